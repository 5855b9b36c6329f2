// common.cuh -- shared host/device helpers for libb200osd (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdint>

#include "../../include/b200osd_capi.h"

namespace b200osd {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;

#define B200_CUDA_TRY(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::b200osd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                                 __FILE__, __LINE__);                                         \
            return B200OSD_ERR_CUDA;                                                          \
        }                                                                                     \
    } while (0)

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return B200OSD_ERR_CUDA;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return B200OSD_OK;
}

// Number of SMs of the current device (148 on B200); cached.
int sm_count();

// Wraps reference-layout DEVICE arrays (built on the device, e.g. by the limit-stencil builder) into a stencil table that
// owns them; unless flags bit 0 is set the arrays are also read back once to build the bucketed layout (stencil.cu).
// On failure the arrays are freed and NULL is returned.
struct AdoptedArrays {
    int numStencils, numControlVertices, numW;
    long long numElements;
    int *sizes, *offsets, *indices;
    float *w[6];
};
::b200osd_stencil_table *adopt_device_table(const AdoptedArrays &a, int flags);

// ---- streaming loads/stores -------------------------------------------------------------------
// Table streams (indices / weights / coords) are read exactly once per launch: bypass L1 allocation
// so the L1 stays available for the primvar gathers.  Outputs are written once: st.global.cs (evict-first).
#ifdef __CUDACC__
__device__ __forceinline__ int4 ld_stream_i4(const int4 *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const uint2 *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ int ld_stream_i1(const int *p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned ld_stream_u1(const unsigned *p) {
    unsigned r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f1(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f1(float *p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream_f2(float *p, float a, float b) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_stream_f4(float *p, float a, float b, float c, float d) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
#endif

}  // namespace b200osd
