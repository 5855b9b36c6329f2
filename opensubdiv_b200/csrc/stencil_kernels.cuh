// stencil_kernels.cuh -- device code of the stencil path (EvalStencils), sm_100a.
//
// Semantics restated from the reference (paths relative to /root/reference/opensubdiv):
//   osd/cpuKernel.cpp:71-240  for row i: out_k[i] = sum_j w_k[off_i + j] * src[idx[off_i + j]]   (k = value,du,dv,duu,duv,dvv)
//   osd/cudaKernel.cu:85-98   row i of [start,end) is written to element i of dst (absolute index)
//
// Two kernel families:
//   csr_*   : plain reference layout (sizes/offsets/indices/weights), one thread per row.  Serves the raw
//             pointer API (any table a client hands us) -- correct for everything, not the fast path.
//   sell_*  : the B200StencilTable layout.  Rows are sorted by size inside windows and cut into slices of
//             32 rows; a slice stores its elements element-major in groups of 4 (lane l owns row l of the
//             slice and finds its elements j=4g..4g+3 in one int4 / float4 at [base + g*32 + l]), so every
//             warp-wide load of the index / weight streams is one fully coalesced 512-byte request and no
//             shuffle or shared-memory reduction is needed: each lane accumulates its own row in registers.
//             Primvar gathers go through L1/L2 (the control-vertex set is small and hot).
#pragma once

#include "common.cuh"

namespace b200osd {

constexpr int kMaxOut = 6;
constexpr int kSliceRows = 32;     // rows per slice == warp size
constexpr int kVec = 4;            // elements per 128-bit load

struct StencilIO {
    const float *src;              // already offset by srcDesc.offset
    int srcStride;                 // floats between vertices
    int L;                         // primvar length
    float *dst[kMaxOut];           // already offset; NULL = skip
    int dstStride[kMaxOut];
    int dstVec[kMaxOut];           // widest aligned store usable for this output: 1, 2 or 4 floats
    int start, end;                // absolute row range
};

struct CsrTable {
    const int *sizes, *offsets, *indices;
    const float *w[kMaxOut];
};

struct SellTable {
    const uint2 *ipool;            // index pool in 8-byte units: per slice either 4 x 16-bit offsets (one uint2) or 4 x 32-bit
                                   // indices (one int4 = two units) per lane slot, element-major like the weights
    const float4 *w4[kMaxOut];     // [totalVec] weight groups, one array per weight stream
    const int4 *meta;              // per slice: {baseVec (weight slots), lenVec, indexBase (>= 0: 16-bit offsets from it,
                                   //             -1: 32-bit indices), first unit of the slice in ipool}
    const int *rows;               // [numSlices*32] original row of each lane slot, -1 = padding
    int sliceBegin, sliceEnd;
};

// ------------------------------------------------------------------------------- vertex access --
template <int L>
__device__ __forceinline__ void load_vertex_scalar(const float *src, int stride, int idx, float (&v)[L]) {
    const float *p = src + (size_t)idx * (size_t)stride;
#pragma unroll
    for (int k = 0; k < L; ++k) v[k] = p[k];
}

// stride is a multiple of 4 floats and src is 16-byte aligned
template <int L>
__device__ __forceinline__ void load_vertex_vec4(const float *src, int stride, int idx, float (&v)[L]) {
    const float4 *p = reinterpret_cast<const float4 *>(src + (size_t)idx * (size_t)stride);
#pragma unroll
    for (int c = 0; c < (L + 3) / 4; ++c) {
        float4 t = p[c];
        if (4 * c + 0 < L) v[4 * c + 0] = t.x;
        if (4 * c + 1 < L) v[4 * c + 1] = t.y;
        if (4 * c + 2 < L) v[4 * c + 2] = t.z;
        if (4 * c + 3 < L) v[4 * c + 3] = t.w;
    }
}

// stride is even and src is 8-byte aligned (e.g. 6-float xyz+normal vertices): L/2 64-bit loads
template <int L>
__device__ __forceinline__ void load_vertex_vec2(const float *src, int stride, int idx, float (&v)[L]) {
    const float2 *p = reinterpret_cast<const float2 *>(src + (size_t)idx * (size_t)stride);
#pragma unroll
    for (int c = 0; c < (L + 1) / 2; ++c) {
        float2 t = p[c];
        if (2 * c + 0 < L) v[2 * c + 0] = t.x;
        if (2 * c + 1 < L) v[2 * c + 1] = t.y;
    }
}

enum { SRC_SCALAR = 0, SRC_VEC2 = 1, SRC_VEC4 = 2 };

template <int L, int SRCMODE>
__device__ __forceinline__ void load_vertex(const float *src, int stride, int idx, float (&v)[L]) {
    if (SRCMODE == SRC_VEC4)      load_vertex_vec4<L>(src, stride, idx, v);
    else if (SRCMODE == SRC_VEC2) load_vertex_vec2<L>(src, stride, idx, v);
    else                          load_vertex_scalar<L>(src, stride, idx, v);
}

template <int L>
__device__ __forceinline__ void store_vertex(float *p, const float (&v)[L], int vec) {
    if ((L % 4 == 0) && vec == 4) {
#pragma unroll
        for (int k = 0; k < L / 4; ++k) st_stream_f4(p + 4 * k, v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    } else if ((L % 2 == 0) && vec >= 2) {
#pragma unroll
        for (int k = 0; k < L / 2; ++k) st_stream_f2(p + 2 * k, v[2 * k], v[2 * k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < L; ++k) st_stream_f1(p + k, v[k]);
    }
}

template <int L, int K>
__device__ __forceinline__ void store_row(const StencilIO &io, int row, const float (&acc)[K][L]) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float *d = io.dst[k];
        if (d) store_vertex<L>(d + (size_t)row * (size_t)io.dstStride[k], acc[k], io.dstVec[k]);
    }
}

template <int L, int K>
__device__ __forceinline__ void accumulate(float (&acc)[K][L], const float (&v)[L], const float (&w)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int c = 0; c < L; ++c) acc[k][c] = fmaf(w[k], v[c], acc[k][c]);
}

// The arithmetic of the reference's CPU kernels, operation for operation: a rounded product added to the running sum
// (osd/cpuKernel.cpp:71-240 compiled for x86-64 contracts nothing), in the table's element order.  With it a row is
// bit-identical to Osd::CpuEvaluator / OmpEvaluator whatever its length -- the fused multiply-add of the fast kernels is the
// more accurate operation, but two correct fp32 summations of n terms may differ by ~n * 2^-23 of sum|w||x|.
template <int L, int K>
__device__ __forceinline__ void accumulate_exact(float (&acc)[K][L], const float (&v)[L], const float (&w)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int c = 0; c < L; ++c) acc[k][c] = __fadd_rn(acc[k][c], __fmul_rn(w[k], v[c]));
}

// ------------------------------------------------------------------------------------ CSR path --
template <int L, int K, int SRCMODE, bool EXACT>
__global__ void __launch_bounds__(128) csr_kernel(StencilIO io, CsrTable t) {
    int row = io.start + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= io.end) return;
    int off = t.offsets[row];
    int n = t.sizes[row];
    float acc[K][L];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int c = 0; c < L; ++c) acc[k][c] = 0.0f;
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
        int idx = t.indices[off + j];
        float w[K];
#pragma unroll
        for (int k = 0; k < K; ++k) w[k] = t.w[k][off + j];
        float v[L];
        load_vertex<L, SRCMODE>(io.src, io.srcStride, idx, v);
        if (EXACT) accumulate_exact<L, K>(acc, v, w);
        else accumulate<L, K>(acc, v, w);
    }
    store_row<L, K>(io, row, acc);
}

// Any primvar length: components are processed in tiles of 4, re-walking the row per tile.
template <int K, bool EXACT>
__global__ void __launch_bounds__(128) csr_kernel_anyL(StencilIO io, CsrTable t) {
    int row = io.start + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= io.end) return;
    int off = t.offsets[row];
    int n = t.sizes[row];
    for (int c0 = 0; c0 < io.L; c0 += 4) {
        int nc = min(4, io.L - c0);
        float acc[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[k][c] = 0.0f;
        for (int j = 0; j < n; ++j) {
            int idx = t.indices[off + j];
            const float *p = io.src + (size_t)idx * (size_t)io.srcStride + c0;
            float v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = (c < nc) ? p[c] : 0.0f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float w = t.w[k][off + j];
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[k][c] = EXACT ? __fadd_rn(acc[k][c], __fmul_rn(w, v[c])) : fmaf(w, v[c], acc[k][c]);
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float *d = io.dst[k];
            if (!d) continue;
            d += (size_t)row * (size_t)io.dstStride[k] + c0;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < nc) d[c] = acc[k][c];
        }
    }
}

// ----------------------------------------------------------------------------------- SELL path --
// One lane per row.  UNROLL index groups are issued back to back so that each lane has 2*UNROLL 128-bit stream
// loads in flight before the first gather is consumed.
// One index group of a lane: 4 control-vertex indices.  `slot` = g*32 + lane inside the slice.  The encoding is a
// per-slice property (warp-uniform branch): 16-bit offsets from the slice's smallest index when the slice spans fewer
// than 65536 control vertices (practically always on refined meshes), plain 32-bit indices otherwise.
__device__ __forceinline__ int4 load_index_group(const uint2 *slicePool, int slot, int indexBase) {
    if (indexBase >= 0) {
        const uint2 q = ld_stream_u2(slicePool + slot);
        return make_int4(indexBase + (int)(q.x & 0xffffu), indexBase + (int)(q.x >> 16),
                         indexBase + (int)(q.y & 0xffffu), indexBase + (int)(q.y >> 16));
    }
    return ld_stream_i4(reinterpret_cast<const int4 *>(slicePool) + slot);
}

template <int L, int K, int SRCMODE, int UNROLL>
__device__ __forceinline__ void sell_slice(const StencilIO &io, const SellTable &t, const int4 m, const int lane,
                                           float (&acc)[K][L]) {
    const size_t base = (size_t)(unsigned)m.x + lane;
    const uint2 *slicePool = t.ipool + (size_t)(unsigned)m.w;
    const float4 *wp[K];
#pragma unroll
    for (int k = 0; k < K; ++k) wp[k] = t.w4[k] + base;
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int c = 0; c < L; ++c) acc[k][c] = 0.0f;

    int g = 0;
    const int ng = m.y;
    for (; g + UNROLL <= ng; g += UNROLL) {
        int4 id[UNROLL];
        float4 w[UNROLL][K];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            id[u] = load_index_group(slicePool, (g + u) * kSliceRows + lane, m.z);
#pragma unroll
            for (int k = 0; k < K; ++k) w[u][k] = ld_stream_f4(wp[k] + (size_t)(g + u) * kSliceRows);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            float v0[L], v1[L], v2[L], v3[L];
            load_vertex<L, SRCMODE>(io.src, io.srcStride, id[u].x, v0);
            load_vertex<L, SRCMODE>(io.src, io.srcStride, id[u].y, v1);
            load_vertex<L, SRCMODE>(io.src, io.srcStride, id[u].z, v2);
            load_vertex<L, SRCMODE>(io.src, io.srcStride, id[u].w, v3);
            float wx[K], wy[K], wz[K], ww[K];
#pragma unroll
            for (int k = 0; k < K; ++k) { wx[k] = w[u][k].x; wy[k] = w[u][k].y; wz[k] = w[u][k].z; ww[k] = w[u][k].w; }
            accumulate<L, K>(acc, v0, wx);
            accumulate<L, K>(acc, v1, wy);
            accumulate<L, K>(acc, v2, wz);
            accumulate<L, K>(acc, v3, ww);
        }
    }
    for (; g < ng; ++g) {
        int4 id = load_index_group(slicePool, g * kSliceRows + lane, m.z);
        float4 w[K];
#pragma unroll
        for (int k = 0; k < K; ++k) w[k] = ld_stream_f4(wp[k] + (size_t)g * kSliceRows);
        float v0[L], v1[L], v2[L], v3[L];
        load_vertex<L, SRCMODE>(io.src, io.srcStride, id.x, v0);
        load_vertex<L, SRCMODE>(io.src, io.srcStride, id.y, v1);
        load_vertex<L, SRCMODE>(io.src, io.srcStride, id.z, v2);
        load_vertex<L, SRCMODE>(io.src, io.srcStride, id.w, v3);
        float wx[K], wy[K], wz[K], ww[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { wx[k] = w[k].x; wy[k] = w[k].y; wz[k] = w[k].z; ww[k] = w[k].w; }
        accumulate<L, K>(acc, v0, wx);
        accumulate<L, K>(acc, v1, wy);
        accumulate<L, K>(acc, v2, wz);
        accumulate<L, K>(acc, v3, ww);
    }
}

// One warp per slice (one-shot grid).  MINB = minimum resident blocks per SM asked of the register allocator.
template <int L, int K, int SRCMODE, int UNROLL, int MINB>
__global__ void __launch_bounds__(256, MINB) sell_kernel(StencilIO io, SellTable t) {
    const int lane = threadIdx.x & 31;
    const int slice = t.sliceBegin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (slice >= t.sliceEnd) return;
    const int4 m = t.meta[slice];
    const int row = t.rows[(size_t)slice * kSliceRows + lane];
    float acc[K][L];
    sell_slice<L, K, SRCMODE, UNROLL>(io, t, m, lane, acc);
    if (row >= io.start && row < io.end) store_row<L, K>(io, row, acc);
}

// Persistent form: the grid is sized to the machine and every warp walks slices with a grid stride, fetching the
// NEXT slice's descriptor and row ids before working on the current one, so the descriptor -> stream dependent
// DRAM round trip is paid once per warp instead of once per slice.
template <int L, int K, int SRCMODE, int UNROLL, int MINB>
__global__ void __launch_bounds__(256, MINB) sell_kernel_persist(StencilIO io, SellTable t) {
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    int slice = t.sliceBegin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (slice >= t.sliceEnd) return;
    int4 m = t.meta[slice];
    int row = t.rows[(size_t)slice * kSliceRows + lane];
    for (;;) {
        const int next = slice + nwarps;
        const bool more = next < t.sliceEnd;
        int4 mn = make_int4(0, 0, 0, 0);
        int rown = -1;
        if (more) {
            mn = t.meta[next];
            rown = t.rows[(size_t)next * kSliceRows + lane];
        }
        float acc[K][L];
        sell_slice<L, K, SRCMODE, UNROLL>(io, t, m, lane, acc);
        if (row >= io.start && row < io.end) store_row<L, K>(io, row, acc);
        if (!more) break;
        slice = next;
        m = mn;
        row = rown;
    }
}

// ------------------------------------------------------------------------- TMA-staged SELL path --
// Same layout, same arithmetic (element order, FMA) as sell_kernel -- results are bit-identical -- but the index and
// weight streams never pass through the load/store unit's register path: every warp owns a ring of shared-memory
// stages and one elected lane streams the next chunks of its slices into it with cp.async.bulk (the 1-D TMA copy;
// SASS: UBLKCP) completing on an mbarrier, C groups (= 4C elements of 32 rows and all K weight streams) per copy set.  The lanes then read their 8/16-byte index groups and 16-byte weight groups from shared memory (conflict-free:
// consecutive lanes, consecutive words).  What this buys over sell_kernel:
//   * the DRAM round trip of the streams is taken by the copy engine `stages-1` chunks ahead, so a warp never waits on
//     HBM -- only the control-vertex gathers (L1/L2 hits) are on its critical path;
//   * no registers hold in-flight stream data, which is what capped the K = 3 / 6 kernels at 3 blocks per SM;
//   * the grid is persistent (one wave), slices are walked with a grid stride and their descriptors / row ids are
//     fetched one slice ahead.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar, unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// bytes of one ring stage: C groups x 32 lanes x (16 B of 32-bit indices at most + K x 16 B of weights)
template <int K, int C> __host__ __device__ constexpr unsigned tma_stage_bytes() { return C * 32u * 16u * (1u + K); }
constexpr unsigned kTmaBarrierBytes = 128;          // the stage barriers of one warp (8 B each), kept on their own line

struct SliceCursor {           // position in a warp's sequence of chunks: slices slice, slice+nwarps, ... x groups g0, g0+kChunk, ...
    int slice, g0;
    int4 m, mNext;             // descriptor of `slice` and of the next slice of this warp (fetched one slice ahead)
};

// C = groups (of 4 elements) per chunk, MINB = resident blocks per SM asked of the register allocator
template <int L, int K, int SRCMODE, int C, int MINB>
__global__ void __launch_bounds__(256, MINB) sell_tma_kernel(StencilIO io, SellTable t, int stages) {
    extern __shared__ __align__(128) unsigned char tma_smem[];
    constexpr unsigned stageBytes = tma_stage_bytes<K, C>();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int nwarps = gridDim.x * W;
    const int first = t.sliceBegin + blockIdx.x * W + warp;
    if (first >= t.sliceEnd) return;                                     // whole warp: nothing to do
    unsigned char *ring = tma_smem + (size_t)warp * ((size_t)stages * stageBytes + kTmaBarrierBytes);
    const unsigned ringAddr = smem_u32(ring);
    const unsigned barAddr = ringAddr + (unsigned)stages * stageBytes;
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(barAddr + 8u * s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const unsigned long long policy = l2_evict_first_policy();

    auto start_cursor = [&](SliceCursor &c) {
        c.slice = first;
        c.g0 = 0;
        c.m = t.meta[first];
        c.mNext = (first + nwarps < t.sliceEnd) ? t.meta[first + nwarps] : make_int4(0, 0, 0, 0);
    };
    auto advance = [&](SliceCursor &c) -> bool {                          // true: moved on to the next slice
        c.g0 += C;
        if (c.g0 < c.m.y) return false;
        c.slice += nwarps;
        c.g0 = 0;
        c.m = c.mNext;
        if (c.slice + nwarps < t.sliceEnd) c.mNext = t.meta[c.slice + nwarps];
        return true;
    };
    // the elected lane streams chunk (c.slice, c.g0 ..) into ring stage s
    auto issue = [&](const SliceCursor &c, int s) {
        if (lane != 0) return;
        const int ng = min(C, c.m.y - c.g0);
        const bool is16 = c.m.z >= 0;
        const unsigned idxBytes = (unsigned)ng * 32u * (is16 ? 8u : 16u), wBytes = (unsigned)ng * 512u;
        const unsigned bar = barAddr + 8u * s, dst = ringAddr + (unsigned)s * stageBytes;
        // generic-proxy reads of this stage (made visible to this lane by the warp barrier) precede the async-proxy writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, idxBytes + (unsigned)K * wBytes);
        if (ng > 0) {
            const uint2 *ip = t.ipool + (size_t)(unsigned)c.m.w + (size_t)c.g0 * 32u * (is16 ? 1u : 2u);
            bulk_g2s(dst, ip, idxBytes, bar, policy);
#pragma unroll
            for (int k = 0; k < K; ++k)
                bulk_g2s(dst + C * 512u * (1u + k), t.w4[k] + (size_t)(unsigned)c.m.x + (size_t)c.g0 * 32u, wBytes, bar, policy);
        }
    };

    SliceCursor p, c;
    start_cursor(p);
    c = p;
    int pStage = 0, cStage = 0;
    unsigned phase = 0;                                                  // bit s: parity the consumer waits for on stage s
    for (int a = 0; a < stages - 1 && p.slice < t.sliceEnd; ++a) {
        issue(p, pStage);
        advance(p);
        pStage = (pStage + 1 == stages) ? 0 : pStage + 1;
    }
    int row = t.rows[(size_t)c.slice * kSliceRows + lane];
    int rowNext = (c.slice + nwarps < t.sliceEnd) ? t.rows[(size_t)(c.slice + nwarps) * kSliceRows + lane] : -1;
    float acc[K][L];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int q = 0; q < L; ++q) acc[k][q] = 0.0f;

    while (c.slice < t.sliceEnd) {
        // refill the stage consumed in the previous iteration (the warp barrier at its end ordered the reads)
        if (p.slice < t.sliceEnd) {
            issue(p, pStage);
            advance(p);
            pStage = (pStage + 1 == stages) ? 0 : pStage + 1;
        }
        const unsigned bar = barAddr + 8u * cStage;
        while (!mbar_try_wait(bar, (phase >> cStage) & 1u)) { }
        phase ^= 1u << cStage;

        const unsigned char *stage = ring + (size_t)cStage * stageBytes;
        const int ng = min(C, c.m.y - c.g0);
        const int indexBase = c.m.z;
#pragma unroll
        for (int gl = 0; gl < C; ++gl) {
            if (gl < ng) {
                int4 id;
                if (indexBase >= 0) {
                    const uint2 q = *reinterpret_cast<const uint2 *>(stage + ((size_t)gl * 32 + lane) * 8);
                    id = make_int4(indexBase + (int)(q.x & 0xffffu), indexBase + (int)(q.x >> 16),
                                   indexBase + (int)(q.y & 0xffffu), indexBase + (int)(q.y >> 16));
                } else {
                    id = *reinterpret_cast<const int4 *>(stage + ((size_t)gl * 32 + lane) * 16);
                }
                float4 w[K];
#pragma unroll
                for (int k = 0; k < K; ++k)
                    w[k] = *reinterpret_cast<const float4 *>(stage + C * 512u * (1u + k) + ((size_t)gl * 32 + lane) * 16);
                float v0[L], v1[L], v2[L], v3[L];
                load_vertex<L, SRCMODE>(io.src, io.srcStride, id.x, v0);
                load_vertex<L, SRCMODE>(io.src, io.srcStride, id.y, v1);
                load_vertex<L, SRCMODE>(io.src, io.srcStride, id.z, v2);
                load_vertex<L, SRCMODE>(io.src, io.srcStride, id.w, v3);
                float wx[K], wy[K], wz[K], ww[K];
#pragma unroll
                for (int k = 0; k < K; ++k) { wx[k] = w[k].x; wy[k] = w[k].y; wz[k] = w[k].z; ww[k] = w[k].w; }
                accumulate<L, K>(acc, v0, wx);
                accumulate<L, K>(acc, v1, wy);
                accumulate<L, K>(acc, v2, wz);
                accumulate<L, K>(acc, v3, ww);
            }
        }
        if (advance(c)) {                                                // that was the slice's last chunk
            if (row >= io.start && row < io.end) store_row<L, K>(io, row, acc);
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int q = 0; q < L; ++q) acc[k][q] = 0.0f;
            row = rowNext;
            rowNext = (c.slice + nwarps < t.sliceEnd) ? t.rows[(size_t)(c.slice + nwarps) * kSliceRows + lane] : -1;
        }
        cStage = (cStage + 1 == stages) ? 0 : cStage + 1;
        __syncwarp();
    }
}

// Batched instances sharing one topology (the reference's pattern is one EvalStencils call per instance
// with shifted descriptors, examples/glShareTopology/meshRefiner.h:68-88).  Instance b reads its control vertices at
// src + b*srcInst and writes its rows at dst + b*dstInst; the index / weight streams of a slice are read ONCE for the B
// instances of a chunk, so table traffic per instance drops B-fold.  Value stencils only (K = 1); per instance the
// arithmetic (order, FMA) is identical to sell_kernel, so results are bit-identical to per-instance calls.
template <int L, int B, int SRCMODE>
__global__ void __launch_bounds__(256) sell_kernel_batched(StencilIO io, SellTable t, long long srcInst, long long dstInst) {
    const int lane = threadIdx.x & 31;
    const int slice = t.sliceBegin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (slice >= t.sliceEnd) return;
    const int4 m = t.meta[slice];
    const int row = t.rows[(size_t)slice * kSliceRows + lane];
    const size_t base = (size_t)(unsigned)m.x + lane;
    const uint2 *slicePool = t.ipool + (size_t)(unsigned)m.w;
    const float4 *wp = t.w4[0] + base;
    float acc[B][L];
#pragma unroll
    for (int b = 0; b < B; ++b)
#pragma unroll
        for (int c = 0; c < L; ++c) acc[b][c] = 0.0f;
    for (int g = 0; g < m.y; ++g) {
        const int4 id = load_index_group(slicePool, g * kSliceRows + lane, m.z);
        const float4 w = ld_stream_f4(wp + (size_t)g * kSliceRows);
        const int ids[4] = { id.x, id.y, id.z, id.w };
        const float ws[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int b = 0; b < B; ++b) {
                float v[L];
                load_vertex<L, SRCMODE>(io.src + (size_t)b * (size_t)srcInst, io.srcStride, ids[q], v);
#pragma unroll
                for (int c = 0; c < L; ++c) acc[b][c] = fmaf(ws[q], v[c], acc[b][c]);
            }
        }
    }
    if (row >= io.start && row < io.end && io.dst[0]) {
#pragma unroll
        for (int b = 0; b < B; ++b)
            store_vertex<L>(io.dst[0] + (size_t)b * (size_t)dstInst + (size_t)row * (size_t)io.dstStride[0], acc[b], io.dstVec[0]);
    }
}

template <int K>
__global__ void __launch_bounds__(256) sell_kernel_anyL(StencilIO io, SellTable t) {
    const int lane = threadIdx.x & 31;
    const int slice = t.sliceBegin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (slice >= t.sliceEnd) return;
    const int4 m = t.meta[slice];
    const int row = t.rows[(size_t)slice * kSliceRows + lane];
    const size_t base = (size_t)(unsigned)m.x + lane;
    const bool live = (row >= io.start && row < io.end);
    for (int c0 = 0; c0 < io.L; c0 += 4) {
        int nc = min(4, io.L - c0);
        float acc[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[k][c] = 0.0f;
        for (int g = 0; g < m.y; ++g) {
            const int4 id = load_index_group(t.ipool + (size_t)(unsigned)m.w, g * kSliceRows + lane, m.z);
            int ids[4] = { id.x, id.y, id.z, id.w };
            float4 w[K];
#pragma unroll
            for (int k = 0; k < K; ++k) w[k] = t.w4[k][base + (size_t)g * kSliceRows];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float *p = io.src + (size_t)ids[q] * (size_t)io.srcStride + c0;
                float v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) v[c] = (c < nc) ? p[c] : 0.0f;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float wk = q == 0 ? w[k].x : (q == 1 ? w[k].y : (q == 2 ? w[k].z : w[k].w));
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[k][c] = fmaf(wk, v[c], acc[k][c]);
                }
            }
        }
        if (live) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float *d = io.dst[k];
                if (!d) continue;
                d += (size_t)row * (size_t)io.dstStride[k] + c0;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < nc) d[c] = acc[k][c];
            }
        }
    }
}

}  // namespace b200osd
