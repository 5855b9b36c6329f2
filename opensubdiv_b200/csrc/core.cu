// core.cu -- library plumbing: error string, launch counter, synchronisation and the device vertex buffer
// (the CudaVertexBuffer equivalent, /root/reference/opensubdiv/osd/cudaVertexBuffer.cpp:35-93).
#include "common.cuh"

#include <cstring>
#include <new>

namespace b200osd {

static thread_local char t_error[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = 0;
    if (!cached) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}

}  // namespace b200osd

using namespace b200osd;

struct b200osd_vertex_buffer {
    int numElements = 0;
    int numVertices = 0;
    float *d = nullptr;
};

extern "C" {

const char *b200osd_version(void) { return "b200osd 0.1 (sm_100a; stencils: csr+sell; patches: separable bspline/gregory/linear/box-spline)"; }
const char *b200osd_last_error(void) { return t_error; }
long long b200osd_launch_count(void) { return g_launches.load(); }
void b200osd_reset_launch_count(void) { g_launches.store(0); }

int b200osd_synchronize(void *stream) {
    if (stream) B200_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    else        B200_CUDA_TRY(cudaDeviceSynchronize());
    return B200OSD_OK;
}

b200osd_vertex_buffer *b200osd_vertex_buffer_create(int numElements, int numVertices) {
    if (numElements <= 0 || numVertices < 0) { set_error("vertex_buffer_create: bad dimensions"); return nullptr; }
    b200osd_vertex_buffer *vb = new (std::nothrow) b200osd_vertex_buffer;
    if (!vb) return nullptr;
    vb->numElements = numElements;
    vb->numVertices = numVertices;
    // size_t arithmetic: the reference computes the byte size in int and overflows above 2 GiB (cudaVertexBuffer.cpp:85)
    const size_t bytes = (size_t)numElements * (size_t)numVertices * sizeof(float);
    cudaError_t e = cudaMalloc((void **)&vb->d, bytes ? bytes : sizeof(float));
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        delete vb;
        return nullptr;       // reference: Create() returns NULL (cudaVertexBuffer.cpp:46-53)
    }
    return vb;
}

void b200osd_vertex_buffer_destroy(b200osd_vertex_buffer *vb) {
    if (!vb) return;
    cudaFree(vb->d);
    delete vb;
}

int b200osd_vertex_buffer_num_elements(const b200osd_vertex_buffer *vb) { return vb ? vb->numElements : 0; }
int b200osd_vertex_buffer_num_vertices(const b200osd_vertex_buffer *vb) { return vb ? vb->numVertices : 0; }
float *b200osd_vertex_buffer_bind(b200osd_vertex_buffer *vb) { return vb ? vb->d : nullptr; }

static int check_range(const b200osd_vertex_buffer *vb, const void *host, int startVertex, int numVertices) {
    if (!vb || !host) { set_error("vertex buffer / host pointer is NULL"); return B200OSD_ERR_INVALID; }
    if (startVertex < 0 || numVertices < 0 || (long long)startVertex + numVertices > vb->numVertices) {
        set_error("vertex range [%d,+%d) outside buffer of %d vertices", startVertex, numVertices, vb->numVertices);
        return B200OSD_ERR_INVALID;
    }
    return B200OSD_OK;
}

int b200osd_vertex_buffer_update(b200osd_vertex_buffer *vb, const float *hostSrc, int startVertex, int numVertices, void *stream) {
    int rc = check_range(vb, hostSrc, startVertex, numVertices);
    if (rc) return rc;
    const size_t bytes = (size_t)vb->numElements * (size_t)numVertices * sizeof(float);
    float *dst = vb->d + (size_t)vb->numElements * (size_t)startVertex;
    if (stream) B200_CUDA_TRY(cudaMemcpyAsync(dst, hostSrc, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    else        B200_CUDA_TRY(cudaMemcpy(dst, hostSrc, bytes, cudaMemcpyHostToDevice));
    return B200OSD_OK;
}

int b200osd_vertex_buffer_read(b200osd_vertex_buffer *vb, float *hostDst, int startVertex, int numVertices, void *stream) {
    int rc = check_range(vb, hostDst, startVertex, numVertices);
    if (rc) return rc;
    const size_t bytes = (size_t)vb->numElements * (size_t)numVertices * sizeof(float);
    const float *src = vb->d + (size_t)vb->numElements * (size_t)startVertex;
    if (stream) B200_CUDA_TRY(cudaMemcpyAsync(hostDst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    else        B200_CUDA_TRY(cudaMemcpy(hostDst, src, bytes, cudaMemcpyDeviceToHost));
    return B200OSD_OK;
}

}  // extern "C"
