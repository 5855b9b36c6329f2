// patchmap.cu -- device-resident patch map: batches of (ptexFace, s, t) samples -> Osd::PatchCoord records.
//
// Replaces the host loop "for each sample: handle = PatchMap::FindPatch(face, s, t); coords.push_back(
// PatchCoord(*handle, s, t))" (examples/glEvalLimit/particles.cpp:91-115,392-394; far/patchMap.h:180-217)
// that feeds EvalPatches; see patchmap.cuh for the tree layout and the descent.
#include "patchmap.cuh"

#include <algorithm>
#include <new>

using namespace b200osd;

namespace {

struct SampleStreams {
    const int *face;
    const float *s, *t;
    int faceStride, sStride, tStride;     // in elements: 1/1/1 for three packed arrays, 3/3/3 for {face,s,t} records
};

constexpr int kFindBlock = 256;

// One thread per sample, grid-stride over a persistent grid (a few blocks per SM) so that the optional hit count costs
// one atomic per block instead of one per warp.  The 20-byte records of a warp are staged in shared memory (row stride
// 5 words: bank conflict free) and written as five fully coalesced 128-byte rows.
__global__ void __launch_bounds__(kFindBlock) patch_map_find_kernel(PatchMapView m, SampleStreams in, int n,
                                                                     b200osd_patch_coord *out, int *numFound) {
    __shared__ int stage[kFindBlock / 32][32 * 5];
    __shared__ int blockHits;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) blockHits = 0;
    __syncthreads();
    int hits = 0;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i - lane < n; i += step) {
        const long long i0 = i - lane;
        int arrayIndex = -1, patchIndex = 0, vertIndex = 0;
        float s = 0.0f, t = 0.0f;
        if (i < n) {
            const int face = ld_stream_i1(in.face + (size_t)i * in.faceStride);
            s = ld_stream_f1(in.s + (size_t)i * in.sStride);
            t = ld_stream_f1(in.t + (size_t)i * in.tStride);
            const int p = patch_map_find(m, face, s, t);
            if (p >= 0) {
                const int2 h = __ldg(m.handles + p);
                arrayIndex = h.x;
                patchIndex = p;
                vertIndex = h.y;
                ++hits;
            }
        }
        int *st = stage[warp];
        __syncwarp();
        st[lane * 5 + 0] = arrayIndex;
        st[lane * 5 + 1] = patchIndex;
        st[lane * 5 + 2] = vertIndex;
        st[lane * 5 + 3] = __float_as_int(s);
        st[lane * 5 + 4] = __float_as_int(t);
        __syncwarp();
        const int words = (int)min(32LL, (long long)n - i0) * 5;
        int *dst = reinterpret_cast<int *>(out + i0);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int e = q * 32 + lane;
            if (e < words) dst[e] = st[e];
        }
    }
    if (numFound) {
        hits = __reduce_add_sync(0xffffffffu, hits);
        if (lane == 0 && hits) atomicAdd(&blockHits, hits);
        __syncthreads();
        if (threadIdx.x == 0 && blockHits) atomicAdd(numFound, blockHits);
    }
}

}  // namespace

struct b200osd_patch_map {
    PatchMapView view{};
    int4 *d_nodes = nullptr;
    int2 *d_handles = nullptr;
    int numNodes = 0, numHandles = 0;
};

extern "C" {

b200osd_patch_map *b200osd_patch_map_create(int numArrays, const b200osd_patch_array *arrays, int numPatches,
                                            const b200osd_patch_param *params, int patchesAreTriangular) {
    if (numArrays < 0 || numPatches < 0 || (numPatches > 0 && (!arrays || !params))) {
        set_error("patch_map_create: bad arguments");
        return nullptr;
    }
    PatchMapHost host;
    const int brc = build_patch_map(numArrays, arrays, numPatches, params, patchesAreTriangular, &host);
    if (brc) {
        set_error(brc == -1 ? "patch_map_create: the patch arrays do not tile the PatchParam table"
                            : "patch_map_create: two patches cover the same parametric cell");
        return nullptr;
    }
    b200osd_patch_map *m = new (std::nothrow) b200osd_patch_map;
    if (!m) return nullptr;
    m->numNodes = (int)host.nodes.size();
    m->numHandles = (int)host.handles.size();
    auto upload = [&](void **d, const void *h, size_t bytes) -> bool {
        *d = nullptr;
        if (!bytes) return true;
        cudaError_t e = cudaMalloc(d, bytes);
        if (e == cudaSuccess) e = cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { set_error("patch_map_create: %s", cudaGetErrorString(e)); return false; }
        return true;
    };
    if (!upload((void **)&m->d_nodes, host.nodes.data(), host.nodes.size() * sizeof(int4)) ||
        !upload((void **)&m->d_handles, host.handles.data(), host.handles.size() * sizeof(int2))) {
        cudaFree(m->d_nodes);
        cudaFree(m->d_handles);
        delete m;
        return nullptr;
    }
    m->view.nodes = m->d_nodes;
    m->view.handles = m->d_handles;
    m->view.minFace = host.minFace;
    m->view.maxFace = host.maxFace;
    m->view.maxDepth = host.maxDepth;
    m->view.triangular = host.triangular;
    return m;
}

void b200osd_patch_map_destroy(b200osd_patch_map *m) {
    if (!m) return;
    cudaFree(m->d_nodes);
    cudaFree(m->d_handles);
    delete m;
}

int b200osd_patch_map_info(const b200osd_patch_map *m, int info[6]) {
    if (!m || !info) { set_error("patch_map_info: NULL argument"); return B200OSD_ERR_INVALID; }
    info[0] = m->view.minFace;
    info[1] = m->view.maxFace;
    info[2] = m->view.maxDepth;
    info[3] = m->view.triangular;
    info[4] = m->numNodes;
    info[5] = m->numHandles;
    return B200OSD_OK;
}

int b200osd_patch_map_find(const b200osd_patch_map *m, int numSamples, const int *ptexFace, int faceStride,
                           const float *s, int sStride, const float *t, int tStride,
                           b200osd_patch_coord *outCoords, int *numFound, void *stream) {
    if (!m) { set_error("patch_map_find: map is NULL"); return B200OSD_ERR_INVALID; }
    if (numSamples <= 0) return B200OSD_OK;
    if (!ptexFace || !s || !t || !outCoords) { set_error("patch_map_find: NULL sample / output pointer"); return B200OSD_ERR_INVALID; }
    if (faceStride < 1 || sStride < 1 || tStride < 1) { set_error("patch_map_find: strides must be >= 1"); return B200OSD_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    if (numFound) B200_CUDA_TRY(cudaMemsetAsync(numFound, 0, sizeof(int), st));
    SampleStreams in{ptexFace, s, t, faceStride, sStride, tStride};
    const int grid = std::min((numSamples + kFindBlock - 1) / kFindBlock, sm_count() * 8);
    patch_map_find_kernel<<<grid, kFindBlock, 0, st>>>(m->view, in, numSamples, outCoords, numFound);
    return check_launch("patch_map_find_kernel");
}

}  // extern "C"
