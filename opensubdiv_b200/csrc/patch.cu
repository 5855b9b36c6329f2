// patch.cu -- host side of the limit-evaluation path: argument validation mirroring the reference
// evaluator, component tiling, device-side binning of the coordinates, kernel dispatch, and the device
// patch-table container.
//
// Reference behaviour mirrored (paths relative to /root/reference/opensubdiv):
//   osd/cpuEvaluator.cpp:165-176,224-241,300-331  src NULL -> false; value-only form with dst NULL -> false;
//                                                 a non-NULL output whose length != srcDesc.length -> false
//   osd/cudaKernel.cu:300-327                     NULL derivative outputs are skipped
//   osd/cudaPatchTable.cpp:69-162                 device copies of PatchArray[], indices, PatchParam[] (+varying, fvar)
//   osd/mesh.h:305-409, osd/glComputeEvaluator.h:98-128   "instantiatable" evaluators own per-use cached state: here
//                                                 the patch plan (a cached binning of one coordinate set)
//
// No entry point below allocates with cudaMalloc, frees, or synchronises the device: tables are immutable after
// _set, plans own their scratch from _create on, and the automatic binning of b200osd_patch_table_eval takes its
// scratch from a stream-ordered memory pool (cudaMallocFromPoolAsync / cudaFreeAsync: asynchronous, legal inside a
// stream capture, reused from the pool after the first call).
#include "patch_kernels.cuh"
#include "limit_kernels.cuh"      // same translation unit: the triangle bases read this unit's __constant__ tables

#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

using namespace b200osd;

namespace {

constexpr int kMaxDevices = 64;

// Stream-ordered scratch: one pool per device that never trims, so that after the first call an allocation is a
// pointer bump on the stream -- no cudaMalloc, no synchronisation, legal during stream capture.
int scratch_pool(cudaMemPool_t *out) {
    static cudaMemPool_t pools[kMaxDevices];
    static std::atomic<bool> ready[kMaxDevices];
    static std::mutex mu;
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) { set_error("device ordinal %d out of range", dev); return B200OSD_ERR_UNSUPPORTED; }
    if (!ready[dev].load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lock(mu);
        if (!ready[dev].load(std::memory_order_relaxed)) {
            cudaMemPoolProps props;
            std::memset(&props, 0, sizeof(props));
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            B200_CUDA_TRY(cudaMemPoolCreate(&pools[dev], &props));
            unsigned long long keep = ~0ULL;
            B200_CUDA_TRY(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
            ready[dev].store(true, std::memory_order_release);
        }
    }
    *out = pools[dev];
    return B200OSD_OK;
}

// ------------------------------------------------------------------------------------ binning --
struct BinScratch {              // carved out of one allocation, every region 256-byte aligned
    BinState *state = nullptr;   // [1]            } zeroed together
    int *count = nullptr;        // [numBins]      }
    int *start = nullptr;        // [numBins]
    int *blockSum = nullptr;     // [scanBlocks]
    int2 *keyRank = nullptr;     // [n]
    int *perm = nullptr;         // [n]
    size_t zeroBytes = 0;
};

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

size_t bin_scratch_bytes(int maxCoords, int numPatches) {
    const size_t bins = (size_t)numPatches + 1, blocks = (bins + kScanTile - 1) / kScanTile;
    return align256(sizeof(BinState)) + align256(bins * 4) + align256(bins * 4) + align256(blocks * 4)
         + align256((size_t)maxCoords * 8) + align256((size_t)maxCoords * 4);
}

BinScratch carve_bin_scratch(void *base, int maxCoords, int numPatches) {
    const size_t bins = (size_t)numPatches + 1, blocks = (bins + kScanTile - 1) / kScanTile;
    char *p = static_cast<char *>(base);
    BinScratch s;
    s.state = reinterpret_cast<BinState *>(p);       p += align256(sizeof(BinState));
    s.count = reinterpret_cast<int *>(p);            p += align256(bins * 4);
    s.zeroBytes = (size_t)(p - static_cast<char *>(base));
    s.start = reinterpret_cast<int *>(p);            p += align256(bins * 4);
    s.blockSum = reinterpret_cast<int *>(p);         p += align256(blocks * 4);
    s.keyRank = reinterpret_cast<int2 *>(p);         p += align256((size_t)maxCoords * 8);
    s.perm = reinterpret_cast<int *>(p);
    return s;
}

// probe (unless forced) -> count -> scan -> scatter, all on `st`; every kernel after the probe returns at once when the
// probe found the coordinates already coherent
int run_binning(const BinScratch &s, const b200osd_patch_coord *coords, int n, int numPatches, bool force, cudaStream_t st) {
    const int bins = numPatches + 1, blocks = (bins + kScanTile - 1) / kScanTile;
    B200_CUDA_TRY(cudaMemsetAsync(s.state, 0, s.zeroBytes, st));
    bin_probe_kernel<<<kBinProbeWarps / 4, 128, 0, st>>>(coords, n, s.state, kPatchModeGrouped, force ? 1 : 0);
    int rc = check_launch("bin_probe_kernel");
    if (rc) return rc;
    bin_count_kernel<<<(n + 255) / 256, 256, 0, st>>>(coords, n, numPatches, s.state, s.count, s.keyRank);
    if ((rc = check_launch("bin_count_kernel"))) return rc;
    bin_scan_local_kernel<<<blocks, kScanThreads, 0, st>>>(s.state, s.count, s.start, s.blockSum, bins);
    if ((rc = check_launch("bin_scan_local_kernel"))) return rc;
    bin_scan_top_kernel<<<1, kScanThreads, 0, st>>>(s.state, s.blockSum, blocks);
    if ((rc = check_launch("bin_scan_top_kernel"))) return rc;
    bin_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(s.state, s.keyRank, s.start, s.blockSum, n, s.perm);
    return check_launch("bin_scatter_kernel");
}

// ----------------------------------------------------------------------------------- dispatch --
int patch_grid(long long n) {
    const long long need = (n + kPatchBlock - 1) / kPatchBlock;
    return (int)std::min<long long>(need, (long long)sm_count() * 16);       // persistent: at most 64 warps per SM
}

template <int ORDER, bool TRI>
int launch_run(const PatchIO &io, int LT, cudaStream_t st) {
    const int grid = patch_grid(io.n);
    const size_t smem = (size_t)(kPatchBlock / 32) * (size_t)io.warpWords * sizeof(float);
    switch (LT) {
        case 1: patch_run_kernel<1, ORDER, TRI><<<grid, kPatchBlock, smem, st>>>(io); break;
        case 2: patch_run_kernel<2, ORDER, TRI><<<grid, kPatchBlock, smem, st>>>(io); break;
        case 3: patch_run_kernel<3, ORDER, TRI><<<grid, kPatchBlock, smem, st>>>(io); break;
        default: patch_run_kernel<4, ORDER, TRI><<<grid, kPatchBlock, smem, st>>>(io); break;
    }
    return check_launch("patch_run_kernel");
}

template <int ORDER, bool TRI>
int launch_direct(const PatchIO &io, int LT, cudaStream_t st) {
    const int grid = patch_grid(io.n);
    const size_t smem = (size_t)(kPatchBlock / 32) * (size_t)io.warpWords * sizeof(float);
    switch (LT) {
        case 1: patch_direct_kernel<1, ORDER, TRI><<<grid, kPatchBlock, smem, st>>>(io); break;
        case 2: patch_direct_kernel<2, ORDER, TRI><<<grid, kPatchBlock, smem, st>>>(io); break;
        case 3: patch_direct_kernel<3, ORDER, TRI><<<grid, kPatchBlock, smem, st>>>(io); break;
        default: patch_direct_kernel<4, ORDER, TRI><<<grid, kPatchBlock, smem, st>>>(io); break;
    }
    return check_launch("patch_direct_kernel");
}

template <int ORDER, bool TRI>
int launch_hull(const PatchIO &io, int LT, const float *hull, cudaStream_t st) {
    const int grid = patch_grid(io.n);
    const size_t smem = (size_t)(kPatchBlock / 32) * (size_t)io.warpWords * sizeof(float);
    switch (LT) {
        case 1: patch_hull_kernel<1, ORDER, TRI><<<grid, kPatchBlock, smem, st>>>(io, hull); break;
        case 2: patch_hull_kernel<2, ORDER, TRI><<<grid, kPatchBlock, smem, st>>>(io, hull); break;
        case 3: patch_hull_kernel<3, ORDER, TRI><<<grid, kPatchBlock, smem, st>>>(io, hull); break;
        default: patch_hull_kernel<4, ORDER, TRI><<<grid, kPatchBlock, smem, st>>>(io, hull); break;
    }
    return check_launch("patch_hull_kernel");
}

int launch_hull_build(const PatchIO &io, int LT, long long rows, float *hull, cudaStream_t st) {
    const int grid = (int)std::min<long long>((rows + 255) / 256, (long long)sm_count() * 32);
    switch (LT) {
        case 1: hull_build_kernel<1><<<grid, 256, 0, st>>>(io.src, io.srcStride, io.indices, rows, hull, io.binState); break;
        case 2: hull_build_kernel<2><<<grid, 256, 0, st>>>(io.src, io.srcStride, io.indices, rows, hull, io.binState); break;
        case 3: hull_build_kernel<3><<<grid, 256, 0, st>>>(io.src, io.srcStride, io.indices, rows, hull, io.binState); break;
        default: hull_build_kernel<4><<<grid, 256, 0, st>>>(io.src, io.srcStride, io.indices, rows, hull, io.binState); break;
    }
    return check_launch("hull_build_kernel");
}

struct PatchShape {            // what the host knows about the table behind a call
    int maxPoints = 20;        // largest control hull (raw device arrays: unknown, assume the largest type)
    bool hasTri = true;        // triangle types may occur
    int options = 0;           // kPatchOpt* (PatchIO::options): any set bit selects the general kernel instantiation
};

// How a call is served.  perm / state: grouped order (state NULL: unconditionally).  hull: scratch for the per-call
// hull cache of hullRows rows (state NULL: unconditionally; otherwise the probe's verdict in *state picks between the
// hull kernels and the caller-order kernel -- both are launched, the one not chosen returns at once).
struct PatchRoute {
    const int *perm = nullptr;
    const BinState *state = nullptr;
    float *hull = nullptr;
    long long hullRows = 0;
};

#define B200_PATCH_DISPATCH(fn, ...)                                                                                   \
    ((shape.hasTri || shape.options) ? (nOut == 1 ? fn<0, true>(__VA_ARGS__) : (nOut == 3 ? fn<1, true>(__VA_ARGS__) : fn<2, true>(__VA_ARGS__))) \
                  : (nOut == 1 ? fn<0, false>(__VA_ARGS__) : (nOut == 3 ? fn<1, false>(__VA_ARGS__) : fn<2, false>(__VA_ARGS__))))

int eval_patches_common(const float *src, const int srcDesc[3], int nOut, float *const dsts[],
                        const int dstDescs[][3], int numPatchCoords, const b200osd_patch_coord *patchCoords,
                        const b200osd_patch_array *patchArrays, const int *patchIndices,
                        const b200osd_patch_param *patchParams, const PatchShape &shape,
                        const PatchRoute &route, cudaStream_t st) {
    const int L = srcDesc[1];
    // components are evaluated in tiles of at most 4 (xyz, uv, rgba fit in one launch)
    for (int c0 = 0; c0 < L; c0 += 4) {
        const int LT = (L - c0) < 4 ? (L - c0) : 4;
        PatchIO io;
        io.src = src + srcDesc[0] + c0;
        io.srcStride = srcDesc[2];
        for (int k = 0; k < kPatchMaxOut; ++k) { io.dst[k] = nullptr; io.dstStride[k] = 0; }
        for (int k = 0; k < nOut; ++k)
            if (dsts[k]) { io.dst[k] = dsts[k] + dstDescs[k][0] + c0; io.dstStride[k] = dstDescs[k][2]; }
        io.n = numPatchCoords;
        io.coords = patchCoords;
        io.arrays = patchArrays;
        io.indices = patchIndices;
        io.params = patchParams;
        io.perm = route.perm;
        io.binState = route.state;
        io.options = shape.options;
        // one record per coordinate: output k at float k*LT of an nOut*LT-float record in one buffer (glEvalLimit-style
        // interleaving, or a single tightly packed output)
        const int R = nOut * LT;
        io.packed = (LT == L) ? 1 : 0;
        for (int k = 0; k < nOut && io.packed; ++k)
            if (!io.dst[k] || io.dst[k] != io.dst[0] + (size_t)k * LT || io.dstStride[k] != R) io.packed = 0;
        io.vecStore = 0;
        if (io.packed) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(io.dst[0]);
            if (a % 16 == 0) io.vecStore |= 1;
            if (a % 8 == 0 && R % 2 == 0) io.vecStore |= 2;
        }
        if (LT == 4 && reinterpret_cast<uintptr_t>(io.src) % 16 == 0 && io.srcStride % 4 == 0) io.vecStore |= 4;
        // shared memory per warp: [ staged hulls, later the results (32 records + 32 indices) | 160 words of prefetched coordinates ]
        const int outWords = 32 * R + 32;
        int rc = B200OSD_OK;
        if (route.hull) {
            io.hullPitch = 0;
            io.coordWords = (outWords + 3) & ~3;
            io.warpWords = io.coordWords + 160;
            rc = launch_hull_build(io, LT, route.hullRows, route.hull, st);
            if (!rc) rc = B200_PATCH_DISPATCH(launch_hull, io, LT, route.hull, st);
            if (rc) return rc;
            if (!route.state) continue;                          // forced: the hull kernels did the tile
        }
        if (shape.maxPoints <= 4) {                              // linear patches: per-lane reads, nothing to stage
            io.hullPitch = 0;
            io.coordWords = (outWords + 3) & ~3;
            io.warpWords = io.coordWords + 160;
            rc = B200_PATCH_DISPATCH(launch_direct, io, LT, st);
            if (rc) return rc;
            continue;
        }
        io.hullPitch = hull_pitch(LT, shape.maxPoints);
        io.coordWords = (std::max(outWords, kHullSlots * io.hullPitch) + 3) & ~3;
        io.warpWords = io.coordWords + 160 + 64;                // + (first index, points) of the tile's distinct patches
        rc = B200_PATCH_DISPATCH(launch_run, io, LT, st);
        if (rc) return rc;
    }
    return B200OSD_OK;
}
#undef B200_PATCH_DISPATCH

// reference-mirroring argument checks (osd/cpuEvaluator.cpp:165-176,224-241,300-331); *empty = nothing to do
int validate_patch_args(const float *src, const int srcDesc[3], int nOut, float *const dsts[], const int dstDescs[][3],
                        int numPatchCoords, bool *empty) {
    *empty = false;
    if (nOut != 1 && nOut != 3 && nOut != 6) { set_error("nOut must be 1, 3 or 6 (got %d)", nOut); return B200OSD_ERR_INVALID; }
    if (!src) { set_error("src is NULL"); return B200OSD_ERR_INVALID; }
    if (nOut == 1 && !dsts[0]) { set_error("dst is NULL"); return B200OSD_ERR_INVALID; }
    const int L = srcDesc[1];
    if (L <= 0) { set_error("srcDesc.length must be positive"); return B200OSD_ERR_INVALID; }
    for (int k = 0; k < nOut; ++k)
        if (dsts[k] && dstDescs[k][1] != L) {
            set_error("output %d length %d != srcDesc.length %d", k, dstDescs[k][1], L);
            return B200OSD_ERR_INVALID;
        }
    if (numPatchCoords <= 0) *empty = true;
    return B200OSD_OK;
}

bool is_tri_type(int t) { return t == PT_TRIANGLES || t == PT_LOOP || t == PT_GREGORY_TRIANGLE; }

}  // namespace

struct b200osd_patch_table {
    struct Triple {
        b200osd_patch_array *arrays = nullptr;
        int *indices = nullptr;
        b200osd_patch_param *params = nullptr;
        int nArrays = 0, nIndices = 0, nParams = 0;
        int maxPoints = 0;             // largest control hull over the slot's arrays
        bool hasTri = false;           // a triangle patch type occurs
    };
    std::vector<Triple> triples;       // 0 vertex, 1 varying, 2+c fvar channel c
    int numFVar = 0;
    int variant = 0;                   // b200osd_patch_table_set_variant
    int options = 0;                   // b200osd_patch_table_set_options
};

struct b200osd_patch_plan {
    const b200osd_patch_table *table = nullptr;
    int maxCoords = 0, numPatches = 0;
    void *block = nullptr;             // one cudaMalloc'ed allocation, carved into `s`
    BinScratch s;
    int n = 0;                         // the coordinate set the permutation was built for
    const b200osd_patch_coord *coords = nullptr;
};

namespace {

void free_triple(b200osd_patch_table::Triple &tr) {
    cudaFree(tr.arrays); cudaFree(tr.indices); cudaFree(tr.params);
    tr = b200osd_patch_table::Triple();
}

template <typename T>
int upload_array(T **d, const T *h, int n) {
    *d = nullptr;
    if (n <= 0 || !h) return B200OSD_OK;
    cudaError_t e = cudaMalloc((void **)d, (size_t)n * sizeof(T));
    if (e != cudaSuccess) { set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); *d = nullptr; return B200OSD_ERR_ALLOC; }
    e = cudaMemcpy(*d, h, (size_t)n * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("cudaMemcpy failed: %s", cudaGetErrorString(e)); cudaFree(*d); *d = nullptr; return B200OSD_ERR_CUDA; }
    return B200OSD_OK;
}

// the (arrays, indices, params) triple a call on slot `which` evaluates, with the varying slot borrowing the vertex params
int resolve_triple(const b200osd_patch_table *t, int which, const b200osd_patch_table::Triple **tr,
                   const b200osd_patch_param **params, int *numPatches) {
    if (!t || which < 0 || which >= (int)t->triples.size()) { set_error("patch table: bad table / slot"); return B200OSD_ERR_INVALID; }
    *tr = &t->triples[which];
    *params = (*tr)->params;
    *numPatches = (*tr)->nParams;
    if (which == 1 && !*params) { *params = t->triples[0].params; *numPatches = t->triples[0].nParams; }   // osd/cudaEvaluator.h:857-878
    if (!(*tr)->arrays || !(*tr)->indices || !*params) { set_error("patch table slot %d is empty", which); return B200OSD_ERR_INVALID; }
    return B200OSD_OK;
}

}  // namespace

extern "C" {

int b200osd_eval_patches(const float *src, const int srcDesc[3], int nOut, float *const dsts[],
                         const int dstDescs[][3], int numPatchCoords, const b200osd_patch_coord *patchCoords,
                         const b200osd_patch_array *patchArrays, const int *patchIndices,
                         const b200osd_patch_param *patchParams, void *stream) {
    return b200osd_eval_patches_ex(src, srcDesc, nOut, dsts, dstDescs, numPatchCoords, patchCoords, patchArrays,
                                   patchIndices, patchParams, 0, stream);
}

int b200osd_eval_patches_ex(const float *src, const int srcDesc[3], int nOut, float *const dsts[],
                            const int dstDescs[][3], int numPatchCoords, const b200osd_patch_coord *patchCoords,
                            const b200osd_patch_array *patchArrays, const int *patchIndices,
                            const b200osd_patch_param *patchParams, int options, void *stream) {
    if (options & ~B200OSD_PATCH_GREGORY_TRUE_DERIVATIVES) { set_error("eval_patches: unknown option bits 0x%x", options); return B200OSD_ERR_INVALID; }
    bool empty = false;
    int rc = validate_patch_args(src, srcDesc, nOut, dsts, dstDescs, numPatchCoords, &empty);
    if (rc || empty) return rc;
    if (!patchCoords || !patchArrays || !patchIndices || !patchParams) { set_error("patch table / coords are NULL"); return B200OSD_ERR_INVALID; }
    // device arrays of unknown shape: any type may occur and nothing is known about the number of patches, so the
    // coordinates are evaluated in the caller's order
    PatchShape shape;
    shape.options = options;
    return eval_patches_common(src, srcDesc, nOut, dsts, dstDescs, numPatchCoords, patchCoords, patchArrays, patchIndices,
                               patchParams, shape, PatchRoute(), (cudaStream_t)stream);
}

int b200osd_patch_table_eval(const b200osd_patch_table *t, int which, const float *src, const int srcDesc[3], int nOut,
                             float *const dsts[], const int dstDescs[][3], int numPatchCoords,
                             const b200osd_patch_coord *patchCoords, void *stream) {
    const b200osd_patch_table::Triple *tr = nullptr;
    const b200osd_patch_param *params = nullptr;
    int numPatches = 0;
    int rc = resolve_triple(t, which, &tr, &params, &numPatches);
    if (rc) return rc;
    bool empty = false;
    rc = validate_patch_args(src, srcDesc, nOut, dsts, dstDescs, numPatchCoords, &empty);
    if (rc || empty) return rc;
    if (!patchCoords) { set_error("patchCoords is NULL"); return B200OSD_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    PatchShape shape;
    shape.maxPoints = std::max(tr->maxPoints, 3);
    shape.hasTri = tr->hasTri;
    shape.options = t->options;

    // How the coordinates are served (variant: 0 automatic, 1 caller's order, 2 grouped by patch per call, 3 hull cache):
    //  * caller's order, hulls staged per warp -- right for coherent sets (sorted by patch, tessellation grids) and for
    //    small calls;
    //  * per-call hull cache -- right for INCOHERENT sets when several coordinates share a patch and the hull is worth
    //    gathering (16-20 points of 3-4 floats, not the 4-point linear patches of varying data): one 192-byte read per
    //    coordinate instead of 18 scattered ones.  Whether the set is coherent is found out on the device by a sampling
    //    probe; both kernels are enqueued and the one not chosen returns at once;
    //  * grouped by patch (counting sort per call): measured slower than the hull cache on B200 -- the random 20-byte
    //    coordinate reads and 72-byte result writes cost more DRAM row activations than the hull reads they save
    //    (profiles/r02b_*) -- so it is never chosen automatically; an evaluator instance can still cache a grouping.
    const int LT0 = std::min(srcDesc[1], 4);
    const bool worth = numPatchCoords >= 65536 && (long long)numPatchCoords >= 2LL * numPatches && shape.maxPoints * LT0 >= 32;
    const int variant = t->variant;
    if (variant == 1 || (variant == 0 && !worth))
        return eval_patches_common(src, srcDesc, nOut, dsts, dstDescs, numPatchCoords, patchCoords, tr->arrays, tr->indices,
                                   params, shape, PatchRoute(), st);
    cudaMemPool_t pool;
    if ((rc = scratch_pool(&pool))) return rc;
    void *block = nullptr;
    PatchRoute route;
    if (variant == 2) {
        const size_t bytes = bin_scratch_bytes(numPatchCoords, numPatches);
        if (cudaMallocFromPoolAsync(&block, bytes, pool, st) != cudaSuccess) { cudaGetLastError(); block = nullptr; }
        if (block) {
            const BinScratch s = carve_bin_scratch(block, numPatchCoords, numPatches);
            rc = run_binning(s, patchCoords, numPatchCoords, numPatches, true, st);
            route.perm = s.perm;
            route.state = s.state;
        }
    } else {
        const size_t hullBytes = align256((size_t)tr->nIndices * LT0 * sizeof(float) + 64);   // + the last hull's partial 16-byte piece
        if (cudaMallocFromPoolAsync(&block, align256(sizeof(BinState)) + hullBytes, pool, st) != cudaSuccess) { cudaGetLastError(); block = nullptr; }
        if (block) {
            BinState *state = static_cast<BinState *>(block);
            route.hull = reinterpret_cast<float *>(static_cast<char *>(block) + align256(sizeof(BinState)));
            route.hullRows = tr->nIndices;
            if (variant == 0) {                                  // the probe decides on the device
                rc = cudaMemsetAsync(state, 0, sizeof(BinState), st) == cudaSuccess ? B200OSD_OK : B200OSD_ERR_CUDA;
                if (!rc) {
                    bin_probe_kernel<<<kBinProbeWarps / 4, 128, 0, st>>>(patchCoords, numPatchCoords, state, kPatchModeHull, 0);
                    rc = check_launch("bin_probe_kernel");
                }
                route.state = state;
            }
        }
    }
    // no scratch: the caller's order is still correct
    if (!rc)
        rc = eval_patches_common(src, srcDesc, nOut, dsts, dstDescs, numPatchCoords, patchCoords, tr->arrays, tr->indices,
                                 params, shape, block ? route : PatchRoute(), st);
    if (block) cudaFreeAsync(block, st);
    return rc;
}

// ---------------------------------------------------------------------------------- patch plan --
b200osd_patch_plan *b200osd_patch_plan_create(const b200osd_patch_table *t, int maxPatchCoords) {
    if (!t || t->triples.empty() || maxPatchCoords <= 0) { set_error("patch_plan_create: bad table / size"); return nullptr; }
    b200osd_patch_plan *p = new (std::nothrow) b200osd_patch_plan;
    if (!p) return nullptr;
    p->table = t;
    p->maxCoords = maxPatchCoords;
    p->numPatches = t->triples[0].nParams;
    for (const auto &tr : t->triples) p->numPatches = std::max(p->numPatches, tr.nParams);
    const size_t bytes = bin_scratch_bytes(maxPatchCoords, p->numPatches);
    cudaError_t e = cudaMalloc(&p->block, bytes);
    if (e != cudaSuccess) {
        set_error("patch_plan_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        delete p;
        return nullptr;
    }
    p->s = carve_bin_scratch(p->block, maxPatchCoords, p->numPatches);
    return p;
}

void b200osd_patch_plan_destroy(b200osd_patch_plan *p) {
    if (!p) return;
    cudaFree(p->block);
    delete p;
}

int b200osd_patch_plan_capacity(const b200osd_patch_plan *p) { return p ? p->maxCoords : 0; }

int b200osd_patch_plan_bin(b200osd_patch_plan *p, int numPatchCoords, const b200osd_patch_coord *patchCoords, void *stream) {
    if (!p) { set_error("patch plan is NULL"); return B200OSD_ERR_INVALID; }
    if (numPatchCoords < 0 || numPatchCoords > p->maxCoords) {
        set_error("patch_plan_bin: %d coordinates exceed the plan's capacity %d", numPatchCoords, p->maxCoords);
        return B200OSD_ERR_INVALID;
    }
    p->n = 0;
    p->coords = nullptr;
    if (numPatchCoords == 0) return B200OSD_OK;
    if (!patchCoords) { set_error("patchCoords is NULL"); return B200OSD_ERR_INVALID; }
    int rc = run_binning(p->s, patchCoords, numPatchCoords, p->numPatches, true, (cudaStream_t)stream);
    if (rc) return rc;
    p->n = numPatchCoords;
    p->coords = patchCoords;
    return B200OSD_OK;
}

int b200osd_patch_plan_eval(const b200osd_patch_plan *p, int which, const float *src, const int srcDesc[3], int nOut,
                            float *const dsts[], const int dstDescs[][3], int numPatchCoords,
                            const b200osd_patch_coord *patchCoords, void *stream) {
    if (!p) { set_error("patch plan is NULL"); return B200OSD_ERR_INVALID; }
    const b200osd_patch_table::Triple *tr = nullptr;
    const b200osd_patch_param *params = nullptr;
    int numPatches = 0;
    int rc = resolve_triple(p->table, which, &tr, &params, &numPatches);
    if (rc) return rc;
    bool empty = false;
    rc = validate_patch_args(src, srcDesc, nOut, dsts, dstDescs, numPatchCoords, &empty);
    if (rc || empty) return rc;
    if (numPatchCoords != p->n || patchCoords != p->coords) {
        set_error("patch_plan_eval: the plan was binned for another coordinate set (%d coordinates at %p)", p->n, (const void *)p->coords);
        return B200OSD_ERR_INVALID;
    }
    PatchShape shape;
    shape.maxPoints = std::max(tr->maxPoints, 3);
    shape.hasTri = tr->hasTri;
    shape.options = p->table->options;
    PatchRoute route;
    route.perm = p->s.perm;
    route.state = p->s.state;
    return eval_patches_common(src, srcDesc, nOut, dsts, dstDescs, numPatchCoords, patchCoords, tr->arrays, tr->indices,
                               params, shape, route, (cudaStream_t)stream);
}

// -------------------------------------------------------------------------------- patch table --
b200osd_patch_table *b200osd_patch_table_create(int numFVarChannels) {
    if (numFVarChannels < 0) return nullptr;
    b200osd_patch_table *t = new (std::nothrow) b200osd_patch_table;
    if (!t) return nullptr;
    t->numFVar = numFVarChannels;
    t->triples.resize(2 + numFVarChannels);
    return t;
}

void b200osd_patch_table_destroy(b200osd_patch_table *t) {
    if (!t) return;
    for (auto &tr : t->triples) free_triple(tr);
    delete t;
}

int b200osd_patch_table_set(b200osd_patch_table *t, int which, int numArrays, const b200osd_patch_array *arrays,
                            int numIndices, const int *indices, int numParams, const b200osd_patch_param *params) {
    if (!t || which < 0 || which >= (int)t->triples.size()) { set_error("patch_table_set: bad table / slot"); return B200OSD_ERR_INVALID; }
    b200osd_patch_table::Triple &tr = t->triples[which];
    free_triple(tr);
    int maxPoints = 0;
    bool hasTri = false;
    for (int a = 0; arrays && a < numArrays; ++a) {
        for (int d : { arrays[a].regDesc, arrays[a].desc }) {
            maxPoints = std::max(maxPoints, patch_type_points(d));
            hasTri = hasTri || is_tri_type(d);
        }
    }
    int rc = upload_array(&tr.arrays, arrays, numArrays);
    if (!rc) rc = upload_array(&tr.indices, indices, numIndices);
    if (!rc) rc = upload_array(&tr.params, params, numParams);
    if (rc) { free_triple(tr); return rc; }
    tr.nArrays = tr.arrays ? numArrays : 0;
    tr.nIndices = tr.indices ? numIndices : 0;
    tr.nParams = tr.params ? numParams : 0;
    tr.maxPoints = maxPoints;
    tr.hasTri = hasTri;
    return B200OSD_OK;
}

int b200osd_patch_table_num_fvar_channels(const b200osd_patch_table *t) { return t ? t->numFVar : 0; }

const void *b200osd_patch_table_buffer(const b200osd_patch_table *t, int which, int kind) {
    if (!t || which < 0 || which >= (int)t->triples.size()) return nullptr;
    const b200osd_patch_table::Triple &tr = t->triples[which];
    // the varying triple shares the vertex PatchParams (osd/cudaEvaluator.h:857-878)
    if (kind == 2 && which == 1 && !tr.params) return t->triples[0].params;
    return kind == 0 ? (const void *)tr.arrays : (kind == 1 ? (const void *)tr.indices : (const void *)tr.params);
}

int b200osd_patch_table_count(const b200osd_patch_table *t, int which, int kind) {
    if (!t || which < 0 || which >= (int)t->triples.size()) return 0;
    const b200osd_patch_table::Triple &tr = t->triples[which];
    if (kind == 2 && which == 1 && !tr.params) return t->triples[0].nParams;
    return kind == 0 ? tr.nArrays : (kind == 1 ? tr.nIndices : tr.nParams);
}

// ------------------------------------------------------------------ limit-stencil construction --
b200osd_stencil_table *b200osd_limit_stencil_table_create(const b200osd_patch_table *pt, const b200osd_stencil_table *cvStencils,
                                                          int numLocations, const b200osd_patch_coord *patchCoords,
                                                          int numWeightSets, int flags, void *stream) {
    if (!pt || pt->triples.empty() || !cvStencils || numLocations < 0 || (numLocations > 0 && !patchCoords) ||
        (numWeightSets != 1 && numWeightSets != 3 && numWeightSets != 6)) {
        set_error("limit_stencil_table_create: bad arguments");
        return nullptr;
    }
    const b200osd_patch_table::Triple &tr = pt->triples[0];
    if (!tr.arrays || !tr.indices || !tr.params) { set_error("limit_stencil_table_create: the patch table has no vertex patches"); return nullptr; }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = numLocations;
    const int tiles = (n + kScanTile - 1) / kScanTile;
    // scratch: per location {size, resolved, offset, row}, tile sums of both scans, {elements, rows, overflow}
    int *scratch = nullptr;
    const size_t words = 4 * (size_t)std::max(n, 1) + 2 * (size_t)std::max(tiles, 1) + 4;
    if (cudaMalloc((void **)&scratch, words * sizeof(int)) != cudaSuccess) {
        set_error("limit_stencil_table_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    int *sizeLoc = scratch, *resolved = scratch + n, *offLoc = scratch + 2 * (size_t)n, *rowLoc = scratch + 3 * (size_t)n;
    int *tileA = scratch + 4 * (size_t)n, *tileB = tileA + std::max(tiles, 1), *totals = tileB + std::max(tiles, 1);
    cudaMemsetAsync(totals, 0, 4 * sizeof(int), st);
    LimitIO io;
    std::memset(&io, 0, sizeof(io));
    io.n = n;
    io.coords = patchCoords;
    io.arrays = tr.arrays;
    io.patchIndices = tr.indices;
    io.params = tr.params;
    io.options = pt->options;
    io.numControlVertices = b200osd_stencil_table_num_control_vertices(cvStencils);
    io.cvSizes = static_cast<const int *>(b200osd_stencil_table_buffer(cvStencils, 0));
    io.cvOffsets = static_cast<const int *>(b200osd_stencil_table_buffer(cvStencils, 1));
    io.cvIndices = static_cast<const int *>(b200osd_stencil_table_buffer(cvStencils, 2));
    io.cvWeights = static_cast<const float *>(b200osd_stencil_table_buffer(cvStencils, 3));
    io.sizeOfLocation = sizeLoc;
    io.resolved = resolved;
    io.offsetOfLocation = offLoc;
    io.rowOfLocation = rowLoc;
    io.overflow = totals + 2;
    const int NW = numWeightSets;
    const int grid = std::max(1, std::min((n + kLimitWarps - 1) / kLimitWarps, sm_count() * 16));
    auto smem = [&](bool fill) { return (size_t)kLimitWarps * ((size_t)kLimitCap * 4 * (1 + (fill ? NW : 0)) + 20 * 6 * 4); };
    int hostTotals[4] = { 0, 0, 0, 0 };
    AdoptedArrays out;
    std::memset(&out, 0, sizeof(out));
    bool ok = true;
    if (n > 0) {
        if (NW == 1) limit_merge_kernel<1, false><<<grid, 32 * kLimitWarps, smem(false), st>>>(io);
        else if (NW == 3) limit_merge_kernel<3, false><<<grid, 32 * kLimitWarps, smem(false), st>>>(io);
        else limit_merge_kernel<6, false><<<grid, 32 * kLimitWarps, smem(false), st>>>(io);
        ok = check_launch("limit_merge_kernel(count)") == B200OSD_OK;
        if (ok) {
            scan_local_kernel<<<tiles, kScanThreads, 0, st>>>(sizeLoc, offLoc, tileA, n);
            scan_top_kernel<<<1, kScanThreads, 0, st>>>(tileA, tiles, totals + 0);
            scan_add_kernel<<<(n + 255) / 256, 256, 0, st>>>(offLoc, tileA, n);
            scan_local_kernel<<<tiles, kScanThreads, 0, st>>>(resolved, rowLoc, tileB, n);
            scan_top_kernel<<<1, kScanThreads, 0, st>>>(tileB, tiles, totals + 1);
            scan_add_kernel<<<(n + 255) / 256, 256, 0, st>>>(rowLoc, tileB, n);
            ok = check_launch("scan kernels") == B200OSD_OK &&
                 cudaMemcpyAsync(hostTotals, totals, sizeof(hostTotals), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                 cudaStreamSynchronize(st) == cudaSuccess;
        }
    }
    const long long ne = hostTotals[0];
    const int rows = hostTotals[1];
    if (ok && hostTotals[2]) { set_error("limit_stencil_table_create: a limit stencil references more than %d control vertices", kLimitCap); ok = false; }
    if (ok) {
        out.numStencils = rows;
        out.numControlVertices = io.numControlVertices;
        out.numW = NW;
        out.numElements = ne;
        auto alloc = [&](void **p, size_t bytes) { return bytes == 0 || cudaMalloc(p, bytes) == cudaSuccess; };
        ok = alloc((void **)&out.sizes, (size_t)rows * 4) && alloc((void **)&out.offsets, (size_t)rows * 4) && alloc((void **)&out.indices, (size_t)ne * 4);
        for (int k = 0; ok && k < NW; ++k) ok = alloc((void **)&out.w[k], (size_t)ne * 4);
        if (!ok) set_error("limit_stencil_table_create: cudaMalloc of the table failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (ok && n > 0 && rows > 0) {
        io.sizes = out.sizes; io.offsets = out.offsets; io.indices = out.indices;
        for (int k = 0; k < NW; ++k) io.w[k] = out.w[k];
        if (NW == 1) limit_merge_kernel<1, true><<<grid, 32 * kLimitWarps, smem(true), st>>>(io);
        else if (NW == 3) limit_merge_kernel<3, true><<<grid, 32 * kLimitWarps, smem(true), st>>>(io);
        else limit_merge_kernel<6, true><<<grid, 32 * kLimitWarps, smem(true), st>>>(io);
        ok = check_launch("limit_merge_kernel(fill)") == B200OSD_OK && cudaStreamSynchronize(st) == cudaSuccess;
        if (!ok) set_error("limit_stencil_table_create: fill failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(scratch);
    if (!ok) {
        cudaFree(out.sizes); cudaFree(out.offsets); cudaFree(out.indices);
        for (int k = 0; k < 6; ++k) cudaFree(out.w[k]);
        return nullptr;
    }
    return adopt_device_table(out, flags);
}

void b200osd_patch_table_set_variant(b200osd_patch_table *t, int variant) { if (t) t->variant = variant; }
int b200osd_patch_table_get_variant(const b200osd_patch_table *t) { return t ? t->variant : 0; }

int b200osd_patch_table_set_options(b200osd_patch_table *t, int options) {
    if (!t) { set_error("patch table is NULL"); return B200OSD_ERR_INVALID; }
    if (options & ~B200OSD_PATCH_GREGORY_TRUE_DERIVATIVES) { set_error("patch_table_set_options: unknown option bits 0x%x", options); return B200OSD_ERR_INVALID; }
    t->options = options;
    return B200OSD_OK;
}
int b200osd_patch_table_get_options(const b200osd_patch_table *t) { return t ? t->options : 0; }

}  // extern "C"
