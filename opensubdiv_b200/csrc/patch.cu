// patch.cu -- host side of the limit-evaluation path: argument validation mirroring the reference
// evaluator, component tiling, kernel dispatch, and the device patch-table container.
//
// Reference behaviour mirrored (paths relative to /root/reference/opensubdiv):
//   osd/cpuEvaluator.cpp:165-176,224-241,300-331  src NULL -> false; value-only form with dst NULL -> false;
//                                                 a non-NULL output whose length != srcDesc.length -> false
//   osd/cudaKernel.cu:300-327                     NULL derivative outputs are skipped
//   osd/cudaPatchTable.cpp:69-162                 device copies of PatchArray[], indices, PatchParam[] (+varying, fvar)
#include "patch_kernels.cuh"

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

using namespace b200osd;

namespace b200osd {
signed char g_box_tab_host[6][12][15];
float g_box_scale_host[6];
}

namespace {

// Quartic box-spline basis of the regular Loop patch: 12 bivariate quartics, coefficients x12 on the monomials
//   1 s t s^2 st t^2 s^3 s^2t st^2 t^3 s^4 s^3t s^2t^2 st^3 t^4     (osd/patchBasis.h:557-572 in expanded form)
const signed char kBox12[12][15] = {
    { 1, -2, -4, 0, 6, 6, 2, 0, -6, -4, -1, -2, 0, 2, 1 },
    { 1, 2, -2, 0, -6, 0, -4, 0, 6, 2, 2, 4, 0, -2, -1 },
    { 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, -1, -2, 0, 0, 0 },
    { 1, -4, -2, 6, 6, 0, -4, -6, 0, 2, 1, 2, 0, -2, -1 },
    { 6, 0, 0, -12, -12, -12, 8, 12, 12, 8, -1, -2, 0, -2, -1 },
    { 1, 4, 2, 6, 6, 0, -4, -6, -12, -4, -1, -2, 0, 4, 2 },
    { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 0, 0, 0 },
    { 1, -2, 2, 0, -6, 0, 2, 6, 0, -4, -1, -2, 0, 4, 2 },
    { 1, 2, 4, 0, 6, 6, -4, -12, -6, -4, 2, 4, 0, -2, -1 },
    { 0, 0, 0, 0, 0, 0, 2, 6, 6, 2, -1, -2, 0, -2, -1 },
    { 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, -2, -1 },
    { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 1 },
};
const signed char kMonoA[15] = { 0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 4, 3, 2, 1, 0 };
const signed char kMonoB[15] = { 0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4 };

int g_box_device = -1;

// The derivative tables are obtained by differentiating kBox12 monomial by monomial.
int upload_box_tables() {
    static const int das[6] = { 0, 1, 0, 2, 1, 0 }, dbs[6] = { 0, 0, 1, 0, 1, 2 };
    static const int divisor[6] = { 1, 2, 2, 12, 6, 12 };
    signed char tab[6][12][15];
    std::memset(tab, 0, sizeof(tab));
    for (int k = 0; k < 6; ++k)
        for (int i = 0; i < 12; ++i)
            for (int m = 0; m < 15; ++m) {
                int a = kMonoA[m], b = kMonoB[m], c = kBox12[i][m];
                if (c == 0 || a < das[k] || b < dbs[k]) continue;
                for (int q = 0; q < das[k]; ++q) c *= (a - q);
                for (int q = 0; q < dbs[k]; ++q) c *= (b - q);
                int n = -1;
                for (int mm = 0; mm < 15; ++mm)
                    if (kMonoA[mm] == a - das[k] && kMonoB[mm] == b - dbs[k]) n = mm;
                tab[k][i][n] += (signed char)(c / divisor[k]);
            }
    const float scale[6] = { 1.0f / 12.0f, 1.0f / 6.0f, 1.0f / 6.0f, 1.0f, 0.5f, 1.0f };
    std::memcpy(g_box_tab_host, tab, sizeof(tab));
    std::memcpy(g_box_scale_host, scale, sizeof(scale));
    B200_CUDA_TRY(cudaMemcpyToSymbol(g_box_tab, tab, sizeof(tab)));
    B200_CUDA_TRY(cudaMemcpyToSymbol(g_box_scale, scale, sizeof(scale)));
    return B200OSD_OK;
}

int ensure_box_tables() {
    // __constant__ symbols are per device: (re)upload when the current device changes.
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (dev != g_box_device) {
        int rc = upload_box_tables();
        if (rc) return rc;
        g_box_device = dev;
    }
    return B200OSD_OK;
}

template <int ORDER, int MODE>
int launch_patches_h(const PatchIO &io, int LT, cudaStream_t st) {
    const int block = kPatchBlock;
    const int grid = (io.n + block - 1) / block;
    const size_t smem = (size_t)(block / 32) * (size_t)io.warpWords * sizeof(float);
    switch (LT) {
        case 1: patch_kernel<1, ORDER, MODE><<<grid, block, smem, st>>>(io); break;
        case 2: patch_kernel<2, ORDER, MODE><<<grid, block, smem, st>>>(io); break;
        case 3: patch_kernel<3, ORDER, MODE><<<grid, block, smem, st>>>(io); break;
        default: patch_kernel<4, ORDER, MODE><<<grid, block, smem, st>>>(io); break;
    }
    return check_launch("patch_kernel");
}

template <int ORDER>
int launch_patches(const PatchIO &io, int LT, int mode, cudaStream_t st) {
    switch (mode) {
        case 3: return launch_patches_h<ORDER, 3>(io, LT, st);
        case 2: return launch_patches_h<ORDER, 2>(io, LT, st);
        case 1: return launch_patches_h<ORDER, 1>(io, LT, st);
        default: return launch_patches_h<ORDER, 0>(io, LT, st);
    }
}

// 0 auto (index buffer for few coordinates per patch, else 4), 1 through the index buffer, 2 hull cache read directly,
// 3 hull cache staged in shared memory, 4 per-warp choice between 2 and 3, 100+T the same with threshold T (sweeps)
int g_patch_variant = 0;
constexpr int kStageThreshold = 8;

struct HullRequest {       // filled by b200osd_patch_table_eval when the hull cache should be used
    float4 *hull4 = nullptr;
    const int *rowsBefore = nullptr;     // device, per patch array
    const b200osd_patch_array *hostArrays = nullptr;
    int hullStride = 0, hullTiles = 0, numArrays = 0, numPatches = 0;
    int staged = 0;        // 0 direct reads, 1 always staged (MODE 2), 2 per-warp choice (MODE 3)
    int threshold = 0;
};

int eval_patches_common(const float *src, const int srcDesc[3], int nOut, float *const dsts[],
                        const int dstDescs[][3], int numPatchCoords, const b200osd_patch_coord *patchCoords,
                        const b200osd_patch_array *patchArrays, const int *patchIndices,
                        const b200osd_patch_param *patchParams, const HullRequest *hull, cudaStream_t st) {
    const int L = srcDesc[1];
    int rc = ensure_box_tables();
    if (rc) return rc;
    if (hull) {
        long long before = 0;
        for (int a = 0; a < hull->numArrays; ++a) {
            const b200osd_patch_array &pa = hull->hostArrays[a];
            const long long rows = (long long)pa.numPatches * (long long)pa.stride;
            if (rows > 0) {
                const dim3 hgrid((unsigned)((rows + 255) / 256), (unsigned)hull->hullTiles);
                hull_gather_kernel<<<hgrid, 256, 0, st>>>(src + srcDesc[0], srcDesc[2], L, patchIndices + pa.indexBase, rows,
                                                          pa.stride, before, hull->hullTiles, hull->hull4);
                rc = check_launch("hull_gather_kernel");
                if (rc) return rc;
            }
            before += rows;
        }
    }
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
        io.hull4 = hull ? hull->hull4 : nullptr;
        io.hullRowsBefore = hull ? hull->rowsBefore : nullptr;
        io.hullStride = hull ? hull->hullStride : 0;
        io.hullTiles = hull ? hull->hullTiles : 0;
        io.tile = c0 / 4;
        // glEvalLimit-style interleaving: output k at float k*LT of an nOut*LT-float record in one buffer
        io.packed = (nOut > 1 && LT == L) ? 1 : 0;
        for (int k = 0; k < nOut && io.packed; ++k)
            if (!io.dst[k] || io.dst[k] != io.dst[0] + (size_t)k * LT || io.dstStride[k] != nOut * LT) io.packed = 0;
        const int nsets = nOut;
        const int recordWords = 32 * (nsets > 1 ? ((nsets * LT) | 1) : LT);
        io.hullPitch = 0; io.stageThreshold = 0;
        int mode = hull ? 1 : 0;
        io.warpWords = recordWords;
        if (hull && hull->staged && hull->hullStride >= 12) {      // 3-4 point (linear) hulls: direct reads
            const int pitch = (hull->hullStride * LT) | 1;
            const int words = std::max(recordWords, 32 * pitch);
            if ((size_t)(kPatchBlock / 32) * words * sizeof(float) <= 48 * 1024) {     // else: direct hull reads
                mode = hull->staged == 1 ? 2 : 3;
                io.stageThreshold = hull->threshold;
                io.warpWords = words;
                io.hullPitch = pitch;
            }
        }
        rc = nOut == 1 ? launch_patches<0>(io, LT, mode, st)
                       : (nOut == 3 ? launch_patches<1>(io, LT, mode, st) : launch_patches<2>(io, LT, mode, st));
        if (rc) return rc;
    }
    return B200OSD_OK;
}

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

}  // namespace

struct b200osd_patch_table {
    struct Triple {
        b200osd_patch_array *arrays = nullptr;
        int *indices = nullptr;
        b200osd_patch_param *params = nullptr;
        int *rowsBefore = nullptr;     // hull cache layout: per array, sum over earlier arrays of numPatches * stride
        long long hullRows = 0;        // total rows of one component tile
        int nArrays = 0, nIndices = 0, nParams = 0;
    };
    std::vector<Triple> triples;   // 0 vertex, 1 varying, 2+c fvar channel c
    int numFVar = 0;
    // hull cache scratch (per-call contents; grown on demand)
    float4 *d_hull = nullptr;
    size_t hullCap = 0;
    bool hullUsed = false;               // the cache is per-call state: remember which stream last wrote / read it
    cudaStream_t hullStream = nullptr;
    std::vector<std::vector<b200osd_patch_array>> hostArrays;   // host copies of the PatchArray descriptors
};

static void free_triple(b200osd_patch_table::Triple &tr) {
    cudaFree(tr.arrays); cudaFree(tr.indices); cudaFree(tr.params); cudaFree(tr.rowsBefore);
    tr = b200osd_patch_table::Triple();
}

template <typename T>
static int upload_array(T **d, const T *h, int n) {
    *d = nullptr;
    if (n <= 0 || !h) return B200OSD_OK;
    cudaError_t e = cudaMalloc((void **)d, (size_t)n * sizeof(T));
    if (e != cudaSuccess) { set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); *d = nullptr; return B200OSD_ERR_ALLOC; }
    e = cudaMemcpy(*d, h, (size_t)n * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("cudaMemcpy failed: %s", cudaGetErrorString(e)); cudaFree(*d); *d = nullptr; return B200OSD_ERR_CUDA; }
    return B200OSD_OK;
}

extern "C" {

int b200osd_eval_patches(const float *src, const int srcDesc[3], int nOut, float *const dsts[],
                         const int dstDescs[][3], int numPatchCoords, const b200osd_patch_coord *patchCoords,
                         const b200osd_patch_array *patchArrays, const int *patchIndices,
                         const b200osd_patch_param *patchParams, void *stream) {
    bool empty = false;
    int rc = validate_patch_args(src, srcDesc, nOut, dsts, dstDescs, numPatchCoords, &empty);
    if (rc || empty) return rc;
    if (!patchCoords || !patchArrays || !patchIndices || !patchParams) { set_error("patch table / coords are NULL"); return B200OSD_ERR_INVALID; }
    return eval_patches_common(src, srcDesc, nOut, dsts, dstDescs, numPatchCoords, patchCoords, patchArrays, patchIndices,
                               patchParams, nullptr, (cudaStream_t)stream);
}

int b200osd_patch_table_eval(const b200osd_patch_table *tc, int which, const float *src, const int srcDesc[3], int nOut,
                             float *const dsts[], const int dstDescs[][3], int numPatchCoords,
                             const b200osd_patch_coord *patchCoords, void *stream) {
    b200osd_patch_table *t = const_cast<b200osd_patch_table *>(tc);
    if (!t || which < 0 || which >= (int)t->triples.size()) { set_error("patch_table_eval: bad table / slot"); return B200OSD_ERR_INVALID; }
    bool empty = false;
    int rc = validate_patch_args(src, srcDesc, nOut, dsts, dstDescs, numPatchCoords, &empty);
    if (rc || empty) return rc;
    const b200osd_patch_table::Triple &tr = t->triples[which];
    const b200osd_patch_param *params = tr.params;
    int numPatches = tr.nParams;
    if (which == 1 && !params) { params = t->triples[0].params; numPatches = t->triples[0].nParams; }   // varying shares vertex params
    if (!patchCoords || !tr.arrays || !tr.indices || !params) { set_error("patch table slot %d is empty", which); return B200OSD_ERR_INVALID; }

    // Hull cache: when many coordinates share few patches, gather every patch's control points once per call into
    // 16-byte rows so that a coordinate reads one compact, aligned block instead of 16-20 scattered vertices.
    HullRequest hull;
    bool useHull = (g_patch_variant >= 2) || (g_patch_variant == 0 && (long long)numPatchCoords >= 4LL * numPatches);
    if (useHull && numPatches > 0) {
        int hs = 0;
        for (int a = 0; a < tr.nArrays; ++a) hs = std::max(hs, t->hostArrays[which][a].stride);
        if (hs < 1 || hs > 32 || tr.hullRows <= 0 || !tr.rowsBefore) useHull = false;   // one warp lane per control point
        const int tiles = (srcDesc[1] + 3) / 4;
        const size_t need = (size_t)std::max(tr.hullRows, 0LL) * tiles;
        if (useHull && need > t->hullCap) {
            cudaFree(t->d_hull);
            t->d_hull = nullptr;
            t->hullCap = 0;
            if (cudaMalloc((void **)&t->d_hull, need * sizeof(float4)) != cudaSuccess) {
                cudaGetLastError();
                useHull = false;                      // not enough memory for the cache: evaluate through the indices
            } else {
                t->hullCap = need;
            }
        }
        if (useHull) {
            // The cache is shared by every call on this table.  Calls on ONE stream are ordered by the stream; when the
            // stream changes, the previous call may still be reading the cache, so wait for the device once (the
            // reference never meets this: all its launches go to the legacy default stream).  While the new stream is
            // being captured no synchronisation is possible -- ordering is then the caller's job (b200osd_capi.h).
            cudaStream_t cur = (cudaStream_t)stream;
            if (t->hullUsed && t->hullStream != cur) {
                cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
                if (cudaStreamIsCapturing(cur, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
                if (cs == cudaStreamCaptureStatusNone) B200_CUDA_TRY(cudaDeviceSynchronize());
            }
            t->hullUsed = true;
            t->hullStream = cur;
            hull.hull4 = t->d_hull;
            hull.rowsBefore = tr.rowsBefore;
            hull.hostArrays = t->hostArrays[which].data();
            hull.hullStride = hs;
            hull.hullTiles = tiles;
            hull.numArrays = tr.nArrays;
            hull.numPatches = numPatches;
            // auto: per-warp choice while the cache is L2-resident (incoherent warps are then bound by L1 wavefronts and
            // staging pays); a cache far larger than L2 makes incoherent reads DRAM-bound and staging only adds work
            const bool l2Resident = need * sizeof(float4) <= (size_t)64 << 20;
            hull.staged = g_patch_variant == 3 ? 1 : (g_patch_variant == 2 ? 0 : (g_patch_variant == 0 && !l2Resident ? 0 : 2));
            hull.threshold = g_patch_variant >= 100 ? g_patch_variant - 100 : kStageThreshold;
        }
    }
    return eval_patches_common(src, srcDesc, nOut, dsts, dstDescs, numPatchCoords, patchCoords, tr.arrays, tr.indices,
                               params, useHull ? &hull : nullptr, (cudaStream_t)stream);
}

void b200osd_set_patch_variant(int v) { g_patch_variant = v; }
int b200osd_get_patch_variant(void) { return g_patch_variant; }

// -------------------------------------------------------------------------------- patch table --
b200osd_patch_table *b200osd_patch_table_create(int numFVarChannels) {
    if (numFVarChannels < 0) return nullptr;
    b200osd_patch_table *t = new (std::nothrow) b200osd_patch_table;
    if (!t) return nullptr;
    t->numFVar = numFVarChannels;
    t->triples.resize(2 + numFVarChannels);
    t->hostArrays.resize(2 + numFVarChannels);
    return t;
}


void b200osd_patch_table_destroy(b200osd_patch_table *t) {
    if (!t) return;
    for (auto &tr : t->triples) free_triple(tr);
    cudaFree(t->d_hull);
    delete t;
}

int b200osd_patch_table_set(b200osd_patch_table *t, int which, int numArrays, const b200osd_patch_array *arrays,
                            int numIndices, const int *indices, int numParams, const b200osd_patch_param *params) {
    if (!t || which < 0 || which >= (int)t->triples.size()) { set_error("patch_table_set: bad table / slot"); return B200OSD_ERR_INVALID; }
    b200osd_patch_table::Triple &tr = t->triples[which];
    free_triple(tr);
    int rc = upload_array(&tr.arrays, arrays, numArrays);
    if (!rc) rc = upload_array(&tr.indices, indices, numIndices);
    if (!rc) rc = upload_array(&tr.params, params, numParams);
    std::vector<int> before((size_t)std::max(numArrays, 0));
    long long rows = 0;
    for (int a = 0; arrays && a < numArrays; ++a) {
        before[(size_t)a] = (int)rows;
        rows += (long long)arrays[a].numPatches * (long long)std::max(arrays[a].stride, 0);
    }
    if (!rc && rows > 0x7fffffffLL) rows = -1;                 // too large for 32-bit row offsets: no hull cache
    if (!rc) rc = upload_array(&tr.rowsBefore, before.data(), arrays ? numArrays : 0);
    if (rc) { free_triple(tr); return rc; }
    tr.hullRows = rows;
    t->hostArrays[which].assign(arrays, arrays + (arrays ? numArrays : 0));
    tr.nArrays = tr.arrays ? numArrays : 0;
    tr.nIndices = tr.indices ? numIndices : 0;
    tr.nParams = tr.params ? numParams : 0;
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

}  // extern "C"
