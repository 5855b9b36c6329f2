// TEST-ONLY numerical harness: runs the *identical* per-coordinate arithmetic of the CUDA patch kernel
// (opensubdiv_b200/csrc/patch_kernels.cuh, __host__ __device__) in a plain CPU loop, so kernel maths can be checked
// against the oracle in this GPU-less container.  Built by tests/emu/build.py into tests/emu/libpatch_emu.so.
// It is not part of libb200osd.so and nothing in opensubdiv_b200/ can reach it.
#include "../../opensubdiv_b200/csrc/patch_kernels.cuh"

#include <cstring>

namespace b200osd {
void set_error(const char *, ...) {}
std::atomic<long long> g_launches{0};
}

using namespace b200osd;

template <int ORDER>
static void run(const PatchIO &io, int LT) {
    for (int i = 0; i < io.n; ++i) {
        switch (LT) {
            case 1: patch_eval_coord<1, ORDER>(io, i); break;
            case 2: patch_eval_coord<2, ORDER>(io, i); break;
            case 3: patch_eval_coord<3, ORDER>(io, i); break;
            default: patch_eval_coord<4, ORDER>(io, i); break;
        }
    }
}

extern "C" __attribute__((visibility("default")))
int emu_eval_patches_ex(const float *src, const int srcDesc[3], int nOut, float *const dsts[], const int dstDescs[][3],
                        int n, const b200osd_patch_coord *coords, const b200osd_patch_array *arrays, const int *indices,
                        const b200osd_patch_param *params, int options) {
    const int L = srcDesc[1];
    for (int c0 = 0; c0 < L; c0 += 4) {
        const int LT = (L - c0) < 4 ? (L - c0) : 4;
        PatchIO io;
        io.src = src + srcDesc[0] + c0;
        io.srcStride = srcDesc[2];
        for (int k = 0; k < kPatchMaxOut; ++k) { io.dst[k] = nullptr; io.dstStride[k] = 0; }
        for (int k = 0; k < nOut; ++k)
            if (dsts[k]) { io.dst[k] = dsts[k] + dstDescs[k][0] + c0; io.dstStride[k] = dstDescs[k][2]; }
        io.packed = 0; io.vecStore = 0; io.perm = nullptr; io.binState = nullptr; io.warpWords = 0; io.coordWords = 0; io.hullPitch = 0;
        io.n = n; io.coords = coords; io.arrays = arrays; io.indices = indices; io.params = params; io.options = options;
        if (nOut == 1) run<0>(io, LT); else if (nOut == 3) run<1>(io, LT); else run<2>(io, LT);
    }
    return 0;
}

extern "C" __attribute__((visibility("default")))
int emu_eval_patches(const float *src, const int srcDesc[3], int nOut, float *const dsts[], const int dstDescs[][3],
                     int n, const b200osd_patch_coord *coords, const b200osd_patch_array *arrays, const int *indices,
                     const b200osd_patch_param *params) {
    return emu_eval_patches_ex(src, srcDesc, nOut, dsts, dstDescs, n, coords, arrays, indices, params, 0);
}

// ---- patch map: the library's host builder + the kernel's descent, run on the CPU (test-only) --------------------
#include "../../opensubdiv_b200/csrc/patchmap.cuh"

extern "C" __attribute__((visibility("default")))
int emu_patch_map_find(int numArrays, const b200osd_patch_array *arrays, int numPatches, const b200osd_patch_param *params,
                       int triangular, int n, const int *face, const float *s, const float *t,
                       b200osd_patch_coord *out, int info[6]) {
    PatchMapHost host;
    const int rc = build_patch_map(numArrays, arrays, numPatches, params, triangular, &host);
    if (rc) return rc;
    PatchMapView m;
    m.nodes = host.nodes.data();
    m.handles = host.handles.data();
    m.minFace = host.minFace; m.maxFace = host.maxFace; m.maxDepth = host.maxDepth; m.triangular = host.triangular;
    if (info) {
        info[0] = m.minFace; info[1] = m.maxFace; info[2] = m.maxDepth; info[3] = m.triangular;
        info[4] = (int)host.nodes.size(); info[5] = (int)host.handles.size();
    }
    int hits = 0;
    for (int i = 0; i < n; ++i) {
        const int p = patch_map_find(m, face[i], s[i], t[i]);
        b200osd_patch_coord c;
        std::memset(&c, 0, sizeof(c));
        c.s = s[i]; c.t = t[i];
        if (p >= 0) { c.arrayIndex = host.handles[p].x; c.patchIndex = p; c.vertIndex = host.handles[p].y; ++hits; }
        else c.arrayIndex = -1;
        out[i] = c;
    }
    return hits;
}
