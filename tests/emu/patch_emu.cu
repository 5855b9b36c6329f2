// TEST-ONLY numerical harness: runs the *identical* per-coordinate arithmetic of the CUDA patch kernel
// (opensubdiv_b200/csrc/patch_kernels.cuh, __host__ __device__) in a plain CPU loop, so kernel maths can be checked
// against the oracle in this GPU-less container.  Built by tests/emu/build.py into tests/emu/libpatch_emu.so.
// It is not part of libb200osd.so and nothing in opensubdiv_b200/ can reach it.
#include "../../opensubdiv_b200/csrc/patch_kernels.cuh"

#include <cstring>

namespace b200osd {
signed char g_box_tab_host[6][12][15];
float g_box_scale_host[6];
void set_error(const char *, ...) {}
std::atomic<long long> g_launches{0};
}

using namespace b200osd;

static void init_tables() {
    static bool done = false;
    if (done) return;
    static const signed char kBox12[12][15] = {
        { 1, -2, -4, 0, 6, 6, 2, 0, -6, -4, -1, -2, 0, 2, 1 },   { 1, 2, -2, 0, -6, 0, -4, 0, 6, 2, 2, 4, 0, -2, -1 },
        { 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, -1, -2, 0, 0, 0 },       { 1, -4, -2, 6, 6, 0, -4, -6, 0, 2, 1, 2, 0, -2, -1 },
        { 6, 0, 0, -12, -12, -12, 8, 12, 12, 8, -1, -2, 0, -2, -1 }, { 1, 4, 2, 6, 6, 0, -4, -6, -12, -4, -1, -2, 0, 4, 2 },
        { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 0, 0, 0 },         { 1, -2, 2, 0, -6, 0, 2, 6, 0, -4, -1, -2, 0, 4, 2 },
        { 1, 2, 4, 0, 6, 6, -4, -12, -6, -4, 2, 4, 0, -2, -1 },  { 0, 0, 0, 0, 0, 0, 2, 6, 6, 2, -1, -2, 0, -2, -1 },
        { 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, -2, -1 },       { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 1 } };
    static const signed char A[15] = { 0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 4, 3, 2, 1, 0 };
    static const signed char B[15] = { 0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4 };
    static const int das[6] = { 0, 1, 0, 2, 1, 0 }, dbs[6] = { 0, 0, 1, 0, 1, 2 }, divisor[6] = { 1, 2, 2, 12, 6, 12 };
    std::memset(g_box_tab_host, 0, sizeof(g_box_tab_host));
    for (int k = 0; k < 6; ++k)
        for (int i = 0; i < 12; ++i)
            for (int m = 0; m < 15; ++m) {
                int a = A[m], b = B[m], c = kBox12[i][m];
                if (c == 0 || a < das[k] || b < dbs[k]) continue;
                for (int q = 0; q < das[k]; ++q) c *= (a - q);
                for (int q = 0; q < dbs[k]; ++q) c *= (b - q);
                for (int mm = 0; mm < 15; ++mm)
                    if (A[mm] == a - das[k] && B[mm] == b - dbs[k]) g_box_tab_host[k][i][mm] += (signed char)(c / divisor[k]);
            }
    const float scale[6] = { 1.0f / 12.0f, 1.0f / 6.0f, 1.0f / 6.0f, 1.0f, 0.5f, 1.0f };
    std::memcpy(g_box_scale_host, scale, sizeof(scale));
    done = true;
}

template <int ORDER>
static void run(const PatchIO &io, int LT) {
    for (int i = 0; i < io.n; ++i) {
        switch (LT) {
            case 1: patch_eval_coord<1, ORDER, false>(io, i, true); break;
            case 2: patch_eval_coord<2, ORDER, false>(io, i, true); break;
            case 3: patch_eval_coord<3, ORDER, false>(io, i, true); break;
            default: patch_eval_coord<4, ORDER, false>(io, i, true); break;
        }
    }
}

extern "C" __attribute__((visibility("default")))
int emu_eval_patches(const float *src, const int srcDesc[3], int nOut, float *const dsts[], const int dstDescs[][3],
                     int n, const b200osd_patch_coord *coords, const b200osd_patch_array *arrays, const int *indices,
                     const b200osd_patch_param *params) {
    init_tables();
    const int L = srcDesc[1];
    for (int c0 = 0; c0 < L; c0 += 4) {
        const int LT = (L - c0) < 4 ? (L - c0) : 4;
        PatchIO io;
        io.src = src + srcDesc[0] + c0;
        io.srcStride = srcDesc[2];
        for (int k = 0; k < kPatchMaxOut; ++k) { io.dst[k] = nullptr; io.dstStride[k] = 0; }
        for (int k = 0; k < nOut; ++k)
            if (dsts[k]) { io.dst[k] = dsts[k] + dstDescs[k][0] + c0; io.dstStride[k] = dstDescs[k][2]; }
        io.hull4 = nullptr; io.hullStride = 0; io.hullTiles = 0; io.tile = 0;
        io.n = n; io.coords = coords; io.arrays = arrays; io.indices = indices; io.params = params;
        if (nOut == 1) run<0>(io, LT); else if (nOut == 3) run<1>(io, LT); else run<2>(io, LT);
    }
    return 0;
}
