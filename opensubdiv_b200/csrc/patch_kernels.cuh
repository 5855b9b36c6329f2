// patch_kernels.cuh -- device code of the limit-evaluation path (EvalPatches*), sm_100a.
//
// Semantics restated from the reference (paths relative to /root/reference/opensubdiv):
//   osd/cpuEvaluator.cpp:157-381   per PatchCoord: array/param lookup, basis weights, gather-sum of control points
//   osd/patchBasisTypes.h:241-426  patch type ids, PatchParam bit fields, (s,t) normalisation into the sub-patch
//   osd/patchBasis.h:53-1610       linear / B-spline / Gregory / box-spline / Gregory-triangle bases, boundary
//                                  folding, derivative scaling (d1 = +-2^depth, d2 = sign*d1*d1)
//
// Design (one thread per PatchCoord, everything in registers):
//   REGULAR (bicubic B-spline, the overwhelmingly common type) is evaluated in SEPARABLE form: the 1-D weight
//   vectors in s and t (value / 1st / 2nd derivative, boundary-folded and pre-scaled) are kept, each of the 4 rows
//   of control points is first contracted with the three s-vectors, and the row results are then combined with
//   the t-vectors.  That is 16*3 + 4*6 = 72 multiply-adds per component instead of the 16*6 = 96 (plus 96
//   products to form the tensor weights) of the reference formulation; boundary folding is linear in the 1-D
//   weights so it commutes with the tensor product.  Summation order differs from the reference, see DESIGN.md.
//   GREGORY_BASIS and QUADS are unrolled per control point with compile-time point tables.
//   LOOP / GREGORY_TRIANGLE / TRIANGLES use a weight-array formulation (local memory), correct but not tuned.
#pragma once

#include "common.cuh"

// Every function below is __host__ __device__ so that tests/emu can run the *identical* arithmetic on the CPU
// (a test-only numerical harness, never a product path: libb200osd.so only ever launches the __global__ kernel).
#define B200_HD __host__ __device__ __forceinline__
#define B200_HD_NOINLINE __host__ __device__

namespace b200osd {

constexpr int kPatchMaxOut = 6;

struct PatchIO {
    const float *src;                 // already offset by srcDesc.offset (+ component tile offset)
    int srcStride;
    float *dst[kPatchMaxOut];         // already offset (+ component tile offset); NULL = skip
    int dstStride[kPatchMaxOut];
    int n;
    const b200osd_patch_coord *coords;
    const b200osd_patch_array *arrays;
    const int *indices;
    const b200osd_patch_param *params;
    // optional per-call hull cache (see hull_gather_kernel): control points of every patch gathered into 16-byte
    // rows, array by array: row(a, p, tile, j) = (hullRowsBefore[a] + (p - primitiveIdBase_a) * stride_a) * tiles
    //                                            + tile * stride_a + j      -- every hull of a 16-point array is 256 B aligned
    const float4 *hull4;
    const int *hullRowsBefore;        // per patch array: sum over earlier arrays of numPatches * stride
    int hullStride;                   // largest array stride (row capacity of a staged hull)
    int hullTiles;                    // ceil(L / 4)
    int tile;                         // component tile evaluated by this launch
    int packed;                       // outputs form one contiguous NSETS*LT-float record per coordinate (see store_outputs)
    // shared memory: one private region of warpWords floats per warp (output transposition; staged hulls alias it)
    int warpWords;
    int hullPitch;                    // staged hulls: floats per lane row = (hullStride * LT) | 1
    int stageThreshold;               // MODE 3: stage when a warp touches more distinct patches than this
};

enum { PT_QUADS = 3, PT_TRIANGLES = 4, PT_LOOP = 5, PT_REGULAR = 6, PT_GREGORY_BASIS = 9, PT_GREGORY_TRIANGLE = 10 };

// Derived box-spline tables (filled once on the host, see patch.cu): g_box_tab[k][i][m], k = value,ds,dt,dss,dst,dtt
__constant__ signed char g_box_tab[6][12][15];
__constant__ float g_box_scale[6];
#ifndef __CUDA_ARCH__
extern signed char g_box_tab_host[6][12][15];     // host mirror (filled by the same derivation) for tests/emu
extern float g_box_scale_host[6];
#endif

B200_HD float box_coeff(int k, int i, int m) {
#ifdef __CUDA_ARCH__
    return (float)g_box_tab[k][i][m];
#else
    return (float)g_box_tab_host[k][i][m];
#endif
}
B200_HD float box_scale(int k) {
#ifdef __CUDA_ARCH__
    return g_box_scale[k];
#else
    return g_box_scale_host[k];
#endif
}
B200_HD float rcp_rn(float x) {
#ifdef __CUDA_ARCH__
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
}
B200_HD int ldg_i(const int *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
B200_HD unsigned ldg_u(const unsigned *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
B200_HD float ldg_f(const float *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
B200_HD int ld_coord_word(const int *p) {
#ifdef __CUDA_ARCH__
    return ld_stream_i1(p);
#else
    return *p;
#endif
}
B200_HD void st_out(float *p, float v) {
#ifdef __CUDA_ARCH__
    st_stream_f1(p, v);
#else
    *p = v;
#endif
}
B200_HD float int_as_float(int v) {
#ifdef __CUDA_ARCH__
    return __int_as_float(v);
#else
    union { int i; float f; } u; u.i = v; return u.f;
#endif
}

// ---------------------------------------------------------------------------------- 1-D bases --
template <int ORDER>
B200_HD void bspline_1d(float t, float (&b)[4], float (&d)[4], float (&dd)[4]) {
    const float c = 1.0f - t;
    const float t2 = t * t, c2 = c * c;
    const float sixth = 1.0f / 6.0f;
    b[0] = sixth * c2 * c;
    b[3] = sixth * t2 * t;
    b[1] = sixth * fmaf(t2, fmaf(3.0f, t, -6.0f), 4.0f);          // (3t^3 - 6t^2 + 4)/6
    b[2] = sixth * fmaf(c2, fmaf(3.0f, c, -6.0f), 4.0f);          // symmetric: b2(t) = b1(1-t)
    if (ORDER >= 1) {
        d[0] = -0.5f * c2;
        d[3] = 0.5f * t2;
        d[1] = t * fmaf(1.5f, t, -2.0f);
        d[2] = -c * fmaf(1.5f, c, -2.0f);
    }
    if (ORDER >= 2) {
        dd[0] = c;
        dd[1] = fmaf(3.0f, t, -2.0f);
        dd[2] = fmaf(3.0f, c, -2.0f);
        dd[3] = t;
    }
}

template <int ORDER>
B200_HD void bezier_1d(float t, float (&b)[4], float (&d)[4], float (&dd)[4]) {
    const float c = 1.0f - t;
    const float t2 = t * t, c2 = c * c;
    b[0] = c2 * c;
    b[1] = 3.0f * c2 * t;
    b[2] = 3.0f * t2 * c;
    b[3] = t2 * t;
    if (ORDER >= 1) {
        d[0] = -3.0f * c2;
        d[1] = 3.0f * c * fmaf(-3.0f, t, 1.0f);                   // 3(1-t)(1-3t)
        d[2] = 3.0f * t * fmaf(-3.0f, t, 2.0f);                   // 3t(2-3t)
        d[3] = 3.0f * t2;
    }
    if (ORDER >= 2) {
        dd[0] = 6.0f * c;
        dd[1] = fmaf(18.0f, t, -12.0f);
        dd[2] = fmaf(-18.0f, t, 6.0f);
        dd[3] = 6.0f * t;
    }
}

// phantom end point folding of a 1-D weight vector: lo = fold index 0 into 1,2 ; hi = fold index 3 into 2,1
B200_HD void fold_lo(float (&w)[4]) { w[2] -= w[0]; w[1] = fmaf(2.0f, w[0], w[1]); w[0] = 0.0f; }
B200_HD void fold_hi(float (&w)[4]) { w[1] -= w[3]; w[2] = fmaf(2.0f, w[3], w[2]); w[3] = 0.0f; }

// Control-point access of one patch: either through the patch's index list into the caller's primvar buffer ...
struct CvIndirect {
    const float *src;
    int stride;
    const int *cvs;
    template <int LT>
    B200_HD void load(int j, float (&v)[LT]) const {
        const float *p = src + (size_t)ldg_i(cvs + j) * (size_t)stride;
#pragma unroll
        for (int c = 0; c < LT; ++c) v[c] = ldg_f(p + c);
    }
};
// ... or from the hull cache: point j of this patch's component tile is one aligned 16-byte row
struct CvHull {
    const float4 *base;
    template <int LT>
    B200_HD void load(int j, float (&v)[LT]) const {
#ifdef __CUDA_ARCH__
        const float4 t = __ldg(base + j);
#else
        const float4 t = base[j];
#endif
        if (LT > 0) v[0] = t.x;
        if (LT > 1) v[1] = t.y;
        if (LT > 2) v[2] = t.z;
        if (LT > 3) v[3] = t.w;
    }
};

// ... or from the warp's shared-memory copy of its 32 hulls (MODE 2): LT floats per point, odd row pitch
struct CvSmem {
    const float *row;
    template <int LT>
    B200_HD void load(int j, float (&v)[LT]) const {
#pragma unroll
        for (int c = 0; c < LT; ++c) v[c] = row[j * LT + c];
    }
};

#ifdef __CUDACC__
extern __shared__ float b200_patch_smem[];
#endif

// first hull-cache row of patch p of the array described by aw (the 6 ints of its PatchArray), component tile `tile`
B200_HD size_t hull_first_row(const int *rowsBefore, int arrayIndex, const int *aw, int p, int tiles, int tile, int *points) {
    const int stride = ldg_i(aw + 4), primBase = ldg_i(aw + 5);
    *points = stride;
    return ((size_t)ldg_i(rowsBefore + arrayIndex) + (size_t)(p - primBase) * (size_t)stride) * (size_t)tiles
           + (size_t)tile * (size_t)stride;
}

constexpr int kPatchBlock = 128;

// Results leave through shared memory: a warp's 32 x LT values of one output are transposed so that consecutive lanes
// write consecutive floats (one 128-byte request per 32 floats when the output is packed, runs of LT otherwise)
// instead of 32 scattered LT-float records.  `live` = this lane holds a real coordinate; i0 = the warp's first one.
// io.packed: all NSETS outputs interleave into ONE record of NSETS*LT floats per coordinate (the glEvalLimit layout,
// examples/glEvalLimit/glEvalLimit.cpp:277-287) -- then the warp's whole 32 x NSETS*LT block is staged and leaves as
// NSETS*LT fully coalesced 128-byte rows instead of LT-float runs at a stride (3.5x fewer L2 write sectors for 18 floats).
template <int LT, int NSETS>
B200_HD void store_outputs(const PatchIO &io, int i, bool live, const float (&out)[NSETS][LT]) {
#ifdef __CUDA_ARCH__
    constexpr int R = NSETS * LT;                 // floats per packed record
    constexpr int RP = R | 1;                     // odd row pitch: conflict-free transposition
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = i - lane;
    const unsigned livemask = __ballot_sync(0xffffffffu, live);   // lanes past the end or holding a miss write nothing
    float *st = b200_patch_smem + (size_t)warp * (size_t)io.warpWords;
    __syncwarp();                                               // the region may still be read as staged hulls
    if (NSETS > 1 && io.packed) {                               // uniform across the grid
#pragma unroll
        for (int k = 0; k < NSETS; ++k)
#pragma unroll
            for (int c = 0; c < LT; ++c) st[lane * RP + k * LT + c] = out[k][c];
        __syncwarp();
        float *base = io.dst[0] + (size_t)i0 * (size_t)R;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int e = q * 32 + lane;
            const int ci = e / R, c = e - ci * R;
            if ((livemask >> ci) & 1u) st_stream_f1(base + e, st[ci * RP + c]);
        }
        return;
    }
    // separate buffers (or a strided / partial record): all outputs are staged at once, [output][coordinate][component]
#pragma unroll
    for (int k = 0; k < NSETS; ++k)
#pragma unroll
        for (int c = 0; c < LT; ++c) st[k * 32 * LT + lane * LT + c] = out[k][c];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        float *d = io.dst[k];
        if (!d) continue;                                   // uniform across the grid
        const size_t stride = (size_t)io.dstStride[k];
        float *base = d + (size_t)i0 * stride;
#pragma unroll
        for (int q = 0; q < LT; ++q) {
            const int e = q * 32 + lane;
            const int ci = e / LT, c = e - ci * LT;
            if ((livemask >> ci) & 1u) st_stream_f1(base + (size_t)ci * stride + c, st[k * 32 * LT + e]);
        }
    }
#else
    if (!live) return;
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        float *d = io.dst[k];
        if (!d) continue;
        d += (size_t)i * (size_t)io.dstStride[k];
#pragma unroll
        for (int c = 0; c < LT; ++c) st_out(d + c, out[k][c]);
    }
#endif
}

// -------------------------------------------------------------------------------- REGULAR path --
template <int LT, int ORDER, typename CV>
B200_HD void eval_regular(const CV &cv, float s, float t, int boundary,
                          float d1, float (&out)[ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6)][LT]) {
    float bs[4], ds[4], dss[4], bt[4], dt[4], dtt[4];
    bspline_1d<ORDER>(s, bs, ds, dss);
    bspline_1d<ORDER>(t, bt, dt, dtt);
    if (boundary) {
        if (boundary & 1) { fold_lo(bt); if (ORDER >= 1) fold_lo(dt); if (ORDER >= 2) fold_lo(dtt); }
        if (boundary & 2) { fold_hi(bs); if (ORDER >= 1) fold_hi(ds); if (ORDER >= 2) fold_hi(dss); }
        if (boundary & 4) { fold_hi(bt); if (ORDER >= 1) fold_hi(dt); if (ORDER >= 2) fold_hi(dtt); }
        if (boundary & 8) { fold_lo(bs); if (ORDER >= 1) fold_lo(ds); if (ORDER >= 2) fold_lo(dss); }
    }
    if (ORDER >= 1) {
        const float d2 = d1 * d1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ds[j] *= d1; dt[j] *= d1;
            if (ORDER >= 2) { dss[j] *= d2; dtt[j] *= d2; }
        }
    }
    constexpr int NSETS = ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6);
#pragma unroll
    for (int k = 0; k < NSETS; ++k)
#pragma unroll
        for (int c = 0; c < LT; ++c) out[k][c] = 0.0f;

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float r0[LT], r1[LT], r2[LT];
#pragma unroll
        for (int c = 0; c < LT; ++c) { r0[c] = 0.0f; r1[c] = 0.0f; r2[c] = 0.0f; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[LT];
            cv.template load<LT>(4 * i + j, v);
#pragma unroll
            for (int c = 0; c < LT; ++c) {
                r0[c] = fmaf(bs[j], v[c], r0[c]);
                if (ORDER >= 1) r1[c] = fmaf(ds[j], v[c], r1[c]);
                if (ORDER >= 2) r2[c] = fmaf(dss[j], v[c], r2[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < LT; ++c) {
            out[0][c] = fmaf(bt[i], r0[c], out[0][c]);
            if (ORDER >= 1) {
                out[1][c] = fmaf(bt[i], r1[c], out[1][c]);
                out[2][c] = fmaf(dt[i], r0[c], out[2][c]);
            }
            if (ORDER >= 2) {
                out[3][c] = fmaf(bt[i], r2[c], out[3][c]);
                out[4][c] = fmaf(dt[i], r1[c], out[4][c]);
                out[5][c] = fmaf(dtt[i], r0[c], out[5][c]);
            }
        }
    }
}

// ---------------------------------------------------------------------------- GREGORY_BASIS path --
// 20 points, 5 per corner c: P, E+, E-, F+, F-; Bezier net position (col,row) per point; the face points carry
// the rational blend G+ = a/(a+b), G- = 1-G+ with (a,b) the distances from corner c along E+ / E-
// (osd/patchBasis.h:345-378); the reciprocal is replaced by 1 when a+b <= 0.  Derivatives use the reference's
// default approximation: Bezier derivative weights times the same G (osd/patchBasis.h:421-440).
template <int LT, int ORDER, typename CV>
B200_HD void eval_gregory(const CV &cv, float s, float t, float d1,
                          float (&out)[ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6)][LT]) {
    constexpr int COL[20] = { 0, 1, 0, 1, 1, 3, 3, 2, 2, 2, 3, 2, 3, 2, 2, 0, 0, 1, 1, 1 };
    constexpr int ROW[20] = { 0, 0, 1, 1, 1, 0, 1, 0, 1, 1, 3, 3, 2, 2, 2, 3, 2, 3, 2, 2 };
    float bs[4], ds[4], dss[4], bt[4], dt[4], dtt[4];
    bezier_1d<ORDER>(s, bs, ds, dss);
    bezier_1d<ORDER>(t, bt, dt, dtt);
    const float sc = 1.0f - s, tc = 1.0f - t;
    float G[8];
    {
        const float a[4] = { s, t, sc, tc };
        const float den[4] = { s + t, sc + t, sc + tc, s + tc };
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float r = (den[c] <= 0.0f) ? 1.0f : rcp_rn(den[c]);
            G[2 * c] = a[c] * r;
            G[2 * c + 1] = 1.0f - G[2 * c];
        }
    }
    if (ORDER >= 1) {
        const float d2 = d1 * d1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ds[j] *= d1; dt[j] *= d1;
            if (ORDER >= 2) { dss[j] *= d2; dtt[j] *= d2; }
        }
    }
    constexpr int NSETS = ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6);
#pragma unroll
    for (int k = 0; k < NSETS; ++k)
#pragma unroll
        for (int c = 0; c < LT; ++c) out[k][c] = 0.0f;
#pragma unroll
    for (int i = 0; i < 20; ++i) {
        const int col = COL[i], row = ROW[i], p = i % 5;
        const float g = (p >= 3) ? G[2 * (i / 5) + (p - 3)] : 1.0f;
        float v[LT];
        cv.template load<LT>(i, v);
        const float gs = bs[col] * g, gt = bt[row];
        float w[NSETS];
        w[0] = gs * gt;
        if (ORDER >= 1) { w[1] = ds[col] * g * gt; w[2] = gs * dt[row]; }
        if (ORDER >= 2) { w[3] = dss[col] * g * gt; w[4] = ds[col] * g * dt[row]; w[5] = gs * dtt[row]; }
#pragma unroll
        for (int k = 0; k < NSETS; ++k)
#pragma unroll
            for (int c = 0; c < LT; ++c) out[k][c] = fmaf(w[k], v[c], out[k][c]);
    }
}

// ---------------------------------------------------------------------------------- QUADS path --
template <int LT, int ORDER, typename CV>
B200_HD void eval_quads(const CV &cv, float s, float t, float d1,
                        float (&out)[ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6)][LT]) {
    const float sc = 1.0f - s, tc = 1.0f - t;
    const float wP[4] = { sc * tc, s * tc, s * t, sc * t };
    const float wS[4] = { -tc * d1, tc * d1, t * d1, -t * d1 };
    const float wT[4] = { -sc * d1, -s * d1, s * d1, sc * d1 };
    const float d2 = d1 * d1;
    const float wST[4] = { d2, -d2, d2, -d2 };
    constexpr int NSETS = ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6);
#pragma unroll
    for (int k = 0; k < NSETS; ++k)
#pragma unroll
        for (int c = 0; c < LT; ++c) out[k][c] = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float v[LT];
        cv.template load<LT>(i, v);
#pragma unroll
        for (int c = 0; c < LT; ++c) {
            out[0][c] = fmaf(wP[i], v[c], out[0][c]);
            if (ORDER >= 1) { out[1][c] = fmaf(wS[i], v[c], out[1][c]); out[2][c] = fmaf(wT[i], v[c], out[2][c]); }
            if (ORDER >= 2) out[4][c] = fmaf(wST[i], v[c], out[4][c]);
        }
    }
}

// ------------------------------------------------------------- triangle bases (weight-array form) --
B200_HD void refl(float *w, int phantom, int plus0, int plus1, int minus) {
    const float v = w[phantom];
    w[plus0] += v;
    w[plus1] += v;
    w[minus] -= v;
}

// Box-spline boundary folding (osd/patchBasis.h:663-886): every phantom point is a reflection B + (B' - I).
B200_HD_NOINLINE void box_fold_boundary(int mask, float *w) {
    const signed char PH[3][3] = { { 0, 1, 2 }, { 6, 9, 11 }, { 10, 7, 3 } };
    const signed char B1[3] = { 4, 5, 8 }, B2[3] = { 5, 8, 4 }, I1[3] = { 8, 4, 5 };
    const signed char B0[3] = { 3, 2, 11 }, I0[3] = { 7, 1, 9 };
    const signed char B3[3] = { 6, 10, 0 }, I2[3] = { 9, 7, 1 };
    const signed char VP[3][2] = { { 3, 0 }, { 2, 6 }, { 11, 10 } };
    const signed char VB0[3] = { 7, 1, 9 }, VI0[3] = { 8, 4, 5 };
    const signed char VB2[3] = { 1, 9, 7 }, VI1[3] = { 5, 8, 4 };
    const int upper = (mask >> 3) & 3;
    int ebits = mask & 7, vbits = 0;
    if (upper == 1) { vbits = ebits; ebits = 0; }
    else if (upper == 2) { vbits = ((ebits & 1) << 2) | (ebits >> 1); }
    for (int e = 0; e < 3; ++e) {
        if (!(ebits & (1 << e))) continue;
        const int prev = (e + 2) % 3, next = (e + 1) % 3;
        if (ebits & (1 << prev)) refl(w, PH[e][0], B1[e], B1[e], I1[e]);
        else                     refl(w, PH[e][0], B1[e], B0[e], I0[e]);
        refl(w, PH[e][1], B1[e], B2[e], I1[e]);
        if (ebits & (1 << next)) refl(w, PH[e][2], B2[e], B2[e], I1[e]);
        else                     refl(w, PH[e][2], B2[e], B3[e], I2[e]);
        w[PH[e][0]] = 0.0f; w[PH[e][1]] = 0.0f; w[PH[e][2]] = 0.0f;
    }
    for (int v = 0; v < 3; ++v) {
        if (!(vbits & (1 << v))) continue;
        refl(w, VP[v][0], B1[v], VB0[v], VI0[v]);
        refl(w, VP[v][1], B1[v], VB2[v], VI1[v]);
        w[VP[v][0]] = 0.0f; w[VP[v][1]] = 0.0f;
    }
}

B200_HD float bern(int n, int i, int j, int k, float u, float v, float w) {
    if (i < 0 || j < 0 || k < 0) return 0.0f;
    const float fact[5] = { 1.0f, 1.0f, 2.0f, 6.0f, 24.0f };
    float r = fact[n] / (fact[i] * fact[j] * fact[k]);
    for (int q = 0; q < i; ++q) r *= u;
    for (int q = 0; q < j; ++q) r *= v;
    for (int q = 0; q < k; ++q) r *= w;
    return r;
}

// Weight arrays for LOOP (12), GREGORY_TRIANGLE (18) and TRIANGLES (3); returns the number of points.
template <int ORDER>
B200_HD_NOINLINE int tri_weights(int type, float s, float t, int boundary, float (*w)[20]) {
    constexpr int NSETS = ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6);
    if (type == PT_TRIANGLES) {
        w[0][0] = 1.0f - s - t; w[0][1] = s; w[0][2] = t;
        if (ORDER >= 1) {
            w[1][0] = -1.0f; w[1][1] = 1.0f; w[1][2] = 0.0f;
            w[2][0] = -1.0f; w[2][1] = 0.0f; w[2][2] = 1.0f;
        }
        if (ORDER >= 2)
            for (int k = 3; k < 6; ++k) { w[k][0] = 0.0f; w[k][1] = 0.0f; w[k][2] = 0.0f; }
        return 3;
    }
    if (type == PT_LOOP) {
        float M[15];
        M[0] = 1.0f; M[1] = s; M[2] = t;
        M[3] = s * s; M[4] = s * t; M[5] = t * t;
        M[6] = M[3] * s; M[7] = M[4] * s; M[8] = M[4] * t; M[9] = M[5] * t;
        M[10] = M[6] * s; M[11] = M[7] * s; M[12] = M[3] * M[5]; M[13] = M[8] * t; M[14] = M[9] * t;
        for (int k = 0; k < NSETS; ++k) {
            for (int i = 0; i < 12; ++i) {
                float acc = 0.0f;
#pragma unroll
                for (int m = 0; m < 15; ++m) acc = fmaf(box_coeff(k, i, m), M[m], acc);
                w[k][i] = box_scale(k) * acc;
            }
            if (boundary) box_fold_boundary(boundary, w[k]);
        }
        return 12;
    }
    // GREGORY_TRIANGLE: quartic Bernstein over the triangle + rational blends on the 3 interior points
    {
        const signed char PI[15] = { 0, 1, 2, 3, 4, 0, 1, 2, 3, 0, 1, 2, 0, 1, 0 };
        const signed char PJ[15] = { 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 4 };
        const signed char SRC[18] = { 0, 1, 5, 6, 6, 4, 8, 3, 7, 7, 14, 12, 13, 10, 10, 2, 11, 9 };
        const signed char GI[18] = { -1, -1, -1, 0, 1, -1, -1, -1, 2, 3, -1, -1, -1, 4, 5, -1, -1, -1 };
        const int DS[6] = { 0, 1, 0, 2, 1, 0 }, DT[6] = { 0, 0, 1, 0, 1, 2 };
        const float u = s, v = t, ww = 1.0f - u - v;
        float G[6] = { 1.0f, 0.0f, 1.0f, 0.0f, 1.0f, 0.0f };
        if ((u + v) > 0.0f)  { G[0] = u / (u + v);   G[1] = v / (u + v); }
        if ((v + ww) > 0.0f) { G[2] = v / (v + ww);  G[3] = ww / (v + ww); }
        if ((ww + u) > 0.0f) { G[4] = ww / (ww + u); G[5] = u / (ww + u); }
        for (int k = 0; k < NSETS; ++k) {
            float B[15];
            const int ds = DS[k], dt = DT[k];
            for (int n = 0; n < 15; ++n) {
                const int i = PI[n], j = PJ[n], kk = 4 - i - j;
                float r;
                if (ds + dt == 0) r = bern(4, i, j, kk, u, v, ww);
                else if (ds + dt == 1)
                    r = 4.0f * ((ds ? bern(3, i - 1, j, kk, u, v, ww) : bern(3, i, j - 1, kk, u, v, ww)) - bern(3, i, j, kk - 1, u, v, ww));
                else if (ds == 2)
                    r = 12.0f * (bern(2, i - 2, j, kk, u, v, ww) - 2.0f * bern(2, i - 1, j, kk - 1, u, v, ww) + bern(2, i, j, kk - 2, u, v, ww));
                else if (dt == 2)
                    r = 12.0f * (bern(2, i, j - 2, kk, u, v, ww) - 2.0f * bern(2, i, j - 1, kk - 1, u, v, ww) + bern(2, i, j, kk - 2, u, v, ww));
                else
                    r = 12.0f * (bern(2, i - 1, j - 1, kk, u, v, ww) - bern(2, i - 1, j, kk - 1, u, v, ww)
                                 - bern(2, i, j - 1, kk - 1, u, v, ww) + bern(2, i, j, kk - 2, u, v, ww));
                B[n] = r;
            }
            for (int i = 0; i < 18; ++i) w[k][i] = (GI[i] < 0) ? B[SRC[i]] : B[SRC[i]] * G[GI[i]];
        }
        return 18;
    }
}

// --------------------------------------------------------------------------------------- kernel --
template <int LT, int ORDER, typename CV>
B200_HD void eval_patch_type(const CV &cv, int type, float s, float t, int boundary, float d1, float sign,
                             float (&out)[ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6)][LT]) {
    constexpr int NSETS = ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6);
    if (type == PT_REGULAR) {
        eval_regular<LT, ORDER>(cv, s, t, boundary, d1, out);
    } else if (type == PT_GREGORY_BASIS) {
        eval_gregory<LT, ORDER>(cv, s, t, d1, out);
    } else if (type == PT_QUADS) {
        eval_quads<LT, ORDER>(cv, s, t, d1, out);
    } else if (type == PT_LOOP || type == PT_GREGORY_TRIANGLE || type == PT_TRIANGLES) {
        float w[NSETS][20];
        const int np = tri_weights<ORDER>(type, s, t, boundary, w);
        const float d2 = sign * d1 * d1;     // osd/patchBasis.h:1598: d2Scale = derivSign * d1Scale * d1Scale
        for (int j = 0; j < np; ++j) {
            float v[LT];
            cv.template load<LT>(j, v);
#pragma unroll
            for (int k = 0; k < NSETS; ++k) {
                const float wk = w[k][j] * (k == 0 ? 1.0f : (k < 3 ? d1 : d2));
#pragma unroll
                for (int c = 0; c < LT; ++c) out[k][c] = fmaf(wk, v[c], out[k][c]);
            }
        }
    }
    // unknown descriptor: the reference evaluates zero points, i.e. writes zeros
}

// MODE: 0 = control points through the index buffer; 1 = from the hull cache, read directly; 2 = from the hull cache,
// staged through shared memory: the warp copies the hulls of its distinct patches cooperatively (consecutive lanes read
// consecutive 16-byte rows: 2 fully used 128-byte lines per 16-point hull instead of one line touch per lane and
// point), then every lane evaluates out of its patch's row -- for incoherent coordinates ~3.5x fewer L1 wavefronts;
// 3 = per warp: staged when the warp touches more than io.stageThreshold distinct patches, direct otherwise (coherent
// warps read the same rows: broadcast loads are cheaper than staging).  See DESIGN.md 4.3.
template <int LT, int ORDER, int MODE>
B200_HD void patch_eval_coord(const PatchIO &io, int i, bool live) {
    constexpr int NSETS = ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6);
    float out[NSETS][LT];
#pragma unroll
    for (int k = 0; k < NSETS; ++k)
#pragma unroll
        for (int c = 0; c < LT; ++c) out[k][c] = 0.0f;

    int arrayIndex = 0, patchIndex = 0;
    float s = 0.0f, t = 0.0f;
    if (live) {
        const int *cw = reinterpret_cast<const int *>(io.coords + i);
        arrayIndex = ld_coord_word(cw + 0);
        patchIndex = ld_coord_word(cw + 1);
        s = int_as_float(ld_coord_word(cw + 3));
        t = int_as_float(ld_coord_word(cw + 4));
        // arrayIndex < 0 marks a sample that hit no patch (b200osd_patch_map_find writes it for holes, where
        // Far::PatchMap::FindPatch returns NULL and the reference's callers skip the sample): its outputs stay untouched
        live = arrayIndex >= 0;
    }
    int type = 0, boundary = 0;
    float d1 = 1.0f, sign = 1.0f;
    const int *aw = nullptr;
    if (live) {
        aw = reinterpret_cast<const int *>(io.arrays + arrayIndex);
        const int regDesc = ldg_i(aw + 0), irrDesc = ldg_i(aw + 1);
        const unsigned field1 = ldg_u(&io.params[patchIndex].field1);

        const int depth = (int)(field1 & 0xfu);
        const int nonquad = (int)((field1 >> 4) & 1u);
        const bool regular = ((field1 >> 5) & 1u) != 0;
        boundary = (int)((field1 >> 7) & 0x1fu);
        const int pv = (int)((field1 >> 12) & 0x3ffu), pu = (int)((field1 >> 22) & 0x3ffu);
        type = regular ? regDesc : irrDesc;

        const float fracInv = (float)(1 << (depth - nonquad));
        const bool isTri = (type == PT_LOOP || type == PT_GREGORY_TRIANGLE || type == PT_TRIANGLES);
        if (isTri && (pu + pv) >= (1 << depth)) {
            const int df = 1 << depth;
            s = (float)(df - pu) - s * fracInv;
            t = (float)(df - pv) - t * fracInv;
            sign = -1.0f;
        } else {
            s = fmaf(s, fracInv, -(float)pu);
            t = fmaf(t, fracInv, -(float)pv);
        }
        d1 = sign * (float)(1 << depth);
    }

    if (MODE >= 2) {
#ifdef __CUDA_ARCH__
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int mine = live ? patchIndex : -1;
        int points = 0;
        size_t row0 = 0;
        if (live) row0 = hull_first_row(io.hullRowsBefore, arrayIndex, aw, patchIndex, io.hullTiles, io.tile, &points);
        bool staged = true;
        if (MODE == 3) {
            // per warp: runs of equal patch indices (>= the number of distinct patches).  Few runs: the lanes read the
            // same rows and broadcast loads straight from the hull cache are cheaper than staging.
            const int prev = __shfl_up_sync(0xffffffffu, mine, 1);
            const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || mine != prev);
            staged = __popc(heads) > io.stageThreshold;
        }
        if (staged) {
            // lanes evaluating the same patch share one staged copy: the lowest such lane (the owner) publishes the row
            const unsigned same = __match_any_sync(0xffffffffu, mine);
            const int owner = __ffs(same) - 1;
            const int np = (owner == lane && live) ? points : 0;         // rows this lane's hull contributes (0: not an owner)
            float *hs = b200_patch_smem + (size_t)warp * (size_t)io.warpWords;
            const int pitch = io.hullPitch;
            // rows 0..15: pass `it` copies hulls it and it+16, one per half warp (256 contiguous bytes each); rows 16
            // apart are 16 banks apart (odd pitch), so the two halves' shared-memory stores do not collide
            const int j = lane & 15, hsel = lane & 16;
#pragma unroll 4
            for (int it = 0; it < 16; ++it) {
                const int h = it + hsel;
                const int n = __shfl_sync(0xffffffffu, np, h);
                const unsigned long long r = __shfl_sync(0xffffffffu, (unsigned long long)row0, h);
                if (j < n) {
                    const float4 v = ld_stream_f4(io.hull4 + r + j);
                    float *d = hs + h * pitch + j * LT;
                    d[0] = v.x;
                    if (LT > 1) d[1] = v.y;
                    if (LT > 2) d[2] = v.z;
                    if (LT > 3) d[3] = v.w;
                }
            }
            // rows 16..: only the few hulls that have them (Gregory end caps), one hull per pass
            unsigned big = __ballot_sync(0xffffffffu, np > 16);
            while (big) {
                const int h = __ffs(big) - 1;
                big &= big - 1;
                const int n = __shfl_sync(0xffffffffu, np, h);
                const unsigned long long r = __shfl_sync(0xffffffffu, (unsigned long long)row0, h);
                const int jj = 16 + lane;
                if (jj < n) {
                    const float4 v = ld_stream_f4(io.hull4 + r + jj);
                    float *d = hs + h * pitch + jj * LT;
                    d[0] = v.x;
                    if (LT > 1) d[1] = v.y;
                    if (LT > 2) d[2] = v.z;
                    if (LT > 3) d[3] = v.w;
                }
            }
            __syncwarp();
            if (live) {
                CvSmem cv;
                cv.row = hs + owner * pitch;
                eval_patch_type<LT, ORDER>(cv, type, s, t, boundary, d1, sign, out);
            }
        } else if (MODE == 3 && live) {
            CvHull cv;
            cv.base = io.hull4 + row0;
            eval_patch_type<LT, ORDER>(cv, type, s, t, boundary, d1, sign, out);
        }
#endif
    } else if (live) {
        if (MODE == 1) {
            // hull cache: the patch's control points as compact 16-byte rows (8 sectors per coordinate instead of 18
            // scattered ones for incoherent coordinates; broadcast reads for coherent ones)
            int points;
            CvHull cv;
            cv.base = io.hull4 + hull_first_row(io.hullRowsBefore, arrayIndex, aw, patchIndex, io.hullTiles, io.tile, &points);
            eval_patch_type<LT, ORDER>(cv, type, s, t, boundary, d1, sign, out);
        } else {
            const int indexBase = ldg_i(aw + 3), stride = ldg_i(aw + 4), primBase = ldg_i(aw + 5);
            CvIndirect cv;
            cv.src = io.src;
            cv.stride = io.srcStride;
            cv.cvs = io.indices + indexBase + stride * (patchIndex - primBase);
            eval_patch_type<LT, ORDER>(cv, type, s, t, boundary, d1, sign, out);
        }
    }
    store_outputs<LT, NSETS>(io, i, live, out);
}

#ifdef __CUDACC__
template <int LT, int ORDER, int MODE>
__global__ void __launch_bounds__(kPatchBlock) patch_kernel(PatchIO io) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i - (int)(threadIdx.x & 31) >= io.n) return;          // whole warp past the end
    patch_eval_coord<LT, ORDER, MODE>(io, i, i < io.n);
}

// Hull cache fill, one launch per patch array: thread r copies control point r of the array's index list (coalesced
// index reads, every lane busy whatever the patch size) into its 16-byte row of the per-array layout described at
// PatchIO::hull4; blockIdx.y = component tile.
__global__ void __launch_bounds__(256) hull_gather_kernel(const float *src, int srcStride, int L, const int *arrayIndices,
                                                          long long rows, int stride, long long rowsBefore, int hullTiles,
                                                          float4 *hull4) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int tile = blockIdx.y;
    if (r >= rows) return;
    const int cv = ld_stream_i1(arrayIndices + r);
    const float *g = src + (size_t)cv * (size_t)srcStride + 4 * tile;
    const int rem = L - 4 * tile;
    float4 v;
    v.x = __ldg(g);
    v.y = rem > 1 ? __ldg(g + 1) : 0.0f;
    v.z = rem > 2 ? __ldg(g + 2) : 0.0f;
    v.w = rem > 3 ? __ldg(g + 3) : 0.0f;
    size_t dst = (size_t)(rowsBefore + r);
    if (hullTiles > 1) {
        const long long q = r / stride;
        dst = (size_t)(rowsBefore + q * stride) * (size_t)hullTiles + (size_t)tile * (size_t)stride + (size_t)(r - q * stride);
    }
    hull4[dst] = v;
}
#endif

}  // namespace b200osd
