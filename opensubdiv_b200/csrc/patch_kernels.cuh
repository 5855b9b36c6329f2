// patch_kernels.cuh -- device code of the limit-evaluation path (EvalPatches*), sm_100a.
//
// Semantics restated from the reference (paths relative to /root/reference/opensubdiv):
//   osd/cpuEvaluator.cpp:157-381   per PatchCoord: array/param lookup, basis weights, gather-sum of control points
//   osd/patchBasisTypes.h:241-426  patch type ids, PatchParam bit fields, (s,t) normalisation into the sub-patch
//   osd/patchBasis.h:53-1610       linear / B-spline / Gregory / box-spline / Gregory-triangle bases, boundary
//                                  folding, derivative scaling (d1 = +-2^depth, d2 = sign*d1*d1)
//
// Design (one thread per PatchCoord, one warp per 32 consecutive coordinates, everything in registers):
//   * A warp first finds the DISTINCT patches among its 32 coordinates (match.any) and copies each distinct control
//     hull ONCE from the index buffer + the caller's primvar buffer into shared memory, two hulls per pass (one per
//     half warp: a coalesced read of the patch's 16-20 indices, then one vertex per lane).  Every lane then evaluates
//     out of its patch's staged row (128-bit shared-memory loads, broadcast when lanes share a patch).  Coordinates
//     that arrive grouped by patch -- sorted by the caller, or binned on the device, see below -- therefore cost one
//     hull fetch per RUN of equal patches instead of one per coordinate.  No per-call cache in global memory.
//   * Device-side binning (SURVEY.md section 7 "bin/sort once per coord set, or a per-call device counting sort"):
//     bin_* kernels build a permutation that groups the coordinates by patch (counting sort on patchIndex with one
//     atomic per coordinate); patch_run_kernel then walks the permutation, gathers its coordinates and writes every
//     result at the CALLER's index i, so outputs are bit-identical per index to the unbinned evaluation.  A sampling
//     probe leaves already coherent coordinate sets alone.
//   * REGULAR (bicubic B-spline, the overwhelmingly common type) is evaluated in SEPARABLE form: the 1-D weight
//     vectors in s and t (value / 1st / 2nd derivative, boundary-folded and pre-scaled) are kept, each of the 4 rows
//     of control points is first contracted with the three s-vectors, and the row results are then combined with
//     the t-vectors.  That is 16*3 + 4*6 = 72 multiply-adds per component instead of the 16*6 = 96 (plus 96
//     products to form the tensor weights) of the reference formulation; boundary folding is linear in the 1-D
//     weights so it commutes with the tensor product.  Summation order differs from the reference, see DESIGN.md.
//   * GREGORY_BASIS and QUADS are unrolled per control point with compile-time point tables.  The triangle types
//     (LOOP / GREGORY_TRIANGLE / TRIANGLES) live in a separate kernel instantiation (TRI) so that quad tables do not
//     carry their code or stack frame.
#pragma once

#include "common.cuh"

#include <utility>

// Every function below is __host__ __device__ so that tests/emu can run the *identical* arithmetic on the CPU
// (a test-only numerical harness, never a product path: libb200osd.so only ever launches the __global__ kernels).
#define B200_HD __host__ __device__ __forceinline__
#define B200_HD_NOINLINE __host__ __device__

namespace b200osd {

constexpr int kPatchMaxOut = 6;

struct BinState {                     // device-side state of one call (probe and bin_* kernels)
    int mode;                         // kPatchMode*: caller's order / per-call hull cache / grouped order (perm is valid)
    int heads, lanes, ticket;         // coherence probe: runs of equal patches / coordinates sampled / warps done
};

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
    int packed;                       // outputs form one contiguous NSETS*LT-float record per coordinate (see store_outputs)
    int vecStore;                     // packed records may leave as 16-byte (contiguous warp block) / 8-byte (binned) stores
    // binned order (patch_run_kernel): position j evaluates coordinate perm[j] when *binState says so
    const int *perm;
    const BinState *binState;
    // shared memory: one private region of warpWords floats per warp (coordinates in, staged hulls, results out)
    int warpWords;
    int coordWords;                   // offset (floats) of the warp's 160-word coordinate prefetch buffer inside its region
    int hullPitch;                    // floats between staged hulls (see hull_pitch / packed_pitch)
    int options;                      // kPatchOptGregoryTrueDerivatives
};

// PatchIO::options / b200osd_patch_table_set_options.  Bit 0: the 8 interior points of GREGORY_BASIS patches get the true
// derivative weights (quotient + product rule), what the reference computes when it is built with
// OPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES (osd/patchBasis.h:441-487); default is its approximation (:421-440).
enum { kPatchOptGregoryTrueDerivatives = 1 };

enum { PT_QUADS = 3, PT_TRIANGLES = 4, PT_LOOP = 5, PT_REGULAR = 6, PT_GREGORY_BASIS = 9, PT_GREGORY_TRIANGLE = 10 };

// Quartic box-spline basis of the regular Loop patch: 12 bivariate quartics, coefficients x12 on the monomials
//   1 s t s^2 st t^2 s^3 s^2t st^2 t^3 s^4 s^3t s^2t^2 st^3 t^4     (osd/patchBasis.h:557-572 in expanded form).
// The five derivative tables follow by differentiating monomial by monomial, all at compile time: the kernels see
// every coefficient as an immediate operand and zero coefficients cost nothing.
struct BoxTables { int c[6][12][15]; };
constexpr BoxTables make_box_tables() {
    constexpr int base[12][15] = {
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
    constexpr int A[15] = { 0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 4, 3, 2, 1, 0 };      // power of s of monomial m
    constexpr int B[15] = { 0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4 };      // power of t
    constexpr int das[6] = { 0, 1, 0, 2, 1, 0 }, dbs[6] = { 0, 0, 1, 0, 1, 2 }; // value, ds, dt, dss, dst, dtt
    constexpr int divisor[6] = { 1, 2, 2, 12, 6, 12 };                          // folded into box_scale()
    BoxTables t{};
    for (int k = 0; k < 6; ++k)
        for (int i = 0; i < 12; ++i)
            for (int m = 0; m < 15; ++m) {
                int a = A[m], b = B[m], c = base[i][m];
                if (c == 0 || a < das[k] || b < dbs[k]) continue;
                for (int q = 0; q < das[k]; ++q) c *= (a - q);
                for (int q = 0; q < dbs[k]; ++q) c *= (b - q);
                for (int mm = 0; mm < 15; ++mm)
                    if (A[mm] == a - das[k] && B[mm] == b - dbs[k]) t.c[k][i][mm] += c / divisor[k];
            }
    return t;
}
template <int K, int I, int M>
struct BoxC { static constexpr int v = make_box_tables().c[K][I][M]; };
template <int K>
B200_HD constexpr float box_scale() {
    return K == 0 ? 1.0f / 12.0f : (K == 1 || K == 2 ? 1.0f / 6.0f : (K == 4 ? 0.5f : 1.0f));
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

// Control-point access of one patch.  load(j): point j; load_row(i): points 4i..4i+3 (one row of a 4 x 4 hull).
// Either through the patch's index list into the caller's primvar buffer (host emulation in tests/emu, and the
// reference formulation) ...
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
    template <int LT>
    B200_HD void load_row(int i, float (&v)[4][LT]) const {
#pragma unroll
        for (int j = 0; j < 4; ++j) load<LT>(4 * i + j, v[j]);
    }
};

// ... or from the warp's shared-memory copy of the hull (patch_run_kernel): point j is one aligned unit of
// hull_unit(LT) floats -- 16 bytes for 3 or 4 components (one 128-bit load), 8 for 2, 4 for 1 ...
__host__ __device__ constexpr int hull_unit(int LT) { return LT >= 3 ? 4 : LT; }
// floats between staged hulls: an odd number of units, so that the rows of different hulls start in different banks
B200_HD int hull_pitch(int LT, int maxPoints) { return hull_unit(LT) * (maxPoints | 1); }

struct CvStaged {
    const float *row;
    template <int LT>
    B200_HD void load(int j, float (&v)[LT]) const {
        if constexpr (LT >= 3) {
            const float4 t = *reinterpret_cast<const float4 *>(row + 4 * j);
            v[0] = t.x; v[1] = t.y; v[2] = t.z;
            if constexpr (LT > 3) v[3] = t.w;
        } else if constexpr (LT == 2) {
            const float2 t = *reinterpret_cast<const float2 *>(row + 2 * j);
            v[0] = t.x; v[1] = t.y;
        } else {
            v[0] = row[j];
        }
    }
    template <int LT>
    B200_HD void load_row(int i, float (&v)[4][LT]) const {
#pragma unroll
        for (int j = 0; j < 4; ++j) load<LT>(4 * i + j, v[j]);
    }
};

// ... or from the per-call hull cache (patch_hull_kernel): the index buffer dereferenced once per call, LT floats per
// point, a patch's points contiguous and its first point 16-byte aligned whenever its hull is (16- and 20-point hulls of
// 1-4 floats are).  A row of a 4 x 4 hull is 4*LT contiguous floats: LT 128-bit loads.

struct CvPacked {
    const float *base;             // first point of the patch in the hull cache
    bool aligned;                  // base is 16-byte aligned
    template <int LT>
    B200_HD void load(int j, float (&v)[LT]) const {
#pragma unroll
        for (int c = 0; c < LT; ++c) v[c] = ldg_f(base + j * LT + c);
    }
    template <int LT>
    B200_HD void load_row(int i, float (&v)[4][LT]) const {
#ifdef __CUDA_ARCH__
        if (aligned) {
            float f[4 * LT];
#pragma unroll
            for (int q = 0; q < LT; ++q) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(base + 4 * LT * i) + q);
                f[4 * q + 0] = t.x; f[4 * q + 1] = t.y; f[4 * q + 2] = t.z; f[4 * q + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < LT; ++c) v[j][c] = f[j * LT + c];
            return;
        }
#endif
#pragma unroll
        for (int j = 0; j < 4; ++j) load<LT>(4 * i + j, v[j]);
    }
};

#ifdef __CUDACC__
extern __shared__ __align__(16) float b200_patch_smem[];
#endif

constexpr int kPatchBlock = 128;

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
        float vv[4][LT];
        cv.template load_row<LT>(i, vv);                          // points 4i .. 4i+3
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int c = 0; c < LT; ++c) {
                r0[c] = fmaf(bs[j], vv[j][c], r0[c]);
                if (ORDER >= 1) r1[c] = fmaf(ds[j], vv[j][c], r1[c]);
                if (ORDER >= 2) r2[c] = fmaf(dss[j], vv[j][c], r2[c]);
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
B200_HD void eval_gregory(const CV &cv, float s, float t, float d1, bool trueDerivatives,
                          float (&out)[ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6)][LT]) {
    constexpr int COL[20] = { 0, 1, 0, 1, 1, 3, 3, 2, 2, 2, 3, 2, 3, 2, 2, 0, 0, 1, 1, 1 };
    constexpr int ROW[20] = { 0, 0, 1, 1, 1, 0, 1, 0, 1, 1, 3, 3, 2, 2, 2, 3, 2, 3, 2, 2 };
    float bs[4], ds[4], dss[4], bt[4], dt[4], dtt[4];
    bezier_1d<ORDER>(s, bs, ds, dss);
    bezier_1d<ORDER>(t, bt, dt, dtt);
    const float sc = 1.0f - s, tc = 1.0f - t;
    float G[8], R[4];
    {
        const float a[4] = { s, t, sc, tc };
        const float den[4] = { s + t, sc + t, sc + tc, s + tc };
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float r = (den[c] <= 0.0f) ? 1.0f : rcp_rn(den[c]);
            G[2 * c] = a[c] * r;
            G[2 * c + 1] = 1.0f - G[2 * c];
            R[c] = r;
        }
    }
    const float d2 = d1 * d1;
    if (ORDER >= 1) {
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
        if (ORDER >= 1 && p >= 3 && trueDerivatives) {
            // G = N / D with constant N', D' (osd/patchBasis.h:441-487); ds.. are already scaled by d1, so G' takes d1 too
            constexpr float NDS[8] = { 1.0f, 0.0f, 0.0f, -1.0f, -1.0f, 0.0f, 0.0f, 1.0f };
            constexpr float NDT[8] = { 0.0f, 1.0f, 1.0f, 0.0f, 0.0f, -1.0f, -1.0f, 0.0f };
            constexpr float DDS[8] = { 1.0f, 1.0f, -1.0f, -1.0f, -1.0f, -1.0f, 1.0f, 1.0f };
            constexpr float DDT[8] = { 1.0f, 1.0f, 1.0f, 1.0f, -1.0f, -1.0f, -1.0f, -1.0f };
            const int k = 2 * (i / 5) + (p - 3);
            const float D = R[i / 5];
            const float g_s = (NDS[k] - DDS[k] * g) * D * d1, g_t = (NDT[k] - DDT[k] * g) * D * d1;
            const float ws = ds[col] * g + bs[col] * g_s;          // d/ds (Bs G)
            const float wt = dt[row] * g + bt[row] * g_t;          // d/dt (Bt G)
            w[1] = ws * bt[row];
            w[2] = wt * bs[col];
            if (ORDER >= 2) {
                const float invD2 = D * D * d2;
                const float g_ss = 2.0f * DDS[k] * invD2 * (g * DDS[k] - NDS[k]);
                const float g_st = invD2 * (2.0f * g * DDS[k] * DDT[k] - NDS[k] * DDT[k] - NDT[k] * DDS[k]);
                const float g_tt = 2.0f * DDT[k] * invD2 * (g * DDT[k] - NDT[k]);
                w[3] = (dss[col] * g + 2.0f * ds[col] * g_s + bs[col] * g_ss) * bt[row];
                w[4] = bt[row] * (bs[col] * g_st + ds[col] * g_t) + dt[row] * ws;
                w[5] = (dtt[row] * g + 2.0f * dt[row] * g_t + bt[row] * g_tt) * bs[col];
            }
        } else {
            if (ORDER >= 1) { w[1] = ds[col] * g * gt; w[2] = gs * dt[row]; }
            if (ORDER >= 2) { w[3] = dss[col] * g * gt; w[4] = ds[col] * g * dt[row]; w[5] = gs * dtt[row]; }
        }
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

// ------------------------------------------------------------------------------- triangle bases --
// Weights live in registers: one derivative set at a time (12 or 18 floats), every index a compile-time constant.
B200_HD void refl(float *w, int phantom, int plus0, int plus1, int minus) {
    const float v = w[phantom];
    w[plus0] += v;
    w[plus1] += v;
    w[minus] -= v;
}

// Box-spline boundary folding (osd/patchBasis.h:663-886): every phantom point is a reflection B + (B' - I).
B200_HD void box_fold_boundary(int mask, float (&w)[12]) {
    constexpr int PH[3][3] = { { 0, 1, 2 }, { 6, 9, 11 }, { 10, 7, 3 } };
    constexpr int B1[3] = { 4, 5, 8 }, B2[3] = { 5, 8, 4 }, I1[3] = { 8, 4, 5 };
    constexpr int B0[3] = { 3, 2, 11 }, I0[3] = { 7, 1, 9 };
    constexpr int B3[3] = { 6, 10, 0 }, I2[3] = { 9, 7, 1 };
    constexpr int VP[3][2] = { { 3, 0 }, { 2, 6 }, { 11, 10 } };
    constexpr int VB0[3] = { 7, 1, 9 }, VI0[3] = { 8, 4, 5 };
    constexpr int VB2[3] = { 1, 9, 7 }, VI1[3] = { 5, 8, 4 };
    const int upper = (mask >> 3) & 3;
    int ebits = mask & 7, vbits = 0;
    if (upper == 1) { vbits = ebits; ebits = 0; }
    else if (upper == 2) { vbits = ((ebits & 1) << 2) | (ebits >> 1); }
#pragma unroll
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
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        if (!(vbits & (1 << v))) continue;
        refl(w, VP[v][0], B1[v], VB0[v], VI0[v]);
        refl(w, VP[v][1], B1[v], VB2[v], VI1[v]);
        w[VP[v][0]] = 0.0f; w[VP[v][1]] = 0.0f;
    }
}

// the 15 monomials of kBox12's columns
B200_HD void box_monomials(float s, float t, float (&M)[15]) {
    M[0] = 1.0f; M[1] = s; M[2] = t;
    M[3] = s * s; M[4] = s * t; M[5] = t * t;
    M[6] = M[3] * s; M[7] = M[4] * s; M[8] = M[4] * t; M[9] = M[5] * t;
    M[10] = M[6] * s; M[11] = M[7] * s; M[12] = M[3] * M[5]; M[13] = M[8] * t; M[14] = M[9] * t;
}

template <int K, int I, int... Ms>
B200_HD float box_row(const float (&M)[15], std::integer_sequence<int, Ms...>) {
    float acc = 0.0f;
    ((acc = (BoxC<K, I, Ms>::v != 0) ? fmaf((float)BoxC<K, I, Ms>::v, M[Ms], acc) : acc), ...);
    return acc;
}

// Weights of derivative set K (0 value, 1 ds, 2 dt, 3 dss, 4 dst, 5 dtt) of the 12-point box spline, boundary folded.
template <int K, int... Is>
B200_HD void box_set_seq(const float (&M)[15], float (&w)[12], std::integer_sequence<int, Is...>) {
    ((w[Is] = box_scale<K>() * box_row<K, Is>(M, std::make_integer_sequence<int, 15>{})), ...);
}
template <int K>
B200_HD void loop_weight_set(const float (&M)[15], int boundary, float (&w)[12]) {
    box_set_seq<K>(M, w, std::make_integer_sequence<int, 12>{});
    if (boundary) box_fold_boundary(boundary, w);
}

// n! / (i! j! k!) u^i v^j w^k, zero outside the triangle (osd/patchBasis.h:1006-1125 in closed form)
template <int N, int I, int J, int K>
B200_HD float bern(float u, float v, float w) {
    if constexpr (I < 0 || J < 0 || K < 0) {
        return 0.0f;
    } else {
        constexpr float fact[5] = { 1.0f, 1.0f, 2.0f, 6.0f, 24.0f };
        float r = fact[N] / (fact[I] * fact[J] * fact[K]);
#pragma unroll
        for (int q = 0; q < I; ++q) r *= u;
#pragma unroll
        for (int q = 0; q < J; ++q) r *= v;
#pragma unroll
        for (int q = 0; q < K; ++q) r *= w;
        return r;
    }
}

// quartic triangular Bernstein function n (row-major over (i,j), k = 4-i-j) differentiated per set K
constexpr int kGtI[15] = { 0, 1, 2, 3, 4, 0, 1, 2, 3, 0, 1, 2, 0, 1, 0 };
constexpr int kGtJ[15] = { 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 4 };
// the 18 Gregory-triangle points: which Bernstein function feeds each, and which rational blend scales it (-1: none)
constexpr int kGtSrc[18] = { 0, 1, 5, 6, 6, 4, 8, 3, 7, 7, 14, 12, 13, 10, 10, 2, 11, 9 };
constexpr int kGtBlend[18] = { -1, -1, -1, 0, 1, -1, -1, -1, 2, 3, -1, -1, -1, 4, 5, -1, -1, -1 };

template <int K, int N>
B200_HD float gt_basis(float u, float v, float w) {
    constexpr int i = kGtI[N], j = kGtJ[N], k = 4 - i - j;
    if constexpr (K == 0) {
        return bern<4, i, j, k>(u, v, w);
    } else if constexpr (K == 1) {
        return 4.0f * (bern<3, i - 1, j, k>(u, v, w) - bern<3, i, j, k - 1>(u, v, w));
    } else if constexpr (K == 2) {
        return 4.0f * (bern<3, i, j - 1, k>(u, v, w) - bern<3, i, j, k - 1>(u, v, w));
    } else if constexpr (K == 3) {
        return 12.0f * (bern<2, i - 2, j, k>(u, v, w) - 2.0f * bern<2, i - 1, j, k - 1>(u, v, w) + bern<2, i, j, k - 2>(u, v, w));
    } else if constexpr (K == 4) {
        return 12.0f * (bern<2, i - 1, j - 1, k>(u, v, w) - bern<2, i - 1, j, k - 1>(u, v, w)
                        - bern<2, i, j - 1, k - 1>(u, v, w) + bern<2, i, j, k - 2>(u, v, w));
    } else {
        return 12.0f * (bern<2, i, j - 2, k>(u, v, w) - 2.0f * bern<2, i, j - 1, k - 1>(u, v, w) + bern<2, i, j, k - 2>(u, v, w));
    }
}

// rational blends of the three interior point pairs (osd/patchBasis.h:1160-1200)
B200_HD void gt_blends(float u, float v, float ww, float (&G)[6]) {
    G[0] = 1.0f; G[1] = 0.0f; G[2] = 1.0f; G[3] = 0.0f; G[4] = 1.0f; G[5] = 0.0f;
    if ((u + v) > 0.0f)  { G[0] = u / (u + v);   G[1] = v / (u + v); }
    if ((v + ww) > 0.0f) { G[2] = v / (v + ww);  G[3] = ww / (v + ww); }
    if ((ww + u) > 0.0f) { G[4] = ww / (ww + u); G[5] = u / (ww + u); }
}

template <int K, int I>
B200_HD float gt_point(float u, float v, float ww, const float (&G)[6]) {
    constexpr int src = kGtSrc[I], blend = kGtBlend[I];
    if constexpr (blend < 0) return gt_basis<K, src>(u, v, ww);
    else return gt_basis<K, src>(u, v, ww) * G[blend];
}
template <int K, int... Is>
B200_HD void gt_set_seq(float u, float v, float ww, const float (&G)[6], float (&w)[18], std::integer_sequence<int, Is...>) {
    ((w[Is] = gt_point<K, Is>(u, v, ww, G)), ...);
}
template <int K>
B200_HD void gregory_tri_weight_set(float u, float v, float ww, const float (&G)[6], float (&w)[18]) {
    gt_set_seq<K>(u, v, ww, G, w, std::make_integer_sequence<int, 18>{});
}

// linear triangle: derivative set K of the 3 weights
template <int K>
B200_HD void triangle_weight_set(float s, float t, float (&w)[3]) {
    if constexpr (K == 0) { w[0] = 1.0f - s - t; w[1] = s; w[2] = t; }
    else if constexpr (K == 1) { w[0] = -1.0f; w[1] = 1.0f; w[2] = 0.0f; }
    else if constexpr (K == 2) { w[0] = -1.0f; w[1] = 0.0f; w[2] = 1.0f; }
    else { w[0] = 0.0f; w[1] = 0.0f; w[2] = 0.0f; }
}

// One derivative set of one triangle type, into registers.  NP = 12 (LOOP), 18 (GREGORY_TRIANGLE), 3 (TRIANGLES).
struct TriParams {
    float s, t, ww;
    int boundary;
    float M[15];     // LOOP
    float G[6];      // GREGORY_TRIANGLE
};
template <int TYPE, int K, int NP>
B200_HD void tri_weight_set(const TriParams &tp, float (&w)[NP]) {
    if constexpr (TYPE == PT_LOOP) loop_weight_set<K>(tp.M, tp.boundary, w);
    else if constexpr (TYPE == PT_GREGORY_TRIANGLE) gregory_tri_weight_set<K>(tp.s, tp.t, tp.ww, tp.G, w);
    else triangle_weight_set<K>(tp.s, tp.t, w);
}
template <int TYPE>
B200_HD void tri_prepare(float s, float t, int boundary, TriParams &tp) {
    tp.s = s; tp.t = t; tp.ww = 1.0f - s - t; tp.boundary = boundary;
    if constexpr (TYPE == PT_LOOP) box_monomials(s, t, tp.M);
    if constexpr (TYPE == PT_GREGORY_TRIANGLE) gt_blends(s, t, tp.ww, tp.G);
}
template <int TYPE>
B200_HD constexpr int tri_points() { return TYPE == PT_LOOP ? 12 : (TYPE == PT_GREGORY_TRIANGLE ? 18 : 3); }

// out[K] += sum_j (w_j * scale) cv_j for set K; recurses over the NSETS sets so K stays a compile-time constant.
template <int LT, int NSETS, int TYPE, int K, typename CV>
B200_HD void tri_accumulate(const CV &cv, const TriParams &tp, float d1, float d2, float (&out)[NSETS][LT]) {
    if constexpr (K < NSETS) {
        constexpr int NP = tri_points<TYPE>();
        float w[NP];
        tri_weight_set<TYPE, K, NP>(tp, w);
        const float scale = K == 0 ? 1.0f : (K < 3 ? d1 : d2);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            float v[LT];
            cv.template load<LT>(j, v);
            const float wk = w[j] * scale;
#pragma unroll
            for (int c = 0; c < LT; ++c) out[K][c] = fmaf(wk, v[c], out[K][c]);
        }
        tri_accumulate<LT, NSETS, TYPE, K + 1>(cv, tp, d1, d2, out);
    }
}
template <int LT, int ORDER, int TYPE, typename CV>
B200_HD void eval_triangle_type(const CV &cv, float s, float t, int boundary, float d1, float sign,
                                float (&out)[ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6)][LT]) {
    constexpr int NSETS = ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6);
    TriParams tp;
    tri_prepare<TYPE>(s, t, boundary, tp);
    const float d2 = sign * d1 * d1;     // osd/patchBasis.h:1598: d2Scale = derivSign * d1Scale * d1Scale
    tri_accumulate<LT, NSETS, TYPE, 0>(cv, tp, d1, d2, out);
}

// Weight arrays w[k][i] of all NSETS sets (the limit-stencil builder wants them side by side); returns the point count.
template <int NSETS, int TYPE, int K>
B200_HD void tri_weights_type(const TriParams &tp, float (*w)[20]) {
    if constexpr (K < NSETS) {
        constexpr int NP = tri_points<TYPE>();
        float wk[NP];
        tri_weight_set<TYPE, K, NP>(tp, wk);
#pragma unroll
        for (int i = 0; i < NP; ++i) w[K][i] = wk[i];
        tri_weights_type<NSETS, TYPE, K + 1>(tp, w);
    }
}
template <int ORDER>
B200_HD_NOINLINE int tri_weights(int type, float s, float t, int boundary, float (*w)[20]) {
    constexpr int NSETS = ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6);
    TriParams tp;
    if (type == PT_TRIANGLES) {
        tri_prepare<PT_TRIANGLES>(s, t, boundary, tp);
        tri_weights_type<NSETS, PT_TRIANGLES, 0>(tp, w);
        return 3;
    }
    if (type == PT_LOOP) {
        tri_prepare<PT_LOOP>(s, t, boundary, tp);
        tri_weights_type<NSETS, PT_LOOP, 0>(tp, w);
        return 12;
    }
    tri_prepare<PT_GREGORY_TRIANGLE>(s, t, boundary, tp);
    tri_weights_type<NSETS, PT_GREGORY_TRIANGLE, 0>(tp, w);
    return 18;
}

// --------------------------------------------------------------------------------------- kernel --
// TRI = true is the general instantiation (128 registers): the triangle types and the optional evaluation modes
// (PatchIO::options).  TRI = false keeps REGULAR / GREGORY_BASIS / QUADS in their default form at 72 registers.
template <int LT, int ORDER, bool TRI, typename CV>
B200_HD void eval_patch_type(const CV &cv, int type, float s, float t, int boundary, float d1, float sign, int options,
                             float (&out)[ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6)][LT]) {
    if (type == PT_REGULAR) {
        eval_regular<LT, ORDER>(cv, s, t, boundary, d1, out);
    } else if (type == PT_GREGORY_BASIS) {
        eval_gregory<LT, ORDER>(cv, s, t, d1, TRI && (options & kPatchOptGregoryTrueDerivatives) != 0, out);
    } else if (type == PT_QUADS) {
        eval_quads<LT, ORDER>(cv, s, t, d1, out);
    } else if (TRI && type == PT_LOOP) {
        eval_triangle_type<LT, ORDER, PT_LOOP>(cv, s, t, boundary, d1, sign, out);
    } else if (TRI && type == PT_GREGORY_TRIANGLE) {
        eval_triangle_type<LT, ORDER, PT_GREGORY_TRIANGLE>(cv, s, t, boundary, d1, sign, out);
    } else if (TRI && type == PT_TRIANGLES) {
        eval_triangle_type<LT, ORDER, PT_TRIANGLES>(cv, s, t, boundary, d1, sign, out);
    }
    // unknown descriptor: the reference evaluates zero points, i.e. writes zeros
}

// number of control points of a patch type (osd/patchBasisTypes.h:241-246; unknown descriptors: none)
B200_HD int patch_type_points(int type) {
    return type == PT_REGULAR ? 16 : (type == PT_GREGORY_BASIS ? 20 : (type == PT_QUADS ? 4 : (type == PT_LOOP ? 12
           : (type == PT_GREGORY_TRIANGLE ? 18 : (type == PT_TRIANGLES ? 3 : 0)))));
}

// What a coordinate needs besides (s,t): decoded from its PatchArray and PatchParam (osd/patchBasisTypes.h:366-426).
struct PatchSite {
    int type, boundary, cvOffset;     // cvOffset: first control-vertex index of the patch in the index buffer
    float s, t, d1, sign;
};

// aw: the six ints of the coordinate's PatchArray, already fetched (registers, shared memory ...)
B200_HD void decode_patch_site_arr(const int *aw, int patchIndex, float s, float t, unsigned field1, PatchSite &ps) {
    const int regDesc = aw[0], irrDesc = aw[1];
    const int indexBase = aw[3], stride = aw[4], primBase = aw[5];
    const int depth = (int)(field1 & 0xfu);
    const int nonquad = (int)((field1 >> 4) & 1u);
    const bool regular = ((field1 >> 5) & 1u) != 0;
    ps.boundary = (int)((field1 >> 7) & 0x1fu);
    const int pv = (int)((field1 >> 12) & 0x3ffu), pu = (int)((field1 >> 22) & 0x3ffu);
    ps.type = regular ? regDesc : irrDesc;
    ps.cvOffset = indexBase + stride * (patchIndex - primBase);
    ps.sign = 1.0f;
    const float fracInv = (float)(1 << (depth - nonquad));
    const bool isTri = (ps.type == PT_LOOP || ps.type == PT_GREGORY_TRIANGLE || ps.type == PT_TRIANGLES);
    if (isTri && (pu + pv) >= (1 << depth)) {
        const int df = 1 << depth;
        ps.s = (float)(df - pu) - s * fracInv;
        ps.t = (float)(df - pv) - t * fracInv;
        ps.sign = -1.0f;
    } else {
        ps.s = fmaf(s, fracInv, -(float)pu);
        ps.t = fmaf(t, fracInv, -(float)pv);
    }
    ps.d1 = ps.sign * (float)(1 << depth);
}

B200_HD void decode_patch_site(const PatchIO &io, int arrayIndex, int patchIndex, float s, float t, PatchSite &ps) {
    const int *g = reinterpret_cast<const int *>(io.arrays + arrayIndex);
    const int aw[6] = { ldg_i(g + 0), ldg_i(g + 1), 0, ldg_i(g + 3), ldg_i(g + 4), ldg_i(g + 5) };
    decode_patch_site_arr(aw, patchIndex, s, t, ldg_u(&io.params[patchIndex].field1), ps);
}

// One coordinate evaluated straight through the index buffer: the reference formulation, used by the host emulation
// of the kernel arithmetic (tests/emu); patch_run_kernel evaluates the same functions out of staged hulls.
template <int LT, int ORDER>
B200_HD void patch_eval_coord(const PatchIO &io, int i) {
    constexpr int NSETS = ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6);
    float out[NSETS][LT];
#pragma unroll
    for (int k = 0; k < NSETS; ++k)
#pragma unroll
        for (int c = 0; c < LT; ++c) out[k][c] = 0.0f;
    const b200osd_patch_coord pc = io.coords[i];
    // arrayIndex < 0 marks a sample that hit no patch (b200osd_patch_map_find writes it for holes, where
    // Far::PatchMap::FindPatch returns NULL and the reference's callers skip the sample): its outputs stay untouched
    if (pc.arrayIndex < 0) return;
    PatchSite ps;
    decode_patch_site(io, pc.arrayIndex, pc.patchIndex, pc.s, pc.t, ps);
    CvIndirect cv;
    cv.src = io.src;
    cv.stride = io.srcStride;
    cv.cvs = io.indices + ps.cvOffset;
    eval_patch_type<LT, ORDER, true>(cv, ps.type, ps.s, ps.t, ps.boundary, ps.d1, ps.sign, io.options, out);
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        float *d = io.dst[k];
        if (!d) continue;
        d += (size_t)i * (size_t)io.dstStride[k];
#pragma unroll
        for (int c = 0; c < LT; ++c) st_out(d + c, out[k][c]);
    }
}

#ifdef __CUDACC__
// Results leave through shared memory.  `ix` holds the caller's index of each lane's coordinate (-1: nothing to write).
//  * packed (all NSETS outputs interleave into ONE record of R = NSETS*LT floats per coordinate, the glEvalLimit layout,
//    examples/glEvalLimit/glEvalLimit.cpp:277-287; also a single tightly packed output): the warp's 32 records are
//    staged at pitch R.  In caller order (contiguous) the whole 32 x R block leaves as 128-bit stores of consecutive
//    addresses; in binned order every record is written by consecutive lanes (64-bit when R is even), so a 72-byte
//    record costs its 3-4 sectors once instead of 18 scattered 4-byte writes;
//  * separate / strided outputs: per output, consecutive lanes write the consecutive floats of a record.
template <int LT, int NSETS>
__device__ __forceinline__ void store_outputs(const PatchIO &io, float *st, int i, bool live, bool contiguous,
                                              const float (&out)[NSETS][LT]) {
    constexpr int R = NSETS * LT;
    const int lane = threadIdx.x & 31;
    int *ix = reinterpret_cast<int *>(st + 32 * R);
    __syncwarp();                                               // the region may still be read as staged hulls
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        if constexpr (LT == 4) {
            *reinterpret_cast<float4 *>(st + lane * R + k * LT) = make_float4(out[k][0], out[k][1], out[k][2], out[k][3]);
        } else {
#pragma unroll
            for (int c = 0; c < LT; ++c) st[lane * R + k * LT + c] = out[k][c];
        }
    }
    ix[lane] = live ? i : -1;
    const unsigned livemask = __ballot_sync(0xffffffffu, live);
    __syncwarp();
    if (io.packed) {                                            // uniform across the grid
        float *d0 = io.dst[0];
        if (contiguous && livemask == 0xffffffffu && (io.vecStore & 1)) {
            float *base = d0 + (size_t)(i - lane) * (size_t)R;
            const float4 *s4 = reinterpret_cast<const float4 *>(st);
#pragma unroll
            for (int q = 0; q < (8 * R + 31) / 32; ++q) {
                const int e = q * 32 + lane;
                if (e < 8 * R) {
                    const float4 v = s4[e];
                    st_stream_f4(base + 4 * e, v.x, v.y, v.z, v.w);
                }
            }
        } else if (R % 2 == 0 && (io.vecStore & 2)) {
            constexpr int R2 = R % 2 == 0 ? R / 2 : 1;
            const float2 *s2 = reinterpret_cast<const float2 *>(st);
#pragma unroll
            for (int q = 0; q < R2; ++q) {
                const int e = q * 32 + lane;                    // float2 index in the staged block
                const int ci = e / R2, c = e - ci * R2;
                const int ii = ix[ci];
                if (ii >= 0) {
                    const float2 v = s2[e];
                    st_stream_f2(d0 + (size_t)ii * (size_t)R + 2 * c, v.x, v.y);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int e = q * 32 + lane;
                const int ci = e / R, c = e - ci * R;
                const int ii = ix[ci];
                if (ii >= 0) st_stream_f1(d0 + (size_t)ii * (size_t)R + c, st[e]);
            }
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        float *d = io.dst[k];
        if (!d) continue;                                       // uniform across the grid
        const size_t stride = (size_t)io.dstStride[k];
#pragma unroll
        for (int q = 0; q < LT; ++q) {
            const int e = q * 32 + lane;
            const int ci = e / LT, c = e - ci * LT;
            const int ii = ix[ci];
            if (ii >= 0) st_stream_f1(d + (size_t)ii * stride + c, st[ci * R + k * LT + c]);
        }
    }
}

// One lane's coordinate: position j of the caller's order, or of the grouped order (perm).  The warp's 32 records are
// 640 contiguous bytes in caller order: five coalesced loads, then a conflict-free transposition through `st`.
struct LaneCoord {
    int i, arrayIndex, patchIndex;
    float s, t;
    bool live;
};

// In caller order the 32 records of a tile are 640 contiguous bytes: they are fetched with asynchronous 4-byte copies
// (cp.async: global -> shared without registers) ONE TILE AHEAD, so that the DRAM latency of the coordinate stream is
// off the warp's critical path; prefetch_tile_coords is called for the next tile once the current one has been read.
__device__ __forceinline__ void prefetch_tile_coords(const PatchIO &io, float *coordBuf, int tile, int tiles, int lane) {
    if (tile < tiles) {
        const long long j0 = (long long)tile << 5;
        const int *g = reinterpret_cast<const int *>(io.coords) + (size_t)j0 * 5;
        const int words = (int)min((long long)32, (long long)io.n - j0) * 5;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(coordBuf);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int e = q * 32 + lane;
            if (e < words) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * e), "l"(g + e) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ LaneCoord read_tile_coords(const PatchIO &io, const float *coordBuf, long long j0, int lane) {
    LaneCoord c;
    const long long j = j0 + lane;
    c.live = j < io.n;
    c.i = (int)j; c.arrayIndex = -1; c.patchIndex = 0; c.s = 0.0f; c.t = 0.0f;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (c.live) {
        const int *cw = reinterpret_cast<const int *>(coordBuf);
        c.arrayIndex = cw[lane * 5 + 0];
        c.patchIndex = cw[lane * 5 + 1];
        c.s = __int_as_float(cw[lane * 5 + 3]);
        c.t = __int_as_float(cw[lane * 5 + 4]);
    }
    __syncwarp();                                               // every lane has its record: the buffer may be refilled
    // arrayIndex < 0 marks a sample that hit no patch (b200osd_patch_map_find writes it for holes, where
    // Far::PatchMap::FindPatch returns NULL and the reference's callers skip the sample): its outputs stay untouched
    c.live = c.live && c.arrayIndex >= 0;
    return c;
}

// grouped order: position j evaluates coordinate perm[j]; a 20-byte record is 1-2 sectors, kept in L1 between the loads
__device__ __forceinline__ LaneCoord gather_lane_coord(const PatchIO &io, long long j0, int lane) {
    LaneCoord c;
    const long long j = j0 + lane;
    c.live = j < io.n;
    c.i = (int)j; c.arrayIndex = -1; c.patchIndex = 0; c.s = 0.0f; c.t = 0.0f;
    if (c.live) {
        c.i = ld_stream_i1(io.perm + j);
        const int *cw = reinterpret_cast<const int *>(io.coords + c.i);
        c.arrayIndex = __ldg(cw + 0);
        c.patchIndex = __ldg(cw + 1);
        c.s = __int_as_float(__ldg(cw + 3));
        c.t = __int_as_float(__ldg(cw + 4));
    }
    c.live = c.live && c.arrayIndex >= 0;
    return c;
}

constexpr int kHullSlots = 8;                                   // distinct hulls staged per round (4 passes of two)
constexpr int kPatchModeDirect = 0, kPatchModeHull = 1, kPatchModeGrouped = 2;

// One tile of 32 coordinates, start to finish: decode, stage the warp's distinct hulls (kHullSlots per round) with
// blocking loads, evaluate, store.  The body of patch_run_kernel.
template <int LT, int ORDER, bool TRI>
__device__ __forceinline__ void patch_tile_sync(const PatchIO &io, float *st, int pitch, const LaneCoord &lc, bool contiguous) {
    constexpr int NSETS = ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6);
    constexpr int LTU = hull_unit(LT);
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int PASSES = kHullSlots / 2;
    const int lane = threadIdx.x & 31;
    const int half = lane >> 4, jj = lane & 15;
    const bool live = lc.live;
    PatchSite ps;
    ps.type = 0; ps.boundary = 0; ps.cvOffset = 0; ps.s = 0.0f; ps.t = 0.0f; ps.d1 = 1.0f; ps.sign = 1.0f;
    if (live) decode_patch_site(io, lc.arrayIndex, lc.patchIndex, lc.s, lc.t, ps);
    const int np = live ? patch_type_points(ps.type) : 0;

    // distinct patches of the warp: the lowest lane of each group of equal patches owns the staged copy
    const int key = live ? lc.patchIndex : (-1 - lane);
    const unsigned same = __match_any_sync(FULL, key);
    const int owner = __ffs(same) - 1;
    const unsigned owners = __ballot_sync(FULL, live && owner == lane);
    const int slot = __popc(owners & ((1u << owner) - 1u));  // dense number of my hull among the warp's hulls

    float out[NSETS][LT];
#pragma unroll
    for (int k = 0; k < NSETS; ++k)
#pragma unroll
        for (int c = 0; c < LT; ++c) out[k][c] = 0.0f;

    // every distinct patch leaves (first index, points) at its slot: the staging lanes below read them from there
    // instead of locating the owner lane and shuffling
    int *slotInfo = reinterpret_cast<int *>(st + io.coordWords + 160);      // [32] first index, [32] points
    if (live && owner == lane) { slotInfo[slot] = ps.cvOffset; slotInfo[32 + slot] = np; }
    const int count = __popc(owners);
    __syncwarp();
#pragma unroll 1
    for (int base = 0; base < count; base += kHullSlots) {
        // this round's hulls, two per pass (one per half warp).  All passes' index loads are issued before the first
        // vertex load: the index -> vertex dependency is paid once per round, not once per pass.
        int cvi[PASSES];
        bool more = false;                                  // some hull of the round has points 16..
#pragma unroll
        for (int q = 0; q < PASSES; ++q) {
            const int hi = base + 2 * q + half;
            const int n_h = hi < count ? slotInfo[32 + hi] : 0;
            more = more || n_h > 16;
            cvi[q] = (jj < n_h) ? ldg_i(io.indices + slotInfo[hi] + jj) : -1;
        }
#pragma unroll
        for (int q = 0; q < PASSES; ++q) {
            if (cvi[q] >= 0) {
                const float *g = io.src + (size_t)cvi[q] * (size_t)io.srcStride;
                float *d = st + (2 * q + half) * pitch + jj * LTU;
                if (LT == 4 && (io.vecStore & 4)) {
                    *reinterpret_cast<float4 *>(d) = __ldg(reinterpret_cast<const float4 *>(g));
                } else if constexpr (LT >= 3) {
                    *reinterpret_cast<float4 *>(d) = make_float4(__ldg(g), __ldg(g + 1), __ldg(g + 2), __ldg(g + LT - 1));
                } else {
#pragma unroll
                    for (int c = 0; c < LT; ++c) d[c] = __ldg(g + c);
                }
            }
        }
        if (__any_sync(FULL, more)) {                       // points 16.. of 18 / 20-point hulls (end caps): rare
#pragma unroll
            for (int q = 0; q < PASSES; ++q) {
                const int hi = base + 2 * q + half;
                const int n_h = hi < count ? slotInfo[32 + hi] : 0;
                const int pnt = jj + 16;
                if (pnt < n_h) {
                    const int ci = ldg_i(io.indices + slotInfo[hi] + pnt);
                    const float *g = io.src + (size_t)ci * (size_t)io.srcStride;
                    float *d = st + (2 * q + half) * pitch + pnt * LTU;
#pragma unroll
                    for (int c = 0; c < LT; ++c) d[c] = __ldg(g + c);
                }
            }
        }
        __syncwarp();
        if (live && slot >= base && slot < base + kHullSlots) {
            CvStaged cv;
            cv.row = st + (slot - base) * pitch;
            eval_patch_type<LT, ORDER, TRI>(cv, ps.type, ps.s, ps.t, ps.boundary, ps.d1, ps.sign, io.options, out);
        }
        __syncwarp();                                       // the next round (or the result staging) overwrites the rows
    }
    store_outputs<LT, NSETS>(io, st, lc.i, live, contiguous, out);
}

// Hulls staged per warp: one warp per 32 coordinates, persistent grid (a block walks tiles with a grid stride).
// Runs in the caller's order, or -- when `perm` is given and the call's state says so -- in the grouped order.
template <int LT, int ORDER, bool TRI>
__global__ void __launch_bounds__(kPatchBlock, TRI ? 4 : 7) patch_run_kernel(PatchIO io) {
    const int mode = io.binState ? io.binState->mode : (io.perm ? kPatchModeGrouped : kPatchModeDirect);   // grid-uniform
    if (mode == kPatchModeHull) return;                         // this call is served by patch_hull_kernel
    const bool grouped = io.perm != nullptr && mode == kPatchModeGrouped;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *st = b200_patch_smem + (size_t)warp * (size_t)io.warpWords;
    const int pitch = io.hullPitch;

    float *coordBuf = st + io.coordWords;
    const int tiles = (int)(((long long)io.n + 31) >> 5);
    const int tileStep = gridDim.x * (kPatchBlock / 32);
    int tile = blockIdx.x * (kPatchBlock / 32) + warp;
    if (!grouped) prefetch_tile_coords(io, coordBuf, tile, tiles, lane);
    for (; tile < tiles; tile += tileStep) {
        LaneCoord lc;
        if (grouped) {
            lc = gather_lane_coord(io, (long long)tile << 5, lane);
        } else {
            lc = read_tile_coords(io, coordBuf, (long long)tile << 5, lane);
            prefetch_tile_coords(io, coordBuf, tile + tileStep, tiles, lane);
        }
        patch_tile_sync<LT, ORDER, TRI>(io, st, pitch, lc, !grouped);
    }
}

// Small hulls (the 3- and 4-point linear patches of varying and linear face-varying data): nothing worth sharing between
// lanes, every lane reads its 3-4 control points straight through the index buffer.
template <int LT, int ORDER, bool TRI>
__global__ void __launch_bounds__(kPatchBlock, 8) patch_direct_kernel(PatchIO io) {
    constexpr int NSETS = ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6);
    const int mode = io.binState ? io.binState->mode : (io.perm ? kPatchModeGrouped : kPatchModeDirect);
    if (mode == kPatchModeHull) return;
    const bool grouped = io.perm != nullptr && mode == kPatchModeGrouped;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *st = b200_patch_smem + (size_t)warp * (size_t)io.warpWords;
    float *coordBuf = st + io.coordWords;
    const int tiles = (int)(((long long)io.n + 31) >> 5);
    const int tileStep = gridDim.x * (kPatchBlock / 32);
    int tile = blockIdx.x * (kPatchBlock / 32) + warp;
    if (!grouped) prefetch_tile_coords(io, coordBuf, tile, tiles, lane);
    for (; tile < tiles; tile += tileStep) {
        LaneCoord lc;
        if (grouped) {
            lc = gather_lane_coord(io, (long long)tile << 5, lane);
        } else {
            lc = read_tile_coords(io, coordBuf, (long long)tile << 5, lane);
            prefetch_tile_coords(io, coordBuf, tile + tileStep, tiles, lane);
        }
        float out[NSETS][LT];
#pragma unroll
        for (int k = 0; k < NSETS; ++k)
#pragma unroll
            for (int c = 0; c < LT; ++c) out[k][c] = 0.0f;
        if (lc.live) {
            PatchSite ps;
            decode_patch_site(io, lc.arrayIndex, lc.patchIndex, lc.s, lc.t, ps);
            CvIndirect cv;
            cv.src = io.src;
            cv.stride = io.srcStride;
            cv.cvs = io.indices + ps.cvOffset;
            eval_patch_type<LT, ORDER, TRI>(cv, ps.type, ps.s, ps.t, ps.boundary, ps.d1, ps.sign, io.options, out);
        }
        store_outputs<LT, NSETS>(io, st, lc.i, lc.live, !grouped, out);
    }
}

// ---- per-call hull cache: for INCOHERENT coordinates (every lane another patch) ------------------------------------
// hull_build_kernel dereferences the index buffer once per call: cache row r = the LT floats of control vertex
// indices[r], so a patch's hull is one contiguous block (192 bytes for 16 xyz points) at its index offset.
// patch_hull_kernel then reads every coordinate's hull with 128-bit loads straight from the cache: ONE random access
// of 192-240 bytes per coordinate instead of 18 scattered ones.  Both return at once unless the call's state (set by
// the coherence probe) asks for them.
template <int LT>
__global__ void __launch_bounds__(256) hull_build_kernel(const float *src, int srcStride, const int *indices, long long rows,
                                                         float *hull, const BinState *state) {
    if (state && state->mode != kPatchModeHull) return;
    // four consecutive rows per thread: one 128-bit index load, LT 128-bit stores
    const long long quads = rows >> 2;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
        const int4 cv = ld_stream_i4(reinterpret_cast<const int4 *>(indices) + q);
        const int c4[4] = { cv.x, cv.y, cv.z, cv.w };
        float f[4 * LT];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float *g = src + (size_t)c4[r] * (size_t)srcStride;
#pragma unroll
            for (int c = 0; c < LT; ++c) f[r * LT + c] = __ldg(g + c);
        }
        float4 *d = reinterpret_cast<float4 *>(hull + (size_t)q * 4 * LT);
#pragma unroll
        for (int k = 0; k < LT; ++k) d[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(rows & 3)) {      // the last 1-3 rows
        const long long r = (quads << 2) + threadIdx.x;
        const float *g = src + (size_t)ld_stream_i1(indices + r) * (size_t)srcStride;
#pragma unroll
        for (int c = 0; c < LT; ++c) hull[(size_t)r * LT + c] = __ldg(g + c);
    }
}

// One lane per coordinate, every lane another patch: each lane reads its hull straight from the cache (LT 128-bit loads
// per row of four points).  Measured alternatives (profiles/r02q_*): copying the warp's 32 hulls cooperatively into
// shared memory first (two full lines per hull instead of one line touch per lane and piece) relieves the L1 wavefront
// pipe (84 % -> lower) but serialises a load -> store -> barrier -> evaluate chain per warp and is 25 % SLOWER.
template <int LT, int ORDER, bool TRI>
__global__ void __launch_bounds__(kPatchBlock, TRI ? 4 : 6) patch_hull_kernel(PatchIO io, const float *hull) {
    constexpr int NSETS = ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6);
    if (io.binState && io.binState->mode != kPatchModeHull) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *st = b200_patch_smem + (size_t)warp * (size_t)io.warpWords;
    float *coordBuf = st + io.coordWords;
    const int tiles = (int)(((long long)io.n + 31) >> 5);
    const int tileStep = gridDim.x * (kPatchBlock / 32);
    int tile = blockIdx.x * (kPatchBlock / 32) + warp;
    prefetch_tile_coords(io, coordBuf, tile, tiles, lane);
    for (; tile < tiles; tile += tileStep) {
        const LaneCoord lc = read_tile_coords(io, coordBuf, (long long)tile << 5, lane);
        prefetch_tile_coords(io, coordBuf, tile + tileStep, tiles, lane);
        float out[NSETS][LT];
#pragma unroll
        for (int k = 0; k < NSETS; ++k)
#pragma unroll
            for (int c = 0; c < LT; ++c) out[k][c] = 0.0f;
        if (lc.live) {
            PatchSite ps;
            decode_patch_site(io, lc.arrayIndex, lc.patchIndex, lc.s, lc.t, ps);
            CvPacked cv;
            cv.base = hull + (size_t)ps.cvOffset * LT;
            cv.aligned = ((size_t)ps.cvOffset * LT) % 4 == 0;   // the cache itself is 256-byte aligned
            eval_patch_type<LT, ORDER, TRI>(cv, ps.type, ps.s, ps.t, ps.boundary, ps.d1, ps.sign, io.options, out);
        }
        store_outputs<LT, NSETS>(io, st, lc.i, lc.live, true, out);
    }
}

// ------------------------------------------------------------------------------------ binning --
// Counting sort of the coordinates by patch (one bin per patch, plus bin numPatches for records that hit no patch).
//   probe   : samples 256 warps' worth of consecutive coordinates; already coherent sets (few distinct patches per
//             warp) are left in the caller's order -- every later bin kernel then returns at once
//   count   : rank of every coordinate inside its bin (one atomic each)
//   scan    : exclusive prefix sums of the bin sizes (per 4096-bin block, then over the block totals)
//   scatter : perm[start[bin] + rank] = i
constexpr int kBinProbeWarps = 256;
constexpr int kScanThreads = 1024, kScanItems = 4, kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(128) bin_probe_kernel(const b200osd_patch_coord *coords, int n, BinState *state,
                                                        int modeIfIncoherent, int force) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int total = (gridDim.x * blockDim.x) >> 5;
    const long long start = ((long long)n * w / total) & ~31LL;
    const long long i = start + lane;
    const bool valid = i < n;
    const int key = valid ? __ldg(reinterpret_cast<const int *>(coords + i) + 1) : -1;
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, valid && (lane == 0 || key != prev));
    const unsigned lanes = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) {
        atomicAdd(&state->heads, __popc(heads));
        atomicAdd(&state->lanes, __popc(lanes));
        __threadfence();
        if (atomicAdd(&state->ticket, 1) == total - 1) {
            __threadfence();
            const int h = atomicAdd(&state->heads, 0), l = atomicAdd(&state->lanes, 0);
            // incoherent: on average, a run of equal patches is shorter than 2 coordinates
            state->mode = (force || 2 * h > l) ? modeIfIncoherent : kPatchModeDirect;
        }
    }
}

__global__ void __launch_bounds__(256) bin_count_kernel(const b200osd_patch_coord *coords, int n, int numPatches,
                                                        const BinState *state, int *count, int2 *keyRank) {
    if (state->mode != kPatchModeGrouped) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int *cw = reinterpret_cast<const int *>(coords + i);
    const int a = __ldg(cw + 0), p = __ldg(cw + 1);
    const int bin = (a < 0 || p < 0 || p >= numPatches) ? numPatches : p;
    const int r = atomicAdd(count + bin, 1);
    keyRank[i] = make_int2(bin, r);
}

__global__ void __launch_bounds__(kScanThreads) bin_scan_local_kernel(const BinState *state, const int *count, int *start,
                                                                      int *blockSum, int numBins) {
    if (state->mode != kPatchModeGrouped) return;
    __shared__ int warpSum[kScanThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = blockIdx.x * kScanTile + tid * kScanItems;
    int v[kScanItems], sum = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
        v[q] = (i0 + q < numBins) ? count[i0 + q] : 0;
        sum += v[q];
    }
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warpSum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int ws = warpSum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, ws, d);
            if (lane >= d) ws += o;
        }
        warpSum[lane] = ws;                                     // inclusive over warps
    }
    __syncthreads();
    int run = incl - sum + (warp > 0 ? warpSum[warp - 1] : 0);  // exclusive prefix of this thread inside the block
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
        if (i0 + q < numBins) start[i0 + q] = run;
        run += v[q];
    }
    if (tid == kScanThreads - 1) blockSum[blockIdx.x] = warpSum[kScanThreads / 32 - 1];
}

// exclusive scan of the block totals in place (single block; loops with a carry when there are more than 1024)
__global__ void __launch_bounds__(kScanThreads) bin_scan_top_kernel(const BinState *state, int *blockSum, int numBlocks) {
    if (state->mode != kPatchModeGrouped) return;
    __shared__ int warpSum[kScanThreads / 32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < numBlocks; b0 += kScanThreads) {
        const int v = (b0 + tid < numBlocks) ? blockSum[b0 + tid] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warpSum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int ws = warpSum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, ws, d);
                if (lane >= d) ws += o;
            }
            warpSum[lane] = ws;
        }
        __syncthreads();
        const int excl = carry + incl - v + (warp > 0 ? warpSum[warp - 1] : 0);
        if (b0 + tid < numBlocks) blockSum[b0 + tid] = excl;
        __syncthreads();
        if (tid == 0) carry += warpSum[kScanThreads / 32 - 1];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) bin_scatter_kernel(const BinState *state, const int2 *keyRank, const int *start,
                                                          const int *blockOff, int n, int *perm) {
    if (state->mode != kPatchModeGrouped) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 kr = keyRank[i];
    perm[__ldg(start + kr.x) + __ldg(blockOff + kr.x / kScanTile) + kr.y] = i;
}
#endif

}  // namespace b200osd
