/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * osd_oracle.c : plain-C, single-threaded restatement of the OpenSubdiv 3.6.0 CPU algorithm for the
 * two Osd evaluator hot paths (stencil application and limit patch evaluation).  It is the checker
 * the B200 kernels are compared against on the GPU box, where /root/reference does not exist.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py compares every function here against
 * the unmodified reference compiled in place (oracle/_ref/libosdref.so: Osd::CpuEvaluator,
 * OsdEvaluatePatchBasis) and tests/test_oracle_golden.py against the committed golden vectors
 * (tests/golden/, incl. the reference's own regression/hbr_regression/baseline/catmark_cube_level3.obj).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  opensubdiv_b200/ never does.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no FMA contraction: the reference x86-64 build
 * has none either, /root/reference/CMakeLists.txt:173-262, which is what fixes the rounding).
 *
 * Reference files restated (paths relative to /root/reference/opensubdiv):
 *   osd/cpuKernel.cpp:71-240      CpuEvalStencils (value, +du/dv, +duu/duv/dvv)
 *   osd/cpuEvaluator.cpp:37-125   EvalStencils argument checks
 *   osd/cpuEvaluator.cpp:157-381  EvalPatches (value, +D1, +D1+D2) and BufferAdapter :127-154
 *   osd/patchBasisTypes.h:241-426 patch descriptor ids, PatchParam bit fields, (s,t) normalisation
 *   osd/patchBasis.h:53-1610      the six bases, boundary folding, derivative scaling
 *   far/patchMap.h:127-217, far/patchMap.cpp:96-188   PatchMap construction and FindPatch (sample location)
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

/* Arithmetic type.  Default float = the reference's arithmetic (this is THE oracle).  Compiled a second time with
 * -DORACLE_F64 every value (primvars, weights, results) is a double: the same algorithm evaluated in double precision on
 * the same fp32 inputs, i.e. the "truth" against which the rounding error of the reference and of the B200 kernels
 * can both be measured (tests/test_accuracy_vs_f64.py).  PatchCoord (s,t) stay fp32 inputs in both builds. */
#ifdef ORACLE_F64
typedef double real;
#define RABS fabs
#else
typedef float real;
#define RABS fabsf
#endif

/* Tolerance scale mode (tests only): when set, every accumulation uses |value| * |weight|, so the evaluators
 * return S = sum_j |w_j| |x_j| -- the magnitude against which a 1e-6 relative bound is meaningful when signed
 * derivative weights cancel ("Parity at 1e-6 relative"). */
static int g_abs_mode = 0;      /* 0 off, 1 plain S = sum |w||x|, 2 conditioned (patches): S = sum (|w|+W)|x| */
void oracle_set_abs_mode(int on) { g_abs_mode = on; }
/* For patches the weights themselves are computed with cancellation (boundary folding subtracts phantom weights,
 * the B-spline polynomials cancel near the knots), so each weight carries an absolute error ~eps*W with
 * W = max_j |unfolded w_j| of its set.  In abs mode EvalPatches therefore returns
 *     S = sum_j (|w_j| + W) |x_j|
 * which bounds the rounding error of ANY correct fp32 evaluation order (reference's included) by ~n*eps*S. */
static real g_wmax[6];
static void note_wmax(real *const w[6], int nsets, int npts)
{
    int k, j;
    for (k = 0; k < nsets; ++k) {
        real m = 0.0f;
        for (j = 0; j < npts; ++j) if (RABS(w[k][j]) > m) m = RABS(w[k][j]);
        g_wmax[k] = m;
    }
}

#define ORACLE_MAX_LEN 64          /* max primvar length handled by the stack temporaries */

typedef struct { int offset, length, stride; } oracle_desc;                          /* osd/bufferDescriptor.h:61-104 */
typedef struct { int arrayIndex, patchIndex, vertIndex; float s, t; } oracle_coord; /* osd/types.h:42-64 (20 B) */
typedef struct { int regDesc, desc, numPatches, indexBase, stride, primitiveIdBase; } oracle_array; /* osd/types.h:66-122 (24 B) */
typedef struct { unsigned int field0, field1; float sharpness; } oracle_param;      /* osd/types.h:127-130 (12 B) */

enum { PT_QUADS = 3, PT_TRIANGLES = 4, PT_LOOP = 5, PT_REGULAR = 6, PT_GREGORY_BASIS = 9, PT_GREGORY_TRIANGLE = 10 };

/* ------------------------------------------------------------------------------------------------
 * Stencils.  osd/cpuKernel.cpp:71-122 (nw=1), :124-170 (nw=3), :172-240 (nw=6).
 *
 * For every row i in [start,end):   out_k[i][0..L) = sum_{j<sizes[i]} w_k[offsets[i]+j] * src[indices[offsets[i]+j]][0..L)
 * accumulated sequentially in j with a separate multiply and add, accumulator starting at 0.
 *
 * Row addressing: the reference is inconsistent for start>0 (generic CPU path writes row start+i to
 * dst element i, cpuKernel.cpp:110-120; the SIMD path cpuKernel.h:105-139, TBB and CUDA write dst
 * element start+i).  Everything agrees for start==0.  This oracle uses the absolute-index convention
 * (row i -> dst element i), the one the CUDA backend (osd/cudaKernel.cu:85-98) and row-range sharding use.
 *
 * Returns 1 on success, 0 where CpuEvaluator::EvalStencils returns false (length mismatch,
 * cpuEvaluator.cpp:47,70-72,105-110).  end<=start is a successful no-op (cpuEvaluator.cpp:46).
 * -----------------------------------------------------------------------------------------------*/
int oracle_eval_stencils(int nw,
                         const real *src, const oracle_desc *srcDesc,
                         real *const *dsts, const oracle_desc *dstDescs,
                         const int *sizes, const int *offsets, const int *indices,
                         const real *const *weights, int start, int end)
{
    int L, i, j, k, w;
    if (end <= start) return 1;
    L = srcDesc->length;
    for (w = 0; w < nw; ++w)
        if (dstDescs[w].length != L) return 0;
    if (L > ORACLE_MAX_LEN || nw > 6) return 0;

    src += srcDesc->offset;
    for (i = start; i < end; ++i) {
        real acc[6][ORACLE_MAX_LEN];
        int off = offsets[i];
        int n = sizes[i];
        memset(acc, 0, sizeof(acc));
        for (j = 0; j < n; ++j) {
            const real *v = src + (ptrdiff_t)indices[off + j] * srcDesc->stride;
            for (w = 0; w < nw; ++w) {
                real wt = weights[w][off + j];
                if (g_abs_mode) { for (k = 0; k < L; ++k) acc[w][k] += RABS(v[k]) * RABS(wt); }
                else            { for (k = 0; k < L; ++k) acc[w][k] += v[k] * wt; }   /* addWithWeight, cpuKernel.cpp:52-61 */
            }
        }
        for (w = 0; w < nw; ++w) {
            if (!dsts[w]) continue;
            memcpy(dsts[w] + dstDescs[w].offset + (ptrdiff_t)i * dstDescs[w].stride, acc[w], (size_t)L * sizeof(real));
        }
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * PatchParam bit fields.  far/patchParam.h:81-96,234-262 == osd/patchBasisTypes.h:310-372.
 *   field0 = faceId:28 | transition:4
 *   field1 = depth:4 | nonquad:1 | regular:1 | unused:1 | boundary:5 | v:10 | u:10   (LSB first)
 * -----------------------------------------------------------------------------------------------*/
static unsigned pp_face(unsigned f0) { return f0 & 0x0fffffffu; }
static int pp_depth(unsigned f1)    { return (int)(f1 & 0xfu); }
static int pp_nonquad(unsigned f1)  { return (int)((f1 >> 4) & 1u); }
static int pp_regular(unsigned f1)  { return (int)((f1 >> 5) & 1u); }
static int pp_boundary(unsigned f1) { return (int)((f1 >> 7) & 0x1fu); }
static int pp_v(unsigned f1)        { return (int)((f1 >> 12) & 0x3ffu); }
static int pp_u(unsigned f1)        { return (int)((f1 >> 22) & 0x3ffu); }

/* ---------------------------------------------------------------------------- 1-D cubic curves --
 * Uniform cubic B-spline (osd/patchBasis.h:99-135) and cubic Bernstein (:247-283) bases with first
 * and second derivatives.  d1/d2 may be NULL.  */
static void bspline3(real t, real *b, real *d1, real *d2)
{
    #ifdef ORACLE_F64
    const real sixth = 1.0 / 6.0;
#else
    const real sixth = (real)(1.0f / 6.0f);
#endif
    real t2 = t * t, t3 = t * t2;
    b[0] = sixth * (1.0f - 3.0f * (t - t2) - t3);
    b[1] = sixth * (4.0f - 6.0f * t2 + 3.0f * t3);
    b[2] = sixth * (1.0f + 3.0f * (t + t2 - t3));
    b[3] = sixth * t3;
    if (d1) {
        d1[0] = -0.5f * t2 + t - 0.5f;
        d1[1] = 1.5f * t2 - 2.0f * t;
        d1[2] = -1.5f * t2 + t + 0.5f;
        d1[3] = 0.5f * t2;
    }
    if (d2) {
        d2[0] = -t + 1.0f;
        d2[1] = 3.0f * t - 2.0f;
        d2[2] = -3.0f * t + 1.0f;
        d2[3] = t;
    }
}

static void bezier3(real t, real *b, real *d1, real *d2)
{
    real t2 = t * t, c = 1.0f - t, c2 = c * c;
    b[0] = c2 * c;
    b[1] = c2 * t * 3.0f;
    b[2] = t2 * c * 3.0f;
    b[3] = t2 * t;
    if (d1) {
        d1[0] = -3.0f * c2;
        d1[1] = 9.0f * t2 - 12.0f * t + 3.0f;
        d1[2] = -9.0f * t2 + 6.0f * t;
        d1[3] = 3.0f * t2;
    }
    if (d2) {
        d2[0] = 6.0f * c;
        d2[1] = 18.0f * t - 12.0f;
        d2[2] = -18.0f * t + 6.0f;
        d2[3] = 6.0f * t;
    }
}

/* w[4*row + col] = cs[col] * ct[row]  (osd/patchBasis.h:218-243) */
static void tensor4(const real *cs, const real *ct, real *w)
{
    int r, c;
    for (r = 0; r < 4; ++r)
        for (c = 0; c < 4; ++c) w[4 * r + c] = cs[c] * ct[r];
}

/* ------------------------------------------------------------------------------------ linear --
 * osd/patchBasis.h:53-97 (bilinear quad) and :493-525 (linear triangle). */
static int basis_quads(real s, real t, real *w[6], int order)
{
    real sc = 1.0f - s, tc = 1.0f - t;
    w[0][0] = sc * tc; w[0][1] = s * tc; w[0][2] = s * t; w[0][3] = sc * t;
    if (order >= 1) {
        w[1][0] = -tc; w[1][1] = tc; w[1][2] = t; w[1][3] = -t;
        w[2][0] = -sc; w[2][1] = -s; w[2][2] = s; w[2][3] = sc;
    }
    if (order >= 2) {
        int i;
        for (i = 0; i < 4; ++i) { w[3][i] = 0.0f; w[5][i] = 0.0f; }
        w[4][0] = 1.0f; w[4][1] = -1.0f; w[4][2] = 1.0f; w[4][3] = -1.0f;
    }
    return 4;
}

static int basis_tris(real s, real t, real *w[6], int order)
{
    w[0][0] = 1.0f - s - t; w[0][1] = s; w[0][2] = t;
    if (order >= 1) {
        w[1][0] = -1.0f; w[1][1] = 1.0f; w[1][2] = 0.0f;
        w[2][0] = -1.0f; w[2][1] = 0.0f; w[2][2] = 1.0f;
    }
    if (order >= 2) {
        int i, k;
        for (k = 3; k < 6; ++k)
            for (i = 0; i < 3; ++i) w[k][i] = 0.0f;
    }
    return 3;
}

/* --------------------------------------------------------------------------- regular B-spline --
 * Tensor product (osd/patchBasis.h:204-245) followed by boundary folding (:138-200): for each boundary
 * edge the phantom row/column of weights w0 is folded into its two neighbours, w1 += 2*w0, w2 -= w0,
 * w0 = 0, edges taken in bit order 1 (row 0), 2 (col 3), 4 (row 3), 8 (col 0).  */
static void fold_line(real *w, int i0, int i1, int i2, int step, int count)
{
    int k;
    for (k = 0; k < count; ++k, i0 += step, i1 += step, i2 += step) {
        w[i2] -= w[i0];
        w[i1] += w[i0] * 2.0f;
        w[i0] = 0.0f;
    }
}

static void bspline_fold_boundary(int mask, real *w)
{
    if (mask & 1) fold_line(w, 0, 4, 8, 1, 4);        /* t = 0 edge: row 0 -> rows 1,2 */
    if (mask & 2) fold_line(w, 3, 2, 1, 4, 4);        /* s = 1 edge: col 3 -> cols 2,1 */
    if (mask & 4) fold_line(w, 12, 8, 4, 1, 4);       /* t = 1 edge: row 3 -> rows 2,1 */
    if (mask & 8) fold_line(w, 0, 1, 2, 4, 4);        /* s = 0 edge: col 0 -> cols 1,2 */
}

static int basis_regular(real s, real t, int boundary, real *w[6], int order)
{
    real bs[4], bt[4], ds[4], dt[4], dss[4], dtt[4];
    int k;
    bspline3(s, bs, order >= 1 ? ds : NULL, order >= 2 ? dss : NULL);
    bspline3(t, bt, order >= 1 ? dt : NULL, order >= 2 ? dtt : NULL);
    tensor4(bs, bt, w[0]);
    if (order >= 1) { tensor4(ds, bt, w[1]); tensor4(bs, dt, w[2]); }
    if (order >= 2) { tensor4(dss, bt, w[3]); tensor4(ds, dt, w[4]); tensor4(bs, dtt, w[5]); }
    note_wmax(w, order == 0 ? 1 : (order == 1 ? 3 : 6), 16);
    if (boundary) {
        int nsets = order == 0 ? 1 : (order == 1 ? 3 : 6);
        for (k = 0; k < nsets; ++k) bspline_fold_boundary(boundary, w[k]);
    }
    return 16;
}

/* ------------------------------------------------------------------------------ Gregory basis --
 * osd/patchBasis.h:332-490.  20 points, 5 per corner c: P (5c), E+ (5c+1), E- (5c+2), F+ (5c+3), F- (5c+4).
 * Each maps onto one position (col,row) of the bicubic Bezier net; the 8 face points additionally
 * carry the rational blend G: with (a,b) the distances from corner c along its E+ / E- directions,
 * G+ = a/(a+b) and G- = 1 - G+ (so each pair sums to one exactly); when a+b <= 0 the reciprocal is
 * replaced by 1 (:369-372).  Derivatives use the reference's default approximation (:421-440): the
 * Bezier derivative weights times the same G; with oracle_set_gregory_true_derivatives(1) they follow the
 * reference's OPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES build instead (:441-487): quotient rule for G = N/D
 * (N' and D' are the constants below), product rule for B*G.  Pinned against oracle/_ref/libosdref_td.so.  */
static int g_gregory_true = 0;
void oracle_set_gregory_true_derivatives(int on) { g_gregory_true = on; }
static const real GREG_NDS[8] = { 1.0f, 0.0f,  0.0f, -1.0f, -1.0f,  0.0f,  0.0f,  1.0f };
static const real GREG_NDT[8] = { 0.0f, 1.0f,  1.0f,  0.0f,  0.0f, -1.0f, -1.0f,  0.0f };
static const real GREG_DDS[8] = { 1.0f, 1.0f, -1.0f, -1.0f, -1.0f, -1.0f,  1.0f,  1.0f };
static const real GREG_DDT[8] = { 1.0f, 1.0f,  1.0f,  1.0f, -1.0f, -1.0f, -1.0f, -1.0f };
static const signed char GREG_COL[20] = { 0, 1, 0, 1, 1,   3, 3, 2, 2, 2,   3, 2, 3, 2, 2,   0, 0, 1, 1, 1 };
static const signed char GREG_ROW[20] = { 0, 0, 1, 1, 1,   0, 1, 0, 1, 1,   3, 3, 2, 2, 2,   3, 2, 3, 2, 2 };

static int basis_gregory(real s, real t, real *w[6], int order)
{
    real bs[4], bt[4], ds[4], dt[4], dss[4], dtt[4], G[8], R[4];
    real sc = 1.0f - s, tc = 1.0f - t;
    real a[4], b[4];
    int c, i;
    bezier3(s, bs, order >= 1 ? ds : NULL, order >= 2 ? dss : NULL);
    bezier3(t, bt, order >= 1 ? dt : NULL, order >= 2 ? dtt : NULL);

    a[0] = s;  b[0] = t;
    a[1] = t;  b[1] = sc;
    a[2] = sc; b[2] = tc;
    a[3] = tc; b[3] = s;
    for (c = 0; c < 4; ++c) {
        real d = (c == 0) ? (s + t) : (c == 1) ? (sc + t) : (c == 2) ? (sc + tc) : (s + tc);
        real r = (d <= 0.0f) ? 1.0f : (1.0f / d);
        G[2 * c] = a[c] * r;
        G[2 * c + 1] = 1.0f - a[c] * r;
        R[c] = r;
        (void)b;
    }
    for (i = 0; i < 20; ++i) {
        int col = GREG_COL[i], row = GREG_ROW[i], p = i % 5;
        int rational = (p >= 3);
        real g = rational ? G[2 * (i / 5) + (p - 3)] : 1.0f;
        if (rational && g_gregory_true) {
            int k = 2 * (i / 5) + (p - 3);
            real D = R[i / 5];
            w[0][i] = bs[col] * bt[row] * g;
            if (order >= 1) {
                real g_s = (GREG_NDS[k] - GREG_DDS[k] * g) * D;
                real g_t = (GREG_NDT[k] - GREG_DDT[k] * g) * D;
                w[1][i] = (ds[col] * g + bs[col] * g_s) * bt[row];
                w[2][i] = (dt[row] * g + bt[row] * g_t) * bs[col];
                if (order >= 2) {
                    real invD2 = D * D;
                    real g_ss = 2.0f * GREG_DDS[k] * invD2 * (g * GREG_DDS[k] - GREG_NDS[k]);
                    real g_st = invD2 * (2.0f * g * GREG_DDS[k] * GREG_DDT[k] - GREG_NDS[k] * GREG_DDT[k] - GREG_NDT[k] * GREG_DDS[k]);
                    real g_tt = 2.0f * GREG_DDT[k] * invD2 * (g * GREG_DDT[k] - GREG_NDT[k]);
                    w[3][i] = (dss[col] * g + 2.0f * ds[col] * g_s + bs[col] * g_ss) * bt[row];
                    w[4][i] = bt[row] * (bs[col] * g_st + ds[col] * g_t) + dt[row] * (ds[col] * g + bs[col] * g_s);
                    w[5][i] = (dtt[row] * g + 2.0f * dt[row] * g_t + bt[row] * g_tt) * bs[col];
                }
            }
        } else if (rational) {
            w[0][i] = bs[col] * bt[row] * g;
            if (order >= 1) {
                w[1][i] = ds[col] * bt[row] * g;
                w[2][i] = dt[row] * bs[col] * g;
            }
            if (order >= 2) {
                w[3][i] = dss[col] * bt[row] * g;
                w[4][i] = ds[col] * dt[row] * g;
                w[5][i] = bs[col] * dtt[row] * g;
            }
        } else {
            w[0][i] = bs[col] * bt[row];
            if (order >= 1) {
                w[1][i] = ds[col] * bt[row];
                w[2][i] = dt[row] * bs[col];
            }
            if (order >= 2) {
                w[3][i] = dss[col] * bt[row];
                w[4][i] = ds[col] * dt[row];
                w[5][i] = bs[col] * dtt[row];
            }
        }
    }
    return 20;
}

/* --------------------------------------------------------------------- Loop quartic box spline --
 * osd/patchBasis.h:527-917.  The 12 basis functions of the regular Loop patch are bivariate quartics;
 * BOX12[i][m] are their coefficients (times 12) on the monomials
 *   m: 0:1  1:s  2:t  3:s^2  4:st  5:t^2  6:s^3  7:s^2t  8:st^2  9:t^3  10:s^4  11:s^3t  12:s^2t^2  13:st^3  14:t^4
 * All derivative tables are DERIVED here by differentiating that one table (d/ds s^a t^b = a s^(a-1) t^b),
 * then normalised to the integer scale the reference uses (1/12, 1/6, 1, 1/2, 1) so that evaluation
 * in increasing-monomial order reproduces the same real sequence.  */
static const signed char BOX12[12][15] = {
    /*        1   s   t  ss  st  tt sss sst stt ttt  s4 s3t s2t2 st3  t4 */
    /* 0*/ {  1, -2, -4,  0,  6,  6,  2,  0, -6, -4, -1, -2,  0,  2,  1 },
    /* 1*/ {  1,  2, -2,  0, -6,  0, -4,  0,  6,  2,  2,  4,  0, -2, -1 },
    /* 2*/ {  0,  0,  0,  0,  0,  0,  2,  0,  0,  0, -1, -2,  0,  0,  0 },
    /* 3*/ {  1, -4, -2,  6,  6,  0, -4, -6,  0,  2,  1,  2,  0, -2, -1 },
    /* 4*/ {  6,  0,  0,-12,-12,-12,  8, 12, 12,  8, -1, -2,  0, -2, -1 },
    /* 5*/ {  1,  4,  2,  6,  6,  0, -4, -6,-12, -4, -1, -2,  0,  4,  2 },
    /* 6*/ {  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  1,  2,  0,  0,  0 },
    /* 7*/ {  1, -2,  2,  0, -6,  0,  2,  6,  0, -4, -1, -2,  0,  4,  2 },
    /* 8*/ {  1,  2,  4,  0,  6,  6, -4,-12, -6, -4,  2,  4,  0, -2, -1 },
    /* 9*/ {  0,  0,  0,  0,  0,  0,  2,  6,  6,  2, -1, -2,  0, -2, -1 },
    /*10*/ {  0,  0,  0,  0,  0,  0,  0,  0,  0,  2,  0,  0,  0, -2, -1 },
    /*11*/ {  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  2,  1 },
};
static const signed char MONO_A[15] = { 0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 4, 3, 2, 1, 0 };   /* power of s */
static const signed char MONO_B[15] = { 0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4 };   /* power of t */

static int mono_index(int a, int b)
{
    int m;
    for (m = 0; m < 15; ++m)
        if (MONO_A[m] == a && MONO_B[m] == b) return m;
    return -1;
}

/* tables[k][i][m], k: 0 value, 1 d/ds, 2 d/dt, 3 dss, 4 dst, 5 dtt; scales[k] the common factor */
static int   g_box_ready = 0;
static int   g_box_tab[6][12][15];
static real g_box_scale[6];

static void box_init(void)
{
    /* divisor that turns (coefficient*12) after differentiation into the reference's integer table */
    static const int das[6] = { 0, 1, 0, 2, 1, 0 }, dbs[6] = { 0, 0, 1, 0, 1, 2 };
    static const int divisor[6] = { 1, 2, 2, 12, 6, 12 };
    int k, i, m;
    memset(g_box_tab, 0, sizeof(g_box_tab));
    for (k = 0; k < 6; ++k) {
        for (i = 0; i < 12; ++i) {
            for (m = 0; m < 15; ++m) {
                int a = MONO_A[m], b = MONO_B[m], c = BOX12[i][m], q, n;
                if (c == 0 || a < das[k] || b < dbs[k]) continue;
                for (q = 0; q < das[k]; ++q) c *= (a - q);
                for (q = 0; q < dbs[k]; ++q) c *= (b - q);
                n = mono_index(a - das[k], b - dbs[k]);
                g_box_tab[k][i][n] += c / divisor[k];       /* exact: every product is a multiple */
            }
        }
    }
#ifdef ORACLE_F64
    g_box_scale[0] = 1.0 / 12.0;
    g_box_scale[1] = g_box_scale[2] = 1.0 / 6.0;
#else
    g_box_scale[0] = (real)(1.0f / 12.0f);
    g_box_scale[1] = g_box_scale[2] = (real)(1.0f / 6.0f);
#endif
    g_box_scale[3] = g_box_scale[5] = 1.0f;
    g_box_scale[4] = (real)(1.0f / 2.0f);
    g_box_ready = 1;
}

/* Boundary folding for the box-spline patch (osd/patchBasis.h:663-886).  The 5-bit mask encodes
 * boundary edges (lower 3 bits) and, through the upper 2 bits, whether those bits instead denote
 * boundary vertices (upper==1) or whether each boundary edge also makes the opposite vertex a
 * boundary vertex (upper==2, vertex bits = edge bits rotated right by one).
 * Every phantom point is a reflection P = B + (B' - I) of patch points, so its weight w is
 * distributed as +w to B, +w to B', -w to I and then cleared.
 * Patch point numbering (12 points):        Edge e has phantom points ph[e][0..2]; corner/neighbour
 *            0   1   2                      roles follow from the triangle symmetry (rotation by 120
 *          3   4   5   6                    degrees maps the edge-0 roles onto edge 1 and edge 2).
 *            7   8   9
 *             10  11
 */
static void refl(real *w, int phantom, int plus0, int plus1, int minus)
{
    real v = w[phantom];
    w[plus0] += v;
    w[plus1] += v;
    w[minus] -= v;
}

static void box_fold_boundary(int mask, real *w)
{
    /* per edge e: phantom triple, the three patch corners' ring points used by the reflections.
     * Roles for edge 0:  phantoms (0,1,2); B1=4, B2=5; I1=8; B0=3,I0=7 (left neighbour), B3=6,I2=9 (right).  */
    static const signed char PH[3][3] = { { 0, 1, 2 }, { 6, 9, 11 }, { 10, 7, 3 } };
    static const signed char B1[3] = { 4, 5, 8 }, B2[3] = { 5, 8, 4 }, I1[3] = { 8, 4, 5 };
    static const signed char B0[3] = { 3, 2, 11 }, I0[3] = { 7, 1, 9 };
    static const signed char B3[3] = { 6, 10, 0 }, I2[3] = { 9, 7, 1 };
    /* per vertex v: two phantoms and their reflection stencils */
    static const signed char VP[3][2] = { { 3, 0 }, { 2, 6 }, { 11, 10 } };
    static const signed char VB1[3] = { 4, 5, 8 };
    static const signed char VB0[3] = { 7, 1, 9 }, VI0[3] = { 8, 4, 5 };
    static const signed char VB2[3] = { 1, 9, 7 }, VI1[3] = { 5, 8, 4 };
    int upper, ebits, vbits = 0, e, v;
    if (mask == 0) return;
    upper = (mask >> 3) & 3;
    ebits = mask & 7;
    if (upper == 1) { vbits = ebits; ebits = 0; }
    else if (upper == 2) { vbits = ((ebits & 1) << 2) | (ebits >> 1); }

    for (e = 0; e < 3; ++e) {
        int prev = (e + 2) % 3, next = (e + 1) % 3;
        if (!(ebits & (1 << e))) continue;
        /* first phantom: if the previous edge is a boundary too, reflect about B1 only */
        if (ebits & (1 << prev)) refl(w, PH[e][0], B1[e], B1[e], I1[e]);
        else                     refl(w, PH[e][0], B1[e], B0[e], I0[e]);
        refl(w, PH[e][1], B1[e], B2[e], I1[e]);
        if (ebits & (1 << next)) refl(w, PH[e][2], B2[e], B2[e], I1[e]);
        else                     refl(w, PH[e][2], B2[e], B3[e], I2[e]);
        w[PH[e][0]] = w[PH[e][1]] = w[PH[e][2]] = 0.0f;
    }
    for (v = 0; v < 3; ++v) {
        if (!(vbits & (1 << v))) continue;
        refl(w, VP[v][0], VB1[v], VB0[v], VI0[v]);
        refl(w, VP[v][1], VB1[v], VB2[v], VI1[v]);
        w[VP[v][0]] = w[VP[v][1]] = 0.0f;
    }
}

static int basis_loop(real s, real t, int boundary, real *w[6], int order)
{
    real M[15];
    int nsets = order == 0 ? 1 : (order == 1 ? 3 : 6), k, i, m;
    if (!g_box_ready) box_init();
    /* monomials built by repeated multiplication exactly as :533-556 */
    M[0] = 1.0f; M[1] = s; M[2] = t;
    M[3] = s * s; M[4] = s * t; M[5] = t * t;
    M[6] = M[3] * s; M[7] = M[4] * s; M[8] = M[4] * t; M[9] = M[5] * t;
    M[10] = M[6] * s; M[11] = M[7] * s; M[12] = M[3] * M[5]; M[13] = M[8] * t; M[14] = M[9] * t;
    for (k = 0; k < nsets; ++k) {
        for (i = 0; i < 12; ++i) {
            real acc = 0.0f;
            for (m = 0; m < 15; ++m) {
                int c = g_box_tab[k][i][m];
                if (c) acc += (real)c * M[m];
            }
            w[k][i] = g_box_scale[k] * acc;
        }
    }
    note_wmax(w, nsets, 12);
    for (k = 0; k < nsets; ++k)
        if (boundary) box_fold_boundary(boundary, w[k]);
    return 12;
}

/* --------------------------------------------------------------------------- Gregory triangle --
 * osd/patchBasis.h:921-1175.  Quartic Bernstein basis over the triangle, B_ijk = 4!/(i!j!k!) u^i v^j w^k
 * with w = 1-u-v, in the reference's 15-point order (rows of constant v power); derivatives w.r.t.
 * s = u and t = v follow from dw/ds = dw/dt = -1, written with the lower-degree Bernstein functions:
 *   dB4_ijk/ds = 4 (B3_{i-1,j,k} - B3_{i,j,k-1}),   d2/ds2 = 12 (B2_{i-2,j,k} - 2 B2_{i-1,j,k-1} + B2_{i,j,k-2}), ...
 * The 18 Gregory-triangle points reuse 12 boundary Bernstein weights and split the 3 interior ones
 * with rational blends G (:1044-1110), default 1/0 when a denominator vanishes (:1132-1145).  */
static real bern(int n, int i, int j, int k, real u, real v, real w)
{
    static const real fact[5] = { 1.0f, 1.0f, 2.0f, 6.0f, 24.0f };
    real r;
    int q;
    if (i < 0 || j < 0 || k < 0) return 0.0f;
    r = fact[n] / (fact[i] * fact[j] * fact[k]);
    for (q = 0; q < i; ++q) r *= u;
    for (q = 0; q < j; ++q) r *= v;
    for (q = 0; q < k; ++q) r *= w;
    return r;
}

static void bezier_tri4(real s, real t, int ds, int dt, real *B)
{
    /* point order: index -> (i = power of u, j = power of v), k = 4-i-j */
    static const signed char PI[15] = { 0, 1, 2, 3, 4, 0, 1, 2, 3, 0, 1, 2, 0, 1, 0 };
    static const signed char PJ[15] = { 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 4 };
    real u = s, v = t, w = 1.0f - u - v;
    int n;
    for (n = 0; n < 15; ++n) {
        int i = PI[n], j = PJ[n], k = 4 - i - j;
        real r;
        if (ds + dt == 0) {
            r = bern(4, i, j, k, u, v, w);
        } else if (ds + dt == 1) {
            real lower = ds ? bern(3, i - 1, j, k, u, v, w) : bern(3, i, j - 1, k, u, v, w);
            r = 4.0f * (lower - bern(3, i, j, k - 1, u, v, w));
        } else if (ds == 2) {
            r = 12.0f * (bern(2, i - 2, j, k, u, v, w) - 2.0f * bern(2, i - 1, j, k - 1, u, v, w) + bern(2, i, j, k - 2, u, v, w));
        } else if (dt == 2) {
            r = 12.0f * (bern(2, i, j - 2, k, u, v, w) - 2.0f * bern(2, i, j - 1, k - 1, u, v, w) + bern(2, i, j, k - 2, u, v, w));
        } else {
            r = 12.0f * (bern(2, i - 1, j - 1, k, u, v, w) - bern(2, i - 1, j, k - 1, u, v, w)
                         - bern(2, i, j - 1, k - 1, u, v, w) + bern(2, i, j, k - 2, u, v, w));
        }
        B[n] = r;
    }
}

static void gregory_tri_from_bezier(const real *B, const real *G, real *w)
{
    /* 18 points = 3 corners x {P, E+, E-, F+, F-} + 3 edge mid points (osd/patchBasis.h:1044-1110) */
    static const signed char SRC[18] = { 0, 1, 5, 6, 6,   4, 8, 3, 7, 7,   14, 12, 13, 10, 10,   2, 11, 9 };
    static const signed char GI[18]  = { -1, -1, -1, 0, 1,  -1, -1, -1, 2, 3,  -1, -1, -1, 4, 5,  -1, -1, -1 };
    int i;
    for (i = 0; i < 18; ++i) w[i] = (GI[i] < 0) ? B[SRC[i]] : B[SRC[i]] * G[GI[i]];
}

static int basis_gregory_tri(real s, real t, real *w[6], int order)
{
    static const int DS[6] = { 0, 1, 0, 2, 1, 0 }, DT[6] = { 0, 0, 1, 0, 1, 2 };
    real G[6] = { 1.0f, 0.0f, 1.0f, 0.0f, 1.0f, 0.0f };
    real u = s, v = t, ww = 1.0f - u - v, B[15];
    int nsets = order == 0 ? 1 : (order == 1 ? 3 : 6), k;
    if ((u + v) > 0.0f)  { G[0] = u / (u + v);   G[1] = v / (u + v); }
    if ((v + ww) > 0.0f) { G[2] = v / (v + ww);  G[3] = ww / (v + ww); }
    if ((ww + u) > 0.0f) { G[4] = ww / (ww + u); G[5] = u / (ww + u); }
    for (k = 0; k < nsets; ++k) {
        bezier_tri4(s, t, DS[k], DT[k], B);
        gregory_tri_from_bezier(B, G, w[k]);
    }
    return 18;
}

/* ------------------------------------------------------------------------------------------------
 * OsdEvaluatePatchBasis (osd/patchBasis.h:1555-1610) = normalise (s,t) into the sub-patch
 * (osd/patchBasisTypes.h:374-426), evaluate (patchBasis.h:1444-1553), scale derivatives by
 * d1 = sign * 2^depth and d2 = sign * d1 * d1 (sign = -1 for a rotated triangle).
 * order: 0 -> wP only, 1 -> +wDs,wDt, 2 -> +wDss,wDst,wDtt.  Returns the number of points.
 * -----------------------------------------------------------------------------------------------*/
int oracle_patch_basis(int patchType, unsigned field0, unsigned field1, real s, real t,
                       real *wP, real *wDs, real *wDt, real *wDss, real *wDst, real *wDtt)
{
    real *w[6];
    int order = (wDs && wDt) ? ((wDss && wDst && wDtt) ? 2 : 1) : 0;
    int depth = pp_depth(field1), n = 0, i;
    int isTri = (patchType == PT_LOOP || patchType == PT_GREGORY_TRIANGLE || patchType == PT_TRIANGLES);
    real sign = 1.0f;
    real fracInv = (real)(1 << (depth - pp_nonquad(field1)));
    (void)field0;
    w[0] = wP; w[1] = wDs; w[2] = wDt; w[3] = wDss; w[4] = wDst; w[5] = wDtt;

    if (isTri && (pp_u(field1) + pp_v(field1)) >= (1 << depth)) {
        int df = 1 << depth;
        s = (real)(df - pp_u(field1)) - (s * fracInv);
        t = (real)(df - pp_v(field1)) - (t * fracInv);
        sign = -1.0f;
    } else {
        s = s * fracInv - (real)pp_u(field1);
        t = t * fracInv - (real)pp_v(field1);
    }

    switch (patchType) {
        case PT_REGULAR:          n = basis_regular(s, t, pp_boundary(field1), w, order); break;
        case PT_LOOP:             n = basis_loop(s, t, pp_boundary(field1), w, order); break;
        case PT_GREGORY_BASIS:    n = basis_gregory(s, t, w, order); break;
        case PT_GREGORY_TRIANGLE: n = basis_gregory_tri(s, t, w, order); break;
        case PT_QUADS:            n = basis_quads(s, t, w, order); break;
        case PT_TRIANGLES:        n = basis_tris(s, t, w, order); break;
        default: return 0;
    }
    if (patchType != PT_REGULAR && patchType != PT_LOOP) note_wmax(w, order == 0 ? 1 : (order == 1 ? 3 : 6), n);
    {
        real a1 = (real)(1 << depth), a2 = a1 * a1;
        g_wmax[1] *= a1; g_wmax[2] *= a1; g_wmax[3] *= a2; g_wmax[4] *= a2; g_wmax[5] *= a2;
    }
    if (order >= 1) {
        real d1 = sign * (real)(1 << depth);
        for (i = 0; i < n; ++i) { wDs[i] *= d1; wDt[i] *= d1; }
        if (order >= 2) {
            real d2 = sign * d1 * d1;
            for (i = 0; i < n; ++i) { wDss[i] *= d2; wDst[i] *= d2; wDtt[i] *= d2; }
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * EvalPatches.  osd/cpuEvaluator.cpp:157-210 (nw=1), :213-282 (nw=3), :285-381 (nw=6).
 * For coord i: array = arrays[coord.arrayIndex]; param = params[coord.patchIndex];
 * type = regular(param) ? array.regDesc : array.desc; weights = basis(type,param,s,t);
 * cvs = indices + array.indexBase + array.stride*(patchIndex - array.primitiveIdBase);
 * out_k[i] = sum_j w_k[j] * src[cvs[j]], sequential in j, separate multiply and add.
 * The same routine serves EvalPatchesVarying / EvalPatchesFaceVarying: callers pass the varying /
 * face-varying (arrays, indices, params) triple (osd/cpuEvaluator.h:823-1226).
 * Returns 0 where the reference returns false: src NULL, value-only form with dst NULL, or a
 * non-NULL output whose length differs from srcDesc.length.  NULL derivative outputs are skipped
 * (the CUDA backend's behaviour, osd/cudaKernel.cu:300-327; the CPU reference would dereference them).
 * -----------------------------------------------------------------------------------------------*/
int oracle_eval_patches(int nw,
                        const real *src, const oracle_desc *srcDesc,
                        real *const *dsts, const oracle_desc *dstDescs,
                        int numPatchCoords, const oracle_coord *coords,
                        const oracle_array *arrays, const int *indices, const oracle_param *params)
{
    real wbuf[6][20];
    int L, i, j, k, q;
    if (!src) return 0;
    if (nw == 1 && !dsts[0]) return 0;
    L = srcDesc->length;
    for (q = 0; q < nw; ++q)
        if (dsts[q] && dstDescs[q].length != L) return 0;
    if (L > ORACLE_MAX_LEN) return 0;
    src += srcDesc->offset;

    for (i = 0; i < numPatchCoords; ++i) {
        const oracle_coord *c = &coords[i];
        const oracle_array *a = &arrays[c->arrayIndex];
        const oracle_param *p = &params[c->patchIndex];
        int type = pp_regular(p->field1) ? a->regDesc : a->desc;
        const int *cvs = indices + a->indexBase + a->stride * (c->patchIndex - a->primitiveIdBase);
        int n = oracle_patch_basis(type, p->field0, p->field1, c->s, c->t,
                                   wbuf[0], nw >= 3 ? wbuf[1] : NULL, nw >= 3 ? wbuf[2] : NULL,
                                   nw >= 6 ? wbuf[3] : NULL, nw >= 6 ? wbuf[4] : NULL, nw >= 6 ? wbuf[5] : NULL);
        for (q = 0; q < nw; ++q) {
            real acc[ORACLE_MAX_LEN];
            if (!dsts[q]) continue;
            for (k = 0; k < L; ++k) acc[k] = 0.0f;
            for (j = 0; j < n; ++j) {
                const real *v = src + (ptrdiff_t)cvs[j] * srcDesc->stride;
                if (g_abs_mode) { for (k = 0; k < L; ++k) acc[k] += RABS(v[k]) * (RABS(wbuf[q][j]) + (g_abs_mode == 2 ? g_wmax[q] : 0.0f)); }
                else            { for (k = 0; k < L; ++k) acc[k] += v[k] * wbuf[q][j]; }
            }
            memcpy(dsts[q] + dstDescs[q].offset + (ptrdiff_t)i * dstDescs[q].stride, acc, (size_t)L * sizeof(real));
        }
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * Patch map: (ptexFace, s, t) -> PatchCoord.  far/patchMap.cpp:96-188 (construction), far/patchMap.h:127-217
 * (quadrant selection and the descent), osd/types.h:53-54 (PatchCoord from a handle).
 *
 * A quadtree per ptex face.  Every patch is filed under the sequence of quadrants that leads from the face to the
 * sub-domain its PatchParam (depth, u, v) describes; a query halves the domain level by level, in DOUBLE exactly as
 * the reference does (u >= median ? u -= median ...), until it reaches a leaf.  This restatement keeps the
 * reference's formulation on purpose: the product descends on integer bits instead, so agreement between the two is
 * a real check.  Quad domains use the PatchParam's (u,v) bits for the path; triangular domains locate an interior
 * point of the sub-triangle and track the 180-degree rotation of centre triangles.
 * Child word: 0 = unset, (index << 2) | 1 = inner node, (index << 2) | 3 = leaf (patch index).
 * Build with stdlib only; single-threaded; not reentrant on one map.
 * -----------------------------------------------------------------------------------------------*/
#include <stdlib.h>

typedef struct { unsigned child[4]; } oracle_qnode;
typedef struct {
    oracle_qnode *nodes;
    int numNodes, capNodes;
    int *arrayOf, *vertOf;            /* per patch: handle.arrayIndex, handle.vertIndex (patchIndex = position) */
    int numPatches, minFace, maxFace, maxDepth, triangular;
} oracle_patch_map;

static int pm_type_points(int type)   /* far/patchDescriptor.h: control points per patch type */
{
    switch (type) {
        case 1: return 1;  case 2: return 2;  case 3: return 4;  case 4: return 3;  case 5: return 12;
        case 6: return 16; case 7: return 4;  case 8: return 4;  case 9: return 20; case 10: return 18;
        default: return 0;
    }
}

/* far/patchMap.h:146-175: quadrant of (u,v) in a triangle whose half-size is `median`; moves (u,v) into it */
static int pm_tri_quadrant(double median, double *u, double *v, int *rotated)
{
    if (!*rotated) {
        if (*u >= median) { *u -= median; return 1; }
        if (*v >= median) { *v -= median; return 2; }
        if ((*u + *v) >= median) { *rotated = 1; return 3; }
        return 0;
    }
    if (*u < median) { *v -= median; return 1; }
    if (*v < median) { *u -= median; return 2; }
    *u -= median;
    *v -= median;
    if ((*u + *v) < median) { *rotated = 0; return 3; }
    return 0;
}

static int pm_new_node(oracle_patch_map *m)
{
    if (m->numNodes == m->capNodes) {
        int cap = m->capNodes ? 2 * m->capNodes : 64;
        oracle_qnode *n = (oracle_qnode *)realloc(m->nodes, (size_t)cap * sizeof(oracle_qnode));
        if (!n) return -1;
        m->nodes = n;
        m->capNodes = cap;
    }
    memset(&m->nodes[m->numNodes], 0, sizeof(oracle_qnode));
    return m->numNodes++;
}

void oracle_patch_map_free(void *h)
{
    oracle_patch_map *m = (oracle_patch_map *)h;
    if (!m) return;
    free(m->nodes); free(m->arrayOf); free(m->vertOf); free(m);
}

void *oracle_patch_map_create(int numArrays, const oracle_array *arrays, int numPatches, const oracle_param *params,
                              int patchesAreTriangular)
{
    oracle_patch_map *m = (oracle_patch_map *)calloc(1, sizeof(oracle_patch_map));
    int a, j, h = 0, p, faces;
    if (!m) return NULL;
    m->triangular = patchesAreTriangular ? 1 : 0;
    m->minFace = 0; m->maxFace = -1;
    m->numPatches = numPatches;
    if (numPatches <= 0) return m;
    m->arrayOf = (int *)malloc((size_t)numPatches * sizeof(int));
    m->vertOf = (int *)malloc((size_t)numPatches * sizeof(int));
    if (!m->arrayOf || !m->vertOf) { oracle_patch_map_free(m); return NULL; }
    /* handles, array by array (far/patchMap.cpp:113-135) */
    for (a = 0; a < numArrays; ++a) {
        int pts = pm_type_points(arrays[a].desc);
        for (j = 0; j < arrays[a].numPatches && h < numPatches; ++j, ++h) { m->arrayOf[h] = a; m->vertOf[h] = j * pts; }
    }
    m->minFace = m->maxFace = (int)pp_face(params[0].field0);
    for (p = 1; p < numPatches; ++p) {
        int f = (int)pp_face(params[p].field0);
        if (f < m->minFace) m->minFace = f;
        if (f > m->maxFace) m->maxFace = f;
    }
    faces = m->maxFace - m->minFace + 1;
    for (p = 0; p < faces; ++p) if (pm_new_node(m) < 0) { oracle_patch_map_free(m); return NULL; }

    for (p = 0; p < numPatches; ++p) {
        unsigned f1 = params[p].field1;
        int depth = pp_depth(f1), root = pp_nonquad(f1) ? 1 : 0, level;
        int node = (int)pp_face(params[p].field0) - m->minFace;
        if (depth > m->maxDepth) m->maxDepth = depth;
        if (depth == root) {                               /* the whole face: all four quadrants are this leaf */
            for (j = 0; j < 4; ++j) m->nodes[node].child[j] = ((unsigned)p << 2) | 3u;
            continue;
        }
        {
            int pu = pp_u(f1), pv = pp_v(f1), rotated = 0;
            double u = 0.25, v = 0.25, median = 0.5;
            if (m->triangular) {                           /* interior point of the sub-triangle, far/patchParam.h:310-323 */
                double frac = (double)(1.0f / (float)(1 << (depth - root)));
                if ((pu + pv) >= (1 << depth)) { u = ((double)((1 << depth) - pu) - u) * frac; v = ((double)((1 << depth) - pv) - v) * frac; }
                else { u = (u + (double)pu) * frac; v = (v + (double)pv) * frac; }
            }
            for (level = root + 1; level <= depth; ++level, median *= 0.5) {
                int quadrant = m->triangular ? pm_tri_quadrant(median, &u, &v, &rotated)
                                             : ((((pv >> (depth - level)) & 1) << 1) | ((pu >> (depth - level)) & 1));
                if (level == depth) {
                    m->nodes[node].child[quadrant] = ((unsigned)p << 2) | 3u;
                } else if (m->nodes[node].child[quadrant] & 1u) {
                    node = (int)(m->nodes[node].child[quadrant] >> 2);
                } else {
                    int fresh = pm_new_node(m);
                    if (fresh < 0) { oracle_patch_map_free(m); return NULL; }
                    m->nodes[node].child[quadrant] = ((unsigned)fresh << 2) | 1u;
                    node = fresh;
                }
            }
        }
    }
    return m;
}

/* n queries; out[i] = PatchCoord{handle, s, t}; a face outside the map or a hole gives arrayIndex = -1 (the reference
 * returns a NULL handle).  Returns the number of hits. */
int oracle_patch_map_find(const void *h, int n, const int *ptexFace, const float *s, const float *t, oracle_coord *out)
{
    const oracle_patch_map *m = (const oracle_patch_map *)h;
    int i, hits = 0;
    for (i = 0; i < n; ++i) {
        oracle_coord c;
        memset(&c, 0, sizeof(c));
        c.arrayIndex = -1; c.s = s[i]; c.t = t[i];
        if (m && ptexFace[i] >= m->minFace && ptexFace[i] <= m->maxFace) {
            const oracle_qnode *node = &m->nodes[ptexFace[i] - m->minFace];
            if (node->child[0] & 1u) {                     /* a root has all quadrants set or none (hole) */
                double u = (double)s[i], v = (double)t[i], median = 0.5;
                int rotated = 0, depth;
                for (depth = 0; depth <= m->maxDepth; ++depth, median *= 0.5) {
                    int quadrant;
                    unsigned w;
                    if (m->triangular) quadrant = pm_tri_quadrant(median, &u, &v, &rotated);
                    else {
                        int uh = (u >= median), vh = (v >= median);
                        if (uh) u -= median;
                        if (vh) v -= median;
                        quadrant = (vh << 1) | uh;
                    }
                    w = node->child[quadrant];
                    if ((w & 3u) == 3u) {
                        int p = (int)(w >> 2);
                        c.arrayIndex = m->arrayOf[p]; c.patchIndex = p; c.vertIndex = m->vertOf[p];
                        ++hits;
                        break;
                    }
                    if (!(w & 1u)) break;
                    node = &m->nodes[w >> 2];
                }
            }
        }
        out[i] = c;
    }
    return hits;
}

/* ------------------------------------------------------------------------------------------------
 * Limit stencil table construction.  far/stencilTableFactory.cpp:559-662 (the per-location loop of
 * LimitStencilTableFactory::Create) + far/stencilBuilder.cpp:154-186,318-384,520-596 (AddWithWeight / merge).
 *
 * For every location that FindPatch resolves: the patch's basis weights (value, 1st, 2nd derivatives) are combined
 * with the stencils of the patch's control points -- a control point below numControlVerts is the control vertex
 * itself (a unit stencil), any other is row (cv - numControlVerts) of the refined + local-point stencil table, which
 * is already expressed in control vertices.  Each source element (index i, weight w != 0) contributes
 * (wP*w, wDs*w, ...) to the entry of control vertex i, entries are created in order of first appearance and later
 * contributions are ADDED to them, in source order -- which fixes both the element order of the result and the
 * floating-point summation order.  A control point whose nw basis weights are all zero is skipped.  Locations outside
 * every patch produce no stencil.  Adaptive (feature-adaptive) tables only.
 * This is the checker for SURVEY.md 8f-4 (table construction on the device); no product code uses it.
 * -----------------------------------------------------------------------------------------------*/
typedef struct {
    int n;                 /* stencils (= resolved locations) */
    int ne, cap;           /* elements */
    int nw;
    int *sizes, *offsets, *indices;
    float *w[6];
} oracle_limit_table;

void oracle_limit_table_free(void *h)
{
    oracle_limit_table *t = (oracle_limit_table *)h;
    int k;
    if (!t) return;
    free(t->sizes); free(t->offsets); free(t->indices);
    for (k = 0; k < 6; ++k) free(t->w[k]);
    free(t);
}

#ifndef ORACLE_F64
static int lt_grow(oracle_limit_table *t)
{
    int k, cap = t->cap ? 2 * t->cap : 4096;
    int *ix = (int *)realloc(t->indices, (size_t)cap * sizeof(int));
    if (!ix) return 0;
    t->indices = ix;
    for (k = 0; k < t->nw; ++k) {
        float *w = (float *)realloc(t->w[k], (size_t)cap * sizeof(float));
        if (!w) return 0;
        t->w[k] = w;
    }
    t->cap = cap;
    return 1;
}

void *oracle_limit_table_create(int nw, const void *patchMap,
                                const oracle_array *arrays, const int *patchIndices, const oracle_param *params,
                                int numControlVerts, const int *cvSizes, const int *cvOffsets, const int *cvIndices,
                                const float *cvWeights,
                                int numLocations, const int *ptexFace, const float *s, const float *t)
{
    oracle_limit_table *T = (oracle_limit_table *)calloc(1, sizeof(oracle_limit_table));
    oracle_coord c;
    float wb[6][20];
    int loc, j, k, q, e;
    if (!T || (nw != 1 && nw != 3 && nw != 6)) { free(T); return NULL; }
    T->nw = nw;
    T->sizes = (int *)malloc((size_t)(numLocations > 0 ? numLocations : 1) * sizeof(int));
    T->offsets = (int *)malloc((size_t)(numLocations > 0 ? numLocations : 1) * sizeof(int));
    if (!T->sizes || !T->offsets) { oracle_limit_table_free(T); return NULL; }
    for (loc = 0; loc < numLocations; ++loc) {
        const oracle_array *a;
        const oracle_param *p;
        const int *cvs;
        int type, np, start;
        if (oracle_patch_map_find(patchMap, 1, &ptexFace[loc], &s[loc], &t[loc], &c) != 1) continue;
        a = &arrays[c.arrayIndex];
        p = &params[c.patchIndex];
        type = pp_regular(p->field1) ? a->regDesc : a->desc;
        cvs = patchIndices + a->indexBase + a->stride * (c.patchIndex - a->primitiveIdBase);
        np = oracle_patch_basis(type, p->field0, p->field1, s[loc], t[loc], wb[0], nw >= 3 ? wb[1] : NULL,
                                nw >= 3 ? wb[2] : NULL, nw >= 6 ? wb[3] : NULL, nw >= 6 ? wb[4] : NULL, nw >= 6 ? wb[5] : NULL);
        start = T->ne;
        for (k = 0; k < np; ++k) {
            const int cv = cvs[k];
            const int unit = cv < numControlVerts;
            const int sz = unit ? 1 : cvSizes[cv - numControlVerts];
            const int off = unit ? 0 : cvOffsets[cv - numControlVerts];
            int allZero = 1;
            for (q = 0; q < nw; ++q) if (wb[q][k] != 0.0f) allZero = 0;
            if (allZero) continue;
            for (j = 0; j < sz; ++j) {
                const float w = unit ? 1.0f : cvWeights[off + j];
                const int src = unit ? cv : cvIndices[off + j];
                if (w == 0.0f) continue;
                for (e = start; e < T->ne; ++e) if (T->indices[e] == src) break;
                if (e == T->ne) {
                    if (T->ne == T->cap && !lt_grow(T)) { oracle_limit_table_free(T); return NULL; }
                    T->indices[e] = src;
                    for (q = 0; q < nw; ++q) T->w[q][e] = wb[q][k] * w;
                    T->ne++;
                } else {
                    for (q = 0; q < nw; ++q) T->w[q][e] += wb[q][k] * w;
                }
            }
        }
        T->sizes[T->n] = T->ne - start;
        T->offsets[T->n] = start;
        T->n++;
    }
    return T;
}
#endif

int oracle_limit_table_num_stencils(const void *h) { return ((const oracle_limit_table *)h)->n; }
int oracle_limit_table_num_elements(const void *h) { return ((const oracle_limit_table *)h)->ne; }
const int *oracle_limit_table_ints(const void *h, int which)   /* 0 sizes, 1 offsets, 2 indices */
{
    const oracle_limit_table *t = (const oracle_limit_table *)h;
    return which == 0 ? t->sizes : (which == 1 ? t->offsets : t->indices);
}
const float *oracle_limit_table_weights(const void *h, int k) { return ((const oracle_limit_table *)h)->w[k]; }

const char *oracle_version(void) { return "osd_oracle 1 (restates OpenSubdiv 3.6.0 osd/cpuKernel.cpp, cpuEvaluator.cpp, patchBasis.h)"; }
