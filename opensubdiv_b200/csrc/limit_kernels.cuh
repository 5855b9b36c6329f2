// limit_kernels.cuh -- limit-stencil TABLE CONSTRUCTION on the device.
//
// What it replaces: the per-location loop of Far::LimitStencilTableFactory::Create
// (/root/reference/opensubdiv/far/stencilTableFactory.cpp:559-662) together with the merge of Far's StencilBuilder
// (far/stencilBuilder.cpp:154-186,318-384,520-596) -- 2.1 s on one host core for the 1 M locations of BASELINE config 3:
//   for every location: the patch's basis weights (value, 1st, 2nd derivatives) are combined with the stencils of the
//   patch's control points -- a control point below numControlVertices is the control vertex itself (a unit stencil),
//   any other is row (cv - numControlVertices) of the refined + local-point stencil table.  Every source element
//   (index i, weight w != 0) contributes (wP*w, wDs*w, ...) to the entry of control vertex i; entries are created in
//   order of first appearance and later contributions are ADDED to them in source order.  A control point whose basis
//   weights are all zero is skipped; a location that hit no patch produces no row.
// The result is the same table, row for row and element for element; for Catmark patches the weights are bit-identical
// to Far's in all six streams, because (a) the basis is evaluated in the reference's own polynomial forms and order
// (osd/patchBasis.h:99-135 cubic B-spline, :204-245 tensor product, :138-200 boundary folding, :247-283 Bernstein,
// :332-490 Gregory; the product's evaluation kernel uses other, faster forms), (b) every contribution is one rounded
// product added to the running sum in source order, and (c) the translation unit is compiled without mul+add
// contraction.  Triangle patches (Loop) use the evaluation kernel's basis: same structure, weights equal to rounding.
//
// One warp per location: lane 0 evaluates the basis into shared memory; control points are then taken one after the
// other (the order is part of the result) while the up to 32 elements of a control point's stencil are merged in
// parallel -- they are distinct vertices, so they touch distinct entries.  Two passes: sizes, then (after a scan) fill.
#pragma once

#include "patch_kernels.cuh"

namespace b200osd {

constexpr int kLimitCap = 512;                  // distinct control vertices one limit stencil may reference
constexpr int kLimitWarps = 2;                  // warps per block (shared memory: cap * 4 * (1 + NW) bytes per warp)

struct LimitIO {
    int n;                                      // locations
    const b200osd_patch_coord *coords;          // located samples (arrayIndex < 0: no patch -> no row)
    const b200osd_patch_array *arrays;
    const int *patchIndices;
    const b200osd_patch_param *params;
    int numControlVertices;
    const int *cvSizes, *cvOffsets, *cvIndices; // refined + local-point stencils, reference layout
    const float *cvWeights;
    // pass 1 out
    int *sizeOfLocation;                        // [n] elements of the location's stencil (0: no row)
    int *resolved;                              // [n] 1 = the location produces a row
    // pass 2 in / out
    const int *offsetOfLocation;                // [n] exclusive scan of sizeOfLocation
    const int *rowOfLocation;                   // [n] exclusive scan of resolved
    int *sizes, *offsets, *indices;             // the table
    float *w[6];
    int *overflow;                              // set when a stencil exceeds kLimitCap entries
    int options;                                // kPatchOpt*: the patch table's evaluation options
};

// ---- the reference's polynomial forms (restated; see the header comment) ----
__device__ __forceinline__ void ref_bspline3(float t, float *b, float *d1, float *d2) {
    const float sixth = 1.0f / 6.0f;
    const float t2 = t * t, t3 = t * t2;
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

__device__ __forceinline__ void ref_bezier3(float t, float *b, float *d1, float *d2) {
    const float t2 = t * t, c = 1.0f - t, c2 = c * c;
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

__device__ __forceinline__ void ref_tensor4(const float *cs, const float *ct, float *w) {
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) w[4 * r + c] = cs[c] * ct[r];
}

__device__ __forceinline__ void ref_fold_line(float *w, int i0, int i1, int i2, int step) {
    for (int k = 0; k < 4; ++k, i0 += step, i1 += step, i2 += step) {
        w[i2] -= w[i0];
        w[i1] += w[i0] * 2.0f;
        w[i0] = 0.0f;
    }
}

// weights of all NW sets into w[set * 20 + point]; returns the number of points.  s, t are already normalised.
template <int NW>
__device__ int ref_patch_weights(int type, float s, float t, int boundary, int options, float *w) {
    constexpr int order = NW == 1 ? 0 : (NW == 3 ? 1 : 2);
    if (type == PT_REGULAR) {
        float bs[4], bt[4], ds[4], dt[4], dss[4], dtt[4];
        ref_bspline3(s, bs, order >= 1 ? ds : nullptr, order >= 2 ? dss : nullptr);
        ref_bspline3(t, bt, order >= 1 ? dt : nullptr, order >= 2 ? dtt : nullptr);
        ref_tensor4(bs, bt, w);
        if (order >= 1) { ref_tensor4(ds, bt, w + 20); ref_tensor4(bs, dt, w + 40); }
        if (order >= 2) { ref_tensor4(dss, bt, w + 60); ref_tensor4(ds, dt, w + 80); ref_tensor4(bs, dtt, w + 100); }
        if (boundary) {
            for (int k = 0; k < NW; ++k) {
                float *wk = w + 20 * k;
                if (boundary & 1) ref_fold_line(wk, 0, 4, 8, 1);      // t = 0 edge: row 0 -> rows 1,2
                if (boundary & 2) ref_fold_line(wk, 3, 2, 1, 4);      // s = 1 edge: col 3 -> cols 2,1
                if (boundary & 4) ref_fold_line(wk, 12, 8, 4, 1);     // t = 1 edge: row 3 -> rows 2,1
                if (boundary & 8) ref_fold_line(wk, 0, 1, 2, 4);      // s = 0 edge: col 0 -> cols 1,2
            }
        }
        return 16;
    }
    if (type == PT_GREGORY_BASIS) {
        const signed char COL[20] = { 0, 1, 0, 1, 1, 3, 3, 2, 2, 2, 3, 2, 3, 2, 2, 0, 0, 1, 1, 1 };
        const signed char ROW[20] = { 0, 0, 1, 1, 1, 0, 1, 0, 1, 1, 3, 3, 2, 2, 2, 3, 2, 3, 2, 2 };
        float bs[4], bt[4], ds[4], dt[4], dss[4], dtt[4], G[8], R[4];
        const float sc = 1.0f - s, tc = 1.0f - t;
        ref_bezier3(s, bs, order >= 1 ? ds : nullptr, order >= 2 ? dss : nullptr);
        ref_bezier3(t, bt, order >= 1 ? dt : nullptr, order >= 2 ? dtt : nullptr);
        const float a[4] = { s, t, sc, tc };
        const float den[4] = { s + t, sc + t, sc + tc, s + tc };
        for (int c = 0; c < 4; ++c) {
            const float r = (den[c] <= 0.0f) ? 1.0f : (1.0f / den[c]);
            G[2 * c] = a[c] * r;
            G[2 * c + 1] = 1.0f - a[c] * r;
            R[c] = r;
        }
        const bool trueDerivatives = (options & kPatchOptGregoryTrueDerivatives) != 0;
        for (int i = 0; i < 20; ++i) {
            const int col = COL[i], row = ROW[i], p = i % 5;
            if (p >= 3 && order >= 1 && trueDerivatives) {
                // far/patchBasis.cpp:483-530 (OPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES), operation for operation
                const float NDS[8] = { 1.0f, 0.0f, 0.0f, -1.0f, -1.0f, 0.0f, 0.0f, 1.0f };
                const float NDT[8] = { 0.0f, 1.0f, 1.0f, 0.0f, 0.0f, -1.0f, -1.0f, 0.0f };
                const float DDS[8] = { 1.0f, 1.0f, -1.0f, -1.0f, -1.0f, -1.0f, 1.0f, 1.0f };
                const float DDT[8] = { 1.0f, 1.0f, 1.0f, 1.0f, -1.0f, -1.0f, -1.0f, -1.0f };
                const int k = 2 * (i / 5) + (p - 3);
                const float g = G[k], D = R[i / 5];
                w[i] = bs[col] * bt[row] * g;
                const float g_s = (NDS[k] - DDS[k] * g) * D;
                const float g_t = (NDT[k] - DDT[k] * g) * D;
                w[20 + i] = (ds[col] * g + bs[col] * g_s) * bt[row];
                w[40 + i] = (dt[row] * g + bt[row] * g_t) * bs[col];
                if (order >= 2) {
                    const float invD2 = D * D;
                    const float g_ss = 2.0f * DDS[k] * invD2 * (g * DDS[k] - NDS[k]);
                    const float g_st = invD2 * (2.0f * g * DDS[k] * DDT[k] - NDS[k] * DDT[k] - NDT[k] * DDS[k]);
                    const float g_tt = 2.0f * DDT[k] * invD2 * (g * DDT[k] - NDT[k]);
                    w[60 + i] = (dss[col] * g + 2.0f * ds[col] * g_s + bs[col] * g_ss) * bt[row];
                    w[80 + i] = bt[row] * (bs[col] * g_st + ds[col] * g_t) + dt[row] * (ds[col] * g + bs[col] * g_s);
                    w[100 + i] = (dtt[row] * g + 2.0f * dt[row] * g_t + bt[row] * g_tt) * bs[col];
                }
            } else if (p >= 3) {
                const float g = G[2 * (i / 5) + (p - 3)];
                w[i] = bs[col] * bt[row] * g;
                if (order >= 1) { w[20 + i] = ds[col] * bt[row] * g; w[40 + i] = dt[row] * bs[col] * g; }
                if (order >= 2) { w[60 + i] = dss[col] * bt[row] * g; w[80 + i] = ds[col] * dt[row] * g; w[100 + i] = bs[col] * dtt[row] * g; }
            } else {
                w[i] = bs[col] * bt[row];
                if (order >= 1) { w[20 + i] = ds[col] * bt[row]; w[40 + i] = dt[row] * bs[col]; }
                if (order >= 2) { w[60 + i] = dss[col] * bt[row]; w[80 + i] = ds[col] * dt[row]; w[100 + i] = bs[col] * dtt[row]; }
            }
        }
        return 20;
    }
    if (type == PT_QUADS) {
        const float sc = 1.0f - s, tc = 1.0f - t;
        w[0] = sc * tc; w[1] = s * tc; w[2] = s * t; w[3] = sc * t;
        if (order >= 1) {
            w[20] = -tc; w[21] = tc; w[22] = t; w[23] = -t;
            w[40] = -sc; w[41] = -s; w[42] = s; w[43] = sc;
        }
        if (order >= 2) {
            for (int i = 0; i < 4; ++i) { w[60 + i] = 0.0f; w[100 + i] = 0.0f; }
            w[80] = 1.0f; w[81] = -1.0f; w[82] = 1.0f; w[83] = -1.0f;
        }
        return 4;
    }
    if (type == PT_LOOP || type == PT_GREGORY_TRIANGLE || type == PT_TRIANGLES) {
        float tw[NW][20];
        const int np = tri_weights<order>(type, s, t, boundary, tw);
        for (int k = 0; k < NW; ++k)
            for (int i = 0; i < np; ++i) w[20 * k + i] = tw[k][i];
        return np;
    }
    return 0;
}

template <int NW, bool FILL>
__global__ void __launch_bounds__(32 * kLimitWarps) limit_merge_kernel(LimitIO io) {
    extern __shared__ __align__(16) unsigned char limit_smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t perWarp = (size_t)kLimitCap * 4 * (1 + (FILL ? NW : 0)) + 20 * 6 * 4;
    unsigned char *base = limit_smem + (size_t)warp * perWarp;
    int *list = reinterpret_cast<int *>(base);                               // [cap] control-vertex indices in order of first appearance
    float *acc = reinterpret_cast<float *>(base + (size_t)kLimitCap * 4);    // [NW][cap] (FILL)
    float *wb = reinterpret_cast<float *>(base + (size_t)kLimitCap * 4 * (1 + (FILL ? NW : 0)));   // [6][20] basis weights
    const int totalWarps = gridDim.x * kLimitWarps;

    for (int loc = blockIdx.x * kLimitWarps + warp; loc < io.n; loc += totalWarps) {
        const int *cw = reinterpret_cast<const int *>(io.coords + loc);
        const int arrayIndex = __ldg(cw + 0), patchIndex = __ldg(cw + 1);
        if (arrayIndex < 0) {                                                // FindPatch == NULL: no row (warp-uniform)
            if (!FILL && lane == 0) { io.sizeOfLocation[loc] = 0; io.resolved[loc] = 0; }
            continue;
        }
        PatchIO pio;
        pio.arrays = io.arrays;
        pio.params = io.params;
        PatchSite ps;
        decode_patch_site(pio, arrayIndex, patchIndex, __int_as_float(__ldg(cw + 3)), __int_as_float(__ldg(cw + 4)), ps);
        int np = 0;
        __syncwarp();
        if (lane == 0) {
            np = ref_patch_weights<NW>(ps.type, ps.s, ps.t, ps.boundary, io.options, wb);
            // derivative scaling (osd/patchBasis.h:1568-1607): d1 = +-2^depth, d2 = sign * d1 * d1
            if (NW >= 3) {
                const float d1 = ps.d1;
                for (int i = 0; i < np; ++i) { wb[20 + i] *= d1; wb[40 + i] *= d1; }
                if (NW >= 6) {
                    const float d2 = ps.sign * d1 * d1;
                    for (int i = 0; i < np; ++i) { wb[60 + i] *= d2; wb[80 + i] *= d2; wb[100 + i] *= d2; }
                }
            }
        }
        np = __shfl_sync(FULL, np, 0);
        __syncwarp();
        int len = 0;
        bool over = false;
        for (int k = 0; k < np; ++k) {
            float wk[NW];
            bool allZero = true;
#pragma unroll
            for (int q = 0; q < NW; ++q) { wk[q] = wb[20 * q + k]; allZero = allZero && (wk[q] == 0.0f); }
            if (allZero) continue;
            const int cv = __ldg(io.patchIndices + ps.cvOffset + k);
            const bool unit = cv < io.numControlVertices;
            const int sz = unit ? 1 : __ldg(io.cvSizes + (cv - io.numControlVertices));
            const int off = unit ? 0 : __ldg(io.cvOffsets + (cv - io.numControlVertices));
            for (int j0 = 0; j0 < sz; j0 += 32) {
                const int j = j0 + lane;
                bool active = j < sz;
                int src = cv;
                float w = 1.0f;
                if (active && !unit) { src = __ldg(io.cvIndices + off + j); w = __ldg(io.cvWeights + off + j); }
                active = active && (w != 0.0f);
                int found = -1;
                const int upto = min(len, kLimitCap);
                for (int e = 0; e < upto; ++e)
                    if (list[e] == src) found = e;
                const bool isNew = active && found < 0;
                const unsigned newMask = __ballot_sync(FULL, isNew);
                const int pos = len + __popc(newMask & ((1u << lane) - 1u));
                if (isNew) {
                    if (pos < kLimitCap) {
                        list[pos] = src;
                        if (FILL) {
#pragma unroll
                            for (int q = 0; q < NW; ++q) acc[q * kLimitCap + pos] = __fmul_rn(wk[q], w);
                        }
                    } else {
                        over = true;
                    }
                } else if (active && FILL) {
#pragma unroll
                    for (int q = 0; q < NW; ++q) acc[q * kLimitCap + found] = __fadd_rn(acc[q * kLimitCap + found], __fmul_rn(wk[q], w));
                }
                len += __popc(newMask);
                __syncwarp();
            }
        }
        if (__any_sync(FULL, over) && lane == 0) *io.overflow = 1;
        const int n = min(len, kLimitCap);
        if (!FILL) {
            if (lane == 0) { io.sizeOfLocation[loc] = n; io.resolved[loc] = 1; }
        } else {
            const int o = io.offsetOfLocation[loc];
            if (lane == 0) {
                const int r = io.rowOfLocation[loc];
                io.sizes[r] = n;
                io.offsets[r] = o;
            }
            for (int e = lane; e < n; e += 32) {
                io.indices[o + e] = list[e];
#pragma unroll
                for (int q = 0; q < NW; ++q) io.w[q][o + e] = acc[q * kLimitCap + e];
            }
        }
        __syncwarp();
    }
}

// ---- exclusive scan of n ints (tiles of 4096; the tile totals are scanned by one block) ----
__global__ void __launch_bounds__(kScanThreads) scan_local_kernel(const int *in, int *out, int *tileSum, int n) {
    __shared__ int warpSum[kScanThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long i0 = (long long)blockIdx.x * kScanTile + (long long)tid * kScanItems;
    int v[kScanItems], sum = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
        v[q] = (i0 + q < n) ? in[i0 + q] : 0;
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
        warpSum[lane] = ws;
    }
    __syncthreads();
    int run = incl - sum + (warp > 0 ? warpSum[warp - 1] : 0);
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
        if (i0 + q < n) out[i0 + q] = run;
        run += v[q];
    }
    if (tid == kScanThreads - 1) tileSum[blockIdx.x] = warpSum[kScanThreads / 32 - 1];
}

__global__ void __launch_bounds__(kScanThreads) scan_top_kernel(int *tileSum, int numTiles, int *total) {
    __shared__ int warpSum[kScanThreads / 32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < numTiles; b0 += kScanThreads) {
        const int v = (b0 + tid < numTiles) ? tileSum[b0 + tid] : 0;
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
        if (b0 + tid < numTiles) tileSum[b0 + tid] = excl;
        __syncthreads();
        if (tid == 0) carry += warpSum[kScanThreads / 32 - 1];
        __syncthreads();
    }
    if (tid == 0) *total = carry;
}

__global__ void __launch_bounds__(256) scan_add_kernel(int *out, const int *tileSum, int n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tileSum[i / kScanTile];
}

}  // namespace b200osd
