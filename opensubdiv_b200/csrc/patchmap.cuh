// patchmap.cuh -- (ptexFace, s, t) -> patch handle: the sample-location step that feeds EvalPatches.
//
// Semantics restated from the reference (paths relative to /root/reference/opensubdiv):
//   far/patchMap.h:180-217     FindPatch: per-face quadtree descent, one quadrant per level, NULL for holes and
//                              for faces outside [minPatchFace, maxPatchFace]
//   far/patchMap.h:127-175     quadrant numbering for quad and (rotating) triangular sub-domains
//   far/patchMap.cpp:96-188    handle = (arrayIndex, patchIndex, vertIndex); a patch is filed under the quadrant
//                              path given by its PatchParam (depth, u, v; triangles: an interior point)
//   osd/types.h:53-54          PatchCoord = handle + (s, t)
//
// Design.  The tree is a flat array of 16-byte nodes (one int4 = the four children, fetched with ONE load per
// level); a child word is 0 = empty, ~patch (negative) = leaf, or a positive node index.  Node 'f - minFace' is
// the root of ptex face f.  For quad domains the descent is integer-only: s and t are scaled by 2^(maxDepth+1)
// (exact in fp32), truncated, and level k consumes bit (maxDepth-k) of each -- the same decisions as the
// reference's repeated "u >= median ? u -= median" in double (every operand is a dyadic multiple, every subtraction
// exact).  Triangular domains need the u+v test and a rotation state, so they descend in double like the reference.
// Everything is __host__ __device__: tests/emu runs the same descent and the same builder on the CPU against
// Far::PatchMap without a GPU; the library itself only ever runs the descent inside kernels.
#pragma once

#include "common.cuh"

#include <vector>

#ifndef B200_HD
#define B200_HD __host__ __device__ __forceinline__
#endif

namespace b200osd {

struct PatchMapView {
    const int4 *nodes;       // quadtree; roots first
    const int2 *handles;     // per patch: {arrayIndex, vertIndex}; patchIndex is the position itself
    int minFace, maxFace;    // maxFace < minFace: empty map
    int maxDepth;
    int triangular;
};

B200_HD int pm_child(const int4 &n, int q) {
    const int lo = (q & 1) ? n.y : n.x;
    const int hi = (q & 1) ? n.w : n.z;
    return (q & 2) ? hi : lo;
}

B200_HD int4 pm_load_node(const int4 *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// Quadrant of (u, v) in a triangular domain at the level whose half-size is 'median'; (u, v) is moved into the
// quadrant's frame.  Quadrant 3 is the centre triangle, which is parametrically rotated by 180 degrees
// (far/patchMap.h:146-175).
B200_HD int pm_tri_quadrant(double median, double &u, double &v, bool &rotated) {
    if (!rotated) {
        if (u >= median) { u -= median; return 1; }
        if (v >= median) { v -= median; return 2; }
        if (u + v >= median) { rotated = true; return 3; }
        return 0;
    }
    if (u < median) { v -= median; return 1; }
    if (v < median) { u -= median; return 2; }
    u -= median;
    v -= median;
    if (u + v < median) { rotated = false; return 3; }
    return 0;
}

// Returns the patch index (== handle index) containing (s, t) of ptex face 'face', or -1.
B200_HD int patch_map_find(const PatchMapView &m, int face, float s, float t) {
    if (face < m.minFace || face > m.maxFace) return -1;
    int4 node = pm_load_node(m.nodes + (face - m.minFace));
    if (node.x == 0) return -1;                       // hole: a root has all or none of its children
    if (!m.triangular) {
        const int levels = m.maxDepth + 1;
        const float scale = (float)(1 << levels);
        const int top = (1 << levels) - 1;
        int iu = (int)(s * scale), iv = (int)(t * scale);
        iu = iu < 0 ? 0 : (iu > top ? top : iu);
        iv = iv < 0 ? 0 : (iv > top ? top : iv);
        for (int bit = levels - 1; bit >= 0; --bit) {
            const int q = (((iv >> bit) & 1) << 1) | ((iu >> bit) & 1);
            const int c = pm_child(node, q);
            if (c < 0) return ~c;
            if (c == 0) return -1;                    // malformed tree (the reference asserts)
            node = pm_load_node(m.nodes + c);
        }
        return -1;
    }
    double u = (double)s, v = (double)t, median = 0.5;
    bool rotated = false;
    for (int depth = 0; depth <= m.maxDepth; ++depth, median *= 0.5) {
        const int q = pm_tri_quadrant(median, u, v, rotated);
        const int c = pm_child(node, q);
        if (c < 0) return ~c;
        if (c == 0) return -1;
        node = pm_load_node(m.nodes + c);
    }
    return -1;
}

// ------------------------------------------------------------------------------------ host build --
struct PatchMapHost {
    std::vector<int4> nodes;
    std::vector<int2> handles;
    int minFace = 0, maxFace = -1, maxDepth = 0, triangular = 0;
};

inline int pm_points_of_type(int type) {
    switch (type) {
        case 1: return 1;    // POINTS
        case 2: return 2;    // LINES
        case 3: return 4;    // QUADS
        case 4: return 3;    // TRIANGLES
        case 5: return 12;   // LOOP
        case 6: return 16;   // REGULAR
        case 7: return 4;    // GREGORY (legacy)
        case 8: return 4;    // GREGORY_BOUNDARY (legacy)
        case 9: return 20;   // GREGORY_BASIS
        case 10: return 18;  // GREGORY_TRIANGLE
        default: return -1;
    }
}

// Builds the tree from the vertex PatchArray[] and PatchParam[] of a flattened patch table.  Returns 0, or a
// negative code: -1 arrays do not tile the parameter table, -2 two patches claim the same cell.
inline int build_patch_map(int numArrays, const b200osd_patch_array *arrays, int numPatches,
                           const b200osd_patch_param *params, int triangular, PatchMapHost *out) {
    PatchMapHost &m = *out;
    m = PatchMapHost();
    m.triangular = triangular ? 1 : 0;
    if (numPatches <= 0 || numArrays <= 0) return 0;

    m.handles.resize((size_t)numPatches);
    int h = 0;
    for (int a = 0; a < numArrays; ++a) {
        const int pts = pm_points_of_type(arrays[a].desc);
        if (arrays[a].primitiveIdBase != h || h + arrays[a].numPatches > numPatches || pts < 0) return -1;
        for (int j = 0; j < arrays[a].numPatches; ++j, ++h) m.handles[(size_t)h] = make_int2(a, j * pts);
    }
    if (h != numPatches) return -1;

    auto faceOf = [&](int p) { return (int)(params[p].field0 & 0x0fffffffu); };
    m.minFace = m.maxFace = faceOf(0);
    for (int p = 1; p < numPatches; ++p) {
        const int f = faceOf(p);
        m.minFace = f < m.minFace ? f : m.minFace;
        m.maxFace = f > m.maxFace ? f : m.maxFace;
    }
    const int numFaces = m.maxFace - m.minFace + 1;
    m.nodes.assign((size_t)numFaces, make_int4(0, 0, 0, 0));
    m.nodes.reserve((size_t)numFaces + (size_t)numPatches);

    for (int p = 0; p < numPatches; ++p) {
        const unsigned f1 = params[p].field1;
        const int depth = (int)(f1 & 0xfu), root = (int)((f1 >> 4) & 1u);
        const int pv = (int)((f1 >> 12) & 0x3ffu), pu = (int)((f1 >> 22) & 0x3ffu);
        m.maxDepth = depth > m.maxDepth ? depth : m.maxDepth;
        size_t cur = (size_t)(faceOf(p) - m.minFace);
        const int steps = depth - root;
        if (steps <= 0) {                              // the patch is the whole face
            int4 &n = m.nodes[cur];
            if (n.x | n.y | n.z | n.w) return -2;
            n = make_int4(~p, ~p, ~p, ~p);
            continue;
        }
        // quadrant path, root first
        int path[16];
        if (!m.triangular) {
            for (int k = 0; k < steps; ++k) {
                const int sh = steps - 1 - k;
                path[k] = (((pv >> sh) & 1) << 1) | ((pu >> sh) & 1);
            }
        } else {
            // an interior point of the sub-triangle, mapped back to the face's domain (far/patchParam.h:310-323)
            const double frac = (double)(1.0f / (float)(1 << steps));
            const int df = 1 << depth;
            double u, v;
            if (pu + pv >= df) { u = ((double)(df - pu) - 0.25) * frac; v = ((double)(df - pv) - 0.25) * frac; }
            else { u = (0.25 + (double)pu) * frac; v = (0.25 + (double)pv) * frac; }
            double median = 0.5;
            bool rotated = false;
            for (int k = 0; k < steps; ++k, median *= 0.5) path[k] = pm_tri_quadrant(median, u, v, rotated);
        }
        for (int k = 0; k < steps; ++k) {
            int *child = &m.nodes[cur].x + path[k];    // int4 is four consecutive ints
            if (k == steps - 1) {
                if (*child != 0) return -2;
                *child = ~p;
            } else if (*child > 0) {
                cur = (size_t)*child;
            } else if (*child == 0) {
                const int fresh = (int)m.nodes.size();
                *child = fresh;                        // before push_back: it may reallocate
                m.nodes.push_back(make_int4(0, 0, 0, 0));
                cur = (size_t)fresh;
            } else {
                return -2;                             // a coarser patch already owns this cell
            }
        }
    }
    return 0;
}

}  // namespace b200osd
