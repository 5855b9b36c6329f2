// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// C-ABI shim over the UNMODIFIED OpenSubdiv 3.6.0 reference (compiled in place
// from /root/reference by oracle/ref/Makefile into oracle/_ref/libosdref.so).
// It exists so that Python tests / bench baselines can
//   (a) build real Far tables (StencilTable, LimitStencilTable, PatchTable,
//       PatchMap handles) for regression shapes and synthetic meshes, and
//   (b) run the reference's own Osd::CpuEvaluator / Osd::OmpEvaluator on them
//       (the ground truth every B200 kernel is compared against).
// Nothing in opensubdiv_b200/ (the product) may link or load this library.
//
// Reference entry points used (file:line relative to /root/reference):
//   Far::TopologyRefinerFactory<>::Create        opensubdiv/far/topologyRefinerFactory.h:84
//   Far::StencilTableFactory::Create             opensubdiv/far/stencilTableFactory.h:94
//   Far::StencilTableFactory::AppendLocalPoint.. opensubdiv/far/stencilTableFactory.h:128-178
//   Far::LimitStencilTableFactory::Create        opensubdiv/far/stencilTableFactory.h:269
//   Far::PatchTableFactory::Create               opensubdiv/far/patchTableFactory.h:175
//   Far::PatchMap::FindPatch                     opensubdiv/far/patchMap.h:74
//   Osd::CpuPatchTable                           opensubdiv/osd/cpuPatchTable.cpp:35
//   Osd::CpuEvaluator::EvalStencils/EvalPatches  opensubdiv/osd/cpuEvaluator.cpp:37-381
//   Osd::OmpEvaluator::EvalStencils/EvalPatches  opensubdiv/osd/ompEvaluator.cpp

#include <opensubdiv/far/topologyDescriptor.h>
#include <opensubdiv/far/topologyRefinerFactory.h>
#include <opensubdiv/far/primvarRefiner.h>
#include <opensubdiv/far/stencilTable.h>
#include <opensubdiv/far/stencilTableFactory.h>
#include <opensubdiv/far/patchTable.h>
#include <opensubdiv/far/patchTableFactory.h>
#include <opensubdiv/far/patchMap.h>
#include <opensubdiv/far/ptexIndices.h>
#include <opensubdiv/osd/cpuEvaluator.h>
#include <opensubdiv/osd/cpuPatchTable.h>
#include <opensubdiv/osd/bufferDescriptor.h>
#include <opensubdiv/osd/types.h>
#include <opensubdiv/osd/patchBasis.h>
#ifdef OPENSUBDIV_HAS_OPENMP
#include <opensubdiv/osd/ompEvaluator.h>
#include <omp.h>
#endif

#include <far_utils.h>    // -I$(REF)/regression/common : TopologyRefinerFactory<Shape>, GetSdcOptions
#include <shapes/all.h>   // -I$(REF)/regression        : the embedded regression OBJ strings

#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace OpenSubdiv;

namespace {

struct RefMesh {
    Far::TopologyRefiner *refiner = nullptr;
    std::vector<float> positions;   // level-0 xyz
    std::vector<float> uvs;         // level-0 fvar channel 0 values (2 floats) if any
    bool hasUV = false;
    int regFaceSize = 4;
};

struct RefStencils {
    Far::StencilTableReal<float> const *st = nullptr;   // owned (base view of lst when limit)
    Far::LimitStencilTable const *lst = nullptr;
};

struct RefPatches {
    Far::PatchTable const *pt = nullptr;
    Osd::CpuPatchTable *cpu = nullptr;
    Far::PatchMap *map = nullptr;
};

std::map<std::string, ShapeDesc> &shapeRegistry() {
    static std::map<std::string, ShapeDesc> reg;
    if (reg.empty()) {
#define REG(name, scheme) reg.insert(std::make_pair(std::string(#name), ShapeDesc(#name, name, scheme)))
        REG(catmark_cube, kCatmark);
        REG(catmark_cube_corner0, kCatmark);
        REG(catmark_cube_corner1, kCatmark);
        REG(catmark_cube_corner2, kCatmark);
        REG(catmark_cube_corner3, kCatmark);
        REG(catmark_cube_corner4, kCatmark);
        REG(catmark_cube_creases0, kCatmark);
        REG(catmark_cube_creases1, kCatmark);
        REG(catmark_cube_creases2, kCatmark);
        REG(catmark_dart_edgecorner, kCatmark);
        REG(catmark_dart_edgeonly, kCatmark);
        REG(catmark_edgecorner, kCatmark);
        REG(catmark_edgeonly, kCatmark);
        REG(catmark_edgenone, kCatmark);
        REG(catmark_fan, kCatmark);
        REG(catmark_flap, kCatmark);
        REG(catmark_flap2, kCatmark);
        REG(catmark_fvar_bound0, kCatmark);
        REG(catmark_fvar_bound1, kCatmark);
        REG(catmark_fvar_bound2, kCatmark);
        REG(catmark_gregory_test0, kCatmark);
        REG(catmark_gregory_test1, kCatmark);
        REG(catmark_gregory_test2, kCatmark);
        REG(catmark_gregory_test3, kCatmark);
        REG(catmark_gregory_test4, kCatmark);
        REG(catmark_gregory_test5, kCatmark);
        REG(catmark_gregory_test6, kCatmark);
        REG(catmark_gregory_test7, kCatmark);
        REG(catmark_helmet, kCatmark);
        REG(catmark_pawn, kCatmark);
        REG(catmark_pole8, kCatmark);
        REG(catmark_pole64, kCatmark);
        REG(catmark_pole360, kCatmark);
        REG(catmark_pyramid, kCatmark);
        REG(catmark_pyramid_creases0, kCatmark);
        REG(catmark_pyramid_creases1, kCatmark);
        REG(catmark_tent, kCatmark);
        REG(catmark_tent_creases0, kCatmark);
        REG(catmark_tent_creases1, kCatmark);
        REG(catmark_torus, kCatmark);
        REG(catmark_torus_creases0, kCatmark);
        REG(catmark_single_crease, kCatmark);
        REG(catmark_smoothtris0, kCatmark);
        REG(catmark_nonquads, kCatmark);
        REG(catmark_bishop, kCatmark);
        REG(catmark_car, kCatmark);
        REG(catmark_rook, kCatmark);
        REG(catmark_chaikin0, kCatmark);
        REG(catmark_hole_test1, kCatmark);
        REG(catmark_hole_test2, kCatmark);
        REG(catmark_hole_test3, kCatmark);
        REG(catmark_hole_test4, kCatmark);
        REG(catmark_square_hedit4, kCatmark);
        REG(loop_cube, kLoop);
        REG(loop_cube_creases0, kLoop);
        REG(loop_cube_creases1, kLoop);
        REG(loop_icosahedron, kLoop);
        REG(loop_saddle_edgecorner, kLoop);
        REG(loop_saddle_edgeonly, kLoop);
        REG(loop_triangle_edgecorner, kLoop);
        REG(loop_triangle_edgeonly, kLoop);
        REG(loop_pole8, kLoop);
        REG(loop_pole64, kLoop);
        REG(loop_toroidal_tet, kLoop);
        REG(loop_tetrahedron, kLoop);
        REG(bilinear_cube, kBilinear);
#undef REG
    }
    return reg;
}

Osd::BufferDescriptor D(const int *d) { return Osd::BufferDescriptor(d[0], d[1], d[2]); }

}  // namespace

extern "C" {

// ---------------------------------------------------------------- meshes ----

int ref_num_shapes() { return (int)shapeRegistry().size(); }

const char *ref_shape_name(int i) {
    auto &reg = shapeRegistry();
    auto it = reg.begin();
    std::advance(it, i);
    return it->first.c_str();
}

void *ref_mesh_from_shape(const char *name) {
    auto &reg = shapeRegistry();
    auto it = reg.find(name);
    if (it == reg.end()) return nullptr;
    Shape *shape = Shape::parseObj(it->second);
    if (!shape) return nullptr;
    typedef Far::TopologyRefinerFactory<Shape> Factory;
    RefMesh *m = new RefMesh;
    m->refiner = Factory::Create(*shape, Factory::Options(GetSdcType(*shape), GetSdcOptions(*shape)));
    m->positions = shape->verts;
    m->hasUV = shape->HasUV();
    if (m->hasUV) m->uvs = shape->uvs;
    m->regFaceSize = (shape->scheme == kLoop) ? 3 : 4;
    delete shape;
    if (!m->refiner) { delete m; return nullptr; }
    return m;
}

// `copies` replicas of a regression shape laid out along x (vertices, UVs, faces, crease / corner tags replicated
// with index offsets): a large adaptive mesh with extraordinary vertices, creases and UV seams for config 4.
void *ref_mesh_from_shape_tiled(const char *name, int copies) {
    auto &reg = shapeRegistry();
    auto it = reg.find(name);
    if (it == reg.end() || copies < 1) return nullptr;
    Shape *base = Shape::parseObj(it->second);
    if (!base) return nullptr;
    Shape big;
    big.scheme = base->scheme;
    big.isLeftHanded = base->isLeftHanded;
    const int nv = base->GetNumVertices(), nuv = (int)base->uvs.size() / 2;
    float xmin = 1e30f, xmax = -1e30f;
    for (int v = 0; v < nv; ++v) { xmin = std::min(xmin, base->verts[3 * v]); xmax = std::max(xmax, base->verts[3 * v]); }
    const float pitch = (xmax - xmin) * 1.25f + 1.0f;
    for (int c = 0; c < copies; ++c) {
        for (int v = 0; v < nv; ++v) {
            big.verts.push_back(base->verts[3 * v] + pitch * c);
            big.verts.push_back(base->verts[3 * v + 1]);
            big.verts.push_back(base->verts[3 * v + 2]);
        }
        big.uvs.insert(big.uvs.end(), base->uvs.begin(), base->uvs.end());
        big.nvertsPerFace.insert(big.nvertsPerFace.end(), base->nvertsPerFace.begin(), base->nvertsPerFace.end());
        for (size_t k = 0; k < base->faceverts.size(); ++k) big.faceverts.push_back(base->faceverts[k] + c * nv);
        for (size_t k = 0; k < base->faceuvs.size(); ++k) big.faceuvs.push_back(base->faceuvs[k] + c * nuv);
        for (size_t k = 0; k < base->tags.size(); ++k) {
            Shape::tag const *t = base->tags[k];
            const bool indexed = (t->name == "crease" || t->name == "corner");
            if (!indexed && c > 0) continue;                     // global options once
            Shape::tag *n = new Shape::tag(*t);
            if (indexed) for (size_t q = 0; q < n->intargs.size(); ++q) n->intargs[q] += c * nv;
            big.tags.push_back(n);
        }
    }
    typedef Far::TopologyRefinerFactory<Shape> Factory;
    RefMesh *m = new RefMesh;
    m->refiner = Factory::Create(big, Factory::Options(GetSdcType(big), GetSdcOptions(big)));
    m->positions = big.verts;
    m->hasUV = big.HasUV();
    if (m->hasUV) m->uvs = big.uvs;
    m->regFaceSize = (big.scheme == kLoop) ? 3 : 4;
    delete base;
    if (!m->refiner) { delete m; return nullptr; }
    return m;
}

// scheme: 0 bilinear, 1 catmark, 2 loop.  boundaryInterp: Sdc::Options::VtxBoundaryInterpolation.
// fvarLinearInterp: Sdc::Options::FVarLinearInterpolation.  fvarIndices may be NULL.
void *ref_mesh_from_topology(int scheme, int numVerts, int numFaces,
                             const int *vertsPerFace, const int *faceVerts,
                             int boundaryInterp, int fvarLinearInterp,
                             int numFVarValues, const int *fvarIndices,
                             int numCreases, const int *creasePairs, const float *creaseWeights,
                             int numCorners, const int *cornerVerts, const float *cornerWeights) {
    Far::TopologyDescriptor desc;
    desc.numVertices = numVerts;
    desc.numFaces = numFaces;
    desc.numVertsPerFace = vertsPerFace;
    desc.vertIndicesPerFace = faceVerts;
    desc.numCreases = numCreases;
    desc.creaseVertexIndexPairs = creasePairs;
    desc.creaseWeights = creaseWeights;
    desc.numCorners = numCorners;
    desc.cornerVertexIndices = cornerVerts;
    desc.cornerWeights = cornerWeights;
    Far::TopologyDescriptor::FVarChannel chan;
    if (fvarIndices && numFVarValues > 0) {
        chan.numValues = numFVarValues;
        chan.valueIndices = fvarIndices;
        desc.numFVarChannels = 1;
        desc.fvarChannels = &chan;
    }
    Sdc::Options sdc;
    sdc.SetVtxBoundaryInterpolation((Sdc::Options::VtxBoundaryInterpolation)boundaryInterp);
    sdc.SetFVarLinearInterpolation((Sdc::Options::FVarLinearInterpolation)fvarLinearInterp);
    Sdc::SchemeType st = scheme == 0 ? Sdc::SCHEME_BILINEAR : (scheme == 2 ? Sdc::SCHEME_LOOP : Sdc::SCHEME_CATMARK);
    typedef Far::TopologyRefinerFactory<Far::TopologyDescriptor> Factory;
    RefMesh *m = new RefMesh;
    m->refiner = Factory::Create(desc, Factory::Options(st, sdc));
    m->regFaceSize = (scheme == 2) ? 3 : 4;
    m->hasUV = (fvarIndices && numFVarValues > 0);
    if (!m->refiner) { delete m; return nullptr; }
    return m;
}

void ref_mesh_free(void *h) {
    RefMesh *m = (RefMesh *)h;
    if (!m) return;
    delete m->refiner;
    delete m;
}

int ref_mesh_num_base_verts(void *h) { return ((RefMesh *)h)->refiner->GetLevel(0).GetNumVertices(); }
int ref_mesh_num_base_faces(void *h) { return ((RefMesh *)h)->refiner->GetLevel(0).GetNumFaces(); }
int ref_mesh_reg_face_size(void *h) { return ((RefMesh *)h)->regFaceSize; }
int ref_mesh_num_fvar_channels(void *h) { return ((RefMesh *)h)->refiner->GetNumFVarChannels(); }
int ref_mesh_num_base_fvar_values(void *h, int ch) { return ((RefMesh *)h)->refiner->GetLevel(0).GetNumFVarValues(ch); }
int ref_mesh_num_verts_total(void *h) { return ((RefMesh *)h)->refiner->GetNumVerticesTotal(); }
int ref_mesh_max_level(void *h) { return ((RefMesh *)h)->refiner->GetMaxLevel(); }
int ref_mesh_level_num_verts(void *h, int level) { return ((RefMesh *)h)->refiner->GetLevel(level).GetNumVertices(); }
int ref_mesh_level_num_faces(void *h, int level) { return ((RefMesh *)h)->refiner->GetLevel(level).GetNumFaces(); }
const float *ref_mesh_positions(void *h) { RefMesh *m = (RefMesh *)h; return m->positions.empty() ? nullptr : m->positions.data(); }
int ref_mesh_num_uvs(void *h) { return (int)((RefMesh *)h)->uvs.size() / 2; }
const float *ref_mesh_uvs(void *h) { RefMesh *m = (RefMesh *)h; return m->uvs.empty() ? nullptr : m->uvs.data(); }

// Face-vertex connectivity of one refined level (used to validate synthetic generators).
int ref_mesh_level_face_verts(void *h, int level, int *out /* may be NULL -> returns count */) {
    Far::TopologyLevel const &lv = ((RefMesh *)h)->refiner->GetLevel(level);
    int n = 0;
    for (int f = 0; f < lv.GetNumFaces(); ++f) {
        Far::ConstIndexArray fv = lv.GetFaceVertices(f);
        for (int k = 0; k < fv.size(); ++k) {
            if (out) out[n] = fv[k];
            ++n;
        }
    }
    return n;
}

void ref_mesh_refine_uniform(void *h, int level, int fullTopologyInLastLevel) {
    RefMesh *m = (RefMesh *)h;
    Far::TopologyRefiner::UniformOptions o(level);
    o.fullTopologyInLastLevel = fullTopologyInLastLevel != 0;
    m->refiner->RefineUniform(o);
}

// Adaptive refinement with explicit options (glStencilViewer pattern, examples/glStencilViewer/glStencilViewer.cpp:340-345).
void ref_mesh_refine_adaptive(void *h, int level, int useSingleCreasePatch, int useInfSharpPatch, int considerFVarChannels) {
    RefMesh *m = (RefMesh *)h;
    Far::TopologyRefiner::AdaptiveOptions o(level);
    o.useSingleCreasePatch = useSingleCreasePatch != 0;
    o.useInfSharpPatch = useInfSharpPatch != 0;
    o.considerFVarChannels = considerFVarChannels != 0;
    m->refiner->RefineAdaptive(o);
}

int ref_mesh_num_ptex_faces(void *h) {
    Far::PtexIndices pi(*((RefMesh *)h)->refiner);
    return pi.GetNumFaces();
}

// --------------------------------------------------------------- patches ----

// endCap: Far::PatchTableFactory::Options::EndCapType.  If refineFirst != 0 the refiner is adaptively refined
// with options.GetRefineAdaptiveOptions() (far/patchTableFactory.h:100-107) before the table is built.
void *ref_patch_table_create(void *meshH, int level, int endCap, int generateFVarTables,
                             int fvarLegacyLinear, int useInfSharpPatch, int useSingleCreasePatch,
                             int legacySharpCorner, int refineFirst) {
    RefMesh *m = (RefMesh *)meshH;
    Far::PatchTableFactory::Options o(level);
    o.SetEndCapType((Far::PatchTableFactory::Options::EndCapType)endCap);
    o.generateFVarTables = generateFVarTables != 0;
    o.generateFVarLegacyLinearPatches = fvarLegacyLinear != 0;
    o.useInfSharpPatch = useInfSharpPatch != 0;
    o.useSingleCreasePatch = useSingleCreasePatch != 0;
    o.generateLegacySharpCornerPatches = legacySharpCorner != 0;
    if (refineFirst) m->refiner->RefineAdaptive(o.GetRefineAdaptiveOptions());
    RefPatches *p = new RefPatches;
    p->pt = Far::PatchTableFactory::Create(*m->refiner, o);
    if (!p->pt) { delete p; return nullptr; }
    p->cpu = new Osd::CpuPatchTable(p->pt);
    p->map = new Far::PatchMap(*p->pt);
    return p;
}

void ref_patch_table_free(void *h) {
    RefPatches *p = (RefPatches *)h;
    if (!p) return;
    delete p->map;
    delete p->cpu;
    delete p->pt;
    delete p;
}

// which: 0 vertex, 1 varying, 2+ch face-varying channel ch
int ref_patch_table_num_arrays(void *h, int which) {
    RefPatches *p = (RefPatches *)h;
    (void)which;
    return (int)p->cpu->GetNumPatchArrays();
}
const void *ref_patch_table_arrays(void *h, int which) {
    RefPatches *p = (RefPatches *)h;
    if (which == 0) return p->cpu->GetPatchArrayBuffer();
    if (which == 1) return p->cpu->GetVaryingPatchArrayBuffer();
    return p->cpu->GetFVarPatchArrayBuffer(which - 2);
}
int ref_patch_table_num_indices(void *h, int which) {
    RefPatches *p = (RefPatches *)h;
    if (which == 0) return (int)p->cpu->GetPatchIndexSize();
    if (which == 1) return (int)p->cpu->GetVaryingPatchIndexSize();
    return (int)p->cpu->GetFVarPatchIndexSize(which - 2);
}
const int *ref_patch_table_indices(void *h, int which) {
    RefPatches *p = (RefPatches *)h;
    if (which == 0) return p->cpu->GetPatchIndexBuffer();
    if (which == 1) return p->cpu->GetVaryingPatchIndexBuffer();
    return p->cpu->GetFVarPatchIndexBuffer(which - 2);
}
int ref_patch_table_num_params(void *h, int which) {
    RefPatches *p = (RefPatches *)h;
    if (which <= 1) return (int)p->cpu->GetPatchParamSize();
    return (int)p->cpu->GetFVarPatchParamSize(which - 2);
}
const void *ref_patch_table_params(void *h, int which) {
    RefPatches *p = (RefPatches *)h;
    if (which <= 1) return p->cpu->GetPatchParamBuffer();
    return p->cpu->GetFVarPatchParamBuffer(which - 2);
}
int ref_patch_table_num_fvar_channels(void *h) { return ((RefPatches *)h)->cpu->GetNumFVarChannels(); }
int ref_patch_table_num_local_points(void *h) { return ((RefPatches *)h)->pt->GetNumLocalPoints(); }
int ref_patch_table_num_local_points_fvar(void *h, int ch) { return ((RefPatches *)h)->pt->GetNumLocalPointsFaceVarying(ch); }

// PatchMap::FindPatch for n samples; out is n x {arrayIndex, patchIndex, vertIndex, s, t} (Osd::PatchCoord, 20 B).
// Returns number of samples that hit a patch; misses (holes) get arrayIndex = -1.
int ref_patch_map_find(void *h, int n, const int *ptexFace, const float *s, const float *t, void *outCoords) {
    RefPatches *p = (RefPatches *)h;
    Osd::PatchCoord *out = (Osd::PatchCoord *)outCoords;
    int hits = 0;
    for (int i = 0; i < n; ++i) {
        Far::PatchTable::PatchHandle const *handle = p->map->FindPatch(ptexFace[i], s[i], t[i]);
        if (handle) {
            out[i] = Osd::PatchCoord(*handle, s[i], t[i]);
            ++hits;
        } else {
            out[i] = Osd::PatchCoord();
            out[i].handle.arrayIndex = -1;
        }
    }
    return hits;
}

// Far's own twin of the basis (far/patchTable.cpp:581-625 EvaluateBasis) for cross-checking the Osd mirror.
void ref_patch_table_far_basis(void *h, int n, const void *coords, float *wP, float *wDs, float *wDt,
                               float *wDss, float *wDst, float *wDtt /* each n x 20 */) {
    RefPatches *p = (RefPatches *)h;
    const Osd::PatchCoord *c = (const Osd::PatchCoord *)coords;
    for (int i = 0; i < n; ++i) {
        p->pt->EvaluateBasis(c[i].handle, c[i].s, c[i].t, wP + 20 * i, wDs + 20 * i, wDt + 20 * i,
                             wDss + 20 * i, wDst + 20 * i, wDtt + 20 * i);
    }
}

// -------------------------------------------------------------- stencils ----

// mode: 0 vertex, 1 varying, 2 face-varying (Far::StencilTableFactory::Mode).
// If patchH != NULL the matching local-point stencil table is appended (osd/mesh.h:639-659).
void *ref_stencil_table_create(void *meshH, int mode, int generateIntermediateLevels,
                               int factorizeIntermediateLevels, int fvarChannel, void *patchH) {
    RefMesh *m = (RefMesh *)meshH;
    Far::StencilTableFactory::Options o;
    o.interpolationMode = mode;
    o.generateOffsets = true;
    o.generateIntermediateLevels = generateIntermediateLevels != 0;
    o.factorizeIntermediateLevels = factorizeIntermediateLevels != 0;
    o.fvarChannel = fvarChannel;
    Far::StencilTable const *st = Far::StencilTableFactory::Create(*m->refiner, o);
    if (!st) return nullptr;
    if (patchH) {
        RefPatches *p = (RefPatches *)patchH;
        Far::StencilTable const *merged = nullptr;
        if (mode == 0) {
            if (p->pt->GetLocalPointStencilTable())
                merged = Far::StencilTableFactory::AppendLocalPointStencilTable(
                    *m->refiner, st, p->pt->GetLocalPointStencilTable());
        } else if (mode == 1) {
            if (p->pt->GetLocalPointVaryingStencilTable())
                merged = Far::StencilTableFactory::AppendLocalPointStencilTableVarying(
                    *m->refiner, st, p->pt->GetLocalPointVaryingStencilTable());
        } else {
            if (p->pt->GetLocalPointFaceVaryingStencilTable(fvarChannel))
                merged = Far::StencilTableFactory::AppendLocalPointStencilTableFaceVarying(
                    *m->refiner, st, p->pt->GetLocalPointFaceVaryingStencilTable(fvarChannel), fvarChannel);
        }
        if (merged) { delete st; st = merged; }
    }
    RefStencils *s = new RefStencils;
    s->st = st;
    return s;
}

// Limit stencils at explicit locations (far/stencilTableFactory.cpp:413-662).
// locations are given flat: for sample i, ptexFace[i], s[i], t[i]; consecutive samples with the same ptex face are grouped.
void *ref_limit_stencil_table_create(void *meshH, int n, const int *ptexFace, const float *s, const float *t,
                                     int gen1st, int gen2nd, void *patchH) {
    RefMesh *m = (RefMesh *)meshH;
    typedef Far::LimitStencilTableFactory::LocationArray LocationArray;
    Far::LimitStencilTableFactory::LocationArrayVec locs;
    int i = 0;
    while (i < n) {
        int j = i;
        while (j < n && ptexFace[j] == ptexFace[i]) ++j;
        LocationArray la;
        la.ptexIdx = ptexFace[i];
        la.numLocations = j - i;
        la.s = s + i;
        la.t = t + i;
        locs.push_back(la);
        i = j;
    }
    Far::LimitStencilTableFactory::Options o;
    o.generate1stDerivatives = gen1st != 0;
    o.generate2ndDerivatives = gen2nd != 0;
    Far::PatchTable const *pt = patchH ? ((RefPatches *)patchH)->pt : nullptr;
    Far::LimitStencilTable const *lst = Far::LimitStencilTableFactory::Create(*m->refiner, locs, nullptr, pt, o);
    if (!lst) return nullptr;
    RefStencils *st = new RefStencils;
    st->lst = lst;
    st->st = lst;
    return st;
}

void ref_stencil_table_free(void *h) {
    RefStencils *s = (RefStencils *)h;
    if (!s) return;
    if (s->lst) delete s->lst; else delete s->st;
    delete s;
}

int ref_stencil_table_num_stencils(void *h) { return ((RefStencils *)h)->st->GetNumStencils(); }
int ref_stencil_table_num_control_verts(void *h) { return ((RefStencils *)h)->st->GetNumControlVertices(); }
int ref_stencil_table_num_elements(void *h) { return (int)((RefStencils *)h)->st->GetControlIndices().size(); }
const int *ref_stencil_table_sizes(void *h) { return ((RefStencils *)h)->st->GetSizes().data(); }
const int *ref_stencil_table_offsets(void *h) { return ((RefStencils *)h)->st->GetOffsets().data(); }
const int *ref_stencil_table_indices(void *h) { return ((RefStencils *)h)->st->GetControlIndices().data(); }
const float *ref_stencil_table_weights(void *h) { return ((RefStencils *)h)->st->GetWeights().data(); }
// which: 1 du, 2 dv, 3 duu, 4 duv, 5 dvv; returns NULL when absent (empty vector)
const float *ref_stencil_table_deriv_weights(void *h, int which) {
    RefStencils *s = (RefStencils *)h;
    if (!s->lst) return nullptr;
    std::vector<float> const *v = nullptr;
    switch (which) {
        case 1: v = &s->lst->GetDuWeights(); break;
        case 2: v = &s->lst->GetDvWeights(); break;
        case 3: v = &s->lst->GetDuuWeights(); break;
        case 4: v = &s->lst->GetDuvWeights(); break;
        case 5: v = &s->lst->GetDvvWeights(); break;
        default: return nullptr;
    }
    return v->empty() ? nullptr : v->data();
}

// Third independent CPU implementation: Far::StencilTable::UpdateValues (far/stencilTable.h:648-674), xyz only.
namespace {
struct P3 {
    float p[3];
    void Clear() { p[0] = p[1] = p[2] = 0.0f; }
    void AddWithWeight(P3 const &s, float w) { p[0] += w * s.p[0]; p[1] += w * s.p[1]; p[2] += w * s.p[2]; }
};
}
void ref_stencil_table_update_values_xyz(void *h, const float *src, float *dst) {
    ((RefStencils *)h)->st->UpdateValues((const P3 *)src, (P3 *)dst);
}

// ------------------------------------------------------------ evaluators ----
// All descriptors are int[3] = {offset, length, stride}.  nw = number of weight streams (1, 3 or 6).
// impl: 0 = Osd::CpuEvaluator, 1 = Osd::OmpEvaluator.  Returns the evaluator's bool as int, -1 if impl unavailable.

int ref_eval_stencils(int impl, int nw,
                      const float *src, const int *srcDesc,
                      float *const *dsts, const int *dstDescs /* nw x 3 */,
                      const int *sizes, const int *offsets, const int *indices,
                      const float *const *weights /* nw */, int start, int end) {
#ifndef OPENSUBDIV_HAS_OPENMP
    if (impl == 1) return -1;
#endif
#define CALL(EV)                                                                                          \
    if (nw == 1)                                                                                          \
        return EV::EvalStencils(src, D(srcDesc), dsts[0], D(dstDescs), sizes, offsets, indices,           \
                                weights[0], start, end);                                                  \
    if (nw == 3)                                                                                          \
        return EV::EvalStencils(src, D(srcDesc), dsts[0], D(dstDescs), dsts[1], D(dstDescs + 3), dsts[2], \
                                D(dstDescs + 6), sizes, offsets, indices, weights[0], weights[1],         \
                                weights[2], start, end);                                                  \
    if (nw == 6)                                                                                          \
        return EV::EvalStencils(src, D(srcDesc), dsts[0], D(dstDescs), dsts[1], D(dstDescs + 3), dsts[2], \
                                D(dstDescs + 6), dsts[3], D(dstDescs + 9), dsts[4], D(dstDescs + 12),     \
                                dsts[5], D(dstDescs + 15), sizes, offsets, indices, weights[0],           \
                                weights[1], weights[2], weights[3], weights[4], weights[5], start, end);
    if (impl == 0) { CALL(Osd::CpuEvaluator) }
#ifdef OPENSUBDIV_HAS_OPENMP
    if (impl == 1) { CALL(Osd::OmpEvaluator) }
#endif
#undef CALL
    return -1;
}

int ref_eval_patches(int impl, int nw,
                     const float *src, const int *srcDesc,
                     float *const *dsts, const int *dstDescs /* nw x 3 */,
                     int numPatchCoords, const void *patchCoords,
                     const void *patchArrays, const int *patchIndices, const void *patchParams) {
#ifndef OPENSUBDIV_HAS_OPENMP
    if (impl == 1) return -1;
#endif
    const Osd::PatchCoord *pc = (const Osd::PatchCoord *)patchCoords;
    const Osd::PatchArray *pa = (const Osd::PatchArray *)patchArrays;
    const Osd::PatchParam *pp = (const Osd::PatchParam *)patchParams;
#define CALL(EV)                                                                                          \
    if (nw == 1)                                                                                          \
        return EV::EvalPatches(src, D(srcDesc), dsts[0], D(dstDescs), numPatchCoords, pc, pa,             \
                               patchIndices, pp);                                                         \
    if (nw == 3)                                                                                          \
        return EV::EvalPatches(src, D(srcDesc), dsts[0], D(dstDescs), dsts[1], D(dstDescs + 3), dsts[2],  \
                               D(dstDescs + 6), numPatchCoords, pc, pa, patchIndices, pp);                \
    if (nw == 6)                                                                                          \
        return EV::EvalPatches(src, D(srcDesc), dsts[0], D(dstDescs), dsts[1], D(dstDescs + 3), dsts[2],  \
                               D(dstDescs + 6), dsts[3], D(dstDescs + 9), dsts[4], D(dstDescs + 12),      \
                               dsts[5], D(dstDescs + 15), numPatchCoords, pc, pa, patchIndices, pp);
    if (impl == 0) { CALL(Osd::CpuEvaluator) }
#ifdef OPENSUBDIV_HAS_OPENMP
    if (impl == 1) { CALL(Osd::OmpEvaluator) }
#endif
#undef CALL
    return -1;
}

// The Osd basis mirror itself (osd/patchBasis.h:1555) for n coords; weights n x 20 each (any may be NULL in groups).
int ref_osd_patch_basis(int patchType, int field0, int field1, float s, float t,
                        float *wP, float *wDs, float *wDt, float *wDss, float *wDst, float *wDtt) {
    Osd::OsdPatchParam param = Osd::OsdPatchParamInit(field0, field1, 0.0f);
    return Osd::OsdEvaluatePatchBasis(patchType, param, s, t, wP, wDs, wDt, wDss, wDst, wDtt);
}

int ref_has_openmp() {
#ifdef OPENSUBDIV_HAS_OPENMP
    return 1;
#else
    return 0;
#endif
}

int ref_omp_max_threads() {
#ifdef OPENSUBDIV_HAS_OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ref_omp_set_threads(int n) {
#ifdef OPENSUBDIV_HAS_OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

const char *ref_version() { return "OpenSubdiv 3.6.0 (unmodified, compiled in place)"; }

}  // extern "C"
