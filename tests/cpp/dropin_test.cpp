// TEST PROGRAM (test infrastructure): the B200 classes used exactly the way reference clients use a backend --
// through Osd::Mesh<> (opensubdiv/osd/mesh.h:413-714) and the EvalOutput-style template calls of
// examples/glEvalLimit/glEvalLimit.cpp:263-472 and examples/glStencilViewer/glStencilViewer.cpp:177-250 --
// next to the reference's own CPU backend on the same topology, comparing the results.
//
// Built by `make -C oracle/ref dropin` against the unmodified reference headers + oracle/_ref/libosdref.so and
// the product's libb200osd.so.  Exit code 0 = parity, 1 = mismatch, 2 = no CUDA device.
#include <opensubdiv/far/topologyDescriptor.h>
#include <opensubdiv/far/topologyRefinerFactory.h>
#include <opensubdiv/far/stencilTableFactory.h>
#include <opensubdiv/far/patchTableFactory.h>
#include <opensubdiv/far/patchMap.h>
#include <opensubdiv/far/ptexIndices.h>
#include <opensubdiv/osd/cpuEvaluator.h>
#include <opensubdiv/osd/cpuPatchTable.h>
#include <opensubdiv/osd/cpuVertexBuffer.h>
#include <opensubdiv/osd/mesh.h>
#include <opensubdiv/osd/patchBasis.h>

#include <b200osd/b200Evaluator.h>
#include <b200osd/b200PatchTable.h>
#include <b200osd/b200PatchMap.h>
#include <b200osd/b200StencilTable.h>
#include <b200osd/b200VertexBuffer.h>

#include <far_utils.h>
#include <shapes/catmark_cube_creases0.h>
#include <shapes/catmark_pyramid.h>
#include <shapes/loop_icosahedron.h>

#include <cmath>
#include <cstring>
#include <cstdio>
#include <random>
#include <vector>

using namespace OpenSubdiv;

// The reference's own CPU classes cannot be given to Osd::Mesh directly (CpuPatchTable has no VertexBufferBinding,
// CpuVertexBuffer no BindVBO -- the reference pairs them with GL classes for that); two trivial adapters fix it.
class CpuPT : public Osd::CpuPatchTable {
public:
    typedef float *VertexBufferBinding;
    static CpuPT *Create(const Far::PatchTable *pt, void * = NULL) { return new CpuPT(pt); }
    explicit CpuPT(const Far::PatchTable *pt) : Osd::CpuPatchTable(pt) {}
};
class CpuVB {
public:
    static CpuVB *Create(int ne, int nv, void * = NULL) { CpuVB *b = new CpuVB; b->_b = Osd::CpuVertexBuffer::Create(ne, nv); return b; }
    ~CpuVB() { delete _b; }
    void UpdateData(const float *src, int start, int n, void * = NULL) { _b->UpdateData(src, start, n); }
    int GetNumElements() const { return _b->GetNumElements(); }
    int GetNumVertices() const { return _b->GetNumVertices(); }
    float *BindCpuBuffer() { return _b->BindCpuBuffer(); }
    float *BindVBO(void * = NULL) { return _b->BindCpuBuffer(); }
private:
    Osd::CpuVertexBuffer *_b;
};
typedef Osd::Mesh<CpuVB, Far::StencilTable, Osd::CpuEvaluator, CpuPT> CpuMesh;
typedef Osd::Mesh<Osd::B200VertexBuffer, Osd::B200StencilTable, Osd::B200Evaluator, Osd::B200PatchTable> B200Mesh;

static Far::TopologyRefiner *makeRefiner(std::string const &str, Scheme scheme) {
    Shape *shape = Shape::parseObj(str.c_str(), scheme);
    Far::TopologyRefiner *r = Far::TopologyRefinerFactory<Shape>::Create(
        *shape, Far::TopologyRefinerFactory<Shape>::Options(GetSdcType(*shape), GetSdcOptions(*shape)));
    delete shape;
    return r;
}

static std::vector<float> positionsOf(std::string const &str, Scheme scheme) {
    Shape *shape = Shape::parseObj(str.c_str(), scheme);
    std::vector<float> p = shape->verts;
    delete shape;
    return p;
}

static double maxRelDiff(const float *a, const float *b, size_t n) {
    double scale = 1e-30, worst = 0;
    for (size_t i = 0; i < n; ++i) scale = std::max(scale, (double)std::fabs(b[i]));
    for (size_t i = 0; i < n; ++i) worst = std::max(worst, std::fabs((double)a[i] - b[i]) / scale);
    return worst;
}

// Error of `got` against `ref` relative to max(|ref|, S) where S = sum_j (|w_j| + max_j|w_j|) |x_j| is the magnitude of
// the terms the reference itself sums for that output (derivative weights are signed and cancel, boundary folding
// subtracts phantom weights: every fp32 evaluation order, the reference's included, carries ~eps*S of rounding).
static double maxCondDiff(const float *got, const float *ref, int n, int k, std::vector<Osd::PatchCoord> const &coords,
                          Osd::CpuPatchTable const *pt, const float *src) {
    double worst = 0;
    for (int i = 0; i < n; ++i) {
        Osd::PatchCoord const &c = coords[i];
        Osd::PatchArray const &a = pt->GetPatchArrayBuffer()[c.handle.arrayIndex];
        Osd::PatchParam const &p = pt->GetPatchParamBuffer()[c.handle.patchIndex];
        Osd::OsdPatchParam param = Osd::OsdPatchParamInit(p.field0, p.field1, p.sharpness);
        int type = Osd::OsdPatchParamIsRegular(param) ? a.GetPatchTypeRegular() : a.GetPatchTypeIrregular();
        float w[6][20];
        int np = Osd::OsdEvaluatePatchBasis(type, param, c.s, c.t, w[0], w[1], w[2], w[3], w[4], w[5]);
        const int *cvs = pt->GetPatchIndexBuffer() + a.GetIndexBase() + a.GetStride() * (c.handle.patchIndex - a.GetPrimitiveIdBase());
        double wmax = 0;
        for (int j = 0; j < np; ++j) wmax = std::max(wmax, (double)std::fabs(w[k][j]));
        for (int comp = 0; comp < 3; ++comp) {
            double S = 0;
            for (int j = 0; j < np; ++j) S += (std::fabs(w[k][j]) + wmax) * std::fabs(src[3 * cvs[j] + comp]);
            double den = std::max(std::max((double)std::fabs(ref[3 * i + comp]), S), 1e-30);
            worst = std::max(worst, std::fabs((double)got[3 * i + comp] - ref[3 * i + comp]) / den);
        }
    }
    return worst;
}

static int g_fail = 0;
static void report(const char *what, double err, double tol) {
    std::printf("%-58s max rel diff %.3e  (tol %.0e)  %s\n", what, err, tol, err <= tol ? "ok" : "MISMATCH");
    if (!(err <= tol)) g_fail = 1;
}

// ---- Osd::Mesh::Refine() through both back ends --------------------------------------------------------------
static int meshCase(const char *name, std::string const &str, Scheme scheme, int level, bool adaptive) {
    std::vector<float> pos = positionsOf(str, scheme);
    int nCV = (int)pos.size() / 3;
    Osd::MeshBitset bits;
    bits.set(Osd::MeshAdaptive, adaptive);
    bits.set(Osd::MeshEndCapGregoryBasis, adaptive);

    CpuMesh cpu(makeRefiner(str, scheme), 3, 0, level, bits);
    B200Mesh *gpu = new B200Mesh(makeRefiner(str, scheme), 3, 0, level, bits);
    if (!gpu->GetVertexBuffer()) { std::printf("no CUDA device: %s\n", b200osd_last_error()); return 2; }

    cpu.UpdateVertexBuffer(&pos[0], 0, nCV);
    gpu->UpdateVertexBuffer(&pos[0], 0, nCV);
    cpu.Refine();
    gpu->Refine();
    gpu->Synchronize();

    int nv = cpu.GetNumVertices();
    if (nv != gpu->GetNumVertices()) { std::printf("%s: vertex count differs\n", name); g_fail = 1; return 1; }
    std::vector<float> got((size_t)nv * 3);
    gpu->GetVertexBuffer()->ReadData(&got[0], 0, nv);
    Osd::B200Evaluator::Synchronize();
    char label[128];
    std::snprintf(label, sizeof(label), "Osd::Mesh::Refine %s level %d %s (%d verts)", name, level, adaptive ? "adaptive" : "uniform", nv);
    report(label, maxRelDiff(&got[0], cpu.GetVertexBuffer()->BindCpuBuffer(), got.size()), 1e-6);

    if (adaptive) {
        // glEvalLimit pattern: PatchCoords travel in a 5-float "vertex buffer"; P, du, dv interleaved in one output buffer
        Far::PatchTable const *farPt = cpu.GetFarPatchTable();
        Far::PatchMap patchMap(*farPt);
        Far::PtexIndices ptex(*cpu.GetTopologyRefiner());
        std::mt19937 rng(2024);
        std::uniform_real_distribution<float> uni(0.0f, 1.0f);
        std::vector<Osd::PatchCoord> coords;
        std::vector<Osd::B200PatchMap::Sample> samples;
        std::vector<Osd::PatchCoord> located;           // one record per sample, arrayIndex = -1 where FindPatch is NULL
        for (int i = 0; i < 20000; ++i) {
            int face = (int)(rng() % (ptex.GetNumFaces() + 2)) - 1;          // includes faces outside the table
            float s = uni(rng), t = uni(rng);
            if (i % 7 == 0) s = (float)(rng() % 17) / 16.0f;                 // exactly on sub-patch boundaries
            if (i % 11 == 0) t = (float)(rng() % 17) / 16.0f;
            if (scheme == kLoop && s + t >= 1.0f) { s = 1.0f - s; t = 1.0f - t; }
            Far::PatchTable::PatchHandle const *h = patchMap.FindPatch(face, s, t);
            if (h) coords.push_back(Osd::PatchCoord(*h, s, t));
            Osd::B200PatchMap::Sample smp = { face, s, t };
            samples.push_back(smp);
            Osd::PatchCoord rec;
            if (h) rec = Osd::PatchCoord(*h, s, t); else { rec.handle.arrayIndex = -1; rec.handle.patchIndex = 0; rec.handle.vertIndex = 0; rec.s = s; rec.t = t; }
            located.push_back(rec);
        }
        int n = (int)coords.size();
        {
            // Far::PatchMap::FindPatch per sample on the host  vs  B200PatchMap::FindPatches, one launch on the device
            int ns = (int)samples.size();
            Osd::B200PatchMap *gpuMap = Osd::B200PatchMap::Create(farPt);
            Osd::B200VertexBuffer *gpuSamples = Osd::B200VertexBuffer::Create(3, ns);
            Osd::B200VertexBuffer *gpuLocated = Osd::B200VertexBuffer::Create(5, ns);
            gpuSamples->UpdateData((const float *)&samples[0], 0, ns);
            bool ok = gpuMap && gpuMap->FindPatches(ns, gpuSamples, gpuLocated);
            std::vector<Osd::PatchCoord> back((size_t)ns);
            gpuLocated->ReadData((float *)&back[0], 0, ns);
            Osd::B200Evaluator::Synchronize();
            int bad = ok ? 0 : ns;
            for (int i = 0; ok && i < ns; ++i) bad += std::memcmp(&back[i], &located[i], sizeof(Osd::PatchCoord)) != 0;
            std::snprintf(label, sizeof(label), "B200PatchMap::FindPatches %s (%d samples, %d hits)", name, ns, n);
            report(label, (double)bad, 0.0);
            delete gpuMap; delete gpuSamples; delete gpuLocated;
        }
        Osd::CpuVertexBuffer *cpuCoords = Osd::CpuVertexBuffer::Create(5, n);
        Osd::B200VertexBuffer *gpuCoords = Osd::B200VertexBuffer::Create(5, n);
        cpuCoords->UpdateData((const float *)&coords[0], 0, n);
        gpuCoords->UpdateData((const float *)&coords[0], 0, n);
        Osd::CpuVertexBuffer *cpuOut = Osd::CpuVertexBuffer::Create(18, n);
        Osd::B200VertexBuffer *gpuOut = Osd::B200VertexBuffer::Create(18, n);
        Osd::BufferDescriptor src(0, 3, 3), p(0, 3, 18), du(3, 3, 18), dv(6, 3, 18), duu(9, 3, 18), duv(12, 3, 18), dvv(15, 3, 18);
        bool a = Osd::CpuEvaluator::EvalPatches(cpu.GetVertexBuffer(), src, cpuOut, p, cpuOut, du, cpuOut, dv, cpuOut, duu,
                                                cpuOut, duv, cpuOut, dvv, n, cpuCoords, cpu.GetPatchTable(),
                                                (Osd::CpuEvaluator const *)NULL);
        bool b = Osd::B200Evaluator::EvalPatches(gpu->GetVertexBuffer(), src, gpuOut, p, gpuOut, du, gpuOut, dv, gpuOut, duu,
                                                 gpuOut, duv, gpuOut, dvv, n, gpuCoords, gpu->GetPatchTable(),
                                                 (Osd::B200Evaluator const *)NULL);
        if (a != b) { std::printf("EvalPatches return values differ\n"); g_fail = 1; }
        std::vector<float> out((size_t)n * 18);
        gpuOut->ReadData(&out[0], 0, n);
        Osd::B200Evaluator::Synchronize();
        const float *ref = cpuOut->BindCpuBuffer();
        // per-output comparison (each derivative order has its own magnitude)
        for (int k = 0; k < 6; ++k) {
            std::vector<float> x((size_t)n * 3), y((size_t)n * 3);
            for (int i = 0; i < n; ++i)
                for (int c = 0; c < 3; ++c) { x[3 * i + c] = out[18 * i + 3 * k + c]; y[3 * i + c] = ref[18 * i + 3 * k + c]; }
            static const char *names[6] = { "P", "du", "dv", "duu", "duv", "dvv" };
            std::snprintf(label, sizeof(label), "EvalPatches %s %s (%d coords)", name, names[k], n);
            report(label, maxCondDiff(&x[0], &y[0], n, k, coords, cpu.GetPatchTable(), cpu.GetVertexBuffer()->BindCpuBuffer()), 1e-6);
        }
        // the instantiatable flavour (osd/mesh.h:305-409): an instance from EvaluatorCacheT, the coordinate set bound once
        // (grouped by patch on the device), the same static call with the instance -- bit-identical to the call without
        {
            Osd::EvaluatorCacheT<Osd::B200Evaluator> cache;
            Osd::B200Evaluator *ev = Osd::GetEvaluator<Osd::B200Evaluator>(&cache, src, p, du, dv, duu, duv, dvv, (void *)NULL);
            Osd::B200VertexBuffer *gpuOut2 = Osd::B200VertexBuffer::Create(18, n);
            bool bound = ev && ev->BindPatchCoords(n, gpuCoords, gpu->GetPatchTable());
            bool c = bound && Osd::B200Evaluator::EvalPatches(gpu->GetVertexBuffer(), src, gpuOut2, p, gpuOut2, du, gpuOut2, dv,
                                                              gpuOut2, duu, gpuOut2, duv, gpuOut2, dvv, n, gpuCoords,
                                                              gpu->GetPatchTable(), ev);
            std::vector<float> out2((size_t)n * 18);
            gpuOut2->ReadData(&out2[0], 0, n);
            Osd::B200Evaluator::Synchronize();
            bool same = c && std::memcmp(&out[0], &out2[0], out.size() * sizeof(float)) == 0;
            std::printf("%-58s %s\n", "EvalPatches through an EvaluatorCacheT instance", same ? "bit-identical  ok" : "MISMATCH");
            if (!same) g_fail = 1;
            delete gpuOut2;
        }
        // varying through the same coords (linear patches)
        bool c1 = Osd::CpuEvaluator::EvalPatchesVarying(cpu.GetVertexBuffer(), src, cpuOut, p, n, cpuCoords, cpu.GetPatchTable(),
                                                        (Osd::CpuEvaluator const *)NULL);
        bool c2 = Osd::B200Evaluator::EvalPatchesVarying(gpu->GetVertexBuffer(), src, gpuOut, p, n, gpuCoords, gpu->GetPatchTable(),
                                                         (Osd::B200Evaluator const *)NULL);
        if (c1 != c2) { std::printf("EvalPatchesVarying return values differ\n"); g_fail = 1; }
        gpuOut->ReadData(&out[0], 0, n);
        Osd::B200Evaluator::Synchronize();
        {
            std::vector<float> x((size_t)n * 3), y((size_t)n * 3);
            for (int i = 0; i < n; ++i)
                for (int c = 0; c < 3; ++c) { x[3 * i + c] = out[18 * i + c]; y[3 * i + c] = cpuOut->BindCpuBuffer()[18 * i + c]; }
            std::snprintf(label, sizeof(label), "EvalPatchesVarying %s (%d coords)", name, n);
            report(label, maxRelDiff(&x[0], &y[0], x.size()), 1e-6);
        }
        delete cpuCoords; delete gpuCoords; delete cpuOut; delete gpuOut;
    }
    delete gpu;
    return 0;
}

// ---- glStencilViewer pattern: LimitStencilTable with derivatives ---------------------------------------------
static void limitCase(const char *name, std::string const &str, Scheme scheme) {
    Far::TopologyRefiner *refiner = makeRefiner(str, scheme);
    Far::TopologyRefiner::AdaptiveOptions opt(3);
    refiner->RefineAdaptive(opt);
    Far::PtexIndices ptex(*refiner);
    int nf = ptex.GetNumFaces(), per = 64;
    std::vector<float> u((size_t)nf * per), v((size_t)nf * per);
    std::mt19937 rng(12345);
    std::uniform_real_distribution<float> uni(0.0f, 1.0f);
    Far::LimitStencilTableFactory::LocationArrayVec locs(nf);
    for (int f = 0; f < nf; ++f) {
        locs[f].ptexIdx = f; locs[f].numLocations = per; locs[f].s = &u[(size_t)f * per]; locs[f].t = &v[(size_t)f * per];
        for (int j = 0; j < per; ++j) {
            float a = uni(rng), b = uni(rng);
            if (scheme == kLoop && a + b >= 1.0f) { a = 1.0f - a; b = 1.0f - b; }
            u[(size_t)f * per + j] = a; v[(size_t)f * per + j] = b;
        }
    }
    Far::LimitStencilTableFactory::Options lopt;
    lopt.generate2ndDerivatives = true;
    Far::LimitStencilTable const *lst = Far::LimitStencilTableFactory::Create(*refiner, locs, 0, 0, lopt);
    int n = lst->GetNumStencils(), nCV = lst->GetNumControlVertices();
    std::vector<float> pos = positionsOf(str, scheme);

    Osd::CpuVertexBuffer *csrc = Osd::CpuVertexBuffer::Create(3, nCV), *cout = Osd::CpuVertexBuffer::Create(18, n);
    Osd::B200VertexBuffer *gsrc = Osd::B200VertexBuffer::Create(3, nCV), *gout = Osd::B200VertexBuffer::Create(18, n);
    csrc->UpdateData(&pos[0], 0, nCV);
    gsrc->UpdateData(&pos[0], 0, nCV);
    Osd::B200StencilTable *gtab = Osd::B200StencilTable::Create(lst);
    Osd::BufferDescriptor src(0, 3, 3), p(0, 3, 18), du(3, 3, 18), dv(6, 3, 18), duu(9, 3, 18), duv(12, 3, 18), dvv(15, 3, 18);
    Osd::CpuEvaluator::EvalStencils(csrc, src, cout, p, cout, du, cout, dv, cout, duu, cout, duv, cout, dvv, lst);
    Osd::B200Evaluator::EvalStencils(gsrc, src, gout, p, gout, du, gout, dv, gout, duu, gout, duv, gout, dvv, gtab);
    std::vector<float> out((size_t)n * 18);
    gout->ReadData(&out[0], 0, n);
    Osd::B200Evaluator::Synchronize();
    for (int k = 0; k < 6; ++k) {
        std::vector<float> x((size_t)n * 3), y((size_t)n * 3);
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c) { x[3 * i + c] = out[18 * i + 3 * k + c]; y[3 * i + c] = cout->BindCpuBuffer()[18 * i + 3 * k + c]; }
        static const char *names[6] = { "P", "du", "dv", "duu", "duv", "dvv" };
        char label[128];
        std::snprintf(label, sizeof(label), "EvalStencils(LimitStencilTable) %s %s (%d pts)", name, names[k], n);
        report(label, maxRelDiff(&x[0], &y[0], x.size()), 1e-6);
    }
    // the raw-pointer overload on the table's reference-layout device arrays
    Osd::B200Evaluator::EvalStencils(gsrc->BindCudaBuffer(), src, gout->BindCudaBuffer(), p,
                                     (const int *)gtab->GetSizesBuffer(), (const int *)gtab->GetOffsetsBuffer(),
                                     (const int *)gtab->GetIndicesBuffer(), (const float *)gtab->GetWeightsBuffer(), 0, n);
    gout->ReadData(&out[0], 0, n);
    Osd::B200Evaluator::Synchronize();
    {
        std::vector<float> x((size_t)n * 3), y((size_t)n * 3);
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c) { x[3 * i + c] = out[18 * i + c]; y[3 * i + c] = cout->BindCpuBuffer()[18 * i + c]; }
        report("EvalStencils raw overload (reference-layout arrays)", maxRelDiff(&x[0], &y[0], x.size()), 1e-6);
    }
    // a table that already lives on the device in the reference layout (what an Osd::CudaStencilTable is) converted once
    {
        Osd::B200StencilTable *conv = Osd::B200StencilTable::CreateFromDevice(gtab, nCV);
        Osd::B200VertexBuffer *gout2 = Osd::B200VertexBuffer::Create(18, n);
        Osd::B200Evaluator::EvalStencils(gsrc, src, gout, p, gout, du, gout, dv, gout, duu, gout, duv, gout, dvv, gtab);
        Osd::B200Evaluator::EvalStencils(gsrc, src, gout2, p, gout2, du, gout2, dv, gout2, duu, gout2, duv, gout2, dvv, conv);
        std::vector<float> out2((size_t)n * 18);
        gout->ReadData(&out[0], 0, n);
        gout2->ReadData(&out2[0], 0, n);
        Osd::B200Evaluator::Synchronize();
        report("EvalStencils through CreateFromDevice(device arrays) vs Create(Far table), bitwise",
               conv && std::memcmp(&out[0], &out2[0], out.size() * sizeof(float)) == 0 ? 0.0 : 1.0, 0.0);
        delete conv; delete gout2;
    }
    delete gtab; delete csrc; delete cout; delete gsrc; delete gout; delete lst; delete refiner;
}

int main() {
    std::printf("%s\n", b200osd_version());
    int rc = meshCase("catmark_cube_creases0", catmark_cube_creases0, kCatmark, 4, false);
    if (rc == 2) return 2;
    meshCase("catmark_cube_creases0", catmark_cube_creases0, kCatmark, 3, true);
    meshCase("catmark_pyramid", catmark_pyramid, kCatmark, 3, true);
    meshCase("loop_icosahedron", loop_icosahedron, kLoop, 3, false);
    meshCase("loop_icosahedron", loop_icosahedron, kLoop, 3, true);
    limitCase("catmark_pyramid", catmark_pyramid, kCatmark);
    limitCase("loop_icosahedron", loop_icosahedron, kLoop);
    std::printf(g_fail ? "DROP-IN TEST FAILED\n" : "DROP-IN TEST PASSED\n");
    return g_fail;
}
