// b200StencilTable.h -- drop-in for Osd::CudaStencilTable (opensubdiv/osd/cudaEvaluator.h:52-92): a deep device copy
// of a Far::StencilTable / Far::LimitStencilTable (Osd::Mesh deletes the Far tables after conversion, osd/mesh.h:676-678).
// Besides the verbatim reference-layout arrays (the Get*Buffer() accessors) the table owns the B200 bucketed layout
// used by B200Evaluator's fast path.
#ifndef B200OSD_STENCIL_TABLE_H
#define B200OSD_STENCIL_TABLE_H

#include <opensubdiv/version.h>
#include <opensubdiv/far/stencilTable.h>

#include "../b200osd_capi.h"

namespace OpenSubdiv {
namespace OPENSUBDIV_VERSION {
namespace Osd {

class B200StencilTable {
public:
    static B200StencilTable *Create(Far::StencilTable const *stencilTable, void *deviceContext = NULL) {
        (void)deviceContext;
        if (!stencilTable) return NULL;
        return wrap(make(stencilTable->GetNumStencils(), stencilTable->GetNumControlVertices(), stencilTable->GetSizes(), stencilTable->GetOffsets(),
                         stencilTable->GetControlIndices(), stencilTable->GetWeights(), NULL, NULL, NULL, NULL, NULL));
    }
    static B200StencilTable *Create(Far::LimitStencilTable const *t, void *deviceContext = NULL) {
        (void)deviceContext;
        if (!t) return NULL;
        return wrap(make(t->GetNumStencils(), t->GetNumControlVertices(), t->GetSizes(), t->GetOffsets(), t->GetControlIndices(), t->GetWeights(),
                         &t->GetDuWeights(), &t->GetDvWeights(), &t->GetDuuWeights(), &t->GetDuvWeights(), &t->GetDvvWeights()));
    }
    /// From a table that already lives on the device in the reference layout, e.g. an Osd::CudaStencilTable
    /// (osd/cudaEvaluator.h:57-90): anything with GetSizesBuffer() ... GetWeightsBuffer() [, GetDuWeightsBuffer() ...]
    /// returning device pointers.  One conversion; afterwards EvalStencils runs on the bucketed layout.
    template <typename DEVICE_STENCIL_TABLE>
    static B200StencilTable *CreateFromDevice(DEVICE_STENCIL_TABLE const *t, int numControlVertices = 0) {
        if (!t) return NULL;
        return wrap(b200osd_stencil_table_create_from_device(
            t->GetNumStencils(), numControlVertices, (const int *)t->GetSizesBuffer(), (const int *)t->GetOffsetsBuffer(),
            (const int *)t->GetIndicesBuffer(), (const float *)t->GetWeightsBuffer(), (const float *)t->GetDuWeightsBuffer(),
            (const float *)t->GetDvWeightsBuffer(), (const float *)t->GetDuuWeightsBuffer(),
            (const float *)t->GetDuvWeightsBuffer(), (const float *)t->GetDvvWeightsBuffer(), 0));
    }
    ~B200StencilTable() { b200osd_stencil_table_destroy(_h); }

    // interfaces needed by the evaluator templates (device pointers, reference layout)
    void *GetSizesBuffer() const { return buf(0); }
    void *GetOffsetsBuffer() const { return buf(1); }
    void *GetIndicesBuffer() const { return buf(2); }
    void *GetWeightsBuffer() const { return buf(3); }
    void *GetDuWeightsBuffer() const { return buf(4); }
    void *GetDvWeightsBuffer() const { return buf(5); }
    void *GetDuuWeightsBuffer() const { return buf(6); }
    void *GetDuvWeightsBuffer() const { return buf(7); }
    void *GetDvvWeightsBuffer() const { return buf(8); }
    int GetNumStencils() const { return b200osd_stencil_table_num_stencils(_h); }

    /// The C-ABI handle (bucketed layout) for B200Evaluator's fast path.
    b200osd_stencil_table const *GetHandle() const { return _h; }

private:
    static b200osd_stencil_table *make(int n, int numControlVertices, std::vector<int> const &sizes, std::vector<Far::Index> const &offsets,
                                       std::vector<Far::Index> const &indices, std::vector<float> const &weights,
                                       std::vector<float> const *du, std::vector<float> const *dv,
                                       std::vector<float> const *duu, std::vector<float> const *duv,
                                       std::vector<float> const *dvv) {
        // the real control-vertex count lets the library recognise unfactorized tables (rows referencing earlier rows)
        return b200osd_stencil_table_create(n, numControlVertices, data(sizes), data(offsets), data(indices), data(weights),
                                            du ? data(*du) : NULL, dv ? data(*dv) : NULL, duu ? data(*duu) : NULL,
                                            duv ? data(*duv) : NULL, dvv ? data(*dvv) : NULL, 0);
    }
    template <typename T> static T const *data(std::vector<T> const &v) { return v.empty() ? NULL : &v[0]; }
    static B200StencilTable *wrap(b200osd_stencil_table *h) { return h ? new B200StencilTable(h) : NULL; }
    void *buf(int which) const { return const_cast<void *>(b200osd_stencil_table_buffer(_h, which)); }
    explicit B200StencilTable(b200osd_stencil_table *h) : _h(h) {}
    B200StencilTable(B200StencilTable const &);
    B200StencilTable &operator=(B200StencilTable const &);
    b200osd_stencil_table *_h;
};

}  // namespace Osd
}  // namespace OPENSUBDIV_VERSION
using namespace OPENSUBDIV_VERSION;
}  // namespace OpenSubdiv

#endif
