// b200PatchTable.h -- drop-in for Osd::CudaPatchTable (opensubdiv/osd/cudaPatchTable.h:51-112).  Like the reference
// (osd/cudaPatchTable.cpp:69-71) it stages through Osd::CpuPatchTable, which flattens the Far::PatchTable into
// (PatchArray[], index buffer, PatchParam[]) triples for vertex, varying and each face-varying channel, then uploads.
#ifndef B200OSD_PATCH_TABLE_H
#define B200OSD_PATCH_TABLE_H

#include <opensubdiv/version.h>
#include <opensubdiv/far/patchTable.h>
#include <opensubdiv/osd/cpuPatchTable.h>
#include <opensubdiv/osd/types.h>

#include "../b200osd_capi.h"

namespace OpenSubdiv {
namespace OPENSUBDIV_VERSION {
namespace Osd {

class B200PatchTable {
public:
    /// Osd::Mesh requires PatchTable::VertexBufferBinding (osd/mesh.h:71,426); headless: a device pointer.
    typedef float *VertexBufferBinding;

    static B200PatchTable *Create(Far::PatchTable const *farPatchTable, void *deviceContext = NULL) {
        (void)deviceContext;
        if (!farPatchTable) return NULL;
        CpuPatchTable cpu(farPatchTable);
        int nfv = cpu.GetNumFVarChannels();
        b200osd_patch_table *h = b200osd_patch_table_create(nfv);
        if (!h) return NULL;
        bool ok = b200osd_patch_table_set(h, 0, (int)cpu.GetNumPatchArrays(), arrays(cpu.GetPatchArrayBuffer()),
                                          (int)cpu.GetPatchIndexSize(), cpu.GetPatchIndexBuffer(),
                                          (int)cpu.GetPatchParamSize(), params(cpu.GetPatchParamBuffer())) == B200OSD_OK;
        if (ok && cpu.GetVaryingPatchArrayBuffer())
            ok = b200osd_patch_table_set(h, 1, (int)cpu.GetNumPatchArrays(), arrays(cpu.GetVaryingPatchArrayBuffer()),
                                         (int)cpu.GetVaryingPatchIndexSize(), cpu.GetVaryingPatchIndexBuffer(), 0, NULL) == B200OSD_OK;
        for (int c = 0; ok && c < nfv; ++c)
            ok = b200osd_patch_table_set(h, 2 + c, (int)cpu.GetNumPatchArrays(), arrays(cpu.GetFVarPatchArrayBuffer(c)),
                                         (int)cpu.GetFVarPatchIndexSize(c), cpu.GetFVarPatchIndexBuffer(c),
                                         (int)cpu.GetFVarPatchParamSize(c), params(cpu.GetFVarPatchParamBuffer(c))) == B200OSD_OK;
#ifdef OPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES
        // the reference's build-time switch (CMakeLists.txt:340,640; osd/patchBasis.h:421-487) carried over per table
        if (ok) ok = b200osd_patch_table_set_options(h, B200OSD_PATCH_GREGORY_TRUE_DERIVATIVES) == B200OSD_OK;
#endif
        if (!ok) { b200osd_patch_table_destroy(h); return NULL; }      // cudaPatchTable.cpp:59-66
        return new B200PatchTable(h);
    }
    ~B200PatchTable() { b200osd_patch_table_destroy(_h); }

    void *GetPatchArrayBuffer() const { return buf(0, 0); }
    void *GetPatchIndexBuffer() const { return buf(0, 1); }
    void *GetPatchParamBuffer() const { return buf(0, 2); }
    void *GetVaryingPatchArrayBuffer() const { return buf(1, 0); }
    void *GetVaryingPatchIndexBuffer() const { return buf(1, 1); }
    int GetNumFVarChannels() const { return b200osd_patch_table_num_fvar_channels(_h); }
    void *GetFVarPatchArrayBuffer(int fvarChannel) const { return buf(2 + fvarChannel, 0); }
    void *GetFVarPatchIndexBuffer(int fvarChannel = 0) const { return buf(2 + fvarChannel, 1); }
    void *GetFVarPatchParamBuffer(int fvarChannel = 0) const { return buf(2 + fvarChannel, 2); }

    /// The C-ABI handle for B200Evaluator's fast path (immutable after Create: shareable across threads and streams).
    b200osd_patch_table const *GetHandle() const { return _h; }

private:
    static_assert(sizeof(PatchArray) == sizeof(b200osd_patch_array), "Osd::PatchArray layout");
    static_assert(sizeof(PatchParam) == sizeof(b200osd_patch_param), "Osd::PatchParam layout");
    static_assert(sizeof(PatchCoord) == sizeof(b200osd_patch_coord), "Osd::PatchCoord layout");
    static b200osd_patch_array const *arrays(PatchArray const *p) { return reinterpret_cast<b200osd_patch_array const *>(p); }
    static b200osd_patch_param const *params(PatchParam const *p) { return reinterpret_cast<b200osd_patch_param const *>(p); }
    void *buf(int which, int kind) const { return const_cast<void *>(b200osd_patch_table_buffer(_h, which, kind)); }
    explicit B200PatchTable(b200osd_patch_table *h) : _h(h) {}
    B200PatchTable(B200PatchTable const &);
    B200PatchTable &operator=(B200PatchTable const &);
    b200osd_patch_table *_h;
};

}  // namespace Osd
}  // namespace OPENSUBDIV_VERSION
using namespace OPENSUBDIV_VERSION;
}  // namespace OpenSubdiv

#endif
