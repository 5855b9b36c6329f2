// b200PatchMap.h -- device-resident counterpart of Far::PatchMap (opensubdiv/far/patchMap.h:48-217).
//
// The reference locates one sample per call on the host -- handle = patchMap.FindPatch(ptexFace, s, t), then
// Osd::PatchCoord(*handle, s, t) (examples/glEvalLimit/particles.cpp:91-115,392-394) -- and uploads the records
// every frame.  B200PatchMap answers a whole batch with one kernel launch and writes the Osd::PatchCoord records
// straight into a device buffer that EvalPatches* can consume.  Results are bit-identical to Far::PatchMap.
#ifndef B200OSD_PATCH_MAP_H
#define B200OSD_PATCH_MAP_H

#include <opensubdiv/version.h>
#include <opensubdiv/far/patchTable.h>
#include <opensubdiv/osd/cpuPatchTable.h>
#include <opensubdiv/osd/types.h>

#include "../b200osd_capi.h"
#include "b200VertexBuffer.h"     // B200DeviceContext

namespace OpenSubdiv {
namespace OPENSUBDIV_VERSION {
namespace Osd {

class B200PatchMap {
public:
    /// One {ptexFace, s, t} sample, 12 bytes; what glEvalLimit keeps per particle (particles.h: Position).
    struct Sample { int ptexFace; float s, t; };

    static B200PatchMap *Create(Far::PatchTable const *farPatchTable) {
        if (!farPatchTable) return NULL;
        CpuPatchTable cpu(farPatchTable);
        bool tri = farPatchTable->GetVaryingPatchDescriptor().GetNumControlVertices() == 3;    // far/patchMap.cpp:93-94
        b200osd_patch_map *h = b200osd_patch_map_create(
            (int)cpu.GetNumPatchArrays(), reinterpret_cast<b200osd_patch_array const *>(cpu.GetPatchArrayBuffer()),
            (int)cpu.GetPatchParamSize(), reinterpret_cast<b200osd_patch_param const *>(cpu.GetPatchParamBuffer()), tri ? 1 : 0);
        return h ? new B200PatchMap(h) : NULL;
    }
    ~B200PatchMap() { b200osd_patch_map_destroy(_h); }

    /// samples: any buffer with BindCudaBuffer() holding numSamples Sample records (e.g. a 3-element B200VertexBuffer);
    /// patchCoords: any buffer with BindCudaBuffer() with room for numSamples Osd::PatchCoord records (a 5-element
    /// B200VertexBuffer, like glEvalLimit.cpp:473-486).  A sample outside every patch (FindPatch == NULL) gets
    /// handle.arrayIndex = -1 and B200Evaluator leaves its outputs untouched.  numFound: optional DEVICE int.
    template <typename SAMPLE_BUFFER, typename COORD_BUFFER>
    bool FindPatches(int numSamples, SAMPLE_BUFFER *samples, COORD_BUFFER *patchCoords, int *numFound = NULL,
                     void *deviceContext = NULL) const {
        Sample const *sp = reinterpret_cast<Sample const *>(samples->BindCudaBuffer());
        return FindPatches(numSamples, sp ? &sp->ptexFace : NULL, 3, sp ? &sp->s : NULL, 3, sp ? &sp->t : NULL, 3,
                           reinterpret_cast<PatchCoord *>(patchCoords->BindCudaBuffer()), numFound, deviceContext);
    }

    /// Raw DEVICE pointers with element strides (1,1,1 for three packed arrays).
    bool FindPatches(int numSamples, const int *ptexFace, int faceStride, const float *s, int sStride, const float *t,
                     int tStride, PatchCoord *patchCoords, int *numFound = NULL, void *deviceContext = NULL) const {
        void *stream = deviceContext ? (void *)static_cast<B200DeviceContext *>(deviceContext)->stream : NULL;
        return b200osd_patch_map_find(_h, numSamples, ptexFace, faceStride, s, sStride, t, tStride,
                                      reinterpret_cast<b200osd_patch_coord *>(patchCoords), numFound, stream) == B200OSD_OK;
    }

    int GetMinPatchFace() const { return info(0); }
    int GetMaxPatchFace() const { return info(1); }
    int GetMaxDepth() const { return info(2); }
    b200osd_patch_map const *GetHandle() const { return _h; }

private:
    int info(int k) const { int v[6] = {0, 0, 0, 0, 0, 0}; b200osd_patch_map_info(_h, v); return v[k]; }
    explicit B200PatchMap(b200osd_patch_map *h) : _h(h) {}
    B200PatchMap(B200PatchMap const &);
    B200PatchMap &operator=(B200PatchMap const &);
    b200osd_patch_map *_h;
};

}  // namespace Osd
}  // namespace OPENSUBDIV_VERSION
using namespace OPENSUBDIV_VERSION;
}  // namespace OpenSubdiv

#endif
