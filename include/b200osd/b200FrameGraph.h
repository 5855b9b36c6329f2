// b200FrameGraph.h -- a whole evaluation frame recorded once and replayed as ONE launch (no reference counterpart;
// Osd::CudaEvaluator issues one launch per kernel on the legacy default stream, osd/cudaEvaluator.cpp:150-372).
//
//   Osd::B200FrameGraph frame;                                   // owns a stream
//   runFrame(frame.GetDeviceContext());  frame.Synchronize();    // once eagerly: first calls allocate
//   frame.Begin();  runFrame(frame.GetDeviceContext());  frame.End();
//   for (;;) { updateControlPointsInPlace();  frame.Launch(); }
//
// where runFrame issues Osd::Mesh::Refine / B200Evaluator::EvalStencils / B200PatchMap::FindPatches /
// B200Evaluator::EvalPatches* with the given deviceContext.
#ifndef B200OSD_FRAME_GRAPH_H
#define B200OSD_FRAME_GRAPH_H

#include <opensubdiv/version.h>

#include "../b200osd_capi.h"
#include "b200VertexBuffer.h"     // B200DeviceContext

namespace OpenSubdiv {
namespace OPENSUBDIV_VERSION {
namespace Osd {

class B200FrameGraph {
public:
    B200FrameGraph() : _h(b200osd_frame_create()), _ctx(_h ? b200osd_frame_stream(_h) : NULL) {}
    ~B200FrameGraph() { b200osd_frame_destroy(_h); }
    bool IsValid() const { return _h != NULL; }

    /// The `void *deviceContext` to hand to every B200 class call that belongs to the frame.
    void *GetDeviceContext() { return &_ctx; }

    bool Begin() { return b200osd_frame_begin(_h) == B200OSD_OK; }
    bool End() { return b200osd_frame_end(_h) == B200OSD_OK; }
    bool Launch() { return b200osd_frame_launch(_h) == B200OSD_OK; }
    bool Synchronize() { return b200osd_frame_synchronize(_h) == B200OSD_OK; }

private:
    B200FrameGraph(B200FrameGraph const &);
    B200FrameGraph &operator=(B200FrameGraph const &);
    b200osd_frame *_h;
    B200DeviceContext _ctx;
};

}  // namespace Osd
}  // namespace OPENSUBDIV_VERSION
using namespace OPENSUBDIV_VERSION;
}  // namespace OpenSubdiv

#endif
