// b200VertexBuffer.h -- drop-in for Osd::CudaVertexBuffer (opensubdiv/osd/cudaVertexBuffer.h:42-80), header-only
// over the C ABI (include/b200osd_capi.h).  Same member names, argument meaning and failure behaviour
// (Create() returns NULL when the device allocation fails, cudaVertexBuffer.cpp:46-53).
#ifndef B200OSD_VERTEX_BUFFER_H
#define B200OSD_VERTEX_BUFFER_H

#include <opensubdiv/version.h>
#include <cstddef>

#include "../b200osd_capi.h"

namespace OpenSubdiv {
namespace OPENSUBDIV_VERSION {
namespace Osd {

/// Optional device context for every B200 class: the opaque `void *deviceContext` argument of the Osd
/// templates may point at one of these to select a CUDA stream (NULL = legacy default stream, as the reference).
struct B200DeviceContext {
    void *stream;   // cudaStream_t
    B200DeviceContext(void *s = NULL) : stream(s) {}
};

inline void *B200StreamOf(void *deviceContext) {
    return deviceContext ? static_cast<B200DeviceContext *>(deviceContext)->stream : NULL;
}

class B200VertexBuffer {
public:
    static B200VertexBuffer *Create(int numElements, int numVertices, void *deviceContext = NULL) {
        (void)deviceContext;
        b200osd_vertex_buffer *h = b200osd_vertex_buffer_create(numElements, numVertices);
        return h ? new B200VertexBuffer(h) : NULL;
    }
    ~B200VertexBuffer() { b200osd_vertex_buffer_destroy(_h); }

    /// Host -> device copy of `numVertices` vertices starting at `startVertex` (cudaVertexBuffer.cpp:56-64).
    void UpdateData(const float *src, int startVertex, int numVertices, void *deviceContext = NULL) {
        b200osd_vertex_buffer_update(_h, src, startVertex, numVertices, B200StreamOf(deviceContext));
    }
    /// Device -> host read-back (what clients do after Synchronize()); not part of the reference class.
    void ReadData(float *dst, int startVertex, int numVertices, void *deviceContext = NULL) {
        b200osd_vertex_buffer_read(_h, dst, startVertex, numVertices, B200StreamOf(deviceContext));
    }
    int GetNumElements() const { return b200osd_vertex_buffer_num_elements(_h); }
    int GetNumVertices() const { return b200osd_vertex_buffer_num_vertices(_h); }

    /// Device pointer; the evaluator templates call exactly this name (osd/cudaEvaluator.h:135-136).
    float *BindCudaBuffer() { return b200osd_vertex_buffer_bind(_h); }
    /// Osd::Mesh calls BindVBO from its virtual accessors (osd/mesh.h:562-568); headless: the device pointer.
    float *BindVBO(void *deviceContext = NULL) { (void)deviceContext; return BindCudaBuffer(); }

private:
    explicit B200VertexBuffer(b200osd_vertex_buffer *h) : _h(h) {}
    B200VertexBuffer(B200VertexBuffer const &);
    B200VertexBuffer &operator=(B200VertexBuffer const &);
    b200osd_vertex_buffer *_h;
};

}  // namespace Osd
}  // namespace OPENSUBDIV_VERSION
using namespace OPENSUBDIV_VERSION;
}  // namespace OpenSubdiv

#endif
