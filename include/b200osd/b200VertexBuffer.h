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
/// The struct is TAGGED: Osd::CudaEvaluator ignores deviceContext altogether, so an existing client may hand over
/// a pointer to something else (e.g. through Osd::Mesh's DEVICE_CONTEXT); anything whose first word is not the tag is
/// treated like NULL instead of being read as a stream.
struct B200DeviceContext {
    static const unsigned kTag = 0xB2000D5Cu;
    unsigned tag;
    void *stream;   // cudaStream_t
    B200DeviceContext(void *s = NULL) : tag(kTag), stream(s) {}
};

inline void *B200StreamOf(void *deviceContext) {
    if (!deviceContext) return NULL;
    B200DeviceContext const *c = static_cast<B200DeviceContext const *>(deviceContext);
    return c->tag == B200DeviceContext::kTag ? c->stream : NULL;
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
