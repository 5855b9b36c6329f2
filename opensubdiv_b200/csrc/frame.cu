// frame.cu -- a whole evaluation frame as ONE replayable launch.
//
// Every compute entry point of this library is stream-ordered and, after its first call, allocation-free.  A frame
// -- e.g. EvalStencils (refinement + local points) -> FindPatches -> EvalPatches* -- issued on the frame's stream
// between b200osd_frame_begin and b200osd_frame_end is therefore recorded by CUDA stream capture into a graph and
// replayed with b200osd_frame_launch: one driver call per frame instead of one per kernel, and no host work between
// the kernels (the refined points of a mid-size mesh stay L2-resident between refinement and patch evaluation).
// The buffers are the caller's: updating the control points in place and relaunching evaluates the new frame.
// The reference has no counterpart (osd/cudaEvaluator.cpp launches on the legacy default stream, one call per kernel).
#include "common.cuh"

#include <new>

using namespace b200osd;

struct b200osd_frame {
    cudaStream_t stream = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool recording = false;
};

extern "C" {

b200osd_frame *b200osd_frame_create(void) {
    b200osd_frame *f = new (std::nothrow) b200osd_frame;
    if (!f) return nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_error("frame_create: cudaStreamCreate failed: %s", cudaGetErrorString(e));
        delete f;
        return nullptr;
    }
    return f;
}

void b200osd_frame_destroy(b200osd_frame *f) {
    if (!f) return;
    if (f->recording) {                       // abandon an unfinished recording
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(f->stream, &g);
        if (g) cudaGraphDestroy(g);
    }
    if (f->exec) cudaGraphExecDestroy(f->exec);
    if (f->graph) cudaGraphDestroy(f->graph);
    if (f->stream) cudaStreamDestroy(f->stream);
    delete f;
}

void *b200osd_frame_stream(const b200osd_frame *f) { return f ? (void *)f->stream : nullptr; }

int b200osd_frame_begin(b200osd_frame *f) {
    if (!f) { set_error("frame_begin: frame is NULL"); return B200OSD_ERR_INVALID; }
    if (f->recording) { set_error("frame_begin: already recording"); return B200OSD_ERR_INVALID; }
    if (f->exec) { cudaGraphExecDestroy(f->exec); f->exec = nullptr; }
    if (f->graph) { cudaGraphDestroy(f->graph); f->graph = nullptr; }
    // thread-local mode: other threads of the application keep full use of the CUDA API while this one records
    B200_CUDA_TRY(cudaStreamBeginCapture(f->stream, cudaStreamCaptureModeThreadLocal));
    f->recording = true;
    return B200OSD_OK;
}

int b200osd_frame_end(b200osd_frame *f) {
    if (!f || !f->recording) { set_error("frame_end: not recording"); return B200OSD_ERR_INVALID; }
    f->recording = false;
    cudaError_t e = cudaStreamEndCapture(f->stream, &f->graph);
    if (e != cudaSuccess || !f->graph) {
        set_error("frame_end: capture failed: %s (run the frame once before recording it: first calls allocate)",
                  cudaGetErrorString(e));
        cudaGetLastError();
        f->graph = nullptr;
        return B200OSD_ERR_CUDA;
    }
    e = cudaGraphInstantiate(&f->exec, f->graph, 0);
    if (e != cudaSuccess) {
        set_error("frame_end: cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
        cudaGraphDestroy(f->graph);
        f->graph = nullptr;
        f->exec = nullptr;
        return B200OSD_ERR_CUDA;
    }
    return B200OSD_OK;
}

int b200osd_frame_launch(b200osd_frame *f) {
    if (!f || !f->exec) { set_error("frame_launch: nothing recorded"); return B200OSD_ERR_INVALID; }
    B200_CUDA_TRY(cudaGraphLaunch(f->exec, f->stream));
    return B200OSD_OK;
}

int b200osd_frame_synchronize(b200osd_frame *f) {
    if (!f) { set_error("frame_synchronize: frame is NULL"); return B200OSD_ERR_INVALID; }
    B200_CUDA_TRY(cudaStreamSynchronize(f->stream));
    return B200OSD_OK;
}

}  // extern "C"
