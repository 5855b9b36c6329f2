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

#include <algorithm>
#include <cstring>
#include <new>

using namespace b200osd;

struct b200osd_frame {
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;          // second branch of the frame (e.g. the control-point broadcast of the NEXT frame)
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool recording = false;
};

extern "C" {

b200osd_frame *b200osd_frame_create(void) {
    b200osd_frame *f = new (std::nothrow) b200osd_frame;
    if (!f) return nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
    // the side branch gets the highest priority: its work (a broadcast of a few MB) must not queue behind the
    // tens of thousands of blocks of an evaluation kernel on the main branch
    int lo = 0, hi = 0;
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&f->side, cudaStreamNonBlocking, hi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->join, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        set_error("frame_create: stream / event creation failed: %s", cudaGetErrorString(e));
        b200osd_frame_destroy(f);
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
    if (f->fork) cudaEventDestroy(f->fork);
    if (f->join) cudaEventDestroy(f->join);
    if (f->side) cudaStreamDestroy(f->side);
    if (f->stream) cudaStreamDestroy(f->stream);
    delete f;
}

void *b200osd_frame_stream(const b200osd_frame *f) { return f ? (void *)f->stream : nullptr; }
void *b200osd_frame_side_stream(const b200osd_frame *f) { return f ? (void *)f->side : nullptr; }

int b200osd_frame_fence(b200osd_frame *f, int mainWaitsForSide) {
    if (!f) { set_error("frame_fence: frame is NULL"); return B200OSD_ERR_INVALID; }
    cudaStream_t from = mainWaitsForSide ? f->side : f->stream, to = mainWaitsForSide ? f->stream : f->side;
    // a fresh event per edge would be needed outside capture only if fences overlapped; inside a capture every record /
    // wait pair becomes a graph dependency at once, so the two events can be reused
    cudaEvent_t ev = mainWaitsForSide ? f->join : f->fork;
    B200_CUDA_TRY(cudaEventRecord(ev, from));
    B200_CUDA_TRY(cudaStreamWaitEvent(to, ev, 0));
    return B200OSD_OK;
}

int b200osd_frame_set_l2_window(b200osd_frame *f, const void *devPtr, size_t bytes, float hitRatio) {
    if (!f) { set_error("frame_set_l2_window: frame is NULL"); return B200OSD_ERR_INVALID; }
    int dev = 0, maxWindow = 0, maxPersist = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    B200_CUDA_TRY(cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    B200_CUDA_TRY(cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, dev));
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof(attr));
    if (devPtr && bytes > 0 && maxWindow > 0 && maxPersist > 0) {
        const size_t window = std::min(bytes, (size_t)maxWindow);
        // set aside as much L2 as the window needs (the device caps it); lines of the window are then kept across the
        // kernels of the frame instead of being evicted by the streaming tables and results
        B200_CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(window, (size_t)maxPersist)));
        attr.accessPolicyWindow.base_ptr = const_cast<void *>(devPtr);
        attr.accessPolicyWindow.num_bytes = window;
        attr.accessPolicyWindow.hitRatio = hitRatio > 0.0f && hitRatio <= 1.0f ? hitRatio : 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        attr.accessPolicyWindow.num_bytes = 0;                  // clears the window
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    B200_CUDA_TRY(cudaStreamSetAttribute(f->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    B200_CUDA_TRY(cudaStreamSetAttribute(f->side, cudaStreamAttributeAccessPolicyWindow, &attr));
    return B200OSD_OK;
}

int b200osd_frame_begin(b200osd_frame *f) {
    if (!f) { set_error("frame_begin: frame is NULL"); return B200OSD_ERR_INVALID; }
    if (f->recording) { set_error("frame_begin: already recording"); return B200OSD_ERR_INVALID; }
    if (f->exec) { cudaGraphExecDestroy(f->exec); f->exec = nullptr; }
    if (f->graph) { cudaGraphDestroy(f->graph); f->graph = nullptr; }
    // thread-local mode: other threads of the application keep full use of the CUDA API while this one records
    B200_CUDA_TRY(cudaStreamBeginCapture(f->stream, cudaStreamCaptureModeThreadLocal));
    f->recording = true;
    // the side stream joins the capture: work issued on it is a parallel branch of the same graph
    cudaError_t e = cudaEventRecord(f->fork, f->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(f->side, f->fork, 0);
    if (e != cudaSuccess) {
        set_error("frame_begin: forking the side stream failed: %s", cudaGetErrorString(e));
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(f->stream, &g);
        if (g) cudaGraphDestroy(g);
        f->recording = false;
        return B200OSD_ERR_CUDA;
    }
    return B200OSD_OK;
}

int b200osd_frame_end(b200osd_frame *f) {
    if (!f || !f->recording) { set_error("frame_end: not recording"); return B200OSD_ERR_INVALID; }
    f->recording = false;
    // every branch must rejoin the origin stream before the capture can end
    cudaError_t e = cudaEventRecord(f->join, f->side);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(f->stream, f->join, 0);
    if (e != cudaSuccess) cudaGetLastError();
    e = cudaStreamEndCapture(f->stream, &f->graph);
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
