/*
 * b200osd_capi.h -- the C ABI of the B200-native Osd evaluator (libb200osd.so).
 *
 * This is the ONLY boundary between host code (C++ wrapper classes in include/b200osd/, the Python
 * mirror in opensubdiv_b200/, or a third-party binding) and the hand-written sm_100a kernels.  It
 * plays the role the three `extern "C"` launchers play in the reference CUDA backend:
 *
 *   reference (OpenSubdiv 3.6.0, paths relative to opensubdiv/)            replaced by
 *   ---------------------------------------------------------------------  -----------------------------------
 *   CudaEvalStencils               osd/cudaEvaluator.cpp:33-44,            b200osd_eval_stencils
 *                                  osd/cudaKernel.cu:351-383               b200osd_stencil_table_eval (fast path)
 *   CudaEvalPatches                osd/cudaEvaluator.cpp:46-53,            b200osd_eval_patches  (nOut = 1)
 *                                  osd/cudaKernel.cu:387-400
 *   CudaEvalPatchesWithDerivatives osd/cudaEvaluator.cpp:55-67,            b200osd_eval_patches  (nOut = 3 or 6)
 *                                  osd/cudaKernel.cu:402-424
 *   CudaStencilTable ctor/dtor     osd/cudaEvaluator.cpp:100-145           b200osd_stencil_table_create/_destroy
 *   CudaPatchTable::allocate       osd/cudaPatchTable.cpp:69-162           b200osd_patch_table_create/_set/_destroy
 *   CudaVertexBuffer               osd/cudaVertexBuffer.cpp:35-93          b200osd_vertex_buffer_*
 *   CudaEvaluator::Synchronize     osd/cudaEvaluator.cpp:377-380           b200osd_synchronize
 *
 * Conventions
 *  - Plain C: pointers, ints, no C++ or torch types.  All `float*`/`int*` data arguments of the
 *    eval functions are DEVICE pointers (anything cudaMalloc-compatible, e.g. a torch tensor's
 *    data_ptr()); `*_create`, `*_update` and `*_read` take HOST pointers and copy.
 *  - Descriptors are int[3] = {offset, length, stride} in floats (Osd::BufferDescriptor,
 *    osd/bufferDescriptor.h:61-104).  Unlike the reference C launchers the descriptor offset is
 *    applied INSIDE this library (pass the un-offset base pointer).
 *  - nOut is 1 (value), 3 (value, du, dv) or 6 (value, du, dv, duu, duv, dvv).  A NULL entry of
 *    dsts[] skips that output (osd/cudaKernel.cu:300-327); dsts[0] == NULL with nOut == 1 is an error.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, as the reference uses).
 *    Launches are asynchronous; call b200osd_synchronize before reading results on the host.
 *  - Return value: B200OSD_OK, or an error code; B200OSD_ERR_INVALID corresponds to the reference
 *    evaluators returning `false` (length mismatch / NULL src or dst, osd/cpuEvaluator.cpp:46-47,
 *    :165-176).  b200osd_last_error() gives a thread-local message.
 *  - Stencil row addressing is absolute: row i of [start,end) is written to element i of each dst
 *    (the CUDA backend's convention, osd/cudaKernel.cu:85-98); end <= start is a successful no-op.
 *  - There is NO CPU fallback: every entry point that computes fails with B200OSD_ERR_CUDA when no
 *    usable sm_100 device is present.
 */
#ifndef B200OSD_CAPI_H
#define B200OSD_CAPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define B200OSD_API
#else
#define B200OSD_API __attribute__((visibility("default")))
#endif

enum {
    B200OSD_OK = 0,
    B200OSD_ERR_INVALID = 1,   /* the reference evaluator would return false */
    B200OSD_ERR_CUDA = 2,      /* CUDA runtime / launch failure, or no device */
    B200OSD_ERR_ALLOC = 3,     /* host or device allocation failed (reference: Create() returns NULL) */
    B200OSD_ERR_UNSUPPORTED = 4
};

/* POD mirrors of the Osd value types (layouts verified by static_assert in the C++ wrappers). */
typedef struct b200osd_patch_coord { int arrayIndex, patchIndex, vertIndex; float s, t; } b200osd_patch_coord;      /* osd/types.h:42-64   */
typedef struct b200osd_patch_array { int regDesc, desc, numPatches, indexBase, stride, primitiveIdBase; } b200osd_patch_array; /* osd/types.h:66-122 */
typedef struct b200osd_patch_param { unsigned int field0, field1; float sharpness; } b200osd_patch_param;           /* osd/types.h:127-130 */

typedef struct b200osd_stencil_table b200osd_stencil_table;
typedef struct b200osd_patch_table   b200osd_patch_table;
typedef struct b200osd_patch_map     b200osd_patch_map;
typedef struct b200osd_vertex_buffer b200osd_vertex_buffer;
typedef struct b200osd_frame         b200osd_frame;

/* ---- library -------------------------------------------------------------------------------- */
B200OSD_API const char *b200osd_version(void);
B200OSD_API const char *b200osd_last_error(void);
/* Number of kernels this library has launched since load (or the last reset) -- bench.py's gpu_launches. */
B200OSD_API long long   b200osd_launch_count(void);
B200OSD_API void        b200osd_reset_launch_count(void);
B200OSD_API int         b200osd_synchronize(void *stream);   /* NULL: cudaDeviceSynchronize (CudaEvaluator::Synchronize) */

/* ---- vertex buffer (CudaVertexBuffer, osd/cudaVertexBuffer.h:42-80) -------------------------- */
B200OSD_API b200osd_vertex_buffer *b200osd_vertex_buffer_create(int numElements, int numVertices);
B200OSD_API void   b200osd_vertex_buffer_destroy(b200osd_vertex_buffer *vb);
B200OSD_API int    b200osd_vertex_buffer_num_elements(const b200osd_vertex_buffer *vb);
B200OSD_API int    b200osd_vertex_buffer_num_vertices(const b200osd_vertex_buffer *vb);
B200OSD_API float *b200osd_vertex_buffer_bind(b200osd_vertex_buffer *vb);     /* BindCudaBuffer(): device pointer */
/* UpdateData(src,startVertex,numVertices): host -> device.  stream == NULL: blocking cudaMemcpy like the
 * reference (cudaVertexBuffer.cpp:56-64); otherwise cudaMemcpyAsync on that stream (pin hostSrc for overlap). */
B200OSD_API int    b200osd_vertex_buffer_update(b200osd_vertex_buffer *vb, const float *hostSrc,
                                                int startVertex, int numVertices, void *stream);
/* device -> host read-back of a vertex range (what clients do after Synchronize()). */
B200OSD_API int    b200osd_vertex_buffer_read(b200osd_vertex_buffer *vb, float *hostDst,
                                              int startVertex, int numVertices, void *stream);

/* ---- stencil table (CudaStencilTable, osd/cudaEvaluator.h:52-92) ------------------------------
 * Host arrays in Far::StencilTable / Far::LimitStencilTable layout (far/stencilTable.h:156-186,434-456):
 * sizes[n], offsets[n], indices[ne], weights[ne] and optional derivative weights (NULL = absent).
 * The table keeps (a) verbatim device copies (the Get*Buffer() pointers of the reference class) and
 * (b) a B200-specific bucketed copy (rows sorted by size inside windows, 32-row slices stored
 * element-major for coalesced 128-bit loads) used by b200osd_stencil_table_eval.
 * numControlVertices: Far::StencilTable::GetNumControlVertices() (far/stencilTable.h:161), or <= 0 for "1 + largest
 * index".  Given the real count, a table built with factorizeIntermediateLevels = false (far/stencilTableFactory.h:66-75,
 * far tutorial 4_3) is recognised by indices that reach past the control vertices: its level-l rows index the vertices of
 * level l-1, so a caller applies it one level at a time -- src = the previous level's block, [start,end) = the level's
 * rows -- which the absolute row range supports as is.  An evaluation of such a table whose source extent overlaps the
 * rows it writes (e.g. all levels in one call on the Osd::Mesh::Refine layout) is order dependent even in the sequential
 * CPU evaluator; it is refused with B200OSD_ERR_UNSUPPORTED instead of racing.
 * flags: bit 0 = skip the bucketed copy (verbatim only); bit 1 = also order rows by locality inside a window;
 * bit 2 = keep 32-bit indices even when a slice fits 16-bit offsets.
 * Summation order: by default the elements of rows of <= 16 terms (every row of a refined regular mesh) are summed in
 * control-index order instead of the table's own order -- neighbouring rows then gather the same vertices at the same
 * step (-7 % time on config 2); same terms, measured <= 3.4e-7 of sum|w||x| away from the reference order on every
 * fixture.  bit 4 (16) = keep the table's order everywhere (bit-identical to the reference's CUDA kernel);
 * bit 3 (8) = sort every row, including rows of 100+ terms (which may then differ by > 1e-6).
 * The bucketed copy is built on the device from the uploaded arrays (the host orders the rows by size and lays out the
 * slices; the passes over the elements are kernels); bit 5 (32) = build it on the host instead (the same bytes; bits 1 and
 * 3 imply it).
 * bit 6 (64) = REFERENCE-EXACT arithmetic: only the verbatim layout is kept and rows are evaluated thread per row with a
 * rounded product added to the running sum in the table's order -- operation for operation what osd/cpuKernel.cpp does,
 * so every output is BIT-IDENTICAL to Osd::CpuEvaluator / Osd::OmpEvaluator whatever the row length (two correct fp32
 * summations of n terms, e.g. the fused multiply-adds of the fast path, may differ by ~n * 2^-23 of sum|w||x|).  About half
 * the speed of the bucketed path; meant for validation and for clients that diff against CPU results.                   */
B200OSD_API b200osd_stencil_table *b200osd_stencil_table_create(
        int numStencils, int numControlVertices,
        const int *sizes, const int *offsets, const int *indices, const float *weights,
        const float *duWeights, const float *dvWeights,
        const float *duuWeights, const float *duvWeights, const float *dvvWeights, int flags);
/* The same from DEVICE arrays in the reference layout -- what a client holds that already built an Osd::CudaStencilTable
 * (osd/cudaEvaluator.h:57-90: GetSizesBuffer() ... GetDvvWeightsBuffer()).  Instead of evaluating those arrays row by row
 * through b200osd_eval_stencils (thread per row, ~55 % of the roofline), convert once and evaluate through the table.
 * Synchronous; the client's arrays are only read (copied device to device; the host sees the row sizes and offsets only). */
B200OSD_API b200osd_stencil_table *b200osd_stencil_table_create_from_device(
        int numStencils, int numControlVertices,
        const int *sizes, const int *offsets, const int *indices, const float *weights,
        const float *duWeights, const float *dvWeights,
        const float *duuWeights, const float *duvWeights, const float *dvvWeights, int flags);
B200OSD_API void b200osd_stencil_table_destroy(b200osd_stencil_table *t);
B200OSD_API int  b200osd_stencil_table_num_stencils(const b200osd_stencil_table *t);
B200OSD_API int  b200osd_stencil_table_num_control_vertices(const b200osd_stencil_table *t);
B200OSD_API int  b200osd_stencil_table_is_factorized(const b200osd_stencil_table *t);  /* 0: indices reach past the control vertices */
B200OSD_API long long b200osd_stencil_table_num_elements(const b200osd_stencil_table *t);
/* which: 0 sizes, 1 offsets, 2 indices, 3 weights, 4 du, 5 dv, 6 duu, 7 duv, 8 dvv -> device pointer or NULL */
B200OSD_API const void *b200osd_stencil_table_buffer(const b200osd_stencil_table *t, int which);
/* bytes of the bucketed device copy actually streamed per full evaluation with nOut weight streams */
B200OSD_API long long b200osd_stencil_table_stream_bytes(const b200osd_stencil_table *t, int nOut);

/* EvalStencils through the table's bucketed layout (fast path). */
B200OSD_API int b200osd_stencil_table_eval(const b200osd_stencil_table *t,
        const float *src, const int srcDesc[3],
        int nOut, float *const dsts[], const int dstDescs[][3],
        int start, int end, void *stream);

/* Batched instances sharing one topology (reference pattern: one EvalStencils call per instance with
 * shifted descriptors, examples/glShareTopology/meshRefiner.h:68-88).  Instance b (0 <= b < numInstances) uses
 * srcDesc.offset + b*srcInstanceStride and dstDesc.offset + b*dstInstanceStride (strides in floats).  The table streams
 * are read once per chunk of up to 4 instances.  Results are bit-identical to numInstances separate calls. */
B200OSD_API int b200osd_stencil_table_eval_batched(const b200osd_stencil_table *t,
        const float *src, const int srcDesc[3], float *dst, const int dstDesc[3],
        int numInstances, long long srcInstanceStride, long long dstInstanceStride,
        int start, int end, void *stream);

/* EvalStencils on plain reference-layout DEVICE arrays (raw CudaEvaluator::EvalStencils overloads,
 * osd/cudaEvaluator.h:171-178,284-295,449-466).  weights[k] for k < nOut. */
B200OSD_API int b200osd_eval_stencils(
        const float *src, const int srcDesc[3],
        int nOut, float *const dsts[], const int dstDescs[][3],
        const int *sizes, const int *offsets, const int *indices, const float *const weights[],
        int start, int end, void *stream);

/* ---- patch table (CudaPatchTable, osd/cudaPatchTable.h:51-112) --------------------------------
 * A patch table is a set of (PatchArray[], index buffer, PatchParam[]) triples as flattened by
 * Osd::CpuPatchTable (osd/cpuPatchTable.cpp:35-156): which = 0 vertex, 1 varying (params shared with
 * vertex: pass NULL/0), 2+c face-varying channel c.                                               */
B200OSD_API b200osd_patch_table *b200osd_patch_table_create(int numFVarChannels);
B200OSD_API void b200osd_patch_table_destroy(b200osd_patch_table *t);
B200OSD_API int  b200osd_patch_table_set(b200osd_patch_table *t, int which,
        int numArrays, const b200osd_patch_array *arrays,
        int numIndices, const int *indices,
        int numParams, const b200osd_patch_param *params);
B200OSD_API int  b200osd_patch_table_num_fvar_channels(const b200osd_patch_table *t);
/* kind: 0 PatchArray buffer, 1 index buffer, 2 PatchParam buffer -> device pointer or NULL */
B200OSD_API const void *b200osd_patch_table_buffer(const b200osd_patch_table *t, int which, int kind);
B200OSD_API int  b200osd_patch_table_count(const b200osd_patch_table *t, int which, int kind);

/* EvalPatches / EvalPatchesVarying / EvalPatchesFaceVarying on DEVICE arrays (raw overloads,
 * osd/cudaEvaluator.h:706-713,752-761,815-827): the caller picks the triple, exactly like the
 * reference templates do (cudaEvaluator.h:502-523, 857-878, 1068-1090). */
B200OSD_API int b200osd_eval_patches(
        const float *src, const int srcDesc[3],
        int nOut, float *const dsts[], const int dstDescs[][3],
        int numPatchCoords, const b200osd_patch_coord *patchCoords,
        const b200osd_patch_array *patchArrays, const int *patchIndices,
        const b200osd_patch_param *patchParams, void *stream);

/* Evaluation options.  The reference chooses at BUILD time (cmake -DOPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES=ON,
 * CMakeLists.txt:340,640) whether the derivative weights of the 8 interior points of a GREGORY_BASIS patch are the
 * default approximation (osd/patchBasis.h:421-440) or the true derivatives of the rational blend (:441-487; Far does
 * the same, far/patchBasis.cpp:462).  Here it is a run-time option of the call or the table; the C++ headers in
 * include/b200osd/ set it when the client is compiled with that macro, so a drop-in build keeps its behaviour. */
#define B200OSD_PATCH_GREGORY_TRUE_DERIVATIVES 1
/* b200osd_eval_patches with an `options` mask (0 = b200osd_eval_patches) */
B200OSD_API int b200osd_eval_patches_ex(
        const float *src, const int srcDesc[3],
        int nOut, float *const dsts[], const int dstDescs[][3],
        int numPatchCoords, const b200osd_patch_coord *patchCoords,
        const b200osd_patch_array *patchArrays, const int *patchIndices,
        const b200osd_patch_param *patchParams, int options, void *stream);

/* EvalPatches through the table handle (fast path): which = 0 vertex, 1 varying, 2+c face-varying channel c.
 * Same contract as b200osd_eval_patches.  The table is immutable: any number of threads and streams may evaluate one
 * table concurrently.  Coordinates are evaluated in the caller's order, one warp per 32 of them; a warp copies every
 * DISTINCT control hull among its coordinates once into shared memory, so a set that arrives grouped by patch (sorted,
 * tessellation grids) costs one hull fetch per run of equal patches.  For large INCOHERENT sets (numPatchCoords >= 65536
 * and >= 2 x the number of patches, 12-20 point hulls) a sampling probe on the device switches the call to a per-call
 * hull cache: the index buffer is dereferenced once into scratch and every coordinate reads its hull as one contiguous
 * 192-240 byte block.  The scratch comes from a stream-ordered memory pool on `stream` (no cudaMalloc, no
 * synchronisation, legal during stream capture); results are bit-identical whichever way a call is served. */
B200OSD_API int b200osd_patch_table_eval(const b200osd_patch_table *t, int which,
        const float *src, const int srcDesc[3],
        int nOut, float *const dsts[], const int dstDescs[][3],
        int numPatchCoords, const b200osd_patch_coord *patchCoords, void *stream);
/* how b200osd_patch_table_eval serves calls on this table (bench / tests): 0 = automatic (above), 1 = always the caller's
 * order, 2 = group the coordinates by patch on the device per call (counting sort; results still land at the caller's
 * index), 3 = always the per-call hull cache */
B200OSD_API void b200osd_patch_table_set_variant(b200osd_patch_table *t, int variant);
B200OSD_API int  b200osd_patch_table_get_variant(const b200osd_patch_table *t);
/* evaluation options (B200OSD_PATCH_*) of every call through this table: b200osd_patch_table_eval, patch plans, and the
 * limit-stencil tables b200osd_limit_stencil_table_create builds from it.  Set before the table is shared. */
B200OSD_API int  b200osd_patch_table_set_options(b200osd_patch_table *t, int options);
B200OSD_API int  b200osd_patch_table_get_options(const b200osd_patch_table *t);

/* ---- patch plan: a cached grouping of ONE coordinate set (the per-use state an "instantiatable" evaluator owns in the
 * reference design: osd/mesh.h:305-409, osd/glComputeEvaluator.h:98-128) ------------------------------------------------
 * _create allocates scratch for up to maxPatchCoords coordinates (the only allocation); _bin groups the coordinates at
 * `patchCoords` by patch on the device (asynchronous on `stream`); _eval evaluates that same set (same pointer, same
 * count: anything else returns B200OSD_ERR_INVALID) through the cached grouping -- vertex, varying and every face-varying
 * slot share one grouping because PatchCoord.handle.patchIndex means the same patch in all of them.  Re-bin after the
 * coordinates change.  One plan serves one stream at a time. */
typedef struct b200osd_patch_plan b200osd_patch_plan;
B200OSD_API b200osd_patch_plan *b200osd_patch_plan_create(const b200osd_patch_table *t, int maxPatchCoords);
B200OSD_API void b200osd_patch_plan_destroy(b200osd_patch_plan *p);
B200OSD_API int  b200osd_patch_plan_capacity(const b200osd_patch_plan *p);
B200OSD_API int  b200osd_patch_plan_bin(b200osd_patch_plan *p, int numPatchCoords,
        const b200osd_patch_coord *patchCoords, void *stream);
B200OSD_API int  b200osd_patch_plan_eval(const b200osd_patch_plan *p, int which,
        const float *src, const int srcDesc[3],
        int nOut, float *const dsts[], const int dstDescs[][3],
        int numPatchCoords, const b200osd_patch_coord *patchCoords, void *stream);

/* ---- limit-stencil construction on the device -------------------------------------------------------
 * The per-location loop of Far::LimitStencilTableFactory::Create (far/stencilTableFactory.cpp:559-662) with the merge of
 * Far's StencilBuilder (far/stencilBuilder.cpp): for each located sample (a DEVICE PatchCoord, e.g. from
 * b200osd_patch_map_find; arrayIndex < 0 produces no row, like FindPatch == NULL) the basis weights of its patch (value;
 * + 1st; + 2nd derivatives: numWeightSets = 1, 3, 6) are combined with the stencils of the patch's control points.
 * cvStencils = the refined + local-point stencil table of the same topology WITHOUT control-vertex rows (row r belongs to
 * vertex numControlVertices + r; the table Osd::Mesh builds, osd/mesh.h:588-659).  The result is the table Far builds --
 * same rows, same element order; for Catmark patches bit-identical weights -- as a stencil table handle that owns its
 * device arrays.  flags as for b200osd_stencil_table_create (bit 0: keep only the reference-layout arrays, no read-back).
 * Synchronous (it allocates the result).  A stencil may reference at most 512 control vertices. */
B200OSD_API b200osd_stencil_table *b200osd_limit_stencil_table_create(
        const b200osd_patch_table *patchTable, const b200osd_stencil_table *cvStencils,
        int numLocations, const b200osd_patch_coord *patchCoords, int numWeightSets, int flags, void *stream);

/* ---- patch map (Far::PatchMap, far/patchMap.h:48-217) ---------------------------
 * Locates samples given as (ptex face, s, t) in the patches of a table and writes Osd::PatchCoord records
 * (osd/types.h:53-54) ready for EvalPatches -- the reference does this one sample at a time on the host
 * (examples/glEvalLimit/particles.cpp:91-115,392-394).  The map is built once per topology, on the host, from the
 * VERTEX PatchArray[] and PatchParam[] (as flattened by Osd::CpuPatchTable) and lives in device memory.
 * patchesAreTriangular mirrors far/patchMap.cpp:93-94 (the varying descriptor has 3 control vertices: Loop tables). */
B200OSD_API b200osd_patch_map *b200osd_patch_map_create(int numArrays, const b200osd_patch_array *vertexArrays,
        int numPatches, const b200osd_patch_param *patchParams, int patchesAreTriangular);
B200OSD_API void b200osd_patch_map_destroy(b200osd_patch_map *m);
/* info = {minPatchFace, maxPatchFace, maxDepth, triangular, numNodes, numHandles} */
B200OSD_API int  b200osd_patch_map_info(const b200osd_patch_map *m, int info[6]);
/* DEVICE arrays in, DEVICE records out; strides in elements (1,1,1 for three packed arrays; 3,3,3 with the pointers
 * offset by one word each for {int face; float s; float t} records).  A sample that hits no patch (a hole, or a
 * face outside the map: FindPatch returns NULL) gets handle.arrayIndex = -1 and keeps its (s,t); EvalPatches in this
 * library leaves the outputs of such records untouched.  numFound (device int, may be NULL) receives the hit count. */
B200OSD_API int  b200osd_patch_map_find(const b200osd_patch_map *m, int numSamples,
        const int *ptexFace, int faceStride, const float *s, int sStride, const float *t, int tStride,
        b200osd_patch_coord *outCoords, int *numFound, void *stream);

/* ---- frame capture ------------------------------------------------------------------------
 * A frame object owns a CUDA stream.  Calls issued with that stream between _begin and _end -- EvalStencils,
 * FindPatches, EvalPatches*, vertex-buffer updates from pinned memory -- are recorded instead of executed; _launch
 * replays them as one graph launch on the same stream.  Run the frame once eagerly first (the stream-ordered scratch
 * pool of b200osd_patch_table_eval fills on first use).  No reference counterpart. */
B200OSD_API b200osd_frame *b200osd_frame_create(void);
B200OSD_API void  b200osd_frame_destroy(b200osd_frame *f);
B200OSD_API void *b200osd_frame_stream(const b200osd_frame *f);     /* the cudaStream_t to pass as `stream` */
/* A second, high-priority stream that belongs to the same frame: while recording, work issued on it becomes a PARALLEL
 * branch of the graph (typically b200osd_comm_broadcast of the next frame's control points next to this frame's
 * evaluation).  _fence orders the two: mainWaitsForSide = 0 makes later side-stream work wait for everything issued so
 * far on the main stream, 1 the other way round.  _begin forks and _end joins the side stream automatically. */
B200OSD_API void *b200osd_frame_side_stream(const b200osd_frame *f);
B200OSD_API int   b200osd_frame_fence(b200osd_frame *f, int mainWaitsForSide);
/* Keeps [devPtr, devPtr + bytes) -- e.g. the refined vertices that EvalStencils writes and EvalPatches reads -- resident
 * in L2 across the kernels of the frame (cudaAccessPolicyWindow, persisting) while everything else streams; call before
 * _begin so that the recorded kernels inherit it.  devPtr = NULL clears the window. */
B200OSD_API int   b200osd_frame_set_l2_window(b200osd_frame *f, const void *devPtr, size_t bytes, float hitRatio);
B200OSD_API int   b200osd_frame_begin(b200osd_frame *f);
B200OSD_API int   b200osd_frame_end(b200osd_frame *f);
B200OSD_API int   b200osd_frame_launch(b200osd_frame *f);           /* asynchronous */
B200OSD_API int   b200osd_frame_synchronize(b200osd_frame *f);

/* ---- multi-GPU data plane (no reference counterpart -- the reference is single-device, its hook is the
 * absolute row range [start, end) of the raw EvalStencils overloads, osd/cudaEvaluator.h:171-178) ------------------------
 * One process per GPU.  Rows / coordinates are cut into contiguous ranges, tables are static and pre-sharded, outputs stay
 * sharded, no row spans ranks (no reduction); the only exchange is the per-frame replication of the control points. */
/* `world` contiguous row ranges [ranges[2r], ranges[2r+1]) of equal cost (a row costs its elements + 1); interior cuts are
 * rounded to a multiple of `align` (e.g. the 2048-row bucketing window).  Host only: needs no device. */
B200OSD_API int b200osd_shard_plan(int numStencils, const int *sizes, int world, int align, int *ranges);
/* Strong scaling of ONE mesh wants more than balance: a contiguous row range of a Far table references control vertices
 * from all over the mesh (its face-, edge- and vertex-points are separate blocks), so every rank would need every control
 * point every frame.  This plan orders the rows by the smallest control vertex they reference (stable; rowOrder[numStencils]
 * = the rows in that order) and cuts THAT order into `world` chunks of equal cost (ranges[2r], ranges[2r+1]) = positions
 * in rowOrder).  A rank builds its table from its rows and, per frame, needs only the control vertices
 * [controlRanges[2r], controlRanges[2r+1]) -- on a mesh whose vertex numbering has any locality about 1/world of them plus
 * a halo, pulled with b200osd_window_get.  Outputs stay sharded, in rowOrder.  Host only: needs no device. */
B200OSD_API int b200osd_shard_plan_locality(int numStencils, const int *sizes, const int *offsets, const int *indices,
                                            int world, int *rowOrder, int *ranges, int *controlRanges);
/* The control vertices a (local) table references, as at most maxRuns index runs [runs[2k], runs[2k+1]) at `granularity`
 * vertices (a closed mesh has a seam: the bounding interval of a chunk can be the whole mesh while two runs are 1/world of
 * it).  The closest runs are merged when there are more than maxRuns.  Returns the number of runs, -1 on bad arguments.
 * These are the per-frame transfers of a rank: one b200osd_window_get per run.  Host only. */
B200OSD_API int b200osd_shard_control_runs(int numStencils, const int *sizes, const int *offsets, const int *indices,
                                           int granularity, int maxRuns, int *runs);
/* the same for a PatchCoord set (every coordinate costs the same) */
B200OSD_API int b200osd_shard_coords(long long numCoords, int world, int align, long long *ranges);
/* Communicator over NCCL (bound at run time with dlopen: no link-time dependency).  Rank 0 calls _unique_id and hands the
 * 128 bytes to the other ranks by any means (a file, MPI, torch.distributed); every rank then calls _create with its own
 * CUDA device current.  All transfer calls are asynchronous on `stream` and may be recorded into a frame. */
typedef struct b200osd_comm b200osd_comm;
B200OSD_API int  b200osd_comm_available(void);
B200OSD_API int  b200osd_comm_unique_id(char id[128]);
B200OSD_API b200osd_comm *b200osd_comm_create(int world, int rank, const char id[128]);
/* maxCTAs > 0 caps the thread blocks NCCL may use for this communicator's transfers (ncclConfig_t.maxCTAs): a few-MB
 * exchange that hides behind an evaluation kernel should cost one or two SMs, not sixteen; 0 = NCCL's default */
B200OSD_API b200osd_comm *b200osd_comm_create_ex(int world, int rank, const char id[128], int maxCTAs);
B200OSD_API void b200osd_comm_destroy(b200osd_comm *c);
B200OSD_API int  b200osd_comm_world(const b200osd_comm *c);
B200OSD_API int  b200osd_comm_rank(const b200osd_comm *c);
/* every rank ends up with root's `count` floats at buf (one mesh, rows sharded: all ranks need all control points) */
B200OSD_API int  b200osd_comm_broadcast(b200osd_comm *c, float *buf, size_t count, int root, void *stream);
/* rank r receives floats [r*countPerRank, (r+1)*countPerRank) of root's sendbuf into recvbuf (N meshes, rank r owns mesh
 * r: each GPU receives only what it reads) */
B200OSD_API int  b200osd_comm_scatter(b200osd_comm *c, const float *sendbuf, float *recvbuf, size_t countPerRank, int root, void *stream);
/* optional: every rank gets every rank's countPerRank floats (a consumer that wants the whole refined buffer) */
B200OSD_API int  b200osd_comm_all_gather(b200osd_comm *c, const float *sendbuf, float *recvbuf, size_t countPerRank, void *stream);

/* Peer-memory window: one-sided exchange over NVLink WITHOUT collective kernels.  _create (collective over `c`) allocates
 * `bytes` of device memory on every rank and maps every rank's block into every other rank (CUDA IPC); _local is this
 * rank's block.  _get copies from any rank's block into local memory by DMA (copy engines: no SM is taken from a kernel
 * running next to it); _signal / _wait order streams ACROSS ranks through per-(slot, sender) counters in peer memory,
 * written and polled by one-thread kernels: work issued on `stream` after _wait(src, slot) runs only once the matching
 * _signal(dst, slot) of rank src -- issued on ITS stream after the data was produced -- has executed.  rank = -1 means
 * every peer.  16 slots.  A wait gives up after 10 s and sets the error flag (b200osd_window_error != 0). */
typedef struct b200osd_window b200osd_window;
B200OSD_API b200osd_window *b200osd_window_create(b200osd_comm *c, size_t bytes);
B200OSD_API void   b200osd_window_destroy(b200osd_window *w);
B200OSD_API void  *b200osd_window_local(const b200osd_window *w);
B200OSD_API size_t b200osd_window_bytes(const b200osd_window *w);
B200OSD_API int    b200osd_window_get(b200osd_window *w, int srcRank, size_t srcOffsetBytes, void *dst, size_t bytes, void *stream);
/* wait + copy + signal as ONE kernel (SM loads over NVLink peer memory instead of a copy engine; for per-frame exchanges of
 * a few hundred KB, where the DMA set-up and two extra launches dominate): waits for the next signal of (waitSlot, srcRank)
 * when waitSlot >= 0, copies numRuns (<= 8) byte ranges [srcOffsetBytes[r], +bytes[r]) of srcRank's window to dsts[r]
 * (multiples of 4 bytes), then signals (signalSlot -> signalRank, -1 = every other rank) when signalSlot >= 0.  One pull at
 * a time per window (issue them on one stream). */
B200OSD_API int    b200osd_window_pull(b200osd_window *w, int srcRank, int waitSlot, int numRuns, const size_t *srcOffsetBytes,
                                       void *const *dsts, const size_t *bytes, int signalRank, int signalSlot, void *stream);
B200OSD_API int    b200osd_window_signal(b200osd_window *w, int dstRank, int slot, void *stream);
B200OSD_API int    b200osd_window_wait(b200osd_window *w, int srcRank, int slot, void *stream);
B200OSD_API int    b200osd_window_error(b200osd_window *w);

/* ---- tuning / introspection (used by bench.py and the tests; not needed by clients) ----------
 * Kernel variant of b200osd_stencil_table_eval for THIS table (there is no process-wide state): 0 = auto,
 * 1 = reference-layout (CSR) kernel, 2 = scalar gathers, 8 = persistent grid, 11 = 8 resident blocks/SM, 12 = 8 + 11. */
B200OSD_API void b200osd_stencil_table_set_variant(b200osd_stencil_table *t, int variant);
B200OSD_API int  b200osd_stencil_table_get_variant(const b200osd_stencil_table *t);

#ifdef __cplusplus
}
#endif
#endif /* B200OSD_CAPI_H */
