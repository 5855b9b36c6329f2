// b200Evaluator.h -- drop-in for Osd::CudaEvaluator (opensubdiv/osd/cudaEvaluator.h:94-1262), header-only over the
// C ABI (include/b200osd_capi.h).  Same static member names, arities, argument meaning and bool results:
//
//   EvalStencils            x3 template forms (value / +du,dv / +du,dv,duu,duv,dvv)   cudaEvaluator.h:125-143,217-242,352-386
//                           x3 raw device-pointer forms                                cudaEvaluator.h:171-178,284-295,449-466
//   EvalPatches             x3 template + x3 raw                                       cudaEvaluator.h:502-677,706-827
//   EvalPatchesVarying      x3 template                                                cudaEvaluator.h:857-1036
//   EvalPatchesFaceVarying  x3 template (+ fvarChannel)                                cudaEvaluator.h:1068-1254
//   Synchronize                                                                        cudaEvaluator.h:1261
//
// Usage is identical to the reference backend:
//   Osd::Mesh<Osd::B200VertexBuffer, Osd::B200StencilTable, Osd::B200Evaluator, Osd::B200PatchTable> mesh(refiner, 3, 0, level, bits);
//
// When the stencil table argument is a B200StencilTable the call goes through its bucketed device layout (fast path);
// any other table type exposing the Get*Buffer() device pointers, and the raw overloads, use the reference-layout
// kernels.  Differences from CudaEvaluator, all deliberate: a length mismatch returns false like Osd::CpuEvaluator
// (cpuEvaluator.cpp:47; the CUDA backend does not check), and `deviceContext` may point at a B200DeviceContext to
// choose a stream.
//
// The evaluator is also INSTANTIATABLE (osd/mesh.h:305-409, modelled on osd/glComputeEvaluator.h:98-128): Create()
// returns an instance that Osd::EvaluatorCacheT can cache per descriptor set and that every static Eval* accepts as
// `instance`.  What an instance caches is the grouping of one PatchCoord set by patch (BindPatchCoords): EvalPatches*
// calls on that same coordinate buffer then skip the per-call device sort.
#ifndef B200OSD_EVALUATOR_H
#define B200OSD_EVALUATOR_H

#include <opensubdiv/version.h>
#include <opensubdiv/osd/bufferDescriptor.h>
#include <opensubdiv/osd/types.h>

#include "../b200osd_capi.h"
#include "b200PatchTable.h"
#include "b200StencilTable.h"
#include "b200VertexBuffer.h"

namespace OpenSubdiv {
namespace OPENSUBDIV_VERSION {
namespace Osd {

class B200Evaluator {
public:
    // ------------------------------------------------------------------------------------- instantiation ----
    typedef bool Instantiatable;

    static B200Evaluator *Create(BufferDescriptor const &srcDesc, BufferDescriptor const &dstDesc,
                                 BufferDescriptor const &duDesc, BufferDescriptor const &dvDesc,
                                 void *deviceContext = NULL) {
        return Create(srcDesc, dstDesc, duDesc, dvDesc, BufferDescriptor(), BufferDescriptor(), BufferDescriptor(), deviceContext);
    }
    static B200Evaluator *Create(BufferDescriptor const &srcDesc, BufferDescriptor const &dstDesc,
                                 BufferDescriptor const &duDesc, BufferDescriptor const &dvDesc,
                                 BufferDescriptor const &duuDesc, BufferDescriptor const &duvDesc,
                                 BufferDescriptor const &dvvDesc, void *deviceContext = NULL) {
        (void)srcDesc; (void)dstDesc; (void)duDesc; (void)dvDesc; (void)duuDesc; (void)duvDesc; (void)dvvDesc; (void)deviceContext;
        // nothing to compile per descriptor set (the kernels are specialised at build time); the instance exists for
        // its cached PatchCoord grouping
        return new B200Evaluator();
    }
    ~B200Evaluator() { b200osd_patch_plan_destroy(_plan); }

    /// Groups the `numPatchCoords` coordinates of `patchCoords` by patch on the device and keeps the grouping: every
    /// later EvalPatches / EvalPatchesVarying / EvalPatchesFaceVarying call given this instance, this table and this
    /// same coordinate buffer reuses it.  Call again after the coordinates change.  (Scratch is allocated here, when the
    /// table changes or the set outgrows the previous one -- never inside an Eval call.)
    template <typename PATCHCOORD_BUFFER>
    bool BindPatchCoords(int numPatchCoords, PATCHCOORD_BUFFER *patchCoords, B200PatchTable const *patchTable,
                         void *deviceContext = NULL) {
        if (!patchTable) return false;
        if (!_plan || _planTable != patchTable->GetHandle() || b200osd_patch_plan_capacity(_plan) < numPatchCoords) {
            b200osd_patch_plan_destroy(_plan);
            _plan = b200osd_patch_plan_create(patchTable->GetHandle(), numPatchCoords);
            _planTable = patchTable->GetHandle();
            if (!_plan) return false;
        }
        _planCoords = (const void *)patchCoords->BindCudaBuffer();
        _planCount = numPatchCoords;
        return b200osd_patch_plan_bin(_plan, numPatchCoords, (const b200osd_patch_coord *)_planCoords,
                                      B200StreamOf(deviceContext)) == B200OSD_OK;
    }

    // ------------------------------------------------------------------------------------------ stencils ----
    template <typename SRC_BUFFER, typename DST_BUFFER, typename STENCIL_TABLE>
    static bool EvalStencils(SRC_BUFFER *srcBuffer, BufferDescriptor const &srcDesc,
                             DST_BUFFER *dstBuffer, BufferDescriptor const &dstDesc,
                             STENCIL_TABLE const *stencilTable,
                             const B200Evaluator *instance = NULL, void *deviceContext = NULL) {
        (void)instance;
        float *dsts[1] = { dstBuffer->BindCudaBuffer() };
        BufferDescriptor descs[1] = { dstDesc };
        return evalTable(srcBuffer->BindCudaBuffer(), srcDesc, 1, dsts, descs, stencilTable, deviceContext);
    }

    template <typename SRC_BUFFER, typename DST_BUFFER, typename STENCIL_TABLE>
    static bool EvalStencils(SRC_BUFFER *srcBuffer, BufferDescriptor const &srcDesc,
                             DST_BUFFER *dstBuffer, BufferDescriptor const &dstDesc,
                             DST_BUFFER *duBuffer, BufferDescriptor const &duDesc,
                             DST_BUFFER *dvBuffer, BufferDescriptor const &dvDesc,
                             STENCIL_TABLE const *stencilTable,
                             const B200Evaluator *instance = NULL, void *deviceContext = NULL) {
        (void)instance;
        float *dsts[3] = { dstBuffer->BindCudaBuffer(), duBuffer->BindCudaBuffer(), dvBuffer->BindCudaBuffer() };
        BufferDescriptor descs[3] = { dstDesc, duDesc, dvDesc };
        return evalTable(srcBuffer->BindCudaBuffer(), srcDesc, 3, dsts, descs, stencilTable, deviceContext);
    }

    template <typename SRC_BUFFER, typename DST_BUFFER, typename STENCIL_TABLE>
    static bool EvalStencils(SRC_BUFFER *srcBuffer, BufferDescriptor const &srcDesc,
                             DST_BUFFER *dstBuffer, BufferDescriptor const &dstDesc,
                             DST_BUFFER *duBuffer, BufferDescriptor const &duDesc,
                             DST_BUFFER *dvBuffer, BufferDescriptor const &dvDesc,
                             DST_BUFFER *duuBuffer, BufferDescriptor const &duuDesc,
                             DST_BUFFER *duvBuffer, BufferDescriptor const &duvDesc,
                             DST_BUFFER *dvvBuffer, BufferDescriptor const &dvvDesc,
                             STENCIL_TABLE const *stencilTable,
                             const B200Evaluator *instance = NULL, void *deviceContext = NULL) {
        (void)instance;
        float *dsts[6] = { dstBuffer->BindCudaBuffer(), duBuffer->BindCudaBuffer(), dvBuffer->BindCudaBuffer(),
                           duuBuffer->BindCudaBuffer(), duvBuffer->BindCudaBuffer(), dvvBuffer->BindCudaBuffer() };
        BufferDescriptor descs[6] = { dstDesc, duDesc, dvDesc, duuDesc, duvDesc, dvvDesc };
        return evalTable(srcBuffer->BindCudaBuffer(), srcDesc, 6, dsts, descs, stencilTable, deviceContext);
    }

    /// Extension (not in the reference): `numInstances` control-point sets sharing one topology refined in one pass.
    /// Instance b uses srcDesc.offset + b*srcInstanceStride and dstDesc.offset + b*dstInstanceStride (in floats) --
    /// the batched form of the per-instance loop in examples/glShareTopology/meshRefiner.h:68-88, bit-identical to it.
    template <typename SRC_BUFFER, typename DST_BUFFER>
    static bool EvalStencilsBatched(SRC_BUFFER *srcBuffer, BufferDescriptor const &srcDesc,
                                    DST_BUFFER *dstBuffer, BufferDescriptor const &dstDesc,
                                    B200StencilTable const *stencilTable, int numInstances,
                                    long long srcInstanceStride, long long dstInstanceStride, void *deviceContext = NULL) {
        int sd[3] = { srcDesc.offset, srcDesc.length, srcDesc.stride };
        int dd[3] = { dstDesc.offset, dstDesc.length, dstDesc.stride };
        return b200osd_stencil_table_eval_batched(stencilTable->GetHandle(), srcBuffer->BindCudaBuffer(), sd,
                                                  dstBuffer->BindCudaBuffer(), dd, numInstances, srcInstanceStride,
                                                  dstInstanceStride, 0, stencilTable->GetNumStencils(),
                                                  B200StreamOf(deviceContext)) == B200OSD_OK;
    }

    // raw device-pointer forms on reference-layout arrays
    static bool EvalStencils(const float *src, BufferDescriptor const &srcDesc,
                             float *dst, BufferDescriptor const &dstDesc,
                             const int *sizes, const int *offsets, const int *indices, const float *weights,
                             int start, int end) {
        float *dsts[1] = { dst };
        BufferDescriptor descs[1] = { dstDesc };
        const float *w[1] = { weights };
        return evalRaw(src, srcDesc, 1, dsts, descs, sizes, offsets, indices, w, start, end, NULL);
    }

    static bool EvalStencils(const float *src, BufferDescriptor const &srcDesc,
                             float *dst, BufferDescriptor const &dstDesc,
                             float *du, BufferDescriptor const &duDesc,
                             float *dv, BufferDescriptor const &dvDesc,
                             const int *sizes, const int *offsets, const int *indices,
                             const float *weights, const float *duWeights, const float *dvWeights,
                             int start, int end) {
        float *dsts[3] = { dst, du, dv };
        BufferDescriptor descs[3] = { dstDesc, duDesc, dvDesc };
        const float *w[3] = { weights, duWeights, dvWeights };
        return evalRaw(src, srcDesc, 3, dsts, descs, sizes, offsets, indices, w, start, end, NULL);
    }

    static bool EvalStencils(const float *src, BufferDescriptor const &srcDesc,
                             float *dst, BufferDescriptor const &dstDesc,
                             float *du, BufferDescriptor const &duDesc,
                             float *dv, BufferDescriptor const &dvDesc,
                             float *duu, BufferDescriptor const &duuDesc,
                             float *duv, BufferDescriptor const &duvDesc,
                             float *dvv, BufferDescriptor const &dvvDesc,
                             const int *sizes, const int *offsets, const int *indices,
                             const float *weights, const float *duWeights, const float *dvWeights,
                             const float *duuWeights, const float *duvWeights, const float *dvvWeights,
                             int start, int end) {
        float *dsts[6] = { dst, du, dv, duu, duv, dvv };
        BufferDescriptor descs[6] = { dstDesc, duDesc, dvDesc, duuDesc, duvDesc, dvvDesc };
        const float *w[6] = { weights, duWeights, dvWeights, duuWeights, duvWeights, dvvWeights };
        return evalRaw(src, srcDesc, 6, dsts, descs, sizes, offsets, indices, w, start, end, NULL);
    }

    // ------------------------------------------------------------------------------------------- patches ----
    // which: 0 vertex, 1 varying, 2+c face-varying channel c (resolved by evalPatchTable below)
#define B200OSD_PATCH_TRIPLE_VERTEX(pt)  (pt), 0
#define B200OSD_PATCH_TRIPLE_VARYING(pt) (pt), 1
#define B200OSD_PATCH_TRIPLE_FVAR(pt, c) (pt), 2 + (c)

    template <typename SRC_BUFFER, typename DST_BUFFER, typename PATCHCOORD_BUFFER, typename PATCH_TABLE>
    static bool EvalPatches(SRC_BUFFER *srcBuffer, BufferDescriptor const &srcDesc,
                            DST_BUFFER *dstBuffer, BufferDescriptor const &dstDesc,
                            int numPatchCoords, PATCHCOORD_BUFFER *patchCoords, PATCH_TABLE *patchTable,
                            B200Evaluator const *instance, void *deviceContext = NULL) {
        float *dsts[1] = { dstBuffer->BindCudaBuffer() };
        BufferDescriptor descs[1] = { dstDesc };
        return evalPatchTable(srcBuffer->BindCudaBuffer(), srcDesc, 1, dsts, descs, numPatchCoords,
                           patchCoords->BindCudaBuffer(), B200OSD_PATCH_TRIPLE_VERTEX(patchTable), instance, deviceContext);
    }

    template <typename SRC_BUFFER, typename DST_BUFFER, typename PATCHCOORD_BUFFER, typename PATCH_TABLE>
    static bool EvalPatches(SRC_BUFFER *srcBuffer, BufferDescriptor const &srcDesc,
                            DST_BUFFER *dstBuffer, BufferDescriptor const &dstDesc,
                            DST_BUFFER *duBuffer, BufferDescriptor const &duDesc,
                            DST_BUFFER *dvBuffer, BufferDescriptor const &dvDesc,
                            int numPatchCoords, PATCHCOORD_BUFFER *patchCoords, PATCH_TABLE *patchTable,
                            B200Evaluator const *instance, void *deviceContext = NULL) {
        float *dsts[3] = { dstBuffer->BindCudaBuffer(), duBuffer->BindCudaBuffer(), dvBuffer->BindCudaBuffer() };
        BufferDescriptor descs[3] = { dstDesc, duDesc, dvDesc };
        return evalPatchTable(srcBuffer->BindCudaBuffer(), srcDesc, 3, dsts, descs, numPatchCoords,
                           patchCoords->BindCudaBuffer(), B200OSD_PATCH_TRIPLE_VERTEX(patchTable), instance, deviceContext);
    }

    template <typename SRC_BUFFER, typename DST_BUFFER, typename PATCHCOORD_BUFFER, typename PATCH_TABLE>
    static bool EvalPatches(SRC_BUFFER *srcBuffer, BufferDescriptor const &srcDesc,
                            DST_BUFFER *dstBuffer, BufferDescriptor const &dstDesc,
                            DST_BUFFER *duBuffer, BufferDescriptor const &duDesc,
                            DST_BUFFER *dvBuffer, BufferDescriptor const &dvDesc,
                            DST_BUFFER *duuBuffer, BufferDescriptor const &duuDesc,
                            DST_BUFFER *duvBuffer, BufferDescriptor const &duvDesc,
                            DST_BUFFER *dvvBuffer, BufferDescriptor const &dvvDesc,
                            int numPatchCoords, PATCHCOORD_BUFFER *patchCoords, PATCH_TABLE *patchTable,
                            B200Evaluator const *instance, void *deviceContext = NULL) {
        float *dsts[6] = { dstBuffer->BindCudaBuffer(), duBuffer->BindCudaBuffer(), dvBuffer->BindCudaBuffer(),
                           duuBuffer->BindCudaBuffer(), duvBuffer->BindCudaBuffer(), dvvBuffer->BindCudaBuffer() };
        BufferDescriptor descs[6] = { dstDesc, duDesc, dvDesc, duuDesc, duvDesc, dvvDesc };
        return evalPatchTable(srcBuffer->BindCudaBuffer(), srcDesc, 6, dsts, descs, numPatchCoords,
                           patchCoords->BindCudaBuffer(), B200OSD_PATCH_TRIPLE_VERTEX(patchTable), instance, deviceContext);
    }

    // raw forms
    static bool EvalPatches(const float *src, BufferDescriptor const &srcDesc,
                            float *dst, BufferDescriptor const &dstDesc,
                            int numPatchCoords, const PatchCoord *patchCoords, const PatchArray *patchArrays,
                            const int *patchIndices, const PatchParam *patchParams) {
        float *dsts[1] = { dst };
        BufferDescriptor descs[1] = { dstDesc };
        return evalPatches(src, srcDesc, 1, dsts, descs, numPatchCoords, patchCoords, patchArrays, patchIndices,
                           patchParams, NULL);
    }

    static bool EvalPatches(const float *src, BufferDescriptor const &srcDesc,
                            float *dst, BufferDescriptor const &dstDesc,
                            float *du, BufferDescriptor const &duDesc,
                            float *dv, BufferDescriptor const &dvDesc,
                            int numPatchCoords, PatchCoord const *patchCoords, PatchArray const *patchArrays,
                            const int *patchIndices, PatchParam const *patchParams) {
        float *dsts[3] = { dst, du, dv };
        BufferDescriptor descs[3] = { dstDesc, duDesc, dvDesc };
        return evalPatches(src, srcDesc, 3, dsts, descs, numPatchCoords, patchCoords, patchArrays, patchIndices,
                           patchParams, NULL);
    }

    static bool EvalPatches(const float *src, BufferDescriptor const &srcDesc,
                            float *dst, BufferDescriptor const &dstDesc,
                            float *du, BufferDescriptor const &duDesc,
                            float *dv, BufferDescriptor const &dvDesc,
                            float *duu, BufferDescriptor const &duuDesc,
                            float *duv, BufferDescriptor const &duvDesc,
                            float *dvv, BufferDescriptor const &dvvDesc,
                            int numPatchCoords, PatchCoord const *patchCoords, PatchArray const *patchArrays,
                            const int *patchIndices, PatchParam const *patchParams) {
        float *dsts[6] = { dst, du, dv, duu, duv, dvv };
        BufferDescriptor descs[6] = { dstDesc, duDesc, dvDesc, duuDesc, duvDesc, dvvDesc };
        return evalPatches(src, srcDesc, 6, dsts, descs, numPatchCoords, patchCoords, patchArrays, patchIndices,
                           patchParams, NULL);
    }

    // varying: same kernel on the varying (linear) patch arrays + the vertex PatchParams (cudaEvaluator.h:857-878)
    template <typename SRC_BUFFER, typename DST_BUFFER, typename PATCHCOORD_BUFFER, typename PATCH_TABLE>
    static bool EvalPatchesVarying(SRC_BUFFER *srcBuffer, BufferDescriptor const &srcDesc,
                                   DST_BUFFER *dstBuffer, BufferDescriptor const &dstDesc,
                                   int numPatchCoords, PATCHCOORD_BUFFER *patchCoords, PATCH_TABLE *patchTable,
                                   B200Evaluator const *instance, void *deviceContext = NULL) {
        float *dsts[1] = { dstBuffer->BindCudaBuffer() };
        BufferDescriptor descs[1] = { dstDesc };
        return evalPatchTable(srcBuffer->BindCudaBuffer(), srcDesc, 1, dsts, descs, numPatchCoords,
                           patchCoords->BindCudaBuffer(), B200OSD_PATCH_TRIPLE_VARYING(patchTable), instance, deviceContext);
    }

    template <typename SRC_BUFFER, typename DST_BUFFER, typename PATCHCOORD_BUFFER, typename PATCH_TABLE>
    static bool EvalPatchesVarying(SRC_BUFFER *srcBuffer, BufferDescriptor const &srcDesc,
                                   DST_BUFFER *dstBuffer, BufferDescriptor const &dstDesc,
                                   DST_BUFFER *duBuffer, BufferDescriptor const &duDesc,
                                   DST_BUFFER *dvBuffer, BufferDescriptor const &dvDesc,
                                   int numPatchCoords, PATCHCOORD_BUFFER *patchCoords, PATCH_TABLE *patchTable,
                                   B200Evaluator const *instance, void *deviceContext = NULL) {
        float *dsts[3] = { dstBuffer->BindCudaBuffer(), duBuffer->BindCudaBuffer(), dvBuffer->BindCudaBuffer() };
        BufferDescriptor descs[3] = { dstDesc, duDesc, dvDesc };
        return evalPatchTable(srcBuffer->BindCudaBuffer(), srcDesc, 3, dsts, descs, numPatchCoords,
                           patchCoords->BindCudaBuffer(), B200OSD_PATCH_TRIPLE_VARYING(patchTable), instance, deviceContext);
    }

    template <typename SRC_BUFFER, typename DST_BUFFER, typename PATCHCOORD_BUFFER, typename PATCH_TABLE>
    static bool EvalPatchesVarying(SRC_BUFFER *srcBuffer, BufferDescriptor const &srcDesc,
                                   DST_BUFFER *dstBuffer, BufferDescriptor const &dstDesc,
                                   DST_BUFFER *duBuffer, BufferDescriptor const &duDesc,
                                   DST_BUFFER *dvBuffer, BufferDescriptor const &dvDesc,
                                   DST_BUFFER *duuBuffer, BufferDescriptor const &duuDesc,
                                   DST_BUFFER *duvBuffer, BufferDescriptor const &duvDesc,
                                   DST_BUFFER *dvvBuffer, BufferDescriptor const &dvvDesc,
                                   int numPatchCoords, PATCHCOORD_BUFFER *patchCoords, PATCH_TABLE *patchTable,
                                   B200Evaluator const *instance, void *deviceContext = NULL) {
        float *dsts[6] = { dstBuffer->BindCudaBuffer(), duBuffer->BindCudaBuffer(), dvBuffer->BindCudaBuffer(),
                           duuBuffer->BindCudaBuffer(), duvBuffer->BindCudaBuffer(), dvvBuffer->BindCudaBuffer() };
        BufferDescriptor descs[6] = { dstDesc, duDesc, dvDesc, duuDesc, duvDesc, dvvDesc };
        return evalPatchTable(srcBuffer->BindCudaBuffer(), srcDesc, 6, dsts, descs, numPatchCoords,
                           patchCoords->BindCudaBuffer(), B200OSD_PATCH_TRIPLE_VARYING(patchTable), instance, deviceContext);
    }

    // face-varying: the channel's own (arrays, indices, params) triple (cudaEvaluator.h:1068-1090)
    template <typename SRC_BUFFER, typename DST_BUFFER, typename PATCHCOORD_BUFFER, typename PATCH_TABLE>
    static bool EvalPatchesFaceVarying(SRC_BUFFER *srcBuffer, BufferDescriptor const &srcDesc,
                                       DST_BUFFER *dstBuffer, BufferDescriptor const &dstDesc,
                                       int numPatchCoords, PATCHCOORD_BUFFER *patchCoords, PATCH_TABLE *patchTable,
                                       int fvarChannel, B200Evaluator const *instance, void *deviceContext = NULL) {
        float *dsts[1] = { dstBuffer->BindCudaBuffer() };
        BufferDescriptor descs[1] = { dstDesc };
        return evalPatchTable(srcBuffer->BindCudaBuffer(), srcDesc, 1, dsts, descs, numPatchCoords,
                           patchCoords->BindCudaBuffer(), B200OSD_PATCH_TRIPLE_FVAR(patchTable, fvarChannel), instance, deviceContext);
    }

    template <typename SRC_BUFFER, typename DST_BUFFER, typename PATCHCOORD_BUFFER, typename PATCH_TABLE>
    static bool EvalPatchesFaceVarying(SRC_BUFFER *srcBuffer, BufferDescriptor const &srcDesc,
                                       DST_BUFFER *dstBuffer, BufferDescriptor const &dstDesc,
                                       DST_BUFFER *duBuffer, BufferDescriptor const &duDesc,
                                       DST_BUFFER *dvBuffer, BufferDescriptor const &dvDesc,
                                       int numPatchCoords, PATCHCOORD_BUFFER *patchCoords, PATCH_TABLE *patchTable,
                                       int fvarChannel, B200Evaluator const *instance, void *deviceContext = NULL) {
        float *dsts[3] = { dstBuffer->BindCudaBuffer(), duBuffer->BindCudaBuffer(), dvBuffer->BindCudaBuffer() };
        BufferDescriptor descs[3] = { dstDesc, duDesc, dvDesc };
        return evalPatchTable(srcBuffer->BindCudaBuffer(), srcDesc, 3, dsts, descs, numPatchCoords,
                           patchCoords->BindCudaBuffer(), B200OSD_PATCH_TRIPLE_FVAR(patchTable, fvarChannel), instance, deviceContext);
    }

    template <typename SRC_BUFFER, typename DST_BUFFER, typename PATCHCOORD_BUFFER, typename PATCH_TABLE>
    static bool EvalPatchesFaceVarying(SRC_BUFFER *srcBuffer, BufferDescriptor const &srcDesc,
                                       DST_BUFFER *dstBuffer, BufferDescriptor const &dstDesc,
                                       DST_BUFFER *duBuffer, BufferDescriptor const &duDesc,
                                       DST_BUFFER *dvBuffer, BufferDescriptor const &dvDesc,
                                       DST_BUFFER *duuBuffer, BufferDescriptor const &duuDesc,
                                       DST_BUFFER *duvBuffer, BufferDescriptor const &duvDesc,
                                       DST_BUFFER *dvvBuffer, BufferDescriptor const &dvvDesc,
                                       int numPatchCoords, PATCHCOORD_BUFFER *patchCoords, PATCH_TABLE *patchTable,
                                       int fvarChannel, B200Evaluator const *instance, void *deviceContext = NULL) {
        float *dsts[6] = { dstBuffer->BindCudaBuffer(), duBuffer->BindCudaBuffer(), dvBuffer->BindCudaBuffer(),
                           duuBuffer->BindCudaBuffer(), duvBuffer->BindCudaBuffer(), dvvBuffer->BindCudaBuffer() };
        BufferDescriptor descs[6] = { dstDesc, duDesc, dvDesc, duuDesc, duvDesc, dvvDesc };
        return evalPatchTable(srcBuffer->BindCudaBuffer(), srcDesc, 6, dsts, descs, numPatchCoords,
                           patchCoords->BindCudaBuffer(), B200OSD_PATCH_TRIPLE_FVAR(patchTable, fvarChannel), instance, deviceContext);
    }

#undef B200OSD_PATCH_TRIPLE_VERTEX
#undef B200OSD_PATCH_TRIPLE_VARYING
#undef B200OSD_PATCH_TRIPLE_FVAR

    /// Waits for all enqueued work (CudaEvaluator::Synchronize, cudaEvaluator.cpp:377-380).
    static void Synchronize(void *deviceContext = NULL) { b200osd_synchronize(B200StreamOf(deviceContext)); }

private:
    static void flatten(int n, BufferDescriptor const *descs, int (*out)[3]) {
        for (int k = 0; k < n; ++k) { out[k][0] = descs[k].offset; out[k][1] = descs[k].length; out[k][2] = descs[k].stride; }
    }

    // fast path: the table owns the bucketed layout
    static bool evalTable(const float *src, BufferDescriptor const &srcDesc, int n, float *const *dsts,
                          BufferDescriptor const *descs, B200StencilTable const *table, void *deviceContext) {
        int sd[3] = { srcDesc.offset, srcDesc.length, srcDesc.stride };
        int dd[6][3];
        flatten(n, descs, dd);
        return b200osd_stencil_table_eval(table->GetHandle(), src, sd, n, dsts, dd, 0, table->GetNumStencils(),
                                          B200StreamOf(deviceContext)) == B200OSD_OK;
    }

    // any other table type with the reference's device-pointer accessors (e.g. Osd::CudaStencilTable)
    template <typename STENCIL_TABLE>
    static bool evalTable(const float *src, BufferDescriptor const &srcDesc, int n, float *const *dsts,
                          BufferDescriptor const *descs, STENCIL_TABLE const *table, void *deviceContext) {
        const float *w[6] = { (const float *)table->GetWeightsBuffer(), NULL, NULL, NULL, NULL, NULL };
        if (n >= 3) { w[1] = (const float *)table->GetDuWeightsBuffer(); w[2] = (const float *)table->GetDvWeightsBuffer(); }
        if (n >= 6) {
            w[3] = (const float *)table->GetDuuWeightsBuffer();
            w[4] = (const float *)table->GetDuvWeightsBuffer();
            w[5] = (const float *)table->GetDvvWeightsBuffer();
        }
        return evalRaw(src, srcDesc, n, dsts, descs, (const int *)table->GetSizesBuffer(), (const int *)table->GetOffsetsBuffer(),
                       (const int *)table->GetIndicesBuffer(), w, 0, table->GetNumStencils(), deviceContext);
    }

    static bool evalRaw(const float *src, BufferDescriptor const &srcDesc, int n, float *const *dsts,
                        BufferDescriptor const *descs, const int *sizes, const int *offsets, const int *indices,
                        const float *const *weights, int start, int end, void *deviceContext) {
        int sd[3] = { srcDesc.offset, srcDesc.length, srcDesc.stride };
        int dd[6][3];
        flatten(n, descs, dd);
        return b200osd_eval_stencils(src, sd, n, dsts, dd, sizes, offsets, indices, weights, start, end,
                                     B200StreamOf(deviceContext)) == B200OSD_OK;
    }

    // fast path: B200PatchTable owns the handle; an instance whose bound coordinate set matches supplies the cached grouping
    static bool evalPatchTable(const float *src, BufferDescriptor const &srcDesc, int n, float *const *dsts,
                               BufferDescriptor const *descs, int numPatchCoords, const void *patchCoords,
                               B200PatchTable const *table, int which, B200Evaluator const *instance, void *deviceContext) {
        int sd[3] = { srcDesc.offset, srcDesc.length, srcDesc.stride };
        int dd[6][3];
        flatten(n, descs, dd);
        if (instance && instance->_plan && instance->_planTable == table->GetHandle() &&
            instance->_planCoords == patchCoords && instance->_planCount == numPatchCoords)
            return b200osd_patch_plan_eval(instance->_plan, which, src, sd, n, dsts, dd, numPatchCoords,
                                           (const b200osd_patch_coord *)patchCoords, B200StreamOf(deviceContext)) == B200OSD_OK;
        return b200osd_patch_table_eval(table->GetHandle(), which, src, sd, n, dsts, dd, numPatchCoords,
                                        (const b200osd_patch_coord *)patchCoords, B200StreamOf(deviceContext)) == B200OSD_OK;
    }
    static bool evalPatchTable(const float *src, BufferDescriptor const &srcDesc, int n, float *const *dsts,
                               BufferDescriptor const *descs, int numPatchCoords, const void *patchCoords,
                               B200PatchTable *table, int which, B200Evaluator const *instance, void *deviceContext) {
        return evalPatchTable(src, srcDesc, n, dsts, descs, numPatchCoords, patchCoords,
                              static_cast<B200PatchTable const *>(table), which, instance, deviceContext);
    }

    // any other patch-table type with the reference's device-pointer accessors (e.g. Osd::CudaPatchTable)
    template <typename PATCH_TABLE>
    static bool evalPatchTable(const float *src, BufferDescriptor const &srcDesc, int n, float *const *dsts,
                               BufferDescriptor const *descs, int numPatchCoords, const void *patchCoords,
                               PATCH_TABLE *table, int which, B200Evaluator const *instance, void *deviceContext) {
        (void)instance;
        if (which == 0)
            return evalPatches(src, srcDesc, n, dsts, descs, numPatchCoords, patchCoords, table->GetPatchArrayBuffer(),
                               table->GetPatchIndexBuffer(), table->GetPatchParamBuffer(), deviceContext);
        if (which == 1)
            return evalPatches(src, srcDesc, n, dsts, descs, numPatchCoords, patchCoords, table->GetVaryingPatchArrayBuffer(),
                               table->GetVaryingPatchIndexBuffer(), table->GetPatchParamBuffer(), deviceContext);
        return evalPatches(src, srcDesc, n, dsts, descs, numPatchCoords, patchCoords, table->GetFVarPatchArrayBuffer(which - 2),
                           table->GetFVarPatchIndexBuffer(which - 2), table->GetFVarPatchParamBuffer(which - 2), deviceContext);
    }

    static bool evalPatches(const float *src, BufferDescriptor const &srcDesc, int n, float *const *dsts,
                            BufferDescriptor const *descs, int numPatchCoords, const void *patchCoords,
                            const void *patchArrays, const void *patchIndices, const void *patchParams, void *deviceContext) {
        int sd[3] = { srcDesc.offset, srcDesc.length, srcDesc.stride };
        int dd[6][3];
        flatten(n, descs, dd);
        // a client built with OPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES keeps that behaviour (osd/patchBasis.h:421-487)
#ifdef OPENSUBDIV_GREGORY_EVAL_TRUE_DERIVATIVES
        const int options = B200OSD_PATCH_GREGORY_TRUE_DERIVATIVES;
#else
        const int options = 0;
#endif
        return b200osd_eval_patches_ex(src, sd, n, dsts, dd, numPatchCoords, (const b200osd_patch_coord *)patchCoords,
                                       (const b200osd_patch_array *)patchArrays, (const int *)patchIndices,
                                       (const b200osd_patch_param *)patchParams, options,
                                       B200StreamOf(deviceContext)) == B200OSD_OK;
    }

    B200Evaluator() : _plan(NULL), _planTable(NULL), _planCoords(NULL), _planCount(0) {}
    B200Evaluator(B200Evaluator const &);
    B200Evaluator &operator=(B200Evaluator const &);
    b200osd_patch_plan *_plan;
    b200osd_patch_table const *_planTable;
    const void *_planCoords;
    int _planCount;
};

}  // namespace Osd
}  // namespace OPENSUBDIV_VERSION
using namespace OPENSUBDIV_VERSION;
}  // namespace OpenSubdiv

#endif
