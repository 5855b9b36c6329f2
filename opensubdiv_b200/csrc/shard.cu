// shard.cu -- the multi-GPU data plane of the evaluator, in C: work partitioning and the per-frame
// replication of the control points over NVLink.
//
// The reference has no multi-GPU code at all (no NCCL / MPI / cudaSetDevice anywhere in
// opensubdiv/); what it does offer is the hook this layer builds on: the raw EvalStencils overloads take an absolute
// row range [start, end) (osd/cudaEvaluator.h:171-178, osd/cudaKernel.cu:85-98).  The path shards naturally -- every
// stencil row and every PatchCoord is independent and the tables are static -- so
//   * b200osd_shard_plan cuts the rows into `world` contiguous ranges of equal COST (elements + a constant per row);
//     each rank builds a table of its rows only (offsets re-based) or passes its range to the table evaluation;
//   * b200osd_shard_coords cuts a PatchCoord set into contiguous ranges (every coordinate costs the same);
//   * b200osd_comm_* replicates the one mutable input, the deformed control points of a frame, to every GPU: one process
//     per GPU, NCCL over NVLink 5 / NVSwitch.  broadcast = every rank gets everything (one mesh, rows sharded: "strong");
//     scatter = rank r gets only slice r (N meshes, rank r owns mesh r: "weak" -- 1/N of the broadcast's bytes per GPU).
//   No reduction exists anywhere: no row spans ranks.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- inside a PyTorch process that is the copy torch already loaded),
// so libb200osd.so itself has no link-time dependency on it and single-GPU users never touch it.
#include "common.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

using namespace b200osd;

// nccl.h supplies the types (and the versioned ncclConfig_t); every FUNCTION is bound with dlsym below
#if __has_include(<nccl.h>)
#include <nccl.h>
#define B200OSD_HAVE_NCCL_CONFIG 1
#else
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
#endif

namespace {

enum { kNcclSuccess = 0, kNcclFloat = 7 };

struct NcclApi {
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
#ifdef B200OSD_HAVE_NCCL_CONFIG
    int (*CommInitRankConfig)(ncclComm_t *, int, ncclUniqueId, int, ncclConfig_t *) = nullptr;
#endif
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

NcclApi *nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
#define B200_NCCL_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name))
        B200_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        B200_NCCL_SYM(CommInitRank, "ncclCommInitRank");
#ifdef B200OSD_HAVE_NCCL_CONFIG
        B200_NCCL_SYM(CommInitRankConfig, "ncclCommInitRankConfig");
#endif
        B200_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        B200_NCCL_SYM(Broadcast, "ncclBroadcast");
        B200_NCCL_SYM(AllGather, "ncclAllGather");
        B200_NCCL_SYM(Send, "ncclSend");
        B200_NCCL_SYM(Recv, "ncclRecv");
        B200_NCCL_SYM(GroupStart, "ncclGroupStart");
        B200_NCCL_SYM(GroupEnd, "ncclGroupEnd");
        B200_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef B200_NCCL_SYM
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.AllGather && api.Send &&
                 api.Recv && api.GroupStart && api.GroupEnd;
    });
    return &api;
}

int nccl_check(int r, const char *what) {
    if (r == kNcclSuccess) return B200OSD_OK;
    NcclApi *a = nccl();
    set_error("%s failed: %s", what, a->GetErrorString ? a->GetErrorString(r) : "NCCL error");
    return B200OSD_ERR_CUDA;
}

}  // namespace

struct b200osd_comm {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
};

extern "C" {

// ------------------------------------------------------------------------------- partitioning --
int b200osd_shard_plan(int numStencils, const int *sizes, int world, int align, int *ranges) {
    if (numStencils < 0 || world < 1 || !ranges || (numStencils > 0 && !sizes)) { set_error("shard_plan: bad arguments"); return B200OSD_ERR_INVALID; }
    const int n = numStencils;
    // cost of a row = its elements + 1 (descriptor + output); cut where the running cost passes r/world of the total
    long long total = 0;
    for (int i = 0; i < n; ++i) total += (long long)sizes[i] + 1;
    std::vector<int> cuts((size_t)world + 1, n);
    cuts[0] = 0;
    long long run = 0;
    int i = 0;
    for (int r = 1; r < world; ++r) {
        const long long target = total * r / world;
        while (i < n && run + sizes[i] + 1 < target) { run += (long long)sizes[i] + 1; ++i; }
        int c = std::min(i + 1, n);                      // first row whose inclusive running cost reaches the target, + 1
        if (align > 1) c = (int)(((long long)c + align / 2) / align) * align;
        c = std::min(std::max(c, cuts[(size_t)r - 1]), n);
        cuts[(size_t)r] = c;
    }
    for (int r = 0; r < world; ++r) { ranges[2 * r] = cuts[(size_t)r]; ranges[2 * r + 1] = cuts[(size_t)r + 1]; }
    return B200OSD_OK;
}

int b200osd_shard_plan_locality(int numStencils, const int *sizes, const int *offsets, const int *indices, int world,
                                int *rowOrder, int *ranges, int *controlRanges) {
    if (numStencils < 0 || world < 1 || !ranges || !controlRanges || (numStencils > 0 && (!sizes || !offsets || !rowOrder))) {
        set_error("shard_plan_locality: bad arguments");
        return B200OSD_ERR_INVALID;
    }
    const int n = numStencils;
    if (!indices)                                        // legal only for a table without elements
        for (int i = 0; i < n; ++i)
            if (sizes[i] > 0) { set_error("shard_plan_locality: indices is NULL"); return B200OSD_ERR_INVALID; }
    // key of a row: the smallest control vertex it references (rows of an empty stencil go first)
    std::vector<int> key((size_t)n, -1);
    int maxKey = -1;
    for (int i = 0; i < n; ++i) {
        if (sizes[i] < 0 || offsets[i] < 0) { set_error("shard_plan_locality: negative size / offset in row %d", i); return B200OSD_ERR_INVALID; }
        int k = -1;
        for (int j = 0; j < sizes[i]; ++j) {
            const int v = indices[(size_t)offsets[i] + j];
            if (v < 0) { set_error("shard_plan_locality: negative index in row %d", i); return B200OSD_ERR_INVALID; }
            k = (k < 0 || v < k) ? v : k;
        }
        key[(size_t)i] = k;
        maxKey = std::max(maxKey, k);
    }
    // counting sort by key, stable: rows of equal key keep the table's order
    std::vector<long long> start((size_t)maxKey + 3, 0);
    for (int i = 0; i < n; ++i) ++start[(size_t)(key[(size_t)i] + 1) + 1];
    for (size_t k = 1; k < start.size(); ++k) start[k] += start[k - 1];
    for (int i = 0; i < n; ++i) rowOrder[start[(size_t)(key[(size_t)i] + 1)]++] = i;
    // cuts balanced on cost (elements + 1 per row), like b200osd_shard_plan
    long long total = 0;
    for (int i = 0; i < n; ++i) total += (long long)sizes[i] + 1;
    std::vector<int> cuts((size_t)world + 1, n);
    cuts[0] = 0;
    long long run = 0;
    int i = 0;
    for (int r = 1; r < world; ++r) {
        const long long target = total * r / world;
        while (i < n && run + sizes[rowOrder[i]] + 1 < target) { run += (long long)sizes[rowOrder[i]] + 1; ++i; }
        cuts[(size_t)r] = std::min(std::max(std::min(i + 1, n), cuts[(size_t)r - 1]), n);
    }
    for (int r = 0; r < world; ++r) {
        ranges[2 * r] = cuts[(size_t)r];
        ranges[2 * r + 1] = cuts[(size_t)r + 1];
        int lo = 0x7fffffff, hi = -1;
        for (int q = cuts[(size_t)r]; q < cuts[(size_t)r + 1]; ++q) {
            const int row = rowOrder[q];
            for (int j = 0; j < sizes[row]; ++j) {
                const int v = indices[(size_t)offsets[row] + j];
                lo = std::min(lo, v);
                hi = std::max(hi, v);
            }
        }
        controlRanges[2 * r] = hi < 0 ? 0 : lo;
        controlRanges[2 * r + 1] = hi + 1;
    }
    return B200OSD_OK;
}

int b200osd_shard_control_runs(int numStencils, const int *sizes, const int *offsets, const int *indices, int granularity,
                               int maxRuns, int *runs) {
    if (numStencils < 0 || granularity < 1 || maxRuns < 1 || !runs || (numStencils > 0 && (!sizes || !offsets))) {
        set_error("shard_control_runs: bad arguments");
        return -1;
    }
    if (!indices)                                        // legal only for a table without elements
        for (int i = 0; i < numStencils; ++i)
            if (sizes[i] > 0) { set_error("shard_control_runs: indices is NULL"); return -1; }
    int maxIdx = -1;
    for (int i = 0; i < numStencils; ++i)
        for (int j = 0; j < sizes[i]; ++j) maxIdx = std::max(maxIdx, indices[(size_t)offsets[i] + j]);
    if (maxIdx < 0) return 0;
    const int blocks = maxIdx / granularity + 1;
    std::vector<char> used((size_t)blocks, 0);
    for (int i = 0; i < numStencils; ++i)
        for (int j = 0; j < sizes[i]; ++j) {
            const int v = indices[(size_t)offsets[i] + j];
            if (v < 0) { set_error("shard_control_runs: negative index in row %d", i); return -1; }
            used[(size_t)(v / granularity)] = 1;
        }
    std::vector<std::pair<int, int>> r;                  // [first block, last block + 1)
    for (int b = 0; b < blocks; ++b) {
        if (!used[(size_t)b]) continue;
        if (!r.empty() && r.back().second == b) r.back().second = b + 1;
        else r.push_back(std::make_pair(b, b + 1));
    }
    while ((int)r.size() > maxRuns) {                    // too many pieces: close the smallest gap
        size_t best = 0;
        for (size_t k = 1; k + 1 < r.size(); ++k)
            if (r[k + 1].first - r[k].second < r[best + 1].first - r[best].second) best = k;
        r[best].second = r[best + 1].second;
        r.erase(r.begin() + (long)best + 1);
    }
    for (size_t k = 0; k < r.size(); ++k) {
        runs[2 * k] = r[k].first * granularity;
        runs[2 * k + 1] = std::min(r[k].second * granularity, maxIdx + 1);
    }
    return (int)r.size();
}

int b200osd_shard_coords(long long numCoords, int world, int align, long long *ranges) {
    if (numCoords < 0 || world < 1 || !ranges) { set_error("shard_coords: bad arguments"); return B200OSD_ERR_INVALID; }
    long long prev = 0;
    for (int r = 0; r < world; ++r) {
        long long c = r + 1 == world ? numCoords : numCoords * (r + 1) / world;
        if (align > 1 && r + 1 < world) c = (c + align / 2) / align * align;
        c = std::min(std::max(c, prev), numCoords);
        ranges[2 * r] = prev;
        ranges[2 * r + 1] = c;
        prev = c;
    }
    return B200OSD_OK;
}

// ------------------------------------------------------------------------------ communicator --
int b200osd_comm_available(void) { return nccl()->ok ? 1 : 0; }

int b200osd_comm_unique_id(char id[128]) {
    NcclApi *a = nccl();
    if (!a->ok) { set_error("NCCL is not available (libnccl.so.2 could not be loaded)"); return B200OSD_ERR_UNSUPPORTED; }
    if (!id) { set_error("comm_unique_id: id is NULL"); return B200OSD_ERR_INVALID; }
    ncclUniqueId u;
    int rc = nccl_check(a->GetUniqueId(&u), "ncclGetUniqueId");
    if (rc) return rc;
    std::memcpy(id, u.internal, 128);
    return B200OSD_OK;
}

b200osd_comm *b200osd_comm_create(int world, int rank, const char id[128]) {
    return b200osd_comm_create_ex(world, rank, id, 0);
}

b200osd_comm *b200osd_comm_create_ex(int world, int rank, const char id[128], int maxCTAs) {
    NcclApi *a = nccl();
    if (!a->ok) { set_error("NCCL is not available (libnccl.so.2 could not be loaded)"); return nullptr; }
    if (world < 1 || rank < 0 || rank >= world || !id) { set_error("comm_create: bad world / rank / id"); return nullptr; }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError())); return nullptr; }
    b200osd_comm *c = new (std::nothrow) b200osd_comm;
    if (!c) return nullptr;
    c->world = world;
    c->rank = rank;
    ncclUniqueId u;
    std::memcpy(u.internal, id, 128);
    int rc = -1;
#ifdef B200OSD_HAVE_NCCL_CONFIG
    if (maxCTAs > 0 && a->CommInitRankConfig) {
        // a transfer of a few MB hidden behind an evaluation kernel needs no bandwidth, only few SMs: cap NCCL's CTAs
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.minCTAs = 1;
        cfg.maxCTAs = maxCTAs;
        rc = a->CommInitRankConfig(&c->comm, world, u, rank, &cfg);
        if (rc != kNcclSuccess) c->comm = nullptr;                 // e.g. a runtime that rejects the header's config version
    }
#endif
    (void)maxCTAs;
    if (rc != kNcclSuccess && nccl_check(a->CommInitRank(&c->comm, world, u, rank), "ncclCommInitRank")) { delete c; return nullptr; }
    return c;
}

void b200osd_comm_destroy(b200osd_comm *c) {
    if (!c) return;
    if (c->comm) nccl()->CommDestroy(c->comm);
    delete c;
}

int b200osd_comm_world(const b200osd_comm *c) { return c ? c->world : 1; }
int b200osd_comm_rank(const b200osd_comm *c) { return c ? c->rank : 0; }

int b200osd_comm_broadcast(b200osd_comm *c, float *buf, size_t count, int root, void *stream) {
    if (!c || !buf) { set_error("comm_broadcast: NULL communicator / buffer"); return B200OSD_ERR_INVALID; }
    if (count == 0 || c->world == 1) return B200OSD_OK;
    return nccl_check(nccl()->Broadcast(buf, buf, count, kNcclFloat, root, c->comm, (cudaStream_t)stream), "ncclBroadcast");
}

int b200osd_comm_scatter(b200osd_comm *c, const float *sendbuf, float *recvbuf, size_t countPerRank, int root, void *stream) {
    if (!c || !recvbuf || (c->rank == root && !sendbuf)) { set_error("comm_scatter: NULL communicator / buffer"); return B200OSD_ERR_INVALID; }
    if (countPerRank == 0) return B200OSD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    NcclApi *a = nccl();
    if (c->world == 1) {
        if (sendbuf != recvbuf) B200_CUDA_TRY(cudaMemcpyAsync(recvbuf, sendbuf, countPerRank * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return B200OSD_OK;
    }
    int rc = nccl_check(a->GroupStart(), "ncclGroupStart");
    if (rc) return rc;
    if (c->rank == root) {
        for (int r = 0; r < c->world && !rc; ++r) {
            if (r == root) continue;
            rc = nccl_check(a->Send(sendbuf + (size_t)r * countPerRank, countPerRank, kNcclFloat, r, c->comm, st), "ncclSend");
        }
    } else {
        rc = nccl_check(a->Recv(recvbuf, countPerRank, kNcclFloat, root, c->comm, st), "ncclRecv");
    }
    const int rc2 = nccl_check(a->GroupEnd(), "ncclGroupEnd");
    if (rc || rc2) return rc ? rc : rc2;
    if (c->rank == root && sendbuf + (size_t)root * countPerRank != recvbuf)
        B200_CUDA_TRY(cudaMemcpyAsync(recvbuf, sendbuf + (size_t)root * countPerRank, countPerRank * sizeof(float),
                                      cudaMemcpyDeviceToDevice, st));
    return B200OSD_OK;
}

int b200osd_comm_all_gather(b200osd_comm *c, const float *sendbuf, float *recvbuf, size_t countPerRank, void *stream) {
    if (!c || !sendbuf || !recvbuf) { set_error("comm_all_gather: NULL communicator / buffer"); return B200OSD_ERR_INVALID; }
    if (countPerRank == 0) return B200OSD_OK;
    if (c->world == 1) {
        if (sendbuf != recvbuf) B200_CUDA_TRY(cudaMemcpyAsync(recvbuf, sendbuf, countPerRank * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return B200OSD_OK;
    }
    return nccl_check(nccl()->AllGather(sendbuf, recvbuf, countPerRank, kNcclFloat, c->comm, (cudaStream_t)stream), "ncclAllGather");
}

}  // extern "C"

// --------------------------------------------------------------------------- peer-memory window --
// One-sided exchange over NVLink peer memory, without NCCL kernels: every rank owns a device buffer that all other ranks
// of the box can address (CUDA IPC), data moves by DMA (copy engines: cudaMemcpyAsync between peer-mapped pointers) and
// ordering is carried by per-(slot, sender) counters in peer memory written / polled by ONE-thread kernels.  An
// evaluation kernel that runs next to such an exchange keeps all of its SMs -- the NCCL path costs it the 8-16 thread
// blocks of the collective's kernel for the duration of the transfer.
namespace {

constexpr int kWindowSlots = 16;

// after everything issued so far on the stream: bump my counter for (slot -> peer p) and store it into p's flag array
__global__ void window_signal_kernel(int *const *peerFlags, int *sendSeq, int world, int rank, int slot, int dst) {
    const int p = threadIdx.x;
    if (p >= world || p == rank || (dst >= 0 && p != dst)) return;
    const int v = sendSeq[slot * world + p] + 1;
    sendSeq[slot * world + p] = v;
    __threadfence_system();
    *reinterpret_cast<volatile int *>(peerFlags[p] + slot * world + rank) = v;
}

// later work on the stream waits until the next signal of (slot, sender p) has arrived; gives up after 10 s (sets *error)
__global__ void window_wait_kernel(const int *flags, int *expect, int world, int rank, int slot, int src, int *error) {
    const int p = threadIdx.x;
    if (p >= world || p == rank || (src >= 0 && p != src)) return;
    const int want = expect[slot * world + p] + 1;
    expect[slot * world + p] = want;
    const volatile int *f = flags + slot * world + p;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*f < want) {
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) { *error = 1; return; }
    }
    __threadfence_system();
}

// wait + copy + signal in ONE kernel: the copy reads the source rank's window with ordinary (uncached) loads over NVLink
// peer memory instead of going through a copy engine, which takes the DMA set-up latency and two kernel launches off a
// per-frame exchange of a few hundred KB.  Every block waits for the flag on its own (no block waits for another one, so
// the grid needs no co-residency); the last block to finish advances the wait counter and signals.
constexpr int kPullMaxRuns = 8, kPullBlocks = 32, kPullThreads = 256, kPullUnroll = 8;
struct PullArgs {
    const char *src[kPullMaxRuns];
    char *dst[kPullMaxRuns];
    unsigned long long bytes[kPullMaxRuns];
    int n;
};

__global__ void __launch_bounds__(kPullThreads) window_pull_kernel(PullArgs a, const int *flags, int *expect, int *const *peerFlags,
                                                                  int *sendSeq, int *error, int *done, int world, int rank,
                                                                  int srcRank, int waitSlot, int sigRank, int sigSlot) {
    if (waitSlot >= 0) {
        if (threadIdx.x == 0) {
            const int want = expect[waitSlot * world + srcRank] + 1;      // advanced by the last block, after every block read it
            const volatile int *f = flags + waitSlot * world + srcRank;
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while (*f < want) {
                __nanosleep(100);
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 10000000000ull) { *error = 1; break; }
            }
            __threadfence_system();
        }
        __syncthreads();
    }
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < a.n; ++r) {
        const bool vec = ((reinterpret_cast<uintptr_t>(a.src[r]) | reinterpret_cast<uintptr_t>(a.dst[r])) & 15) == 0;
        const size_t nv = vec ? a.bytes[r] / 16 : 0;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(a.src[r]);
        uint4 *d4 = reinterpret_cast<uint4 *>(a.dst[r]);
        // kPullUnroll independent 16-byte loads in flight per thread: a load over NVLink takes ~2 us, so the bytes in
        // flight (blocks x threads x unroll x 16 = 1 MB) decide the copy rate, not the instruction count
        for (size_t i = tid; i < nv; i += nth * kPullUnroll) {
            uint4 v[kPullUnroll];
#pragma unroll
            for (int u = 0; u < kPullUnroll; ++u)
                if (i + u * nth < nv) v[u] = __ldcv(s4 + i + u * nth);
#pragma unroll
            for (int u = 0; u < kPullUnroll; ++u)
                if (i + u * nth < nv) d4[i + u * nth] = v[u];
        }
        const unsigned *s1 = reinterpret_cast<const unsigned *>(a.src[r] + nv * 16);
        unsigned *d1 = reinterpret_cast<unsigned *>(a.dst[r] + nv * 16);
        const size_t nw = (a.bytes[r] - nv * 16) / 4;
        for (size_t i = tid; i < nw; i += nth) d1[i] = __ldcv(s1 + i);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(done, 1) == (int)gridDim.x - 1) {                  // the last block
            *done = 0;
            if (waitSlot >= 0) expect[waitSlot * world + srcRank] += 1;
            if (sigSlot >= 0) {
                for (int p = 0; p < world; ++p) {
                    if (p == rank || (sigRank >= 0 && p != sigRank)) continue;
                    const int v = sendSeq[sigSlot * world + p] + 1;
                    sendSeq[sigSlot * world + p] = v;
                    __threadfence_system();
                    *reinterpret_cast<volatile int *>(peerFlags[p] + sigSlot * world + rank) = v;
                }
            }
        }
    }
}

}  // namespace

struct b200osd_window {
    b200osd_comm *comm = nullptr;
    int world = 1, rank = 0;
    size_t bytes = 0;
    void *block = nullptr;                 // one cudaMalloc: [ data (bytes, 256-aligned) | flags | sendSeq | expect | error | peerFlags table ]
    char *data = nullptr;
    int *flags = nullptr, *sendSeq = nullptr, *expect = nullptr, *error = nullptr;
    int **d_peerFlags = nullptr;
    std::vector<char *> peerData;          // every rank's data region as mapped here (own: local pointer)
    std::vector<void *> opened;            // IPC mappings to close
};

extern "C" {

b200osd_window *b200osd_window_create(b200osd_comm *c, size_t bytes) {
    if (!c || bytes == 0) { set_error("window_create: NULL communicator / empty window"); return nullptr; }
    NcclApi *a = nccl();
    b200osd_window *w = new (std::nothrow) b200osd_window;
    if (!w) return nullptr;
    w->comm = c;
    w->world = c->world;
    w->rank = c->rank;
    w->bytes = bytes;
    const size_t dataBytes = (bytes + 255) & ~(size_t)255;
    const size_t ctr = (size_t)kWindowSlots * w->world * sizeof(int);
    const size_t total = dataBytes + 3 * ((ctr + 255) & ~(size_t)255) + 256 + (((size_t)w->world * sizeof(int *) + 255) & ~(size_t)255);
    cudaError_t e = cudaMalloc(&w->block, total);
    if (e != cudaSuccess) { set_error("window_create: cudaMalloc(%zu) failed: %s", total, cudaGetErrorString(e)); delete w; return nullptr; }
    cudaMemset(w->block, 0, total);
    char *p = static_cast<char *>(w->block);
    w->data = p;                                         p += dataBytes;
    w->flags = reinterpret_cast<int *>(p);               p += (ctr + 255) & ~(size_t)255;
    w->sendSeq = reinterpret_cast<int *>(p);             p += (ctr + 255) & ~(size_t)255;
    w->expect = reinterpret_cast<int *>(p);              p += (ctr + 255) & ~(size_t)255;
    w->error = reinterpret_cast<int *>(p);               p += 256;
    w->d_peerFlags = reinterpret_cast<int **>(p);
    w->peerData.assign((size_t)w->world, nullptr);
    std::vector<int *> peerFlags((size_t)w->world, nullptr);
    w->peerData[(size_t)w->rank] = w->data;
    peerFlags[(size_t)w->rank] = w->flags;
    bool ok = true;
    if (w->world > 1) {
        // every rank's IPC handle of its block travels through one all-gather; offsets inside the block are identical
        cudaIpcMemHandle_t mine;
        ok = cudaIpcGetMemHandle(&mine, w->block) == cudaSuccess;
        const size_t hb = sizeof(cudaIpcMemHandle_t);
        const size_t slotFloats = (hb + 3) / 4;
        float *dsend = nullptr, *drecv = nullptr;
        ok = ok && cudaMalloc((void **)&dsend, slotFloats * 4) == cudaSuccess && cudaMalloc((void **)&drecv, slotFloats * 4 * w->world) == cudaSuccess;
        std::vector<cudaIpcMemHandle_t> all((size_t)w->world);
        if (ok) {
            cudaMemcpy(dsend, &mine, hb, cudaMemcpyHostToDevice);
            ok = a->AllGather(dsend, drecv, slotFloats, kNcclFloat, c->comm, (cudaStream_t)0) == kNcclSuccess &&
                 cudaStreamSynchronize(0) == cudaSuccess;
            for (int r = 0; ok && r < w->world; ++r)
                ok = cudaMemcpy(&all[(size_t)r], reinterpret_cast<char *>(drecv) + (size_t)r * slotFloats * 4, hb, cudaMemcpyDeviceToHost) == cudaSuccess;
        }
        cudaFree(dsend);
        cudaFree(drecv);
        for (int r = 0; ok && r < w->world; ++r) {
            if (r == w->rank) continue;
            void *base = nullptr;
            if (cudaIpcOpenMemHandle(&base, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
            w->opened.push_back(base);
            w->peerData[(size_t)r] = static_cast<char *>(base);
            peerFlags[(size_t)r] = reinterpret_cast<int *>(static_cast<char *>(base) + dataBytes);
        }
    }
    if (ok) ok = cudaMemcpy(w->d_peerFlags, peerFlags.data(), (size_t)w->world * sizeof(int *), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        set_error("window_create: exchanging / opening the peer mappings failed: %s", cudaGetErrorString(cudaGetLastError()));
        b200osd_window_destroy(w);
        return nullptr;
    }
    return w;
}

void b200osd_window_destroy(b200osd_window *w) {
    if (!w) return;
    for (void *p : w->opened) cudaIpcCloseMemHandle(p);
    cudaFree(w->block);
    delete w;
}

void *b200osd_window_local(const b200osd_window *w) { return w ? (void *)w->data : nullptr; }
size_t b200osd_window_bytes(const b200osd_window *w) { return w ? w->bytes : 0; }

int b200osd_window_get(b200osd_window *w, int srcRank, size_t srcOffsetBytes, void *dst, size_t bytes, void *stream) {
    if (!w || !dst || srcRank < 0 || srcRank >= w->world) { set_error("window_get: bad window / rank / destination"); return B200OSD_ERR_INVALID; }
    if (srcOffsetBytes + bytes > w->bytes) { set_error("window_get: [%zu,+%zu) outside the %zu-byte window", srcOffsetBytes, bytes, w->bytes); return B200OSD_ERR_INVALID; }
    if (bytes == 0) return B200OSD_OK;
    B200_CUDA_TRY(cudaMemcpyAsync(dst, w->peerData[(size_t)srcRank] + srcOffsetBytes, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return B200OSD_OK;
}

int b200osd_window_pull(b200osd_window *w, int srcRank, int waitSlot, int numRuns, const size_t *srcOffsetBytes, void *const *dsts,
                        const size_t *bytes, int signalRank, int signalSlot, void *stream) {
    if (!w || srcRank < 0 || srcRank >= w->world || numRuns < 0 || numRuns > kPullMaxRuns || waitSlot >= kWindowSlots ||
        signalSlot >= kWindowSlots || signalRank >= w->world || (numRuns > 0 && (!srcOffsetBytes || !dsts || !bytes))) {
        set_error("window_pull: bad window / rank / slot / runs (at most %d runs)", kPullMaxRuns);
        return B200OSD_ERR_INVALID;
    }
    PullArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n = numRuns;
    for (int r = 0; r < numRuns; ++r) {
        if (!dsts[r] || bytes[r] % 4 != 0 || srcOffsetBytes[r] % 4 != 0 || srcOffsetBytes[r] + bytes[r] > w->bytes) {
            set_error("window_pull: run %d: [%zu,+%zu) outside the %zu-byte window, NULL destination or not a multiple of 4 bytes",
                      r, srcOffsetBytes[r], bytes[r], w->bytes);
            return B200OSD_ERR_INVALID;
        }
        a.src[r] = w->peerData[(size_t)srcRank] + srcOffsetBytes[r];
        a.dst[r] = static_cast<char *>(dsts[r]);
        a.bytes[r] = bytes[r];
    }
    if (w->world == 1) { waitSlot = -1; signalSlot = -1; }
    window_pull_kernel<<<kPullBlocks, kPullThreads, 0, (cudaStream_t)stream>>>(a, w->flags, w->expect, w->d_peerFlags, w->sendSeq, w->error,
                                                                            w->error + 1, w->world, w->rank, srcRank, waitSlot,
                                                                            signalRank, signalSlot);
    return check_launch("window_pull_kernel");
}

int b200osd_window_signal(b200osd_window *w, int dstRank, int slot, void *stream) {
    if (!w || slot < 0 || slot >= kWindowSlots || dstRank >= w->world) { set_error("window_signal: bad window / slot / rank"); return B200OSD_ERR_INVALID; }
    if (w->world == 1) return B200OSD_OK;
    window_signal_kernel<<<1, 32 * ((w->world + 31) / 32), 0, (cudaStream_t)stream>>>(w->d_peerFlags, w->sendSeq, w->world, w->rank, slot, dstRank);
    return check_launch("window_signal_kernel");
}

int b200osd_window_wait(b200osd_window *w, int srcRank, int slot, void *stream) {
    if (!w || slot < 0 || slot >= kWindowSlots || srcRank >= w->world) { set_error("window_wait: bad window / slot / rank"); return B200OSD_ERR_INVALID; }
    if (w->world == 1) return B200OSD_OK;
    window_wait_kernel<<<1, 32 * ((w->world + 31) / 32), 0, (cudaStream_t)stream>>>(w->flags, w->expect, w->world, w->rank, slot, srcRank, w->error);
    return check_launch("window_wait_kernel");
}

int b200osd_window_error(b200osd_window *w) {
    if (!w) return 0;
    int e = 0;
    if (cudaMemcpy(&e, w->error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return e;
}

}  // extern "C"
