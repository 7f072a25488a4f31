// common.cuh -- context, device buffers and launch helpers of libamr_b200.so
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <unordered_map>
#include <atomic>
#include <thread>

#include "amr_b200.h"

typedef int32_t label;
typedef uint8_t u8;

#define AMR_API extern "C" __attribute__((visibility("default")))

constexpr int kThreads = 256;           // CTA size of the cell/point kernels (8 warps)
constexpr int kSlice = 32;              // SELL slice height = one warp
constexpr double kSMALL = 1e-15;
constexpr double kGREAT = 1e15;
constexpr int kMaxIter = 62;            // queue iterations of a closure loop enqueued before the host looks (speculation cap)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct Patch {
    int32_t start, size, type, nbr;
};

// Sliced-ELL (SELL-32) adjacency: rows of one warp-slice are stored column-major so that
// column j of 32 consecutive rows is one 128-byte line.  Entries beyond a row's length hold
// `pad` (the row's own index for cell->cell, -1 for point->cell).
struct Sell {
    int64_t rows = 0;
    int64_t slices = 0;
    int64_t entries = 0;       // 32 * sum(width)
    int32_t *sliceOff = nullptr;  // [slices+1], exclusive prefix of the slice widths
    int32_t *col = nullptr;       // [entries]
    u8 *count = nullptr;          // [rows] true row length (saturated at 255)
    int maxTileInts = 0;          // largest tile, in int32 (ring stage capacity of the TMA pipeline)
    int32_t *tileFirst = nullptr; // [tiles + 1] first slice of every tile: up to 8 consecutive slices whose columns fit kTileCapCols
    int64_t tiles = 0;
    u8 *sliceGhost = nullptr;     // [slices] slice holds a cell that owns a coupled face (the estimators keep those raw until
                                  // the boundary fix-up has folded the patch faces in); cell->cell only
};

struct TraceRec { const char *name; cudaEvent_t ev; };

// Pinned host memory that the last kernel of a speculative loop writes directly (zero-copy): queue sizes, list size,
// by-product counts and the peer error word arrive with ONE store sequence over PCIe instead of four small DMA copies
// and a stream synchronisation; the host polls `seq`.
struct Mailbox {
    int q[64];
    long long marked[16];
    int total;
    int peerError;
    volatile unsigned seq;
    unsigned pad;
};

struct amr_ctx {
    int device = 0;
    // AMR_B200_TRACE=<file prefix>: a CUDA event after every launch / host round trip of this context; the time line
    // (name, microseconds since the first event) is written to <prefix>.<rank>.<ctx address> at amr_ctx_destroy.
    // Diagnosis only: the events themselves cost ~1 us each.
    bool faceBatch = false;    // Lohner / Gauss gradient: faces of a cell four at a time (lohner.cuh); AMR_B200_FACE_BATCH=1 selects it: measured slower (151 registers), profiles/r2_lohner_timing.json
    bool hostTime = false;     // AMR_B200_HOSTTIME=1: wall-clock marks of the mesh hand-over calls on stderr (diagnosis)
    double htLast = 0.0;
    bool traceOn = false;
    std::string tracePath;
    std::vector<TraceRec> trace;
    cudaStream_t stream = nullptr;
    cudaStream_t upStream = nullptr, downStream = nullptr;   // host<->host field mapping: both PCIe directions at once
    struct UploadLane { void *pin[2] = {nullptr, nullptr}; cudaStream_t s = nullptr; cudaEvent_t ev[2] = {nullptr, nullptr}; };
    std::vector<UploadLane> lanes;   // pageable -> pinned -> device staging lanes (bulkUpload)
    bool ownStream = true;
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    // NVLink peer-memory halos (p2p.cuh)
    bool usePeer = false;
    char *region = nullptr;            // my IPC-exported region
    std::vector<char *> peerHost;      // mapped regions of all ranks (mine included)
    char **peerDev = nullptr;          // same, on the device
    unsigned *dSeq = nullptr;          // device: [0] last completed halo swap, [1] last completed reduction (p2p.cuh)
    unsigned long long peerTimeoutNs = 20000000000ull;   // bound of the device-side spins (AMR_B200_PEER_TIMEOUT_S)
    bool peerTouched = false;          // a peer-memory exchange was enqueued since the error word was last read
    void *haloPlan = nullptr;          // HaloPlan (host copy), rebuilt at amr_mesh_set
    bool havePlan = false;
    void *pointPlanPtr = nullptr;      // HaloPlan of the point swaps (syncPointList), rebuilt at amr_mesh_set_patch_points
    bool havePointPlan = false;
    unsigned long long *redVal = nullptr;   // device staging word for reductions
    std::string err;
    int numSMs = 148;
    int pipeStagesForced = 0;                      // AMR_B200_PIPE_STAGES
    bool usePipeline = true;                       // TMA-staged SELL kernels (AMR_B200_NO_PIPELINE=1 disables)
    std::unordered_map<const void *, int> occCache; // kernel -> resident CTAs per SM at the current smem size
    int64_t launches = 0;
    int64_t devBytes = 0;

    // mesh
    int64_t n = 0, f = 0, b = 0, p = 0;
    label *owner = nullptr, *neighbour = nullptr;   // face lists as sent by amr_mesh_set (absent after amr_mesh_update: faceLists)
    label *bOwner = nullptr;   // [b] faceOwner of the boundary faces (= owner + f, or its own allocation after amr_mesh_update)
    bool bOwnerOwned = false, meshSet = false, faceLists = false;
    std::vector<Patch> patches;
    int64_t nCoupled = 0;
    u8 *bCoupled = nullptr;    // [b] boundary face lies on a processor patch
    label *cplFace = nullptr;  // [nCoupled] boundary-face indices (face - f) of the coupled faces
    label *cplCell = nullptr;  // [nCplCells] unique owners of coupled faces, ascending
    int64_t nCplCells = 0;
    u8 *cellGhost = nullptr;   // [n] cell owns a coupled face
    int32_t *lkOff = nullptr;  // [nCplCells + 1] push-on-change addressing of the linked closure loops (kernels_linked.cuh)
    unsigned *lkDst = nullptr; // [nCoupled] rank << 26 | byte offset in that rank's ghost array
    bool haveLink = false;
    Sell cc;                   // cell -> face-neighbour cells (+ ghost slots n + bface for coupled faces)
    Sell pc;                   // point -> cells
    bool havePoints = false;
    // what amr_mesh_update keeps of the OLD mesh for the amr_mesh_update_points that follows it
    struct Carry { bool valid = false; Sell pcOld; u8 *bndPointOld = nullptr; label *pointLevelOld = nullptr; label *rev = nullptr; } carry;
    // polyRefiner
    bool haveFaces = false;
    label *faceOff = nullptr, *facePts = nullptr;
    u8 *polyBad = nullptr;     // [p] polyRefiner split-point markup (points that cannot be split points), mesh-constant
    int polyBadKind = -1;      // -1: not built; 0 polyhedralRefinement, 1 prismatic2DRefinement
    int optEdgeConsistency = 1, optPolyForce = 0;
    // 2-D cutter: edges as split elements
    bool haveEdges = false;
    int64_t nEdges = 0;
    Sell ec;                   // edge -> cells
    Sell ce;                   // cell -> cells sharing an edge (polyRefiner closure on the queue machinery), built on first use
    int ceState = 0;           // 0 not built, 1 built, -1 not available (transient list too large): dense sweeps
    int balancedEdges = -1;    // |dLevel| <= 1 over all pairs of cells sharing an edge, on every rank (-1: not checked yet)
    u8 *bndEdge = nullptr;
    label *edgePts = nullptr;  // [2*nEdges]
    // geometry (Lohner)
    bool haveGeom = false;
    double *gWeights = nullptr, *gSf = nullptr, *gMagSf = nullptr, *gDelta = nullptr, *gV = nullptr;
    Sell cf;                   // cell -> faces (2*face + role), ascending
    double *grad = nullptr;    // [3n]
    double *xbf = nullptr;     // [b] boundaryField of x
    int mapForm = 1;                // amr_map_fields on device buffers: 0 fused (56 registers), 1 fused with 40 registers (default), 2 one launch per field
    int64_t pushMinCells = 1 << 16;
    int balancedLocal = -1;         // internal faces of the mesh satisfy |dLevel| <= 1 on EVERY rank (-1: not checked yet)
    int balancedGlobal = -1;        // ... and the coupled faces, on every rank
    long long *hMarked = nullptr;   // pinned: source cells seen per buffer-layer / sweep site by the previous call
    unsigned long long *dMarked = nullptr;   // [16] device: the counts of the current call, by site
    long long *hMarkedStage = nullptr;       // pinned read-back staging (= hMarked + 16)
    unsigned sitesTouched = 0;               // sites a layer kernel reported for in this call
    int lookResult = 0;                      // queueLook: index of the first empty queue, or 0
    Mailbox *mailbox = nullptr;              // pinned (device-visible through UVA)
    unsigned mailSeq = 0;
    int pullNum = 1, pullDen = 2;   // ... and the sparse pull form (rows of the unmarked cells) when sources/n > pullNum/pullDen
    int pushNum = 1, pushDen = 4;   // buffer layers take the sparse push form when marked/n < pushNum/pushDen (0 = never)
    label *ppPts = nullptr;         // points of the processor patches, patch after patch (syncPointList pairing)
    std::vector<int32_t> ppOff;     // [npatch+1] offsets into ppPts
    label *visible = nullptr, *splitParent = nullptr;   // refinement history (hexRefRefinementHistory), kept for cluster queries
    int64_t nSplit = 0;
    u8 *bEmpty = nullptr;      // [b] boundary face lies on an empty patch (no surface-field entries)
    double *lsP = nullptr, *lsN = nullptr;   // leastSquaresVectors pVectors [3(f+b)], nVectors [3f]
    bool haveLs = false;
    u8 *bndPoint = nullptr;
    u8 *pcClass = nullptr;     // [p] split-point pre-filter (see k_point_class); rebuilt when points/levels change
    bool pcClassDirty = true;
    int64_t nSplitCand = 0;
    label *candList = nullptr;  // [nSplitCand] ascending candidate points (sparse case)

    // levels / history
    bool haveLevels = false, haveHistory = false;
    label *lvl = nullptr;      // [n] cellLevel
    u8 *lvl8 = nullptr;        // [n] cellLevel as bytes (gather-friendly)
    label *parentIdx = nullptr;// [n] history.parentIndex(c) or -1 (not visible / unrefined)
    label *pointLevel = nullptr;
    u8 *protCell = nullptr;    // [n] or null

    // resident results
    double *error = nullptr;   // [n]
    u8 *refineCell = nullptr;  // [n] marking kept between the refine and unrefine stages
    int64_t refineCellN = 0;
    bool refineCellOversize = false;   // marking held in a private (non-pool) buffer until the next amr_mesh_set
    bool errorResident = false;        // ctx->error holds the last estimate
    int64_t flagCap = 0;        // capacity (bytes) of flagA/B/C and refineCell: they are swapped by pointer
    int anyProtCached = -1;     // returnReduce(protectedCell_.size()) > 0, cached per mesh
    int64_t nTotalCells = -1;   // mesh_.globalData().nTotalCells(), reduced once per mesh

    // scratch (sized at mesh_set)
    u8 *flagA = nullptr, *flagB = nullptr;      // [max(n,p)]
    u8 *flagC = nullptr;                        // [max(n,p)]
    u8 *lf = nullptr;                           // [n] level +/- flag
    unsigned *srcBits = nullptr;                // [n/32] source mask of a sparse-pull layer as bits (L2-resident gathers)
    u8 *ghost8 = nullptr;       // [b] received byte payloads (marks, level +/- flag), boundary-face layout
    double *ghost64 = nullptr;  // [b] received field values
    unsigned long long *lbStatus = nullptr;     // [tileCap] tile status words of the single-pass compaction (scan.cuh)
    unsigned *lbTicket = nullptr;               // dynamic tile ticket
    unsigned scanTag = 0;                       // per-call tag of the status words
    int32_t *dTotals = nullptr;                 // [8] list sizes written by the compaction kernels
    int64_t tileCap = 0;
    // work queues of the closure loops (kernels_sparse.cuh): queue j lives in queue[j & 1], its size in dQ[j], the size
    // summed over ranks in dGQ[j]
    label *queue[2] = {nullptr, nullptr};
    int64_t queueCap = 0;
    int *dQ = nullptr, *dGQ = nullptr;          // [kMaxIter + 2] each
    int *hQ = nullptr;                          // pinned mirror of dGQ then dQ
    int itersSeen[2] = {-1, -1};                // iterations the refine / unrefine closure needed last time (speculation depth)
    // split-point table of hexRef3D (mesh-constant: rebuilt when points / levels / history change)
    bool spDirty = true;
    int64_t nSp = 0;
    label *spPoint = nullptr;                   // [nSp] ascending point labels (getSplitElems)
    label *spCells = nullptr;                   // [8 * nSp] k-th cell of split point i at [k * nSp + i]
    label *cellSp = nullptr;                    // [n] split point of the cell, or -1
    u8 *spFlag = nullptr;                       // [nSp + 16]
    label *spList = nullptr;                    // [nSp] compact list of the split points that passed the filter
    int refineCellGen = 1;                      // largest generation tag in the resident marking
    int *dFlags = nullptr;                      // [16] device scalars (changed flags, counters)
    int64_t *dCount = nullptr;                  // [4]
    int *hFlags = nullptr;                      // pinned mirror
    int64_t *hCount = nullptr;

    // per-call staging arena (host<->device copies for host-pointer arguments)
    std::vector<DevBuf> arena;
    size_t arenaUsed = 0;       // within arena.back()
    struct PendingOut { void *host; const void *dev; size_t bytes; };
    std::vector<PendingOut> outs;
    bool hostTouched = false;   // this call copied from/to host memory: it must be complete on return
};

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
            return AMR_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define NCK(call)                                                                             \
    do {                                                                                      \
        ncclResult_t r_ = (call);                                                             \
        if (r_ != ncclSuccess) {                                                              \
            ctx->err = std::string(#call) + ": " + ncclGetErrorString(r_);                    \
            return AMR_ERR_NCCL;                                                              \
        }                                                                                     \
    } while (0)

#define FAIL(code, msg)                                                                       \
    do {                                                                                      \
        ctx->err = (msg);                                                                     \
        return (code);                                                                        \
    } while (0)

#define TRY(expr)                                                                             \
    do {                                                                                      \
        int rc_ = (expr);                                                                     \
        if (rc_ != AMR_OK) return rc_;                                                        \
    } while (0)

#include <chrono>
static inline void hostMark(amr_ctx *ctx, const char *what) {
    if (!ctx->hostTime) return;
    cudaStreamSynchronize(ctx->stream);
    const double now = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    if (what) fprintf(stderr, "[amr host +%8.2f ms] %s\n", 1e3 * (now - ctx->htLast), what);
    ctx->htLast = now;
}
static inline void traceMark(amr_ctx *ctx, const char *name) {
    if (!ctx->traceOn || ctx->trace.size() >= 200000) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, ctx->stream);
    ctx->trace.push_back({name, e});
}

// kernel launch with bookkeeping (gpu_launches of bench.py is this counter)
#define LAUNCH(kern, grid, block, ...)                                                        \
    do {                                                                                      \
        kern<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__);                               \
        ctx->launches++;                                                                      \
        traceMark(ctx, #kern);                                                                \
    } while (0)

static inline unsigned gridFor(int64_t items, int perBlock = kThreads) {
    int64_t g = (items + perBlock - 1) / perBlock;
    return (unsigned)(g < 1 ? 1 : g);
}

template <typename T>
static int devAlloc(amr_ctx *ctx, T **ptr, int64_t count) {
    size_t bytes = sizeof(T) * (size_t)(count > 0 ? count : 1);
    CK(cudaMalloc((void **)ptr, bytes));
    ctx->devBytes += (int64_t)bytes;
    return AMR_OK;
}
template <typename T>
static void devFree(amr_ctx *ctx, T **ptr) {
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    (void)ctx;
}

static inline bool isDevicePtr(const void *ptr) {
    if (!ptr) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---- bulk upload of pageable host arrays -------------------------------------------------
// cudaMemcpy from pageable memory moves ~12 GB/s (one driver thread bouncing through pinned buffers); the mesh
// addressing handed over after every topology change is several GB.  Here kUploadLanes host threads copy interleaved
// 16 MiB chunks into their own double-buffered pinned staging and queue the DMA on their own stream, so the host
// memcpy runs at multi-core bandwidth and overlaps the PCIe transfer.  Synchronous: complete on return.
constexpr int kUploadLanesMax = 6;
constexpr size_t kUploadChunk = (size_t)16 << 20;
// lanes per context: the host cores are shared by the ranks of the node (one context each): cores / ranks, at most 6
// (AMR_B200_UPLOAD_LANES overrides)
static int uploadLanes(const amr_ctx *ctx) {
    const char *e = getenv("AMR_B200_UPLOAD_LANES");
    if (e && atoi(e) > 0) return atoi(e) > 64 ? 64 : atoi(e);
    int cores = (int)std::thread::hardware_concurrency();
    if (cores <= 0) cores = 8;
    int lanes = cores / (ctx->nranks > 1 ? ctx->nranks : 1);
    return lanes < 1 ? 1 : (lanes > kUploadLanesMax ? kUploadLanesMax : lanes);
}
static int bulkUpload(amr_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (ctx->lanes.empty()) {
        ctx->lanes.resize((size_t)uploadLanes(ctx));
        for (auto &L : ctx->lanes) {
            for (int k = 0; k < 2; k++) {
                if (cudaMallocHost(&L.pin[k], kUploadChunk) != cudaSuccess) { cudaGetLastError(); L.pin[k] = nullptr; }
                cudaEventCreateWithFlags(&L.ev[k], cudaEventDisableTiming);
            }
            cudaStreamCreateWithFlags(&L.s, cudaStreamNonBlocking);
        }
    }
    bool ok = true;
    for (auto &L : ctx->lanes) ok = ok && L.pin[0] && L.pin[1] && L.s && L.ev[0] && L.ev[1];
    if (!ok) {          // no pinned memory to be had: plain copy
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return AMR_OK;
    }
    CK(cudaStreamSynchronize(ctx->stream));       // whatever still uses dst is done
    std::atomic<int> err{0};
    const size_t nLanes = ctx->lanes.size();
    auto work = [&](int t) {
        if (cudaSetDevice(ctx->device) != cudaSuccess) { err = 1; return; }
        amr_ctx::UploadLane &L = ctx->lanes[(size_t)t];
        int slot = 0;
        for (size_t off = (size_t)t * kUploadChunk; off < bytes; off += nLanes * kUploadChunk) {
            const size_t len = bytes - off < kUploadChunk ? bytes - off : kUploadChunk;
            if (cudaEventSynchronize(L.ev[slot]) != cudaSuccess) { err = 1; return; }     // the slot's previous DMA has drained
            memcpy(L.pin[slot], (const char *)src + off, len);
            if (cudaMemcpyAsync((char *)dst + off, L.pin[slot], len, cudaMemcpyHostToDevice, L.s) != cudaSuccess) { err = 1; return; }
            cudaEventRecord(L.ev[slot], L.s);
            slot ^= 1;
        }
        if (cudaStreamSynchronize(L.s) != cudaSuccess) err = 1;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < (int)nLanes; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    if (err) { ctx->err = "bulk upload failed"; cudaGetLastError(); return AMR_ERR_CUDA; }
    return AMR_OK;
}
// host or device source -> device; large pageable sources take the staged multi-thread path
static int uploadAuto(amr_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return AMR_OK;
    if (bytes >= ((size_t)32 << 20)) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, src) != cudaSuccess) { cudaGetLastError(); a.type = cudaMemoryTypeUnregistered; }
        if (a.type == cudaMemoryTypeUnregistered) return bulkUpload(ctx, dst, src, bytes);
    }
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    return AMR_OK;
}

// ---- staging arena ---------------------------------------------------------------------
static void arenaReset(amr_ctx *ctx) {
    if (ctx->arena.size() > 1) {
        size_t total = 0;
        for (auto &bf : ctx->arena) { total += bf.cap; cudaFree(bf.p); ctx->devBytes -= (int64_t)bf.cap; }
        ctx->arena.clear();
        DevBuf nb;
        if (cudaMalloc(&nb.p, total) == cudaSuccess) { nb.cap = total; ctx->devBytes += (int64_t)total; ctx->arena.push_back(nb); }
        else cudaGetLastError();
    }
    ctx->arenaUsed = 0;
    ctx->outs.clear();
    ctx->hostTouched = false;
}
static int arenaAlloc(amr_ctx *ctx, void **out, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    if (ctx->arena.empty() || ctx->arenaUsed + bytes > ctx->arena.back().cap) {
        DevBuf nb;
        size_t cap = bytes > ((size_t)1 << 20) ? bytes : ((size_t)1 << 20);
        CK(cudaMalloc(&nb.p, cap));
        nb.cap = cap; ctx->devBytes += (int64_t)cap;
        ctx->arena.push_back(nb);
        ctx->arenaUsed = 0;
    }
    *out = (char *)ctx->arena.back().p + ctx->arenaUsed;
    ctx->arenaUsed += bytes;
    return AMR_OK;
}
// input: device pointer used as is, host pointer staged
template <typename T>
static int stageIn(amr_ctx *ctx, const T *ptr, int64_t count, const T **dev) {
    if (!ptr) { *dev = nullptr; return AMR_OK; }
    if (isDevicePtr(ptr)) { *dev = ptr; return AMR_OK; }
    void *q;
    TRY(arenaAlloc(ctx, &q, sizeof(T) * (size_t)(count > 0 ? count : 1)));
    if (count > 0) TRY(uploadAuto(ctx, q, ptr, sizeof(T) * (size_t)count));
    ctx->hostTouched = true;
    *dev = (const T *)q;
    return AMR_OK;
}
// output: device pointer written directly, host pointer gets a staged buffer copied back by flushOuts
template <typename T>
static int stageOut(amr_ctx *ctx, T *ptr, int64_t count, T **dev) {
    if (!ptr) { *dev = nullptr; return AMR_OK; }
    if (isDevicePtr(ptr)) { *dev = ptr; return AMR_OK; }
    void *q;
    TRY(arenaAlloc(ctx, &q, sizeof(T) * (size_t)(count > 0 ? count : 1)));
    *dev = (T *)q;
    ctx->outs.push_back({(void *)ptr, q, sizeof(T) * (size_t)(count > 0 ? count : 0)});
    return AMR_OK;
}
// in/out: staged copy in, copy back at flush
template <typename T>
static int stageInOut(amr_ctx *ctx, T *ptr, int64_t count, T **dev) {
    if (!ptr) { *dev = nullptr; return AMR_OK; }
    if (isDevicePtr(ptr)) { *dev = ptr; return AMR_OK; }
    void *q;
    TRY(arenaAlloc(ctx, &q, sizeof(T) * (size_t)(count > 0 ? count : 1)));
    if (count > 0) TRY(uploadAuto(ctx, q, ptr, sizeof(T) * (size_t)count));
    *dev = (T *)q;
    ctx->outs.push_back({(void *)ptr, q, sizeof(T) * (size_t)(count > 0 ? count : 0)});
    return AMR_OK;
}
// ---- peer-memory error word ---------------------------------------------------------------------------------
// A device-side spin that timed out (p2p.cuh) raises the error word of this rank's region.  It is fetched with every
// host round trip of a call that enqueued a peer exchange and turned into AMR_ERR_STATE; reporting clears it.  Calls
// that are only stream-ordered (all data pointers on the device) report it at the next synchronising call or at
// amr_synchronize, like an asynchronous CUDA error.
constexpr size_t kPeerErrOffset = 2048;    // offsetof(P2PRegionHeader, error), asserted in amr_b200.cu
static int peerErrFetch(amr_ctx *ctx) {
    ctx->hFlags[9] = 0;
    if (ctx->usePeer && ctx->peerTouched)
        CK(cudaMemcpyAsync(ctx->hFlags + 9, ctx->region + kPeerErrOffset, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return AMR_OK;
}
// after the stream was synchronised
static int peerErrCheck(amr_ctx *ctx) {
    ctx->peerTouched = false;
    if (ctx->hFlags[9]) {
        ctx->hFlags[9] = 0;
        cudaMemsetAsync(ctx->region + kPeerErrOffset, 0, sizeof(int), ctx->stream);
        FAIL(AMR_ERR_STATE, "peer-memory halo swap / reduction timed out: a neighbour rank did not enter the same collective call "
                            "(results of this call are invalid; AMR_B200_PEER_TIMEOUT_S sets the limit)");
    }
    return AMR_OK;
}

// Copy staged outputs back and make the call complete with respect to the host.  A call whose data
// pointers were all device pointers is only stream-ordered (normal CUDA semantics for device buffers):
// no host synchronisation, the next call on the context's stream sees the results.
static int flushOuts(amr_ctx *ctx) {
    const bool sync = ctx->hostTouched || !ctx->outs.empty();
    for (auto &o : ctx->outs)
        if (o.bytes) CK(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->outs.clear();
    if (sync) {
        TRY(peerErrFetch(ctx));
        CK(cudaStreamSynchronize(ctx->stream));
        TRY(peerErrCheck(ctx));
    }
    return AMR_OK;
}
