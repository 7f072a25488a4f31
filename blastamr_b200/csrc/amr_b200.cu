// amr_b200.cu -- C-ABI of the B200-native AMR decision path (see include/amr_b200.h).
//
// Host-side orchestration only: argument staging, the sweep loops with their convergence test,
// processor-patch halo swaps over NCCL, and the ascending-list compaction.  All arithmetic is in
// kernels.cuh.  There is no CPU code path: every entry point needs a live CUDA context.
#include "common.cuh"
#include "scan.cuh"
#include "sell.cuh"
#include "kernels.cuh"
#include "kernels_pipe.cuh"
#include "kernels_sparse.cuh"
#include "kernels_linked.cuh"
#include "lohner.cuh"
#include "fields.cuh"
#include "p2p.cuh"

#include <algorithm>
#include <new>

// ============================================================================ pipeline launch

constexpr size_t kPipeSmemLimit = 96 * 1024;   // beyond this the ring would cost too much occupancy: direct kernels

static bool pipeOk(const amr_ctx *ctx, const Sell &S) {
    return ctx->usePipeline && S.rows > 0 && pipeSmemBytes(S.maxTileInts, 2) <= kPipeSmemLimit;
}

template <class K, class... Args>
static int launchPipe(amr_ctx *ctx, K kern, const Sell &S, Args... args) {
    // meshes without processor patches have no cell that waits for a boundary fix-up: no per-slice flag to fetch (the
    // flag load sits right before the tile wait and cost ~20 % of kp_estimate's stall samples, profiles/r1_ncu_full.csv)
    SellView V{S.sliceOff, S.col, S.rows, S.slices, S.tileFirst, S.tiles, S.maxTileInts, kStages, ctx->nCoupled > 0 ? S.sliceGhost : (const u8 *)nullptr};
    // Ring depth: three stages (two only when three do not fit at all); the stage is kept small by capping the columns of
    // a tile (sell.cuh), not by a shallower ring.
    const void *key = (const void *)kern;
    auto it = ctx->occCache.find(key);
    int perSM;
    if (it == ctx->occCache.end()) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPipeSmemLimit));
        int p3 = 0, p2 = 0;
        if (pipeSmemBytes(V.stageInts, 3) <= kPipeSmemLimit) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p3, kern, kPipeThreads, pipeSmemBytes(V.stageInts, 3)));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p2, kern, kPipeThreads, pipeSmemBytes(V.stageInts, 2)));
        int stages = p3 >= 1 ? 3 : 2;         // (two stages with more resident CTAs were measured SLOWER: 0.54 vs 0.35 ms for kp_estimate at 23.7 M cells)
        if (ctx->pipeStagesForced == 2 || ctx->pipeStagesForced == 3) stages = (ctx->pipeStagesForced == 3 && p3 < 1) ? 2 : ctx->pipeStagesForced;
        perSM = stages == 3 ? p3 : p2;
        if (perSM < 1) perSM = 1;
        ctx->occCache[key] = perSM * 4 + stages;
    } else perSM = it->second / 4;
    V.stages = ctx->occCache[key] % 4;
    const size_t smem = pipeSmemBytes(V.stageInts, V.stages);
    const int64_t ntiles = S.tiles;
    int64_t grid = (int64_t)ctx->numSMs * perSM;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kPipeThreads, smem, ctx->stream>>>(V, args...);
    ctx->launches++;
    traceMark(ctx, "pipelined SELL kernel");
    return AMR_OK;
}

// ============================================================================ context

AMR_API int amr_ctx_create(int device, amr_ctx **out) {
    if (!out) return AMR_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return AMR_ERR_CUDA;   // no CPU fallback by design
    }
    amr_ctx *ctx = new (std::nothrow) amr_ctx();
    if (!ctx) return AMR_ERR_CUDA;
    ctx->device = device;
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) ctx->numSMs = sms;
        const char *np = getenv("AMR_B200_NO_PIPELINE");
        if (np && np[0] == '1') ctx->usePipeline = false;
        // tuning knob: "num/den" = marked fraction below which a buffer layer takes the sparse push form; "0" = never
        const char *pf = getenv("AMR_B200_PUSH_FRACTION");
        if (pf) {
            int a = 0, b = 1;
            if (sscanf(pf, "%d/%d", &a, &b) >= 1 && a >= 0 && b > 0) { ctx->pushNum = a; ctx->pushDen = b; }
        }
        const char *lf = getenv("AMR_B200_PULL_FRACTION");
        if (lf) {
            int a = 0, b = 1;
            if (sscanf(lf, "%d/%d", &a, &b) >= 1 && a >= 0 && b > 0) { ctx->pullNum = a; ctx->pullDen = b; }
        }
        const char *fb = getenv("AMR_B200_FACE_BATCH");          // tuning / A-B: 1 = four faces at a time in the Lohner / Gauss-gradient walks
        if (fb && fb[0] == '1') ctx->faceBatch = true;
        const char *ps = getenv("AMR_B200_PIPE_STAGES");          // tuning / A-B: 2 or 3 ring stages (default: chosen per mesh)
        if (ps) ctx->pipeStagesForced = atoi(ps);
        const char *ht = getenv("AMR_B200_HOSTTIME");
        if (ht && ht[0] == '1') ctx->hostTime = true;
        const char *tr = getenv("AMR_B200_TRACE");
        if (tr && tr[0]) { ctx->traceOn = true; ctx->tracePath = tr; }
        const char *mf = getenv("AMR_B200_MAP");                 // tuning: fused40 (default) | fused | separate
        if (mf && strcmp(mf, "fused") == 0) ctx->mapForm = 0;
        else if (mf && strcmp(mf, "fused40") == 0) ctx->mapForm = 1;
        else if (mf && strcmp(mf, "separate") == 0) ctx->mapForm = 2;
        const char *pm = getenv("AMR_B200_PUSH_MIN_CELLS");      // below this mesh size the extra launches do not pay
        if (pm) ctx->pushMinCells = atoll(pm);
    }
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return AMR_ERR_CUDA;
    }
    cudaMalloc((void **)&ctx->dFlags, 16 * sizeof(int));
    cudaMalloc((void **)&ctx->dCount, 4 * sizeof(int64_t));
    cudaMalloc((void **)&ctx->dSeq, 2 * sizeof(unsigned));
    cudaMalloc((void **)&ctx->lbTicket, sizeof(unsigned));
    cudaMalloc((void **)&ctx->dTotals, 8 * sizeof(int32_t));
    cudaMalloc((void **)&ctx->dQ, 2 * (kMaxIter + 2) * sizeof(int) + 16 * sizeof(unsigned long long));   // queue sizes + site counts: one memset per call
    cudaMallocHost((void **)&ctx->hFlags, 16 * sizeof(int));
    cudaMallocHost((void **)&ctx->hCount, 4 * sizeof(int64_t));
    cudaMallocHost((void **)&ctx->hMarked, 32 * sizeof(long long));
    cudaMallocHost((void **)&ctx->hQ, 2 * (kMaxIter + 2) * sizeof(int));
    cudaMallocHost((void **)&ctx->mailbox, sizeof(Mailbox));
    if (ctx->mailbox) memset(ctx->mailbox, 0, sizeof(Mailbox));
    if (ctx->hMarked) { for (int i = 0; i < 32; i++) ctx->hMarked[i] = -1; ctx->hMarkedStage = ctx->hMarked + 16; }
    if (!ctx->dFlags || !ctx->dCount || !ctx->hFlags || !ctx->hCount || !ctx->hMarked || !ctx->dSeq || !ctx->lbTicket || !ctx->dTotals ||
        !ctx->dQ || !ctx->hQ || !ctx->mailbox) { delete ctx; return AMR_ERR_CUDA; }
    ctx->dGQ = ctx->dQ + (kMaxIter + 2);
    ctx->dMarked = (unsigned long long *)(ctx->dQ + 2 * (kMaxIter + 2));
    static_assert((2 * (kMaxIter + 2) * sizeof(int)) % 8 == 0, "site counts must be 8-byte aligned");
    cudaMemset(ctx->dFlags, 0, 16 * sizeof(int));
    cudaMemset(ctx->dCount, 0, 4 * sizeof(int64_t));
    cudaMemset(ctx->dSeq, 0, 2 * sizeof(unsigned));
    cudaMemset(ctx->lbTicket, 0, sizeof(unsigned));
    cudaMemset(ctx->dTotals, 0, 8 * sizeof(int32_t));
    cudaMemset(ctx->dQ, 0, 2 * (kMaxIter + 2) * sizeof(int) + 16 * sizeof(unsigned long long));
    *out = ctx;
    return AMR_OK;
}

static void freeMesh(amr_ctx *ctx) {
    if (ctx->bOwnerOwned) devFree(ctx, &ctx->bOwner);
    ctx->bOwner = nullptr; ctx->bOwnerOwned = false; ctx->meshSet = false; ctx->faceLists = false;
    devFree(ctx, &ctx->owner); devFree(ctx, &ctx->neighbour); devFree(ctx, &ctx->bCoupled);
    devFree(ctx, &ctx->cplFace); devFree(ctx, &ctx->cplCell); devFree(ctx, &ctx->cellGhost); ctx->nCplCells = 0;
    devFree(ctx, &ctx->lkOff); devFree(ctx, &ctx->lkDst); ctx->haveLink = false;
    sellFree(ctx, ctx->cc); sellFree(ctx, ctx->pc);
    sellFree(ctx, ctx->carry.pcOld); devFree(ctx, &ctx->carry.bndPointOld); devFree(ctx, &ctx->carry.pointLevelOld); devFree(ctx, &ctx->carry.rev);
    ctx->carry.valid = false;
    devFree(ctx, &ctx->bndPoint); devFree(ctx, &ctx->pcClass); devFree(ctx, &ctx->candList); ctx->pcClassDirty = true; ctx->balancedLocal = -1;
    devFree(ctx, &ctx->lvl); devFree(ctx, &ctx->lvl8); devFree(ctx, &ctx->parentIdx); devFree(ctx, &ctx->pointLevel);
    devFree(ctx, &ctx->protCell);
    devFree(ctx, &ctx->error); devFree(ctx, &ctx->refineCell);
    devFree(ctx, &ctx->flagA); devFree(ctx, &ctx->flagB); devFree(ctx, &ctx->flagC); devFree(ctx, &ctx->lf); devFree(ctx, &ctx->srcBits);
    devFree(ctx, &ctx->ghost8); devFree(ctx, &ctx->ghost64);
    devFree(ctx, &ctx->lbStatus);
    devFree(ctx, &ctx->queue[0]); devFree(ctx, &ctx->queue[1]); ctx->queueCap = 0;
    devFree(ctx, &ctx->spPoint); devFree(ctx, &ctx->spCells); devFree(ctx, &ctx->cellSp); devFree(ctx, &ctx->spFlag); devFree(ctx, &ctx->spList);
    ctx->nSp = 0; ctx->spDirty = true; ctx->balancedGlobal = -1; ctx->refineCellGen = 1;
    devFree(ctx, &ctx->gWeights); devFree(ctx, &ctx->gSf); devFree(ctx, &ctx->gMagSf); devFree(ctx, &ctx->gDelta); devFree(ctx, &ctx->gV);
    devFree(ctx, &ctx->grad); devFree(ctx, &ctx->xbf);
    devFree(ctx, &ctx->bEmpty); devFree(ctx, &ctx->lsP); devFree(ctx, &ctx->lsN); ctx->haveLs = false;
    devFree(ctx, &ctx->visible); devFree(ctx, &ctx->splitParent); ctx->nSplit = 0;
    devFree(ctx, &ctx->ppPts); ctx->ppOff.clear(); ctx->havePointPlan = false;
    sellFree(ctx, ctx->cf);
    ctx->haveGeom = false;
    sellFree(ctx, ctx->ec); devFree(ctx, &ctx->bndEdge); devFree(ctx, &ctx->edgePts);
    sellFree(ctx, ctx->ce); ctx->ceState = 0; ctx->balancedEdges = -1;
    devFree(ctx, &ctx->faceOff); devFree(ctx, &ctx->facePts); ctx->haveFaces = false;
    devFree(ctx, &ctx->polyBad); ctx->polyBadKind = -1;
    ctx->haveEdges = false; ctx->nEdges = 0;
    ctx->havePoints = ctx->haveLevels = ctx->haveHistory = false;
    ctx->occCache.clear();
    ctx->refineCellN = 0;
    ctx->devBytes = 0;
    for (auto &bf : ctx->arena) ctx->devBytes += (int64_t)bf.cap;
}

static void traceDump(amr_ctx *ctx) {
    if (!ctx->traceOn || ctx->trace.empty()) return;
    char path[512];
    snprintf(path, sizeof(path), "%s.%d.%p", ctx->tracePath.c_str(), ctx->rank, (void *)ctx);
    FILE *fh = fopen(path, "w");
    if (fh) {
        const size_t first = ctx->trace.size() > 6000 ? ctx->trace.size() - 6000 : 0;     // the last calls
        double t = 0.0;
        for (size_t i = first; i < ctx->trace.size(); i++) {
            float ms = 0.f;
            if (i > first) cudaEventElapsedTime(&ms, ctx->trace[i - 1].ev, ctx->trace[i].ev);     // event to event: full resolution
            t += (double)ms * 1e3;
            fprintf(fh, "%10.1f %8.1f %s\n", t, (double)ms * 1e3, ctx->trace[i].name);
        }
        fclose(fh);
    }
    for (auto &r : ctx->trace) cudaEventDestroy(r.ev);
    ctx->trace.clear();
}

AMR_API int amr_ctx_destroy(amr_ctx *ctx) {
    if (!ctx) return AMR_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    traceDump(ctx);
    freeMesh(ctx);
    for (auto &bf : ctx->arena) cudaFree(bf.p);
    if (ctx->usePeer) {
        for (int r = 0; r < ctx->nranks; r++)
            if (r != ctx->rank && ctx->peerHost[(size_t)r]) cudaIpcCloseMemHandle(ctx->peerHost[(size_t)r]);
        cudaFree(ctx->peerDev); cudaFree(ctx->redVal);
        delete (HaloPlan *)ctx->haloPlan;
        delete (HaloPlan *)ctx->pointPlanPtr;
    }
    if (ctx->region) cudaFree(ctx->region);
    if (ctx->comm) ncclCommDestroy(ctx->comm);
    cudaFree(ctx->dFlags); cudaFree(ctx->dCount); cudaFree(ctx->dSeq); cudaFree(ctx->lbTicket); cudaFree(ctx->dTotals); cudaFree(ctx->dQ);
    cudaFreeHost(ctx->hFlags); cudaFreeHost(ctx->hCount); cudaFreeHost(ctx->hMarked); cudaFreeHost(ctx->hQ); cudaFreeHost(ctx->mailbox);
    for (auto &L : ctx->lanes) {
        for (int k = 0; k < 2; k++) { if (L.pin[k]) cudaFreeHost(L.pin[k]); if (L.ev[k]) cudaEventDestroy(L.ev[k]); }
        if (L.s) cudaStreamDestroy(L.s);
    }
    if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->upStream) cudaStreamDestroy(ctx->upStream);
    if (ctx->downStream) cudaStreamDestroy(ctx->downStream);
    delete ctx;
    return AMR_OK;
}

AMR_API const char *amr_last_error(const amr_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

AMR_API int amr_set_stream(amr_ctx *ctx, void *s) {
    if (!ctx) return AMR_ERR_ARG;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)s;
    ctx->ownStream = false;
    return AMR_OK;
}

AMR_API int amr_comm_unique_id(void *id128) {
    if (!id128) return AMR_ERR_ARG;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return AMR_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(id128, &id, 128);
    return AMR_OK;
}

static int peerSetup(amr_ctx *ctx);

AMR_API int amr_comm_init(amr_ctx *ctx, int nranks, int rank, const void *id128) {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return AMR_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NCK(ncclCommInitRank(&ctx->comm, nranks, id, rank));
    ctx->nranks = nranks; ctx->rank = rank;
    return peerSetup(ctx);
}

AMR_API int64_t amr_launch_count(const amr_ctx *ctx) { return ctx ? ctx->launches : 0; }
AMR_API int64_t amr_device_bytes(const amr_ctx *ctx) { return ctx ? ctx->devBytes : 0; }

#define ENTER()                                                                               \
    if (!ctx) return AMR_ERR_ARG;                                                             \
    CK(cudaSetDevice(ctx->device));                                                           \
    arenaReset(ctx);                                                                          \
    ctx->err.clear();                                                                         \
    traceMark(ctx, __func__)

// ============================================================================ halo swaps

// ---- NVLink peer-memory transport (p2p.cuh) ----------------------------------------------------------

// all ranks map each other's region; NCCL is only the transport for the 64-byte IPC handles.  Every rank reaches
// the "everybody or nobody" all-reduce whatever failed locally, so no rank is left blocked in NCCL.
static int peerSetup(amr_ctx *ctx) {
    if (ctx->nranks < 2) return AMR_OK;
    const char *mode = getenv("AMR_B200_HALO");
    {
        const char *ts = getenv("AMR_B200_PEER_TIMEOUT_S");
        double sec = ts ? atof(ts) : 20.0;
        if (!(sec > 0.0)) sec = 20.0;
        ctx->peerTimeoutNs = (unsigned long long)(sec * 1e9);
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    static_assert(offsetof(P2PRegionHeader, error) == kPeerErrOffset, "kPeerErrOffset");
    // scratch of the set-up collectives: the one allocation whose failure cannot be agreed on
    char *scratch = nullptr;
    CK(cudaMalloc((void **)&scratch, (size_t)64 * (ctx->nranks + 1)));
    char *dSend = scratch, *dAll = scratch + 64;
    int ok = 1;
    if ((mode && strcmp(mode, "nccl") == 0) || ctx->nranks > kMaxRanks) ok = 0;
    if (ok && cudaMalloc((void **)&ctx->region, (size_t)kRegionBytes) != cudaSuccess) { cudaGetLastError(); ctx->region = nullptr; ok = 0; }
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok) {
        if (cudaMemsetAsync(ctx->region, 0, (size_t)kRegionBytes, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaIpcGetMemHandle(&mine, ctx->region) != cudaSuccess) {
            cudaGetLastError(); ok = 0;
        }
    }
    std::vector<cudaIpcMemHandle_t> all((size_t)ctx->nranks);
    int rc = AMR_OK;
    if (cudaMemcpyAsync(dSend, &mine, 64, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    if (ncclAllGather(dSend, dAll, 64, ncclUint8, ctx->comm, ctx->stream) != ncclSuccess) { ctx->err = "peerSetup: ncclAllGather failed"; rc = AMR_ERR_NCCL; }
    if (rc == AMR_OK && (cudaMemcpyAsync(all.data(), dAll, (size_t)64 * ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                         cudaStreamSynchronize(ctx->stream) != cudaSuccess)) { cudaGetLastError(); ok = 0; }
    ctx->peerHost.assign((size_t)ctx->nranks, nullptr);
    for (int r = 0; r < ctx->nranks && ok && rc == AMR_OK; r++) {
        if (r == ctx->rank) { ctx->peerHost[(size_t)r] = ctx->region; continue; }
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        ctx->peerHost[(size_t)r] = (char *)p;
    }
    if (ok && rc == AMR_OK) {
        if (cudaMalloc((void **)&ctx->peerDev, sizeof(char *) * (size_t)ctx->nranks) != cudaSuccess ||
            cudaMemcpy(ctx->peerDev, ctx->peerHost.data(), sizeof(char *) * (size_t)ctx->nranks, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMalloc((void **)&ctx->redVal, 16) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    // everybody or nobody
    if (rc == AMR_OK) {
        int *word = (int *)scratch;
        if (cudaMemcpyAsync(word, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { cudaGetLastError(); rc = AMR_ERR_CUDA; ctx->err = "peerSetup: copy failed"; }
        else if (ncclAllReduce(word, word, 1, ncclInt32, ncclMin, ctx->comm, ctx->stream) != ncclSuccess) { rc = AMR_ERR_NCCL; ctx->err = "peerSetup: ncclAllReduce failed"; }
        else if (cudaMemcpyAsync(&ok, word, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                 cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); rc = AMR_ERR_CUDA; ctx->err = "peerSetup: copy failed"; }
    }
    cudaFree(scratch);
    if (rc != AMR_OK || !ok) {
        for (int r = 0; r < (int)ctx->peerHost.size(); r++)
            if (r != ctx->rank && ctx->peerHost[(size_t)r]) cudaIpcCloseMemHandle(ctx->peerHost[(size_t)r]);
        ctx->peerHost.clear();
        if (ctx->peerDev) { cudaFree(ctx->peerDev); ctx->peerDev = nullptr; }
        if (ctx->redVal) { cudaFree(ctx->redVal); ctx->redVal = nullptr; }
        if (ctx->region) { cudaFree(ctx->region); ctx->region = nullptr; }
        return rc;          // AMR_OK: the NCCL path stays
    }
    ctx->haloPlan = new HaloPlan();
    ctx->pointPlanPtr = new HaloPlan();
    ctx->usePeer = true;
    return AMR_OK;
}

// Per-mesh plan: where my patch ranges land in each neighbour's boundary-face array.  Collective over ALL ranks of
// the communicator (amr_mesh_set after a topology change is): the ranks agree (all-reduce) whether this mesh uses
// the peer transport or NCCL -- one rank with more patches than the plan holds, or a boundary that does not fit a
// parity buffer, moves everybody to NCCL -- and resynchronise their exchange epochs, so that a rank that had no
// processor patch on the previous mesh cannot wait on a stale arrival word.
static int peerPlan(amr_ctx *ctx) {
    ctx->havePlan = false;
    if (!ctx->usePeer) return AMR_OK;
    HaloPlan &P = *(HaloPlan *)ctx->haloPlan;
    P.nPatch = 0; P.prefix[0] = 0;
    int ok = 1;
    std::vector<int32_t> mine, nbrs;                 // (offset, size) per processor patch, ALL of them
    for (const Patch &q : ctx->patches) {
        if (q.type != AMR_PATCH_PROCESSOR || q.size == 0) continue;
        mine.push_back(q.start - (int32_t)ctx->f); mine.push_back(q.size); nbrs.push_back(q.nbr);
        if (q.nbr < 0 || q.nbr >= ctx->nranks || q.nbr == ctx->rank) FAIL(AMR_ERR_ARG, "amr_mesh_set: processor patch with a bad neighbour rank");
    }
    const int np = (int)nbrs.size();
    if (np > 32) ok = 0;
    if (ctx->b * 24 > kHaloCap) ok = 0;             // the widest payload (vector field) must fit one parity buffer
    std::vector<int32_t> theirs((size_t)(2 * np > 0 ? 2 * np : 1), 0);
    int32_t *dMine = nullptr, *dTheirs = nullptr;
    void *q;
    TRY(arenaAlloc(ctx, &q, (size_t)(2 * np + 4) * 4)); dMine = (int32_t *)q;
    TRY(arenaAlloc(ctx, &q, (size_t)(2 * np + 4) * 4)); dTheirs = (int32_t *)q;
    if (np) CK(cudaMemcpyAsync(dMine, mine.data(), (size_t)2 * np * 4, cudaMemcpyHostToDevice, ctx->stream));
    NCK(ncclGroupStart());
    for (int k = 0; k < np; k++) {
        NCK(ncclSend(dMine + 2 * k, 2, ncclInt32, nbrs[(size_t)k], ctx->comm, ctx->stream));
        NCK(ncclRecv(dTheirs + 2 * k, 2, ncclInt32, nbrs[(size_t)k], ctx->comm, ctx->stream));
    }
    NCK(ncclGroupEnd());
    if (np) CK(cudaMemcpyAsync(theirs.data(), dTheirs, (size_t)2 * np * 4, cudaMemcpyDeviceToHost, ctx->stream));
    // agreement word: [0] = min(ok), [1] = max(halo exchange number), [2] = max(reduction number)   (as -min of negatives)
    unsigned seqNow[2] = {0, 0};
    CK(cudaMemcpyAsync(seqNow, ctx->dSeq, sizeof(seqNow), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int32_t agree[3] = {ok, -(int32_t)(seqNow[0] & 0x3fffffffu), -(int32_t)(seqNow[1] & 0x3fffffffu)};
    int32_t *dAgree = dMine + 2 * np;
    for (int k = 0; k < np; k++)
        if (theirs[(size_t)2 * k + 1] != mine[(size_t)2 * k + 1]) agree[0] = 0;      // the two sides of a patch disagree on its size
    CK(cudaMemcpyAsync(dAgree, agree, sizeof(agree), cudaMemcpyHostToDevice, ctx->stream));
    NCK(ncclAllReduce(dAgree, dAgree, 3, ncclInt32, ncclMin, ctx->comm, ctx->stream));
    CK(cudaMemcpyAsync(agree, dAgree, sizeof(agree), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // common exchange numbers: above every value any rank has used (stale arrival words are below them)
    seqNow[0] = (unsigned)(-agree[1]) + 2u;
    seqNow[1] = (unsigned)(-agree[2]) + 2u;
    CK(cudaMemcpyAsync(ctx->dSeq, seqNow, sizeof(seqNow), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (!agree[0]) return AMR_OK;                    // agreed: NCCL transport for this mesh
    for (int k = 0; k < np; k++) {
        P.nbr[k] = nbrs[(size_t)k]; P.offMine[k] = mine[(size_t)2 * k]; P.size[k] = mine[(size_t)2 * k + 1];
        P.offTheirs[k] = theirs[(size_t)2 * k]; P.prefix[k + 1] = P.prefix[k] + P.size[k];
    }
    P.nPatch = np;
    ctx->havePlan = true;
    return AMR_OK;
}

// Push-on-change addressing of the linked closure loops (kernels_linked.cuh): for every cell that owns processor faces,
// where its value lives in the neighbours' ghost arrays.  Host bookkeeping, once per mesh, after peerPlan.
static int linkPlan(amr_ctx *ctx, const std::vector<label> &bOwn, const std::vector<label> &cells) {
    devFree(ctx, &ctx->lkOff); devFree(ctx, &ctx->lkDst);
    ctx->haveLink = false;
    if (!ctx->usePeer || !ctx->havePlan) return AMR_OK;
    const char *e = getenv("AMR_B200_LINKED");               // tuning / A-B: 0 = three launches per iteration (same on every rank)
    if (e && atoi(e) == 0) return AMR_OK;
    const HaloPlan &P = *(const HaloPlan *)ctx->haloPlan;
    const size_t nCpl = cells.size();
    std::vector<int32_t> off(nCpl + 1, 0);
    std::vector<unsigned> dst((size_t)(ctx->nCoupled > 0 ? ctx->nCoupled : 1), 0u);
    auto slot = [&](label c) { return (size_t)(std::lower_bound(cells.begin(), cells.end(), c) - cells.begin()); };
    for (int q = 0; q < P.nPatch; q++)
        for (int k = 0; k < P.size[q]; k++) off[slot(bOwn[(size_t)(P.offMine[q] + k)]) + 1]++;
    for (size_t i = 0; i < nCpl; i++) off[i + 1] += off[i];
    std::vector<int32_t> fill(off.begin(), off.end() - 1);
    for (int q = 0; q < P.nPatch; q++)
        for (int k = 0; k < P.size[q]; k++) {
            const int64_t there = (int64_t)P.offTheirs[q] + k;
            if (there >= kHaloCap) FAIL(AMR_ERR_STATE, "linkPlan: ghost offset outside the parity buffer");
            dst[(size_t)fill[slot(bOwn[(size_t)(P.offMine[q] + k)])]++] = ((unsigned)P.nbr[q] << 26) | (unsigned)there;
        }
    TRY(devAlloc(ctx, &ctx->lkOff, (int64_t)nCpl + 1));
    TRY(devAlloc(ctx, &ctx->lkDst, (int64_t)dst.size()));
    CK(cudaMemcpy(ctx->lkOff, off.data(), (nCpl + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->lkDst, dst.data(), dst.size() * 4, cudaMemcpyHostToDevice));
    ctx->haveLink = true;
    return AMR_OK;
}
// the closure loops run linked when the peer transport is planned for this mesh (agreed over the ranks in peerPlan)
static inline bool linkedOk(const amr_ctx *ctx) { return ctx->comm && ctx->nranks > 1 && ctx->usePeer && ctx->havePlan && ctx->haveLink; }
static LinkOut makeLinkOut(const amr_ctx *ctx) {
    LinkOut L;
    L.cplCell = ctx->cplCell; L.off = ctx->lkOff; L.dst = ctx->lkDst; L.cellGhost = ctx->cellGhost; L.nCpl = (int)ctx->nCplCells;
    return L;
}

// NCCL transport of a swap (fallback / AMR_B200_HALO=nccl): for every processor patch send
// my boundary-face values [patchStart-f, +size) and receive the neighbour's into the same range
// of `recv` (processor patch pairs list their faces in the same order on both sides).
static int haloExchange(amr_ctx *ctx, const void *send, void *recv, size_t elemBytes) {
    if (ctx->nCoupled == 0) return AMR_OK;
    if (!ctx->comm) FAIL(AMR_ERR_STATE, "mesh has processor patches but no communicator (amr_comm_init) and no host-side neighbour values");
    NCK(ncclGroupStart());
    for (const Patch &q : ctx->patches) {
        if (q.type != AMR_PATCH_PROCESSOR || q.size == 0) continue;
        size_t off = (size_t)(q.start - ctx->f) * elemBytes;
        NCK(ncclSend((const char *)send + off, (size_t)q.size * elemBytes, ncclUint8, q.nbr, ctx->comm, ctx->stream));
        NCK(ncclRecv((char *)recv + off, (size_t)q.size * elemBytes, ncclUint8, q.nbr, ctx->comm, ctx->stream));
    }
    NCK(ncclGroupEnd());
    return AMR_OK;
}

// swap of "the value at the owner cell of every boundary face" (patchInternalField / owner levels / owner
// marks): what every syncTools call site on this path sends.  Peer path: fused gather + push + signal, and the
// receive side (spin on the neighbours' arrival words + unpack) as a second launch that the caller may issue AFTER
// its ghost-free main kernel (haloBegin ... main kernel ... haloEnd ... boundary fix-up): the neighbours' skew and
// the NVLink latency then hide behind the main kernel.  A rank never pushes exchange k+1 before it unpacked
// exchange k, so a neighbour can run at most one exchange ahead and the two parity buffers suffice.
// At most one swap is in flight per context (haloBegin ... haloEnd); exchange numbers live on the device (p2p.cuh).
// gate (optional, device int, identical on all ranks): the swap only happens when *gate != 0.
// aligned destination words of a byte-payload push (k_halo_push_cells<1>)
static inline int64_t pushWords(const HaloPlan &P) {
    int64_t w = 0;
    for (int q = 0; q < P.nPatch; q++) w += (P.size[q] + (P.offTheirs[q] & 3) + 3) >> 2;
    return w;
}
struct HaloToken {
    bool pending = false;
    size_t elemBytes = 0;
    void *recv = nullptr;
    const int *gate = nullptr;
};
static int haloBegin(amr_ctx *ctx, const void *cells, size_t elemBytes, void *recv, HaloToken *tok, const u8 *minus = nullptr,
                     const int *gate = nullptr) {
    tok->pending = false;
    if (ctx->nCoupled == 0) return AMR_OK;
    if (!ctx->comm) FAIL(AMR_ERR_STATE, "mesh has processor patches but no communicator (amr_comm_init) and no host-side neighbour values");
    const int64_t b = ctx->b;
    if (ctx->usePeer && ctx->havePlan) {
        const HaloPlan &P = *(const HaloPlan *)ctx->haloPlan;
        unsigned g = gridFor(P.prefix[P.nPatch]);
        if (elemBytes == 1) g = gridFor(pushWords(P));
        unsigned *counter = (unsigned *)(ctx->dFlags + 6);
        const label *bOwner = ctx->bOwner;
        if (elemBytes == 1) k_halo_push_cells<1><<<g, kThreads, 0, ctx->stream>>>(P, ctx->peerDev, ctx->rank, ctx->dSeq, bOwner, (const char *)cells, minus, counter, gate);
        else if (elemBytes == 4) k_halo_push_cells<4><<<g, kThreads, 0, ctx->stream>>>(P, ctx->peerDev, ctx->rank, ctx->dSeq, bOwner, (const char *)cells, minus, counter, gate);
        else if (elemBytes == 8) k_halo_push_cells<8><<<g, kThreads, 0, ctx->stream>>>(P, ctx->peerDev, ctx->rank, ctx->dSeq, bOwner, (const char *)cells, minus, counter, gate);
        else k_halo_push_cells<24><<<g, kThreads, 0, ctx->stream>>>(P, ctx->peerDev, ctx->rank, ctx->dSeq, bOwner, (const char *)cells, minus, counter, gate);
        ctx->launches++;
        traceMark(ctx, "k_halo_push_cells");
        ctx->peerTouched = true;
        tok->pending = true; tok->elemBytes = elemBytes; tok->recv = recv; tok->gate = gate;
        return AMR_OK;
    }
    if (gate) FAIL(AMR_ERR_STATE, "gated halo swap on the NCCL transport");      // callers use the host-synchronised loops there
    // NCCL path: pack into a send buffer laid out like the boundary-face array, then grouped send/recv (complete here)
    void *send = nullptr;
    TRY(arenaAlloc(ctx, &send, (size_t)b * elemBytes));
    if (elemBytes == 1 && minus) LAUNCH(k_pack_level_minus, gridFor(b), kThreads, b, ctx->bOwner, (const u8 *)cells, minus, (u8 *)send);
    else if (elemBytes == 1) LAUNCH(k_pack_u8, gridFor(b), kThreads, b, ctx->bOwner, (const u8 *)cells, (u8 *)send);
    else if (elemBytes == 4) LAUNCH(k_pack_i32, gridFor(b), kThreads, b, ctx->bOwner, (const label *)cells, (label *)send);
    else if (elemBytes == 8) LAUNCH(k_pack_f64, gridFor(b), kThreads, b, ctx->bOwner, (const double *)cells, (double *)send);
    else LAUNCH(k_pack_vec3, gridFor(b), kThreads, b, ctx->bOwner, (const double *)cells, (double *)send);
    return haloExchange(ctx, send, recv, elemBytes);
}
static int haloEnd(amr_ctx *ctx, HaloToken *tok) {
    if (!tok->pending) return AMR_OK;
    tok->pending = false;
    const HaloPlan &P = *(const HaloPlan *)ctx->haloPlan;
    const unsigned g = gridFor(P.prefix[P.nPatch]);
    const unsigned long long to = ctx->peerTimeoutNs;
    unsigned *counter = (unsigned *)(ctx->dFlags + 7);
    if (tok->elemBytes == 1) k_halo_unpack<1><<<g, kThreads, 0, ctx->stream>>>(P, ctx->region, ctx->dSeq, (char *)tok->recv, to, counter, tok->gate);
    else if (tok->elemBytes == 4) k_halo_unpack<4><<<g, kThreads, 0, ctx->stream>>>(P, ctx->region, ctx->dSeq, (char *)tok->recv, to, counter, tok->gate);
    else if (tok->elemBytes == 8) k_halo_unpack<8><<<g, kThreads, 0, ctx->stream>>>(P, ctx->region, ctx->dSeq, (char *)tok->recv, to, counter, tok->gate);
    else k_halo_unpack<24><<<g, kThreads, 0, ctx->stream>>>(P, ctx->region, ctx->dSeq, (char *)tok->recv, to, counter, tok->gate);
    ctx->launches++;
    return AMR_OK;
}
static int haloSwapCells(amr_ctx *ctx, const void *cells, size_t elemBytes, void *recv, const u8 *minus = nullptr) {
    HaloToken tok;
    TRY(haloBegin(ctx, cells, elemBytes, recv, &tok, minus));
    return haloEnd(ctx, &tok);
}
// How the boundary fix-up kernel that follows haloBegin gets the ghost values (GhostLink, p2p.cuh): on the peer
// transport the kernel does the receive side itself (no unpack launch) and, with fuseReduce, the all-reduce of the
// sweep as well; otherwise the values are already in `recvArray` (NCCL, or supplied by the host).
static GhostLink makeLink(amr_ctx *ctx, HaloToken *tok, const void *recvArray, bool fuseReduce) {
    GhostLink G;
    memset(&G, 0, sizeof(G));
    G.seq = ctx->dSeq;
    G.counter = (unsigned *)(ctx->dFlags + 7);
    G.timeoutNs = ctx->peerTimeoutNs;
    G.nranks = ctx->nranks; G.rank = ctx->rank;
    if (tok && tok->pending) {
        G.P = *(const HaloPlan *)ctx->haloPlan;
        G.region = ctx->region;
        tok->pending = false;
    } else {
        G.ghost = (const char *)recvArray;
        G.region = ctx->region;
    }
    G.peer = (fuseReduce && ctx->comm && ctx->nranks > 1 && ctx->usePeer) ? ctx->peerDev : nullptr;
    return G;
}
// the speculative (device-gated) loops need the peer transport; with NCCL the host-synchronised loops run
static inline bool gatedSwapsOk(const amr_ctx *ctx) { return (!ctx->comm || ctx->nranks < 2) ? true : (ctx->usePeer && ctx->havePlan); }

// all-rank reduction of one device word (reduce()/returnReduce()/gMin/gMax of the reference): *word = op over ranks of *src
// (src == nullptr: in place).  op: 0 max(int32), 1 sum(int64), 2 min(f64), 3 max(f64), 4 sum(int32)
static int allReduceWord(amr_ctx *ctx, void *word, int op, const void *src = nullptr, const int *gate = nullptr) {
    if (!ctx->comm || ctx->nranks < 2) {
        if (src && src != word) CK(cudaMemcpyAsync(word, src, (op == 0 || op == 4) ? 4 : 8, cudaMemcpyDeviceToDevice, ctx->stream));
        return AMR_OK;
    }
    if (ctx->usePeer) {
        k_red_all<<<1, kMaxRanks, 0, ctx->stream>>>(ctx->peerDev, ctx->region, ctx->nranks, ctx->rank, ctx->dSeq, op, src, word,
                                                    ctx->peerTimeoutNs, gate);
        ctx->launches++;
        ctx->peerTouched = true;
        return AMR_OK;
    }
    if (gate) FAIL(AMR_ERR_STATE, "gated reduction on the NCCL transport");
    if (!src) src = word;
    if (op == 0) NCK(ncclAllReduce(src, word, 1, ncclInt32, ncclMax, ctx->comm, ctx->stream));
    else if (op == 4) NCK(ncclAllReduce(src, word, 1, ncclInt32, ncclSum, ctx->comm, ctx->stream));
    else if (op == 1) NCK(ncclAllReduce(src, word, 1, ncclInt64, ncclSum, ctx->comm, ctx->stream));
    else NCK(ncclAllReduce(src, word, 1, ncclFloat64, op == 2 ? ncclMin : ncclMax, ctx->comm, ctx->stream));
    return AMR_OK;
}

// read the device flag (after an all-reduce over ranks when a communicator is attached):
// the reduce(nChanged, sumOp) / returnReduce(bool, orOp) sites of the reference.  Every host round trip also
// fetches the peer region's error word (peerErrFetch / peerErrCheck in common.cuh).
#define PEER_ERR_FETCH() TRY(peerErrFetch(ctx))
#define PEER_ERR_CHECK() TRY(peerErrCheck(ctx))

static int globalFlag(amr_ctx *ctx, int slot, int *value) {
    TRY(allReduceWord(ctx, ctx->dFlags + slot, 0));
    CK(cudaMemcpyAsync(ctx->hFlags + slot, ctx->dFlags + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PEER_ERR_FETCH();
    CK(cudaStreamSynchronize(ctx->stream));
    PEER_ERR_CHECK();
    *value = ctx->hFlags[slot];
    return AMR_OK;
}
static int globalSum(amr_ctx *ctx, int slot, int64_t *value) {
    TRY(allReduceWord(ctx, ctx->dCount + slot, 1));
    CK(cudaMemcpyAsync(ctx->hCount + slot, ctx->dCount + slot, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    PEER_ERR_FETCH();
    CK(cudaStreamSynchronize(ctx->stream));
    PEER_ERR_CHECK();
    *value = ctx->hCount[slot];
    return AMR_OK;
}

AMR_API int amr_synchronize(amr_ctx *ctx) {
    if (!ctx) return AMR_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    PEER_ERR_FETCH();
    CK(cudaStreamSynchronize(ctx->stream));
    PEER_ERR_CHECK();
    return AMR_OK;
}

// ============================================================================ mesh state

// coupled-face lists, transport plans: everything amr_mesh_set / amr_mesh_update derive from the boundary faces
// (ctx->bOwner [b] on the device, ctx->patches, ctx->nCoupled are set)
static int meshBoundarySetup(amr_ctx *ctx, const std::vector<u8> &coupled) {
    const int64_t n = ctx->n, b = ctx->b;
    std::vector<label> bOwn, cells;
    if (ctx->nCoupled > 0) {
        // compact lists of the coupled faces and of the cells that own them (host bookkeeping, once per mesh)
        bOwn.resize((size_t)b);
        CK(cudaMemcpy(bOwn.data(), ctx->bOwner, (size_t)b * 4, cudaMemcpyDeviceToHost));
        std::vector<label> faces;
        faces.reserve((size_t)ctx->nCoupled);
        for (int64_t i = 0; i < b; i++) if (coupled[(size_t)i]) { faces.push_back((label)i); cells.push_back(bOwn[(size_t)i]); }
        std::sort(cells.begin(), cells.end());
        cells.erase(std::unique(cells.begin(), cells.end()), cells.end());
        ctx->nCplCells = (int64_t)cells.size();
        TRY(devAlloc(ctx, &ctx->cplFace, (int64_t)faces.size()));
        TRY(devAlloc(ctx, &ctx->cplCell, (int64_t)cells.size()));
        TRY(devAlloc(ctx, &ctx->cellGhost, n + 16));
        CK(cudaMemcpy(ctx->cplFace, faces.data(), faces.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->cplCell, cells.data(), cells.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemsetAsync(ctx->cellGhost, 0, (size_t)n + 16, ctx->stream));
        LAUNCH(k_list_to_flags, gridFor(ctx->nCplCells), kThreads, ctx->nCplCells, (const label *)ctx->cplCell, ctx->cellGhost);
    }
    TRY(peerPlan(ctx));
    TRY(linkPlan(ctx, bOwn, cells));
    return AMR_OK;
}
// per-mesh scratch and resident results; `pending` = a marking remapped through the topology change (amr_remap_refine_cell)
static int meshScratchAlloc(amr_ctx *ctx, u8 *pending, int64_t pendingN) {
    const int64_t n = ctx->n, b = ctx->b, p = ctx->p;
    int64_t m = std::max(n, p);
    ctx->flagCap = m + 16;
    TRY(devAlloc(ctx, &ctx->flagA, m + 16)); TRY(devAlloc(ctx, &ctx->flagB, m + 16)); TRY(devAlloc(ctx, &ctx->flagC, m + 16));
    TRY(devAlloc(ctx, &ctx->lf, n + 16));
    TRY(devAlloc(ctx, &ctx->srcBits, n / 32 + 16));
    TRY(devAlloc(ctx, &ctx->error, n));
    TRY(devAlloc(ctx, &ctx->refineCell, m + 16));
    TRY(devAlloc(ctx, &ctx->ghost8, b + 16));
    TRY(devAlloc(ctx, &ctx->ghost64, b + 2));
    ctx->tileCap = (m + kTile - 1) / kTile + 8;
    TRY(devAlloc(ctx, &ctx->lbStatus, ctx->tileCap));
    CK(cudaMemsetAsync(ctx->lbStatus, 0, (size_t)ctx->tileCap * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->refineCell, 0, (size_t)m + 16, ctx->stream));
    if (pending) {
        CK(cudaMemcpyAsync(ctx->refineCell, pending, (size_t)pendingN, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaMemsetAsync(ctx->ghost8, 0, (size_t)b + 16, ctx->stream));
    CK(cudaMemsetAsync(ctx->ghost64, 0, (size_t)(b + 2) * 8, ctx->stream));
    ctx->refineCellN = n;
    CK(cudaStreamSynchronize(ctx->stream));
    return AMR_OK;
}

AMR_API int amr_mesh_set(amr_ctx *ctx, int64_t n, int64_t f, int64_t b, int64_t p, const int32_t *owner,
                         const int32_t *neighbour, int npatch, const int32_t *patchStart, const int32_t *patchSize,
                         const int32_t *patchType, const int32_t *patchNbrRank) {
    ENTER();
    if (n < 0 || f < 0 || b < 0 || p < 0 || !owner || (f > 0 && !neighbour) || npatch < 0) FAIL(AMR_ERR_ARG, "amr_mesh_set: bad sizes/pointers");
    if (n >= ((int64_t)1 << 31) - 2 - b) FAIL(AMR_ERR_ARG, "amr_mesh_set: label overflow (n + b must fit int32)");
    // validate BEFORE the current mesh is released: a failed call leaves the context as it was
    std::vector<u8> coupled((size_t)(b > 0 ? b : 1), 0), emptyFace((size_t)(b > 0 ? b : 1), 0);
    std::vector<Patch> patches;
    int64_t nCoupled = 0;
    for (int i = 0; i < npatch; i++) {
        if (!patchStart || !patchSize) FAIL(AMR_ERR_ARG, "amr_mesh_set: patch arrays are null");
        Patch q{patchStart[i], patchSize[i], patchType ? patchType[i] : 0, patchNbrRank ? patchNbrRank[i] : -1};
        if (q.size < 0 || q.start < f || (int64_t)q.start + q.size > f + b) FAIL(AMR_ERR_ARG, "amr_mesh_set: patch range outside boundary faces");
        if (q.type == AMR_PATCH_PROCESSOR) {
            if (ctx->comm && (q.nbr < 0 || q.nbr >= ctx->nranks || q.nbr == ctx->rank)) FAIL(AMR_ERR_ARG, "amr_mesh_set: processor patch with a bad neighbour rank");
            for (int32_t k = 0; k < q.size; k++) coupled[(size_t)(q.start - f + k)] = 1;
            nCoupled += q.size;
        } else if (q.type == AMR_PATCH_EMPTY) {
            for (int32_t k = 0; k < q.size; k++) emptyFace[(size_t)(q.start - f + k)] = 1;
        }
        patches.push_back(q);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    hostMark(ctx, nullptr);
    // a marking remapped through the topology change (amr_remap_refine_cell) survives the mesh switch
    u8 *pending = nullptr;
    int64_t pendingN = 0;
    if (ctx->refineCell && ctx->refineCellN == n && n > 0) { pending = ctx->refineCell; pendingN = n; ctx->refineCell = nullptr; }
    struct PendingGuard { u8 *&p; ~PendingGuard() { if (p) cudaFree(p); } } pendingGuard{pending};   // freed on every exit path
    freeMesh(ctx);
    hostMark(ctx, "amr_mesh_set: old mesh released");
    ctx->nTotalCells = -1;
    ctx->anyProtCached = -1;
    ctx->n = n; ctx->f = f; ctx->b = b; ctx->p = p;
    ctx->patches = patches;
    ctx->nCoupled = nCoupled;
    TRY(devAlloc(ctx, &ctx->owner, f + b));
    TRY(devAlloc(ctx, &ctx->neighbour, f));
    TRY(devAlloc(ctx, &ctx->bCoupled, b));
    TRY(devAlloc(ctx, &ctx->bEmpty, b));
    TRY(uploadAuto(ctx, ctx->owner, owner, (size_t)(f + b) * 4));
    if (f) TRY(uploadAuto(ctx, ctx->neighbour, neighbour, (size_t)f * 4));
    if (b) CK(cudaMemcpyAsync(ctx->bCoupled, coupled.data(), (size_t)b, cudaMemcpyHostToDevice, ctx->stream));
    if (b) CK(cudaMemcpyAsync(ctx->bEmpty, emptyFace.data(), (size_t)b, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->bOwner = ctx->owner + f;
    ctx->meshSet = true; ctx->faceLists = true;
    hostMark(ctx, "amr_mesh_set: owner/neighbour uploaded");
    TRY(meshBoundarySetup(ctx, coupled));
    hostMark(ctx, "amr_mesh_set: boundary set-up");
    TRY(sellBuildCellCells(ctx));
    hostMark(ctx, "amr_mesh_set: cell adjacency built");
    TRY(meshScratchAlloc(ctx, pending, pendingN));
    hostMark(ctx, "amr_mesh_set: scratch allocated");
    return AMR_OK;
}

AMR_API int amr_mesh_set_points(amr_ctx *ctx, const int32_t *pcOff, const int32_t *pcCells, const uint8_t *bndPoint) {
    ENTER();
    if (ctx->n == 0 && ctx->p == 0) FAIL(AMR_ERR_STATE, "amr_mesh_set_points: call amr_mesh_set first");
    if (!pcOff || !bndPoint) FAIL(AMR_ERR_ARG, "amr_mesh_set_points: null pointer");
    int64_t p = ctx->p;
    const int32_t *dOff = nullptr, *dIdx = nullptr;
    TRY(stageIn(ctx, pcOff, p + 1, &dOff));
    int32_t nnz = 0;
    CK(cudaMemcpyAsync(&nnz, dOff + p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (nnz < 0 || (nnz > 0 && !pcCells)) FAIL(AMR_ERR_ARG, "amr_mesh_set_points: bad pointCells CSR");
    TRY(stageIn(ctx, pcCells, nnz, &dIdx));
    TRY(sellBuildFromCsr(ctx, ctx->pc, p, dOff, dIdx));
    devFree(ctx, &ctx->bndPoint);
    TRY(devAlloc(ctx, &ctx->bndPoint, p));
    if (p) TRY(uploadAuto(ctx, ctx->bndPoint, bndPoint, (size_t)p));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->havePoints = true;
    ctx->pcClassDirty = true; ctx->balancedLocal = -1; ctx->balancedGlobal = -1; ctx->spDirty = true; ctx->polyBadKind = -1;
    return AMR_OK;
}

static int growFlagPool(amr_ctx *ctx, int64_t need);
// ---- N1 (second half): incremental hand-over of the addressing after a topology change ------------------------------
// hexRef::updateMesh (hexRef.C:2422-2608) and fvMeshHexRefiner::updateMesh (fvMeshHexRefiner.C:229-240) renumber what
// the refiner keeps through mapPolyMesh; the addressing itself (cells(), pointCells()) is rebuilt by OpenFOAM on demand.
// amr_mesh_set is that rebuild for the device: it uploads the whole mesh.  amr_mesh_update carries the adjacency the
// context already holds through the change instead: the host sends the maps it has anyway and the rows that changed.
static int parsePatches(amr_ctx *ctx, int64_t f, int64_t b, int npatch, const int32_t *patchStart, const int32_t *patchSize, const int32_t *patchType,
                        const int32_t *patchNbrRank, std::vector<Patch> &patches, std::vector<u8> &coupled, std::vector<u8> &emptyFace, int64_t &nCoupled) {
    coupled.assign((size_t)(b > 0 ? b : 1), 0); emptyFace.assign((size_t)(b > 0 ? b : 1), 0);
    nCoupled = 0;
    for (int i = 0; i < npatch; i++) {
        if (!patchStart || !patchSize) FAIL(AMR_ERR_ARG, "patch arrays are null");
        Patch q{patchStart[i], patchSize[i], patchType ? patchType[i] : 0, patchNbrRank ? patchNbrRank[i] : -1};
        if (q.size < 0 || q.start < f || (int64_t)q.start + q.size > f + b) FAIL(AMR_ERR_ARG, "patch range outside boundary faces");
        if (q.type == AMR_PATCH_PROCESSOR) {
            if (ctx->comm && (q.nbr < 0 || q.nbr >= ctx->nranks || q.nbr == ctx->rank)) FAIL(AMR_ERR_ARG, "processor patch with a bad neighbour rank");
            for (int32_t k = 0; k < q.size; k++) coupled[(size_t)(q.start - f + k)] = 1;
            nCoupled += q.size;
        } else if (q.type == AMR_PATCH_EMPTY) {
            for (int32_t k = 0; k < q.size; k++) emptyFace[(size_t)(q.start - f + k)] = 1;
        }
        patches.push_back(q);
    }
    return AMR_OK;
}
static int fetchInt(amr_ctx *ctx, const int32_t *ptrHostOrDev, int64_t index, int32_t *out) {
    if (isDevicePtr(ptrHostOrDev)) {
        CK(cudaMemcpyAsync(out, ptrHostOrDev + index, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else *out = ptrHostOrDev[index];
    return AMR_OK;
}

AMR_API int amr_mesh_update(amr_ctx *ctx, int64_t nNew, int64_t fNew, int64_t bNew, int64_t pNew, const int32_t *cellMap,
                            const int32_t *reverseCellMap, int64_t nDirtyCells, const int32_t *dirtyCells, const int32_t *dirtyCellOff,
                            const int32_t *dirtyCellNbrs, const int32_t *dirtyCellLevel, const int32_t *dirtyParentIndex,
                            const int32_t *boundaryOwner, int npatch, const int32_t *patchStart, const int32_t *patchSize,
                            const int32_t *patchType, const int32_t *patchNbrRank) {
    ENTER();
    if (!ctx->meshSet || !ctx->cc.col) FAIL(AMR_ERR_STATE, "amr_mesh_update: no mesh to update (amr_mesh_set first)");
    if (nNew < 0 || fNew < 0 || bNew < 0 || pNew < 0 || nDirtyCells < 0 || npatch < 0 || (nNew > 0 && !cellMap) || (bNew > 0 && !boundaryOwner))
        FAIL(AMR_ERR_ARG, "amr_mesh_update: bad sizes/pointers");
    if (nDirtyCells > 0 && (!dirtyCells || !dirtyCellOff || !dirtyCellNbrs)) FAIL(AMR_ERR_ARG, "amr_mesh_update: dirty rows are null");
    if (nNew >= ((int64_t)1 << 31) - 2 - bNew) FAIL(AMR_ERR_ARG, "amr_mesh_update: label overflow (n + b must fit int32)");
    if (ctx->haveLevels && nDirtyCells > 0 && !dirtyCellLevel) FAIL(AMR_ERR_ARG, "amr_mesh_update: the context holds cell levels: dirtyCellLevel is needed");
    const int64_t nOld = ctx->n;
    std::vector<u8> coupled, emptyFace;
    std::vector<Patch> patches;
    int64_t nCoupled = 0;
    TRY(parsePatches(ctx, fNew, bNew, npatch, patchStart, patchSize, patchType, patchNbrRank, patches, coupled, emptyFace, nCoupled));
    hostMark(ctx, nullptr);
    // inputs (host pointers are staged; the arena lives until the next call)
    const int32_t *dMap = nullptr, *dRevIn = nullptr, *dDirty = nullptr, *dOff = nullptr, *dNbr = nullptr, *dLvl = nullptr, *dPar = nullptr, *dBOwn = nullptr;
    int32_t nnz = 0;
    if (nDirtyCells > 0) TRY(fetchInt(ctx, dirtyCellOff, nDirtyCells, &nnz));
    if (nnz < 0) FAIL(AMR_ERR_ARG, "amr_mesh_update: bad dirty-row CSR");
    TRY(stageIn(ctx, cellMap, nNew, &dMap));
    TRY(stageIn(ctx, reverseCellMap, nOld, &dRevIn));
    TRY(stageIn(ctx, dirtyCells, nDirtyCells, &dDirty));
    TRY(stageIn(ctx, dirtyCellOff, nDirtyCells + 1, &dOff));
    TRY(stageIn(ctx, dirtyCellNbrs, (int64_t)nnz, &dNbr));
    TRY(stageIn(ctx, dirtyCellLevel, nDirtyCells, &dLvl));
    TRY(stageIn(ctx, dirtyParentIndex, nDirtyCells, &dPar));
    TRY(stageIn(ctx, boundaryOwner, bNew, &dBOwn));
    CK(cudaStreamSynchronize(ctx->stream));
    hostMark(ctx, "amr_mesh_update: inputs staged");
    // what survives of the old mesh
    Sell ccOld = ctx->cc; ctx->cc = Sell();
    amr_ctx::Carry keep;
    keep.valid = ctx->havePoints;
    keep.pcOld = ctx->pc; ctx->pc = Sell();
    keep.bndPointOld = ctx->bndPoint; ctx->bndPoint = nullptr;
    keep.pointLevelOld = ctx->pointLevel; ctx->pointLevel = nullptr;
    label *lvlOld = ctx->lvl, *parentOld = ctx->parentIdx;
    ctx->lvl = nullptr; ctx->parentIdx = nullptr;
    const bool hadLevels = ctx->haveLevels, hadHistory = ctx->haveHistory;
    u8 *pending = nullptr;
    int64_t pendingN = 0;
    if (ctx->refineCell && ctx->refineCellN == nNew && nNew > 0) { pending = ctx->refineCell; pendingN = nNew; ctx->refineCell = nullptr; }
    struct Guard {
        amr_ctx *c; Sell &cc; amr_ctx::Carry &k; label *&l; label *&q; u8 *&pend; bool armed = true;
        ~Guard() { if (pend) cudaFree(pend); if (l) cudaFree(l); if (q) cudaFree(q); sellFree(c, cc);
                   if (armed) { sellFree(c, k.pcOld); if (k.bndPointOld) cudaFree(k.bndPointOld); if (k.pointLevelOld) cudaFree(k.pointLevelOld); if (k.rev) cudaFree(k.rev); } }
    } guard{ctx, ccOld, keep, lvlOld, parentOld, pending};
    freeMesh(ctx);
    hostMark(ctx, "amr_mesh_update: old mesh released");
    ctx->nTotalCells = -1;
    ctx->anyProtCached = -1;
    ctx->n = nNew; ctx->f = fNew; ctx->b = bNew; ctx->p = pNew;
    ctx->patches = patches;
    ctx->nCoupled = nCoupled;
    TRY(devAlloc(ctx, &ctx->bOwner, bNew));
    ctx->bOwnerOwned = true;
    TRY(devAlloc(ctx, &ctx->bCoupled, bNew));
    TRY(devAlloc(ctx, &ctx->bEmpty, bNew));
    if (bNew) {
        CK(cudaMemcpyAsync(ctx->bOwner, dBOwn, (size_t)bNew * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->bCoupled, coupled.data(), (size_t)bNew, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->bEmpty, emptyFace.data(), (size_t)bNew, cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->faceLists = false;                       // (meshSet only once everything below has succeeded: a failed update leaves
                                                  //  a context that asks for amr_mesh_set)
    TRY(meshBoundarySetup(ctx, coupled));
    // the renumbering of the cells stays on the device for the point rows (amr_mesh_update_points)
    if (dRevIn && keep.valid) {
        CK(cudaMalloc((void **)&keep.rev, (size_t)(nOld > 0 ? nOld : 1) * 4));
        CK(cudaMemcpyAsync(keep.rev, dRevIn, (size_t)nOld * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    int32_t *dirtyIdx = nullptr;
    CK(cudaMalloc((void **)&dirtyIdx, (size_t)(nNew > 0 ? nNew : 1) * 4));
    struct IdxGuard { int32_t *p; ~IdxGuard() { cudaFree(p); } } idxGuard{dirtyIdx};
    hostMark(ctx, "amr_mesh_update: boundary set-up");
    TRY(sellCarryOver(ctx, ctx->cc, ccOld, nNew, dMap, dRevIn, nDirtyCells, dDirty, dOff, dNbr, nNew, 1, true, dirtyIdx));
    hostMark(ctx, "amr_mesh_update: cell adjacency carried over");
    if (cudaMalloc((void **)&ctx->cc.sliceGhost, (size_t)ctx->cc.slices + 16) == cudaSuccess) {
        cudaMemsetAsync(ctx->cc.sliceGhost, 0, (size_t)ctx->cc.slices + 16, ctx->stream);
        if (nNew > 0 && ctx->nCoupled > 0 && ctx->cellGhost) LAUNCH(k_slice_ghost_cells, gridFor(nNew), kThreads, (const u8 *)ctx->cellGhost, nNew, ctx->cc.sliceGhost);
    }
    // cellLevel_ and the history's parent index through the change (hexRef.C:2455-2485): kept for the clean cells
    if (hadLevels && lvlOld) {
        TRY(devAlloc(ctx, &ctx->lvl, nNew)); TRY(devAlloc(ctx, &ctx->lvl8, nNew + 16)); TRY(devAlloc(ctx, &ctx->parentIdx, nNew));
        if (nNew) {
            LAUNCH(k_upd_values<label>, gridFor(nNew), kThreads, nNew, dMap, (const int32_t *)dirtyIdx, dLvl, (const label *)lvlOld, ctx->lvl);
            LAUNCH(k_lvl_to_u8, gridFor(nNew), kThreads, nNew, ctx->lvl, ctx->lvl8);
            if (hadHistory && parentOld && (dPar || nDirtyCells == 0))
                LAUNCH(k_upd_values<label>, gridFor(nNew), kThreads, nNew, dMap, (const int32_t *)dirtyIdx, dPar, (const label *)parentOld, ctx->parentIdx);
            else CK(cudaMemsetAsync(ctx->parentIdx, 0xFF, (size_t)nNew * 4, ctx->stream));
        }
        ctx->haveLevels = true;
        ctx->haveHistory = hadHistory && parentOld && (dPar || nDirtyCells == 0);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    hostMark(ctx, "amr_mesh_update: levels");
    TRY(meshScratchAlloc(ctx, pending, pendingN));
    hostMark(ctx, "amr_mesh_update: scratch allocated");
    ctx->carry = keep;
    guard.armed = false;
    ctx->meshSet = true;
    ctx->pcClassDirty = true; ctx->balancedLocal = -1; ctx->balancedGlobal = -1; ctx->spDirty = true;
    return AMR_OK;
}

AMR_API int amr_mesh_update_points(amr_ctx *ctx, const int32_t *pointMap, int64_t nDirtyPoints, const int32_t *dirtyPoints,
                                   const int32_t *dirtyPointOff, const int32_t *dirtyPointCells, const uint8_t *dirtyBndPoint,
                                   const int32_t *dirtyPointLevel) {
    ENTER();
    if (!ctx->meshSet || !ctx->carry.valid) FAIL(AMR_ERR_STATE, "amr_mesh_update_points: call amr_mesh_update first (on a context that held point addressing)");
    const int64_t p = ctx->p;
    if ((p > 0 && !pointMap) || nDirtyPoints < 0) FAIL(AMR_ERR_ARG, "amr_mesh_update_points: bad arguments");
    if (nDirtyPoints > 0 && (!dirtyPoints || !dirtyPointOff || !dirtyPointCells || !dirtyBndPoint)) FAIL(AMR_ERR_ARG, "amr_mesh_update_points: dirty rows are null");
    amr_ctx::Carry &K = ctx->carry;
    if (K.pointLevelOld && nDirtyPoints > 0 && !dirtyPointLevel) FAIL(AMR_ERR_ARG, "amr_mesh_update_points: the context holds point levels: dirtyPointLevel is needed");
    const int32_t *dMap = nullptr, *dDirty = nullptr, *dOff = nullptr, *dCells = nullptr, *dPl = nullptr;
    const u8 *dBnd = nullptr;
    int32_t nnz = 0;
    if (nDirtyPoints > 0) TRY(fetchInt(ctx, dirtyPointOff, nDirtyPoints, &nnz));
    if (nnz < 0) FAIL(AMR_ERR_ARG, "amr_mesh_update_points: bad dirty-row CSR");
    TRY(stageIn(ctx, pointMap, p, &dMap));
    TRY(stageIn(ctx, dirtyPoints, nDirtyPoints, &dDirty));
    TRY(stageIn(ctx, dirtyPointOff, nDirtyPoints + 1, &dOff));
    TRY(stageIn(ctx, dirtyPointCells, (int64_t)nnz, &dCells));
    TRY(stageIn(ctx, dirtyBndPoint, nDirtyPoints, &dBnd));
    TRY(stageIn(ctx, dirtyPointLevel, nDirtyPoints, &dPl));
    int32_t *dirtyIdx = nullptr;
    CK(cudaMalloc((void **)&dirtyIdx, (size_t)(p > 0 ? p : 1) * 4));
    struct IdxGuard { int32_t *p; ~IdxGuard() { cudaFree(p); } } idxGuard{dirtyIdx};
    sellFree(ctx, ctx->pc);
    TRY(sellCarryOver(ctx, ctx->pc, K.pcOld, p, dMap, (const label *)K.rev, nDirtyPoints, dDirty, dOff, dCells, ctx->n, 0, false, dirtyIdx));
    devFree(ctx, &ctx->bndPoint);
    TRY(devAlloc(ctx, &ctx->bndPoint, p));
    if (p) LAUNCH(k_upd_values<u8>, gridFor(p), kThreads, p, dMap, (const int32_t *)dirtyIdx, dBnd, (const u8 *)K.bndPointOld, ctx->bndPoint);
    devFree(ctx, &ctx->pointLevel);
    if (K.pointLevelOld) {
        TRY(devAlloc(ctx, &ctx->pointLevel, p));
        if (p) LAUNCH(k_upd_values<label>, gridFor(p), kThreads, p, dMap, (const int32_t *)dirtyIdx, dPl, (const label *)K.pointLevelOld, ctx->pointLevel);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    sellFree(ctx, K.pcOld); devFree(ctx, &K.bndPointOld); devFree(ctx, &K.pointLevelOld); devFree(ctx, &K.rev);
    K.valid = false;
    if (p + 16 > ctx->flagCap) TRY(growFlagPool(ctx, p + 16));
    ctx->havePoints = true;
    ctx->pcClassDirty = true; ctx->balancedLocal = -1; ctx->balancedGlobal = -1; ctx->spDirty = true; ctx->polyBadKind = -1;
    return AMR_OK;
}

// grow the pointer-swapped flag pool (flagA/B/C + refineCell) to hold `need` bytes each
static int growFlagPool(amr_ctx *ctx, int64_t need) {
    if (need <= ctx->flagCap) return AMR_OK;
    u8 **bufs[4] = {&ctx->flagA, &ctx->flagB, &ctx->flagC, &ctx->refineCell};
    for (auto pb : bufs) {
        u8 *nb = nullptr;
        CK(cudaMalloc((void **)&nb, (size_t)need));
        CK(cudaMemsetAsync(nb, 0, (size_t)need, ctx->stream));
        if (*pb) { CK(cudaMemcpyAsync(nb, *pb, (size_t)ctx->flagCap, cudaMemcpyDeviceToDevice, ctx->stream)); }
        CK(cudaStreamSynchronize(ctx->stream));
        if (*pb) cudaFree(*pb);
        *pb = nb;
        ctx->devBytes += need - ctx->flagCap;
    }
    ctx->flagCap = need;
    int64_t tiles = (need + kTile - 1) / kTile + 8;
    if (tiles > ctx->tileCap) {
        devFree(ctx, &ctx->lbStatus);
        ctx->tileCap = tiles;
        TRY(devAlloc(ctx, &ctx->lbStatus, ctx->tileCap));
        CK(cudaMemsetAsync(ctx->lbStatus, 0, (size_t)ctx->tileCap * 8, ctx->stream));
    }
    return AMR_OK;
}

AMR_API int amr_mesh_set_edges(amr_ctx *ctx, int64_t nEdges, const int32_t *edgePoints, const int32_t *ecOff,
                               const int32_t *ecCells, const uint8_t *bndEdge) {
    ENTER();
    if (!ctx->meshSet) FAIL(AMR_ERR_STATE, "amr_mesh_set_edges: call amr_mesh_set first");
    if (nEdges < 0 || !edgePoints || !ecOff || !bndEdge) FAIL(AMR_ERR_ARG, "amr_mesh_set_edges: bad arguments");
    const int32_t *dOff = nullptr, *dIdx = nullptr;
    TRY(stageIn(ctx, ecOff, nEdges + 1, &dOff));
    int32_t nnz = 0;
    CK(cudaMemcpyAsync(&nnz, dOff + nEdges, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (nnz < 0 || (nnz > 0 && !ecCells)) FAIL(AMR_ERR_ARG, "amr_mesh_set_edges: bad edgeCells CSR");
    TRY(stageIn(ctx, ecCells, nnz, &dIdx));
    TRY(sellBuildFromCsr(ctx, ctx->ec, nEdges, dOff, dIdx));
    devFree(ctx, &ctx->bndEdge); devFree(ctx, &ctx->edgePts);
    TRY(devAlloc(ctx, &ctx->bndEdge, nEdges)); TRY(devAlloc(ctx, &ctx->edgePts, 2 * nEdges));
    if (nEdges) {
        TRY(uploadAuto(ctx, ctx->bndEdge, bndEdge, (size_t)nEdges));
        TRY(uploadAuto(ctx, ctx->edgePts, edgePoints, (size_t)nEdges * 8));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    TRY(growFlagPool(ctx, nEdges + 16));
    ctx->nEdges = nEdges;
    ctx->haveEdges = true;
    ctx->polyBadKind = -1;
    sellFree(ctx, ctx->ce); ctx->ceState = 0; ctx->balancedEdges = -1;
    return AMR_OK;
}

AMR_API int amr_mesh_set_faces(amr_ctx *ctx, const int32_t *faceOff, const int32_t *facePts) {
    ENTER();
    if (!ctx->meshSet) FAIL(AMR_ERR_STATE, "amr_mesh_set_faces: call amr_mesh_set first");
    if (!faceOff || !facePts) FAIL(AMR_ERR_ARG, "amr_mesh_set_faces: null pointer");
    const int64_t nf = ctx->f + ctx->b;
    devFree(ctx, &ctx->faceOff); devFree(ctx, &ctx->facePts);
    TRY(devAlloc(ctx, &ctx->faceOff, nf + 1));
    CK(cudaMemcpyAsync(ctx->faceOff, faceOff, (size_t)(nf + 1) * 4, cudaMemcpyDefault, ctx->stream));
    int32_t nnz = 0;
    CK(cudaMemcpyAsync(&nnz, ctx->faceOff + nf, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (nnz < 0) FAIL(AMR_ERR_ARG, "amr_mesh_set_faces: bad CSR");
    TRY(devAlloc(ctx, &ctx->facePts, nnz));
    if (nnz) TRY(uploadAuto(ctx, ctx->facePts, facePts, (size_t)nnz * 4));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->haveFaces = true;
    ctx->polyBadKind = -1;
    return AMR_OK;
}

// per-mesh plan of the POINT swaps (syncTools::syncPointList sites): like peerPlan, over the patches' point lists.
// Collective: every rank of the communicator calls amr_mesh_set_patch_points (with empty lists when it has none).
static int pointPlan(amr_ctx *ctx) {
    ctx->havePointPlan = false;
    if (!ctx->comm || ctx->nranks < 2 || !ctx->usePeer) return AMR_OK;
    HaloPlan &P = *(HaloPlan *)ctx->pointPlanPtr;
    P.nPatch = 0; P.prefix[0] = 0;
    int ok = ctx->havePlan ? 1 : 0;
    std::vector<int32_t> mine, nbrs;
    const size_t npatch = ctx->patches.size();
    for (size_t k = 0; k < npatch; k++) {
        const Patch &q = ctx->patches[k];
        const int32_t cnt = ctx->ppOff[k + 1] - ctx->ppOff[k];
        if (q.type != AMR_PATCH_PROCESSOR || q.size == 0) continue;       // same patches as peerPlan's send/recv pairs
        mine.push_back(ctx->ppOff[k]); mine.push_back(cnt); nbrs.push_back(q.nbr);
    }
    const int np = (int)nbrs.size();
    if (np > 32) ok = 0;
    if ((int64_t)ctx->ppOff[npatch] * 8 > kHaloCap) ok = 0;
    std::vector<int32_t> theirs((size_t)(2 * np > 0 ? 2 * np : 1), 0);
    void *q;
    TRY(arenaAlloc(ctx, &q, (size_t)(2 * np + 4) * 4)); int32_t *dMine = (int32_t *)q;
    TRY(arenaAlloc(ctx, &q, (size_t)(2 * np + 4) * 4)); int32_t *dTheirs = (int32_t *)q;
    if (np) CK(cudaMemcpyAsync(dMine, mine.data(), (size_t)2 * np * 4, cudaMemcpyHostToDevice, ctx->stream));
    NCK(ncclGroupStart());
    for (int k = 0; k < np; k++) {
        NCK(ncclSend(dMine + 2 * k, 2, ncclInt32, nbrs[(size_t)k], ctx->comm, ctx->stream));
        NCK(ncclRecv(dTheirs + 2 * k, 2, ncclInt32, nbrs[(size_t)k], ctx->comm, ctx->stream));
    }
    NCK(ncclGroupEnd());
    if (np) CK(cudaMemcpyAsync(theirs.data(), dTheirs, (size_t)2 * np * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < np; k++)
        if (theirs[(size_t)2 * k + 1] != mine[(size_t)2 * k + 1]) ok = 0;     // the two sides list different numbers of points
    int32_t *dAgree = dMine + 2 * np;
    CK(cudaMemcpyAsync(dAgree, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    NCK(ncclAllReduce(dAgree, dAgree, 1, ncclInt32, ncclMin, ctx->comm, ctx->stream));
    CK(cudaMemcpyAsync(&ok, dAgree, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (!ok) return AMR_OK;                          // agreed: NCCL transport for the point swaps of this mesh
    for (int k = 0; k < np; k++) {
        P.nbr[k] = nbrs[(size_t)k]; P.offMine[k] = mine[(size_t)2 * k]; P.size[k] = mine[(size_t)2 * k + 1];
        P.offTheirs[k] = theirs[(size_t)2 * k]; P.prefix[k + 1] = P.prefix[k] + P.size[k];
    }
    P.nPatch = np;
    ctx->havePointPlan = true;
    return AMR_OK;
}

AMR_API int amr_mesh_set_patch_points(amr_ctx *ctx, const int32_t *patchPointOff, const int32_t *patchPoints) {
    ENTER();
    if (!ctx->meshSet) FAIL(AMR_ERR_STATE, "amr_mesh_set_patch_points: call amr_mesh_set first");
    if (!patchPointOff) FAIL(AMR_ERR_ARG, "amr_mesh_set_patch_points: null pointer");
    const size_t np = ctx->patches.size();
    ctx->ppOff.assign(patchPointOff, patchPointOff + np + 1);
    for (size_t q = 0; q < np; q++)
        if (ctx->ppOff[q + 1] < ctx->ppOff[q] || ctx->ppOff[0] != 0) { ctx->ppOff.clear(); FAIL(AMR_ERR_ARG, "amr_mesh_set_patch_points: bad offsets"); }
    const int64_t tot = ctx->ppOff[np];
    if (tot > 0 && !patchPoints) { ctx->ppOff.clear(); FAIL(AMR_ERR_ARG, "amr_mesh_set_patch_points: null pointer"); }
    devFree(ctx, &ctx->ppPts);
    TRY(devAlloc(ctx, &ctx->ppPts, tot > 0 ? tot : 1));
    if (tot) CK(cudaMemcpyAsync(ctx->ppPts, patchPoints, (size_t)tot * 4, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return pointPlan(ctx);
}

// swap of a per-point array over the processor patches' point lists: recv[i] = the neighbour's value at the point that
// pairs with my i-th patch point (processorPolyPatch::meshPoints() against neighbPoints()).  Peer memory when the ranks
// agreed on it for this mesh, else grouped ncclSend / ncclRecv.  recv is laid out like the patch point lists.
static int pointSwap(amr_ctx *ctx, const void *values, size_t elemBytes, void *recv) {
    const size_t np = ctx->ppOff.empty() ? 0 : ctx->patches.size();
    const int64_t tot = np ? ctx->ppOff[np] : 0;
    if (ctx->havePointPlan) {
        const HaloPlan &P = *(const HaloPlan *)ctx->pointPlanPtr;
        const unsigned g = gridFor(P.prefix[P.nPatch]);
        unsigned *c1 = (unsigned *)(ctx->dFlags + 6), *c2 = (unsigned *)(ctx->dFlags + 7);
        const unsigned long long to = ctx->peerTimeoutNs;
        if (elemBytes == 1) {
            k_halo_push_cells<1><<<gridFor(pushWords(P)), kThreads, 0, ctx->stream>>>(P, ctx->peerDev, ctx->rank, ctx->dSeq, (const label *)ctx->ppPts, (const char *)values, (const u8 *)nullptr, c1, (const int *)nullptr);
            k_halo_unpack<1><<<g, kThreads, 0, ctx->stream>>>(P, ctx->region, ctx->dSeq, (char *)recv, to, c2, (const int *)nullptr);
        } else {
            k_halo_push_cells<4><<<g, kThreads, 0, ctx->stream>>>(P, ctx->peerDev, ctx->rank, ctx->dSeq, (const label *)ctx->ppPts, (const char *)values, (const u8 *)nullptr, c1, (const int *)nullptr);
            k_halo_unpack<4><<<g, kThreads, 0, ctx->stream>>>(P, ctx->region, ctx->dSeq, (char *)recv, to, c2, (const int *)nullptr);
        }
        ctx->launches += 2;
        ctx->peerTouched = true;
        return AMR_OK;
    }
    void *q;
    TRY(arenaAlloc(ctx, &q, (size_t)(tot > 0 ? tot : 1) * elemBytes));
    char *send = (char *)q;
    if (tot) {
        if (elemBytes == 1) LAUNCH(k_pack_u8, gridFor(tot), kThreads, tot, (const label *)ctx->ppPts, (const u8 *)values, (u8 *)send);
        else LAUNCH(k_pack_i32, gridFor(tot), kThreads, tot, (const label *)ctx->ppPts, (const label *)values, (label *)send);
    }
    NCK(ncclGroupStart());
    for (size_t k = 0; k < np; k++) {
        const Patch &pp = ctx->patches[k];
        const int64_t cnt = ctx->ppOff[k + 1] - ctx->ppOff[k];
        if (pp.type != AMR_PATCH_PROCESSOR || cnt == 0) continue;
        NCK(ncclSend(send + (size_t)ctx->ppOff[k] * elemBytes, (size_t)cnt * elemBytes, ncclUint8, pp.nbr, ctx->comm, ctx->stream));
        NCK(ncclRecv((char *)recv + (size_t)ctx->ppOff[k] * elemBytes, (size_t)cnt * elemBytes, ncclUint8, pp.nbr, ctx->comm, ctx->stream));
    }
    NCK(ncclGroupEnd());
    return AMR_OK;
}

// syncTools::syncPointList(pflag, orEqOp<bool>(), false) over the processor patches (fvMeshRefiner.C:952-958).  One
// pairwise exchange covers points shared by two ranks; a point on an edge or corner of the decomposition reaches the
// ranks it shares no FACE with through a common neighbour, so the exchange repeats until no rank changed (OR is
// idempotent: same fixed point as the reference's global shared-point pass).  Collective.
static int syncPointsOr(amr_ctx *ctx, u8 *pflag) {
    if (!ctx->comm || ctx->nranks < 2) return AMR_OK;
    if (ctx->nCoupled > 0 && ctx->ppOff.empty())
        FAIL(AMR_ERR_STATE, "point sync across processor patches needs their point lists (amr_mesh_set_patch_points)");
    const size_t np = ctx->ppOff.empty() ? 0 : ctx->patches.size();
    const int64_t tot = np ? ctx->ppOff[np] : 0;
    void *q;
    TRY(arenaAlloc(ctx, &q, (size_t)(tot > 0 ? tot : 1)));
    u8 *recv = (u8 *)q;
    for (int round = 0; round < 64; round++) {
        CK(cudaMemsetAsync(ctx->dFlags + 11, 0, sizeof(int), ctx->stream));
        TRY(pointSwap(ctx, pflag, 1, recv));
        if (tot) LAUNCH(k_or_scatter, gridFor(tot), kThreads, tot, (const label *)ctx->ppPts, (const u8 *)recv, pflag, ctx->dFlags + 11);
        int changed = 0;
        TRY(globalFlag(ctx, 11, &changed));
        if (!changed) return AMR_OK;
    }
    FAIL(AMR_ERR_INVARIANT, "point sync did not settle");
}

// "Check pointLevel is synchronized" of hexRef::checkRefinementLevels (hexRef.C:3027-3057): the reference min-syncs a
// copy of pointLevel_ over the coupled points and aborts where the copy differs from the original, i.e. where the
// ranks sharing a point do not all hold the same level.  Every sharer of a processor-patch point is linked to
// another one by a patch that lists the point, so one pairwise swap over the patches' point lists finds every
// disagreement.  Counts the points of THIS rank whose level differs from a neighbour's; collective.
__global__ void k_point_level_mismatch(int64_t tot, const label *__restrict__ ppPts, const label *__restrict__ pointLevel,
                                       const label *__restrict__ recv, u8 *__restrict__ badPoint) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < tot && recv[i] != pointLevel[ppPts[i]]) badPoint[ppPts[i]] = 1;
}
AMR_API int amr_check_point_levels(amr_ctx *ctx, const int32_t *pointLevel, int64_t *nBad) {
    ENTER();
    if (nBad) *nBad = 0;
    const int64_t p = ctx->p;
    const int32_t *dPl = nullptr;
    TRY(stageIn(ctx, pointLevel, p, &dPl));
    if (!dPl) dPl = ctx->pointLevel;
    if (!dPl) FAIL(AMR_ERR_STATE, "amr_check_point_levels: no pointLevel (amr_levels_set)");
    if (!ctx->comm || ctx->nranks < 2) return AMR_OK;            // nothing is coupled on a single rank
    if (ctx->nCoupled > 0 && ctx->ppOff.empty())
        FAIL(AMR_ERR_STATE, "amr_check_point_levels: the processor patches' point lists are needed (amr_mesh_set_patch_points)");
    const size_t np = ctx->ppOff.empty() ? 0 : ctx->patches.size();
    const int64_t tot = np ? ctx->ppOff[np] : 0;
    void *q;
    TRY(arenaAlloc(ctx, &q, (size_t)(tot > 0 ? tot : 1) * 4));
    label *recv = (label *)q;
    u8 *bad = ctx->flagC;
    if (p) CK(cudaMemsetAsync(bad, 0, (size_t)p, ctx->stream));
    TRY(pointSwap(ctx, dPl, 4, recv));
    if (tot) LAUNCH(k_point_level_mismatch, gridFor(tot), kThreads, tot, (const label *)ctx->ppPts, dPl, (const label *)recv, bad);
    CK(cudaMemsetAsync(ctx->dCount, 0, 8, ctx->stream));
    if (p) LAUNCH(k_count_flags, gridFor(p), kThreads, p, (const u8 *)bad, (unsigned long long *)ctx->dCount);
    CK(cudaMemcpyAsync(ctx->hCount, ctx->dCount, 8, cudaMemcpyDeviceToHost, ctx->stream));   // this rank's count, before the reduction
    int64_t total = 0;
    TRY(globalSum(ctx, 0, &total));
    if (nBad) *nBad = ctx->hCount[0] > total ? total : ctx->hCount[0];
    CK(cudaGetLastError());
    if (total > 0) FAIL(AMR_ERR_INVARIANT, "PointLevel is not consistent across coupled patches. (hexRef.C:3047)");
    return AMR_OK;
}

AMR_API int amr_set_option(amr_ctx *ctx, int option, int64_t value) {
    if (!ctx) return AMR_ERR_ARG;
    if (option == AMR_OPT_EDGE_BASED_CONSISTENCY) ctx->optEdgeConsistency = value ? 1 : 0;
    else if (option == AMR_OPT_POLY_FORCE) ctx->optPolyForce = value ? 1 : 0;
    else { ctx->err = "amr_set_option: unknown option"; return AMR_ERR_ARG; }
    return AMR_OK;
}

AMR_API int amr_mesh_set_geometry(amr_ctx *ctx, const double *weights, const double *Sf, const double *magSf,
                                  const double *nonOrthDeltaCoeffs, const double *V) {
    ENTER();
    if (!ctx->meshSet) FAIL(AMR_ERR_STATE, "amr_mesh_set_geometry: call amr_mesh_set first");
    if (!ctx->faceLists) FAIL(AMR_ERR_STATE, "amr_mesh_set_geometry: needs owner/neighbour (amr_mesh_update hands over the cell adjacency only; face lists come with amr_mesh_set)");
    if (!weights || !Sf || !magSf || !nonOrthDeltaCoeffs || !V) FAIL(AMR_ERR_ARG, "amr_mesh_set_geometry: null pointer");
    const int64_t n = ctx->n, f = ctx->f, b = ctx->b, nf = f + b;
    devFree(ctx, &ctx->gWeights); devFree(ctx, &ctx->gSf); devFree(ctx, &ctx->gMagSf); devFree(ctx, &ctx->gDelta); devFree(ctx, &ctx->gV);
    devFree(ctx, &ctx->grad); devFree(ctx, &ctx->xbf);
    TRY(devAlloc(ctx, &ctx->gWeights, nf)); TRY(devAlloc(ctx, &ctx->gSf, 3 * nf)); TRY(devAlloc(ctx, &ctx->gMagSf, nf));
    TRY(devAlloc(ctx, &ctx->gDelta, nf)); TRY(devAlloc(ctx, &ctx->gV, n)); TRY(devAlloc(ctx, &ctx->grad, 3 * n)); TRY(devAlloc(ctx, &ctx->xbf, b));
    TRY(uploadAuto(ctx, ctx->gWeights, weights, (size_t)nf * 8));
    TRY(uploadAuto(ctx, ctx->gSf, Sf, (size_t)nf * 24));
    TRY(uploadAuto(ctx, ctx->gMagSf, magSf, (size_t)nf * 8));
    TRY(uploadAuto(ctx, ctx->gDelta, nonOrthDeltaCoeffs, (size_t)nf * 8));
    if (n) TRY(uploadAuto(ctx, ctx->gV, V, (size_t)n * 8));
    // cell -> faces SELL, rows ascending in face label
    sellFree(ctx, ctx->cf);
    int32_t *deg = nullptr, *cursor = nullptr;
    CK(cudaMalloc((void **)&deg, (size_t)(n + 1) * 4));
    CK(cudaMalloc((void **)&cursor, (size_t)(n + 1) * 4));
    CK(cudaMemsetAsync(deg, 0, (size_t)(n + 1) * 4, ctx->stream));
    CK(cudaMemsetAsync(cursor, 0, (size_t)(n + 1) * 4, ctx->stream));
    if (nf > 0) LAUNCH(k_cf_deg, gridFor(nf), kThreads, ctx->owner, ctx->neighbour, f, b, deg);
    int rc = sellAllocFromDeg(ctx, ctx->cf, n, deg, 0);
    if (rc == AMR_OK) {
        if (nf > 0) LAUNCH(k_cf_fill, gridFor(nf), kThreads, ctx->owner, ctx->neighbour, f, b, ctx->cf.sliceOff, cursor, ctx->cf.col);
        if (n > 0) LAUNCH(k_sell_sort_rows, gridFor(n), kThreads, ctx->cf.col, ctx->cf.sliceOff, deg, n);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = std::string("amr_mesh_set_geometry: ") + cudaGetErrorString(e); rc = AMR_ERR_CUDA; }
    }
    cudaFree(deg); cudaFree(cursor);
    if (rc != AMR_OK) return rc;
    ctx->haveGeom = true;
    return AMR_OK;
}

AMR_API int amr_levels_set(amr_ctx *ctx, const int32_t *cellLevel, const int32_t *pointLevel, const int32_t *visibleCells,
                           const int32_t *splitParent, int64_t nSplit) {
    ENTER();
    if (!cellLevel) FAIL(AMR_ERR_ARG, "amr_levels_set: cellLevel is null");
    int64_t n = ctx->n, p = ctx->p;
    if (!ctx->lvl) { TRY(devAlloc(ctx, &ctx->lvl, n)); TRY(devAlloc(ctx, &ctx->lvl8, n + 16)); TRY(devAlloc(ctx, &ctx->parentIdx, n)); }
    if (n) TRY(uploadAuto(ctx, ctx->lvl, cellLevel, (size_t)n * 4));
    if (n) LAUNCH(k_lvl_to_u8, gridFor(n), kThreads, n, ctx->lvl, ctx->lvl8);
    if (pointLevel && p) {
        if (!ctx->pointLevel) TRY(devAlloc(ctx, &ctx->pointLevel, p));
        TRY(uploadAuto(ctx, ctx->pointLevel, pointLevel, (size_t)p * 4));
    }
    ctx->haveHistory = false;
    if (visibleCells) {
        const int32_t *dVis = nullptr, *dPar = nullptr;
        TRY(stageIn(ctx, visibleCells, n, &dVis));
        TRY(stageIn(ctx, splitParent, nSplit, &dPar));
        if (nSplit > 0 && !dPar) FAIL(AMR_ERR_ARG, "amr_levels_set: splitParent is null");
        if (n) LAUNCH(k_parent_index, gridFor(n), kThreads, n, dVis, dPar, nSplit, ctx->parentIdx);
        // the history itself stays resident for the cluster queries of the load-balance constraint
        devFree(ctx, &ctx->visible); devFree(ctx, &ctx->splitParent);
        TRY(devAlloc(ctx, &ctx->visible, n)); TRY(devAlloc(ctx, &ctx->splitParent, nSplit > 0 ? nSplit : 1));
        if (n) CK(cudaMemcpyAsync(ctx->visible, dVis, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        if (nSplit) CK(cudaMemcpyAsync(ctx->splitParent, dPar, (size_t)nSplit * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->nSplit = nSplit;
        ctx->haveHistory = true;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->haveLevels = true;
    ctx->pcClassDirty = true; ctx->balancedLocal = -1; ctx->balancedGlobal = -1; ctx->spDirty = true; ctx->polyBadKind = -1;
    ctx->balancedEdges = -1;
    return AMR_OK;
}

AMR_API int amr_protected_cells_set(amr_ctx *ctx, const uint8_t *protectedCell) {
    ENTER();
    ctx->anyProtCached = -1;
    if (!protectedCell) { devFree(ctx, &ctx->protCell); return AMR_OK; }
    if (!ctx->protCell) TRY(devAlloc(ctx, &ctx->protCell, ctx->n + 16));
    if (ctx->n) TRY(uploadAuto(ctx, ctx->protCell, protectedCell, (size_t)ctx->n));
    CK(cudaStreamSynchronize(ctx->stream));
    return AMR_OK;
}

// ============================================================================ stage 1

static Thresholds mkThr(const double *t) {
    Thresholds r{0, 0, 0, 0};
    if (t) { r.lowerRefine = t[0]; r.upperRefine = t[1]; r.lowerUnrefine = t[2]; r.upperUnrefine = t[3]; }
    return r;
}

AMR_API int amr_estimate(amr_ctx *ctx, int kind, const double *x, const double *nbrValues, double p0, double p1, int scale,
                         const double *thr, double *errorOut) {
    ENTER();
    if (!ctx->meshSet) FAIL(AMR_ERR_STATE, "amr_estimate: mesh not set");
    if (!x) FAIL(AMR_ERR_ARG, "amr_estimate: x is null");
    if (scale && !thr) FAIL(AMR_ERR_ARG, "amr_estimate: scale requested without thresholds");
    const int64_t n = ctx->n, b = ctx->b;
    const double *dx = nullptr;
    TRY(stageIn(ctx, x, n, &dx));
    // a device errorOut is written in place (no second copy); a host errorOut is read back from the
    // resident field, which is what later stages use when they are passed error == NULL
    const bool devOut = errorOut && isDevicePtr(errorOut);
    double *target = devOut ? errorOut : ctx->error;
    ctx->errorResident = !devOut;
    Thresholds t = mkThr(thr);
    if (n > 0) {
        if (kind == AMR_EST_FIELD_VALUE) {
            if (scale) LAUNCH((k_pointwise<true, true>), gridFor(n), kThreads, n, dx, t, target);
            else LAUNCH((k_pointwise<true, false>), gridFor(n), kThreads, n, dx, t, target);
        } else if (kind == AMR_EST_DELTA || kind == AMR_EST_SCALED_DELTA || kind == AMR_EST_CODED_DELTA ||
                   kind == AMR_EST_DENSITY_GRADIENT) {
            const double *ghost = nullptr;
            HaloToken halo;
            if (ctx->nCoupled > 0) {
                if (nbrValues) TRY(stageIn(ctx, nbrValues, b, &ghost));
                else {
                    TRY(haloBegin(ctx, dx, 8, ctx->ghost64, &halo));      // received after the ghost-free main kernel
                    ghost = ctx->ghost64;
                }
            }
            const int32_t *so = ctx->cc.sliceOff, *col = ctx->cc.col;
#define EST(K)                                                                                                   \
    do {                                                                                                         \
        const u8 *cg = ctx->nCoupled > 0 ? (const u8 *)ctx->cellGhost : (const u8 *)nullptr;                     \
        if (pipeOk(ctx, ctx->cc)) {                                                                              \
            if (scale) TRY(launchPipe(ctx, kp_estimate<K, true>, ctx->cc, dx, p0, p1, t, cg, target));            \
            else TRY(launchPipe(ctx, kp_estimate<K, false>, ctx->cc, dx, p0, p1, t, cg, target));                 \
        } else if (scale) LAUNCH((k_estimate<K, true>), gridFor(n), kThreads, n, so, col, dx, cg, p0, p1, t, target); \
        else LAUNCH((k_estimate<K, false>), gridFor(n), kThreads, n, so, col, dx, cg, p0, p1, t, target);        \
        if (ctx->nCoupled > 0) {                                                                                 \
            GhostLink G = makeLink(ctx, &halo, ghost, false);                                                    \
            LAUNCH((k_fix_estimate_link<K>), gridFor(ctx->nCoupled), kThreads, G, ctx->nCoupled, (const label *)ctx->cplFace, \
                   (const label *)ctx->bOwner, dx, p0, p1, target);                                    \
            if (scale) LAUNCH(k_fix_normalize, gridFor(ctx->nCplCells), kThreads, ctx->nCplCells, (const label *)ctx->cplCell, t, target); \
        }                                                                                                        \
    } while (0)
            if (kind == AMR_EST_DELTA) EST(0);
            else if (kind == AMR_EST_SCALED_DELTA || kind == AMR_EST_DENSITY_GRADIENT) EST(1);
            else EST(4);
#undef EST
        } else {
            FAIL(AMR_ERR_ARG, "amr_estimate: estimator kind not available on the device path");
        }
        if (errorOut && !devOut) { CK(cudaMemcpyAsync(errorOut, target, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream)); ctx->hostTouched = true; }
    }
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_estimate_lohner(amr_ctx *ctx, const double *x, const double *xBoundary, const double *nbrValues,
                                double epsilon, int scale, const double *thr, double *errorOut) {
    ENTER();
    if (!ctx->meshSet) FAIL(AMR_ERR_STATE, "amr_estimate_lohner: mesh not set");
    if (!ctx->haveGeom) FAIL(AMR_ERR_STATE, "amr_estimate_lohner: geometry not set (amr_mesh_set_geometry)");
    if (!x) FAIL(AMR_ERR_ARG, "amr_estimate_lohner: x is null");
    if (scale && !thr) FAIL(AMR_ERR_ARG, "amr_estimate_lohner: scale requested without thresholds");
    const int64_t n = ctx->n, f = ctx->f, b = ctx->b;
    const double *dx = nullptr, *dxb = nullptr, *dxn = nullptr;
    TRY(stageIn(ctx, x, n, &dx));
    TRY(stageIn(ctx, xBoundary, b, &dxb));
    const bool devOut = errorOut && isDevicePtr(errorOut);
    double *target = devOut ? errorOut : ctx->error;
    ctx->errorResident = !devOut;
    if (ctx->nCoupled > 0) {
        if (nbrValues) TRY(stageIn(ctx, nbrValues, b, &dxn));
        else {
            TRY(haloSwapCells(ctx, dx, 8, ctx->ghost64));
            dxn = ctx->ghost64;
        }
    } else dxn = ctx->ghost64;
    Geom G{ctx->gWeights, ctx->gSf, ctx->gMagSf, ctx->gDelta, ctx->gV};
    Thresholds t = mkThr(thr);
    if (b) LAUNCH(k_boundary_field, gridFor(b), kThreads, b, (const label *)ctx->bOwner, dx, dxb, dxn, ctx->bCoupled, ctx->xbf);
    if (n) {
        if (ctx->faceBatch) {
            LAUNCH(k_gauss_grad_batched, gridFor(n), kThreads, n, f, ctx->cf.sliceOff, ctx->cf.col, ctx->owner, ctx->neighbour, G, dx,
                   (const double *)ctx->xbf, dxn, (const u8 *)ctx->bCoupled, ctx->grad);
            if (scale) LAUNCH((k_lohner_batched<true>), gridFor(n), kThreads, n, f, ctx->cf.sliceOff, ctx->cf.col, ctx->owner, ctx->neighbour, G, dx,
                              (const double *)ctx->grad, (const double *)ctx->xbf, dxn, (const u8 *)ctx->bCoupled, epsilon, t, target);
            else LAUNCH((k_lohner_batched<false>), gridFor(n), kThreads, n, f, ctx->cf.sliceOff, ctx->cf.col, ctx->owner, ctx->neighbour, G, dx,
                        (const double *)ctx->grad, (const double *)ctx->xbf, dxn, (const u8 *)ctx->bCoupled, epsilon, t, target);
        } else {
        LAUNCH(k_gauss_grad, gridFor(n), kThreads, n, f, ctx->cf.sliceOff, ctx->cf.col, ctx->owner, ctx->neighbour, G, dx,
               (const double *)ctx->xbf, dxn, (const u8 *)ctx->bCoupled, ctx->grad);
        if (scale) LAUNCH((k_lohner<true>), gridFor(n), kThreads, n, f, ctx->cf.sliceOff, ctx->cf.col, ctx->owner, ctx->neighbour, G, dx,
                          (const double *)ctx->grad, (const double *)ctx->xbf, dxn, (const u8 *)ctx->bCoupled, epsilon, t, target);
        else LAUNCH((k_lohner<false>), gridFor(n), kThreads, n, f, ctx->cf.sliceOff, ctx->cf.col, ctx->owner, ctx->neighbour, G, dx,
                    (const double *)ctx->grad, (const double *)ctx->xbf, dxn, (const u8 *)ctx->bCoupled, epsilon, t, target);
        }
        if (errorOut && !devOut) { CK(cudaMemcpyAsync(errorOut, target, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream)); ctx->hostTouched = true; }
    }
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_normalize(amr_ctx *ctx, double lowerRefine, double upperRefine, double lowerUnrefine, double upperUnrefine,
                          double *errorInOut) {
    ENTER();
    const int64_t n = ctx->n;
    double *dE = nullptr;
    TRY(stageInOut(ctx, errorInOut, n, &dE));
    double *e = dE ? dE : ctx->error;
    if (!e) FAIL(AMR_ERR_STATE, "amr_normalize: no error field");
    Thresholds t{lowerRefine, upperRefine, lowerUnrefine, upperUnrefine};
    if (n) LAUNCH((k_pointwise<false, true>), gridFor(n), kThreads, n, e, t, e);
    if (dE && n && dE != ctx->error) CK(cudaMemcpyAsync(ctx->error, dE, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

// gMax/gMin of the error field, the range thresholds and (optionally) normalize: gradientRange.C:102-109
static int normalizeRange(amr_ctx *ctx, double *e, double refineThreshold, double unrefineThreshold, int scale, double *thrOut,
                          double *minMaxOut) {
    const int64_t n = ctx->n;
    const int nPart = (int)std::min<int64_t>((int64_t)ctx->numSMs * 8, std::max<int64_t>(1, (n + kThreads - 1) / kThreads));
    void *q;
    TRY(arenaAlloc(ctx, &q, (size_t)(2 * nPart + 2) * 8));
    double *part = (double *)q, *res = part + 2 * nPart;
    double mm[2] = {kGREAT * 1e285, -kGREAT * 1e285};
    if (n > 0) {
        LAUNCH(k_minmax_partial, (unsigned)nPart, kThreads, n, (const double *)e, part);
        LAUNCH(k_minmax_final, 1, 1, nPart, (const double *)part, res);
    } else {
        CK(cudaMemcpyAsync(res, mm, 16, cudaMemcpyHostToDevice, ctx->stream));
    }
    TRY(allReduceWord(ctx, res, 2));         // gMin
    TRY(allReduceWord(ctx, res + 1, 3));     // gMax
    CK(cudaMemcpyAsync(mm, res, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const double minE = mm[0], maxE = mm[1];
    Thresholds t;
    t.lowerRefine = minE + refineThreshold * (maxE - minE);      // gradientRange.C:105
    t.upperRefine = kGREAT;
    t.lowerUnrefine = minE + unrefineThreshold * (maxE - minE);  // gradientRange.C:107
    t.upperUnrefine = kGREAT;
    if (thrOut) { thrOut[0] = t.lowerRefine; thrOut[1] = t.upperRefine; thrOut[2] = t.lowerUnrefine; thrOut[3] = t.upperUnrefine; }
    if (minMaxOut) { minMaxOut[0] = minE; minMaxOut[1] = maxE; }
    if (scale && n) LAUNCH((k_pointwise<false, true>), gridFor(n), kThreads, n, e, t, e);
    return AMR_OK;
}

AMR_API int amr_normalize_range(amr_ctx *ctx, double refineThreshold, double unrefineThreshold, int scale,
                                double *errorInOut, double *thrOut, double *minMaxOut) {
    ENTER();
    const int64_t n = ctx->n;
    double *dE = nullptr;
    TRY(stageInOut(ctx, errorInOut, n, &dE));
    double *e = dE ? dE : ctx->error;
    if (!e) FAIL(AMR_ERR_STATE, "amr_normalize_range: no error field");
    TRY(normalizeRange(ctx, e, refineThreshold, unrefineThreshold, scale, thrOut, minMaxOut));
    if (dE && n && dE != ctx->error) CK(cudaMemcpyAsync(ctx->error, dE, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

// One face layer of the marking `cur`, IN PLACE (kernels_sparse.cuh (1)): sources are the cells with generation
// 0 < v <= gen, or the non-zero cells of a separate source array (the first forced layer of the hex refiner,
// fvMeshRefiner.C:805-823); new marks get gen + 1.  `site` identifies the call site (which layer of which stage): the
// number of sources seen there by the PREVIOUS call picks the sparse push or the dense pull form -- AMR markings move
// slowly from step to step, a wrong prediction only costs time, both forms give the same flags -- and every layer
// kernel reports the exact count of this call as a by-product (read back with the call's final round trip).
constexpr int kSiteProtect = 0, kSiteRefine = 4, kSiteUnrefine = 8, kSiteSweep = 12;   // 4 slots each in ctx->hMarked[16]
static inline int siteOf(int base, int layer) { return base + (layer < 3 ? layer : 3); }

static int sitesBegin(amr_ctx *ctx) {
    ctx->sitesTouched = 0;
    // the by-product counts of the layers AND the queue sizes of the closure loop that follows: one memset per call
    CK(cudaMemsetAsync(ctx->dQ, 0, 2 * (kMaxIter + 2) * sizeof(int) + 16 * sizeof(unsigned long long), ctx->stream));
    return AMR_OK;
}
// enqueue the read-back of the counts; sitesCommit after the stream was synchronised
static int sitesFetch(amr_ctx *ctx) {
    if (ctx->sitesTouched) CK(cudaMemcpyAsync(ctx->hMarkedStage, ctx->dMarked, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    return AMR_OK;
}
static void sitesCommit(amr_ctx *ctx) {
    for (int s = 0; s < 16; s++)
        if (ctx->sitesTouched & (1u << s)) ctx->hMarked[s] = ctx->hMarkedStage[s];
    ctx->sitesTouched = 0;
}
static inline bool sparseForm(const amr_ctx *ctx, int site, int64_t n) {
    if (ctx->pushNum <= 0 || n < ctx->pushMinCells || n <= 0) return false;
    if (ctx->pushNum >= ctx->pushDen) return true;                 // AMR_B200_PUSH_FRACTION >= 1: always (tests force the form)
    const long long seen = ctx->hMarked[site];                     // sources at this site last time (-1: never seen -> dense)
    return seen >= 0 && seen * ctx->pushDen < n * ctx->pushNum;
}

// more than pullNum/pullDen of the cells were SOURCES at this site last time: at least that many are marked, the rows of
// the unmarked rest are the cheaper walk (k_layer_pull_sparse).  AMR_B200_PULL_FRACTION, "0" = never.
static inline bool sparsePullForm(const amr_ctx *ctx, int site, int64_t n) {
    if (ctx->pullNum <= 0 || n < ctx->pushMinCells || n <= 0) return false;
    const long long seen = ctx->hMarked[site];
    if (ctx->pullNum >= ctx->pullDen) return true;                 // >= 1: always (tests force the form)
    return seen >= 0 && seen * ctx->pullDen > n * ctx->pullNum;
}

static int extendLayer(amr_ctx *ctx, u8 *cur, const u8 *srcSeparate, int gen, int site) {
    const int64_t n = ctx->n;
    if (gen >= 254) FAIL(AMR_ERR_ARG, "more than 253 buffer layers in one call sequence");
    const u8 *srcArr = srcSeparate ? srcSeparate : cur;
    const bool separate = srcSeparate != nullptr;
    HaloToken halo;
    if (ctx->nCoupled > 0) TRY(haloBegin(ctx, srcArr, 1, ctx->ghost8, &halo));     // the values as they are BEFORE this layer
    unsigned long long *cnt = ctx->dMarked + site;
    if (n > 0) {
        const bool al = ((((uintptr_t)srcArr) | ((uintptr_t)cur)) & 15) == 0;
        if (sparseForm(ctx, site, n)) {
            LAUNCH(k_layer_push, gridFor(n, kThreads * 16), kThreads, n, ctx->cc.sliceOff, ctx->cc.col, srcArr, separate, gen, al, cur, cnt);
        } else if (sparsePullForm(ctx, site, n)) {
            // source mask as bits (n/8 bytes: L2-resident for the gathers), then the rows of the unmarked cells
            LAUNCH(k_pack_source_bits, gridFor((n + 31) / 32), kThreads, n, srcArr, separate, gen, al, ctx->srcBits, cnt);
            LAUNCH(k_layer_pull_sparse, gridFor(n, kThreads * 16), kThreads, n, ctx->cc.sliceOff, ctx->cc.col, (const unsigned *)ctx->srcBits, gen, al, cur);
        } else if (pipeOk(ctx, ctx->cc)) TRY(launchPipe(ctx, kp_layer_pull, ctx->cc, srcArr, separate, gen, cur, cnt));
        else LAUNCH(k_layer_pull, gridFor(n), kThreads, n, ctx->cc.sliceOff, ctx->cc.col, srcArr, separate, gen, cur, cnt);
        ctx->sitesTouched |= 1u << site;
    }
    if (ctx->nCoupled > 0) {
        GhostLink G = makeLink(ctx, &halo, ctx->ghost8, false);
        LAUNCH(k_fix_layer_link, gridFor(ctx->nCoupled), kThreads, G, ctx->nCoupled, (const label *)ctx->cplFace, (const label *)ctx->bOwner,
               separate, gen, cur);
    }
    return AMR_OK;
}

AMR_API int amr_protect_patches(amr_ctx *ctx, const int32_t *patchIds, int nPatchIds, int nLayers, double *errorInOut,
                                int64_t *nProtected) {
    ENTER();
    const int64_t n = ctx->n;
    if (nProtected) *nProtected = 0;
    if (nPatchIds == 0) return AMR_OK;   // errorEstimator.C:331-334
    if (!patchIds) FAIL(AMR_ERR_ARG, "amr_protect_patches: patchIds is null");
    double *dE = nullptr;
    TRY(stageInOut(ctx, errorInOut, n, &dE));
    double *e = dE ? dE : ctx->error;
    u8 *cur = ctx->flagA;
    TRY(sitesBegin(ctx));
    CK(cudaMemsetAsync(cur, 0, (size_t)n, ctx->stream));
    for (int k = 0; k < nPatchIds; k++) {
        int q = patchIds[k];
        if (q < 0 || q >= (int)ctx->patches.size()) FAIL(AMR_ERR_ARG, "amr_protect_patches: patch id out of range");
        const Patch &pp = ctx->patches[q];
        if (pp.size) LAUNCH(k_patch_cells_set, gridFor(pp.size), kThreads, ctx->bOwner, (int64_t)pp.start - ctx->f, (int64_t)pp.size, cur, (u8)1);
    }
    int gen = 1;
    for (int i = 0; i < nLayers; i++) {
        if (gen >= 250) {                 // renormalise the generations (errorEstimator.C:347 can ask for many layers)
            if (n) LAUNCH(k_flags_normalize, gridFor(n, kThreads * 16), kThreads, n, (const u8 *)cur, cur, (((uintptr_t)cur) & 15) == 0);
            gen = 1;
        }
        TRY(extendLayer(ctx, cur, nullptr, gen, siteOf(kSiteProtect, i)));
        gen++;
    }
    CK(cudaMemsetAsync(ctx->dCount, 0, sizeof(int64_t), ctx->stream));
    if (n) LAUNCH(k_zero_protected, gridFor(n), kThreads, n, cur, e, (unsigned long long *)ctx->dCount);
    if (dE && n && dE != ctx->error) CK(cudaMemcpyAsync(ctx->error, dE, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->hCount, ctx->dCount, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sitesFetch(ctx));
    ctx->hostTouched = true;              // the count is a host output: complete on return
    CK(cudaGetLastError());
    TRY(flushOuts(ctx));
    sitesCommit(ctx);
    if (nProtected) *nProtected = ctx->hCount[0];
    return AMR_OK;
}

// ============================================================================ stages 2+3

// Ascending list from flags.  emitListAsync enqueues the (single-pass) compaction; finishList reads the count back
// (one sync, shared with the caller's convergence flag when it passes flagSlot >= 0) and copies
// the entries to a host destination.
struct ListJob {
    int32_t *dList = nullptr;
    int32_t *dTotal = nullptr;
    int32_t *userOut = nullptr;
    bool hostOut = false;
};
static int listTarget(amr_ctx *ctx, int64_t N, int32_t *userOut, ListJob *job) {
    job->userOut = userOut;
    job->hostOut = userOut && !isDevicePtr(userOut);
    if (userOut) {
        if (job->hostOut) {
            if (!job->dList) { void *q; TRY(arenaAlloc(ctx, &q, (size_t)(N > 0 ? N : 1) * 4)); job->dList = (int32_t *)q; }
        } else job->dList = userOut;
    }
    return AMR_OK;
}
static int emitListAsync(amr_ctx *ctx, const u8 *flags, int64_t N, int32_t *userOut, ListJob *job) {
    TRY(listTarget(ctx, N, userOut, job));
    TRY(compactFlags(ctx, flags, N, job->dList, &job->dTotal));
    return AMR_OK;
}
// flagSlot >= 0: also fetch ctx->dFlags[flagSlot] (all-reduced over ranks) in the same round trip
static int finishList(amr_ctx *ctx, ListJob *job, int flagSlot, int *flagValue, int64_t *nOut, bool copyOut) {
    if (flagSlot >= 0) {
        TRY(allReduceWord(ctx, ctx->dFlags + flagSlot, 0));
        CK(cudaMemcpyAsync(ctx->hFlags + flagSlot, ctx->dFlags + flagSlot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaMemcpyAsync(ctx->hFlags + 8, job->dTotal, 4, cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sitesFetch(ctx));
    PEER_ERR_FETCH();
    CK(cudaStreamSynchronize(ctx->stream));
    PEER_ERR_CHECK();
    sitesCommit(ctx);
    if (flagSlot >= 0 && flagValue) *flagValue = ctx->hFlags[flagSlot];
    const int32_t total = ctx->hFlags[8];
    if (nOut) *nOut = total;
    if (copyOut && job->hostOut && total > 0) {
        CK(cudaMemcpyAsync(job->userOut, job->dList, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->hostTouched = true;
    }
    return AMR_OK;
}
static int finishListCopy(amr_ctx *ctx, ListJob *job) {
    const int32_t total = ctx->hFlags[8];
    if (job->hostOut && total > 0) {
        CK(cudaMemcpyAsync(job->userOut, job->dList, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->hostTouched = true;
    }
    return AMR_OK;
}
static int emitList(amr_ctx *ctx, const u8 *flags, int64_t N, int32_t *userOut, int64_t *nOut) {
    ListJob job;
    TRY(emitListAsync(ctx, flags, N, userOut, &job));
    return finishList(ctx, &job, -1, nullptr, nOut, true);
}

// |dLevel| <= 1 over the internal faces of the mesh, checked once per mesh (hexRef::checkRefinementLevels is what
// guarantees it in the reference: hexRef.C:2934-3025, called after every topology change).  With a communicator the
// answer is agreed over all ranks (min): it selects between loops with different collective structure, so every rank
// must take the same one.  Collective on the first call after a mesh / level change.
static int ensureBalanceChecked(amr_ctx *ctx) {
    if (ctx->balancedLocal >= 0 || !ctx->lvl) return AMR_OK;
    const int64_t n = ctx->n;
    CK(cudaMemsetAsync(ctx->dCount + 3, 0, 8, ctx->stream));
    if (n) LAUNCH(k_check_levels, gridFor(n), kThreads, n, ctx->cc.sliceOff, ctx->cc.col, (const label *)ctx->lvl, (unsigned long long *)(ctx->dCount + 3));
    int64_t bad = 0;
    TRY(globalSum(ctx, 3, &bad));              // all ranks (sum of the violating faces)
    ctx->balancedLocal = bad == 0 ? 1 : 0;
    return AMR_OK;
}
// The polyRefiner's closure on the queue machinery needs (a) the cell -> edge-neighbour graph and (b) a mesh that is 2:1
// balanced over every pair of cells sharing an edge (what refinement::edgeConsistentRefinement maintains).  Both are
// established once per mesh; the verdict is agreed over the ranks (it selects between loops with different collectives).
// AMR_B200_POLY_QUEUE=0 keeps the dense sweeps.
static int ensureEdgeNeighbours(amr_ctx *ctx) {
    if (ctx->balancedEdges >= 0) return AMR_OK;
    const char *e = getenv("AMR_B200_POLY_QUEUE");
    int ok = (e && e[0] == '0') ? 0 : 1;
    if (ok && ctx->ceState == 0) {
        TRY(sellBuildEdgeNeighbours(ctx, ctx->ce, ctx->ec, ctx->nEdges, ctx->n));
        ctx->ceState = ctx->ce.col ? 1 : -1;
    }
    // a rank without cells has an empty graph: fine; a rank whose graph could not be built vetoes the queue form
    if (ctx->ceState < 0 && ctx->n > 0 && ctx->nEdges > 0) ok = 0;
    const int64_t n = ctx->n;
    CK(cudaMemsetAsync(ctx->dCount + 3, 0, 8, ctx->stream));
    if (ok && n && ctx->ce.col)
        LAUNCH(k_check_levels, gridFor(n), kThreads, n, ctx->ce.sliceOff, ctx->ce.col, (const label *)ctx->lvl, (unsigned long long *)(ctx->dCount + 3));
    if (!ok) { const long long one = 1; CK(cudaMemcpyAsync(ctx->dCount + 3, &one, 8, cudaMemcpyHostToDevice, ctx->stream)); }
    int64_t bad = 0;
    TRY(globalSum(ctx, 3, &bad));
    ctx->balancedEdges = bad == 0 ? 1 : 0;
    return AMR_OK;
}
static int ensureQueues(amr_ctx *ctx, int64_t need) {
    if (need <= ctx->queueCap) return AMR_OK;
    devFree(ctx, &ctx->queue[0]); devFree(ctx, &ctx->queue[1]);
    TRY(devAlloc(ctx, &ctx->queue[0], need)); TRY(devAlloc(ctx, &ctx->queue[1], need));
    ctx->queueCap = need;
    return AMR_OK;
}
// grid of a queue kernel: enough CTAs to fill the device, the kernels stride over the queue
static inline unsigned queueGrid(const amr_ctx *ctx) { return (unsigned)(ctx->numSMs * 8); }

// One host look at a speculatively enqueued queue loop: sizes of queues [1, issued] (summed over ranks), the list size
// (optional) and the by-product counts in ONE round trip.  Returns the index of the first empty queue, or 0.
// the mailbox post rides on the compaction that ends the loop (MailArgs, scan.cuh)
static MailArgs queueMail(amr_ctx *ctx, const int *gq, int issued) {
    MailArgs m;
    m.mb = ctx->mailbox;
    m.gq = gq; m.nq = issued + 1;
    m.marked = ctx->dMarked;
    m.peerError = (ctx->usePeer && ctx->peerTouched) ? (const int *)(ctx->region + kPeerErrOffset) : nullptr;
    m.seq = ++ctx->mailSeq;
    return m;
}
static int queueLook(amr_ctx *ctx, int issued) {
    Mailbox *mb = ctx->mailbox;
    const unsigned seq = ctx->mailSeq;
    CK(cudaGetLastError());
    // poll the mailbox (a PCIe write away) instead of synchronising the stream; fall back to the stream on anything odd
    bool arrived = false;
    for (long spins = 0; ; spins++) {
        if (__atomic_load_n(&mb->seq, __ATOMIC_ACQUIRE) == seq) { arrived = true; break; }
        if ((spins & 0xfffff) == 0xfffff) {
            const cudaError_t e = cudaStreamQuery(ctx->stream);
            if (e == cudaSuccess) { arrived = __atomic_load_n(&mb->seq, __ATOMIC_ACQUIRE) == seq; break; }
            if (e != cudaErrorNotReady) { ctx->err = std::string("queue loop: ") + cudaGetErrorString(e); return AMR_ERR_CUDA; }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    if (!arrived) CK(cudaStreamSynchronize(ctx->stream));
    traceMark(ctx, "host look returned");
    ctx->peerTouched = false;
    if (mb->peerError) {
        cudaMemsetAsync(ctx->region + kPeerErrOffset, 0, sizeof(int), ctx->stream);
        FAIL(AMR_ERR_STATE, "peer-memory halo swap / reduction timed out: a neighbour rank did not enter the same collective call "
                            "(results of this call are invalid; AMR_B200_PEER_TIMEOUT_S sets the limit)");
    }
    for (int s2 = 0; s2 < 16; s2++)
        if (ctx->sitesTouched & (1u << s2)) ctx->hMarked[s2] = mb->marked[s2];
    ctx->sitesTouched = 0;
    ctx->hFlags[8] = mb->total;
    ctx->lookResult = 0;
    for (int j = 1; j <= issued; j++)
        if (mb->q[j] == 0) { ctx->lookResult = j; break; }
    return AMR_OK;
}

// hexRef::consistentRefinement(maxSet = true) as queue iterations (kernels_sparse.cuh (2)) + the ascending list
// (hexRef.C:1440-1460).  The iterations the previous call needed are enqueued in one go, each launch gated on the
// device by the (all-reduced) size of its queue, then the list is compacted and the host looks ONCE; only if the
// closure needed more iterations than last time does the host look again.
static int refineClosureQueue(amr_ctx *ctx, const Sell &A, u8 *set, u8 *lf, int *sweepsOut, int32_t *cellsOut, int64_t *nOut) {
    const int64_t n = ctx->n;
    TRY(ensureQueues(ctx, n + 16));
    const bool multi = ctx->comm && ctx->nranks > 1;
    const bool speculate = gatedSwapsOk(ctx);
    int *qc = ctx->dQ, *gq = multi ? ctx->dGQ : ctx->dQ;
    const int64_t cap = ctx->queueCap;
    const label *cf = ctx->cplFace, *bo = ctx->bOwner;
    ListJob job;                          // (queue sizes were cleared by sitesBegin at the start of the call)
    TRY(listTarget(ctx, n, cellsOut, &job));
    const bool linked = multi && speculate && linkedOk(ctx);
    const unsigned nFix = linked ? gridFor(std::max<int64_t>(ctx->nCoupled, 1)) : 0u;
    const LinkOut L = linked ? makeLinkOut(ctx) : LinkOut();
    // sweep 0: the rows of all marked cells
    if (linked) {
        // ONE swap for the whole closure, then one launch per iteration (kernels_linked.cuh)
        HaloToken halo;
        if (ctx->nCoupled > 0) TRY(haloBegin(ctx, lf, 1, ctx->ghost8, &halo));
        GhostLink G = makeLink(ctx, &halo, ctx->ghost8, true);
        LAUNCH((k_refine_linked<true>), nFix + gridFor(std::max<int64_t>(n, 1), kThreads * 16), kThreads, G, L, nFix, ctx->nCoupled, cf, bo, n, A.sliceOff,
               A.col, (const u8 *)ctx->lvl8, lf, set, ((((uintptr_t)set)) & 15) == 0, (const label *)nullptr, (const int *)nullptr, ctx->queue[1],
               qc + 1, cap, gq + 1, (const int *)nullptr);
        ctx->peerTouched = true;
    } else {
        HaloToken halo;
        if (ctx->nCoupled > 0) TRY(haloBegin(ctx, lf, 1, ctx->ghost8, &halo));     // lf as it is before the sweep, like the reference's neiLevel
        if (n) LAUNCH(k_refine_push0, gridFor(n, kThreads * 16), kThreads, n, A.sliceOff, A.col, (const u8 *)ctx->lvl8, lf, set,
                      ((((uintptr_t)set)) & 15) == 0, ctx->queue[1], qc + 1, cap);
        if (multi) {                  // coupled faces (hexRef.C:758-780) + reduce(nChanged): one launch on the peer transport
            GhostLink G = makeLink(ctx, &halo, ctx->ghost8, true);
            LAUNCH(k_fix_refine_queue_link, gridFor(std::max<int64_t>(ctx->nCoupled, 1)), kThreads, G, ctx->nCoupled, cf, bo, (const u8 *)ctx->lvl8, lf, set,
                   ctx->queue[1], qc + 1, cap, gq + 1, (const int *)nullptr);
            if (!G.peer) TRY(allReduceWord(ctx, gq + 1, 4, qc + 1));
        }
    }
    int issued = 1;                       // queues [1, issued] have been produced
    int batch = speculate ? std::max(1, ctx->itersSeen[0] - 1) : 0;
    int iters = 0;
    while (true) {
        for (int k = 0; k < batch && issued < kMaxIter; k++) {
            const int j = issued;
            const int *gate = gq + j;
            const int *xgate = speculate ? gate : nullptr;             // host-checked iterations are never dead
            if (linked) {
                GhostLink G = makeLink(ctx, nullptr, ctx->ghost8, true);
                G.ghost = nullptr;                                     // the ghost array is the parity buffer of the closure's swap
                LAUNCH((k_refine_linked<false>), nFix + queueGrid(ctx), kThreads, G, L, nFix, ctx->nCoupled, cf, bo, n, A.sliceOff, A.col,
                       (const u8 *)ctx->lvl8, lf, set, false, (const label *)ctx->queue[j & 1], (const int *)(qc + j), ctx->queue[(j + 1) & 1], qc + j + 1,
                       cap, gq + j + 1, gate);
                ctx->peerTouched = true;
                issued++;
                continue;
            }
            HaloToken halo;
            if (ctx->nCoupled > 0) TRY(haloBegin(ctx, lf, 1, ctx->ghost8, &halo, nullptr, xgate));
            LAUNCH(k_refine_frontier, queueGrid(ctx), kThreads, A.sliceOff, A.col, (const u8 *)ctx->lvl8, lf, set,
                   (const label *)ctx->queue[j & 1], (const int *)(qc + j), ctx->queue[(j + 1) & 1], qc + j + 1, cap, gate);
            if (multi) {
                GhostLink G = makeLink(ctx, &halo, ctx->ghost8, true);
                LAUNCH(k_fix_refine_queue_link, gridFor(std::max<int64_t>(ctx->nCoupled, 1)), kThreads, G, ctx->nCoupled, cf, bo, (const u8 *)ctx->lvl8, lf, set,
                       ctx->queue[(j + 1) & 1], qc + j + 1, cap, gq + j + 1, xgate);
                if (!G.peer) TRY(allReduceWord(ctx, gq + j + 1, 4, qc + j + 1, xgate));
            }
            issued++;
        }
        const MailArgs mail = queueMail(ctx, gq, issued);
        TRY(compactFlagsEx(ctx, set, n, nullptr, nullptr, nullptr, job.dList, &job.dTotal, 0, nullptr, &mail));   // speculative: valid if the loop has converged
        TRY(queueLook(ctx, issued));
        if (ctx->lookResult) { iters = ctx->lookResult; break; }
        if (issued >= kMaxIter) FAIL(AMR_ERR_INVARIANT, "2:1 closure did not settle (more queue iterations than refinement levels can need)");
        batch = speculate ? 2 : 1;
    }
    ctx->itersSeen[0] = iters;
    if (nOut) *nOut = ctx->hFlags[8];
    TRY(finishListCopy(ctx, &job));
    if (sweepsOut) *sweepsOut = iters;
    return AMR_OK;
}

// hexRef::consistentRefinement loop (hexRef.C:1421-1437) with dense sweeps: maxSet = false, the polyRefiner's edge rule
// (refinement::consistentRefinement, refinement.C:776-780: the edge rule first, then the face rule), and meshes that
// are not 2:1 balanced.  One host round trip per sweep; the ascending list (hexRef.C:1440-1460) after the last one.
static int refineSweepsDense(amr_ctx *ctx, int maxSet, u8 *set, u8 *lf, int *sweepsOut, int32_t *cellsOut, int64_t *nOut, bool edgeRule) {
    const int64_t n = ctx->n;
    int sweeps = 0;
    while (true) {
        CK(cudaMemsetAsync(ctx->dFlags, 0, sizeof(int), ctx->stream));
        if (edgeRule && ctx->nEdges) {
            if (maxSet) LAUNCH((k_edge_refine_sweep<true>), gridFor(ctx->nEdges), kThreads, ctx->nEdges, ctx->ec.sliceOff, ctx->ec.col, lf, set, ctx->dFlags);
            else LAUNCH((k_edge_refine_sweep<false>), gridFor(ctx->nEdges), kThreads, ctx->nEdges, ctx->ec.sliceOff, ctx->ec.col, lf, set, ctx->dFlags);
        }
        HaloToken halo;
        if (ctx->nCoupled > 0) TRY(haloBegin(ctx, lf, 1, ctx->ghost8, &halo));
        if (n) {
            if (pipeOk(ctx, ctx->cc)) {
                if (maxSet) TRY(launchPipe(ctx, kp_refine_sweep<true>, ctx->cc, lf, set, ctx->dFlags, (const int *)nullptr));
                else TRY(launchPipe(ctx, kp_refine_sweep<false>, ctx->cc, lf, set, ctx->dFlags, (const int *)nullptr));
            } else if (maxSet) LAUNCH(k_refine_sweep_max, gridFor(n), kThreads, n, ctx->cc.sliceOff, ctx->cc.col, lf, set, ctx->dFlags, (const int *)nullptr);
            else LAUNCH(k_refine_sweep_min, gridFor(n), kThreads, n, ctx->cc.sliceOff, ctx->cc.col, lf, set, ctx->dFlags);
        }
        TRY(haloEnd(ctx, &halo));
        if (ctx->nCoupled > 0 && n) {      // hexRef.C:758-780
            if (maxSet) LAUNCH((k_fix_refine_sweep<true>), gridFor(ctx->nCoupled), kThreads, ctx->nCoupled, (const label *)ctx->cplFace,
                               (const label *)ctx->bOwner, (const u8 *)ctx->lvl8, (const u8 *)ctx->ghost8, lf, set, ctx->dFlags);
            else LAUNCH((k_fix_refine_sweep<false>), gridFor(ctx->nCoupled), kThreads, ctx->nCoupled, (const label *)ctx->cplFace,
                        (const label *)ctx->bOwner, (const u8 *)ctx->lvl8, (const u8 *)ctx->ghost8, lf, set, ctx->dFlags);
        }
        sweeps++;
        int changed = 0;
        TRY(globalFlag(ctx, 0, &changed));
        if (!changed) break;
        if (sweeps > 100000) FAIL(AMR_ERR_INVARIANT, "2:1 sweeps did not converge");
    }
    ListJob job;
    TRY(emitListAsync(ctx, set, n, cellsOut, &job));
    TRY(finishList(ctx, &job, -1, nullptr, nOut, true));
    if (sweepsOut) *sweepsOut = sweeps;
    return AMR_OK;
}

static int refineSweeps(amr_ctx *ctx, int maxSet, u8 *set, u8 *lf, int *sweepsOut, int32_t *cellsOut, int64_t *nOut,
                        bool edgeRule = false) {
    // the queue form is only valid on a locally 2:1-balanced mesh: verified once per mesh
    bool queueOk = maxSet && ctx->pushNum > 0 && ctx->lvl;      // the same on every rank
    if (queueOk && edgeRule) {
        // polyRefiner (refinement.C:744-810: edge rule + face rule per sweep): one rule on the graph of cells that share an
        // edge, which contains the internal face pairs (sell.cuh); coupled faces as in the hexRefiner's loop
        TRY(ensureEdgeNeighbours(ctx));
        queueOk = ctx->balancedEdges == 1;
        if (queueOk) return refineClosureQueue(ctx, ctx->ce, set, lf, sweepsOut, cellsOut, nOut);
    } else if (queueOk) {
        TRY(ensureBalanceChecked(ctx));
        queueOk = ctx->balancedLocal == 1;
    }
    if (queueOk) return refineClosureQueue(ctx, ctx->cc, set, lf, sweepsOut, cellsOut, nOut);
    return refineSweepsDense(ctx, maxSet, set, lf, sweepsOut, cellsOut, nOut, edgeRule);
}

// fvMeshHexRefiner::calculateProtectedCells (fvMeshHexRefiner.C:60-179): closure of protectedCell_
// over faces towards finer neighbours.  Uses the extend kernel on a source mask that is rebuilt
// each round: src[q] = unrefineable[q] restricted to "neighbour is finer", which needs the levels
// of both sides -> dedicated kernel.
__global__ void __launch_bounds__(kThreads)
k_protected_spread(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
                   const u8 *__restrict__ lvl8, u8 *__restrict__ u, int *changed) {
    // cell c becomes unrefineable when a face neighbour q is unrefineable and level[c] > level[q]
    // (seedFace rule, fvMeshHexRefiner.C:101-128, then both cells of a seed face are set :137-170;
    // the coarser one already is)
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    int flip = 0;
    if (c < n && !u[c]) {
        const int32_t base = sliceOff[c >> 5];
        const int w = sliceOff[(c >> 5) + 1] - base;
        const int32_t *row = col + ((size_t)base << 5) + (c & 31);
        const int mine = lvl8[c];
        for (int j = 0; j < w; j++) {
            int32_t q = row[(size_t)j << 5];
            if (q == (int32_t)c) continue;
            if (u[q] && mine > (int)lvl8[q]) { flip = 1; break; }   // coupled faces: k_fix_protected
        }
        if (flip) u[c] = 1;
    }
    if (__syncthreads_or(flip) && threadIdx.x == 0) *changed = 1;
}

static int calcProtected(amr_ctx *ctx, u8 *unrefineable, bool *have) {
    const int64_t n = ctx->n, b = ctx->b;
    // returnReduce(protectedCell_.size(), sumOp) == 0 -> empty list (fvMeshHexRefiner.C:65-69)
    if (ctx->anyProtCached < 0) {      // reduced once per mesh / protectedCell change (both are set collectively)
        CK(cudaMemsetAsync(ctx->dFlags + 2, 0, sizeof(int), ctx->stream));
        if (ctx->protCell) { int one = 1; CK(cudaMemcpyAsync(ctx->dFlags + 2, &one, sizeof(int), cudaMemcpyHostToDevice, ctx->stream)); }
        int anyR = 0;
        TRY(globalFlag(ctx, 2, &anyR));
        ctx->anyProtCached = anyR ? 1 : 0;
    }
    const int any = ctx->anyProtCached;
    *have = any != 0;
    if (!any) return AMR_OK;
    if (ctx->protCell) CK(cudaMemcpyAsync(unrefineable, ctx->protCell, (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    else CK(cudaMemsetAsync(unrefineable, 0, (size_t)n, ctx->stream));
    u8 *ghostLvl = ctx->flagC;   // [b] fits: flagC has max(n,p) >= ... guard below
    if (ctx->nCoupled > 0) {
        if (b > std::max(ctx->n, ctx->p)) FAIL(AMR_ERR_STATE, "scratch too small for boundary levels");
        TRY(haloSwapCells(ctx, ctx->lvl8, 1, ghostLvl));
    }
    while (true) {
        CK(cudaMemsetAsync(ctx->dFlags + 1, 0, sizeof(int), ctx->stream));
        if (ctx->nCoupled > 0) {
            TRY(haloSwapCells(ctx, unrefineable, 1, ctx->ghost8));
        }
        if (n) LAUNCH(k_protected_spread, gridFor(n), kThreads, n, ctx->cc.sliceOff, ctx->cc.col, ctx->lvl8, unrefineable, ctx->dFlags + 1);
        if (ctx->nCoupled > 0)
            LAUNCH(k_fix_protected, gridFor(ctx->nCoupled), kThreads, ctx->nCoupled, (const label *)ctx->cplFace, (const label *)ctx->bOwner,
                   (const u8 *)ctx->lvl8, (const u8 *)ctx->ghost8, (const u8 *)ghostLvl, unrefineable, ctx->dFlags + 1);
        int changed = 0;
        TRY(globalFlag(ctx, 1, &changed));
        if (!changed) break;
    }
    return AMR_OK;
}

AMR_API int amr_select_refine(amr_ctx *ctx, const double *error, double lo, double hi, const int32_t *maxCellLevel,
                              int32_t maxLevelUniform, int nBufferLayers, const int32_t *protectedPatchIds,
                              int nProtectedPatches, int64_t maxCells, int refinerKind, uint8_t *refineCellOut,
                              int32_t *cellsOut, int64_t *nOut, int *sweepsOut) {
    ENTER();
    if (!ctx->meshSet) FAIL(AMR_ERR_STATE, "amr_select_refine: mesh not set");
    if (!ctx->haveLevels) FAIL(AMR_ERR_STATE, "amr_select_refine: levels not set");
    if (refinerKind != AMR_REFINER_HEX3D && refinerKind != AMR_REFINER_HEX2D && refinerKind != AMR_REFINER_POLY3D &&
        refinerKind != AMR_REFINER_POLY2D)
        FAIL(AMR_ERR_ARG, "amr_select_refine: unknown refiner kind");
    const bool poly = refinerKind == AMR_REFINER_POLY3D || refinerKind == AMR_REFINER_POLY2D;
    if (poly && ctx->optEdgeConsistency && !ctx->haveEdges)
        FAIL(AMR_ERR_STATE, "amr_select_refine: polyRefiner with edgeBasedConsistency needs amr_mesh_set_edges");
    if (nProtectedPatches > 0 && !protectedPatchIds) FAIL(AMR_ERR_ARG, "amr_select_refine: protectedPatchIds is null");
    if (nBufferLayers > 250) FAIL(AMR_ERR_ARG, "amr_select_refine: more than 250 buffer layers");
    const int64_t n = ctx->n;
    const double *dErr = nullptr;
    const int32_t *dMax = nullptr;
    TRY(stageIn(ctx, error, n, &dErr));
    if (!dErr) {
        if (!ctx->errorResident) FAIL(AMR_ERR_STATE, "amr_select_refine: error == NULL but the last estimate wrote to a caller-owned device buffer");
        dErr = ctx->error;
    }
    TRY(stageIn(ctx, maxCellLevel, n, &dMax));
    u8 *dRcOut = nullptr;
    TRY(stageOut(ctx, refineCellOut, n, &dRcOut));
    TRY(sitesBegin(ctx));

    // the marking (the reference's refineCell) is built in place in the resident buffer and stays there for the
    // unrefinement stage
    u8 *cur = ctx->refineCell, *src0 = ctx->flagC;
    // hexRefiner always passes force=true (fvMeshHexRefiner.C:1511), polyRefiner passes force_ (fvMeshPolyRefiner.C:338)
    const bool forced = poly ? (ctx->optPolyForce != 0) : true;
    // M1 (+ source mask of the first forced layer)
    {
        const bool al = ((((uintptr_t)dErr) | ((uintptr_t)dMax)) & 15) == 0;
        if (n) LAUNCH(k_candidates, gridFor(n, kThreads * 4), kThreads, n, dErr, lo, hi, ctx->lvl8, dMax, maxLevelUniform, cur,
                      (nBufferLayers > 0 && forced) ? src0 : (u8 *)nullptr, al);
    }
    // M2: extendMarkedCells(refineCell, maxRefinement, i == 0, force = true), fvMeshHexRefiner.C:1509-1512
    int gen = 1;
    for (int i = 0; i < nBufferLayers; i++) {
        TRY(extendLayer(ctx, cur, (i == 0 && forced) ? src0 : (const u8 *)nullptr, gen, siteOf(kSiteRefine, i)));
        gen++;
    }
    // M3: protections, fvMeshHexRefiner.C:1514-1533 (polyRefiner: patches only, fvMeshPolyRefiner.C:341-350)
    for (int k = 0; k < nProtectedPatches; k++) {
        int q = protectedPatchIds[k];
        if (q < 0 || q >= (int)ctx->patches.size()) FAIL(AMR_ERR_ARG, "amr_select_refine: patch id out of range");
        const Patch &pp = ctx->patches[q];
        if (pp.size) LAUNCH(k_patch_cells_set, gridFor(pp.size), kThreads, ctx->bOwner, (int64_t)pp.start - ctx->f, (int64_t)pp.size, cur, (u8)0);
    }
    if (!poly && ctx->protCell && n) LAUNCH(k_masked_set, gridFor(n), kThreads, n, ctx->protCell, cur, (u8)0);
    ctx->refineCellN = n;
    ctx->refineCellGen = gen;
    if (dRcOut && n) LAUNCH(k_flags_normalize, gridFor(n, kThreads * 16), kThreads, n, (const u8 *)cur, dRcOut, ((((uintptr_t)cur) | ((uintptr_t)dRcOut)) & 15) == 0);

    // M4/M5: selectRefineCells, fvMeshHexRefiner.C:306-391
    u8 *set = ctx->flagA;
    u8 *unref = ctx->flagB;
    bool haveU = false;
    if (poly) {
        // base selectRefineCells (fvMeshRefiner.C:190-219): level < maxRefinement && candidate; no quota, no protected cells
        if (n && !dMax) LAUNCH(k_select_filter16, gridFor(n, kThreads * 16), kThreads, n, (const u8 *)cur, (const u8 *)ctx->lvl8, (int)maxLevelUniform, (const u8 *)nullptr, set, ctx->lf);
        else if (n) LAUNCH(k_select_filter, gridFor(n, kThreads * 4), kThreads, n, cur, ctx->lvl, dMax, maxLevelUniform, (u8 *)nullptr, -1, set, ctx->lf, ((((uintptr_t)dMax)) & 15) == 0);
        TRY(refineSweeps(ctx, 1, set, ctx->lf, sweepsOut, cellsOut, nOut, ctx->optEdgeConsistency != 0));
        CK(cudaGetLastError());
        return flushOuts(ctx);
    }
    TRY(calcProtected(ctx, unref, &haveU));
    // nTotalCells / nCandidates reductions only matter when the maxCells quota can bind
    if (ctx->nTotalCells < 0) {
        int64_t nTotal = n;
        if (ctx->comm && ctx->nranks > 1) {
            CK(cudaMemcpyAsync(ctx->dCount + 1, &nTotal, 8, cudaMemcpyHostToDevice, ctx->stream));
            TRY(globalSum(ctx, 1, &nTotal));
        }
        ctx->nTotalCells = nTotal;
    }
    const int64_t nTotal = ctx->nTotalCells;
    int64_t nTotToRefine = (maxCells - nTotal) / 7;
    bool quota = false;
    if (nTotal >= nTotToRefine) {     // otherwise nCandidates <= nTotal < nTotToRefine always holds
        CK(cudaMemsetAsync(ctx->dCount, 0, 8, ctx->stream));
        if (n) LAUNCH(k_count_flags, gridFor(n), kThreads, n, cur, (unsigned long long *)ctx->dCount);
        int64_t nCand = 0;
        TRY(globalSum(ctx, 0, &nCand));
        quota = !(nCand < nTotToRefine);
    }
    if (!quota) {
        if (n && !dMax) LAUNCH(k_select_filter16, gridFor(n, kThreads * 16), kThreads, n, (const u8 *)cur, (const u8 *)ctx->lvl8, (int)maxLevelUniform, haveU ? (const u8 *)unref : (const u8 *)nullptr, set, ctx->lf);
        else if (n) LAUNCH(k_select_filter, gridFor(n, kThreads * 4), kThreads, n, cur, ctx->lvl, dMax, maxLevelUniform, haveU ? unref : (u8 *)nullptr, -1, set, ctx->lf, ((((uintptr_t)dMax)) & 15) == 0);
    } else {
        // level-by-level truncation, fvMeshHexRefiner.C:346-374; gMax(maxRefinement)
        int32_t maxLevel = maxLevelUniform;
        if (dMax) {
            std::vector<int32_t> h((size_t)n);
            CK(cudaMemcpyAsync(h.data(), dMax, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            maxLevel = n ? *std::max_element(h.begin(), h.end()) : 0;
        }
        if (ctx->comm && ctx->nranks > 1) {
            CK(cudaMemcpyAsync(ctx->dFlags + 3, &maxLevel, 4, cudaMemcpyHostToDevice, ctx->stream));
            int g = 0; TRY(globalFlag(ctx, 3, &g)); maxLevel = g;
        }
        CK(cudaMemsetAsync(set, 0, (size_t)n, ctx->stream));
        for (int32_t level = 0; level < maxLevel; level++) {
            if (n) LAUNCH(k_select_filter, gridFor(n, kThreads * 4), kThreads, n, cur, ctx->lvl, dMax, maxLevelUniform, haveU ? unref : (u8 *)nullptr, level, set, ctx->lf, ((((uintptr_t)dMax)) & 15) == 0);
            CK(cudaMemsetAsync(ctx->dCount, 0, 8, ctx->stream));
            if (n) LAUNCH(k_count_flags, gridFor(n), kThreads, n, set, (unsigned long long *)ctx->dCount);
            int64_t tot = 0;
            TRY(globalSum(ctx, 0, &tot));
            if (tot > nTotToRefine) break;
        }
        if (maxLevel <= 0 && n) LAUNCH(k_select_filter, gridFor(n, kThreads * 4), kThreads, n, cur, ctx->lvl, dMax, maxLevelUniform, haveU ? unref : (u8 *)nullptr, 1 << 20, set, ctx->lf, ((((uintptr_t)dMax)) & 15) == 0);
    }
    // B1: hexRef::consistentRefinement(candidates, true) + ascending list
    TRY(refineSweeps(ctx, 1, set, ctx->lf, sweepsOut, cellsOut, nOut));
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_consistent_refinement(amr_ctx *ctx, const int32_t *cellsIn, int64_t nIn, int maxSet, int32_t *cellsOut,
                                      int64_t *nOut, int *sweepsOut) {
    ENTER();
    if (!ctx->haveLevels) FAIL(AMR_ERR_STATE, "amr_consistent_refinement: levels not set");
    const int64_t n = ctx->n;
    const int32_t *dIn = nullptr;
    TRY(stageIn(ctx, cellsIn, nIn, &dIn));
    TRY(sitesBegin(ctx));
    u8 *set = ctx->flagA;
    CK(cudaMemsetAsync(set, 0, (size_t)n, ctx->stream));
    if (nIn > 0) LAUNCH(k_list_to_flags, gridFor(nIn), kThreads, nIn, dIn, set);
    if (n) LAUNCH(k_select_filter, gridFor(n, kThreads * 4), kThreads, n, set, ctx->lvl, (const label *)nullptr, (label)0x7fffffff, (const u8 *)nullptr, -1, set, ctx->lf, true);
    TRY(refineSweeps(ctx, maxSet, set, ctx->lf, sweepsOut, cellsOut, nOut));
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_remap_refine_cell(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *cellMap,
                                  const int32_t *reverseCellMap, const uint8_t *refineCell, uint8_t *newRefineCellOut) {
    ENTER();
    if (!cellMap || !reverseCellMap || nOld < 0 || nNew < 0) FAIL(AMR_ERR_ARG, "amr_remap_refine_cell: bad arguments");
    if (!refineCell && nOld != ctx->refineCellN) FAIL(AMR_ERR_STATE, "amr_remap_refine_cell: resident marking has a different size than nOld");
    // Called after amr_mesh_set of the NEW mesh or before it: the old marking is either passed in
    // or the resident one (which amr_mesh_set preserves in `pendingFlags`) -- to stay simple the
    // call must come BEFORE amr_mesh_set of the new mesh when the resident marking is used.
    const int32_t *dMap = nullptr, *dRev = nullptr;
    const u8 *dOld = nullptr;
    TRY(stageIn(ctx, cellMap, nNew, &dMap));
    TRY(stageIn(ctx, reverseCellMap, nOld, &dRev));
    TRY(stageIn(ctx, refineCell, nOld, &dOld));
    if (!dOld) dOld = ctx->refineCell;
    if (!dOld) FAIL(AMR_ERR_STATE, "amr_remap_refine_cell: no marking");
    if (nNew + 16 <= ctx->flagCap) {
        // the context already holds the new mesh (or one at least as large): remap into a pool buffer
        // and make it the resident marking by pointer swap
        u8 *dNew = ctx->flagA;
        if (nNew) LAUNCH(k_remap_refine_cell, gridFor(nNew, kThreads * 4), kThreads, nNew, dMap, dRev, dOld, dNew);
        if (newRefineCellOut && nNew) {
            CK(cudaMemcpyAsync(newRefineCellOut, dNew, (size_t)nNew, cudaMemcpyDefault, ctx->stream));
            if (!isDevicePtr(newRefineCellOut)) ctx->hostTouched = true;
        }
        if (ctx->hostTouched) CK(cudaStreamSynchronize(ctx->stream));
        std::swap(ctx->flagA, ctx->refineCell);
        ctx->refineCellN = nNew;
    } else {
        u8 *dNew = nullptr;
        const int64_t capNew = std::max<int64_t>(nNew + 16, ctx->flagCap);
        CK(cudaMalloc((void **)&dNew, (size_t)capNew));
        if (nNew) LAUNCH(k_remap_refine_cell, gridFor(nNew, kThreads * 4), kThreads, nNew, dMap, dRev, dOld, dNew);
        if (newRefineCellOut && nNew) CK(cudaMemcpyAsync(newRefineCellOut, dNew, (size_t)nNew, cudaMemcpyDefault, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        // the larger buffer leaves the pool convention; amr_mesh_set of the new mesh adopts its content
        if (ctx->refineCell) cudaFree(ctx->refineCell);
        ctx->refineCell = dNew; ctx->refineCellN = nNew;
        ctx->refineCellOversize = true;
    }
    ctx->refineCellGen = 1;              // the remapped marking is a plain boolList again
    CK(cudaGetLastError());
    return AMR_OK;
}

// split-point pre-filter (mesh-constant) and the fused U1+U2+U3 kernel.  When only a minority of the points can
// be split points the direct kernel (threads of non-candidates exit, their rows are never read) beats the
// TMA pipeline, which streams every tile.
static int selectUnrefinePoints(amr_ctx *ctx, const double *dErr, double unrefineLevel, const u8 *dRc, u8 *isSplit, u8 *keep) {
    const int64_t p = ctx->p;
    if (!p) return AMR_OK;
    if (ctx->pcClassDirty) {
        if (!ctx->pcClass) TRY(devAlloc(ctx, &ctx->pcClass, p + 16));
        CK(cudaMemsetAsync(ctx->dCount + 2, 0, 8, ctx->stream));
        LAUNCH(k_point_class, gridFor(p), kThreads, p, (const u8 *)ctx->pc.count, (const u8 *)ctx->bndPoint,
               (const label *)ctx->pointLevel, ctx->pcClass, (unsigned long long *)(ctx->dCount + 2));
        CK(cudaMemcpyAsync(ctx->hCount + 2, ctx->dCount + 2, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->nSplitCand = ctx->hCount[2];
        devFree(ctx, &ctx->candList);
        if (ctx->nSplitCand * 2 <= p) {       // sparse: keep an ascending candidate list (mesh-constant, like pcClass)
            TRY(devAlloc(ctx, &ctx->candList, ctx->nSplitCand));
            int32_t *dTotal = nullptr;
            TRY(compactFlags(ctx, ctx->pcClass, p, ctx->candList, &dTotal));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        ctx->pcClassDirty = false;
    }
    if (pipeOk(ctx, ctx->pc) && ctx->nSplitCand * 2 > p)
        TRY(launchPipe(ctx, kp_select_unrefine_points, ctx->pc, (const u8 *)ctx->pcClass, (const label *)ctx->parentIdx,
                       (const u8 *)ctx->lvl8, dErr, unrefineLevel, dRc, isSplit, keep));
    else if (ctx->candList) {
        if (isSplit) CK(cudaMemsetAsync(isSplit, 0, (size_t)p, ctx->stream));
        if (keep) CK(cudaMemsetAsync(keep, 0, (size_t)p, ctx->stream));
        if (ctx->nSplitCand)
            LAUNCH(k_select_unrefine_cand, gridFor(ctx->nSplitCand), kThreads, ctx->nSplitCand, (const label *)ctx->candList, ctx->pc.sliceOff,
                   ctx->pc.col, (const label *)ctx->parentIdx, (const u8 *)ctx->lvl8, dErr, unrefineLevel, dRc, isSplit, keep);
    } else
        LAUNCH(k_select_unrefine_points, gridFor(p), kThreads, p, ctx->pc.sliceOff, ctx->pc.col, (const u8 *)ctx->pcClass,
               (const label *)ctx->parentIdx, (const u8 *)ctx->lvl8, dErr, unrefineLevel, dRc, isSplit, keep);
    return AMR_OK;
}

// hexRef3D::getSplitElems (hexRef3D.C:2183-2321) depends on the mesh, the levels and the refinement history only: the
// ascending split-point list, the 8 cells of every split point and the split point of every cell are built ONCE per
// mesh (amr_mesh_set_points / amr_levels_set invalidate them).  The per-step work of the unrefinement branch then
// never touches the pointCells addressing again.
static int ensureSplitTable(amr_ctx *ctx) {
    if (!ctx->spDirty) return AMR_OK;
    const int64_t n = ctx->n, p = ctx->p;
    devFree(ctx, &ctx->spPoint); devFree(ctx, &ctx->spCells); devFree(ctx, &ctx->cellSp); devFree(ctx, &ctx->spFlag); devFree(ctx, &ctx->spList);
    ctx->nSp = 0;
    u8 *isSplit = ctx->flagC;
    TRY(selectUnrefinePoints(ctx, (const double *)nullptr, 0.0, (const u8 *)nullptr, isSplit, (u8 *)nullptr));
    CK(cudaMemsetAsync(ctx->dCount + 2, 0, 8, ctx->stream));
    if (p) LAUNCH(k_count_flags, gridFor(p), kThreads, p, (const u8 *)isSplit, (unsigned long long *)(ctx->dCount + 2));
    CK(cudaMemcpyAsync(ctx->hCount + 2, ctx->dCount + 2, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int64_t nSp = ctx->hCount[2];
    TRY(devAlloc(ctx, &ctx->spPoint, nSp)); TRY(devAlloc(ctx, &ctx->spCells, 8 * nSp)); TRY(devAlloc(ctx, &ctx->cellSp, n + 4));
    TRY(devAlloc(ctx, &ctx->spFlag, nSp + 16)); TRY(devAlloc(ctx, &ctx->spList, nSp));
    CK(cudaMemsetAsync(ctx->cellSp, 0xFF, (size_t)(n + 4) * 4, ctx->stream));
    int32_t *dTotal = nullptr;
    TRY(compactFlags(ctx, isSplit, p, ctx->spPoint, &dTotal));
    if (nSp) LAUNCH(k_sp_table, gridFor(nSp), kThreads, nSp, (const label *)ctx->spPoint, ctx->pc.sliceOff, ctx->pc.col, ctx->spCells, ctx->cellSp);
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->nSp = nSp;
    ctx->spDirty = false;
    return AMR_OK;
}

// hexRef3D::selectUnrefineElems + consistentUnrefinement on the split-point table (kernels_sparse.cuh (3)) + the
// ascending list (hexRef3D.C:1930-1950).  Same speculation as refineClosureQueue: one host look per call.
static int unrefineClosureQueue(amr_ctx *ctx, const double *dErr, double unrefineLevel, const u8 *marked, int *sweepsOut,
                                int32_t *elemsOut, int64_t *nOut) {
    const int64_t n = ctx->n, nSp = ctx->nSp;
    TRY(ensureQueues(ctx, std::max(n, nSp) + 16));
    const bool multi = ctx->comm && ctx->nranks > 1;
    const bool speculate = gatedSwapsOk(ctx);
    int *qc = ctx->dQ, *gq = multi ? ctx->dGQ : ctx->dQ;
    const int64_t cap = ctx->queueCap;
    const label *cf = ctx->cplFace, *bo = ctx->bOwner;
    u8 *uc = ctx->flagB, *spFlag = ctx->spFlag;
    ListJob job;
    TRY(listTarget(ctx, std::max<int64_t>(nSp, 1), elemsOut, &job));
    CK(cudaMemsetAsync(uc, 0, (size_t)n, ctx->stream));      // (queue sizes were cleared by sitesBegin at the start of the call)
    CK(cudaMemsetAsync(spFlag, 1, (size_t)nSp, ctx->stream));
    // U1 + U3: pFld < unrefineLevel and no marked cell, cell-centric (the split points themselves are mesh-constant)
    if (n && nSp) LAUNCH(k_sp_filter, gridFor(n, kThreads * 4), kThreads, n, (const label *)ctx->cellSp, dErr, unrefineLevel, marked, spFlag);
    // the candidates as a compact list; unrefineCell = their cells
    int32_t *nList = nullptr;
    TRY(compactFlagsEx(ctx, spFlag, nSp, nullptr, nullptr, nullptr, ctx->spList, &nList, 1, nullptr));
    const unsigned qg = queueGrid(ctx);
    if (nSp) LAUNCH(k_sp_to_cells, qg, kThreads, (const label *)ctx->spList, (const int32_t *)nList, nSp, (const label *)ctx->spCells, uc);
    const bool linked = multi && speculate && linkedOk(ctx);
    const unsigned nFix = linked ? gridFor(std::max<int64_t>(ctx->nCoupled, 1)) : 0u;
    const LinkOut L = linked ? makeLinkOut(ctx) : LinkOut();
    // sweep 0
    if (linked) {
        HaloToken halo;
        if (ctx->nCoupled > 0) TRY(haloBegin(ctx, ctx->lvl8, 1, ctx->ghost8, &halo, uc));     // cellLevel - unrefineCell (hexRef3D.C:1843-1853)
        GhostLink G = makeLink(ctx, &halo, ctx->ghost8, true);
        LAUNCH((k_unref_linked<true>), nFix + qg, kThreads, G, L, nFix, ctx->nCoupled, cf, bo, (const u8 *)ctx->lvl8, nSp, (const label *)ctx->spCells,
               (const label *)ctx->cellSp, spFlag, uc, (const label *)ctx->spList, (const int *)nList, ctx->cc.sliceOff, ctx->cc.col, ctx->queue[1], qc + 1,
               cap, gq + 1, (const int *)nullptr);
        ctx->peerTouched = true;
    } else {
        HaloToken halo;
        if (ctx->nCoupled > 0) TRY(haloBegin(ctx, ctx->lvl8, 1, ctx->ghost8, &halo, uc));     // cellLevel - unrefineCell (hexRef3D.C:1843-1853)
        if (nSp) LAUNCH(k_unref_sweep0, qg, kThreads, (const label *)ctx->spList, (const int32_t *)nList, nSp, (const label *)ctx->spCells,
                        ctx->cc.sliceOff, ctx->cc.col, (const u8 *)ctx->lvl8, spFlag, uc, ctx->queue[1], qc + 1, cap);
        if (multi) {                  // coupled faces (hexRef3D.C:1855-1889) + reduce(nChanged): one launch on the peer transport
            GhostLink G = makeLink(ctx, &halo, ctx->ghost8, true);
            LAUNCH(k_fix_unref_queue_link, gridFor(std::max<int64_t>(ctx->nCoupled, 1)), kThreads, G, ctx->nCoupled, cf, bo, (const u8 *)ctx->lvl8, nSp,
                   (const label *)ctx->spCells, (const label *)ctx->cellSp, spFlag, uc, ctx->queue[1], qc + 1, cap, gq + 1, (const int *)nullptr);
            if (!G.peer) TRY(allReduceWord(ctx, gq + 1, 4, qc + 1));
        }
    }
    int issued = 1;
    int batch = speculate ? std::max(1, ctx->itersSeen[1] - 1) : 0;
    int iters = 0;
    while (true) {
        for (int k = 0; k < batch && issued < kMaxIter; k++) {
            const int j = issued;
            const int *gate = gq + j;
            const int *xgate = speculate ? gate : nullptr;
            if (linked) {
                GhostLink G = makeLink(ctx, nullptr, ctx->ghost8, true);
                G.ghost = nullptr;
                LAUNCH((k_unref_linked<false>), nFix + qg, kThreads, G, L, nFix, ctx->nCoupled, cf, bo, (const u8 *)ctx->lvl8, nSp, (const label *)ctx->spCells,
                       (const label *)ctx->cellSp, spFlag, uc, (const label *)ctx->queue[j & 1], (const int *)(qc + j), ctx->cc.sliceOff, ctx->cc.col,
                       ctx->queue[(j + 1) & 1], qc + j + 1, cap, gq + j + 1, gate);
                ctx->peerTouched = true;
                issued++;
                continue;
            }
            HaloToken halo;
            if (ctx->nCoupled > 0) TRY(haloBegin(ctx, ctx->lvl8, 1, ctx->ghost8, &halo, uc, xgate));
            LAUNCH(k_unref_frontier, qg, kThreads, (const label *)ctx->queue[j & 1], (const int *)(qc + j), nSp, (const label *)ctx->spCells,
                   (const label *)ctx->cellSp, ctx->cc.sliceOff, ctx->cc.col, (const u8 *)ctx->lvl8, spFlag, uc, ctx->queue[(j + 1) & 1], qc + j + 1, cap, gate);
            if (multi) {
                GhostLink G = makeLink(ctx, &halo, ctx->ghost8, true);
                LAUNCH(k_fix_unref_queue_link, gridFor(std::max<int64_t>(ctx->nCoupled, 1)), kThreads, G, ctx->nCoupled, cf, bo, (const u8 *)ctx->lvl8, nSp,
                       (const label *)ctx->spCells, (const label *)ctx->cellSp, spFlag, uc, ctx->queue[(j + 1) & 1], qc + j + 1, cap, gq + j + 1, xgate);
                if (!G.peer) TRY(allReduceWord(ctx, gq + j + 1, 4, qc + j + 1, xgate));
            }
            issued++;
        }
        // ascending point list of the surviving candidates (speculative: valid if the loop has converged)
        const MailArgs mail = queueMail(ctx, gq, issued);
        TRY(compactFlagsEx(ctx, spFlag, nSp, nList, (const label *)ctx->spList, (const label *)ctx->spPoint, job.dList, &job.dTotal, 0, nullptr, &mail));
        TRY(queueLook(ctx, issued));
        if (ctx->lookResult) { iters = ctx->lookResult; break; }
        if (issued >= kMaxIter) FAIL(AMR_ERR_INVARIANT, "unrefinement closure did not settle");
        batch = speculate ? 2 : 1;
    }
    ctx->itersSeen[1] = iters;
    // hexRef3D.C:1808-1833 "problem": only an unbalanced mesh can produce it
    TRY(ensureBalanceChecked(ctx));
    {
        // every rank takes the same branch: balancedLocal is agreed over the ranks, balancedGlobal is set by all together
        bool check = ctx->balancedLocal == 0;
        if (multi && ctx->balancedGlobal < 0) check = true;                  // coupled faces not verified yet: look once
        if (check) {
            CK(cudaMemsetAsync(ctx->dFlags + 1, 0, sizeof(int), ctx->stream));
            if (ctx->nCoupled > 0) TRY(haloSwapCells(ctx, ctx->lvl8, 1, ctx->ghost8, uc));
            if (n) LAUNCH(k_unref_problem, gridFor(n), kThreads, n, ctx->cc.sliceOff, ctx->cc.col, (const u8 *)ctx->lvl8, (const u8 *)uc, ctx->dFlags + 1);
            if (ctx->nCoupled > 0)
                LAUNCH(k_fix_unref_problem, gridFor(ctx->nCoupled), kThreads, ctx->nCoupled, cf, bo, (const u8 *)ctx->lvl8, (const u8 *)ctx->ghost8,
                       (const u8 *)uc, ctx->dFlags + 1);
            int problem = 0;
            TRY(globalFlag(ctx, 1, &problem));
            if (problem) FAIL(AMR_ERR_INVARIANT, "problem (hexRef3D.C:1811): an unmarked cell is more than one level coarser than a face neighbour");
            if (ctx->balancedLocal == 1) ctx->balancedGlobal = 1;          // holds for every marking: levels alone decide it
        }
    }
    if (nOut) *nOut = ctx->hFlags[8];
    TRY(finishListCopy(ctx, &job));
    if (sweepsOut) *sweepsOut = iters;
    return AMR_OK;
}

// hexRef3D::consistentUnrefinement loop (hexRef3D.C:1758-1927) on point (edge) flags + ascending list
// (hexRef3D.C:1930-1950): the generic form for split elements given by a point->cells / edge->cells addressing
// (hexRef2D, polyRefiner).  One host round trip per sweep.
static int unrefineSweeps(amr_ctx *ctx, const Sell &EL, u8 *pflag, int *sweepsOut, int32_t *elemsOut, int64_t *nOut,
                          bool edgeRule = false) {
    const int64_t n = ctx->n, p = EL.rows;
    // polyRefiner: the edge rule (refinement.C:511-604) + the face rule over the internal faces are ONE cell-centric rule on
    // the graph of cells that share an edge (sell.cuh), walked for the marked cells only -- no pass over all mesh edges
    const Sell *A = &ctx->cc;
    bool edgeKernel = edgeRule && ctx->nEdges;
    if (edgeKernel && ctx->lvl) {
        TRY(ensureEdgeNeighbours(ctx));
        if (ctx->ceState == 1) { A = &ctx->ce; edgeKernel = false; }
    }
    u8 *uc = ctx->flagB;     // unrefineCell
    const bool alP = (((uintptr_t)pflag) & 15) == 0, alC = (((uintptr_t)uc) & 15) == 0;
    const unsigned gP = gridFor(p, kThreads * 16), gC = gridFor(n, kThreads * 16);
    int sweeps = 0;
    while (true) {
        CK(cudaMemsetAsync(uc, 0, (size_t)n, ctx->stream));
        if (p) LAUNCH(k_points_to_cells, gP, kThreads, p, EL.sliceOff, EL.col, pflag, alP, uc, (const u8 *)nullptr);
        CK(cudaMemsetAsync(ctx->dFlags, 0, sizeof(int), ctx->stream));
        if (ctx->nCoupled > 0) {
            TRY(haloSwapCells(ctx, ctx->lvl8, 1, ctx->ghost8, uc));
        }
        if (n) LAUNCH(k_unrefine_sweep, gC, kThreads, n, A->sliceOff, A->col, ctx->lvl8, uc, alC, ctx->dFlags);
        if (ctx->nCoupled > 0)      // hexRef3D.C:1855-1889
            LAUNCH(k_fix_unrefine_sweep, gridFor(ctx->nCoupled), kThreads, ctx->nCoupled, (const label *)ctx->cplFace,
                   (const label *)ctx->bOwner, (const u8 *)ctx->lvl8, (const u8 *)ctx->ghost8, uc, ctx->dFlags);
        // polyhedralRefinement.C:2278-2286: faceConsistentUnrefinement, then edgeConsistentUnrefinement
        if (edgeKernel)
            LAUNCH(k_edge_unrefine_sweep, gridFor(ctx->nEdges), kThreads, ctx->nEdges, ctx->ec.sliceOff, ctx->ec.col, (const u8 *)ctx->lvl8, uc, ctx->dFlags);
        sweeps++;
        int changed = 0;
        TRY(globalFlag(ctx, 0, &changed));
        if (!changed) break;
        if (p) LAUNCH(k_knock_points, gP, kThreads, p, EL.sliceOff, EL.col, pflag, alP, uc);
        if (sweeps > 100000) FAIL(AMR_ERR_INVARIANT, "unrefinement sweeps did not converge");
    }
    ListJob job;
    TRY(emitListAsync(ctx, pflag, p, elemsOut, &job));
    TRY(finishList(ctx, &job, -1, nullptr, nOut, true));
    if (sweepsOut) *sweepsOut = sweeps;
    return AMR_OK;
}

// split-point markup of the two polyRefiner engines: polyhedralRefinement (faces) / prismatic2DRefinement (edges)
static int polySplitMarkup(amr_ctx *ctx, bool prismatic, u8 *bad);
// The markup depends on the mesh and pointLevel only: built once per mesh and engine (faces / edges / levels / point
// hand-overs invalidate it), like the hexRefiner's split-point table.
static int ensurePolyMarkup(amr_ctx *ctx, bool prismatic, const u8 **bad) {
    const int kind = prismatic ? 1 : 0;
    if (ctx->polyBadKind != kind || !ctx->polyBad) {
        devFree(ctx, &ctx->polyBad);
        TRY(devAlloc(ctx, &ctx->polyBad, ctx->p + 16));
        TRY(polySplitMarkup(ctx, prismatic, ctx->polyBad));
        ctx->polyBadKind = kind;
    }
    *bad = ctx->polyBad;
    return AMR_OK;
}
static int polySplitMarkup(amr_ctx *ctx, bool prismatic, u8 *bad) {
    const int64_t p = ctx->p;
    if (p) CK(cudaMemsetAsync(bad, 0, (size_t)p, ctx->stream));
    if (prismatic) {
        if (ctx->nEdges) LAUNCH(k_edge_point_levels, gridFor(ctx->nEdges), kThreads, ctx->nEdges, (const label *)ctx->edgePts,
                                (const label *)ctx->pointLevel, bad);
        if (ctx->nCoupled > 0) LAUNCH(k_coupled_face_points, gridFor(ctx->nCoupled), kThreads, ctx->nCoupled, (const label *)ctx->cplFace, ctx->f,
                                      (const label *)ctx->faceOff, (const label *)ctx->facePts, bad);
    } else {
        const int64_t nf = ctx->f + ctx->b;
        if (nf) LAUNCH(k_face_point_levels, gridFor(nf), kThreads, nf, ctx->f, (const label *)ctx->faceOff, (const label *)ctx->facePts,
                       (const label *)ctx->pointLevel, bad);
    }
    return AMR_OK;
}

AMR_API int amr_select_unrefine(amr_ctx *ctx, const double *error, double unrefineLevel, uint8_t *refineCellInOut,
                                int nBufferLayers, const int32_t *protectedPatchIds, int nProtectedPatches,
                                int refinerKind, int32_t *elemsOut, int64_t *nOut, int *sweepsOut) {
    ENTER();
    if (!ctx->meshSet) FAIL(AMR_ERR_STATE, "amr_select_unrefine: mesh not set");
    if (!ctx->havePoints) FAIL(AMR_ERR_STATE, "amr_select_unrefine: point addressing not set (amr_mesh_set_points)");
    const bool prismatic = refinerKind == AMR_REFINER_POLY2D;
    const bool poly = refinerKind == AMR_REFINER_POLY3D || prismatic;
    if (!ctx->haveLevels || (!poly && !ctx->haveHistory)) FAIL(AMR_ERR_STATE, "amr_select_unrefine: levels/history not set");
    if (refinerKind != AMR_REFINER_HEX3D && refinerKind != AMR_REFINER_HEX2D && !poly)
        FAIL(AMR_ERR_ARG, "amr_select_unrefine: unknown refiner kind");
    if (poly) {
        if (!ctx->haveFaces || !ctx->pointLevel) FAIL(AMR_ERR_STATE, "amr_select_unrefine: polyRefiner needs amr_mesh_set_faces and pointLevel");
        if ((ctx->optEdgeConsistency || prismatic) && !ctx->haveEdges)
            FAIL(AMR_ERR_STATE, "amr_select_unrefine: edgeBasedConsistency / prismatic2DRefinement need amr_mesh_set_edges (all mesh edges)");
        if (ctx->nCoupled > 0 && nBufferLayers > 0 && ctx->ppOff.empty())
            FAIL(AMR_ERR_STATE, "amr_select_unrefine: polyRefiner point buffer layers across processor patches need the patches' point lists (amr_mesh_set_patch_points)");
    }
    if (refinerKind == AMR_REFINER_HEX2D && !ctx->haveEdges) FAIL(AMR_ERR_STATE, "amr_select_unrefine: hexRef2D needs amr_mesh_set_edges");
    if (nProtectedPatches > 0 && !protectedPatchIds) FAIL(AMR_ERR_ARG, "amr_select_unrefine: protectedPatchIds is null");
    if (nBufferLayers > 200) FAIL(AMR_ERR_ARG, "amr_select_unrefine: more than 200 buffer layers");
    const int64_t n = ctx->n, p = ctx->p;
    const double *dErr = nullptr;
    TRY(stageIn(ctx, error, n, &dErr));
    if (!dErr) dErr = ctx->error;
    u8 *dRc = nullptr;
    TRY(stageInOut(ctx, refineCellInOut, n, &dRc));
    TRY(sitesBegin(ctx));
    // the marking lives in the resident buffer; the layers grow it in place
    u8 *cur = ctx->refineCell;
    int gen = ctx->refineCellGen;
    if (dRc) {
        if (n) LAUNCH(k_flags_normalize, gridFor(n, kThreads * 16), kThreads, n, (const u8 *)dRc, cur, ((((uintptr_t)cur) | ((uintptr_t)dRc)) & 15) == 0);
        ctx->refineCellN = n;
        gen = 1;
    } else if (ctx->refineCellN != n) {
        FAIL(AMR_ERR_STATE, "amr_select_unrefine: resident marking has the wrong size (amr_remap_refine_cell after the topology change)");
    }
    if (gen + nBufferLayers >= 250) {
        if (n) LAUNCH(k_flags_normalize, gridFor(n, kThreads * 16), kThreads, n, (const u8 *)cur, cur, (((uintptr_t)cur) & 15) == 0);
        gen = 1;
    }
    // nUnrefinementBufferLayers x extendMarkedCellsAcrossFaces, fvMeshHexRefiner.C:1595-1598
    if (poly) {
        // extendMarkedCellsAcrossPoints (fvMeshRefiner.C:924-977): cells -> points -> cells, in place (flags only grow)
        for (int i = 0; i < nBufferLayers; i++) {
            // flagA: the point still has an unmarked cell on this rank -- only those points have anything to scatter (the marking
            // covers whole regions, the points that change it lie on its boundary)
            if (p) LAUNCH(k_cells_to_points_or, gridFor(p), kThreads, p, ctx->pc.sliceOff, ctx->pc.col, (const u8 *)cur, ctx->flagC, ctx->flagA);
            TRY(syncPointsOr(ctx, ctx->flagC));                      // syncTools::syncPointList, fvMeshRefiner.C:952-958
            if (p) LAUNCH(k_points_to_cells, gridFor(p, kThreads * 16), kThreads, p, ctx->pc.sliceOff, ctx->pc.col, (const u8 *)ctx->flagC,
                          (((uintptr_t)ctx->flagC) & 15) == 0, cur, (const u8 *)ctx->flagA);
        }
    } else {
        for (int i = 0; i < nBufferLayers; i++) {
            TRY(extendLayer(ctx, cur, nullptr, gen, siteOf(kSiteUnrefine, i)));
            gen++;
        }
    }
    ctx->refineCellGen = gen;
    // protections, fvMeshHexRefiner.C:1600-1622
    if (!poly && ctx->protCell && n) LAUNCH(k_masked_set, gridFor(n), kThreads, n, ctx->protCell, cur, (u8)1);
    for (int k = 0; k < nProtectedPatches; k++) {
        int q = protectedPatchIds[k];
        if (q < 0 || q >= (int)ctx->patches.size()) FAIL(AMR_ERR_ARG, "amr_select_unrefine: patch id out of range");
        const Patch &pp = ctx->patches[q];
        if (pp.size) LAUNCH(k_patch_cells_set, gridFor(pp.size), kThreads, ctx->bOwner, (int64_t)pp.start - ctx->f, (int64_t)pp.size, cur, (u8)1);
    }
    if (dRc && n) LAUNCH(k_flags_normalize, gridFor(n, kThreads * 16), kThreads, n, (const u8 *)cur, dRc, ((((uintptr_t)cur) | ((uintptr_t)dRc)) & 15) == 0);
    const u8 *marked = cur;
    // U1+U2+U3 fused: maxCellField + getSplitElems + filter
    u8 *pflag = ctx->flagC;
    if (poly) {
        const u8 *bad = nullptr;
        TRY(ensurePolyMarkup(ctx, prismatic, &bad));
        if (p) LAUNCH(k_poly_select_points, gridFor(p), kThreads, p, ctx->pc.sliceOff, ctx->pc.col, (const label *)ctx->pointLevel, (const u8 *)bad,
                      dErr, unrefineLevel, marked, (u8 *)nullptr, pflag);
        TRY(unrefineSweeps(ctx, ctx->pc, pflag, sweepsOut, elemsOut, nOut, ctx->optEdgeConsistency != 0));
        CK(cudaGetLastError());
        return flushOuts(ctx);
    }
    if (refinerKind == AMR_REFINER_HEX2D) {
        const int64_t ne = ctx->nEdges;
        if (ne) LAUNCH(k_select_unrefine_edges, gridFor(ne), kThreads, ne, ctx->ec.sliceOff, ctx->ec.col, ctx->ec.count, ctx->bndEdge,
                       ctx->edgePts, ctx->pc.sliceOff, ctx->pc.col, ctx->parentIdx, ctx->lvl8, dErr, unrefineLevel, marked, (u8 *)nullptr, pflag);
        TRY(unrefineSweeps(ctx, ctx->ec, pflag, sweepsOut, elemsOut, nOut));
        CK(cudaGetLastError());
        return flushOuts(ctx);
    }
    // hexRef3D: B3 on the mesh-constant split-point table
    TRY(ensureSplitTable(ctx));
    TRY(unrefineClosureQueue(ctx, dErr, unrefineLevel, marked, sweepsOut, elemsOut, nOut));
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_get_split_elems(amr_ctx *ctx, int refinerKind, int32_t *elemsOut, int64_t *nOut) {
    ENTER();
    if (!ctx->havePoints) FAIL(AMR_ERR_STATE, "amr_get_split_elems: points not set");
    const int64_t p = ctx->p;
    u8 *pflag = ctx->flagC;
    if (refinerKind == AMR_REFINER_POLY3D || refinerKind == AMR_REFINER_POLY2D) {
        if (!ctx->haveFaces || !ctx->pointLevel) FAIL(AMR_ERR_STATE, "amr_get_split_elems: polyRefiner needs amr_mesh_set_faces and pointLevel");
        if (refinerKind == AMR_REFINER_POLY2D && !ctx->haveEdges) FAIL(AMR_ERR_STATE, "amr_get_split_elems: prismatic2DRefinement needs amr_mesh_set_edges");
        u8 *bad = ctx->flagA;
        TRY(polySplitMarkup(ctx, refinerKind == AMR_REFINER_POLY2D, bad));
        if (p) LAUNCH(k_poly_select_points, gridFor(p), kThreads, p, ctx->pc.sliceOff, ctx->pc.col, (const label *)ctx->pointLevel, (const u8 *)bad,
                      (const double *)nullptr, 0.0, (const u8 *)nullptr, pflag, (u8 *)nullptr);
        TRY(emitList(ctx, pflag, p, elemsOut, nOut));
        CK(cudaGetLastError());
        return flushOuts(ctx);
    }
    if (!ctx->haveHistory) FAIL(AMR_ERR_STATE, "amr_get_split_elems: history not set");
    if (refinerKind != AMR_REFINER_HEX3D && refinerKind != AMR_REFINER_HEX2D) FAIL(AMR_ERR_ARG, "amr_get_split_elems: unknown refiner kind");
    if (refinerKind == AMR_REFINER_HEX2D) {
        if (!ctx->haveEdges) FAIL(AMR_ERR_STATE, "amr_get_split_elems: hexRef2D needs amr_mesh_set_edges");
        const int64_t ne = ctx->nEdges;
        if (ne) LAUNCH(k_select_unrefine_edges, gridFor(ne), kThreads, ne, ctx->ec.sliceOff, ctx->ec.col, ctx->ec.count, ctx->bndEdge,
                       ctx->edgePts, ctx->pc.sliceOff, ctx->pc.col, ctx->parentIdx, ctx->lvl8, (const double *)nullptr, 0.0,
                       (const u8 *)nullptr, pflag, (u8 *)nullptr);
        TRY(emitList(ctx, pflag, ne, elemsOut, nOut));
        CK(cudaGetLastError());
        return flushOuts(ctx);
    }
    // hexRef3D: the mesh-constant list
    TRY(ensureSplitTable(ctx));
    if (nOut) *nOut = ctx->nSp;
    if (elemsOut && ctx->nSp) {
        CK(cudaMemcpyAsync(elemsOut, ctx->spPoint, (size_t)ctx->nSp * 4, cudaMemcpyDefault, ctx->stream));
        if (!isDevicePtr(elemsOut)) ctx->hostTouched = true;
    }
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_max_cell_field(amr_ctx *ctx, const double *vFld, double *pFld) {
    ENTER();
    if (!ctx->havePoints) FAIL(AMR_ERR_STATE, "amr_max_cell_field: point addressing not set");
    if (!pFld) FAIL(AMR_ERR_ARG, "amr_max_cell_field: pFld is null");
    const double *dV = nullptr;
    double *dP = nullptr;
    TRY(stageIn(ctx, vFld, ctx->n, &dV));
    if (!dV) dV = ctx->error;
    TRY(stageOut(ctx, pFld, ctx->p, &dP));
    if (ctx->p) LAUNCH(k_max_cell_field, gridFor(ctx->p), kThreads, ctx->p, ctx->pc.sliceOff, ctx->pc.col, dV, dP);
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_check_levels(amr_ctx *ctx, const int32_t *cellLevel, int64_t *nBad) {
    ENTER();
    const int64_t n = ctx->n, b = ctx->b;
    const int32_t *dL = nullptr;
    TRY(stageIn(ctx, cellLevel, n, &dL));
    if (!dL) dL = ctx->lvl;
    if (!dL) FAIL(AMR_ERR_STATE, "amr_check_levels: no levels");
    label *ghost = nullptr;
    if (ctx->nCoupled > 0) {
        void *q;
        TRY(arenaAlloc(ctx, &q, (size_t)b * 4)); ghost = (label *)q;
        TRY(haloSwapCells(ctx, dL, 4, ghost));
    }
    CK(cudaMemsetAsync(ctx->dCount, 0, 8, ctx->stream));
    if (n) LAUNCH(k_check_levels, gridFor(n), kThreads, n, ctx->cc.sliceOff, ctx->cc.col, dL, (unsigned long long *)ctx->dCount);
    if (ctx->nCoupled > 0)
        LAUNCH(k_fix_check_levels, gridFor(ctx->nCoupled), kThreads, ctx->nCoupled, (const label *)ctx->cplFace, (const label *)ctx->bOwner,
               dL, (const label *)ghost, (unsigned long long *)ctx->dCount);
    CK(cudaMemcpyAsync(ctx->hCount, ctx->dCount, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (nBad) *nBad = ctx->hCount[0];
    CK(cudaGetLastError());
    if (ctx->hCount[0] > 0) FAIL(AMR_ERR_INVARIANT, "Celllevel does not satisfy 2:1 constraint (hexRef.C:2975)");
    return AMR_OK;
}

// ============================================================================ stage 4

// Host -> host mapping of a large field: the old field goes up in chunks on one stream while mapped
// chunks of the new field come down on another, so both PCIe directions are busy at once.  A chunk of
// new cells may start as soon as the old rows it reads (max of cellMap over the chunk) have arrived;
// for the maps of hexRef (identity prefix + appended children) that is almost immediately.
static int mapFieldsHostPipelined(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *cellMap, int nFields,
                                  const int *nComps, const double *const *oldFs, double *const *newFs) {
    const int64_t chunk = (int64_t)1 << 22;                    // 4 M cells per chunk
    const int64_t nUp = (nOld + chunk - 1) / chunk, nDown = (nNew + chunk - 1) / chunk;
    if (!ctx->upStream) CK(cudaStreamCreateWithFlags(&ctx->upStream, cudaStreamNonBlocking));
    if (!ctx->downStream) CK(cudaStreamCreateWithFlags(&ctx->downStream, cudaStreamNonBlocking));
    const int32_t *dMap = nullptr;
    TRY(stageIn(ctx, cellMap, nNew, &dMap));
    void *q;
    TRY(arenaAlloc(ctx, &q, (size_t)(nDown + 1) * 4)); int *dMax = (int *)q;
    CK(cudaMemsetAsync(dMax, 0xff, (size_t)(nDown + 1) * 4, ctx->stream));      // -1
    LAUNCH(k_chunk_max_src, gridFor(nNew), kThreads, nNew, chunk, dMap, dMax);
    std::vector<int> hMax((size_t)nDown);
    CK(cudaMemcpyAsync(hMax.data(), dMax, (size_t)nDown * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int64_t j = 0; j < nDown; j++)
        if (hMax[(size_t)j] >= nOld) FAIL(AMR_ERR_ARG, "amr_map_field: cellMap entry out of range");
    std::vector<cudaEvent_t> ev;
    int rc = AMR_OK;
    for (int fi = 0; fi < nFields && rc == AMR_OK; fi++) {
        const int nComp = nComps[fi];
        const double *oldF = oldFs[fi];
        double *newF = newFs[fi];
        double *dOld = nullptr, *dNew = nullptr;
        if ((rc = arenaAlloc(ctx, &q, (size_t)nOld * nComp * 8)) != AMR_OK) break;
        dOld = (double *)q;
        if ((rc = arenaAlloc(ctx, &q, (size_t)nNew * nComp * 8)) != AMR_OK) break;
        dNew = (double *)q;
        const size_t e0 = ev.size();
        for (int64_t k = 0; k < nUp; k++) {
            const int64_t r0 = k * chunk, r1 = std::min(nOld, r0 + chunk);
            CK(cudaMemcpyAsync(dOld + r0 * nComp, oldF + r0 * nComp, (size_t)(r1 - r0) * nComp * 8, cudaMemcpyHostToDevice, ctx->upStream));
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            CK(cudaEventRecord(e, ctx->upStream));
            ev.push_back(e);
        }
        int64_t waited = -1;
        for (int64_t j = 0; j < nDown; j++) {
            const int64_t i0 = j * chunk, i1 = std::min(nNew, i0 + chunk);
            const int64_t need = hMax[(size_t)j] < 0 ? -1 : hMax[(size_t)j] / chunk;     // last upload chunk this one reads
            for (int64_t k = waited + 1; k <= need; k++) CK(cudaStreamWaitEvent(ctx->downStream, ev[e0 + (size_t)k], 0));
            if (need > waited) waited = need;
            const int64_t items = (i1 - i0) * nComp;
            k_map_field_range<<<gridFor(items, kThreads * kMapItems), kThreads, 0, ctx->downStream>>>(i0, i1, nComp, dMap, dOld, dNew);
            ctx->launches++;
            CK(cudaMemcpyAsync(newF + i0 * nComp, dNew + i0 * nComp, (size_t)items * 8, cudaMemcpyDeviceToHost, ctx->downStream));
        }
    }
    cudaStreamSynchronize(ctx->upStream);
    cudaStreamSynchronize(ctx->downStream);
    for (auto &e : ev) cudaEventDestroy(e);
    if (rc != AMR_OK) return rc;
    CK(cudaGetLastError());
    return AMR_OK;
}

// fvMesh::mapFields maps EVERY registered field through one mapPolyMesh (MapGeometricFields over the object
// registry): the batched entry point shares the cellMap and, for host buffers, keeps uploading field k+1 while
// field k comes down.
// fvMesh::mapFields direct injection of one field on the device
static int launchMapField(amr_ctx *ctx, int64_t nNew, int nComp, const int32_t *dMap, const double *dOld, double *dNew) {
    const int64_t items = nNew * nComp;
    if (items <= 0) return AMR_OK;
    const bool al16 = ((((uintptr_t)dOld) | ((uintptr_t)dNew)) & 15) == 0;
    const unsigned gE = gridFor(items, kThreads * kMapItems), gP = gridFor((items + 1) / 2, kThreads * kMapItems);
    switch (nComp) {
    case 1: LAUNCH((k_map_field<1>), gE, kThreads, nNew, dMap, dOld, dNew); break;
    case 3:
        if (al16) LAUNCH((k_map_field_pair<3>), gP, kThreads, nNew, dMap, dOld, dNew);
        else LAUNCH((k_map_field<3>), gE, kThreads, nNew, dMap, dOld, dNew);
        break;
    case 6:
        if (al16) LAUNCH((k_map_field_pair<6>), gP, kThreads, nNew, dMap, dOld, dNew);
        else LAUNCH((k_map_field<6>), gE, kThreads, nNew, dMap, dOld, dNew);
        break;
    case 9:
        if (al16) LAUNCH((k_map_field_pair<9>), gP, kThreads, nNew, dMap, dOld, dNew);
        else LAUNCH((k_map_field<9>), gE, kThreads, nNew, dMap, dOld, dNew);
        break;
    default: LAUNCH(k_map_field_n, gE, kThreads, nNew, nComp, dMap, dOld, dNew); break;
    }
    return AMR_OK;
}

// all fields (<= kMaxMapFields per launch) and, optionally, the marking through the map in one pass (k_map_fields_fused)
static int launchMapFused(amr_ctx *ctx, int64_t nNew, const int32_t *dMap, int nFields, const int *nComp, const double *const *dOld,
                          double *const *dNew, const int32_t *dRev, const u8 *dOldFlag, u8 *dNewFlag) {
    if (nNew <= 0) return AMR_OK;
    if (ctx->mapForm == 2) {          // one launch per field (re-reads cellMap): AMR_B200_MAP=separate
        for (int i = 0; i < nFields; i++) TRY(launchMapField(ctx, nNew, nComp[i], dMap, dOld[i], dNew[i]));
        if (dNewFlag) LAUNCH(k_remap_refine_cell, gridFor(nNew, kThreads * 4), kThreads, nNew, dMap, dRev, dOldFlag, dNewFlag);
        return AMR_OK;
    }
    const int64_t chunks = (nNew + kMapChunk - 1) / kMapChunk;
    int done = 0;
    bool flagsDone = dNewFlag == nullptr;
    while (done < nFields || !flagsDone) {
        MapFieldSet F;
        F.nFields = 0;
        while (done < nFields && F.nFields < kMaxMapFields) {
            F.nc[F.nFields] = nComp[done]; F.oldF[F.nFields] = dOld[done]; F.newF[F.nFields] = dNew[done];
            F.nFields++; done++;
        }
        for (int k = F.nFields; k < kMaxMapFields; k++) { F.nc[k] = 0; F.oldF[k] = nullptr; F.newF[k] = nullptr; }
        const bool withFlags = !flagsDone;
        const int nStreams = F.nFields + (withFlags ? 1 : 0);
        if (chunks * nStreams > 0x7fffffff) FAIL(AMR_ERR_ARG, "amr_map_fields: too many cells for one launch");
        if (ctx->mapForm == 1)
            LAUNCH((k_map_fields_fused<6>), (unsigned)(chunks * nStreams), kThreads, nNew, dMap, F, nStreams, withFlags ? dRev : (const int32_t *)nullptr,
                   withFlags ? dOldFlag : (const u8 *)nullptr, withFlags ? dNewFlag : (u8 *)nullptr);
        else
            LAUNCH((k_map_fields_fused<1>), (unsigned)(chunks * nStreams), kThreads, nNew, dMap, F, nStreams, withFlags ? dRev : (const int32_t *)nullptr,
                   withFlags ? dOldFlag : (const u8 *)nullptr, withFlags ? dNewFlag : (u8 *)nullptr);
        flagsDone = true;
    }
    return AMR_OK;
}

AMR_API int amr_map_fields_and_marking(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *cellMap, int nFields, const int *nComp,
                                       const double *const *oldF, double *const *newF, const int32_t *reverseCellMap,
                                       const uint8_t *refineCell, uint8_t *newRefineCellOut) {
    ENTER();
    if (!cellMap || nFields < 0 || nOld < 0 || nNew < 0 || (nFields > 0 && (!nComp || !oldF || !newF))) FAIL(AMR_ERR_ARG, "amr_map_fields: bad arguments");
    const bool marking = reverseCellMap != nullptr;
    bool allHost = nFields > 0;
    for (int i = 0; i < nFields; i++) {
        if (!oldF[i] || !newF[i] || nComp[i] < 1) FAIL(AMR_ERR_ARG, "amr_map_fields: null field / bad nComp");
        if (isDevicePtr(oldF[i]) || isDevicePtr(newF[i])) allHost = false;
    }
    if (marking && !refineCell && nOld != ctx->refineCellN) FAIL(AMR_ERR_STATE, "amr_map_fields_and_marking: resident marking has a different size than nOld");
    if (allHost) {
        // solver-owned host fields: chunked two-way PCIe pipeline; the marking (device resident) goes through the map on its own
        TRY(mapFieldsHostPipelined(ctx, nOld, nNew, cellMap, nFields, nComp, oldF, newF));
        if (!marking) return AMR_OK;
        nFields = 0;
    }
    const int32_t *dMap = nullptr, *dRev = nullptr;
    const u8 *dOldFlag = nullptr;
    u8 *dNewFlag = nullptr;
    TRY(stageIn(ctx, cellMap, nNew, &dMap));
    std::vector<const double *> dOld((size_t)std::max(nFields, 1), nullptr);
    std::vector<double *> dNew((size_t)std::max(nFields, 1), nullptr);
    for (int i = 0; i < nFields; i++) {
        TRY(stageIn(ctx, oldF[i], nOld * nComp[i], &dOld[(size_t)i]));
        TRY(stageOut(ctx, newF[i], nNew * nComp[i], &dNew[(size_t)i]));
    }
    bool oversize = false;
    if (marking) {
        TRY(stageIn(ctx, reverseCellMap, nOld, &dRev));
        TRY(stageIn(ctx, refineCell, nOld, &dOldFlag));
        if (!dOldFlag) dOldFlag = ctx->refineCell;
        if (!dOldFlag) FAIL(AMR_ERR_STATE, "amr_map_fields_and_marking: no marking");
        if (nNew + 16 <= ctx->flagCap) dNewFlag = ctx->flagA;        // becomes the resident marking by pointer swap
        else {
            CK(cudaMalloc((void **)&dNewFlag, (size_t)std::max<int64_t>(nNew + 16, ctx->flagCap)));
            oversize = true;
        }
    }
    TRY(launchMapFused(ctx, nNew, dMap, nFields, nComp, dOld.data(), dNew.data(), dRev, dOldFlag, dNewFlag));
    if (marking) {
        if (newRefineCellOut && nNew) {
            CK(cudaMemcpyAsync(newRefineCellOut, dNewFlag, (size_t)nNew, cudaMemcpyDefault, ctx->stream));
            if (!isDevicePtr(newRefineCellOut)) ctx->hostTouched = true;
        }
        if (oversize) {
            // the larger buffer leaves the pool convention; amr_mesh_set of the new mesh adopts its content
            CK(cudaStreamSynchronize(ctx->stream));
            if (ctx->refineCell) cudaFree(ctx->refineCell);
            ctx->refineCell = dNewFlag;
            ctx->refineCellOversize = true;
        } else std::swap(ctx->flagA, ctx->refineCell);
        ctx->refineCellN = nNew;
        ctx->refineCellGen = 1;
    }
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

// fvMesh::mapFields maps EVERY registered field through one mapPolyMesh (MapGeometricFields over the object
// registry): the batched entry point reads the cellMap once for all fields and, for host buffers, keeps uploading
// field k+1 while field k comes down.
AMR_API int amr_map_fields(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *cellMap, int nFields, const int *nComp,
                           const double *const *oldF, double *const *newF) {
    return amr_map_fields_and_marking(ctx, nOld, nNew, cellMap, nFields, nComp, oldF, newF, nullptr, nullptr, nullptr);
}

AMR_API int amr_map_field(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *cellMap, int nComp, const double *oldF,
                          double *newF, int mode, const int32_t *reverseCellMap, const double *volOld) {
    ENTER();
    if (!cellMap || !oldF || !newF || nComp < 1) FAIL(AMR_ERR_ARG, "amr_map_field: null pointer / bad nComp");
    if (mode == AMR_MAP_DIRECT && !isDevicePtr(oldF) && !isDevicePtr(newF) && nNew * nComp >= ((int64_t)1 << 23))
        return mapFieldsHostPipelined(ctx, nOld, nNew, cellMap, 1, &nComp, &oldF, &newF);
    const int32_t *dMap = nullptr;
    const double *dOld = nullptr;
    double *dNew = nullptr;
    TRY(stageIn(ctx, cellMap, nNew, &dMap));
    TRY(stageIn(ctx, oldF, nOld * nComp, &dOld));
    TRY(stageOut(ctx, newF, nNew * nComp, &dNew));
    TRY(launchMapField(ctx, nNew, nComp, dMap, dOld, dNew));
    if (mode == AMR_MAP_CONSERVATIVE) {
        if (!reverseCellMap || !volOld) FAIL(AMR_ERR_ARG, "amr_map_field: conservative mode needs reverseCellMap and volOld");
        // merge lists are tiny compared with the field (8 cells per unrefined parent): build the CSR on the host
        std::vector<int32_t> rev((size_t)nOld);
        CK(cudaMemcpyAsync(rev.data(), reverseCellMap, (size_t)nOld * 4, cudaMemcpyDefault, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        std::vector<int32_t> masters;
        std::vector<int64_t> pairs;   // (master << 32) | old cell
        std::vector<char> isMaster((size_t)(nNew > 0 ? nNew : 1), 0);
        for (int64_t c = 0; c < nOld; c++) if (rev[c] < -1) isMaster[(size_t)(-rev[c] - 2)] = 1;
        for (int64_t c = 0; c < nOld; c++) {
            int64_t t = rev[c] < -1 ? -rev[c] - 2 : rev[c];
            if (t >= 0 && isMaster[(size_t)t]) pairs.push_back((t << 32) | c);
        }
        std::sort(pairs.begin(), pairs.end());
        std::vector<int32_t> off, idx;
        for (size_t i = 0; i < pairs.size(); i++) {
            int32_t m = (int32_t)(pairs[i] >> 32);
            if (masters.empty() || masters.back() != m) { masters.push_back(m); off.push_back((int32_t)idx.size()); }
            idx.push_back((int32_t)(pairs[i] & 0xffffffff));
        }
        off.push_back((int32_t)idx.size());
        if (!masters.empty()) {
            const int32_t *dM = nullptr, *dO = nullptr, *dI = nullptr;
            const double *dV = nullptr;
            TRY(stageIn(ctx, masters.data(), (int64_t)masters.size(), &dM));
            TRY(stageIn(ctx, off.data(), (int64_t)off.size(), &dO));
            TRY(stageIn(ctx, idx.data(), (int64_t)idx.size(), &dI));
            TRY(stageIn(ctx, volOld, nOld, &dV));
            int64_t it = (int64_t)masters.size() * nComp;
            LAUNCH(k_map_conservative, gridFor(it), kThreads, (int64_t)masters.size(), dM, dO, dI, dV, nComp, dOld, dNew);
            CK(cudaStreamSynchronize(ctx->stream));   // host vectors must outlive the staged copies
        }
    }
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_update_levels(amr_ctx *ctx, int64_t nNew, const int32_t *cellMap, const int32_t *oldLevel,
                              const uint8_t *refinedOld, const uint8_t *unrefinedOld, int32_t *newLevel) {
    ENTER();
    if (!cellMap || !newLevel) FAIL(AMR_ERR_ARG, "amr_update_levels: null pointer");
    const int64_t nOld = ctx->n;
    const int32_t *dMap = nullptr, *dOldL = nullptr;
    const u8 *dR = nullptr, *dU = nullptr;
    int32_t *dNew = nullptr;
    TRY(stageIn(ctx, cellMap, nNew, &dMap));
    TRY(stageIn(ctx, oldLevel, nOld, &dOldL));
    if (!dOldL) dOldL = ctx->lvl;
    if (!dOldL) FAIL(AMR_ERR_STATE, "amr_update_levels: no levels");
    TRY(stageIn(ctx, refinedOld, nOld, &dR));
    TRY(stageIn(ctx, unrefinedOld, nOld, &dU));
    TRY(stageOut(ctx, newLevel, nNew, &dNew));
    if (nNew) LAUNCH(k_update_levels, gridFor(nNew), kThreads, nNew, dMap, dOldL, dR, dU, dNew);
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

// ============================================================================ rows either side of the path (fields.cuh)

AMR_API int amr_remap_labels(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *map, int mode, const int32_t *oldValues,
                             int32_t *newValues) {
    ENTER();
    if (!map || !oldValues || !newValues) FAIL(AMR_ERR_ARG, "amr_remap_labels: null pointer");
    if (mode != AMR_REMAP_REORDER && mode != AMR_REMAP_GATHER) FAIL(AMR_ERR_ARG, "amr_remap_labels: unknown mode");
    const int32_t *dMap = nullptr, *dOld = nullptr;
    int32_t *dNew = nullptr;
    TRY(stageIn(ctx, map, mode == AMR_REMAP_REORDER ? nOld : nNew, &dMap));
    TRY(stageIn(ctx, oldValues, nOld, &dOld));
    TRY(stageOut(ctx, newValues, nNew, &dNew));
    if (mode == AMR_REMAP_REORDER) {
        if (nNew) CK(cudaMemsetAsync(dNew, 0xFF, (size_t)nNew * 4, ctx->stream));    // -1
        if (nOld) LAUNCH(k_reorder_labels, gridFor(nOld), kThreads, nOld, dMap, dOld, dNew);
    } else if (nNew) LAUNCH(k_gather_labels, gridFor(nNew), kThreads, nNew, dMap, dOld, dNew);
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_remap_protected_cells(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *cellMap, const uint8_t *oldProtected,
                                      uint8_t *newProtected) {
    ENTER();
    if (!cellMap || !oldProtected || !newProtected) FAIL(AMR_ERR_ARG, "amr_remap_protected_cells: null pointer");
    const int32_t *dMap = nullptr;
    const u8 *dOld = nullptr;
    u8 *dNew = nullptr;
    TRY(stageIn(ctx, cellMap, nNew, &dMap));
    TRY(stageIn(ctx, oldProtected, nOld, &dOld));
    TRY(stageOut(ctx, newProtected, nNew, &dNew));
    if (nNew) LAUNCH(k_gather_protected, gridFor(nNew), kThreads, nNew, dMap, dOld, dNew);
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_map_old_volumes(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *cellMap, const int32_t *reverseCellMap,
                                int nVolumes, const double *const *oldV, double *const *newV, int64_t *nMerged) {
    ENTER();
    if (!cellMap || !reverseCellMap || nOld < 0 || nNew < 0 || nVolumes < 0 || (nVolumes > 0 && (!oldV || !newV)))
        FAIL(AMR_ERR_ARG, "amr_map_old_volumes: bad arguments");
    const int32_t *dMap = nullptr, *dRev = nullptr;
    TRY(stageIn(ctx, cellMap, nNew, &dMap));
    TRY(stageIn(ctx, reverseCellMap, nOld, &dRev));
    std::vector<const double *> dOld((size_t)std::max(nVolumes, 1), nullptr);
    std::vector<double *> dNew((size_t)std::max(nVolumes, 1), nullptr);
    for (int k = 0; k < nVolumes; k++) {
        if (!oldV[k] || !newV[k]) FAIL(AMR_ERR_ARG, "amr_map_old_volumes: null volume field");
        TRY(stageIn(ctx, oldV[k], nOld, &dOld[(size_t)k]));
        TRY(stageOut(ctx, newV[k], nNew, &dNew[(size_t)k]));
        if (nNew) LAUNCH(k_map_old_volume, gridFor(nNew), kThreads, nNew, dMap, dOld[(size_t)k], dNew[(size_t)k]);
    }
    // masters of the merges, ascending, and the slot blocks of their merged cells
    void *q;
    const int64_t tiles = (nNew + kTile - 1) / kTile + 8;
    if (tiles > ctx->tileCap) {                       // the map may belong to a larger mesh than the one held by the context
        devFree(ctx, &ctx->lbStatus);
        ctx->tileCap = tiles;
        TRY(devAlloc(ctx, &ctx->lbStatus, ctx->tileCap));
        CK(cudaMemsetAsync(ctx->lbStatus, 0, (size_t)ctx->tileCap * 8, ctx->stream));
    }
    TRY(arenaAlloc(ctx, &q, (size_t)nNew + 16)); u8 *isMaster = (u8 *)q;
    TRY(arenaAlloc(ctx, &q, (size_t)(nNew > 0 ? nNew : 1) * 4)); label *masters = (label *)q;
    TRY(arenaAlloc(ctx, &q, (size_t)(nNew > 0 ? nNew : 1) * 4)); label *indexOf = (label *)q;
    CK(cudaMemsetAsync(isMaster, 0, (size_t)nNew + 16, ctx->stream));
    CK(cudaMemsetAsync(ctx->dCount + 3, 0, 8, ctx->stream));
    if (nOld) LAUNCH(k_merge_masters, gridFor(nOld), kThreads, nOld, dRev, isMaster, (unsigned long long *)(ctx->dCount + 3));
    int32_t *nM = nullptr;
    TRY(compactFlags(ctx, isMaster, nNew, masters, &nM));
    CK(cudaMemcpyAsync(ctx->hFlags + 8, nM, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->hCount + 3, ctx->dCount + 3, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int64_t nMasters = ctx->hFlags[8];
    if (nMerged) *nMerged = ctx->hCount[3];
    if (nMasters > 0 && nVolumes > 0) {
        const unsigned g = (unsigned)std::min<int64_t>((int64_t)ctx->numSMs * 8, (nMasters + kThreads - 1) / kThreads);
        LAUNCH(k_merge_index, g, kThreads, (const int32_t *)nM, (const label *)masters, indexOf);
        TRY(arenaAlloc(ctx, &q, (size_t)nMasters * 4)); int *cursor = (int *)q;
        int K = 8;                                    // hexRef3D merges 7 cells into the 8th; retried with the real maximum otherwise
        for (int attempt = 0; attempt < 2; attempt++) {
            TRY(arenaAlloc(ctx, &q, (size_t)nMasters * K * 4)); label *slots = (label *)q;
            CK(cudaMemsetAsync(cursor, 0, (size_t)nMasters * 4, ctx->stream));
            CK(cudaMemsetAsync(ctx->dFlags + 10, 0, sizeof(int), ctx->stream));
            LAUNCH(k_merge_slots, gridFor(nOld), kThreads, nOld, dRev, (const label *)indexOf, K, cursor, slots, ctx->dFlags + 10);
            CK(cudaMemcpyAsync(ctx->hFlags + 10, ctx->dFlags + 10, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (ctx->hFlags[10] > K) { K = ctx->hFlags[10]; continue; }
            for (int k = 0; k < nVolumes; k++)
                LAUNCH(k_merge_inject, g, kThreads, (const int32_t *)nM, (const label *)masters, K, (const int *)cursor, slots, dOld[(size_t)k], dNew[(size_t)k]);
            break;
        }
    }
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_correct_fluxes(amr_ctx *ctx, int64_t nOldFaces, const int32_t *faceMap, const int32_t *reverseFaceMap,
                               const double *U, const double *UBoundary, const double *nbrU, double *phi, int64_t *nCorrected) {
    ENTER();
    if (!ctx->haveGeom) FAIL(AMR_ERR_STATE, "amr_correct_fluxes: geometry of the new mesh not set (amr_mesh_set_geometry)");
    if (!faceMap || !reverseFaceMap || !U || !phi) FAIL(AMR_ERR_ARG, "amr_correct_fluxes: null pointer");
    const int64_t n = ctx->n, f = ctx->f, b = ctx->b, nf = f + b;
    const int32_t *dFm = nullptr, *dRf = nullptr;
    const double *dU = nullptr, *dUb = nullptr, *dUn = nullptr;
    double *dPhi = nullptr;
    TRY(stageIn(ctx, faceMap, nf, &dFm));
    TRY(stageIn(ctx, reverseFaceMap, nOldFaces, &dRf));
    TRY(stageIn(ctx, U, 3 * n, &dU));
    TRY(stageIn(ctx, UBoundary, 3 * b, &dUb));
    TRY(stageInOut(ctx, phi, nf, &dPhi));
    if (ctx->nCoupled > 0) {
        if (nbrU) TRY(stageIn(ctx, nbrU, 3 * b, &dUn));
        else {
            void *q;
            TRY(arenaAlloc(ctx, &q, (size_t)b * 24));
            TRY(haloSwapCells(ctx, dU, 24, q));
            dUn = (const double *)q;
        }
    }
    void *q;
    TRY(arenaAlloc(ctx, &q, (size_t)nf + 16));
    u8 *fix = (u8 *)q;
    CK(cudaMemsetAsync(fix, 0, (size_t)nf + 16, ctx->stream));
    CK(cudaMemsetAsync(ctx->dFlags + 10, 0, sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(ctx->dCount + 3, 0, 8, ctx->stream));
    if (nf) {
        LAUNCH(k_flux_mark, gridFor(nf), kThreads, nf, dFm, dRf, fix, ctx->dFlags + 10);
        LAUNCH(k_flux_apply, gridFor(nf), kThreads, nf, f, (const label *)ctx->owner, (const label *)ctx->neighbour,
               (const double *)ctx->gWeights, (const double *)ctx->gSf, dU, dUb, dUn, (const u8 *)ctx->bCoupled, (const u8 *)ctx->bEmpty,
               (const u8 *)fix, dPhi, (unsigned long long *)(ctx->dCount + 3));
    }
    CK(cudaMemcpyAsync(ctx->hFlags + 10, ctx->dFlags + 10, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->hCount + 3, ctx->dCount + 3, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->hFlags[10]) FAIL(AMR_ERR_INVARIANT, "Problem: should not have removed faces when refining.");   // adaptiveFvMesh.C:141-144
    if (nCorrected) *nCorrected = ctx->hCount[3];
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_mesh_set_ls_vectors(amr_ctx *ctx, const double *pVectors, const double *nVectors) {
    ENTER();
    if (!ctx->meshSet) FAIL(AMR_ERR_STATE, "amr_mesh_set_ls_vectors: call amr_mesh_set first");
    if (!pVectors || (!nVectors && ctx->f)) FAIL(AMR_ERR_ARG, "amr_mesh_set_ls_vectors: null pointer");
    const int64_t f = ctx->f, nf = ctx->f + ctx->b;
    devFree(ctx, &ctx->lsP); devFree(ctx, &ctx->lsN);
    TRY(devAlloc(ctx, &ctx->lsP, 3 * nf)); TRY(devAlloc(ctx, &ctx->lsN, 3 * f));
    if (nf) TRY(uploadAuto(ctx, ctx->lsP, pVectors, (size_t)nf * 24));
    if (f) TRY(uploadAuto(ctx, ctx->lsN, nVectors, (size_t)f * 24));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->haveLs = true;
    return AMR_OK;
}

AMR_API int amr_estimate_gradient(amr_ctx *ctx, int scheme, const double *x, const double *xBoundary, const double *nbrValues,
                                  double totalVolume, double refineThreshold, double unrefineThreshold, int scale, double *thrOut,
                                  double *gradOut, double *errorOut) {
    ENTER();
    if (!ctx->haveGeom) FAIL(AMR_ERR_STATE, "amr_estimate_gradient: geometry not set (amr_mesh_set_geometry)");
    if (scheme != AMR_GRAD_GAUSS_LINEAR && scheme != AMR_GRAD_LEAST_SQUARES) FAIL(AMR_ERR_ARG, "amr_estimate_gradient: unknown gradient scheme");
    if (scheme == AMR_GRAD_LEAST_SQUARES && !ctx->haveLs) FAIL(AMR_ERR_STATE, "amr_estimate_gradient: leastSquares vectors not set (amr_mesh_set_ls_vectors)");
    if (!x) FAIL(AMR_ERR_ARG, "amr_estimate_gradient: x is null");
    const int64_t n = ctx->n, f = ctx->f, b = ctx->b;
    const double *dx = nullptr, *dxb = nullptr, *dxn = nullptr;
    TRY(stageIn(ctx, x, n, &dx));
    TRY(stageIn(ctx, xBoundary, b, &dxb));
    const bool devOut = errorOut && isDevicePtr(errorOut);
    double *target = devOut ? errorOut : ctx->error;
    ctx->errorResident = !devOut;
    if (ctx->nCoupled > 0) {
        if (nbrValues) TRY(stageIn(ctx, nbrValues, b, &dxn));
        else {
            TRY(haloSwapCells(ctx, dx, 8, ctx->ghost64));
            dxn = ctx->ghost64;
        }
    } else dxn = ctx->ghost64;
    Geom G{ctx->gWeights, ctx->gSf, ctx->gMagSf, ctx->gDelta, ctx->gV};
    if (b) LAUNCH(k_boundary_field, gridFor(b), kThreads, b, (const label *)ctx->bOwner, dx, dxb, dxn, ctx->bCoupled, ctx->xbf);
    if (n) {
        if (scheme == AMR_GRAD_GAUSS_LINEAR && ctx->faceBatch)
            LAUNCH(k_gauss_grad_batched, gridFor(n), kThreads, n, f, ctx->cf.sliceOff, ctx->cf.col, ctx->owner, ctx->neighbour, G, dx,
                   (const double *)ctx->xbf, dxn, (const u8 *)ctx->bCoupled, ctx->grad);
        else if (scheme == AMR_GRAD_GAUSS_LINEAR)
            LAUNCH(k_gauss_grad, gridFor(n), kThreads, n, f, ctx->cf.sliceOff, ctx->cf.col, ctx->owner, ctx->neighbour, G, dx,
                   (const double *)ctx->xbf, dxn, (const u8 *)ctx->bCoupled, ctx->grad);
        else
            LAUNCH(k_ls_grad, gridFor(n), kThreads, n, f, ctx->cf.sliceOff, ctx->cf.col, (const label *)ctx->owner, (const label *)ctx->neighbour,
                   (const double *)ctx->lsP, (const double *)ctx->lsN, dx, (const double *)ctx->xbf, ctx->grad);
        LAUNCH(k_mag_grad, gridFor(n), kThreads, n, (const double *)ctx->grad, (const double *)ctx->gV, totalVolume, target);
    }
    TRY(normalizeRange(ctx, target, refineThreshold, unrefineThreshold, scale, thrOut, nullptr));
    if (gradOut && n) { CK(cudaMemcpyAsync(gradOut, ctx->grad, (size_t)n * 24, cudaMemcpyDefault, ctx->stream)); if (!isDevicePtr(gradOut)) ctx->hostTouched = true; }
    if (errorOut && !devOut && n) { CK(cudaMemcpyAsync(errorOut, target, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream)); ctx->hostTouched = true; }
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_mesh_size(amr_ctx *ctx, const int32_t *geometricD, double *dx, double *dX) {
    ENTER();
    if (!ctx->haveGeom) FAIL(AMR_ERR_STATE, "amr_mesh_size: geometry not set (amr_mesh_set_geometry)");
    if (!geometricD) FAIL(AMR_ERR_ARG, "amr_mesh_size: geometricD is null");
    const int64_t n = ctx->n;
    double *dDx = nullptr, *dDX = nullptr;
    TRY(stageOut(ctx, dx, n, &dDx));
    TRY(stageOut(ctx, dX, 3 * n, &dDX));
    int nGeo = 0;
    for (int k = 0; k < 3; k++) if (geometricD[k] > 0) nGeo++;
    if (dDx && n) {
        if (nGeo == 3) LAUNCH(k_mesh_dx_cbrt, gridFor(n), kThreads, n, (const double *)ctx->gV, dDx);
        else LAUNCH(k_mesh_dx_planar, gridFor(n), kThreads, n, ctx->f, ctx->cf.sliceOff, ctx->cf.col, (const double *)ctx->gSf,
                    geometricD[0] < 0 ? 1.0 : 0.0, geometricD[1] < 0 ? 1.0 : 0.0, geometricD[2] < 0 ? 1.0 : 0.0, dDx);
    }
    if (dDX && n) LAUNCH(k_mesh_dX, gridFor(n), kThreads, n, ctx->cf.sliceOff, ctx->cf.col, (const double *)ctx->gSf, (const double *)ctx->gV, dDX);
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_max_refinement(amr_ctx *ctx, int32_t maxLevel, double minDx, const double *dx, const double *error,
                               int32_t *maxRefinementOut) {
    ENTER();
    if (!maxRefinementOut) FAIL(AMR_ERR_ARG, "amr_max_refinement: output is null");
    const int64_t n = ctx->n;
    int32_t *dOut = nullptr;
    TRY(stageOut(ctx, maxRefinementOut, n, &dOut));
    if (maxLevel >= 0 || !ctx->haveLevels) {                       // errorEstimator.C:264-267
        if (n) LAUNCH(k_fill_labels, gridFor(n), kThreads, n, (label)maxLevel, dOut);
    } else {
        if (!dx) FAIL(AMR_ERR_ARG, "amr_max_refinement: dx is null (amr_mesh_size)");
        const double *dDx = nullptr, *dErr = nullptr;
        TRY(stageIn(ctx, dx, n, &dDx));
        TRY(stageIn(ctx, error, n, &dErr));
        if (!dErr) dErr = ctx->error;
        if (!dErr) FAIL(AMR_ERR_STATE, "amr_max_refinement: no error field");
        if (n) LAUNCH(k_max_refinement, gridFor(n), kThreads, n, (const label *)ctx->lvl, minDx, dDx, dErr, dOut);
    }
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_estimate_cell_size(amr_ctx *ctx, int sizeType, const int32_t *geometricD, double lowerUnrefine, double lowerRefine,
                                   const int32_t *cmpts, int nCmpts, const double *minDX, const double *maxDX, const double *dx,
                                   const double *dX, double *errorInOut) {
    ENTER();
    if (!ctx->haveGeom) FAIL(AMR_ERR_STATE, "amr_estimate_cell_size: geometry not set (amr_mesh_set_geometry)");
    if (sizeType < AMR_SIZE_VOLUME || sizeType > AMR_SIZE_CMPT) FAIL(AMR_ERR_ARG, "amr_estimate_cell_size: unknown size type");
    if (!geometricD) FAIL(AMR_ERR_ARG, "amr_estimate_cell_size: geometricD is null");
    const int64_t n = ctx->n;
    CellSizeArgs A;
    A.sizeType = sizeType;
    for (int k = 0; k < 3; k++) { A.geoValid[k] = geometricD[k] == 1; A.cmpts[k] = 0; A.minDX[k] = 0; A.maxDX[k] = 0; }
    A.lowerUnrefine = lowerUnrefine; A.lowerRefine = lowerRefine;
    A.nCmpts = 0;
    if (sizeType == AMR_SIZE_CMPT) {
        if (nCmpts < 0 || nCmpts > 3 || (nCmpts && (!cmpts || !minDX || !maxDX))) FAIL(AMR_ERR_ARG, "amr_estimate_cell_size: bad component list");
        A.nCmpts = nCmpts;
        for (int i = 0; i < nCmpts; i++) {
            if (cmpts[i] < 0 || cmpts[i] > 2) FAIL(AMR_ERR_ARG, "amr_estimate_cell_size: component out of range");
            A.cmpts[i] = cmpts[i];
        }
        for (int k = 0; k < 3; k++) { A.minDX[k] = minDX[k]; A.maxDX[k] = maxDX[k]; }
    }
    const double *dDx = nullptr, *dDX = nullptr;
    TRY(stageIn(ctx, dx, n, &dDx));
    TRY(stageIn(ctx, dX, 3 * n, &dDX));
    if (sizeType == AMR_SIZE_CHARACTERISTIC && !dDx) FAIL(AMR_ERR_ARG, "amr_estimate_cell_size: dx is null (amr_mesh_size)");
    if ((sizeType == AMR_SIZE_MAG || sizeType == AMR_SIZE_CMPT) && !dDX) FAIL(AMR_ERR_ARG, "amr_estimate_cell_size: dX is null (amr_mesh_size)");
    double *dE = nullptr;
    TRY(stageInOut(ctx, errorInOut, n, &dE));
    double *e = dE ? dE : ctx->error;
    if (n) LAUNCH(k_cell_size, gridFor(n), kThreads, n, A, (const double *)ctx->gV, dDx, dDX, e);
    if (dE && n && dE != ctx->error) CK(cudaMemcpyAsync(ctx->error, dE, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->errorResident = true;
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_multicomponent(amr_ctx *ctx, int nEstimators, const double *const *errors, const int32_t *const *maxCellLevels,
                               double *errorOut, int32_t *maxRefinementOut, int *wavesOut) {
    ENTER();
    if (!ctx->meshSet) FAIL(AMR_ERR_STATE, "amr_multicomponent: mesh not set");
    if (nEstimators < 0 || (nEstimators && !errors)) FAIL(AMR_ERR_ARG, "amr_multicomponent: errors is null");
    const int64_t n = ctx->n;
    std::vector<const double *> dErr((size_t)nEstimators, nullptr);
    for (int i = 0; i < nEstimators; i++) {
        if (!errors[i]) FAIL(AMR_ERR_ARG, "amr_multicomponent: errors[i] is null");
        TRY(stageIn(ctx, errors[i], n, &dErr[(size_t)i]));
    }
    // update(): error = max over the sub-estimators, starting from -1 (multicomponent.C:100-118)
    const bool devOut = errorOut && isDevicePtr(errorOut);
    double *target = devOut ? errorOut : ctx->error;
    ctx->errorResident = !devOut;
    if (n) {
        if (nEstimators == 0) LAUNCH(k_fill_scalar, gridFor(n), kThreads, n, -1.0, target);
        for (int i = 0; i < nEstimators; i++) LAUNCH(k_max_combine, gridFor(n), kThreads, n, dErr[(size_t)i], target, i == 0 ? 1 : 0);
        if (errorOut && !devOut) { CK(cudaMemcpyAsync(errorOut, target, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream)); ctx->hostTouched = true; }
    }
    int waves = 0;
    if (maxRefinementOut) {
        if (nEstimators && !maxCellLevels) FAIL(AMR_ERR_ARG, "amr_multicomponent: maxCellLevels is null");
        int32_t *dOut = nullptr;
        TRY(stageOut(ctx, maxRefinementOut, n, &dOut));
        if (n) CK(cudaMemsetAsync(dOut, 0, (size_t)n * 4, ctx->stream));
        u8 *state = ctx->flagA;
        for (int i = 0; i < nEstimators && n; i++) {
            const int32_t *dLvl = nullptr;
            if (!maxCellLevels[i]) FAIL(AMR_ERR_ARG, "amr_multicomponent: maxCellLevels[i] is null");
            TRY(stageIn(ctx, maxCellLevels[i], n, &dLvl));
            LAUNCH(k_mc_init, gridFor(n), kThreads, n, dErr[(size_t)i], state);
            for (;;) {
                CK(cudaMemsetAsync(ctx->dFlags + 10, 0, 2 * sizeof(int), ctx->stream));
                LAUNCH(k_mc_wave, gridFor(n), kThreads, n, ctx->cc.sliceOff, ctx->cc.col, dLvl, (const label *)dOut, state, ctx->dFlags + 10);
                CK(cudaMemcpyAsync(ctx->hFlags + 10, ctx->dFlags + 10, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
                waves++;
                if (!ctx->hFlags[11]) break;                      // nothing left undetermined
                if (!ctx->hFlags[10]) FAIL(AMR_ERR_INVARIANT, "amr_multicomponent: wavefront made no progress");
            }
            LAUNCH(k_mc_apply, gridFor(n), kThreads, n, ctx->cc.sliceOff, ctx->cc.col, dLvl, (const u8 *)state, dOut);
        }
    }
    if (wavesOut) *wavesOut = waves;
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

// ---- inputs of the load-balance decision

AMR_API int amr_imbalance(amr_ctx *ctx, double *maxImbalanceRatio, int64_t *nGlobalCells) {
    ENTER();
    if (!maxImbalanceRatio) FAIL(AMR_ERR_ARG, "amr_imbalance: output is null");
    const int nProcs = (ctx->comm && ctx->nranks > 1) ? ctx->nranks : 1;
    int64_t nGlobal = ctx->n;
    if (nProcs > 1) {
        CK(cudaMemcpyAsync(ctx->dCount + 1, &nGlobal, 8, cudaMemcpyHostToDevice, ctx->stream));
        TRY(globalSum(ctx, 1, &nGlobal));                     // returnReduce(mesh_.nCells(), sumOp<label>())
    }
    const double ideal = (double)nGlobal / (double)nProcs;    // fvMeshBalance.C:463-464
    double localImbalance = fabs((double)ctx->n - ideal);     // :465
    if (nProcs > 1) {
        void *q;
        TRY(arenaAlloc(ctx, &q, 8));
        CK(cudaMemcpyAsync(q, &localImbalance, 8, cudaMemcpyHostToDevice, ctx->stream));
        TRY(allReduceWord(ctx, q, 3));                        // returnReduce(localImbalance, maxOp<scalar>())
        CK(cudaMemcpyAsync(&localImbalance, q, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    *maxImbalanceRatio = localImbalance / ideal;              // :467
    if (nGlobalCells) *nGlobalCells = nGlobal;
    return AMR_OK;
}

// cellToCluster on the device (arena): hexRefRefinementHistory::markCommonCells
static int historyClusters(amr_ctx *ctx, label **cellToClusterDev, int32_t **nClustersDev) {
    const int64_t n = ctx->n, ns = ctx->nSplit;
    void *q;
    TRY(arenaAlloc(ctx, &q, (size_t)(n > 0 ? n : 1) * 4)); label *root = (label *)q;
    TRY(arenaAlloc(ctx, &q, (size_t)(ns > 0 ? ns : 1) * 4)); int *firstCell = (int *)q;
    TRY(arenaAlloc(ctx, &q, (size_t)(n > 0 ? n : 1) * 4)); label *c2c = (label *)q;
    TRY(arenaAlloc(ctx, &q, (size_t)(n > 0 ? n : 1) * 4)); int32_t *firstList = (int32_t *)q;
    if (ns) LAUNCH(k_fill_int, gridFor(ns), kThreads, ns, 0x7fffffff, firstCell);
    u8 *isFirst = ctx->flagA;
    int32_t *total = nullptr;
    if (n) {
        LAUNCH(k_history_root, gridFor(n), kThreads, n, (const label *)ctx->visible, (const label *)ctx->splitParent, root, firstCell);
        LAUNCH(k_cluster_first_flag, gridFor(n), kThreads, n, (const label *)root, (const int *)firstCell, isFirst);
    }
    TRY(compactFlags(ctx, isFirst, n, firstList, &total));
    if (n) {
        LAUNCH(k_cluster_of_root, gridFor(n), kThreads, (const int32_t *)total, (const int32_t *)firstList, (const label *)root, firstCell);
        LAUNCH(k_cluster_cells, gridFor(n), kThreads, n, (const label *)root, (const int *)firstCell, c2c);
    }
    *cellToClusterDev = c2c;
    *nClustersDev = total;
    return AMR_OK;
}

AMR_API int amr_history_clusters(amr_ctx *ctx, const int32_t *cellToClusterIn, int32_t *cellToClusterOut, uint8_t *blockedFace,
                                 int64_t *nClusters, int64_t *nUnblocked) {
    ENTER();
    if (!ctx->meshSet) FAIL(AMR_ERR_STATE, "amr_history_clusters: mesh not set");
    if (!ctx->faceLists) FAIL(AMR_ERR_STATE, "amr_history_clusters: needs owner/neighbour (amr_mesh_update hands over the cell adjacency only; face lists come with amr_mesh_set)");
    const int64_t n = ctx->n, f = ctx->f, nf = ctx->f + ctx->b;
    label *c2c = nullptr;
    int32_t *total = nullptr;
    if (cellToClusterIn) {                                    // polyRefiner: refinement::getCellClusters ids from the host (any numbering)
        const int32_t *d = nullptr;
        TRY(stageIn(ctx, cellToClusterIn, n, &d));
        c2c = (label *)d;
    } else {
        if (!ctx->haveHistory) FAIL(AMR_ERR_STATE, "amr_history_clusters: refinement history not set (amr_levels_set)");
        TRY(historyClusters(ctx, &c2c, &total));
    }
    if (cellToClusterOut && n) {
        CK(cudaMemcpyAsync(cellToClusterOut, c2c, (size_t)n * 4, cudaMemcpyDefault, ctx->stream));
        if (!isDevicePtr(cellToClusterOut)) ctx->hostTouched = true;
    }
    CK(cudaMemsetAsync(ctx->dCount + 3, 0, 8, ctx->stream));
    if (blockedFace) {
        u8 *dB = nullptr;
        TRY(stageOut(ctx, blockedFace, nf, &dB));
        if (nf) LAUNCH(k_blocked_faces, gridFor(nf), kThreads, nf, f, (const label *)ctx->owner, (const label *)ctx->neighbour, (const label *)c2c, dB,
                       (unsigned long long *)(ctx->dCount + 3));
    }
    if (nClusters || nUnblocked) {
        ctx->hFlags[10] = 0;
        if (total) CK(cudaMemcpyAsync(ctx->hFlags + 10, total, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->hCount + 3, ctx->dCount + 3, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (nClusters) *nClusters = total ? ctx->hFlags[10] : -1;
        if (nUnblocked) *nUnblocked = ctx->hCount[3];
    }
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_history_constrain_decomposition(amr_ctx *ctx, int32_t *decomposition, int64_t *nChanged) {
    ENTER();
    if (!ctx->meshSet) FAIL(AMR_ERR_STATE, "amr_history_constrain_decomposition: mesh not set");
    if (!ctx->faceLists) FAIL(AMR_ERR_STATE, "amr_history_constrain_decomposition: needs owner/neighbour (amr_mesh_update hands over the cell adjacency only; face lists come with amr_mesh_set)");
    if (!ctx->haveHistory) FAIL(AMR_ERR_STATE, "amr_history_constrain_decomposition: refinement history not set (amr_levels_set)");
    if (!decomposition) FAIL(AMR_ERR_ARG, "amr_history_constrain_decomposition: decomposition is null");
    const int64_t n = ctx->n, f = ctx->f;
    label *c2c = nullptr;
    int32_t *total = nullptr;
    TRY(historyClusters(ctx, &c2c, &total));
    int32_t *dDec = nullptr;
    TRY(stageInOut(ctx, decomposition, n, &dDec));
    void *q;
    TRY(arenaAlloc(ctx, &q, (size_t)(n > 0 ? n : 1) * 4)); label *decIn = (label *)q;       // the values before the loop
    TRY(arenaAlloc(ctx, &q, (size_t)(n > 0 ? n : 1) * 4)); int *firstFace = (int *)q;       // clusters <= cells
    TRY(arenaAlloc(ctx, &q, (size_t)(n > 0 ? n : 1) * 4)); int *clusterToProc = (int *)q;
    u8 *changed = ctx->flagB;
    if (n) {
        CK(cudaMemcpyAsync(decIn, dDec, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemsetAsync(changed, 0, (size_t)n, ctx->stream));
        LAUNCH(k_fill_int, gridFor(n), kThreads, n, 0x7fffffff, firstFace);
    }
    if (f) LAUNCH(k_cluster_first_face, gridFor(f), kThreads, f, (const label *)ctx->owner, (const label *)ctx->neighbour, (const label *)c2c, firstFace);
    if (n) LAUNCH(k_cluster_proc, gridFor(n), kThreads, n, (const int *)firstFace, (const label *)ctx->owner, (const label *)decIn, clusterToProc);
    if (f) LAUNCH(k_cluster_apply, gridFor(f), kThreads, f, (const label *)ctx->owner, (const label *)ctx->neighbour, (const label *)c2c,
                  (const int *)clusterToProc, (const label *)decIn, dDec, changed);
    if (nChanged) {
        CK(cudaMemsetAsync(ctx->dCount + 3, 0, 8, ctx->stream));
        if (n) LAUNCH(k_count_flags, gridFor(n), kThreads, n, (const u8 *)changed, (unsigned long long *)(ctx->dCount + 3));
        CK(cudaMemcpyAsync(ctx->hCount + 3, ctx->dCount + 3, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        *nChanged = ctx->hCount[3];
    }
    CK(cudaGetLastError());
    return flushOuts(ctx);
}

AMR_API int amr_proc_load(amr_ctx *ctx, const int32_t *distribution, int nProcs, int64_t *procLoad) {
    ENTER();
    if (!distribution || !procLoad || nProcs < 1 || nProcs > 8192) FAIL(AMR_ERR_ARG, "amr_proc_load: bad arguments");
    const int64_t n = ctx->n;
    const int32_t *dD = nullptr;
    TRY(stageIn(ctx, distribution, n, &dD));
    void *q;
    TRY(arenaAlloc(ctx, &q, (size_t)nProcs * 8));
    unsigned long long *load = (unsigned long long *)q;
    CK(cudaMemsetAsync(load, 0, (size_t)nProcs * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->dFlags + 10, 0, sizeof(int), ctx->stream));
    if (n) {
        const unsigned g = (unsigned)std::min<int64_t>((int64_t)ctx->numSMs * 8, (n + kThreads - 1) / kThreads);
        k_proc_load<<<g, kThreads, (size_t)nProcs * 4, ctx->stream>>>(n, dD, nProcs, load, ctx->dFlags + 10);
        ctx->launches++;
    }
    if (ctx->comm && ctx->nranks > 1)
        for (int k = 0; k < nProcs; k++) TRY(allReduceWord(ctx, load + k, 1));     // reduce(procLoadNew, sumOp<labelList>())
    CK(cudaMemcpyAsync(procLoad, load, (size_t)nProcs * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->hFlags + 10, ctx->dFlags + 10, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->hFlags[10]) FAIL(AMR_ERR_ARG, "amr_proc_load: distribution holds a processor outside [0, nProcs)");
    return AMR_OK;
}
