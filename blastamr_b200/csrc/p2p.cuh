// p2p.cuh -- processor-patch halo swaps and small reductions over NVLink peer memory.
//
// NCCL's grouped send/recv walks a schedule over ALL ranks of the communicator, so a 160 KB halo swap
// with 3 neighbours costs ~0.2 ms on 8 ranks (measured: 0.64 ms/step of pure latency).  Here every rank
// owns one IPC-exported region (two parity buffers + flag words); a swap is
//   k_halo_push_cells : gather my boundary-cell values and store them straight into the neighbours' regions
//                       (NVLink stores); the last CTA release-stores the epoch into their flag words
//   k_halo_unpack     : acquire-spin on my own flag words, then copy the received ranges into `recv`
// and a reduction (the reduce()/gMin/gMax of the reference) is the same pattern to all ranks in ONE launch.
//
// Exchange numbers live in DEVICE memory (seq[0] halo swaps, seq[1] reductions): a kernel reads the number of the last
// completed exchange, uses the next one, and the receiving kernel stores it back when it is done.  The host never
// passes an epoch, so a sweep loop can be enqueued speculatively, every launch gated by a device word that all ranks
// hold identically (the all-reduced size of the work queues): a gated-off exchange consumes no number on any rank and
// the parity buffers stay in step.  At most ONE halo swap is in flight per context (push ... unpack).
// Double buffering by parity makes reuse safe: a rank can be at most one exchange ahead of a neighbour (it needs that
// neighbour's signal to finish its own wait).  Spins are bounded (wall-clock limit from the context:
// AMR_B200_PEER_TIMEOUT_S, default 20 s -- long enough for rank-imbalanced host work between two collective calls)
// and raise the region's error word instead of hanging the GPU; the host reads that word at every synchronising
// return (flushOuts / the sweep round trips) and fails the call.  NCCL stays as set-up transport (handle exchange)
// and as fallback (AMR_B200_HALO=nccl, or when IPC mapping is not possible).
#pragma once
#include "common.cuh"

constexpr int kMaxRanks = 64;
constexpr int64_t kHaloCap = (int64_t)64 << 20;        // bytes per parity buffer

__device__ __forceinline__ unsigned long long globalTimerNs() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct P2PRegionHeader {                                 // lives at the start of every rank's region
    unsigned arrive[2][kMaxRanks];                       // halo epoch written by each neighbour
    unsigned redEpoch[2][kMaxRanks];
    unsigned long long red[2][kMaxRanks];
    int error;
    int pad[63];
};
constexpr int64_t kRegionBytes = (int64_t)sizeof(P2PRegionHeader) + 2 * kHaloCap;

struct HaloPlan {                                        // device-side description of one rank's coupled patches
    int nPatch;
    int nbr[32];
    int offMine[32], offTheirs[32], size[32];
    int prefix[33];                                      // prefix sum of sizes
};

__device__ __forceinline__ void stReleaseSys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ldAcquireSys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// fused pack + push + signal: value of a per-cell array at the owner of every coupled boundary face goes
// straight into the neighbour's region; the last CTA to finish publishes the epoch (threadfence-reduction
// pattern), so a swap is two launches (this one and the receiving kernel).  `minus` (optional, bytes): v -= minus[o].
// gate (optional): the exchange only happens when *gate != 0 -- the same value on every rank.
// Byte payloads (marks, level +/- flag: every swap of the sweep loops) travel as 32-bit words: a thread owns 4
// consecutive destination bytes, aligned in the neighbour's buffer, so NVLink carries one store per 4 faces.
// No per-thread system fence: the CTA barrier orders every thread's stores before thread 0's fence (cumulativity),
// and the last CTA's release store of the epoch is what the receiver acquires.
template <int BYTES>
__global__ void k_halo_push_cells(HaloPlan P, char *const *__restrict__ peer, int myRank, const unsigned *__restrict__ seq,
                                  const label *__restrict__ bOwner, const char *__restrict__ cells, const u8 *__restrict__ minus,
                                  unsigned *counter, const int *__restrict__ gate) {
    if (gate && *gate == 0) return;
    const unsigned epoch = seq[0] + 1u;                  // stored back by the receiving kernel, later in stream order
    const int parity = (int)(epoch & 1u);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (BYTES == 1) {
        // thread i <-> the i-th aligned destination word over all patches (prefixW = prefix of the word counts)
        int q = 0, w = i;
        bool live = false;
        for (; q < P.nPatch; q++) {
            const int a = P.offTheirs[q] & 3;
            const int words = (P.size[q] + a + 3) >> 2;
            if (w < words) { live = true; break; }
            w -= words;
        }
        if (live) {
            const int a = P.offTheirs[q] & 3;
            const int k0 = 4 * w - a;                    // first element of this word (may be < 0 / reach past the patch)
            unsigned pack = 0;
            bool full = k0 >= 0 && k0 + 4 <= P.size[q];
            u8 vals[4] = {0, 0, 0, 0};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int k = k0 + t;
                if (k >= 0 && k < P.size[q]) {
                    const label o = bOwner[P.offMine[q] + k];
                    u8 v = (u8)cells[o];
                    if (minus) v = (u8)(v - minus[o]);
                    vals[t] = v;
                }
                pack |= (unsigned)vals[t] << (8 * t);
            }
            char *dst = peer[P.nbr[q]] + sizeof(P2PRegionHeader) + (size_t)parity * kHaloCap + (size_t)(P.offTheirs[q] + k0);
            if (full) *reinterpret_cast<unsigned *>(dst) = pack;
            else {
#pragma unroll
                for (int t = 0; t < 4; t++) { const int k = k0 + t; if (k >= 0 && k < P.size[q]) dst[t] = (char)vals[t]; }
            }
        }
    } else if (i < P.prefix[P.nPatch]) {
        int q = 0;
        while (i >= P.prefix[q + 1]) q++;
        const int k = i - P.prefix[q];
        const label o = bOwner[P.offMine[q] + k];
        char *dst = peer[P.nbr[q]] + sizeof(P2PRegionHeader) + (size_t)parity * kHaloCap + (size_t)(P.offTheirs[q] + k) * BYTES;
        if (BYTES == 4) *reinterpret_cast<int *>(dst) = reinterpret_cast<const int *>(cells)[o];
        else if (BYTES == 8) *reinterpret_cast<double *>(dst) = reinterpret_cast<const double *>(cells)[o];
        else {                                               // vector field: 3 doubles per cell
            const double *s3 = reinterpret_cast<const double *>(cells) + 3 * (size_t)o;
            double *d3 = reinterpret_cast<double *>(dst);
            d3[0] = s3[0]; d3[1] = s3[1]; d3[2] = s3[2];
        }
    }
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                              // this CTA's stores (ordered before by the barrier) are out
        last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        if (threadIdx.x < P.nPatch) {
            __threadfence_system();
            P2PRegionHeader *h = reinterpret_cast<P2PRegionHeader *>(peer[P.nbr[threadIdx.x]]);
            stReleaseSys(&h->arrive[parity][myRank], epoch);
        }
        if (threadIdx.x == 0) *counter = 0;
    }
}
// receive side; the last CTA stores the exchange number back (seq[0] = epoch)
template <int BYTES>
__global__ void k_halo_unpack(HaloPlan P, char *mine, unsigned *seq, char *__restrict__ recv, unsigned long long timeoutNs,
                              unsigned *counter, const int *__restrict__ gate) {
    if (gate && *gate == 0) return;
    __shared__ int ok;
    const unsigned epoch = seq[0] + 1u;
    const int parity = (int)(epoch & 1u);
    P2PRegionHeader *h = reinterpret_cast<P2PRegionHeader *>(mine);
    if (threadIdx.x == 0) {
        int good = h->error == 0;                        // an earlier timeout of this call chain: do not touch recv
        const unsigned long long t0 = globalTimerNs();
        for (int q = 0; q < P.nPatch && good; q++) {
            unsigned spins = 0;
            while ((int)(ldAcquireSys(&h->arrive[parity][P.nbr[q]]) - epoch) < 0) {
                if ((++spins & 1023u) == 0 && globalTimerNs() - t0 > timeoutNs) { good = 0; h->error = 1; break; }
            }
        }
        ok = good;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (ok && i < P.prefix[P.nPatch]) {
        int q = 0;
        while (i >= P.prefix[q + 1]) q++;
        const int k = i - P.prefix[q];
        const char *src = mine + sizeof(P2PRegionHeader) + (size_t)parity * kHaloCap + (size_t)(P.offMine[q] + k) * BYTES;
        char *dst = recv + (size_t)(P.offMine[q] + k) * BYTES;
        if (BYTES == 1) *dst = *src;
        else if (BYTES == 4) *reinterpret_cast<int *>(dst) = *reinterpret_cast<const int *>(src);
        else if (BYTES == 8) *reinterpret_cast<double *>(dst) = *reinterpret_cast<const double *>(src);
        else {
            const double *s3 = reinterpret_cast<const double *>(src);
            double *d3 = reinterpret_cast<double *>(dst);
            d3[0] = s3[0]; d3[1] = s3[1]; d3[2] = s3[2];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(counter, 1u) == gridDim.x - 1) { *counter = 0; seq[0] = epoch; }
    }
}

// ---- all-rank reduction of one 8-byte word in ONE launch (one CTA, thread r <-> rank r): push my word to every rank,
// wait for everybody's, combine, store the result in place.  op: 0 max(int32), 1 sum(int64), 2 min(f64), 3 max(f64),
// 4 sum(int32).  The value sent is *src (default: *word), the result goes to *word.
__global__ void k_red_all(char *const *__restrict__ peer, char *mine, int nranks, int myRank, unsigned *seq, int op,
                          const void *src, void *word, unsigned long long timeoutNs, const int *__restrict__ gate) {
    if (gate && *gate == 0) return;                      // a gated-off reduction leaves *word as it is (0: cleared by the caller)
    if (!src) src = word;
    __shared__ unsigned long long vals[kMaxRanks];
    __shared__ int bad;
    const unsigned epoch = seq[1] + 1u;
    const int parity = (int)(epoch & 1u);
    const bool is32 = (op == 0 || op == 4);
    const int r = threadIdx.x;
    if (r == 0) bad = 0;
    __syncthreads();
    if (r < nranks) {
        P2PRegionHeader *h = reinterpret_cast<P2PRegionHeader *>(peer[r]);
        h->red[parity][myRank] = is32 ? (unsigned long long)(unsigned)(*reinterpret_cast<const int *>(src))
                                      : *reinterpret_cast<const unsigned long long *>(src);
        __threadfence_system();
        stReleaseSys(&h->redEpoch[parity][myRank], epoch);
        P2PRegionHeader *me = reinterpret_cast<P2PRegionHeader *>(mine);
        const unsigned long long t0 = globalTimerNs();
        unsigned spins = 0;
        while ((int)(ldAcquireSys(&me->redEpoch[parity][r]) - epoch) < 0) {
            if ((++spins & 1023u) == 0 && globalTimerNs() - t0 > timeoutNs) { bad = 1; me->error = 1; break; }
        }
        vals[r] = *reinterpret_cast<volatile unsigned long long *>(&me->red[parity][r]);
    }
    __syncthreads();
    if (r == 0) {
        if (!bad) {
            unsigned long long acc = vals[0];
            for (int k = 1; k < nranks; k++) {
                const unsigned long long v = vals[k];
                if (op == 0) { int a = (int)acc, b2 = (int)v; acc = (unsigned long long)(unsigned)(a > b2 ? a : b2); }
                else if (op == 4) acc = (unsigned long long)(unsigned)((int)acc + (int)v);
                else if (op == 1) acc = (unsigned long long)((long long)acc + (long long)v);
                else {
                    const double a = __longlong_as_double((long long)acc), b2 = __longlong_as_double((long long)v);
                    const double m = (op == 2) ? ((a < b2) ? a : b2) : ((a > b2) ? a : b2);
                    acc = (unsigned long long)__double_as_longlong(m);
                }
            }
            if (is32) *reinterpret_cast<int *>(word) = (int)acc;
            else *reinterpret_cast<unsigned long long *>(word) = acc;
        } else if (is32) *reinterpret_cast<int *>(word) = 0;     // a timed-out exchange ends a speculative loop; the host reports it
        seq[1] = epoch;
    }
}

// ---- boundary fix-up kernels fused with the receive side of a swap (and with the reduction that follows it) ----------
// The neighbour stores the value of my boundary face bf at offset bf of my parity buffer -- the parity buffer IS the
// ghost array in boundary-face layout -- so a fix-up kernel needs no unpack launch: it waits for the arrival words
// itself (thread 0 of every CTA), reads the ghost values in place, and its last CTA stores the exchange number back
// and, where the caller wants it, runs the one-word all-reduce of the sweep (queue size) right there.
struct GhostLink {
    const char *ghost;            // transport already delivered the values (NCCL / host-provided): boundary-face layout; else nullptr
    HaloPlan P;                   // peer transport: the patches to wait for
    char *region;                 // my region
    char *const *peer;            // all regions (fused reduction), nullptr = no fused reduction
    unsigned *seq;
    unsigned *counter;
    unsigned long long timeoutNs;
    int nranks, rank;
};
// returns the ghost array, or nullptr after a timeout (the caller then skips its work but still calls ghostRelease)
__device__ __forceinline__ const char *ghostAcquire(const GhostLink &G) {
    if (G.ghost) return G.ghost;
    __shared__ int ok;
    const unsigned epoch = G.seq[0] + 1u;
    const int parity = (int)(epoch & 1u);
    if (threadIdx.x == 0) {
        P2PRegionHeader *h = reinterpret_cast<P2PRegionHeader *>(G.region);
        int good = h->error == 0;
        const unsigned long long t0 = globalTimerNs();
        for (int q = 0; q < G.P.nPatch && good; q++) {
            unsigned spins = 0;
            while ((int)(ldAcquireSys(&h->arrive[parity][G.P.nbr[q]]) - epoch) < 0) {
                if ((++spins & 1023u) == 0 && globalTimerNs() - t0 > G.timeoutNs) { good = 0; h->error = 1; break; }
            }
        }
        ok = good;
    }
    __syncthreads();
    return ok ? G.region + sizeof(P2PRegionHeader) + (size_t)parity * kHaloCap : nullptr;
}
// end of a fused fix-up kernel: true in the last CTA to finish (all CTAs' writes are visible to it)
__device__ __forceinline__ bool ghostRelease(const GhostLink &G) {
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = (atomicAdd(G.counter, 1u) == gridDim.x - 1);
        if (last) { *G.counter = 0; if (!G.ghost) G.seq[0] = G.seq[0] + 1u; }
    }
    __syncthreads();
    return last != 0;
}
// all-reduce (sum, int32) of *src into *dst by ONE CTA of at least nranks threads (the last CTA of a fix-up kernel)
__device__ __forceinline__ void ctaAllReduceSum(const GhostLink &G, const int *src, int *dst) {
    __shared__ int vals[kMaxRanks];
    __shared__ int bad;
    const unsigned epoch = G.seq[1] + 1u;
    const int parity = (int)(epoch & 1u);
    const int r = threadIdx.x;
    if (r == 0) bad = 0;
    __syncthreads();
    if (r < G.nranks) {
        P2PRegionHeader *h = reinterpret_cast<P2PRegionHeader *>(G.peer[r]);
        h->red[parity][G.rank] = (unsigned long long)(unsigned)(*reinterpret_cast<const volatile int *>(src));
        __threadfence_system();
        stReleaseSys(&h->redEpoch[parity][G.rank], epoch);
        P2PRegionHeader *me = reinterpret_cast<P2PRegionHeader *>(G.region);
        const unsigned long long t0 = globalTimerNs();
        unsigned spins = 0;
        while ((int)(ldAcquireSys(&me->redEpoch[parity][r]) - epoch) < 0) {
            if ((++spins & 1023u) == 0 && globalTimerNs() - t0 > G.timeoutNs) { bad = 1; me->error = 1; break; }
        }
        vals[r] = (int)(unsigned)*reinterpret_cast<volatile unsigned long long *>(&me->red[parity][r]);
    }
    __syncthreads();
    if (r == 0) {
        int acc = 0;
        if (!bad) for (int k = 0; k < G.nranks; k++) acc += vals[k];
        *dst = acc;                                       // a timed-out exchange ends a speculative loop; the host reports it
        G.seq[1] = epoch;
    }
}
