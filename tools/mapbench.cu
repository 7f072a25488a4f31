// mapbench.cu -- stand-alone timing of field-map kernel variants (new[i] = old[cellMap[i]], NC interleaved
// components) on the bench's shape: 64 M old cells, 2.05 M refined parents, 7 appended children each.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_mapbench tools/mapbench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int kThreads = 256;

template <int NC, int ITEMS>
__global__ void __launch_bounds__(kThreads) v_elem(int64_t nNew, const int *__restrict__ cellMap, const double *__restrict__ oldF, double *__restrict__ newF) {
    const int64_t total = nNew * NC;
    const int64_t e0 = (int64_t)blockIdx.x * (kThreads * ITEMS) + threadIdx.x;
    int64_t src[ITEMS];
#pragma unroll
    for (int t = 0; t < ITEMS; t++) {
        const int64_t e = e0 + (int64_t)t * kThreads;
        if (e < total) { const int64_t i = e / NC; src[t] = (int64_t)cellMap[i] * NC + (e - i * NC); } else src[t] = -1;
    }
    double v[ITEMS];
#pragma unroll
    for (int t = 0; t < ITEMS; t++) if (src[t] >= 0) v[t] = oldF[src[t]];
#pragma unroll
    for (int t = 0; t < ITEMS; t++) if (src[t] >= 0) newF[e0 + (int64_t)t * kThreads] = v[t];
}

// pair variant: a thread owns ITEMS pairs of consecutive elements (e even); one 128-bit load when the two sources are
// consecutive and aligned (identity runs, same-cell components), always one 128-bit store.  HINT: 0 default, 1 streaming
template <int NC, int ITEMS, int HINT>
__global__ void __launch_bounds__(kThreads) v_pair(int64_t nNew, const int *__restrict__ cellMap, const double *__restrict__ oldF, double *__restrict__ newF) {
    const int64_t total = nNew * NC;
    const int64_t p0 = (int64_t)blockIdx.x * (kThreads * ITEMS) + threadIdx.x;
    int64_t s0[ITEMS], s1[ITEMS];
#pragma unroll
    for (int t = 0; t < ITEMS; t++) {
        const int64_t e = 2 * (p0 + (int64_t)t * kThreads);
        s0[t] = s1[t] = -1;
        if (e < total) {
            const int64_t i = e / NC; const int k = (int)(e - i * NC);
            const int m0 = cellMap[i];
            s0[t] = (int64_t)m0 * NC + k;
            if (e + 1 < total) {
                if (k + 1 < NC) s1[t] = s0[t] + 1;
                else s1[t] = (int64_t)cellMap[i + 1] * NC;
            }
        }
    }
    double2 v[ITEMS];
#pragma unroll
    for (int t = 0; t < ITEMS; t++) {
        if (s0[t] < 0) continue;
        if (s1[t] == s0[t] + 1 && !(s0[t] & 1)) {
            v[t] = HINT ? __ldcs(reinterpret_cast<const double2 *>(oldF + s0[t])) : *reinterpret_cast<const double2 *>(oldF + s0[t]);
        } else {
            v[t].x = HINT ? __ldcs(oldF + s0[t]) : oldF[s0[t]];
            if (s1[t] >= 0) v[t].y = HINT ? __ldcs(oldF + s1[t]) : oldF[s1[t]];
        }
    }
#pragma unroll
    for (int t = 0; t < ITEMS; t++) {
        if (s0[t] < 0) continue;
        const int64_t e = 2 * (p0 + (int64_t)t * kThreads);
        if (s1[t] >= 0) { if (HINT) __stcs(reinterpret_cast<double2 *>(newF + e), v[t]); else *reinterpret_cast<double2 *>(newF + e) = v[t]; }
        else newF[e] = v[t].x;
    }
}
template <int NC, int ITEMS>
__global__ void __launch_bounds__(kThreads) v_elem_cs(int64_t nNew, const int *__restrict__ cellMap, const double *__restrict__ oldF, double *__restrict__ newF) {
    const int64_t total = nNew * NC;
    const int64_t e0 = (int64_t)blockIdx.x * (kThreads * ITEMS) + threadIdx.x;
    int64_t src[ITEMS];
#pragma unroll
    for (int t = 0; t < ITEMS; t++) {
        const int64_t e = e0 + (int64_t)t * kThreads;
        if (e < total) { const int64_t i = e / NC; src[t] = (int64_t)__ldcs(cellMap + i) * NC + (e - i * NC); } else src[t] = -1;
    }
    double v[ITEMS];
#pragma unroll
    for (int t = 0; t < ITEMS; t++) if (src[t] >= 0) v[t] = __ldcs(oldF + src[t]);
#pragma unroll
    for (int t = 0; t < ITEMS; t++) if (src[t] >= 0) __stcs(newF + e0 + (int64_t)t * kThreads, v[t]);
}
// persistent grid-stride form of the element kernel
template <int NC, int ITEMS>
__global__ void __launch_bounds__(kThreads) v_elem_persist(int64_t nNew, const int *__restrict__ cellMap, const double *__restrict__ oldF, double *__restrict__ newF) {
    const int64_t total = nNew * NC;
    for (int64_t base = (int64_t)blockIdx.x * (kThreads * ITEMS); base < total; base += (int64_t)gridDim.x * (kThreads * ITEMS)) {
        const int64_t e0 = base + threadIdx.x;
        int64_t src[ITEMS];
#pragma unroll
        for (int t = 0; t < ITEMS; t++) {
            const int64_t e = e0 + (int64_t)t * kThreads;
            if (e < total) { const int64_t i = e / NC; src[t] = (int64_t)cellMap[i] * NC + (e - i * NC); } else src[t] = -1;
        }
        double v[ITEMS];
#pragma unroll
        for (int t = 0; t < ITEMS; t++) if (src[t] >= 0) v[t] = oldF[src[t]];
#pragma unroll
        for (int t = 0; t < ITEMS; t++) if (src[t] >= 0) newF[e0 + (int64_t)t * kThreads] = v[t];
    }
}

// tile variant: a CTA owns TILE consecutive new cells, stages their cellMap entries in shared memory once and then
// streams every field of the batch through them (coalesced element-wise)
constexpr int TILE = 1024;
struct Fields { int n; int nc[4]; const double *oldF[4]; double *newF[4]; };
template <int BATCH>
__global__ void __launch_bounds__(kThreads) v_tile(int64_t nNew, const int *__restrict__ cellMap, Fields F) {
    __shared__ int sm[TILE];
    const int64_t c0 = (int64_t)blockIdx.x * TILE;
    const int cnt = (int)min((int64_t)TILE, nNew - c0);
    for (int k = threadIdx.x; k < cnt; k += kThreads) sm[k] = cellMap[c0 + k];
    __syncthreads();
    for (int fi = 0; fi < F.n; fi++) {
        const int nc = F.nc[fi];
        const double *__restrict__ o = F.oldF[fi];
        double *__restrict__ d = F.newF[fi] + c0 * nc;
        const int total = cnt * nc;
        for (int e0 = threadIdx.x; e0 < total; e0 += kThreads * BATCH) {
            double v[BATCH];
#pragma unroll
            for (int t = 0; t < BATCH; t++) {
                const int e = e0 + t * kThreads;
                if (e < total) { const int i = e / nc; v[t] = o[(int64_t)sm[i] * nc + (e - i * nc)]; }
            }
#pragma unroll
            for (int t = 0; t < BATCH; t++) { const int e = e0 + t * kThreads; if (e < total) d[e] = v[t]; }
        }
    }
}
// same with 128-bit stores for the vector field: thread handles 2 consecutive elements
template <int BATCH>
__global__ void __launch_bounds__(kThreads) v_tile2(int64_t nNew, const int *__restrict__ cellMap, Fields F) {
    __shared__ int sm[TILE];
    const int64_t c0 = (int64_t)blockIdx.x * TILE;
    const int cnt = (int)min((int64_t)TILE, nNew - c0);
    for (int k = threadIdx.x; k < cnt; k += kThreads) sm[k] = cellMap[c0 + k];
    __syncthreads();
    for (int fi = 0; fi < F.n; fi++) {
        const int nc = F.nc[fi];
        const double *__restrict__ o = F.oldF[fi];
        double *__restrict__ d = F.newF[fi] + c0 * nc;       // c0*nc even because TILE is even
        const int total = cnt * nc;                            // even unless the last tile
        for (int p0 = threadIdx.x; 2 * p0 < total; p0 += kThreads * BATCH) {
            double2 v[BATCH];
#pragma unroll
            for (int t = 0; t < BATCH; t++) {
                const int e = 2 * (p0 + t * kThreads);
                if (e < total) {
                    const int i = e / nc, k = e - i * nc;
                    v[t].x = o[(int64_t)sm[i] * nc + k];
                    if (e + 1 < total) { const int i2 = (k + 1 == nc) ? i + 1 : i, k2 = (k + 1 == nc) ? 0 : k + 1; v[t].y = o[(int64_t)sm[i2] * nc + k2]; }
                }
            }
#pragma unroll
            for (int t = 0; t < BATCH; t++) {
                const int e = 2 * (p0 + t * kThreads);
                if (e + 1 < total) *reinterpret_cast<double2 *>(d + e) = v[t];
                else if (e < total) d[e] = v[t].x;
            }
        }
    }
}

int main() {
    const int64_t n = 64000000, nRef = 2050416, nNew = n + 7 * nRef;
    std::vector<int> map((size_t)nNew);
    for (int64_t i = 0; i < n; i++) map[(size_t)i] = (int)i;
    for (int64_t r = 0; r < nRef; r++) for (int k = 0; k < 7; k++) map[(size_t)(n + 7 * r + k)] = (int)((r * 31) % n);
    int *dMap; double *o1, *o3, *oe, *n1, *n3, *ne;
    CK(cudaMalloc(&dMap, nNew * 4)); CK(cudaMemcpy(dMap, map.data(), nNew * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&o1, n * 8)); CK(cudaMalloc(&o3, n * 24)); CK(cudaMalloc(&oe, n * 8));
    CK(cudaMalloc(&n1, nNew * 8)); CK(cudaMalloc(&n3, nNew * 24)); CK(cudaMalloc(&ne, nNew * 8));
    CK(cudaMemset(o1, 1, n * 8)); CK(cudaMemset(o3, 1, n * 24)); CK(cudaMemset(oe, 1, n * 8));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto timeit = [&](const char *name, double bytes, auto fn) {
        for (int w = 0; w < 3; w++) fn();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(a);
        const int reps = 10;
        for (int r = 0; r < reps; r++) fn();
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
        printf("%-44s %8.4f ms  %8.1f GB/s\n", name, ms, bytes / ms * 1e-6);
        CK(cudaGetLastError());
    };
    const double b1 = 4.0 * nNew + 8.0 * n + 8.0 * nNew, b3 = 4.0 * nNew + 24.0 * n + 24.0 * nNew;
    const double b5 = 4.0 * nNew + 40.0 * n + 40.0 * nNew;
    timeit("memcpy D2D 1.7 GB (r+w counted)", 2.0 * nNew * 8 * 1.0, [&] { cudaMemcpyAsync(n1, o1 /*512MB*/, n * 8, cudaMemcpyDeviceToDevice); });
    timeit("memcpy D2D vec (r+w)", 2.0 * n * 24, [&] { cudaMemcpyAsync(n3, o3, n * 24, cudaMemcpyDeviceToDevice); });
    auto g = [&](int64_t items, int per) { return (unsigned)((items + per - 1) / per); };
    timeit("elem NC=1 x4", b1, [&] { v_elem<1, 4><<<g(nNew, kThreads * 4), kThreads>>>(nNew, dMap, o1, n1); });
    timeit("elem NC=1 x8", b1, [&] { v_elem<1, 8><<<g(nNew, kThreads * 8), kThreads>>>(nNew, dMap, o1, n1); });
    timeit("elem NC=3 x4 (current)", b3, [&] { v_elem<3, 4><<<g(nNew * 3, kThreads * 4), kThreads>>>(nNew, dMap, o3, n3); });
    timeit("elem NC=3 x8", b3, [&] { v_elem<3, 8><<<g(nNew * 3, kThreads * 8), kThreads>>>(nNew, dMap, o3, n3); });
    timeit("elem NC=3 x2", b3, [&] { v_elem<3, 2><<<g(nNew * 3, kThreads * 2), kThreads>>>(nNew, dMap, o3, n3); });
    timeit("elem NC=3 x4 streaming hints", b3, [&] { v_elem_cs<3, 4><<<g(nNew * 3, kThreads * 4), kThreads>>>(nNew, dMap, o3, n3); });
    timeit("elem NC=1 x4 streaming hints", b1, [&] { v_elem_cs<1, 4><<<g(nNew, kThreads * 4), kThreads>>>(nNew, dMap, o1, n1); });
    timeit("elem NC=3 x4 persistent 148x8", b3, [&] { v_elem_persist<3, 4><<<148 * 8, kThreads>>>(nNew, dMap, o3, n3); });
    timeit("pair NC=3 x2", b3, [&] { v_pair<3, 2, 0><<<g((nNew * 3 + 1) / 2, kThreads * 2), kThreads>>>(nNew, dMap, o3, n3); });
    timeit("pair NC=3 x4", b3, [&] { v_pair<3, 4, 0><<<g((nNew * 3 + 1) / 2, kThreads * 4), kThreads>>>(nNew, dMap, o3, n3); });
    timeit("pair NC=3 x6", b3, [&] { v_pair<3, 6, 0><<<g((nNew * 3 + 1) / 2, kThreads * 6), kThreads>>>(nNew, dMap, o3, n3); });
    timeit("pair NC=3 x8", b3, [&] { v_pair<3, 8, 0><<<g((nNew * 3 + 1) / 2, kThreads * 8), kThreads>>>(nNew, dMap, o3, n3); });
    timeit("pair NC=1 x4", b1, [&] { v_pair<1, 4, 0><<<g((nNew + 1) / 2, kThreads * 4), kThreads>>>(nNew, dMap, o1, n1); });
    timeit("pair NC=1 x8", b1, [&] { v_pair<1, 8, 0><<<g((nNew + 1) / 2, kThreads * 8), kThreads>>>(nNew, dMap, o1, n1); });
    timeit("pair x4: 1,3,1 separately", b1 * 2 + b3, [&] {
        v_pair<1, 4, 0><<<g((nNew + 1) / 2, kThreads * 4), kThreads>>>(nNew, dMap, o1, n1);
        v_pair<3, 4, 0><<<g((nNew * 3 + 1) / 2, kThreads * 4), kThreads>>>(nNew, dMap, o3, n3);
        v_pair<1, 4, 0><<<g((nNew + 1) / 2, kThreads * 4), kThreads>>>(nNew, dMap, oe, ne); });
    timeit("pair NC=3 x2 streaming", b3, [&] { v_pair<3, 2, 1><<<g((nNew * 3 + 1) / 2, kThreads * 2), kThreads>>>(nNew, dMap, o3, n3); });
    timeit("pair NC=3 x4 streaming", b3, [&] { v_pair<3, 4, 1><<<g((nNew * 3 + 1) / 2, kThreads * 4), kThreads>>>(nNew, dMap, o3, n3); });
    timeit("pair NC=1 x2", b1, [&] { v_pair<1, 2, 0><<<g((nNew + 1) / 2, kThreads * 2), kThreads>>>(nNew, dMap, o1, n1); });
    timeit("pair NC=1 x4 streaming", b1, [&] { v_pair<1, 4, 1><<<g((nNew + 1) / 2, kThreads * 4), kThreads>>>(nNew, dMap, o1, n1); });
    Fields f3{1, {3, 0, 0, 0}, {o3, nullptr, nullptr, nullptr}, {n3, nullptr, nullptr, nullptr}};
    Fields f5{3, {1, 3, 1, 0}, {o1, o3, oe, nullptr}, {n1, n3, ne, nullptr}};
    const unsigned gt = g(nNew, TILE);
    timeit("tile NC=3 batch4", b3, [&] { v_tile<4><<<gt, kThreads>>>(nNew, dMap, f3); });
    timeit("tile NC=3 batch6", b3, [&] { v_tile<6><<<gt, kThreads>>>(nNew, dMap, f3); });
    timeit("tile NC=3 batch12", b3, [&] { v_tile<12><<<gt, kThreads>>>(nNew, dMap, f3); });
    timeit("tile2 NC=3 batch3 (128-bit stores)", b3, [&] { v_tile2<3><<<gt, kThreads>>>(nNew, dMap, f3); });
    timeit("tile2 NC=3 batch6 (128-bit stores)", b3, [&] { v_tile2<6><<<gt, kThreads>>>(nNew, dMap, f3); });
    timeit("tile fused 1+3+1 batch4", b5, [&] { v_tile<4><<<gt, kThreads>>>(nNew, dMap, f5); });
    timeit("tile fused 1+3+1 batch6", b5, [&] { v_tile<6><<<gt, kThreads>>>(nNew, dMap, f5); });
    timeit("tile2 fused 1+3+1 batch4", b5, [&] { v_tile2<4><<<gt, kThreads>>>(nNew, dMap, f5); });
    timeit("separate 1,3,1 (current path)", b1 * 2 + b3, [&] {
        v_elem<1, 4><<<g(nNew, kThreads * 4), kThreads>>>(nNew, dMap, o1, n1);
        v_elem<3, 4><<<g(nNew * 3, kThreads * 4), kThreads>>>(nNew, dMap, o3, n3);
        v_elem<1, 4><<<g(nNew, kThreads * 4), kThreads>>>(nNew, dMap, oe, ne); });
    return 0;
}
