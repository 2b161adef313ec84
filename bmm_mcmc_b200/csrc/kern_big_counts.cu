// Sufficient statistics of a grid-path sweep at large P: c_k = #{i: z_i = k} and V_kd = sum_{i: z_i = k} x_id
// (/root/reference/src/full_gibbs.cpp:182-200, stickbreaking.cpp:164-186) straight from the bit-packed
// rows, without expanding them.
//
//   cnt_hist_kernel     c_k: shared-memory histogram of the allocations, one global atomic per bin and block
//   cnt_offsets_kernel  exclusive scan of c_k -> start of every cluster's segment
//   cnt_scatter_kernel  counting sort: observation indices grouped by cluster (warp-aggregated cursor atomics)
//   cnt_accum_kernel    each block takes a slice of the sorted indices (one or two clusters), thread w owns
//                       32-bit word w of the rows and adds the rows' words with bit-sliced counters: seven
//                       carry-save bit planes hold per-bit counts up to 127, folded into 32 integer counters
//                       every 127 rows, flushed with one atomic per (cluster, variable) at the end of a run.
//
// Integer arithmetic throughout, every row is read exactly once (W consecutive words = one coalesced
// read).  Replaced a tcgen05 formulation ([X]^T onehot(z) as 128 x 128 x 128 MMAs over re-expanded bf16
// rows) that took 2.0 ms per 1e6 x 4096 sweep against 0.1-0.2 ms for this one: the contraction wastes
// the tensor pipe on a one-hot operand and its operand expansion was the bottleneck.
#include "common.cuh"
#include "kernels.h"

namespace bmm {
namespace {

constexpr int CNT_SLICE = 512;   // sorted observations per block of cnt_accum_kernel

__global__ void cnt_hist_kernel(long long N, int K, const uint8_t *__restrict__ z, int *__restrict__ ck) {
    __shared__ int h[256];
    for (int k = threadIdx.x; k < 256; k += blockDim.x) h[k] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
        atomicAdd(&h[z[i] - 1], 1);
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) if (h[k]) atomicAdd(&ck[k], h[k]);
}

// off[k] = sum_{l<k} c_l (k = 0..K), cursor[k] = off[k]
__global__ void cnt_offsets_kernel(int K, const int *__restrict__ ck, int *__restrict__ off, int *__restrict__ cursor) {
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < K; ++k) { off[k] = acc; cursor[k] = acc; acc += ck[k]; }
        off[K] = acc;
    }
}

// block = 1024 threads x 4 observations: ranks within the block from shared-memory atomics, one cursor
// atomic per (block, cluster)
constexpr int SC_PER = 4;
__global__ void __launch_bounds__(1024) cnt_scatter_kernel(long long N, int K, const uint8_t *__restrict__ z,
                                                           int *__restrict__ cursor, int *__restrict__ sorted) {
    __shared__ int h[256];
    for (int k = threadIdx.x; k < 256; k += blockDim.x) h[k] = 0;
    __syncthreads();
    const long long i0 = (long long)blockIdx.x * (1024 * SC_PER);
    int kk[SC_PER], rk[SC_PER];
#pragma unroll
    for (int u = 0; u < SC_PER; ++u) {
        const long long i = i0 + u * 1024 + threadIdx.x;
        kk[u] = i < N ? (int)z[i] - 1 : -1;
        rk[u] = kk[u] >= 0 ? atomicAdd(&h[kk[u]], 1) : 0;
    }
    __syncthreads();
    if ((int)threadIdx.x < K) { const int c = h[threadIdx.x]; h[threadIdx.x] = c ? atomicAdd(&cursor[threadIdx.x], c) : 0; }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < SC_PER; ++u)
        if (kk[u] >= 0) sorted[h[kk[u]] + rk[u]] = (int)(i0 + u * 1024 + threadIdx.x);
}

// grid: (ceil(N / CNT_SLICE), ceil(W / 128)); thread t owns word w = blockIdx.y * 128 + t
__global__ void __launch_bounds__(128) cnt_accum_kernel(long long N, int K, int P, int W, const uint32_t *__restrict__ xbits,
                                                        const int *__restrict__ sorted, const int *__restrict__ off,
                                                        int *__restrict__ Vkd) {
    __shared__ int s_off[257];
    __shared__ int s_idx[CNT_SLICE];
    for (int k = threadIdx.x; k <= K; k += blockDim.x) s_off[k] = off[k];
    const int w = blockIdx.y * 128 + threadIdx.x;
    const long long n0 = (long long)blockIdx.x * CNT_SLICE, n1 = min(n0 + (long long)CNT_SLICE, N);
    const bool active = w < W;
    int k = 0;
    while (k < K - 1 && (long long)s_off[k + 1] <= n0) ++k;      // cluster of the first entry of the slice
    uint32_t pl[7] = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
    int cnt[32];
#pragma unroll
    for (int b = 0; b < 32; ++b) cnt[b] = 0;
    for (int t = threadIdx.x; t < (int)(n1 - n0); t += blockDim.x) s_idx[t] = sorted[n0 + t];
    __syncthreads();
    auto add_row = [&](uint32_t carry) {       // carry-save increment of the per-bit counters
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const uint32_t t = pl[q] & carry;
            pl[q] ^= carry;
            carry = t;
        }
    };
    const uint32_t *xcol = xbits + (active ? w : 0);
    constexpr int U = 8, FOLD = 120;           // rows in flight per thread; rows per fold (planes hold <= 127)
    int n = 0;                                 // position within the slice
    const int nend = (int)(n1 - n0);
    while (n < nend) {
        while (k < K - 1 && (long long)s_off[k + 1] <= n0 + n) ++k;
        const int seg_end = (k == K - 1) ? nend : (int)min((long long)nend, (long long)s_off[k + 1] - n0);
        while (n < seg_end) {                  // one run of cluster k, FOLD rows at a time
            const int m_end = min(n + FOLD, seg_end);
            for (; n + U <= m_end; n += U) {
                uint32_t xw[U];
#pragma unroll
                for (int u = 0; u < U; ++u) xw[u] = __ldg(xcol + (size_t)s_idx[n + u] * W);
#pragma unroll
                for (int u = 0; u < U; ++u) add_row(xw[u]);
            }
            for (; n < m_end; ++n) add_row(__ldg(xcol + (size_t)s_idx[n] * W));
#pragma unroll
            for (int b = 0; b < 32; ++b) {     // fold the planes into the integer counters
                int c = 0;
#pragma unroll
                for (int q = 0; q < 7; ++q) c += (int)((pl[q] >> b) & 1u) << q;
                cnt[b] += c;
            }
#pragma unroll
            for (int q = 0; q < 7; ++q) pl[q] = 0u;
        }
        if (active) {
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const int d = 32 * w + b;
                if (cnt[b] && d < P) atomicAdd(&Vkd[k + (size_t)K * d], cnt[b]);
            }
        }
#pragma unroll
        for (int b = 0; b < 32; ++b) cnt[b] = 0;
    }
}

}  // namespace

// counts: [K + K*P] (c_k then V_kd at K + k + K*d), zero on entry; ws: 2*(K+1) + N ints of scratch
cudaError_t launch_big_counts(long long N, int K, int P, int W, const uint32_t *xbits, const uint8_t *z, int *counts,
                              int *ws, int sm_count, cudaStream_t st) {
    int *off = ws, *cursor = ws + (K + 1), *sorted = ws + 2 * (K + 1);
    cnt_hist_kernel<<<sm_count * 4, 256, 0, st>>>(N, K, z, counts);
    cnt_offsets_kernel<<<1, 32, 0, st>>>(K, counts, off, cursor);
    cnt_scatter_kernel<<<(unsigned)((N + 1024 * SC_PER - 1) / (1024 * SC_PER)), 1024, 0, st>>>(N, K, z, cursor, sorted);
    dim3 grid((unsigned)((N + CNT_SLICE - 1) / CNT_SLICE), (unsigned)((W + 127) / 128));
    cnt_accum_kernel<<<grid, 128, 0, st>>>(N, K, P, W, xbits, sorted, off, counts + K);
    g_launches += 4;
    return cudaGetLastError();
}

}  // namespace bmm
