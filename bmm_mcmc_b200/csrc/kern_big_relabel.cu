// Stephens relabelling for the grid path (one chain over N observations, possibly N-sharded):
// /root/reference/src/stephens.cpp:6-94 restated as streaming kernels over row-major float matrices
// P (this sweep's conditional probabilities, N x K) and Q (the running reference, N x K).
//
//   online (stephens.cpp:66-94), per sweep j >= burnin:
//     grid_cost_kernel    G(k,l) = sum_i log q_ik * p_il   and  s_l = sum_i p_il^2, a register-tiled
//                         K x K contraction over the observations (fp32 tiles, fp64 cross-block sums);
//                         C(k,l) = s_l - G(k,l)  (the reference's p*(p - log q), quirk 4)
//     [all-reduce]        K*K + K doubles when N-sharded
//     grid_assign_kernel  K! enumeration / warp-parallel Jonker-Volgenant -> perm (replaces my_lpsolve)
//     grid_qupdate_kernel Q' = j (Q + P[:, perm]) / (j + 1)   (quirks 3, 5)
//   batch (stephens.cpp:6-64), once at j == burnin - 1 over the M stored sweeps: 100 x { Q = mean of the
//     permuted slices; per slice C_t(k,l) = sum_i p (log p - log q); assignment } with the same kernels.
#include "assign.cuh"
#include "kernels.h"

namespace bmm {
namespace {

constexpr int GC_TP = 32;   // observations per shared-memory tile

// out[k + K*l] += sum_i LQ_ik * P_il ; out[K*K + l] += sum_i (use_logp ? p log p : p*p)
template <int T>
__global__ void __launch_bounds__(256) grid_cost_kernel(long long N, int K, const float *__restrict__ P,
                                                        const float *__restrict__ Q, int use_logp, double *out) {
    constexpr int KP = 16 * T;
    __shared__ __align__(16) float sP[GC_TP][KP];
    __shared__ __align__(16) float sLQ[GC_TP][KP];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[T][T];
#pragma unroll
    for (int m = 0; m < T; ++m)
#pragma unroll
        for (int n = 0; n < T; ++n) acc[m][n] = 0.f;
    float sacc[T];
#pragma unroll
    for (int n = 0; n < T; ++n) sacc[n] = 0.f;
    const long long ntiles = (N + GC_TP - 1) / GC_TP;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long i0 = tile * GC_TP;
        for (int e = tid; e < GC_TP * KP; e += 256) {
            const int i = e / KP, k = e % KP;
            float p = 0.f, lq = 0.f;
            if (k < K && i0 + i < N) {
                p = P[(size_t)(i0 + i) * K + k];
                lq = __logf(Q[(size_t)(i0 + i) * K + k]);
            }
            sP[i][k] = p; sLQ[i][k] = lq;
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < GC_TP; ++i) {
            float a[T], b[T];
#pragma unroll
            for (int m = 0; m < T; ++m) a[m] = sLQ[i][ty * T + m];
#pragma unroll
            for (int n = 0; n < T; ++n) b[n] = sP[i][tx * T + n];
#pragma unroll
            for (int m = 0; m < T; ++m)
#pragma unroll
                for (int n = 0; n < T; ++n) acc[m][n] = fmaf(a[m], b[n], acc[m][n]);
            if (ty == 0) {
#pragma unroll
                for (int n = 0; n < T; ++n) sacc[n] += use_logp ? (b[n] > 0.f ? b[n] * __logf(b[n]) : 0.f) : b[n] * b[n];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < T; ++m)
#pragma unroll
        for (int n = 0; n < T; ++n) {
            const int k = ty * T + m, l = tx * T + n;
            if (k < K && l < K) atomicAdd(&out[k + (size_t)K * l], (double)acc[m][n]);
        }
    if (ty == 0) {
#pragma unroll
        for (int n = 0; n < T; ++n) {
            const int l = tx * T + n;
            if (l < K) atomicAdd(&out[(size_t)K * K + l], (double)sacc[n]);
        }
    }
}

// Warp-parallel shortest-augmenting-path assignment (Jonker-Volgenant potentials): lanes share the
// column scan, the arg-min is a warp reduction.  Same optimum as assign_jv_thread; ties broken towards
// the smallest column index like the serial scan.  Workspace as assign_ws_bytes(K).
__device__ void assign_jv_warp(int K, const double *cost, void *ws, int *col_to_row) {
    const int lane = threadIdx.x & 31;
    const double INF = 1e300;
    double *u = (double *)ws, *v = u + (K + 1), *minv = v + (K + 1);
    int *p = (int *)(minv + (K + 1)), *way = p + (K + 1);
    unsigned char *used = (unsigned char *)(way + (K + 1));
    for (int j = lane; j <= K; j += 32) { u[j] = 0.0; v[j] = 0.0; p[j] = 0; way[j] = 0; }
    __syncwarp();
    for (int i = 1; i <= K; ++i) {
        if (lane == 0) p[0] = i;
        for (int j = lane; j <= K; j += 32) { minv[j] = INF; used[j] = 0; }
        __syncwarp();
        int j0 = 0;
        for (;;) {
            if (lane == 0) used[j0] = 1;
            __syncwarp();
            const int i0 = p[j0];
            const double ui0 = u[i0];
            double best = INF;
            int bj = 0x7fffffff;
            for (int j = 1 + lane; j <= K; j += 32)
                if (!used[j]) {
                    const double cur = cost[(i0 - 1) + (size_t)K * (j - 1)] - ui0 - v[j];
                    if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
                    if (minv[j] < best) { best = minv[j]; bj = j; }
                }
#pragma unroll
            for (int off = 16; off; off >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
                if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
            }
            int j1 = bj;
            double delta = best;
            if (j1 == 0x7fffffff) {  // non-finite costs: any free column, so the loop terminates
                delta = 0.0;
                j1 = 0;
                for (int j = 1; j <= K; ++j) if (!used[j]) { j1 = j; break; }
            }
            __syncwarp();
            for (int j = lane; j <= K; j += 32) {
                if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
                else minv[j] -= delta;
            }
            __syncwarp();
            j0 = j1;
            if (p[j0] == 0) break;
        }
        if (lane == 0) {
            do { const int j1 = way[j0]; p[j0] = p[j1]; j0 = j1; } while (j0);
        }
        __syncwarp();
    }
    for (int j = 1 + lane; j <= K; j += 32) col_to_row[j - 1] = p[j] - 1;
    __syncwarp();
}

// One warp: cost = s_l - G(k,l) (column-major k + K*l), assignment, permutation bookkeeping.
// perm_dst[l * perm_stride] receives perm[l]; perm_cur (optional) the same, contiguous.
__global__ void grid_assign_kernel(int K, double *acc, char *ws, int *perm_cur, int *perm_dst, int perm_stride) {
    const int lane = threadIdx.x;
    for (int e = lane; e < K * K; e += 32) acc[e] = acc[(size_t)K * K + e / K] - acc[e];
    __syncwarp();
    __shared__ int c2r[256];
    if (K <= ASSIGN_ENUM_MAXK) { if (lane == 0) assign_enum_thread(K, acc, c2r); }
    else assign_jv_warp(K, acc, ws, c2r);
    __syncwarp();
    for (int l = lane; l < K; l += 32) {
        if (perm_cur) perm_cur[l] = c2r[l];
        if (perm_dst) perm_dst[(size_t)l * perm_stride] = c2r[l];
    }
}

__global__ void grid_qupdate_kernel(long long N, int K, float *__restrict__ Q, const float *__restrict__ P,
                                    const int *__restrict__ perm, int sample_num) {
    const float sn = (float)sample_num, inv = 1.f / (float)(sample_num + 1);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < N * K; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / K;
        const int k = (int)(e % K);
        Q[e] = sn * (Q[e] + P[i * K + perm[k]]) * inv;   // (stephens.cpp:87-92)
    }
}

// batch: p.replace(0, 1e-6) (stephens.cpp:30-31)
__global__ void grid_clamp_kernel(long long n, float *p) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        if (p[e] == 0.f) p[e] = 1e-6f;
}

// batch: Q[:, k] = mean_t cube[t][:, perm[t][k]] (stephens.cpp:37-43); perm is [M][K]
__global__ void grid_qmean_kernel(long long N, int K, int M, const float *__restrict__ cube, const int *__restrict__ perm,
                                  float *__restrict__ Q) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < N * K; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / K;
        const int k = (int)(e % K);
        float acc = 0.f;
        for (int t = 0; t < M; ++t) acc += cube[(size_t)t * N * K + i * K + perm[t * K + k]];
        Q[e] = acc / (float)M;
    }
}

__global__ void grid_identity_perm_kernel(int n, int K, int *perm) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) perm[e] = e % K;
}

}  // namespace

cudaError_t launch_grid_cost(long long N, int K, const float *P, const float *Q, int use_logp, double *acc,
                             int sm_count, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(acc, 0, ((size_t)K * K + K) * sizeof(double), st);
    if (e != cudaSuccess) return e;
    const long long ntiles = (N + GC_TP - 1) / GC_TP;
    const int grid = (int)(ntiles < 2LL * sm_count ? (ntiles < 1 ? 1 : ntiles) : 2LL * sm_count);
    const int T = (K + 15) / 16;
    if (T <= 1) grid_cost_kernel<1><<<grid, 256, 0, st>>>(N, K, P, Q, use_logp, acc);
    else if (T <= 2) grid_cost_kernel<2><<<grid, 256, 0, st>>>(N, K, P, Q, use_logp, acc);
    else if (T <= 4) grid_cost_kernel<4><<<grid, 256, 0, st>>>(N, K, P, Q, use_logp, acc);
    else if (T <= 8) grid_cost_kernel<8><<<grid, 256, 0, st>>>(N, K, P, Q, use_logp, acc);
    else return cudaErrorInvalidValue;
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_grid_assign(int K, double *acc, char *ws, int *perm_cur, int *perm_dst, int perm_stride, cudaStream_t st) {
    grid_assign_kernel<<<1, 32, 0, st>>>(K, acc, ws, perm_cur, perm_dst, perm_stride);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_grid_qupdate(long long N, int K, float *Q, const float *P, const int *perm, int sample_num,
                                int sm_count, cudaStream_t st) {
    grid_qupdate_kernel<<<sm_count * 8, 256, 0, st>>>(N, K, Q, P, perm, sample_num);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_grid_clamp(long long n, float *p, int sm_count, cudaStream_t st) {
    grid_clamp_kernel<<<sm_count * 8, 256, 0, st>>>(n, p);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_grid_qmean(long long N, int K, int M, const float *cube, const int *perm, float *Q, int sm_count,
                              cudaStream_t st) {
    grid_qmean_kernel<<<sm_count * 8, 256, 0, st>>>(N, K, M, cube, perm, Q);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_grid_identity_perm(int n, int K, int *perm, cudaStream_t st) {
    grid_identity_perm_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, K, perm);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace bmm
