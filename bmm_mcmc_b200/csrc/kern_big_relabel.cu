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
#include <cstdlib>

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

// Warp-parallel shortest-augmenting-path assignment (Jonker-Volgenant potentials).  Lane `l` owns
// columns j = 1 + l + 32 m and keeps their whole state in registers -- column potential v, reduced
// slack minv, predecessor way, matched row p and THAT ROW's potential ucol (a row's potential travels
// with the row when the matching changes) -- so one step of the path search is a conflict-free row
// read from shared memory plus a warp arg-min, not a chain of dependent shared-memory round trips
// (measured: 2 900 -> ~350 cycles per step at K = 128).  costT is row-major: costT[(i-1)*K + (j-1)].
// Starts from the row reduction u_i = min_j c_ij (dual feasible with v = 0): a row whose minimum
// column is free is matched at once, which settles every row once relabelling has converged.
// Ties go to the smallest column index, like the serial scan of assign_jv_thread.
// Warm start (vws != NULL and vws[K] != 0): the column potentials of the previous solve are the starting duals, v = 0
// otherwise.  Any v is dual feasible with u_i = min_j (c_ij - v_j), so the result is optimal either way; between two
// sweeps the cost matrix moves little and nearly every row finds its previous column free and tight in the row
// reduction, instead of a ~K-step augmenting path per row when the columns of unoccupied clusters tie (measured at C4,
// K = 32: 212 us per solve cold).
template <int M>
__device__ void assign_jv_warp(int K, const double *costT, double *urow, int *col_to_row, double *vws) {
    const int lane = threadIdx.x & 31;
    const double INF = 1e300;
    double v[M], minv[M], ucol[M];
    int p[M], way[M];
    unsigned usedm = 0;
#pragma unroll
    for (int m = 0; m < M; ++m) { v[m] = 0.0; minv[m] = INF; ucol[m] = 0.0; p[m] = 0; way[m] = 0; }
    if (vws && vws[K] != 0.0) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const int j = 1 + lane + 32 * m;
            if (j <= K) { const double w = vws[j - 1]; v[m] = (w == w && fabs(w) < 1e300) ? w : 0.0; }
        }
    }
    auto sel_i = [&](const int (&a)[M], int m0) { int r = a[0];
#pragma unroll
        for (int m = 1; m < M; ++m) r = (m0 == m) ? a[m] : r; return r; };
    auto sel_d = [&](const double (&a)[M], int m0) { double r = a[0];
#pragma unroll
        for (int m = 1; m < M; ++m) r = (m0 == m) ? a[m] : r; return r; };
    auto argmin = [&](double &best, int &bj) {
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
            if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
        }
    };
    // ---- row reduction ----
    unsigned long long done_lo = 0, done_hi = 0, done_2 = 0, done_3 = 0;   // rows 1..K matched at start (bitset, uniform)
    auto set_done = [&](int i) { if (i < 64) done_lo |= 1ull << i; else if (i < 128) done_hi |= 1ull << (i - 64);
                                 else if (i < 192) done_2 |= 1ull << (i - 128); else done_3 |= 1ull << (i - 192); };
    auto is_done = [&](int i) { return (int)(((i < 64) ? done_lo >> i : (i < 128) ? done_hi >> (i - 64)
                                              : (i < 192) ? done_2 >> (i - 128) : done_3 >> (i - 192)) & 1ull); };
    for (int i = 1; i <= K; ++i) {
        double best = INF;
        int bj = 0x7fffffff;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const int j = 1 + lane + 32 * m;
            if (j <= K) {
                const double c = costT[(size_t)(i - 1) * K + (j - 1)] - v[m];
                if (c < best) { best = c; bj = j; }
            }
        }
        argmin(best, bj);
        if (lane == 0) urow[i] = bj == 0x7fffffff ? 0.0 : best;
        if (bj != 0x7fffffff) {
            const int owner = (bj - 1) & 31, m0 = (bj - 1) >> 5;
            const int taken = __shfl_sync(0xffffffffu, sel_i(p, m0), owner);
            if (!taken) {
                if (lane == owner) {
#pragma unroll
                    for (int m = 0; m < M; ++m) if (m == m0) { p[m] = i; ucol[m] = best; }
                }
                set_done(i);
            }
        }
    }
    __syncwarp();
    // ---- augmenting paths for the rows still unmatched ----
    for (int i = 1; i <= K; ++i) {
        if (is_done(i)) continue;
        int p0 = i;
        double u0 = urow[i];
        usedm = 0;
#pragma unroll
        for (int m = 0; m < M; ++m) minv[m] = INF;
        int j0 = 0;
        for (;;) {
            int i0;
            double ui0;
            if (j0 == 0) { i0 = p0; ui0 = u0; }
            else {
                const int owner = (j0 - 1) & 31, m0 = (j0 - 1) >> 5;
                if (lane == owner) usedm |= 1u << m0;
                i0 = __shfl_sync(0xffffffffu, sel_i(p, m0), owner);
                ui0 = __shfl_sync(0xffffffffu, sel_d(ucol, m0), owner);
            }
            const double *crow = costT + (size_t)(i0 - 1) * K;
            double best = INF;
            int bj = 0x7fffffff;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int j = 1 + lane + 32 * m;
                if (j <= K && !((usedm >> m) & 1u)) {
                    const double cur = crow[j - 1] - ui0 - v[m];
                    if (cur < minv[m]) { minv[m] = cur; way[m] = j0; }
                    if (minv[m] < best) { best = minv[m]; bj = j; }
                }
            }
            argmin(best, bj);
            double delta = best;
            int j1 = bj;
            if (j1 == 0x7fffffff) {  // non-finite costs: any free column, so the loop terminates
                delta = 0.0;
                int cand = 0x7fffffff;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const int j = 1 + lane + 32 * m;
                    if (j <= K && !((usedm >> m) & 1u) && j < cand) cand = j;
                }
                j1 = __reduce_min_sync(0xffffffffu, cand);
                const int owner = (j1 - 1) & 31, m0 = (j1 - 1) >> 5;
                if (lane == owner) {
#pragma unroll
                    for (int m = 0; m < M; ++m) if (m == m0) way[m] = j0;
                }
            }
#pragma unroll
            for (int m = 0; m < M; ++m) {
                if ((usedm >> m) & 1u) { ucol[m] += delta; v[m] -= delta; }
                else minv[m] -= delta;
            }
            u0 += delta;
            j0 = j1;
            const int owner = (j0 - 1) & 31, m0 = (j0 - 1) >> 5;
            if (__shfl_sync(0xffffffffu, sel_i(p, m0), owner) == 0) break;
        }
        // unwind: column j0 takes the row (and its potential) of its predecessor column
        do {
            const int owner = (j0 - 1) & 31, m0 = (j0 - 1) >> 5;
            const int j1 = __shfl_sync(0xffffffffu, sel_i(way, m0), owner);
            int pj1 = p0;
            double uj1 = u0;
            if (j1 != 0) {
                const int o1 = (j1 - 1) & 31, m1 = (j1 - 1) >> 5;
                pj1 = __shfl_sync(0xffffffffu, sel_i(p, m1), o1);
                uj1 = __shfl_sync(0xffffffffu, sel_d(ucol, m1), o1);
            }
            if (lane == owner) {
#pragma unroll
                for (int m = 0; m < M; ++m) if (m == m0) { p[m] = pj1; ucol[m] = uj1; }
            }
            j0 = j1;
        } while (j0);
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int j = 1 + lane + 32 * m;
        if (j <= K) { col_to_row[j - 1] = p[m] - 1; if (vws) vws[j - 1] = v[m]; }
    }
    if (vws && lane == 0) vws[K] = 1.0;
    __syncwarp();
}

// One warp: cost = s_l - G(k,l) (column-major k + K*l), assignment, permutation bookkeeping.
// perm_dst[l * perm_stride] receives perm[l]; perm_cur (optional) the same, contiguous.
// The solver's potentials / labels live in shared memory, and so does the cost matrix when it fits
// (cost_in_smem): every step of the augmenting-path search is a dependent chain of such accesses.
// acc holds G(k,l) at k + K*l and s_l at K*K + l; the cost C(k,l) = s_l - G(k,l) has rows k = reference
// labels and columns l = sample labels (stephens.cpp:78-84).
__global__ void grid_assign_kernel(int K, double *acc, int cost_in_smem, int *perm_cur, int *perm_dst, int perm_stride,
                                   int *changed, double *vws) {
    extern __shared__ __align__(16) char sm[];
    const int lane = threadIdx.x;
    __shared__ int c2r[256];
    if (K <= ASSIGN_ENUM_MAXK) {
        for (int e = lane; e < K * K; e += 32) acc[e] = acc[(size_t)K * K + e / K] - acc[e];
        __syncwarp();
        if (lane == 0) assign_enum_thread(K, acc, c2r);
    } else {
        // row-major copy (costT[k*K + l]) in shared memory when it fits, else in place in global memory
        double *urow = (double *)sm;
        double *costT = cost_in_smem ? urow + (K + 1) : acc;
        if (cost_in_smem) {
            for (int e = lane; e < K * K; e += 32) { const int k = e / K, l = e % K; costT[e] = acc[(size_t)K * K + l] - acc[k + (size_t)K * l]; }
        } else {
            for (int e = lane; e < K * K; e += 32) acc[e] = acc[(size_t)K * K + e / K] - acc[e];
            __syncwarp();
            for (int k = 0; k < K; ++k)        // in-place transpose to row-major
                for (int l = k + 1 + lane; l < K; l += 32) { const double t = acc[k + (size_t)K * l]; acc[k + (size_t)K * l] = acc[l + (size_t)K * k]; acc[l + (size_t)K * k] = t; }
        }
        __syncwarp();
        if (K <= 32) assign_jv_warp<1>(K, costT, urow, c2r, vws);
        else if (K <= 64) assign_jv_warp<2>(K, costT, urow, c2r, vws);
        else if (K <= 128) assign_jv_warp<4>(K, costT, urow, c2r, vws);
        else assign_jv_warp<8>(K, costT, urow, c2r, vws);
    }
    __syncwarp();
    for (int l = lane; l < K; l += 32) {
        if (perm_cur) perm_cur[l] = c2r[l];
        if (perm_dst) {
            if (changed && perm_dst[(size_t)l * perm_stride] != c2r[l]) *changed = 1;
            perm_dst[(size_t)l * perm_stride] = c2r[l];
        }
    }
}

// Index type of the element-wise relabelling kernels: 32-bit while N * K allows it (a 64-bit division per element by the
// run-time K held grid_qmean_kernel at 1.5 TB/s)
// fixed (BMM_FLAG_STEPHENS_FIXED): perm is the inverse permutation and Q' the running mean (j Q + p) / (j + 1)
template <typename I>
__global__ void grid_qupdate_kernel(long long N, int K, float *__restrict__ Q, const float *__restrict__ P,
                                    const int *__restrict__ perm, int sample_num, int fixed) {
    const float sn = (float)sample_num, inv = 1.f / (float)(sample_num + 1);
    const I total = (I)(N * K), step = (I)((long long)gridDim.x * blockDim.x);
    for (I e = (I)((long long)blockIdx.x * blockDim.x + threadIdx.x); e < total; e += step) {
        const I i = e / (I)K;
        const int k = (int)(e - i * (I)K);
        const float pr = P[i * K + perm[k]];
        Q[e] = fixed ? (sn * Q[e] + pr) * inv : sn * (Q[e] + pr) * inv;   // (stephens.cpp:87-92)
    }
}

// inv[t][perm[t][l]] = l for n = M permutations of K
__global__ void grid_invert_perm_kernel(int n, int K, const int *__restrict__ perm, int *__restrict__ inv) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * K; e += gridDim.x * blockDim.x)
        inv[(e / K) * K + perm[e]] = e % K;
}

// batch: p.replace(0, 1e-6) (stephens.cpp:30-31)
__global__ void grid_clamp_kernel(long long n, float *p) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        if (p[e] == 0.f) p[e] = 1e-6f;
}

// batch: Q[:, k] = mean_t cube[t][:, perm[t][k]] (stephens.cpp:37-43); perm is [M][K]
template <typename I>
__global__ void grid_qmean_kernel(long long N, int K, int M, const float *__restrict__ cube, const int *__restrict__ perm,
                                  float *__restrict__ Q) {
    const I total = (I)(N * K), step = (I)((long long)gridDim.x * blockDim.x);
    for (I e = (I)((long long)blockIdx.x * blockDim.x + threadIdx.x); e < total; e += step) {
        const I i = e / (I)K;
        const int k = (int)(e - i * (I)K);
        float acc = 0.f;
        for (int t = 0; t < M; ++t) acc += cube[(size_t)t * N * K + (size_t)i * K + perm[t * K + k]];
        Q[e] = acc / (float)M;
    }
}

// theta_rel[perm(s, k), :] = theta[k, :] for every kept sweep s at once (full_gibbs.cpp:221-223); perm is S x K column-major
__global__ void grid_theta_rel_kernel(int K, int P, int S, const double *__restrict__ theta, const int *__restrict__ perm,
                                      double *__restrict__ theta_rel) {
    const size_t KP = (size_t)K * P, n = KP * S;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const size_t s = e / KP, r = e % KP;
        const int k = (int)(r % K);
        const size_t d = r / K;
        theta_rel[KP * s + perm[s + (size_t)S * k] + (size_t)K * d] = theta[e];
    }
}

__global__ void grid_identity_perm_kernel(int n, int K, int *perm) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) perm[e] = e % K;
}

__global__ void grid_zfreq_kernel(long long N, int K, const uint8_t *__restrict__ z, const int *__restrict__ perm,
                                  unsigned *__restrict__ zfreq) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const int zi = (int)z[i] - 1;
        if (zi >= 0 && zi < K) zfreq[i + N * (perm ? perm[zi] : zi)] += 1u;
    }
}

}  // namespace

cudaError_t launch_grid_zfreq(long long N, int K, const uint8_t *z, const int *perm, unsigned *zfreq, int sm_count, cudaStream_t st) {
    grid_zfreq_kernel<<<sm_count * 8, 256, 0, st>>>(N, K, z, perm, zfreq);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_grid_cost(long long N, int K, const float *P, const float *Q, int use_logp, double *acc,
                             int sm_count, cudaStream_t st, int tc, int *status) {
    cudaError_t e = cudaMemsetAsync(acc, 0, ((size_t)K * K + K) * sizeof(double), st);
    if (e != cudaSuccess) return e;
    static const bool no_tc = getenv("BMM_NO_TC") != nullptr;  // A/B switch
    if (tc && status && !no_tc && grid_cost_tc_supported(N, K))
        return launch_grid_cost_tc(N, K, P, Q, use_logp, acc, status, sm_count, st);
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

// ws: (K + 1) doubles of warm-start state (column potentials + "valid"), zero before the first solve; NULL = cold start
cudaError_t launch_grid_assign(int K, double *acc, char *ws, int *perm_cur, int *perm_dst, int perm_stride, cudaStream_t st,
                               int *changed) {
    const size_t wsb = ((size_t)(K + 1) * sizeof(double) + 15) & ~(size_t)15, costb = (size_t)K * K * sizeof(double);
    const int in_smem = wsb + costb <= 200 * 1024;
    const size_t smem = wsb + (in_smem ? costb : 0);
    static FuncAttrCache attr;
    if (cudaError_t e = attr.ensure_smem(grid_assign_kernel, (int)smem)) return e;
    grid_assign_kernel<<<1, 32, smem, st>>>(K, acc, in_smem, perm_cur, perm_dst, perm_stride, changed, (double *)ws);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_grid_invert_perm(int n, int K, const int *perm, int *inv, cudaStream_t st) {
    grid_invert_perm_kernel<<<(n * K + 255) / 256, 256, 0, st>>>(n, K, perm, inv);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_grid_qupdate(long long N, int K, float *Q, const float *P, const int *perm, int sample_num,
                                int sm_count, cudaStream_t st, int fixed) {
    if (N * K < (1LL << 31) - (long long)sm_count * 8 * 256) grid_qupdate_kernel<unsigned><<<sm_count * 8, 256, 0, st>>>(N, K, Q, P, perm, sample_num, fixed);
    else grid_qupdate_kernel<long long><<<sm_count * 8, 256, 0, st>>>(N, K, Q, P, perm, sample_num, fixed);
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
    if (N * K < (1LL << 31) - (long long)sm_count * 8 * 256) grid_qmean_kernel<unsigned><<<sm_count * 8, 256, 0, st>>>(N, K, M, cube, perm, Q);
    else grid_qmean_kernel<long long><<<sm_count * 8, 256, 0, st>>>(N, K, M, cube, perm, Q);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_grid_theta_rel(int K, int P, int S, const double *theta, const int *perm, double *theta_rel, int sm_count, cudaStream_t st) {
    if (S < 1) return cudaSuccess;
    grid_theta_rel_kernel<<<sm_count * 4, 256, 0, st>>>(K, P, S, theta, perm, theta_rel);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_grid_identity_perm(int n, int K, int *perm, cudaStream_t st) {
    grid_identity_perm_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, K, perm);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace bmm
