// Stand-alone relabelling kernels: Stephens batch initialisation, one online step, batched
// assignment solves, and the Dirichlet helper.  Device code lives in stephens.cuh / assign.cuh.
#include "kernels.h"
#include "stephens.cuh"

namespace bmm {
namespace {

struct StephensBatchParams {
    int U, K, M;
    const int *wt;     // [U] or nullptr
    double *cube;      // [chain][M][U*K]   probabilities (zeros are replaced in place)
    double *logp;      // [chain][M][U*K]   scratch
    double *Q, *logQ;  // [chain][U*K]      out
    int *perm;         // [chain][M*K cm]   scratch / out (perm[t + M*k])
    double *cost;      // [chain][M][K*K]   scratch
    char *assign_ws;   // [chain][M][assign_ws_bytes(K)]
    int fixed;         // BMM_FLAG_STEPHENS_FIXED: store the inverse permutation (reference label -> sample column)
};

// One block per chain.  my_stephens_batch (stephens.cpp:6-64).
__global__ void stephens_batch_kernel(StephensBatchParams sp) {
    const int c = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int U = sp.U, K = sp.K, M = sp.M;
    const size_t UK = (size_t)U * K;
    double *cube = sp.cube + (size_t)c * M * UK, *logp = sp.logp + (size_t)c * M * UK;
    double *Q = sp.Q + (size_t)c * UK, *logQ = sp.logQ + (size_t)c * UK;
    int *perm = sp.perm + (size_t)c * M * K;
    double *cost = sp.cost + (size_t)c * M * K * K;
    const size_t wsb = assign_ws_bytes(K);
    char *ws = sp.assign_ws + (size_t)c * M * wsb;
    for (size_t e = tid; e < (size_t)M * UK; e += nthr) {
        double v = cube[e];
        if (v == 0.0) { v = 0.000001; cube[e] = v; }  // p.replace(0, min_prob) (:30-31)
        logp[e] = log(v);
    }
    for (int e = tid; e < M * K; e += nthr) perm[e] = e / M;  // perm(t,k) = k (:15-19)
    __syncthreads();
    for (int iter = 0; iter < 100; ++iter) {  // threshold 10^(-6) == -16: always maxiter (:24-25,33)
        for (size_t e = tid; e < UK; e += nthr) {
            const int u = (int)(e % U), k = (int)(e / U);
            double acc = 0.0;
            for (int t = 0; t < M; ++t) acc += cube[(size_t)t * UK + u + (size_t)U * perm[t + M * k]];
            acc /= M;
            Q[e] = acc;
            logQ[e] = log(acc);
        }
        __syncthreads();
        for (int e = tid; e < M * K * K; e += nthr) {
            const int t = e / (K * K), r = e % (K * K), k = r % K, l = r / K;
            const double *pl = cube + (size_t)t * UK + (size_t)U * l, *lp = logp + (size_t)t * UK + (size_t)U * l;
            const double *lq = logQ + (size_t)U * k;
            double acc = 0.0;
            if (sp.wt) { for (int u = 0; u < U; ++u) acc += sp.wt[u] * (pl[u] * (lp[u] - lq[u])); }
            else       { for (int u = 0; u < U; ++u) acc += pl[u] * (lp[u] - lq[u]); }
            cost[(size_t)t * K * K + k + K * l] = acc;
        }
        __syncthreads();
        int changed = 0;
        for (int t = tid; t < M; t += nthr) {
            int c2r[256];
            int *out = c2r;
            assign_thread(K, cost + (size_t)t * K * K, ws + (size_t)t * wsb, out);
            if (sp.fixed) {   // what `perm.row(iter) <- sort_index(...)` (:56) was meant to do
                for (int l = 0; l < K; ++l) { changed |= perm[t + M * out[l]] != l; perm[t + M * out[l]] = l; }
            } else {
                for (int k = 0; k < K; ++k) { changed |= perm[t + M * k] != out[k]; perm[t + M * k] = out[k]; }
            }
        }
        // Fixed point: with unchanged permutations the next iteration recomputes the same Q, the same costs
        // and the same assignments, so the remaining iterations of the reference's fixed 100 are no-ops
        // and Q already holds what they would return.
        if (!__syncthreads_or(changed)) break;
    }
}



__global__ void stephens_online_kernel(int U, int K, double *Q, double *logQ, const double *p, int sample_num,
                                       double *cost_out, int *perm_out, char *ws, int fixed) {
    extern __shared__ __align__(16) char sm[];
    double *cost = (double *)sm;
    int *perm = (int *)(cost + K * K);
    for (size_t e = threadIdx.x; e < (size_t)U * K; e += blockDim.x) logQ[e] = log(Q[e]);
    __syncthreads();
    stephens_online_block(U, K, nullptr, Q, logQ, p, sample_num, cost, perm, ws, fixed != 0);
    for (int e = threadIdx.x; e < K * K; e += blockDim.x) if (cost_out) cost_out[e] = cost[e];
    for (int e = threadIdx.x; e < K; e += blockDim.x) perm_out[e] = perm[e];
}

// my_lpsolve (my_lpsolve.cpp:6-31): one thread per problem; writes the 0/1 solution matrix.
__global__ void assign_kernel(int K, int batch, const double *cost, int *solution, char *ws) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    int c2r[256];
    assign_thread(K, cost + (size_t)b * K * K, ws + (size_t)b * assign_ws_bytes(K), c2r);
    int *sol = solution + (size_t)b * K * K;
    for (int e = 0; e < K * K; ++e) sol[e] = 0;
    for (int col = 0; col < K; ++col) sol[c2r[col] + K * col] = 1;
}

// rdirichlet_cpp (full_gibbs.cpp:10-27)
__global__ void rdirichlet_kernel(int K, const double *alpha_m, unsigned long long seed, double *out) {
    extern __shared__ double g[];
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        Stream st(seed, 0u, 0u, ST_MISC, (uint32_t)k);
        g[k] = st.gamma(alpha_m[k]);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double sum = 0.0;
        for (int j = 0; j < K; ++j) sum += g[j];
        out[k] = g[k] / sum;
    }
}

}  // namespace

cudaError_t launch_stephens_batch(int n_chains, int U, int K, int M, const int *wt, double *cube, double *logp,
                                  double *Q, double *logQ, int *perm, double *cost, char *assign_ws,
                                  cudaStream_t st, int fixed) {
    StephensBatchParams sp{U, K, M, wt, cube, logp, Q, logQ, perm, cost, assign_ws, fixed};
    stephens_batch_kernel<<<n_chains, 128, 0, st>>>(sp);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_stephens_online(int U, int K, double *Q, double *logQ, const double *p, int sample_num,
                                   double *cost, int *perm, char *assign_ws, cudaStream_t st, int fixed) {
    size_t smem = (size_t)K * K * 8 + (size_t)K * 4 + 16;
    cudaError_t e = cudaFuncSetAttribute(stephens_online_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    stephens_online_kernel<<<1, 256, smem, st>>>(U, K, Q, logQ, p, sample_num, cost, perm, assign_ws, fixed);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_assign(int K, int batch, const double *cost, int *solution, char *ws, cudaStream_t st) {
    assign_kernel<<<(batch + 63) / 64, 64, 0, st>>>(K, batch, cost, solution, ws);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_rdirichlet(int K, const double *alpha_m, unsigned long long seed, double *out, cudaStream_t st) {
    rdirichlet_kernel<<<1, 128, (size_t)K * 8, st>>>(K, alpha_m, seed, out);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace bmm
