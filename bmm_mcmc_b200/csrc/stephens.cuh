// Stephens (2000b) relabelling on the GPU: the online step (block-cooperative device function,
// called once per sweep from inside the sampler kernels) and the batch initialisation kernel.
// Restates /root/reference/src/stephens.cpp:6-94 *with its quirks* (SURVEY.md Appendix D 1-6):
// 100 fixed batch iterations, permutations never inverted, online cost uses p*(p - log q),
// Q' = j*(Q + p_reordered)/(j+1), zero->1e-6 clamping only in batch.  `fixed` (BMM_FLAG_STEPHENS_FIXED)
// selects the corrected variant of the online step instead; the default is the reference's.
//
// "Rows" are either observations (U = N, wt == nullptr) or unique data rows with multiplicities
// (the uncollapsed samplers, where every observation with the same x shares one probability row);
// sums over observations become weighted sums over rows.
#pragma once
#include "assign.cuh"

namespace bmm {

// cost (K x K, cm) and perm (K) live in shared memory; Q / logQ / p are U x K column-major.
// The cooperating threads are `nthr` threads with ranks `tid` that `sync()` synchronises: the whole block with
// __syncthreads, or one warp with __syncwarp when the sampler kernel relabels on one warp while another
// draws the parameters.
template <class Sync>
__device__ inline void stephens_online_group(int U, int K, const int *__restrict__ wt, double *Q, double *logQ,
                                             const double *p, int sample_num, double *cost, int *perm,
                                             void *assign_ws, bool fixed, const int tid, const int nthr, Sync sync) {
    if (fixed) {   // BMM_FLAG_STEPHENS_FIXED: log p in the cost, inverse permutation and running mean in the update
        for (int t = tid; t < K * K; t += nthr) {
            const int k = t % K, l = t / K;
            const double *pl = p + (size_t)U * l, *lq = logQ + (size_t)U * k;
            double acc = 0.0;
            for (int u = 0; u < U; ++u) {
                const double pv = pl[u], term = pv > 0.0 ? pv * (log(pv) - lq[u]) : 0.0;
                acc += wt ? wt[u] * term : term;
            }
            cost[k + K * l] = acc;
        }
        sync();
        if (tid == 0) assign_thread(K, cost, assign_ws, perm);
        sync();
        const double sn = (double)sample_num, sn1 = (double)(sample_num + 1);
        for (int e = tid; e < U * K; e += nthr) {
            const int u = e % U, k = e / U;
            int inv = 0;
            for (int l = 0; l < K; ++l) if (perm[l] == k) inv = l;   // sample column assigned to reference label k
            const double qn = (sn * Q[e] + p[u + (size_t)U * inv]) / sn1;
            Q[e] = qn;
            logQ[e] = log(qn);
        }
        sync();
        return;
    }
    for (int t = tid; t < K * K; t += nthr) {
        const int k = t % K, l = t / K;
        const double *pl = p + (size_t)U * l, *lq = logQ + (size_t)U * k;
        double acc = 0.0;
        if (wt) {
            for (int u = 0; u < U; ++u) { double pv = pl[u]; acc += wt[u] * (pv * (pv - lq[u])); }
        } else {
            for (int u = 0; u < U; ++u) { double pv = pl[u]; acc += pv * (pv - lq[u]); }  // p, not log p (:79)
        }
        cost[k + K * l] = acc;
    }
    sync();
    if (tid == 0) assign_thread(K, cost, assign_ws, perm);  // perm[l] = index_max(solution.col(l)) (:82-84)
    sync();
    const double sn = (double)sample_num, sn1 = (double)(sample_num + 1);
    for (int e = tid; e < U * K; e += nthr) {
        const int u = e % U, k = e / U;
        double qn = (sn * (Q[e] + p[u + (size_t)U * perm[k]])) / sn1;  // (:87-92)
        Q[e] = qn;
        logQ[e] = log(qn);
    }
    sync();
}

__device__ inline void stephens_online_block(int U, int K, const int *__restrict__ wt, double *Q, double *logQ,
                                             const double *p, int sample_num, double *cost, int *perm,
                                             void *assign_ws, bool fixed = false) {
    stephens_online_group(U, K, wt, Q, logQ, p, sample_num, cost, perm, assign_ws, fixed, (int)threadIdx.x, (int)blockDim.x,
                          [] { __syncthreads(); });
}

}  // namespace bmm
