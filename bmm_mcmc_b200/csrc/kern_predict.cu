// Posterior predictive distribution of a fitted Bernoulli mixture (the reference's TODO list, /root/reference/TODO:6,
// "Implement predictive distribution").  From the S kept draws (theta^(s), pi^(s)) an uncollapsed sampler
// returns (full_gibbs.cpp:233-248):
//     p(x* | data) ~= 1/S sum_s sum_k pi_k^(s) prod_d theta_kd^(s)^x*_d (1 - theta_kd^(s))^(1 - x*_d)
// and the averaged responsibilities r_k(x*) = 1/S sum_s pi_k L_k / sum_l pi_l L_l.  New rows are bit-packed like the data;
// one thread per new row walks the draws with a running log-sum-exp in fp64; the per-draw log tables are built once.
#include "kernels.h"

namespace bmm {
namespace {

// tab[s][k][0..P) = log theta_kd, [P..2P) = log(1 - theta_kd), [2P] = log pi_k   (theta: K x P x S, k + K d + K P s; pi: S x K cm)
__global__ void predict_tables_kernel(int K, int P, int S, const double *__restrict__ theta, const double *__restrict__ pi,
                                      double *__restrict__ tab) {
    const size_t n = (size_t)S * K * (2 * P + 1), row = 2 * P + 1;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const size_t sk = e / row;
        const int c = (int)(e % row), s = (int)(sk / K), k = (int)(sk % K);
        double v;
        if (c == 2 * P) v = log(pi[s + (size_t)S * k]);
        else {
            const int d = c < P ? c : c - P;
            const double th = theta[k + (size_t)K * d + (size_t)K * P * s];
            v = c < P ? log(th) : log(1.0 - th);
        }
        tab[e] = v;
    }
}

__global__ void __launch_bounds__(128) predict_kernel(int M, int P, int W, int K, int S, const uint32_t *__restrict__ xbits,
                                                      const double *__restrict__ tab, double *__restrict__ logpred,
                                                      double *__restrict__ member) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const uint32_t *xb = xbits + (size_t)m * W;
    const size_t row = 2 * P + 1;
    double mx = -INFINITY, acc = 0.0;     // running log-sum-exp over (s, k)
    for (int s = 0; s < S; ++s) {
        double lmx = -INFINITY, lsum = 0.0;
        for (int pass = 0; pass < (member ? 2 : 1); ++pass) {
            for (int k = 0; k < K; ++k) {
                const double *t = tab + ((size_t)s * K + k) * row;
                double ll = t[2 * P];
                for (int d = 0; d < P; ++d) ll += ((xb[d >> 5] >> (d & 31)) & 1u) ? t[d] : t[P + d];
                if (pass == 0) {
                    if (ll > lmx) { lsum = lsum * exp(lmx - ll) + 1.0; lmx = ll; }
                    else if (ll > -INFINITY) lsum += exp(ll - lmx);
                } else if (lsum > 0.0) {
                    member[m + (size_t)M * k] += exp(ll - lmx) / lsum / S;     // responsibility of cluster k under draw s
                }
            }
        }
        if (lsum > 0.0) {
            const double lse = lmx + log(lsum);                              // log p(x* | draw s)
            if (lse > mx) { acc = acc * exp(mx - lse) + 1.0; mx = lse; }
            else acc += exp(lse - mx);
        }
    }
    logpred[m] = acc > 0.0 ? mx + log(acc) - log((double)S) : -INFINITY;
}

}  // namespace

cudaError_t launch_predict(int M, int P, int W, int K, int S, const uint32_t *xbits, const double *theta, const double *pi,
                           double *tab, double *logpred, double *member, cudaStream_t st) {
    const size_t n = (size_t)S * K * (2 * P + 1);
    predict_tables_kernel<<<(unsigned)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096), 256, 0, st>>>(K, P, S, theta, pi, tab);
    g_launches++;
    predict_kernel<<<(M + 127) / 128, 128, 0, st>>>(M, P, W, K, S, xbits, tab, logpred, member);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace bmm
