// TEST INFRASTRUCTURE -- CPU oracle, never linked into or called by the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.
//
// Line-faithful CPU restatement (plain C++17, no Rcpp/Armadillo) of bmm-mcmc's hot path:
//   gibbs_cpp                 /root/reference/src/full_gibbs.cpp:32-249
//   gibbs_stickbreaking_cpp   /root/reference/src/stickbreaking.cpp:10-255
//   collapsed_gibbs_cpp       /root/reference/src/collapsed_gibbs.cpp:24-244
//   collapsed_gibbs_dp_cpp    /root/reference/src/collapsed_gibbs_dp.cpp:27-300
//   my_stephens_batch/online  /root/reference/src/stephens.cpp:6-94
//   my_lpsolve                /root/reference/src/my_lpsolve.cpp:6-31  (+ lp_solve itself, loaded
//                             from oracle/_ref/liblpsolve_ref.so = the reference's own C code)
//   update_alpha              /root/reference/src/utils.cpp:6-14
//   rdirichlet_cpp            /root/reference/src/full_gibbs.cpp:10-27
// It keeps the reference's algorithmic costs (per-point log recomputation, member-list
// rescans and vector copies, one-hot scans, 100 fixed batch iterations) and its quirks
// (SURVEY.md Appendix D), so it doubles as the CPU baseline.
//
// PARITY PINNING: the reference has no tests.  Pinned: unif_rand/rbinom core (bundled data
// regenerated bit-exactly), the assignment step (reference lp_solve compiled here).  Everything
// else -- the sampler restatements themselves, rmultinom/sample composition, Armadillo
// summation order, the Gamma/Beta generators -- is "parity unpinned": there is no R here to
// run the reference's samplers.  All matrices use R's column-major layout.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <queue>
#include <string>
#include <vector>
#include <dlfcn.h>

#include "rrng.h"

using oracle::RRng;

namespace {

typedef void (*lp_transbig_fn)(int, int, double *, double *);
lp_transbig_fn g_lp = nullptr;
bool g_lp_tried = false;

lp_transbig_fn load_lp() {
    if (g_lp_tried) return g_lp;
    g_lp_tried = true;
    Dl_info info;
    std::string dir = ".";
    if (dladdr((void *)&load_lp, &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        size_t s = p.rfind('/');
        if (s != std::string::npos) dir = p.substr(0, s);
    }
    std::string path = dir + "/_ref/liblpsolve_ref.so";
    void *h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (h) g_lp = (lp_transbig_fn)dlsym(h, "lp_transbig_edit");
    return g_lp;
}

// Exact assignment (Jonker-Volgenant style shortest augmenting path), the "port" stand-in
// when the reference lp_solve library is not present.  cost is K x K col-major; out row->col.
void hungarian(int K, const double *cost_cm, int *row_to_col) {
    const double INF = 1e300;
    std::vector<double> u(K + 1, 0.0), v(K + 1, 0.0), minv(K + 1);
    std::vector<int> p(K + 1, 0), way(K + 1, 0);
    std::vector<char> used(K + 1);
    for (int i = 1; i <= K; ++i) {
        p[0] = i;
        int j0 = 0;
        std::fill(minv.begin(), minv.end(), INF);
        std::fill(used.begin(), used.end(), 0);
        do {
            used[j0] = 1;
            int i0 = p[j0], j1 = 0;
            double delta = INF;
            for (int j = 1; j <= K; ++j)
                if (!used[j]) {
                    double cur = cost_cm[(i0 - 1) + (size_t)K * (j - 1)] - u[i0] - v[j];
                    if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
                    if (minv[j] < delta) { delta = minv[j]; j1 = j; }
                }
            for (int j = 0; j <= K; ++j)
                if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
                else minv[j] -= delta;
            j0 = j1;
        } while (p[j0] != 0);
        do { int j1 = way[j0]; p[j0] = p[j1]; j0 = j1; } while (j0);
    }
    for (int j = 1; j <= K; ++j) row_to_col[p[j] - 1] = j - 1;
}

// my_lpsolve (my_lpsolve.cpp:6-31): x is K x K col-major; sol is K x K col-major ints.
// use_ref=1 -> the reference's lp_solve; 0 -> Hungarian port.  Returns 0 ok, 1 ref missing.
int my_lpsolve(int K, const double *x, int *sol, int use_ref) {
    if (use_ref) {
        lp_transbig_fn f = load_lp();
        if (!f) return 1;
        std::vector<double> objective(1 + (size_t)K * K), solution((size_t)K * K);
        objective[0] = 0;
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) {
                objective[(size_t)i * K + j + 1] = x[i + (size_t)K * j];
                solution[(size_t)i * K + j] = 0;
            }
        f(K, K, objective.data(), solution.data());
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) sol[i + (size_t)K * j] = (int)solution[(size_t)i * K + j];  // truncation :26
        return 0;
    }
    std::vector<int> r2c(K);
    hungarian(K, x, r2c.data());
    std::fill(sol, sol + (size_t)K * K, 0);
    for (int i = 0; i < K; ++i) sol[i + (size_t)K * r2c[i]] = 1;
    return 0;
}

int index_max_col(int K, const int *sol, int col) {  // arma::index_max: first maximum
    int best = 0;
    for (int r = 1; r < K; ++r)
        if (sol[r + (size_t)K * col] > sol[best + (size_t)K * col]) best = r;
    return best;
}

// Correctness-fixed Stephens mode (NOT the reference's behaviour, off by default):
//   * threshold = 1e-6 instead of 10 ^ (-6) == -16, criterion on the total cost over the M slices (:24,58-59)
//   * the permutation used to re-order columns is the inverse of index_max(solution.col(.)), which is what
//     the no-op `perm <- sort_index(perm)` was meant to compute (:56,85)
//   * online cost with log p, like the batch cost (:79)
//   * Q' = (j Q + p_reordered) / (j + 1), a running mean (:92)
// The permutation handed back to the samplers (sample label -> reference label) is unchanged.
static int g_stephens_fixed = 0;
static int g_sample_stable = 0;

// my_stephens_batch (stephens.cpp:6-64).  p: N x K x M cube (by value in the reference).
int stephens_batch(int N, int K, int M, const double *p_in, double *q, int use_ref, int *perm_out) {
    std::vector<double> p(p_in, p_in + (size_t)N * K * M);
    std::vector<double> cost((size_t)K * K), sub((size_t)N * K), logq(N);
    std::vector<int> perm((size_t)M * K), solution((size_t)K * K);
    for (int k = 0; k < K; ++k)
        for (int t = 0; t < M; ++t) perm[t + (size_t)M * k] = k;
    // threshold = 10^(-6) is integer XOR = -16 (:24): the loop always runs maxiter = 100 times.
    double previous = -99, current, criterion = 99, threshold = g_stephens_fixed ? 1e-6 : (double)(10 ^ (-6));
    std::vector<int> a(K);
    int maxiter = 100, t = 0;
    const double min_prob = 0.000001;
    for (auto &v : p) if (v == 0.0) v = min_prob;  // p.replace(0, min_prob) :31
    while ((criterion > threshold) && (t < maxiter)) {
        t++;
        double total = 0;
        std::fill(q, q + (size_t)N * K, 0.0);
        for (int k = 0; k < K; ++k)
            for (int it = 0; it < M; ++it) {
                const double *src = &p[(size_t)N * K * it + (size_t)N * perm[it + (size_t)M * k]];
                for (int i = 0; i < N; ++i) q[i + (size_t)N * k] += src[i];
            }
        for (size_t e = 0; e < (size_t)N * K; ++e) q[e] /= M;
        for (int it = 0; it < M; ++it) {
            const double *ps = &p[(size_t)N * K * it];
            for (int k = 0; k < K; ++k) {
                for (size_t e = 0; e < (size_t)N * K; ++e) sub[e] = std::log(ps[e]);  // recomputed K times :49
                for (int i = 0; i < N; ++i) logq[i] = std::log(q[i + (size_t)N * k]);
                for (int l = 0; l < K; ++l) {
                    // arma::sum(expr, 0): two running sums over even / odd rows (op_sum::apply_noalias_proxy)
                    double acc1 = 0, acc2 = 0;
                    int i = 0;
                    for (; i + 1 < N; i += 2) {
                        acc1 += ps[i + (size_t)N * l] * (sub[i + (size_t)N * l] - logq[i]);
                        acc2 += ps[i + 1 + (size_t)N * l] * (sub[i + 1 + (size_t)N * l] - logq[i + 1]);
                    }
                    if (i < N) acc1 += ps[i + (size_t)N * l] * (sub[i + (size_t)N * l] - logq[i]);
                    cost[k + (size_t)K * l] = acc1 + acc2;
                }
            }
            if (my_lpsolve(K, cost.data(), solution.data(), use_ref)) return 1;
            for (int k = 0; k < K; ++k) perm[it + (size_t)M * k] = index_max_col(K, solution.data(), k);
            // perm.row(iter) <- sort_index(...) (:56) is `perm.row(iter) < -sort_index(...)`: a no-op.
            if (g_stephens_fixed) {
                for (int k = 0; k < K; ++k) a[k] = perm[it + (size_t)M * k];
                for (int l = 0; l < K; ++l) perm[it + (size_t)M * a[l]] = l;
                for (size_t e = 0; e < (size_t)K * K; ++e) total += cost[e] * solution[e];
            }
        }
        current = 0;
        for (size_t e = 0; e < (size_t)K * K; ++e) current += cost[e] * solution[e];
        if (g_stephens_fixed) current = total;
        criterion = std::fabs(previous - current);
        previous = current;
    }
    if (perm_out) std::copy(perm.begin(), perm.end(), perm_out);
    return 0;
}

// my_stephens_online (stephens.cpp:66-94).
int stephens_online(int N, int K, const double *q, const double *p, int sample_num, int *perm,
                    double *q_new, int use_ref, double *cost_out) {
    std::vector<double> cost((size_t)K * K), logq(N);
    std::vector<int> solution((size_t)K * K);
    for (int k = 0; k < K; ++k) {
        for (int i = 0; i < N; ++i) logq[i] = std::log(q[i + (size_t)N * k]);
        for (int l = 0; l < K; ++l) {
            double acc[2] = {0, 0};  // arma::sum(expr, 0): even / odd rows accumulate separately
            for (int i = 0; i < N; ++i) {
                double pv = p[i + (size_t)N * l];
                if (g_stephens_fixed) acc[i & 1] += pv > 0 ? pv * (std::log(pv) - logq[i]) : 0.0;
                else acc[i & 1] += pv * (pv - logq[i]);  // p, not log p (:79)
            }
            cost[k + (size_t)K * l] = acc[0] + acc[1];
        }
    }
    if (cost_out) std::copy(cost.begin(), cost.end(), cost_out);
    if (my_lpsolve(K, cost.data(), solution.data(), use_ref)) return 1;
    for (int k = 0; k < K; ++k) perm[k] = index_max_col(K, solution.data(), k);
    // perm <- sort_index(perm) (:85): no-op.
    if (g_stephens_fixed) {
        std::vector<int> inv(K);
        for (int l = 0; l < K; ++l) inv[perm[l]] = l;
        for (int k = 0; k < K; ++k)
            for (int i = 0; i < N; ++i)
                q_new[i + (size_t)N * k] = (sample_num * q[i + (size_t)N * k] + p[i + (size_t)N * inv[k]]) / (double)(sample_num + 1);
        return 0;
    }
    for (int k = 0; k < K; ++k)
        for (int i = 0; i < N; ++i)
            q_new[i + (size_t)N * k] =
                (sample_num * (q[i + (size_t)N * k] + p[i + (size_t)N * perm[k]])) / (double)(sample_num + 1);  // :92
    return 0;
}

// update_alpha (utils.cpp:6-14)
double update_alpha(RRng &R, double alpha_old, double a, double b, int N, int K) {
    double b_eps = b - std::log(R.rbeta(alpha_old + 1, N));
    double pi1 = a + K - 1, pi2 = N * b_eps, pi = pi1 / (pi1 + pi2);
    double g1 = R.rgamma(a + K, 1 / b_eps);
    double g2 = R.rgamma(a + K - 1, 1 / b_eps);
    return pi * g1 + (1 - pi) * g2;
}

struct Relabel {  // shared relabelling state of the four samplers
    int N, K, burnin, burnrelabel, use_ref;
    std::vector<double> probs_out, probs_sample, Q, Qn;
    std::vector<int> perm_sample;
    Relabel(int N_, int K_, int burnin_, int burnrelabel_, int use_ref_)
        : N(N_), K(K_), burnin(burnin_), burnrelabel(burnrelabel_), use_ref(use_ref_),
          probs_out((size_t)N_ * K_ * std::max(burnrelabel_, 0), 0.0), probs_sample((size_t)N_ * K_, 0.0),
          Q((size_t)N_ * K_, 0.0), Qn((size_t)N_ * K_, 0.0), perm_sample(K_, 0) {}
    inline void stash(int j, int i, int k, double v) {
        if (j < burnin && j >= (burnin - burnrelabel)) probs_out[i + (size_t)N * k + (size_t)N * K * (j - burnin + burnrelabel)] = v;
        else if (j >= burnin) probs_sample[i + (size_t)N * k] = v;
    }
    // returns 1 if permutations were produced this sweep, <0 on error
    int after_sweep(int j) {
        if (j == burnin - 1) {
            if (stephens_batch(N, K, burnrelabel, probs_out.data(), Q.data(), use_ref, nullptr)) return -1;
            return 0;
        } else if (j >= burnin) {
            if (stephens_online(N, K, Q.data(), probs_sample.data(), j, perm_sample.data(), Qn.data(), use_ref, nullptr)) return -1;
            Q.swap(Qn);
            return 1;
        }
        return 0;
    }
};

}  // namespace

extern "C" {

struct oracle_out {
    double *pi;         // nsamples x K (cm), full / stick-breaking only
    double *alpha;      // nsamples
    int *permutations;  // (nsamples-burnin) x K (cm)
    int *z;             // nsamples x N (cm), original labels 1..K (row 0: initial state or 0)
    int *z_rel;         // nsamples x N (cm), relabelled (rows >= burnin)
    double *theta;      // K x P x nsamples
    double *theta_rel;  // K x P x nsamples (slices >= burnin)
    double *u_rec;      // nsamples x N x slots: uniforms consumed by each z draw (-1 = unused)
    int u_slots;        // slots per draw in u_rec
    double *probs;      // optional nsamples x (N x K cm): conditional probabilities per sweep
    double *loglik;     // optional nsamples x (N x K cm): Bernoulli log-likelihood (full / SB)
    double *Q_final;    // optional N x K
    int *Kactive;       // optional nsamples: DP cluster count at sweep end
};

int oracle_has_lpsolve_ref() { return load_lp() != nullptr; }

void oracle_set_stephens_fixed(int on) { g_stephens_fixed = on; }

// Tie order of RcppArmadillo::sample()'s descending sort in the DP sampler: 0 (default) = std::sort like the
// reference build, 1 = stable for every candidate count (see rrng.h::sample1).
void oracle_set_sample_stable(int on) { g_sample_stable = on; }

int oracle_assign(int K, const double *cost_cm, int *sol_cm, int use_ref) { return my_lpsolve(K, cost_cm, sol_cm, use_ref); }

int oracle_stephens_batch(int N, int K, int M, const double *p, double *q, int use_ref, int *perm_out) {
    return stephens_batch(N, K, M, p, q, use_ref, perm_out);
}

int oracle_stephens_online(int N, int K, const double *q, const double *p, int sample_num, int *perm,
                           double *q_new, int use_ref, double *cost_out) {
    return stephens_online(N, K, q, p, sample_num, perm, q_new, use_ref, cost_out);
}

void oracle_unif_rand(unsigned seed, int n, double *out) {
    RRng R; R.set_seed(seed);
    for (int i = 0; i < n; ++i) out[i] = R.unif_rand();
}

// rbinom(n, 1, p) stream, used by the fixture known-answer test
void oracle_rbinom1(unsigned seed, int nblocks, const int *counts, const double *ps, int *out) {
    RRng R; R.set_seed(seed);
    size_t o = 0;
    for (int b = 0; b < nblocks; ++b)
        for (int i = 0; i < counts[b]; ++i) out[o++] = R.rbinom1(ps[b]);
}

// rmultinom(1, prob, K) from a given uniform list (replay check of the draw rule itself)
int oracle_rmultinom1_seeded(unsigned seed, const double *prob, int K, int *rN, double *u_used, int *n_used) {
    RRng R; R.set_seed(seed);
    R.rec = u_used; R.rec_cap = K; R.rec_n = 0;
    int rc = R.rmultinom1(prob, K, rN);
    *n_used = R.rec_n;
    return rc;
}

// rdirichlet_cpp (full_gibbs.cpp:10-27)
void oracle_rdirichlet(unsigned seed, int K, const double *alpha_m, double *out) {
    RRng R; R.set_seed(seed);
    double sum_term = 0;
    for (int j = 0; j < K; ++j) { out[j] = R.rgamma(alpha_m[j], 1.0); sum_term += out[j]; }
    for (int j = 0; j < K; ++j) out[j] /= sum_term;
}

void oracle_rgamma(unsigned seed, int n, double shape, double scale, double *out) {
    RRng R; R.set_seed(seed);
    for (int i = 0; i < n; ++i) out[i] = R.rgamma(shape, scale);
}
void oracle_rbeta(unsigned seed, int n, double a, double b, double *out) {
    RRng R; R.set_seed(seed);
    for (int i = 0; i < n; ++i) out[i] = R.rbeta(a, b);
}

// ---------------------------------------------------------------------------------------------
// Shared z-sweep of the uncollapsed samplers (full_gibbs.cpp:87-157 == stickbreaking.cpp:70-140)
// ---------------------------------------------------------------------------------------------
static int uncollapsed_zsweep(RRng &R, int N, int P, int K, const int *df, const double *thisTheta,
                              const double *pi_prev, int j, int nsamples, bool relabel, Relabel &rl,
                              double *onehot /*N x K*/, oracle_out *o, int stabilise) {
    std::vector<double> s(K), ll(K);
    std::vector<int> this_z(K);
    for (int i = 0; i < N; ++i) {
        double cum_probs = 0;
        for (int k = 0; k < K; ++k) {
            double loglh = 0;
            for (int d = 0; d < P; ++d) {
                int x = df[i + (size_t)N * d];
                loglh += x * std::log(thisTheta[k + (size_t)K * d]) + (1 - x) * std::log(1 - thisTheta[k + (size_t)K * d]);
            }
            ll[k] = loglh;
            if (o->loglik) o->loglik[(size_t)j * N * K + i + (size_t)N * k] = loglh;
        }
        double mx = 0;
        if (stabilise) {  // NOT in the reference (quirk 13); only for shapes where it yields NaN
            mx = -INFINITY;
            for (int k = 0; k < K; ++k) mx = std::max(mx, std::log(pi_prev[k]) + ll[k]);
        }
        for (int k = 0; k < K; ++k) {
            double dummy = stabilise ? std::exp(std::log(pi_prev[k]) + ll[k] - mx) : std::exp(std::log(pi_prev[k]) + ll[k]);
            s[k] = dummy;
            cum_probs += dummy;
        }
        for (int p = 0; p < K; ++p) s[p] /= cum_probs;
        if (o->probs) for (int k = 0; k < K; ++k) o->probs[(size_t)j * N * K + i + (size_t)N * k] = s[k];
        if (o->u_rec) {
            R.rec = o->u_rec + ((size_t)j * N + i) * o->u_slots; R.rec_cap = o->u_slots; R.rec_n = 0;
        }
        int rc = R.rmultinom1(s.data(), K, this_z.data());
        R.rec = nullptr;
        if (rc) return -2;
        for (int k = 0; k < K; ++k) onehot[i + (size_t)N * k] = this_z[k];
        for (int k = 0; k < K; ++k)
            if (this_z[k] == 1) o->z[j + (size_t)nsamples * i] = k + 1;
        if (relabel) for (int k = 0; k < K; ++k) rl.stash(j, i, k, s[k]);
    }
    return 0;
}

// gibbs_cpp (full_gibbs.cpp:32-249).  df: N x P int (cm).  Histories are full length (nsamples);
// the R return block's tail slicing (:233-248) is done by the caller.
int oracle_gibbs_full(const int *df, int N, int P, const double *initialPi, const double *initialTheta,
                      int nsamples, int K, double alpha, double beta, double gamma, double a, double b,
                      int burnin, int relabel, int burnrelabel, unsigned seed, int use_ref, int stabilise,
                      oracle_out *o) {
    RRng R; R.set_seed(seed);
    Relabel rl(N, K, burnin, relabel ? burnrelabel : 0, use_ref);
    std::vector<double> onehot((size_t)N * K);
    std::vector<double> thisTheta((size_t)K * P);
    for (int k = 0; k < K; ++k) o->pi[0 + (size_t)nsamples * k] = initialPi[k];
    std::copy(initialTheta, initialTheta + (size_t)K * P, o->theta);
    if (alpha == 0) o->alpha[0] = 1; else for (int j = 0; j < nsamples; ++j) o->alpha[j] = alpha;
    std::vector<double> pi_prev(K), dirich(K);
    for (int j = 1; j < nsamples; ++j) {
        std::copy(o->theta + (size_t)K * P * (j - 1), o->theta + (size_t)K * P * j, thisTheta.begin());
        for (int k = 0; k < K; ++k) pi_prev[k] = o->pi[(j - 1) + (size_t)nsamples * k];
        int rc = uncollapsed_zsweep(R, N, P, K, df, thisTheta.data(), pi_prev.data(), j, nsamples, relabel, rl, onehot.data(), o, stabilise);
        if (rc) return rc;
        bool have_perm = false;
        if (relabel) {
            int r = rl.after_sweep(j);
            if (r < 0) return -3;
            if (r == 1) {
                have_perm = true;
                for (int k = 0; k < K; ++k) o->permutations[(j - burnin) + (size_t)(nsamples - burnin) * k] = rl.perm_sample[k];
                for (int i = 0; i < N; ++i) o->z_rel[j + (size_t)nsamples * i] = rl.perm_sample[o->z[j + (size_t)nsamples * i] - 1] + 1;
            }
        }
        // sufficient statistics from the one-hot matrix (:182-200)
        std::vector<int> ck(K, 0), Vkd((size_t)K * P, 0);
        for (int k = 0; k < K; ++k)
            for (int i = 0; i < N; ++i) {
                int Znk = (int)onehot[i + (size_t)N * k];
                ck[k] += Znk;
                for (int d = 0; d < P; ++d) Vkd[k + (size_t)K * d] += Znk * df[i + (size_t)N * d];
            }
        double sum_term = 0;  // rdirichlet_cpp (:10-27)
        for (int k = 0; k < K; ++k) { dirich[k] = R.rgamma((o->alpha[j - 1] / K) + ck[k], 1.0); sum_term += dirich[k]; }
        for (int k = 0; k < K; ++k) o->pi[j + (size_t)nsamples * k] = dirich[k] / sum_term;
        for (int k = 0; k < K; ++k)
            for (int d = 0; d < P; ++d) {
                double th = R.rbeta(beta + Vkd[k + (size_t)K * d], gamma + ck[k] - Vkd[k + (size_t)K * d]);
                o->theta[k + (size_t)K * d + (size_t)K * P * j] = th;
                if (relabel && j >= burnin && have_perm) o->theta_rel[rl.perm_sample[k] + (size_t)K * d + (size_t)K * P * j] = th;
            }
        if (alpha == 0) o->alpha[j] = update_alpha(R, o->alpha[j - 1], a, b, N, K);
    }
    if (o->Q_final && relabel) std::copy(rl.Q.begin(), rl.Q.end(), o->Q_final);
    return 0;
}

// gibbs_stickbreaking_cpp (stickbreaking.cpp:10-255)
int oracle_gibbs_stickbreaking(const int *df, int N, int P, const double *initialPi, const double *initialTheta,
                               int nsamples, int maxK, double alpha, double beta, double gamma, double a, double b,
                               int burnin, int relabel, int burnrelabel, unsigned seed, int use_ref, int stabilise,
                               oracle_out *o) {
    RRng R; R.set_seed(seed);
    const int K = maxK;
    Relabel rl(N, K, burnin, relabel ? burnrelabel : 0, use_ref);
    std::vector<double> onehot((size_t)N * K), thisTheta((size_t)K * P), pi_prev(K), v(K), this_pi(K);
    const double viable_threshold = 0.01;
    int K_viable = maxK;
    for (int k = 0; k < K; ++k) o->pi[0 + (size_t)nsamples * k] = initialPi[k];
    std::copy(initialTheta, initialTheta + (size_t)K * P, o->theta);
    if (alpha == 0) o->alpha[0] = 1; else for (int j = 0; j < nsamples; ++j) o->alpha[j] = alpha;
    for (int j = 1; j < nsamples; ++j) {
        std::copy(o->theta + (size_t)K * P * (j - 1), o->theta + (size_t)K * P * j, thisTheta.begin());
        for (int k = 0; k < K; ++k) pi_prev[k] = o->pi[(j - 1) + (size_t)nsamples * k];
        int rc = uncollapsed_zsweep(R, N, P, K, df, thisTheta.data(), pi_prev.data(), j, nsamples, relabel, rl, onehot.data(), o, stabilise);
        if (rc) return rc;
        bool have_perm = false;
        if (relabel) {
            int r = rl.after_sweep(j);
            if (r < 0) return -3;
            if (r == 1) {
                have_perm = true;
                for (int k = 0; k < K; ++k) o->permutations[(j - burnin) + (size_t)(nsamples - burnin) * k] = rl.perm_sample[k];
                for (int i = 0; i < N; ++i) o->z_rel[j + (size_t)nsamples * i] = rl.perm_sample[o->z[j + (size_t)nsamples * i] - 1] + 1;
            }
        }
        std::vector<int> ck(K, 0), Vkd((size_t)K * P, 0);
        int num_previous_clusters = 0;
        for (int k = K - 1; k >= 0; --k) {  // reverse order (:170)
            for (int i = 0; i < N; ++i) {
                int Znk = (int)onehot[i + (size_t)N * k];
                ck[k] += Znk;
                for (int d = 0; d < P; ++d) Vkd[k + (size_t)K * d] += Znk * df[i + (size_t)N * d];
            }
            double beta1 = 1 + ck[k], beta2 = o->alpha[j - 1] + num_previous_clusters;
            v[k] = R.rbeta(beta1, beta2);
            num_previous_clusters += ck[k];
        }
        v[K - 1] = 1;
        K_viable = 0;
        this_pi[0] = v[0];
        if (this_pi[0] > viable_threshold) K_viable++;
        double cumprod = 1 - v[0];
        for (int k = 1; k < K; k++) {
            this_pi[k] = cumprod * v[k];
            if (this_pi[k] > viable_threshold) K_viable++;
            cumprod *= (1 - v[k]);
        }
        for (int k = 0; k < K; ++k) o->pi[j + (size_t)nsamples * k] = this_pi[k];
        for (int k = 0; k < K; ++k)
            for (int d = 0; d < P; ++d) {
                double th = R.rbeta(beta + Vkd[k + (size_t)K * d], gamma + ck[k] - Vkd[k + (size_t)K * d]);
                o->theta[k + (size_t)K * d + (size_t)K * P * j] = th;
                if (relabel && j >= burnin && have_perm) o->theta_rel[rl.perm_sample[k] + (size_t)K * d + (size_t)K * P * j] = th;
            }
        if (alpha == 0) o->alpha[j] = update_alpha(R, o->alpha[j - 1], a, b, N, K_viable);
    }
    if (o->Q_final && relabel) std::copy(rl.Q.begin(), rl.Q.end(), o->Q_final);
    return 0;
}

// collapsed_gibbs_cpp (collapsed_gibbs.cpp:24-244)
int oracle_gibbs_collapsed(const int *df, int N, int P, const int *initialK, int nsamples, int K, double alpha,
                           double beta, double gamma, double a, double b, int burnin, int relabel, int burnrelabel,
                           unsigned seed, int use_ref, oracle_out *o) {
    RRng R; R.set_seed(seed);
    Relabel rl(N, K, burnin, relabel ? burnrelabel : 0, use_ref);
    for (int i = 0; i < N; ++i) o->z[0 + (size_t)nsamples * i] = initialK[i];
    if (alpha == 0) o->alpha[0] = 1; else for (int j = 0; j < nsamples; ++j) o->alpha[j] = alpha;
    std::vector<std::vector<int>> clusters(K);
    for (int i = 0; i < N; ++i) clusters[initialK[i] - 1].push_back(i);
    std::vector<int> Ck, this_z(K);
    std::vector<double> probs(K);
    for (int j = 1; j < nsamples; ++j) {
        for (int i = 0; i < N; ++i) {
            int curr_cluster = o->z[(j - 1) + (size_t)nsamples * i] - 1;
            auto &cc = clusters[curr_cluster];
            cc.erase(std::remove(cc.begin(), cc.end(), i), cc.end());
            double probs_sum = 0;
            for (int k = 0; k < K; ++k) {
                Ck = clusters[k];  // vector copy (:101)
                int Nk = (int)Ck.size();
                double dummy;
                if (Nk > 0) {
                    double LHS = std::log(Nk + (o->alpha[j - 1] / K)) - std::log(N - 1 + o->alpha[j - 1]);
                    double logLH = 0;
                    for (int d = 0; d < P; ++d) {
                        int sum_xd = 0;
                        for (int c : Ck) sum_xd += df[c + (size_t)N * d];  // rescan (:111-114)
                        int xnd = df[i + (size_t)N * d];
                        double left = xnd * std::log(beta + sum_xd);
                        double right = (1 - xnd) * std::log(gamma + Nk - sum_xd);
                        double denom = std::log(beta + gamma + Nk);
                        logLH += left + right - denom;
                    }
                    dummy = std::exp(LHS + logLH);
                } else {
                    dummy = 0;  // empty cluster can never be re-occupied (quirk 7)
                }
                probs_sum += dummy;
                probs[k] = dummy;
            }
            for (int k = 0; k < K; ++k) probs[k] /= probs_sum;
            if (o->probs) for (int k = 0; k < K; ++k) o->probs[(size_t)j * N * K + i + (size_t)N * k] = probs[k];
            if (o->u_rec) { R.rec = o->u_rec + ((size_t)j * N + i) * o->u_slots; R.rec_cap = o->u_slots; R.rec_n = 0; }
            int rc = R.rmultinom1(probs.data(), K, this_z.data());
            R.rec = nullptr;
            if (rc) return -2;
            if (relabel) for (int k = 0; k < K; ++k) rl.stash(j, i, k, probs[k]);
            for (int k = 0; k < K; ++k)
                if (this_z[k] == 1) { o->z[j + (size_t)nsamples * i] = k + 1; clusters[k].push_back(i); }
        }
        bool have_perm = false;
        if (relabel) {
            int r = rl.after_sweep(j);
            if (r < 0) return -3;
            if (r == 1) {
                have_perm = true;
                for (int k = 0; k < K; ++k) o->permutations[(j - burnin) + (size_t)(nsamples - burnin) * k] = rl.perm_sample[k];
                for (int i = 0; i < N; ++i) o->z_rel[j + (size_t)nsamples * i] = rl.perm_sample[o->z[j + (size_t)nsamples * i] - 1] + 1;
            }
        }
        for (int k = 0; k < K; ++k) {  // theta point estimates (:205-219); NaN when empty (quirk 8)
            Ck = clusters[k];
            int Nk = (int)Ck.size();
            for (int d = 0; d < P; ++d) {
                int dsum = 0;
                for (int c : Ck) dsum += df[c + (size_t)N * d];
                double th = dsum / (double)Nk;
                o->theta[k + (size_t)K * d + (size_t)K * P * j] = th;
                if (relabel && j >= burnin && have_perm) o->theta_rel[rl.perm_sample[k] + (size_t)K * d + (size_t)K * P * j] = th;
            }
        }
        if (alpha == 0) o->alpha[j] = update_alpha(R, o->alpha[j - 1], a, b, N, K);
    }
    if (o->Q_final && relabel) std::copy(rl.Q.begin(), rl.Q.end(), o->Q_final);
    return 0;
}

// collapsed_gibbs_dp_cpp (collapsed_gibbs_dp.cpp:27-300).  Returns -4 for beta != gamma (:48-50),
// -5 when the free-label heap is empty (:166-168), -6 when the reference's state goes
// inconsistent after a truncation fallback (K != used_clusters.size(): undefined behaviour there).
int oracle_gibbs_dp(const int *df, int N, int P, int nsamples, double alpha, double beta, double gamma, double a,
                    double b, int burnin, int relabel, int burnrelabel, int maxK, unsigned seed, int use_ref,
                    oracle_out *o) {
    if (beta != gamma) return -4;
    RRng R; R.set_seed(seed);
    R.stable_ties = g_sample_stable != 0;
    int K = 0;
    Relabel rl(N, maxK, burnin, relabel ? burnrelabel : 0, use_ref);
    std::vector<std::vector<int>> clusters(maxK);
    std::vector<int> used_clusters;
    std::priority_queue<int, std::vector<int>, std::greater<int>> unused_clusters;
    for (int i = 0; i < N; ++i) unused_clusters.push(i);
    double RHS_newk = P * (std::log(beta) - std::log(beta + gamma));
    if (alpha == 0) o->alpha[0] = 1; else for (int j = 0; j < nsamples; ++j) o->alpha[j] = alpha;
    std::vector<int> Ck;
    std::vector<double> sp; std::vector<int> spp;
    for (int j = 1; j < nsamples; ++j) {
        if ((size_t)K != used_clusters.size()) return -6;
        double left_denom = std::log(N - 1 + o->alpha[j - 1]);
        double probs_newk = std::log(o->alpha[j - 1]) - left_denom + RHS_newk;
        for (int i = 0; i < N; ++i) {
            if (j > 1) {
                int curr_cluster = o->z[(j - 1) + (size_t)nsamples * i] - 1;
                auto &cc = clusters[curr_cluster];
                cc.erase(std::remove(cc.begin(), cc.end(), i), cc.end());
                if (cc.size() == 0) {
                    used_clusters.erase(std::remove(used_clusters.begin(), used_clusters.end(), curr_cluster), used_clusters.end());
                    unused_clusters.push(curr_cluster);
                    K--;
                }
            }
            if (K < 0 || (size_t)K != used_clusters.size()) return -6;
            std::vector<double> probs(K + 1), probs_norm(K + 1);
            std::vector<int> choices(K + 1);
            for (int k = 0; k < K; ++k) {
                Ck = clusters[used_clusters[k]];
                int Nk = (int)Ck.size();
                double LHS = std::log((double)Nk) - left_denom;
                double logLH = 0;
                double denom = std::log(beta + gamma + Nk);
                for (int d = 0; d < P; ++d) {
                    int sum_xd = 0;
                    for (int c : Ck) sum_xd += df[c + (size_t)N * d];
                    int xnd = df[i + (size_t)N * d];
                    double left = xnd * std::log(beta + sum_xd);
                    double right = (1 - xnd) * std::log(gamma + Nk - sum_xd);
                    logLH += left + right - denom;
                }
                probs[k] = LHS + logLH;
                choices[k] = used_clusters[k];
            }
            if (unused_clusters.size() == 0) return -5;
            int new_cluster = unused_clusters.top();
            choices[K] = new_cluster;
            probs[K] = probs_newk;
            double max_prob = *std::max_element(probs.begin(), probs.end());
            double sumprob = 0;
            for (int k = 0; k <= K; ++k) { double f = std::exp(probs[k] - max_prob); probs_norm[k] = f; sumprob += f; }
            for (int k = 0; k <= K; ++k) probs_norm[k] /= sumprob;
            if (relabel) for (int k = 0; k <= K; ++k) if (choices[k] < maxK) rl.stash(j, i, choices[k], probs_norm[k]);
            if (o->probs) for (int k = 0; k <= K; ++k) if (choices[k] < maxK) o->probs[(size_t)j * N * maxK + i + (size_t)N * choices[k]] = probs_norm[k];
            if (o->u_rec) { R.rec = o->u_rec + ((size_t)j * N + i) * o->u_slots; R.rec_cap = o->u_slots; R.rec_n = 0; }
            sp.resize(K + 1); spp.resize(K + 1);
            int idx = R.sample1(probs_norm.data(), K + 1, sp.data(), spp.data());
            R.rec = nullptr;
            int ret = choices[idx];
            if (ret == new_cluster) {
                if (K < (maxK - 1)) {
                    unused_clusters.pop();
                    used_clusters.push_back(new_cluster);
                    K++;
                } else {
                    int smallest_cluster = 0, smallest_size = N + 1;
                    for (int k = 0; k < K; ++k) {
                        int sz = (int)clusters[used_clusters[k]].size();
                        if (sz < smallest_size) { smallest_size = sz; smallest_cluster = k; }
                        ret = smallest_cluster;  // index, not label (quirk 9)
                    }
                }
            }
            if (ret < 0 || ret >= maxK) return -6;
            clusters[ret].push_back(i);
            o->z[j + (size_t)nsamples * i] = ret + 1;
            if (alpha == 0) o->alpha[j] = update_alpha(R, o->alpha[j - 1], a, b, N, K);  // per point (quirk 10)
        }
        if (o->Kactive) o->Kactive[j] = K;
        bool have_perm = false;
        if (relabel) {
            int r = rl.after_sweep(j);
            if (r < 0) return -3;
            if (r == 1) {
                have_perm = true;
                for (int k = 0; k < maxK; ++k) o->permutations[(j - burnin) + (size_t)(nsamples - burnin) * k] = rl.perm_sample[k];
                for (int i = 0; i < N; ++i) o->z_rel[j + (size_t)nsamples * i] = rl.perm_sample[o->z[j + (size_t)nsamples * i] - 1] + 1;
            }
        }
        for (int k : used_clusters) {
            Ck = clusters[k];
            int Nk = (int)Ck.size();
            for (int d = 0; d < P; ++d) {
                int dsum = 0;
                for (int c : Ck) dsum += df[c + (size_t)N * d];
                double th = dsum / (double)Nk;
                o->theta[k + (size_t)maxK * d + (size_t)maxK * P * j] = th;
                if (relabel && j >= burnin && have_perm) o->theta_rel[rl.perm_sample[k] + (size_t)maxK * d + (size_t)maxK * P * j] = th;
            }
        }
    }
    if (o->Q_final && relabel) std::copy(rl.Q.begin(), rl.Q.end(), o->Q_final);
    return 0;
}

}  // extern "C"
