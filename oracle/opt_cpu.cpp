// TEST / BENCH INFRASTRUCTURE -- not part of the product path.
// "Optimised CPU" baseline of the uncollapsed sampler with online relabelling ("optionally also report an
// optimised CPU ... so algorithmic and hardware speed-ups are separable"): the same ALGORITHM the chain-per-block GPU
// kernel runs -- rows de-duplicated into U patterns, log tables hoisted out of the observation loop, per-pattern
// conditional probabilities, one uniform per allocation, (pattern, label) histogram folded into the sufficient
// statistics, Marsaglia-Tsang Gamma / Beta draws, Stephens' online step on the U x K matrices with K! enumeration --
// written for one host core per chain (g++ -O3 -march=native), one std::thread per chain.  It follows
// /root/reference/src/full_gibbs.cpp:83-231 and stephens.cpp:66-94 in distribution, not in operation order; it is
// never compared bit for bit with anything, only timed.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct Rng {          // xoshiro256++
    uint64_t s[4];
    explicit Rng(uint64_t seed) {
        for (int i = 0; i < 4; ++i) { seed += 0x9E3779B97F4A7C15ull; uint64_t z = seed; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; s[i] = z ^ (z >> 31); }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t r = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double unif() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
    double normal() { const double u1 = unif(), u2 = unif(); return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2); }
    double gamma(double a) {
        double boost = 1.0;
        if (a < 1.0) { boost = std::exp(std::log(unif()) / a); a += 1.0; }
        const double d = a - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
        for (;;) {
            const double x = normal();
            double v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            if (std::log(unif()) < 0.5 * x * x + d - d * v + d * std::log(v)) return boost * d * v;
        }
    }
    double beta(double a, double b) { const double x = gamma(a), y = gamma(b); return x / (x + y); }
};

unsigned one_chain(const int32_t *X, int N, int P, int K, int nsamples, int burnin, uint64_t seed) {
    // patterns
    std::unordered_map<uint64_t, int> seen;
    std::vector<uint64_t> pat;
    std::vector<int> wt, rowid(N);
    for (int i = 0; i < N; ++i) {
        uint64_t b = 0;
        for (int d = 0; d < P; ++d) b |= (uint64_t)(X[i + (size_t)N * d] & 1) << d;
        auto it = seen.find(b);
        if (it == seen.end()) { it = seen.emplace(b, (int)pat.size()).first; pat.push_back(b); wt.push_back(0); }
        wt[it->second]++; rowid[i] = it->second;
    }
    const int U = (int)pat.size();
    Rng rng(seed);
    std::vector<double> theta((size_t)K * P), pi(K), w1((size_t)K * P), w0((size_t)K * P), prob((size_t)U * K), cum((size_t)U * K),
        Q((size_t)U * K, 1.0 / K), cost((size_t)K * K);
    std::vector<int> hist((size_t)U * K), ck(K), V((size_t)K * P), perm(K), best(K);
    std::vector<uint8_t> zh((size_t)nsamples * N);     // the allocation history, one byte per draw like the device keeps it
    for (auto &t : theta) t = rng.unif();
    { double s = 0; for (int k = 0; k < K; ++k) { pi[k] = std::exp(rng.unif()); s += pi[k]; } for (auto &p : pi) p /= s; }
    double alpha = 1.0;
    for (int j = 1; j < nsamples; ++j) {
        for (size_t e = 0; e < (size_t)K * P; ++e) { w1[e] = std::log(theta[e]); w0[e] = std::log(1.0 - theta[e]); }
        for (int u = 0; u < U; ++u) {          // conditional probabilities per pattern
            double s = 0;
            for (int k = 0; k < K; ++k) {
                double ll = std::log(pi[k]);
                for (int d = 0; d < P; ++d) ll += ((pat[u] >> d) & 1) ? w1[k + (size_t)K * d] : w0[k + (size_t)K * d];
                prob[(size_t)u * K + k] = std::exp(ll);
                s += prob[(size_t)u * K + k];
            }
            double c = 0;
            for (int k = 0; k < K; ++k) { prob[(size_t)u * K + k] /= s; c += prob[(size_t)u * K + k]; cum[(size_t)u * K + k] = c; }
        }
        std::fill(hist.begin(), hist.end(), 0);
        for (int i = 0; i < N; ++i) {          // allocations
            const double u01 = rng.unif();
            const double *c = &cum[(size_t)rowid[i] * K];
            int k = 0;
            while (k < K - 1 && u01 >= c[k]) ++k;
            zh[(size_t)j * N + i] = (uint8_t)(k + 1);
            hist[(size_t)rowid[i] * K + k]++;
        }
        if (j >= burnin) {                      // online relabelling on the patterns (weights = multiplicities)
            for (int k = 0; k < K; ++k)
                for (int l = 0; l < K; ++l) {
                    double c = 0;
                    for (int u = 0; u < U; ++u) { const double p = prob[(size_t)u * K + l]; c += wt[u] * p * (p - std::log(Q[(size_t)u * K + k])); }
                    cost[(size_t)k * K + l] = c;
                }
            for (int k = 0; k < K; ++k) perm[k] = k;
            double bc = 1e300;
            std::vector<int> cand(perm);
            do {
                double c = 0;
                for (int l = 0; l < K; ++l) c += cost[(size_t)cand[l] * K + l];
                if (c < bc) { bc = c; best = cand; }
            } while (K <= 8 && std::next_permutation(cand.begin(), cand.end()));
            for (int u = 0; u < U; ++u) {
                double tmp[64];
                for (int k = 0; k < K; ++k) tmp[k] = j * (Q[(size_t)u * K + k] + prob[(size_t)u * K + best[k]]) / (j + 1.0);
                for (int k = 0; k < K; ++k) Q[(size_t)u * K + k] = tmp[k];
            }
        }
        std::fill(ck.begin(), ck.end(), 0); std::fill(V.begin(), V.end(), 0);
        for (int u = 0; u < U; ++u)
            for (int k = 0; k < K; ++k) {
                const int h = hist[(size_t)u * K + k];
                ck[k] += h;
                for (int d = 0; d < P; ++d) V[k + (size_t)K * d] += h * (int)((pat[u] >> d) & 1);
            }
        double s = 0;
        for (int k = 0; k < K; ++k) { pi[k] = rng.gamma(alpha / K + ck[k]); s += pi[k]; }
        for (auto &p : pi) p /= s;
        for (int k = 0; k < K; ++k)
            for (int d = 0; d < P; ++d) theta[k + (size_t)K * d] = rng.beta(0.5 + V[k + (size_t)K * d], 0.5 + ck[k] - V[k + (size_t)K * d]);
        {   // Escobar-West alpha update with the reference's convex combination (utils.cpp:6-14)
            const double be = 1.0 - std::log(rng.beta(alpha + 1.0, N)), p1 = 1.0 + K - 1.0, p2 = N * be, w = p1 / (p1 + p2);
            alpha = w * rng.gamma(1.0 + K) / be + (1 - w) * rng.gamma(1.0 + K - 1.0) / be;
        }
    }
    unsigned chk = 0;
    for (size_t e = (size_t)N; e < zh.size(); e += 997) chk += zh[e];
    return chk;
}

// Collapsed finite-K sampler with maintained counts and the conditional in product form (what collapsed_prod_kernel runs):
// (N_k + alpha/K) prod_d (x ? beta + S_kd : gamma + N_k - S_kd) / (beta + gamma + N_k)^P, empty clusters skipped
// (collapsed_gibbs.cpp:104,131-133), O(K P) per update instead of the reference's member rescans.
unsigned one_chain_collapsed(const int32_t *X, int N, int P, int K, int nsamples, uint64_t seed) {
    Rng rng(seed);
    std::vector<uint8_t> x((size_t)N * P), z(N);
    for (int i = 0; i < N; ++i)
        for (int d = 0; d < P; ++d) x[(size_t)i * P + d] = (uint8_t)(X[i + (size_t)N * d] & 1);
    std::vector<int> Nk(K, 0), S((size_t)K * P, 0);
    for (int i = 0; i < N; ++i) {
        const int k = (int)(rng.unif() * K) % K;
        z[i] = (uint8_t)k; Nk[k]++;
        for (int d = 0; d < P; ++d) S[(size_t)k * P + d] += x[(size_t)i * P + d];
    }
    std::vector<uint8_t> zh((size_t)nsamples * N);
    const double alpha = 1.0, beta = 0.5, gamma = 0.5;
    double pr[64];
    for (int j = 1; j < nsamples; ++j) {
        for (int i = 0; i < N; ++i) {
            const uint8_t *xi = &x[(size_t)i * P];
            const int zo = z[i];
            Nk[zo]--;
            for (int d = 0; d < P; ++d) S[(size_t)zo * P + d] -= xi[d];
            double tot = 0;
            for (int k = 0; k < K; ++k) {
                double v = 0;
                if (Nk[k] > 0) {
                    v = Nk[k] + alpha / K;
                    const double den = beta + gamma + Nk[k];
                    for (int d = 0; d < P; ++d) v *= (xi[d] ? beta + S[(size_t)k * P + d] : gamma + Nk[k] - S[(size_t)k * P + d]) / den;
                }
                pr[k] = v; tot += v;
            }
            const double u = rng.unif() * tot;
            int k = 0;
            double c = pr[0];
            while (k < K - 1 && u >= c) c += pr[++k];
            z[i] = (uint8_t)k; Nk[k]++;
            for (int d = 0; d < P; ++d) S[(size_t)k * P + d] += xi[d];
            zh[(size_t)j * N + i] = (uint8_t)(k + 1);
        }
    }
    unsigned chk = 0;
    for (size_t e = (size_t)N; e < zh.size(); e += 997) chk += zh[e];
    return chk;
}

}  // namespace

// chains x (nsamples - 1) sweeps of gibbs_collapsed (fixed alpha = 1, no relabelling) on `threads` host threads; wall seconds
extern "C" double opt_cpu_collapsed_gibbs(const int32_t *X, int N, int P, int K, int nsamples, int chains, int threads) {
    if (K > 64 || threads < 1) return -1.0;
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([=] {
            volatile unsigned sink = 0;
            for (int c = t; c < chains; c += threads) sink = sink + one_chain_collapsed(X, N, P, K, nsamples, 5000 + c);
        });
    for (auto &th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// chains x (nsamples - 1) sweeps of gibbs_full with relabelling on `threads` host threads; returns seconds of wall time
extern "C" double opt_cpu_full_gibbs(const int32_t *X, int N, int P, int K, int nsamples, int burnin, int chains, int threads) {
    if (P > 64 || K > 64 || threads < 1) return -1.0;
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([=] {
            volatile unsigned sink = 0;
            for (int c = t; c < chains; c += threads) sink = sink + one_chain(X, N, P, K, nsamples, burnin, 1000 + c);
        });
    for (auto &th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
