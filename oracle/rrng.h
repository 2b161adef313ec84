// TEST INFRASTRUCTURE -- CPU oracle, never linked into or called by the product path.
//
// R-compatible random-number shim for the oracle restatement of bmm-mcmc's samplers.
// The reference draws through R's global RNG (Rcpp::RNGScope, src/RcppExports.cpp:14):
//   rmultinom  full_gibbs.cpp:133, stickbreaking.cpp:116, collapsed_gibbs.cpp:154
//   R::rbeta   full_gibbs.cpp:219, stickbreaking.cpp:190,223, utils.cpp:8
//   R::rgamma  full_gibbs.cpp:18, utils.cpp:12
//   RcppArmadillo::sample  collapsed_gibbs_dp.cpp:207
// R nmath / RcppArmadillo are NOT under /root/reference (third-party, version unpinned:
// DESCRIPTION:10 only asks Rcpp >= 1.0.2; fixtures were written by R 3.6.1).  Restated here:
//   * set.seed scrambling + MT19937 unif_rand + rbinom(1,p) inversion: PINNED bit-exactly by
//     regenerating data/*.RData from set.seed(17) (tests/test_fixtures.py).
//   * rmultinom(1,..) (sequential conditional binomials, long double running total) and
//     RcppArmadillo::sample(x,1,false,prob) (descending sort + cumulative walk): restated
//     from the published algorithms; parity unpinned beyond the rbinom/unif_rand core.
//     Tie rule for equal probabilities in sample(): std::sort on (value, index) packets, as
//     Armadillo's sort_index does (libstdc++: stable up to 16 candidates, introsort above).
//   * rgamma / rbeta: NOT R's Ahrens-Dieter / Cheng generators.  Any exact sampler has the
//     same law, so the oracle uses Marsaglia-Tsang + Box-Muller on the same uniform stream.
//     Consequence: theta/pi/alpha draws agree with R only in distribution ("parity
//     unpinned" for the parameter-draw values; pinned for everything deterministic).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace oracle {

struct RRng {
    uint32_t mt[624];
    int mti = 625;
    // Optional recorder: every unif_rand() consumed while rec != nullptr is appended.
    double *rec = nullptr;
    int rec_n = 0, rec_cap = 0;

    void set_seed(uint32_t seed) {
        for (int j = 0; j < 50; ++j) seed = 69069u * seed + 1u;
        uint32_t first = 0;
        for (int j = 0; j < 625; ++j) {
            seed = 69069u * seed + 1u;
            if (j == 0) first = seed; else mt[j - 1] = seed;
        }
        (void)first;  // i_seed[0] is overwritten by mti = 624 (FixupSeeds)
        mti = 624;
    }

    uint32_t genrand() {
        static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
        const int N = 624, M = 397;
        uint32_t y;
        if (mti >= N) {
            int kk;
            for (kk = 0; kk < N - M; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + M] ^ (y >> 1) ^ mag01[y & 0x1];
            }
            for (; kk < N - 1; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ mag01[y & 0x1];
            }
            y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ mag01[y & 0x1];
            mti = 0;
        }
        y = mt[mti++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }

    double unif_rand() {
        const double i2_32m1 = 2.328306437080797e-10;
        double v = (double)genrand() * 2.3283064365386963e-10;
        if (v <= 0.0) v = 0.5 * i2_32m1;
        if (1.0 - v <= 0.0) v = 1.0 - 0.5 * i2_32m1;
        if (rec && rec_n < rec_cap) rec[rec_n] = v;
        if (rec) rec_n++;
        return v;
    }

    // rbinom(n = 1, pp): inversion branch of nmath/rbinom.c (np < 30).
    int rbinom1(double pp) {
        if (pp == 0.0) return 0;
        if (pp == 1.0) return 1;
        double p = std::fmin(pp, 1.0 - pp), q = 1.0 - p, r = p / q, g = r * 2.0;
        int ix;
        for (;;) {
            ix = 0;
            double f = q, u = unif_rand();
            bool done = false;
            for (;;) {
                if (u < f) { done = true; break; }
                if (ix > 110) break;
                u -= f;
                ix++;
                f *= (g / ix - r);
            }
            if (done) break;
        }
        if (pp > 0.5) ix = 1 - ix;
        return ix;
    }

    // rmultinom(1, prob, K, rN).  Returns 0, or -1 when R would yield NA / raise.
    int rmultinom1(const double *prob, int K, int *rN) {
        long double p_tot = 0.0L;
        for (int k = 0; k < K; ++k) {
            double pp = prob[k];
            if (!std::isfinite(pp) || pp < 0.0 || pp > 1.0) { rN[k] = std::numeric_limits<int>::min(); return -1; }
            p_tot += pp;
            rN[k] = 0;
        }
        if (std::fabs((double)(p_tot - 1.0L)) > 1e-7) return -1;
        int n = 1;
        for (int k = 0; k < K - 1; ++k) {
            if (prob[k] != 0.0) {
                double pp = (double)(prob[k] / p_tot);
                rN[k] = (pp < 1.0) ? rbinom1(pp) : n;
                n -= rN[k];
            } else {
                rN[k] = 0;
            }
            if (n <= 0) return 0;
            p_tot -= prob[k];
        }
        rN[K - 1] = n;
        return 0;
    }

    // RcppArmadillo::sample(x, 1, false, prob)(0): index into x of the drawn element.
    // prob is normalised (FixProb), sorted descending, walked cumulatively.  Armadillo's sort_index is
    // std::sort over (value, index) packets with the comparator a.val > b.val: NOT a stable sort.  With
    // libstdc++ it is an insertion sort (ties keep index order) up to 16 candidates and an introsort above;
    // the compiled reference (oracle/_ref/libbmm_ref.so) shows the difference from sweep to sweep once a
    // DP chain holds 16+ clusters with tied probabilities.  stable_ties = true selects the stable order for
    // every n (the documented contract of the GPU's Philox path, where the order is immaterial in law).
    bool stable_ties = false;
    struct SortPacket { double val; int index; };
    int sample1(const double *prob_in, int n, double *scratch_p, int *scratch_perm) {
        double sum = 0.0;
        for (int i = 0; i < n; ++i) sum += prob_in[i];
        SortPacket pk_small[64];
        std::vector<SortPacket> pk_big;
        SortPacket *pk = pk_small;
        if (n > 64) { pk_big.resize(n); pk = pk_big.data(); }
        for (int i = 0; i < n; ++i) { pk[i].val = prob_in[i] / sum; pk[i].index = i; }
        auto cmp = [](const SortPacket &a, const SortPacket &b) { return a.val > b.val; };
        if (stable_ties) std::stable_sort(pk, pk + n, cmp);
        else std::sort(pk, pk + n, cmp);
        for (int i = 0; i < n; ++i) { scratch_p[i] = pk[i].val; scratch_perm[i] = pk[i].index; }
        double rT = 1.0 * unif_rand(), mass = 0.0;
        int jj;
        for (jj = 0; jj < n - 1; ++jj) {
            mass += scratch_p[jj];
            if (rT <= mass) break;
        }
        return scratch_perm[jj];
    }

    // ---- parameter draws: same laws as R::rgamma / R::rbeta, different algorithm ----
    double norm_rand() {
        double u1 = unif_rand(), u2 = unif_rand();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586477 * u2);
    }
    double rgamma(double shape, double scale) {
        if (!(shape > 0.0) || !(scale > 0.0)) return (shape == 0.0) ? 0.0 : std::numeric_limits<double>::quiet_NaN();
        double boost = 1.0;
        if (shape < 1.0) {
            boost = std::pow(unif_rand(), 1.0 / shape);
            shape += 1.0;
        }
        double d = shape - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
        for (;;) {
            double x = norm_rand(), v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            double u = unif_rand();
            if (std::log(u) < 0.5 * x * x + d - d * v + d * std::log(v)) return scale * boost * d * v;
        }
    }
    double rbeta(double a, double b) {
        double x = rgamma(a, 1.0), y = rgamma(b, 1.0);
        return x / (x + y);
    }
};

}  // namespace oracle
