// TEST INFRASTRUCTURE -- stand-in for <RcppArmadilloExtensions/sample.h>
// (included by /root/reference/src/collapsed_gibbs_dp.cpp:3, called at :207 as
//  RcppArmadillo::sample(choices, 1, false, probs_norm)).
// RcppArmadillo is not under /root/reference; restated from its published algorithm [memory]:
// FixProb (validate, normalise) then ProbSampleNoReplace (descending sort_index + sort, one unif_rand per
// draw, cumulative walk over the first n-1 entries, drawn entry removed) or ProbSampleReplace; the
// Walker-alias branch (>= 200 sizeable probabilities) is not reachable from the path and is rejected.
#pragma once
#include "../RcppArmadillo.h"

namespace Rcpp {
namespace RcppArmadillo {

inline void FixProb(arma::vec &prob, const int size, const bool replace) {
    double sum = 0.0;
    int nPos = 0;
    const int nn = (int)prob.n_elem;
    for (int ii = 0; ii < nn; ii++) {
        double p = prob.mem[ii];
        if (!std::isfinite(p)) throw std::range_error("NAs not allowed in probability");
        if (p < 0.0) throw std::range_error("Negative probabilities not allowed");
        if (p > 0.0) {
            nPos++;
            sum += p;
        }
    }
    if (nPos == 0 || (!replace && size > nPos)) throw std::range_error("Not enough positive probabilities");
    for (int ii = 0; ii < nn; ii++) prob.mem[ii] = prob.mem[ii] / sum;
}

inline void ProbSampleNoReplace(arma::uvec &index, int nOrig, int size, arma::vec &prob) {
    int ii, jj, kk;
    int nOrig_1 = nOrig - 1;
    double rT, mass, totalmass = 1.0;
    arma::uvec perm = arma::sort_index(prob, "descend");
    prob = arma::sort(prob, "descend");
    for (ii = 0; ii < size; ii++, nOrig_1--) {
        rT = totalmass * unif_rand();
        mass = 0;
        for (jj = 0; jj < nOrig_1; jj++) {
            mass += prob.mem[jj];
            if (rT <= mass) break;
        }
        index.mem[ii] = perm.mem[jj];
        totalmass -= prob.mem[jj];
        for (kk = jj; kk < nOrig_1; kk++) {
            prob.mem[kk] = prob.mem[kk + 1];
            perm.mem[kk] = perm.mem[kk + 1];
        }
    }
}

inline void ProbSampleReplace(arma::uvec &index, int nOrig, int size, arma::vec &prob) {
    int ii, jj;
    int nOrig_1 = nOrig - 1;
    double rU;
    arma::uvec perm = arma::sort_index(prob, "descend");
    prob = arma::sort(prob, "descend");
    for (ii = 1; ii < nOrig; ii++) prob.mem[ii] += prob.mem[ii - 1];  // cumulative
    for (ii = 0; ii < size; ii++) {
        rU = unif_rand();
        for (jj = 0; jj < nOrig_1; jj++)
            if (rU <= prob.mem[jj]) break;
        index.mem[ii] = perm.mem[jj];
    }
}

template <class T>
T sample(const T &x, const int size, const bool replace, NumericVector prob_ = NumericVector(0)) {
    int nOrig = x.size();
    int probsize = prob_.size();
    T ret(size);
    if (size > nOrig && !replace) throw std::range_error("Tried to sample more elements than in x without replacement");
    if (probsize == 0) throw std::range_error("shim sample(): the unweighted branch is not on the path");
    if (probsize != nOrig) throw std::range_error("Number of probabilities must equal input vector length.");
    arma::uvec index((arma::uword)size);
    arma::vec fixprob(prob_.d.data(), (arma::uword)probsize);
    FixProb(fixprob, size, replace);
    if (replace) {
        int walker_test = 0;
        for (int i = 0; i < nOrig; ++i) walker_test += (fixprob.mem[i] * nOrig) > 0.1;
        if (walker_test >= 200) throw std::range_error("shim sample(): Walker alias branch is not on the path");
        ProbSampleReplace(index, nOrig, size, fixprob);
    } else {
        ProbSampleNoReplace(index, nOrig, size, fixprob);
    }
    for (int ii = 0; ii < size; ii++) ret[ii] = x[(int)index.mem[ii]];
    return ret;
}

}  // namespace RcppArmadillo
}  // namespace Rcpp
