// TEST INFRASTRUCTURE.  The Rcpp host of the R package (r-package/src/host.cpp), compiled UNMODIFIED against the Rcpp
// stand-in the oracle uses for the reference's sources (oracle/shim/) and linked with libbmm_b200.so, driven the way R's
// generated RcppExports.cpp would drive it: Rcpp matrices in, the returned list out.  No R in this image, so this is how
// the host's argument forwarding, output allocation (names, shapes, storage modes of full_gibbs.cpp:233-248 etc.) and
// error propagation get exercised; tests/test_rhost.py compares the lists with the ctypes path.
#include <cstring>
#include <string>

#include "../../r-package/src/host.cpp"

namespace {
thread_local std::string g_err;

template <typename T>
bool copy_out(const List &l, const char *name, T *dst, long long cap, int *dims3) {
    for (size_t i = 0; i < l.p->names.size(); ++i)
        if (l.p->names[i] == name) {
            const SEXPREC *s = l.p->elts[i].get();
            const std::vector<T> &v = sexp_traits<T>::vec(*const_cast<SEXPREC *>(s));
            if ((long long)v.size() > cap) return false;
            if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(T));
            if (dims3) { dims3[0] = dims3[1] = dims3[2] = 0; for (size_t k = 0; k < s->dim.size() && k < 3; ++k) dims3[k] = s->dim[k]; }
            return true;
        }
    return false;
}
}  // namespace

extern "C" {

const char *rhost_last_error() { return g_err.c_str(); }
void rhost_set_seed(unsigned seed) { bmm_shim::rng().set_seed(seed); }

// names of the returned list, comma separated (the contract: R/utils.R readers use obj$theta, obj$z, obj$pi ...)
// sampler: 0 gibbs_cpp, 1 gibbs_stickbreaking_cpp, 2 collapsed_gibbs_cpp, 3 collapsed_gibbs_dp_cpp
int rhost_gibbs(int sampler, const int *X, int N, int P, const double *init_pi, const double *init_theta, const int *init_z,
                int nsamples, int K, double alpha, double beta, double gamma, double a, double b, int burnin, int relabel,
                int burnrelabel, char *names_out, int names_cap, double *pi, double *alpha_out, int *perms, int *z, double *theta,
                int *z_orig, double *theta_orig, int *theta_dims) {
    try {
        IntegerMatrix df(N, P);
        std::memcpy(df.begin(), X, sizeof(int) * (size_t)N * P);
        List r;
        if (sampler <= 1) {
            NumericVector ip(K);
            std::memcpy(ip.begin(), init_pi, sizeof(double) * K);
            NumericMatrix th(K, P);
            std::memcpy(th.begin(), init_theta, sizeof(double) * (size_t)K * P);
            r = sampler == 0 ? gibbs_cpp(df, ip, th, nsamples, K, alpha, beta, gamma, a, b, burnin, relabel != 0, burnrelabel, false)
                             : gibbs_stickbreaking_cpp(df, ip, th, nsamples, K, alpha, beta, gamma, a, b, burnin, relabel != 0, burnrelabel, false);
        } else if (sampler == 2) {
            IntegerVector iz(N);
            std::memcpy(iz.begin(), init_z, sizeof(int) * N);
            r = collapsed_gibbs_cpp(df, iz, nsamples, K, alpha, beta, gamma, a, b, burnin, relabel != 0, burnrelabel, false);
        } else {
            r = collapsed_gibbs_dp_cpp(df, nsamples, alpha, beta, gamma, a, b, burnin, relabel != 0, burnrelabel, K, false);
        }
        std::string names;
        for (const std::string &n : r.p->names) names += (names.empty() ? "" : ",") + n;
        if ((int)names.size() + 1 > names_cap) { g_err = "names buffer too small"; return -1; }
        std::strcpy(names_out, names.c_str());
        const long long S = nsamples - burnin;
        copy_out<double>(r, "pi", pi, S * K, nullptr);
        copy_out<double>(r, "alpha", alpha_out, S, nullptr);
        copy_out<int>(r, "permutations", perms, S * K, nullptr);
        copy_out<int>(r, "z", z, S * N, nullptr);
        copy_out<double>(r, "theta", theta, (long long)K * P * S, theta_dims);
        copy_out<int>(r, "z_original", z_orig, S * N, nullptr);
        copy_out<double>(r, "theta_original", theta_orig, (long long)K * P * S, nullptr);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// predict_gibbs: log_pred[M], membership[M x K cm]
int rhost_predictive(const int *Xnew, int M, int P, int K, int S, const double *theta, const double *pi, double *log_pred,
                     double *membership) {
    try {
        IntegerMatrix nd(M, P);
        std::memcpy(nd.begin(), Xnew, sizeof(int) * (size_t)M * P);
        NumericVector th(Dimension(K, P, S));
        std::memcpy(th.begin(), theta, sizeof(double) * (size_t)K * P * S);
        NumericMatrix pm(S, K);
        std::memcpy(pm.begin(), pi, sizeof(double) * (size_t)S * K);
        List r = predictive_cpp(nd, th, pm);
        copy_out<double>(r, "log_pred", log_pred, M, nullptr);
        copy_out<double>(r, "membership", membership, (long long)M * K, nullptr);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

int rhost_lpsolve(const double *cost, int K, int *sol) {
    try {
        NumericMatrix c(K, K);
        std::memcpy(c.begin(), cost, sizeof(double) * (size_t)K * K);
        IntegerMatrix s = my_lpsolve(c);
        std::memcpy(sol, s.begin(), sizeof(int) * (size_t)K * K);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"
