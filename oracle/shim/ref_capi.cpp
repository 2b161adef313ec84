// TEST INFRASTRUCTURE -- never linked into or called by the product path.
//
// C entry points of oracle/_ref/libbmm_ref.so: the reference's own, UNMODIFIED C++ sources
// (/root/reference/src/*.cpp, compiled where they lie by oracle/build_ref.sh against the header shim in
// this directory) driven the way R drives them: R_init_bmmmcmc registers the seven .Call symbols
// (src/RcppExports.cpp:137-151) and ref_dotcall looks one up by name and calls it with SEXP arguments.
// my_stephens_online and update_alpha are not registered with R (no [[Rcpp::export]]); they are reached
// through the reference's own headers (src/stephens.h:6, src/utils.h:3).
#include <RcppArmadillo.h>

#include "stephens.h"  // /root/reference/src/stephens.h (pulls my_lpsolve.h)
#include "utils.h"     // /root/reference/src/utils.h

extern "C" void R_init_bmmmcmc(DllInfo *dll);

namespace {

DllInfo g_dll;
bool g_init = false;
thread_local std::string g_err;

struct Handle {
    std::shared_ptr<SEXPREC> p;
};

const R_CallMethodDef *lookup(const char *name) {
    if (!g_init) {
        R_init_bmmmcmc(&g_dll);
        g_init = true;
    }
    for (const R_CallMethodDef *e = g_dll.call_entries; e && e->name; ++e)
        if (std::strcmp(e->name, name) == 0) return e;
    return nullptr;
}

typedef SEXP (*F1)(SEXP);
typedef SEXP (*F2)(SEXP, SEXP);
typedef SEXP (*F12)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F13)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F14)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

}  // namespace

extern "C" {

struct ref_arg {
    int type;  // INTSXP 13, REALSXP 14, LGLSXP 10
    int ndim;  // 0 = plain vector
    int dim[3];
    const void *data;  // int32 / double, column-major
    long long n;
};

const char *ref_last_error() { return g_err.c_str(); }

// set.seed(seed) for this thread's generator
void ref_set_seed(unsigned seed) { bmm_shim::rng().set_seed(seed); }

// record every unif_rand() consumed from now on (flat stream); cap = 0 stops recording
void ref_record_uniforms(double *buf, int cap) {
    oracle::RRng &r = bmm_shim::rng();
    r.rec = cap > 0 ? buf : nullptr;
    r.rec_cap = cap;
    r.rec_n = 0;
}
int ref_recorded() { return bmm_shim::rng().rec_n; }

int ref_n_registered() {
    lookup("");
    int n = 0;
    for (const R_CallMethodDef *e = g_dll.call_entries; e && e->name; ++e) ++n;
    return n;
}
const char *ref_registered_name(int i, int *nargs) {
    lookup("");
    const R_CallMethodDef *e = g_dll.call_entries + i;
    if (nargs) *nargs = e->numArgs;
    return e->name;
}

// .Call(symbol, args...): returns an opaque handle (ref_free) or NULL with ref_last_error() set
void *ref_dotcall(const char *symbol, int nargs, const ref_arg *args) {
    g_err.clear();
    const R_CallMethodDef *e = lookup(symbol);
    if (!e) {
        g_err = std::string("symbol not registered: ") + symbol;
        return nullptr;
    }
    if (e->numArgs != nargs) {
        g_err = "Incorrect number of arguments (" + std::to_string(nargs) + "), expecting " + std::to_string(e->numArgs) +
                " for '" + symbol + "'";
        return nullptr;
    }
    std::vector<SEXPREC> store((size_t)nargs);
    std::vector<SEXP> a((size_t)nargs);
    for (int i = 0; i < nargs; ++i) {
        SEXPREC &s = store[(size_t)i];
        s.type = args[i].type;
        if (s.type == REALSXP) s.real.assign((const double *)args[i].data, (const double *)args[i].data + args[i].n);
        else s.integer.assign((const int *)args[i].data, (const int *)args[i].data + args[i].n);
        s.dim.assign(args[i].dim, args[i].dim + args[i].ndim);
        a[(size_t)i] = &s;
    }
    bmm_shim::arena().clear();
    Rcpp::last_condition().clear();
    SEXP r = nullptr;
    void *f = (void *)e->fun;
    switch (nargs) {
        case 1: r = ((F1)f)(a[0]); break;
        case 2: r = ((F2)f)(a[0], a[1]); break;
        case 12: r = ((F12)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11]); break;
        case 13: r = ((F13)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12]); break;
        case 14: r = ((F14)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13]); break;
        default: g_err = "ref_dotcall: arity not supported"; return nullptr;
    }
    if (!r) {
        g_err = Rcpp::last_condition().empty() ? "NULL result" : Rcpp::last_condition();
        bmm_shim::arena().clear();
        return nullptr;
    }
    Handle *h = new Handle();
    for (auto &sp : bmm_shim::arena())
        if (sp.get() == r) h->p = sp;
    bmm_shim::arena().clear();
    if (!h->p) {
        delete h;
        g_err = "ref_dotcall: result not found in the arena";
        return nullptr;
    }
    return h;
}

void ref_free(void *h) { delete (Handle *)h; }

int ref_type(void *h) { return ((Handle *)h)->p->type; }
long long ref_length(void *h) { return (long long)((Handle *)h)->p->length(); }
int ref_ndim(void *h) { return (int)((Handle *)h)->p->dim.size(); }
int ref_dim(void *h, int i) { return ((Handle *)h)->p->dim[(size_t)i]; }
const void *ref_data(void *h) {
    SEXPREC *s = ((Handle *)h)->p.get();
    return s->type == REALSXP ? (const void *)s->real.data() : (const void *)s->integer.data();
}
const char *ref_list_name(void *h, int i) { return ((Handle *)h)->p->names[(size_t)i].c_str(); }
void *ref_list_elt(void *h, int i) {
    Handle *e = new Handle();
    e->p = ((Handle *)h)->p->elts[(size_t)i];
    return e;
}

// my_stephens_online (src/stephens.cpp:66-94): q, p are N x K column-major
int ref_stephens_online(int N, int K, const double *q, const double *p, int sample_num, int *perm, double *q_new) {
    g_err.clear();
    try {
        arma::mat Q((arma::uword)N, (arma::uword)K), Pm((arma::uword)N, (arma::uword)K);
        std::memcpy(Q.mem, q, sizeof(double) * (size_t)N * K);
        std::memcpy(Pm.mem, p, sizeof(double) * (size_t)N * K);
        std::pair<arma::Row<int>, arma::mat> out = my_stephens_online(Q, Pm, sample_num, false);
        for (int k = 0; k < K; ++k) perm[k] = out.first(k);
        std::memcpy(q_new, out.second.mem, sizeof(double) * (size_t)N * K);
        return 0;
    } catch (std::exception &ex) {
        g_err = ex.what();
        return -1;
    }
}

// update_alpha (src/utils.cpp:6-14), drawing from this thread's generator
double ref_update_alpha(double alpha_old, double a, double b, int N, int K) { return update_alpha(alpha_old, a, b, N, K); }

}  // extern "C"
