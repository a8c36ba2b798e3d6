// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE ONLY (see limbo_oracle.hpp).
// extern "C" surface over the templated restatement so tests/, smoke() and
// bench.py's cpu_baseline leg can drive it through ctypes.  prec: 0 = double
// (the reference's arithmetic), 1 = long double (x87 80-bit), 2 = __float128.
#include "limbo_oracle.hpp"
#include <cstring>
#include <thread>

namespace {

struct IGP {
    virtual ~IGP() {}
    virtual void set_data(long N, int D, int P, const double* X, const double* om) = 0;
    virtual void set_kernel(int id, const double* hp, int nh, double noise) = 0;
    virtual long fit() = 0;
    virtual void refit_alpha(const double* om) = 0;
    virtual void append(const double* x, const double* om) = 0;
    virtual void query(long M, const double* Xq, double* mu, double* s2, int nthreads) const = 0;
    virtual double log_lik() const = 0;
    virtual void grad(double* g, int optimize_noise) = 0;
    virtual double loo_cv() = 0;
    virtual void loo_grad(double* g, int optimize_noise) = 0;
    virtual void kinv_obs(double* out) = 0;
    virtual void get(int what, double* dst) = 0;
    virtual long n() const = 0;
    virtual IGP* clone() const = 0;
};

template <typename T>
struct GPImpl : IGP {
    lbo::GP<T> gp;
    void set_data(long N, int D, int P, const double* X, const double* om) override
    {
        gp.set_data(N, D, P, X, om);
        gp.kern.D = D;
    }
    void set_kernel(int id, const double* hp, int nh, double noise) override
    {
        gp.kern.id = id;
        gp.kern.D = gp.D;
        gp.kern.noise = T(noise);
        gp.kern.klam = (id == lbo::K_SE_ARD && gp.D > 0) ? (nh - 1 - gp.D) / gp.D : 0; // Params::kernel_squared_exp_ard::k()
        gp.kern.set_params(hp);
    }
    long fit() override { gp.compute_full_kernel(); return gp.chol_info; }
    void refit_alpha(const double* om) override
    {
        gp.obs_mean.assign(om, om + (size_t)gp.N * gp.P);
        gp.compute_alpha(); // gp.hpp:250-251
    }
    void append(const double* x, const double* om) override { gp.append(x, om); }
    void query(long M, const double* Xq, double* mu, double* s2, int nthreads) const override
    {
        // the reference evaluates candidates one at a time; tools::par fans
        // them over host threads (tools/parallel.hpp:116-229)
        auto work = [&](long lo, long hi) {
            std::vector<T> k, m(gp.P);
            for (long q = lo; q < hi; ++q) {
                T sg;
                gp.query(Xq + q * gp.D, m.data(), &sg, k);
                for (int p = 0; p < gp.P; ++p) mu[q * gp.P + p] = (double)m[p];
                s2[q] = (double)sg;
            }
        };
        if (nthreads <= 1 || M < 2) { work(0, M); return; }
        std::vector<std::thread> th;
        long chunk = (M + nthreads - 1) / nthreads;
        for (int t = 0; t < nthreads; ++t) {
            long lo = t * chunk, hi = std::min(M, lo + chunk);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto& t : th) t.join();
    }
    double log_lik() const override { return (double)gp.log_lik(); }
    void grad(double* g, int on) override
    {
        int nh = gp.kern.n_params() + (on ? 1 : 0);
        std::vector<T> gg(nh);
        gp.kernel_grad_log_lik(gg.data(), on != 0);
        for (int i = 0; i < nh; ++i) g[i] = (double)gg[i];
    }
    double loo_cv() override { return (double)gp.log_loo_cv(); }
    void loo_grad(double* g, int on) override
    {
        int nh = gp.kern.n_params() + (on ? 1 : 0);
        std::vector<T> gg(nh);
        gp.kernel_grad_log_loo_cv(gg.data(), on != 0);
        for (int i = 0; i < nh; ++i) g[i] = (double)gg[i];
    }
    void kinv_obs(double* out) override
    {
        std::vector<T> w((size_t)gp.N * gp.P);
        gp.kinv_obs(w.data());
        for (size_t i = 0; i < w.size(); ++i) out[i] = (double)w[i];
    }
    void get(int what, double* dst) override
    {
        const std::vector<T>* src = nullptr;
        if (what == 0) src = &gp.K;
        else if (what == 1) src = &gp.L;
        else if (what == 2) src = &gp.alpha;
        else if (what == 3) { if (!gp.inv_updated) gp.compute_inv_kernel(); src = &gp.Kinv; }
        if (!src) return;
        for (size_t i = 0; i < src->size(); ++i) dst[i] = (double)(*src)[i];
    }
    long n() const override { return gp.N; }
    IGP* clone() const override { return new GPImpl<T>(*this); }
};

IGP* make(int prec)
{
    if (prec == 1) return new GPImpl<long double>();
#ifdef LBO_HAVE_QUAD
    if (prec == 2) return new GPImpl<__float128>();
#endif
    return new GPImpl<double>();
}

} // namespace

extern "C" {

void* lbo_create(int prec) { return make(prec); }
void lbo_destroy(void* h) { delete (IGP*)h; }
void* lbo_clone(void* h) { return ((IGP*)h)->clone(); }
void lbo_set_data(void* h, long N, int D, int P, const double* X, const double* om) { ((IGP*)h)->set_data(N, D, P, X, om); }
void lbo_set_kernel(void* h, int id, const double* hp, int nh, double noise) { ((IGP*)h)->set_kernel(id, hp, nh, noise); }
long lbo_fit(void* h) { return ((IGP*)h)->fit(); }
void lbo_refit_alpha(void* h, const double* om) { ((IGP*)h)->refit_alpha(om); }
void lbo_append(void* h, const double* x, const double* om) { ((IGP*)h)->append(x, om); }
void lbo_query(void* h, long M, const double* Xq, double* mu, double* s2, int nthreads) { ((IGP*)h)->query(M, Xq, mu, s2, nthreads); }
double lbo_log_lik(void* h) { return ((IGP*)h)->log_lik(); }
void lbo_grad(void* h, double* g, int optimize_noise) { ((IGP*)h)->grad(g, optimize_noise); }
double lbo_loo_cv(void* h) { return ((IGP*)h)->loo_cv(); }
void lbo_loo_grad(void* h, double* g, int optimize_noise) { ((IGP*)h)->loo_grad(g, optimize_noise); }
void lbo_kinv_obs(void* h, double* out) { ((IGP*)h)->kinv_obs(out); }
void lbo_get(void* h, int what, double* dst) { ((IGP*)h)->get(what, dst); }
long lbo_n(void* h) { return ((IGP*)h)->n(); }

// scalar kernel / gradient evaluation (test_kernel.cpp-style checks)
double lbo_kernel_eval(int id, int D, const double* hp, double noise, const double* x1, const double* x2, int same_index)
{
    lbo::Kernel<double> k;
    k.id = id; k.D = D; k.noise = noise;
    k.set_params(hp);
    return same_index ? k(x1, x2, 0, 0) : k(x1, x2);
}
void lbo_kernel_grad(int id, int D, const double* hp, const double* x1, const double* x2, double* g)
{
    lbo::Kernel<double> k;
    k.id = id; k.D = D;
    k.set_params(hp);
    k.gradient(x1, x2, g);
}

// acquisition epilogues, vectorised over M candidates
void lbo_ucb(long M, const double* mu0, const double* s2, double alpha, double* out)
{
    for (long i = 0; i < M; ++i) out[i] = lbo::ucb<double>(mu0[i], s2[i], alpha);
}
void lbo_ei(long M, const double* mu0, const double* s2, double f_max, double jitter, double* out)
{
    for (long i = 0; i < M; ++i) out[i] = lbo::ei<double>(mu0[i], s2[i], f_max, jitter);
}
double lbo_gp_ucb_beta(int iteration, int dim_in, double delta) { return lbo::gp_ucb_beta(iteration, dim_in, delta); }

// KernelLFOpt objective (model/gp/kernel_lf_opt.hpp:77-92): copy the GP, set
// h-params, recompute(false), log-lik (+ gradient).
double lbo_lml_eval(void* h, const double* hp, int nh, double noise, int kernel_id, double* grad, int optimize_noise)
{
    IGP* g = ((IGP*)h)->clone();
    double nz = optimize_noise ? std::exp(2.0 * hp[nh - 1]) : noise; // kernel.hpp:116-123
    g->set_kernel(kernel_id, hp, nh, nz);
    g->fit();
    double lik = g->log_lik();
    if (grad) g->grad(grad, optimize_noise);
    delete g;
    return lik;
}

// opt::Rprop driving KernelLFOptimization; returns best-seen params.
void lbo_rprop_lml(void* h, int kernel_id, const double* init, int nh, double noise, int iterations, double eps_stop,
    int optimize_noise, double* out_params, long* n_evals)
{
    std::vector<double> p0(init, init + nh);
    long ne = 0;
    auto f = [&](const std::vector<double>& p, double* g) {
        return lbo_lml_eval(h, p.data(), nh, noise, kernel_id, g, optimize_noise);
    };
    std::vector<double> best = lbo::rprop(f, p0, iterations, eps_stop, false, &ne);
    std::memcpy(out_params, best.data(), sizeof(double) * nh);
    if (n_evals) *n_evals = ne;
}

// standalone dense helpers (used to cross-check the restated LLT)
long lbo_cholesky(long n, double* A) { return lbo::chol_blocked<double>(A, n, n); }

int lbo_has_quad()
{
#ifdef LBO_HAVE_QUAD
    return 1;
#else
    return 0;
#endif
}
}
