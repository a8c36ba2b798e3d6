// tests/cpp/dropin_test.cpp — the drop-in claim, compiled: the reference's OWN policy templates
// (acqui::UCB, acqui::EI, model::gp::KernelLFOpt / KernelLooOpt / KernelMeanLFOpt<Params, opt::Rprop>, kernel::*, mean::* from
// /root/reference/src/limbo) instantiated over limbo_b200::model::GP and, side by side, over the
// reference's limbo::model::GP; results must agree to the fp64 bar.  Needs a GPU to run.
// (Eigen is the stand-in from oracle/ref_shim because the image has no Eigen.)
#include <cstdio>
#include <limbo/acqui/ei.hpp>
#include <limbo/acqui/ucb.hpp>
#include <limbo/kernel/matern_five_halves.hpp>
#include <limbo/kernel/squared_exp_ard.hpp>
#include <limbo/mean/constant.hpp>
#include <limbo/mean/data.hpp>
#include <limbo/mean/function_ard.hpp>
#include <limbo/model/gp.hpp>
#include <limbo/model/gp/kernel_lf_opt.hpp>
#include <limbo/model/gp/kernel_loo_opt.hpp>
#include <limbo/model/gp/kernel_mean_lf_opt.hpp>
#include <limbo/model/multi_gp.hpp>
#include <limbo/opt/rprop.hpp>

#include <limbo_b200/model/gp.hpp>
#include <limbo_b200/opt/batched_random.hpp>

using namespace limbo;

struct Params {
    struct kernel : public defaults::kernel {};
    struct kernel_squared_exp_ard : public defaults::kernel_squared_exp_ard {};
    struct kernel_maternfivehalves : public defaults::kernel_maternfivehalves {};
    struct opt_rprop {
        BO_PARAM(int, iterations, 6);
        BO_PARAM(double, eps_stop, 0.0);
    };
    struct acqui_ucb : public defaults::acqui_ucb {};
    struct acqui_ei : public defaults::acqui_ei {};
    struct mean_constant {
        BO_PARAM(double, constant, 0.5);
    };
    struct opt_batchedrandom : public limbo_b200::defaults::opt_batchedrandom {
        BO_PARAM(int, candidates, 4000);
    };
};

struct FirstElem {
    double operator()(const Eigen::VectorXd& x) const { return x(0); }
};

static double u01(unsigned long long& s)
{ // splitmix64, as limbo_b200/synth.py
    s += 0x9E3779B97F4A7C15ULL;
    unsigned long long z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// SquaredExpARD with two Lambda columns (squared_exp_ard.hpp:109-126,142-146)
struct ParamsLambda : Params {
    struct kernel_squared_exp_ard {
        BO_PARAM(int, k, 2);
        BO_PARAM(double, sigma_sq, 1);
    };
};

template <typename Kernel, typename Params = ::Params>
int run_case(const char* name, int N, int D)
{
    using HP = model::gp::KernelLFOpt<Params, opt::Rprop<Params>>;
    using RefGP = model::GP<Params, Kernel, mean::Data<Params>, HP>;
    using NewGP = limbo_b200::model::GP<Params, Kernel, mean::Data<Params>, HP>;
    unsigned long long seed = 42;
    std::vector<Eigen::VectorXd> X, Y, Q;
    for (int i = 0; i < N + 3; ++i) {
        Eigen::VectorXd x((Eigen::Index)D), y((Eigen::Index)1);
        double s = 0;
        for (int d = 0; d < D; ++d) { x(d) = u01(seed); s += std::cos(3.0 * x(d)); }
        y(0) = s;
        X.push_back(x);
        Y.push_back(y);
    }
    for (int i = 0; i < 64; ++i) {
        Eigen::VectorXd q((Eigen::Index)D);
        for (int d = 0; d < D; ++d) q(d) = u01(seed);
        Q.push_back(q);
    }
    std::vector<Eigen::VectorXd> X0(X.begin(), X.begin() + N), Y0(Y.begin(), Y.begin() + N);
    RefGP ref(D, 1);
    NewGP gpu(D, 1);
    ref.compute(X0, Y0);
    gpu.compute(X0, Y0);
    for (int i = N; i < N + 3; ++i) { ref.add_sample(X[i], Y[i]); gpu.add_sample(X[i], Y[i]); } // gp.hpp:126-152
    ref.optimize_hyperparams(); // the reference's KernelLFOpt<Rprop> driving each model
    gpu.optimize_hyperparams();
    double dh = (ref.kernel_function().h_params() - gpu.kernel_function().h_params()).norm();
    double dll = std::fabs(ref.get_log_lik() - gpu.get_log_lik()) / std::fabs(ref.get_log_lik());
    acqui::UCB<Params, RefGP> ucb_r(ref);
    acqui::UCB<Params, NewGP> ucb_n(gpu);
    acqui::EI<Params, RefGP> ei_r(ref);
    acqui::EI<Params, NewGP> ei_n(gpu);
    FirstElem afun;
    double dmu = 0, ds = 0, ducb = 0, dei = 0;
    for (auto& q : Q) {
        Eigen::VectorXd m1, m2;
        double s1, s2;
        std::tie(m1, s1) = ref.query(q);
        std::tie(m2, s2) = gpu.query(q);
        dmu = std::max(dmu, std::fabs(m1(0) - m2(0)));
        ds = std::max(ds, std::fabs(s1 - s2));
        ducb = std::max(ducb, std::fabs(opt::fun(ucb_r(q, afun, false)) - opt::fun(ucb_n(q, afun, false))));
        dei = std::max(dei, std::fabs(opt::fun(ei_r(q, afun, false)) - opt::fun(ei_n(q, afun, false))));
    }
    double dL = (ref.matrixL() - gpu.matrixL()).norm();
    NewGP copy(gpu); // value semantics (kernel_lf_opt.hpp:79)
    double dcopy = std::fabs(copy.sigma(Q[0]) - gpu.sigma(Q[0]));
    std::printf("%s N=%d D=%d |dh|=%.3e dll_rel=%.3e dmu=%.3e dsigma2=%.3e ducb=%.3e dei=%.3e |dL|=%.3e dcopy=%.3e\n", name, N + 3, D, dh,
        dll, dmu, ds, ducb, dei, dL, dcopy);
    bool ok = dh < 1e-8 && dll < 1e-10 && dmu < 1e-9 && ds < 1e-10 && ducb < 1e-9 && dei < 1e-9 && dL < 1e-8 && dcopy == 0.0;
    return ok ? 0 : 1;
}

// The reference's KernelLooOpt<Rprop> (model/gp/kernel_loo_opt.hpp) and KernelMeanLFOpt<Rprop> with
// mean::FunctionARD<mean::Constant> (kernel_mean_lf_opt.hpp, mean/function_ard.hpp) driving both model types.
template <typename HP, typename Mean>
int run_hp_case(const char* name, int N, int D, bool loo)
{
    using Kernel = kernel::SquaredExpARD<Params>;
    using RefGP = model::GP<Params, Kernel, Mean, HP>;
    using NewGP = limbo_b200::model::GP<Params, Kernel, Mean, HP>;
    unsigned long long seed = 77;
    std::vector<Eigen::VectorXd> X, Y;
    for (int i = 0; i < N; ++i) {
        Eigen::VectorXd x((Eigen::Index)D), y((Eigen::Index)1);
        double s = 1.5;
        for (int d = 0; d < D; ++d) { x(d) = u01(seed); s += std::cos(3.0 * x(d)); }
        y(0) = s;
        X.push_back(x);
        Y.push_back(y);
    }
    RefGP ref(D, 1);
    NewGP gpu(D, 1);
    ref.compute(X, Y);
    gpu.compute(X, Y);
    double dg;
    if (loo) dg = (ref.compute_kernel_grad_log_loo_cv() - gpu.compute_kernel_grad_log_loo_cv()).norm() / ref.compute_kernel_grad_log_loo_cv().norm();
    else dg = (ref.compute_mean_grad_log_lik() - gpu.compute_mean_grad_log_lik()).norm() / ref.compute_mean_grad_log_lik().norm();
    ref.optimize_hyperparams();
    gpu.optimize_hyperparams();
    double dh = (ref.kernel_function().h_params() - gpu.kernel_function().h_params()).norm();
    double dm = (ref.mean_function().h_params() - gpu.mean_function().h_params()).norm();
    double v1 = loo ? ref.get_log_loo_cv() : ref.get_log_lik(), v2 = loo ? gpu.get_log_loo_cv() : gpu.get_log_lik();
    double dv = std::fabs(v1 - v2) / std::fabs(v1);
    std::printf("%s N=%d D=%d dgrad_rel=%.3e |dh|=%.3e |dmean_h|=%.3e dobjective_rel=%.3e\n", name, N, D, dg, dh, dm, dv);
    return (dg < 1e-9 && dh < 1e-7 && dm < 1e-7 && dv < 1e-9) ? 0 : 1;
}

// model::MultiGP (model/multi_gp.hpp:60-63: template-template GP parameter) instantiated over the drop-in and over the
// reference's GP, side by side: compute, incremental add_sample, query.
int run_multigp()
{
    using Kernel = kernel::MaternFiveHalves<Params>;
    using Mean = mean::Constant<Params>;
    using RefM = model::MultiGP<Params, model::GP, Kernel, Mean>;
    using NewM = model::MultiGP<Params, limbo_b200::model::GP, Kernel, Mean>;
    unsigned long long seed = 5;
    const int N = 90, D = 2, P = 3;
    std::vector<Eigen::VectorXd> X, Y;
    for (int i = 0; i < N + 4; ++i) {
        Eigen::VectorXd x((Eigen::Index)D), y((Eigen::Index)P);
        for (int d = 0; d < D; ++d) x(d) = u01(seed);
        y(0) = std::cos(3.0 * x(0)) + x(1);
        y(1) = std::sin(4.0 * x(1));
        y(2) = x(0) * x(1);
        X.push_back(x);
        Y.push_back(y);
    }
    std::vector<Eigen::VectorXd> X0(X.begin(), X.begin() + N), Y0(Y.begin(), Y.begin() + N);
    RefM ref(D, P);
    NewM gpu(D, P);
    ref.compute(X0, Y0);
    gpu.compute(X0, Y0);
    for (int i = N; i < N + 4; ++i) { ref.add_sample(X[i], Y[i]); gpu.add_sample(X[i], Y[i]); } // multi_gp.hpp:139-176
    double dmu = 0, ds = 0;
    for (int q = 0; q < 40; ++q) {
        Eigen::VectorXd v((Eigen::Index)D);
        for (int d = 0; d < D; ++d) v(d) = u01(seed);
        Eigen::VectorXd m1, s1, m2, s2;
        std::tie(m1, s1) = ref.query(v);
        std::tie(m2, s2) = gpu.query(v);
        for (int p = 0; p < P; ++p) {
            dmu = std::max(dmu, std::fabs(m1(p) - m2(p)));
            ds = std::max(ds, std::fabs(s1(p) - s2(p)));
        }
        Eigen::VectorXd mm = gpu.mu(v), ss = gpu.sigma(v);
        for (int p = 0; p < P; ++p) {
            dmu = std::max(dmu, std::fabs(mm(p) - m2(p)));
            ds = std::max(ds, std::fabs(ss(p) - s2(p)));
        }
    }
    const bool dims = gpu.dim_in() == D && gpu.dim_out() == P && gpu.nb_samples() == N + 4 && (int)gpu.gp_models().size() == P;
    std::printf("MultiGP<limbo_b200::model::GP> N=%d D=%d P=%d dmu=%.3e dsigma2=%.3e dims_ok=%d\n", N + 4, D, P, dmu, ds, (int)dims);
    return (dmu < 1e-9 && ds < 1e-10 && dims) ? 0 : 1;
}

// The inner loop of bayes_opt::BOptimizer::optimize (boptimizer.hpp:139-170; BOptimizer itself needs Boost.Parameter, which
// this image lacks) with the batched acquisition optimiser: acquisition functor built per iteration, the optimiser only
// sees the closure f(x, gradient), add_sample after every evaluation.
template <template <typename, typename> class Acq>
int run_bo_loop(const char* name)
{
    using Kernel = kernel::MaternFiveHalves<Params>;
    using GP_t = limbo_b200::model::GP<Params, Kernel, mean::Data<Params>, model::gp::NoLFOpt<Params>>;
    using Acq_t = Acq<Params, GP_t>;
    unsigned long long seed = 99;
    const int D = 2;
    Eigen::VectorXd sol((Eigen::Index)D);
    sol(0) = 0.25; sol(1) = 0.75;
    auto sfun = [&](const Eigen::VectorXd& x) { Eigen::VectorXd y((Eigen::Index)1); y(0) = -(x - sol).squaredNorm(); return y; };
    std::vector<Eigen::VectorXd> S, O;
    for (int i = 0; i < 10; ++i) { // init::RandomSampling
        Eigen::VectorXd x((Eigen::Index)D);
        for (int d = 0; d < D; ++d) x(d) = u01(seed);
        S.push_back(x);
        O.push_back(sfun(x));
    }
    GP_t model(D, 1);
    model.compute(S, O);
    limbo_b200::opt::BatchedRandom<Params> acqui_optimizer;
    FirstElem afun;
    for (int it = 0; it < 30; ++it) {
        Acq_t acqui(model, it);
        auto acqui_optimization = [&](const Eigen::VectorXd& x, bool g) { return acqui(x, afun, g); };
        Eigen::VectorXd start((Eigen::Index)D);
        for (int d = 0; d < D; ++d) start(d) = u01(seed);
        Eigen::VectorXd nx = acqui_optimizer(acqui_optimization, start, true);
        S.push_back(nx);
        O.push_back(sfun(nx));
        model.add_sample(S.back(), O.back());
    }
    double best = -1e300;
    Eigen::VectorXd bx;
    for (size_t i = 0; i < O.size(); ++i)
        if (O[i](0) > best) { best = O[i](0); bx = S[i]; }
    const double err = (bx - sol).squaredNorm();
    std::printf("BO loop with %s + BatchedRandom: 10 + 30 evaluations, |x* - sol|^2 = %.3e\n", name, err);
    return err < 2e-3 ? 0 : 1; // test_boptimizer.cpp:202-281 accepts 1e-3 after 190 iterations of a tuned inner optimiser
}

// BatchedRandom over a batch-aware functor and over the reference's own one-point acqui::UCB must pick the same candidate.
struct ArgmaxProbe : limbo_b200::opt::BatchedRandom<Params> {
    template <typename F>
    static std::pair<double, long> run(const F& f, const std::vector<Eigen::VectorXd>& c) { return _argmax(f, c); }
};
int run_batch_equivalence()
{
    using Kernel = kernel::SquaredExpARD<Params>;
    using GP_t = limbo_b200::model::GP<Params, Kernel, mean::Data<Params>, model::gp::NoLFOpt<Params>>;
    unsigned long long seed = 7;
    const int N = 80, D = 3, M = 600;
    std::vector<Eigen::VectorXd> X, Y, C;
    for (int i = 0; i < N; ++i) {
        Eigen::VectorXd x((Eigen::Index)D), y((Eigen::Index)1);
        double s = 0;
        for (int d = 0; d < D; ++d) { x(d) = u01(seed); s += std::cos(3.0 * x(d)); }
        y(0) = s;
        X.push_back(x);
        Y.push_back(y);
    }
    for (int i = 0; i < M; ++i) {
        Eigen::VectorXd c((Eigen::Index)D);
        for (int d = 0; d < D; ++d) c(d) = u01(seed);
        C.push_back(c);
    }
    GP_t gp(D, 1);
    gp.compute(X, Y);
    FirstElem afun;
    int bad = 0;
    {
        limbo_b200::acqui::UCB<Params, GP_t> batch(gp);
        acqui::UCB<Params, GP_t> single(gp); // the reference's functor over the drop-in model: one query() per candidate
        auto fb = [&](const Eigen::VectorXd& x, bool g) { return batch(x, afun, g); };
        auto fs = [&](const Eigen::VectorXd& x, bool g) { return single(x, afun, g); };
        auto rb = ArgmaxProbe::run(fb, C), rs = ArgmaxProbe::run(fs, C);
        std::printf("UCB argmax over %d candidates: batched (%.12g, %ld) one-by-one (%.12g, %ld)\n", M, rb.first, rb.second, rs.first, rs.second);
        bad += !(rb.second == rs.second && std::fabs(rb.first - rs.first) < 1e-10);
    }
    {
        limbo_b200::acqui::EI<Params, GP_t> batch(gp);
        acqui::EI<Params, GP_t> single(gp);
        auto fb = [&](const Eigen::VectorXd& x, bool g) { return batch(x, afun, g); };
        auto fs = [&](const Eigen::VectorXd& x, bool g) { return single(x, afun, g); };
        auto rb = ArgmaxProbe::run(fb, C), rs = ArgmaxProbe::run(fs, C);
        std::printf("EI  argmax over %d candidates: batched (%.12g, %ld) one-by-one (%.12g, %ld)\n", M, rb.first, rb.second, rs.first, rs.second);
        bad += !(rb.second == rs.second && std::fabs(rb.first - rs.first) < 1e-10);
    }
    return bad;
}

int main()
{
    int bad = 0;
    bad += run_multigp();
    bad += run_batch_equivalence();
    bad += run_bo_loop<limbo_b200::acqui::UCB>("limbo_b200::acqui::UCB");
    bad += run_bo_loop<limbo_b200::acqui::EI>("limbo_b200::acqui::EI");
    bad += run_hp_case<model::gp::KernelLooOpt<Params, opt::Rprop<Params>>, mean::Data<Params>>("KernelLooOpt", 70, 2, true);
    bad += run_hp_case<model::gp::KernelMeanLFOpt<Params, opt::Rprop<Params>>, mean::FunctionARD<Params, mean::Constant<Params>>>(
        "KernelMeanLFOpt", 70, 2, false);
    bad += run_case<kernel::SquaredExpARD<Params>>("SquaredExpARD", 60, 3);
    bad += run_case<kernel::MaternFiveHalves<Params>>("MaternFiveHalves", 150, 2);
    bad += run_case<kernel::SquaredExpARD<ParamsLambda>, ParamsLambda>("SquaredExpARD_k2", 70, 3);
    std::printf(bad ? "DROPIN FAIL\n" : "DROPIN OK\n");
    return bad;
}
