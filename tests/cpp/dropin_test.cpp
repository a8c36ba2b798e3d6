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
#include <limbo/opt/rprop.hpp>

#include <limbo_b200/model/gp.hpp>

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

int main()
{
    int bad = 0;
    bad += run_hp_case<model::gp::KernelLooOpt<Params, opt::Rprop<Params>>, mean::Data<Params>>("KernelLooOpt", 70, 2, true);
    bad += run_hp_case<model::gp::KernelMeanLFOpt<Params, opt::Rprop<Params>>, mean::FunctionARD<Params, mean::Constant<Params>>>(
        "KernelMeanLFOpt", 70, 2, false);
    bad += run_case<kernel::SquaredExpARD<Params>>("SquaredExpARD", 60, 3);
    bad += run_case<kernel::MaternFiveHalves<Params>>("MaternFiveHalves", 150, 2);
    bad += run_case<kernel::SquaredExpARD<ParamsLambda>, ParamsLambda>("SquaredExpARD_k2", 70, 3);
    std::printf(bad ? "DROPIN FAIL\n" : "DROPIN OK\n");
    return bad;
}
