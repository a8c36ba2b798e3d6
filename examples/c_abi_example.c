/* examples/c_abi_example.c — the C ABI of include/limbo_b200.h used from plain C (what a cgo / JNI / ctypes binding calls).
 * Fits a GP on N points of f(x) = cos(3 (x0 + x1)), predicts M candidates, runs the fused UCB argmax and prints the numbers
 * tests/test_gpu_c_example.py compares with the CPU oracle on the same inputs.
 *   gcc -std=c99 -Iinclude examples/c_abi_example.c -Llimbo_b200/lib -llimbo_b200 -lm -o c_abi_example            */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "limbo_b200.h"

static double u01(unsigned long long* s)
{ /* splitmix64, the generator of limbo_b200/synth.py */
    unsigned long long z;
    *s += 0x9E3779B97F4A7C15ULL;
    z = *s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

#define CHECK(call)                                                                         \
    do {                                                                                    \
        int rc_ = (call);                                                                   \
        if (rc_ != LB_OK) {                                                                 \
            fprintf(stderr, "%s -> %d (%s) %s\n", #call, rc_, lb_strerror(rc_), lb_last_cuda_error()); \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

int main(void)
{
    enum { N = 300, D = 2, M = 1000 };
    static double X[N * D], obs_mean[N], Xq[M * D], mu[M], s2[M], acq[M];
    unsigned long long seed = 2024;
    double ymean = 0.0, loglik = 0.0, best = 0.0;
    double hp[D + 1] = {-0.5, -0.3, 0.1}; /* log ell_0, log ell_1, log sigma_f (kernel/squared_exp_ard.hpp:96-105) */
    double ucb_params[2] = {0.5, 0.0};     /* acqui/ucb.hpp:58 */
    int64_t idx = -1;
    lb_gp* gp = NULL;
    int i;

    for (i = 0; i < N; ++i) {
        X[i * D] = u01(&seed);
        X[i * D + 1] = u01(&seed);
        obs_mean[i] = cos(3.0 * (X[i * D] + X[i * D + 1]));
        ymean += obs_mean[i] / N;
    }
    for (i = 0; i < N; ++i) obs_mean[i] -= ymean; /* mean::Data stays with the caller (gp.hpp:547) */
    for (i = 0; i < M * D; ++i) Xq[i] = u01(&seed);

    CHECK(lb_create(&gp, 0, LB_PREC_FP64));
    CHECK(lb_set_data(gp, N, D, 1, X, obs_mean));
    CHECK(lb_set_kernel(gp, LB_KERNEL_SQUARED_EXP_ARD, hp, D + 1, 0.01));
    CHECK(lb_fit(gp));
    CHECK(lb_log_lik(gp, &loglik));
    CHECK(lb_query(gp, M, Xq, mu, s2));
    CHECK(lb_acq_argmax(gp, LB_ACQ_UCB, ucb_params, M, Xq, NULL, ymean, acq, &best, &idx));
    printf("loglik %.17g\n", loglik);
    printf("mu0 %.17g sigma2_0 %.17g\n", mu[0] + ymean, s2[0]);
    printf("mu_last %.17g sigma2_last %.17g\n", mu[M - 1] + ymean, s2[M - 1]);
    printf("best %.17g idx %lld n %lld launches %lld\n", best, (long long)idx, (long long)lb_nb_samples(gp), lb_launch_count(gp));
    CHECK(lb_destroy(gp));
    return 0;
}
