// Host-side accuracy harness for dlsm::fast_log1pexp (the same __host__ __device__ source the
// kernels inline).  Prints the max abs / rel error against long double log1pl(expl(x)).
#include "../../dynetlsm_b200/csrc/dlsm_device.cuh"
#include <cmath>
#include <cstdio>
#include <random>

static long double ref(long double x)
{
    return (x > 0 ? x : 0.0L) + log1pl(expl(-fabsl(x)));
}

int main()
{
    double max_abs = 0, max_rel = 0, worst = 0, max_abs_neg = 0;
    std::mt19937_64 g(1);
    std::uniform_real_distribution<double> U(-40.0, 40.0);
    auto probe = [&](double x) {
        const double got = dlsm::fast_log1pexp(x);
        const long double want = ref((long double)x);
        const double ae = (double)fabsl((long double)got - want);
        const double re = (want > 1e-290L) ? (double)(ae / fabsl(want)) : 0.0; // below: clamped tail
        if (ae > max_abs) { max_abs = ae; worst = x; }
        if (re > max_rel && x > -15.0) max_rel = re;   // relative accuracy is only claimed where the value is > ~3e-7
        if (x <= 0 && ae > max_abs_neg) max_abs_neg = ae;
    };
    for (int k = 0; k < 4000000; k++) probe(U(g));
    for (int k = -4000; k <= 4000; k++) probe(k * 0.01);             // grid incl. 0 and the +-36 seam
    for (int k = 0; k < 2000; k++) { probe(std::ldexp(1.0, -k % 60)); probe(-std::ldexp(1.0, -k % 60)); }
    probe(36.0); probe(-36.0); probe(36.0000001); probe(-36.0000001); probe(700.0); probe(-700.0); probe(100.0); probe(-100.0); probe(-300.0); probe(1e6); probe(-1e6);
    const int special_ok = std::isnan(dlsm::fast_log1pexp(NAN)) && 
                           dlsm::fast_log1pexp(-1e9) < 1e-290 && dlsm::fast_log1pexp(-1e9) >= 0.0 && dlsm::fast_log1pexp(1e9) == 1e9 && dlsm::fast_log1pexp(0.0) == std::log(2.0);
    // the fused Bernoulli term against its definition
    double max_term = 0;
    for (int k = 0; k < 200000; k++) {
        const double x = U(g) * 0.5;
        for (int yb = 0; yb < 2; yb++) {
            const long double want = (long double)yb * x - ref((long double)x);
            const double got = dlsm::logit_term(yb - 0.5, x);
            const double ae = (double)fabsl((long double)got - want);
            if (ae > max_term) max_term = ae;
        }
    }
    // fast_exp against expl
    double max_exp_rel = 0;
    std::uniform_real_distribution<double> V(-700.0, 700.0);
    for (int k = 0; k < 2000000; k++) {
        const double x = (k & 1) ? V(g) : U(g);
        const long double want = expl((long double)x);
        const double re = (double)(fabsl((long double)dlsm::fast_exp(x) - want) / want);
        if (re > max_exp_rel) max_exp_rel = re;
    }
    const int exp_special = dlsm::fast_exp(-800.0) == 0.0 && std::isinf(dlsm::fast_exp(720.0)) &&
                            std::isnan(dlsm::fast_exp(NAN)) && dlsm::fast_exp(0.0) == 1.0;
    printf("%.3e %.3e %.6f %d %.3e %.3e %d %.3e\n", max_abs, max_rel, worst, special_ok, max_term,
           max_exp_rel, exp_special, max_abs_neg);
    return 0;
}
