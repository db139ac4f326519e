// eval_math.cuh -- Student-t tail used by the Spearman gate of the evaluation metrics (host + device).
//
// scipy.stats.spearmanr (called at reference utils/evaluation.py:166) reports the two-sided p-value of
// t = rho sqrt(dof / ((1+rho)(1-rho))) under Student's t with dof = n - 2.  That tail is the regularised
// incomplete beta function  p = I_x(dof/2, 1/2),  x = dof / (dof + t^2),  evaluated here in double with the
// modified-Lentz continued fraction (Numerical Recipes' betacf form), switching to the mirrored fraction
// 1 - I_{1-x}(1/2, dof/2) where the direct one converges slowly.  1 - x is formed without cancellation.
#pragma once

#include <math.h>

#ifdef __CUDACC__
#define ARVAE_HD __host__ __device__
#else
#define ARVAE_HD
#endif

namespace arvae {

// continued fraction of I_x(a, b); converges fast for x < (a + 1) / (a + b + 2)
ARVAE_HD static inline double beta_cf(double a, double b, double x) {
    const double tiny = 1e-300, eps = 1e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 20000; ++m) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) <= eps) break;
    }
    return h;
}

// log(Gamma(a + 1/2) / Gamma(a)): the difference of two lgamma values loses log10(a lgamma(a)) digits for large a
// (the evaluation set has dof/2 ~ 1e4), so from a = 64 the asymptotic series is used instead (next term 3e-17).
ARVAE_HD static inline double log_gamma_ratio_half(double a) {
    if (a < 64.0) return lgamma(a + 0.5) - lgamma(a);
    const double r = 1.0 / a, r2 = r * r;
    return 0.5 * log(a) - r * (0.125 - r2 * (1.0 / 192.0 - r2 * (1.0 / 640.0 - r2 * (17.0 / 14336.0))));
}

// P(|T_dof| >= |t|); NaN in -> NaN out, |t| = inf -> 0, t = 0 -> 1.  dof > 0.
ARVAE_HD static inline double student_t_two_sided(double t, double dof) {
    if (t != t || !(dof > 0.0)) return NAN;
    if (isinf(t)) return 0.0;
    if (t == 0.0) return 1.0;
    const double a = 0.5 * dof, b = 0.5;
    const double t2 = t * t;
    const double y = t2 / (dof + t2);           // 1 - x
    const double x = dof / (dof + t2);
    const double log_x = -log1p(t2 / dof);      // log x without cancellation near x = 1
    const double log_y = log(y);
    const double lbeta = log_gamma_ratio_half(a) - 0.57236494292470008707;  // - lgamma(1/2) = - log(sqrt(pi))
    const double front = exp(lbeta + a * log_x + b * log_y);
    double p;
    if (x < (a + 1.0) / (a + b + 2.0)) p = front * beta_cf(a, b, x) / a;
    else p = 1.0 - front * beta_cf(b, a, y) / b;
    return p < 0.0 ? 0.0 : (p > 1.0 ? 1.0 : p);
}

// t statistic of a correlation coefficient, as scipy forms it (division by zero -> inf, negatives clipped to 0)
ARVAE_HD static inline double correlation_t(double rho, double dof) {
    const double den = (rho + 1.0) * (1.0 - rho);
    double q = dof / den;                        // den == 0 -> +inf
    if (!(q > 0.0)) q = (q != q) ? q : 0.0;      // clip(0); NaN stays NaN
    return rho * sqrt(q);
}

}  // namespace arvae
