// ssme_b200/csrc/models/linear_gaussian_optimal.cuh -- the linear-Gaussian model of linear_gaussian.cuh filtered with the
// OPTIMAL proposal q(x_t | x_{t-1}, y_t) instead of the transition density: a model that uses the general SISR hooks
// (the reference's qSamp / logQEv / logFEv, liu_west_filter.h:1495-1516; pf's SISR filters) through the device model concept.
//   theta = (phi, sigma, tau);  x_t | x_{t-1} ~ N(phi x_{t-1}, sigma^2),  y_t | x_t ~ N(x_t, tau^2),  x_1 ~ N(0, p0), p0 = sigma^2 / (1 - phi^2)
//   q(x_t | x_{t-1}, y_t) = N(ax x_{t-1} + ay y_t, s^2),  s^2 = 1 / (1/sigma^2 + 1/tau^2),  ax = s^2 phi / sigma^2,  ay = s^2 / tau^2
//   log g + log f - log q = log N(y_t; phi x_{t-1}, sigma^2 + tau^2)      (depends on the OLD state only: a fully adapted filter)
//   time 1: q1 = N(ay0 y_1, s0^2), s0^2 = 1 / (1/p0 + 1/tau^2), ay0 = s0^2 / tau^2;  log mu + log g - log q1 = log N(y_1; 0, p0 + tau^2)
// Its exact likelihood is the Kalman filter's (tests/test_gpu_kalman.py); the estimator's variance is far below the
// bootstrap filter's on the same model.
#pragma once
#include "model_api.cuh"

namespace ssme {

struct LinearGaussianOptimalModel {
    static constexpr int kId = 3;  // SSME_B200_MODEL_LINEAR_GAUSSIAN_OPTIMAL
    static constexpr int kNumParams = 3;
    static constexpr int kObsStride = 1;
    static constexpr bool kHasF32 = false;

    struct Params {
        double phi, s, ax, ay, hw, cw, s0, ay0, hw0, cw0;
    };
    struct Step {
        double y;
    };
    static __device__ __forceinline__ Params init(const double* th)
    {
        Params m;
        m.phi = th[0];
        const double sig2 = __dmul_rn(th[1], th[1]), tau2 = __dmul_rn(th[2], th[2]);
        const double s2 = __ddiv_rn(1.0, __dadd_rn(__ddiv_rn(1.0, sig2), __ddiv_rn(1.0, tau2)));
        m.s = __dsqrt_rn(s2);
        m.ax = __ddiv_rn(__dmul_rn(s2, m.phi), sig2);
        m.ay = __ddiv_rn(s2, tau2);
        const double v = __dadd_rn(sig2, tau2);
        m.hw = __ddiv_rn(0.5, v);
        m.cw = __dsub_rn(__dmul_rn(-0.5, dlog(v)), SSME_DM_HALF_LOG_2PI);
        const double p0 = __ddiv_rn(sig2, __dsub_rn(1.0, __dmul_rn(m.phi, m.phi)));
        const double s02 = __ddiv_rn(1.0, __dadd_rn(__ddiv_rn(1.0, p0), __ddiv_rn(1.0, tau2)));
        m.s0 = __dsqrt_rn(s02);
        m.ay0 = __ddiv_rn(s02, tau2);
        const double v0 = __dadd_rn(p0, tau2);
        m.hw0 = __ddiv_rn(0.5, v0);
        m.cw0 = __dsub_rn(__dmul_rn(-0.5, dlog(v0)), SSME_DM_HALF_LOG_2PI);
        return m;
    }
    static __device__ __forceinline__ Step step(const Params&, const double* row) { return Step{row[0]}; }
    // qSamp at time 1 and afterwards: the proposal sees the observation
    static __device__ __forceinline__ double q1(const Params& m, const Step& s, double z) { return __fma_rn(m.s0, z, __dmul_rn(m.ay0, s.y)); }
    static __device__ __forceinline__ double f(const Params& m, const Step& s, double x, double z)
    {
        return __fma_rn(m.s, z, __fma_rn(m.ax, x, __dmul_rn(m.ay, s.y)));
    }
    // the observation density itself (not used by the filter: the weights below already contain it)
    static __device__ __forceinline__ double logg(const Params& m, const Step& s, double x) { return logw(m, s, x, x); }
    static __device__ __forceinline__ double logw(const Params& m, const Step& s, double, double x_prev)
    {
        const double d = __dsub_rn(s.y, __dmul_rn(m.phi, x_prev));
        return __fma_rn(-m.hw, __dmul_rn(d, d), m.cw);
    }
    static __device__ __forceinline__ double logw1(const Params& m, const Step& s, double)
    {
        return __fma_rn(-m.hw0, __dmul_rn(s.y, s.y), m.cw0);
    }
};

}  // namespace ssme
