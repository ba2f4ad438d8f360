// ssme_b200/csrc/models/linear_gaussian.cuh -- AR(1) state observed in Gaussian noise.
//   theta = (phi, sigma, tau):  x_1 = z sigma / sqrt(1 - phi^2);  x_t = phi x_{t-1} + sigma z;  y_t | x_t ~ N(x_t, tau^2).
// Not a model of the reference: it is here because its exact log-likelihood is known (Kalman filter), which gives the
// one parity check that depends on neither oracle (tests/test_gpu_kalman.py), and because it was added through the
// model concept alone -- this header, one line in models.cuh, one id in include/ssme_b200.h -- without touching a kernel.
// Canonical arithmetic: log g = fma(-h, d*d, c0), d = y - x, h = 1/2 / tau^2, c0 = -log tau - 1/2 log 2pi.
#pragma once
#include "model_api.cuh"

namespace ssme {

struct LinearGaussianModel {
    static constexpr int kId = 2;  // SSME_B200_MODEL_LINEAR_GAUSSIAN
    static constexpr int kNumParams = 3;
    static constexpr int kObsStride = 1;
    static constexpr bool kHasF32 = false;

    struct Params {
        double phi, sigma, sd0, c0, h;
    };
    struct Step {
        double y;
    };
    static __device__ __forceinline__ Params init(const double* th)
    {
        Params m;
        m.phi = th[0];
        m.sigma = th[1];
        const double tau = th[2];
        m.sd0 = __ddiv_rn(m.sigma, __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(m.phi, m.phi))));
        m.c0 = __dsub_rn(-dlog(tau), SSME_DM_HALF_LOG_2PI);
        m.h = __ddiv_rn(0.5, __dmul_rn(tau, tau));
        return m;
    }
    static __device__ __forceinline__ Step step(const Params&, const double* row) { return Step{row[0]}; }
    static __device__ __forceinline__ double q1(const Params& m, const Step&, double z) { return __dmul_rn(z, m.sd0); }
    static __device__ __forceinline__ double f(const Params& m, const Step&, double x, double z)
    {
        return __fma_rn(m.phi, x, __dmul_rn(m.sigma, z));
    }
    static __device__ __forceinline__ double logg(const Params& m, const Step& s, double x)
    {
        const double d = __dsub_rn(s.y, x);
        return __fma_rn(-m.h, __dmul_rn(d, d), m.c0);
    }
};

}  // namespace ssme
