// ssme_b200/csrc/models/sv_leverage.cuh -- stochastic volatility with leverage, the model of the reference's Liu-West
// and swarm tests (test/test_liu_west.cpp:83-157, test/test_pswarm.cpp:33-141): theta = (phi, mu, sigma, rho);
// x_1 = z sigma / sqrt(1 - phi^2);  x_t = mu + phi (x_{t-1} - mu) + rho sigma z_t e^{-x_{t-1}/2} + sigma sqrt(1 - rho^2) z,
// covariate z_t = y_{t-1} (second column of the observation row);  y_t | x_t ~ N(0, e^{x_t}).
// Innovation sd sigma sqrt(1 - rho^2) as in the Liu-West tests.
#pragma once
#include "model_api.cuh"

namespace ssme {

struct SvLeverageModel {
    static constexpr int kId = 1;  // SSME_B200_MODEL_SV_LEVERAGE
    static constexpr int kNumParams = 4;
    static constexpr int kObsStride = 2;
    static constexpr bool kHasF32 = true;

    struct Params {
        double phi, mu, sd0, c0, rho_sigma, sdv;
    };
    struct Step {
        double h, cz;
    };
    static __device__ __forceinline__ Params init(const double* th)
    {
        Params m;
        m.phi = th[0];
        m.mu = th[1];
        const double sigma = th[2], rho = th[3];
        m.sd0 = __ddiv_rn(sigma, __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(m.phi, m.phi))));
        m.c0 = -SSME_DM_HALF_LOG_2PI;  // -log(beta) - 1/2 log 2pi with beta = 1
        m.rho_sigma = __dmul_rn(rho, sigma);
        m.sdv = __dmul_rn(sigma, __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(rho, rho))));
        return m;
    }
    static __device__ __forceinline__ Step step(const Params& m, const double* row)
    {
        const double y = row[0];
        // inv2b2 = 0.5 / (1 * 1) = 0.5 exactly (beta = 1)
        return Step{__dmul_rn(__dmul_rn(y, y), 0.5), __dmul_rn(m.rho_sigma, row[1])};
    }
    static __device__ __forceinline__ double q1(const Params& m, const Step&, double z) { return __dmul_rn(z, m.sd0); }
    static __device__ __forceinline__ double f(const Params& m, const Step& s, double x, double z)
    {
        const double e2 = dexp(__dmul_rn(-0.5, x));
        double mean = __fma_rn(m.phi, __dsub_rn(x, m.mu), m.mu);
        mean = __fma_rn(s.cz, e2, mean);
        return __fma_rn(m.sdv, z, mean);
    }
    static __device__ __forceinline__ double logg(const Params& m, const Step& s, double x)
    {
        const double e = dexp(-x);
        return __fma_rn(-s.h, e, __fma_rn(-0.5, x, m.c0));
    }

    struct ParamsF {
        float phi, mu, sd0, c0, sdv;
        double rho_sigma;
    };
    struct StepF {
        float h, cz;
    };
    static __device__ __forceinline__ ParamsF init_f32(const double* th)
    {
        const Params m = init(th);
        return ParamsF{(float)m.phi, (float)m.mu, (float)m.sd0, (float)m.c0, (float)m.sdv, m.rho_sigma};
    }
    static __device__ __forceinline__ StepF step_f32(const ParamsF& m, const double* row)
    {
        const double y = row[0];
        return StepF{(float)__dmul_rn(__dmul_rn(y, y), 0.5), (float)__dmul_rn(m.rho_sigma, row[1])};
    }
    static __device__ __forceinline__ float q1_f32(const ParamsF& m, const StepF&, float z) { return __fmul_rn(z, m.sd0); }
    static __device__ __forceinline__ float f_f32(const ParamsF& m, const StepF& s, float x, float z)
    {
        const float e2 = fexp(__fmul_rn(-0.5f, x));
        float mean = __fmaf_rn(m.phi, __fsub_rn(x, m.mu), m.mu);
        mean = __fmaf_rn(s.cz, e2, mean);
        return __fmaf_rn(m.sdv, z, mean);
    }
    static __device__ __forceinline__ float logg_f32(const ParamsF& m, const StepF& s, float x)
    {
        const float e = fexp(-x);
        return __fmaf_rn(-s.h, e, __fmaf_rn(-0.5f, x, m.c0));
    }
};

}  // namespace ssme
