// ssme_b200/csrc/models/sv.cuh -- univariate stochastic volatility, the reference's example model svol_bs
// (example/univ_svol_bootstrap_filter.h):  theta = (beta, phi, sigma^2) :54-61;  x_1 = z sigma / sqrt(1 - phi^2) :68;
// x_t = phi x_{t-1} + sigma z :77;  y_t | x_t ~ N(0, (beta e^{x_t/2})^2) :85.
// Canonical arithmetic: log g = fma(-h, exp(-x), fma(-1/2, x, c0)), c0 = -log beta - 1/2 log 2pi,
// h = y^2 (1/2 / beta^2) -- one exp, no log, no divide per particle.
#pragma once
#include "model_api.cuh"

namespace ssme {

struct SvModel {
    static constexpr int kId = 0;  // SSME_B200_MODEL_SV
    static constexpr int kNumParams = 3;
    static constexpr int kObsStride = 1;
    static constexpr bool kHasF32 = true;

    struct Params {
        double phi, sigma, sd0, c0, inv2b2;
    };
    struct Step {
        double h;
    };
    static __device__ __forceinline__ Params init(const double* th)
    {
        Params m;
        const double beta = th[0];
        m.phi = th[1];
        m.sigma = __dsqrt_rn(th[2]);
        m.sd0 = __ddiv_rn(m.sigma, __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(m.phi, m.phi))));
        m.c0 = __dsub_rn(-dlog(beta), SSME_DM_HALF_LOG_2PI);
        m.inv2b2 = __ddiv_rn(0.5, __dmul_rn(beta, beta));
        return m;
    }
    static __device__ __forceinline__ Step step(const Params& m, const double* row)
    {
        const double y = row[0];
        return Step{__dmul_rn(__dmul_rn(y, y), m.inv2b2)};
    }
    static __device__ __forceinline__ double q1(const Params& m, const Step&, double z) { return __dmul_rn(z, m.sd0); }
    static __device__ __forceinline__ double f(const Params& m, const Step&, double x, double z)
    {
        return __fma_rn(m.phi, x, __dmul_rn(m.sigma, z));
    }
    static __device__ __forceinline__ double logg(const Params& m, const Step& s, double x)
    {
        const double e = dexp(-x);
        return __fma_rn(-s.h, e, __fma_rn(-0.5, x, m.c0));
    }

    // ---- fp32 mode: per-filter constants formed in double and rounded once ----
    struct ParamsF {
        float phi, sigma, sd0, c0;
        double inv2b2;
    };
    struct StepF {
        float h;
    };
    static __device__ __forceinline__ ParamsF init_f32(const double* th)
    {
        const Params m = init(th);
        return ParamsF{(float)m.phi, (float)m.sigma, (float)m.sd0, (float)m.c0, m.inv2b2};
    }
    static __device__ __forceinline__ StepF step_f32(const ParamsF& m, const double* row)
    {
        const double y = row[0];
        return StepF{(float)__dmul_rn(__dmul_rn(y, y), m.inv2b2)};
    }
    static __device__ __forceinline__ float q1_f32(const ParamsF& m, const StepF&, float z) { return __fmul_rn(z, m.sd0); }
    static __device__ __forceinline__ float f_f32(const ParamsF& m, const StepF&, float x, float z)
    {
        return __fmaf_rn(m.phi, x, __fmul_rn(m.sigma, z));
    }
    static __device__ __forceinline__ float logg_f32(const ParamsF& m, const StepF& s, float x)
    {
        const float e = fexp(-x);
        return __fmaf_rn(-s.h, e, __fmaf_rn(-0.5f, x, m.c0));
    }
};

}  // namespace ssme
