// ssme_b200/csrc/models/model_api.cuh -- the DEVICE MODEL CONCEPT: what a state-space model supplies to the kernels.
//
// The reference's API is "subclass the filter and supply the densities" (example/univ_svol_bootstrap_filter.h:37-41:
// logQ1Ev, logMuEv, logGEv, fSamp, q1Samp; include/ssme/liu_west_filter.h:1456-1522 for the Liu-West family).  Host virtuals
// cannot run on the device, so here a model is a TYPE with static device functions, and every filter kernel (K1
// pf_kernel.cuh, K1f pf_kernel_f32.cuh, K2 cluster_kernel.cuh, K3 spill_kernel.cuh) is a template over that type.
// The kernels contain no model-specific code.
//
//   struct MyModel {
//       static constexpr int kId;         // value of ssme_b200_config::model that selects it (include/ssme_b200.h)
//       static constexpr int kNumParams;  // length of one untransformed theta row (param::pack::get_untrans_params)
//       static constexpr int kObsStride;  // doubles per observation row: y_t [, covariate z_t, ...]
//       static constexpr bool kHasF32;    // also provides the float hooks below (fp32 mode, K1f)
//       struct Params;                    // per-filter constants derived once from theta        (the model's constructor)
//       struct Step;                      // per-time-step quantities shared by all particles    (functions of the observation row)
//       static __device__ Params init(const double* theta);
//       static __device__ Step   step(const Params&, const double* obs_row);
//       static __device__ double q1  (const Params&, const Step&, double z);            // x_1 from one N(0,1) draw      (q1Samp)
//       static __device__ double f   (const Params&, const Step&, double x, double z);  // x_t | x_{t-1} from one draw   (fSamp)
//       static __device__ double logg(const Params&, const Step&, double x);            // log g(y_t | x_t)              (logGEv)
//       // kHasF32: ParamsF, StepF, init_f32, step_f32, q1_f32, f_f32, logg_f32 with float states
//       // OPTIONAL, for a proposal other than the transition density (general SISR, the qSamp / logQEv / logFEv hooks of
//       // liu_west_filter.h:1495-1516 and of pf's SISR filters): f and q1 then SAMPLE THE PROPOSAL (they see the observation
//       // through Step), and the model supplies the incremental log-weights
//       static __device__ double logw (const Params&, const Step&, double x, double x_prev);  // log g + log f - log q
//       static __device__ double logw1(const Params&, const Step&, double x);                 // log mu + log g - log q1
//       // OPTIONAL, expectation functions E[h_k(x_t) | y_{1:t}] formed before resampling -- the device counterpart of the
//       // std::function callbacks the reference's filter() / Swarm take (pswarm_filter.h:47, 340; liu_west_filter.h:1662-1683).
//       // Without them the two built-in functions h = x, x^2 are used.
//       static constexpr int kNumExpect;                                                        // 1 .. 8
//       static __device__ double expect_fn(const Params&, const Step&, double x, int k);        // h_k(x_t), k < kNumExpect
//   };
//
// Contract.  Without logw / logw1 the bootstrap proposal is meant (q = f, q1 = mu), so the weight is logg alone: logMuEv -
// logQ1Ev cancels exactly (univ_svol_bootstrap_filter.h:92-95 vs :102).  models/linear_gaussian_optimal.cuh is a model with
// the optimal proposal p(x_t | x_{t-1}, y_t).  One N(0,1) variate per particle per step (scalar state).
// Every arithmetic operation must be an explicit round-to-nearest intrinsic (__fma_rn, __dmul_rn, ...) or a det_math.cuh
// function, because the same sequence is restated in oracle/pf_oracle.c (can_q1 / can_f / can_logg) and the two are tested
// bit for bit.
//
// Adding a model = one header in this directory + one line in models.cuh (SSME_FOR_EACH_MODEL) + its id in
// include/ssme_b200.h; plus, for the parity tests, its restatement in oracle/pf_oracle.c.  No kernel changes
// (models/linear_gaussian.cuh was added that way).
#pragma once
#include <type_traits>

#include "../det_math.cuh"

namespace ssme {

template <typename M, typename = void>
struct model_has_logw : std::false_type {};
template <typename M>
struct model_has_logw<M, std::void_t<decltype(&M::logw), decltype(&M::logw1)>> : std::true_type {};

template <typename M, typename = void>
struct model_has_expect : std::false_type {};
template <typename M>
struct model_has_expect<M, std::void_t<decltype(&M::expect_fn), decltype(M::kNumExpect)>> : std::true_type {};
// number of expectation functions of a model (2 built-in ones, h = x and x^2, unless the model brings its own)
template <typename M, bool = model_has_expect<M>::value>
struct model_num_expect {
    static constexpr int value = 2;
};
template <typename M>
struct model_num_expect<M, true> {
    static constexpr int value = M::kNumExpect;
    static_assert(M::kNumExpect >= 1 && M::kNumExpect <= 8, "1 .. 8 expectation functions");
};

// incremental log-weight of a particle that moved from x_prev to x (first: the time-1 draw): log g for a bootstrap model
// (identical code to calling logg directly), log g + log f - log q for a model that brings its own proposal
template <typename M>
__device__ __forceinline__ double model_log_weight(const typename M::Params& mc, const typename M::Step& ms, double x, double x_prev, bool first)
{
    if constexpr (model_has_logw<M>::value) return first ? M::logw1(mc, ms, x) : M::logw(mc, ms, x, x_prev);
    else return M::logg(mc, ms, x);
}

}  // namespace ssme
