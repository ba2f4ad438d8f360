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
//   };
//
// Contract.  The bootstrap proposal is built in (q = f, q1 = mu), so the time-1 weight is logg alone: logMuEv - logQ1Ev
// cancels exactly (univ_svol_bootstrap_filter.h:92-95 vs :102).  One N(0,1) variate per particle per step (scalar state).
// Every arithmetic operation must be an explicit round-to-nearest intrinsic (__fma_rn, __dmul_rn, ...) or a det_math.cuh
// function, because the same sequence is restated in oracle/pf_oracle.c (can_q1 / can_f / can_logg) and the two are tested
// bit for bit.
//
// Adding a model = one header in this directory + one line in models.cuh (SSME_FOR_EACH_MODEL) + its id in
// include/ssme_b200.h; plus, for the parity tests, its restatement in oracle/pf_oracle.c.  No kernel changes
// (models/linear_gaussian.cuh was added that way).
#pragma once
#include "../det_math.cuh"
