// ssme_b200/csrc/models/models.cuh -- the list of device models the library is built with (see model_api.cuh).
// Adding a model: write models/<name>.cuh, include it here and add one X(...) line; give it an id in include/ssme_b200.h.
#pragma once
#include "linear_gaussian.cuh"
#include "linear_gaussian_optimal.cuh"
#include "sv.cuh"
#include "sv_leverage.cuh"
#include "sv_volatility.cuh"

// X(ModelType): expanded where kernels are instantiated (pf_inst.cu, cluster and global-memory launchers in capi.cu /
// spill_capi.cu) and where the C ABI validates a configuration (model_info in capi.cu).
#define SSME_FOR_EACH_MODEL(X) \
    X(SvModel)                 \
    X(SvLeverageModel)         \
    X(LinearGaussianModel)     \
    X(LinearGaussianOptimalModel) \
    X(SvVolatilityModel)

namespace ssme {

struct ModelInfo {
    int id, num_params, obs_stride;
    bool has_f32;
    int num_expect;  // expectation functions E[h_k(x_t) | y_{1:t}] the tracing kernel forms (model_api.cuh)
};

// host-side description of model `id`; returns false for an unknown id
inline bool model_info(int id, ModelInfo* out)
{
#define SSME_MODEL_INFO_CASE(M)                                            \
    if (id == M::kId) {                                                    \
        *out = ModelInfo{M::kId, M::kNumParams, M::kObsStride, M::kHasF32, model_num_expect<M>::value}; \
        return true;                                                       \
    }
    SSME_FOR_EACH_MODEL(SSME_MODEL_INFO_CASE)
#undef SSME_MODEL_INFO_CASE
    return false;
}

}  // namespace ssme
