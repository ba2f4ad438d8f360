// ssme_b200/csrc/models/sv_volatility.cuh -- the SV model of sv.cuh with EXPECTATION FUNCTIONS OF ITS OWN.
//
// The reference's filters and its Swarm take a vector of std::function callbacks h(x_t) and return E[h(x_t) | y_{1:t}]
// (pswarm_filter.h:47, 340-349; in-tree twin liu_west_filter.h:1662-1683).  Host callbacks cannot run inside a kernel; on the
// device they are two more members of the model type (model_api.cuh: kNumExpect, expect_fn).  This model asks for three:
// the filtering mean and second moment of the log-volatility and the volatility scale exp(x_t / 2) itself -- what a user of the
// SV example reads off a filter.  Added like linear_gaussian.cuh: this header, one line in models.cuh, one id; no kernel changed.
#pragma once
#include "sv.cuh"

namespace ssme {

struct SvVolatilityModel : SvModel {
    static constexpr int kId = 4;  // SSME_B200_MODEL_SV_VOLATILITY
    static constexpr bool kHasF32 = false;
    static constexpr int kNumExpect = 3;
    static __device__ __forceinline__ double expect_fn(const Params&, const Step&, double x, int k)
    {
        if (k == 0) return x;
        if (k == 1) return __dmul_rn(x, x);
        return dexp(__dmul_rn(0.5, x));
    }
};

}  // namespace ssme
