// oracle/refshim/pf/resamplers.h -- TEST INFRASTRUCTURE.  Stand-in for the two pf resamplers the reference names:
//   mn_resampler     example/estimate_univ_svol.h:119     multinomial via std::discrete_distribution
//   mn_resamp_fast1  test/test_pswarm.cpp:149,157          the sorted-uniform algorithm; its first-party twin in the
//                    reference tree is mn_resamp_states_and_params (include/ssme/liu_west_filter.h:91-145), followed here
// Both copy the particle array and reset the log-weights to zero.
#ifndef SSME_REFSHIM_PF_RESAMPLERS_H
#define SSME_REFSHIM_PF_RESAMPLERS_H
#include <Eigen/Dense>
#include <algorithm>
#include <array>
#include <cmath>
#include <random>

#include "shim_streams.h"

namespace pf {
namespace resamplers {

template <size_t nparts, size_t dimx, typename float_t>
class rbase {
public:
    using ssv = Eigen::Matrix<float_t, dimx, 1>;
    using arrayVec = std::array<ssv, nparts>;
    using arrayFloat = std::array<float_t, nparts>;
    rbase() : m_gen{shim::next_seed()} {}

protected:
    std::mt19937 m_gen;
};

template <size_t nparts, size_t dimx, typename float_t>
class mn_resampler : public rbase<nparts, dimx, float_t> {
public:
    using typename rbase<nparts, dimx, float_t>::arrayVec;
    using typename rbase<nparts, dimx, float_t>::arrayFloat;
    void resampLogWts(arrayVec& oldParts, arrayFloat& oldLogUnNormWts)
    {
        arrayFloat w;
        const float_t m = *std::max_element(oldLogUnNormWts.begin(), oldLogUnNormWts.end());
        std::transform(oldLogUnNormWts.begin(), oldLogUnNormWts.end(), w.begin(), [&m](float_t& d) -> float_t { return std::exp(d - m); });
        arrayVec tmpPartics = oldParts;
        std::discrete_distribution<> idxSampler(w.begin(), w.end());
        if (shim::resamp_stream().active()) {
            shim::playback_engine eng(shim::resamp_stream());
            for (size_t i = 0; i < nparts; ++i) tmpPartics[i] = oldParts[idxSampler(eng)];
        } else {
            for (size_t i = 0; i < nparts; ++i) tmpPartics[i] = oldParts[idxSampler(this->m_gen)];
        }
        oldParts = std::move(tmpPartics);
        std::fill(oldLogUnNormWts.begin(), oldLogUnNormWts.end(), 0.0);
    }
};

template <size_t nparts, size_t dimx, typename float_t>
class mn_resamp_fast1 : public rbase<nparts, dimx, float_t> {
public:
    using typename rbase<nparts, dimx, float_t>::arrayVec;
    using typename rbase<nparts, dimx, float_t>::arrayFloat;
    mn_resamp_fast1() : m_u_sampler(0.0, 1.0) {}
    void resampLogWts(arrayVec& oldParts, arrayFloat& oldLogUnNormWts)
    {
        // liu_west_filter.h:96-139 with the parameter array dropped
        arrayFloat w, e;
        const float_t m = *std::max_element(oldLogUnNormWts.begin(), oldLogUnNormWts.end());
        std::transform(oldLogUnNormWts.begin(), oldLogUnNormWts.end(), w.begin(), [&m](float_t& d) -> float_t { return std::exp(d - m); });
        float_t norm(0.0), G(0.0);
        for (size_t i = 0; i < nparts; ++i) {
            norm += w[i];
            e[i] = -std::log(unif());
            G += e[i];
        }
        G -= std::log(unif());
        arrayVec tmpPartics = oldParts;
        float_t ustat(0.0), running(w[0] / norm), one_less(0.0);
        unsigned int idx = 0;
        for (size_t i = 0; i < nparts; ++i) {
            ustat += e[i] / G;
            while (!((one_less < ustat) && (ustat <= running))) {
                idx++;
                running += w[idx] / norm;
                one_less += w[idx - 1] / norm;
            }
            tmpPartics[i] = oldParts[idx];
        }
        oldParts = std::move(tmpPartics);
        std::fill(oldLogUnNormWts.begin(), oldLogUnNormWts.end(), 0.0);
    }

private:
    float_t unif()
    {
        if (shim::resamp_stream().active()) return static_cast<float_t>(shim::resamp_stream().next());
        return m_u_sampler(this->m_gen);
    }
    std::uniform_real_distribution<float_t> m_u_sampler;
};

}  // namespace resamplers
}  // namespace pf
#endif
