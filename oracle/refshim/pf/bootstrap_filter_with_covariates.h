// oracle/refshim/pf/bootstrap_filter_with_covariates.h -- TEST INFRASTRUCTURE.  Stand-in for pf::filters::BSFilterWC
// (test/test_pswarm.cpp:4,33): BSFilter whose model hooks also see a covariate z_t.  Restated as in bootstrap_filter.h
// from the reference's first-party twin LWFilter2WithCovs::filter (include/ssme/liu_west_filter.h:2191-2343).
#ifndef SSME_REFSHIM_PF_BOOTSTRAP_FILTER_WC_H
#define SSME_REFSHIM_PF_BOOTSTRAP_FILTER_WC_H
#include <Eigen/Dense>
#include <algorithm>
#include <array>
#include <cmath>
#include <limits>
#include <vector>

#include "pf_base.h"

namespace pf {
namespace filters {

template <size_t nparts, size_t dimx, size_t dimy, size_t dimcov, typename resamp_t, typename float_t, bool debug = false>
class BSFilterWC : public bases::pf_withcov_base<float_t, dimy, dimx, dimcov> {
protected:
    using ssv = Eigen::Matrix<float_t, dimx, 1>;
    using osv = Eigen::Matrix<float_t, dimy, 1>;
    using cvsv = Eigen::Matrix<float_t, dimcov, 1>;
    using Mat = Eigen::Matrix<float_t, Eigen::Dynamic, Eigen::Dynamic>;
    using arrayStates = std::array<ssv, nparts>;
    using arrayFloat = std::array<float_t, nparts>;
    using filt_func = std::function<const Mat(const ssv&, const cvsv&)>;

public:
    explicit BSFilterWC(const unsigned int& rs = 1) : m_now(0), m_logLastCondLike(0.0), m_resampSched(rs)
    {
        std::fill(m_logUnNormWeights.begin(), m_logUnNormWeights.end(), 0.0);
    }
    virtual ~BSFilterWC() = default;
    float_t getLogCondLike() const override { return m_logLastCondLike; }
    std::vector<Mat> getExpectations() const override { return m_expectations; }

    virtual float_t logQ1Ev(const ssv& x1, const osv& y1, const cvsv& z1) = 0;
    virtual float_t logMuEv(const ssv& x1, const cvsv& z1) = 0;
    virtual float_t logGEv(const osv& yt, const ssv& xt, const cvsv& zt) = 0;
    virtual ssv fSamp(const ssv& xtm1, const cvsv& zt) = 0;
    virtual ssv q1Samp(const osv& y1, const cvsv& z1) = 0;

    void filter(const osv& data, const cvsv& cov, const std::vector<filt_func>& fs = std::vector<filt_func>()) override
    {
        float_t maxW;
        if (m_now > 0) {
            arrayFloat oldLogUnNormWts = m_logUnNormWeights;
            float_t maxOld(-std::numeric_limits<float_t>::infinity());
            for (size_t ii = 0; ii < nparts; ++ii) {
                if (m_logUnNormWeights[ii] > maxOld) maxOld = m_logUnNormWeights[ii];
                ssv newSamp = fSamp(m_particles[ii], cov);
                m_logUnNormWeights[ii] += logGEv(data, newSamp, cov);
                m_particles[ii] = newSamp;
            }
            maxW = *std::max_element(m_logUnNormWeights.begin(), m_logUnNormWeights.end());
            float_t sumExp1(0.0), sumExp2(0.0);
            for (size_t i = 0; i < nparts; ++i) {
                sumExp1 += std::exp(m_logUnNormWeights[i] - maxW);
                sumExp2 += std::exp(oldLogUnNormWts[i] - maxOld);
            }
            m_logLastCondLike = maxW + std::log(sumExp1) - maxOld - std::log(sumExp2);
        } else {
            for (size_t ii = 0; ii < nparts; ++ii) {
                m_particles[ii] = q1Samp(data, cov);
                m_logUnNormWeights[ii] = logMuEv(m_particles[ii], cov);
                m_logUnNormWeights[ii] += logGEv(data, m_particles[ii], cov);
                m_logUnNormWeights[ii] -= logQ1Ev(m_particles[ii], data, cov);
            }
            maxW = *std::max_element(m_logUnNormWeights.begin(), m_logUnNormWeights.end());
            float_t sumExp(0.0);
            for (size_t i = 0; i < nparts; ++i) sumExp += std::exp(m_logUnNormWeights[i] - maxW);
            m_logLastCondLike = -std::log(nparts) + maxW + std::log(sumExp);
            m_expectations.resize(fs.size());
        }
        unsigned int fId(0);
        for (auto& h : fs) {
            Mat first = h(m_particles[0], cov);
            Mat numer = Mat::Zero(first.rows(), first.cols());
            float_t denom(0.0);
            for (size_t p = 0; p < nparts; ++p) {
                numer += h(m_particles[p], cov) * std::exp(m_logUnNormWeights[p] - maxW);
                denom += std::exp(m_logUnNormWeights[p] - maxW);
            }
            m_expectations[fId++] = numer / denom;
        }
        if ((m_now + 1) % m_resampSched == 0) m_resampler.resampLogWts(m_particles, m_logUnNormWeights);
        m_now += 1;
    }

protected:
    arrayStates m_particles;
    arrayFloat m_logUnNormWeights;
    unsigned int m_now;
    float_t m_logLastCondLike;
    resamp_t m_resampler;
    std::vector<Mat> m_expectations;
    unsigned int m_resampSched;
};

}  // namespace filters
}  // namespace pf
#endif
