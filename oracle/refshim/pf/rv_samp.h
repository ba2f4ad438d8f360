// oracle/refshim/pf/rv_samp.h -- TEST INFRASTRUCTURE.  Stand-in for the pf::rvsamp classes the reference uses:
//   UnivNormSampler  (example/univ_svol_bootstrap_filter.h:30,68,77; test/test_liu_west.cpp:38)
//   UniformSampler   (ada_pmmh_mvn.h:330,348; test/test_liu_west.cpp:41-44)
//   MVNSampler       (ada_pmmh_mvn.h:112,266-268; liu_west_filter.h:325,426-427,582)
//   k_gen            (liu_west_filter.h:328,404)
// pf (tbrown122387/pf, unpinned, absent from the image) is restated from its published behaviour:
// std::mt19937 per object + std::normal_distribution / std::uniform_real_distribution / std::discrete_distribution.
// Differences, declared: (1) seeding and playback hooks (shim_streams.h); (2) MVNSampler factorises the covariance by
// Cholesky (pf: symmetric eigendecomposition) -- same law, different sample for the same normals; a non-positive
// pivot gives a zero column so that a zero covariance (delta = 1) draws the mean exactly.
#ifndef SSME_REFSHIM_PF_RV_SAMP_H
#define SSME_REFSHIM_PF_RV_SAMP_H
#include <Eigen/Dense>
#include <algorithm>
#include <array>
#include <cmath>
#include <random>

#include "shim_streams.h"

namespace pf {
namespace rvsamp {

class rvsamp_base {
public:
    rvsamp_base() : m_rng{shim::next_seed()} {}

protected:
    std::mt19937 m_rng;
};

template <typename float_t>
class UnivNormSampler : public rvsamp_base {
public:
    UnivNormSampler() : m_z_gen(0.0, 1.0), m_mu(0.0), m_sigma(1.0) {}
    UnivNormSampler(float_t mu, float_t sigma) : m_z_gen(0.0, 1.0), m_mu(mu), m_sigma(sigma) {}
    void setStdDev(float_t sigma) { m_sigma = sigma; }
    void setMean(float_t mu) { m_mu = mu; }
    float_t sample()
    {
        const float_t z = shim::normal_stream().active() ? static_cast<float_t>(shim::normal_stream().next()) : m_z_gen(m_rng);
        return m_mu + m_sigma * z;
    }

private:
    std::normal_distribution<float_t> m_z_gen;
    float_t m_mu, m_sigma;
};

template <typename float_t>
class UniformSampler : public rvsamp_base {
public:
    UniformSampler() : m_unif_gen(0.0, 1.0), m_lo(0.0), m_hi(1.0) {}
    UniformSampler(float_t lower, float_t upper) : m_unif_gen(lower, upper), m_lo(lower), m_hi(upper) {}
    float_t sample()
    {
        if (shim::uniform_stream().active())  // what uniform_real_distribution computes: u * (b - a) + a
            return static_cast<float_t>(shim::uniform_stream().next()) * (m_hi - m_lo) + m_lo;
        return m_unif_gen(m_rng);
    }

private:
    std::uniform_real_distribution<float_t> m_unif_gen;
    float_t m_lo, m_hi;
};

template <size_t dim, typename float_t>
class MVNSampler : public rvsamp_base {
public:
    using Vec = Eigen::Matrix<float_t, dim, 1>;
    using Mat = Eigen::Matrix<float_t, dim, dim>;
    MVNSampler() : m_z_gen(0.0, 1.0)
    {
        setMean(Vec::Zero());
        setCovar(Mat::Identity());
    }
    MVNSampler(const Vec& meanVec, const Mat& covMat) : m_z_gen(0.0, 1.0)
    {
        setMean(meanVec);
        setCovar(covMat);
    }
    void setMean(const Vec& meanVec) { m_mean = meanVec; }
    void setCovar(const Mat& covMat)
    {
        // lower Cholesky factor, row by row; a non-positive pivot zeroes its column
        m_scale = Mat::Zero();
        for (size_t i = 0; i < dim; ++i)
            for (size_t j = 0; j <= i; ++j) {
                float_t acc = covMat(i, j);
                for (size_t k = 0; k < j; ++k) acc = acc - m_scale(i, k) * m_scale(j, k);
                if (i == j) m_scale(i, j) = (acc > 0.0) ? std::sqrt(acc) : float_t(0.0);
                else m_scale(i, j) = (m_scale(j, j) > 0.0) ? acc / m_scale(j, j) : float_t(0.0);
            }
    }
    Vec sample()
    {
        Vec Z;
        for (size_t i = 0; i < dim; ++i)
            Z(i) = shim::mvn_stream().active() ? static_cast<float_t>(shim::mvn_stream().next()) : m_z_gen(m_rng);
        return m_mean + m_scale * Z;  // pf: mean + scale_mat * Z (matrix-vector product first, then the mean)
    }

private:
    std::normal_distribution<float_t> m_z_gen;
    Mat m_scale;
    Vec m_mean;
};

// draws N indices i.i.d. from the distribution proportional to exp(logWts) (liu_west_filter.h:404)
template <size_t N, typename float_t>
class k_gen : public rvsamp_base {
public:
    std::array<unsigned int, N> sample(const std::array<float_t, N>& logWts)
    {
        std::array<float_t, N> w;
        const float_t m = *std::max_element(logWts.begin(), logWts.end());
        std::transform(logWts.begin(), logWts.end(), w.begin(), [&m](const float_t& d) -> float_t { return std::exp(d - m); });
        std::discrete_distribution<> kGen(w.begin(), w.end());
        std::array<unsigned int, N> ks;
        if (shim::kgen_stream().active()) {
            shim::playback_engine eng(shim::kgen_stream());
            for (size_t i = 0; i < N; ++i) ks[i] = kGen(eng);
        } else {
            for (size_t i = 0; i < N; ++i) ks[i] = kGen(m_rng);
        }
        return ks;
    }
};

}  // namespace rvsamp
}  // namespace pf
#endif
