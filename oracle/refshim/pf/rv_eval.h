// oracle/refshim/pf/rv_eval.h -- TEST INFRASTRUCTURE.  Stand-in for the four pf::rveval functions the reference calls
// (univ_svol_bootstrap_filter.h:85,92,102; estimate_univ_svol.h:95,98,101,153; test/test_liu_west.cpp:88,131,138).
// pf (tbrown122387/pf, unpinned, absent) is restated from its published formulas:
//   evalUnivNorm(x, mu, sigma, log)      -log sigma - 1/2 log 2pi - 1/2 (x-mu)^2 / sigma^2, sigma > 0 else -inf
//   evalUniform(x, a, b, log)            -log(b - a) inside [a, b], else -inf
//   evalUnivInvGamma(x, alpha, beta, log) alpha log beta - lgamma alpha - (alpha+1) log x - beta / x, x > 0
//   twiceFisher(phi)                      log(1 + phi) - log(1 - phi)   (= parameters.h:356)
#ifndef SSME_REFSHIM_PF_RV_EVAL_H
#define SSME_REFSHIM_PF_RV_EVAL_H
#include <cmath>
#include <limits>
#include <stdexcept>

namespace pf {
namespace rveval {

template <typename float_t>
constexpr float_t inv_sqrt_2pi = float_t(0.3989422804014327);
template <typename float_t>
constexpr float_t log_two_pi = float_t(1.8378770664093453);  // log(2 pi)

template <typename float_t>
float_t evalUnivNorm(const float_t& x, const float_t& mu, const float_t& sigma, bool log = false)
{
    float_t exponent = -.5 * (x - mu) * (x - mu) / (sigma * sigma);
    if (sigma > 0.0) {
        if (log) return -std::log(sigma) - .5 * log_two_pi<float_t> + exponent;
        return inv_sqrt_2pi<float_t> * std::exp(exponent) / sigma;
    }
    return log ? -std::numeric_limits<float_t>::infinity() : float_t(0.0);
}

template <typename float_t>
float_t evalUniform(const float_t& x, const float_t& lower, const float_t& upper, bool log = false)
{
    if (x > lower && x <= upper) {
        float_t width = upper - lower;
        return log ? -std::log(width) : float_t(1.0) / width;
    }
    return log ? -std::numeric_limits<float_t>::infinity() : float_t(0.0);
}

template <typename float_t>
float_t evalUnivInvGamma(const float_t& x, const float_t& alpha, const float_t& beta, bool log = false)
{
    if (x > 0.0 && alpha > 0.0 && beta > 0.0) {
        float_t lv = alpha * std::log(beta) - std::lgamma(alpha) - (alpha + 1.0) * std::log(x) - beta / x;
        return log ? lv : std::exp(lv);
    }
    return log ? -std::numeric_limits<float_t>::infinity() : float_t(0.0);
}

template <typename float_t>
float_t twiceFisher(const float_t& phi)
{
    if (phi <= -1.0 || phi >= 1.0) throw std::invalid_argument("error: phi was not between -1 and 1");
    return std::log(1.0 + phi) - std::log(1.0 - phi);
}

template <typename float_t>
float_t invTwiceFisher(const float_t& psi)
{
    float_t ans = (1.0 - std::exp(psi)) / (-1.0 - std::exp(psi));
    return ans;
}

}  // namespace rveval
}  // namespace pf
#endif
