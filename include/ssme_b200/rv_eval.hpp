// include/ssme_b200/rv_eval.hpp -- the handful of pf::rveval functions SSME's call sites use
// (reference call sites: example/estimate_univ_svol.h:95,98,101,153; univ_svol_bootstrap_filter.h:85,92,102).
// The pf library is an external, unpinned dependency of the reference (CMakeLists.txt:12); these are
// restatements of its published formulas, host-side only (priors, start values).
#ifndef SSME_B200_RV_EVAL_HPP
#define SSME_B200_RV_EVAL_HPP

#include <cmath>
#include <limits>

namespace ssme_b200 {
namespace rveval {

template <typename float_t>
float_t evalUnivNorm(const float_t& x, const float_t& mu, const float_t& sigma, bool log)
{
    const float_t exponent = -.5 * (x - mu) * (x - mu) / (sigma * sigma);
    if (sigma > 0.0) {
        const float_t lg = -std::log(sigma) - .5 * std::log(2.0 * M_PI) + exponent;
        return log ? lg : std::exp(lg);
    }
    return log ? -std::numeric_limits<float_t>::infinity() : 0.0;
}

template <typename float_t>
float_t evalUniform(const float_t& x, const float_t& lower, const float_t& upper, bool log)
{
    if ((x > lower) && (x <= upper)) return log ? -std::log(upper - lower) : 1.0 / (upper - lower);
    return log ? -std::numeric_limits<float_t>::infinity() : 0.0;
}

template <typename float_t>
float_t evalUnivInvGamma(const float_t& x, const float_t& alpha, const float_t& beta, bool log)
{
    if ((x > 0.0) && (alpha > 0.0) && (beta > 0.0)) {
        const float_t lg = alpha * std::log(beta) - std::lgamma(alpha) - (alpha + 1.0) * std::log(x) - beta / x;
        return log ? lg : std::exp(lg);
    }
    return log ? -std::numeric_limits<float_t>::infinity() : 0.0;
}

// log(1+phi) - log(1-phi): the same map as param::twice_fisher_trans::trans (parameters.h:356)
template <typename float_t>
float_t twiceFisher(const float_t& phi)
{
    return std::log(1.0 + phi) - std::log(1.0 - phi);
}

}  // namespace rveval
}  // namespace ssme_b200
#endif
