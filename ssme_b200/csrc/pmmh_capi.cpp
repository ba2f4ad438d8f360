// ssme_b200/csrc/pmmh_capi.cpp -- ssme_b200_pmmh_run: the C++ multi-chain PMMH host loop behind the C ABI.
// Host-only translation unit (no device code); the likelihoods come from ssme_b200_loglike_batch_sharded.
#include <chrono>
#include <string>
#include <vector>

#include "../../include/ssme_b200.h"
#include "../../include/ssme_b200/pmmh_multichain.hpp"
#include "../../include/ssme_b200/rv_eval.hpp"

namespace {

using evaluator_t = std::function<void(const double*, size_t, unsigned, std::uint64_t, double*)>;

template <size_t NP>
int run_chains(const evaluator_t& evaluator, const ssme_b200_pmmh_config* cfg, const double* start_theta, const std::vector<std::string>& tts,
               typename ssme_b200::pmmh_multichain<NP, double>::prior_t prior, double* final_theta, double* mean_theta,
               double* accept_rate, double* last_loglik, double* seconds)
{
    using driver = ssme_b200::pmmh_multichain<NP, double>;
    const size_t C = (size_t)cfg->num_chains;
    std::vector<typename driver::psv> start(C);
    for (size_t c = 0; c < C; ++c) {
        typename driver::psv untrans;
        for (size_t k = 0; k < NP; ++k) untrans(k) = start_theta[c * NP + k];
        start[c] = typename driver::pack(untrans, tts, false).get_trans_params();
    }
    driver d(start, tts, (unsigned)cfg->num_pfilters, (unsigned)cfg->t0, (unsigned)cfg->t1, driver::psm::Identity() * cfg->c0_diag, prior,
             evaluator, (unsigned long)cfg->proposal_seed);
    std::vector<double> sum(C * NP, 0.0);
    const auto t_begin = std::chrono::steady_clock::now();
    for (int it = 0; it < cfg->iterations; ++it) {
        d.step();
        if (mean_theta)
            for (size_t c = 0; c < C; ++c) {
                const auto th = d.chain(c).current_theta.get_untrans_params();
                for (size_t k = 0; k < NP; ++k) sum[c * NP + k] += th(k);
            }
    }
    const auto t_end = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t_end - t_begin).count();
    for (size_t c = 0; c < C; ++c) {
        const auto th = d.chain(c).current_theta.get_untrans_params();
        for (size_t k = 0; k < NP; ++k) {
            if (final_theta) final_theta[c * NP + k] = th(k);
            if (mean_theta) mean_theta[c * NP + k] = sum[c * NP + k] / cfg->iterations;
        }
        if (accept_rate) accept_rate[c] = d.chain(c).ma_accept_rate;
        if (last_loglik) last_loglik[c] = d.chain(c).old_log_like;
    }
    return SSME_B200_OK;
}

}  // namespace

// defined in capi.cu
extern "C" const char* ssme_b200_last_error(void);
namespace ssme { int set_last_error(int code, const char* msg); }

static int run_model(int32_t model, const evaluator_t& evaluator, const ssme_b200_pmmh_config* cfg, const double* start_theta,
                     double* final_theta, double* mean_theta, double* accept_rate, double* last_loglik, double* seconds)
{
    if (!cfg || !start_theta) return ssme::set_last_error(SSME_B200_EINVAL, "null argument");
    if (cfg->struct_size != (int32_t)sizeof(ssme_b200_pmmh_config)) return ssme::set_last_error(SSME_B200_EINVAL, "ssme_b200_pmmh_config size mismatch");
    if (cfg->num_chains < 1 || cfg->num_pfilters < 1 || cfg->iterations < 1)
        return ssme::set_last_error(SSME_B200_EINVAL, "num_chains, num_pfilters and iterations must be positive");
    namespace rv = ssme_b200::rveval;
    try {
        if (model == SSME_B200_MODEL_SV) {
            auto prior = [](const param::pack<double, 3>& theta) {
                const auto p = theta.get_untrans_params();
                return rv::evalUnivNorm<double>(p(0), 1.0, 1.0, true) + rv::evalUniform<double>(p(1), 0.0, 1.0, true) +
                       rv::evalUnivInvGamma<double>(p(2), .001, .001, true);
            };
            return run_chains<3>(evaluator, cfg, start_theta, {"null", "twice_fisher", "log"}, prior, final_theta, mean_theta, accept_rate,
                                 last_loglik, seconds);
        }
        if (model != SSME_B200_MODEL_SV_LEVERAGE) return ssme::set_last_error(SSME_B200_EINVAL, "unknown model id");
        auto prior = [](const param::pack<double, 4>& theta) {
            const auto p = theta.get_untrans_params();
            return rv::evalUniform<double>(p(0), 0.0, 1.0, true) + rv::evalUnivNorm<double>(p(1), 0.0, 1.0, true) +
                   rv::evalUniform<double>(p(2), 0.0, 5.0, true) + rv::evalUniform<double>(p(3), -1.0, 1.0, true);
        };
        return run_chains<4>(evaluator, cfg, start_theta, {"logit", "null", "log", "twice_fisher"}, prior, final_theta, mean_theta,
                             accept_rate, last_loglik, seconds);
    } catch (const std::invalid_argument& e) {
        return ssme::set_last_error(SSME_B200_EINVAL, e.what());
    } catch (const std::length_error& e) {
        return ssme::set_last_error(SSME_B200_ELENGTH, e.what());
    } catch (const std::exception& e) {
        return ssme::set_last_error(SSME_B200_ERUNTIME, e.what());
    }
}

extern "C" int ssme_b200_pmmh_run(ssme_b200_handle h, const ssme_b200_pmmh_config* cfg, const double* start_theta, double* final_theta,
                                  double* mean_theta, double* accept_rate, double* last_loglik, double* seconds)
{
    if (!h) return ssme::set_last_error(SSME_B200_EINVAL, "null handle");
    return run_model(ssme_b200_model(h), ssme_b200::make_gpu_evaluator(h), cfg, start_theta, final_theta, mean_theta, accept_rate,
                     last_loglik, seconds);
}

extern "C" int ssme_b200_pmmh_run_custom(int32_t model, const ssme_b200_pmmh_config* cfg, ssme_b200_evaluator_fn evaluator, void* user,
                                         const double* start_theta, double* final_theta, double* mean_theta, double* accept_rate,
                                         double* last_loglik, double* seconds)
{
    if (!evaluator) return ssme::set_last_error(SSME_B200_EINVAL, "null evaluator");
    evaluator_t ev = [evaluator, user](const double* theta, size_t C, unsigned R, std::uint64_t base, double* out) {
        if (evaluator(user, theta, C, R, base, out) != 0) throw std::runtime_error("the likelihood evaluator reported a failure");
    };
    return run_model(model, ev, cfg, start_theta, final_theta, mean_theta, accept_rate, last_loglik, seconds);
}
