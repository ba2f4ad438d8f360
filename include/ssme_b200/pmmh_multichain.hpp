// include/ssme_b200/pmmh_multichain.hpp -- many adaptive PMMH chains advanced in lock step, one GPU
// launch (or one launch per rank) per iteration.
//
// Each chain is the reference's sampler (include/ssme/ada_pmmh_mvn.h:325-372: iteration 0 evaluates
// without proposing, MVN random walk on the transformed scale :260-269, Haario adaptation inside
// (t0, t1) :212-250, accept-rate recursion :351,356, NaN = reject :349).  What the reference does for ONE
// chain with `m_pool.work(theta)` (:344) is done here for C chains at once: the C proposals x R replicate
// filters form one batch (BASELINE.json config 3: 64 chains), evaluated through an `evaluator` callback:
//   * single GPU: ssme_b200_loglike_batch (make_gpu_evaluator)
//   * N GPUs:     ssme_b200_loglike_batch_sharded -- every rank runs a contiguous range of the C*R filters
//                 and the per-filter log-likelihoods are all-gathered over NCCL, so that EVERY rank holds
//                 the same numbers, takes the same accept/adapt decisions with the same proposal RNG, and
//                 the chain states stay replicated without any further exchange.
// The log-mean-exp over the R replicates is thread_pool's (thread_pool.h:263-268), evaluated on the host
// in index order.
#ifndef SSME_B200_PMMH_MULTICHAIN_HPP
#define SSME_B200_PMMH_MULTICHAIN_HPP

#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <random>
#include <string>
#include <vector>

#include "../ssme_b200.h"
#include "fixed.hpp"
#include "gpu_pool.hpp"
#include "parameters.hpp"

namespace ssme_b200 {

// out = m + log(sum exp(v_i - m)) - log(n)   (thread_pool.h:263-268)
inline double log_mean_exp_host(const double* v, size_t n)
{
    double m = -std::numeric_limits<double>::infinity();
    for (size_t i = 0; i < n; ++i)
        if (v[i] > m) m = v[i];
    double sum_exp = 0;
    for (size_t i = 0; i < n; ++i) sum_exp += std::exp(v[i] - m);
    return m + std::log(sum_exp) - std::log((double)n);
}

template <size_t numparams, typename float_t = double>
class pmmh_multichain {
public:
    using pack = param::pack<float_t, numparams>;
    using psv = vec<float_t, numparams>;
    using psm = mat<float_t, numparams>;
    // fills per_filter[C*R] (filter (c, r) at c*R + r) with log-likelihood estimates at the C untransformed thetas;
    // must return the same bits on every rank
    using evaluator_t = std::function<void(const double* theta, size_t C, unsigned R, std::uint64_t stream_base, double* per_filter)>;
    using prior_t = std::function<float_t(const pack&)>;

    struct chain_state {
        pack current_theta;
        psm sigma_hat = psm::Zero();
        psv mean_trans_theta = psv::Zero();
        psm Ct;
        float_t ma_accept_rate = 0;
        float_t old_log_like = 0, new_log_like = 0, old_log_prior = 0, new_log_prior = 0;
        float_t log_accept_prob = -std::numeric_limits<float_t>::infinity();
        bool accepted = false;
        std::mt19937 gen;
    };

    pmmh_multichain(const std::vector<psv>& start_trans_thetas, std::vector<std::string> tts, unsigned num_pfilters, unsigned t0,
                    unsigned t1, const psm& C0, prior_t prior, evaluator_t evaluator, unsigned long proposal_seed)
        : m_tts(std::move(tts)), m_R(num_pfilters), m_t0(t0), m_t1(t1), m_iter(0), m_sd(2.4 * 2.4 / numparams), m_eps(.01),
          m_prior(std::move(prior)), m_eval(std::move(evaluator))
    {
        if (start_trans_thetas.empty()) throw std::invalid_argument("need at least one chain");
        if (num_pfilters == 0) throw std::invalid_argument("num_pfilters must be positive");
        m_chains.resize(start_trans_thetas.size());
        for (size_t c = 0; c < m_chains.size(); ++c) {
            m_chains[c].current_theta = pack(start_trans_thetas[c], m_tts);
            m_chains[c].Ct = C0;
            m_chains[c].gen.seed(static_cast<std::uint32_t>(proposal_seed + 7919u * (unsigned long)c));
        }
    }

    size_t num_chains() const { return m_chains.size(); }
    unsigned iterations_done() const { return m_iter; }
    const chain_state& chain(size_t c) const { return m_chains[c]; }
    std::uint64_t proposals_decided() const { return m_iter == 0 ? 0 : (std::uint64_t)(m_iter - 1) * m_chains.size(); }

    // one iteration of every chain (the body of the while loop at ada_pmmh_mvn.h:332-370)
    void step()
    {
        const size_t C = m_chains.size();
        std::vector<pack> proposed(C);
        std::vector<double> theta(C * numparams), per_filter(C * (size_t)m_R);
        for (size_t c = 0; c < C; ++c) {
            chain_state& s = m_chains[c];
            if (m_iter > 0) {
                update_moments_and_Ct(s);
                proposed[c] = pack(q_samp(s), m_tts);
                s.new_log_prior = m_prior(proposed[c]) + proposed[c].get_log_jacobian();
            } else {
                proposed[c] = s.current_theta;
            }
            const psv th = proposed[c].get_untrans_params();
            for (size_t k = 0; k < numparams; ++k) theta[c * numparams + k] = (double)th(k);
        }
        // the hot call: C x R filters in one batch (reference: one m_pool.work per chain)
        m_eval(theta.data(), C, m_R, (std::uint64_t)m_iter * C * m_R, per_filter.data());
        for (size_t c = 0; c < C; ++c) {
            chain_state& s = m_chains[c];
            const float_t ll = (float_t)log_mean_exp_host(per_filter.data() + c * m_R, m_R);
            if (m_iter > 0) {
                s.new_log_like = ll;
                s.log_accept_prob = s.new_log_prior + s.new_log_like - s.old_log_prior - s.old_log_like;
                std::uniform_real_distribution<float_t> runif(0.0, 1.0);
                const float_t log_uniform_draw = std::log(runif(s.gen));
                s.accepted = log_uniform_draw < s.log_accept_prob;
                if (s.accepted) {
                    s.ma_accept_rate = 1.0 / (m_iter + 1.0) + m_iter * s.ma_accept_rate / (m_iter + 1.0);
                    s.current_theta = proposed[c];
                    s.old_log_prior = s.new_log_prior;
                    s.old_log_like = s.new_log_like;
                } else {
                    s.ma_accept_rate = 0.0 / (m_iter + 1.0) + m_iter * s.ma_accept_rate / (m_iter + 1.0);
                }
            } else {
                s.old_log_like = ll;
                s.old_log_prior = m_prior(s.current_theta) + s.current_theta.get_log_jacobian();
            }
        }
        m_iter++;
    }

    void run(unsigned iters)
    {
        for (unsigned i = 0; i < iters; ++i) step();
    }

private:
    void update_moments_and_Ct(chain_state& s)
    {
        const psv x = s.current_theta.get_trans_params();
        const float_t n = (float_t)m_iter;
        if (m_iter == 1) {
            s.mean_trans_theta += x;
        } else if (m_iter == 2) {
            s.sigma_hat = outer(s.mean_trans_theta, s.mean_trans_theta) + outer(x, x) - outer(s.mean_trans_theta, x) - outer(x, s.mean_trans_theta);
            s.sigma_hat *= (float_t).5;
            s.mean_trans_theta = (float_t).5 * s.mean_trans_theta + (float_t).5 * x;
        } else {
            const psv d = x - s.mean_trans_theta;
            s.sigma_hat = s.sigma_hat * ((n - (float_t)2.0) / (n - (float_t)1.0)) + outer(d, d) / n;
            s.mean_trans_theta = ((n - (float_t)1.0) * s.mean_trans_theta + x) / n;
        }
        if ((m_t1 > m_iter) && (m_iter > m_t0)) s.Ct = m_sd * (s.sigma_hat + m_eps * psm::Identity());
    }

    psv q_samp(chain_state& s)
    {
        const psm A = cholesky(s.Ct);
        std::normal_distribution<float_t> rnorm(0.0, 1.0);
        psv z;
        for (size_t i = 0; i < numparams; ++i) z(i) = rnorm(s.gen);
        return s.current_theta.get_trans_params() + A * z;
    }

    std::vector<std::string> m_tts;
    unsigned m_R, m_t0, m_t1, m_iter;
    float_t m_sd, m_eps;
    prior_t m_prior;
    evaluator_t m_eval;
    std::vector<chain_state> m_chains;
};

// evaluator bound to a backend handle; works for one GPU and, after ssme_b200_comm_init, for N ranks
inline std::function<void(const double*, size_t, unsigned, std::uint64_t, double*)> make_gpu_evaluator(ssme_b200_handle h)
{
    return [h](const double* theta, size_t C, unsigned R, std::uint64_t stream_base, double* per_filter) {
        throw_on_error(ssme_b200_loglike_batch_sharded(h, theta, C, R, stream_base, nullptr, per_filter));
    };
}

}  // namespace ssme_b200
#endif
