// include/ssme_b200/ada_pmmh_mvn.hpp -- adaptive particle marginal Metropolis-Hastings with the
// likelihood evaluated on the GPU.
//
// Drop-in for the reference class template include/ssme/ada_pmmh_mvn.h (same name, template
// parameters, constructor argument list, public methods and file formats):
//   ctor                      ada_pmmh_mvn.h:54-67, 160-209   (+ two trailing defaulted arguments)
//   commence_sampling         ada_pmmh_mvn.h:325-372   iteration 0 evaluates without proposing (:362-365)
//   update_moments_and_Ct     ada_pmmh_mvn.h:212-250   Haario-style C_t = (2.4^2/d)(Sigma_hat + 0.01 I), t0 < iter < t1
//   q_samp                    ada_pmmh_mvn.h:260-269   MVN random walk on the transformed scale
//   record_params / messages  ada_pmmh_mvn.h:272-322   same CSV layouts, same time-stamped file names (:375-383)
// What changed: `m_pool.work(theta)` (ada_pmmh_mvn.h:344,363) no longer fans a user-written C++
// filter out over std::threads; it launches num_pfilters bootstrap filters on the B200 through the
// C ABI (gpu_pool.hpp).  The state-space model therefore is a device model id (gpu_options::model)
// instead of the pure virtual log_like_eval; log_prior_eval stays a host-side pure virtual.
// `mc` and `num_threads` are accepted for source compatibility and ignored.
#ifndef SSME_B200_ADA_PMMH_MVN_HPP
#define SSME_B200_ADA_PMMH_MVN_HPP

#include <chrono>
#include <cmath>
#include <ctime>
#include <fstream>
#include <iostream>
#include <limits>
#include <random>
#include <string>
#include <vector>

#include "fixed.hpp"
#include "gpu_pool.hpp"
#include "parameters.hpp"
#include "utils.hpp"

template <size_t numparams, size_t dimobs, size_t numparts, typename float_t, bool debug = false>
class ada_pmmh_mvn {
public:
    using osv = ssme_b200::vec<float_t, dimobs>;
    using psv = ssme_b200::vec<float_t, numparams>;
    using psm = ssme_b200::mat<float_t, numparams>;
    using dyn_data_t = param::pack<float_t, numparams>;
    using static_data_t = std::vector<osv>;

    ada_pmmh_mvn(const psv& start_trans_theta, std::vector<std::string> tts, const unsigned int& num_mcmc_iters,
                 const unsigned int& num_pfilters, const std::string& data_file, const std::string& sample_file_base_name,
                 const std::string& message_file_base_name, const bool& mc, const unsigned int& t0, const unsigned int& t1,
                 const psm& C0, bool print_to_console, unsigned int print_every_k, unsigned int num_threads,
                 const ssme_b200::gpu_options& gpu = ssme_b200::gpu_options(), unsigned long proposal_seed = 0)
        : m_current_theta(start_trans_theta, tts)
        , m_tts(tts)
        , m_sigma_hat(psm::Zero())
        , m_mean_trans_theta(psv::Zero())
        , m_ma_accept_rate(0.0)
        , m_t0(t0)
        , m_t1(t1)
        , m_Ct(C0)
        , m_num_mcmc_iters(num_mcmc_iters)
        , m_iter(0)
        , m_sd(2.4 * 2.4 / numparams)
        , m_eps(.01)
        , m_print_to_console(print_to_console)
        , m_print_every_k(print_every_k)
        , m_pool(num_pfilters, (unsigned)numparts, gpu)
        , m_gen(proposal_seed ? static_cast<std::uint32_t>(proposal_seed)
                              : static_cast<std::uint32_t>(std::chrono::high_resolution_clock::now().time_since_epoch().count()))
        , m_old_log_like(0.0)
        , m_new_log_like(0.0)
        , m_old_log_prior(0.0)
        , m_new_log_prior(0.0)
        , m_log_accept_prob(-std::numeric_limits<float_t>::infinity())
        , m_accepted(false)
    {
        (void)mc;
        (void)num_threads;
        m_data = utils::read_data<dimobs, float_t>(data_file);
        m_pool.add_observed_data(m_data);  // throws std::length_error on an empty series (estimate_univ_svol.h:112-113)
        std::cerr << "first row of data: \n" << m_data[0].transpose() << "\n";
        m_samples_file_stream.open(gen_string_with_time(sample_file_base_name));
        m_message_stream.open(gen_string_with_time(message_file_base_name));
    }
    virtual ~ada_pmmh_mvn() = default;

    psm get_ct() const { return m_Ct; }

    void commence_sampling()
    {
        std::uniform_real_distribution<float_t> runif(0.0, 1.0);
        psv proposed_trans_theta;
        while (m_iter < m_num_mcmc_iters) {
            if (m_iter > 0) {
                update_moments_and_Ct(m_current_theta);
                proposed_trans_theta = q_samp(m_current_theta);
                dyn_data_t proposed_theta(proposed_trans_theta, m_tts);
                m_new_log_prior = log_prior_eval(proposed_theta) + proposed_theta.get_log_jacobian();
                m_new_log_like = m_pool.work(proposed_theta);  // <- the hot call, now one GPU launch

                m_log_accept_prob = m_new_log_prior + m_new_log_like - m_old_log_prior - m_old_log_like;
                float_t log_uniform_draw = std::log(runif(m_gen));
                m_accepted = log_uniform_draw < m_log_accept_prob;  // false when the probability is NaN
                if (m_accepted) {
                    m_ma_accept_rate = 1.0 / (m_iter + 1.0) + m_iter * m_ma_accept_rate / (m_iter + 1.0);
                    m_current_theta = proposed_theta;
                    m_old_log_prior = m_new_log_prior;
                    m_old_log_like = m_new_log_like;
                } else {
                    m_ma_accept_rate = 0.0 / (m_iter + 1.0) + m_iter * m_ma_accept_rate / (m_iter + 1.0);
                    if (std::isnan(m_log_accept_prob)) std::cerr << "accept proability had a nan in it\n";
                }
            } else {
                m_old_log_like = m_pool.work(m_current_theta);
                m_old_log_prior = log_prior_eval(m_current_theta) + m_current_theta.get_log_jacobian();
            }
            record_params();
            record_messages();
            m_iter++;
        }
        m_samples_file_stream.flush();
        m_message_stream.flush();
    }

    // prior density on the untransformed (constrained) scale; the Jacobian is added by the sampler
    virtual float_t log_prior_eval(const param::pack<float_t, numparams>& theta) = 0;

    // One particle-filter estimate of the log-likelihood at theta (reference: the user's pure virtual,
    // ada_pmmh_mvn.h:99-100).  Here it is implemented by the device model; `data` must be the series the
    // object was constructed with (it already lives in HBM).
    virtual float_t log_like_eval(const param::pack<float_t, numparams>& theta, const std::vector<osv>& data)
    {
        if (data.empty()) throw std::length_error("can't read in data\n");
        const auto th = theta.get_untrans_params();
        double in[numparams], out = 0.0, one = 0.0;
        for (size_t k = 0; k < numparams; ++k) in[k] = (double)th(k);
        ssme_b200::throw_on_error(ssme_b200_loglike_batch(m_pool.handle(), in, 1, 1, m_single_stream++, &out, &one));
        return (float_t)one;
    }

    // read-only views of the chain state (tests, drivers)
    unsigned int iterations_done() const { return m_iter; }
    float_t accept_rate() const { return m_ma_accept_rate; }
    const dyn_data_t& current_theta() const { return m_current_theta; }
    float_t current_log_like() const { return m_old_log_like; }
    ssme_b200::gpu_pool<numparams, dimobs, float_t, debug>& pool() { return m_pool; }

private:
    void update_moments_and_Ct(const dyn_data_t& new_theta)
    {
        const psv x = new_theta.get_trans_params();
        const float_t n = (float_t)m_iter;
        if (m_iter == 1) {
            m_mean_trans_theta += x;
        } else if (m_iter == 2) {
            // two-point sample covariance (n-1 denominator) and mean
            m_sigma_hat = ssme_b200::outer(m_mean_trans_theta, m_mean_trans_theta) + ssme_b200::outer(x, x) -
                          ssme_b200::outer(m_mean_trans_theta, x) - ssme_b200::outer(x, m_mean_trans_theta);
            m_sigma_hat *= (float_t).5;
            m_mean_trans_theta = (float_t).5 * m_mean_trans_theta + (float_t).5 * x;
        } else if (m_iter > 2) {
            const psv d = x - m_mean_trans_theta;
            m_sigma_hat = m_sigma_hat * ((n - (float_t)2.0) / (n - (float_t)1.0)) + ssme_b200::outer(d, d) / n;
            m_mean_trans_theta = ((n - (float_t)1.0) * m_mean_trans_theta + x) / n;
        } else {
            std::cerr << "something went wrong\n";
        }
        if ((m_t1 > m_iter) && (m_iter > m_t0)) m_Ct = m_sd * (m_sigma_hat + m_eps * psm::Identity());
    }

    psv q_samp(const dyn_data_t& old_theta)
    {
        // theta' ~ N(theta, C_t) on the transformed scale.  The reference factorises C_t inside
        // pf::rvsamp::MVNSampler (external); any factor A with A A^T = C_t gives the same law.
        const psm A = ssme_b200::cholesky(m_Ct);
        std::normal_distribution<float_t> rnorm(0.0, 1.0);
        psv z;
        for (size_t i = 0; i < numparams; ++i) z(i) = rnorm(m_gen);
        return old_theta.get_trans_params() + A * z;
    }

    void record_params()
    {
        if (m_iter % m_print_every_k == 0) {
            if (m_samples_file_stream.is_open()) {
                const psv p = m_current_theta.get_untrans_params();
                for (size_t i = 0; i < numparams; ++i) {
                    if (i == 0) m_samples_file_stream << p(i);
                    else m_samples_file_stream << "," << p(i);
                }
                m_samples_file_stream << "\n";
            } else {
                std::cerr << "tried to write to a closed ofstream! " << "\n";
                m_message_stream << "tried to write to a closed ofstream! " << "\n";
            }
        }
    }

    void record_messages()
    {
        static const char* header = "iter number, accept rate, old_ll, new_ll, old_lprior, new_lprior, accept prob, outcome\n";
        if (m_iter == 0) {
            m_message_stream << header;
            if (m_print_to_console) std::cout << header;
        }
        auto emit = [&](std::ostream& os) {
            os << m_iter + 1 << ", " << m_ma_accept_rate << ", " << m_old_log_like << ", " << m_new_log_like << ", " << m_old_log_prior
               << ", " << m_new_log_prior << ", " << m_log_accept_prob << ", " << m_accepted << "\n";
        };
        emit(m_message_stream);
        if (m_print_to_console) emit(std::cout);
    }

    static std::string gen_string_with_time(const std::string& str)
    {
        time_t now = time(0);
        struct tm tstruct = *localtime(&now);
        char buf[80];
        strftime(buf, sizeof(buf), "%Y-%m-%d.%H-%M-%S", &tstruct);
        return str + "_" + buf;
    }

    dyn_data_t m_current_theta;
    std::vector<std::string> m_tts;
    psm m_sigma_hat;
    psv m_mean_trans_theta;
    float_t m_ma_accept_rate;
    unsigned int m_t0, m_t1;
    psm m_Ct;
    std::ofstream m_samples_file_stream, m_message_stream;
    unsigned int m_num_mcmc_iters, m_iter;
    float_t m_sd, m_eps;
    bool m_print_to_console;
    unsigned int m_print_every_k;
    ssme_b200::gpu_pool<numparams, dimobs, float_t, debug> m_pool;
    static_data_t m_data;
    std::mt19937 m_gen;
    std::uint64_t m_single_stream = (std::uint64_t)1 << 59;  // stream ids of direct log_like_eval calls: the upper half of the 60-bit id space
    float_t m_old_log_like, m_new_log_like, m_old_log_prior, m_new_log_prior, m_log_accept_prob;
    bool m_accepted;
};

#endif
