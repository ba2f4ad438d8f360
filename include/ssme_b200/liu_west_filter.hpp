// include/ssme_b200/liu_west_filter.hpp -- Liu-West joint state / parameter filter on the GPU backend.
//
// Mirrors the user-facing part of LWFilter2WithCovs (include/ssme/liu_west_filter.h:1949-2188) and of LWFilterWithCovs
// (:741-968, the auxiliary-particle form; LWFilterWithCovs_svol below): constructed from the
// transform names, delta and the resampling schedule (:2100-2113), `getLogCondLike()` (:2180-2184), and the prior the
// user's subclass supplies through `paramPriorSamp` (test/test_liu_west.cpp:339-349: independent uniforms).  The
// reference is streaming -- `filter(y_t, z_t)` once per observation, with the state-space model given as seven pure
// virtuals on Eigen vectors; here the model is the device functor of svol_lw_2_par (test/test_liu_west.cpp:213-358)
// and the whole series is filtered by one call, after which the per-step results are read back.
#ifndef SSME_B200_LIU_WEST_FILTER_HPP
#define SSME_B200_LIU_WEST_FILTER_HPP

#include <array>
#include <stdexcept>
#include <string>
#include <vector>

#include "../ssme_b200.h"
#include "gpu_pool.hpp"

namespace ssme_b200 {

namespace detail {
template <size_t nparts, typename float_t, int form>
class lw_svol_base;
}

// LWFilter2WithCovs on svol_lw_2_par (SISR form, bootstrap proposal)
template <size_t nparts, typename float_t = double>
using LWFilter2WithCovs_svol = detail::lw_svol_base<nparts, float_t, SSME_B200_LW_SISR>;
// LWFilterWithCovs on svol_lw_1_par (auxiliary particle filter: first-stage weights through propMu; test_liu_west.cpp:26-157)
template <size_t nparts, typename float_t = double>
using LWFilterWithCovs_svol = detail::lw_svol_base<nparts, float_t, SSME_B200_LW_APF>;

// The covariate-free twins LWFilter2 (liu_west_filter.h:1470-1761) and LWFilter (:236-552): the same filters on the same
// model with the covariate held at zero (no leverage term in the transition) -- filter(y_t) instead of filter(y_t, z_t).
template <size_t nparts, typename float_t = double>
class LWFilter2_svol;
template <size_t nparts, typename float_t = double>
class LWFilter_svol;

namespace detail {
template <size_t nparts, typename float_t, int form>
class lw_svol_base {
public:
    using psv = vec<float_t, 4>;  // phi, mu, sigma, rho

    // transforms must be {"logit", "null", "log", "twice_fisher"} -- the ones svol_lw_2_par passes (test_liu_west.cpp:263)
    lw_svol_base(std::vector<std::string> transforms, float_t delta, const psv& prior_lower, const psv& prior_upper,
                 const unsigned int& rs = 1, const gpu_options& opt = gpu_options())
        : m_delta(delta), m_lo(prior_lower), m_hi(prior_upper)
    {
        const std::vector<std::string> want{"logit", "null", "log", "twice_fisher"};
        if (transforms != want) throw std::invalid_argument("the device model uses the transforms logit, null, log, twice_fisher");
        // the resampling schedule of the reference's constructors (liu_west_filter.h:1686, 1754): the SISR form takes it, the
        // auxiliary-particle form resamples at every step
        if (rs < 1 || (rs != 1 && form != SSME_B200_LW_SISR))
            throw std::invalid_argument("the auxiliary-particle GPU Liu-West filter resamples at every step (rs = 1)");
        ssme_b200_config c{};
        c.struct_size = (int32_t)sizeof(c);
        c.device = opt.device;
        c.model = SSME_B200_MODEL_SV_LEVERAGE;
        c.num_particles = (int32_t)nparts;
        c.resampler = opt.resampler;
        c.resample_every = (int32_t)rs;
        c.dtype = SSME_B200_DTYPE_F64;
        c.rng_mode = SSME_B200_RNG_PHILOX;
        c.seed = opt.seed;
        c.force_global_memory = 1;
        throw_on_error(ssme_b200_create(&c, &m_h));
    }
    ~lw_svol_base() { ssme_b200_destroy(m_h); }
    lw_svol_base(const lw_svol_base&) = delete;
    lw_svol_base& operator=(const lw_svol_base&) = delete;

    // filter(obs_data, cov_data) for t = 0 .. T-1 in one call; cov[t] is the covariate z_t (the lagged observation)
    void filter_series(const std::vector<float_t>& obs, const std::vector<float_t>& cov, std::uint64_t stream_id = 0)
    {
        if (obs.empty() || obs.size() != cov.size()) throw std::length_error("observations and covariates must be non-empty and of equal length");
        std::vector<double> rows(2 * obs.size());
        for (size_t t = 0; t < obs.size(); ++t) { rows[2 * t] = (double)obs[t]; rows[2 * t + 1] = (double)cov[t]; }
        throw_on_error(ssme_b200_replace_observations(m_h, rows.data(), obs.size(), 2));
        m_streaming = false;
        m_cond_like.assign(obs.size(), 0.0);
        m_theta_bar.assign(obs.size() * 4, 0.0);
        double lo[4], hi[4];
        for (int k = 0; k < 4; ++k) { lo[k] = (double)m_lo(k); hi[k] = (double)m_hi(k); }
        throw_on_error(ssme_b200_lw_filter_form(m_h, form, lo, hi, (double)m_delta, stream_id, &m_loglik, m_cond_like.data(),
                                                m_theta_bar.data(), m_final_mean.data(), nullptr, nullptr));
    }

    // filter(obs, cov, fs) over the series: also E[h | y_{1:t}] before resampling for h = x_t, phi, mu, sigma, rho
    // (getExpectations, liu_west_filter.h:1087-1101 / :2263-2276); read back with getExpectation(t, which)
    void filter_series_with_expectations(const std::vector<float_t>& obs, const std::vector<float_t>& cov, std::uint64_t stream_id = 0)
    {
        if (obs.empty() || obs.size() != cov.size()) throw std::length_error("observations and covariates must be non-empty and of equal length");
        std::vector<double> rows(2 * obs.size());
        for (size_t t = 0; t < obs.size(); ++t) { rows[2 * t] = (double)obs[t]; rows[2 * t + 1] = (double)cov[t]; }
        throw_on_error(ssme_b200_replace_observations(m_h, rows.data(), obs.size(), 2));
        m_streaming = false;
        m_cond_like.assign(obs.size(), 0.0);
        m_expect.assign(obs.size() * 5, 0.0);
        double lo[4], hi[4];
        for (int k = 0; k < 4; ++k) { lo[k] = (double)m_lo(k); hi[k] = (double)m_hi(k); }
        throw_on_error(ssme_b200_lw_expectations(m_h, form, lo, hi, (double)m_delta, stream_id, &m_loglik, m_cond_like.data(), m_expect.data()));
    }
    float_t getExpectation(size_t t, size_t which) const { return (float_t)m_expect.at(5 * t + which); }

    // The reference's streaming call: filter(obs_data, cov_data) once per observation (:971, :2191).  The first call
    // starts the run (stream id fixed by set_stream); results equal filter_series on the same data bit for bit.
    void set_stream(std::uint64_t stream_id) { m_stream = stream_id; }
    void filter(float_t y_t, float_t z_t)
    {
        if (!m_streaming) {
            double lo[4], hi[4];
            for (int k = 0; k < 4; ++k) { lo[k] = (double)m_lo(k); hi[k] = (double)m_hi(k); }
            throw_on_error(ssme_b200_lw_begin(m_h, form, lo, hi, (double)m_delta, m_stream));
            m_cond_like.clear();
            m_theta_bar.clear();
            m_streaming = true;
        }
        double cl = 0.0, tb[4];
        throw_on_error(ssme_b200_lw_step(m_h, (double)y_t, (double)z_t, &cl, tb));
        m_cond_like.push_back(cl);
        m_theta_bar.insert(m_theta_bar.end(), tb, tb + 4);
        throw_on_error(ssme_b200_lw_state(m_h, &m_loglik, m_final_mean.data(), nullptr));
    }

    // *FutureSimulator::sim_future_obs(num_steps, last_obs) (liu_west_filter.h:1315-1360): from the particles of the streaming
    // run, num_steps simulated observations per particle; result[time][particle] as the reference's obsSamples.  The filter's
    // state is not modified.  Every call draws fresh randomness unless the same sim_stream is passed again.
    std::vector<std::vector<float_t>> sim_future_obs(unsigned num_steps, float_t last_obs, std::uint64_t sim_stream = 0)
    {
        if (!m_streaming) throw std::runtime_error("sim_future_obs follows filter(y_t, z_t)");
        std::vector<double> flat((size_t)num_steps * nparts);
        throw_on_error(ssme_b200_lw_sim_future(m_h, num_steps, (double)last_obs, sim_stream ? sim_stream : ++m_sim_calls, flat.data()));
        std::vector<std::vector<float_t>> out(num_steps, std::vector<float_t>(nparts));
        for (size_t t = 0; t < num_steps; ++t)
            for (size_t i = 0; i < nparts; ++i) out[t][i] = (float_t)flat[t * nparts + i];
        return out;
    }

    // log p(y_t | y_{1:t-1}) of step t (the reference returns the latest one; :2180-2184)
    float_t getLogCondLike(size_t t) const { return (float_t)m_cond_like.at(t); }
    float_t getLogCondLike() const { return (float_t)m_cond_like.back(); }
    float_t getLogLike() const { return (float_t)m_loglik; }
    // mean of the untransformed parameter particles after the last step: E[theta | y_{1:T}]
    psv getParamMeans() const
    {
        psv r;
        for (int k = 0; k < 4; ++k) r(k) = (float_t)m_final_mean[k];
        return r;
    }

private:
    ssme_b200_handle m_h = nullptr;
    float_t m_delta;
    psv m_lo, m_hi;
    double m_loglik = 0.0;
    bool m_streaming = false;
    std::uint64_t m_stream = 0, m_sim_calls = 0;
    std::vector<double> m_cond_like, m_theta_bar, m_expect;
    std::array<double, 4> m_final_mean{};
};
}  // namespace detail

template <size_t nparts, typename float_t>
class LWFilter2_svol : public detail::lw_svol_base<nparts, float_t, SSME_B200_LW_SISR> {
public:
    using base = detail::lw_svol_base<nparts, float_t, SSME_B200_LW_SISR>;
    using base::base;
    void filter(float_t y_t) { base::filter(y_t, float_t(0)); }
    void filter_series(const std::vector<float_t>& obs, std::uint64_t stream_id = 0)
    {
        base::filter_series(obs, std::vector<float_t>(obs.size(), float_t(0)), stream_id);
    }
};

template <size_t nparts, typename float_t>
class LWFilter_svol : public detail::lw_svol_base<nparts, float_t, SSME_B200_LW_APF> {
public:
    using base = detail::lw_svol_base<nparts, float_t, SSME_B200_LW_APF>;
    using base::base;
    void filter(float_t y_t) { base::filter(y_t, float_t(0)); }
    void filter_series(const std::vector<float_t>& obs, std::uint64_t stream_id = 0)
    {
        base::filter_series(obs, std::vector<float_t>(obs.size(), float_t(0)), stream_id);
    }
};

}  // namespace ssme_b200
#endif
