// include/ssme_b200/pswarm_filter.hpp -- the "particle swarm filter" of SSME on the GPU backend.
//
// Mirrors Swarm / SwarmWithCovs (include/ssme/pswarm_filter.h:22-305, 309-605): `nparamparts` bootstrap filters, each with
// its own parameter vector drawn ONCE (samp_untrans_params, called from finish_construction :280-304), advanced over the
// observations; per observation the swarm reports the average over the filters of their log conditional likelihoods
// (comp_func :86-92, aggregation :96-160 -- an average of logs).  The reference streams one observation per `update`
// call through split_data_thread_pool; here the whole series is filtered by one launch (one filter per CTA) and the
// per-step values are read back.  Expectations E[h(x_t) | y_{1:t}] (getExpectations) are not computed on the device yet.
#ifndef SSME_B200_PSWARM_FILTER_HPP
#define SSME_B200_PSWARM_FILTER_HPP

#include <functional>
#include <stdexcept>
#include <vector>

#include "../ssme_b200.h"
#include "gpu_pool.hpp"

namespace ssme_b200 {

template <size_t nstateparts, size_t nparamparts, size_t dimparam, typename float_t = double>
class Swarm {
public:
    using psv = vec<float_t, dimparam>;

    explicit Swarm(const gpu_options& opt = gpu_options()) : m_num_obs(0)
    {
        const size_t want = (opt.model == SSME_B200_MODEL_SV) ? 3 : 4;
        if (dimparam != want) throw std::invalid_argument("dimparam does not match the device model");
        ssme_b200_config c{};
        c.struct_size = (int32_t)sizeof(c);
        c.device = opt.device;
        c.model = opt.model;
        c.num_particles = (int32_t)nstateparts;
        c.resampler = opt.resampler;
        c.resample_every = 1;
        c.dtype = SSME_B200_DTYPE_F64;
        c.rng_mode = SSME_B200_RNG_PHILOX;
        c.seed = opt.seed;
        throw_on_error(ssme_b200_create(&c, &m_h));
    }
    virtual ~Swarm() { ssme_b200_destroy(m_h); }
    Swarm(const Swarm&) = delete;
    Swarm& operator=(const Swarm&) = delete;

    // the reference's pure virtual: one untransformed parameter draw per model (pswarm_filter.h:262)
    virtual psv samp_untrans_params() = 0;

    // update(y_t) for t = 0 .. T-1 in one call; obs is row-major [T][dimy], dimy = 1 or 2 (y_t, z_t)
    // with_expectations: also form E[h_k(x_t) | y_{1:t}] averaged over the parameter particles (Swarm::update(yt, fs) +
    // getExpectations, pswarm_filter.h:223-239, 96-160).  The functions h_k are members of the device model type
    // (csrc/models/model_api.cuh); num_expectations() of them, h(x) = x and h(x) = x^2 unless the model brings its own.
    void update_series(const std::vector<double>& obs, size_t dimy, std::uint64_t stream_base = 0, bool with_expectations = false)
    {
        if (obs.empty() || obs.size() % dimy != 0) throw std::length_error("bad observation array");
        const size_t T = obs.size() / dimy;
        std::vector<double> theta(nparamparts * dimparam);
        for (size_t j = 0; j < nparamparts; ++j) {  // finish_construction: draw every model's parameters once
            const psv p = samp_untrans_params();
            for (size_t k = 0; k < dimparam; ++k) theta[j * dimparam + k] = (double)p(k);
        }
        throw_on_error(ssme_b200_replace_observations(m_h, obs.data(), T, dimy));
        m_log_cond_like.assign(T, 0.0);
        if (with_expectations) {
            m_expectations.assign(num_expectations() * T, 0.0);
            throw_on_error(ssme_b200_swarm_expectations(m_h, theta.data(), nparamparts, stream_base, m_log_cond_like.data(),
                                                        m_expectations.data(), nullptr));
        } else {
            m_expectations.clear();
            throw_on_error(ssme_b200_swarm_filter(m_h, theta.data(), nparamparts, stream_base, m_log_cond_like.data(), nullptr));
        }
        m_num_obs = (unsigned)T;
    }

    // The reference's streaming call: update(y_t) once per observation (pswarm_filter.h:223-239); row = (y_t) or (y_t, z_t).
    // The first call draws every model's parameters (finish_construction) and fixes the random streams.
    void update(const std::vector<double>& row, bool with_expectations = false, std::uint64_t stream_base = 0)
    {
        if (!m_streaming) {
            std::vector<double> theta(nparamparts * dimparam);
            for (size_t j = 0; j < nparamparts; ++j) {
                const psv p = samp_untrans_params();
                for (size_t k = 0; k < dimparam; ++k) theta[j * dimparam + k] = (double)p(k);
            }
            throw_on_error(ssme_b200_swarm_begin(m_h, theta.data(), nparamparts, stream_base));
            m_log_cond_like.clear();
            m_expectations.clear();
            m_num_obs = 0;
            m_streaming = true;
        }
        double cl = 0.0, ex[8] = {0.0};
        throw_on_error(ssme_b200_swarm_step(m_h, row.data(), &cl, with_expectations ? ex : nullptr));
        m_log_cond_like.push_back(cl);
        if (with_expectations) m_expectations.insert(m_expectations.end(), ex, ex + num_expectations());
        m_num_obs += 1;
    }

    float_t getLogCondLike(size_t t) const { return (float_t)m_log_cond_like.at(t); }
    float_t getLogCondLike() const { return (float_t)m_log_cond_like.back(); }
    unsigned num_obs() const { return m_num_obs; }
    // size of the reference's std::vector<func> of expectation callbacks (pswarm_filter.h:47): fixed by the device model type
    size_t num_expectations() const { return (size_t)ssme_b200_num_expectations(m_h); }
    // getExpectations()[which] at step t: E[h_which(x_t) | y_{1:t}] (default functions: which = 0 -> x_t, 1 -> x_t^2)
    float_t getExpectation(size_t t, size_t which) const { return (float_t)m_expectations.at(num_expectations() * t + which); }

private:
    ssme_b200_handle m_h = nullptr;
    std::vector<double> m_log_cond_like, m_expectations;
    unsigned m_num_obs = 0;
    bool m_streaming = false;
};

}  // namespace ssme_b200
#endif
