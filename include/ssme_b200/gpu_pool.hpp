// include/ssme_b200/gpu_pool.hpp -- thread_pool-shaped front end of the GPU likelihood backend.
//
// Same three-method shape as the reference dispatcher (include/ssme/thread_pool.h):
//   ctor(num_comps, ...)          thread_pool.h:118   (f is replaced by a device model id)
//   add_observed_data(data)       thread_pool.h:166   once; a second call throws std::runtime_error
//   work(theta) -> float_t        thread_pool.h:189   num_comps replicate filters + log-mean-exp (:263-268)
// plus work_batch(thetas) for callers that hold several proposals (multi-chain PMMH, swarm).
// Status codes of the C ABI are re-thrown as the exception types the reference throws.
#ifndef SSME_B200_GPU_POOL_HPP
#define SSME_B200_GPU_POOL_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../ssme_b200.h"
#include "fixed.hpp"
#include "parameters.hpp"

namespace ssme_b200 {

inline void throw_on_error(int rc)
{
    if (rc == SSME_B200_OK) return;
    const std::string msg = ssme_b200_last_error();
    switch (rc) {
    case SSME_B200_EINVAL: throw std::invalid_argument(msg);
    case SSME_B200_ELENGTH: throw std::length_error(msg);
    default: throw std::runtime_error(msg);
    }
}

struct gpu_options {
    int device = 0;
    int model = SSME_B200_MODEL_SV;
    int resampler = SSME_B200_RESAMP_MULTINOMIAL;
    int resample_every = 1;
    std::uint64_t seed = 20260101;  // Philox key of the filters' random streams
    int scan_items_per_lane = 0, threads_per_filter = 0;
    int use_cluster = 0;  // 1 = one filter per thread-block cluster (fewer filters than SMs: lower step latency)
};

template <size_t numparams, size_t dimobs, typename float_t, bool debug = false>
class gpu_pool {
public:
    using dyn_data_t = param::pack<float_t, numparams>;
    using osv = vec<float_t, dimobs>;
    using static_data_t = std::vector<osv>;

    gpu_pool() = delete;
    gpu_pool(const gpu_pool&) = delete;
    gpu_pool& operator=(const gpu_pool&) = delete;

    // num_comps: filters per proposal (ada_pmmh_mvn's num_pfilters); num_particles: nparts of the filter
    gpu_pool(unsigned num_comps, unsigned num_particles, const gpu_options& opt = gpu_options())
        : m_total_calcs(num_comps), m_next_stream(0)
    {
        if (num_comps == 0) throw std::invalid_argument("num_comps must be positive");
        ssme_b200_config c{};
        c.struct_size = (int32_t)sizeof(c);
        c.device = opt.device;
        c.model = opt.model;
        c.num_particles = (int32_t)num_particles;
        c.resampler = opt.resampler;
        c.resample_every = opt.resample_every;
        c.dtype = SSME_B200_DTYPE_F64;
        c.rng_mode = SSME_B200_RNG_PHILOX;
        c.seed = opt.seed;
        c.scan_items_per_lane = opt.scan_items_per_lane;
        c.threads_per_filter = opt.threads_per_filter;
        c.use_cluster = opt.use_cluster;
        throw_on_error(ssme_b200_create(&c, &m_h));
    }
    ~gpu_pool() { ssme_b200_destroy(m_h); }

    void add_observed_data(const static_data_t& obs_data)
    {
        std::vector<double> flat(obs_data.size() * dimobs);
        for (size_t t = 0; t < obs_data.size(); ++t)
            for (size_t k = 0; k < dimobs; ++k) flat[t * dimobs + k] = (double)obs_data[t](k);
        throw_on_error(ssme_b200_set_observations(m_h, flat.data(), obs_data.size(), dimobs));
    }

    // thread_pool::work: log-mean-exp of num_comps particle-filter log-likelihood estimates at new_param
    float_t work(dyn_data_t new_param)
    {
        const auto th = new_param.get_untrans_params();
        double in[numparams], out = 0.0;
        for (size_t k = 0; k < numparams; ++k) in[k] = (double)th(k);
        throw_on_error(ssme_b200_loglike_batch(m_h, in, 1, m_total_calcs, m_next_stream, &out, nullptr));
        m_next_stream += m_total_calcs;
        return (float_t)out;
    }

    // P proposals in one launch; returns P log-mean-exp estimates
    std::vector<float_t> work_batch(const std::vector<dyn_data_t>& params)
    {
        const size_t P = params.size();
        std::vector<double> in(P * numparams), out(P);
        for (size_t p = 0; p < P; ++p) {
            const auto th = params[p].get_untrans_params();
            for (size_t k = 0; k < numparams; ++k) in[p * numparams + k] = (double)th(k);
        }
        throw_on_error(ssme_b200_loglike_batch(m_h, in.data(), P, m_total_calcs, m_next_stream, out.data(), nullptr));
        m_next_stream += (std::uint64_t)P * m_total_calcs;
        return std::vector<float_t>(out.begin(), out.end());
    }

    // untransformed thetas, row-major [P][numparams]; explicit stream ids (reproducible multi-rank runs)
    void work_raw(const double* theta, size_t P, std::uint64_t stream_base, double* out, double* per_filter = nullptr)
    {
        throw_on_error(ssme_b200_loglike_batch(m_h, theta, P, m_total_calcs, stream_base, out, per_filter));
    }

    unsigned num_comps() const { return m_total_calcs; }
    ssme_b200_handle handle() const { return m_h; }
    void set_next_stream(std::uint64_t s) { m_next_stream = s; }

private:
    ssme_b200_handle m_h = nullptr;
    const unsigned m_total_calcs;
    std::uint64_t m_next_stream;
};

}  // namespace ssme_b200
#endif
