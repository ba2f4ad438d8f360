// oracle/refshim/pf/pf_base.h -- TEST INFRASTRUCTURE.  Stand-in for the pf base classes the reference names:
// pf::bases::pf_base / pf_withcov_base (include/ssme/pswarm_filter.h:29,41,44,59,352: typedefs float_type,
// dynamic_matrix, func) and GenFutureSimulator (test/test_pswarm.cpp:25,34).  Interface only; restated from pf's
// published API (tbrown122387/pf, unpinned, absent).
#ifndef SSME_REFSHIM_PF_BASE_H
#define SSME_REFSHIM_PF_BASE_H
#include <Eigen/Dense>
#include <array>
#include <functional>
#include <vector>

namespace pf {
namespace bases {

template <typename float_t, size_t dimobs, size_t dimstate>
class pf_base {
public:
    using float_type = float_t;
    using observation_sized_vector = Eigen::Matrix<float_t, dimobs, 1>;
    using state_sized_vector = Eigen::Matrix<float_t, dimstate, 1>;
    using dynamic_matrix = Eigen::Matrix<float_t, Eigen::Dynamic, Eigen::Dynamic>;
    using func = std::function<const dynamic_matrix(const state_sized_vector&)>;
    using func_vec = std::vector<func>;
    static constexpr unsigned int dim_obs = dimobs;
    static constexpr unsigned int dim_state = dimstate;
    virtual void filter(const observation_sized_vector& data, const func_vec& fs = func_vec()) = 0;
    virtual float_t getLogCondLike() const = 0;
    virtual std::vector<dynamic_matrix> getExpectations() const = 0;
    virtual ~pf_base() = default;
};

template <typename float_t, size_t dimobs, size_t dimstate, size_t dimcov>
class pf_withcov_base {
public:
    using float_type = float_t;
    using observation_sized_vector = Eigen::Matrix<float_t, dimobs, 1>;
    using state_sized_vector = Eigen::Matrix<float_t, dimstate, 1>;
    using cov_sized_vector = Eigen::Matrix<float_t, dimcov, 1>;
    using dynamic_matrix = Eigen::Matrix<float_t, Eigen::Dynamic, Eigen::Dynamic>;
    using func = std::function<const dynamic_matrix(const state_sized_vector&, const cov_sized_vector&)>;
    using func_vec = std::vector<func>;
    static constexpr unsigned int dim_obs = dimobs;
    static constexpr unsigned int dim_state = dimstate;
    static constexpr unsigned int dim_cov = dimcov;
    virtual void filter(const observation_sized_vector& data, const cov_sized_vector& cov, const func_vec& fs = func_vec()) = 0;
    virtual float_t getLogCondLike() const = 0;
    virtual std::vector<dynamic_matrix> getExpectations() const = 0;
    virtual ~pf_withcov_base() = default;
};

// forward simulation add-on: out of the hot path; kept so the reference's tests compile
template <size_t dimx, size_t dimy, typename float_t, size_t nparts>
class GenFutureSimulator {
public:
    using ssv = Eigen::Matrix<float_t, dimx, 1>;
    using osv = Eigen::Matrix<float_t, dimy, 1>;
    virtual std::array<ssv, nparts> get_uwtd_samps() const = 0;
    virtual ssv fSamp(const ssv& xtm1, const osv& ytm1) = 0;
    virtual osv gSamp(const ssv& xt) = 0;
    virtual ~GenFutureSimulator() = default;
    std::vector<std::array<osv, nparts>> sim_future_obs(unsigned int num_future_steps, const osv& yt)
    {
        std::array<ssv, nparts> states = get_uwtd_samps();
        std::vector<std::array<osv, nparts>> out(num_future_steps);
        std::array<osv, nparts> prev;
        prev.fill(yt);
        for (unsigned int s = 0; s < num_future_steps; ++s)
            for (size_t j = 0; j < nparts; ++j) {
                states[j] = fSamp(states[j], prev[j]);
                out[s][j] = gSamp(states[j]);
                prev[j] = out[s][j];
            }
        return out;
    }
};

}  // namespace bases
}  // namespace pf
#endif
