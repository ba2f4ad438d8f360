// examples/estimate_univ_svol.hpp -- the reference's only end-to-end program, on the GPU backend.
// Mirrors example/estimate_univ_svol.h of the reference: univ_svol_estimator supplies the prior
// (beta ~ N(1,1), phi ~ U(0,1), sigma^2 ~ InvGamma(.001,.001), :95-101); the bootstrap filter for the
// univariate SV model (example/univ_svol_bootstrap_filter.h) is the device model SSME_B200_MODEL_SV.
#ifndef SSME_B200_EXAMPLE_ESTIMATE_UNIV_SVOL_HPP
#define SSME_B200_EXAMPLE_ESTIMATE_UNIV_SVOL_HPP

#include <string>
#include <vector>

#include <ssme_b200/ada_pmmh_mvn.hpp>
#include <ssme_b200/rv_eval.hpp>

template <size_t numparams, size_t dimstate, size_t dimobs, size_t numparts, typename float_t>
class univ_svol_estimator : public ada_pmmh_mvn<numparams, dimobs, numparts, float_t> {
public:
    using base = ada_pmmh_mvn<numparams, dimobs, numparts, float_t>;
    using psv = typename base::psv;
    using psm = typename base::psm;
    using base::base;

    float_t log_prior_eval(const param::pack<float_t, 3>& theta) override
    {
        namespace rveval = ssme_b200::rveval;
        float_t returnThis(0.0);
        const float_t beta = theta.get_untrans_params(0, 0)[0];
        const float_t phi = theta.get_untrans_params(1, 1)[0];
        const float_t ss = theta.get_untrans_params(2, 2)[0];
        returnThis += rveval::evalUnivNorm<float_t>(beta, 1.0, 1.0, true);
        returnThis += rveval::evalUniform<float_t>(phi, 0.0, 1.0, true);
        returnThis += rveval::evalUnivInvGamma<float_t>(ss, .001, .001, true);
        return returnThis;
    }
};

// do_ada_pmmh_univ_svol of the reference (:139-178): same start values, transforms, C0, t0, t1
template <size_t numparams, size_t dimstate, size_t dimobs, size_t numparts, typename float_t>
void do_ada_pmmh_univ_svol(const std::string& datafile, const std::string& samples_base_name, const std::string& messages_base_name,
                           unsigned int num_mcmc_iters, unsigned int num_pfilters, bool multicore,
                           const ssme_b200::gpu_options& gpu = ssme_b200::gpu_options(), unsigned long proposal_seed = 0)
{
    using est = univ_svol_estimator<numparams, dimstate, dimobs, numparts, float_t>;
    typename est::psv start_trans_theta{(float_t)1.0, ssme_b200::rveval::twiceFisher<float_t>(.5), (float_t)std::log(2.0e-4)};
    std::vector<std::string> tts{"null", "twice_fisher", "log"};  // beta, phi, sigma squared
    typename est::psm C0 = est::psm::Identity() * (float_t).15;
    unsigned int t0 = 150, t1 = 1000;
    est mcmcobj(start_trans_theta, tts, num_mcmc_iters, num_pfilters, datafile, samples_base_name, messages_base_name, multicore, t0,
                t1, C0, false, 1, 0, gpu, proposal_seed);
    mcmcobj.commence_sampling();
}

#endif
