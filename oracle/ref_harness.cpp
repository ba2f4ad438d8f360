// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// C-ABI harness around the reference's OWN headers, compiled UNMODIFIED from where they lie under /root/reference
// (oracle/Makefile target `refhdr` -> oracle/_ref/libssme_refhdr.so; nothing of the reference is copied into this repo):
//   include/ssme/parameters.h        param::transform / param::pack                      -> ssme_refhdr_transform, _pack
//   include/ssme/liu_west_filter.h   mn_resamp_states_and_params::resampLogWts (:91-145)  -> ssme_refhdr_resample_sorted
//                                    LWFilter2::filter (:1608-1761)                       -> ssme_refhdr_lwfilter2_sv
//                                    LWFilter2WithCovs::filter (:2191-2343), LWFilterWithCovs::filter (:971-1159),
//                                    update_parameter_proposal_components                  -> ssme_refhdr_lw_leverage
//   include/ssme/ada_pmmh_mvn.h      commence_sampling, update_moments_and_Ct, q_samp      -> ssme_refhdr_pmmh_chain
//   include/ssme/thread_pool.h       thread_pool::work (log-mean-exp)                     -> (through ada_pmmh_mvn)
//   example/univ_svol_bootstrap_filter.h  svol_bs model hooks (:54-103)                    -> ssme_refhdr_bsfilter_sv
//   test/test_liu_west.cpp           the reference's own test models svol_lw_1_par / svol_lw_2_par (:26-157, :213-358)
// What is NOT reference code here: Eigen3 and tbrown122387/pf are absent from the image, so their surface is a stand-in
// under oracle/refshim/ (Eigen/Dense; pf/rv_samp.h, rv_eval.h, resamplers.h, bootstrap_filter.h -- see each header).
// Randomness: the pf stand-in plays back the streams the caller passes (oracle/refshim/pf/shim_streams.h); the
// reference's in-tree resampler owns a private std::mt19937 (liu_west_filter.h:66), which is reseeded per step through
// `#define private public` (test-only) and whose uniforms are reproduced here for the caller.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <future>
#include <iostream>
#include <limits>
#include <memory>
#include <mutex>
#include <random>
#include <shared_mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <atomic>
#include <tuple>
#include <utility>
#include <vector>

#include <Eigen/Dense>
#include <catch2/catch_all.hpp>
#include <pf/bootstrap_filter.h>
#include <pf/resamplers.h>
#include <pf/rv_eval.h>
#include <pf/rv_samp.h>

// test-only access to the reference classes' private state (resampler engine, particle arrays)
#define private public
#define protected public
#include <ssme/parameters.h>
#include <ssme/liu_west_filter.h>
#include <ssme/ada_pmmh_mvn.h>
#include <estimate_univ_svol.h>          // example/ (univ_svol_estimator; pulls in univ_svol_bootstrap_filter.h: svol_bs)
#include <test_liu_west.cpp>             // test/    (svol_lw_1_par, svol_lw_2_par; its TEST_CASEs only register)
#undef private
#undef protected

namespace {

const char* const kTransNames[4] = {"null", "twice_fisher", "logit", "log"};  // oracle / C-ABI numbering

thread_local std::string g_err;
int fail(const std::exception& e)
{
    g_err = e.what();
    return -1;
}

// the uniforms std::uniform_real_distribution<double>(0,1) draws from std::mt19937{seed}: what
// mn_resamp_states_and_params::resampLogWts consumes (liu_west_filter.h:110,113), N + 1 per call
void resampler_uniforms(std::uint32_t seed, int n, double* out)
{
    std::mt19937 g{seed};
    std::uniform_real_distribution<double> u(0.0, 1.0);
    for (int i = 0; i < n; ++i) out[i] = u(g);
}

template <size_t N>
int resample_sorted(const double* lw, std::uint32_t seed, int32_t* anc, double* uniforms)
{
    mn_resamp_states_and_params<N, 1, 1, double> rs(seed);
    typename mn_resamp_states_and_params<N, 1, 1, double>::arrayVec states;
    typename mn_resamp_states_and_params<N, 1, 1, double>::arrayParams params;
    typename mn_resamp_states_and_params<N, 1, 1, double>::arrayFloat w;
    Eigen::Matrix<double, 1, 1> one;
    one(0) = 0.0;
    for (size_t i = 0; i < N; ++i) {
        states[i](0) = (double)i;  // the state carries its own index: after resampling it is the ancestor
        params[i] = param::pack<double, 1>(one, std::vector<std::string>{"null"});
        w[i] = lw[i];
    }
    rs.resampLogWts(states, params, w);
    for (size_t i = 0; i < N; ++i) {
        anc[i] = (int32_t)states[i](0);
        if (w[i] != 0.0) throw std::runtime_error("log-weights were not reset");
    }
    if (uniforms) resampler_uniforms(seed, (int)N + 1, uniforms);
    return 0;
}

// ---- SV model on the reference's LWFilter2 (SISR step, liu_west_filter.h:1608-1761) ---------------------------------------
// Hooks follow example/univ_svol_bootstrap_filter.h:64-103 with (beta, phi, sigma^2) read from the particle's parameter.
// With delta = 1 the shrinkage is a = 1 and h^2 = 0 (liu_west_filter.h:1583,1776): every particle keeps its parameter
// exactly, so LWFilter2 IS the bootstrap/SISR filter the external pf::BSFilter implements.
template <size_t N>
class sv_on_lwfilter2 : public LWFilter2<N, 1, 1, 3, double> {
public:
    using base = LWFilter2<N, 1, 1, 3, double>;
    using ssv = typename base::ssv;
    using osv = typename base::osv;
    using psv = typename base::psv;
    sv_on_lwfilter2(const psv& theta, double delta, unsigned rs) : base({"null", "null", "null"}, delta, rs), m_theta(theta) {}
    double logMuEv(const ssv& x1, const psv& p) override
    {
        return rveval::evalUnivNorm<double>(x1(0), 0.0, std::sqrt(p(2)) / std::sqrt(1.0 - p(1) * p(1)), true);
    }
    ssv q1Samp(const osv&, const psv& p) override
    {
        ssv x;
        x(0) = m_z.sample() * std::sqrt(p(2)) / std::sqrt(1. - p(1) * p(1));
        return x;
    }
    double logQ1Ev(const ssv& x1, const osv&, const psv& p) override
    {
        return rveval::evalUnivNorm<double>(x1(0), 0.0, std::sqrt(p(2)) / std::sqrt(1.0 - p(1) * p(1)), true);
    }
    double logGEv(const osv& yt, const ssv& xt, const psv& p) override
    {
        return rveval::evalUnivNorm<double>(yt(0), 0.0, p(0) * std::exp(.5 * xt(0)), true);
    }
    double logFEv(const ssv& xt, const ssv& xtm1, const psv& p) override
    {
        return rveval::evalUnivNorm<double>(xt(0), p(1) * xtm1(0), std::sqrt(p(2)), true);
    }
    ssv qSamp(const ssv& xtm1, const osv&, const psv& p) override
    {
        ssv x;
        x(0) = p(1) * xtm1(0) + m_z.sample() * std::sqrt(p(2));
        return x;
    }
    double logQEv(const ssv& xt, const ssv& xtm1, const osv&, const psv& p) override { return logFEv(xt, xtm1, p); }
    psv paramPriorSamp() override { return m_theta; }

private:
    psv m_theta;
    rvsamp::UnivNormSampler<double> m_z;
};

template <size_t N>
int lwfilter2_sv(const double* theta, const double* y, int T, int rs, double delta, const double* z, const uint32_t* seeds,
                 double* cond_like, double* x_post, double* u_out)
{
    typename sv_on_lwfilter2<N>::psv th;
    th << theta[0], theta[1], theta[2];
    sv_on_lwfilter2<N> f(th, delta, (unsigned)rs);
    pf::shim::normal_stream().arm(z, (size_t)T * N);
    std::vector<double> zeros((size_t)3 * N, 0.0);
    for (int t = 0; t < T; ++t) {
        f.m_resampler.m_gen.seed(seeds[t]);
        if (u_out) resampler_uniforms(seeds[t], (int)N + 1, u_out + (size_t)t * (N + 1));
        pf::shim::mvn_stream().arm(zeros.data(), zeros.size());  // the jitter normals multiply a zero factor when delta = 1
        typename sv_on_lwfilter2<N>::osv yt;
        yt(0) = y[t];
        f.filter(yt);
        cond_like[t] = f.getLogCondLike();
        for (size_t i = 0; i < N; ++i) x_post[(size_t)t * N + i] = f.m_state_particles[i](0);
    }
    pf::shim::normal_stream().disarm();
    pf::shim::mvn_stream().disarm();
    return 0;
}

// ---- the example's own model class on the pf stand-in (BSFilter + mn_resampler) -------------------------------------
template <size_t N>
int bsfilter_sv(const double* theta, const double* y, int T, const double* z, const double* u, double* cond_like, double* x_post)
{
    using svol_t = svol_bs<N, 1, 1, pf::resamplers::mn_resampler<N, 1, double>, double>;
    struct mod_t : svol_t {  // read access to the particle array (protected in BSFilter)
        using svol_t::svol_t;
        double particle(size_t i) const { return this->m_particles[i](0); }
    };
    Eigen::Matrix<double, 3, 1> th;
    th << theta[0], theta[1], theta[2];
    param::pack<double, 3> pp(th, std::vector<std::string>{"null", "null", "null"});
    mod_t mod(pp);  // svol_bs(const pack&): beta, phi, sqrt(ss)  (univ_svol_bootstrap_filter.h:54-61)
    if (z) pf::shim::normal_stream().arm(z, (size_t)T * N);  // NULL streams: the samplers' own mt19937 (independent RNG)
    if (u) pf::shim::resamp_stream().arm(u, (size_t)T * N);
    for (int t = 0; t < T; ++t) {
        Eigen::Matrix<double, 1, 1> yt;
        yt(0) = y[t];
        mod.filter(yt);
        cond_like[t] = mod.getLogCondLike();
        if (x_post)
            for (size_t i = 0; i < N; ++i) x_post[(size_t)t * N + i] = mod.particle(i);
    }
    pf::shim::normal_stream().disarm();
    pf::shim::resamp_stream().disarm();
    return 0;
}

// ---- Liu-West filters on the reference's own test models -----------------------------------------------------------
struct lw_out {
    double* cond_like;   // [T]
    double* theta_bar;   // [T][4]  m_thetaBar entering step t (row 0 unused)
    double* x_post;      // [T][N]  states after the step's resampling
    double* th_post;     // [T][N][4] transformed parameters after the step's resampling
    double* expect;      // [T][5]  E[x], E[phi], E[mu], E[sigma], E[rho] before resampling
    double* u_resamp;    // [T][N+1] the uniforms the in-tree resampler consumed
};

template <typename Filter, size_t N>
int lw_run(Filter& f, const double* y, const double* cov, int T, const double* u_prior, const double* z_state, const double* z_jitter,
           const double* u_aux, const uint32_t* seeds, const lw_out& o)
{
    using ssv = Eigen::Matrix<double, 1, 1>;
    using psv = Eigen::Matrix<double, 4, 1>;
    using Mat = Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic>;
    using func = std::function<const Mat(const ssv&, const ssv&, const psv&)>;
    std::vector<func> fs;
    if (o.expect)
        fs.push_back([](const ssv& xt, const ssv&, const psv& pt) -> const Mat {
            Mat m(5, 1);
            m(0, 0) = xt(0);
            for (int k = 0; k < 4; ++k) m(k + 1, 0) = pt(k);
            return m;
        });
    pf::shim::uniform_stream().arm(u_prior, (size_t)N * 4);
    pf::shim::normal_stream().arm(z_state, (size_t)T * N);
    for (int t = 0; t < T; ++t) {
        f.m_resampler.m_gen.seed(seeds[t]);
        if (o.u_resamp) resampler_uniforms(seeds[t], (int)N + 1, o.u_resamp + (size_t)t * (N + 1));
        if (t > 0) {
            pf::shim::mvn_stream().arm(z_jitter + (size_t)t * N * 4, (size_t)N * 4);
            if (u_aux) pf::shim::kgen_stream().arm(u_aux + (size_t)t * N, (size_t)N);
        }
        ssv yt, zt;
        yt(0) = y[t];
        zt(0) = (t > 0) ? (cov ? cov[t] : y[t - 1]) : 0.0;
        f.filter(yt, zt, fs);
        if (o.cond_like) o.cond_like[t] = f.getLogCondLike();
        if (o.theta_bar && t > 0)
            for (int k = 0; k < 4; ++k) o.theta_bar[(size_t)t * 4 + k] = f.m_thetaBar(k);
        if (o.expect) {
            const auto e = f.getExpectations();
            for (int k = 0; k < 5; ++k) o.expect[(size_t)t * 5 + k] = e[0](k, 0);
        }
        for (size_t i = 0; i < N; ++i) {
            if (o.x_post) o.x_post[(size_t)t * N + i] = f.m_state_particles[i](0);
            if (o.th_post) {
                const auto tp = f.m_param_particles[i].get_trans_params();
                for (int k = 0; k < 4; ++k) o.th_post[((size_t)t * N + i) * 4 + k] = tp(k);
            }
        }
    }
    pf::shim::uniform_stream().disarm();
    pf::shim::normal_stream().disarm();
    pf::shim::mvn_stream().disarm();
    pf::shim::kgen_stream().disarm();
    return 0;
}

// rs: the resampling schedule.  The test model's constructor does not forward one (test/test_liu_west.cpp:257-263 calls the base
// with (transforms, delta)), so the base class's member is set directly (test-only access, see the #define above): the filter
// code that runs is the reference's LWFilter2WithCovs::filter, which resamples when (m_now + 1) % m_resampSched == 0 (:2272, :2340).
template <size_t N>
int lw_leverage_form0(const double* lo, const double* hi, double delta, const double* y, const double* cov, int T, const double* u_prior,
                      const double* z_state, const double* z_jitter, const uint32_t* seeds, const lw_out& o, unsigned rs = 1)
{
    svol_lw_2_par<N, double> f(delta, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], lo[3], hi[3], 1);
    f.m_resampSched = rs;
    return lw_run<svol_lw_2_par<N, double>, N>(f, y, cov, T, u_prior, z_state, z_jitter, nullptr, seeds, o);
}

// ---- the reference's PMMH loop on a closed-form "likelihood" --------------------------------------------------------
// log_like_eval returns a deterministic function of theta, so the chain depends only on the proposal / accept draws:
// the reference's commence_sampling, update_moments_and_Ct and q_samp are compared with ours on identical streams.
class toy_pmmh : public ada_pmmh_mvn<3, 1, 8, double> {
public:
    using base = ada_pmmh_mvn<3, 1, 8, double>;
    using base::base;
    double log_prior_eval(const param::pack<double, 3>& theta) override
    {
        const auto p = theta.get_untrans_params();
        double r = rveval::evalUnivNorm<double>(p(0), 1.0, 1.0, true);
        r += rveval::evalUniform<double>(p(1), 0.0, 1.0, true);
        r += rveval::evalUnivInvGamma<double>(p(2), .001, .001, true);
        return r;
    }
    double log_like_eval(const param::pack<double, 3>& theta, const std::vector<osv>& data) override
    {
        const auto p = theta.get_untrans_params();
        double s = 0.0;
        for (size_t i = 0; i < data.size(); ++i) s += -0.5 * (data[i](0) - p(0)) * (data[i](0) - p(0));
        const double l2 = std::log(p(2)) + 1.0;
        return s - 2.0 * (p(1) - 0.3) * (p(1) - 0.3) - 0.5 * l2 * l2;
    }
};

// ---- the reference's example estimator (example/estimate_univ_svol.h:17-131, unmodified): PMMH on the SV model with the
// likelihood evaluated by svol_bs filters fanned out through the reference's thread_pool -- the CPU path end to end
template <size_t N>
int example_pmmh(const char* data_file, const char* tmp_dir, const double* start_trans, int iters, int num_pfilters, int t0, int t1,
                 double c0_diag, int num_threads, double* samples, double* loglikes)
{
    using est_t = univ_svol_estimator<3, 1, 1, N, double>;
    Eigen::Matrix<double, 3, 1> st;
    st << start_trans[0], start_trans[1], start_trans[2];
    Eigen::Matrix<double, 3, 3> C0 = Eigen::Matrix<double, 3, 3>::Identity() * c0_diag;
    const std::string dir(tmp_dir);
    est_t m(st, {"null", "twice_fisher", "log"}, (unsigned)iters, (unsigned)num_pfilters, data_file, dir + "/ref_samples", dir + "/ref_messages",
            num_threads != 1, (unsigned)t0, (unsigned)t1, C0, false, 1u, (unsigned)num_threads);
    for (int it = 0; it < iters; ++it) {
        m.m_num_mcmc_iters = (unsigned)(it + 1);
        m.commence_sampling();
        const auto p = m.m_current_theta.get_untrans_params();
        for (int k = 0; k < 3; ++k) samples[it * 3 + k] = p(k);
        if (loglikes) loglikes[it] = m.m_old_log_like;
    }
    return 0;
}

}  // namespace

extern "C" {

// deterministic seeding of every pf stand-in sampler constructed afterwards (0 = clock seeds, as pf does)
void ssme_refhdr_set_seed(uint64_t s)
{
    if (s) pf::shim::set_base_seed(s);
    else pf::shim::use_clock_seeds();
}

// univ_svol_estimator<3,1,1,N,double>::commence_sampling with its own (mt19937) randomness; samples[iters][3] untransformed
// (beta, phi, sigma^2), loglikes[iters] the current log-likelihood estimate.  N in {100, 500}.
int ssme_refhdr_example_pmmh(int N, const char* data_file, const char* tmp_dir, const double* start_trans, int iters, int num_pfilters,
                             int t0, int t1, double c0_diag, int num_threads, double* samples, double* loglikes)
{
    try {
        switch (N) {
        case 100: return example_pmmh<100>(data_file, tmp_dir, start_trans, iters, num_pfilters, t0, t1, c0_diag, num_threads, samples, loglikes);
        case 500: return example_pmmh<500>(data_file, tmp_dir, start_trans, iters, num_pfilters, t0, t1, c0_diag, num_threads, samples, loglikes);
        default: throw std::invalid_argument("unsupported N");
        }
    } catch (const std::exception& e) {
        return fail(e);
    }
}

const char* ssme_refhdr_last_error(void) { return g_err.c_str(); }

// sizes the templates are instantiated for (the reference's particle counts are template parameters)
int ssme_refhdr_supported_sizes(int32_t* out, int cap)
{
    static const int32_t s[] = {16, 100, 500};
    const int n = (int)(sizeof(s) / sizeof(s[0]));
    for (int i = 0; i < n && i < cap; ++i) out[i] = s[i];
    return n;
}

// param::transform<double> (parameters.h:262-449): op 0 trans, 1 inv_trans, 2 log_jacobian; type 0 null, 1 twice_fisher, 2 logit, 3 log
int ssme_refhdr_transform(int type, int op, double x, double* out)
{
    try {
        if (type < 0 || type > 3) throw std::invalid_argument("transform type");
        auto t = param::transform<double>::create(std::string(kTransNames[type]));
        *out = (op == 0) ? t->trans(x) : (op == 1) ? t->inv_trans(x) : t->log_jacobian(x);
        return 0;
    } catch (const std::exception& e) {
        return fail(e);
    }
}

// param::pack<double,4> (parameters.h:462-631): built from transformed (from_transformed = 1) or untransformed values
int ssme_refhdr_pack4(const int* types, const double* vals, int from_transformed, double* trans_out, double* untrans_out, double* logjac_out)
{
    try {
        Eigen::Matrix<double, 4, 1> v;
        v << vals[0], vals[1], vals[2], vals[3];
        std::vector<std::string> names;
        for (int k = 0; k < 4; ++k) names.push_back(kTransNames[types[k]]);
        param::pack<double, 4> pp(v, names, from_transformed != 0);
        param::pack<double, 4> copy(pp), assigned;
        assigned = copy;  // deep copies (parameters.h:488-500, 547-561)
        const auto tp = assigned.get_trans_params();
        const auto up = assigned.get_untrans_params();
        for (int k = 0; k < 4; ++k) {
            trans_out[k] = tp(k);
            untrans_out[k] = up(k);
            if (assigned.get_untrans_params(k, k)(0) != up(k)) throw std::runtime_error("sub-range getter disagrees");
        }
        *logjac_out = assigned.get_log_jacobian();
        return 0;
    } catch (const std::exception& e) {
        return fail(e);
    }
}

// mn_resamp_states_and_params::resampLogWts (liu_west_filter.h:91-145) on N log-weights; anc[N], uniforms[N+1] (may be NULL)
int ssme_refhdr_resample_sorted(int N, const double* lw, uint32_t seed, int32_t* anc, double* uniforms)
{
    try {
        switch (N) {
        case 16: return resample_sorted<16>(lw, seed, anc, uniforms);
        case 100: return resample_sorted<100>(lw, seed, anc, uniforms);
        case 500: return resample_sorted<500>(lw, seed, anc, uniforms);
        default: throw std::invalid_argument("unsupported N");
        }
    } catch (const std::exception& e) {
        return fail(e);
    }
}

// LWFilter2::filter driven over y[0..T) on the SV model with fixed theta = (beta, phi, sigma^2) and delta (1 = no jitter).
// z[T][N] normals; seeds[T] reseed the in-tree resampler before each step.  Out: cond_like[T], x_post[T][N], u_out[T][N+1].
int ssme_refhdr_lwfilter2_sv(int N, const double* theta, const double* y, int T, int rs, double delta, const double* z,
                             const uint32_t* seeds, double* cond_like, double* x_post, double* u_out)
{
    try {
        switch (N) {
        case 16: return lwfilter2_sv<16>(theta, y, T, rs, delta, z, seeds, cond_like, x_post, u_out);
        case 100: return lwfilter2_sv<100>(theta, y, T, rs, delta, z, seeds, cond_like, x_post, u_out);
        case 500: return lwfilter2_sv<500>(theta, y, T, rs, delta, z, seeds, cond_like, x_post, u_out);
        default: throw std::invalid_argument("unsupported N");
        }
    } catch (const std::exception& e) {
        return fail(e);
    }
}

// the example's svol_bs (unmodified) on the pf stand-in's BSFilter + mn_resampler: z[T][N] normals, u[T][N] uniforms
int ssme_refhdr_bsfilter_sv(int N, const double* theta, const double* y, int T, const double* z, const double* u, double* cond_like,
                            double* x_post)
{
    try {
        switch (N) {
        case 16: return bsfilter_sv<16>(theta, y, T, z, u, cond_like, x_post);
        case 100: return bsfilter_sv<100>(theta, y, T, z, u, cond_like, x_post);
        case 500: return bsfilter_sv<500>(theta, y, T, z, u, cond_like, x_post);
        default: throw std::invalid_argument("unsupported N");
        }
    } catch (const std::exception& e) {
        return fail(e);
    }
}

// form 0: LWFilter2WithCovs::filter on svol_lw_2_par (test/test_liu_west.cpp:213-358), N in {16,100,500};
// form 1: LWFilterWithCovs::filter on svol_lw_1_par (:26-157), whose particle count is the test file's NPARTS = 10.
int ssme_refhdr_lw_leverage(int form, int N, const double* lo, const double* hi, double delta, const double* y, const double* cov, int T,
                            const double* u_prior, const double* z_state, const double* z_jitter, const double* u_aux,
                            const uint32_t* seeds, double* cond_like, double* theta_bar, double* x_post, double* th_post, double* expect,
                            double* u_resamp)
{
    try {
        const lw_out o{cond_like, theta_bar, x_post, th_post, expect, u_resamp};
        if (form == 0) {
            switch (N) {
            case 16: return lw_leverage_form0<16>(lo, hi, delta, y, cov, T, u_prior, z_state, z_jitter, seeds, o);
            case 100: return lw_leverage_form0<100>(lo, hi, delta, y, cov, T, u_prior, z_state, z_jitter, seeds, o);
            case 500: return lw_leverage_form0<500>(lo, hi, delta, y, cov, T, u_prior, z_state, z_jitter, seeds, o);
            default: throw std::invalid_argument("unsupported N");
            }
        }
        if (form == 1) {
            if (N != NPARTS) throw std::invalid_argument("svol_lw_1_par is compiled for the test file's NPARTS only");
            svol_lw_1_par<NPARTS, double> f(delta, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], lo[3], hi[3], 1);
            return lw_run<svol_lw_1_par<NPARTS, double>, NPARTS>(f, y, cov, T, u_prior, z_state, z_jitter, u_aux, seeds, o);
        }
        throw std::invalid_argument("form");
    } catch (const std::exception& e) {
        return fail(e);
    }
}

// form 0 with a resampling schedule rs >= 1 (LWFilter2WithCovs(transforms, delta, rs), liu_west_filter.h:2047)
int ssme_refhdr_lw_leverage_rs(int N, int rs, const double* lo, const double* hi, double delta, const double* y, const double* cov, int T,
                               const double* u_prior, const double* z_state, const double* z_jitter, const uint32_t* seeds,
                               double* cond_like, double* theta_bar, double* x_post, double* th_post, double* expect, double* u_resamp)
{
    try {
        if (rs < 1) throw std::invalid_argument("rs");
        const lw_out o{cond_like, theta_bar, x_post, th_post, expect, u_resamp};
        switch (N) {
        case 16: return lw_leverage_form0<16>(lo, hi, delta, y, cov, T, u_prior, z_state, z_jitter, seeds, o, (unsigned)rs);
        case 100: return lw_leverage_form0<100>(lo, hi, delta, y, cov, T, u_prior, z_state, z_jitter, seeds, o, (unsigned)rs);
        case 500: return lw_leverage_form0<500>(lo, hi, delta, y, cov, T, u_prior, z_state, z_jitter, seeds, o, (unsigned)rs);
        default: throw std::invalid_argument("unsupported N");
        }
    } catch (const std::exception& e) {
        return fail(e);
    }
}

// The reference's ada_pmmh_mvn::commence_sampling on a closed-form likelihood (toy_pmmh above).  z_prop[iters][3] are the
// normals of q_samp, u_acc[iters] the accept uniforms (row 0 of both unused: iteration 0 does not propose).
// Out: samples[iters][3] (untransformed, as written to the samples file), accept[iters], ct[9] = get_ct() at the end.
int ssme_refhdr_pmmh_chain(const double* start_trans, const double* data, int ndata, int iters, int t0, int t1, const double* c0,
                           const double* z_prop, const double* u_acc, const char* tmp_dir, double* samples, int32_t* accept, double* ct)
{
    try {
        const std::string dir(tmp_dir);
        {
            std::ofstream d(dir + "/data.csv");
            d.precision(17);
            for (int i = 0; i < ndata; ++i) d << data[i] << "\n";
        }
        Eigen::Matrix<double, 3, 1> st;
        st << start_trans[0], start_trans[1], start_trans[2];
        Eigen::Matrix<double, 3, 3> C0;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) C0(i, j) = c0[i * 3 + j];
        toy_pmmh m(st, {"null", "twice_fisher", "log"}, (unsigned)iters, 3u, dir + "/data.csv", dir + "/samples", dir + "/messages", false,
                   (unsigned)t0, (unsigned)t1, C0, false, 1u, 1u);
        pf::shim::mvn_stream().arm(z_prop + 3, (size_t)(iters - 1) * 3);
        pf::shim::uniform_stream().arm(u_acc + 1, (size_t)(iters - 1));
        // one iteration at a time so the accept flags can be read back (m_num_mcmc_iters bounds the loop: ada_pmmh_mvn.h:332)
        for (int it = 0; it < iters; ++it) {
            m.m_num_mcmc_iters = (unsigned)(it + 1);
            m.commence_sampling();
            const auto p = m.m_current_theta.get_untrans_params();
            for (int k = 0; k < 3; ++k) samples[it * 3 + k] = p(k);
            accept[it] = (it > 0 && m.m_accepted) ? 1 : 0;
        }
        pf::shim::mvn_stream().disarm();
        pf::shim::uniform_stream().disarm();
        const auto C = m.get_ct();
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) ct[i * 3 + j] = C(i, j);
        return 0;
    } catch (const std::exception& e) {
        return fail(e);
    }
}

}  // extern "C"
