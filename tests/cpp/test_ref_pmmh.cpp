// tests/cpp/test_ref_pmmh.cpp -- the reference's OWN ada_pmmh_mvn::commence_sampling (include/ssme/ada_pmmh_mvn.h:325-372,
// compiled unmodified into oracle/_ref/libssme_refhdr.so, see oracle/ref_harness.cpp) against this repository's
// multi-chain driver (include/ssme_b200/pmmh_multichain.hpp) on a closed-form likelihood and identical proposal /
// accept streams: same accept decisions, same chain, same adapted proposal covariance.
#include <cstdint>
#include <random>
#include <string>
#include <vector>

#include <ssme_b200/pmmh_multichain.hpp>
#include <ssme_b200/rv_eval.hpp>

#include "check.hpp"

extern "C" int ssme_refhdr_pmmh_chain(const double* start_trans, const double* data, int ndata, int iters, int t0, int t1, const double* c0,
                                      const double* z_prop, const double* u_acc, const char* tmp_dir, double* samples, int32_t* accept,
                                      double* ct);
extern "C" const char* ssme_refhdr_last_error(void);
extern "C" int ssme_refhdr_pack4(const int* types, const double* vals, int from_transformed, double* trans_out, double* untrans_out,
                                 double* logjac_out);

int main(int argc, char** argv)
{
    const std::string tmpdir = argc > 1 ? argv[1] : "/tmp";
    using driver = ssme_b200::pmmh_multichain<3, double>;
    namespace rv = ssme_b200::rveval;
    const int iters = 400, t0 = 20, t1 = 300;  // adaptation window inside the run (ada_pmmh_mvn.h:247)
    const unsigned long seed = 4242;
    std::vector<double> data{0.8, 1.4, 0.9, 1.2, 1.1};
    const std::vector<std::string> tts{"null", "twice_fisher", "log"};
    driver::psv start;
    start(0) = 1.0;
    start(1) = rv::twiceFisher<double>(.5);
    start(2) = std::log(0.3);
    const driver::psm C0 = driver::psm::Identity() * .15;

    // the draws chain 0 of pmmh_multichain will consume, replayed from its generator (pmmh_multichain.hpp: q_samp draws
    // numparams normals from a fresh normal_distribution, then the accept step one uniform, per iteration >= 1)
    std::vector<double> z_prop((size_t)iters * 3, 0.0), u_acc((size_t)iters, 0.0);
    {
        std::mt19937 gen(static_cast<std::uint32_t>(seed));
        for (int it = 1; it < iters; ++it) {
            std::normal_distribution<double> rnorm(0.0, 1.0);
            for (int k = 0; k < 3; ++k) z_prop[(size_t)it * 3 + k] = rnorm(gen);
            std::uniform_real_distribution<double> runif(0.0, 1.0);
            u_acc[(size_t)it] = runif(gen);
        }
    }

    TEST_CASE("reference commence_sampling == pmmh_multichain on identical streams");
    std::vector<double> ref_samples((size_t)iters * 3), ref_ct(9);
    std::vector<int32_t> ref_accept((size_t)iters);
    std::vector<double> c0(9, 0.0);
    c0[0] = c0[4] = c0[8] = .15;
    double st[3] = {start(0), start(1), start(2)};
    const int rc = ssme_refhdr_pmmh_chain(st, data.data(), (int)data.size(), iters, t0, t1, c0.data(), z_prop.data(), u_acc.data(),
                                          tmpdir.c_str(), ref_samples.data(), ref_accept.data(), ref_ct.data());
    if (rc != 0) std::printf("refhdr error: %s\n", ssme_refhdr_last_error());
    REQUIRE(rc == 0);

    auto prior = [](const param::pack<double, 3>& theta) {
        const auto p = theta.get_untrans_params();
        return rv::evalUnivNorm<double>(p(0), 1.0, 1.0, true) + rv::evalUniform<double>(p(1), 0.0, 1.0, true) +
               rv::evalUnivInvGamma<double>(p(2), .001, .001, true);
    };
    auto evaluator = [&data](const double* theta, size_t C, unsigned R, std::uint64_t, double* per_filter) {
        for (size_t c = 0; c < C; ++c) {
            const double* p = theta + c * 3;
            double s = 0.0;
            for (size_t i = 0; i < data.size(); ++i) s += -0.5 * (data[i] - p[0]) * (data[i] - p[0]);
            const double l2 = std::log(p[2]) + 1.0;
            const double v = s - 2.0 * (p[1] - 0.3) * (p[1] - 0.3) - 0.5 * l2 * l2;
            for (unsigned r = 0; r < R; ++r) per_filter[c * R + r] = v;
        }
    };
    driver d({start}, tts, 3u, (unsigned)t0, (unsigned)t1, C0, prior, evaluator, seed);
    int accepted = 0, flag_mismatch = 0;
    double worst = 0.0;
    for (int it = 0; it < iters; ++it) {
        d.step();
        const auto p = d.chain(0).current_theta.get_untrans_params();
        for (int k = 0; k < 3; ++k) {
            const double e = std::fabs(p(k) - ref_samples[(size_t)it * 3 + k]) / std::fabs(ref_samples[(size_t)it * 3 + k]);
            if (e > worst) worst = e;
        }
        const int acc = (it > 0 && d.chain(0).accepted) ? 1 : 0;
        accepted += acc;
        if (acc != ref_accept[(size_t)it]) ++flag_mismatch;
    }
    std::printf("accepted %d of %d proposals; worst relative difference of the chains %.3g\n", accepted, iters - 1, worst);
    REQUIRE(flag_mismatch == 0);
    REQUIRE(accepted > 20 && accepted < iters - 20);  // both branches exercised
    REQUIRE(worst < 1e-12);
    double worst_ct = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const double e = std::fabs(d.chain(0).Ct(i, j) - ref_ct[i * 3 + j]);
            if (e > worst_ct) worst_ct = e;
        }
    REQUIRE(worst_ct < 1e-12);  // get_ct() after the adaptation window (ada_pmmh_mvn.h:76,247-248)

    TEST_CASE("param::pack of parameters.hpp == the reference's param::pack (parameters.h:462-631) on random packs");
    {
        const char* names[4] = {"null", "twice_fisher", "logit", "log"};
        std::mt19937 g(99);
        std::normal_distribution<double> nrm(0.0, 2.0);
        int bad = 0;
        for (int trial = 0; trial < 200; ++trial) {
            int types[4];
            double vals[4], tp[4], up[4], lj = 0.0;
            std::vector<std::string> ts;
            ssme_b200::vec<double, 4> v;
            for (int k = 0; k < 4; ++k) {
                types[k] = (int)(g() % 4);
                vals[k] = nrm(g);
                v(k) = vals[k];
                ts.push_back(names[types[k]]);
            }
            if (ssme_refhdr_pack4(types, vals, 1, tp, up, &lj) != 0) { ++bad; continue; }
            param::pack<double, 4> pp(v, ts, true);
            const auto mine_u = pp.get_untrans_params();
            for (int k = 0; k < 4; ++k)
                if (mine_u(k) != up[k] || pp.get_trans_params()(k) != tp[k]) ++bad;
            if (pp.get_log_jacobian() != lj) ++bad;
            // and from the untransformed side (from_transformed = false, parameters.h:475-480)
            if (ssme_refhdr_pack4(types, up, 0, tp, up, &lj) == 0) {
                ssme_b200::vec<double, 4> u2;
                for (int k = 0; k < 4; ++k) u2(k) = mine_u(k);
                param::pack<double, 4> p2(u2, ts, false);
                for (int k = 0; k < 4; ++k)
                    if (p2.get_trans_params()(k) != tp[k]) ++bad;
            }
        }
        REQUIRE(bad == 0);
    }
    return finish();
}
