// tests/cpp/test_gpu_host.cpp -- GPU tests of the C++ host layer (gpu_pool, ada_pmmh_mvn) through the C ABI.
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <ssme_b200/ada_pmmh_mvn.hpp>
#include <ssme_b200/gpu_pool.hpp>
#include <ssme_b200/liu_west_filter.hpp>
#include <ssme_b200/pswarm_filter.hpp>
#include <ssme_b200/rv_eval.hpp>

#include "../../examples/estimate_univ_svol.hpp"
#include "check.hpp"

static std::vector<double> sv_series(size_t T, unsigned seed)
{
    std::mt19937 g(seed);
    std::normal_distribution<double> n(0, 1);
    std::vector<double> y(T);
    double x = n(g) * 0.25 / std::sqrt(1 - 0.95 * 0.95);
    for (size_t t = 0; t < T; ++t) {
        if (t) x = 0.95 * x + 0.25 * n(g);
        y[t] = std::exp(0.5 * x) * n(g);
    }
    return y;
}

int main(int argc, char** argv)
{
    const std::string tmpdir = argc > 1 ? argv[1] : "/tmp";
    using pool_t = ssme_b200::gpu_pool<3, 1, double>;
    using pack3 = param::pack<double, 3>;
    const std::vector<std::string> tts{"null", "twice_fisher", "log"};
    const auto y = sv_series(300, 5);
    std::vector<ssme_b200::vec<double, 1>> data(y.size());
    for (size_t t = 0; t < y.size(); ++t) data[t](0) = y[t];

    TEST_CASE("test thread pool [thread_pool] -- log-mean-exp known answer")  // test_thread_pool.cpp:36-47
    {
        std::vector<double> vals(10000, 3.0);  // NUMCALCS 1e4 evaluations of d({1,1,1}) = 3
        double out = 0;
        for (int i = 0; i < 100; ++i) {
            ssme_b200::throw_on_error(ssme_b200_log_mean_exp(0, vals.data(), 1, 10000, &out));
            REQUIRE(std::abs(out - 3.0) < .001);
        }
    }
    TEST_CASE("gpu_pool: thread_pool error conventions")
    {
        pool_t pool(4, 500);
        pack3 theta(ssme_b200::vec<double, 3>{1.0, ssme_b200::rveval::twiceFisher(.95), std::log(0.0625)}, tts);
        REQUIRE_THROWS_AS(pool.work(theta), std::runtime_error);  // thread_pool.h:192
        REQUIRE_THROWS_AS(pool.add_observed_data({}), std::length_error);  // estimate_univ_svol.h:112-113
        pool.add_observed_data(data);
        REQUIRE_THROWS_AS(pool.add_observed_data(data), std::runtime_error);  // thread_pool.h:169
        const double a = pool.work(theta), b = pool.work(theta);
        REQUIRE(std::isfinite(a) && std::isfinite(b));
        REQUIRE(a != b);                  // fresh random streams on every call, like the clock-seeded reference
        REQUIRE(std::abs(a - b) < 5.0);   // but the same likelihood
        pool.set_next_stream(0);
        const double a2 = pool.work(theta);
        REQUIRE(a2 == a);                 // and reproducible when the stream ids are pinned
        std::vector<pack3> batch{theta, theta};
        pool.set_next_stream(0);
        auto v = pool.work_batch(batch);
        REQUIRE(v.size() == 2 && v[0] == a && v[1] == b);
    }
    TEST_CASE("ada_pmmh_mvn on the GPU backend: file formats and chain behaviour")
    {
        const std::string f = tmpdir + "/ssme_b200_pmmh_data.csv";
        { std::ofstream o(f); o.precision(17); for (double v : y) o << v << "\n"; }
        using est = univ_svol_estimator<3, 1, 1, 200, double>;
        est::psv start{1.0, ssme_b200::rveval::twiceFisher(.9), std::log(0.05)};
        est::psm C0 = est::psm::Identity() * .01;
        const unsigned iters = 60;
        std::string samples_file, messages_file;
        {
            est m(start, tts, iters, 2, f, tmpdir + "/ssme_b200_samples", tmpdir + "/ssme_b200_messages", false, 20, 1000, C0, false, 1, 0,
                  ssme_b200::gpu_options(), 12345);
            m.commence_sampling();
            REQUIRE(m.iterations_done() == iters);
            REQUIRE(m.accept_rate() > 0.0 && m.accept_rate() < 1.0);
            REQUIRE(std::isfinite(m.current_log_like()));
            // adaptation happened inside the window (t0 = 20): C_t is no longer the diagonal C0
            REQUIRE(std::abs(m.get_ct()(0, 1)) > 0.0);
            REQUIRE(std::isfinite(m.log_like_eval(m.current_theta(), data)));
        }
        // locate the two time-stamped files
        FILE* p = popen(("ls " + tmpdir + "/ssme_b200_samples_* " + tmpdir + "/ssme_b200_messages_* 2>/dev/null").c_str(), "r");
        char line[512];
        while (p && fgets(line, sizeof(line), p)) {
            std::string s(line);
            s.erase(s.find_last_not_of("\n") + 1);
            if (s.find("samples") != std::string::npos) samples_file = s; else messages_file = s;
        }
        if (p) pclose(p);
        std::ifstream sf(samples_file), mf(messages_file);
        std::string l;
        unsigned nrows = 0;
        while (std::getline(sf, l)) {
            ++nrows;
            REQUIRE(std::count(l.begin(), l.end(), ',') == 2);  // numparams - 1 commas, no header
        }
        REQUIRE(nrows == iters);
        std::getline(mf, l);
        REQUIRE(l == "iter number, accept rate, old_ll, new_ll, old_lprior, new_lprior, accept prob, outcome");  // ada_pmmh_mvn.h:308
        unsigned mrows = 0;
        while (std::getline(mf, l)) { ++mrows; REQUIRE(std::count(l.begin(), l.end(), ',') == 7); }
        REQUIRE(mrows == iters);
        std::remove(samples_file.c_str());
        std::remove(messages_file.c_str());
    }
    TEST_CASE("test filter without funcs for type 2 filters with covariates [filter method]")  // test_liu_west.cpp:365-375
    {
        // svol_lw_2_par<NPARTS,FLOATTYPE> mod(.99, .8, .99, -.1, .1, .01, .1, -.5, -.01, 10); one filter() call; logCondLike^2 > 0
        using lw_t = ssme_b200::LWFilter2WithCovs_svol<10, double>;
        lw_t mod({"logit", "null", "log", "twice_fisher"}, .99, lw_t::psv{.8, -.1, .01, -.5}, lw_t::psv{.99, .1, .1, -.01});
        mod.filter_series({0.3}, {0.0});
        REQUIRE(std::pow(mod.getLogCondLike(), 2) > 0.0);
        REQUIRE_THROWS_AS((lw_t({"null", "null", "log", "twice_fisher"}, .99, lw_t::psv{.8, -.1, .01, -.5}, lw_t::psv{.99, .1, .1, -.01})),
                          std::invalid_argument);
        // a longer run learns something: posterior means stay inside the prior box and cond-likes are finite
        ssme_b200::LWFilter2WithCovs_svol<20000, double> big({"logit", "null", "log", "twice_fisher"}, .99, lw_t::psv{.8, -.1, .01, -.5},
                                                             lw_t::psv{.99, .1, .3, -.01});
        std::vector<double> cov(y.size(), 0.0);
        for (size_t t = 1; t < y.size(); ++t) cov[t] = y[t - 1];
        big.filter_series(y, cov);
        const auto pm = big.getParamMeans();
        REQUIRE(pm(0) > .8 && pm(0) < .99 && pm(2) > .01 && pm(2) < .3 && pm(3) > -.5 && pm(3) < -.01);
        REQUIRE(std::isfinite(big.getLogLike()));
        // the object filters a second series (ssme_b200_replace_observations): same stream id, same series -> the same bits
        const double first = big.getLogLike();
        big.filter_series(y, cov);
        REQUIRE(big.getLogLike() == first);
        std::vector<double> y2(y.begin(), y.begin() + y.size() / 2), cov2(cov.begin(), cov.begin() + y.size() / 2);
        big.filter_series(y2, cov2);
        REQUIRE((std::isfinite(big.getLogLike()) && big.getLogLike() != first));
    }
    TEST_CASE("test filter without funcs for type 1 filters with covariates [filter method]")  // test_liu_west.cpp:163-173
    {
        // svol_lw_1_par<NPARTS,FLOATTYPE> mod(.99, .8, .99, -.1, .1, .01, .1, -.5, -.01, 10); filter(y1, z1); logCondLike^2 > 0
        using lw_t = ssme_b200::LWFilterWithCovs_svol<10, double>;
        lw_t mod({"logit", "null", "log", "twice_fisher"}, .99, lw_t::psv{.8, -.1, .01, -.5}, lw_t::psv{.99, .1, .1, -.01});
        mod.filter_series({0.3, -0.2, 0.5}, {0.0, 0.3, -0.2});
        for (size_t t = 0; t < 3; ++t) REQUIRE(std::pow(mod.getLogCondLike(t), 2) > 0.0);
        ssme_b200::LWFilterWithCovs_svol<20000, double> big({"logit", "null", "log", "twice_fisher"}, .99, lw_t::psv{.8, -.1, .01, -.5},
                                                            lw_t::psv{.99, .1, .3, -.01});
        ssme_b200::LWFilter2WithCovs_svol<20000, double> big2({"logit", "null", "log", "twice_fisher"}, .99, lw_t::psv{.8, -.1, .01, -.5},
                                                              lw_t::psv{.99, .1, .3, -.01});
        std::vector<double> cov(y.size(), 0.0);
        for (size_t t = 1; t < y.size(); ++t) cov[t] = y[t - 1];
        big.filter_series(y, cov);
        big2.filter_series(y, cov);
        const auto pm = big.getParamMeans();
        REQUIRE(pm(0) > .8 && pm(0) < .99 && pm(2) > .01 && pm(2) < .3 && pm(3) > -.5 && pm(3) < -.01);
        REQUIRE(std::isfinite(big.getLogLike()));
        REQUIRE(big.getLogLike() != big2.getLogLike());  // two different estimators of the same quantity
        // the reference's streaming call, one observation at a time, gives the same numbers as the whole-series call
        ssme_b200::LWFilterWithCovs_svol<20000, double> online({"logit", "null", "log", "twice_fisher"}, .99, lw_t::psv{.8, -.1, .01, -.5},
                                                               lw_t::psv{.99, .1, .3, -.01});
        for (size_t t = 0; t < y.size(); ++t) online.filter(y[t], cov[t]);
        REQUIRE(online.getLogLike() == big.getLogLike());
        REQUIRE(online.getLogCondLike() == big.getLogCondLike());
        REQUIRE(online.getParamMeans()(0) == big.getParamMeans()(0));
        // LWFilterWithCovsFutureSimulator::sim_future_obs(num_steps, last_obs) (liu_west_filter.h:1315-1360): [time][particle]
        const auto sim = online.sim_future_obs(5, y.back(), 77);
        REQUIRE(sim.size() == 5);
        REQUIRE(sim[0].size() == 20000);
        double m1 = 0, m2 = 0;
        for (const auto& row : sim)
            for (double v : row) { REQUIRE(std::isfinite(v)); m1 += v; m2 += v * v; }
        m1 /= 1e5; m2 /= 1e5;
        REQUIRE(std::abs(m1) < 0.05);
        REQUIRE(m2 > 0.05);
        REQUIRE(online.sim_future_obs(5, y.back(), 77) == sim);   // same stream, same simulation; the filter is untouched:
        online.filter(y[0], cov[0]);
        REQUIRE(std::isfinite(online.getLogCondLike()));
        REQUIRE(std::abs(big.getLogLike() - big2.getLogLike()) < 0.02 * std::abs(big2.getLogLike()));
    }
    TEST_CASE("Liu-West SISR filter with a resampling schedule (the reference's constructor argument rs)")
    {
        using lw_t = ssme_b200::LWFilter2WithCovs_svol<5000, double>;
        const std::vector<std::string> tr{"logit", "null", "log", "twice_fisher"};
        const lw_t::psv lo{.8, -.1, .01, -.5}, hi{.99, .1, .3, -.01};
        std::vector<double> y{0.3, -0.2, 0.5, 0.1, -0.4, 0.2, 0.05, -0.6}, cov(y.size(), 0.0);
        for (size_t t = 1; t < y.size(); ++t) cov[t] = y[t - 1];
        lw_t every(tr, .99, lo, hi, 1), second(tr, .99, lo, hi, 2);   // LWFilter2WithCovs(transforms, delta, rs), liu_west_filter.h:2047
        every.filter_series(y, cov, 3);
        second.filter_series(y, cov, 3);
        REQUIRE(std::isfinite(second.getLogLike()));
        REQUIRE(second.getLogLike() != every.getLogLike());
        REQUIRE(second.getLogCondLike(0) == every.getLogCondLike(0));   // step 0 is the same draw in both
        REQUIRE(std::abs(second.getLogLike() - every.getLogLike()) < 0.05 * std::abs(every.getLogLike()));
        lw_t online(tr, .99, lo, hi, 2);
        online.set_stream(3);
        for (size_t t = 0; t < y.size(); ++t) online.filter(y[t], cov[t]);
        REQUIRE(online.getLogLike() == second.getLogLike());
        bool threw = false;
        try {
            ssme_b200::LWFilterWithCovs_svol<5000, double> apf(tr, .99, lo, hi, 2);   // the auxiliary form has no schedule
        } catch (const std::invalid_argument&) {
            threw = true;
        }
        REQUIRE(threw);
    }
    TEST_CASE("covariate-free Liu-West twins: filter(y_t)")
    {
        using lw_t = ssme_b200::LWFilter2_svol<5000, double>;
        lw_t a({"logit", "null", "log", "twice_fisher"}, .99, lw_t::psv{.8, -.1, .01, -.5}, lw_t::psv{.99, .1, .3, -.01});
        ssme_b200::LWFilter_svol<5000, double> b({"logit", "null", "log", "twice_fisher"}, .99, lw_t::psv{.8, -.1, .01, -.5},
                                                 lw_t::psv{.99, .1, .3, -.01});
        std::vector<double> head(y.begin(), y.begin() + 30);
        a.filter_series(head);
        for (double v : head) b.filter(v);
        REQUIRE(std::isfinite(a.getLogLike()));
        REQUIRE(std::isfinite(b.getLogLike()));
        REQUIRE(std::pow(a.getLogCondLike(), 2) > 0.0);
        REQUIRE(std::abs(a.getLogLike() - b.getLogLike()) < 0.05 * std::abs(a.getLogLike()));
    }
    TEST_CASE("swarm: 10 x 10 particles, assertions of test_pswarm.cpp:251-252")
    {
        struct my_swarm : ssme_b200::Swarm<10, 10, 4, double> {
            std::mt19937 g{3};
            my_swarm() : ssme_b200::Swarm<10, 10, 4, double>([] { ssme_b200::gpu_options o; o.model = SSME_B200_MODEL_SV_LEVERAGE; return o; }()) {}
            psv samp_untrans_params() override
            {  // the uniform priors of svol_swarm_1 (test_pswarm.cpp:244): phi, mu, sigma, rho
                std::uniform_real_distribution<double> a(.8, .99), b(-.1, .1), c(.01, .1), d(-.5, -.01);
                return psv{a(g), b(g), c(g), d(g)};
            }
        } sw;
        std::vector<double> rows(2 * 40);
        for (size_t t = 0; t < 40; ++t) { rows[2 * t] = y[t]; rows[2 * t + 1] = t ? y[t - 1] : 0.0; }
        sw.update_series(rows, 2);
        REQUIRE(sw.num_obs() == 40);
        for (size_t t = 0; t < 40; ++t) REQUIRE(std::pow(sw.getLogCondLike(t), 2) > 0.0);
        // "test update with funcs" (test_pswarm.cpp:323-346): expectations come back finite and consistent
        my_swarm sw3;  // the reference's streaming call, one observation at a time
        for (size_t t = 0; t < 40; ++t) sw3.update({rows[2 * t], rows[2 * t + 1]});
        REQUIRE(sw3.num_obs() == 40);
        for (size_t t = 0; t < 40; ++t) REQUIRE(sw3.getLogCondLike(t) == sw.getLogCondLike(t));
        my_swarm sw2;  // same parameter draws (same generator seed), same random streams
        sw2.update_series(rows, 2, 0, true);
        for (size_t t = 0; t < 40; ++t) {
            REQUIRE(sw2.getLogCondLike(t) == sw.getLogCondLike(t));
            REQUIRE(std::isfinite(sw2.getExpectation(t, 0)));
            REQUIRE(sw2.getExpectation(t, 1) >= 0.0);
        }
    }
    return finish();
}
