// tests/cpp/test_host.cpp -- CPU tests of the host-side mirror of the reference interface.
// The known answers are the reference's own (test/test_parameters.cpp, test/test_utils.cpp); the
// test bodies read like the reference's tests on purpose.
#include <fstream>
#include <string>
#include <vector>

#include <ssme_b200/fixed.hpp>
#include <ssme_b200/parameters.hpp>
#include <ssme_b200/rv_eval.hpp>
#include <ssme_b200/utils.hpp>
#include <ssme_b200/gpu_pool.hpp>

#include "check.hpp"

using vec4 = ssme_b200::vec<double, 4>;

int main(int argc, char** argv)
{
    const std::string tmpdir = argc > 1 ? argv[1] : "/tmp";
    const std::vector<std::string> ts{"null", "log", "logit", "twice_fisher"};
    const vec4 trans_params{1.0, -1.3, 9.5, .89};
    const vec4 ideal_un_trans_params{1.0, 0.2725318, 0.9999252, 0.4177803};  // test_parameters.cpp:114

    TEST_CASE("test constructors [pack]")  // test_parameters.cpp:59-91 (the parts the reference left commented out, enabled)
    {
        param::pack<double, 4> p1;
        REQUIRE(p1.size() == 0);
        REQUIRE(p1.capacity() == 4);
        REQUIRE_THROWS_AS(p1.get_trans_params(), std::length_error);
        REQUIRE_THROWS_AS(p1.get_untrans_params(), std::length_error);
        REQUIRE_THROWS_AS(p1.get_log_jacobian(), std::length_error);
        param::pack<double, 4> p2(trans_params, ts, true);
        REQUIRE(p2.size() == 4);
        param::pack<double, 4> p3(p2);
        REQUIRE(p3.size() == 4);
        REQUIRE_THROWS_AS((param::pack<double, 4>(p1)), std::invalid_argument);  // parameters.h:498
        REQUIRE_THROWS_AS((param::pack<double, 4>(trans_params, std::vector<std::string>{"null"})), std::invalid_argument);
        REQUIRE_THROWS_AS((param::pack<double, 4>(trans_params, std::vector<std::string>{"null", "log", "logit", "bogus"})),
                          std::invalid_argument);
    }
    TEST_CASE("test assignment [pack]")  // test_parameters.cpp:94-109
    {
        param::pack<double, 4> pp1(trans_params, ts);
        param::pack<double, 4> pp2(vec4{1.0, 1.0, 1.0, 1.0}, ts);
        pp1 = pp2;
        for (size_t i = 0; i < 4; ++i) REQUIRE(std::abs(1.0 - pp1.get_trans_params()(i)) < .00001);
        param::pack<double, 4> empty;
        REQUIRE_THROWS_AS(pp1 = empty, std::invalid_argument);
    }
    TEST_CASE("test transformations [pack]")  // test_parameters.cpp:112-121
    {
        param::pack<double, 4> pp(trans_params, ts);
        for (size_t i = 0; i < 4; ++i) REQUIRE(std::abs(ideal_un_trans_params(i) - pp.get_untrans_params()(i)) < .0001);
    }
    TEST_CASE("test transformations pt 2 [pack]")  // test_parameters.cpp:123-137
    {
        param::pack<double, 4> pp(trans_params, ts);
        auto first_three = pp.get_untrans_params(0, 2);
        REQUIRE(first_three.size() == 3);
        for (size_t i = 0; i < 3; ++i) REQUIRE(std::abs(ideal_un_trans_params(i) - first_three[i]) < .0001);
    }
    TEST_CASE("test LogJacobians [pack]")  // test_parameters.cpp:139-146
    {
        param::pack<double, 4> pp(trans_params, ts);
        REQUIRE(std::abs(-11.6851 - pp.get_log_jacobian()) < .0001);
    }
    TEST_CASE("test subsetting [pack]")  // test_parameters.cpp:149-165
    {
        param::pack<double, 4> pp(trans_params, ts);
        for (unsigned i = 0; i < 4; ++i) REQUIRE(std::abs(trans_params(i) - pp.get_trans_params(i, i)[0]) < .0001);
        for (unsigned i = 0; i < 4; ++i) REQUIRE(std::abs(ideal_un_trans_params(i) - pp.get_untrans_params(i, i)[0]) < .0001);
    }
    TEST_CASE("from untransformed, add_param_and_transform, round trips")
    {
        param::pack<double, 4> pp(ideal_un_trans_params, ts, false);
        for (size_t i = 0; i < 4; ++i) REQUIRE(std::abs(pp.get_trans_params()(i) - trans_params(i)) < 2e-3);
        param::pack<double, 2> q;
        q.add_param_and_transform(.5, param::trans_type::TT_logit);  // logit(.5) == 0
        REQUIRE(q.size() == 1);
        REQUIRE_THROWS_AS(q.get_trans_params(), std::length_error);
        q.add_param_and_transform(0.0, "log", true);  // exp(0) == 1
        REQUIRE(std::abs(q.get_trans_params()(0)) < 1e-12);
        REQUIRE(std::abs(q.get_untrans_params()(1) - 1.0) < 1e-12);
        REQUIRE_THROWS_AS(q.add_param_and_transform(1.0, "null"), std::length_error);  // parameters.h:521
        REQUIRE_THROWS_AS((param::transform<double>::trans(param::trans_type::TT_twice_fisher, 1.0)), std::invalid_argument);
        REQUIRE_THROWS_AS((param::transform<double>::trans(param::trans_type::TT_logit, 1.5)), std::invalid_argument);
        REQUIRE_THROWS_AS((param::transform<double>::trans(param::trans_type::TT_log, -1.0)), std::invalid_argument);
        for (double x : {-30.0, -2.0, 0.0, 0.7, 12.0}) {
            for (auto tt : {param::trans_type::TT_null, param::trans_type::TT_twice_fisher, param::trans_type::TT_logit, param::trans_type::TT_log}) {
                const double u = param::transform<double>::inv_trans(tt, x);
                REQUIRE(std::abs(param::transform<double>::trans(tt, u) - x) < 1e-6 * (1 + std::abs(x)) || std::abs(x) > 20);
            }
        }
        // float instantiation, as the example uses (example/main.cpp:13)
        param::pack<float, 3> pf(ssme_b200::vec<float, 3>{1.0f, 1.0986123f, -8.517193f}, {"null", "twice_fisher", "log"});
        REQUIRE(std::abs(pf.get_untrans_params()(1) - 0.5f) < 1e-5f);
        REQUIRE(std::abs(pf.get_untrans_params()(2) - 2.0e-4f) < 1e-8f);
    }
    TEST_CASE("data_reader_test [read_in_data]")  // test_utils.cpp:9-19, fixture test/test_data.csv = "1.23, 4.56"
    {
        const std::string f = tmpdir + "/ssme_b200_test_data.csv";
        { std::ofstream o(f); o << "1.23, 4.56\n"; }
        auto data = utils::read_data<2, double>(f);
        REQUIRE(data.size() == 1);
        REQUIRE(std::abs(1.23 - data[0](0)) < .0001);
        REQUIRE(std::abs(4.56 - data[0](1)) < .0001);
        { std::ofstream o(f); o << "0.5\nnot_a_number\n-0.25\n\n1e-3\n"; }
        auto col = utils::read_data<1, double>(f);  // bad rows are skipped (utils.h:53-56)
        REQUIRE(col.size() == 3);
        REQUIRE(col[1](0) == -0.25 && col[2](0) == 1e-3);
        REQUIRE(utils::read_data<1, double>(tmpdir + "/does_not_exist.csv").empty());
        { std::ofstream o(f); for (int i = 0; i < 33; ++i) o << ".9,0.0,1.0,-.1\n"; }  // test/test_svol_leverage_samples.csv
        utils::csv_param_sampler<4, double> s(f, 7);
        REQUIRE(s.num_rows() == 33);
        auto row = s.samp();
        REQUIRE(row(0) == .9 && row(2) == 1.0 && row(3) == -.1);
    }
    TEST_CASE("rv_eval")
    {
        namespace rv = ssme_b200::rveval;
        REQUIRE(std::abs(rv::evalUnivNorm<double>(0.3, 0.1, 2.0, true) - (-std::log(2.0) - 0.5 * std::log(2 * M_PI) - 0.5 * 0.01)) < 1e-14);
        REQUIRE(rv::evalUnivNorm<double>(0.3, 0.1, -1.0, true) == -INFINITY);
        REQUIRE(std::abs(rv::evalUniform<double>(0.5, 0.0, 4.0, true) + std::log(4.0)) < 1e-15);
        REQUIRE(rv::evalUniform<double>(4.5, 0.0, 4.0, true) == -INFINITY);
        REQUIRE(std::abs(rv::evalUnivInvGamma<double>(2.0, 3.0, 1.5, true) - (3 * std::log(1.5) - std::lgamma(3.0) - 4 * std::log(2.0) - 0.75)) < 1e-13);
        REQUIRE(std::abs(rv::twiceFisher<double>(.5) - std::log(3.0)) < 1e-15);
    }
    TEST_CASE("fixed-size algebra")
    {
        ssme_b200::mat<double, 3> a;
        const double vals[9] = {4, 1, 0.5, 1, 3, 0.2, 0.5, 0.2, 2};
        for (int i = 0; i < 9; ++i) a.m[i] = vals[i];
        auto l = ssme_b200::cholesky(a);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += l(i, k) * l(j, k);
                REQUIRE(std::abs(s - a(i, j)) < 1e-14);
            }
        a(2, 2) = -1;
        REQUIRE_THROWS_AS(ssme_b200::cholesky(a), std::runtime_error);
        ssme_b200::vec<double, 3> x{1, 2, 3}, y{0.5, -1, 2};
        REQUIRE(ssme_b200::outer(x, y)(2, 1) == -3.0);
        REQUIRE(((x + y) * 2.0)(0) == 3.0);
    }
    TEST_CASE("gpu_pool argument checks that need no device")
    {
        using pool_t = ssme_b200::gpu_pool<3, 1, double>;
        REQUIRE_THROWS_AS(pool_t(0, 100), std::invalid_argument);
        REQUIRE_THROWS_AS(pool_t(4, 0), std::invalid_argument);
        ssme_b200::gpu_options bad;
        bad.model = 9;
        REQUIRE_THROWS_AS(pool_t(4, 100, bad), std::invalid_argument);
    }
    return finish();
}
