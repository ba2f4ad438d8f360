// oracle/refshim/catch2/catch_all.hpp -- TEST INFRASTRUCTURE.  The three Catch2 macros the reference's test suite uses
// (TEST_CASE, TEST_CASE_METHOD, REQUIRE: test/*.cpp) so that the UNMODIFIED reference tests compile and run here
// (Catch2 is not installed).  CATCH_CONFIG_MAIN (test/test-main.cpp:1) emits main(): runs every registered case,
// prints a one-line summary per case, returns the number of failed cases.
#ifndef SSME_REFSHIM_CATCH_ALL_HPP
#define SSME_REFSHIM_CATCH_ALL_HPP
#include <algorithm>  // the real catch_all.hpp pulls these in; the reference's tests rely on it (thread_pool.h uses std::exp,
#include <cmath>      // std::max_element and std::function without including their headers: thread_pool.h:4-12 vs :55,263)
#include <chrono>
#include <cstdio>
#include <exception>
#include <functional>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace catch_shim {
struct test_case {
    std::string name, tags;
    std::function<void()> body;
};
inline std::vector<test_case>& registry()
{
    static std::vector<test_case> r;
    return r;
}
struct registrar {
    registrar(const char* name, const char* tags, std::function<void()> f) { registry().push_back({name, tags, std::move(f)}); }
};
struct failure : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline int& assertion_count()
{
    static int n = 0;
    return n;
}
inline int run_all()
{
    int failed = 0;
    for (auto& t : registry()) {
        try {
            t.body();
            std::printf("PASSED  %s %s\n", t.name.c_str(), t.tags.c_str());
        } catch (const std::exception& e) {
            ++failed;
            std::printf("FAILED  %s %s: %s\n", t.name.c_str(), t.tags.c_str(), e.what());
        }
    }
    std::printf("test cases: %zu | failed: %d | assertions: %d\n", registry().size(), failed, assertion_count());
    return failed;
}
}  // namespace catch_shim

#define CATCH_SHIM_CAT2(a, b) a##b
#define CATCH_SHIM_CAT(a, b) CATCH_SHIM_CAT2(a, b)

#define CATCH_SHIM_TEST_CASE(fn, ...)                                      \
    static void fn();                                                      \
    static catch_shim::registrar CATCH_SHIM_CAT(fn, _reg)(__VA_ARGS__, fn); \
    static void fn()
#define TEST_CASE(...) CATCH_SHIM_TEST_CASE(CATCH_SHIM_CAT(catch_shim_case_, __COUNTER__), __VA_ARGS__)

#define CATCH_SHIM_TEST_CASE_METHOD(cls, fixture, ...)                                              \
    namespace {                                                                                     \
    struct cls : fixture {                                                                          \
        void test();                                                                                \
    };                                                                                              \
    static catch_shim::registrar CATCH_SHIM_CAT(cls, _reg)(__VA_ARGS__, [] { cls obj; obj.test(); }); \
    }                                                                                               \
    void cls::test()
#define TEST_CASE_METHOD(fixture, ...) CATCH_SHIM_TEST_CASE_METHOD(CATCH_SHIM_CAT(catch_shim_fixture_, __COUNTER__), fixture, __VA_ARGS__)

#define REQUIRE(...)                                                                                                        \
    do {                                                                                                                    \
        ++catch_shim::assertion_count();                                                                                    \
        if (!(__VA_ARGS__)) throw catch_shim::failure(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": REQUIRE( " #__VA_ARGS__ " )"); \
    } while (0)

#ifdef CATCH_CONFIG_MAIN
int main() { return catch_shim::run_all(); }
#endif
#endif
