// tests/cpp/check.hpp -- a few lines of test scaffolding (Catch2, which the reference uses, is not in this image).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

static int g_checks = 0, g_failed = 0;
#define REQUIRE(...)                                                                   \
    do {                                                                                \
        ++g_checks;                                                                     \
        if (!(__VA_ARGS__)) { ++g_failed; std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #__VA_ARGS__); } \
    } while (0)
#define REQUIRE_THROWS_AS(expr, exc)                                                    \
    do {                                                                                \
        ++g_checks;                                                                     \
        bool ok_ = false;                                                               \
        try { (void)(expr); } catch (const exc&) { ok_ = true; } catch (...) {}         \
        if (!ok_) { ++g_failed; std::printf("FAILED %s:%d: %s did not throw %s\n", __FILE__, __LINE__, #expr, #exc); } \
    } while (0)
#define TEST_CASE(name) std::printf("-- %s\n", name);
static int finish() { std::printf("%d checks, %d failed\n", g_checks, g_failed); return g_failed ? 1 : 0; }
