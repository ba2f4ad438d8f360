// oracle/ref_cpu.cpp -- TEST / BASELINE INFRASTRUCTURE.  The reference's CPU path for the hot path,
// timed by bench.py (cpu_baseline, --impl reference).  Never linked into the product.
//
// What is the reference's own code here: the dispatcher.  When /root/reference is present at build
// time (oracle/Makefile `ref`, -DSSME_HAVE_REFERENCE_POOL) the proposals are fanned out through the
// UNMODIFIED include/ssme/thread_pool.h, compiled from where it lies (std-only header; it needs
// <cmath>/<algorithm>/<functional> included first: thread_pool.h:4-12 vs :55,263).
// What is restated (kind = "port"): the filter arithmetic, which lives in the external library
// tbrown122387/pf (absent, unpinned): bootstrap filter step per include/ssme/liu_west_filter.h:1608-1761,
// model per example/univ_svol_bootstrap_filter.h:54-103, multinomial resampling through
// std::discrete_distribution exactly as pf::resamplers::mn_resampler does, RNG = std::mt19937 +
// std::normal_distribution as pf::rvsamp::UnivNormSampler does.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <functional>
#include <future>
#include <limits>
#include <mutex>
#include <random>
#include <thread>
#include <vector>
#include <array>

#ifdef SSME_HAVE_REFERENCE_POOL
#include <ssme/thread_pool.h>
#define SSME_POOL_KIND 1
#else
#define SSME_POOL_KIND 0
// API-identical local pool (ctor / add_observed_data / work), used only when the reference tree was
// not available when this file was compiled.
template <typename dyn_data_t, typename static_data_t, typename func_output_t, bool debug = false>
class thread_pool {
public:
    using F = std::function<func_output_t(dyn_data_t, static_data_t)>;
    thread_pool(F f, unsigned num_comps, bool mt = true, unsigned num_threads = 0)
        : m_f(f), m_total(num_comps), m_threads(num_threads)
    {
        if (m_threads == 0) m_threads = mt ? std::max(1u, std::thread::hardware_concurrency()) : 1u;
        if (m_threads > 1 && !mt) throw std::invalid_argument("can't request single threaded and multiple threads at the same time!");
        if (mt && m_threads == 1) throw std::runtime_error("requested multiple threads but only one is available");
    }
    void add_observed_data(const static_data_t& d)
    {
        if (m_have) throw std::runtime_error("you already called add_observed_data once before!");
        m_data = d;
        m_have = true;
    }
    func_output_t work(dyn_data_t p)
    {
        if (!m_have) throw std::runtime_error("must add observed data before calculating anything\n");
        std::vector<func_output_t> vals(m_total);
        std::atomic_uint next{0};
        auto body = [&]() {
            for (unsigned i = next++; i < m_total; i = next++) vals[i] = m_f(p, m_data);
        };
        std::vector<std::thread> ts;
        for (unsigned t = 1; t < m_threads; ++t) ts.emplace_back(body);
        body();
        for (auto& t : ts) t.join();
        func_output_t m = *std::max_element(vals.begin(), vals.end());
        func_output_t s = 0;
        for (auto v : vals) s += std::exp(v - m);
        return m + std::log(s) - std::log(m_total);
    }

private:
    F m_f;
    unsigned m_total, m_threads;
    static_data_t m_data;
    bool m_have = false;
};
#endif

namespace {

std::atomic<uint64_t> g_seed_counter{1};

double eval_univ_norm_log(double x, double mu, double sigma)
{
    double exponent = -.5 * (x - mu) * (x - mu) / (sigma * sigma);
    if (sigma > 0.0) return -std::log(sigma) - .5 * std::log(2.0 * M_PI) + exponent;
    return -std::numeric_limits<double>::infinity();
}

struct params_t {
    int model;  // 0 SV (beta, phi, sigma^2), 1 leverage (phi, mu, sigma, rho)
    int N;
    uint64_t seed;
    double th[4];
};

// one bootstrap filter over the whole series; its own engines, like each pf filter object
double run_filter(const params_t& p, const std::vector<double>& y)
{
    const int N = p.N;
    const uint64_t s = p.seed + 0x9E3779B97F4A7C15ull * g_seed_counter.fetch_add(1);
    std::mt19937 gen_norm{static_cast<std::uint32_t>(s)}, gen_res{static_cast<std::uint32_t>(s >> 32)};
    std::normal_distribution<double> norm(0.0, 1.0);
    double beta = 1.0, phi, sigma, mu = 0.0, rho = 0.0;
    if (p.model == 0) { beta = p.th[0]; phi = p.th[1]; sigma = std::sqrt(p.th[2]); }
    else { phi = p.th[0]; mu = p.th[1]; sigma = p.th[2]; rho = p.th[3]; }
    std::vector<double> x(N), xn(N), lw(N, 0.0), w(N);
    double loglik = 0.0;
    for (size_t t = 0; t < y.size(); ++t) {
        const double yt = y[t];
        double cl;
        if (t == 0) {
            for (int i = 0; i < N; ++i) {
                x[i] = norm(gen_norm) * sigma / std::sqrt(1. - phi * phi);
                const double sd0 = sigma / std::sqrt(1.0 - phi * phi);
                lw[i] = eval_univ_norm_log(x[i], 0.0, sd0);
                lw[i] += eval_univ_norm_log(yt, 0.0, beta * std::exp(.5 * x[i]));
                lw[i] -= eval_univ_norm_log(x[i], 0.0, sd0);
            }
            double m = *std::max_element(lw.begin(), lw.end());
            double sumexp = 0.0;
            for (int i = 0; i < N; ++i) sumexp += std::exp(lw[i] - m);
            cl = -std::log(N) + m + std::log(sumexp);
        } else {
            double maxOld = -std::numeric_limits<double>::infinity();
            std::vector<double>& old = w;  // reuse storage for the old weights
            for (int i = 0; i < N; ++i) {
                if (lw[i] > maxOld) maxOld = lw[i];
                old[i] = lw[i];
                if (p.model == 0) {
                    x[i] = phi * x[i] + norm(gen_norm) * sigma;
                } else {
                    double xt = mu + phi * (x[i] - mu) + y[t - 1] * rho * sigma * std::exp(-.5 * x[i]);
                    xt += norm(gen_norm) * sigma * std::sqrt(1.0 - rho * rho);
                    x[i] = xt;
                }
                lw[i] += eval_univ_norm_log(yt, 0.0, beta * std::exp(.5 * x[i]));
            }
            double maxNumer = *std::max_element(lw.begin(), lw.end());
            double s1 = 0.0, s2 = 0.0;
            for (int i = 0; i < N; ++i) {
                s1 += std::exp(lw[i] - maxNumer);
                s2 += std::exp(old[i] - maxOld);
            }
            cl = maxNumer + std::log(s1) - maxOld - std::log(s2);
        }
        loglik += cl;
        // mn_resampler::resampLogWts: exp(lw - max) -> std::discrete_distribution -> N draws -> lw = 0
        double m = *std::max_element(lw.begin(), lw.end());
        for (int i = 0; i < N; ++i) w[i] = std::exp(lw[i] - m);
        std::discrete_distribution<> idx(w.begin(), w.end());
        for (int j = 0; j < N; ++j) xn[j] = x[idx(gen_res)];
        x.swap(xn);
        std::fill(lw.begin(), lw.end(), 0.0);
    }
    return loglik;
}

}  // namespace

extern "C" {

// 1 = dispatched through the reference's own thread_pool.h, 0 = local API-identical pool
int ssme_refcpu_pool_kind(void) { return SSME_POOL_KIND; }

unsigned ssme_refcpu_hardware_threads(void) { return std::max(1u, std::thread::hardware_concurrency()); }

// P proposals, each evaluated as thread_pool::work(theta): R replicate filters + log-mean-exp.
// num_threads = 0 -> the reference's default policy (hardware_concurrency, thread_pool.h:131-137);
// num_threads = 1 -> single-threaded as the shipped example runs (example/main.cpp:42, mc = false).
// Returns 0 and the wall time of the work() calls only (pool construction excluded).
int ssme_refcpu_loglike_batch(int model, int N, const double* y, int64_t T, const double* theta, int num_params, int P,
                              unsigned R, unsigned num_threads, uint64_t seed, double* out, double* seconds, unsigned* threads_used)
{
    try {
        using dyn_t = params_t;
        using static_t = std::vector<double>;
        const bool mt = (num_threads != 1);
        if (mt && num_threads == 0 && std::thread::hardware_concurrency() < 2) num_threads = 1;
        const bool mt2 = (num_threads != 1);
        // pool_func takes its arguments BY VALUE, as ada_pmmh_mvn::pool_func does (ada_pmmh_mvn.h:147)
        thread_pool<dyn_t, static_t, double> pool([](dyn_t p, static_t data) { return run_filter(p, data); }, R, mt2, num_threads);
        pool.add_observed_data(static_t(y, y + T));
        if (threads_used) *threads_used = mt2 ? (num_threads ? num_threads : std::max(1u, std::thread::hardware_concurrency())) : 1u;
        auto t0 = std::chrono::steady_clock::now();
        for (int p = 0; p < P; ++p) {
            dyn_t d;
            d.model = model;
            d.N = N;
            d.seed = seed + (uint64_t)p;
            for (int k = 0; k < 4; ++k) d.th[k] = (k < num_params) ? theta[(size_t)p * num_params + k] : 0.0;
            out[p] = pool.work(d);
        }
        auto t1 = std::chrono::steady_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

}  // extern "C"
