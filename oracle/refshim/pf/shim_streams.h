// oracle/refshim/pf/shim_streams.h -- TEST INFRASTRUCTURE.
// Random-number plumbing of the pf stand-in (oracle/refshim/pf/*.h).  The real pf samplers own a std::mt19937 seeded
// from the clock.  The stand-in keeps that behaviour by default and adds two test hooks:
//   * pf::shim::set_base_seed(s): every sampler constructed afterwards is seeded deterministically (s, s+1, ...);
//   * playback streams: when a stream is armed, the samplers of that kind return its values in call order instead
//     of drawing -- "identical pre-generated normal and uniform streams" of BASELINE.json's north star.
// Not thread-safe by design (tests drive one filter at a time).
#ifndef SSME_REFSHIM_PF_STREAMS_H
#define SSME_REFSHIM_PF_STREAMS_H
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <random>
#include <stdexcept>

namespace pf {
namespace shim {

struct stream {
    const double* p = nullptr;
    size_t n = 0, pos = 0;
    bool active() const { return p != nullptr; }
    void arm(const double* q, size_t len) { p = q; n = len; pos = 0; }
    void disarm() { p = nullptr; n = pos = 0; }
    double next()
    {
        if (pos >= n) throw std::runtime_error("pf shim: injected stream exhausted");
        return p[pos++];
    }
};

inline stream& normal_stream() { static stream s; return s; }   // UnivNormSampler::sample
inline stream& uniform_stream() { static stream s; return s; }  // UniformSampler::sample (unit uniforms, scaled to [lo,hi))
inline stream& mvn_stream() { static stream s; return s; }      // MVNSampler::sample (dim normals per call)
inline stream& resamp_stream() { static stream s; return s; }   // resamplers (unit uniforms)
inline stream& kgen_stream() { static stream s; return s; }     // k_gen::sample (unit uniforms)

inline uint64_t& seed_state() { static uint64_t s = 0; return s; }
inline bool& seed_fixed() { static bool b = false; return b; }
inline void set_base_seed(uint64_t s) { seed_state() = s; seed_fixed() = true; }
inline void use_clock_seeds() { seed_fixed() = false; }
inline std::uint32_t next_seed()
{
    if (seed_fixed()) return static_cast<std::uint32_t>(seed_state()++);
    // pf: static_cast<uint32_t>(high_resolution_clock::now().time_since_epoch().count()), made distinct per object
    static std::uint32_t salt = 0;
    return static_cast<std::uint32_t>(std::chrono::high_resolution_clock::now().time_since_epoch().count()) + 0x9E3779B9u * (salt++);
}

// A 32-bit uniform random bit generator that makes std::generate_canonical<double,53> -- and therefore
// std::discrete_distribution / std::uniform_real_distribution<double> -- return exactly the next value of a playback
// stream: libstdc++ takes two 32-bit words g1, g2 and returns (g1 + g2 * 2^32) / 2^64 (bits/random.tcc, generate_canonical).
// A stream value u = m / 2^53 is reproduced by g1 + g2 * 2^32 = m * 2^11.
class playback_engine {
    stream* s;
    uint64_t bits = 0;
    int phase = 0;

public:
    using result_type = std::uint32_t;
    explicit playback_engine(stream& st) : s(&st) {}
    static constexpr result_type min() { return 0; }
    static constexpr result_type max() { return std::numeric_limits<result_type>::max(); }
    result_type operator()()
    {
        if (phase == 0) {
            const double u = s->next();
            if (!(u >= 0.0 && u < 1.0)) throw std::invalid_argument("pf shim: playback uniform outside [0,1)");
            bits = static_cast<uint64_t>(u * 9007199254740992.0) << 11;  // exact: u has at most 53 significant bits
            phase = 1;
            return static_cast<result_type>(bits & 0xffffffffu);
        }
        phase = 0;
        return static_cast<result_type>(bits >> 32);
    }
};

}  // namespace shim
}  // namespace pf
#endif
