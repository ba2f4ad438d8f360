// examples/main.cpp -- adaptive PMMH for the univariate stochastic-volatility model on the GPU backend.
//
// Takes the five positional arguments of the reference's example program (example/main.cpp:17-37), in the same order, so
// existing run scripts keep working:
//   ssme_example DATA.csv SAMPLES_PREFIX MESSAGES_PREFIX MCMC_ITERATIONS FILTERS_PER_PROPOSAL
// The chain runs in single precision on the host with 500 particles per filter, as the reference's example does
// (example/main.cpp:9-13); the filters themselves run on the GPU.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <string>

#include "estimate_univ_svol.hpp"

namespace {

constexpr std::size_t kParticles = 500;
constexpr std::size_t kParams = 3;  // beta, phi, sigma^2
constexpr std::size_t kDimState = 1;
constexpr std::size_t kDimObs = 1;
using chain_float = float;

struct options {
    std::string data, samples_prefix, messages_prefix;
    unsigned iterations = 0, filters = 0;
};

bool parse_count(const char* text, unsigned& out)
{
    errno = 0;
    char* end = nullptr;
    const unsigned long v = std::strtoul(text, &end, 10);
    if (errno != 0 || end == text || *end != '\0' || v == 0 || v > 0xffffffffUL) return false;
    out = static_cast<unsigned>(v);
    return true;
}

bool parse(int argc, char** argv, options& o)
{
    if (argc != 6) return false;
    o.data = argv[1];
    o.samples_prefix = argv[2];
    o.messages_prefix = argv[3];
    return parse_count(argv[4], o.iterations) && parse_count(argv[5], o.filters);
}

}  // namespace

int main(int argc, char** argv)
{
    options o;
    if (!parse(argc, argv, o)) {
        std::fprintf(stderr,
                     "usage: %s DATA.csv SAMPLES_PREFIX MESSAGES_PREFIX MCMC_ITERATIONS FILTERS_PER_PROPOSAL\n"
                     "  DATA.csv              one observation per row\n"
                     "  SAMPLES_PREFIX        posterior draws go to <prefix>_<timestamp> files\n"
                     "  MESSAGES_PREFIX       acceptance / adaptation log, same naming\n"
                     "  MCMC_ITERATIONS       length of the chain (positive integer)\n"
                     "  FILTERS_PER_PROPOSAL  particle filters averaged per likelihood estimate (positive integer)\n",
                     argc > 0 ? argv[0] : "ssme_example");
        return 2;
    }
    try {
        do_ada_pmmh_univ_svol<kParams, kDimState, kDimObs, kParticles, chain_float>(o.data, o.samples_prefix, o.messages_prefix, o.iterations,
                                                                                   o.filters, /*multicore (ignored: the GPU is the pool)*/ false);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ssme_example: %s\n", e.what());
        return 1;
    }
    return 0;
}
