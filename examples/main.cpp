// examples/main.cpp -- same command line as the reference's example/main.cpp:
//   ssme_example <datafile> <samples_base_name> <messages_base_name> <num mcmc iters> <num pfilters>
// 500 particles, float32 host arithmetic for the chain (example/main.cpp:9-13); the filters run in
// fp64 on the GPU.
#include <cstdlib>
#include <iostream>
#include <string>

#include "estimate_univ_svol.hpp"

#define NUMPARTS 500
#define DIMOBS 1
#define NUMPARAMS 3
#define DIMSTATE 1
#define FLOATTYPE float

int main(int argc, char* argv[])
{
    if (argc != 6) {
        std::cerr << "Please enter:\n"
                     "1.) datafile location, \n"
                     "2.) samples_base_name, \n"
                     "3.) messages file base name, \n"
                     "4.) number of mcmc iterations. \n"
                     "5.) number of pfilters. \n";
        return 0;
    }
    std::string data_loc = argv[1], samples_base_name = argv[2], messages_base_name = argv[3];
    unsigned int num_mcmc_iters = atoi(argv[4]), num_pfilters = atoi(argv[5]);
    try {
        do_ada_pmmh_univ_svol<NUMPARAMS, DIMSTATE, DIMOBS, NUMPARTS, FLOATTYPE>(data_loc, samples_base_name, messages_base_name,
                                                                               num_mcmc_iters, num_pfilters, false);
    } catch (const std::exception& e) {
        std::cerr << "ssme_example: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
