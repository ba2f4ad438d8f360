// include/ssme_b200/utils.hpp -- CSV input of SSME (reference include/ssme/utils.h), Eigen-free.
//   utils::read_data<nc,float_t>       utils.h:25-64   header-less, comma separated, one row per line
//   utils::csv_param_sampler           utils.h:75-141  uniform re-draws of rows of a parameter CSV
// Reader behaviour kept: a missing file prints to stderr and yields an empty vector (:35-37); a line
// with an unparsable number is reported and skipped (:53-56).  Tightened: a row with fewer than nc
// numbers is skipped with a message instead of being read past its end (:59 has no length check).
#ifndef SSME_B200_UTILS_HPP
#define SSME_B200_UTILS_HPP

#include <chrono>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "fixed.hpp"

namespace utils {

template <size_t nc, typename float_t>
std::vector<ssme_b200::vec<float_t, nc>> read_data(const std::string& file_loc)
{
    std::vector<ssme_b200::vec<float_t, nc>> data;
    std::string line;
    std::ifstream ifs;
    ifs.open(file_loc);
    std::string one_number;
    if (!ifs.is_open()) std::cerr << "utils::read_data() failed to read data from: " << file_loc << "\n";
    while (std::getline(ifs, line)) {
        std::vector<float_t> data_row;
        try {
            std::istringstream buff(line);
            while (std::getline(buff, one_number, ',')) data_row.push_back(std::stod(one_number));
        } catch (const std::invalid_argument& ia) {
            std::cerr << "Invalid Argument: " << ia.what() << "\n";
            continue;
        }
        if (data_row.size() < nc) {
            std::cerr << "utils::read_data(): skipping a row with " << data_row.size() << " of " << nc << " columns\n";
            continue;
        }
        ssme_b200::vec<float_t, nc> drw;
        for (size_t i = 0; i < nc; ++i) drw(i) = data_row[i];
        data.push_back(drw);
    }
    return data;
}

template <size_t dimparam, typename float_t>
class csv_param_sampler {
public:
    using psv = ssme_b200::vec<float_t, dimparam>;
    csv_param_sampler() = delete;
    explicit csv_param_sampler(const std::string& param_csv_filename)
        : m_gen{static_cast<std::uint32_t>(std::chrono::high_resolution_clock::now().time_since_epoch().count())}
    {
        init(param_csv_filename);
    }
    csv_param_sampler(const std::string& param_csv_filename, unsigned long seed) : m_gen{static_cast<std::uint32_t>(seed)}
    {
        init(param_csv_filename);
    }
    psv samp() { return m_param_samps[m_idx_sampler(m_gen)]; }
    size_t num_rows() const { return m_param_samps.size(); }

private:
    void init(const std::string& f)
    {
        m_param_samps = read_data<dimparam, float_t>(f);
        if (m_param_samps.empty()) throw std::runtime_error("csv_param_sampler: no parameter rows in " + f);
        m_idx_sampler = std::uniform_int_distribution<int>(0, (int)m_param_samps.size() - 1);
    }
    std::mt19937 m_gen;
    std::uniform_int_distribution<int> m_idx_sampler;
    std::vector<psv> m_param_samps;
};

}  // namespace utils
#endif
