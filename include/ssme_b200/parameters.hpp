// include/ssme_b200/parameters.hpp -- the param::pack<> parameter container of SSME, Eigen-free.
//
// Mirrors include/ssme/parameters.h of the reference: same namespace, class and method names,
// same exception types and messages' meaning, same transforms:
//   trans_type / string names   parameters.h:27, 290-313
//   null, twice_fisher, logit, log transforms and their log-Jacobians   parameters.h:317-449
//   pack ctor / copy / assign / add_param_and_transform / getters / get_log_jacobian   parameters.h:462-631
// Difference by design: the reference keeps one heap-allocated polymorphic transform per element
// (deep-cloned on every copy, parameters.h:488-500); here a pack is a trivially copyable pair of
// arrays, because a PMMH proposal crosses the C ABI as plain untransformed doubles.
#ifndef SSME_B200_PARAMETERS_HPP
#define SSME_B200_PARAMETERS_HPP

#include <array>
#include <cmath>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "fixed.hpp"

namespace param {

enum class trans_type { TT_null, TT_twice_fisher, TT_logit, TT_log };

inline trans_type trans_type_from_string(const std::string& tt)
{
    if (tt == "null") return trans_type::TT_null;
    if (tt == "twice_fisher") return trans_type::TT_twice_fisher;
    if (tt == "logit") return trans_type::TT_logit;
    if (tt == "log") return trans_type::TT_log;
    throw std::invalid_argument("that transform type was not accounted for");  // parameters.h:308
}

// the three operations of param::transform<float_t> (parameters.h:36-82), dispatched on the enum
template <typename float_t>
struct transform {
    static float_t trans(trans_type tt, const float_t& p)
    {
        switch (tt) {
        case trans_type::TT_null: return p;
        case trans_type::TT_twice_fisher:
            if ((p <= -1.0) || (p >= 1.0)) throw std::invalid_argument("error: phi was not between -1 and 1");
            return std::log(1.0 + p) - std::log(1.0 - p);
        case trans_type::TT_logit:
            if ((p < 0.0) || (p > 1.0)) throw std::invalid_argument("error: p was not between 0 and 1 \n");
            return std::log(p) - std::log(1.0 - p);
        case trans_type::TT_log:
            if (p < 0.0) throw std::invalid_argument("p is negative\n");
            return std::log(p);
        }
        throw std::invalid_argument("that transform type was not accounted for");
    }
    static float_t inv_trans(trans_type tt, const float_t& trans_p)
    {
        float_t ans;
        switch (tt) {
        case trans_type::TT_null: return trans_p;
        case trans_type::TT_twice_fisher:
            if (trans_p >= 0.0) ans = 2 / (1.0 + std::exp(-trans_p)) - 1.0;
            else ans = 1.0 - 2.0 / (1.0 + std::exp(trans_p));
            if ((ans <= -1.0) || (ans >= 1.0)) throw std::invalid_argument("error: there was probably overflow for exp(trans_p) \n");
            return ans;
        case trans_type::TT_logit:
            if (trans_p >= 0.0) ans = 1.0 / (1.0 + std::exp(-trans_p));
            else ans = std::exp(trans_p) / (1.0 + std::exp(trans_p));
            if ((ans <= 0.0) || (ans >= 1.0)) std::cerr << "error: there was probably underflow for exp(-r) \n";
            return ans;
        case trans_type::TT_log: return std::exp(trans_p);
        }
        throw std::invalid_argument("that transform type was not accounted for");
    }
    static float_t log_jacobian(trans_type tt, const float_t& trans_p)
    {
        switch (tt) {
        case trans_type::TT_null: return 0.0;
        case trans_type::TT_twice_fisher: return std::log(2.0) + trans_p - 2.0 * std::log(1.0 + std::exp(trans_p));
        case trans_type::TT_logit: return -trans_p - 2.0 * std::log(1.0 + std::exp(-trans_p));
        case trans_type::TT_log: return trans_p;
        }
        throw std::invalid_argument("that transform type was not accounted for");
    }
};

template <typename float_t, size_t numelem>
class pack {
public:
    using eig_vec = ssme_b200::vec<float_t, numelem>;
    using dyn_vec = std::vector<float_t>;  // the reference returns Eigen::Matrix<float_t, Dynamic, 1> for sub-ranges

    pack() : m_add_idx(0) {}

    pack(const eig_vec& params, const std::vector<std::string>& transform_names, bool from_transformed = true)
    {
        if (numelem != transform_names.size()) throw std::invalid_argument("params needs to be the right size (full)");
        m_add_idx = numelem;
        for (size_t i = 0; i < numelem; ++i) {
            m_ts[i] = trans_type_from_string(transform_names[i]);
            m_trans_params(i) = from_transformed ? params(i) : transform<float_t>::trans(m_ts[i], params(i));
        }
    }

    // copy construction / assignment of a partially filled pack throw, as the reference does (parameters.h:498, 558)
    pack(const pack& other) { copy_from(other, "copy ctor can only work with full parameter packs"); }
    pack& operator=(const pack& other)
    {
        copy_from(other, "pack assignment can only work with full parameter packs");
        return *this;
    }

    void add_param_and_transform(float_t elem, trans_type tt, bool is_transformed = false)
    {
        if (m_add_idx >= numelem) throw std::length_error("can't add any more transformations");
        m_ts[m_add_idx] = tt;
        m_trans_params(m_add_idx) = is_transformed ? elem : transform<float_t>::trans(tt, elem);
        m_add_idx++;
    }
    void add_param_and_transform(float_t elem, const std::string& trans_name, bool is_transformed = false)
    {
        add_param_and_transform(elem, trans_type_from_string(trans_name), is_transformed);
    }

    unsigned size() const { return m_add_idx; }
    size_t capacity() const { return numelem; }

    eig_vec get_trans_params() const
    {
        require_full();
        return m_trans_params;
    }
    eig_vec get_untrans_params() const
    {
        require_full();
        eig_vec p;
        for (size_t i = 0; i < numelem; ++i) p(i) = transform<float_t>::inv_trans(m_ts[i], m_trans_params(i));
        return p;
    }
    // inclusive index range, "not like python indexing" (parameters.h:222-225)
    dyn_vec get_trans_params(const unsigned int& start, const unsigned int& end) const
    {
        dyn_vec r;
        for (unsigned i = start; i <= end; ++i) r.push_back(m_trans_params(i));
        return r;
    }
    dyn_vec get_untrans_params(const unsigned int& start, const unsigned int& end) const
    {
        require_full();
        dyn_vec r;
        for (unsigned i = start; i <= end; ++i) r.push_back(transform<float_t>::inv_trans(m_ts[i], m_trans_params(i)));
        return r;
    }
    float_t get_log_jacobian() const
    {
        require_full();
        float_t result(0.0);
        for (size_t i = 0; i < m_add_idx; ++i) result += transform<float_t>::log_jacobian(m_ts[i], m_trans_params(i));
        return result;
    }
    trans_type get_transform(unsigned i) const { return m_ts[i]; }

private:
    void require_full() const
    {
        if (m_add_idx < numelem) throw std::length_error("the parameter container is not full");
    }
    void copy_from(const pack& other, const char* msg)
    {
        if (other.size() != numelem) throw std::invalid_argument(msg);
        m_add_idx = numelem;
        m_trans_params = other.m_trans_params;
        m_ts = other.m_ts;
    }
    eig_vec m_trans_params;
    std::array<trans_type, numelem> m_ts{};
    unsigned m_add_idx;
};

}  // namespace param
#endif
