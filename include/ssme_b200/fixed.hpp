// include/ssme_b200/fixed.hpp -- minimal fixed-size vector / matrix types.
//
// The reference's signatures are written on Eigen::Matrix<float_t,n,1> / <float_t,n,n>
// (e.g. ada_pmmh_mvn.h:34-36 psv/psm/osv).  Eigen is not a dependency of this backend; these two
// small aggregates provide the operations the PMMH host loop needs (element access with (), +, -,
// scalar *, outer product, Cholesky) with the same call syntax, so user code written against the
// reference's typedefs compiles against these with a `using` change only.
#ifndef SSME_B200_FIXED_HPP
#define SSME_B200_FIXED_HPP

#include <array>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <ostream>
#include <stdexcept>

namespace ssme_b200 {

template <typename T, std::size_t N>
struct vec {
    std::array<T, N> v{};
    vec() = default;
    vec(std::initializer_list<T> il)
    {
        if (il.size() != N) throw std::invalid_argument("vec: wrong number of initialisers");
        std::size_t i = 0;
        for (T e : il) v[i++] = e;
    }
    static vec Zero() { return vec(); }
    static constexpr std::size_t size() { return N; }
    T& operator()(std::size_t i) { return v[i]; }
    const T& operator()(std::size_t i) const { return v[i]; }
    T& operator[](std::size_t i) { return v[i]; }
    const T& operator[](std::size_t i) const { return v[i]; }
    const T* data() const { return v.data(); }
    T* data() { return v.data(); }
    vec& operator+=(const vec& o) { for (std::size_t i = 0; i < N; ++i) v[i] += o.v[i]; return *this; }
    vec& operator-=(const vec& o) { for (std::size_t i = 0; i < N; ++i) v[i] -= o.v[i]; return *this; }
    vec& operator*=(T s) { for (auto& e : v) e *= s; return *this; }
    const vec& transpose() const { return *this; }  // printing helper, as osv.transpose() in the reference
};
template <typename T, std::size_t N> vec<T, N> operator+(vec<T, N> a, const vec<T, N>& b) { return a += b; }
template <typename T, std::size_t N> vec<T, N> operator-(vec<T, N> a, const vec<T, N>& b) { return a -= b; }
template <typename T, std::size_t N> vec<T, N> operator*(T s, vec<T, N> a) { return a *= s; }
template <typename T, std::size_t N> vec<T, N> operator*(vec<T, N> a, T s) { return a *= s; }
template <typename T, std::size_t N> vec<T, N> operator/(vec<T, N> a, T s) { for (auto& e : a.v) e /= s; return a; }
template <typename T, std::size_t N>
std::ostream& operator<<(std::ostream& os, const vec<T, N>& a)
{
    for (std::size_t i = 0; i < N; ++i) os << (i ? " " : "") << a(i);
    return os;
}

template <typename T, std::size_t N>
struct mat {
    std::array<T, N * N> m{};  // row-major
    static mat Zero() { return mat(); }
    static mat Identity()
    {
        mat r;
        for (std::size_t i = 0; i < N; ++i) r(i, i) = T(1);
        return r;
    }
    T& operator()(std::size_t i, std::size_t j) { return m[i * N + j]; }
    const T& operator()(std::size_t i, std::size_t j) const { return m[i * N + j]; }
    mat& operator+=(const mat& o) { for (std::size_t i = 0; i < N * N; ++i) m[i] += o.m[i]; return *this; }
    mat& operator-=(const mat& o) { for (std::size_t i = 0; i < N * N; ++i) m[i] -= o.m[i]; return *this; }
    mat& operator*=(T s) { for (auto& e : m) e *= s; return *this; }
};
template <typename T, std::size_t N> mat<T, N> operator+(mat<T, N> a, const mat<T, N>& b) { return a += b; }
template <typename T, std::size_t N> mat<T, N> operator-(mat<T, N> a, const mat<T, N>& b) { return a -= b; }
template <typename T, std::size_t N> mat<T, N> operator*(T s, mat<T, N> a) { return a *= s; }
template <typename T, std::size_t N> mat<T, N> operator*(mat<T, N> a, T s) { return a *= s; }
template <typename T, std::size_t N> mat<T, N> operator/(mat<T, N> a, T s) { for (auto& e : a.m) e /= s; return a; }
template <typename T, std::size_t N>
vec<T, N> operator*(const mat<T, N>& a, const vec<T, N>& x)
{
    vec<T, N> r;
    for (std::size_t i = 0; i < N; ++i)
        for (std::size_t j = 0; j < N; ++j) r(i) += a(i, j) * x(j);
    return r;
}
// a b^T
template <typename T, std::size_t N>
mat<T, N> outer(const vec<T, N>& a, const vec<T, N>& b)
{
    mat<T, N> r;
    for (std::size_t i = 0; i < N; ++i)
        for (std::size_t j = 0; j < N; ++j) r(i, j) = a(i) * b(j);
    return r;
}
// lower Cholesky factor; throws std::runtime_error if the matrix is not positive definite
template <typename T, std::size_t N>
mat<T, N> cholesky(const mat<T, N>& a)
{
    mat<T, N> l;
    for (std::size_t i = 0; i < N; ++i) {
        for (std::size_t j = 0; j <= i; ++j) {
            T s = a(i, j);
            for (std::size_t k = 0; k < j; ++k) s -= l(i, k) * l(j, k);
            if (i == j) {
                if (!(s > T(0))) throw std::runtime_error("cholesky: matrix is not positive definite");
                l(i, i) = std::sqrt(s);
            } else {
                l(i, j) = s / l(j, j);
            }
        }
    }
    return l;
}

}  // namespace ssme_b200
#endif
