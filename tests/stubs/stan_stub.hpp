// Minimal stand-in for the slice of Stan Math (2.19 API) that the Stan shim touches:
// var, value_of, precomputed_gradients, return_type.  One-level reverse sweep only.
#pragma once
#include <stdexcept>
#include <type_traits>
#include <vector>
#include <Eigen/Dense>
namespace stan {
namespace math {
struct vari {
    double val, adj;
    std::vector<vari*> ops;
    std::vector<double> g;
    explicit vari(double v) : val(v), adj(0) {}
};
struct var {
    vari* vi_;
    var() : vi_(new vari(0)) {}
    var(double v) : vi_(new vari(v)) {}
    double val() const { return vi_->val; }
    double adj() const { return vi_->adj; }
    void grad() {
        vi_->adj = 1;
        for (size_t i = 0; i < vi_->ops.size(); ++i) vi_->ops[i]->adj += vi_->g[i];
    }
};
inline double value_of(double x) { return x; }
inline double value_of(const var& v) { return v.vi_->val; }
inline var precomputed_gradients(double value, const std::vector<var>& operands, const std::vector<double>& gradients) {
    if (operands.size() != gradients.size()) throw std::invalid_argument("precomputed_gradients: size mismatch");
    var r(value);
    for (size_t i = 0; i < operands.size(); ++i) r.vi_->ops.push_back(operands[i].vi_);
    r.vi_->g = gradients;
    return r;
}
}  // namespace math
template <typename... T>
struct return_type;
template <>
struct return_type<> { typedef double type; };
template <typename T, typename... Rest>
struct return_type<T, Rest...> {
    typedef typename std::conditional<std::is_same<T, math::var>::value, math::var,
                                      typename return_type<Rest...>::type>::type type;
};
}  // namespace stan
