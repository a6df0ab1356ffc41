// phylo_b200_stan.hpp -- Stan <-> libphylo_b200 boundary shim.
//
// Drop-in for the reference's eigen/prune_stan.hpp:9-17 (var overload -> precomputed_gradients) and
// eigen/eigen.j2:171-177 (double overload), generalised from "branch lengths, baked JC Q, one
// category" to every differentiable input of the production likelihood
// (phylostan/generate_script.py:755-892, 961-1055).
//
// Use exactly as the reference uses prune_stan.hpp (eigen/eigen.py:79-87, eigen/util.py:117-126):
//   pystan.StanModel(file=..., allow_undefined=True,
//                    includes=["phylo_b200_stan.hpp"], include_dirs=[<repo>/phylostan_b200/stan, <repo>/include],
//                    extra_compile_args=["--std=c++14"], + link against libphylo_b200.so)
// pystan pastes the include inside the generated model namespace after `using namespace stan::math;`
// which is why everything below is fully qualified and `inline`.
//
// Stan declarations (functions block, no body):
//   real pruning_loglik(vector blens);                                         // eigen/example.stan:3
//   real phylo_loglik(vector blens, vector subst, vector freqs, vector rs, vector ps);
//   real phylo_loglik_heights(real[] heights, real[] rates, int[,] map, real[] lowers,
//                             vector subst, vector freqs, vector rs, vector ps);         // clock trees
//
// Unlike eigen.hpp, the tree and alignment are not baked in at compile time: the host program creates
// a handle (phylo_b200_create) and publishes it with phylo_b200_set_default before sampling starts.
//
// Errors follow Stan's convention: std::domain_error rejects the current draw,
// anything else is fatal.
#ifndef PHYLO_B200_STAN_HPP
#define PHYLO_B200_STAN_HPP

#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "phylo_b200.h"

namespace phylo_b200_stan {

inline phylo_b200_handle handle() {
    phylo_b200_handle h = phylo_b200_get_default();
    if (!h)
        throw std::runtime_error("phylo_b200: no default handle; call phylo_b200_set_default() "
                                 "(phylostan_b200.likelihood.set_default) before running the model");
    return h;
}

inline void check(int rc) {
    if (rc == 0) return;
    const std::string msg = std::string("phylo_b200: ") + phylo_b200_last_error();
    if (rc == PHYLO_B200_EDOMAIN) throw std::domain_error(msg);  // Stan rejects the draw
    throw std::runtime_error(msg);
}

// value_of + operand collection that works for both `double` and `stan::math::var` vectors
template <typename T>
inline void gather(const Eigen::Matrix<T, Eigen::Dynamic, 1>& v, std::vector<double>& x) {
    x.reserve(x.size() + v.rows());
    for (int i = 0; i < v.rows(); ++i) x.push_back(stan::math::value_of(v(i)));
}
inline void push_operands(const Eigen::Matrix<stan::math::var, Eigen::Dynamic, 1>& v, const double* g,
                          std::vector<stan::math::var>& ops, std::vector<double>& grads) {
    for (int i = 0; i < v.rows(); ++i) {
        ops.push_back(v(i));
        grads.push_back(g[i]);
    }
}
inline void push_operands(const Eigen::Matrix<double, Eigen::Dynamic, 1>&, const double*,
                          std::vector<stan::math::var>&, std::vector<double>&) {}

template <typename R>
struct finish;
template <>
struct finish<double> {
    static double go(double v, const std::vector<stan::math::var>&, const std::vector<double>&) { return v; }
};
template <>
struct finish<stan::math::var> {
    static stan::math::var go(double v, const std::vector<stan::math::var>& ops, const std::vector<double>& g) {
        return stan::math::precomputed_gradients(v, ops, g);  // operands.size() == gradients.size()
    }
};

template <typename T>
struct is_double { enum { value = 0 }; };
template <>
struct is_double<double> { enum { value = 1 }; };

}  // namespace phylo_b200_stan

// real phylo_loglik(vector blens, vector subst, vector freqs, vector rs, vector ps)
//   blens  [bcount]  branch above node k at position k (generate_script.py:660-679)
//   subst  [] JC69 | [kappa] HKY | [6] GTR rates (AC,AG,AT,CG,CT,GT)
//   freqs  [4] (ignored for JC69), rs [C], ps [C]
template <typename T_bl, typename T_su, typename T_fr, typename T_rs, typename T_ps>
inline typename stan::return_type<T_bl, T_su, T_fr, T_rs, T_ps>::type phylo_loglik(
    const Eigen::Matrix<T_bl, Eigen::Dynamic, 1>& blens, const Eigen::Matrix<T_su, Eigen::Dynamic, 1>& subst,
    const Eigen::Matrix<T_fr, Eigen::Dynamic, 1>& freqs, const Eigen::Matrix<T_rs, Eigen::Dynamic, 1>& rs,
    const Eigen::Matrix<T_ps, Eigen::Dynamic, 1>& ps, std::ostream* /*pstream__*/) {
    namespace p = phylo_b200_stan;
    typedef typename stan::return_type<T_bl, T_su, T_fr, T_rs, T_ps>::type R;
    phylo_b200_handle h = p::handle();
    const int bcount = phylo_b200_bcount(h), nsubst = phylo_b200_nsubst(h), C = phylo_b200_ncat(h);
    if (blens.rows() != bcount || subst.rows() != nsubst || rs.rows() != C || ps.rows() != C ||
        (nsubst > 0 && freqs.rows() != 4))
        throw std::invalid_argument("phylo_loglik: argument sizes do not match the published tree/alignment");
    std::vector<double> xb, xs, xf, xr, xp;
    p::gather(blens, xb); p::gather(subst, xs); p::gather(freqs, xf); p::gather(rs, xr); p::gather(ps, xp);
    const int want_grad = !p::is_double<R>::value;
    double logp = 0.0;
    std::vector<double> gb(bcount), gs(nsubst > 0 ? nsubst : 1), gf(4), gr(C), gp(C);
    p::check(phylo_b200_eval(h, xb.data(), nsubst ? xs.data() : 0, xf.size() == 4 ? xf.data() : 0, xr.data(),
                             xp.data(), want_grad, &logp, gb.data(), gs.data(), gf.data(), gr.data(), gp.data()));
    std::vector<stan::math::var> ops;
    std::vector<double> grads;
    if (want_grad) {
        p::push_operands(blens, gb.data(), ops, grads);
        p::push_operands(subst, gs.data(), ops, grads);
        if (nsubst > 0 && freqs.rows() == 4) p::push_operands(freqs, gf.data(), ops, grads);  // JC69 fixes freqs: no operand
        p::push_operands(rs, gr.data(), ops, grads);
        p::push_operands(ps, gp.data(), ops, grads);
    }
    return p::finish<R>::go(logp, ops, grads);
}

// real phylo_loglik_heights(real[] heights, real[] rates, int[,] map, real[] lowers,
//                           vector subst, vector freqs, vector rs, vector ps)
// Clock trees: replaces the generated heights -> blens loop (generate_script.py:660-679) together with
// the likelihood.  rates = {rate} (strict clock) or substrates[2S-2]; lowers = tip dates (2S-1 entries).
namespace phylo_b200_stan {
template <typename T>
inline void gather(const std::vector<T>& v, std::vector<double>& x) {
    x.reserve(x.size() + v.size());
    for (size_t i = 0; i < v.size(); ++i) x.push_back(stan::math::value_of(v[i]));
}
inline void push_operands(const std::vector<stan::math::var>& v, const double* g, std::vector<stan::math::var>& ops,
                          std::vector<double>& grads) {
    for (size_t i = 0; i < v.size(); ++i) {
        ops.push_back(v[i]);
        grads.push_back(g[i]);
    }
}
inline void push_operands(const std::vector<double>&, const double*, std::vector<stan::math::var>&,
                          std::vector<double>&) {}
}  // namespace phylo_b200_stan

namespace phylo_b200_stan {
template <typename T_h, typename T_r, typename T_su, typename T_fr, typename T_rs, typename T_ps>
inline typename stan::return_type<T_h, T_r, T_su, T_fr, T_rs, T_ps>::type loglik_heights(
    bool autocorr, const std::vector<T_h>& heights, const std::vector<T_r>& rates,
    const std::vector<std::vector<int> >& map, const std::vector<double>& lowers,
    const Eigen::Matrix<T_su, Eigen::Dynamic, 1>& subst, const Eigen::Matrix<T_fr, Eigen::Dynamic, 1>& freqs,
    const Eigen::Matrix<T_rs, Eigen::Dynamic, 1>& rs, const Eigen::Matrix<T_ps, Eigen::Dynamic, 1>& ps) {
    namespace p = phylo_b200_stan;
    typedef typename stan::return_type<T_h, T_r, T_su, T_fr, T_rs, T_ps>::type R;
    phylo_b200_handle h = p::handle();
    const int bcount = phylo_b200_bcount(h), nsubst = phylo_b200_nsubst(h), C = phylo_b200_ncat(h);
    const int S = bcount / 2 + 1;
    if ((int)heights.size() != S - 1 || (int)map.size() != 2 * S - 1 || (int)lowers.size() != 2 * S - 1 ||
        ((int)rates.size() != bcount && (autocorr || (int)rates.size() != 1)) || subst.rows() != nsubst || rs.rows() != C ||
        ps.rows() != C || (nsubst > 0 && freqs.rows() != 4))
        throw std::invalid_argument("phylo_loglik_heights: argument sizes do not match the published tree/alignment");
    std::vector<int32_t> fmap(2 * map.size());
    for (size_t i = 0; i < map.size(); ++i) {
        fmap[2 * i] = map[i][0];
        fmap[2 * i + 1] = map[i][1];
    }
    std::vector<double> xh, xr, xs, xf, xrs, xp;
    p::gather(heights, xh); p::gather(rates, xr); p::gather(subst, xs); p::gather(freqs, xf); p::gather(rs, xrs);
    p::gather(ps, xp);
    const int want_grad = !p::is_double<R>::value;
    double logp = 0.0;
    std::vector<double> gh(S - 1), gr(rates.size()), gs(nsubst > 0 ? nsubst : 1), gf(4), grs(C), gp(C);
    p::check((autocorr ? phylo_b200_eval_heights_autocorr : phylo_b200_eval_heights)(
        h, fmap.data(), xh.data(), lowers.data(), xr.data(), (int)xr.size(), nsubst ? xs.data() : 0,
        xf.size() == 4 ? xf.data() : 0, xrs.data(), xp.data(), want_grad, &logp, gh.data(), gr.data(), gs.data(),
        gf.data(), grs.data(), gp.data()));
    std::vector<stan::math::var> ops;
    std::vector<double> grads;
    if (want_grad) {
        p::push_operands(heights, gh.data(), ops, grads);
        p::push_operands(rates, gr.data(), ops, grads);
        p::push_operands(subst, gs.data(), ops, grads);
        if (nsubst > 0 && freqs.rows() == 4) p::push_operands(freqs, gf.data(), ops, grads);  // JC69 fixes freqs: no operand
        p::push_operands(rs, grs.data(), ops, grads);
        p::push_operands(ps, gp.data(), ops, grads);
    }
    return p::finish<R>::go(logp, ops, grads);
}
}  // namespace phylo_b200_stan

template <typename T_h, typename T_r, typename T_su, typename T_fr, typename T_rs, typename T_ps>
inline typename stan::return_type<T_h, T_r, T_su, T_fr, T_rs, T_ps>::type phylo_loglik_heights(
    const std::vector<T_h>& heights, const std::vector<T_r>& rates, const std::vector<std::vector<int> >& map,
    const std::vector<double>& lowers, const Eigen::Matrix<T_su, Eigen::Dynamic, 1>& subst,
    const Eigen::Matrix<T_fr, Eigen::Dynamic, 1>& freqs, const Eigen::Matrix<T_rs, Eigen::Dynamic, 1>& rs,
    const Eigen::Matrix<T_ps, Eigen::Dynamic, 1>& ps, std::ostream* /*pstream__*/) {
    return phylo_b200_stan::loglik_heights(false, heights, rates, map, lowers, subst, freqs, rs, ps);
}

// real phylo_loglik_heights_autocorr(real[] heights, real[] substrates, int[,] map, real[] lowers,
//                                    vector subst, vector freqs, vector rs, vector ps)
// Autocorrelated clocks: replaces heights_to_blens_autocorr (generate_script.py:682-708) + the likelihood.
template <typename T_h, typename T_r, typename T_su, typename T_fr, typename T_rs, typename T_ps>
inline typename stan::return_type<T_h, T_r, T_su, T_fr, T_rs, T_ps>::type phylo_loglik_heights_autocorr(
    const std::vector<T_h>& heights, const std::vector<T_r>& rates, const std::vector<std::vector<int> >& map,
    const std::vector<double>& lowers, const Eigen::Matrix<T_su, Eigen::Dynamic, 1>& subst,
    const Eigen::Matrix<T_fr, Eigen::Dynamic, 1>& freqs, const Eigen::Matrix<T_rs, Eigen::Dynamic, 1>& rs,
    const Eigen::Matrix<T_ps, Eigen::Dynamic, 1>& ps, std::ostream* /*pstream__*/) {
    return phylo_b200_stan::loglik_heights(true, heights, rates, map, lowers, subst, freqs, rs, ps);
}

// real pruning_loglik(vector blens) -- the reference's own operator (eigen/prune_stan.hpp:9-17,
// eigen/eigen.j2:171-177): substitution model, frequencies and site rates are whatever the published
// handle was created with (JC69, one category for the eigen/ examples).  Returns the TRUE gradient;
// eigen.j2:165 multiplies each component by its branch length ("reference quirks").
template <typename T_bl>
inline typename stan::return_type<T_bl>::type pruning_loglik(const Eigen::Matrix<T_bl, Eigen::Dynamic, 1>& blens,
                                                             std::ostream* /*pstream__*/) {
    namespace p = phylo_b200_stan;
    typedef typename stan::return_type<T_bl>::type R;
    phylo_b200_handle h = p::handle();
    const int bcount = phylo_b200_bcount(h);
    // eigen/vbsky_fix_rate.stan:238 passes 2S-1 entries (one unused slot)
    if (blens.rows() < bcount) throw std::invalid_argument("pruning_loglik: blens is shorter than the branch count");
    if (phylo_b200_nsubst(h) != 0 || phylo_b200_ncat(h) != 1)
        throw std::invalid_argument("pruning_loglik(blens) needs a JC69, single-category handle; use phylo_loglik");
    std::vector<double> xb;
    p::gather(blens, xb);
    const int want_grad = !p::is_double<R>::value;
    double logp = 0.0;
    std::vector<double> gb(blens.rows(), 0.0);
    p::check(phylo_b200_eval(h, xb.data(), 0, 0, 0, 0, want_grad, &logp, gb.data(), 0, 0, 0, 0));
    std::vector<stan::math::var> ops;
    std::vector<double> grads;
    if (want_grad) p::push_operands(blens, gb.data(), ops, grads);  // extra slots get gradient 0
    return p::finish<R>::go(logp, ops, grads);
}

#endif  // PHYLO_B200_STAN_HPP
