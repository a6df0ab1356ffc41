// Drives phylostan_b200/stan/phylo_b200_stan.hpp the way a pystan-compiled model would: the include is
// pasted inside a model namespace after `using namespace stan::math;` (eigen/eigen.py:79-87).
//   shim_driver --no-gpu          : no handle published -> the shim must throw std::runtime_error
//   shim_driver problem.bin       : value (double overload) and value+gradient (var overload)
#include <cstdint>
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>
#include "stan_stub.hpp"
#include "phylo_b200.h"  // global scope for main(); inside the model namespace the guard makes it a no-op

namespace model_namespace {
using namespace stan::math;
#include "phylo_b200_stan.hpp"
}  // namespace model_namespace

typedef Eigen::Matrix<double, Eigen::Dynamic, 1> VecD;
typedef Eigen::Matrix<stan::math::var, Eigen::Dynamic, 1> VecV;

template <class V>
static V make(const std::vector<double>& x) {
    V v((int)x.size());
    for (size_t i = 0; i < x.size(); ++i) v((int)i) = x[i];
    return v;
}
static void dump(const char* name, const VecV& v) {
    std::printf("\"%s\": [", name);
    for (int i = 0; i < v.rows(); ++i) std::printf("%s%.17g", i ? ", " : "", v(i).adj());
    std::printf("]");
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    if (!std::strcmp(argv[1], "--no-gpu")) {
        try {
            VecD b(4);
            model_namespace::pruning_loglik(b, &std::cout);
        } catch (const std::runtime_error& e) {
            std::printf("ok: %s\n", e.what());
            return 0;
        }
        return 1;
    }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int32_t hdr[5];  // S, L, C, model, flags
    if (std::fread(hdr, 4, 5, f) != 5) return 4;
    const int S = hdr[0], L = hdr[1], C = hdr[2], model = hdr[3], flags = hdr[4];
    std::vector<int32_t> peel(3 * (S - 1));
    std::vector<uint8_t> tips((size_t)S * L);
    std::vector<double> w(L);
    const int bcount = (flags & PHYLO_B200_ROOTED) ? 2 * S - 2 : 2 * S - 3;
    const int ns = model == PHYLO_B200_GTR ? 6 : (model == PHYLO_B200_HKY ? 1 : 0);
    std::vector<double> bl(bcount), su(ns), fr(4), rs(C), ps(C);
    bool ok = std::fread(peel.data(), 4, peel.size(), f) == peel.size() && std::fread(tips.data(), 1, tips.size(), f) == tips.size() &&
              std::fread(w.data(), 8, L, f) == (size_t)L && std::fread(bl.data(), 8, bcount, f) == (size_t)bcount &&
              std::fread(su.data(), 8, ns, f) == (size_t)ns && std::fread(fr.data(), 8, 4, f) == 4 &&
              std::fread(rs.data(), 8, C, f) == (size_t)C && std::fread(ps.data(), 8, C, f) == (size_t)C;
    std::fclose(f);
    if (!ok) return 5;
    phylo_b200_handle h = 0;
    // PHYLO_SHIM_DEVICES="0,1,...": one handle over several GPUs (phylo_b200_create_multi); the Stan-facing
    // calls below are unchanged
    std::vector<int> devs;
    if (const char* dv = std::getenv("PHYLO_SHIM_DEVICES"))
        for (const char* q = dv; *q;) {
            devs.push_back(std::atoi(q));
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
    if (devs.size() > 1 ? phylo_b200_create_multi(&h, S, L, C, model, flags, peel.data(), tips.data(), w.data(), devs.data(),
                                                   (int)devs.size())
                        : phylo_b200_create(&h, S, L, C, model, flags, peel.data(), tips.data(), w.data(), 0)) {
        std::fprintf(stderr, "create: %s\n", phylo_b200_last_error());
        return 6;
    }
    phylo_b200_set_default(h);
    // data-only call (ADVI's ELBO draws): every argument double -> double overload
    const double v0 = model_namespace::phylo_loglik(make<VecD>(bl), make<VecD>(su), make<VecD>(fr), make<VecD>(rs),
                                                    make<VecD>(ps), &std::cout);
    // all parameters: var overload -> precomputed_gradients
    VecV vb = make<VecV>(bl), vs = make<VecV>(su), vf = make<VecV>(fr), vr = make<VecV>(rs), vp = make<VecV>(ps);
    stan::math::var r = model_namespace::phylo_loglik(vb, vs, vf, vr, vp, &std::cout);
    r.grad();
    // mixed: only branch lengths are parameters (fixed substitution model)
    VecV vb2 = make<VecV>(bl);
    stan::math::var r2 = model_namespace::phylo_loglik(vb2, make<VecD>(su), make<VecD>(fr), make<VecD>(rs),
                                                       make<VecD>(ps), &std::cout);
    r2.grad();
    std::printf("{\"value_double\": %.17g, \"value_var\": %.17g, \"n_operands\": %d, \"n_operands_mixed\": %d, ", v0, r.val(),
                (int)r.vi_->ops.size(), (int)r2.vi_->ops.size());
    dump("blens", vb); std::printf(", "); dump("subst", vs); std::printf(", "); dump("freqs", vf); std::printf(", ");
    dump("rs", vr); std::printf(", "); dump("ps", vp); std::printf(", "); dump("blens_mixed", vb2);
    if (ns == 0 && C == 1) {  // the reference's own operator surface
        VecV vb3 = make<VecV>(bl);
        stan::math::var r3 = model_namespace::pruning_loglik(vb3, &std::cout);
        r3.grad();
        std::printf(", \"pruning_loglik\": %.17g, ", r3.val());
        dump("pruning_grad", vb3);
        // domain error -> std::domain_error (Stan rejects the draw)
        VecD bad = make<VecD>(bl);
        bad(0) = -1.0;
        try {
            model_namespace::pruning_loglik(bad, &std::cout);
            std::printf(", \"domain_error\": false");
        } catch (const std::domain_error&) {
            std::printf(", \"domain_error\": true");
        }
    }
    if ((flags & PHYLO_B200_ROOTED) && argc > 2) {  // clock-tree front end: heights + rate through Stan types
        FILE* g = std::fopen(argv[2], "rb");
        std::vector<int32_t> map(2 * (2 * S - 1));
        std::vector<double> hts(S - 1), lowers(2 * S - 1), rate(1);
        bool ok2 = g && std::fread(map.data(), 4, map.size(), g) == map.size() && std::fread(hts.data(), 8, S - 1, g) == (size_t)(S - 1) &&
                   std::fread(lowers.data(), 8, 2 * S - 1, g) == (size_t)(2 * S - 1) && std::fread(rate.data(), 8, 1, g) == 1;
        if (g) std::fclose(g);
        if (!ok2) return 7;
        std::vector<std::vector<int> > m2(2 * S - 1, std::vector<int>(2));
        for (int i = 0; i < 2 * S - 1; ++i) { m2[i][0] = map[2 * i]; m2[i][1] = map[2 * i + 1]; }
        std::vector<stan::math::var> vh(hts.begin(), hts.end()), vrate(rate.begin(), rate.end());
        stan::math::var r4 = model_namespace::phylo_loglik_heights(vh, vrate, m2, lowers, make<VecD>(su), make<VecD>(fr),
                                                                   make<VecD>(rs), make<VecD>(ps), &std::cout);
        r4.grad();
        std::printf(", \"heights_value\": %.17g, \"heights_nops\": %d, \"heights_grad\": [", r4.val(), (int)r4.vi_->ops.size());
        for (int i = 0; i < S - 1; ++i) std::printf("%s%.17g", i ? ", " : "", vh[i].adj());
        std::printf("], \"rate_grad\": %.17g", vrate[0].adj());
        // autocorrelated variant with every substrate equal to the strict rate: same branch lengths
        std::vector<stan::math::var> vh2(hts.begin(), hts.end()), vsub;
        for (int i = 0; i < 2 * S - 2; ++i) vsub.push_back(stan::math::var(rate[0]));  // distinct operands
        stan::math::var r5 = model_namespace::phylo_loglik_heights_autocorr(vh2, vsub, m2, lowers, make<VecD>(su), make<VecD>(fr),
                                                                            make<VecD>(rs), make<VecD>(ps), &std::cout);
        r5.grad();
        double hdiff = 0.0, rsum = 0.0;
        for (int i = 0; i < S - 1; ++i) hdiff = std::max(hdiff, std::fabs(vh2[i].adj() - vh[i].adj()));
        for (int i = 0; i < 2 * S - 2; ++i) rsum += vsub[i].adj();
        std::printf(", \"autocorr_value\": %.17g, \"autocorr_heights_maxdiff\": %.17g, \"autocorr_rate_grad_sum\": %.17g",
                    r5.val(), hdiff, rsum);
    }
    std::printf("}\n");
    phylo_b200_destroy(h);
    return 0;
}
