/*
 * phylo_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the reference's tree-likelihood hot path.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call this.  The product (phylostan_b200/csrc) never does.
 *
 * Follows (paths relative to /root/reference):
 *   eigen/eigen.j2:56-168            post-order, pre-order, analytic branch gradient
 *   phylostan/generate_script.py:755-892   JC69 / HKY / GTR P-matrices
 *   phylostan/generate_script.py:998-1040  rooted / unrooted mixture likelihood
 *   pruner/tree.cpp:201-242          p / q recursion (sum_s p q = L invariant)
 *
 * Parity pinning: the 3-taxon closed form of eigen/test_ll_3tax.py (executed
 * from the reference with torch standing in for jax; tests/golden/ll_3tax.json).
 * HKY/GTR/Weibull and substitution-parameter gradients are NOT pinned by any
 * reference test ("parity unpinned" for those); they are checked against
 * finite differences and an independent torch-autograd restatement instead.
 */
#ifndef PHYLO_ORACLE_H
#define PHYLO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_JC69 = 0, ORACLE_HKY = 1, ORACLE_GTR = 2 };

/* flags */
enum {
    ORACLE_ROOTED      = 1,  /* clock tree: bcount = 2S-2; else unrooted, bcount = 2S-3 */
    ORACLE_NO_NORMQ    = 2,  /* keep Q un-normalised (eigen/eigen.py:47-50 uses mu=.25 JC) */
    ORACLE_RESCALE     = 4,  /* per-(pattern,category) rescaling; off = reference arithmetic */
    ORACLE_QUIRK_TIMES = 8,  /* multiply d/dblens[b] by blens[b] (eigen/eigen.j2:165) */
    ORACLE_DP_EIGEN    = 16  /* dP/dtheta by the eigen F-matrix formula instead of Van Loan's
                                block exponential (faster; used for CPU-baseline timing) */
};

/*
 * peel     int32 [S-1][3], 1-based node ids (child1, child2, parent) in post-order
 *          (phylostan/utils.py:75-81); root = 2S-1.
 * tipmask  uint8 [S][L]; bit s set <=> tipdata[tip][pattern][s] == 1
 *          (phylostan/utils.py:180-188: one-hot A,C,G,T, anything else 1,1,1,1).
 * weights  double [L] pattern weights.
 * blens    double [bcount]; blens[k] = branch above node k+1.
 * subst    JC69: ignored; HKY: kappa[1]; GTR: rates[6] (AC,AG,AT,CG,CT,GT).
 * grads    any output pointer may be NULL.  grad_subst has 0/1/6 entries and is the
 *          UNCONSTRAINED partial derivative of the formula (Stan applies simplex
 *          Jacobians upstream).  grad_freqs includes the root term.
 * nthreads OpenMP threads over patterns (<=0: all).
 * returns 0 on success.
 */
int oracle_loglik_grad(int S, int L, int C, const int32_t *peel, const uint8_t *tipmask,
                       const double *weights, int model, int flags, const double *blens,
                       const double *subst, const double *freqs, const double *rs,
                       const double *ps, int want_grad, int nthreads, double *logp,
                       double *grad_blens, double *grad_subst, double *grad_freqs,
                       double *grad_rs, double *grad_ps);

/* per-pattern site log-likelihoods (length L), same conventions; no gradient */
int oracle_site_loglik(int S, int L, int C, const int32_t *peel, const uint8_t *tipmask,
                       int model, int flags, const double *blens, const double *subst,
                       const double *freqs, const double *rs, const double *ps,
                       double *site_logl);

/* P(t) for one branch/category exactly as the generated Stan does (row-major 4x4). */
int oracle_pmatrix(int model, int flags, const double *subst, const double *freqs, double tau,
                   double *P);

/* max over nodes of |sum_s p(s) q(s) - L| / L for pattern `l`, category `c`
 * (pruner/test.cpp:44 invariant).  No rescaling. */
double oracle_pq_invariant(int S, int L, int C, const int32_t *peel, const uint8_t *tipmask,
                           int model, int flags, const double *blens, const double *subst,
                           const double *freqs, const double *rs, const double *ps, int l, int c);

int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
