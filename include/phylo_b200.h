/*
 * phylo_b200.h -- C ABI of libphylo_b200.so: Felsenstein-pruning log-likelihood and its
 * analytic gradient on one B200 (sm_100a), behind the reference's operator surface.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *
 *   phylo_b200_create[_tipdata]  the compile-time constants baked into eigen.hpp by
 *                                eigen/eigen.j2:19-38 (child_parent, postorder, tip partials, pi, Q)
 *                                == the Stan data block peel/tipdata/weights/C of
 *                                phylostan/generate_script.py:1186-1197 filled by
 *                                phylostan/phylostan.py:183,204,255-272
 *   phylo_b200_eval              value_grad vbsky_loglik(const vector<double>&)  eigen/eigen.j2:56-168
 *                                and the model-block loops of generate_script.py:961-1055 with the
 *                                P-matrix functions of generate_script.py:755-892
 *   phylo_b200_eval (want_grad=0) double pruning_loglik(Matrix<double,-1,1>, ostream*) eigen/eigen.j2:171-177
 *   phylo_b200_eval (want_grad=1) var pruning_loglik(Matrix<var,-1,1>, ostream*)  eigen/prune_stan.hpp:9-17
 *                                (value + the gradient vector handed to precomputed_gradients)
 *   phylo_b200_eval_batch        no reference counterpart (Stan evaluates draws serially,
 *                                phylostan/phylostan.py:311-313); used for batches of ELBO draws
 *   struct of outputs            struct value_grad { double log_P; VectorXd grad; }  eigen/value_grad.hpp:5-8
 *
 * Conventions are the reference's: node ids 1-based, tips 1..S, internals S+1..2S-1 in post-order,
 * root 2S-1 (phylostan/utils.py:59-72); blens[k-1] is the branch above node k
 * (generate_script.py:660-679); states A,C,G,T (utils.py:180).  The gradient returned is the TRUE
 * d logL / d blens (eigen/eigen.j2:165 returns blens[k] times that).
 *
 * All functions return 0 on success, a negative PHYLO_B200_E* code otherwise;
 * phylo_b200_last_error() describes the last failure on the calling thread.  There is no CPU
 * fallback: without a usable sm_100 device phylo_b200_create fails.
 */
#ifndef PHYLO_B200_H
#define PHYLO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHYLO_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define PHYLO_B200_API __attribute__((visibility("default")))
#else
#define PHYLO_B200_API
#endif

/* substitution models (generate_script.py:755, :783, :839) */
#define PHYLO_B200_JC69 0 /* no parameters, freqs fixed to 1/4                      */
#define PHYLO_B200_HKY  1 /* subst = {kappa}                                       */
#define PHYLO_B200_GTR  2 /* subst = {AC, AG, AT, CG, CT, GT} (generate_script.py:855-858) */

/* flags */
#define PHYLO_B200_ROOTED   1 /* clock tree, bcount = 2S-2; otherwise unrooted, bcount = 2S-3
                                 and node 2S-2 carries no branch (generate_script.py:1034)   */
#define PHYLO_B200_NO_NORMQ 2 /* do not normalise Q to one substitution per unit time
                                 (eigen/eigen.py:47-50 bakes such a Q)                       */

/* error codes */
#define PHYLO_B200_EINVAL   -1 /* bad argument (sizes, peel order, NULL)         */
#define PHYLO_B200_ECUDA    -2 /* CUDA runtime / launch failure                  */
#define PHYLO_B200_ENODEV   -3 /* no sm_100 device                               */
#define PHYLO_B200_EDOMAIN  -4 /* non-finite or out-of-domain parameter / result */
#define PHYLO_B200_ENOMEM   -5

typedef struct phylo_b200_ctx *phylo_b200_handle;

/*
 * peel     int32 [S-1][3]  (child1, child2, parent), post-order, 1-based
 * tipmask  uint8 [S][L]    bit s set <=> tipdata[tip][pattern][s] != 0
 * weights  double [L]      pattern weights (NULL = all 1)
 * C        number of rate categories (>= 1)
 * device   CUDA device ordinal
 */
PHYLO_B200_API int phylo_b200_create(phylo_b200_handle *out, int S, int L, int C, int model, int flags,
                      const int32_t *peel, const uint8_t *tipmask, const double *weights,
                      int device);

/* Same, with the reference's own Stan-data layout: tipdata double [S][L][4] in {0,1}
 * ("real tipdata[S,L,4]", generate_script.py:1188). */
PHYLO_B200_API int phylo_b200_create_tipdata(phylo_b200_handle *out, int S, int L, int C, int model, int flags,
                              const int32_t *peel, const double *tipdata, const double *weights,
                              int device);

/* Same as phylo_b200_create, with the alignment already resident on the GPU: d_tipmask uint8 [S][L] and
 * d_weights double [L] (NULL = all 1) are DEVICE pointers on `device` (the "device-resident buffers" of the
 * data plumbing; phylostan/phylostan.py:183-204 builds the same arrays on the host).  Padding, the
 * one-hot / all-ones classification and the re-coding run on the device; the inputs are not kept. */
PHYLO_B200_API int phylo_b200_create_device(phylo_b200_handle *out, int S, int L, int C, int model, int flags,
                                            const int32_t *peel, const uint8_t *d_tipmask, const double *d_weights,
                                            int device);

/*
 * One handle, N devices of this box (`devices, ndev`): the site patterns are cut into ndev
 * contiguous shards (shard i = patterns [L i / ndev, L (i+1) / ndev) on devices[i]), every device holds the
 * full tree, and every evaluation entry point of this header (eval, eval_batch, eval_heights*, upload / run /
 * download) runs all shards concurrently on their own streams and adds the shards' result rows on
 * devices[0] -- peer loads over NVLink when the devices can access each other, staged peer copies otherwise
 * -- so the Stan-facing call shape (eigen/prune_stan.hpp:9-17) reaches every GPU from one process.  The
 * exchange is the one the pattern sum `target += log(...) * weights[i]` (generate_script.py:1010) needs:
 * [B][nout] doubles per evaluation.  device_out / download / get_timing refer to devices[0]; a device may be
 * listed more than once (two shards on one GPU; used by the single-GPU tests).  ndev == 1 is phylo_b200_create.
 * phylo_b200_info(h, 12) returns the number of shards.
 */
PHYLO_B200_API int phylo_b200_create_multi(phylo_b200_handle *out, int S, int L, int C, int model, int flags,
                                           const int32_t *peel, const uint8_t *tipmask, const double *weights,
                                           const int *devices, int ndev);

PHYLO_B200_API void phylo_b200_destroy(phylo_b200_handle h);

/* sizes of the per-evaluation arrays */
PHYLO_B200_API int phylo_b200_bcount(phylo_b200_handle h);  /* 2S-2 rooted, 2S-3 unrooted */
PHYLO_B200_API int phylo_b200_nsubst(phylo_b200_handle h);  /* 0 / 1 / 6                   */
PHYLO_B200_API int phylo_b200_ncat(phylo_b200_handle h);
/* length of one packed output row: 1 + bcount + nsubst + 4 + C + C
 * = [logL | d/dblens | d/dsubst | d/dfreqs | d/drs | d/dps] */
PHYLO_B200_API int phylo_b200_nout(phylo_b200_handle h);

/*
 * One evaluation with HOST arrays (blocking).  Inputs: blens[bcount], subst[nsubst] (may be NULL
 * when nsubst == 0), freqs[4] (ignored for JC69; may be NULL), rs[C], ps[C] (NULL: rs = 1, ps = 1/C).
 * Outputs: logp[1]; with want_grad != 0 any non-NULL g_* array is filled:
 * g_blens[bcount], g_subst[nsubst], g_freqs[4], g_rs[C], g_ps[C].  d/dsubst and d/dfreqs are the
 * unconstrained partial derivatives of the formulas at generate_script.py:799-812 / 855-868;
 * d/dfreqs includes the root term of generate_script.py:1007.  JC69 handles fix the frequencies to 1/4
 * (generate_script.py:755-780 takes no freqs argument): freqs is not read and g_freqs is returned as zeros.
 */
PHYLO_B200_API int phylo_b200_eval(phylo_b200_handle h, const double *blens, const double *subst,
                    const double *freqs, const double *rs, const double *ps, int want_grad,
                    double *logp, double *g_blens, double *g_subst, double *g_freqs,
                    double *g_rs, double *g_ps);

/* B independent parameter draws, row-major leading dimension B on every array. */
PHYLO_B200_API int phylo_b200_eval_batch(phylo_b200_handle h, int B, const double *blens, const double *subst,
                          const double *freqs, const double *rs, const double *ps, int want_grad,
                          double *logp, double *g_blens, double *g_subst, double *g_freqs,
                          double *g_rs, double *g_ps);

/* The same, but a draw that phylo_b200_eval would reject does not fail the batch: status[d] = 0 (evaluated),
 * 1 (parameters out of domain: negative or non-finite branch length, rates, frequencies ... -- not evaluated) or
 * 2 (log-likelihood not finite: impossible pattern or underflow).  For status[d] != 0, logp[d] = -inf and the
 * draw's gradient rows are zero, which is what Stan does with a rejected draw (std::domain_error in
 * eigen/prune_stan.hpp's caller).  Returns 0 whenever the call itself ran, whatever the draws' states. */
PHYLO_B200_API int phylo_b200_eval_batch_status(phylo_b200_handle h, int B, const double *blens, const double *subst,
                          const double *freqs, const double *rs, const double *ps, int want_grad,
                          double *logp, double *g_blens, double *g_subst, double *g_freqs,
                          double *g_rs, double *g_ps, int32_t *status);

/*
 * Node-height front end for clock trees (replaces the generated Stan loop `heights_to_blens`,
 * phylostan/generate_script.py:660-679, and its reverse sweep on Stan's tape):
 *     blens[node] = rate_node * (heights[parent] - heights[node])        internal node
 *     blens[node] = rate_node * (heights[parent] - lowers[node])         tip (lowers = NULL: 0)
 * map      int32 [2S-1][2] rows (node, parent) in pre-order, 1-based, row 0 = (root, 0)  (utils.py:84-90)
 * heights  double [S-1], heights[k] belongs to internal node S+1+k
 * rates    nrates == 1: strict clock `rate`; nrates == 2S-2: per-branch `substrates[node]`
 * Returns d logL / d heights [S-1] and d logL / d rates [nrates]; the other outputs as phylo_b200_eval.
 * Rooted handles only.
 */
PHYLO_B200_API int phylo_b200_eval_heights(phylo_b200_handle h, const int32_t *map, const double *heights,
                                           const double *lowers, const double *rates, int nrates,
                                           const double *subst, const double *freqs, const double *rs,
                                           const double *ps, int want_grad, double *logp, double *g_heights,
                                           double *g_rates, double *g_subst, double *g_freqs, double *g_rs,
                                           double *g_ps);

/*
 * The same front end for the autocorrelated clocks (acln, acg, ace, aoup, hsmrf, gmrf; replaces
 * `heights_to_blens_autocorr`, phylostan/generate_script.py:682-708): the rate of a branch is the mean
 * of the rates at its two ends,
 *     blens[node] = 1/2 (substrates[node] + substrates[parent]) * (heights[parent] - heights|lowers[node])
 * with substrates[map[2,1]] standing in for the root's rate, and the branch above map[2,1] (the first
 * child of the root in the pre-order map) using substrates[map[2,1]] alone.  nrates must be 2S-2.
 */
PHYLO_B200_API int phylo_b200_eval_heights_autocorr(phylo_b200_handle h, const int32_t *map, const double *heights,
                                                    const double *lowers, const double *rates, int nrates,
                                                    const double *subst, const double *freqs, const double *rs,
                                                    const double *ps, int want_grad, double *logp,
                                                    double *g_heights, double *g_rates, double *g_subst,
                                                    double *g_freqs, double *g_rs, double *g_ps);

/*
 * The same front ends for B draws with the heights -> blens step and its reverse sweep ON THE DEVICE
 * (generate_script.py:660-679 / :682-708 as a kernel in front of the P-matrix
 * kernel, the d/dheights and d/drates gather as a kernel behind the contraction; one H2D of
 * [heights | rates], one D2H of the results, no host loop over branches).  autocorr selects the :682-708 form.
 * heights [B][S-1], rates [B][nrates], outputs with leading dimension B; the map and `lowers` are uploaded once
 * per distinct map.  A draw whose heights give a negative or non-finite branch length fails the call with
 * PHYLO_B200_EDOMAIN (as phylo_b200_eval_batch does for such blens).
 */
PHYLO_B200_API int phylo_b200_eval_heights_batch(phylo_b200_handle h, int autocorr, const int32_t *map, int B,
                                                 const double *heights, const double *lowers, const double *rates,
                                                 int nrates, const double *subst, const double *freqs,
                                                 const double *rs, const double *ps, int want_grad, double *logp,
                                                 double *g_heights, double *g_rates, double *g_subst,
                                                 double *g_freqs, double *g_rs, double *g_ps);

/*
 * ... and with the ratio parametrisation of the heights on the device as well (generate_script.py:711-735
 * `transform`, :738-752 its log-Jacobian): props [B][S-2] (pre-order of the internal non-root nodes) and
 * root_height [B] in; logp [B] (likelihood only), logjac [B], heights [B][S-1] (optional) out.  The reverse
 * sweep returns d(logp + logjac)/dprops [B][S-2] and /droot_height [B]; hbar_extra [B][S-1] (optional) is an
 * additional adjoint of the heights -- a tree prior's gradient -- pushed through the same sweep.
 */
PHYLO_B200_API int phylo_b200_eval_ratios_batch(phylo_b200_handle h, int autocorr, const int32_t *map, int B,
                                                const double *props, const double *root_height, const double *lowers,
                                                const double *rates, int nrates, const double *subst,
                                                const double *freqs, const double *rs, const double *ps,
                                                const double *hbar_extra, int want_grad, double *logp, double *logjac,
                                                double *heights, double *g_props, double *g_root, double *g_rates,
                                                double *g_subst, double *g_freqs, double *g_rs, double *g_ps);

/*
 * Host-only helpers (no GPU) for drivers that keep the ratio parametrisation of the node heights outside
 * Stan: heights = transform(props, root_height, map, lowers) of phylostan/generate_script.py:711-735
 * with its log-Jacobian (:738-752), for B draws, and the reverse sweep through it.  props [B][S-2] are
 * consumed in pre-order of the internal non-root nodes; lowers may be NULL (contemporaneous tips).
 * ratios_reverse: hbar [B][S-1] holds d(everything downstream)/dheights on entry and is used as scratch;
 * the log-Jacobian's own derivative is added inside.
 */
PHYLO_B200_API int phylo_b200_ratios_forward(int S, const int32_t *map, const double *lowers, int B,
                                             const double *props, const double *root_height, double *heights,
                                             double *logjac);
PHYLO_B200_API int phylo_b200_ratios_reverse(int S, const int32_t *map, const double *lowers, int B,
                                             const double *props, const double *heights, double *hbar,
                                             double *g_props, double *g_root);

/*
 * Split form of eval_batch for callers that keep parameters resident on the device between
 * evaluations (benchmarks, batched drivers, multi-GPU ranks):
 *   upload   host parameters -> device (one H2D copy); also derives the eigen system per draw
 *   run      P-matrix kernel + sweep kernel(s) + contraction on the handle's stream, no host sync;
 *            results stay in the device output buffer [B][nout]
 *   device_out   device pointer of that buffer (e.g. for an NCCL all-reduce across pattern shards)
 *   download synchronises the stream and copies [B][nout] to the host
 */
PHYLO_B200_API int phylo_b200_upload(phylo_b200_handle h, int B, const double *blens, const double *subst,
                      const double *freqs, const double *rs, const double *ps);
PHYLO_B200_API int phylo_b200_run(phylo_b200_handle h, int B, int want_grad);
PHYLO_B200_API int phylo_b200_device_out(phylo_b200_handle h, void **dptr, int *ld);
PHYLO_B200_API int phylo_b200_download(phylo_b200_handle h, int B, double *out /* [B][nout] */);

/* Use a caller-owned CUDA stream (cudaStream_t as void*); NULL restores the handle's own. */
PHYLO_B200_API int phylo_b200_set_stream(phylo_b200_handle h, void *stream);
PHYLO_B200_API int phylo_b200_sync(phylo_b200_handle h);

/* Tuning: patterns per thread (1, 2 or 4) and pattern blocks (warps per category) per CTA;
 * 0 = automatic.  Takes effect on the next run. */
PHYLO_B200_API int phylo_b200_set_tiling(phylo_b200_handle h, int patterns_per_thread, int pattern_blocks);

/* Tuning / testing: shared-memory stack slots for gradient runs (0 = automatic).  Fewer slots than the
 * tree's stack depth park the top stack positions in the per-CTA HBM scratch (phylo_b200_info 0 =
 * depth, 11 = slots used by the last run).  Value-only runs always keep the whole stack on chip. */
PHYLO_B200_API int phylo_b200_set_stack_slots(phylo_b200_handle h, int slots);

/* Tuning / testing: where fp64 gradient runs with 4 patterns per thread keep their stack of pending vectors.
 * ctas_per_sm = 0: shared memory (two CTAs per SM); 2 or 3: tensor memory (tcgen05.ld / tcgen05.st as a per-thread
 * scratchpad), the children's partials staged through a shared-memory operand ring, that many CTAs per SM.
 * -1 = the library's default (also what the environment variable PHYLO_B200_SWEEP_TM=0|2|3 overrides at create
 * time).  Runs the variant does not cover (value only, fp32, K != 4, JC69 scalar kernel) ignore it;
 * phylo_b200_info 13 = what the last run used. */
PHYLO_B200_API int phylo_b200_set_sweep_variant(phylo_b200_handle h, int ctas_per_sm);

/* Tuning: cherry tables (on by default).  In message-statistic runs with 4 patterns per thread, the message of a
 * cherry -- an internal node whose two children are tips -- is looked up in a 25-entry table per (draw, category,
 * cherry) built from the 5 x 5 code pairs of its tips, instead of being stored by the post-order sweep and read back
 * by the pre-order sweep: a third less scratch traffic on coalescent trees (the site-repeat idea of the reference's
 * pruner/tree.cpp:140-174 in the form that fits this design).  Pitchforks -- a cherry and a tip -- get 125-entry
 * tables the same way (PHYLO_B200_TABLE_TIPS=2 in the environment at create time: cherries only).  PHYLO_B200_CHERRY=0 in the environment at create time
 * turns them off as well; phylo_b200_info 15 = whether the last run used them. */
PHYLO_B200_API int phylo_b200_set_cherry_tables(phylo_b200_handle h, int enabled);

/* Arithmetic of the sweeps: 64 (default; the parity-tested product path) or 32, the optional
 * "fp32 with scaling" mode: partials, transition matrices and 4x4 statistics in float with
 * power-of-two rescaling in units of 2^24, log-likelihood and gradient sums in double.  Its error is
 * reported separately; it is not covered by the 1e-10 / 1e-8 parity guarantee. */
PHYLO_B200_API int phylo_b200_set_precision(phylo_b200_handle h, int bits);

/* Device time of the kernels of the last run (CUDA events on the handle's stream):
 * ms[0] P-matrix kernel, ms[1] sweep kernel, ms[2] contraction kernel, ms[3] whole run.
 * Enable before the run; disabled by default (events force a sync when read). */
PHYLO_B200_API int phylo_b200_set_timing(phylo_b200_handle h, int enabled);
PHYLO_B200_API int phylo_b200_get_timing(phylo_b200_handle h, double ms[4]);

/* Introspection: what = 0 stack depth D, 1 patterns per thread, 2 threads per CTA, 3 grid size,
 * 4 dynamic shared memory bytes, 5 padded pattern count, 6 kernels launched by the last run,
 * 7 scratch bytes allocated on the device, 8 / 9 post- / pre-order stack depth, 10 pattern tiles,
 * 11 on-chip stack slots of the last run, 12 pattern shards (devices) behind the handle,
 * 13 sweep variant of the last run (0 shared-memory stack, 2 / 3 tensor-memory stack with that many CTAs per SM),
 * 14 whether the last gradient run used the message statistic (fp64, simple tips, 128-thread CTAs, more than one
 * pattern per thread and
 * (largest branch length) x (largest site rate) x (spread of Q's eigenvalues) + log(|m1|_F |m2|_F / 4) < 12 for every
 * draw of the batch;
 * PHYLO_B200_MSG=0 in the environment at create time turns it off),
 * 15 whether that run took the messages of cherries (internal nodes with two tip children) from per-(draw, category)
 * 25-entry tables instead of storing and re-reading them (phylo_b200_set_cherry_tables),
 * 16 whether its post-order took them from the tables as well (second plan; PHYLO_B200_POST_TABLES=0 turns that off). */
PHYLO_B200_API long long phylo_b200_info(phylo_b200_handle h, int what);

/*
 * Host-only hooks (no GPU needed), used by the CPU test-suite:
 *   plan    the depth-first traversal plan derived from `peel` (replaces the run-time std::map
 *           bookkeeping of eigen/eigen.j2:82-108).  The most recent vector stays in registers (TOS);
 *           operand sources are -1 tip, -2 TOS, >= 0 shared-memory slot.  post gets S-1 rows of 8 int32
 *           (a, b, src_a, src_b, spill_slot, node, 0, 0), pre gets S-1 rows of 12 int32
 *           (node, a, b, src_node, dst_b, a_internal, row_node, row_a, row_b, park_node, park_b, 0),
 *           depth[2] = {post-order, pre-order} shared-memory stack depth (TOS excluded).
 *   derive  the per-draw model algebra of generate_script.py:799-825 / 855-881:
 *           out = [pi 4 | lambda 4 | m1 16 | m2 16 | Q 16 | X_theta ntheta*16]; returns ntheta.
 */
PHYLO_B200_API int phylo_b200_plan(int S, const int32_t *peel, int32_t *post, int32_t *pre, int32_t *depth);
/* plan_tables: the message-table nodes of the tree (internal nodes with two or, when max_tips = 3, three tips below
 *           them; node_tab[2S-1] = table index or -1) and the second plan, whose post-order treats them as leaves:
 *           post gets info[0] rows (a leafified child has source -1), pre gets S-1 rows in which a node the post-order
 *           skips has row_node = -1 and fields 9 / 10 hold the parking rows of the node and of b;
 *           info[4] = {post-order steps, table nodes, table entries per (draw, category), stack depth}. */
PHYLO_B200_API int phylo_b200_plan_tables(int S, const int32_t *peel, int max_tips, int32_t *node_tab, int32_t *post,
                                          int32_t *pre, int32_t *info);
PHYLO_B200_API int phylo_b200_derive(int model, int flags, const double *subst, const double *freqs, double *out);

/*
 * Process-wide default handle.  The reference bakes tree + alignment into the generated eigen.hpp
 * (eigen/eigen.j2:19-38), so its Stan-facing pruning_loglik(blens) takes no data argument.  Here the
 * host program (phylostan run) creates the handle once and publishes it; the Stan shim
 * (phylostan_b200/stan/phylo_b200_stan.hpp), compiled into the model and linked against this
 * library, fetches it.  set_default does not take ownership.
 */
PHYLO_B200_API int phylo_b200_set_default(phylo_b200_handle h);
PHYLO_B200_API phylo_b200_handle phylo_b200_get_default(void);

PHYLO_B200_API const char *phylo_b200_last_error(void);
PHYLO_B200_API int phylo_b200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PHYLO_B200_H */
