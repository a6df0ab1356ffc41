/*
 * phylo_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See phylo_oracle.h.
 *
 * Plain C restatement of the reference algorithm (paths relative to /root/reference):
 *   P-matrices      phylostan/generate_script.py:755-892
 *   post-order      eigen/eigen.j2:122-141  == generate_script.py:998-1005 / 1025-1034
 *   pre-order       eigen/eigen.j2:144-157  == pruner/tree.cpp:228-242
 *   L and gradient  eigen/eigen.j2:160-167, generate_script.py:1006-1010 / 1035-1040
 * Parity pins (tests/test_oracle.py): the closed form of eigen/test_ll_3tax.py (tests/golden/ll_3tax.json); the
 * reference's own C++ vbsky_loglik, eigen/eigen.j2 rendered, compiled and run by oracle/build_ref.py, for JC69 / HKY /
 * GTR rate matrices -- log_P and the branch gradient (tests/golden/ref_eigen_cpp.json); the NumPy GTR class of
 * scripts/phylo.py for P(t) (tests/golden/gtr_pt.json); SURVEY App. B anchors.
 * Extensions the reference only gets through Stan's autodiff (unpinned by reference tests):
 *   d/d(rates|kappa), d/dfreqs, d/drs, d/dps  -- done here through per-branch 4x4
 *   statistics contracted with dP/dtheta, where dP/dtheta comes from Van Loan's
 *   block-triangular matrix exponential (default) or the eigen "F-matrix" formula.
 */
#include "phylo_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------- small dense helpers */

static void mat4_mul(const double *A, const double *B, double *Cm) {
    double T[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
            T[4 * i + j] = s;
        }
    memcpy(Cm, T, sizeof T);
}

/* cyclic Jacobi for a symmetric 4x4; eigenvalues ascending like eigenvalues_sym */
static void eig_sym4(const double *Ain, double *lam, double *U) {
    double A[16];
    memcpy(A, Ain, sizeof A);
    for (int i = 0; i < 16; ++i) U[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0;
        for (int i = 0; i < 4; ++i)
            for (int j = i + 1; j < 4; ++j) off += A[4 * i + j] * A[4 * i + j];
        if (off < 1e-300) break;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) {
                double apq = A[4 * p + q];
                if (apq == 0.0) continue;
                double theta = (A[4 * q + q] - A[4 * p + p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) { /* A <- A J */
                    double akp = A[4 * k + p], akq = A[4 * k + q];
                    A[4 * k + p] = c * akp - s * akq;
                    A[4 * k + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; ++k) { /* A <- J^T A */
                    double apk = A[4 * p + k], aqk = A[4 * q + k];
                    A[4 * p + k] = c * apk - s * aqk;
                    A[4 * q + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 4; ++k) {
                    double ukp = U[4 * k + p], ukq = U[4 * k + q];
                    U[4 * k + p] = c * ukp - s * ukq;
                    U[4 * k + q] = s * ukp + c * ukq;
                }
            }
    }
    for (int i = 0; i < 4; ++i) lam[i] = A[5 * i];
    for (int i = 0; i < 3; ++i) /* sort ascending, permuting columns of U */
        for (int j = i + 1; j < 4; ++j)
            if (lam[j] < lam[i]) {
                double t = lam[i];
                lam[i] = lam[j];
                lam[j] = t;
                for (int k = 0; k < 4; ++k) {
                    t = U[4 * k + i];
                    U[4 * k + i] = U[4 * k + j];
                    U[4 * k + j] = t;
                }
            }
}

/* ---------------------------------------------------------------- substitution model */

typedef struct {
    int model, normalize, jc_closed;
    double R[16];      /* symmetric exchangeabilities, zero diagonal */
    double pi[4];
    double Q0[16], s;  /* un-normalised Q and its normaliser */
    double Q[16];
    double lam[4], m1[16], m2[16];
    int ntheta;        /* 0 / 1 / 6 */
} subst_t;

/* generate_script.py:799-812 (HKY), :855-868 (GTR); JC69 == all exchangeabilities 1, pi=1/4 */
static void subst_setup(subst_t *m, int model, int flags, const double *subst, const double *freqs) {
    memset(m, 0, sizeof *m);
    m->model = model;
    m->normalize = !(flags & ORACLE_NO_NORMQ);
    m->jc_closed = (model == ORACLE_JC69) && m->normalize;
    double *R = m->R;
    if (model == ORACLE_JC69) {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) R[4 * i + j] = (i == j) ? 0.0 : 1.0;
        for (int i = 0; i < 4; ++i) m->pi[i] = 0.25;
        m->ntheta = 0;
    } else if (model == ORACLE_HKY) {
        double k = subst[0];
        double r[16] = {0, 1, k, 1, 1, 0, 1, k, k, 1, 0, 1, 1, k, 1, 0};
        memcpy(R, r, sizeof r);
        memcpy(m->pi, freqs, 4 * sizeof(double));
        m->ntheta = 1;
    } else {
        const double *r = subst;
        double rr[16] = {0, r[0], r[1], r[2], r[0], 0, r[3], r[4], r[1], r[3], 0, r[5], r[2], r[4], r[5], 0};
        memcpy(R, rr, sizeof rr);
        memcpy(m->pi, freqs, 4 * sizeof(double));
        m->ntheta = 6;
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) {
        double row = 0;
        for (int j = 0; j < 4; ++j) {
            m->Q0[4 * i + j] = (i == j) ? 0.0 : R[4 * i + j] * m->pi[j];
            row += m->Q0[4 * i + j];
        }
        m->Q0[5 * i] = -row;
        s -= m->Q0[5 * i] * m->pi[i];
    }
    m->s = s;
    for (int i = 0; i < 16; ++i) m->Q[i] = m->normalize ? m->Q0[i] / s : m->Q0[i];
    /* A = D^1/2 Q D^-1/2 ; m1 = D^-1/2 U ; m2 = U^T D^1/2   (:814-825 / :870-881) */
    double A[16], U[16], sq[4];
    for (int i = 0; i < 4; ++i) sq[i] = sqrt(m->pi[i]);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) A[4 * i + j] = sq[i] * m->Q[4 * i + j] / sq[j];
    for (int i = 0; i < 4; ++i) /* symmetrise away rounding */
        for (int j = i + 1; j < 4; ++j) A[4 * i + j] = A[4 * j + i] = 0.5 * (A[4 * i + j] + A[4 * j + i]);
    eig_sym4(A, m->lam, U);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            m->m1[4 * i + j] = U[4 * i + j] / sq[i];
            m->m2[4 * i + j] = U[4 * j + i] * sq[j];
        }
}

static void subst_pmatrix(const subst_t *m, double tau, double *P) {
    if (m->jc_closed) { /* generate_script.py:765-766; expm1 form: 1/4 - 1/4 e^{-x} cancels when x is tiny */
        double e1 = expm1(-tau / 0.75);
        for (int i = 0; i < 16; ++i) P[i] = -0.25 * e1;
        for (int i = 0; i < 4; ++i) P[5 * i] = 1.0 + 0.75 * e1;
        return;
    }
    /* generate_script.py:824-829.  Short branches (max|lambda| tau < 1): the mathematically identical
     * P = I + m1 diag(expm1(lambda tau)) m2; the plain form computes the O(tau) off-diagonal entries by
     * cancelling m1 m2 against the identity (absolute error 1e-16 on entries of size tau), which no
     * two implementations round alike -- see DESIGN.md "reference quirks". */
    double T[16], lmax = 0;
    for (int j = 0; j < 4; ++j)
        if (fabs(m->lam[j]) > lmax) lmax = fabs(m->lam[j]);
    int small = lmax * tau < 1.0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            T[4 * i + j] = m->m1[4 * i + j] * (small ? expm1(m->lam[j] * tau) : exp(m->lam[j] * tau));
    mat4_mul(T, m->m2, P);
    if (small)
        for (int i = 0; i < 4; ++i) P[5 * i] += 1.0;
}

/* dQ/dtheta_k (unconstrained partials).  k < ntheta: exchangeability parameter;
 * k = ntheta .. ntheta+3: frequency pi_{k-ntheta}. */
static void subst_dQ(const subst_t *m, int k, double *dQ) {
    double dQ0[16] = {0}, ds = 0;
    const double *pi = m->pi, *R = m->R;
    static const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
    if (k < m->ntheta) {
        int np = 1, A_[2], B_[2];
        if (m->model == ORACLE_HKY) {
            np = 2;
            A_[0] = 0; B_[0] = 2; A_[1] = 1; B_[1] = 3;
        } else {
            A_[0] = pa[k]; B_[0] = pb[k];
        }
        for (int e = 0; e < np; ++e) {
            int a = A_[e], b = B_[e];
            dQ0[4 * a + b] += pi[b];
            dQ0[4 * b + a] += pi[a];
            dQ0[5 * a] -= pi[b];
            dQ0[5 * b] -= pi[a];
            ds += 2.0 * pi[a] * pi[b];
        }
    } else {
        int f = k - m->ntheta;
        for (int i = 0; i < 4; ++i)
            if (i != f) {
                dQ0[4 * i + f] += R[4 * i + f];
                dQ0[5 * i] -= R[4 * i + f];
            }
        ds = -2.0 * m->Q0[5 * f];
    }
    for (int i = 0; i < 16; ++i)
        dQ[i] = m->normalize ? dQ0[i] / m->s - m->Q0[i] * ds / (m->s * m->s) : dQ0[i];
}

/* ---- Van Loan: expm([[A,E],[0,A]]) = [[e^A, D],[0,e^A]],  D = d/de e^{A+eE}|0 ---------- */
static void mat8_mul(const double *A, const double *B, double *Cm) {
    double T[64];
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            double s = 0;
            for (int k = 0; k < 8; ++k) s += A[8 * i + k] * B[8 * k + j];
            T[8 * i + j] = s;
        }
    memcpy(Cm, T, sizeof T);
}

static void expm8(const double *Min, double *E) {
    double M[64], nrm = 0;
    memcpy(M, Min, sizeof M);
    for (int i = 0; i < 8; ++i) {
        double r = 0;
        for (int j = 0; j < 8; ++j) r += fabs(M[8 * i + j]);
        if (r > nrm) nrm = r;
    }
    int sq = 0;
    while (nrm > 0.25) { nrm *= 0.5; ++sq; }
    double sc = ldexp(1.0, -sq);
    for (int i = 0; i < 64; ++i) M[i] *= sc;
    double term[64];
    for (int i = 0; i < 64; ++i) E[i] = term[i] = (i % 9 == 0) ? 1.0 : 0.0;
    for (int k = 1; k <= 24; ++k) {
        mat8_mul(term, M, term);
        for (int i = 0; i < 64; ++i) { term[i] /= k; E[i] += term[i]; }
    }
    for (int i = 0; i < sq; ++i) mat8_mul(E, E, E);
}

static void dP_vanloan(const subst_t *m, const double *dQ, double tau, double *dP) {
    double M[64] = {0}, E[64];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            M[8 * i + j] = m->Q[4 * i + j] * tau;
            M[8 * (i + 4) + j + 4] = m->Q[4 * i + j] * tau;
            M[8 * i + j + 4] = dQ[4 * i + j] * tau;
        }
    expm8(M, E);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) dP[4 * i + j] = E[8 * i + j + 4];
}

/* eigen route: dP = m1 (F o (m2 dQ m1)) m2, F_ij = (e^{li t}-e^{lj t})/(li-lj), F_ii = t e^{li t} */
static void dP_eigen(const subst_t *m, const double *dQ, double tau, double *dP) {
    double X[16], T[16];
    mat4_mul(m->m2, dQ, T);
    mat4_mul(T, m->m1, X);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            /* factored around the larger exponential: expm1 never overflows on long branches */
            double x = fabs(m->lam[i] - m->lam[j]) * tau;
            double hi = exp((m->lam[i] > m->lam[j] ? m->lam[i] : m->lam[j]) * tau);
            double f = tau * hi * (x < 1e-9 ? 1.0 - 0.5 * x : -expm1(-x) / x);
            X[4 * i + j] *= f;
        }
    mat4_mul(m->m1, X, T);
    mat4_mul(T, m->m2, dP);
}

/* ---------------------------------------------------------------- tree bookkeeping */

typedef struct {
    int S, L, C, nnode, rooted, bcount, root, uc1, uc2;
    const int32_t *peel;
} tree_t;

static int tree_setup(tree_t *t, int S, int L, int C, const int32_t *peel, int flags) {
    t->S = S; t->L = L; t->C = C; t->nnode = 2 * S - 1; t->peel = peel;
    t->rooted = (flags & ORACLE_ROOTED) != 0;
    t->bcount = t->rooted ? 2 * S - 2 : 2 * S - 3;
    t->root = peel[3 * (S - 2) + 2] - 1;
    t->uc1 = peel[3 * (S - 2) + 0] - 1;
    t->uc2 = peel[3 * (S - 2) + 1] - 1;
    if (S < 2 || t->root != 2 * S - 2) return 1;
    if (!t->rooted && (S < 3 || t->uc2 != 2 * S - 3)) return 2;
    return 0;
}

static inline void tip_partial(uint8_t mask, double *p) {
    for (int s = 0; s < 4; ++s) p[s] = (mask >> s) & 1 ? 1.0 : 0.0;
}

static inline void matvec(const double *P, const double *x, double *y) {
    for (int i = 0; i < 4; ++i)
        y[i] = P[4 * i] * x[0] + P[4 * i + 1] * x[1] + P[4 * i + 2] * x[2] + P[4 * i + 3] * x[3];
}
static inline void matTvec(const double *P, const double *x, double *y) {
    for (int j = 0; j < 4; ++j)
        y[j] = P[j] * x[0] + P[4 + j] * x[1] + P[8 + j] * x[2] + P[12 + j] * x[3];
}
static inline double max4(const double *x) {
    double m = x[0];
    for (int i = 1; i < 4; ++i) if (x[i] > m) m = x[i];
    return m;
}

/* does node k (0-based, non-root) carry a branch? (unrooted: uc2 does not) */
static inline int has_branch(const tree_t *t, int k) { return k < t->bcount; }

/* post-order for one pattern and one category.  p, msg: [nnode][4]; returns log of the
 * accumulated scale of the root partial (0 when rescaling is off). */
static double postorder(const tree_t *t, const uint8_t *tipmask, int l, const double *Pc,
                        int rescale, double *p, double *msg) {
    int S = t->S;
    double logscale = 0;
    for (int k = 0; k < S; ++k) tip_partial(tipmask[(size_t)k * t->L + l], p + 4 * k);
    for (int n = 0; n < S - 1; ++n) {
        int a = t->peel[3 * n] - 1, b = t->peel[3 * n + 1] - 1, par = t->peel[3 * n + 2] - 1;
        matvec(Pc + 16 * a, p + 4 * a, msg + 4 * a);
        if (has_branch(t, b)) matvec(Pc + 16 * b, p + 4 * b, msg + 4 * b);
        else memcpy(msg + 4 * b, p + 4 * b, 4 * sizeof(double)); /* generate_script.py:1034 */
        for (int s = 0; s < 4; ++s) p[4 * par + s] = msg[4 * a + s] * msg[4 * b + s];
        if (rescale) {
            double mx = max4(p + 4 * par);
            if (mx > 0) {
                for (int s = 0; s < 4; ++s) p[4 * par + s] /= mx;
                logscale += log(mx);
            }
        }
    }
    return logscale;
}

typedef struct {
    double logp;
    double *gtau;   /* [nnode][C]  d logL / d(t_b r_c) */
    double *G;      /* [nnode][C][16] sum_l w*omega*A(x)p(y)/den, or NULL */
    double gpi_root[4];
    double *gps;    /* [C] */
} accum_t;

static void pattern_eval(const tree_t *t, const uint8_t *tipmask, double w, int l,
                         const subst_t *sm, const double *Pall, const double *ps, int rescale,
                         int want_grad, double *work, accum_t *acc, double *site_logl) {
    const int nn = t->nnode, C = t->C;
    double *p = work, *msg = p + (size_t)4 * nn * C, *q = msg + (size_t)4 * nn * C;
    double lc[64], lsc[64];
    double mxl = -INFINITY;
    for (int c = 0; c < C; ++c) {
        double *pc = p + (size_t)4 * nn * c, *mc = msg + (size_t)4 * nn * c;
        lsc[c] = postorder(t, tipmask, l, Pall + (size_t)16 * nn * c, rescale, pc, mc);
        double r = 0;
        for (int s = 0; s < 4; ++s) r += pc[4 * t->root + s] * sm->pi[s]; /* :1007 */
        lc[c] = (ps[c] * r > 0) ? log(ps[c] * r) + lsc[c] : -INFINITY;
        if (lc[c] > mxl) mxl = lc[c];
    }
    double sum = 0;
    for (int c = 0; c < C; ++c) sum += exp(lc[c] - mxl);
    double sitel = mxl + log(sum);
    if (!rescale) { /* reference arithmetic: log(sum(probs)) with no shifting (:1010) */
        double tot = 0;
        for (int c = 0; c < C; ++c) {
            double r = 0;
            const double *pc = p + (size_t)4 * nn * c;
            for (int s = 0; s < 4; ++s) r += pc[4 * t->root + s] * sm->pi[s];
            tot += ps[c] * r;
        }
        sitel = log(tot);
    }
    if (site_logl) site_logl[l] = sitel;
    acc->logp += w * sitel;
    if (!want_grad) return;

    for (int c = 0; c < C; ++c) {
        double omega = exp(lc[c] - sitel); /* ps_c L_c / L */
        double *pc = p + (size_t)4 * nn * c, *mc = msg + (size_t)4 * nn * c, *qc = q + (size_t)4 * nn * c;
        const double *Pc = Pall + (size_t)16 * nn * c;
        double rootdot = 0;
        for (int s = 0; s < 4; ++s) rootdot += pc[4 * t->root + s] * sm->pi[s];
        if (rootdot > 0) {
            for (int s = 0; s < 4; ++s) acc->gpi_root[s] += w * omega * pc[4 * t->root + s] / rootdot;
            if (ps[c] > 0) acc->gps[c] += w * omega / ps[c];
        }
        if (!(omega > 0)) continue;
        for (int s = 0; s < 4; ++s) qc[4 * t->root + s] = sm->pi[s]; /* eigen.j2:144 */
        for (int n = t->S - 2; n >= 0; --n) { /* reverse post-order == a pre-order */
            int ch[2] = {t->peel[3 * n] - 1, t->peel[3 * n + 1] - 1}, par = t->peel[3 * n + 2] - 1;
            for (int e = 0; e < 2; ++e) {
                int b = ch[e], sib = ch[1 - e];
                double A[4];
                for (int s = 0; s < 4; ++s) A[s] = qc[4 * par + s] * mc[4 * sib + s]; /* eq (7) */
                if (has_branch(t, b)) matTvec(Pc + 16 * b, A, qc + 4 * b);
                else memcpy(qc + 4 * b, A, sizeof A);
                if (rescale) {
                    double mx = max4(qc + 4 * b);
                    if (mx > 0) {
                        for (int s = 0; s < 4; ++s) { qc[4 * b + s] /= mx; A[s] /= mx; }
                    }
                }
                if (!has_branch(t, b)) continue;
                /* eq (9): ((Q^T q_b) . p_b) / (q_b . p_b) */
                double Qtq[4], num = 0, den = 0;
                matTvec(sm->Q, qc + 4 * b, Qtq);
                for (int s = 0; s < 4; ++s) { num += Qtq[s] * pc[4 * b + s]; den += qc[4 * b + s] * pc[4 * b + s]; }
                if (!(den > 0)) continue;
                double f = w * omega / den;
                acc->gtau[(size_t)b * C + c] += f * num;
                if (acc->G) {
                    double *G = acc->G + ((size_t)b * C + c) * 16;
                    for (int x = 0; x < 4; ++x)
                        for (int y = 0; y < 4; ++y) G[4 * x + y] += f * A[x] * pc[4 * b + y];
                }
            }
        }
    }
}

static double *all_pmatrices(const tree_t *t, const subst_t *sm, const double *blens, const double *rs) {
    const int nn = t->nnode, C = t->C;
    double *P = (double *)calloc((size_t)16 * nn * C, sizeof(double));
    for (int c = 0; c < C; ++c)
        for (int b = 0; b < t->bcount; ++b)
            subst_pmatrix(sm, blens[b] * rs[c], P + ((size_t)nn * c + b) * 16);
    return P;
}

int oracle_loglik_grad(int S, int L, int C, const int32_t *peel, const uint8_t *tipmask,
                       const double *weights, int model, int flags, const double *blens,
                       const double *subst, const double *freqs, const double *rs,
                       const double *ps, int want_grad, int nthreads, double *logp,
                       double *grad_blens, double *grad_subst, double *grad_freqs,
                       double *grad_rs, double *grad_ps) {
    tree_t t;
    int rc = tree_setup(&t, S, L, C, peel, flags);
    if (rc || C > 64) return 10 + rc;
    subst_t sm;
    subst_setup(&sm, model, flags, subst, freqs);
    double *Pall = all_pmatrices(&t, &sm, blens, rs);
    const int nn = t.nnode, rescale = (flags & ORACLE_RESCALE) != 0;
    const int need_G = want_grad && model != ORACLE_JC69 && (grad_subst || grad_freqs);
#ifdef _OPENMP
    int nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#else
    int nt = 1;
    (void)nthreads;
#endif
    if (nt > L) nt = L > 0 ? L : 1;
    accum_t *accs = (accum_t *)calloc(nt, sizeof(accum_t));
    for (int i = 0; i < nt; ++i) {
        accs[i].gtau = (double *)calloc((size_t)nn * C, sizeof(double));
        accs[i].gps = (double *)calloc(C, sizeof(double));
        accs[i].G = need_G ? (double *)calloc((size_t)nn * C * 16, sizeof(double)) : NULL;
    }
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        double *work = (double *)malloc((size_t)3 * 4 * nn * C * sizeof(double));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int l = 0; l < L; ++l)
            pattern_eval(&t, tipmask, weights ? weights[l] : 1.0, l, &sm, Pall, ps, rescale, want_grad,
                         work, &accs[tid], NULL);
        free(work);
    }
    /* deterministic reduction over threads */
    accum_t *a0 = &accs[0];
    for (int i = 1; i < nt; ++i) {
        a0->logp += accs[i].logp;
        for (size_t k = 0; k < (size_t)nn * C; ++k) a0->gtau[k] += accs[i].gtau[k];
        for (int c = 0; c < C; ++c) a0->gps[c] += accs[i].gps[c];
        for (int s = 0; s < 4; ++s) a0->gpi_root[s] += accs[i].gpi_root[s];
        if (need_G)
            for (size_t k = 0; k < (size_t)nn * C * 16; ++k) a0->G[k] += accs[i].G[k];
    }
    if (logp) *logp = a0->logp;
    if (want_grad) {
        if (grad_blens)
            for (int b = 0; b < t.bcount; ++b) {
                double g = 0;
                for (int c = 0; c < C; ++c) g += rs[c] * a0->gtau[(size_t)b * C + c];
                grad_blens[b] = (flags & ORACLE_QUIRK_TIMES) ? blens[b] * g : g; /* eigen.j2:165 */
            }
        if (grad_rs)
            for (int c = 0; c < C; ++c) {
                double g = 0;
                for (int b = 0; b < t.bcount; ++b) g += blens[b] * a0->gtau[(size_t)b * C + c];
                grad_rs[c] = g;
            }
        if (grad_ps)
            for (int c = 0; c < C; ++c) grad_ps[c] = a0->gps[c];
        int ntheta = sm.ntheta;
        if (grad_freqs)
            for (int s = 0; s < 4; ++s) grad_freqs[s] = a0->gpi_root[s];
        if (grad_subst)
            for (int k = 0; k < ntheta; ++k) grad_subst[k] = 0;
        if (need_G) {
            for (int k = 0; k < ntheta + 4; ++k) {
                double *dst = (k < ntheta) ? (grad_subst ? grad_subst + k : NULL)
                                           : (grad_freqs ? grad_freqs + (k - ntheta) : NULL);
                if (!dst) continue;
                double dQ[16], dP[16], tot = 0;
                subst_dQ(&sm, k, dQ);
                for (int b = 0; b < t.bcount; ++b)
                    for (int c = 0; c < C; ++c) {
                        double tau = blens[b] * rs[c];
                        if (flags & ORACLE_DP_EIGEN) dP_eigen(&sm, dQ, tau, dP);
                        else dP_vanloan(&sm, dQ, tau, dP);
                        const double *G = a0->G + ((size_t)b * C + c) * 16;
                        for (int i = 0; i < 16; ++i) tot += G[i] * dP[i];
                    }
                *dst += tot;
            }
        }
    }
    for (int i = 0; i < nt; ++i) { free(accs[i].gtau); free(accs[i].gps); free(accs[i].G); }
    free(accs);
    free(Pall);
    return 0;
}

int oracle_site_loglik(int S, int L, int C, const int32_t *peel, const uint8_t *tipmask,
                       int model, int flags, const double *blens, const double *subst,
                       const double *freqs, const double *rs, const double *ps,
                       double *site_logl) {
    tree_t t;
    int rc = tree_setup(&t, S, L, C, peel, flags);
    if (rc || C > 64) return 10 + rc;
    subst_t sm;
    subst_setup(&sm, model, flags, subst, freqs);
    double *Pall = all_pmatrices(&t, &sm, blens, rs);
    accum_t acc;
    memset(&acc, 0, sizeof acc);
    double *work = (double *)malloc((size_t)3 * 4 * t.nnode * C * sizeof(double));
    for (int l = 0; l < L; ++l)
        pattern_eval(&t, tipmask, 1.0, l, &sm, Pall, ps, (flags & ORACLE_RESCALE) != 0, 0, work, &acc,
                     site_logl);
    free(work);
    free(Pall);
    return 0;
}

int oracle_pmatrix(int model, int flags, const double *subst, const double *freqs, double tau,
                   double *P) {
    subst_t sm;
    subst_setup(&sm, model, flags, subst, freqs);
    subst_pmatrix(&sm, tau, P);
    return 0;
}

double oracle_pq_invariant(int S, int L, int C, const int32_t *peel, const uint8_t *tipmask,
                           int model, int flags, const double *blens, const double *subst,
                           const double *freqs, const double *rs, const double *ps, int l, int c) {
    (void)ps;
    tree_t t;
    if (tree_setup(&t, S, L, C, peel, flags)) return NAN;
    subst_t sm;
    subst_setup(&sm, model, flags, subst, freqs);
    double *Pall = all_pmatrices(&t, &sm, blens, rs);
    const int nn = t.nnode;
    const double *Pc = Pall + (size_t)16 * nn * c;
    double *p = (double *)calloc((size_t)12 * nn, sizeof(double)), *msg = p + 4 * nn, *q = msg + 4 * nn;
    postorder(&t, tipmask, l, Pc, 0, p, msg);
    double Lk = 0;
    for (int s = 0; s < 4; ++s) { q[4 * t.root + s] = sm.pi[s]; Lk += sm.pi[s] * p[4 * t.root + s]; }
    double worst = 0;
    for (int n = S - 2; n >= 0; --n) {
        int ch[2] = {peel[3 * n] - 1, peel[3 * n + 1] - 1}, par = peel[3 * n + 2] - 1;
        for (int e = 0; e < 2; ++e) {
            int b = ch[e], sib = ch[1 - e];
            double A[4], d = 0;
            for (int s = 0; s < 4; ++s) A[s] = q[4 * par + s] * msg[4 * sib + s];
            if (has_branch(&t, b)) matTvec(Pc + 16 * b, A, q + 4 * b); /* pruner/tree.cpp:239 */
            else memcpy(q + 4 * b, A, sizeof A);
            for (int s = 0; s < 4; ++s) d += q[4 * b + s] * p[4 * b + s];
            double e_ = fabs(d - Lk) / Lk;
            if (e_ > worst) worst = e_;
        }
    }
    free(p);
    free(Pall);
    return worst;
}
