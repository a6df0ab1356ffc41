// subst.cpp -- see subst.hpp.  Pure host C++.
#include "subst.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace phylo {

namespace {

struct M4 {
    double v[16];
    double& operator()(int i, int j) { return v[4 * i + j]; }
    double operator()(int i, int j) const { return v[4 * i + j]; }
};

M4 mul(const M4& A, const M4& B) {
    M4 C;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A(i, k) * B(k, j);
            C(i, j) = s;
        }
    return C;
}

// exchangeability pairs in GTR order AC, AG, AT, CG, CT, GT (generate_script.py:855-858)
constexpr int kPairA[6] = {0, 0, 0, 1, 1, 2};
constexpr int kPairB[6] = {1, 2, 3, 2, 3, 3};

}  // namespace

void eigh4(const double* Ain, double* lam, double* U) {
    double A[4][4], V[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            A[i][j] = Ain[4 * i + j];
            V[i][j] = i == j;
        }
    for (int sweep = 0; sweep < 50; ++sweep) {
        double off = 0, diag = 0;
        for (int i = 0; i < 4; ++i) {
            diag += A[i][i] * A[i][i];
            for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j];
        }
        if (off <= 1e-34 * diag || off == 0.0) break;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 4; ++q) {
                if (A[p][q] == 0.0) continue;
                // symmetric Schur rotation annihilating A[p][q]
                double tau = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                double t = std::copysign(1.0, tau) / (std::fabs(tau) + std::hypot(1.0, tau));
                double c = 1.0 / std::hypot(1.0, t), s = t * c;
                for (int k = 0; k < 4; ++k) {
                    double x = A[k][p], y = A[k][q];
                    A[k][p] = c * x - s * y;
                    A[k][q] = s * x + c * y;
                }
                for (int k = 0; k < 4; ++k) {
                    double x = A[p][k], y = A[q][k];
                    A[p][k] = c * x - s * y;
                    A[q][k] = s * x + c * y;
                }
                for (int k = 0; k < 4; ++k) {
                    double x = V[k][p], y = V[k][q];
                    V[k][p] = c * x - s * y;
                    V[k][q] = s * x + c * y;
                }
            }
    }
    int idx[4] = {0, 1, 2, 3};
    std::sort(idx, idx + 4, [&](int a, int b) { return A[a][a] < A[b][b]; });
    for (int j = 0; j < 4; ++j) {
        lam[j] = A[idx[j]][idx[j]];
        for (int i = 0; i < 4; ++i) U[4 * i + j] = V[i][idx[j]];
    }
}

bool derive(int model, bool normalize, const double* subst, const double* freqs, Derived& d) {
    std::memset(&d, 0, sizeof d);
    M4 R{};
    const int ns = n_subst(model);
    if (model == JC69) {
        for (int i = 0; i < 4; ++i) {
            d.pi[i] = 0.25;
            for (int j = 0; j < 4; ++j) R(i, j) = i != j;
        }
    } else {
        if (!subst || !freqs) return false;
        for (int i = 0; i < 4; ++i) {
            if (!(freqs[i] > 0.0) || !std::isfinite(freqs[i])) return false;
            d.pi[i] = freqs[i];
        }
        for (int k = 0; k < ns; ++k)
            if (!(subst[k] >= 0.0) || !std::isfinite(subst[k])) return false;
        if (model == HKY) {
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) R(i, j) = i != j;
            R(0, 2) = R(2, 0) = R(1, 3) = R(3, 1) = subst[0];  // transitions A<->G, C<->T
        } else {
            for (int k = 0; k < 6; ++k) R(kPairA[k], kPairB[k]) = R(kPairB[k], kPairA[k]) = subst[k];
        }
    }
    // Q0 = R diag(pi) with zero row sums; s = -sum_i Q0_ii pi_i  (generate_script.py:862-868)
    M4 Q0{};
    double s = 0;
    for (int i = 0; i < 4; ++i) {
        double row = 0;
        for (int j = 0; j < 4; ++j)
            if (i != j) {
                Q0(i, j) = R(i, j) * d.pi[j];
                row += Q0(i, j);
            }
        Q0(i, i) = -row;
        s += row * d.pi[i];
    }
    if (!(s > 0.0)) return false;
    const double inv_s = normalize ? 1.0 / s : 1.0;
    M4 Q;
    for (int i = 0; i < 16; ++i) Q.v[i] = Q0.v[i] * inv_s;
    std::memcpy(d.Q, Q.v, sizeof d.Q);

    // symmetrise with D = diag(pi): A = D^{1/2} Q D^{-1/2}; m1 = D^{-1/2} U, m2 = U^T D^{1/2}
    double rt[4], A[16], U[16];
    for (int i = 0; i < 4; ++i) rt[i] = std::sqrt(d.pi[i]);
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j) {
            double aij = rt[i] * Q(i, j) / rt[j], aji = rt[j] * Q(j, i) / rt[i];
            A[4 * i + j] = A[4 * j + i] = 0.5 * (aij + aji);
        }
    eigh4(A, d.lam, U);
    M4 m1, m2;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            m1(i, j) = U[4 * i + j] / rt[i];
            m2(i, j) = U[4 * j + i] * rt[j];
        }
    std::memcpy(d.m1, m1.v, sizeof d.m1);
    std::memcpy(d.m2, m2.v, sizeof d.m2);

    // X_theta = m2 dQ/dtheta m1 with dQ = dQ0/s - Q0 ds/s^2 (normalised) or dQ0
    d.ntheta = model == JC69 ? 0 : ns + 4;
    for (int k = 0; k < d.ntheta; ++k) {
        M4 dQ0{};
        double ds = 0;
        auto bump_pair = [&](int a, int b) {
            dQ0(a, b) += d.pi[b];
            dQ0(b, a) += d.pi[a];
            dQ0(a, a) -= d.pi[b];
            dQ0(b, b) -= d.pi[a];
            ds += 2.0 * d.pi[a] * d.pi[b];
        };
        if (k < ns) {
            if (model == HKY) { bump_pair(0, 2); bump_pair(1, 3); }
            else bump_pair(kPairA[k], kPairB[k]);
        } else {
            const int f = k - ns;  // d/dpi_f: column f of the off-diagonal part
            for (int i = 0; i < 4; ++i)
                if (i != f) {
                    dQ0(i, f) += R(i, f);
                    dQ0(i, i) -= R(i, f);
                    ds += 2.0 * R(i, f) * d.pi[i];
                }
        }
        M4 dQ;
        for (int i = 0; i < 16; ++i)
            dQ.v[i] = normalize ? (dQ0.v[i] - Q.v[i] * ds) * inv_s : dQ0.v[i];
        M4 X = mul(mul(m2, dQ), m1);
        std::memcpy(d.X[k], X.v, sizeof d.X[k]);
    }
    return true;
}

}  // namespace phylo
