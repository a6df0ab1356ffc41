// subst.hpp -- host-side substitution-model algebra: everything that is O(1) per parameter draw.
//
// Builds what the device kernels consume for one draw: the normalised rate matrix Q, its
// symmetric eigen-system (lambda, m1, m2) exactly as the generated Stan functions do
// (phylostan/generate_script.py:799-825 HKY, :855-881 GTR; JC69 is the all-ones special case of
// :755-780), and for every substitution parameter theta the eigen-basis matrix
// X_theta = m2 * dQ/dtheta * m1 used by the contraction kernel:
//     dP/dtheta = m1 (F o X_theta) m2,  F_ij = (e^{l_i tau} - e^{l_j tau}) / (l_i - l_j).
#pragma once

namespace phylo {

enum Model { JC69 = 0, HKY = 1, GTR = 2 };

inline int n_subst(int model) { return model == GTR ? 6 : (model == HKY ? 1 : 0); }

struct Derived {
    double pi[4];
    double lam[4];
    double m1[16], m2[16];  // row-major; P(tau) = m1 diag(exp(lam tau)) m2
    double Q[16];
    double X[10][16];       // [theta] : n_subst exchangeability parameters, then 4 frequencies
    int ntheta;             // n_subst + 4 (JC69: 0 -- frequencies are fixed)
};

// Returns false when a parameter is non-finite or out of domain (freqs <= 0, rates < 0, ...).
bool derive(int model, bool normalize, const double* subst, const double* freqs, Derived& out);

// Symmetric 4x4 eigen-solver (cyclic Jacobi); eigenvalues ascending, eigenvectors in columns.
void eigh4(const double* A, double* lam, double* U);

}  // namespace phylo
