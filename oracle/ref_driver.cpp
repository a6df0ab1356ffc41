// Driver for the reference's own C++ likelihood (eigen/eigen.j2 rendered to eigen.hpp by oracle/build_ref.py):
// reads sets of branch lengths from stdin ("n t_0 ... t_{n-1}" per line), prints log_P and the gradient vector the
// reference returns (eigen.j2:160-167: entry i is times[i] * dlogP/dtimes[i]) with 17 significant digits.
// TEST INFRASTRUCTURE ONLY.  Compiled against oracle/mini_eigen (Eigen is not in this image).
#include <cstdio>
#include <iostream>
#include <vector>

#include "eigen.hpp"

int main() {
    int n;
    while (std::scanf("%d", &n) == 1) {
        std::vector<double> t;
        t.reserve(n + 1);  // eigen.j2:165 reads times[i] for the root as well (one past the end): keep that slot defined
        for (int i = 0; i < n; ++i) {
            double x;
            if (std::scanf("%lf", &x) != 1) return 2;
            t.push_back(x);
        }
        t.data()[n] = 0.0;
        value_grad r = vbsky_loglik(t);
        std::printf("%.17g", r.log_P);
        for (int i = 0; i < n; ++i) std::printf(" %.17g", r.grad(i));
        std::printf("\n");
    }
    return 0;
}
