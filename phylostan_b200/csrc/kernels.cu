// kernels.cu -- see kernels.cuh.  sm_100a only.
#include "kernels.cuh"

namespace phylo {

namespace {

// threads per CTA allowed for K patterns per thread (register budget: 128 / 128 / 255)
constexpr int max_threads(int K) { return K == 1 ? 512 : (K == 2 ? 256 : 128); }
constexpr int min_blocks(int K) { return K == 1 ? 1 : 2; }

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ double pow2_64k(int k) {  // 2^(64 k), -15 <= k <= 15
    return __hiloint2double((1023 + 64 * k) << 20, 0);
}

__device__ __forceinline__ void load_mat(const double* __restrict__ M, double (&m)[16]) {
    const double2* q = reinterpret_cast<const double2*>(M);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double2 v = __ldg(q + i);
        m[2 * i] = v.x;
        m[2 * i + 1] = v.y;
    }
}

// y = M x  (row-major)
__device__ __forceinline__ void matvec(const double (&m)[16], const double (&x)[4], double (&y)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        y[i] = fma(m[4 * i + 3], x[3], fma(m[4 * i + 2], x[2], fma(m[4 * i + 1], x[1], m[4 * i] * x[0])));
}

// y = M^T x
__device__ __forceinline__ void matTvec(const double (&m)[16], const double (&x)[4], double (&y)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
        y[j] = fma(m[12 + j], x[3], fma(m[8 + j], x[2], fma(m[4 + j], x[1], m[j] * x[0])));
}

__device__ __forceinline__ void tip_vec(unsigned code, double (&p)[4]) {
#pragma unroll
    for (int s = 0; s < 4; ++s) p[s] = ((code >> s) & 1u) ? 1.0 : 0.0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum 16 per-lane values over the 32 lanes of a warp with a recursive-halving exchange
// (8+4+2+1+1 = 16 shuffles instead of 80), then one 128-byte RED per warp: even lane 2i adds
// entry i to dst[i].
__device__ __forceinline__ void warp_reduce16_atomic(const double (&v)[16], double* __restrict__ dst, int lane) {
    double a8[8], a4[4], a2[2], a1;
    bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double send = hi ? v[i] : v[i + 8], keep = hi ? v[i + 8] : v[i];
        a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double send = hi ? a8[i] : a8[i + 4], keep = hi ? a8[i + 4] : a8[i];
        a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        double send = hi ? a4[i] : a4[i + 2], keep = hi ? a4[i + 2] : a4[i];
        a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    hi = lane & 2;
    {
        double send = hi ? a2[0] : a2[1], keep = hi ? a2[1] : a2[0];
        a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    if (!(lane & 1)) atomicAdd(dst + ((lane >> 1) & 15), a1);
}

// ------------------------------------------------------------------------------------------
// K1: transition matrices
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128) pmat_kernel(const double* __restrict__ params, ParamLayout lay, int bcount,
                                                   int jc_closed, double* __restrict__ P, int total) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k = idx % lay.nn, c = (idx / lay.nn) % lay.C, d = idx / (lay.nn * lay.C);
    const double* prm = params + (size_t)d * lay.stride;
    double* out = P + (size_t)idx * 16;
    double m[16];
    if (k >= bcount) {  // root, or the unrooted second root child: no branch
#pragma unroll
        for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0;
    } else {
        const double tau = prm[lay.off_t + k] * prm[lay.off_rs + c];
        if (jc_closed) {  // generate_script.py:765-766
            const double e = exp(-tau / 0.75);
#pragma unroll
            for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 0.25 + 0.75 * e : 0.25 - 0.25 * e;
        } else {  // generate_script.py:824-829
            double ex[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) ex[j] = exp(prm[lay.off_lam + j] * tau);
            const double* m1 = prm + lay.off_m1;
            const double* m2 = prm + lay.off_m2;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double s = 0.0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) s += (m1[4 * i + q] * ex[q]) * m2[4 * q + j];
                    m[4 * i + j] = s;
                }
        }
    }
    double2* o2 = reinterpret_cast<double2*>(out);
#pragma unroll
    for (int i = 0; i < 8; ++i) o2[i] = make_double2(m[2 * i], m[2 * i + 1]);
}

// ------------------------------------------------------------------------------------------
// K2/K3: fused depth-first post-order + pre-order sweep
// ------------------------------------------------------------------------------------------

template <int K, bool GRAD>
__global__ void __launch_bounds__(max_threads(K), min_blocks(K)) sweep_kernel(const SweepArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = a.C;
    const int c = warp % C, pb = warp / C;
    double2* st = reinterpret_cast<double2*>(smem_raw);                    // [D][K][2][NT]
    double* ex_l = reinterpret_cast<double*>(st + (size_t)a.D * K * 2 * NT);  // [K][NT]
    int* ex_e = reinterpret_cast<int*>(ex_l + K * NT);                     // [K][NT]
    const int tpat = (NT / (32 * C)) * 32 * K;

    double2* sc = a.scratch + (size_t)blockIdx.x * a.scratch_stride;
    int8_t* dl = a.dscr + (size_t)blockIdx.x * a.dscr_stride;

#define ST(slot, j, h) st[(((slot)*K + (j)) * 2 + (h)) * NT + tid]
#define SC(row, j, h) sc[(((size_t)(row)*K + (j)) * 2 + (h)) * NT + tid]
#define DL(row, j) dl[((size_t)(row)*K + (j)) * NT + tid]

    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int d = item / a.ntiles, tile = item - d * a.ntiles;
        const double* prm = a.params + (size_t)d * a.lay.stride;
        const double* Pd = a.P + ((size_t)d * C + c) * a.nn * 16;
        const int pat0 = tile * tpat + pb * 32 * K + lane;  // pattern of sub-index j: pat0 + 32 j

        // -------------------------------------------------------------- post-order
        int etot[K];
#pragma unroll
        for (int j = 0; j < K; ++j) etot[j] = 0;
        int so_last = 0;
        for (int i = 0; i < a.nsteps; ++i) {
            const int4 s0 = __ldg(reinterpret_cast<const int4*>(a.post + i));
            const int4 s1 = __ldg(reinterpret_cast<const int4*>(a.post + i) + 1);
            const int na = s0.x, nb = s0.y, sa = s0.z, sb = s0.w, so = s1.x;
            so_last = so;
            double M[16], ma[K][4];
            load_mat(Pd + (size_t)na * 16, M);
#pragma unroll
            for (int j = 0; j < K; ++j) {
                double p[4];
                if (sa < 0) {
                    tip_vec(a.tips[(size_t)na * a.Lpad + pat0 + 32 * j], p);
                } else {
                    double2 u = ST(sa, j, 0), v = ST(sa, j, 1);
                    p[0] = u.x; p[1] = u.y; p[2] = v.x; p[3] = v.y;
                }
                matvec(M, p, ma[j]);
            }
            load_mat(Pd + (size_t)nb * 16, M);
#pragma unroll
            for (int j = 0; j < K; ++j) {
                double p[4], mb[4];
                if (sb < 0) {
                    tip_vec(a.tips[(size_t)nb * a.Lpad + pat0 + 32 * j], p);
                } else {
                    double2 u = ST(sb, j, 0), v = ST(sb, j, 1);
                    p[0] = u.x; p[1] = u.y; p[2] = v.x; p[3] = v.y;
                }
                matvec(M, p, mb);
#pragma unroll
                for (int s = 0; s < 4; ++s) p[s] = ma[j][s] * mb[s];
                // per-(pattern,category) rescaling by exact powers of 2^64
                const double mx = fmax(fmax(p[0], p[1]), fmax(p[2], p[3]));
                int kexp = 0;
                if (mx < 2.938735877055719e-39 /* 2^-128 */ && mx > 0.0) {
                    const int e = ((__double2hiint(mx) >> 20) & 0x7ff) - 1023;  // floor(log2 mx), -1023 if subnormal
                    kexp = min((-e) >> 6, 15);
                    const double f = pow2_64k(kexp);
#pragma unroll
                    for (int s = 0; s < 4; ++s) p[s] *= f;
                    etot[j] += kexp;
                }
                ST(so, j, 0) = make_double2(p[0], p[1]);
                ST(so, j, 1) = make_double2(p[2], p[3]);
                if (GRAD) {
                    SC(i, j, 0) = make_double2(p[0], p[1]);
                    SC(i, j, 1) = make_double2(p[2], p[3]);
                    DL(i, j) = (int8_t)kexp;
                }
            }
        }

        // -------------------------------------------------------------- root: site likelihoods
        const double ps_c = prm[a.lay.off_ps + c];
        double pi[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) pi[s] = prm[a.lay.off_pi + s];
        double proot[K][4], rdot[K];
        __syncthreads();  // previous item's readers of ex_* are done
#pragma unroll
        for (int j = 0; j < K; ++j) {
            double2 u = ST(so_last, j, 0), v = ST(so_last, j, 1);
            proot[j][0] = u.x; proot[j][1] = u.y; proot[j][2] = v.x; proot[j][3] = v.y;
            rdot[j] = pi[0] * u.x + pi[1] * u.y + pi[2] * v.x + pi[3] * v.y;  // generate_script.py:1007
            ex_l[j * NT + tid] = ps_c * rdot[j];
            ex_e[j * NT + tid] = etot[j];
        }
        __syncthreads();
        double acc_logl = 0.0, acc_dps = 0.0, acc_dpi[4] = {0.0, 0.0, 0.0, 0.0};
        double fac[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int base = j * NT + pb * C * 32 + lane;
            int emin = ex_e[base];
            for (int cc = 1; cc < C; ++cc) emin = min(emin, ex_e[base + cc * 32]);
            double sum = 0.0;
            for (int cc = 0; cc < C; ++cc) {
                const int de = ex_e[base + cc * 32] - emin;
                sum += de > 15 ? 0.0 : ex_l[base + cc * 32] * pow2_64k(-de);
            }
            const double w = a.weights[pat0 + 32 * j];
            if (c == 0) acc_logl += w * (log(sum) - (double)emin * 44.361419555836500 /* 64 ln 2 */);
            const int de = etot[j] - emin;
            fac[j] = de > 15 ? 0.0 : w * pow2_64k(-de) / sum;
            if (GRAD) {
                acc_dps += fac[j] * rdot[j];
#pragma unroll
                for (int s = 0; s < 4; ++s) acc_dpi[s] += fac[j] * ps_c * proot[j][s];
            }
        }

        // -------------------------------------------------------------- pre-order
        if (GRAD) {
            {
                const int sroot = __ldg(&a.pre[0].sn);
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const double f = fac[j] * ps_c;
                    ST(sroot, j, 0) = make_double2(pi[0] * f, pi[1] * f);
                    ST(sroot, j, 1) = make_double2(pi[2] * f, pi[3] * f);
                }
            }
            double* Gd = a.G + ((size_t)d * a.nn * C + c) * 16;  // + node * C * 16
            for (int i = 0; i < a.nsteps; ++i) {
                const int4 s0 = __ldg(reinterpret_cast<const int4*>(a.pre + i));
                const int4 s1 = __ldg(reinterpret_cast<const int4*>(a.pre + i) + 1);
                const int4 s2 = __ldg(reinterpret_cast<const int4*>(a.pre + i) + 2);
                const int na = s0.y, nb = s0.z, sn = s0.w;
                const int sa = s1.x, sb = s1.y, rown = s1.z, rowa = s1.w, rowb = s2.x;
                double qn[K][4];
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    double2 u = ST(sn, j, 0), v = ST(sn, j, 1);
                    const double f = pow2_64k((int)DL(rown, j));
                    qn[j][0] = u.x * f; qn[j][1] = u.y * f; qn[j][2] = v.x * f; qn[j][3] = v.y * f;
                }
                double MA[16], MB[16];
                load_mat(Pd + (size_t)na * 16, MA);
                load_mat(Pd + (size_t)nb * 16, MB);
                double GA[16], GB[16];
#pragma unroll
                for (int x = 0; x < 16; ++x) GA[x] = GB[x] = 0.0;
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    double pa[4], pbv[4], ma[4], mb[4], AA[4], AB[4];
                    if (rowa < 0) {
                        tip_vec(a.tips[(size_t)na * a.Lpad + pat0 + 32 * j], pa);
                    } else {
                        double2 u = SC(rowa, j, 0), v = SC(rowa, j, 1);
                        pa[0] = u.x; pa[1] = u.y; pa[2] = v.x; pa[3] = v.y;
                    }
                    if (rowb < 0) {
                        tip_vec(a.tips[(size_t)nb * a.Lpad + pat0 + 32 * j], pbv);
                    } else {
                        double2 u = SC(rowb, j, 0), v = SC(rowb, j, 1);
                        pbv[0] = u.x; pbv[1] = u.y; pbv[2] = v.x; pbv[3] = v.y;
                    }
                    matvec(MA, pa, ma);
                    matvec(MB, pbv, mb);
#pragma unroll
                    for (int s = 0; s < 4; ++s) {  // eq (7) of eigen.j2:148
                        AA[s] = qn[j][s] * mb[s];
                        AB[s] = qn[j][s] * ma[s];
                    }
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                            GA[4 * x + y] = fma(AA[x], pa[y], GA[4 * x + y]);
                            GB[4 * x + y] = fma(AB[x], pbv[y], GB[4 * x + y]);
                        }
                    if (sa >= 0) {
                        double q[4];
                        matTvec(MA, AA, q);  // eigen.j2:151-153
                        ST(sa, j, 0) = make_double2(q[0], q[1]);
                        ST(sa, j, 1) = make_double2(q[2], q[3]);
                    }
                    if (sb >= 0) {
                        double q[4];
                        matTvec(MB, AB, q);
                        ST(sb, j, 0) = make_double2(q[0], q[1]);
                        ST(sb, j, 1) = make_double2(q[2], q[3]);
                    }
                }
                warp_reduce16_atomic(GA, Gd + (size_t)na * C * 16, lane);
                warp_reduce16_atomic(GB, Gd + (size_t)nb * C * 16, lane);
            }
        }

        // -------------------------------------------------------------- per-item scalars
        double* od = a.out + (size_t)d * a.nout;
        if (c == 0) {
            acc_logl = warp_sum(acc_logl);
            if (lane == 0) atomicAdd(od, acc_logl);
        }
        if (GRAD) {
            acc_dps = warp_sum(acc_dps);
#pragma unroll
            for (int s = 0; s < 4; ++s) acc_dpi[s] = warp_sum(acc_dpi[s]);
            if (lane == 0) {
                atomicAdd(od + a.off_out_ps + c, acc_dps);
#pragma unroll
                for (int s = 0; s < 4; ++s) atomicAdd(od + a.off_out_freqs + s, acc_dpi[s]);
            }
        }
    }
#undef ST
#undef SC
#undef DL
}

// ------------------------------------------------------------------------------------------
// K4: contraction of the branch statistics
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128) contract_kernel(const ContractArgs a) {
    const int d = blockIdx.y;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = b < a.bcount;
    const double* prm = a.params + (size_t)d * a.lay.stride;
    double* od = a.out + (size_t)d * a.nout;
    double Q[16], m1[16], m2[16], lam[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        Q[i] = prm[a.lay.off_Q + i];
        m1[i] = prm[a.lay.off_m1 + i];
        m2[i] = prm[a.lay.off_m2 + i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) lam[i] = prm[a.lay.off_lam + i];
    const double t = live ? prm[a.lay.off_t + b] : 0.0;
    double dt = 0.0;
    double th[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) th[k] = 0.0;
    for (int c = 0; c < a.C; ++c) {
        double g = 0.0;
        if (live) {
            const double r = prm[a.lay.off_rs + c];
            double G[16], P[16];
            load_mat(a.G + (((size_t)d * a.nn + b) * a.C + c) * 16, G);
            load_mat(a.P + (((size_t)d * a.C + c) * a.nn + b) * 16, P);
            // d logL / d tau = <G, Q P>   (dP/dtau = Q P)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double qp = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) qp = fma(Q[4 * i + k], P[4 * k + j], qp);
                    g = fma(G[4 * i + j], qp, g);
                }
            dt = fma(r, g, dt);
            if (a.lay.ntheta > 0) {
                // H = m1^T G m2^T ; d logL/dtheta += sum_ij H_ij F_ij X_ij
                const double tau = t * r;
                double T[16], H[16], ex[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) ex[i] = exp(lam[i] * tau);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double s = 0.0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) s = fma(m1[4 * k + i], G[4 * k + j], s);
                        T[4 * i + j] = s;
                    }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double s = 0.0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) s = fma(T[4 * i + k], m2[4 * j + k], s);
                        const double x = (lam[i] - lam[j]) * tau;
                        const double f = tau * ex[j] * (fabs(x) < 1e-8 ? 1.0 + 0.5 * x : expm1(x) / x);
                        H[4 * i + j] = s * f;
                    }
#pragma unroll
                for (int k = 0; k < 10; ++k)
                    if (k < a.lay.ntheta) {
                        const double* X = prm + a.lay.off_X + 16 * k;
                        double s = 0.0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) s = fma(H[i], X[i], s);
                        th[k] += s;
                    }
            }
        }
        const double gr = warp_sum(t * g);
        if (lane == 0 && gr != 0.0) atomicAdd(od + a.off_out_rs + c, gr);
    }
    if (live) od[1 + b] = dt;
#pragma unroll
    for (int k = 0; k < 10; ++k)
        if (k < a.lay.ntheta) {
            const double s = warp_sum(th[k]);
            if (lane == 0 && s != 0.0)
                atomicAdd(od + (k < a.nsubst ? a.off_out_subst + k : a.off_out_freqs + (k - a.nsubst)), s);
        }
}

template <int K, bool GRAD>
cudaError_t launch_sweep_t(const SweepArgs& a, int grid, int nthreads, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(sweep_kernel<K, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    sweep_kernel<K, GRAD><<<grid, nthreads, smem, stream>>>(a);
    return cudaGetLastError();
}

template <int K, bool GRAD>
cudaError_t occupancy_t(int nthreads, size_t smem, int* n) {
    cudaError_t e = cudaFuncSetAttribute(sweep_kernel<K, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, sweep_kernel<K, GRAD>, nthreads, smem);
}

}  // namespace

int sweep_max_threads(int K) { return max_threads(K); }

size_t sweep_smem_bytes(int D, int K, int nthreads) {
    return (size_t)D * K * 2 * nthreads * sizeof(double2) + (size_t)K * nthreads * (sizeof(double) + sizeof(int));
}

void launch_pmat(const double* params, ParamLayout lay, int bcount, int jc_closed, double* P, int B,
                 cudaStream_t stream) {
    const int total = B * lay.C * lay.nn;
    pmat_kernel<<<(total + 127) / 128, 128, 0, stream>>>(params, lay, bcount, jc_closed, P, total);
}

cudaError_t launch_sweep(const SweepArgs& a, int K, bool grad, int grid, int nthreads, size_t smem,
                         cudaStream_t stream) {
    switch (K * 2 + (grad ? 1 : 0)) {
        case 2: return launch_sweep_t<1, false>(a, grid, nthreads, smem, stream);
        case 3: return launch_sweep_t<1, true>(a, grid, nthreads, smem, stream);
        case 4: return launch_sweep_t<2, false>(a, grid, nthreads, smem, stream);
        case 5: return launch_sweep_t<2, true>(a, grid, nthreads, smem, stream);
        case 8: return launch_sweep_t<4, false>(a, grid, nthreads, smem, stream);
        case 9: return launch_sweep_t<4, true>(a, grid, nthreads, smem, stream);
    }
    return cudaErrorInvalidValue;
}

cudaError_t sweep_occupancy(int K, bool grad, int nthreads, size_t smem, int* n) {
    switch (K * 2 + (grad ? 1 : 0)) {
        case 2: return occupancy_t<1, false>(nthreads, smem, n);
        case 3: return occupancy_t<1, true>(nthreads, smem, n);
        case 4: return occupancy_t<2, false>(nthreads, smem, n);
        case 5: return occupancy_t<2, true>(nthreads, smem, n);
        case 8: return occupancy_t<4, false>(nthreads, smem, n);
        case 9: return occupancy_t<4, true>(nthreads, smem, n);
    }
    return cudaErrorInvalidValue;
}

void launch_contract(const ContractArgs& a, int B, cudaStream_t stream) {
    dim3 grid((a.bcount + 127) / 128, B);
    contract_kernel<<<grid, 128, 0, stream>>>(a);
}

}  // namespace phylo
