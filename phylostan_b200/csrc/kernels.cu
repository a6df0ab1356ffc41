// kernels.cu -- see kernels.cuh.  sm_100a only.
#include "kernels.cuh"

#include <algorithm>

// Build-time switches kept for A/B measurements (csrc/build_variant.sh, tools/ab_sweep.py; results in
// profiles/README.md, round 2); the defaults are the product.  On the 1000 x 100k bench shape: the shared-memory
// reduction removes ~130 instructions per warp-step but needs a stack slot's worth of shared memory at K = 4 (136.6 vs
// 142.4 evaluations/s) and changes nothing at K = 2, where it fits (117.7 vs 119.9); without the L2 prefetch 137.5; the
// pre-order tip shortcut removes 13 % of the FP64 work and is neutral to slower (register spills).
#ifndef PHYLO_RSM
#define PHYLO_RSM 0     // K > 1: 4x4 statistics summed over the warp through shared memory (0: shuffle exchange)
#endif
#ifndef PHYLO_PF
#define PHYLO_PF 1      // pre-order: L2 prefetch of the scratch lines two steps ahead
#endif
#ifndef PHYLO_TIPRING
#define PHYLO_TIPRING 1  // value-only kernels (128-thread CTAs, simple-tip handles) take their tip codes through the per-warp
                         // TipRing: +6 / +10 / +17 % at K = 4 / 2 / 1 (r2_ab_6.log); 2: gradient kernels too (+0.4 %, and a
                         // second traversal-ordered copy of the alignment: not worth it)
#endif
#ifndef PHYLO_TMRED
#define PHYLO_TMRED 1  // K > 1 fp64 gradient kernels of 128-thread CTAs: the two 4x4 statistics of a step are summed over the
                       // warp through TENSOR MEMORY used as a transpose unit (tmr_* below) instead of the select / shuffle exchange
#endif
#ifndef PHYLO_CHERRY_HOIST
#define PHYLO_CHERRY_HOIST 1  // cherry-table kernels (2: every kernel): the operand preload branches once per child instead of once per pattern
#endif
#ifndef PHYLO_ABL_CHERRY
#define PHYLO_ABL_CHERRY 0  // TIMING ABLATION ONLY (wrong results): 1 = the messages of cherries are not stored, 2 = and the
                            // pre-order reads one cached row instead of theirs -- what a cherry message table could save
#endif
#ifndef PHYLO_TMRED_DPACK
#define PHYLO_TMRED_DPACK 1   // tmr_store16: doubles unpacked inside the asm statement instead of by __double2loint / hiint
                              // (+1.1 %: no spills, fewer moves); 2: the loads pack theirs inside the statement too
#endif
#ifndef PHYLO_MSGTIP_EARLY
#define PHYLO_MSGTIP_EARLY 0  // message-statistic pre-order: a tip child's message (a column of P) is fetched where an internal
                              // child's row is, at the end of the previous step, so both define the operand registers at one place
#endif
#ifndef PHYLO_POSTSPLIT
#define PHYLO_POSTSPLIT 1  // post-order, child a: one matrix-vector loop per operand kind instead of copies into a common array
#endif
#ifndef PHYLO_PRETIP
#define PHYLO_PRETIP 0  // pre-order: a simple tip child's message is a column of P (tip records column-major)
#endif

namespace phylo {

namespace {

// ------------------------------------------------------------------------------------------
// precision traits: fp64 is the product path; fp32 is the optional "fp32 with scaling" mode
// ------------------------------------------------------------------------------------------

template <typename T>
struct Real;
template <>
struct Real<double> {
    typedef double2 vec;                      // storage vector: a 4-state entry is two of these
    static constexpr int kVec = 2;
    static constexpr int kRec = kRecBytes;    // [desc 64 | P_a slot 160 | P_b slot 160]
    static constexpr int kMat = 160;          // 4x4 row-major (128 B) or, for simple tips, 4x5 column-major
    static constexpr int kUnit = 64;          // rescale by powers of 2^64 ...
    static constexpr int kMaxK = 15;          // ... at most 2^960 at once
    __device__ static __forceinline__ double tiny() { return 2.938735877055719e-39; }  // 2^-128
    __device__ static __forceinline__ double pow2(int k) {  // 2^(64 k)
        return __hiloint2double((1023 + 64 * k) << 20, 0);
    }
    __device__ static __forceinline__ int exponent(double x) {  // floor(log2 x); -1023 if subnormal
        return ((__double2hiint(x) >> 20) & 0x7ff) - 1023;
    }
};
template <>
struct Real<float> {
    typedef float4 vec;
    static constexpr int kVec = 1;
    static constexpr int kRec = kRecBytesF32;  // [desc 64 | P_a slot 80 | P_b slot 80]
    static constexpr int kMat = 80;
    static constexpr int kUnit = 24;
    static constexpr int kMaxK = 4;
    __device__ static __forceinline__ float tiny() { return 5.9604644775390625e-08f; }  // 2^-24
    __device__ static __forceinline__ float pow2(int k) { return __int_as_float((127 + 24 * k) << 23); }
    __device__ static __forceinline__ int exponent(float x) { return ((__float_as_int(x) >> 23) & 0xff) - 127; }
};

// 2^(-bits) in double, 0 <= bits <= 1000
__device__ __forceinline__ double pow2_neg(int bits) { return __hiloint2double((1023 - bits) << 20, 0); }

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------

// 4-state entry <-> registers; `nt` is the distance (in vectors) between the two halves of an fp64 entry
__device__ __forceinline__ void ld4(const double2* p, int nt, double (&x)[4]) {
    const double2 u = p[0], v = p[nt];
    x[0] = u.x; x[1] = u.y; x[2] = v.x; x[3] = v.y;
}
__device__ __forceinline__ void ld4(const float4* p, int, float (&x)[4]) {
    const float4 u = p[0];
    x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w;
}
__device__ __forceinline__ void st4(double2* p, int nt, const double (&x)[4]) {
    p[0] = make_double2(x[0], x[1]);
    p[nt] = make_double2(x[2], x[3]);
}
__device__ __forceinline__ void st4(float4* p, int, const float (&x)[4]) { p[0] = make_float4(x[0], x[1], x[2], x[3]); }
// streaming (evict-first) variants for the scratch rows: written once, read once much later
__device__ __forceinline__ void ld4cs(const double2* p, int nt, double (&x)[4]) {
    const double2 u = __ldcs(p), v = __ldcs(p + nt);
    x[0] = u.x; x[1] = u.y; x[2] = v.x; x[3] = v.y;
}
__device__ __forceinline__ void ld4cs(const float4* p, int, float (&x)[4]) {
    const float4 u = __ldcs(p);
    x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w;
}
__device__ __forceinline__ void st4cs(double2* p, int nt, const double (&x)[4]) {
    __stcs(p, make_double2(x[0], x[1]));
    __stcs(p + nt, make_double2(x[2], x[3]));
}
__device__ __forceinline__ void st4cs(float4* p, int, const float (&x)[4]) {
    __stcs(p, make_float4(x[0], x[1], x[2], x[3]));
}

// 4x4 row-major matrix from shared memory (warp-uniform address: broadcast LDS.128)
__device__ __forceinline__ void lds_mat(const unsigned char* p, double (&m)[16]) {
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double2 v = q[i];
        m[2 * i] = v.x;
        m[2 * i + 1] = v.y;
    }
}
__device__ __forceinline__ void lds_mat(const unsigned char* p, float (&m)[16]) {
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 v = q[i];
        m[4 * i] = v.x; m[4 * i + 1] = v.y; m[4 * i + 2] = v.z; m[4 * i + 3] = v.w;
    }
}

// y = M x  (row-major)
template <typename T>
__device__ __forceinline__ void matvec(const T (&m)[16], const T (&x)[4], T (&y)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        y[i] = fma(m[4 * i + 3], x[3], fma(m[4 * i + 2], x[2], fma(m[4 * i + 1], x[1], m[4 * i] * x[0])));
}

// y = M^T x
template <typename T>
__device__ __forceinline__ void matTvec(const T (&m)[16], const T (&x)[4], T (&y)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
        y[j] = fma(m[12 + j], x[3], fma(m[8 + j], x[2], fma(m[4 + j], x[1], m[j] * x[0])));
}

// tip cell -> 0/1 partial.  IDX = false: `code` is the 4-bit state mask; IDX = true (simple tips):
// `code` is the column index 0..3, or 4 for an all-ones cell.
template <bool IDX, typename T>
__device__ __forceinline__ void tip_vec(unsigned code, T (&p)[4]) {
    if (IDX) code = (0xF8421u >> (4u * code)) & 0xFu;  // index -> mask: 0..3 -> 1,2,4,8 ; 4 -> 15
#pragma unroll
    for (int s = 0; s < 4; ++s) p[s] = ((code >> s) & 1u) ? T(1) : T(0);
}

// Simple-tip fast path (TIPS kernels): every tip cell is one-hot or all ones (what the reference's
// encoder produces, phylostan/utils.py:180-188) and is stored as a column index 0..3 / 4.  The record
// then holds P column-major with a fifth all-ones column (row sums of a stochastic matrix), so the
// message of a tip child is one 32-byte load instead of a matrix-vector product.
template <typename V, typename T>
__device__ __forceinline__ void tip_msg(const unsigned char* slot, unsigned idx, T (&m)[4]) {
    ld4(reinterpret_cast<const V*>(slot) + idx * (int)(4 * sizeof(T) / sizeof(V)), 1, m);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum 16 per-lane values over the 32 lanes of a warp with a recursive-halving exchange
// (8+4+2+1+1 = 16 shuffles instead of 80), then one 128-byte RED per warp: even lane 2i adds
// entry i to dst[i].
template <typename T>
__device__ __forceinline__ void warp_reduce16_atomic(const T (&v)[16], double* __restrict__ dst, int lane) {
    T a8[8], a4[4], a2[2], a1;
    bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        T send = hi ? v[i] : v[i + 8], keep = hi ? v[i + 8] : v[i];
        a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        T send = hi ? a8[i] : a8[i + 4], keep = hi ? a8[i + 4] : a8[i];
        a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        T send = hi ? a4[i] : a4[i + 2], keep = hi ? a4[i + 2] : a4[i];
        a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    hi = lane & 2;
    {
        T send = hi ? a2[0] : a2[1], keep = hi ? a2[1] : a2[0];
        a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    if (!(lane & 1)) atomicAdd(dst + ((lane >> 1) & 15), (double)a1);
}

// The same reduction split in two so that the low-parallelism tail of TWO reductions can be
// interleaved: head = the xor-16 and xor-8 levels (16 -> 4 values per lane), tail = xor-4, -2, -1 of
// both and the two REDs.
template <typename T>
__device__ __forceinline__ void warp_reduce16_head(const T (&v)[16], T (&a4)[4], int lane) {
    T a8[8];
    bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        T send = hi ? v[i] : v[i + 8], keep = hi ? v[i + 8] : v[i];
        a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        T send = hi ? a8[i] : a8[i + 4], keep = hi ? a8[i + 4] : a8[i];
        a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
}
template <typename T>
__device__ __forceinline__ void warp_reduce4x2_tail_atomic(const T (&x4)[4], const T (&y4)[4], double* __restrict__ dx,
                                                           double* __restrict__ dy, int lane) {
    T x2[2], y2[2], x1, y1;
    bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        T sx = hi ? x4[i] : x4[i + 2], kx = hi ? x4[i + 2] : x4[i];
        T sy = hi ? y4[i] : y4[i + 2], ky = hi ? y4[i + 2] : y4[i];
        x2[i] = kx + __shfl_xor_sync(0xffffffffu, sx, 4);
        y2[i] = ky + __shfl_xor_sync(0xffffffffu, sy, 4);
    }
    hi = lane & 2;
    {
        T sx = hi ? x2[0] : x2[1], kx = hi ? x2[1] : x2[0];
        T sy = hi ? y2[0] : y2[1], ky = hi ? y2[1] : y2[0];
        x1 = kx + __shfl_xor_sync(0xffffffffu, sx, 2);
        y1 = ky + __shfl_xor_sync(0xffffffffu, sy, 2);
    }
    // last level: the even lane finishes x, the odd lane finishes y -- one shuffle serves both
    const bool odd = lane & 1;
    const T mine = odd ? y1 : x1, other = odd ? x1 : y1;
    const T tot = mine + __shfl_xor_sync(0xffffffffu, other, 1);
    atomicAdd((odd ? dy : dx) + ((lane >> 1) & 15), (double)tot);
}

__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
// K consecutive bytes (K = 1, 2, 4) of a lane as one packed word: byte j = (w >> 8 j) & 0xff
template <int K>
__device__ __forceinline__ unsigned ldg_bytes(const uint8_t* p) {
    if (K == 4) return __ldg(reinterpret_cast<const unsigned*>(p));
    if (K == 2) return __ldg(reinterpret_cast<const unsigned short*>(p));
    return __ldg(p);
}
// same, through the coherent path: the rescale exponents are written earlier in the SAME kernel, so
// the read-only (.nc) path of __ldg must not be used for them
template <int K>
__device__ __forceinline__ unsigned ld_bytes(const uint8_t* p) {
    if (K == 4) return __ldcs(reinterpret_cast<const unsigned*>(p));
    if (K == 2) return __ldcs(reinterpret_cast<const unsigned short*>(p));
    return __ldcs(p);
}
template <int K>
__device__ __forceinline__ void stcs_bytes(uint8_t* p, unsigned w) {
    if (K == 4) __stcs(reinterpret_cast<unsigned*>(p), w);
    else if (K == 2) __stcs(reinterpret_cast<unsigned short*>(p), (unsigned short)w);
    else __stcs(p, (uint8_t)w);
}
#define BYTE_OF(w, j) (((w) >> (8 * (j))) & 0xffu)

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
}
__device__ __forceinline__ void prefetch_l1(const void* p) {
    asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p));
}

// message of a cherry for one pattern: entry `code` = 5 x + y of its 25 x 4 table (global memory, read-only path)
__device__ __forceinline__ void cherry_msg(const double* __restrict__ tab, unsigned code, double (&m)[4]) {
    const double2* t = reinterpret_cast<const double2*>(tab) + 2 * code;
    const double2 u = __ldg(t), v = __ldg(t + 1);
    m[0] = u.x; m[1] = u.y; m[2] = v.x; m[3] = v.y;
}
__device__ __forceinline__ void cherry_msg(const double*, unsigned, float (&)[4]) {}  // fp32 kernels never use the tables

// ------------------------------------------------------------------------------------------
// Warp sum of 32 per-lane doubles through tensor memory (no MMA): TMEM as a transpose unit
// ------------------------------------------------------------------------------------------
//
// Every lane writes its 32 values (64 words) into its own TMEM lane with tcgen05.st.32x32b; tcgen05.ld.16x256b then
// hands thread t, for g = 0..7, entry 4 g + (t & 3) of lanes t/4 and t/4 + 8 (second load, lane base 16: t/4 + 16 and
// t/4 + 24) -- measured layout, tools/micro/tmem_transpose_probe.cu.  Four lanes are summed in registers, the other
// eight (threads with the same t & 3) by three recursive-halving exchange levels, after which every thread owns ONE of
// the 32 sums: entry 4 (4 b4 + 2 b3 + b2) + (t & 3), with b4 b3 b2 the bits 4, 3, 2 of t.  ~90 instructions instead of
// the ~300 of two 16-value select / shuffle exchanges.
__device__ __forceinline__ void tmr_store16(uint32_t addr, const double (&v)[16]) {
#if PHYLO_TMRED_DPACK
#pragma unroll
    for (int q = 0; q < 4; ++q)
        asm volatile(
            "{\n .reg .b32 t<8>;\n mov.b64 {t0, t1}, %1;\n mov.b64 {t2, t3}, %2;\n mov.b64 {t4, t5}, %3;\n mov.b64 {t6, t7}, %4;\n"
            " tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {t0, t1, t2, t3, t4, t5, t6, t7};\n}\n" ::"r"(addr + 8 * q),
            "d"(v[4 * q]), "d"(v[4 * q + 1]), "d"(v[4 * q + 2]), "d"(v[4 * q + 3]));
#else
#pragma unroll
    for (int q = 0; q < 4; ++q)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(addr + 8 * q),
                     "r"(__double2loint(v[4 * q])), "r"(__double2hiint(v[4 * q])), "r"(__double2loint(v[4 * q + 1])),
                     "r"(__double2hiint(v[4 * q + 1])), "r"(__double2loint(v[4 * q + 2])), "r"(__double2hiint(v[4 * q + 2])),
                     "r"(__double2loint(v[4 * q + 3])), "r"(__double2hiint(v[4 * q + 3])));
#endif
}
// s[g] += entry 4 g + (t & 3) of the two lanes the load at `addr` (lane base 0 or 16) hands this thread
__device__ __forceinline__ void tmr_load_add(uint32_t addr, double (&s)[8], bool first) {
#if PHYLO_TMRED_DPACK > 1
    double d[16];
    asm volatile(
        "{\n .reg .b32 t<32>;\n"
        " tcgen05.ld.sync.aligned.16x256b.x8.b32 {t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, "
        "t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31}, [%16];\n"
        " tcgen05.wait::ld.sync.aligned;\n"
        " mov.b64 %0, {t0, t1};\n mov.b64 %1, {t2, t3};\n mov.b64 %2, {t4, t5};\n mov.b64 %3, {t6, t7};\n"
        " mov.b64 %4, {t8, t9};\n mov.b64 %5, {t10, t11};\n mov.b64 %6, {t12, t13};\n mov.b64 %7, {t14, t15};\n"
        " mov.b64 %8, {t16, t17};\n mov.b64 %9, {t18, t19};\n mov.b64 %10, {t20, t21};\n mov.b64 %11, {t22, t23};\n"
        " mov.b64 %12, {t24, t25};\n mov.b64 %13, {t26, t27};\n mov.b64 %14, {t28, t29};\n mov.b64 %15, {t30, t31};\n}\n"
        : "=d"(d[0]), "=d"(d[1]), "=d"(d[2]), "=d"(d[3]), "=d"(d[4]), "=d"(d[5]), "=d"(d[6]), "=d"(d[7]), "=d"(d[8]), "=d"(d[9]),
          "=d"(d[10]), "=d"(d[11]), "=d"(d[12]), "=d"(d[13]), "=d"(d[14]), "=d"(d[15])
        : "r"(addr));
#pragma unroll
    for (int g = 0; g < 8; ++g) s[g] = first ? d[2 * g] + d[2 * g + 1] : s[g] + (d[2 * g] + d[2 * g + 1]);
#else
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr));
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const double lo = __hiloint2double((int)r[4 * g + 1], (int)r[4 * g]);
        const double hi = __hiloint2double((int)r[4 * g + 3], (int)r[4 * g + 2]);
        s[g] = first ? lo + hi : s[g] + (lo + hi);
    }
#endif
}
// `tm` = this warp's TMEM address (its lane quadrant, 64 columns): columns 0..31 hold child b's 16 values of every lane,
// 32..63 child a's (tmr_store16); adds the 32 warp sums to db[0..15] / da[0..15]
__device__ __forceinline__ void tmr_finish(uint32_t tm, double* __restrict__ db, double* __restrict__ da, int lane) {
    asm volatile("tcgen05.wait::st.sync.aligned;\n");
    double s[8];
    tmr_load_add(tm, s, true);
    tmr_load_add(tm + (16u << 16), s, false);
    double a4[4], a2[2];
    bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double send = hi ? s[i] : s[i + 4], keep = hi ? s[i + 4] : s[i];
        a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double send = hi ? a4[i] : a4[i + 2], keep = hi ? a4[i + 2] : a4[i];
        a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    hi = lane & 4;
    const double send = hi ? a2[0] : a2[1], keep = hi ? a2[1] : a2[0];
    const double tot = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    const int e = ((lane >> 2) & 7) * 4 + (lane & 3);  // bits 4 3 2 of the lane pick g, bits 1 0 the entry within it
    atomicAdd(e < 16 ? db + e : da + (e - 16), tot);
}

// ------------------------------------------------------------------------------------------
// K1: instruction streams (step record = descriptor + both transition matrices)
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ void pmatrix(const double* __restrict__ prm, const ParamLayout& lay, int node, int c,
                                        int bcount, int jc_closed, double (&m)[16]) {
    if (node >= bcount) {  // root, or the unrooted second root child: no branch
#pragma unroll
        for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0;
        return;
    }
    const double tau = prm[lay.off_t + node] * prm[lay.off_rs + c];
    if (jc_closed) {  // generate_script.py:765-766, written with expm1: 1/4 - 1/4 e^{-x} cancels for short branches
        const double e1 = expm1(-tau / 0.75);
#pragma unroll
        for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 + 0.75 * e1 : -0.25 * e1;
        return;
    }
    // generate_script.py:824-829: P = m1 diag(e^{lambda tau}) m2.  For short branches (max |lambda| tau < 1) the
    // identical P = I + m1 diag(expm1(lambda tau)) m2 is used: the plain form gets the O(tau) off-diagonals
    // by cancelling m1 m2 against the identity, i.e. with absolute error 1e-16 on values of size tau.
    double lam[4], ex[4], lmax = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        lam[j] = prm[lay.off_lam + j];
        lmax = fmax(lmax, fabs(lam[j]));
    }
    const bool small = lmax * tau < 1.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) ex[j] = small ? expm1(lam[j] * tau) : exp(lam[j] * tau);
    const double* m1 = prm + lay.off_m1;
    const double* m2 = prm + lay.off_m2;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) s += (m1[4 * i + q] * ex[q]) * m2[4 * q + j];
            m[4 * i + j] = (small && i == j) ? 1.0 + s : s;
        }
}

__device__ __forceinline__ void store_mat(unsigned char* dst, const double (&m)[16], double) {
    double2* o2 = reinterpret_cast<double2*>(dst);
#pragma unroll
    for (int q = 0; q < 8; ++q) o2[q] = make_double2(m[2 * q], m[2 * q + 1]);
}
__device__ __forceinline__ void store_mat(unsigned char* dst, const double (&m)[16], float) {
    float4* o4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int q = 0; q < 4; ++q)
        o4[q] = make_float4((float)m[4 * q], (float)m[4 * q + 1], (float)m[4 * q + 2], (float)m[4 * q + 3]);
}

// column-major 4x5: columns 0..3 of P, then a column of ones
__device__ __forceinline__ void store_tipmat(unsigned char* dst, const double (&m)[16], double) {
    double* o = reinterpret_cast<double*>(dst);
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int x = 0; x < 4; ++x) o[4 * y + x] = m[4 * x + y];
#pragma unroll
    for (int x = 0; x < 4; ++x) o[16 + x] = 1.0;
}
__device__ __forceinline__ void store_tipmat(unsigned char* dst, const double (&m)[16], float) {
    float* o = reinterpret_cast<float*>(dst);
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int x = 0; x < 4; ++x) o[4 * y + x] = (float)m[4 * x + y];
#pragma unroll
    for (int x = 0; x < 4; ++x) o[16 + x] = 1.0f;
}

// Descriptors carry ready-made offsets (tip rows, stack slots, scratch rows, G blocks) so the sweep
// does no index arithmetic beyond pointer + offset.  One thread per (sweep, draw, category, step, child).
template <typename T>
__global__ void __launch_bounds__(128) stream_kernel(const StreamArgs a) {
    const int nrec_post = a.B * a.lay.C * a.npost, nrec_pre = a.B * a.lay.C * a.nsteps;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2 * (nrec_post + nrec_pre)) return;
    const int child = idx & 1, rr = idx >> 1;
    const int which = rr >= nrec_post;  // 0 post-order stream, 1 pre-order stream
    const int r = which ? rr - nrec_post : rr;
    const int ns = which ? a.nsteps : a.npost;
    const int d = r / (a.lay.C * ns), c = (r / ns) % a.lay.C, i = r % ns;
    const double* prm = a.params + (size_t)d * a.lay.stride;
    unsigned char* rec = (which ? a.spre : a.spost) + (size_t)r * Real<T>::kRec;
    int na, nb;
    int4* rd = reinterpret_cast<int4*>(rec);
    if (which == 0) {
        const int4* s = reinterpret_cast<const int4*>(a.post + i);
        const int4 s0 = __ldg(s);           // a, b, src_a, src_b
        na = s0.x; nb = s0.y;
        if (child == 0) {
            const int spill = __ldg(reinterpret_cast<const int*>(s + 1));
            PostRec pr;
            pr.tip_a = (long long)na * a.Lpad;
            pr.tip_b = (long long)nb * a.Lpad;
            const bool a_hbm = s0.z >= a.slots;  // parked above the capped shared-memory stack
            pr.off_a = s0.z < 0 ? 0 : a_hbm ? __ldg(a.node_row + na) * a.SS : s0.z * a.slot_stride;
            pr.off_b = 0;  // an internal second child is always the previous result (TOS)
            pr.off_spill = spill >= 0 && spill < a.slots ? spill * a.slot_stride : -1;
            pr.flags = (s0.z == kSrcTip ? 1 : 0) | (s0.w == kSrcTip ? 2 : 0) | (s0.z == kSrcTos ? 4 : 0) |
                       (s0.w == kSrcTos ? 8 : 0) | (a_hbm ? 16 : 0) | (spill >= a.slots ? 32 : 0);
            pr.row_a = na >= a.S ? __ldg(a.node_row + na) * a.SS : -1;
            pr.row_b = nb >= a.S ? __ldg(a.node_row + nb) * a.SS : -1;
            if (PHYLO_ABL_CHERRY) {
                if (na >= a.S) { const PostStep& q = a.post[__ldg(a.node_row + na)]; if (q.a < a.S && q.b < a.S) pr.row_a = -1; }
                if (nb >= a.S) { const PostStep& q = a.post[__ldg(a.node_row + nb)]; if (q.a < a.S && q.b < a.S) pr.row_b = -1; }
            }
            if (a.node_cherry) {  // a cherry's message comes from its table: nothing to store
                if (__ldg(a.node_cherry + na) >= 0) pr.row_a = -1;
                if (__ldg(a.node_cherry + nb) >= 0) pr.row_b = -1;
            }
            pr.pad0 = pr.pad1 = 0;
            pr.ctab_a = pr.ctab_b = 0;
            if (a.post_tables) {  // a leafified child: tip-like, its code row and its table
                const int ka = __ldg(a.node_cherry + na), kb = __ldg(a.node_cherry + nb);
                const double* tab = a.ctab + ((size_t)d * a.lay.C + c) * a.tab_entries * 4;
                if (ka >= 0) { pr.tip_a = a.ctips_off + (long long)ka * a.Lpad; pr.ctab_a = (long long)(tab + 4 * __ldg(a.tab_off + ka)); }
                if (kb >= 0) { pr.tip_b = a.ctips_off + (long long)kb * a.Lpad; pr.ctab_b = (long long)(tab + 4 * __ldg(a.tab_off + kb)); }
            }
            const int4* src = reinterpret_cast<const int4*>(&pr);
            rd[0] = src[0]; rd[1] = src[1]; rd[2] = src[2]; rd[3] = src[3];
        }
    } else {
        const int4* s = reinterpret_cast<const int4*>(a.pre + i);
        const int4 s0 = __ldg(s);  // node a b src_n
        na = s0.y; nb = s0.z;
        if (child == 0) {
            const int4 s1 = __ldg(s + 1);  // dst_b a_internal rown rowa
            const int4 s2 = __ldg(s + 2);  // rowb parkn parkb -
            const int rowb = s2.x;
            PreRec pr;
            pr.tip_a = (long long)na * a.Lpad;
            pr.tip_b = (long long)nb * a.Lpad;
            pr.row_a = s1.w >= 0 ? s1.w * a.SS : -1;
            pr.row_b = rowb >= 0 ? rowb * a.SS : -1;
            if (PHYLO_ABL_CHERRY == 2) {
                if (s1.w >= 0) { const PostStep& q = a.post[s1.w]; if (q.a < a.S && q.b < a.S) pr.row_a = 0; }
                if (rowb >= 0) { const PostStep& q = a.post[rowb]; if (q.a < a.S && q.b < a.S) pr.row_b = 0; }
            }
            pr.dl_n = s1.z >= 0 ? s1.z * a.KNT : -1;   // s1.z = rown; -1: the post-order skipped this node, no exponents
            const bool n_hbm = s0.w >= a.slots, b_hbm = s1.x >= a.slots;
            pr.off_n = s0.w < 0 ? -1 : n_hbm ? s2.y * a.SS : s0.w * a.slot_stride;   // parked: in the node's parking row
            pr.off_b = s1.x < 0 ? -1 : b_hbm ? s2.z * a.SS : s1.x * a.slot_stride;
            pr.g_a = na * a.lay.C * 16;
            pr.g_b = nb * a.lay.C * 16;
            pr.flags = (s1.y ? 1 : 0) | (n_hbm ? 2 : 0) | (b_hbm ? 4 : 0);
            pr.ctab_a = pr.ctab_b = 0;
            if (a.node_cherry) {
                const int ka = __ldg(a.node_cherry + na), kb = __ldg(a.node_cherry + nb);
                const double* tab = a.ctab + ((size_t)d * a.lay.C + c) * a.tab_entries * 4;
                if (ka >= 0) { pr.row_a = -1; pr.tip_a = a.ctips_off + (long long)ka * a.Lpad; pr.ctab_a = (long long)(tab + 4 * __ldg(a.tab_off + ka)); }
                if (kb >= 0) { pr.row_b = -1; pr.tip_b = a.ctips_off + (long long)kb * a.Lpad; pr.ctab_b = (long long)(tab + 4 * __ldg(a.tab_off + kb)); }
            }
            const int4* src = reinterpret_cast<const int4*>(&pr);
            rd[0] = src[0]; rd[1] = src[1]; rd[2] = src[2]; rd[3] = src[3];
        }
    }
    double m[16];
    const int node = child ? nb : na;
    pmatrix(prm, a.lay, node, c, a.bcount, a.jc_closed, m);
    if (a.tips_simple && (which == 0 || PHYLO_PRETIP || a.msg) && node < a.S) store_tipmat(rec + 64 + Real<T>::kMat * child, m, T());
    else store_mat(rec + 64 + Real<T>::kMat * child, m, T());
}

// ------------------------------------------------------------------------------------------
// K2/K3: fused depth-first post-order + pre-order sweep
// ------------------------------------------------------------------------------------------

// Per-warp record ring fed by cp.async.  Records sit contiguously: record i lives in ring slot
// i % 6 (3 chunks of 2 records); chunk n+2 is issued when chunk n starts being consumed.
template <int REC>
struct Ring {
    unsigned char* buf;          // this warp's ring (generic pointer into shared memory)
    unsigned sbuf;               // the same as a 32-bit shared-window address
    const unsigned char* next;   // next chunk to fetch from the stream
    int remaining;               // records not yet issued
    int wchunk;                  // chunk buffer (0..2) the next issue writes
    int slot;                    // ring slot (0..5) of the current record
    int lane;

    template <class F>
    __device__ __forceinline__ void issue(F&& extra) {
        const unsigned dst = sbuf + wchunk * (kRecChunk * REC) + lane * 16;
        const unsigned char* s = next + lane * 16;
        constexpr int kTail = kRecChunk * REC - 512;  // bytes of a two-record chunk beyond the first 32 x 16 B (fp64: 768 B per chunk)
        if (remaining >= 2) {
            if (kTail >= 0 || lane < (kRecChunk * REC) / 16) cp_async16(dst, s);
            if (kTail > 0 && lane < kTail / 16) cp_async16(dst + 512, s + 512);
        } else if (remaining == 1 && lane < REC / 16) {
            cp_async16(dst, s);
        }
        extra();            // the tip-code ring rides in the same commit group
        cp_async_commit();  // always commit: keeps the group count uniform
        next += kRecChunk * REC;
        remaining -= kRecChunk;
        wchunk = wchunk == kRecBufs - 1 ? 0 : wchunk + 1;
    }
    // afterwards records 0, 1 (and 2, 3 once step 0 ran) are in flight; call step(0) before reading
    template <class F>
    __device__ __forceinline__ void start(const unsigned char* stream, int n, F&& extra) {
        cp_async_wait<0>();
        __syncwarp();
        next = stream;
        remaining = n;
        wchunk = 0;
        slot = 0;
#pragma unroll
        for (int q = 0; q < kRecBufs - 1; ++q) issue(extra);
    }
    // top of step i: afterwards the records of steps i, i+1, i+2 (even i: and i+3) are readable, and so
    // is every byte-operand block committed before this call
    template <class F>
    __device__ __forceinline__ void step(int i, F&& extra) {
        if (i) slot = slot == kRecChunk * kRecBufs - 1 ? 0 : slot + 1;
        if ((i & 1) == 0) {
            __syncwarp();  // every lane is done with the chunk about to be overwritten
            issue(extra);
            cp_async_wait<kRecBufs - 2>();
            __syncwarp();
        }
    }
    __device__ __forceinline__ const unsigned char* rec(int ahead) const {
        int s = slot + ahead;
        if (s >= kRecChunk * kRecBufs) s -= kRecChunk * kRecBufs;
        return buf + s * REC;
    }
};

// Per-warp ring of TIP CODES in consumption order (TR kernels: 128-thread CTAs of simple-tip handles).  The handle
// keeps, per sweep, a copy of the tip codes laid out [tile][slot][32 K]: slot s of a tile holds the codes of the
// s-th tip child the sweep meets (child a before child b within a step).  A warp's codes therefore form ONE
// contiguous stream per tile, and a chunk of four slots is a single cp.async per lane that rides in the record
// ring's commit group -- no per-tip address arithmetic, no registers waiting on loads, and a look-ahead of 8-12
// slots (the two-step register look-ahead this replaces was the kernel's largest single stall: a tip-code line
// takes longer to arrive from HBM than two post-order steps run).  Twelve slots; a chunk is fetched at every
// even step while at most eight slots are unread, which keeps at least four complete slots ahead of the two
// steps that follow (a step consumes at most two).
template <int K>
struct TipRing {
    static constexpr int kSlot = 32 * K;   // bytes: this warp's codes of one tip
    static constexpr int kSlots = 12;
    static constexpr int kBytes = kSlots * kSlot;
    static constexpr int kLps = kSlot / 16;  // lanes (16 B each) per slot
    const unsigned char* buf;
    unsigned sbuf;
    int lane;
    const uint8_t* src;  // next slot to fetch of this tile's stream
    int left;            // slots of the stream not yet fetched
    int unread;          // slots fetched and not yet consumed
    int wpos, rpos;      // ring positions of the next chunk / the next read
    __device__ __forceinline__ void start(const uint8_t* stream, int nslots) {
        src = stream; left = nslots; unread = 0; wpos = 0; rpos = 0;
    }
    __device__ __forceinline__ void fetch() {
        if (unread <= kSlots - 4 && left > 0) {
            const int q = lane / kLps;  // slot of the chunk this lane copies into
            if (lane < 4 * kLps && q < left) cp_async16(sbuf + (wpos + q) * kSlot + (lane % kLps) * 16, src + lane * 16);
            src += 4 * kSlot;
            left -= 4;
            unread += 4;
            wpos = wpos == kSlots - 4 ? 0 : wpos + 4;
        }
    }
    // this lane's K codes of the next tip: byte j = (w >> 8 j) & 0xff
    __device__ __forceinline__ unsigned get() {
        const unsigned char* p = buf + rpos * kSlot + lane * K;
        rpos = rpos == kSlots - 1 ? 0 : rpos + 1;
        --unread;
        if (K == 4) return *reinterpret_cast<const unsigned*>(p);
        if (K == 2) return *reinterpret_cast<const unsigned short*>(p);
        return *p;
    }
};

// Sum 16 per-lane values over the warp through shared memory: every lane parks its 16 values
// (row x of red holds entry x of all lanes, rows padded to 34), then lane l adds up entry l & 15 over
// the 16 source lanes of its half (8 conflict-free 16-byte loads), one shuffle joins the halves and
// lanes 0..15 issue one 128-byte RED.  ~45 instructions and a short dependency chain instead of the
// ~110 of the select/shuffle exchange.
template <typename T>
__device__ __forceinline__ void warp_reduce16_smem(const T (&v)[16], T* __restrict__ red, double* __restrict__ dst,
                                                   int lane) {
#pragma unroll
    for (int x = 0; x < 16; ++x) red[x * 34 + lane] = v[x];
    __syncwarp();
    const T* row = red + (lane & 15) * 34 + (lane & 16);
    T s[4] = {T(0), T(0), T(0), T(0)};
    if (sizeof(T) == 8) {
        const double2* r2 = reinterpret_cast<const double2*>(row);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double2 u = r2[k];
            s[(2 * k) & 3] += (T)u.x;
            s[(2 * k + 1) & 3] += (T)u.y;
        }
    } else {
        const float4* r4 = reinterpret_cast<const float4*>(row);  // rows start 8-byte aligned only: see below
        (void)r4;
        const float2* r2 = reinterpret_cast<const float2*>(row);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 u = r2[k];
            s[(2 * k) & 3] += (T)u.x;
            s[(2 * k + 1) & 3] += (T)u.y;
        }
    }
    T t = (s[0] + s[1]) + (s[2] + s[3]);
    t += __shfl_xor_sync(0xffffffffu, t, 16);
    if (lane < 16) atomicAdd(dst + lane, (double)t);
    __syncwarp();  // the rows are rewritten by the next reduction
}

// NTC > 0: the CTA size is the compile-time constant NTC (all shared/scratch offsets fold into
// immediates); NTC == 0: generic CTA size read from blockDim (up to 512 threads).
// DEEP (gradient runs of deep trees): stack entries flagged by the stream live in the HBM scratch.
// JC (gradient runs of JC69 handles): Q = mu (J/4 - I), so Q P p = mu (mean(p) 1 - P p) and the branch
// derivative needs ONE number per (pattern, branch) -- the reference's own formula, eigen.j2:160-167 --
// instead of the 4x4 statistic: with T = sum_x q_n(x) (P_a p_a)(x) (P_b p_b)(x),
//   <A_b p_b^T, Q P_b> / mu = mean(p_b) sum(A_b) - T,   <A_a p_a^T, Q P_a> / mu = mean(p_a) sum(A_a) - T.
// The scalar goes to entry 0 of the branch's G block; the contraction multiplies by mu.
// TR: tip codes come through the per-warp TipRing from the handle's traversal-ordered copies (a.tips_post /
// a.tips_pre) instead of per-step loads from the [S][Lpad] rows.
// MSG (gradient runs of simple-tip handles): the MESSAGE statistic.  The post-order stores, at
// the parent's step, the messages mu_c = P_c p_c of its internal children instead of the parent's own partial (the
// same number of rows), so the pre-order never multiplies a partial by P again: A_b = q_n o mu_a, A_a = q_n o mu_b,
// a tip child's message is a column of P, and the statistic is G~_b = sum A_b mu_b^T = G_b P_b^T (72 instead of 104
// FP64 instructions per pattern and step with two internal children, 40 instead of 72 for a cherry).  The
// contraction undoes the factor analytically: <G_b, Q P_b> = <G~_b, Q> (Q and P_b commute), and
// <G_b, dP_b/dtheta> = <m1^T G~_b m2^T o Phi, X_theta> with Phi_ij = (e^{(l_i - l_j) tau} - 1) / (l_i - l_j), which
// amplifies rounding by e^{|l_i - l_j| tau}: the host only chooses MSG while that stays below e^12 for every
// branch and category of the batch.
// CH (with MSG): cherry tables -- see K3b below.
template <typename T, int K, bool GRAD, bool TIPS, int NTC, int MINB, bool DEEP, bool JC, bool TR, bool MSG = false,
          bool CH = false>
__global__ void __launch_bounds__(NTC ? NTC : 512, MINB) sweep_kernel(const SweepArgs a) {
    typedef Real<T> R;
    typedef typename R::vec V;
    constexpr int VP = R::kVec;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = NTC ? NTC : (int)blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = a.C, nsteps = a.nsteps;
    const int c = warp % C, pb = warp / C;
    // shared memory: [stack | ring | byte ring | reduction rows]; the root's category exchange borrows the
    // head of the stack region, which is empty between the two sweeps (see sweep_smem_bytes)
    V* st = reinterpret_cast<V*>(smem_raw);                                    // [D][K][VP][NT]
    double* ex_l = reinterpret_cast<double*>(smem_raw);                        // [K][NT]
    int* ex_e = reinterpret_cast<int*>(ex_l + K * NT);                         // [K][NT]
    unsigned char* sm_p = smem_raw + a.stack_bytes;
    Ring<R::kRec> ring;
    ring.buf = sm_p + warp * (R::kRec * kRecChunk * kRecBufs);
    sm_p += (NT / 32) * (R::kRec * kRecChunk * kRecBufs);
    TipRing<K> tr;
    tr.buf = sm_p + warp * TipRing<K>::kBytes;
    tr.sbuf = (unsigned)__cvta_generic_to_shared(tr.buf);
    tr.lane = lane;
    tr.start(nullptr, 0);
    if (TR) sm_p += (NT / 32) * TipRing<K>::kBytes;
    auto tipfetch = [&]() { if (TR) tr.fetch(); };
    // K == 1 (small, latency-bound problems): the two 4x4 statistics of a step are summed over the warp
    // in one pass, [32 entries][33] per warp; K > 1: one child at a time, [16 entries][34] per warp
    T* const red = reinterpret_cast<T*>(sm_p) + warp * (K == 1 ? 32 * 33 : 16 * 34);
    ring.sbuf = (unsigned)__cvta_generic_to_shared(ring.buf);
    ring.lane = lane;
    ring.next = nullptr;
    ring.remaining = 0;
    ring.wchunk = 0;
    ring.slot = 0;
    const int tpat = (NT / (32 * C)) * 32 * K;
    const int SS = K * VP * NT;  // vectors per stack slot / scratch row
    // the statistics' warp sums go through tensor memory: 64 columns, every warp its own lane quadrant
    constexpr bool kTmRed = PHYLO_TMRED && GRAD && !JC && K > 1 && NTC == 128 && sizeof(T) == 8;
    __shared__ uint32_t tmr_base;
    uint32_t tmr = 0u;
    if (kTmRed) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(&tmr_base)), "r"(64));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        tmr = tmr_base + ((uint32_t)(32 * (warp & 3)) << 16);
    }

    V* const stt = st + tid;
    V* const sct = reinterpret_cast<V*>(a.scratch) + (size_t)blockIdx.x * a.scratch_stride + tid;
    uint8_t* const dlt = a.dscr + (size_t)blockIdx.x * a.dscr_stride + tid * K;  // K exponents per lane, packed
    const uint8_t* const dlw = a.dscr + (size_t)blockIdx.x * a.dscr_stride + (tid - lane) * K;  // the warp's block

    // entry j of the stack slot / scratch row at vector offset `off`
#define ST(off, j) (stt + (off) + (j) * (VP * NT))
#define SC(off, j) (sct + (off) + (j) * (VP * NT))

    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int d = item / a.ntiles, tile = item - d * a.ntiles;
        const double* prm = a.params + (size_t)d * a.lay.stride;
        const int pat0 = tile * tpat + pb * 32 * K + lane * K;  // this lane's K consecutive patterns: pat0 + j
        const uint8_t* tipw = a.tips + (pat0 - lane * K);       // the warp's 32 K tip codes of a row start here
        const size_t stream_off = ((size_t)d * C + c) * nsteps * R::kRec;
        const int npost = a.npost;  // fewer than nsteps when the table nodes are leaves of the post-order
        const size_t stream_off_post = ((size_t)d * C + c) * npost * R::kRec;
        // post-order message tables (kTab kernels): a leafified child is tip-like, its message comes from its table by
        // the combined code of the tips below it, gathered one step ahead (ta / tb)
        constexpr bool kTab = CH && MSG && !TR && sizeof(T) == 8;

        // -------------------------------------------------------------- post-order
        int etot[K];
        T tos[K][4];  // most recent partial (top of stack), kept in registers
#pragma unroll
        for (int j = 0; j < K; ++j) {
            etot[j] = 0;
            tos[j][0] = tos[j][1] = tos[j][2] = tos[j][3] = T(0);
        }
        const size_t wblock = (size_t)(pat0 - lane * K) / (32 * K);  // this warp's block of 32 K patterns
        if (TR) tr.start(a.tips_post + wblock * a.S * TipRing<K>::kSlot, a.S);
        ring.start(a.spost + stream_off_post, npost, tipfetch);
        ring.step(0, tipfetch);
        // tip codes of the children of steps i (ca, cb), i+1 (ca1, cb1) and, inside the loop, i+2:
        // K codes per lane packed in one word, loaded two steps ahead of their use (TR: from the tip ring)
        const uint8_t* tipp = tipw + lane * K;
        unsigned ca = 0u, cb = 0u, ca1 = 0u, cb1 = 0u;
        if (!TR) {
            const PostRec* r0 = reinterpret_cast<const PostRec*>(ring.rec(0));
            if (r0->flags & 1) ca = ldg_bytes<K>(tipp + r0->tip_a);
            if (r0->flags & 2) cb = ldg_bytes<K>(tipp + r0->tip_b);
            if (npost > 1) {
                const PostRec* r1 = reinterpret_cast<const PostRec*>(ring.rec(1));
                if (r1->flags & 1) ca1 = ldg_bytes<K>(tipp + r1->tip_a);
                if (r1->flags & 2) cb1 = ldg_bytes<K>(tipp + r1->tip_b);
            }
        }
        // kTab: the table entries a step will gather are prefetched into L1 one step ahead (the codes are known by then),
        // so the gather itself costs what a tip's column lookup in shared memory costs and no register lives across steps
        if (kTab) {
            const longlong2 ct = *reinterpret_cast<const longlong2*>(ring.rec(0) + 48);
#pragma unroll
            for (int j = 0; j < K; ++j) {
                if (ct.x) prefetch_l1(reinterpret_cast<const double*>(ct.x) + 4 * BYTE_OF(ca, j));
                if (ct.y) prefetch_l1(reinterpret_cast<const double*>(ct.y) + 4 * BYTE_OF(cb, j));
            }
        }
        V* srow = sct;  // scratch row of step i
        uint8_t* drow = dlt;
        // value-only kernels unroll by two (measured +6 % at K = 4, +13 % at K = 2; nothing in the gradient
        // kernels, whose code is larger)
        constexpr int kPostUnroll = GRAD ? 1 : 2;
#pragma unroll kPostUnroll
        for (int i = 0; i < npost; ++i) {
            if (i) ring.step(i, tipfetch);
            const unsigned char* rec = ring.rec(0);
            const int4 s1 = *reinterpret_cast<const int4*>(rec + 16);  // off_a, off_b, off_spill, flags
            const int fl = s1.w;
            const longlong2 ctp = kTab ? *reinterpret_cast<const longlong2*>(rec + 48) : make_longlong2(0, 0);
            unsigned ca2 = 0u, cb2 = 0u;
            if (TR) {
                if (fl & 1) ca = tr.get();
                if (fl & 2) cb = tr.get();
            } else if (i + 2 < npost) {  // tip codes of step i+2 -> registers
                const PostRec* n = reinterpret_cast<const PostRec*>(ring.rec(2));
                const int nf = n->flags;
                if (nf & 1) ca2 = ldg_bytes<K>(tipp + n->tip_a);
                if (nf & 2) cb2 = ldg_bytes<K>(tipp + n->tip_b);
            }
            T ma[K][4], mb[K][4];
            if (DEEP && (fl & 16)) {  // rare: child a was parked above the capped stack, in its scratch row
                T M[16];
                lds_mat(rec + 64, M);
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    T p[4];
                    ld4cs(SC(s1.x, j), NT, p);
                    matvec(M, p, ma[j]);
                }
            } else if (kTab && ctp.x) {     // leafified child: its message comes from its table (prefetched into L1)
#pragma unroll
                for (int j = 0; j < K; ++j) cherry_msg(reinterpret_cast<const double*>(ctp.x), BYTE_OF(ca, j), ma[j]);
            } else if (TIPS && (fl & 1)) {  // tip child: its message is a column of P_a
#pragma unroll
                for (int j = 0; j < K; ++j) tip_msg<V>(rec + 64, BYTE_OF(ca, j), ma[j]);
            } else {
                T M[16];
                lds_mat(rec + 64, M);
#if PHYLO_POSTSPLIT
                if (fl & 4) {  // one loop per operand kind: no copies of the TOS into a common operand array
#pragma unroll
                    for (int j = 0; j < K; ++j) matvec(M, tos[j], ma[j]);
                } else if (!TIPS && (fl & 1)) {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        T p[4];
                        tip_vec<false>(BYTE_OF(ca, j), p);
                        matvec(M, p, ma[j]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        T p[4];
                        ld4(ST(s1.x, j), NT, p);
                        matvec(M, p, ma[j]);
                    }
                }
#else
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    T p[4];
                    if (!TIPS && (fl & 1)) {
                        tip_vec<false>(BYTE_OF(ca, j), p);
                    } else if (fl & 4) {
#pragma unroll
                        for (int s = 0; s < 4; ++s) p[s] = tos[j][s];
                    } else {
                        ld4(ST(s1.x, j), NT, p);
                    }
                    matvec(M, p, ma[j]);
                }
#endif
            }
            if (kTab && ctp.y) {
#pragma unroll
                for (int j = 0; j < K; ++j) cherry_msg(reinterpret_cast<const double*>(ctp.y), BYTE_OF(cb, j), mb[j]);
            } else if (TIPS && (fl & 2)) {
#pragma unroll
                for (int j = 0; j < K; ++j) tip_msg<V>(rec + 64 + R::kMat, BYTE_OF(cb, j), mb[j]);
            } else {
                T M[16];
                lds_mat(rec + 64 + R::kMat, M);
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    T p[4];
                    if (!TIPS && (fl & 2)) {
                        tip_vec<false>(BYTE_OF(cb, j), p);
                    } else {  // an internal second child is always the previous step's result
#pragma unroll
                        for (int s = 0; s < 4; ++s) p[s] = tos[j][s];
                    }
                    matvec(M, p, mb[j]);
                }
            }
            if (s1.z >= 0) {  // the previous result still waits for its sibling: park it in shared memory
#pragma unroll
                for (int j = 0; j < K; ++j) st4(ST(s1.z, j), NT, tos[j]);
            }
            if (DEEP && MSG && (fl & 32)) {  // rare: it is parked above the capped stack -- in its own scratch row, which the
                                             // message-statistic sweep has not written (its parent stores the message there later)
#pragma unroll
                for (int j = 0; j < K; ++j) st4cs(srow - SS + j * (VP * NT), NT, tos[j]);
            }
            if (GRAD && MSG) {  // the children's messages go to the children's rows; tips have none
                const int2 rw = *reinterpret_cast<const int2*>(rec + 32);  // row_a, row_b
                if (rw.x >= 0) {
#pragma unroll
                    for (int j = 0; j < K; ++j) st4cs(SC(rw.x, j), NT, ma[j]);
                }
                if (rw.y >= 0) {
#pragma unroll
                    for (int j = 0; j < K; ++j) st4cs(SC(rw.y, j), NT, mb[j]);
                }
            }
            unsigned kpack = 0u;
            bool tiny = false;
            const T kTiny = R::tiny();
#pragma unroll
            for (int j = 0; j < K; ++j) {
#pragma unroll
                for (int s = 0; s < 4; ++s) tos[j][s] = ma[j][s] * mb[j][s];
                tiny |= tos[j][0] < kTiny && tos[j][1] < kTiny && tos[j][2] < kTiny && tos[j][3] < kTiny;
            }
            if (tiny) {  // rare: per-(pattern,category) rescaling by exact powers of two
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    T (&p)[4] = tos[j];
                    if (p[0] < kTiny && p[1] < kTiny && p[2] < kTiny && p[3] < kTiny) {
                        const T mx = fmax(fmax(p[0], p[1]), fmax(p[2], p[3]));
                        if (mx > T(0)) {
                            const int kexp = min((-R::exponent(mx)) / R::kUnit, R::kMaxK);
                            const T f = R::pow2(kexp);
#pragma unroll
                            for (int s = 0; s < 4; ++s) p[s] *= f;
                            etot[j] += kexp;
                            kpack |= (unsigned)kexp << (8 * j);
                        }
                    }
                }
            }
            if (GRAD && !MSG) {
#pragma unroll
                for (int j = 0; j < K; ++j)
                    st4cs(srow + j * (VP * NT), NT, tos[j]);
            }
            if (GRAD) stcs_bytes<K>(drow, kpack);
            srow += SS;
            drow += K * NT;
            if (kTab && i + 1 < npost) {  // the next step's table entries -> L1 (its codes arrived a step ago)
                const longlong2 ct = *reinterpret_cast<const longlong2*>(ring.rec(1) + 48);
                if (ct.x) {
#pragma unroll
                    for (int j = 0; j < K; ++j) prefetch_l1(reinterpret_cast<const double*>(ct.x) + 4 * BYTE_OF(ca1, j));
                }
                if (ct.y) {
#pragma unroll
                    for (int j = 0; j < K; ++j) prefetch_l1(reinterpret_cast<const double*>(ct.y) + 4 * BYTE_OF(cb1, j));
                }
            }
            if (!TR) { ca = ca1; cb = cb1; ca1 = ca2; cb1 = cb2; }
        }

        // -------------------------------------------------------------- root: site likelihoods
        const double ps_c = prm[a.lay.off_ps + c];
        double pi[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) pi[s] = prm[a.lay.off_pi + s];
        if (GRAD) {  // overlaps the root exchange
            if (TR) tr.start(a.tips_pre + wblock * a.S * TipRing<K>::kSlot, a.S);
            ring.start(a.spre + stream_off, nsteps, tipfetch);
        }
        double rdot[K];
        __syncthreads();  // every warp has emptied its stack: the exchange arrays may borrow it
#pragma unroll
        for (int j = 0; j < K; ++j) {
            rdot[j] = pi[0] * (double)tos[j][0] + pi[1] * (double)tos[j][1] + pi[2] * (double)tos[j][2] +
                      pi[3] * (double)tos[j][3];  // generate_script.py:1007
            ex_l[j * NT + tid] = ps_c * rdot[j];
            ex_e[j * NT + tid] = etot[j];
        }
        __syncthreads();
        double acc_logl = 0.0, acc_dps = 0.0, acc_dpi[4] = {0.0, 0.0, 0.0, 0.0};
        constexpr int kMaxDe = 960 / R::kUnit;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int base = j * NT + pb * C * 32 + lane;
            int emin = ex_e[base];
            for (int cc = 1; cc < C; ++cc) emin = min(emin, ex_e[base + cc * 32]);
            double sum = 0.0;
            for (int cc = 0; cc < C; ++cc) {
                const int de = ex_e[base + cc * 32] - emin;
                sum += de > kMaxDe ? 0.0 : ex_l[base + cc * 32] * pow2_neg(R::kUnit * de);
            }
            const double w = a.weights[pat0 + j];
            if (c == 0) acc_logl += w * (log(sum) - (double)emin * (R::kUnit * 0.6931471805599453));
            const int de = etot[j] - emin;
            const double fac = de > kMaxDe ? 0.0 : w * pow2_neg(R::kUnit * de) / sum;
            if (GRAD) {
                acc_dps += fac * rdot[j];
                const double f = fac * ps_c;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    acc_dpi[s] += f * (double)tos[j][s];
                    tos[j][s] = (T)(pi[s] * f);  // q(root): weight, 1/L and the category share folded in
                }
            }
        }
        __syncthreads();  // the exchange has been read: the stack region is a stack again

        // -------------------------------------------------------------- pre-order
        if (GRAD) {
            ring.step(0, tipfetch);
            // packed byte operands of steps i (ca, cb, dcur = rescale exponents of the node) and i+1
            unsigned dcur, d1 = 0u;
            ca = cb = ca1 = cb1 = 0u;
            {
                const PreRec* r0 = reinterpret_cast<const PreRec*>(ring.rec(0));
                if (!TR && r0->row_a < 0) ca = ldg_bytes<K>(tipp + r0->tip_a);
                if (!TR && r0->row_b < 0) cb = ldg_bytes<K>(tipp + r0->tip_b);
                dcur = r0->dl_n >= 0 ? ld_bytes<K>(dlt + r0->dl_n) : 0u;  // -1: the post-order skipped this node
                if (nsteps > 1) {
                    const PreRec* r1 = reinterpret_cast<const PreRec*>(ring.rec(1));
                    if (!TR && r1->row_a < 0) ca1 = ldg_bytes<K>(tipp + r1->tip_a);
                    if (!TR && r1->row_b < 0) cb1 = ldg_bytes<K>(tipp + r1->tip_b);
                    d1 = r1->dl_n >= 0 ? ld_bytes<K>(dlt + r1->dl_n) : 0u;
                }
            }
            double* Gd = a.G + ((size_t)d * a.nn * C + c) * 16;  // + node * C * 16
            // children's partials of the current step; loaded from scratch during the previous step's
            // tail (after their last use there), so no extra registers and a reduction's worth of cover
            T pa[K][4], pbv[K][4];
            constexpr bool kTipEarly = PHYLO_MSGTIP_EARLY && MSG && !TR;
            // cherry tables (the stream marks such children: row = -1, tip_* = combined-code row, ctab_* >= 0): their
            // message is gathered where an internal child's row is loaded, one step ahead
            constexpr bool kCherry = CH && MSG && !TR && sizeof(T) == 8;
            {
                const unsigned char* r0 = ring.rec(0);
                const int4 r1 = *reinterpret_cast<const int4*>(r0 + 16);  // row_a, row_b, ...
                const longlong2 ct = kCherry ? *reinterpret_cast<const longlong2*>(r0 + 48) : make_longlong2(0, 0);
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    if (r1.x >= 0) ld4cs(SC(r1.x, j), NT, pa[j]);
                    else if (kCherry && ct.x) cherry_msg(reinterpret_cast<const double*>(ct.x), BYTE_OF(ca, j), pa[j]);
                    else if (kTipEarly) tip_msg<V>(r0 + 64, BYTE_OF(ca, j), pa[j]);
                    if (r1.y >= 0) ld4cs(SC(r1.y, j), NT, pbv[j]);
                    else if (kCherry && ct.y) cherry_msg(reinterpret_cast<const double*>(ct.y), BYTE_OF(cb, j), pbv[j]);
                    else if (kTipEarly) tip_msg<V>(r0 + 64 + R::kMat, BYTE_OF(cb, j), pbv[j]);
                }
            }
            for (int i = 0; i < nsteps; ++i) {
                if (i) ring.step(i, tipfetch);
                const unsigned char* rec = ring.rec(0);
                const int4 s1 = *reinterpret_cast<const int4*>(rec + 16);  // row_a, row_b, dl_n, off_n
                const int4 s2 = *reinterpret_cast<const int4*>(rec + 32);  // off_b, g_a, g_b, flags
                const int rowa = s1.x, rowb = s1.y;
                unsigned ca2 = 0u, cb2 = 0u, d2 = 0u;
                if (TR) {
                    if (rowa < 0) ca = tr.get();
                    if (rowb < 0) cb = tr.get();
                }
                if (i + 2 < nsteps) {  // operands of step i+2: bytes -> registers, scratch lines -> L2
                    const PreRec* n = reinterpret_cast<const PreRec*>(ring.rec(2));
                    const int4 n1 = *reinterpret_cast<const int4*>(reinterpret_cast<const unsigned char*>(n) + 16);
                    if (n1.x < 0) {
                        if (!TR) ca2 = ldg_bytes<K>(tipp + n->tip_a);
                    } else if (PHYLO_PF) {
#pragma unroll
                        for (int j = 0; j < K; ++j)
#pragma unroll
                            for (int h = 0; h < VP; ++h) prefetch_l2(SC(n1.x, j) + h * NT);
                    }
                    if (n1.y < 0) {
                        if (!TR) cb2 = ldg_bytes<K>(tipp + n->tip_b);
                    } else if (PHYLO_PF) {
#pragma unroll
                        for (int j = 0; j < K; ++j)
#pragma unroll
                            for (int h = 0; h < VP; ++h) prefetch_l2(SC(n1.y, j) + h * NT);
                    }
                    d2 = n1.z >= 0 ? ld_bytes<K>(dlt + n1.z) : 0u;
                }
                // q(node) lives in the TOS registers for the whole step: either it is still there (the
                // node was the previous step's first child) or it is popped from the shared-memory stack
                T (&qn)[K][4] = tos;
                if (DEEP && (s2.w & 2)) {  // rare: q(node) was parked above the capped stack, in p(node)'s scratch row
#pragma unroll
                    for (int j = 0; j < K; ++j) ld4cs(SC(s1.w, j), NT, qn[j]);
                }
                const bool pop = s1.w >= 0 && !(DEEP && (s2.w & 2));
#pragma unroll
                for (int j = 0; j < K; ++j)
                    if (pop) ld4(ST(s1.w, j), NT, qn[j]);
                if (dcur) {  // rare: this node was rescaled in the post-order for some pattern of this lane
#pragma unroll
                    for (int j = 0; j < K; ++j)
                        if (BYTE_OF(dcur, j)) {
                            const T f = R::pow2((int)BYTE_OF(dcur, j));
#pragma unroll
                            for (int s = 0; s < 4; ++s) qn[j][s] *= f;
                        }
                }
                // A_a = q_n o (P_b p_b), A_b = q_n o (P_a p_a)   (eq (7), eigen.j2:148).  One (warp-uniform)
                // branch per operand kind; a simple tip's message is a column of P (PRETIP streams).
                T Aa[K][4], Ab[K][4];
                const longlong2 ct0 = kCherry ? *reinterpret_cast<const longlong2*>(rec + 48) : make_longlong2(0, 0);
                if (!kTipEarly && rowa < 0 && !ct0.x) {  // MSG: pa / pbv hold MESSAGES, a tip's is a column of P
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        if (MSG) tip_msg<V>(rec + 64, BYTE_OF(ca, j), pa[j]);
                        else tip_vec<TIPS>(BYTE_OF(ca, j), pa[j]);
                    }
                }
                if (!kTipEarly && rowb < 0 && !ct0.y) {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        if (MSG) tip_msg<V>(rec + 64 + R::kMat, BYTE_OF(cb, j), pbv[j]);
                        else tip_vec<TIPS>(BYTE_OF(cb, j), pbv[j]);
                    }
                }
                if (MSG) {
#pragma unroll
                    for (int j = 0; j < K; ++j)
#pragma unroll
                        for (int s = 0; s < 4; ++s) Ab[j][s] = qn[j][s] * pa[j][s];
                } else if (PHYLO_PRETIP && TIPS && rowa < 0) {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        T m[4];
                        tip_msg<V>(rec + 64, BYTE_OF(ca, j), m);
#pragma unroll
                        for (int s = 0; s < 4; ++s) Ab[j][s] = qn[j][s] * m[s];
                    }
                } else {
                    T M[16], m[4];
                    lds_mat(rec + 64, M);
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        matvec(M, pa[j], m);
#pragma unroll
                        for (int s = 0; s < 4; ++s) Ab[j][s] = qn[j][s] * m[s];
                    }
                }
                T gb4[4];  // child b's statistics after the first two reduction levels (shuffle exchange)
                T jb = T(0), ja = T(0);  // JC: the scalar statistics of both branches
                (void)gb4;
                {   // child b: q(b) waits in shared memory
                    T G[16];
#pragma unroll
                    for (int x = 0; x < 16; ++x) G[x] = T(0);
                    const bool btip = PHYLO_PRETIP && TIPS && rowb < 0;  // tip: the message is a column of P_b
                    T M[16];
                    if (MSG ? s2.x >= 0 : !btip) lds_mat(rec + 64 + R::kMat, M);  // MSG: only q(b) needs P_b
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        T m[4];
                        if (MSG) {
#pragma unroll
                            for (int s = 0; s < 4; ++s) m[s] = pbv[j][s];
                        } else if (btip) tip_msg<V>(rec + 64 + R::kMat, BYTE_OF(cb, j), m);
                        else matvec(M, pbv[j], m);
#pragma unroll
                        for (int s = 0; s < 4; ++s) Aa[j][s] = qn[j][s] * m[s];
                        if (JC) {
                            const T t = fma(Ab[j][3], m[3], fma(Ab[j][2], m[2], fma(Ab[j][1], m[1], Ab[j][0] * m[0])));
                            const T sAb = (Ab[j][0] + Ab[j][1]) + (Ab[j][2] + Ab[j][3]);
                            const T sAa = (Aa[j][0] + Aa[j][1]) + (Aa[j][2] + Aa[j][3]);
                            const T mb_ = T(0.25) * ((pbv[j][0] + pbv[j][1]) + (pbv[j][2] + pbv[j][3]));
                            const T ma_ = T(0.25) * ((pa[j][0] + pa[j][1]) + (pa[j][2] + pa[j][3]));
                            jb += fma(mb_, sAb, -t);
                            ja += fma(ma_, sAa, -t);
                        } else {
#pragma unroll
                            for (int x = 0; x < 4; ++x)
#pragma unroll
                                for (int y = 0; y < 4; ++y)
                                    G[4 * x + y] = fma(Ab[j][x], pbv[j][y], G[4 * x + y]);
                        }
                    }
                    if (s2.x >= 0) {  // b is internal: q(b) = P_b^T A_b   (eigen.j2:151-153)
#pragma unroll
                        for (int j = 0; j < K; ++j) {
                            T q[4];
                            matTvec(M, Ab[j], q);
                            if (DEEP && (s2.w & 4)) st4cs(SC(s2.x, j), NT, q);  // over p(b)'s scratch row (just consumed)
                            else st4(ST(s2.x, j), NT, q);
                        }
                    }
                    if (JC) {
                    } else if (K == 1) {
#pragma unroll
                        for (int x = 0; x < 16; ++x) red[x * 33 + lane] = G[x];
                    } else if (kTmRed) {
                        tmr_store16(tmr, reinterpret_cast<const double (&)[16]>(G));
                    } else if (PHYLO_RSM) {
                        warp_reduce16_smem(G, red, Gd + s2.z, lane);
                    } else {
                        warp_reduce16_head(G, gb4, lane);
                    }
                }
                {   // child a: processed next when internal, so q(a) stays in the TOS registers
                    T G[16];
#pragma unroll
                    for (int x = 0; x < 16; ++x) G[x] = T(0);
                    if (!JC) {
#pragma unroll
                        for (int j = 0; j < K; ++j)
#pragma unroll
                            for (int x = 0; x < 4; ++x)
#pragma unroll
                                for (int y = 0; y < 4; ++y)
                                    G[4 * x + y] = fma(Aa[j][x], pa[j][y], G[4 * x + y]);
                    }
                    if (s2.w & 1) {
                        T M[16];
                        asm volatile("" ::: "memory");  // reload P_a instead of keeping it alive in registers
                        lds_mat(rec + 64, M);
#pragma unroll
                        for (int j = 0; j < K; ++j) matTvec(M, Aa[j], tos[j]);
                    }
                    if (i + 1 < nsteps) {  // pa / pbv are dead: fetch the next step's operands now
                        const unsigned char* nrec = ring.rec(1);
                        const int4 r1 = *reinterpret_cast<const int4*>(nrec + 16);
                        const longlong2 ct = kCherry ? *reinterpret_cast<const longlong2*>(nrec + 48) : make_longlong2(0, 0);
                        if ((kCherry && PHYLO_CHERRY_HOIST) || (PHYLO_CHERRY_HOIST > 1 && !kTipEarly)) {  // one warp-uniform branch per child, not one per pattern
                            if (r1.x >= 0) {
#pragma unroll
                                for (int j = 0; j < K; ++j) ld4cs(SC(r1.x, j), NT, pa[j]);
                            } else if (ct.x) {
#pragma unroll
                                for (int j = 0; j < K; ++j) cherry_msg(reinterpret_cast<const double*>(ct.x), BYTE_OF(ca1, j), pa[j]);
                            }
                            if (r1.y >= 0) {
#pragma unroll
                                for (int j = 0; j < K; ++j) ld4cs(SC(r1.y, j), NT, pbv[j]);
                            } else if (ct.y) {
#pragma unroll
                                for (int j = 0; j < K; ++j) cherry_msg(reinterpret_cast<const double*>(ct.y), BYTE_OF(cb1, j), pbv[j]);
                            }
                        } else {
#pragma unroll
                        for (int j = 0; j < K; ++j) {
                            if (r1.x >= 0) ld4cs(SC(r1.x, j), NT, pa[j]);
                            else if (kCherry && ct.x) cherry_msg(reinterpret_cast<const double*>(ct.x), BYTE_OF(ca1, j), pa[j]);
                            else if (kTipEarly) tip_msg<V>(nrec + 64, BYTE_OF(ca1, j), pa[j]);
                            if (r1.y >= 0) ld4cs(SC(r1.y, j), NT, pbv[j]);
                            else if (kCherry && ct.y) cherry_msg(reinterpret_cast<const double*>(ct.y), BYTE_OF(cb1, j), pbv[j]);
                            else if (kTipEarly) tip_msg<V>(nrec + 64 + R::kMat, BYTE_OF(cb1, j), pbv[j]);
                        }
                        }
                    }
                    if (JC) {  // two scalars per lane: lanes 0..15 finish child b's sum, lanes 16..31 child a's
                        const bool hi = lane & 16;
                        T v = (hi ? ja : jb) + __shfl_xor_sync(0xffffffffu, hi ? jb : ja, 16);
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        if ((lane & 15) == 0) atomicAdd(Gd + (hi ? s2.y : s2.z), (double)v);
                    } else if (K == 1) {
#pragma unroll
                        for (int x = 0; x < 16; ++x) red[(16 + x) * 33 + lane] = G[x];
                        __syncwarp();
                        // lane l owns entry l: 0..15 of child b, 16..31 of child a; four partial sums keep the chain short
                        const T* row = red + lane * 33;
                        T s0 = T(0), s1_ = T(0), s2s = T(0), s3 = T(0);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            s0 += row[j];
                            s1_ += row[j + 1];
                            s2s += row[j + 2];
                            s3 += row[j + 3];
                        }
                        atomicAdd(Gd + (lane < 16 ? s2.z + lane : s2.y + lane - 16), (double)((s0 + s1_) + (s2s + s3)));
                        __syncwarp();  // the rows are rewritten by the next step
                    } else if (kTmRed) {
                        tmr_store16(tmr + 32, reinterpret_cast<const double (&)[16]>(G));
                        tmr_finish(tmr, Gd + s2.z, Gd + s2.y, lane);
                    } else if (PHYLO_RSM) {
                        warp_reduce16_smem(G, red, Gd + s2.y, lane);
                    } else {
                        T ga4[4];
                        warp_reduce16_head(G, ga4, lane);
                        warp_reduce4x2_tail_atomic(gb4, ga4, Gd + s2.z, Gd + s2.y, lane);
                    }
                }
                if (!TR) { ca = ca1; cb = cb1; ca1 = ca2; cb1 = cb2; }
                dcur = d1; d1 = d2;
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
                if (a.lay.ntheta > 0) {  // JC69 fixes the frequencies: no derivative to report
#pragma unroll
                    for (int s = 0; s < 4; ++s) atomicAdd(od + a.off_out_freqs + s, acc_dpi[s]);
                }
            }
        }
    }
    cp_async_wait<0>();
    if (kTmRed) {
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmr_base), "r"(64));
    }
#undef ST
#undef SC
}

// Parallel build (csrc/build.sh): this file is compiled several times with -DPHYLO_PART=p, each part instantiating
// its share of sweep_kernel (1: fp64 simple tips, 2: fp64 masks, 3: fp32, 4: JC69 scalar / message statistic / cherry
// tables) and part 0 everything else (the kernels below, the launchers).  Without PHYLO_PART: one translation unit.
#if !defined(PHYLO_PART) || PHYLO_PART == 0
// ------------------------------------------------------------------------------------------
// K2b: the gradient sweep with its stack in TENSOR MEMORY (fp64, K patterns per lane, 128-thread CTAs)
// ------------------------------------------------------------------------------------------
//
// The sweep above is bound by the number of warps an SM can hold: 255 registers per thread and a 98 KB
// shared-memory stack per CTA allow two CTAs = two warps per scheduler, which leaves the FP64 pipe idle whenever both
// are between their FMA sections (profiles/README.md, round 2).  Blackwell's 256 KB of tensor memory per SM is
// otherwise unused here and can be addressed without any MMA: tcgen05.st / tcgen05.ld .32x32b give every thread of
// a warp its own TMEM lane, i.e. a private scratchpad of up to 512 words.  This variant keeps the stack of pending
// 4-state vectors there (one slot = K x 4 doubles = 8 K columns; measured as fast as shared memory,
// tools/micro/tmem_stack_bench.cu), which frees the shared memory for an OPERAND RING: the children's partials of
// step i+1 are copied from the HBM scratch into shared memory with cp.async while step i runs (one 4 KB slot per
// child and warp; a tip child's 0/1 vectors are decoded into the same kind of slot), so no partial waits in
// registers across a step boundary and every pass of a step reads its operands with two LDS.128.  With the
// operands out of the register file the step is written as short passes with ONE 4x4 accumulator live
// (A_b -> G_b -> q(b) -> A_a -> G_a -> q(a)), which fits 168 registers: three CTAs per SM.
// Stack positions beyond the TMEM slots are parked in the HBM scratch exactly as in the DEEP variant above.

// ("memory" clobbers or not, pops issued early or where they are needed, "=d" outputs instead of the 32-bit halves:
// all measured, profiles/r2_logs/r2_ab_13.log / r2_ab_14.log -- no difference beyond what the register allocator makes of them)
__device__ __forceinline__ void tm_st4(uint32_t addr, const double (&x)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(addr),
                 "r"(__double2loint(x[0])), "r"(__double2hiint(x[0])), "r"(__double2loint(x[1])), "r"(__double2hiint(x[1])),
                 "r"(__double2loint(x[2])), "r"(__double2hiint(x[2])), "r"(__double2loint(x[3])), "r"(__double2hiint(x[3]))
                 : "memory");
}
__device__ __forceinline__ void tm_ld4_raw(uint32_t addr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(addr)
                 : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
// The destination registers are guarded by the scoreboard like those of any load (the wait compiles to a scoreboard
// wait); the PTX model asks for tcgen05.wait::ld before they are read, and the values are passed through it so that
// no use is scheduled above it.
__device__ __forceinline__ void tm_wait_ld(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])::"memory");
}
template <int K>
__device__ __forceinline__ void tm_push(uint32_t addr, const double (&x)[K][4]) {
#pragma unroll
    for (int j = 0; j < K; ++j) tm_st4(addr + 8 * j, x[j]);
    tm_wait_st();
}
template <int K>
__device__ __forceinline__ void tm_pop(uint32_t addr, double (&x)[K][4]) {
    uint32_t r[K][8];
#pragma unroll
    for (int j = 0; j < K; ++j) tm_ld4_raw(addr + 8 * j, r[j]);
#pragma unroll
    for (int j = 0; j < K; ++j) tm_wait_ld(r[j]);  // the first one waits, the others only order the reads
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int s = 0; s < 4; ++s) x[j][s] = __hiloint2double((int)r[j][2 * s + 1], (int)r[j][2 * s]);
}

template <int MINB> struct TmCfg;
template <> struct TmCfg<3> { static constexpr int colsA = 128, colsB = 32; };  // 3 x 160 of the SM's 512 columns
template <> struct TmCfg<2> { static constexpr int colsA = 256, colsB = 0; };

// MSG: the message statistic (see sweep_kernel), simple-tip handles only: the post-order stores the children's
// messages, a tip child's operand slot receives its column of P, and the two passes over the operands become
// A_b / G~_b and A_a / G~_a without any product with P.
template <int K, bool TIPS, int MINB, bool MSG = false>
__global__ void __launch_bounds__(128, MINB) sweep_tm_kernel(const SweepArgs a) {
    typedef double T;
    typedef Real<double> R;
    typedef double2 V;
    constexpr int NT = 128, VP = 2;
    constexpr int kOps = kTmOpSlots;                 // operand slots per warp: both children of steps i and i+1
    constexpr int kOpBytes = K * VP * 32 * 16;       // one child's K partials of a warp, [j][half][lane] x 16 B
    constexpr int colsA = TmCfg<MINB>::colsA, colsB = TmCfg<MINB>::colsB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t tm_base[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = a.C, nsteps = a.nsteps;
    const int c = warp % C, pb = warp / C;
    // shared memory: [operand rings | record rings]; the root's category exchange borrows the operand rings, which
    // are empty between the two sweeps
    unsigned char* const ops = smem_raw + warp * (kOps * kOpBytes) + lane * 16;   // this lane's column of the warp's slots
    const unsigned ops_s = (unsigned)__cvta_generic_to_shared(ops);
    double* ex_l = reinterpret_cast<double*>(smem_raw);                        // [K][NT]
    int* ex_e = reinterpret_cast<int*>(ex_l + K * NT);                         // [K][NT]
    Ring<R::kRec> ring;
    ring.buf = smem_raw + (NT / 32) * (kOps * kOpBytes) + warp * (R::kRec * kRecChunk * kRecBufs);
    ring.sbuf = (unsigned)__cvta_generic_to_shared(ring.buf);
    ring.lane = lane;
    ring.next = nullptr;
    ring.remaining = 0;
    ring.wchunk = 0;
    ring.slot = 0;
    auto noop = []() {};

    // tensor memory: one warp allocates, everybody reads the base addresses
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tm_base[0])), "r"(colsA));
        if (colsB)
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(&tm_base[1])), "r"(colsB));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tm_lane = (uint32_t)(32 * (warp & 3)) << 16;
    const uint32_t tmA = tm_base[0] + tm_lane, tmB = (colsB ? tm_base[1] : 0u) + tm_lane - colsA;
    // column offset of a stack slot (8 K per slot, from the record) -> TMEM address of this warp's lanes
#define TM(off) ((colsB && (off) >= colsA) ? tmB + (uint32_t)(off) : tmA + (uint32_t)(off))

    const int tpat = (NT / (32 * C)) * 32 * K;
    V* const sct = reinterpret_cast<V*>(a.scratch) + (size_t)blockIdx.x * a.scratch_stride + tid;
    uint8_t* const dlt = a.dscr + (size_t)blockIdx.x * a.dscr_stride + tid * K;  // K exponents per lane, packed
#define SC(off, j) (sct + (off) + (j) * (VP * NT))
    // operand slot s of this warp: entry j of this lane
#define OP(s, j) (reinterpret_cast<const V*>(ops + (s) * kOpBytes) + (j) * (VP * 32))

    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int d = item / a.ntiles, tile = item - d * a.ntiles;
        const double* prm = a.params + (size_t)d * a.lay.stride;
        const int pat0 = tile * tpat + pb * 32 * K + lane * K;  // this lane's K consecutive patterns: pat0 + j
        const uint8_t* tipp = a.tips + pat0;
        const size_t stream_off = ((size_t)d * C + c) * nsteps * R::kRec;

        // -------------------------------------------------------------- post-order
        int etot[K];
        T tos[K][4];  // most recent partial (top of stack), kept in registers
#pragma unroll
        for (int j = 0; j < K; ++j) {
            etot[j] = 0;
            tos[j][0] = tos[j][1] = tos[j][2] = tos[j][3] = T(0);
        }
        ring.start(a.spost + stream_off, nsteps, noop);
        ring.step(0, noop);
        unsigned ca = 0u, cb = 0u, ca1 = 0u, cb1 = 0u;
        {
            const PostRec* r0 = reinterpret_cast<const PostRec*>(ring.rec(0));
            if (r0->flags & 1) ca = ldg_bytes<K>(tipp + r0->tip_a);
            if (r0->flags & 2) cb = ldg_bytes<K>(tipp + r0->tip_b);
            if (nsteps > 1) {
                const PostRec* r1 = reinterpret_cast<const PostRec*>(ring.rec(1));
                if (r1->flags & 1) ca1 = ldg_bytes<K>(tipp + r1->tip_a);
                if (r1->flags & 2) cb1 = ldg_bytes<K>(tipp + r1->tip_b);
            }
        }
        V* srow = sct;  // scratch row of step i
        uint8_t* drow = dlt;
        for (int i = 0; i < nsteps; ++i) {
            if (i) ring.step(i, noop);
            const unsigned char* rec = ring.rec(0);
            const int4 s1 = *reinterpret_cast<const int4*>(rec + 16);  // off_a, off_b, off_spill, flags
            const int fl = s1.w;
            unsigned ca2 = 0u, cb2 = 0u;
            if (i + 2 < nsteps) {  // tip codes of step i+2 -> registers
                const PostRec* n = reinterpret_cast<const PostRec*>(ring.rec(2));
                const int nf = n->flags;
                if (nf & 1) ca2 = ldg_bytes<K>(tipp + n->tip_a);
                if (nf & 2) cb2 = ldg_bytes<K>(tipp + n->tip_b);
            }
            T ma[K][4], mb[K][4];
            // child a: tip, the TOS, a TMEM slot or (rare) its scratch row
            if (TIPS && (fl & 1)) {
#pragma unroll
                for (int j = 0; j < K; ++j) tip_msg<V>(rec + 64, BYTE_OF(ca, j), ma[j]);
            } else {
                T pa[K][4];
                if (fl & 16) {
#pragma unroll
                    for (int j = 0; j < K; ++j) ld4cs(SC(s1.x, j), NT, pa[j]);
                } else if (fl & 1) {
#pragma unroll
                    for (int j = 0; j < K; ++j) tip_vec<false>(BYTE_OF(ca, j), pa[j]);
                } else if (fl & 4) {
#pragma unroll
                    for (int j = 0; j < K; ++j)
#pragma unroll
                        for (int s = 0; s < 4; ++s) pa[j][s] = tos[j][s];
                } else {
                    tm_pop<K>(TM(s1.x), pa);
                }
                T M[16];
                lds_mat(rec + 64, M);
#pragma unroll
                for (int j = 0; j < K; ++j) matvec(M, pa[j], ma[j]);
            }
            if (TIPS && (fl & 2)) {
#pragma unroll
                for (int j = 0; j < K; ++j) tip_msg<V>(rec + 64 + R::kMat, BYTE_OF(cb, j), mb[j]);
            } else {
                T M[16];
                lds_mat(rec + 64 + R::kMat, M);
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    T p[4];
                    if (!TIPS && (fl & 2)) {
                        tip_vec<false>(BYTE_OF(cb, j), p);
                    } else {  // an internal second child is always the previous step's result
#pragma unroll
                        for (int s = 0; s < 4; ++s) p[s] = tos[j][s];
                    }
                    matvec(M, p, mb[j]);
                }
            }
            if (s1.z >= 0) tm_push<K>(TM(s1.z), tos);  // the previous result still waits for its sibling
            if (MSG) {
                if (fl & 32) {  // rare: it is parked beyond the TMEM slots, in its own scratch row (not written otherwise)
#pragma unroll
                    for (int j = 0; j < K; ++j) st4cs(srow - K * VP * NT + j * (VP * NT), NT, tos[j]);
                }
                const int2 rw = *reinterpret_cast<const int2*>(rec + 32);  // row_a, row_b: the children's messages go there
                if (rw.x >= 0) {
#pragma unroll
                    for (int j = 0; j < K; ++j) st4cs(SC(rw.x, j), NT, ma[j]);
                }
                if (rw.y >= 0) {
#pragma unroll
                    for (int j = 0; j < K; ++j) st4cs(SC(rw.y, j), NT, mb[j]);
                }
            }
            unsigned kpack = 0u;
            bool tiny = false;
            const T kTiny = R::tiny();
#pragma unroll
            for (int j = 0; j < K; ++j) {
#pragma unroll
                for (int s = 0; s < 4; ++s) tos[j][s] = ma[j][s] * mb[j][s];
                tiny |= tos[j][0] < kTiny && tos[j][1] < kTiny && tos[j][2] < kTiny && tos[j][3] < kTiny;
            }
            if (tiny) {  // rare: per-(pattern,category) rescaling by exact powers of two
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    T (&p)[4] = tos[j];
                    if (p[0] < kTiny && p[1] < kTiny && p[2] < kTiny && p[3] < kTiny) {
                        const T mx = fmax(fmax(p[0], p[1]), fmax(p[2], p[3]));
                        if (mx > T(0)) {
                            const int kexp = min((-R::exponent(mx)) / R::kUnit, R::kMaxK);
                            const T f = R::pow2(kexp);
#pragma unroll
                            for (int s = 0; s < 4; ++s) p[s] *= f;
                            etot[j] += kexp;
                            kpack |= (unsigned)kexp << (8 * j);
                        }
                    }
                }
            }
            if (!MSG) {
#pragma unroll
                for (int j = 0; j < K; ++j) st4cs(srow + j * (VP * NT), NT, tos[j]);
            }
            stcs_bytes<K>(drow, kpack);
            srow += K * VP * NT;
            drow += K * NT;
            ca = ca1; cb = cb1; ca1 = ca2; cb1 = cb2;
        }

        // -------------------------------------------------------------- root: site likelihoods
        const double ps_c = prm[a.lay.off_ps + c];
        double pi[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) pi[s] = prm[a.lay.off_pi + s];
        ring.start(a.spre + stream_off, nsteps, noop);  // overlaps the root exchange
        double rdot[K];
        __syncthreads();  // every warp is done with its operand ring (previous item): the exchange may borrow it
#pragma unroll
        for (int j = 0; j < K; ++j) {
            rdot[j] = pi[0] * tos[j][0] + pi[1] * tos[j][1] + pi[2] * tos[j][2] + pi[3] * tos[j][3];  // generate_script.py:1007
            ex_l[j * NT + tid] = ps_c * rdot[j];
            ex_e[j * NT + tid] = etot[j];
        }
        __syncthreads();
        double acc_logl = 0.0, acc_dps = 0.0, acc_dpi[4] = {0.0, 0.0, 0.0, 0.0};
        constexpr int kMaxDe = 960 / R::kUnit;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int base = j * NT + pb * C * 32 + lane;
            int emin = ex_e[base];
            for (int cc = 1; cc < C; ++cc) emin = min(emin, ex_e[base + cc * 32]);
            double sum = 0.0;
            for (int cc = 0; cc < C; ++cc) {
                const int de = ex_e[base + cc * 32] - emin;
                sum += de > kMaxDe ? 0.0 : ex_l[base + cc * 32] * pow2_neg(R::kUnit * de);
            }
            const double w = a.weights[pat0 + j];
            if (c == 0) acc_logl += w * (log(sum) - (double)emin * (R::kUnit * 0.6931471805599453));
            const int de = etot[j] - emin;
            const double fac = de > kMaxDe ? 0.0 : w * pow2_neg(R::kUnit * de) / sum;
            acc_dps += fac * rdot[j];
            const double f = fac * ps_c;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                acc_dpi[s] += f * tos[j][s];
                tos[j][s] = pi[s] * f;  // q(root): weight, 1/L and the category share folded in
            }
        }
        __syncthreads();  // the exchange has been read: the region is the operand ring again

        // -------------------------------------------------------------- pre-order
        {
            // operand slots: child a of step i -> slot (2 i) & 3, child b -> slot (2 i + 1) & 3.  An internal child's
            // partial is copied from its scratch row one step ahead (cp.async, in the record ring's commit
            // group); a tip child's 0/1 vectors are written into its slot at the top of its own step.
            auto fetch = [&](const unsigned char* r, int st2) {   // st2 = (2 * step) & 3
                const int4 r1 = *reinterpret_cast<const int4*>(r + 16);  // row_a, row_b, ...
                if (r1.x >= 0) {
                    const V* src = SC(r1.x, 0);
                    const unsigned dst = ops_s + st2 * kOpBytes;
#pragma unroll
                    for (int q = 0; q < K * VP; ++q) cp_async16(dst + q * 512, src + q * NT);
                }
                if (r1.y >= 0) {
                    const V* src = SC(r1.y, 0);
                    const unsigned dst = ops_s + (st2 + 1) * kOpBytes;
#pragma unroll
                    for (int q = 0; q < K * VP; ++q) cp_async16(dst + q * 512, src + q * NT);
                }
            };
            cp_async_wait<0>();   // records 0..3 of the pre-order stream
            __syncwarp();
            fetch(ring.rec(0), 0);
            cp_async_commit();
            // packed byte operands of steps i (ca, cb, dcur = rescale exponents of the node) and i+1
            unsigned dcur, d1 = 0u;
            ca = cb = ca1 = cb1 = 0u;
            {
                const PreRec* r0 = reinterpret_cast<const PreRec*>(ring.rec(0));
                if (r0->row_a < 0) ca = ldg_bytes<K>(tipp + r0->tip_a);
                if (r0->row_b < 0) cb = ldg_bytes<K>(tipp + r0->tip_b);
                dcur = r0->dl_n >= 0 ? ld_bytes<K>(dlt + r0->dl_n) : 0u;
                if (nsteps > 1) {
                    const PreRec* r1 = reinterpret_cast<const PreRec*>(ring.rec(1));
                    if (r1->row_a < 0) ca1 = ldg_bytes<K>(tipp + r1->tip_a);
                    if (r1->row_b < 0) cb1 = ldg_bytes<K>(tipp + r1->tip_b);
                    d1 = r1->dl_n >= 0 ? ld_bytes<K>(dlt + r1->dl_n) : 0u;
                }
            }
            double* Gd = a.G + ((size_t)d * a.nn * C + c) * 16;  // + node * C * 16
            for (int i = 0; i < nsteps; ++i) {
                if (i) ring.slot = ring.slot == kRecChunk * kRecBufs - 1 ? 0 : ring.slot + 1;
                cp_async_wait<0>();  // this step's operands and the records up to i + 2 (even i: i + 3)
                __syncwarp();        // every lane is done with the record chunk about to be overwritten
                const int sa = (2 * i) & 3, sb = sa + 1, sn = (sa + 2) & 3;
                const unsigned char* rec = ring.rec(0);
                const int4 s1 = *reinterpret_cast<const int4*>(rec + 16);  // row_a, row_b, dl_n, off_n
                const int4 s2 = *reinterpret_cast<const int4*>(rec + 32);  // off_b, g_a, g_b, flags
                const bool npop = s1.w >= 0 && !(s2.w & 2);  // q(node) waits in tensor memory
                if ((i & 1) == 0) {
                    ring.issue([&]() { if (i + 1 < nsteps) fetch(ring.rec(1), sn); });
                } else {
                    if (i + 1 < nsteps) fetch(ring.rec(1), sn);
                    cp_async_commit();
                }
                const int rowa = s1.x, rowb = s1.y;
                unsigned ca2 = 0u, cb2 = 0u, d2 = 0u;
                if (i + 2 < nsteps) {  // byte operands of step i+2 -> registers
                    const PreRec* n = reinterpret_cast<const PreRec*>(ring.rec(2));
                    const int4 n1 = *reinterpret_cast<const int4*>(reinterpret_cast<const unsigned char*>(n) + 16);
                    if (n1.x < 0) ca2 = ldg_bytes<K>(tipp + n->tip_a);
                    if (n1.y < 0) cb2 = ldg_bytes<K>(tipp + n->tip_b);
                    d2 = n1.z >= 0 ? ld_bytes<K>(dlt + n1.z) : 0u;
                }
                // tip children: decode into their operand slots
                if (rowa < 0) {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        T p[4];
                        if (MSG) tip_msg<V>(rec + 64, BYTE_OF(ca, j), p);  // the message: a column of P_a
                        else tip_vec<TIPS>(BYTE_OF(ca, j), p);
                        st4(const_cast<V*>(OP(sa, j)), 32, p);
                    }
                }
                if (rowb < 0) {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        T p[4];
                        if (MSG) tip_msg<V>(rec + 64 + R::kMat, BYTE_OF(cb, j), p);
                        else tip_vec<TIPS>(BYTE_OF(cb, j), p);
                        st4(const_cast<V*>(OP(sb, j)), 32, p);
                    }
                }
                // q(node): still in the TOS registers, popped from tensor memory, or (rare) parked in p(node)'s scratch row
                T (&qn)[K][4] = tos;
                if (s2.w & 2) {
#pragma unroll
                    for (int j = 0; j < K; ++j) ld4cs(SC(s1.w, j), NT, qn[j]);
                } else if (npop) {
                    tm_pop<K>(TM(s1.w), qn);
                }
                if (dcur) {  // rare: this node was rescaled in the post-order for some pattern of this lane
#pragma unroll
                    for (int j = 0; j < K; ++j)
                        if (BYTE_OF(dcur, j)) {
                            const T f = R::pow2((int)BYTE_OF(dcur, j));
#pragma unroll
                            for (int s = 0; s < 4; ++s) qn[j][s] *= f;
                        }
                }
                // A_b = q_n o (P_a p_a)   (eq (7), eigen.j2:148)
                T Ab[K][4];
                if (MSG) {  // the operand slots hold messages: A_b = q_n o mu_a
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        T m[4];
                        ld4(OP(sa, j), 32, m);
#pragma unroll
                        for (int s = 0; s < 4; ++s) Ab[j][s] = qn[j][s] * m[s];
                    }
                } else {
                    T M[16];
                    lds_mat(rec + 64, M);
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        T p[4], m[4];
                        ld4(OP(sa, j), 32, p);
                        matvec(M, p, m);
#pragma unroll
                        for (int s = 0; s < 4; ++s) Ab[j][s] = qn[j][s] * m[s];
                    }
                }
                // G_b = sum_j A_b p_b^T
                T gb4[4];
                {
                    T G[16];
#pragma unroll
                    for (int x = 0; x < 16; ++x) G[x] = T(0);
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        T p[4];
                        ld4(OP(sb, j), 32, p);
#pragma unroll
                        for (int x = 0; x < 4; ++x)
#pragma unroll
                            for (int y = 0; y < 4; ++y) G[4 * x + y] = fma(Ab[j][x], p[y], G[4 * x + y]);
                    }
                    warp_reduce16_head(G, gb4, lane);
                }
                // q(b) = P_b^T A_b (eigen.j2:151-153) waits in tensor memory; A_a = q_n o (P_b p_b) takes q_n's registers
                {
                    T M[16];
                    if (!MSG || s2.x >= 0) lds_mat(rec + 64 + R::kMat, M);  // MSG: only q(b) needs P_b
                    if (s2.x >= 0) {
                        T q[K][4];
#pragma unroll
                        for (int j = 0; j < K; ++j) matTvec(M, Ab[j], q[j]);
                        if (s2.w & 4) {  // rare: parked over p(b)'s scratch row (already copied into the operand ring)
#pragma unroll
                            for (int j = 0; j < K; ++j) st4cs(SC(s2.x, j), NT, q[j]);
                        } else {
                            tm_push<K>(TM(s2.x), q);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        T p[4], m[4];
                        if (MSG) ld4(OP(sb, j), 32, m);
                        else {
                            ld4(OP(sb, j), 32, p);
                            matvec(M, p, m);
                        }
#pragma unroll
                        for (int s = 0; s < 4; ++s) qn[j][s] *= m[s];  // now A_a
                    }
                }
                T (&Aa)[K][4] = tos;
                {
                    T G[16], ga4[4];
#pragma unroll
                    for (int x = 0; x < 16; ++x) G[x] = T(0);
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        T p[4];
                        ld4(OP(sa, j), 32, p);
#pragma unroll
                        for (int x = 0; x < 4; ++x)
#pragma unroll
                            for (int y = 0; y < 4; ++y) G[4 * x + y] = fma(Aa[j][x], p[y], G[4 * x + y]);
                    }
                    warp_reduce16_head(G, ga4, lane);
                    warp_reduce4x2_tail_atomic(gb4, ga4, Gd + s2.z, Gd + s2.y, lane);
                }
                if (s2.w & 1) {  // a is internal and is processed next: q(a) = P_a^T A_a becomes the TOS
                    T M[16];
                    lds_mat(rec + 64, M);
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        T q[4];
                        matTvec(M, Aa[j], q);
#pragma unroll
                        for (int s = 0; s < 4; ++s) tos[j][s] = q[s];
                    }
                }
                ca = ca1; cb = cb1; ca1 = ca2; cb1 = cb2;
                dcur = d1; d1 = d2;
            }
        }

        // -------------------------------------------------------------- per-item scalars
        double* od = a.out + (size_t)d * a.nout;
        if (c == 0) {
            acc_logl = warp_sum(acc_logl);
            if (lane == 0) atomicAdd(od, acc_logl);
        }
        acc_dps = warp_sum(acc_dps);
#pragma unroll
        for (int s = 0; s < 4; ++s) acc_dpi[s] = warp_sum(acc_dpi[s]);
        if (lane == 0) {
            atomicAdd(od + a.off_out_ps + c, acc_dps);
            if (a.lay.ntheta > 0) {  // JC69 fixes the frequencies: no derivative to report
#pragma unroll
                for (int s = 0; s < 4; ++s) atomicAdd(od + a.off_out_freqs + s, acc_dpi[s]);
            }
        }
    }
    cp_async_wait<0>();
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm_base[0]), "r"(colsA));
        if (colsB) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm_base[1]), "r"(colsB));
    }
#undef TM
#undef SC
#undef OP
}

// ------------------------------------------------------------------------------------------
// K3b: cherry tables (message-statistic runs)
// ------------------------------------------------------------------------------------------
//
// SURVEY 8f rank 3 in the form that fits this design (the reference's unfinished pruner/ re-uses a subtree's partial
// when the tip states under it repeat, pruner/tree.cpp:140-174): a cherry's partial depends on the pattern only through
// the states of its two tips, 5 x 5 code pairs, so its MESSAGE to its parent, P_n ((P_a)[:, x] o (P_b)[:, y]), is a
// 25-entry table per (draw, category, cherry).  The post-order then stores nothing for a cherry and the pre-order
// gathers its message from the table (L1 / L2) instead of reading a scratch row back: a third of the internal nodes of
// a coalescent tree are cherries, so a third of the scratch traffic goes away (26.4 -> 17.6 GB per evaluation on
// config 3).  The arithmetic is the sweep's own, rescaling rule included (it depends on the partial only).  The same
// holds one level up: a pitchfork (a cherry and a tip, three tips below) has 125 code triples, 25 x + 5 y + z -- one
// byte still -- and gets a 125-entry table (PHYLO_B200_TABLE_TIPS=2 restricts the tables to cherries).

__global__ void __launch_bounds__(256) cherry_codes_kernel(const uint8_t* __restrict__ tips, uint8_t* __restrict__ ctips,
                                                            const int32_t* __restrict__ cherries, int Lpad, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t k = i / (size_t)Lpad, l = i - k * (size_t)Lpad;
        const int32_t* r = cherries + kTabRec * k;
        int code = 5 * tips[(size_t)r[2] * Lpad + l] + tips[(size_t)r[3] * Lpad + l];
        if (r[1] == 3) code = 5 * code + tips[(size_t)r[4] * Lpad + l];
        ctips[i] = (uint8_t)code;
    }
}

// p <- p rescaled by the sweep's rule (all four entries below 2^-128: an exact power of 2^64)
__device__ __forceinline__ void table_rescale(double (&p)[4]) {
    typedef Real<double> R;
    const double kTiny = R::tiny();
    if (p[0] < kTiny && p[1] < kTiny && p[2] < kTiny && p[3] < kTiny) {
        const double mx = fmax(fmax(p[0], p[1]), fmax(p[2], p[3]));
        if (mx > 0.0) {
            const double f = R::pow2(min((-R::exponent(mx)) / R::kUnit, R::kMaxK));
#pragma unroll
            for (int s = 0; s < 4; ++s) p[s] *= f;
        }
    }
}
// column x of P (x = 4: the all-ones column of an ambiguous cell), P row-major
__device__ __forceinline__ void table_col(const double (&P)[16], int x, double (&m)[4]) {
#pragma unroll
    for (int s = 0; s < 4; ++s) m[s] = x < 4 ? P[4 * s + x] : 1.0;
}

__global__ void __launch_bounds__(128) cherry_table_kernel(const CherryArgs a) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * a.C * a.ncherry) return;
    const int k = idx % a.ncherry, c = (idx / a.ncherry) % a.C, d = idx / (a.ncherry * a.C);
    const double* prm = a.params + (size_t)d * a.lay.stride;
    const int32_t* r = a.cherries + kTabRec * k;
    const int n = r[0], ntips = r[1], shape = r[6];
    double* out = a.ctab + (((size_t)d * a.C + c) * a.tab_entries + r[7]) * 4;
    double P0[16], P1[16], P2[16], Pc[16], Pn[16];
    pmatrix(prm, a.lay, r[2], c, a.bcount, a.jc_closed, P0);
    pmatrix(prm, a.lay, r[3], c, a.bcount, a.jc_closed, P1);
    pmatrix(prm, a.lay, n, c, a.bcount, a.jc_closed, Pn);
    if (ntips == 3) {
        pmatrix(prm, a.lay, r[4], c, a.bcount, a.jc_closed, P2);
        pmatrix(prm, a.lay, r[5], c, a.bcount, a.jc_closed, Pc);
    }
    const int nz = ntips == 3 ? 5 : 1;
    for (int x = 0; x < 5; ++x)
        for (int y = 0; y < 5; ++y)
            for (int z = 0; z < nz; ++z) {
                double u[4], v[4], p[4], m[4];
                if (ntips == 2) {
                    table_col(P0, x, u);
                    table_col(P1, y, v);
                } else if (shape == 0) {  // ((t0, t1), t2): the inner cherry's message, then the third tip's column
                    double cu[4], cv[4], cp[4];
                    table_col(P0, x, cu);
                    table_col(P1, y, cv);
#pragma unroll
                    for (int s = 0; s < 4; ++s) cp[s] = cu[s] * cv[s];
                    if (!a.norescale) table_rescale(cp);
                    matvec(Pc, cp, u);
                    table_col(P2, z, v);
                } else {                  // (t0, (t1, t2))
                    double cu[4], cv[4], cp[4];
                    table_col(P1, y, cu);
                    table_col(P2, z, cv);
#pragma unroll
                    for (int s = 0; s < 4; ++s) cp[s] = cu[s] * cv[s];
                    if (!a.norescale) table_rescale(cp);
                    table_col(P0, x, u);
                    matvec(Pc, cp, v);
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) p[s] = u[s] * v[s];
                if (!a.norescale) table_rescale(p);
                matvec(Pn, p, m);
                double2* o2 = reinterpret_cast<double2*>(out + 4 * ((5 * x + y) * nz + z));
                o2[0] = make_double2(m[0], m[1]);
                o2[1] = make_double2(m[2], m[3]);
            }
}

// ------------------------------------------------------------------------------------------
// K4: contraction of the branch statistics
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ void ldg_mat(const double* __restrict__ M, double (&m)[16]) {
    const double2* q = reinterpret_cast<const double2*>(M);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double2 v = __ldg(q + i);
        m[2 * i] = v.x;
        m[2 * i + 1] = v.y;
    }
}
__device__ __forceinline__ void ldg_pmat(const unsigned char* p, double (&m)[16], double) {
    ldg_mat(reinterpret_cast<const double*>(p), m);
}
__device__ __forceinline__ void ldg_pmat(const unsigned char* p, double (&m)[16], float) {
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 v = __ldg(q + i);
        m[4 * i] = v.x; m[4 * i + 1] = v.y; m[4 * i + 2] = v.z; m[4 * i + 3] = v.w;
    }
}

// One thread per (branch, category); sums over categories / branches go through warp shuffles and
// a few atomics per warp.
template <typename T>
__global__ void __launch_bounds__(128) contract_kernel(const ContractArgs a) {
    const int d = blockIdx.y;
    const int t_ = blockIdx.x * blockDim.x + threadIdx.x;
    const int C = a.C;
    const int b = t_ / C, c = t_ - b * C;
    const int lane = threadIdx.x & 31;
    const bool live = b < a.bcount;
    const double* prm = a.params + (size_t)d * a.lay.stride;
    double* od = a.out + (size_t)d * a.nout;
    double g = 0.0, tb = 0.0, r = 0.0;
    double th[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) th[k] = 0.0;
    if (live) {
        tb = prm[a.lay.off_t + b];
        r = prm[a.lay.off_rs + c];
        const int pos = a.node_pos[b];
        double G[16], P[16];
        ldg_mat(a.G + (((size_t)d * a.nn + b) * C + c) * 16, G);
        const unsigned char* rec = a.spost + (((size_t)d * C + c) * a.nsteps + (pos >> 1)) * Real<T>::kRec;
        ldg_pmat(rec + 64 + Real<T>::kMat * (pos & 1), P, T());
        if (a.tips_simple && b < a.S) {  // tip records hold P column-major: transpose back
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = i + 1; j < 4; ++j) {
                    const double t = P[4 * i + j];
                    P[4 * i + j] = P[4 * j + i];
                    P[4 * j + i] = t;
                }
        }
        // d logL / d tau = <G, Q P>   (dP/dtau = Q P)
        const double* Q = prm + a.lay.off_Q;
        if (a.jc_scalar) {  // the sweep already contracted with (J/4 - P): entry 0 holds <G, Q P> / mu, mu = -4/3 Q_00
            g = G[0] * (-Q[0] * (4.0 / 3.0));
        } else if (a.msg) {  // G holds G~ = G P^T, and <G, Q P> = tr(P^-1 G~^T Q P) = <G~, Q> because Q and P commute
#pragma unroll
            for (int i = 0; i < 16; ++i) g = fma(G[i], Q[i], g);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double qp = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) qp = fma(Q[4 * i + k], P[4 * k + j], qp);
                    g = fma(G[4 * i + j], qp, g);
                }
        }
        if (a.lay.ntheta > 0) {
            // H = m1^T G m2^T ; d logL/dtheta += sum_ij H_ij F_ij X_ij
            const double* m1 = prm + a.lay.off_m1;
            const double* m2 = prm + a.lay.off_m2;
            double lam[4], ex[4], Tm[16], H[16];
            const double tau = tb * r;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                lam[i] = prm[a.lay.off_lam + i];
                ex[i] = exp(lam[i] * tau);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) s = fma(m1[4 * k + i], G[4 * k + j], s);
                    Tm[4 * i + j] = s;
                }
            // F_ij = (e^{l_i tau} - e^{l_j tau}) / (l_i - l_j), factored around the LARGER exponential so that
            // expm1 only ever sees a non-positive argument (long branches: 0 * inf otherwise).  F is
            // symmetric: six expm1 for the off-diagonal pairs, F_ii = tau e^{l_i tau}.
            double F[16];
            if (a.msg) {
                // message statistic: <G, dP> = <G~, dP P^-1> and dP P^-1 = m1 (X o Phi) m2 with
                // Phi_ij = F_ij e^{-l_j tau} = (e^{(l_i - l_j) tau} - 1) / (l_i - l_j), Phi_ii = tau (not symmetric)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double x = (lam[i] - lam[j]) * tau;
                        F[4 * i + j] = i == j ? tau : tau * (fabs(x) < 1e-8 ? 1.0 + 0.5 * x : expm1(x) / x);
                    }
            } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                F[5 * i] = tau * ex[i];
#pragma unroll
                for (int j = i + 1; j < 4; ++j) {
                    const double x = fabs(lam[i] - lam[j]) * tau;
                    F[4 * i + j] = F[4 * j + i] = tau * fmax(ex[i], ex[j]) * (x < 1e-8 ? 1.0 - 0.5 * x : -expm1(-x) / x);
                }
            }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) s = fma(Tm[4 * i + k], m2[4 * j + k], s);
                    H[4 * i + j] = s * F[4 * i + j];
                }
#pragma unroll
            for (int k = 0; k < 10; ++k)
                if (k < a.lay.ntheta) {
                    const double* X = prm + a.lay.off_X + 16 * k;
                    double s = 0.0;
#pragma unroll
                    for (int i = 0; i < 16; ++i) s = fma(H[i], X[i], s);
                    th[k] = s;
                }
        }
    }
    // d/dblens[b] = sum_c r_c g_bc ; d/drs[c] = sum_b t_b g_bc
    double dt = r * g, dr = tb * g;
    if ((C & (C - 1)) == 0 && C <= 32) {  // categories of one branch sit in adjacent lanes
        for (int o = 1; o < C; o <<= 1) dt += __shfl_xor_sync(0xffffffffu, dt, o);
        for (int o = C; o < 32; o <<= 1) dr += __shfl_xor_sync(0xffffffffu, dr, o);
        if (live && c == 0) od[1 + b] = dt;
        if (lane < C && dr != 0.0) atomicAdd(od + a.off_out_rs + lane, dr);
    } else {
        if (live) {
            atomicAdd(od + 1 + b, dt);
            atomicAdd(od + a.off_out_rs + c, dr);
        }
    }
#pragma unroll
    for (int k = 0; k < 10; ++k)
        if (k < a.lay.ntheta) {
            const double s = warp_sum(th[k]);
            if (lane == 0 && s != 0.0)
                atomicAdd(od + (k < a.nsubst ? a.off_out_subst + k : a.off_out_freqs + (k - a.nsubst)), s);
        }
}

// ------------------------------------------------------------------------------------------
// K0: device-resident tip data -> the handle's padded code rows (phylo_b200_create_device)
// ------------------------------------------------------------------------------------------

// [S][L] 4-bit masks -> [S][Lpad] masks (padding = all ones); flags[0] is raised when some cell is neither
// one-hot nor all ones, i.e. when the column-index fast path does not apply
__global__ void __launch_bounds__(256) tips_pad_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                        int L, int Lpad, size_t total, int* __restrict__ flags) {
    bool odd = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t s = i / (size_t)Lpad;
        const int l = (int)(i - s * (size_t)Lpad);
        const uint8_t m = l < L ? (uint8_t)(src[s * (size_t)L + l] & 0xF) : (uint8_t)0xF;
        dst[i] = m;
        odd |= !(m == 15 || m == 1 || m == 2 || m == 4 || m == 8);
    }
    if (__any_sync(0xffffffffu, odd) && (threadIdx.x & 31) == 0) atomicOr(flags, 1);
}

// masks -> column indices 0..3 / 4 (all ones), in place
__global__ void __launch_bounds__(256) tips_index_kernel(uint8_t* __restrict__ t, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t m = t[i];
        t[i] = m == 15 ? 4 : (m == 1 ? 0 : (m == 2 ? 1 : (m == 4 ? 2 : 3)));
    }
}

// [S][Lpad] code rows -> [ntiles][S][T] in consumption order: slot s of every tile holds the T codes of tip order[s]
__global__ void __launch_bounds__(256) tips_reorder_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                            const int32_t* __restrict__ order, int S, int Lpad, int T,
                                                            size_t total) {
    // one thread per 4 bytes (T is a multiple of 32, Lpad of 512)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total / 4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i * 4;
        const int j = (int)(b % T);
        const size_t ts = b / T;
        const int s = (int)(ts % S);
        const size_t tile = ts / S;
        *reinterpret_cast<uint32_t*>(dst + b) =
            *reinterpret_cast<const uint32_t*>(src + (size_t)order[s] * Lpad + tile * T + j);
    }
}

// weights [L] (or NULL: ones) -> [Lpad] (padding = 0); flags[1] is raised on a non-finite weight
__global__ void __launch_bounds__(256) weights_pad_kernel(const double* __restrict__ w, double* __restrict__ dst, int L,
                                                           int Lpad, int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Lpad) return;
    const double v = i < L ? (w ? w[i] : 1.0) : 0.0;
    dst[i] = v;
    if (!isfinite(v)) atomicOr(flags + 1, 1);
}

// ------------------------------------------------------------------------------------------
// K6: clock-tree front end -- heights / ratios + rates -> branch lengths, and the chain rule back
// ------------------------------------------------------------------------------------------

// rate multiplier of the branch in pre-order row j and the (up to two) rate slots it reads
// (generate_script.py:660-679: strict / per-branch; :682-708: mean of the rates at the two ends, the root's
// rate being substrates[map[2,1]] and the branch above map[2,1] using its own rate alone)
__device__ __forceinline__ void clock_slots(const ClockArgs& a, int j, int node, int par, int& s0, int& s1) {
    if (!a.autocorr) { s0 = a.nrates == 1 ? 0 : node - 1; s1 = -1; return; }
    s0 = node - 1;
    s1 = j == 1 ? -1 : (par == a.nn ? a.map[2] - 1 : par - 1);
}

// One CTA per draw.  The ratio transform is a recursion down the tree (a node's height needs its parent's):
// thread 0 walks the pre-order rows; everything else is one thread per branch.
__global__ void __launch_bounds__(128) clock_forward_kernel(const ClockArgs a) {
    const int b = blockIdx.x, S = a.S, nn = a.nn;
    const double* in = a.in + (size_t)b * a.in_ld;
    double* h = a.hwork + (size_t)b * 2 * (S - 1);
    double* ho = a.hout + (size_t)b * a.hout_ld;
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    if (a.ratios) {
        if (threadIdx.x == 0) {  // generate_script.py:711-735 and its log-Jacobian :738-752
            const double* p = in;
            h[a.map[0] - S - 1] = in[S - 2];  // root height follows the S-2 proportions
            double lj = 0.0;
            int k = 0;
            for (int j = 1; j < nn; ++j) {
                const int node = a.map[2 * j], par = a.map[2 * j + 1];
                if (node <= S) continue;
                const double lo = a.lowers ? a.lowers[node - 1] : 0.0, span = h[par - S - 1] - lo;
                h[node - S - 1] = lo + span * p[k++];
                lj += log(span);
            }
            ho[(S - 1) + a.nrates + (S - 2) + 1] = lj;
        }
    } else {
        for (int k = threadIdx.x; k < S - 1; k += blockDim.x) h[k] = in[k];
    }
    __syncthreads();
    const double* rates = in + (S - 1);
    double* t = a.params + (size_t)b * a.stride + a.off_t;
    for (int j = 1 + threadIdx.x; j < nn; j += blockDim.x) {
        const int node = a.map[2 * j], par = a.map[2 * j + 1];
        int s0, s1;
        clock_slots(a, j, node, par, s0, s1);
        const double r = s1 < 0 ? rates[s0] : 0.5 * (rates[s0] + rates[s1]);
        const double lo = node > S ? h[node - S - 1] : (a.lowers ? a.lowers[node - 1] : 0.0);
        const double bl = r * (h[par - S - 1] - lo);
        t[node - 1] = bl;
        if (!(bl >= 0.0) || !isfinite(bl)) bad = 1;  // what pack_draw rejects on the host path
    }
    __syncthreads();
    if (threadIdx.x == 0) ho[a.hout_ld - 1] = bad ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(128) clock_reverse_kernel(const ClockArgs a) {
    const int b = blockIdx.x, S = a.S, nn = a.nn;
    const double* in = a.in + (size_t)b * a.in_ld;
    const double* rates = in + (S - 1);
    const double* h = a.hwork + (size_t)b * 2 * (S - 1);
    double* hb = a.hwork + (size_t)b * 2 * (S - 1) + (S - 1);
    double* ho = a.hout + (size_t)b * a.hout_ld;
    const double* gbl = a.out + (size_t)b * a.nout + 1;  // d/dblens[node - 1]
    double* g_rates = ho + (S - 1);
    // r g and span g of the branch in pre-order row j
    auto branch = [&](int j, double& rg, double& sg, int& s0, int& s1) {
        const int node = a.map[2 * j], par = a.map[2 * j + 1];
        clock_slots(a, j, node, par, s0, s1);
        const double r = s1 < 0 ? rates[s0] : 0.5 * (rates[s0] + rates[s1]);
        const double lo = node > S ? h[node - S - 1] : (a.lowers ? a.lowers[node - 1] : 0.0);
        const double g = gbl[node - 1];
        rg = r * g;
        sg = (h[par - S - 1] - lo) * g;
    };
    // d/dheights: an internal node gathers + r g from the branches of its two children and - r g from its own
    for (int k = threadIdx.x; k < S - 1; k += blockDim.x) {
        double rg, sg, acc = 0.0;
        int s0, s1;
        branch(a.kids[2 * k], rg, sg, s0, s1); acc += rg;
        branch(a.kids[2 * k + 1], rg, sg, s0, s1); acc += rg;
        const int own = a.row_of[S + k];
        if (own > 0) { branch(own, rg, sg, s0, s1); acc -= rg; }
        ho[k] = acc;
        hb[k] = acc + (a.has_extra ? in[(S - 1) + a.nrates + k] : 0.0);
    }
    // d/drates
    if (a.nrates == 1) {  // strict clock: sum over branches (fixed tree order -> reproducible)
        __shared__ double part[128];
        double acc = 0.0;
        for (int j = 1 + threadIdx.x; j < nn; j += blockDim.x) {
            double rg, sg;
            int s0, s1;
            branch(j, rg, sg, s0, s1);
            acc += sg;
        }
        part[threadIdx.x] = acc;
        __syncthreads();
        for (int o = 64; o > 0; o >>= 1) {
            if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) g_rates[0] = part[0];
    } else if (!a.autocorr) {
        for (int j = 1 + threadIdx.x; j < nn; j += blockDim.x) {
            double rg, sg;
            int s0, s1;
            branch(j, rg, sg, s0, s1);
            g_rates[s0] = sg;
        }
    } else {  // slot m-1 gathers its own branch and half of each child's; the root's stand-in gathers the root's children
        const int first = a.map[2];
        for (int m = 1 + threadIdx.x; m < nn; m += blockDim.x) {  // node m, slot m - 1
            double rg, sg, acc = 0.0;
            int s0, s1;
            const int own = a.row_of[m - 1];
            branch(own, rg, sg, s0, s1);
            acc += s1 < 0 ? sg : 0.5 * sg;
            if (m > S) {
                for (int c = 0; c < 2; ++c) {
                    const int jr = a.kids[2 * (m - S - 1) + c];
                    branch(jr, rg, sg, s0, s1);
                    if (s1 == m - 1) acc += 0.5 * sg;
                }
            }
            if (m == first) {
                for (int c = 0; c < 2; ++c) {
                    const int jr = a.kids[2 * (nn - S - 1) + c];
                    branch(jr, rg, sg, s0, s1);
                    if (s1 == m - 1) acc += 0.5 * sg;
                }
            }
            g_rates[m - 1] = acc;
        }
    }
    if (!a.ratios) return;
    __syncthreads();
    if (threadIdx.x == 0) {  // reverse sweep of the ratio transform and of its log-Jacobian: children before parents
        const double* p = in;
        double* gp = ho + (S - 1) + a.nrates;
        int k = S - 2;
        for (int j = nn - 1; j >= 1; --j) {
            const int node = a.map[2 * j], par = a.map[2 * j + 1];
            if (node <= S) continue;
            --k;
            const double lo = a.lowers ? a.lowers[node - 1] : 0.0, span = h[par - S - 1] - lo;
            const double nb = hb[node - S - 1];
            gp[k] = nb * span;
            hb[par - S - 1] += nb * p[k] + 1.0 / span;
        }
        gp[S - 2] = hb[a.map[0] - S - 1];  // d/droot_height
    }
}

// ------------------------------------------------------------------------------------------
// K5: multi-device handles -- add the other shards' result rows (peer memory) to this device's
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) peer_sum_kernel(double* __restrict__ out, const PeerRows peers, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        double s = out[i];
        for (int p = 0; p < peers.n; ++p) s += peers.src[p][i];  // shard order: the sum is reproducible
        out[i] = s;
    }
}

#endif  // part 0

// ------------------------------------------------------------------------------------------
// launch plumbing
// ------------------------------------------------------------------------------------------

// Specialised for 128-thread CTAs (one pattern block of 4 categories, the common case) with a
// per-K register budget; any other CTA size takes the generic variant.
template <typename T, int K> struct Cfg;
template <> struct Cfg<double, 1> { static constexpr int minb = 4; };
template <> struct Cfg<double, 2> { static constexpr int minb = 3; };
template <> struct Cfg<double, 4> { static constexpr int minb = 1; };
template <> struct Cfg<float, 1> { static constexpr int minb = 6; };
template <> struct Cfg<float, 2> { static constexpr int minb = 5; };
template <> struct Cfg<float, 4> { static constexpr int minb = 3; };

typedef void (*SweepFn)(const SweepArgs);

// 128-thread CTAs of simple-tip handles take their tip codes through the TipRing (PHYLO_TIPRING)
template <typename T, int K, bool GRAD, bool TIPS, bool DEEP, bool JC = false>
SweepFn pick_kernel(int nthreads) {
    if (nthreads == 128)
        return sweep_kernel<T, K, GRAD, TIPS, 128, Cfg<T, K>::minb, DEEP, JC, TIPS && (PHYLO_TIPRING == 2 || (PHYLO_TIPRING == 1 && !GRAD))>;
    return sweep_kernel<T, K, GRAD, TIPS, 0, 1, DEEP, JC, false>;
}

// message-statistic gradient kernels (fp64, simple tips, 128-thread CTAs); ch: with cherry tables (K = 4 only)
template <bool DEEP>
SweepFn pick_kernel_msg(int K, bool ch) {
    if (ch) return K == 4 ? sweep_kernel<double, 4, true, true, 128, Cfg<double, 4>::minb, DEEP, false, false, true, true> : nullptr;
    switch (K) {
#if !defined(PHYLO_FAST_BUILD) || PHYLO_FAST_BUILD < 2
        case 1: return sweep_kernel<double, 1, true, true, 128, Cfg<double, 1>::minb, DEEP, false, PHYLO_TIPRING == 2, true>;
        case 2: return sweep_kernel<double, 2, true, true, 128, Cfg<double, 2>::minb, DEEP, false, PHYLO_TIPRING == 2, true>;
#endif
        case 4: return sweep_kernel<double, 4, true, true, 128, Cfg<double, 4>::minb, DEEP, false, PHYLO_TIPRING == 2, true>;
    }
    return nullptr;
}

// scalar-statistic gradient kernels (JC69, fp64, whole stack on chip)
template <bool TIPS>
SweepFn pick_kernel_jc(int K, int nthreads) {
    switch (K) {
        case 1: return pick_kernel<double, 1, true, TIPS, false, true>(nthreads);
        case 2: return pick_kernel<double, 2, true, TIPS, false, true>(nthreads);
        case 4: return pick_kernel<double, 4, true, TIPS, false, true>(nthreads);
    }
    return nullptr;
}

template <typename T, bool TIPS>
SweepFn pick_kernel_k(int K, bool grad, bool deep, int nthreads) {
    switch (K * 4 + (grad ? (deep ? 2 : 1) : 0)) {
        case 4: return pick_kernel<T, 1, false, TIPS, false>(nthreads);
        case 5: return pick_kernel<T, 1, true, TIPS, false>(nthreads);
        case 6: return pick_kernel<T, 1, true, TIPS, true>(nthreads);
        case 8: return pick_kernel<T, 2, false, TIPS, false>(nthreads);
        case 9: return pick_kernel<T, 2, true, TIPS, false>(nthreads);
        case 10: return pick_kernel<T, 2, true, TIPS, true>(nthreads);
        case 16: return pick_kernel<T, 4, false, TIPS, false>(nthreads);
        case 17: return pick_kernel<T, 4, true, TIPS, false>(nthreads);
        case 18: return pick_kernel<T, 4, true, TIPS, true>(nthreads);
    }
    return nullptr;
}

#ifndef PHYLO_PART
SweepFn pick(int prec, bool tips, int K, bool grad, bool deep, int nthreads, bool jc = false, bool msg = false,
             bool ch = false) {
    if (msg)
        return prec == 64 && tips && grad && !jc && nthreads == 128
                   ? (deep ? pick_kernel_msg<true>(K, ch) : pick_kernel_msg<false>(K, ch))
                   : nullptr;
#ifdef PHYLO_FAST_BUILD  // compile-time experiments only: the fp64 K = 4 / 2 gradient kernels of simple-tip handles
    if (prec != 64 || !tips || !grad || jc || nthreads != 128) return nullptr;
    if (K == 4) return deep ? pick_kernel<double, 4, true, true, true>(128) : pick_kernel<double, 4, true, true, false>(128);
#if PHYLO_FAST_BUILD < 2
    if (K == 2) return deep ? pick_kernel<double, 2, true, true, true>(128) : pick_kernel<double, 2, true, true, false>(128);
#endif
    return nullptr;
#else
    if (jc && prec == 64 && grad && !deep) return tips ? pick_kernel_jc<true>(K, nthreads) : pick_kernel_jc<false>(K, nthreads);
    if (prec == 32)
        return tips ? pick_kernel_k<float, true>(K, grad, deep, nthreads) : pick_kernel_k<float, false>(K, grad, deep, nthreads);
    return tips ? pick_kernel_k<double, true>(K, grad, deep, nthreads) : pick_kernel_k<double, false>(K, grad, deep, nthreads);
#endif
}

#endif

}  // namespace

#ifdef PHYLO_PART
// the parts' share of the instantiations, one external function each (kernel stubs are ordinary host functions: a
// pointer obtained in one translation unit launches, and is queried, from another)
typedef void (*SweepFnX)(const SweepArgs);
SweepFnX pick_part_f64_tips(int K, bool grad, bool deep, int nthreads);
SweepFnX pick_part_f64_masks(int K, bool grad, bool deep, int nthreads);
SweepFnX pick_part_f32(bool tips, int K, bool grad, bool deep, int nthreads);
SweepFnX pick_part_special(bool tips, int K, bool deep, int nthreads, bool jc, bool msg, bool ch);
#if PHYLO_PART == 1
SweepFnX pick_part_f64_tips(int K, bool grad, bool deep, int nthreads) { return pick_kernel_k<double, true>(K, grad, deep, nthreads); }
#elif PHYLO_PART == 2
SweepFnX pick_part_f64_masks(int K, bool grad, bool deep, int nthreads) { return pick_kernel_k<double, false>(K, grad, deep, nthreads); }
#elif PHYLO_PART == 3
SweepFnX pick_part_f32(bool tips, int K, bool grad, bool deep, int nthreads) {
    return tips ? pick_kernel_k<float, true>(K, grad, deep, nthreads) : pick_kernel_k<float, false>(K, grad, deep, nthreads);
}
#elif PHYLO_PART == 4
SweepFnX pick_part_special(bool tips, int K, bool deep, int nthreads, bool jc, bool msg, bool ch) {
    if (msg) return deep ? pick_kernel_msg<true>(K, ch) : pick_kernel_msg<false>(K, ch);
    if (jc) return tips ? pick_kernel_jc<true>(K, nthreads) : pick_kernel_jc<false>(K, nthreads);
    return nullptr;
}
#endif
#endif

#if !defined(PHYLO_PART) || PHYLO_PART == 0
namespace {
#ifdef PHYLO_PART
SweepFn pick(int prec, bool tips, int K, bool grad, bool deep, int nthreads, bool jc = false, bool msg = false,
             bool ch = false) {
    if (msg) return prec == 64 && tips && grad && !jc && nthreads == 128 ? pick_part_special(tips, K, deep, nthreads, false, true, ch) : nullptr;
    if (jc && prec == 64 && grad && !deep) return pick_part_special(tips, K, deep, nthreads, true, false, false);
    if (prec == 32) return pick_part_f32(tips, K, grad, deep, nthreads);
    return tips ? pick_part_f64_tips(K, grad, deep, nthreads) : pick_part_f64_masks(K, grad, deep, nthreads);
}
#endif
}  // namespace

int sweep_max_threads(int) { return 512; }

bool sweep_uses_tipring(bool tips, int nthreads, bool grad) {
    return tips && nthreads == 128 && (PHYLO_TIPRING == 2 || (PHYLO_TIPRING == 1 && !grad));
}

int record_bytes(int prec) { return prec == 32 ? kRecBytesF32 : kRecBytes; }

size_t sweep_stack_bytes(int D, int K, int nthreads, int prec) {
    const size_t entry = prec == 32 ? 16 : 32;  // bytes per 4-state vector
    // the root's category exchange (a double and an int per thread and pattern) borrows the stack region
    return std::max((size_t)D * K * nthreads * entry, (size_t)K * nthreads * (sizeof(double) + sizeof(int)));
}

// [stack | record rings | byte rings | reduction rows], all per warp except the stack
size_t sweep_smem_bytes(int D, int K, int nthreads, int prec, bool jc, bool tips, bool grad) {
    const size_t val = prec == 32 ? 4 : 8;
    const size_t warps = nthreads / 32;
    size_t red = 0;
    if (K == 1) red = warps * 32 * 33 * val;                       // both children in one pass
    else if (PHYLO_RSM && !jc) red = warps * 16 * 34 * val;        // one child at a time
    return sweep_stack_bytes(D, K, nthreads, prec) + warps * record_bytes(prec) * kRecChunk * kRecBufs +
           (sweep_uses_tipring(tips, nthreads, grad) ? warps * (size_t)(12 * 32 * K) : 0) + red;
}

void launch_stream(const StreamArgs& a, int prec, cudaStream_t stream) {
    const int total = 2 * a.B * a.lay.C * (a.npost + a.nsteps);
    if (prec == 32) stream_kernel<float><<<(total + 127) / 128, 128, 0, stream>>>(a);
    else stream_kernel<double><<<(total + 127) / 128, 128, 0, stream>>>(a);
}

bool sweep_msg_available(int prec, bool tips, bool grad, bool, int nthreads, bool jc) {
    return prec == 64 && tips && grad && !jc && nthreads == 128;
}

bool sweep_cherry_available(int K) { return K == 4 && PHYLO_TIPRING != 2; }

cudaError_t launch_sweep(const SweepArgs& a, int prec, bool tips, int K, bool grad, bool deep, int grid, int nthreads,
                         size_t smem, cudaStream_t stream, bool jc, bool msg, bool ch) {
    SweepFn kern = pick(prec, tips, K, grad, deep, nthreads, jc, msg, ch);
    if (!kern) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, nthreads, smem, stream>>>(a);
    return cudaGetLastError();
}

// Resident CTAs per SM from registers and shared memory alone.  The occupancy calculator answers 1 for every kernel
// that allocates tensor memory (it cannot know how many of the 512 columns the kernel asks for); the kernels here take
// 64 (the statistics' warp sums) or 160 / 256 (the TMEM stack), and the hardware co-schedules them accordingly.
static cudaError_t occupancy_by_hand(const void* kern, int nthreads, size_t smem, int* n) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    int dev = 0, regs_sm = 0, smem_sm = 0, thr_sm = 0, reserved = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&thr_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
    cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, dev);
    const int warps = (nthreads + 31) / 32;
    const int regs_warp = ((fa.numRegs * 32 + 255) / 256) * 256;  // allocated per warp in units of 256
    const int by_regs = regs_sm / std::max(1, regs_warp * warps);
    const int by_smem = (int)(smem_sm / (smem + fa.sharedSizeBytes + reserved));
    *n = std::max(0, std::min(std::min(by_regs, by_smem), thr_sm / nthreads));
    return cudaSuccess;
}

cudaError_t sweep_occupancy(int prec, bool tips, int K, bool grad, bool deep, int nthreads, size_t smem, int* n,
                            bool jc, bool msg, bool ch) {
    SweepFn kern = pick(prec, tips, K, grad, deep, nthreads, jc, msg, ch);
    if (!kern) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (PHYLO_TMRED && grad && !jc && K > 1 && nthreads == 128 && prec == 64)  // the kernels that allocate tensor memory
        return occupancy_by_hand(reinterpret_cast<const void*>(kern), nthreads, smem, n);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, kern, nthreads, smem);
}

void launch_tips_prepare(const uint8_t* d_src, uint8_t* d_dst, int S, int L, int Lpad, const double* d_w, double* d_wdst,
                         int* d_flags, cudaStream_t stream) {
    const size_t total = (size_t)S * Lpad;
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 32);
    tips_pad_kernel<<<grid, 256, 0, stream>>>(d_src, d_dst, L, Lpad, total, d_flags);
    weights_pad_kernel<<<(Lpad + 255) / 256, 256, 0, stream>>>(d_w, d_wdst, L, Lpad, d_flags);
}

void launch_tips_reorder(const uint8_t* d_tips, uint8_t* d_dst, const int32_t* d_order, int S, int Lpad, int T, int ntiles,
                         cudaStream_t stream) {
    const size_t total = (size_t)ntiles * S * T;
    const int grid = (int)std::min<size_t>((total / 4 + 255) / 256, (size_t)148 * 32);
    tips_reorder_kernel<<<grid, 256, 0, stream>>>(d_tips, d_dst, d_order, S, Lpad, T, total);
}

void launch_tips_index(uint8_t* d_tips, int S, int Lpad, cudaStream_t stream) {
    const size_t total = (size_t)S * Lpad;
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 32);
    tips_index_kernel<<<grid, 256, 0, stream>>>(d_tips, total);
}

void launch_clock_forward(const ClockArgs& a, cudaStream_t stream) { clock_forward_kernel<<<a.B, 128, 0, stream>>>(a); }
void launch_clock_reverse(const ClockArgs& a, cudaStream_t stream) { clock_reverse_kernel<<<a.B, 128, 0, stream>>>(a); }

void launch_peer_sum(double* out, const PeerRows& peers, size_t count, cudaStream_t stream) {
    const int grid = (int)std::min<size_t>((count + 255) / 256, 148);
    peer_sum_kernel<<<grid, 256, 0, stream>>>(out, peers, count);
}

typedef void (*SweepTmFn)(const SweepArgs);
static SweepTmFn pick_tm(bool tips, int K, int ctas, bool msg) {
    if (K != 4 || (msg && !tips)) return nullptr;
#ifdef PHYLO_FAST_BUILD  // compile-time experiments only
    if (ctas == 3 && tips) return msg ? sweep_tm_kernel<4, true, 3, true> : sweep_tm_kernel<4, true, 3>;
    return nullptr;
#else
    if (msg) return ctas == 3 ? sweep_tm_kernel<4, true, 3, true> : ctas == 2 ? sweep_tm_kernel<4, true, 2, true> : nullptr;
    if (ctas == 3) return tips ? sweep_tm_kernel<4, true, 3> : sweep_tm_kernel<4, false, 3>;
    if (ctas == 2) return tips ? sweep_tm_kernel<4, true, 2> : sweep_tm_kernel<4, false, 2>;
    return nullptr;
#endif
}

bool sweep_tm_available(int prec, int K, int nthreads, bool grad, bool jc) {
    return prec == 64 && K == 4 && nthreads == 128 && grad && !jc;
}

int sweep_tm_slots(int K, int ctas) { return (ctas == 3 ? 160 : 256) / (8 * K); }

size_t sweep_tm_smem_bytes(int K) {
    return (size_t)4 * (kTmOpSlots * (size_t)K * 2 * 32 * 16 + kRecBytes * kRecChunk * kRecBufs);
}

cudaError_t sweep_tm_prepare(bool tips, int K, int ctas, int* n, bool msg) {
    SweepTmFn kern = pick_tm(tips, K, ctas, msg);
    if (!kern) return cudaErrorInvalidValue;
    const size_t smem = sweep_tm_smem_bytes(K);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // the occupancy calculator does not know how many TMEM columns the kernel allocates and answers 1; what bounds
    // the residency is registers (launch bounds) and shared memory, both sized for `ctas`
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, kern, 128, smem);
    if (e == cudaSuccess && *n >= 1) *n = ctas;
    return e;
}

cudaError_t launch_sweep_tm(const SweepArgs& a, bool tips, int K, int ctas, int grid, cudaStream_t stream, bool msg) {
    SweepTmFn kern = pick_tm(tips, K, ctas, msg);
    if (!kern) return cudaErrorInvalidValue;
    const size_t smem = sweep_tm_smem_bytes(K);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, 128, smem, stream>>>(a);
    return cudaGetLastError();
}

void launch_cherry_tables(const CherryArgs& a, cudaStream_t stream) {
    const int total = a.B * a.C * a.ncherry;
    cherry_table_kernel<<<(total + 127) / 128, 128, 0, stream>>>(a);
}

void launch_cherry_codes(const uint8_t* d_tips, uint8_t* d_ctips, const int32_t* d_cherries, int ncherry, int Lpad,
                         cudaStream_t stream) {
    const size_t total = (size_t)ncherry * Lpad;
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 32);
    cherry_codes_kernel<<<grid, 256, 0, stream>>>(d_tips, d_ctips, d_cherries, Lpad, total);
}

void launch_contract(const ContractArgs& a, int prec, int B, cudaStream_t stream) {
    dim3 grid((a.bcount * a.C + 127) / 128, B);
    if (prec == 32) contract_kernel<float><<<grid, 128, 0, stream>>>(a);
    else contract_kernel<double><<<grid, 128, 0, stream>>>(a);
}

#endif  // part 0

}  // namespace phylo
