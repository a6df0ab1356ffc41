// Tensor memory as a per-thread scratchpad (no MMA anywhere): can the sweep kernel's stack of pending 4-state
// vectors live in TMEM instead of shared memory?  Each CTA of 128 threads allocates COLS_A (+ COLS_B) columns;
// thread t of warp w owns TMEM lane 32 (w % 4) + t and reads / writes 32-word entries ("slots": K = 4 patterns x
// 4 doubles) with tcgen05.st / tcgen05.ld .32x32b.  Checks: (1) contents survive arbitrary push / pop orders,
// (2) CTAS_PER_SM CTAs of one SM hold their allocations at the same time, (3) cycles per entry store / load
// against the same traffic through shared memory (LDS.128 / STS.128, conflict-free layout).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_stack_bench tmem_stack_bench.cu && ./tmem_stack_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#ifndef CTAS_PER_SM
#define CTAS_PER_SM 3
#endif
#if CTAS_PER_SM == 3
#define COLS_A 128
#define COLS_B 32
#define SMEM_BYTES (72 * 1024)
#else
#define COLS_A 256
#define COLS_B 0
#define SMEM_BYTES (100 * 1024)
#endif
constexpr int kSlots = (COLS_A + COLS_B) / 32;

__device__ __forceinline__ void tm_st8(uint32_t a, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(a), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tm_ld8(uint32_t a, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(a)
                 : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ uint32_t word(int tid, int ver, int k) { return (uint32_t)(tid * 2654435761u) ^ (uint32_t)(ver * 40503u + k * 97u + 1u); }

struct Out {
    unsigned long long errors, t_start, t_end, cyc_tm_st, cyc_tm_ld, cyc_sm_st, cyc_sm_ld, cyc_tm_rt, cyc_sm_rt;
    unsigned long long cyc_tm_sparse, cyc_sm_sparse;
    unsigned smid, baseA, baseB;
};

__global__ void __launch_bounds__(128, CTAS_PER_SM) bench(Out* out, int iters) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t tm_base[2];
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tm_base[0])),
                     "r"(COLS_A));
        if (COLS_B)
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(&tm_base[1])),
                         "r"(COLS_B));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t baseA = tm_base[0], baseB = COLS_B ? tm_base[1] : 0u;
    const uint32_t lane_off = (uint32_t)(32 * (warp & 3)) << 16;
    auto slot_addr = [&](int s) -> uint32_t {
        const int col = 32 * s;
        return (col < COLS_A ? baseA + col : baseB + (col - COLS_A)) + lane_off;
    };

    // (1) correctness: random overwrite / read-back of the slots
    unsigned long long errors = 0;
    int ver[kSlots];
#pragma unroll
    for (int s = 0; s < kSlots; ++s) ver[s] = -1;
    uint32_t rng = 12345u + blockIdx.x * 977u;  // warp-uniform choices (the instructions are .aligned)
    for (int it = 0; it < iters; ++it) {
        rng = rng * 1664525u + 1013904223u;
        const int s = (rng >> 8) % kSlots;
        rng = rng * 1664525u + 1013904223u;
        const int r = (rng >> 8) % kSlots;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = word(tid + 128 * blockIdx.x, it, 8 * j + k);
            tm_st8(slot_addr(s) + 8 * j, v);
        }
        tm_wait_st();
#pragma unroll
        for (int q = 0; q < kSlots; ++q)
            if (q == s) ver[q] = it;
        int rv = -1;
#pragma unroll
        for (int q = 0; q < kSlots; ++q)
            if (q == r) rv = ver[q];
        if (rv >= 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t v[8];
                tm_ld8(slot_addr(r) + 8 * j, v);
                tm_wait_ld();
#pragma unroll
                for (int k = 0; k < 8; ++k) errors += v[k] != word(tid + 128 * blockIdx.x, rv, 8 * j + k);
            }
        }
    }

    // (3) cycles: 64 entry stores, 64 entry loads (dependent through an accumulator), 64 store->load round trips
    uint32_t acc[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    long long c0 = clock64();
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) tm_st8(slot_addr(it % kSlots) + 8 * j, acc);
        tm_wait_st();
        acc[0] += it;
    }
    long long c1 = clock64();
    for (int it = 0; it < 64; ++it) {
        uint32_t v[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) tm_ld8(slot_addr(it % kSlots) + 8 * j, v[j]);
        tm_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += v[j][k];
    }
    long long c2 = clock64();
    for (int it = 0; it < 64; ++it) {  // round trip: the load depends on the store just made
#pragma unroll
        for (int j = 0; j < 4; ++j) tm_st8(slot_addr(it % kSlots) + 8 * j, acc);
        tm_wait_st();
        uint32_t v[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) tm_ld8(slot_addr(it % kSlots) + 8 * j, v[j]);
        tm_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += v[j][k];
    }
    long long c3 = clock64();
    // the same through shared memory: [slot][8 x 16 B][128 threads]
    uint4* sm = reinterpret_cast<uint4*>(smem);
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) sm[((it & 3) * 8 + q) * 128 + tid] = make_uint4(acc[0], acc[1], acc[2], acc[3] + q);
        acc[0] += it;
    }
    long long c4 = clock64();
    for (int it = 0; it < 64; ++it) {
        uint4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = sm[((it & 3) * 8 + q) * 128 + tid];
#pragma unroll
        for (int q = 0; q < 8; ++q) { acc[0] += v[q].x; acc[1] += v[q].y; acc[2] += v[q].z; acc[3] += v[q].w; }
    }
    long long c5 = clock64();
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) sm[((it & 3) * 8 + q) * 128 + tid] = make_uint4(acc[0], acc[1], acc[2], acc[3] + q);
        uint4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = sm[((it & 3) * 8 + q) * 128 + tid];
#pragma unroll
        for (int q = 0; q < 8; ++q) { acc[0] += v[q].x; acc[1] += v[q].y; acc[2] += v[q].z; acc[3] += v[q].w; }
    }
    long long c6 = clock64();

    // (4) the sweep kernel's access pattern: one entry load every few thousand cycles, FP64 work in between
    double f[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) f[k] = 1.0 + 1e-9 * (tid + k);
    unsigned long long sparse_tm = 0, sparse_sm = 0;
    for (int it = 0; it < 128; ++it) {
#pragma unroll 1
        for (int r = 0; r < 24; ++r)
#pragma unroll
            for (int k = 0; k < 16; ++k) f[k] = fma(f[k], 1.0000001, 1e-12);
        long long a0 = clock64();
        uint32_t v[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) tm_ld8(slot_addr(it % kSlots) + 8 * j, v[j]);
        tm_wait_ld();
        uint32_t x = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) x += v[j][k];
        acc[1] += x;
        long long a1 = clock64();
        sparse_tm += (acc[1] == 0xdeadbeefu) ? 0 : (a1 - a0);
#pragma unroll 1
        for (int r = 0; r < 24; ++r)
#pragma unroll
            for (int k = 0; k < 16; ++k) f[k] = fma(f[k], 1.0000001, 1e-12);
        long long b0 = clock64();
        uint4 w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) w[q] = sm[((it & 3) * 8 + q) * 128 + tid];
        uint32_t y = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) y += w[q].x + w[q].y + w[q].z + w[q].w;
        acc[2] += y;
        long long b1 = clock64();
        sparse_sm += (acc[2] == 0xdeadbeefu) ? 0 : (b1 - b0);
    }
    if (f[0] + f[7] + f[15] == 12345.0) acc[3] += 1;

    // every CTA of the grid must hold its columns at the same time: spin until all have arrived
    __shared__ unsigned long long serr[4];
    for (int o = 16; o > 0; o >>= 1) errors += __shfl_xor_sync(0xffffffffu, errors, o);
    if ((tid & 31) == 0) serr[warp] = errors;
    __syncthreads();
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (tid == 0) {
        Out o;
        o.errors = serr[0] + serr[1] + serr[2] + serr[3];
        o.t_start = t0; o.t_end = t1;
        o.cyc_tm_st = c1 - c0; o.cyc_tm_ld = c2 - c1; o.cyc_tm_rt = c3 - c2;
        o.cyc_sm_st = c4 - c3; o.cyc_sm_ld = c5 - c4; o.cyc_sm_rt = c6 - c5;
        o.cyc_tm_sparse = sparse_tm; o.cyc_sm_sparse = sparse_sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(o.smid));
        o.baseA = baseA; o.baseB = baseB;
        out[blockIdx.x] = o;
        if (acc[0] + acc[5] == 0xdeadbeef) out[blockIdx.x].errors += 1u << 30;
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(baseA), "r"(COLS_A));
        if (COLS_B) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(baseB), "r"(COLS_B));
    }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bench, 128, SMEM_BYTES);
    const int grid = sms * CTAS_PER_SM;
    Out* d;
    cudaMalloc(&d, sizeof(Out) * grid);
    bench<<<grid, 128, SMEM_BYTES>>>(d, 20000);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<Out> h(grid);
    cudaMemcpy(h.data(), d, sizeof(Out) * grid, cudaMemcpyDeviceToHost);
    unsigned long long errors = 0, smax = 0, emin = ~0ull;
    double st = 0, ld = 0, rt = 0, sst = 0, sld = 0, srt = 0, sp_tm = 0, sp_sm = 0;
    std::vector<int> per_sm(sms, 0);
    for (auto& o : h) {
        errors += o.errors;
        smax = o.t_start > smax ? o.t_start : smax;
        emin = o.t_end < emin ? o.t_end : emin;
        st += o.cyc_tm_st; ld += o.cyc_tm_ld; rt += o.cyc_tm_rt; sst += o.cyc_sm_st; sld += o.cyc_sm_ld; srt += o.cyc_sm_rt;
        sp_tm += o.cyc_tm_sparse; sp_sm += o.cyc_sm_sparse;
        per_sm[o.smid % sms]++;
    }
    int maxper = 0;
    for (int v : per_sm) maxper = v > maxper ? v : maxper;
    printf("CTAs per SM requested %d, occupancy API %d, grid %d, slots per CTA %d (%d + %d columns)\n", CTAS_PER_SM, occ, grid, kSlots, COLS_A, COLS_B);
    printf("errors %llu; all CTAs alive together: %s (latest start %s earliest end); max CTAs seen on one SM %d\n", errors,
           smax < emin ? "yes" : "NO", smax < emin ? "<" : ">=", maxper);
    printf("first CTA: baseA 0x%08x baseB 0x%08x; second: 0x%08x 0x%08x; third: 0x%08x 0x%08x\n", h[0].baseA, h[0].baseB, h[1].baseA,
           h[1].baseB, h[2].baseA, h[2].baseB);
    const double n = 64.0 * grid;
    printf("cycles per 32-word entry and thread (all %d warps of the SM busy with the same loop):\n", 4 * CTAS_PER_SM);
    printf("  TMEM  store+wait %.1f   load+wait %.1f   store->load round trip %.1f\n", st / n, ld / n, rt / n);
    printf("  smem  store      %.1f   load      %.1f   store->load round trip %.1f\n", sst / n, sld / n, srt / n);
    printf("one entry load + use between blocks of 384 DFMA per thread (the sweep's access pattern), cycles: TMEM %.1f   smem %.1f\n",
           sp_tm / (128.0 * grid), sp_sm / (128.0 * grid));
    return errors ? 2 : 0;
}
