// Microbenchmark (B200): FP64 mma.sync throughput alone and interleaved with DFMA chains -- does the FP64 tensor
// path share the vector FP64 pipe?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                 "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
                   "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}

// MODE 0: DFMA only (NF independent chains); 1: m8n8k4 only (NM independent accumulators); 2: both; 3: m16n8k16; 4: m16n8k4
template <int MODE, int NF, int NM>
__global__ void __launch_bounds__(128) k(double* out, int iters, double x) {
    double f[NF > 0 ? NF : 1];
    double d0[NM > 0 ? NM : 1], d1[NM > 0 ? NM : 1];
    double d4[NM > 0 ? NM : 1][4];
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
    for (int i = 0; i < NM; ++i) { d0[i] = i; d1[i] = -i; for (int j = 0; j < 4; ++j) d4[i][j] = i + j; }
    double a = x + threadIdx.x, b = x - threadIdx.x;
    double a8[8], b4[4], a2[2];
    for (int j = 0; j < 8; ++j) a8[j] = a + j;
    for (int j = 0; j < 4; ++j) b4[j] = b + j;
    a2[0] = a; a2[1] = b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NF > NM ? NF : NM); ++i) {
            if ((MODE == 0 || MODE == 2) && i < NF) f[i] = fma(f[i], a, b);
            if ((MODE == 1 || MODE == 2) && i < NM) dmma884(d0[i], d1[i], a, b);
            if (MODE == 3 && i < NM) dmma16816(d4[i], a8, b4);
            if (MODE == 4 && i < NM) dmma1684(d4[i], a2, b);
        }
    }
    double s = 0;
    for (int i = 0; i < NF; ++i) s += f[i];
    for (int i = 0; i < NM; ++i) s += d0[i] + d1[i] + d4[i][0] + d4[i][1] + d4[i][2] + d4[i][3];
    if (s == 1.2345) out[0] = s;
}

template <int MODE, int NF, int NM>
void run(const char* name, int warps_per_sm) {
    double* out;
    cudaMalloc(&out, 8);
    const int iters = 20000;
    const int ctas = 148 * warps_per_sm / 4;
    k<MODE, NF, NM><<<ctas, 128>>>(out, 100, 1.0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE, NF, NM><<<ctas, 128>>>(out, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double nfma = (double)iters * NF * ctas * 4, nmma = (double)iters * NM * ctas * 4;
    printf("%-28s warps/SM %2d  %8.3f ms   DFMA warp-instr/s %.3e   MMA/s %.3e  (per SM per clk @1.9GHz: dfma %.3f  mma %.3f)\n", name,
           warps_per_sm, ms, (MODE == 0 || MODE == 2) ? nfma / ms * 1e3 : 0.0, MODE ? nmma / ms * 1e3 : 0.0,
           (MODE == 0 || MODE == 2) ? nfma / ms * 1e3 / 148 / 1.9e9 : 0.0, MODE ? nmma / ms * 1e3 / 148 / 1.9e9 : 0.0);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0, 8, 0>("dfma x8", w);
        run<1, 0, 8>("m8n8k4 x8", w);
        run<2, 8, 8>("dfma x8 + m8n8k4 x8", w);
        run<2, 8, 2>("dfma x8 + m8n8k4 x2", w);
        run<2, 8, 1>("dfma x8 + m8n8k4 x1", w);
        run<3, 0, 4>("m16n8k16 x4", w);
        run<4, 0, 8>("m16n8k4 x8", w);
    }
    // latency: one dependent chain, one warp per SM sub-partition
    run<0, 1, 0>("dfma chain (latency)", 4);
    run<1, 0, 1>("m8n8k4 chain (latency)", 4);
    run<4, 0, 1>("m16n8k4 chain (latency)", 4);
    run<3, 0, 1>("m16n8k16 chain (latency)", 4);
    return 0;
}
