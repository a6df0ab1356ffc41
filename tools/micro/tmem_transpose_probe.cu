// Which (TMEM lane, column) does each thread's register receive for the cross-lane tcgen05.ld shapes?  One warp writes
// lane l, column c := 1000 l + c with .32x32b (every thread its own lane), then reads the same columns back with
// .16x256b.x8 / .16x128b.x8 / .16x64b.x8 at lane bases 0 and 16, and prints what every thread got.  The sweep kernel
// wants to use this as a transpose: 32 lanes x 32 doubles in, each thread a few ENTRIES of many LANES out.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_transpose_probe tmem_transpose_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(128) probe(uint32_t* out) {
    __shared__ uint32_t base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t base = base_s + ((uint32_t)(32 * warp) << 16);
    // write: thread = lane, 64 columns
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 1000u * (32 * warp + lane) + (8 * g + k);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(base + 8 * g), "r"(v[0]),
                     "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n");
    // read back with the three cross-lane shapes, lane bases 0 and 16 of the warp's quadrant
    for (int half = 0; half < 2; ++half) {
        const uint32_t a = base + ((uint32_t)(16 * half) << 16);
        uint32_t r[32];
        // .16x256b.x8: 32 registers per thread, 64 columns
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
            "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(a));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n");
        for (int k = 0; k < 32; ++k) out[((0 * 2 + half) * 128 + tid) * 32 + k] = r[k];
        // .16x128b.x8: 16 registers per thread, 32 columns
        asm volatile(
            "tcgen05.ld.sync.aligned.16x128b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(a));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n");
        for (int k = 0; k < 32; ++k) out[((1 * 2 + half) * 128 + tid) * 32 + k] = k < 16 ? r[k] : 0xffffffffu;
        // .16x64b.x8: 8 registers per thread, 16 columns
        asm volatile("tcgen05.ld.sync.aligned.16x64b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(a));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n");
        for (int k = 0; k < 32; ++k) out[((2 * 2 + half) * 128 + tid) * 32 + k] = k < 8 ? r[k] : 0xffffffffu;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base_s), "r"(64));
}

int main() {
    uint32_t* d;
    const size_t n = 3 * 2 * 128 * 32;
    cudaMalloc(&d, n * 4);
    cudaMemset(d, 0xff, n * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    static uint32_t h[3 * 2 * 128 * 32];
    cudaMemcpy(h, d, n * 4, cudaMemcpyDeviceToHost);
    const char* names[3] = {"16x256b.x8", "16x128b.x8", "16x64b.x8"};
    const int nreg[3] = {32, 16, 8};
    for (int s = 0; s < 3; ++s)
        for (int half = 0; half < 2; ++half) {
            printf("== %s, lane base %d (warp 1 shown: its lanes are 32..63); entries are lane:column\n", names[s], 16 * half);
            for (int t = 32; t < 64; ++t) {
                if (t > 40 && t < 60 && t != 48) continue;
                printf("  thread %2d:", t - 32);
                for (int k = 0; k < nreg[s]; ++k) {
                    const uint32_t v = h[((s * 2 + half) * 128 + t) * 32 + k];
                    printf(" %u:%u", v / 1000 - 32, v % 1000);
                }
                printf("\n");
            }
        }
    return 0;
}
