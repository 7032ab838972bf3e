// Micro-benchmark (B200): issue rate of the packed-fp16 instructions the ray-cast is made of.
// Each variant runs 8 independent dependency chains per thread, 16 warps per SMSP worth of threads, and reports
// warp-instructions per cycle per SM sub-partition (1.0 = one instruction every cycle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp16 tools/ubench_fp16.cu && ./ubench_fp16
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t* out, int iters, uint32_t seed) {
    __half2 a[8], b, c;
    float f[8];
    uint32_t u[8];
    b = __floats2half2_rn(1.0001f + seed, 0.9999f);
    c = __floats2half2_rn(0.5f, 0.25f + seed);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = __floats2half2_rn(1.0f + i + threadIdx.x, 2.0f + i);
        f[i] = 1.0f + i + threadIdx.x;
        u[i] = i * 77u + threadIdx.x;
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = __hfma2(a[i], b, c);                                    // HFMA2
            if (MODE == 1) a[i] = __hadd2_rn(a[i], b);                                    // add.rn.f16x2
            if (MODE == 2) a[i] = __hmul2_rn(a[i], b);                                    // mul.rn.f16x2
            if (MODE == 3) { a[i] = __hmul2_rn(a[i], b); a[i] = __hadd2_rn(a[i], c); }    // mul then add, not contracted
            if (MODE == 4) f[i] = __fmaf_rn(f[i], 1.0001f, 0.5f);                         // FFMA
            if (MODE == 5) u[i] = __byte_perm(u[i], seed, 0x5410) + 1u;                   // PRMT + IADD
            if (MODE == 6) { a[i] = __hmul2_rn(a[i], b); u[i] = (u[i] ^ seed) + 3u; }     // fp16 + integer interleaved
            if (MODE == 7) { a[i] = __hfma2(a[i], b, c); f[i] = __fmaf_rn(f[i], 1.0001f, 0.5f); }   // HFMA2 + FFMA interleaved
            if (MODE == 8) { uint32_t m = __hge2_mask(a[i], b); a[i] = __hadd2_rn(a[i], c); u[i] ^= m; }   // HSET2 + HADD2 + LOP
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= *reinterpret_cast<uint32_t*>(&a[i]) ^ __float_as_uint(f[i]) ^ u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r ^ (uint32_t)(t1 - t0);
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (uint32_t)(t1 - t0);
}

template <int MODE>
void run(const char* name, int per_iter) {
    uint32_t* d;
    cudaMalloc(&d, 148 * 4 * 512 * 4);
    const int iters = 4096;
    k<MODE><<<148 * 4, 512>>>(d, 16, 0);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 512>>>(d, iters, 0);     // 4 CTAs x 16 warps per SM = 16 warps per sub-partition
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    uint32_t cyc;
    cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
    // per sub-partition: 16 warps x iters x 8 x per_iter warp-instructions in `cyc` cycles (CTA 0's view)
    const double inst = 16.0 * iters * 8 * per_iter;
    printf("%-44s %7.3f ms  %9u cycles  %.3f warp-inst/cycle/SMSP  (event-time clock %.0f MHz)\n", name, ms, cyc, inst / cyc,
           cyc / (ms * 1e3));
    cudaFree(d);
}

int main() {
    run<0>("HFMA2 (fused)", 1);
    run<1>("HADD2 .rn", 1);
    run<2>("HMUL2 .rn", 1);
    run<3>("HMUL2 + HADD2 (un-contracted pair)", 2);
    run<4>("FFMA", 1);
    run<5>("PRMT + IADD", 2);
    run<6>("HMUL2 + LOP3/IADD", 3);
    run<7>("HFMA2 + FFMA", 2);
    run<8>("HSET2 + HADD2 + LOP3", 3);
    return 0;
}
