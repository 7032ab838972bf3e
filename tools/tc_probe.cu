// Stand-alone probe of the tcgen05 building blocks the policy kernel uses (sm_100a): shared-memory operands in the canonical
// K-major SWIZZLE_128B layout written by plain stores, kind::tf32 MMAs with the accumulator in TMEM, a weight matrix split into
// tf32 hi + lo parts (two MMAs per k-step), tcgen05.ld of the result.  D[128 x N] = A[128 x K] . W[N x K]^T, fp32 result against
// an fp64 reference.  A holds fp16-representable values (like the heightmap observation columns), which tf32 keeps exactly.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_probe tools/tc_probe.cu && timeout 60 tools/tc_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int M = 128, N = 80, KC = 32;      // KC floats = one 128-byte swizzle row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor: start >> 4 | LBO << 16 | SBO << 32 | version 1 << 46 | layout 2 << 61)
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// rows x 32 floats, row r at (r / 8) * 1024 + (r % 8) * 128 bytes, 16-byte chunk c at position c ^ (r % 8)
__device__ __forceinline__ void store_row_swizzled(float* tile, int r, const float4* v8) {
    float* row = tile + (r >> 3) * 256 + (r & 7) * 32;
#pragma unroll
    for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(row + ((c ^ (r & 7)) << 2)) = v8[c];
}

template <int NCHUNK>
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ Wh, const float* __restrict__ Wl,
                                                    float* __restrict__ D, int K, int* status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* sA = reinterpret_cast<float*>(smem_raw);                    // [NCHUNK][128 * 32]
    float* sBh = sA + NCHUNK * M * KC;                                  // [NCHUNK][80 * 32]
    float* sBl = sBh + NCHUNK * N * KC;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int kc = 0; kc < NCHUNK; ++kc) {
        float4 v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(A + (size_t)tid * K + kc * KC + c * 4);
        store_row_swizzled(sA + kc * M * KC, tid, v);
        if (tid < N) {
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(Wh + (size_t)tid * K + kc * KC + c * 4);
            store_row_swizzled(sBh + kc * N * KC, tid, v);
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(Wl + (size_t)tid * K + kc * KC + c * 4);
            store_row_swizzled(sBl + kc * N * KC, tid, v);
        }
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy stores -> visible to the tensor core's reads
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        uint32_t acc = 0;
        for (int kc = 0; kc < NCHUNK; ++kc) {
            const uint64_t da = make_desc(smem_u32(sA + kc * M * KC)), dh = make_desc(smem_u32(sBh + kc * N * KC)),
                           dl = make_desc(smem_u32(sBl + kc * N * KC));
#pragma unroll
            for (int k = 0; k < 4; ++k) {                                 // 8 tf32 = 32 bytes = 2 descriptor units per MMA
                mma_tf32(tmem_base, da + 2 * k, dh + 2 * k, idesc, acc);
                acc = 1;
                mma_tf32(tmem_base, da + 2 * k, dl + 2 * k, idesc, 1);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // bounded wait for the MMAs (a wrong descriptor must not hang the box)
    {
        uint32_t done = 0;
        for (int it = 0; it < (1 << 22) && !done; ++it)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(smem_u32(&bar)), "r"(0)
                         : "memory");
        if (!done) {
            if (tid == 0) *status = 1;
            __syncthreads();
            if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
            return;
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
}

static float tf32_hi(float x) {          // round to nearest tf32 (10 explicit mantissa bits)
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

int main() {
    constexpr int NCHUNK = 2;
    const int K = NCHUNK * KC;
    float *hA = (float*)malloc(sizeof(float) * M * K), *hW = (float*)malloc(sizeof(float) * N * K), *hWh = (float*)malloc(sizeof(float) * N * K),
          *hWl = (float*)malloc(sizeof(float) * N * K), *hD = (float*)malloc(sizeof(float) * M * N);
    srand(7);
    for (int i = 0; i < M * K; ++i) hA[i] = __half2float(__float2half((float)rand() / RAND_MAX * 5.5f));       // fp16 values, like dist / 2
    for (int i = 0; i < N * K; ++i) {
        hW[i] = ((float)rand() / RAND_MAX - 0.5f) * 0.2f;
        hWh[i] = tf32_hi(hW[i]);
        hWl[i] = tf32_hi(hW[i] - hWh[i]);
    }
    float *dA, *dWh, *dWl, *dD;
    int* dS;
    cudaMalloc(&dA, sizeof(float) * M * K); cudaMalloc(&dWh, sizeof(float) * N * K); cudaMalloc(&dWl, sizeof(float) * N * K);
    cudaMalloc(&dD, sizeof(float) * M * N); cudaMalloc(&dS, sizeof(int));
    cudaMemcpy(dA, hA, sizeof(float) * M * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dWh, hWh, sizeof(float) * N * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dWl, hWl, sizeof(float) * N * K, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, sizeof(float) * M * N); cudaMemset(dS, 0, sizeof(int));
    const size_t smem = sizeof(float) * NCHUNK * (M + 2 * N) * KC + 1024;
    cudaFuncSetAttribute(probe_kernel<NCHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<NCHUNK><<<1, 128, smem>>>(dA, dWh, dWl, dD, K, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0;
    cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(hD, dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
    printf("launch: %s, status %d (1 = the MMAs never completed)\n", cudaGetErrorString(e), st);
    double maxerr = 0, maxerr_hi = 0, maxref = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0, ref_hi = 0;
            for (int k = 0; k < K; ++k) {
                ref += (double)hA[m * K + k] * (double)hW[n * K + k];
                ref_hi += (double)hA[m * K + k] * (double)hWh[n * K + k];
            }
            maxerr = fmax(maxerr, fabs(hD[m * N + n] - ref));
            maxerr_hi = fmax(maxerr_hi, fabs(hD[m * N + n] - ref_hi));
            maxref = fmax(maxref, fabs(ref));
        }
    printf("D[0][0..3] = %g %g %g %g\n", hD[0], hD[1], hD[2], hD[3]);
    printf("max |D - A.W| = %.3e (max |ref| %.3f); against the hi part alone: %.3e (the lo MMAs matter iff the first is much smaller)\n", maxerr,
           maxref, maxerr_hi);
    printf(maxerr < 2e-5 && st == 0 ? "PROBE OK\n" : "PROBE FAILED\n");
    return 0;
}
