// Policy-inference epilogue: the direct consumer of obs_buf.
//
// Replaces StochasticActorHeightmap.compute / DeterministicHeightmap.compute (learning/model.py:152-195, 197-241):
//   x0 = encoder0(obs[:, p : p+S])            Linear(S,80)  + act, Linear(80,60) + act      (model.py:118-144,186)
//   x1 = encoder1(obs[:, p+S : p+S+D])        Linear(D,80)  + act, Linear(80,60) + act      (model.py:187)
//   x  = cat(obs[:, 0:p], x0, x1)             (model.py:188-189)
//   x  = Linear(124,256)+act, Linear(256,160)+act, Linear(160,128)+act, Linear(128,A) [+ Tanh for the actor] (model.py:169-177,190-191)
// with the layer widths the reference hard-wires (train.py:95, cfg/trainSKRL/RoverPPOSKRL.yaml:4-9).
//
// Arithmetic: the reference runs torch fp32 nn.Linear (TF32 is off by default for matmuls), so this kernel is an fp32
// FMA chain -- no tensor cores: a bf16/TF32 contraction would change the answer by 1e-3, and the whole net is
// 0.27 MFMA per env (2.2 GFLOP per 4096-env step), i.e. tens of microseconds of fp32 pipe.  Sums are accumulated in
// k order per output with fmaf (cuBLAS's order is unspecified; the parity gate is 2e-5 absolute against an fp64 oracle).
//
// One CTA of 256 threads per tile of TM = 32 envs; every activation of the tile stays in shared memory from the
// observation row to the action (k-major [k][env], row stride 36 floats), so obs is read exactly once from HBM/L2 and
// nothing but the A outputs is written.  Weights are re-packed once per handle into k-major, 32-column-padded panels
// (Wt[k][OUTP]) that every CTA streams from L2 through a 16 KB staging buffer.  Thread (mg = warp, ng = lane) owns
// 4 envs x CN columns {ng + 32 j}: per k one broadcast 128-bit activation load + CN conflict-free weight loads feed
// 4 CN FMAs.  99 KB shared memory -> 2 CTAs per SM.
#include <new>

#include "common.cuh"

#define PL_TM 32
#define PL_LDA 36
#define PL_THREADS 256
#define PL_WS_FLOATS 4096          // weight staging buffer: KC x OUTP floats
#define PL_E1 80
#define PL_E2 60
#define PL_M1 256
#define PL_M2 160
#define PL_M3 128
#define PL_MAX_PROPRIO 8
#define PL_MAX_HEAD 4

static inline int pad32(int n) { return (n + 31) / 32 * 32; }

struct PackedLinear {
    float* wt;     // [in][outp]  k-major, zero padded columns
    float* bias;   // [outp]
    int in, out, outp;
};

// What the kernel reads: a copy lives at the start of the handle's device allocation, so one launch can serve two
// networks (blockIdx.y) without indexing kernel parameters dynamically.
struct PolicyDev {
    PackedLinear es1, es2, ed1, ed2, m1, m2, m3;
    float* head_w;   // [A][PL_M3]  torch layout
    float* head_b;   // [A]
    int n_proprio, n_sparse, n_dense, n_head;
    int act, head_tanh;
};

struct rvb_policy : PolicyDev {
    int device;
    float* storage;        // the one allocation everything points into: [PolicyDev | panels | head]
    int64_t storage_floats;
    const PolicyDev* dev;  // = storage
};
#define PL_DESC_FLOATS 128   // room reserved for the descriptor (keeps the panels 512-byte aligned)
static_assert(sizeof(PolicyDev) <= PL_DESC_FLOATS * sizeof(float), "descriptor does not fit its slot");

__device__ __forceinline__ float pl_act(float v, int kind) {
    switch (kind) {
        case RVB_ACT_LEAKYRELU: return v > 0.f ? v : __fmul_rn(v, 0.01f);   // nn.LeakyReLU() default slope (model.py:107)
        case RVB_ACT_RELU: return fmaxf(v, 0.f);
        case RVB_ACT_ELU: return v > 0.f ? v : expm1f(v);
        case RVB_ACT_TANH: return tanhf(v);
        case RVB_ACT_SIGMOID: return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-v)));
        case RVB_ACT_RELU6: return fminf(fmaxf(v, 0.f), 6.f);
        default: return v;
    }
}

// Two fp32 FMAs in one issue slot (Blackwell FFMA2: d.xy = a.xy * b + c.xy; ptxas folds the {b, b} pair into the
// instruction's broadcast operand).  Each half is an IEEE fma, so the result equals two fmaf() calls bit for bit.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a0), "f"(a1), "f"(b));
}

// torch [out][in] -> k-major padded panel
__global__ void pl_pack_kernel(const float* __restrict__ w, const float* __restrict__ b, int in, int out, int outp,
                               float* __restrict__ wt, float* __restrict__ bias) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < (int64_t)in * outp) {
        const int k = (int)(i / outp), n = (int)(i % outp);
        wt[i] = n < out ? w[(int64_t)n * in + k] : 0.f;
    }
    if (i < outp) bias[i] = i < out ? b[i] : 0.f;
}

// One dense layer of the tile.  Input activations: shared memory `act_in` [K][PL_LDA] (GLOBAL_IN = false) or columns
// [col0, col0 + K) of the tile's observation rows in global memory, staged transposed through `xs` (GLOBAL_IN = true).
// Output: act(in . Wt + bias) into shared memory `act_out` [n][PL_LDA], rows n < OUT only.
template <int CN, bool GLOBAL_IN, bool F2>
__device__ __forceinline__ void pl_dense(const float* __restrict__ act_in, const float* __restrict__ obs, int64_t obs_ld, int64_t m0,
                                         int64_t N, int col0, int K, const PackedLinear& L, float* __restrict__ xs,
                                         float* __restrict__ ws, float* __restrict__ act_out, int act_kind) {
    constexpr int OUTP = 32 * CN;
    constexpr int KC = (PL_WS_FLOATS / OUTP) < 32 ? (PL_WS_FLOATS / OUTP) : 32;
    const int tid = threadIdx.x, ng = tid & 31, mg = tid >> 5;
    float acc[4][CN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CN; ++j) acc[i][j] = 0.f;

    // Software pipeline: chunk i+1 travels global -> registers while chunk i is multiplied out of shared memory.
    constexpr int WV = (KC * (OUTP / 4) + PL_THREADS - 1) / PL_THREADS;
    float4 wreg[WV];
    float xreg[PL_TM / 8];
    auto fetch = [&](int k0) {
        const int kc = min(KC, K - k0);
        const float4* src = reinterpret_cast<const float4*>(L.wt + (int64_t)k0 * OUTP);   // contiguous rows, 16-byte aligned
#pragma unroll
        for (int v = 0; v < WV; ++v) {
            const int i = tid + v * PL_THREADS;
            wreg[v] = i < kc * (OUTP / 4) ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (GLOBAL_IN) {                                   // a warp reads 32 consecutive floats of one env's row
#pragma unroll
            for (int r = 0; r < PL_TM / 8; ++r) {
                const int m = mg + 8 * r;
                xreg[r] = (ng < kc && m0 + m < N) ? __ldg(obs + (m0 + m) * obs_ld + col0 + k0 + ng) : 0.f;
            }
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += KC) {
        const int kc = min(KC, K - k0);
        __syncthreads();                                   // previous chunk fully consumed (and act_in fully written)
#pragma unroll
        for (int v = 0; v < WV; ++v) {
            const int i = tid + v * PL_THREADS;
            if (i < kc * (OUTP / 4)) reinterpret_cast<float4*>(ws)[i] = wreg[v];
        }
        if (GLOBAL_IN) {                                   // transposed: xs[k][env]
#pragma unroll
            for (int r = 0; r < PL_TM / 8; ++r) xs[ng * PL_LDA + mg + 8 * r] = xreg[r];
        }
        __syncthreads();
        if (k0 + KC < K) fetch(k0 + KC);
        const float* a_base = GLOBAL_IN ? xs : act_in + (int64_t)k0 * PL_LDA;
        if (F2) {
            auto step = [&](int kk) {
                const float4 a = *reinterpret_cast<const float4*>(a_base + kk * PL_LDA + mg * 4);
                float w[CN];
#pragma unroll
                for (int j = 0; j < CN; ++j) w[j] = ws[kk * OUTP + ng + 32 * j];
#pragma unroll
                for (int j = 0; j < CN; ++j) {
                    ffma2(acc[0][j], acc[1][j], a.x, a.y, w[j]);
                    ffma2(acc[2][j], acc[3][j], a.z, a.w, w[j]);
                }
            };
            if (kc == KC) {                                // full chunks: no loop overhead
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) step(kk);
            } else {
#pragma unroll 4
                for (int kk = 0; kk < kc; ++kk) step(kk);
            }
        } else {
#pragma unroll 4
            for (int kk = 0; kk < kc; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(a_base + kk * PL_LDA + mg * 4);
                float w[CN];
#pragma unroll
                for (int j = 0; j < CN; ++j) w[j] = ws[kk * OUTP + ng + 32 * j];
#pragma unroll
                for (int j = 0; j < CN; ++j) {
                    acc[0][j] = fmaf(a.x, w[j], acc[0][j]);
                    acc[1][j] = fmaf(a.y, w[j], acc[1][j]);
                    acc[2][j] = fmaf(a.z, w[j], acc[2][j]);
                    acc[3][j] = fmaf(a.w, w[j], acc[3][j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CN; ++j) {
        const int n = ng + 32 * j;
        if (n < L.out) {
            const float b = __ldg(L.bias + n);
            float4 o;
            o.x = pl_act(__fadd_rn(acc[0][j], b), act_kind);
            o.y = pl_act(__fadd_rn(acc[1][j], b), act_kind);
            o.z = pl_act(__fadd_rn(acc[2][j], b), act_kind);
            o.w = pl_act(__fadd_rn(acc[3][j], b), act_kind);
            *reinterpret_cast<float4*>(act_out + n * PL_LDA + mg * 4) = o;
        }
    }
}

// shared-memory plan (floats)
#define PL_SM_A 0                                   // concat [<=128][LDA], later the third hidden layer [128][LDA]
#define PL_SM_B (PL_SM_A + 128 * PL_LDA)            // first hidden layer [256][LDA]
#define PL_SM_C (PL_SM_B + PL_M1 * PL_LDA)          // encoder hidden [80][LDA], later the second hidden layer [160][LDA]
#define PL_SM_XS (PL_SM_C + PL_M2 * PL_LDA)         // obs chunk [32][LDA]
#define PL_SM_WS (PL_SM_XS + 32 * PL_LDA)           // weight chunk
#define PL_SM_FLOATS (PL_SM_WS + PL_WS_FLOATS)
#define PL_SMEM_BYTES (PL_SM_FLOATS * 4)

template <bool F2>
__global__ void __launch_bounds__(PL_THREADS, 2)
policy_forward_kernel(const PolicyDev* __restrict__ p0, const PolicyDev* __restrict__ p1, const float* __restrict__ obs,
                      int64_t obs_ld, int64_t N, float* __restrict__ out0, int64_t out0_ld, float* __restrict__ out1,
                      int64_t out1_ld) {
    // blockIdx.y selects the network (actor / critic of one PPO step read the same obs tile, hot in L2)
    const PolicyDev& P = *(blockIdx.y ? p1 : p0);
    float* __restrict__ out = blockIdx.y ? out1 : out0;
    const int64_t out_ld = blockIdx.y ? out1_ld : out0_ld;
    extern __shared__ __align__(16) float sm[];
    float* A = sm + PL_SM_A;
    float* B = sm + PL_SM_B;
    float* Cb = sm + PL_SM_C;
    float* xs = sm + PL_SM_XS;
    float* ws = sm + PL_SM_WS;
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * PL_TM;
    const int p = P.n_proprio;

    // proprioceptive columns -> concat rows 0..p-1 (model.py:188)
    if (tid < PL_TM * p) {
        const int m = tid & 31, k = tid >> 5;
        A[k * PL_LDA + m] = (m0 + m < N) ? __ldg(obs + (m0 + m) * obs_ld + k) : 0.f;
    }
    // sparse encoder (model.py:186) -> concat rows p .. p+59
    pl_dense<3, true, F2>(nullptr, obs, obs_ld, m0, N, p, P.n_sparse, P.es1, xs, ws, Cb, P.act);
    pl_dense<2, false, F2>(Cb, nullptr, 0, 0, 0, 0, PL_E1, P.es2, xs, ws, A + p * PL_LDA, P.act);
    // dense encoder (model.py:187) -> concat rows p+60 .. p+119
    pl_dense<3, true, F2>(nullptr, obs, obs_ld, m0, N, p + P.n_sparse, P.n_dense, P.ed1, xs, ws, Cb, P.act);
    pl_dense<2, false, F2>(Cb, nullptr, 0, 0, 0, 0, PL_E1, P.ed2, xs, ws, A + (p + PL_E2) * PL_LDA, P.act);
    // MLP (model.py:190-191)
    pl_dense<8, false, F2>(A, nullptr, 0, 0, 0, 0, p + 2 * PL_E2, P.m1, xs, ws, B, P.act);
    pl_dense<5, false, F2>(B, nullptr, 0, 0, 0, 0, PL_M1, P.m2, xs, ws, Cb, P.act);
    pl_dense<4, false, F2>(Cb, nullptr, 0, 0, 0, 0, PL_M2, P.m3, xs, ws, A, P.act);
    __syncthreads();
    // head: Linear(128, A) [+ Tanh] -- warp o computes output o for the tile's 32 envs
    const int m = tid & 31, o = tid >> 5;
    if (o < P.n_head && m0 + m < N) {
        const float* w = P.head_w + o * PL_M3;
        float acc = 0.f;
#pragma unroll 8
        for (int k = 0; k < PL_M3; ++k) acc = fmaf(A[k * PL_LDA + m], __ldg(w + k), acc);
        acc = __fadd_rn(acc, __ldg(P.head_b + o));
        out[(m0 + m) * out_ld + o] = P.head_tanh ? tanhf(acc) : acc;
    }
}

static int check_linear(const rvb_linear* L, int in, int out, const char* what) {
    if (!L || !L->weight || !L->bias) return rvb_set_error(RVB_ERR_INVALID, "rvb_policy_create: null layer", what);
    if (L->in_features != in || L->out_features != out)
        return rvb_set_error(RVB_ERR_UNSUPPORTED, "rvb_policy_create: layer shape differs from the reference network "
                             "(encoders [80,60], mlp [256,160,128]; train.py:95)", what);
    return RVB_OK;
}

extern "C" int rvb_policy_create(rvb_policy** out, int32_t n_proprio, int32_t n_sparse, int32_t n_dense,
                                 const rvb_linear* enc_sparse, const rvb_linear* enc_dense, const rvb_linear* mlp,
                                 const rvb_linear* head, int32_t activation, int32_t head_tanh, int device, void* stream) {
    RVB_REQUIRE(out, "rvb_policy_create: null pointer");
    *out = nullptr;
    RVB_REQUIRE(enc_sparse && enc_dense && mlp && head, "rvb_policy_create: null pointer");
    RVB_REQUIRE(n_proprio >= 0 && n_proprio <= PL_MAX_PROPRIO && n_sparse > 0 && n_dense > 0,
                "rvb_policy_create: bad observation split");
    RVB_REQUIRE(activation >= RVB_ACT_LEAKYRELU && activation <= RVB_ACT_RELU6, "rvb_policy_create: unknown activation");
    RVB_REQUIRE(head->out_features >= 1 && head->out_features <= PL_MAX_HEAD, "rvb_policy_create: head width must be 1..4");
    int rc;
    if ((rc = check_linear(&enc_sparse[0], n_sparse, PL_E1, "sparse encoder layer 0"))) return rc;
    if ((rc = check_linear(&enc_sparse[1], PL_E1, PL_E2, "sparse encoder layer 1"))) return rc;
    if ((rc = check_linear(&enc_dense[0], n_dense, PL_E1, "dense encoder layer 0"))) return rc;
    if ((rc = check_linear(&enc_dense[1], PL_E1, PL_E2, "dense encoder layer 1"))) return rc;
    if ((rc = check_linear(&mlp[0], n_proprio + 2 * PL_E2, PL_M1, "mlp layer 0"))) return rc;
    if ((rc = check_linear(&mlp[1], PL_M1, PL_M2, "mlp layer 1"))) return rc;
    if ((rc = check_linear(&mlp[2], PL_M2, PL_M3, "mlp layer 2"))) return rc;
    if ((rc = check_linear(head, PL_M3, head->out_features, "head"))) return rc;

    int cur_dev = -1;                  // like every other entry point: the caller's current device, never changed here
    RVB_CUDA(cudaGetDevice(&cur_dev));
    RVB_REQUIRE(device == cur_dev, "rvb_policy_create: `device` is not the calling thread's current CUDA device");
    rvb_policy* P = new (std::nothrow) rvb_policy();
    if (!P) return rvb_set_error(RVB_ERR_NOMEM, "rvb_policy_create", "host allocation failed");
    P->n_proprio = n_proprio; P->n_sparse = n_sparse; P->n_dense = n_dense; P->n_head = head->out_features;
    P->act = activation; P->head_tanh = head_tanh ? 1 : 0; P->device = device;
    struct Slot { PackedLinear* dst; const rvb_linear* src; };
    Slot slots[7] = {{&P->es1, &enc_sparse[0]}, {&P->es2, &enc_sparse[1]}, {&P->ed1, &enc_dense[0]}, {&P->ed2, &enc_dense[1]},
                     {&P->m1, &mlp[0]}, {&P->m2, &mlp[1]}, {&P->m3, &mlp[2]}};
    int64_t total = 0;
    for (auto& s : slots) {
        s.dst->in = s.src->in_features; s.dst->out = s.src->out_features; s.dst->outp = pad32(s.src->out_features);
        total += (int64_t)s.dst->in * s.dst->outp + s.dst->outp;
    }
    total += (int64_t)P->n_head * PL_M3 + 32 + PL_DESC_FLOATS;
    cudaError_t e = cudaMalloc((void**)&P->storage, sizeof(float) * total);
    if (e != cudaSuccess) { delete P; return rvb_set_error(RVB_ERR_NOMEM, "rvb_policy_create: cudaMalloc", cudaGetErrorString(e)); }
    P->storage_floats = total;
    cudaStream_t st = as_stream(stream);
    P->dev = reinterpret_cast<const PolicyDev*>(P->storage);
    float* cur = P->storage + PL_DESC_FLOATS;
    for (auto& s : slots) {
        PackedLinear& L = *s.dst;
        L.wt = cur; cur += (int64_t)L.in * L.outp;       // every panel size is a multiple of 32 floats: 16-byte alignment holds
        L.bias = cur; cur += L.outp;
        const int64_t n = (int64_t)L.in * L.outp;
        pl_pack_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(s.src->weight, s.src->bias, L.in, L.out, L.outp, L.wt, L.bias);
    }
    P->head_w = cur; cur += (int64_t)P->n_head * PL_M3;
    P->head_b = cur;
    e = cudaMemcpyAsync(P->head_w, head->weight, sizeof(float) * P->n_head * PL_M3, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(P->head_b, head->bias, sizeof(float) * P->n_head, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess)       // the descriptor the kernel reads (P outlives the copy: the stream is synchronised below)
        e = cudaMemcpyAsync(P->storage, static_cast<const PolicyDev*>(P), sizeof(PolicyDev), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaFuncSetAttribute(policy_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PL_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(policy_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PL_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // the caller may free its weight tensors on return
    if (e != cudaSuccess) {
        cudaFree(P->storage);
        delete P;
        return rvb_set_error(RVB_ERR_CUDA, "rvb_policy_create", cudaGetErrorString(e));
    }
    *out = P;
    return RVB_OK;
}

extern "C" int rvb_policy_destroy(rvb_policy* P) {
    if (!P) return RVB_OK;
    cudaFree(P->storage);
    delete P;
    return RVB_OK;
}

extern "C" int64_t rvb_policy_bytes(const rvb_policy* P) { return P ? P->storage_floats * (int64_t)sizeof(float) : 0; }

// 1 = packed FFMA2 inner loop (default), 0 = scalar FFMA.  Bit-identical results; the switch exists for A/B measurement.
static int g_policy_variant = 1;
extern "C" int rvb_policy_variant(int v) {
    const int prev = g_policy_variant;
    if (v == 0 || v == 1) g_policy_variant = v;
    return prev;
}

static int check_forward(const rvb_policy* P, const float* obs, int64_t obs_ld, const float* out, int64_t out_ld) {
    RVB_REQUIRE(obs && out, "rvb_policy_forward: null pointer");
    RVB_REQUIRE(obs_ld >= (int64_t)P->n_proprio + P->n_sparse + P->n_dense, "rvb_policy_forward: obs rows are shorter than the network's input");
    RVB_REQUIRE(out_ld >= P->n_head, "rvb_policy_forward: out rows are shorter than the head");
    int cur_dev = -1;
    RVB_CUDA(cudaGetDevice(&cur_dev));
    RVB_REQUIRE(P->device == cur_dev, "rvb_policy_forward: the network lives on another device than the calling thread's current one");
    return RVB_OK;
}

extern "C" int rvb_policy_forward(const rvb_policy* P, const float* obs, int64_t obs_ld, int64_t N, float* out, int64_t out_ld,
                                  void* stream) {
    RVB_REQUIRE(P, "rvb_policy_forward: null handle");
    if (N <= 0) return RVB_OK;
    int rc;
    if ((rc = check_forward(P, obs, obs_ld, out, out_ld))) return rc;
    RVB_REQUIRE(N <= (int64_t)PL_TM * 0x7fffffff, "rvb_policy_forward: too many envs");
    auto kern = g_policy_variant ? policy_forward_kernel<true> : policy_forward_kernel<false>;
    kern<<<dim3((unsigned)ceil_div(N, PL_TM), 1), PL_THREADS, PL_SMEM_BYTES, as_stream(stream)>>>(
        P->dev, nullptr, obs, obs_ld, N, out, out_ld, nullptr, 0);
    RVB_LAUNCH_CHECK();
    return RVB_OK;
}

extern "C" int rvb_policy_forward_pair(const rvb_policy* A, const rvb_policy* B, const float* obs, int64_t obs_ld, int64_t N,
                                       float* out_a, int64_t out_a_ld, float* out_b, int64_t out_b_ld, void* stream) {
    RVB_REQUIRE(A && B, "rvb_policy_forward_pair: null handle");
    if (N <= 0) return RVB_OK;
    int rc;
    if ((rc = check_forward(A, obs, obs_ld, out_a, out_a_ld))) return rc;
    if ((rc = check_forward(B, obs, obs_ld, out_b, out_b_ld))) return rc;
    RVB_REQUIRE(A->device == B->device, "rvb_policy_forward_pair: the two networks live on different devices");
    RVB_REQUIRE(N <= (int64_t)PL_TM * 0x7fffffff, "rvb_policy_forward: too many envs");
    auto kern = g_policy_variant ? policy_forward_kernel<true> : policy_forward_kernel<false>;
    kern<<<dim3((unsigned)ceil_div(N, PL_TM), 2), PL_THREADS, PL_SMEM_BYTES, as_stream(stream)>>>(
        A->dev, B->dev, obs, obs_ld, N, out_a, out_a_ld, out_b, out_b_ld);
    RVB_LAUNCH_CHECK();
    return RVB_OK;
}
