// Dev tool (GPU box): verifies the tcgen05 operand layouts / descriptors of csrc/tc_prims.cuh against a host matmul
// and measures the issue->commit->wait chain.  One configuration per process (an illegal shape kills the context).
//   tc_probe <fmt: tf32|bf16> <ts: 0|1> <N> <swap_lbo_sbo: 0|1> <iters> <reps>
//   D[128 x N] = A[128 x 64] * B[N x 64]^T ; reps = how many times the K loop is re-issued per commit.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../neural-tape-modeling_b200/csrc/tc_prims.cuh"

using namespace ntm::tc;

struct Args {
    const float* A;
    const float* B;
    float* D;
    long long* cyc;
    int N, ts, swap, iters, reps, bmn;
};

template <int FMT>
__global__ void __launch_bounds__(128, 1) probe_kernel(Args p)
{
    constexpr int ELT = FMT == FMT_TF32 ? 4 : 2;
    constexpr int KSTEP = 32 / ELT;          // elements per MMA
    constexpr int NK = 64 / KSTEP;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int N = p.N;
    uint8_t* sA = smem;                       // 128 x 64 x ELT
    uint8_t* sB = smem + 128 * 64 * ELT;      // N x 64 x ELT
    const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16, sbo = 128;

    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 512);
        tmem_relinquish();
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t colA = 256;

    // ---- operands ---------------------------------------------------------------------------------
    for (int idx = tid; idx < N * 64; idx += 128) {
        const int n = idx / 64, k = idx % 64;
        const float v = p.B[idx];
        if (p.bmn) {
            // MN-major B, no swizzle: 16-byte vectors of T consecutive n; 8 consecutive k at 16-byte stride (one 128-byte
            // core matrix); n-vector groups at SBO = 128, k groups of 8 at LBO = (N / T) * 128
            constexpr int T = 16 / ELT;
            const uint32_t off = (uint32_t)(n % T) * ELT + (uint32_t)(k % 8) * 16 + (uint32_t)(n / T) * 128 + (uint32_t)(k / 8) * ((uint32_t)(N / T) * 128);
            if (FMT == FMT_TF32) *(uint32_t*)(sB + off) = to_tf32(v);
            else *(__nv_bfloat16*)(sB + off) = __float2bfloat16_rn(v);
        } else if (FMT == FMT_TF32) *(uint32_t*)(sB + kmajor_off<4>(n, k, lboB, sbo)) = to_tf32(v);
        else *(__nv_bfloat16*)(sB + kmajor_off<2>(n, k, lboB, sbo)) = __float2bfloat16_rn(v);
    }
    if (!p.ts) {
        for (int idx = tid; idx < 128 * 64; idx += 128) {
            const int m = idx / 64, k = idx % 64;
            const float v = p.A[idx];
            if (FMT == FMT_TF32) *(uint32_t*)(sA + kmajor_off<4>(m, k, lboA, sbo)) = to_tf32(v);
            else *(__nv_bfloat16*)(sA + kmajor_off<2>(m, k, lboA, sbo)) = __float2bfloat16_rn(v);
        }
    } else {
        const float* arow = p.A + tid * 64;   // thread = row = TMEM lane
        if (FMT == FMT_TF32) {
            for (int c = 0; c < 64; c += 16) {
                uint32_t r[16];
                for (int i = 0; i < 16; ++i) r[i] = to_tf32(arow[c + i]);
                tmem_st16(tmem + lane_base + colA + c, r);
            }
        } else {
            for (int c = 0; c < 32; c += 16) {
                uint32_t r[16];
                for (int i = 0; i < 16; ++i) r[i] = pack_bf16(arow[2 * (c + i)], arow[2 * (c + i) + 1]);
                tmem_st16(tmem + lane_base + colA + c, r);
            }
        }
        tmem_st_wait();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---- MMA issue ---------------------------------------------------------------------------------
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = instr_desc(FMT, 128, N) | (p.bmn ? (1u << 16) : 0u);
        const uint32_t aL = p.swap ? sbo : lboA, aS = p.swap ? lboA : sbo;
        constexpr int TB = 16 / ELT;
        const uint32_t lboB_mn = (uint32_t)(N / TB) * 128;          // k-group (8 k) stride
        const uint32_t bL = p.bmn ? (p.swap ? 128u : lboB_mn) : (p.swap ? sbo : lboB);
        const uint32_t bS = p.bmn ? (p.swap ? lboB_mn : 128u) : (p.swap ? lboB : sbo);
        const uint32_t bstep = p.bmn ? (KSTEP / 8) * lboB_mn : 2 * lboB;   // bytes between consecutive MMAs along K
        long long t_issue = 0;
        const long long t0 = clock64();
        for (int it = 0; it < p.iters; ++it) {
            for (int rep = 0; rep < p.reps; ++rep) {
                for (int ks = 0; ks < NK; ++ks) {
                    const uint64_t bd = smem_desc(smem_u32(sB) + ks * bstep, bL, bS);
                    if (p.ts) mma_ts<FMT>(tmem, tmem + colA + ks * 8, bd, idesc, ks > 0);
                    else mma_ss<FMT>(tmem, smem_desc(smem_u32(sA) + ks * 2 * lboA, aL, aS), bd, idesc, ks > 0);
                }
            }
            mma_commit(&bar);
            if (it == 0) t_issue = clock64() - t0;
            mbar_wait(&bar, it & 1);
        }
        const long long t1 = clock64();
        p.cyc[0] = t1 - t0;
        p.cyc[1] = t_issue;
    }
    __syncthreads();
    tc_fence_after();

    // ---- D -> global: thread = row ----------------------------------------------------------------------
    for (int c = 0; c < N; c += 8) {
        uint32_t r[8];
        tmem_ld8(tmem + lane_base + c, r);
        tmem_ld_wait();
        for (int i = 0; i < 8; ++i) p.D[tid * N + c + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}


// ---- chain latency with everything compile-time: descriptors precomputed, loop unrolled ---------------------
template <int FMT, int TS, int N, int REPS, int M = 128>
__global__ void __launch_bounds__(128, 1) chain_kernel(long long* cyc, int iters)
{
    constexpr int ELT = FMT == FMT_TF32 ? 4 : 2;
    constexpr int NK = 64 / (32 / ELT);
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* sA = smem;
    uint8_t* sB = smem + 128 * 64 * ELT;
    constexpr uint32_t lboA = 128 * 16, lboB = N * 16, sbo = 128;
    for (int i = tid; i < (128 * 64 * ELT + N * 64 * ELT) / 4; i += 128) ((uint32_t*)smem)[i] = 0;
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (__shfl_sync(0xffffffffu, warp, 0) == 1 && elect_one()) {
        constexpr uint32_t idesc = instr_desc(FMT, M, N);
        uint64_t ad[NK], bd[NK];
#pragma unroll
        for (int ks = 0; ks < NK; ++ks) {
            ad[ks] = smem_desc(smem_u32(sA) + ks * 2 * lboA, lboA, sbo);
            bd[ks] = smem_desc(smem_u32(sB) + ks * 2 * lboB, lboB, sbo);
        }
        long long t_issue = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const long long ti = clock64();
#pragma unroll
            for (int rep = 0; rep < REPS; ++rep) {
#pragma unroll
                for (int ks = 0; ks < NK; ++ks) {
                    if (TS) mma_ts<FMT>(tmem + rep * N, tmem + 448 + ks * 8, bd[ks], idesc, ks > 0);
                    else mma_ss<FMT>(tmem + rep * N, ad[ks], bd[ks], idesc, ks > 0);
                }
            }
            mma_commit(&bar);
            t_issue += clock64() - ti;
            mbar_wait(&bar, it & 1);
        }
        cyc[0] = clock64() - t0;
        cyc[1] = t_issue;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int FMT, int TS, int N, int REPS, int M = 128>
static int run_chain(const char* name, int iters)
{
    long long* dC;
    cudaMalloc(&dC, 16);
    const int smem = 128 * 64 * 4 + N * 64 * 4 + 1024;
    cudaFuncSetAttribute(chain_kernel<FMT, TS, N, REPS, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    chain_kernel<FMT, TS, N, REPS, M><<<1, 128, smem>>>(dC, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc[2] = {0, 0};
    cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost);
    constexpr int NK = FMT == FMT_TF32 ? 8 : 4;
    printf("chain %-26s mmas/iter=%2d : cycles/iter=%7.1f  issue/iter=%7.1f  (%s)\n", name, NK * REPS,
           (double)cyc[0] / iters, (double)cyc[1] / iters, cudaGetErrorString(e));
    return 0;
}


// ---- legacy warp-level mma.sync latency / throughput (m16n8k16 f16, m16n8k8 tf32) ------------------------------
__device__ __forceinline__ void hmma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void hmma1688tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int CHAINS, int TF32>
__global__ void hmma_kernel(long long* cyc, float* sink, int iters)
{
    float c[CHAINS][4];
    uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x3c003c00u};
    if (TF32) { a[0] = a[1] = a[2] = a[3] = 0x3f800000u; b[0] = b[1] = 0x3f800000u; }
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                if (TF32) hmma1688tf32(c[i], a, b);
                else hmma16816(c[i], a, b);
            }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    sink[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CHAINS, int TF32>
static void run_hmma(int warps, int blocks)
{
    long long* dC; float* dS;
    cudaMalloc(&dC, 16); cudaMalloc(&dS, 4 * 32 * warps * blocks);
    const int iters = 1000;
    hmma_kernel<CHAINS, TF32><<<blocks, 32 * warps>>>(dC, dS, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    printf("mma.sync %s chains=%d warps/CTA=%d CTAs=%d: %.1f cycles per 4-deep dependent chain set (%d mma) -> %.1f clk/mma/warp (%s)\n",
           TF32 ? "m16n8k8.tf32" : "m16n8k16.f16", CHAINS, warps, blocks, (double)cyc / iters, 4 * CHAINS,
           (double)cyc / iters / (4 * CHAINS), cudaGetErrorString(e));
}

static float round_tf32(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}
static float round_bf16(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
    memcpy(&x, &u, 4);
    return x;
}

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)

int main(int argc, char** argv)
{
    if (argc >= 2 && !strcmp(argv[1], "chain")) {
        const int it = 500;
        run_chain<FMT_TF32, 0, 16, 1>("tf32 SS N=16 x1", it);
        run_chain<FMT_TF32, 0, 16, 3>("tf32 SS N=16 x3", it);
        run_chain<FMT_TF32, 0, 8, 3>("tf32 SS N=8 x3", it);
        run_chain<FMT_TF32, 0, 32, 3>("tf32 SS N=32 x3", it);
        run_chain<FMT_TF32, 0, 64, 3>("tf32 SS N=64 x3", it);
        run_chain<FMT_TF32, 1, 192, 1>("tf32 TS N=192 x1", it);
        run_chain<FMT_TF32, 1, 192, 2>("tf32 TS N=192 x2", it);
        run_chain<FMT_TF32, 0, 192, 1>("tf32 SS N=192 x1", it);
        run_chain<FMT_BF16, 0, 16, 1>("bf16 SS N=16 x1", it);
        run_chain<FMT_BF16, 0, 16, 3>("bf16 SS N=16 x3", it);
        run_chain<FMT_BF16, 0, 16, 9>("bf16 SS N=16 x9", it);
        run_chain<FMT_BF16, 1, 192, 1>("bf16 TS N=192 x1", it);
        run_chain<FMT_BF16, 1, 192, 2>("bf16 TS N=192 x2", it);
        run_chain<FMT_BF16, 0, 192, 1>("bf16 SS N=192 x1", it);
        run_chain<FMT_BF16, 0, 8, 3, 64>("bf16 SS M=64 N=8 x3", it);
        run_chain<FMT_BF16, 0, 8, 9, 64>("bf16 SS M=64 N=8 x9", it);
        run_chain<FMT_BF16, 0, 16, 9, 64>("bf16 SS M=64 N=16 x9", it);
        run_chain<FMT_BF16, 0, 8, 2>("bf16 SS N=8 x2", it);
        run_chain<FMT_BF16, 0, 64, 3>("bf16 SS N=64 x3", it);
        run_chain<FMT_BF16, 0, 64, 9>("bf16 SS N=64 x9", it);
        run_chain<FMT_BF16, 0, 128, 3>("bf16 SS N=128 x3", it);
        return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "hmma")) {
        run_hmma<1, 0>(1, 1); run_hmma<3, 0>(1, 1); run_hmma<3, 0>(4, 1); run_hmma<6, 0>(4, 1); run_hmma<3, 0>(8, 1); run_hmma<3, 0>(4, 148);
        run_hmma<1, 1>(1, 1); run_hmma<3, 1>(1, 1); run_hmma<3, 1>(4, 1); run_hmma<6, 1>(4, 1); run_hmma<3, 1>(8, 1);
        return 0;
    }
    if (argc < 7) {
        printf("usage: tc_probe tf32|bf16 ts N swap iters reps\n");
        return 1;
    }
    const bool tf32 = !strcmp(argv[1], "tf32");
    Args a{};
    a.ts = atoi(argv[2]);
    a.N = atoi(argv[3]);
    a.swap = atoi(argv[4]);
    a.iters = atoi(argv[5]);
    a.reps = atoi(argv[6]);
    a.bmn = argc > 7 ? atoi(argv[7]) : 0;
    const int N = a.N;
    std::vector<float> A(128 * 64), B(N * 64), D(128 * N);
    srand(1234);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    long long* dC;
    CK(cudaMalloc(&dA, A.size() * 4));
    CK(cudaMalloc(&dB, B.size() * 4));
    CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMalloc(&dC, 16));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, D.size() * 4));
    a.A = dA; a.B = dB; a.D = dD; a.cyc = dC;
    const int smem = 128 * 64 * 4 + N * 64 * 4 + 1024;
    if (tf32) {
        CK(cudaFuncSetAttribute(probe_kernel<FMT_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        probe_kernel<FMT_TF32><<<1, 128, smem>>>(a);
    } else {
        CK(cudaFuncSetAttribute(probe_kernel<FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        probe_kernel<FMT_BF16><<<1, 128, smem>>>(a);
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long cyc[2];
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < 64; ++k) {
                const float x = tf32 ? round_tf32(A[m * 64 + k]) : round_bf16(A[m * 64 + k]);
                const float y = tf32 ? round_tf32(B[n * 64 + k]) : round_bf16(B[n * 64 + k]);
                s += (double)x * y;
            }
            maxerr = fmax(maxerr, fabs(s - D[m * N + n]));
            maxref = fmax(maxref, fabs(s));
        }
    printf("%s ts=%d N=%3d swap=%d iters=%d reps=%d bmn=%d : max|err|=%.3e (max|ref|=%.2f) %s   cycles/iter=%.1f first-issue=%lld\n",
           argv[1], a.ts, N, a.swap, a.iters, a.reps, a.bmn, maxerr, maxref, maxerr < 1e-4 ? "OK " : "BAD",
           (double)cyc[0] / a.iters, cyc[1]);
    if (maxerr >= 1e-4) {
        printf("  D[0][0..7]   =");
        for (int i = 0; i < 8; ++i) printf(" %8.4f", D[i]);
        printf("\n  D[1][0..7]   =");
        for (int i = 0; i < 8; ++i) printf(" %8.4f", D[N + i]);
        printf("\n");
    }
    return 0;
}
