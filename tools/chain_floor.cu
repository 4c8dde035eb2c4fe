// Dev tool (GPU box): on-chip latency floor of ONE dependent GRU step at batch 1 (cfg 5).
//
// Part 1 measures the primitive dependent latencies on this GPU (clk): FFMA, SHFL, MUFU.EX2, MUFU.RCP, LDS (broadcast
// float4), STS -> BAR.SYNC -> LDS round trip.  Part 2 runs the minimal dependent chain of a GRU-HS[64] step with nothing
// else in the loop -- state broadcast from shared memory, 3 x 64/KS dependent FMAs per thread, log2(KS) shuffle levels,
// the four dependent MUFU stages of the gates (ex2 -> rcp -> ex2 -> rcp), state blend, STS, one barrier -- for every
// k-split KS, on one SM.  The smallest figure is the floor the batch-1 kernels are compared against.
//   build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/chain_floor.bin tools/chain_floor.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int N_IT = 4096;

// ---- part 1: primitive dependent latencies (one warp) -----------------------------------------------------------
__global__ void prim_kernel(float* out, long long* cyc, float seed)
{
    __shared__ float sm[256];
    const int lane = threadIdx.x;
    sm[lane] = seed; sm[lane + 32] = seed; sm[lane + 64] = seed; sm[lane + 96] = seed;
    __syncthreads();
    float v = seed + lane * 1e-3f;
    long long t0, t1;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, 0f3A83126F;" : "+f"(v));
    t1 = clock64(); if (lane == 0) cyc[0] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) asm volatile("{.reg .f32 t; shfl.sync.bfly.b32 t, %0, 1, 31, 0xffffffff; add.f32 %0, t, 0f3A83126F;}" : "+f"(v));
    t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) asm volatile("{.reg .f32 t; ex2.approx.ftz.f32 t, %0; add.f32 %0, t, 0fBF800000;}" : "+f"(v));
    t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0;
    v = 1.0f + lane * 1e-3f + (v == 12345.0f ? 1.0f : 0.0f);     // keeps the loops above alive
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) asm volatile("{.reg .f32 t; rcp.approx.ftz.f32 t, %0; add.f32 %0, t, 0f3F000000;}" : "+f"(v));
    t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0;
    int idx = lane & 3;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) { const float4 q = *reinterpret_cast<const float4*>(&sm[4 * idx]); idx = ((int)q.x + idx + 1) & 15; }
    t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0;
    t0 = clock64();
    for (int i = 0; i < N_IT; ++i) { sm[lane] = v; __syncthreads(); v = sm[(lane + 1) & 31] + 1e-3f; }
    t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0;
    out[lane] = v + idx;
}

// ---- part 2: the minimal GRU step chain, 64 * KS threads ---------------------------------------------------------
template <int KS, int ACC>
__global__ void __launch_bounds__(64 * KS) chain_kernel(const float* __restrict__ w, float* out, long long* cyc)
{
    constexpr int KL = 64 / KS;                       // k elements per thread
    __shared__ __align__(16) float hs[2][64];
    const int tid = threadIdx.x, j = tid / KS, q = tid % KS;
    float wr[KL], wz[KL], wn[KL];
#pragma unroll
    for (int k = 0; k < KL; ++k) {
        wr[k] = w[(0 * 64 + j) * 64 + q * KL + k];
        wz[k] = w[(1 * 64 + j) * 64 + q * KL + k];
        wn[k] = w[(2 * 64 + j) * 64 + q * KL + k];
    }
    float h = 0.01f * j;
    if (q == 0) hs[0][j] = h;
    __syncthreads();
    int cur = 0;
    const long long t0 = clock64();
    for (int it = 0; it < N_IT; ++it) {
        float pr[ACC] = {}, pz[ACC] = {}, pn[ACC] = {};
        const float4* hp = reinterpret_cast<const float4*>(&hs[cur][q * KL]);
#pragma unroll
        for (int k4 = 0; k4 < (KL + 3) / 4; ++k4) {
            const float4 hv = hp[k4];
            const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int i = 0; i < 4 && 4 * k4 + i < KL; ++i) {
                pr[(4 * k4 + i) % ACC] = fmaf(wr[4 * k4 + i], hh[i], pr[(4 * k4 + i) % ACC]);
                pz[(4 * k4 + i) % ACC] = fmaf(wz[4 * k4 + i], hh[i], pz[(4 * k4 + i) % ACC]);
                pn[(4 * k4 + i) % ACC] = fmaf(wn[4 * k4 + i], hh[i], pn[(4 * k4 + i) % ACC]);
            }
        }
        float ar = pr[0], az = pz[0], an = pn[0];
#pragma unroll
        for (int i = 1; i < ACC; ++i) { ar += pr[i]; az += pz[i]; an += pn[i]; }
#pragma unroll
        for (int m = 1; m < KS; m <<= 1) {
            ar += __shfl_xor_sync(0xffffffffu, ar, m);
            az += __shfl_xor_sync(0xffffffffu, az, m);
            an += __shfl_xor_sync(0xffffffffu, an, m);
        }
        const float dr = 1.0f + ex2a(ar), dz = 1.0f + ex2a(az);
        const float ri = rcpa(dr * dz);
        const float r = dz * ri, z = dr * ri;
        const float dn = 1.0f + ex2a(fmaf(r, an, 0.1f));
        const float n = fmaf(-2.0f, rcpa(dn), 1.0f);
        h = fmaf(z, h - n, n);
        if (q == 0) hs[cur ^ 1][j] = h;
        cur ^= 1;
        __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0) cyc[0] = t1 - t0;
    out[tid] = h;
}

template <int KS, int ACC>
void run_chain(const float* w, float* out, long long* cyc, double ghz)
{
    chain_kernel<KS, ACC><<<1, 64 * KS>>>(w, out, cyc);
    CK(cudaDeviceSynchronize());
    chain_kernel<KS, ACC><<<1, 64 * KS>>>(w, out, cyc);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
    printf("GRU step chain, k-split %2d x %d accumulators (%4d threads, %2d dependent FMAs, %d shuffle levels): %7.1f clk/step = %6.1f ns at %.3f GHz\n",
           KS, ACC, 64 * KS, 64 / KS / ACC, KS == 1 ? 0 : (KS == 2 ? 1 : KS == 4 ? 2 : KS == 8 ? 3 : 4), (double)c / N_IT, (double)c / N_IT / ghz, ghz);
}

int main()
{
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const double ghz = clk_khz * 1e-6;
    float *w, *out;
    long long* cyc;
    CK(cudaMalloc(&w, 192 * 64 * sizeof(float)));
    CK(cudaMalloc(&out, 4096 * sizeof(float)));
    CK(cudaMalloc(&cyc, 8 * sizeof(long long)));
    float hw[192 * 64];
    srand(1);
    for (int i = 0; i < 192 * 64; ++i) hw[i] = 0.25f * ((float)rand() / RAND_MAX - 0.5f);
    CK(cudaMemcpy(w, hw, sizeof(hw), cudaMemcpyHostToDevice));
    prim_kernel<<<1, 32>>>(out, cyc, 0.5f);
    CK(cudaDeviceSynchronize());
    prim_kernel<<<1, 32>>>(out, cyc, 0.5f);
    CK(cudaDeviceSynchronize());
    long long c[6];
    CK(cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
    const char* names[6] = {"FFMA (dependent)", "SHFL.BFLY + FADD", "MUFU.EX2 + FADD", "MUFU.RCP + FADD", "LDS.128 (address-dependent)",
                            "STS -> BAR.SYNC -> LDS + FADD (1 warp)"};
    printf("# primitive dependent latencies (max SM clock %.3f GHz)\n", ghz);
    for (int i = 0; i < 6; ++i) printf("%-42s %6.1f clk\n", names[i], (double)c[i] / N_IT);
    printf("# minimal dependent chain of one GRU-HS[64] step at batch 1 (one SM)\n");
    run_chain<1, 1>(w, out, cyc, ghz);
    run_chain<1, 4>(w, out, cyc, ghz);
    run_chain<2, 1>(w, out, cyc, ghz);
    run_chain<2, 2>(w, out, cyc, ghz);
    run_chain<2, 4>(w, out, cyc, ghz);
    run_chain<4, 1>(w, out, cyc, ghz);
    run_chain<4, 2>(w, out, cyc, ghz);
    run_chain<8, 1>(w, out, cyc, ghz);
    run_chain<16, 1>(w, out, cyc, ghz);
    return 0;
}
