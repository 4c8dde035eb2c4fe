// Tensor-core persistent GRU kernel (tcgen05.mma + TMEM) -- the batched throughput path.
//
// Replaces, for many concurrent streams, the per-timestep loop behind `self.GRU(x, self.hidden)` + `self.output(x)`
// of RNN.forward / DiffDelRNN.forward (code/model.py:81-82, :412-413; gate equations torch rnn.py:1221-1224) and the
// delay read of code/model.py:422.
//
// Weight-stationary, transposed formulation.  Per timestep and per group of N streams one CTA evaluates
//       G^T[192 x N] = W_hh[192 x 64] . H^T[64 x N]
// as three M=128 accumulator tiles, one per gate, each holding the gate's 64 rows TWICE (rows m and m+64 are the
// same hidden unit).  Accumulator row m lives in TMEM lane m, column n is stream n, so thread `lane` of the
// epilogue reads r, z and n pre-activations of ITS hidden unit straight out of TMEM with no cross-thread
// exchange; lanes 0..63 take streams [0, N/2), lanes 64..127 streams [N/2, N).
//   A operand  W_hh (rounded once to f16 / bf16 / tf32, pre-scaled by -log2(e) resp. 2 log2(e) so that the gates
//              need a bare ex2) is packed by ntm_gru_prepare and sits in shared memory for the whole kernel.
//   B operand  H (N x 64, K-major) is rewritten in shared memory by the epilogue every step.
//   state      the fp32 hidden state never leaves registers; only its rounded copy feeds the tensor core.
//   head       y = w_out . h + b is evaluated in fp32 from the registers with a transposing warp butterfly, AFTER
//              the thread has released the next MMA, i.e. off the recurrence's critical path.
// Roles: 4 epilogue warps + 1 MMA-issue warp per group; G (1 or 2) independent groups per CTA overlap one
// group's MMA latency with the other group's epilogue.  x / y are staged per 32-step chunk in shared memory.
#include <string.h>

#include "gates.cuh"
#include "tc_prims.cuh"

namespace ntm {

using namespace tc;

namespace {

constexpr float LOG2E = 1.4426950408889634f;

template <int FMT, int N, int G>
struct TcCfg {
    static constexpr int ELT = FMT == FMT_TF32 ? 4 : 2;
    static constexpr int KCH = 64 * ELT / 16;          // 16-byte K chunks per operand row
    static constexpr int NK = KCH / 2;                 // MMAs along K
    static constexpr int NS = N / 2;                   // streams per thread
    static constexpr int SC = NS < 8 ? NS : 8;         // streams per TMEM load
    static constexpr int CH = 32;                      // steps per staged chunk
    static constexpr int NT = 32 * 5 * G;
    static constexpr uint32_t A_LBO = 128 * 16, A_SBO = 128, A_TILE = 128 * 64 * ELT;
    static constexpr uint32_t B_LBO = N * 16 + 16, B_SBO = 128;   // +16: conflict-free element stores
    static constexpr uint32_t B_BYTES = KCH * B_LBO;
    static constexpr int YP_LD = N + 1;
    // shared memory map (bytes)
    static constexpr uint32_t OFF_A = 0;
    static constexpr uint32_t OFF_GRP = 3 * A_TILE;
    static constexpr uint32_t GRP_B = 0;
    static constexpr uint32_t GRP_XS = (B_BYTES + 127) / 128 * 128;
    static constexpr uint32_t GRP_YP = GRP_XS + 2 * CH * N * 4;
    static constexpr uint32_t GRP_BYTES = (GRP_YP + 2 * CH * YP_LD * 4 + 127) / 128 * 128;
    static constexpr uint32_t OFF_BAR = OFF_GRP + G * GRP_BYTES;
    static constexpr uint32_t SMEM_BYTES = OFF_BAR + 64;
    static constexpr uint32_t TMEM_COLS = 3 * N * G <= 32 ? 32 : 3 * N * G <= 64 ? 64 : 3 * N * G <= 128 ? 128 :
                                          3 * N * G <= 256 ? 256 : 512;
};

template <int NREG>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, uint32_t (&r)[NREG]);
template <>
__device__ __forceinline__ void tmem_ldn<4>(uint32_t taddr, uint32_t (&r)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
template <>
__device__ __forceinline__ void tmem_ldn<8>(uint32_t taddr, uint32_t (&r)[8]) { tmem_ld8(taddr, r); }

template <int FMT>
__device__ __forceinline__ void store_operand(uint8_t* p, float v)
{
    if (FMT == FMT_TF32) *reinterpret_cast<uint32_t*>(p) = to_tf32(v);
    else if (FMT == FMT_BF16) *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(v);
    else *reinterpret_cast<__half*>(p) = __float2half_rn(v);
}

__device__ __forceinline__ void bar_sync_named(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Sum p[i] over the 32 lanes of the warp for every i; lane l ends up with the total of stream
// (l >> (5 - log2 NS)) (all lanes of that sub-group hold the same value).
template <int NS>
__device__ __forceinline__ float warp_transpose_reduce(float (&p)[NS], int lane)
{
    int mask = 16;
#pragma unroll
    for (int n = NS; n > 1; n >>= 1) {
        const bool up = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float keep = up ? p[i + n / 2] : p[i];
            const float send = up ? p[i] : p[i + n / 2];
            p[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
        mask >>= 1;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1)
        if (m <= mask) p[0] += __shfl_xor_sync(0xffffffffu, p[0], m);
    return p[0];
}

template <int FMT, int N, int G>
__global__ void __launch_bounds__(32 * 5 * G, 1) gru_tc_kernel(const GruArgs a)
{
    using C = TcCfg<FMT, N, G>;
    constexpr int NS = C::NS, SC = C::SC, CH = C::CH, NK = C::NK;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);   // [g]: h_ready, [G+g]: acc_full
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_BAR + 48);

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;

    // ---- one-time setup: weights image -> smem, TMEM, barriers ---------------------------------------------
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.blob + BlobLayout::tc_image(FMT));
        uint4* dst = reinterpret_cast<uint4*>(smem + C::OFF_A);
        for (int i = tid; i < (int)(3 * C::A_TILE / 16); i += C::NT) dst[i] = src[i];
    }
    if (warp == 4 * G) {
        tmem_alloc(tmem_slot, C::TMEM_COLS);
        tmem_relinquish();
    }
    if (tid == 0) {
        for (int g = 0; g < G; ++g) {
            mbar_init(&bars[g], 4);          // one arrive per epilogue warp
            mbar_init(&bars[G + g], 1);      // tcgen05.commit
        }
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp >= 4 * G) {
        // ================================ MMA issue warp of group g ===========================================
        const int g = warp - 4 * G;
        const long long b0 = ((long long)blockIdx.x * G + g) * N;
        if (b0 < a.B && elect_one()) {
            constexpr uint32_t idesc = instr_desc(FMT, 128, N);
            const uint32_t a_base = smem_u32(smem + C::OFF_A);
            const uint32_t b_base = smem_u32(smem + C::OFF_GRP + g * C::GRP_BYTES + C::GRP_B);
            const uint32_t d_base = tmem + (uint32_t)(g * 3 * N);
            for (long long t = 0; t < a.T; ++t) {
                mbar_wait(&bars[g], (uint32_t)(t & 1));
                tc_fence_after();
#pragma unroll
                for (int gate = 0; gate < 3; ++gate)
#pragma unroll
                    for (int ks = 0; ks < NK; ++ks)
                        mma_ss<FMT>(d_base + gate * N,
                                    smem_desc(a_base + gate * C::A_TILE + ks * 2 * C::A_LBO, C::A_LBO, C::A_SBO),
                                    smem_desc(b_base + ks * 2 * C::B_LBO, C::B_LBO, C::B_SBO), idesc, ks > 0);
                mma_commit(&bars[G + g]);
            }
        }
    } else {
        // ================================ epilogue warps of group g ===========================================
        const int g = warp >> 2;
        const int wq = warp & 3;                 // TMEM lane quarter
        const int gt = tid - g * 128;            // thread index inside the group = TMEM lane
        const int j = gt & 63;                   // hidden unit
        const int half = gt >> 6;                // which half of the group's streams
        const int wa = wq & 1;                   // which 32-unit block of the head partial sums
        const long long b0 = ((long long)blockIdx.x * G + g) * N;
        const int ns = (int)((a.B - b0) < (long long)N ? (a.B - b0) : (long long)N);   // may be <= 0
        if (ns > 0) {
            uint8_t* const grp = smem + C::OFF_GRP + g * C::GRP_BYTES;
            uint8_t* const bop = grp + C::GRP_B;
            float* const xs = reinterpret_cast<float*>(grp + C::GRP_XS);     // [2][CH][N]
            float* const yp = reinterpret_cast<float*>(grp + C::GRP_YP);     // [2][CH][YP_LD]
            const uint32_t d_base = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(g * 3 * N + half * NS);

            // per-unit constants, pre-scaled like the packed weights
            const float* __restrict__ blob = a.blob;
            const UnitConst uc = load_unit_const(blob, j);
            const float wo = uc.wo;
            const float bo = blob[BlobLayout::B_OUT];

            const bool delay = a.d != nullptr;
            float* __restrict__ head_out = delay ? a.pre : a.y;
            const long long ldo = delay ? a.ldp : a.ldy;

            auto load_x = [&](int buf, long long t0) {
                const int n = (int)((a.T - t0) < (long long)CH ? (a.T - t0) : (long long)CH);
                float* dstb = xs + buf * CH * N;
                for (int idx = gt; idx < CH * N; idx += 128) {
                    const int s = idx % N, tt = idx / N;
                    if (s < ns && tt < n) cp_async4(dstb + tt * N + s, a.x + (b0 + s) * a.ldx + t0 + tt);
                    else dstb[tt * N + s] = 0.0f;
                }
                cp_async_commit();
            };

            // ---- initial state: fp32 in registers, rounded copy into the B operand -------------------------
            float hst[NS];
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const int s = half * NS + i;
                hst[i] = (s < ns && a.h_in) ? a.h_in[(b0 + s) * 64 + j] : 0.0f;
                store_operand<FMT>(bop + kmajor_off<C::ELT>(s, j, C::B_LBO, C::B_SBO), hst[i]);
            }
            load_x(0, 0);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[g]);

            const long long nchunks = (a.T + CH - 1) / CH;
            long long t = 0;
            for (long long c = 0; c < nchunks; ++c) {
                const long long t0 = c * CH;
                const int n = (int)((a.T - t0) < (long long)CH ? (a.T - t0) : (long long)CH);
                const int xb = (int)(c & 1);
                const float* xcur = xs + xb * CH * N;
                cp_async_wait_all();
                bar_sync_named(1 + g, 128);          // xs[xb] landed; previous flush finished reading yp / xs
                if (c + 1 < nchunks) load_x(xb ^ 1, t0 + CH);

                for (int tt = 0; tt < n; ++tt, ++t) {
                    mbar_wait(&bars[G + g], (uint32_t)(t & 1));
                    tc_fence_after();
                    float p[NS];
#pragma unroll
                    for (int c0 = 0; c0 < NS; c0 += SC) {
                        uint32_t ar[SC], az[SC], an[SC];
                        tmem_ldn<SC>(d_base + c0, ar);
                        tmem_ldn<SC>(d_base + N + c0, az);
                        tmem_ldn<SC>(d_base + 2 * N + c0, an);
                        float xv[SC];
#pragma unroll
                        for (int i = 0; i < SC; i += 4) {
                            const float4 v = *reinterpret_cast<const float4*>(xcur + tt * N + half * NS + c0 + i);
                            xv[i] = v.x; xv[i + 1] = v.y; xv[i + 2] = v.z; xv[i + 3] = v.w;
                        }
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < SC; i += 2) {
                            float z0, z1, dn0, dn1, hn0, hn1;
                            gates_rz_dn(uc, __uint_as_float(ar[i]), __uint_as_float(az[i]), __uint_as_float(an[i]), xv[i], z0, dn0);
                            gates_rz_dn(uc, __uint_as_float(ar[i + 1]), __uint_as_float(az[i + 1]), __uint_as_float(an[i + 1]),
                                        xv[i + 1], z1, dn1);
                            hn0 = gates_blend1(z0, dn0, hst[c0 + i]);
                            hn1 = gates_blend1(z1, dn1, hst[c0 + i + 1]);
                            hst[c0 + i] = hn0;
                            hst[c0 + i + 1] = hn1;
                            p[c0 + i] = wo * hn0;
                            p[c0 + i + 1] = wo * hn1;
                            store_operand<FMT>(bop + kmajor_off<C::ELT>(half * NS + c0 + i, j, C::B_LBO, C::B_SBO), hn0);
                            store_operand<FMT>(bop + kmajor_off<C::ELT>(half * NS + c0 + i + 1, j, C::B_LBO, C::B_SBO), hn1);
                        }
                    }
                    // release the next MMA of this group, then do the head off the critical path
                    tc_fence_before();
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars[g]);
                    const float tot = warp_transpose_reduce<NS>(p, lane);
                    if ((lane & (32 / NS - 1)) == 0) yp[(wa * CH + tt) * C::YP_LD + half * NS + (lane / (32 / NS))] = tot;
                }

                // ---- flush the chunk: y = sum of the two 32-unit partials + bias (+ x) -------------------------
                bar_sync_named(1 + g, 128);
                for (int idx = gt; idx < N * CH; idx += 128) {
                    const int s = idx / CH, tt = idx % CH;
                    if (s < ns && tt < n) {
                        float v = yp[tt * C::YP_LD + s] + yp[(CH + tt) * C::YP_LD + s] + bo;
                        if (a.skip) v += xcur[tt * N + s];
                        head_out[(b0 + s) * ldo + t0 + tt] = v;
                        if (delay && a.warmup) a.y[(b0 + s) * a.ldy + t0 + tt] = v;
                    }
                }
                if (delay && !a.warmup) {
                    bar_sync_named(1 + g, 128);      // this chunk's pre_d is visible group-wide (L2 reads below)
                    for (int idx = gt; idx < N * CH; idx += 128) {
                        const int s = idx / CH, tt = idx % CH;
                        if (s < ns && tt < n) {
                            const long long tg = t0 + tt;
                            const float* prow = a.pre + (b0 + s) * a.ldp;
                            const float* hrow = a.hist_in + (b0 + s) * (long long)a.D;
                            a.y[(b0 + s) * a.ldy + tg] =
                                delay_read(a.d[(b0 + s) * a.ldd + tg], tg, a.D,
                                           [&](long long i) { return i >= 0 ? __ldcg(prow + i) : hrow[a.D + i]; });
                        }
                    }
                }
            }

            // ---- final state; rolled delay history (code/model.py:314-315) -----------------------------------
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const int s = half * NS + i;
                if (s < ns) a.h_out[(b0 + s) * 64 + j] = hst[i];
            }
            if (delay) {
                bar_sync_named(1 + g, 128);
                for (long long idx = gt; idx < (long long)ns * a.D; idx += 128) {
                    const int s = (int)(idx / a.D);
                    const long long i = idx % a.D;
                    const long long src = a.T - a.D + i;
                    a.hist_out[(b0 + s) * (long long)a.D + i] =
                        src >= 0 ? __ldcg(a.pre + (b0 + s) * a.ldp + src)
                                 : a.hist_in[(b0 + s) * (long long)a.D + a.D + src];
                }
            }
        }
    }

    // ---- teardown: every MMA has completed (the epilogue consumed its last accumulator) ------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 4 * G) tmem_dealloc(tmem, TcCfg<FMT, N, G>::TMEM_COLS);
}

template <int FMT, int N, int G>
cudaError_t launch_tc_one(const GruArgs& a, cudaStream_t st)
{
    using C = TcCfg<FMT, N, G>;
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && !configured[dev]) {
        e = cudaFuncSetAttribute(gru_tc_kernel<FMT, N, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)C::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const long long grid = (a.B + (long long)N * G - 1) / ((long long)N * G);
    gru_tc_kernel<FMT, N, G><<<(unsigned)grid, C::NT, C::SMEM_BYTES, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

template <int FMT>
cudaError_t launch_tc_fmt(const GruArgs& a, int n, int g, cudaStream_t st)
{
    if (g == 2) {
        switch (n) {
            case 8: return launch_tc_one<FMT, 8, 2>(a, st);
            case 16: return launch_tc_one<FMT, 16, 2>(a, st);
            case 32: return launch_tc_one<FMT, 32, 2>(a, st);
            default: return launch_tc_one<FMT, 64, 2>(a, st);
        }
    }
    switch (n) {
        case 8: return launch_tc_one<FMT, 8, 1>(a, st);
        case 16: return launch_tc_one<FMT, 16, 1>(a, st);
        case 32: return launch_tc_one<FMT, 32, 1>(a, st);
        default: return launch_tc_one<FMT, 64, 1>(a, st);
    }
}

}  // namespace

// Host side of the A operand: for every operand format, gate tile `gate` holds rows m = 0..127 = hidden unit m % 64
// of that gate (PyTorch row gate*64 + unit), scaled so that the epilogue needs a bare ex2:
//   r, z rows by -log2(e)   (sigmoid(a) = 1 / (1 + 2^(-a log2 e)))
//   n rows    by 2 log2(e)  (tanh(a)    = 1 - 2 / (1 + 2^(2 a log2 e)))
void pack_tc_images(const float* w_hh, float* blob_host)
{
    const uint32_t lbo = 128 * 16, sbo = 128;
    uint8_t* f16 = reinterpret_cast<uint8_t*>(blob_host + BlobLayout::IMG_F16);
    uint8_t* b16 = reinterpret_cast<uint8_t*>(blob_host + BlobLayout::IMG_BF16);
    uint8_t* t32 = reinterpret_cast<uint8_t*>(blob_host + BlobLayout::IMG_TF32);
    for (int gate = 0; gate < 3; ++gate) {
        const float scale = gate < 2 ? -LOG2E : 2.0f * LOG2E;
        for (int m = 0; m < 128; ++m)
            for (int k = 0; k < 64; ++k) {
                const float v = scale * w_hh[(gate * 64 + (m & 63)) * 64 + k];
                const __half_raw h = static_cast<__half_raw>(__float2half_rn(v));
                const __nv_bfloat16_raw b = static_cast<__nv_bfloat16_raw>(__float2bfloat16_rn(v));
                uint32_t u;
                memcpy(&u, &v, 4);
                u = (u + 0x1000u) & 0xFFFFE000u;                 // cvt.rna.tf32.f32
                memcpy(f16 + gate * (128 * 64 * 2) + kmajor_off<2>(m, k, lbo, sbo), &h.x, 2);
                memcpy(b16 + gate * (128 * 64 * 2) + kmajor_off<2>(m, k, lbo, sbo), &b.x, 2);
                memcpy(t32 + gate * (128 * 64 * 4) + kmajor_off<4>(m, k, lbo, sbo), &u, 4);
            }
    }
}

// fmt: FMT_F16 / FMT_BF16 / FMT_TF32.  tune_n / tune_g: streams per group / groups per CTA (0 = automatic).
cudaError_t launch_gru_tc(const GruArgs& a, int fmt, int sm_count, int tune_n, int tune_g, cudaStream_t st)
{
    if (a.B <= 0 || a.T <= 0) return cudaSuccess;
    int n = tune_n, g = tune_g;
    if (n <= 0 || g <= 0) {
        // spread the streams over all SMs first; once an SM owns >= 16 streams split them into two groups so that
        // one group's MMA latency hides behind the other group's epilogue
        const long long per_sm = (a.B + sm_count - 1) / sm_count;
        g = per_sm >= 16 ? 2 : 1;
        const long long per_grp = (per_sm + g - 1) / g;
        n = 8;
        while (n < 64 && n < per_grp) n <<= 1;
    }
    if (n != 8 && n != 16 && n != 32) n = 64;
    if (g != 2) g = 1;
    switch (fmt) {
        case FMT_TF32: return launch_tc_fmt<FMT_TF32>(a, n, g, st);
        case FMT_BF16: return launch_tc_fmt<FMT_BF16>(a, n, g, st);
        default: return launch_tc_fmt<FMT_F16>(a, n, g, st);
    }
}

}  // namespace ntm
