// Tensor-core persistent GRU kernel (tcgen05.mma + TMEM) -- the THROUGHPUT regime of the batched path (many streams per SM).
//
// Replaces, for many concurrent streams, the per-timestep loop behind `self.GRU(x, self.hidden)` + `self.output(x)`
// of RNN.forward / DiffDelRNN.forward (code/model.py:81-82, :412-413; gate equations torch rnn.py:1221-1224) and the
// delay read of code/model.py:422.
//
// Weight-stationary, transposed formulation.  Per timestep and per group of N streams the tensor core evaluates
//       G^T[192 x N] = W_hh[192 x 64] . H^T[64 x N]
// as three M=128 accumulator tiles, one per gate, each holding the gate's 64 rows TWICE (rows m and m+64 are the same
// hidden unit).
// Accumulator row m lives in TMEM lane m, column n is stream n, so the epilogue thread on TMEM lane L reads r, z and n
// pre-activations of hidden unit L % 64 straight out of TMEM with no cross-thread exchange.
//   A operand  packed once by ntm_gru_prepare (rows pre-scaled by -log2 e / 2 log2 e so the gates need a bare ex2) and
//              loaded into TENSOR MEMORY at kernel start (".ts" MMA form): the weights are never re-read from shared
//              memory, whose bandwidth otherwise bounds these small-N MMAs (measured 71 vs 40 clk per MMA at N = 64).
//   K = 80     the contraction is augmented by one k-step: B rows 64..68 hold [x_hi, x_lo, x_hi, 1, 1] of the step's
//              input sample (x split into two operand-format limbs) and the matching A columns hold
//              [w_i hi, w_i hi, w_i lo, b hi, b lo], so W_i x + b_i + b_h of the r and z gates, b_hn of the n gate and
//              the head bias come out of the MMA at ~2^-21 relative accuracy and cost the CUDA cores nothing.
//   B operand  the rounded state H, MN-major (streams contiguous): a thread owns one hidden unit (= one K row) and a
//              block of consecutive streams, so it publishes its new states with 16-byte stores.
//   state      the fp32 hidden state never leaves registers; only its rounded copy feeds the tensor core.
//   head       y = w_out . h' + b in fp32 from the registers (transposing warp butterfly + two partial sums per stream),
//              AFTER the thread has released the next MMA batch, i.e. off the recurrence's critical path.
// Roles per group: 8 epilogue warps (two per TMEM lane quarter, each taking half of the quarter's stream columns) +
// 1 MMA-issue warp (one elected lane).  G = 2 independent groups per CTA overlap one group's MMA latency with the other
// group's epilogue.  mbarrier hand-off: tcgen05.commit -> acc_full; fence.proxy.async + arrive (one per warp) -> h_ready.
// x / y are staged per 32-step chunk in shared memory.
#include <string.h>

#include "gates.cuh"
#include "tc_prims.cuh"

namespace ntm {

using namespace tc;

namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr int TC_TILES = 3;              // r, z, n gate tiles

template <int N, int G>
struct TcCfg {
    static constexpr int ELT = 2;                      // f16 / bf16 operands
    static constexpr int NK = 5;                       // MMAs along K = 64 + 16 (K = 16 each)
    static constexpr int KG = 2 * NK;                  // k groups of 8
    static constexpr int NS = N / 4;                   // streams per epilogue thread
    static constexpr int SC = 8;                       // streams per TMEM load / operand vector
    static constexpr int CH = 32;                      // steps per staged chunk
    static constexpr int EPI_WARPS = 8;
    static constexpr int NT = 32 * (EPI_WARPS + 1) * G;
    static constexpr int A_COLS = 16 * NK / 2;         // TMEM columns of one A tile (two 16-bit elements per column)
    static constexpr int A_ROW_WORDS = A_COLS;         // 32-bit words per row of the row-major image in the blob
    // B operand, MN-major, no swizzle: 16-byte vector = 8 consecutive streams of one k; 8 consecutive k = one 128-byte
    // core matrix; stream-vector groups at SBO, k groups of 8 at LBO
    static constexpr uint32_t B_SBO = 128, B_LBO = (N / 8) * 128, B_BYTES = KG * B_LBO;
    static constexpr int YS_LD = N + 1;                // padded: conflict-free flush reads
    // shared memory map (bytes)
    static constexpr uint32_t OFF_GRP = 0;
    static constexpr uint32_t GRP_B = 0;
    static constexpr int XROWS = CH + 1;               // a chunk also stages the first sample of its successor
    static constexpr uint32_t GRP_XS = B_BYTES;                        // [2][XROWS][N] floats
    static constexpr uint32_t GRP_YS = GRP_XS + 2 * XROWS * N * 4;        // [2 (unit halves)][CH][YS_LD] floats
    static constexpr uint32_t GRP_BYTES = (GRP_YS + 2 * CH * YS_LD * 4 + 127) / 128 * 128;
    static constexpr uint32_t OFF_BAR = OFF_GRP + G * GRP_BYTES;
    static constexpr uint32_t SMEM_BYTES = OFF_BAR + 64;
    static constexpr uint32_t TMEM_A = TC_TILES * N * G;   // first column of the resident A tiles
    static constexpr uint32_t TMEM_NEED = TMEM_A + TC_TILES * A_COLS;
    static_assert(TMEM_NEED <= 512, "accumulators + resident weights exceed the 512 TMEM columns");
    static constexpr uint32_t TMEM_COLS = TMEM_NEED <= 256 ? 256 : 512;
};

__device__ __forceinline__ void bar_sync_named(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int FMT>
__device__ __forceinline__ uint32_t pack_pair(float lo, float hi)
{
    if (FMT == FMT_BF16) return pack_bf16(lo, hi);
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// Sum p[i] over the 32 lanes of the warp for every i; lane l ends up with the total of stream (l >> (5 - log2 NS))
// (all lanes of that sub-group hold the same value).
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
__global__ void __launch_bounds__(32 * 9 * G, 1) gru_tc_kernel(const GruArgs a)
{
    using C = TcCfg<N, G>;
    constexpr int NS = C::NS, SC = C::SC, CH = C::CH, NK = C::NK, EW = C::EPI_WARPS;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);   // [g]: h_ready, [G+g]: acc_full
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_BAR + 48);

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;

    // ---- one-time setup: TMEM, barriers, weights -> TMEM ----------------------------------------------------
    if (warp == EW * G) {
        tmem_alloc(tmem_slot, C::TMEM_COLS);
        tmem_relinquish();
    }
    if (tid == 0) {
        for (int g = 0; g < G; ++g) {
            mbar_init(&bars[g], EW);         // one arrive per epilogue warp
            mbar_init(&bars[G + g], 1);      // tcgen05.commit
        }
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp < 4) {
        // W_hh / w_out tiles become TMEM-resident A operands: thread = row = TMEM lane, 40 columns per tile
        const uint32_t* img = reinterpret_cast<const uint32_t*>(a.blob + BlobLayout::tc_image(FMT));
        const int row = warp * 32 + lane;
#pragma unroll 1
        for (int tile = 0; tile < TC_TILES; ++tile)
#pragma unroll 1
            for (int c0 = 0; c0 < C::A_COLS; c0 += 8) {
                const uint4* src = reinterpret_cast<const uint4*>(img + (tile * 128 + row) * C::A_ROW_WORDS + c0);
                const uint4 v0 = src[0], v1 = src[1];
                const uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + C::TMEM_A + tile * C::A_COLS + c0, r);
            }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp >= EW * G) {
        // ================================ MMA issue warp of group g ===========================================
        const int g = warp - EW * G;
        const long long b0 = ((long long)blockIdx.x * G + g) * N;
        if (b0 < a.B && elect_one()) {
            constexpr uint32_t idesc = instr_desc(FMT, 128, N) | (1u << 16);      // B operand MN-major
            const uint32_t a_base = tmem + C::TMEM_A;
            const uint32_t b_base = smem_u32(smem + C::OFF_GRP + g * C::GRP_BYTES + C::GRP_B);
            const uint32_t d_base = tmem + (uint32_t)(g * TC_TILES * N);
            for (long long t = 0; t < a.T; ++t) {
                mbar_wait(&bars[g], (uint32_t)(t & 1));
                tc_fence_after();
#pragma unroll
                for (int tile = 0; tile < TC_TILES; ++tile) {
#pragma unroll
                    for (int ks = 0; ks < NK; ++ks)
                        mma_ts<FMT>(d_base + tile * N, a_base + tile * C::A_COLS + ks * 8,
                                    smem_desc(b_base + ks * 2 * C::B_LBO, C::B_LBO, C::B_SBO), idesc, ks > 0);
                }
                mma_commit(&bars[G + g]);
            }
        }
    } else {
        // ================================ epilogue warps of group g ===========================================
        const int g = warp / EW;
        const int wl = warp % EW;
        const int wq = wl & 3;                   // TMEM lane quarter (== warp id % 4)
        const int gt = tid - g * (EW * 32);      // thread index inside the group
        const int L = wq * 32 + lane;            // TMEM lane
        const int j = L & 63;                    // hidden unit
        const int sb = (L >> 6) * 2 + (wl >> 2); // which quarter of the group's streams
        const int s0 = sb * NS;                  // first stream of this thread
        const long long b0 = ((long long)blockIdx.x * G + g) * N;
        const int ns = (int)((a.B - b0) < (long long)N ? (a.B - b0) : (long long)N);   // may be <= 0
        if (ns > 0) {
            uint8_t* const grp = smem + C::OFF_GRP + g * C::GRP_BYTES;
            uint8_t* const bop = grp + C::GRP_B;
            float* const xs = reinterpret_cast<float*>(grp + C::GRP_XS);     // [2][CH][N]
            float* const ys = reinterpret_cast<float*>(grp + C::GRP_YS);     // [2][CH][YS_LD]
            const uint32_t d_base = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(g * TC_TILES * N);
            const int bar_id = 1 + g;
            constexpr int GT = EW * 32;

            const float* __restrict__ blob = a.blob;
            const float wo = blob[BlobLayout::W_OUT + j];
            const float bo = blob[BlobLayout::B_OUT];
            const float cn_w = 2.0f * LOG2E * blob[BlobLayout::W_IH + 128 + j];     // the only terms left to the CUDA cores:
            const float cn_b = 2.0f * LOG2E * blob[BlobLayout::B_IH + 128 + j];     // W_in x + b_in (scaled like the n rows)

            const bool delay = a.d != nullptr;
            float* __restrict__ head_out = delay ? a.pre : a.y;
            const long long ldo = delay ? a.ldp : a.ldy;

            auto load_x = [&](int buf, long long t0) {
                const int n = (int)((a.T - t0) < (long long)C::XROWS ? (a.T - t0) : (long long)C::XROWS);
                float* dstb = xs + buf * C::XROWS * N;
                for (int idx = gt; idx < C::XROWS * N; idx += GT) {
                    const int s = idx % N, tt = idx / N;
                    if (s < ns && tt < n) cp_async4(dstb + tt * N + s, a.x + (b0 + s) * a.ldx + t0 + tt);
                    else dstb[tt * N + s] = 0.0f;
                }
                cp_async_commit();
            };
            // rounded states of 8 consecutive streams of unit j -> one 16-byte vector of the MN-major B operand
            auto store_operand8 = [&](int s, const float* v) {
                uint4 w;
                w.x = pack_pair<FMT>(v[0], v[1]); w.y = pack_pair<FMT>(v[2], v[3]);
                w.z = pack_pair<FMT>(v[4], v[5]); w.w = pack_pair<FMT>(v[6], v[7]);
                *reinterpret_cast<uint4*>(bop + (j & 7) * 16 + (j >> 3) * C::B_LBO + (s >> 3) * C::B_SBO) = w;
            };
            // K augmentation: B rows 64 (x_hi), 65 (x_lo), 66 (x_hi) of this thread's streams for the step that consumes
            // sample row `xrow`; done by the threads of units 0..2 (one B row each)
            auto store_x_aug = [&](const float* xrow) {
                if (j < 3) {
#pragma unroll
                    for (int c0 = 0; c0 < NS; c0 += SC) {
                        float v[SC];
#pragma unroll
                        for (int i = 0; i < SC; ++i) {
                            const float x = xrow[s0 + c0 + i];
                            const float hi = FMT == FMT_BF16 ? __bfloat162float(__float2bfloat16_rn(x))
                                                             : __half2float(__float2half_rn(x));
                            v[i] = j == 1 ? x - hi : hi;
                        }
                        uint4 w;
                        w.x = pack_pair<FMT>(v[0], v[1]); w.y = pack_pair<FMT>(v[2], v[3]);
                        w.z = pack_pair<FMT>(v[4], v[5]); w.w = pack_pair<FMT>(v[6], v[7]);
                        *reinterpret_cast<uint4*>(bop + j * 16 + 8 * C::B_LBO + ((s0 + c0) >> 3) * C::B_SBO) = w;
                    }
                }
            };
            auto flush = [&](long long t0, int n, const float* xchunk) {
                for (int idx = gt; idx < N * CH; idx += GT) {
                    const int s = idx / CH, tt = idx % CH;
                    if (s < ns && tt < n) {
                        float v = ys[tt * C::YS_LD + s] + ys[(CH + tt) * C::YS_LD + s] + bo;
                        if (a.skip) v += xchunk[tt * N + s];
                        head_out[(b0 + s) * ldo + t0 + tt] = v;
                        if (delay && a.warmup) a.y[(b0 + s) * a.ldy + t0 + tt] = v;
                    }
                }
                if (delay && !a.warmup) {
                    bar_sync_named(bar_id, GT);      // this chunk's pre_d is visible group-wide (L2 reads below)
                    for (int idx = gt; idx < N * CH; idx += GT) {
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
            };

            // ---- initial state: fp32 in registers, rounded copy into the B operand -------------------------
            float hst[NS];
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const int s = s0 + i;
                hst[i] = (s < ns && a.h_in) ? a.h_in[(b0 + s) * 64 + j] : 0.0f;
            }
#pragma unroll
            for (int c0 = 0; c0 < NS; c0 += SC) store_operand8(s0 + c0, hst + c0);
            // constant B rows 67, 68 (= 1: bias columns) and 69..79 (= 0), written once
            for (int idx = gt; idx < 13 * (N / 8); idx += GT) {
                const int k = 67 + idx / (N / 8), v8 = idx % (N / 8);
                const uint32_t one2 = pack_pair<FMT>(1.0f, 1.0f), val = k < 69 ? one2 : 0u;
                *reinterpret_cast<uint4*>(bop + (k & 7) * 16 + (k >> 3) * C::B_LBO + v8 * C::B_SBO) = make_uint4(val, val, val, val);
            }
            load_x(0, 0);
            cp_async_wait_all();
            bar_sync_named(bar_id, GT);          // xs[0] visible: the first step's x rows can be published
            store_x_aug(xs);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[g]);

            const long long nchunks = (a.T + CH - 1) / CH;
            long long t = 0;
            for (long long c = 0; c < nchunks; ++c) {
                const long long t0 = c * CH;
                const int n = (int)((a.T - t0) < (long long)CH ? (a.T - t0) : (long long)CH);
                const int xb = (int)(c & 1);
                const float* xcur = xs + xb * C::XROWS * N;

                cp_async_wait_all();
                bar_sync_named(bar_id, GT);          // xs[xb] landed; the previous flush is done with ys and xs[xb ^ 1]
                if (c + 1 < nchunks) load_x(xb ^ 1, t0 + CH);

                for (int tt = 0; tt < n; ++tt, ++t) {
                    mbar_wait(&bars[G + g], (uint32_t)(t & 1));
                    tc_fence_after();
                    float p[NS];
#pragma unroll
                    for (int c0 = 0; c0 < NS; c0 += SC) {
                        uint32_t ar[SC], az[SC], an[SC];
                        tmem_ld8(d_base + s0 + c0, ar);
                        tmem_ld8(d_base + N + s0 + c0, az);
                        tmem_ld8(d_base + 2 * N + s0 + c0, an);
                        float xv[SC], hn[SC];
#pragma unroll
                        for (int i = 0; i < SC; i += 4) {
                            const float4 v = *reinterpret_cast<const float4*>(xcur + tt * N + s0 + c0 + i);
                            xv[i] = v.x; xv[i + 1] = v.y; xv[i + 2] = v.z; xv[i + 3] = v.w;
                        }
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < SC; ++i) {
                            // accumulators already hold the complete scaled pre-activations of r, z and W_hn h + b_hn
                            const float r = rcp_approx(1.0f + ex2_approx(__uint_as_float(ar[i])));
                            const float z = rcp_approx(1.0f + ex2_approx(__uint_as_float(az[i])));
                            const float en = ex2_approx(fmaf(r, __uint_as_float(an[i]), fmaf(cn_w, xv[i], cn_b)));
                            const float nn = fmaf(-2.0f, rcp_approx(1.0f + en), 1.0f);
                            hn[i] = fmaf(z, hst[c0 + i] - nn, nn);
                            hst[c0 + i] = hn[i];
                            p[c0 + i] = wo * hn[i];
                        }
                        store_operand8(s0 + c0, hn);
                    }
                    store_x_aug(xcur + (tt + 1) * N);      // input sample of the NEXT step (row CH = successor chunk's first)
                    // release the next MMA batch of this group
                    tc_fence_before();
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars[g]);
                    // fp32 output head, off the critical path: sum w_out[j] * h'[j] over this warp's 32 hidden units
                    const float tot = warp_transpose_reduce<NS>(p, lane);
                    if ((lane & (32 / NS - 1)) == 0) ys[((wq & 1) * CH + tt) * C::YS_LD + s0 + lane / (32 / NS)] = tot;
                }
                bar_sync_named(bar_id, GT);
                flush(t0, n, xcur);
            }

            // ---- final state; rolled delay history (code/model.py:314-315) -----------------------------------
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const int s = s0 + i;
                if (s < ns) a.h_out[(b0 + s) * 64 + j] = hst[i];
            }
            if (delay) {
                bar_sync_named(bar_id, GT);
                for (long long idx = gt; idx < (long long)ns * a.D; idx += GT) {
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

    // ---- teardown: every MMA has completed (the epilogue consumed the last accumulator) ------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == EW * G) tmem_dealloc(tmem, TcCfg<N, G>::TMEM_COLS);
}

template <int FMT, int N, int G>
cudaError_t launch_tc_one(const GruArgs& a, cudaStream_t st)
{
    using C = TcCfg<N, G>;
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
    if (g == 2) return n == 32 ? launch_tc_one<FMT, 32, 2>(a, st) : launch_tc_one<FMT, 64, 2>(a, st);
    return n == 32 ? launch_tc_one<FMT, 32, 1>(a, st) : launch_tc_one<FMT, 64, 1>(a, st);
}

}  // namespace

// Host side of the A operand (called by ntm_gru_prepare once the fp32 part of the blob is filled).  For each operand
// format: tiles 0..2 = r, z, n gates, rows m = 0..127 = hidden unit m % 64 (PyTorch row gate*64 + unit), scaled so
// that the epilogue needs a bare ex2:
//   r, z rows by -log2(e)   (sigmoid(a) = 1 / (1 + 2^(-a log2 e)))        n rows by 2 log2(e)   (tanh(a) = 1 - 2 / (1 + 2^(2 a log2 e)))
// columns 0..63 = W_h*, columns 64..68 = [w_i hi, w_i hi, w_i lo, b hi, b lo] matching the B rows [x_hi, x_lo, x_hi, 1, 1]
// (n tile: w_i = 0 and b = b_hn only, because W_in x + b_in sits outside the r * (...) product).
void pack_tc_images(float* blob_host)
{
    using L = BlobLayout;
    const float* w_hh = blob_host + L::W_HH;
    const float* w_ih = blob_host + L::W_IH;
    const float* b_ih = blob_host + L::B_IH;
    const float* b_hh = blob_host + L::B_HH;
    const uint32_t row_bytes = 80 * 2, tile_bytes = 128 * row_bytes;     // row-major [tile][row m][k], k contiguous
    for (int fmt = 0; fmt < 2; ++fmt) {
        uint8_t* img = reinterpret_cast<uint8_t*>(blob_host + L::tc_image(fmt));
        memset(img, 0, TC_TILES * tile_bytes);
        auto rnd = [&](float v) {
            return fmt == FMT_BF16 ? __bfloat162float(__float2bfloat16_rn(v)) : __half2float(__float2half_rn(v));
        };
        auto put = [&](int tile, int m, int k, float v) {
            uint16_t bits;
            if (fmt == FMT_BF16) bits = static_cast<__nv_bfloat16_raw>(__float2bfloat16_rn(v)).x;
            else bits = static_cast<__half_raw>(__float2half_rn(v)).x;
            memcpy(img + tile * tile_bytes + m * row_bytes + k * 2, &bits, 2);
        };
        auto put_split = [&](int tile, int m, int k_hi, int k_lo, float v) {
            const float hi = rnd(v);
            put(tile, m, k_hi, hi);
            put(tile, m, k_lo, v - hi);
        };
        for (int gate = 0; gate < 3; ++gate) {
            const float scale = gate < 2 ? -LOG2E : 2.0f * LOG2E;
            for (int m = 0; m < 128; ++m) {
                const int row = gate * 64 + (m & 63);
                for (int k = 0; k < 64; ++k) put(gate, m, k, scale * w_hh[row * 64 + k]);
                if (gate < 2) {
                    const float wi = scale * w_ih[row];
                    put(gate, m, 64, wi);                 // x_hi * w_hi
                    put_split(gate, m, 65, 66, wi);       // x_lo * w_hi + x_hi * w_lo
                    put_split(gate, m, 67, 68, scale * (b_ih[row] + b_hh[row]));
                } else {
                    put_split(gate, m, 67, 68, scale * b_hh[row]);
                }
            }
        }
    }
}

// fmt: FMT_F16 / FMT_BF16 (tf32 operands run on the mma.sync kernel).  tune_n / tune_g: streams per group (32, 64) /
// groups per CTA (1, 2); 0 = automatic.
cudaError_t launch_gru_tc(const GruArgs& a, int fmt, int sm_count, int tune_n, int tune_g, cudaStream_t st)
{
    if (a.B <= 0 || a.T <= 0) return cudaSuccess;
    int n = tune_n, g = tune_g;
    if (n <= 0 || g <= 0) {
        const long long per_sm = (a.B + sm_count - 1) / sm_count;
        g = 2;
        n = per_sm > 64 ? 64 : 32;
    }
    if (n != 32) n = 64;
    if (g != 2) g = 1;
    return fmt == FMT_BF16 ? launch_tc_fmt<FMT_BF16>(a, n, g, st) : launch_tc_fmt<FMT_F16>(a, n, g, st);
}

}  // namespace ntm
