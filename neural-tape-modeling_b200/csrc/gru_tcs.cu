// Stream-major tensor-core persistent GRU kernel (tcgen05.mma + TMEM) -- the THROUGHPUT regime of the batched path
// (>= ~128 streams per SM; BASELINE cfg 4: 65 536 streams).
//
// Replaces, for many concurrent streams, the per-timestep loop behind `self.GRU(x, self.hidden)` + `self.output(x)` of
// RNN.forward (code/model.py:81-82; gate equations torch rnn.py:1221-1224).
//
// Formulation.  One tile = 128 streams = the M rows of the MMA.  Per timestep the tensor core evaluates
//       G[128 streams x 192] = [H | x_hi x_lo x_hi 1 1 0..][128 x 80] . W^T[80 x 192]
//   A operand  the rounded state of the tile, IN TENSOR MEMORY (".ts" MMA form): TMEM lane = stream, so the thread that
//              owns a stream writes its new rounded state with tcgen05.st straight from registers -- no shared-memory
//              round trip, no proxy fence, no cross-thread exchange anywhere in the step.
//   B operand  W_hh (+ the K augmentation carrying W_i x + b for r, z and b_hn for n), K-major, staged once in shared
//              memory from the image ntm_gru_prepare packed (rows pre-scaled by -log2 e / 2 log2 e).
//   N = 192    five MMAs of M=128, N=192, K=16 per tile-step run at the tensor pipe's full rate (measured 106 clk each,
//              profiles/r01_tc_probe.txt) = 4 clk per stream-step, against >= 16 clk per stream-step of MUFU work: the
//              tensor pipe is never the bound here, which the weight-stationary kernel (gru_tc.cu) could not achieve.
//   epilogue   accumulator lane = stream, column = gate row: ONE THREAD OWNS ONE STREAM -- its 64 fp32 states live in
//              registers for the whole launch, r/z/n of every unit arrive by tcgen05.ld, per-unit constants are
//              warp-uniform (kernel-parameter constant bank), the output head is a plain in-thread fp32 dot product,
//              and reciprocals are shared between the r and z gates of a unit and the n gates of two units
//              (4.5 MUFU per unit-step, gates.cuh).
// Roles: 4 epilogue warps per tile (one per TMEM lane quarter) + 1 MMA-issue warp per tile; two tiles per CTA ping-pong
// so one tile's MMA + hand-off latency hides behind the other tile's gate math.  mbarrier hand-off:
//   epilogue: tcgen05.st (new A) -> wait::st -> fence::before_thread_sync -> arrive(h_ready[tile])      (count 4)
//   issuer:   wait(h_ready) -> fence::after_thread_sync -> 5 x tcgen05.mma -> tcgen05.commit(acc_full[tile])
// x is prefetched two steps ahead by the owning thread, y is written by the owning thread (partial sectors merge in L2).
#include <math.h>

#include "gates.cuh"
#include "tc_prims.cuh"

namespace ntm {

using namespace tc;

namespace {

constexpr int TS_M = 128;                  // streams per tile
constexpr int TS_N = 192;                  // gate rows
constexpr int TS_NK = 5;                   // MMAs along K = 64 + 16
constexpr int TS_TILE_COLS = 256;          // TMEM columns reserved per tile: 192 accumulator + 40 operand (+ 24 spare)
constexpr int TS_A_OFF = 192;              // first operand column inside a tile's TMEM block
constexpr uint32_t TS_SBO = 128, TS_LBO = (TS_N / 8) * 128, TS_B_BYTES = 2 * TS_NK * TS_LBO;   // K-major, no swizzle
constexpr uint32_t TS_OFF_BAR = TS_B_BYTES;                       // 6 mbarriers + the TMEM base slot
constexpr uint32_t TS_OFF_YP = TS_OFF_BAR + 64;                   // [2 tiles][2][128] floats
constexpr uint32_t TS_SMEM_BYTES = TS_OFF_YP + 2 * 2 * 128 * 4;
constexpr int TS_UG = 8;                   // hidden units per TMEM load group
bool g_tcs_dynamic = true;                  // experiments: (var & 32) switches the dynamic schedule off
constexpr int TCS_DEFAULT_UW = 2;           // two threads per stream (measured best)
constexpr int TCS_DEFAULT_VAR = 3;          // staggered tiles + reciprocal shared by two units

#ifdef NTM_TCS_TRACE
// debug build only (NTM_EXTRA_NVCC_FLAGS=-DNTM_TCS_TRACE): per-step clock stamps of CTA 0, [tile][step][start, mid, end]
__device__ long long g_tcs_trace[2 * 256 * 3];
#define TCS_STAMP(slot)                                                                          \
    do {                                                                                         \
        if (blockIdx.x == 0 && lane == 0 && wq == 0 && uh == 0 && t >= 64 && t < 64 + 256)       \
            g_tcs_trace[(tile * 256 + (int)(t - 64)) * 3 + (slot)] = clock64();                  \
    } while (0)
#else
#define TCS_STAMP(slot) do {} while (0)
#endif

// dynamic schedule of one launch (all null / 0: static)
struct TcsSched {
    unsigned long long* counter;   // next job
    int* flags;                    // [group]: time chunks completed
    int n_groups, n_chunks, chunk_T;
};

template <int FMT>
__device__ __forceinline__ uint32_t pack_op(float lo, float hi)
{
    if (FMT == FMT_BF16) return pack_bf16(lo, hi);
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
template <int FMT>
__device__ __forceinline__ float round_op(float x)
{
    return FMT == FMT_BF16 ? __bfloat162float(__float2bfloat16_rn(x)) : __half2float(__float2half_rn(x));
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3])
                 : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t r0, uint32_t r1)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(r0), "r"(r1) : "memory");
}

// Gate math of one tile: one thread = one stream (UW == 1) or one NU-unit slice of it (UW == 2, slice UH).
template <int FMT, int TILES, int UW, int VAR, int UH>
__device__ __forceinline__ void tcs_epilogue(const GruArgs& a, const TcsConsts& kc, int tile, int wq, int lane, uint32_t tmem,
                                             uint64_t* bars, float* ypart, long long group, long long t0, int nsteps,
                                             long long base)
{
    constexpr bool STAGGER = (VAR & 1) != 0 && TILES == 2, SHARE4 = (VAR & 2) != 0;
    constexpr int NPOLY = (VAR >> 2) & 3;
    constexpr int NU = 64 / UW;                  // hidden units per thread
    constexpr int NG = NU / TS_UG;               // TMEM load groups per thread and step
    constexpr bool PREFETCH = UW == 1;           // double-buffered accumulator loads (registers allow it only for UW == 1)
    constexpr int uh = UH, u0 = UH * NU;
    const int s = wq * 32 + lane;            // stream inside the tile == TMEM lane
    const long long b0 = (group * TILES + tile) * TS_M;
    const int ns = (int)((a.B - b0) < (long long)TS_M ? (a.B - b0) : (long long)TS_M);   // may be <= 0
    // both tiles of the group are live (the stagger protocol needs a partner)
    const bool stagger = STAGGER && (group * TILES + 1) * TS_M < a.B;
    if (ns > 0) {
        const bool valid = s < ns;
        const long long row = b0 + (valid ? s : 0);
        const float* __restrict__ xp = a.x + row * a.ldx + t0;      // this job's time window
        float* __restrict__ yp = a.y + row * a.ldy + t0;
        const long long Trem = a.T - t0;                             // samples of x readable from xp
        const uint32_t t_acc = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(tile * TS_TILE_COLS);
        const uint32_t t_op = t_acc + TS_A_OFF;
        float* const yslot = ypart + tile * 2 * TS_M + s;

        // ---- initial state: fp32 in registers, rounded copy + the first input sample into the A operand ----
        float h[NU];
#pragma unroll
        for (int j = 0; j < NU; ++j) {
            // a later time chunk continues from the state its predecessor (possibly another SM) left in h_out: L2 reads
            if (t0 > 0) h[j] = valid ? __ldcg(a.h_out + row * 64 + u0 + j) : 0.0f;
            else h[j] = (valid && a.h_in) ? a.h_in[row * 64 + u0 + j] : 0.0f;
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            uint32_t w[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) w[p] = pack_op<FMT>(h[8 * g + 2 * p], h[8 * g + 2 * p + 1]);
            tmem_st4(t_op + u0 / 2 + 4 * g, w);
        }
        float x0 = valid ? xp[0] : 0.0f;
        float x1 = (valid && Trem > 1) ? xp[1] : 0.0f;
        float xprev = 0.0f, yprev = 0.0f;    // UW == 2: this thread's head partial / input sample of the previous step
        if (uh == 0) {
            // K augmentation, columns 32..39 = k 64..79: [x_hi, x_lo | x_hi, 1 | 1, 0 | 0 ...]
            const float xh = round_op<FMT>(x0);
            const uint32_t w[4] = {pack_op<FMT>(1.0f, 0.0f), 0u, 0u, 0u};
            tmem_st2(t_op + 32, pack_op<FMT>(xh, x0 - xh), pack_op<FMT>(xh, 1.0f));
            tmem_st4(t_op + 34, w);
            tmem_st2(t_op + 38, 0u, 0u);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[tile]);

        for (int t = 0; t < nsteps; ++t) {
            const float x2 = (valid && t + 2 < Trem) ? __ldg(xp + t + 2) : 0.0f;
            const long long gt = base + t;       // steps this CTA has run so far: the mbarrier phase counter
            if (stagger && (tile == 1 || gt > 0)) mbar_wait_sleep(&bars[4 + (tile ^ 1)], (uint32_t)((gt - (tile == 0)) & 1));
            mbar_wait_sleep(&bars[2 + tile], (uint32_t)(gt & 1));
            tc_fence_after();
            TCS_STAMP(0);
            if (UW == 2 && uh == 0 && t > 0) {
                // the partner's head partial of the previous step (published before its h_ready arrive, which
                // happens-before the commit this thread just observed)
                float v = yprev + yslot[((gt - 1) & 1) * TS_M];
                if (a.skip) v += xprev;
                if (valid) yp[t - 1] = v;
            }

            float ys[4] = {uh == 0 ? kc.bo : 0.0f, 0.0f, 0.0f, 0.0f};
            uint32_t acc[PREFETCH ? 2 : 1][3][TS_UG];
            tmem_ld8(t_acc + u0, acc[0][0]);
            tmem_ld8(t_acc + 64 + u0, acc[0][1]);
            tmem_ld8(t_acc + 128 + u0, acc[0][2]);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                constexpr int NB = PREFETCH ? 2 : 1;
                tmem_ld_wait();
                if (PREFETCH && g + 1 < NG) {
                    tmem_ld8(t_acc + u0 + (g + 1) * TS_UG, acc[(g + 1) % NB][0]);
                    tmem_ld8(t_acc + 64 + u0 + (g + 1) * TS_UG, acc[(g + 1) % NB][1]);
                    tmem_ld8(t_acc + 128 + u0 + (g + 1) * TS_UG, acc[(g + 1) % NB][2]);
                }
                float pre[3][TS_UG];
#pragma unroll
                for (int i = 0; i < TS_UG; ++i) {
                    pre[0][i] = __uint_as_float(acc[g % NB][0][i]);
                    pre[1][i] = __uint_as_float(acc[g % NB][1][i]);
                    pre[2][i] = __uint_as_float(acc[g % NB][2][i]);
                }
                if (!PREFETCH && g + 1 < NG) {       // the registers are free again: next group's loads fly during the math
                    tmem_ld8(t_acc + u0 + (g + 1) * TS_UG, acc[0][0]);
                    tmem_ld8(t_acc + 64 + u0 + (g + 1) * TS_UG, acc[0][1]);
                    tmem_ld8(t_acc + 128 + u0 + (g + 1) * TS_UG, acc[0][2]);
                }
                uint32_t w[TS_UG / 2];
#pragma unroll
                for (int p = 0; p < TS_UG / 2; ++p) {
                    const int jl = g * TS_UG + 2 * p, j = u0 + jl;       // local / global unit index (j static per uh)
                    float hn0, hn1;
                    // accumulators hold the complete scaled pre-activations of r, z and W_hn h + b_hn
                    const float g0 = fmaf(kc.cn_w[j], x0, kc.cn_b[j]);
                    const float g1 = fmaf(kc.cn_w[j + 1], x0, kc.cn_b[j + 1]);
                    gates_unit_pair<SHARE4, NPOLY>(pre[0][2 * p], pre[1][2 * p], pre[2][2 * p], g0, pre[0][2 * p + 1],
                                                   pre[1][2 * p + 1], pre[2][2 * p + 1], g1, h[jl], h[jl + 1], hn0, hn1);
                    h[jl] = hn0;
                    h[jl + 1] = hn1;
                    const float w0 = kc.wo[j];
                    const float w1 = kc.wo[j + 1];
                    ys[p & 3] = fmaf(w0, hn0, ys[p & 3]);
                    ys[(p + 2) & 3] = fmaf(w1, hn1, ys[(p + 2) & 3]);
                    w[p] = pack_op<FMT>(hn0, hn1);
                }
                tmem_st4(t_op + u0 / 2 + g * (TS_UG / 2), w);
                if (g == NG / 2 - 1) TCS_STAMP(1);
                if (stagger && g == NG / 2 - 1) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars[4 + tile]);
                }
            }
            if (uh == 0) {
                const float xh = round_op<FMT>(x1);      // input sample of the NEXT step
                tmem_st2(t_op + 32, pack_op<FMT>(xh, x1 - xh), pack_op<FMT>(xh, 1.0f));
            }
            float v = (ys[0] + ys[1]) + (ys[2] + ys[3]);
            if (UW == 2 && uh == 1) yslot[(gt & 1) * TS_M] = v;
            // release the next MMA batch of this tile (the job's last step has no successor here)
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0 && t + 1 < nsteps) mbar_arrive(&bars[tile]);
            TCS_STAMP(2);

            if (UW == 1) {
                if (a.skip) v += x0;
                if (valid) yp[t] = v;
            } else {
                yprev = v;                   // uh == 0: completed by the partner's partial at the next step
                xprev = x0;
            }
            x0 = x1;
            x1 = x2;
        }
        if (UW == 2) {
            asm volatile("bar.sync %0, %1;" ::"r"(1 + tile), "r"(64 * 4) : "memory");     // last partials visible
            if (uh == 0 && valid) {
                float v = yprev + yslot[((base + nsteps - 1) & 1) * TS_M];
                if (a.skip) v += xprev;
                yp[nsteps - 1] = v;
            }
        }

        if (valid) {
#pragma unroll
            for (int j = 0; j < NU; ++j) a.h_out[row * 64 + u0 + j] = h[j];
        }
    }
}

// UW: epilogue warps per TMEM lane quarter (1: a thread owns all 64 units of its stream; 2: two threads own 32 units each).
// VAR bit 0: staggered tiles (a tile may start a step's gate math only after the other tile passed the middle of its
//            own -- keeps the two tiles of a CTA in anti-phase); bit 1: reciprocal shared by two
//            units (4.0 MUFU per unit-step); bits 2-3: ex2 evaluations per unit pair moved to the FMA pipe.
template <int FMT, int TILES, int UW, int VAR>
__global__ void __launch_bounds__(32 * (4 * UW + 1) * TILES, 1) gru_tcs_kernel(const GruArgs a, const __grid_constant__ TcsConsts kc, const TcsSched sc)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* const bop = smem;
    uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + TS_OFF_BAR);    // [tile]: h_ready, [2 + tile]: acc_full, [4 + tile]: mid
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + TS_OFF_BAR + 48);
    float* const ypart = reinterpret_cast<float*>(smem + TS_OFF_YP);          // [tile][2][128] head partials (UW == 2)

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    constexpr int EPI_WARPS = 4 * UW * TILES;
    constexpr uint32_t TMEM_COLS = TILES * TS_TILE_COLS;

    // ---- one-time setup: TMEM, barriers, weights -> shared-memory B operand -----------------------------------
    if (warp == EPI_WARPS) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    if (tid == 0) {
        for (int t = 0; t < TILES; ++t) {
            mbar_init(&bars[t], 4 * UW);         // one arrive per epilogue warp of the tile
            mbar_init(&bars[2 + t], 1);          // tcgen05.commit
            mbar_init(&bars[4 + t], 4 * UW);     // middle of the tile's gate math
        }
        fence_mbar_init();
    }
    {
        // image rows: [gate tile][row m (unit = m % 64)][80 k] 16-bit, k contiguous (pack_tc_images, gru_tc.cu); the B
        // operand row n = gate * 64 + unit is K-major: 16-byte k chunks at LBO, 8-row groups at SBO
        const uint4* img = reinterpret_cast<const uint4*>(a.blob + BlobLayout::tc_image(FMT));
        for (int idx = tid; idx < TS_N * 2 * TS_NK; idx += blockDim.x) {
            const int n = idx / (2 * TS_NK), c = idx % (2 * TS_NK);
            const uint4 v = img[((n >> 6) * 128 + (n & 63)) * (2 * TS_NK) + c];
            *reinterpret_cast<uint4*>(bop + c * TS_LBO + (n >> 3) * TS_SBO + (n & 7) * 16) = v;
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // ---- job loop.  A job = one group of TILES consecutive 128-stream tiles x one time chunk.  Static schedule
    // (sc.counter == nullptr): the CTA runs group blockIdx.x over all T steps.  Dynamic schedule (more groups than SMs):
    // a persistent grid pulls (group, chunk) jobs from a global counter, chunk-major, so that whole SMs never idle in a
    // partial last wave; a group's state travels between chunks (and SMs) through h_out, guarded by a per-group
    // progress flag (release/acquire at GPU scope).
    int* const job_slot = reinterpret_cast<int*>(smem + TS_OFF_BAR + 52);
    long long base = 0;                          // steps this CTA has run so far
    for (long long it = 0;; ++it) {
        long long group, t0;
        int nsteps;
        if (sc.counter == nullptr) {
            if (it > 0) break;
            group = blockIdx.x; t0 = 0; nsteps = (int)a.T;
        } else {
            if (tid == 0) {
                const unsigned long long job = atomicAdd(sc.counter, 1ull);
                int chunk = -1;
                if (job < (unsigned long long)sc.n_groups * sc.n_chunks) {
                    chunk = (int)(job / sc.n_groups);
                    const long long grp = (long long)(job % sc.n_groups);
                    while (*reinterpret_cast<volatile int*>(sc.flags + grp) < chunk) __nanosleep(200);
                    __threadfence();
                    job_slot[1] = (int)grp;
                }
                job_slot[0] = chunk;
            }
            __syncthreads();
            const int chunk = job_slot[0];
            if (chunk < 0) break;
            group = job_slot[1];
            t0 = (long long)chunk * sc.chunk_T;
            nsteps = (int)((a.T - t0) < (long long)sc.chunk_T ? (a.T - t0) : (long long)sc.chunk_T);
        }

        if (warp >= EPI_WARPS) {
            // ================================ MMA issue warp of one tile ======================================
            const int tile = warp - EPI_WARPS;
            const long long b0 = (group * TILES + tile) * TS_M;
            if (b0 < a.B && elect_one()) {
                constexpr uint32_t idesc = instr_desc(FMT, TS_M, TS_N);
                const uint32_t d_base = tmem + (uint32_t)(tile * TS_TILE_COLS);
                const uint32_t a_base = d_base + TS_A_OFF;
                const uint32_t b_base = smem_u32(bop);
                for (int t = 0; t < nsteps; ++t) {
                    mbar_wait_sleep(&bars[tile], (uint32_t)((base + t) & 1));
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < TS_NK; ++ks)
                        mma_ts<FMT>(d_base, a_base + ks * 8, smem_desc(b_base + ks * 2 * TS_LBO, TS_LBO, TS_SBO), idesc, ks > 0);
                    mma_commit(&bars[2 + tile]);
                }
            }
        } else {
            // ================================ epilogue warps ======================================================
            const int tile = warp / (4 * UW);
            const int wq = warp & 3;                 // TMEM lane quarter (== warp id % 4)
            if (UW == 1 || ((warp >> 2) & 1) == 0)
                tcs_epilogue<FMT, TILES, UW, VAR, 0>(a, kc, tile, wq, lane, tmem, bars, ypart, group, t0, nsteps, base);
            else
                tcs_epilogue<FMT, TILES, UW, VAR, UW - 1>(a, kc, tile, wq, lane, tmem, bars, ypart, group, t0, nsteps, base);
        }
        base += nsteps;

        // end of the job: every role is done with TMEM and the barriers' phases agree again
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (sc.counter != nullptr && tid == 0) {
            __threadfence();                                       // the group's h_out is visible GPU-wide ...
            atomicExch(sc.flags + group, (int)(t0 / sc.chunk_T) + 1);   // ... before its next chunk may start
        }
    }

    // ---- teardown: every MMA has completed (the epilogue consumed the last accumulator) ------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS) tmem_dealloc(tmem, TMEM_COLS);
}

template <int FMT, int TILES, int UW, int VAR>
cudaError_t launch_tcs_one(const GruArgs& a, const TcsConsts& kc, int sm_count, cudaStream_t st)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && !configured[dev]) {
        e = cudaFuncSetAttribute(gru_tcs_kernel<FMT, TILES, UW, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)TS_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const long long per_group = (long long)TS_M * TILES;
    const long long groups = (a.B + per_group - 1) / per_group;
    TcsSched sc{};
    long long grid = groups;
    // More groups than SMs and a partial last wave: pull (group, time chunk) jobs dynamically.  Chunk length: a job
    // costs ~7 steps of overhead (state round trip through L2, pipeline fill/drain; measured) and the tail idles half a
    // job per SM on average, so ct ~ sqrt(2 * 7 * T * groups / SMs), as a power of two in 64 .. 1024.
    if (groups > sm_count && groups % sm_count != 0 && groups < (1ll << 30) && g_tcs_dynamic) {
        const double ct = sqrt(14.0 * (double)a.T * (double)groups / (double)sm_count);
        long long p2 = 64;
        while (p2 * 1.41 <= ct && p2 < 1024) p2 *= 2;
        const long long chunks = (a.T + p2 - 1) / p2;
        if (chunks > 1 && chunks < (1ll << 30)) {
            void* scratch = nullptr;
            const size_t bytes = 16 + sizeof(int) * (size_t)groups;
            e = cudaMallocAsync(&scratch, bytes, st);               // stream-ordered: no device synchronisation
            if (e != cudaSuccess) return e;
            e = cudaMemsetAsync(scratch, 0, bytes, st);
            if (e != cudaSuccess) return e;
            sc.counter = static_cast<unsigned long long*>(scratch);
            sc.flags = reinterpret_cast<int*>(static_cast<char*>(scratch) + 16);
            sc.n_groups = (int)groups; sc.n_chunks = (int)chunks; sc.chunk_T = (int)p2;
            grid = sm_count;
        }
    }
    gru_tcs_kernel<FMT, TILES, UW, VAR><<<(unsigned)grid, 32 * (4 * UW + 1) * TILES, TS_SMEM_BYTES, st>>>(a, kc, sc);
    ++g_launches;
    e = cudaGetLastError();
    if (sc.counter) {
        const cudaError_t e2 = cudaFreeAsync(sc.counter, st);
        if (e == cudaSuccess) e = e2;
    }
    return e;
}

}  // namespace

#ifdef NTM_TCS_TRACE
extern "C" __attribute__((visibility("default"))) int ntm_debug_tcs_trace(long long* dst)
{
    return (int)cudaMemcpyFromSymbol(dst, g_tcs_trace, sizeof(long long) * 2 * 256 * 3);
}
#endif

// host: the warp-uniform per-unit constants of the stream-major kernel, from the fp32 part of the blob
void fill_tcs_consts(const float* blob_host, TcsConsts* kc)
{
    constexpr float L = 1.4426950408889634f;
    for (int j = 0; j < 64; ++j) {
        kc->cn_w[j] = 2.0f * L * blob_host[BlobLayout::W_IH + 128 + j];
        kc->cn_b[j] = 2.0f * L * blob_host[BlobLayout::B_IH + 128 + j];
        kc->wo[j] = blob_host[BlobLayout::W_OUT + j];
    }
    kc->bo = blob_host[BlobLayout::B_OUT];
}

namespace {

template <int FMT, int TILES>
cudaError_t launch_tcs_var(const GruArgs& a, const TcsConsts& kc, int var, int sm_count, cudaStream_t st)
{
    g_tcs_dynamic = var < 0 || (var & 32) == 0;
    if (var >= 0) var &= 31;
    switch (var) {      // experiments: (var & 15) = VAR bits, (var & 16) = two threads per stream, (var & 32) = static
        case 0: return launch_tcs_one<FMT, TILES, 1, 0>(a, kc, sm_count, st);
        case 1: return launch_tcs_one<FMT, TILES, 1, 1>(a, kc, sm_count, st);
        case 2: return launch_tcs_one<FMT, TILES, 1, 2>(a, kc, sm_count, st);
        case 3: return launch_tcs_one<FMT, TILES, 1, 3>(a, kc, sm_count, st);
        case 16: return launch_tcs_one<FMT, TILES, 2, 0>(a, kc, sm_count, st);
        case 17: return launch_tcs_one<FMT, TILES, 2, 1>(a, kc, sm_count, st);
        case 18: return launch_tcs_one<FMT, TILES, 2, 2>(a, kc, sm_count, st);
        case 19: return launch_tcs_one<FMT, TILES, 2, 3>(a, kc, sm_count, st);
        case 23: return launch_tcs_one<FMT, TILES, 2, 7>(a, kc, sm_count, st);
        case 31: return launch_tcs_one<FMT, TILES, TCS_DEFAULT_UW, TCS_DEFAULT_VAR>(a, kc, sm_count, st);
        default: return launch_tcs_one<FMT, TILES, TCS_DEFAULT_UW, TCS_DEFAULT_VAR>(a, kc, sm_count, st);
    }
}

}  // namespace

// fmt: FMT_F16 / FMT_BF16.  tiles: 128-stream tiles per CTA (1 or 2; 0 = automatic); var: kernel variant (experiments;
// -1 = default).  Plain GRU only (a.d == nullptr).
cudaError_t launch_gru_tcs(const GruArgs& a, const TcsConsts& kc, int fmt, int sm_count, int tiles, int var, cudaStream_t st)
{
    if (a.B <= 0 || a.T <= 0) return cudaSuccess;
    if (a.d != nullptr) return cudaErrorInvalidValue;
    // one tile per SM (time-multiplexed by the job queue) beats two resident tiles until ~190 streams per SM (measured)
    if (tiles <= 0) tiles = a.B >= 190ll * sm_count ? 2 : 1;
    if (fmt == FMT_BF16) return tiles >= 2 ? launch_tcs_var<FMT_BF16, 2>(a, kc, var, sm_count, st) : launch_tcs_var<FMT_BF16, 1>(a, kc, var, sm_count, st);
    return tiles >= 2 ? launch_tcs_var<FMT_F16, 2>(a, kc, var, sm_count, st) : launch_tcs_var<FMT_F16, 1>(a, kc, var, sm_count, st);
}

}  // namespace ntm
