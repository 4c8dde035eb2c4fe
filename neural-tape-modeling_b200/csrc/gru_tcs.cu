// Stream-major tensor-core persistent GRU kernel (tcgen05.mma + TMEM) -- the THROUGHPUT regime of the batched path
// (>= ~110 streams per SM; BASELINE cfg 4: 65 536 streams).
//
// Replaces, for many concurrent streams, the per-timestep loop behind `self.GRU(x, self.hidden)` + `self.output(x)` of
// RNN.forward / DiffDelRNN.forward (code/model.py:81-82, :412-413; gate equations torch rnn.py:1221-1224).
//
// Formulation.  One tile = 128 streams = the M rows of the MMA.  Per timestep the tensor core evaluates
//       G[128 streams x 192] = A[128 x K] . W^T[K x 192]
//   A operand  the (rounded) state of the tile, IN TENSOR MEMORY (".ts" MMA form): TMEM lane = stream, so the thread that
//              owns a stream writes its new state with tcgen05.st straight from registers -- no shared-memory round trip,
//              no proxy fence, no cross-thread exchange anywhere in the step.
//   B operand  W_hh^T, K-major in shared memory, staged once from the image ntm_gru_prepare packed (rows pre-scaled by
//              -log2 e / 2 log2 e).
//   N = 192    MMAs of M=128, N=192 run at the tensor pipe's full rate (measured 106 clk per K=16 f16 MMA,
//              profiles/r01_tc_probe.txt) = 4..10 clk per stream-step, against >= 16 clk per stream-step of MUFU work:
//              the tensor pipe is never the bound here.
//   epilogue   accumulator lane = stream, column = gate row: TWO THREADS OWN ONE STREAM (32 hidden units each) -- their
//              fp32 states live in registers for the whole launch, r/z/n of every unit arrive by tcgen05.ld, per-unit
//              constants are warp-uniform (kernel-parameter constant bank), the output head is an in-thread fp32 dot
//              product, reciprocals are shared between the r and z gates of two units and the n gates of four units
//              (3.75 MUFU per unit-step; 4.0 in the formats without the K augmentation, gates.cuh).
// Operand formats (template FMT):
//   f16 / bf16   K = 80: [H | x_hi x_lo x_hi 1 1 0..]: the K augmentation carries W_i x + b (r, z) and b_hn (n) through the MMA
//   f16x3        STRICT, fp32-grade: K = 192 = [h_hi | h_lo' | h_hi] . [G W_hi | G W_hi / 2^8 | (G W)_lo]^T with
//                h_hi = f16(h), h_lo' = f16((h - h_hi) 2^8), G a power of two that lifts the residual (G W)_lo out of the f16
//                subnormals; the third block re-reads the h_hi columns of TMEM.  Input projection and biases in fp32, the
//                accumulator is scaled back by 1 / G in the same FMA; Newton-refined reciprocals (gates.cuh, rcp_strict).
//   tf32         K = 64 of kind::tf32 (one element per TMEM column), input projection and biases in fp32.
// Roles: 8 epilogue warps per tile (two per TMEM lane quarter) + 1 MMA-issue warp per tile; two tiles per CTA ping-pong
// so one tile's MMA + hand-off latency hides behind the other tile's gate math.  mbarrier hand-off:
//   epilogue: tcgen05.st (new A) -> wait::st -> fence::before_thread_sync -> arrive(h_ready[tile])      (count 8)
//   issuer:   wait(h_ready) -> fence::after_thread_sync -> NMMA x tcgen05.mma -> tcgen05.commit(acc_full[tile])
// x is prefetched two steps ahead by the owning thread, y is written by the owning thread (partial sectors merge in L2).
#include <math.h>
#include <string.h>

#include "gates.cuh"
#include "tc_prims.cuh"

namespace ntm {

using namespace tc;

namespace {

constexpr int TS_M = 128;                  // streams per tile
constexpr int TS_N = 192;                  // gate rows
constexpr int TS_TILE_COLS = 256;          // TMEM columns reserved per tile: 192 accumulator + up to 64 operand columns
constexpr int TS_A_OFF = 192;              // first operand column inside a tile's TMEM block
constexpr uint32_t TS_SBO = 128, TS_LBO = (TS_N / 8) * 128;       // K-major, no swizzle
constexpr int TS_UG = 8;                   // hidden units per TMEM load group
constexpr int TS_UW = 2;                   // threads per stream
constexpr float TS_LO_SCALE = 256.0f;      // strict form: h_lo' = (h - h_hi) * 2^8
#ifndef NTM_TCS_STRICT_OWN_RCP
// strict form: 1 = every gate its own reciprocal, 0 = r/z of a unit and n of two units share one.  Measured (profiles/
// r02_strict_tcs_variants.txt): own reciprocals change NOTHING in the achieved error (cfg-1 pulse train, the one chaotic
// golden signal: 3.3e-5 either way) and cost 16 % (9.18 vs 10.96 Gsamples/s at 37 888 streams) -- sharing stays.
#define NTM_TCS_STRICT_OWN_RCP 0
#endif
constexpr int TCS_DEFAULT_VAR = 15;        // staggered tiles + r/z reciprocal shared by two units + packed fp32 arithmetic + n reciprocal
                                           // shared by four units (the last only where the K augmentation exists: f16, bf16)

template <int FMT>
struct TsFmt {
    static constexpr bool AUG = FMT < 2;                       // the K augmentation carries W_i x + b
    static constexpr bool STRICT = FMT == FMT_F16X3;
    static constexpr int KIND = FMT == FMT_TF32 ? FMT_TF32 : FMT == FMT_BF16 ? FMT_BF16 : FMT_F16;   // MMA kind / descriptor format
    static constexpr int ELT = FMT == FMT_TF32 ? 4 : 2;
    static constexpr int KB = AUG ? 80 : STRICT ? 192 : 64;    // K of the B image
    static constexpr int CHUNKS = KB * ELT / 16;               // 16-byte k chunks per row
    static constexpr int NMMA = CHUNKS / 2;                    // one MMA consumes 32 bytes of K
    static constexpr uint32_t B_BYTES = CHUNKS * TS_LBO;
    static constexpr uint32_t OFF_BAR = B_BYTES;               // 6 mbarriers + the TMEM base slot + the job slot
    static constexpr uint32_t OFF_YP = OFF_BAR + 64;           // [2 tiles][2][128] floats
    static constexpr uint32_t SMEM_BYTES = OFF_YP + 2 * 2 * 128 * 4;
};

#ifdef NTM_TCS_TRACE
// debug build only (NTM_EXTRA_NVCC_FLAGS=-DNTM_TCS_TRACE): per-step clock stamps of CTA 0, [tile][step][start, mid, end]
__device__ long long g_tcs_trace[2 * 256 * 3];
#define TCS_STAMP(slot)                                                                          \
    do {                                                                                         \
        if (blockIdx.x == 0 && lane == 0 && wq == 0 && UH == 0 && t >= 64 && t < 64 + 256)       \
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

// ---- packed fp32 arithmetic (sm_100: add / mul / fma .f32x2 = FADD2 / FMUL2 / FFMA2) ------------------------------------
// The epilogue is co-limited by the MUFU pipe and by ISSUE SLOTS (ncu r01: issue 68 %, XU 78 %, FMA pipe 37 %): every
// operation on the two hidden units of a pair is the same instruction twice, so it is issued once on a register pair.
// New states of the two hidden units (j, j + 1) of one stream from their accumulator values.
//   AUG formats: ar, az are complete scaled pre-activations, an = W_hn h + b_hn.
//   others:      ar, az, an = (gain x) W_h* h; input projection and biases are added here in fp32.
// SHARE4: one reciprocal serves r and z of BOTH units (4.0 MUFU per unit-step; ex2 arguments of r, z clamped to 30 so the
// product of four denominators stays finite: sigmoid saturates at 2^-30); STRICT: Newton-refined reciprocals.
template <int FMT, bool SHARE4, bool PACK2>
__device__ __forceinline__ void tcs_gate_pair(const TcsConsts& kc, int j, float x, float ar0, float ar1, float az0, float az1,
                                              float an0, float an1, float& h0, float& h1, float& ys)
{
    using F = TsFmt<FMT>;
    constexpr float C = SHARE4 ? 30.0f : EX2_CLAMP;
    auto rcp_f = [](float d) { return F::STRICT ? rcp_strict(d) : rcp_approx(d); };
    if (PACK2) {
        const f32x2 one = pk(1.0f, 1.0f), x2 = pk(x, x);
        f32x2 pr = pk(ar0, ar1), pz = pk(az0, az1), ahn = pk(an0, an1);
        if (!F::AUG) {
            const f32x2 gi = pk(kc.gain_inv, kc.gain_inv);
            const f32x2 ir = fma2(pk(kc.cr_w[j], kc.cr_w[j + 1]), x2, pk(kc.cr_b[j], kc.cr_b[j + 1]));
            const f32x2 iz = fma2(pk(kc.cz_w[j], kc.cz_w[j + 1]), x2, pk(kc.cz_b[j], kc.cz_b[j + 1]));
            const f32x2 hb = pk(kc.ch_b[j], kc.ch_b[j + 1]);
            if (F::STRICT) { pr = fma2(pr, gi, ir); pz = fma2(pz, gi, iz); ahn = fma2(ahn, gi, hb); }
            else { pr = add2(pr, ir); pz = add2(pz, iz); ahn = add2(ahn, hb); }
        }
        float r0, r1, z0, z1;
        upk(pr, r0, r1);
        upk(pz, z0, z1);
        const f32x2 dr = add2(pk(ex2_approx(fminf(r0, C)), ex2_approx(fminf(r1, C))), one);
        const f32x2 dz = add2(pk(ex2_approx(fminf(z0, C)), ex2_approx(fminf(z1, C))), one);
        const f32x2 p = mul2(dr, dz);
        float p0, p1;
        upk(p, p0, p1);
        f32x2 r, z;
        if (F::STRICT && NTM_TCS_STRICT_OWN_RCP) {       // every gate its own Newton-refined reciprocal (as gates_strict)
            float a0, a1, b0, b1;
            upk(dr, a0, a1);
            upk(dz, b0, b1);
            r = pk(rcp_f(a0), rcp_f(a1));
            z = pk(rcp_f(b0), rcp_f(b1));
        } else {
            f32x2 inv;
            if (SHARE4) {
                const float q = rcp_f(p0 * p1);
                inv = mul2(pk(p1, p0), pk(q, q));
            } else {
                inv = pk(rcp_f(p0), rcp_f(p1));
            }
            r = mul2(dz, inv);
            z = mul2(dr, inv);
        }
        const f32x2 gin = fma2(pk(kc.cn_w[j], kc.cn_w[j + 1]), x2, pk(kc.cn_b[j], kc.cn_b[j + 1]));
        float a0, a1;
        upk(fma2(r, ahn, gin), a0, a1);
        const f32x2 dn = add2(pk(ex2_approx(fminf(a0, EX2_CLAMP)), ex2_approx(fminf(a1, EX2_CLAMP))), one);
        float d0, d1;
        upk(dn, d0, d1);
        f32x2 n;
        if (F::STRICT && NTM_TCS_STRICT_OWN_RCP) {
            n = fma2(pk(rcp_f(d0), rcp_f(d1)), pk(-2.0f, -2.0f), one);
        } else {
            const float qn = rcp_f(d0 * d1);
            n = fma2(mul2(pk(d1, d0), pk(qn, qn)), pk(-2.0f, -2.0f), one);          // 1 - 2 / dn
        }
        const f32x2 h = pk(h0, h1);
        const f32x2 hn = fma2(z, fma2(n, pk(-1.0f, -1.0f), h), n);                          // n + z (h - n)
        upk(hn, h0, h1);
        ys = fmaf(kc.wo[j], h0, ys);
        ys = fmaf(kc.wo[j + 1], h1, ys);
    } else {
        if (!F::AUG) {
            const float gi = F::STRICT ? kc.gain_inv : 1.0f;
            ar0 = fmaf(ar0, gi, fmaf(kc.cr_w[j], x, kc.cr_b[j]));
            ar1 = fmaf(ar1, gi, fmaf(kc.cr_w[j + 1], x, kc.cr_b[j + 1]));
            az0 = fmaf(az0, gi, fmaf(kc.cz_w[j], x, kc.cz_b[j]));
            az1 = fmaf(az1, gi, fmaf(kc.cz_w[j + 1], x, kc.cz_b[j + 1]));
            an0 = fmaf(an0, gi, kc.ch_b[j]);
            an1 = fmaf(an1, gi, kc.ch_b[j + 1]);
        }
        const float dr0 = 1.0f + ex2_approx(fminf(ar0, C)), dr1 = 1.0f + ex2_approx(fminf(ar1, C));
        const float dz0 = 1.0f + ex2_approx(fminf(az0, C)), dz1 = 1.0f + ex2_approx(fminf(az1, C));
        float r0, z0, r1, z1;
        const float p0 = dr0 * dz0, p1 = dr1 * dz1;
        if (SHARE4) {
            const float q = rcp_f(p0 * p1);
            const float i0 = p1 * q, i1 = p0 * q;
            r0 = dz0 * i0; z0 = dr0 * i0; r1 = dz1 * i1; z1 = dr1 * i1;
        } else {
            const float i0 = rcp_f(p0), i1 = rcp_f(p1);
            r0 = dz0 * i0; z0 = dr0 * i0; r1 = dz1 * i1; z1 = dr1 * i1;
        }
        const float g0 = fmaf(kc.cn_w[j], x, kc.cn_b[j]), g1 = fmaf(kc.cn_w[j + 1], x, kc.cn_b[j + 1]);
        const float dn0 = 1.0f + ex2_approx(fminf(fmaf(r0, an0, g0), EX2_CLAMP));
        const float dn1 = 1.0f + ex2_approx(fminf(fmaf(r1, an1, g1), EX2_CLAMP));
        const float qn = rcp_f(dn0 * dn1);
        const float n0 = fmaf(-2.0f, dn1 * qn, 1.0f), n1 = fmaf(-2.0f, dn0 * qn, 1.0f);
        h0 = fmaf(z0, h0 - n0, n0);
        h1 = fmaf(z1, h1 - n1, n1);
        ys = fmaf(kc.wo[j], h0, ys);
        ys = fmaf(kc.wo[j + 1], h1, ys);
    }
}

// FOUR hidden units (j .. j+3) of one stream, AUG formats, packed arithmetic: r and z of each pair share a reciprocal (as in
// tcs_gate_pair, SHARE4), and the n gates of all four units share ONE (3.75 MUFU per unit-step instead of 4.0; ex2 arguments of n
// clamped to 30 so that the product of four denominators stays below 2^121: tanh saturates at 1 - 2^-29).  Measured at 37 888
// streams (profiles/r02_tcs_quad.txt): 19.44 vs 20.08 clk per stream-step per SM; taking the second pair's z-gate 2^x from the FMA
// pipe on top (3.25 MUFU per unit-step, 12 more instructions per quad) gave 19.32 -- the MUFU pipe is no longer the bound, not kept.
__device__ __forceinline__ void tcs_gate_quad(const TcsConsts& kc, int j, float x, const float (&ar)[4], const float (&az)[4],
                                              const float (&an)[4], float* h, float& ys0, float& ys1)
{
    constexpr float C = 30.0f;
    const f32x2 one = pk(1.0f, 1.0f), x2 = pk(x, x);
    f32x2 z[2], dn[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int jj = j + 2 * q;
        const f32x2 dr = add2(pk(ex2_approx(fminf(ar[2 * q], C)), ex2_approx(fminf(ar[2 * q + 1], C))), one);
        const f32x2 dz = add2(pk(ex2_approx(fminf(az[2 * q], C)), ex2_approx(fminf(az[2 * q + 1], C))), one);
        const f32x2 p = mul2(dr, dz);
        float p0, p1;
        upk(p, p0, p1);
        const float qi = rcp_approx(p0 * p1);
        const f32x2 inv = mul2(pk(p1, p0), pk(qi, qi));
        const f32x2 r = mul2(dz, inv);
        z[q] = mul2(dr, inv);
        const f32x2 gin = fma2(pk(kc.cn_w[jj], kc.cn_w[jj + 1]), x2, pk(kc.cn_b[jj], kc.cn_b[jj + 1]));
        float a0, a1;
        upk(fma2(r, pk(an[2 * q], an[2 * q + 1]), gin), a0, a1);
        dn[q] = add2(pk(ex2_approx(fminf(a0, C)), ex2_approx(fminf(a1, C))), one);
    }
    float d0, d1, d2, d3;
    upk(dn[0], d0, d1);
    upk(dn[1], d2, d3);
    const float pa = d0 * d1, pb = d2 * d3;
    const float qn = rcp_approx(pa * pb);
    const float ia = pb * qn, ib = pa * qn;                      // 1 / (d0 d1), 1 / (d2 d3)
    const f32x2 m2 = pk(-2.0f, -2.0f), m1 = pk(-1.0f, -1.0f);
    const f32x2 n0 = fma2(mul2(pk(d1, d0), pk(ia, ia)), m2, one);          // 1 - 2 / dn
    const f32x2 n1 = fma2(mul2(pk(d3, d2), pk(ib, ib)), m2, one);
    const f32x2 h0 = fma2(z[0], fma2(n0, m1, pk(h[0], h[1])), n0);         // n + z (h - n)
    const f32x2 h1 = fma2(z[1], fma2(n1, m1, pk(h[2], h[3])), n1);
    upk(h0, h[0], h[1]);
    upk(h1, h[2], h[3]);
    ys0 = fmaf(kc.wo[j], h[0], ys0);
    ys0 = fmaf(kc.wo[j + 1], h[1], ys0);
    ys1 = fmaf(kc.wo[j + 2], h[2], ys1);
    ys1 = fmaf(kc.wo[j + 3], h[3], ys1);
}

// Publish the (rounded) states of TS_UG consecutive units, first unit `jl` of this thread's slice, into the tile's A operand.
template <int FMT>
__device__ __forceinline__ void tcs_store_state(uint32_t t_op, int u0, int jl, const float* h)
{
    if (FMT == FMT_TF32) {                       // one tf32 element per column
        uint32_t w[TS_UG];
#pragma unroll
        for (int i = 0; i < TS_UG; ++i) w[i] = to_tf32(h[jl + i]);
        tmem_st8(t_op + u0 + jl, w);
    } else if (FMT == FMT_F16X3) {               // hi pairs at columns 0..31, scaled residual pairs at 32..63
        uint32_t hi[TS_UG / 2], lo[TS_UG / 2];
#pragma unroll
        for (int p = 0; p < TS_UG / 2; ++p) {
            const __half2 hh = __floats2half2_rn(h[jl + 2 * p], h[jl + 2 * p + 1]);
            const float2 hf = __half22float2(hh);
            const __half2 hl = __floats2half2_rn((h[jl + 2 * p] - hf.x) * TS_LO_SCALE, (h[jl + 2 * p + 1] - hf.y) * TS_LO_SCALE);
            hi[p] = *reinterpret_cast<const uint32_t*>(&hh);
            lo[p] = *reinterpret_cast<const uint32_t*>(&hl);
        }
        tmem_st4(t_op + (u0 + jl) / 2, hi);
        tmem_st4(t_op + 32 + (u0 + jl) / 2, lo);
    } else {
        uint32_t w[TS_UG / 2];
#pragma unroll
        for (int p = 0; p < TS_UG / 2; ++p) w[p] = pack_op<FMT>(h[jl + 2 * p], h[jl + 2 * p + 1]);
        tmem_st4(t_op + (u0 + jl) / 2, w);
    }
}

// Gate math of one tile: one thread = one 32-unit slice (UH = 0, 1) of one stream.
template <int FMT, int TILES, int VAR, int UH>
__device__ __forceinline__ void tcs_epilogue(const GruArgs& a, const TcsConsts& kc, int tile, int wq, int lane, uint32_t tmem,
                                             uint64_t* bars, float* ypart, long long group, long long t0, int nsteps,
                                             long long base)
{
    using F = TsFmt<FMT>;
    // (the strict form shares a reciprocal only between the r and z gate of ONE unit: a product of four denominators
    // carries three more roundings into every gate value)
    constexpr bool STAGGER = (VAR & 1) != 0 && TILES == 2, SHARE4 = (VAR & 2) != 0 && !F::STRICT, PACK2 = (VAR & 4) != 0;
    // bit 3: the n gates of four units share one reciprocal (3.75 MUFU per unit-step; formats with the K augmentation only)
    constexpr bool QUAD = (VAR & 8) != 0 && F::AUG && SHARE4 && PACK2;
    constexpr int NU = 64 / TS_UW;               // hidden units per thread
    constexpr int NG = NU / TS_UG;               // TMEM load groups per thread and step
    constexpr int u0 = UH * NU;
    const int s = wq * 32 + lane;                // stream inside the tile == TMEM lane
    const long long b0 = (group * TILES + tile) * TS_M;
    const int ns = (int)((a.B - b0) < (long long)TS_M ? (a.B - b0) : (long long)TS_M);   // may be <= 0
    // both tiles of the group are live (the stagger protocol needs a partner)
    const bool stagger = STAGGER && (group * TILES + 1) * TS_M < a.B;
    if (ns <= 0) return;
    const bool valid = s < ns;
    const long long row = b0 + (valid ? s : 0);
    const float* __restrict__ xp = a.x + row * a.ldx + t0;      // this job's time window
    float* __restrict__ yp = a.y + row * a.ldy + t0;
    const long long Trem = a.T - t0;                             // samples of x readable from xp
    const uint32_t t_acc = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(tile * TS_TILE_COLS);
    const uint32_t t_op = t_acc + TS_A_OFF;
    float* const yslot = ypart + tile * 2 * TS_M + s;

    // ---- initial state: fp32 in registers, rounded copy (+ the first input sample) into the A operand ----
    float h[NU];
#pragma unroll
    for (int j = 0; j < NU; ++j) {
        // a later time chunk continues from the state its predecessor (possibly another SM) left in h_out: L2 reads
        if (t0 > 0) h[j] = valid ? __ldcg(a.h_out + row * 64 + u0 + j) : 0.0f;
        else h[j] = (valid && a.h_in) ? a.h_in[row * 64 + u0 + j] : 0.0f;
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) tcs_store_state<FMT>(t_op, u0, g * TS_UG, h);
    float x0 = valid ? xp[0] : 0.0f;
    float x1 = (valid && Trem > 1) ? xp[1] : 0.0f;
    float xprev = 0.0f, yprev = 0.0f;    // this thread's head partial / input sample of the previous step
    if (F::AUG && UH == 0) {
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
        if (UH == 0 && t > 0) {
            // the partner's head partial of the previous step (published before its h_ready arrive, which
            // happens-before the commit this thread just observed)
            float v = yprev + yslot[((gt - 1) & 1) * TS_M];
            if (a.skip) v += xprev;
            if (valid) yp[t - 1] = v;
        }

        float ys[2] = {UH == 0 ? kc.bo : 0.0f, 0.0f};
        uint32_t acc[3][TS_UG];
        tmem_ld8(t_acc + u0, acc[0]);
        tmem_ld8(t_acc + 64 + u0, acc[1]);
        tmem_ld8(t_acc + 128 + u0, acc[2]);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            tmem_ld_wait();
            float pre[3][TS_UG];
#pragma unroll
            for (int i = 0; i < TS_UG; ++i) {
                pre[0][i] = __uint_as_float(acc[0][i]);
                pre[1][i] = __uint_as_float(acc[1][i]);
                pre[2][i] = __uint_as_float(acc[2][i]);
            }
            if (g + 1 < NG) {       // the registers are free again: next group's loads fly during the math
                tmem_ld8(t_acc + u0 + (g + 1) * TS_UG, acc[0]);
                tmem_ld8(t_acc + 64 + u0 + (g + 1) * TS_UG, acc[1]);
                tmem_ld8(t_acc + 128 + u0 + (g + 1) * TS_UG, acc[2]);
            }
            if (QUAD) {
#pragma unroll
                for (int q = 0; q < TS_UG / 4; ++q) {
                    const int jl = g * TS_UG + 4 * q;
                    const float qr[4] = {pre[0][4 * q], pre[0][4 * q + 1], pre[0][4 * q + 2], pre[0][4 * q + 3]};
                    const float qz[4] = {pre[1][4 * q], pre[1][4 * q + 1], pre[1][4 * q + 2], pre[1][4 * q + 3]};
                    const float qn[4] = {pre[2][4 * q], pre[2][4 * q + 1], pre[2][4 * q + 2], pre[2][4 * q + 3]};
                    tcs_gate_quad(kc, u0 + jl, x0, qr, qz, qn, &h[jl], ys[0], ys[1]);
                }
            } else {
#pragma unroll
            for (int p = 0; p < TS_UG / 2; ++p) {
                const int jl = g * TS_UG + 2 * p;                      // local unit index; global = u0 + jl (static per UH)
                tcs_gate_pair<FMT, SHARE4, PACK2>(kc, u0 + jl, x0, pre[0][2 * p], pre[0][2 * p + 1], pre[1][2 * p],
                                                  pre[1][2 * p + 1], pre[2][2 * p], pre[2][2 * p + 1], h[jl], h[jl + 1], ys[p & 1]);
            }
            }
            tcs_store_state<FMT>(t_op, u0, g * TS_UG, h);
            if (g == NG / 2 - 1) TCS_STAMP(1);
            if (stagger && g == NG / 2 - 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[4 + tile]);
            }
        }
        if (F::AUG && UH == 0) {
            const float xh = round_op<FMT>(x1);      // input sample of the NEXT step
            tmem_st2(t_op + 32, pack_op<FMT>(xh, x1 - xh), pack_op<FMT>(xh, 1.0f));
        }
        const float v = ys[0] + ys[1];
        if (UH == 1) yslot[(gt & 1) * TS_M] = v;
        // release the next MMA batch of this tile (the job's last step has no successor here)
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0 && t + 1 < nsteps) mbar_arrive(&bars[tile]);
        TCS_STAMP(2);

        yprev = v;                   // UH == 0: completed by the partner's partial at the next step
        xprev = x0;
        x0 = x1;
        x1 = x2;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + tile), "r"(64 * 4) : "memory");     // last partials visible
    if (UH == 0 && valid) {
        float v = yprev + yslot[((base + nsteps - 1) & 1) * TS_M];
        if (a.skip) v += xprev;
        yp[nsteps - 1] = v;
    }
    if (valid) {
#pragma unroll
        for (int j = 0; j < NU; ++j) a.h_out[row * 64 + u0 + j] = h[j];
    }
}

// VAR bit 0: staggered tiles (a tile may start a step's gate math only after the other tile passed the middle of its
//            own -- keeps the two tiles of a CTA in anti-phase); bit 1: reciprocal shared by two
//            units (4.0 MUFU per unit-step); bit 2: packed fp32 arithmetic (FFMA2 / FMUL2 / FADD2).
template <int FMT, int TILES, int VAR>
__global__ void __launch_bounds__(32 * (4 * TS_UW + 1) * TILES, 1) gru_tcs_kernel(const GruArgs a, const __grid_constant__ TcsConsts kc, const TcsSched sc)
{
    using F = TsFmt<FMT>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* const bop = smem;
    uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + F::OFF_BAR);    // [tile]: h_ready, [2 + tile]: acc_full, [4 + tile]: mid
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + F::OFF_BAR + 48);
    float* const ypart = reinterpret_cast<float*>(smem + F::OFF_YP);          // [tile][2][128] head partials

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    constexpr int EPI_WARPS = 4 * TS_UW * TILES;
    constexpr uint32_t TMEM_COLS = TILES * TS_TILE_COLS;

    // ---- one-time setup: TMEM, barriers, weights -> shared-memory B operand -----------------------------------
    if (warp == EPI_WARPS) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    if (tid == 0) {
        for (int t = 0; t < TILES; ++t) {
            mbar_init(&bars[t], 4 * TS_UW);         // one arrive per epilogue warp of the tile
            mbar_init(&bars[2 + t], 1);             // tcgen05.commit
            mbar_init(&bars[4 + t], 4 * TS_UW);     // middle of the tile's gate math
        }
        fence_mbar_init();
    }
    {
        // image: [row n = gate * 64 + unit][KB elements], k contiguous (pack_tc_images); the B operand is K-major:
        // 16-byte k chunks at LBO, 8-row groups at SBO
        const uint4* img = reinterpret_cast<const uint4*>(a.blob + BlobLayout::tc_image(FMT));
        for (int idx = tid; idx < TS_N * F::CHUNKS; idx += blockDim.x) {
            const int n = idx / F::CHUNKS, c = idx % F::CHUNKS;
            *reinterpret_cast<uint4*>(bop + c * TS_LBO + (n >> 3) * TS_SBO + (n & 7) * 16) = img[idx];
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
    int* const job_slot = reinterpret_cast<int*>(smem + F::OFF_BAR + 52);
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
                constexpr uint32_t idesc = instr_desc(F::KIND, TS_M, TS_N);
                const uint32_t d_base = tmem + (uint32_t)(tile * TS_TILE_COLS);
                const uint32_t a_base = d_base + TS_A_OFF;
                const uint32_t b_base = smem_u32(bop);
                for (int t = 0; t < nsteps; ++t) {
                    mbar_wait_sleep(&bars[tile], (uint32_t)((base + t) & 1));
                    tc_fence_after();
#pragma unroll
                    for (int i = 0; i < F::NMMA; ++i) {
                        // one MMA = 8 operand columns of TMEM (16 two-byte or 8 tf32 elements); the strict form's third
                        // K block (ks 8..11, the weight residual) re-reads the h_hi columns.  The strict form issues its two
                        // correction blocks FIRST: the tensor core truncates every accumulation, and a 2^-11-sized term
                        // added to the full-sized sum would lose its low bits to that truncation (a bias the recurrence
                        // integrates); added while the accumulator is still small they are summed essentially exactly.
                        const int ks = F::STRICT ? (i + 4) % F::NMMA : i;
                        const int acol = (F::STRICT && ks >= 8) ? (ks - 8) * 8 : ks * 8;
                        mma_ts<F::KIND>(d_base, a_base + acol, smem_desc(b_base + ks * 2 * TS_LBO, TS_LBO, TS_SBO), idesc, i > 0);
                    }
                    mma_commit(&bars[2 + tile]);
                }
            }
        } else {
            // ================================ epilogue warps ======================================================
            const int tile = warp / (4 * TS_UW);
            const int wq = warp & 3;                 // TMEM lane quarter (== warp id % 4)
            if (((warp >> 2) & 1) == 0)
                tcs_epilogue<FMT, TILES, VAR, 0>(a, kc, tile, wq, lane, tmem, bars, ypart, group, t0, nsteps, base);
            else
                tcs_epilogue<FMT, TILES, VAR, 1>(a, kc, tile, wq, lane, tmem, bars, ypart, group, t0, nsteps, base);
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

template <int FMT, int TILES, int VAR>
cudaError_t launch_tcs_one(const GruArgs& a, const TcsConsts& kc, int sm_count, bool dynamic, cudaStream_t st)
{
    using F = TsFmt<FMT>;
    static OncePerDevice once;
    cudaError_t e = once.run([] {
        return cudaFuncSetAttribute(gru_tcs_kernel<FMT, TILES, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F::SMEM_BYTES);
    });
    if (e != cudaSuccess) return e;
    const long long per_group = (long long)TS_M * TILES;
    const long long groups = (a.B + per_group - 1) / per_group;
    TcsSched sc{};
    long long grid = groups;
    // More groups than SMs and a partial last wave: pull (group, time chunk) jobs dynamically.  Chunk length: a job
    // costs ~7 steps of overhead (state round trip through L2, pipeline fill/drain; measured) and the tail idles half a
    // job per SM on average, so ct ~ sqrt(2 * 7 * T * groups / SMs), as a power of two in 64 .. 1024.
    if (groups > sm_count && groups % sm_count != 0 && groups < (1ll << 30) && dynamic) {
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
    gru_tcs_kernel<FMT, TILES, VAR><<<(unsigned)grid, 32 * (4 * TS_UW + 1) * TILES, F::SMEM_BYTES, st>>>(a, kc, sc);
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
    using B = BlobLayout;
    float wmax = 0.0f;
    for (int j = 0; j < 64; ++j) {
        kc->cn_w[j] = 2.0f * L * blob_host[B::W_IH + 128 + j];
        kc->cn_b[j] = 2.0f * L * blob_host[B::B_IH + 128 + j];
        kc->ch_b[j] = 2.0f * L * blob_host[B::B_HH + 128 + j];
        kc->cr_w[j] = -L * blob_host[B::W_IH + j];
        kc->cr_b[j] = -L * (blob_host[B::B_IH + j] + blob_host[B::B_HH + j]);
        kc->cz_w[j] = -L * blob_host[B::W_IH + 64 + j];
        kc->cz_b[j] = -L * (blob_host[B::B_IH + 64 + j] + blob_host[B::B_HH + 64 + j]);
        kc->wo[j] = blob_host[B::W_OUT + j];
    }
    kc->bo = blob_host[B::B_OUT];
    for (int i = 0; i < G192 * H64; ++i) wmax = fmaxf(wmax, fabsf(2.0f * L * blob_host[B::W_HH + i]));
    // strict form: the largest power of two that keeps gain * |scaled W| below 2^13 (f16 range with room for the 12-term
    // accumulation), at most 2^6: the weight residual (G W)_lo then stays a normal f16 number down to |W| ~ 2^-9
    float gain = 64.0f;
    while (gain > 1.0f && gain * wmax > 8192.0f) gain *= 0.5f;
    kc->gain_inv = 1.0f / gain;
}

// host: the B-operand images of the stream-major kernel (BlobLayout::IMG_*), from the fp32 part of the blob
void pack_tc_images(float* blob_host)
{
    constexpr float L = 1.4426950408889634f;
    using B = BlobLayout;
    const float* w_hh = blob_host + B::W_HH;
    const float* w_ih = blob_host + B::W_IH;
    const float* b_ih = blob_host + B::B_IH;
    const float* b_hh = blob_host + B::B_HH;
    TcsConsts kc;
    fill_tcs_consts(blob_host, &kc);
    const float gain = 1.0f / kc.gain_inv;
    auto bits16 = [](int fmt, float v) -> uint16_t {
        if (fmt == FMT_BF16) return static_cast<__nv_bfloat16_raw>(__float2bfloat16_rn(v)).x;
        return static_cast<__half_raw>(__float2half_rn(v)).x;
    };
    auto rnd16 = [](int fmt, float v) {
        return fmt == FMT_BF16 ? __bfloat162float(__float2bfloat16_rn(v)) : __half2float(__float2half_rn(v));
    };
    for (int n = 0; n < G192; ++n) {
        const int gate = n >> 6;
        const float scale = gate < 2 ? -L : 2.0f * L;
        // ---- f16 / bf16: 64 weights + the K augmentation [x_hi, x_lo, x_hi, 1, 1, 0 ...] -----------------------------
        for (int fmt = 0; fmt < 2; ++fmt) {
            uint16_t* row = reinterpret_cast<uint16_t*>(blob_host + B::tc_image(fmt)) + n * 80;
            memset(row, 0, 80 * 2);
            for (int k = 0; k < 64; ++k) row[k] = bits16(fmt, scale * w_hh[n * 64 + k]);
            auto put_split = [&](int k_hi, int k_lo, float v) {
                const float hi = rnd16(fmt, v);
                row[k_hi] = bits16(fmt, hi);
                row[k_lo] = bits16(fmt, v - hi);
            };
            if (gate < 2) {
                const float wi = scale * w_ih[n];
                row[64] = bits16(fmt, wi);            // x_hi * w_hi
                put_split(65, 66, wi);                // x_lo * w_hi + x_hi * w_lo
                put_split(67, 68, scale * (b_ih[n] + b_hh[n]));
            } else {
                put_split(67, 68, scale * b_hh[n]);
            }
        }
        // ---- strict: [G W_hi | G W_hi / 2^8 | (G W)_lo] ------------------------------------------------------------------
        {
            uint16_t* row = reinterpret_cast<uint16_t*>(blob_host + B::IMG_F16X3) + n * 192;
            for (int k = 0; k < 64; ++k) {
                const float w = gain * (scale * w_hh[n * 64 + k]);
                const float hi = rnd16(FMT_F16, w);
                row[k] = bits16(FMT_F16, hi);
                row[64 + k] = bits16(FMT_F16, hi * (1.0f / TS_LO_SCALE));
                row[128 + k] = bits16(FMT_F16, w - hi);
            }
        }
        // ---- tf32 (cvt.rna: round to nearest, ties away; low 13 bits zero) --------------------------------------------------
        {
            uint32_t* row = reinterpret_cast<uint32_t*>(blob_host + B::IMG_TF32) + n * 64;
            for (int k = 0; k < 64; ++k) {
                const float w = scale * w_hh[n * 64 + k];
                uint32_t u;
                memcpy(&u, &w, 4);
                if ((u & 0x7f800000u) != 0x7f800000u) u += 0x1000u;
                row[k] = u & 0xffffe000u;
            }
        }
    }
}

namespace {

template <int FMT, int TILES>
cudaError_t launch_tcs_var(const GruArgs& a, const TcsConsts& kc, int var, int sm_count, cudaStream_t st)
{
    const bool dynamic = var < 0 || (var & 32) == 0;
    if (var >= 0) var &= 31;
    switch (var) {      // experiments: (var & 15) = VAR bits, (var & 32) = static schedule
        case 3: return launch_tcs_one<FMT, TILES, 3>(a, kc, sm_count, dynamic, st);
        case 7: if (FMT < 2) return launch_tcs_one<FMT < 2 ? FMT : 0, TILES, 7>(a, kc, sm_count, dynamic, st);      // (elsewhere 7 == 15)
        default: return launch_tcs_one<FMT, TILES, TCS_DEFAULT_VAR>(a, kc, sm_count, dynamic, st);
    }
}

}  // namespace

// fmt: FMT_F16 / FMT_BF16 / FMT_TF32 / FMT_F16X3.  tiles: 128-stream tiles per CTA (1 or 2; 0 = automatic); var: kernel
// variant (experiments; -1 = default).  DiffDelRNN batches (a.d != nullptr): the GRU + head pass writes pre_d, the delay
// read runs as a second, HBM-bound pass over it (csrc/delay.cu: 12 bytes per sample at ~4.7 TB/s against ~75 ps per
// sample of recurrent work: +3 %) -- the taps reach up to D samples back into other time chunks' output, which the
// job-queue schedule of this kernel may still be producing on another SM.
cudaError_t launch_gru_tcs(const GruArgs& a0, const TcsConsts& kc, int fmt, int sm_count, int tiles, int var, cudaStream_t st)
{
    if (a0.B <= 0 || a0.T <= 0) return cudaSuccess;
    GruArgs a = a0;
    const bool delay = a0.d != nullptr;
    if (delay) {                               // pass 1: plain GRU + head into pre_d
        a.y = a0.pre; a.ldy = a0.ldp;
        a.d = nullptr; a.pre = nullptr;
    }
    // one tile per SM (time-multiplexed by the job queue) beats two resident tiles until ~190 streams per SM (measured)
    if (tiles <= 0) tiles = a.B >= 190ll * sm_count ? 2 : 1;
    cudaError_t e;
    switch (fmt) {
        case FMT_BF16: e = tiles >= 2 ? launch_tcs_var<FMT_BF16, 2>(a, kc, var, sm_count, st) : launch_tcs_var<FMT_BF16, 1>(a, kc, var, sm_count, st); break;
        case FMT_TF32: e = tiles >= 2 ? launch_tcs_var<FMT_TF32, 2>(a, kc, var, sm_count, st) : launch_tcs_var<FMT_TF32, 1>(a, kc, var, sm_count, st); break;
        case FMT_F16X3: e = tiles >= 2 ? launch_tcs_var<FMT_F16X3, 2>(a, kc, var, sm_count, st) : launch_tcs_var<FMT_F16X3, 1>(a, kc, var, sm_count, st); break;
        default: e = tiles >= 2 ? launch_tcs_var<FMT_F16, 2>(a, kc, var, sm_count, st) : launch_tcs_var<FMT_F16, 1>(a, kc, var, sm_count, st); break;
    }
    if (e != cudaSuccess || !delay) return e;
    // pass 2: y = delay(pre_d, d) with the carried history (warm-up: y = pre_d, history still rolled)
    return launch_delay(a0.pre, a0.ldp, a0.d, a0.ldd, a0.y, a0.ldy, a0.hist_in, a0.hist_out, a0.B, a0.T, a0.D, a0.warmup, st);
}

}  // namespace ntm
