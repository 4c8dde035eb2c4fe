// GRU-HS[64] forward, warp-level mma.sync, EIGHT warps per CTA and eight streams per CTA: the form for one CTA per SM with
// a full n8 tile (593 .. 1184 streams on 148 SMs = BASELINE config 2).
//
// Follows RNN.forward, code/model.py:67-88 (torch rnn.py:1221-1224 gate equations), like gru_mma.cu.  Why another form:
// with two 4-stream CTAs per SM (gru_mma.cu, HALF) every SM sub-partition issues 26 HMMAs per step, half of them on the
// unused odd columns of the n8 tiles -- 208 clk of tensor-pipe time in a 434-clk step.  Here warp w owns the eight hidden
// units [8w, 8w + 8) as TWO m16n8k16 tiles
//     tile RZ: rows 0-7 = r of the warp's units, rows 8-15 = z of the same units
//     tile NH: rows 0-7 = n of the warp's units, row 8 = w_out rounded, row 9 = its rounding residual, rows 10-15 zero
// so that a thread (gid, tig) holds r, z, n of ONE unit (8w + gid) for the two streams 2 tig, 2 tig + 1 -- the same two
// (unit, stream) pairs of gate math per thread as the 4-stream form -- and the output head comes for free out of the
// padding rows of the n tile (every warp computes it, warp 0 stores it).  Per sub-partition and step: 16 HMMAs on full
// tiles instead of 26 on half-empty ones, the same 22 MUFU instructions.  Plain GRU, f16 / bf16 operands.
//
// MEASURED: 262.6 ns/step at 593 .. 1184 streams against 221-224 ns for the two 4-stream CTAs -- the eight warps advance in
// lock-step behind one barrier (all in their HMMA phase, then all in their MUFU phase), while two independent CTAs overlap
// one's MMAs with the other's gate math; the tensor-pipe time saved does not make up for that.  Parity-checked and kept
// selectable (ntm_set_tuning(8, 5)) as the record of that experiment; the dispatcher never picks it.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gates.cuh"
#include "ntm_common.cuh"

namespace ntm {

namespace {

constexpr int M8_S = 8;                         // streams per CTA
constexpr int M8_CH = 128;                      // steps per staged chunk
constexpr int M8_ROW = 64 * 2 + 16;             // padded row of the state tile (bytes)
constexpr int M8_HB = M8_S * M8_ROW;            // one state tile
constexpr int M8_THREADS = 256;

template <bool BF16>
__device__ __forceinline__ uint32_t m8_pack2(float lo, float hi)
{
    if (BF16) {
        uint32_t y;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
        return y;
    }
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

template <bool BF16>
__device__ __forceinline__ float m8_round(float v)
{
    return BF16 ? __bfloat162float(__float2bfloat16_rn(v)) : __half2float(__float2half_rn(v));
}

template <bool BF16>
__device__ __forceinline__ void m8_store1(uint8_t* p, float v)
{
    if (BF16) *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(v);
    else *reinterpret_cast<__half*>(p) = __float2half_rn(v);
}

template <bool BF16>
__device__ __forceinline__ void m8_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    if (BF16)
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool BF16>
__global__ void __launch_bounds__(M8_THREADS, 1) gru_mma8_kernel(const GruArgs a)
{
    constexpr float L2E = 1.4426950408889634f;
    __shared__ __align__(128) uint8_t hb[2 * M8_HB];              // rounded state, double-buffered: [stream][64 units]
    __shared__ __align__(16) float xs[2 * M8_CH * M8_S];          // staged input: [buf][step][stream]
    __shared__ __align__(16) float yp[M8_CH * M8_S];              // head outputs of the chunk: [step][stream]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const int u = 8 * warp + gid;                                 // this thread's hidden unit
    const long long b0 = (long long)blockIdx.x * M8_S;
    const int ns = (int)((a.B - b0) < (long long)M8_S ? (a.B - b0) : (long long)M8_S);
    const float* __restrict__ blob = a.blob;

    // ---- A fragments (weights) -> registers.  k-step ks, thread tig <-> hidden indices tig * 16 + 4 ks + {0..3}, the
    // bijection of gru_mma.cu: a thread's B fragments of a whole step are two 16-byte loads of a state row.
    uint32_t arz[4][4], anh[4][4];
    {
        const float* wr = blob + BlobLayout::W_HH + (0 * 64 + u) * 64;
        const float* wz = blob + BlobLayout::W_HH + (1 * 64 + u) * 64;
        const float* wn = blob + BlobLayout::W_HH + (2 * 64 + u) * 64;
        const float* wo = blob + BlobLayout::W_OUT;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int k = tig * 16 + 4 * ks;
            float hd[4];                                          // row gid + 8 of the n tile: head weights / residual / zero
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float w = wo[k + i], hi = m8_round<BF16>(w);
                hd[i] = gid == 0 ? hi : (gid == 1 ? w - hi : 0.0f);
            }
            arz[ks][0] = m8_pack2<BF16>(-L2E * wr[k], -L2E * wr[k + 1]);
            arz[ks][1] = m8_pack2<BF16>(-L2E * wz[k], -L2E * wz[k + 1]);
            arz[ks][2] = m8_pack2<BF16>(-L2E * wr[k + 2], -L2E * wr[k + 3]);
            arz[ks][3] = m8_pack2<BF16>(-L2E * wz[k + 2], -L2E * wz[k + 3]);
            anh[ks][0] = m8_pack2<BF16>(2.0f * L2E * wn[k], 2.0f * L2E * wn[k + 1]);
            anh[ks][1] = m8_pack2<BF16>(hd[0], hd[1]);
            anh[ks][2] = m8_pack2<BF16>(2.0f * L2E * wn[k + 2], 2.0f * L2E * wn[k + 3]);
            anh[ks][3] = m8_pack2<BF16>(hd[2], hd[3]);
        }
    }
    const UnitConst uc = load_unit_const(blob, u);
    const float bo = blob[BlobLayout::B_OUT];

    auto load_x = [&](int buf, long long t0) {
        const int n = (int)((a.T - t0) < (long long)M8_CH ? (a.T - t0) : (long long)M8_CH);
        float* dstb = xs + buf * M8_CH * M8_S;
        for (int idx = tid; idx < M8_CH * M8_S; idx += M8_THREADS) {
            const int s = idx % M8_S, tt = idx / M8_S;
            if (s < ns && tt < n) cp_async4(dstb + tt * M8_S + s, a.x + (b0 + s) * a.ldx + t0 + tt);
            else dstb[tt * M8_S + s] = 0.0f;
        }
        cp_async_commit();
    };

    // ---- initial state: fp32 in registers (hst[e]: stream 2 tig + e), rounded copy into state tile 0 ---------------
    float hst[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int s = 2 * tig + e;
        hst[e] = (s < ns && a.h_in) ? a.h_in[(b0 + s) * 64 + u] : 0.0f;
        m8_store1<BF16>(hb + s * M8_ROW + u * 2, hst[e]);
        m8_store1<BF16>(hb + M8_HB + s * M8_ROW + u * 2, 0.0f);
    }
    const long long nchunks = (a.T + M8_CH - 1) / M8_CH;
    int cur = 0;
    load_x(0, 0);
    for (long long c = 0; c < nchunks; ++c) {
        const long long t0 = c * M8_CH;
        const int n = (int)((a.T - t0) < (long long)M8_CH ? (a.T - t0) : (long long)M8_CH);
        const int xb = (int)(c & 1);
        const float* xcur = xs + xb * M8_CH * M8_S;
        cp_async_wait_all();
        __syncthreads();                       // xs[xb] landed; state tile `cur` complete; previous flush done
        if (c + 1 < nchunks) load_x(xb ^ 1, t0 + M8_CH);

        for (int tt = 0; tt < n; ++tt) {
            const uint8_t* hcur = hb + cur * M8_HB;
            uint8_t* hnext = hb + (cur ^ 1) * M8_HB;
            uint32_t b[8];
            {
                const uint8_t* src = hcur + gid * M8_ROW + tig * 32;
                const uint4 v0 = *reinterpret_cast<const uint4*>(src), v1 = *reinterpret_cast<const uint4*>(src + 16);
                b[0] = v0.x; b[1] = v0.y; b[2] = v0.z; b[3] = v0.w;
                b[4] = v1.x; b[5] = v1.y; b[6] = v1.z; b[7] = v1.w;
            }
            const float2 xv = *reinterpret_cast<const float2*>(xcur + tt * M8_S + 2 * tig);
            // input projection and biases enter through the accumulators; the head rows start at zero
            float crz[4] = {fmaf(uc.cr_w, xv.x, uc.cr_b), fmaf(uc.cr_w, xv.y, uc.cr_b),
                            fmaf(uc.cz_w, xv.x, uc.cz_b), fmaf(uc.cz_w, xv.y, uc.cz_b)};
            float cnh[4] = {uc.ch_b, uc.ch_b, 0.0f, 0.0f};
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                m8_mma<BF16>(crz, arz[ks], b[2 * ks], b[2 * ks + 1]);
                m8_mma<BF16>(cnh, anh[ks], b[2 * ks], b[2 * ks + 1]);
            }
            // head of the PREVIOUS step (the B fragments are its state): row 8 (gid 0) + residual row 9 (gid 1)
            // (consuming it one iteration later, as the 4-stream form does, measured slower here: 272.5 vs 262.6 ns/step)
            if (warp == 0) {
                const float lo0 = __shfl_down_sync(0xffffffffu, cnh[2], 4), lo1 = __shfl_down_sync(0xffffffffu, cnh[3], 4);
                if (gid == 0 && tt > 0)
                    *reinterpret_cast<float2*>(yp + (tt - 1) * M8_S + 2 * tig) = make_float2(cnh[2] + lo0, cnh[3] + lo1);
            }
            // gates, state blend, rounded state for the next step
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float xx = e ? xv.y : xv.x;
                const float r = rcp_approx(1.0f + ex2_approx(crz[e]));
                const float dn = 1.0f + ex2_approx(fminf(fmaf(r, cnh[e], fmaf(uc.cn_w, xx, uc.cn_b)), EX2_CLAMP));
                const float z = rcp_approx(1.0f + ex2_approx(crz[2 + e]));
                hst[e] = gates_blend1(z, dn, hst[e]);
                m8_store1<BF16>(hnext + (2 * tig + e) * M8_ROW + u * 2, hst[e]);
            }
            cur ^= 1;
            __syncthreads();                   // next state tile published; all reads of the old one are done
        }

        if (n > 0 && warp == 0) {              // head of the chunk's last step: the n tile once more on the final state
            const uint8_t* src = hb + cur * M8_HB + gid * M8_ROW + tig * 32;
            const uint4 v0 = *reinterpret_cast<const uint4*>(src), v1 = *reinterpret_cast<const uint4*>(src + 16);
            const uint32_t b[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            float ch[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) m8_mma<BF16>(ch, anh[ks], b[2 * ks], b[2 * ks + 1]);
            const float lo0 = __shfl_down_sync(0xffffffffu, ch[2], 4), lo1 = __shfl_down_sync(0xffffffffu, ch[3], 4);
            if (gid == 0) *reinterpret_cast<float2*>(yp + (n - 1) * M8_S + 2 * tig) = make_float2(ch[2] + lo0, ch[3] + lo1);
        }
        __syncthreads();
        // ---- flush the chunk: y = head + bias (+ x) ----------------------------------------------------------------
        for (int idx = tid; idx < M8_S * M8_CH; idx += M8_THREADS) {
            const int s = idx / M8_CH, tt = idx % M8_CH;
            if (s < ns && tt < n) {
                float v = yp[tt * M8_S + s] + bo;
                if (a.skip) v += xcur[tt * M8_S + s];
                a.y[(b0 + s) * a.ldy + t0 + tt] = v;
            }
        }
    }
    // ---- final state ---------------------------------------------------------------------------------------------------
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int s = 2 * tig + e;
        if (s < ns) a.h_out[(b0 + s) * 64 + u] = hst[e];
    }
}

}  // namespace

// fmt: FMT_F16 (0) / FMT_BF16 (1); plain GRU only (a.d == nullptr).
cudaError_t launch_gru_mma8(const GruArgs& a, int fmt, cudaStream_t st)
{
    if (a.B <= 0 || a.T <= 0) return cudaSuccess;
    if (a.d != nullptr || (fmt != 0 && fmt != 1)) return cudaErrorInvalidValue;
    const long long grid = (a.B + M8_S - 1) / M8_S;
    if (fmt == 1) gru_mma8_kernel<true><<<(unsigned)grid, M8_THREADS, 0, st>>>(a);
    else gru_mma8_kernel<false><<<(unsigned)grid, M8_THREADS, 0, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace ntm
