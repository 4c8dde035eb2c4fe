// Exact-fp32 persistent GRU kernel (CUDA cores) -- the parity anchor and the batch-1 real-time path.
//
// Replaces the per-timestep loop behind `self.GRU(x, self.hidden)` + `self.output(x)` of
// RNN.forward / DiffDelRNN.forward (code/model.py:81-82, :412-413; gate equations torch rnn.py:1221-1224)
// and, for DiffDelRNN, the delay read of code/model.py:422 fused at every 64-sample flush.
//
// Layout: one CTA owns S streams for ALL T timesteps (no per-step launch, no HBM round trip of the state).
//   - W_hh (192x64) lives in REGISTERS: thread (j, q) of 64*KS threads holds the three gate rows of hidden
//     unit j restricted to the k-slice q (3 * 64/KS floats), for the whole kernel.
//   - the hidden state of the S streams is double-buffered in shared memory; every step each thread reads
//     its k-slice with broadcast LDS.128, does 3*64/KS FMAs per stream, and the KS partial sums are combined
//     with a transposing butterfly (so the KS lanes of a hidden unit end up owning different streams and no
//     activation is computed twice).
//   - gates, state blend and the 64-wide head dot are fused; x is staged in 64-sample chunks with cp.async
//     (double-buffered), y is staged in shared memory and flushed with coalesced stores.
#include "ntm_common.cuh"

namespace ntm {

namespace {

constexpr int CH = 64;   // samples per staged chunk (even: the h double-buffer parity restarts per chunk)

// accurate activations (no fast-math): expf / tanhf are libm-grade (<= 2 ulp), division is IEEE
__device__ __forceinline__ float sigmoid_acc(float a) { return 1.0f / (1.0f + expf(-a)); }

template <int KS, int SP>
struct Transpose;   // combine the KS partial sums; lane q ends up owning streams i*KS + q

template <int SP>
struct Transpose<1, SP> {
    static __device__ __forceinline__ void run(const float (&acc)[SP], float (&own)[SP], int) {
#pragma unroll
        for (int i = 0; i < SP; ++i) own[i] = acc[i];
    }
};
template <int SP>
struct Transpose<2, SP> {
    static __device__ __forceinline__ void run(const float (&acc)[SP], float (&own)[SP / 2], int q) {
        const bool hi = q & 1;
#pragma unroll
        for (int i = 0; i < SP / 2; ++i) {
            const float keep = hi ? acc[2 * i + 1] : acc[2 * i];
            const float send = hi ? acc[2 * i] : acc[2 * i + 1];
            own[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
    }
};
template <int SP>
struct Transpose<4, SP> {
    static __device__ __forceinline__ void run(const float (&acc)[SP], float (&own)[SP / 4], int q) {
        const bool b0 = q & 1, b1 = q & 2;
        float t1[SP / 2];
#pragma unroll
        for (int p = 0; p < SP / 2; ++p) {
            const float keep = b0 ? acc[2 * p + 1] : acc[2 * p];
            const float send = b0 ? acc[2 * p] : acc[2 * p + 1];
            t1[p] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
#pragma unroll
        for (int i = 0; i < SP / 4; ++i) {
            const float keep = b1 ? t1[2 * i + 1] : t1[2 * i];
            const float send = b1 ? t1[2 * i] : t1[2 * i + 1];
            own[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
    }
};

template <int KS, int S>
__global__ void __launch_bounds__(64 * KS) gru_fp32_kernel(const GruArgs a)
{
    constexpr int NT = 64 * KS;          // threads
    constexpr int NW = NT / 32;          // warps
    constexpr int KK = 64 / KS;          // k values per thread
    constexpr int HS = 64 + 32 / KS;     // smem row stride of a stream's state (conflict-free STS)
    constexpr int SO = (S + KS - 1) / KS;  // streams owned per lane after the transpose
    constexpr int SP = SO * KS;          // S padded to a multiple of KS

    __shared__ __align__(16) float hbuf[2][S][HS];
    __shared__ float xs[2][S][CH];
    __shared__ float ys[S][CH];
    __shared__ float ypart[2][NW][SP];

    const int tid = threadIdx.x;
    const int j = tid / KS;              // hidden unit
    const int q = tid % KS;              // k-slice, later: owned-stream residue
    const int lane = tid & 31, warp = tid >> 5;
    const long long b0 = (long long)blockIdx.x * S;
    const int ns = (int)((a.B - b0) < (long long)S ? (a.B - b0) : (long long)S);

    // ---- parameters -> registers -------------------------------------------------------------
    const float* __restrict__ blob = a.blob;
    float w[3][KK];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int i = 0; i < KK / 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(blob + BlobLayout::W_HH + (g * 64 + j) * 64 +
                                                              4 * (q + KS * i));
            w[g][4 * i + 0] = v.x; w[g][4 * i + 1] = v.y; w[g][4 * i + 2] = v.z; w[g][4 * i + 3] = v.w;
        }
    const float wir = blob[BlobLayout::W_IH + j], wiz = blob[BlobLayout::W_IH + 64 + j],
                win = blob[BlobLayout::W_IH + 128 + j];
    const float br = blob[BlobLayout::B_IH + j] + blob[BlobLayout::B_HH + j];
    const float bz = blob[BlobLayout::B_IH + 64 + j] + blob[BlobLayout::B_HH + 64 + j];
    const float bin = blob[BlobLayout::B_IH + 128 + j], bhn = blob[BlobLayout::B_HH + 128 + j];
    const float wo = blob[BlobLayout::W_OUT + j];
    const float bo = blob[BlobLayout::B_OUT];

    const bool delay = a.d != nullptr;
    float* __restrict__ head_out = delay ? a.pre : a.y;
    const long long ldo = delay ? a.ldp : a.ldy;

    // ---- initial state -------------------------------------------------------------------------
    for (int idx = tid; idx < S * 64; idx += NT) {
        const int s = idx >> 6, k = idx & 63;
        hbuf[0][s][k] = (s < ns && a.h_in) ? a.h_in[(b0 + s) * 64 + k] : 0.0f;
    }
    auto load_x = [&](int buf, long long t0) {
        const int n = (int)((a.T - t0) < (long long)CH ? (a.T - t0) : (long long)CH);
        for (int idx = tid; idx < S * CH; idx += NT) {
            const int s = idx / CH, tt = idx % CH;
            if (s < ns && tt < n) cp_async4(&xs[buf][s][tt], a.x + (b0 + s) * a.ldx + t0 + tt);
            else xs[buf][s][tt] = 0.0f;
        }
        cp_async_commit();
    };
    load_x(0, 0);
    __syncthreads();
    float hown[SO];
#pragma unroll
    for (int i = 0; i < SO; ++i) {
        const int s = i * KS + q;
        hown[i] = (s < S) ? hbuf[0][s][j] : 0.0f;
    }

    const long long nchunks = (a.T + CH - 1) / CH;
    for (long long c = 0; c < nchunks; ++c) {
        const long long t0 = c * CH;
        const int n = (int)((a.T - t0) < (long long)CH ? (a.T - t0) : (long long)CH);
        const int xb = (int)(c & 1);
        cp_async_wait_all();
        __syncthreads();                       // xs[xb] landed; ys and xs[xb^1] free again
        if (c + 1 < nchunks) load_x(xb ^ 1, t0 + CH);

        for (int tt = 0; tt < n; ++tt) {
            const int cur = tt & 1;
            // ---- partial mat-vec: 3 gate rows x k-slice, all S streams --------------------------
            float acc[3][SP];
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int s = 0; s < SP; ++s) acc[g][s] = 0.0f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
#pragma unroll
                for (int i = 0; i < KK / 4; ++i) {
                    const float4 hv = *reinterpret_cast<const float4*>(&hbuf[cur][s][4 * (q + KS * i)]);
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        float v = acc[g][s];
                        v = fmaf(w[g][4 * i + 0], hv.x, v);
                        v = fmaf(w[g][4 * i + 1], hv.y, v);
                        v = fmaf(w[g][4 * i + 2], hv.z, v);
                        v = fmaf(w[g][4 * i + 3], hv.w, v);
                        acc[g][s] = v;
                    }
                }
            }
            float own[3][SO];
            Transpose<KS, SP>::run(acc[0], own[0], q);
            Transpose<KS, SP>::run(acc[1], own[1], q);
            Transpose<KS, SP>::run(acc[2], own[2], q);

            // ---- gates + state blend + head partial for the owned (j, stream) pairs --------------
#pragma unroll
            for (int i = 0; i < SO; ++i) {
                const int s = i * KS + q;
                const bool live = s < S;             // padded lanes compute garbage in lock-step, store nothing
                const float xv = xs[xb][live ? s : 0][tt];
                const float r = sigmoid_acc(fmaf(wir, xv, br) + own[0][i]);
                const float z = sigmoid_acc(fmaf(wiz, xv, bz) + own[1][i]);
                const float nn = tanhf(fmaf(r, own[2][i] + bhn, fmaf(win, xv, bin)));
                const float hn = fmaf(hown[i] - nn, z, nn);      // (h - n) * z + n
                hown[i] = hn;
                if (live) hbuf[cur ^ 1][s][j] = hn;
                float p = wo * hn;
#pragma unroll
                for (int off = KS; off < 32; off <<= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
                if (live && lane < KS) ypart[cur][warp][s] = p;
            }
            __syncthreads();                   // new state + head partials published
            if (tid < S) {
                float v = bo;
#pragma unroll
                for (int wv = 0; wv < NW; ++wv) v += ypart[cur][wv][tid];
                if (a.skip) v += xs[xb][tid][tt];
                ys[tid][tt] = v;
            }
        }
        __syncthreads();                       // ys complete for this chunk
        for (int idx = tid; idx < S * CH; idx += NT) {
            const int s = idx / CH, tt = idx % CH;
            if (s < ns && tt < n) head_out[(b0 + s) * ldo + t0 + tt] = ys[s][tt];
        }
        if (delay) {
            __syncthreads();                   // this chunk's pre_d is visible CTA-wide (L2 reads below)
            for (int idx = tid; idx < S * CH; idx += NT) {
                const int s = idx / CH, tt = idx % CH;
                if (s < ns && tt < n) {
                    const long long t = t0 + tt;
                    float v;
                    if (a.warmup) {
                        v = ys[s][tt];
                    } else {
                        const float* prow = a.pre + (b0 + s) * a.ldp;
                        const float* hrow = a.hist_in + (b0 + s) * (long long)a.D;
                        v = delay_read(a.d[(b0 + s) * a.ldd + t], t, a.D, [&](long long i) {
                            return i >= 0 ? __ldcg(prow + i) : hrow[a.D + i];
                        });
                    }
                    a.y[(b0 + s) * a.ldy + t] = v;
                }
            }
        }
    }

    // ---- final state; rolled delay history (code/model.py:314-315) -------------------------------
#pragma unroll
    for (int i = 0; i < SO; ++i) {
        const int s = i * KS + q;
        if (s < ns) a.h_out[(b0 + s) * 64 + j] = hown[i];
    }
    if (delay) {
        __syncthreads();
        for (long long idx = tid; idx < (long long)ns * a.D; idx += NT) {
            const int s = (int)(idx / a.D);
            const long long i = idx % a.D;
            const long long src = a.T - a.D + i;
            a.hist_out[(b0 + s) * (long long)a.D + i] =
                src >= 0 ? __ldcg(a.pre + (b0 + s) * a.ldp + src) : a.hist_in[(b0 + s) * (long long)a.D + a.D + src];
        }
    }
}

template <int KS, int S>
cudaError_t launch_one(const GruArgs& a, cudaStream_t st)
{
    const long long grid = (a.B + S - 1) / S;
    gru_fp32_kernel<KS, S><<<(unsigned)grid, 64 * KS, 0, st>>>(a);
    ++g_launches;
    return cudaGetLastError();
}

template <int KS>
cudaError_t launch_ks(const GruArgs& a, int s, cudaStream_t st)
{
    switch (s) {
        case 1: return launch_one<KS, 1>(a, st);
        case 2: return launch_one<KS, 2>(a, st);
        case 4: return launch_one<KS, 4>(a, st);
        case 8: return launch_one<KS, 8>(a, st);
        default: return launch_one<KS, 16>(a, st);
    }
}

}  // namespace

cudaError_t launch_gru_fp32(const GruArgs& a, int sm_count, int tune_s, int tune_ks, cudaStream_t st)
{
    if (a.B <= 0 || a.T <= 0) return cudaSuccess;
    int s = tune_s;
    if (s <= 0) {
        // fill the SMs first (one CTA per SM), then grow the streams per CTA
        const long long per_sm = (a.B + sm_count - 1) / sm_count;
        s = 1;
        while (s < 16 && s < per_sm) s <<= 1;
    }
    if (s != 1 && s != 2 && s != 4 && s != 8) s = 16;
    const int ks = tune_ks > 0 ? tune_ks : 4;
    return ks == 2 ? launch_ks<2>(a, s, st) : launch_ks<4>(a, s, st);
}

}  // namespace ntm
