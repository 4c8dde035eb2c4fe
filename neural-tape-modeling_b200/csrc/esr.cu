// Evaluation losses on the device: ESR and DC-pre-emphasised ESR ("DCPreESR") of (B, 1, T) output / target pairs.
//
// Replaces (the step right after the recurrent path in the reference's evaluation, code/test-model.py:250-253,386-388):
//   ESRLoss            code/Automated_GuitarAmpModelling/CoreAudioML/training.py:5-16
//   ESRLoss(dc_pre)    code/GreyBoxDRC/loss_funcs.py:6-52: both signals zero-padded in front and filtered with the 2000-tap
//                      truncation of H(z) = (1 - z^-1) / (1 - R z^-1), R = 0.995 (h[0] = 1, h[k] = (R-1) R^(k-1)), then
//                      mean((f_t - f_o)^2) / (mean(f_t^2) + 1e-5) over all B*T elements.
// The kernel returns the two SUMS (double); the caller forms the ratio.
//
// The reference evaluates the FIR as a dense conv1d: 2 x 2000 multiply-adds per sample.  Here the filter is a scan:
//     f[n] = x[n] + (R - 1) S[n-1],     S[n] = sum_{j=0..1998} R^j x[n-j] = R S[n-1] + x[n] - R^1999 x[n-1999]
// (exactly the truncated FIR in exact arithmetic; the difference e = t - o is filtered instead of o, the filter being
// linear).  S is a first-order linear recurrence: a warp scans 128 consecutive samples per iteration (4 per lane, carries
// combined with 5 shuffle steps), in double precision because f nearly cancels for low-frequency content.  Because the
// window is finite, a scan may start from zero state anywhere >= 1999 samples before the first sample it is asked for:
// every warp owns one (stream, chunk) work item and warms up over the 2048 samples in front of its chunk, so items are
// independent and the pass is HBM-bound: 8 algorithmic bytes per sample (o and t read once), O(1) flops per sample.
#include "ntm_common.cuh"

namespace ntm {

namespace {

constexpr int ESR_TAPS = 2000;
constexpr int ESR_WARM = 2048;           // >= ESR_TAPS - 1, multiple of the 128-sample tile
constexpr int ESR_WARPS = 4;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

template <bool DCPRE>
__global__ void __launch_bounds__(32 * ESR_WARPS) esr_kernel(const float* __restrict__ out, long long ldo,
                                                             const float* __restrict__ tgt, long long ldt, long long B,
                                                             long long T, long long chunk, long long chunks, double R,
                                                             double Rw /* R^(ESR_TAPS-1) */, double* __restrict__ sums)
{
    const int lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * ESR_WARPS + (threadIdx.x >> 5);
    if (item >= B * chunks) return;
    const long long b = item / chunks, c = item % chunks;
    const long long c0 = c * chunk, c1 = (c0 + chunk) < T ? (c0 + chunk) : T;
    const float* __restrict__ o = out + b * ldo;
    const float* __restrict__ t = tgt + b * ldt;
    double num = 0.0, den = 0.0;

    if (!DCPRE) {
        for (long long n = c0 + lane; n < c1; n += 32) {
            const double tv = (double)__ldg(t + n), e = tv - (double)__ldg(o + n);
            num = fma(e, e, num);
            den = fma(tv, tv, den);
        }
    } else {
        const double R2 = R * R, D = R2 * R2;                    // decay over one lane's 4 samples
        const double Dk[5] = {D, D * D, (D * D) * (D * D), ((D * D) * (D * D)) * ((D * D) * (D * D)),
                              (((D * D) * (D * D)) * ((D * D) * (D * D))) * (((D * D) * (D * D)) * ((D * D) * (D * D)))};
        const double D32 = Dk[4] * Dk[4];
        double Dl = 1.0;                                          // D^lane
        for (int i = 0; i < lane; ++i) Dl *= D;
        double Dm[5];                                             // scan multipliers, zero where the partner lane is out of range
#pragma unroll
        for (int k = 0; k < 5; ++k) Dm[k] = lane >= (1 << k) ? Dk[k] : 0.0;
        const double g = R - 1.0;
        double se_tile = 0.0, st_tile = 0.0;                      // S of e / of t at the sample before the tile
        const long long n0 = c0 >= ESR_WARM ? c0 - ESR_WARM : 0;
        // raw samples of one tile: current and delayed-by-1999 values of t and o (zero outside [0, T))
        auto fetch = [&](long long nb, float (&tv)[4], float (&ov)[4], float (&td)[4], float (&od)[4]) {
            if (nb >= ESR_TAPS - 1 && nb + 128 <= T) {            // interior tile (warp-uniform): no bounds checks
                const float* tp = t + nb + 4 * lane;
                const float* op = o + nb + 4 * lane;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    tv[j] = __ldg(tp + j);
                    ov[j] = __ldg(op + j);
                    td[j] = __ldg(tp + j - (ESR_TAPS - 1));
                    od[j] = __ldg(op + j - (ESR_TAPS - 1));
                }
                return;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long nn = nb + 4 * lane + j, nd = nn - (ESR_TAPS - 1);
                const bool in = nn < T, ind = nd >= 0 && nd < T;
                tv[j] = in ? __ldg(t + nn) : 0.0f;
                ov[j] = in ? __ldg(o + nn) : 0.0f;
                td[j] = ind ? __ldg(t + nd) : 0.0f;
                od[j] = ind ? __ldg(o + nd) : 0.0f;
            }
        };
        float ntv[4], nov[4], ntd[4], nod[4];
        fetch(n0, ntv, nov, ntd, nod);
        for (long long nb = n0; nb < c1; nb += 128) {
            const long long n = nb + 4 * lane;
            double xe[4], xt[4], ue[4], ut[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                xt[j] = (double)ntv[j];
                xe[j] = (double)ntv[j] - (double)nov[j];
                ut[j] = fma(-Rw, (double)ntd[j], xt[j]);
                ue[j] = fma(-Rw, (double)ntd[j] - (double)nod[j], xe[j]);
            }
            // the next tile's loads are in flight while this one is scanned (the scan is a long dependent chain)
            if (nb + 128 < c1) fetch(nb + 128, ntv, nov, ntd, nod);
            // lane aggregates from zero state, inclusive decayed scan over the lanes
            double ae = fma(fma(fma(ue[0], R, ue[1]), R, ue[2]), R, ue[3]);
            double at = fma(fma(fma(ut[0], R, ut[1]), R, ut[2]), R, ut[3]);
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const double pe = __shfl_up_sync(0xffffffffu, ae, 1 << k), pt = __shfl_up_sync(0xffffffffu, at, 1 << k);
                ae = fma(Dm[k], pe, ae);
                at = fma(Dm[k], pt, at);
            }
            double se = __shfl_up_sync(0xffffffffu, ae, 1), st = __shfl_up_sync(0xffffffffu, at, 1);
            if (lane == 0) { se = 0.0; st = 0.0; }
            se = fma(Dl, se_tile, se);                            // S at the sample before this lane's first
            st = fma(Dl, st_tile, st);
            if (nb >= c0 && nb + 128 <= c1) {                     // tile entirely inside the chunk (warp-uniform)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double fe = fma(g, se, xe[j]), ft = fma(g, st, xt[j]);
                    num = fma(fe, fe, num);
                    den = fma(ft, ft, den);
                    se = fma(R, se, ue[j]);
                    st = fma(R, st, ut[j]);
                }
            } else if (nb >= c0) {                                // the chunk's last, partial tile
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double fe = fma(g, se, xe[j]), ft = fma(g, st, xt[j]);
                    if (n + j < c1) { num = fma(fe, fe, num); den = fma(ft, ft, den); }
                    se = fma(R, se, ue[j]);
                    st = fma(R, st, ut[j]);
                }
            }                                                     // warm-up tiles only feed the carry below
            se_tile = fma(D32, se_tile, __shfl_sync(0xffffffffu, ae, 31));
            st_tile = fma(D32, st_tile, __shfl_sync(0xffffffffu, at, 31));
        }
    }
    num = warp_sum(num);
    den = warp_sum(den);
    if (lane == 0) {
        atomicAdd(sums, num);
        atomicAdd(sums + 1, den);
    }
}

}  // namespace

// sums[0] = sum (f(t) - f(o))^2, sums[1] = sum f(t)^2 over B x T (sums is zeroed here, on the stream).
cudaError_t launch_esr(const float* out, long long ldo, const float* tgt, long long ldt, long long B, long long T,
                       int dc_pre, double* sums, int sm_count, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(sums, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess || B <= 0 || T <= 0) return e;
    // chunk: enough work items to fill the GPU (>= ~8 warps per SM sub-partition) but long against the warm-up
    long long chunk = 16384;
    while (chunk > 2048 && B * ((T + chunk - 1) / chunk) < 32ll * sm_count * ESR_WARPS) chunk >>= 1;
    const long long chunks = (T + chunk - 1) / chunk;
    const long long grid = (B * chunks + ESR_WARPS - 1) / ESR_WARPS;
    if (grid > 0x7fffffffll) return cudaErrorInvalidValue;
    const double R = 0.995;
    double Rw = 1.0;
    for (int i = 0; i < ESR_TAPS - 1; ++i) Rw *= R;
    if (dc_pre) esr_kernel<true><<<(unsigned)grid, 32 * ESR_WARPS, 0, st>>>(out, ldo, tgt, ldt, B, T, chunk, chunks, R, Rw, sums);
    else esr_kernel<false><<<(unsigned)grid, 32 * ESR_WARPS, 0, st>>>(out, ldo, tgt, ldt, B, T, chunk, chunks, R, Rw, sums);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace ntm
