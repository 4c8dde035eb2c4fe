// Stand-alone time-varying fractional delay line (HBM-bound elementwise gather).
// Replaces TimeVaryingDelayLine.forward, code/model.py:269-320, as used on its own by apply_delay
// (code/test-model.py:259-290).  The reference materialises three (B,1,T,D+1) temporaries; here every
// output sample reads two past samples (history or this call's input) -- 12 algorithmic bytes per sample.
#include "ntm_common.cuh"

namespace ntm {

namespace {

// VEC: rows of d and y are 16-byte aligned (base and leading dimension): a thread owns DELAY_SPT consecutive samples,
// loads their delays with 16-byte loads, issues all of its 2 x DELAY_SPT tap gathers as independent instructions, and
// stores 16 bytes at a time.  One sample per thread and iteration left the pass latency-bound (d -> index -> x -> y is a
// dependent chain with 4 bytes in flight per thread: 2.44 TB/s); this form reaches 3.73 TB/s.  (Tried: the same with lane-
// strided samples so that every access of a warp is one 128-byte segment -- 2.57 TB/s: scalar loads put fewer bytes in
// flight per thread, and bytes in flight are what bounds this two-phase gather; and a persistent block that stages the
// delays and the reachable x window of a 2048-sample tile in shared memory with double-buffered 16-byte cp.async -- the
// same 4.66 TB/s at D = 365 and 3.9 TB/s at D = 2000, where consecutive windows overlap by half: dropped.)
constexpr int DELAY_SPT = 8;

template <bool VEC>
__global__ void __launch_bounds__(256, 4) delay_kernel(const float* __restrict__ x, long long ldx,
                                                    const float* __restrict__ d, long long ldd,
                                                    float* __restrict__ y, long long ldy,
                                                    const float* __restrict__ hist_in, long long B, long long T,
                                                    int D, int warmup)
{
    const long long b = blockIdx.y;
    const float* xr = x + b * ldx;
    const float* hr = hist_in + b * (long long)D;
    const float* dr = d + b * ldd;
    float* yr = y + b * ldy;
    auto past = [&](long long i) { return i >= 0 ? __ldg(xr + i) : __ldg(hr + D + i); };
    {
        const long long groups = T / DELAY_SPT;
        for (long long gq = (long long)blockIdx.x * blockDim.x + threadIdx.x; gq < groups;
             gq += (long long)gridDim.x * blockDim.x) {
            const long long t = gq * DELAY_SPT;
            float dv[DELAY_SPT], v[DELAY_SPT];
            if (VEC) {
#pragma unroll
                for (int q = 0; q < DELAY_SPT / 4; ++q) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>((warmup ? xr : dr) + t) + q);
                    dv[4 * q] = a.x; dv[4 * q + 1] = a.y; dv[4 * q + 2] = a.z; dv[4 * q + 3] = a.w;
                }
            } else {                                              // rows not 16-byte aligned: same grouping, 4-byte accesses
#pragma unroll
                for (int j = 0; j < DELAY_SPT; ++j) dv[j] = __ldg((warmup ? xr : dr) + t + j);
            }
            if (!warmup) {
                // the arithmetic of delay_read (ntm_common.cuh) in three passes: tap weights and addresses, loads, sums.
                // A tap outside j in [0, D] gets weight 0 and a harmless address (adding an exact zero = skipping the tap).
                float w[DELAY_SPT][2], p[DELAY_SPT][2];
                int off[DELAY_SPT][2];                            // tap address relative to x[t] (32-bit: fewer registers, more
                                                                  // resident threads = more bytes in flight)
#pragma unroll
                for (int j = 0; j < DELAY_SPT; ++j) {
                    const float fl = floorf(dv[j]);
#pragma unroll
                    for (int tap = 0; tap < 2; ++tap) {
                        const float jf = fl + (float)tap;
                        const bool ok = jf >= 0.0f && jf <= (float)D;
                        w[j][tap] = ok ? fmaxf(__fsub_rn(1.0f, fabsf(__fsub_rn(jf, dv[j]))), 0.0f) : 0.0f;
                        off[j][tap] = ok ? j - (int)jf : j;
                    }
                }
                const float* xt = xr + t;
                if (t > (long long)D) {                           // no tap of this group can reach the carried history
#pragma unroll
                    for (int j = 0; j < DELAY_SPT; ++j) {
                        p[j][0] = __ldg(xt + off[j][0]);
                        p[j][1] = __ldg(xt + off[j][1]);
                    }
                } else {
                    const float* ht = hr + D + t;
#pragma unroll
                    for (int j = 0; j < DELAY_SPT; ++j) {
                        p[j][0] = __ldg((t + off[j][0] >= 0 ? xt : ht) + off[j][0]);
                        p[j][1] = __ldg((t + off[j][1] >= 0 ? xt : ht) + off[j][1]);
                    }
                }
#pragma unroll
                for (int j = 0; j < DELAY_SPT; ++j)
                    v[j] = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(w[j][1], p[j][1])), __fmul_rn(w[j][0], p[j][0]));
            } else {
#pragma unroll
                for (int j = 0; j < DELAY_SPT; ++j) v[j] = dv[j];
            }
            if (VEC) {
#pragma unroll
                for (int q = 0; q < DELAY_SPT / 4; ++q)
                    reinterpret_cast<float4*>(yr + t)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < DELAY_SPT; ++j) yr[t + j] = v[j];
            }
        }
        // the ragged tail (T % DELAY_SPT samples) by the first block
        if (blockIdx.x == 0) {
            const long long t = groups * DELAY_SPT + threadIdx.x;
            if (t < T) yr[t] = warmup ? xr[t] : delay_read(dr[t], t, D, past);
        }
    }
}

// new history = last D samples of (history || x), any T (code/model.py:314-315)
__global__ void __launch_bounds__(256) roll_history_kernel(const float* __restrict__ x, long long ldx,
                                                           const float* __restrict__ hist_in,
                                                           float* __restrict__ hist_out, long long T, int D)
{
    const long long b = blockIdx.y;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < D;
         i += (long long)gridDim.x * blockDim.x) {
        const long long src = T - D + i;
        hist_out[b * D + i] = src >= 0 ? x[b * ldx + src] : hist_in[b * D + D + src];
    }
}

__global__ void __launch_bounds__(256) delay_check_kernel(const float* __restrict__ d, long long ldd, long long T,
                                                          float Dmax, int* flag)
{
    const long long b = blockIdx.y;
    bool bad = false;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < T;
         t += (long long)gridDim.x * blockDim.x)
        bad |= !(d[b * ldd + t] <= Dmax);      // (NaN trips the check like the reference's assert, code/model.py:283)
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

unsigned grid_x(long long n)
{
    long long g = (n + 255) / 256;
    if (g > 2048) g = 2048;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace

cudaError_t launch_delay(const float* x, long long ldx, const float* d, long long ldd, float* y, long long ldy,
                         const float* hist_in, float* hist_out, long long B, long long T, long long D, int warmup,
                         cudaStream_t st)
{
    if (B <= 0) return cudaSuccess;
    for (long long b0 = 0; b0 < B; b0 += 65535) {       // gridDim.y limit
        const long long nb = (B - b0) < 65535 ? (B - b0) : 65535;
        if (T > 0) {
            // warm-up copies x -> y: then x takes d's place in the alignment test
            const float* src = warmup ? x : d;
            const long long lds = warmup ? ldx : ldd;
            const bool vec = (((unsigned long long)src | (unsigned long long)y) & 15ull) == 0 && (lds & 3) == 0 && (ldy & 3) == 0;
            const dim3 grid(grid_x((T + DELAY_SPT - 1) / DELAY_SPT), (unsigned)nb);
            if (vec)
                delay_kernel<true><<<grid, 256, 0, st>>>(x + b0 * ldx, ldx, d + b0 * ldd, ldd, y + b0 * ldy, ldy,
                                                         hist_in + b0 * D, nb, T, (int)D, warmup);
            else
                delay_kernel<false><<<grid, 256, 0, st>>>(x + b0 * ldx, ldx, d + b0 * ldd, ldd, y + b0 * ldy, ldy,
                                                          hist_in + b0 * D, nb, T, (int)D, warmup);
            ++g_launches;
        }
        if (D > 0) {
            roll_history_kernel<<<dim3(grid_x(D), (unsigned)nb), 256, 0, st>>>(x + b0 * ldx, ldx, hist_in + b0 * D,
                                                                                hist_out + b0 * D, T, (int)D);
            ++g_launches;
        }
    }
    return cudaGetLastError();
}

cudaError_t launch_delay_check(const float* d, long long ldd, long long B, long long T, long long D, int* flag_dev,
                               cudaStream_t st)
{
    if (B <= 0 || T <= 0) return cudaSuccess;
    for (long long b0 = 0; b0 < B; b0 += 65535) {
        const long long nb = (B - b0) < 65535 ? (B - b0) : 65535;
        delay_check_kernel<<<dim3(grid_x(T), (unsigned)nb), 256, 0, st>>>(d + b0 * ldd, ldd, T, (float)D, flag_dev);
        ++g_launches;
    }
    return cudaGetLastError();
}

}  // namespace ntm
