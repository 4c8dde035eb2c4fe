// Stand-alone time-varying fractional delay line (HBM-bound elementwise gather).
// Replaces TimeVaryingDelayLine.forward, code/model.py:269-320, as used on its own by apply_delay
// (code/test-model.py:259-290).  The reference materialises three (B,1,T,D+1) temporaries; here every
// output sample reads two past samples (history or this call's input) -- 12 algorithmic bytes per sample.
#include "ntm_common.cuh"

namespace ntm {

namespace {

__global__ void __launch_bounds__(256) delay_kernel(const float* __restrict__ x, long long ldx,
                                                    const float* __restrict__ d, long long ldd,
                                                    float* __restrict__ y, long long ldy,
                                                    const float* __restrict__ hist_in, long long B, long long T,
                                                    int D, int warmup)
{
    const long long b = blockIdx.y;
    const float* xr = x + b * ldx;
    const float* hr = hist_in + b * (long long)D;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < T;
         t += (long long)gridDim.x * blockDim.x) {
        float v;
        if (warmup) {
            v = xr[t];
        } else {
            v = delay_read(d[b * ldd + t], t, D, [&](long long i) { return i >= 0 ? xr[i] : hr[D + i]; });
        }
        y[b * ldy + t] = v;
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
        bad |= d[b * ldd + t] > Dmax;
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
            delay_kernel<<<dim3(grid_x(T), (unsigned)nb), 256, 0, st>>>(x + b0 * ldx, ldx, d + b0 * ldd, ldd,
                                                                         y + b0 * ldy, ldy, hist_in + b0 * D, nb, T,
                                                                         (int)D, warmup);
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
