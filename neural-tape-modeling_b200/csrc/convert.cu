// binary16 <-> float32 conversion passes of the 16-bit host transport (ntm_gru_predict_host_f16): the samples cross the host
// link as IEEE binary16 and are widened / narrowed on the device next to the recurrent kernel.  HBM-bound, 6 bytes per sample
// each -- ~1 % of the recurrent kernel's time at cfg 2.
#include <cuda_fp16.h>

#include "ntm_common.cuh"

namespace ntm {

namespace {

__global__ void __launch_bounds__(256) half_to_float_kernel(const __half* __restrict__ src, float* __restrict__ dst, long long n8)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const uint4 v = reinterpret_cast<const uint4*>(src)[i];
        const __half2* h = reinterpret_cast<const __half2*>(&v);
        const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), c = __half22float2(h[2]), d = __half22float2(h[3]);
        reinterpret_cast<float4*>(dst)[2 * i] = make_float4(a.x, a.y, b.x, b.y);
        reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
    }
}

__global__ void __launch_bounds__(256) float_to_half_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n8)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
        uint4 v;
        __half2* h = reinterpret_cast<__half2*>(&v);
        h[0] = __floats2half2_rn(a.x, a.y); h[1] = __floats2half2_rn(a.z, a.w);
        h[2] = __floats2half2_rn(b.x, b.y); h[3] = __floats2half2_rn(b.z, b.w);
        reinterpret_cast<uint4*>(dst)[i] = v;
    }
}

}  // namespace

// n: elements, a multiple of 8; both buffers 16-byte aligned (the staging slabs of predict_host are)
cudaError_t launch_half_to_float(const void* src, float* dst, long long n, int sm_count, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const long long n8 = n / 8;
    const int grid = (int)((n8 + 255) / 256 < (long long)sm_count * 8 ? (n8 + 255) / 256 : (long long)sm_count * 8);
    half_to_float_kernel<<<grid, 256, 0, st>>>(static_cast<const __half*>(src), dst, n8);
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_float_to_half(const float* src, void* dst, long long n, int sm_count, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const long long n8 = n / 8;
    const int grid = (int)((n8 + 255) / 256 < (long long)sm_count * 8 ? (n8 + 255) / 256 : (long long)sm_count * 8);
    float_to_half_kernel<<<grid, 256, 0, st>>>(src, static_cast<__half*>(dst), n8);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace ntm
