// Dev tool (GPU box): error statistics of MUFU ex2.approx.ftz / rcp.approx.ftz against double precision -- is the error of the
// approximate activations a BIAS (which the GRU recurrence integrates) or noise?   nvcc -arch=sm_100a -o mufu_probe mufu_probe.cu
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// per block: sum, sum of squares, max |rel error| in ulp (2^-23) for the two functions over a slice of arguments
__global__ void probe(double lo, double hi, int n, double* out)
{
    double s[2] = {0, 0}, q[2] = {0, 0}, m[2] = {0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = (float)(lo + (hi - lo) * ((double)i + 0.5) / n);
        const double e = ((double)ex2_approx(x) - exp2((double)x)) / exp2((double)x) * 8388608.0;
        const float d = 1.0f + ex2_approx(x);
        const double r = ((double)rcp_approx(d) - 1.0 / (double)d) * (double)d * 8388608.0;
        s[0] += e; q[0] += e * e; m[0] = fmax(m[0], fabs(e));
        s[1] += r; q[1] += r * r; m[1] = fmax(m[1], fabs(r));
    }
    for (int k = 0; k < 2; ++k) {
        atomicAdd(out + 3 * k, s[k]);
        atomicAdd(out + 3 * k + 1, q[k]);
        // max via atomicMax on the bit pattern of a non-negative double
        atomicMax(reinterpret_cast<unsigned long long*>(out + 3 * k + 2), (unsigned long long)__double_as_longlong(m[k]));
    }
}

int main()
{
    double* d;
    cudaMalloc(&d, 6 * sizeof(double));
    const int n = 1 << 24;
    const double ranges[][2] = {{-30, 30}, {-8, 8}, {-1, 1}, {-0.5, 0.5}, {0, 0.125}, {4, 4.125}, {-20, -10}};
    for (auto& r : ranges) {
        cudaMemset(d, 0, 6 * sizeof(double));
        probe<<<592, 256>>>(r[0], r[1], n, d);
        double h[6];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("x in [%g, %g]: ex2.approx mean %+.4f ulp rms %.4f max %.3f | rcp.approx(1+2^x) mean %+.4f ulp rms %.4f max %.3f\n",
               r[0], r[1], h[0] / n, sqrt(h[1] / n), h[2], h[3] / n, sqrt(h[4] / n), h[5]);
    }
    return 0;
}
