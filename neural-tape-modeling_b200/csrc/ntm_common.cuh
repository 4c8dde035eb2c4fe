// Shared declarations of the engine's translation units (internal; the public ABI is include/ntm_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ntm {

constexpr int H64 = 64;      // hidden size all shipped checkpoints use
constexpr int G192 = 192;    // 3 gates x 64, row order r,z,n (torch rnn.py:1221-1224)

// Device parameter blob (floats), written once by ntm_gru_prepare.
struct BlobLayout {
    static constexpr int W_HH = 0;                    // 192 x 64, PyTorch layout
    static constexpr int W_IH = W_HH + G192 * H64;    // 192
    static constexpr int B_IH = W_IH + G192;          // 192
    static constexpr int B_HH = B_IH + G192;          // 192
    static constexpr int W_OUT = B_HH + G192;         // 64
    static constexpr int B_OUT = W_OUT + H64;         // 1 (0 when the head has no bias)
    static constexpr int FP32_END = B_OUT + 4;
    // tensor-core A-operand images (gru_tc.cu): 3 gate tiles x 128 rows x 80 k (64 + the input/bias augmentation) of
    // 2-byte elements, row-major; one image per operand format
    static constexpr int IMG_F16 = FP32_END;
    static constexpr int IMG_BF16 = IMG_F16 + 3 * 128 * 80 / 2;
    static constexpr int TOTAL = IMG_BF16 + 3 * 128 * 80 / 2;
    __host__ __device__ static constexpr int tc_image(int fmt) { return fmt == 1 ? IMG_BF16 : IMG_F16; }
};
static_assert(BlobLayout::FP32_END % 4 == 0, "tensor-core images must stay 16-byte aligned");

// Mailbox of a real-time stream (ntm_rt_*): page-locked, mapped host memory shared by the host thread and the resident
// server kernel (gru_mma.cu, RT form).  The host writes a block into x, then bumps seq_in; the kernel writes y, then sets
// seq_out = seq_in.  seq_in == RT_STOP ends the kernel.
constexpr int RT_MAXBLK = 256;          // samples per block, at most
constexpr int RT_MAXSTREAMS = 4;
constexpr unsigned RT_STOP = 0xffffffffu;
struct RtMailbox {
    volatile unsigned seq_in;
    unsigned pad0[31];
    float x[RT_MAXSTREAMS][RT_MAXBLK];
    volatile unsigned seq_out;
    unsigned pad1[31];
    float y[RT_MAXSTREAMS][RT_MAXBLK];
};

// Arguments of one recurrent launch; d == nullptr selects the plain RNN.forward path.
struct GruArgs {
    const float* blob;
    const float* x;
    float* y;
    const float* h_in;      // B x 64 or nullptr (zero state)
    float* h_out;           // B x 64 (may alias h_in)
    const float* d;         // delay trajectory (samples) or nullptr
    float* pre;             // GRU head output before the delay (DiffDelRNN only)
    const float* hist_in;   // B x D
    float* hist_out;        // B x D
    long long B, T, ldx, ldy, ldd, ldp;
    int D;
    int warmup;
    int skip;
    int ring_len;                   // mma.sync kernel, DiffDelRNN: samples of pre_d kept per stream in shared memory
                                    // (power of two >= D + chunk; 0: read the delay taps back through L2)
    RtMailbox* rt;                  // real-time server form only (x, y point into it)
    unsigned long long rt_idle_ns;  // the server leaves after this long without a block
};

// Warp-uniform per-unit constants of the stream-major tcgen05 kernel (gru_tcs.cu); passed as a kernel parameter so
// that they are constant-bank operands.  Scaled like the n-gate rows (2 log2 e).
struct TcsConsts {
    float cn_w[64];   // W_in
    float cn_b[64];   // b_in
    float wo[64];     // output head
    float bo;
};

// per-TU launchers -------------------------------------------------------------------------------
cudaError_t launch_gru_fp32(const GruArgs& a, int sm_count, int tune_s, int tune_ks, int fast_act, cudaStream_t st);
cudaError_t launch_gru_tc(const GruArgs& a, int fmt, int sm_count, int tune_n, int tune_g, cudaStream_t st);
cudaError_t launch_gru_mma8(const GruArgs& a, int fmt, cudaStream_t st);     // 8 warps x 8 streams per CTA (gru_mma8.cu)
cudaError_t launch_gru_tcs(const GruArgs& a, const TcsConsts& kc, int fmt, int sm_count, int tiles, int var, cudaStream_t st);
void fill_tcs_consts(const float* blob_host, TcsConsts* kc);
cudaError_t launch_gru_mma(const GruArgs& a, int fmt, int n_tiles, cudaStream_t st);
cudaError_t launch_gru_mma_rt(const GruArgs& a, int fmt, cudaStream_t st);    // resident real-time server, <= 4 streams
void pack_tc_images(float* blob_host);   // host: fills BlobLayout::IMG_* from the fp32 part of the blob
cudaError_t launch_delay(const float* x, long long ldx, const float* d, long long ldd, float* y, long long ldy,
                         const float* hist_in, float* hist_out, long long B, long long T, long long D, int warmup,
                         cudaStream_t st);
cudaError_t launch_delay_check(const float* d, long long ldd, long long B, long long T, long long D, int* flag_dev,
                               cudaStream_t st);

cudaError_t launch_esr(const float* out, long long ldo, const float* tgt, long long ldt, long long B, long long T,
                       int dc_pre, double* sums, int sm_count, cudaStream_t st);

extern unsigned long long g_launches;   // engine kernels launched by this process (ntm_query)

// ------------------------------------------------------------------------------------------------
// Fractional delay read shared by the fused and the stand-alone kernels.
// Reference: TimeVaryingDelayLine.forward, code/model.py:294-311 -- weights relu(1 - |j - d|) over the taps
// j = 0..D; only j = floor(d) and floor(d)+1 can be non-zero.  Weights and products are rounded exactly as
// the reference rounds them (float sub/abs/sub, float mul, one float add) so the result is bit-identical.
template <class LoadPast>
__device__ __forceinline__ float delay_read(float dt, long long t, int D, LoadPast past)
{
    const float fl = floorf(dt);
    float acc = 0.0f;
#pragma unroll
    for (int tap = 1; tap >= 0; --tap) {
        const float jf = fl + (float)tap;
        if (jf >= 0.0f && jf <= (float)D) {
            const float w = fmaxf(__fsub_rn(1.0f, fabsf(__fsub_rn(jf, dt))), 0.0f);
            acc = __fadd_rn(acc, __fmul_rn(w, past(t - (long long)jf)));
        }
    }
    return acc;
}

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

}  // namespace ntm
