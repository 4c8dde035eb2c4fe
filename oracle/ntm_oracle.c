/*
 * ntm_oracle.c -- CPU restatement of the neural-tape-modeling recurrent forward pass.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (neural-tape-modeling_b200/) may
 * link, import or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * What it restates (reference = /root/reference, third-party arithmetic = torch 2.11):
 *   - GRU cell, gate order r,z,n:            torch/nn/modules/rnn.py:1221-1224 (equations),
 *                                            called from code/model.py:81 and :412
 *   - Linear head (+ optional input skip):   code/model.py:82-84 (bias) and :413-415 (no bias)
 *   - f64 -> f32 down-cast of the input:     code/model.py:76, :403   (done by the caller)
 *   - TimeVaryingDelayLine.forward:          code/model.py:269-320
 *
 * Parity pinning: the reference holds no golden vectors for this path,
 * so this restatement is pinned against outputs of the reference itself, generated in the build
 * container by oracle/make_golden.py (imports /root/reference/code/model.py) and committed under
 * tests/golden/.  tests/test_oracle.py checks every function here against those fixtures.
 *
 * Two precisions: *_f32 follows ATen's CPU op order in float; *_f64 carries the state and
 * all arithmetic in double (ground truth used to measure the reference's own fp32 noise floor).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NTM_ORACLE_EXPORT __attribute__((visibility("default")))

static inline float sigmoid_f32(float a) { return 1.0f / (1.0f + expf(-a)); }
static inline double sigmoid_f64(double a) { return 1.0 / (1.0 + exp(-a)); }

/*
 * One GRU layer (input size 1, hidden size H) followed by the 1-output Linear head.
 *   x, y : B rows of T samples, row strides ldx / ldy   (the (B,1,T) tensors of code/model.py:67-88)
 *   h    : B x H hidden state, read as h_{-1}, overwritten with h_{T-1}  (self.hidden, code/model.py:81)
 *   b_out may be NULL (DiffDelRNN head has bias=False, code/model.py:365)
 * ATen CPU op order (aten/src/ATen/native/RNN.cpp GRUCell): gi = W_ih x + b_ih ; gh = W_hh h + b_hh ;
 *   r = sig(gi_r + gh_r) ; z = sig(gi_z + gh_z) ; n = tanh(gi_n + r * gh_n) ; h' = (h - n) * z + n.
 */
NTM_ORACLE_EXPORT
int ntm_oracle_gru_f32(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                       const float* w_out, const float* b_out, int H,
                       const float* x, float* y, float* h,
                       int64_t B, int64_t T, int64_t ldx, int64_t ldy, int skip)
{
    if (H <= 0 || B < 0 || T < 0) return -1;
    float* gh = (float*)malloc(sizeof(float) * 3 * (size_t)H);
    float* hn = (float*)malloc(sizeof(float) * (size_t)H);
    if (!gh || !hn) { free(gh); free(hn); return -2; }
    for (int64_t b = 0; b < B; ++b) {
        float* hb = h + b * H;
        for (int64_t t = 0; t < T; ++t) {
            const float xt = x[b * ldx + t];
            for (int g = 0; g < 3 * H; ++g) {
                float acc = 0.0f;
                const float* wr = w_hh + (size_t)g * H;
                for (int k = 0; k < H; ++k) acc += wr[k] * hb[k];
                gh[g] = acc + b_hh[g];
            }
            for (int j = 0; j < H; ++j) {
                const float gi_r = w_ih[j] * xt + b_ih[j];
                const float gi_z = w_ih[H + j] * xt + b_ih[H + j];
                const float gi_n = w_ih[2 * H + j] * xt + b_ih[2 * H + j];
                const float r = sigmoid_f32(gi_r + gh[j]);
                const float z = sigmoid_f32(gi_z + gh[H + j]);
                const float n = tanhf(gi_n + r * gh[2 * H + j]);
                hn[j] = (hb[j] - n) * z + n;
            }
            memcpy(hb, hn, sizeof(float) * (size_t)H);
            float acc = 0.0f;
            for (int j = 0; j < H; ++j) acc += w_out[j] * hb[j];
            if (b_out) acc += b_out[0];
            if (skip) acc += xt;
            y[b * ldy + t] = acc;
        }
    }
    free(gh); free(hn);
    return 0;
}

/* Same recurrence with double state/arithmetic; x is the (already down-cast) float input, y/h are double. */
NTM_ORACLE_EXPORT
int ntm_oracle_gru_f64(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                       const float* w_out, const float* b_out, int H,
                       const float* x, double* y, double* h,
                       int64_t B, int64_t T, int64_t ldx, int64_t ldy, int skip)
{
    if (H <= 0 || B < 0 || T < 0) return -1;
    double* gh = (double*)malloc(sizeof(double) * 3 * (size_t)H);
    double* hn = (double*)malloc(sizeof(double) * (size_t)H);
    if (!gh || !hn) { free(gh); free(hn); return -2; }
    for (int64_t b = 0; b < B; ++b) {
        double* hb = h + b * H;
        for (int64_t t = 0; t < T; ++t) {
            const double xt = (double)x[b * ldx + t];
            for (int g = 0; g < 3 * H; ++g) {
                double acc = 0.0;
                const float* wr = w_hh + (size_t)g * H;
                for (int k = 0; k < H; ++k) acc += (double)wr[k] * hb[k];
                gh[g] = acc + (double)b_hh[g];
            }
            for (int j = 0; j < H; ++j) {
                const double gi_r = (double)w_ih[j] * xt + (double)b_ih[j];
                const double gi_z = (double)w_ih[H + j] * xt + (double)b_ih[H + j];
                const double gi_n = (double)w_ih[2 * H + j] * xt + (double)b_ih[2 * H + j];
                const double r = sigmoid_f64(gi_r + gh[j]);
                const double z = sigmoid_f64(gi_z + gh[H + j]);
                const double n = tanh(gi_n + r * gh[2 * H + j]);
                hn[j] = (1.0 - z) * n + z * hb[j];
            }
            memcpy(hb, hn, sizeof(double) * (size_t)H);
            double acc = 0.0;
            for (int j = 0; j < H; ++j) acc += (double)w_out[j] * hb[j];
            if (b_out) acc += (double)b_out[0];
            if (skip) acc += xt;
            y[b * ldy + t] = acc;
        }
    }
    free(gh); free(hn);
    return 0;
}

/* p[i] for i in [-D, T): carried history (i < 0) followed by this call's input (code/model.py:286). */
static inline float padded_at(const float* hist, const float* xrow, int64_t D, int64_t i)
{
    return i < 0 ? hist[D + i] : xrow[i];
}

/* new history = last D samples of (history || x), any T (code/model.py:314-315). */
static void roll_history(const float* hist_in, float* hist_out, const float* xrow, int64_t D, int64_t T)
{
    float* tmp = (float*)malloc(sizeof(float) * (size_t)(D > 0 ? D : 1));
    for (int64_t i = 0; i < D; ++i) tmp[i] = padded_at(hist_in, xrow, D, T - D + i);
    memcpy(hist_out, tmp, sizeof(float) * (size_t)D);
    free(tmp);
}

/*
 * TimeVaryingDelayLine.forward, literal form (code/model.py:294-311): for every output sample the
 * window of D+1 past samples is weighted by relu(1 - |j - d|), j = delay in samples of each tap,
 * products rounded to float, then summed.  O(T*D): small cases only.
 *   x, d, y : B rows (strides ldx / ldd / ldy);  hist_in/hist_out : B x D.
 * Returns -3 if max(d) > D (the reference's assert, code/model.py:283).
 */
NTM_ORACLE_EXPORT
int ntm_oracle_delay_window_f32(const float* x, const float* d, float* y,
                                const float* hist_in, float* hist_out,
                                int64_t B, int64_t T, int64_t D,
                                int64_t ldx, int64_t ldd, int64_t ldy, int warmup)
{
    for (int64_t b = 0; b < B; ++b)
        for (int64_t t = 0; t < T; ++t)
            if (d[b * ldd + t] > (float)D) return -3;
    for (int64_t b = 0; b < B; ++b) {
        const float* xr = x + b * ldx;
        const float* hi = hist_in + b * D;
        if (warmup) {
            for (int64_t t = 0; t < T; ++t) y[b * ldy + t] = xr[t];
        } else {
            for (int64_t t = 0; t < T; ++t) {
                const float dt = d[b * ldd + t];
                /* torch.sum over the window: window index i = 0..D holds tap delay j = D - i. */
                float acc = 0.0f;
                for (int64_t i = 0; i <= D; ++i) {
                    const float j = (float)(D - i);
                    float w = 1.0f - fabsf(j - dt);
                    w = w > 0.0f ? w : 0.0f;
                    acc += w * padded_at(hi, xr, D, t - (D - i));
                }
                y[b * ldy + t] = acc;
            }
        }
        roll_history(hi, hist_out + b * D, xr, D, T);
    }
    return 0;
}

/*
 * Same operator in its two-tap form: only taps j = floor(d) and floor(d)+1 can have non-zero weight,
 * the weights are still evaluated as float(1 - |j - d|) so the result is bit-identical to the window
 * form whenever the window sum adds the two products in one rounding (all other terms are +-0).
 */
NTM_ORACLE_EXPORT
int ntm_oracle_delay_f32(const float* x, const float* d, float* y,
                         const float* hist_in, float* hist_out,
                         int64_t B, int64_t T, int64_t D,
                         int64_t ldx, int64_t ldd, int64_t ldy, int warmup)
{
    for (int64_t b = 0; b < B; ++b)
        for (int64_t t = 0; t < T; ++t)
            if (d[b * ldd + t] > (float)D) return -3;
    for (int64_t b = 0; b < B; ++b) {
        const float* xr = x + b * ldx;
        const float* hi = hist_in + b * D;
        if (warmup) {
            for (int64_t t = 0; t < T; ++t) y[b * ldy + t] = xr[t];
        } else {
            for (int64_t t = 0; t < T; ++t) {
                const float dt = d[b * ldd + t];
                const float fl = floorf(dt);
                float acc = 0.0f;
                for (int tap = 1; tap >= 0; --tap) {      /* larger delay first = window order */
                    const float jf = fl + (float)tap;
                    if (jf < 0.0f || jf > (float)D) continue;
                    float w = 1.0f - fabsf(jf - dt);
                    w = w > 0.0f ? w : 0.0f;
                    acc += w * padded_at(hi, xr, D, t - (int64_t)jf);
                }
                y[b * ldy + t] = acc;
            }
        }
        roll_history(hi, hist_out + b * D, xr, D, T);
    }
    return 0;
}

/* Error-to-signal ratio, CoreAudioML/training.py:10-16: mean((t-o)^2) / (mean(t^2) + 1e-5), in double. */
NTM_ORACLE_EXPORT
double ntm_oracle_esr(const float* out, const float* target, int64_t n)
{
    double num = 0.0, den = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const double e = (double)target[i] - (double)out[i];
        num += e * e;
        den += (double)target[i] * (double)target[i];
    }
    if (n > 0) { num /= (double)n; den /= (double)n; }
    return num / (den + 1e-5);
}

/*
 * DC pre-emphasised error-to-signal ratio ("DCPreESR"), code/GreyBoxDRC/loss_funcs.py:6-52 as called by
 * code/test-model.py:250-253,386-388 on (B,1,T) tensors.
 *   DC_PreEmph (:6-30): both signals are zero-padded by 1999 samples at the front and correlated with the flipped
 *   2000-sample impulse response of H(z) = (1 - z^-1) / (1 - R z^-1), R = 0.995 (scipy.signal.dimpulse, cast to
 *   float32), i.e. a causal 2000-tap FIR:  f[n] = sum_{k=0..1999} h[k] x[n-k],  h[0] = 1, h[k] = (R-1) R^(k-1).
 *   ESRLoss (:32-52): mean((f_t - f_o)^2) / (mean(f_t^2) + 1e-5), means over ALL B*T elements.
 * `taps` holds h[0..ntaps-1] in float32 exactly as the reference builds them (tests pass the scipy values);
 * products and sums are carried in double (the reference's fp32 conv1d differs from this by its own round-off).
 * ntaps == 0: no filter (CoreAudioML/training.py:10-16).  Returns the loss; sums[0..1] (optional) get the two sums.
 */
NTM_ORACLE_EXPORT
double ntm_oracle_dcpre_esr(const float* out, const float* target, int64_t B, int64_t T, int64_t ldo, int64_t ldt,
                            const float* taps, int64_t ntaps, double* sums)
{
    double num = 0.0, den = 0.0;
    for (int64_t b = 0; b < B; ++b) {
        const float* o = out + b * ldo;
        const float* t = target + b * ldt;
        for (int64_t n = 0; n < T; ++n) {
            double ft, fe;
            if (ntaps == 0) {
                ft = (double)t[n];
                fe = (double)t[n] - (double)o[n];
            } else {
                ft = 0.0;
                double fo = 0.0;
                const int64_t kmax = n < ntaps - 1 ? n : ntaps - 1;
                for (int64_t k = kmax; k >= 0; --k) {
                    ft += (double)taps[k] * (double)t[n - k];
                    fo += (double)taps[k] * (double)o[n - k];
                }
                fe = ft - fo;
            }
            num += fe * fe;
            den += ft * ft;
        }
    }
    if (sums) { sums[0] = num; sums[1] = den; }
    const double cnt = (double)(B * T);
    return cnt > 0 ? (num / cnt) / (den / cnt + 1e-5) : 0.0;
}
