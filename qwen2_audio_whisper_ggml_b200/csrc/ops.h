// ops.h -- internal launch wrappers of the sm_100a kernels (device pointers + stream).
// The public boundary is include/q2w_b200.h; these are what the engine (engine.cu) strings together
// and what the kernel-level parity tests reach through the q2w_op_* C-ABI shims.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace q2w {

// sets the calling thread's q2w_last_error() text and returns `code` (engine.cu)
int set_last_error(int code, const char* msg);

// ---- GEMM: out[M,N] = epilogue( A[M,K] (f16, row-major, lda) x W[N,K]^T (f16, row-major, ldw) ), F32 accumulate in TMEM
enum GemmEpilogue : int {
    EPI_BIAS_F16       = 0,  // out f16 = (acc + bias[n]) * (n < scale_cols ? scale : 1)        (fused QKV projection)
    EPI_BIAS_GELU_F16  = 1,  // out f16 = gelu_tanh(acc + bias[n])                              (fc1, conv1)
    EPI_BIAS_RESID_F32 = 2,  // out f32 = resid[m,n] + acc + bias[n]   (resid may alias out)    (out-proj, fc2)
    EPI_BIAS_GELU_POS_F32 = 3,  // out f32 = gelu_tanh(acc + bias[n]) + pos[(m % pos_period), n] (conv2 + positional embedding)
    EPI_BIAS_F32       = 4,  // out f32 = acc + bias[n]                                          (tests / generic)
};

struct GemmArgs {
    const __half* A;  int lda;    // activations [M, K]
    const __half* W;  int ldw;    // weights     [N, K]   (ggml mul_mat layout: ne0 = K contiguous); raw ggml blocks if wtype is Q8_0 / Q4_0
    int wtype;                    // 0 or 1 = F16 (TMA-loaded), 8 = Q8_0, 2 = Q4_0 (decoded inside the kernel; needs K % 64 == 0)
    int M, N, K;
    const float* bias;            // [N] or nullptr (treated as 0)
    void* out;        int ldo;    // [M, N] f16 or f32 depending on epilogue
    const float* resid;           // [M, ldo] f32 (EPI_BIAS_RESID_F32)
    const float* pos; int pos_period;  // [pos_period, N] f32 (EPI_BIAS_GELU_POS_F32)
    int scale_cols;   float scale;     // EPI_BIAS_F16
    int w_static;                 // 1: W is a model weight that no kernel writes (never the dequant scratch): the kernel may fetch its first
                                  // tiles before waiting for the previous kernel in the stream (programmatic dependent launch)
};

// returns cudaSuccess or the failing error; never aborts
cudaError_t gemm_f16_tcgen05(const GemmArgs& a, GemmEpilogue epi, cudaStream_t st);
int gemm_num_launches();  // bookkeeping for bench "gpu_launches"
// split-K of the residual epilogue at small M: 0 off (bit-reproducible single pass), 1 whole k-ranges per tile, 2 balanced unit
// ranges, -1 back to the default (env Q2W_GEMM_SPLITK, else 1)
void gemm_set_splitk_mode(int mode);

// ---- LayerNorm: y f16 [M, D] = (x - mean) * rsqrt(var + eps) * gamma + beta, x f32 [M, D]
cudaError_t layernorm_f32_to_f16(const float* x, const float* gamma, const float* beta, __half* y, int M, int D,
                                 float eps, cudaStream_t st);
// ---- tail: avg-pool(k=2,s=2) over time then LayerNorm, f32 out.  x [B*T, D] -> y [B*(T/2), D]
// y16 (optional): the same rows rounded to F16 -- the A operand of the multi-modal projector GEMM, written by the same pass
cudaError_t pool2_layernorm_f32(const float* x, const float* gamma, const float* beta, float* y, int B, int T, int D,
                                float eps, cudaStream_t st, __half* y16 = nullptr);

// ---- attention (non-causal, no mask, Q pre-scaled): qkv f16 [B*T, 3*D] (q | k | v blocks of D = H*64), out f16 [B*T, D]
// up to four matrices (one encoder block) in ONE launch: src[i] raw ggml blocks (Q8_0 = 8 / Q4_0 = 2), dst[i] f16, nblocks[i] 32-element blocks
struct DequantJob {
    const uint8_t* src[4];
    __half* dst[4];
    unsigned long long nblocks[4];
};
// sched: two ints of device memory, zero before the first launch (the kernel leaves them zero again); one buffer per stream.
// job (optional): quantised weight matrices (ggml_type job_type = 8 / 2, an even number of blocks each) that the kernel's two idle
// warps per CTA decode to F16 while the attention runs -- complete when the kernel is.
cudaError_t attention_f16_tcgen05(const __half* qkv, __half* out, int B, int T, int H, int* sched, cudaStream_t st,
                                  const DequantJob* job = nullptr, int job_type = 0);

// ---- mel front-end
struct MelPlan;  // filterbank + tables resident on device
cudaError_t mel_plan_create(MelPlan** plan, const float* filters_host, int n_mel, int n_fft_bins, cudaStream_t st);
void mel_plan_destroy(MelPlan* plan);
// PCM windows -> un-normalised log10 mel, mel-major [B][n_mel][ld_frames] f32, plus per-window max (float bits, ordered)
//   pcm: B windows, window b at pcm + b*pcm_stride, n_samples[b] valid samples (device int array or nullptr => all = n_max)
//   frames [0, n_frames) are written for every window; frames past the signal get log10(1e-10) = -10 exactly as the reference
cudaError_t mel_logpower(const MelPlan* plan, const float* pcm, size_t pcm_stride, const int* n_samples_dev, int n_max,
                         int B, int n_frames, float* logmel, int ld_frames, float* win_max, cudaStream_t st);
// clamp to (max - 8), (x + 4) / 4 in place; mel-major f32 (the reference's whisper_mel layout)
cudaError_t mel_normalize(float* logmel, int ld_frames, int n_frames, int n_mel, const float* win_max, int B,
                          cudaStream_t st);
// normalise + build the conv1 im2col operand: A1 f16 [B*n_ctx2, 3*n_mel], column = ic*3 + k (ggml im2col order),
// reading frames [offset, offset + n_ctx2) of each window (zero beyond n_frames_valid: src/qwen2-whisper.cpp:2274-2283)
cudaError_t mel_to_conv1_operand(const float* mel, int ld_frames, int n_frames_valid, int n_mel, const float* win_max,
                                 int normalise, int offset, int n_ctx2, int B, __half* A1, cudaStream_t st);
// conv2 im2col: h1 f16 [B*T2, C] (time-major) -> A2 f16 [B*(T2/2), 3*C], column = ic*3 + k, stride 2, pad 1
cudaError_t conv2_im2col(const __half* h1, __half* A2, int B, int T2, int C, cudaStream_t st);

// ---- ggml block decode: Q8_0 / Q4_0 / F32 rows -> f16 [rows, K] (K % 32 == 0)
cudaError_t dequant_to_f16(const void* src, int ggml_type, __half* dst, size_t rows, int K, cudaStream_t st);
cudaError_t dequant_multi_to_f16(const DequantJob& job, int ggml_type, cudaStream_t st);

}  // namespace q2w
