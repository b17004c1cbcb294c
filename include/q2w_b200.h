/* q2w_b200.h -- C ABI of the B200 (sm_100a) audio front-end library  libq2w_b200.so
 *
 * This is the seam SURVEY.md section 1 / section 8(b) cuts into the reference: the host code of
 * src/qwen2-whisper.cpp keeps its public API (include/qwen2-whisper.h) but calls THIS library at the
 * points where it used to go through ggml's backend scheduler.  Each entry point names the reference
 * interface it replaces (file:line relative to /root/reference).
 *
 * Conventions: plain C, opaque handles, plain pointers and sizes, no C++ types, no exceptions.
 * int functions return 0 on success and a negative Q2W_E_* code on failure (the reference's
 * "0 ok / negative on failure" convention, src/qwen2-whisper.cpp:2341-2375); q2w_last_error() returns
 * a thread-local diagnostic string.  CUDA errors become Q2W_E_CUDA; the library never abort()s.
 * There is no CPU fallback: every compute entry point requires a CUDA device of compute capability 10.x.
 * A handle is single-caller (the reference's rule, include/qwen2-whisper.h:44-45); several states may
 * share one read-only model (src/qwen2-whisper.cpp:769-770).
 */
#ifndef Q2W_B200_H
#define Q2W_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Q2W_API __attribute__((visibility("default")))

enum {
    Q2W_OK            =  0,
    Q2W_E_INVALID     = -1,  /* bad argument / shape / state */
    Q2W_E_CUDA        = -2,  /* CUDA runtime or driver error (message in q2w_last_error) */
    Q2W_E_NO_DEVICE   = -3,  /* no sm_100 device: this library has no fallback path */
    Q2W_E_UNKNOWN_TENSOR = -4,   /* src/qwen2-whisper.cpp:1807 */
    Q2W_E_BAD_SHAPE   = -5,  /* :1814, :1821 */
    Q2W_E_BAD_SIZE    = -6,  /* :1829 */
    Q2W_E_INCOMPLETE  = -7,  /* :1861 not all tensors loaded */
    Q2W_E_UNSUPPORTED = -8,  /* ggml type / hparams outside the path (e.g. head_dim != 64) */
    Q2W_E_NOMEM       = -9
};

/* ggml_type values accepted for weights (ggml/include/ggml.h enum ggml_type) */
enum { Q2W_TYPE_F32 = 0, Q2W_TYPE_F16 = 1, Q2W_TYPE_Q4_0 = 2, Q2W_TYPE_Q8_0 = 8 };

/* The 11 int32 header fields of the model file, in file order (src/qwen2-whisper.cpp:1374-1384). */
typedef struct q2w_hparams {
    int32_t n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
    int32_t n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, ftype;
} q2w_hparams;

typedef struct q2w_model q2w_model;   /* weights on one device; replaces whisper_model + its backend buffer (:732-775, :1765) */
typedef struct q2w_state q2w_state;   /* workspaces, mel, embeddings, timers; replaces whisper_state (:795-864, :2779) */

/* ---- model -------------------------------------------------------------------------------------- */
/* wtype: ggml_type of the 2-D weight matrices (F32/F16/Q8_0/Q4_0), from ftype (:1414-1424).  device: CUDA ordinal. */
Q2W_API int  q2w_model_create(q2w_model** out, const q2w_hparams* hp, int wtype, int device);
/* mel filterbank float[n_mel][n_fft] (n_fft = 201), replaces whisper_filters load (:1442-1451) */
Q2W_API int  q2w_model_upload_filters(q2w_model* m, const float* filters, int n_mel, int n_fft);
/* one tensor of the file's tensor stream in ggml byte layout; replaces ggml_backend_tensor_set (:1839-1850).
 * Validates name / element count / shape / byte size exactly like :1807-1833. ne[] is ggml order (innermost first). */
Q2W_API int  q2w_model_upload_tensor(q2w_model* m, const char* name, int ggml_type, int n_dims, const int32_t* ne,
                                     const void* data, size_t nbytes);
/* bytes the file must carry for tensor `name` in the model's weight type (0 = unknown name): validate a record BEFORE reading its payload */
Q2W_API size_t q2w_model_tensor_bytes(const q2w_model* m, const char* name);
/* all 7 + 15*L tensors present? (:1861)  Builds the fused QKV weight views. */
Q2W_API int  q2w_model_finalize(q2w_model* m);
Q2W_API void q2w_model_free(q2w_model* m);
Q2W_API int  q2w_model_device(const q2w_model* m);
Q2W_API int  q2w_model_n_tensors_expected(const q2w_model* m);
Q2W_API int  q2w_model_n_tensors_loaded(const q2w_model* m);
Q2W_API size_t q2w_model_weight_bytes(const q2w_model* m);

/* ---- state -------------------------------------------------------------------------------------- */
/* max_batch: windows processed per micro-batch (workspace is sized for it; larger batches are chunked). */
Q2W_API int  q2w_state_create(q2w_state** out, q2w_model* m, int max_batch);
Q2W_API void q2w_state_free(q2w_state* s);
/* resize the per-batch scratch in place: the state's mel, embeddings, timers and streams survive (a state is otherwise what
 * whisper_init_state builds once, :2779) */
Q2W_API int  q2w_state_set_max_batch(q2w_state* s, int max_batch);
Q2W_API int  q2w_state_max_batch(const q2w_state* s);

/* PCM (host, float [-1,1], 16 kHz) -> log-mel kept on the device in the reference's layout float[n_mel][n_len],
 * n_len = (n_samples + 480000) / 160.  Replaces log_mel_spectrogram via whisper_pcm_to_mel_with_state (:3268, :2575). */
Q2W_API int  q2w_pcm_to_mel(q2w_state* s, const float* pcm_host, int n_samples);
/* caller-provided mel, float[n_mel][n_len] (whisper_set_mel_with_state :3281-3300) */
Q2W_API int  q2w_set_mel(q2w_state* s, const float* mel_host, int n_len, int n_mel);
Q2W_API int  q2w_mel_n_len(const q2w_state* s);       /* whisper_n_len_from_state */
Q2W_API int  q2w_mel_n_len_org(const q2w_state* s);
Q2W_API int  q2w_get_mel(q2w_state* s, float* out_host, size_t n_floats);   /* additive: mel accessor */

/* conv stem + encoder on frames [mel_offset, mel_offset + 2*n_audio_ctx) of the state's mel, zero-filled past n_len.
 * Replaces whisper_encode_qwen2_internal (:2241-2339).  Result: embeddings float[n_audio_ctx/2][n_audio_state] on device. */
Q2W_API int  q2w_encode(q2w_state* s, int mel_offset);

/* Batched hot path (additive API): B independent windows of PCM, window b = pcm + b*stride floats with
 * n_samples[b] <= 2*n_audio_ctx*160 valid samples (n_samples == NULL: all windows full).  Each window gets its own
 * mel normalisation (chunk-then-mel semantics).  Output: B x [n_audio_ctx/2][n_audio_state] f32.
 * _host: pcm in host memory (pinned for full overlap), embeddings copied to out_host if non-NULL.
 * _device: pcm already resident in HBM; results stay on the device (q2w_embeddings_device). */
Q2W_API int  q2w_encode_batch_host(q2w_state* s, const float* pcm_host, size_t stride, const int32_t* n_samples, int B,
                                   float* out_host);
Q2W_API int  q2w_encode_batch_device(q2w_state* s, const float* pcm_dev, size_t stride, const int32_t* n_samples_host,
                                     int B);

/* Asynchronous host batches: queue a batch and return; at most two are in flight per state (a third submit first waits for the oldest).
 * pcm_host / out_host must stay valid (and should be pinned) until q2w_encode_batch_wait(ticket) returns. Lets a caller overlap the
 * H2D / D2H copies of one batch with the compute of the next -- what a serving loop over whisper_encode_batch would do. */
Q2W_API int  q2w_encode_batch_host_async(q2w_state* s, const float* pcm_host, size_t stride, const int32_t* n_samples, int B, float* out_host,
                                         int* ticket);
Q2W_API int  q2w_encode_batch_wait(q2w_state* s, int ticket);

/* Whole-file streaming: n windows of the state's (globally normalised) mel, starting at the given frame offsets,
 * encoded as one batch -- the batched form of n x whisper_full(ctx, {offset_ms}, NULL, 0)  (:2349-2369). */
Q2W_API int  q2w_encode_offsets(q2w_state* s, const int32_t* mel_offsets, int n, float* out_host);

/* embeddings of the last encode / encode_batch (replaces the D2H in whisper_print_emb_enc :4196) */
Q2W_API int  q2w_embd_dims(const q2w_state* s, int* n_windows, int* n_out, int* n_state);
Q2W_API int  q2w_get_embeddings(q2w_state* s, float* out_host, size_t offset_floats, size_t n_floats);
Q2W_API const float* q2w_embeddings_device(const q2w_state* s);
/* per-window mel of the last encode_batch (un-normalised log10 power, float[B][n_mel][ld]) -- debugging / parity */
Q2W_API int  q2w_get_batch_mel(q2w_state* s, int window, float* out_host /* [n_mel][2*n_audio_ctx] normalised */);

/* timers with the reference's meaning (t_mel_us, t_encode_us, n_encode: :796-809, :2651, :2335) */
Q2W_API void q2w_get_timings(const q2w_state* s, int64_t* t_mel_us, int64_t* t_encode_us, int32_t* n_encode);
Q2W_API void q2w_reset_timings(q2w_state* s);
/* per-kernel-class device timing with CUDA events recorded on the state's stream (bench.py's live roofline).
 * classes: 0 weight GEMMs (tcgen05), 1 attention, 2 LayerNorm (+pool tail), 3 mel, 4 im2col/operand builders, 5 ggml block decode.
 * total_flops / total_bytes are the ALGORITHMIC figures of DESIGN.md for the launches recorded since enable(1). */
Q2W_API int  q2w_profile_enable(q2w_state* s, int on);
Q2W_API int  q2w_profile_read(q2w_state* s, int kernel_class, double* total_ms, long* count, double* total_flops, double* total_bytes);
/* the stream all work of this state is enqueued on (cudaStream_t), for callers that time with CUDA events */
Q2W_API void* q2w_state_stream(const q2w_state* s);
Q2W_API int  q2w_sync(q2w_state* s);

/* ---- the step after the path, additive and optional -------------------------------- */
/* Qwen2-Audio's multi_modal_projector: Linear(n_audio_state -> n_out) + bias on every embedding row (HF
 * Qwen2AudioMultiModalProjector.linear; the reference stops at the final LayerNorm, src/qwen2-whisper.cpp:2175-2185).
 * W: [n_out][n_audio_state] row-major, F32 or F16 (rounded to F16 once); bias float[n_out] or NULL.  Once uploaded, the pool +
 * final-LayerNorm kernel also writes its rows in F16 (same pass) and q2w_project runs one tcgen05 GEMM over the embeddings of
 * the last encode / encode_batch: float[n_windows * n_audio_ctx/2][n_out], kept on the device, copied to out_host if non-NULL. */
Q2W_API int  q2w_model_upload_projector(q2w_model* m, int ggml_type, int n_out, const void* w_host, size_t nbytes, const float* bias_host);
Q2W_API int  q2w_model_projector_width(const q2w_model* m);
Q2W_API int  q2w_project(q2w_state* s, float* out_host, size_t n_floats);
Q2W_API int  q2w_projection_dims(const q2w_state* s, int* n_rows, int* n_out);
Q2W_API const float* q2w_projected_device(const q2w_state* s);

/* ---- all GPUs of the box from one process (SURVEY 8(b) additive item 3, 8(e)) -------------------- */
/* One weight replica per device (the caller creates and uploads one q2w_model per device; the same device may be listed more than
 * once -- replicas are independent), one state + one host worker thread per replica.  A batch is cut into contiguous blocks, window
 * w -> replica floor(w * G / B); there is no collective on the compute path.  Results land in caller order in out_host (if non-NULL)
 * and stay on the producing devices; gather_device >= 0 additionally assembles all B x [n_out][n_state] embeddings, ordered by window
 * index, in one buffer on that device (peer copies over NVLink; q2w_multi_gathered_device).  The reference picks ONE device
 * (whisper_context_params.gpu_device, include/qwen2-whisper.h:118); this is the "-1 = all visible" extension. */
typedef struct q2w_multi q2w_multi;
Q2W_API int  q2w_multi_create(q2w_multi** out, q2w_model* const* models, int n_models, int max_batch_per_device);
Q2W_API void q2w_multi_free(q2w_multi* mm);                       /* frees the states and workers, NOT the models */
Q2W_API int  q2w_multi_n_devices(const q2w_multi* mm);
Q2W_API int  q2w_multi_device(const q2w_multi* mm, int i);        /* CUDA ordinal of replica i */
Q2W_API q2w_state* q2w_multi_state(q2w_multi* mm, int i);         /* replica i's state (single-window API, accessors) */
Q2W_API int  q2w_multi_shard_bounds(const q2w_multi* mm, int B, int i, int* lo, int* hi);
Q2W_API int  q2w_multi_set_max_batch(q2w_multi* mm, int max_batch_per_device);
Q2W_API int  q2w_multi_encode_batch_host(q2w_multi* mm, const float* pcm_host, size_t stride, const int32_t* n_samples, int B,
                                         float* out_host, int gather_device /* -1: no gather */);
/* asynchronous form (no gather): every device queues its shard and the call returns a ticket; at most two batches in flight, a third
 * submit first waits for the oldest; pcm_host / out_host must stay valid until q2w_multi_encode_batch_wait(ticket) */
Q2W_API int  q2w_multi_encode_batch_host_async(q2w_multi* mm, const float* pcm_host, size_t stride, const int32_t* n_samples, int B,
                                               float* out_host, int* ticket);
Q2W_API int  q2w_multi_encode_batch_wait(q2w_multi* mm, int ticket);
Q2W_API const float* q2w_multi_gathered_device(const q2w_multi* mm);
Q2W_API int  q2w_multi_get_gathered(q2w_multi* mm, float* out_host, size_t n_floats);
Q2W_API double q2w_multi_last_device_ms(const q2w_multi* mm, int i);   /* device time of replica i's last shard (CUDA events) */

/* ---- stage taps for the parity tests (drift must be localisable) ---------------------- */
/* stop every following forward pass after n_layers encoder blocks (0 = conv stem + positional embedding only; -1 = all, default);
 * q2w_debug_get_residual then copies the F32 residual stream x[n_audio_ctx][n_audio_state] of one window of the last forward:
 * the reference's `cur` / `inpL` at the same point of whisper_build_graph_encoder (:2005, :2154). */
Q2W_API int  q2w_debug_forward_layers(q2w_state* s, int n_layers);
Q2W_API int  q2w_debug_get_residual(q2w_state* s, int window, float* out_host);

/* ---- diagnostics -------------------------------------------------------------------------------- */
Q2W_API const char* q2w_last_error(void);
Q2W_API long q2w_kernel_launches(void);      /* number of this library's kernels launched so far (bench "gpu_launches") */
Q2W_API int  q2w_device_count(void);         /* sm_100 devices visible, 0 if none */
Q2W_API const char* q2w_build_info(void);

/* ---- kernel-level entry points (device pointers; stream = cudaStream_t or NULL) used by the parity tests ----- */
Q2W_API int q2w_op_gemm(const void* A_f16, int lda, const void* W_f16, int ldw, int M, int N, int K, const float* bias,
                        void* out, int ldo, int epilogue, const float* resid, const float* pos, int pos_period,
                        int scale_cols, float scale, void* stream);
/* same GEMM with W as raw ggml blocks (wtype = Q2W_TYPE_Q8_0 / Q2W_TYPE_Q4_0, K % 64 == 0): decoded inside the kernel */
Q2W_API int q2w_op_gemm_q(const void* A_f16, int lda, const void* W_raw, int wtype, int M, int N, int K, const float* bias,
                          void* out, int ldo, int epilogue, const float* resid, int scale_cols, float scale, void* stream);
Q2W_API int q2w_op_layernorm(const float* x, const float* gamma, const float* beta, void* y_f16, int M, int D, float eps,
                             void* stream);
Q2W_API int q2w_op_pool_layernorm(const float* x, const float* gamma, const float* beta, float* y, int B, int T, int D,
                                  float eps, void* stream);
Q2W_API int q2w_op_attention(const void* qkv_f16, void* out_f16, int B, int T, int H, void* stream);
Q2W_API int q2w_op_dequant(const void* src, int ggml_type, void* dst_f16, size_t rows, int K, void* stream);
/* n <= 4 quantised matrices (raw ggml blocks, Q2W_TYPE_Q8_0 / Q4_0, an even number of 32-element blocks each) -> f16, either by the
 * stand-alone kernel or (with_attention != 0) by the idle warps of ONE attention launch over qkv -> att_out, as the engine does per block */
Q2W_API int q2w_op_dequant_multi(const void* const* src, void* const* dst_f16, const unsigned long long* nblocks, int n, int ggml_type,
                                 int with_attention, const void* qkv_f16, void* att_out_f16, int B, int T, int H, void* stream);
Q2W_API int q2w_op_conv2_im2col(const void* h1_f16, void* A2_f16, int B, int T2, int C, void* stream);
/* window slice [offset, offset + n_ctx2) of B mel-major log-mels (zero past n_frames_valid, :2274-2283), optional clamp/normalise from
 * the per-window ordered-int max keys, conv1 im2col layout: A1 f16 [B * n_ctx2][3 * n_mel], column = ic * 3 + k */
Q2W_API int q2w_op_conv1_operand(const float* mel_dev, int ld_frames, int n_frames_valid, int n_mel, const void* win_max_keys_dev,
                                 int normalise, int offset, int n_ctx2, int B, void* A1_f16, void* stream);
/* split-K of the residual-epilogue GEMM at small M: 0 off (bit-reproducible single pass), 1 whole k-ranges, 2 balanced, -1 default */
Q2W_API void q2w_op_set_gemm_splitk(int mode);
/* mel: filters host [n_mel][201]; pcm device; logmel device [B][n_mel][ld]; win_max device int32[B] (ordered keys) */
Q2W_API int q2w_op_mel(const float* filters_host, int n_mel, const float* pcm_dev, size_t stride, const int32_t* n_samples_dev,
                       int n_max, int B, int n_frames, float* logmel_dev, int ld, void* win_max_dev, int normalise,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* Q2W_B200_H */
