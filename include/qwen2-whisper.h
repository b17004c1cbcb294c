/* qwen2-whisper.h -- drop-in public C API of the B200-native audio front-end.
 *
 * Same symbols, signatures, struct layouts and return conventions as the reference's
 * include/qwen2-whisper.h for everything on the PCM -> log-mel -> encoder -> embeddings path
 * (SURVEY.md section 8(b)); a caller of the reference's libwhisper.so relinks against libq2w_b200.so.
 * Differences, all deliberate and documented in INTEGRATION.md:
 *   - whisper_encode / whisper_encode_with_state are DEFINED here (the reference only declares them, h:245-254).
 *   - whisper_full_default_params() actually returns its struct (reference: missing return, src:4231-4295).
 *   - decoder-era declarations the reference never defines (whisper_full_n_segments, ... h:452-510) are omitted.
 *   - additive entry points at the bottom: embeddings / mel accessors and the batched window API.
 *   - use_gpu = false or flash_attn = true are rejected at init (NULL + log): this build is CUDA-only and the
 *     reference's flash branch is an empty stub (src:2057-2079).
 * No ggml header is needed: the two callback typedefs the reference borrows from ggml.h are restated below.
 */
#ifndef QWEN2_WHISPER_H
#define QWEN2_WHISPER_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifndef WHISPER_API
#  define WHISPER_API __attribute__((visibility("default")))
#endif

#define WHISPER_SAMPLE_RATE 16000
#define WHISPER_N_FFT       400
#define WHISPER_HOP_LENGTH  160
#define WHISPER_CHUNK_SIZE  30

#ifdef __cplusplus
extern "C" {
#endif

/* ---- types borrowed from ggml.h by the reference header (ggml/include/ggml.h:560-567, :620, :2175) ---- */
#ifndef GGML_API   /* if the real ggml.h was included first, use its definitions */
enum ggml_log_level {
    GGML_LOG_LEVEL_NONE = 0, GGML_LOG_LEVEL_INFO = 1, GGML_LOG_LEVEL_WARN = 2,
    GGML_LOG_LEVEL_ERROR = 3, GGML_LOG_LEVEL_DEBUG = 4, GGML_LOG_LEVEL_CONT = 5
};
typedef bool (*ggml_abort_callback)(void * data);
typedef void (*ggml_log_callback)(enum ggml_log_level level, const char * text, void * user_data);
#endif

struct whisper_context;
struct whisper_state;
struct whisper_full_params;

typedef int32_t whisper_pos;
typedef int32_t whisper_token;
typedef int32_t whisper_seq_id;

/* kept only so that struct whisper_context_params has the reference's exact layout (h:87-128) */
enum whisper_alignment_heads_preset {
    WHISPER_AHEADS_NONE, WHISPER_AHEADS_N_TOP_MOST, WHISPER_AHEADS_CUSTOM, WHISPER_AHEADS_TINY_EN, WHISPER_AHEADS_TINY,
    WHISPER_AHEADS_BASE_EN, WHISPER_AHEADS_BASE, WHISPER_AHEADS_SMALL_EN, WHISPER_AHEADS_SMALL, WHISPER_AHEADS_MEDIUM_EN,
    WHISPER_AHEADS_MEDIUM, WHISPER_AHEADS_LARGE_V1, WHISPER_AHEADS_LARGE_V2, WHISPER_AHEADS_LARGE_V3,
    WHISPER_AHEADS_LARGE_V3_TURBO
};
typedef struct whisper_ahead  { int n_text_layer; int n_head; } whisper_ahead;
typedef struct whisper_aheads { size_t n_heads; const whisper_ahead * heads; } whisper_aheads;

struct whisper_context_params {
    bool use_gpu;       /* must be true: there is no CPU path */
    bool flash_attn;    /* must be false (see header comment) */
    int  gpu_device;    /* CUDA ordinal */
    bool dtw_token_timestamps;                           /* decoder-era, ignored */
    enum whisper_alignment_heads_preset dtw_aheads_preset;
    int  dtw_n_top;
    struct whisper_aheads dtw_aheads;
    size_t dtw_mem_size;
};

typedef struct whisper_model_loader {
    void * context;
    size_t (*read)(void * ctx, void * output, size_t read_size);
    bool   (*eof)(void * ctx);
    void   (*close)(void * ctx);
} whisper_model_loader;

/* ---- init / free (h:141-149, :176, :203-206) ---- */
WHISPER_API struct whisper_context * whisper_init_from_file_with_params  (const char * path_model, struct whisper_context_params params);
WHISPER_API struct whisper_context * whisper_init_from_buffer_with_params(void * buffer, size_t buffer_size, struct whisper_context_params params);
WHISPER_API struct whisper_context * whisper_init_with_params            (struct whisper_model_loader * loader, struct whisper_context_params params);
WHISPER_API struct whisper_context * whisper_init_from_file_with_params_no_state  (const char * path_model, struct whisper_context_params params);
WHISPER_API struct whisper_context * whisper_init_from_buffer_with_params_no_state(void * buffer, size_t buffer_size, struct whisper_context_params params);
WHISPER_API struct whisper_context * whisper_init_with_params_no_state            (struct whisper_model_loader * loader, struct whisper_context_params params);
/* the deprecated default-params spellings (h:151-174, src:3184-3206); still exported so that old callers relink */
WHISPER_API struct whisper_context * whisper_init_from_file  (const char * path_model);
WHISPER_API struct whisper_context * whisper_init_from_buffer(void * buffer, size_t buffer_size);
WHISPER_API struct whisper_context * whisper_init            (struct whisper_model_loader * loader);
WHISPER_API struct whisper_context * whisper_init_from_file_no_state  (const char * path_model);
WHISPER_API struct whisper_context * whisper_init_from_buffer_no_state(void * buffer, size_t buffer_size);
WHISPER_API struct whisper_context * whisper_init_no_state            (struct whisper_model_loader * loader);
WHISPER_API struct whisper_state   * whisper_init_state(struct whisper_context * ctx);
WHISPER_API void whisper_free      (struct whisper_context * ctx);
WHISPER_API void whisper_free_state(struct whisper_state * state);
WHISPER_API void whisper_free_params(struct whisper_full_params * params);
WHISPER_API void whisper_free_context_params(struct whisper_context_params * params);

/* ---- mel (h:211-239) ---- */
WHISPER_API int whisper_pcm_to_mel(struct whisper_context * ctx, const float * samples, int n_samples, int n_threads);
WHISPER_API int whisper_pcm_to_mel_with_state(struct whisper_context * ctx, struct whisper_state * state, const float * samples, int n_samples, int n_threads);
WHISPER_API int whisper_set_mel(struct whisper_context * ctx, const float * data, int n_len, int n_mel);
WHISPER_API int whisper_set_mel_with_state(struct whisper_context * ctx, struct whisper_state * state, const float * data, int n_len, int n_mel);

/* ---- encoder (h:245-254; defined here) ---- */
WHISPER_API int whisper_encode(struct whisper_context * ctx, int offset, int n_threads);
WHISPER_API int whisper_encode_with_state(struct whisper_context * ctx, struct whisper_state * state, int offset, int n_threads);

/* ---- getters (h:288-306) ---- */
WHISPER_API int whisper_n_len           (struct whisper_context * ctx);
WHISPER_API int whisper_n_len_from_state(struct whisper_state * state);
WHISPER_API int whisper_n_vocab         (struct whisper_context * ctx);
WHISPER_API int whisper_n_text_ctx      (struct whisper_context * ctx);
WHISPER_API int whisper_n_audio_ctx     (struct whisper_context * ctx);
WHISPER_API int whisper_model_n_vocab      (struct whisper_context * ctx);
WHISPER_API int whisper_model_n_audio_ctx  (struct whisper_context * ctx);
WHISPER_API int whisper_model_n_audio_state(struct whisper_context * ctx);
WHISPER_API int whisper_model_n_audio_head (struct whisper_context * ctx);
WHISPER_API int whisper_model_n_audio_layer(struct whisper_context * ctx);
WHISPER_API int whisper_model_n_text_ctx   (struct whisper_context * ctx);
WHISPER_API int whisper_model_n_text_state (struct whisper_context * ctx);
WHISPER_API int whisper_model_n_text_head  (struct whisper_context * ctx);
WHISPER_API int whisper_model_n_text_layer (struct whisper_context * ctx);
WHISPER_API int whisper_model_n_mels       (struct whisper_context * ctx);
WHISPER_API int whisper_model_ftype        (struct whisper_context * ctx);
WHISPER_API int whisper_model_type         (struct whisper_context * ctx);
WHISPER_API const char * whisper_model_type_readable(struct whisper_context * ctx);

/* ---- timings / info / log (h:335-339, :526-527) ---- */
WHISPER_API void whisper_print_timings(struct whisper_context * ctx);
WHISPER_API void whisper_reset_timings(struct whisper_context * ctx);
WHISPER_API const char * whisper_print_system_info(void);
WHISPER_API void whisper_log_set(ggml_log_callback log_callback, void * user_data);
WHISPER_API void whisper_print_emb_enc(struct whisper_context * ctx);

/* ---- whisper_full (h:346-450): layout identical to the reference; the encoder path reads only
 *      n_threads, offset_ms, duration_ms, abort_callback(+user_data)  (src:2351-2369) ---- */
typedef void (*whisper_new_segment_callback)(struct whisper_context * ctx, struct whisper_state * state, int n_new, void * user_data);
typedef void (*whisper_progress_callback)(struct whisper_context * ctx, struct whisper_state * state, int progress, void * user_data);
typedef bool (*whisper_encoder_begin_callback)(struct whisper_context * ctx, struct whisper_state * state, void * user_data);

struct whisper_full_params {
    int n_threads;            /* accepted, ignored (GPU path) */
    int n_max_text_ctx;
    int offset_ms;            /* seek = offset_ms / 10 mel frames */
    int duration_ms;

    bool translate, no_context, no_timestamps, single_segment, print_special, print_progress, print_realtime, print_timestamps;

    bool  token_timestamps;
    float thold_pt, thold_ptsum;
    int   max_len;
    bool  split_on_word;
    int   max_tokens;

    bool debug_mode;
    int  audio_ctx;

    bool tdrz_enable;

    const char * suppress_regex;

    const char * initial_prompt;
    const whisper_token * prompt_tokens;
    int prompt_n_tokens;

    const char * language;
    bool detect_language;

    bool suppress_blank, suppress_non_speech_tokens;

    float temperature, max_initial_ts, length_penalty;
    float temperature_inc, entropy_thold, logprob_thold, no_speech_thold;

    whisper_new_segment_callback new_segment_callback;
    void * new_segment_callback_user_data;
    whisper_progress_callback progress_callback;
    void * progress_callback_user_data;
    whisper_encoder_begin_callback encoder_begin_callback;
    void * encoder_begin_callback_user_data;
    ggml_abort_callback abort_callback;
    void * abort_callback_user_data;

    size_t i_start_rule;
};

WHISPER_API struct whisper_context_params * whisper_context_default_params_by_ref(void);
WHISPER_API struct whisper_context_params   whisper_context_default_params(void);
WHISPER_API struct whisper_full_params      whisper_full_default_params(void);
WHISPER_API struct whisper_full_params *    whisper_full_default_params_by_ref(void);   /* additive, for FFI callers */

/* mel (only if n_samples > 0) + one encoder pass at seek = offset_ms / 10.  0 ok, -1 encode failed, -2 mel failed,
 * 0 + warning if fewer than 100 frames remain (src:2341-2383). */
WHISPER_API int whisper_full(struct whisper_context * ctx, struct whisper_full_params params, const float * samples, int n_samples);
WHISPER_API int whisper_full_with_state(struct whisper_context * ctx, struct whisper_state * state, struct whisper_full_params params, const float * samples, int n_samples);

/* ================= additive API (SURVEY 8(b) "Additive API needed") ================= */
/* embeddings of the last encode: [n_windows][n_out = n_audio_ctx/2][n_state] float */
WHISPER_API int whisper_embd_dims(struct whisper_context * ctx, int * n_windows, int * n_out, int * n_state);
WHISPER_API int whisper_get_embeddings(struct whisper_context * ctx, float * dst, size_t n_floats);            /* copies to host */
WHISPER_API int whisper_get_embeddings_from_state(struct whisper_state * state, float * dst, size_t n_floats);
WHISPER_API const float * whisper_get_embeddings_device(struct whisper_context * ctx);                       /* device pointer */
/* mel of the default state, float[n_mel][n_len]; n_len is the padded length (n_samples + 30 s) / 160, whereas whisper_n_len()
 * returns n_len_org like the reference (src:3440-3446) */
WHISPER_API int whisper_get_mel_dims(struct whisper_context * ctx, int * n_len, int * n_len_org, int * n_mel);
WHISPER_API int whisper_get_mel(struct whisper_context * ctx, float * dst, size_t n_floats);
/* B independent 30 s windows of PCM (window b at samples + b*stride, n_samples[b] valid, NULL = full windows):
 * per-window mel + encoder, results [B][n_out][n_state] copied to dst if non-NULL.  max windows per micro-batch is
 * set once with whisper_set_max_batch (default 16). */
WHISPER_API int whisper_encode_batch(struct whisper_context * ctx, const float * samples, size_t stride, const int32_t * n_samples, int n_windows, float * dst);
/* asynchronous form: returns a ticket (>= 0) as soon as the batch is queued, -1 on error; samples / dst must stay valid (pinned for
 * real overlap) until whisper_encode_batch_wait(ticket) returns 0. At most two batches are in flight per context. */
WHISPER_API int whisper_encode_batch_async(struct whisper_context * ctx, const float * samples, size_t stride, const int32_t * n_samples, int n_windows, float * dst);
WHISPER_API int whisper_encode_batch_wait(struct whisper_context * ctx, int ticket);
WHISPER_API int whisper_encode_batch_device(struct whisper_context * ctx, const float * samples_dev, size_t stride, const int32_t * n_samples, int n_windows);
/* whole-file streaming: after ONE whisper_pcm_to_mel over the full audio (global normalisation), encode n windows starting at the
 * given mel-frame offsets (offset_ms / 10) as a batch == n x whisper_full(ctx, {offset_ms}, NULL, 0) of the reference */
WHISPER_API int whisper_encode_offsets(struct whisper_context * ctx, const int32_t * mel_offsets, int n_windows, float * dst);
WHISPER_API int whisper_set_max_batch(struct whisper_context * ctx, int max_batch);   /* per device; resizes scratch in place, mel + embeddings survive */
/* ---- every GPU of the box from one process.  whisper_context_params.gpu_device = -1 loads one weight
 *      replica on every visible sm_100 device (the reference's field picks one device, h:118); the _multi initialisers take an explicit
 *      list (an ordinal may repeat: independent replicas).  whisper_encode_batch then shards the windows over the replicas, window w ->
 *      replica floor(w * G / n_windows), no collective; whisper_encode_batch_multi(..., gather_device) additionally gathers all
 *      embeddings, in window order, on one device (peer copies over NVLink).  The single-window API runs on replica 0. */
WHISPER_API struct whisper_context * whisper_init_from_file_multi  (const char * path_model, struct whisper_context_params params, const int * devices, int n_devices);
WHISPER_API struct whisper_context * whisper_init_from_buffer_multi(void * buffer, size_t buffer_size, struct whisper_context_params params, const int * devices, int n_devices);
WHISPER_API int whisper_n_devices(struct whisper_context * ctx);
WHISPER_API int whisper_device   (struct whisper_context * ctx, int i);
WHISPER_API int whisper_encode_batch_multi(struct whisper_context * ctx, const float * samples, size_t stride, const int32_t * n_samples, int n_windows, float * dst, int gather_device);
WHISPER_API const float * whisper_get_gathered_device(struct whisper_context * ctx);
WHISPER_API void * whisper_q2w_multi(struct whisper_context * ctx);   /* q2w_multi* (include/q2w_b200.h) or NULL */
/* the step after the path: Qwen2-Audio's multi_modal_projector, Linear(n_audio_state -> n_out) + bias, applied to the
 * embeddings of the last encode by one more tcgen05 GEMM.  weight: [n_out][n_audio_state] row-major, ggml_type 0 (F32) or 1 (F16);
 * bias float[n_out] or NULL.  whisper_project -> float[n_windows * n_audio_ctx/2][n_out] (copied to dst if non-NULL). */
WHISPER_API int whisper_set_projector(struct whisper_context * ctx, int ggml_type, int n_out, const void * weight, size_t nbytes, const float * bias);
WHISPER_API int whisper_project(struct whisper_context * ctx, float * dst, size_t n_floats);
WHISPER_API int whisper_projection_dims(struct whisper_context * ctx, int * n_rows, int * n_out);
/* the underlying C-ABI state handle (q2w_state*, include/q2w_b200.h) for callers that need streams / device pointers */
WHISPER_API void * whisper_q2w_state(struct whisper_context * ctx);

#ifdef __cplusplus
}
#endif
#endif /* QWEN2_WHISPER_H */
