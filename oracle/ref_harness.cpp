// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Builds the *unmodified* reference (whisper.cpp fork, CPU ggml backend) into oracle/_ref/ by
// #including its single translation unit where it lies under $(REF)/src/qwen2-whisper.cpp, and
// exposes a tiny C ABI so tests / bench.py (cpu_baseline leg) can drive it through ctypes.
// No reference source is copied into this repo.
//
// One deviation, stated wherever parity is reported: the fork's
// ggml_conv_1d() emits an F32 im2col (ggml/src/ggml.c:6642) whose mul_mat against an F16 conv
// kernel no backend accepts, so F16/Q8_0/Q4_0 model files abort as shipped.  The shim below
// upcasts the F16 conv kernel to F32 (lossless) before the reference's own ggml_conv_1d_ph.
//
// Reference entry points used (file:line in /root/reference):
//   whisper_init_from_file_with_params   src/qwen2-whisper.cpp:3139
//   whisper_init_from_buffer_with_params src/qwen2-whisper.cpp:3154
//   whisper_pcm_to_mel                   src/qwen2-whisper.cpp:3277  (log_mel_spectrogram :2575)
//   whisper_set_mel                      src/qwen2-whisper.cpp:3302
//   whisper_full                         src/qwen2-whisper.cpp:2377  (whisper_encode_qwen2_internal :2241)
//   state->mel / state->embd_enc         src/qwen2-whisper.cpp:795-864 (opaque; reachable because we include the TU)
//   ggml_quantize_chunk                  ggml/src/ggml.c (decl ggml/include/ggml.h:2345)
#include "ggml.h"
#include "ggml-backend.h"

static struct ggml_tensor * oracle_conv_1d_ph(struct ggml_context * ctx, struct ggml_tensor * a,
                                              struct ggml_tensor * b, int s, int d) {
    if (a->type != GGML_TYPE_F32) {
        a = ggml_cast(ctx, a, GGML_TYPE_F32);
    }
    return ggml_conv_1d_ph(ctx, a, b, s, d);
}
#define ggml_conv_1d_ph oracle_conv_1d_ph
#include "src/qwen2-whisper.cpp"
#undef ggml_conv_1d_ph

#include <cstring>

static void q2wref_quiet_log(ggml_log_level, const char *, void *) {}

// ggml fills its F16 <-> F32 lookup tables inside the first ggml_init() (ggml/src/ggml.c); the bare
// quantise / dequantise helpers below need them even when no model has been loaded yet.
static void q2wref_ensure_tables() {
    static bool done = false;
    if (!done) {
        struct ggml_init_params ip = { 1024, nullptr, false };
        struct ggml_context * c = ggml_init(ip);
        ggml_free(c);
        done = true;
    }
}

extern "C" {

__attribute__((visibility("default"))) void q2wref_set_quiet(int quiet) {
    whisper_log_set(quiet ? q2wref_quiet_log : nullptr, nullptr);
}

static whisper_context_params q2wref_cparams() {
    whisper_context_params cp = whisper_context_default_params();
    cp.use_gpu    = false;
    cp.flash_attn = false;   // the flash branch is commented out in the fork
    return cp;
}

__attribute__((visibility("default"))) void * q2wref_init_from_file(const char * path) {
    return whisper_init_from_file_with_params(path, q2wref_cparams());
}

__attribute__((visibility("default"))) void * q2wref_init_from_buffer(void * buf, size_t n) {
    return whisper_init_from_buffer_with_params(buf, n, q2wref_cparams());
}

__attribute__((visibility("default"))) void q2wref_free(void * h) { whisper_free((whisper_context *) h); }

__attribute__((visibility("default"))) int q2wref_pcm_to_mel(void * h, const float * pcm, int n, int n_threads) {
    return whisper_pcm_to_mel((whisper_context *) h, pcm, n, n_threads);
}

__attribute__((visibility("default"))) int q2wref_set_mel(void * h, const float * data, int n_len, int n_mel) {
    return whisper_set_mel((whisper_context *) h, data, n_len, n_mel);
}

// dims[0]=n_len dims[1]=n_len_org dims[2]=n_mel
__attribute__((visibility("default"))) void q2wref_mel_dims(void * h, int * dims) {
    auto * ctx = (whisper_context *) h;
    dims[0] = ctx->state->mel.n_len;
    dims[1] = ctx->state->mel.n_len_org;
    dims[2] = ctx->state->mel.n_mel;
}

__attribute__((visibility("default"))) void q2wref_get_mel(void * h, float * out) {
    auto * ctx = (whisper_context *) h;
    memcpy(out, ctx->state->mel.data.data(), ctx->state->mel.data.size() * sizeof(float));
}

// whisper_full with a hand-built params struct: whisper_full_default_params() has no return
// statement in the fork; the encoder path reads only n_threads, offset_ms,
// duration_ms and abort_callback (src/qwen2-whisper.cpp:2351-2369).
__attribute__((visibility("default"))) int q2wref_full(void * h, const float * pcm, int n, int n_threads, int offset_ms) {
    whisper_full_params p;
    memset(&p, 0, sizeof(p));
    p.n_threads = n_threads;
    p.offset_ms = offset_ms;
    return whisper_full((whisper_context *) h, p, pcm, n);
}

// dims[0]=n_state (ne[0]) dims[1]=n_out (ne[1]); returns element count or -1
__attribute__((visibility("default"))) long q2wref_embd_dims(void * h, int * dims) {
    auto * ctx = (whisper_context *) h;
    if (!ctx->state->embd_enc) return -1;
    dims[0] = (int) ctx->state->embd_enc->ne[0];
    dims[1] = (int) ctx->state->embd_enc->ne[1];
    return (long) ggml_nelements(ctx->state->embd_enc);
}

__attribute__((visibility("default"))) int q2wref_get_embd(void * h, float * out) {
    auto * ctx = (whisper_context *) h;
    if (!ctx->state->embd_enc) return -1;
    ggml_backend_tensor_get(ctx->state->embd_enc, out, 0, ggml_nbytes(ctx->state->embd_enc));
    return 0;
}

// t[0]=t_mel_us t[1]=t_encode_us t[2]=n_encode t[3]=t_load_us
__attribute__((visibility("default"))) void q2wref_timings(void * h, long long * t) {
    auto * ctx = (whisper_context *) h;
    t[0] = ctx->state->t_mel_us;
    t[1] = ctx->state->t_encode_us;
    t[2] = ctx->state->n_encode;
    t[3] = ctx->t_load_us;
}

__attribute__((visibility("default"))) void q2wref_reset_timings(void * h) { whisper_reset_timings((whisper_context *) h); }

__attribute__((visibility("default"))) void q2wref_hparams(void * h, int * hp) {
    auto * ctx = (whisper_context *) h;
    const auto & p = ctx->model.hparams;
    hp[0] = p.n_vocab; hp[1] = p.n_audio_ctx; hp[2] = p.n_audio_state; hp[3] = p.n_audio_head;
    hp[4] = p.n_audio_layer; hp[5] = p.n_mels; hp[6] = p.ftype; hp[7] = (int) ctx->wtype;
}

// ggml's own quantiser / dequantiser, for pinning the numpy restatement in oracle/ggml_quants.py
__attribute__((visibility("default"))) size_t q2wref_quantize(int type, const float * src, void * dst, long nrows, long n_per_row) {
    q2wref_ensure_tables();
    return ggml_quantize_chunk((ggml_type) type, src, dst, 0, nrows, n_per_row, nullptr);
}

__attribute__((visibility("default"))) void q2wref_dequantize(int type, const void * src, float * dst, long n) {
    q2wref_ensure_tables();
    ggml_internal_get_type_traits((ggml_type) type).to_float(src, dst, n);
}

__attribute__((visibility("default"))) size_t q2wref_row_size(int type, long ne) { return ggml_row_size((ggml_type) type, ne); }

// the F16 GELU table the CPU backend evaluates through (ggml/src/ggml.c:2556-2570)
__attribute__((visibility("default"))) void q2wref_gelu(const float * x, float * y, int n) {
    struct ggml_init_params ip = { (size_t) 16*1024*1024 + (size_t) n * 8 + 4096, nullptr, false };
    struct ggml_context * c = ggml_init(ip);
    struct ggml_tensor * a = ggml_new_tensor_1d(c, GGML_TYPE_F32, n);
    memcpy(a->data, x, (size_t) n * sizeof(float));
    struct ggml_tensor * g = ggml_gelu(c, a);
    struct ggml_cgraph * gf = ggml_new_graph(c);
    ggml_build_forward_expand(gf, g);
    ggml_graph_compute_with_ctx(c, gf, 1);
    memcpy(y, g->data, (size_t) n * sizeof(float));
    ggml_free(c);
}

} // extern "C"
