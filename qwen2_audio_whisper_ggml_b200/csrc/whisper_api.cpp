// whisper_api.cpp -- host side of the drop-in boundary: the reference's public C API (include/qwen2-whisper.h)
// implemented over the C-ABI CUDA library (include/q2w_b200.h).  Plain C++, no CUDA, no ggml.
//
// This is the reference's L2 "fork runtime" minus everything ggml: model-file loader (src/qwen2-whisper.cpp:1350-1872),
// context/state lifetime (:2779-2951, :3012-3266), mel setters/getters (:3268-3308), the encode driver
// (:2341-2383), timings and logging (:3516-3555, :4186-4229) -- with the same names, argument meaning, return codes
// and log texts, so parity tests read like calls against the reference.
#include "qwen2-whisper.h"
#include "q2w_b200.h"

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

namespace {

void default_log(ggml_log_level, const char* text, void*) {
    fputs(text, stderr);
    fflush(stderr);
}
ggml_log_callback g_log = default_log;
void* g_log_ud = nullptr;

__attribute__((format(printf, 2, 3))) void wlog(ggml_log_level lvl, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_log(lvl, buf, g_log_ud);
}
#define LOG_INFO(...) wlog(GGML_LOG_LEVEL_INFO, __VA_ARGS__)
#define LOG_WARN(...) wlog(GGML_LOG_LEVEL_WARN, __VA_ARGS__)
#define LOG_ERROR(...) wlog(GGML_LOG_LEVEL_ERROR, __VA_ARGS__)

int64_t now_us() {
    return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

constexpr uint32_t FILE_MAGIC = 0x67676d6c;        // "ggml"
constexpr int QNT_VERSION_FACTOR = 1000;           // ggml.h GGML_QNT_VERSION_FACTOR

// enum ggml_ftype -> ggml_type for the types on this path (ggml_ftype_to_ggml_type)
int ftype_to_type(int ftype) {
    switch (ftype) {
        case 0: return Q2W_TYPE_F32;
        case 1: return Q2W_TYPE_F16;
        case 2: return Q2W_TYPE_Q4_0;
        case 7: return Q2W_TYPE_Q8_0;
        default: return -1;
    }
}

size_t type_size(int t) { return t == 0 ? 4 : t == 1 ? 2 : t == 2 ? 18 : t == 8 ? 34 : 0; }
int blck_size(int t) { return (t == 2 || t == 8) ? 32 : 1; }

template <typename T>
bool read_safe(whisper_model_loader* l, T& dst) {
    return l->read(l->context, &dst, sizeof(T)) == sizeof(T);
}

}  // namespace

struct whisper_state {
    q2w_state* qs = nullptr;
    bool owned = true;       // false: replica 0's state of a multi-device context (owned by the q2w_multi handle)
};

struct whisper_context {
    int64_t t_load_us = 0;
    int64_t t_start_us = 0;
    whisper_context_params params{};
    q2w_hparams hp{};
    int ftype = 1;           // hparams.ftype % factor
    int wtype = Q2W_TYPE_F16;
    int model_type = 0;      // e_model
    q2w_model* model = nullptr;            // replica 0 (the only one unless gpu_device == -1 / an explicit device list)
    std::vector<q2w_model*> replicas;      // one per device, replicas[0] == model
    std::vector<int> devices;              // CUDA ordinals, parallel to replicas
    q2w_multi* multi = nullptr;            // workers + states over all replicas (multi-device contexts only)
    whisper_state* state = nullptr;
    int max_batch = 16;
    std::string path_model;
};

static void free_models(whisper_context& c) {
    for (q2w_model* m : c.replicas) q2w_model_free(m);
    c.replicas.clear();
    c.model = nullptr;
}

namespace {

bool model_load(whisper_model_loader* loader, whisper_context& wctx) {
    LOG_INFO("%s: loading model\n", "whisper_model_load");
    const int64_t t_start_us = now_us();
    wctx.t_start_us = t_start_us;

    // every read is checked against the requested size: a file truncated anywhere -- inside the filterbank, a name, the last tensor's
    // payload -- fails here instead of uploading stale bytes (the reference only notices a missing tensor, :1861)
    auto read_exact = [&](void* dst, size_t n) { return n == 0 || loader->read(loader->context, dst, n) == n; };
    uint32_t magic = 0;
    read_safe(loader, magic);
    if (magic != FILE_MAGIC) {
        LOG_ERROR("%s: invalid model data (bad magic)\n", "whisper_model_load");
        return false;
    }
    q2w_hparams& hp = wctx.hp;
    int32_t* fields[11] = {&hp.n_vocab, &hp.n_audio_ctx, &hp.n_audio_state, &hp.n_audio_head, &hp.n_audio_layer, &hp.n_text_ctx,
                           &hp.n_text_state, &hp.n_text_head, &hp.n_text_layer, &hp.n_mels, &hp.ftype};
    for (int32_t* f : fields) {
        if (!read_safe(loader, *f)) {
            LOG_ERROR("%s: truncated model header\n", "whisper_model_load");
            return false;
        }
    }
    switch (hp.n_audio_layer) {   // e_model, :1390-1412
        case 4: wctx.model_type = 1; break;
        case 6: wctx.model_type = 2; break;
        case 12: wctx.model_type = 3; break;
        case 24: wctx.model_type = 4; break;
        case 32: wctx.model_type = 5; break;
        default: wctx.model_type = 0;
    }
    const int qntvr = hp.ftype / QNT_VERSION_FACTOR;
    wctx.ftype = hp.ftype % QNT_VERSION_FACTOR;
    wctx.wtype = ftype_to_type(wctx.ftype);
    if (wctx.wtype < 0) {
        LOG_ERROR("%s: invalid model (bad ftype value %d)\n", "whisper_model_load", wctx.ftype);
        return false;
    }
    LOG_INFO("%s: n_vocab       = %d\n", "whisper_model_load", hp.n_vocab);
    LOG_INFO("%s: n_audio_ctx   = %d\n", "whisper_model_load", hp.n_audio_ctx);
    LOG_INFO("%s: n_audio_state = %d\n", "whisper_model_load", hp.n_audio_state);
    LOG_INFO("%s: n_audio_head  = %d\n", "whisper_model_load", hp.n_audio_head);
    LOG_INFO("%s: n_audio_layer = %d\n", "whisper_model_load", hp.n_audio_layer);
    LOG_INFO("%s: n_mels        = %d\n", "whisper_model_load", hp.n_mels);
    LOG_INFO("%s: ftype         = %d\n", "whisper_model_load", wctx.ftype);
    LOG_INFO("%s: qntvr         = %d\n", "whisper_model_load", qntvr);

    // mel filters (:1442-1451)
    int32_t n_mel = 0, n_fft = 0;
    if (!read_safe(loader, n_mel) || !read_safe(loader, n_fft) || n_mel <= 0 || n_fft <= 0 || n_mel > 1024 || n_fft > 4096) {
        LOG_ERROR("%s: invalid mel filterbank header (%d x %d)\n", "whisper_model_load", n_mel, n_fft);
        return false;
    }
    std::vector<float> filters(static_cast<size_t>(n_mel) * n_fft);
    if (!read_exact(filters.data(), filters.size() * sizeof(float))) {
        LOG_ERROR("%s: truncated model file (mel filterbank)\n", "whisper_model_load");
        return false;
    }

    // vocab (:1454-1487): parsed and dropped -- nothing on the encoder path reads it
    int32_t n_vocab = 0;
    if (!read_safe(loader, n_vocab) || n_vocab < 0) {
        LOG_ERROR("%s: truncated model file (vocab header)\n", "whisper_model_load");
        return false;
    }
    std::vector<char> tmp;
    for (int i = 0; i < n_vocab; ++i) {
        uint32_t len = 0;
        if (!read_safe(loader, len) || len > (1u << 20)) {
            LOG_ERROR("%s: truncated vocab\n", "whisper_model_load");
            return false;
        }
        tmp.resize(len);
        if (!read_exact(tmp.data(), len)) {
            LOG_ERROR("%s: truncated vocab\n", "whisper_model_load");
            return false;
        }
    }

    hp.ftype = wctx.ftype;
    // one weight replica per device: gpu_device >= 0 -> that device (the reference's meaning, include/qwen2-whisper.h:118);
    // gpu_device == -1 -> every visible sm_100 device; an explicit list comes from whisper_init_*_multi
    if (wctx.devices.empty()) {
        if (wctx.params.gpu_device >= 0) {
            wctx.devices.push_back(wctx.params.gpu_device);
        } else {
            const int n = q2w_device_count();
            for (int d = 0; d < n; ++d) wctx.devices.push_back(d);
            if (wctx.devices.empty()) {
                LOG_ERROR("%s: gpu_device = -1 but no sm_100 device is visible\n", "whisper_model_load");
                return false;
            }
        }
    }
    int rc = Q2W_OK;
    for (size_t r = 0; r < wctx.devices.size(); ++r) {
        q2w_model* m = nullptr;
        rc = q2w_model_create(&m, &hp, wctx.wtype, wctx.devices[r]);
        if (rc != Q2W_OK) {
            LOG_ERROR("%s: failed to allocate memory for the model on device %d: %s\n", "whisper_model_load", wctx.devices[r], q2w_last_error());
            return false;
        }
        wctx.replicas.push_back(m);
        char nm[16];
        snprintf(nm, sizeof(nm), "CUDA%d", wctx.devices[r]);
        LOG_INFO("%s: %8s total size = %8.2f MB\n", "whisper_model_load", nm, q2w_model_weight_bytes(m) / 1e6);
        if ((rc = q2w_model_upload_filters(m, filters.data(), n_mel, n_fft)) != Q2W_OK) {
            LOG_ERROR("%s: %s\n", "whisper_model_load", q2w_last_error());
            return false;
        }
    }
    wctx.model = wctx.replicas[0];

    // tensor stream (:1782-1855)
    size_t total_size = 0;
    int n_loaded = 0;
    std::vector<char> read_buf;
    while (true) {
        int32_t n_dims = 0, length = 0, ttype = 0;
        const bool got_dims = read_safe(loader, n_dims);
        if (!got_dims && loader->eof(loader->context)) break;          // clean end of the tensor stream (:1787-1790)
        if (!got_dims || !read_safe(loader, length) || !read_safe(loader, ttype)) {
            LOG_ERROR("%s: truncated model file (tensor header)\n", "whisper_model_load");
            return false;
        }
        if (n_dims < 1 || n_dims > 4 || length < 0 || length > 4096) {
            LOG_ERROR("%s: corrupt tensor record (n_dims %d, name length %d)\n", "whisper_model_load", n_dims, length);
            return false;
        }
        int64_t nelements = 1;
        int32_t ne[4] = {1, 1, 1, 1};
        for (int i = 0; i < n_dims; ++i) {
            if (!read_safe(loader, ne[i]) || ne[i] <= 0 || (nelements *= ne[i]) > (int64_t(1) << 31)) {   // the largest real tensor has 6.5 M elements
                LOG_ERROR("%s: truncated or corrupt tensor record (dims)\n", "whisper_model_load");
                return false;
            }
        }
        std::string name(static_cast<size_t>(length), '\0');
        if (!read_exact(&name[0], name.size())) {
            LOG_ERROR("%s: truncated model file (tensor name)\n", "whisper_model_load");
            return false;
        }
        const size_t bpe = type_size(ttype);
        if (bpe == 0 || ne[0] % blck_size(ttype)) {
            LOG_ERROR("%s: tensor '%s' has unsupported type %d or bad shape\n", "whisper_model_load", name.c_str(), ttype);
            return false;
        }
        const size_t nbytes = static_cast<size_t>(nelements) / blck_size(ttype) * bpe;
        {   // reject a record the model cannot hold before reading (or allocating for) its payload: same messages as :1807 / :1829
            const size_t want = q2w_model_tensor_bytes(wctx.model, name.c_str());
            if (want == 0) {
                LOG_ERROR("%s: unknown tensor '%s' in model file\n", "whisper_model_load", name.c_str());
                return false;
            }
            if (want != nbytes) {
                LOG_ERROR("%s: tensor '%s' has wrong size in model file: got %zu, expected %zu\n", "whisper_model_load", name.c_str(), nbytes, want);
                return false;
            }
        }
        read_buf.resize(nbytes);
        if (!read_exact(read_buf.data(), nbytes)) {
            LOG_ERROR("%s: truncated model file: tensor '%s' needs %zu bytes\n", "whisper_model_load", name.c_str(), nbytes);
            return false;
        }
        for (q2w_model* m : wctx.replicas) {
            rc = q2w_model_upload_tensor(m, name.c_str(), ttype, std::min(n_dims, 3), ne, read_buf.data(), nbytes);
            if (rc != Q2W_OK) {
                LOG_ERROR("%s: %s\n", "whisper_model_load", q2w_last_error());
                return false;
            }
        }
        total_size += nbytes;
        n_loaded++;
    }
    LOG_INFO("%s: model size    = %7.2f MB\n", "whisper_model_load", total_size / 1e6);
    for (q2w_model* m : wctx.replicas) {
        if ((rc = q2w_model_finalize(m)) != Q2W_OK) {
            LOG_ERROR("%s: ERROR %s\n", "whisper_model_load", q2w_last_error());
            return false;
        }
    }
    (void) n_loaded;
    wctx.t_load_us = now_us() - t_start_us;
    return true;
}

}  // namespace

extern "C" {

struct whisper_context_params whisper_context_default_params(void) {
    whisper_context_params r{};
    r.use_gpu = true;
    r.flash_attn = false;
    r.gpu_device = 0;
    r.dtw_token_timestamps = false;
    r.dtw_aheads_preset = WHISPER_AHEADS_NONE;
    r.dtw_n_top = -1;
    r.dtw_aheads.n_heads = 0;
    r.dtw_aheads.heads = nullptr;
    r.dtw_mem_size = 1024 * 1024 * 128;
    return r;
}

struct whisper_context_params* whisper_context_default_params_by_ref(void) {
    whisper_context_params* p = new whisper_context_params;
    *p = whisper_context_default_params();
    return p;
}

struct whisper_full_params whisper_full_default_params(void) {
    whisper_full_params r{};
    r.n_threads = std::min(4, static_cast<int>(std::thread::hardware_concurrency()));
    r.n_max_text_ctx = 16384;
    r.no_context = true;
    r.print_progress = true;
    r.print_timestamps = true;
    r.thold_pt = 0.01f;
    r.thold_ptsum = 0.01f;
    r.language = "en";
    r.suppress_blank = true;
    r.max_initial_ts = 1.0f;
    r.length_penalty = -1.0f;
    r.temperature_inc = 0.2f;
    r.entropy_thold = 2.4f;
    r.logprob_thold = -1.0f;
    r.no_speech_thold = 0.6f;
    return r;   // the reference forgets this line (src:4231-4295)
}

struct whisper_full_params* whisper_full_default_params_by_ref(void) {
    whisper_full_params* p = new whisper_full_params;
    *p = whisper_full_default_params();
    return p;
}

void whisper_free_context_params(struct whisper_context_params* params) { delete params; }
void whisper_free_params(struct whisper_full_params* params) { delete params; }

static whisper_context* init_no_state(struct whisper_model_loader* loader, struct whisper_context_params params, const int* devices, int n_devices) {
    if (!loader || !loader->read || !loader->eof || !loader->close) {
        LOG_ERROR("%s: invalid model loader\n", __func__);
        return nullptr;
    }
    LOG_INFO("%s: use gpu    = %d\n", __func__, params.use_gpu);
    LOG_INFO("%s: flash attn = %d\n", __func__, params.flash_attn);
    LOG_INFO("%s: gpu_device = %d\n", __func__, params.gpu_device);
    if (!params.use_gpu) {
        loader->close(loader->context);
        LOG_ERROR("%s: use_gpu = false is not supported: this build has no CPU path\n", __func__);
        return nullptr;
    }
    if (params.flash_attn) {
        loader->close(loader->context);
        LOG_ERROR("%s: flash_attn = true is not supported (the reference's flash branch skips attention; attention here is always fused)\n", __func__);
        return nullptr;
    }
    whisper_context* ctx = new whisper_context;
    ctx->params = params;
    for (int i = 0; i < n_devices; ++i) ctx->devices.push_back(devices[i]);
    if (!model_load(loader, *ctx)) {
        loader->close(loader->context);
        LOG_ERROR("%s: failed to load model\n", __func__);
        free_models(*ctx);
        delete ctx;
        return nullptr;
    }
    loader->close(loader->context);
    return ctx;
}

struct whisper_context* whisper_init_with_params_no_state(struct whisper_model_loader* loader, struct whisper_context_params params) {
    return init_no_state(loader, params, nullptr, 0);
}

namespace {
// std::ifstream / memory-buffer loaders shared by the _from_file / _from_buffer entry points
struct buf_context { uint8_t* p; size_t size; size_t off; bool hit_end; };
whisper_model_loader file_loader(std::ifstream* fin) {
    whisper_model_loader loader = {};
    loader.context = fin;
    loader.read = [](void* c, void* out, size_t n) -> size_t {
        std::ifstream* f = static_cast<std::ifstream*>(c);
        f->read(static_cast<char*>(out), static_cast<std::streamsize>(n));
        return static_cast<size_t>(f->gcount());
    };
    loader.eof = [](void* c) -> bool { return static_cast<std::ifstream*>(c)->eof(); };
    loader.close = [](void* c) { static_cast<std::ifstream*>(c)->close(); };
    return loader;
}
whisper_model_loader buffer_loader(buf_context* bc) {
    whisper_model_loader loader = {};
    loader.context = bc;
    loader.read = [](void* c, void* out, size_t n) -> size_t {
        buf_context* b = static_cast<buf_context*>(c);
        const size_t k = std::min(n, b->size - b->off);
        if (k < n) b->hit_end = true;
        memcpy(out, b->p + b->off, k);
        b->off += k;
        return k;
    };
    // like an ifstream, eof only turns true once a read ran past the end (so a file ending exactly after a tensor works)
    loader.eof = [](void* c) -> bool { return static_cast<buf_context*>(c)->hit_end; };
    loader.close = [](void*) {};
    return loader;
}
}  // namespace

static whisper_context* init_from_file_no_state(const char* path_model, struct whisper_context_params params, const int* devices, int n_devices,
                                                const char* fn) {
    LOG_INFO("%s: loading model from '%s'\n", fn, path_model ? path_model : "(null)");
    std::ifstream fin;
    if (path_model) fin.open(path_model, std::ios::binary);
    if (!path_model || !fin) {
        LOG_ERROR("%s: failed to open '%s'\n", fn, path_model ? path_model : "(null)");
        return nullptr;
    }
    whisper_model_loader loader = file_loader(&fin);
    whisper_context* ctx = init_no_state(&loader, params, devices, n_devices);
    if (ctx) ctx->path_model = path_model;
    return ctx;
}

static whisper_context* init_from_buffer_no_state(void* buffer, size_t buffer_size, struct whisper_context_params params, const int* devices,
                                                  int n_devices, const char* fn) {
    LOG_INFO("%s: loading model from buffer\n", fn);
    if (!buffer) {
        LOG_ERROR("%s: null buffer\n", fn);
        return nullptr;
    }
    buf_context bc = {static_cast<uint8_t*>(buffer), buffer_size, 0, false};
    whisper_model_loader loader = buffer_loader(&bc);
    return init_no_state(&loader, params, devices, n_devices);
}

struct whisper_context* whisper_init_from_file_with_params_no_state(const char* path_model, struct whisper_context_params params) {
    return init_from_file_no_state(path_model, params, nullptr, 0, __func__);
}

struct whisper_context* whisper_init_from_buffer_with_params_no_state(void* buffer, size_t buffer_size, struct whisper_context_params params) {
    return init_from_buffer_no_state(buffer, buffer_size, params, nullptr, 0, __func__);
}

struct whisper_state* whisper_init_state(struct whisper_context* ctx) {
    if (!ctx || !ctx->model) return nullptr;
    whisper_state* st = new whisper_state;
    const int rc = q2w_state_create(&st->qs, ctx->model, ctx->max_batch);
    if (rc != Q2W_OK) {
        LOG_ERROR("%s: whisper_backend_init() failed: %s\n", __func__, q2w_last_error());
        delete st;
        return nullptr;
    }
    return st;
}

// default state. One device: a plain state. Several replicas: the multi-device handle owns one state per replica (plus one host
// worker thread each) and the context's default state is replica 0's, so the single-window API keeps working unchanged.
static whisper_context* with_state(whisper_context* ctx) {
    if (!ctx) return nullptr;
    if (ctx->replicas.size() > 1) {
        const int rc = q2w_multi_create(&ctx->multi, ctx->replicas.data(), static_cast<int>(ctx->replicas.size()), ctx->max_batch);
        if (rc != Q2W_OK) {
            LOG_ERROR("%s: multi-device init failed: %s\n", __func__, q2w_last_error());
            whisper_free(ctx);
            return nullptr;
        }
        ctx->state = new whisper_state;
        ctx->state->qs = q2w_multi_state(ctx->multi, 0);
        ctx->state->owned = false;
        return ctx;
    }
    ctx->state = whisper_init_state(ctx);
    if (!ctx->state) {
        whisper_free(ctx);
        return nullptr;
    }
    return ctx;
}

struct whisper_context* whisper_init_from_file_with_params(const char* path_model, struct whisper_context_params params) {
    return with_state(whisper_init_from_file_with_params_no_state(path_model, params));
}
struct whisper_context* whisper_init_from_buffer_with_params(void* buffer, size_t buffer_size, struct whisper_context_params params) {
    return with_state(whisper_init_from_buffer_with_params_no_state(buffer, buffer_size, params));
}
struct whisper_context* whisper_init_with_params(struct whisper_model_loader* loader, struct whisper_context_params params) {
    return with_state(whisper_init_with_params_no_state(loader, params));
}

// the deprecated spellings (src/qwen2-whisper.cpp:3184-3206): default params
struct whisper_context* whisper_init_from_file(const char* path_model) {
    return whisper_init_from_file_with_params(path_model, whisper_context_default_params());
}
struct whisper_context* whisper_init_from_buffer(void* buffer, size_t buffer_size) {
    return whisper_init_from_buffer_with_params(buffer, buffer_size, whisper_context_default_params());
}
struct whisper_context* whisper_init(struct whisper_model_loader* loader) {
    return whisper_init_with_params(loader, whisper_context_default_params());
}
struct whisper_context* whisper_init_from_file_no_state(const char* path_model) {
    return whisper_init_from_file_with_params_no_state(path_model, whisper_context_default_params());
}
struct whisper_context* whisper_init_from_buffer_no_state(void* buffer, size_t buffer_size) {
    return whisper_init_from_buffer_with_params_no_state(buffer, buffer_size, whisper_context_default_params());
}
struct whisper_context* whisper_init_no_state(struct whisper_model_loader* loader) {
    return whisper_init_with_params_no_state(loader, whisper_context_default_params());
}

// additive: an explicit device list (one weight replica per entry; the same ordinal may appear more than once)
struct whisper_context* whisper_init_from_file_multi(const char* path_model, struct whisper_context_params params, const int* devices, int n_devices) {
    if (!devices || n_devices < 1) return nullptr;
    return with_state(init_from_file_no_state(path_model, params, devices, n_devices, __func__));
}
struct whisper_context* whisper_init_from_buffer_multi(void* buffer, size_t buffer_size, struct whisper_context_params params, const int* devices,
                                                       int n_devices) {
    if (!devices || n_devices < 1) return nullptr;
    return with_state(init_from_buffer_no_state(buffer, buffer_size, params, devices, n_devices, __func__));
}

void whisper_free_state(struct whisper_state* state) {
    if (!state) return;
    if (state->owned) q2w_state_free(state->qs);
    delete state;
}

void whisper_free(struct whisper_context* ctx) {
    if (!ctx) return;
    whisper_free_state(ctx->state);
    if (ctx->multi) q2w_multi_free(ctx->multi);
    free_models(*ctx);
    delete ctx;
}

// ------------------------------------------------------------------------------------------------ mel
int whisper_pcm_to_mel_with_state(struct whisper_context* ctx, struct whisper_state* state, const float* samples, int n_samples, int /*n_threads*/) {
    if (!ctx || !state || q2w_pcm_to_mel(state->qs, samples, n_samples) != Q2W_OK) {
        LOG_ERROR("%s: failed to compute mel spectrogram: %s\n", __func__, q2w_last_error());
        return -1;
    }
    return 0;
}
int whisper_pcm_to_mel(struct whisper_context* ctx, const float* samples, int n_samples, int n_threads) {
    return whisper_pcm_to_mel_with_state(ctx, ctx ? ctx->state : nullptr, samples, n_samples, n_threads);
}

int whisper_set_mel_with_state(struct whisper_context* ctx, struct whisper_state* state, const float* data, int n_len, int n_mel) {
    if (!ctx || !state) return -1;
    if (n_mel != ctx->hp.n_mels) {
        LOG_ERROR("%s: invalid number of mel bands: %d (expected %d)\n", __func__, n_mel, ctx->hp.n_mels);
        return -1;
    }
    if (q2w_set_mel(state->qs, data, n_len, n_mel) != Q2W_OK) {
        LOG_ERROR("%s: %s\n", __func__, q2w_last_error());
        return -1;
    }
    return 0;
}
int whisper_set_mel(struct whisper_context* ctx, const float* data, int n_len, int n_mel) {
    return whisper_set_mel_with_state(ctx, ctx ? ctx->state : nullptr, data, n_len, n_mel);
}

// ------------------------------------------------------------------------------------------------ encode
int whisper_encode_with_state(struct whisper_context* ctx, struct whisper_state* state, int offset, int /*n_threads*/) {
    if (!ctx || !state) return -1;
    if (q2w_encode(state->qs, offset) != Q2W_OK) {
        LOG_ERROR("%s: failed to eval: %s\n", __func__, q2w_last_error());
        return -1;
    }
    return 0;
}
int whisper_encode(struct whisper_context* ctx, int offset, int n_threads) {
    return whisper_encode_with_state(ctx, ctx ? ctx->state : nullptr, offset, n_threads);
}

int whisper_full_with_state(struct whisper_context* ctx, struct whisper_state* state, struct whisper_full_params params, const float* samples,
                            int n_samples) {
    if (!ctx || !state) return -1;
    if (n_samples > 0) {
        if (whisper_pcm_to_mel_with_state(ctx, state, samples, n_samples, params.n_threads) != 0) {
            LOG_ERROR("%s: failed to compute log mel spectrogram\n", __func__);
            return -2;
        }
    }
    const int seek_start = params.offset_ms / 10;
    const int seek_end = params.duration_ms == 0 ? whisper_n_len_from_state(state) : seek_start + params.duration_ms / 10;
    if (seek_end < seek_start + 100) {
        LOG_WARN("%s: input is too short - %d ms < 1000 ms. consider padding the input audio with silence\n", __func__,
                 (seek_end - seek_start) * 10);
        return 0;
    }
    if (q2w_encode(state->qs, seek_start) != Q2W_OK) {
        LOG_ERROR("%s: failed to encode: %s\n", __func__, q2w_last_error());
        return -1;
    }
    if (params.abort_callback && params.abort_callback(params.abort_callback_user_data)) {   // polled once, after the pass (:2338)
        LOG_ERROR("%s: failed to encode\n", __func__);
        return -1;
    }
    return 0;
}
int whisper_full(struct whisper_context* ctx, struct whisper_full_params params, const float* samples, int n_samples) {
    return whisper_full_with_state(ctx, ctx ? ctx->state : nullptr, params, samples, n_samples);
}

// ------------------------------------------------------------------------------------------------ getters
// the reference returns mel.n_len_org here (src/qwen2-whisper.cpp:3440-3446): frames that carry audio, not the 30 s-padded length,
// which is what makes whisper_full skip clips shorter than ~1 s (:2357-2365)
int whisper_n_len_from_state(struct whisper_state* state) { return state ? q2w_mel_n_len_org(state->qs) : 0; }
int whisper_n_len(struct whisper_context* ctx) { return ctx ? whisper_n_len_from_state(ctx->state) : 0; }
int whisper_n_vocab(struct whisper_context* ctx) { return ctx->hp.n_vocab; }
int whisper_n_text_ctx(struct whisper_context* ctx) { return ctx->hp.n_text_ctx; }
int whisper_n_audio_ctx(struct whisper_context* ctx) { return ctx->hp.n_audio_ctx; }
int whisper_model_n_vocab(struct whisper_context* ctx) { return ctx->hp.n_vocab; }
int whisper_model_n_audio_ctx(struct whisper_context* ctx) { return ctx->hp.n_audio_ctx; }
int whisper_model_n_audio_state(struct whisper_context* ctx) { return ctx->hp.n_audio_state; }
int whisper_model_n_audio_head(struct whisper_context* ctx) { return ctx->hp.n_audio_head; }
int whisper_model_n_audio_layer(struct whisper_context* ctx) { return ctx->hp.n_audio_layer; }
int whisper_model_n_text_ctx(struct whisper_context* ctx) { return ctx->hp.n_text_ctx; }
int whisper_model_n_text_state(struct whisper_context* ctx) { return ctx->hp.n_text_state; }
int whisper_model_n_text_head(struct whisper_context* ctx) { return ctx->hp.n_text_head; }
int whisper_model_n_text_layer(struct whisper_context* ctx) { return ctx->hp.n_text_layer; }
int whisper_model_n_mels(struct whisper_context* ctx) { return ctx->hp.n_mels; }
int whisper_model_ftype(struct whisper_context* ctx) { return ctx->ftype; }
int whisper_model_type(struct whisper_context* ctx) { return ctx->model_type; }
const char* whisper_model_type_readable(struct whisper_context* ctx) {
    static const char* names[] = {"unknown", "tiny", "base", "small", "medium", "large"};
    return names[std::max(0, std::min(5, ctx->model_type))];
}

// ------------------------------------------------------------------------------------------------ timings / log
void whisper_print_timings(struct whisper_context* ctx) {
    const int64_t t_end_us = now_us();
    LOG_INFO("\n");
    LOG_INFO("%s:     load time = %8.2f ms\n", __func__, ctx->t_load_us / 1000.0f);
    if (ctx->state) {
        int64_t t_mel = 0, t_enc = 0;
        int32_t n_enc = 0;
        q2w_get_timings(ctx->state->qs, &t_mel, &t_enc, &n_enc);
        const int32_t n = std::max(1, n_enc);
        LOG_INFO("%s:      mel time = %8.2f ms\n", __func__, t_mel / 1000.0f);
        LOG_INFO("%s:   encode time = %8.2f ms / %5d runs (%8.2f ms per run)\n", __func__, 1e-3f * t_enc, n, 1e-3f * t_enc / n);
    }
    LOG_INFO("%s:    total time = %8.2f ms\n", __func__, (t_end_us - ctx->t_start_us) / 1000.0f);
}

void whisper_reset_timings(struct whisper_context* ctx) {
    ctx->t_start_us = now_us();
    if (ctx->state) q2w_reset_timings(ctx->state->qs);
}

const char* whisper_print_system_info(void) {
    static std::string s;
    s = std::string("CUDA = 1 | SM100A = ") + std::to_string(q2w_device_count() > 0 ? 1 : 0) + " | TCGEN05 = 1 | TMA = 1 | CPU_FALLBACK = 0 | " +
        q2w_build_info();
    return s.c_str();
}

void whisper_log_set(ggml_log_callback log_callback, void* user_data) {
    g_log = log_callback ? log_callback : default_log;
    g_log_ud = user_data;
}

void whisper_print_emb_enc(struct whisper_context* ctx) {
    // D2H of the first 20 outputs, " %.3f" each, newline (src:4191-4203)
    float v[20] = {0};
    if (ctx && ctx->state) q2w_get_embeddings(ctx->state->qs, v, 0, 20);
    for (int i = 0; i < 20; ++i) printf(" %.3f", v[i]);
    printf("\n");
}

// ------------------------------------------------------------------------------------------------ additive API
int whisper_embd_dims(struct whisper_context* ctx, int* n_windows, int* n_out, int* n_state) {
    if (!ctx || !ctx->state) return -1;
    return q2w_embd_dims(ctx->state->qs, n_windows, n_out, n_state) == Q2W_OK ? 0 : -1;
}
int whisper_get_embeddings_from_state(struct whisper_state* state, float* dst, size_t n_floats) {
    if (!state) return -1;
    if (q2w_get_embeddings(state->qs, dst, 0, n_floats) != Q2W_OK) {
        LOG_ERROR("%s: %s\n", __func__, q2w_last_error());
        return -1;
    }
    return 0;
}
int whisper_get_embeddings(struct whisper_context* ctx, float* dst, size_t n_floats) {
    return whisper_get_embeddings_from_state(ctx ? ctx->state : nullptr, dst, n_floats);
}
const float* whisper_get_embeddings_device(struct whisper_context* ctx) {
    return (ctx && ctx->state) ? q2w_embeddings_device(ctx->state->qs) : nullptr;
}
int whisper_get_mel(struct whisper_context* ctx, float* dst, size_t n_floats) {
    if (!ctx || !ctx->state) return -1;
    if (q2w_get_mel(ctx->state->qs, dst, n_floats) != Q2W_OK) {
        LOG_ERROR("%s: %s\n", __func__, q2w_last_error());
        return -1;
    }
    return 0;
}
int whisper_get_mel_dims(struct whisper_context* ctx, int* n_len, int* n_len_org, int* n_mel) {
    if (!ctx || !ctx->state) return -1;
    if (n_len) *n_len = q2w_mel_n_len(ctx->state->qs);
    if (n_len_org) *n_len_org = q2w_mel_n_len_org(ctx->state->qs);
    if (n_mel) *n_mel = ctx->hp.n_mels;
    return 0;
}
// windows per micro-batch (per device). The default state's scratch is resized in place: its mel and embeddings survive.
int whisper_set_max_batch(struct whisper_context* ctx, int max_batch) {
    if (!ctx || max_batch < 1) return -1;
    ctx->max_batch = max_batch;            // states created from now on (whisper_init_state) use it too
    int rc = Q2W_OK;
    if (ctx->multi) rc = q2w_multi_set_max_batch(ctx->multi, max_batch);
    else if (ctx->state) rc = q2w_state_set_max_batch(ctx->state->qs, max_batch);
    if (rc != Q2W_OK) {
        LOG_ERROR("%s: %s\n", __func__, q2w_last_error());
        return -1;
    }
    return 0;
}
int whisper_n_devices(struct whisper_context* ctx) { return ctx ? static_cast<int>(ctx->replicas.size()) : 0; }
int whisper_device(struct whisper_context* ctx, int i) { return (ctx && i >= 0 && i < static_cast<int>(ctx->devices.size())) ? ctx->devices[i] : -1; }
// all devices of the context: window w -> device floor(w * G / n_windows), results in caller order in dst (if non-NULL);
// gather_device >= 0 also assembles every embedding on that device (whisper_get_gathered_device). One device: == whisper_encode_batch.
int whisper_encode_batch_multi(struct whisper_context* ctx, const float* samples, size_t stride, const int32_t* n_samples, int n_windows, float* dst,
                               int gather_device) {
    if (!ctx || !ctx->state) return -1;
    if (!ctx->multi) {
        if (gather_device >= 0 && gather_device != ctx->devices[0]) {
            LOG_ERROR("%s: gather device %d is not a device of this context\n", __func__, gather_device);
            return -1;
        }
        return whisper_encode_batch(ctx, samples, stride, n_samples, n_windows, dst);
    }
    if (q2w_multi_encode_batch_host(ctx->multi, samples, stride, n_samples, n_windows, dst, gather_device) != Q2W_OK) {
        LOG_ERROR("%s: failed to encode: %s\n", __func__, q2w_last_error());
        return -1;
    }
    return 0;
}
const float* whisper_get_gathered_device(struct whisper_context* ctx) {
    if (!ctx) return nullptr;
    return ctx->multi ? q2w_multi_gathered_device(ctx->multi) : (ctx->state ? q2w_embeddings_device(ctx->state->qs) : nullptr);
}
void* whisper_q2w_multi(struct whisper_context* ctx) { return ctx ? ctx->multi : nullptr; }
int whisper_encode_batch(struct whisper_context* ctx, const float* samples, size_t stride, const int32_t* n_samples, int n_windows, float* dst) {
    if (!ctx || !ctx->state) return -1;
    if (ctx->multi) return whisper_encode_batch_multi(ctx, samples, stride, n_samples, n_windows, dst, -1);
    if (q2w_encode_batch_host(ctx->state->qs, samples, stride, n_samples, n_windows, dst) != Q2W_OK) {
        LOG_ERROR("%s: failed to encode: %s\n", __func__, q2w_last_error());
        return -1;
    }
    return 0;
}
int whisper_encode_batch_async(struct whisper_context* ctx, const float* samples, size_t stride, const int32_t* n_samples, int n_windows, float* dst) {
    if (!ctx || !ctx->state) return -1;
    int ticket = -1;
    if (ctx->multi) {   // every device queues its shard
        if (q2w_multi_encode_batch_host_async(ctx->multi, samples, stride, n_samples, n_windows, dst, &ticket) != Q2W_OK) {
            LOG_ERROR("%s: failed to queue the batch: %s\n", __func__, q2w_last_error());
            return -1;
        }
        return ticket;
    }
    if (q2w_encode_batch_host_async(ctx->state->qs, samples, stride, n_samples, n_windows, dst, &ticket) != Q2W_OK) {
        LOG_ERROR("%s: failed to queue the batch: %s\n", __func__, q2w_last_error());
        return -1;
    }
    return ticket;
}
int whisper_encode_batch_wait(struct whisper_context* ctx, int ticket) {
    if (!ctx || !ctx->state) return -1;
    if (ctx->multi) {
        if (q2w_multi_encode_batch_wait(ctx->multi, ticket) != Q2W_OK) {
            LOG_ERROR("%s: %s\n", __func__, q2w_last_error());
            return -1;
        }
        return 0;
    }
    if (q2w_encode_batch_wait(ctx->state->qs, ticket) != Q2W_OK) {
        LOG_ERROR("%s: %s\n", __func__, q2w_last_error());
        return -1;
    }
    return 0;
}
int whisper_encode_batch_device(struct whisper_context* ctx, const float* samples_dev, size_t stride, const int32_t* n_samples, int n_windows) {
    if (!ctx || !ctx->state) return -1;
    if (q2w_encode_batch_device(ctx->state->qs, samples_dev, stride, n_samples, n_windows) != Q2W_OK) {
        LOG_ERROR("%s: failed to encode: %s\n", __func__, q2w_last_error());
        return -1;
    }
    return 0;
}
int whisper_encode_offsets(struct whisper_context* ctx, const int32_t* mel_offsets, int n_windows, float* dst) {
    if (!ctx || !ctx->state) return -1;
    if (q2w_encode_offsets(ctx->state->qs, mel_offsets, n_windows, dst) != Q2W_OK) {
        LOG_ERROR("%s: failed to encode: %s\n", __func__, q2w_last_error());
        return -1;
    }
    return 0;
}
// multi-modal projector (additive): uploaded to every replica; whisper_project runs it on the default state's last embeddings
int whisper_set_projector(struct whisper_context* ctx, int ggml_type, int n_out, const void* weight, size_t nbytes, const float* bias) {
    if (!ctx) return -1;
    for (q2w_model* m : ctx->replicas) {
        if (q2w_model_upload_projector(m, ggml_type, n_out, weight, nbytes, bias) != Q2W_OK) {
            LOG_ERROR("%s: %s\n", __func__, q2w_last_error());
            return -1;
        }
    }
    return 0;
}
int whisper_project(struct whisper_context* ctx, float* dst, size_t n_floats) {
    if (!ctx || !ctx->state) return -1;
    if (q2w_project(ctx->state->qs, dst, n_floats) != Q2W_OK) {
        LOG_ERROR("%s: %s\n", __func__, q2w_last_error());
        return -1;
    }
    return 0;
}
int whisper_projection_dims(struct whisper_context* ctx, int* n_rows, int* n_out) {
    if (!ctx || !ctx->state) return -1;
    return q2w_projection_dims(ctx->state->qs, n_rows, n_out) == Q2W_OK ? 0 : -1;
}
void* whisper_q2w_state(struct whisper_context* ctx) { return (ctx && ctx->state) ? ctx->state->qs : nullptr; }

}  // extern "C"
