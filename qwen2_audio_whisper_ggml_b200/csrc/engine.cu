// engine.cu -- the C-ABI of include/q2w_b200.h: model upload, state/workspaces, and the forward pass
// PCM -> log-mel -> conv stem -> 32 pre-LN blocks -> avg-pool(2) -> final LayerNorm, as a sequence of sm_100a kernels
// on one stream.  This file replaces what src/qwen2-whisper.cpp did through ggml_backend_sched_* (graph build,
// alloc, split, compute): there is no graph IR, no allocator and no backend dispatch -- workspaces are static per
// micro-batch and the op order below IS the encoder (Appendix C of SURVEY.md; reference :1892-1952, :1954-2203).
#include "../../include/q2w_b200.h"
#include "ops.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <mutex>
#include <vector>

using namespace q2w;

namespace {

thread_local char g_err[512] = "";
std::atomic<long> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace

int q2w::set_last_error(int code, const char* msg) { return fail(code, "%s", msg ? msg : ""); }

namespace {

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(Q2W_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

// a kernel launch wrapper result: count it, then check
#define CKL(call)                                                                                      \
    do {                                                                                               \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                            \
        CK(call);                                                                                      \
    } while (0)

int64_t now_us() {
    return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

size_t type_row_bytes(int type, int64_t ne0) {
    switch (type) {
        case Q2W_TYPE_F32: return static_cast<size_t>(ne0) * 4;
        case Q2W_TYPE_F16: return static_cast<size_t>(ne0) * 2;
        case Q2W_TYPE_Q8_0: return static_cast<size_t>(ne0 / 32) * 34;   // block_q8_0 ggml-common.h:186-191
        case Q2W_TYPE_Q4_0: return static_cast<size_t>(ne0 / 32) * 18;   // block_q4_0 ggml-common.h:144-148
        default: return 0;
    }
}

struct Tensor {
    void* d = nullptr;        // device storage (type dev_type)
    int file_type = 0;        // ggml type the model file must carry
    int dev_type = 0;         // type kept on the device (F32 matrices/conv kernels are converted to F16 once)
    int n_dims = 1;
    int64_t ne[3] = {1, 1, 1};
    bool loaded = false;
    bool owns = true;
    int64_t nelements() const { return ne[0] * ne[1] * ne[2]; }
    int64_t nrows() const { return ne[1] * ne[2]; }
    size_t file_bytes() const { return type_row_bytes(file_type, ne[0]) * static_cast<size_t>(nrows()); }
    size_t dev_bytes() const { return type_row_bytes(dev_type, ne[0]) * static_cast<size_t>(nrows()); }
};

struct Layer {
    Tensor ln1_w, ln1_b, q_w, q_b, k_w, v_w, v_b, o_w, o_b, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b;
    void* qkv_w = nullptr;    // q_w | k_w | v_w rows, contiguous (ggml row layout keeps blocks row-local)
    float* qkv_b = nullptr;   // q_b | 0 | v_b
};

}  // namespace

struct q2w_model {
    q2w_hparams hp{};
    int wtype = Q2W_TYPE_F16;      // file type of the 2-D matrices
    int wtype_dev = Q2W_TYPE_F16;  // what the GEMM path sees (F16 / Q8_0 / Q4_0)
    int device = 0;
    Tensor pe, conv1_w, conv1_b, conv2_w, conv2_b, ln_w, ln_b;
    std::vector<Layer> layers;
    std::map<std::string, Tensor*> by_name;
    MelPlan* mel = nullptr;
    // optional step AFTER the path: Qwen2-Audio's multi_modal_projector, one Linear(n_audio_state -> n_out) with bias
    __half* proj_w = nullptr;   // [n_out][n_audio_state] f16
    float* proj_b = nullptr;    // [n_out]
    int proj_out = 0;
    int n_loaded = 0;
    bool finalized = false;
    size_t weight_bytes = 0;
    cudaStream_t stream = nullptr;
};

struct q2w_state {
    q2w_model* m = nullptr;
    int max_batch = 1;
    cudaStream_t stream = nullptr;
    // dims
    int T = 0, T2 = 0, D = 0, H = 0, FF = 0, n_mel = 0, win_samples = 0, ld_mel = 0, n_frames_batch = 0;
    // workspaces (sized for max_batch windows)
    float* x = nullptr;        // residual stream f32 [B*T, D]
    __half* ln = nullptr;      // LayerNorm out f16 [B*T, D]
    __half* qkv = nullptr;     // [B*T, 3D]        (aliases conv2 operand A2)
    __half* att = nullptr;     // [B*T, D]         (aliases conv1 operand A1)
    __half* h = nullptr;       // [B*T, 4D]        (aliases conv1 output h1 [B*T2, D])
    // quantised weights: the F16 copy of ONE encoder block (QKV | out | fc1 | fc2, 39 MB), rewritten block after block by the decode
    // warps that ride inside the attention kernel
    __half* wlayer = nullptr;
    float* pcm_dev = nullptr;  // [2][B, win_samples]  double-buffered host staging (copy of micro-batch i+1 overlaps compute of i)
    int* nsamp_dev = nullptr;  // [2][B]
    cudaStream_t s_in = nullptr, s_out = nullptr;   // copy-in / copy-out streams of the host-buffer batch path
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    float* logmel = nullptr;   // [B, n_mel, ld_mel]
    float* winmax = nullptr;   // [B] ordered int keys
    int* att_sched = nullptr;  // [2] work counter of the persistent attention kernel (zero between launches)
    bool att_sched_dirty = false;   // a forward pass failed part-way: re-zero the counter before the next launch
    int dbg_layers = -1;       // q2w_debug_forward_layers: stop after this many encoder blocks (-1 = all); eager launches only
    int dbg_windows = 0;       // windows of the last forward still resident in x
    int fused_dequant = 0;     // Q2W_FUSED_DEQUANT=1: decode inside the GEMM (A/B); default: decode warps riding in the attention kernel
    int e2e_split = 2;         // Q2W_E2E_SPLIT: micro-batches a synchronous host batch is cut into
    // asynchronous host batches: at most two in flight; ticket t owns embedding region t & 1 and completion event ev_ticket[t & 1]
    cudaEvent_t ev_ticket[2] = {nullptr, nullptr};
    int ticket_B[2] = {0, 0};
    int64_t ticket_t0[2] = {0, 0};
    bool ticket_live[2] = {false, false};
    int next_ticket = 0;
    size_t emb_off_windows = 0;     // where the embeddings of the last completed call start inside emb
    uint64_t mb_seq = 0;            // running micro-batch index: staging slot = mb_seq & 1, across calls
    // results
    float* emb = nullptr;      // [n_windows, T/2, D]
    __half* emb16 = nullptr;   // the same rows in F16 (only when the model carries a projector: written by the pool + LN tail)
    float* proj = nullptr;     // [n_windows * T/2, proj_out] f32, q2w_project
    size_t proj_cap_rows = 0;  // capacity of proj in floats
    int proj_rows = 0;
    size_t emb_cap_windows = 0;
    int emb_windows = 0;
    // API mel (whisper_pcm_to_mel / whisper_set_mel)
    float* api_mel = nullptr;
    size_t api_mel_cap = 0;
    int api_n_len = 0, api_n_len_org = 0, api_ld = 0;
    float* api_pcm = nullptr;
    size_t api_pcm_cap = 0;
    float* api_max = nullptr;
    // CUDA graph of the single-window forward (launch-bound: ~330 kernels of 5-25 us each); captured on the second B = 1 call
    cudaGraphExec_t g1 = nullptr;
    float* g1_emb = nullptr;
    int g1_state = 0;          // 0 cold, 1 warmed eagerly, 2 graph ready, -1 capture failed (stay eager)
    // timers
    int64_t t_mel_us = 0, t_encode_us = 0;
    int32_t n_encode = 0;
    // optional per-kernel-class CUDA-event timing (bench.py roofline): records live on the stream the kernels run on
    struct ProfRec { int cls; double flops; double bytes; cudaEvent_t a, b; };
    bool prof_on = false;
    std::vector<ProfRec> prof;
    size_t prof_used = 0;
};

namespace {

int alloc_tensor(Tensor& t, int file_type, int dev_type, int n_dims, int64_t ne0, int64_t ne1, int64_t ne2) {
    t.file_type = file_type;
    t.dev_type = dev_type;
    t.n_dims = n_dims;
    t.ne[0] = ne0; t.ne[1] = ne1; t.ne[2] = ne2;
    if (t.owns) CK(cudaMalloc(&t.d, t.dev_bytes()));
    return Q2W_OK;
}

void free_tensor(Tensor& t) {
    if (t.owns && t.d) cudaFree(t.d);
    t.d = nullptr;
}

int check_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(Q2W_E_NO_DEVICE, "no CUDA device visible: libq2w_b200 has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(Q2W_E_INVALID, "device ordinal %d out of range (0..%d)", device, n - 1);
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) return fail(Q2W_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
    return Q2W_OK;
}

struct ProfScope {
    q2w_state* s;
    cudaStream_t st;
    long idx = -1;   // index, not pointer: the record vector may grow while a scope is open
    ProfScope(q2w_state* s_, int cls, double flops, double bytes, cudaStream_t st_ = nullptr) : s(s_), st(st_ ? st_ : s_->stream) {
        if (!s->prof_on) return;
        if (s->prof_used == s->prof.size()) {
            q2w_state::ProfRec n{};
            if (cudaEventCreate(&n.a) != cudaSuccess || cudaEventCreate(&n.b) != cudaSuccess) return;
            s->prof.push_back(n);
        }
        idx = static_cast<long>(s->prof_used++);
        q2w_state::ProfRec& r = s->prof[idx];
        r.cls = cls; r.flops = flops; r.bytes = bytes;
        cudaEventRecord(r.a, st);
    }
    ~ProfScope() { if (idx >= 0) cudaEventRecord(s->prof[idx].b, st); }
};
enum { PC_GEMM = 0, PC_ATTN = 1, PC_LN = 2, PC_MEL = 3, PC_IM2COL = 4, PC_DEQUANT = 5, PC_COUNT = 6 };

// y = A x W^T.  wtype: F16 (TMA-fed W) or Q8_0 / Q4_0 raw ggml blocks (decoded by the GEMM's own dequant warpgroup).
// w_static: W is a model tensor no kernel writes (its first tiles may be fetched before the previous kernel has finished).
int weight_gemm(q2w_state* s, const __half* A, int lda, const void* W, int wtype, bool w_static, int M, int N, int K, const float* bias,
                void* out, int ldo, GemmEpilogue epi, const float* resid, const float* pos, int pos_period, int scale_cols,
                float scale) {
    ProfScope ps(s, PC_GEMM, 2.0 * M * static_cast<double>(N) * K, 0.0);
    GemmArgs g{};
    g.A = A; g.lda = lda; g.W = static_cast<const __half*>(W); g.ldw = K; g.wtype = wtype; g.M = M; g.N = N; g.K = K; g.bias = bias; g.out = out; g.ldo = ldo;
    g.resid = resid; g.pos = pos; g.pos_period = pos_period; g.scale_cols = scale_cols; g.scale = scale;
    g.w_static = w_static ? 1 : 0;
    CKL(gemm_f16_tcgen05(g, epi, s->stream));
    return Q2W_OK;
}

// decode job over the layer buffer: entries (matrix k of block il) -> its slot; k: 0 QKV, 1 out, 2 fc1, 3 fc2
void add_decode(q2w_state* s, DequantJob& job, int slot, int il, int k) {
    Layer& L = s->m->layers[il];
    const size_t D = s->D, FF = s->FF;
    const void* src[4] = {L.qkv_w, L.o_w.d, L.fc1_w.d, L.fc2_w.d};
    const size_t elems[4] = {3 * D * D, D * D, FF * D, D * FF};
    const size_t off[4] = {0, 3 * D * D, 4 * D * D, 4 * D * D + FF * D};
    job.src[slot] = static_cast<const uint8_t*>(src[k]);
    job.dst[slot] = s->wlayer + off[k];
    job.nblocks[slot] = elems[k] / 32;
}

double decode_bytes(const q2w_state* s, const DequantJob& job) {
    double b = 0;
    for (int i = 0; i < 4; ++i) b += static_cast<double>(job.nblocks[i]) * ((s->m->wtype_dev == Q2W_TYPE_Q8_0 ? 34 : 18) + 64);
    return b;
}

int forward_eager(q2w_state* s, int Bm, int w0);
int forward_dispatch(q2w_state* s, int Bm, int w0);

// conv stem + encoder for Bm windows whose conv1 operand A1 (in s->att) is ready; writes emb rows [w0, w0+Bm)
int forward_from_a1(q2w_state* s, int Bm, int w0) {
    if (s->att_sched_dirty) {
        // a previous pass failed between launches: the attention work counter may be non-zero (the kernel re-zeroes it only when it
        // runs to completion), which would silently mis-schedule every later launch
        CK(cudaMemsetAsync(s->att_sched, 0, 2 * sizeof(int), s->stream));
        s->att_sched_dirty = false;
    }
    const int rc = forward_dispatch(s, Bm, w0);
    if (rc != Q2W_OK) s->att_sched_dirty = true;
    else s->dbg_windows = Bm;
    return rc;
}

int forward_dispatch(q2w_state* s, int Bm, int w0) {
    if (Bm != 1 || w0 != 0 || s->prof_on || s->g1_state < 0 || s->dbg_layers >= 0) return forward_eager(s, Bm, w0);
    if (s->g1_state == 2 && s->g1_emb == s->emb) {
        CK(cudaGraphLaunch(s->g1, s->stream));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return Q2W_OK;
    }
    if (s->g1_state == 0) {            // first call runs eagerly: one-time function attributes / driver entry points get resolved
        s->g1_state = 1;
        return forward_eager(s, Bm, w0);
    }
    if (s->g1) { cudaGraphExecDestroy(s->g1); s->g1 = nullptr; }
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); s->g1_state = -1; return forward_eager(s, Bm, w0); }
    const int rc = forward_eager(s, Bm, w0);
    const cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
    if (rc != Q2W_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&s->g1, graph, 0) != cudaSuccess) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        s->g1 = nullptr;
        s->g1_state = -1;
        return forward_eager(s, Bm, w0);
    }
    cudaGraphDestroy(graph);
    s->g1_state = 2;
    s->g1_emb = s->emb;
    CK(cudaGraphLaunch(s->g1, s->stream));
    return Q2W_OK;
}

int forward_eager(q2w_state* s, int Bm, int w0) {
    q2w_model* m = s->m;
    const int T = s->T, T2 = s->T2, D = s->D, H = s->H, FF = s->FF;
    const int M = Bm * T;
    const float eps = 1e-5f;  // hparams.eps, src/qwen2-whisper.cpp:579
    __half* A1 = s->att;
    __half* h1 = s->h;
    __half* A2 = s->qkv;
    int rc;
    // conv1 (k3 s1 p1) + bias + GELU   (:1922-1925)
    if ((rc = weight_gemm(s, A1, 3 * s->n_mel, m->conv1_w.d, Q2W_TYPE_F16, true, Bm * T2, D, 3 * s->n_mel,
                          static_cast<const float*>(m->conv1_b.d), h1, D, EPI_BIAS_GELU_F16, nullptr, nullptr, 0, 0, 1.f)))
        return rc;
    // conv2 (k3 s2 p1) + bias + GELU, transposed to time-major and + positional embedding   (:1927-1930, :2001-2005)
    {
        ProfScope ps(s, PC_IM2COL, 0.0, 2.0 * Bm * (static_cast<double>(T2) * D + static_cast<double>(T) * 3 * D));
        CKL(conv2_im2col(h1, A2, Bm, T2, D, s->stream));
    }
    if ((rc = weight_gemm(s, A2, 3 * D, m->conv2_w.d, Q2W_TYPE_F16, true, M, D, 3 * D, static_cast<const float*>(m->conv2_b.d),
                          s->x, D, EPI_BIAS_GELU_POS_F32, nullptr, static_cast<const float*>(m->pe.d), T, 0, 1.f)))
        return rc;
    const float kq_scale = 1.0f / sqrtf(static_cast<float>(D / H));  // :1985
    const int n_layers = s->dbg_layers >= 0 ? std::min(s->dbg_layers, m->hp.n_audio_layer) : m->hp.n_audio_layer;
    // Quantised weights stay Q8_0 / Q4_0 in HBM exactly as in the model file. Two decode strategies:
    //   rider   (default) the F16 copy of one encoder block lives in a 39 MB layer buffer and is rewritten block after block by the two
    //           otherwise idle warps of every attention CTA: while attention(l) runs they decode out / fc1 / fc2 of block l (read right
    //           after it) and QKV of block l + 1. No extra launches, no second stream; the GEMMs are the plain TMA-fed F16 kernels
    //   fused   (Q2W_FUSED_DEQUANT=1) raw ggml blocks go straight into the GEMM and are decoded by its dequant warpgroup: every W tile
    //           is re-decoded by each M-tile that uses it -- measured 2.7x slower at M = 96000 and 1.7x slower at M = 1500
    const bool quant = m->wtype_dev != Q2W_TYPE_F16;
    const bool fused = quant && s->fused_dequant > 0 && D % 64 == 0;
    const bool by_layer = quant && !fused;
    const size_t w_off[4] = {0, static_cast<size_t>(3) * D * D, static_cast<size_t>(4) * D * D, static_cast<size_t>(4) * D * D + static_cast<size_t>(FF) * D};
    if (by_layer && n_layers > 0) {
        DequantJob job{};                                   // block 0's QKV has no attention kernel before it: one stand-alone launch
        add_decode(s, job, 0, 0, 0);
        ProfScope ps(s, PC_DEQUANT, 0.0, decode_bytes(s, job));
        CKL(dequant_multi_to_f16(job, m->wtype_dev, s->stream));
    }
    for (int il = 0; il < n_layers; ++il) {
        Layer& L = m->layers[il];
        const void* w_qkv = L.qkv_w; const void* w_o = L.o_w.d; const void* w_fc1 = L.fc1_w.d; const void* w_fc2 = L.fc2_w.d;
        int wt = m->wtype_dev;
        bool w_static = true;
        DequantJob job{};
        if (by_layer) {
            add_decode(s, job, 0, il, 1);
            add_decode(s, job, 1, il, 2);
            add_decode(s, job, 2, il, 3);
            if (il + 1 < n_layers) add_decode(s, job, 3, il + 1, 0);
            const __half* base = s->wlayer;
            w_qkv = base + w_off[0]; w_o = base + w_off[1]; w_fc1 = base + w_off[2]; w_fc2 = base + w_off[3];
            wt = Q2W_TYPE_F16;
            w_static = false;                                  // written by a kernel of this stream: never fetch it early
        }
        // pre-LN + fused QKV projection (+bias, Q * KQscale)   (:2019-2055)
        {
            ProfScope ps(s, PC_LN, 0.0, 6.0 * M * D);
            CKL(layernorm_f32_to_f16(s->x, static_cast<const float*>(L.ln1_w.d), static_cast<const float*>(L.ln1_b.d), s->ln, M, D,
                                     eps, s->stream));
        }
        if ((rc = weight_gemm(s, s->ln, D, w_qkv, wt, w_static, M, 3 * D, D, L.qkv_b, s->qkv, 3 * D, EPI_BIAS_F16, nullptr,
                              nullptr, 0, D, kq_scale)))
            return rc;
        // softmax(Q K^T) V per head   (:2080-2106)
        {
            ProfScope ps(s, PC_ATTN, 4.0 * Bm * static_cast<double>(T) * T * D, 8.0 * M * D);
            CKL(attention_f16_tcgen05(s->qkv, s->att, Bm, T, H, s->att_sched, s->stream, by_layer ? &job : nullptr, by_layer ? m->wtype_dev : 0));
        }
        // out-proj + bias + residual   (:2112-2120)
        if ((rc = weight_gemm(s, s->att, D, w_o, wt, w_static, M, D, D, static_cast<const float*>(L.o_b.d), s->x, D,
                              EPI_BIAS_RESID_F32, s->x, nullptr, 0, 0, 1.f)))
            return rc;
        // MLP: LN, fc1 + GELU, fc2 + residual   (:2128-2154)
        {
            ProfScope ps(s, PC_LN, 0.0, 6.0 * M * D);
            CKL(layernorm_f32_to_f16(s->x, static_cast<const float*>(L.ln2_w.d), static_cast<const float*>(L.ln2_b.d), s->ln, M, D,
                                     eps, s->stream));
        }
        if ((rc = weight_gemm(s, s->ln, D, w_fc1, wt, w_static, M, FF, D, static_cast<const float*>(L.fc1_b.d), s->h, FF,
                              EPI_BIAS_GELU_F16, nullptr, nullptr, 0, 0, 1.f)))
            return rc;
        if ((rc = weight_gemm(s, s->h, FF, w_fc2, wt, w_static, M, D, FF, static_cast<const float*>(L.fc2_b.d), s->x, D,
                              EPI_BIAS_RESID_F32, s->x, nullptr, 0, 0, 1.f)))
            return rc;
    }
    // avg-pool(2,2) over time + final LayerNorm   (:2160-2181)
    float* out = s->emb + static_cast<size_t>(w0) * (T / 2) * D;
    __half* out16 = s->emb16 ? s->emb16 + static_cast<size_t>(w0) * (T / 2) * D : nullptr;
    ProfScope ps(s, PC_LN, 0.0, 6.0 * M * D);
    CKL(pool2_layernorm_f32(s->x, static_cast<const float*>(m->ln_w.d), static_cast<const float*>(m->ln_b.d), out, Bm, T, D, eps,
                            s->stream, out16));
    return Q2W_OK;
}

int ensure_emb(q2w_state* s, int n_windows) {
    const bool want16 = s->m->proj_w != nullptr;
    if (static_cast<size_t>(n_windows) > s->emb_cap_windows || (want16 && !s->emb16)) {
        const size_t cap = std::max(static_cast<size_t>(n_windows), s->emb_cap_windows);
        if (s->emb) cudaFree(s->emb);
        if (s->emb16) cudaFree(s->emb16);
        s->emb = nullptr;
        s->emb16 = nullptr;
        s->emb_cap_windows = 0;
        if (s->g1) { cudaGraphExecDestroy(s->g1); s->g1 = nullptr; }     // the captured single-window graph wrote into the old buffers
        if (s->g1_state == 2) s->g1_state = 1;
        CK(cudaMalloc(&s->emb, cap * (s->T / 2) * s->D * sizeof(float)));
        if (want16) CK(cudaMalloc(&s->emb16, cap * (s->T / 2) * s->D * sizeof(__half)));
        s->emb_cap_windows = cap;
    }
    return Q2W_OK;
}

int ticket_wait(q2w_state* s, int ticket) {
    const int r = ticket & 1;
    if (ticket < 0 || !s->ticket_live[r] || (s->next_ticket - ticket) > 2 || ticket >= s->next_ticket)
        return fail(Q2W_E_INVALID, "ticket %d is not in flight", ticket);
    CK(cudaSetDevice(s->m->device));
    CK(cudaEventSynchronize(s->ev_ticket[r]));
    s->ticket_live[r] = false;
    s->emb_windows = s->ticket_B[r];
    s->emb_off_windows = static_cast<size_t>(r) * s->emb_cap_windows / 2;
    s->t_encode_us += now_us() - s->ticket_t0[r];
    s->n_encode += s->ticket_B[r];
    return Q2W_OK;
}

// waits for every asynchronous batch still in flight
int drain_tickets(q2w_state* s) {
    for (int r = 0; r < 2; ++r)
        if (s->ticket_live[r]) {
            const int t = (s->next_ticket - 1 - ((s->next_ticket - 1 - r) & 1));
            int rc = ticket_wait(s, t);
            if (rc) return rc;
        }
    return Q2W_OK;
}

// mel + conv1 operand for Bm windows resident in s->pcm_dev, then the encoder
// nsamp_dev == nullptr: every window of the chunk has n_max valid samples (no per-window length array needed on the device)
int batch_chunk(q2w_state* s, const float* pcm_dev, size_t stride, const int* nsamp_dev, int n_max, int Bm, int w0) {
    {
        // algorithmic bytes: PCM in + used mel frames out (4*480000 + 4*128*3000 per window)
        ProfScope ps(s, PC_MEL, 0.0, static_cast<double>(Bm) * (4.0 * s->win_samples + 4.0 * s->n_mel * s->T2));
        CKL(mel_logpower(s->m->mel, pcm_dev, stride, nsamp_dev, n_max, Bm, s->n_frames_batch, s->logmel, s->ld_mel,
                         s->winmax, s->stream));
        g_launches.fetch_add(1);  // mel_logpower issues two kernels (key init + main)
    }
    {
        ProfScope ps(s, PC_IM2COL, 0.0, static_cast<double>(Bm) * s->T2 * s->n_mel * (4.0 + 6.0));
        CKL(mel_to_conv1_operand(s->logmel, s->ld_mel, s->n_frames_batch, s->n_mel, s->winmax, 1, 0, s->T2, Bm, s->att, s->stream));
    }
    return forward_from_a1(s, Bm, w0);
}

}  // namespace

// =============================================================================================== model
extern "C" {

int q2w_model_create(q2w_model** out, const q2w_hparams* hp, int wtype, int device) {
    if (!out || !hp) return fail(Q2W_E_INVALID, "null argument");
    *out = nullptr;
    int rc = check_device(device);
    if (rc) return rc;
    if (wtype != Q2W_TYPE_F32 && wtype != Q2W_TYPE_F16 && wtype != Q2W_TYPE_Q8_0 && wtype != Q2W_TYPE_Q4_0)
        return fail(Q2W_E_UNSUPPORTED, "weight type %d not on this path (F32/F16/Q8_0/Q4_0 only)", wtype);
    const int D = hp->n_audio_state, H = hp->n_audio_head, L = hp->n_audio_layer, T = hp->n_audio_ctx, NM = hp->n_mels;
    if (D <= 0 || H <= 0 || L <= 0 || T <= 0 || NM <= 0) return fail(Q2W_E_INVALID, "bad hparams");
    if (D % H || D / H != 64) return fail(Q2W_E_UNSUPPORTED, "head_dim %d unsupported (attention kernel is specialised for 64)", H ? D / H : 0);
    if (D % 32 || D > 1280 || (T & 1) || (3 * NM) % 8)
        return fail(Q2W_E_UNSUPPORTED, "hparams outside the supported envelope (n_audio_state %% 32, <= 1280; even n_audio_ctx)");
    CK(cudaSetDevice(device));
    q2w_model* m = new q2w_model();
    m->hp = *hp;
    m->wtype = wtype;
    m->wtype_dev = (wtype == Q2W_TYPE_F32) ? Q2W_TYPE_F16 : wtype;
    m->device = device;
    m->layers.resize(L);
    const int vtype = (wtype == Q2W_TYPE_F32) ? Q2W_TYPE_F32 : Q2W_TYPE_F16;  // conv kernel type, :1543
    const int FF = 4 * D;
#define ALLOC(t, ft, dt, nd, a, b, c) do { if ((rc = alloc_tensor(t, ft, dt, nd, a, b, c))) { q2w_model_free(m); return rc; } } while (0)
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) { q2w_model_free(m); return fail(Q2W_E_CUDA, "stream create failed"); }
    // shapes in ggml order (innermost first), :1591-1638
    ALLOC(m->pe, Q2W_TYPE_F32, Q2W_TYPE_F32, 2, D, T, 1);
    ALLOC(m->conv1_w, vtype, Q2W_TYPE_F16, 3, 3, NM, D);
    ALLOC(m->conv1_b, Q2W_TYPE_F32, Q2W_TYPE_F32, 2, 1, D, 1);
    ALLOC(m->conv2_w, vtype, Q2W_TYPE_F16, 3, 3, D, D);
    ALLOC(m->conv2_b, Q2W_TYPE_F32, Q2W_TYPE_F32, 2, 1, D, 1);
    ALLOC(m->ln_w, Q2W_TYPE_F32, Q2W_TYPE_F32, 1, D, 1, 1);
    ALLOC(m->ln_b, Q2W_TYPE_F32, Q2W_TYPE_F32, 1, D, 1, 1);
    m->by_name["embed_positions.weight"] = &m->pe;
    m->by_name["conv1.weight"] = &m->conv1_w;
    m->by_name["conv1.bias"] = &m->conv1_b;
    m->by_name["conv2.weight"] = &m->conv2_w;
    m->by_name["conv2.bias"] = &m->conv2_b;
    m->by_name["layer_norm.weight"] = &m->ln_w;
    m->by_name["layer_norm.bias"] = &m->ln_b;
    const size_t wrow = type_row_bytes(m->wtype_dev, D);
    for (int i = 0; i < L; ++i) {
        Layer& ly = m->layers[i];
        if (cudaMalloc(&ly.qkv_w, wrow * 3 * D) != cudaSuccess || cudaMalloc(&ly.qkv_b, sizeof(float) * 3 * D) != cudaSuccess) {
            q2w_model_free(m);
            return fail(Q2W_E_NOMEM, "cudaMalloc failed for layer %d", i);
        }
        cudaMemsetAsync(ly.qkv_b, 0, sizeof(float) * 3 * D, m->stream);
        ly.q_w.owns = ly.k_w.owns = ly.v_w.owns = false;
        ly.q_b.owns = ly.v_b.owns = false;
        ly.q_w.d = ly.qkv_w;
        ly.k_w.d = static_cast<uint8_t*>(ly.qkv_w) + wrow * D;
        ly.v_w.d = static_cast<uint8_t*>(ly.qkv_w) + wrow * 2 * D;
        ly.q_b.d = ly.qkv_b;
        ly.v_b.d = ly.qkv_b + 2 * D;
        ALLOC(ly.ln1_w, Q2W_TYPE_F32, Q2W_TYPE_F32, 1, D, 1, 1);
        ALLOC(ly.ln1_b, Q2W_TYPE_F32, Q2W_TYPE_F32, 1, D, 1, 1);
        ALLOC(ly.q_w, wtype, m->wtype_dev, 2, D, D, 1);
        ALLOC(ly.q_b, Q2W_TYPE_F32, Q2W_TYPE_F32, 1, D, 1, 1);
        ALLOC(ly.k_w, wtype, m->wtype_dev, 2, D, D, 1);
        ALLOC(ly.v_w, wtype, m->wtype_dev, 2, D, D, 1);
        ALLOC(ly.v_b, Q2W_TYPE_F32, Q2W_TYPE_F32, 1, D, 1, 1);
        ALLOC(ly.o_w, wtype, m->wtype_dev, 2, D, D, 1);
        ALLOC(ly.o_b, Q2W_TYPE_F32, Q2W_TYPE_F32, 1, D, 1, 1);
        ALLOC(ly.ln2_w, Q2W_TYPE_F32, Q2W_TYPE_F32, 1, D, 1, 1);
        ALLOC(ly.ln2_b, Q2W_TYPE_F32, Q2W_TYPE_F32, 1, D, 1, 1);
        ALLOC(ly.fc1_w, wtype, m->wtype_dev, 2, D, FF, 1);
        ALLOC(ly.fc1_b, Q2W_TYPE_F32, Q2W_TYPE_F32, 1, FF, 1, 1);
        ALLOC(ly.fc2_w, wtype, m->wtype_dev, 2, FF, D, 1);
        ALLOC(ly.fc2_b, Q2W_TYPE_F32, Q2W_TYPE_F32, 1, D, 1, 1);
        const std::string p = "layers." + std::to_string(i) + ".";
        m->by_name[p + "self_attn_layer_norm.weight"] = &ly.ln1_w;
        m->by_name[p + "self_attn_layer_norm.bias"] = &ly.ln1_b;
        m->by_name[p + "self_attn.q_proj.weight"] = &ly.q_w;
        m->by_name[p + "self_attn.q_proj.bias"] = &ly.q_b;
        m->by_name[p + "self_attn.k_proj.weight"] = &ly.k_w;
        m->by_name[p + "self_attn.v_proj.weight"] = &ly.v_w;
        m->by_name[p + "self_attn.v_proj.bias"] = &ly.v_b;
        m->by_name[p + "self_attn.out_proj.weight"] = &ly.o_w;
        m->by_name[p + "self_attn.out_proj.bias"] = &ly.o_b;
        m->by_name[p + "final_layer_norm.weight"] = &ly.ln2_w;
        m->by_name[p + "final_layer_norm.bias"] = &ly.ln2_b;
        m->by_name[p + "fc1.weight"] = &ly.fc1_w;
        m->by_name[p + "fc1.bias"] = &ly.fc1_b;
        m->by_name[p + "fc2.weight"] = &ly.fc2_w;
        m->by_name[p + "fc2.bias"] = &ly.fc2_b;
    }
#undef ALLOC
    for (auto& kv : m->by_name) m->weight_bytes += kv.second->dev_bytes();
    *out = m;
    return Q2W_OK;
}

int q2w_model_upload_filters(q2w_model* m, const float* filters, int n_mel, int n_fft) {
    if (!m || !filters) return fail(Q2W_E_INVALID, "null argument");
    if (n_mel != m->hp.n_mels) return fail(Q2W_E_BAD_SHAPE, "filterbank has %d mel bands, model expects %d", n_mel, m->hp.n_mels);
    if (n_fft != 201) return fail(Q2W_E_BAD_SHAPE, "filterbank has %d bins, expected 1 + WHISPER_N_FFT/2 = 201", n_fft);
    CK(cudaSetDevice(m->device));
    if (m->mel) { mel_plan_destroy(m->mel); m->mel = nullptr; }
    CK(mel_plan_create(&m->mel, filters, n_mel, n_fft, m->stream));
    return Q2W_OK;
}

int q2w_model_upload_tensor(q2w_model* m, const char* name, int ggml_type, int n_dims, const int32_t* ne, const void* data,
                            size_t nbytes) {
    if (!m || !name || !ne || !data) return fail(Q2W_E_INVALID, "null argument");
    auto it = m->by_name.find(name);
    if (it == m->by_name.end()) return fail(Q2W_E_UNKNOWN_TENSOR, "unknown tensor '%s' in model file", name);
    Tensor& t = *it->second;
    int64_t got[3] = {1, 1, 1};
    int64_t nel = 1;
    if (n_dims < 1 || n_dims > 3) return fail(Q2W_E_BAD_SHAPE, "tensor '%s': n_dims %d", name, n_dims);
    for (int i = 0; i < n_dims; ++i) { got[i] = ne[i]; nel *= ne[i]; }
    if (nel != t.nelements()) return fail(Q2W_E_BAD_SHAPE, "tensor '%s' has wrong size in model file", name);
    if (got[0] != t.ne[0] || got[1] != t.ne[1] || got[2] != t.ne[2])
        return fail(Q2W_E_BAD_SHAPE, "tensor '%s' has wrong shape in model file: got [%lld, %lld, %lld], expected [%lld, %lld, %lld]", name,
                    (long long) got[0], (long long) got[1], (long long) got[2], (long long) t.ne[0], (long long) t.ne[1], (long long) t.ne[2]);
    if (ggml_type != t.file_type || nbytes != t.file_bytes())
        return fail(Q2W_E_BAD_SIZE, "tensor '%s' has wrong size in model file: got type %d / %zu bytes, expected type %d / %zu bytes", name,
                    ggml_type, nbytes, t.file_type, t.file_bytes());
    CK(cudaSetDevice(m->device));
    if (t.file_type == t.dev_type) {
        CK(cudaMemcpyAsync(t.d, data, nbytes, cudaMemcpyHostToDevice, m->stream));
        CK(cudaStreamSynchronize(m->stream));
    } else {
        // F32 matrix / conv kernel in an F32 model file: round to F16 once (the tensor-core operand type)
        void* tmp = nullptr;
        CK(cudaMalloc(&tmp, nbytes));
        cudaError_t e = cudaMemcpyAsync(tmp, data, nbytes, cudaMemcpyHostToDevice, m->stream);
        if (e == cudaSuccess) e = dequant_to_f16(tmp, Q2W_TYPE_F32, static_cast<__half*>(t.d), static_cast<size_t>(t.nrows()), static_cast<int>(t.ne[0]) , m->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
        cudaFree(tmp);
        if (e != cudaSuccess) return fail(Q2W_E_CUDA, "upload of '%s' failed: %s", name, cudaGetErrorString(e));
    }
    if (!t.loaded) { t.loaded = true; m->n_loaded++; }
    return Q2W_OK;
}

// bytes the model file must carry for tensor `name` (0: no such tensor) -- lets a loader reject a corrupt record before it reads it
size_t q2w_model_tensor_bytes(const q2w_model* m, const char* name) {
    if (!m || !name) return 0;
    auto it = m->by_name.find(name);
    return it == m->by_name.end() ? 0 : it->second->file_bytes();
}

int q2w_model_finalize(q2w_model* m) {
    if (!m) return fail(Q2W_E_INVALID, "null argument");
    if (m->n_loaded != static_cast<int>(m->by_name.size()))
        return fail(Q2W_E_INCOMPLETE, "not all tensors loaded from model file - expected %zu, got %d", m->by_name.size(), m->n_loaded);
    if (!m->mel) return fail(Q2W_E_INCOMPLETE, "mel filterbank not uploaded");
    CK(cudaSetDevice(m->device));
    CK(cudaStreamSynchronize(m->stream));
    m->finalized = true;
    return Q2W_OK;
}

void q2w_model_free(q2w_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    for (auto& kv : m->by_name) free_tensor(*kv.second);
    for (auto& ly : m->layers) {
        // every member, registered by name or not (creation may have failed midway through this layer); free_tensor is idempotent
        Tensor* ts[] = {&ly.ln1_w, &ly.ln1_b, &ly.q_w, &ly.q_b, &ly.k_w, &ly.v_w, &ly.v_b, &ly.o_w, &ly.o_b, &ly.ln2_w, &ly.ln2_b,
                        &ly.fc1_w, &ly.fc1_b, &ly.fc2_w, &ly.fc2_b};
        for (Tensor* t : ts) free_tensor(*t);
        if (ly.qkv_w) cudaFree(ly.qkv_w);
        if (ly.qkv_b) cudaFree(ly.qkv_b);
    }
    // tensors not yet registered by name (creation failed midway)
    free_tensor(m->pe); free_tensor(m->conv1_w); free_tensor(m->conv1_b); free_tensor(m->conv2_w); free_tensor(m->conv2_b);
    free_tensor(m->ln_w); free_tensor(m->ln_b);
    if (m->mel) mel_plan_destroy(m->mel);
    if (m->proj_w) cudaFree(m->proj_w);
    if (m->proj_b) cudaFree(m->proj_b);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

int q2w_model_device(const q2w_model* m) { return m ? m->device : -1; }

// The step after the path: Qwen2-Audio's multi_modal_projector = Linear(n_audio_state -> n_out) + bias on every
// embedding row (HF Qwen2AudioMultiModalProjector.linear). The reference stops at the final LayerNorm (src/qwen2-whisper.cpp:2175-2185);
// this is additive and optional: a model without it behaves exactly as before.
int q2w_model_upload_projector(q2w_model* m, int ggml_type, int n_out, const void* w_host, size_t nbytes, const float* bias_host) {
    if (!m || !w_host) return fail(Q2W_E_INVALID, "null argument");
    const int D = m->hp.n_audio_state;
    if (ggml_type != Q2W_TYPE_F32 && ggml_type != Q2W_TYPE_F16) return fail(Q2W_E_UNSUPPORTED, "projector weights must be F32 or F16");
    if (n_out <= 0 || n_out % 8) return fail(Q2W_E_BAD_SHAPE, "projector width %d must be a positive multiple of 8", n_out);
    if (nbytes != type_row_bytes(ggml_type, D) * static_cast<size_t>(n_out)) return fail(Q2W_E_BAD_SIZE, "projector weight has %zu bytes, expected [%d][%d]", nbytes, n_out, D);
    CK(cudaSetDevice(m->device));
    if (m->proj_w) { cudaFree(m->proj_w); m->proj_w = nullptr; }
    if (m->proj_b) { cudaFree(m->proj_b); m->proj_b = nullptr; }
    m->proj_out = 0;
    CK(cudaMalloc(reinterpret_cast<void**>(&m->proj_w), static_cast<size_t>(n_out) * D * sizeof(__half)));
    CK(cudaMalloc(reinterpret_cast<void**>(&m->proj_b), static_cast<size_t>(n_out) * sizeof(float)));
    cudaError_t e;
    if (ggml_type == Q2W_TYPE_F16) {
        e = cudaMemcpyAsync(m->proj_w, w_host, nbytes, cudaMemcpyHostToDevice, m->stream);
    } else {
        void* tmp = nullptr;
        CK(cudaMalloc(&tmp, nbytes));
        e = cudaMemcpyAsync(tmp, w_host, nbytes, cudaMemcpyHostToDevice, m->stream);
        if (e == cudaSuccess) e = dequant_to_f16(tmp, Q2W_TYPE_F32, m->proj_w, static_cast<size_t>(n_out), D, m->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
        cudaFree(tmp);
    }
    if (e == cudaSuccess) e = bias_host ? cudaMemcpyAsync(m->proj_b, bias_host, sizeof(float) * n_out, cudaMemcpyHostToDevice, m->stream)
                                        : cudaMemsetAsync(m->proj_b, 0, sizeof(float) * n_out, m->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) return fail(Q2W_E_CUDA, "projector upload failed: %s", cudaGetErrorString(e));
    m->proj_out = n_out;
    return Q2W_OK;
}
int q2w_model_projector_width(const q2w_model* m) { return m ? m->proj_out : 0; }
int q2w_model_n_tensors_expected(const q2w_model* m) { return m ? static_cast<int>(m->by_name.size()) : 0; }
int q2w_model_n_tensors_loaded(const q2w_model* m) { return m ? m->n_loaded : 0; }
size_t q2w_model_weight_bytes(const q2w_model* m) { return m ? m->weight_bytes : 0; }

// =============================================================================================== state
// workspaces that scale with max_batch (everything else in the state -- API mel, embeddings, timers, streams -- survives a resize)
static void free_workspaces(q2w_state* s) {
    void** ptrs[] = {reinterpret_cast<void**>(&s->x), reinterpret_cast<void**>(&s->ln), reinterpret_cast<void**>(&s->qkv),
                     reinterpret_cast<void**>(&s->att), reinterpret_cast<void**>(&s->h), reinterpret_cast<void**>(&s->pcm_dev),
                     reinterpret_cast<void**>(&s->nsamp_dev), reinterpret_cast<void**>(&s->logmel), reinterpret_cast<void**>(&s->winmax)};
    for (void** p : ptrs) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    if (s->g1) { cudaGraphExecDestroy(s->g1); s->g1 = nullptr; }
    s->g1_state = 0;
    s->dbg_windows = 0;
}

static cudaError_t alloc_workspaces(q2w_state* s, int max_batch) {
    const size_t B = max_batch, T = s->T, D = s->D;
    const size_t att_elems = std::max(T * D, static_cast<size_t>(s->T2) * 3 * s->n_mel);
    cudaError_t e = cudaSuccess;
#define SALLOC(ptr, bytes) if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&(ptr)), (bytes))
    SALLOC(s->x, B * T * D * sizeof(float));
    SALLOC(s->ln, B * T * D * sizeof(__half));
    SALLOC(s->qkv, B * T * 3 * D * sizeof(__half));
    SALLOC(s->att, B * att_elems * sizeof(__half));
    SALLOC(s->h, B * T * 4 * D * sizeof(__half));
    SALLOC(s->pcm_dev, 2 * B * s->win_samples * sizeof(float));
    SALLOC(s->nsamp_dev, 2 * B * sizeof(int));
    SALLOC(s->logmel, B * s->n_mel * s->ld_mel * sizeof(float));
    SALLOC(s->winmax, B * sizeof(float));
#undef SALLOC
    if (e == cudaSuccess) s->max_batch = max_batch;
    return e;
}

int q2w_state_create(q2w_state** out, q2w_model* m, int max_batch) {
    if (!out || !m) return fail(Q2W_E_INVALID, "null argument");
    *out = nullptr;
    if (!m->finalized) return fail(Q2W_E_INCOMPLETE, "model not finalized");
    if (max_batch < 1) return fail(Q2W_E_INVALID, "max_batch must be >= 1");
    CK(cudaSetDevice(m->device));
    q2w_state* s = new q2w_state();
    s->m = m;
    s->max_batch = max_batch;
    s->T = m->hp.n_audio_ctx; s->T2 = 2 * s->T; s->D = m->hp.n_audio_state; s->H = m->hp.n_audio_head; s->FF = 4 * s->D;
    s->n_mel = m->hp.n_mels;
    s->win_samples = s->T2 * 160;
    s->n_frames_batch = s->T2 + 2;                      // frames T2, T2+1 still see real samples and enter the max (:2522, :2634)
    s->ld_mel = (s->n_frames_batch + 15) / 16 * 16;
    const size_t D = s->D;
    const size_t wlayer_elems = static_cast<size_t>(12) * D * D;     // QKV 3 D^2 + out D^2 + fc1 4 D^2 + fc2 4 D^2
    {   // decode strategy for quantised weights, resolved once per state
        const char* e = getenv("Q2W_FUSED_DEQUANT");
        s->fused_dequant = e ? atoi(e) : 0;              // 1: decode inside the GEMM; 0 (default): decode warps riding in the attention kernel
        const char* sp = getenv("Q2W_E2E_SPLIT");
        s->e2e_split = sp ? std::max(1, atoi(sp)) : 2;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = alloc_workspaces(s, max_batch);
    if (m->wtype_dev != Q2W_TYPE_F16 && e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->wlayer), wlayer_elems * sizeof(__half));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->s_out, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_ticket[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->api_max), sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->att_sched), 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(s->att_sched, 0, 2 * sizeof(int));
    if (e != cudaSuccess) {
        q2w_state_free(s);
        return fail(e == cudaErrorMemoryAllocation ? Q2W_E_NOMEM : Q2W_E_CUDA, "state allocation failed: %s", cudaGetErrorString(e));
    }
    *out = s;
    return Q2W_OK;
}

// Resize the per-batch workspaces in place. The state's mel (whisper_pcm_to_mel / whisper_set_mel), its embeddings, timers and
// streams are kept; only scratch is reallocated (and the captured single-window graph dropped, its buffers are gone).
int q2w_state_set_max_batch(q2w_state* s, int max_batch) {
    if (!s) return fail(Q2W_E_INVALID, "null argument");
    if (max_batch < 1) return fail(Q2W_E_INVALID, "max_batch must be >= 1");
    if (max_batch == s->max_batch) return Q2W_OK;
    CK(cudaSetDevice(s->m->device));
    int rc = drain_tickets(s);
    if (rc) return rc;
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaStreamSynchronize(s->s_in));
    CK(cudaStreamSynchronize(s->s_out));
    const int old = s->max_batch;
    free_workspaces(s);
    cudaError_t e = alloc_workspaces(s, max_batch);
    if (e != cudaSuccess) {
        cudaGetLastError();
        free_workspaces(s);
        if (alloc_workspaces(s, old) != cudaSuccess) { cudaGetLastError(); free_workspaces(s); s->max_batch = 0; }
        return fail(e == cudaErrorMemoryAllocation ? Q2W_E_NOMEM : Q2W_E_CUDA, "workspace allocation for max_batch %d failed: %s", max_batch,
                    cudaGetErrorString(e));
    }
    s->mb_seq = 0;
    return Q2W_OK;
}

int q2w_state_max_batch(const q2w_state* s) { return s ? s->max_batch : 0; }

void q2w_state_free(q2w_state* s) {
    if (!s) return;
    cudaSetDevice(s->m->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->s_in) cudaStreamSynchronize(s->s_in);      // asynchronous batches may still be copying from / into caller memory
    if (s->s_out) cudaStreamSynchronize(s->s_out);
    free_workspaces(s);
    void* ptrs[] = {s->wlayer, s->emb, s->emb16, s->proj, s->api_mel, s->api_pcm, s->api_max, s->att_sched};
    for (void* p : ptrs) if (p) cudaFree(p);

    for (auto& r : s->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (int i = 0; i < 2; ++i) {
        if (s->ev_in[i]) cudaEventDestroy(s->ev_in[i]);
        if (s->ev_done[i]) cudaEventDestroy(s->ev_done[i]);
        if (s->ev_ticket[i]) cudaEventDestroy(s->ev_ticket[i]);
    }
    if (s->s_in) cudaStreamDestroy(s->s_in);
    if (s->s_out) cudaStreamDestroy(s->s_out);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int q2w_pcm_to_mel(q2w_state* s, const float* pcm_host, int n_samples) {
    if (!s || !pcm_host || n_samples <= 0) return fail(Q2W_E_INVALID, "bad argument");
    CK(cudaSetDevice(s->m->device));
    const int64_t t0 = now_us();
    // n_len = (n + 30 s + 2*200 - 400) / 160 ; n_len_org = 1 + (n + 200 - 400) / 160   (:2594-2613)
    const int n_len = static_cast<int>((static_cast<int64_t>(n_samples) + 480000) / 160);
    const int n_len_org = 1 + (n_samples + 200 - 400) / 160;
    const int ld = (n_len + 3) / 4 * 4;
    const size_t need = static_cast<size_t>(s->n_mel) * ld;
    if (need > s->api_mel_cap) {
        if (s->api_mel) cudaFree(s->api_mel);
        s->api_mel = nullptr; s->api_mel_cap = 0;
        CK(cudaMalloc(&s->api_mel, need * sizeof(float)));
        s->api_mel_cap = need;
    }
    if (static_cast<size_t>(n_samples) > s->api_pcm_cap) {
        if (s->api_pcm) cudaFree(s->api_pcm);
        s->api_pcm = nullptr; s->api_pcm_cap = 0;
        CK(cudaMalloc(&s->api_pcm, static_cast<size_t>(n_samples) * sizeof(float)));
        s->api_pcm_cap = n_samples;
    }
    CK(cudaMemcpyAsync(s->api_pcm, pcm_host, static_cast<size_t>(n_samples) * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    CKL(mel_logpower(s->m->mel, s->api_pcm, 0, nullptr, n_samples, 1, n_len, s->api_mel, ld, s->api_max, s->stream));
    g_launches.fetch_add(1);
    CKL(mel_normalize(s->api_mel, ld, n_len, s->n_mel, s->api_max, 1, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->api_n_len = n_len; s->api_n_len_org = n_len_org; s->api_ld = ld;
    s->t_mel_us += now_us() - t0;
    return Q2W_OK;
}

int q2w_set_mel(q2w_state* s, const float* mel_host, int n_len, int n_mel) {
    if (!s || !mel_host || n_len <= 0) return fail(Q2W_E_INVALID, "bad argument");
    if (n_mel != s->n_mel) return fail(Q2W_E_INVALID, "invalid number of mel bands: %d (expected %d)", n_mel, s->n_mel);
    CK(cudaSetDevice(s->m->device));
    const size_t need = static_cast<size_t>(n_mel) * n_len;
    if (need > s->api_mel_cap) {
        if (s->api_mel) cudaFree(s->api_mel);
        s->api_mel = nullptr; s->api_mel_cap = 0;
        CK(cudaMalloc(&s->api_mel, need * sizeof(float)));
        s->api_mel_cap = need;
    }
    CK(cudaMemcpyAsync(s->api_mel, mel_host, need * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->api_n_len = n_len; s->api_n_len_org = n_len; s->api_ld = n_len;
    return Q2W_OK;
}

int q2w_mel_n_len(const q2w_state* s) { return s ? s->api_n_len : 0; }
int q2w_mel_n_len_org(const q2w_state* s) { return s ? s->api_n_len_org : 0; }

int q2w_get_mel(q2w_state* s, float* out_host, size_t n_floats) {
    if (!s || !out_host) return fail(Q2W_E_INVALID, "null argument");
    if (!s->api_mel || s->api_n_len <= 0) return fail(Q2W_E_INVALID, "no mel in this state");
    if (n_floats < static_cast<size_t>(s->n_mel) * s->api_n_len) return fail(Q2W_E_INVALID, "output buffer too small");
    CK(cudaSetDevice(s->m->device));
    CK(cudaMemcpy2DAsync(out_host, static_cast<size_t>(s->api_n_len) * sizeof(float), s->api_mel, static_cast<size_t>(s->api_ld) * sizeof(float),
                         static_cast<size_t>(s->api_n_len) * sizeof(float), s->n_mel, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return Q2W_OK;
}

int q2w_encode(q2w_state* s, int mel_offset) {
    if (!s) return fail(Q2W_E_INVALID, "null argument");
    if (!s->api_mel || s->api_n_len <= 0) return fail(Q2W_E_INVALID, "no mel in this state: call q2w_pcm_to_mel or q2w_set_mel first");
    if (mel_offset < 0) return fail(Q2W_E_INVALID, "negative mel offset");
    CK(cudaSetDevice(s->m->device));
    const int64_t t0 = now_us();
    int rc = drain_tickets(s);
    if (rc) return rc;
    if ((rc = ensure_emb(s, 1))) return rc;
    s->emb_off_windows = 0;
    // window [offset, offset + 2*n_ctx) of the (already normalised) mel, zero past n_len   (:2264-2285; i0 = min(mel_offset, n_len))
    mel_offset = std::min(mel_offset, s->api_n_len);
    CKL(mel_to_conv1_operand(s->api_mel, s->api_ld, s->api_n_len, s->n_mel, nullptr, 0, mel_offset, s->T2, 1, s->att, s->stream));
    if ((rc = forward_from_a1(s, 1, 0))) return rc;
    CK(cudaStreamSynchronize(s->stream));
    s->emb_windows = 1;
    s->t_encode_us += now_us() - t0;
    s->n_encode++;
    return Q2W_OK;
}

// Whole-file streaming semantics: the state's mel was computed ONCE over the full PCM (global max, like
// whisper_pcm_to_mel), and n windows starting at the given frame offsets are encoded as one batch -- the batched form of calling
// whisper_full(ctx, {offset_ms}, NULL, 0) n times (src/qwen2-whisper.cpp:2349-2369: the mel is only recomputed when n_samples > 0).
int q2w_encode_offsets(q2w_state* s, const int32_t* mel_offsets, int n, float* out_host) {
    if (!s || !mel_offsets || n <= 0) return fail(Q2W_E_INVALID, "bad argument");
    if (!s->api_mel || s->api_n_len <= 0) return fail(Q2W_E_INVALID, "no mel in this state: call q2w_pcm_to_mel or q2w_set_mel first");
    for (int i = 0; i < n; ++i)
        if (mel_offsets[i] < 0) return fail(Q2W_E_INVALID, "negative mel offset");
    CK(cudaSetDevice(s->m->device));
    const int64_t t0 = now_us();
    int rc = drain_tickets(s);
    if (rc) return rc;
    if ((rc = ensure_emb(s, n))) return rc;
    s->emb_off_windows = 0;
    const size_t a1_per_window = static_cast<size_t>(s->T2) * 3 * s->n_mel;
    const size_t out_per_window = static_cast<size_t>(s->T / 2) * s->D;
    for (int w0 = 0; w0 < n; w0 += s->max_batch) {
        const int Bm = std::min(s->max_batch, n - w0);
        for (int b = 0; b < Bm; ++b)   // window slice + zero fill past n_len + im2col, one launch per window (same mel, different offset)
            CKL(mel_to_conv1_operand(s->api_mel, s->api_ld, s->api_n_len, s->n_mel, nullptr, 0, std::min(mel_offsets[w0 + b], s->api_n_len), s->T2, 1,
                                     s->att + static_cast<size_t>(b) * a1_per_window, s->stream));
        if ((rc = forward_from_a1(s, Bm, w0))) return rc;
        if (out_host)
            CK(cudaMemcpyAsync(out_host + static_cast<size_t>(w0) * out_per_window, s->emb + static_cast<size_t>(w0) * out_per_window,
                               static_cast<size_t>(Bm) * out_per_window * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    }
    CK(cudaStreamSynchronize(s->stream));
    s->emb_windows = n;
    s->t_encode_us += now_us() - t0;
    s->n_encode += n;
    return Q2W_OK;
}

// async_ticket == nullptr: synchronous (returns when the embeddings are on the host / device).
// async_ticket != nullptr: returns as soon as the work is queued; *async_ticket identifies the batch for q2w_encode_batch_wait.
static int encode_batch_impl(q2w_state* s, const float* pcm, bool pcm_on_host, size_t stride, const int32_t* n_samples, int B,
                             float* out_host, int* async_ticket) {
    if (!s || !pcm || B <= 0) return fail(Q2W_E_INVALID, "bad argument");
    if (stride == 0) return fail(Q2W_E_INVALID, "stride must be > 0");
    if (async_ticket && !out_host) return fail(Q2W_E_INVALID, "an asynchronous batch needs a host output buffer");
    CK(cudaSetDevice(s->m->device));
    const int64_t t0 = now_us();
    int rc;
    int region = 0;
    if (async_ticket) {
        region = s->next_ticket & 1;
        if (s->ticket_live[region] && (rc = ticket_wait(s, s->next_ticket - 2))) return rc;   // at most two batches in flight
        if (2 * static_cast<size_t>(B) > s->emb_cap_windows || (s->m->proj_w && !s->emb16)) {   // (re)allocating: nothing may be in flight
            if ((rc = drain_tickets(s))) return rc;
            if ((rc = ensure_emb(s, 2 * B))) return rc;
        }
    } else {
        if ((rc = drain_tickets(s))) return rc;
        if ((rc = ensure_emb(s, B))) return rc;
    }
    const size_t emb_base = static_cast<size_t>(region) * (s->emb_cap_windows / 2);   // in windows
    const size_t width = std::min(stride, static_cast<size_t>(s->win_samples));
    std::vector<int> ns(B);
    for (int b = 0; b < B; ++b) {
        int n = n_samples ? n_samples[b] : static_cast<int>(width);
        if (n < 0 || static_cast<size_t>(n) > width) return fail(Q2W_E_INVALID, "window %d: n_samples %d exceeds the window (%zu samples)", b, n, width);
        ns[b] = n;
    }
    const size_t out_per_window = static_cast<size_t>(s->T / 2) * s->D;
    // micro-batches: at most max_batch windows each.  With host buffers a batch that would fit in one micro-batch is still cut
    // in two, so the H2D copy of the second half and the D2H copy of the first half's embeddings overlap compute.
    int mb = std::min(s->max_batch, B);
    if (pcm_on_host && !async_ticket && B <= s->max_batch && B >= 16) {   // (queued batches overlap with each other instead: no cut)
        mb = (B + s->e2e_split - 1) / s->e2e_split;
    }
    for (int w0 = 0; w0 < B; w0 += mb, ++s->mb_seq) {
        const int Bm = std::min(mb, B - w0);
        const int slot = static_cast<int>(s->mb_seq & 1);
        int* nsamp = s->nsamp_dev + static_cast<size_t>(slot) * s->max_batch;
        // windows of equal length (the common case: full windows) need no length array on the device -- and a single-window call
        // then has no host-to-device copy at all in front of its first kernel
        bool uniform = true;
        for (int b = 1; b < Bm; ++b) uniform = uniform && ns[w0 + b] == ns[w0];
        const int n_max = uniform ? ns[w0] : s->win_samples;
        const float* pcm_dev = nullptr;
        size_t dev_stride = stride;
        if (pcm_on_host) {
            float* stage = s->pcm_dev + static_cast<size_t>(slot) * s->max_batch * s->win_samples;
            if (s->mb_seq >= 2) CK(cudaStreamWaitEvent(s->s_in, s->ev_done[slot], 0));   // compute of micro-batch mb_seq - 2 has consumed this slot
            if (!uniform) CK(cudaMemcpyAsync(nsamp, ns.data() + w0, sizeof(int) * Bm, cudaMemcpyHostToDevice, s->s_in));
            CK(cudaMemcpy2DAsync(stage, static_cast<size_t>(s->win_samples) * sizeof(float), pcm + static_cast<size_t>(w0) * stride,
                                 stride * sizeof(float), width * sizeof(float), Bm, cudaMemcpyHostToDevice, s->s_in));
            CK(cudaEventRecord(s->ev_in[slot], s->s_in));
            CK(cudaStreamWaitEvent(s->stream, s->ev_in[slot], 0));
            pcm_dev = stage;
            dev_stride = s->win_samples;
        } else {
            if (s->mb_seq >= 2) CK(cudaStreamWaitEvent(s->stream, s->ev_done[slot], 0));
            if (!uniform) CK(cudaMemcpyAsync(nsamp, ns.data() + w0, sizeof(int) * Bm, cudaMemcpyHostToDevice, s->stream));
            pcm_dev = pcm + static_cast<size_t>(w0) * stride;
        }
        if ((rc = batch_chunk(s, pcm_dev, dev_stride, uniform ? nullptr : nsamp, n_max, Bm, static_cast<int>(emb_base) + w0))) return rc;
        CK(cudaEventRecord(s->ev_done[slot], s->stream));
        if (out_host) {
            CK(cudaStreamWaitEvent(s->s_out, s->ev_done[slot], 0));
            CK(cudaMemcpyAsync(out_host + static_cast<size_t>(w0) * out_per_window, s->emb + (emb_base + w0) * out_per_window,
                               static_cast<size_t>(Bm) * out_per_window * sizeof(float), cudaMemcpyDeviceToHost, s->s_out));
        }
    }
    if (async_ticket) {
        CK(cudaEventRecord(s->ev_ticket[region], s->s_out));   // after the last D2H, which itself waits for the last micro-batch
        s->ticket_live[region] = true;
        s->ticket_B[region] = B;
        s->ticket_t0[region] = t0;
        *async_ticket = s->next_ticket++;
        return Q2W_OK;
    }
    if (out_host) CK(cudaStreamSynchronize(s->s_out));
    CK(cudaStreamSynchronize(s->stream));
    s->emb_windows = B;
    s->emb_off_windows = 0;
    s->t_encode_us += now_us() - t0;
    s->n_encode += B;
    return Q2W_OK;
}

int q2w_encode_batch_host(q2w_state* s, const float* pcm_host, size_t stride, const int32_t* n_samples, int B, float* out_host) {
    return encode_batch_impl(s, pcm_host, true, stride, n_samples, B, out_host, nullptr);
}

int q2w_encode_batch_device(q2w_state* s, const float* pcm_dev, size_t stride, const int32_t* n_samples_host, int B) {
    return encode_batch_impl(s, pcm_dev, false, stride, n_samples_host, B, nullptr, nullptr);
}

int q2w_encode_batch_host_async(q2w_state* s, const float* pcm_host, size_t stride, const int32_t* n_samples, int B, float* out_host, int* ticket) {
    if (!ticket) return fail(Q2W_E_INVALID, "null ticket");
    return encode_batch_impl(s, pcm_host, true, stride, n_samples, B, out_host, ticket);
}

int q2w_encode_batch_wait(q2w_state* s, int ticket) {
    if (!s) return fail(Q2W_E_INVALID, "null argument");
    return ticket_wait(s, ticket);
}

int q2w_embd_dims(const q2w_state* s, int* n_windows, int* n_out, int* n_state) {
    if (!s) return fail(Q2W_E_INVALID, "null argument");
    if (n_windows) *n_windows = s->emb_windows;
    if (n_out) *n_out = s->T / 2;
    if (n_state) *n_state = s->D;
    return Q2W_OK;
}

int q2w_get_embeddings(q2w_state* s, float* out_host, size_t offset_floats, size_t n_floats) {
    if (!s || !out_host) return fail(Q2W_E_INVALID, "null argument");
    const size_t total = static_cast<size_t>(s->emb_windows) * (s->T / 2) * s->D;
    if (!s->emb || offset_floats + n_floats > total) return fail(Q2W_E_INVALID, "embedding range [%zu, %zu) outside the %zu floats available", offset_floats, offset_floats + n_floats, total);
    CK(cudaSetDevice(s->m->device));
    CK(cudaMemcpyAsync(out_host, s->emb + s->emb_off_windows * (s->T / 2) * s->D + offset_floats, n_floats * sizeof(float), cudaMemcpyDeviceToHost,
                       s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return Q2W_OK;
}

// projector on the embeddings of the last encode / encode_batch: one tcgen05 GEMM, A = the F16 rows the pool + LN tail wrote
int q2w_project(q2w_state* s, float* out_host, size_t n_floats) {
    if (!s) return fail(Q2W_E_INVALID, "null argument");
    q2w_model* m = s->m;
    if (!m->proj_w || m->proj_out <= 0) return fail(Q2W_E_INVALID, "the model has no projector (q2w_model_upload_projector)");
    if (!s->emb16 || s->emb_windows <= 0) return fail(Q2W_E_INVALID, "no embeddings to project: encode after the projector was uploaded");
    CK(cudaSetDevice(m->device));
    const int rows = s->emb_windows * (s->T / 2);
    const int N = m->proj_out, D = s->D;
    if (static_cast<size_t>(rows) * N > s->proj_cap_rows) {          // capacity in floats (the projector width may change between calls)
        if (s->proj) cudaFree(s->proj);
        s->proj = nullptr; s->proj_cap_rows = 0;
        CK(cudaMalloc(reinterpret_cast<void**>(&s->proj), static_cast<size_t>(rows) * N * sizeof(float)));
        s->proj_cap_rows = static_cast<size_t>(rows) * N;
    }
    const __half* A = s->emb16 + s->emb_off_windows * (s->T / 2) * D;
    int rc = weight_gemm(s, A, D, m->proj_w, Q2W_TYPE_F16, true, rows, N, D, m->proj_b, s->proj, N, EPI_BIAS_F32, nullptr, nullptr, 0, 0, 1.f);
    if (rc) return rc;
    s->proj_rows = rows;
    if (out_host) {
        if (n_floats < static_cast<size_t>(rows) * N) return fail(Q2W_E_INVALID, "output buffer too small for [%d][%d]", rows, N);
        CK(cudaMemcpyAsync(out_host, s->proj, static_cast<size_t>(rows) * N * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    }
    CK(cudaStreamSynchronize(s->stream));
    return Q2W_OK;
}
int q2w_projection_dims(const q2w_state* s, int* n_rows, int* n_out) {
    if (!s) return fail(Q2W_E_INVALID, "null argument");
    if (n_rows) *n_rows = s->proj_rows;
    if (n_out) *n_out = s->m->proj_out;
    return Q2W_OK;
}
const float* q2w_projected_device(const q2w_state* s) { return s ? s->proj : nullptr; }

const float* q2w_embeddings_device(const q2w_state* s) { return s && s->emb ? s->emb + s->emb_off_windows * (s->T / 2) * s->D : nullptr; }

int q2w_get_batch_mel(q2w_state* s, int window, float* out_host) {
    if (!s || !out_host) return fail(Q2W_E_INVALID, "null argument");
    if (window < 0 || window >= s->max_batch) return fail(Q2W_E_INVALID, "window index out of range");
    CK(cudaSetDevice(s->m->device));
    // normalise a copy of the window's log-mel exactly as the conv1 operand builder does
    float* tmp = nullptr;
    const size_t n = static_cast<size_t>(s->n_mel) * s->ld_mel;
    CK(cudaMalloc(&tmp, n * sizeof(float)));
    cudaError_t e = cudaMemcpyAsync(tmp, s->logmel + static_cast<size_t>(window) * n, n * sizeof(float), cudaMemcpyDeviceToDevice, s->stream);
    if (e == cudaSuccess) e = mel_normalize(tmp, s->ld_mel, s->T2, s->n_mel, s->winmax + window, 1, s->stream);
    if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(out_host, static_cast<size_t>(s->T2) * sizeof(float), tmp, static_cast<size_t>(s->ld_mel) * sizeof(float),
                              static_cast<size_t>(s->T2) * sizeof(float), s->n_mel, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(Q2W_E_CUDA, "get_batch_mel: %s", cudaGetErrorString(e));
    return Q2W_OK;
}

void q2w_get_timings(const q2w_state* s, int64_t* t_mel_us, int64_t* t_encode_us, int32_t* n_encode) {
    if (!s) return;
    if (t_mel_us) *t_mel_us = s->t_mel_us;
    if (t_encode_us) *t_encode_us = s->t_encode_us;
    if (n_encode) *n_encode = s->n_encode;
}

void q2w_reset_timings(q2w_state* s) {
    if (!s) return;
    s->t_mel_us = s->t_encode_us = 0;
    s->n_encode = 0;
}

int q2w_profile_enable(q2w_state* s, int on) {
    if (!s) return fail(Q2W_E_INVALID, "null argument");
    CK(cudaSetDevice(s->m->device));
    CK(cudaStreamSynchronize(s->stream));
    s->prof_on = on != 0;
    s->prof_used = 0;
    return Q2W_OK;
}

int q2w_profile_read(q2w_state* s, int cls, double* total_ms, long* count, double* total_flops, double* total_bytes) {
    if (!s || cls < 0 || cls >= PC_COUNT) return fail(Q2W_E_INVALID, "bad argument");
    CK(cudaSetDevice(s->m->device));
    CK(cudaStreamSynchronize(s->stream));
    double ms = 0, fl = 0, by = 0;
    long n = 0;
    for (size_t i = 0; i < s->prof_used; ++i) {
        const auto& r = s->prof[i];
        if (r.cls != cls) continue;
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, r.a, r.b));
        ms += t; fl += r.flops; by += r.bytes; n++;
    }
    if (total_ms) *total_ms = ms;
    if (count) *count = n;
    if (total_flops) *total_flops = fl;
    if (total_bytes) *total_bytes = by;
    return Q2W_OK;
}

void* q2w_state_stream(const q2w_state* s) { return s ? static_cast<void*>(s->stream) : nullptr; }

int q2w_sync(q2w_state* s) {
    if (!s) return fail(Q2W_E_INVALID, "null argument");
    CK(cudaSetDevice(s->m->device));
    CK(cudaStreamSynchronize(s->stream));
    return Q2W_OK;
}

// =============================================================================================== stage taps (parity tests)
int q2w_debug_forward_layers(q2w_state* s, int n_layers) {
    if (!s) return fail(Q2W_E_INVALID, "null argument");
    s->dbg_layers = n_layers < 0 ? -1 : n_layers;
    return Q2W_OK;
}

int q2w_debug_get_residual(q2w_state* s, int window, float* out_host) {
    if (!s || !out_host) return fail(Q2W_E_INVALID, "null argument");
    if (window < 0 || window >= s->dbg_windows) return fail(Q2W_E_INVALID, "window %d not resident (last forward held %d)", window, s->dbg_windows);
    CK(cudaSetDevice(s->m->device));
    const size_t n = static_cast<size_t>(s->T) * s->D;
    CK(cudaMemcpyAsync(out_host, s->x + static_cast<size_t>(window) * n, n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return Q2W_OK;
}

// =============================================================================================== diagnostics
const char* q2w_last_error(void) { return g_err; }
long q2w_kernel_launches(void) { return g_launches.load(); }

int q2w_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++;
    }
    return ok;
}

const char* q2w_build_info(void) { return "libq2w_b200: sm_100a, tcgen05/TMEM GEMM + TMA, built " __DATE__ " " __TIME__; }

// =============================================================================================== kernel-level shims
int q2w_op_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, void* out, int ldo,
                int epilogue, const float* resid, const float* pos, int pos_period, int scale_cols, float scale, void* stream) {
    GemmArgs g{};
    g.A = static_cast<const __half*>(A); g.lda = lda; g.W = static_cast<const __half*>(W); g.ldw = ldw;
    g.M = M; g.N = N; g.K = K; g.bias = bias; g.out = out; g.ldo = ldo; g.resid = resid; g.pos = pos; g.pos_period = pos_period;
    g.scale_cols = scale_cols; g.scale = scale;
    static const int w_static = [] { const char* e = getenv("Q2W_OP_W_STATIC"); return e ? atoi(e) : 0; }();   // tools/gemm_b1.py: W is a constant there
    g.w_static = w_static;
    CKL(gemm_f16_tcgen05(g, static_cast<GemmEpilogue>(epilogue), static_cast<cudaStream_t>(stream)));
    return Q2W_OK;
}

int q2w_op_gemm_q(const void* A, int lda, const void* W_raw, int wtype, int M, int N, int K, const float* bias, void* out, int ldo,
                  int epilogue, const float* resid, int scale_cols, float scale, void* stream) {
    GemmArgs g{};
    g.A = static_cast<const __half*>(A); g.lda = lda; g.W = static_cast<const __half*>(W_raw); g.ldw = K; g.wtype = wtype;
    g.M = M; g.N = N; g.K = K; g.bias = bias; g.out = out; g.ldo = ldo; g.resid = resid; g.scale_cols = scale_cols; g.scale = scale;
    CKL(gemm_f16_tcgen05(g, static_cast<GemmEpilogue>(epilogue), static_cast<cudaStream_t>(stream)));
    return Q2W_OK;
}

int q2w_op_layernorm(const float* x, const float* gamma, const float* beta, void* y, int M, int D, float eps, void* stream) {
    CKL(layernorm_f32_to_f16(x, gamma, beta, static_cast<__half*>(y), M, D, eps, static_cast<cudaStream_t>(stream)));
    return Q2W_OK;
}

int q2w_op_pool_layernorm(const float* x, const float* gamma, const float* beta, float* y, int B, int T, int D, float eps, void* stream) {
    CKL(pool2_layernorm_f32(x, gamma, beta, y, B, T, D, eps, static_cast<cudaStream_t>(stream)));
    return Q2W_OK;
}

int q2w_op_attention(const void* qkv, void* out, int B, int T, int H, void* stream) {
    // stand-alone op (tests, tools): one work counter per device, so callers on one device must not overlap launches of this op
    static std::mutex mu;
    static int* sched[64] = {};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(Q2W_E_INVALID, "device index out of range");
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!sched[dev]) {
            CK(cudaMalloc(reinterpret_cast<void**>(&sched[dev]), 2 * sizeof(int)));
            CK(cudaMemset(sched[dev], 0, 2 * sizeof(int)));
        }
    }
    CKL(attention_f16_tcgen05(static_cast<const __half*>(qkv), static_cast<__half*>(out), B, T, H, sched[dev], static_cast<cudaStream_t>(stream)));
    return Q2W_OK;
}

int q2w_op_dequant(const void* src, int ggml_type, void* dst, size_t rows, int K, void* stream) {
    CKL(dequant_to_f16(src, ggml_type, static_cast<__half*>(dst), rows, K, static_cast<cudaStream_t>(stream)));
    return Q2W_OK;
}

// up to four quantised matrices decoded to F16: by the stand-alone kernel (with_attention == 0) or by the idle warps of one attention launch
// over qkv / out (with_attention != 0) -- the way the engine decodes a block's weights
int q2w_op_dequant_multi(const void* const* src, void* const* dst_f16, const unsigned long long* nblocks, int n, int ggml_type, int with_attention,
                         const void* qkv, void* att_out, int B, int T, int H, void* stream) {
    if (!src || !dst_f16 || !nblocks || n < 1 || n > 4) return fail(Q2W_E_INVALID, "bad argument");
    DequantJob job{};
    for (int i = 0; i < n; ++i) {
        job.src[i] = static_cast<const uint8_t*>(src[i]);
        job.dst[i] = static_cast<__half*>(dst_f16[i]);
        job.nblocks[i] = nblocks[i];
    }
    if (!with_attention) {
        CKL(dequant_multi_to_f16(job, ggml_type, static_cast<cudaStream_t>(stream)));
        return Q2W_OK;
    }
    static std::mutex mu;
    static int* sched[64] = {};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(Q2W_E_INVALID, "device index out of range");
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!sched[dev]) {
            CK(cudaMalloc(reinterpret_cast<void**>(&sched[dev]), 2 * sizeof(int)));
            CK(cudaMemset(sched[dev], 0, 2 * sizeof(int)));
        }
    }
    CKL(attention_f16_tcgen05(static_cast<const __half*>(qkv), static_cast<__half*>(att_out), B, T, H, sched[dev], static_cast<cudaStream_t>(stream), &job,
                              ggml_type));
    return Q2W_OK;
}

int q2w_op_conv2_im2col(const void* h1, void* A2, int B, int T2, int C, void* stream) {
    CKL(conv2_im2col(static_cast<const __half*>(h1), static_cast<__half*>(A2), B, T2, C, static_cast<cudaStream_t>(stream)));
    return Q2W_OK;
}

int q2w_op_conv1_operand(const float* mel_dev, int ld_frames, int n_frames_valid, int n_mel, const void* win_max_keys_dev, int normalise,
                          int offset, int n_ctx2, int B, void* A1_f16, void* stream) {
    CKL(mel_to_conv1_operand(mel_dev, ld_frames, n_frames_valid, n_mel, static_cast<const float*>(win_max_keys_dev), normalise, offset, n_ctx2, B,
                             static_cast<__half*>(A1_f16), static_cast<cudaStream_t>(stream)));
    return Q2W_OK;
}

void q2w_op_set_gemm_splitk(int mode) { gemm_set_splitk_mode(mode); }

int q2w_op_mel(const float* filters_host, int n_mel, const float* pcm_dev, size_t stride, const int32_t* n_samples_dev, int n_max,
               int B, int n_frames, float* logmel_dev, int ld, void* win_max_dev, int normalise, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MelPlan* plan = nullptr;
    CK(mel_plan_create(&plan, filters_host, n_mel, 201, st));
    cudaError_t e = mel_logpower(plan, pcm_dev, stride, n_samples_dev, n_max, B, n_frames, logmel_dev, ld, static_cast<float*>(win_max_dev), st);
    g_launches.fetch_add(2);
    if (e == cudaSuccess && normalise) {
        e = mel_normalize(logmel_dev, ld, n_frames, n_mel, static_cast<const float*>(win_max_dev), B, st);
        g_launches.fetch_add(1);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    mel_plan_destroy(plan);
    if (e != cudaSuccess) return fail(Q2W_E_CUDA, "q2w_op_mel: %s", cudaGetErrorString(e));
    return Q2W_OK;
}

}  // extern "C"
