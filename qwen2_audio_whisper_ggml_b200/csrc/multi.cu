// multi.cu -- one-process, all-GPU data parallelism behind the C ABI (include/q2w_b200.h, q2w_multi_*).
//
// every 30 s window is an independent unit (own mel max, own attention), so a batch shards with no exchange
// step: window w -> device floor(w * G / B) (contiguous blocks), a full weight replica per device, one host worker thread +
// one q2w_state (stream, workspaces) per device, results stay on the producing device.  The reference has no counterpart
// (its whisper_context_params.gpu_device picks ONE device, include/qwen2-whisper.h:118); a C++ host such as
// examples/main/main.cpp:455-591 gets all GPUs of the box through this handle without torch / torchrun.
// No collective on the compute path.  Only on request the embeddings are gathered onto one device, ordered by window index:
// peer copies over NVLink (cudaMemcpyPeerAsync on the producing device's stream, issued by its own worker as soon as its shard
// is done) -- the single-process form of the NCCL gather that qwen2_audio_whisper_ggml_b200/parallel.py issues between processes.
#include "../../include/q2w_b200.h"
#include "ops.h"

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Job {
    const float* pcm = nullptr;
    size_t stride = 0;
    const int32_t* n_samples = nullptr;
    int B = 0;
    float* out_host = nullptr;
    int gather_slot = -1;      // index into q2w_multi::workers of the gather target, -1: none
    float* gather_buf = nullptr;
    int mode = 0;              // 0: synchronous batch; 1: queue the shard (asynchronous host batch); 2: wait for the shard queued under `slot`
    int slot = 0;              // asynchronous batches: which of the two in-flight batches
};

struct Worker {
    int index = 0;
    int device = 0;
    q2w_model* model = nullptr;
    q2w_state* state = nullptr;
    std::thread thread;
    // hand-off (guarded by the owner's mutex)
    bool has_job = false, done = false, quit = false;
    Job job;
    int lo = 0, hi = 0;
    int rc = Q2W_OK;
    std::string err;
    double ms = 0.0;           // device time of the last shard (CUDA events on the worker's stream)
    int shard_ticket[2] = {-1, -1};   // the state-level tickets of the two asynchronous batches that may be in flight (-1: empty shard)
};

}  // namespace

struct q2w_multi {
    std::vector<Worker> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    int max_batch = 1;
    int n_out = 0, n_state = 0;
    float* gather_buf = nullptr;       // on workers[gather_slot].device
    size_t gather_cap_windows = 0;
    int gather_slot = -1;
    int last_B = 0;
    bool peer_enabled = false;
    int next_ticket = 0;               // asynchronous batches: ticket t lives in slot t & 1
    bool slot_live[2] = {false, false};
};

namespace {

void shard_bounds(int B, int g, int G, int& lo, int& hi) {
    // window w belongs to device floor(w * G / B): start = ceil(g * B / G)   (same rule as parallel.shard_bounds)
    lo = static_cast<int>((static_cast<long long>(g) * B + G - 1) / G);
    hi = static_cast<int>((static_cast<long long>(g + 1) * B + G - 1) / G);
    if (hi > B) hi = B;
    if (lo > hi) lo = hi;
}

void run_job(q2w_multi* mm, Worker& w) {
    const Job& j = w.job;
    w.rc = Q2W_OK;
    w.err.clear();
    w.ms = 0.0;
    const size_t opw = static_cast<size_t>(mm->n_out) * mm->n_state;
    const int n = w.hi - w.lo;
    cudaSetDevice(w.device);
    if (j.mode == 1) {                                       // queue this shard and return: copies and kernels overlap across batches
        w.shard_ticket[j.slot] = -1;
        if (n <= 0) return;
        w.rc = q2w_encode_batch_host_async(w.state, j.pcm + static_cast<size_t>(w.lo) * j.stride, j.stride, j.n_samples ? j.n_samples + w.lo : nullptr, n,
                                           j.out_host + static_cast<size_t>(w.lo) * opw, &w.shard_ticket[j.slot]);
        if (w.rc != Q2W_OK) w.err = q2w_last_error();
        return;
    }
    if (j.mode == 2) {
        if (w.shard_ticket[j.slot] < 0) return;
        w.rc = q2w_encode_batch_wait(w.state, w.shard_ticket[j.slot]);
        if (w.rc != Q2W_OK) w.err = q2w_last_error();
        w.shard_ticket[j.slot] = -1;
        return;
    }
    if (n <= 0) return;                                      // empty shard (B < number of devices)
    cudaStream_t st = static_cast<cudaStream_t>(q2w_state_stream(w.state));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    w.rc = q2w_encode_batch_host(w.state, j.pcm + static_cast<size_t>(w.lo) * j.stride, j.stride, j.n_samples ? j.n_samples + w.lo : nullptr, n,
                                 j.out_host ? j.out_host + static_cast<size_t>(w.lo) * opw : nullptr);
    if (w.rc != Q2W_OK) {
        w.err = q2w_last_error();
    } else if (j.gather_buf) {
        // this shard's rows of the gathered [B][n_out][n_state] tensor, written straight into the target device's memory
        const float* src = q2w_embeddings_device(w.state);
        const int dst_dev = mm->workers[j.gather_slot].device;
        cudaError_t e = cudaMemcpyPeerAsync(j.gather_buf + static_cast<size_t>(w.lo) * opw, dst_dev, src, w.device, static_cast<size_t>(n) * opw * sizeof(float), st);
        if (e != cudaSuccess) {
            w.rc = Q2W_E_CUDA;
            w.err = std::string("gather peer copy failed: ") + cudaGetErrorString(e);
        }
    }
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess && w.rc == Q2W_OK) {
        w.rc = Q2W_E_CUDA;
        w.err = "device work of the shard failed";
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) w.ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

void worker_main(q2w_multi* mm, int idx) {
    Worker& w = mm->workers[idx];
    cudaSetDevice(w.device);
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(mm->mu);
            mm->cv_job.wait(lk, [&] { return w.has_job || w.quit; });
            if (w.quit) return;
        }
        run_job(mm, w);
        {
            std::lock_guard<std::mutex> lk(mm->mu);
            w.has_job = false;
            w.done = true;
        }
        mm->cv_done.notify_all();
    }
}

// hands every worker its job and waits until all have reported (the workers run concurrently, one per device)
int dispatch(q2w_multi* mm, const Job& proto, int B) {
    const int G = static_cast<int>(mm->workers.size());
    {
        std::lock_guard<std::mutex> lk(mm->mu);
        for (int i = 0; i < G; ++i) {
            Worker& w = mm->workers[i];
            w.job = proto;
            if (proto.mode != 2) shard_bounds(B, i, G, w.lo, w.hi);
            w.done = false;
            w.has_job = true;
        }
    }
    mm->cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lk(mm->mu);
        mm->cv_done.wait(lk, [&] {
            for (auto& w : mm->workers)
                if (!w.done) return false;
            return true;
        });
    }
    for (auto& w : mm->workers)
        if (w.rc != Q2W_OK) return q2w::set_last_error(w.rc, (std::string("device ") + std::to_string(w.device) + ": " + w.err).c_str());
    return Q2W_OK;
}

}  // namespace

extern "C" {

int q2w_multi_create(q2w_multi** out, q2w_model* const* models, int n, int max_batch) {
    if (!out || !models || n < 1 || max_batch < 1) return q2w::set_last_error(Q2W_E_INVALID, "q2w_multi_create: bad argument");
    *out = nullptr;
    q2w_multi* mm = new q2w_multi();
    mm->workers.resize(n);
    mm->max_batch = max_batch;
    for (int i = 0; i < n; ++i) {
        Worker& w = mm->workers[i];
        w.index = i;
        w.model = models[i];
        if (!w.model) { q2w_multi_free(mm); return q2w::set_last_error(Q2W_E_INVALID, "q2w_multi_create: null model"); }
        w.device = q2w_model_device(w.model);
        const int rc = q2w_state_create(&w.state, w.model, max_batch);
        if (rc != Q2W_OK) { q2w_multi_free(mm); return rc; }
    }
    q2w_embd_dims(mm->workers[0].state, nullptr, &mm->n_out, &mm->n_state);
    // direct peer access between every pair of distinct devices where the hardware offers it (NVLink / NVSwitch); without it
    // cudaMemcpyPeerAsync still works, staged by the driver
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const int a = mm->workers[i].device, b = mm->workers[j].device;
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) {
                cudaSetDevice(a);
                const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) mm->peer_enabled = true;
                cudaGetLastError();
            }
        }
    for (int i = 0; i < n; ++i) mm->workers[i].thread = std::thread(worker_main, mm, i);
    *out = mm;
    return Q2W_OK;
}

void q2w_multi_free(q2w_multi* mm) {
    if (!mm) return;
    {
        std::lock_guard<std::mutex> lk(mm->mu);
        for (auto& w : mm->workers) w.quit = true;
    }
    mm->cv_job.notify_all();
    for (auto& w : mm->workers)
        if (w.thread.joinable()) w.thread.join();
    for (auto& w : mm->workers)
        if (w.state) q2w_state_free(w.state);
    if (mm->gather_buf) {
        cudaSetDevice(mm->workers[mm->gather_slot].device);
        cudaFree(mm->gather_buf);
    }
    delete mm;
}

int q2w_multi_n_devices(const q2w_multi* mm) { return mm ? static_cast<int>(mm->workers.size()) : 0; }
int q2w_multi_device(const q2w_multi* mm, int i) { return (mm && i >= 0 && i < static_cast<int>(mm->workers.size())) ? mm->workers[i].device : -1; }
q2w_state* q2w_multi_state(q2w_multi* mm, int i) { return (mm && i >= 0 && i < static_cast<int>(mm->workers.size())) ? mm->workers[i].state : nullptr; }

int q2w_multi_shard_bounds(const q2w_multi* mm, int B, int i, int* lo, int* hi) {
    if (!mm || B < 0 || i < 0 || i >= static_cast<int>(mm->workers.size()) || !lo || !hi) return q2w::set_last_error(Q2W_E_INVALID, "q2w_multi_shard_bounds: bad argument");
    shard_bounds(B, i, static_cast<int>(mm->workers.size()), *lo, *hi);
    return Q2W_OK;
}

int q2w_multi_set_max_batch(q2w_multi* mm, int max_batch) {
    if (!mm || max_batch < 1) return q2w::set_last_error(Q2W_E_INVALID, "q2w_multi_set_max_batch: bad argument");
    for (auto& w : mm->workers) {
        const int rc = q2w_state_set_max_batch(w.state, max_batch);
        if (rc != Q2W_OK) return rc;
    }
    mm->max_batch = max_batch;
    return Q2W_OK;
}

int q2w_multi_encode_batch_host(q2w_multi* mm, const float* pcm_host, size_t stride, const int32_t* n_samples, int B, float* out_host, int gather_device) {
    if (!mm || !pcm_host || B <= 0 || stride == 0) return q2w::set_last_error(Q2W_E_INVALID, "q2w_multi_encode_batch_host: bad argument");
    const int G = static_cast<int>(mm->workers.size());
    int gslot = -1;
    if (gather_device >= 0) {
        for (int i = 0; i < G; ++i)
            if (mm->workers[i].device == gather_device) { gslot = i; break; }
        if (gslot < 0) return q2w::set_last_error(Q2W_E_INVALID, "gather device is not one of this handle's devices");
        if (gslot != mm->gather_slot || static_cast<size_t>(B) > mm->gather_cap_windows) {
            if (mm->gather_buf) {
                cudaSetDevice(mm->workers[mm->gather_slot].device);
                cudaFree(mm->gather_buf);
                mm->gather_buf = nullptr;
                mm->gather_cap_windows = 0;
            }
            cudaSetDevice(gather_device);
            const cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&mm->gather_buf), static_cast<size_t>(B) * mm->n_out * mm->n_state * sizeof(float));
            if (e != cudaSuccess) {
                cudaGetLastError();
                return q2w::set_last_error(Q2W_E_NOMEM, "gather buffer allocation failed");
            }
            mm->gather_slot = gslot;
            mm->gather_cap_windows = B;
        }
    }
    for (int t = mm->next_ticket - 2; t < mm->next_ticket; ++t)          // a synchronous call first retires whatever is still in flight
        if (t >= 0 && mm->slot_live[t & 1]) {
            const int rc = q2w_multi_encode_batch_wait(mm, t);
            if (rc != Q2W_OK) return rc;
        }
    Job j;
    j.pcm = pcm_host; j.stride = stride; j.n_samples = n_samples; j.B = B; j.out_host = out_host;
    j.gather_slot = gslot;
    j.gather_buf = gslot >= 0 ? mm->gather_buf : nullptr;
    const int rc = dispatch(mm, j, B);
    mm->last_B = B;
    return rc;
}

// Asynchronous form: every device queues its shard (H2D, kernels, D2H on its own streams) and the call returns; at most two batches are
// in flight (a third submit first waits for the oldest). pcm_host / out_host must stay valid until q2w_multi_encode_batch_wait(ticket).
int q2w_multi_encode_batch_host_async(q2w_multi* mm, const float* pcm_host, size_t stride, const int32_t* n_samples, int B, float* out_host, int* ticket) {
    if (!mm || !pcm_host || !out_host || !ticket || B <= 0 || stride == 0) return q2w::set_last_error(Q2W_E_INVALID, "q2w_multi_encode_batch_host_async: bad argument");
    const int slot = mm->next_ticket & 1;
    if (mm->slot_live[slot]) {
        const int rc = q2w_multi_encode_batch_wait(mm, mm->next_ticket - 2);
        if (rc != Q2W_OK) return rc;
    }
    Job j;
    j.pcm = pcm_host; j.stride = stride; j.n_samples = n_samples; j.B = B; j.out_host = out_host; j.mode = 1; j.slot = slot;
    const int rc = dispatch(mm, j, B);
    if (rc != Q2W_OK) return rc;
    mm->slot_live[slot] = true;
    mm->last_B = B;
    *ticket = mm->next_ticket++;
    return Q2W_OK;
}

int q2w_multi_encode_batch_wait(q2w_multi* mm, int ticket) {
    if (!mm) return q2w::set_last_error(Q2W_E_INVALID, "null argument");
    const int slot = ticket & 1;
    if (ticket < 0 || ticket >= mm->next_ticket || mm->next_ticket - ticket > 2 || !mm->slot_live[slot])
        return q2w::set_last_error(Q2W_E_INVALID, "ticket is not in flight");
    Job j;
    j.mode = 2; j.slot = slot;
    const int rc = dispatch(mm, j, 0);
    mm->slot_live[slot] = false;
    return rc;
}

const float* q2w_multi_gathered_device(const q2w_multi* mm) { return mm ? mm->gather_buf : nullptr; }

int q2w_multi_get_gathered(q2w_multi* mm, float* out_host, size_t n_floats) {
    if (!mm || !out_host || !mm->gather_buf) return q2w::set_last_error(Q2W_E_INVALID, "no gathered embeddings");
    const size_t total = static_cast<size_t>(mm->last_B) * mm->n_out * mm->n_state;
    if (n_floats > total) return q2w::set_last_error(Q2W_E_INVALID, "more floats requested than gathered");
    cudaSetDevice(mm->workers[mm->gather_slot].device);
    const cudaError_t e = cudaMemcpy(out_host, mm->gather_buf, n_floats * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return q2w::set_last_error(Q2W_E_CUDA, cudaGetErrorString(e));
    return Q2W_OK;
}

double q2w_multi_last_device_ms(const q2w_multi* mm, int i) {
    return (mm && i >= 0 && i < static_cast<int>(mm->workers.size())) ? mm->workers[i].ms : 0.0;
}

}  // extern "C"
