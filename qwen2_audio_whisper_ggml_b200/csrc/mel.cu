// mel.cu -- fused log-mel front-end for sm_100a.
//
// Replaces log_mel_spectrogram() + log_mel_spectrogram_worker_thread() + fft()/dft()
// (/root/reference/src/qwen2-whisper.cpp:2443-2665):
//   x      = [ reflect(pcm[1..200]) | pcm[0..n) | zeros ]                              (:2594-2606)
//   frame f: x[160 f .. 160 f + 400) * hann_periodic_400                               (:2526-2533)
//   X      = rfft_400(frame), P[k] = re^2 + im^2, k = 0..200                            (:2536-2542)
//   m[j,f] = log10(max(sum_k P[k] * filt[j,k], 1e-10))                                  (:2545-2561)
//   g = max m ; m = max(m, g - 8) ; m = (m + 4) / 4                                     (:2634-2649)
//
// One CTA = 16 consecutive frames of one window.  The 2800 PCM samples those frames touch are staged once in
// shared memory (coalesced), each frame's 400-point real FFT is computed in FP32 as a 200-point complex FFT
// (mixed radix 8 x 5 x 5, decimation in time, twiddles from an exact table) followed by the real-input split,
// the 128 x 201 filterbank is applied as a banded product (only each row's non-zero span, found at upload time,
// so any filter matrix from a model file is handled exactly), and log10 + the per-window running max
// (atomicMax on an order-preserving integer key) finish the pass.  The clamp/normalise step needs the global
// max, so it is fused into the consumer (conv1 operand builder) or run as mel_normalize() for the API mel.
#include "ops.h"

#include <cmath>
#include <vector>

namespace q2w {

struct MelPlan {
    float* filters = nullptr;   // [n_mel][n_bins]
    int2* ranges = nullptr;     // [n_mel] {first non-zero bin, one past last}
    float* hann = nullptr;      // [400]
    float2* tw = nullptr;       // [400] exp(-2 pi i k / 400)
    int n_mel = 0, n_bins = 0;
};

namespace {

constexpr int N_FFT = 400, HOP = 160, PAD = 200, NZ = 200, NBINS = 201;
constexpr int FPB = 16;                               // frames per block
constexpr int SEG = (FPB - 1) * HOP + N_FFT;          // 2800 samples staged per block
constexpr int ZPITCH = NZ + NZ / 8;                   // 225 float2 per frame (1 pad per 8 -> conflict-free radix-8 stores)
constexpr int THREADS = 256;
constexpr int P_BYTES = ((NBINS * FPB > SEG + N_FFT ? NBINS * FPB : SEG + N_FFT) * 4 + 15) / 16 * 16;   // max(s_p, s_x + s_hann)

__host__ __device__ inline int zidx(int p) { return p + (p >> 3); }

__device__ __forceinline__ int float_to_key(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float key_to_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

__device__ __forceinline__ void dft4(float2 b0, float2 b1, float2 b2, float2 b3, float2& y0, float2& y1, float2& y2,
                                     float2& y3) {
    const float2 s0 = cadd(b0, b2), s1 = csub(b0, b2), s2 = cadd(b1, b3), s3 = mul_mi(csub(b1, b3));
    y0 = cadd(s0, s2); y2 = csub(s0, s2); y1 = cadd(s1, s3); y3 = csub(s1, s3);
}

__device__ __forceinline__ void dft5(const float2 (&v)[5], float2 (&o)[5]) {
    constexpr float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    constexpr float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
    o[0] = make_float2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
    const float2 m1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
    const float2 m2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
    const float2 n1 = mul_mi(make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
    const float2 n2 = mul_mi(make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
    o[1] = cadd(m1, n1); o[4] = csub(m1, n1);
    o[2] = cadd(m2, n2); o[3] = csub(m2, n2);
}

__global__ void init_keys_kernel(int* keys, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = float_to_key(-INFINITY);
}

// grid = (ceil(n_frames / FPB), B)
__global__ void __launch_bounds__(THREADS)
mel_logpower_kernel(const float* __restrict__ filters, const int2* __restrict__ ranges, const float* __restrict__ hann_g,
                    const float2* __restrict__ tw_g, int n_mel, const float* __restrict__ pcm, size_t pcm_stride,
                    const int* __restrict__ n_samples_dev, int n_max, int n_frames, float* __restrict__ logmel,
                    int ld_frames, int* __restrict__ win_max_key) {
    extern __shared__ __align__(16) uint8_t smem_mel[];
    // s_x and s_hann are dead after pass 1, so the power spectrum s_p reuses their bytes: 44.9 KB per block -> 5 blocks per SM
    float* s_x = reinterpret_cast<float*>(smem_mel);                 // [SEG]
    float* s_hann = s_x + SEG;                                       // [400]
    float* s_p = reinterpret_cast<float*>(smem_mel);                 // [NBINS][FPB]   (bin-major, frame-minor), aliases s_x | s_hann | pad
    float2* s_tw = reinterpret_cast<float2*>(smem_mel + P_BYTES);    // [400]
    float2* s_z = s_tw + N_FFT;                                      // [FPB][ZPITCH]
    __shared__ float s_red[THREADS / 32];

    const int b = blockIdx.y;
    const int f0 = blockIdx.x * FPB;
    const int tid = threadIdx.x;
    const int n = n_samples_dev ? min(n_samples_dev[b], n_max) : n_max;
    const float* pw = pcm + static_cast<size_t>(b) * pcm_stride;
    float* out = logmel + static_cast<size_t>(b) * n_mel * ld_frames;

    const long x0 = static_cast<long>(f0) * HOP;  // index into the padded signal
    if (x0 >= static_cast<long>(n) + PAD) {
        // every frame of this block lies in the zero padding: log10(1e-10) exactly as :2566-2571
        for (int idx = tid; idx < n_mel * FPB; idx += THREADS) {
            const int j = idx / FPB, f = f0 + (idx - j * FPB);
            if (f < n_frames) out[static_cast<size_t>(j) * ld_frames + f] = -10.0f;
        }
        if (tid == 0) atomicMax(&win_max_key[b], float_to_key(-10.0f));
        return;
    }

    // ---- stage the padded signal segment, window table and twiddles.  All loads of a thread are issued before its stores
    //      (round-1 ncu: the one-load-one-store loop spent 31 % of the kernel's stall samples waiting on its own LDG)
    {
        constexpr int PER_THREAD = (SEG + THREADS - 1) / THREADS;   // 11
        float v[PER_THREAD];
#pragma unroll
        for (int it = 0; it < PER_THREAD; ++it) {
            const int i = tid + it * THREADS;
            const long xi = x0 + i;
            const long sidx = xi < PAD ? PAD - xi : xi - PAD;        // reflect: x[i] = pcm[200 - i]
            v[it] = (i < SEG && sidx < n) ? __ldg(pw + sidx) : 0.f;
        }
#pragma unroll
        for (int it = 0; it < PER_THREAD; ++it) {
            const int i = tid + it * THREADS;
            if (i < SEG) s_x[i] = v[it];
        }
    }
    for (int i = tid; i < N_FFT; i += THREADS) {
        s_hann[i] = hann_g[i];
        s_tw[i] = tw_g[i];
    }
    __syncthreads();

    // ---- pass 1: radix-8 on z[n] = x[2n] + i x[2n+1], n = w + 25 t  (w = j3 + 5 j2), out at 40 j3 + 8 j2 + k
    for (int item = tid; item < FPB * 25; item += THREADS) {
        const int f = item / 25, w = item - f * 25;
        const int j3 = w % 5, j2 = w / 5;
        const float* xf = s_x + f * HOP;
        float2 a[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int nn = w + 25 * t;
            const float2 xv = *reinterpret_cast<const float2*>(xf + 2 * nn);
            const float2 hv = *reinterpret_cast<const float2*>(s_hann + 2 * nn);
            a[t] = make_float2(xv.x * hv.x, xv.y * hv.y);
        }
        float2 e0, e1, e2, e3, o0, o1, o2, o3;
        dft4(a[0], a[2], a[4], a[6], e0, e1, e2, e3);
        dft4(a[1], a[3], a[5], a[7], o0, o1, o2, o3);
        constexpr float r = 0.70710678118654752440f;
        const float2 w1 = make_float2(r * (o1.x + o1.y), r * (o1.y - o1.x));      // o1 * (1 - i)/sqrt2
        const float2 w2 = mul_mi(o2);                                              // o2 * (-i)
        const float2 w3 = make_float2(r * (o3.y - o3.x), -r * (o3.x + o3.y));      // o3 * (-1 - i)/sqrt2
        float2* z = s_z + f * ZPITCH;
        const int base = 40 * j3 + 8 * j2;
        z[zidx(base + 0)] = cadd(e0, o0); z[zidx(base + 4)] = csub(e0, o0);
        z[zidx(base + 1)] = cadd(e1, w1); z[zidx(base + 5)] = csub(e1, w1);
        z[zidx(base + 2)] = cadd(e2, w2); z[zidx(base + 6)] = csub(e2, w2);
        z[zidx(base + 3)] = cadd(e3, w3); z[zidx(base + 7)] = csub(e3, w3);
    }
    __syncthreads();

    // ---- pass 2: radix-5, sub-FFT length 8 -> 40.  twiddle W_40^{jk} = W_400^{10 j k}
    for (int item = tid; item < FPB * 40; item += THREADS) {
        const int f = item / 40, w = item - f * 40;
        const int blk = w >> 3, k = w & 7;
        float2* z = s_z + f * ZPITCH;
        const int base = 40 * blk + k;
        float2 v[5], o[5];
        v[0] = z[zidx(base)];
#pragma unroll
        for (int j = 1; j < 5; ++j) v[j] = cmul(z[zidx(base + 8 * j)], s_tw[10 * j * k]);
        dft5(v, o);
#pragma unroll
        for (int q = 0; q < 5; ++q) z[zidx(base + 8 * q)] = o[q];
    }
    __syncthreads();

    // ---- pass 3: radix-5, 40 -> 200.  twiddle W_200^{jk} = W_400^{2 j k}
    for (int item = tid; item < FPB * 40; item += THREADS) {
        const int f = item / 40, k = item - f * 40;
        float2* z = s_z + f * ZPITCH;
        float2 v[5], o[5];
        v[0] = z[zidx(k)];
#pragma unroll
        for (int j = 1; j < 5; ++j) v[j] = cmul(z[zidx(k + 40 * j)], s_tw[2 * j * k]);
        dft5(v, o);
#pragma unroll
        for (int q = 0; q < 5; ++q) z[zidx(k + 40 * q)] = o[q];
    }
    __syncthreads();

    // ---- real-input split + power:  X[k] = E[k] + W_400^k O[k],  E = (Z[k] + conj Z[200-k]) / 2,  O = -i (Z[k] - conj Z[200-k]) / 2
    for (int item = tid; item < FPB * NBINS; item += THREADS) {
        const int f = item & (FPB - 1), k = item >> 4;
        const float2* z = s_z + f * ZPITCH;
        const float2 zk = z[zidx(k == NZ ? 0 : k)];
        float2 zc = z[zidx(k == 0 ? 0 : NZ - k)];
        zc.y = -zc.y;
        const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
        const float2 d = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
        const float2 x = cadd(e, cmul(s_tw[k], mul_mi(d)));
        s_p[k * FPB + f] = x.x * x.x + x.y * x.y;
    }
    __syncthreads();

    // ---- banded filterbank + log10 + running max.  thread = (mel j, group of 8 frames)
    float lmax = -INFINITY;
    const int groups = FPB / 8;
    for (int item = tid; item < n_mel * groups; item += THREADS) {
        const int j = item % n_mel, g = item / n_mel;
        const int2 rg = ranges[j];
        const float* frow = filters + static_cast<size_t>(j) * NBINS;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = rg.x; k < rg.y; ++k) {
            const float w = __ldg(frow + k);
            const float4 p0 = *reinterpret_cast<const float4*>(s_p + k * FPB + g * 8);
            const float4 p1 = *reinterpret_cast<const float4*>(s_p + k * FPB + g * 8 + 4);
            acc[0] = fmaf(w, p0.x, acc[0]); acc[1] = fmaf(w, p0.y, acc[1]);
            acc[2] = fmaf(w, p0.z, acc[2]); acc[3] = fmaf(w, p0.w, acc[3]);
            acc[4] = fmaf(w, p1.x, acc[4]); acc[5] = fmaf(w, p1.y, acc[5]);
            acc[6] = fmaf(w, p1.z, acc[6]); acc[7] = fmaf(w, p1.w, acc[7]);
        }
        float* orow = out + static_cast<size_t>(j) * ld_frames + f0 + g * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float v = acc[i] > 1e-10f ? log10f(acc[i]) : -10.0f;
            acc[i] = v;
            if (f0 + g * 8 + i < n_frames) lmax = fmaxf(lmax, v);
        }
        if (f0 + g * 8 + 8 <= n_frames && (ld_frames & 3) == 0) {
            *reinterpret_cast<float4*>(orow) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4*>(orow + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (f0 + g * 8 + i < n_frames) orow[i] = acc[i];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = lmax;
    __syncthreads();
    if (tid == 0) {
        float m = s_red[0];
#pragma unroll
        for (int i = 1; i < THREADS / 32; ++i) m = fmaxf(m, s_red[i]);
        if (m > -INFINITY) atomicMax(&win_max_key[b], float_to_key(m));
    }
}

__global__ void __launch_bounds__(256)
mel_normalize_kernel(float* __restrict__ mel, int ld_frames, int n_frames, int n_mel, const int* __restrict__ win_max_key) {
    const int b = blockIdx.y;
    const float thr = key_to_float(win_max_key[b]) - 8.0f;
    float* mb = mel + static_cast<size_t>(b) * n_mel * ld_frames;
    const size_t total = static_cast<size_t>(n_mel) * n_frames;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t j = i / n_frames, f = i - j * n_frames;
        float* p = mb + j * ld_frames + f;
        *p = (fmaxf(*p, thr) + 4.0f) * 0.25f;
    }
}

constexpr size_t MEL_SMEM = P_BYTES + sizeof(float2) * (N_FFT + FPB * ZPITCH);

}  // namespace

cudaError_t mel_plan_create(MelPlan** plan_out, const float* filters_host, int n_mel, int n_bins, cudaStream_t st) {
    if (!plan_out || !filters_host || n_mel <= 0 || n_bins != NBINS) return cudaErrorInvalidValue;
    MelPlan* p = new MelPlan();
    p->n_mel = n_mel;
    p->n_bins = n_bins;
    std::vector<int2> ranges(n_mel);
    for (int j = 0; j < n_mel; ++j) {
        int lo = n_bins, hi = 0;
        for (int k = 0; k < n_bins; ++k) {
            if (filters_host[static_cast<size_t>(j) * n_bins + k] != 0.0f) {
                if (k < lo) lo = k;
                hi = k + 1;
            }
        }
        if (hi == 0) lo = 0;
        ranges[j] = make_int2(lo, hi);
    }
    // Hann window exactly as the reference builds it (:2428-2436): 0.5 * (1.0 - cosf(2 pi i / 400)) evaluated in double
    std::vector<float> hann(N_FFT);
    for (int i = 0; i < N_FFT; ++i) hann[i] = static_cast<float>(0.5 * (1.0 - cosf(static_cast<float>((2.0 * M_PI * i) / N_FFT))));
    std::vector<float2> tw(N_FFT);
    for (int i = 0; i < N_FFT; ++i) {
        const double th = -2.0 * M_PI * i / N_FFT;
        tw[i] = make_float2(static_cast<float>(cos(th)), static_cast<float>(sin(th)));
    }
    cudaError_t e;
#define Q2W_TRY(x) do { e = (x); if (e != cudaSuccess) { mel_plan_destroy(p); return e; } } while (0)
    Q2W_TRY(cudaMalloc(&p->filters, sizeof(float) * n_mel * n_bins));
    Q2W_TRY(cudaMalloc(&p->ranges, sizeof(int2) * n_mel));
    Q2W_TRY(cudaMalloc(&p->hann, sizeof(float) * N_FFT));
    Q2W_TRY(cudaMalloc(&p->tw, sizeof(float2) * N_FFT));
    Q2W_TRY(cudaMemcpyAsync(p->filters, filters_host, sizeof(float) * n_mel * n_bins, cudaMemcpyHostToDevice, st));
    Q2W_TRY(cudaMemcpyAsync(p->ranges, ranges.data(), sizeof(int2) * n_mel, cudaMemcpyHostToDevice, st));
    Q2W_TRY(cudaMemcpyAsync(p->hann, hann.data(), sizeof(float) * N_FFT, cudaMemcpyHostToDevice, st));
    Q2W_TRY(cudaMemcpyAsync(p->tw, tw.data(), sizeof(float2) * N_FFT, cudaMemcpyHostToDevice, st));
    Q2W_TRY(cudaStreamSynchronize(st));  // host vectors go out of scope
    Q2W_TRY(cudaFuncSetAttribute(mel_logpower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(MEL_SMEM)));
#undef Q2W_TRY
    *plan_out = p;
    return cudaSuccess;
}

void mel_plan_destroy(MelPlan* p) {
    if (!p) return;
    cudaFree(p->filters);
    cudaFree(p->ranges);
    cudaFree(p->hann);
    cudaFree(p->tw);
    delete p;
}

cudaError_t mel_logpower(const MelPlan* plan, const float* pcm, size_t pcm_stride, const int* n_samples_dev, int n_max,
                         int B, int n_frames, float* logmel, int ld_frames, float* win_max, cudaStream_t st) {
    if (!plan || B <= 0 || n_frames <= 0 || ld_frames < n_frames) return cudaErrorInvalidValue;
    int* keys = reinterpret_cast<int*>(win_max);
    init_keys_kernel<<<(B + 255) / 256, 256, 0, st>>>(keys, B);
    dim3 grid((n_frames + FPB - 1) / FPB, B);
    mel_logpower_kernel<<<grid, THREADS, MEL_SMEM, st>>>(plan->filters, plan->ranges, plan->hann, plan->tw, plan->n_mel, pcm,
                                                         pcm_stride, n_samples_dev, n_max, n_frames, logmel, ld_frames, keys);
    return cudaGetLastError();
}

cudaError_t mel_normalize(float* logmel, int ld_frames, int n_frames, int n_mel, const float* win_max, int B,
                          cudaStream_t st) {
    if (B <= 0 || n_frames <= 0) return cudaErrorInvalidValue;
    const size_t total = static_cast<size_t>(n_mel) * n_frames;
    unsigned gx = static_cast<unsigned>((total + 255) / 256);
    if (gx > 148 * 8) gx = 148 * 8;
    mel_normalize_kernel<<<dim3(gx, B), 256, 0, st>>>(logmel, ld_frames, n_frames, n_mel, reinterpret_cast<const int*>(win_max));
    return cudaGetLastError();
}

}  // namespace q2w
