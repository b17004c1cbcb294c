// elementwise.cu -- HBM-bound kernels of the encoder path: LayerNorm, avg-pool + final LayerNorm,
// conv-stem im2col operands, ggml block decode (Q8_0 / Q4_0 / F32 -> f16).
//
// Reference semantics (file:line in /root/reference):
//   ggml_norm                ggml/src/ggml.c:11941-11990  mean / variance two-pass, 1/sqrtf(var + eps); gamma/beta by
//                            separate mul/add nodes (src/qwen2-whisper.cpp:2019-2024, :2128-2133, :2175-2180)
//   ggml_pool_1d(AVG,2,2,0)  ggml/src/ggml.c:15077-15125  (src/qwen2-whisper.cpp:2160-2171)
//   im2col_f32               ggml/src/ggml.c:14717        column index = ic*K + k, zero padding
//   dequantize_row_q8_0      ggml/src/ggml-quants.c:1616 ; dequantize_row_q4_0 :1522 ; block layouts ggml-common.h:144,186
#include "launch.cuh"
#include "ops.h"
#include "ptx.cuh"

#include <cstdint>

namespace q2w {

namespace {

constexpr int LN_MAX_VPL = 10;  // float4 per lane -> D <= 1280

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per row; the row lives in registers between the two statistics passes
template <bool POOL, typename OutT>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 OutT* __restrict__ y, int rows_out, int D, float eps, int T /*POOL: input rows per window*/,
                 __half* __restrict__ y16 /*optional second output in F16 (the projector's A operand), F32-output variant only*/) {
    pdl_wait();               // predecessor grid complete + visible (launch.cuh)
    pdl_launch_dependents();
    const int widx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (widx >= rows_out) return;
    // Rows are taken LAST FIRST. The producer (a persistent GEMM walking its tiles in row order) wrote the last rows of x most recently,
    // so at large M (x = 491 MB at 64 windows, L2 = 126 MB) they are the ones still in L2; and the consumer GEMM starts at row 0, which
    // this order writes last. Same bytes, fewer of them from HBM.
    const int warp = rows_out - 1 - widx;
    const int nvec = D >> 2;
    const float4* src0;
    const float4* src1 = nullptr;
    if constexpr (POOL) {
        const int To = T >> 1;
        const int b = warp / To, t = warp - b * To;
        const size_t r0 = static_cast<size_t>(b) * T + 2 * t;
        src0 = reinterpret_cast<const float4*>(x + r0 * D);
        src1 = reinterpret_cast<const float4*>(x + (r0 + 1) * D);
    } else {
        src0 = reinterpret_cast<const float4*>(x + static_cast<size_t>(warp) * D);
    }
    float4 v[LN_MAX_VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_VPL; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            float4 a = __ldcs(src0 + idx);
            if constexpr (POOL) {
                const float4 c = __ldcs(src1 + idx);
                a.x = (a.x + c.x) * 0.5f; a.y = (a.y + c.y) * 0.5f; a.z = (a.z + c.z) * 0.5f; a.w = (a.w + c.w) * 0.5f;
            }
            v[i] = a;
            s += (a.x + a.y) + (a.z + a.w);
        } else {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const float mean = warp_sum(s) / static_cast<float>(D);
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_VPL; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            s2 += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
    }
    const float var = warp_sum(s2) / static_cast<float>(D);
    const float rstd = 1.0f / sqrtf(var + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < LN_MAX_VPL; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            const float4 g = __ldg(g4 + idx), bb = __ldg(b4 + idx);
            const float o0 = v[i].x * rstd * g.x + bb.x;
            const float o1 = v[i].y * rstd * g.y + bb.y;
            const float o2 = v[i].z * rstd * g.z + bb.z;
            const float o3 = v[i].w * rstd * g.w + bb.w;
            if constexpr (sizeof(OutT) == 2) {
                __half2 h0 = __floats2half2_rn(o0, o1), h1 = __floats2half2_rn(o2, o3);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0);
                pk.y = *reinterpret_cast<uint32_t*>(&h1);
                reinterpret_cast<uint2*>(y + static_cast<size_t>(warp) * D)[idx] = pk;
            } else {
                reinterpret_cast<float4*>(y + static_cast<size_t>(warp) * D)[idx] = make_float4(o0, o1, o2, o3);
                if (y16 != nullptr) {
                    __half2 h0 = __floats2half2_rn(o0, o1), h1 = __floats2half2_rn(o2, o3);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&h0);
                    pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    reinterpret_cast<uint2*>(y16 + static_cast<size_t>(warp) * D)[idx] = pk;
                }
            }
        }
    }
}

// conv1 operand.  One block = 32 output frames of one window: stage [n_mel][34] normalised mel values in smem
// (coalesced along time), then emit 32 rows of 3*n_mel f16 (coalesced along the row).
__global__ void __launch_bounds__(256)
conv1_operand_kernel(const float* __restrict__ mel, int ld_frames, int n_frames_valid, int n_mel,
                     const float* __restrict__ win_max, int normalise, int offset, int n_ctx2, __half* __restrict__ A1) {
    pdl_wait();               // predecessor grid complete + visible (launch.cuh)
    pdl_launch_dependents();
    extern __shared__ float s_mel[];  // [n_mel][35]
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * 32;
    const float* mb = mel + static_cast<size_t>(b) * n_mel * ld_frames;
    float thr = 0.f;
    if (normalise) {  // per-window max is stored as an order-preserving integer key (mel.cu: float_to_key)
        const int key = reinterpret_cast<const int*>(win_max)[b];
        thr = __int_as_float(key >= 0 ? key : key ^ 0x7FFFFFFF) - 8.0f;
    }
    for (int idx = threadIdx.x; idx < n_mel * 34; idx += blockDim.x) {
        const int ic = idx / 34, j = idx - ic * 34;
        const int tt = t0 - 1 + j;
        float v = 0.f;
        if (tt >= 0 && tt < n_ctx2 && offset + tt < n_frames_valid) {
            v = mb[static_cast<size_t>(ic) * ld_frames + offset + tt];
            if (normalise) {
                v = fmaxf(v, thr);
                v = (v + 4.0f) * 0.25f;
            }
        }
        s_mel[ic * 35 + j] = v;
    }
    __syncthreads();
    const int ncol = 3 * n_mel;
    const int rows = min(32, n_ctx2 - t0);
    __half* out = A1 + (static_cast<size_t>(b) * n_ctx2 + t0) * ncol;
    for (int idx = threadIdx.x; idx < rows * ncol; idx += blockDim.x) {
        const int r = idx / ncol, col = idx - r * ncol;
        const int ic = col / 3, k = col - ic * 3;
        out[idx] = __float2half_rn(s_mel[ic * 35 + r + k]);
    }
}

// conv2 operand: A2[(b*T + t)][ic*3 + k] = h1[(b*T2 + 2t + k - 1)][ic], zero outside [0, T2).  8 outputs (16 B) per thread.
__global__ void __launch_bounds__(256)
conv2_im2col_kernel(const __half* __restrict__ h1, __half* __restrict__ A2, int B, int T2, int C) {
    pdl_wait();               // predecessor grid complete + visible (launch.cuh)
    pdl_launch_dependents();
    const int T = T2 >> 1;
    const int ncol = 3 * C;
    const int vec_per_row = ncol >> 3;
    const size_t total = static_cast<size_t>(B) * T * vec_per_row;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t row = i / vec_per_row;
        const int c0 = static_cast<int>(i - row * vec_per_row) << 3;
        const int b = static_cast<int>(row / T), t = static_cast<int>(row - static_cast<size_t>(b) * T);
        const __half* base = h1 + static_cast<size_t>(b) * T2 * C;
        __align__(16) __half o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = c0 + j;
            const int ic = col / 3, k = col - ic * 3;
            const int tt = 2 * t + k - 1;
            o[j] = (tt >= 0 && tt < T2) ? base[static_cast<size_t>(tt) * C + ic] : __float2half_rn(0.f);
        }
        *reinterpret_cast<uint4*>(A2 + row * ncol + c0) = *reinterpret_cast<const uint4*>(o);
    }
}

// ggml block decode, one thread per 32-element block (weights are L2-resident, output 64 B / thread)
__global__ void __launch_bounds__(256)
dequant_q8_0_kernel(const uint8_t* __restrict__ src, __half* __restrict__ dst, size_t nblocks) {
    pdl_wait();               // predecessor grid complete + visible (launch.cuh)
    pdl_launch_dependents();
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    const uint16_t* p = reinterpret_cast<const uint16_t*>(src + i * 34);  // {f16 d; int8 qs[32]}, 2-byte aligned
    const float d = __half2float(__ushort_as_half(p[0]));
    __align__(16) __half o[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint16_t w = p[1 + j];
        o[2 * j]     = __float2half_rn(static_cast<float>(static_cast<int8_t>(w & 0xFF)) * d);
        o[2 * j + 1] = __float2half_rn(static_cast<float>(static_cast<int8_t>(w >> 8)) * d);
    }
    uint4* out = reinterpret_cast<uint4*>(dst + i * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = reinterpret_cast<const uint4*>(o)[j];
}

__global__ void __launch_bounds__(256)
dequant_q4_0_kernel(const uint8_t* __restrict__ src, __half* __restrict__ dst, size_t nblocks) {
    pdl_wait();               // predecessor grid complete + visible (launch.cuh)
    pdl_launch_dependents();
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    const uint16_t* p = reinterpret_cast<const uint16_t*>(src + i * 18);  // {f16 d; u8 qs[16]}
    const float d = __half2float(__ushort_as_half(p[0]));
    __align__(16) __half o[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint16_t w = p[1 + j];
        const int q0 = w & 0xFF, q1 = w >> 8;  // qs[2j], qs[2j+1]
        o[2 * j]          = __float2half_rn(static_cast<float>((q0 & 0xF) - 8) * d);   // element 2j
        o[2 * j + 1]      = __float2half_rn(static_cast<float>((q1 & 0xF) - 8) * d);   // element 2j+1
        o[2 * j + 16]     = __float2half_rn(static_cast<float>((q0 >> 4) - 8) * d);    // element 2j+16
        o[2 * j + 17]     = __float2half_rn(static_cast<float>((q1 >> 4) - 8) * d);    // element 2j+17
    }
    uint4* out = reinterpret_cast<uint4*>(dst + i * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = reinterpret_cast<const uint4*>(o)[j];
}

// One launch decodes up to four matrices (a whole encoder block: QKV | out | fc1 | fc2). Four threads per 32-element ggml block, one
// 16-byte store each: a warp writes 512 contiguous bytes (whole sectors), and reads its 8 blocks' 272 / 144 contiguous bytes.
// Values are bit-identical to dequant_q8_0 / q4_0_kernel above: F16(q * d) with one rounding.
template <int TYPE>
__global__ void __launch_bounds__(256)
dequant_multi_kernel(const DequantJob job, unsigned long long total_threads) {
    pdl_wait();
    pdl_launch_dependents();
    // grid-stride over a SMALL grid (two CTAs per SM): this kernel runs on the decode stream next to the compute stream's GEMMs, and a
    // grid that fills every thread slot of the machine keeps their CTAs (one per SM, 213 KB of shared memory) from launching at all
    for (unsigned long long gid = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; gid < total_threads;
         gid += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        unsigned long long blk = gid >> 2;
        const int part = static_cast<int>(gid & 3);
        int t = 0;
        while (t < 3 && blk >= job.nblocks[t]) { blk -= job.nblocks[t]; ++t; }
        __align__(16) __half o[8];
        if constexpr (TYPE == 8) {
            const uint16_t* p = reinterpret_cast<const uint16_t*>(job.src[t] + blk * 34);   // {f16 d; int8 qs[32]}
            const float d = __half2float(__ushort_as_half(__ldg(p)));
            uint16_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = __ldg(p + 1 + part * 4 + j);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o[2 * j]     = __float2half_rn(static_cast<float>(static_cast<int8_t>(w[j] & 0xFF)) * d);
                o[2 * j + 1] = __float2half_rn(static_cast<float>(static_cast<int8_t>(w[j] >> 8)) * d);
            }
        } else {
            const uint16_t* p = reinterpret_cast<const uint16_t*>(job.src[t] + blk * 18);   // {f16 d; u8 qs[16]}: element j low nibble of qs[j], j + 16 high
            const float d = __half2float(__ushort_as_half(__ldg(p)));
            const int sh = (part >> 1) * 4;
            uint16_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = __ldg(p + 1 + (part & 1) * 4 + j);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o[2 * j]     = __float2half_rn(static_cast<float>((((w[j] & 0xFF) >> sh) & 0xF) - 8) * d);
                o[2 * j + 1] = __float2half_rn(static_cast<float>((((w[j] >> 8) >> sh) & 0xF) - 8) * d);
            }
        }
        *reinterpret_cast<uint4*>(job.dst[t] + blk * 32 + part * 8) = *reinterpret_cast<const uint4*>(o);
    }
}

__global__ void __launch_bounds__(256)
f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
    pdl_wait();               // predecessor grid complete + visible (launch.cuh)
    pdl_launch_dependents();
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        dst[i] = __float2half_rn(src[i]);
}

}  // namespace

cudaError_t layernorm_f32_to_f16(const float* x, const float* gamma, const float* beta, __half* y, int M, int D,
                                 float eps, cudaStream_t st) {
    if (D % 4 || D > LN_MAX_VPL * 128 || M <= 0) return cudaErrorInvalidValue;
    const int warps_per_block = 8;
    const int grid = (M + warps_per_block - 1) / warps_per_block;
    return launch_pdl(layernorm_kernel<false, __half>, dim3(grid), dim3(warps_per_block * 32), 0, st, x, gamma, beta, y, M, D, eps, 0,
                      static_cast<__half*>(nullptr));
}

cudaError_t pool2_layernorm_f32(const float* x, const float* gamma, const float* beta, float* y, int B, int T, int D,
                                float eps, cudaStream_t st, __half* y16) {
    if (D % 4 || D > LN_MAX_VPL * 128 || B <= 0 || T < 2) return cudaErrorInvalidValue;
    const int rows_out = B * (T / 2);
    const int warps_per_block = 8;
    const int grid = (rows_out + warps_per_block - 1) / warps_per_block;
    return launch_pdl(layernorm_kernel<true, float>, dim3(grid), dim3(warps_per_block * 32), 0, st, x, gamma, beta, y, rows_out, D, eps, T, y16);
}

cudaError_t mel_to_conv1_operand(const float* mel, int ld_frames, int n_frames_valid, int n_mel, const float* win_max,
                                 int normalise, int offset, int n_ctx2, int B, __half* A1, cudaStream_t st) {
    if (B <= 0 || n_ctx2 <= 0 || n_mel <= 0 || (3 * n_mel) % 8) return cudaErrorInvalidValue;
    dim3 grid((n_ctx2 + 31) / 32, B);
    const size_t smem = static_cast<size_t>(n_mel) * 35 * sizeof(float);
    return launch_pdl(conv1_operand_kernel, grid, dim3(256), smem, st, mel, ld_frames, n_frames_valid, n_mel, win_max, normalise, offset,
                      n_ctx2, A1);
}

cudaError_t conv2_im2col(const __half* h1, __half* A2, int B, int T2, int C, cudaStream_t st) {
    if (B <= 0 || T2 < 2 || (3 * C) % 8) return cudaErrorInvalidValue;
    const size_t total = static_cast<size_t>(B) * (T2 / 2) * (3 * C / 8);
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    return launch_pdl(conv2_im2col_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, h1, A2, B, T2, C);
}

cudaError_t dequant_multi_to_f16(const DequantJob& job, int ggml_type, cudaStream_t st) {
    unsigned long long total = 0;
    for (int i = 0; i < 4; ++i) total += job.nblocks[i];
    if (total == 0) return cudaSuccess;
    const unsigned long long threads = total * 4;
    DeviceInfo di;
    cudaError_t e = current_device_info(di);
    if (e != cudaSuccess) return e;
    const unsigned long long want = (threads + 255) / 256;
    const unsigned grid = static_cast<unsigned>(want < 2ull * di.num_sms ? want : 2ull * di.num_sms);
    switch (ggml_type) {
        case 8: return launch_pdl(dequant_multi_kernel<8>, dim3(grid), dim3(256), 0, st, job, threads);
        case 2: return launch_pdl(dequant_multi_kernel<2>, dim3(grid), dim3(256), 0, st, job, threads);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t dequant_to_f16(const void* src, int ggml_type, __half* dst, size_t rows, int K, cudaStream_t st) {
    if (ggml_type != 0 && (K % 32)) return cudaErrorInvalidValue;   // quantised rows are whole 32-element blocks
    const size_t nblocks = rows * static_cast<size_t>(K / 32);
    const unsigned grid = static_cast<unsigned>((nblocks + 255) / 256);
    switch (ggml_type) {
        case 8: return launch_pdl(dequant_q8_0_kernel, dim3(grid), dim3(256), 0, st, static_cast<const uint8_t*>(src), dst, nblocks);
        case 2: return launch_pdl(dequant_q4_0_kernel, dim3(grid), dim3(256), 0, st, static_cast<const uint8_t*>(src), dst, nblocks);
        case 0: {
            const size_t n = rows * static_cast<size_t>(K);
            size_t blocks = (n + 255) / 256;
            if (blocks > 148 * 32) blocks = 148 * 32;
            return launch_pdl(f32_to_f16_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, static_cast<const float*>(src), dst, n);
        }
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace q2w
