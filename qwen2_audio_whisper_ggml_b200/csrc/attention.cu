// attention.cu -- fused non-causal self-attention over one window's T frames, head_dim 64.
//
// Replaces the reference's per-head  KQ = mul_mat(K, Q) ; soft_max(KQ) ; mul_mat(V, KQ_soft_max)  chain and the
// permute/cont copies around it (/root/reference/src/qwen2-whisper.cpp:2052-2106; the flash_attn branch there is
// commented out).  Q arrives already multiplied by 1/sqrt(64) (ggml_scale at :2054 is folded into the
// QKV GEMM epilogue).  Softmax follows ggml_soft_max (ggml/src/ggml.c:13854-13940): scale 1, no mask, row max
// subtracted, normalised by the row sum -- evaluated online (flash style) in FP32, probabilities rounded to F16
// for the PV product, FP32 accumulation.
//
// v0: one CTA = 64 query rows of one (window, head); 4 warps x 16 rows; K/V tiles of 64 rows double-buffered with
// cp.async into XOR-swizzled shared memory; ldmatrix + mma.sync.m16n8k16.  (The tcgen05/TMEM version replaces this.)
#include "ops.h"

#include <cstdint>

namespace q2w {

namespace {

constexpr int HD = 64, BQ = 64, BKV = 64, ATT_THREADS = 128;
constexpr int TILE_BYTES = 64 * 128;  // 64 rows x 64 f16

__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// load a 64 x 64 f16 tile (rows r0.. of a [T, ld] matrix, 64 columns at col0) into swizzled smem; rows >= T zero-filled
__device__ __forceinline__ void load_tile(uint32_t s_base, const __half* g, int ld, int r0, int T, int tid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = tid + i * ATT_THREADS;  // 512 chunks of 16 B
        const int row = c >> 3, chunk = c & 7;
        const int gr = r0 + row;
        const bool valid = gr < T;
        const __half* src = g + static_cast<size_t>(valid ? gr : (T - 1)) * ld + chunk * 8;
        cp_async16(s_base + tile_off(row, chunk), src, valid);
    }
}

__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int T, int D) {
    __shared__ __align__(128) uint8_t s_q[TILE_BYTES];
    __shared__ __align__(128) uint8_t s_k[2][TILE_BYTES];
    __shared__ __align__(128) uint8_t s_v[2][TILE_BYTES];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    const int ld = 3 * D;
    const __half* gq = qkv + static_cast<size_t>(b) * T * ld + h * HD;
    const __half* gk = gq + D;
    const __half* gv = gq + 2 * D;
    const uint32_t sq = static_cast<uint32_t>(__cvta_generic_to_shared(s_q));
    const uint32_t sk[2] = {static_cast<uint32_t>(__cvta_generic_to_shared(s_k[0])),
                            static_cast<uint32_t>(__cvta_generic_to_shared(s_k[1]))};
    const uint32_t sv[2] = {static_cast<uint32_t>(__cvta_generic_to_shared(s_v[0])),
                            static_cast<uint32_t>(__cvta_generic_to_shared(s_v[1]))};

    const int n_kv = (T + BKV - 1) / BKV;
    load_tile(sq, gq, ld, q0, T, tid);
    load_tile(sk[0], gk, ld, 0, T, tid);
    load_tile(sv[0], gv, ld, 0, T, tid);
    cp_async_commit();

    uint32_t qa[4][4];
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    constexpr float LOG2E = 1.4426950408889634f;

    for (int it = 0; it < n_kv; ++it) {
        const int st = it & 1;
        if (it + 1 < n_kv) {
            load_tile(sk[st ^ 1], gk, ld, (it + 1) * BKV, T, tid);
            load_tile(sv[st ^ 1], gv, ld, (it + 1) * BKV, T, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        if (it == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int chunk = 2 * ks + (lane >> 4);
                ldsm_x4(qa[ks], sq + tile_off(row, chunk));
            }
        }

        // ---- S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int nt2 = 0; nt2 < 4; ++nt2) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int mi = lane >> 3;
                const int row = nt2 * 16 + (lane & 7) + (mi >> 1) * 8;
                const int chunk = 2 * ks + (mi & 1);
                uint32_t kb[4];
                ldsm_x4(kb, sk[st] + tile_off(row, chunk));
                mma16816(s[2 * nt2], qa[ks], kb[0], kb[1]);
                mma16816(s[2 * nt2 + 1], qa[ks], kb[2], kb[3]);
            }
        }
        // ---- mask the ragged last tile
        const int kv0 = it * BKV;
        if (kv0 + BKV > T) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int c = kv0 + nt * 8 + 2 * (lane & 3);
                if (c >= T)     { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
                if (c + 1 >= T) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            }
        }
        // ---- online softmax (base-2 domain)
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float alpha[2], mnew[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mnew[r] = fmaxf(m_run[r], mx[r] * LOG2E);
            alpha[r] = exp2f(m_run[r] - mnew[r]);
            m_run[r] = mnew[r];
        }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = exp2f(fmaf(s[nt][0], LOG2E, -mnew[0]));
            s[nt][1] = exp2f(fmaf(s[nt][1], LOG2E, -mnew[0]));
            s[nt][2] = exp2f(fmaf(s[nt][2], LOG2E, -mnew[1]));
            s[nt][3] = exp2f(fmaf(s[nt][3], LOG2E, -mnew[1]));
            rs[0] += s[nt][0] + s[nt][1];
            rs[1] += s[nt][2] + s[nt][3];
        }
        l_run[0] = l_run[0] * alpha[0] + rs[0];
        l_run[1] = l_run[1] * alpha[1] + rs[1];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            o[nt][0] *= alpha[0]; o[nt][1] *= alpha[0];
            o[nt][2] *= alpha[1]; o[nt][3] *= alpha[1];
        }
        // ---- O += P V
#pragma unroll
        for (int ks2 = 0; ks2 < 4; ++ks2) {
            uint32_t pa[4];
            pa[0] = pack_h2(s[2 * ks2][0], s[2 * ks2][1]);
            pa[1] = pack_h2(s[2 * ks2][2], s[2 * ks2][3]);
            pa[2] = pack_h2(s[2 * ks2 + 1][0], s[2 * ks2 + 1][1]);
            pa[3] = pack_h2(s[2 * ks2 + 1][2], s[2 * ks2 + 1][3]);
#pragma unroll
            for (int nt2 = 0; nt2 < 4; ++nt2) {
                const int mi = lane >> 3;
                const int row = ks2 * 16 + (lane & 7) + (mi & 1) * 8;
                const int chunk = 2 * nt2 + (mi >> 1);
                uint32_t vb[4];
                ldsm_x4_t(vb, sv[st] + tile_off(row, chunk));
                mma16816(o[2 * nt2], pa, vb[0], vb[1]);
                mma16816(o[2 * nt2 + 1], pa, vb[2], vb[3]);
            }
        }
        __syncthreads();  // stage st is overwritten by the prefetch of iteration it+1
    }

    // ---- finalise: divide by the row sum, write f16
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
    const int r0 = q0 + warp * 16 + (lane >> 2);
    __half* ob = out + static_cast<size_t>(b) * T * D + h * HD + 2 * (lane & 3);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        if (r0 < T)
            *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r0) * D + nt * 8) = pack_h2(o[nt][0] * inv0, o[nt][1] * inv0);
        if (r0 + 8 < T)
            *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r0 + 8) * D + nt * 8) = pack_h2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
}

}  // namespace

cudaError_t attention_f16(const __half* qkv, __half* out, int B, int T, int H, cudaStream_t st) {
    if (B <= 0 || T <= 0 || H <= 0) return cudaErrorInvalidValue;
    const int D = H * HD;
    dim3 grid((T + BQ - 1) / BQ, H, B);
    attention_kernel<<<grid, ATT_THREADS, 0, st>>>(qkv, out, T, D);
    return cudaGetLastError();
}

}  // namespace q2w
