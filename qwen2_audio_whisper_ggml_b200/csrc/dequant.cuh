// dequant.cuh -- ggml block decode in registers, shared by the GEMM's in-kernel dequant warpgroup (gemm_tcgen05.cu) and the decode
// warps that ride inside the attention kernel (attention_tcgen05.cu).
#pragma once
#include <cstdint>
#include <cuda_fp16.h>

namespace q2w {

constexpr int WT_F16 = 1, WT_Q4_0 = 2, WT_Q8_0 = 8;          // ggml_type values

// ---- ggml block decode helpers (ggml-common.h:144-148, :186-191; ggml-quants.c:1522-1540, :1616-1630)
// four biased bytes -> two half2 holding (1024 + b) exactly (0x64xx is 1024 + xx in F16), then subtract the bias and scale
__device__ __forceinline__ void dq4(uint32_t biased, __half2 bias, __half2 d2, uint32_t& o0, uint32_t& o1) {
    const uint32_t lo = __byte_perm(biased, 0x64646464u, 0x4140);
    const uint32_t hi = __byte_perm(biased, 0x64646464u, 0x4342);
    __half2 a = __hmul2(__hsub2(*reinterpret_cast<const __half2*>(&lo), bias), d2);
    __half2 b = __hmul2(__hsub2(*reinterpret_cast<const __half2*>(&hi), bias), d2);
    o0 = *reinterpret_cast<uint32_t*>(&a);
    o1 = *reinterpret_cast<uint32_t*>(&b);
}

template <int WT> struct DqTraits;
template <> struct DqTraits<WT_Q8_0> { static constexpr int WORDS = 17; static constexpr int BLOCK_BYTES = 34; };
template <> struct DqTraits<WT_Q4_0> { static constexpr int WORDS = 9;  static constexpr int BLOCK_BYTES = 18; };

// decode the two blocks of one k-step (64 elements) of one W row into 32 packed half2 words
template <int WT>
__device__ __forceinline__ void decode_row(const uint32_t (&w)[DqTraits<WT>::WORDS], uint32_t (&out)[32]) {
    if constexpr (WT == WT_Q8_0) {
        // bytes: [d0:2][q0:32][d1:2][q1:32]  ->  w[0] = d0 | q0[0..1], w[8] = q0[30..31] | d1, w[9..16] = q1 (aligned)
        const __half2 bias = __float2half2_rn(1152.0f);            // 1024 + 128
        const __half2 d0 = __half2half2(__ushort_as_half(static_cast<unsigned short>(w[0] & 0xFFFF)));
        const __half2 d1 = __half2half2(__ushort_as_half(static_cast<unsigned short>(w[8] >> 16)));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t q = __byte_perm(w[j], w[j + 1], 0x5432) ^ 0x80808080u;   // re-align block 0 by two bytes, bias to unsigned
            dq4(q, bias, d0, out[2 * j], out[2 * j + 1]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) dq4(w[9 + j] ^ 0x80808080u, bias, d1, out[16 + 2 * j], out[16 + 2 * j + 1]);
    } else {
        // bytes: [d0:2][qs0:16][d1:2][qs1:16]; element j = low nibble of qs[j], element j+16 = high nibble of qs[j]
        const __half2 bias = __float2half2_rn(1032.0f);            // 1024 + 8
        const __half2 d0 = __half2half2(__ushort_as_half(static_cast<unsigned short>(w[0] & 0xFFFF)));
        const __half2 d1 = __half2half2(__ushort_as_half(static_cast<unsigned short>(w[4] >> 16)));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t q = __byte_perm(w[j], w[j + 1], 0x5432);
            dq4(q & 0x0F0F0F0Fu, bias, d0, out[2 * j], out[2 * j + 1]);                  // elements 4j .. 4j+3
            dq4((q >> 4) & 0x0F0F0F0Fu, bias, d0, out[8 + 2 * j], out[8 + 2 * j + 1]);   // elements 16+4j ..
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t q = w[5 + j];
            dq4(q & 0x0F0F0F0Fu, bias, d1, out[16 + 2 * j], out[16 + 2 * j + 1]);
            dq4((q >> 4) & 0x0F0F0F0Fu, bias, d1, out[24 + 2 * j], out[24 + 2 * j + 1]);
        }
    }
}


}  // namespace q2w
