// gemm_tcgen05.cu -- persistent, warp-specialised F16 x F16 -> F32 GEMM for sm_100a.
//
// Replaces, on the encoder path, every weight ggml_mul_mat of the reference graph
// (src/qwen2-whisper.cpp:2029-2046 Q/K/V, :2112 out-proj, :2137 fc1, :2147 fc2, and the conv stem
// mul_mats produced by ggml_conv_1d_ph :1922/:1927) together with the elementwise nodes that follow
// them (ggml_add bias, ggml_scale, ggml_gelu, residual ggml_add, positional-embedding add).
//
// Design (one CTA per SM, 192 threads):
//   warp 0   : TMA producer   -- cp.async.bulk.tensor 2D loads of A (128 x 64) and W (256 x 64) f16 tiles,
//                                SWIZZLE_128B, 4-stage mbarrier ring
//   warp 1   : MMA issuer     -- one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M128 N256 K16),
//                                accumulators live in TMEM, double-buffered (2 x 256 columns)
//   warps 2-5: epilogue       -- tcgen05.ld 32x32b.x32 -> registers -> bias / scale / GELU / residual / pos -> global
// The epilogue of tile i overlaps the main loop of tile i+1 through the two TMEM accumulator stages.
// M, N, K tails are handled by TMA zero-fill on the load side and predication on the store side.
#include "ops.h"
#include "ptx.cuh"

#include <atomic>
#include <mutex>

namespace q2w {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_BYTES = BN * BK * 2;   // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int UMMA_K = 16;

struct KParams {
    int M, N, K;
    const float* bias;
    void* out;
    int ldo;
    const float* resid;
    const float* pos;
    int pos_period;
    int scale_cols;
    float scale;
    int m_tiles, n_tiles;
};

__device__ __forceinline__ float gelu_tanh(float x) {
    // 0.5 x (1 + tanh(u)) == x * sigmoid(2u),  u = sqrt(2/pi) x (1 + 0.044715 x^2)   (ggml.c:2541-2547)
    const float u = 0.79788456080286535588f * x * (1.0f + 0.044715f * x * x);
    const float e = __expf(-2.0f * u);
    return __fdividef(x, 1.0f + e);
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full_bar = bars;                    // [STAGES]
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.m_tiles * p.n_tiles;
    const int nkb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4);  // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / p.n_tiles) * BM;
                const int n0 = (tile % p.n_tiles) * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * STAGE_BYTES;
                    uint8_t* sB = sA + A_BYTES;
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tma_load_2d(sA, &tmA, &full_bar[stage], kb * BK, m0);
                    tma_load_2d(sB, &tmB, &full_bar[stage], kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t a_desc = make_sw128_kmajor_desc(a_addr);
                    const uint64_t b_desc = make_sw128_kmajor_desc(a_addr + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 16 elements (32 B) along K inside the 128-byte swizzle row: +2 in (addr >> 4) units
                        umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);        // accumulator complete -> epilogue
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;            // TMEM lane quadrant this warp may access
        const int row_in_tile = q * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / p.n_tiles) * BM;
            const int n0 = (tile % p.n_tiles) * BN;
            const int m = m0 + row_in_tile;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
            const float* pos_row = nullptr;
            if constexpr (EPI == EPI_BIAS_GELU_POS_F32) {
                pos_row = p.pos + static_cast<size_t>(m % p.pos_period) * p.N;
            }
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                const int n = n0 + c * 32;
                if (n >= p.N) break;  // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32b_x32(t_row + c * 32, r);
                tmem_ld_wait();
                if (m < p.M) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {  // 8 columns per group
                        const int ng = n + g * 8;
                        if (ng < p.N) {
                            float v[8];
                            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                            if (p.bias) {
                                b0 = __ldg(reinterpret_cast<const float4*>(p.bias + ng));
                                b1 = __ldg(reinterpret_cast<const float4*>(p.bias + ng + 4));
                            }
                            v[0] = __uint_as_float(r[g * 8 + 0]) + b0.x;
                            v[1] = __uint_as_float(r[g * 8 + 1]) + b0.y;
                            v[2] = __uint_as_float(r[g * 8 + 2]) + b0.z;
                            v[3] = __uint_as_float(r[g * 8 + 3]) + b0.w;
                            v[4] = __uint_as_float(r[g * 8 + 4]) + b1.x;
                            v[5] = __uint_as_float(r[g * 8 + 5]) + b1.y;
                            v[6] = __uint_as_float(r[g * 8 + 6]) + b1.z;
                            v[7] = __uint_as_float(r[g * 8 + 7]) + b1.w;
                            if constexpr (EPI == EPI_BIAS_F16) {
                                if (ng < p.scale_cols) {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v[i] *= p.scale;
                                }
                            }
                            if constexpr (EPI == EPI_BIAS_GELU_F16 || EPI == EPI_BIAS_GELU_POS_F32) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = gelu_tanh(v[i]);
                            }
                            if constexpr (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16) {
                                __half* o = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(m) * p.ldo + ng;
                                uint4 pk;
                                __half2 h0 = __floats2half2_rn(v[0], v[1]);
                                __half2 h1 = __floats2half2_rn(v[2], v[3]);
                                __half2 h2 = __floats2half2_rn(v[4], v[5]);
                                __half2 h3 = __floats2half2_rn(v[6], v[7]);
                                pk.x = *reinterpret_cast<uint32_t*>(&h0);
                                pk.y = *reinterpret_cast<uint32_t*>(&h1);
                                pk.z = *reinterpret_cast<uint32_t*>(&h2);
                                pk.w = *reinterpret_cast<uint32_t*>(&h3);
                                *reinterpret_cast<uint4*>(o) = pk;
                            } else {
                                float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(m) * p.ldo + ng;
                                if constexpr (EPI == EPI_BIAS_RESID_F32) {
                                    const float* rs = p.resid + static_cast<size_t>(m) * p.ldo + ng;
                                    const float4 r0 = *reinterpret_cast<const float4*>(rs);
                                    const float4 r1 = *reinterpret_cast<const float4*>(rs + 4);
                                    v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
                                    v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
                                }
                                if constexpr (EPI == EPI_BIAS_GELU_POS_F32) {
                                    const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos_row + ng));
                                    const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos_row + ng + 4));
                                    v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w;
                                    v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
                                }
                                *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                                *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
                            }
                        }
                    }
                }
            }
            // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<PFN_encodeTiled>(f);
        }
    });
    return fn;
}

// 2-D f16 row-major [rows, cols] with leading dimension ld (elements); box = box_rows x 64 columns, 128B swizzle
bool make_tmap_f16_2d(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * sizeof(__half)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

std::atomic<int> g_launches{0};
int g_num_sms = 0;

template <int EPI>
cudaError_t launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& kp, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int tiles = kp.m_tiles * kp.n_tiles;
    const int grid = tiles < g_num_sms ? tiles : g_num_sms;
    gemm_kernel<EPI><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmA, tmB, kp);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

}  // namespace

int gemm_num_launches() { return g_launches.load(); }

cudaError_t gemm_f16_tcgen05(const GemmArgs& a, GemmEpilogue epi, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0 || a.K <= 0) return cudaErrorInvalidValue;
    if ((a.K % 8) || (a.N % 8) || (a.lda % 8) || (a.ldw % 8) || (a.ldo % 8)) return cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.W) | reinterpret_cast<uintptr_t>(a.out)) & 15)
        return cudaErrorMisalignedAddress;
    CUtensorMap tmA, tmB;
    if (!make_tmap_f16_2d(&tmA, a.A, a.M, a.K, a.lda, BM)) return cudaErrorInvalidValue;
    if (!make_tmap_f16_2d(&tmB, a.W, a.N, a.K, a.ldw, BN)) return cudaErrorInvalidValue;
    KParams kp;
    kp.M = a.M; kp.N = a.N; kp.K = a.K;
    kp.bias = a.bias; kp.out = a.out; kp.ldo = a.ldo; kp.resid = a.resid;
    kp.pos = a.pos; kp.pos_period = a.pos_period > 0 ? a.pos_period : 1;
    kp.scale_cols = a.scale_cols; kp.scale = a.scale;
    kp.m_tiles = (a.M + BM - 1) / BM;
    kp.n_tiles = (a.N + BN - 1) / BN;
    switch (epi) {
        case EPI_BIAS_F16:          return launch<EPI_BIAS_F16>(tmA, tmB, kp, st);
        case EPI_BIAS_GELU_F16:     return launch<EPI_BIAS_GELU_F16>(tmA, tmB, kp, st);
        case EPI_BIAS_RESID_F32:
            if (!a.resid) return cudaErrorInvalidValue;
            return launch<EPI_BIAS_RESID_F32>(tmA, tmB, kp, st);
        case EPI_BIAS_GELU_POS_F32:
            if (!a.pos) return cudaErrorInvalidValue;
            return launch<EPI_BIAS_GELU_POS_F32>(tmA, tmB, kp, st);
        case EPI_BIAS_F32:          return launch<EPI_BIAS_F32>(tmA, tmB, kp, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace q2w
