// gemm_tcgen05.cu -- persistent, warp-specialised F16 x F16 -> F32 GEMM for sm_100a.
//
// Replaces, on the encoder path, every weight ggml_mul_mat of the reference graph
// (src/qwen2-whisper.cpp:2029-2046 Q/K/V, :2112 out-proj, :2137 fc1, :2147 fc2, and the conv stem
// mul_mats produced by ggml_conv_1d_ph :1922/:1927) together with the elementwise nodes that follow
// them (ggml_add bias, ggml_scale, ggml_gelu, residual ggml_add, positional-embedding add).
//
// Design (one CTA per SM, 320 threads):
//   warp 0   : TMA producer   -- cp.async.bulk.tensor 2D loads of A (128 x 64) and W (256 x 64) f16 tiles,
//                                SWIZZLE_128B, 4-stage mbarrier ring
//   warp 1   : MMA issuer     -- one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M128 N256 K16),
//                                accumulators live in TMEM, double-buffered (2 x 256 columns)
//   warps 2-9: epilogue       -- two warps per TMEM lane quadrant (128 columns each): tcgen05.ld 32x32b.x32 -> registers ->
//                                smem transpose -> bias / scale / GELU / residual / pos -> coalesced global stores
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
constexpr int EPI_WARPS = 8;
constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * 33 * 4;  // per-epilogue-warp 32 x 33 f32 transpose buffers
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_STAGE_BYTES;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
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
            mbar_init(&tempty_bar[s], EPI_WARPS);  // one arrive per epilogue warp
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
        // ------------------------------------------------------------ epilogue (warps 2..9)
        // TMEM -> registers (lane = row) -> per-warp smem transpose (pitch 33, conflict-free both ways) -> lane = column:
        // every global access below is a contiguous row segment (128 B per warp instruction), bias / residual / positional
        // reads included, and the residual rows of chunk c+1 are in flight while chunk c is processed.
        // (Round-1 profile: row-per-lane stores cost 32 sectors per request and made the K=1280 GEMMs epilogue-bound.)
        const int q = warp & 3;                    // TMEM lane quadrant this warp may access
        const int chalf = (warp - 2) >> 2;         // which 128-column half of the tile
        float* stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256) + (warp - 2) * (32 * 33);
        int acc = 0;
        uint32_t acc_phase = 0;
        constexpr bool F16OUT = (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16);
        constexpr bool HAS_ADD = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_GELU_POS_F32);
        constexpr int CHUNKS = BN / 32 / 2;        // chunks per warp
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / p.n_tiles) * BM + q * 32;
            const int n0 = (tile % p.n_tiles) * BN + chalf * (BN / 2);
            const int rows = min(32, p.M - m0);       // valid rows of this warp's quadrant (may be <= 0)
            int pm = 0;
            if constexpr (EPI == EPI_BIAS_GELU_POS_F32) pm = m0 % p.pos_period;
            // rows of the residual / positional operand for one 32-column chunk: 32 independent coalesced loads per lane
            auto load_add = [&](int n, float (&add)[32]) {
                const int nc = n + lane;
                const bool ok = nc < p.N;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    add[i] = 0.f;
                    if (i < rows && ok) {
                        if constexpr (EPI == EPI_BIAS_RESID_F32) add[i] = p.resid[static_cast<size_t>(m0 + i) * p.ldo + nc];
                        if constexpr (EPI == EPI_BIAS_GELU_POS_F32) {
                            int pr = pm + i;
                            if (pr >= p.pos_period) pr -= p.pos_period;
                            add[i] = __ldg(p.pos + static_cast<size_t>(pr) * p.N + nc);
                        }
                    }
                }
            };
            float add[32];
            if constexpr (HAS_ADD) {
                if (n0 < p.N) load_add(n0, add);      // does not depend on the accumulator: issued before the wait
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * BN + chalf * (BN / 2) + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < CHUNKS; ++c) {
                const int n = n0 + c * 32;
                if (n >= p.N) break;  // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32b_x32(t_row + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]);
                __syncwarp();
                if constexpr (F16OUT) {
                    // two rows per instruction: lanes 0-15 -> row rr, lanes 16-31 -> row rr+1, two columns per lane
                    const int half = lane >> 4, l2 = (lane & 15) * 2;
                    const int nc = n + l2;
                    const bool col_ok = nc < p.N;     // N % 8 == 0 -> the pair is valid together
                    float b0 = 0.f, b1 = 0.f;
                    if (p.bias && col_ok) { b0 = __ldg(p.bias + nc); b1 = __ldg(p.bias + nc + 1); }
                    float sc = 1.0f;
                    if constexpr (EPI == EPI_BIAS_F16) sc = (nc < p.scale_cols) ? p.scale : 1.0f;
                    __half* obase = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(m0) * p.ldo + nc;
#pragma unroll 4
                    for (int rr = 0; rr < 32; rr += 2) {
                        const int row = rr + half;
                        float v0 = stg[row * 33 + l2] + b0;
                        float v1 = stg[row * 33 + l2 + 1] + b1;
                        if constexpr (EPI == EPI_BIAS_F16) { v0 *= sc; v1 *= sc; }
                        if constexpr (EPI == EPI_BIAS_GELU_F16) { v0 = gelu_tanh(v0); v1 = gelu_tanh(v1); }
                        if (row < rows && col_ok) {
                            __half2 h = __floats2half2_rn(v0, v1);
                            *reinterpret_cast<__half2*>(obase + static_cast<size_t>(row) * p.ldo) = h;
                        }
                    }
                } else {
                    float addn[32];
                    if constexpr (HAS_ADD) {
                        if (c + 1 < CHUNKS && n + 32 < p.N) load_add(n + 32, addn);   // prefetch the next chunk's rows
                    }
                    const int nc = n + lane;
                    const bool col_ok = nc < p.N;
                    const float bv = (p.bias && col_ok) ? __ldg(p.bias + nc) : 0.f;
                    float* obase = reinterpret_cast<float*>(p.out) + static_cast<size_t>(m0) * p.ldo + nc;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float v = stg[i * 33 + lane] + bv;
                        if constexpr (EPI == EPI_BIAS_GELU_POS_F32) v = gelu_tanh(v);
                        if constexpr (HAS_ADD) v += add[i];
                        if (i < rows && col_ok) obase[static_cast<size_t>(i) * p.ldo] = v;
                    }
                    if constexpr (HAS_ADD) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) add[i] = addn[i];
                    }
                }
                __syncwarp();  // the staging buffer is rewritten by the next chunk
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
