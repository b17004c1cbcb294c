// gemm_tcgen05.cu -- persistent, warp-specialised F16 x F16 -> F32 GEMM for sm_100a on CTA PAIRS (cta_group::2).
//
// Replaces, on the encoder path, every weight ggml_mul_mat of the reference graph
// (src/qwen2-whisper.cpp:2029-2046 Q/K/V, :2112 out-proj, :2137 fc1, :2147 fc2, and the conv stem
// mul_mats produced by ggml_conv_1d_ph :1922/:1927) together with the elementwise nodes that follow
// them (ggml_add bias, ggml_scale, ggml_gelu, residual ggml_add, positional-embedding add).
//
// Design (cluster of 2 CTAs = one 256 x 256 output tile at a time, one cluster per SM pair, 320 threads per CTA):
//   warp 0    TMA producer (both CTAs): its own 128 x 64 slice of A and its own 128-row half of the 256 x 64 W tile
//             (cp.async.bulk.tensor.2d.cta_group::2, SWIZZLE_128B), completion bytes of BOTH CTAs land on the leader's
//             full barrier.  4-stage ring, 32 KB per stage per CTA -- each CTA fetches 1/3 less from L2 than a 128 x 256
//             single-CTA tile would, and the freed shared memory pays for the epilogue buffers below.
//   warp 1    MMA issuer (leader CTA only): tcgen05.mma.cta_group::2.kind::f16, M256 N256 K16; accumulators live in TMEM
//             (128 lanes x 256 columns per CTA), double-buffered; tcgen05.commit ... multicast::cluster releases the smem
//             stage / publishes the accumulator in both CTAs.
//   warps 10-13 (quantised weights only) in-kernel ggml block decode: W stays Q8_0 / Q4_0 in HBM exactly as in the model file; each
//             thread owns one row of this CTA's 128-row W half, reads its two 32-element blocks of the k-step (34 B / 18 B each,
//             one k-step ahead), decodes them with the integer->F16 magic-number trick (q*d rounded once to F16, bit-identical
//             to dequantize_row_q8_0 / q4_0 followed by an F16 store) and writes the 128-byte row into the SWIZZLE_128B tile
//             the MMA reads; the smem stage becomes "full" when the A bytes of both CTAs AND the 8 decode warps have arrived.
//   warps 2-9 epilogue, two per TMEM lane quadrant.  Everything stays in the "lane = output row" domain the accumulator
//             arrives in: tcgen05.ld -> bias (broadcast loads) / scale / GELU in registers -> result written into the
//             128B-swizzled smem chunk -> one TMA store per 32-row chunk; the residual epilogue issues a TMA *reduction* store
//             instead (cp.reduce.async.bulk.tensor ... add: out += acc + bias, summed at L2 -- exactly one add per element, so
//             still deterministic), which keeps the residual tile off the SM altogether.  No per-element global address
//             arithmetic, no predication: M/N tails are clipped by the TMA unit on store and zero-filled on load.  (Round-1 profiles: the per-element LDG/STG epilogue spent
//             ~45 % of its issue slots on 64-bit address math and kept the K = 1280 GEMMs at 45-55 % tensor-pipe activity.)
#include "ops.h"
#include "launch.cuh"
#include "ptx.cuh"
#include "dequant.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace q2w {

namespace {

constexpr int BM = 128;            // rows per CTA (256 per pair)
constexpr int BN = 256;            // columns per pair tile; each CTA stages BN/2 rows of W
constexpr int BK = 64;
// smem budget (227 KB): 2 chunk buffers per epilogue warp and a 5-stage ring. The residual epilogue is a TMA *reduction* store
// (out += acc + bias, added at L2): no residual tile ever travels to the SM, so it needs no third buffer either (the round-1
// scheme that TMA-loaded residual chunks into a 3-buffer rotation left room for 4 stages only and measured 22 % (fc2) / 34 %
// (out-proj) of the main loop waiting for operands; it is gone).
template <int EPI> struct Cfg {
    static constexpr bool F16OUT = (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16);
    static constexpr int STAGES = 5;
    static constexpr int EPI_BUFS = 2;
};
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int B_BYTES = (BN / 2) * BK * 2;    // 16 KB (this CTA's half of W)
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int CHUNK_BYTES = 32 * 128;         // 32 rows x 128 B (32 f32 or 64 f16 columns)
constexpr int BAR_BYTES = 512;
template <int EPI> constexpr int epi_bytes() { return EPI_WARPS * Cfg<EPI>::EPI_BUFS * CHUNK_BYTES; }   // 64 / 96 KB
template <int EPI> constexpr int smem_bytes() { return Cfg<EPI>::STAGES * STAGE_BYTES + epi_bytes<EPI>() + BAR_BYTES + 1024 /*align slack*/; }
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;          // F16 weights (W tile by TMA)
constexpr int DQ_WARPS = 4;
constexpr int NUM_THREADS_Q = NUM_THREADS + 32 * DQ_WARPS;   // quantised weights: + decode warpgroup
constexpr int TMEM_COLS = 512;
constexpr int UMMA_K = 16;

struct KParams {
    int M, N, K;
    const float* bias;
    const float* pos;
    int pos_period;
    int scale_cols;
    float scale;
    int m_tiles, n_tiles;   // in units of 256 x 256
    const uint8_t* wraw;    // quantised W: raw ggml blocks, row pitch wrow_bytes
    int wrow_bytes;
    int sk_upc;             // split-K (residual epilogue, small M only): k-block units per cluster; 0 = whole tiles, strided over the clusters
    int w_static;           // W is a model weight no kernel ever writes: its first tiles may be fetched before griddepcontrol.wait
};

// Work distribution.  Classic: cluster c owns the whole tiles c, c + C, c + 2C, ...  Split-K (sk_upc > 0): the (tile, k-block) units
// are numbered tile-major and cluster c owns the contiguous range [c * upc, (c + 1) * upc), cut into per-tile segments; every segment
// ends in its own reduction-store epilogue (out += partial, summed at L2), the segment that holds k-block 0 also adds the bias.
// All roles of a cluster (producer, MMA issuer, decode warps, epilogue) enumerate the same segments with this iterator.
struct WorkIter {
    int nkb, a, b, step;
    bool sk;
    __device__ __forceinline__ WorkIter(const KParams& p, int nkb_, int cluster_id, int num_clusters) {
        nkb = nkb_;
        sk = p.sk_upc > 0;
        const int num_tiles = p.m_tiles * p.n_tiles;
        if (sk) {
            a = cluster_id * p.sk_upc;
            b = min(a + p.sk_upc, num_tiles * nkb);
            step = 0;
        } else {
            a = cluster_id;
            b = num_tiles;
            step = num_clusters;
        }
    }
    __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1) {
        if (a >= b) return false;
        if (sk) {
            tile = a / nkb;
            kb0 = a - tile * nkb;
            const int n = min(nkb - kb0, b - a);
            kb1 = kb0 + n;
            a += n;
        } else {
            tile = a;
            kb0 = 0;
            kb1 = nkb;
            a += step;
        }
        return true;
    }
};

__device__ __forceinline__ float gelu_tanh(float x) {
    // 0.5 x (1 + tanh(u)),  u = sqrt(2/pi) x (1 + 0.044715 x^2)   (ggml.c:2541-2547), with the hardware tanh: ONE MUFU op per element
    // (the exp + reciprocal form needs two, and at 64 elements per thread per chunk the MUFU bounded the fc1 epilogue: ncu showed
    // the tensor pipe 74.8 % active on fc1 against 84.1 % on the QKV GEMM of the same K).  tanh.approx.f32 is accurate to 2^-11
    // relative, i.e. an absolute error <= 2.5e-4 |x| on the result -- below what the reference's own evaluation carries (it looks the
    // value up in an F16 table indexed by x rounded to F16, ggml.c:2556-2570: 4.9e-4 |x|), and below the F16 rounding of the output.
    const float u = x * fmaf(0.0356774081363001f, x * x, 0.79788456080286535588f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

#ifdef Q2W_GEMM_WALL
// wall-clock stamps (make EXTRA_NVFLAGS=-DQ2W_GEMM_WALL, tools/gemm_wall.py): where a small-M launch spends its time
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define WALL(var) const unsigned long long var = gtime()
#else
#define WALL(var)
#endif

template <int EPI, int WT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WT == WT_F16 ? NUM_THREADS : NUM_THREADS_Q, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmO, const KParams p) {
    constexpr int STAGES = Cfg<EPI>::STAGES;
    constexpr int EPI_BUFS = Cfg<EPI>::EPI_BUFS;
    constexpr int EPI_BYTES = epi_bytes<EPI>();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + EPI_BYTES);
    uint64_t* full_bar = bars;                       // [STAGES]  used in the leader CTA only
    uint64_t* empty_bar = bars + STAGES;             // [STAGES]  one per CTA, released by the leader's multicast commit
    uint64_t* tfull_bar = bars + 2 * STAGES;         // [2]       accumulator ready, multicast to both CTAs
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;    // [2]       leader only: 2 x EPI_WARPS arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int nkb = (p.K + BK - 1) / BK;
    WALL(w_entry);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], WT == WT_F16 ? 1 : 1 + 2 * DQ_WARPS);   // producer's expect_tx (+ decode warps of both CTAs)
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 2 * EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc_cta2(tmem_slot, TMEM_COLS);
        tmem_relinquish_cta2();
    }
    tc_fence_before();
    // The mbarrier inits were published cluster-wide by fence.mbarrier_init.release.cluster (the initialising thread, above); the TMEM base
    // address needs CTA-level ordering only. So: a CTA barrier, then an execution-only rendezvous of the pair -- not the release / acquire
    // cluster barrier, whose MEMBAR.ALL.GPU + ERRBAR drain sits on the critical path whenever this CTA could not start early.
    __syncthreads();
    cluster_sync_exec_only();    // both CTAs' barriers exist before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // The weights do not depend on the previous kernel: the W tiles of the first ring pass are requested BEFORE griddepcontrol.wait,
    // so their HBM round trip runs under the predecessor's tail (weights are read once per forward at small M: always a DRAM miss).
    int w_early = 0;             // (producer thread) k-blocks whose W tile and expect_tx are already issued
    if constexpr (WT == WT_F16) {
        if (warp == 0 && lane == 0 && p.w_static) {
            WorkIter wi(p, nkb, cluster_id, num_clusters);
            int tile, kb0, kb1;
            if (wi.next(tile, kb0, kb1)) {
                const int n0 = (tile % p.n_tiles) * BN + static_cast<int>(cta_rank) * (BN / 2);
                w_early = min(STAGES, kb1 - kb0);
                for (int i = 0; i < w_early; ++i) {
                    if (leader) mbar_expect_tx(&full_bar[i], 2 * STAGE_BYTES);
                    tma_load_2d_cta2(smem + i * STAGE_BYTES + A_BYTES, &tmB, smem_u32(&full_bar[i]) & kPeerBitMask, (kb0 + i) * BK, n0);
                }
            }
        }
    }
    WALL(w_prolog);
    pdl_wait();                  // everything above overlapped the previous kernel's tail; from here on its output is visible
    pdl_launch_dependents();
    WALL(w_wait);

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int issued = 0;
            WorkIter wi(p, nkb, cluster_id, num_clusters);
            int tile, kb0, kb1;
            while (wi.next(tile, kb0, kb1)) {
                const int m0 = (tile / p.n_tiles) * (2 * BM) + static_cast<int>(cta_rank) * BM;
                const int n0 = (tile % p.n_tiles) * BN + static_cast<int>(cta_rank) * (BN / 2);
                for (int kb = kb0; kb < kb1; ++kb, ++issued) {
                    uint8_t* sA = smem + stage * STAGE_BYTES;
                    uint8_t* sB = sA + A_BYTES;
                    const uint32_t full_leader = smem_u32(&full_bar[stage]) & kPeerBitMask;
                    if (issued < w_early) {          // first ring pass: stage free by construction, W and the byte count already in flight
                        tma_load_2d_cta2(sA, &tmA, full_leader, kb * BK, m0);
                    } else {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (leader) mbar_expect_tx(&full_bar[stage], WT == WT_F16 ? 2 * STAGE_BYTES : 2 * A_BYTES);   // TMA bytes of both CTAs
                        tma_load_2d_cta2(sA, &tmA, full_leader, kb * BK, m0);
                        if constexpr (WT == WT_F16) tma_load_2d_cta2(sB, &tmB, full_leader, kb * BK, n0);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();   // reconverge before the (aligned) cluster barrier below
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA)
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
#ifdef Q2W_GEMM_TIMELINE
            long long t_full = 0, t_acc = 0, n_slow = 0;
            int my_tiles = 0;
            const long long t_begin = clock64();
#endif
#ifdef Q2W_GEMM_WALL
            unsigned long long w_first = 0;
#endif
            WorkIter wi(p, nkb, cluster_id, num_clusters);
            int tile, kb0, kb1;
            while (wi.next(tile, kb0, kb1)) {
#ifdef Q2W_GEMM_TIMELINE
                long long w0 = clock64();
                ++my_tiles;
#endif
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
#ifdef Q2W_GEMM_TIMELINE
                t_acc += clock64() - w0;
#endif
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
#ifdef Q2W_GEMM_TIMELINE
                    w0 = clock64();
#endif
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
#ifdef Q2W_GEMM_WALL
                    if (w_first == 0) w_first = gtime();
#endif
#ifdef Q2W_GEMM_TIMELINE
                    { const long long dt = clock64() - w0; t_full += dt; if (dt > 200) ++n_slow; }
#endif
                    const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t a_desc = make_sw128_kmajor_desc(a_addr);
                    const uint64_t b_desc = make_sw128_kmajor_desc(a_addr + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_f16_ss_cta2(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb != kb0) || (k != 0));
                    umma_commit_cta2_mcast(&empty_bar[stage], 0x3);   // both CTAs may refill this stage
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_cta2_mcast(&tfull_bar[acc], 0x3);         // accumulator complete in both CTAs
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
#ifdef Q2W_GEMM_WALL
            if (cluster_id == 0 || cluster_id == num_clusters - 1)
                printf("GW mma c%d/%d N%d K%d entry %llu prolog %llu wait %llu first_operands %llu issued_all %llu\n", cluster_id, num_clusters, p.N, p.K, w_entry,
                       w_prolog, w_wait, w_first, gtime());
#endif
#ifdef Q2W_GEMM_TIMELINE
            if (cluster_id == 3 && my_tiles > 0) {
                const long long tot = clock64() - t_begin;
                printf("gemm M-tiles %d N-tiles %d K %d: MMA thread %lld clk for %d segments (%lld per segment, MMA floor %d per whole tile); waiting for operands %lld (%.1f %%, %lld k-blocks > 200 clk), for a free accumulator %lld (%.1f %%)\n",
                       p.m_tiles, p.n_tiles, p.K, tot, my_tiles, tot / my_tiles, nkb * 512, t_full, 100.0 * t_full / tot, n_slow, t_acc, 100.0 * t_acc / tot);
            }
#endif
        }
        __syncwarp();
    } else if (WT != WT_F16 && warp >= 2 + EPI_WARPS) {
        // ------------------------------------------------------------ in-kernel ggml block decode (warps 10..13, both CTAs)
        if constexpr (WT != WT_F16) {
            constexpr int WORDS = DqTraits<WT>::WORDS;
            const int row = (warp - 2 - EPI_WARPS) * 32 + lane;            // row of this CTA's W half
            const uint32_t row_off = static_cast<uint32_t>(row) * 128;
            const uint32_t sw = static_cast<uint32_t>(row & 7);
            const int kblocks = p.K / 32;                                   // ggml blocks per row
            int stage = 0;
            uint32_t phase = 0;
            WorkIter wi(p, nkb, cluster_id, num_clusters);
            int tile, kb0, kb1;
            while (wi.next(tile, kb0, kb1)) {
                const int n = (tile % p.n_tiles) * BN + static_cast<int>(cta_rank) * (BN / 2) + row;
                const bool valid = n < p.N;
                const uint32_t* wrow = reinterpret_cast<const uint32_t*>(p.wraw + static_cast<size_t>(valid ? n : 0) * p.wrow_bytes);
                uint32_t cur[WORDS], nxt[WORDS];
                auto fetch = [&](int kb, uint32_t (&dst)[WORDS]) {
                    // two blocks = 2 * BLOCK_BYTES bytes, 4-byte aligned because K % 64 == 0; the second block may lie past K
                    const bool two = 2 * kb + 1 < kblocks;
                    const uint32_t* src = wrow + static_cast<size_t>(kb) * (2 * DqTraits<WT>::BLOCK_BYTES / 4);
#pragma unroll
                    for (int i = 0; i < WORDS; ++i) dst[i] = (valid && (two || i < WORDS / 2 + 1)) ? __ldg(src + i) : 0u;
                };
                fetch(kb0, cur);
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (kb + 1 < kb1) fetch(kb + 1, nxt);                    // one k-step ahead, overlaps the wait + decode below
                    uint32_t o[32];
                    decode_row<WT>(cur, o);
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sB = smem + stage * STAGE_BYTES + A_BYTES;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<uint4*>(sB + row_off + ((static_cast<uint32_t>(c) ^ sw) << 4)) =
                            make_uint4(o[4 * c + 0], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                    fence_proxy_async_smem();                                // generic-proxy writes -> visible to the tensor core
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(&full_bar[stage], 0); // the leader's full barrier, from either CTA
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
#pragma unroll
                    for (int i = 0; i < WORDS; ++i) cur[i] = nxt[i];
                }
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9)
        const int ew = warp - 2;
        const int q = warp & 3;                    // TMEM lane quadrant this warp may access
        const int chalf = ew >> 2;                 // which 128-column half of the tile
        uint8_t* bufs = epi_smem + ew * (EPI_BUFS * CHUNK_BYTES);
        constexpr bool F16OUT = (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16);
        constexpr bool REDUCE = (EPI == EPI_BIAS_RESID_F32);   // out += v through a TMA reduction store
        constexpr int CCOLS = F16OUT ? 64 : 32;    // columns per 128-byte chunk row
        constexpr int CHUNKS = (BN / 2) / CCOLS;   // chunks per warp per tile
        const uint32_t row_off = static_cast<uint32_t>(lane) * 128;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);

        long gc = 0;                               // running chunk counter: buffer rotation
        int acc = 0;
        uint32_t acc_phase = 0;
#ifdef Q2W_GEMM_WALL
        unsigned long long w_acc = 0;
#endif
        WorkIter wi(p, nkb, cluster_id, num_clusters);
        int tile, kb0, kb1;
        while (wi.next(tile, kb0, kb1)) {
            const int m0 = (tile / p.n_tiles) * (2 * BM) + static_cast<int>(cta_rank) * BM + q * 32;
            const int nbase = (tile % p.n_tiles) * BN + chalf * (BN / 2);
            const bool add_bias = p.bias != nullptr && kb0 == 0;   // split-K: only the segment that starts the K range carries the bias
            // this warp's 128 bias values, four per lane, requested BEFORE the accumulator wait and handed out by shuffles below: a
            // per-chunk global load put one L2 / HBM round trip per chunk on the critical path of a single-tile launch (measured
            // 4-5 us between "accumulator ready" and "last store issued" at M = 1500)
            float4 breg = make_float4(0.f, 0.f, 0.f, 0.f);
            if (add_bias && nbase + 4 * lane < p.N) breg = __ldg(reinterpret_cast<const float4*>(p.bias + nbase + 4 * lane));
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
#ifdef Q2W_GEMM_WALL
            w_acc = gtime();
#endif
            const uint32_t t_row = tmem_base + acc * BN + chalf * (BN / 2) + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < CHUNKS; ++c, ++gc) {
                const int n = nbase + c * CCOLS;
                const int b = static_cast<int>(gc % EPI_BUFS);
                uint8_t* buf = bufs + b * CHUNK_BYTES;
                float v[CCOLS];
#ifdef Q2W_GEMM_WALL
                const unsigned long long wc0 = gtime();
#endif
                {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_row + c * CCOLS, r);
                    if constexpr (F16OUT) {
                        uint32_t r2[32];
                        tmem_ld_32x32b_x32(t_row + c * CCOLS + 32, r2);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r[j]); v[32 + j] = __uint_as_float(r2[j]); }
                    } else {
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    }
                }
                if (c == CHUNKS - 1) {
                    // every TMEM read of this warp for this tile is done: hand the accumulator back to the leader's MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster_cta_release(&tempty_bar[acc], 0);   // TMEM reads are ordered by the tcgen05 fence, not by a memory fence
                }
                // ---- bias (+ scale / GELU / positional) in registers; the column index is uniform across the warp
                if (add_bias) {                                    // warp-uniform
#pragma unroll
                    for (int g = 0; g < CCOLS / 4; ++g) {
                        const int src = c * (CCOLS / 4) + g;           // lane that holds columns [n + 4g, n + 4g + 4)
                        v[4 * g + 0] += __shfl_sync(0xffffffffu, breg.x, src);
                        v[4 * g + 1] += __shfl_sync(0xffffffffu, breg.y, src);
                        v[4 * g + 2] += __shfl_sync(0xffffffffu, breg.z, src);
                        v[4 * g + 3] += __shfl_sync(0xffffffffu, breg.w, src);
                    }
                }
                if constexpr (EPI == EPI_BIAS_F16) {
#pragma unroll
                    for (int j = 0; j < CCOLS; ++j) v[j] *= (n + j < p.scale_cols) ? p.scale : 1.0f;
                }
                if constexpr (EPI == EPI_BIAS_GELU_F16 || EPI == EPI_BIAS_GELU_POS_F32) {
#pragma unroll
                    for (int j = 0; j < CCOLS; ++j) v[j] = gelu_tanh(v[j]);
                }
                if constexpr (EPI == EPI_BIAS_GELU_POS_F32) {
                    const int m = m0 + lane;
                    if (m < p.M) {
                        const float* pr = p.pos + static_cast<size_t>(m % p.pos_period) * p.N + n;
#pragma unroll
                        for (int g = 0; g < CCOLS / 4; ++g) {
                            if (n + 4 * g < p.N) {
                                const float4 pv = __ldg(reinterpret_cast<const float4*>(pr + 4 * g));
                                v[4 * g + 0] += pv.x; v[4 * g + 1] += pv.y; v[4 * g + 2] += pv.z; v[4 * g + 3] += pv.w;
                            }
                        }
                    }
                }
                // ---- the chunk buffer: buffer b was last read by the store of chunk gc - EPI_BUFS; at most EPI_BUFS - 1 newer groups may
                //      still be pending
#ifdef Q2W_GEMM_WALL
                const unsigned long long wc1 = gtime();
#endif
                if (lane == 0) bulk_wait_group_read<EPI_BUFS - 1>();
                __syncwarp();
#ifdef Q2W_GEMM_WALL
                const unsigned long long wc2 = gtime();
#endif
                // ---- write the 128-byte row of this chunk, 16-byte pieces XOR-swizzled like SWIZZLE_128B expects
                if constexpr (F16OUT) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        uint4 pk;
                        __half2 h0 = __floats2half2_rn(v[8 * g + 0], v[8 * g + 1]);
                        __half2 h1 = __floats2half2_rn(v[8 * g + 2], v[8 * g + 3]);
                        __half2 h2 = __floats2half2_rn(v[8 * g + 4], v[8 * g + 5]);
                        __half2 h3 = __floats2half2_rn(v[8 * g + 6], v[8 * g + 7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(buf + row_off + ((static_cast<uint32_t>(g) ^ sw) << 4)) = pk;
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        *reinterpret_cast<float4*>(buf + row_off + ((static_cast<uint32_t>(g) ^ sw) << 4)) =
                            make_float4(v[4 * g + 0], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                }
#ifdef Q2W_GEMM_WALL
                const unsigned long long wc3 = gtime();
#endif
                fence_proxy_async_smem();
                __syncwarp();
#ifdef Q2W_GEMM_WALL
                const unsigned long long wc4 = gtime();
#endif
                if (lane == 0) {
                    if (n < p.N) {                                  // rows >= M / columns >= N are clipped by the TMA unit
                        if constexpr (REDUCE) tma_reduce_add_2d(&tmO, buf, n, m0);
                        else tma_store_2d(&tmO, buf, n, m0);
                    }
                    bulk_commit_group();
                }
#ifdef Q2W_GEMM_WALL
                if (leader && ew == 0 && lane == 0 && cluster_id == 0)
                    printf("GC N%d K%d chunk %d: ld+math %llu wait_buf %llu st.shared %llu fence %llu issue %llu ns\n", p.N, p.K, c, wc1 - wc0, wc2 - wc1, wc3 - wc2, wc4 - wc3, gtime() - wc4);
#endif
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        // shared memory must outlive the reads of the last stores; their global writes are ordinary outstanding memory operations of this
        // grid, complete and visible before any dependent grid passes its griddepcontrol.wait / starts in stream order
#ifdef Q2W_GEMM_WALL
        const unsigned long long w_issued = gtime();
#endif
        if (lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
#ifdef Q2W_GEMM_WALL
        if (leader && ew == 0 && lane == 0 && (cluster_id == 0 || cluster_id == num_clusters - 1))
            printf("GW epi c%d/%d N%d K%d acc_ready %llu stores_issued %llu stores_read %llu\n", cluster_id, num_clusters, p.N, p.K, w_acc, w_issued, gtime());
#endif
    }

    tc_fence_before();
    cluster_sync_exec_only();    // no CTA leaves (or frees TMEM) while its peer may still signal it; nothing to publish: relaxed arrive
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_cta2(tmem_base, TMEM_COLS);
    }
#ifdef Q2W_GEMM_WALL
    if (leader && threadIdx.x == 0 && (cluster_id == 0 || cluster_id == num_clusters - 1))
        printf("GW end c%d/%d N%d K%d exit %llu\n", cluster_id, num_clusters, p.N, p.K, gtime());
#endif
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

// 2-D row-major [rows, cols] with leading dimension ld (elements); box = box_rows x box_cols (128 bytes wide), 128B swizzle
bool make_tmap_2d(CUtensorMap* tm, const void* ptr, CUtensorMapDataType dt, size_t esize, uint64_t rows, uint64_t cols, uint64_t ld,
                  uint32_t box_rows, uint32_t box_cols) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * esize};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

std::atomic<int> g_launches{0};

// split-K policy for the residual epilogue (Q2W_GEMM_SPLITK: 0 = off, 1 = equal whole k-ranges per tile, 2 = balanced unit ranges that may
// straddle tiles; default 1.  Measured at M = 1500 on B200, graph replay, us per launch: out-proj 13.2 / 10.7 / 13.4, fc2 30.2 / 20.1 / 20.7
// for modes 0 / 1 / 2 -- the balanced ranges pay a second 256 KB reduction epilogue per cluster for 33 instead of 40 k-blocks)
std::atomic<int> g_splitk_override{-1};   // gemm_set_splitk_mode(): tests pin the bit-exact single-pass path with 0
int splitk_mode() {
    static const int env_mode = [] {
        const char* e = std::getenv("Q2W_GEMM_SPLITK");
        return e ? std::atoi(e) : 1;
    }();
    const int o = g_splitk_override.load(std::memory_order_relaxed);
    return o >= 0 ? o : env_mode;
}
constexpr int SK_MIN_KB = 16;      // only K >= 1024 is worth cutting (and the tiny test models keep their bit-exact single-pass sums)
constexpr int SK_MIN_SEG = 4;      // k-blocks per cluster, at least

template <int EPI, int WT>
cudaError_t launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, KParams kp, cudaStream_t st) {
    static std::atomic<unsigned long long> configured{0};
    DeviceInfo di;
    cudaError_t e = current_device_info(di);
    if (e != cudaSuccess) return e;
    if ((e = smem_optin_once(gemm_kernel<EPI, WT>, smem_bytes<EPI>(), di.dev, configured)) != cudaSuccess) return e;
    const int tiles = kp.m_tiles * kp.n_tiles;
    const int max_clusters = di.num_sms / 2;
    int clusters = tiles < max_clusters ? tiles : max_clusters;
    kp.sk_upc = 0;
    if constexpr (EPI == EPI_BIAS_RESID_F32) {
        // Small M leaves most of the machine idle (M = 1500: 30 tiles for 74 CTA pairs, and fc2 runs 80 k-blocks on them). The
        // epilogue already is `out += partial` at L2, so cutting K costs nothing but the order of the F32 adds: with two or more
        // partial sums per element the result is no longer bit-reproducible from run to run (differences of one F32 ulp of the
        // residual stream). That trade is confined to this small-M path: large batches never take it.
        const int nkb = (kp.K + BK - 1) / BK;
        const int mode = splitk_mode();
        if (mode > 0 && nkb >= SK_MIN_KB && 2 * tiles <= max_clusters) {
            int upc;
            if (mode == 1) {
                int split = max_clusters / tiles;
                while (split > 1 && (nkb % split || nkb / split < SK_MIN_SEG)) --split;
                upc = nkb / split;
            } else {
                const int units = tiles * nkb;
                upc = (units + max_clusters - 1) / max_clusters;
                if (upc < SK_MIN_SEG) upc = SK_MIN_SEG;
            }
            if (upc < nkb) {
                kp.sk_upc = upc;
                clusters = (tiles * nkb + upc - 1) / upc;
            }
        }
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return launch_pdl(gemm_kernel<EPI, WT>, dim3(2 * clusters), dim3(WT == WT_F16 ? NUM_THREADS : NUM_THREADS_Q), smem_bytes<EPI>(), st, tmA, tmB,
                      tmO, kp);
}

}  // namespace

int gemm_num_launches() { return g_launches.load(); }
void gemm_set_splitk_mode(int mode) { g_splitk_override.store(mode, std::memory_order_relaxed); }

cudaError_t gemm_f16_tcgen05(const GemmArgs& a, GemmEpilogue epi, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0 || a.K <= 0) return cudaErrorInvalidValue;
    if ((a.K % 8) || (a.N % 8) || (a.lda % 8) || (a.ldw % 8) || (a.ldo % 8)) return cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.W) | reinterpret_cast<uintptr_t>(a.out)) & 15)
        return cudaErrorMisalignedAddress;
    const bool f16out = (epi == EPI_BIAS_F16 || epi == EPI_BIAS_GELU_F16);
    CUtensorMap tmA, tmB, tmO;
    if (!make_tmap_2d(&tmA, a.A, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a.M, a.K, a.lda, BM, BK)) return cudaErrorInvalidValue;
    const int wt = a.wtype == 0 ? WT_F16 : a.wtype;
    if (wt != WT_F16 && wt != WT_Q8_0 && wt != WT_Q4_0) return cudaErrorInvalidValue;
    if (wt == WT_F16) {
        if (!make_tmap_2d(&tmB, a.W, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a.N, a.K, a.ldw, BN / 2, BK)) return cudaErrorInvalidValue;
    } else {
        if (a.K % 64) return cudaErrorInvalidValue;     // two whole ggml blocks per k-step keep the raw reads 4-byte aligned
        tmB = tmA;                                      // unused
    }
    if (f16out) {
        if (!make_tmap_2d(&tmO, a.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a.M, a.N, a.ldo, 32, 64)) return cudaErrorInvalidValue;
    } else {
        if (!make_tmap_2d(&tmO, a.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.M, a.N, a.ldo, 32, 32)) return cudaErrorInvalidValue;
    }
    if (epi == EPI_BIAS_RESID_F32 && a.resid != a.out) return cudaErrorInvalidValue;   // the residual epilogue accumulates into `out`: it must BE the residual
    KParams kp;
    kp.M = a.M; kp.N = a.N; kp.K = a.K;
    kp.bias = a.bias;
    kp.pos = a.pos; kp.pos_period = a.pos_period > 0 ? a.pos_period : 1;
    kp.scale_cols = a.scale_cols; kp.scale = a.scale;
    kp.m_tiles = (a.M + 2 * BM - 1) / (2 * BM);
    kp.n_tiles = (a.N + BN - 1) / BN;
    kp.wraw = static_cast<const uint8_t*>(static_cast<const void*>(a.W));
    kp.wrow_bytes = wt == WT_Q8_0 ? a.K / 32 * 34 : wt == WT_Q4_0 ? a.K / 32 * 18 : 0;
    kp.sk_upc = 0;
    kp.w_static = a.w_static;
    if (epi == EPI_BIAS_GELU_POS_F32 && !a.pos) return cudaErrorInvalidValue;
#define Q2W_DISPATCH(E)                                                                 \
    case E:                                                                             \
        if (wt == WT_F16) return launch<E, WT_F16>(tmA, tmB, tmO, kp, st);         \
        if (wt == WT_Q8_0) return launch<E, WT_Q8_0>(tmA, tmB, tmO, kp, st);       \
        return launch<E, WT_Q4_0>(tmA, tmB, tmO, kp, st);
    switch (epi) {
        Q2W_DISPATCH(EPI_BIAS_F16)
        Q2W_DISPATCH(EPI_BIAS_GELU_F16)
        Q2W_DISPATCH(EPI_BIAS_RESID_F32)
        Q2W_DISPATCH(EPI_BIAS_F32)
        case EPI_BIAS_GELU_POS_F32:     // conv stem only: its kernels are always F16 in the model file (vtype, :1543)
            if (wt != WT_F16) return cudaErrorInvalidValue;
            return launch<EPI_BIAS_GELU_POS_F32, WT_F16>(tmA, tmB, tmO, kp, st);
    }
#undef Q2W_DISPATCH
    return cudaErrorInvalidValue;
}

}  // namespace q2w
