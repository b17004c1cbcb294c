// attention_tcgen05.cu -- fused non-causal self-attention on the 5th-gen tensor cores (sm_100a).
//
// Replaces the reference's per-head  KQ = mul_mat(K, Q) ; soft_max(KQ) ; mul_mat(V, KQ_soft_max)  chain and the
// permute/cont copies around it (/root/reference/src/qwen2-whisper.cpp:2052-2106; the flash_attn branch there is an empty
// stub).  Q arrives pre-multiplied by 1/sqrt(64) (ggml_scale :2054 folded into the QKV GEMM epilogue).
// Softmax follows ggml_soft_max (ggml/src/ggml.c:13854-13940): scale 1, no mask, row max subtracted, normalised by the
// row sum; it is evaluated online in FP32, probabilities are rounded to F16 for the PV product, FP32 accumulation.
//
// Persistent kernel, two CTAs per SM (256 TMEM columns each), 256 threads per CTA. A work item = 128 query rows of one (window, head);
// CTA c starts with item c and pulls the rest from a global counter (both CTAs of an SM finish together); the KV tile sequence is
// continuous across items, so nothing drains at an item boundary.
//   warp 0      TMA producer: Q (two buffers, the next item's Q half an item ahead), K_g / V_g tiles (128 x 64 f16, SWIZZLE_128B, two
//               stages each) through a 3-D tensor map over qkv[B][T][3D] -- rows past T are out of bounds for the map and arrive as
//               zeros, never as the next window
//   warp 1      MMA issuer:  S = Q K_g^T   tcgen05.mma kind::f16  M128 N128 K16 x4   (A, B K-major)          -> TMEM cols [0,128)
//                            O += P_g V_g  tcgen05.mma kind::f16  M128 N64  K16 x8   (A = P read from TMEM cols [192,256),
//                                                                                      B = V MN-major as loaded) -> TMEM cols [128,192)
//               QK_{g+1} is issued before PV_g (also across items), so S_{g+1} is ready when softmax_g ends
//   warps 4-7   softmax, one thread per query row: tcgen05.ld the S row (single pass, 128 registers), mask the ragged last
//               tile, running max / sum in the log2 domain (FMNMX3, FFMA2, FADD2: two elements per issue slot), ex2, pack to F16,
//               tcgen05.st the P row back into TMEM. O stays in TMEM for the whole KV loop; it is rescaled (tcgen05.ld -> mul ->
//               tcgen05.st) only when the running max grew by more than 2^8 since the last rescale -- the stale max is exact algebra,
//               it only bounds the magnitude of P (<= 256 in F16), and the final division by the row sum uses the same reference point.
//               Item epilogue (O / l -> F16) is deferred into the next item's first tile, staged 128B-swizzled in the finished item's
//               own Q buffer and written by ONE TMA store (rows >= T clipped by the TMA unit).
// Registers are rebalanced with setmaxnreg (producer / MMA / rider-decode warpgroup 72, softmax warpgroup 184).
// Measured (tools/ubench/xu_pipe.cu, tools/att_clk.py, make EXTRA_NVFLAGS=-DQ2W_ATT_TIMELINE): MUFU.EX2 issues at 8 clk per warp per
// scheduler and the whole softmax instruction mix fits under it (64.6 clk per 8 elements with two warps per scheduler); the steady
// state is ~2200 clk per tile per CTA against the 2048-clk MUFU floor of two co-resident CTAs; run back to back the kernel sits at the
// 1000 W power cap (SM clock 1.75 GHz of 1.965), so removing idle cycles now buys clocks, not time.
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

constexpr int HD = 64, BQ = 128, BKV = 128;
constexpr int THREADS = 256;
constexpr int TILE_BYTES = 128 * 128;           // 128 rows x 64 f16 = 16 KB
constexpr int SMEM_DATA = 6 * TILE_BYTES;       // Q x2, K x2, V x2 = 96 KB per CTA, two CTAs per SM (P never touches shared memory;
                                                // the finished item's Q buffer doubles as the staging tile of its output store)
constexpr int SMEM_BYTES = SMEM_DATA + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 256;
constexpr int S_COL = 0, O_COL = 128, P_COL = 192;   // S f32 [0,128) | O f32 [128,192) | P f16x2 [192,256)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;       // log2 units

// MN-major operand tile written by TMA with SWIZZLE_128B: each K row is 128 B (64 f16 along N), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(TILE_BYTES >> 4) << 16;   // LBO: stride between 64-element MN atoms (unused: N = 64 is one atom)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;         // SBO: stride between 8-row K groups
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- packed FP32x2 arithmetic (FFMA2 / FADD2: one issue slot per two elements) and the 3-input FMNMX3
__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t add2_rm(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (M = 128 rows in lanes, K packed two f16 per column) comes from TMEM
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Persistent: grid = 2 CTAs per SM; each CTA starts with work item blockIdx.x (window, head, 128-query tile) and pulls every further
// item from a global counter. The KV tile sequence is continuous across items (global tile index g): the producer prefetches the next item's Q (two Q buffers) and its
// first K/V tiles while the current item's last tiles are still in the softmax, and the MMA warp issues S = Q' K_0'^T of the next
// item under the current item's last exp phase -- so TMEM allocation, barrier setup, descriptor fetch and the first TMA round trip
// (a third of a one-item CTA's lifetime, measured) are paid once per CTA instead of once per item.
__global__ void __launch_bounds__(THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_out, int T, int D, int n_qt, int H, int n_items,
                    int* __restrict__ sched /* [0] next item, [1] finished CTAs; zero on entry, zero again on exit */,
                    const DequantJob job /* weight matrices to decode meanwhile (warps 2-3), all nblocks 0 = nothing */, int job_type) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                         // [2]
    uint8_t* sK = smem + 2 * TILE_BYTES;        // [2]  (a single K / V buffer exposes the TMA round trip every tile: measured ~600 clk
    uint8_t* sV = smem + 4 * TILE_BYTES;        // [2]   of s_full wait per 3000-clk tile, and PV_g queued behind the late QK_{g+1})
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_DATA);
    uint64_t* q_full = bars + 0;    // [2]
    uint64_t* q_empty = bars + 2;   // [2]  last QK of the item that used this Q buffer has completed
    uint64_t* k_full = bars + 4;    // [2]
    uint64_t* k_empty = bars + 6;   // [2]  QK_g complete
    uint64_t* v_full = bars + 8;    // [2]
    uint64_t* v_empty = bars + 10;  // [2]  PV_g complete
    uint64_t* pv_done = bars + 12;  // PV_g complete: P buffer free, O accumulated
    uint64_t* s_full = bars + 13;
    uint64_t* s_empty = bars + 14;  // softmax has pulled S_g into registers
    uint64_t* p_full = bars + 15;
    uint64_t* q_free = bars + 16;   // [2]  the output store staged in this Q buffer has been read out
    uint64_t* id_full = bars + 18;  // [2]  item_ring[it & 3] holds the id of this CTA's it-th item (or -1: no more work)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    volatile int* item_ring = reinterpret_cast<volatile int*>(tmem_slot + 1);   // [4]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kv = (T + BKV - 1) / BKV;
#ifdef Q2W_ATT_TIMELINE   // diagnostic build (make EXTRA_NVFLAGS=-DQ2W_ATT_TIMELINE): per-tile clock stamps of one softmax thread, printed at exit
    __shared__ long long tl[40][6];
    const long long t_start = clock64();
    const bool tl_on = blockIdx.x == 5 && warp == 4 && lane == 0;
#define TL(e) if (tl_on && g < 40) tl[g][e] = clock64() - t_start
#else
#define TL(e)
#endif
    int items_done = 0;   // (softmax warps) how many items this CTA ended up processing

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm);
        tma_prefetch_desc(&tm_out);
        for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&q_free[i], 1); mbar_init(&id_full[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        mbar_init(pv_done, 1);
        mbar_init(s_full, 1);
        mbar_init(s_empty, 128);
        mbar_init(p_full, 128);
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
    pdl_wait();                  // the QKV GEMM's output is visible from here on; the prologue above overlapped its tail
    pdl_launch_dependents();

    // Work is PULLED: the producer thread takes the next item from a global counter and publishes it to the other roles through
    // item_ring / id_full. With a static split the two CTAs of an SM drift apart (whichever wins issue priority finishes up to 35 %
    // earlier, measured) and the loser runs its tail alone, where one softmax warp per scheduler cannot keep the MUFU fed.
    // item -> (query tile, head, window); query tiles of one (window, head) are adjacent items so that co-running CTAs share K/V in L2
    auto item_id = [&](int it) {   // id of this CTA's it-th item, -1 when there is none
        mbar_wait(&id_full[it & 1], (it >> 1) & 1);
        return item_ring[it & 3];
    };
    auto item_coords = [&](int item, int& q0, int& h, int& b) {
        const int qt = item % n_qt;
        const int r = item / n_qt;
        h = r % H;
        b = r / H;
        q0 = qt * BQ;
    };

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        if (warp == 0 && lane == 0) {
            // ------------------------------------------------------------ TMA producer
            // take the next item, wait until the Q buffer (and id slot) of item it - 2 is really retired, publish, start the Q load
            auto next_item = [&](int it) {
                // every CTA's first item is its own index (no round trip before the first loads); the counter hands out the rest
                // (and when the grid already covers every item there is nothing to hand out: no round trip at all)
                const int id = it == 0 ? static_cast<int>(blockIdx.x)
                               : static_cast<int>(gridDim.x) >= n_items ? n_items : static_cast<int>(gridDim.x) + atomicAdd(sched, 1);
                const int qb = it & 1;
                if (it >= 2) {
                    mbar_wait(&q_empty[qb], ((it >> 1) - 1) & 1);   // QKs of item it - 2 done with the buffer
                    mbar_wait(&q_free[qb], ((it >> 1) - 1) & 1);    // ... and its output tile, staged there, stored
                }
                item_ring[it & 3] = id < n_items ? id : -1;
                mbar_arrive(&id_full[qb]);                          // release: the slot is written
                if (id >= n_items) return -1;
                int q0, h, b;
                item_coords(id, q0, h, b);
                mbar_expect_tx(&q_full[qb], TILE_BYTES);
                tma_load_3d(sQ + qb * TILE_BYTES, &tm, &q_full[qb], h * HD, q0, b);
                return id;
            };
            int g = 0;
            int cur = next_item(0);
            int nxt = -1;
            for (int it = 0; cur >= 0; ++it) {
                int q0, h, b;
                item_coords(cur, q0, h, b);
                for (int j = 0; j < n_kv; ++j, ++g) {
                    const int st = g & 1;
                    if (g >= 2) mbar_wait(&k_empty[st], ((g >> 1) - 1) & 1);
                    mbar_expect_tx(&k_full[st], TILE_BYTES);
                    tma_load_3d(sK + st * TILE_BYTES, &tm, &k_full[st], D + h * HD, j * BKV, b);
                    if (g >= 2) mbar_wait(&v_empty[st], ((g >> 1) - 1) & 1);
                    mbar_expect_tx(&v_full[st], TILE_BYTES);
                    tma_load_3d(sV + st * TILE_BYTES, &tm, &v_full[st], 2 * D + h * HD, j * BKV, b);
                    // the item after this one: its Q goes into the buffer of item it - 1, free once that item's output (staged there
                    // during this item's first tile) has been stored -- half an item ahead of its first use
                    // (the second item right after the first tile's loads: both Q buffers are still unused)
                    if (j == (it == 0 ? 0 : n_kv / 2)) nxt = next_item(it + 1);
                }
                cur = nxt;
            }
        } else if (warp == 1 && lane == 0) {
            // ------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc_qk = make_idesc_f16(BQ, BKV, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_f16(BQ, HD, 0, 1);     // B = V is MN-major
            // software pipeline: S_{g+1} = Q K_{g+1}^T is issued BEFORE waiting for P_g, so it runs under softmax_g's exp phase
            // (the single S buffer is free as soon as softmax_g has pulled its row into registers: s_empty) -- also across items.
            auto issue_qk = [&](int it, int j, int g) {
                const int qb = it & 1;
                if (j == 0) mbar_wait(&q_full[qb], (it >> 1) & 1);
                mbar_wait(&k_full[g & 1], (g >> 1) & 1);
                if (g > 0) mbar_wait(s_empty, (g - 1) & 1);
                tc_fence_after();
                const uint64_t q_desc = make_sw128_kmajor_desc(smem_u32(sQ + qb * TILE_BYTES));
                const uint64_t k_desc = make_sw128_kmajor_desc(smem_u32(sK + (g & 1) * TILE_BYTES));
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_f16_ss(tmem_base + S_COL, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k != 0);
                umma_commit(s_full);
                umma_commit(&k_empty[g & 1]);
                if (j == n_kv - 1) umma_commit(&q_empty[qb]);
            };
            int g = 0;
            bool more = item_id(0) >= 0;
            if (more) issue_qk(0, 0, 0);
            for (int it = 0; more; ++it) {
                for (int j = 0; j < n_kv; ++j, ++g) {
                    if (j + 1 < n_kv) {
                        issue_qk(it, j + 1, g + 1);
                    } else {
                        more = item_id(it + 1) >= 0;
                        if (more) issue_qk(it + 1, 0, g + 1);
                    }
                    mbar_wait(p_full, g & 1);
                    mbar_wait(&v_full[g & 1], (g >> 1) & 1);
                    tc_fence_after();
                    const uint64_t v_desc = make_sw128_mnmajor_desc(smem_u32(sV + (g & 1) * TILE_BYTES));
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k) {
                        // P is the A operand straight from TMEM: lane = query row, 16 f16 (one K step) = 8 packed 32-bit columns
                        // V: 16 K rows (kv) per step = 2 KB
                        const uint64_t vb = v_desc + static_cast<uint64_t>(k * (2048 >> 4));
                        umma_f16_ts(tmem_base + O_COL, tmem_base + P_COL + k * 8, vb, idesc_pv, (j | k) != 0);
                    }
                    umma_commit(pv_done);
                    umma_commit(&v_empty[g & 1]);
                }
            }
        } else if (warp >= 2 && job_type != 0) {
            // ------------------------------------------------------------ rider: ggml block decode (warps 2 and 3 are otherwise idle)
            // Quantised weights stay Q8_0 / Q4_0 in HBM; the F16 copy the next GEMMs read is produced HERE, under the attention kernel,
            // because nothing else can run next to the encoder's kernels at small batch: a GEMM CTA owns its SM's shared memory and TMEM,
            // two attention CTAs own the register file -- a decode kernel on a second stream only runs in the gaps and delays their CTAs
            // (measured: single-window p50 4.2 ms against 3.45 ms for an F16 file). These 64 threads per CTA have 72 registers, no
            // shared memory and nothing to do; they walk the job in pairs of ggml blocks (68 / 36 bytes, 4-byte aligned), one pair per
            // thread per step, values bit-identical to dequantize_row_* + one F16 rounding (decode_row, dequant.cuh).
            const unsigned long long tid = static_cast<unsigned long long>(blockIdx.x) * 64 + (threadIdx.x - 64);
            const unsigned long long nthreads = static_cast<unsigned long long>(gridDim.x) * 64;
            // (constant indices only: a runtime index into the by-value job would put it on the local-memory stack)
            const unsigned long long n0 = job.nblocks[0] >> 1, n1 = job.nblocks[1] >> 1, n2 = job.nblocks[2] >> 1, n3 = job.nblocks[3] >> 1;
            const unsigned long long pairs_total = n0 + n1 + n2 + n3;
            // software-pipelined: the words of pair i + 1 are requested before pair i is decoded (these warps are latency-bound)
            auto locate = [&](unsigned long long pi, const uint8_t*& sb, __half*& db, unsigned long long& pr) {
                pr = pi; sb = job.src[0]; db = job.dst[0];
                if (pr >= n0) {
                    pr -= n0; sb = job.src[1]; db = job.dst[1];
                    if (pr >= n1) {
                        pr -= n1; sb = job.src[2]; db = job.dst[2];
                        if (pr >= n2) { pr -= n2; sb = job.src[3]; db = job.dst[3]; }
                    }
                }
            };
            if (job_type == WT_Q8_0) {
                uint32_t cur[17], nxt[17];
                const uint8_t* sb; __half* db; unsigned long long pr;
                unsigned long long pi = tid;
                if (pi < pairs_total) {
                    locate(pi, sb, db, pr);
                    const uint32_t* src = reinterpret_cast<const uint32_t*>(sb + pr * 68);
#pragma unroll
                    for (int i = 0; i < 17; ++i) cur[i] = __ldg(src + i);
                }
                for (; pi < pairs_total; pi += nthreads) {
                    __half* dst = db + pr * 64;
                    const unsigned long long pn = pi + nthreads;
                    if (pn < pairs_total) {
                        locate(pn, sb, db, pr);
                        const uint32_t* src = reinterpret_cast<const uint32_t*>(sb + pr * 68);
#pragma unroll
                        for (int i = 0; i < 17; ++i) nxt[i] = __ldg(src + i);
                    }
                    uint32_t o[32];
                    decode_row<WT_Q8_0>(cur, o);
#pragma unroll
                    for (int c = 0; c < 8; ++c) reinterpret_cast<uint4*>(dst)[c] = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
#pragma unroll
                    for (int i = 0; i < 17; ++i) cur[i] = nxt[i];
                }
            } else {
                uint32_t cur[9], nxt[9];
                const uint8_t* sb; __half* db; unsigned long long pr;
                unsigned long long pi = tid;
                if (pi < pairs_total) {
                    locate(pi, sb, db, pr);
                    const uint32_t* src = reinterpret_cast<const uint32_t*>(sb + pr * 36);
#pragma unroll
                    for (int i = 0; i < 9; ++i) cur[i] = __ldg(src + i);
                }
                for (; pi < pairs_total; pi += nthreads) {
                    __half* dst = db + pr * 64;
                    const unsigned long long pn = pi + nthreads;
                    if (pn < pairs_total) {
                        locate(pn, sb, db, pr);
                        const uint32_t* src = reinterpret_cast<const uint32_t*>(sb + pr * 36);
#pragma unroll
                        for (int i = 0; i < 9; ++i) nxt[i] = __ldg(src + i);
                    }
                    uint32_t o[32];
                    decode_row<WT_Q4_0>(cur, o);
#pragma unroll
                    for (int c = 0; c < 8; ++c) reinterpret_cast<uint4*>(dst)[c] = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
#pragma unroll
                    for (int i = 0; i < 9; ++i) cur[i] = nxt[i];
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
        // ------------------------------------------------------------ softmax (thread = query row)
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
        // item epilogue: O / l -> f16 -> the item's own (now idle) Q buffer, 128-byte swizzled -> ONE TMA store of the 128 x 64 tile.
        // Per-thread global stores of a 128-byte row cost 32 L1 line transactions per instruction (1024 per item, measured ~1500 clk);
        // the TMA store writes whole lines and clips rows >= T itself. The producer reloads the buffer only after q_free.
        const bool epi_leader = warp == 4 && lane == 0;
        auto item_epilogue = [&](float inv, int q0, int h, int b, int qb) {
            uint8_t* stage = sQ + qb * TILE_BYTES;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t o[32];
                tmem_ld_32x32b_x32(t_lane + O_COL + hh * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 4; ++c) {                  // SWIZZLE_128B: 16-byte chunk index ^= row & 7
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        __half2 v = __floats2half2_rn(__uint_as_float(o[8 * c + 2 * i]) * inv, __uint_as_float(o[8 * c + 2 * i + 1]) * inv);
                        w[i] = *reinterpret_cast<uint32_t*>(&v);
                    }
                    *reinterpret_cast<uint4*>(stage + row * 128 + (((hh * 4 + c) ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (epi_leader) {
                tma_store_3d(&tm_out, stage, h * HD, q0, b);
                bulk_commit_group();
            }
        };
        int g = 0;
        float prev_inv = 0.f;
        int prev_q0 = 0, prev_h = 0, prev_b = 0;
        const int release_tile = n_kv > 1 ? 1 : 0;   // the tile after which the leader hands the staged Q buffer back to the producer
        for (int it = 0;; ++it) {
            const int item = item_id(it);
            if (item < 0) break;
            items_done = it + 1;
            int q0, h, b;
            item_coords(item, q0, h, b);
            float m_used = -INFINITY;   // reference point of P and O, log2 domain
            float l_sum = 0.f;
            for (int j = 0; j < n_kv; ++j, ++g) {
                TL(0);
                mbar_wait(s_full, g & 1);
                tc_fence_after();
                TL(1);
                uint32_t s[4][32];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) tmem_ld_32x32b_x32(t_lane + S_COL + c4 * 32, s[c4]);
                tmem_ld_wait();
                TL(2);
                tc_fence_before();
                mbar_arrive(s_empty);
                // ---- row max (log2 domain), ragged last tile masked
                const int valid = T - j * BKV;   // >= 1
                float mx = -INFINITY;
                if (valid < BKV) {               // warp-uniform
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c4 * 32 + i >= valid) s[c4][i] = __float_as_uint(-INFINITY);
                }
                {
                    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
                        for (int i = 0; i < 32; i += 2)   // FMNMX3: two elements per issue slot, 4 independent chains
                            m4[c4] = max3(m4[c4], __uint_as_float(s[c4][i]), __uint_as_float(s[c4][i + 1]));
                    mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                }
                mx *= LOG2E;   // log2 domain; p = ex2(s * log2e - m) is one FFMA + one MUFU per element
                // ---- lazy rescale decision (warp-uniform because tcgen05.ld/st are warp-collective)
                float factor = 1.0f;
                bool need = false;
                if (j == 0) {
                    m_used = mx;
                } else if (mx - m_used > RESCALE_THRESHOLD) {
                    need = true;
                    factor = ex2(m_used - mx);
                    m_used = mx;
                    l_sum *= factor;
                }
                const bool any_need = __any_sync(0xffffffffu, need);
                // ---- probabilities (packed FFMA2 / FADD2: one issue slot per two elements)
                uint32_t pk[4][16];
                uint64_t rs2[4] = {0ull, 0ull, 0ull, 0ull};   // packed (even, odd) partial row sums
                const uint64_t l2e2 = pack2(LOG2E, LOG2E), nm2 = pack2(-m_used, -m_used);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const uint64_t x = fma2(pack2(__uint_as_float(s[c4][i]), __uint_as_float(s[c4][i + 1])), l2e2, nm2);
                        float x0, x1;
                        unpack2(x, x0, x1);
                        const float p0 = ex2(x0), p1 = ex2(x1);
                        rs2[c4] = add2(rs2[c4], pack2(p0, p1));
                        __half2 hh = __floats2half2_rn(p0, p1);
                        pk[c4][i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
                    }
                float rs4[4];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    float a, bb;
                    unpack2(rs2[c4], a, bb);
                    rs4[c4] = a + bb;
                }
                l_sum += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
                // ---- P buffer free, O_{g-1} accumulated
                if (g > 0) {
                    mbar_wait(pv_done, (g - 1) & 1);
                    tc_fence_after();
                    if (j == 0) {
                        // the previous item's O is complete and this item's first PV (accumulate = 0 into the same columns) is only
                        // issued after the p_full arrival below: its epilogue runs here, a whole exp phase after its last PV was
                        // issued, instead of waiting for that PV at the item boundary (measured 1300-1700 clk of idle MUFU)
                        item_epilogue(prev_inv, prev_q0, prev_h, prev_b, (it - 1) & 1);
                        tc_fence_before();
                    } else if (any_need) {
                        uint32_t o[32];
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            tmem_ld_32x32b_x32(t_lane + O_COL + hh * 32, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                            tmem_st_32x32b_x32(t_lane + O_COL + hh * 32, o);
                        }
                        tmem_st_wait();
                    }
                }
                // ---- P row -> TMEM columns [192,256) (two f16 per 32-bit column): no shared memory, no proxy fence
                tmem_st_32x32b_x32(t_lane + P_COL, *reinterpret_cast<uint32_t (*)[32]>(&pk[0][0]));
                tmem_st_32x32b_x32(t_lane + P_COL + 32, *reinterpret_cast<uint32_t (*)[32]>(&pk[2][0]));
                TL(3);
                tmem_st_wait();
                TL(4);
                tc_fence_before();
                mbar_arrive(p_full);
                TL(5);
                if (it > 0 && j == release_tile && epi_leader) {
                    bulk_wait_group_read<0>();                 // issued a tile ago: normally long complete
                    mbar_arrive(&q_free[(it - 1) & 1]);
                }
                __syncwarp();
            }
            prev_inv = 1.0f / l_sum;
            prev_q0 = q0;
            prev_h = h;
            prev_b = b;
        }
        if (items_done > 0) {   // the last item's epilogue
            mbar_wait(pv_done, (g - 1) & 1);
            tc_fence_after();
            item_epilogue(prev_inv, prev_q0, prev_h, prev_b, (items_done - 1) & 1);
            tc_fence_before();
        }
        if (epi_leader) bulk_wait_group_read<0>();   // shared memory must outlive the last store's reads; the writes complete with the grid
    }

#ifdef Q2W_ATT_TIMELINE
    if (warp == 4 && lane == 0) {
        uint32_t smid;
        unsigned long long gt;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        printf("cta %d sm %u items %d cycles %lld end_ns %llu\n", (int)blockIdx.x, smid, items_done, clock64() - t_start, gt);
    }
    if (tl_on)
        for (int g = 0; g < 40 && g < items_done * n_kv; ++g)
            printf("g%2d top %6lld s_full %6lld (+%4lld) ld %6lld (+%4lld) st_issued %6lld (+%5lld) st_done %6lld (+%4lld) arrive +%lld\n", g, tl[g][0], tl[g][1],
                   tl[g][1] - tl[g][0], tl[g][2], tl[g][2] - tl[g][1], tl[g][3], tl[g][3] - tl[g][2], tl[g][4], tl[g][4] - tl[g][3], tl[g][5] - tl[g][4]);
#endif
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
    // (when the grid covers every item the work counter is never touched: nothing to count, no atomic round trip in the kernel's tail)
    if (static_cast<int>(gridDim.x) < n_items && threadIdx.x == 0 &&
        atomicAdd(sched + 1, 1) == static_cast<int>(gridDim.x) - 1) {   // last CTA out: every CTA has taken its final id
        sched[0] = 0;
        sched[1] = 0;
        __threadfence();
    }
}

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
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(f);
    });
    return fn;
}

}  // namespace

cudaError_t attention_f16_tcgen05(const __half* qkv, __half* out, int B, int T, int H, int* sched, cudaStream_t st, const DequantJob* job,
                                  int job_type) {
    if (!sched) return cudaErrorInvalidValue;
    if (B <= 0 || T <= 0 || H <= 0) return cudaErrorInvalidValue;
    const int D = H * HD;
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return cudaErrorInvalidValue;
    // qkv viewed as [B][T][3D] f16: rows >= T of a window are out of bounds for the map (zero fill), never the next window
    CUtensorMap tm;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(3 * D), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(3 * D) * sizeof(__half), static_cast<cuuint64_t>(T) * 3 * D * sizeof(__half)};
    cuuint32_t box[3] = {HD, BKV, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(qkv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    // out viewed as [B][T][D] f16, boxes of 128 rows x 64 columns (one head): rows >= T are clipped by the TMA unit
    CUtensorMap tm_out;
    {
        cuuint64_t odims[3] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
        cuuint64_t ostrides[2] = {static_cast<cuuint64_t>(D) * sizeof(__half), static_cast<cuuint64_t>(T) * D * sizeof(__half)};
        cuuint32_t obox[3] = {HD, BQ, 1};
        if (enc(&tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, out, odims, ostrides, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    static std::atomic<unsigned long long> configured{0};   // per device (launch.cuh)
    DeviceInfo di;
    cudaError_t err = current_device_info(di);
    if (err != cudaSuccess) return err;
    if ((err = smem_optin_once(attention_tc_kernel, SMEM_BYTES, di.dev, configured)) != cudaSuccess) return err;
    const int num_sms = di.num_sms;
    const int n_qt = (T + BQ - 1) / BQ;
    const long long items = static_cast<long long>(n_qt) * H * B;
    if (items > 0x7fffffffLL / ((T + BKV - 1) / BKV)) return cudaErrorInvalidValue;
    const int n_items = static_cast<int>(items);
    const int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;   // persistent: two CTAs per SM
    DequantJob j{};
    int jt = 0;
    if (job != nullptr && (job_type == WT_Q8_0 || job_type == WT_Q4_0)) {
        j = *job;
        jt = job_type;
        for (int i = 0; i < 4; ++i)
            if ((j.nblocks[i] & 1) || (reinterpret_cast<uintptr_t>(j.src[i]) & 3) || (reinterpret_cast<uintptr_t>(j.dst[i]) & 15)) return cudaErrorInvalidValue;
    }
    return launch_pdl(attention_tc_kernel, dim3(grid), dim3(THREADS), SMEM_BYTES, st, tm, tm_out, T, D, n_qt, H, n_items, sched, j, jt);
}

}  // namespace q2w
