// One kernel per residual MLP sublayer (F16C arithmetic):   x' = x + rowscale * (fc2(gelu(fc1(LN(x)))) + b2)
//
//   reference: lib/model/DSTformer.py:79-85 (MLP.forward: fc1 -> GELU -> fc2), :242,244,247,249 (x + drop_path(mlp(norm(x))))
//
// The split form (gemm2_kernel<2, EPI_LN_GELU_SPLIT> then gemm2_kernel<2, EPI_RESID>) writes the hidden activation
// (M x hidden x 4 B as F16C rows = 4.33 GB at BASELINE config 2) to HBM and reads it back: 8.66 GB of the 17.8 GB the two
// launches move.  A single 128-row tile cannot hold the fc2 accumulator (C fp32 columns) AND a hidden accumulator in
// the 512 TMEM columns at C = 512, so the fusion here is a PANEL schedule inside one persistent launch instead:
//
//   * a CTA pair owns 256-row token blocks and walks their   fc1 tiles 0 .. NT1-1   (256 hidden columns each, K = C)
//     and   fc2 tiles 0 .. NT2-1   (256 output columns each, K = hidden)   through the SAME TMA ring / UMMA issue path /
//     two TMEM accumulators as gemm2_kernel: the mainloop never drains between the two linears;
//   * the fc1 epilogue (LayerNorm fold, GELU, F16C encode) stores its hidden chunk by TMA; the producer warp of each CTA
//     reloads ITS OWN 128 hidden rows as the fc2 A operand.  The only dependency is intra-CTA: hready[set][n] (an
//     mbarrier, one arrival per epilogue warp) is signalled once every store of fc1 tile n has completed
//     (cp.async.bulk.wait_group: both sides of the hand-over are async-proxy accesses of L2) and gates the 8 K blocks of
//     fc2 that read those 256 hidden columns;
//   * tile order: block by block, fc1 0..NT1-1 then fc2 0..NT2-1.  The hidden rows live in a per-pair slot of a small
//     ring (pairs x 256 rows x hidden x 4 B = 77 MB for 74 pairs) that is rewritten every token block, with an L2 policy
//     per access class: ring stores / reloads L2::evict_last, everything touched once (fp32 residual in, both outputs,
//     the last pass over the x rows) L2::evict_first.  Measured at BASELINE config 2 (ncu, profiles/r02l_mlp_dram_*.csv):
//     15.9 GB of DRAM traffic per launch against 17.4 GB without the policies and 18.0 GB for the two GEMM launches
//     -- the reloads hit L2 more often (read hit rate 69 -> 72-75 %), but ~300 MB of streaming traffic pass through
//     the 126 MB L2 per token-block period and the ring's WRITES still miss (write hit rate < 5 %): the ring is a
//     partial win, not an HBM-free hand-over.  fc2 tile 0 starts on hidden columns 0..767 while the epilogue of the
//     last fc1 tile still runs; only its last quarter waits for it.  (Measured and dropped, profiles/README.md round 2:
//     interleaving the fc2 tiles of block k-1 with the fc1 tiles of block k -- the same number of SM cycles, 5 % slower
//     at the board's power cap.)
//     WAR on a ring slot is excluded by the pipeline itself: the first hidden store of token block k+1 follows the tfull
//     commit of its fc1 tile 0, which follows (in-order tensor pipe) every MMA -- hence every operand load -- of token
//     block k's fc2 tiles.
//
// The fc2 epilogue is gemm2_kernel's EPI_RESID (residual tile in by TMA, fp32 + F16C rows + LN statistics out).
// Arithmetic is identical to the split form instruction for instruction (same MMA order per K block, same epilogue
// math), so the two forms agree BIT FOR BIT: tests/test_gpu_mlp_fused.py.
#pragma once
#include "gemm_tc2.cuh"

namespace mb {

struct MlpParams {
    int M, C, H;               // tokens, model width (fc1 K = fc2 N), hidden width (fc1 N = fc2 K)
    const float* c1;           // fc1: c[n] = sum_k beta_k W1[n,k] + b1[n]          [H]
    const float* s1;           // fc1: s[n] = sum_k (W1 gamma)[n,k]                 [H]
    const float* b2;           // fc2 bias                                          [C]
    const float* stats_in;     // LN partial statistics of x: [M][nh_in][3]
    int nh_in;
    float ln_dim, eps;
    const float* row_scale;    // optional DropPath scale per frame (row / J)
    int J;
    float* stats_out;          // LN partial statistics of x' [M][C/128][3], or null (block-final sublayer)
    int split_out;             // also emit x' as F16C rows (tmS)
    int ring;                  // hidden rows indexed by CTA pair (L2-resident ring) instead of by token block
    int l2_hint;               // hidden stores / loads carry L2::evict_last
};

constexpr int MLPF_STAGES = 4;
constexpr int MLPF_THREADS = 320;
constexpr int MLPF_MAX_NT1 = 8;                 // hidden <= 2048
struct MlpFusedCfg {
    static constexpr int STAGE_BYTES = 2 * 128 * 128;           // A: 128 rows x 128 B, B: 128 rows x 128 B
    static constexpr int A_BYTES = 128 * 128;
    static constexpr int STAGING_PER_WARP = 12288;              // buf0 | buf1 | bufS, 4 KB each
    static constexpr int OFF_STAGING = MLPF_STAGES * STAGE_BYTES;
    static constexpr int OFF_BAR = OFF_STAGING + 8 * STAGING_PER_WARP;
    static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
};

constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;      // createpolicy.fractional.L2::evict_last, fraction 1.0
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;     // createpolicy.fractional.L2::evict_first, fraction 1.0

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_done() {     // <= N groups may still be in flight (writes included)
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* m, const void* smem_src, int c0, int c1, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_2cta_hint(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr,
                                                      int c0, int c1, int c2, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MLPF_THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmA1,   // x as F16C rows        3D (2C, M, 1)  box (64, 128, 1)
                 const __grid_constant__ CUtensorMap tmB1,   // W1' F16C rows [H][C]  3D            box (64, 128, 1)
                 const __grid_constant__ CUtensorMap tmA2,   // hidden F16C rows      3D (2H, M, 1)  box (64, 128, 1)
                 const __grid_constant__ CUtensorMap tmB2,   // W2 F16C rows [C][H]   3D            box (64, 128, 1)
                 const __grid_constant__ CUtensorMap tmH,    // hidden store map      2D (2H, M)     box (64, 32)
                 const __grid_constant__ CUtensorMap tmR,    // fp32 x (residual)     2D (C, M)      box (32, 32)
                 const __grid_constant__ CUtensorMap tmX,    // fp32 x' out           2D (C, M)      box (32, 32)
                 const __grid_constant__ CUtensorMap tmS,    // x' F16C rows out      2D (2C, M)     box (64, 32)
                 const MlpParams p) {
    using Cfg = MlpFusedCfg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* full_bar = bars;                                  // [STAGES] leader's is the one in use
    uint64_t* empty_bar = bars + MLPF_STAGES;                   // [STAGES] per CTA, multicast-committed by the leader
    uint64_t* tfull_bar = bars + 2 * MLPF_STAGES;               // [2]
    uint64_t* tempty_bar = bars + 2 * MLPF_STAGES + 2;          // [2]      leader's: all epilogue warps of the pair
    uint64_t* rbar = bars + 2 * MLPF_STAGES + 4;                // [8 warps][2] residual chunk landed
    uint64_t* hready = bars + 2 * MLPF_STAGES + 4 + 16;         // [MLPF_MAX_NT1] hidden columns of fc1 tile n are in L2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hready + MLPF_MAX_NT1);

    const int warp = warp_uniform(threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int npairs = gridDim.x >> 1;

    const int num_mp = (p.M + 255) / 256;
    const int NT1 = p.H / 256, NT2 = p.C / 256;
    const int KB1 = p.C / 32, KB2 = p.H / 32;
    const int rounds = pair < num_mp ? (num_mp - pair + npairs - 1) / npairs : 0;   // token blocks of this pair
    // L2 policy per access class (see the header): ring lines evict_last, touched-once traffic evict_first
    const uint64_t hpol = p.l2_hint ? L2_EVICT_LAST : L2_EVICT_NORMAL;
    const uint64_t spol = p.l2_hint ? L2_EVICT_FIRST : L2_EVICT_NORMAL;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2);
        tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmR); tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmS);
        for (int i = 0; i < MLPF_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 16);                       // one arrival per epilogue warp of the pair
        }
        for (int i = 0; i < 16; ++i) mbar_init(&rbar[i], 1);
        for (int i = 0; i < MLPF_MAX_NT1; ++i) mbar_init(&hready[i], 8);   // one arrival per epilogue warp of THIS CTA
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2cta<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // hidden rows of this pair's k-th token block (global block index mblk)
    auto hidden_row0 = [&](int mblk) { return (p.ring ? pair : mblk) * 256 + static_cast<int>(rank) * 128; };

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        int stage = 0;
        uint32_t phase = 0;
        for (int k = 0; k < rounds; ++k) {                         // k: which of this pair's token blocks
            for (int e = 0; e < NT1 + NT2; ++e) {
                const bool is_fc2 = e >= NT1;
                const int n = is_fc2 ? e - NT1 : e;
                const int mblk = pair + k * npairs;
                const int a_row = is_fc2 ? hidden_row0(mblk) : mblk * 256 + static_cast<int>(rank) * 128;
                const int b_row = n * 256 + static_cast<int>(rank) * 128;
                const CUtensorMap* ma = is_fc2 ? &tmA2 : &tmA1;
                const CUtensorMap* mw = is_fc2 ? &tmB2 : &tmB1;
                const int nkb = is_fc2 ? KB2 : KB1;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (is_fc2 && n == 0 && (kb & 7) == 0) {
                        // hidden columns [32 kb, 32 kb + 256) = fc1 tile kb/8 of my own 128 rows: stores complete?
                        mbar_wait(&hready[kb >> 3], k & 1);
                        fence_proxy_async_global();
                    }
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one()) {
                        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
                        uint8_t* sB = sA + Cfg::A_BYTES;
                        const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        tma_load_3d_2cta_hint(sA, ma, full_leader, kb * 64, a_row, 0,
                                              is_fc2 ? hpol : (n == NT1 - 1 ? spol : L2_EVICT_NORMAL));
                        tma_load_3d_2cta(sB, mw, full_leader, kb * 64, b_row, 0);
                    }
                    __syncwarp();
                    if (++stage == MLPF_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t IDESC = umma_idesc_fmt(256, 256, 0, 0, 0, 0);        // f16 x f16
            constexpr uint32_t IDESC8 = umma_idesc_fmt(256, 256, 1, 1, 0, 0);       // e5m2 x e5m2
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int k = 0; k < rounds; ++k) {
                for (int e = 0; e < NT1 + NT2; ++e) {
                    const int nkb = e >= NT1 ? KB2 : KB1;
                    mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * 256;
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                            const uint32_t sB = sA + Cfg::A_BYTES;
                            const uint64_t a = umma_smem_desc(sA, 16, 8 * 128, 2u);
                            const uint64_t b = umma_smem_desc(sB, 16, 8 * 128, 2u);
                            // K-slices of the 128-byte block: f16 at +0 / +32 B, lo8 at +64 B, hi8 at +96 B (units of 16 B)
                            umma_ss_2cta(d_tmem, a, b, IDESC, kb != 0);
                            umma_ss_2cta(d_tmem, a + 2, b + 2, IDESC, 1);
                            umma_ss_2cta_f8(d_tmem, a + 4, b + 6, IDESC8, 1);     // (al 2^6) * (wh 2^-6)
                            umma_ss_2cta_f8(d_tmem, a + 6, b + 4, IDESC8, 1);     // (ah 2^-6) * (wl 2^6)
                            tc_commit_2cta(&empty_bar[stage], 3);
                            if (kb == nkb - 1) tc_commit_2cta(&tfull_bar[acc], 3);
                        }
                        __syncwarp();
                        if (++stage == MLPF_STAGES) { stage = 0; phase ^= 1; }
                    }
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
        const int ew = warp - 2;
        const int quad = warp & 3;
        const int half = ew >> 2;                       // 128-column half of the 256-column tile
        uint8_t* stg = smem + Cfg::OFF_STAGING + ew * Cfg::STAGING_PER_WARP;
        uint8_t* buf[2] = {stg, stg + 4096};
        uint8_t* bufS = stg + 8192;
        uint64_t* my_rbar = rbar + 2 * ew;
        const int ngrp_out = p.C / STATS_GROUP;
        const uint32_t sw128 = static_cast<uint32_t>(lane & 7);          // SWIZZLE_128B: chunk16 ^= row % 8
        uint32_t rc = 0;    // residual chunks consumed: buffer parity / rbar phase (4 per fc2 tile: rc & 1 == 0 at every tile start)
        uint32_t hc = 0;    // hidden chunks stored: staging parity (buf[1] / bufS)
        int acc = 0;
        uint32_t acc_phase = 0;
        uint64_t* pend = nullptr;    // hready barrier of the last fc1 tile, not yet signalled (its stores may be in flight)
        int k_stats = -1;            // token block whose LN statistics (mean, rstd) are loaded
        float mean = 0.f, rstd = 1.f;

        // one F16C block per row into a SWIZZLE_128B staging tile: 16-byte units 0..3 = 32 f16, 4..5 = 32 lo8, 6..7 = 32 hi8
        auto stage_f16c = [&](const float (&v)[32], uint32_t ss_row) {
            uint32_t l8[8], g8[8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float xv[8] = {v[8 * g], v[8 * g + 1], v[8 * g + 2], v[8 * g + 3],
                                     v[8 * g + 4], v[8 * g + 5], v[8 * g + 6], v[8 * g + 7]};
                uint32_t h4[4], l2[2], g2[2];
                split8_f16c(xv, h4, l2, g2);
                sts_v4(ss_row + ((g ^ sw128) << 4), h4[0], h4[1], h4[2], h4[3]);
                l8[2 * g] = l2[0]; l8[2 * g + 1] = l2[1];
                g8[2 * g] = g2[0]; g8[2 * g + 1] = g2[1];
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                sts_v4(ss_row + (((4 + u) ^ sw128) << 4), l8[4 * u], l8[4 * u + 1], l8[4 * u + 2], l8[4 * u + 3]);
                sts_v4(ss_row + (((6 + u) ^ sw128) << 4), g8[4 * u], g8[4 * u + 1], g8[4 * u + 2], g8[4 * u + 3]);
            }
        };

        for (int k = 0; k < rounds; ++k) {
            for (int e = 0; e < NT1 + NT2; ++e) {
                const bool is_fc2 = e >= NT1;
                const int n = is_fc2 ? e - NT1 : e;
                const int mblk = pair + k * npairs;
                const int rowb = mblk * 256 + static_cast<int>(rank) * 128 + quad * 32;
                const int row = rowb + lane;
                const bool row_ok = row < p.M;
                const uint32_t t_row = tmem_base + acc * 256 + half * 128 + (static_cast<uint32_t>(quad * 32) << 16);

                if (!is_fc2) {
                    // ---------------- fc1 tile: h = gelu(rstd (acc - mean s) + c) -> F16C rows of the hidden buffer
                    if (k != k_stats) {
                        mean = 0.f; rstd = 1.f;
                        if (row_ok) ln_row_stats(p.stats_in + static_cast<size_t>(row) * p.nh_in * 3, p.nh_in, p.ln_dim, p.eps, mean, rstd);
                        k_stats = k;
                    }
                    const int hrowb = hidden_row0(mblk) + quad * 32;
                    mbar_wait(&tfull_bar[acc], acc_phase);
                    tc_fence_after();
                    uint32_t racc[2][32];
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch, ++hc) {
                        const int col0 = n * 256 + half * 128 + ch * 32;
                        if (elect_one()) {
                            if (ch == 1 && pend) {
                                // every store group of the previous fc1 tile has COMPLETED (only this tile's chunk 0 may be
                                // in flight): its 32 rows x 128 hidden columns are in L2 for the producer's reload
                                tma_store_wait_done<1>();
                                mbar_arrive(pend);
                            } else if (ch == 0) {
                                tma_store_wait_read<0>();                  // the previous tile (possibly an fc2 tile, whose groups
                                                                           // read buf AND bufS) has released all staging
                            } else {
                                tma_store_wait_read<1>();                  // group hc-2 no longer reads this staging buffer
                            }
                        }
                        if (ch == 1) pend = nullptr;
                        uint32_t (&r)[32] = racc[ch & 1];
                        if (ch == 0) tmem_ld32(t_row, racc[0]);
                        tmem_ld_wait();
                        if (ch + 1 < 4) tmem_ld32(t_row + (ch + 1) * 32, racc[(ch + 1) & 1]);
                        if (ch == 3) {
                            tc_fence_before();
                            __syncwarp();
                            if (elect_one()) mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
                        }
                        float v[32];
                        {
                            const float4* c4 = reinterpret_cast<const float4*>(p.c1 + col0);
                            const float4* s4 = reinterpret_cast<const float4*>(p.s1 + col0);
                            const float ms = -mean * rstd;
                            const float2 ms2 = make_float2(ms, ms), rstd2 = make_float2(rstd, rstd);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 c = __ldg(c4 + i);
                                const float4 s = __ldg(s4 + i);
                                float2 a = __ffma2_rn(rstd2, make_float2(__uint_as_float(r[4 * i + 0]), __uint_as_float(r[4 * i + 1])),
                                                      __ffma2_rn(ms2, make_float2(s.x, s.y), make_float2(c.x, c.y)));
                                float2 b = __ffma2_rn(rstd2, make_float2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])),
                                                      __ffma2_rn(ms2, make_float2(s.z, s.w), make_float2(c.z, c.w)));
                                a = gelu_erf2(a);
                                b = gelu_erf2(b);
                                v[4 * i + 0] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = b.x; v[4 * i + 3] = b.y;
                            }
                        }
                        __syncwarp();                                      // the elected lane has seen the older store group retire
                        uint8_t* ss = (hc & 1) ? bufS : buf[(rc & 1) ^ 1];
                        stage_f16c(v, smem_u32(ss) + lane * 128);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (elect_one()) {
                            tma_store_2d_hint(&tmH, ss, col0 * 2, hrowb, hpol);
                            tma_store_commit();
                        }
                    }
                    pend = &hready[n];
                    if (n == NT1 - 1) {
                        // the last hidden columns gate the tail of fc2 tile 0, whose epilogue is the next thing this warp
                        // does: signal now
                        if (elect_one()) {
                            tma_store_wait_done<0>();
                            mbar_arrive(pend);
                        }
                        __syncwarp();
                        pend = nullptr;
                    }
                } else {
                    // ---------------- fc2 tile: x' = x + rowscale (acc + b2); fp32 + F16C rows + LN statistics
                    float rscale = 1.f;
                    if (row_ok && p.row_scale) rscale = p.row_scale[row / p.J];
                    // first residual chunk of this tile -> buf[rc & 1] (lands while this warp waits for the accumulator)
                    if (elect_one()) {
                        tma_store_wait_read<0>();                          // older store groups have read all staging
                        mbar_arrive_expect_tx(&my_rbar[rc & 1], 4096);
                        tma_load_2d_hint(buf[rc & 1], &tmR, &my_rbar[rc & 1], n * 256 + half * 128, rowb, spol);
                    }
                    __syncwarp();
                    float st_shift = 0.f;
                    float2 st_sum2 = make_float2(0.f, 0.f), st_sq2 = make_float2(0.f, 0.f);
                    mbar_wait(&tfull_bar[acc], acc_phase);
                    tc_fence_after();
                    uint32_t r[32];
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch, ++rc) {
                        const int b = rc & 1;
                        const int col0 = n * 256 + half * 128 + ch * 32;
                        mbar_wait(&my_rbar[b], (rc >> 1) & 1);           // residual chunk landed in buf[b]
                        if (ch < 3 && elect_one()) {
                            tma_store_wait_read<0>();                      // older groups no longer read buf[b^1] / bufS
                            mbar_arrive_expect_tx(&my_rbar[b ^ 1], 4096);
                            tma_load_2d_hint(buf[b ^ 1], &tmR, &my_rbar[b ^ 1], col0 + 32, rowb, spol);
                        }
                        tmem_ld32(t_row + ch * 32, r);
                        tmem_ld_wait();
                        if (ch == 3) {
                            tc_fence_before();
                            __syncwarp();
                            if (elect_one()) {
                                mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
                                tma_store_wait_read<0>();                  // (no prefetch this chunk) bufS is free again
                            }
                        }
                        float v[32];
                        {
                            const float4* b4 = reinterpret_cast<const float4*>(p.b2 + col0);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 bb = __ldg(b4 + i);
                                const float4 x = lds_v4f(smem_u32(buf[b]) + lane * 128 + ((i ^ sw128) << 4));
                                v[4 * i + 0] = x.x + rscale * (__uint_as_float(r[4 * i + 0]) + bb.x);
                                v[4 * i + 1] = x.y + rscale * (__uint_as_float(r[4 * i + 1]) + bb.y);
                                v[4 * i + 2] = x.z + rscale * (__uint_as_float(r[4 * i + 2]) + bb.z);
                                v[4 * i + 3] = x.w + rscale * (__uint_as_float(r[4 * i + 3]) + bb.w);
                            }
                            if (ch == 0) st_shift = v[0];
                            {
                                // packed fp32x2 accumulation (two independent chains; combined after the last chunk)
                                const float2 nsh = make_float2(-st_shift, -st_shift);
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    const float2 d = __fadd2_rn(make_float2(v[2 * i], v[2 * i + 1]), nsh);
                                    st_sum2 = __fadd2_rn(st_sum2, d);
                                    st_sq2 = __ffma2_rn(d, d, st_sq2);
                                }
                            }
                        }
                        __syncwarp();                                      // all lanes have consumed buf[b]; older stores retired
                        const uint32_t xs_row = smem_u32(buf[b]) + lane * 128;
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            sts_v4f(xs_row + ((i ^ sw128) << 4), v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        if (p.split_out) stage_f16c(v, smem_u32(bufS) + lane * 128);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (elect_one()) {
                            tma_store_2d_hint(&tmX, buf[b], col0, rowb, spol);
                            if (p.split_out) tma_store_2d_hint(&tmS, bufS, col0 * 2, rowb, spol);
                            tma_store_commit();
                        }
                    }
                    if (row_ok && p.stats_out) {
                        float* so = p.stats_out + (static_cast<size_t>(row) * ngrp_out + n * 2 + half) * 3;
                        so[0] = st_shift;
                        so[1] = st_sum2.x + st_sum2.y;
                        so[2] = st_sq2.x + st_sq2.y;
                    }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        if (elect_one()) tma_store_wait_all();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta<512>(tmem_base);
    }
}

}  // namespace mb
