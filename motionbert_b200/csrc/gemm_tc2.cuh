// 2-CTA (cta_group::2) tcgen05 GEMM with a TMA-fed / TMA-drained epilogue  --  the production GEMM.
//
//   D[M,N] = A[M,K] * W[N,K]^T, operands as bf16 hi/lo planes (see gemm_tc.cuh for the arithmetic modes and
//   the fused epilogues; this kernel shares GemmParams / EPI_* with it).
//
// Why a second kernel: ncu on the 1-CTA version (profiles/r01_*) showed (a) L1TEX saturated (65-70 %) by the
// epilogue's row-strided 16-byte global accesses, which also starves the TMA->smem operand feed, and (b) 48 KB
// of operands per 768 MMA-cycles per SM.  Here:
//   * a CTA PAIR (cluster 2x1, same TPC) computes a 256 x 256 tile with UMMA M=256: each CTA stages only its own
//     128 A rows and HALF of the B tile (32 KB / stage instead of 48 KB), the leader CTA's single thread issues the
//     MMAs for both SMs, completion is multicast to both CTAs' barriers (tcgen05.commit ... multicast::cluster);
//   * the epilogue never touches global memory with LSU instructions for tile data: the residual tile arrives by
//     TMA into swizzled smem (one 32x32 fp32 box per warp, prefetched one chunk ahead), results are written to
//     swizzled smem staging and leave through TMA stores (cp.async.bulk.tensor ... bulk_group), fully coalesced,
//     rows beyond M clipped by the tensor map.
// Warp roles per CTA (320 threads): w0 TMA producer, w1 MMA issuer (leader CTA only) + TMEM alloc,
// w2..w9 epilogue (lane quadrant = warp%4, column half = (warp-2)/4, 4 chunks of 32 columns each).
#pragma once
#include "gemm_tc.cuh"

namespace mb {

constexpr int G2_STAGES_RESID = 4;   // residual epilogue needs 12 KB staging per warp
constexpr int G2_STAGES_OTHER = 5;   // split-only / fp32-only epilogues need 8 KB -> one more operand stage
constexpr int G2_THREADS = 320;        // default: 8 epilogue warps
constexpr int G2_EPI_WARPS = 8;
// Single-pass (bf16) GEMMs have a 3x shorter mainloop per tile and ncu showed their epilogue to be LATENCY-bound
// (tensor pipe 32-48 % active, issue slots 24-46 %): the bf16-plane epilogues can therefore run with 16 epilogue
// warps (EW = 16: each warp owns 32 rows x 64 columns, 4 warps per scheduler hide tcgen05.ld / MUFU / TMA-store
// latency; no register double-buffering at the 112-register cap of 576 threads).
constexpr int g2_threads(int ew) { return (2 + ew) * 32; }

// MB_F16C_DEEP (A/B build switch): the F16C qkv / fc1 / tail GEMMs trade the second epilogue staging buffer of every warp
// for a sixth operand stage (a stage lasts 512 MMA-cycles in F16C against 768 in BF16x3: less look-ahead per stage).
#ifndef MB_F16C_DEEP
#define MB_F16C_DEEP 0
#endif
template <int PASSES, int EPI, int EW = G2_EPI_WARPS>
struct Gemm2Cfg {
    static constexpr bool DEEP = MB_F16C_DEEP && PASSES == 2 && EPI != EPI_RESID && EW == 8;
    static constexpr int STAGES = DEEP ? 6 : (EPI == EPI_RESID) ? G2_STAGES_RESID : G2_STAGES_OTHER;
    // 8 warps: buf0 4 KB | buf1 4 KB | [split 4 KB].  16 warps (bf16-plane outputs only): one 2 KB plane tile per
    // buffer, double-buffered (4 KB), or -- two output planes -- a single 4 KB buffer
    static constexpr bool TWO_PLANES = (PASSES == 3) || (EPI == EPI_BIAS_GELU_PAIR);
    static constexpr int NBUF = ((EW == 16 && TWO_PLANES) || DEEP) ? 1 : 2;
    static constexpr int BUF_BYTES = (EW == 16 && !TWO_PLANES) ? 2048 : 4096;
    static constexpr int STAGING_PER_WARP = (EPI == EPI_RESID) ? 12288 : NBUF * BUF_BYTES;
    // PASSES == 2 is the F16C mode (ptx.cuh): one SWIZZLE_128B row of [32 f16 | 32 lo8 | 32 hi8] per 32-element K block,
    // 2 fp16 MMAs + 2 e5m2 MMAs per block = 2 pass-equivalents of the 16-bit tensor rate.
    static constexpr bool F16C = (PASSES == 2);
    static constexpr int BK = (PASSES == 1) ? 64 : 32;
    static constexpr int SWZ = F16C ? 128 : BK * 2;
    static constexpr int KSTEP16 = SWZ / 2;                    // tensor-map (16-bit unit) coordinate step per stage
    static constexpr uint32_t LAYOUT = (SWZ == 128) ? 2u : 4u;
    static constexpr int PLANES = (PASSES == 3) ? 2 : 1;
    static constexpr int A_PLANE = 128 * SWZ;                  // this CTA's 128 rows of A
    static constexpr int B_PLANE = 128 * SWZ;                  // this CTA's half (128 of 256 rows) of the W tile
    static constexpr int A_BYTES = PLANES * A_PLANE;
    static constexpr int B_BYTES = PLANES * B_PLANE;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;      // 32 KB
    static constexpr int OFF_STAGING = STAGES * STAGE_BYTES;
    static constexpr int OFF_BAR = OFF_STAGING + EW * STAGING_PER_WARP;
    static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
};

// B_MN = true is the data-gradient form  dX[M, Kf] = G[M, Nf] * W[Nf, Kf]  (backward groundwork, row a15): the
// contraction runs over W's ROW index, so the very same packed W planes are consumed as an MN-major B operand
// (64-column SWIZZLE_128B blocks, LBO = block stride) instead of packing a transposed copy of every weight.
// OUT16C: the split output (tmS) is an F16C row buffer (2-D map, box (64 x 16-bit, 32 rows), SWIZZLE_128B) instead of
// bf16 hi/lo planes.
template <int PASSES, int EPI, bool B_MN = false, int EW = G2_EPI_WARPS, bool OUT16C = (PASSES == 2)>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(g2_threads(EW), 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA,   // bf16 3D (K, M, plane), box (BK, 128, PLANES)
             const __grid_constant__ CUtensorMap tmB,   // bf16 3D (K, N, plane), box (BK, 128, PLANES); B_MN: (Kf, Nf, plane), box (64, BK, 1)
             const __grid_constant__ CUtensorMap tmR,   // fp32 2D (N, M) residual,    box (32, 32)         [RESID]
             const __grid_constant__ CUtensorMap tmX,   // fp32 2D (N, M) output,      box (32, 32)         [RESID/F32]
             const __grid_constant__ CUtensorMap tmS,   // bf16 3D (N, M, plane) out,  box (32, 32, PLANES) [RESID/SPLIT]
             const GemmParams p) {
    using Cfg = Gemm2Cfg<PASSES, EPI, EW>;
    constexpr int G2_STAGES = Cfg::STAGES;
    static_assert(EW == 8 || (EW == 16 && PASSES == 1 && (EPI == EPI_LN_SPLIT || EPI == EPI_LN_GELU_SPLIT || EPI == EPI_BIAS_SPLIT ||
                                                          EPI == EPI_BIAS_GELU_PAIR || EPI == EPI_GELUBWD_SPLIT)),
                  "16 epilogue warps: single-pass bf16-plane epilogues only");
    constexpr bool kResid = (EPI == EPI_RESID);
    constexpr bool kF32Out = (EPI == EPI_RESID || EPI == EPI_LN_TANH_F32 || EPI == EPI_BIAS_F32);
    constexpr bool kSplitOut = (EPI == EPI_RESID || EPI == EPI_LN_SPLIT || EPI == EPI_LN_GELU_SPLIT || EPI == EPI_BIAS_SPLIT ||
                                EPI == EPI_BIAS_GELU_PAIR || EPI == EPI_GELUBWD_SPLIT);
    constexpr bool kTwoPlanes = (PASSES == 3 || (PASSES == 2 && !OUT16C)) || (EPI == EPI_BIAS_GELU_PAIR);   // second bf16 plane: lo, or gelu(y)
    static_assert(!(B_MN && PASSES == 2), "the F16C mode has no MN-major weight form (backward runs in bf16)");
    static_assert(!OUT16C || EW == 8, "F16C output: 8 epilogue warps");
    constexpr bool kDoubleLd = !kResid && EW == 8;                               // register double-buffered tcgen05.ld
    constexpr bool kLn = (EPI == EPI_LN_SPLIT || EPI == EPI_LN_GELU_SPLIT || EPI == EPI_LN_TANH_F32 || EPI == EPI_LN_TANH_POOL);
    constexpr bool kPool = (EPI == EPI_LN_TANH_POOL);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* full_bar = bars;                             // [STAGES] leader's is the one in use
    uint64_t* empty_bar = bars + G2_STAGES;                // [STAGES] per CTA, multicast-committed by the leader
    uint64_t* tfull_bar = bars + 2 * G2_STAGES;            // [2]      per CTA, multicast-committed by the leader
    uint64_t* tempty_bar = bars + 2 * G2_STAGES + 2;       // [2]      leader's: all epilogue threads of the pair
    uint64_t* rbar = bars + 2 * G2_STAGES + 4;             // [8 warps][2] residual-tile landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G2_STAGES + 4 + 2 * EW);

    const int warp = warp_uniform(threadIdx.x >> 5);     // role dispatch on a value nvcc knows is warp-uniform
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int npairs = gridDim.x >> 1;

    const int num_mp = (p.M + 255) / 256;
    const int num_n = p.N / 256;
    const int num_tiles = num_mp * num_n;
    const int num_kb = p.K / Cfg::BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (kResid) tma_prefetch_desc(&tmR);
        if (kF32Out) tma_prefetch_desc(&tmX);
        if (kSplitOut) tma_prefetch_desc(&tmS);
        for (int i = 0; i < G2_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 2 * EW);             // one arrival per epilogue warp of the pair
        }
        for (int i = 0; i < 2 * EW; ++i) mbar_init(&rbar[i], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2cta<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        // every lane walks the (warp-uniform) loop and one elected lane issues: nvcc keeps coordinates, shared addresses
        // and descriptors in uniform registers (with `if (lane == 0)` around the loop every UTMALDG / UTCHMMA was wrapped
        // in an ELECT + R2UR waterfall of ~10 instructions)
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += npairs) {
                const int m_pair = tile / num_n, n_idx = tile % num_n;
                const int a_row = m_pair * 256 + static_cast<int>(rank) * 128;
                const int b_row = n_idx * 256 + static_cast<int>(rank) * 128;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one()) {
                        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
                        uint8_t* sB = sA + Cfg::A_BYTES;
                        const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        tma_load_3d_2cta(sA, &tmA, full_leader, kb * Cfg::KSTEP16, a_row, 0);
                        if (!B_MN) {
                            tma_load_3d_2cta(sB, &tmB, full_leader, kb * Cfg::KSTEP16, b_row, 0);
                        } else {
                            // this CTA's 128 output columns = two 64-column blocks of [BK contraction rows][128 B] per plane
                            for (int pl = 0; pl < Cfg::PLANES; ++pl)
                                for (int blk = 0; blk < 2; ++blk)
                                    tma_load_3d_2cta(sB + pl * Cfg::B_PLANE + blk * (Cfg::BK * 128), &tmB, full_leader,
                                                     b_row + blk * 64, kb * Cfg::BK, pl);
                        }
                    }
                    __syncwarp();
                    if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t IDESC = Cfg::F16C ? umma_idesc_fmt(256, 256, 0, 0, 0, 0)      // f16 x f16
                                                 : umma_idesc_bf16(256, 256, 0, B_MN ? 1 : 0);
            constexpr uint32_t IDESC8 = umma_idesc_fmt(256, 256, 1, 1, 0, 0);                  // e5m2 x e5m2
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < num_tiles; tile += npairs) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);      // no memory is handed over: TMEM reads are ordered by the tcgen05 fences
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                        const uint32_t sB = sA + Cfg::A_BYTES;
                        const uint64_t a_hi = umma_smem_desc(sA, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                        const uint64_t a_lo = umma_smem_desc(sA + Cfg::A_PLANE, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                        // B: K-major rows of SWZ bytes, or (B_MN) MN-major 64-column blocks of [BK rows][128 B], SWIZZLE_128B
                        const uint64_t b_hi = B_MN ? umma_smem_desc(sB, Cfg::BK * 128, 1024, 2u)
                                                   : umma_smem_desc(sB, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                        const uint64_t b_lo = B_MN ? umma_smem_desc(sB + Cfg::B_PLANE, Cfg::BK * 128, 1024, 2u)
                                                   : umma_smem_desc(sB + Cfg::B_PLANE, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                        if (Cfg::F16C) {
                            // K-slices of the 128-byte block: f16 at +0 / +32 B, lo8 at +64 B, hi8 at +96 B (units of 16 B)
                            umma_ss_2cta(d_tmem, a_hi, b_hi, IDESC, kb != 0);
                            umma_ss_2cta(d_tmem, a_hi + 2, b_hi + 2, IDESC, 1);
                            umma_ss_2cta_f8(d_tmem, a_hi + 4, b_hi + 6, IDESC8, 1);     // (al 2^6) * (wh 2^-6)
                            umma_ss_2cta_f8(d_tmem, a_hi + 6, b_hi + 4, IDESC8, 1);     // (ah 2^-6) * (wl 2^6)
                        }
#pragma unroll
                        for (int ks = 0; ks < (Cfg::F16C ? 0 : Cfg::BK / 16); ++ks) {
                            const uint64_t koff = static_cast<uint64_t>((ks * 32) >> 4);
                            const uint64_t boff = B_MN ? static_cast<uint64_t>((ks * 16 * 128) >> 4) : koff;   // 16 rows down
                            if (PASSES == 3) {
                                umma_ss_2cta(d_tmem, a_lo + koff, b_hi + boff, IDESC, (kb | ks) != 0);
                                umma_ss_2cta(d_tmem, a_hi + koff, b_lo + boff, IDESC, 1);
                                umma_ss_2cta(d_tmem, a_hi + koff, b_hi + boff, IDESC, 1);
                            } else {
                                umma_ss_2cta(d_tmem, a_hi + koff, b_hi + boff, IDESC, (kb | ks) != 0);
                            }
                        }
                        tc_commit_2cta(&empty_bar[stage], 3);
                        if (kb == num_kb - 1) tc_commit_2cta(&tfull_bar[acc], 3);
                    }
                    __syncwarp();
                    if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
        const int ew = warp - 2;
        const int quad = warp & 3;
        const int half = ew >> 2;                      // column group of this warp: 128 columns (EW = 8) / 64 (EW = 16)
        constexpr int NCH = 32 / EW;                   // 32-column chunks per warp per tile: 4 / 2
        constexpr int COLS_PER_WARP = NCH * 32;
        uint8_t* stg = smem + Cfg::OFF_STAGING + ew * Cfg::STAGING_PER_WARP;
        uint8_t* buf[2] = {stg, stg + (Cfg::NBUF == 2 ? Cfg::BUF_BYTES : 0)};
        uint8_t* bufS = stg + 8192;   // only exists (and is only used) for EPI_RESID
        uint64_t* my_rbar = rbar + 2 * ew;
        const int ngrp_out = p.N / STATS_GROUP;
        const uint32_t sw128 = static_cast<uint32_t>(lane & 7);          // SWIZZLE_128B: chunk16 ^= row % 8
        const uint32_t sw64 = static_cast<uint32_t>((lane >> 1) & 3);    // SWIZZLE_64B : chunk16 ^= (row / 2) % 4

        auto chunk_coords = [&](int tile, int ch, int& col0, int& rowb) {
            const int m_pair = tile / num_n, n_idx = tile % num_n;
            col0 = n_idx * 256 + half * COLS_PER_WARP + ch * 32;
            rowb = m_pair * 256 + static_cast<int>(rank) * 128 + quad * 32;
        };
        uint32_t ci = 0;   // chunks processed by this warp (buffer parity / rbar phase)
        if (kResid && pair < num_tiles && elect_one()) {
            int c0, r0;
            chunk_coords(pair, 0, c0, r0);
            mbar_arrive_expect_tx(&my_rbar[0], 4096);
            tma_load_2d(buf[0], &tmR, &my_rbar[0], c0, r0);
        }
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = pair; tile < num_tiles; tile += npairs) {
            const int m_pair = tile / num_n, n_idx = tile % num_n;
            const int row = m_pair * 256 + static_cast<int>(rank) * 128 + quad * 32 + lane;
            const bool row_ok = row < p.M;

            float mean = 0.f, rstd = 1.f, rscale = 1.f;
            if (kLn) {
                if (row_ok) ln_row_stats(p.stats_in + static_cast<size_t>(row) * p.nh_in * 3, p.nh_in, p.ln_dim,
                                         p.eps, mean, rstd);
            }
            if (kResid) {
                if (row_ok && p.row_scale) rscale = p.row_scale[row / p.J];
            }
            float st_shift = 0.f;
            float2 st_sum2 = make_float2(0.f, 0.f), st_sq2 = make_float2(0.f, 0.f);

            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * 256 + half * COLS_PER_WARP + (static_cast<uint32_t>(quad * 32) << 16);
#ifdef MB_EXP_NO_EPI
            // timing experiment (A/B builds only, results are garbage): the accumulator is released unread
            if (Cfg::F16C && !kResid) {
                tc_fence_before();
                __syncwarp();
                if (elect_one()) mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
#endif
            uint32_t racc[kDoubleLd ? 2 : 1][32];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch, ++ci) {
                const int b = ci & 1;
                int col0, rowb;
                chunk_coords(tile, ch, col0, rowb);
                if (kResid) {
                    mbar_wait(&my_rbar[b], (ci >> 1) & 1);            // residual chunk landed in buf[b]
                    if (elect_one()) {
                        tma_store_wait_read<0>();                      // group ci-1 no longer reads buf[b^1] / bufS
                        int nt = tile, nc = ch + 1;
                        if (nc == NCH) { nc = 0; nt = tile + npairs; }
                        if (nt < num_tiles) {
                            int c1, r1;
                            chunk_coords(nt, nc, c1, r1);
                            mbar_arrive_expect_tx(&my_rbar[b ^ 1], 4096);
                            tma_load_2d(buf[b ^ 1], &tmR, &my_rbar[b ^ 1], c1, r1);
                        }
                    }
                } else if (Cfg::NBUF == 2) {
                    if (elect_one()) tma_store_wait_read<1>();           // group ci-2 no longer reads buf[b]
                } else {
                    if (elect_one()) tma_store_wait_read<0>();           // single staging buffer: previous store has read it
                }
                // non-residual epilogues double-buffer the accumulator chunk: tcgen05.ld of chunk ch+1 is in flight while
                // chunk ch is processed (the residual variant has no registers to spare at 10 warps / 168 registers)
                uint32_t (&r)[32] = racc[kDoubleLd ? (ch & 1) : 0];
                if (!kDoubleLd) {
                    tmem_ld32(t_row + ch * 32, r);
                    tmem_ld_wait();
                } else {
                    if (ch == 0) tmem_ld32(t_row, racc[0]);
                    tmem_ld_wait();
                    if (ch + 1 < NCH) tmem_ld32(t_row + (ch + 1) * 32, racc[(ch + 1) & 1]);
                }
                if (ch == NCH - 1) {
                    // every TMEM read of this accumulator is complete (the last chunk sits in registers) -> hand it back to
                    // the leader's MMA warp BEFORE the last chunk is processed and stored: each lane fences its
                    // tcgen05.ld's, the warp converges, ONE lane arrives (a 32-way release-arrive cost 11 % of the
                    // epilogue's issue slots in fences, profiles/r02a)
                    tc_fence_before();
                    __syncwarp();
                    if (elect_one()) mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
                }
                float v[32];
                if (kResid) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.vec0 + col0);
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
                } else if (EPI == EPI_BIAS_F32 || EPI == EPI_BIAS_SPLIT || EPI == EPI_BIAS_GELU_PAIR) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.vec0 + col0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 bb = __ldg(b4 + i);
                        v[4 * i + 0] = __uint_as_float(r[4 * i + 0]) + bb.x;
                        v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bb.y;
                        v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bb.z;
                        v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bb.w;
                    }
                } else if (EPI == EPI_GELUBWD_SPLIT) {
                    // d h_pre = d h * gelu'(h_pre): my row's 32 pre-activations (64 contiguous bytes) straight from global
                    uint4 a[4];
                    if (row_ok) {
                        const uint4* a4 = reinterpret_cast<const uint4*>(p.aux + static_cast<size_t>(row) * p.N + col0);
#pragma unroll
                        for (int i = 0; i < 4; ++i) a[i] = __ldg(a4 + i);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) a[i] = make_uint4(0u, 0u, 0u, 0u);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t w[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            v[8 * i + 2 * e] = __uint_as_float(r[8 * i + 2 * e]) * gelu_grad_fast(__uint_as_float(w[e] << 16));
                            v[8 * i + 2 * e + 1] = __uint_as_float(r[8 * i + 2 * e + 1]) * gelu_grad_fast(__uint_as_float(w[e] & 0xffff0000u));
                        }
                    }
                } else {
                    const float4* c4 = reinterpret_cast<const float4*>(p.vec0 + col0);
                    const float4* s4 = reinterpret_cast<const float4*>(p.vec1 + col0);
                    const float ms = -mean * rstd;
                    const float2 ms2 = make_float2(ms, ms), rstd2 = make_float2(rstd, rstd);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 c = __ldg(c4 + i);
                        const float4 s = __ldg(s4 + i);
                        // packed fp32x2 FMAs (bit-identical to two scalar fmaf each; half the issue slots)
                        float2 a = __ffma2_rn(rstd2, make_float2(__uint_as_float(r[4 * i + 0]), __uint_as_float(r[4 * i + 1])),
                                              __ffma2_rn(ms2, make_float2(s.x, s.y), make_float2(c.x, c.y)));
                        float2 b = __ffma2_rn(rstd2, make_float2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])),
                                              __ffma2_rn(ms2, make_float2(s.z, s.w), make_float2(c.z, c.w)));
                        if (EPI == EPI_LN_GELU_SPLIT) {
                            a = gelu_erf2(a);
                            b = gelu_erf2(b);
                        }
                        v[4 * i + 0] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = b.x; v[4 * i + 3] = b.y;
                    }
                    if (EPI == EPI_LN_TANH_F32 || EPI == EPI_LN_TANH_POOL) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = tanhf(v[i]);
                    }
                    if (kPool && row_ok) {
                        // temporal mean pool of the representation: rows (b, f, j) -> (b, j); fp32 vector reductions into L2
                        const int bj = (row / (p.pool_F * p.J)) * p.J + row % p.J;
                        float* dst = p.out_f32 + static_cast<size_t>(bj) * p.N + col0;
                        const float w = 1.0f / static_cast<float>(p.pool_F);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * i), "f"(v[4 * i] * w),
                                         "f"(v[4 * i + 1] * w), "f"(v[4 * i + 2] * w), "f"(v[4 * i + 3] * w)
                                         : "memory");
                    }
                }
                // all lanes have consumed buf[b] (residual) and lane 0 has seen the older store groups retire
                __syncwarp();
                uint8_t* xs = buf[b];                                   // fp32 staging (aliases the residual tile)
                uint8_t* ss = kResid ? bufS : buf[b];                   // split staging
                const uint32_t xs_row = smem_u32(xs) + lane * 128;     // this lane's 128-byte staging rows
                const uint32_t ss_row = smem_u32(ss) + lane * 128;
                const uint32_t ss_row64 = smem_u32(ss) + lane * 64;
                if (kF32Out) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        sts_v4f(xs_row + ((i ^ sw128) << 4), v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
                const bool do_split = kSplitOut && (!kResid || p.out_hi != nullptr);   // block-final residuals feed
                if (do_split && OUT16C) {
                    // one F16C block per row: 16-byte units 0..3 = 32 f16, 4..5 = 32 lo8, 6..7 = 32 hi8 (SWIZZLE_128B staging)
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
                } else if (do_split) {                                                 // only the fp32 fusion kernel
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (EPI == EPI_BIAS_GELU_PAIR) {
                            const __nv_bfloat162 y2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                            const float2 gl = gelu_erf2(make_float2(v[2 * i], v[2 * i + 1]));
                            const __nv_bfloat162 g2 = __floats2bfloat162_rn(gl.x, gl.y);
                            hi[i] = *reinterpret_cast<const uint32_t*>(&y2);
                            lo[i] = *reinterpret_cast<const uint32_t*>(&g2);
                        } else {
                            split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        sts_v4(ss_row64 + ((i ^ sw64) << 4), hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                        if (kTwoPlanes)
                            sts_v4(ss_row64 + 2048 + ((i ^ sw64) << 4), lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {       // elect.sync is deterministic: the same lane owns every bulk-store group of the warp
                    if (kF32Out) tma_store_2d(&tmX, xs, col0, rowb);
                    if (do_split) {
                        if (OUT16C) tma_store_2d(&tmS, ss, col0 * 2, rowb);
                        else tma_store_3d(&tmS, ss, col0, rowb, 0);
                    }
                    tma_store_commit();
                }
            }
            if (kResid) {
                if (row_ok && p.stats_out) {
                    float* so = p.stats_out + (static_cast<size_t>(row) * ngrp_out + n_idx * 2 + half) * 3;
                    so[0] = st_shift;
                    so[1] = st_sum2.x + st_sum2.y;
                    so[2] = st_sq2.x + st_sq2.y;
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
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
