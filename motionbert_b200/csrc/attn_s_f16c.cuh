// Spatial multi-head attention (and temporal attention of clips with F <= 32) in the F16C arithmetic mode
// (DSTformer.py:178-186 `Attention.forward_spatial`; :188-200 for the packed temporal mode).
//
// Same tiling and pipeline as attn_s_tc.cuh -- four sequences of <= 32 rows packed block-diagonally into one 128-row
// tile, two problems in flight (operand sets in smem, S/P/O sets in TMEM) -- with F16C operands (ptx.cuh): per 32 head
// channels one 128-byte block [32 f16 | 32 lo8 | 32 hi8], S = Q K^T and O = P V as 2 fp16 + 2 e5m2 MMAs per block, the
// probabilities written back over S as [16 cols f16 | 8 cols lo8 | 8 cols hi8] per 32 keys.  Output: F16C rows.
#pragma once
#include "attn_s_tc.cuh"

namespace mb {

struct AttnS16Params {
    int nseq;     // sequences: B*F frames (spatial) or B*J (batch, joint) pairs (temporal-packed)
    int L;        // valid rows per sequence: J (spatial) or F <= 32 (temporal-packed)
    int F, J;
    int C, H;
    float scale_log2e;
    uint8_t* out;            // F16C rows [M][C]
};

template <int HD>
struct AttnS16Cfg {
    static constexpr int NBLK = HD / 32;
    static constexpr int BLK = 128 * 128;                      // one block column of a 128-row tile: 16 KB
    static constexpr int TILE_BYTES = NBLK * BLK;
    static constexpr int SET_BYTES = 3 * TILE_BYTES;           // Q | K | V of one problem
    static constexpr int OFF_BAR = 2 * SET_BYTES;
    static constexpr int OFF_RED = OFF_BAR + 256;
    static constexpr int SMEM_BYTES = OFF_RED + 3 * 128 * 4 + 1024;
    static constexpr int TMEM_SET = 192;                       // S/P 128 columns + O 64 columns per problem
};

// softmax of one row of <= 32 scores (thread = row): probabilities as F16C pieces, returns the row sum.
// LC > 0: row length known at compile time (J = 17 joints: the 15 padding columns cost no exponentials, no compares).
template <int LC>
__device__ __forceinline__ float softmax32_f16c(const uint32_t (&r)[32], int L, float sl2, uint32_t (&hh)[16],
                                                uint32_t (&l8)[8], uint32_t (&g8)[8]) {
    const int n = LC > 0 ? LC : L;
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 32; ++k)
        if (k < n) mx = fmaxf(mx, __uint_as_float(r[k]));
    const float mxs = mx * sl2;
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float pv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            pv[k] = (8 * q + k < n) ? ex2_approx(fmaf(__uint_as_float(r[8 * q + k]), sl2, -mxs)) : 0.f;
            sum += pv[k];
        }
        uint32_t h4[4], l2[2], g2[2];
        split8_f16c(pv, h4, l2, g2);
#pragma unroll
        for (int k = 0; k < 4; ++k) hh[4 * q + k] = h4[k];
        l8[2 * q] = l2[0]; l8[2 * q + 1] = l2[1];
        g8[2 * q] = g2[0]; g8[2 * q + 1] = g2[1];
    }
    return sum;
}

template <int HD, bool TEMPORAL>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_s16_kernel(const __grid_constant__ CUtensorMap tmQKV,   // spatial : 3-D (6C x 16-bit, J, BF), box (64, 32, 4)
                                                             // temporal: 4-D (6C x 16-bit, J, F, B), box (64, 1, 32, 1)
                const AttnS16Params p) {
    using Cfg = AttnS16Cfg<HD>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* qk_full = bars + 0;    // [2]
    uint64_t* qk_empty = bars + 2;   // [2]
    uint64_t* v_full = bars + 4;     // [2]
    uint64_t* v_empty = bars + 6;    // [2]
    uint64_t* s_full = bars + 8;     // [2]
    uint64_t* p_full = bars + 10;    // [2]
    uint64_t* o_full = bars + 12;    // [2]
    uint64_t* o_empty = bars + 14;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = warp_uniform(threadIdx.x >> 5);          // role dispatch on a value nvcc knows is warp-uniform
    const int lane = threadIdx.x & 31;
    const int num_groups = (p.nseq + ATS_FRAMES - 1) / ATS_FRAMES;
    const int num_prob = num_groups * p.H;
    const int n_mine = (num_prob > static_cast<int>(blockIdx.x))
                           ? (num_prob - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qk_full[i], 1);
            mbar_init(&qk_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], ATT_SM_THREADS);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], ATT_SM_THREADS);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        // (every lane walks the warp-uniform loop, one elected lane issues: operands stay in uniform registers)
        {
            for (int i = 0; i < n_mine; ++i) {
                const int prob = blockIdx.x + i * gridDim.x;
                const int h = prob % p.H, g = prob / p.H;
                const int f0 = g * ATS_FRAMES;
                const int s = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                uint8_t* set = smem + s * Cfg::SET_BYTES;
                // one operand tile = [block][4 slabs x 32 rows][128 B]; col16 = 16-bit-unit column of the operand's first block
                auto load_tile = [&](uint8_t* dst, uint64_t* bar, int col16) {
                    for (int blk = 0; blk < Cfg::NBLK; ++blk) {
                        if (!TEMPORAL) {
                            tma_load_3d(dst + blk * Cfg::BLK, &tmQKV, bar, col16 + blk * 64, 0, f0);
                        } else {
                            for (int f = 0; f < ATS_FRAMES; ++f) {
                                const int seq = f0 + f;
                                const int b = seq / p.J, j = seq % p.J;      // b >= B when seq >= nseq: OOB -> zero fill
                                tma_load_4d(dst + blk * Cfg::BLK + f * ATS_SLAB * 128, &tmQKV, bar, col16 + blk * 64, j, 0, b);
                            }
                        }
                    }
                };
                const int col16 = (h * HD / 32) * 64;
                mbar_wait(&qk_empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&qk_full[s], 2 * Cfg::TILE_BYTES);
                    load_tile(set, &qk_full[s], col16);                                   // Q
                    load_tile(set + Cfg::TILE_BYTES, &qk_full[s], 2 * p.C + col16);       // K
                }
                mbar_wait(&v_empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&v_full[s], Cfg::TILE_BYTES);
                    load_tile(set + 2 * Cfg::TILE_BYTES, &v_full[s], 4 * p.C + col16);    // V
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        constexpr uint32_t idesc_s_h = umma_idesc_fmt(128, 128, 0, 0, 0, 0);
        constexpr uint32_t idesc_s_8 = umma_idesc_fmt(128, 128, 1, 1, 0, 0);
        constexpr uint32_t idesc_o_h = umma_idesc_fmt(128, 32, 0, 0, 0, 1);
        constexpr uint32_t idesc_o_8 = umma_idesc_fmt(128, 32, 1, 1, 0, 1);
        auto issue_S = [&](int i) {
            const int s = i & 1;
            mbar_wait(&qk_full[s], (i >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sQ = smem_u32(smem + s * Cfg::SET_BYTES);
                const uint32_t sK = sQ + Cfg::TILE_BYTES;
                const uint32_t tS = tmem_base + s * Cfg::TMEM_SET;
#pragma unroll
                for (int blk = 0; blk < Cfg::NBLK; ++blk) {
                    const uint64_t q = umma_smem_desc(sQ + blk * Cfg::BLK, 16, 1024, 2u);
                    const uint64_t k = umma_smem_desc(sK + blk * Cfg::BLK, 16, 1024, 2u);
                    umma_ss(tS, q, k, idesc_s_h, blk != 0);
                    umma_ss(tS, q + 2, k + 2, idesc_s_h, 1);
                    umma_ss_f8(tS, q + 4, k + 6, idesc_s_8, 1);
                    umma_ss_f8(tS, q + 6, k + 4, idesc_s_8, 1);
                }
                tc_commit(&s_full[s]);
                tc_commit(&qk_empty[s]);
            }
            __syncwarp();
        };
        if (n_mine > 0) issue_S(0);
        for (int i = 0; i < n_mine; ++i) {
            if (i + 1 < n_mine) issue_S(i + 1);
            const int s = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            mbar_wait(&p_full[s], ph);
            mbar_wait(&v_full[s], ph);
            mbar_wait(&o_empty[s], ph ^ 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sV = smem_u32(smem + s * Cfg::SET_BYTES + 2 * Cfg::TILE_BYTES);
                const uint32_t tS = tmem_base + s * Cfg::TMEM_SET;
                const uint32_t tO = tS + 128;
#pragma unroll
                for (int c = 0; c < 4; ++c) {                    // 128 keys = 4 chunks of 32 (= the 4 packed sequences)
#pragma unroll
                    for (int blk = 0; blk < Cfg::NBLK; ++blk) {
                        const uint32_t vb = sV + blk * Cfg::BLK + c * 32 * 128;
                        const uint32_t d = tO + blk * 32;
                        umma_ts(d, tS + 32 * c, umma_smem_desc(vb, 1024, 1024, 2u), idesc_o_h, c != 0);
                        umma_ts(d, tS + 32 * c + 8, umma_smem_desc(vb + 16 * 128, 1024, 1024, 2u), idesc_o_h, 1);
                        umma_ts_f8(d, tS + 32 * c + 16, umma_smem_desc(vb + 96, 1024, 1024, 2u), idesc_o_8, 1);
                        umma_ts_f8(d, tS + 32 * c + 24, umma_smem_desc(vb + 64, 1024, 1024, 2u), idesc_o_8, 1);
                    }
                }
                tc_commit(&o_full[s]);
                tc_commit(&v_empty[s]);
            }
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------------------- softmax + output (warps 2..9)
        const int quad = warp & 3;                      // == sequence slot inside the tile
        const int half = (warp - 2) >> 2;
        const int r_in_tile = quad * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t red_sum = smem_u32(smem + Cfg::OFF_RED);           // float [3][128], slot = problem index % 3
        const float sl2 = p.scale_log2e;

        auto softmax = [&](int i) {
            const int s = i & 1;
            const uint32_t tS = tmem_base + s * Cfg::TMEM_SET;
            mbar_wait(&s_full[s], (i >> 1) & 1);
            tc_fence_after();
            if (half == 0) {
                uint32_t r[32];
                tmem_ld32(tS + lane_off + quad * 32, r);
                tmem_ld_wait();
                uint32_t hh[16], l8[8], g8[8];
                const float sum = (p.L == 17) ? softmax32_f16c<17>(r, 17, sl2, hh, l8, g8) : softmax32_f16c<0>(r, p.L, sl2, hh, l8, g8);
                tmem_st16(tS + lane_off + quad * 32, hh);
                tmem_st8(tS + lane_off + quad * 32 + 16, l8);
                tmem_st8(tS + lane_off + quad * 32 + 24, g8);
                sts_f32(red_sum + ((i % 3) * 128 + r_in_tile) * 4, sum);
            } else {
                // zero the three off-diagonal blocks of these rows so that sequences do not mix in P V
                uint32_t z[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) z[k] = 0u;
#pragma unroll
                for (int blk = 0; blk < 4; ++blk) {
                    if (blk != quad) {
                        tmem_st16(tS + lane_off + blk * 32, z);
                        tmem_st16(tS + lane_off + blk * 32 + 16, z);
                    }
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[s]);
        };
        auto output = [&](int i) {
            const int s = i & 1;
            const int prob = blockIdx.x + i * gridDim.x;
            const int h = prob % p.H, g = prob / p.H;
            const uint32_t tO = tmem_base + s * Cfg::TMEM_SET + 128;
            named_bar_sync(1, ATT_SM_THREADS);           // red_sum slot written by the half-0 warps is visible
            const float inv = 1.0f / lds_f32(red_sum + ((i % 3) * 128 + r_in_tile) * 4);
            mbar_wait(&o_full[s], (i >> 1) & 1);
            tc_fence_after();
            const int seq = g * ATS_FRAMES + quad;
            const bool ok = (lane < p.L) && (seq < p.nseq);
            const size_t tok = !TEMPORAL ? static_cast<size_t>(seq) * p.J + lane
                                         : (static_cast<size_t>(seq / p.J) * p.F + lane) * p.J + (seq % p.J);
            if (HD == 64 || half == 0) {
                const int c0 = (HD == 64) ? half * 32 : 0;
                uint32_t r[32];
                tmem_ld32(tO + lane_off + c0, r);
                tmem_ld_wait();
                if (ok) {
                    uint8_t* rowp = p.out + tok * p.C * 4;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        float xv[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) xv[k] = __uint_as_float(r[16 * q + k]) * inv;
                        store16_f16c(rowp, h * HD + c0 + 16 * q, xv);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&o_empty[s]);
        };
        if (n_mine > 0) softmax(0);
        for (int i = 0; i < n_mine; ++i) {
            if (i + 1 < n_mine) softmax(i + 1);
            output(i);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace mb
