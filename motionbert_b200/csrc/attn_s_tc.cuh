// Spatial multi-head attention on tcgen05 (DSTformer.py:178-186, `Attention.forward_spatial`).
//
// One (frame, head) problem is softmax(q k^T d^-1/2) v over the J = 17 joints of the frame: far below the
// 128-row UMMA tile.  Four frames are therefore packed into one tile: a 4-D TMA tensor map
// (col, joint, frame, plane) with box (d, 32, 4, planes) lands each frame's 17 joint rows in its own 32-row
// slab (rows 17..31 zero-filled by TMA), giving Q, K and V tiles of 128 rows:
//   S = Q K^T (128 x 128): only the four 32x32 diagonal blocks are meaningful (same frame);
//   softmax: thread == row; warps of column-half 0 read their frame's 32-column block, write the probabilities
//            back in place as bf16 hi/lo, warps of half 1 ZERO the three off-diagonal blocks of the same rows;
//   O = P V: tcgen05.mma with P from TMEM (block-diagonal, so frames do not mix), V MN-major from smem;
//   epilogue: O / rowsum -> bf16 hi/lo planes of the token-major (M, C) attention output.
// Same warp roles / barrier protocol as attn_t_tc.cuh (one q-tile per problem).
#pragma once
#include "attn_t_tc.cuh"

namespace mb {

constexpr int ATS_FRAMES = 4;     // frames per tile
constexpr int ATS_SLAB = 32;      // rows reserved per frame (J <= 32)

// The same packed kernel also serves TEMPORAL attention of short clips (F <= 32 frames: T=27 of BASELINE config 5,
// T=16 / T=30 of the mesh / PoseTrack datasets): there a "sequence" is one (batch, joint) pair, its rows are the F
// frames (token stride J), and four such sequences share a tile -- instead of one 128-row tile per sequence with
// 79 % padding in attn_t_tc.cuh.
struct AttnSParams {
    int nseq;     // sequences: B*F frames (spatial) or B*J (batch, joint) pairs (temporal-packed)
    int L;        // valid rows per sequence: J (spatial) or F <= 32 (temporal-packed)
    int F, J;     // clip length and joints (token index math of the temporal-packed mode)
    int C, H;
    float scale_log2e;
    __nv_bfloat16* out_hi;   // [M, C]
    __nv_bfloat16* out_lo;
    int out_f16c;            // != 0: out_hi is an F16C row buffer [M][C] (see AttnTParams)
};

template <int HD, int PASSES>
struct AttnSCfg {
    static constexpr int SWZ = HD * 2;
    static constexpr uint32_t LAYOUT = (SWZ == 128) ? 2u : 4u;
    static constexpr int PLANES = (PASSES == 3) ? 2 : 1;
    static constexpr int PLANE = 128 * SWZ;
    static constexpr int TILE_BYTES = PLANES * PLANE;          // 32 KB (d=64, 3 passes)
    static constexpr int SET_BYTES = 3 * TILE_BYTES;           // Q | K | V of one problem
    static constexpr int OFF_BAR = 2 * SET_BYTES;              // two problems in flight
    static constexpr int OFF_RED = OFF_BAR + 256;
    static constexpr int SMEM_BYTES = OFF_RED + 3 * 128 * 4 + 1024;
    static constexpr int TMEM_SET = 192;                       // S/P 128 columns + O 64 columns per problem
};

// Software-pipelined over problems i = 0, 1, ... of this CTA (two operand sets in smem, two S/P/O sets in TMEM):
//   TMA      : loads set i+1 while set i is in use
//   MMA      : S(i+1) = Q K^T is issued BEFORE waiting for the probabilities of problem i, so it runs under
//              softmax(i); O(i) = P V runs under softmax(i+1)
//   softmax  : softmax(i+1) precedes the output epilogue of problem i
template <int HD, int PASSES, bool TEMPORAL>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_s_tc_kernel(const __grid_constant__ CUtensorMap tmQKV,   // spatial : 4-D (3C, J, BF, plane), box (HD, 32, 4, PLANES)
                                                              // temporal: 5-D (3C, J, F, B, plane), box (HD, 1, 32, 1, 1)
                 const AttnSParams p) {
    using Cfg = AttnSCfg<HD, PASSES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* qk_full = bars + 0;    // [2] Q and K of a problem landed
    uint64_t* qk_empty = bars + 2;   // [2] S = Q K^T of that problem retired (Q/K slots reusable early)
    uint64_t* v_full = bars + 4;     // [2]
    uint64_t* v_empty = bars + 6;    // [2] P V retired
    uint64_t* s_full = bars + 8;     // [2]
    uint64_t* p_full = bars + 10;    // [2]
    uint64_t* o_full = bars + 12;    // [2]
    uint64_t* o_empty = bars + 14;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = warp_uniform(threadIdx.x >> 5);      // uniform role dispatch (see ptx.cuh elect_one)
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
        if (elect_one()) {
            for (int i = 0; i < n_mine; ++i) {
                const int prob = blockIdx.x + i * gridDim.x;
                const int h = prob % p.H, g = prob / p.H;
                const int f0 = g * ATS_FRAMES;
                const int s = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                uint8_t* set = smem + s * Cfg::SET_BYTES;
                // one operand tile = [plane][4 slabs x 32 rows][HD]; in the temporal mode every (slab, plane) is its own
                // 32-frame box of one (batch, joint) sequence (sequences past the end are fully out of bounds -> zeros)
                auto load_tile = [&](uint8_t* dst, uint64_t* bar, int col) {
                    if (!TEMPORAL) {
                        tma_load_4d(dst, &tmQKV, bar, col, 0, f0, 0);
                    } else {
                        for (int pl = 0; pl < Cfg::PLANES; ++pl)
                            for (int f = 0; f < ATS_FRAMES; ++f) {
                                const int seq = f0 + f;
                                const int b = seq / p.J, j = seq % p.J;      // b >= B when seq >= nseq: OOB -> zero fill
                                tma_load_5d(dst + pl * Cfg::PLANE + f * ATS_SLAB * Cfg::SWZ, &tmQKV, bar, col, j, 0, b, pl);
                            }
                    }
                };
                mbar_wait(&qk_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&qk_full[s], 2 * Cfg::TILE_BYTES);
                load_tile(set, &qk_full[s], h * HD);                                 // Q
                load_tile(set + Cfg::TILE_BYTES, &qk_full[s], p.C + h * HD);         // K
                mbar_wait(&v_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[s], Cfg::TILE_BYTES);
                load_tile(set + 2 * Cfg::TILE_BYTES, &v_full[s], 2 * p.C + h * HD);  // V
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);
        auto issue_S = [&](int i) {
            const int s = i & 1;
            mbar_wait(&qk_full[s], (i >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sQ = smem_u32(smem + s * Cfg::SET_BYTES);
                const uint32_t sK = sQ + Cfg::TILE_BYTES;
                const uint32_t tS = tmem_base + s * Cfg::TMEM_SET;
                const uint64_t q_hi = umma_smem_desc(sQ, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                const uint64_t q_lo = umma_smem_desc(sQ + Cfg::PLANE, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                const uint64_t k_hi = umma_smem_desc(sK, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                const uint64_t k_lo = umma_smem_desc(sK + Cfg::PLANE, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
#pragma unroll
                for (int ks = 0; ks < HD / 16; ++ks) {
                    const uint64_t koff = static_cast<uint64_t>((ks * 32) >> 4);
                    if (PASSES == 3) {
                        umma_ss(tS, q_lo + koff, k_hi + koff, idesc_s, ks != 0);
                        umma_ss(tS, q_hi + koff, k_lo + koff, idesc_s, 1);
                        umma_ss(tS, q_hi + koff, k_hi + koff, idesc_s, 1);
                    } else {
                        umma_ss(tS, q_hi + koff, k_hi + koff, idesc_s, ks != 0);
                    }
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
                for (int ks = 0; ks < 8; ++ks) {                 // 128 keys = 8 K-steps of 16
                    const uint32_t a_hi = tS + 32 * (ks >> 1) + 8 * (ks & 1);
                    const uint32_t a_lo = a_hi + 16;
                    const uint32_t voff = static_cast<uint32_t>(ks) * 16 * Cfg::SWZ;
                    const uint64_t v_hi = umma_smem_desc(sV + voff, 8 * Cfg::SWZ, 8 * Cfg::SWZ, Cfg::LAYOUT);
                    const uint64_t v_lo = umma_smem_desc(sV + Cfg::PLANE + voff, 8 * Cfg::SWZ, 8 * Cfg::SWZ, Cfg::LAYOUT);
                    if (PASSES == 3) {
                        umma_ts(tO, a_lo, v_hi, idesc_o, ks != 0);
                        umma_ts(tO, a_hi, v_lo, idesc_o, 1);
                        umma_ts(tO, a_hi, v_hi, idesc_o, 1);
                    } else {
                        umma_ts(tO, a_hi, v_hi, idesc_o, ks != 0);
                    }
                }
                tc_commit(&o_full[s]);
                tc_commit(&v_empty[s]);
            }
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------------------- softmax + output (warps 2..9)
        const int quad = warp & 3;                      // == frame slot inside the tile
        const int half = (warp - 2) >> 2;
        const int r_in_tile = quad * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        float* red_sum = reinterpret_cast<float*>(smem + Cfg::OFF_RED);   // [3][128], slot = problem index % 3
        const float sl2 = p.scale_log2e;

        auto softmax = [&](int i) {
            const int s = i & 1;
            const uint32_t tS = tmem_base + s * Cfg::TMEM_SET;
            mbar_wait(&s_full[s], (i >> 1) & 1);
            tc_fence_after();
            if (half == 0) {
                // my frame's 32-column diagonal block: columns [32*quad, 32*quad + 32), first J valid
                uint32_t r[32];
                tmem_ld32(tS + lane_off + quad * 32, r);
                tmem_ld_wait();
                float mx = -INFINITY;
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (k < p.L) mx = fmaxf(mx, __uint_as_float(r[k]));
                const float mxs = mx * sl2;
                float sum = 0.f;
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    float p0 = (2 * k < p.L) ? ex2_approx(fmaf(__uint_as_float(r[2 * k]), sl2, -mxs)) : 0.f;
                    float p1 = (2 * k + 1 < p.L) ? ex2_approx(fmaf(__uint_as_float(r[2 * k + 1]), sl2, -mxs)) : 0.f;
                    sum += p0 + p1;
                    split2(p0, p1, hi[k], lo[k]);
                }
                tmem_st16(tS + lane_off + quad * 32, hi);
                if (PASSES == 3) tmem_st16(tS + lane_off + quad * 32 + 16, lo);
                red_sum[(i % 3) * 128 + r_in_tile] = sum;
            } else {
                // zero the three off-diagonal blocks of these rows so that frames do not mix in P V
                uint32_t z[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) z[k] = 0u;
#pragma unroll
                for (int blk = 0; blk < 4; ++blk) {
                    if (blk != quad) {
                        tmem_st16(tS + lane_off + blk * 32, z);
                        if (PASSES == 3) tmem_st16(tS + lane_off + blk * 32 + 16, z);
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
            const float inv = 1.0f / red_sum[(i % 3) * 128 + r_in_tile];
            mbar_wait(&o_full[s], (i >> 1) & 1);
            tc_fence_after();
            const int seq = g * ATS_FRAMES + quad;
            const bool ok = (lane < p.L) && (seq < p.nseq);
            // token row of (sequence, lane): frame-major spatial tokens, or frame `lane` of the (b, j) sequence
            const size_t tok = !TEMPORAL ? static_cast<size_t>(seq) * p.J + lane
                                         : (static_cast<size_t>(seq / p.J) * p.F + lane) * p.J + (seq % p.J);
            if (HD == 64 || half == 0) {
                const int c0 = (HD == 64) ? half * 32 : 0;
                uint32_t r[32];
                tmem_ld32(tO + lane_off + c0, r);
                tmem_ld_wait();
                if (ok && p.out_f16c) {
                    uint8_t* rowp = reinterpret_cast<uint8_t*>(p.out_hi) + tok * p.C * 4;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        float xv[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) xv[k] = __uint_as_float(r[16 * q + k]) * inv;
                        store16_f16c(rowp, h * HD + c0 + 16 * q, xv);
                    }
                } else if (ok) {
                    const size_t ob = tok * p.C + h * HD + c0;
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        split2(__uint_as_float(r[2 * k]) * inv, __uint_as_float(r[2 * k + 1]) * inv, hi[k], lo[k]);
                    uint4* h4 = reinterpret_cast<uint4*>(p.out_hi + ob);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        h4[k] = make_uint4(hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
                    if (p.out_lo) {
                        uint4* l4 = reinterpret_cast<uint4*>(p.out_lo + ob);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            l4[k] = make_uint4(lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
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
