// Temporal multi-head attention in the F16C arithmetic mode (DSTformer.py:188-200, `Attention.forward_temporal`).
//
// Same problem decomposition as attn_t_tc.cuh -- one problem = one (batch, joint, head), softmax(Q K^T d^-1/2) V over
// the F <= 256 frames of that joint, Q / K / V gathered from the token-major (B*F*J, 3C) qkv buffer by TMA -- but
//   * operands are F16C rows (ptx.cuh): per 32 head channels one 128-byte block [32 f16 | 32 lo8 | 32 hi8];
//     S = Q K^T is 2 fp16 MMAs + 2 e5m2 MMAs per block (2 pass-equivalents instead of 3), O = P V likewise with the
//     probabilities written back into TMEM in place as [16 cols f16 | 8 cols lo8 | 8 cols hi8] per 32 keys;
//   * the keys are split into two halves of <= 128 and the TMEM holds THREE 128-column score buffers plus TWO output
//     accumulators (3 x 128 + 2 x 64 = 512 columns), so the tensor pipe never waits for the softmax threads:
//       S(t+1) is issued while softmax(t) runs, P_a V_a runs under the exponentials of half b, and the output
//       epilogue of tile t-1 is executed between the two halves of tile t.
//     Buffer of half g (global half counter) = g % 3; a half may be overwritten once the P V product of half g - 3 has
//     been issued, which the single MMA-issuing thread guarantees by program order (tcgen05.mma executes in order).
// Warp roles (576 threads): w0 TMA producer, w1 MMA issuer (+ TMEM alloc), w2..w17 softmax / output: 4 threads per
// query row, thread `grp` owns key chunk `grp` (32 keys) of each half and 16 of the output columns.
#pragma once
#include "attn_t_tc.cuh"

namespace mb {

struct AttnT16Params {
    int B, F, J, C, H;
    int NK;                  // round_up(F, 32) keys per problem
    float scale_log2e;       // d^-1/2 * log2(e)
    uint8_t* out;            // F16C rows [M][C]
};

template <int HD>
struct AttnT16Cfg {
    static constexpr int NBLK = HD / 32;                       // F16C blocks per head row
    static constexpr int Q_BLK = ATT_BM * 128;                 // one block column of the 128-row Q tile
    static constexpr int Q_BYTES = NBLK * Q_BLK;
    static constexpr int KV_MAX_BYTES = NBLK * ATT_MAXK * 128;
    static constexpr int OFF_K = 0;
    static constexpr int OFF_V = KV_MAX_BYTES;
    static constexpr int OFF_Q = 2 * KV_MAX_BYTES;
    static constexpr int OFF_BAR = OFF_Q + 2 * Q_BYTES;
    static constexpr int OFF_RED = OFF_BAR + 256;
    static constexpr int SMEM_BYTES = OFF_RED + 2 * ATT_T_GROUPS * 128 * 4 + 1024;
    static constexpr uint32_t TM_X = 0;                        // three score / probability buffers of 128 columns
    static constexpr uint32_t TM_O = 384;                      // two output accumulators of HD columns
};

template <int HD>
__global__ void __launch_bounds__(ATT_T_THREADS, 1)
attn_t16_kernel(const __grid_constant__ CUtensorMap tmQ,    // 4-D (6C x 16-bit, J, F, B), box (64, 1, 128, 1)
                const __grid_constant__ CUtensorMap tmKV,   //                              box (64, 1, NK , 1)
                const AttnT16Params p) {
    using Cfg = AttnT16Cfg<HD>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* k_full = bars + 0;
    uint64_t* k_empty = bars + 1;
    uint64_t* v_full = bars + 2;
    uint64_t* v_empty = bars + 3;
    uint64_t* q_full = bars + 4;    // [2]
    uint64_t* q_empty = bars + 6;   // [2]
    uint64_t* s_full = bars + 8;
    uint64_t* pa_full = bars + 9;
    uint64_t* pb_full = bars + 10;
    uint64_t* o_full = bars + 11;   // [2]
    uint64_t* o_empty = bars + 13;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = warp_uniform(threadIdx.x >> 5);          // role dispatch on a value nvcc knows is warp-uniform
    const int lane = threadIdx.x & 31;
    const int num_prob = p.B * p.J * p.H;
    const int num_qt = (p.F + ATT_BM - 1) / ATT_BM;
    const int n_mine = (num_prob > static_cast<int>(blockIdx.x))
                           ? (num_prob - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
    const int T = n_mine * num_qt;                            // q-tiles of this CTA
    const int nh = p.NK > 128 ? 2 : 1;                        // key halves per tile
    const int nk_a = p.NK > 128 ? 128 : p.NK;
    const int nk_b = p.NK - nk_a;
    const int kv_blk = p.NK * 128;                            // bytes per block column of K (or V) in smem
    const uint32_t kv_bytes = Cfg::NBLK * kv_blk;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        mbar_init(k_full, 1);  mbar_init(k_empty, 1);
        mbar_init(v_full, 1);  mbar_init(v_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], ATT_T_SM_THREADS);
        }
        mbar_init(s_full, 1);
        mbar_init(pa_full, ATT_T_SM_THREADS);
        mbar_init(pb_full, ATT_T_SM_THREADS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto prob_of = [&](int t, int& h, int& j, int& b) {
        const int prob = blockIdx.x + (t / num_qt) * gridDim.x;
        h = prob % p.H; j = (prob / p.H) % p.J; b = prob / (p.H * p.J);
    };

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        // (every lane walks the warp-uniform loop, one elected lane issues: operands stay in uniform registers)
        for (int t = 0; t < T; ++t) {
            int h, j, b;
            prob_of(t, h, j, b);
            const int i = t / num_qt, qt = t % num_qt;
            const uint32_t kv_ph = i & 1;
            const int col16 = (h * HD / 32) * 64;                 // 16-bit-unit column of the head's first block (q part)
            if (qt == 0) {
                mbar_wait(k_empty, kv_ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(k_full, kv_bytes);
                    for (int blk = 0; blk < Cfg::NBLK; ++blk)
                        tma_load_4d(smem + Cfg::OFF_K + blk * kv_blk, &tmKV, k_full, 2 * p.C + col16 + blk * 64, j, 0, b);
                }
            }
            const int qs = t & 1;
            mbar_wait(&q_empty[qs], ((t >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&q_full[qs], Cfg::Q_BYTES);
                for (int blk = 0; blk < Cfg::NBLK; ++blk)
                    tma_load_4d(smem + Cfg::OFF_Q + qs * Cfg::Q_BYTES + blk * Cfg::Q_BLK, &tmQ, &q_full[qs],
                                col16 + blk * 64, j, qt * ATT_BM, b);
            }
            if (qt == 0) {
                mbar_wait(v_empty, kv_ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(v_full, kv_bytes);
                    for (int blk = 0; blk < Cfg::NBLK; ++blk)
                        tma_load_4d(smem + Cfg::OFF_V + blk * kv_blk, &tmKV, v_full, 4 * p.C + col16 + blk * 64, j, 0, b);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        const uint32_t sK = smem_u32(smem + Cfg::OFF_K);
        const uint32_t sV = smem_u32(smem + Cfg::OFF_V);
        constexpr uint32_t idesc_o_h = umma_idesc_fmt(ATT_BM, 32, 0, 0, 0, 1);     // O block += P V : V is MN-major
        constexpr uint32_t idesc_o_8 = umma_idesc_fmt(ATT_BM, 32, 1, 1, 0, 1);

        // S(t) = Q K^T for both key halves of q-tile t
        auto issue_S = [&](int t) {
            const int i = t / num_qt, qt = t % num_qt;
            if (qt == 0) mbar_wait(k_full, i & 1);
            const int qs = t & 1;
            mbar_wait(&q_full[qs], (t >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sQ = smem_u32(smem + Cfg::OFF_Q + qs * Cfg::Q_BYTES);
                for (int hf = 0; hf < nh; ++hf) {
                    const int n = hf == 0 ? nk_a : nk_b;
                    const uint32_t tS = tmem_base + Cfg::TM_X + 128u * static_cast<uint32_t>((t * nh + hf) % 3);
                    const uint32_t idesc_h = umma_idesc_fmt(ATT_BM, n, 0, 0, 0, 0);
                    const uint32_t idesc_8 = umma_idesc_fmt(ATT_BM, n, 1, 1, 0, 0);
#pragma unroll
                    for (int blk = 0; blk < Cfg::NBLK; ++blk) {
                        const uint64_t q = umma_smem_desc(sQ + blk * Cfg::Q_BLK, 16, 1024, 2u);
                        const uint64_t k = umma_smem_desc(sK + blk * kv_blk + hf * 128 * 128, 16, 1024, 2u);
                        umma_ss(tS, q, k, idesc_h, blk != 0);                 // f16 slices at +0 / +32 B
                        umma_ss(tS, q + 2, k + 2, idesc_h, 1);
                        umma_ss_f8(tS, q + 4, k + 6, idesc_8, 1);              // (ql 2^6)(kh 2^-6)
                        umma_ss_f8(tS, q + 6, k + 4, idesc_8, 1);              // (qh 2^-6)(kl 2^6)
                    }
                }
                tc_commit(s_full);
                tc_commit(&q_empty[qs]);
                if (qt == num_qt - 1) tc_commit(k_empty);
            }
            __syncwarp();
        };
        // O(t) (+)= P V over key half hf of q-tile t
        auto issue_PV = [&](int t, int hf) {
            if (elect_one()) {
                const int n = hf == 0 ? nk_a : nk_b;
                const uint32_t tP = tmem_base + Cfg::TM_X + 128u * static_cast<uint32_t>((t * nh + hf) % 3);
                const uint32_t tO = tmem_base + Cfg::TM_O + static_cast<uint32_t>((t & 1) * HD);
                for (int c = 0; c < n / 32; ++c) {
                    const uint32_t row0 = static_cast<uint32_t>(hf * 128 + c * 32) * 128u;      // byte offset of the chunk's first key row
#pragma unroll
                    for (int blk = 0; blk < Cfg::NBLK; ++blk) {
                        const uint32_t vb = sV + blk * kv_blk + row0;
                        const uint32_t d = tO + blk * 32;
                        const uint32_t first = (hf == 0 && c == 0) ? 0u : 1u;
                        // MN-major B operand: 32 channels x K keys, SBO = 8 key rows (1024 B); slices of the 128-byte block
                        umma_ts(d, tP + 32 * c, umma_smem_desc(vb, 1024, 1024, 2u), idesc_o_h, first);
                        umma_ts(d, tP + 32 * c + 8, umma_smem_desc(vb + 16 * 128, 1024, 1024, 2u), idesc_o_h, 1);
                        umma_ts_f8(d, tP + 32 * c + 16, umma_smem_desc(vb + 96, 1024, 1024, 2u), idesc_o_8, 1);   // pl8 x vh8
                        umma_ts_f8(d, tP + 32 * c + 24, umma_smem_desc(vb + 64, 1024, 1024, 2u), idesc_o_8, 1);   // ph8 x vl8
                    }
                }
            }
            __syncwarp();
        };

        if (T > 0) issue_S(0);
        for (int t = 0; t < T; ++t) {
            const int i = t / num_qt, qt = t % num_qt;
            mbar_wait(pa_full, t & 1);
            if (qt == 0) mbar_wait(v_full, i & 1);
            mbar_wait(&o_empty[t & 1], ((t >> 1) & 1) ^ 1);
            tc_fence_after();
            issue_PV(t, 0);
            if (t + 1 < T) issue_S(t + 1);
            if (nh == 2) {
                mbar_wait(pb_full, t & 1);
                tc_fence_after();
                issue_PV(t, 1);
            }
            if (elect_one()) {
                tc_commit(&o_full[t & 1]);
                if (qt == num_qt - 1) tc_commit(v_empty);
            }
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------------------- softmax + output (warps 2..17)
        const int quad = warp & 3;
        const int grp = (warp - 2) >> 2;                // key chunk of each half / 16-column slice of the output
        const int r_in_tile = quad * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t red_max = smem_u32(smem + Cfg::OFF_RED) + 4u * r_in_tile;     // float [4][128], my row
        const uint32_t red_sum = red_max + ATT_T_GROUPS * 128 * 4;                   // float [4][128]
        const float sl2 = p.scale_log2e;
        const bool has_a = grp * 32 < nk_a;
        const bool has_b = nh == 2 && grp * 32 < nk_b;
        const int key_a = grp * 32, key_b = 128 + grp * 32;
        const bool full_a = key_a + 32 <= p.F, full_b = key_b + 32 <= p.F;

        // (clip, joint, head) of a q-tile; advanced incrementally (one set of integer divisions per PROBLEM, by hand-over)
        struct TileId { int h, j, b, qt; };
        auto tile_id = [&](int t) {
            TileId id;
            const int prob = blockIdx.x + (t / num_qt) * gridDim.x;
            id.h = prob % p.H; id.j = (prob / p.H) % p.J; id.b = prob / (p.H * p.J); id.qt = t % num_qt;
            return id;
        };
        auto epilogue = [&](int t, const TileId& id, float inv) {
            mbar_wait(&o_full[t & 1], (t >> 1) & 1);
            tc_fence_after();
            const int fr = id.qt * ATT_BM + r_in_tile;
            if (grp * 16 < HD) {
                uint32_t r[16];
                tmem_ld16(tmem_base + Cfg::TM_O + (t & 1) * HD + lane_off + grp * 16, r);
                tmem_ld_wait();
                if (fr < p.F) {
                    const size_t tok = (static_cast<size_t>(id.b) * p.F + fr) * p.J + id.j;
                    float xv[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) xv[i] = __uint_as_float(r[i]) * inv;
                    store16_f16c(p.out + tok * p.C * 4, id.h * HD + grp * 16, xv);
                }
            }
            tc_fence_before();
            mbar_arrive(&o_empty[t & 1]);
        };
        auto s_addr = [&](int t, int hf) {
            return tmem_base + Cfg::TM_X + 128u * static_cast<uint32_t>((t * nh + hf) % 3) + lane_off + grp * 32;
        };
        auto row_max = [&](const uint32_t (&r)[32], int key0, bool full, float mx) {
            if (full) {
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (key0 + i < p.F) mx = fmaxf(mx, __uint_as_float(r[i]));
            }
            return mx;
        };
        // probabilities of one chunk (32 scores in r): exponentials, partial row sum, in-place F16C write-back over S
        auto exp_store = [&](uint32_t tS, const uint32_t (&r)[32], int key0, bool full, float mxs, float& sum) {
            uint32_t hh[16], l8[8], g8[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float pv[8];
                if (full) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) pv[i] = ex2_approx(fmaf(__uint_as_float(r[8 * q + i]), sl2, -mxs));
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        pv[i] = key0 + 8 * q + i < p.F ? ex2_approx(fmaf(__uint_as_float(r[8 * q + i]), sl2, -mxs)) : 0.f;
                }
                sum += ((pv[0] + pv[1]) + (pv[2] + pv[3])) + ((pv[4] + pv[5]) + (pv[6] + pv[7]));
                uint32_t h4[4], l2[2], g2[2];
                split8_f16c(pv, h4, l2, g2);
#pragma unroll
                for (int i = 0; i < 4; ++i) hh[4 * q + i] = h4[i];
                l8[2 * q] = l2[0]; l8[2 * q + 1] = l2[1];
                g8[2 * q] = g2[0]; g8[2 * q + 1] = g2[1];
            }
            tmem_st16(tS, hh);
            tmem_st8(tS + 16, l8);
            tmem_st8(tS + 24, g8);
        };

        float inv_prev = 0.f;
        TileId id_prev = {0, 0, 0, 0};
        for (int t = 0; t < T; ++t) {
            mbar_wait(s_full, t & 1);
            tc_fence_after();
            // pass 1: row max of the raw scores over my chunks (scale > 0 commutes with max); half b first, so that the
            // scores of half a stay in registers for the exponentials (one TMEM read less per tile)
            float mx = -INFINITY;
            uint32_t ra[32];
            if (has_b) {
                uint32_t rb[32];
                tmem_ld32(s_addr(t, 1), rb);
                tmem_ld_wait();
                mx = row_max(rb, key_b, full_b, mx);
            }
            if (has_a) {
                tmem_ld32(s_addr(t, 0), ra);
                tmem_ld_wait();
                mx = row_max(ra, key_a, full_a, mx);
            }
            sts_f32(red_max + grp * 512, mx);
            named_bar_sync(1, ATT_T_SM_THREADS);
            mx = fmaxf(fmaxf(lds_f32(red_max), lds_f32(red_max + 512)), fmaxf(lds_f32(red_max + 1024), lds_f32(red_max + 1536)));
            const float mxs = mx * sl2;
            float sum = 0.f;
            if (has_a) exp_store(s_addr(t, 0), ra, key_a, full_a, mxs, sum);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(pa_full);
            if (t > 0) epilogue(t - 1, id_prev, inv_prev);             // under the tensor pipe's P_a V_a / S(t+1)
            if (nh == 2) {
                if (has_b) {
                    uint32_t rb[32];
                    tmem_ld32(s_addr(t, 1), rb);
                    tmem_ld_wait();
                    exp_store(s_addr(t, 1), rb, key_b, full_b, mxs, sum);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(pb_full);
            }
            sts_f32(red_sum + grp * 512, sum);
            id_prev = tile_id(t);
            named_bar_sync(1, ATT_T_SM_THREADS);
            inv_prev = 1.0f / ((lds_f32(red_sum) + lds_f32(red_sum + 512)) + (lds_f32(red_sum + 1024) + lds_f32(red_sum + 1536)));
        }
        if (T > 0) epilogue(T - 1, id_prev, inv_prev);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace mb
