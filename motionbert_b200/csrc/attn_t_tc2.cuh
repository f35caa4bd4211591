// EXPERIMENTAL (test flag MB_FLAG_ATTN_T_V2, not the product path): measured SLOWER than attn_t_tc.cuh on B200
// (60.5 ms vs 44 ms per forward at config 2, profiles/README.md) -- the 256 softmax threads, not the tensor pipe,
// bound this kernel (MUFU.EX2 + two F2FP per score pair), and the smem ring adds STS + proxy fences to exactly them.
// Kept as the documented negative result of round 1.
//
// Temporal attention, second generation: probabilities go to the tensor core through SHARED memory
// (DSTformer.py:188-200, `Attention.forward_temporal`).
//
// attn_t_tc.cuh keeps P in TMEM (in place over S).  At T = 243 that pins all the TMEM a tile can use
// (S 256 columns + O 64), so S-MMA -> softmax -> PV-MMA of one 128-query tile run strictly one after the other and
// the tensor pipe idles while 256 threads do the softmax (profiles/r01v6: 26.6 % active).  Here the softmax threads
// write P (bf16 hi/lo) into a 4-slot smem ring of [128 queries x 32 keys] K-major SWIZZLE_64B operand tiles:
//   * S is dead as soon as the softmax has READ it (s_free), so the MMA warp issues S(t+1) = Q K^T while the
//     softmax of tile t is still producing its last chunk, and P V of tile t streams chunk by chunk behind it;
//   * O is double buffered in TMEM (S 256 + 2 x 64 columns), the output epilogue of tile t-1 runs after the
//     softmax of tile t, i.e. under P V (t).
// Everything else (5-D TMA gather of the strided (b, joint) sequence, BF16x3 passes, MN-major V) is as in v1.
#pragma once
#include "attn_t_tc.cuh"

namespace mb {

constexpr int AT2_SLOTS = 4;          // P ring: slots 0,1 <- key half 0 (chunks 0..3), slots 2,3 <- key half 1 (chunks 4..7)
constexpr int AT2_CHUNK_KEYS = 32;

template <int HD, int PASSES>
struct Attn2Cfg {
    static constexpr int SWZ = HD * 2;                            // Q/K/V rows: 128 B (d=64) or 64 B (d=32)
    static constexpr uint32_t LAYOUT = (SWZ == 128) ? 2u : 4u;
    static constexpr int PLANES = (PASSES == 3) ? 2 : 1;
    static constexpr int Q_PLANE = ATT_BM * SWZ;
    static constexpr int Q_BYTES = PLANES * Q_PLANE;
    static constexpr int KV_MAX_BYTES = PLANES * ATT_MAXK * SWZ;
    static constexpr int P_PLANE = ATT_BM * 64;                   // 128 rows x 32 keys bf16 (64-byte rows, SWIZZLE_64B)
    static constexpr int P_SLOT = PLANES * P_PLANE;               // 16 KB
    static constexpr int OFF_K = 0;
    static constexpr int OFF_V = KV_MAX_BYTES;
    static constexpr int OFF_Q = 2 * KV_MAX_BYTES;
    static constexpr int OFF_P = OFF_Q + Q_BYTES;
    static constexpr int OFF_BAR = OFF_P + AT2_SLOTS * P_SLOT;
    static constexpr int OFF_RED = OFF_BAR + 256;                 // float red[2 (tile parity)][2 (key half)][128]
    static constexpr int SMEM_BYTES = OFF_RED + 4 * 128 * 4;      // 231,680 B at d=64 / 3 passes: no room for slack,
                                                                  // the dynamic smem symbol is declared 1024-aligned
};

template <int HD, int PASSES>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_t2_kernel(const __grid_constant__ CUtensorMap tmQ,    // box (HD, 1, 128, 1, PLANES)
               const __grid_constant__ CUtensorMap tmKV,   // box (HD, 1, NK , 1, PLANES)
               const AttnTParams p) {
    using Cfg = Attn2Cfg<HD, PASSES>;
    extern __shared__ __align__(1024) uint8_t smem_al1024[];
    uint8_t* smem = smem_al1024;
    if ((smem_u32(smem) & 1023u) != 0u) {
        if (threadIdx.x == 0) printf("[mb] attn_t2: dynamic smem base not 1024-byte aligned\n");
        __trap();
    }
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* k_full = bars + 0;
    uint64_t* k_empty = bars + 1;
    uint64_t* v_full = bars + 2;
    uint64_t* v_empty = bars + 3;
    uint64_t* q_full = bars + 4;
    uint64_t* q_empty = bars + 5;
    uint64_t* s_full = bars + 6;
    uint64_t* s_free = bars + 7;     // 256 softmax threads: "I have read all my scores of this tile"
    uint64_t* p_full = bars + 8;     // [4] 128 threads of one key half wrote a P chunk
    uint64_t* p_empty = bars + 12;   // [4] the P V MMAs of that chunk retired
    uint64_t* o_full = bars + 16;    // [2]
    uint64_t* o_empty = bars + 18;   // [2] 256 threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_prob = p.B * p.J * p.H;
    const int num_qt = (p.F + ATT_BM - 1) / ATT_BM;
    const int n_prob_mine = (num_prob > static_cast<int>(blockIdx.x))
                                ? (num_prob - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
    const int n_tiles = n_prob_mine * num_qt;
    const int kv_plane = p.NK * Cfg::SWZ;
    const uint32_t kv_bytes = Cfg::PLANES * kv_plane;
    const int nch = (p.NK + AT2_CHUNK_KEYS - 1) / AT2_CHUNK_KEYS;       // 32-key chunks per tile (<= 8)
    const int nch0 = nch < 4 ? nch : 4;                                   // chunks of key half 0
    const int nch1 = nch - nch0;                                          // chunks of key half 1

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        mbar_init(k_full, 1);  mbar_init(k_empty, 1);
        mbar_init(v_full, 1);  mbar_init(v_empty, 1);
        mbar_init(q_full, 1);  mbar_init(q_empty, 1);
        mbar_init(s_full, 1);
        mbar_init(s_free, ATT_SM_THREADS);
        for (int i = 0; i < AT2_SLOTS; ++i) {
            mbar_init(&p_full[i], 128);
            mbar_init(&p_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
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
    const uint32_t tmem_S = tmem_base;                 // columns [0, 256)
    const uint32_t tmem_O0 = tmem_base + ATT_MAXK;     // columns [256, 256 + 2*HD): two output accumulators

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (lane == 0) {
            uint32_t q_it = 0;
            for (int ip = 0; ip < n_prob_mine; ++ip) {
                const int prob = blockIdx.x + ip * gridDim.x;
                const int h = prob % p.H, j = (prob / p.H) % p.J, b = prob / (p.H * p.J);
                const uint32_t kv_ph = ip & 1;
                mbar_wait(k_empty, kv_ph ^ 1);
                mbar_arrive_expect_tx(k_full, kv_bytes);
                tma_load_5d(smem + Cfg::OFF_K, &tmKV, k_full, p.C + h * HD, j, 0, b, 0);
                for (int qt = 0; qt < num_qt; ++qt, ++q_it) {
                    mbar_wait(q_empty, (q_it & 1) ^ 1);
                    mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
                    tma_load_5d(smem + Cfg::OFF_Q, &tmQ, q_full, h * HD, j, qt * ATT_BM, b, 0);
                    if (qt == 0) {
                        mbar_wait(v_empty, kv_ph ^ 1);
                        mbar_arrive_expect_tx(v_full, kv_bytes);
                        tma_load_5d(smem + Cfg::OFF_V, &tmKV, v_full, 2 * p.C + h * HD, j, 0, b, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        const uint32_t idesc_s = umma_idesc_bf16(ATT_BM, p.NK, 0, 0);
        const uint32_t idesc_o = umma_idesc_bf16(ATT_BM, HD, 0, 1);       // B (=V) MN-major
        const uint32_t sK = smem_u32(smem + Cfg::OFF_K);
        const uint32_t sV = smem_u32(smem + Cfg::OFF_V);
        const uint32_t sQ = smem_u32(smem + Cfg::OFF_Q);
        const uint32_t sP = smem_u32(smem + Cfg::OFF_P);
        // S(t) = Q K^T for tile t (problem ip = t / num_qt, q-tile qt = t % num_qt)
        auto issue_S = [&](int t) {
            const int ip = t / num_qt, qt = t % num_qt;
            if (qt == 0) mbar_wait(k_full, ip & 1);
            mbar_wait(q_full, t & 1);
            if (t > 0) mbar_wait(s_free, (t - 1) & 1);          // softmax has read S(t-1) completely
            tc_fence_after();
            if (lane == 0) {
                const uint64_t q_hi = umma_smem_desc(sQ, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                const uint64_t q_lo = umma_smem_desc(sQ + Cfg::Q_PLANE, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                const uint64_t k_hi = umma_smem_desc(sK, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                const uint64_t k_lo = umma_smem_desc(sK + kv_plane, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
#pragma unroll
                for (int ks = 0; ks < HD / 16; ++ks) {
                    const uint64_t koff = static_cast<uint64_t>((ks * 32) >> 4);
                    if (PASSES == 3) {
                        umma_ss(tmem_S, q_lo + koff, k_hi + koff, idesc_s, ks != 0);
                        umma_ss(tmem_S, q_hi + koff, k_lo + koff, idesc_s, 1);
                        umma_ss(tmem_S, q_hi + koff, k_hi + koff, idesc_s, 1);
                    } else {
                        umma_ss(tmem_S, q_hi + koff, k_hi + koff, idesc_s, ks != 0);
                    }
                }
                tc_commit(s_full);
                tc_commit(q_empty);
                if (qt == num_qt - 1) tc_commit(k_empty);
            }
            __syncwarp();
        };
        uint32_t slot_use[AT2_SLOTS] = {0, 0, 0, 0};
        if (n_tiles > 0) issue_S(0);
        for (int t = 0; t < n_tiles; ++t) {
            const int ip = t / num_qt, qt = t % num_qt;
            const uint32_t tO = tmem_O0 + (t & 1) * HD;
            // consumption order alternates the two key halves: 0,4,1,5,2,6,3,7
            int k0 = 0, k1 = 0;
            for (int i = 0; i < nch; ++i) {
                const bool from1 = (k1 < nch1) && ((i & 1) || k0 >= nch0);
                const int c = from1 ? 4 + k1 : k0;
                const int slot = from1 ? 2 + (k1 & 1) : (k0 & 1);
                if (from1) ++k1; else ++k0;
                // S(t+1) is issued as soon as every softmax thread has STARTED its last chunk (s_free), i.e. before the
                // MMA waits for the last chunk of either key half, so it executes under the tail of softmax(t)
                if (i == (nch >= 2 ? nch - 2 : 0) && t + 1 < n_tiles) issue_S(t + 1);
                mbar_wait(&p_full[slot], slot_use[slot] & 1);
                if (i == 0) {
                    if (qt == 0) mbar_wait(v_full, ip & 1);
                    mbar_wait(&o_empty[t & 1], ((t >> 1) & 1) ^ 1);
                }
                tc_fence_after();
                if (lane == 0) {
                    const int keys = (p.NK - c * AT2_CHUNK_KEYS) < AT2_CHUNK_KEYS ? (p.NK - c * AT2_CHUNK_KEYS) : AT2_CHUNK_KEYS;
                    const uint32_t sPs = sP + slot * Cfg::P_SLOT;
                    const uint64_t p_hi = umma_smem_desc(sPs, 16, 8 * 64, 4u);                 // K-major SWIZZLE_64B
                    const uint64_t p_lo = umma_smem_desc(sPs + Cfg::P_PLANE, 16, 8 * 64, 4u);
                    for (int ks = 0; ks < keys / 16; ++ks) {
                        const uint64_t koff = static_cast<uint64_t>((ks * 32) >> 4);
                        const uint32_t voff = static_cast<uint32_t>(c * AT2_CHUNK_KEYS + ks * 16) * Cfg::SWZ;
                        const uint64_t v_hi = umma_smem_desc(sV + voff, 8 * Cfg::SWZ, 8 * Cfg::SWZ, Cfg::LAYOUT);
                        const uint64_t v_lo = umma_smem_desc(sV + kv_plane + voff, 8 * Cfg::SWZ, 8 * Cfg::SWZ, Cfg::LAYOUT);
                        const uint32_t accum = (i != 0 || ks != 0) ? 1u : 0u;
                        if (PASSES == 3) {
                            umma_ss(tO, p_lo + koff, v_hi, idesc_o, accum);
                            umma_ss(tO, p_hi + koff, v_lo, idesc_o, 1);
                            umma_ss(tO, p_hi + koff, v_hi, idesc_o, 1);
                        } else {
                            umma_ss(tO, p_hi + koff, v_hi, idesc_o, accum);
                        }
                    }
                    tc_commit(&p_empty[slot]);
                    if (i == nch - 1) {
                        tc_commit(&o_full[t & 1]);
                        if (qt == num_qt - 1) tc_commit(v_empty);
                    }
                }
                __syncwarp();
                ++slot_use[slot];
            }
        }
    } else {
        // ---------------------------------------------------------------- softmax + output (warps 2..9)
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int r_in_tile = quad * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const int my_nch = half == 0 ? nch0 : nch1;
        const int my_c0 = half * 4;
        // One exchange array serves both the row max of tile t (written before the per-tile barrier, read right
        // after it) and the row sum of tile t (written at the end of tile t, read after the barrier of tile t+1):
        // slot [t & 1] is not rewritten before S(t+2) exists, which needs s_free(t+1) from every thread, i.e. after
        // every thread has read sum(t).
        float* red = reinterpret_cast<float*>(smem + Cfg::OFF_RED);       // [2 (tile parity)][2 (half)][128]
        uint8_t* sP = smem + Cfg::OFF_P;
        const float sl2 = p.scale_log2e;
        const uint32_t sw64 = static_cast<uint32_t>((r_in_tile >> 1) & 3);
        uint32_t slot_use[2] = {0, 0};       // my two ring slots: half*2 + {0,1}

        auto output = [&](int t, float inv) {                // epilogue of tile t (O accumulator t & 1)
            const int ip = t / num_qt, qt = t % num_qt;
            const int prob = blockIdx.x + ip * gridDim.x;
            const int h = prob % p.H, j = (prob / p.H) % p.J, b = prob / (p.H * p.J);
            const uint32_t tO = tmem_O0 + (t & 1) * HD;
            mbar_wait(&o_full[t & 1], (t >> 1) & 1);
            tc_fence_after();
            const int tq = qt * ATT_BM + r_in_tile;
            const bool ok = tq < p.F;
            const size_t tok = (static_cast<size_t>(b) * p.F + (ok ? tq : 0)) * p.J + j;
            if (HD == 64 || half == 0) {
                const int c0 = (HD == 64) ? half * 32 : 0;
                const size_t ob = tok * p.C + h * HD + c0;
                uint32_t r[32];
                tmem_ld32(tO + lane_off + c0, r);
                tmem_ld_wait();
                if (ok) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        split2(__uint_as_float(r[2 * i]) * inv, __uint_as_float(r[2 * i + 1]) * inv, hi[i], lo[i]);
                    uint4* h4 = reinterpret_cast<uint4*>(p.out_hi + ob);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        h4[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                    if (p.out_lo) {
                        uint4* l4 = reinterpret_cast<uint4*>(p.out_lo + ob);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            l4[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&o_empty[t & 1]);
        };

        for (int t = 0; t < n_tiles; ++t) {
            const int par = t & 1;
            mbar_wait(s_full, t & 1);
            tc_fence_after();
            // pass 1: row max over my key chunks (raw scores; scale > 0 commutes with max); the tcgen05.ld of chunk
            // k+1 is in flight while chunk k is reduced (two register buffers, statically indexed)
            float mx = -INFINITY;
            uint32_t rr[2][32];
            if (my_nch > 0) tmem_ld32(tmem_S + lane_off + my_c0 * 32, rr[0]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k < my_nch) {
                    const int c = my_c0 + k;
                    tmem_ld_wait();
                    if (k + 1 < my_nch) tmem_ld32(tmem_S + lane_off + (c + 1) * 32, rr[(k + 1) & 1]);
                    if (c * 32 + 32 <= p.F) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(rr[k & 1][i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c * 32 + i < p.F) mx = fmaxf(mx, __uint_as_float(rr[k & 1][i]));
                    }
                }
            }
            red[(par * 2 + half) * 128 + r_in_tile] = mx;
            named_bar_sync(1, ATT_SM_THREADS);
            mx = fmaxf(red[(par * 2 + 0) * 128 + r_in_tile], red[(par * 2 + 1) * 128 + r_in_tile]);
            // row sum of the PREVIOUS tile (both halves) became visible at the barrier above
            float inv_prev = 0.f;
            if (t > 0)
                inv_prev = 1.0f / (red[((par ^ 1) * 2 + 0) * 128 + r_in_tile] + red[((par ^ 1) * 2 + 1) * 128 + r_in_tile]);
            const float mxs = mx * sl2;
            // pass 2: p = 2^(s*c - max*c) -> bf16 hi/lo -> smem ring slot (K-major, SWIZZLE_64B)
            float sum = 0.f;
            if (my_nch == 0) {
                tc_fence_before();
                mbar_arrive(s_free);
            }
            if (my_nch > 0) tmem_ld32(tmem_S + lane_off + my_c0 * 32, rr[0]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k < my_nch) {
                    const int c = my_c0 + k;
                    const int ls = k & 1;
                    const int slot = half * 2 + ls;
                    tmem_ld_wait();
                    if (k + 1 < my_nch) {
                        tmem_ld32(tmem_S + lane_off + (c + 1) * 32, rr[(k + 1) & 1]);
                    } else {                                    // that was my last read of S(t)
                        tc_fence_before();
                        mbar_arrive(s_free);
                    }
                    uint32_t hi[16], lo[16];
                    const bool full = c * 32 + 32 <= p.F;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float p0 = ex2_approx(fmaf(__uint_as_float(rr[k & 1][2 * i]), sl2, -mxs));
                        float p1 = ex2_approx(fmaf(__uint_as_float(rr[k & 1][2 * i + 1]), sl2, -mxs));
                        if (!full) {
                            if (c * 32 + 2 * i >= p.F) p0 = 0.f;
                            if (c * 32 + 2 * i + 1 >= p.F) p1 = 0.f;
                        }
                        sum += p0 + p1;
                        split2(p0, p1, hi[i], lo[i]);
                    }
                    mbar_wait(&p_empty[slot], (slot_use[ls] & 1) ^ 1);     // P V of the previous user of this slot retired
                    uint8_t* dst = sP + slot * Cfg::P_SLOT + r_in_tile * 64;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        *reinterpret_cast<uint4*>(dst + ((i ^ sw64) << 4)) =
                            make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                        if (PASSES == 3)
                            *reinterpret_cast<uint4*>(dst + Cfg::P_PLANE + ((i ^ sw64) << 4)) =
                                make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&p_full[slot]);
                    ++slot_use[ls];
                }
            }
            named_bar_sync(2, ATT_SM_THREADS);          // both halves have read max(t) before the slot takes sum(t)
            red[(par * 2 + half) * 128 + r_in_tile] = sum;
            // output of the previous tile (its P V ran under this tile's softmax)
            if (t > 0) output(t - 1, inv_prev);
        }
        if (n_tiles > 0) {
            named_bar_sync(1, ATT_SM_THREADS);
            const int par = (n_tiles - 1) & 1;
            const float inv = 1.0f / (red[(par * 2 + 0) * 128 + r_in_tile] + red[(par * 2 + 1) * 128 + r_in_tile]);
            output(n_tiles - 1, inv);
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
