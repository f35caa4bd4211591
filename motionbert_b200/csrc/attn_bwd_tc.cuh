// Attention-core backward on tcgen05 (bf16 single-pass arithmetic, fp32 accumulate).
//
// Forward (DSTformer.py:178-200):  S = Q K^T * c,  P = softmax_row(S),  O = P V          (per batch b, joint j, head h)
// Backward, flash-style (P is recomputed on chip, never stored):
//   delta_i = sum_d dO[i,d] O[i,d]        dP = dO V^T        dS = P (dP - delta) * c
//   dQ = dS K            dK = dS^T Q            dV = P^T dO
// Two kernels, both shaped like the forward temporal kernel (one 128-row tile at a time, 5-D TMA gathers):
//   attn_bwd_q_kernel : rows = queries.  S and dP as two UMMA accumulators (2 x 256 TMEM columns), softmax statistics,
//                       dS written back in place as packed bf16, dQ = dS K with dS read from TMEM and K as an
//                       MN-major operand; also stores log2-sum-exp and delta per query for the second kernel.
//   attn_bwd_kv_kernel: rows = keys.  S^T = K Q^T and dP^T = V dO^T, P^T / dS^T rebuilt from the stored statistics
//                       (column vectors now), dV = P^T dO and dK = dS^T Q with dO / Q as MN-major operands.
// The spatial attention is the same problem with (B' = B*F, F' = J, J' = 1): the host passes those dimensions.
// PACK = true (sequences of <= 32 rows: the 17-joint spatial attention, short temporal clips): FOUR sequences share
// one 128-row tile, each in its own 32-row slab (four 32-row TMA boxes per operand); scores are 128 x 128 with only
// the diagonal 32 x 32 blocks live, the off-diagonal blocks of P / dS are written as zeros so the dQ / dK / dV
// contractions over all 128 rows pick up exactly the rows of the own sequence.
#pragma once
#include "attn_t_tc.cuh"

namespace mb {

// dQ kernel: w0 TMA, w1 MMA, w2..w17 SIMT = FOUR threads per row (TMEM lane quadrant = warp % 4, column quarter =
// (warp - 2) / 4): 11 % faster than two per row (ncu: 549 -> 486 us at B=32).  dK/dV kernel: two threads per row
// (w2..w9) -- with four it needs 16-column steps at the 96-register cap and came out 12 % slower.
constexpr int ABW_THREADS = 576;
constexpr int ABW_SIMT = 512;
constexpr int ABW_KV_THREADS = 320;
constexpr int ABW_KV_SIMT = 256;

struct AttnBwdParams {
    int B, F, J, C, H;
    int NK;                     // round_up(F, 16)
    float scale;                // d^-1/2
    float scale_log2e;
    const __nv_bfloat16* O;     // [M, C]   forward attention output (for delta)          (q kernel)
    const __nv_bfloat16* dO;    // [M, C]   gradient w.r.t. the attention output            (q kernel: row reads)
    float* lse2;                // [B*J*H*F] log2-sum-exp of the scaled scores per query   (q kernel writes, kv reads)
    float* delta;               // [B*J*H*F]
    __nv_bfloat16* dqkv;        // [M, 3C]  output gradient (q part: q kernel; k, v parts: kv kernel)
};

template <int HD>
struct AttnBwdCfg {
    static constexpr int SWZ = HD * 2;
    static constexpr uint32_t LAYOUT = (SWZ == 128) ? 2u : 4u;
    static constexpr int TILE = ATT_BM * SWZ;                 // a 128-row operand tile
    static constexpr int SEQ = ATT_MAXK * SWZ;                // a whole-sequence operand (<= 256 rows)
    static constexpr int OFF_A = 0;                           // q kernel: Q tile   | kv kernel: K tile
    static constexpr int OFF_B = TILE;                        // q kernel: dO tile  | kv kernel: V tile
    static constexpr int OFF_C = 2 * TILE;                    // q kernel: K (seq)  | kv kernel: Q (seq)
    static constexpr int OFF_D = 2 * TILE + SEQ;              // q kernel: V (seq)  | kv kernel: dO (seq)
    static constexpr int OFF_E = 2 * TILE + 2 * SEQ;          // q kernel: O tile (row-wise delta = dO . O, read by the SIMT threads)
    static constexpr int OFF_BAR = 3 * TILE + 2 * SEQ;
    static constexpr int OFF_VEC = OFF_BAR + 128;             // floats: q kernel max/sum/delta [3][4][128] | kv kernel lse2[256], delta[256]
    static constexpr int SMEM_BYTES = OFF_VEC + 3 * 4 * 128 * 4 + 1024;
};

// ------------------------------------------------------------------------------------------------- dQ kernel
template <int HD, bool PACK = false>
__global__ void __launch_bounds__(ABW_THREADS, 1)
attn_bwd_q_kernel(const __grid_constant__ CUtensorMap tmQKV_t,   // qkv 5-D, box (HD, 1, 128, 1, 1)
                  const __grid_constant__ CUtensorMap tmQKV_s,   // qkv 5-D, box (HD, 1, NK , 1, 1)
                  const __grid_constant__ CUtensorMap tmDO_t,    // dO  5-D, box (HD, 1, 128, 1, 1)
                  const __grid_constant__ CUtensorMap tmO_t,     // O   5-D, box (HD, 1, 128, 1, 1)
                  const AttnBwdParams p) {
    using Cfg = AttnBwdCfg<HD>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* kv_full = bars + 0;
    uint64_t* kv_empty = bars + 1;
    uint64_t* t_full = bars + 2;      // Q and dO tile landed
    uint64_t* t_empty = bars + 3;
    uint64_t* sd_full = bars + 4;     // S and dP accumulators ready
    uint64_t* ds_full = bars + 5;     // dS written (256 threads)
    uint64_t* dq_full = bars + 6;
    uint64_t* dq_empty = bars + 7;    // 256 threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    float* red = reinterpret_cast<float*>(smem + Cfg::OFF_VEC);   // [2][128] max, then [2][128] sum

    const int warp = warp_uniform(threadIdx.x >> 5);      // uniform role dispatch (see ptx.cuh elect_one)
    const int lane = threadIdx.x & 31;
    const int nseq = p.B * p.J;
    const int num_prob = PACK ? ((nseq + 3) / 4) * p.H : nseq * p.H;
    const int num_qt = PACK ? 1 : (p.F + ATT_BM - 1) / ATT_BM;
    const int NKe = PACK ? ATT_BM : p.NK;                    // key columns of the score tile
    const uint32_t seq_bytes = static_cast<uint32_t>(NKe) * Cfg::SWZ;
    constexpr uint32_t SLAB = 32 * Cfg::SWZ;                 // one 32-row slab of a packed tile

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV_t); tma_prefetch_desc(&tmQKV_s); tma_prefetch_desc(&tmDO_t); tma_prefetch_desc(&tmO_t);
        mbar_init(kv_full, 1);  mbar_init(kv_empty, 1);
        mbar_init(t_full, 1);   mbar_init(t_empty, 1 + ABW_SIMT);   // MMA commit + the SIMT readers of the dO / O tiles
        mbar_init(sd_full, 1);  mbar_init(ds_full, ABW_SIMT);
        mbar_init(dq_full, 1);  mbar_init(dq_empty, ABW_SIMT);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base;             // [0,256): S, then dS (packed bf16) in place
    const uint32_t tmem_dP = tmem_base + 256;      // [256,512): dP, then dQ accumulator in its first HD columns

    if (warp == 0) {
        if (elect_one()) {
            uint32_t t_it = 0;
            int ip = 0;
            for (int prob = blockIdx.x; prob < num_prob; prob += gridDim.x, ++ip) {
                const int h = prob % p.H, j = (prob / p.H) % p.J, b = prob / (p.H * p.J);
                mbar_wait(kv_empty, (ip & 1) ^ 1);
                mbar_arrive_expect_tx(kv_full, 2 * seq_bytes);
                if (PACK) {
                    // four sequences, one 32-row box each (sequences past the end re-load the last one: finite data,
                    // their rows are never stored)
                    for (int s4 = 0; s4 < 4; ++s4) {
                        int sq = (prob / p.H) * 4 + s4;
                        if (sq >= nseq) sq = nseq - 1;
                        tma_load_5d(smem + Cfg::OFF_C + s4 * SLAB, &tmQKV_s, kv_full, p.C + h * HD, sq % p.J, 0, sq / p.J, 0);
                        tma_load_5d(smem + Cfg::OFF_D + s4 * SLAB, &tmQKV_s, kv_full, 2 * p.C + h * HD, sq % p.J, 0, sq / p.J, 0);
                    }
                    mbar_wait(t_empty, (t_it & 1) ^ 1);
                    mbar_arrive_expect_tx(t_full, 3 * Cfg::TILE);
                    for (int s4 = 0; s4 < 4; ++s4) {
                        int sq = (prob / p.H) * 4 + s4;
                        if (sq >= nseq) sq = nseq - 1;
                        tma_load_5d(smem + Cfg::OFF_A + s4 * SLAB, &tmQKV_t, t_full, h * HD, sq % p.J, 0, sq / p.J, 0);
                        tma_load_5d(smem + Cfg::OFF_B + s4 * SLAB, &tmDO_t, t_full, h * HD, sq % p.J, 0, sq / p.J, 0);
                        tma_load_5d(smem + Cfg::OFF_E + s4 * SLAB, &tmO_t, t_full, h * HD, sq % p.J, 0, sq / p.J, 0);
                    }
                    ++t_it;
                    continue;
                }
                tma_load_5d(smem + Cfg::OFF_C, &tmQKV_s, kv_full, p.C + h * HD, j, 0, b, 0);       // K
                tma_load_5d(smem + Cfg::OFF_D, &tmQKV_s, kv_full, 2 * p.C + h * HD, j, 0, b, 0);   // V
                for (int qt = 0; qt < num_qt; ++qt, ++t_it) {
                    mbar_wait(t_empty, (t_it & 1) ^ 1);
                    mbar_arrive_expect_tx(t_full, 3 * Cfg::TILE);
                    tma_load_5d(smem + Cfg::OFF_A, &tmQKV_t, t_full, h * HD, j, qt * ATT_BM, b, 0);   // Q tile
                    tma_load_5d(smem + Cfg::OFF_B, &tmDO_t, t_full, h * HD, j, qt * ATT_BM, b, 0);    // dO tile
                    tma_load_5d(smem + Cfg::OFF_E, &tmO_t, t_full, h * HD, j, qt * ATT_BM, b, 0);     // O tile
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc_s = umma_idesc_bf16(ATT_BM, NKe, 0, 0);
        const uint32_t idesc_q = umma_idesc_bf16(ATT_BM, HD, 0, 1);     // dQ = dS K: B (=K) MN-major
        const uint32_t sQ = smem_u32(smem + Cfg::OFF_A), sDO = smem_u32(smem + Cfg::OFF_B);
        const uint32_t sK = smem_u32(smem + Cfg::OFF_C), sV = smem_u32(smem + Cfg::OFF_D);
        uint32_t t_it = 0;
        int ip = 0;
        for (int prob = blockIdx.x; prob < num_prob; prob += gridDim.x, ++ip) {
            mbar_wait(kv_full, ip & 1);
            for (int qt = 0; qt < num_qt; ++qt, ++t_it) {
                const uint32_t ph = t_it & 1;
                mbar_wait(t_full, ph);
                mbar_wait(dq_empty, ph ^ 1);          // previous tile's dQ (aliases dP) has been read out
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t dq_ = umma_smem_desc(sQ, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                    const uint64_t ddo = umma_smem_desc(sDO, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                    const uint64_t dk = umma_smem_desc(sK, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                    const uint64_t dv = umma_smem_desc(sV, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks) {
                        const uint64_t koff = static_cast<uint64_t>((ks * 32) >> 4);
                        umma_ss(tmem_S, dq_ + koff, dk + koff, idesc_s, ks != 0);      // S  = Q  K^T
                    }
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks) {
                        const uint64_t koff = static_cast<uint64_t>((ks * 32) >> 4);
                        umma_ss(tmem_dP, ddo + koff, dv + koff, idesc_s, ks != 0);     // dP = dO V^T
                    }
                    tc_commit(sd_full);
                    tc_commit(t_empty);
                }
                __syncwarp();
                mbar_wait(ds_full, ph);
                tc_fence_after();
                if (elect_one()) {
                    const int nks = NKe / 16;
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint32_t a = tmem_S + 32 * (ks >> 1) + 8 * (ks & 1);      // packed bf16 dS
                        const uint32_t koff = static_cast<uint32_t>(ks) * 16 * Cfg::SWZ;
                        const uint64_t kmn = umma_smem_desc(sK + koff, 8 * Cfg::SWZ, 8 * Cfg::SWZ, Cfg::LAYOUT);
                        umma_ts(tmem_dP, a, kmn, idesc_q, ks != 0);                     // dQ = dS K
                    }
                    tc_commit(dq_full);
                    if (qt == num_qt - 1) tc_commit(kv_empty);
                }
                __syncwarp();
            }
        }
    } else {
        const int quad = warp & 3;
        const int part = (warp - 2) >> 2;                 // 0..3: the four threads of a row split its key columns
        const int r_in_tile = quad * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const int nch = (NKe + 31) / 32;
        // unpacked: part p owns the 32-column chunks [2p, 2p+2); packed: only the diagonal chunk (= quad) is live, part 0
        // owns it and parts 1..3 zero-fill one off-diagonal chunk each
        const int ch_lo = PACK ? (part == 0 ? quad : 0) : (part * 2 < nch ? part * 2 : nch);
        const int ch_hi = PACK ? (part == 0 ? quad + 1 : 0) : (part * 2 + 2 < nch ? part * 2 + 2 : nch);
        const float sl2 = p.scale_log2e;
        float* red_max = red;                             // [4][128]
        float* red_sum = red + 512;
        float* red_del = red + 1024;
        uint32_t t_it = 0;
        for (int prob = blockIdx.x; prob < num_prob; prob += gridDim.x) {
            int h = prob % p.H, j = (prob / p.H) % p.J, b = prob / (p.H * p.J);
            for (int qt = 0; qt < num_qt; ++qt, ++t_it) {
                const uint32_t ph = t_it & 1;
                int tq = qt * ATT_BM + r_in_tile;
                bool ok = tq < p.F;
                if (PACK) {
                    const int sq = (prob / p.H) * 4 + quad;          // my slab's sequence
                    tq = lane;
                    ok = lane < p.F && sq < nseq;
                    j = ok ? sq % p.J : 0;
                    b = ok ? sq / p.J : 0;
                }
                const size_t tok = (static_cast<size_t>(b) * p.F + (ok ? tq : 0)) * p.J + j;
                // delta = dO . O of my row, from the TMA-staged (swizzled) tiles: each of the four threads takes HD/4 columns.
                // (Reading the rows from global memory here -- 128 scattered 128-byte rows per warp instruction -- was the
                // top stall of this kernel.)
                float dpart = 0.f;
                mbar_wait(t_full, ph);
                {
                    constexpr int NCHK = HD / 32;                    // 16-byte chunks per thread: 2 (HD = 64) or 1
                    const uint32_t sw = (Cfg::SWZ == 128) ? static_cast<uint32_t>(r_in_tile & 7)
                                                          : static_cast<uint32_t>((r_in_tile >> 1) & 3);
                    const uint8_t* so = smem + Cfg::OFF_E + r_in_tile * Cfg::SWZ;
                    const uint8_t* sg = smem + Cfg::OFF_B + r_in_tile * Cfg::SWZ;
#pragma unroll
                    for (int i = 0; i < NCHK; ++i) {
                        const uint32_t chunk = (static_cast<uint32_t>(part * NCHK + i) ^ sw) << 4;
                        const uint4 a = *reinterpret_cast<const uint4*>(so + chunk);
                        const uint4 g = *reinterpret_cast<const uint4*>(sg + chunk);
                        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            dpart = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(gw[e] << 16), dpart);
                            dpart = fmaf(__uint_as_float(aw[e] & 0xffff0000u), __uint_as_float(gw[e] & 0xffff0000u), dpart);
                        }
                    }
                }
                mbar_arrive(t_empty);                                // my reads of the dO / O tiles are done
                red_del[part * 128 + r_in_tile] = dpart;
                mbar_wait(sd_full, ph);
                tc_fence_after();
                float mx = -INFINITY;
                for (int ch = ch_lo; ch < ch_hi; ++ch) {
                    uint32_t r[32];
                    tmem_ld32(tmem_S + lane_off + ch * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if ((PACK ? i : ch * 32 + i) < p.F) mx = fmaxf(mx, __uint_as_float(r[i]));
                }
                red_max[part * 128 + r_in_tile] = mx;
                named_bar_sync(1, ABW_SIMT);
                mx = fmaxf(fmaxf(red_max[r_in_tile], red_max[128 + r_in_tile]), fmaxf(red_max[256 + r_in_tile], red_max[384 + r_in_tile]));
                const float delta = (red_del[r_in_tile] + red_del[128 + r_in_tile]) + (red_del[256 + r_in_tile] + red_del[384 + r_in_tile]);
                const float mxs = mx * sl2;
                float sum = 0.f;
                for (int ch = ch_lo; ch < ch_hi; ++ch) {
                    uint32_t r[32];
                    tmem_ld32(tmem_S + lane_off + ch * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if ((PACK ? i : ch * 32 + i) < p.F) sum += ex2_approx(fmaf(__uint_as_float(r[i]), sl2, -mxs));
                }
                red_sum[part * 128 + r_in_tile] = sum;
                named_bar_sync(1, ABW_SIMT);
                sum = (red_sum[r_in_tile] + red_sum[128 + r_in_tile]) + (red_sum[256 + r_in_tile] + red_sum[384 + r_in_tile]);
                const float lse2 = mxs + log2f(sum);                 // P = 2^(s*c*log2e - lse2)
                if (ok && part == 0) {
                    const size_t si = PACK ? static_cast<size_t>(prob) * ATT_BM + r_in_tile
                                           : static_cast<size_t>(prob) * p.F + tq;
                    p.lse2[si] = lse2;
                    p.delta[si] = delta;
                }
                // dS = P (dP - delta) * scale  -> packed bf16 over S, 16 source columns at a time (register budget: 96)
                for (int ch = ch_lo; ch < ch_hi; ++ch) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t sv[16], gv[16];
                        tmem_ld16(tmem_S + lane_off + ch * 32 + hh * 16, sv);
                        tmem_ld16(tmem_dP + lane_off + ch * 32 + hh * 16, gv);
                        tmem_ld_wait();
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int c = hh * 16 + 2 * i;
                            float d0 = 0.f, d1 = 0.f;
                            if ((PACK ? c : ch * 32 + c) < p.F) {
                                const float pv = ex2_approx(fmaf(__uint_as_float(sv[2 * i]), sl2, -lse2));
                                d0 = pv * (__uint_as_float(gv[2 * i]) - delta) * p.scale;
                            }
                            if ((PACK ? c + 1 : ch * 32 + c + 1) < p.F) {
                                const float pv = ex2_approx(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, -lse2));
                                d1 = pv * (__uint_as_float(gv[2 * i + 1]) - delta) * p.scale;
                            }
                            const __nv_bfloat162 t2 = __floats2bfloat162_rn(d0, d1);
                            pk[i] = *reinterpret_cast<const uint32_t*>(&t2);
                        }
                        tmem_st8(tmem_S + lane_off + ch * 32 + hh * 8, pk);
                    }
                }
                if (PACK && part != 0) {
                    uint32_t z[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) z[i] = 0u;
                    tmem_st16(tmem_S + lane_off + ((quad + part) & 3) * 32, z);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(ds_full);
                // dQ tile: 16 columns per thread
                mbar_wait(dq_full, ph);
                tc_fence_after();
                if (HD == 64 || part < 2) {
                    const int c0 = part * 16;
                    uint32_t r[16];
                    tmem_ld16(tmem_dP + lane_off + c0, r);
                    tmem_ld_wait();
                    if (ok) {
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
                            pk[i] = *reinterpret_cast<const uint32_t*>(&t2);
                        }
                        uint4* d4 = reinterpret_cast<uint4*>(p.dqkv + tok * (3 * p.C) + h * HD + c0);
                        d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
                tc_fence_before();
                mbar_arrive(dq_empty);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------- dK / dV kernel
template <int HD, bool PACK = false>
__global__ void __launch_bounds__(ABW_KV_THREADS, 1)
attn_bwd_kv_kernel(const __grid_constant__ CUtensorMap tmQKV_t,   // qkv 5-D, box (HD, 1, 128, 1, 1)
                   const __grid_constant__ CUtensorMap tmQKV_s,   // qkv 5-D, box (HD, 1, NK , 1, 1)
                   const __grid_constant__ CUtensorMap tmDO_s,    // dO  5-D, box (HD, 1, NK , 1, 1)
                   const AttnBwdParams p) {
    using Cfg = AttnBwdCfg<HD>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* seq_full = bars + 0;    // Q and dO of the whole sequence
    uint64_t* seq_empty = bars + 1;
    uint64_t* t_full = bars + 2;      // K and V tile
    uint64_t* t_empty = bars + 3;
    uint64_t* sd_full = bars + 4;
    uint64_t* ps_full = bars + 5;     // P^T and dS^T written (256 threads)
    uint64_t* g_full = bars + 6;      // dV, dK accumulators ready
    uint64_t* g_empty = bars + 7;     // 256 threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    float* v_lse = reinterpret_cast<float*>(smem + Cfg::OFF_VEC);   // [256]
    float* v_delta = v_lse + 256;                                   // [256]

    const int warp = warp_uniform(threadIdx.x >> 5);      // uniform role dispatch (see ptx.cuh elect_one)
    const int lane = threadIdx.x & 31;
    const int nseq = p.B * p.J;
    const int num_prob = PACK ? ((nseq + 3) / 4) * p.H : nseq * p.H;
    const int num_kt = PACK ? 1 : (p.F + ATT_BM - 1) / ATT_BM;
    const int NKe = PACK ? ATT_BM : p.NK;
    const uint32_t seq_bytes = static_cast<uint32_t>(NKe) * Cfg::SWZ;
    constexpr uint32_t SLAB = 32 * Cfg::SWZ;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV_t); tma_prefetch_desc(&tmQKV_s); tma_prefetch_desc(&tmDO_s);
        mbar_init(seq_full, 1); mbar_init(seq_empty, 1);
        mbar_init(t_full, 1);   mbar_init(t_empty, 1);
        mbar_init(sd_full, 1);  mbar_init(ps_full, ABW_KV_SIMT);
        mbar_init(g_full, 1);   mbar_init(g_empty, ABW_KV_SIMT);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base;             // [0,256): S^T, then per 32-column chunk: 16 cols P^T | 16 cols dS^T
    const uint32_t tmem_dP = tmem_base + 256;      // [256,512): dP^T, then dV at [256,256+HD), dK at [256+HD,256+2HD)

    if (warp == 0) {
        if (elect_one()) {
            uint32_t t_it = 0;
            int ip = 0;
            for (int prob = blockIdx.x; prob < num_prob; prob += gridDim.x, ++ip) {
                const int h = prob % p.H, j = (prob / p.H) % p.J, b = prob / (p.H * p.J);
                mbar_wait(seq_empty, (ip & 1) ^ 1);
                mbar_arrive_expect_tx(seq_full, 2 * seq_bytes);
                if (PACK) {
                    for (int s4 = 0; s4 < 4; ++s4) {
                        int sq = (prob / p.H) * 4 + s4;
                        if (sq >= nseq) sq = nseq - 1;
                        tma_load_5d(smem + Cfg::OFF_C + s4 * SLAB, &tmQKV_s, seq_full, h * HD, sq % p.J, 0, sq / p.J, 0);
                        tma_load_5d(smem + Cfg::OFF_D + s4 * SLAB, &tmDO_s, seq_full, h * HD, sq % p.J, 0, sq / p.J, 0);
                    }
                    mbar_wait(t_empty, (t_it & 1) ^ 1);
                    mbar_arrive_expect_tx(t_full, 2 * Cfg::TILE);
                    for (int s4 = 0; s4 < 4; ++s4) {
                        int sq = (prob / p.H) * 4 + s4;
                        if (sq >= nseq) sq = nseq - 1;
                        tma_load_5d(smem + Cfg::OFF_A + s4 * SLAB, &tmQKV_t, t_full, p.C + h * HD, sq % p.J, 0, sq / p.J, 0);
                        tma_load_5d(smem + Cfg::OFF_B + s4 * SLAB, &tmQKV_t, t_full, 2 * p.C + h * HD, sq % p.J, 0, sq / p.J, 0);
                    }
                    ++t_it;
                    continue;
                }
                tma_load_5d(smem + Cfg::OFF_C, &tmQKV_s, seq_full, h * HD, j, 0, b, 0);     // Q (all queries)
                tma_load_5d(smem + Cfg::OFF_D, &tmDO_s, seq_full, h * HD, j, 0, b, 0);      // dO (all queries)
                for (int kt = 0; kt < num_kt; ++kt, ++t_it) {
                    mbar_wait(t_empty, (t_it & 1) ^ 1);
                    mbar_arrive_expect_tx(t_full, 2 * Cfg::TILE);
                    tma_load_5d(smem + Cfg::OFF_A, &tmQKV_t, t_full, p.C + h * HD, j, kt * ATT_BM, b, 0);       // K tile
                    tma_load_5d(smem + Cfg::OFF_B, &tmQKV_t, t_full, 2 * p.C + h * HD, j, kt * ATT_BM, b, 0);   // V tile
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc_s = umma_idesc_bf16(ATT_BM, NKe, 0, 0);
        const uint32_t idesc_g = umma_idesc_bf16(ATT_BM, HD, 0, 1);
        const uint32_t sK = smem_u32(smem + Cfg::OFF_A), sV = smem_u32(smem + Cfg::OFF_B);
        const uint32_t sQ = smem_u32(smem + Cfg::OFF_C), sDO = smem_u32(smem + Cfg::OFF_D);
        uint32_t t_it = 0;
        int ip = 0;
        for (int prob = blockIdx.x; prob < num_prob; prob += gridDim.x, ++ip) {
            mbar_wait(seq_full, ip & 1);
            for (int kt = 0; kt < num_kt; ++kt, ++t_it) {
                const uint32_t ph = t_it & 1;
                mbar_wait(t_full, ph);
                mbar_wait(g_empty, ph ^ 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t dk = umma_smem_desc(sK, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                    const uint64_t dv = umma_smem_desc(sV, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                    const uint64_t dq_ = umma_smem_desc(sQ, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                    const uint64_t ddo = umma_smem_desc(sDO, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks) {
                        const uint64_t koff = static_cast<uint64_t>((ks * 32) >> 4);
                        umma_ss(tmem_S, dk + koff, dq_ + koff, idesc_s, ks != 0);       // S^T  = K Q^T
                    }
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks) {
                        const uint64_t koff = static_cast<uint64_t>((ks * 32) >> 4);
                        umma_ss(tmem_dP, dv + koff, ddo + koff, idesc_s, ks != 0);      // dP^T = V dO^T
                    }
                    tc_commit(sd_full);
                    tc_commit(t_empty);
                }
                __syncwarp();
                mbar_wait(ps_full, ph);
                tc_fence_after();
                if (elect_one()) {
                    const int nks = NKe / 16;
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint32_t a_p = tmem_S + 32 * (ks >> 1) + 8 * (ks & 1);     // P^T  : first 16 columns of the chunk
                        const uint32_t a_ds = a_p + 16;                                  // dS^T : last 16 columns
                        const uint32_t roff = static_cast<uint32_t>(ks) * 16 * Cfg::SWZ;
                        const uint64_t do_mn = umma_smem_desc(sDO + roff, 8 * Cfg::SWZ, 8 * Cfg::SWZ, Cfg::LAYOUT);
                        const uint64_t q_mn = umma_smem_desc(sQ + roff, 8 * Cfg::SWZ, 8 * Cfg::SWZ, Cfg::LAYOUT);
                        umma_ts(tmem_dP, a_p, do_mn, idesc_g, ks != 0);                  // dV = P^T  dO
                        umma_ts(tmem_dP + HD, a_ds, q_mn, idesc_g, ks != 0);             // dK = dS^T Q
                    }
                    tc_commit(g_full);
                    if (kt == num_kt - 1) tc_commit(seq_empty);
                }
                __syncwarp();
            }
        }
    } else {
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int r_in_tile = quad * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const int nch = (NKe + 31) / 32;
        const int ch_lo = PACK ? (half == 0 ? quad : 0) : (half * 4 < nch ? half * 4 : nch);
        const int ch_hi = PACK ? (half == 0 ? quad + 1 : 0) : ((half * 4 + 4 < nch) ? half * 4 + 4 : nch);
        const float sl2 = p.scale_log2e;
        const int sid = threadIdx.x - 64;             // 0..255
        uint32_t t_it = 0;
        for (int prob = blockIdx.x; prob < num_prob; prob += gridDim.x) {
            int h = prob % p.H, j = (prob / p.H) % p.J, b = prob / (p.H * p.J);
            // per-query statistics of this sequence -> smem (queries >= F get p = 0)
            named_bar_sync(1, ABW_KV_SIMT);               // previous problem's readers are done
            if (PACK) {
                const bool qok = sid < ATT_BM && (sid & 31) < p.F && (prob / p.H) * 4 + (sid >> 5) < nseq;
                const size_t si = static_cast<size_t>(prob) * ATT_BM;
                v_lse[sid] = qok ? p.lse2[si + sid] : 3.0e38f;
                v_delta[sid] = qok ? p.delta[si + sid] : 0.f;
            } else {
                const size_t si = static_cast<size_t>(prob) * p.F;
                v_lse[sid] = sid < p.F ? p.lse2[si + sid] : 3.0e38f;
                v_delta[sid] = sid < p.F ? p.delta[si + sid] : 0.f;
            }
            named_bar_sync(1, ABW_KV_SIMT);
            for (int kt = 0; kt < num_kt; ++kt, ++t_it) {
                const uint32_t ph = t_it & 1;
                mbar_wait(sd_full, ph);
                tc_fence_after();
                for (int ch = ch_lo; ch < ch_hi; ++ch) {
                    uint32_t s[32], g[32];
                    tmem_ld32(tmem_S + lane_off + ch * 32, s);
                    tmem_ld32(tmem_dP + lane_off + ch * 32, g);
                    tmem_ld_wait();
                    uint32_t pp[16], pd[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int q0 = ch * 32 + 2 * i;
                        const float p0 = ex2_approx(fmaf(__uint_as_float(s[2 * i]), sl2, -v_lse[q0]));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(s[2 * i + 1]), sl2, -v_lse[q0 + 1]));
                        const float d0 = p0 * (__uint_as_float(g[2 * i]) - v_delta[q0]) * p.scale;
                        const float d1 = p1 * (__uint_as_float(g[2 * i + 1]) - v_delta[q0 + 1]) * p.scale;
                        const __nv_bfloat162 a2 = __floats2bfloat162_rn(p0, p1);
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(d0, d1);
                        pp[i] = *reinterpret_cast<const uint32_t*>(&a2);
                        pd[i] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
                    tmem_st16(tmem_S + lane_off + ch * 32, pp);
                    tmem_st16(tmem_S + lane_off + ch * 32 + 16, pd);
                }
                if (PACK && half == 1) {
                    uint32_t z[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch)
                        if (ch != quad) {
                            tmem_st16(tmem_S + lane_off + ch * 32, z);
                            tmem_st16(tmem_S + lane_off + ch * 32 + 16, z);
                        }
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(ps_full);
                mbar_wait(g_full, ph);
                tc_fence_after();
                int tk = kt * ATT_BM + r_in_tile;
                bool ok = tk < p.F;
                if (PACK) {
                    const int sq = (prob / p.H) * 4 + quad;
                    tk = lane;
                    ok = lane < p.F && sq < nseq;
                    j = ok ? sq % p.J : 0;
                    b = ok ? sq / p.J : 0;
                }
                const size_t tok = (static_cast<size_t>(b) * p.F + (ok ? tk : 0)) * p.J + j;
                // half 0 drains dV, half 1 drains dK (HD columns each)
                {
                    const uint32_t src = tmem_dP + half * HD;
                    const int part = half == 0 ? 2 : 1;              // v part / k part of the qkv gradient
#pragma unroll
                    for (int c0 = 0; c0 < HD; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32(src + lane_off + c0, r);
                        tmem_ld_wait();
                        if (ok) {
                            uint32_t pk[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
                                pk[i] = *reinterpret_cast<const uint32_t*>(&t2);
                            }
                            uint4* d4 = reinterpret_cast<uint4*>(p.dqkv + tok * (3 * p.C) + part * p.C + h * HD + c0);
#pragma unroll
                            for (int i = 0; i < 4; ++i) d4[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(g_empty);
            }
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
