// Temporal multi-head attention on tcgen05 (DSTformer.py:188-200, `Attention.forward_temporal`).
//
// One problem = one (batch b, joint j, head h): softmax(Q K^T * d^-1/2) V over the F <= 256 frames of
// that joint.  The reference materialises a (B,H,J,F,F) fp32 score tensor three times in HBM plus
// permute copies; here scores never leave the SM:
//   * Q/K/V tiles are gathered straight out of the (B*F*J, 3C) token-major QKV buffer by a 5-D TMA
//     tensor map (col, joint, frame, batch, plane) - the temporal "permute" is free.
//   * S = Q K^T  : tcgen05.mma, M=128 query frames x N=NK keys, fp32 accumulator in TMEM (NK <= 256 columns)
//   * softmax    : 512 threads, 4 per query row (= TMEM lane), each owning a quarter of the keys; probabilities are written back
//                  IN PLACE over S as packed bf16 hi / lo planes (32 score columns -> 16 hi + 16 lo columns)
//   * O = P V    : tcgen05.mma with the A operand read from TMEM (P) and V as an MN-major smem operand
//   * epilogue   : O / rowsum -> bf16 hi/lo planes of the (M, C) attention output (token-major again)
// BF16x3 (PASSES == 3): Qh Kl + Ql Kh + Qh Kh and Ph Vl + Pl Vh + Ph Vh, fp32 accumulate.
// Persistent: grid = #SMs, each CTA loops over problems; K / V / Q buffers have their own full/empty
// barriers so the TMA warp prefetches the next problem's K and Q while the current P V product runs.
#pragma once
#include "ptx.cuh"

namespace mb {

constexpr int ATT_THREADS = 320;      // (attn_s_tc / attn_t_tc2) TMA warp, MMA warp, 8 softmax/output warps
constexpr int ATT_SM_THREADS = 256;
constexpr int ATT_T_GROUPS = 4;       // temporal kernel: 4 threads per query row, each owning a quarter of the keys
constexpr int ATT_T_SM_THREADS = ATT_T_GROUPS * 128;      // 16 softmax/output warps: the softmax is latency-bound,
constexpr int ATT_T_THREADS = 64 + ATT_T_SM_THREADS;      // 4 warps per scheduler hide MUFU / tcgen05.ld latency
constexpr int ATT_BM = 128;      // query rows per tile
constexpr int ATT_MAXK = 256;    // max keys (frames)

struct AttnTParams {
    int B, F, J, C, H;
    int NK;                  // round_up(F, 16) keys per problem
    float scale_log2e;       // d^-1/2 * log2(e)
    __nv_bfloat16* out_hi;   // [M, C]
    __nv_bfloat16* out_lo;   // may be null (PASSES == 1)
    int out_f16c;            // != 0: out_hi is an F16C row buffer [M][C] (ptx.cuh) for an F16C-mode projection GEMM
};

// 16 consecutive output values of one row -> F16C block pieces at column `col` (multiple of 16) of row buffer `rowp`
__device__ __forceinline__ void store16_f16c(uint8_t* rowp, int col, const float (&x)[16]) {
    uint8_t* blk = rowp + static_cast<size_t>(col >> 5) * 128;
    const int e = col & 31;
    uint32_t h[8], l[4], g[4];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float xv[8] = {x[8 * q], x[8 * q + 1], x[8 * q + 2], x[8 * q + 3], x[8 * q + 4], x[8 * q + 5], x[8 * q + 6], x[8 * q + 7]};
        uint32_t h4[4], l2[2], g2[2];
        split8_f16c(xv, h4, l2, g2);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[4 * q + i] = h4[i];
        l[2 * q] = l2[0]; l[2 * q + 1] = l2[1];
        g[2 * q] = g2[0]; g[2 * q + 1] = g2[1];
    }
    uint4* hp = reinterpret_cast<uint4*>(blk + 2 * e);
    hp[0] = make_uint4(h[0], h[1], h[2], h[3]);
    hp[1] = make_uint4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<uint4*>(blk + 64 + e) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4*>(blk + 96 + e) = make_uint4(g[0], g[1], g[2], g[3]);
}

template <int HD, int PASSES>
struct AttnCfg {
    static constexpr int SWZ = HD * 2;                            // 128 B (d=64) or 64 B (d=32)
    static constexpr uint32_t LAYOUT = (SWZ == 128) ? 2u : 4u;
    static constexpr int PLANES = (PASSES == 3) ? 2 : 1;
    static constexpr int Q_PLANE = ATT_BM * SWZ;
    static constexpr int Q_BYTES = PLANES * Q_PLANE;
    static constexpr int KV_MAX_BYTES = PLANES * ATT_MAXK * SWZ;
    static constexpr int OFF_K = 0;
    static constexpr int OFF_V = KV_MAX_BYTES;
    static constexpr int OFF_Q = 2 * KV_MAX_BYTES;
    static constexpr int OFF_BAR = OFF_Q + 2 * Q_BYTES;
    static constexpr int OFF_RED = OFF_BAR + 256;                 // float red_max[4][128], red_sum[4][128]
    static constexpr int SMEM_BYTES = OFF_RED + 2 * ATT_T_GROUPS * 128 * 4 + 1024;
};

template <int HD, int PASSES>
__global__ void __launch_bounds__(ATT_T_THREADS, 1)
attn_t_tc_kernel(const __grid_constant__ CUtensorMap tmQ,    // box (HD, 1, 128, 1, PLANES)
                 const __grid_constant__ CUtensorMap tmKV,   // box (HD, 1, NK , 1, PLANES)
                 const AttnTParams p) {
    using Cfg = AttnCfg<HD, PASSES>;
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
    uint64_t* p_full = bars + 9;
    uint64_t* o_full = bars + 10;
    uint64_t* o_empty = bars + 11;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = warp_uniform(threadIdx.x >> 5);      // uniform role dispatch (see ptx.cuh elect_one)
    const int lane = threadIdx.x & 31;
    const int num_prob = p.B * p.J * p.H;
    const int num_qt = (p.F + ATT_BM - 1) / ATT_BM;
    const int kv_plane = p.NK * Cfg::SWZ;                    // bytes per K (or V) plane in smem
    const uint32_t kv_bytes = Cfg::PLANES * kv_plane;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        mbar_init(k_full, 1);  mbar_init(k_empty, 1);
        mbar_init(v_full, 1);  mbar_init(v_empty, 1);
        mbar_init(&q_full[0], 1);  mbar_init(&q_full[1], 1);
        mbar_init(&q_empty[0], 1); mbar_init(&q_empty[1], 1);
        mbar_init(s_full, 1);
        mbar_init(p_full, ATT_T_SM_THREADS);
        mbar_init(o_full, 1);
        mbar_init(o_empty, ATT_T_SM_THREADS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base;               // columns [0, 256): S, then P (in place)
    const uint32_t tmem_O = tmem_base + ATT_MAXK;    // columns [256, 256+HD)

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (elect_one()) {
            uint32_t kv_it = 0, q_it = 0;
            for (int prob = blockIdx.x; prob < num_prob; prob += gridDim.x) {
                const int h = prob % p.H, j = (prob / p.H) % p.J, b = prob / (p.H * p.J);
                const uint32_t kv_ph = kv_it & 1;
                mbar_wait(k_empty, kv_ph ^ 1);
                mbar_arrive_expect_tx(k_full, kv_bytes);
                tma_load_5d(smem + Cfg::OFF_K, &tmKV, k_full, p.C + h * HD, j, 0, b, 0);
                for (int qt = 0; qt < num_qt; ++qt) {
                    const int qs = q_it & 1;
                    const uint32_t q_ph = (q_it >> 1) & 1;
                    mbar_wait(&q_empty[qs], q_ph ^ 1);
                    mbar_arrive_expect_tx(&q_full[qs], Cfg::Q_BYTES);
                    tma_load_5d(smem + Cfg::OFF_Q + qs * Cfg::Q_BYTES, &tmQ, &q_full[qs], h * HD, j, qt * ATT_BM, b, 0);
                    ++q_it;
                    if (qt == 0) {
                        mbar_wait(v_empty, kv_ph ^ 1);
                        mbar_arrive_expect_tx(v_full, kv_bytes);
                        tma_load_5d(smem + Cfg::OFF_V, &tmKV, v_full, 2 * p.C + h * HD, j, 0, b, 0);
                    }
                }
                ++kv_it;
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        const uint32_t idesc_s = umma_idesc_bf16(ATT_BM, p.NK, 0, 0);     // S = Q K^T : both K-major
        const uint32_t idesc_o = umma_idesc_bf16(ATT_BM, HD, 0, 1);       // O = P V   : B (=V) is MN-major
        const uint32_t sK = smem_u32(smem + Cfg::OFF_K);
        const uint32_t sV = smem_u32(smem + Cfg::OFF_V);
        uint32_t kv_it = 0, q_it = 0, t_it = 0;
        for (int prob = blockIdx.x; prob < num_prob; prob += gridDim.x) {
            const uint32_t kv_ph = kv_it & 1;
            mbar_wait(k_full, kv_ph);
            for (int qt = 0; qt < num_qt; ++qt) {
                const int qs = q_it & 1;
                const uint32_t q_ph = (q_it >> 1) & 1;
                const uint32_t t_ph = t_it & 1;
                mbar_wait(&q_full[qs], q_ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sQ = smem_u32(smem + Cfg::OFF_Q + qs * Cfg::Q_BYTES);
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
                    tc_commit(&q_empty[qs]);
                    if (qt == num_qt - 1) tc_commit(k_empty);
                }
                __syncwarp();
                ++q_it;
                // wait for the probabilities (written into TMEM by the softmax warps), V and a free O
                mbar_wait(p_full, t_ph);
                if (qt == 0) mbar_wait(v_full, kv_ph);
                mbar_wait(o_empty, t_ph ^ 1);
                tc_fence_after();
                if (elect_one()) {
                    const int nks = p.NK / 16;
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint32_t a_hi = tmem_S + 32 * (ks >> 1) + 8 * (ks & 1);
                        const uint32_t a_lo = a_hi + 16;
                        const uint32_t voff = static_cast<uint32_t>(ks) * 16 * Cfg::SWZ;
                        const uint64_t v_hi = umma_smem_desc(sV + voff, 8 * Cfg::SWZ, 8 * Cfg::SWZ, Cfg::LAYOUT);
                        const uint64_t v_lo =
                            umma_smem_desc(sV + kv_plane + voff, 8 * Cfg::SWZ, 8 * Cfg::SWZ, Cfg::LAYOUT);
                        if (PASSES == 3) {
                            umma_ts(tmem_O, a_lo, v_hi, idesc_o, ks != 0);
                            umma_ts(tmem_O, a_hi, v_lo, idesc_o, 1);
                            umma_ts(tmem_O, a_hi, v_hi, idesc_o, 1);
                        } else {
                            umma_ts(tmem_O, a_hi, v_hi, idesc_o, ks != 0);
                        }
                    }
                    tc_commit(o_full);
                    if (qt == num_qt - 1) tc_commit(v_empty);
                }
                __syncwarp();
                ++t_it;
            }
            ++kv_it;
        }
    } else {
        // ---------------------------------------------------------------- softmax + output (warps 2..17)
        const int quad = warp & 3;
        const int grp = (warp - 2) >> 2;                // which quarter of the key chunks / output columns
        const int r_in_tile = quad * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const int nch = (p.NK + 31) / 32;
        const int ch_lo = grp * 2 < nch ? grp * 2 : nch;
        const int ch_hi = (grp * 2 + 2 < nch) ? grp * 2 + 2 : nch;       // my chunks [ch_lo, ch_hi)
        float* red_max = reinterpret_cast<float*>(smem + Cfg::OFF_RED);   // [4][128]
        float* red_sum = red_max + ATT_T_GROUPS * 128;                    // [4][128]
        const float sl2 = p.scale_log2e;
        uint32_t t_it = 0;
        for (int prob = blockIdx.x; prob < num_prob; prob += gridDim.x) {
            const int h = prob % p.H, j = (prob / p.H) % p.J, b = prob / (p.H * p.J);
            for (int qt = 0; qt < num_qt; ++qt) {
                const uint32_t t_ph = t_it & 1;
                mbar_wait(s_full, t_ph);
                tc_fence_after();
                // pass 1: row max of the raw scores over my key chunks (scale > 0 commutes with max)
                float mx = -INFINITY;
                for (int ch = ch_lo; ch < ch_hi; ++ch) {
                    uint32_t r[32];
                    tmem_ld32(tmem_S + lane_off + ch * 32, r);
                    tmem_ld_wait();
                    if (ch * 32 + 32 <= p.F) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (ch * 32 + i < p.F) mx = fmaxf(mx, __uint_as_float(r[i]));
                    }
                }
                red_max[grp * 128 + r_in_tile] = mx;
                named_bar_sync(1, ATT_T_SM_THREADS);
                mx = fmaxf(fmaxf(red_max[r_in_tile], red_max[128 + r_in_tile]),
                           fmaxf(red_max[256 + r_in_tile], red_max[384 + r_in_tile]));
                const float mxs = mx * sl2;
                // pass 2: p = 2^(s*c - max*c), partial row sum, bf16 hi/lo split written back over S
                float sum = 0.f;
                for (int ch = ch_lo; ch < ch_hi; ++ch) {
                    uint32_t r[32];
                    tmem_ld32(tmem_S + lane_off + ch * 32, r);
                    tmem_ld_wait();
                    uint32_t hi[16], lo[16];
                    const bool full = ch * 32 + 32 <= p.F;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * i]), sl2, -mxs));
                        float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), sl2, -mxs));
                        if (!full) {
                            if (ch * 32 + 2 * i >= p.F) p0 = 0.f;
                            if (ch * 32 + 2 * i + 1 >= p.F) p1 = 0.f;
                        }
                        sum += p0 + p1;
                        split2(p0, p1, hi[i], lo[i]);
                    }
                    tmem_st16(tmem_S + lane_off + ch * 32, hi);
                    if (PASSES == 3) tmem_st16(tmem_S + lane_off + ch * 32 + 16, lo);
                }
                red_sum[grp * 128 + r_in_tile] = sum;
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(p_full);
                named_bar_sync(1, ATT_T_SM_THREADS);
                const float inv = 1.0f / ((red_sum[r_in_tile] + red_sum[128 + r_in_tile]) +
                                          (red_sum[256 + r_in_tile] + red_sum[384 + r_in_tile]));

                // output tile: thread (row, grp) writes 16 of the HD columns (HD == 32: groups 0 and 1 only)
                mbar_wait(o_full, t_ph);
                tc_fence_after();
                const int t = qt * ATT_BM + r_in_tile;
                const bool ok = t < p.F;
                const size_t tok = (static_cast<size_t>(b) * p.F + (ok ? t : 0)) * p.J + j;
                if (grp * 16 < HD) {
                    const int c0 = grp * 16;
                    const size_t ob = tok * p.C + h * HD + c0;
                    uint32_t r[16];
                    tmem_ld16(tmem_O + lane_off + c0, r);
                    tmem_ld_wait();
                    if (ok && p.out_f16c) {
                        float xv[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) xv[i] = __uint_as_float(r[i]) * inv;
                        store16_f16c(reinterpret_cast<uint8_t*>(p.out_hi) + tok * p.C * 4, h * HD + c0, xv);
                    } else if (ok) {
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            split2(__uint_as_float(r[2 * i]) * inv, __uint_as_float(r[2 * i + 1]) * inv, hi[i], lo[i]);
                        uint4* h4 = reinterpret_cast<uint4*>(p.out_hi + ob);
                        h4[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        h4[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                        if (p.out_lo) {
                            uint4* l4 = reinterpret_cast<uint4*>(p.out_lo + ob);
                            l4[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            l4[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(o_empty);
                ++t_it;
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
