// Persistent, warp-specialised tcgen05 GEMM for the DSTformer linears (sm_100a).
//
//   D[M,N] = A[M,K] * W[N,K]^T     (nn.Linear: W is [out,in] = K-major, exactly what UMMA wants)
//
// Operands live in HBM as bf16 "hi" and "lo" planes (x ~= hi + lo, 16 mantissa bits).
//   PASSES == 3 : BF16x3 math  hi*hi + hi*lo + lo*hi  -> fp32-parity mode (rel err ~1e-5)
//   PASSES == 1 : plain bf16 (hi plane only)          -> bf16 training configs
// Tiles: 128 x 256 x BK (BK = 32 / SWIZZLE_64B for 3 passes, 64 / SWIZZLE_128B for 1 pass), 4-stage
// TMA->smem ring (48 KB / stage), two 256-column fp32 accumulators in TMEM (all 512 columns) so the
// epilogue of tile i overlaps the MMAs of tile i+1.
// Warp roles (320 threads): w0 = TMA producer, w1 = MMA issuer (+TMEM alloc), w2..w9 = epilogue
// (TMEM lane quadrant = warp_idx % 4, column half = (warp_idx-2)/4; thread = one output row x 128 columns,
// 32-column chunks via tcgen05.ld; residual loads are register double-buffered one chunk ahead).
//
// The epilogue is where the reference's elementwise ops are folded (DSTformer.py line refs):
//   EPI_LN_SPLIT      y = rstd*(acc - mean*s[n]) + c[n]            LayerNorm folded algebraically   (:241,243 qkv)
//   EPI_LN_GELU_SPLIT y = gelu_erf(...)                                                              (:242,244 fc1 + :81)
//   EPI_RESID         x' = x + rowscale*(acc + b[n]); also emits bf16 hi/lo of x' and LN partial
//                     statistics of x' for the next sublayer                                         (:241-249 residual)
//   EPI_LN_TANH_F32   rep = tanh(...)                                                               (:352-354)
//   EPI_BIAS_F32      y = acc + b[n]                                                                 (plain; tests)
#pragma once
#include "ptx.cuh"

namespace mb {

enum : int { EPI_LN_SPLIT = 0, EPI_LN_GELU_SPLIT = 1, EPI_RESID = 2, EPI_LN_TANH_F32 = 3, EPI_BIAS_F32 = 4,
              EPI_BIAS_SPLIT = 5,      /* y = acc + b[n] -> bf16 plane                          (backward: qkv recompute, dgrad) */
              EPI_BIAS_GELU_PAIR = 6,  /* plane 0 = y = acc + b[n], plane 1 = gelu(y)             (backward: fc1 recompute)         */
              EPI_GELUBWD_SPLIT = 7,   /* y = acc * gelu'(aux[m, n]) -> bf16 plane                (backward: fc2 dgrad -> d h_pre)  */
              EPI_LN_TANH_POOL = 8     /* rep = tanh(LN-folded linear), never stored: its mean over the F frames of each
                                          (clip, joint) is accumulated into out_f32[(b J + j), n]  (model_action.py:20-21)  */
};                                     /* 5..8 exist in the 2-CTA kernel only */

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 256;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 320;
constexpr int GEMM_EPI_THREADS = 256;
constexpr int STATS_GROUP = 128;   // LayerNorm partial statistics are kept per 128-column group

struct GemmParams {
    int M, N, K;
    const float* vec0;        // bias[n] (RESID/BIAS) or c[n] (LN modes)
    const float* vec1;        // s[n] = sum_k W'[n,k] (LN modes)
    const float* stats_in;    // LN modes: [M][nh_in][3] = (shift, sum(x-shift), sum((x-shift)^2)) per 128-col group
    int nh_in;                // groups per row of the LN input (= C/128)
    float ln_dim;             // C (number of normalised features)
    float eps;
    const float* resid;       // RESID: fp32 [M,N]
    const float* row_scale;   // RESID: optional DropPath scale per frame (row / J), may be null
    int J;
    float* out_f32;           // RESID / *_F32
    __nv_bfloat16* out_hi;    // split outputs [M,N]
    __nv_bfloat16* out_lo;    // may be null when PASSES == 1
    float* stats_out;         // RESID: [M][N/128][3]
    const __nv_bfloat16* aux; // GELUBWD: pre-activation plane [M,N] (bf16)
    int pool_F;               // LN_TANH_POOL: frames per clip (rows are (b F + f) J + j)
};

template <int PASSES>
struct GemmCfg {
    static constexpr int BK = (PASSES == 3) ? 32 : 64;
    static constexpr int SWZ = BK * 2;                          // bytes per smem row == swizzle span
    static constexpr uint32_t LAYOUT = (SWZ == 128) ? 2u : 4u;  // SWIZZLE_128B : SWIZZLE_64B
    static constexpr int PLANES = (PASSES == 3) ? 2 : 1;
    static constexpr int A_PLANE = GEMM_BM * SWZ;               // bytes
    static constexpr int B_PLANE = GEMM_BN * SWZ;
    static constexpr int A_BYTES = PLANES * A_PLANE;
    static constexpr int B_BYTES = PLANES * B_PLANE;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;       // 48 KB in both modes
    static constexpr int SMEM_BYTES = GEMM_STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// nn.GELU() exact (erf) form: gelu(x) = x * Phi(x), Phi(x) = 0.5 * erfc(-x / sqrt(2)).
// erfc(|z|) by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7 in exact arithmetic, ~4e-7 in fp32 -- the same
// as an fp32 evaluation through erff), branch-free on MUFU.RCP / MUFU.EX2: ~14 instructions instead of ~35,
// which is what keeps the fc1 epilogue under its mainloop (profiles/r01: fc1 was epilogue-ALU bound with erff).
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    poly *= t;
    const float q = 0.5f * poly * ex2_approx(-z * z * 1.4426950408889634f);   // 0.5 * erfc(|z|)
    return x * (x >= 0.f ? 1.0f - q : q);
}

// Two values at once on Blackwell's packed fp32x2 pipe instructions (FFMA2 / FMUL2 / FADD2: one issue slot for two lanes'
// worth of work -- the F16C GEMM epilogues are issue-bound, profiles/r02a) and ONE MUFU per value instead of two:
//   Phi(-t) = 0.5 erfc(t / sqrt 2) = 2^P(t),  t = min(|x|, 5.75),  P = degree-8 weighted-minimax fit of log2(0.5 erfc(t / sqrt 2))
//   (|Phi error| <= 2.3e-7 on the whole axis; beyond 5.75 Phi(-t) < 5e-9: clamped);  gelu(x) = x (0.5 + copysign(0.5 - 2^P, x)).
// |gelu error| <= 4.5e-7 absolute (2.6e-7 for |x| < 3; the rcp + ex2 Abramowitz-Stegun form of gelu_erf: 6.1e-7 / 4.1e-7,
// both measured against float64 erf over [-9, 9] with every fp32 rounding emulated).  17 instructions per pair (8 FFMA2,
// 2 FMNMX, 2 MUFU.EX2, 2 FADD2, 2 LOP3, 1 FMUL2) against 19 with 4 MUFU.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
    const float2 t = make_float2(fminf(fabsf(x.x), 5.75f), fminf(fabsf(x.y), 5.75f));
    float2 p = __ffma2_rn(make_float2(-1.800373980e-06f, -1.800373980e-06f), t, make_float2(2.663698069e-05f, 2.663698069e-05f));
    p = __ffma2_rn(p, t, make_float2(-1.234105584e-04f, -1.234105584e-04f));
    p = __ffma2_rn(p, t, make_float2(-2.961509454e-04f, -2.961509454e-04f));
    p = __ffma2_rn(p, t, make_float2(7.286241278e-03f, 7.286241278e-03f));
    p = __ffma2_rn(p, t, make_float2(-5.266715959e-02f, -5.266715959e-02f));
    p = __ffma2_rn(p, t, make_float2(-4.591412842e-01f, -4.591412842e-01f));
    p = __ffma2_rn(p, t, make_float2(-1.151116371e+00f, -1.151116371e+00f));
    p = __ffma2_rn(p, t, make_float2(-1.0f, -1.0f));
    const float2 d = __fadd2_rn(make_float2(0.5f, 0.5f), make_float2(-ex2_approx(p.x), -ex2_approx(p.y)));   // 0.5 - Phi(-t)
    const float2 phi = __fadd2_rn(make_float2(0.5f, 0.5f), make_float2(copysignf(d.x, x.x), copysignf(d.y, x.y)));
    return __fmul2_rn(x, phi);
}

// d/dx [x Phi(x)] = Phi(x) + x phi(x), with Phi from the same erfc polynomial and phi from the SAME exponential
// (exp(-z^2) with z = |x|/sqrt(2) is exp(-x^2/2)): one MUFU.RCP + one MUFU.EX2.
__device__ __forceinline__ float gelu_grad_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    poly *= t;
    const float e = ex2_approx(-z * z * 1.4426950408889634f);   // exp(-x^2 / 2)
    const float q = 0.5f * poly * e;                              // 0.5 * erfc(|z|)
    return fmaf(x, 0.3989422804014327f * e, x >= 0.f ? 1.0f - q : q);
}

// Combine per-group partial statistics (Chan et al.) -> mean, rstd of a row.
__device__ __forceinline__ void ln_row_stats(const float* __restrict__ st, int nh, float dim, float eps,
                                             float& mean, float& rstd) {
    float n_a = 0.f, mean_a = 0.f, m2_a = 0.f;
    const float n_h = static_cast<float>(STATS_GROUP);
    for (int h = 0; h < nh; ++h) {
        const float shift = st[3 * h + 0], s1 = st[3 * h + 1], s2 = st[3 * h + 2];
        const float mean_h = shift + s1 / n_h;
        const float m2_h = fmaxf(s2 - s1 * s1 / n_h, 0.f);
        const float delta = mean_h - mean_a;
        const float n = n_a + n_h;
        mean_a += delta * (n_h / n);
        m2_a += m2_h + delta * delta * (n_a * n_h / n);
        n_a = n;
    }
    mean = mean_a;
    rstd = 1.0f / sqrtf(m2_a / dim + eps);
}

#ifdef MB_TEST_KERNELS   // first-generation 1-CTA kernel: kept for tests (libmotionbert_b200_test.so) only
template <int PASSES, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,   // 3D (K, M, plane), box (BK, 128, PLANES)
               const __grid_constant__ CUtensorMap tmB,   // 3D (K, N, plane), box (BK, 256, PLANES)
               const GemmParams p) {
    using Cfg = GemmCfg<PASSES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                       // [STAGES]  TMA -> MMA
    uint64_t* empty_bar = bars + GEMM_STAGES;        // [STAGES]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * GEMM_STAGES;    // [2]       MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * GEMM_STAGES + 2;   // [2]   epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int num_m = (p.M + GEMM_BM - 1) / GEMM_BM;
    const int num_n = p.N / GEMM_BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = p.K / Cfg::BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < GEMM_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], GEMM_EPI_THREADS);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_idx = tile / num_n, n_idx = tile % num_n;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sB = sA + Cfg::A_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    tma_load_3d(sA, &tmA, &full_bar[stage], kb * Cfg::BK, m_idx * GEMM_BM, 0);
                    tma_load_3d(sB, &tmB, &full_bar[stage], kb * Cfg::BK, n_idx * GEMM_BN, 0);
                    if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t IDESC = umma_idesc_bf16(GEMM_BM, GEMM_BN, 0, 0);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * GEMM_BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sB = sA + Cfg::A_BYTES;
                    const uint64_t a_hi = umma_smem_desc(sA, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                    const uint64_t b_hi = umma_smem_desc(sB, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                    const uint64_t a_lo = umma_smem_desc(sA + Cfg::A_PLANE, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
                    const uint64_t b_lo = umma_smem_desc(sB + Cfg::B_PLANE, 16, 8 * Cfg::SWZ, Cfg::LAYOUT);
#pragma unroll
                    for (int ks = 0; ks < Cfg::BK / 16; ++ks) {
                        const uint64_t koff = static_cast<uint64_t>((ks * 32) >> 4);   // +32 B per K=16 step
                        if (PASSES == 3) {
                            // small cross terms first, dominant hi*hi last
                            umma_ss(d_tmem, a_lo + koff, b_hi + koff, IDESC, (kb | ks) != 0);
                            umma_ss(d_tmem, a_hi + koff, b_lo + koff, IDESC, 1);
                            umma_ss(d_tmem, a_hi + koff, b_hi + koff, IDESC, 1);
                        } else {
                            umma_ss(d_tmem, a_hi + koff, b_hi + koff, IDESC, (kb | ks) != 0);
                        }
                    }
                    tc_commit(&empty_bar[stage]);                 // smem slot free once these MMAs retire
                    if (kb == num_kb - 1) tc_commit(&tfull_bar[acc]);   // accumulator ready
                }
                __syncwarp();
                if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9)
        const int quad = warp & 3;                      // TMEM lanes [32*quad, 32*quad+32)
        const int half = (warp - 2) >> 2;               // columns [128*half, 128*half+128) of the tile
        constexpr int NCH = GEMM_BN / 2 / 32;           // 4 chunks of 32 columns per thread
        const int ngrp_out = p.N / STATS_GROUP;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_idx = tile / num_n, n_idx = tile % num_n;
            const int row = m_idx * GEMM_BM + quad * 32 + lane;
            const bool row_ok = row < p.M;
            const size_t row_off = static_cast<size_t>(row) * p.N;
            const int colbase = n_idx * GEMM_BN + half * (GEMM_BN / 2);

            float mean = 0.f, rstd = 1.f, rscale = 1.f;
            if (EPI == EPI_LN_SPLIT || EPI == EPI_LN_GELU_SPLIT || EPI == EPI_LN_TANH_F32) {
                if (row_ok) ln_row_stats(p.stats_in + static_cast<size_t>(row) * p.nh_in * 3, p.nh_in, p.ln_dim,
                                         p.eps, mean, rstd);
            }
            // residual prefetch (one chunk ahead, issued before the accumulator is even ready)
            float4 xr[2][8];
            if (EPI == EPI_RESID) {
                if (row_ok && p.row_scale) rscale = p.row_scale[row / p.J];
                const float4* x4 = reinterpret_cast<const float4*>(p.resid + row_off + colbase);
#pragma unroll
                for (int i = 0; i < 8; ++i) xr[0][i] = row_ok ? x4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float st_shift = 0.f, st_sum = 0.f, st_sq = 0.f;

            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * GEMM_BN + half * (GEMM_BN / 2) +
                                   (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int col0 = colbase + ch * 32;
                if (EPI == EPI_RESID) {
                    if (ch + 1 < NCH) {
                        const float4* x4 = reinterpret_cast<const float4*>(p.resid + row_off + col0 + 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            xr[(ch + 1) & 1][i] = row_ok ? x4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                uint32_t r[32];
                tmem_ld32(t_row + ch * 32, r);
                tmem_ld_wait();
                float v[32];
                if (EPI == EPI_RESID) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.vec0 + col0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = __ldg(b4 + i);
                        const float4 x = xr[ch & 1][i];
                        v[4 * i + 0] = x.x + rscale * (__uint_as_float(r[4 * i + 0]) + b.x);
                        v[4 * i + 1] = x.y + rscale * (__uint_as_float(r[4 * i + 1]) + b.y);
                        v[4 * i + 2] = x.z + rscale * (__uint_as_float(r[4 * i + 2]) + b.z);
                        v[4 * i + 3] = x.w + rscale * (__uint_as_float(r[4 * i + 3]) + b.w);
                    }
                    if (ch == 0) st_shift = v[0];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float d = v[i] - st_shift;
                        st_sum += d;
                        st_sq = fmaf(d, d, st_sq);
                    }
                } else if (EPI == EPI_BIAS_F32) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.vec0 + col0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = __ldg(b4 + i);
                        v[4 * i + 0] = __uint_as_float(r[4 * i + 0]) + b.x;
                        v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + b.y;
                        v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + b.z;
                        v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + b.w;
                    }
                } else {
                    const float4* c4 = reinterpret_cast<const float4*>(p.vec0 + col0);
                    const float4* s4 = reinterpret_cast<const float4*>(p.vec1 + col0);
                    const float ms = -mean * rstd;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 c = __ldg(c4 + i);
                        const float4 s = __ldg(s4 + i);
                        v[4 * i + 0] = fmaf(rstd, __uint_as_float(r[4 * i + 0]), fmaf(ms, s.x, c.x));
                        v[4 * i + 1] = fmaf(rstd, __uint_as_float(r[4 * i + 1]), fmaf(ms, s.y, c.y));
                        v[4 * i + 2] = fmaf(rstd, __uint_as_float(r[4 * i + 2]), fmaf(ms, s.z, c.z));
                        v[4 * i + 3] = fmaf(rstd, __uint_as_float(r[4 * i + 3]), fmaf(ms, s.w, c.w));
                    }
                    if (EPI == EPI_LN_GELU_SPLIT) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
                    }
                    if (EPI == EPI_LN_TANH_F32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = tanhf(v[i]);
                    }
                }
                if (row_ok) {
                    if (EPI == EPI_RESID || EPI == EPI_LN_TANH_F32 || EPI == EPI_BIAS_F32) {
                        float4* o4 = reinterpret_cast<float4*>(p.out_f32 + row_off + col0);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                    if ((EPI == EPI_RESID || EPI == EPI_LN_SPLIT || EPI == EPI_LN_GELU_SPLIT) && p.out_hi) {
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
                        uint4* h4 = reinterpret_cast<uint4*>(p.out_hi + row_off + col0);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            h4[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                        if (p.out_lo) {
                            uint4* l4 = reinterpret_cast<uint4*>(p.out_lo + row_off + col0);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                l4[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
                        }
                    }
                }
            }
            // all TMEM reads of this accumulator are complete -> hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
            if (EPI == EPI_RESID) {
                if (row_ok && p.stats_out) {
                    float* so = p.stats_out + (static_cast<size_t>(row) * ngrp_out + n_idx * 2 + half) * 3;
                    so[0] = st_shift;
                    so[1] = st_sum;
                    so[2] = st_sq;
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

#endif  // MB_TEST_KERNELS

}  // namespace mb
