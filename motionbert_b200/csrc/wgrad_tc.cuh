// Weight-gradient GEMM on tcgen05 (groundwork for SURVEY.md section 8 row a15, the native backward):
//
//   dW[N, K] += G[M, N]^T * X[M, K]          (nn.Linear y = x W^T:  dW = dy^T x, reduction over the M tokens)
//
// Both operands are token-major in HBM, i.e. the reduction index m is the SLOW index of both: exactly UMMA's
// "MN-major" operand form (a_major = b_major = 1), so no transposed copy of the activations is ever made:
//   A = G^T tile  [128 n] x [BT tokens]   smem: per 64-n block [BT token rows][128 B], SWIZZLE_128B, LBO = block stride
//   B = X   tile  [256 k] x [BT tokens]   smem: per 64-k block [BT token rows][128 B]
// One CTA owns one 128 x 256 tile of dW and one slice of the token range (split-K over M, the long dimension:
// ~0.5 M tokens vs N, K <= 1536); fp32 accumulation in TMEM, then `red.global.add.f32` into the (pre-zeroed,
// later NCCL-reduced) gradient buffer.  BF16x3 (hi/lo planes) or single-pass bf16 like the forward kernels.
#pragma once
#include "ptx.cuh"

namespace mb {

constexpr int WG_THREADS = 192;   // w0 TMA, w1 MMA, w2..w5 epilogue (TMEM lane quadrant = warp % 4)
constexpr int WG_STAGES = 4;
// tokens per pipeline stage: 64 in single-pass mode (48 KB per stage, 192 KB in flight -- with 32-token stages the
// 4-deep ring held only ~1000 MMA-cycles of work, less than one TMA round trip: ncu showed the tensor pipe 49 % active),
// 32 in BF16x3 mode (two planes per operand)
template <int PASSES>
struct WgBt { static constexpr int value = (PASSES == 3) ? 32 : 64; };

struct WgradParams {
    int M, N, K;
    int tokens_per_split;   // multiple of the stage size WgBt<PASSES>::value
    float* dW;              // [N, K] fp32, accumulated with atomics
};

template <int PASSES>
struct WgradCfg {
    static constexpr int PLANES = (PASSES == 3) ? 2 : 1;
    static constexpr int BT = WgBt<PASSES>::value;
    static constexpr int BLK = BT * 128;                       // one 64-element MN block: [BT token rows][128 B]
    static constexpr int A_PLANE = 2 * BLK;                    // 128 n  = 2 blocks
    static constexpr int B_PLANE = 4 * BLK;                    // 256 k  = 4 blocks
    static constexpr int A_BYTES = PLANES * A_PLANE;
    static constexpr int B_BYTES = PLANES * B_PLANE;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;      // 48 KB in both modes
    static constexpr int SMEM_BYTES = WG_STAGES * STAGE_BYTES + 256 + 1024;
};

template <int PASSES>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmG,   // bf16 3D (N, M, plane), box (64, BT, 1)
             const __grid_constant__ CUtensorMap tmX,   // bf16 3D (K, M, plane), box (64, BT, 1)
             const WgradParams p) {
    using Cfg = WgradCfg<PASSES>;
    constexpr int WG_BT = Cfg::BT;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + WG_STAGES;
    uint64_t* done_bar = bars + 2 * WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);

    const int warp = warp_uniform(threadIdx.x >> 5);      // uniform role dispatch (see ptx.cuh elect_one)
    const int lane = threadIdx.x & 31;
    const int tiles_k = p.K / 256;
    const int tiles = (p.N / 128) * tiles_k;
    const int tile = blockIdx.x % tiles;
    const int split = blockIdx.x / tiles;
    const int n0 = (tile / tiles_k) * 128;
    const int k0 = (tile % tiles_k) * 256;
    const int m_begin = split * p.tokens_per_split;
    int m_end = m_begin + p.tokens_per_split;
    if (m_end > p.M) m_end = p.M;
    const int nblk = m_begin < m_end ? (m_end - m_begin + WG_BT - 1) / WG_BT : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmG);
        tma_prefetch_desc(&tmX);
        for (int i = 0; i < WG_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < nblk; ++kb) {
                const int m = m_begin + kb * WG_BT;          // rows past M are zero-filled by TMA
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
                uint8_t* sB = sA + Cfg::A_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                for (int pl = 0; pl < Cfg::PLANES; ++pl) {
                    for (int blk = 0; blk < 2; ++blk)
                        tma_load_3d(sA + pl * Cfg::A_PLANE + blk * Cfg::BLK, &tmG, &full_bar[stage], n0 + blk * 64, m, pl);
                    for (int blk = 0; blk < 4; ++blk)
                        tma_load_3d(sB + pl * Cfg::B_PLANE + blk * Cfg::BLK, &tmX, &full_bar[stage], k0 + blk * 64, m, pl);
                }
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t IDESC = umma_idesc_bf16(128, 256, 1, 1);      // A and B MN-major
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < nblk; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
                for (int ks = 0; ks < WG_BT / 16; ++ks) {
                    const uint32_t roff = ks * 16 * 128;                  // 16 token rows further down each block
                    // MN-major, SWIZZLE_128B: LBO = distance between 64-element MN blocks, SBO = 8 rows x 128 B
                    const uint64_t a_hi = umma_smem_desc(sA + roff, Cfg::BLK, 1024, 2u);
                    const uint64_t b_hi = umma_smem_desc(sB + roff, Cfg::BLK, 1024, 2u);
                    const uint64_t a_lo = umma_smem_desc(sA + Cfg::A_PLANE + roff, Cfg::BLK, 1024, 2u);
                    const uint64_t b_lo = umma_smem_desc(sB + Cfg::B_PLANE + roff, Cfg::BLK, 1024, 2u);
                    if (PASSES == 3) {
                        umma_ss(tmem_base, a_lo, b_hi, IDESC, (kb | ks) != 0);
                        umma_ss(tmem_base, a_hi, b_lo, IDESC, 1);
                        umma_ss(tmem_base, a_hi, b_hi, IDESC, 1);
                    } else {
                        umma_ss(tmem_base, a_hi, b_hi, IDESC, (kb | ks) != 0);
                    }
                }
                tc_commit(&empty_bar[stage]);
                if (kb == nblk - 1) tc_commit(done_bar);
            }
            __syncwarp();
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (nblk > 0) {
        // epilogue: TMEM -> red.global.add (thread = dW row n, 32 consecutive k per chunk)
        const int quad = warp & 3;
        const int n = n0 + quad * 32 + lane;
        mbar_wait(done_bar, 0);
        tc_fence_after();
        float* dst = p.dW + static_cast<size_t>(n) * p.K + k0;
#pragma unroll 1
        for (int ch = 0; ch < 8; ++ch) {
            uint32_t r[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + ch * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + ch * 32 + 4 * i),
                             "f"(__uint_as_float(r[4 * i])), "f"(__uint_as_float(r[4 * i + 1])),
                             "f"(__uint_as_float(r[4 * i + 2])), "f"(__uint_as_float(r[4 * i + 3]))
                             : "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

}  // namespace mb
