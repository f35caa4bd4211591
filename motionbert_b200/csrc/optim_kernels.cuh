// Optimizer step and weight re-pack of the training loop (train.py:289 `optim.AdamW`,
// train.py:206 `optimizer.step()`):
//   * adamw_group_kernel : decoupled-weight-decay Adam over up to 48 parameter tensors per launch (pointer table in
//     kernel-parameter space), torch.optim.AdamW's arithmetic (p *= 1 - lr wd; m, v moments; bias corrections;
//     p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps));
//   * pack_group_kernel  : the LayerNorm-folded tensor-core operand packing of pack_linear_kernel for up to 40 linears per
//     launch (81 linears -> 3 launches instead of 81).
#pragma once
#include "simt_kernels.cuh"

namespace mb {

constexpr int ADAMW_GROUP = 48;
constexpr int ADAMW_CHUNK = 4096;          // elements per CTA

struct AdamWGroup {
    float* p[ADAMW_GROUP];
    const float* g[ADAMW_GROUP];
    float* m[ADAMW_GROUP];
    float* v[ADAMW_GROUP];
    int chunk_end[ADAMW_GROUP];            // inclusive prefix of chunks per tensor
    int numel[ADAMW_GROUP];
    int n;
    float lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt;   // bc1 = 1 - beta1^t, bc2_sqrt = sqrt(1 - beta2^t)
};

__global__ void __launch_bounds__(256) adamw_group_kernel(const __grid_constant__ AdamWGroup G) {
    int t = 0;
    while (t < G.n - 1 && static_cast<int>(blockIdx.x) >= G.chunk_end[t]) ++t;
    const int chunk = blockIdx.x - (t ? G.chunk_end[t - 1] : 0);
    const int n = G.numel[t];
    float* __restrict__ p = G.p[t];
    const float* __restrict__ g = G.g[t];
    float* __restrict__ m = G.m[t];
    float* __restrict__ v = G.v[t];
    const float decay = 1.0f - G.lr * G.weight_decay;
    const float step_size = G.lr / G.bc1;
    const int lo = chunk * ADAMW_CHUNK;
    const int hi = lo + ADAMW_CHUNK < n ? lo + ADAMW_CHUNK : n;
    for (int i = lo + threadIdx.x; i < hi; i += 256) {
        const float gi = g[i];
        const float mi = G.beta1 * m[i] + (1.0f - G.beta1) * gi;
        const float vi = G.beta2 * v[i] + (1.0f - G.beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / G.bc2_sqrt + G.eps;
        p[i] = p[i] * decay - step_size * (mi / denom);
    }
}

constexpr int PACK_GROUP = 40;
struct PackGroup {
    const float* W[PACK_GROUP];
    const float* b[PACK_GROUP];
    const float* gamma[PACK_GROUP];        // null: no LayerNorm fold
    const float* beta[PACK_GROUP];
    __nv_bfloat16* hi[PACK_GROUP];
    __nv_bfloat16* lo[PACK_GROUP];
    float* vec_c[PACK_GROUP];
    float* vec_s[PACK_GROUP];              // null when gamma is null
    int N[PACK_GROUP], K[PACK_GROUP];
    int block_end[PACK_GROUP];             // inclusive prefix of CTAs (8 rows each) per linear
    int n;
    int f16c;
};

__device__ __forceinline__ void pack_linear_row(const float* __restrict__ W, const float* __restrict__ b,
                                                const float* __restrict__ gamma, const float* __restrict__ beta, int n,
                                                int K, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                float* __restrict__ vec_c, float* __restrict__ vec_s, int f16c) {
    const int lane = lane_id();
    float s = 0.f, c = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float w = W[static_cast<size_t>(n) * K + k];
        const float wp = gamma ? w * gamma[k] : w;
        if (f16c) {
            uint8_t* blk = reinterpret_cast<uint8_t*>(hi) + (static_cast<size_t>(n) * K + (k & ~31)) * 4;
            uint32_t h2, l2, g2;
            split2_f16c(wp, 0.f, h2, l2, g2);
            *reinterpret_cast<uint16_t*>(blk + 2 * lane) = static_cast<uint16_t>(h2 & 0xffffu);
            blk[64 + lane] = static_cast<uint8_t>(l2 & 0xffu);
            blk[96 + lane] = static_cast<uint8_t>(g2 & 0xffu);
            s += wp;
        } else {
            __nv_bfloat16 h, l;
            split_bf16(wp, h, l);
            hi[static_cast<size_t>(n) * K + k] = h;
            lo[static_cast<size_t>(n) * K + k] = l;
            s += __bfloat162float(h) + __bfloat162float(l);
        }
        if (beta) c = fmaf(beta[k], w, c);
    }
    s = warp_sum(s);
    c = warp_sum(c);
    if (lane == 0) {
        vec_c[n] = c + (b ? b[n] : 0.f);
        if (vec_s) vec_s[n] = s;
    }
}

__global__ void __launch_bounds__(256) pack_group_kernel(const __grid_constant__ PackGroup G) {
    int t = 0;
    while (t < G.n - 1 && static_cast<int>(blockIdx.x) >= G.block_end[t]) ++t;
    const int blk = blockIdx.x - (t ? G.block_end[t - 1] : 0);
    const int n = blk * 8 + (threadIdx.x >> 5);
    if (n >= G.N[t]) return;
    pack_linear_row(G.W[t], G.b[t], G.gamma[t], G.beta[t], n, G.K[t], G.hi[t], G.lo[t], G.vec_c[t], G.vec_s[t], G.f16c);
}

}  // namespace mb
