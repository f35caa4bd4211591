// CUDA-core kernels of the native backward: everything that is not a GEMM or the
// attention core.  Backward arithmetic is bf16 single-pass on the tensor cores (the reference's training configs are
// bf16), fp32 for every reduction / residual-stream quantity.  Warp-per-token-row like the forward row kernels.
#pragma once
#include "simt_kernels.cuh"

namespace mb {

// mean / rstd of a token row from the per-128-column partial statistics the forward stored
__device__ __forceinline__ void row_mean_rstd(const float* __restrict__ stats, size_t row, int C, float eps, float& mean,
                                              float& rstd) {
    ln_row_stats(stats + row * (C / STATS_GROUP) * 3, C / STATS_GROUP, static_cast<float>(C), eps, mean, rstd);
}

// xhat = (x - mean) * rstd  ->  bf16 plane (the A operand of every backward GEMM that touches a LayerNorm'ed input)
template <int NV>
__global__ void __launch_bounds__(256) ln_xhat_kernel(const float* __restrict__ x, const float* __restrict__ stats, int M,
                                                       int C, float eps, __nv_bfloat16* __restrict__ xhat) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = lane_id();
    float mean, rstd;
    row_mean_rstd(stats, row, C, eps, mean, rstd);
    const size_t base = static_cast<size_t>(row) * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = 128 * i + 4 * lane;
        const float4 v = *reinterpret_cast<const float4*>(x + base + c);
        const __nv_bfloat162 a = __floats2bfloat162_rn((v.x - mean) * rstd, (v.y - mean) * rstd);
        const __nv_bfloat162 b = __floats2bfloat162_rn((v.z - mean) * rstd, (v.w - mean) * rstd);
        *reinterpret_cast<uint2*>(xhat + base + c) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
}

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&t);
}

// h = gelu(h_pre) on bf16 planes (recompute of the fc2 input)
__global__ void __launch_bounds__(256) gelu_plane_kernel(const __nv_bfloat16* __restrict__ hpre, size_t n8,
                                                          __nv_bfloat16* __restrict__ h) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const uint4 v = reinterpret_cast<const uint4*>(hpre)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = pack2(gelu_erf(bf16lo(w[e])), gelu_erf(bf16hi(w[e])));
    reinterpret_cast<uint4*>(h)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// d/dx [x Phi(x)] = Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_grad(float x) { return gelu_grad_fast(x); }
// dh_pre = dh * gelu'(h_pre):  dh fp32 (dgrad output), h_pre bf16 -> dh_pre bf16 plane
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const float* __restrict__ dh, const __nv_bfloat16* __restrict__ hpre,
                                                        size_t n8, __nv_bfloat16* __restrict__ dhpre) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const uint4 v = reinterpret_cast<const uint4*>(hpre)[i];
    const float4 g0 = reinterpret_cast<const float4*>(dh)[2 * i], g1 = reinterpret_cast<const float4*>(dh)[2 * i + 1];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = pack2(g[2 * e] * gelu_grad(bf16lo(w[e])), g[2 * e + 1] * gelu_grad(bf16hi(w[e])));
    reinterpret_cast<uint4*>(dhpre)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// same with dh already rounded to a bf16 plane by the data-gradient GEMM's epilogue (may run in place over dh)
__global__ void __launch_bounds__(256) gelu_bwd_plane_kernel(const __nv_bfloat16* dh, const __nv_bfloat16* __restrict__ hpre,
                                                              size_t n8, __nv_bfloat16* dhpre) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const uint4 v = reinterpret_cast<const uint4*>(hpre)[i];
    const uint4 gq = reinterpret_cast<const uint4*>(dh)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const uint32_t g[4] = {gq.x, gq.y, gq.z, gq.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
        o[e] = pack2(bf16lo(g[e]) * gelu_grad(bf16lo(w[e])), bf16hi(g[e]) * gelu_grad(bf16hi(w[e])));
    reinterpret_cast<uint4*>(dhpre)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// fp32 -> bf16 plane, optionally multiplied elementwise by (1 - t^2) (tanh backward)
__global__ void __launch_bounds__(256) to_plane_kernel(const float* __restrict__ g, const float* __restrict__ tanh_out,
                                                        size_t n2, __nv_bfloat16* __restrict__ out) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    float2 v = reinterpret_cast<const float2*>(g)[i];
    if (tanh_out) {
        const float2 t = reinterpret_cast<const float2*>(tanh_out)[i];
        v.x *= (1.0f - t.x * t.x);
        v.y *= (1.0f - t.y * t.y);
    }
    reinterpret_cast<uint32_t*>(out)[i] = pack2(v.x, v.y);
}

// column sums of a token-major matrix (bias gradients): out[n] += sum_m g[m, n].  One CTA owns 128 rows x 256 columns.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ g, int M, int N, float* __restrict__ out) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    const int m0 = blockIdx.y * 128;
    if (n >= N) return;
    float acc = 0.f;
    const int m1 = m0 + 128 < M ? m0 + 128 : M;
    for (int m = m0; m < m1; ++m) {
        if constexpr (sizeof(T) == 2) acc += __bfloat162float(g[static_cast<size_t>(m) * N + n]);
        else acc += g[static_cast<size_t>(m) * N + n];
    }
    atomicAdd(out + n, acc);
}

// Weight-space part of "LayerNorm folded into the next Linear" (y = xhat W'^T + c, W' = W*gamma, c = W beta + b):
// given dW' (= dY^T xhat) and dc (= sum_m dY):   dW = dW'*gamma + dc (x) beta ; dgamma = sum_n dW'*W ; dbeta = W^T dc ; db = dc
__global__ void __launch_bounds__(256) ln_linear_grad_kernel(const float* __restrict__ dWp, const float* __restrict__ dc,
                                                              const float* __restrict__ W, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int N, int K,
                                                              float* __restrict__ dW, float* __restrict__ db,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int k = blockIdx.x * 256 + threadIdx.x;          // one thread per input feature, loops over a slab of rows
    const int n0 = blockIdx.y * 64;
    if (k >= K) return;
    const float gk = gamma[k], bk = beta[k];
    float ag = 0.f, ab = 0.f;
    const int n1 = n0 + 64 < N ? n0 + 64 : N;
    for (int n = n0; n < n1; ++n) {
        const size_t o = static_cast<size_t>(n) * K + k;
        const float dwp = dWp[o], w = W[o], dcn = dc[n];
        dW[o] = fmaf(dwp, gk, dcn * bk);
        ag = fmaf(dwp, w, ag);
        ab = fmaf(dcn, w, ab);
        if (k == 0) db[n] = dcn;
    }
    atomicAdd(dgamma + k, ag);
    atomicAdd(dbeta + k, ab);
}

// LayerNorm backward + residual:  dx = dy_resid [+ extra] + rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))
// (dxhat = dL/d xhat from the data-gradient GEMM on the folded weights).  Emits fp32 and the bf16 plane that feeds
// the GEMMs of the previous sublayer.  with_ln = 0: plain pass-through sum (dx = dy_resid + extra).
template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_finalize_kernel(const float* __restrict__ dxhat, const float* __restrict__ x,
                                                               const float* __restrict__ stats, const float* __restrict__ dy,
                                                               const float* __restrict__ extra, int M, int C, float eps,
                                                               float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_plane) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = lane_id();
    const size_t base = static_cast<size_t>(row) * C;
    float mean = 0.f, rstd = 0.f;
    float4 g[NV], xh[NV];
    float s1 = 0.f, s2 = 0.f;
    if (dxhat) {
        row_mean_rstd(stats, row, C, eps, mean, rstd);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = 128 * i + 4 * lane;
            g[i] = *reinterpret_cast<const float4*>(dxhat + base + c);
            const float4 v = *reinterpret_cast<const float4*>(x + base + c);
            xh[i] = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
            s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
            s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
        }
        s1 = warp_sum(s1) / static_cast<float>(C);
        s2 = warp_sum(s2) / static_cast<float>(C);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = 128 * i + 4 * lane;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dy) o = *reinterpret_cast<const float4*>(dy + base + c);
        if (extra) {
            const float4 e = *reinterpret_cast<const float4*>(extra + base + c);
            o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
        }
        if (dxhat) {
            o.x += rstd * (g[i].x - s1 - xh[i].x * s2);
            o.y += rstd * (g[i].y - s1 - xh[i].y * s2);
            o.z += rstd * (g[i].z - s1 - xh[i].z * s2);
            o.w += rstd * (g[i].w - s1 - xh[i].w * s2);
        }
        if (dx) *reinterpret_cast<float4*>(dx + base + c) = o;
        if (dx_plane) *reinterpret_cast<uint2*>(dx_plane + base + c) = make_uint2(pack2(o.x, o.y), pack2(o.z, o.w));
    }
}

// DropPath backward (lib/model/drop.py:17-32): the branch of a residual sublayer sees scale[frame] * dy
template <int NV>
__global__ void __launch_bounds__(256) scale_rows_kernel(const float* __restrict__ g, const float* __restrict__ scale, int J,
                                                          int M, int C, float* __restrict__ out,
                                                          __nv_bfloat16* __restrict__ out_plane) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = lane_id();
    const float sc = scale[row / J];
    const size_t base = static_cast<size_t>(row) * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = 128 * i + 4 * lane;
        float4 v = *reinterpret_cast<const float4*>(g + base + c);
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
        *reinterpret_cast<float4*>(out + base + c) = v;
        *reinterpret_cast<uint2*>(out_plane + base + c) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
    }
}

// gradient w.r.t. the pose input: d x_in[m, k] = sum_c d x0[m, c] We[c, k]   (DSTformer.py:333)
__global__ void __launch_bounds__(256) embed_dx_kernel(const float* __restrict__ dx0, const float* __restrict__ We, int M,
                                                        int C, int dim_in, float* __restrict__ dxin) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = lane_id();
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float g = dx0[static_cast<size_t>(row) * C + c];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < dim_in) acc[k] = fmaf(g, We[static_cast<size_t>(c) * dim_in + k], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float t = warp_sum(acc[k]);
        if (lane == 0 && k < dim_in) dxin[static_cast<size_t>(row) * dim_in + k] = t;
    }
}

// S/T fusion backward (DSTformer.py:343-349):  x = a0 x_st + a1 x_ts,  a = softmax([x_st, x_ts] Wa^T + ba)
// One warp walks FUSE_ROWS consecutive token rows and keeps its share of dWa (2 logits x 2C columns) in registers;
// the 8 warps of a CTA then merge through shared-memory atomics and issue ONE global atomic per dWa element per CTA.
constexpr int FUSE_ROWS = 32;
template <int NV>
__global__ void __launch_bounds__(256) fuse_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ xst,
                                                        const float* __restrict__ xts, const float* __restrict__ Wa,
                                                        const float* __restrict__ ba, int M, int C, float* __restrict__ dxst,
                                                        float* __restrict__ dxts, __nv_bfloat16* __restrict__ dxst_plane,
                                                        __nv_bfloat16* __restrict__ dxts_plane, float* __restrict__ dWa,
                                                        float* __restrict__ dba) {
    extern __shared__ float s_acc[];     // [4*C] dWa partials + [2] dba partials
    const int warp = threadIdx.x >> 5, lane = lane_id();
    for (int i = threadIdx.x; i < 4 * C + 2; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    float4 w0a[NV], w0b[NV], w1a[NV], w1b[NV];
    float4 a00[NV], a01[NV], a10[NV], a11[NV];     // dl0*x_st, dl0*x_ts, dl1*x_st, dl1*x_ts
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = 128 * i + 4 * lane;
        w0a[i] = __ldg(reinterpret_cast<const float4*>(Wa + c));
        w0b[i] = __ldg(reinterpret_cast<const float4*>(Wa + C + c));
        w1a[i] = __ldg(reinterpret_cast<const float4*>(Wa + 2 * C + c));
        w1b[i] = __ldg(reinterpret_cast<const float4*>(Wa + 3 * C + c));
        a00[i] = a01[i] = a10[i] = a11[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float b0 = ba[0], b1 = ba[1];
    float sdl0 = 0.f, sdl1 = 0.f;
    const int row0 = (blockIdx.x * (blockDim.x >> 5) + warp) * FUSE_ROWS;
    for (int rr = 0; rr < FUSE_ROWS; ++rr) {
        const int row = row0 + rr;
        if (row >= M) break;
        const size_t base = static_cast<size_t>(row) * C;
        float4 a[NV], b[NV], g[NV];
        float l0 = 0.f, l1 = 0.f, da0 = 0.f, da1 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = 128 * i + 4 * lane;
            a[i] = *reinterpret_cast<const float4*>(xst + base + c);
            b[i] = *reinterpret_cast<const float4*>(xts + base + c);
            g[i] = *reinterpret_cast<const float4*>(dx + base + c);
            l0 += a[i].x * w0a[i].x + a[i].y * w0a[i].y + a[i].z * w0a[i].z + a[i].w * w0a[i].w + b[i].x * w0b[i].x + b[i].y * w0b[i].y + b[i].z * w0b[i].z + b[i].w * w0b[i].w;
            l1 += a[i].x * w1a[i].x + a[i].y * w1a[i].y + a[i].z * w1a[i].z + a[i].w * w1a[i].w + b[i].x * w1b[i].x + b[i].y * w1b[i].y + b[i].z * w1b[i].z + b[i].w * w1b[i].w;
            da0 += g[i].x * a[i].x + g[i].y * a[i].y + g[i].z * a[i].z + g[i].w * a[i].w;
            da1 += g[i].x * b[i].x + g[i].y * b[i].y + g[i].z * b[i].z + g[i].w * b[i].w;
        }
        l0 = warp_sum(l0) + b0;
        l1 = warp_sum(l1) + b1;
        da0 = warp_sum(da0);
        da1 = warp_sum(da1);
        const float mx = fmaxf(l0, l1);
        const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
        const float inv = 1.0f / (e0 + e1);
        const float al0 = e0 * inv, al1 = e1 * inv;
        const float dot = al0 * da0 + al1 * da1;
        const float dl0 = al0 * (da0 - dot), dl1 = al1 * (da1 - dot);       // d logits
        sdl0 += dl0;
        sdl1 += dl1;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = 128 * i + 4 * lane;
            const float4 ds = make_float4(al0 * g[i].x + dl0 * w0a[i].x + dl1 * w1a[i].x, al0 * g[i].y + dl0 * w0a[i].y + dl1 * w1a[i].y,
                                          al0 * g[i].z + dl0 * w0a[i].z + dl1 * w1a[i].z, al0 * g[i].w + dl0 * w0a[i].w + dl1 * w1a[i].w);
            const float4 dt = make_float4(al1 * g[i].x + dl0 * w0b[i].x + dl1 * w1b[i].x, al1 * g[i].y + dl0 * w0b[i].y + dl1 * w1b[i].y,
                                          al1 * g[i].z + dl0 * w0b[i].z + dl1 * w1b[i].z, al1 * g[i].w + dl0 * w0b[i].w + dl1 * w1b[i].w);
            *reinterpret_cast<float4*>(dxst + base + c) = ds;
            *reinterpret_cast<float4*>(dxts + base + c) = dt;
            *reinterpret_cast<uint2*>(dxst_plane + base + c) = make_uint2(pack2(ds.x, ds.y), pack2(ds.z, ds.w));
            *reinterpret_cast<uint2*>(dxts_plane + base + c) = make_uint2(pack2(dt.x, dt.y), pack2(dt.z, dt.w));
            a00[i].x = fmaf(dl0, a[i].x, a00[i].x); a00[i].y = fmaf(dl0, a[i].y, a00[i].y); a00[i].z = fmaf(dl0, a[i].z, a00[i].z); a00[i].w = fmaf(dl0, a[i].w, a00[i].w);
            a01[i].x = fmaf(dl0, b[i].x, a01[i].x); a01[i].y = fmaf(dl0, b[i].y, a01[i].y); a01[i].z = fmaf(dl0, b[i].z, a01[i].z); a01[i].w = fmaf(dl0, b[i].w, a01[i].w);
            a10[i].x = fmaf(dl1, a[i].x, a10[i].x); a10[i].y = fmaf(dl1, a[i].y, a10[i].y); a10[i].z = fmaf(dl1, a[i].z, a10[i].z); a10[i].w = fmaf(dl1, a[i].w, a10[i].w);
            a11[i].x = fmaf(dl1, b[i].x, a11[i].x); a11[i].y = fmaf(dl1, b[i].y, a11[i].y); a11[i].z = fmaf(dl1, b[i].z, a11[i].z); a11[i].w = fmaf(dl1, b[i].w, a11[i].w);
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = 128 * i + 4 * lane;
        const float4* srcs[4] = {&a00[i], &a01[i], &a10[i], &a11[i]};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            atomicAdd(&s_acc[q * C + c + 0], srcs[q]->x);
            atomicAdd(&s_acc[q * C + c + 1], srcs[q]->y);
            atomicAdd(&s_acc[q * C + c + 2], srcs[q]->z);
            atomicAdd(&s_acc[q * C + c + 3], srcs[q]->w);
        }
    }
    if (lane == 0) {
        atomicAdd(&s_acc[4 * C + 0], sdl0);
        atomicAdd(&s_acc[4 * C + 1], sdl1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) atomicAdd(dWa + i, s_acc[i]);
    if (threadIdx.x < 2) atomicAdd(dba + threadIdx.x, s_acc[4 * C + threadIdx.x]);
}

// head backward (DSTformer.py:357): out = rep Wh^T + bh  ->  d_rep (+)= d_out Wh ; dWh += d_out^T rep ; dbh += sum d_out
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ rep,
                                                        const float* __restrict__ Wh, int M, int R, int dim_out,
                                                        const float* __restrict__ drep_in, float* __restrict__ drep,
                                                        float* __restrict__ dWh, float* __restrict__ dbh) {
    // one CTA = 64 rows; thread t owns rep column(s) t, t+256 ...; dWh accumulated in registers over the CTA's rows
    const int m0 = blockIdx.x * 64;
    const int m1 = m0 + 64 < M ? m0 + 64 : M;
    for (int r = threadIdx.x; r < R; r += 256) {
        float acc[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] = 0.f;
        for (int m = m0; m < m1; ++m) {
            const float rv = rep[static_cast<size_t>(m) * R + r];
            float d = drep_in ? drep_in[static_cast<size_t>(m) * R + r] : 0.f;
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                if (o < dim_out && dout) {
                    const float go = dout[static_cast<size_t>(m) * dim_out + o];
                    d = fmaf(go, Wh[static_cast<size_t>(o) * R + r], d);
                    acc[o] = fmaf(go, rv, acc[o]);
                }
            }
            drep[static_cast<size_t>(m) * R + r] = d;
        }
        if (dout) {
#pragma unroll
            for (int o = 0; o < 8; ++o)
                if (o < dim_out) atomicAdd(dWh + static_cast<size_t>(o) * R + r, acc[o]);
        }
    }
    if (dout && threadIdx.x < dim_out) {
        float t = 0.f;
        for (int m = m0; m < m1; ++m) t += dout[static_cast<size_t>(m) * dim_out + threadIdx.x];
        atomicAdd(dbh + threadIdx.x, t);
    }
}

// embed backward (DSTformer.py:333-337): x0 = xin We^T + be + pos[j] + temp[f]
//   dWe[c, k] += sum_m dx[m, c] xin[m, k] ; dbe[c] += sum_m dx[m, c] ; dpos[j, c] += sum_{b,f} dx ; dtemp[f, c] += sum_{b,j} dx
// One CTA = one frame index f and a slab of EMB_BATCH clips; thread = channel; per-joint sums live in registers, so
// the global atomics are (J + 2 + dim_in) * C per CTA.
constexpr int EMB_BATCH = 16;
constexpr int EMB_MAXJ = 32;
__global__ void __launch_bounds__(256) embed_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ xin, int dim_in,
                                                         int B, int F, int J, int C, float* __restrict__ dWe,
                                                         float* __restrict__ dbe, float* __restrict__ dpos,
                                                         float* __restrict__ dtemp) {
    const int f = blockIdx.x;
    const int b0 = blockIdx.y * EMB_BATCH;
    const int b1 = b0 + EMB_BATCH < B ? b0 + EMB_BATCH : B;
    for (int c = threadIdx.x; c < C; c += 256) {
        float sp[EMB_MAXJ], sw[8], st = 0.f;
#pragma unroll
        for (int j = 0; j < EMB_MAXJ; ++j) sp[j] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) sw[k] = 0.f;
        for (int b = b0; b < b1; ++b) {
#pragma unroll
            for (int j = 0; j < EMB_MAXJ; ++j) {
                if (j < J) {
                    const size_t m = (static_cast<size_t>(b) * F + f) * J + j;
                    const float g = dx[m * C + c];
                    sp[j] += g;
                    st += g;
                    #pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k < dim_in) sw[k] = fmaf(g, xin[m * dim_in + k], sw[k]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < EMB_MAXJ; ++j)
            if (j < J) atomicAdd(dpos + static_cast<size_t>(j) * C + c, sp[j]);
        atomicAdd(dtemp + static_cast<size_t>(f) * C + c, st);
        atomicAdd(dbe + c, st);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < dim_in) atomicAdd(dWe + static_cast<size_t>(c) * dim_in + k, sw[k]);
    }
}

}  // namespace mb
