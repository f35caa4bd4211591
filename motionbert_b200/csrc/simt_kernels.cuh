// CUDA-core kernels of the DSTformer path: everything that is not GEMM-shaped
// (embed, S/T fusion, head, 17-joint spatial attention, weight packing), plus slow
// bring-up/reference versions of the tensor-core kernels that tests use to localise bugs
// on the device (gemm_ref, attn_t_ref).  All HBM access is 128-bit and row-contiguous.
#pragma once
#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace mb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// One warp owns one token row of C channels held as float4 per lane per 128-channel slab
// (channels 128*i + 4*lane .. +3).  Writes fp32 x, bf16 hi/lo planes and the per-128-group
// LayerNorm partials consumed by the next GEMM's LN-folded epilogue.
// F16C: `hi` is an F16C row buffer (ptx.cuh: 4 bytes per element, 128-byte blocks of 32 elements), `lo` is unused.
template <int MAXV, bool F16C = false>
__device__ __forceinline__ void emit_row(const float4 (&v)[MAXV], int nv, size_t row, int C, float* __restrict__ x,
                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                         float* __restrict__ stats) {
    const int lane = lane_id();
    const size_t base = row * static_cast<size_t>(C);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        if (i < nv) {
            const int c = 128 * i + 4 * lane;
            if (x) *reinterpret_cast<float4*>(x + base + c) = v[i];
            if (F16C) {
                uint8_t* blk = reinterpret_cast<uint8_t*>(hi) + base * 4 + static_cast<size_t>(c >> 5) * 128;
                const int e = c & 31;
                uint32_t h0, l0, g0, h1, l1, g1;
                split2_f16c(v[i].x, v[i].y, h0, l0, g0);
                split2_f16c(v[i].z, v[i].w, h1, l1, g1);
                *reinterpret_cast<uint2*>(blk + 2 * e) = make_uint2(h0, h1);
                *reinterpret_cast<uint32_t*>(blk + 64 + e) = l0 | (l1 << 16);
                *reinterpret_cast<uint32_t*>(blk + 96 + e) = g0 | (g1 << 16);
            } else {
                uint32_t h0, l0, h1, l1;
                split2(v[i].x, v[i].y, h0, l0);
                split2(v[i].z, v[i].w, h1, l1);
                *reinterpret_cast<uint2*>(hi + base + c) = make_uint2(h0, h1);
                if (lo) *reinterpret_cast<uint2*>(lo + base + c) = make_uint2(l0, l1);
            }
        }
    }
    if (stats) {
        // one 128-channel slab (= one float4 per lane) is exactly one statistics group
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            if (i < nv) {
                const float shift = __shfl_sync(0xffffffffu, v[i].x, 0);
                const float d0 = v[i].x - shift, d1 = v[i].y - shift, d2 = v[i].z - shift, d3 = v[i].w - shift;
                float s1 = (d0 + d1) + (d2 + d3);
                float s2 = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
                s1 = warp_sum(s1);
                s2 = warp_sum(s2);
                if (lane == 0) {
                    float* so = stats + (row * nv + i) * 3;
                    so[0] = shift;
                    so[1] = s1;
                    so[2] = s2;
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// embed (DSTformer.py:330-337): joints_embed Linear(3->C) + pos_embed[j] + temp_embed[f]
// ---------------------------------------------------------------------------------------------
template <int NV, bool F16C = false>
__global__ void __launch_bounds__(256) embed_kernel(const float* __restrict__ xin, int dim_in,
                                                     const float* __restrict__ W,      // [C, dim_in]
                                                     const float* __restrict__ bias,   // [C]
                                                     const float* __restrict__ pos,    // [J, C]
                                                     const float* __restrict__ temp,   // [maxlen, C]
                                                     int M, int F, int J, int C, float* __restrict__ x,
                                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                     float* __restrict__ stats) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = lane_id();
    const int j = row % J;
    const int f = (row / J) % F;
    float in[8];
    for (int k = 0; k < dim_in && k < 8; ++k) in[k] = xin[static_cast<size_t>(row) * dim_in + k];
    float4 v[NV];
    constexpr int nv = NV;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (i < nv) {
            const int c = 128 * i + 4 * lane;
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float a = 0.f;
                for (int k = 0; k < dim_in && k < 8; ++k) a = fmaf(in[k], W[(c + e) * dim_in + k], a);
                o[e] = a + bias[c + e];
            }
            const float4 pe = *reinterpret_cast<const float4*>(pos + static_cast<size_t>(j) * C + c);
            const float4 te = *reinterpret_cast<const float4*>(temp + static_cast<size_t>(f) * C + c);
            v[i] = make_float4((o[0] + pe.x) + te.x, (o[1] + pe.y) + te.y, (o[2] + pe.z) + te.z,
                               (o[3] + pe.w) + te.w);
        }
    }
    emit_row<NV, F16C>(v, nv, row, C, x, hi, lo, stats);
}

// ---------------------------------------------------------------------------------------------
// S/T stream fusion (DSTformer.py:343-349): alpha = softmax(Linear(2C->2)(cat[x_st, x_ts])) per token
// ---------------------------------------------------------------------------------------------
template <int NV, bool F16C = false>
__global__ void __launch_bounds__(256, 4) fuse_kernel(const float* __restrict__ xst, const float* __restrict__ xts,
                                                    const float* __restrict__ Wa,   // [2, 2C]
                                                    const float* __restrict__ ba,   // [2]
                                                    int M, int C, float* __restrict__ x,
                                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                    float* __restrict__ stats) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = lane_id();
    constexpr int nv = NV;
    const size_t base = static_cast<size_t>(row) * C;
    float4 a[NV], b[NV];
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (i < nv) {
            const int c = 128 * i + 4 * lane;
            a[i] = *reinterpret_cast<const float4*>(xst + base + c);
            b[i] = *reinterpret_cast<const float4*>(xts + base + c);
            const float4 w0a = __ldg(reinterpret_cast<const float4*>(Wa + c));
            const float4 w0b = __ldg(reinterpret_cast<const float4*>(Wa + C + c));
            const float4 w1a = __ldg(reinterpret_cast<const float4*>(Wa + 2 * C + c));
            const float4 w1b = __ldg(reinterpret_cast<const float4*>(Wa + 3 * C + c));
            d0 += a[i].x * w0a.x + a[i].y * w0a.y + a[i].z * w0a.z + a[i].w * w0a.w + b[i].x * w0b.x +
                  b[i].y * w0b.y + b[i].z * w0b.z + b[i].w * w0b.w;
            d1 += a[i].x * w1a.x + a[i].y * w1a.y + a[i].z * w1a.z + a[i].w * w1a.w + b[i].x * w1b.x +
                  b[i].y * w1b.y + b[i].z * w1b.z + b[i].w * w1b.w;
        }
    }
    d0 = warp_sum(d0) + ba[0];
    d1 = warp_sum(d1) + ba[1];
    const float mx = fmaxf(d0, d1);
    const float e0 = expf(d0 - mx), e1 = expf(d1 - mx);
    const float inv = 1.0f / (e0 + e1);
    const float al0 = e0 * inv, al1 = e1 * inv;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (i < nv)
            v[i] = make_float4(a[i].x * al0 + b[i].x * al1, a[i].y * al0 + b[i].y * al1, a[i].z * al0 + b[i].z * al1,
                               a[i].w * al0 + b[i].w * al1);
    }
    emit_row<NV, F16C>(v, nv, row, C, x, hi, lo, stats);
}

// ---------------------------------------------------------------------------------------------
// head (DSTformer.py:357): out[m, o] = rep[m, :] . Wh[o, :] + bh[o]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ rep, const float* __restrict__ Wh,
                                                    const float* __restrict__ bh, int M, int R, int dim_out,
                                                    float* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = lane_id();
    const float* r = rep + static_cast<size_t>(row) * R;
    for (int o = 0; o < dim_out; ++o) {
        float acc = 0.f;
        for (int c = 4 * lane; c < R; c += 128) {
            const float4 x = *reinterpret_cast<const float4*>(r + c);
            const float4 w = __ldg(reinterpret_cast<const float4*>(Wh + static_cast<size_t>(o) * R + c));
            acc += x.x * w.x + x.y * w.y + x.z * w.z + x.w * w.w;
        }
        acc = warp_sum(acc);
        if (lane == 0) out[static_cast<size_t>(row) * dim_out + o] = acc + bh[o];
    }
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 nn.Linear weight [N,K] -> bf16 hi/lo planes (+ LayerNorm fold)
//   LN fold:  LN(x) W^T + b  =  rstd * (x W'^T - mean * s) + c,  W' = W * gamma,  s[n] = sum_k W'[n,k]
//             (s uses the *split* W' so the mean term cancels exactly),  c[n] = sum_k beta[k] W[n,k] + b[n]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_linear_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                                           const float* __restrict__ gamma,   // null: no LN fold
                                                           const float* __restrict__ beta, int N, int K,
                                                           __nv_bfloat16* __restrict__ hi,
                                                           __nv_bfloat16* __restrict__ lo, float* __restrict__ vec_c,
                                                           float* __restrict__ vec_s, int f16c) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= N) return;
    const int lane = lane_id();
    float s = 0.f, c = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float w = W[static_cast<size_t>(n) * K + k];
        const float wp = gamma ? w * gamma[k] : w;
        if (f16c) {
            // F16C rows [N][K]: block k/32 of row n, element `lane` (k % 32 == lane): f16 | lo8 | hi8
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

// fp32 [M,K] -> hi/lo planes + LN partial stats (used by the test hooks to feed the GEMM)
template <int NV, bool F16C = false>
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ xin, int M, int C,
                                                          __nv_bfloat16* __restrict__ hi,
                                                          __nv_bfloat16* __restrict__ lo, float* __restrict__ stats) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = lane_id();
    constexpr int nv = NV;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i)
        v[i] = *reinterpret_cast<const float4*>(xin + static_cast<size_t>(row) * C + 128 * i + 4 * lane);
    emit_row<NV, F16C>(v, nv, row, C, nullptr, hi, lo, stats);
}

#ifdef MB_TEST_KERNELS   // CUDA-core bring-up / reference kernels: libmotionbert_b200_test.so only
// ---------------------------------------------------------------------------------------------
// Spatial attention (DSTformer.py:178-186): per (frame, head) softmax(q k^T d^-1/2) v over J joints.
// One CTA per frame, one warp per head (looped), lane i < J owns query row i: q row and the output
// row live in registers, k/v rows are smem broadcasts (conflict-free).  fp32 math on hi+lo inputs.
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256) attn_s_kernel(const __nv_bfloat16* __restrict__ qkv_hi,
                                                      const __nv_bfloat16* __restrict__ qkv_lo,   // may be null
                                                      int BF, int J, int C, int H, float scale,
                                                      __nv_bfloat16* __restrict__ out_hi,
                                                      __nv_bfloat16* __restrict__ out_lo) {
    extern __shared__ float sm[];   // [J][3C]
    const int frame = blockIdx.x;
    const int tid = threadIdx.x;
    const int C3 = 3 * C;
    const size_t tok0 = static_cast<size_t>(frame) * J;
    {
        const uint4* gh = reinterpret_cast<const uint4*>(qkv_hi + tok0 * C3);
        const uint4* gl = qkv_lo ? reinterpret_cast<const uint4*>(qkv_lo + tok0 * C3) : nullptr;
        const int nvec = J * C3 / 8;
        for (int i = tid; i < nvec; i += blockDim.x) {
            const uint4 h = gh[i];
            uint4 l = make_uint4(0, 0, 0, 0);
            if (gl) l = gl[i];
            const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
            const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
            float f[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                f[2 * e] = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
                f[2 * e + 1] = __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
            }
            float4* d = reinterpret_cast<float4*>(sm + static_cast<size_t>(i) * 8);
            d[0] = make_float4(f[0], f[1], f[2], f[3]);
            d[1] = make_float4(f[4], f[5], f[6], f[7]);
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int h = warp; h < H; h += nwarps) {
        if (lane < J) {
            float q[HD];
            const float* qp = sm + static_cast<size_t>(lane) * C3 + h * HD;
#pragma unroll
            for (int d4 = 0; d4 < HD / 4; ++d4) {
                const float4 t = *reinterpret_cast<const float4*>(qp + 4 * d4);
                q[4 * d4] = t.x; q[4 * d4 + 1] = t.y; q[4 * d4 + 2] = t.z; q[4 * d4 + 3] = t.w;
            }
            float s[32];
            float mx = -INFINITY;
#pragma unroll
            for (int kj = 0; kj < 32; ++kj) {
                if (kj < J) {
                    const float* kp = sm + static_cast<size_t>(kj) * C3 + C + h * HD;
                    float a = 0.f;
#pragma unroll
                    for (int d4 = 0; d4 < HD / 4; ++d4) {
                        const float4 t = *reinterpret_cast<const float4*>(kp + 4 * d4);
                        a = fmaf(q[4 * d4], t.x, a); a = fmaf(q[4 * d4 + 1], t.y, a);
                        a = fmaf(q[4 * d4 + 2], t.z, a); a = fmaf(q[4 * d4 + 3], t.w, a);
                    }
                    s[kj] = a * scale;
                    mx = fmaxf(mx, s[kj]);
                }
            }
            float sum = 0.f;
#pragma unroll
            for (int kj = 0; kj < 32; ++kj)
                if (kj < J) { s[kj] = expf(s[kj] - mx); sum += s[kj]; }
            const float inv = 1.0f / sum;
            float o[HD];
#pragma unroll
            for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
            for (int kj = 0; kj < 32; ++kj) {
                if (kj < J) {
                    const float* vp = sm + static_cast<size_t>(kj) * C3 + 2 * C + h * HD;
                    const float pj = s[kj] * inv;
#pragma unroll
                    for (int d4 = 0; d4 < HD / 4; ++d4) {
                        const float4 t = *reinterpret_cast<const float4*>(vp + 4 * d4);
                        o[4 * d4] = fmaf(pj, t.x, o[4 * d4]); o[4 * d4 + 1] = fmaf(pj, t.y, o[4 * d4 + 1]);
                        o[4 * d4 + 2] = fmaf(pj, t.z, o[4 * d4 + 2]); o[4 * d4 + 3] = fmaf(pj, t.w, o[4 * d4 + 3]);
                    }
                }
            }
            const size_t ob = (tok0 + lane) * C + h * HD;
#pragma unroll
            for (int d8 = 0; d8 < HD / 8; ++d8) {
                uint32_t hh[4], ll[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split2(o[8 * d8 + 2 * e], o[8 * d8 + 2 * e + 1], hh[e], ll[e]);
                *reinterpret_cast<uint4*>(out_hi + ob + 8 * d8) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                if (out_lo) *reinterpret_cast<uint4*>(out_lo + ob + 8 * d8) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Bring-up / test reference for temporal attention (DSTformer.py:188-200), CUDA cores, fp32.
// One CTA per (b, joint, head); K and V of the sequence in smem; thread = query frame.
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128) attn_t_ref_kernel(const __nv_bfloat16* __restrict__ qkv_hi,
                                                          const __nv_bfloat16* __restrict__ qkv_lo, int B, int F,
                                                          int J, int C, int H, float scale,
                                                          __nv_bfloat16* __restrict__ out_hi,
                                                          __nv_bfloat16* __restrict__ out_lo) {
    extern __shared__ float sm[];   // K [F][HD], V [F][HD]
    float* sK = sm;
    float* sV = sm + static_cast<size_t>(F) * HD;
    const int h = blockIdx.x % H;
    const int j = (blockIdx.x / H) % J;
    const int b = blockIdx.x / (H * J);
    const int C3 = 3 * C;
    auto ld = [&](size_t idx) -> float {
        float v = __bfloat162float(qkv_hi[idx]);
        if (qkv_lo) v += __bfloat162float(qkv_lo[idx]);
        return v;
    };
    for (int i = threadIdx.x; i < F * HD; i += blockDim.x) {
        const int t = i / HD, d = i % HD;
        const size_t tok = (static_cast<size_t>(b) * F + t) * J + j;
        sK[i] = ld(tok * C3 + C + h * HD + d);
        sV[i] = ld(tok * C3 + 2 * C + h * HD + d);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < F; t += blockDim.x) {
        const size_t tok = (static_cast<size_t>(b) * F + t) * J + j;
        float q[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) q[d] = ld(tok * C3 + h * HD + d);
        float mx = -INFINITY;
        for (int u = 0; u < F; ++u) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) a = fmaf(q[d], sK[u * HD + d], a);
            mx = fmaxf(mx, a * scale);
        }
        float o[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = 0.f;
        float sum = 0.f;
        for (int u = 0; u < F; ++u) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) a = fmaf(q[d], sK[u * HD + d], a);
            const float pe = expf(a * scale - mx);
            sum += pe;
#pragma unroll
            for (int d = 0; d < HD; ++d) o[d] = fmaf(pe, sV[u * HD + d], o[d]);
        }
        const float inv = 1.0f / sum;
        const size_t ob = tok * C + h * HD;
#pragma unroll
        for (int d2 = 0; d2 < HD / 2; ++d2) {
            uint32_t hh, ll;
            split2(o[2 * d2] * inv, o[2 * d2 + 1] * inv, hh, ll);
            *reinterpret_cast<uint32_t*>(out_hi + ob + 2 * d2) = hh;
            if (out_lo) *reinterpret_cast<uint32_t*>(out_lo + ob + 2 * d2) = ll;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Bring-up / test reference GEMM on CUDA cores with the same fused epilogues as gemm_tc_kernel.
// One warp per (row, 128-column group); lanes stride the group's columns.
// ---------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256) gemm_ref_kernel(const __nv_bfloat16* __restrict__ a_hi,
                                                        const __nv_bfloat16* __restrict__ a_lo,
                                                        const __nv_bfloat16* __restrict__ w_hi,
                                                        const __nv_bfloat16* __restrict__ w_lo, const GemmParams p) {
    const int ngrp = p.N / STATS_GROUP;
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= p.M * ngrp) return;
    const int row = wid / ngrp, grp = wid % ngrp;
    const int lane = lane_id();
    float mean = 0.f, rstd = 1.f, rscale = 1.f;
    if (EPI == EPI_LN_SPLIT || EPI == EPI_LN_GELU_SPLIT || EPI == EPI_LN_TANH_F32)
        ln_row_stats(p.stats_in + static_cast<size_t>(row) * p.nh_in * 3, p.nh_in, p.ln_dim, p.eps, mean, rstd);
    if (EPI == EPI_RESID && p.row_scale) rscale = p.row_scale[row / p.J];
    float vals[STATS_GROUP / 32];
    for (int i = 0; i < STATS_GROUP / 32; ++i) {
        const int n = grp * STATS_GROUP + i * 32 + lane;
        float acc = 0.f;
        for (int k = 0; k < p.K; ++k) {
            float a = __bfloat162float(a_hi[static_cast<size_t>(row) * p.K + k]);
            float w = __bfloat162float(w_hi[static_cast<size_t>(n) * p.K + k]);
            if (a_lo) a += __bfloat162float(a_lo[static_cast<size_t>(row) * p.K + k]);
            if (w_lo) w += __bfloat162float(w_lo[static_cast<size_t>(n) * p.K + k]);
            acc = fmaf(a, w, acc);
        }
        float v;
        if (EPI == EPI_RESID) v = p.resid[static_cast<size_t>(row) * p.N + n] + rscale * (acc + p.vec0[n]);
        else if (EPI == EPI_BIAS_F32) v = acc + p.vec0[n];
        else {
            v = fmaf(rstd, acc, fmaf(-mean * rstd, p.vec1[n], p.vec0[n]));
            if (EPI == EPI_LN_GELU_SPLIT) v = gelu_erf(v);
            if (EPI == EPI_LN_TANH_F32) v = tanhf(v);
        }
        vals[i] = v;
        const size_t o = static_cast<size_t>(row) * p.N + n;
        if (EPI == EPI_RESID || EPI == EPI_LN_TANH_F32 || EPI == EPI_BIAS_F32) p.out_f32[o] = v;
        if ((EPI == EPI_RESID || EPI == EPI_LN_SPLIT || EPI == EPI_LN_GELU_SPLIT) && p.out_hi) {
            __nv_bfloat16 h, l;
            split_bf16(v, h, l);
            p.out_hi[o] = h;
            if (p.out_lo) p.out_lo[o] = l;
        }
    }
    if (EPI == EPI_RESID && p.stats_out) {
        const float shift = __shfl_sync(0xffffffffu, vals[0], 0);
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < STATS_GROUP / 32; ++i) {
            const float d = vals[i] - shift;
            s1 += d;
            s2 = fmaf(d, d, s2);
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) {
            float* so = p.stats_out + (static_cast<size_t>(row) * ngrp + grp) * 3;
            so[0] = shift; so[1] = s1; so[2] = s2;
        }
    }
}

#endif  // MB_TEST_KERNELS

}  // namespace mb
