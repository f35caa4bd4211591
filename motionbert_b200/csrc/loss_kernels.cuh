// Pretrain-step losses fused into one pass over the (B, T, J, 3) pose output:
//   loss_mpjpe        (lib/model/loss.py:56-63)    L1 = mean_{b,t,j} |p - g|
//   n_mpjpe           (lib/model/loss.py:80-89)    L2 = mean |s_bt p - g|,  s_bt = sum_j <g,p> / sum_j <p,p>
//   loss_velocity     (lib/model/loss.py:133-142)  L3 = mean_{b,t>=1,j} |(p_t - p_{t-1}) - (g_t - g_{t-1})|
//   loss_2d_weighted  (lib/model/loss.py:73-78)    L  = mean |(p_xy - g_xy) * conf|
// and the gradient of  total = L1 + lambda_scale L2 + lambda_velocity L3  (train.py:178-191) w.r.t. p, analytically,
// in the same launch -- the reference reads the output ~7 times, launches ~40 elementwise kernels and synchronises 8
// times per step for .item() (train.py:192-199).  One warp per (b, t) frame, lane = joint (J <= 32).
#pragma once
#include "simt_kernels.cuh"

namespace mb {

struct PoseLossParams {
    const float* pred;      // (B, T, J, 3)
    const float* target;    // (B, T, J, 3)
    const float* conf;      // (B, T, J) or null: 2-D re-projection mode when non-null
    int B, T, J;
    float lambda_scale, lambda_velocity;
    double* acc;            // [4] zero-initialised accumulators: sum|e1|, sum|e2|, sum|dv|, unused
    float* d_pred;          // (B, T, J, 3) gradient of `total` (or of the 2-D loss), may be null
};

__device__ __forceinline__ float3 ld3(const float* p) { return make_float3(p[0], p[1], p[2]); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// unit vector of v and its length; the zero vector maps to zero (torch.norm's sub-gradient at 0)
__device__ __forceinline__ float3 unit3(float3 v, float& n) {
    n = sqrtf(dot3(v, v));
    const float inv = n > 0.f ? 1.0f / n : 0.f;
    return make_float3(v.x * inv, v.y * inv, v.z * inv);
}

__global__ void __launch_bounds__(256) pose_loss_kernel(const PoseLossParams p) {
    const int frame = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = lane_id();
    const int nframes = p.B * p.T;
    float l1 = 0.f, l2 = 0.f, l3 = 0.f;
    if (frame < nframes) {
        const int t = frame % p.T;
        const bool jv = lane < p.J;
        const size_t row = static_cast<size_t>(frame) * p.J + (jv ? lane : 0);
        const float3 pj = jv ? ld3(p.pred + row * 3) : make_float3(0.f, 0.f, 0.f);
        const float3 gj = jv ? ld3(p.target + row * 3) : make_float3(0.f, 0.f, 0.f);
        float3 grad = make_float3(0.f, 0.f, 0.f);
        if (p.conf) {
            // 2-D re-projection loss on (x, y), weighted by the detector confidence
            const float c = jv ? p.conf[row] : 0.f;
            const float3 e = make_float3((pj.x - gj.x) * c, (pj.y - gj.y) * c, 0.f);
            float n;
            const float3 u = unit3(e, n);
            l1 = jv ? n : 0.f;
            const float w = 1.0f / (static_cast<float>(nframes) * p.J);
            grad = make_float3(u.x * c * w, u.y * c * w, 0.f);
        } else {
            const float inv_n = 1.0f / (static_cast<float>(nframes) * p.J);
            // L1
            float n1;
            const float3 u1 = unit3(make_float3(pj.x - gj.x, pj.y - gj.y, pj.z - gj.z), n1);
            l1 = jv ? n1 : 0.f;
            // L2: per-frame scale
            const float a = warp_sum(jv ? dot3(gj, pj) : 0.f);
            const float c = warp_sum(jv ? dot3(pj, pj) : 0.f);
            const float s = a / c;
            float n2;
            const float3 u2 = unit3(make_float3(s * pj.x - gj.x, s * pj.y - gj.y, s * pj.z - gj.z), n2);
            l2 = jv ? n2 : 0.f;
            const float q = warp_sum(jv ? dot3(u2, pj) : 0.f) / c;          // sum_k <u_k, p_k> / c
            grad.x = inv_n * (u1.x + p.lambda_scale * (s * u2.x + q * (gj.x - 2.f * s * pj.x)));
            grad.y = inv_n * (u1.y + p.lambda_scale * (s * u2.y + q * (gj.y - 2.f * s * pj.y)));
            grad.z = inv_n * (u1.z + p.lambda_scale * (s * u2.z + q * (gj.z - 2.f * s * pj.z)));
            // L3: velocity error of (t-1 -> t) is accounted to frame t; its gradient touches frames t and t-1
            if (p.T > 1) {
                const float inv_n3 = p.lambda_velocity / (static_cast<float>(p.B) * (p.T - 1) * p.J);
                if (t >= 1 && jv) {
                    const float3 pp = ld3(p.pred + (row - p.J) * 3), gp = ld3(p.target + (row - p.J) * 3);
                    float n3;
                    const float3 w = unit3(make_float3((pj.x - pp.x) - (gj.x - gp.x), (pj.y - pp.y) - (gj.y - gp.y),
                                                       (pj.z - pp.z) - (gj.z - gp.z)), n3);
                    l3 = n3;
                    grad.x += inv_n3 * w.x; grad.y += inv_n3 * w.y; grad.z += inv_n3 * w.z;
                }
                if (t + 1 < p.T && jv) {
                    const float3 pn = ld3(p.pred + (row + p.J) * 3), gn = ld3(p.target + (row + p.J) * 3);
                    float n3;
                    const float3 w = unit3(make_float3((pn.x - pj.x) - (gn.x - gj.x), (pn.y - pj.y) - (gn.y - gj.y),
                                                       (pn.z - pj.z) - (gn.z - gj.z)), n3);
                    grad.x -= inv_n3 * w.x; grad.y -= inv_n3 * w.y; grad.z -= inv_n3 * w.z;
                }
            }
        }
        if (p.d_pred && jv) {
            float* d = p.d_pred + row * 3;
            d[0] = grad.x; d[1] = grad.y; d[2] = grad.z;
        }
    }
    // block-level sums -> one double atomic per quantity per CTA
    __shared__ float s_red[3][8];
    l1 = warp_sum(l1); l2 = warp_sum(l2); l3 = warp_sum(l3);
    const int warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[0][warp] = l1; s_red[1][warp] = l2; s_red[2][warp] = l3; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += static_cast<double>(s_red[threadIdx.x][w]);
        atomicAdd(p.acc + threadIdx.x, t);
    }
}

// losses[0..3] = (L1 | 2-D loss, L2, L3, total)
__global__ void pose_loss_finalize_kernel(const double* acc, int B, int T, int J, int mode_2d, float lambda_scale,
                                          float lambda_velocity, float* losses) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double n = static_cast<double>(B) * T * J;
    const double l1 = acc[0] / n;
    if (mode_2d) {
        losses[0] = static_cast<float>(l1); losses[1] = 0.f; losses[2] = 0.f; losses[3] = static_cast<float>(l1);
        return;
    }
    const double l2 = acc[1] / n;
    const double l3 = T > 1 ? acc[2] / (static_cast<double>(B) * (T - 1) * J) : 0.0;
    losses[0] = static_cast<float>(l1);
    losses[1] = static_cast<float>(l2);
    losses[2] = static_cast<float>(l3);
    losses[3] = static_cast<float>(l1 + lambda_scale * l2 + lambda_velocity * l3);
}

}  // namespace mb

// ---------------------------------------------------------------------------------------------
// Augmenter2D (lib/data/augmentation.py:29-74; called in train.py:162-172 right before the encoder):
//   add_noise: per (clip, 27 key frames, joint) a displacement drawn from a per-joint Gaussian (probability weight[j]) or a
//              uniform range, interpolated linearly over the F frames (F.interpolate trilinear, align_corners=True: only
//              the frame axis changes size), plus per-(frame, joint) jitter; the detector confidence is re-synthesised
//              from the displacement length: conf = clip(a / (d + a) + b d + (shift s + m), 0, 1);
//   add_mask : per-(clip, frame, joint) and per-frame Bernoulli masks multiply (x, y, conf).
// The reference runs ~25 elementwise / interpolate kernels over (B, F, J, .) tensors; here ONE pass reads the clip and the
// random draws (made by the caller in the reference's own order, so a seed reproduces the reference's augmentation) and
// writes the augmented (B, F, J, 3) clip.  One thread per (b, f, j).
// ---------------------------------------------------------------------------------------------
struct Augment2DParams {
    const float* x;            // (B, F, J, cin) input clip, cin >= 2 (only x, y are read when noise is on)
    int cin;
    int B, F, J, K;            // K = key frames of the noise model (27)
    int do_noise, do_mask;
    const float* sel;          // (B, K, J)     U[0,1)
    const float* gauss;        // (B, K, J, 2)  N(0,1)
    const float* unif;         // (B, K, J, 2)  U[0,1)
    const float* jitter;       // (F, J, 2)     N(0,1)
    const float* shift;        // (B, F, J)     N(0,1)
    const float* mean;         // (J, 2)
    const float* stdv;         // (J, 2)
    const float* weight;       // (J)
    float uniform_range, noise_std, a, b, m, s;
    const float* mask_u;       // (B, F, J) U[0,1)
    const float* maskT_u;      // (F)       U[0,1)
    float mask_ratio, mask_T_ratio;
    float* out;                // (B, F, J, 3)
};

__global__ void __launch_bounds__(256) augment2d_kernel(const Augment2DParams p) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t n = static_cast<size_t>(p.B) * p.F * p.J;
    if (i >= n) return;
    const int j = static_cast<int>(i % p.J);
    const int f = static_cast<int>((i / p.J) % p.F);
    const int b = static_cast<int>(i / (static_cast<size_t>(p.J) * p.F));
    float x0 = p.x[i * p.cin], x1 = p.x[i * p.cin + 1];
    float c = p.cin > 2 ? p.x[i * p.cin + 2] : 1.f;
    if (p.do_noise) {
        // ATen upsample_trilinear3d, align_corners=True: src = f * (K-1)/(F-1); lambda in fp32
        const float scale = p.F > 1 ? static_cast<float>(p.K - 1) / static_cast<float>(p.F - 1) : 0.f;
        const float src = scale * static_cast<float>(f);
        const int k0 = static_cast<int>(src);
        const int k1 = k0 + (k0 < p.K - 1 ? 1 : 0);
        const float l1 = src - static_cast<float>(k0), l0 = 1.0f - l1;
        const float w = p.weight[j];
        float d[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float v[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int k = q == 0 ? k0 : k1;
                const size_t kj = (static_cast<size_t>(b) * p.K + k) * p.J + j;
                const float g = p.gauss[kj * 2 + e] * p.stdv[j * 2 + e] + p.mean[j * 2 + e];
                const float u = (p.unif[kj * 2 + e] - 0.5f) * p.uniform_range;
                const bool pick_g = p.sel[kj] < w;
                v[q] = g * (pick_g ? 1.f : 0.f) + u * (pick_g ? 0.f : 1.f);
            }
            d[e] = (l0 * v[0] + l1 * v[1]) + (p.jitter[(static_cast<size_t>(f) * p.J + j) * 2 + e] * p.noise_std + 0.f);
        }
        x0 += d[0];
        x1 += d[1];
        const float dis = sqrtf(d[0] * d[0] + d[1] * d[1]);
        const float fconf = p.a / (dis + p.a) + p.b * dis;
        c = fminf(fmaxf(fconf + (p.shift[i] * p.s + p.m), 0.f), 1.f);
    }
    if (p.do_mask) {
        const float mk = (p.mask_u[i] > p.mask_ratio ? 1.f : 0.f);
        const float mt = (p.maskT_u[f] > p.mask_T_ratio ? 1.f : 0.f);
        x0 = x0 * mk * mt; x1 = x1 * mk * mt; c = c * mk * mt;
    }
    p.out[i * 3] = x0; p.out[i * 3 + 1] = x1; p.out[i * 3 + 2] = c;
}
