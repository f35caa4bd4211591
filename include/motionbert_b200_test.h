/* Kernel-level test hooks and CUDA-core / first-generation reference kernels of libmotionbert_b200_test.so.
 *
 * The test library is the SAME source as libmotionbert_b200.so compiled with -DMB_TEST_KERNELS: it exports the whole
 * product ABI (include/motionbert_b200.h) plus the hooks below, and it honours the TEST ONLY kernel flags
 * (MB_FLAG_REF_GEMM, MB_FLAG_REF_ATTN_T, MB_FLAG_REF_ATTN_S, MB_FLAG_GEMM_1CTA).  Nothing in the product path loads it. */
#ifndef MOTIONBERT_B200_TEST_H
#define MOTIONBERT_B200_TEST_H

#include "motionbert_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ kernel-level test hooks ----------
 * Exercise one kernel in isolation so tests/ can localise a failure on the device.  Not used by the
 * product path. */

/* y[M,N] = epilogue(A[M,K] . W[N,K]^T): A, W fp32 on device; scratch >= mb_test_linear_scratch_bytes().
 * mode: 0 LN-folded split (returns hi+lo as fp32), 1 LN+GELU split, 2 residual (+stats), 3 LN+tanh, 4 bias.
 * gamma/beta (LN modes) and resid (mode 2) may be NULL otherwise.  stats_out (mode 2): [M][N/128][3]. */
int mb_test_linear_scratch_bytes(int M, int N, int K, size_t* bytes);
int mb_test_linear(int mode, int math, int use_ref /* 0: 2-CTA tcgen05 (product), 1: CUDA-core reference, 2: 1-CTA tcgen05 */, int M, int N, int K, const float* A, const float* W,
                   const float* bias, const float* gamma, const float* beta, const float* resid, float eps,
                   float* y, float* stats_out, void* scratch, size_t scratch_bytes, void* stream);

/* The F16C operand encoder (ptx.cuh) on its own: x fp32 [rows][cols] (cols % 32 == 0) -> out [rows][cols * 4] bytes. */
int mb_test_f16c_encode(const float* x, int rows, int cols, void* out, void* stream);

/* Attention over a fp32 qkv buffer [B*F*J, 3C] -> y fp32 [B*F*J, C].  temporal=1: forward_temporal
 * (DSTformer.py:188-200), 0: forward_spatial (:178-186). */
int mb_test_attention_scratch_bytes(int B, int F, int J, int C, size_t* bytes);
int mb_test_attention(int temporal, int math, int use_ref /* 0 product, 1 CUDA-core ref, 3 unpacked T */, int B, int F, int J, int C, int H, const float* qkv,
                      float* y, void* scratch, size_t scratch_bytes, void* stream);

/* Backward groundwork (not yet wired into a native backward pass):
 * dW[N,K] = G[M,N]^T . X[M,K] with the split-K tcgen05 weight-gradient kernel (both operands MN-major, no transposes).
 * G = dL/dy and X = the layer input, token-major fp32 on the device; dW fp32 (overwritten).  N % 128 == 0, K % 256 == 0. */
int mb_test_wgrad_scratch_bytes(int M, int N, int K, size_t* bytes);
int mb_test_wgrad(int math, int M, int N, int K, const float* G, const float* X, float* dW, void* scratch,
                  size_t scratch_bytes, void* stream);
/* dX[M,K] = G[M,N] . W[N,K]: the production 2-CTA GEMM reading W in its forward layout as an MN-major operand
 * (no transposed weight copy).  N % 64 == 0 (32 for BF16x3), K % 256 == 0. */
/* d(qkv) [M,3C] of the attention core softmax(q k^T d^-1/2) v for a given d(out) [M,C] (fp32 in/out on the device):
 * flash-style tcgen05 backward in bf16 single-pass arithmetic (attn_bwd_tc.cuh). */
int mb_test_attention_backward_scratch_bytes(int B, int F, int J, int C, size_t* bytes);
int mb_test_attention_backward(int temporal, int B, int F, int J, int C, int H, const float* qkv, const float* dO,
                               float* dqkv, void* scratch, size_t scratch_bytes, void* stream);
int mb_test_dgrad_scratch_bytes(int M, int N, int K, size_t* bytes);
int mb_test_dgrad(int math, int M, int N, int K, const float* G, const float* W, float* dX, void* scratch,
                  size_t scratch_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
