/* motionbert_b200.h -- C ABI of the B200-native DSTformer encoder (libmotionbert_b200.so).
 *
 * Drop-in boundary for ONE hot path of Walter0807/MotionBERT: the forward pass of
 *     lib/model/DSTformer.py:269-361   class DSTformer  (.forward / .get_representation)
 * Everything below replaces, for that path, the chain of torch.nn calls the reference issues
 * (nn.Linear / nn.LayerNorm / nn.GELU / softmax / matmul: DSTformer.py:79-85,138-200,239-249,329-358).
 * The reference has no native/FFI layer of its own; the binding a maintainer adds is
 * the ctypes stub shown in INTEGRATION.md (shipped as motionbert_b200/_lib.py).
 *
 * Conventions
 *   - plain C types only; every device buffer is allocated and owned by the CALLER (PyTorch's caching
 *     allocator in the Python host); the library allocates no device memory and keeps no reference to
 *     activations after a call returns.  Host-side immutable state (TMA tensor maps, offsets) lives in
 *     the opaque handle.
 *   - all work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*);
 *     no internal synchronisation, CUDA-graph capturable.
 *   - return 0 on success, negative MbStatus otherwise; never throws, never exits.
 *     mb_last_error() returns a thread-local message for the last failing call on this thread.
 *   - thread-safe: distinct handles may be used concurrently from different host threads / devices
 *     (nn.DataParallel drives one replica per GPU from its own Python thread).
 *   - there is NO CPU fallback: on a device that is not sm_100 every compute entry point fails with
 *     MB_ERR_ARCH.
 */
#ifndef MOTIONBERT_B200_H_
#define MOTIONBERT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_ABI_VERSION 1

typedef enum MbStatus {
    MB_OK = 0,
    MB_ERR_INVALID = -1,    /* bad argument / unsupported shape (F > maxlen, J mismatch, C % 256 != 0 ...) */
    MB_ERR_NULL = -2,       /* required pointer is NULL                                                    */
    MB_ERR_ALIGN = -3,      /* device pointer not 16-byte aligned                                          */
    MB_ERR_ARCH = -4,       /* device is not compute capability 10.x                                        */
    MB_ERR_CUDA = -5,       /* a CUDA runtime / driver call failed (message has cudaGetErrorString)         */
    MB_ERR_WORKSPACE = -6   /* workspace too small                                                          */
} MbStatus;

/* Arithmetic mode of the GEMM-shaped work (tensor-core operand formats; fp32 accumulate in all modes). */
typedef enum MbMath {
    MB_MATH_BF16X3 = 0,     /* x = hi + lo bf16 split, 3 MMA passes: fp32 parity (~1.5e-5 rel); training forward      */
    MB_MATH_BF16 = 1,       /* single bf16 pass: for the reference's bf16/AMP-style training configs                 */
    MB_MATH_F16C = 2        /* "F16C": x = h + l, one fp16 pass + both cross terms as ONE e5m2 pass (K = 32 per
                             * instruction at twice the rate) = 2 pass-equivalents: fp32 parity (~7e-5 rel, inside the
                             * 1e-3 / 0.1 mm targets) -- the inference default.  Forward only (mb_forward).          */
} MbMath;

/* Mirrors the constructor arguments of DSTformer.__init__ (DSTformer.py:270-273) as the factory passes
 * them (lib/utils/learning.py:83-85).  hidden = int(dim_feat * mlp_ratio) (DSTformer.py:232). */
typedef struct MbDesc {
    int32_t dim_in;       /* 3                                    */
    int32_t dim_out;      /* 3                                    */
    int32_t dim_feat;     /* C: 512 (base) / 256 (Lite)           */
    int32_t dim_rep;      /* 512                                  */
    int32_t depth;        /* 5                                    */
    int32_t num_heads;    /* 8   (head_dim = C / heads in {32,64})*/
    int32_t hidden;       /* 1024                                 */
    int32_t num_joints;   /* 17                                   */
    int32_t maxlen;       /* 243                                  */
    float eps;            /* LayerNorm eps, 1e-6 from the factory */
    float qk_scale;       /* 0 => head_dim ** -0.5 (DSTformer.py:94) */
    int32_t math;         /* MbMath                               */
} MbDesc;

typedef struct MbEncoder MbEncoder;   /* opaque host-side handle */

/* flags for mb_forward */
/* The TEST ONLY kernels are compiled into libmotionbert_b200_test.so only (include/motionbert_b200_test.h); the
 * product library rejects their flags with MB_ERR_INVALID. */
#define MB_FLAG_REF_GEMM   0x1u   /* TEST ONLY: CUDA-core reference GEMM instead of tcgen05           */
#define MB_FLAG_REF_ATTN_T 0x2u   /* TEST ONLY: CUDA-core reference temporal attention                */
#define MB_FLAG_REF_ATTN_S 0x8u   /* TEST ONLY: CUDA-core reference spatial attention                 */
#define MB_FLAG_GEMM_1CTA  0x4u   /* TEST ONLY: first-generation 1-CTA tcgen05 GEMM (LSU epilogue)    */
#define MB_FLAG_ATTN_T_UNPACKED 0x20u /* one-sequence-per-tile temporal kernel even when F <= 32 (both libraries)  */
#define MB_FLAG_ATTN_BF16X3 0x40u /* F16C mode A/B: qkv as bf16 hi/lo planes + the BF16x3 attention kernels (both)  */
#define MB_FLAG_MLP_SPLIT   0x100u /* F16C mode: the MLP sublayer as two GEMM launches (fc1, fc2) instead of the fused kernel */
#define MB_FLAG_MLP_NO_RING 0x200u /* fused MLP kernel: hidden rows indexed by token block (full buffer) instead of the L2 ring */
#define MB_FLAG_MLP_NO_HINT 0x800u /* fused MLP kernel: no L2::evict_last on the hidden stores / loads */

int mb_version(void);
const char* mb_last_error(void);

/* Create / destroy the host-side handle for one module replica on the CURRENT device. */
int mb_create(const MbDesc* desc, MbEncoder** out);
void mb_destroy(MbEncoder* enc);

/* Parameter tree: the 4 + depth*2*24 + 6 + depth*2 tensors of DSTformer.state_dict() in registration
 * order (DSTformer.py:276-311).  mb_param_info lets the host verify names and sizes. */
int mb_param_count(const MbEncoder* enc);
int mb_param_info(const MbEncoder* enc, int index, char* name, int name_cap, int64_t* numel);

/* Packed weights: bf16 hi/lo planes of every nn.Linear weight with the preceding LayerNorm's affine
 * folded in, plus the small fp32 vectors.  Re-run after every optimizer step / load_state_dict.
 *   params : host array of mb_param_count() device pointers (fp32, contiguous, 16-BYTE ALIGNED: the kernels read
 *            parameters with 128-bit accesses -- slices of a coalesced buffer such as nn.DataParallel's broadcast replicas
 *            are not; the Python class hands over aligned copies) in state_dict order. */
int mb_packed_bytes(const MbEncoder* enc, size_t* bytes);
int mb_pack_weights(MbEncoder* enc, const float* const* params, void* packed, void* stream);

/* Scratch for one forward call of shape (B, F, num_joints, dim_in). */
int mb_workspace_bytes(const MbEncoder* enc, int B, int F, size_t* bytes);

/* DSTformer.forward (DSTformer.py:329-358).
 *   x   : (B, F, J, dim_in) fp32 contiguous, device
 *   out : (B, F, J, dim_out) fp32 or NULL   -- forward(x)
 *   rep : (B, F, J, dim_rep) fp32 or NULL   -- get_representation(x) (DSTformer.py:360-361)
 *   drop_path_scale : NULL (eval / rate 0: the shipped configs) or device fp32 [depth*2*4][B*F] per-frame
 *         DropPath factors mask/keep_prob (lib/model/drop.py:17-32), one vector per residual sublayer
 *         in call order (st block: S-attn, S-mlp, T-attn, T-mlp; then ts block: T-attn, T-mlp, S-attn, S-mlp). */
int mb_forward(MbEncoder* enc, const void* packed, const float* x, float* out, float* rep,
               const float* drop_path_scale, void* workspace, size_t workspace_bytes, int B, int F,
               uint32_t flags, void* stream);

/* Same call with HOST buffers (pageable or pinned): H2D of x, forward, D2H of out/rep on `stream`,
 * synchronised before returning.  `workspace` must additionally hold the device copies:
 * mb_workspace_bytes_host() bytes. */
int mb_workspace_bytes_host(const MbEncoder* enc, int B, int F, int want_out, int want_rep, size_t* bytes);
int mb_forward_host(MbEncoder* enc, const void* packed, const float* x_host, float* out_host, float* rep_host,
                    void* workspace, size_t workspace_bytes, int B, int F, uint32_t flags, void* stream);

/* ---- training: forward with saved activations + analytic backward (loss.backward() through DSTformer.forward,
 * train.py:149-176 / lib/model/DSTformer.py:329-358; row a15 of SURVEY.md section 8) ------------------------------
 * mb_forward_train : mb_forward that additionally keeps every residual-stream tensor
 *                    (fp32 + LayerNorm partial statistics, 9 per depth + 1) in the caller's `saved` region of
 *                    mb_saved_bytes() bytes (1024-byte aligned).  `rep` is mandatory (the backward needs it).
 * mb_backward      : gradients of all mb_param_count() parameters for given d_out (B,F,J,dim_out) and/or d_rep
 *                    (B,F,J,dim_rep) (either may be NULL).  bf16 single-pass tensor-core arithmetic with fp32
 *                    accumulation; qkv / hidden activations / attention probabilities are recomputed, not stored.
 *     params : the same device pointers given to mb_pack_weights (fp32, state_dict order, 16-byte aligned);
 *              d_out / d_rep 16-byte aligned as well
 *     grads  : device pointers, same order and sizes, ZERO-FILLED by the caller (the kernels accumulate into them)
 *     x, rep, saved : the input / output / saved region of the matching mb_forward_train call
 *     workspace : mb_backward_workspace_bytes() bytes, 1024-byte aligned (independent of the forward workspace)
 *     drop_path_scale : the SAME vector given to mb_forward_train (or NULL)
 *     d_x : optional (B,F,J,dim_in) gradient w.r.t. the pose input (NULL: not computed; the reference's training
 *           scripts never need it)
 *     phase_events : NULL, or depth + 2 cudaEvent_t handles recorded on `stream` as the gradients of a phase become
 *           final -- [0] tail (norm, pre_logits, head), [1 + k] depth (depth-1-k) (blocks_st / blocks_ts / ts_attn of
 *           that depth), [depth + 1] embed -- so that a data-parallel caller can all-reduce phase k on a side stream
 *           while phase k+1 is still computing (the reference's nn.DataParallel reduces after the whole backward) */
int mb_saved_bytes(const MbEncoder* enc, int B, int F, size_t* bytes);
int mb_forward_train(MbEncoder* enc, const void* packed, const float* x, float* out, float* rep,
                     const float* drop_path_scale, void* saved, size_t saved_bytes, void* workspace,
                     size_t workspace_bytes, int B, int F, uint32_t flags, void* stream);
int mb_backward_workspace_bytes(const MbEncoder* enc, int B, int F, size_t* bytes);
int mb_backward_launch_count(const MbEncoder* enc, int has_drop_path, int want_dx);   /* kernels per mb_backward call */
int mb_backward(MbEncoder* enc, const void* packed, const float* const* params, const float* x, const float* rep,
                const void* saved, size_t saved_bytes, const float* drop_path_scale, const float* d_out,
                const float* d_rep, float* const* grads, float* d_x, void* workspace, size_t workspace_bytes, int B,
                int F, void* const* phase_events, void* stream);

/* ---- pretrain-step losses on the pose output, fused with their gradient ------------
 * 3-D mode (conf == NULL):  losses[0] = loss_mpjpe (lib/model/loss.py:56-63), [1] = n_mpjpe (:80-89),
 *   [2] = loss_velocity (:133-142), [3] = total = [0] + lambda_scale [1] + lambda_velocity [2]  (train.py:178-191);
 * 2-D mode (conf != NULL, (B,T,J) detector confidences): losses[0] = losses[3] = loss_2d_weighted (:73-78).
 * d_pred (optional) receives d total / d pred, computed analytically in the same launch.  All pointers are device
 * pointers; pred / target are (B,T,J,3) fp32 contiguous; scratch = 32 bytes, 8-byte aligned. */
int mb_pretrain_loss(const float* pred, const float* target, const float* conf, int B, int T, int J,
                     float lambda_scale, float lambda_velocity, float* losses, float* d_pred, void* scratch,
                     void* stream);

/* ---- action-recognition tail (lib/model/model_action.py:15-21, :62-71) ---------------
 * `ActionHeadClassification` / `ActionHeadEmbed` consume the representation only through its mean over the T frames.
 * mb_forward_pooled runs the same forward as mb_forward but the tail GEMM's epilogue accumulates
 *     rep_pool[b, j, :] = mean_f tanh(Linear(LN(x)))[b, f, j, :]           (B, J, dim_rep) fp32, 16-byte aligned
 * and the (B, F, J, dim_rep) representation (2.2 GB at B=256, T=243) is never written.  Same workspace as mb_forward. */
int mb_forward_pooled(MbEncoder* enc, const void* packed, const float* x, float* rep_pool, void* workspace,
                      size_t workspace_bytes, int B, int F, uint32_t flags, void* stream);

/* ---- optimizer step (train.py:289 optim.AdamW, train.py:206 optimizer.step()) ---------
 * torch.optim.AdamW's arithmetic (decoupled weight decay, bias-corrected moments) over the encoder's parameter tensors
 * in grouped launches (48 tensors per kernel, pointer table in kernel-parameter space) instead of a 260-tensor walk.
 * All arrays have mb_param_count() entries in mb_param_info() order; active[i] == 0 (or active == NULL: all active)
 * skips tensor i (frozen / no gradient).  t = 1-based step count.  mb_pack_weights re-packs all 81 linears in three
 * grouped launches afterwards. */
int mb_adamw_step(MbEncoder* enc, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const uint8_t* active, int t, float lr, float beta1, float beta2, float eps,
                  float weight_decay, void* stream);

/* ---- Augmenter2D on the GPU in one pass (lib/data/augmentation.py:29-74, called at
 * train.py:162-172 immediately before the encoder) ----------------------------------------------------------------
 * noise != 0: `add_noise` (:29-65).  The caller supplies the random draws in the reference's own order and shapes (so
 *   the reference's seed reproduces the reference's augmentation): sel (B,K,J) ~ U[0,1), gauss (B,K,J,2) ~ N(0,1),
 *   unif (B,K,J,2) ~ U[0,1), jitter (F,J,2) ~ N(0,1), shift (B,F,J) ~ N(0,1); K = 27 key frames; mean/std (J,2) and
 *   weight (J) are params/synthetic_noise.pth, (a, b, m, s) params/d2c_params.pkl, noise_std = 0.002, uniform_range
 *   = 0.06.  Per (b, key frame, joint): delta = sel < weight[j] ? gauss*std+mean : (unif-0.5)*range; linear
 *   interpolation over the F frames (align_corners), + jitter*noise_std; x,y += delta; conf = clip(a/(d+a) + b d +
 *   shift s + m, 0, 1) with d = |delta|.
 * mask != 0: `add_mask` (:67-74): out *= (mask_u (B,F,J) > mask_ratio) * (maskT_u (F) > mask_T_ratio).
 * x: (B,F,J,cin) fp32, cin >= 2 (>= 3 when noise == 0); out: (B,F,J,3) fp32.  All device pointers. */
int mb_augment2d(const float* x, int cin, int B, int F, int J, int K, int noise, int mask, const float* sel,
                 const float* gauss, const float* unif, const float* jitter, const float* shift, const float* mean,
                 const float* stdv, const float* weight, float uniform_range, float noise_std, float a, float b, float m,
                 float s, const float* mask_u, const float* maskT_u, float mask_ratio, float mask_T_ratio, float* out,
                 void* stream);

/* Number of kernels one mb_forward call launches (for bench.py's gpu_launches): 88 for depth 5 in F16C mode (one kernel
 * per MLP sublayer), 108 with MB_FLAG_MLP_SPLIT and in the bf16 modes. */
int mb_forward_launch_count(const MbEncoder* enc, int want_out, uint32_t flags);

/* Per-kernel-class device timing for bench.py's roofline: while enabled, mb_forward brackets every launch
 * with CUDA events on the caller's stream.  mb_profile_read synchronises, returns the accumulated
 * milliseconds and launch counts per class since the last read and resets them.
 * classes: 0 gemm LN->qkv, 1 the MLP sublayer's first launch (F16C: the fused fc1+GELU+fc2+residual kernel = the whole
 *          sublayer; MB_FLAG_MLP_SPLIT / bf16 modes: gemm LN->fc1+GELU), 2 gemm proj+residual (and fc2+residual in the
 *          two-GEMM form), 3 gemm tail (rep), 4 temporal attention, 5 spatial attention, 6 embed, 7 fusion, 8 head.
 *          Not thread-safe; measurement only. */
#define MB_PROFILE_CLASSES 9
int mb_profile_enable(MbEncoder* enc, int on);
int mb_profile_read(MbEncoder* enc, float* ms_by_class, int* launches_by_class);

#ifdef __cplusplus
}
#endif
#endif /* MOTIONBERT_B200_H_ */
