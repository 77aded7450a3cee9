/*
 * gftorf_train.h — C ABI of the fused operators either side of the rasterizer in one training
 * iteration of brownvc/gftorf (the "next" rows).  Same rules as gftorf.h: plain C,
 * raw device pointers, caller-owned memory, caller's stream, <0 = failure + gft_last_error().
 *
 * The reference implements these steps as chains of PyTorch operators (there is no native code to
 * bind), so each entry point cites the Python lines it replaces:
 *
 *   gft_assemble_forward / _backward   gaussian_renderer/__init__.py:81-105 and the activations of
 *                                      scene/gaussian_model.py:35-43,123-157 (f1)
 *   gft_fused_loss                     utils/loss_utils.py:17-33,84-114 + train.py:204-223 (f3)
 *   gft_adam_step                      scene/gaussian_model.py:247-272 (13 param groups through
 *                                      torch.optim.Adam(lr=0, eps=1e-15)) + train.py:467-474 (f4)
 */
#ifndef GFTORF_TRAIN_H_INCLUDED
#define GFTORF_TRAIN_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#include "gftorf.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * f1 — assembly of the rasterizer inputs from the raw (pre-activation) parameters.
 *
 * Reference: seven zero tensors + 14 masked index-copies (gaussian_renderer/__init__.py:81-105)
 * around sigmoid / exp / F.normalize / torch.cat (scene/gaussian_model.py:123-157).  One launch
 * here; every output element is written exactly once.
 *
 *   static  Gaussians (motion mask false), if include_static:
 *       means3D = xyz, opacity = sigmoid(opacity_raw), scales = exp(scaling_raw),
 *       rotations = normalize(rotation_raw), shs = [f_dc | f_rest], shs_p = [[ph_dc, amp_dc] | ...]
 *   dynamic Gaussians (motion mask true), if include_dynamic, with j = dyn_index[i]:
 *       means3D = xyz + d_xyz[j], rotations = normalize(rotation_raw + d_rot[j]),
 *       shs += d_sh[j], shs_p += d_sh_p[j]
 *   Gaussians of an excluded region: every output row is zero (the reference's untouched zeros).
 * ---------------------------------------------------------------------------------------- */
typedef struct GftAssembleArgs {
  int P;
  int M;                 /* SH coefficients per Gaussian: 1 + rest (16 at sh_degree 3) */
  int isotropic;         /* scaling_raw is [P,1] and is repeated over 3 axes (gaussian_model.py:125) */
  int include_static, include_dynamic;   /* render_regions */
  /* raw parameters */
  const float* xyz;            /* [P,3]     */
  const float* opacity_raw;    /* [P,1]     */
  const float* scaling_raw;    /* [P,3] or [P,1] */
  const float* rotation_raw;   /* [P,4]     */
  const float* f_dc_color;     /* [P,1,3]   */
  const float* f_rest_color;   /* [P,M-1,3] */
  const float* f_dc_phase;     /* [P,1,1]   */
  const float* f_rest_phase;   /* [P,M-1,1] */
  const float* f_dc_amp;       /* [P,1,1]   */
  const float* f_rest_amp;     /* [P,M-1,1] */
  /* dynamic region: dyn_index[i] = position of Gaussian i among the masked ones, or -1 (NULL:
   * no dynamic Gaussians).  The d_* arrays are in masked order (the deformation MLP's output,
   * scene/gaussian_model.py:170-174); any of them may be NULL (= zero deformation). */
  const int32_t* dyn_index;    /* [P]        */
  const float* d_xyz;          /* [Nd,3]     */
  const float* d_rot;          /* [Nd,4]     */
  const float* d_sh;           /* [Nd,M,3]   */
  const float* d_sh_p;         /* [Nd,M,2]   */
  /* outputs: the rasterizer's inputs */
  float* means3D;              /* [P,3]   */
  float* opacities;            /* [P,1]   */
  float* scales;               /* [P,3]   */
  float* rotations;            /* [P,4]   */
  float* shs;                  /* [P,M,3] */
  float* shs_p;                /* [P,M,2] */
} GftAssembleArgs;

int gft_assemble_forward(const GftAssembleArgs* args, gft_stream_t stream);

/* Backward of the assembly: gradients w.r.t. the rasterizer inputs -> gradients w.r.t. the raw
 * parameters and the deformation outputs (what autograd derives from the reference's operator
 * chain).  `fwd` is the forward's argument block (inputs and OUTPUTS are read: sigmoid / exp /
 * normalize are differentiated from their saved results).  Every gradient row is written once
 * (zeros for excluded Gaussians); d_* gradient pointers may be NULL. */
typedef struct GftAssembleGrads {
  /* incoming */
  const float* g_means3D; const float* g_opacities; const float* g_scales; const float* g_rotations;
  const float* g_shs; const float* g_shs_p;
  /* outgoing, raw parameters */
  float* g_xyz; float* g_opacity_raw; float* g_scaling_raw; float* g_rotation_raw;
  float* g_f_dc_color; float* g_f_rest_color; float* g_f_dc_phase; float* g_f_rest_phase;
  float* g_f_dc_amp; float* g_f_rest_amp;
  /* outgoing, deformation outputs (masked order) */
  float* g_d_xyz; float* g_d_rot; float* g_d_sh; float* g_d_sh_p;
} GftAssembleGrads;

int gft_assemble_backward(const GftAssembleArgs* fwd, const GftAssembleGrads* grads,
                          gft_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f3 — photometric / ToF loss on a rendered image, value and gradient in one call.
 *
 *   loss = lambda * ( (1 - lambda_dssim) * L(img, gt) + lambda_dssim * (1 - SSIM(img, gt)) )
 *   L: 0 = l1_loss (loss_utils.py:17-18), 1 = l2_loss (:20-21),
 *      2 = weighted_l1_loss over the first `nch` channels, weight = w + sqrt(sum_c img_c^2) detached
 *          (:23-25), 3 = weighted_l1_loss_quad, weight = w + |img| detached (:27-29),
 *      4 = weighted_l2_loss_quad (:31-33)
 *   SSIM: 11x11 Gaussian window, sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2, mean over all
 *   channels and pixels (:84-114).
 * Writes d loss / d img into `grad` ([C,H,W], every element) and adds the loss value to *loss.
 * `scratch` holds 3 floats per image element (gft_fused_loss_scratch_bytes).
 * ---------------------------------------------------------------------------------------- */
typedef struct GftLossArgs {
  int C, H, W;
  int kind;              /* 0..4, see above */
  int nch;               /* kind 2: channels entering the L1 term (<= C) */
  float w;               /* kinds 2..4: weight offset */
  float lambda;          /* overall factor (opt.lambda_color / lambda_tof / lambda_depth) */
  float lambda_dssim;    /* 0 disables the SSIM term (and its two passes) */
  const float* img;      /* [C,H,W] rendered */
  const float* gt;       /* [C,H,W] target   */
  float* grad;           /* [C,H,W] out      */
  float* loss;           /* [1] in/out: loss value is ADDED */
  float* scratch;
} GftLossArgs;

size_t gft_fused_loss_scratch_bytes(int C, int H, int W);
int gft_fused_loss(const GftLossArgs* args, gft_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f4 — Adam over one flat parameter buffer, all parameter groups in one launch.
 *
 * torch.optim.Adam(betas=(0.9,0.999), eps=1e-15, weight_decay=0, amsgrad=False) as the reference
 * configures it; a segment = one param group (its own learning rate; lr = 0 still updates the
 * moments, as torch does).  `step` is the 1-based step count AFTER this update (bias correction).
 * Segment bounds are element offsets, multiples of 4; gaps between segments are left untouched.
 * zero_grad != 0 clears the gradient buffer in the same pass (optimizer.zero_grad, train.py:474).
 * ---------------------------------------------------------------------------------------- */
#define GFT_ADAM_MAX_SEGMENTS 24
typedef struct GftAdamSegment {
  long long begin, end;
  double lr;              /* a Python float in the reference: the step size is formed in double */
} GftAdamSegment;

typedef struct GftAdamArgs {
  float* param;
  float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int n_segments;
  int step;
  float beta1, beta2, eps;
  int zero_grad;
  GftAdamSegment seg[GFT_ADAM_MAX_SEGMENTS];
} GftAdamArgs;

int gft_adam_step(const GftAdamArgs* args, gft_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f2 — adaptive density control: GaussianModel.densify_and_prune (scene/gaussian_model.py:631-646)
 * with densify_and_clone (:607-629), densify_and_split (:568-605, N = 2), prune_points (:493-513)
 * and the optimizer-state surgery (:473-536), as a plan + one gather per parameter group.
 *
 * gft_densify_plan decides, per original Gaussian, what the whole sequence does to it and lays
 * out the new Gaussian set in the reference's final order:
 *     [ kept originals | clones | split children, copy 0 | split children, copy 1 ]
 * counts[0..4] = kept, clones, Gaussians selected for splitting, children kept per copy, new P.
 * plan_* need room for 2*P entries (the new set never exceeds 2*P).  plan_noise[k] is the row of
 * the caller's normal samples a child uses: copy * counts[2] + rank among the selected — the rows
 * of torch.normal(mean = 0, std = scaling[selected].repeat(2, 1)) (:579-581).  split_src
 * (optional, P entries) lists the selected Gaussians so the caller can draw those samples.
 * max_grad must be > 0 (clones carry a zero gradient into the split test, :572-573).
 * size_prune: the world-size criteria of :640-643 (`if max_screen_size`); the screen-size criterion
 * of :639 is always false in the reference because max_radii2D is reset by the appends (:566).
 * ---------------------------------------------------------------------------------------- */
typedef struct GftDensifyPlanArgs {
  int P;
  int isotropic;
  int size_prune;
  float max_grad, min_opacity, extent, percent_dense;
  const float* grad_accum;     /* [P,1] xyz_gradient_accum */
  const float* denom;          /* [P,1] */
  const float* opacity_raw;    /* [P,1] */
  const float* scaling_raw;    /* [P,3] or [P,1] */
  int32_t* plan_src;           /* [2P] out */
  int32_t* plan_kind;          /* [2P] out: 0 kept original, 1 clone, 2 split child */
  int32_t* plan_noise;         /* [2P] out: sample row for children, -1 otherwise */
  int32_t* split_src;          /* [P]  out, optional */
  int32_t* counts;             /* [5]  out */
  char* workspace;             /* gft_densify_workspace_bytes(P) */
} GftDensifyPlanArgs;

size_t gft_densify_workspace_bytes(int P);
int gft_densify_plan(const GftDensifyPlanArgs* args, gft_stream_t stream);

/* Gathers ONE parameter group (row width `width` floats) and its two Adam moments into the new
 * set: kept originals carry their moments, new Gaussians start from zero moments (:527-528).
 * mode 0: rows are copied; mode 1 (xyz, width 3): a child's position is
 * R(rotation_raw / |rotation_raw|) * (noise row) + xyz, noise = the caller's samples (already
 * scaled by the standard deviations); mode 2 (scaling): a child's row is log(exp(s) / (0.8 * 2)). */
typedef struct GftDensifyApplyArgs {
  int P_new;
  int width;
  int mode;
  const int32_t* plan_src; const int32_t* plan_kind; const int32_t* plan_noise;
  const float* param_in; const float* exp_avg_in; const float* exp_avg_sq_in;
  float* param_out; float* exp_avg_out; float* exp_avg_sq_out;   /* moment pointers may be NULL */
  const float* noise;          /* [2*counts[2], 3], mode 1 */
  const float* rotation_raw;   /* [P,4] of the OLD set, mode 1 */
} GftDensifyApplyArgs;

int gft_densify_apply(const GftDensifyApplyArgs* args, gft_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * e — the exchange step of a data-parallel iteration: in-place sum-allreduce of
 * the flat gradient bucket over the GPUs of one NVSwitch domain, through the switch's multicast /
 * in-network reduction (NVLS).  The reference has no multi-GPU code; this stands where an
 * ncclAllReduce(sum, fp32) would.
 *
 * `multicast_ptr` is the MULTICAST address of the bucket: the same symmetric allocation on every
 * rank, bound to an NVLink multicast object (the Python side gets it from
 * torch.distributed._symmetric_memory: rendezvous(t).multicast_ptr + offset).  Rank r reduces and
 * re-broadcasts the r-th 1/world slice.  The caller must run a cross-GPU barrier on the stream
 * before the call (every rank's bucket complete) and after it (results visible everywhere).
 * ---------------------------------------------------------------------------------------- */
int gft_nvls_allreduce_sum(float* multicast_ptr, long long n_floats, int rank, int world,
                           gft_stream_t stream);

/* The same exchange as ONE kernel: a cross-GPU barrier over the symmetric-memory signal pads, the
 * reduce + broadcast of this rank's slice, and a second barrier — no host-launched barrier
 * kernels around it.  `signal_pads_dev`: device array of `world` pointers to the ranks' signal
 * pads (torch: SymmetricMemory.signal_pad_ptrs_dev), each `pad_words` 32-bit words, zero-filled;
 * the library uses their upper half.  `blocks` (0 = automatic) and `unroll` (2, 4 or 8) must be
 * the same on every rank.  Collective: every rank of the group must call it. */
int gft_nvls_allreduce_fused(float* multicast_ptr, long long n_floats, int rank, int world,
                             void* const* signal_pads_dev, int pad_words, int blocks, int unroll,
                             gft_stream_t stream);

/* The same exchange with plain peer-to-peer accesses instead of the multicast object: rank r sums
 * slice r from every rank's symmetric buffer and stores the sum into every buffer (barriers as
 * above).  Moves 2(N-1)/N bucket sizes per GPU and direction against (1 + 1/N) through the switch:
 * the better choice at N = 2.  `buffer_ptrs_host`: HOST array of `world` device pointers to the
 * ranks' symmetric buffers (torch: SymmetricMemory.buffer_ptrs), `byte_offset` the bucket's offset
 * inside them (a multiple of 16).  world <= 8; `unroll` 1, 2 or 4.  Collective. */
int gft_p2p_allreduce_fused(void* const* buffer_ptrs_host, long long byte_offset, long long n_floats, int rank,
                            int world, void* const* signal_pads_dev, int pad_words, int blocks, int unroll,
                            gft_stream_t stream);

/* The peer-to-peer exchange with posted stores only (remote loads are the slow half of the kernel
 * above): every rank pushes the slices it does not own into region [rank] of the owner's scratch,
 * barrier, the owner sums own values + its scratch regions and pushes the sums into every bucket,
 * barrier.  `scratch_ptrs_host`: HOST array of `world` device pointers to a second symmetric
 * buffer of at least world * ceil(n_floats / 4 / world) * 16 bytes per rank, used by nothing else
 * while the call runs.  world <= 8.  Collective. */
int gft_push_allreduce_fused(void* const* buffer_ptrs_host, long long byte_offset, void* const* scratch_ptrs_host,
                             long long n_floats, int rank, int world, void* const* signal_pads_dev, int pad_words,
                             int blocks, gft_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GFTORF_TRAIN_H_INCLUDED */
