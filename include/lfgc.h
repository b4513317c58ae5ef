/*
 * lfgc.h -- C ABI of liblfgc.so, the sm_100a implementation of the fV-SRN
 * latent-feature-grid hot path of Bussler/Latent_Feature_Grid_Compression.
 *
 * The reference has no FFI layer: its operator API for this path is the Python
 * nn.Module surface (model/Feature_Grid_Model.py, model/model_utils.py, ...).
 * The Python modules in latent_feature_grid_compression_b200/ keep that surface
 * and reach the GPU only through the entry points below (ctypes).  Each entry
 * point names the reference code it replaces (file:line relative to the
 * reference root).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions (all functions):
 *   - extern "C", plain pointers and sizes, no torch / C++ types;
 *   - return 0 on success or a negative LFGC_E* code; never throw; the message
 *     of the last failure on the calling thread is lfgc_last_error();
 *   - every pointer is a caller-owned DEVICE pointer (fp32 unless stated) that
 *     must stay valid until the work is complete in stream order;
 *   - the last argument is the cudaStream_t to launch on (as void*); the
 *     functions never synchronise and never allocate device memory, so they are
 *     CUDA-graph capturable; host-side descriptor structs are read at call time;
 *   - re-entrant across host threads and streams.
 */
#ifndef LFGC_H
#define LFGC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LFGC_ABI_VERSION 6 /* 6: tensor-core operand image (lfgc_tc_panel_*); 5: lfgc_peer_announce; 4: + lfgc_train_step_accumulate; 3: lfgc_grid_step / lfgc_train_step_partials replace lfgc_step_glue */
#define LFGC_MAX_LEVELS 12 /* coefficient tensors per model (1 low-pass + up to 11 detail levels) */
#define LFGC_MAX_TAPS 16   /* longest supported 1-D reconstruction filter */
#define LFGC_MAX_LAYERS 8  /* hidden layers of the decoder MLP */
#define LFGC_MAX_PEERS 16  /* ranks whose gradient buffers lfgc_grid_step can sum */

enum {
    LFGC_OK = 0,
    LFGC_E_INVALID = -1,     /* bad argument (null pointer, bad size, unsupported shape) */
    LFGC_E_UNSUPPORTED = -2, /* configuration outside what the kernels are built for */
    LFGC_E_CUDA = -3,        /* a CUDA runtime call failed; see lfgc_last_error() */
    LFGC_E_WORKSPACE = -4    /* caller-provided workspace too small */
};

/* mask-layer kinds (model/Smallify_Dropout.py, model/Variational_Dropout_Layer.py,
 * model/Straight_Through_Dropout.py) */
enum {
    LFGC_MASK_IDENTITY = 0,
    LFGC_MASK_DIRECT = 1,      /* m = p0                (Smallify betas, Smallify_Dropout.py:57; baked d_mask, :59) */
    LFGC_MASK_VARIATIONAL = 2, /* m = exp(p0) + exp(p1/2) * noise  (Variational_Dropout_Layer.py:104-110) */
    LFGC_MASK_STE_SIGMOID = 3, /* m = [sigmoid(p0) >= thr], aux = sigmoid(p0)  (Straight_Through_Dropout.py:55-58) */
    LFGC_MASK_BERNOULLI = 4    /* m = [noise < p0]      (Straight_Through_Dropout.py:28-29) */
};

/* flags for lfgc_forward / lfgc_reconstruct / lfgc_train_step */
enum {
    LFGC_F_CLAMP = 1         /* clamp output to [-1, 1] (eval mode, Feature_Grid_Model.py:78) */
};

/* Static shape of one model instance: what model_utils.setup_model fixes (model/model_utils.py:23-59). */
typedef struct lfgc_model_desc {
    int32_t C;      /* grid_features */
    int32_t Cp;     /* channel stride of the channels-last grid, multiple of 4, >= C (pad channels are zero) */
    int32_t G[3];   /* decoded grid extents (D, H, W) = (z, y, x); the reference uses G,G,G */
    int32_t H;      /* n_hidden_size */
    int32_t L;      /* n_layers (hidden layers, each followed by SnakeAlt) */
    int32_t F;      /* n_embedding_freq; input width = 3 + 6F + C */
} lfgc_model_desc;

/* Wavelet synthesis chain: Feature_Grid_Model.decode_volume (model/Feature_Grid_Model.py:102-108) over
 * _WaveletFilterNd.decode (wavelet_transform/Torch_Wavelet_Transform.py:91-104). */
typedef struct lfgc_wavelet_desc {
    int32_t n_coeff;                    /* number of coefficient tensors: 1 + number of synthesis levels */
    int32_t C;                          /* channels */
    int32_t n_taps;                     /* reconstruction filter length (even) */
    float rec_lo[LFGC_MAX_TAPS];        /* pywt rec_lo, rounded to fp32 as the reference stores it (:41) */
    float rec_hi[LFGC_MAX_TAPS];
    int32_t dims[LFGC_MAX_LEVELS][3];   /* spatial extent of coefficient tensor l: level 0 is (C,d,d,d), l>=1 (C,7,d,d,d) */
    int32_t target[LFGC_MAX_LEVELS][3]; /* output extent of synthesis level l (l>=1), = shape_array[l-1] */
} lfgc_wavelet_desc;

int lfgc_abi_version(void);
/* sizeof(lfgc_wavelet_desc), sizeof(lfgc_model_desc), sizeof(lfgc_peer_announce), sizeof(lfgc_grid_step_args) into out[0..n);
 * returns how many there are.  Bindings compare their own struct layouts with these when they load the library. */
int lfgc_struct_sizes(size_t* out, int n);
const char* lfgc_last_error(void);
/* number of SMs of the current device (grid sizing); negative on error */
int lfgc_sm_count(void);
/* kernels launched by this library so far in this process (host-side counter; a CUDA-graph replay re-issues the
 * launches recorded during capture without passing through here) */
long long lfgc_launch_count(void);

/* ---- mask layers ------------------------------------------------------------------------------------------- */

/* Mask multiplier of one mask tensor (n elements).  Replaces the forward of SmallifyDropout /
 * VariationalDropout / MaskedWavelet_Straight_Through_Dropout / Straight_Through_Dropout
 * (Smallify_Dropout.py:54-61, Variational_Dropout_Layer.py:101-113, Straight_Through_Dropout.py:26-30,53-61).
 * mult_out[n] is the value multiplier; aux_out[n] (nullable) is the gradient multiplier (sigmoid for the STE,
 * otherwise equal to mult).  noise is the caller-drawn xi ~ N(0,1) or u ~ U[0,1) (same torch generator order
 * as the reference). */
int lfgc_mask_multiplier(int mode, int64_t n, const float* p0, const float* p1, const float* noise,
                         float threshold, float* mult_out, float* aux_out, void* stream);

/* d mult -> d mask parameters (autograd of the above).  g0/g1 receive (accumulate != 0: +=) the gradients of
 * p0/p1.  For LFGC_MASK_BERNOULLI and LFGC_MASK_IDENTITY nothing is written. */
int lfgc_mask_param_grad(int mode, int64_t n, const float* p0, const float* p1, const float* noise,
                         const float* gmult, float* g0, float* g1, int accumulate, void* stream);

/* Smallify sign-variance tracker, device side: replaces SmallifySignVarianceTracker.sign_variance_pruning_onlyVar
 * (Smallify_Dropout.py:103-112), which round-trips through the host every step. */
int lfgc_smallify_ema(const float* betas, float* ema, float* emavar, int64_t n, float momentum, void* stream);

/* ---- wavelet synthesis ------------------------------------------------------------------------------------- */

/* One analysis level, used at model construction only: _WaveletFilterNd.encode
 * (wavelet_transform/Torch_Wavelet_Transform.py:75-89) as called by Feature_Grid_Model.encode_volume
 * (model/Feature_Grid_Model.py:83-99).  x (C,d0,d1,d2) -> out (C,8,e0,e1,e2); dec_lo/dec_hi are HOST arrays of the
 * pywt decomposition taps (the flip of :56 is applied inside).  e_out receives the output extents; with x or out
 * NULL only e_out is computed. */
int lfgc_dwt_level(const float* x, int C, const int32_t d[3], int n_taps, const float* dec_lo, const float* dec_hi,
                   float* out, int32_t e_out[3], void* stream);

/* bytes of scratch lfgc_decode_fwd / lfgc_decode_bwd need for this chain */
size_t lfgc_decode_scratch_bytes(const lfgc_wavelet_desc* w);

/* grid_cl[z][y][x][Cp] = synthesis of (coeff[l] * mult[l]); mult[l] may be NULL (identity).
 * coeff / mult are HOST arrays of n_coeff DEVICE pointers.  Pad channels [C, Cp) are written as zero.
 * With n_coeff == 1 this is a masked NCDHW -> channels-last transpose.
 * also_zero (nullable): a buffer shaped like grid_cl that is cleared in the same pass -- the gradient accumulator
 * the coming lfgc_backward / lfgc_train_step scatters into (saves a separate memset launch per step). */
int lfgc_decode_fwd(const lfgc_wavelet_desc* w, const float* const* coeff, const float* const* mult,
                    float* scratch, float* grid_cl, int Cp, float* also_zero, void* stream);

/* Adjoint of lfgc_decode_fwd (autograd of conv_transpose3d + mul, training/training.py:137).
 * grad_coeff[l] (+)= d(coeff*mult) * gmul[l]  where gmul[l] is the gradient multiplier (aux of the mask; NULL = 1)
 * grad_mult[l]  (+)= sum_c coeff[l][c] * d(coeff*mult)[c]      (NULL entries are skipped)
 * accumulate: 0 overwrite (grad_mult is cleared first, one memset per level), 1 add to both, 2 overwrite grad_coeff and
 * add into grad_mult, which the caller guarantees to be zero (lfgc_adam_reg clears it): no memsets. */
int lfgc_decode_bwd(const lfgc_wavelet_desc* w, const float* grad_grid_cl, int Cp, const float* const* coeff,
                    const float* const* gmul, float* scratch, float* const* grad_coeff, float* const* grad_mult,
                    int accumulate, void* stream);

/* ---- per-sample path ----------------------------------------------------------------------------------------- */

/* number of fp32 values in the packed MLP parameter block:
 * [W0 (H x in) | b0 (H) | W1 (H x H) | b1 | ... | Wf (1 x H) | bf (1)], nn.Linear (out,in) row-major,
 * i.e. net_layers.i.weight/.bias then final_layer.weight/.bias in state_dict order. */
int64_t lfgc_mlp_param_count(const lfgc_model_desc* m);

/* out[n] = MLP([xyz | Fourier(xyz) | trilinear(grid_cl, xyz)]).  Replaces Feature_Grid_Model.forward after
 * decode_volume (model/Feature_Grid_Model.py:62-78): F.grid_sample(bilinear, zeros, align_corners=False),
 * FourierEmbedding.embed (model/Feature_Embedding.py:14-16), 4 x SnakeAlt(Linear), final Linear, optional clamp. */
int lfgc_forward(const lfgc_model_desc* m, const float* coords, int64_t n, const float* grid_cl,
                 const float* mlp, float* out, int flags, void* stream);

/* workspace (bytes) for lfgc_backward / lfgc_train_step: per-CTA partial sums of the MLP gradient */
size_t lfgc_backward_workspace_bytes(const lfgc_model_desc* m);

/* Backward of lfgc_forward for d(loss)/d(out) = grad_out[n] (recomputes the forward; nothing is saved):
 *   grad_grid_cl[z][y][x][Cp] += trilinear scatter of d(features)       (L2 vector atomics)
 *   grad_mlp[P]               (+)= MLP weight/bias gradients             (deterministic two-stage reduction)
 *   grad_coords                 must be NULL: the coordinate gradient is NOT produced (the reference's autograd computes
 *                               one because training.py:99 sets requires_grad, and never reads it); a non-NULL
 *                               pointer is refused with LFGC_E_UNSUPPORTED rather than silently left unwritten
 * Replaces the autograd backward of grid_sampler_3d, cat, addmm, sin/pow (training/training.py:137). */
int lfgc_backward(const lfgc_model_desc* m, const float* coords, int64_t n, const float* grad_out,
                  const float* grid_cl, const float* mlp, float* grad_grid_cl, float* grad_mlp,
                  float* grad_coords, int accumulate_mlp, void* workspace, size_t workspace_bytes, void* stream);

/* Fused training step core: in-kernel Philox voxel sampler (IndexDataset.__getitem__, data/IndexDataset.py:90-96),
 * exact ground-truth lookup (data/Interpolation.py:8-44 at integer positions), forward, MSE loss
 * (training/training.py:130,201) and backward, in one launch.
 *   volume[R0][R1][R2] normalised volume; sample i of the step uses Philox(seed, counter = sample_offset + i
 *   + *step_dev * step_stride); step_dev (nullable) is a DEVICE int32 step counter (the one lfgc_adam increments),
 *   which lets a captured CUDA graph draw fresh samples on every replay
 *   loss_scale: d(loss)/d(pred) = loss_scale * 2 * (pred - gt); pass 1/N_global for the mean over the global batch
 *   loss_sum[0] = sum (pred-gt)^2 over this call's samples (overwritten; summed with the MLP gradient partials)
 *   explicit_idx (nullable, int64[n]): use these flat voxel indices instead of Philox (parity tests)
 *   explicit_coords / explicit_gt (nullable, both or neither; fp32 [n][3] normalised positions and [n] target
 *   values): caller-supplied samples instead of the in-kernel sampler -- the host-fed training step (the
 *   reference's DataLoader contract, training/training.py:89-109); volume may then be NULL */
int lfgc_train_step(const lfgc_model_desc* m, const float* volume, const int32_t R[3], int64_t n,
                    uint64_t seed, uint64_t sample_offset, const int32_t* step_dev, uint64_t step_stride,
                    const int64_t* explicit_idx, const float* explicit_coords, const float* explicit_gt,
                    float loss_scale, const float* grid_cl, const float* mlp, float* grad_grid_cl, float* grad_mlp,
                    float* loss_sum, int accumulate_mlp, void* workspace, size_t workspace_bytes, void* stream);

/* lfgc_train_step with the Gaussian log-likelihood of VariationalDropoutLoss (model/Variational_Dropout_Layer.py:24-30,
 * 54-69) instead of the plain MSE: log_sigma[n] (nullable) is the per-sample log sigma v_i (the Variance_Model output,
 * training/training.py:119-121);  d(loss)/d(pred_i) = 2 loss_scale (pred_i - gt_i) exp(-2 v_i)  and
 * dlog_sigma[n] (nullable) receives  d(loss)/d(v_i) = 2 loss_scale (1 - (pred_i - gt_i)^2 exp(-2 v_i)).
 * Pass loss_scale = n_voxels / (2 N_global) for the reference's batch_scale.  loss_sum stays sum (pred-gt)^2. */
int lfgc_train_step_weighted(const lfgc_model_desc* m, const float* volume, const int32_t R[3], int64_t n,
                             uint64_t seed, uint64_t sample_offset, const int32_t* step_dev, uint64_t step_stride,
                             const int64_t* explicit_idx, const float* explicit_coords, const float* explicit_gt,
                             float loss_scale, const float* log_sigma, float* dlog_sigma, const float* grid_cl,
                             const float* mlp, float* grad_grid_cl, float* grad_mlp, float* loss_sum, int accumulate_mlp,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ---- Variance_Model (model/Variational_Dropout_Layer.py:159-175) -------------------------------------------------- */

/* out[n] = final(ReLU(net_{L-1}(... ReLU(net_0(x))))) for x[n][3]: the plain coordinate MLP 3 -> H x L -> 1 that
 * predicts log sigma per sample.  mlp is packed like lfgc_mlp_param_count with an input width of 3
 * (net_layers.i.weight/.bias, final_layer.weight/.bias).  H <= 32, L <= 4 (the reference builds 32 x 4). */
int lfgc_plain_mlp_forward(int H, int L, const float* x, int64_t n, const float* mlp, float* out, void* stream);
size_t lfgc_plain_mlp_workspace_bytes(int H, int L);
/* grad_mlp (+)= parameter gradients for d(loss)/d(out) = grad_out[n] (forward recomputed, deterministic reduction) */
int lfgc_plain_mlp_backward(int H, int L, const float* x, int64_t n, const float* grad_out, const float* mlp,
                            float* grad_mlp, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ---- sampler / ground truth ---------------------------------------------------------------------------------- */

/* Random voxel sampler: replaces IndexDataset.__getitem__ + DataLoader collation (data/IndexDataset.py:90-96) and
 * the ground-truth lookup trilinear_f_interpolation at integer positions (training/training.py:107-109).
 * raw_out[n][3] = (i,j,k) as fp32, norm_out[n][3] = scales * (2*raw/max_idx - 1) with the reference's fp32
 * operation order, gt_out[n] = volume[i][j][k].  Any output may be NULL.  explicit_idx as above. */
int lfgc_sample(const float* volume, const int32_t R[3], int64_t n, uint64_t seed, uint64_t sample_offset,
                const int64_t* explicit_idx, float* raw_out, float* norm_out, float* gt_out, void* stream);

/* lfgc_sample whose Philox counter also advances with a DEVICE step counter (counter = sample_offset + i +
 * *step_dev * step_stride, as in lfgc_train_step), so a captured CUDA graph draws fresh samples on every replay. */
int lfgc_sample_stream(const float* volume, const int32_t R[3], int64_t n, uint64_t seed, uint64_t sample_offset,
                       const int32_t* step_dev, uint64_t step_stride, const int64_t* explicit_idx, float* raw_out,
                       float* norm_out, float* gt_out, void* stream);

/* General trilinear_f_interpolation(p, f, min_bb, max_bb, res) (data/Interpolation.py:8-44): fp32 lattice
 * coordinates, fp64 alphas, f[x][y][z] indexing, lerp order x -> y -> z. min_bb/max_bb are HOST float[3]. */
int lfgc_trilinear(const float* p, int64_t n, const float* volume, const int32_t R[3], const float min_bb[3],
                   const float max_bb[3], float* out, void* stream);

/* ---- reconstruction ------------------------------------------------------------------------------------------ */

/* Full-volume reconstruction of the slab [slab_begin, slab_end) along volume dim 0: replaces field_from_net
 * (visualization/OutputToVTK.py:7-47): eval forward with clamp over every voxel of the slab.
 * axis0[R0], axis1[R1], axis2[R2] are the normalised coordinates of each voxel index along the three volume
 * dims (the host builds them with the reference's per-tile fp32 linspace * 2 - 1, times scales, so the
 * coordinates are bit-identical to the reference's); sample (i,j,k) is evaluated at (axis0[i], axis1[j], axis2[k]).
 * out_slab[(slab_end-slab_begin)][R1][R2]. */
int lfgc_reconstruct(const lfgc_model_desc* m, const float* grid_cl, const float* mlp, const int32_t R[3],
                     const float* axis0, const float* axis1, const float* axis2, int32_t slab_begin,
                     int32_t slab_end, float* out_slab, int flags, void* stream);

/* Deviation statistics (visualization/OutputToVTK.py:53-60), accumulated in fp64 on the device:
 * acc[0] += sum (gt-pred)^2, acc[1] += sum |gt-pred|, acc[2] = max(acc[2], max gt), acc[3] = min(acc[3], min gt).
 * The caller initialises acc = {0, 0, -inf, +inf} and finishes PSNR = 10 log10((max-min)^2 / (acc[0]/n)). */
int lfgc_deviation_stats(const float* pred, const float* gt, int64_t n, double* acc, void* stream);

/* ---- optimiser ----------------------------------------------------------------------------------------------- */

/* torch.optim.Adam (training/training.py:199,232) on flat buffers, one launch.  step_count is a DEVICE int32[2]:
 * [0] = number of steps taken so far (incremented by the kernel, so the call is graph-replayable; this is the
 * counter lfgc_train_step reads), [1] = scratch ticket counter that must be zero-initialised;
 * lr is a DEVICE float (the host decay strategies write it).  grad_scale multiplies the gradient first
 * (1/world for data parallel means).  The hyper-parameters are doubles, as in torch: 1 - beta and log(beta) are formed
 * in fp64 on the host and then rounded (1 - fl32(0.999) would be 4.7e-5 off); the bias corrections 1 - beta^step are
 * evaluated on the device in fp32 as -expm1(step * fl32(log(beta))).  The sample-independent regulariser gradients are separate calls (lfgc_add_l1_grad,
 * lfgc_add_l2_grad, lfgc_variational_dkl_grad) issued before this one. */
int lfgc_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* lr, int32_t* step_count,
              double beta1, double beta2, double eps, double grad_scale, void* stream);

/* p-gradient of the sample-independent regularisers added in place: g += w_l2 * 2 * p (n_l2 leading elements) */
/* lfgc_adam with the SmallifyLoss terms (model/Smallify_Dropout.py:22-40) folded into the gradient: + 2 weight_l2 p on
 * elements [l2_begin, l2_end) (the wavelet coefficients) and + weight_l1 sign(p) on [l1_begin, l1_end) (the mask
 * parameters); one launch instead of lfgc_add_l2_grad + lfgc_add_l1_grad + lfgc_adam.  Optional bookkeeping of the mask
 * range in the same pass: ema / emavar (nullable pair, l1_end - l1_begin floats) receive the sign-variance tracker update
 * of lfgc_smallify_ema with the PRE-update parameter (the value this step's forward used); zero_l1_grad != 0 clears the
 * gradient elements of the range after they were consumed (the adjoint accumulates into them with atomics). */
int lfgc_adam_reg(float* p, float* g, float* m, float* v, int64_t n, const float* lr, int32_t* step_count,
                  double beta1, double beta2, double eps, double grad_scale, int64_t l2_begin, int64_t l2_end,
                  double weight_l2, int64_t l1_begin, int64_t l1_end, double weight_l1, float* ema, float* emavar,
                  float momentum, int zero_l1_grad, void* stream);
int lfgc_add_l2_grad(float* g, const float* p, int64_t n, float weight, void* stream);
int lfgc_add_l1_grad(float* g, const float* p, int64_t n, float weight, void* stream);

/* lfgc_train_step that ADDS its MLP-gradient sums and its loss sum to grad_mlp_loss: n_slices rows of
 * lfgc_mlp_param_count + 1 floats (the last float of a row is the loss); the buffer is a running sum the caller cleared (or
 * carries over) and the result is the SUM OF THE ROWS (lfgc_grid_step's nslices / pstride take exactly this).  The CTAs are
 * spread over the rows because same-address reductions of 148 CTAs arriving together serialise in L2 (one row: +3.5 us);
 * kernels other than the tensor-core one add everything to row 0.  The tensor-core kernel adds with
 * atomics from its own epilogue, so there is no reduction launch and no use of the workspace slices; the summation order is
 * then not fixed (like the grid gradient's, which always accumulates with atomics).  This is the per-sample launch of the
 * data-parallel step: the buffer is the MLP section of the [grid gradient | MLP gradient | loss] message lfgc_peer_sum reads. */
/* Optional early announcement for lfgc_peer_sum: once the LAST CTA of the per-sample kernel has added its sums (all of this
 * rank's contribution to the step is then complete), it stores epoch[0] + 1 into slot [rank] of every rank's flag array --
 * what lfgc_peer_sum would otherwise do at its start -- so that the flags travel over NVLink while the next launch is still
 * being set up.  Pass the same flags / rank / epoch as to lfgc_peer_sum (and announced = 1 there); ticket: device int32,
 * 0 at the start, reset by the kernel. */
typedef struct lfgc_peer_announce {
    int32_t n_peers, rank;
    int32_t* flags[LFGC_MAX_PEERS];
    const int32_t* epoch;
    int32_t* ticket;
} lfgc_peer_announce;

int lfgc_train_step_accumulate(const lfgc_model_desc* m, const float* volume, const int32_t R[3], int64_t n, uint64_t seed,
                               uint64_t sample_offset, const int32_t* step_dev, uint64_t step_stride,
                               const int64_t* explicit_idx, const float* explicit_coords, const float* explicit_gt,
                               float loss_scale, const float* grid_cl, const float* mlp, float* grad_grid_cl,
                               float* grad_mlp_loss, int n_slices, const lfgc_peer_announce* announce /* nullable */,
                               const float* tc_panels /* nullable: lfgc_tc_panel_build */, void* workspace,
                               size_t workspace_bytes, void* stream);

/* Weight operands of the tensor-core training kernel as a ready-made image in device memory (csrc/tc_panels.cuh): the
 * tf32 hi / lo split of every MLP weight in the kernel's two shared-memory operand layouts plus the biases, so that the
 * kernel's per-launch setup is a straight copy instead of a rebuild (12 k of 64 k cycles per launch at 32768 samples).
 * lfgc_tc_panel_bytes: size of the image, 0 when the tensor-core kernel does not cover the model (nothing to pass then).
 * lfgc_tc_panel_build: (re)builds the image from the packed parameter block `mlp` -- call it after the parameters changed
 * outside lfgc_grid_step, which keeps the image current itself (lfgc_grid_step_args.panel_image).  Passing a stale image to
 * lfgc_train_step_partials / _accumulate trains on stale weights: the owner of the parameters owns the image. */
size_t lfgc_tc_panel_bytes(const lfgc_model_desc* m);
int lfgc_tc_panel_build(const lfgc_model_desc* m, const float* mlp, float* image, void* stream);

/* lfgc_train_step that LEAVES the MLP-gradient partial sums in the workspace instead of reducing them: *nslices_out (host
 * int, written at call time) rows of (lfgc_mlp_param_count + 1) floats, the last float of a row being that slice's
 * share of the loss sum.  lfgc_grid_step reduces them (same fixed order -> deterministic) together with the rest of the
 * optimiser step, which saves a dependent launch.  n must be positive. */
int lfgc_train_step_partials(const lfgc_model_desc* m, const float* volume, const int32_t R[3], int64_t n, uint64_t seed,
                             uint64_t sample_offset, const int32_t* step_dev, uint64_t step_stride,
                             const int64_t* explicit_idx, const float* explicit_coords, const float* explicit_gt,
                             float loss_scale, const float* grid_cl, const float* mlp, float* grad_grid_cl,
                             const float* tc_panels /* nullable: lfgc_tc_panel_build */, void* workspace,
                             size_t workspace_bytes, int32_t* nslices_out, void* stream);

/* Everything of one optimiser step that is not per-sample, for mask-free models, in ONE launch: [sum over the
 * data-parallel ranks' gradient buffers, read from peer memory] -> reduction of the MLP-gradient partial sums (+ loss) ->
 * synthesis adjoint (autograd of Feature_Grid_Model.decode_volume, model/Feature_Grid_Model.py:102-108) -> [+ 2 weight_l2
 * coeff, SmallifyLoss' weight term, model/Smallify_Dropout.py:29,37] -> Adam over coefficients and MLP (torch.optim.Adam,
 * training/training.py:199,232) -> synthesis of the UPDATED coefficients into the channels-last grid the next step gathers
 * from, gradient accumulator cleared.  Replaces lfgc_decode_bwd + lfgc_adam + the next step's lfgc_decode_fwd (+ the
 * partial reduction of lfgc_train_step).  One CTA per channel, the channel's wavelet pyramid in shared memory, levels
 * evaluated separably; no grid-wide barrier.  step_count as for lfgc_adam ([0] steps taken, [1] ticket scratch = 0).
 * All pointers inside the struct are DEVICE pointers; the struct itself is read on the host at call time. */
typedef struct lfgc_grid_step_args {
    int32_t n_srcs;                               /* gradient sources summed in index order (normally 1) */
    const float* grad_grid[LFGC_MAX_PEERS];       /* channels-last grid gradients (G0,G1,G2,Cp), one per source */
    const float* mlp_partials[LFGC_MAX_PEERS];    /* per source: nslices rows of pstride floats (MLP gradient, loss at [pcount]) */
    int32_t nslices;                              /* rows per source (lfgc_train_step_partials' nslices_out; 1 after a reduction) */
    int32_t pstride, pcount;                      /* row stride (>= pcount + 1) and MLP parameter count (0: no MLP block) */
    float* zero_grid;                             /* gradient accumulator to clear (normally grad_grid[0]); NULL: none */
    float* grid_cl;                               /* out: decoded grid of the updated coefficients, channels-last */
    float* p;                                     /* flat parameter buffer ... */
    float* g;                                     /* ... gradients (written: coefficient and MLP sections) ... */
    float* m;                                     /* ... Adam first ... */
    float* v;                                     /* ... and second moments */
    int64_t coeff_off[LFGC_MAX_LEVELS];           /* offset (floats) of coefficient tensor l inside the flat buffers */
    int64_t mlp_off;                              /* offset of the packed MLP block */
    float* loss_out;                              /* sum of the loss slots of all sources / slices; NULL: none */
    const float* lr;                              /* device scalar */
    int32_t* step_count;
    double beta1, beta2, eps, grad_scale, weight_l2;
    /* Optional scratch of lfgc_grid_step_scratch_bytes(): with it (and n_srcs == 1, >= 1 wavelet level) the FINEST level
     * runs on the whole GPU -- its adjoint with the Adam update of its detail bands fused in, and its synthesis -- and only
     * the coarser levels go through the per-channel shared-memory kernel: three dependent launches. */
    float* scratch;
    size_t scratch_bytes;
    /* Optional: the tensor-core operand image of the model whose MLP block this step updates (lfgc_tc_panel_bytes > 0);
     * the thread that applies Adam to a parameter also stores its hi / lo parts at their operand positions. */
    const lfgc_model_desc* panel_model;
    float* panel_image;
} lfgc_grid_step_args;
int lfgc_grid_step(const lfgc_wavelet_desc* w, int Cp, const lfgc_grid_step_args* a, void* stream);
/* dynamic shared memory lfgc_grid_step needs for this pyramid; 0 = not supported (does not fit: use the separate kernels) */
size_t lfgc_grid_step_smem_bytes(const lfgc_wavelet_desc* w);
size_t lfgc_grid_step_scratch_bytes(const lfgc_wavelet_desc* w);
/* 1 when lfgc_grid_step covers this pyramid (whole in shared memory, or split with the scratch), else 0 */
int lfgc_grid_step_supported(const lfgc_wavelet_desc* w);

/* Data-parallel gradient sum without a collective: out[i] = sum_r srcs[r][i] (rank order: all ranks get bit-identical sums),
 * the sources being every rank's buffer as mapped into THIS process (peer memory, e.g. torch symmetric memory), behind a
 * barrier inside the kernel: this rank stores its epoch (epoch[0] + 1) into slot [rank] of every rank's flag array
 * (flags[r]: int32[n_srcs]) and waits until all n_srcs slots of its own array carry it; the kernel then publishes the new
 * epoch (epoch: device int32[2] = {epochs so far, ticket scratch}, both 0 at the start).  Every rank must issue the same
 * sequence of launches.  The caller double-buffers the sources by step parity, so one barrier per step suffices; `zero`
 * (nullable, n floats: the other parity's local buffer) is cleared in the same pass.  n must be a multiple of 4 and the
 * buffers 16-byte aligned.  Replaces the NCCL all-reduce of the path when the ranks share a node. */
int lfgc_peer_sum(const float* const* srcs, int32_t* const* flags, int n_srcs, int rank, int32_t* epoch, float* out,
                  float* zero, int64_t n, int announced /* 1: lfgc_train_step_accumulate already stored this epoch's flags */,
                  void* stream);

/* Gradient of the KL regulariser of VariationalDropoutLoss (model/Variational_Dropout_Layer.py:54-69,115-122) added
 * in place to the mask-parameter gradients, and the per-step ramp of its weight (:57-58).  mask_params / mask_grads
 * point at a buffer holding, per mask layer i, [log_thetas_i (n_i) | log_var_i (n_i)]; layer_sizes is a HOST array
 * of the n_i.  w_dkl is a DEVICE double[2] (both entries initialised to the starting weight), indexed by the parity
 * of *step_count (the lfgc_adam counter): w <- w * ramp while w < w_max, then
 * grad += w * scale * d DKL / d(param), with scale = batch_scale = n_voxels / N. */
int lfgc_variational_dkl_grad(const float* mask_params, float* mask_grads, int n_layers, const int64_t* layer_sizes,
                              double* w_dkl, const int32_t* step_count, double ramp, double w_max, float scale,
                              void* stream);

/* All live variational mask layers at once (instead of one lfgc_mask_multiplier / lfgc_mask_param_grad per layer).  The
 * flat mask section holds, per layer i, [log_thetas_i (n_i) | log_var_i (n_i)] (as for lfgc_variational_dkl_grad);
 * noise / mult_out / gmult are the layers' masks concatenated (sum n_i floats).
 *   lfgc_variational_multiplier: mult = exp(log_thetas) + exp(log_var / 2) * noise (Variational_Dropout_Layer.py:104-108);
 *                                zero_out (nullable, sum n_i floats) is cleared in the same pass
 *   lfgc_variational_param_grad: mask_grads (flat mask section, overwritten) = autograd of the above for d loss/d mult = gmult */
int lfgc_variational_multiplier(const float* mask_params, const float* noise, int n_layers, const int64_t* layer_sizes,
                                float* mult_out, float* zero_out, void* stream);
int lfgc_variational_param_grad(const float* mask_params, const float* noise, const float* gmult, int n_layers,
                                const int64_t* layer_sizes, float* mask_grads, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LFGC_H */
