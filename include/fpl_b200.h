/* fpl_b200.h -- C ABI of the B200-native T-bar detection hot path (libfplb200.so).
 *
 * This is the drop-in boundary for ONE path of janelia-flyem/flypylib:
 *
 *     EM volume --[3-D CNN forward, tiled]--> probability map --[voxel2obj]--> point detections
 *
 * Every entry point states the reference interface it replaces (file:line relative to the
 * reference tree).  The reference is pure Python; its "FFI" for this path is the Keras /
 * SciPy / NumPy call sites, so the binding a maintainer adds is a ctypes stub.
 *
 * Conventions
 *   - plain C: pointers, sizes, scalars.  No torch / C++ types.
 *   - every function returns 0 on success, a negative FPL_E* code on failure;
 *     fpl_last_error() returns a thread-local human-readable message.
 *   - pointers named d_* are DEVICE pointers (sm_100a B200 global memory) owned by the caller;
 *     h_* are host pointers.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - the library allocates only inside an explicit fpl_ctx (grow-only workspace arena,
 *     activation buffer pool, packed weights); fpl_ctx_destroy frees everything.
 *   - one fpl_ctx per (process, device).  Calls on one context are stream-ordered: its workspace and
 *     activation buffers are recycled in launch order, so use ONE stream at a time per context
 *     (synchronise before switching streams).
 *   - volumes are C-order (Z,Y,X); detections are rows (x, y, z, conf) of float64, exactly the
 *     columns of the reference's obj_pred array (flypylib/fplobjdetect.py:211,233-257).
 *   - there is NO CPU fallback: every compute entry point fails with FPL_ENODEV when no sm_100
 *     device is present.
 */
#ifndef FPL_B200_H
#define FPL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPL_OK        0
#define FPL_EINVAL   -1   /* bad argument */
#define FPL_ECUDA    -2   /* CUDA runtime / driver error (see fpl_last_error) */
#define FPL_ENOMEM   -3   /* device allocation failed */
#define FPL_ENODEV   -4   /* no sm_100 device */
#define FPL_EOVERFLOW -5  /* caller-provided output capacity too small */
#define FPL_ESTATE   -6   /* object used in the wrong state (e.g. network without weights) */

typedef struct fpl_ctx fpl_ctx;
typedef struct fpl_net fpl_net;

/* ---------------------------------------------------------------------------------------------
 * library / context
 * ------------------------------------------------------------------------------------------- */
int         fpl_version(void);                 /* ABI version, currently 1 */
const char *fpl_last_error(void);
int         fpl_device_count(int *count);      /* number of visible sm_100 devices */

/* One context per (process, GPU).  Replaces nothing in the reference (Keras/TF session state). */
int fpl_ctx_create(int device, fpl_ctx **out);
int fpl_ctx_destroy(fpl_ctx *ctx);
/* bytes currently held by the context's workspace arena */
int fpl_ctx_workspace_bytes(fpl_ctx *ctx, int64_t *bytes);
/* give the grow-only workspace arena and the activation buffer pool back to the driver (they regrow on demand);
 * synchronises the device.  For callers that alternate between very different problem sizes. */
int fpl_ctx_release_workspace(fpl_ctx *ctx);
/* number of kernel launches issued by this context since creation (bench.py "gpu_launches") */
int fpl_ctx_launch_count(fpl_ctx *ctx, int64_t *launches);

/* Device-side timing of kernel families with CUDA events recorded on the launching stream
 * (bench.py roofline).  profile_end synchronises the device and returns, per family
 * (0 conv3x3x3 tcgen05, 1 conv1x1x1 tcgen05, 2 first-layer conv, 3 pool/upsample/concat/final,
 * 4 Gaussian passes, 5 radix select, 6 NMS (compaction, rounds, sort), 7 tile gather/scatter):
 * summed milliseconds, summed algorithmic work (FLOP for 0-2, bytes otherwise) and launch-group
 * count.  Arrays have 8 entries. */
int fpl_ctx_profile_begin(fpl_ctx *ctx);
int fpl_ctx_profile_end(fpl_ctx *ctx, double *ms_by_tag, double *work_by_tag, int64_t *count_by_tag);

/* ---------------------------------------------------------------------------------------------
 * voxel2obj: smoothing + percentile threshold + greedy NMS       (flypylib/fplobjdetect.py:132-257)
 * ------------------------------------------------------------------------------------------- */

/* Parameters of one voxel2obj call.  Mirrors the reference signature
 *   voxel2obj(pred, obj_min_dist, smoothing_sigma, volume_offset, buffer_sz, thd)
 * (fplobjdetect.py:132-135), seg=None branch.  The Gaussian taps are supplied by the host shim
 * exactly as SciPy builds them (fplobjdetect.py:167-168 -> scipy _gaussian_kernel1d). */
typedef struct fpl_v2o_params {
    int32_t  obj_min_dist;      /* r: pad width, border width, suppression-ball radius            */
    int32_t  lw;                /* Gaussian half width = int(2*sigma+0.5); -1: sigma==0, no filter */
    const double *h_weights;    /* 2*lw+1 taps (host), correlate order                            */
    double   thd;               /* thd as numpy would promote it against a float32 (see shim)     */
    int64_t  rank_lo, rank_hi;  /* 0-based order statistics of the PADDED volume used by          */
    float    gamma;             /*   np.percentile(.,97) and its float32 lerp weight (:183)       */
    int32_t  buffer_xyz[3];     /* buffer_sz in (x,y,z) order as the reference applies it (:239-250) */
    double   offset_xyz[3];     /* volume_offset (x,y,z) (:252-253)                               */
} fpl_v2o_params;

/* Stage A (fplobjdetect.py:158-175): interior (Z,Y,X) of  zero_border(gaussian_filter(pad(pred, r))).
 * The r-wide border of the padded map is identically zero and is never materialised.
 * d_smooth may not alias d_pred.  Uses ctx workspace: 1 x Z*Y*X floats. */
int fpl_v2o_smooth(fpl_ctx *ctx, const float *d_pred, int64_t Z, int64_t Y, int64_t X,
                   const fpl_v2o_params *p, float *d_smooth, void *stream);

/* Stage B (fplobjdetect.py:183): threshold = max(percentile_97(padded map), thd).
 * h_out[0] = threshold (as double), h_out[1] = value at rank_lo, h_out[2] = value at rank_hi,
 * h_out[3] = number of NaNs seen.  Synchronises `stream`. */
int fpl_v2o_threshold(fpl_ctx *ctx, const float *d_smooth, int64_t Z, int64_t Y, int64_t X,
                      const fpl_v2o_params *p, double *h_out, void *stream);

/* Stage C (fplobjdetect.py:184-257): candidates > threshold, greedy NMS with ball suppression,
 * emission order (conf desc, flat index asc), un-pad, buffer crop, offset.
 * d_dets: capacity rows x 4 doubles (x,y,z,conf).  *h_count receives the number of rows.
 * h_stats (optional, 8 int64): candidates, rounds, ball checks, selected before crop, ...
 * Synchronises `stream`. */
int fpl_v2o_detect(fpl_ctx *ctx, const float *d_smooth, int64_t Z, int64_t Y, int64_t X,
                   const fpl_v2o_params *p, double threshold, double *d_dets, int64_t capacity,
                   int64_t *h_count, int64_t *h_stats, void *stream);

/* Stage C with a segmentation: the seg / seg_dilate / seg_force branch of voxel2obj (fplobjdetect.py:192-195, 213-224).
 * The suppression set of a selected point is  ball AND dilate_D(seg cube == seg[point])  (OR the forced inner ball of
 * radius seg_force); the loop runs sequentially on the device in one persistent CTA.  d_val / d_idx: the candidates
 * (smoothed value > threshold) sorted by (value desc, flat interior index asc); d_seg: (Z,Y,X) int64 labels (voxels
 * outside the volume count as label 0, the reference's zero padding) or NULL; d_supp: ceil(Z*Y*X/32) zeroed uint32
 * words; seg_dilate < 0: no dilation (None); seg_force <= 0: none; obj_min_dist <= 31.  d_rows: capacity x 4 doubles
 * (x, y, z, conf) after un-padding, buffer crop and offset; d_count: 3 int64 (rows written, points selected before the
 * crop, overflow flag).  Does not synchronise. */
int fpl_v2o_detect_seg(fpl_ctx *ctx, const float *d_val, const int64_t *d_idx, int64_t n_cand, const int64_t *d_seg,
                       uint32_t *d_supp, int64_t Z, int64_t Y, int64_t X, const fpl_v2o_params *p, int32_t seg_dilate,
                       int32_t seg_force, double *d_rows, int64_t capacity, int64_t *d_count, void *stream);

/* A+B+C in one call: the replacement for fplobjdetect.voxel2obj(pred, ...) on a device-resident
 * float32 probability map.  h_threshold (optional) receives the threshold used. */
int fpl_voxel2obj(fpl_ctx *ctx, const float *d_pred, int64_t Z, int64_t Y, int64_t X,
                  const fpl_v2o_params *p, double *d_dets, int64_t capacity, int64_t *h_count,
                  double *h_threshold, int64_t *h_stats, void *stream);

/* --- exact multi-GPU voxel2obj (semantics S2): building blocks for z-slab ranks.  The host
 * program (flypylib_b200/multi_gpu.py:voxel2obj_global) does the collectives between these calls: halo
 * exchange of the probability map, all-reduce of the radix histograms (np.percentile over the WHOLE padded
 * volume, fplobjdetect.py:183), all-gather of the points selected in each round of the greedy loop (:187-231). */

/* Histogram of ((key >> shift) & (bins-1)) over the values whose monotone key matches prefix under prefix_mask
 * (one level of the radix select).  d_hist: 2048 uint64 (device); d_nan (optional): NaN count (device). */
int fpl_v2o_hist_level(fpl_ctx *ctx, const float *d_v, int64_t n, uint32_t prefix, uint32_t prefix_mask,
                       int shift, int bins, uint64_t *d_hist, uint64_t *d_nan, void *stream);

/* Open a session on an extended slab (Ze,Y,X) of the smoothed map: planes [own_lo, own_hi) are owned by this
 * rank, the rest is halo (>= obj_min_dist planes towards every neighbouring rank).  Candidates = values >
 * threshold.  *session is released by fpl_v2o_slab_end. */
int fpl_v2o_slab_begin(fpl_ctx *ctx, const float *d_smooth_ext, int64_t Ze, int64_t Y, int64_t X,
                       const fpl_v2o_params *p, double threshold, int64_t own_lo, int64_t own_hi,
                       int64_t list_cap, int64_t det_cap, void **session, int64_t *h_n_candidates,
                       void *stream);
/* Decision half of one round: owned valid candidates without a better valid voxel in their ball are selected;
 * their (z,y,x) slab coordinates go to d_sel_zyx (device, sel_cap x 3 int64). */
int fpl_v2o_slab_round(void *session, int64_t *d_sel_zyx, int64_t sel_cap, int64_t *h_n_sel,
                       int64_t *h_alive_owned, void *stream);
/* Update half: suppress the balls of the points selected by ALL ranks (slab coordinates; points outside the
 * slab are legal, only the part of their ball inside counts). */
int fpl_v2o_slab_suppress(void *session, const int64_t *d_zyx, int64_t n_pts, void *stream);
/* The same round without a host round trip between its halves.  round_pack: decision half; d_block (device,
 * (1 + sel_cap) x 3 int64) receives row 0 = (points selected, valid owned candidates at round start, overflow flag)
 * and rows 1.. = (z + z_offset, y, x) of the selected points.  The caller all-gathers the blocks of all ranks (ONE
 * collective per round) and passes them to apply_blocks, which suppresses every ball that reaches this slab
 * (z_offset = global z of this slab's plane 0 in both calls).  Neither call synchronises; the caller reads the
 * gathered header rows once per round to decide termination. */
int fpl_v2o_slab_round_pack(void *session, int64_t *d_block, int64_t sel_cap, int64_t z_offset, void *stream);
int fpl_v2o_slab_apply_blocks(void *session, const int64_t *d_blocks, int32_t n_blocks, int64_t sel_cap,
                              int64_t z_offset, void *stream);
/* Close: rows (z, y, x, conf) of the owned detections in slab coordinates, unordered. */
int fpl_v2o_slab_end(void *session, double *d_rows, int64_t capacity, int64_t *h_count, int64_t *h_rounds,
                     void *stream);

/* ---------------------------------------------------------------------------------------------
 * network: model builders + FplNetwork.infer         (flypylib/fplmodels.py, fplnetwork.py:99-189)
 * ------------------------------------------------------------------------------------------- */

/* architectures (flypylib/fplmodels.py:102-136, :138-172, :258-304) */
#define FPL_ARCH_VGG_LIKE    1
#define FPL_ARCH_VGG_LIKE2   2
#define FPL_ARCH_UNET_LIKE2  3
/* further builders of flypylib/fplmodels.py (same layer vocabulary): :73-100, :206-256, :306-357, :359-410, :412-467 */
#define FPL_ARCH_BASELINE    4
#define FPL_ARCH_UNET_LIKE   5
#define FPL_ARCH_UNET_LIKE3  6
#define FPL_ARCH_UNET_LIKE4  7
#define FPL_ARCH_UNET_LIKE4B 8
/* flypylib/fplmodels.py:174-208: residual blocks (add of a cropped shortcut); bf16 and fp32 precisions */
#define FPL_ARCH_RESNET_LIKE 9
/* flypylib/fplmodels.py:470-526: U-Net without BatchNormalization */
#define FPL_ARCH_UNET_LIKE_VOL 10

/* arithmetic of the conv stack */
#define FPL_PREC_FP32  0   /* CUDA-core fp32 direct convolution (validation path)            */
#define FPL_PREC_BF16  1   /* tcgen05 kind::f16 implicit GEMM, bf16 operands, fp32 accumulate */
#define FPL_PREC_TF32  2   /* high-precision tensor-core path: bf16 hi/lo split operands, 3 bf16 contractions
                            * per convolution, fp32 accumulate (~16 mantissa bits, >= TF32 accuracy)        */

/* Build a network object for `arch`; replaces the Keras graph construction in the builders and
 * FplNetwork._set_infer (fplnetwork.py:99-110: the x rf_stride nearest up-sampling of the VGG
 * output is part of the network). */
int fpl_net_create(fpl_ctx *ctx, int arch, fpl_net **out);
int fpl_net_destroy(fpl_net *net);

/* receptive-field info the builders return: rf_size, rf_offset, rf_stride, infer_sz
 * (fplmodels.py:136,172,304) */
int fpl_net_info(const fpl_net *net, int32_t *rf_size, int32_t *rf_offset, int32_t *rf_stride,
                 int32_t *infer_sz);

/* Number of weight arrays in Keras get_weights() order and the element count of each
 * (conv kernel (kd,kh,kw,Cin,Cout); BatchNormalization gamma,beta,moving_mean,moving_variance;
 * last VGG conv kernel,bias).  Replaces Model.get_weights()/set_weights (fplnetwork.py:109-110). */
int fpl_net_num_weights(const fpl_net *net, int32_t *n);
int fpl_net_weight_size(const fpl_net *net, int32_t index, int64_t *elems);
/* h_arrays[i] points to host float32 data of array i.  Folds BN into scale/bias and packs the
 * kernels for the selected precision. */
int fpl_net_set_weights(fpl_net *net, const float *const *h_arrays, int32_t n, int precision);

/* VGG nets only: evaluate super-tiles of edge m*(infer_sz-2*rf_offset)+2*rf_offset in
 * fpl_net_infer_volume.  Their origins stay on the reference grid and the nets are shift-equivariant
 * by rf_stride, so the result equals the reference tiling with fewer halo FLOPs.
 * m = 1 (default) is the reference grid itself; the U-Net must keep m = 1. */
int fpl_net_set_tile_multiplier(fpl_net *net, int32_t m);

/* Forward pass of a batch of tiles: replaces infer_network.predict (fplnetwork.py:175-176).
 * d_tiles: n_tiles x in^3 float32 (channels = 1), in = tile input edge (any valid size for the
 * architecture); d_out: n_tiles x out^3 float32 (VGG: already x4 up-sampled). */
int fpl_net_out_size(const fpl_net *net, int32_t in_sz, int32_t *out_sz);
int fpl_net_forward_tiles(fpl_net *net, const float *d_tiles, int32_t n_tiles, int32_t in_sz,
                          float *d_out, void *stream);

/* Whole-volume inference: replaces FplNetwork.infer(image) (fplnetwork.py:136-189): reference tile
 * grid (origins k*(infer_sz-2*rf_offset)), zero padding of far-edge tiles, scatter of tile
 * interiors, rf_offset-wide border left at 0.
 * d_image: (Z,Y,X) float32 (already normalised, as the reference passes it) or uint8 with
 * norm_mean/norm_std applied on the fly ((x-mean)/std, fplobjdetect.py:1106-1107);
 * image_is_u8 selects.  z_tile_begin/z_tile_end restrict the call to a range of tile layers
 * (multi-GPU z-slab sharding; pass 0,-1 for all).  d_pred: (Z,Y,X) float32, fully written. */
int fpl_net_infer_volume(fpl_net *net, const void *d_image, int image_is_u8, float norm_mean,
                         float norm_std, int64_t Z, int64_t Y, int64_t X, int32_t z_tile_begin,
                         int32_t z_tile_end, float *d_pred, void *stream);

/* One z-slab of a larger volume: the multi-GPU / streaming form of fpl_net_infer_volume.  The reference
 * replicates the graph on n_gpu towers and feeds one tile per tower and predict step
 * (flypylib/multi_gpu.py:20-61, fplnetwork.py:130-134,175-176); here every rank (or every chunk of a host
 * volume) evaluates the planes [z0,z1) of the (Z,Y,X) image as one slab and writes ITS planes of the one
 * prediction volume: [z0+rf_offset, z1-rf_offset), plus the zero border [0,rf_offset) when z0 == 0 and
 * [Z-rf_offset,Z) when z1 == Z.  The values equal those of the whole-volume call bit for bit provided the cuts
 * lie on the network's grid: z0 a multiple of rf_stride (VGG builders, shift-equivariant) or of
 * infer_sz-2*rf_offset (U-Net builders: tile phase matters), and an inner slab holds 2*rf_offset + k*that
 * many planes (checked).  d_image_slab / d_pred_slab point at plane z0 of the image / prediction volume;
 * only the planes listed above are written, so slabs of neighbouring calls may share one buffer. */
int fpl_net_infer_slab(fpl_net *net, const void *d_image_slab, int image_is_u8, float norm_mean,
                       float norm_std, int64_t Z, int64_t z0, int64_t z1, int64_t Y, int64_t X,
                       float *d_pred_slab, void *stream);

/* ---------------------------------------------------------------------------------------------
 * data-parallel training step of the VGG builders (BASELINE config 5)
 *   replaces one batch of Keras fit_generator (flypylib/fplnetwork.py:112-128) with the default
 *   compile args loss='binary_crossentropy', optimizer='adam' (fplnetwork.py:74-77)
 * ------------------------------------------------------------------------------------------- */
typedef struct fpl_trainer fpl_trainer;

/* patch_sz = rf_size of the builder (18 / 24: one output voxel per patch, fplobjdetect.py:77),
 * batch = patches per GPU (64 in scripts/fpl_cx1_0_vgg_4ss.py:11-17) */
int fpl_train_create(fpl_ctx *ctx, int arch, int patch_sz, int batch, fpl_trainer **out);
int fpl_train_destroy(fpl_trainer *t);
/* Arithmetic of the convolution contractions (forward, dgrad, wgrad) of the step.  FPL_PREC_TF32 (default): tcgen05
 * tensor cores on bf16 hi/lo split operands, three bf16 contractions with fp32 accumulation -- fp32-class results
 * (the reference trains in float32); FPL_PREC_BF16: one bf16 contraction; FPL_PREC_FP32: fp32 CUDA-core kernels
 * (validation path). */
int fpl_train_set_precision(fpl_trainer *t, int precision);
/* n_params: floats of the flat parameter vector (Keras get_weights() order, concatenated);
 * n_bn: floats of the batch-statistics vector (per BN layer: mean[C], biased var[C]) */
int fpl_train_sizes(const fpl_trainer *t, int64_t *n_params, int64_t *n_bn);
/* forward (training mode: BN batch statistics per rank, Dropout(0.5) from a counter-based hash of
 * dropout_seed) + loss + backward.  d_x: (batch, s,s,s) float32; d_labels: batch uint8 in {0,1};
 * d_params / d_grads: n_params floats (grads of non-trainable slots are 0); loss_scale = 1/global_batch
 * (Keras mean over the whole batch; gradients of all ranks are then SUMMED by the caller's all-reduce).
 * h_loss_sum = sum of per-sample binary cross-entropies of this rank, h_correct = #(round(p) == y). */
int fpl_train_forward_backward(fpl_trainer *t, const float *d_x, const uint8_t *d_labels, const float *d_params,
                               float *d_grads, float *d_bn_batch, float loss_scale, uint64_t dropout_seed,
                               double *h_loss_sum, int64_t *h_correct, void *stream);
/* Adam step (Keras: lr_t = lr*sqrt(1-b2^t)/(1-b1^t), p -= lr_t*m/(sqrt(v)+eps)) on the trainable slots and
 * moving-average update of the BN statistics (moving = moving*momentum + batch*(1-momentum)). */
int fpl_train_apply(fpl_trainer *t, float *d_params, const float *d_grads, float *d_m, float *d_v,
                    const float *d_bn_batch, int64_t step, float lr, float beta1, float beta2, float eps,
                    float bn_momentum, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FPL_B200_H */
