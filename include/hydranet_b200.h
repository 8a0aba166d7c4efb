/* hydranet_b200 -- C ABI of the B200-native HydraNet hot path.
 *
 * Boundary: the reference's Python `HydraNet.forward` (model/model.py:159-198)
 * and the three static decoders (head_seg/segmentation.py:107-125, head_detect/detection.py:232-245
 * -> head_detect/detection_loss.py:70-108, head_lane/lanedetect.py:103-116 ->
 * head_lane/lane_codec.py:116-219 + lane_codec_utils.py:487-542) bottom out in the entry points
 * below.  Plain pointers and sizes only; device pointers are raw CUDA addresses, `stream` is a
 * cudaStream_t passed as void*.  Every function returns 0 on success or an HN_ERR_* code
 * (cf. the reference's own int-returning C API, deploy/src/interface/Hydranet.h:83-111);
 * hn_last_error() returns a thread-local message.  There is no CPU fallback behind any of them.
 *
 * Activations are NHWC bf16 "views": a base pointer plus element strides, so padded buffers,
 * channel slices and strided (stride-2) sub-samplings are all expressed without copies.
 */
#ifndef HYDRANET_B200_H
#define HYDRANET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HN_OK 0
#define HN_ERR_ARG 1
#define HN_ERR_CUDA 2
#define HN_ERR_UNSUPPORTED 3

#define HN_ACT_NONE 0
#define HN_ACT_RELU 1
#define HN_ACT_SWISH 2
#define HN_ACT_ELU 3
#define HN_ACT_SIGMOID 4

#define HN_HALO_NONE 0
#define HN_HALO_REFLECT 1   /* nn.ReflectionPad2d(1): head_seg/segmentation.py:40 */
#define HN_HALO_REPLICATE 2 /* reflect-pad of a nearest-x2 upsample == replicate-pad of its source */

#define HN_EPI_STD 0
#define HN_EPI_SEGOUT 1 /* sub-pixel seg logits (fp32 NCHW) + fused argmax */

#define HN_MAX_SRC 6
#define HN_MAX_TAPS 96
#define HN_MAX_GROUPS 8

/* 4-D NHWC bf16 view; channel stride is 1; strides in elements. */
typedef struct hn_view {
    const void* ptr;
    int32_t N, H, W, C;
    int64_t stride_n, stride_y, stride_x;
} hn_view;

/* One K-step (64 input channels) of the implicit GEMM: which source, which spatial shift, which
 * channel offset.  The packed weight matrix stores its K columns in exactly this order. */
typedef struct hn_tap {
    int8_t src, dy, dx, rsv0;
    int16_t c0, rsv1;
} hn_tap;

/* Convolution as implicit GEMM on tcgen05 (replaces every nn.Conv2d(+BatchNorm2d)(+activation)
 * call of model/net/anynet.py:8-76, net/common.py:35-114, net/bifpn.py:57-99,
 * head_seg/segmentation.py:16-48, head_detect/detection.py:28-83, head_lane/lanedetect.py:45-64).
 * out[n, y*s+oy, x*s+ox, :cout] = act(sum_taps src[tap.src][n, y+dy, x+dx, c0:c0+64] . W + bias) (+res) */
typedef struct hn_conv_desc {
    hn_view src[HN_MAX_SRC];
    int32_t n_src;
    const void* weight; /* bf16 [w_rows][num_taps*64], row = output channel */
    int32_t w_rows;
    int32_t num_taps;
    hn_tap taps[HN_MAX_TAPS];
    int32_t flat;           /* 1: rows = src[0].W consecutive pixels (H=N=1), tile = 128 rows */
    int32_t tile_h, tile_w; /* spatial tile, tile_h*tile_w == 128 */
    int32_t n_img, out_h, out_w; /* tile-space extent per image (before out_scale) */
    int32_t flat_hw;        /* flat: rows per image (for out_stride_n addressing) */
    int32_t cout, bn, stages;
    const float* bias; /* fp32 [>= n_tiles*bn] */
    int32_t act;
    int32_t epi;
    void* out;
    int32_t out_fp32;
    int64_t out_stride_n, out_stride_y, out_stride_x; /* elements of the output dtype */
    int32_t out_scale, out_oy, out_ox;
    int32_t halo;
    const void* res; /* bf16 residual view of the output pixels (own strides) */
    int64_t res_stride_n, res_stride_y, res_stride_x;
    int32_t res_relu;
    int32_t grouped; /* 1: block-diagonal (grouped) conv: tap channel offset += n-tile origin; needs bn == 64 */
    void* out2;      /* HN_EPI_SEGOUT: uint8 class map [N][2*out_h][2*out_w] or NULL */
    int32_t n_cls;   /* HN_EPI_SEGOUT: classes per sub-pixel (columns = 4 parities x 8) */
    /* Row groups (flat mode): one launch over several stacked problems that share the weights -- the five pyramid
     * levels of a detection tower (detection.py:30-37: shared conv, per-level BatchNorm).  Rows [group_end[g-1],
     * group_end[g]) form group g. */
    int32_t n_groups; /* 0 = off */
    int32_t group_end[HN_MAX_GROUPS];
    const float* group_scale; /* fp32 [n_groups][n_tiles*bn]: epilogue = act(acc * scale + shift); NULL = plain bias */
    const float* group_shift;
    int32_t group_addr;       /* 1: output row address = n*out_stride_n + group_out_base[g] + pix*out_stride_x with
                                 (n, pix) = divmod(row - group start, group_hw[g]) (fp32 head outputs) */
    int32_t group_hw[HN_MAX_GROUPS];
    int64_t group_out_base[HN_MAX_GROUPS];
} hn_conv_desc;

/* Stem: 3x3 s2 p1 conv 3->32 + folded BN + ReLU, fp32 NCHW in, bf16 NHWC out (anynet.py:8-20). */
typedef struct hn_stem_desc {
    const float* x; /* [N][3][H][W] */
    int32_t N, H, W;
    const float* w; /* [27][32] fp32, BN folded: index (ci*9+ky*3+kx)*32+co */
    const float* b; /* [32] */
    hn_view out;    /* [N][H/2][W/2][32] */
    int32_t no_relu; /* 1: raw convolution output (training: BatchNorm follows as its own op) */
} hn_stem_desc;

#define HN_IN_SAME 0
#define HN_IN_UP2 1  /* nearest x2 of a half-resolution view */
#define HN_IN_POOL 2 /* 3x3 s2 max over a double-resolution view zero-padded right/bottom (common.py:138-151) */

/* BiFPN fusion node front half: weighted sum -> swish -> depthwise 3x3 (zero pad 1)
 * (net/bifpn.py:170-231 + common.py:91-101).  With n_in==1, w={1}, swish=0 it is the plain
 * depthwise conv of the detection towers (detection.py:33-37). */
typedef struct hn_node_desc {
    int32_t n_in;
    hn_view in[3];
    int32_t mode[3];
    float w[3];
    int32_t swish;
    const float* dw; /* fp32 [9][C] */
    hn_view out;
} hn_node_desc;

/* Plain depthwise 3x3 (zero pad 1) over up to HN_MAX_GROUPS independent (input, output) view pairs that share the
 * weights: all pyramid levels of a detection-tower layer in one launch (detection.py:33-37). */
typedef struct hn_dw_multi_desc {
    int32_t n;
    hn_view in[HN_MAX_GROUPS];
    hn_view out[HN_MAX_GROUPS];
    const float* dw; /* fp32 [9][C] */
} hn_dw_multi_desc;

#define HN_POOL_ZERO_RB 0 /* MaxPool2dStaticSamePadding(3,2): zero pad right/bottom, zeros take part */
#define HN_POOL_NEGINF 1  /* nn.MaxPool2d(3,2,padding=1) (lanedetect.py:41) */
typedef struct hn_pool_desc {
    hn_view in, out;
    int32_t mode;
} hn_pool_desc;

/* Lane fuse to stride 32: cat(maxpool(maxpool(P3)), maxpool(P4), P5, up2(P6)) (lanedetect.py:76-80)
 * or to stride 16: cat(maxpool(P3), up2(P5), P4, up4(P6)) (lanedetect.py:70-74). */
typedef struct hn_lanefuse_desc {
    hn_view p3, p4, p5, p6, out;
    int32_t stride; /* 16 or 32 */
} hn_lanefuse_desc;

/* Squeeze-excite (anynet.py:39-47,68-69) = hn_se_pool_fwd (pool + both FC layers) -> hn_se_scale_fwd.
 * pool: mean over H*W of every channel, bf16 [N][C].  Deterministic: per-chunk partial sums are added in
 * chunk order by the block that arrives last for its image.  `counter` must be zero on first use.
 * With S > 0 that block also runs the two FC layers of its image: hidden = ReLU(W1 . mean + b1) (rounded to
 * bf16), gate = sigmoid(W2 . hidden + b2) -> bf16 [N][C]: one launch instead of three. */
typedef struct hn_se_pool_desc {
    hn_view x;
    int32_t pix_per_block; /* pixels summed by one block (multiple of 128) */
    float* partial;   /* scratch fp32 [N][ceil(H*W/pix_per_block)][C] */
    int32_t* counter; /* scratch int32 [N] */
    void* mean;       /* bf16 [N][C] */
    int32_t S;        /* hidden width padded to a multiple of 8 (zero rows / columns); 0 = pool only */
    const void* w1;   /* bf16 [S][C] */
    const float* b1;  /* fp32 [S] */
    const void* w2;   /* bf16 [C][S] */
    const float* b2;  /* fp32 [C] */
    void* gate;       /* bf16 [N][C] */
} hn_se_pool_desc;

/* scale: x[n, :, :, c] *= scale[n][c] in place, scale bf16 [N][C] */
typedef struct hn_se_scale_desc {
    hn_view x;
    const void* scale;
} hn_se_scale_desc;

/* An XBlock's grouped 3x3 convolution (group width 8, stride 1, BN folded, ReLU; anynet.py:60-66) and its squeeze-excite
 * (anynet.py:39-47,68-69) in ONE launch for small maps (hn_gconv_se_supported): a cluster of 4 CTAs per image, each owning a
 * quarter of the groups; the 3x3 runs as warp-level tensor-core MMAs out of shared memory, the result is pooled, gated and
 * scaled before it is written to se.x.  `weight`: bf16 [C/8][10][8 out][8 in] (tap = ky*3+kx, the 10th tap zero);
 * `bias`: fp32 [C].  `in` and se.x have the same shape and must not alias. */
typedef struct hn_gconv_se_desc {
    hn_view in;
    const void* weight;
    const float* bias;
    hn_se_pool_desc se;
} hn_gconv_se_desc;

/* Image pre-processing, the step in front of HydraNet.forward (model/demo.py:191-196 with imagenet_normalize
 * demo.py:26-40; C++ twin deploy/hydranet_model.cpp:159-248): BGR -> RGB, cv2.resize (uint8 INTER_LINEAR in OpenCV's
 * 11-bit fixed point; INTER_AREA for an exact 2x2 down-scale), (x/255 - mean)/std in float64 rounded to fp32,
 * HWC -> planar.  Bit-exact against cv2 4.13 + numpy (oracle/preprocess_ref.py). */
typedef struct hn_preprocess_desc {
    const uint8_t* src;  /* [N] images of src_h x src_w x 3 (B,G,R), row pitch src_pitch bytes, image stride src_stride bytes */
    int32_t N, src_h, src_w;
    int64_t src_pitch, src_stride;
    float* dst;          /* [N][3 (R,G,B)][dst_h][dst_w] fp32, contiguous: the `x` of HydraNet.forward */
    int32_t dst_h, dst_w;
} hn_preprocess_desc;

/* Detection decode + NMS (detection_loss.py:7-108; torchvision.ops.boxes.batched_nms semantics). */
#define HN_NMS_AUTO_CUDA 0 /* coordinate trick iff 4*n <= 100000 (torchvision boxes.py, CUDA tensors) */
#define HN_NMS_AUTO_CPU 1  /* coordinate trick iff 4*n <= 4000 */
#define HN_NMS_TRICK 2
#define HN_NMS_VANILLA 3
typedef struct hn_det_desc {
    const float* anchors;        /* [A][4] y1,x1,y2,x2 */
    const float* regression;     /* [N][A][4] dy,dx,dh,dw */
    const float* classification; /* [N][A][ncls] post-sigmoid */
    int32_t N, A, ncls;
    int32_t img_h, img_w;
    float conf_thres, iou_thres;
    int32_t nms_mode;
    /* workspace, device: see hn_det_workspace_bytes */
    void* workspace;
    int64_t workspace_bytes;
    /* outputs, device */
    float* out_boxes;   /* [N][A][4] xyxy, kept boxes in score-descending order */
    float* out_scores;  /* [N][A] */
    int64_t* out_class; /* [N][A] */
    int32_t* out_count; /* [N] kept per image */
    int32_t* out_cand;  /* [N] candidates over threshold per image (diagnostic) */
    /* optional: skip the decode and use these pre-decoded candidates ("identical pre-NMS inputs") */
    const float* pre_boxes; /* [N][A][4] or NULL */
} hn_det_desc;

/* Lane decode + lane NMS (lane_codec.py:116-219, lane_codec_utils.py:487-542). One image per row. */
typedef struct hn_lane_desc {
    const float* cls; /* [N][fh*fw][2] logits, or probabilities if cls_is_prob */
    const float* loc; /* [N][fh*fw][2*ppl+2] */
    int32_t N, fh, fw, ppl;
    int32_t cls_is_prob;
    float conf_thres, nms_thres;
    int32_t use_mean;
    double step_w, interval, points_per_anchor; /* Python floats of the codec (lane_codec.py:33-46) */
    float input_width, margin_width;
    void* workspace; /* device scratch, hn_lane_workspace_bytes(N, fh*fw, ppl) */
    /* outputs, device */
    int32_t* out_count; /* [N] lanes after NMS */
    int32_t* out_meta;  /* [N][fh*fw][4]: anchor index, start_pos, end_pos, n_points */
    float* out_prob;    /* [N][fh*fw] */
    float* out_x;       /* [N][fh*fw][ppl] x of every point, ordered bottom(start_pos) -> top */
    int32_t* out_cand;  /* [N] lanes before NMS */
} hn_lane_desc;

/* ---- single-shot entry points (each enqueues on `stream` and returns) ---- */
int hn_conv_fwd(const hn_conv_desc* d, void* stream);
int hn_stem_fwd(const hn_stem_desc* d, void* stream);
void hn_stem_set_mma(int on); /* 1 (default): tensor-core stem with (hi, lo) bf16 splits; 0: the fp32 CUDA-core kernel */
int hn_node_fwd(const hn_node_desc* d, void* stream);
int hn_dw_multi_fwd(const hn_dw_multi_desc* d, void* stream);
int hn_pool_fwd(const hn_pool_desc* d, void* stream);
int hn_lanefuse_fwd(const hn_lanefuse_desc* d, void* stream);
int hn_se_pool_fwd(const hn_se_pool_desc* d, void* stream);
int hn_se_scale_fwd(const hn_se_scale_desc* d, void* stream);
/* The whole squeeze-excite of a block (anynet.py:39-47,68-69) in ONE launch, for small maps: a cluster of 4 CTAs per image, each
 * owning a quarter of the channels: pool -> FC1 + ReLU -> FC2 + sigmoid -> x *= gate in place.  Same descriptor as
 * hn_se_pool_fwd (S > 0 required); `partial`, `pix_per_block` and `counter` are unused.
 * Outputs: mean, gate, and x scaled in place.  hn_se_fused_supported tells whether a shape qualifies (H*W <= 4096,
 * C >= 64, the channel slice fits in shared memory). */
int hn_se_fused_fwd(const hn_se_pool_desc* d, void* stream);
int hn_se_fused_supported(int32_t H, int32_t W, int32_t C, int32_t S);
int hn_gconv_se_fwd(const hn_gconv_se_desc* d, void* stream);
int hn_gconv_se_supported(int32_t H, int32_t W, int32_t C, int32_t S);
/* profiling aid: device buffer of 8 int64 globaltimer stamps per CTA (N * 4 CTAs) written at the phase boundaries of the two
 * fused launches above (start, operands staged, conv done, pooled, cluster barrier 1, FC1 + barrier 2, gate done, end); NULL = off */
void hn_se_fused_set_debug(long long* buf);
int hn_preprocess_fwd(const hn_preprocess_desc* d, void* stream);
int hn_seg_argmax(const float* logits, int32_t N, int32_t C, int64_t HW, int64_t* out_i64, uint8_t* out_u8,
                  void* stream);
int hn_u8_to_i64(const uint8_t* in, int64_t* out, int64_t n, void* stream);
int64_t hn_det_workspace_bytes(int32_t N, int32_t A);
int hn_det_decode_nms(const hn_det_desc* d, void* stream);
int64_t hn_lane_workspace_bytes(int32_t N, int32_t n_anchor, int32_t ppl);
int hn_lane_decode_nms(const hn_lane_desc* d, void* stream);

/* Host tails of the decoders on the device.
 * hn_det_invert_affine: DetectionHeader.invert_affine (detection.py:218-230) in place on the kept boxes of hn_det_decode_nms;
 *   scale_xy: device fp32 [N][2] = float32(new_w / old_w), float32(new_h / old_h).
 * hn_lane_scale_to_org: LaneHeader.scale_to_org's arithmetic (lanedetect.py:118-124, lane_codec_utils.py:128-182,236-282) for the
 *   kept lanes of hn_lane_decode_nms: ordering keys [N][n_anchor][4] = (cross_x, k, first x, last x) and the points scaled to
 *   the original frame (x fp32, y fp64, as numpy / Python compute them).  The final ordering (a non-transitive comparator
 *   under Python's sort) and the dict stay on the host. */
int hn_det_invert_affine(float* boxes, const int32_t* count, int32_t N, int32_t A, const float* scale_xy, void* stream);
int hn_lane_scale_to_org(const int32_t* count, const int32_t* meta, const float* xs, int32_t N, int32_t n_anchor, int32_t ppl,
                         double input_height, double interval, double cross_y, double scale_x, double scale_y, float* keys, float* x_out,
                         double* y_out, void* stream);

/* ---- plan: a recorded schedule of the ops above, replayed with one call (or as a CUDA graph) ---- */
typedef struct hn_plan hn_plan;
int hn_plan_create(hn_plan** out);
int hn_plan_destroy(hn_plan* p);
int hn_plan_add_conv(hn_plan* p, const hn_conv_desc* d);
int hn_plan_set_branch(hn_plan* p, int branch); /* ops added next belong to `branch`: 0 = trunk, 1..4 = independent branches forked after the trunk by hn_plan_run (e.g. the three heads) */
int hn_plan_add_wait(hn_plan* p, int branch); /* the current branch continues only after branch `branch` (a smaller index) has completed */
int hn_plan_add_stem(hn_plan* p, const hn_stem_desc* d);
int hn_plan_add_node(hn_plan* p, const hn_node_desc* d);
int hn_plan_add_dw_multi(hn_plan* p, const hn_dw_multi_desc* d);
int hn_plan_add_pool(hn_plan* p, const hn_pool_desc* d);
int hn_plan_add_lanefuse(hn_plan* p, const hn_lanefuse_desc* d);
int hn_plan_add_se_pool(hn_plan* p, const hn_se_pool_desc* d);
int hn_plan_add_se_scale(hn_plan* p, const hn_se_scale_desc* d);
int hn_plan_add_se_fused(hn_plan* p, const hn_se_pool_desc* d);
int hn_plan_add_gconv_se(hn_plan* p, const hn_gconv_se_desc* d);
int hn_plan_add_det(hn_plan* p, const hn_det_desc* d);
int hn_plan_add_lane(hn_plan* p, const hn_lane_desc* d);
int hn_plan_size(const hn_plan* p);
int hn_plan_num_launches(const hn_plan* p); /* kernels one replay launches */
int hn_plan_run(hn_plan* p, void* stream);
int hn_plan_run_range(hn_plan* p, int first, int last, void* stream);
int hn_plan_graph_capture(hn_plan* p, void* stream); /* instantiate a CUDA graph of the whole plan */
int hn_plan_graph_launch(hn_plan* p, void* stream);


/* =====================================================================================================================
 * Training step (reference: model/train.py:241-269 -- forward in train mode, cal_loss,
 * backward, Adam).  The reference runs it through PyTorch autograd over cuDNN; here every layer's forward / backward is
 * one of the entry points below (plus hn_conv_fwd, which also serves as the data-gradient GEMM: dgrad of a convolution
 * is a convolution of the output gradient with the transposed, spatially flipped filter).
 * Activations and activation gradients: bf16, NHWC, as `hn_view`s or as row matrices `hn_mat` ([pixels][channels]).
 * Parameters and parameter gradients: fp32 in the reference's own layouts (OIHW filters).
 * ===================================================================================================================== */

/* bf16 row-major matrix [rows][cols], `ld` elements between rows; cols and ld multiples of 8, ptr 16-byte aligned. */
typedef struct hn_mat {
    void* ptr;
    int64_t rows;
    int32_t cols;
    int64_t ld;
} hn_mat;

#define HN_MAX_SEG 8

/* BatchNorm2d in training mode (+ activation, + residual): nn.BatchNorm2d.forward with batch statistics and the
 * running-statistics update (anynet.py:14-17,31-60; common.py:95-99; detection.py:20-24 per-level lists; lanedetect.py:48).
 * Row segments carry independent statistics / parameters: the pyramid levels of a detection tower layer are one
 * stacked matrix with five BatchNorms.
 *   fwd: mean/var over the rows of each segment (deterministic two-level sum), running stats updated in place
 *        (momentum, unbiased variance), y = act(z * scale + shift [+ res]).
 *   bwd: dz_act = dy * act'(.), dgamma = sum dz_act * xhat, dbeta = sum dz_act,
 *        dz = gamma * invstd * (dz_act - mean(dz_act) - xhat * mean(dz_act * xhat)); dres = dz_act (optional). */
typedef struct hn_bn_desc {
    hn_mat z;
    int32_t n_seg;
    int64_t seg_end[HN_MAX_SEG]; /* exclusive row end of each segment */
    const float* gamma[HN_MAX_SEG];
    const float* beta[HN_MAX_SEG];
    float* running_mean[HN_MAX_SEG]; /* may be NULL (no update) */
    float* running_var[HN_MAX_SEG];
    float eps, momentum;
    float* stats;   /* fp32 [n_seg][4][cols]: mean, invstd, scale, shift -- written by fwd, read by bwd */
    int32_t act;    /* HN_ACT_NONE / HN_ACT_RELU / HN_ACT_SWISH */
    hn_mat res;     /* optional residual (ptr NULL = none), added before the activation */
    hn_mat y;       /* fwd output; bwd: the saved output (ReLU mask) */
    float* scratch; /* device scratch for partial sums */
    int64_t scratch_bytes;
    hn_mat dy;      /* bwd in */
    hn_mat dz;      /* bwd out: gradient of z */
    hn_mat dres;    /* bwd out (optional): gradient of the residual input */
    float* dgamma[HN_MAX_SEG]; /* fp32 [cols], written */
    float* dbeta[HN_MAX_SEG];
} hn_bn_desc;
int hn_bn_train_fwd(const hn_bn_desc* d, void* stream);
int hn_bn_train_bwd(const hn_bn_desc* d, void* stream);

/* Column reductions over the rows of `a` (deterministic): per segment of `rows_per_seg` consecutive rows
 * (0 = one segment).  mode 0: out0 = sum a, out1 = sum a^2 (out1 may be NULL); mode 1: out0 = sum a*b.
 * Serves bias gradients, squeeze-excite pooling (segments = images) and the fusion-weight gradients. */
int hn_col_reduce(const hn_mat* a, const hn_mat* b, int32_t mode, int64_t rows_per_seg, float* out0, float* out1,
                  float scale, float* scratch, int64_t scratch_bytes, void* stream);

/* dz = dy * act'(ref): ref is the layer OUTPUT for ReLU / ELU / sigmoid and the PRE-activation for swish.
 * Up to 3 extra scaled copies out[k] = w[k] * dz (the inputs of a BiFPN fusion node, bifpn.py:170-231). */
typedef struct hn_actbwd_desc {
    hn_mat dy, ref, dz;
    int32_t act;
    int32_t n_scaled;
    hn_mat scaled[3];
    const float* w; /* device, fp32 [n_scaled] (the normalised fusion weights live on the device: no host sync) */
} hn_actbwd_desc;
int hn_act_bwd(const hn_actbwd_desc* d, void* stream);

/* s = sum_i w[i] * in[i]; a = swish(s) (both stored: s feeds the backward).  bifpn.py:170-231. */
typedef struct hn_wsum_desc {
    int32_t n_in;
    hn_mat in[3];
    const float* w; /* device, fp32 [n_in] */
    hn_mat s, a;
} hn_wsum_desc;
int hn_wsum_swish_fwd(const hn_wsum_desc* d, void* stream);

/* Re-sampling ops and their adjoints.  mode: HN_RS_UP2 nearest x2 (F.interpolate / upsample), HN_RS_POOL_ZERO
 * (MaxPool2dStaticSamePadding(3,2), common.py:117-151), HN_RS_POOL_NEGINF (nn.MaxPool2d(3,2,1), lanedetect.py:41).
 * bwd: din = adjoint(dout); pooling routes each output gradient to the FIRST maximum of its window in scan order
 * (PyTorch's max_pool2d backward); `x` is the forward input. */
#define HN_RS_UP2 0
#define HN_RS_POOL_ZERO 1
#define HN_RS_POOL_NEGINF 2
typedef struct hn_resample_desc {
    int32_t mode;
    hn_view x;    /* forward input */
    hn_view y;    /* forward output (fwd: written) */
    hn_view dy;   /* bwd in */
    hn_view dx;   /* bwd out */
} hn_resample_desc;
int hn_resample_fwd(const hn_resample_desc* d, void* stream);
int hn_resample_bwd(const hn_resample_desc* d, void* stream);

/* Segmentation-decoder input assembly (segmentation.py:84-105): out = ReflectionPad2d(1)(cat(up2(low), skip)),
 * either part optional (low NULL: plain reflect pad of skip).  bwd folds the halo back and sums the 2x2 blocks. */
typedef struct hn_seggather_desc {
    hn_view low;   /* [N][H/2][W/2][Cl] or ptr NULL */
    hn_view skip;  /* [N][H][W][Cs] or ptr NULL */
    hn_view out;   /* [N][H+2][W+2][Cl+Cs] (fwd out / bwd in: gradient of the padded tensor) */
    hn_view dlow, dskip; /* bwd out */
} hn_seggather_desc;
int hn_seggather_fwd(const hn_seggather_desc* d, void* stream);
int hn_seggather_bwd(const hn_seggather_desc* d, void* stream);

/* fp32 head tensors <-> bf16 gradient rows.  The reference's head outputs are [B, sum_l H_l*W_l*A, k] (detection.py:40-43),
 * [B, H*W, k] (lanedetect.py:84-96) and NCHW logits (segmentation.py:105); their gradients arrive from the PyTorch
 * losses in those layouts.  dz[row][c] = dout[addr(row) + c * stride_c] * act'(out[...]) for c < cols_valid, 0 up to dz.cols.
 * Row addressing: plain (n = row / rows_per_img, pix = row % rows_per_img: n*stride_n + pix*stride_pix) or by row groups
 * as in hn_conv_desc (group_addr). */
typedef struct hn_headgrad_desc {
    const float* dout;
    const float* out; /* saved forward output (sigmoid) or NULL */
    int32_t act;      /* HN_ACT_NONE or HN_ACT_SIGMOID */
    int32_t cols_valid;
    int64_t stride_n, stride_pix, stride_c;
    int64_t rows_per_img;
    int32_t n_groups;
    int64_t group_end[HN_MAX_GROUPS];
    int64_t group_hw[HN_MAX_GROUPS];
    int64_t group_out_base[HN_MAX_GROUPS];
    hn_mat dz;
} hn_headgrad_desc;
int hn_head_grad(const hn_headgrad_desc* d, void* stream);

/* Depthwise 3x3 weight gradient: dw[tap][c] = sum_{n,y,x} dy[n,y,x,c] * x[n,y+ky-1,x+kx-1,c] (zero padding),
 * fp32 [9][C]; accumulate != 0 adds to dw (shared filters applied to several pyramid levels). */
int hn_dw_wgrad(const hn_view* x, const hn_view* dy, float* dw, int32_t accumulate, float* scratch, int64_t scratch_bytes,
                void* stream);

/* Stem weight gradient (anynet.py:8-20): dW[co][ci][ky][kx] = sum dz[n,oy,ox,co] * x[n,ci,2oy+ky-1,2ox+kx-1]. */
int hn_stem_wgrad(const float* x, int32_t N, int32_t H, int32_t W, const hn_view* dz, float* dw, float* scratch,
                  int64_t scratch_bytes, void* stream);

/* Squeeze-excite in training (anynet.py:39-47,68-69), fp32 vectors: mean [N][C] -> h = relu(W1 mean + b1) [N][S]
 * -> gate = sigmoid(W2 h + b2) [N][C]; parameters fp32 in the reference layout ([S][C], [C][S]). */
typedef struct hn_sefc_desc {
    int32_t N, C, S;
    const float* mean;
    const float *w1, *b1, *w2, *b2;
    float *h, *gate;
    /* bwd */
    const float* dgate; /* [N][C] gradient of the gate */
    float* dmean;       /* [N][C] out */
    float *dw1, *db1, *dw2, *db2; /* out (written) */
    float* tmp;         /* scratch fp32 [N][C + S] */
} hn_sefc_desc;
int hn_se_fc_fwd(const hn_sefc_desc* d, void* stream);
int hn_se_fc_bwd(const hn_sefc_desc* d, void* stream);
/* y[n, pix, c] = x[n, pix, c] * gate[n][c] (+ add[n][c]);  rows_per_img rows per image. */
int hn_se_apply(const hn_mat* x, const float* gate, const float* add, int64_t rows_per_img, const hn_mat* y, void* stream);

/* Weight packing for the GEMM kernels, all layers in one launch: fp32 parameters -> bf16 K-major blocks.
 * One entry = one 64-wide K block of one packed matrix: dst[r][j] = src[off + r*s_r + j*s_c] for r < rows, j < cols
 * (0 elsewhere up to rows_pad x 64); grouped: only the 8x8 diagonal blocks of a group-width-8 convolution. */
typedef struct hn_pack_entry {
    const float* src;
    void* dst;       /* bf16, first element of the block (row 0, column j0) */
    int64_t dst_ld;  /* elements between packed rows */
    int32_t rows, rows_pad, cols;
    int64_t s_r, s_c; /* source strides (elements) */
    int32_t grouped;  /* 0 plain; 1 forward grouped (row = co, col = ci within the 64-block); 2 dgrad grouped */
    int32_t rsv;
} hn_pack_entry;
int hn_pack_weights(const hn_pack_entry* entries_device, int32_t n, int32_t max_rows_pad, void* stream);

/* Convolution weight gradient on tcgen05: dW[co][tap][ci] += sum_pixels dY[pix][co] * X[pix + tap][ci].
 * Both operands are read through TMA as MN-major tiles ([pixel][channel] rows), K = pixels; split-K partial sums are
 * added with fp32 reductions into `dw`, which must be zero (or hold the value to accumulate onto) on entry.
 * Element (co, tap t, ci) lives at dw[co*s_co + ci*s_ci + tap_off[t]]. */
typedef struct hn_wgrad_desc {
    hn_view dy;              /* [N][H][W][Cout] (or flat rows: N=H=1) */
    hn_view src[HN_MAX_SRC]; /* input views, as in the forward conv */
    int32_t n_src;
    int32_t num_taps;
    hn_tap taps[HN_MAX_TAPS]; /* (source, dy, dx, c0): c0 = first input channel of this K block in its source */
    int64_t tap_off[HN_MAX_TAPS]; /* dw offset of (co = 0, ci = c0) for this tap */
    int32_t tap_cin[HN_MAX_TAPS]; /* valid input channels of the block (<= 64) */
    int32_t flat;
    int32_t tile_h, tile_w;
    int32_t cout;
    int64_t s_co, s_ci;
    int32_t grouped; /* 1: group width 8 -- only ci in the row's group are stored, at ci_local * s_ci */
    float* dw;
} hn_wgrad_desc;
int hn_conv_wgrad(const hn_wgrad_desc* d, void* stream);

/* Detection loss (head_detect/detection_loss.py:111-267 FocalLoss): IoU anchor assignment (< 0.4
 * negative, >= 0.5 positive, between ignored), focal classification loss (alpha, gamma) and smooth-L1 box loss (beta 1/9) per
 * image -- cls_loss[b], reg_loss[b] normalised as the reference does -- and the gradients of mean_b(cls_loss) w.r.t. the
 * classification tensor [B][A][K] and of mean_b(reg_loss) w.r.t. the regression tensor [B][A][4], in three launches.
 * anchors [A][4] (y1,x1,y2,x2); annotations [B][M][5] (x1,y1,x2,y2,class), class -1 = padding, M <= 64.
 * workspace: B*A*4 + B*24 + 64 bytes. */
int hn_det_loss(const float* classification, const float* regression, const float* anchors, const float* annotations, int32_t B, int32_t A,
                int32_t K, int32_t M, float alpha, float gamma, void* workspace, int64_t workspace_bytes, float* cls_loss, float* reg_loss,
                float* dcls, float* dreg, void* stream);

/* Segmentation loss (head_segment/segmentation_loss.py:48-65 with use_top_k): class-weighted cross-entropy
 * per pixel (ignore_index pixels count as 0), the k largest values of every image, mean over the N*k kept values.  The k-th
 * value is found by a 3-pass radix select (no sort); ties at the threshold share their weight in the gradient.
 * logits fp32 [N][C][HW], target int64 [N][HW], weight fp32 [C].  hn_seg_loss_fwd writes loss[0] and keeps what the backward needs
 * in the workspace (hn_seg_loss_workspace_bytes); hn_seg_loss_bwd writes dlogits = gout[0] * dloss/dlogits (gout: device scalar). */
typedef struct hn_segloss_desc {
    const float* logits;
    const int64_t* target;
    const float* weight;
    int32_t N, C;
    int64_t HW, k;
    int32_t ignore_index;
    void* workspace;
    int64_t workspace_bytes;
    float* loss;          /* fwd: [1] */
    const float* gout;    /* bwd: [1] on the device */
    float* dlogits;       /* bwd: [N][C][HW] */
} hn_segloss_desc;
int64_t hn_seg_loss_workspace_bytes(int32_t N, int64_t HW);
int hn_seg_loss_fwd(const hn_segloss_desc* d, void* stream);
int hn_seg_loss_bwd(const hn_segloss_desc* d, void* stream);

/* Lane losses (head_lane/lanedetect_loss.py:18-78) in three launches: two-class log-softmax with online
 * hard-negative mining (negative_num = clamp(n_pos * negative_ratio, 1, n_neg) negatives with the smallest background
 * log-probability, found by a radix select; ties at the threshold included, as `bg <= kth` does) and the masked Huber loss of
 * the positive anchors (entries weighted_index, weighted_index + 1 of a row weigh alpha; zeros of the target are skipped).
 * cls_targets / cls_preds fp32 [T][2], loc_targets / loc_preds fp32 [T][L].  out4 = (total_pos, total_neg, loc, positive_num);
 * dcls [2][T][2] = the gradients of total_pos and of total_neg w.r.t. cls_preds; dloc [T][L] = gradient of loc w.r.t. loc_preds. */
int64_t hn_lane_loss_workspace_bytes(int32_t T);
int hn_lane_loss(const float* cls_targets, const float* cls_preds, const float* loc_targets, const float* loc_preds, int32_t T, int32_t L,
                 int32_t weighted_index, float negative_ratio, float alpha, void* workspace, int64_t workspace_bytes, float* out4, float* dcls,
                 float* dloc, void* stream);

/* Adam step over many tensors in one launch (torch.optim.Adam semantics, train.py:147: L2 weight decay added to the
 * gradient, bias-corrected moments).  Tensor table on the device. */
typedef struct hn_adam_tensor {
    float* p;
    const float* g;
    float* m;
    float* v;
    int64_t n;
} hn_adam_tensor;
int hn_adam_step(const hn_adam_tensor* tensors_device, const int32_t* chunk_tensor_device, const int32_t* chunk_index_device,
                 int32_t n_chunks, int32_t chunk_elems, float lr, float beta1, float beta2, float eps, float weight_decay,
                 int32_t step, float grad_scale, float* dyn_device, void* stream);
/* dyn_device (optional): fp32 [2] = {learning rate, number of steps taken so far}.  When given, `lr` / `step` are ignored:
 * the kernel reads them from the device and a second one-thread kernel advances the step count -- so a CUDA graph that
 * captured the call keeps stepping correctly, and a scheduler changes the rate with one small device write. */

/* diagnostics: when set (before a conv is prepared), every conv CTA stores 16 int64 globaltimer stamps */
void hn_conv_set_debug_buffer(void* device_i64);
void hn_se_set_split_fc(int on); /* tuning knob: squeeze-excite FC layers as two batched launches after the pooling, or (default: measured faster) fused into the pooling kernel's tail */
int hn_se_pool_num_launches(const hn_se_pool_desc* d);
void hn_conv_set_cluster(int ctas); /* tuning knob: CTAs per cluster sharing a weight tile (0 = default, 1 = off) */
void hn_conv_set_pair_min_bn(int bn); /* tuning knob: narrowest N tile run on CTA pairs (default 192) */
void hn_det_set_rounds_ctas_per_sm(int n); /* tuning knob: CTAs per SM of the cooperative NMS rounds kernel (default 2) */
void hn_plan_set_branch_priority(int on); /* tuning knob: side branches of a plan on high-priority streams (default on) */
void hn_det_set_rounds_passes(int n); /* tuning knob: passes over a short NMS worklist per grid barrier (default 3) */
void hn_set_pdl(int on); /* tuning knob: programmatic dependent launch of the plan's kernels (default off: measured slower in graph replay) */
void hn_conv_set_tap_runs(int mode); /* tuning knob: dy taps sharing one A box (0 = off, 1 = default policy, 2 = wherever they fit) */
void hn_det_set_debug_buffer(void* device_i64); /* [N*16][8] int64 cycle counters of the NMS kernel */
void hn_det_force_sequential(int on);           /* tests: run the sequential per-class NMS kernel only */
int hn_version(void);
const char* hn_last_error(void);
int hn_device_sm_count(void);

#ifdef __cplusplus
}
#endif
#endif /* HYDRANET_B200_H */
