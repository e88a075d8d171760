/*
 * centerpose_b200.h -- C ABI of libcenterpose_b200.so (sm_100a).
 *
 * The drop-in boundary for the CenterPose inference hot path
 * (SURVEY.md section 8b).  Plain pointers and sizes only; every function
 * returns 0 on success or a negative cp_status, and cp_last_error() returns a
 * thread-local message.  The library never allocates in steady state: the
 * caller (PyTorch on the Python side) owns every input / output buffer, a
 * plan owns its packed weights and activation arena.  All device work is
 * enqueued on the cudaStream_t passed as `stream` (pass
 * torch.cuda.current_stream().cuda_stream); no call synchronises the device
 * unless documented.
 *
 * Reference interfaces each entry point replaces (paths relative to
 * /root/reference/src/lib):
 *
 *   cp_plan_create / cp_plan_load_weights / cp_plan_destroy
 *       models/model.py:26-31 create_model(), :34-87 load_model()
 *       models/networks/pose_dla_dcn.py:457-521 DLASeg.__init__, :573-590 factories
 *   cp_forward
 *       models/networks/pose_dla_dcn.py:523-570 DLASeg.forward
 *       (DLA :310-322, DLAUp :437-443, IDAUp :411-417, DeformConv :377-389,
 *        convGRU.py:72-94, GN.py:4-9)
 *   cp_dcn_v2_forward
 *       models/networks/DCNv2/src/vision.cpp:4-9  _ext.dcn_v2_forward
 *       models/networks/DCNv2/src/dcn_v2.h:9-45, src/cuda/dcn_v2_cuda.cu:42-172,
 *       src/cuda/dcn_v2_im2col_cuda.cu:125-195
 *   cp_decode_pnp
 *       detectors/object_pose.py:131-165 process() [sigmoid + object_pose_decode]
 *       models/decode.py:72-375, utils/post_process.py:12-68,
 *       detectors/object_pose.py:184-197 merge_outputs + :27-124 soft_nms_nvidia,
 *       detectors/base_detector.py:548-654 point assembly + pnp_shell,
 *       utils/pnp/cuboid_pnp_shell.py:11-93, utils/pnp/cuboid_pnp_solver.py:91-239
 *   cp_infer
 *       detectors/base_detector.py:473-654 (process -> post_process -> merge -> PnP)
 *   cp_preprocess
 *       detectors/base_detector.py:91-148 pre_process (resize + affine warp + normalise)
 */
#ifndef CENTERPOSE_B200_H_
#define CENTERPOSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CP_ABI_VERSION 1

/* ---- status codes --------------------------------------------------------- */
enum cp_status {
  CP_OK = 0,
  CP_ERR_INVALID = -1,      /* bad argument / unsupported configuration        */
  CP_ERR_CUDA = -2,         /* a CUDA runtime call failed (message has details) */
  CP_ERR_NOT_LOADED = -3,   /* plan used before cp_plan_load_weights           */
  CP_ERR_MISSING_KEY = -4,  /* a state_dict key the plan needs was not supplied */
  CP_ERR_SHAPE = -5         /* tensor element count does not match the plan     */
};

/* ---- architecture / precision -------------------------------------------- */
enum cp_arch {
  CP_ARCH_DLA34 = 0,        /* 'dla_34'   : DLA-34 + DCNv2            (model.py:19) */
  CP_ARCH_DLAV1_34 = 1      /* 'dlav1_34' : + convGRU + GroupNorm heads (model.py:20) */
};

enum cp_precision {
  CP_PREC_FP32 = 0,         /* fp32 operands and accumulation on CUDA cores (parity mode)      */
  CP_PREC_TF32X3 = 1,       /* tcgen05 kind::tf32, 3-term split + promoted accumulation: fp32-equivalent
                             * tensor-core mode, meets the same parity bar as CP_PREC_FP32; the Python host's
                             * default                                                                      */
  CP_PREC_BF16 = 2,         /* tcgen05 kind::f16 bf16 operands, fp32 accumulation (fast mode)  */
  CP_PREC_TF32 = 3          /* tcgen05 kind::tf32 single pass -- the math PyTorch's cuDNN convolutions use by
                             * default (allow_tf32); stride-2 convs run the 3-term split gather kernel     */
};

#define CP_MAX_HEADS 16
#define CP_POSE_RECORD 192  /* floats per detection slot in `poses`  */
#define CP_DETS_RECORD 128  /* floats per candidate slot in `dets`   */
#define CP_META_DOUBLES 16  /* doubles per image in `meta`           */
#define CP_MAX_K 128
#define CP_MAX_CLASSES 80   /* hm channels the decode stage merges (decode.py:52-68 _topk) */

typedef struct cp_plan cp_plan;

typedef struct cp_config {
  int32_t arch;                 /* cp_arch                                                */
  int32_t tracking;             /* 1: pre_img / pre_hm / pre_hm_hp stems (pose_dla_dcn.py:253-271) */
  int32_t tracking_task_gru;    /* dlav1 only: 4 GRU steps + tracking routing (:473-477,545-555)   */
  int32_t max_batch;
  int32_t height, width;        /* network input, multiples of 32                         */
  int32_t precision;            /* cp_precision                                           */
  int32_t device;               /* CUDA ordinal                                           */
  int32_t head_conv;            /* 256                                                    */
  int32_t num_heads;
  const char* head_names[CP_MAX_HEADS];    /* in opt.heads order (opts.py:394-426)        */
  int32_t head_channels[CP_MAX_HEADS];
} cp_config;

/* Create a plan: builds the static layer schedule and allocates the
 * activation arena + packed-weight storage on cfg->device.  Synchronous. */
int cp_plan_create(const cp_config* cfg, cp_plan** out);
int cp_plan_destroy(cp_plan* plan);

/* Ingest a reference state_dict (Appendix A of SURVEY.md).  names[i] is the
 * reference key (an optional leading "module." is ignored, model.py:43-47),
 * dev_ptrs[i] a DEVICE pointer to contiguous fp32 data in the reference's
 * layout (OIHW conv weights, [C] vectors), numel[i] its element count.  The
 * pointers are only borrowed for the duration of the call: BatchNorm is
 * folded, weights are repacked (K-major, padded) into plan-owned storage.
 * Unknown keys (e.g. base.fc.*, num_batches_tracked) are ignored; a key the
 * plan needs but does not get returns CP_ERR_MISSING_KEY.  Enqueued on
 * `stream`; the borrowed tensors must stay alive until that work completes. */
int cp_plan_load_weights(cp_plan* plan, const char* const* names, const void* const* dev_ptrs,
                         const int64_t* numel, int32_t n, void* stream);

/* Network forward.  images: device fp32 NCHW [batch,3,H,W]; pre_img [batch,3,H,W],
 * pre_hm [batch,1,H,W], pre_hm_hp [batch,8,H,W] or NULL (tracking plans only).
 * head_out[i]: device fp32 NCHW [batch, head_channels[i], H/4, W/4] receiving the
 * LOGITS of head i (the reference applies sigmoid later, object_pose.py:136-138). */
int cp_forward(cp_plan* plan, int32_t batch, const float* images, const float* pre_img,
               const float* pre_hm, const float* pre_hm_hp, float* const* head_out, void* stream);

/* Per-op timing of one forward (the reference only has wall-clock stamps around whole stages,
 * base_detector.py:466-498): CUDA events are recorded on `stream` between the ops of the
 * schedule; the call synchronises on the last event.  `flops` / `bytes` are the ALGORITHMIC
 * work of the op (2*MAC; fp32 tensors touched once), used for the roofline in bench.py.
 * kind = op_type*10 + igemm_mode  (0: NCHW stem conv, 1: NHWC conv, 2: deformable conv,
 * 10: max-pool, 20: up-sample+add, 30: GroupNorm+ReLU, 40: GRU gate). */
typedef struct cp_op_stat {
  char name[96];
  int32_t kind;
  float ms;
  double flops;
  double bytes;
} cp_op_stat;
int cp_plan_num_ops(const cp_plan* plan);
int cp_plan_profile(cp_plan* plan, int32_t batch, const float* images, const float* pre_img,
                    const float* pre_hm, const float* pre_hm_hp, float* const* head_out, void* stream,
                    cp_op_stat* stats, int32_t max_stats, int32_t* n_stats);

/* Arena / weight bytes owned by the plan (for logging). */
int64_t cp_plan_bytes(const cp_plan* plan);
/* Number of kernel launches one cp_forward enqueues (for bench gpu_launches). */
int32_t cp_plan_forward_launches(const cp_plan* plan);

/* ---- decode + grouping + post-process + soft-NMS + PnP --------------------- */
typedef struct cp_heads {
  /* device fp32 NCHW [batch, C, out_h, out_w]; NULL when the head is absent */
  const float* hm;                /* [B,num_classes,..] logits (or probabilities if !apply_sigmoid) */
  const float* wh;                /* [B,2,..]   */
  const float* hps;               /* [B,2J,..]  */
  const float* reg;               /* [B,2,..]   or NULL */
  const float* hm_hp;             /* [B,J,..]   */
  const float* hp_offset;         /* [B,2,..]   or NULL */
  const float* scale;             /* [B,3,..]   or NULL */
  const float* hps_uncertainty;   /* [B,2J,..]  or NULL */
  const float* scale_uncertainty; /* [B,3,..]   or NULL */
  const float* tracking;          /* [B,2,..]   or NULL */
  const float* tracking_hp;       /* [B,2J,..]  or NULL */
} cp_heads;

typedef struct cp_decode_params {
  int32_t batch, out_h, out_w;
  int32_t num_classes;     /* opt.num_classes = hm channels, 1..CP_MAX_CLASSES (Objectron models: 1, opts.py:434): per-class
                            * top-K, then the K best of the num_classes x K candidates (decode.py:52-68)               */
  int32_t num_joints;      /* 8                                                            */
  int32_t K;               /* opt.K = 100, <= CP_MAX_K                                     */
  int32_t rep_mode;        /* opt.rep_mode: 0,1,3,4 (2 = random GMM sampling, unsupported) */
  int32_t use_moments;     /* opt.tracking_task || opt.refined_Kalman (decode.py:222)      */
  int32_t nms;             /* opt.nms (demo.py:114 sets True)                              */
  int32_t visible_thresh;  /* cuboid_pnp_shell.py:59-66: 6 book/chair/cereal_box, 3 camera/bottle/cup, 0 bike/laptop/shoe */
  int32_t opencv_return;   /* opt.show_axes: return the OpenCV pose instead of OpenGL      */
  int32_t apply_sigmoid;   /* 0: hm / hm_hp are probabilities; 1: both are logits; 2: only hm is a logit
                            * (opt.mse_loss: hm_hp is decoded raw, object_pose.py:136-138)  */
  int32_t use_pnp;         /* opt.use_pnp                                                  */
  float vis_thresh;        /* opt.vis_thresh (0.3)                                         */
  float balance;           /* opt.balance_coefficient[opt.c] (2)                           */
  int32_t modern_bool_semantics; /* 0 (default): the pinned torch==1.1.0 meaning of models/decode.py:183-188, where the
                            * seven comparison results are ADDED as integers and `mask_2 == 7` = "all seven gates hold".
                            * 1: what the unmodified reference computes on torch >= 1.2 (bool + bool is a logical OR, so
                            * `== 7` is never true): the heat-map keypoint representation is never used, kps_heatmap_* stay
                            * at the -10000 sentinel and the PnP sees the 8 displacement points only.                 */
  float test_scale;        /* opt.test_scales[0] (1; 0 is read as 1): the scale the frame was resized by in pre_process; != 1 divides bbox,
                            * kps, kps_displacement_mean / _std, kps_heatmap_mean, tracking and tracking_hp by it in
                            * float32 before soft-NMS and the PnP (object_pose.py:171-177)                              */
  int32_t num_scales;      /* len(opt.test_scales): > 1 forces the soft-NMS (object_pose.py:193); merge_outputs keeps the
                            * detections of test_scales[0] only (object_pose.py:188 reads detections[0])               */
} cp_decode_params;

/* meta: device fp64 [batch, CP_META_DOUBLES] per image:
 *   [0] c_x  [1] c_y  [2] s (src width of the affine, base_detector.py:114)
 *   [3] image width  [4] image height  [5..13] camera matrix row-major  [14,15] unused
 * dets:  device fp32 [batch, K, CP_DETS_RECORD] or NULL -- the 13 arrays of
 *        decode.py:348-361 in output-map pixels (layout: cp_dets_field).
 * poses: device fp32 [batch, K, CP_POSE_RECORD] -- slots [0, n_valid[b]) hold the
 *        reference's `results` (after score filter + soft-NMS) in order, image pixels,
 *        with the PnP output of each (layout: cp_pose_field).
 * n_valid: device int32 [batch].
 * workspace: device scratch of at least cp_decode_workspace_bytes(prm) bytes. */
size_t cp_decode_workspace_bytes(const cp_decode_params* prm);
int cp_decode_pnp(const cp_decode_params* prm, const cp_heads* heads, const double* meta,
                  float* dets, float* poses, int32_t* n_valid,
                  void* workspace, size_t workspace_bytes, void* stream);

/* forward + decode in one call; head maps stay in plan-owned buffers.
 * `heads_out` may be NULL, or an array of num_heads device pointers that
 * additionally receive the head logits (NCHW). */
int cp_infer(cp_plan* plan, int32_t batch, const float* images, const float* pre_img,
             const float* pre_hm, const float* pre_hm_hp, const cp_decode_params* prm,
             const double* meta, float* const* heads_out, float* dets, float* poses,
             int32_t* n_valid, void* stream);

/* Offsets (in floats) inside one CP_POSE_RECORD slot. */
enum cp_pose_field {
  CP_P_SCORE = 0, CP_P_CLS = 1, CP_P_STATUS = 2, CP_P_NPTS = 3,
  CP_P_BBOX = 4,            /* 4  */
  CP_P_CT = 8,              /* 2  */
  CP_P_KPS = 10,            /* 16 */
  CP_P_KPS_DISP_MEAN = 26,  /* 16 */
  CP_P_KPS_HM_MEAN = 42,    /* 16 */
  CP_P_KPS_HM_STD = 58,     /* 16 */
  CP_P_KPS_HM_HEIGHT = 74,  /* 8  */
  CP_P_KPS_DISP_STD = 82,   /* 16 */
  CP_P_OBJ_SCALE = 98,      /* 3  */
  CP_P_OBJ_SCALE_UNC = 101, /* 3  */
  CP_P_TRACKING = 104,      /* 2  */
  CP_P_TRACKING_HP = 106,   /* 16 */
  CP_P_LOCATION = 122,      /* 3  */
  CP_P_QUAT = 125,          /* 4 xyzw */
  CP_P_REPROJ = 129,        /* 1  */
  CP_P_PROJ_CUBOID = 130,   /* 16 */
  CP_P_KPS_3D_CAM = 146,    /* 27 */
  CP_P_KPS_PNP = 173,       /* 18 */
  CP_P_SRC_INDEX = 191      /* index k of the candidate in the top-K list */
};

/* PnP status stored in CP_P_STATUS (mirrors the reference's None paths). */
enum cp_pnp_status {
  CP_PNP_NOT_RUN = 0,
  CP_PNP_OK = 1,          /* pnp_shell returned a tuple -> goes into `boxes`                    */
  CP_PNP_INVISIBLE = 2,   /* pose stored in the result, visibility gate returned None (:59-79)  */
  CP_PNP_BEHIND = 3,      /* z < 0 (cuboid_pnp_solver.py:208-220)                               */
  CP_PNP_FEW_POINTS = 4,  /* < 4 valid points (cuboid_pnp_solver.py:157-160); 4-5 points are solved by EPnP   */
  CP_PNP_SOLVER_FAIL = 5
};

/* Offsets inside one CP_DETS_RECORD slot (output-map pixel units). */
enum cp_dets_field {
  CP_D_BBOX = 0, CP_D_SCORE = 4, CP_D_CLS = 5,
  CP_D_KPS = 6,             /* 16 */
  CP_D_OBJ_SCALE = 22,      /* 3  */
  CP_D_OBJ_SCALE_UNC = 25,  /* 3  */
  CP_D_TRACKING = 28,       /* 2  */
  CP_D_TRACKING_HP = 30,    /* 16 */
  CP_D_KPS_DISP_MEAN = 46,  /* 16 */
  CP_D_KPS_DISP_STD = 62,   /* 16 */
  CP_D_KPS_HM_MEAN = 78,    /* 16 */
  CP_D_KPS_HM_STD = 94,     /* 16 */
  CP_D_KPS_HM_HEIGHT = 110, /* 8  */
  CP_D_IND = 118            /* flat index of the centre cell */
};

/* ---- CenterPoseTrack state on the device (SURVEY.md rows a-T / f-2) ----------- */
/* Replaces utils/tracker.py:15-302 (Tracker: association, 32-state Kalman filter per object, scale pool, second PnP
 * with the filtered keypoints), the gaussian_fusion closure of detectors/base_detector.py:502-544 and the rendering of
 * the previous-frame heat maps, base_detector.py:150-388 (_get_additional_inputs, default options: use_pnp,
 * render_hm_mode 1, render_hmhp_mode 0-3).  One tracker holds `streams` independent videos (the batch dimension);
 * state lives in device memory, every call is enqueued on `stream`, calls on one tracker must be issued in frame order
 * on one stream.  Not supported: --hungarian, meta['pre_dets'] seeding, gt_pre_hm_hmhp. */
#define CP_TRACK_RECORD 320
typedef struct cp_tracker cp_tracker;
typedef struct cp_tracker_config {
  int32_t streams;            /* independent videos = max batch of cp_tracker_step                         */
  int32_t max_tracks;         /* per stream, <= CP_MAX_K                                                    */
  int32_t kalman;             /* opt.kalman                                                                 */
  int32_t scale_pool;         /* opt.scale_pool                                                             */
  int32_t use_pnp;            /* opt.use_pnp                                                                */
  int32_t hps_uncertainty;    /* opt.hps_uncertainty                                                        */
  int32_t max_age;            /* opt.max_age (5)                                                            */
  int32_t visible_thresh;     /* as in cp_decode_params                                                     */
  int32_t opencv_return;      /* opt.show_axes                                                              */
  int32_t render_hm_mode;     /* opt.render_hm_mode (1: centre heat = score)                                */
  int32_t render_hmhp_mode;   /* opt.render_hmhp_mode (2: PnP keypoints, heat = filter confidence)          */
  int32_t device;
  float new_thresh;           /* opt.new_thresh                                                             */
  float pre_thresh;           /* opt.pre_thresh                                                             */
  float R;                    /* opt.R (20)                                                                 */
  float conf_lo, conf_hi;     /* opt.conf_border[opt.c] (3, 9)                                              */
} cp_tracker_config;

int cp_tracker_create(const cp_tracker_config* cfg, cp_tracker** out);
int cp_tracker_destroy(cp_tracker* trk);
/* Tracker.reset(): forget every track of stream `index` (or of all streams when index < 0). */
int cp_tracker_reset(cp_tracker* trk, int32_t index, void* stream);
/* Tracker.step(results, boxes) for `batch` streams (stream b <- poses[b]).  poses / n_valid / meta as produced by
 * cp_decode_pnp / cp_infer (K slots per image).  tracks_out: device fp32 [batch, max_tracks, CP_TRACK_RECORD], slots
 * [0, n_tracks[b]) = the reference's self.tracker.tracks in order (layout: cp_track_field); n_tracks: device int32. */
int cp_tracker_step(cp_tracker* trk, int32_t batch, const float* poses, const int32_t* n_valid, int32_t K,
                    const double* meta, float* tracks_out, int32_t* n_tracks, void* stream);
/* _get_additional_inputs(): render the tracks of every stream into pre_hm [batch,1,inp_h,inp_w] and pre_hm_hp
 * [batch,8,inp_h,inp_w] (device fp32, overwritten).  trans_input: device fp64 [batch,6] = the row-major 2x3 affine
 * meta['trans_input'] (original image -> network input). */
int cp_tracker_render(cp_tracker* trk, int32_t batch, const double* meta, const double* trans_input, int32_t inp_h,
                      int32_t inp_w, float* pre_hm, float* pre_hm_hp, void* stream);

/* Offsets (in floats) inside one CP_TRACK_RECORD slot.  [0, CP_POSE_RECORD) is the pose record of the detection the
 * track carries; its PnP fields hold the SECOND (filtered) solve whenever that produced a pose (pnp_shell mutates the
 * track dict, cuboid_pnp_shell.py:27-54). */
enum cp_track_field {
  CP_T_ID = 192, CP_T_AGE = 193, CP_T_ACTIVE = 194,
  CP_T_IN_BOXES = 195,         /* 1: the second PnP returned a tuple and conf_avg > 0.25 -> in `boxes` (tracker.py:278-281) */
  CP_T_PNP2_STATUS = 196,      /* cp_pnp_status of the second PnP                                          */
  CP_T_CONF_AVG = 197,
  CP_T_KPS_FUSION_MEAN = 200,  /* 16 */
  CP_T_KPS_FUSION_STD = 216,   /* 16 */
  CP_T_KPS_MEAN_KF = 232,      /* 16, -10000 where the filter confidence is < 0.15                         */
  CP_T_KPS_STD_KF = 248,       /* 16 */
  CP_T_OBJ_SCALE_KF = 264,     /* 3  */
  CP_T_OBJ_SCALE_UNC_KF = 267, /* 3  */
  CP_T_KPS_PNP_KF = 270,       /* 18 */
  CP_T_KPS_3D_CAM_KF = 288     /* 27 */
};

/* ---- stand-alone modulated deformable convolution (the `_ext` replacement) -- */
/* input [B,C,H,W], weight [Co,C,3,3], bias [Co], offset [B,18,H,W]
 * (channel 2k = dy, 2k+1 = dx of tap k), mask [B,9,H,W] (already sigmoid'ed),
 * output [B,Co,H,W]; all device fp32 NCHW.  3x3, stride 1, pad 1, dilation 1,
 * deformable_group 1 -- the only configuration CenterPose instantiates
 * (pose_dla_dcn.py:384).  Scratch is taken from the stream-ordered allocator. */
int cp_dcn_v2_forward(const float* input, const float* weight, const float* bias,
                      const float* offset, const float* mask, float* output,
                      int32_t B, int32_t C, int32_t H, int32_t W, int32_t Co, void* stream);

/* Same op with an explicit cp_precision (CP_PREC_FP32: CUDA cores; CP_PREC_TF32X3 / CP_PREC_BF16: tcgen05). */
int cp_dcn_v2_forward_ex(const float* input, const float* weight, const float* bias, const float* offset,
                         const float* mask, float* output, int32_t B, int32_t C, int32_t H, int32_t W,
                         int32_t Co, int32_t precision, void* stream);

/* Backward of the same op: replaces `_ext.dcn_v2_backward` (DCNv2/dcn_v2.py:63-76 -> src/cuda/dcn_v2_cuda.cu:206-335,
 * kernels dcn_v2_im2col_cuda.cu:197-330).  Inputs as in the forward plus grad_output [B,Co,H,W]; the five gradients
 * (grad_input [B,C,H,W], grad_offset [B,18,H,W], grad_mask [B,9,H,W], grad_weight [Co,C,3,3], grad_bias [Co]) are
 * OVERWRITTEN (the reference returns fresh tensors).  `precision` selects the kernel family of the column-gradient GEMM
 * (CP_PREC_FP32: CUDA cores; CP_PREC_TF32X3: tcgen05, fp32-equivalent); the sampling pass and the weight-gradient GEMM
 * run in fp32.  grad_input is accumulated with float atomics (as in the reference), everything else in a fixed order.
 * Scratch (about B*H*W*(11*C + Co + 32) floats) comes from the stream-ordered allocator. */
int cp_dcn_v2_backward(const float* input, const float* weight, const float* offset, const float* mask,
                       const float* grad_output, float* grad_input, float* grad_offset, float* grad_mask,
                       float* grad_weight, float* grad_bias, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Co,
                       int32_t precision, void* stream);

/* ---- single fused convolution (building block of the plan, exposed for layer-level parity tests) ----
 * out = [relu]( conv2d(x, weight, stride, pad) + bias [+ residual] ); x / residual / out are device fp32
 * NHWC ([B,H,W,Cin] / [B,Ho,Wo,Cout]), weight is OIHW like nn.Conv2d (pose_dla_dcn.py:37-44);
 * Cin % 16 == 0, Cout % 4 == 0.  bias / residual may be NULL. */
int cp_conv2d(const float* x, const float* weight, const float* bias, const float* residual, float* out,
              int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t k, int32_t stride,
              int32_t pad, int32_t relu, int32_t precision, void* stream);

/* ---- batched pre-process (next-row f-1) ------------------------------------ */
/* frames: device uint8 [B, src_h, src_w, 3] (BGR as cv2.imread gives);
 * out: device fp32 NCHW [B,3,dst_h,dst_w] = (warpAffine(frame)/255 - mean)/std with the
 * reference's fix_res affine (c = src centre, s = max(src_h, src_w)).  The warp is a bit-for-bit restatement of
 * cv2.warpAffine(INTER_LINEAR) for 8-bit frames (OpenCV's fixed-point remap: AB_SCALE 1024, INTER_BITS 5,
 * 15-bit integer weights), so the batched path feeds the network exactly what base_detector.py:128-134 does. */
int cp_preprocess(const uint8_t* frames, float* out, int32_t B, int32_t src_h, int32_t src_w,
                  int32_t dst_h, int32_t dst_w, const float mean[3], const float std[3], void* stream);
/* Same with an explicit forward affine: trans_input = HOST pointer to the row-major 2x3 double matrix
 * meta['trans_input'] (source frame -> network input), e.g. for fix_short / keep_res or rotated crops. */
int cp_preprocess_affine(const uint8_t* frames, float* out, int32_t B, int32_t src_h, int32_t src_w, int32_t dst_h,
                         int32_t dst_w, const double trans_input[6], const float mean[3], const float std[3], void* stream);

/* ---- misc ------------------------------------------------------------------ */
int cp_version(void);
const char* cp_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* CENTERPOSE_B200_H_ */
