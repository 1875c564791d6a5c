/*
 * yolob200.h — C ABI of libyolob200.so, the B200 (sm_100a) implementation of the
 * YOLO-Infer-pt inference hot path.
 *
 * The reference (t0saki/YOLO-Infer-pt) has no FFI layer of its own: its boundary is the Python
 * surface `nets.nn.YOLO.forward` (nets/nn.py:294-297) and `utils.util.non_max_suppression`
 * (utils/util.py:123-169).  Every entry point below names the reference interface it stands in
 * for.  The Python host side (yolo_infer_pt_b200/nets/nn.py, utils/util.py) binds these symbols
 * with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain C types only; no torch / pybind types cross this boundary;
 *   - every device buffer (input, output, weights, workspace) is owned by the caller;
 *     the library owns plan structs, TMA descriptors and CUDA-graph handles only;
 *   - all work is ordered on the caller's stream (small-batch plans run independent branches of the
 *     network on side streams the plan owns, forked from and joined to the caller's stream by events
 *     inside the call); no hidden synchronisation (except yb_plan_bind, which may run one-off set-up
 *     kernels and is not on the hot path);
 *   - return value: 0 = ok, < 0 = error; the message is in yb_last_error() (thread-local);
 *   - the library never falls back to the CPU: if no sm_100 device is present the calls fail.
 */
#ifndef YOLOB200_H
#define YOLOB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YB_OK 0
#define YB_ERR_ARG (-1)
#define YB_ERR_CUDA (-2)
#define YB_ERR_UNSUPPORTED (-3)
#define YB_ERR_STATE (-4)

/* element types of the NCHW input image tensor handed to yb_forward */
#define YB_F32 0
#define YB_F16 1
#define YB_BF16 2
#define YB_U8 3 /* uint8 0..255; the kernel divides by 255 (main.py:265-267 does it on the host side) */

typedef struct yb_plan yb_plan;

/* Architecture of one YOLOv11 variant: the three lists the reference's constructors pass to
 * YOLO(width, depth, csp, num_classes) (nets/nn.py:308-347). */
typedef struct yb_arch_desc {
  int width[6];
  int depth[6];
  int csp[2];
  int num_classes;
} yb_arch_desc;

/* One convolution of the network, in execution order, as the weight packer needs it.
 * `name` is the state_dict prefix of the module in the reference's naming
 * (e.g. "net.p2.1.conv1" for a Conv wrapper, "head.box.0.2" for a plain Conv2d tail). */
typedef struct yb_conv_info {
  char name[96];
  int cout;          /* real output channels                                         */
  int cin;           /* real input channels per group (weight.shape[1])              */
  int ksize;         /* 1 or 3                                                        */
  int stride;        /* 1 or 2                                                        */
  int groups;        /* 1 (dense) or cout (depthwise)                                 */
  int act;           /* 1 = SiLU, 0 = identity                                        */
  int wrapped;       /* 1 = reference `Conv` wrapper (conv+norm), 0 = bare Conv2d     */
  int kind;          /* 0 stem, 1 dense tensor-core GEMM, 2 depthwise                 */
  size_t blob_offset;/* byte offset of this conv's packed weights in the weight blob */
  size_t blob_bytes;
} yb_conv_info;

/* ---- plan life-cycle: stands in for YOLO.__init__/fuse()/forward wiring, nets/nn.py:282-305 ---- */

/* Build the static execution plan (op list, activation-buffer aliasing, TMA descriptors) for one
 * (architecture, batch, height, width) on `device`. H and W must be multiples of 32
 * (nets/nn.py:205-206 concatenates stride-2 pyramids).
 * act_dtype: storage type of activations and packed weights, YB_F16 or YB_BF16 (accumulation, bias,
 * SiLU, DFL decode and sigmoid are always fp32).  YB_F16 is what the reference's own evaluation uses
 * (main.py:251 `model.half()`, main.py:266 `samples.half()`) and is the default of the Python host side;
 * YB_BF16 trades 3 mantissa bits for range.  device < 0: host-only plan (layout / packer, never bound). */
int yb_plan_create(const yb_arch_desc* arch, int batch, int height, int width, int act_dtype, int device,
                   yb_plan** out);
void yb_plan_destroy(yb_plan* plan);

size_t yb_plan_workspace_bytes(const yb_plan* plan); /* activation arena, device memory  */
size_t yb_plan_weight_bytes(const yb_plan* plan);    /* packed weight blob               */
int yb_plan_num_anchors(const yb_plan* plan);        /* A = sum over levels of H_i*W_i   */
int yb_plan_num_outputs(const yb_plan* plan);        /* 4 + num_classes                  */
int yb_plan_num_convs(const yb_plan* plan);
int yb_plan_conv_info(const yb_plan* plan, int index, yb_conv_info* out);
int yb_plan_num_launches(const yb_plan* plan);       /* kernels one yb_forward enqueues  */

/* Host-side weight packer: takes the BN-folded fp32 OIHW weight and bias of conv `index`
 * (what fuse_conv produces, nets/nn.py:8-25) and writes the kernel layout into host_blob
 * (a host buffer of yb_plan_weight_bytes bytes). */
int yb_plan_pack_conv(const yb_plan* plan, int index, const float* weight_oihw, const float* bias,
                      void* host_blob);

/* Attach device buffers: the packed weight blob (already copied to the device by the caller) and
 * the workspace arena. Must be called before yb_forward; may be called again after a re-pack. */
int yb_plan_bind(yb_plan* plan, const void* dev_weights, void* dev_workspace);

/* ---- forward: stands in for YOLO.forward in eval mode, nets/nn.py:294-297 + Head 255-270 ---- */

/* in_nchw : (B,3,H,W) image tensor of dtype in_dtype, values in [0,1] (or 0..255 for YB_U8)
 * out     : (B, 4+nc, A) fp32 — rows [cx,cy,w,h,score_0..score_nc-1], pixel units,
 *           anchors ordered level 8,16,32, row-major (nets/nn.py:262-270). */
int yb_forward(yb_plan* plan, const void* in_nchw, int in_dtype, float* out, void* cuda_stream);

/* yb_forward whose class-score epilogues also do non_max_suppression's candidate filter (utils/util.py:130,
 * 147: scores > conf): every such (anchor, class) is appended to the per-image key lists of
 * `nms_workspace` (yb_nms_workspace_bytes bytes, headers zero: yb_nms_workspace_init, or last used by a
 * yb_nms* call of the same batch) while the scores are still in registers, instead of re-reading the
 * (B, nc, A) scores in the NMS.  Follow with yb_nms_prefiltered on the same workspace, conf and max_nms. */
int yb_forward_nms(yb_plan* plan, const void* in_nchw, int in_dtype, float* out, float conf, int max_nms,
                   void* nms_workspace, size_t nms_workspace_bytes, void* cuda_stream);

/* Same, but stops before the DFL decode and returns the pre-decode logits:
 * raw : (B, A, 64+nc) fp32, channel order [box 4x16, cls nc] — the training-mode head output
 * of nets/nn.py:256-259 flattened and level-concatenated (a transposed view of nn.py:262). */
int yb_forward_raw(yb_plan* plan, const void* in_nchw, int in_dtype, float* raw, void* cuda_stream);

/* Capture the forward into a CUDA graph keyed on (in, out, stream) and replay it on later calls
 * with the same pointers. enable = 0 returns to plain stream launches. */
int yb_plan_use_graph(yb_plan* plan, int enable);

/* Per-op device timing for bench.py's roofline: while enabled (and graphs are off), every forward
 * brackets each op with CUDA events on the caller's stream. yb_plan_profile_read synchronises,
 * writes the mean milliseconds of each op over the forwards recorded since the last read into
 * op_ms[0 .. num_launches) and returns how many forwards were averaged. */
int yb_plan_profile(yb_plan* plan, int enable);
int yb_plan_profile_read(yb_plan* plan, float* op_ms, int capacity);

/* Debug / validation switch: 0 = tcgen05 tensor-core convolutions (the product path),
 * 1 = scalar direct-convolution CUDA kernel used only to cross-check the tensor-core kernel. */
int yb_plan_set_conv_impl(yb_plan* plan, int impl);

/* Copy an intermediate activation (by producing conv name) to host fp32 NHWC for layer-level
 * parity tests. Synchronises the stream. Returns the number of floats written or < 0. */
long long yb_plan_debug_read(yb_plan* plan, const char* conv_name, float* host_out,
                             size_t host_capacity_floats, int* out_h, int* out_w, int* out_c);

/* Teacher-forced single-op execution for the per-op parity tests: overwrite activation buffer
 * `buf_index` (as listed by yb_plan_describe) from host memory in the buffer's own element type,
 * then run op `op_index` alone. `in_nchw` is only read by the stem op. Both synchronise. */
int yb_plan_debug_write(yb_plan* plan, int buf_index, const void* host_data, size_t bytes);
int yb_plan_run_op(yb_plan* plan, int op_index, const void* in_nchw, int in_dtype, float* out,
                   void* cuda_stream);

/* Writes a JSON description of the plan (buffers, ops, slices, GEMM shapes) into buf; returns the
 * number of bytes needed (call with capacity 0 to size the buffer). Used by the host-logic tests
 * to replay the dataflow on the CPU. */
long long yb_plan_describe(const yb_plan* plan, char* buf, size_t capacity);

/* ---- NMS: stands in for utils.util.non_max_suppression, utils/util.py:123-169 ---- */

size_t yb_nms_workspace_bytes(int batch, int num_classes, int num_anchors, int max_nms);

/* pred       : (B, 4+nc, A) fp32 device tensor, boxes as (cx,cy,w,h)
 * out        : (B, max_det, 6) fp32 rows [x1,y1,x2,y2,score,class]; rows >= counts[b] untouched
 * out_counts : (B) int32 detections kept per image
 * conf       : candidates are scores strictly greater than conf (fp32 compare, util.py:130,147)
 * iou        : suppress when (double)IoU > iou (strict, util.py:162 → torchvision CPU nms)
 * max_wh     : per-class coordinate offset (util.py:124,160), max_nms: sort cap (util.py:126,157),
 * max_det    : kept-box cap (util.py:125,163). */
int yb_nms(const float* pred, int batch, int num_classes, int num_anchors, float conf, double iou,
           int max_det, int max_nms, float max_wh, float* out, int* out_counts, void* workspace,
           size_t workspace_bytes, void* cuda_stream);
/* Same call without the memset node in front of the first kernel, for a workspace whose headers are
 * known to be zero: one that yb_nms_workspace_init cleared, or that the previous yb_nms / yb_nms_clean
 * call used with the same (batch, max_nms) - the per-image kernel re-zeroes its header on exit.  Keeps the
 * forward -> NMS kernel chain free of non-kernel nodes (programmatic dependent launch, batch-1 latency). */
int yb_nms_workspace_init(void* workspace, size_t workspace_bytes, void* cuda_stream);
int yb_nms_clean(const float* pred, int batch, int num_classes, int num_anchors, float conf, double iou,
                 int max_det, int max_nms, float max_wh, float* out, int* out_counts, void* workspace,
                 size_t workspace_bytes, void* cuda_stream);

/* yb_nms for predictions produced by yb_forward_nms with this workspace: the candidate lists are already
 * filled, the pass over the scores is skipped (images that overflowed their list are still rebuilt from
 * pred).  Results are identical to yb_nms. */
int yb_nms_prefiltered(const float* pred, int batch, int num_classes, int num_anchors, float conf, double iou,
                       int max_det, int max_nms, float max_wh, float* out, int* out_counts, void* workspace,
                       size_t workspace_bytes, void* cuda_stream);

/* ---- misc ---- */
/* ---- pre-processing (the step in front of YOLO.forward) ------------------------
 * Replaces Dataset.load_image's cv2.resize(INTER_LINEAR) (utils/dataset.py:95-103), resize()'s
 * letterbox with a constant-0 border (utils/dataset.py:292-313) and the HWC->CHW / BGR->RGB shuffle
 * (utils/dataset.py:86-88) for a whole batch in one kernel, bit-exact with OpenCV's 8-bit bilinear.
 * desc : device array [batch][3] of int64 = (device pointer of an HWC uint8 BGR image, height, width)
 * out  : device (batch, 3, input_size, input_size) uint8 RGB, the tensor yb_forward takes as YB_U8
 * meta : optional device [batch][3] double = (ratio, pad_w, pad_h) as resize() returns them        */
int yb_letterbox(const long long* desc, int batch, int input_size, uint8_t* out_nchw_rgb, double* meta,
                 void* cuda_stream);

/* ---- detection consumer (the step behind non_max_suppression) -------------------
 * Replaces compute_metric (utils/util.py:99-120) for the whole padded batch yb_nms produced: IoU of every
 * detection with the class-matched labels and the greedy label assignment per IoU threshold, on the
 * device (the reference round-trips every image through numpy).
 * det (batch, max_det, 6) + counts (batch): yb_nms's outputs;  targets (batch, max_targets, 5) rows
 * [class, x1, y1, x2, y2] in pixels + target_counts (batch);  iou_v (n_iou <= 16) thresholds;
 * correct (batch, max_det, n_iou) uint8: the reference's boolean matrix, zero past counts[b].          */
int yb_compute_metric(const float* det, const int* counts, const float* targets, const int* target_counts,
                      int batch, int max_det, int max_targets, const float* iou_v, int n_iou, uint8_t* correct,
                      void* cuda_stream);

const char* yb_last_error(void);
unsigned long long yb_launch_count(void); /* kernels launched by this library so far (process-wide) */
int yb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* YOLOB200_H */
