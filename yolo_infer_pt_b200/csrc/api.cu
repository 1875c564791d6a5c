// extern "C" surface of libyolob200.so (see include/yolob200.h).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <new>

#include "yb_internal.h"

namespace yb {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("YB_NO_PDL") ? 0 : 1;
  return v != 0;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

size_t nms_workspace_bytes(int B, int nc, int A, int max_nms);
int nms_run(const float* pred, int B, int nc, int A, float conf, double iou, int max_det, int max_nms,
            float max_wh, float* out, int* out_counts, void* ws, size_t ws_bytes, cudaStream_t st, int ws_clean);
int nms_sink_layout(void* ws, size_t ws_bytes, int B, int nc, int A, int max_nms, int** hdr,
                    unsigned long long** keys, int* cap);

int metric_run(const float* det, const int* counts, const float* tgt, const int* tcounts, int B, int max_det,
               int max_t, const float* iou_v, int n_iou, uint8_t* correct, cudaStream_t st);
int letterbox_run(const long long* desc, int B, int S, uint8_t* out, double* meta, cudaStream_t st);

static int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error("no CUDA device available (%s); libyolob200 has no CPU path",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return YB_ERR_CUDA;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (%d devices)", device, n);
    return YB_ERR_ARG;
  }
  cudaDeviceProp prop;
  YB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major,
              prop.minor);
    return YB_ERR_UNSUPPORTED;
  }
  return YB_OK;
}

static int run_one(yb_plan* p, const Op& op, const void* in, int in_dtype, float* out, int raw,
                   cudaStream_t st) {
  switch (op.kind) {
    case OP_STEM:
      return launch_stem(p, op, in, in_dtype, st);
    case OP_CONV: {
      if (p->conv_impl == 1) return launch_conv_naive(p, op, st);
      bool fuse = p->fuse_decode && !raw && op.head_part;
      return launch_conv_tc(p, op, st, fuse ? out : nullptr);
    }
    case OP_DW:
      return launch_dw(p, op, st);
    case OP_POOL:
      return launch_pool(p, op, st);
    case OP_ATTN:
      return launch_attn(p, op, st);
    case OP_DECODE: {
      const Buf& lb = p->bufs[p->logits_buf];
      const float* logits = reinterpret_cast<const float*>(buf_ptr(p, p->logits_buf));
      if (raw) {
        size_t rows = (size_t)p->B * p->A;
        YB_CUDA(cudaMemcpy2DAsync(out, (size_t)p->no * 4, logits, (size_t)lb.C * 4, (size_t)p->no * 4, rows,
                                  cudaMemcpyDeviceToDevice, st));
        return YB_OK;
      }
      if (p->fuse_decode && p->conv_impl == 0) return YB_OK;  // decoded in the head tails' epilogues
      return launch_decode(p, logits, out, st);
    }
  }
  return YB_OK;
}

static int run_ops(yb_plan* p, const void* in, int in_dtype, float* out, int raw, cudaStream_t st) {
  int rc = YB_OK;
  std::vector<cudaEvent_t>* ev = nullptr;
  if (p->profiling && !p->use_graph) {
    if (p->prof_used == (int)p->prof_events.size()) {
      if (p->prof_events.size() < 256) {
        std::vector<cudaEvent_t> v(p->ops.size() + 1);
        for (auto& e : v) YB_CUDA(cudaEventCreate(&e));
        p->prof_events.push_back(v);
      }
    }
    if (p->prof_used < (int)p->prof_events.size()) ev = &p->prof_events[p->prof_used++];
  }
  size_t op_i = 0;
  if (ev) YB_CUDA(cudaEventRecord((*ev)[0], st));
  // Stream lanes (plan.cu): independent branches run on the plan's side streams; cross-lane dependencies are
  // events, and every side lane joins the caller's stream again before this function returns.  Per-op profiling
  // and the scalar cross-check path keep everything on the caller's stream.
  const bool lanes = p->num_lanes > 1 && !ev && p->conv_impl == 0;
  bool lane_used[YB_MAX_LANES] = {false, false, false, false};
  if (lanes) {
    if (p->op_events.size() != p->ops.size()) p->op_events.assign(p->ops.size(), nullptr);
    for (int l = 1; l < p->num_lanes; l++) {
      if (!p->lane_streams[l]) YB_CUDA(cudaStreamCreateWithFlags(&p->lane_streams[l], cudaStreamNonBlocking));
      if (!p->lane_join[l]) YB_CUDA(cudaEventCreateWithFlags(&p->lane_join[l], cudaEventDisableTiming));
    }
  }
  for (const Op& op : p->ops) {
    cudaStream_t os = st;
    if (lanes) {
      if (op.lane > 0) os = p->lane_streams[op.lane];
      for (int d : op.xdeps) YB_CUDA(cudaStreamWaitEvent(os, p->op_events[d], 0));
      lane_used[op.lane] = true;
    }
    // a depthwise conv fused into its consumer's kernel is only launched for the scalar cross-check path
    if (!(op.fused_away && p->conv_impl == 0)) rc = run_one(p, op, in, in_dtype, out, raw, os);
    if (rc) break;
    if (lanes && op.signal) {
      if (!p->op_events[op_i]) YB_CUDA(cudaEventCreateWithFlags(&p->op_events[op_i], cudaEventDisableTiming));
      YB_CUDA(cudaEventRecord(p->op_events[op_i], os));
    }
    op_i++;
    if (ev) YB_CUDA(cudaEventRecord((*ev)[op_i], st));
  }
  if (lanes)
    for (int l = 1; l < p->num_lanes; l++)
      if (lane_used[l]) {   // (also on the error path: a capture must not end with a forked stream)
        YB_CUDA(cudaEventRecord(p->lane_join[l], p->lane_streams[l]));
        YB_CUDA(cudaStreamWaitEvent(st, p->lane_join[l], 0));
      }
  return rc;
}

static int forward_impl(yb_plan* p, const void* in, int in_dtype, float* out, int raw, void* stream) {
  if (!p || !in || !out) {
    set_error("yb_forward: null argument");
    return YB_ERR_ARG;
  }
  if (!p->bound) {
    set_error("yb_forward: plan has no weights/workspace bound (call yb_plan_bind)");
    return YB_ERR_STATE;
  }
  DeviceGuard guard(p->device);
  if (guard.rc) return guard.rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!p->use_graph) return run_ops(p, in, in_dtype, out, raw, st);
  for (GraphEntry& g : p->graphs) {
    if (g.in == in && g.out == out && g.stream == stream && g.dtype == in_dtype && g.raw == raw &&
        g.impl == p->conv_impl && g.sink == (const void*)p->sink_keys && g.sink_conf == p->sink_conf &&
        g.sink_hdr == (const void*)p->sink_hdr && g.sink_cap == p->sink_cap) {
      YB_CUDA(cudaGraphLaunch(g.exec, st));
      g_launches.fetch_add((unsigned long long)yb_plan_num_launches(p), std::memory_order_relaxed);
      return YB_OK;
    }
  }
  cudaGraph_t graph = nullptr;
  unsigned long long before = g_launches.load();
  // The caller's stream may be the legacy default stream, which cannot be captured: record the op
  // list on a private stream and replay the instantiated graph on the caller's stream.
  if (!p->capture_stream)
    YB_CUDA(cudaStreamCreateWithFlags(&p->capture_stream, cudaStreamNonBlocking));
  YB_CUDA(cudaStreamBeginCapture(p->capture_stream, cudaStreamCaptureModeThreadLocal));
  int rc = run_ops(p, in, in_dtype, out, raw, p->capture_stream);
  cudaError_t e = cudaStreamEndCapture(p->capture_stream, &graph);
  g_launches.store(before);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) {
    set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    return YB_ERR_CUDA;
  }
  GraphEntry g;
  g.in = in;
  g.out = out;
  g.stream = stream;
  g.dtype = in_dtype;
  g.raw = raw;
  g.impl = p->conv_impl;
  g.sink = p->sink_keys;
  g.sink_conf = p->sink_conf;
  g.sink_hdr = p->sink_hdr;
  g.sink_cap = p->sink_cap;
  e = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    return YB_ERR_CUDA;
  }
  if (p->graphs.size() >= 8) {
    cudaGraphExecDestroy(p->graphs.front().exec);
    p->graphs.erase(p->graphs.begin());
  }
  p->graphs.push_back(g);
  YB_CUDA(cudaGraphLaunch(g.exec, st));
  g_launches.fetch_add((unsigned long long)yb_plan_num_launches(p), std::memory_order_relaxed);
  return YB_OK;
}

static void drop_graphs(yb_plan* p) {
  for (GraphEntry& g : p->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  p->graphs.clear();
}

}  // namespace yb

using namespace yb;

extern "C" {

int yb_plan_create(const yb_arch_desc* arch, int batch, int height, int width, int act_dtype, int device,
                   yb_plan** out) {
  if (!arch || !out || batch <= 0) {
    set_error("yb_plan_create: bad argument");
    return YB_ERR_ARG;
  }
  if (act_dtype != YB_F16 && act_dtype != YB_BF16) {
    set_error("yb_plan_create: act_dtype must be YB_F16 or YB_BF16 (got %d)", act_dtype);
    return YB_ERR_ARG;
  }
  *out = nullptr;
  int rc = YB_OK;
  if (device >= 0) {  // device < 0: host-only plan (op list, weight layout); it can never be bound
    rc = check_device(device);
    if (rc) return rc;
  }
  yb_plan* p = new (std::nothrow) yb_plan();
  if (!p) {
    set_error("out of host memory");
    return YB_ERR_ARG;
  }
  p->arch = *arch;
  p->B = batch;
  p->H = height;
  p->W = width;
  p->device = device;
  p->act_f16 = act_dtype == YB_F16 ? 1 : 0;
  cudaDeviceProp prop;
  if (device >= 0 && cudaGetDeviceProperties(&prop, device) == cudaSuccess)
    p->num_sms = prop.multiProcessorCount;
  rc = build_plan(p);
  if (rc) {
    delete p;
    return rc;
  }
  *out = p;
  return YB_OK;
}

void yb_plan_destroy(yb_plan* plan) {
  if (!plan) return;
  drop_graphs(plan);
  for (auto& v : plan->prof_events)
    for (auto& e : v) cudaEventDestroy(e);
  if (plan->capture_stream) cudaStreamDestroy(plan->capture_stream);
  for (cudaEvent_t e : plan->op_events)
    if (e) cudaEventDestroy(e);
  for (int l = 0; l < YB_MAX_LANES; l++) {
    if (plan->lane_join[l]) cudaEventDestroy(plan->lane_join[l]);
    if (plan->lane_streams[l]) cudaStreamDestroy(plan->lane_streams[l]);
  }
  delete plan;
}

size_t yb_plan_workspace_bytes(const yb_plan* plan) { return plan ? plan->workspace_bytes : 0; }
size_t yb_plan_weight_bytes(const yb_plan* plan) { return plan ? plan->weight_bytes : 0; }
int yb_plan_num_anchors(const yb_plan* plan) { return plan ? plan->A : 0; }
int yb_plan_num_outputs(const yb_plan* plan) { return plan ? 4 + plan->nc : 0; }
int yb_plan_num_convs(const yb_plan* plan) { return plan ? (int)plan->convs.size() : 0; }
int yb_plan_num_launches(const yb_plan* plan) {
  if (!plan) return 0;
  int n = (int)plan->ops.size() - (plan->fuse_decode && plan->conv_impl == 0 ? 1 : 0);
  if (plan->conv_impl == 0)
    for (const Op& op : plan->ops) n -= op.fused_away;   // depthwise convs computed inside their consumer's kernel
  return n;
}

int yb_plan_conv_info(const yb_plan* plan, int index, yb_conv_info* out) {
  if (!plan || !out || index < 0 || index >= (int)plan->convs.size()) {
    set_error("yb_plan_conv_info: bad argument");
    return YB_ERR_ARG;
  }
  *out = plan->convs[index].info;
  return YB_OK;
}

int yb_plan_pack_conv(const yb_plan* plan, int index, const float* w, const float* bias,
                      void* host_blob) {
  if (!plan || !w || !host_blob || index < 0 || index >= (int)plan->convs.size()) {
    set_error("yb_plan_pack_conv: bad argument");
    return YB_ERR_ARG;
  }
  const ConvW& cw = plan->convs[index];
  const Op& op = plan->ops[cw.op];
  uint8_t* dst = reinterpret_cast<uint8_t*>(host_blob) + cw.info.blob_offset;
  memset(dst, 0, cw.info.blob_bytes);
  const int cout = cw.info.cout, cin = cw.info.cin, k = cw.info.ksize;
  if (cw.info.kind == 1) {
    act_t* W = reinterpret_cast<act_t*>(dst);
    float* B = reinterpret_cast<float*>(dst + (size_t)op.N_pad * op.K_pad * 2);
    const bool f16 = plan->act_f16 != 0;
    int per_tap = 0;
    for (int s = 0; s < op.nseg; s++) per_tap += op.seg_kpad[s];   // (> padded channels for tap-aligned 3x3 layers)
    for (int co = 0; co < cout; co++) {
      act_t* row = W + (size_t)co * op.K_pad;
      for (int tap = 0; tap < k * k; tap++) {
        int kpos = op.a_tma ? 0 : tap * per_tap;
        int ci0 = 0;
        for (int s = 0; s < op.nseg; s++) {
          for (int c = 0; c < op.src[s].C; c++) {
            const int ci = ci0 + c;
            float v = w[((size_t)co * cin + ci) * k * k + tap];
            if (ci >= op.wfold_dst && ci < op.wfold_dst + op.wfold_n)   // residual folded into this conv (plan.cu, C3k2)
              v += w[((size_t)co * cin + (ci - op.wfold_dst + op.wfold_src)) * k * k + tap];
            row[kpos + c] = host_to_act16(v, f16);
          }
          ci0 += op.src[s].C;
          kpos += op.seg_kpad[s];
        }
      }
      B[co] = bias ? bias[co] : 0.f;
    }
  } else if (cw.info.kind == 0) {
    int Cp = cpad8(cout);
    float* W = reinterpret_cast<float*>(dst);
    for (int co = 0; co < cout; co++) {
      for (int t = 0; t < 27; t++) W[(size_t)t * Cp + co] = w[(size_t)co * 27 + t];
      W[(size_t)27 * Cp + co] = bias ? bias[co] : 0.f;
    }
  } else {
    int Cp = cpad8(cout);
    float* W = reinterpret_cast<float*>(dst);
    for (int co = 0; co < cout; co++) {
      for (int t = 0; t < 9; t++) W[(size_t)t * Cp + co] = w[(size_t)co * 9 + t];
      W[(size_t)9 * Cp + co] = bias ? bias[co] : 0.f;
    }
  }
  return YB_OK;
}

int yb_plan_bind(yb_plan* plan, const void* dev_weights, void* dev_workspace) {
  if (!plan || !dev_weights || !dev_workspace) {
    set_error("yb_plan_bind: null argument");
    return YB_ERR_ARG;
  }
  if (plan->device < 0) {
    set_error("yb_plan_bind: host-only plan (created with device < 0) cannot run");
    return YB_ERR_STATE;
  }
  DeviceGuard guard(plan->device);
  if (guard.rc) return guard.rc;
  drop_graphs(plan);
  plan->d_weights = reinterpret_cast<const uint8_t*>(dev_weights);
  plan->d_ws = reinterpret_cast<uint8_t*>(dev_workspace);
  for (Op& op : plan->ops) {
    if (op.kind == OP_CONV) {
      int rc = conv_tc_prepare(plan, op);
      if (rc) return rc;
    }
  }
  plan->bound = true;
  return YB_OK;
}

int yb_forward(yb_plan* plan, const void* in_nchw, int in_dtype, float* out, void* cuda_stream) {
  if (plan) plan->sink_keys = nullptr, plan->sink_hdr = nullptr;
  return forward_impl(plan, in_nchw, in_dtype, out, 0, cuda_stream);
}

int yb_forward_nms(yb_plan* plan, const void* in_nchw, int in_dtype, float* out, float conf, int max_nms,
                   void* nms_workspace, size_t nms_workspace_bytes, void* cuda_stream) {
  if (!plan || !nms_workspace) {
    set_error("yb_forward_nms: null argument");
    return YB_ERR_ARG;
  }
  if (!plan->fuse_decode || plan->conv_impl != 0) {
    set_error("yb_forward_nms: needs the fused head epilogues (tensor-core path, decode fusion on)");
    return YB_ERR_STATE;
  }
  int rc = nms_sink_layout(nms_workspace, nms_workspace_bytes, plan->B, plan->nc, plan->A, max_nms, &plan->sink_hdr,
                           &plan->sink_keys, &plan->sink_cap);
  if (rc) return rc;
  plan->sink_conf = conf;
  rc = forward_impl(plan, in_nchw, in_dtype, out, 0, cuda_stream);
  plan->sink_keys = nullptr;
  plan->sink_hdr = nullptr;
  return rc;
}

int yb_forward_raw(yb_plan* plan, const void* in_nchw, int in_dtype, float* raw, void* cuda_stream) {
  return forward_impl(plan, in_nchw, in_dtype, raw, 1, cuda_stream);
}

int yb_plan_use_graph(yb_plan* plan, int enable) {
  if (!plan) return YB_ERR_ARG;
  plan->use_graph = enable ? 1 : 0;
  if (!enable) drop_graphs(plan);
  return YB_OK;
}

int yb_plan_profile(yb_plan* plan, int enable) {
  if (!plan) return YB_ERR_ARG;
  plan->profiling = enable ? 1 : 0;
  plan->prof_used = 0;
  return YB_OK;
}

int yb_plan_profile_read(yb_plan* plan, float* op_ms, int capacity) {
  if (!plan || !op_ms || capacity < (int)plan->ops.size()) {
    set_error("yb_plan_profile_read: bad argument");
    return YB_ERR_ARG;
  }
  DeviceGuard guard(plan->device);
  if (guard.rc) return guard.rc;
  YB_CUDA(cudaDeviceSynchronize());
  int n = plan->prof_used;
  for (size_t i = 0; i < plan->ops.size(); i++) op_ms[i] = 0.f;
  for (int f = 0; f < n; f++) {
    for (size_t i = 0; i < plan->ops.size(); i++) {
      float ms = 0.f;
      YB_CUDA(cudaEventElapsedTime(&ms, plan->prof_events[f][i], plan->prof_events[f][i + 1]));
      op_ms[i] += ms / (float)n;
    }
  }
  plan->prof_used = 0;
  return n;
}

int yb_plan_set_conv_impl(yb_plan* plan, int impl) {
  if (!plan || impl < 0 || impl > 1) {
    set_error("yb_plan_set_conv_impl: bad argument");
    return YB_ERR_ARG;
  }
  plan->conv_impl = impl;
  return YB_OK;
}

long long yb_plan_debug_read(yb_plan* plan, const char* conv_name, float* host_out,
                             size_t host_capacity_floats, int* out_h, int* out_w, int* out_c) {
  if (!plan || !conv_name || !host_out || !plan->bound) {
    set_error("yb_plan_debug_read: bad argument or unbound plan");
    return YB_ERR_ARG;
  }
  for (const Op& op : plan->ops) {
    if (op.name != conv_name) continue;
    if (op.kind == OP_DECODE) break;
    const Buf& b = plan->bufs[op.dst.buf];
    int C = op.dst.C;
    if (op.kind == OP_POOL) C = op.dst.C;
    size_t rows = (size_t)plan->B * b.rows_per_img;
    size_t n = rows * C;
    if (n > host_capacity_floats) {
      set_error("yb_plan_debug_read: host buffer too small (%zu floats needed)", n);
      return YB_ERR_ARG;
    }
    DeviceGuard guard(plan->device);
  if (guard.rc) return guard.rc;
    YB_CUDA(cudaDeviceSynchronize());
    size_t row_bytes = (size_t)b.C * b.elem_bytes;
    std::vector<uint8_t> tmp(rows * row_bytes);
    YB_CUDA(cudaMemcpy(tmp.data(), buf_ptr(plan, op.dst.buf), tmp.size(), cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < rows; r++) {
      const uint8_t* rp = tmp.data() + (b.s2d ? buf_row_elem(b, r / b.rows_per_img, r % b.rows_per_img) * b.elem_bytes : r * row_bytes) +
                          (size_t)op.dst.c_off * b.elem_bytes;
      for (int c = 0; c < C; c++) {
        float v;
        if (b.elem_bytes == 4) {
          v = reinterpret_cast<const float*>(rp)[c];
        } else {
          v = host_from_act16(reinterpret_cast<const uint16_t*>(rp)[c], plan->act_f16 != 0);
        }
        host_out[r * C + c] = v;
      }
    }
    if (out_h) *out_h = b.H;
    if (out_w) *out_w = b.rows_per_img / (b.H ? b.H : 1);
    if (out_c) *out_c = C;
    return (long long)n;
  }
  set_error("yb_plan_debug_read: no op named %s", conv_name);
  return YB_ERR_ARG;
}

int yb_plan_debug_write(yb_plan* plan, int buf_index, const void* host_data, size_t bytes) {
  if (!plan || !plan->bound || !host_data || buf_index < 0 || buf_index >= (int)plan->bufs.size() ||
      bytes != plan->bufs[buf_index].bytes) {
    set_error("yb_plan_debug_write: bad argument (buffer %d, %zu bytes)", buf_index, bytes);
    return YB_ERR_ARG;
  }
  DeviceGuard guard(plan->device);
  if (guard.rc) return guard.rc;
  YB_CUDA(cudaDeviceSynchronize());
  const Buf& b = plan->bufs[buf_index];
  if (b.s2d) {   // host data is (B, H, W, C); the buffer is stored space-to-depth
    std::vector<uint8_t> tmp(bytes);
    const size_t px = (size_t)b.C * b.elem_bytes;
    for (size_t n = 0; n < (size_t)plan->B; n++)
      for (size_t r = 0; r < (size_t)b.rows_per_img; r++)
        memcpy(tmp.data() + buf_row_elem(b, n, r) * b.elem_bytes,
               reinterpret_cast<const uint8_t*>(host_data) + (n * b.rows_per_img + r) * px, px);
    YB_CUDA(cudaMemcpy(buf_ptr(plan, buf_index), tmp.data(), bytes, cudaMemcpyHostToDevice));
    return YB_OK;
  }
  YB_CUDA(cudaMemcpy(buf_ptr(plan, buf_index), host_data, bytes, cudaMemcpyHostToDevice));
  return YB_OK;
}

int yb_plan_run_op(yb_plan* plan, int op_index, const void* in_nchw, int in_dtype, float* out,
                   void* cuda_stream) {
  if (!plan || !plan->bound || op_index < 0 || op_index >= (int)plan->ops.size()) {
    set_error("yb_plan_run_op: bad argument");
    return YB_ERR_ARG;
  }
  DeviceGuard guard(plan->device);
  if (guard.rc) return guard.rc;
  int saved = plan->fuse_decode;
  plan->fuse_decode = 0;  // head tails write their logits slice, so the test can read it back
  int rc = run_one(plan, plan->ops[op_index], in_nchw, in_dtype, out, 0, (cudaStream_t)cuda_stream);
  plan->fuse_decode = saved;
  if (rc) return rc;
  YB_CUDA(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  return YB_OK;
}

long long yb_plan_describe(const yb_plan* plan, char* buf, size_t capacity) {
  if (!plan) return YB_ERR_ARG;
  std::string j;
  char t[512];
  auto slice = [&](const Slice& s) {
    snprintf(t, sizeof(t), "{\"buf\":%d,\"c_off\":%d,\"C\":%d,\"up\":%d}", s.buf, s.c_off, s.C, s.up);
    return std::string(t);
  };
  snprintf(t, sizeof(t), "{\"act_f16\":%d,\"num_lanes\":%d,", plan->act_f16, plan->num_lanes);
  j += t;
  snprintf(t, sizeof(t), "\"B\":%d,\"H\":%d,\"W\":%d,\"nc\":%d,\"A\":%d,\"logits_buf\":%d,"
           "\"workspace_bytes\":%zu,\"weight_bytes\":%zu,\"lvl_off\":[%d,%d,%d],\"bufs\":[",
           plan->B, plan->H, plan->W, plan->nc, plan->A, plan->logits_buf, plan->workspace_bytes,
           plan->weight_bytes, plan->lvl_off[0], plan->lvl_off[1], plan->lvl_off[2]);
  j += t;
  for (size_t i = 0; i < plan->bufs.size(); i++) {
    const Buf& b = plan->bufs[i];
    snprintf(t, sizeof(t), "%s{\"H\":%d,\"W\":%d,\"C\":%d,\"elem_bytes\":%d,\"rows_per_img\":%d,"
             "\"offset\":%zu,\"bytes\":%zu,\"first_def\":%d,\"last_use\":%d,\"s2d\":%d,\"tag\":\"%s\"}",
             i ? "," : "", b.H, b.W, b.C, b.elem_bytes, b.rows_per_img, b.offset, b.bytes, b.first_def,
             b.last_use, b.s2d, b.tag.c_str());
    j += t;
  }
  j += "],\"ops\":[";
  for (size_t i = 0; i < plan->ops.size(); i++) {
    const Op& o = plan->ops[i];
    snprintf(t, sizeof(t), "%s{\"kind\":%d,\"name\":\"%s\",\"k\":%d,\"stride\":%d,\"Hin\":%d,\"Win\":%d,"
             "\"Hout\":%d,\"Wout\":%d,\"act\":%d,\"out_f32\":%d,\"dst_row_off\":%d,\"conv_index\":%d,"
             "\"a_tma\":%d,\"patch\":%d,\"pair\":%d,\"resident\":%d,\"occ\":%d,\"stages\":%d,\"dw_fused\":%d,\"fused_away\":%d,\"K\":%d,\"K_pad\":%d,\"N_pad\":%d,\"BN\":%d,\"has_res\":%d,",
             i ? "," : "", (int)o.kind, o.name.c_str(), o.k, o.stride, o.Hin, o.Win, o.Hout, o.Wout, o.act,
             o.out_f32, o.dst_row_off, o.conv_index, o.a_tma, o.patch, o.pair, o.b_resident, o.occ, o.stages, o.dw_fused, o.fused_away, o.K, o.K_pad, o.N_pad, o.BN, o.has_res);
    j += t;
    snprintf(t, sizeof(t), "\"seg_kpad\":[%d,%d,%d,%d],\"dw\":[%d,%d,%d,%d],\"heads\":%d,\"scale\":%.9g,",
             o.seg_kpad[0], o.seg_kpad[1], o.seg_kpad[2], o.seg_kpad[3], o.dw_gsz, o.dw_gstride, o.dw_goff,
             o.dw_add, o.heads, o.scale);
    j += t;
    j += "\"src\":[";
    for (int s = 0; s < o.nseg; s++) j += (s ? "," : "") + slice(o.src[s]);
    j += "],\"s2d\":" + std::to_string(o.s2d) + ",\"wfold\":[" + std::to_string(o.wfold_dst) + "," + std::to_string(o.wfold_src) + "," + std::to_string(o.wfold_n) + "]";
    j += ",\"lane\":" + std::to_string(o.lane) + ",\"signal\":" + std::to_string(o.signal) + ",\"xdeps\":[";
    for (size_t s = 0; s < o.xdeps.size(); s++) j += (s ? "," : "") + std::to_string(o.xdeps[s]);
    j += "],\"dst\":" + slice(o.dst) + ",\"res\":" + slice(o.res) + "}";
  }
  j += "]}";
  if (buf && capacity > j.size()) memcpy(buf, j.c_str(), j.size() + 1);
  return (long long)j.size() + 1;
}

// device that owns a device pointer (the NMS / metric / letterbox entry points take no plan)
static int device_of(const void* ptr) {
  cudaPointerAttributes at;
  if (ptr && cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice) return at.device;
  cudaGetLastError();
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

size_t yb_nms_workspace_bytes(int batch, int num_classes, int num_anchors, int max_nms) {
  return nms_workspace_bytes(batch, num_classes, num_anchors, max_nms);
}

int yb_nms(const float* pred, int batch, int num_classes, int num_anchors, float conf, double iou,
           int max_det, int max_nms, float max_wh, float* out, int* out_counts, void* workspace,
           size_t workspace_bytes, void* cuda_stream) {
  if (!pred || !out || !out_counts) {
    set_error("yb_nms: null argument");
    return YB_ERR_ARG;
  }
  DeviceGuard guard(device_of(pred));
  if (guard.rc) return guard.rc;
  return nms_run(pred, batch, num_classes, num_anchors, conf, iou, max_det, max_nms, max_wh, out,
                 out_counts, workspace, workspace_bytes, (cudaStream_t)cuda_stream, 0);
}

int yb_nms_prefiltered(const float* pred, int batch, int num_classes, int num_anchors, float conf, double iou,
                       int max_det, int max_nms, float max_wh, float* out, int* out_counts, void* workspace,
                       size_t workspace_bytes, void* cuda_stream) {
  if (!pred || !out || !out_counts) {
    set_error("yb_nms_prefiltered: null argument");
    return YB_ERR_ARG;
  }
  DeviceGuard guard(device_of(pred));
  if (guard.rc) return guard.rc;
  return nms_run(pred, batch, num_classes, num_anchors, conf, iou, max_det, max_nms, max_wh, out,
                 out_counts, workspace, workspace_bytes, (cudaStream_t)cuda_stream, 2);
}

int yb_nms_workspace_init(void* workspace, size_t workspace_bytes, void* cuda_stream) {
  if (!workspace) {
    set_error("yb_nms_workspace_init: null workspace");
    return YB_ERR_ARG;
  }
  DeviceGuard guard(device_of(workspace));
  if (guard.rc) return guard.rc;
  YB_CUDA(cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)cuda_stream));
  return YB_OK;
}

int yb_nms_clean(const float* pred, int batch, int num_classes, int num_anchors, float conf, double iou,
                 int max_det, int max_nms, float max_wh, float* out, int* out_counts, void* workspace,
                 size_t workspace_bytes, void* cuda_stream) {
  if (!pred || !out || !out_counts) {
    set_error("yb_nms: null argument");
    return YB_ERR_ARG;
  }
  DeviceGuard guard(device_of(pred));
  if (guard.rc) return guard.rc;
  return nms_run(pred, batch, num_classes, num_anchors, conf, iou, max_det, max_nms, max_wh, out,
                 out_counts, workspace, workspace_bytes, (cudaStream_t)cuda_stream, 1);
}

int yb_letterbox(const long long* desc, int batch, int input_size, uint8_t* out_nchw_rgb, double* meta,
                 void* cuda_stream) {
  DeviceGuard guard(device_of(out_nchw_rgb));
  if (guard.rc) return guard.rc;
  return letterbox_run(desc, batch, input_size, out_nchw_rgb, meta, (cudaStream_t)cuda_stream);
}

int yb_compute_metric(const float* det, const int* counts, const float* targets, const int* target_counts,
                      int batch, int max_det, int max_targets, const float* iou_v, int n_iou, uint8_t* correct,
                      void* cuda_stream) {
  DeviceGuard guard(device_of(det));
  if (guard.rc) return guard.rc;
  return metric_run(det, counts, targets, target_counts, batch, max_det, max_targets, iou_v, n_iou, correct,
                    (cudaStream_t)cuda_stream);
}

const char* yb_last_error(void) { return g_err; }
unsigned long long yb_launch_count(void) { return g_launches.load(); }
int yb_version(void) { return 200; }

}  // extern "C"
