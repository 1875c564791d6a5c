// Internal structures shared by the plan builder and the kernel launchers of libyolob200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/yolob200.h"

namespace yb {

void set_error(const char* fmt, ...);
void count_launch();

#define YB_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      yb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                    __LINE__);                                                         \
      return YB_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

// Programmatic dependent launch: every kernel of the forward / NMS chain is launched with the
// programmatic-stream-serialization attribute, triggers its dependents at once and executes
// griddepcontrol.wait before its first global-memory access.  A dependent grid therefore starts as
// soon as all CTAs of its predecessor are running (or done) and overlaps its prologue (barrier
// init, TMEM allocation, descriptor prefetch) with the predecessor's tail; data is only touched
// after the predecessor has completed and flushed.  Every CTA of every kernel executes the wait, so
// completion of kernel N implies completion of kernel N-1 (buffer-reuse hazards stay ordered).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_prologue_done() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

#ifdef __CUDACC__
// The two 16-bit activation / weight storage types (plan->act_f16).  Arithmetic is fp32 everywhere; only
// the pack / unpack at loads and stores and the tensor-core operand format differ.
template <bool F16>
struct Act16 {
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    if (F16) {
      __half2 h = __floats2half2_rn(a, b);
      return *reinterpret_cast<uint32_t*>(&h);
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack2(uint32_t v) {
    if (F16) return __half22float2(*reinterpret_cast<const __half2*>(&v));
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
  }
  static __device__ __forceinline__ uint16_t pack1(float a) {
    if (F16) return __half_as_ushort(__float2half_rn(a));
    return __bfloat16_as_ushort(__float2bfloat16(a));
  }
  static __device__ __forceinline__ float unpack1(uint16_t v) {
    if (F16) return __half2float(__ushort_as_half(v));
    return __uint_as_float((uint32_t)v << 16);
  }
  static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
    if (F16) {
      __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
      return *reinterpret_cast<uint32_t*>(&r);
    }
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  // mma.sync.m16n8k16, fp32 accumulate
  static __device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    if (F16)
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
          : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
          : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
          : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
          : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
#endif
// host-side conversion of one fp32 value to the plan's 16-bit storage type
static inline uint16_t host_to_act16(float v, bool f16) {
  if (f16) return __half_as_ushort(__float2half_rn(v));
  return __bfloat16_as_ushort(__float2bfloat16(v));
}
static inline float host_from_act16(uint16_t v, bool f16) {
  if (f16) return __half2float(__ushort_as_half(v));
  return __bfloat162float(__ushort_as_bfloat16(v));
}
static constexpr int YB_MAX_DEVICES = 64;   // per-device caches of function attributes (power of two)
typedef uint16_t act_t;   // one activation / weight element in global memory (fp16 or bf16 bits)

// Selects a device for the duration of a yb_* call and restores the caller's current device on exit
// (torch tracks the current device itself; a library call must not change it behind its back).
struct DeviceGuard {
  int prev = -1, rc = YB_OK;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) {
      cudaError_t e = cudaSetDevice(dev);
      if (e != cudaSuccess) {
        yb::set_error("cudaSetDevice(%d) failed: %s", dev, cudaGetErrorString(e));
        rc = YB_ERR_CUDA;
        prev = -1;
      }
    } else {
      prev = -1;
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int cpad8(int c) { return round_up(c, 8); }

// ---------------------------------------------------------------------------------------------
// Activation buffers: NHWC bf16 (or fp32 for the head logits) inside one workspace arena.
// A "slice" is a channel range [c_off, c_off + C) of a buffer; producers write straight into the
// slice of their consumer's concat buffer, so torch.cat / chunk never exist as kernels.
// ---------------------------------------------------------------------------------------------
static constexpr int YB_MAX_LANES = 4;

struct Buf {
  int H = 0, W = 0, C = 0;  // C = total (padded) channels = row stride in elements
  int elem_bytes = 2;
  int rows_per_img = 0;     // H*W, or A for the logits buffer
  size_t offset = 0;        // byte offset in the workspace arena
  size_t bytes = 0;
  int first_def = 1 << 30, last_use = -1;  // liveness (op indices) for arena reuse
  std::vector<int> touches;                // every op that reads or writes the buffer (stream-lane aware reuse)
  // space-to-depth storage (plan.cu): the (H, W, C) tensor lives as (H/2, W/2, 4C), channel block (y&1)*2 + (x&1) -
  // written that way by its producer for a single stride-2 3x3 consumer, which then reads TMA halo patches
  int s2d = 0;
  std::string tag;
};

struct Slice {
  int buf = -1;
  int c_off = 0;
  int C = 0;      // real channels
  int up = 0;     // 1: consumer reads this source nearest-upsampled x2 (nn.py:195,205-206)
};

enum OpKind { OP_STEM = 0, OP_CONV = 1, OP_DW = 2, OP_POOL = 3, OP_ATTN = 4, OP_DECODE = 5 };

struct Op {
  OpKind kind = OP_CONV;
  std::string name;
  // geometry
  int k = 1, stride = 1;
  int Hin = 0, Win = 0;    // input spatial dims at the conv's own resolution (after upsample)
  int Hout = 0, Wout = 0;
  int nseg = 0;
  Slice src[4];
  Slice dst;               // dst.C = real cout
  int dst_row_off = 0;     // row offset inside an image of the dst buffer (anchor offset of a level)
  int has_res = 0;
  Slice res;
  int act = 1;
  int out_f32 = 0;         // epilogue writes fp32 (head logits)
  int conv_index = -1;     // index into Plan::convs
  // tensor-core GEMM parameters (OP_CONV)
  int a_tma = 0;           // A operand via TMA tiled loads (1x1, no upsample), else software im2col
  int K = 0, K_pad = 0;    // GEMM K (real, padded to 64)
  int seg_kpad[4] = {0, 0, 0, 0};  // per-segment padded K (a_tma mode)
  int N_pad = 0, BN = 0;   // padded Cout, N tile
  int stages = 0;
  int tile2d = 0;          // 8 x 16 spatial output tiles (else 128 consecutive flattened rows)
  int patch = 0;           // 3x3/s1: A operand read from TMA halo patches through shifted descriptors
  int patch_stage_bytes = 0, patch_stages = 0;
  int pair = 0;            // patch layers: 16 x 16 pixel super-tiles (two accumulators per weight k-block)
  int b_resident = 0;      // weight matrix stays in shared memory across the CTA's tiles
  int c_bufs = 1;          // output staging buffers
  int dw_fused = 0;        // OP_CONV: the depthwise conv `dw_op` in front of this 1x1 is computed by its producer warps
  int dw_op = -1;
  int fused_away = 0;      // OP_DW: computed inside the consumer's kernel, not launched in a forward
  int occ = 1;             // resident CTAs per SM of the persistent GEMM kernel
  int head_part = 0;       // 1: box tail (fusable DFL decode), 2: cls tail (fusable sigmoid)
  size_t smem_bytes = 0;
  CUtensorMap tmap_b;
  CUtensorMap tmap_a[4];
  CUtensorMap tmap_c;      // bf16 output slice, written by the epilogue's TMA store
  // depthwise: channel gather of the source (attention `pe` conv reads the v rows of qkv)
  int dw_gsz = 0, dw_gstride = 0, dw_goff = 0, dw_add = 0;
  // attention
  int heads = 0, dk = 0, dh = 0;
  float scale = 0.f;
  // residual folded into the consumer's weights (plan.cu, C3k2): input channels [wfold_dst, +wfold_n) of this conv
  // take the sum of their own weights and those of channels [wfold_src, +wfold_n) when the blob is packed
  int wfold_dst = 0, wfold_src = 0, wfold_n = 0;
  int s2d = 0;             // stride-2 3x3 conv whose source buffer is stored space-to-depth (halo-patch path): 1 = 16, 2 = 64 channels
  // stream lanes (plan.cu): independent branches of the graph are enqueued on separate streams
  int lane = 0;            // 0 = the caller's stream
  int signal = 0;          // an op on another lane waits for this one: record an event behind it
  std::vector<int> xdeps;  // ops on other lanes this op waits for
};

struct ConvW {
  yb_conv_info info;
  int op = -1;
};

struct GraphEntry {
  const void* in = nullptr;
  void* out = nullptr;
  void* stream = nullptr;
  int dtype = 0, raw = 0, impl = 0;
  const void* sink = nullptr;   // NMS candidate sink captured in the graph (yb_forward_nms)
  float sink_conf = 0.f;
  const void* sink_hdr = nullptr;   // the sink's header block and per-image key stride are baked into the graph too
  int sink_cap = 0;
  cudaGraphExec_t exec = nullptr;
};

}  // namespace yb

struct yb_plan {
  yb_arch_desc arch;
  int B = 0, H = 0, W = 0, device = 0;
  int act_f16 = 1;         // activation / weight storage type: 1 = IEEE fp16 (default), 0 = bf16
  int nc = 0, no = 0, A = 0;
  int lvl_h[3], lvl_w[3], lvl_off[3];
  float lvl_stride[3];
  std::vector<yb::Buf> bufs;
  std::vector<yb::Op> ops;
  std::vector<yb::ConvW> convs;
  size_t workspace_bytes = 0, weight_bytes = 0;
  int logits_buf = -1;
  const uint8_t* d_weights = nullptr;
  uint8_t* d_ws = nullptr;
  bool bound = false;
  int conv_impl = 0;
  int use_graph = 0;
  int fuse_decode = 1;     // head tails decode in their epilogue; the logits buffer is skipped
  int num_sms = 148;
  // candidate sink of the current forward (yb_forward_nms): NMS key lists filled by the class-score epilogues
  int* sink_hdr = nullptr;
  unsigned long long* sink_keys = nullptr;
  int sink_cap = 0;
  float sink_conf = 0.f;
  std::vector<yb::GraphEntry> graphs;
  cudaStream_t capture_stream = nullptr;
  int profiling = 0;
  std::vector<std::vector<cudaEvent_t>> prof_events;  // one vector of ops+1 events per recorded forward
  int prof_used = 0;
  // stream lanes: side streams + the events of cross-lane dependencies (created on first use)
  int num_lanes = 1;
  cudaStream_t lane_streams[yb::YB_MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t lane_join[yb::YB_MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> op_events;
};

namespace yb {
// plan.cu
int build_plan(yb_plan* p);
// launchers (each enqueues exactly one kernel on `st`)
int launch_conv_tc(const yb_plan* p, const Op& op, cudaStream_t st, float* fused_out);
int launch_conv_naive(const yb_plan* p, const Op& op, cudaStream_t st);
int launch_stem(const yb_plan* p, const Op& op, const void* in, int in_dtype, cudaStream_t st);
int launch_dw(const yb_plan* p, const Op& op, cudaStream_t st);
int launch_pool(const yb_plan* p, const Op& op, cudaStream_t st);
int launch_attn(const yb_plan* p, const Op& op, cudaStream_t st);
int launch_decode(const yb_plan* p, const float* logits, float* out, cudaStream_t st);
int conv_tc_prepare(yb_plan* p, Op& op);  // tensor maps + smem attribute, needs bound buffers

inline uint8_t* buf_ptr(const yb_plan* p, int buf) { return p->d_ws + p->bufs[buf].offset; }
// element offset of logical pixel row r = y * W + x of image n inside a buffer (space-to-depth aware)
inline size_t buf_row_elem(const Buf& b, size_t n, size_t r) {
  if (!b.s2d) return (n * (size_t)b.rows_per_img + r) * (size_t)b.C;
  const size_t y = r / (size_t)b.W, x = r % (size_t)b.W;
  return ((n * (size_t)(b.H / 2) + y / 2) * (size_t)(b.W / 2) + x / 2) * (size_t)(4 * b.C) + ((y & 1) * 2 + (x & 1)) * (size_t)b.C;
}
}  // namespace yb
