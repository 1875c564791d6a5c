// non_max_suppression on the device (reference utils/util.py:123-169, torchvision.ops.nms CPU
// semantics for the greedy step).  No host synchronisation, no data-dependent launch shapes:
//
//   1. append    one coalesced pass over (B, 4+nc, A): every score > conf becomes a 64-bit key
//                  key = ~orderable(score) << 32 | (anchor * nc + class)
//                ascending key order == descending score, ties by ascending candidate index
//                (anchor-major, class-minor: the row-major order of util.py:147's nonzero()).
//                Keys are appended to a per-image list while it has room.
//   2. select    only for images with more than max_nms candidates (util.py:157's [:max_nms]):
//                an exact radix select of the max_nms-th smallest key (6 histogram passes over
//                the scores, gated per image), then a re-compaction of keys <= that threshold.
//   3. sort      per-image bitonic sort of <= max_nms keys (shared-memory tiles of 4096 keys,
//                global compare-exchange steps above that).
//   4. greedy    one CTA per image walks the sorted candidates in chunks of 256: each candidate
//                is tested against the boxes kept so far (<= max_det of them), survivors are
//                resolved inside the chunk with a ballot-built IoU bitmask and a serial scan.
//                Greedy NMS only ever needs the first max_det kept boxes (util.py:163), so the
//                walk stops there and the n x n mask of torchvision's CUDA kernel is never built.
//
// IoU arithmetic replicates torchvision's CPU kernel operation by operation in fp32 with
// round-to-nearest intrinsics (no FMA contraction), on the class-offset boxes of util.py:160-161,
// and compares (double)iou > iou_threshold.
#include <stdio.h>
#include <string.h>

#include "yb_internal.h"

namespace yb {

static constexpr int HIST_BINS = 2048;
static constexpr int SORT_TILE = 4096;
static constexpr int NUM_PASSES = 6;  // 11,11,11,11,11,9 bits

struct NmsHeader {  // per image, zeroed at the start of every call
  int cand_count;   // candidates found by the scan
  int sel_count;    // keys appended by the scan (may exceed capacity; clamp on read)
  int sel2_count;   // keys appended by the re-compaction
  int overflow;     // cand_count > max_nms
  int need;         // remaining rank inside the current prefix
  int n_final;      // keys to sort / walk
  int pad[2];
  unsigned long long prefix;  // radix-select prefix; after the last pass: the threshold key
  unsigned long long pad2;
};

struct NmsArgs {
  const float* pred;
  int B, nc, A;
  float conf;
  double iou;
  int max_det, max_nms, cap;
  float max_wh;
  NmsHeader* hdr;
  unsigned int* hist;        // [B][HIST_BINS]
  unsigned long long* keys;  // [B][cap]
  float* out;
  int* out_counts;
};

__device__ __forceinline__ unsigned int orderable(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
__device__ __forceinline__ unsigned long long make_key(float score, unsigned int idx) {
  return ((unsigned long long)(~orderable(score)) << 32) | idx;
}
__device__ __forceinline__ int pass_shift(int pass) { return pass < 5 ? 64 - 11 * (pass + 1) : 0; }
__device__ __forceinline__ int pass_bits(int pass) { return pass < 5 ? 11 : 9; }

// Pass over the scores of every image: each score > conf becomes a key appended to the image's
// list (unordered; the sort orders them).  Reads 4 anchors per thread (float4 when aligned), one
// atomic per warp iteration.
__global__ void __launch_bounds__(256) nms_append_kernel(const NmsArgs a, int vec4) {
  const int b = blockIdx.y;
  NmsHeader* h = a.hdr + b;
  const long long total = (long long)a.nc * a.A;
  const float* sp = a.pred + ((size_t)b * (4 + a.nc) + 4) * a.A;
  unsigned long long* keys = a.keys + (size_t)b * a.cap;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  const long long iters = (total + stride - 1) / stride;
  long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  int local_count = 0;
  for (long long it = 0; it < iters; it++, e += stride) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    int nv = 0;
    if (e < total) {
      nv = (int)min((long long)4, total - e);
      if (vec4 && nv == 4) {
        float4 q = __ldg(reinterpret_cast<const float4*>(sp + e));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
        for (int j = 0; j < nv; j++) v[j] = __ldg(sp + e + j);
      }
    }
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) cnt += (j < nv && v[j] > a.conf) ? 1 : 0;
    // warp-wide exclusive prefix of cnt
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int wtotal = __shfl_sync(0xffffffffu, incl, 31);
    if (wtotal == 0) continue;
    int base = 0;
    if (lane == 0) base = atomicAdd(&h->sel_count, wtotal);
    base = __shfl_sync(0xffffffffu, base, 0);
    int slot = base + incl - cnt;
    local_count += cnt;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (j < nv && v[j] > a.conf) {
        long long ee = e + j;
        int c = (int)(ee / a.A);
        int an = (int)(ee - (long long)c * a.A);
        if (slot < a.cap) keys[slot] = make_key(v[j], (unsigned int)an * (unsigned int)a.nc + (unsigned int)c);
        slot++;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) local_count += __shfl_xor_sync(0xffffffffu, local_count, o);
  if (lane == 0 && local_count) atomicAdd(&h->cand_count, local_count);
}

// Overflow images only (more than max_nms candidates):
// mode 1: histogram of digit `pass` among keys matching the radix-select prefix
// mode 2: re-compaction of keys <= threshold
template <int MODE>
__global__ void __launch_bounds__(256) nms_scan_kernel(const NmsArgs a, int pass) {
  __shared__ unsigned int hist_s[HIST_BINS];
  const int b = blockIdx.y;
  NmsHeader* h = a.hdr + b;
  if (!h->overflow) return;
  if (MODE == 1) {
    for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x) hist_s[i] = 0;
    __syncthreads();
  }
  const long long total = (long long)a.nc * a.A;
  const float* sp = a.pred + ((size_t)b * (4 + a.nc) + 4) * a.A;
  unsigned long long* keys = a.keys + (size_t)b * a.cap;
  const unsigned long long prefix = h->prefix;
  const int shift = pass_shift(pass), bits = pass_bits(pass);
  const int lane = threadIdx.x & 31;
  // all lanes of a warp run the same number of iterations (ballots below)
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long iters = (total + stride - 1) / stride;
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long it = 0; it < iters; it++, e += stride) {
    bool cand = false;
    unsigned long long key = 0;
    if (e < total) {
      float s = __ldg(sp + e);
      if (s > a.conf) {
        int c = (int)(e / a.A);
        int an = (int)(e - (long long)c * a.A);
        key = make_key(s, (unsigned int)an * (unsigned int)a.nc + (unsigned int)c);
        cand = true;
      }
    }
    if (MODE == 1) {
      bool match = pass == 0 ? true : ((key >> (shift + bits)) == prefix);
      if (cand && match) atomicAdd(&hist_s[(unsigned int)(key >> shift) & ((1u << bits) - 1u)], 1u);
    } else {
      bool take = cand && key <= prefix;
      unsigned int m = __ballot_sync(0xffffffffu, take);
      if (m) {
        int leader = __ffs(m) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&h->sel2_count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (take) {
          int slot = base + __popc(m & ((1u << lane) - 1u));
          if (slot < a.cap) keys[slot] = key;
        }
      }
    }
  }
  if (MODE == 1) {
    __syncthreads();
    unsigned int* hg = a.hist + (size_t)b * HIST_BINS;
    for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x)
      if (hist_s[i]) atomicAdd(&hg[i], hist_s[i]);
  }
}

// One block per image: consume the histogram of digit `pass`, extend the prefix.
__global__ void __launch_bounds__(1024) nms_pick_kernel(const NmsArgs a, int pass) {
  __shared__ unsigned int cum[HIST_BINS];
  const int b = blockIdx.x;
  NmsHeader* h = a.hdr + b;
  unsigned int* hg = a.hist + (size_t)b * HIST_BINS;
  if (pass < 0) {  // decision step: does this image exceed max_nms candidates?
    if (threadIdx.x == 0) {
      int total = h->cand_count;
      if (total > a.max_nms) {
        h->overflow = 1;
        h->need = a.max_nms;
        h->prefix = 0ull;
        h->n_final = a.max_nms;
      } else {
        h->overflow = 0;
        h->n_final = total;
      }
    }
    return;
  }
  if (!h->overflow) return;
  const int nb = 1 << pass_bits(pass);
  for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x) cum[i] = i < nb ? hg[i] : 0u;
  __syncthreads();
  // inclusive scan over 2048 bins (Hillis-Steele in shared memory, 11 rounds)
  for (int off = 1; off < HIST_BINS; off <<= 1) {
    unsigned int v0 = 0, v1 = 0;
    int i0 = threadIdx.x, i1 = threadIdx.x + 1024;
    if (i0 >= off) v0 = cum[i0 - off];
    if (i1 >= off) v1 = cum[i1 - off];
    __syncthreads();
    cum[i0] += v0;
    cum[i1] += v1;
    __syncthreads();
  }
  const unsigned int need = (unsigned int)h->need;
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    unsigned int before = i ? cum[i - 1] : 0u;
    if (before < need && cum[i] >= need) {  // exactly one bin satisfies this
      h->prefix = (h->prefix << pass_bits(pass)) | (unsigned long long)i;
      h->need = (int)(need - before);
    }
  }
  for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x) hg[i] = 0u;
}

// ---------------------------------------------------------------------------------------------
// Bitonic sort of each image's key list (ascending).  Lists are padded with all-ones sentinels
// up to P = max(SORT_TILE, next_pow2(n)); tiles beyond P are skipped.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int padded_len(int n) {
  int p = SORT_TILE;
  while (p < n) p <<= 1;
  return p;
}

// phase 0: load (+pad), full sort of each tile;  phase 1: finish stage k (j = SORT_TILE/2 .. 1)
__global__ void __launch_bounds__(1024) nms_sort_tile_kernel(const NmsArgs a, int phase, int k) {
  __shared__ unsigned long long s[SORT_TILE];
  const int b = blockIdx.y;
  const int n = min(a.hdr[b].n_final, a.cap);
  if (n <= 1 && phase == 0) return;
  const int P = padded_len(n);
  const int t0 = blockIdx.x * SORT_TILE;
  if (t0 >= P) return;
  if (phase == 1 && k > P) return;
  unsigned long long* keys = a.keys + (size_t)b * a.cap + t0;
  for (int i = threadIdx.x; i < SORT_TILE; i += blockDim.x)
    s[i] = (phase == 0 && t0 + i >= n) ? ~0ull : keys[i];
  __syncthreads();
  const int kk_begin = phase == 0 ? 2 : k;
  const int kk_end = phase == 0 ? SORT_TILE : k;
  for (int kk = kk_begin; kk <= kk_end; kk <<= 1) {
    for (int j = min(kk >> 1, SORT_TILE >> 1); j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < SORT_TILE / 2; t += blockDim.x) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair
        int l = i | j;
        bool asc = (((t0 + i) & kk) == 0);
        unsigned long long x = s[i], y = s[l];
        if ((x > y) == asc) {
          s[i] = y;
          s[l] = x;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < SORT_TILE; i += blockDim.x) keys[i] = s[i];
}

// one global compare-exchange step (k, j) with j >= SORT_TILE
__global__ void __launch_bounds__(256) nms_sort_global_kernel(const NmsArgs a, int k, int j) {
  const int b = blockIdx.y;
  const int n = min(a.hdr[b].n_final, a.cap);
  const int P = padded_len(n);
  if (k > P) return;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P / 2) return;
  int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
  int l = i | j;
  unsigned long long* keys = a.keys + (size_t)b * a.cap;
  bool asc = ((i & k) == 0);
  unsigned long long x = keys[i], y = keys[l];
  if ((x > y) == asc) {
    keys[i] = y;
    keys[l] = x;
  }
}

// ---------------------------------------------------------------------------------------------
// Greedy walk.
// ---------------------------------------------------------------------------------------------
struct BoxF {
  float x1, y1, x2, y2, area;
};

// torchvision CPU nms_kernel_impl, operation by operation (std::max(a,b) = a < b ? b : a).
// Two exits ahead of the IEEE division leave the decision unchanged: boxes that do not intersect
// have inter == 0 (never > thr for thr >= 0; NaN compares false), and a 2-ulp fast quotient that is
// more than 1e-4 (relative) away from the threshold decides the same way as the rounded one.
__device__ __forceinline__ bool suppresses(const BoxF& i, const BoxF& j, double thr, float thr_lo, float thr_hi) {
  float xx1 = (i.x1 < j.x1) ? j.x1 : i.x1;
  float yy1 = (i.y1 < j.y1) ? j.y1 : i.y1;
  float xx2 = (j.x2 < i.x2) ? j.x2 : i.x2;
  float yy2 = (j.y2 < i.y2) ? j.y2 : i.y2;
  float dw = __fsub_rn(xx2, xx1), dh = __fsub_rn(yy2, yy1);
  if (thr_lo >= 0.f && (!(dw > 0.f) || !(dh > 0.f))) return false;
  float w = (0.f < dw) ? dw : 0.f;
  float h = (0.f < dh) ? dh : 0.f;
  float inter = __fmul_rn(w, h);
  float uni = __fsub_rn(__fadd_rn(i.area, j.area), inter);
  if (uni > 1e-30f && uni < 1e30f && inter < 1e30f) {
    float q = __fdividef(inter, uni);
    if (q > thr_hi) return true;
    if (q < thr_lo) return false;
  }
  float ovr = __fdiv_rn(inter, uni);
  return (double)ovr > thr;
}

static constexpr int G_CHUNK = 256;
static constexpr int G_MAXDET = 1024;  // shared-memory capacity for kept boxes

__device__ __forceinline__ void load_candidate(const NmsArgs& a, int b, unsigned long long key, BoxF& off,
                                               float* raw6) {
  unsigned int idx = (unsigned int)(key & 0xFFFFFFFFull);
  int an = (int)(idx / (unsigned int)a.nc);
  int c = (int)(idx - (unsigned int)an * (unsigned int)a.nc);
  const float* pb = a.pred + (size_t)b * (4 + a.nc) * a.A + an;
  float cx = __ldg(pb), cy = __ldg(pb + a.A), w = __ldg(pb + 2 * (size_t)a.A), h = __ldg(pb + 3 * (size_t)a.A);
  // wh2xy, util.py:76-82: x -/+ w / 2
  float hw = __fdiv_rn(w, 2.f), hh = __fdiv_rn(h, 2.f);
  float x1 = __fsub_rn(cx, hw), y1 = __fsub_rn(cy, hh), x2 = __fadd_rn(cx, hw), y2 = __fadd_rn(cy, hh);
  float fc = (float)c;
  float o = __fmul_rn(fc, a.max_wh);  // util.py:160
  off.x1 = __fadd_rn(x1, o);
  off.y1 = __fadd_rn(y1, o);
  off.x2 = __fadd_rn(x2, o);
  off.y2 = __fadd_rn(y2, o);
  off.area = __fmul_rn(__fsub_rn(off.x2, off.x1), __fsub_rn(off.y2, off.y1));
  if (raw6) {
    raw6[0] = x1;
    raw6[1] = y1;
    raw6[2] = x2;
    raw6[3] = y2;
    raw6[4] = from_orderable(~(unsigned int)(key >> 32));
    raw6[5] = fc;
  }
}

__global__ void __launch_bounds__(G_CHUNK) nms_greedy_kernel(const NmsArgs a) {
  __shared__ BoxF kept[G_MAXDET];
  __shared__ int kept_pos[G_MAXDET];            // position in the sorted list
  __shared__ BoxF live[G_CHUNK];
  __shared__ int live_pos[G_CHUNK];
  __shared__ unsigned int mask[G_CHUNK][G_CHUNK / 32];
  __shared__ int warp_cnt[G_CHUNK / 32];
  __shared__ int s_K, s_m;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(a.hdr[b].n_final, a.cap);
  const unsigned long long* keys = a.keys + (size_t)b * a.cap;
  const double thr = a.iou;
  // brackets of the threshold for the fast quotient (only used when 0 <= thr < 1e30)
  const bool thr_ok = thr >= 0.0 && thr < 1e30;
  const float thr_lo = thr_ok ? (float)thr * 0.9999f - 1e-30f : -1.f;
  const float thr_hi = thr_ok ? (float)thr * 1.0001f + 1e-30f : 3.0e38f;
  if (tid == 0) s_K = 0;
  __syncthreads();
  for (int base = 0; base < n; base += G_CHUNK) {
    const int K = s_K;
    if (K >= a.max_det) break;
    const int ci = base + tid;
    bool alive = ci < n;
    BoxF me;
    me.x1 = me.y1 = me.x2 = me.y2 = me.area = 0.f;
    if (alive) {
      load_candidate(a, b, keys[ci], me, nullptr);
      for (int k = 0; k < K; k++) {
        if (suppresses(kept[k], me, thr, thr_lo, thr_hi)) {
          alive = false;
          break;
        }
      }
    }
    // ordered compaction of the survivors
    unsigned int bal = __ballot_sync(0xffffffffu, alive);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0;
    for (int wi = 0; wi < warp; wi++) before += warp_cnt[wi];
    if (alive) {
      int pos = before + __popc(bal & ((1u << lane) - 1u));
      live[pos] = me;
      live_pos[pos] = ci;
    }
    if (tid == 0) {
      int m = 0;
      for (int wi = 0; wi < G_CHUNK / 32; wi++) m += warp_cnt[wi];
      s_m = m;
    }
    __syncthreads();
    const int m = s_m;
    // suppression bitmask inside the chunk: mask[i][w] bit l <=> live[i] suppresses live[32w+l], 32w+l > i
    for (int task = warp; task < m * (G_CHUNK / 32); task += G_CHUNK / 32) {
      int i = task / (G_CHUNK / 32), wcol = task - i * (G_CHUNK / 32);
      int j = wcol * 32 + lane;
      bool sup = false;
      if (j > i && j < m) sup = suppresses(live[i], live[j], thr, thr_lo, thr_hi);
      unsigned int bits = __ballot_sync(0xffffffffu, sup);
      if (lane == 0) mask[i][wcol] = bits;
    }
    __syncthreads();
    if (warp == 0) {
      unsigned int remv = 0;  // lane w (< 8) holds removed-bits word w
      int Kc = K;
      for (int i = 0; i < m; i++) {
        unsigned int word = __shfl_sync(0xffffffffu, remv, i >> 5);
        if (!((word >> (i & 31)) & 1u)) {
          if (lane == 0) {
            kept[Kc] = live[i];
            kept_pos[Kc] = live_pos[i];
          }
          Kc++;
          if (Kc >= a.max_det) break;
          if (lane < G_CHUNK / 32) remv |= mask[i][lane];
        }
      }
      if (lane == 0) s_K = Kc;
    }
    __syncthreads();
  }
  const int K = s_K;
  if (tid == 0) a.out_counts[b] = K;
  for (int k = tid; k < K; k += blockDim.x) {
    BoxF tmp;
    float r[6];
    load_candidate(a, b, keys[kept_pos[k]], tmp, r);
    float* op = a.out + ((size_t)b * a.max_det + k) * 6;
#pragma unroll
    for (int q = 0; q < 6; q++) op[q] = r[q];
  }
}

// ---------------------------------------------------------------------------------------------
static int cap_for(int max_nms) {
  int c = SORT_TILE;
  while (c < max_nms) c <<= 1;
  return c;
}

size_t nms_workspace_bytes(int B, int nc, int A, int max_nms) {
  (void)nc;
  (void)A;
  size_t hdr = ((size_t)B * sizeof(NmsHeader) + 255) / 256 * 256;
  size_t hist = (size_t)B * HIST_BINS * 4;
  size_t keys = (size_t)B * cap_for(max_nms) * 8;
  return hdr + hist + keys;
}

int nms_run(const float* pred, int B, int nc, int A, float conf, double iou, int max_det, int max_nms,
            float max_wh, float* out, int* out_counts, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (B <= 0 || nc <= 0 || A <= 0 || max_det <= 0 || max_nms <= 0) {
    set_error("yb_nms: bad sizes B=%d nc=%d A=%d max_det=%d max_nms=%d", B, nc, A, max_det, max_nms);
    return YB_ERR_ARG;
  }
  if (max_det > G_MAXDET) {
    set_error("yb_nms: max_det %d exceeds the kernel capacity %d", max_det, G_MAXDET);
    return YB_ERR_UNSUPPORTED;
  }
  if ((unsigned long long)A * (unsigned long long)nc >= (1ull << 32)) {
    set_error("yb_nms: A*nc does not fit the 32-bit candidate index");
    return YB_ERR_UNSUPPORTED;
  }
  size_t need = nms_workspace_bytes(B, nc, A, max_nms);
  if (ws_bytes < need || !ws) {
    set_error("yb_nms: workspace too small (%zu < %zu)", ws_bytes, need);
    return YB_ERR_ARG;
  }
  NmsArgs a;
  a.pred = pred;
  a.B = B;
  a.nc = nc;
  a.A = A;
  a.conf = conf;
  a.iou = iou;
  a.max_det = max_det;
  a.max_nms = max_nms;
  a.cap = cap_for(max_nms);
  a.max_wh = max_wh;
  size_t hdr_bytes = ((size_t)B * sizeof(NmsHeader) + 255) / 256 * 256;
  a.hdr = reinterpret_cast<NmsHeader*>(ws);
  a.hist = reinterpret_cast<unsigned int*>((uint8_t*)ws + hdr_bytes);
  a.keys = reinterpret_cast<unsigned long long*>((uint8_t*)ws + hdr_bytes + (size_t)B * HIST_BINS * 4);
  a.out = out;
  a.out_counts = out_counts;
  YB_CUDA(cudaMemsetAsync(ws, 0, hdr_bytes + (size_t)B * HIST_BINS * 4, st));
  long long total = (long long)nc * A;
  int gx = (int)std::min<long long>((total + 256 * 16 - 1) / (256 * 16), 1024);
  int vec4 = (A % 4 == 0) && ((reinterpret_cast<uintptr_t>(pred) & 15) == 0);
  nms_append_kernel<<<dim3(gx, B), 256, 0, st>>>(a, vec4);
  count_launch();
  nms_pick_kernel<<<B, 32, 0, st>>>(a, -1);
  count_launch();
  // radix select + re-compaction: every block returns at once unless its image overflowed
  dim3 ogrid(std::min(gx, 48), B);
  for (int pass = 0; pass < NUM_PASSES; pass++) {
    nms_scan_kernel<1><<<ogrid, 256, 0, st>>>(a, pass);
    count_launch();
    nms_pick_kernel<<<B, 1024, 0, st>>>(a, pass);
    count_launch();
  }
  nms_scan_kernel<2><<<ogrid, 256, 0, st>>>(a, 0);
  count_launch();
  dim3 tgrid(a.cap / SORT_TILE, B);
  nms_sort_tile_kernel<<<tgrid, 1024, 0, st>>>(a, 0, 0);
  count_launch();
  for (int k = 2 * SORT_TILE; k <= a.cap; k <<= 1) {
    for (int j = k >> 1; j >= SORT_TILE; j >>= 1) {
      dim3 ggrid(a.cap / 2 / 256, B);
      nms_sort_global_kernel<<<ggrid, 256, 0, st>>>(a, k, j);
      count_launch();
    }
    nms_sort_tile_kernel<<<tgrid, 1024, 0, st>>>(a, 1, k);
    count_launch();
  }
  nms_greedy_kernel<<<B, G_CHUNK, 0, st>>>(a);
  count_launch();
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

}  // namespace yb
