// non_max_suppression on the device (reference utils/util.py:123-169, torchvision.ops.nms CPU
// semantics for the greedy step).  No host synchronisation, no data-dependent launch shapes:
//
//   1. append    one coalesced pass over (B, 4+nc, A): every score > conf becomes a 64-bit key
//                  key = ~orderable(score) << 32 | (anchor * nc + class)
//                ascending key order == descending score, ties by ascending candidate index
//                (anchor-major, class-minor: the row-major order of util.py:147's nonzero()).
//                Keys are appended to a per-image list while it has room.
//   2. walk      one CTA (1024 threads) per image consumes the candidates lazily, in descending
//                score order, one *band* at a time: a 2048-bin histogram of the score bits picks a
//                key range holding <= 4096 candidates, the band is compacted into shared memory,
//                bitonic-sorted there and handed to the greedy step; the next band is fetched only
//                if fewer than max_det boxes are kept so far.  Greedy NMS only ever needs the first
//                max_det kept boxes (util.py:163) and only the max_nms best candidates
//                (util.py:157), so a full sort of the list and the n x n mask of torchvision's CUDA
//                kernel are never built.  A histogram bin with more than 4096 candidates (massive
//                score ties) is split exactly by re-histogramming the lower key bits.  Images with
//                more candidates than the key list holds are banded straight from the scores.
//      greedy    256 candidates at a time: each is tested against the boxes kept so far (4 threads
//                per candidate), survivors are resolved inside the chunk with a ballot-built IoU
//                bitmask and a serial scan.
//
// IoU arithmetic replicates torchvision's CPU kernel operation by operation in fp32 with
// round-to-nearest intrinsics (no FMA contraction), on the class-offset boxes of util.py:160-161,
// and compares (double)iou > iou_threshold.
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <string.h>

#include "yb_internal.h"

namespace yb {

static constexpr int HIST_BINS = 2048;
static constexpr int SORT_TILE = 4096;

struct NmsHeader {  // per image, zero at the start of every call
  int cand_count;   // candidates found by the scan
  int sel_count;    // keys appended by the scan (may exceed capacity; clamp on read)
  int sel2_count;   // overflow images: keys re-appended by the select pass
  int selected;     // overflow images: 1 = the list holds every key of the max_nms best (and possibly a few more)
  int pad[4];
};

struct NmsArgs {
  const float* pred;
  int B, nc, A;
  float conf;
  double iou;
  int max_det, max_nms, cap;
  int first_band, next_band;   // candidates per band (<= SORT_TILE)
  float max_wh;
  NmsHeader* hdr;
  unsigned int* ghist;       // [B][HIST_BINS], zero between calls (overflow images only)
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

// Pass over the scores of every image: each score > conf becomes a key appended to the image's
// list (unordered; the sort orders them).  Reads 4 anchors per thread (float4 when aligned), one
// atomic per warp iteration.
__global__ void __launch_bounds__(256) nms_append_kernel(const NmsArgs a, int vec4) {
  pdl_prologue_done();
  pdl_wait();
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

#ifndef YB_G_CHUNK
#define YB_G_CHUNK 128
#endif
static constexpr int G_CHUNK = YB_G_CHUNK;   // candidates per greedy step
static constexpr int G_MAXDET = 1024;  // shared-memory capacity for kept boxes
static constexpr int FIRST_BAND = 1024;  // preferred size of the first band (most images finish inside it)
static constexpr int BAND_PRE = 1024;    // candidates of a band whose boxes are gathered in one go (one latency per band)
static constexpr int NEXT_BAND = 1024;   // and of the following ones: sorting 4 x 1024 keys costs less than 1 x 4096

// Returns the candidate's class code: its class when every raw coordinate lies inside
// (-max_wh/2, max_wh/2) - then boxes of different classes cannot intersect after the per-class offset of
// util.py:160, so only same-class pairs need an IoU - and -1 otherwise (NaN / huge boxes: tested against
// everything, exactly as the reference's single nms over offset boxes would).
__device__ __forceinline__ int load_candidate(const NmsArgs& a, int b, unsigned long long key, BoxF& off,
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
  const float lim = 0.5f * a.max_wh;
  const bool normal = fabsf(x1) < lim && fabsf(y1) < lim && fabsf(x2) < lim && fabsf(y2) < lim;
  return normal ? c : -1;
}

// Histogram bins over the high key word (= bit-inverted orderable score): bin 0 starts at score 1.0
// and every bin spans 2^16 consecutive fp32 values, so [conf = 0.001, 1] covers ~1277 bins; anything
// outside is clamped, which keeps bins monotone in the key order (all that correctness needs).
static constexpr long long KH_ONE = 0x407FFFFFll;  // high key word of score 1.0f
__device__ __forceinline__ int bin_of(unsigned long long key) {
  long long d = ((long long)(unsigned int)(key >> 32) - KH_ONE) >> 16;
  return (int)(d < 0 ? 0 : (d > HIST_BINS - 1 ? HIST_BINS - 1 : d));
}
// smallest key of bin b (b in [0, HIST_BINS]); ~0 is never a real key
__device__ __forceinline__ unsigned long long edge_key(int b) {
  if (b <= 0) return 0ull;
  if (b >= HIST_BINS) return ~0ull;
  return (unsigned long long)(KH_ONE + ((long long)b << 16)) << 32;
}

struct ImgSmem {
  unsigned long long tile[SORT_TILE];
  unsigned int hist[HIST_BINS];
  unsigned int hist2[HIST_BINS];
  BoxF kept[G_MAXDET];
  unsigned long long kept_key[G_MAXDET];
  short kept_code[G_MAXDET];                 // class code of the kept box (-1: abnormal, tested against everything)
  BoxF band_box[BAND_PRE];                   // boxes of the band's first candidates, gathered once after the sort
  short band_code[BAND_PRE];
  BoxF live[G_CHUNK];      // candidates of the chunk, then (compacted in place) its survivors
  unsigned long long live_key[G_CHUNK];
  int live_code[G_CHUNK];
  unsigned int maskT[G_CHUNK][G_CHUNK / 32]; // bit i of row j: live[i] (i < j) suppresses live[j]
  unsigned int decided[G_CHUNK / 32], keptbits[G_CHUNK / 32];
  int dead[G_CHUNK];
  int warp_cnt[G_CHUNK / 32];
  int K, m, cnt, walk_end;
  unsigned int walk_cum;
  unsigned int scan_tot[32];
  int st_bands, st_chunks, st_surv, st_tests;   // -DYB_NMS_STATS: work counters, copied to the header pad
  int st_cyc[8];                                // cycles: 0 hist, 1 select+compact, 2 sort, 3 load+A, 4 compaction, 5 B, 6 C, 7 D
};

// The image's candidates: the appended key list when it is complete, else the raw scores.
struct Src {
  const unsigned long long* keys;
  int n;
  const float* sp;
  long long total;
  int A, nc;
  float conf;
  bool complete;
};
// Calls f(valid, key) with a uniform trip count over the CTA (f may use warp collectives).
template <int IMG_T, typename F>
__device__ __forceinline__ void scan_src(const Src& s, F f) {
  if (s.complete) {
    // four independent loads in flight per thread: the list lives in L2 / HBM, its latency is the cost
    const int iters = (s.n + IMG_T - 1) / IMG_T;
    int i = threadIdx.x;
    for (int it = 0; it < iters; it += 4, i += 4 * IMG_T) {
      unsigned long long k[4];
#pragma unroll
      for (int u = 0; u < 4; u++) k[u] = (i + u * IMG_T < s.n) ? s.keys[i + u * IMG_T] : ~0ull;
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (it + u < iters) f(i + u * IMG_T < s.n, k[u]);
    }
  } else {
    const long long iters = (s.total + IMG_T - 1) / IMG_T;
    long long e = threadIdx.x;
    for (long long it = 0; it < iters; it++, e += IMG_T) {
      bool v = false;
      unsigned long long key = ~0ull;
      if (e < s.total) {
        const float sc = __ldg(s.sp + e);
        if (sc > s.conf) {
          const int c = (int)(e / s.A);
          const int an = (int)(e - (long long)c * s.A);
          key = make_key(sc, (unsigned int)an * (unsigned int)s.nc + (unsigned int)c);
          v = true;
        }
      }
      f(v, key);
    }
  }
}

// Longest run of bins [start, end) with sum <= limit (trailing empty bins included), by the whole CTA:
// per-thread partial sums, a two-level scan, then every thread tests its own bins.  Results in
// sm.walk_end / sm.walk_cum.  Must be called by all threads; ends with a barrier.
template <int IMG_T>
__device__ __forceinline__ void walk_bins(ImgSmem& sm, const unsigned int* h, int start, int nb, unsigned int limit) {
  constexpr int BPT = HIST_BINS / IMG_T;   // consecutive bins per thread
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned int v[BPT], tot = 0;
#pragma unroll
  for (int q = 0; q < BPT; q++) {
    const int i = tid * BPT + q;
    v[q] = (i >= start && i < nb) ? h[i] : 0u;
    tot += v[q];
  }
  unsigned int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) sm.scan_tot[warp] = incl;
  if (tid == 0) {
    sm.walk_end = start;
    sm.walk_cum = 0u;
  }
  __syncthreads();
  unsigned int base = incl - tot;
  for (int w = 0; w < warp; w++) base += sm.scan_tot[w];
  // cum is monotone over bins: the run ends after the last bin whose inclusive sum still fits
  unsigned int cum = base;
#pragma unroll
  for (int q = 0; q < BPT; q++) {
    const int i = tid * BPT + q;
    cum += v[q];
    if (i >= start && i < nb && cum <= limit) atomicMax(&sm.walk_end, i + 1);
  }
  __syncthreads();
  cum = base;
#pragma unroll
  for (int q = 0; q < BPT; q++) {
    const int i = tid * BPT + q;
    cum += v[q];
    if (i + 1 == sm.walk_end && i >= start) sm.walk_cum = cum;
  }
  __syncthreads();
}

#ifdef YB_NMS_STATS
#define YB_T(slot)                                      \
  do {                                                  \
    if (tid == 0) {                                     \
      const long long _n = clock64();                   \
      sm.st_cyc[slot] += (int)(_n - st_last);           \
      st_last = _n;                                     \
    }                                                   \
  } while (0)
#else
#define YB_T(slot) do { } while (0)
#endif
// ---- images with more candidates than the key list holds (dense predictions, util.py:157's [:max_nms] cut) ----
// Only the max_nms best candidates can ever reach the greedy step.  Two grid-wide passes over the scores of such
// an image (every block of the other images returns at once) rebuild its key list so that it holds all of them:
// a histogram of the score bins, then a compaction of every key up to the bin in which rank max_nms falls.  The
// per-image kernel then bands from the list as usual; if even that prefix does not fit the list (a single fat
// bin), it falls back to banding from the raw scores.
__global__ void __launch_bounds__(256) nms_ovf_hist_kernel(const NmsArgs a) {
  __shared__ unsigned int hs[HIST_BINS];
  pdl_prologue_done();
  pdl_wait();
  const int b = blockIdx.y;
  if (a.hdr[b].sel_count <= a.cap) return;
  for (int i = threadIdx.x; i < HIST_BINS; i += 256) hs[i] = 0u;
  __syncthreads();
  const long long total = (long long)a.nc * a.A;
  const float* sp = a.pred + ((size_t)b * (4 + a.nc) + 4) * a.A;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const float sc = __ldg(sp + e);
    if (sc > a.conf) atomicAdd(&hs[bin_of(make_key(sc, 0u))], 1u);
  }
  __syncthreads();
  unsigned int* g = a.ghist + (size_t)b * HIST_BINS;
  for (int i = threadIdx.x; i < HIST_BINS; i += 256)
    if (hs[i]) atomicAdd(&g[i], hs[i]);
}

__global__ void __launch_bounds__(256) nms_ovf_select_kernel(const NmsArgs a) {
  __shared__ unsigned int part[256];
  __shared__ int s_T;
  pdl_prologue_done();
  pdl_wait();
  const int b = blockIdx.y;
  NmsHeader* h = a.hdr + b;
  if (h->sel_count <= a.cap) return;
  // threshold bin: the first bin T with count(bins <= T) >= max_nms (every block derives it on its own)
  const unsigned int* g = a.ghist + (size_t)b * HIST_BINS;
  constexpr int BPT = HIST_BINS / 256;
  unsigned int v[BPT], tot = 0;
#pragma unroll
  for (int q = 0; q < BPT; q++) {
    v[q] = g[threadIdx.x * BPT + q];
    tot += v[q];
  }
  part[threadIdx.x] = tot;
  if (threadIdx.x == 0) s_T = HIST_BINS - 1;
  __syncthreads();
  unsigned int base = 0;
  for (int i = 0; i < (int)threadIdx.x; i++) base += part[i];
  unsigned int cum = base;
#pragma unroll
  for (int q = 0; q < BPT; q++) {
    const unsigned int before = cum;
    cum += v[q];
    if (before < (unsigned int)a.max_nms && cum >= (unsigned int)a.max_nms) s_T = threadIdx.x * BPT + q;   // unique
  }
  __syncthreads();
  const int T = s_T;
  unsigned int upto = 0;   // keys with bin <= T
  for (int i = 0; i <= T; i++) upto += g[i];
  if (upto > (unsigned int)a.cap) return;   // does not fit: the per-image kernel bands from the raw scores
  if (blockIdx.x == 0 && threadIdx.x == 0) h->selected = 1;
  const long long total = (long long)a.nc * a.A;
  const float* sp = a.pred + ((size_t)b * (4 + a.nc) + 4) * a.A;
  unsigned long long* keys = a.keys + (size_t)b * a.cap;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * 256;
  const long long iters = (total + stride - 1) / stride;
  long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  for (long long it = 0; it < iters; it++, e += stride) {
    bool take = false;
    unsigned long long key = 0;
    if (e < total) {
      const float sc = __ldg(sp + e);
      if (sc > a.conf) {
        const int c = (int)(e / a.A);
        const int an = (int)(e - (long long)c * a.A);
        key = make_key(sc, (unsigned int)an * (unsigned int)a.nc + (unsigned int)c);
        take = bin_of(key) <= T;
      }
    }
    const unsigned int m = __ballot_sync(0xffffffffu, take);
    if (m) {
      const int leader = __ffs(m) - 1;
      int slot0 = 0;
      if (lane == leader) slot0 = atomicAdd(&h->sel2_count, __popc(m));
      slot0 = __shfl_sync(0xffffffffu, slot0, leader);
      if (take) {
        const int slot = slot0 + __popc(m & ((1u << lane) - 1u));
        if (slot < a.cap) keys[slot] = key;
      }
    }
  }
}

// IMG_T threads per image: 1024 for small batches (latency), 512 (two CTAs per SM) for large ones.
template <int IMG_T>
__global__ void __launch_bounds__(IMG_T) nms_image_kernel(const NmsArgs a) {
  extern __shared__ __align__(16) uint8_t nms_smem_raw[];
  ImgSmem& sm = *reinterpret_cast<ImgSmem*>(nms_smem_raw);
  pdl_prologue_done();
  pdl_wait();
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const NmsHeader h = a.hdr[b];
  Src src;
  src.complete = h.sel_count <= a.cap || h.selected;   // selected: the list was rebuilt with the max_nms best
  src.keys = a.keys + (size_t)b * a.cap;
  src.n = h.selected ? min(h.sel2_count, a.cap) : min(h.sel_count, a.cap);
  src.sp = a.pred + ((size_t)b * (4 + a.nc) + 4) * a.A;
  src.total = (long long)a.nc * a.A;
  src.A = a.A;
  src.nc = a.nc;
  src.conf = a.conf;
  const double thr = a.iou;
  // brackets of the threshold for the fast quotient (only used when 0 <= thr < 1e30)
  const bool thr_ok = thr >= 0.0 && thr < 1e30;
  const float thr_lo = thr_ok ? (float)thr * 0.9999f - 1e-30f : -1.f;
  const float thr_hi = thr_ok ? (float)thr * 1.0001f + 1e-30f : 3.0e38f;

  for (int i = tid; i < HIST_BINS; i += IMG_T) sm.hist[i] = 0u;
  if (tid == 0) sm.K = 0;
#ifdef YB_NMS_STATS
  long long st_last = clock64();
  if (tid == 0) {
    sm.st_bands = sm.st_chunks = sm.st_surv = sm.st_tests = 0;
    for (int i = 0; i < 8; i++) sm.st_cyc[i] = 0;
  }
#endif
  // class-restricted pair tests need "no intersection => not suppressed" (thr >= 0) and a sane offset
  const bool by_class = thr_ok && a.max_wh > 0.f && a.nc <= 0x7FFF;
  __syncthreads();
  if (h.cand_count > 0) {
    scan_src<IMG_T>(src, [&](bool v, unsigned long long k) {
      if (v) atomicAdd(&sm.hist[bin_of(k)], 1u);
    });
  }
  __syncthreads();

  YB_T(0);
  int cur_bin = 0;
  int consumed = 0;                 // candidates handed to the greedy step so far (sorted positions)
  unsigned long long lo_done = 0;   // every key below this has been consumed
  bool first = true;
  while (h.cand_count > 0 && cur_bin < HIST_BINS && consumed < a.max_nms && sm.K < a.max_det) {
    // ---- choose the band [lo_done, band_hi)
    walk_bins<IMG_T>(sm, sm.hist, cur_bin, HIST_BINS, first ? a.first_band : a.next_band);
    int b_end = sm.walk_end;
    unsigned int cnt = sm.walk_cum;
    unsigned long long band_hi;
    int next_bin;
    if (b_end == HIST_BINS && cnt == 0) break;  // nothing left
    if (b_end > cur_bin && cnt > 0) {
      band_hi = edge_key(b_end);
      next_bin = b_end;
    } else {
      // skip the empty bins walked over; the bin at b_end alone exceeds the limit
      const int fat = b_end;
      const unsigned int fat_cnt = sm.hist[fat];
      if (fat_cnt <= (unsigned int)SORT_TILE) {
        band_hi = edge_key(fat + 1);
        cnt = fat_cnt;
        next_bin = fat + 1;
      } else {
        // exact split of a bin with more than SORT_TILE keys: re-histogram the lower key bits
        unsigned long long lo = lo_done > edge_key(fat) ? lo_done : edge_key(fat);
        unsigned long long hi = edge_key(fat + 1);
        for (;;) {
          const unsigned long long span = hi - lo;
          const int need_bits = 64 - __clzll((long long)(span - 1));
          const int shift = need_bits > 11 ? need_bits - 11 : 0;
          __syncthreads();
          for (int i = tid; i < HIST_BINS; i += IMG_T) sm.hist2[i] = 0u;
          __syncthreads();
          scan_src<IMG_T>(src, [&](bool v, unsigned long long k) {
            if (v && k >= lo && k < hi) atomicAdd(&sm.hist2[(unsigned int)((k - lo) >> shift)], 1u);
          });
          __syncthreads();
          walk_bins<IMG_T>(sm, sm.hist2, 0, HIST_BINS, SORT_TILE);
          const int e = sm.walk_end;
          if (e == 0) {  // the first sub-bin alone is still too fat: descend into it
            hi = lo + (1ull << shift);
            continue;
          }
          cnt = sm.walk_cum;
          band_hi = (((span - 1) >> shift) < (unsigned long long)e) ? hi : lo + ((unsigned long long)e << shift);
          break;
        }
        __syncthreads();
        if (tid == 0) sm.hist[fat] -= cnt;  // the rest of the bin stays for the next round
        next_bin = fat;
      }
    }
    YB_T(1);
    // ---- compact the band into the tile, pad, sort ascending (= descending score)
    if (tid == 0) sm.cnt = 0;
    __syncthreads();
    {
      const unsigned long long lo = lo_done;
      scan_src<IMG_T>(src, [&](bool v, unsigned long long k) {
        const bool take = v && k >= lo && k < band_hi;
        const unsigned int m = __ballot_sync(0xffffffffu, take);
        if (m) {
          const int leader = __ffs(m) - 1;
          int base = 0;
          if (lane == leader) base = atomicAdd(&sm.cnt, __popc(m));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (take) {
            const int slot = base + __popc(m & ((1u << lane) - 1u));
            if (slot < SORT_TILE) sm.tile[slot] = k;
          }
        }
      });
    }
    __syncthreads();
    YB_T(2);
    const int n_band = min(sm.cnt, SORT_TILE);
    int P = 32;
    while (P < n_band) P <<= 1;
    for (int i = n_band + tid; i < P; i += IMG_T) sm.tile[i] = ~0ull;
    __syncthreads();
    // bitonic network; element e lives at thread e % IMG_T.  Stages with partner distance j >= 32 run in
    // shared memory (one barrier each); the j = 16 .. 1 tail of every kk (and all of kk <= 32) runs in
    // registers with warp shuffles: one barrier per kk instead of five.
    for (int kk = 2; kk <= P; kk <<= 1) {
      int j = kk >> 1;
      for (; j >= 32; j >>= 1) {
        for (int t = tid; t < P / 2; t += IMG_T) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair
          const int l = i | j;
          const bool asc = ((i & kk) == 0);
          const unsigned long long x = sm.tile[i], y = sm.tile[l];
          if ((x > y) == asc) {
            sm.tile[i] = y;
            sm.tile[l] = x;
          }
        }
        __syncthreads();
      }
      for (int e = tid; e < P; e += IMG_T) {   // P >= 32 and IMG_T % 32 == 0: whole warps take this branch
        unsigned long long x = sm.tile[e];
        const bool asc = ((e & kk) == 0);
        for (int jj = j; jj > 0; jj >>= 1) {
          const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, jj);
          const bool lower = (lane & jj) == 0;
          const bool take_min = lower == asc;
          x = take_min ? (x < y ? x : y) : (x < y ? y : x);
        }
        sm.tile[e] = x;
      }
      __syncthreads();
    }
    YB_T(3);
#ifdef YB_NMS_STATS
    if (tid == 0) sm.st_bands++;
#endif
    // ---- greedy step over the sorted band, 256 candidates at a time
    const int n_use = min(n_band, a.max_nms - consumed);  // util.py:157: only the max_nms best
    for (int i = tid; i < min(n_use, BAND_PRE); i += IMG_T) {
      BoxF bx;
      const int code = load_candidate(a, b, sm.tile[i], bx, nullptr);
      sm.band_box[i] = bx;
      sm.band_code[i] = (short)((by_class && code >= 0) ? code : -1);
    }
    __syncthreads();
    for (int base = 0; base < n_use; base += G_CHUNK) {
      const int K = sm.K;
      if (K >= a.max_det) break;
      const int j = tid & (G_CHUNK - 1), part = tid / G_CHUNK;
      const int ci = base + j;
      // -- A: load the chunk's candidates, then test each against the kept boxes it can intersect (same
      //       class code, or either side abnormal); IMG_T / 256 threads share a candidate's kept list
      if (tid < G_CHUNK) {
        int code = -1;
        BoxF cand;
        cand.x1 = cand.y1 = cand.x2 = cand.y2 = cand.area = 0.f;
        if (ci < n_use) {
          if (ci < BAND_PRE) {
            cand = sm.band_box[ci];
            code = sm.band_code[ci];
          } else {
            code = load_candidate(a, b, sm.tile[ci], cand, nullptr);
            if (!by_class) code = -1;
          }
        }
        sm.live[j] = cand;
        sm.live_code[j] = code;
        sm.dead[j] = ci < n_use ? 0 : 1;
      }
      if (tid < G_CHUNK / 32) {
        sm.decided[tid] = 0u;
        sm.keptbits[tid] = 0u;
      }
      __syncthreads();
      if (ci < n_use) {
        const BoxF me = sm.live[j];
        const int code = sm.live_code[j];
        // (measured: testing four kept boxes per step for instruction-level parallelism, with an early exit once a
        // sibling thread has found a suppressor, changes nothing - 0.717 -> 0.727 ms batch-1 forward + NMS)
        for (int k = part; k < K; k += IMG_T / G_CHUNK) {
          const int kc = sm.kept_code[k];
          if (code >= 0 && kc >= 0 && kc != code) continue;
          if (suppresses(sm.kept[k], me, thr, thr_lo, thr_hi)) {
            sm.dead[j] = 1;
            break;
          }
        }
      }
      __syncthreads();
      // -- ordered compaction of the survivors (threads 0..255, in place: pos <= j)
      bool alive = false;
      BoxF me;
      int code = -1;
      unsigned int bal = 0;
      if (tid < G_CHUNK) {
        alive = !sm.dead[j];
        me = sm.live[j];
        code = sm.live_code[j];
        bal = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) sm.warp_cnt[warp] = __popc(bal);
      }
      __syncthreads();
      if (tid < G_CHUNK) {
        int before = 0;
        for (int wi = 0; wi < warp; wi++) before += sm.warp_cnt[wi];
        if (alive) {
          const int pos = before + __popc(bal & ((1u << lane) - 1u));
          sm.live[pos] = me;
          sm.live_key[pos] = sm.tile[ci];
          sm.live_code[pos] = code;
        }
        if (tid == 0) {
          int m = 0;
          for (int wi = 0; wi < G_CHUNK / 32; wi++) m += sm.warp_cnt[wi];
          sm.m = m;
        }
      }
      __syncthreads();
      YB_T(4);
      const int m = sm.m;
#ifdef YB_NMS_STATS
      if (tid == 0) { sm.st_chunks++; sm.st_surv += m; }
#endif
      // -- B: who suppresses whom inside the chunk (transposed: row j holds its potential suppressors i < j)
      for (int task = warp; task < m * (G_CHUNK / 32); task += IMG_T / 32) {
        const int jj = task / (G_CHUNK / 32), wcol = task - jj * (G_CHUNK / 32);
        if (wcol * 32 >= jj) {   // no i < jj in this word
          if (lane == 0) sm.maskT[jj][wcol] = 0u;
          continue;
        }
        const int i = wcol * 32 + lane;
        bool sup = false;
        if (i < jj) {
          const int cj = sm.live_code[jj], cI = sm.live_code[i];
          if (cj < 0 || cI < 0 || cj == cI) sup = suppresses(sm.live[i], sm.live[jj], thr, thr_lo, thr_hi);
        }
        const unsigned int bits = __ballot_sync(0xffffffffu, sup);
        if (lane == 0) sm.maskT[jj][wcol] = bits;
      }
      __syncthreads();
      YB_T(5);
      // -- C: greedy resolution in rounds.  live[j] is dead once a kept earlier box suppresses it, kept once
      //       all its potential suppressors are decided and none of them is kept.  Chains only form inside a
      //       class, so a couple of rounds settle the chunk (the fixed point is the serial greedy result).
      //       Only the warps that hold survivors take part: their rounds synchronise through a named barrier
      //       over those warps (a single warp: __syncwarp / vote) instead of two block-wide barriers per round -
      //       after the test against the kept boxes a chunk of 128 rarely has more than a few dozen survivors.
      const int nw = (m + 31) >> 5;   // uniform
      if (warp < nw) {
        const int nthr = nw * 32;
        bool done_j = tid >= m;   // threads beyond the survivors have nothing to decide
        for (int round = 0; round < G_CHUNK + 1; round++) {
          bool now = false, keep = false;
          if (!done_j) {
            bool pend = false, dead = false;
            for (int w = 0; w <= (tid >> 5); w++) {
              const unsigned int sp = sm.maskT[tid][w];
              if (sp & ~sm.decided[w]) pend = true;
              if (sp & sm.keptbits[w]) dead = true;
            }
            if (dead) now = true;
            else if (!pend) now = keep = true;
          }
          // every participating thread has read this round's snapshot
          if (nw == 1) __syncwarp();
          else asm volatile("bar.sync 2, %0;" ::"r"(nthr) : "memory");
          if (now) {
            if (keep) atomicOr(&sm.keptbits[tid >> 5], 1u << (tid & 31));
            atomicOr(&sm.decided[tid >> 5], 1u << (tid & 31));
            done_j = true;
          }
          bool all_done;
          if (nw == 1) {
            __syncwarp();
            all_done = __all_sync(0xffffffffu, done_j);
          } else {
            unsigned int r;
            asm volatile(
                "{ .reg .pred p, q; setp.ne.u32 p, %1, 0; barrier.cta.red.and.pred q, 2, %2, p; selp.u32 %0, 1, 0, q; }"
                : "=r"(r)
                : "r"((unsigned int)done_j), "r"(nthr)
                : "memory");
            all_done = r != 0u;
          }
#ifdef YB_NMS_STATS
          if (tid == 0) sm.st_tests++;   // (stats build: resolution rounds)
#endif
          if (all_done) break;
        }
        YB_T(6);
        // -- D: append the kept survivors in order (up to max_det)
        if (tid < m && ((sm.keptbits[tid >> 5] >> (tid & 31)) & 1u)) {
          int rank = __popc(sm.keptbits[tid >> 5] & ((1u << (tid & 31)) - 1u));
          for (int w = 0; w < (tid >> 5); w++) rank += __popc(sm.keptbits[w]);
          const int pos = K + rank;
          if (pos < a.max_det) {
            sm.kept[pos] = sm.live[tid];
            sm.kept_key[pos] = sm.live_key[tid];
            const int code_k = sm.live_code[tid];
            sm.kept_code[pos] = (short)(code_k >= 0 ? (code_k & 0x7FFF) : -1);
          }
        }
        if (tid == 0) {
          int total = 0;
          for (int w = 0; w < G_CHUNK / 32; w++) total += __popc(sm.keptbits[w]);
          sm.K = min(K + total, a.max_det);
        }
      }
      __syncthreads();
      YB_T(7);
    }
    consumed += n_band;
    lo_done = band_hi;
    cur_bin = next_bin;
    first = false;
    __syncthreads();
  }
  __syncthreads();
  const int K = sm.K;
  if (tid == 0) {
    a.out_counts[b] = K;
    // leave the header zeroed for the next call on this workspace (yb_nms then needs no memset)
    a.hdr[b].cand_count = 0;
    a.hdr[b].sel_count = 0;
    a.hdr[b].sel2_count = 0;
    a.hdr[b].selected = 0;
#ifdef YB_NMS_STATS
    // (stats build only: the header is not left zero)  -DYB_NMS_STATS=1: cycles per phase, =2: work counters
    int* hw = reinterpret_cast<int*>(&a.hdr[b]);
    if (YB_NMS_STATS == 2) {
      hw[0] = sm.st_bands; hw[1] = sm.st_chunks; hw[2] = sm.st_surv; hw[3] = sm.st_tests;
      for (int i = 4; i < 8; i++) hw[i] = 0;
    } else {
      for (int i = 0; i < 8; i++) hw[i] = sm.st_cyc[i];
    }
#endif
  }
  if (h.sel_count > a.cap)   // overflow image: leave its global histogram zeroed for the next call
    for (int i = tid; i < HIST_BINS; i += IMG_T) a.ghist[(size_t)b * HIST_BINS + i] = 0u;
  for (int k = tid; k < K; k += IMG_T) {
    BoxF tmp;
    float r[6];
    load_candidate(a, b, sm.kept_key[k], tmp, r);
    float* op = a.out + ((size_t)b * a.max_det + k) * 6;
#pragma unroll
    for (int q = 0; q < 6; q++) op[q] = r[q];
  }
  // rows past the last detection are zero (the caller does not have to clear the buffer)
  float* tail = a.out + ((size_t)b * a.max_det + K) * 6;
  for (int i = tid; i < (a.max_det - K) * 6; i += IMG_T) tail[i] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// Key-list capacity per image.  The list only has to hold max_nms keys for correctness, but an image whose
// candidates all fit is banded from the list (8 B per candidate) instead of from the raw scores (4 B per
// anchor x class, every band): room for 1/8 of the score slots, between 32 K and 256 K keys.
static int cap_for(int max_nms, long long slots) {
  int c = SORT_TILE;
  while (c < max_nms) c <<= 1;
  long long want = slots / 8;
  if (want > 262144) want = 262144;
  while (c < want) c <<= 1;
  return c;
}

size_t nms_workspace_bytes(int B, int nc, int A, int max_nms) {
  size_t hdr = ((size_t)B * sizeof(NmsHeader) + 255) / 256 * 256;
  size_t hist = (size_t)B * HIST_BINS * 4;
  size_t keys = (size_t)B * cap_for(max_nms, (long long)nc * A) * 8;
  return hdr + hist + keys;
}

// Where the class-score epilogues of the forward (conv_tc.cu, out_mode 3) append candidate keys when the
// caller runs yb_forward_nms + yb_nms_prefiltered instead of yb_forward + yb_nms.
int nms_sink_layout(void* ws, size_t ws_bytes, int B, int nc, int A, int max_nms, int** hdr,
                    unsigned long long** keys, int* cap) {
  static_assert(sizeof(NmsHeader) == 32, "conv_tc.cu indexes the header as 8 ints per image");
  if (B <= 0 || nc <= 0 || A <= 0 || max_nms <= 0 || ws_bytes < nms_workspace_bytes(B, nc, A, max_nms)) {
    set_error("yb_forward_nms: NMS workspace too small or bad sizes");
    return YB_ERR_ARG;
  }
  size_t hdr_bytes = ((size_t)B * sizeof(NmsHeader) + 255) / 256 * 256;
  *hdr = reinterpret_cast<int*>(ws);
  *keys = reinterpret_cast<unsigned long long*>((uint8_t*)ws + hdr_bytes + (size_t)B * HIST_BINS * 4);
  *cap = cap_for(max_nms, (long long)nc * A);
  return YB_OK;
}

// ws_clean: 0 = clear the headers first, 1 = headers known to be zero, 2 = the candidate lists were already
// filled by the forward's epilogues (skip the append pass)
int nms_run(const float* pred, int B, int nc, int A, float conf, double iou, int max_det, int max_nms,
            float max_wh, float* out, int* out_counts, void* ws, size_t ws_bytes, cudaStream_t st, int ws_clean) {
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
  a.cap = cap_for(max_nms, (long long)nc * A);
  a.first_band = FIRST_BAND;
  a.next_band = NEXT_BAND;
  if (const char* e = getenv("YB_NMS_FIRST_BAND")) a.first_band = std::max(32, std::min(SORT_TILE, atoi(e)));
  if (const char* e = getenv("YB_NMS_NEXT_BAND")) a.next_band = std::max(32, std::min(SORT_TILE, atoi(e)));
  a.max_wh = max_wh;
  size_t hdr_bytes = ((size_t)B * sizeof(NmsHeader) + 255) / 256 * 256;
  a.hdr = reinterpret_cast<NmsHeader*>(ws);
  a.ghist = reinterpret_cast<unsigned int*>((uint8_t*)ws + hdr_bytes);
  a.keys = reinterpret_cast<unsigned long long*>((uint8_t*)ws + hdr_bytes + (size_t)B * HIST_BINS * 4);
  a.out = out;
  a.out_counts = out_counts;
  // per-device function attribute (the caller has made the tensor's device current)
  static bool attr_dev[YB_MAX_DEVICES] = {false};
  int cur_dev = 0;
  YB_CUDA(cudaGetDevice(&cur_dev));
  bool& attr_set = attr_dev[cur_dev & (YB_MAX_DEVICES - 1)];
  if (!attr_set) {
    YB_CUDA(cudaFuncSetAttribute(nms_image_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(ImgSmem)));
    YB_CUDA(cudaFuncSetAttribute(nms_image_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(ImgSmem)));
    attr_set = true;
  }
  // the per-image kernel leaves the headers zeroed; a workspace last used by yb_nms / yb_nms_workspace_init
  // with the same batch needs no memset node in front of the append kernel
  if (!ws_clean) YB_CUDA(cudaMemsetAsync(ws, 0, hdr_bytes + (size_t)B * HIST_BINS * 4, st));
  long long total = (long long)nc * A;
  int gx = (int)std::min<long long>((total + 256 * 16 - 1) / (256 * 16), 1024);
  int vec4 = (A % 4 == 0) && ((reinterpret_cast<uintptr_t>(pred) & 15) == 0);
  if (ws_clean != 2) {
    YB_CUDA(launch_pdl(nms_append_kernel, dim3(gx, B), dim3(256), 0, st, a, vec4));
    count_launch();
  }
  // overflow images only (every other block exits at once: keep the grids small, ~2 K CTAs)
  const int ogx = std::max(8, std::min(std::min(gx, 128), 2048 / B));
  YB_CUDA(launch_pdl(nms_ovf_hist_kernel, dim3(ogx, B), dim3(256), 0, st, a));
  count_launch();
  YB_CUDA(launch_pdl(nms_ovf_select_kernel, dim3(ogx, B), dim3(256), 0, st, a));
  count_launch();
  if (B <= 160) YB_CUDA(launch_pdl(nms_image_kernel<1024>, dim3(B), dim3(1024), sizeof(ImgSmem), st, a));
  else YB_CUDA(launch_pdl(nms_image_kernel<512>, dim3(B), dim3(512), sizeof(ImgSmem), st, a));
  count_launch();
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

}  // namespace yb
