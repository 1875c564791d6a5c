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
#include <string.h>

#include "yb_internal.h"

namespace yb {

static constexpr int HIST_BINS = 2048;
static constexpr int SORT_TILE = 4096;

struct NmsHeader {  // per image, zeroed at the start of every call
  int cand_count;   // candidates found by the scan
  int sel_count;    // keys appended by the scan (may exceed capacity; clamp on read)
  int pad[6];
};

struct NmsArgs {
  const float* pred;
  int B, nc, A;
  float conf;
  double iou;
  int max_det, max_nms, cap;
  float max_wh;
  NmsHeader* hdr;
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

static constexpr int G_CHUNK = 256;
static constexpr int G_MAXDET = 1024;  // shared-memory capacity for kept boxes
static constexpr int FIRST_BAND = 1024;  // preferred size of the first band (most images finish inside it)

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
  BoxF live[G_CHUNK];      // candidates of the chunk, then (compacted in place) its survivors
  unsigned long long live_key[G_CHUNK];
  unsigned int mask[G_CHUNK][G_CHUNK / 32];
  int dead[G_CHUNK];
  int warp_cnt[G_CHUNK / 32];
  int K, m, cnt, walk_end;
  unsigned int walk_cum;
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
    const int iters = (s.n + IMG_T - 1) / IMG_T;
    int i = threadIdx.x;
    for (int it = 0; it < iters; it++, i += IMG_T) {
      const bool v = i < s.n;
      f(v, v ? s.keys[i] : ~0ull);
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

// Warp 0: longest run of bins [start, end) with sum <= limit.  Results in sm.walk_end / sm.walk_cum.
__device__ __forceinline__ void walk_bins(ImgSmem& sm, const unsigned int* h, int start, int nb, unsigned int limit) {
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  unsigned int cum = 0;
  int end = start;
  for (int b0 = start; b0 < nb; b0 += 32) {
    unsigned int incl = (b0 + lane < nb) ? h[b0 + lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const bool ok = (b0 + lane < nb) && (cum + incl <= limit);
    const int nok = __popc(__ballot_sync(0xffffffffu, ok));  // ok is a prefix of the lanes (incl is monotone)
    if (nok > 0) cum += __shfl_sync(0xffffffffu, incl, nok - 1);
    end = b0 + nok;
    if (nok < 32) break;
  }
  if (lane == 0) {
    sm.walk_end = end;
    sm.walk_cum = cum;
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
  src.complete = h.sel_count <= a.cap;
  src.keys = a.keys + (size_t)b * a.cap;
  src.n = min(h.sel_count, a.cap);
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
  __syncthreads();
  if (h.cand_count > 0) {
    scan_src<IMG_T>(src, [&](bool v, unsigned long long k) {
      if (v) atomicAdd(&sm.hist[bin_of(k)], 1u);
    });
  }
  __syncthreads();

  int cur_bin = 0;
  int consumed = 0;                 // candidates handed to the greedy step so far (sorted positions)
  unsigned long long lo_done = 0;   // every key below this has been consumed
  bool first = true;
  while (h.cand_count > 0 && cur_bin < HIST_BINS && consumed < a.max_nms && sm.K < a.max_det) {
    // ---- choose the band [lo_done, band_hi)
    walk_bins(sm, sm.hist, cur_bin, HIST_BINS, first ? FIRST_BAND : SORT_TILE);
    __syncthreads();
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
          walk_bins(sm, sm.hist2, 0, HIST_BINS, SORT_TILE);
          __syncthreads();
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
    const int n_band = min(sm.cnt, SORT_TILE);
    int P = 32;
    while (P < n_band) P <<= 1;
    for (int i = n_band + tid; i < P; i += IMG_T) sm.tile[i] = ~0ull;
    __syncthreads();
    for (int kk = 2; kk <= P; kk <<= 1) {
      for (int j = kk >> 1; j > 0; j >>= 1) {
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
    }
    // ---- greedy step over the sorted band, 256 candidates at a time
    const int n_use = min(n_band, a.max_nms - consumed);  // util.py:157: only the max_nms best
    for (int base = 0; base < n_use; base += G_CHUNK) {
      const int K = sm.K;
      if (K >= a.max_det) break;
      const int j = tid & (G_CHUNK - 1), part = tid >> 8;
      const int ci = base + j;
      if (tid < G_CHUNK) {
        sm.dead[j] = ci < n_use ? 0 : 1;
        if (ci < n_use) load_candidate(a, b, sm.tile[ci], sm.live[j], nullptr);
      }
      __syncthreads();
      if (ci < n_use) {
        const BoxF me = sm.live[j];
        for (int k = part; k < K; k += IMG_T / G_CHUNK) {
          if (suppresses(sm.kept[k], me, thr, thr_lo, thr_hi)) {
            sm.dead[j] = 1;
            break;
          }
        }
      }
      __syncthreads();
      // ordered compaction of the survivors (threads 0..255, in place: pos <= j)
      bool alive = false;
      BoxF me;
      unsigned int bal = 0;
      if (tid < G_CHUNK) {
        alive = !sm.dead[j];
        me = sm.live[j];
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
        }
        if (tid == 0) {
          int m = 0;
          for (int wi = 0; wi < G_CHUNK / 32; wi++) m += sm.warp_cnt[wi];
          sm.m = m;
        }
      }
      __syncthreads();
      const int m = sm.m;
      // suppression bitmask inside the chunk: mask[i][w] bit l <=> live[i] suppresses live[32w+l], 32w+l > i
      for (int task = warp; task < m * (G_CHUNK / 32); task += IMG_T / 32) {
        const int i = task / (G_CHUNK / 32), wcol = task - i * (G_CHUNK / 32);
        const int jj = wcol * 32 + lane;
        bool sup = false;
        if (jj > i && jj < m) sup = suppresses(sm.live[i], sm.live[jj], thr, thr_lo, thr_hi);
        const unsigned int bits = __ballot_sync(0xffffffffu, sup);
        if (lane == 0) sm.mask[i][wcol] = bits;
      }
      __syncthreads();
      if (warp == 0) {
        unsigned int remv = 0;  // lane w (< 8) holds removed-bits word w
        int Kc = K;
        for (int i = 0; i < m; i++) {
          const unsigned int word = __shfl_sync(0xffffffffu, remv, i >> 5);
          if (!((word >> (i & 31)) & 1u)) {
            if (lane == 0) {
              sm.kept[Kc] = sm.live[i];
              sm.kept_key[Kc] = sm.live_key[i];
            }
            Kc++;
            if (Kc >= a.max_det) break;
            if (lane < G_CHUNK / 32) remv |= sm.mask[i][lane];
          }
        }
        if (lane == 0) sm.K = Kc;
      }
      __syncthreads();
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
  }
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
static int cap_for(int max_nms) {
  int c = SORT_TILE;
  while (c < max_nms) c <<= 1;
  return c;
}

size_t nms_workspace_bytes(int B, int nc, int A, int max_nms) {
  (void)nc;
  (void)A;
  size_t hdr = ((size_t)B * sizeof(NmsHeader) + 255) / 256 * 256;
  size_t keys = (size_t)B * cap_for(max_nms) * 8;
  return hdr + keys;
}

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
  a.cap = cap_for(max_nms);
  a.max_wh = max_wh;
  size_t hdr_bytes = ((size_t)B * sizeof(NmsHeader) + 255) / 256 * 256;
  a.hdr = reinterpret_cast<NmsHeader*>(ws);
  a.keys = reinterpret_cast<unsigned long long*>((uint8_t*)ws + hdr_bytes);
  a.out = out;
  a.out_counts = out_counts;
  static bool attr_set = false;
  if (!attr_set) {
    YB_CUDA(cudaFuncSetAttribute(nms_image_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(ImgSmem)));
    YB_CUDA(cudaFuncSetAttribute(nms_image_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(ImgSmem)));
    attr_set = true;
  }
  // the per-image kernel leaves the headers zeroed; a workspace last used by yb_nms / yb_nms_workspace_init
  // with the same batch needs no memset node in front of the append kernel
  if (!ws_clean) YB_CUDA(cudaMemsetAsync(ws, 0, hdr_bytes, st));
  long long total = (long long)nc * A;
  int gx = (int)std::min<long long>((total + 256 * 16 - 1) / (256 * 16), 1024);
  int vec4 = (A % 4 == 0) && ((reinterpret_cast<uintptr_t>(pred) & 15) == 0);
  YB_CUDA(launch_pdl(nms_append_kernel, dim3(gx, B), dim3(256), 0, st, a, vec4));
  count_launch();
  if (B <= 160) YB_CUDA(launch_pdl(nms_image_kernel<1024>, dim3(B), dim3(1024), sizeof(ImgSmem), st, a));
  else YB_CUDA(launch_pdl(nms_image_kernel<512>, dim3(B), dim3(512), sizeof(ImgSmem), st, a));
  count_launch();
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

}  // namespace yb
