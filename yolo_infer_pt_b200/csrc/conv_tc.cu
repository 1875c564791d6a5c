// Dense convolution (1x1 / 3x3, stride 1 / 2, groups = 1) as an implicit GEMM on the sm_100a
// 5th-generation tensor cores.  Stands in for every `Conv.fuse_forward` / bare Conv2d call of the
// reference (nets/nn.py:38-39, 246, 252) together with the SiLU, residual add (nn.py:49,135,136),
// torch.cat (nn.py:63,80,94,148,205-208,257) and Upsample (nn.py:195) around it.
//
//   D[m, n] = sum_k A[m, k] * W[n, k]       m = flattened (image, oy, ox) output pixel
//                                           n = output channel
//                                           k = (tap, source slice, channel)
//
//   A : NHWC bf16 activations, four ways into shared memory (template parameter MODE):
//       MODE_ATMA   1x1/stride-1 over plain slices: 128x64 tiles by TMA, one tensor map per source of a concat;
//       MODE_PATCH  3x3/stride-1: the halo patch of a 16x8 pixel tile by one 4-D TMA box (zero-filled outside
//                   the image = padding); every tap is a UMMA descriptor shifted by the tap's pixel offset
//                   (the swizzle is a function of the smem address bits).  MODE_PATCH2: 16x16 super-tiles,
//                   two accumulators share each weight k-block;
//       MODE_DW     depthwise 3x3 + bias + SiLU computed by 8 extra warps from a TMA patch straight into the
//                   swizzled A stage of the 1x1 conv that follows (nn.py:248-251);
//       MODE_GATHER stride-2 and upsampled sources: 128 threads gather 16-byte channel granules with cp.async
//                   into the 128B-swizzled K-major layout.
//   W : packed bf16 [N_pad][K_pad], K-major, fetched with TMA (SWIZZLE_128B); resident in shared memory for
//       the whole kernel when the matrix fits next to the A ring.
//   D : fp32 accumulators in TMEM (128 lanes x BN columns), tcgen05.mma cta_group::1 kind::f16,
//       issued by one thread; tcgen05.commit releases smem stages and signals the epilogue.
//   Epilogue: tcgen05.ld -> bias + SiLU -> +residual -> bf16 staged in swizzled smem -> TMA store into the
//       channel slice of the consumer's buffer (or fp32 head outputs, below).
//
// The kernel is persistent: grid = SMs x resident CTAs, each CTA walks tiles t = blockIdx.x + i*grid.
// Warp roles (352 threads; 608 with the depthwise warps): warps 0-3 epilogue (TMEM lane = output row), warps
// 4-7 im2col producers or a second epilogue group, warp 8 TMEM allocator + MMA issuer, warp 9 TMA producer,
// warp 10 halo-patch producer / proxy-fence relay.  Each role executes griddepcontrol.wait right before its
// first access to activations (constants - bias, weights - are requested before it).  Two accumulator stages in
// TMEM (2 x BN columns) let the epilogue of tile i overlap the main loop of tile i+1; the smem
// stage ring runs continuously across tiles.
// Epilogue modes: bf16 NHWC slice (+residual) | fp32 head logits | fused DFL box decode
// (nets/nn.py:222-225,265-268 + make_anchors utils/util.py:85-96) | fused class sigmoid (nn.py:270),
// the last two writing the final (B, 4+nc, A) fp32 tensor directly; the class mode can also append the NMS
// candidate keys of its scores > conf (utils/util.py:130,147) to per-image lists (yb_forward_nms).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>

#include "yb_internal.h"

namespace yb {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
static constexpr int MAX_STAGES = 12;
static constexpr int C_GROUP_BYTES = BM * 128;    // 128 rows x 64 bf16 output staging sub-tile
static constexpr int MAX_PATCH_STAGES = 6;
static constexpr int NUM_BARS = 3 * MAX_STAGES + 4 + 2 * MAX_PATCH_STAGES;
static constexpr int PT_H = 16, PT_W = 8;          // PATCH mode output tile (pixels)
static constexpr int PP_H = PT_H + 2, PP_W = PT_W + 2;  // its input halo patch
enum { MODE_GATHER = 0, MODE_ATMA = 1, MODE_PATCH = 2, MODE_DW = 3, MODE_PATCH2 = 4 };
static constexpr int PP2_W = 2 * PT_W + 2;        // MODE_PATCH2: one 18 x 18 patch feeds two 16 x 8 tiles side by side

struct ConvParams {
  const act_t* src[4];   // fp16 or bf16 bits (template parameter F16)
  int src_cp[4];      // padded channels of the slice (multiple of 8): its K extent per tap
  int src_ld[4];      // row stride of the source buffer, elements
  int src_up[4];
  int nseg;
  int ksize, stride, pad;
  int Hin, Win, Hout, Wout;
  int M;              // B * Hout * Wout
  int K, num_kb;
  int per_tap;        // sum of src_cp (tap-aligned layers: padded to a multiple of 64)
  int tap_c;          // tap-aligned layers: channels >= tap_c of a tap are padding (never gathered)
  int seg_kb[4];      // a_tma: k-blocks per segment
  void* dst;
  int dst_ld;         // elements
  int dst_rows_per_img, dst_row_off, hw_out;
  int cout_store;     // channels actually stored (padded to 8 for bf16, 4 for fp32)
  uint8_t kb_kv[32];  // valid K16 steps of k-block kb (4 unless the block ends a source slice / the real K)
  int out_f32;
  const float* bias;
  const act_t* res;
  int res_ld;
  int act;
  int BN, stages, a_tma, tmem_cols;
  int n_tiles, total_tiles;
  uint32_t hw_mul, hw_shr, w_mul, w_shr;  // magic numbers: division by hw_out and by Wout
  // 2-D spatial tiles (8 x 16 output pixels) when Hout % 8 == 0 and Wout % 16 == 0
  int tile2d, tiles_x, tiles_per_img;
  uint32_t tpi_mul, tpi_shr, tx_mul, tx_shr;
  uint32_t pt_mul, pt_shr;   // division by per_tap
  // 3x3 / stride-1 convs fed from TMA halo patches (mode PATCH): an 18 x 10 pixel patch of `cblk`
  // channels per 16 x 8 output tile; each tap is a shifted UMMA descriptor into the patch
  int patch, cblk, ncb, a_layout;
  int s2d;              // stride-2 3x3 over a space-to-depth source: 1 = 16 channels (the four parity blocks are one 64-channel patch), 2 = 64 channels (one patch per parity block)
  int c_s2d;            // the destination buffer is stored space-to-depth (8 x 16 pixel tiles through a 5-D TMA map)
  int patch_creal;      // real input channels (tap-aligned layers: K steps past them are zero padding and are skipped)
  int pair;             // MODE_PATCH2: 16 x 16 pixel super-tiles = two accumulators sharing every weight k-block
  int patch_tx_bytes, patch_stage_bytes, patch_stages;
  int a_region_bytes;   // shared memory of the A ring (stages x 16 KB, or the patch ring)
  int b_resident;       // weights stay in shared memory for the whole kernel: slot kb, loaded during the first tile
  int b_slots;          // weight slots in shared memory (num_kb when resident, else stages)
  int c_bufs;           // output staging buffers (2: the epilogue never waits for the previous tile's TMA store)
  int alt_epilogue;     // the two epilogue warp groups take alternate tiles (when the kernel has two)
  // MODE_DW: a depthwise 3x3 (+bias, SiLU) fused in front of this 1x1 conv.  Its input comes in as
  // TMA halo patches (unswizzled), warps 4-7 compute the depthwise output of the tile straight into
  // the swizzled A stage; the intermediate tensor never exists in global memory.
  int dw;
  const float* dw_w;    // [9][dw_cp] fp32 taps + [dw_cp] bias
  int dw_C, dw_cp, dw_act;
  int dw_patch_bytes;   // shared memory of the patch ring (the A ring follows it)
  int wait_sleep_ns;    // sleep between polls of the off-critical-path waits (YB_WAIT_SLEEP_NS, 0 = tight try_wait)
  int dw_dbg;           // YB_DW_DBG bit 0: fp32 depthwise math, bit 1: fp32 SiLU in the epilogue (A/B switches)
  // fused head decode (out_mode 2 / 3): dst is the (B, 4+nc, A) fp32 output tensor
  int out_mode;       // 0 bf16 slice, 1 fp32 logits, 2 DFL box decode, 3 class sigmoid
  int A_total, nc;
  // out_mode 3 with a candidate sink: every score > nms_conf also becomes a key of the image's NMS list
  // (what nms_append_kernel would find by re-reading the scores)
  int* nms_hdr;                   // [B] NmsHeader (8 ints): [0] cand_count, [1] sel_count
  unsigned long long* nms_keys;   // [B][nms_cap]
  int nms_cap;
  float nms_conf;
  float lvl_stride;
  // naive path only
  const act_t* w;
  int K_pad, cout;
  int src_c[4];       // real channels
  int seg_kpad[4];
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait suspends the thread in hardware for a bounded time, so this loop polls rarely.
// -DYB_BOUNDED_WAIT turns a protocol bug into a trap (a CUDA error) instead of a hang.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#ifdef YB_BOUNDED_WAIT
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > (1u << 27)) __trap();
  }
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
#endif
}
// Wait for roles that are NOT on the critical path (epilogue warps waiting for an accumulator, the TMA / patch
// producers waiting for a free stage): poll with test_wait and sleep between polls.  A warp parked in try_wait is
// woken by every mbarrier event of the CTA and re-executes the check (ncu: the try_wait loop was 35 % of all warp
// instructions of a fused-depthwise launch, issued from the same schedulers the math warps need); a sleeping
// warp issues nothing.  `ns` bounds the extra latency; sleep_ns == 0 falls back to the tight wait.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t sleep_ns) {
  if (sleep_ns == 0) {
    mbar_wait(bar, parity);
    return;
  }
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    asm volatile("nanosleep.u32 %0;" ::"r"(sleep_ns));
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, "
      "%3}], [%4];" ::"r"(dst),
      "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// q = n / d for n < 2^31 with host-computed (mul, shr): mul = ceil(2^(31+ceil_log2 d) / d)
__device__ __forceinline__ int fast_div(int n, uint32_t mul, uint32_t shr, int d) {
  return d == 1 ? n : (int)(__umulhi((uint32_t)n, mul) >> shr);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, "
      "%3, %4, %5}], [%6];" ::"r"(dst),
      "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
// Tile -> position.  Linear tiles: 128 consecutive rows of the flattened (image, y, x) index.
// 2-D tiles: an 8 (y) x 16 (x) patch of one image; tile row r = ly * 16 + lx.
struct TilePos {
  int m0;            // linear: first row
  int n, oy0, ox0;   // 2-D: image and patch origin
};
template <bool T2D, int TW, int TH = BM / TW>
__device__ __forceinline__ TilePos tile_pos(const ConvParams& P, int mt) {
  TilePos t;
  t.m0 = mt * BM;
  t.n = t.oy0 = t.ox0 = 0;
  if (T2D) {
    t.n = fast_div(mt, P.tpi_mul, P.tpi_shr, P.tiles_per_img);
    int r = mt - t.n * P.tiles_per_img;
    int ty = fast_div(r, P.tx_mul, P.tx_shr, P.tiles_x);
    t.oy0 = ty * TH;
    t.ox0 = (r - ty * P.tiles_x) * TW;
  }
  return t;
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  // K-major, SWIZZLE_128B canonical layout: 8-row x 128B atoms, SBO = 1024 B, LBO = 1 (unused),
  // descriptor version 1 (sm_100), layout type 2.
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// 16-byte global->shared async copy; src_bytes = 0 zero-fills the destination (out-of-image taps)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_pending(int n) {
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
  }
}
// SiLU(x) = x * sigmoid(x) = h + h * tanh(h), h = x/2: one MUFU op (tanh.approx) instead of two
// (ex2 + rcp).  The epilogue is MUFU-throughput bound on the small-channel layers.
__device__ __forceinline__ float silu_f(float x) {
  float h = 0.5f * x, t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

__device__ __forceinline__ float silu_half(float h) {  // SiLU(2h)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}


// SiLU(2h) on a pair of fp16 values with ONE MUFU op (tanh.approx.f16x2) and one HFMA2: h + h * tanh(h).
// Used on the class branch only (fused depthwise layers), whose 1e-2 score tolerance has a 10x margin; its
// output is rounded to fp16 storage anyway.
__device__ __forceinline__ uint32_t silu_half_h2(uint32_t h2) {
  uint32_t t, o;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(h2));
  asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(o) : "r"(h2), "r"(t));
  return o;
}
__device__ __forceinline__ uint32_t hfma2_u32(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// All MMAs of one halo patch (one channel block, 9 taps) against weights resident in shared memory,
// fully unrolled: tap offsets and K steps are compile-time constants, so the single issuing thread
// spends two integer adds per tcgen05.mma instead of divisions and descriptor assembly (with
// N <= 64 an MMA retires in ~48 cycles; a longer issue sequence is the bottleneck).
// a_lo / b_lo: low descriptor words (address >> 4) of the patch and of weight slot `kb0`; bstep =
// slot stride >> 4; first: the accumulator is overwritten by the first MMA.
template <int CBLK, bool PAIR, bool FULLK>
__device__ __forceinline__ void issue_patch_steady(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                   uint32_t b_hi, uint32_t bstep, uint32_t idesc, bool first,
                                                   uint32_t bn, int kvalid) {
  constexpr int ROWB = CBLK * 2;
  constexpr int PP_W = PAIR ? PP2_W : yb::PP_W;   // patch row pitch in pixels (shadows the single-tile constant)
  constexpr int NMMA = CBLK == 8 ? 5 : 9 * CBLK / 16;
  // Low descriptor words: (shared-memory address >> 4) in bits 0-13, LBO in bits 16-29.  Every operand lies
  // inside the CTA's shared-memory window (< 256 KB), so base + offset never carries out of the 14 address bits:
  // one 32-bit add with a compile-time constant (offset and, for 8-channel patches, the tap-pair LBO) per MMA
  // for A, and a running weight-slot base for B.
  const uint32_t a0 = a_lo | (1u << 16), b0 = b_lo | (1u << 16);
  uint32_t bslot = b0;
#pragma unroll
  for (int i = 0; i < NMMA; i++) {
    uint32_t a_add;   // compile-time: (offset >> 4) + ((lbo - 1) << 16)
    int kslot, k;
    if (CBLK == 8) {
      const int t0 = 2 * i, t1 = i < 4 ? 2 * i + 1 : 2 * i;
      const int o0 = ((t0 / 3) * PP_W + t0 % 3) * 16, o1 = ((t1 / 3) * PP_W + t1 % 3) * 16;
      const uint32_t lbo = i < 4 ? (uint32_t)((o1 - o0) >> 4) : 1u;
      a_add = (uint32_t)(o0 >> 4) + ((lbo - 1u) << 16);
      kslot = i / 4;
      k = i % 4;
    } else {
      const int kg = i * 16, tap = kg / CBLK, cin = kg % CBLK;
      a_add = (uint32_t)((((tap / 3) * PP_W + tap % 3) * ROWB + cin * 2) >> 4);
      kslot = kg / 64;
      k = (kg % 64) / 16;
    }
    if (i > 0 && k == 0 && kslot > 0) bslot += bstep;       // next weight k-block (compile-time condition)
    if (!FULLK && CBLK == 64 && k >= kvalid) continue;      // zero-padded channels of a tap-aligned layer
    const uint64_t da = ((uint64_t)a_hi << 32) | (uint64_t)(a0 + a_add);
    const uint64_t db = ((uint64_t)b_hi << 32) | (uint64_t)(bslot + 2u * (uint32_t)k);
    umma_bf16(d_tmem, da, db, idesc, (uint32_t)(!(first && i == 0)));
    if (PAIR)   // right half: same weights, patch shifted by 8 pixels, second accumulator
      umma_bf16(d_tmem + bn, da + (uint64_t)((PT_W * ROWB) >> 4), db, idesc, (uint32_t)(!(first && i == 0)));
  }
}

// ------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------
static constexpr int NUM_THREADS = 352;
static constexpr int EPI_WARPS = 4;      // warps 0-3
static constexpr int PROD_WARP0 = 4;     // warps 4-7
static constexpr int MMA_WARP = 8;
static constexpr int TMA_WARP = 9;
static constexpr int RELAY_WARP = 10;    // proxy-fence relay between the im2col warps and the MMA warp
static constexpr int DW_WARP0 = 11;      // MODE_DW: warps 11-18 compute the depthwise conv (both epilogue groups stay)
static constexpr int DW_THREADS = 256;
static constexpr int NUM_THREADS_DW = (DW_WARP0 + DW_THREADS / 32) * 32;

// Specialised at compile time on the A-operand path (TMA tiles vs im2col gather), the tile shape
// (8x16 spatial patch vs 128 flattened rows) and the epilogue family (bf16 slice vs head modes):
// every instantiation carries only the code of its own roles, which keeps it inside the
// instruction cache (11 warps run disjoint code).
template <int MODE, bool T2D, bool HEAD, bool F16>
__global__ void __launch_bounds__(MODE == MODE_DW ? NUM_THREADS_DW : NUM_THREADS, MODE == MODE_PATCH ? 3 : MODE == MODE_DW ? 1 : 2)  // (PATCH2: 2)
    conv_gemm_tcgen05_kernel(const ConvParams P, const __grid_constant__ CUtensorMap tmap_b,
                             const __grid_constant__ CUtensorMap tmap_a0,
                             const __grid_constant__ CUtensorMap tmap_a1,
                             const __grid_constant__ CUtensorMap tmap_a2,
                             const __grid_constant__ CUtensorMap tmap_a3,
                             const __grid_constant__ CUtensorMap tmap_c) {
  using A16 = Act16<F16>;   // 16-bit storage / operand type of activations and weights
  constexpr bool DW = MODE == MODE_DW;
  constexpr bool A_TMA = MODE != MODE_GATHER;   // warps 4-7 are not im2col producers: they join the epilogue
  constexpr bool PAIR = MODE == MODE_PATCH2;        // two 16 x 8 tiles per iteration (one 16 x 16 super-tile)
  constexpr bool PATCH = MODE == MODE_PATCH || PAIR;
  constexpr bool PGEO = PATCH || DW;               // 16 x 8 pixel tiles
  constexpr int TW = PAIR ? 16 : PGEO ? PT_W : 16; // 2-D tile width used to place the tile in the image
  constexpr int TH = PAIR ? 16 : BM / TW;
  constexpr int EW = PGEO ? PT_W : 16;             // width of one accumulator's pixel block (row r = ly * EW + lx)
  constexpr int EWS = PGEO ? 3 : 4;
  constexpr int PPW = PAIR ? PP2_W : PP_W;         // patch row pitch, pixels
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int S = P.stages;
  const int BN = P.BN;
  const uint32_t b_stage_bytes = (uint32_t)BN * 128u;
  const uint32_t a_base = base;
  const uint32_t b_base = base + (uint32_t)P.a_region_bytes;
  // output staging for the TMA store: one 128 x 64 bf16 swizzled sub-tile (16 KB) per 64 channels
  const uint32_t c_groups = (uint32_t)(BN + 63) / 64;
  const uint32_t c_base = b_base + (uint32_t)P.b_slots * b_stage_bytes;
  uint8_t* tail = smem + (size_t)P.a_region_bytes + (size_t)P.b_slots * b_stage_bytes +
                  (size_t)c_groups * C_GROUP_BYTES * (size_t)P.c_bufs * (PAIR ? 2 : 1);
  const bool RES = P.b_resident != 0;
  const uint32_t ring_base = a_base + (DW ? (uint32_t)P.dw_patch_bytes : 0u);   // A stage ring
  // Two epilogue warp groups exist when no im2col producers are needed.  With one N tile and two
  // staging buffers they take alternate tiles (group g drains TMEM stage g into staging buffer g, its
  // own named barrier and TMA stores), so two epilogue chains are in flight; otherwise they split
  // the columns of every tile.
  const bool alt_epi = A_TMA && P.n_tiles == 1 && (HEAD || P.c_bufs == 2) && P.alt_epilogue;
  // barriers: full[MAX_STAGES], empty[MAX_STAGES], tmem_full[2], tmem_empty[2], gathered[MAX_STAGES],
  //           patch_full[MAX_PATCH_STAGES], patch_empty[MAX_PATCH_STAGES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + NUM_BARS * 8);
  float* bias_s = reinterpret_cast<float*>(tail + NUM_BARS * 8 + 16);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar0 + 8u * (2 * MAX_STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar0 + 8u * (2 * MAX_STAGES + 2 + a); };
  auto gathered_bar = [&](int s) { return bar0 + 8u * (2 * MAX_STAGES + 4 + s); };
  auto patch_full_bar = [&](int s) { return bar0 + 8u * (3 * MAX_STAGES + 4 + s); };
  auto patch_empty_bar = [&](int s) { return bar0 + 8u * (3 * MAX_STAGES + 4 + MAX_PATCH_STAGES + s); };

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int num_kb = P.num_kb;

  if (tid == 0) {
    for (int s = 0; s < S; s++) {
      // TMA expect_tx (+ the relay warp on im2col layers, + one arrival per depthwise warp in MODE_DW: every
      // mbarrier arrival wakes the warps sleeping in try_wait, so arrivals are per warp, not per thread)
      mbar_init(full_bar(s), MODE == MODE_GATHER ? 2u : DW ? (uint32_t)(DW_THREADS / 32) + 1u : 1u);
      mbar_init(empty_bar(s), 1u);
      mbar_init(gathered_bar(s), 128u);            // one cp.async-completion arrive per im2col thread
    }
    for (int s = 0; s < MAX_PATCH_STAGES; s++) {
      mbar_init(patch_full_bar(s), 1u);
      mbar_init(patch_empty_bar(s), DW ? (uint32_t)(DW_THREADS / 32) : 1u);   // one arrival per depthwise warp
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tmem_full_bar(a), 1u);
      mbar_init(tmem_empty_bar(a), (A_TMA && !alt_epi) ? 8u : 4u);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (PATCH && P.cblk == 8 && tid < P.patch_stages) {
    // 8-channel patches pair two taps per K=16 MMA; the unpaired ninth tap reads 16 bytes past the
    // patch (against zero weights), which must therefore hold finite data
    *reinterpret_cast<uint4*>(smem + (size_t)tid * P.patch_stage_bytes + P.patch_tx_bytes) = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // prologue done (barriers, TMEM): let the next kernel start its own.  Each role waits for the producer
  // kernel of our inputs (griddepcontrol.wait) right before its first access to activations; bias and
  // weights are constants of the plan and are fetched before that wait, under the previous kernel's tail
  pdl_prologue_done();

  // With A fed by TMA the im2col warps have nothing to gather: they join the epilogue as a second
  // group that converts the odd 16-column chunks of the accumulator (same TMEM lane quadrants).
  if (warp < EPI_WARPS || (A_TMA && warp < MMA_WARP)) {
    // ============================ epilogue =================================================
    const int grp = warp >> 2;                  // 0: warps 0-3, 1: warps 4-7
    const int ngrp = A_TMA ? 2 : 1;
    const int etid = tid & 127;                 // row of the tile owned by this thread
    const int qwarp = warp & 3;                 // TMEM lane quadrant
    const int cstep = (alt_epi || (HEAD && P.out_mode == 2)) ? 16 : 16 * ngrp;   // DFL decode needs all 4 sides in one thread
    const int cfirst = alt_epi ? 0 : (HEAD && P.out_mode == 2) ? (grp ? BN : 0) : 16 * grp;
    const bool leader = alt_epi ? etid == 0 : tid == 0;
    auto epi_sync = [&]() {
      if (alt_epi) {
        if (grp) asm volatile("bar.sync 2, 128;" ::: "memory");
        else asm volatile("bar.sync 1, 128;" ::: "memory");
      } else if (ngrp == 2) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
      } else {
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    };
    int ti = 0;
    // One N tile: the bias vector is loaded once.  With two staging buffers the tile loop then needs
    // a single CTA-level barrier per tile (before the TMA store), and never waits on the previous
    // tile's store: that one is drained (wait_group.read) a full tile later.
    const bool bias_once = P.n_tiles == 1;
    const bool fast_flow = !alt_epi && bias_once && (HEAD || P.c_bufs == 2);
    // SiLU(x) = h + h*tanh(h) with h = x/2: the 1/2 is folded into the accumulator FMA (h = acc*0.5 +
    // bias*0.5, bit-identical to (acc + bias)*0.5), so a SiLU element costs FFMA + MUFU + FFMA
    if (bias_once) {
      for (int i = tid; i < BN; i += 128 * ngrp) bias_s[i] = (P.act ? 0.5f : 1.f) * __ldg(P.bias + i);
      if (ngrp == 2) asm volatile("bar.sync 1, 256;" ::: "memory");
      else asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    pdl_wait();
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ti++) {
      if (alt_epi && (ti & 1) != grp) continue;
      const int mt = P.n_tiles == 1 ? tile : tile / P.n_tiles;
      const TilePos tp = tile_pos<T2D, TW, TH>(P, mt);
      const int m0 = tp.m0;
      const int n0 = (tile - mt * P.n_tiles) * BN;
      const int acc = ti & 1;
      for (int half = 0; half < (PAIR ? 2 : 1); half++) {   // PATCH2: left / right 16 x 8 block of the super-tile
      int m = m0 + etid;
      bool row_ok = m < P.M;
      int n_img = 0, r = 0;
      if (T2D) {
        n_img = tp.n;
        const int oy = tp.oy0 + (etid >> EWS), ox = tp.ox0 + PT_W * half + (etid & (EW - 1));
        r = oy * P.Wout + ox;
        m = n_img * P.hw_out + r;
        row_ok = !PGEO || (oy < P.Hout && ox < P.Wout);   // 16 x 8 tiles may hang over the image edge
      } else if (row_ok) {
        n_img = fast_div(m, P.hw_mul, P.hw_shr, P.hw_out);
        r = m - n_img * P.hw_out;
      }
      const size_t drow = (size_t)n_img * P.dst_rows_per_img + P.dst_row_off + r;
      const act_t* resp = (!HEAD && P.res && row_ok) ? P.res + (size_t)m * P.res_ld : nullptr;
      // residual operand: prefetched two chunks ahead so its global-load latency hides behind the
      // wait for the accumulator and the math of the previous chunks
      uint4 ra0, ra1, rb0, rb1;
      ra0 = ra1 = rb0 = rb1 = make_uint4(0u, 0u, 0u, 0u);
      auto res_fetch = [&](int c0, uint4& r0, uint4& r1) {
        if (resp && c0 < BN) {
          const int nb = n0 + c0;
          if (nb < P.cout_store) r0 = __ldg(reinterpret_cast<const uint4*>(resp + nb));
          if (nb + 8 < P.cout_store) r1 = __ldg(reinterpret_cast<const uint4*>(resp + nb + 8));
        }
      };
      res_fetch(cfirst, ra0, ra1);
      res_fetch(cfirst + cstep, rb0, rb1);
      if (half == 0) {
      mbar_wait_sleep(tmem_full_bar(acc), (uint32_t)(ti >> 1) & 1u, (uint32_t)P.wait_sleep_ns);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      if (half == 0 && !fast_flow && !(alt_epi && HEAD)) {
        // the previous tile's TMA stores (this group's, when the groups alternate: issued a whole
        // tile ago) must have finished reading the staging buffer
        if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (!bias_once)
          for (int i = tid; i < BN; i += 128 * ngrp) bias_s[i] = (P.act ? 0.5f : 1.f) * __ldg(P.bias + n0 + i);
        epi_sync();
      }
      const uint32_t c_buf = c_base + (((fast_flow && (ti & 1)) || (alt_epi && grp) || (PAIR && half)) ? c_groups * C_GROUP_BYTES : 0u);
      const uint32_t t_row = tmem_base + ((uint32_t)(qwarp * 32) << 16) + (uint32_t)(acc * (PAIR ? 2 * BN : BN) + half * BN);
      float dist[4];
      // (measured: keeping the next chunk's tcgen05.ld in flight across the conversion of this one - issue after
      // the fp32 results exist, reuse the registers - made the halo-patch and residual variants 5-10 % slower and
      // nothing faster: the epilogue is not what these kernels wait for, and the longer live ranges cost scheduling
      // freedom.  One load + wait per 16-column chunk it stays.)
      auto do_chunk = [&](int c0, const uint4& rv0, const uint4& rv1, auto with_res) {
        constexpr bool WITH_RES = decltype(with_res)::value;
        uint32_t v[16];
        tmem_ld16(t_row + (uint32_t)c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int nb = n0 + c0;
        if (nb >= P.cout_store || (!row_ok && HEAD && P.out_mode != 3)) return;   // class mode keeps the warp whole (shuffles)
        float f[16];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * q);   // broadcast LDS.128
          const float bq[4] = {b4.x, b4.y, b4.z, b4.w};
          if (P.act) {   // uniform
            // (fused-depthwise class branch with fp16 storage: SiLU runs on packed halves at the store below)
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const float h = fmaf(__uint_as_float(v[4 * q + j]), 0.5f, bq[j]);
              f[4 * q + j] = (DW && F16 && !(P.dw_dbg & 2)) ? h : silu_half(h);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; j++) f[4 * q + j] = __uint_as_float(v[4 * q + j]) + bq[j];
          }
        }
        if (!HEAD) {
          // bf16 tile staged in shared memory in the TMA SWIZZLE_128B layout (16-byte chunk index
          // XOR row%8 inside each 128-byte row), then written with one TMA store per 64 channels
          const uint32_t srow = c_buf + (uint32_t)(c0 >> 6) * C_GROUP_BYTES + (uint32_t)etid * 128u;
#pragma unroll
          for (int h = 0; h < 2; h++) {
            if (nb + 8 * h < P.cout_store) {
              if (WITH_RES && resp) {
                const uint4 rv = h ? rv1 : rv0;
                const uint32_t r2[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                  float2 rf = A16::unpack2(r2[j]);
                  f[8 * h + 2 * j] += rf.x;
                  f[8 * h + 2 * j + 1] += rf.y;
                }
              }
              const uint32_t chunk = (uint32_t)(((c0 & 63) >> 3) + h);
              const uint32_t addr = srow + ((chunk ^ (uint32_t)(etid & 7)) << 4);
              uint32_t pk[4];
#pragma unroll
              for (int j = 0; j < 4; j++) {
                pk[j] = A16::pack2(f[8 * h + 2 * j], f[8 * h + 2 * j + 1]);
                if (DW && F16 && P.act && !(P.dw_dbg & 2)) pk[j] = silu_half_h2(pk[j]);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                           "r"(pk[3])
                           : "memory");
            }
          }
        } else if (P.out_mode == 1) {
          float* dp = reinterpret_cast<float*>(P.dst) + drow * (size_t)P.dst_ld + nb;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            if (nb + 4 * q < P.cout_store)
              *reinterpret_cast<float4*>(dp + 4 * q) =
                  make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
          }
        } else if (P.out_mode == 2) {
          // one 16-column chunk = the 16 DFL bins of one box side: softmax expectation
          float mx = f[0];
#pragma unroll
          for (int j = 1; j < 16; j++) mx = fmaxf(mx, f[j]);
          float se = 0.f, sw = 0.f;
          const float mxl = mx * -1.4426950408889634f;
#pragma unroll
          for (int j = 0; j < 16; j++) {
            float e;   // exp(f - mx) as one FFMA + MUFU.EX2 (ncu: __expf's range handling made this line a third of the kernel)
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(f[j], 1.4426950408889634f, mxl)));
            se += e;
            sw = fmaf((float)j, e, sw);
          }
          float rs;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(se));   // se >= 1 (the maximum contributes exp(0))
          dist[(c0 >> 4) & 3] = sw * rs;
        } else {
          // class scores: plane-major fp32 stores, one anchor per lane -> 128 B per warp store
          float* ob = reinterpret_cast<float*>(P.dst) + ((size_t)n_img * (4 + P.nc) + 4 + nb) * P.A_total +
                      P.dst_row_off + r;
          // (ncu on head.cls.0.4: this loop was 68 % of the kernel's instructions at ~25 per score - range checks of
          // __expf / __fdividef, a 64-bit multiply per store address, a branch per class.  Now FMUL + EX2 + FADD + RCP,
          // a running plane pointer and predicated stores; the candidate count only when a sink is attached.)
          int cnt = 0;
          const bool sink = P.nms_keys != nullptr;   // uniform
          float* o = ob;
          if (nb + 16 <= P.nc) {   // uniform: a whole chunk of classes - no bound check per class, one store predicate
#pragma unroll
            for (int j = 0; j < 16; j++) {
              float t, sg;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(f[j] * -1.4426950408889634f));
              asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.f + t));   // 1 / (1 + exp(-x)): inf -> 0, 0 -> 1
              f[j] = sg;
              if (row_ok) *o = sg;
              o += P.A_total;
            }
            if (sink) {
#pragma unroll
              for (int j = 0; j < 16; j++) cnt += f[j] > P.nms_conf ? 1 : 0;
              if (!row_ok) cnt = 0;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j++) {
              float t, sg;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(f[j] * -1.4426950408889634f));
              asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.f + t));
              f[j] = sg;
              if (row_ok && nb + j < P.nc) {
                *o = sg;
                cnt += sg > P.nms_conf ? 1 : 0;
              }
              o += P.A_total;
            }
          }
          if (P.nms_keys) {
            // warp-aggregated append (the whole warp is here: rows outside the tensor carry cnt = 0): one
            // scan + one pair of atomics per image the warp's 32 anchors belong to (one or two, more only on
            // feature maps smaller than a warp)
            unsigned rem = __ballot_sync(0xffffffffu, cnt > 0);
            if (rem) {
              const int lane_ = (int)(tid & 31);
              int slot = 0;
              while (rem) {
                const int leader = __ffs((int)rem) - 1;
                const int img = __shfl_sync(0xffffffffu, n_img, leader);
                const bool in = n_img == img;
                int incl = in ? cnt : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                  const int tt = __shfl_up_sync(0xffffffffu, incl, o);
                  if (lane_ >= o) incl += tt;
                }
                const int tot = __shfl_sync(0xffffffffu, incl, 31);
                int base = 0;
                if (lane_ == leader) {
                  base = atomicAdd(P.nms_hdr + 8 * img + 1, tot);   // sel_count: slots handed out
                  atomicAdd(P.nms_hdr + 8 * img, tot);              // cand_count
                }
                base = __shfl_sync(0xffffffffu, base, leader);
                if (in) slot = base + incl - cnt;
                rem &= ~__ballot_sync(0xffffffffu, in);
              }
              if (cnt) {
                unsigned long long* kl = P.nms_keys + (size_t)n_img * P.nms_cap;
                const unsigned int anchor = (unsigned int)(P.dst_row_off + r);
#pragma unroll
                for (int j = 0; j < 16; j++) {
                  if (nb + j < P.nc && f[j] > P.nms_conf) {
                    if (slot < P.nms_cap) {
                      const unsigned int u = __float_as_uint(f[j]);
                      const unsigned int ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                      kl[slot] = ((unsigned long long)(~ord) << 32) | (anchor * (unsigned int)P.nc + (unsigned int)(nb + j));
                    }
                    slot++;
                  }
                }
              }
            }
          }
        }
      };
      // two copies of the chunk body; a kernel only ever runs one of them.  The copy without a residual
      // carries no prefetch registers and no rotation moves.
      if ((DW && (P.dw_dbg & 8)) || (!HEAD && (P.dw_dbg & 16))) {
        // (ablation switch: epilogue without TMEM loads / math / staging)
      } else if (!HEAD && P.res != nullptr) {
        for (int c0 = cfirst; c0 < BN; c0 += cstep) {
          do_chunk(c0, ra0, ra1, std::true_type{});
          ra0 = rb0;
          ra1 = rb1;
          res_fetch(c0 + 2 * cstep, rb0, rb1);
        }
      } else {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int c0 = cfirst; c0 < BN; c0 += cstep) do_chunk(c0, z, z, std::false_type{});
      }
      if (HEAD && P.out_mode == 2 && row_ok && (alt_epi || grp == 0)) {
        const int y = fast_div(r, P.w_mul, P.w_shr, P.Wout), x = r - y * P.Wout;
        const float ax = (float)x + 0.5f, ay = (float)y + 0.5f, st = P.lvl_stride;
        const float x1 = ax - dist[0], y1 = ay - dist[1], x2 = ax + dist[2], y2 = ay + dist[3];
        float* ob = reinterpret_cast<float*>(P.dst) + (size_t)n_img * (4 + P.nc) * P.A_total + P.dst_row_off + r;
        ob[0] = (x1 + x2) / 2.f * st;
        ob[(size_t)P.A_total] = (y1 + y2) / 2.f * st;
        ob[(size_t)2 * P.A_total] = (x2 - x1) * st;
        ob[(size_t)3 * P.A_total] = (y2 - y1) * st;
      }
      }   // half
      // accumulator stage drained: hand it back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
      if (!HEAD) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // fast flow: the store of tile t-1 (other buffer) has had this whole tile to finish reading;
        // after the barrier below every thread may overwrite that buffer for tile t+1
        if (fast_flow && leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        epi_sync();
        if (leader) {
          const uint32_t c_buf = c_base + (((fast_flow && (ti & 1)) || (alt_epi && grp)) ? c_groups * C_GROUP_BYTES : 0u);
          for (int half = 0; half < (PAIR ? 2 : 1); half++)
          for (uint32_t g = 0; g < c_groups; g++) {
            const int cg0 = n0 + (int)g * 64;
            if (cg0 < P.cout_store) {
              if (T2D && !PGEO && P.c_s2d) {
                // space-to-depth destination: the 8 x 16 pixel tile is the box {64, 2, 8, 2, 4} of {c, x&1, x/2, y&1, n*H/2 + y/2}
                asm volatile(
                    "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(
                        (uint64_t)&tmap_c),
                    "r"(cg0), "r"(0), "r"(tp.ox0 >> 1), "r"(0), "r"(tp.n * (P.Hout >> 1) + (tp.oy0 >> 1)),
                    "r"(c_buf + g * C_GROUP_BYTES)
                    : "memory");
              } else if (T2D) {
                asm volatile(
                    "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(
                        (uint64_t)&tmap_c),
                    "r"(cg0), "r"(tp.ox0 + PT_W * half), "r"(tp.oy0), "r"(tp.n),
                    "r"(c_buf + (uint32_t)half * c_groups * C_GROUP_BYTES + g * C_GROUP_BYTES)
                    : "memory");
              } else {
                asm volatile(
                    "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                        (uint64_t)&tmap_c),
                    "r"(cg0), "r"(m0), "r"(c_buf + g * C_GROUP_BYTES)
                    : "memory");
              }
            }
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  } else if (warp < MMA_WARP) {
    // ============================ im2col producer ==========================================
    pdl_wait();
    {
      const int ptid = tid - PROD_WARP0 * 32;
      const int g = ptid & 7;       // 16-byte granule (8 channels) inside the 128-byte K row
      const int rbase = ptid >> 3;  // rows rbase + 16*i
      const uint32_t sw_off = (uint32_t)((g ^ (rbase & 7)) << 4);
      // cp.async gathers need no staging registers and no wait in the producer: each thread's
      // copies of a k-block arrive on the stage's full barrier by themselves when they land
      // (cp.async.mbarrier.arrive.noinc), so a thread runs ahead as far as free stages allow.  The
      // generic-proxy writes are made visible to the tensor core's async-proxy reads by the MMA
      // thread's fence.proxy.async after it acquires the barrier.
      // Address arithmetic is hoisted: per tile each row keeps the pixel index of its tap (0,0) and
      // a 9-bit in-bounds mask; a k-block then costs one add + one wide multiply per row.
      int stage = 0;
      uint32_t phase = 0;
      const int ksize = P.ksize, Win = P.Win, Hin = P.Hin, K = P.K, per_tap = P.per_tap, nseg = P.nseg;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const TilePos tp = tile_pos<T2D, TW, TH>(P, P.n_tiles == 1 ? tile : tile / P.n_tiles);
        if (T2D) {
          // ---- fast path: 8 x 16 spatial tile.  This thread's 8 rows are the 8 image rows of one
          // tile column (ly = i, lx = rbase), so x validity is shared and y validity per tap row is
          // a byte over i: only the first / last row of the patch can fall outside the image.
          const int x0 = (tp.ox0 + rbase) * P.stride - P.pad;
          const int y0 = tp.oy0 * P.stride - P.pad;
          uint32_t xbits = 1u, ybits[3] = {0xFFu, 0xFFu, 0xFFu};
          if (ksize == 3) {
            xbits = 0;
#pragma unroll
            for (int t = 0; t < 3; t++) {
              if ((unsigned)(x0 + t) < (unsigned)Win) xbits |= 1u << t;
              if (y0 + t < 0) ybits[t] &= ~1u;
              if (y0 + t + 7 * P.stride >= Hin) ybits[t] &= ~0x80u;
            }
          }
          // byte offset of tap (0,0) of each row inside its image (32-bit), image base 64-bit
          uint32_t offb[8];
          const char* img_base = nullptr;
          uint32_t ld2 = 0;
          int Ws = 0, cur_seg = -1;
          for (int kb = 0; kb < num_kb; kb++) {
            const int s = stage;
            const uint32_t ph = phase;
            const int k0 = kb * BK + g * 8;
            int tap = fast_div(k0, P.pt_mul, P.pt_shr, per_tap);
            int c = k0 - tap * per_tap;
            int seg = 0;
            if (nseg > 1) {
              while (seg + 1 < nseg && c >= P.src_cp[seg]) {
                c -= P.src_cp[seg];
                seg++;
              }
            }
            if (seg != cur_seg) {  // first k-block of the tile, or the K walk entered the next concat source
              cur_seg = seg;
              const int up = P.src_up[seg];
              Ws = Win >> up;
              ld2 = (uint32_t)P.src_ld[seg] * 2u;
              img_base = reinterpret_cast<const char*>(P.src[seg]) +
                         (size_t)tp.n * (size_t)((Hin >> up) * Ws) * ld2;
#pragma unroll
              for (int i = 0; i < 8; i++)
                offb[i] = (uint32_t)(((y0 + i * P.stride) >> up) * Ws + (x0 >> up)) * ld2;
            }
            uint32_t deltab = (uint32_t)c * 2u;
            uint32_t ok8 = (k0 < K && c < P.tap_c) ? 0xFFu : 0u;
            if (ksize == 3) {
              const int dy = (tap * 11) >> 5;  // tap / 3 for tap < 16
              const int dx = tap - dy * 3;
              deltab += (uint32_t)(dy * Ws + dx) * ld2;
              ok8 = (k0 < K && c < P.tap_c && ((xbits >> dx) & 1u)) ? (dy == 0 ? ybits[0] : dy == 1 ? ybits[1] : ybits[2]) : 0u;
            }
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t a_s = a_base + (uint32_t)s * A_STAGE_BYTES + sw_off + (uint32_t)rbase * 128u;
            if (k0 < ((K + 15) & ~15)) {   // granules of K16 steps past the real K are never read by the MMA
              // (through L1 - cp.async.ca, hoping the 2.25x / 9x tap re-reads of a tile hit there - measured slower:
              // net.p2.0 328 -> 382 us; the carve-out left next to two 113 KB CTAs is too small)
#pragma unroll
              for (int i = 0; i < 8; i++)  // masked taps are zero-filled; their source is never read
                cp_async16(a_s + (uint32_t)i * 2048u, img_base + (uint32_t)(offb[i] + deltab),
                           ((ok8 >> i) & 1u) ? 16u : 0u);
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(gathered_bar(s)) : "memory");
            if (++stage == S) {
              stage = 0;
              phase ^= 1u;
            }
          }
          continue;
        }
        // ---- general path: 128 consecutive rows of the flattened (image, y, x) index
        int row_n[8], row_y[8], row_x[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          int m = tp.m0 + rbase + 16 * i;
          if (m < P.M) {
            int n = fast_div(m, P.hw_mul, P.hw_shr, P.hw_out);
            int r = m - n * P.hw_out;
            int oy = fast_div(r, P.w_mul, P.w_shr, P.Wout);
            row_n[i] = n;
            row_y[i] = oy * P.stride - P.pad;
            row_x[i] = (r - oy * P.Wout) * P.stride - P.pad;
          } else {
            row_n[i] = -1;
            row_y[i] = 0;
            row_x[i] = 0;
          }
        }
        int pix[8];
        uint32_t okmask[8];
        int cur_seg = -1;
        for (int kb = 0; kb < num_kb; kb++) {
          const int s = stage;
          const uint32_t ph = phase;
          const int k0 = kb * BK + g * 8;
          int tap = fast_div(k0, P.pt_mul, P.pt_shr, P.per_tap);
          int c = k0 - tap * P.per_tap;
          const bool k_ok = k0 < P.K && c < P.tap_c;
          int seg = 0;
          while (seg + 1 < P.nseg && c >= P.src_cp[seg]) {
            c -= P.src_cp[seg];
            seg++;
          }
          const int up = P.src_up[seg];
          const int Ws = P.Win >> up;
          if (seg != cur_seg) {  // new tile, or the K walk crossed into the next source of a concat
            cur_seg = seg;
            const int Hs = P.Hin >> up;
#pragma unroll
            for (int i = 0; i < 8; i++) {
              uint32_t mk = 0;
              if (row_n[i] >= 0) {
                if (P.ksize == 3) {
                  uint32_t xm = 0, ym = 0;
#pragma unroll
                  for (int t = 0; t < 3; t++) {
                    if ((unsigned)(row_x[i] + t) < (unsigned)P.Win) xm |= 1u << t;
                    if ((unsigned)(row_y[i] + t) < (unsigned)P.Hin) ym |= 7u << (3 * t);
                  }
                  mk = (xm * 0x49u) & ym;
                } else {
                  mk = 1u;
                }
              }
              okmask[i] = mk;
              // ksize 3 never reads an upsampled source, so (y >> up) only matters for tap (0,0)
              pix[i] = (row_n[i] * Hs + (row_y[i] >> up)) * Ws + (row_x[i] >> up);
            }
          }
          int delta = 0, tbit = 0;
          if (P.ksize == 3) {
            int dy = (tap * 11) >> 5;
            delta = dy * Ws + (tap - dy * 3);
            tbit = tap;
          }
          const act_t* sp = P.src[seg] + c;
          const int ld = P.src_ld[seg];
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t a_s = a_base + (uint32_t)s * A_STAGE_BYTES + sw_off;
          if (k0 < ((P.K + 15) & ~15)) {   // granules of K16 steps past the real K are never read by the MMA
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const bool ok = k_ok && ((okmask[i] >> tbit) & 1u);
              const int idx = ok ? pix[i] + delta : 0;
              cp_async16(a_s + (uint32_t)(rbase + 16 * i) * 128u, sp + (long long)idx * ld, ok ? 16u : 0u);
            }
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(gathered_bar(s)) : "memory");
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ============================ MMA issuer ==============================================
    // One thread runs the whole issue loop (tcgen05.mma / commit are single-thread instructions):
    // with N <= 64 an MMA retires in ~48 cycles, so every instruction on this thread's path counts.
    // (elect.sync instead of lane == 0: the compiler then knows a single thread issues and drops the
    // elect / retry wrapper it otherwise puts around every tcgen05.mma)
    if (elect_one()) {
    // instruction descriptor: D fp32, A/B format 0 = fp16 / 1 = bf16, K-major both, N, M
    const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);
    const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) |
                             ((uint64_t)1 << 16);
    int ti = 0, stage = 0, pstage = 0;
    uint32_t phase = 0, pphase = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ti++) {
      const int acc = ti & 1;
      mbar_wait(tmem_empty_bar(acc), ((uint32_t)(ti >> 1) & 1u) ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * (PAIR ? 2 * BN : BN));
      if (PATCH) {
        // k-blocks in consumption order: (channel block, tap) for >= 64 channels, else the packed
        // (tap, channel) order.  Every K=16 step reads the patch through a descriptor shifted by the
        // tap's pixel offset: the swizzle is a function of the shared-memory address bits, so any
        // 16-byte-aligned start inside the TMA-written patch is a valid operand origin.
        const int cblk = P.cblk;
        const uint32_t row_bytes = (uint32_t)cblk * 2u;
        const uint64_t adesc_hi = ((uint64_t)((PPW * row_bytes) >> 4) << 32) | ((uint64_t)1 << 46) |
                                  ((uint64_t)P.a_layout << 61);
        if (P.s2d) {
          // stride-2 3x3 over a space-to-depth source (16 channels: one 64-channel patch holds the four parity
          // blocks): original tap (ky, kx) reads parity block (ky != 1, kx != 1) of the pixel one row up / one
          // column left when ky == 0 / kx == 0 - nine K=16 MMAs, weights in their usual tap-major packing
          if (P.s2d == 2) {
            // 64 channels: one patch per parity block, loaded in block order; block (py, px) serves the taps with
            // (ky != 1) == py, (kx != 1) == px - 1 + 2 + 2 + 4 taps of four K=16 steps, weight k-block = tap
            if (ti == 0)
              for (int kb = 0; kb < num_kb; kb++) mbar_wait(full_bar(kb), 0u);
            const uint32_t a_hi = (uint32_t)(adesc_hi >> 32), b_hi = (uint32_t)(desc_hi >> 32);
            const uint32_t b_lo = (b_base >> 4) | (1u << 16), bstep = b_stage_bytes >> 4;
#pragma unroll
            for (int pl = 0; pl < 4; pl++) {
              mbar_wait(patch_full_bar(pstage), pphase);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t a_lo = ((a_base + (uint32_t)pstage * (uint32_t)P.patch_stage_bytes) >> 4) | (1u << 16);
#pragma unroll
              for (int t = 0; t < 9; t++) {
                const int ky = t / 3, kx = t % 3;
                if (((ky != 1) * 2 + (kx != 1)) != pl) continue;   // compile-time
                const int a_off = (((ky != 0) * PPW + (kx != 0)) * 128) >> 4;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                  const uint64_t da = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)(a_off + 2 * k));
                  const uint64_t db = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)t * bstep + 2u * (uint32_t)k);
                  umma_bf16(d_tmem, da, db, idesc, (uint32_t)!(pl == 0 && t == 4 && k == 0));
                }
              }
              umma_commit(patch_empty_bar(pstage));
              if (pl == 3) umma_commit(tmem_full_bar(acc));
              if (++pstage == P.patch_stages) {
                pstage = 0;
                pphase ^= 1u;
              }
            }
            continue;
          }
          mbar_wait(patch_full_bar(pstage), pphase);
          if (ti == 0)
            for (int kb = 0; kb < num_kb; kb++) mbar_wait(full_bar(kb), 0u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          {
            const uint32_t a_lo = ((a_base + (uint32_t)pstage * (uint32_t)P.patch_stage_bytes) >> 4) | (1u << 16);
            const uint32_t a_hi = (uint32_t)(adesc_hi >> 32), b_hi = (uint32_t)(desc_hi >> 32);
            const uint32_t b_lo = (b_base >> 4) | (1u << 16), bstep = b_stage_bytes >> 4;
#pragma unroll
            for (int t = 0; t < 9; t++) {
              const int ky = t / 3, kx = t % 3;
              const int a_off = (((ky != 0) * PPW + (kx != 0)) * 128 + ((ky != 1) * 2 + (kx != 1)) * 32) >> 4;
              const uint64_t da = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)a_off);
              const uint64_t db = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)(t / 4) * bstep + 2u * (uint32_t)(t % 4));
              umma_bf16(d_tmem, da, db, idesc, (uint32_t)(t != 0));
            }
            umma_commit(patch_empty_bar(pstage));
            umma_commit(tmem_full_bar(acc));
          }
          if (++pstage == P.patch_stages) {
            pstage = 0;
            pphase ^= 1u;
          }
          continue;
        }
        if (RES && ti > 0) {
          // steady state with resident weights: nothing to wait for but the patches
          for (int cb = 0; cb < P.ncb; cb++) {
            mbar_wait(patch_full_bar(pstage), pphase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
              const uint32_t a_lo = (a_base + (uint32_t)pstage * (uint32_t)P.patch_stage_bytes) >> 4;
              const uint32_t a_hi = (uint32_t)(adesc_hi >> 32), b_hi = (uint32_t)(desc_hi >> 32);
              const uint32_t bstep = b_stage_bytes >> 4;
              const uint32_t b_lo = (b_base >> 4) + (uint32_t)(cb * 9) * bstep;
              const bool first = cb == 0;
              const int kvalid = min(4, (P.patch_creal - cb * 64 + 15) >> 4);
              if (cblk == 64) {
                if (kvalid >= 4) issue_patch_steady<64, PAIR, true>(d_tmem, a_lo, a_hi, b_lo, b_hi, bstep, idesc, first, (uint32_t)BN, 4);
                else issue_patch_steady<64, PAIR, false>(d_tmem, a_lo, a_hi, b_lo, b_hi, bstep, idesc, first, (uint32_t)BN, kvalid);
              } else if (cblk == 32) issue_patch_steady<32, PAIR, true>(d_tmem, a_lo, a_hi, b_lo, b_hi, bstep, idesc, first, (uint32_t)BN, 4);
              else if (cblk == 16) issue_patch_steady<16, PAIR, true>(d_tmem, a_lo, a_hi, b_lo, b_hi, bstep, idesc, first, (uint32_t)BN, 4);
              else issue_patch_steady<8, PAIR, true>(d_tmem, a_lo, a_hi, b_lo, b_hi, bstep, idesc, first, (uint32_t)BN, 4);
              umma_commit(patch_empty_bar(pstage));
              if (cb == P.ncb - 1) umma_commit(tmem_full_bar(acc));
            }
              if (++pstage == P.patch_stages) {
              pstage = 0;
              pphase ^= 1u;
            }
          }
          continue;
        }
        for (int kb = 0; kb < num_kb; kb++) {
          const bool first_of_patch = cblk >= 64 ? (kb % 9 == 0) : (kb == 0);
          const bool last_of_patch = cblk >= 64 ? (kb % 9 == 8) : (kb == num_kb - 1);
          if (first_of_patch) mbar_wait(patch_full_bar(pstage), pphase);
          if (!RES || ti == 0) mbar_wait(full_bar(stage), phase);   // resident weights: stage == kb, loaded once
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          {
            const uint32_t pa = a_base + (uint32_t)pstage * (uint32_t)P.patch_stage_bytes;
            const uint64_t db = desc_hi | (uint64_t)(((b_base + (uint32_t)stage * b_stage_bytes) >> 4) & 0x3FFF);
#pragma unroll
            for (int k = 0; k < BK / 16; k++) {
              uint32_t a_addr;
              uint64_t lbo = 1;
              if (cblk >= 64) {
                if ((kb / 9) * 64 + k * 16 >= P.patch_creal) break;   // zero-padded channels of a tap-aligned layer
                const int tap = kb % 9, dy = (tap * 11) >> 5, dx = tap - dy * 3;
                a_addr = pa + (uint32_t)(dy * PPW + dx) * row_bytes + 32u * k;
              } else if (cblk >= 16) {
                const int kg = kb * BK + k * 16;
                const int tap = kg / cblk;
                if (tap >= 9) break;
                const int dy = (tap * 11) >> 5, dx = tap - dy * 3;
                a_addr = pa + (uint32_t)(dy * PPW + dx) * row_bytes + (uint32_t)(kg - tap * cblk) * 2u;
              } else {
                // 8 channels: one MMA covers taps (2j, 2j+1); LBO = distance between their pixels
                const int j = kb * 4 + k;
                if (j >= 5) break;
                const int t0 = 2 * j, t1 = j < 4 ? 2 * j + 1 : 2 * j;
                const int dy0 = (t0 * 11) >> 5, dy1 = (t1 * 11) >> 5;
                const uint32_t o0 = (uint32_t)(dy0 * PPW + (t0 - dy0 * 3)) * 16u;
                const uint32_t o1 = (uint32_t)(dy1 * PPW + (t1 - dy1 * 3)) * 16u;
                a_addr = pa + o0;
                lbo = j < 4 ? (uint64_t)((o1 - o0) >> 4) : 1;
              }
              const uint64_t da = adesc_hi | (lbo << 16) | (uint64_t)((a_addr >> 4) & 0x3FFF);
              umma_bf16(d_tmem, da, db + 2u * k, idesc, (uint32_t)((kb | k) != 0));
              if (PAIR)
                umma_bf16(d_tmem + (uint32_t)BN, da + (uint64_t)((PT_W * row_bytes) >> 4), db + 2u * k, idesc,
                          (uint32_t)((kb | k) != 0));
            }
            umma_commit(empty_bar(stage));
            if (last_of_patch) umma_commit(patch_empty_bar(pstage));
            if (kb == num_kb - 1) umma_commit(tmem_full_bar(acc));
          }
          if (last_of_patch && ++pstage == P.patch_stages) {
            pstage = 0;
            pphase ^= 1u;
          }
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
        continue;
      }
      // Low descriptor words as running 32-bit values (address >> 4; every operand lies inside the CTA's shared
      // memory window, so adding offsets never carries out of the 14 address bits): per k-block two adds, then
      // four MMAs at compile-time offsets.  Only the last k-block can hold K16 steps past the real K.
      const uint32_t dhi = (uint32_t)(desc_hi >> 32);
      const uint32_t a_ring = (ring_base >> 4) | (1u << 16), b_ring = (b_base >> 4) | (1u << 16);
      const uint32_t a_step = A_STAGE_BYTES >> 4, b_step = b_stage_bytes >> 4;
      for (int kb = 0; kb < num_kb; kb++) {
        mbar_wait(full_bar(stage), phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
          const uint32_t da_lo = a_ring + (uint32_t)stage * a_step;
          const uint32_t db_lo = b_ring + (uint32_t)(RES ? kb : stage) * b_step;
          const uint64_t da = ((uint64_t)dhi << 32) | da_lo, db = ((uint64_t)dhi << 32) | db_lo;
          // K16 steps past the end of a source slice (concat walked as separate K segments, each padded to a
          // 64-channel block) or past the real K hold zeros: not issued
          const int kv = kb < 32 ? (int)P.kb_kv[kb] : BK / 16;
          if (kb == 0) {
            umma_bf16(d_tmem, da, db, idesc, 0u);
          } else {
            umma_bf16(d_tmem, da, db, idesc, 1u);
          }
          if (kv == BK / 16) {
#pragma unroll
            for (int k = 1; k < BK / 16; k++) umma_bf16(d_tmem, da + 2u * k, db + 2u * k, idesc, 1u);
          } else {
#pragma unroll
            for (int k = 1; k < BK / 16; k++)
              if (k < kv) umma_bf16(d_tmem, da + 2u * k, db + 2u * k, idesc, 1u);
          }
          umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs have read it
          if (kb == num_kb - 1) umma_commit(tmem_full_bar(acc));  // accumulator complete -> epilogue
        }
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    }
  } else if (DW && warp >= DW_WARP0) {
      // ============================ depthwise producer =======================================
      // thread = 2 channels (cpair) x 4 pixel columns (xh) x 4 tile rows (rq).  Patch rows stream
      // through registers once; a row feeds the three output rows it overlaps (ky = 0, 1, 2), whose
      // accumulators rotate through three slots.  Lanes of a warp are consecutive channel pairs, so
      // every shared-memory access of a warp is one 128-byte row.
      const int ptid = tid - DW_WARP0 * 32;
      const int cpair = ptid & 31, q = ptid >> 5, xh = q & 1, rq = q >> 1;
      constexpr int RPT = PT_H / (DW_THREADS / 64);   // tile rows per thread: 64 threads (32 pairs x 2 halves) per row group
      const int CW = P.ncb * 64;
      float* dww = bias_s + 256;   // [10][CW]: 9 taps + bias, zero beyond the real channels
      for (int i = ptid; i < 10 * CW; i += DW_THREADS) {
        const int tap = i / CW, c = i - tap * CW;
        // SiLU(x) = h + h*tanh(h), h = x/2: the 1/2 is folded into the depthwise weights and bias (exact)
        dww[i] = c < P.dw_C ? (P.dw_act ? 0.5f : 1.f) * __ldg(P.dw_w + tap * P.dw_cp + c) : 0.f;
      }
      asm volatile("bar.sync 3, %0;" ::"n"(DW_THREADS) : "memory");
      pdl_wait();
      int stage = 0, pstage = 0;
      uint32_t phase = 0, pphase = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        for (int cb = 0; cb < P.ncb; cb++) {
          float2 w[9];
#pragma unroll
          for (int t9 = 0; t9 < 9; t9++) w[t9] = *reinterpret_cast<const float2*>(dww + t9 * CW + cb * 64 + 2 * cpair);
          const float2 bias2 = *reinterpret_cast<const float2*>(dww + 9 * CW + cb * 64 + 2 * cpair);
          mbar_wait(patch_full_bar(pstage), pphase);
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t pbase = a_base + (uint32_t)pstage * (uint32_t)P.patch_stage_bytes +
                                 (uint32_t)((RPT * rq) * PP_W + 4 * xh) * 128u + (uint32_t)cpair * 4u;
          const uint32_t sbase = ring_base + (uint32_t)stage * A_STAGE_BYTES + (uint32_t)(cpair & 3) * 4u;
          if (P.dw_dbg & 4) {
            // (ablation switch: no depthwise math at all)
          } else if (F16 && !(P.dw_dbg & 1)) {
            // fp16 storage: the whole depthwise conv runs on packed halves - the loaded channel pair IS the HFMA2
            // operand (no unpack), the accumulator pair goes through one tanh.approx.f16x2 + one HFMA2 for SiLU
            // and is stored as is (no pack).  9 fp16 roundings on the accumulator instead of one on the output:
            // class branch only (score tolerance 1e-2; measured score error stays < 1e-3).
            uint32_t wh[9];
#pragma unroll
            for (int t9 = 0; t9 < 9; t9++) wh[t9] = A16::pack2(w[t9].x, w[t9].y);
            const uint32_t biash = A16::pack2(bias2.x, bias2.y);
            uint32_t acch[3][4];
#pragma unroll
            for (int pr = 0; pr < RPT + 2; pr++) {
              uint32_t f[6];
#pragma unroll
              for (int i = 0; i < 6; i++)
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(f[i]) : "r"(pbase + (uint32_t)(pr * PP_W + i) * 128u));
#pragma unroll
              for (int ky = 0; ky < 3; ky++) {
                const int r = pr - ky;
                if (r < 0 || r >= RPT) continue;
                const int slot = r % 3;
#pragma unroll
                for (int px = 0; px < 4; px++) {
                  if (ky == 0) acch[slot][px] = biash;
#pragma unroll
                  for (int kx = 0; kx < 3; kx++) acch[slot][px] = hfma2_u32(f[px + kx], wh[ky * 3 + kx], acch[slot][px]);
                }
                if (ky == 2) {
#pragma unroll
                  for (int px = 0; px < 4; px++) {
                    const uint32_t o = P.dw_act ? silu_half_h2(acch[slot][px]) : acch[slot][px];
                    const uint32_t m = (uint32_t)((RPT * rq + r) * PT_W + 4 * xh + px);
                    const uint32_t addr = sbase + m * 128u + ((((uint32_t)cpair >> 2) ^ (m & 7u)) << 4);
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(o) : "memory");
                  }
                }
              }
            }
          } else {
          float2 acc[3][4];
#pragma unroll
          for (int pr = 0; pr < RPT + 2; pr++) {
            float2 f[6];
#pragma unroll
            for (int i = 0; i < 6; i++) {
              uint32_t v;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(pbase + (uint32_t)(pr * PP_W + i) * 128u));
              f[i] = A16::unpack2(v);
            }
#pragma unroll
            for (int ky = 0; ky < 3; ky++) {
              const int r = pr - ky;          // output row (of this thread's RPT rows) fed through tap row ky
              if (r < 0 || r >= RPT) continue;
              const int slot = r % 3;
#pragma unroll
              for (int px = 0; px < 4; px++) {
                if (ky == 0) acc[slot][px] = bias2;
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                  // packed fp32 FMA on the channel pair
                  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&f[px + kx]);
                  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&w[ky * 3 + kx]);
                  unsigned long long rc = *reinterpret_cast<unsigned long long*>(&acc[slot][px]), rd;
                  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
                  acc[slot][px] = *reinterpret_cast<float2*>(&rd);
                }
              }
              if (ky == 2) {   // row r complete: activation, bf16, swizzled K-major A stage
#pragma unroll
                for (int px = 0; px < 4; px++) {
                  float2 o = acc[slot][px];
                  if (P.dw_act) {
                    o.x = silu_half(o.x);
                    o.y = silu_half(o.y);
                  }
                  const uint32_t m = (uint32_t)((RPT * rq + r) * PT_W + 4 * xh + px);
                  const uint32_t addr = sbase + m * 128u + ((((uint32_t)cpair >> 2) ^ (m & 7u)) << 4);
                  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(A16::pack2(o.x, o.y)) : "memory");
                }
              }
            }
          }
          }
          // generic-proxy writes -> visible to the tensor core's async-proxy reads; one arrival per warp
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(full_bar(stage));
            mbar_arrive(patch_empty_bar(pstage));
          }
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
          if (++pstage == P.patch_stages) {
            pstage = 0;
            pphase ^= 1u;
          }
        }
      }
  } else if (warp == RELAY_WARP) {
    // ============================ proxy-fence relay ========================================
    // im2col data is written by cp.async (generic proxy) but read by the tensor core through the
    // async proxy.  This thread acquires a gathered stage, issues the proxy fence and forwards the
    // arrival to the MMA warp, keeping the (expensive) fence off the MMA issue path.
    pdl_wait();
    if (PGEO && lane == 0) {
      // ---- halo-patch producer: one 4-D TMA box {cblk, 10, 18, 1} per (tile, channel block);
      // coordinates start one pixel above / left of the tile, out-of-image pixels are zero-filled
      // by the TMA unit (= the convolution's padding)
      int pstage = 0;
      uint32_t pphase = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const TilePos tp = tile_pos<T2D, TW, TH>(P, P.n_tiles == 1 ? tile : tile / P.n_tiles);
        for (int cb = 0; cb < P.ncb; cb++) {
          mbar_wait_sleep(patch_empty_bar(pstage), pphase ^ 1u, (uint32_t)P.wait_sleep_ns);
          mbar_expect_tx(patch_full_bar(pstage), (uint32_t)P.patch_tx_bytes);
          tma_load_4d(a_base + (uint32_t)pstage * (uint32_t)P.patch_stage_bytes, &tmap_a0, cb * 64, tp.ox0 - 1,
                      tp.oy0 - 1, tp.n, patch_full_bar(pstage));
          if (++pstage == P.patch_stages) {
            pstage = 0;
            pphase ^= 1u;
          }
        }
      }
    }
    if (MODE == MODE_GATHER && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(gathered_bar(stage), phase);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(full_bar(stage));
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ============================ TMA producer ============================================
    if (lane == 0) {
      const uint32_t a_tx = MODE == MODE_ATMA ? (uint32_t)A_STAGE_BYTES : 0u;
      int stage = 0;
      uint32_t phase = 0;
      // The weight k-blocks of the CTA's first tile that fit the (still empty) stage ring are requested
      // before the wait for the previous kernel: their DRAM / L2 latency is off the critical path.
      int pre = 0;
      if ((int)blockIdx.x < P.total_tiles) {
        const int n0f = ((int)blockIdx.x - (P.n_tiles == 1 ? (int)blockIdx.x : (int)blockIdx.x / P.n_tiles) * P.n_tiles) * BN;
        pre = min(num_kb, S);
        for (int kb = 0; kb < pre; kb++) {
          mbar_expect_tx(full_bar(kb), a_tx + b_stage_bytes);
          const int kb_w = (PATCH && P.cblk >= 64 && !P.s2d) ? (kb % 9) * P.ncb + kb / 9 : kb;
          tma_load_2d(b_base + (uint32_t)kb * b_stage_bytes, &tmap_b, kb_w * BK, n0f, full_bar(kb));
        }
      }
      pdl_wait();
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        // resident weights are fetched during the CTA's first tile only
        const bool load_b = !RES || tile == (int)blockIdx.x;
        if (PATCH && !load_b) break;   // nothing left to load: the MMA warp no longer waits on these barriers
        const uint32_t tx = a_tx + (load_b ? b_stage_bytes : 0u);
        const int mt = P.n_tiles == 1 ? tile : tile / P.n_tiles;
        const TilePos tp = tile_pos<T2D, TW, TH>(P, mt);
        const int n0 = (tile - mt * P.n_tiles) * BN;
        int seg = 0, kk = 0;
        for (int kb = 0; kb < num_kb; kb++) {
          const int s = stage;
          const uint32_t ph = phase;
          const bool early = tile == (int)blockIdx.x && kb < pre;   // expect_tx + weight load already issued
          if (!early) {
            mbar_wait_sleep(empty_bar(s), ph ^ 1u, (uint32_t)P.wait_sleep_ns);
            if (tx) mbar_expect_tx(full_bar(s), tx);
            else mbar_arrive(full_bar(s));   // im2col layer with resident weights: only the gather feeds this stage
          }
          // PATCH with >= 64 channels consumes k-blocks channel-block-major; they are packed tap-major
          const int kb_w = (PATCH && P.cblk >= 64 && !P.s2d) ? (kb % 9) * P.ncb + kb / 9 : kb;
          if (load_b && !early)
            tma_load_2d(b_base + (uint32_t)(RES ? kb : s) * b_stage_bytes, &tmap_b, kb_w * BK, n0, full_bar(s));
          if (MODE == MODE_ATMA) {
            const CUtensorMap* ma =
                seg == 0 ? &tmap_a0 : seg == 1 ? &tmap_a1 : seg == 2 ? &tmap_a2 : &tmap_a3;
            if (T2D)
              tma_load_4d(a_base + (uint32_t)s * A_STAGE_BYTES, ma, kk * BK, tp.ox0, tp.oy0, tp.n, full_bar(s));
            else
              tma_load_2d(a_base + (uint32_t)s * A_STAGE_BYTES, ma, kk * BK, tp.m0, full_bar(s));
            if (++kk == P.seg_kb[seg]) {
              kk = 0;
              seg++;
            }
          }
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// Scalar cross-check kernel (validation only; selected with yb_plan_set_conv_impl(plan, 1)).
// One thread per (output row, output channel), same packed weights, fp32 accumulation.
// ------------------------------------------------------------------------------------------
template <bool F16>
__global__ void conv_direct_check_kernel(const ConvParams P) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int cs = P.cout_store;
  if (idx >= (long long)P.M * cs) return;
  int m = (int)(idx / cs);
  int n = (int)(idx - (long long)m * cs);
  int img = m / P.hw_out;
  int r = m - img * P.hw_out;
  int oy = r / P.Wout, ox = r - oy * P.Wout;
  float acc = 0.f;
  const act_t* wrow = P.w + (size_t)n * P.K_pad;
  int taps = P.ksize * P.ksize;
  for (int tap = 0; tap < taps; tap++) {
    int dy = P.ksize == 3 ? tap / 3 : 0, dx = P.ksize == 3 ? tap % 3 : 0;
    int iy = oy * P.stride - P.pad + dy, ix = ox * P.stride - P.pad + dx;
    bool ok = (unsigned)iy < (unsigned)P.Hin && (unsigned)ix < (unsigned)P.Win;
    int kpos = P.a_tma ? 0 : tap * P.per_tap;
    for (int s = 0; s < P.nseg; s++) {
      if (ok) {
        int up = P.src_up[s];
        int Hs = P.Hin >> up, Ws = P.Win >> up;
        const act_t* sp =
            P.s2d ? P.src[s] + ((((size_t)img * (Hs >> 1) + (iy >> 1)) * (Ws >> 1) + (ix >> 1)) * 4 + (iy & 1) * 2 + (ix & 1)) * (size_t)P.src_ld[s]
                  : P.src[s] + ((size_t)(img * Hs + (iy >> up)) * Ws + (ix >> up)) * (size_t)P.src_ld[s];
        for (int c = 0; c < P.src_c[s]; c++)
          acc += Act16<F16>::unpack1(sp[c]) * Act16<F16>::unpack1(wrow[kpos + c]);
      }
      kpos += P.seg_kpad[s];
    }
  }
  float x = acc + P.bias[n];
  if (P.act) x = x / (1.f + expf(-x));
  size_t drow = (size_t)img * P.dst_rows_per_img + P.dst_row_off + r;
  if (P.out_f32) {
    reinterpret_cast<float*>(P.dst)[drow * (size_t)P.dst_ld + n] = x;
  } else {
    if (P.res) x += Act16<F16>::unpack1(P.res[(size_t)m * P.res_ld + n]);
    if (P.c_s2d)
      drow = (((size_t)img * (P.Hout >> 1) + (oy >> 1)) * (P.Wout >> 1) + (ox >> 1)) * 4 + (oy & 1) * 2 + (ox & 1);
    reinterpret_cast<act_t*>(P.dst)[drow * (size_t)P.dst_ld + n] = Act16<F16>::pack1(x);
  }
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// L2 promotion (the granularity at which a TMA miss fetches from DRAM) must not exceed the bytes a row
// of the box really uses: a 16-channel slice of a 32-channel buffer is 32 useful bytes per 64-byte pixel,
// and 256-byte promotion would drag the other slice through DRAM as well (measured: 3x the algorithmic
// reads on the C2f/C3k2 bottleneck inputs).
static CUtensorMapL2promotion promo_for(uint64_t used_bytes, uint64_t pitch_bytes) {
  static const bool force256 = getenv("YB_TMA_PROMO256") != nullptr;
  if (force256 || used_bytes >= pitch_bytes || used_bytes >= 256) return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  if (used_bytes >= 128) return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (used_bytes >= 64) return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
  return CU_TENSOR_MAP_L2_PROMOTION_NONE;
}

static thread_local bool g_tmap_f16 = true;   // element type of the maps being encoded (set by conv_tc_prepare)

static int make_tmap_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows,
                        uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return YB_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, g_tmap_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   promo_for(std::min<uint64_t>(box_inner, inner) * 2, row_stride_bytes),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with %d (inner=%llu rows=%llu stride=%llu box=%ux%u)",
              (int)r, (unsigned long long)inner, (unsigned long long)rows,
              (unsigned long long)row_stride_bytes, box_inner, box_rows);
    return YB_ERR_CUDA;
  }
  return YB_OK;
}

// 4-D map over an NHWC slice {C, W, H, N} with an 8 (y) x 16 (x) x 64-channel box: smem rows come
// out as r = ly * 16 + lx, 128 bytes each, SWIZZLE_128B — the same image as a 128-row 2-D box.
static int make_tmap_nhwc(CUtensorMap* map, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N,
                          uint64_t ld_elems, uint32_t box_c = 64, uint32_t box_w = 16, uint32_t box_h = 8,
                          bool no_swizzle = false) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return YB_ERR_CUDA;
  }
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {ld_elems * 2, W * ld_elems * 2, H * W * ld_elems * 2};
  cuuint32_t box[4] = {box_c, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  // rows of box_c channels: the swizzle span equals the row size (16-byte rows are not swizzled)
  const CUtensorMapSwizzle sw = no_swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE : box_c >= 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : box_c == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(map, g_tmap_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo_for(std::min<uint64_t>(box_c, C) * 2, ld_elems * 2),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (4-D NHWC) failed with %d (C=%llu W=%llu H=%llu N=%llu ld=%llu)", (int)r,
              (unsigned long long)C, (unsigned long long)W, (unsigned long long)H, (unsigned long long)N,
              (unsigned long long)ld_elems);
    return YB_ERR_CUDA;
  }
  return YB_OK;
}

// Store map of a space-to-depth buffer (logical (H, W, C) kept as (H/2, W/2, 4C), channel block (y&1)*2 + (x&1)): dims
// {c, x&1, x/2, y&1, n*H/2 + y/2}; the box {64, 2, 8, 2, 4} is the 8 x 16 pixel output tile in its usual staging order.
static int make_tmap_s2d_store(CUtensorMap* map, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return YB_ERR_CUDA;
  }
  cuuint64_t dims[5] = {C, 2, W / 2, 2, N * (H / 2)};
  cuuint64_t strides[4] = {C * 2, 4 * C * 2, 2 * C * 2, (W / 2) * 4 * C * 2};
  cuuint32_t box[5] = {64, 2, 8, 2, 4};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, g_tmap_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (space-to-depth store) failed with %d (C=%llu W=%llu H=%llu N=%llu)", (int)r,
              (unsigned long long)C, (unsigned long long)W, (unsigned long long)H, (unsigned long long)N);
    return YB_ERR_CUDA;
  }
  return YB_OK;
}

// a_region: bytes of the A ring when it is not `stages` x 16 KB (PATCH mode), else 0;
// b_slots: weight slots (num_kb when the weights are resident), 0 = one per stage
static size_t conv_smem_bytes(int stages, int BN, size_t a_region = 0, int b_slots = 0, int c_bufs = 1) {
  const int groups = (BN + 63) / 64;
  return 1024 + (a_region ? a_region : (size_t)stages * A_STAGE_BYTES) + (size_t)(b_slots ? b_slots : stages) * BN * 128 +
         (size_t)groups * C_GROUP_BYTES * c_bufs + NUM_BARS * 8 + 16 + 256 * 4 + 64;
}

static bool patch_eligible(const yb_plan* p, const Op& op) {
  if (getenv("YB_NO_PATCH")) return false;
  if (op.s2d) return true;   // decided with the buffer layout when the plan was built
  if (op.k != 3 || op.stride != 1 || op.nseg != 1 || op.src[0].up || op.out_f32) return false;
  const int C = op.seg_kpad[0];   // channels per tap as packed (a multiple of 64 for tap-aligned layers)
  if (!(C == 8 || C == 16 || C == 32 || (C % 64 == 0 && C > 0))) return false;
  int min_hw = 40;  // below this the 16 x 8 tiles hang too far over the image edge
  if (const char* e = getenv("YB_PATCH_MIN_HW")) min_hw = atoi(e);
  (void)p;
  return op.Hout >= min_hw && op.Wout >= min_hw;
}

static int tmem_cols_for(int BN) {
  int cols = 32;
  while (cols < 2 * BN) cols <<= 1;  // two accumulator stages
  return cols;
}

int conv_tc_prepare(yb_plan* p, Op& op) {
  g_tmap_f16 = p->act_f16 != 0;
  const ConvW& cw = p->convs[op.conv_index];
  const uint8_t* wbase = p->d_weights + cw.info.blob_offset;
  // Resident CTAs per SM: bounded by TMEM (512 columns per SM, 2 accumulator stages per CTA) and
  // capped at 2.  The shared-memory request is sized so that exactly `occ` CTAs fit, which keeps
  // tcgen05.alloc from ever waiting on a co-resident persistent CTA.
  const size_t SMEM_MAX = 227 * 1024;
  // patch layers with tiny N tiles are bound by per-tile latency chains: a third resident CTA hides them
  // (space-to-depth sources: 24 KB patches, two CTAs per SM)
  const bool tiny_patch = patch_eligible(p, op) && !op.s2d && op.BN <= 32 && op.Hout >= 80 && getenv("YB_NO_OCC3") == nullptr;
  int occ = std::min(tiny_patch ? 3 : 2, 512 / tmem_cols_for(op.BN));
  if (op.dw_fused) occ = 1;   // 19-warp CTA (8 depthwise warps): one per SM
  if (const char* e = getenv("YB_OCC")) occ = std::max(1, std::min(occ, atoi(e)));
  if (!op.a_tma)
    if (const char* e = getenv("YB_OCC_GATHER")) occ = std::max(1, std::min(occ, atoi(e)));
  op.patch = patch_eligible(p, op) ? 1 : 0;
  // pair tiles: two accumulators share every weight k-block (halves the weight traffic per output pixel and
  // the per-tile fixed costs); needs 4 x BN TMEM columns and one N tile
  // (measured: wins on the wide 160 x 160 layers - YOLO11x C=48: 0.45 -> 0.34 ms; neutral on YOLO11n's 80 x 80 box
  // branch, a loss on 40 x 40 maps and on the 8-32 channel layers, which prefer single tiles with alternating
  // epilogue groups.  YB_PAIR=0/1 overrides)
  const int pair_env = getenv("YB_PAIR") ? atoi(getenv("YB_PAIR")) : -1;
  op.pair = (op.patch && op.N_pad == op.BN && 4 * tmem_cols_for(op.BN) / 2 <= 512 && op.Wout >= 16 &&
             (pair_env < 0 ? (op.BN >= 48 && op.Hout >= 160) : pair_env != 0)) ? 1 : 0;
  if (op.s2d) op.pair = 0;
  if (op.pair) occ = std::min(occ, std::max(1, 512 / tmem_cols_for(2 * op.BN)));
  const int num_kb = op.K_pad / BK;
  const size_t w_bytes = (size_t)num_kb * op.BN * 128;
  // Weights resident in shared memory (loaded once per CTA instead of once per tile) whenever the
  // whole [BN x K_pad] matrix fits next to a useful A ring: L2 -> SM bandwidth (~43 B/clk/SM) is the
  // scarce resource of the small-channel layers, and the weight tile is 30-60 % of their traffic.
  const bool res_ok = op.N_pad == op.BN && num_kb <= MAX_STAGES && getenv("YB_NO_RESIDENT") == nullptr;
  const bool cb2_ok = op.N_pad == op.BN && op.BN <= 128 && !op.out_f32 && !op.pair && getenv("YB_ONE_CBUF") == nullptr;
  const int halves = op.pair ? 2 : 1;   // staging buffers per tile iteration
  // shared-memory plan: (occupancy, resident weights, staging buffers, minimum A stages) -> stages
  auto plan_smem = [&](int occ_try, bool resident, int cb, int min_st, int& st_out, int& pst_out) -> bool {
    const size_t bud = SMEM_MAX / occ_try;
    if (op.dw_fused) {
      // patch ring (2 stages) + depthwise weights sit next to the usual A ring
      op.patch_stage_bytes = round_up(PP_H * PP_W * 128, 1024);
      // measured (B200, ablations in DESIGN.md): these layers are bound by the instruction streams of the depthwise
      // and epilogue warps, not by patch bytes in flight - more patch stages at the cost of the second output
      // staging buffer (alternating epilogue groups) lose 30 %; three patch stages + four A stages + two staging
      // buffers is the best split of the 227 KB
      const int pst_max = getenv("YB_DW_PST") ? atoi(getenv("YB_DW_PST")) : 3;
      const int st_want = getenv("YB_DW_ST") ? atoi(getenv("YB_DW_ST")) : 4;
      for (int pst = std::min(pst_max, MAX_PATCH_STAGES); pst >= 2; pst--) {
        const size_t extra = (size_t)pst * op.patch_stage_bytes + (size_t)10 * num_kb * 64 * 4;
        for (int st = std::max(2, std::min(st_want, 4)); st >= 2; st--)
          if (conv_smem_bytes(st, op.BN, 0, resident ? num_kb : 0, cb) + extra <= bud) { st_out = st; pst_out = pst; return true; }
      }
      (void)min_st;
      return false;
    }
    if (op.patch) {
      const int C = op.s2d ? 64 : op.seg_kpad[0], cblk = std::min(C, 64);
      op.patch_stage_bytes = round_up(PP_H * (op.pair ? PP2_W : PP_W) * cblk * 2 + 16, 1024);
      for (int pst = MAX_PATCH_STAGES; pst >= 2; pst--) {
        const size_t a_region = (size_t)pst * op.patch_stage_bytes;
        if (resident) {
          if (conv_smem_bytes(num_kb, op.BN, a_region, num_kb, cb * halves) <= bud) { st_out = num_kb; pst_out = pst; return true; }
        } else {
          for (int st = std::min(MAX_STAGES, 8); st >= min_st; st--)
            if (conv_smem_bytes(st, op.BN, a_region, 0, cb * halves) <= bud) { st_out = st; pst_out = pst; return true; }
        }
      }
      return false;
    }
    for (int st = std::min(MAX_STAGES, 8); st >= min_st; st--)
      if (conv_smem_bytes(st, op.BN, 0, resident ? num_kb : 0, cb) <= bud) { st_out = st; pst_out = 0; return true; }
    return false;
  };
  int st = 2, pst = 0, cb = 1;
  bool found = false;
  op.b_resident = 0;
  struct Try { int occ; bool res; int cb; int min_st; };
  // (patch layers: resident weights on one CTA per SM beat streamed weights on two - the streamed path
  // re-reads the matrix per tile and runs the generic MMA issue loop)
  const Try tries[] = {
      {occ, true, 2, 4}, {occ, true, 1, 4}, {1, true, 2, 4}, {1, true, 1, 4}, {occ, false, 2, 5}, {occ, false, 1, 4},
      {occ, true, 1, 3}, {occ, false, 1, 2}};
  const int dw_cb = getenv("YB_DW_CB") ? atoi(getenv("YB_DW_CB")) : 2;
  for (const Try& t : tries) {
    if (t.res && !res_ok) continue;
    if (t.cb == 2 && !cb2_ok) continue;
    if (op.dw_fused && t.cb != dw_cb && cb2_ok) continue;   // (A/B switch YB_DW_CB)
    if (t.occ != occ && !op.dw_fused &&
        !(op.patch && occ > 1 && (w_bytes >= 24 * 1024 || op.s2d) && getenv("YB_NO_RESIDENT_OCC1") == nullptr))
      continue;
    if (plan_smem(t.occ, t.res, t.cb, t.min_st, st, pst)) {
      op.b_resident = t.res ? 1 : 0;
      cb = t.cb;
      occ = t.occ;
      found = true;
      break;
    }
  }
  if (!found && op.dw_fused) {   // last resort for the fused depthwise path: one CTA per SM
    for (int cbt = 2; cbt >= 1 && !found; cbt--)
      for (int res = 1; res >= 0 && !found; res--) {
        if ((res && !res_ok) || (cbt == 2 && !cb2_ok)) continue;
        if (plan_smem(1, res != 0, cbt, cbt == 1 && !res ? 2 : 3, st, pst)) {
          op.b_resident = res;
          cb = cbt;
          occ = 1;
          found = true;
        }
      }
    if (!found) {
      set_error("conv %s: fused depthwise tile does not fit shared memory", op.name.c_str());
      return YB_ERR_UNSUPPORTED;
    }
  }
  if (op.s2d && !(found && op.b_resident)) {
    set_error("conv %s: the space-to-depth path needs its weights resident in shared memory", op.name.c_str());
    return YB_ERR_UNSUPPORTED;
  }
  if (!found && op.patch) {  // patches + streamed weights do not fit: fall back to the gather path
    op.patch = 0;
    plan_smem(occ, false, 1, 2, st, pst);
  }
  op.c_bufs = cb;
  if (!op.patch) op.pair = 0;
  if (const char* e = getenv("YB_STAGES"))
    if (!op.b_resident && !op.patch) st = std::max(1, std::min(st, atoi(e)));
  op.stages = st;
  op.patch_stages = pst;
  const size_t a_region = op.patch ? (size_t)pst * op.patch_stage_bytes : 0;
  const size_t dw_extra = op.dw_fused ? (size_t)pst * op.patch_stage_bytes + (size_t)10 * num_kb * 64 * 4 : 0;
  op.smem_bytes = conv_smem_bytes(st, op.BN, a_region, op.b_resident ? num_kb : 0, cb * halves) + dw_extra;
  if (getenv("YB_SMEM_EXACT") == nullptr) op.smem_bytes = std::max(op.smem_bytes, SMEM_MAX / (occ + 1) + 1024);
  if (op.smem_bytes > SMEM_MAX) {
    set_error("conv %s: tile needs %zu bytes of shared memory", op.name.c_str(), op.smem_bytes);
    return YB_ERR_UNSUPPORTED;
  }
  op.occ = occ;
  int rc = make_tmap_2d(&op.tmap_b, wbase, (uint64_t)op.K_pad, (uint64_t)op.N_pad,
                        (uint64_t)op.K_pad * 2, BK, (uint32_t)op.BN);
  if (rc) return rc;
  for (int i = 0; i < 4; i++) op.tmap_a[i] = op.tmap_b;
  op.tile2d = (op.Hout % 8 == 0 && op.Wout % 16 == 0 && !op.out_f32 && getenv("YB_NO_TILE2D") == nullptr) ? 1 : 0;
  if (op.dw_fused) {
    // the A operand is computed in-kernel from halo patches of the depthwise conv's own input
    const Op& dwop = p->ops[op.dw_op];
    op.tile2d = 1;
    const Buf& b = p->bufs[dwop.src[0].buf];
    const uint8_t* base = buf_ptr(p, dwop.src[0].buf) + (size_t)dwop.src[0].c_off * 2;
    rc = make_tmap_nhwc(&op.tmap_a[0], base, (uint64_t)dwop.src[0].C, (uint64_t)b.W, (uint64_t)b.H, (uint64_t)p->B,
                        (uint64_t)b.C, 64, PP_W, PP_H, true);
    if (rc) return rc;
  } else if (op.patch) {
    op.tile2d = 1;
    const Buf& b = p->bufs[op.src[0].buf];
    const uint8_t* base = buf_ptr(p, op.src[0].buf) + (size_t)op.src[0].c_off * 2;
    if (op.s2d)   // the source as stored: (H/2, W/2, 4C)
      rc = make_tmap_nhwc(&op.tmap_a[0], base, (uint64_t)4 * b.C, (uint64_t)b.W / 2, (uint64_t)b.H / 2, (uint64_t)p->B,
                          (uint64_t)4 * b.C, 64, PP_W, PP_H);
    else
    rc = make_tmap_nhwc(&op.tmap_a[0], base, (uint64_t)op.src[0].C, (uint64_t)b.W, (uint64_t)b.H, (uint64_t)p->B,
                        (uint64_t)b.C, (uint32_t)std::min(op.seg_kpad[0], 64), op.pair ? PP2_W : PP_W, PP_H);
    if (rc) return rc;
  }
  if (op.a_tma && !op.dw_fused) {
    for (int i = 0; i < op.nseg; i++) {
      const Buf& b = p->bufs[op.src[i].buf];
      const uint8_t* base = buf_ptr(p, op.src[i].buf) + (size_t)op.src[i].c_off * 2;
      if (op.tile2d)
        rc = make_tmap_nhwc(&op.tmap_a[i], base, (uint64_t)cpad8(op.src[i].C), (uint64_t)b.W, (uint64_t)b.H,
                            (uint64_t)p->B, (uint64_t)b.C);
      else
        rc = make_tmap_2d(&op.tmap_a[i], base, (uint64_t)cpad8(op.src[i].C),
                          (uint64_t)p->B * b.rows_per_img, (uint64_t)b.C * 2, BK, BM);
      if (rc) return rc;
    }
  }
  op.tmap_c = op.tmap_b;
  if (!op.out_f32) {
    const Buf& db = p->bufs[op.dst.buf];
    const uint8_t* dbase = buf_ptr(p, op.dst.buf) + (size_t)op.dst.c_off * 2;
    if (db.s2d) {
      if (!(op.a_tma && op.tile2d && !op.dw_fused && op.N_pad == op.BN)) {
        set_error("conv %s: cannot store its output space-to-depth", op.name.c_str());
        return YB_ERR_UNSUPPORTED;
      }
      rc = make_tmap_s2d_store(&op.tmap_c, dbase, (uint64_t)db.C, (uint64_t)db.W, (uint64_t)db.H, (uint64_t)p->B);
    } else if (op.patch || op.dw_fused)
      rc = make_tmap_nhwc(&op.tmap_c, dbase, (uint64_t)cpad8(op.dst.C), (uint64_t)db.W, (uint64_t)db.H,
                          (uint64_t)p->B, (uint64_t)db.C, 64, PT_W, PT_H);
    else if (op.tile2d)
      rc = make_tmap_nhwc(&op.tmap_c, dbase, (uint64_t)cpad8(op.dst.C), (uint64_t)db.W, (uint64_t)db.H,
                          (uint64_t)p->B, (uint64_t)db.C);
    else
      rc = make_tmap_2d(&op.tmap_c, dbase, (uint64_t)cpad8(op.dst.C), (uint64_t)p->B * db.rows_per_img,
                        (uint64_t)db.C * 2, 64, BM);
    if (rc) return rc;
  }
  // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute
  static bool attr_set[YB_MAX_DEVICES] = {false};
  bool& done = attr_set[p->device & (YB_MAX_DEVICES - 1)];
  if (!done) {
    const int smem_max = 227 * 1024;
#define YB_ATTR(MODE, T2D, HEAD)                                                                                                                                   \
    YB_CUDA(cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<MODE, T2D, HEAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));  \
    YB_CUDA(cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<MODE, T2D, HEAD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max))
    YB_ATTR(MODE_ATMA, false, true);
    YB_ATTR(MODE_GATHER, false, true);
    YB_ATTR(MODE_ATMA, true, false);
    YB_ATTR(MODE_ATMA, false, false);
    YB_ATTR(MODE_GATHER, true, false);
    YB_ATTR(MODE_GATHER, false, false);
    YB_ATTR(MODE_PATCH, true, false);
    YB_ATTR(MODE_DW, true, false);
    YB_ATTR(MODE_PATCH2, true, false);
#undef YB_ATTR
    done = true;
  }
  return YB_OK;
}

static void fill_params(const yb_plan* p, const Op& op, ConvParams& P) {
  memset(&P, 0, sizeof(P));
  const ConvW& cw = p->convs[op.conv_index];
  P.nseg = op.nseg;
  int per_tap = 0;
  for (int i = 0; i < op.nseg; i++) {
    const Buf& b = p->bufs[op.src[i].buf];
    P.src[i] = reinterpret_cast<const act_t*>(buf_ptr(p, op.src[i].buf)) + op.src[i].c_off;
    P.src_cp[i] = cpad8(op.src[i].C);
    P.src_c[i] = op.src[i].C;
    P.src_ld[i] = b.C;
    P.src_up[i] = op.src[i].up;
    P.seg_kb[i] = op.seg_kpad[i] / BK;
    P.seg_kpad[i] = op.seg_kpad[i];
    per_tap += op.a_tma ? P.src_cp[i] : op.seg_kpad[i];   // tap-aligned 3x3 layers: per-tap K padded to 64
  }
  P.tap_c = (!op.a_tma && op.nseg == 1 && op.seg_kpad[0] != P.src_cp[0]) ? P.src_cp[0] : 0x7fffffff;
  P.per_tap = per_tap;
  P.ksize = op.k;
  P.stride = op.stride;
  P.pad = op.k / 2;
  P.Hin = op.Hin;
  P.Win = op.Win;
  P.Hout = op.Hout;
  P.Wout = op.Wout;
  P.hw_out = op.Hout * op.Wout;
  P.M = p->B * P.hw_out;
  P.K = op.K;
  P.K_pad = op.K_pad;
  P.num_kb = op.K_pad / BK;
  const Buf& db = p->bufs[op.dst.buf];
  P.dst = buf_ptr(p, op.dst.buf) + (size_t)op.dst.c_off * db.elem_bytes;
  P.dst_ld = db.C;
  P.dst_rows_per_img = db.rows_per_img;
  P.dst_row_off = op.dst_row_off;
  P.out_f32 = op.out_f32;
  P.cout = op.dst.C;
  P.cout_store = op.out_f32 ? round_up(op.dst.C, 4) : cpad8(op.dst.C);
  // valid K16 steps per k-block: TMA-fed sources are padded to 64-channel blocks one by one, everything else is
  // one run whose tail is padding
  for (int kb = 0; kb < 32; kb++) P.kb_kv[kb] = BK / 16;
  if (op.a_tma && !op.dw_fused) {
    int kb = 0;
    for (int i = 0; i < op.nseg; i++) {
      const int blocks = op.seg_kpad[i] / BK, c = cpad8(op.src[i].C);
      for (int j = 0; j < blocks; j++, kb++)
        if (kb < 32) P.kb_kv[kb] = (uint8_t)std::max(1, std::min(BK / 16, (c - j * BK + 15) / 16));
    }
  } else if (!op.patch) {
    const int last = op.K_pad / BK - 1;
    if (last >= 0 && last < 32) P.kb_kv[last] = (uint8_t)std::max(1, std::min(BK / 16, (op.K - last * BK + 15) / 16));
  }
  const uint8_t* wbase = p->d_weights + cw.info.blob_offset;
  P.w = reinterpret_cast<const act_t*>(wbase);
  P.bias = reinterpret_cast<const float*>(wbase + (size_t)op.N_pad * op.K_pad * 2);
  if (op.has_res) {
    const Buf& rb = p->bufs[op.res.buf];
    P.res = reinterpret_cast<const act_t*>(buf_ptr(p, op.res.buf)) + op.res.c_off;
    P.res_ld = rb.C;
  }
  P.act = op.act;
  P.BN = op.BN;
  P.stages = op.stages;
  P.a_tma = op.a_tma;
  P.tmem_cols = tmem_cols_for(op.pair ? 2 * op.BN : op.BN);
  P.n_tiles = op.N_pad / op.BN;
  P.total_tiles = ((P.M + BM - 1) / BM) * P.n_tiles;
  auto magic = [](int d, uint32_t& mul, uint32_t& shr) {
    if (d <= 1) {
      mul = 0;
      shr = 0;
      return;
    }
    int lg = 0;
    while ((1u << lg) < (uint32_t)d) lg++;
    int pshift = 31 + lg;
    mul = (uint32_t)((((unsigned long long)1 << pshift) + (unsigned)d - 1) / (unsigned)d);
    shr = (uint32_t)(pshift - 32);
  };
  magic(P.hw_out, P.hw_mul, P.hw_shr);
  magic(P.Wout, P.w_mul, P.w_shr);
  magic(P.per_tap, P.pt_mul, P.pt_shr);
  P.tile2d = op.tile2d;
  P.a_region_bytes = op.stages * A_STAGE_BYTES;
  P.b_resident = op.b_resident;
  P.b_slots = op.b_resident ? P.num_kb : op.stages;
  P.c_bufs = op.c_bufs;
  // measured: alternate tiles win except on patch layers with >= 32 output channels
  P.alt_epilogue = (getenv("YB_NO_ALT_EPI") || (op.patch && op.BN >= 32) || op.pair) ? 0 : 1;
  P.s2d = op.s2d;
  P.c_s2d = (!op.out_f32 && p->bufs[op.dst.buf].s2d) ? 1 : 0;
  if (op.patch) {
    const int C = op.s2d ? 64 : op.seg_kpad[0];
    P.patch = 1;
    P.cblk = std::min(C, 64);
    P.ncb = op.s2d == 2 ? 4 : (C + 63) / 64;
    P.a_layout = P.cblk == 64 ? 2 : P.cblk == 32 ? 4 : P.cblk == 16 ? 6 : 0;
    P.pair = op.pair;
    P.patch_creal = op.s2d == 2 ? 256 : op.s2d ? 64 : op.src[0].C;
    P.patch_tx_bytes = PP_H * (op.pair ? PP2_W : PP_W) * P.cblk * 2;
    P.patch_stage_bytes = op.patch_stage_bytes;
    P.patch_stages = op.patch_stages;
    P.a_region_bytes = op.patch_stages * op.patch_stage_bytes;
    P.tiles_x = (op.Wout + (op.pair ? 2 : 1) * PT_W - 1) / ((op.pair ? 2 : 1) * PT_W);
    P.tiles_per_img = P.tiles_x * ((op.Hout + PT_H - 1) / PT_H);
    P.total_tiles = p->B * P.tiles_per_img * P.n_tiles;
    magic(P.tiles_per_img, P.tpi_mul, P.tpi_shr);
    magic(P.tiles_x, P.tx_mul, P.tx_shr);
  } else if (op.dw_fused) {
    const Op& dwop = p->ops[op.dw_op];
    const ConvW& dcw = p->convs[dwop.conv_index];
    P.dw = 1;
    P.dw_w = reinterpret_cast<const float*>(p->d_weights + dcw.info.blob_offset);
    P.dw_C = dwop.dst.C;
    P.dw_cp = cpad8(dwop.dst.C);
    P.dw_act = dwop.act;

    P.ncb = P.num_kb;
    P.patch_tx_bytes = PP_H * PP_W * 128;
    P.patch_stage_bytes = op.patch_stage_bytes;
    P.patch_stages = op.patch_stages;
    P.dw_patch_bytes = op.patch_stages * op.patch_stage_bytes;
    P.a_region_bytes = P.dw_patch_bytes + op.stages * A_STAGE_BYTES;
    P.tiles_x = (op.Wout + PT_W - 1) / PT_W;
    P.tiles_per_img = P.tiles_x * ((op.Hout + PT_H - 1) / PT_H);
    P.total_tiles = p->B * P.tiles_per_img * P.n_tiles;
    magic(P.tiles_per_img, P.tpi_mul, P.tpi_shr);
    magic(P.tiles_x, P.tx_mul, P.tx_shr);
  } else if (op.tile2d) {
    P.tiles_x = op.Wout / 16;
    P.tiles_per_img = P.tiles_x * (op.Hout / 8);
    magic(P.tiles_per_img, P.tpi_mul, P.tpi_shr);
    magic(P.tiles_x, P.tx_mul, P.tx_shr);
  }
  {
    static const int sleep_ns = getenv("YB_WAIT_SLEEP_NS") ? atoi(getenv("YB_WAIT_SLEEP_NS")) : 0;
    P.wait_sleep_ns = sleep_ns;
  }
  P.dw_dbg = getenv("YB_DW_DBG") ? atoi(getenv("YB_DW_DBG")) : 0;
  P.out_mode = op.out_f32 ? 1 : 0;
  P.A_total = p->A;
  P.nc = p->nc;
  P.nms_hdr = nullptr;
  P.nms_keys = nullptr;
  P.nms_cap = 0;
  P.nms_conf = INFINITY;
}

int launch_conv_tc(const yb_plan* p, const Op& op, cudaStream_t st, float* fused_out) {
  ConvParams P;
  fill_params(p, op, P);
  if (fused_out && op.head_part) {
    // head tails write the final (B, 4+nc, A) tensor: DFL box decode (part 1) / class sigmoid (part 2)
    P.out_mode = op.head_part == 1 ? 2 : 3;
    P.dst = fused_out;
    P.cout_store = op.head_part == 1 ? 64 : round_up(p->nc, 16);
    int lvl = op.dst_row_off == p->lvl_off[2] ? 2 : op.dst_row_off == p->lvl_off[1] ? 1 : 0;
    P.lvl_stride = p->lvl_stride[lvl];
    if (op.head_part == 2 && p->sink_keys) {   // class scores also feed the NMS candidate lists
      P.nms_hdr = p->sink_hdr;
      P.nms_keys = p->sink_keys;
      P.nms_cap = p->sink_cap;
      P.nms_conf = p->sink_conf;
    }
  }
  int grid = std::min(P.total_tiles, p->num_sms * op.occ);
#define YB_LAUNCH(ATMA, T2D, HEAD)                                                                       \
  do {                                                                                                   \
    if (p->act_f16)                                                                                      \
      YB_CUDA(launch_pdl(conv_gemm_tcgen05_kernel<ATMA, T2D, HEAD, true>, dim3(grid),                    \
                         dim3(ATMA == MODE_DW ? NUM_THREADS_DW : NUM_THREADS), op.smem_bytes, st,        \
                         P, op.tmap_b, op.tmap_a[0], op.tmap_a[1], op.tmap_a[2], op.tmap_a[3], op.tmap_c)); \
    else                                                                                                 \
      YB_CUDA(launch_pdl(conv_gemm_tcgen05_kernel<ATMA, T2D, HEAD, false>, dim3(grid),                   \
                         dim3(ATMA == MODE_DW ? NUM_THREADS_DW : NUM_THREADS), op.smem_bytes, st,        \
                         P, op.tmap_b, op.tmap_a[0], op.tmap_a[1], op.tmap_a[2], op.tmap_a[3], op.tmap_c)); \
  } while (0)
  const bool head = P.out_mode != 0;
  if (head) {
    if (P.a_tma) YB_LAUNCH(MODE_ATMA, false, true);
    else YB_LAUNCH(MODE_GATHER, false, true);
  } else if (P.dw) {
    YB_LAUNCH(MODE_DW, true, false);
  } else if (P.patch && P.pair) {
    YB_LAUNCH(MODE_PATCH2, true, false);
  } else if (P.patch) {
    YB_LAUNCH(MODE_PATCH, true, false);
  } else if (P.a_tma) {
    if (P.tile2d) YB_LAUNCH(MODE_ATMA, true, false);
    else YB_LAUNCH(MODE_ATMA, false, false);
  } else {
    if (P.tile2d) YB_LAUNCH(MODE_GATHER, true, false);
    else YB_LAUNCH(MODE_GATHER, false, false);
  }
#undef YB_LAUNCH
  count_launch();
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

int launch_conv_naive(const yb_plan* p, const Op& op, cudaStream_t st) {
  ConvParams P;
  fill_params(p, op, P);
  long long total = (long long)P.M * P.cout_store;
  int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (p->act_f16) conv_direct_check_kernel<true><<<(unsigned)blocks, threads, 0, st>>>(P);
  else conv_direct_check_kernel<false><<<(unsigned)blocks, threads, 0, st>>>(P);
  count_launch();
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

}  // namespace yb
