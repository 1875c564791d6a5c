// Plan builder: walks the YOLOv11 topology (reference nets/nn.py:151-270) once
// per (arch, B, H, W) and emits a static op list over NHWC bf16 buffers in one workspace arena.
// torch.cat / chunk / Upsample / residual adds of the reference do not exist here as ops:
//   - producers write into channel slices of their consumer's concat buffer,
//   - 1x1 consumers of an FPN concat walk their GEMM K dimension over several source slices,
//   - nearest-upsampled sources are gathered at (y>>1, x>>1) on load,
//   - residual adds happen in the producing GEMM's epilogue.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "yb_internal.h"

namespace yb {

namespace {

struct Builder {
  yb_plan* p;
  int err = 0;

  int new_buf(int H, int W, int C, int elem_bytes, const std::string& tag, int rows_per_img = -1) {
    Buf b;
    b.H = H;
    b.W = W;
    b.C = C;
    b.elem_bytes = elem_bytes;
    b.rows_per_img = rows_per_img < 0 ? H * W : rows_per_img;
    b.bytes = (size_t)p->B * b.rows_per_img * C * elem_bytes;
    b.tag = tag;
    p->bufs.push_back(b);
    return (int)p->bufs.size() - 1;
  }
  Slice whole(int buf) {
    Slice s;
    s.buf = buf;
    s.c_off = 0;
    s.C = p->bufs[buf].C;
    return s;
  }
  Slice sub(int buf, int off, int C) {
    Slice s;
    s.buf = buf;
    s.c_off = off;
    s.C = C;
    if (off % 8) {
      set_error("channel slice offset %d of %s is not a multiple of 8 (unsupported width)", off,
                p->bufs[buf].tag.c_str());
      err = YB_ERR_UNSUPPORTED;
    }
    return s;
  }
  void touch(const Slice& s, int op_index) {
    Buf& b = p->bufs[s.buf];
    b.first_def = std::min(b.first_def, op_index);
    b.last_use = std::max(b.last_use, op_index);
    b.touches.push_back(op_index);
  }
  int src_h(const Slice& s) { return p->bufs[s.buf].H << s.up; }
  int src_w(const Slice& s) { return p->bufs[s.buf].W << s.up; }

  int add_conv_record(Op& op, int cout, int cin, int groups, int wrapped, int kind) {
    ConvW cw;
    memset(&cw.info, 0, sizeof(cw.info));
    snprintf(cw.info.name, sizeof(cw.info.name), "%s", op.name.c_str());
    cw.info.cout = cout;
    cw.info.cin = cin;
    cw.info.ksize = op.k;
    cw.info.stride = op.stride;
    cw.info.groups = groups;
    cw.info.act = op.act;
    cw.info.wrapped = wrapped;
    cw.info.kind = kind;
    cw.op = (int)p->ops.size();
    p->convs.push_back(cw);
    return (int)p->convs.size() - 1;
  }

  // Dense convolution (reference `Conv` wrapper nn.py:28-39 or bare Conv2d tail nn.py:246,252).
  Slice conv(const std::string& name, std::vector<Slice> srcs, int cout, int k, int s, int act,
             const Slice* dst = nullptr, const Slice* res = nullptr, int wrapped = 1,
             int out_f32 = 0, int dst_row_off = 0) {
    Op op;
    op.kind = OP_CONV;
    op.name = name;
    op.k = k;
    op.stride = s;
    op.act = act;
    op.out_f32 = out_f32;
    op.dst_row_off = dst_row_off;
    op.nseg = (int)srcs.size();
    if (op.nseg > 4 || (k == 3 && op.nseg != 1)) {
      set_error("conv %s: unsupported source count", name.c_str());
      err = YB_ERR_UNSUPPORTED;
      return Slice();
    }
    op.Hin = src_h(srcs[0]);
    op.Win = src_w(srcs[0]);
    int cin = 0;
    for (int i = 0; i < op.nseg; i++) {
      op.src[i] = srcs[i];
      cin += srcs[i].C;
      if (src_h(srcs[i]) != op.Hin || src_w(srcs[i]) != op.Win) {
        set_error("conv %s: source %d has mismatching spatial size", name.c_str(), i);
        err = YB_ERR_ARG;
      }
    }
    int pad = k / 2;
    op.Hout = (op.Hin + 2 * pad - k) / s + 1;
    op.Wout = (op.Win + 2 * pad - k) / s + 1;
    if (dst) {
      op.dst = *dst;
      op.dst.C = cout;
    } else {
      int b = new_buf(op.Hout, op.Wout, cpad8(cout), 2, name);
      op.dst = whole(b);
      op.dst.C = cout;
    }
    if (res) {
      op.has_res = 1;
      op.res = *res;
    }
    // GEMM shape. a_tma: all sources are plain (not upsampled) and the conv is 1x1/stride 1.
    op.a_tma = (k == 1 && s == 1);
    for (int i = 0; i < op.nseg; i++)
      if (op.src[i].up) op.a_tma = 0;
    if (getenv("YB_NO_ATMA")) op.a_tma = 0;
    int K = 0, Kp = 0;
    if (op.a_tma) {
      for (int i = 0; i < op.nseg; i++) {
        op.seg_kpad[i] = round_up(op.src[i].C, 64);
        K += op.src[i].C;
        Kp += op.seg_kpad[i];
      }
    } else {
      int per_tap = 0;
      for (int i = 0; i < op.nseg; i++) {
        op.seg_kpad[i] = cpad8(op.src[i].C);
        per_tap += op.seg_kpad[i];
      }
      // 3x3 / stride-1 layers that will read TMA halo patches need every tap's K range to start on a
      // 64-channel block: channel counts that are no power of two <= 32 and no multiple of 64 (YOLO11x: 48,
      // 96) get their per-tap K padded ("tap-aligned" packing; the pad columns hold zero weights)
      int min_hw = 40;
      if (const char* e = getenv("YB_PATCH_MIN_HW")) min_hw = atoi(e);
      if (k == 3 && s == 1 && op.nseg == 1 && !op.src[0].up && !getenv("YB_NO_PATCH") && !getenv("YB_NO_TAP_ALIGN") &&
          op.Hout >= min_hw && op.Wout >= min_hw && per_tap > 32 && per_tap % 64 != 0) {
        op.seg_kpad[0] = round_up(per_tap, 64);
        per_tap = op.seg_kpad[0];
      }
      K = per_tap * k * k;
      Kp = round_up(K, 64);
    }
    op.K = K;
    op.K_pad = Kp;
    op.N_pad = round_up(cout, 16);
    int nt = (op.N_pad + 255) / 256;
    while (op.N_pad % (16 * nt)) nt++;
    op.BN = op.N_pad / nt;
    op.conv_index = add_conv_record(op, cout, cin, 1, wrapped, 1);
    int idx = (int)p->ops.size();
    for (int i = 0; i < op.nseg; i++) touch(op.src[i], idx);
    touch(op.dst, idx);
    if (op.has_res) touch(op.res, idx);
    p->ops.push_back(op);
    return op.dst;
  }

  // Depthwise 3x3 (+BN folded, optional SiLU): head cls branches nn.py:248,250 and attention pe nn.py:109.
  Slice dwconv(const std::string& name, const Slice& src, int C, int act, const Slice* dst = nullptr,
               int gsz = 0, int gstride = 0, int goff = 0, int add = 0) {
    Op op;
    op.kind = OP_DW;
    op.name = name;
    op.k = 3;
    op.stride = 1;
    op.act = act;
    op.nseg = 1;
    op.src[0] = src;
    op.Hin = op.Hout = src_h(src);
    op.Win = op.Wout = src_w(src);
    if (dst) {
      op.dst = *dst;
      op.dst.C = C;
    } else {
      int b = new_buf(op.Hout, op.Wout, cpad8(C), 2, name);
      op.dst = whole(b);
      op.dst.C = C;
    }
    op.dw_gsz = gsz ? gsz : C;
    op.dw_gstride = gsz ? gstride : C;
    op.dw_goff = goff;
    op.dw_add = add;
    if (C % 8) {
      set_error("depthwise conv %s: %d channels not a multiple of 8", name.c_str(), C);
      err = YB_ERR_UNSUPPORTED;
    }
    op.conv_index = add_conv_record(op, C, 1, C, 1, 2);
    int idx = (int)p->ops.size();
    touch(op.src[0], idx);
    touch(op.dst, idx);
    p->ops.push_back(op);
    return op.dst;
  }

  // Residual bottleneck nn.py:42-49: x + conv2(conv1(x)); the add sits in conv2's epilogue.
  // (add = false: the caller folds the `x +` into the weights of the 1x1 conv that consumes x and this output)
  Slice residual(const std::string& name, const Slice& x, double e, const Slice* dst = nullptr, bool add = true) {
    int c = x.C;
    int h = (int)(c * e);
    Slice t = conv(name + ".conv1", {x}, h, 3, 1, 1);
    return conv(name + ".conv2", {t}, c, 3, 1, 1, dst, add ? &x : nullptr);
  }

  // C3k nn.py:52-63.
  Slice csp_module(const std::string& name, const Slice& x, int out_ch, const Slice* dst) {
    int half = out_ch / 2;
    int cm = new_buf(src_h(x), src_w(x), 2 * half, 2, name + ".cat");
    Slice t1 = conv(name + ".conv1", {x}, half, 1, 1, 1);
    Slice d2 = sub(cm, half, half);
    conv(name + ".conv2", {x}, half, 1, 1, 1, &d2);
    Slice r1 = residual(name + ".res_m.0", t1, 1.0);
    Slice d1 = sub(cm, 0, half);
    residual(name + ".res_m.1", r1, 1.0, &d1);
    return conv(name + ".conv3", {whole(cm)}, out_ch, 1, 1, 1, dst);
  }

  // C3k2 nn.py:66-80.
  Slice csp(const std::string& name, std::vector<Slice> srcs, int out_ch, int n, int use_csp, int r,
            const Slice* dst = nullptr) {
    int c = out_ch / r;
    int H = src_h(srcs[0]), W = src_w(srcs[0]);
    // (Dense parts - a buffer of its own for every part narrower than 64 bytes per pixel, conv1 storing its two
    // halves to two tensors, conv2 walking K over 2 + n sources - were built and measured in round 2: the 2.3x DRAM
    // over-fetch of net.p2.1's 16-channel slices goes away, but the block gets 111 us SLOWER: 32-byte rows are bound
    // by the TMA engine's row rate on the read and on the write side, not by DRAM.  A milder split - conv1's two
    // halves together, every bottleneck output dense, conv2 over 1 + n sources - gained 40 us while the bottleneck
    // still read its residual, nothing once that read was folded away (below).  Both dropped)
    // The last bottleneck's `x + f(x)` (nn.py:49) is only read by conv2, which also reads x itself as the part in
    // front of it: W_x x + W_m (x + f) = (W_x + W_m) x + W_m f.  The bottleneck stores f alone (no residual read:
    // 420 MB of the 735 MB net.p2.1.res_m.0.conv2 pulled through DRAM) and conv2's weights for x are summed when
    // the blob is packed (yb_plan_pack_conv).  YB_NO_RES_FOLD=1 keeps the add in the bottleneck's epilogue.
    const bool fold = !use_csp && n >= 1 && getenv("YB_NO_RES_FOLD") == nullptr;
    auto mark_fold = [&]() {
      if (!fold) return;
      Op& o = p->ops.back();
      o.wfold_dst = n * c;
      o.wfold_src = (n + 1) * c;
      o.wfold_n = c;
    };
    int cat = new_buf(H, W, (2 + n) * c, 2, name + ".cat");
    Slice d = sub(cat, 0, 2 * c);
    conv(name + ".conv1", srcs, 2 * c, 1, 1, 1, &d);
    for (int i = 0; i < n; i++) {
      Slice in = sub(cat, (1 + i) * c, c);
      Slice di = sub(cat, (2 + i) * c, c);
      std::string mn = name + ".res_m." + std::to_string(i);
      if (use_csp)
        csp_module(mn, in, c, &di);
      else
        residual(mn, in, 0.5, &di, !(fold && i == n - 1));
    }
    Slice o2 = conv(name + ".conv2", {whole(cat)}, out_ch, 1, 1, 1, dst);
    mark_fold();
    return o2;
  }

  // SPPF nn.py:83-94: the 5/9/13 windows equal the three cascaded 5x5 pools (max, -inf padding).
  Slice spp(const std::string& name, const Slice& x, int C) {
    int half = C / 2;
    int sp = new_buf(src_h(x), src_w(x), 4 * half, 2, name + ".cat");
    Slice d = sub(sp, 0, half);
    conv(name + ".conv1", {x}, half, 1, 1, 1, &d);
    Op op;
    op.kind = OP_POOL;
    op.name = name + ".res_m";
    op.nseg = 1;
    op.src[0] = d;
    op.dst = sub(sp, half, 3 * half);
    op.Hin = op.Hout = src_h(x);
    op.Win = op.Wout = src_w(x);
    if (half % 8) {
      set_error("SPP %s: %d channels not a multiple of 8", name.c_str(), half);
      err = YB_ERR_UNSUPPORTED;
    }
    int idx = (int)p->ops.size();
    touch(op.src[0], idx);
    p->ops.push_back(op);
    return conv(name + ".conv2", {whole(sp)}, C, 1, 1, 1);
  }

  // PSABlock nn.py:126-136 with Attention nn.py:97-123, updating y in place.
  void psa_block(const std::string& name, const Slice& y, int ch, int heads) {
    int dh = ch / heads, dk = dh / 2;
    if (dh != 64 || dk != 32) {
      set_error("attention %s: head dim %d unsupported (kernel is specialised for 64/32)",
                name.c_str(), dh);
      err = YB_ERR_UNSUPPORTED;
      return;
    }
    Slice qkv = conv(name + ".conv1.qkv", {y}, ch + dk * heads * 2, 1, 1, 0);
    int ao = new_buf(src_h(y), src_w(y), ch, 2, name + ".attn");
    Op op;
    op.kind = OP_ATTN;
    op.name = name + ".conv1.attn";
    op.nseg = 1;
    op.src[0] = qkv;
    op.dst = whole(ao);
    op.Hin = op.Hout = src_h(y);
    op.Win = op.Wout = src_w(y);
    op.heads = heads;
    op.dk = dk;
    op.dh = dh;
    op.scale = 1.0f / sqrtf((float)dk);
    int idx = (int)p->ops.size();
    touch(op.src[0], idx);
    touch(op.dst, idx);
    p->ops.push_back(op);
    Slice aos = whole(ao);
    dwconv(name + ".conv1.conv1", qkv, ch, 0, &aos, dh, 2 * dk + dh, 2 * dk, 1);
    conv(name + ".conv1.conv2", {aos}, ch, 1, 1, 0, &y, &y);
    Slice f1 = conv(name + ".conv2.0", {y}, 2 * ch, 1, 1, 1);
    conv(name + ".conv2.1", {f1}, ch, 1, 1, 0, &y, &y);
  }

  // C2PSA nn.py:139-148.
  Slice psa(const std::string& name, const Slice& x, int C, int n) {
    int ps = new_buf(src_h(x), src_w(x), C, 2, name + ".cat");
    Slice d = whole(ps);
    conv(name + ".conv1", {x}, C, 1, 1, 1, &d);
    Slice y = sub(ps, C / 2, C / 2);
    for (int i = 0; i < n; i++) psa_block(name + ".res_m." + std::to_string(i), y, C / 2, C / 128);
    return conv(name + ".conv2", {whole(ps)}, C, 1, 1, 1);
  }
};

}  // namespace

int build_plan(yb_plan* p) {
  const int* w = p->arch.width;
  const int* d = p->arch.depth;
  const int* c = p->arch.csp;
  Builder b;
  b.p = p;
  p->nc = p->arch.num_classes;
  p->no = 64 + p->nc;
  if (p->H % 32 || p->W % 32 || p->H <= 0 || p->W <= 0) {
    set_error("input size %dx%d must be a positive multiple of 32", p->H, p->W);
    return YB_ERR_ARG;
  }
  if (w[0] != 3 || w[5] % 128 || p->nc < 1) {
    set_error("unsupported architecture (width[0]=%d, width[5]=%d, nc=%d)", w[0], w[5], p->nc);
    return YB_ERR_UNSUPPORTED;
  }
  for (int i = 0; i < 3; i++) {
    p->lvl_h[i] = p->H >> (3 + i);
    p->lvl_w[i] = p->W >> (3 + i);
    p->lvl_stride[i] = (float)(8 << i);
  }
  p->lvl_off[0] = 0;
  p->lvl_off[1] = p->lvl_h[0] * p->lvl_w[0];
  p->lvl_off[2] = p->lvl_off[1] + p->lvl_h[1] * p->lvl_w[1];
  p->A = p->lvl_off[2] + p->lvl_h[2] * p->lvl_w[2];

  // ---- stem: net.p1.0, Conv(3 -> w1, 3, s2) reading the caller's NCHW image (nn.py:161) ----
  Slice p1;
  {
    Op op;
    op.kind = OP_STEM;
    op.name = "net.p1.0";
    op.k = 3;
    op.stride = 2;
    op.act = 1;
    op.Hin = p->H;
    op.Win = p->W;
    op.Hout = p->H / 2;
    op.Wout = p->W / 2;
    int buf = b.new_buf(op.Hout, op.Wout, cpad8(w[1]), 2, op.name);
    op.dst = b.whole(buf);
    op.dst.C = w[1];
    op.conv_index = b.add_conv_record(op, w[1], 3, 1, 1, 0);
    b.touch(op.dst, 0);
    p->ops.push_back(op);
    p1 = op.dst;
  }
  // ---- backbone nn.py:163-175 ----
  Slice t = b.conv("net.p2.0", {p1}, w[2], 3, 2, 1);
  Slice p2 = b.csp("net.p2.1", {t}, w[3], d[0], c[0], 4);
  t = b.conv("net.p3.0", {p2}, w[3], 3, 2, 1);
  Slice P3 = b.csp("net.p3.1", {t}, w[4], d[1], c[0], 4);
  t = b.conv("net.p4.0", {P3}, w[4], 3, 2, 1);
  Slice P4 = b.csp("net.p4.1", {t}, w[4], d[2], c[1], 2);
  t = b.conv("net.p5.0", {P4}, w[5], 3, 2, 1);
  t = b.csp("net.p5.1", {t}, w[5], d[3], c[1], 2);
  t = b.spp("net.p5.2", t, w[5]);
  Slice P5 = b.psa("net.p5.3", t, w[5], d[4]);
  if (b.err) return b.err;
  // ---- head nn.py:240-257: tails write fp32 logits (B, A, 64+nc) at the level's anchor offset.  The three ops of
  // a box tower and the three kernels of a class tower depend on their FPN level only: each level's head is emitted
  // right behind the tensor it reads, so that the stream lanes below can run it next to the rest of the neck
  // (default with one lane, and with YB_HEAD_LAST=1: all heads after the neck, the order of the reference's module list).
  int no_p = 64 + round_up(p->nc, 4);
  p->logits_buf = b.new_buf(1, p->A, no_p, 4, "head.logits", p->A);
  int Bc = std::max(64, w[3] / 4);
  int Cc = std::max(std::max(80, w[3]), p->nc);
  auto head_level = [&](int i, const Slice& x) {
    std::string bi = "head.box." + std::to_string(i);
    std::string ci = "head.cls." + std::to_string(i);
    Slice b0 = b.conv(bi + ".0", {x}, Bc, 3, 1, 1);
    Slice b1 = b.conv(bi + ".1", {b0}, Bc, 3, 1, 1);
    Slice dbox = b.sub(p->logits_buf, 0, 64);
    b.conv(bi + ".2", {b1}, 64, 1, 1, 0, &dbox, nullptr, 0, 1, p->lvl_off[i]);
    p->ops.back().head_part = 1;
    Slice c0 = b.dwconv(ci + ".0", x, x.C, 1);
    Slice c1 = b.conv(ci + ".1", {c0}, Cc, 1, 1, 1);
    Slice c2 = b.dwconv(ci + ".2", c1, Cc, 1);
    Slice c3 = b.conv(ci + ".3", {c2}, Cc, 1, 1, 1);
    Slice dcls = b.sub(p->logits_buf, 64, p->nc);
    b.conv(ci + ".4", {c3}, p->nc, 1, 1, 0, &dcls, nullptr, 0, 1, p->lvl_off[i]);
    p->ops.back().head_part = 2;
  };
  // stream lanes (below): on by default where the kernels of one layer do not fill the GPU - measured on the B200, forward
  // of YOLO11n at 640 x 640: B = 1 -18 %, 4 -12 %, 16 -5 %, 32 and up +1 % (YOLO11x: B = 1 -15 %, 8 -5 %, 16 -1 %)
  const int lanes_env = getenv("YB_LANES") ? atoi(getenv("YB_LANES"))
                                           : ((long long)p->B * p->H * p->W <= 16LL * 640 * 640 ? 4 : 1);
  const bool head_last = getenv("YB_HEAD_LAST") ? atoi(getenv("YB_HEAD_LAST")) != 0 : lanes_env <= 1;
  // ---- neck nn.py:203-209 ----
  Slice P5u = P5;
  P5u.up = 1;
  Slice T4 = b.csp("fpn.h1", {P5u, P4}, w[4], d[5], c[0], 2);
  Slice T4u = T4;
  T4u.up = 1;
  Slice N3 = b.csp("fpn.h2", {T4u, P3}, w[3], d[5], c[0], 2);
  if (!head_last) head_level(0, N3);
  Slice h3 = b.conv("fpn.h3", {N3}, w[3], 3, 2, 1);
  Slice N4 = b.csp("fpn.h4", {h3, T4}, w[4], d[5], c[0], 2);
  if (!head_last) head_level(1, N4);
  Slice h5 = b.conv("fpn.h5", {N4}, w[4], 3, 2, 1);
  Slice N5 = b.csp("fpn.h6", {h5, P5}, w[5], d[5], c[1], 2);
  if (b.err) return b.err;
  if (head_last) {
    head_level(0, N3);
    head_level(1, N4);
  }
  head_level(2, N5);
  if (b.err) return b.err;
  {
    Op op;
    op.kind = OP_DECODE;
    op.name = "head.decode";
    int idx = (int)p->ops.size();
    Slice lg = b.whole(p->logits_buf);
    b.touch(lg, idx);
    b.touch(lg, idx + 1);  // keep alive for yb_forward_raw's copy-out
    p->ops.push_back(op);
  }

  // ---- depthwise 3x3 -> 1x1 fusion (head cls branches, nn.py:248-251): the depthwise output is
  // produced tile by tile inside the 1x1 conv's kernel when that conv is its only consumer.
  if (!getenv("YB_NO_DWFUSE")) {
    for (size_t i = 0; i + 1 < p->ops.size(); i++) {
      Op& d = p->ops[i];
      Op& c = p->ops[i + 1];
      if (d.kind != OP_DW || c.kind != OP_CONV) continue;
      if (d.dw_add || d.dw_goff || d.dw_gsz != d.dst.C || d.src[0].up) continue;
      if (c.k != 1 || c.stride != 1 || c.nseg != 1 || !c.a_tma || c.out_f32 || c.has_res) continue;
      if (c.src[0].buf != d.dst.buf || c.src[0].c_off != d.dst.c_off || c.src[0].C != d.dst.C) continue;
      if (p->bufs[d.dst.buf].last_use != (int)i + 1) continue;
      const int max_kb = getenv("YB_DWFUSE_MAX_KB") ? atoi(getenv("YB_DWFUSE_MAX_KB")) : 4;
      const int max_nt = getenv("YB_DWFUSE_MAX_NT") ? atoi(getenv("YB_DWFUSE_MAX_NT")) : 1;
      if (c.N_pad / c.BN > max_nt || c.K_pad / 64 > max_kb || c.Hout < 20) continue;
      c.dw_fused = 1;
      c.dw_op = (int)i;
      d.fused_away = 1;
      // the consumer's kernel now reads the depthwise conv's INPUT: keep that buffer alive through op i + 1,
      // or the arena may place the consumer's own output (first defined at i + 1) on top of it
      Buf& in = p->bufs[d.src[0].buf];
      in.last_use = std::max(in.last_use, (int)i + 1);
      in.touches.push_back((int)i + 1);
    }
  }
  if (getenv("YB_NO_FUSE_DECODE")) p->fuse_decode = 0;
  // ---- weight blob layout ----
  size_t off = 0;
  for (auto& cw : p->convs) {
    const Op& op = p->ops[cw.op];
    size_t bytes = 0;
    if (cw.info.kind == 1)
      bytes = (size_t)op.N_pad * op.K_pad * 2 + (size_t)op.N_pad * 4;
    else if (cw.info.kind == 0)
      bytes = (size_t)27 * cpad8(cw.info.cout) * 4 + (size_t)cpad8(cw.info.cout) * 4;
    else
      bytes = (size_t)9 * cpad8(cw.info.cout) * 4 + (size_t)cpad8(cw.info.cout) * 4;
    cw.info.blob_offset = off;
    cw.info.blob_bytes = bytes;
    off += (bytes + 255) / 256 * 256;
  }
  p->weight_bytes = off;

  // ---- space-to-depth sources for stride-2 3x3 convs ------------------------------------------------------
  // A stride-2 conv reads every input pixel 2.25 times through the im2col gather.  Where its source tensor has no
  // other reader, the producer stores it as (H/2, W/2, 4C) - channel block = pixel parity - and the conv becomes a
  // 2x2-tap stride-1 conv over those blocks on the TMA halo-patch path: every original tap (ky, kx) is one parity
  // block at one of four patch shifts, the weights keep their packing.  Producers that can store this way: the
  // stem (plain stores; 16 channels) and a 1x1 conv on 8 x 16 pixel tiles (the same staging tile through a 5-D TMA
  // map; 64 channels).  YOLO11n: net.p1 -> net.p2.0 and net.p2 -> net.p3.0.
  if (!getenv("YB_NO_S2D") && !getenv("YB_STEM_DIRECT") && !getenv("YB_NO_PATCH") && !getenv("YB_NO_RESIDENT")) {   // (needs resident weights)
    int min_hw = 40;
    if (const char* e = getenv("YB_PATCH_MIN_HW")) min_hw = atoi(e);
    const int max_c = getenv("YB_S2D_MAXC") ? atoi(getenv("YB_S2D_MAXC")) : 64;
    for (size_t i = 0; i < p->ops.size(); i++) {
      Op& o = p->ops[i];
      if (o.kind != OP_CONV || o.k != 3 || o.stride != 2 || o.nseg != 1 || o.src[0].up || o.out_f32) continue;
      Buf& sb = p->bufs[o.src[0].buf];
      if (o.src[0].c_off != 0 || o.src[0].C != sb.C || (sb.C != 16 && sb.C != 64) || sb.C > max_c || sb.H % 2 || sb.W % 2) continue;
      if (o.Hout < min_hw || o.Wout < min_hw || o.N_pad != o.BN) continue;
      if (sb.touches.size() != 2 || sb.touches[1] != (int)i) continue;
      const Op& pr = p->ops[sb.touches[0]];
      if (sb.C == 16) {
        if (pr.kind != OP_STEM) continue;
      } else {
        if (pr.kind != OP_CONV || !pr.a_tma || pr.out_f32 || pr.dw_fused || pr.N_pad != pr.BN || pr.dst.c_off != 0 ||
            pr.dst.C != sb.C || pr.Hout % 8 || pr.Wout % 16 || getenv("YB_NO_TILE2D"))
          continue;
      }
      sb.s2d = 1;
      o.s2d = sb.C == 16 ? 1 : 2;
    }
  }

  // ---- stream lanes (YB_LANES=n; default 4 for small batches, 1 = one stream otherwise) ---------------------
  // Branches of the graph that do not depend on each other (the two towers of every head level against the
  // rest of the neck, the two 1x1 convs in front of a C3k) are enqueued on separate streams.  Dependencies
  // come from the (buffer, channel range, anchor rows) every op reads and writes; an op continues the lane of
  // its latest producer if that producer is still the lane's last op, otherwise it forks onto a lane whose last
  // op is complete anyway (no false ordering), else onto the lane that has been idle longest.
  // Small batches (kernels of a few dozen CTAs, latency-bound): the batch-1 forward of YOLO11n drops from 0.62 to
  // 0.50 ms.  Large batches (YOLO11n, B = 256: 45.0-45.2 k img/s with one lane, 44.9-45.4 k with four): nothing -
  // the persistent kernels already fill every SM slot, a dependent kernel launched early by PDL takes the slots
  // its predecessor frees before another lane's kernel can, and running two chains on half the slots each sums
  // to the same time.
  const size_t n_ops = p->ops.size();
  const int nl = std::max(1, std::min(YB_MAX_LANES, lanes_env));
  p->num_lanes = nl;
  struct Acc { int buf, c0, c1, rows, write; };
  auto accesses = [&](size_t i) {
    std::vector<Acc> a;
    const Op& o = p->ops[i];
    auto rd = [&](const Slice& sl) { if (sl.buf >= 0) a.push_back({sl.buf, sl.c_off, sl.c_off + cpad8(sl.C), -1, 0}); };
    auto wr = [&](const Slice& sl, int rows) { if (sl.buf >= 0) a.push_back({sl.buf, sl.c_off, sl.c_off + cpad8(sl.C), rows, 1}); };
    switch (o.kind) {
      case OP_STEM: wr(o.dst, -1); break;
      case OP_CONV:
        for (int k = 0; k < o.nseg; k++) rd(o.src[k]);
        if (o.dw_fused) rd(p->ops[o.dw_op].src[0]);
        if (o.has_res) rd(o.res);
        wr(o.dst, o.out_f32 ? o.dst_row_off : -1);
        break;
      case OP_DW:
        rd(o.src[0]);
        if (o.dw_add) rd(o.dst);
        wr(o.dst, -1);
        break;
      case OP_POOL: case OP_ATTN: rd(o.src[0]); wr(o.dst, -1); break;
      case OP_DECODE: { Slice lg; lg.buf = p->logits_buf; lg.c_off = 0; lg.C = p->bufs[p->logits_buf].C; rd(lg); break; }
    }
    return a;
  };
  std::vector<std::vector<Acc>> acc(n_ops);
  for (size_t i = 0; i < n_ops; i++) acc[i] = accesses(i);
  // reach[i][j]: op j is complete before op i starts (stream order of a lane + cross-lane events, transitive)
  std::vector<std::vector<uint8_t>> reach(n_ops, std::vector<uint8_t>(n_ops, 0));
  int lane_tail[YB_MAX_LANES];
  for (int l = 0; l < YB_MAX_LANES; l++) lane_tail[l] = -1;
  for (size_t i = 0; i < n_ops; i++) {
    Op& o = p->ops[i];
    std::vector<int> deps;
    for (size_t j = 0; j < i; j++) {
      bool hit = false;
      for (const Acc& x : acc[i])
        for (const Acc& y : acc[j]) {
          if (x.buf != y.buf || !(x.write || y.write) || x.c1 <= y.c0 || y.c1 <= x.c0) continue;
          if (x.write && y.write && x.rows >= 0 && y.rows >= 0 && x.rows != y.rows) continue;  // head tails of different levels
          hit = true;
        }
      if (hit) deps.push_back((int)j);
    }
    int lane = 0;
    if (nl > 1 && !deps.empty()) {
      const int m = deps.back();
      if (lane_tail[p->ops[m].lane] == m) {
        lane = p->ops[m].lane;
      } else {
        // a lane whose last op is complete anyway once the dependencies are (no false ordering), else the lane
        // that has been idle longest
        std::vector<uint8_t> done(n_ops, 0);
        for (int dj : deps) {
          done[dj] = 1;
          for (size_t k = 0; k < n_ops; k++) done[k] |= reach[dj][k];
        }
        lane = -1;
        for (int l = 0; l < nl && lane < 0; l++)
          if (lane_tail[l] < 0 || done[lane_tail[l]]) lane = l;
        if (lane < 0) {
          lane = 0;
          for (int l = 1; l < nl; l++)
            if (lane_tail[l] < lane_tail[lane]) lane = l;
        }
      }
    }
    o.lane = lane;
    const int prev = lane_tail[lane];
    if (prev >= 0) {
      reach[i] = reach[prev];
      reach[i][prev] = 1;
    }
    // cross-lane waits: the latest dependency on every other lane, unless it is already implied
    for (int l = 0; l < nl; l++) {
      if (l == lane) continue;
      int latest = -1;
      for (int dj : deps)
        if (p->ops[dj].lane == l) latest = dj;
      if (latest < 0 || reach[i][latest]) continue;
      o.xdeps.push_back(latest);
      p->ops[latest].signal = 1;
      for (size_t k = 0; k < n_ops; k++) reach[i][k] |= reach[latest][k];
      reach[i][latest] = 1;
    }
    lane_tail[lane] = (int)i;
  }

  // ---- arena assignment with lifetime reuse (first-fit over live intervals) ----
  // Two buffers may share memory only if the earlier one is dead in op order AND every op that touches it is
  // complete before any op that touches the later one starts (with one lane the second condition is implied).
  auto ordered = [&](const Buf& early, const Buf& late) {
    for (int y : late.touches)
      for (int x : early.touches) {
        if (x >= y) return false;
        if (y >= (int)n_ops) continue;   // (the copy-out behind the last op runs after the lanes have joined)
        if (!reach[y][x]) return false;
      }
    return true;
  };
  bool reuse = getenv("YB_NO_REUSE") == nullptr;
  std::vector<int> order(p->bufs.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
  std::sort(order.begin(), order.end(), [&](int a, int bb) {
    return p->bufs[a].first_def < p->bufs[bb].first_def;
  });
  std::vector<int> placed;
  size_t top = 0;
  for (int bi : order) {
    Buf& nb = p->bufs[bi];
    size_t sz = (nb.bytes + 1023) / 1024 * 1024;
    size_t cand = 0;
    if (reuse) {
      // collect memory intervals of buffers whose lifetime overlaps, then first-fit
      std::vector<std::pair<size_t, size_t>> busy;
      for (int pi : placed) {
        const Buf& ob = p->bufs[pi];
        if ((ob.last_use < nb.first_def && ordered(ob, nb)) || (nb.last_use < ob.first_def && ordered(nb, ob))) continue;
        busy.push_back({ob.offset, ob.offset + (ob.bytes + 1023) / 1024 * 1024});
      }
      std::sort(busy.begin(), busy.end());
      for (auto& iv : busy) {
        if (cand + sz <= iv.first) break;
        cand = std::max(cand, iv.second);
      }
    } else {
      cand = top;
    }
    nb.offset = cand;
    top = std::max(top, cand + sz);
    placed.push_back(bi);
  }
  p->workspace_bytes = top;
  return YB_OK;
}

}  // namespace yb
