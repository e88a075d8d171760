// Plan = static layer schedule + activation arena + packed weights for one
// (arch, heads, batch, resolution) configuration.  The topology below follows
// the reference definitions (citations relative to
// /root/reference/src/lib/models/networks):
//   pose_dla_dcn.py:227-322  DLA-34 base    (dla34(): levels [1,1,1,2,2,1],
//                                            channels [16,32,64,128,256,512], :340-346)
//   pose_dla_dcn.py:171-224  Tree / Root / BasicBlock wiring and concat order
//   pose_dla_dcn.py:392-443  IDAUp / DLAUp  (proj DCN -> depthwise ConvT up -> + skip -> node DCN)
//   pose_dla_dcn.py:457-570  DLASeg: ida_up, heads (3x3 -> [GN] -> ReLU -> 1x1), convGRU routing
//   DCNv2/dcn_v2.py:97-128   DCN = 3x3 conv -> 27 ch (18 offsets + 9 mask logits) + deformable 3x3
//   convGRU.py:20-94, GN.py:4-9
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"

namespace cp {

thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
thread_local long long g_launch_counter = 0;
thread_local int g_pdl = 0;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

namespace {

struct Act {
  size_t off = 0;  // floats into the activation arena
  int C = 0, H = 0, W = 0;
  int stride = 0;  // pixel stride in floats
  int ext = -1;    // >= 0: external NCHW input index (0 images, 1 pre_img, 2 pre_hm, 3 pre_hm_hp)
};

enum OpType { OP_IGEMM, OP_MAXPOOL, OP_UPADD, OP_GN_RELU, OP_GRU };

struct Op {
  OpType type = OP_IGEMM;
  std::string name;
  // igemm
  Act src[4];
  int nsrc = 0;
  int mode = IGEMM_NHWC_VEC;
  Act out;
  int out_head = -1;  // >= 0: result goes to head output `out_head` as NCHW
  bool has_res = false, res_after_relu = false, relu = false;
  Act res;
  int kh = 1, kw = 1, stride = 1, pad = 0, Cin = 0, Cout = 0, CoutPad = 0, Kpad = 0;
  size_t w_off = 0, b_off = 0;
  int w_ld = 0;            // leading dimension of the fp32 packed weight matrix
  bool use_umma = false;   // run on the tcgen05 gather kernel
  bool use_tma = false;    // run on the TMA-fed tcgen05 kernel (conv_tma.cu)
  bool use_dcn_tma = false;   // deformable conv on the TMA-staged tcgen05 kernel (dcn_tma.cu)
  std::vector<int> head_children;   // merged heads 3x3 conv: indices of the per-head 1x1 ops that read its slices
  bool fuse_heads = false;          // ... which run inside its epilogue (conv_tma.cu), never touching HBM
  bool fused_away = false;          // this 1x1 op is computed by its parent's epilogue
  int tma_cslab = 32;
  std::vector<unsigned char> tma_maps;   // 4 CUtensorMap, encoded once the arena exists
  size_t umma_off = 0;     // bytes into the plan's tensor-core weight-tile buffer
  Act om;
  // up-sample
  int f = 0;
  size_t upw_off = 0;
  bool has_skip = false;
  Act skip;
  // group-norm
  size_t gamma_off = 0, beta_off = 0;
  int groups = 0, chanOffset = 0;
  // gru
  Act gx, gh, gprev;
  bool first_step = false;
};

enum PackType { PACK_CONV, PACK_BIAS, PACK_UP, PACK_VEC };

struct PackJob {
  PackType type;
  std::string key;       // main tensor key
  std::string bias_key;  // conv bias key ("" = none)
  std::string bn;        // BN prefix ("" = none)
  int Cout = 0, Cin = 0, kh = 0, kw = 0, CoutPad = 0, Kpad = 0, ld = 0, colOff = 0;
  size_t dst = 0;        // floats into the weight arena (weights / bias / up / vec)
  size_t scale = 0;      // scratch scale vector (PACK_BIAS writes, PACK_CONV reads); 0 = none
  bool use_scale = false;
};

struct WRef {
  const float* p;
  int64_t n;
};

}  // namespace
}  // namespace cp

using namespace cp;

struct cp_plan {
  cp_config cfg;
  std::vector<std::string> head_names;
  int B = 0, H = 0, W = 0;
  std::vector<Op> ops;
  std::vector<PackJob> jobs;
  float* act = nullptr;
  size_t act_floats = 0;
  float* wts = nullptr;
  size_t w_floats = 0;
  std::vector<Act> head_bufs;  // plan-owned NCHW head logits (cp_infer)
  bool loaded = false;
  int launches = 0;
  void* decode_ws = nullptr;
  size_t decode_ws_bytes = 0;
  double* gn_stats = nullptr;
  int prec = -1;                 // -1 fp32 CUDA cores, 0 bf16 tcgen05, 1 tf32x3 tcgen05, 2 tf32 (TMA) + tf32x3 elsewhere
  int min_tc_cin = 32;           // ops with fewer input channels stay on the CUDA-core kernels (CP_MIN_TC_CIN overrides)
  bool no_fuse_heads = false;    // CP_NO_FUSE_HEADS=1: keep the per-head 1x1 convs as separate launches
  bool no_dcn_tma = false;       // CP_NO_DCN_TMA=1: deformable convs on the global-gather kernel (A/B measurements)
  bool no_umma = false;          // CP_NO_UMMA=1: ops the TMA kernels do not take stay on the fp32 CUDA-core kernel (diagnostics)
  unsigned char* umma_wts = nullptr;
  size_t umma_bytes = 0;
  float* splitk_ws = nullptr;      // conv_tma / dcn_tma split-K partial sums (kSplitkWsFloats)
};

namespace cp {
namespace {

struct Builder {
  cp_plan* P;
  size_t act_cur = 0, w_cur = 0;
  int B;

  size_t walloc(size_t n) {
    size_t o = w_cur;
    w_cur += (n + 63) / 64 * 64;
    return o;
  }
  Act new_act(int C, int H, int W) {
    Act a;
    a.off = act_cur;
    a.C = C;
    a.H = H;
    a.W = W;
    a.stride = C;
    act_cur += ((size_t)B * H * W * C + 63) / 64 * 64;
    return a;
  }

  // generic conv (+ folded BN) op
  Act conv(const std::vector<Act>& srcs, const std::string& wkey, const std::string& bias_key,
           const std::string& bn, int Cout, int k, int stride, int pad, bool relu, const Act* res = nullptr,
           bool res_after = false, const Act* out_slice = nullptr, int colOff = 0, int ld = 0,
           size_t shared_w = (size_t)-1, size_t shared_b = (size_t)-1, int out_head = -1) {
    Op op;
    op.type = OP_IGEMM;
    op.nsrc = (int)srcs.size();
    int Cin = 0;
    for (int i = 0; i < op.nsrc; ++i) {
      op.src[i] = srcs[i];
      Cin += srcs[i].C;
    }
    op.mode = srcs[0].ext >= 0 ? IGEMM_NCHW_SCALAR : IGEMM_NHWC_VEC;
    op.kh = op.kw = k;
    op.stride = stride;
    op.pad = pad;
    op.Cin = Cin;
    op.Cout = Cout;
    op.CoutPad = round_up(Cout, 16);
    if (op.CoutPad > 16 && op.CoutPad % 32) op.CoutPad = round_up(Cout, 32);
    if (op.CoutPad > 32 && op.CoutPad % 64) op.CoutPad = round_up(Cout, 64);
    op.Kpad = round_up(k * k * Cin, 16);
    op.relu = relu;
    if (res) {
      op.has_res = true;
      op.res = *res;
      op.res_after_relu = res_after;
    }
    int Hin = srcs[0].H, Win = srcs[0].W;
    int Ho = (Hin + 2 * pad - k) / stride + 1, Wo = (Win + 2 * pad - k) / stride + 1;
    op.out_head = out_head;
    if (out_slice) {
      op.out = *out_slice;
    } else if (out_head < 0) {
      op.out = new_act(Cout, Ho, Wo);
      if (op.CoutPad != Cout) {
        // padded channels are never stored (the kernel masks n >= Cout); the
        // pixel stride is widened for the 27 -> 32 offset/mask tensor only
      }
    } else {
      op.out.C = Cout;
      op.out.H = Ho;
      op.out.W = Wo;
    }
    // weights
    op.w_ld = op.CoutPad;
    if (shared_w != (size_t)-1) {
      op.w_off = shared_w;
      op.b_off = shared_b;
    } else {
      int ldw = ld ? ld : op.CoutPad;
      op.w_off = walloc((size_t)op.Kpad * ldw);
      op.b_off = walloc(ldw);
      add_pack(wkey, bias_key, bn, Cout, Cin, k, op.CoutPad, op.Kpad, ldw, colOff, op.w_off, op.b_off);
    }
    op.name = wkey.empty() ? std::string("conv") + std::to_string(k) + "x" + std::to_string(k) + "_merged_" + std::to_string(Cout)
                           : wkey.substr(0, wkey.rfind('.'));
    P->ops.push_back(op);
    return op.out;
  }

  void add_pack(const std::string& wkey, const std::string& bias_key, const std::string& bn, int Cout, int Cin,
                int k, int CoutPad, int Kpad, int ld, int colOff, size_t w_off, size_t b_off) {
    PackJob jb;
    jb.type = PACK_BIAS;
    jb.key = bias_key;
    jb.bn = bn;
    jb.Cout = Cout;
    jb.CoutPad = CoutPad;
    jb.dst = b_off + colOff;
    jb.scale = walloc(CoutPad);
    P->jobs.push_back(jb);
    PackJob jw;
    jw.type = PACK_CONV;
    jw.key = wkey;
    jw.Cout = Cout;
    jw.Cin = Cin;
    jw.kh = jw.kw = k;
    jw.CoutPad = CoutPad;
    jw.Kpad = Kpad;
    jw.ld = ld;
    jw.colOff = colOff;
    jw.dst = w_off;
    jw.scale = jb.scale;
    jw.use_scale = !bn.empty();
    P->jobs.push_back(jw);
  }

  Act conv_bn(const Act& x, const std::string& convkey, const std::string& bnkey, int Cout, int k, int stride,
              int pad, bool relu, const Act* res = nullptr, bool res_after = false) {
    return conv({x}, convkey + ".weight", "", bnkey, Cout, k, stride, pad, relu, res, res_after);
  }

  Act maxpool(const Act& x) {
    Op op;
    op.type = OP_MAXPOOL;
    op.src[0] = x;
    op.out = new_act(x.C, x.H / 2, x.W / 2);
    op.name = "maxpool2";
    P->ops.push_back(op);
    return op.out;
  }

  Act basic_block(const Act& x, const std::string& p, int Cout, int stride, const Act& residual) {
    Act y = conv_bn(x, p + ".conv1", p + ".bn1", Cout, 3, stride, 1, true);
    return conv_bn(y, p + ".conv2", p + ".bn2", Cout, 3, 1, 1, true, &residual, false);
  }

  // levels == 1 Tree
  Act tree1(const Act& x, const std::string& p, int Cin, int Cout, int stride, bool level_root,
            const std::vector<Act>& extra) {
    std::vector<Act> children = extra;
    Act bottom = stride > 1 ? maxpool(x) : x;
    Act residual = bottom;
    if (Cin != Cout) residual = conv_bn(bottom, p + ".project.0", p + ".project.1", Cout, 1, 1, 0, false);
    if (level_root) children.push_back(bottom);
    Act x1 = basic_block(x, p + ".tree1", Cout, stride, residual);
    Act x2 = basic_block(x1, p + ".tree2", Cout, 1, x1);
    std::vector<Act> cat = {x2, x1};
    for (auto& c : children) cat.push_back(c);
    return conv(cat, p + ".root.conv.weight", "", p + ".root.bn", Cout, 1, 1, 0, true);
  }

  // levels == 2 Tree with level_root (level3 / level4); the outer `project` is dead compute
  Act tree2(const Act& x, const std::string& p, int Cin, int Cout, int stride) {
    Act bottom = maxpool(x);
    Act x1 = tree1(x, p + ".tree1", Cin, Cout, stride, false, {});
    return tree1(x1, p + ".tree2", Cout, Cout, 1, false, {bottom, x1});
  }

  Act deform_conv(const Act& x, const std::string& p, int Cout) {
    // offset / mask conv: 3x3 -> 27 channels stored with pixel stride 32
    Op om;
    om.type = OP_IGEMM;
    om.nsrc = 1;
    om.src[0] = x;
    om.mode = IGEMM_NHWC_VEC;
    om.kh = om.kw = 3;
    om.stride = 1;
    om.pad = 1;
    om.Cin = x.C;
    om.Cout = 27;
    om.CoutPad = 32;
    om.Kpad = 9 * x.C;
    om.out = new_act(32, x.H, x.W);
    om.out.C = 27;
    om.w_ld = 32;
    om.w_off = walloc((size_t)om.Kpad * 32);
    om.b_off = walloc(32);
    add_pack(p + ".conv.conv_offset_mask.weight", p + ".conv.conv_offset_mask.bias", "", 27, x.C, 3, 32,
             om.Kpad, 32, 0, om.w_off, om.b_off);
    om.name = p + ".conv.conv_offset_mask";
    P->ops.push_back(om);

    Op op;
    op.type = OP_IGEMM;
    op.nsrc = 1;
    op.src[0] = x;
    op.mode = IGEMM_DCN;
    op.kh = op.kw = 3;
    op.stride = 1;
    op.pad = 1;
    op.Cin = x.C;
    op.Cout = Cout;
    op.CoutPad = round_up(Cout, 64);
    op.Kpad = 9 * x.C;
    op.relu = true;
    op.om = om.out;
    op.out = new_act(Cout, x.H, x.W);
    op.w_ld = op.CoutPad;
    op.w_off = walloc((size_t)op.Kpad * op.CoutPad);
    op.b_off = walloc(op.CoutPad);
    add_pack(p + ".conv.weight", p + ".conv.bias", p + ".actf.0", Cout, x.C, 3, op.CoutPad, op.Kpad,
             op.CoutPad, 0, op.w_off, op.b_off);
    op.name = p + ".conv(dcn)";
    P->ops.push_back(op);
    return op.out;
  }

  Act up_add(const Act& x, const std::string& key, int f, const Act& skip) {
    Op op;
    op.type = OP_UPADD;
    op.src[0] = x;
    op.f = f;
    op.has_skip = true;
    op.skip = skip;
    op.out = new_act(x.C, x.H * f, x.W * f);
    op.upw_off = walloc((size_t)x.C * 4 * f * f);
    PackJob j;
    j.type = PACK_UP;
    j.key = key;
    j.Cout = x.C;
    j.kh = 2 * f;
    j.dst = op.upw_off;
    P->jobs.push_back(j);
    op.name = key.substr(0, key.rfind('.'));
    P->ops.push_back(op);
    return op.out;
  }

  // IDAUp.forward (pose_dla_dcn.py:411-417); up_f[j] = up-sampling factor of proj_j
  void ida_up(std::vector<Act>& layers, const std::string& p, int startp, int endp, int o,
              const std::vector<int>& up_f) {
    for (int i = startp + 1; i < endp; ++i) {
      int j = i - startp;
      Act t = deform_conv(layers[i], p + ".proj_" + std::to_string(j), o);
      Act u = up_add(t, p + ".up_" + std::to_string(j) + ".weight", up_f[j], layers[i - 1]);
      layers[i] = deform_conv(u, p + ".node_" + std::to_string(j), o);
    }
  }
};

int build_graph(cp_plan* P) {
  const cp_config& c = P->cfg;
  Builder b;
  b.P = P;
  b.B = P->B;
  P->ops.clear();
  P->jobs.clear();
  const int H = P->H, W = P->W;

  auto ext = [&](int idx, int C) {
    Act a;
    a.ext = idx;
    a.C = C;
    a.H = H;
    a.W = W;
    return a;
  };
  // ---- DLA-34 base (pose_dla_dcn.py:310-322)
  Act x = b.conv({ext(0, 3)}, "base.base_layer.0.weight", "", "base.base_layer.1", 16, 7, 1, 3, true);
  if (c.tracking) {
    x = b.conv({ext(1, 3)}, "base.pre_img_layer.0.weight", "", "base.pre_img_layer.1", 16, 7, 1, 3, true, &x, true);
    x = b.conv({ext(2, 1)}, "base.pre_hm_layer.0.weight", "", "base.pre_hm_layer.1", 16, 7, 1, 3, true, &x, true);
    x = b.conv({ext(3, 8)}, "base.pre_hm_hp_layer.0.weight", "", "base.pre_hm_hp_layer.1", 16, 7, 1, 3, true, &x,
               true);
  }
  std::vector<Act> lv(6);
  lv[0] = b.conv_bn(x, "base.level0.0", "base.level0.1", 16, 3, 1, 1, true);
  lv[1] = b.conv_bn(lv[0], "base.level1.0", "base.level1.1", 32, 3, 2, 1, true);
  lv[2] = b.tree1(lv[1], "base.level2", 32, 64, 2, false, {});
  lv[3] = b.tree2(lv[2], "base.level3", 64, 128, 2);
  lv[4] = b.tree2(lv[3], "base.level4", 128, 256, 2);
  lv[5] = b.tree1(lv[4], "base.level5", 256, 512, 2, true, {});

  // ---- DLAUp (pose_dla_dcn.py:420-443), first_level = 2
  const int first = 2;
  std::vector<int> channels = {64, 128, 256, 512};
  std::vector<int> in_channels = channels;
  std::vector<int> scales = {1, 2, 4, 8};
  struct Ida {
    int o;
    std::vector<int> up_f;
  };
  std::vector<Ida> idas;
  const int nch = (int)channels.size();
  for (int i = 0; i < nch - 1; ++i) {
    int j = nch - i - 2;
    Ida d;
    d.o = channels[j];
    for (int t = j; t < nch; ++t) d.up_f.push_back(scales[t] / scales[j]);
    idas.push_back(d);
    for (int t = j + 1; t < nch; ++t) {
      scales[t] = scales[j];
      in_channels[t] = channels[j];
    }
  }
  std::vector<Act> layers = lv;
  std::vector<Act> out = {layers.back()};
  for (int i = 0; i < (int)layers.size() - first - 1; ++i) {
    b.ida_up(layers, "dla_up.ida_" + std::to_string(i), (int)layers.size() - i - 2, (int)layers.size(), idas[i].o,
             idas[i].up_f);
    out.insert(out.begin(), layers.back());
  }
  // ---- ida_up over out[0:3] (pose_dla_dcn.py:487-488, 533-536), last_level = 5
  std::vector<Act> y(out.begin(), out.begin() + 3);
  b.ida_up(y, "ida_up", 0, 3, 64, {1, 2, 4});
  Act F = y.back();

  // ---- optional convGRU (convGRU.py:72-94): the three input convs see the same x at every step -> hoisted
  std::vector<Act> feat_for_head(c.num_heads, F);
  const bool gru = c.arch == CP_ARCH_DLAV1_34;
  if (gru) {
    const int steps = c.tracking_task_gru ? 4 : 3;
    const int HC = 64;
    // xi = [Wir x + b | Wiz x + b | Win x + b]
    Act xi = b.new_act(3 * HC, F.H, F.W);
    size_t wx = b.walloc((size_t)9 * 64 * 3 * HC), bx = b.walloc(3 * HC);
    const char* xin[3] = {"Wir", "Wiz", "Win"};
    const char* hin[3] = {"Whr", "Whz", "Whn"};
    for (int g = 0; g < 3; ++g)
      b.add_pack(std::string("convGRU.cell0.") + xin[g] + ".weight", std::string("convGRU.cell0.") + xin[g] + ".bias",
                 "", HC, 64, 3, HC, 9 * 64, 3 * HC, g * HC, wx, bx);
    b.conv({F}, "", "", "", 3 * HC, 3, 1, 1, false, nullptr, false, &xi, 0, 0, wx, bx);
    size_t wh = b.walloc((size_t)9 * HC * 3 * HC), bh = b.walloc(3 * HC);
    for (int g = 0; g < 3; ++g)
      b.add_pack(std::string("convGRU.cell0.") + hin[g] + ".weight", "", "", HC, HC, 3, HC, 9 * HC, 3 * HC, g * HC, wh,
                 bh);
    std::vector<Act> hs;
    Act hprev;
    for (int s = 0; s < steps; ++s) {
      Act hh;
      if (s > 0) {
        hh = b.new_act(3 * HC, F.H, F.W);
        b.conv({hprev}, "", "", "", 3 * HC, 3, 1, 1, false, nullptr, false, &hh, 0, 0, wh, bh);
      }
      Op op;
      op.type = OP_GRU;
      op.gx = xi;
      op.gh = hh;
      op.gprev = hprev;
      op.first_step = (s == 0);
      op.out = b.new_act(HC, F.H, F.W);
      op.name = "convGRU.gates.step" + std::to_string(s);
      P->ops.push_back(op);
      hprev = op.out;
      hs.push_back(op.out);
    }
    for (int h = 0; h < c.num_heads; ++h) {
      const std::string& n = P->head_names[h];
      int r = -1;
      if (c.tracking_task_gru) {
        if (n == "tracking" || n == "tracking_hp") r = 0;
        else if (n == "hm" || n == "wh" || n == "reg") r = 1;
        else if (n == "hm_hp" || n == "hp_offset" || n == "hps" || n == "hps_uncertainty") r = 2;
        else if (n == "scale" || n == "scale_uncertainty") r = 3;
      } else {
        if (n == "hm" || n == "wh" || n == "reg") r = 0;
        else if (n == "hm_hp" || n == "hp_offset" || n == "hps") r = 1;
        else if (n == "scale") r = 2;
      }
      if (r < 0) return fail(CP_ERR_INVALID, "head '" + n + "' has no convGRU route (pose_dla_dcn.py:545-563)");
      feat_for_head[h] = hs[r];
    }
  }

  // ---- heads: all heads that read the same feature share one 3x3 conv launch (N = n_heads * head_conv)
  const int HCV = c.head_conv;
  std::vector<bool> done(c.num_heads, false);
  for (int h0 = 0; h0 < c.num_heads; ++h0) {
    if (done[h0]) continue;
    std::vector<int> grp;
    for (int h = h0; h < c.num_heads; ++h)
      if (!done[h] && feat_for_head[h].off == feat_for_head[h0].off) grp.push_back(h);
    const Act& f = feat_for_head[h0];
    const int N = (int)grp.size() * HCV;
    Act mid = b.new_act(N, f.H, f.W);
    size_t wm = b.walloc((size_t)9 * f.C * N), bm = b.walloc(N);
    for (size_t gi = 0; gi < grp.size(); ++gi) {
      const std::string& n = P->head_names[grp[gi]];
      b.add_pack(n + ".0.weight", n + ".0.bias", "", HCV, f.C, 3, HCV, 9 * f.C, N, (int)gi * HCV, wm, bm);
    }
    b.conv({f}, "", "", "", N, 3, 1, 1, !gru, nullptr, false, &mid, 0, 0, wm, bm);
    const size_t merged_idx = P->ops.size() - 1;
    std::vector<int> children;
    for (size_t gi = 0; gi < grp.size(); ++gi) {
      int h = grp[gi];
      const std::string& n = P->head_names[h];
      done[h] = true;
      Act slice = mid;
      slice.off += gi * HCV;
      slice.C = HCV;
      slice.stride = N;
      std::string last = n + ".2";
      if (gru) {
        Op g;
        g.type = OP_GN_RELU;
        g.out = slice;
        g.groups = (HCV % 32 == 0) ? 32 : 16;
        g.gamma_off = b.walloc(HCV);
        g.beta_off = b.walloc(HCV);
        PackJob j1;
        j1.type = PACK_VEC;
        j1.key = n + ".1.weight";
        j1.Cout = HCV;
        j1.dst = g.gamma_off;
        P->jobs.push_back(j1);
        PackJob j2 = j1;
        j2.key = n + ".1.bias";
        j2.dst = g.beta_off;
        P->jobs.push_back(j2);
        g.name = n + ".1(groupnorm+relu)";
        P->ops.push_back(g);
        last = n + ".3";
      }
      b.conv({slice}, last + ".weight", last + ".bias", "", c.head_channels[h], 1, 1, 0, false, nullptr, false,
             nullptr, 0, 0, (size_t)-1, (size_t)-1, h);
      children.push_back((int)P->ops.size() - 1);
    }
    if (!gru) P->ops[merged_idx].head_children = children;     // GroupNorm sits between the two convs in dlav1
  }
  // plan-owned head buffers (NCHW) for cp_infer
  P->head_bufs.clear();
  for (int h = 0; h < c.num_heads; ++h) {
    Act a = b.new_act(c.head_channels[h], H / 4, W / 4);
    P->head_bufs.push_back(a);
  }
  P->act_floats = b.act_cur;
  P->w_floats = b.w_cur;
  // tensor-core eligibility + weight-tile storage
  P->umma_bytes = 0;
  if (P->prec >= 0) {
    for (auto& op : P->ops) {
      if (op.type != OP_IGEMM) continue;
      IgemmParams q{};
      q.mode = op.mode;
      q.nsrc = op.nsrc;
      q.Cin = op.Cin;
      q.CoutPad = op.CoutPad;
      for (int i = 0; i < op.nsrc; ++i) {
        q.srcC[i] = op.src[i].C;
        q.srcStride[i] = op.src[i].stride;
      }
      q.kh = op.kh;
      q.kw = op.kw;
      q.stride = op.stride;
      q.pad = op.pad;
      q.Win = op.src[0].W;
      const int gather_prec = P->prec == 2 ? 1 : P->prec;
      if (op.src[0].ext >= 0) continue;
      // 16-channel layers (level0 / level1): 133 K single-tile CTAs of almost no MMA work are dominated by the fixed
      // per-CTA cost of a tcgen05 kernel (measured 5.2 ms vs 2.5 ms on the FFMA kernel) -> keep them on CUDA cores
      if (op.Cin < P->min_tc_cin) continue;
      q.Hin = op.src[0].H;
      if ((P->prec == 2 || P->prec == 1) && !P->no_dcn_tma && dcn_tma_supported(q, P->prec == 1)) {
        op.use_dcn_tma = true;
        op.umma_off = P->umma_bytes;
        P->umma_bytes += (tma_weight_bytes(op.Cin, 9, op.CoutPad, P->prec == 1) + 1023) / 1024 * 1024;
      } else if ((P->prec == 2 || P->prec == 1) && tma_conv_supported(q, P->prec == 1)) {
        op.use_tma = true;
        op.umma_off = P->umma_bytes;
        P->umma_bytes += (tma_weight_bytes(op.Cin, op.kh * op.kw, op.CoutPad, P->prec == 1) + 1023) / 1024 * 1024;
      } else if (!P->no_umma && umma_supported(q, gather_prec)) {
        op.use_umma = true;
        op.umma_off = P->umma_bytes;
        P->umma_bytes += (umma_weight_bytes(op.kh * op.kw * op.Cin, op.CoutPad, gather_prec) + 1023) / 1024 * 1024;
      }
    }
  }
  // The per-head 1x1 convs move into the epilogue of the merged heads conv when that one runs on conv_tma: hidden =
  // relu(conv3x3) never reaches HBM (3.7 GB of writes + 3.7 GB of reads and seven launches at batch 32).  Single-pass
  // tf32: the epilogue thread that owns a position multiplies it with the head's [256][16] weights (3.8 -> 3.4 ms).
  // tf32x3: the same threads also promote the accumulation groups of the NEXT tile, so the 1x1 weights + 3x3 bias of
  // the tile are staged in shared memory by a dedicated warp and only the float4 groups holding real output channels are
  // multiplied (heads 7.7 -> 5.2 ms; the first version, 16 padded outputs through __ldg, stalled the promotion: 9.9 ms).
  if (P->prec == 2 || P->prec == 1) {
    for (auto& op : P->ops) {
      if (op.head_children.empty() || !op.use_tma || P->no_fuse_heads) continue;
      const int bn = tma_tile_n(op.CoutPad, P->prec == 1);
      bool ok = op.relu && !op.has_res && (c.head_conv % bn == 0) && (int)op.head_children.size() <= 16 &&
                op.CoutPad == (int)op.head_children.size() * c.head_conv && (P->prec != 1 || bn == 128);
      for (int ci : op.head_children) {
        const Op& ch = P->ops[ci];
        ok = ok && ch.CoutPad == 16 && ch.w_ld == 16 && ch.Cin == c.head_conv && ch.out_head >= 0 && !ch.relu && !ch.has_res;
      }
      if (!ok) continue;
      op.fuse_heads = true;
      for (int ci : op.head_children) P->ops[ci].fused_away = true;
    }
  }
  return CP_OK;
}


}  // namespace
}  // namespace cp

// ------------------------------------------------------------------------------------
extern "C" {

int cp_version(void) { return CP_ABI_VERSION; }
const char* cp_last_error(void) { return cp::g_last_error.c_str(); }

int cp_plan_create(const cp_config* cfg, cp_plan** out) {
  if (!cfg || !out) return fail(CP_ERR_INVALID, "cp_plan_create: null argument");
  if (cfg->arch != CP_ARCH_DLA34 && cfg->arch != CP_ARCH_DLAV1_34)
    return fail(CP_ERR_INVALID, "cp_plan_create: unknown arch");
  if (cfg->precision != CP_PREC_FP32 && cfg->precision != CP_PREC_TF32X3 && cfg->precision != CP_PREC_BF16 &&
      cfg->precision != CP_PREC_TF32)
    return fail(CP_ERR_INVALID, "cp_plan_create: unknown precision");
  if (cfg->height % 32 || cfg->width % 32 || cfg->height <= 0 || cfg->width <= 0)
    return fail(CP_ERR_INVALID, "cp_plan_create: height/width must be positive multiples of 32");
  if (cfg->max_batch <= 0 || cfg->num_heads <= 0 || cfg->num_heads > CP_MAX_HEADS)
    return fail(CP_ERR_INVALID, "cp_plan_create: bad batch / head count");
  if (cfg->head_conv <= 0 || cfg->head_conv % 64)
    return fail(CP_ERR_INVALID, "cp_plan_create: head_conv must be a positive multiple of 64");
  if (cfg->arch == CP_ARCH_DLAV1_34 && cfg->head_conv != 256)
    return fail(CP_ERR_INVALID, "cp_plan_create: dlav1 GroupNorm path needs head_conv == 256");
  std::unique_ptr<cp_plan> P(new cp_plan());
  P->cfg = *cfg;
  for (int i = 0; i < cfg->num_heads; ++i) {
    if (!cfg->head_names[i] || cfg->head_channels[i] <= 0 || cfg->head_channels[i] > 16)
      return fail(CP_ERR_INVALID, "cp_plan_create: head channels must be in 1..16");
    P->head_names.push_back(cfg->head_names[i]);
  }
  for (int i = 0; i < cfg->num_heads; ++i) P->cfg.head_names[i] = P->head_names[i].c_str();
  P->B = cfg->max_batch;
  P->H = cfg->height;
  P->W = cfg->width;
  P->prec = cfg->precision == CP_PREC_BF16 ? 0 : (cfg->precision == CP_PREC_TF32X3 ? 1 : (cfg->precision == CP_PREC_TF32 ? 2 : -1));
  if (const char* e = getenv("CP_MIN_TC_CIN")) P->min_tc_cin = atoi(e);
  if (const char* e = getenv("CP_NO_FUSE_HEADS")) P->no_fuse_heads = atoi(e) != 0;
  if (const char* e = getenv("CP_NO_DCN_TMA")) P->no_dcn_tma = atoi(e) != 0;
  if (const char* e = getenv("CP_NO_UMMA")) P->no_umma = atoi(e) != 0;
  // the plan lives on cfg->device; the caller's current device is restored on every exit path
  struct DeviceGuard {
    int prev = -1;
    ~DeviceGuard() {
      if (prev >= 0) cudaSetDevice(prev);
    }
  } guard;
  CP_CUDA_CHECK(cudaGetDevice(&guard.prev));
  CP_CUDA_CHECK(cudaSetDevice(cfg->device));
  int rc = build_graph(P.get());
  if (rc) return rc;
  CP_CUDA_CHECK(cudaMalloc(&P->act, P->act_floats * sizeof(float)));
  CP_CUDA_CHECK(cudaMalloc(&P->wts, P->w_floats * sizeof(float)));
  CP_CUDA_CHECK(cudaMemset(P->wts, 0, P->w_floats * sizeof(float)));
  CP_CUDA_CHECK(cudaMalloc(&P->gn_stats, sizeof(double) * (size_t)P->B * 64 * 2));
  if (P->umma_bytes) CP_CUDA_CHECK(cudaMalloc(&P->umma_wts, P->umma_bytes));
  if (P->prec == 1 || P->prec == 2) {
    CP_CUDA_CHECK(cudaMalloc(&P->splitk_ws, kSplitkWsFloats * sizeof(float)));
  }
  for (auto& op : P->ops) {
    if (!op.use_dcn_tma) continue;
    IgemmParams q{};
    q.nsrc = 1;
    q.src[0] = P->act + op.src[0].off;
    q.srcC[0] = op.src[0].C;
    q.srcStride[0] = op.src[0].stride;
    q.Hin = op.src[0].H;
    q.Win = op.src[0].W;
    op.tma_maps.resize(512 + 64);
    unsigned char* mp = (unsigned char*)(((uintptr_t)op.tma_maps.data() + 63) & ~(uintptr_t)63);
    int rc2 = dcn_tma_encode(q, P->B, mp);
    if (rc2) return rc2;
  }
  for (auto& op : P->ops) {
    if (!op.use_tma) continue;
    IgemmParams q{};
    q.nsrc = op.nsrc;
    for (int i = 0; i < op.nsrc; ++i) {
      q.src[i] = P->act + op.src[i].off;
      q.srcC[i] = op.src[i].C;
      q.srcStride[i] = op.src[i].stride;
    }
    q.kh = op.kh;
    q.kw = op.kw;
    q.Cin = op.Cin;          // Cin / kh / Win / CoutPad select 32- vs 16-channel slabs (tma_cslab)
    q.CoutPad = op.CoutPad;
    q.Hin = op.src[0].H;
    q.Win = op.src[0].W;
    op.tma_cslab = tma_cslab(q, P->prec == 1);
    op.tma_maps.resize(512 + 64);
    unsigned char* mp = (unsigned char*)(((uintptr_t)op.tma_maps.data() + 63) & ~(uintptr_t)63);
    int rc2 = tma_conv_encode(q, P->B, P->prec == 1, mp);
    if (rc2) return rc2;
  }
  int n = 0;
  for (auto& op : P->ops) n += op.fused_away ? 0 : ((op.type == OP_GN_RELU) ? 2 : 1);
  P->launches = n;
  *out = P.release();
  return CP_OK;
}

int cp_plan_destroy(cp_plan* P) {
  if (!P) return CP_OK;
  cudaFree(P->act);
  cudaFree(P->wts);
  cudaFree(P->gn_stats);
  if (P->umma_wts) cudaFree(P->umma_wts);
  if (P->splitk_ws) cudaFree(P->splitk_ws);
  if (P->decode_ws) cudaFree(P->decode_ws);
  delete P;
  return CP_OK;
}

int64_t cp_plan_bytes(const cp_plan* P) { return P ? (int64_t)((P->act_floats + P->w_floats) * sizeof(float)) : 0; }
int32_t cp_plan_forward_launches(const cp_plan* P) { return P ? P->launches : 0; }

int cp_plan_load_weights(cp_plan* P, const char* const* names, const void* const* ptrs, const int64_t* numel,
                         int32_t n, void* stream_) {
  if (!P || !names || !ptrs || !numel) return fail(CP_ERR_INVALID, "cp_plan_load_weights: null argument");
  cudaStream_t s = (cudaStream_t)stream_;
  std::map<std::string, WRef> m;
  for (int i = 0; i < n; ++i) {
    std::string k = names[i];
    if (k.rfind("module.", 0) == 0 && k.rfind("module_list", 0) != 0) k = k.substr(7);
    m[k] = WRef{(const float*)ptrs[i], numel[i]};
  }
  auto get = [&](const std::string& k, int64_t want, const float** out) -> int {
    auto it = m.find(k);
    if (it == m.end()) return fail(CP_ERR_MISSING_KEY, "state_dict key missing: " + k);
    if (it->second.n != want)
      return fail(CP_ERR_SHAPE, "state_dict key " + k + " has " + std::to_string(it->second.n) + " elements, plan needs " +
                                    std::to_string(want));
    *out = it->second.p;
    return CP_OK;
  };
  int rc;
  for (auto& j : P->jobs) {
    switch (j.type) {
      case PACK_BIAS: {
        const float *cb = nullptr, *g = nullptr, *be = nullptr, *mu = nullptr, *var = nullptr;
        if (!j.key.empty() && (rc = get(j.key, j.Cout, &cb))) return rc;
        if (!j.bn.empty()) {
          if ((rc = get(j.bn + ".weight", j.Cout, &g))) return rc;
          if ((rc = get(j.bn + ".bias", j.Cout, &be))) return rc;
          if ((rc = get(j.bn + ".running_mean", j.Cout, &mu))) return rc;
          if ((rc = get(j.bn + ".running_var", j.Cout, &var))) return rc;
        }
        if ((rc = launch_pack_bias(cb, g, be, mu, var, P->wts + j.scale, P->wts + j.dst, j.Cout, j.CoutPad, 1e-5f, s)))
          return rc;
        break;
      }
      case PACK_CONV: {
        const float* w = nullptr;
        if ((rc = get(j.key, (int64_t)j.Cout * j.Cin * j.kh * j.kw, &w))) return rc;
        if ((rc = launch_pack_conv_weight(w, j.use_scale ? P->wts + j.scale : nullptr, P->wts + j.dst, j.Cout, j.Cin,
                                          j.kh, j.kw, j.CoutPad, j.Kpad, j.ld, j.colOff, s)))
          return rc;
        break;
      }
      case PACK_UP: {
        const float* w = nullptr;
        if ((rc = get(j.key, (int64_t)j.Cout * j.kh * j.kh, &w))) return rc;
        if ((rc = launch_pack_up_weight(w, P->wts + j.dst, j.Cout, j.kh, s))) return rc;
        break;
      }
      case PACK_VEC: {
        const float* w = nullptr;
        if ((rc = get(j.key, j.Cout, &w))) return rc;
        CP_CUDA_CHECK(cudaMemcpyAsync(P->wts + j.dst, w, sizeof(float) * j.Cout, cudaMemcpyDeviceToDevice, s));
        break;
      }
    }
  }
  // second pass: tensor-core weight tiles are cut from the finished fp32 matrices (merged matrices are complete now)
  for (auto& op : P->ops) {
    if (op.type != OP_IGEMM) continue;
    if (op.use_dcn_tma) {
      if ((rc = launch_pack_tma_weight(P->wts + op.w_off, op.w_ld, op.Cin, 9, op.Cout, op.CoutPad, 1, P->prec == 1, 16,
                                       P->umma_wts + op.umma_off, s, dcn_tma_tile_n(op.CoutPad, P->prec == 1))))
        return rc;
    } else if (op.use_tma) {
      if ((rc = launch_pack_tma_weight(P->wts + op.w_off, op.w_ld, op.Cin, op.kh * op.kw, op.Cout, op.CoutPad, 1,
                                       P->prec == 1, op.tma_cslab, P->umma_wts + op.umma_off, s)))
        return rc;
    } else if (op.use_umma) {
      if ((rc = launch_pack_umma_weight(P->wts + op.w_off, op.w_ld, op.kh * op.kw * op.Cin, op.Cout, op.CoutPad,
                                        P->prec == 2 ? 1 : P->prec, P->umma_wts + op.umma_off, s)))
        return rc;
    }
  }
  P->loaded = true;
  return CP_OK;
}

struct ProfCtx {
  std::vector<cudaEvent_t> ev;
};

// PDL hides ~2 us of launch latency + prologue per kernel: a constant ~0.19 ms of the forward (scripts/pdl_sweep.py, ms with /
// without: batch 1 1.99 / 2.18, 2 2.77 / 2.97, 4 4.45 / 4.64, 8 7.73 / 7.83, 16 14.70 / 14.65, 32 27.42 / 27.17).  Beyond
// batch 8 the kernels run for 100s of us and the early CTAs of the next launch only add scheduling work, so it is switched
// on by the amount of work.  CP_PDL=1 / CP_NO_PDL=1 force it.
static bool pdl_wanted(long long pixels) {
  if (getenv("CP_NO_PDL")) return false;
  if (const char* e = getenv("CP_PDL")) return atoi(e) != 0;
  return pixels <= 8ll * 512 * 512;
}

static int run_forward(cp_plan* P, int batch, const float* const ext[4], float* const* head_out, cudaStream_t s,
                       ProfCtx* prof = nullptr) {
  int rc;
  const long long launches0 = g_launch_counter;
  // programmatic dependent launch for the whole schedule (common.cuh); off while profiling (events sit between the ops)
  struct PdlScope {
    explicit PdlScope(int on) { g_pdl = on; }
    ~PdlScope() { g_pdl = 0; }
  } pdl_scope(!prof && pdl_wanted((long long)batch * P->H * P->W));
  for (auto& op : P->ops) {
    if (prof) {
      cudaEvent_t e;
      CP_CUDA_CHECK(cudaEventCreate(&e));
      CP_CUDA_CHECK(cudaEventRecord(e, s));
      prof->ev.push_back(e);
    }
    if (op.fused_away) continue;          // computed inside the epilogue of the merged heads conv
    switch (op.type) {
      case OP_IGEMM: {
        IgemmParams p{};
        p.nsrc = op.nsrc;
        for (int i = 0; i < op.nsrc; ++i) {
          const Act& a = op.src[i];
          if (a.ext >= 0) {
            if (!ext[a.ext]) return fail(CP_ERR_INVALID, "cp_forward: a tracking input tensor is null");
            p.src[i] = ext[a.ext];
          } else {
            p.src[i] = P->act + a.off;
          }
          p.srcC[i] = a.C;
          p.srcStride[i] = a.stride;
        }
        p.B = batch;
        p.Hin = op.src[0].H;
        p.Win = op.src[0].W;
        p.Cin = op.Cin;
        p.kh = op.kh;
        p.kw = op.kw;
        p.stride = op.stride;
        p.pad = op.pad;
        p.Hout = (p.Hin + 2 * op.pad - op.kh) / op.stride + 1;
        p.Wout = (p.Win + 2 * op.pad - op.kw) / op.stride + 1;
        p.Cout = op.Cout;
        p.CoutPad = op.CoutPad;
        p.Kpad = op.Kpad;
        p.wgt = P->wts + op.w_off;
        p.bias = P->wts + op.b_off;
        p.residual = op.has_res ? P->act + op.res.off : nullptr;
        p.resStride = op.res.stride;
        p.relu = op.relu;
        p.res_after_relu = op.res_after_relu;
        if (op.out_head >= 0) {
          p.out = head_out[op.out_head];
          p.out_nchw = 1;
          p.outStride = 0;
        } else {
          p.out = P->act + op.out.off;
          p.outStride = op.out.stride;
        }
        if (op.mode == IGEMM_DCN) {
          p.offmask = P->act + op.om.off;
          p.omStride = op.om.stride;
          p.mask_is_logit = 1;
        }
        p.mode = op.mode;
        if (op.fuse_heads) {
          p.fuse_n = (int)op.head_children.size();
          p.fuse_hidden = P->cfg.head_conv;
          for (int i = 0; i < p.fuse_n; ++i) {
            const Op& ch = P->ops[op.head_children[i]];
            p.fuse_w[i] = P->wts + ch.w_off;
            p.fuse_b[i] = P->wts + ch.b_off;
            p.fuse_out[i] = head_out[ch.out_head];
            p.fuse_cout[i] = ch.Cout;
          }
        }
        if (op.use_dcn_tma) {
          p.wgt_umma = P->umma_wts + op.umma_off;
          p.splitk_ws = P->splitk_ws;
          p.splitk_ws_floats = P->splitk_ws ? kSplitkWsFloats : 0;
          const unsigned char* mp = (const unsigned char*)(((uintptr_t)op.tma_maps.data() + 63) & ~(uintptr_t)63);
          if ((rc = launch_dcn_tma(p, mp, P->prec == 1, P->prec == 2, s))) return rc;
        } else if (op.use_tma) {
          p.wgt_umma = P->umma_wts + op.umma_off;
          p.splitk_ws = P->splitk_ws;
          p.splitk_ws_floats = P->splitk_ws ? kSplitkWsFloats : 0;
          const unsigned char* mp = (const unsigned char*)(((uintptr_t)op.tma_maps.data() + 63) & ~(uintptr_t)63);
          if ((rc = launch_conv_tma(p, mp, P->prec == 2, P->prec == 1, s))) return rc;
        } else if (op.use_umma) {
          p.wgt_umma = P->umma_wts + op.umma_off;
          if ((rc = launch_igemm_umma(p, P->prec == 2 ? 1 : P->prec, s))) return rc;
        } else if (stem_supported(p)) {
          if ((rc = launch_stem_conv(p, s))) return rc;
        } else if (conv3_c16_supported(p)) {
          if ((rc = launch_conv3_c16(p, s))) return rc;
        } else if ((rc = launch_igemm_fp32(p, s))) {
          return rc;
        }
        break;
      }
      case OP_MAXPOOL:
        if ((rc = launch_maxpool2(P->act + op.src[0].off, P->act + op.out.off, batch, op.src[0].H, op.src[0].W,
                                  op.src[0].C, s)))
          return rc;
        break;
      case OP_UPADD:
        if ((rc = launch_upsample_add(P->act + op.src[0].off, P->wts + op.upw_off,
                                      op.has_skip ? P->act + op.skip.off : nullptr, P->act + op.out.off, batch,
                                      op.src[0].H, op.src[0].W, op.src[0].C, op.f, s)))
          return rc;
        break;
      case OP_GN_RELU:
        if ((rc = launch_group_norm_relu(P->act + op.out.off, P->wts + op.gamma_off, P->wts + op.beta_off, batch,
                                         op.out.H * op.out.W, op.out.C, op.out.stride, 0, op.groups, 1e-5f,
                                         (float*)P->gn_stats, s)))
          return rc;
        break;
      case OP_GRU:
        if ((rc = launch_gru_gates(P->act + op.gx.off, op.first_step ? nullptr : P->act + op.gh.off,
                                   op.first_step ? nullptr : P->act + op.gprev.off, P->act + op.out.off,
                                   batch * op.out.H * op.out.W, op.out.C, op.first_step, s)))
          return rc;
        break;
    }
  }
  if (prof) {
    cudaEvent_t e;
    CP_CUDA_CHECK(cudaEventCreate(&e));
    CP_CUDA_CHECK(cudaEventRecord(e, s));
    prof->ev.push_back(e);
  }
  P->launches = (int)(g_launch_counter - launches0);      // measured, replaces the estimate of cp_plan_create
  return CP_OK;
}

// Algorithmic work of one op at batch `batch` (2*MAC; fp32 bytes of the tensors it must touch once).
static void op_work(const Op& op, int batch, double* flops, double* bytes) {
  *flops = 0;
  *bytes = 0;
  const double B = batch;
  switch (op.type) {
    case OP_IGEMM: {
      int Hin = op.src[0].H, Win = op.src[0].W;
      int Ho = (Hin + 2 * op.pad - op.kh) / op.stride + 1, Wo = (Win + 2 * op.pad - op.kw) / op.stride + 1;
      double M = B * Ho * Wo;
      *flops = 2.0 * M * op.kh * op.kw * op.Cin * op.Cout;
      *bytes = 4.0 * (B * Hin * Win * op.Cin + (double)op.kh * op.kw * op.Cin * op.Cout + (op.fuse_heads ? 0.0 : M * op.Cout) +
                      (op.has_res ? M * op.Cout : 0.0) + (op.mode == IGEMM_DCN ? M * 27 : 0.0));
      if (op.fused_away) {      // accounted to the parent below
        *flops = 0;
        *bytes = 0;
      }
      break;
    }
    case OP_MAXPOOL:
      *bytes = 4.0 * B * op.src[0].H * op.src[0].W * op.src[0].C * 1.25;
      break;
    case OP_UPADD: {
      double o = B * op.out.H * op.out.W * op.out.C;
      *flops = 2.0 * o * 4;
      *bytes = 4.0 * (o * 2 + B * op.src[0].H * op.src[0].W * op.src[0].C);
      break;
    }
    case OP_GN_RELU:
      *bytes = 4.0 * B * op.out.H * op.out.W * op.out.C * 3;
      break;
    case OP_GRU:
      *bytes = 4.0 * B * op.out.H * op.out.W * op.out.C * 8;
      break;
  }
}

int cp_forward(cp_plan* P, int32_t batch, const float* images, const float* pre_img, const float* pre_hm,
               const float* pre_hm_hp, float* const* head_out, void* stream) {
  if (!P || !images || !head_out) return fail(CP_ERR_INVALID, "cp_forward: null argument");
  if (!P->loaded) return fail(CP_ERR_NOT_LOADED, "cp_forward: call cp_plan_load_weights first");
  if (batch <= 0 || batch > P->B) return fail(CP_ERR_INVALID, "cp_forward: batch exceeds the plan's max_batch");
  const float* ext[4] = {images, pre_img, pre_hm, pre_hm_hp};
  return run_forward(P, batch, ext, head_out, (cudaStream_t)stream);
}

int cp_plan_num_ops(const cp_plan* P) { return P ? (int)P->ops.size() : 0; }

int cp_plan_profile(cp_plan* P, int32_t batch, const float* images, const float* pre_img, const float* pre_hm,
                    const float* pre_hm_hp, float* const* head_out, void* stream, cp_op_stat* stats, int32_t max_stats,
                    int32_t* n_stats) {
  if (!P || !images || !head_out || !stats || !n_stats) return fail(CP_ERR_INVALID, "cp_plan_profile: null argument");
  if (!P->loaded) return fail(CP_ERR_NOT_LOADED, "cp_plan_profile: call cp_plan_load_weights first");
  if (batch <= 0 || batch > P->B) return fail(CP_ERR_INVALID, "cp_plan_profile: batch exceeds the plan's max_batch");
  const float* ext[4] = {images, pre_img, pre_hm, pre_hm_hp};
  ProfCtx ctx;
  int rc = run_forward(P, batch, ext, head_out, (cudaStream_t)stream, &ctx);
  if (rc == CP_OK) {
    CP_CUDA_CHECK(cudaEventSynchronize(ctx.ev.back()));
    int n = 0;
    for (size_t i = 0; i + 1 < ctx.ev.size() && n < max_stats; ++i, ++n) {
      const Op& op = P->ops[i];
      cp_op_stat& st = stats[n];
      memset(&st, 0, sizeof(st));
      snprintf(st.name, sizeof(st.name), "%s", op.name.c_str());
      st.kind = (int)op.type * 10 + (op.type == OP_IGEMM ? op.mode : 0);
      cudaEventElapsedTime(&st.ms, ctx.ev[i], ctx.ev[i + 1]);
      op_work(op, batch, &st.flops, &st.bytes);
    }
    *n_stats = n;
  }
  for (auto e : ctx.ev) cudaEventDestroy(e);
  return rc;
}

int cp_infer(cp_plan* P, int32_t batch, const float* images, const float* pre_img, const float* pre_hm,
             const float* pre_hm_hp, const cp_decode_params* prm, const double* meta, float* const* heads_out,
             float* dets, float* poses, int32_t* n_valid, void* stream) {
  if (!P || !images || !prm || !meta || !poses || !n_valid) return fail(CP_ERR_INVALID, "cp_infer: null argument");
  if (!P->loaded) return fail(CP_ERR_NOT_LOADED, "cp_infer: call cp_plan_load_weights first");
  if (batch <= 0 || batch > P->B) return fail(CP_ERR_INVALID, "cp_infer: batch exceeds the plan's max_batch");
  float* hp[CP_MAX_HEADS];
  for (int h = 0; h < P->cfg.num_heads; ++h)
    hp[h] = (heads_out && heads_out[h]) ? heads_out[h] : P->act + P->head_bufs[h].off;
  const float* ext[4] = {images, pre_img, pre_hm, pre_hm_hp};
  int rc = run_forward(P, batch, ext, hp, (cudaStream_t)stream);
  if (rc) return rc;
  cp_heads hd{};
  for (int h = 0; h < P->cfg.num_heads; ++h) {
    const std::string& n = P->head_names[h];
    if (n == "hm" && P->cfg.head_channels[h] != prm->num_classes)
      return fail(CP_ERR_INVALID, "cp_infer: prm->num_classes differs from the plan's hm channels");
    if (n == "hm") hd.hm = hp[h];
    else if (n == "wh") hd.wh = hp[h];
    else if (n == "hps") hd.hps = hp[h];
    else if (n == "reg") hd.reg = hp[h];
    else if (n == "hm_hp") hd.hm_hp = hp[h];
    else if (n == "hp_offset") hd.hp_offset = hp[h];
    else if (n == "scale") hd.scale = hp[h];
    else if (n == "hps_uncertainty") hd.hps_uncertainty = hp[h];
    else if (n == "scale_uncertainty") hd.scale_uncertainty = hp[h];
    else if (n == "tracking") hd.tracking = hp[h];
    else if (n == "tracking_hp") hd.tracking_hp = hp[h];
  }
  cp_decode_params q = *prm;
  q.batch = batch;
  q.out_h = P->H / 4;
  q.out_w = P->W / 4;
  q.apply_sigmoid = prm->apply_sigmoid == 2 ? 2 : 1;     // the plan's heads are logits; 2 = opt.mse_loss (raw hm_hp)
  size_t need = cp_decode_workspace_bytes(&q);
  if (need > P->decode_ws_bytes) {
    // grows only on the first call for a given K / batch (not steady state)
    if (P->decode_ws) cudaFree(P->decode_ws);
    cp_decode_params qmax = q;
    qmax.batch = P->B;
    size_t cap = cp_decode_workspace_bytes(&qmax);
    CP_CUDA_CHECK(cudaMalloc(&P->decode_ws, cap));
    P->decode_ws_bytes = cap;
  }
  return cp_decode_pnp(&q, &hd, meta, dets, poses, n_valid, P->decode_ws, P->decode_ws_bytes, stream);
}

}  // extern "C"
