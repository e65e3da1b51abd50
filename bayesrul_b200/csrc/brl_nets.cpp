#include "brl_nets.h"

#include <stdexcept>

namespace brl {

namespace {

struct Builder {
  NetSpec n;
  long long off = 0;

  int layer(int cout, int cin, int kh, int kw, int ph, int pw, int wndim, float drop_factor, int out_elems) {
    LayerSpec L{};
    L.cout = cout; L.cin = cin; L.kh = kh; L.kw = kw; L.ph = ph; L.pw = pw;
    L.w_off = off;
    off += (long long)cout * cin * kh * kw;
    L.b_off = off;
    off += cout;
    L.wndim = wndim;
    L.drop_factor = drop_factor;
    L.out_elems = out_elems;
    n.layers.push_back(L);
    return (int)n.layers.size() - 1;
  }
  int buf(int C, int H, int W, int shared = 0) {
    n.bufs.push_back(BufSpec{C, H, W, shared});
    return (int)n.bufs.size() - 1;
  }
  void conv(int layer, ViewSpec in, int out_buf, int co_off, int Hout, int Wout, int relu = 1, int head = 0) {
    n.ops.push_back(OpSpec{OP_CONV, layer, in, out_buf, co_off, Hout, Wout, relu, head});
  }
  void pool(int kind, ViewSpec in, int out_buf, int Hout) {
    n.ops.push_back(OpSpec{kind, -1, in, out_buf, 0, Hout, in.W, 0, 0});
  }
  NetSpec finish(int id, long long flops) {
    n.id = id;
    n.P = off;
    n.flops_fwd = flops;
    n.out_buf = (int)n.bufs.size() - 1;
    for (auto& L : n.layers) {
      n.site_off.push_back(L.w_off);
      n.site_off.push_back(L.b_off);
    }
    n.site_off.push_back(off);
    return n;
  }
};

NetSpec build_inception() {
  Builder b;
  b.n.xC = 18; b.n.xH = 30; b.n.xW = 1; b.n.xsC = 1; b.n.xsH = 18; b.n.xsW = 0;  // x.transpose(2,1), inception.py:213
  // conv1d weights [Cout, Cin, k]: kh = k (time), kw = 1
  auto c1d = [&](int co, int ci, int k, float df) {
    int l = b.layer(co, ci, k, 1, (k - 1) / 2, 0, 3, df, co * 30);
    auto& L = b.n.layers[l];
    L.wshape[0] = co; L.wshape[1] = ci; L.wshape[2] = k;
    return l;
  };
  auto lin = [&](int co, int ci, float df) {
    int l = b.layer(co, ci, 1, 1, 0, 0, 2, df, co);
    auto& L = b.n.layers[l];
    L.wshape[0] = co; L.wshape[1] = ci;
    return l;
  };
  const int l0 = c1d(27, 18, 1, .25f), l1 = c1d(27, 18, 3, .25f), l2 = c1d(27, 18, 5, .25f), l3 = c1d(27, 18, 3, .25f);
  const int l4 = c1d(16, 108, 1, .25f), l5 = c1d(64, 108, 1, 0.f), l6 = c1d(16, 64, 3, .25f);
  const int l7 = c1d(64, 108, 1, 0.f), l8 = c1d(16, 64, 5, .25f), l9 = c1d(32, 108, 1, .25f);
  const int l10 = lin(64, 2400, 1.0f), l11 = lin(2, 64, 0.f);
  const int XP = b.buf(18, 30, 1, 1), M1 = b.buf(108, 30, 1), T2 = b.buf(64, 30, 1), T3 = b.buf(64, 30, 1);
  const int M1P = b.buf(108, 30, 1), M2 = b.buf(80, 30, 1), H = b.buf(64, 1, 1), O = b.buf(2, 1, 1);
  const ViewSpec X{-1, 18, 30, 1};
  b.conv(l0, X, M1, 0, 30, 1);
  b.conv(l1, X, M1, 27, 30, 1);
  b.conv(l2, X, M1, 54, 30, 1);
  b.pool(OP_MAXPOOL3, X, XP, 30);
  b.conv(l3, ViewSpec{XP, 18, 30, 1}, M1, 81, 30, 1);
  b.conv(l4, ViewSpec{M1, 108, 30, 1}, M2, 0, 30, 1);
  b.conv(l5, ViewSpec{M1, 108, 30, 1}, T2, 0, 30, 1);
  b.conv(l6, ViewSpec{T2, 64, 30, 1}, M2, 16, 30, 1);
  b.conv(l7, ViewSpec{M1, 108, 30, 1}, T3, 0, 30, 1);
  b.conv(l8, ViewSpec{T3, 64, 30, 1}, M2, 32, 30, 1);
  b.pool(OP_MAXPOOL3, ViewSpec{M1, 108, 30, 1}, M1P, 30);
  b.conv(l9, ViewSpec{M1P, 108, 30, 1}, M2, 48, 30, 1);
  b.conv(l10, ViewSpec{M2, 2400, 1, 1}, H, 0, 1, 1);  // Flatten: index c*30+t is the memory order of M2
  b.conv(l11, ViewSpec{H, 64, 1, 1}, O, 0, 1, 1, 0, 1);
  return b.finish(0, 2289376);
}

NetSpec build_conv() {
  Builder b;
  b.n.xC = 1; b.n.xH = 30; b.n.xW = 18; b.n.xsC = 0; b.n.xsH = 18; b.n.xsW = 1;  // x.unsqueeze(1), conv.py:75
  auto c2d = [&](int co, int ci, int kh, int kw, int out_elems) {
    int l = b.layer(co, ci, kh, kw, 0, 0, 4, 1.0f, out_elems);
    auto& L = b.n.layers[l];
    L.wshape[0] = co; L.wshape[1] = ci; L.wshape[2] = kh; L.wshape[3] = kw;
    return l;
  };
  const int l0 = c2d(16, 1, 5, 9, 16 * 26 * 10), l1 = c2d(32, 16, 2, 10, 32 * 25), l2 = c2d(64, 32, 2, 1, 64 * 11);
  const int l3 = b.layer(2, 320, 1, 1, 0, 0, 2, 0.f, 2);
  b.n.layers[l3].wshape[0] = 2; b.n.layers[l3].wshape[1] = 320;
  const int A1 = b.buf(16, 26, 10), A2 = b.buf(32, 25, 1), A2P = b.buf(32, 12, 1), A3 = b.buf(64, 11, 1);
  const int A3P = b.buf(64, 5, 1), O = b.buf(2, 1, 1);
  b.conv(l0, ViewSpec{-1, 1, 30, 18}, A1, 0, 26, 10);
  b.conv(l1, ViewSpec{A1, 16, 26, 10}, A2, 0, 25, 1);
  b.pool(OP_AVGPOOL2, ViewSpec{A2, 32, 25, 1}, A2P, 12);
  b.conv(l2, ViewSpec{A2P, 32, 12, 1}, A3, 0, 11, 1);
  b.pool(OP_AVGPOOL2, ViewSpec{A3, 64, 11, 1}, A3P, 5);
  b.conv(l3, ViewSpec{A3P, 320, 1, 1}, O, 0, 1, 1, 0, 1);
  return b.finish(1, 977792);
}

NetSpec build_linear() {
  Builder b;
  b.n.xC = 540; b.n.xH = 1; b.n.xW = 1; b.n.xsC = 1; b.n.xsH = 0; b.n.xsW = 0;  // Flatten of [1,30,18]
  const int dims[6] = {540, 256, 128, 128, 32, 2};
  int prev = -1;
  for (int i = 0; i < 5; ++i) {
    const int l = b.layer(dims[i + 1], dims[i], 1, 1, 0, 0, 2, i < 4 ? 1.0f : 0.f, dims[i + 1]);
    b.n.layers[l].wshape[0] = dims[i + 1];
    b.n.layers[l].wshape[1] = dims[i];
    const int ob = b.buf(dims[i + 1], 1, 1);
    b.conv(l, ViewSpec{prev, dims[i], 1, 1}, ob, 0, 1, 1, i < 4, i == 4);
    prev = ob;
  }
  return b.finish(2, 383104);
}

}  // namespace

const NetSpec& get_net(int id) {
  static const NetSpec nets[3] = {build_inception(), build_conv(), build_linear()};
  if (id < 0 || id > 2) throw std::invalid_argument("unknown net id");
  return nets[id];
}

}  // namespace brl
