// Static description of the three N-CMAPSS regressors as a tape of conv / pool ops.
// Every layer is expressed as a stride-1 conv2d over a [C,H,W] view (a Linear over a flattened
// buffer is a 1x1 conv over the [C*H*W,1,1] view of the same memory).
// Reference: models/nets/inception.py:142-217, conv.py:14-79, linear.py:10-72.
#pragma once
#include <cstdint>
#include <vector>

namespace brl {

struct LayerSpec {
  int cout, cin, kh, kw, ph, pw;
  long long w_off, b_off;
  int wndim;
  long long wshape[4];
  float drop_factor;  // fraction of p_dropout applied at the site after this layer (0: no site)
  int out_elems;      // per-window elements of the layer output (Cout * Hout * Wout)
};

struct ViewSpec {
  int buf;  // -1: the input windows x; else activation buffer id
  int C, H, W;
};

enum OpKind { OP_CONV = 0, OP_MAXPOOL3 = 1, OP_AVGPOOL2 = 2 };

struct OpSpec {
  int kind;
  int layer;  // OP_CONV
  ViewSpec in;
  int out_buf, co_off;
  int Hout, Wout;
  int relu, head;
};

struct BufSpec {
  int C, H, W;
  int shared;  // 1: identical for every MC sample (derived from x only)
  long long elems() const { return (long long)C * H * W; }
};

struct NetSpec {
  int id;
  long long P;
  int xC, xH, xW, xsC, xsH, xsW;  // how the [30,18] window is viewed by the first layer(s)
  std::vector<LayerSpec> layers;
  std::vector<BufSpec> bufs;
  std::vector<OpSpec> ops;
  int out_buf;
  long long flops_fwd;
  std::vector<long long> site_off;  // n_sites + 1 entries
};

const NetSpec& get_net(int id);  // throws std::invalid_argument for unknown ids

}  // namespace brl
