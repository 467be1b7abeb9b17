// EfficientNet-B3 trunk table, evaluated from the published compound-scaling rule
// (width 1.2, depth 1.4, divisor 8, TF-"same" padding frozen for a 300x300 image).
// Reference behaviour: cosypose/models/efficientnet_utils.py:59-81 (filter / repeat rounding),
// :123-146 (static same padding), :169 (B3 coefficients), :259-264 (base stages);
// cosypose/models/efficientnet.py:136-157 (stage unrolling).  Mirrors cosypose_b200/effnet_spec.py;
// tests/test_spec.py checks the two agree through cosyb200_effnet_block().
#pragma once
#include <cmath>
#include <vector>

#include "common.h"

namespace cosyb {

inline int effnet_round_filters(int f) {
  const double width = 1.2;
  const int div = 8;
  double ff = f * width;
  int nf = std::max(div, int(ff + div / 2.0) / div * div);
  if (nf < 0.9 * ff) nf += div;
  return nf;
}

inline void effnet_same_pad(int k, int s, int* lo, int* hi) {
  const int img = 300;
  int out = (img + s - 1) / s;
  int pad = std::max((out - 1) * s + (k - 1) + 1 - img, 0);
  *lo = pad / 2;
  *hi = pad - pad / 2;
}

inline int conv_out(int n, int k, int s, int lo, int hi) { return (n + lo + hi - k) / s + 1; }

constexpr int STEM_OUT = 40;

inline std::vector<BlockSpec> make_effnet_b3() {
  struct Stage { int r, k, s, e, i, o; };
  const Stage stages[7] = {{1, 3, 1, 1, 32, 16}, {2, 3, 2, 6, 16, 24}, {2, 5, 2, 6, 24, 40},
                           {3, 3, 2, 6, 40, 80}, {3, 5, 1, 6, 80, 112}, {4, 5, 2, 6, 112, 192},
                           {1, 3, 1, 6, 192, 320}};
  std::vector<BlockSpec> out;
  int lo, hi;
  effnet_same_pad(3, 2, &lo, &hi);
  int h = conv_out(RENDER_H, 3, 2, lo, hi), w = conv_out(RENDER_W, 3, 2, lo, hi);
  for (const Stage& st : stages) {
    int cin = effnet_round_filters(st.i), cout = effnet_round_filters(st.o);
    int reps = int(std::ceil(1.4 * st.r));
    for (int r = 0; r < reps; ++r) {
      BlockSpec b;
      b.k = st.k;
      b.s = r == 0 ? st.s : 1;
      b.e = st.e;
      b.cin = r == 0 ? cin : cout;
      b.cexp = b.cin * st.e;
      b.cse = std::max(1, int(b.cin * 0.25));
      b.cout = cout;
      effnet_same_pad(b.k, b.s, &b.pad_lo, &b.pad_hi);
      b.skip = (b.s == 1 && b.cin == b.cout) ? 1 : 0;
      b.hin = h;
      b.win = w;
      b.hout = conv_out(h, b.k, b.s, b.pad_lo, b.pad_hi);
      b.wout = conv_out(w, b.k, b.s, b.pad_lo, b.pad_hi);
      h = b.hout;
      w = b.wout;
      out.push_back(b);
    }
  }
  return out;
}

}  // namespace cosyb
