// Geometry shared by every 3x3 conv kernel (CUDA-core and tcgen05 paths).
//
// All four uses of a 3x3 / pad-1 convolution on the hot path are one
// "gather-GEMM":   Out[out_pix(j), n] = sum_{t in taps(phase)} sum_c
//                                       In[in_pix(j, t), c] * Wq[n][wt(t)][c]
// over an iteration grid j in [B] x [Hj] x [Wj]:
//   fprop  stride s : in_pix = s*j + (kh-1, kw-1), out_pix = j,        1 phase, 9 taps
//   dgrad  stride 1 : in_pix = j + (1-kh, 1-kw),   out_pix = j,        1 phase, 9 taps
//   dgrad  stride 2 : in_pix = j + d(kh,kw),       out_pix = 2*j + (ph,pw),
//                     4 phases (output parity) with 1/2/2/4 taps  -- this is also
//                     the ConvTranspose2d(k3,s2,p1,op1) forward (models.py:66-68,77).
#pragma once
#include <stdint.h>

namespace vae {

struct TapTable {
  int ntaps;
  int ph, pw;
  int8_t dh[9], dw[9], wt[9];
};

struct GatherGeom {
  int B, Hin, Win, Cin;   // gathered tensor (NHWC)
  int Hout, Wout, Nout;   // produced tensor (NHWC)
  int Hj, Wj;             // iteration grid per phase
  int in_mul, out_mul;    // pixel multipliers
  int nphase;
  TapTable phase[4];
};

// forward conv, stride s: X[B,Hi,Wi,Ci] -> Y[B,Hi/s,Wi/s,Co]
static inline GatherGeom geom_fprop(int B, int Hi, int Wi, int Ci, int Co, int s) {
  GatherGeom g{};
  g.B = B; g.Hin = Hi; g.Win = Wi; g.Cin = Ci;
  g.Hout = Hi / s; g.Wout = Wi / s; g.Nout = Co;
  g.Hj = g.Hout; g.Wj = g.Wout;
  g.in_mul = s; g.out_mul = 1; g.nphase = 1;
  TapTable& t = g.phase[0];
  t.ntaps = 9; t.ph = 0; t.pw = 0;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      int i = kh * 3 + kw;
      t.dh[i] = (int8_t)(kh - 1); t.dw[i] = (int8_t)(kw - 1); t.wt[i] = (int8_t)i;
    }
  return g;
}

// data gradient of that conv: dY[B,Hi/s,Wi/s,Co] -> dX[B,Hi,Wi,Ci]
static inline GatherGeom geom_dgrad(int B, int Hi, int Wi, int Ci, int Co, int s) {
  GatherGeom g{};
  g.B = B; g.Hin = Hi / s; g.Win = Wi / s; g.Cin = Co;
  g.Hout = Hi; g.Wout = Wi; g.Nout = Ci;
  g.Hj = Hi / s; g.Wj = Wi / s;
  g.in_mul = 1; g.out_mul = s; g.nphase = s * s;
  for (int ph = 0; ph < s; ++ph)
    for (int pw = 0; pw < s; ++pw) {
      TapTable& t = g.phase[ph * s + pw];
      t.ntaps = 0; t.ph = ph; t.pw = pw;
      for (int kh = 0; kh < 3; ++kh) {
        if ((ph + 1 - kh) % s != 0) continue;
        for (int kw = 0; kw < 3; ++kw) {
          if ((pw + 1 - kw) % s != 0) continue;
          int i = t.ntaps++;
          // (ph+1-kh) may be negative (-1 for s=1, kh=2): exact division by s
          t.dh[i] = (int8_t)((ph + 1 - kh) / s);
          t.dw[i] = (int8_t)((pw + 1 - kw) / s);
          t.wt[i] = (int8_t)(kh * 3 + kw);
        }
      }
    }
  return g;
}

}  // namespace vae
