// Depthwise kxk + bias + swish + squeeze-excite for the small-spatial MBConv blocks (output <= 30x40), fp32.
// (reference: models/efficientnet.py:81-90: depthwise conv, BN, swish, then the squeeze-excite gate.)
//
// k_dwconv_roll walks an image column by column with one thread per (channel vector, 4 columns); at 15x20
// and 7x10 that leaves ~14 warps per SM, each one a serial chain of row loads, and the launch runs 5-10x
// below both the HBM and the FMA bound.  Here a CTA owns (image, 32 channels, tile of output rows x columns):
//   1. the zero-padded input strip [rows][cols][32 channels] is staged in shared memory with 16-byte loads
//      (a pixel's 32 channels are one 128-byte line of the NHWC activation);
//   2. lane = channel, warp = output row: every input value is read once from shared memory (stride-32
//      floats across pixels, consecutive across lanes: conflict free) and feeds the <= KS outputs of the row it
//      touches; the KS*KS taps of the lane's channel stay in registers;
//   3. bias + swish, 128-byte coalesced row stores, per-channel sums reduced over the CTA's warps in a fixed order.
// Squeeze-excite: the reduce FC is linear in the channel sums, so every CTA adds its 32 channels' share
//   partial_r[b][strip][chunk][j] = sum_{c in chunk} Wr[j][c] * sum_strip[c]
// and the small k_se_fc2 launch finishes the gate for all images:
//   gate[c] = sigmoid(be[c] + sum_j We[c][j] * swish(br[j] + inv_hw * sum_{strip,chunk} partial_r[..][j]))
// No atomics; the order of every floating-point sum is fixed, so results are run-to-run identical.
#pragma once
#include "common.h"
#include "kernels_backbone.cuh"

namespace cosyb {

constexpr int DWT_THREADS = 256;
constexpr int DWT_WARPS = DWT_THREADS / 32;
constexpr int DWT_CC = 32;   // channels per CTA

// A CTA's output tile is R rows x (WO * XU) columns; a warp works on units of (row, WO adjacent columns).
// grid = (ceil(C / 32), n_strips * n_xtiles, B); dynamic smem = ((R-1)*S + KS) * ((WO*XU-1)*S + KS) * 32 floats
template <int KS, int S, int WO, int XU>
__global__ void __launch_bounds__(DWT_THREADS)
k_dw_tile(const float* __restrict__ in, const float* __restrict__ w /*[KS*KS][C]*/, const float* __restrict__ bias,
          float* __restrict__ out, float* __restrict__ partial_r /*[B][n_strips * n_xtiles][n_chunks][Cse]*/, int H,
          int W, int C, int Ho, int Wo, int pad, int R, int n_xt, int Cse, const float* __restrict__ wr /*[Cse][C]*/) {
  constexpr int WP = (WO * XU - 1) * S + KS;   // staged columns
  constexpr int UP = (WO - 1) * S + KS;        // columns read by one unit
  extern __shared__ __align__(16) float s_tile[];
  __shared__ float s_red[DWT_WARPS][DWT_CC];
  __shared__ __align__(16) float s_csum[DWT_CC];
  const int strip = blockIdx.y / n_xt, xt = blockIdx.y % n_xt, b = blockIdx.z, n_strips = gridDim.y;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const int c0 = blockIdx.x * DWT_CC;
  const int oy0 = strip * R, ox0 = xt * WO * XU;
  const int rows = min(R, Ho - oy0);
  const int HP = (rows - 1) * S + KS;
  const int iy0 = oy0 * S - pad;

  // ---- stage the padded input strip
  const float* inb = in + (size_t)b * H * W * C + c0;
  const int n_v4 = HP * WP * (DWT_CC / 4);
#pragma unroll 4
  for (int i = tid; i < n_v4; i += DWT_THREADS) {
    const int pix = i >> 3, q = i & 7;
    const int py = pix / WP, px = pix - py * WP;
    const int iy = iy0 + py, ix = ox0 * S + px - pad;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W && c0 + 4 * q < C)
      v = __ldg(reinterpret_cast<const float4*>(inb + ((size_t)iy * W + ix) * C + 4 * q));
    *reinterpret_cast<float4*>(s_tile + pix * DWT_CC + 4 * q) = v;
  }
  const int c = c0 + lane;
  const bool cok = c < C;
  float wreg[KS * KS];
#pragma unroll
  for (int t = 0; t < KS * KS; ++t) wreg[t] = cok ? __ldg(w + (size_t)t * C + c) : 0.f;
  const float bv = cok ? __ldg(bias + c) : 0.f;
  __syncthreads();

  // ---- depthwise: warp = (output row, column group), lane = channel
  float psum = 0.f;
  for (int u = wp; u < rows * XU; u += DWT_WARPS) {
    const int r = u / XU, xu = u % XU;
    float acc[WO];
#pragma unroll
    for (int x = 0; x < WO; ++x) acc[x] = 0.f;
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
      const float* srow = s_tile + (size_t)((r * S + ky) * WP + xu * WO * S) * DWT_CC + lane;
#pragma unroll
      for (int px = 0; px < UP; ++px) {
        const float v = srow[px * DWT_CC];
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          if (px - kx >= 0 && (px - kx) % S == 0 && (px - kx) / S < WO)
            acc[(px - kx) / S] = fmaf(v, wreg[ky * KS + kx], acc[(px - kx) / S]);
        }
      }
    }
    float* orow = out + ((size_t)(b * Ho + oy0 + r) * Wo + ox0 + xu * WO) * C + c;
#pragma unroll
    for (int x = 0; x < WO; ++x) {
      const float o = swishf(acc[x] + bv);
      psum += o;
      if (cok) orow[(size_t)x * C] = o;
    }
  }
  s_red[wp][lane] = psum;
  __syncthreads();
  if (tid < DWT_CC) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < DWT_WARPS; ++k) s += s_red[k][tid];
    s_csum[tid] = s;   // lanes beyond C hold exact zeros (zero taps, zero bias -> swish(0) = 0)
  }
  __syncthreads();
  // ---- this CTA's share of the squeeze-excite reduce FC
  if (tid < Cse) {
    const float* wrow = wr + (size_t)tid * C + c0;
    float a = 0.f;
#pragma unroll
    for (int q = 0; q < DWT_CC / 4; ++q) {
      if (c0 + 4 * q < C) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow + 4 * q));
        const float4 m = *reinterpret_cast<const float4*>(s_csum + 4 * q);
        a = fmaf(wv.x, m.x, fmaf(wv.y, m.y, fmaf(wv.z, m.z, fmaf(wv.w, m.w, a))));
      }
    }
    partial_r[(((size_t)b * n_strips + blockIdx.y) * gridDim.x + blockIdx.x) * Cse + tid] = a;
  }
}

// Finishes the squeeze-excite gate from the reduce-FC partial sums of k_dw_tile.
// grid = (ceil(C / 256), B), 256 threads; we_t is the expand weight transposed to [Cse][C].
constexpr int SE2_THREADS = 256;
__global__ void __launch_bounds__(SE2_THREADS)
k_se_fc2(const float* __restrict__ partial_r, int n_part /*strips * chunks*/, int C, int Cse, float inv_hw,
         const float* __restrict__ br, const float* __restrict__ we_t, const float* __restrict__ be,
         float* __restrict__ gate) {
  __shared__ float s_half[2][128];
  __shared__ float s_r[128];
  const int b = blockIdx.y, tid = threadIdx.x;
  {
    // thread (half, j): sums every second-half / first-half partial of output j, 8 loads in flight
    const int j = tid & 127, half = tid >> 7;
    const int p0 = half == 0 ? 0 : n_part / 2, p1 = half == 0 ? n_part / 2 : n_part;
    float s = 0.f;
    if (j < Cse) {
      const float* src = partial_r + (size_t)b * n_part * Cse + j;
      int p = p0;
      for (; p + 8 <= p1; p += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (size_t)(p + u) * Cse);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
      }
      for (; p < p1; ++p) s += __ldg(src + (size_t)p * Cse);
    }
    s_half[half][j] = s;
  }
  __syncthreads();
  if (tid < Cse) s_r[tid] = swishf((s_half[0][tid] + s_half[1][tid]) * inv_hw + __ldg(br + tid));
  __syncthreads();
  const int c = blockIdx.x * SE2_THREADS + tid;
  if (c < C) {
    float a0 = __ldg(be + c), a1 = 0.f;
    int j = 0;
    for (; j + 8 <= Cse; j += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(we_t + (size_t)(j + u) * C + c);
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        a0 = fmaf(v[u], s_r[j + u], a0);
        a1 = fmaf(v[u + 1], s_r[j + u + 1], a1);
      }
    }
    for (; j < Cse; ++j) a0 = fmaf(__ldg(we_t + (size_t)j * C + c), s_r[j], a0);
    gate[(size_t)b * C + c] = sigmoidf_(a0 + a1);
  }
}

// ---- host side -------------------------------------------------------------------------------------------
struct DwTilePlan { bool ok; int R, n_strips, n_xt, wo, xu, n_chunks, smem_bytes; };

inline DwTilePlan dw_tile_plan(const BlockSpec& b) {
  DwTilePlan p{};
  p.ok = false;
  if (!((b.k == 3 || b.k == 5) && (b.s == 1 || b.s == 2))) return p;
  // instances (launch_dw_tile): unit width x units per tile row
  if (b.wout == 10) { p.wo = 10; p.xu = 1; }
  else if (b.wout == 20) { p.wo = 20; p.xu = 1; }
  else if (b.wout == 40 && b.k == 5 && b.s == 1) { p.wo = 40; p.xu = 1; }
  else return p;   // the 120x160 and 60x80 blocks keep the rolling kernel (measured: the tiled kernel is slower there)
  const int tile_w = p.wo * p.xu;
  p.n_xt = b.wout / tile_w;
  auto smem = [&](int R) { return ((R - 1) * b.s + b.k) * ((tile_w - 1) * b.s + b.k) * DWT_CC * 4; };
  const int limit = 80 * 1024;
  p.R = b.hout;
  while (smem(p.R) > limit && p.R > 1) p.R = (p.R + 1) / 2;
  p.n_strips = (b.hout + p.R - 1) / p.R;
  p.R = (b.hout + p.n_strips - 1) / p.n_strips;
  p.n_chunks = (b.cexp + DWT_CC - 1) / DWT_CC;
  p.smem_bytes = smem(p.R);
  p.ok = true;
  return p;
}

}  // namespace cosyb
