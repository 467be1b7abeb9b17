// NOT BUILT.  Generalisation of kernels_xdw.cuh to the small-spatial blocks 9-24 (in-image rows only + zero ring in the E
// tile, single-buffered A operand for Kp up to 240 TMEM columns, channel-group items for the 7x10 blocks).  Correct
// (all 28 taps within 1e-4), but measured slower overall than the shipped split (profiles/r02_block_times.txt): the
// late blocks are bound by the depthwise phase itself (k5, 25 taps), not by the expanded tensor crossing HBM/L2, the
// per-chunk geometry costs the early blocks 15 %, and 7-15 hi*hi MMAs accumulated in TMEM cost accuracy (z-up case
// 1.5e-4 vs 7e-5).  Kept as a record of the experiment.
// Fused MBConv front half for sm_100a: 1x1 expand (tcgen05, 3xFP16 split) + BN + swish + depthwise kxk + BN +
// swish + squeeze-excite pooling in ONE kernel; the 6x-wide expanded activation never leaves the SM.
//
//   E[p][c]   = swish(x[p][:] . We[c][:] + be[c])            p over a spatial tile WITH its depthwise halo
//   D[o][c]   = swish(sum_taps E[o*s + tap][c] * Wd[tap][c] + bd[c])
//   pool[c]  += sum_o D[o][c]                                 (per-tile partial sums, fixed order)
//
// Reference op chain: models/efficientnet.py:71-90 (expand conv + BN + swish, depthwise conv + BN + swish,
// adaptive_avg_pool2d of the squeeze-excite branch); static "same" padding models/efficientnet_utils.py:123-146:
// the zero padding applies to the EXPANDED activation, so halo pixels outside the image are exact zeros in E.
// That mask costs nothing here: the expand bias rides in the GEMM (A gets a column of ones at k = Cin, the weights
// a row of biases), an out-of-image pixel is an all-zero A row, its accumulator is exactly 0 and swish(0) = 0.
//
// One CTA per SM, persistent over items = (hypothesis, spatial tile); an item walks over all chunks of `cc`
// (48 or 64) expanded channels:
//   workers (16 warps)  two groups of 8 warps that take ALTERNATE chunks (group g owns accumulator buffer g and its own
//                       E tile), so the MUFU-bound drain of one chunk runs against the FMA / shared-memory bound
//                       depthwise phase of the other instead of all 16 warps sitting on the same pipe;
//                       per item: x rows of the halo tile -> fp16 hi/lo split -> TMEM (A operand, resident for all
//                       chunks; thread = tile row = TMEM lane; double buffered, each group converts half of the k
//                       units of the next item's rows under its last chunk of the current item);
//                       per chunk: drain the accumulator (tcgen05.ld), swish -> shared-memory tile E[pixel][channel];
//                       then the depthwise convolution from shared memory with a rolling register window (lane =
//                       channel PAIR: 64-bit conflict-free loads, packed FFMA2), bias + swish, 256-byte row stores
//                       of D and the pooling partial sums.
//   warp 16             MMA issuer: per chunk and m-tile 3 * Kp/16 kind::f16 MMAs (a_lo*b_hi, a_hi*b_lo, then a_hi*b_hi),
//                       A from TMEM, B from shared memory; accumulators double buffered in TMEM so the MMAs of chunk
//                       c+1 run under the CUDA-core work of chunk c.
//   warp 17             weight loader: one bulk copy (cp.async.bulk + mbarrier) per chunk into a 2-slot ring.
// The kernel is bound by CUDA-core issue and the MUFU pipe (two swishes per expanded element), not by HBM or the
// tensor pipe: swish is evaluated on pairs with ONE reciprocal (1 / (d0 * d1), then * d1 and * d0), 1.5 MUFU ops
// per element instead of 2.
// Precision: as kernels_pw2.cuh (fp16 hi/lo split of both operands, power-of-two weight scale, fp32 accumulate).
#pragma once
#include "kernels_pw2.cuh"

namespace cosyb {
namespace xdw {

using namespace tc;
using pw2::make_desc;
using pw2::make_idesc_f16;
using pw2::pack_f16x2;
using pw2::split11;
using pw2::tmem_ld_wait;
using pw2::umma_commit_elect;
using pw2::umma_f16_ts_pred;

constexpr int NWW = 16;                 // worker warps
constexpr int WORKERS = NWW * 32;
constexpr int GROUP_WARPS = NWW / 2, GROUP_THREADS = GROUP_WARPS * 32;   // two worker groups on alternate chunks
constexpr int MMA_WARP = NWW, LOADER_WARP = NWW + 1;
constexpr int THREADS = (NWW + 4) * 32;      // warps 18-19 only complete the fifth warpgroup (setmaxnreg is per warpgroup)
constexpr int CC_MAX = 64;              // expanded channels per chunk (= MMA N): 48 or 64
constexpr int E_SLACK_ROWS = 16;        // the last x-segment of a tile may read (never use) a few pixels past the tile
constexpr int MAX_UNITS = 32;
constexpr int MAX_XU = 2;               // 16-k units of the A row a worker thread converts
constexpr uint32_t TMEM_COLS = 512;

struct Plan {
  int ok, MT, TH, TW, IH, IW, tiles_y, tiles_x, cc, n_chunks, n_cg, cpg, a_double, Kp, NX, NYS, RH, e_rows, smem_bytes;
};

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(u64 r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float ex2f(float t) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  return e;
}
__device__ __forceinline__ float rcpf(float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r;
}
// (swish(v0), swish(v1)) with one reciprocal: 1/d0 = d1 / (d0 d1).  The exponent is clamped so that d0 * d1 stays
// finite (for v < -55 the result is |v| 2^-80 instead of |v| e^v: both far below one ulp of anything they are added to).
__device__ __forceinline__ u64 swish2(u64 v) {
  const u64 t = fmul2(v, pk2(-1.4426950408889634f, -1.4426950408889634f));
  float t0, t1;
  upk2(t, t0, t1);
  const float d0 = 1.0f + ex2f(fminf(t0, 80.f)), d1 = 1.0f + ex2f(fminf(t1, 80.f));
  const float r = rcpf(d0 * d1);
  return fmul2(v, pk2(r * d1, r * d0));
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// 8 consecutive accumulator columns as 4 packed pairs
__device__ __forceinline__ void tmem_ld8_pairs(uint32_t taddr, u64* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 4; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "r"(r[2 * i]), "r"(r[2 * i + 1]));
}

// Wx: [n_cg * cpg chunks][hi|lo][Kp/8][cc][8] fp16, k = Cin holds the bias row, k > Cin and channels >= Cexp zero
// x [B][H][W][Cin], out [B][Ho][Wo][Cexp], partial [B][tiles][Cexp]
// item = (hypothesis, spatial tile, channel group of `cpg` chunks).  The A rows of an item are the IN-IMAGE pixels of
// its halo window (row-major over the clipped window); the out-of-image ring of the E tile is zero filled once per
// item, so padding costs neither MMA rows nor swishes.  a_double: two A buffers in TMEM (the next item is converted
// under the current one); otherwise one buffer, refilled when the MMA warp reports the item's last MMA complete.
template <int KS, int S, int NX, int CCT, int RH>
__global__ void __launch_bounds__(THREADS, 1)
k_xdw(const float* __restrict__ x, const __half* __restrict__ Wx, float inv_wscale, const float* __restrict__ dw_w,
      const float* __restrict__ dw_bias, float* __restrict__ out, float* __restrict__ partial, int B, int H, int W,
      int Cin, int Cexp, int Ho, int Wo, int pad, int MT, int TH, int TW, int IH, int IW, int tiles_y, int tiles_x,
      int n_cg, int cpg, int Kp, int NYS, int e_rows, int a_double, int do_trace) {
  constexpr int cc = CCT;                                           // expanded channels per chunk (MMA N)
  constexpr int NIN = (NX - 1) * S + KS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[12];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) float s_ps[2][MAX_UNITS * 64];
  const uint32_t b_base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t b_bytes = (uint32_t)Kp * cc * 4u;                  // hi + lo image of one chunk
  const uint32_t e_base0 = b_base + 2 * b_bytes;
  constexpr int EP = cc + 4;                                        // floats per E row (16-byte row stores conflict free)
  const uint32_t e_bytes = (uint32_t)e_rows * EP * 4u;              // one E tile (one per worker group)
  const int tid = threadIdx.x, lane = tid % 32;
  const int warp = __shfl_sync(0xffffffffu, tid / 32, 0);
  const int tiles = tiles_y * tiles_x;
  const int n_items = B * tiles * n_cg;
  const uint32_t ACC_STRIDE = (uint32_t)MT * cc;                    // columns of one accumulator buffer
  const uint32_t A_COL0 = 2 * ACC_STRIDE;
  const uint32_t A_STRIDE = (uint32_t)MT * Kp;                      // columns of one A buffer
  auto fullA = [&]() { return smem_u32(&bars[0]); };
  auto fullB = [&](int s) { return smem_u32(&bars[1 + s]); };
  auto emptyB = [&](int s) { return smem_u32(&bars[3 + s]); };
  auto acc_full = [&](int b) { return smem_u32(&bars[5 + b]); };
  auto acc_empty = [&](int b) { return smem_u32(&bars[7 + b]); };
  auto a_free = [&]() { return smem_u32(&bars[9]); };

  if (tid == 0) {
    mbar_init(fullA(), WORKERS);
    mbar_init(a_free(), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(fullB(s), 1);
      mbar_init(emptyB(s), 1);
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), GROUP_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(&s_tmem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, s_tmem, 0);
  const int my_items = (int)blockIdx.x < n_items ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp < NWW) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");   // 16 x 32 x (112 - 96) = what the fifth warpgroup releases (96 -> 32)
    // ------------------------------------------------------------------ workers
    const int group = warp / GROUP_WARPS, gw = warp % GROUP_WARPS;
    const int q = warp % 4;
    const int ncs = 2 / MT;                                    // MT == 1: the two warp sets of a group split the columns
    const int mt = (gw / 4) % MT, csplit = (gw / 4) / MT;
    const int CW = cc / ncs;
    const int sub = group;                                     // k units of the A rows this group converts: u = sub, sub + 2, ..
    const int row = mt * 128 + q * 32 + lane;                  // A row == TMEM lane of m-tile mt
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t e_base = e_base0 + (uint32_t)group * e_bytes;
    float* s_psg = s_ps[group];
    const int gtid = tid - group * GROUP_THREADS;
    const int n_xu = Kp / 16;

    // geometry of item `it`: clipped halo window, this thread's pixel, its slot in the E tile
    struct Geo { int img, tile, cg, oy0, ox0, dy0, dx0, chh, cw, n_rows; bool row_ok; int e_slot; const float* xr; };
    auto geometry = [&](int it) {
      Geo g;
      g.cg = it % n_cg;
      g.tile = (it / n_cg) % tiles;
      g.img = it / (n_cg * tiles);
      g.oy0 = (g.tile / tiles_x) * TH;
      g.ox0 = (g.tile % tiles_x) * TW;
      const int iy0 = g.oy0 * S - pad, ix0 = g.ox0 * S - pad;
      const int wy0 = max(iy0, 0), wx0 = max(ix0, 0);
      g.chh = min(iy0 + IH, H) - wy0;
      g.cw = min(ix0 + IW, W) - wx0;
      g.dy0 = wy0 - iy0;
      g.dx0 = wx0 - ix0;
      g.n_rows = g.chh * g.cw;
      g.row_ok = row < g.n_rows;
      const int cy = row / g.cw, cx = row - cy * g.cw;
      g.e_slot = (g.dy0 + cy) * IW + g.dx0 + cx;
      g.xr = x + (((size_t)g.img * H + (g.row_ok ? wy0 + cy : 0)) * W + (g.row_ok ? wx0 + cx : 0)) * Cin;
      return g;
    };
    auto load_unit = [&](const Geo& g, int u, float* v) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.row_ok && u < n_xu && u * 16 + c * 4 < Cin) t = __ldg(reinterpret_cast<const float4*>(g.xr + u * 16 + c * 4));
        v[c * 4] = t.x; v[c * 4 + 1] = t.y; v[c * 4 + 2] = t.z; v[c * 4 + 3] = t.w;
      }
    };
    // 16 k of this thread's row -> hi/lo fp16 -> TMEM A buffer `ab` (+ the ones column at k = Cin)
    auto store_unit = [&](bool row_ok, int u, float* v, int ab) {
      const float one = row_ok ? 1.f : 0.f;
      if (Cin == u * 16) v[0] = one;
      else if (Cin == u * 16 + 8) v[8] = one;
      uint32_t ph[8], pl[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float h0, l0, h1, l1;
        split11(v[2 * c], h0, l0);
        split11(v[2 * c + 1], h1, l1);
        ph[c] = pack_f16x2(h0, h1);
        pl[c] = pack_f16x2(l0, l1);
      }
      const uint32_t a_hi = t_lane + A_COL0 + ab * A_STRIDE + (uint32_t)(mt * Kp) + (uint32_t)(u * 8);
      tmem_st8(a_hi, ph);
      tmem_st8(a_hi + (uint32_t)(Kp / 2), pl);
    };
    auto publish_A = [&]() {
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(fullA());
    };
    // whole conversion of one item without prefetch (single-A mode, and the first item)
    auto convert_direct = [&](const Geo& g, int ab) {
      if (csplit == 0) {
        for (int u = sub; u < n_xu; u += 2) {
          float v[16];
          load_unit(g, u, v);
          store_unit(g.row_ok, u, v, ab);
        }
      }
      publish_A();
    };

    const int NXS = (TW + NX - 1) / NX;
    const int n_units = NXS * NYS;
    const int total_chunks = my_items * cpg;
    float xv[MAX_XU][16];                                      // a_double: prefetched rows of the next item
    bool nx_ok = false;
    if (my_items > 0) convert_direct(geometry(blockIdx.x), 0);
    for (int gch = group; gch < total_chunks; gch += 2) {      // this group's chunks; accumulator buffer == group
      const int local_it = gch / cpg, ch = gch - local_it * cpg;
      const int it = blockIdx.x + local_it * gridDim.x;
      const Geo g = geometry(it);
      const int chg = g.cg * cpg + ch;                         // chunk index in the weight image
      const int next = it + gridDim.x;
      const int buf = group;
      const bool first_in_item = gch - 2 < local_it * cpg;     // this group's first / last chunk inside the item
      const bool last_in_item = gch + 2 >= (local_it + 1) * cpg;
      const bool prep_next = a_double && last_in_item && next < n_items;
      const bool tr = do_trace && tid == 0 && gch < 32;
      if (tr) trace(512 + 8 * (gch >> 1));
      if (first_in_item) {
        if (!a_double && local_it > 0) {                       // the single A buffer is free once the previous item's MMAs are done
          mbar_wait_warp(a_free(), (local_it - 1) & 1);
          tc_fence_after();
          convert_direct(g, 0);
        }
        if (g.n_rows != IH * IW) {                             // zero the out-of-image ring of this group's E tile
          for (int e = gtid; e < IH * IW; e += GROUP_THREADS) {
            const int ey = e / IW - g.dy0, ex = e % IW - g.dx0;
            if (ey < 0 || ey >= g.chh || ex < 0 || ex >= g.cw) {
              const uint32_t dst = e_base + (uint32_t)(e * EP) * 4u;
#pragma unroll
              for (int c = 0; c < cc; c += 4)
                asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst + c * 4), "r"(0u) : "memory");
            }
          }
        }
      }
      Geo gn;
      if (prep_next) {                                         // in flight under the accumulator wait and the drain
        gn = geometry(next);
        nx_ok = gn.row_ok;
        if (csplit == 0) {
#pragma unroll
          for (int j = 0; j < MAX_XU; ++j) load_unit(gn, sub + 2 * j, xv[j]);
        }
      }
      // depthwise taps and bias of this lane's channel pair (the same for every unit of the chunk)
      const int c_local = 2 * lane;
      const int c_glob = chg * cc + c_local;
      const bool c_ok = c_local < cc && c_glob < Cexp;
      u64 wreg[KS * KS];
      u64 bv2;
      auto load_w = [&]() {
#pragma unroll
        for (int t = 0; t < KS * KS; ++t)
          wreg[t] = c_ok ? __ldg(reinterpret_cast<const u64*>(dw_w + (size_t)t * Cexp + c_glob)) : 0ull;
        bv2 = c_ok ? __ldg(reinterpret_cast<const u64*>(dw_bias + c_glob)) : 0ull;
      };
      if (!prep_next) load_w();                                // (with prep_next the registers hold the next item's rows)
      mbar_wait_warp(acc_full(buf), (gch >> 1) & 1);
      tc_fence_after();
      if (tr) trace(512 + 8 * (gch >> 1) + 1);
      // ---- drain: swish of the accumulator (bias included) into this pixel's slot of the E tile
      {
        const uint32_t e_row = e_base + (uint32_t)(g.e_slot * EP + csplit * CW) * 4u;
        const uint32_t t_acc = t_lane + buf * ACC_STRIDE + (uint32_t)(mt * cc + csplit * CW);
        const u64 inv2 = pk2(inv_wscale, inv_wscale);
#pragma unroll 2
        for (int c0 = 0; c0 < CW; c0 += 8) {
          u64 v[4];
          tmem_ld8_pairs(t_acc + c0, v);
          if (g.row_ok) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = swish2(fmul2(v[i], inv2));
            asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(e_row + c0 * 4), "l"(v[0]), "l"(v[1]) : "memory");
            asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(e_row + c0 * 4 + 16), "l"(v[2]), "l"(v[3]) : "memory");
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty(buf));
      if (tr) trace(512 + 8 * (gch >> 1) + 2);
      if (prep_next) {                                         // the MMAs of the next item start under this chunk's dw
        if (csplit == 0) {
#pragma unroll
          for (int j = 0; j < MAX_XU; ++j)
            if (sub + 2 * j < n_xu) store_unit(nx_ok, sub + 2 * j, xv[j], (local_it + 1) & 1);
        }
        publish_A();
        load_w();
      }
      if (tr) trace(512 + 8 * (gch >> 1) + 3);
      named_bar_sync(1 + group, GROUP_THREADS);                // #1: this group's E complete
      if (tr) trace(512 + 8 * (gch >> 1) + 4);
      // ---- depthwise from E: unit = (x segment, y segment); lane = channel pair
      const int img = g.img, oy0 = g.oy0, ox0 = g.ox0;
      for (int u = gw; u < n_units; u += GROUP_WARPS) {
        const int xs = u % NXS, ys = u / NXS;
        const int oyr0 = ys * RH, oxr0 = xs * NX;
        const int rows_here = min(RH, min(TH, Ho - oy0) - oyr0);
        // Fully unrolled over the NR input rows of the unit: every E value is loaded once and feeds the output rows
        // it touches (known at compile time); the pre-activations of all RH x NX output pairs stay in registers,
        // then all swishes are evaluated together (RH * NX independent chains for the MUFU latency).
        constexpr int NR = (RH - 1) * S + KS;
        u64 o[RH][NX];
#pragma unroll
        for (int oy = 0; oy < RH; ++oy)
#pragma unroll
          for (int xx = 0; xx < NX; ++xx) o[oy][xx] = bv2;
        uint32_t e_ptr = e_base + (uint32_t)(((oyr0 * S) * IW + oxr0 * S) * EP + c_local) * 4u;
        const uint32_t e_row_bytes = (uint32_t)(IW * EP) * 4u;
        constexpr uint32_t e_px_bytes = (uint32_t)EP * 4u;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          u64 v[NIN];
#pragma unroll
          for (int kx = 0; kx < NIN; ++kx)
            asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v[kx]) : "r"(e_ptr + kx * e_px_bytes));
          e_ptr += e_row_bytes;
#pragma unroll
          for (int ky = 0; ky < KS; ++ky) {
            if (r >= ky && (r - ky) % S == 0 && (r - ky) / S < RH) {
              const int oy = (r - ky) / S;
#pragma unroll
              for (int xx = 0; xx < NX; ++xx)
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) o[oy][xx] = ffma2(v[xx * S + kx], wreg[ky * KS + kx], o[oy][xx]);
            }
          }
        }
#pragma unroll
        for (int oy = 0; oy < RH; ++oy)
#pragma unroll
          for (int xx = 0; xx < NX; ++xx) o[oy][xx] = swish2(o[oy][xx]);
        u64 psum = 0ull;
        float* o_ptr = out + (((size_t)img * Ho + oy0 + oyr0) * Wo + ox0 + oxr0) * Cexp + c_glob;
        const int o_row = Wo * Cexp;
#pragma unroll
        for (int oy = 0; oy < RH; ++oy) {
#pragma unroll
          for (int xx = 0; xx < NX; ++xx) {
            if (oy < rows_here && c_ok && oxr0 + xx < TW && ox0 + oxr0 + xx < Wo) {
              psum = fadd2(psum, o[oy][xx]);
              float a, b;
              upk2(o[oy][xx], a, b);
              *reinterpret_cast<float2*>(o_ptr + oy * o_row + xx * Cexp) = make_float2(a, b);
            }
          }
        }
        float pa, pb;
        upk2(psum, pa, pb);
        *reinterpret_cast<float2*>(&s_psg[u * 64 + c_local]) = make_float2(pa, pb);
      }
      if (tr) trace(512 + 8 * (gch >> 1) + 5);
      named_bar_sync(1 + group, GROUP_THREADS);                // #2: E consumed, s_ps complete
      if (tr) trace(512 + 8 * (gch >> 1) + 6);
      if (gtid < cc && chg * cc + gtid < Cexp) {
        float s = 0.f;
        for (int v = 0; v < n_units; ++v) s += s_psg[v * 64 + gtid];
        partial[((size_t)img * tiles + g.tile) * Cexp + chg * cc + gtid] = s;
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
  if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_f16(cc);
    const uint32_t lbo = (uint32_t)cc * 16u;
    const int ksteps = Kp / 16;
    int gch = 0;
    for (int local_it = 0; local_it < my_items; ++local_it) {
      mbar_wait_warp(fullA(), local_it & 1);
      tc_fence_after();
      const uint32_t a_base = tmem_base + A_COL0 + (a_double ? (uint32_t)(local_it & 1) * A_STRIDE : 0u);
      for (int ch = 0; ch < cpg; ++ch, ++gch) {
        const int buf = gch & 1;
        mbar_wait_warp(fullB(buf), (gch >> 1) & 1);
        if (gch >= 2) mbar_wait_warp(acc_empty(buf), ((gch >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t b_hi = b_base + buf * b_bytes, b_lo = b_hi + b_bytes / 2;
        const uint64_t dbh0 = make_desc(b_hi, lbo, 128), dbl0 = make_desc(b_lo, lbo, 128);
        for (int m = 0; m < MT; ++m) {
          const uint32_t d = tmem_base + buf * ACC_STRIDE + (uint32_t)(m * cc);
          const uint32_t a_hi = a_base + (uint32_t)(m * Kp), a_lo = a_hi + (uint32_t)(Kp / 2);
          for (int j = 0; j < ksteps; ++j) {                       // small terms first
            const uint64_t koff = (uint64_t)((j * 2 * lbo) >> 4);
            umma_f16_ts_pred(d, a_lo + j * 8, dbh0 + koff, idesc, j == 0 ? 0u : 1u);
            umma_f16_ts_pred(d, a_hi + j * 8, dbl0 + koff, idesc, 1);
          }
          for (int j = 0; j < ksteps; ++j) {
            const uint64_t koff = (uint64_t)((j * 2 * lbo) >> 4);
            umma_f16_ts_pred(d, a_hi + j * 8, dbh0 + koff, idesc, 1);
          }
        }
        umma_commit_elect(emptyB(buf));
        umma_commit_elect(acc_full(buf));
        if (ch == cpg - 1) umma_commit_elect(a_free());           // every MMA that reads this item's A is done
        __syncwarp();
      }
    }
    tc_fence_before();
  } else if (warp == LOADER_WARP) {
    // ------------------------------------------------------------------ weight loader
    if (lane == 0) {
      int gch = 0;
      for (int local_it = 0; local_it < my_items; ++local_it) {
        const int it = blockIdx.x + local_it * gridDim.x;
        const int cg = it % n_cg;
        for (int ch = 0; ch < cpg; ++ch, ++gch) {
          const int slot = gch & 1;
          if (gch >= 2) mbar_wait(emptyB(slot), ((gch >> 1) - 1) & 1);
          mbar_arrive_expect_tx(fullB(slot), b_bytes);
          bulk_copy_g2s(b_base + slot * b_bytes, Wx + (size_t)(cg * cpg + ch) * Kp * cc * 2, b_bytes, fullB(slot));
        }
      }
    }
  }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side ---------------------------------------------------------------------------------------
inline int kp_for(int cin) { return (cin + 1 + 15) / 16 * 16; }   // + the ones column that carries the bias

// Tile shapes per block.  Early blocks (2-8): the halo tile fills two 128-row MMA tiles, all chunks in one item, two A
// buffers.  Late blocks: the image is split into at most two column strips (15x20) or taken whole (7x10), so that the
// in-image rows fit MT m-tiles; the A operand (Kp up to 240 TMEM columns per m-tile) is single buffered; the 7x10
// blocks split their chunks over two items per hypothesis to occupy the SMs.  Block 25 (Cin = 384) is not planned: its
// double-buffered weight chunks (2 x 77 KB) do not fit next to the two E tiles.
inline Plan make_plan(const BlockSpec& b) {
  Plan p{};
  if (b.e == 1) return p;
  p.Kp = kp_for(b.cin);
  p.n_cg = 1;
  p.MT = 2;
  p.cc = b.cexp % 64 == 0 ? 64 : 48;
  if (b.k == 3 && b.s == 2 && b.hout == 60) { p.TH = 6; p.TW = 8; p.NX = 1; p.NYS = 2; }          // block 2
  else if (b.k == 3 && b.s == 1 && b.hout == 60) { p.TH = 12; p.TW = 16; p.NX = 2; p.NYS = 2; }   // blocks 3-4
  else if (b.k == 5 && b.s == 2 && b.hout == 30) { p.TH = 5; p.TW = 8; p.NX = 1; p.NYS = 2; }     // block 5
  else if (b.k == 5 && b.s == 1 && b.hout == 30) { p.TH = 10; p.TW = 14; p.NX = 2; p.NYS = 2; }   // blocks 6-7
  else if (b.k == 3 && b.s == 2 && b.hout == 15) { p.TH = 5; p.TW = 10; p.NX = 2; p.NYS = 3; }    // block 8
  else if (b.k == 3 && b.s == 1 && b.hout == 15) { p.TH = 15; p.TW = 10; p.NX = 2; p.NYS = 3; }   // blocks 9-12
  else if (b.k == 5 && b.s == 1 && b.hout == 15) { p.TH = 15; p.TW = 10; p.NX = 2; p.NYS = 3; p.cc = 48; }   // 13-17
  else if (b.k == 5 && b.s == 2 && b.hout == 7) { p.TH = 7; p.TW = 5; p.NX = 1; p.NYS = 3; p.cc = 48; }      // 18
  else if (b.k == 5 && b.s == 1 && b.hout == 7) { p.TH = 7; p.TW = 10; p.NX = 2; p.NYS = 2; p.cc = 48; p.MT = 1; p.n_cg = 2; }   // 19-23
  else if (b.k == 3 && b.s == 1 && b.hout == 7 && b.cin < 256) { p.TH = 7; p.TW = 10; p.NX = 2; p.NYS = 2; p.cc = 48; p.MT = 1; p.n_cg = 2; }   // 24
  else return p;
  if (b.cexp % p.cc) return p;
  p.n_chunks = b.cexp / p.cc;
  p.cpg = (p.n_chunks + p.n_cg - 1) / p.n_cg;                  // chunks per item; the weight image is padded to n_cg * cpg
  p.IH = (p.TH - 1) * b.s + b.k;
  p.IW = (p.TW - 1) * b.s + b.k;
  p.tiles_y = (b.hout + p.TH - 1) / p.TH;
  p.tiles_x = (b.wout + p.TW - 1) / p.TW;
  // most in-image rows any tile has
  int max_rows = 0;
  for (int ty = 0; ty < p.tiles_y; ++ty)
    for (int tx = 0; tx < p.tiles_x; ++tx) {
      const int iy0 = ty * p.TH * b.s - b.pad_lo, ix0 = tx * p.TW * b.s - b.pad_lo;
      const int chh = std::min(iy0 + p.IH, b.hin) - std::max(iy0, 0), cw = std::min(ix0 + p.IW, b.win) - std::max(ix0, 0);
      max_rows = std::max(max_rows, chh * cw);
    }
  if (max_rows > p.MT * 128) return p;
  p.a_double = p.MT * (2 * p.cc + 2 * p.Kp) <= 512 && p.Kp / 16 <= 2 * MAX_XU ? 1 : 0;
  if (!p.a_double && p.MT * (2 * p.cc + p.Kp) > 512) return p;
  if (((p.TW + p.NX - 1) / p.NX) * p.NYS > MAX_UNITS) return p;
  p.RH = (p.TH + p.NYS - 1) / p.NYS;
  // E rows: the unclipped window; the last y segment computes (never stores) rows past the tile when TH % RH != 0
  const int e_rows = std::max(p.IH * p.IW, ((p.NYS * p.RH - 1) * b.s + b.k) * p.IW);
  p.e_rows = e_rows + E_SLACK_ROWS;
  p.smem_bytes = 128 + 2 * p.Kp * p.cc * 4 + 2 * p.e_rows * (p.cc + 4) * 4;
  if (p.cpg < 2 || p.smem_bytes > 208 * 1024) return p;   // + 16.5 KB static = the 227 KB an SM offers
  p.ok = 1;
  return p;
}

// Expected relative loss of the accumulator's round-toward-zero adds over the Kp/16 hi*hi MMAs of one chunk (the
// cross terms go in first, while the accumulator is small): sum_j (j / n) half-ulps; the measured unit on same-sign
// operands is 5.1e-8 (kernels_pw2.cuh: 7.7e-8 for n = 2).  Folded into the epilogue scale.
inline float rz_compensation(int Kp) {
  const int n = Kp / 16;
  return 1.0f + 5.1e-8f * 0.5f * (float)(n + 1);
}

// max |.| over the expand weights and their biases: one power-of-two scale serves both (the bias is a weight row)
inline float weight_scale(const float* W_nk, const float* bias, int N, int K) {
  float mx = 0.f;
  for (size_t i = 0; i < (size_t)N * K; ++i) mx = std::max(mx, std::fabs(W_nk[i]));
  for (int i = 0; i < N; ++i) mx = std::max(mx, std::fabs(bias[i]));
  if (!(mx > 0.f) || !std::isfinite(mx)) return 1.f;
  int e;
  std::frexp(mx, &e);
  return std::ldexp(1.f, 13 - e);
}

// W_nk [Cexp][Cin] (BN scale folded), bias [Cexp] -> [n_chunks_padded][hi|lo][Kp/8][cc][8] fp16 bits, bias at k = Cin
inline std::vector<uint16_t> pack_weights(const float* W_nk, const float* bias, int N, int K, int cc, int n_chunks_padded,
                                          float wscale) {
  const int Kp = kp_for(K), nch = n_chunks_padded;
  std::vector<uint16_t> o((size_t)nch * 2 * (Kp / 8) * cc * 8, 0);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k <= K; ++k) {
      const float w = (k < K ? W_nk[(size_t)n * K + k] : bias[n]) * wscale;
      const float h = pw2::host_round11(w);
      const int ch = n / cc, nn = n % cc;
      const size_t hi = ((((size_t)ch * 2 + 0) * (Kp / 8) + k / 8) * cc + nn) * 8 + k % 8;
      const size_t lo = ((((size_t)ch * 2 + 1) * (Kp / 8) + k / 8) * cc + nn) * 8 + k % 8;
      o[hi] = pw2::host_f16_bits(h);
      o[lo] = pw2::host_f16_bits(w - h);
    }
  return o;
}

}  // namespace xdw
}  // namespace cosyb
