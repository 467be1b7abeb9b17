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
//
// One CTA per SM, persistent over items = (hypothesis, spatial tile); an item walks over all chunks of 64 expanded
// channels:
//   workers (16 warps)  per item: x rows of the halo tile -> fp16 hi/lo split -> TMEM (A operand, resident for all
//                       chunks; thread = tile row = TMEM lane);
//                       per chunk: drain the accumulator (tcgen05.ld), bias + swish, zero the out-of-image rows,
//                       -> shared-memory tile E[pixel][channel]; then the depthwise convolution from shared memory
//                       with a rolling register window (lane = channel, conflict free), bias + swish, 128-byte row
//                       stores of D and the pooling partial sums.
//   warp 16             MMA issuer: per chunk and m-tile 3 * Kp/16 kind::f16 MMAs (a_lo*b_hi, a_hi*b_lo, then a_hi*b_hi),
//                       A from TMEM, B from shared memory; accumulators double buffered in TMEM so the MMAs of chunk
//                       c+1 run under the CUDA-core work of chunk c.
//   warp 17             weight loader: one bulk copy (cp.async.bulk + mbarrier) per chunk into a 2-slot ring.
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
using pw2::tmem_ld16_nowait;
using pw2::tmem_ld_wait;
using pw2::umma_commit_elect;
using pw2::umma_f16_ts_pred;

constexpr int NWW = 16;                 // worker warps
constexpr int WORKERS = NWW * 32;
constexpr int MMA_WARP = NWW, LOADER_WARP = NWW + 1;
constexpr int THREADS = (NWW + 2) * 32;
constexpr int CC = 64;                  // expanded channels per chunk (= MMA N)
constexpr int EP = CC + 4;              // floats per E row: 16-byte row stores of 8 consecutive rows are conflict free
constexpr int E_SLACK_ROWS = 16;        // the last x-segment of a tile may read (never use) a few pixels past the tile
constexpr int MAX_UNITS = 64;
constexpr uint32_t TMEM_COLS = 512;

struct Plan {
  int ok, MT, TH, TW, IH, IW, tiles_y, tiles_x, n_chunks, Kp, NX, NYS, smem_bytes;
};

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// Wx: [n_chunks][hi|lo][Kp/8][CC][8] fp16 (rows >= Cexp zero); ebias [n_chunks*CC] (zero padded)
// x [B][H][W][Cin], out [B][Ho][Wo][Cexp], partial [B][tiles][Cexp]
template <int KS, int S, int NX>
__global__ void __launch_bounds__(THREADS, 1)
k_xdw(const float* __restrict__ x, const __half* __restrict__ Wx, const float* __restrict__ ebias, float inv_wscale,
      const float* __restrict__ dw_w, const float* __restrict__ dw_bias, float* __restrict__ out,
      float* __restrict__ partial, int B, int H, int W, int Cin, int Cexp, int Ho, int Wo, int pad, int MT, int TH,
      int TW, int IH, int IW, int tiles_y, int tiles_x, int n_chunks, int Kp, int NYS) {
  constexpr int NSLOT = (KS + S - 1) / S;
  constexpr int PERIOD = S * NSLOT;
  constexpr int NIN = (NX - 1) * S + KS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[10];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_ebias[CC];
  __shared__ float s_ps[MAX_UNITS * 32];
  const uint32_t b_base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t b_bytes = (uint32_t)Kp * CC * 4u;                  // hi + lo image of one chunk
  float* E = reinterpret_cast<float*>(smem_raw + (b_base - smem_u32(smem_raw)) + 2 * b_bytes);
  const int tid = threadIdx.x, lane = tid % 32;
  const int warp = __shfl_sync(0xffffffffu, tid / 32, 0);
  const int tiles = tiles_y * tiles_x;
  const int n_items = B * tiles;
  const int n_rows = IH * IW;
  const uint32_t ACC_STRIDE = (uint32_t)MT * CC;                    // columns of one accumulator buffer
  const uint32_t A_COL0 = 2 * ACC_STRIDE;
  auto fullA = [&]() { return smem_u32(&bars[0]); };
  auto fullB = [&](int s) { return smem_u32(&bars[1 + s]); };
  auto emptyB = [&](int s) { return smem_u32(&bars[3 + s]); };
  auto acc_full = [&](int b) { return smem_u32(&bars[5 + b]); };
  auto acc_empty = [&](int b) { return smem_u32(&bars[7 + b]); };

  if (tid == 0) {
    mbar_init(fullA(), WORKERS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(fullB(s), 1);
      mbar_init(emptyB(s), 1);
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), WORKERS);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(&s_tmem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, s_tmem, 0);

  if (warp < NWW) {
    // ------------------------------------------------------------------ workers
    const int q = warp % 4, part = warp / 4;
    const int nsub = 4 / MT;                                   // sub-parts (k units / column ranges) per m-tile
    const int mt = part % MT, sub = part / MT;
    const bool part_on = part < MT * nsub;
    const int row = mt * 128 + q * 32 + lane;                  // tile row (halo pixel) == TMEM lane of m-tile mt
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int py = row / IW, px = row % IW;
    const int CW = CC / nsub;                                  // accumulator columns this thread drains per chunk
    int gch = 0;                                               // chunks processed by this CTA so far
    int local_it = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++local_it) {
      const int img = it / tiles, tile = it % tiles;
      const int oy0 = (tile / tiles_x) * TH, ox0 = (tile % tiles_x) * TW;
      const int gy = oy0 * S - pad + py, gx = ox0 * S - pad + px;
      const bool row_valid = row < n_rows && gy >= 0 && gy < H && gx >= 0 && gx < W;
      // ---- A operand: this thread's halo pixel, 16 k per unit, -> hi/lo fp16 -> TMEM
      if (part_on) {
        const float* xr = x + (((size_t)img * H + (row_valid ? gy : 0)) * W + (row_valid ? gx : 0)) * Cin;
        for (int u = sub; u < Kp / 16; u += nsub) {
          float v[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_valid && u * 16 + c * 4 < Cin) t = __ldg(reinterpret_cast<const float4*>(xr + u * 16 + c * 4));
            v[c * 4] = t.x; v[c * 4 + 1] = t.y; v[c * 4 + 2] = t.z; v[c * 4 + 3] = t.w;
          }
          uint32_t ph[8], pl[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float h0, l0, h1, l1;
            split11(v[2 * c], h0, l0);
            split11(v[2 * c + 1], h1, l1);
            ph[c] = pack_f16x2(h0, h1);
            pl[c] = pack_f16x2(l0, l1);
          }
          const uint32_t a_hi = t_lane + A_COL0 + (uint32_t)(mt * Kp) + (uint32_t)(u * 8);
          tmem_st8(a_hi, ph);
          tmem_st8(a_hi + (uint32_t)(Kp / 2), pl);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(fullA());

      for (int ch = 0; ch < n_chunks; ++ch, ++gch) {
        const int buf = gch & 1;
        const int cc_here = min(CC, Cexp - ch * CC);
        if (tid < CC) s_ebias[tid] = __ldg(ebias + ch * CC + tid);   // read after barrier #0 below
        mbar_wait_warp(acc_full(buf), (gch >> 1) & 1);
        tc_fence_after();
        named_bar_sync(1, WORKERS);                                   // #0: s_ebias visible, previous chunk's readers done
        // ---- drain: bias + swish, out-of-image rows are exact zeros
        if (part_on) {
          float* erow = E + (size_t)row * EP + sub * CW;
          const uint32_t t_acc = t_lane + buf * ACC_STRIDE + (uint32_t)(mt * CC + sub * CW);
          for (int c0 = 0; c0 < CW; c0 += 16) {
            float v[16];
            tmem_ld16_nowait(t_acc + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float o = swishf(fmaf(v[i], inv_wscale, s_ebias[sub * CW + c0 + i]));
              v[i] = row_valid ? o : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(erow + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive(acc_empty(buf));
        named_bar_sync(1, WORKERS);                                   // #1: E complete
        // ---- depthwise from E: unit = (32-channel group, x segment, y segment); lane = channel
        const int n_groups = (cc_here + 31) / 32;
        const int NXS = (TW + NX - 1) / NX;
        const int RH = (TH + NYS - 1) / NYS;
        const int n_units = n_groups * NXS * NYS;
        for (int u = warp; u < n_units; u += NWW) {
          const int g = u % n_groups, xs = (u / n_groups) % NXS, ys = u / (n_groups * NXS);
          const int c_local = g * 32 + lane;
          const int c_glob = ch * CC + c_local;
          const bool c_ok = c_local < cc_here;
          const int oyr0 = ys * RH, oxr0 = xs * NX;
          const int rows_here = min(RH, min(TH, Ho - oy0) - oyr0);
          float wreg[KS * KS];
#pragma unroll
          for (int t = 0; t < KS * KS; ++t) wreg[t] = c_ok ? __ldg(dw_w + (size_t)t * Cexp + c_glob) : 0.f;
          const float bv = c_ok ? __ldg(dw_bias + c_glob) : 0.f;
          float acc[NSLOT][NX];
#pragma unroll
          for (int s = 0; s < NSLOT; ++s)
#pragma unroll
            for (int xx = 0; xx < NX; ++xx) acc[s][xx] = 0.f;
          float psum = 0.f;
          const float* e0 = E + ((size_t)(oyr0 * S) * IW + oxr0 * S) * EP + c_local;
          float* o0 = out + (((size_t)img * Ho + oy0 + oyr0) * Wo + ox0 + oxr0) * Cexp + c_glob;
          const int n_in_rows = rows_here > 0 ? (rows_here - 1) * S + KS : 0;
          for (int r0 = 0; r0 < n_in_rows; r0 += PERIOD) {
#pragma unroll
            for (int j = 0; j < PERIOD; ++j) {
              const int r = r0 + j;
              if (r < n_in_rows) {
                float v[NIN];
#pragma unroll
                for (int kx = 0; kx < NIN; ++kx) v[kx] = e0[((size_t)r * IW + kx) * EP];
#pragma unroll
                for (int ky = 0; ky < KS; ++ky) {
                  if ((j - ky + PERIOD * KS) % S == 0) {
                    const int slot = ((j - ky + PERIOD * KS) / S) % NSLOT;
#pragma unroll
                    for (int xx = 0; xx < NX; ++xx)
#pragma unroll
                      for (int kx = 0; kx < KS; ++kx)
                        acc[slot][xx] = fmaf(v[xx * S + kx], wreg[ky * KS + kx], acc[slot][xx]);
                  }
                }
              }
              if ((j - (KS - 1) + PERIOD * KS) % S == 0) {
                const int slot_done = ((j - (KS - 1) + PERIOD * KS) / S) % NSLOT;
                const int num = r - (KS - 1);
                const int oyr = num >= 0 ? num / S : -1;
                if (oyr >= 0 && oyr < rows_here && c_ok) {
#pragma unroll
                  for (int xx = 0; xx < NX; ++xx) {
                    if (oxr0 + xx < TW && ox0 + oxr0 + xx < Wo) {
                      const float o = swishf(acc[slot_done][xx] + bv);
                      psum += o;
                      o0[((size_t)oyr * Wo + xx) * Cexp] = o;
                    }
                  }
                }
#pragma unroll
                for (int xx = 0; xx < NX; ++xx) acc[slot_done][xx] = 0.f;
              }
            }
          }
          s_ps[u * 32 + lane] = psum;
        }
        named_bar_sync(1, WORKERS);                                   // #2: E consumed, s_ps complete
        if (tid < cc_here) {
          const int g = tid / 32, l = tid % 32;
          float s = 0.f;
          for (int v = 0; v < NXS * NYS; ++v) s += s_ps[(v * n_groups + g) * 32 + l];
          partial[((size_t)img * tiles + tile) * Cexp + ch * CC + tid] = s;
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_f16(CC);
    const uint32_t lbo = (uint32_t)CC * 16u;
    const int ksteps = Kp / 16;
    int gch = 0, local_it = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++local_it) {
      mbar_wait_warp(fullA(), local_it & 1);
      tc_fence_after();
      for (int ch = 0; ch < n_chunks; ++ch, ++gch) {
        const int buf = gch & 1;
        mbar_wait_warp(fullB(buf), (gch >> 1) & 1);
        if (gch >= 2) mbar_wait_warp(acc_empty(buf), ((gch >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t b_hi = b_base + buf * b_bytes, b_lo = b_hi + b_bytes / 2;
        const uint64_t dbh0 = make_desc(b_hi, lbo, 128), dbl0 = make_desc(b_lo, lbo, 128);
        for (int m = 0; m < MT; ++m) {
          const uint32_t d = tmem_base + buf * ACC_STRIDE + (uint32_t)(m * CC);
          const uint32_t a_hi = tmem_base + A_COL0 + (uint32_t)(m * Kp), a_lo = a_hi + (uint32_t)(Kp / 2);
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
        __syncwarp();
      }
    }
    tc_fence_before();
  } else {
    // ------------------------------------------------------------------ weight loader
    if (lane == 0) {
      int gch = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        for (int ch = 0; ch < n_chunks; ++ch, ++gch) {
          const int slot = gch & 1;
          if (gch >= 2) mbar_wait(emptyB(slot), ((gch >> 1) - 1) & 1);
          mbar_arrive_expect_tx(fullB(slot), b_bytes);
          bulk_copy_g2s(b_base + slot * b_bytes, Wx + (size_t)ch * Kp * CC * 2, b_bytes, fullB(slot));
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
inline int kp_for(int cin) { return (cin + 15) / 16 * 16; }
inline int n_chunks_for(int cexp) { return (cexp + CC - 1) / CC; }

// tile shapes chosen so that the halo tile fills two 128-row MMA tiles; blocks whose input has too many channels
// for a TMEM-resident A operand (MT * (128 + Kp) > 512) are not planned
inline Plan make_plan(const BlockSpec& b) {
  Plan p{};
  if (b.e == 1) return p;
  p.Kp = kp_for(b.cin);
  p.n_chunks = n_chunks_for(b.cexp);
  p.MT = 2;
  if (b.k == 3 && b.s == 2 && b.hout == 60) { p.TH = 6; p.TW = 8; p.NX = 2; p.NYS = 2; }          // block 2
  else if (b.k == 3 && b.s == 1 && b.hout == 60) { p.TH = 12; p.TW = 16; p.NX = 4; p.NYS = 2; }   // blocks 3-4
  else if (b.k == 5 && b.s == 2 && b.hout == 30) { p.TH = 5; p.TW = 8; p.NX = 2; p.NYS = 2; }     // block 5
  else if (b.k == 5 && b.s == 1 && b.hout == 30) { p.TH = 10; p.TW = 14; p.NX = 4; p.NYS = 2; }   // blocks 6-7
  else if (b.k == 3 && b.s == 2 && b.hout == 15) { p.TH = 5; p.TW = 10; p.NX = 2; p.NYS = 2; }    // block 8
  else return p;
  p.IH = (p.TH - 1) * b.s + b.k;
  p.IW = (p.TW - 1) * b.s + b.k;
  if (p.IH * p.IW > p.MT * 128 || p.MT * (128 + p.Kp) > 512) return p;
  p.tiles_y = (b.hout + p.TH - 1) / p.TH;
  p.tiles_x = (b.wout + p.TW - 1) / p.TW;
  p.smem_bytes = 128 + 2 * p.Kp * CC * 4 + (p.MT * 128 + E_SLACK_ROWS) * EP * 4;
  p.ok = 1;
  return p;
}

// W_nk [Cexp][Cin] (BN scale folded) -> [n_chunks][hi|lo][Kp/8][CC][8] fp16 bits
inline std::vector<uint16_t> pack_weights(const float* W_nk, int N, int K, float wscale) {
  const int Kp = kp_for(K), nch = n_chunks_for(N);
  std::vector<uint16_t> o((size_t)nch * 2 * (Kp / 8) * CC * 8, 0);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float w = W_nk[(size_t)n * K + k] * wscale;
      const float h = pw2::host_round11(w);
      const int ch = n / CC, nn = n % CC;
      const size_t hi = ((((size_t)ch * 2 + 0) * (Kp / 8) + k / 8) * CC + nn) * 8 + k % 8;
      const size_t lo = ((((size_t)ch * 2 + 1) * (Kp / 8) + k / 8) * CC + nn) * 8 + k % 8;
      o[hi] = pw2::host_f16_bits(h);
      o[lo] = pw2::host_f16_bits(w - h);
    }
  return o;
}

}  // namespace xdw
}  // namespace cosyb
