// EfficientNet-B3 trunk kernels, fp32, NHWC activations (reference: models/efficientnet.py:71-98,
// 174-190; BN eval eps 1e-3 folded at load; swish = x*sigmoid(x), efficientnet_utils.py:37-57).
//   k_stem        3x3 s2 6->40 + bias + swish; reads the crop and render planes directly (the
//                 6-channel concat of pose.py:104 never exists in memory)
//   k_pw_gemm     1x1 convolution as a row-major GEMM on CUDA cores with fused bias / swish /
//                 SE gate on the A operand / residual add (parity anchor; the product path runs
//                 kernels_tc.cuh: tcgen05 3xTF32)
//   k_dwconv_roll depthwise kxk (k3/k5, s1/s2, static asymmetric "same" padding) + bias + swish with
//                 a register rolling window; emits deterministic per-tile channel sums for the squeeze step
//   k_se_gate     squeeze-excite: tile sums -> mean -> FC+swish -> FC+sigmoid -> gate[B][Cexp]
//   k_pool_fc_update  mean pool over 7x10, Linear(1536,9), 6D->R + image-space pose update
#pragma once
#include "common.h"
#include "kernels_geometry.cuh"

namespace cosyb {

// x * sigmoid(x) with the hardware exp2 / reciprocal approximations (5 instructions: FMUL, MUFU.EX2, FADD,
// MUFU.RCP, FMUL; ~2 ulp).  The accurate expf costs ~15 instructions per element and made the epilogue of
// the expand convolutions the slowest stage of the pipeline (measured with the in-kernel cycle trace).
// ex2.approx.ftz instead of __expf: the non-ftz form wraps MUFU.EX2 in a compare and two predicated multiplies to
// keep denormal results, which 1 + e then rounds away anyway - same bits out, 3 instructions fewer per element
// in kernels that are bound by instruction issue.
__device__ __forceinline__ float exp_neg_fast(float v) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  return e;
}
__device__ __forceinline__ float swishf(float v) { return __fdividef(v, 1.0f + exp_neg_fast(v)); }
__device__ __forceinline__ float sigmoidf_(float v) { return __fdividef(1.0f, 1.0f + exp_neg_fast(v)); }

// ------------------------------------------------------------------------------------------ stem
constexpr int STEM_TX = 32, STEM_TY = 8;                    // output tile
constexpr int STEM_IW = 2 * STEM_TX + 1, STEM_IH = 2 * STEM_TY + 1;  // 65 x 17 input tile
constexpr int STEM_IWP = STEM_IW + 1;                        // padded row
constexpr int STEM_CO = 40;
constexpr int STEM_SMEM = STEM_TX * STEM_TY * STEM_CO;       // 10240 floats, aliased for the epilogue
static_assert(IN_CH * STEM_IH * STEM_IWP + 54 * STEM_CO <= STEM_SMEM, "stem smem");

// U8: the rendered view arrives as uint8 NHWC [B][240][320][3], the format the reference's renderer
// hands over before `.float() / 255` (rendering/bullet_batch_renderer.py:70-83); converted here.
template <bool U8>
__global__ void __launch_bounds__(STEM_TX* STEM_TY)
k_stem(const float* __restrict__ crops, const void* __restrict__ renders_any, const __grid_constant__ StemWeights wts,
       float* __restrict__ out) {
  constexpr int H = RENDER_H, W = RENDER_W, HO = H / 2, WO = W / 2;
  __shared__ __align__(16) float smem[STEM_SMEM];
  float* s_in = smem;                                // [6][17][66]
  const int b = blockIdx.z;
  const int ox0 = blockIdx.x * STEM_TX, oy0 = blockIdx.y * STEM_TY;
  const int tid = threadIdx.y * STEM_TX + threadIdx.x;
  for (int i = tid; i < IN_CH * STEM_IH * STEM_IW; i += STEM_TX * STEM_TY) {
    int c = i / (STEM_IH * STEM_IW), r = i % (STEM_IH * STEM_IW);
    int iy = r / STEM_IW, ix = r % STEM_IW;
    int gy = 2 * oy0 + iy, gx = 2 * ox0 + ix;   // pad (0,1): only the high side can fall outside
    float v = 0.f;
    if (gy < H && gx < W) {
      if (c < 3) {
        v = __ldg(crops + (((size_t)b * 3 + c) * H + gy) * W + gx);
      } else if (U8) {
        const uint8_t* r8 = static_cast<const uint8_t*>(renders_any);
        v = __fdiv_rn((float)__ldg(r8 + (((size_t)b * H + gy) * W + gx) * 3 + (c - 3)), 255.0f);
      } else {
        const float* rf = static_cast<const float*>(renders_any);
        v = __ldg(rf + (((size_t)b * 3 + (c - 3)) * H + gy) * W + gx);
      }
    }
    s_in[(c * STEM_IH + iy) * STEM_IWP + ix] = v;
  }
  __syncthreads();
  float acc[STEM_CO];
#pragma unroll
  for (int i = 0; i < STEM_CO; ++i) acc[i] = 0.f;
  const int tx = threadIdx.x, ty = threadIdx.y;
  // fully unrolled: every weight is a constant-bank operand of its FFMA (the shared-memory pipe was the co-bottleneck
  // with 10 broadcast LDS.128 per 40 FFMA: profiles/r02_ncu_full_summary.csv, 54 % issue-active)
#pragma unroll
  for (int c = 0; c < IN_CH; ++c) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float v = s_in[(c * STEM_IH + 2 * ty + ky) * STEM_IWP + 2 * tx + kx];
#pragma unroll
        for (int q = 0; q < STEM_CO; ++q) acc[q] = fmaf(v, wts.w[(ky * 3 + kx) * IN_CH + c][q], acc[q]);
      }
    }
  }
  __syncthreads();  // everyone is done with s_in: reuse as the output staging tile
  float4* s_out = reinterpret_cast<float4*>(smem);
#pragma unroll
  for (int q = 0; q < STEM_CO / 4; ++q) {
    float4 o;
    o.x = swishf(acc[4 * q + 0] + wts.bias[4 * q + 0]);
    o.y = swishf(acc[4 * q + 1] + wts.bias[4 * q + 1]);
    o.z = swishf(acc[4 * q + 2] + wts.bias[4 * q + 2]);
    o.w = swishf(acc[4 * q + 3] + wts.bias[4 * q + 3]);
    s_out[tid * (STEM_CO / 4) + q] = o;
  }
  __syncthreads();
  // each tile row is STEM_TX*40 contiguous floats in the NHWC output
  constexpr int ROW_V4 = STEM_TX * STEM_CO / 4;
  for (int i = tid; i < STEM_TY * ROW_V4; i += STEM_TX * STEM_TY) {
    int r = i / ROW_V4, j = i % ROW_V4;
    float4* dst = reinterpret_cast<float4*>(out + (((size_t)b * HO + oy0 + r) * WO + ox0) * STEM_CO);
    dst[j] = s_out[r * ROW_V4 + j];
  }
}

// ------------------------------------------------------------------------------- pointwise GEMM
// C[M][N] = epi( (A[M][K] (*gate[m / rows_per_img][k])) @ Wkn[K][N] + bias[N] ) (+ resid[M][N])
// K % 4 == 0 and N % 4 == 0 (all trunk widths are multiples of 8).
constexpr int GEMM_BK = 16;
constexpr int GEMM_THREADS = 256;

template <int BM, int BN, int TM, int TN, bool GATE, bool SWISH, bool RESID>
__global__ void __launch_bounds__(GEMM_THREADS)
k_pw_gemm(const float* __restrict__ A, const float* __restrict__ Wkn, const float* __restrict__ bias,
          const float* __restrict__ gate, const float* __restrict__ resid, float* __restrict__ C,
          int M, int N, int K, int rows_per_img) {
  static_assert((BM / TM) * (BN / TN) == GEMM_THREADS, "thread tiling");
  static_assert(TM % 4 == 0 && TN % 4 == 0, "float4 fragments");
  constexpr int APAD = 4;
  constexpr int A_V4 = BM * GEMM_BK / 4 / GEMM_THREADS;  // float4 loads of A per thread
  constexpr int B_V4 = GEMM_BK * BN / 4;                 // float4 loads of B per CTA
  static_assert(BM * GEMM_BK / 4 % GEMM_THREADS == 0, "A tile load");
  __shared__ __align__(16) float As[GEMM_BK][BM + APAD];
  __shared__ __align__(16) float Bs[GEMM_BK][BN];
  const int tid = threadIdx.x;
  const int tn = tid % (BN / TN), tm = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 a_reg[A_V4];
  float4 b_reg[(B_V4 + GEMM_THREADS - 1) / GEMM_THREADS];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_V4; ++i) {
      int idx = tid + i * GEMM_THREADS;
      int row = idx / (GEMM_BK / 4), kq = idx % (GEMM_BK / 4);
      int m = m0 + row, k = k0 + kq * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M && k < K) {
        v = *reinterpret_cast<const float4*>(A + (size_t)m * K + k);
        if (GATE) {
          float4 g = __ldg(reinterpret_cast<const float4*>(gate + (size_t)(m / rows_per_img) * K + k));
          v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
        }
      }
      a_reg[i] = v;
    }
#pragma unroll
    for (int i = 0; i < (B_V4 + GEMM_THREADS - 1) / GEMM_THREADS; ++i) {
      int idx = tid + i * GEMM_THREADS;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < B_V4) {
        int kr = idx / (BN / 4), nq = idx % (BN / 4);
        int k = k0 + kr, n = n0 + nq * 4;
        if (k < K && n < N) v = __ldg(reinterpret_cast<const float4*>(Wkn + (size_t)k * N + n));
      }
      b_reg[i] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < A_V4; ++i) {
      int idx = tid + i * GEMM_THREADS;
      int row = idx / (GEMM_BK / 4), kq = idx % (GEMM_BK / 4);
      As[kq * 4 + 0][row] = a_reg[i].x;
      As[kq * 4 + 1][row] = a_reg[i].y;
      As[kq * 4 + 2][row] = a_reg[i].z;
      As[kq * 4 + 3][row] = a_reg[i].w;
    }
#pragma unroll
    for (int i = 0; i < (B_V4 + GEMM_THREADS - 1) / GEMM_THREADS; ++i) {
      int idx = tid + i * GEMM_THREADS;
      if (idx < B_V4) {
        int kr = idx / (BN / 4), nq = idx % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[kr][nq * 4]) = b_reg[i];
      }
    }
  };

  load_tiles(0);
  for (int k0 = 0; k0 < K; k0 += GEMM_BK) {
    store_tiles();
    __syncthreads();
    if (k0 + GEMM_BK < K) load_tiles(k0 + GEMM_BK);
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; ++kk) {
      float a[TM], bb[TN];
#pragma unroll
      for (int i = 0; i < TM / 4; ++i) {
        float4 v = *reinterpret_cast<const float4*>(&As[kk][tm * TM + i * 4]);
        a[i * 4 + 0] = v.x; a[i * 4 + 1] = v.y; a[i * 4 + 2] = v.z; a[i * 4 + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN / 4; ++j) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[kk][tn * TN + j * 4]);
        bb[j * 4 + 0] = v.x; bb[j * 4 + 1] = v.y; bb[j * 4 + 2] = v.z; bb[j * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + tm * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN / 4; ++j) {
      int n = n0 + tn * TN + j * 4;
      if (n >= N) continue;
      float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n));
      float4 o;
      o.x = acc[i][j * 4 + 0] + bv.x;
      o.y = acc[i][j * 4 + 1] + bv.y;
      o.z = acc[i][j * 4 + 2] + bv.z;
      o.w = acc[i][j * 4 + 3] + bv.w;
      if (SWISH) { o.x = swishf(o.x); o.y = swishf(o.y); o.z = swishf(o.z); o.w = swishf(o.w); }
      if (RESID) {
        float4 r = *reinterpret_cast<const float4*>(resid + (size_t)m * N + n);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      *reinterpret_cast<float4*>(C + (size_t)m * N + n) = o;
    }
  }
}

// ------------------------------------------------------------------------------------ depthwise
// in [B][H][W][C] -> out [B][Ho][Wo][C]; deterministic per-tile channel sums (no atomics).
constexpr int DW_MAX_THREADS = 256;

// Depthwise with a rolling window (stride 1 or 2): a thread owns (channel vector, output column) and
// walks down TH output rows.  Each input row is loaded once per thread (KS vector loads, horizontally
// shared through L1) and accumulated into the ceil(KS/S) output rows it touches, held in a register
// ring; the KS*KS taps of the thread's channels stay in registers.  Loads per output drop from KS^2
// to ~KS*S.
//   grid = (tiles_x * tiles_y, n_chunks, B); block = Gc * PX threads
template <int V> struct VecT;
template <> struct VecT<4> { using T = float4; };
template <> struct VecT<2> { using T = float2; };
__device__ __forceinline__ float4 pack_vec(const float (&a)[4]) { return make_float4(a[0], a[1], a[2], a[3]); }
__device__ __forceinline__ float2 pack_vec(const float (&a)[2]) { return make_float2(a[0], a[1]); }
template <int V> __device__ __forceinline__ void vzero(float* a) {
#pragma unroll
  for (int i = 0; i < V; ++i) a[i] = 0.f;
}

template <int KS, int S, int V, int NX>
__global__ void __launch_bounds__(DW_MAX_THREADS)
k_dwconv_roll(const float* __restrict__ in, const float* __restrict__ w /*[KS*KS][C]*/,
              const float* __restrict__ bias, float* __restrict__ out, float* __restrict__ partial,
              int H, int W, int C, int Ho, int Wo, int pad, int Gc, int PX, int TH, int tiles_x,
              int tiles_per_img) {
  using VT = typename VecT<V>::T;
  constexpr int NSLOT = (KS + S - 1) / S;      // output rows in flight per thread
  constexpr int PERIOD = S * NSLOT;            // the (input row -> slot, tap row) pattern repeats with this period
  constexpr int NIN = (NX - 1) * S + KS;       // input columns feeding the thread's NX adjacent outputs
  __shared__ float sred[DW_MAX_THREADS * V];
  const int b = blockIdx.z, tile = blockIdx.x;
  const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
  const int tid = threadIdx.x;
  const int c = (blockIdx.y * Gc + tid % Gc) * V;
  const int ox0 = (tile_x * PX + tid / Gc) * NX;
  const int oy0 = tile_y * TH;
  const bool col_ok = ox0 < Wo;
  const float* inb = in + (size_t)b * H * W * C + c;
  float* outb = out + (size_t)b * Ho * Wo * C + c;

  float wreg[KS * KS][V];
#pragma unroll
  for (int t = 0; t < KS * KS; ++t) {
    VT wv = __ldg(reinterpret_cast<const VT*>(w + (size_t)t * C + c));
    const float* wp = reinterpret_cast<const float*>(&wv);
#pragma unroll
    for (int i = 0; i < V; ++i) wreg[t][i] = wp[i];
  }
  float bv[V];
  {
    VT t = __ldg(reinterpret_cast<const VT*>(bias + c));
    const float* tp = reinterpret_cast<const float*>(&t);
#pragma unroll
    for (int i = 0; i < V; ++i) bv[i] = tp[i];
  }
  float acc[NSLOT][NX][V];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s)
#pragma unroll
    for (int x = 0; x < NX; ++x) vzero<V>(acc[s][x]);
  float psum[V];
  vzero<V>(psum);

  // input rows r = 0 .. (TH-1)*S + KS - 1 relative to iy0 = oy0*S - pad; input row r feeds output row
  // (r - ky) / S with tap row ky whenever that division is exact.
  const int n_in_rows = (TH - 1) * S + KS;
  const int iy0 = oy0 * S - pad, ix0 = ox0 * S - pad;
  for (int r0 = 0; r0 < n_in_rows; r0 += PERIOD) {
#pragma unroll
    for (int j = 0; j < PERIOD; ++j) {
      const int r = r0 + j;
      const int iy = iy0 + r;
      if (r < n_in_rows && iy >= 0 && iy < H && col_ok) {
        float v[NIN][V];
#pragma unroll
        for (int kx = 0; kx < NIN; ++kx) {
          const int ix = ix0 + kx;
          if (ix >= 0 && ix < W) {
            VT t = *reinterpret_cast<const VT*>(inb + ((size_t)iy * W + ix) * C);
            const float* tp = reinterpret_cast<const float*>(&t);
#pragma unroll
            for (int i = 0; i < V; ++i) v[kx][i] = tp[i];
          } else {
            vzero<V>(v[kx]);
          }
        }
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          if ((j - ky + PERIOD * KS) % S == 0) {
            const int slot = ((j - ky + PERIOD * KS) / S) % NSLOT;
#pragma unroll
            for (int x = 0; x < NX; ++x)
#pragma unroll
              for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                for (int i = 0; i < V; ++i)
                  acc[slot][x][i] = fmaf(v[x * S + kx][i], wreg[ky * KS + kx][i], acc[slot][x][i]);
          }
        }
      }
      // output row (r - KS + 1) / S has now seen all of its input rows
      if ((j - (KS - 1) + PERIOD * KS) % S == 0) {
        const int slot_done = ((j - (KS - 1) + PERIOD * KS) / S) % NSLOT;
        const int num = r - (KS - 1);
        const int oyr = num >= 0 ? num / S : -1;
        if (oyr >= 0 && oyr < TH && oy0 + oyr < Ho && col_ok) {
#pragma unroll
          for (int x = 0; x < NX; ++x) {
            if (ox0 + x < Wo) {
              float o[V];
#pragma unroll
              for (int i = 0; i < V; ++i) {
                o[i] = swishf(acc[slot_done][x][i] + bv[i]);
                psum[i] += o[i];
              }
              *reinterpret_cast<VT*>(outb + ((size_t)(oy0 + oyr) * Wo + ox0 + x) * C) = pack_vec(o);
            }
          }
        }
#pragma unroll
        for (int x = 0; x < NX; ++x) vzero<V>(acc[slot_done][x]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) sred[tid * V + i] = psum[i];
  __syncthreads();
  if (tid < Gc) {
    float s[V];
#pragma unroll
    for (int i = 0; i < V; ++i) s[i] = sred[tid * V + i];
    for (int q = 1; q < PX; ++q)
#pragma unroll
      for (int i = 0; i < V; ++i) s[i] += sred[(q * Gc + tid) * V + i];
    *reinterpret_cast<VT*>(partial + ((size_t)b * tiles_per_img + tile) * C + c) = pack_vec(s);
  }
}

// -------------------------------------------------------------------------------- squeeze-excite
// gate[b][c] = sigmoid(be[c] + sum_j We[c][j] * swish(br[j] + sum_c' Wr[j][c'] * mean[c']))
// (reference: models/efficientnet.py:85-88).  grid = (ceil(B / SE_IMGS), n_split): a CTA serves SE_IMGS
// hypotheses so that each weight element it reads from L2 is used SE_IMGS times; every CTA rebuilds the
// channel means and the reduced vectors (cheap), then produces its slice of the gates.  `we_t` is the
// expand weight transposed to [Cse][C] so that the last stage reads contiguously.
// dynamic smem: SE_IMGS * (C + Cse) floats.
constexpr int SE_THREADS = 1024;
constexpr int SE_IMGS = 8;       // wide blocks: a CTA serves 8 hypotheses so that each FC weight read from L2 is used 8 times
constexpr int SE_IMGS_SMALL = 2; // blocks 0-8 (C <= 288, tiny FCs, up to 100 tile partials per channel): 4x the CTAs instead

template <int SE_IMGS>
__global__ void __launch_bounds__(SE_THREADS)
k_se_gate(int B, const float* __restrict__ partial, int tiles_per_img, int C, int Cse, float inv_hw,
          const float* __restrict__ wr, const float* __restrict__ br, const float* __restrict__ we_t,
          const float* __restrict__ be, float* __restrict__ gate) {
  extern __shared__ float se_smem[];
  float* s_mean = se_smem;                 // [SE_IMGS][C]
  float* s_r = se_smem + SE_IMGS * C;      // [SE_IMGS][Cse]
  const int b0 = blockIdx.x * SE_IMGS, tid = threadIdx.x;
  const int n_img = min(SE_IMGS, B - b0);
  for (int i = tid; i < SE_IMGS * C; i += SE_THREADS) {
    const int im = i / C, c = i % C;
    float s = 0.f;
    if (im < n_img) {
      const float* pb = partial + ((size_t)(b0 + im) * tiles_per_img) * C + c;
      int t = 0;
      for (; t + 4 <= tiles_per_img; t += 4) {      // independent loads in flight, fixed summation order
        const float v0 = pb[(size_t)t * C], v1 = pb[(size_t)(t + 1) * C], v2 = pb[(size_t)(t + 2) * C], v3 = pb[(size_t)(t + 3) * C];
        s += v0; s += v1; s += v2; s += v3;
      }
      for (; t < tiles_per_img; ++t) s += pb[(size_t)t * C];
    }
    s_mean[i] = s * inv_hw;
  }
  __syncthreads();
  const int warp = tid / 32, lane = tid % 32;
  for (int j = warp; j < Cse; j += SE_THREADS / 32) {
    const float* wrow = wr + (size_t)j * C;
    float acc[SE_IMGS];
#pragma unroll
    for (int im = 0; im < SE_IMGS; ++im) acc[im] = 0.f;
    for (int c = lane * 4; c < C; c += 128) {   // C is a multiple of 8
      const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow + c));
#pragma unroll
      for (int im = 0; im < SE_IMGS; ++im) {
        const float4 mv = *reinterpret_cast<const float4*>(s_mean + im * C + c);
        acc[im] = fmaf(wv.x, mv.x, fmaf(wv.y, mv.y, fmaf(wv.z, mv.z, fmaf(wv.w, mv.w, acc[im]))));
      }
    }
#pragma unroll
    for (int im = 0; im < SE_IMGS; ++im) {
      float s = acc[im];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) s_r[im * Cse + j] = swishf(s + __ldg(br + j));
    }
  }
  __syncthreads();
  const int per = (C + gridDim.y - 1) / gridDim.y;
  const int c_end = min(C, (int)(blockIdx.y + 1) * per);
  for (int c = blockIdx.y * per + tid; c < c_end; c += SE_THREADS) {
    float acc[SE_IMGS];
    const float bias = __ldg(be + c);
#pragma unroll
    for (int im = 0; im < SE_IMGS; ++im) acc[im] = bias;
    for (int j = 0; j < Cse; ++j) {
      const float wv = __ldg(we_t + (size_t)j * C + c);
#pragma unroll
      for (int im = 0; im < SE_IMGS; ++im) acc[im] = fmaf(wv, s_r[im * Cse + j], acc[im]);
    }
#pragma unroll
    for (int im = 0; im < SE_IMGS; ++im)
      if (im < n_img) gate[(size_t)(b0 + im) * C + c] = sigmoidf_(acc[im]);
  }
}

// ------------------------------------------------------------------- pool + FC + pose update
// feat [B][70][1536] (head output, swish applied) -> pose9 [B][9] -> TCO_out (optional).
// (reference: models/pose.py:83-86 and :69-79).
constexpr int HEAD_THREADS = 256;

__global__ void __launch_bounds__(HEAD_THREADS)
k_pool_fc_update(const float* __restrict__ feat, int n_pos, const float* __restrict__ fc_w,
                 const float* __restrict__ fc_b, float* __restrict__ pose9,
                 const float* __restrict__ TCO_in, const float* __restrict__ K_crop,
                 float* __restrict__ TCO_out) {
  __shared__ float s_pool[N_FEATURES];
  __shared__ float s_pose[POSE_DIM];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* fb = feat + (size_t)b * n_pos * N_FEATURES;
  const float inv = 1.0f / (float)n_pos;
  for (int c = tid; c < N_FEATURES; c += HEAD_THREADS) {
    float s = 0.f;
    for (int p = 0; p < n_pos; ++p) s += fb[(size_t)p * N_FEATURES + c];
    s_pool[c] = s * inv;
  }
  __syncthreads();
  const int warp = tid / 32, lane = tid % 32;
  for (int j = warp; j < POSE_DIM; j += HEAD_THREADS / 32) {
    const float* wrow = fc_w + (size_t)j * N_FEATURES;
    float s = 0.f;
    for (int c = lane; c < N_FEATURES; c += 32) s = fmaf(__ldg(wrow + c), s_pool[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      s += __ldg(fc_b + j);
      s_pose[j] = s;
      pose9[b * POSE_DIM + j] = s;
    }
  }
  __syncthreads();
  if (tid == 0 && TCO_out != nullptr) {
    float To[16];
    pose_update_one(TCO_in + b * 16, K_crop + b * 9, s_pose, To);
#pragma unroll
    for (int i = 0; i < 16; ++i) TCO_out[b * 16 + i] = To[i];
  }
}

}  // namespace cosyb
