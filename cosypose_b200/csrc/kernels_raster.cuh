// Device rasteriser for the "render" half of render-and-compare.
// (reference contract: models/pose.py:100-102 calls rendering/bullet_batch_renderer.py:46-90, one pybullet/OpenGL
//  getCameraImage per hypothesis in worker processes, uint8 frames through a multiprocessing queue, pinned copy,
//  /255; camera from K via simulator/camera.py:10-34 (near 0.01, far 10), background 0 (bullet_scene_renderer.py:49).)
//
// pybullet is not in this image, so the shading of the reference's renderer cannot be pinned ("parity unpinned"); what is
// restated here is the geometry of that camera: a pinhole with intrinsics K (skew honoured), sample points at pixel
// centres (j + 0.5, i + 0.5) as OpenGL does with the projection camera.py:10-34 builds, nearest surface wins, background 0,
// triangles with a vertex in front of the near plane (z < 0.01) are dropped instead of clipped.  Surface colour is the
// perspective-correct interpolation of per-vertex colours, unlit.
//
// Three launches per batch, no atomics on floats and no order dependence:
//   clear      depth/id buffer [B][240][320] of 64-bit keys = all ones, large-triangle counters = 0 (cudaMemsetAsync)
//   k_raster_tris   one lane per (hypothesis, triangle): camera transform, projection, bounding box, then by box size
//                   <= 32 pixels   the lane walks its own box: edge functions at the pixel centres, atomicMin of
//                                  (depth bits << 32 | triangle id) -- dense scans, pixel-sized triangles;
//                   <= 1024 pixels the warp takes these one at a time, 32 pixels per step (setup by shuffle);
//                   larger         the triangle id goes to the hypothesis's large-triangle list (coarse CAD meshes,
//                                  close-ups: a handful of triangles spanning the view would serialise a warp).
//   k_raster_resolve one thread per pixel, one CTA per 16x16 tile: first the large triangles whose box meets the tile
//                   (setups staged in shared memory, 256 at a time; the owner thread needs no atomic), then the
//                   winning triangle -> barycentric weights (the same arithmetic, so the same bits) -> colour ->
//                   uint8 NHWC [B][240][320][3], the layout k_stem<true> reads directly, or fp32 NCHW.
// Every floating-point step is an individually rounded IEEE operation (__fmul_rn / __fadd_rn / __fdiv_rn: no FMA
// contraction), edge functions are evaluated on a canonical ordering of their end points so the two triangles sharing an
// edge see bit-identical values with opposite sign (watertight: a pixel centre is never dropped between them); a plain
// float32 restatement on the host therefore reproduces the frames bit for bit (tests/test_gpu_render.py).
#pragma once
#include "common.h"

namespace cosyb {
namespace raster {

constexpr float NEAR_Z = 0.01f;   // simulator/camera.py:45
constexpr int RT_THREADS = 256;
constexpr int LANE_PIX = 32, WARP_PIX = 1024;   // box sizes up to which a lane / a warp rasterises a triangle
constexpr int TILE = 16;
constexpr unsigned long long EMPTY_KEY = ~0ull;

struct Tri {
  float u[3], v[3], iz[3];   // projected vertices (pixels) and 1 / camera z
  float area;                // twice the signed area after orientation (> 0)
  int x0, x1, y0, y1;        // inclusive pixel box, empty if x1 < x0 or y1 < y0
  int vid[3];                // vertex ids in the oriented order
  bool ok;
};

// (b - a) x (p - a), end points in canonical (lexicographic) order so that edge(a, b, p) == -edge(b, a, p) exactly
__device__ __forceinline__ float edge_fn(float ax, float ay, float bx, float by, float px, float py) {
  const bool swap = (bx < ax) || (bx == ax && by < ay);
  const float sx = swap ? bx : ax, sy = swap ? by : ay, ex = swap ? ax : bx, ey = swap ? ay : by;
  const float e = __fsub_rn(__fmul_rn(__fsub_rn(ex, sx), __fsub_rn(py, sy)), __fmul_rn(__fsub_rn(ey, sy), __fsub_rn(px, sx)));
  return swap ? -e : e;
}

// camera transform + projection + orientation + pixel box of one triangle; T = TCO [4][4] row-major, K [3][3]
__device__ __forceinline__ Tri setup_tri(const float* __restrict__ verts, const int32_t* __restrict__ face,
                                         const float* T, const float* K) {
  Tri t;
  t.ok = true;
  const float fx = K[0], sk = K[1], cx = K[2], fy = K[4], cy = K[5];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int vi = face[i];
    t.vid[i] = vi;
    const float X = __ldg(verts + 3 * (size_t)vi), Y = __ldg(verts + 3 * (size_t)vi + 1), Z = __ldg(verts + 3 * (size_t)vi + 2);
    const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], X), __fmul_rn(T[1], Y)), __fmul_rn(T[2], Z)), T[3]);
    const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], X), __fmul_rn(T[5], Y)), __fmul_rn(T[6], Z)), T[7]);
    const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], X), __fmul_rn(T[9], Y)), __fmul_rn(T[10], Z)), T[11]);
    if (!(zc >= NEAR_Z)) t.ok = false;   // also drops NaN poses (the reference renders a black frame for those)
    t.u[i] = __fadd_rn(__fdiv_rn(__fadd_rn(__fmul_rn(fx, xc), __fmul_rn(sk, yc)), zc), cx);
    t.v[i] = __fadd_rn(__fdiv_rn(__fmul_rn(fy, yc), zc), cy);
    t.iz[i] = __fdiv_rn(1.0f, zc);
  }
  float area = edge_fn(t.u[0], t.v[0], t.u[1], t.v[1], t.u[2], t.v[2]);
  if (area < 0.f) {   // orient counter-clockwise (both faces are drawn)
    float s;
    int si;
    s = t.u[1]; t.u[1] = t.u[2]; t.u[2] = s;
    s = t.v[1]; t.v[1] = t.v[2]; t.v[2] = s;
    s = t.iz[1]; t.iz[1] = t.iz[2]; t.iz[2] = s;
    si = t.vid[1]; t.vid[1] = t.vid[2]; t.vid[2] = si;
    area = -area;
  }
  t.area = area;
  if (!(area > 0.f)) t.ok = false;
  const float umin = fminf(t.u[0], fminf(t.u[1], t.u[2])), umax = fmaxf(t.u[0], fmaxf(t.u[1], t.u[2]));
  const float vmin = fminf(t.v[0], fminf(t.v[1], t.v[2])), vmax = fmaxf(t.v[0], fmaxf(t.v[1], t.v[2]));
  // pixel centres j + 0.5 inside [umin, umax]; the float -> int conversions saturate, so far-away boxes stay empty
  t.x0 = max(0, __float2int_ru(__fsub_rn(umin, 0.5f)));
  t.x1 = min(RENDER_W - 1, __float2int_rd(__fsub_rn(umax, 0.5f)));
  t.y0 = max(0, __float2int_ru(__fsub_rn(vmin, 0.5f)));
  t.y1 = min(RENDER_H - 1, __float2int_rd(__fsub_rn(vmax, 0.5f)));
  if (!(umin == umin && umax == umax && vmin == vmin && vmax == vmax)) t.ok = false;
  if (!t.ok) { t.x1 = -1; t.x0 = 0; t.y1 = -1; t.y0 = 0; }
  return t;
}

// barycentric weights of pixel (j, i); false if the centre is outside
__device__ __forceinline__ bool weights(const float* u, const float* v, float area, int j, int i, float* b) {
  const float px = __fadd_rn((float)j, 0.5f), py = __fadd_rn((float)i, 0.5f);
  const float w0 = edge_fn(u[1], v[1], u[2], v[2], px, py);
  const float w1 = edge_fn(u[2], v[2], u[0], v[0], px, py);
  const float w2 = edge_fn(u[0], v[0], u[1], v[1], px, py);
  if (!(w0 >= 0.f && w1 >= 0.f && w2 >= 0.f)) return false;
  b[0] = __fdiv_rn(w0, area);
  b[1] = __fdiv_rn(w1, area);
  b[2] = __fdiv_rn(w2, area);
  return true;
}

__device__ __forceinline__ float inv_depth(const float* b, const float* iz) {
  return __fadd_rn(__fadd_rn(__fmul_rn(b[0], iz[0]), __fmul_rn(b[1], iz[1])), __fmul_rn(b[2], iz[2]));
}

// depth key of the triangle at pixel (j, i), EMPTY_KEY if its centre is not covered
__device__ __forceinline__ unsigned long long pixel_key(const float* u, const float* v, const float* iz, float area,
                                                        int j, int i, unsigned int tri_id) {
  float b[3];
  if (!weights(u, v, area, j, i, b)) return EMPTY_KEY;
  const float q = inv_depth(b, iz);
  if (!(q > 0.f)) return EMPTY_KEY;
  const float z = __fdiv_rn(1.0f, q);
  return ((unsigned long long)__float_as_uint(z) << 32) | tri_id;
}

__device__ __forceinline__ void shade_pixel(const float* u, const float* v, const float* iz, float area, int j, int i,
                                            unsigned int tri_id, unsigned long long* __restrict__ zrow) {
  const unsigned long long key = pixel_key(u, v, iz, area, j, i, tri_id);
  if (key != EMPTY_KEY) atomicMin(zrow + (size_t)i * RENDER_W + j, key);
}

// grid = (ceil(max faces / 256), B); zbuf [B][240][320]; big_cnt [B], big_list [B][max_faces] (global face ids)
__global__ void __launch_bounds__(RT_THREADS)
k_raster_tris(const float* __restrict__ verts, const int32_t* __restrict__ faces, const int32_t* __restrict__ face_off,
              int n_labels, const int32_t* __restrict__ label_ids, const float* __restrict__ TCO,
              const float* __restrict__ Kc, unsigned long long* __restrict__ zbuf, int* __restrict__ big_cnt,
              int* __restrict__ big_list, int max_faces) {
  __shared__ float s_T[12], s_K[6];
  const int b = blockIdx.y;
  const int lab = min(max(__ldg(label_ids + b), 0), n_labels - 1);
  const int f0 = __ldg(face_off + lab), nf = __ldg(face_off + lab + 1) - f0;
  if ((int)(blockIdx.x * RT_THREADS) >= nf) return;
  if (threadIdx.x < 12) s_T[threadIdx.x] = __ldg(TCO + (size_t)b * 16 + threadIdx.x);
  if (threadIdx.x >= 32 && threadIdx.x < 38) s_K[threadIdx.x - 32] = __ldg(Kc + (size_t)b * 9 + (threadIdx.x - 32));
  __syncthreads();
  const int fl = blockIdx.x * RT_THREADS + threadIdx.x;
  unsigned long long* zrow = zbuf + (size_t)b * RENDER_H * RENDER_W;
  Tri t;
  t.x0 = t.y0 = 0; t.x1 = t.y1 = -1; t.area = 1.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) t.u[i] = t.v[i] = t.iz[i] = 0.f;
  if (fl < nf) t = setup_tri(verts, faces + 3 * (size_t)(f0 + fl), s_T, s_K);
  const int bw = t.x1 - t.x0 + 1, bh = t.y1 - t.y0 + 1;
  const int npix = (bw > 0 && bh > 0) ? bw * bh : 0;
  const unsigned int tri_id = (unsigned int)(f0 + fl);
  const bool big = npix > LANE_PIX && npix <= WARP_PIX;
  if (npix > 0 && npix <= LANE_PIX) {
    for (int i = t.y0; i <= t.y1; ++i)
      for (int j = t.x0; j <= t.x1; ++j) shade_pixel(t.u, t.v, t.iz, t.area, j, i, tri_id, zrow);
  }
  if (npix > WARP_PIX) big_list[(size_t)b * max_faces + atomicAdd(big_cnt + b, 1)] = (int)tri_id;   // any order: min wins
  // medium boxes: the whole warp takes them one at a time
  unsigned int todo = __ballot_sync(0xffffffffu, big);
  const int lane = threadIdx.x & 31;
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    float u[3], v[3], iz[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      u[i] = __shfl_sync(0xffffffffu, t.u[i], src);
      v[i] = __shfl_sync(0xffffffffu, t.v[i], src);
      iz[i] = __shfl_sync(0xffffffffu, t.iz[i], src);
    }
    const float area = __shfl_sync(0xffffffffu, t.area, src);
    const int x0 = __shfl_sync(0xffffffffu, t.x0, src), y0 = __shfl_sync(0xffffffffu, t.y0, src);
    const int w = __shfl_sync(0xffffffffu, bw, src), n = __shfl_sync(0xffffffffu, npix, src);
    const unsigned int id = __shfl_sync(0xffffffffu, tri_id, src);
    for (int p = lane; p < n; p += 32) {
      const int dy = p / w, dx = p - dy * w;
      shade_pixel(u, v, iz, area, x0 + dx, y0 + dy, id, zrow);
    }
  }
}

// one thread per pixel, one CTA per 16x16 tile: grid = (320 / 16, 240 / 16, B), 256 threads.
// out_u8 [B][240][320][3] or out_f [B][3][240][320] in [0, 1] (= uint8 / 255 as the reference does)
struct TriS { float u[3], v[3], iz[3], area; int x0, x1, y0, y1; unsigned int id; };
template <bool U8>
__global__ void __launch_bounds__(256)
k_raster_resolve(const float* __restrict__ verts, const float* __restrict__ colors, const int32_t* __restrict__ faces,
                 const float* __restrict__ TCO, const float* __restrict__ Kc,
                 const unsigned long long* __restrict__ zbuf, const int* __restrict__ big_cnt,
                 const int* __restrict__ big_list, int max_faces, void* __restrict__ out,
                 float* __restrict__ depth /*[B][240][320] camera z, 0 = background; may be null*/) {
  __shared__ TriS s_tri[256];
  __shared__ float s_T[12], s_K[6];
  const int b = blockIdx.z, tid = threadIdx.x;
  const int tx0 = blockIdx.x * TILE, ty0 = blockIdx.y * TILE;
  const int j = tx0 + (tid & (TILE - 1)), i = ty0 + (tid >> 4);
  const int pix = i * RENDER_W + j;
  const size_t idx = (size_t)b * RENDER_H * RENDER_W + pix;
  if (tid < 12) s_T[tid] = __ldg(TCO + (size_t)b * 16 + tid);
  if (tid >= 32 && tid < 38) s_K[tid - 32] = __ldg(Kc + (size_t)b * 9 + (tid - 32));
  unsigned long long key = zbuf[idx];
  const int n_big = min(__ldg(big_cnt + b), max_faces);
  __syncthreads();
  for (int base = 0; base < n_big; base += 256) {
    const int n = min(256, n_big - base);
    if (tid < n) {
      const int f = __ldg(big_list + (size_t)b * max_faces + base + tid);
      const Tri t = setup_tri(verts, faces + 3 * (size_t)f, s_T, s_K);
      TriS& d = s_tri[tid];
#pragma unroll
      for (int k = 0; k < 3; ++k) { d.u[k] = t.u[k]; d.v[k] = t.v[k]; d.iz[k] = t.iz[k]; }
      d.area = t.area; d.x0 = t.x0; d.x1 = t.x1; d.y0 = t.y0; d.y1 = t.y1; d.id = (unsigned int)f;
    }
    __syncthreads();
    for (int k = 0; k < n; ++k) {
      const TriS& d = s_tri[k];
      if (d.x1 < tx0 || d.x0 >= tx0 + TILE || d.y1 < ty0 || d.y0 >= ty0 + TILE) continue;   // uniform over the CTA
      if (j < d.x0 || j > d.x1 || i < d.y0 || i > d.y1) continue;
      key = min(key, pixel_key(d.u, d.v, d.iz, d.area, j, i, d.id));
    }
    __syncthreads();
  }
  float rgb[3] = {0.f, 0.f, 0.f};
  if (key != EMPTY_KEY) {
    const unsigned int tri_id = (unsigned int)(key & 0xffffffffu);
    const Tri t = setup_tri(verts, faces + 3 * (size_t)tri_id, s_T, s_K);
    float bw[3];
    if (t.ok && weights(t.u, t.v, t.area, j, i, bw)) {
      const float z = __uint_as_float((unsigned int)(key >> 32));
      float wq[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) wq[k] = __fmul_rn(bw[k], t.iz[k]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float c0 = __ldg(colors + 3 * (size_t)t.vid[0] + c), c1 = __ldg(colors + 3 * (size_t)t.vid[1] + c),
                    c2 = __ldg(colors + 3 * (size_t)t.vid[2] + c);
        const float sum = __fadd_rn(__fadd_rn(__fmul_rn(wq[0], c0), __fmul_rn(wq[1], c1)), __fmul_rn(wq[2], c2));
        rgb[c] = fminf(fmaxf(__fmul_rn(sum, z), 0.f), 1.f);
      }
    }
  }
  if (depth) depth[idx] = key != EMPTY_KEY ? __uint_as_float((unsigned int)(key >> 32)) : 0.f;
  unsigned char q[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) q[c] = (unsigned char)__float2int_rn(__fmul_rn(rgb[c], 255.f));
  if (U8) {
    unsigned char* o = (unsigned char*)out + idx * 3;
    o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
  } else {
    float* o = (float*)out + (size_t)b * 3 * RENDER_H * RENDER_W + pix;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[(size_t)c * RENDER_H * RENDER_W] = __fdiv_rn((float)q[c], 255.f);
  }
}

}  // namespace raster
}  // namespace cosyb
