// C-ABI of the engine (declarations and reference citations: include/cosyb200.h).
#include <cstdarg>
#include <cstring>
#include <map>
#include <string>

#include <dlfcn.h>

#include "common.h"
#include "effnet_table.h"
#include "kernels_backbone.cuh"
#include "kernels_crop.cuh"
#include "kernels_geometry.cuh"
#include "kernels_ransac.cuh"
#include "kernels_ba.cuh"
#include "kernels_eval.cuh"
#include "kernels_tc.cuh"
#include "kernels_pw2.cuh"
#include "kernels_xdw.cuh"
#include "kernels_dwtile.cuh"
#include "kernels_raster.cuh"

namespace cosyb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

int prof_resolve(cosyb200_handle* h) {
  if (h->ev_cat.empty()) return 0;
  CB_CUDA(cudaEventSynchronize(h->ev_pool[(h->ev_cat.size() - 1) * 2 + 1]));
  for (size_t i = 0; i < h->ev_cat.size(); ++i) {
    float ms = 0.f;
    CB_CUDA(cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]));
    h->cat_ms[h->ev_cat[i]] += ms;
    h->blk_ms[h->ev_cat[i]][h->ev_blk[i]] += ms;
  }
  h->ev_cat.clear();
  h->ev_blk.clear();
  return 0;
}

static int dev_alloc(void** p, size_t bytes) {
  CB_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  return 0;
}

static int upload(PoseModel& m, float** dst, const std::vector<float>& v) {
  void* p = nullptr;
  int rc = dev_alloc(&p, v.size() * sizeof(float));
  if (rc) return rc;
  m.allocs.push_back(p);
  CB_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  *dst = (float*)p;
  return 0;
}

// fp16 hi/lo image of a [N][K] weight for kernels_pw2.cuh
static int upload_p2(PoseModel& m, void** dst, float* inv_scale, const float* W_nk, int N, int K) {
  const float sc = pw2::weight_scale(W_nk, (size_t)N * K);
  const std::vector<uint16_t> img = pw2::pack_weights(W_nk, N, K, sc);
  void* p = nullptr;
  int rc = dev_alloc(&p, img.size() * 2);
  if (rc) return rc;
  m.allocs.push_back(p);
  CB_CUDA(cudaMemcpy(p, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  *dst = p;
  *inv_scale = 1.0f / sc;
  return 0;
}

// fp16 hi/lo chunk images of an expand weight (bias as an extra k row) for kernels_xdw.cuh
static int upload_xdw(PoseModel& m, BlockWeights& w, const float* W_nk, const std::vector<float>& bias, int N, int K, int cc) {
  const float sc = xdw::weight_scale(W_nk, bias.data(), N, K);
  const std::vector<uint16_t> img = xdw::pack_weights(W_nk, bias.data(), N, K, cc, sc);
  void* p = nullptr;
  int rc = dev_alloc(&p, img.size() * 2);
  if (rc) return rc;
  m.allocs.push_back(p);
  CB_CUDA(cudaMemcpy(p, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  w.expand_x = p;
  w.expand_x_inv = 1.0f / sc;
  return 0;
}

// ---- multi-GPU exchange: NCCL bound at run time ---------------------------------------------------
namespace {
struct NcclId { char b[128]; };   // ncclUniqueId (passed by value)
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
int nccl_bind() {
  if (g_nccl.lib) return 0;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy already loaded (PyTorch's), if any
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
  if (!lib) { set_error("NCCL is not available: %s", dlerror()); return COSYB200_ESTATE; }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(lib, "ncclAllGather");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather) {
    set_error("NCCL symbols missing in the loaded libnccl");
    return COSYB200_ESTATE;
  }
  g_nccl.lib = lib;
  return 0;
}
#define CB_NCCL(expr)                                                                             \
  do {                                                                                            \
    const int r_ = (expr);                                                                        \
    if (r_ != 0) {                                                                                \
      set_error("%s -> %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error"); \
      return COSYB200_ECUDA;                                                                      \
    }                                                                                             \
  } while (0)
}  // namespace


static void clear_graphs(cosyb200_handle* h) {
  for (auto& e : h->graphs) cudaGraphExecDestroy(e.exec);
  h->graphs.clear();
  h->model_epoch += 1;
}

static void free_model(PoseModel& m) {
  for (void* p : m.allocs) cudaFree(p);
  m = PoseModel();
}

// ---- GEMM dispatch --------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
static void launch_gemm_tile(bool gate, bool swish, bool resid, const float* A, const float* Wkn,
                             const float* bias, const float* g, const float* r, float* C, int M,
                             int N, int K, int rows_per_img, cudaStream_t st) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  if (!gate && swish && !resid)
    k_pw_gemm<BM, BN, TM, TN, false, true, false><<<grid, GEMM_THREADS, 0, st>>>(A, Wkn, bias, g, r, C, M, N, K, rows_per_img);
  else if (gate && !swish && !resid)
    k_pw_gemm<BM, BN, TM, TN, true, false, false><<<grid, GEMM_THREADS, 0, st>>>(A, Wkn, bias, g, r, C, M, N, K, rows_per_img);
  else if (gate && !swish && resid)
    k_pw_gemm<BM, BN, TM, TN, true, false, true><<<grid, GEMM_THREADS, 0, st>>>(A, Wkn, bias, g, r, C, M, N, K, rows_per_img);
}

static void launch_gemm(bool gate, bool swish, bool resid, const float* A, const float* Wkn,
                        const float* bias, const float* g, const float* r, float* C, int M, int N,
                        int K, int rows_per_img, cudaStream_t st) {
  // pick the N tile with the least padded work per unit of kernel efficiency
  const int bns[3] = {32, 64, 128};
  const float eff[3] = {0.6f, 0.8f, 1.0f};
  int best = 0;
  float best_cost = 1e30f;
  for (int i = 0; i < 3; ++i) {
    float cost = float((N + bns[i] - 1) / bns[i] * bns[i]) / eff[i];
    if (cost < best_cost - 1e-6f) { best_cost = cost; best = i; }
  }
  if (best == 0) launch_gemm_tile<128, 32, 4, 4>(gate, swish, resid, A, Wkn, bias, g, r, C, M, N, K, rows_per_img, st);
  else if (best == 1) launch_gemm_tile<128, 64, 8, 4>(gate, swish, resid, A, Wkn, bias, g, r, C, M, N, K, rows_per_img, st);
  else launch_gemm_tile<128, 128, 8, 8>(gate, swish, resid, A, Wkn, bias, g, r, C, M, N, K, rows_per_img, st);
}

// tensor-core path: Wpk = tc::pack_weights image of the [N][K] weight
// No more tiles than SMs: one bigger CTA per SM with two producer groups (measured at 140 tiles: 52 vs 65 us; at
// 210 tiles the 296 two-per-SM slots finish in one round and win, 73 vs 86 us).
static int tc_groups_for(int tiles) { return tiles <= 148 ? 2 : 1; }
template <int BN_MAX, bool G, bool S, bool R, int NG>
static int launch_gemm_tc_inst(cosyb200_handle* h, const float* A, const float* Wpk, const float* bias, const float* g,
                               const float* r, float* C, int M, int N, int K, int rows_per_img, cudaStream_t st) {
  const tc::Plan p = tc::make_plan(N, K, NG);
  const int m_tiles = (M + tc::BM - 1) / tc::BM;
  const int slots = NG == 1 ? 2 * h->n_sms : h->n_sms;
  const int grid = std::min(m_tiles, std::max(1, slots / p.n_tiles)) * p.n_tiles;   // multiple of n_tiles
  tc::k_pw_gemm_tc<BN_MAX, G, S, R, NG><<<grid, tc::threads_for(NG), p.smem_bytes, st>>>(
      A, Wpk, bias, g, r, C, M, N, K, rows_per_img, p.bn, p.n_tiles, p.nb, p.resident);
  CB_LAUNCH_CHECK();
  return 0;
}

static int launch_gemm_tc(cosyb200_handle* h, bool gate, bool swish, bool resid, const float* A, const float* Wpk,
                          const float* bias, const float* g, const float* r, float* C, int M, int N, int K,
                          int rows_per_img, cudaStream_t st) {
  const tc::Plan p1 = tc::make_plan(N, K, 1);
  const int tiles = ((M + tc::BM - 1) / tc::BM) * p1.n_tiles;
  const int ng = h->tc_groups ? h->tc_groups : tc_groups_for(tiles);
#define TC_ARGS h, A, Wpk, bias, g, r, C, M, N, K, rows_per_img, st
#define TC_DISPATCH(G, S, R) (ng == 2 ? launch_gemm_tc_inst<64, G, S, R, 2>(TC_ARGS) : launch_gemm_tc_inst<64, G, S, R, 1>(TC_ARGS))
  if (!gate && swish && !resid) return TC_DISPATCH(false, true, false);
  if (gate && !swish && !resid) return TC_DISPATCH(true, false, false);
  if (gate && !swish && resid) return TC_DISPATCH(true, false, true);
  if (!gate && !swish && !resid) return TC_DISPATCH(false, false, false);
#undef TC_DISPATCH
#undef TC_ARGS
  set_error("launch_gemm_tc: unsupported epilogue");
  return COSYB200_EINVAL;
}

// 3xFP16 tensor-core path (kernels_pw2.cuh): Wp2 = pw2::pack_weights image, inv_wscale = 1 / its power-of-two scale
static int launch_pw2(cosyb200_handle* h, bool gate, bool swish, bool resid, const float* A, const void* Wp2,
                      float inv_wscale, const float* bias, const float* g, const float* r, float* C, int M, int N,
                      int K, int rows_per_img, cudaStream_t st) {
  const pw2::Plan p = pw2::make_plan(M, N, K, h->n_sms, h->pw2_nt);
  if (p.bn == 0) { set_error("launch_pw2: no plan for M=%d N=%d K=%d", M, N, K); return COSYB200_EINVAL; }
  const int gate_smem = rows_per_img >= 64 ? 1 : 0;
  if (p.grid > h->n_sms) { set_error("launch_pw2: grid %d exceeds the SM count", p.grid); return COSYB200_EINVAL; }
#define PW2_LAUNCH_S(G, S, R, SM, SK)                                                                              \
  pw2::k_pw2<G, S, R, SM, SK><<<p.grid, pw2::THREADS, p.smem_bytes, st>>>(                                         \
      A, (const __half*)Wp2, bias, g, r, C, M, N, K, rows_per_img, p.bn, p.n_tiles, p.nb, p.resident,             \
      pw2::n_alloc_for(N), inv_wscale, gate_smem, h->pw_ws, h->pw_flags)
#define PW2_LAUNCH(G, S, R)                                                                                        \
  do {                                                                                                             \
    if (p.small) { if (p.split_k) PW2_LAUNCH_S(G, S, R, true, true); else PW2_LAUNCH_S(G, S, R, true, false); }    \
    else { if (p.split_k) PW2_LAUNCH_S(G, S, R, false, true); else PW2_LAUNCH_S(G, S, R, false, false); }          \
  } while (0)
  if (!gate && swish && !resid) PW2_LAUNCH(false, true, false);
  else if (gate && !swish && !resid) PW2_LAUNCH(true, false, false);
  else if (gate && !swish && resid) PW2_LAUNCH(true, false, true);
  else if (!gate && !swish && !resid) PW2_LAUNCH(false, false, false);
  else { set_error("launch_pw2: unsupported epilogue"); return COSYB200_EINVAL; }
#undef PW2_LAUNCH_S
#undef PW2_LAUNCH
  CB_LAUNCH_CHECK();
  return 0;
}

// One 1x1 convolution of the trunk through the implementation selected by the handle.
struct PwWeights { const float* kn; const float* tc; const void* p2; float p2_inv; const float* bias; };
static int pointwise(cosyb200_handle* h, bool gate, bool swish, bool resid, const float* A, const PwWeights& w,
                     const float* g, const float* r, float* C, int M, int N, int K, int rows_per_img, cudaStream_t st) {
  if (h->gemm_impl == 2) return launch_pw2(h, gate, swish, resid, A, w.p2, w.p2_inv, w.bias, g, r, C, M, N, K, rows_per_img, st);
  if (h->gemm_impl == 1) return launch_gemm_tc(h, gate, swish, resid, A, w.tc, w.bias, g, r, C, M, N, K, rows_per_img, st);
  launch_gemm(gate, swish, resid, A, w.kn, w.bias, g, r, C, M, N, K, rows_per_img, st);
  CB_LAUNCH_CHECK();
  return 0;
}

// Dynamic shared memory opt-ins are per device and per function: set for every kernel that needs more than
// 48 KB whenever a handle is created on a device (cosyb200_create, under its DeviceGuard).
template <typename F>
static int opt_in_smem(F* fn, int bytes) {
  CB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}
template <bool G, bool S, bool R>
static int opt_in_gemm_kernels() {
  int rc = 0;
  rc |= opt_in_smem(tc::k_pw_gemm_tc<64, G, S, R, 1>, 112 * 1024);
  rc |= opt_in_smem(tc::k_pw_gemm_tc<64, G, S, R, 2>, 202 * 1024);
  rc |= opt_in_smem(pw2::k_pw2<G, S, R, false, false>, 224 * 1024);
  rc |= opt_in_smem(pw2::k_pw2<G, S, R, true, false>, 224 * 1024);
  rc |= opt_in_smem(pw2::k_pw2<G, S, R, false, true>, 224 * 1024);
  rc |= opt_in_smem(pw2::k_pw2<G, S, R, true, true>, 224 * 1024);
  return rc;
}

// ---- depthwise dispatch ---------------------------------------------------------------------
struct DwPlan { int n_chunks, Gc, P, tiles; int V, TH, tiles_x, NX; };
static DwPlan dw_plan(const BlockSpec& b) {
  // k_dwconv_roll: thread = (channel vector, output column), rolls down TH output rows
  DwPlan p;
  p.V = b.k == 3 ? 4 : 2;
  int G = b.cexp / p.V;
  p.n_chunks = (G + DW_MAX_THREADS - 1) / DW_MAX_THREADS;
  while (G % p.n_chunks) ++p.n_chunks;
  p.Gc = G / p.n_chunks;
  p.P = DW_MAX_THREADS / p.Gc;
  p.TH = std::min(b.hout, 30);
  p.NX = 4;                                   // adjacent output columns per thread (register blocking)
  p.tiles_x = (b.wout + p.P * p.NX - 1) / (p.P * p.NX);
  p.tiles = p.tiles_x * ((b.hout + p.TH - 1) / p.TH);
  return p;
}

static int launch_dw(const BlockSpec& b, const BlockWeights& w, const float* in, float* out,
                     float* partial, int B, cudaStream_t st) {
  DwPlan p = dw_plan(b);
  dim3 grid(p.tiles, p.n_chunks, B);
  int threads = p.Gc * p.P;
#define DW_ARGS in, w.dw_w, w.dw_bias, out, partial, b.hin, b.win, b.cexp, b.hout, b.wout, b.pad_lo, p.Gc, p.P, p.TH, p.tiles_x, p.tiles
  if (b.k == 3 && b.s == 1) k_dwconv_roll<3, 1, 4, 4><<<grid, threads, 0, st>>>(DW_ARGS);
  else if (b.k == 3 && b.s == 2) k_dwconv_roll<3, 2, 4, 4><<<grid, threads, 0, st>>>(DW_ARGS);
  else if (b.k == 5 && b.s == 1) k_dwconv_roll<5, 1, 2, 4><<<grid, threads, 0, st>>>(DW_ARGS);
  else if (b.k == 5 && b.s == 2) k_dwconv_roll<5, 2, 2, 4><<<grid, threads, 0, st>>>(DW_ARGS);
  else { set_error("unsupported depthwise k=%d s=%d", b.k, b.s); return COSYB200_EINVAL; }
#undef DW_ARGS
  CB_LAUNCH_CHECK();
  return 0;
}

// smem-tiled depthwise (kernels_dwtile.cuh); the squeeze-excite gate is finished by k_se_fc2
template <int KS, int S, int WO, int XU>
static int launch_dw_tile_inst(const DwTilePlan& p, const BlockSpec& b, const BlockWeights& w, const float* in, float* out,
                               cosyb200_handle* h, int B, cudaStream_t st) {
  dim3 grid(p.n_chunks, p.n_strips * p.n_xt, B);
  k_dw_tile<KS, S, WO, XU><<<grid, DWT_THREADS, p.smem_bytes, st>>>(
      in, w.dw_w, w.dw_bias, out, h->pool_partial, b.hin, b.win, b.cexp, b.hout, b.wout, b.pad_lo, p.R, p.n_xt, b.cse,
      w.se_r_w);
  CB_LAUNCH_CHECK();
  return 0;
}

static int launch_dw_tile(const DwTilePlan& p, const BlockSpec& b, const BlockWeights& w, const float* in, float* out,
                          cosyb200_handle* h, int B, cudaStream_t st) {
#define DWT(KS, S, WO, XU) \
  if (b.k == KS && b.s == S && p.wo == WO && p.xu == XU) return launch_dw_tile_inst<KS, S, WO, XU>(p, b, w, in, out, h, B, st)
  DWT(5, 1, 40, 1); DWT(3, 2, 20, 1); DWT(3, 1, 20, 1); DWT(5, 1, 20, 1);
  DWT(5, 2, 10, 1); DWT(5, 1, 10, 1); DWT(3, 1, 10, 1);
#undef DWT
  set_error("launch_dw_tile: no instance for k=%d s=%d wout=%d", b.k, b.s, b.wout);
  return COSYB200_EINVAL;
}

static bool use_dw_tile(const cosyb200_handle* h, const BlockSpec& b) {
  return h->dw_impl != 0 && dw_tile_plan(b).ok;
}

// fused expand + depthwise + pooling (kernels_xdw.cuh)
static bool use_xdw(const cosyb200_handle* h, const BlockSpec& b) { return h->xdw != 0 && xdw::make_plan(b).ok; }

template <int KS, int S, int NX, int CCT, int RH>
static int launch_xdw_inst(const xdw::Plan& p, cosyb200_handle* h, const BlockSpec& b, const BlockWeights& w,
                           const float* x, float* out, int B, cudaStream_t st) {
  const int n_items = B * p.tiles_y * p.tiles_x;
  xdw::k_xdw<KS, S, NX, CCT, RH><<<std::min(n_items, h->n_sms), xdw::THREADS, p.smem_bytes, st>>>(
      x, (const __half*)w.expand_x, w.expand_x_inv, w.dw_w, w.dw_bias, out, h->pool_partial, B, b.hin, b.win, b.cin,
      b.cexp, b.hout, b.wout, b.pad_lo, p.MT, p.TH, p.TW, p.IH, p.IW, p.tiles_y, p.tiles_x, p.n_chunks, p.Kp,
      p.NYS, p.e_rows, h->cur_block == h->trace_block ? 1 : 0);
  CB_LAUNCH_CHECK();
  return 0;
}

static int launch_xdw(cosyb200_handle* h, const BlockSpec& b, const BlockWeights& w, const float* x, float* out, int B,
                      cudaStream_t st) {
  const xdw::Plan p = xdw::make_plan(b);
#define XDW(KS, S, NXV, CCV, RHV) if (b.k == KS && b.s == S && p.NX == NXV && p.cc == CCV && p.RH == RHV) return launch_xdw_inst<KS, S, NXV, CCV, RHV>(p, h, b, w, x, out, B, st)
  XDW(3, 2, 1, 48, 3); XDW(3, 1, 2, 64, 6); XDW(5, 2, 1, 64, 3); XDW(5, 1, 2, 48, 5); XDW(3, 2, 2, 48, 2);
#undef XDW
  set_error("launch_xdw: no instance for k=%d s=%d NX=%d", b.k, b.s, p.NX);
  return COSYB200_EINVAL;
}

// ---- trunk forward --------------------------------------------------------------------------
static int net_forward(cosyb200_handle* h, int slot, int B, const float* crops, const void* renders,
                       int render_u8, float* pose9, float* const* taps, const float* TCO_in, const float* K_crop,
                       float* TCO_out, cudaStream_t st) {
  const PoseModel& m = h->models[slot];
  {
    dim3 grid(RENDER_W / 2 / STEM_TX, RENDER_H / 2 / STEM_TY, B), block(STEM_TX, STEM_TY);
    LaunchScope ls(h, CAT_STEM, st);
    if (render_u8) k_stem<true><<<grid, block, 0, st>>>(crops, renders, m.stem_host, h->act[0]);
    else k_stem<false><<<grid, block, 0, st>>>(crops, renders, m.stem_host, h->act[0]);
    CB_LAUNCH_CHECK();
  }
  int cur = 0;
  auto tap = [&](int i, const float* src, size_t n) -> int {
    if (taps && taps[i]) CB_CUDA(cudaMemcpyAsync(taps[i], src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
  };
  int rc = tap(0, h->act[0], (size_t)B * (RENDER_H / 2) * (RENDER_W / 2) * STEM_OUT);
  if (rc) return rc;
  for (size_t i = 0; i < h->blocks.size(); ++i) {
    const BlockSpec& b = h->blocks[i];
    const BlockWeights& w = m.blocks[i];
    h->cur_block = (int)i;
    const float* x = h->act[cur];
    float* y = h->act[cur ^ 1];
    const int Min = B * b.hin * b.win, Mout = B * b.hout * b.wout;
    const float* dw_in = x;
    const bool fused = use_xdw(h, b);
    if (fused) {
      LaunchScope ls(h, CAT_DW, st);
      if (int rc2 = launch_xdw(h, b, w, x, h->buf_d, B, st)) return rc2;
    }
    if (b.e != 1 && !fused) {
      LaunchScope ls(h, CAT_EXPAND, st);
      const PwWeights pw{w.expand_kn, w.expand_tc, w.expand_p2, w.expand_p2_inv, w.expand_bias};
      if (int rc2 = pointwise(h, false, true, false, x, pw, nullptr, nullptr, h->buf_e, Min, b.cexp, b.cin, 1, st)) return rc2;
      dw_in = h->buf_e;
      if ((int)i == h->dump_block && h->dump_e)
        CB_CUDA(cudaMemcpyAsync(h->dump_e, h->buf_e, (size_t)Min * b.cexp * 4, cudaMemcpyDeviceToDevice, st));
    }
    const bool tile_dw = !fused && use_dw_tile(h, b);
    if (!fused) {
      LaunchScope ls(h, CAT_DW, st);
      if (tile_dw) rc = launch_dw_tile(dw_tile_plan(b), b, w, dw_in, h->buf_d, h, B, st);
      else rc = launch_dw(b, w, dw_in, h->buf_d, h->pool_partial, B, st);
    }
    if (rc) return rc;
    DwPlan p = dw_plan(b);
    if (fused) {
      const xdw::Plan xp = xdw::make_plan(b);
      p.tiles = xp.tiles_y * xp.tiles_x;
    }
    if (tile_dw) {
      const DwTilePlan tp = dw_tile_plan(b);
      LaunchScope ls(h, CAT_SE, st);
      const dim3 g2((b.cexp + SE2_THREADS - 1) / SE2_THREADS, B);
      k_se_fc2<<<g2, SE2_THREADS, 0, st>>>(h->pool_partial, tp.n_strips * tp.n_xt * tp.n_chunks, b.cexp, b.cse,
                                           1.0f / float(b.hout * b.wout), w.se_r_b, w.se_e_w, w.se_e_b, h->gate);
    } else {
      LaunchScope ls(h, CAT_SE, st);
      if (b.cexp <= 288) {
        const dim3 se_grid((B + SE_IMGS_SMALL - 1) / SE_IMGS_SMALL, 1);
        const size_t se_smem = (size_t)SE_IMGS_SMALL * (b.cexp + b.cse) * sizeof(float);
        k_se_gate<SE_IMGS_SMALL><<<se_grid, SE_THREADS, se_smem, st>>>(B, h->pool_partial, p.tiles, b.cexp, b.cse,
                                                                     1.0f / float(b.hout * b.wout), w.se_r_w, w.se_r_b,
                                                                     w.se_e_w, w.se_e_b, h->gate);
      } else {
        const dim3 se_grid((B + SE_IMGS - 1) / SE_IMGS, b.cexp >= 1024 ? 8 : 4);
        const size_t se_smem = (size_t)SE_IMGS * (b.cexp + b.cse) * sizeof(float);
        k_se_gate<SE_IMGS><<<se_grid, SE_THREADS, se_smem, st>>>(B, h->pool_partial, p.tiles, b.cexp, b.cse,
                                                               1.0f / float(b.hout * b.wout), w.se_r_w, w.se_r_b, w.se_e_w,
                                                               w.se_e_b, h->gate);
      }
    }
    CB_LAUNCH_CHECK();
    if ((int)i == h->dump_block) {
      if (h->dump_d) CB_CUDA(cudaMemcpyAsync(h->dump_d, h->buf_d, (size_t)Mout * b.cexp * 4, cudaMemcpyDeviceToDevice, st));
      if (h->dump_gate) CB_CUDA(cudaMemcpyAsync(h->dump_gate, h->gate, (size_t)B * b.cexp * 4, cudaMemcpyDeviceToDevice, st));
    }
    {
      LaunchScope ls(h, CAT_PROJECT, st);
      const PwWeights pw{w.proj_kn, w.proj_tc, w.proj_p2, w.proj_p2_inv, w.proj_bias};
      if (int rc2 = pointwise(h, true, false, b.skip != 0, h->buf_d, pw, h->gate, x, y, Mout, b.cout, b.cexp,
                              b.hout * b.wout, st)) return rc2;
    }
    cur ^= 1;
    rc = tap(1 + (int)i, y, (size_t)Mout * b.cout);
    if (rc) return rc;
  }
  h->cur_block = cosyb200_handle::N_BLK - 1;
  const BlockSpec& last = h->blocks.back();
  const int n_pos = last.hout * last.wout;
  {
    LaunchScope ls(h, CAT_HEAD, st);
    const PwWeights pw{m.head_kn, m.head_tc, m.head_p2, m.head_p2_inv, m.head_bias};
    if (int rc2 = pointwise(h, false, true, false, h->act[cur], pw, nullptr, nullptr, h->buf_e, B * n_pos, N_FEATURES,
                            last.cout, 1, st)) return rc2;
  }
  rc = tap(1 + (int)h->blocks.size(), h->buf_e, (size_t)B * n_pos * N_FEATURES);
  if (rc) return rc;
  {
    LaunchScope ls(h, CAT_POOL_FC, st);
    k_pool_fc_update<<<B, HEAD_THREADS, 0, st>>>(h->buf_e, n_pos, m.fc_w, m.fc_b, pose9, TCO_in, K_crop, TCO_out);
  }
  CB_LAUNCH_CHECK();
  return 0;
}

static int check_batch(cosyb200_handle* h, int B, const char* fn) {
  CB_CHECK_ARG(h != nullptr, "%s: null handle", fn);
  CB_CHECK_ARG(B >= 1 && B <= h->max_batch, "%s: batch %d outside [1, %d]", fn, B, h->max_batch);
  return 0;
}

}  // namespace cosyb

using namespace cosyb;

extern "C" {

const char* cosyb200_last_error(void) { return g_err; }
int cosyb200_version(void) { return 100; }

int cosyb200_effnet_block(int idx, int32_t* out) {
  static const std::vector<BlockSpec> blocks = make_effnet_b3();
  CB_CHECK_ARG(out != nullptr && idx >= 0 && idx < (int)blocks.size(), "effnet_block: bad index %d", idx);
  const BlockSpec& b = blocks[idx];
  int32_t v[11] = {b.k, b.s, b.e, b.cin, b.cexp, b.cse, b.cout, b.pad_lo, b.pad_hi, b.skip, (int32_t)blocks.size()};
  memcpy(out, v, sizeof(v));
  return COSYB200_OK;
}

int cosyb200_launch_plan(int idx, int batch, int32_t* out) {
  static const std::vector<BlockSpec> blocks = make_effnet_b3();
  CB_CHECK_ARG(out != nullptr && idx >= 0 && idx < (int)blocks.size() && batch >= 1, "launch_plan: bad arguments");
  const BlockSpec& b = blocks[idx];
  const DwPlan rp = dw_plan(b);
  const DwTilePlan tp = dw_tile_plan(b);
  const int Min = batch * b.hin * b.win, Mout = batch * b.hout * b.wout;
  const tc::Plan pe = tc::make_plan(b.cexp, b.cin, 1), pp = tc::make_plan(b.cout, b.cexp, 1);
  const int tiles_e = ((Min + tc::BM - 1) / tc::BM) * pe.n_tiles, tiles_p = ((Mout + tc::BM - 1) / tc::BM) * pp.n_tiles;
  const int ng_e = tc_groups_for(tiles_e), ng_p = tc_groups_for(tiles_p);
  const tc::Plan pe2 = tc::make_plan(b.cexp, b.cin, ng_e), pp2 = tc::make_plan(b.cout, b.cexp, ng_p);
  int32_t v[32] = {
      // depthwise: 0 tiled?, 1 rows per tile, 2 row strips, 3 x tiles, 4 unit width, 5 units per row, 6 channel chunks,
      // 7 shared memory bytes; rolling kernel: 8 tiles, 9 channel chunks, 10 threads, 11 rows per tile
      tp.ok ? 1 : 0, tp.R, tp.n_strips, tp.n_xt, tp.wo, tp.xu, tp.n_chunks, tp.smem_bytes,
      rp.tiles, rp.n_chunks, rp.Gc * rp.P, rp.TH,
      // expand 1x1: 12 bn, 13 n tiles, 14 k stages, 15 weight slots, 16 resident, 17 shared memory bytes, 18 producer
      // groups, 19 tiles
      pe2.bn, pe2.n_tiles, pe2.nk, pe2.nb, pe2.resident, pe2.smem_bytes, ng_e, tiles_e,
      // project 1x1: 20..27 the same
      pp2.bn, pp2.n_tiles, pp2.nk, pp2.nb, pp2.resident, pp2.smem_bytes, ng_p, tiles_p,
      0, 0, 0, 0};
  memcpy(out, v, sizeof(v));
  return COSYB200_OK;
}

int cosyb200_pw2_plan(int M, int N, int K, int n_sms, int32_t* out) {
  CB_CHECK_ARG(out != nullptr && M >= 1 && N >= 1 && K >= 1 && n_sms >= 1, "pw2_plan: bad arguments");
  const pw2::Plan p = pw2::make_plan(M, N, K, n_sms);
  const int32_t v[9] = {p.bn, p.small, p.n_tiles, p.nk, p.nb, p.resident, p.smem_bytes, p.grid, p.split_k};
  memcpy(out, v, sizeof(v));
  return p.bn ? COSYB200_OK : COSYB200_EINVAL;
}

int cosyb200_create(cosyb200_handle** out, int device, int max_batch) {
  CB_CHECK_ARG(out != nullptr, "create: null out");
  CB_CHECK_ARG(max_batch >= 1 && max_batch <= 4096, "create: max_batch %d outside [1, 4096]", max_batch);
  int n_dev = 0;
  CB_CUDA(cudaGetDeviceCount(&n_dev));
  CB_CHECK_ARG(device >= 0 && device < n_dev, "create: device %d not present (%d devices)", device, n_dev);
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  CB_CUDA(cudaGetDeviceProperties(&prop, device));
  CB_CHECK_ARG(prop.major == 10, "create: device %d is sm_%d%d; this engine is built for sm_100a only",
               device, prop.major, prop.minor);
  {
    // per-device, per-function opt-ins (idempotent; repeated for every handle so that each device gets them)
    int rc = 0;
    rc |= opt_in_gemm_kernels<false, true, false>();
    rc |= opt_in_gemm_kernels<true, false, false>();
    rc |= opt_in_gemm_kernels<true, false, true>();
    rc |= opt_in_gemm_kernels<false, false, false>();
    rc |= opt_in_smem(xdw::k_xdw<3, 2, 1, 48, 3>, 208 * 1024);
    rc |= opt_in_smem(xdw::k_xdw<3, 1, 2, 64, 6>, 208 * 1024);
    rc |= opt_in_smem(xdw::k_xdw<5, 2, 1, 64, 3>, 208 * 1024);
    rc |= opt_in_smem(xdw::k_xdw<5, 1, 2, 48, 5>, 208 * 1024);
    rc |= opt_in_smem(xdw::k_xdw<3, 2, 2, 48, 2>, 208 * 1024);
    rc |= opt_in_smem(k_lm_solve<true>, LM_SMEM_MAX_N * (LM_SMEM_MAX_N + 1) / 2 * 8);
    rc |= opt_in_smem(k_pose_errors, 16000 * 12);
    rc |= opt_in_smem(k_roi_crop, (int)(CROP_SMEM_FLOATS * sizeof(float)));
    rc |= opt_in_smem(k_se_gate<SE_IMGS>, 100 * 1024);
    rc |= opt_in_smem(k_dw_tile<5, 1, 40, 1>, 80 * 1024);
    rc |= opt_in_smem(k_dw_tile<3, 2, 20, 1>, 80 * 1024);
    rc |= opt_in_smem(k_dw_tile<3, 1, 20, 1>, 80 * 1024);
    rc |= opt_in_smem(k_dw_tile<5, 1, 20, 1>, 80 * 1024);
    rc |= opt_in_smem(k_dw_tile<5, 2, 10, 1>, 80 * 1024);
    rc |= opt_in_smem(k_dw_tile<5, 1, 10, 1>, 80 * 1024);
    rc |= opt_in_smem(k_dw_tile<3, 1, 10, 1>, 80 * 1024);
    if (rc) return COSYB200_ECUDA;
  }
  cosyb200_handle* h = new cosyb200_handle();
  h->device = device;
  h->max_batch = max_batch;
  h->n_sms = prop.multiProcessorCount;
  h->blocks = make_effnet_b3();
  size_t act = (size_t)(RENDER_H / 2) * (RENDER_W / 2) * STEM_OUT, e = 0, d = 0, part = 0, cmax = 0;
  for (const BlockSpec& b : h->blocks) {
    act = std::max(act, (size_t)b.hout * b.wout * b.cout);
    if (b.e != 1) e = std::max(e, (size_t)b.hin * b.win * b.cexp);
    d = std::max(d, (size_t)b.hout * b.wout * b.cexp);
    part = std::max(part, (size_t)dw_plan(b).tiles * b.cexp);
    if (xdw::make_plan(b).ok) part = std::max(part, (size_t)xdw::make_plan(b).tiles_y * xdw::make_plan(b).tiles_x * b.cexp);
    if (dw_tile_plan(b).ok)
      part = std::max(part, (size_t)dw_tile_plan(b).n_strips * dw_tile_plan(b).n_xt * dw_tile_plan(b).n_chunks * b.cse);
    cmax = std::max(cmax, (size_t)b.cexp);
  }
  e = std::max(e, (size_t)h->blocks.back().hout * h->blocks.back().wout * N_FEATURES);
  h->act_elems = act; h->e_elems = e; h->d_elems = d; h->partial_elems = part;
  const size_t B = (size_t)max_batch;
  int rc = 0;
  rc |= dev_alloc((void**)&h->act[0], B * act * 4);
  rc |= dev_alloc((void**)&h->act[1], B * act * 4);
  rc |= dev_alloc((void**)&h->buf_e, B * e * 4);
  rc |= dev_alloc((void**)&h->buf_d, B * d * 4);
  rc |= dev_alloc((void**)&h->pool_partial, B * part * 4);
  rc |= dev_alloc((void**)&h->gate, B * cmax * 4);
  rc |= dev_alloc((void**)&h->crops, B * 3 * RENDER_H * RENDER_W * 4);
  {   // k_pw2: partial-sum slots + flags of the k-stage work split, one per CTA (flags start at 0, owners reset them)
    void* p0 = nullptr;
    rc |= dev_alloc(&p0, (size_t)h->n_sms * pw2::WS_SLOT_BYTES + (size_t)h->n_sms * 4);
    if (p0) {
      h->pw_ws = (float*)p0;
      h->pw_flags = (int*)((char*)p0 + (size_t)h->n_sms * pw2::WS_SLOT_BYTES);
      if (cudaMemset(h->pw_flags, 0, (size_t)h->n_sms * 4) != cudaSuccess) rc |= 1;
    }
  }
  rc |= dev_alloc((void**)&h->pose9, B * POSE_DIM * 4);
  if (rc) { cosyb200_destroy(h); return COSYB200_ENOMEM; }
  *out = h;
  return COSYB200_OK;
}

int cosyb200_destroy(cosyb200_handle* h) {
  if (!h) return COSYB200_OK;
  DeviceGuard guard(h->device);
  clear_graphs(h);
  if (h->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->nccl_comm);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  free_model(h->models[0]);
  free_model(h->models[1]);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  void* ptrs[] = {h->act[0], h->act[1], h->buf_e, h->buf_d, h->pool_partial, h->gate, h->crops, h->pose9,
                  h->pts_sampled, h->sym, h->n_sym, h->aabb, h->io_buf, h->ba_ws, h->lm_ws, h->vote_ws,
                  h->r_verts, h->r_colors, h->r_faces, h->r_face_off, h->r_zbuf, h->r_frames, h->r_big_cnt, h->r_big_list,
                  h->pw_ws};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete h;
  return COSYB200_OK;
}

int cosyb200_load_pose_model(cosyb200_handle* h, int slot, int n, const char* const* names,
                             const float* const* ptrs, const int64_t* numels) {
  CB_CHECK_ARG(h != nullptr, "load_pose_model: null handle");
  CB_CHECK_ARG(slot == 0 || slot == 1, "load_pose_model: slot %d", slot);
  CB_CHECK_ARG(n > 0 && names && ptrs && numels, "load_pose_model: empty state dict");
  DeviceGuard guard(h->device);
  clear_graphs(h);
  std::map<std::string, std::pair<const float*, int64_t>> sd;
  for (int i = 0; i < n; ++i) sd[names[i]] = {ptrs[i], numels[i]};
  bool ok = true;
  std::string missing;
  auto get = [&](const std::string& key, int64_t numel) -> const float* {
    auto it = sd.find(key);
    if (it == sd.end() || it->second.second != numel || it->second.first == nullptr) {
      if (ok) missing = key + (it == sd.end() ? " (missing)" : " (wrong size)");
      ok = false;
      return nullptr;
    }
    return it->second.first;
  };
  // BN eval: y = x * g/sqrt(var+eps) + (b - mean * g/sqrt(var+eps))
  auto bn_fold = [&](const std::string& prefix, int c, std::vector<float>& scale, std::vector<float>& shift) {
    const float* g = get(prefix + ".weight", c);
    const float* b = get(prefix + ".bias", c);
    const float* mu = get(prefix + ".running_mean", c);
    const float* var = get(prefix + ".running_var", c);
    scale.assign(c, 1.f);
    shift.assign(c, 0.f);
    if (!g || !b || !mu || !var) return;
    for (int i = 0; i < c; ++i) {
      float inv = 1.0f / sqrtf(var[i] + BN_EPS);
      float a = g[i] * inv;
      scale[i] = a;
      shift[i] = b[i] - mu[i] * a;
    }
  };
  // [N][K] row-scaled + its transpose
  auto pack_pw = [&](const float* W, int N, int K, const std::vector<float>& scale, std::vector<float>& nk,
                     std::vector<float>& kn) {
    nk.assign((size_t)N * K, 0.f);
    kn.assign((size_t)N * K, 0.f);
    if (!W) return;
    for (int o = 0; o < N; ++o)
      for (int k = 0; k < K; ++k) {
        float v = W[(size_t)o * K + k] * scale[o];
        nk[(size_t)o * K + k] = v;
        kn[(size_t)k * N + o] = v;
      }
  };

  PoseModel m;
  m.blocks.resize(h->blocks.size());
  std::vector<float> scale, shift, nk, kn, tmp;
  int rc = 0;
  {
    const float* W = get("backbone._conv_stem.weight", (int64_t)STEM_OUT * IN_CH * 9);
    bn_fold("backbone._bn0", STEM_OUT, scale, shift);
    tmp.assign(54 * STEM_OUT, 0.f);
    if (W)
      for (int co = 0; co < STEM_OUT; ++co)
        for (int ci = 0; ci < IN_CH; ++ci)
          for (int t = 0; t < 9; ++t)
            tmp[(t * IN_CH + ci) * STEM_OUT + co] = W[((size_t)co * IN_CH + ci) * 9 + t] * scale[co];
    rc |= upload(m, &m.stem_w, tmp);
    rc |= upload(m, &m.stem_bias, shift);
    if (tmp.size() == 54 * STEM_OUT && shift.size() >= (size_t)STEM_OUT) {
      memcpy(m.stem_host.w, tmp.data(), sizeof(m.stem_host.w));
      memcpy(m.stem_host.bias, shift.data(), sizeof(m.stem_host.bias));
    }
  }
  for (size_t i = 0; i < h->blocks.size() && !rc; ++i) {
    const BlockSpec& b = h->blocks[i];
    BlockWeights& w = m.blocks[i];
    const std::string p = "backbone._blocks." + std::to_string(i);
    if (b.e != 1) {
      const float* W = get(p + "._expand_conv.weight", (int64_t)b.cexp * b.cin);
      bn_fold(p + "._bn0", b.cexp, scale, shift);
      pack_pw(W, b.cexp, b.cin, scale, nk, kn);
      rc |= upload(m, &w.expand_nk, nk);
      rc |= upload(m, &w.expand_kn, kn);
      rc |= upload(m, &w.expand_tc, tc::pack_weights(nk.data(), b.cexp, b.cin));
      rc |= upload_p2(m, &w.expand_p2, &w.expand_p2_inv, nk.data(), b.cexp, b.cin);
      rc |= upload(m, &w.expand_bias, shift);
      if (xdw::make_plan(b).ok) rc |= upload_xdw(m, w, nk.data(), shift, b.cexp, b.cin, xdw::make_plan(b).cc);
    }
    {
      const int kk = b.k * b.k;
      const float* W = get(p + "._depthwise_conv.weight", (int64_t)b.cexp * kk);
      bn_fold(p + "._bn1", b.cexp, scale, shift);
      tmp.assign((size_t)kk * b.cexp, 0.f);
      if (W)
        for (int c = 0; c < b.cexp; ++c)
          for (int t = 0; t < kk; ++t) tmp[(size_t)t * b.cexp + c] = W[(size_t)c * kk + t] * scale[c];
      rc |= upload(m, &w.dw_w, tmp);
      rc |= upload(m, &w.dw_bias, shift);
    }
    {
      auto copy_up = [&](const std::string& key, int64_t numel, float** dst) {
        const float* src = get(key, numel);
        tmp.assign((size_t)numel, 0.f);
        if (src) memcpy(tmp.data(), src, (size_t)numel * 4);
        rc |= upload(m, dst, tmp);
      };
      copy_up(p + "._se_reduce.weight", (int64_t)b.cse * b.cexp, &w.se_r_w);
      copy_up(p + "._se_reduce.bias", b.cse, &w.se_r_b);
      {   // expand weight transposed to [cse][cexp]
        const float* src = get(p + "._se_expand.weight", (int64_t)b.cexp * b.cse);
        tmp.assign((size_t)b.cexp * b.cse, 0.f);
        if (src)
          for (int c = 0; c < b.cexp; ++c)
            for (int j = 0; j < b.cse; ++j) tmp[(size_t)j * b.cexp + c] = src[(size_t)c * b.cse + j];
        rc |= upload(m, &w.se_e_w, tmp);
      }
      copy_up(p + "._se_expand.bias", b.cexp, &w.se_e_b);
    }
    {
      const float* W = get(p + "._project_conv.weight", (int64_t)b.cout * b.cexp);
      bn_fold(p + "._bn2", b.cout, scale, shift);
      pack_pw(W, b.cout, b.cexp, scale, nk, kn);
      rc |= upload(m, &w.proj_nk, nk);
      rc |= upload(m, &w.proj_kn, kn);
      rc |= upload(m, &w.proj_tc, tc::pack_weights(nk.data(), b.cout, b.cexp));
      rc |= upload_p2(m, &w.proj_p2, &w.proj_p2_inv, nk.data(), b.cout, b.cexp);
      rc |= upload(m, &w.proj_bias, shift);
    }
  }
  if (!rc) {
    const int cin = h->blocks.back().cout;
    const float* W = get("backbone._conv_head.weight", (int64_t)N_FEATURES * cin);
    bn_fold("backbone._bn1", N_FEATURES, scale, shift);
    pack_pw(W, N_FEATURES, cin, scale, nk, kn);
    rc |= upload(m, &m.head_nk, nk);
    rc |= upload(m, &m.head_kn, kn);
    rc |= upload(m, &m.head_tc, tc::pack_weights(nk.data(), N_FEATURES, cin));
    rc |= upload_p2(m, &m.head_p2, &m.head_p2_inv, nk.data(), N_FEATURES, cin);
    rc |= upload(m, &m.head_bias, shift);
    const float* fw = get("pose_fc.weight", (int64_t)POSE_DIM * N_FEATURES);
    const float* fb = get("pose_fc.bias", POSE_DIM);
    tmp.assign((size_t)POSE_DIM * N_FEATURES, 0.f);
    if (fw) memcpy(tmp.data(), fw, tmp.size() * 4);
    rc |= upload(m, &m.fc_w, tmp);
    tmp.assign(POSE_DIM, 0.f);
    if (fb) memcpy(tmp.data(), fb, POSE_DIM * 4);
    rc |= upload(m, &m.fc_b, tmp);
  }
  if (rc) { free_model(m); return rc; }
  if (!ok) {
    free_model(m);
    set_error("load_pose_model: state_dict entry %s", missing.c_str());
    return COSYB200_EINVAL;
  }
  free_model(h->models[slot]);
  m.loaded = true;
  h->models[slot] = m;
  return COSYB200_OK;
}

int cosyb200_set_meshes(cosyb200_handle* h, int n_labels, int n_points, const float* points,
                        int n_sample, const int64_t* point_ids, int s_max, const float* sym,
                        const int32_t* n_sym, const float* aabb) {
  CB_CHECK_ARG(h != nullptr, "set_meshes: null handle");
  CB_CHECK_ARG(n_labels >= 1 && s_max >= 1 && sym && n_sym && aabb, "set_meshes: bad tables");
  CB_CHECK_ARG((points == nullptr) || (n_sample == N_SAMPLE && point_ids && n_points >= n_sample),
               "set_meshes: need %d sampled point ids out of >= %d points", N_SAMPLE, N_SAMPLE);
  DeviceGuard guard(h->device);
  clear_graphs(h);
  // points == NULL replaces the symmetry / AABB tables only: an engine shared with the pose predictors keeps
  // its sampled points (they belong to the same label set)
  CB_CHECK_ARG(points != nullptr || h->pts_sampled == nullptr || n_labels == h->n_labels,
               "set_meshes: %d labels without points on an engine holding points of %d labels", n_labels, h->n_labels);
  if (points && h->pts_sampled) { cudaFree(h->pts_sampled); h->pts_sampled = nullptr; }
  for (void* p : {(void*)h->sym, (void*)h->n_sym, (void*)h->aabb}) if (p) cudaFree(p);
  h->sym = nullptr; h->n_sym = nullptr; h->aabb = nullptr;
  h->n_labels = n_labels;
  h->s_max = s_max;
  for (int l = 0; l < n_labels; ++l)
    CB_CHECK_ARG(n_sym[l] >= 1 && n_sym[l] <= s_max, "set_meshes: n_sym[%d]=%d outside [1,%d]", l, n_sym[l], s_max);
  if (points) {
    std::vector<float> ps((size_t)n_labels * N_SAMPLE * 3);
    for (int l = 0; l < n_labels; ++l)
      for (int i = 0; i < N_SAMPLE; ++i) {
        int64_t id = point_ids[i];
        CB_CHECK_ARG(id >= 0 && id < n_points, "set_meshes: point id %lld out of range", (long long)id);
        memcpy(&ps[((size_t)l * N_SAMPLE + i) * 3], &points[((size_t)l * n_points + id) * 3], 12);
      }
    CB_CUDA(cudaMalloc((void**)&h->pts_sampled, ps.size() * 4));
    CB_CUDA(cudaMemcpy(h->pts_sampled, ps.data(), ps.size() * 4, cudaMemcpyHostToDevice));
  }
  CB_CUDA(cudaMalloc((void**)&h->sym, (size_t)n_labels * s_max * 64));
  CB_CUDA(cudaMemcpy(h->sym, sym, (size_t)n_labels * s_max * 64, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMalloc((void**)&h->n_sym, (size_t)n_labels * 4));
  CB_CUDA(cudaMemcpy(h->n_sym, n_sym, (size_t)n_labels * 4, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMalloc((void**)&h->aabb, (size_t)n_labels * 96));
  CB_CUDA(cudaMemcpy(h->aabb, aabb, (size_t)n_labels * 96, cudaMemcpyHostToDevice));
  return COSYB200_OK;
}

#define NEED_MESH_PTS(fn) CB_CHECK_ARG(h->pts_sampled != nullptr, fn ": meshes not set");

int cosyb200_tco_init(cosyb200_handle* h, int B, int zup, const float* boxes, const float* K,
                      const int32_t* label_ids, float* TCO, void* stream) {
  CB_CHECK_ARG(h != nullptr && B >= 1, "tco_init: bad arguments");
  if (zup) { if (!h->pts_sampled) { set_error("tco_init: meshes not set"); return COSYB200_ESTATE; } }
  DeviceGuard guard(h->device);
  LaunchScope ls(h, CAT_GEOMETRY, (cudaStream_t)stream);
  k_tco_init<<<B, GEO_THREADS, 0, (cudaStream_t)stream>>>(B, zup, boxes, K, label_ids, h->n_labels, h->pts_sampled, N_SAMPLE, TCO);
  CB_LAUNCH_CHECK();
  return COSYB200_OK;
}

int cosyb200_prepare_iter(cosyb200_handle* h, int B, int img_h, int img_w, const float* K,
                          const float* TCO, const int32_t* label_ids, float* boxes_rend,
                          float* boxes_crop, float* K_crop, void* stream) {
  CB_CHECK_ARG(h != nullptr && B >= 1, "prepare_iter: bad arguments");
  if (!h->pts_sampled) { set_error("prepare_iter: meshes not set"); return COSYB200_ESTATE; }
  CB_CHECK_ARG(img_h > 0 && img_w > 0, "prepare_iter: image size");
  DeviceGuard guard(h->device);
  float aspect = (float)((double)std::max(img_h, img_w) / (double)std::min(img_h, img_w));
  LaunchScope ls(h, CAT_GEOMETRY, (cudaStream_t)stream);
  k_project_boxes<<<B, GEO_THREADS, 0, (cudaStream_t)stream>>>(B, K, TCO, label_ids, h->n_labels, h->pts_sampled, N_SAMPLE,
                                                              aspect, boxes_rend, boxes_crop, K_crop);
  CB_LAUNCH_CHECK();
  return COSYB200_OK;
}

static int launch_crop(cosyb200_handle* h, int B, const float* images, int n_images, int img_h, int img_w,
                       const int32_t* im_ids, const float* boxes_crop, float* crops, cudaStream_t st) {
  const size_t smem = CROP_SMEM_FLOATS * sizeof(float);
  dim3 grid(RENDER_H / CROP_PH, B);
  {
    LaunchScope ls(h, CAT_CROP, st);
    k_roi_crop<<<grid, CROP_THREADS, smem, st>>>(B, images, n_images, img_h, img_w, im_ids, boxes_crop, crops);
  }
  CB_LAUNCH_CHECK();
  return 0;
}

int cosyb200_roi_crop(cosyb200_handle* h, int B, const float* images, int n_images, int img_h,
                      int img_w, const int32_t* im_ids, const float* boxes_crop, float* crops,
                      void* stream) {
  CB_CHECK_ARG(h != nullptr && B >= 1, "roi_crop: bad arguments");
  CB_CHECK_ARG(n_images >= 1 && img_h >= 2 && img_w >= 2, "roi_crop: image size");
  DeviceGuard guard(h->device);
  return launch_crop(h, B, images, n_images, img_h, img_w, im_ids, boxes_crop, crops, (cudaStream_t)stream);
}

int cosyb200_net_forward(cosyb200_handle* h, int slot, int B, const float* crops, const float* renders,
                         float* pose9, float* const* taps, void* stream) {
  if (int rc = check_batch(h, B, "net_forward")) return rc;
  CB_CHECK_ARG(slot == 0 || slot == 1, "net_forward: slot %d", slot);
  if (!h->models[slot].loaded) { set_error("net_forward: model slot %d not loaded", slot); return COSYB200_ESTATE; }
  DeviceGuard guard(h->device);
  return net_forward(h, slot, B, crops, renders, 0, pose9, taps, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int cosyb200_update_pose(cosyb200_handle* h, int B, const float* TCO_in, const float* K_crop,
                         const float* pose9, float* TCO_out, void* stream) {
  CB_CHECK_ARG(h != nullptr && B >= 1, "update_pose: bad arguments");
  DeviceGuard guard(h->device);
  LaunchScope ls(h, CAT_GEOMETRY, (cudaStream_t)stream);
  k_update_pose<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, TCO_in, K_crop, pose9, TCO_out);
  CB_LAUNCH_CHECK();
  return COSYB200_OK;
}

// ---- device rasteriser (kernels_raster.cuh) ----------------------------------------------------------------
int cosyb200_set_render_meshes(cosyb200_handle* h, int n_labels, int64_t n_vertices, const float* vertices,
                               const float* colors, int64_t n_faces, const int32_t* faces, const int32_t* face_offsets) {
  CB_CHECK_ARG(h != nullptr, "set_render_meshes: null handle");
  CB_CHECK_ARG(n_labels >= 1 && n_vertices >= 1 && n_faces >= 1 && vertices && colors && faces && face_offsets,
               "set_render_meshes: bad tables");
  CB_CHECK_ARG(n_vertices < (1ll << 31) && n_faces < (1ll << 31), "set_render_meshes: more than 2^31 vertices or faces");
  CB_CHECK_ARG(face_offsets[0] == 0 && face_offsets[n_labels] == n_faces, "set_render_meshes: face_offsets must run 0..n_faces");
  int max_faces = 0;
  for (int l = 0; l < n_labels; ++l) {
    CB_CHECK_ARG(face_offsets[l + 1] >= face_offsets[l], "set_render_meshes: face_offsets must not decrease");
    max_faces = std::max(max_faces, face_offsets[l + 1] - face_offsets[l]);
  }
  for (int64_t i = 0; i < 3 * n_faces; ++i)
    CB_CHECK_ARG(faces[i] >= 0 && faces[i] < n_vertices, "set_render_meshes: face %lld names vertex %d of %lld",
                 (long long)(i / 3), faces[i], (long long)n_vertices);
  DeviceGuard guard(h->device);
  clear_graphs(h);
  h->model_epoch += 1;
  for (void* p : {(void*)h->r_verts, (void*)h->r_colors, (void*)h->r_faces, (void*)h->r_face_off}) if (p) cudaFree(p);
  h->r_verts = h->r_colors = nullptr; h->r_faces = h->r_face_off = nullptr;
  h->r_labels = 0;
  CB_CUDA(cudaMalloc((void**)&h->r_verts, (size_t)n_vertices * 12));
  CB_CUDA(cudaMemcpy(h->r_verts, vertices, (size_t)n_vertices * 12, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMalloc((void**)&h->r_colors, (size_t)n_vertices * 12));
  CB_CUDA(cudaMemcpy(h->r_colors, colors, (size_t)n_vertices * 12, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMalloc((void**)&h->r_faces, (size_t)n_faces * 12));
  CB_CUDA(cudaMemcpy(h->r_faces, faces, (size_t)n_faces * 12, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMalloc((void**)&h->r_face_off, (size_t)(n_labels + 1) * 4));
  CB_CUDA(cudaMemcpy(h->r_face_off, face_offsets, (size_t)(n_labels + 1) * 4, cudaMemcpyHostToDevice));
  if (!h->r_zbuf) CB_CUDA(cudaMalloc((void**)&h->r_zbuf, (size_t)h->max_batch * RENDER_H * RENDER_W * 8));
  if (!h->r_frames) CB_CUDA(cudaMalloc((void**)&h->r_frames, (size_t)h->max_batch * RENDER_H * RENDER_W * 3));
  if (!h->r_big_cnt) CB_CUDA(cudaMalloc((void**)&h->r_big_cnt, (size_t)h->max_batch * 4));
  if (h->r_big_list) { cudaFree(h->r_big_list); h->r_big_list = nullptr; }
  CB_CUDA(cudaMalloc((void**)&h->r_big_list, (size_t)h->max_batch * max_faces * 4));
  h->r_labels = n_labels;
  h->r_max_faces = max_faces;
  return COSYB200_OK;
}

static int launch_render(cosyb200_handle* h, int B, const int32_t* label_ids, const float* TCO, const float* K,
                         void* out, int out_u8, float* depth, cudaStream_t st) {
  if (h->r_labels == 0) { set_error("render: render meshes not set (cosyb200_set_render_meshes)"); return COSYB200_ESTATE; }
  const size_t npix = (size_t)B * RENDER_H * RENDER_W;
  static_assert(RENDER_H % raster::TILE == 0 && RENDER_W % raster::TILE == 0, "resolve tiles must cover the view");
  CB_CUDA(cudaMemsetAsync(h->r_zbuf, 0xff, npix * 8, st));
  CB_CUDA(cudaMemsetAsync(h->r_big_cnt, 0, (size_t)B * 4, st));
  {
    LaunchScope ls(h, CAT_RENDER, st);
    const dim3 grid((h->r_max_faces + raster::RT_THREADS - 1) / raster::RT_THREADS, B);
    raster::k_raster_tris<<<grid, raster::RT_THREADS, 0, st>>>(h->r_verts, h->r_faces, h->r_face_off, h->r_labels, label_ids,
                                                               TCO, K, h->r_zbuf, h->r_big_cnt, h->r_big_list,
                                                               h->r_max_faces);
  }
  {
    LaunchScope ls(h, CAT_RENDER, st);
    const dim3 grid(RENDER_W / raster::TILE, RENDER_H / raster::TILE, B);
    if (out_u8)
      raster::k_raster_resolve<true><<<grid, 256, 0, st>>>(h->r_verts, h->r_colors, h->r_faces, TCO, K, h->r_zbuf,
                                                           h->r_big_cnt, h->r_big_list, h->r_max_faces, out, depth);
    else
      raster::k_raster_resolve<false><<<grid, 256, 0, st>>>(h->r_verts, h->r_colors, h->r_faces, TCO, K, h->r_zbuf,
                                                            h->r_big_cnt, h->r_big_list, h->r_max_faces, out, depth);
  }
  CB_LAUNCH_CHECK();
  return COSYB200_OK;
}

int cosyb200_render(cosyb200_handle* h, int B, const int32_t* label_ids, const float* TCO, const float* K, void* out,
                    int out_u8, float* depth, void* stream) {
  if (int rc = check_batch(h, B, "render")) return rc;
  CB_CHECK_ARG(label_ids && TCO && K && out, "render: null argument");
  DeviceGuard guard(h->device);
  return launch_render(h, B, label_ids, TCO, K, out, out_u8, depth, (cudaStream_t)stream);
}

int cosyb200_refine_iter(cosyb200_handle* h, int slot, int B, const float* images, int n_images,
                         int img_h, int img_w, const int32_t* im_ids, const float* boxes_crop,
                         const void* renders, int render_u8, const float* K_crop, const float* TCO_in,
                         float* pose9, float* TCO_out, void* stream) {
  if (int rc = check_batch(h, B, "refine_iter")) return rc;
  CB_CHECK_ARG(slot == 0 || slot == 1, "refine_iter: slot %d", slot);
  if (!h->models[slot].loaded) { set_error("refine_iter: model slot %d not loaded", slot); return COSYB200_ESTATE; }
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = launch_crop(h, B, images, n_images, img_h, img_w, im_ids, boxes_crop, h->crops, st)) return rc;
  return net_forward(h, slot, B, h->crops, renders, render_u8, pose9 ? pose9 : h->pose9, nullptr, TCO_in, K_crop, TCO_out, st);
}

int cosyb200_refine_n(cosyb200_handle* h, int slot, int B, int n_iter, const float* images,
                      int n_images, int img_h, int img_w, const int32_t* im_ids, const float* K,
                      const int32_t* label_ids, const void* renders, int render_u8, const float* TCO_in,
                      float* TCO_out, float* K_crop, float* boxes_rend, float* boxes_crop,
                      float* pose9, void* stream) {
  if (int rc = check_batch(h, B, "refine_n")) return rc;
  CB_CHECK_ARG(n_iter >= 1, "refine_n: n_iter %d", n_iter);
  if (!renders && h->r_labels == 0) { set_error("refine_n: no views given and no render meshes set"); return COSYB200_ESTATE; }
  if (!renders && h->r_labels != h->n_labels) {
    set_error("refine_n: render meshes cover %d labels, the mesh tables %d (the same label ids index both)", h->r_labels, h->n_labels);
    return COSYB200_ESTATE;
  }
  struct Io { const int32_t* im_ids; const float* K; const int32_t* label_ids; const float* TCO_in;
              float *TCO_out, *K_crop, *boxes_rend, *boxes_crop, *pose9; };
  auto run = [&](void* st, const Io& io) -> int {
    for (int n = 0; n < n_iter; ++n) {
      const float* tin = n == 0 ? io.TCO_in : io.TCO_out + (size_t)(n - 1) * B * 16;
      int rc = cosyb200_prepare_iter(h, B, img_h, img_w, io.K, tin, io.label_ids, io.boxes_rend + (size_t)n * B * 4,
                                     io.boxes_crop + (size_t)n * B * 4, io.K_crop + (size_t)n * B * 9, st);
      if (rc) return rc;
      const void* views = (const char*)renders + (size_t)n * B * 3 * RENDER_H * RENDER_W * (render_u8 ? 1 : 4);
      int views_u8 = render_u8;
      if (!renders) {   // no views handed in: rasterise the hypotheses at their current poses (pose.py:100-102)
        DeviceGuard guard(h->device);
        rc = launch_render(h, B, io.label_ids, tin, io.K_crop + (size_t)n * B * 9, h->r_frames, 1, nullptr, (cudaStream_t)st);
        if (rc) return rc;
        views = h->r_frames;
        views_u8 = 1;
      }
      rc = cosyb200_refine_iter(h, slot, B, images, n_images, img_h, img_w, io.im_ids,
                                io.boxes_crop + (size_t)n * B * 4, views, views_u8, io.K_crop + (size_t)n * B * 9, tin,
                                io.pose9 + (size_t)n * B * POSE_DIM, io.TCO_out + (size_t)n * B * 16, st);
      if (rc) return rc;
    }
    return COSYB200_OK;
  };
  const Io direct{im_ids, K, label_ids, TCO_in, TCO_out, K_crop, boxes_rend, boxes_crop, pose9};
  const bool graph_ok = h->use_graph && !h->profiling && h->dump_block < 0 && h->trace_block < 0 &&
                        n_iter <= cosyb200_handle::GRAPH_MAX_ITER;
  if (!graph_ok) return run(stream, direct);

  // Graph path.  The small per-hypothesis inputs and outputs go through buffers owned by the handle (plain async
  // copies around the graph launch), so the captured launches only depend on the shape of the call, the frame and
  // view buffers and the engine options: the graph is replayed whenever those repeat.
  DeviceGuard guard(h->device);
  cudaStream_t cst = (cudaStream_t)stream;
  const size_t Bmax = (size_t)h->max_batch, NI = cosyb200_handle::GRAPH_MAX_ITER;
  if (!h->io_buf) {
    // layout (floats): K 9B | TCO_in 16B | im_ids B | label_ids B | TCO_out 16B NI | K_crop 9B NI | boxes_rend 4B NI |
    //                  boxes_crop 4B NI | pose9 9B NI
    void* p = nullptr;
    if (int rc = dev_alloc(&p, Bmax * (9 + 16 + 2 + NI * (16 + 9 + 4 + 4 + POSE_DIM)) * 4)) return rc;
    h->io_buf = (float*)p;
  }
  float* b_K = h->io_buf;
  float* b_TCO = b_K + Bmax * 9;
  int32_t* b_im = (int32_t*)(b_TCO + Bmax * 16);
  int32_t* b_lab = b_im + Bmax;
  float* b_out = (float*)(b_lab + Bmax);
  float* b_Kc = b_out + Bmax * NI * 16;
  float* b_br = b_Kc + Bmax * NI * 9;
  float* b_bc = b_br + Bmax * NI * 4;
  float* b_p9 = b_bc + Bmax * NI * 4;
  const Io staged{b_im, b_K, b_lab, b_TCO, b_out, b_Kc, b_br, b_bc, b_p9};
  const std::vector<uint64_t> key = {
      (uint64_t)slot, (uint64_t)B, (uint64_t)n_iter, (uint64_t)n_images, (uint64_t)img_h, (uint64_t)img_w,
      (uint64_t)render_u8, (uint64_t)images, (uint64_t)renders,
      (uint64_t)h->gemm_impl, (uint64_t)h->xdw, (uint64_t)h->dw_impl, (uint64_t)h->tc_groups, h->model_epoch};
  cosyb200_handle::RefineGraph* g = nullptr;
  for (auto& e : h->graphs)
    if (e.key == key) g = &e;
  if (!g) {
    // Callers that hand over fresh frame / view buffers on every call would pay capture + instantiation each time:
    // after a few misses in a row go back to plain launches and only retry a capture now and then.
    h->graph_miss_streak += 1;
    if (h->graph_miss_streak > 3 && h->graph_miss_streak % 16 != 0) return run(stream, direct);
    if (!h->cap_stream) CB_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    int64_t before[cosyb200_handle::N_CAT];
    for (int i = 0; i < cosyb200_handle::N_CAT; ++i) before[i] = h->launches[i];
    CB_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeRelaxed));
    const int rc = run(h->cap_stream, staged);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) { set_error("refine_n: stream capture failed: %s", cudaGetErrorString(ce)); return COSYB200_ECUDA; }
    cosyb200_handle::RefineGraph e;
    e.key = key;
    const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { set_error("refine_n: cudaGraphInstantiate: %s", cudaGetErrorString(ie)); return COSYB200_ECUDA; }
    for (int i = 0; i < cosyb200_handle::N_CAT; ++i) {
      e.launches[i] = h->launches[i] - before[i];
      h->launches[i] = before[i];
    }
    if (h->graphs.size() >= 16) {            // evict the least recently used entry
      size_t lru = 0;
      for (size_t i = 1; i < h->graphs.size(); ++i)
        if (h->graphs[i].last_use < h->graphs[lru].last_use) lru = i;
      cudaGraphExecDestroy(h->graphs[lru].exec);
      h->graphs.erase(h->graphs.begin() + lru);
    }
    h->graphs.push_back(e);
    g = &h->graphs.back();
  }
  else h->graph_miss_streak = 0;
  g->last_use = ++h->graph_clock;
  const size_t nb = (size_t)n_iter * B;
  CB_CUDA(cudaMemcpyAsync(b_K, K, (size_t)B * 9 * 4, cudaMemcpyDeviceToDevice, cst));
  CB_CUDA(cudaMemcpyAsync(b_TCO, TCO_in, (size_t)B * 16 * 4, cudaMemcpyDeviceToDevice, cst));
  CB_CUDA(cudaMemcpyAsync(b_im, im_ids, (size_t)B * 4, cudaMemcpyDeviceToDevice, cst));
  CB_CUDA(cudaMemcpyAsync(b_lab, label_ids, (size_t)B * 4, cudaMemcpyDeviceToDevice, cst));
  CB_CUDA(cudaGraphLaunch(g->exec, cst));
  CB_CUDA(cudaMemcpyAsync(TCO_out, b_out, nb * 16 * 4, cudaMemcpyDeviceToDevice, cst));
  CB_CUDA(cudaMemcpyAsync(K_crop, b_Kc, nb * 9 * 4, cudaMemcpyDeviceToDevice, cst));
  CB_CUDA(cudaMemcpyAsync(boxes_rend, b_br, nb * 4 * 4, cudaMemcpyDeviceToDevice, cst));
  CB_CUDA(cudaMemcpyAsync(boxes_crop, b_bc, nb * 4 * 4, cudaMemcpyDeviceToDevice, cst));
  CB_CUDA(cudaMemcpyAsync(pose9, b_p9, nb * POSE_DIM * 4, cudaMemcpyDeviceToDevice, cst));
  for (int i = 0; i < cosyb200_handle::N_CAT; ++i) h->launches[i] += g->launches[i];
  return COSYB200_OK;
}

int cosyb200_set_option(cosyb200_handle* h, const char* name, int value) {
  CB_CHECK_ARG(h != nullptr && name != nullptr, "set_option: bad arguments");
  if (strcmp(name, "gemm_impl") == 0) {
    CB_CHECK_ARG(value >= 0 && value <= 2, "set_option: gemm_impl must be 0 (cuda cores), 1 (tcgen05 3xTF32) or 2 (tcgen05 3xFP16)");
    h->gemm_impl = value;
    return COSYB200_OK;
  }
  if (strcmp(name, "tc_groups") == 0) {
    CB_CHECK_ARG(value >= 0 && value <= 2, "set_option: tc_groups must be 0 (per layer), 1 or 2");
    h->tc_groups = value;
    return COSYB200_OK;
  }
  if (strcmp(name, "graph") == 0) {
    CB_CHECK_ARG(value == 0 || value == 1, "set_option: graph must be 0 (plain launches) or 1 (refine_n replays CUDA graphs)");
    h->use_graph = value;
    return COSYB200_OK;
  }
  if (strcmp(name, "xdw") == 0) {
    CB_CHECK_ARG(value == 0 || value == 1, "set_option: xdw must be 0 (separate expand and depthwise kernels) or 1 (fused)");
    h->xdw = value;
    return COSYB200_OK;
  }
  if (strcmp(name, "trace_block") == 0) {
    h->trace_block = value;
    return COSYB200_OK;
  }
  if (strcmp(name, "pw2_nt") == 0) {   // tuning aid: force the number of n-tile columns of k_pw2 (0 = cost model)
    CB_CHECK_ARG(value >= 0 && value <= 64, "set_option: pw2_nt must be in [0, 64]");
    h->pw2_nt = value;
    clear_graphs(h);
    return COSYB200_OK;
  }
  if (strcmp(name, "dw_impl") == 0) {
    CB_CHECK_ARG(value == 0 || value == 1, "set_option: dw_impl must be 0 (rolling window + separate SE) or 1 (tiled + split SE)");
    h->dw_impl = value;
    return COSYB200_OK;
  }
  set_error("set_option: unknown option %s", name);
  return COSYB200_EINVAL;
}

int cosyb200_debug_pointwise(cosyb200_handle* h, int impl, int M, int N, int K, const float* A,
                             const float* W_nk_host, const float* bias_host, const float* gate,
                             int rows_per_img, const float* resid, int swish, float* C, void* stream) {
  CB_CHECK_ARG(h != nullptr && M >= 1 && N >= 8 && K >= 8 && N % 8 == 0 && K % 8 == 0, "debug_pointwise: bad sizes");
  CB_CHECK_ARG(A && W_nk_host && bias_host && C, "debug_pointwise: null pointer");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<float> w;
  float p2_inv = 1.f;
  if (impl == 2) {
    const float sc = pw2::weight_scale(W_nk_host, (size_t)N * K);
    const std::vector<uint16_t> img = pw2::pack_weights(W_nk_host, N, K, sc);
    w.resize((img.size() + 1) / 2);
    memcpy(w.data(), img.data(), img.size() * 2);
    p2_inv = 1.0f / sc;
  } else if (impl == 1) {
    w = tc::pack_weights(W_nk_host, N, K);
  } else {
    w.resize((size_t)N * K);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) w[(size_t)k * N + n] = W_nk_host[(size_t)n * K + k];
  }
  float *dW = nullptr, *dB = nullptr;
  CB_CUDA(cudaMalloc((void**)&dW, w.size() * 4));
  CB_CUDA(cudaMalloc((void**)&dB, (size_t)N * 4));
  CB_CUDA(cudaMemcpy(dW, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  CB_CUDA(cudaMemcpy(dB, bias_host, (size_t)N * 4, cudaMemcpyHostToDevice));
  int rc = 0;
  {
    LaunchScope ls(h, CAT_EXPAND, st);
    if (impl == 2) rc = launch_pw2(h, gate != nullptr, swish != 0, resid != nullptr, A, dW, p2_inv, dB, gate, resid, C, M,
                                   N, K, rows_per_img, st);
    else if (impl == 1) rc = launch_gemm_tc(h, gate != nullptr, swish != 0, resid != nullptr, A, dW, dB, gate, resid, C, M, N, K,
                                       rows_per_img, st);
    else launch_gemm(gate != nullptr, swish != 0, resid != nullptr, A, dW, dB, gate, resid, C, M, N, K, rows_per_img, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(dW);
  cudaFree(dB);
  if (rc) return rc;
  if (e != cudaSuccess) { set_error("debug_pointwise: %s", cudaGetErrorString(e)); return COSYB200_ECUDA; }
  CB_LAUNCH_CHECK();
  return COSYB200_OK;
}

int cosyb200_debug_dump(cosyb200_handle* h, int block, float* expanded, float* dw_out, float* gate) {
  CB_CHECK_ARG(h != nullptr, "debug_dump: null handle");
  h->dump_block = block;
  h->dump_e = expanded; h->dump_d = dw_out; h->dump_gate = gate;
  return COSYB200_OK;
}

int cosyb200_debug_trace(cosyb200_handle* h, long long* trace_dev) {
  CB_CHECK_ARG(h != nullptr, "debug_trace: null handle");
  DeviceGuard guard(h->device);
  CB_CUDA(cudaMemcpyToSymbol(tc::g_trace, &trace_dev, sizeof(trace_dev)));
  return COSYB200_OK;
}

// ---- multi-GPU exchange ---------------------------------------------------------------------------
int cosyb200_nccl_unique_id(char* id128_host) {
  CB_CHECK_ARG(id128_host != nullptr, "nccl_unique_id: null pointer");
  if (int rc = nccl_bind()) return rc;
  CB_NCCL(g_nccl.GetUniqueId(id128_host));
  return COSYB200_OK;
}

int cosyb200_nccl_comm_init(cosyb200_handle* h, int world, int rank, const char* id128_host) {
  CB_CHECK_ARG(h != nullptr && id128_host != nullptr && world >= 1 && rank >= 0 && rank < world, "nccl_comm_init: bad arguments");
  if (int rc = nccl_bind()) return rc;
  DeviceGuard guard(h->device);
  if (h->nccl_comm) { g_nccl.CommDestroy(h->nccl_comm); h->nccl_comm = nullptr; }
  NcclId id;
  memcpy(id.b, id128_host, 128);
  CB_NCCL(g_nccl.CommInitRank(&h->nccl_comm, world, id, rank));
  h->nccl_world = world;
  h->nccl_rank = rank;
  return COSYB200_OK;
}

int cosyb200_nccl_comm_destroy(cosyb200_handle* h) {
  CB_CHECK_ARG(h != nullptr, "nccl_comm_destroy: null handle");
  if (h->nccl_comm) {
    DeviceGuard guard(h->device);
    g_nccl.CommDestroy(h->nccl_comm);
    h->nccl_comm = nullptr;
  }
  return COSYB200_OK;
}

int cosyb200_allgather_candidates(cosyb200_handle* h, const float* local_dev, float* all_dev,
                                  int64_t count_per_rank, void* stream) {
  CB_CHECK_ARG(h != nullptr && all_dev != nullptr && count_per_rank >= 0, "allgather_candidates: bad arguments");
  if (!h->nccl_comm) { set_error("allgather_candidates: call cosyb200_nccl_comm_init first"); return COSYB200_ESTATE; }
  DeviceGuard guard(h->device);
  const float* send = local_dev ? local_dev : all_dev + (size_t)h->nccl_rank * count_per_rank;   // in place
  CB_NCCL(g_nccl.AllGather(send, all_dev, (size_t)count_per_rank, /* ncclFloat32 */ 7, h->nccl_comm, (cudaStream_t)stream));
  return COSYB200_OK;
}

// ---- launch accounting ------------------------------------------------------------------------
int cosyb200_profile_enable(cosyb200_handle* h, int on) {
  CB_CHECK_ARG(h != nullptr, "profile_enable: null handle");
  DeviceGuard guard(h->device);
  if (int rc = prof_resolve(h)) return rc;
  h->profiling = on != 0;
  return COSYB200_OK;
}

int cosyb200_profile_read(cosyb200_handle* h, int reset, int64_t* launches10, double* ms10) {
  CB_CHECK_ARG(h != nullptr, "profile_read: null handle");
  DeviceGuard guard(h->device);
  if (int rc = prof_resolve(h)) return rc;
  for (int i = 0; i < cosyb200_handle::N_CAT; ++i) {
    if (launches10) launches10[i] = h->launches[i];
    if (ms10) ms10[i] = h->cat_ms[i];
    if (reset) { h->launches[i] = 0; h->cat_ms[i] = 0; }
  }
  return COSYB200_OK;
}

int cosyb200_profile_read_blocks(cosyb200_handle* h, int reset, double* ms /*[10][32]*/) {
  CB_CHECK_ARG(h != nullptr && ms != nullptr, "profile_read_blocks: bad arguments");
  DeviceGuard guard(h->device);
  if (int rc = prof_resolve(h)) return rc;
  for (int c = 0; c < cosyb200_handle::N_CAT; ++c)
    for (int b = 0; b < cosyb200_handle::N_BLK; ++b) {
      ms[c * cosyb200_handle::N_BLK + b] = h->blk_ms[c][b];
      if (reset) h->blk_ms[c][b] = 0;
    }
  return COSYB200_OK;
}

// ---- multiview ------------------------------------------------------------------------------
static int pick_gs(int s_max) {
  int gs = 1;
  while (gs < s_max && gs < 32) gs <<= 1;
  return gs;
}

#define DISPATCH_GS(gs, KERNEL, n_rows, ...)                                              \
  do {                                                                                    \
    int64_t threads_ = (int64_t)(n_rows) * (gs);                                          \
    unsigned blocks_ = (unsigned)((threads_ + RANSAC_THREADS - 1) / RANSAC_THREADS);      \
    switch (gs) {                                                                         \
      case 1: KERNEL<1><<<blocks_, RANSAC_THREADS, 0, st>>>(__VA_ARGS__); break;          \
      case 2: KERNEL<2><<<blocks_, RANSAC_THREADS, 0, st>>>(__VA_ARGS__); break;          \
      case 4: KERNEL<4><<<blocks_, RANSAC_THREADS, 0, st>>>(__VA_ARGS__); break;          \
      case 8: KERNEL<8><<<blocks_, RANSAC_THREADS, 0, st>>>(__VA_ARGS__); break;          \
      case 16: KERNEL<16><<<blocks_, RANSAC_THREADS, 0, st>>>(__VA_ARGS__); break;        \
      default: KERNEL<32><<<blocks_, RANSAC_THREADS, 0, st>>>(__VA_ARGS__); break;        \
    }                                                                                     \
  } while (0)

int cosyb200_ransac_models(cosyb200_handle* h, int64_t n_seeds, const float* poses,
                           const int32_t* cand_labels, const int32_t* seeds, float* TC1C2, void* stream) {
  CB_CHECK_ARG(h != nullptr && n_seeds >= 0, "ransac_models: bad arguments");
  if (!h->sym) { set_error("ransac_models: meshes not set"); return COSYB200_ESTATE; }
  if (n_seeds == 0) return COSYB200_OK;
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  int gs = pick_gs(h->s_max);
  LaunchScope ls(h, CAT_RANSAC, st);
  DISPATCH_GS(gs, k_ransac_models, n_seeds, n_seeds, poses, cand_labels, seeds, h->aabb, h->sym, h->n_sym, h->s_max, TC1C2);
  CB_LAUNCH_CHECK();
  return COSYB200_OK;
}

int cosyb200_ransac_score(cosyb200_handle* h, int64_t n, const float* poses, const int32_t* cand_labels,
                          const int32_t* tmatches, const float* TC1C2, float* dists, void* stream) {
  CB_CHECK_ARG(h != nullptr && n >= 0, "ransac_score: bad arguments");
  if (!h->sym) { set_error("ransac_score: meshes not set"); return COSYB200_ESTATE; }
  if (n == 0) return COSYB200_OK;
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  int gs = pick_gs(h->s_max);
  LaunchScope ls(h, CAT_RANSAC, st);
  DISPATCH_GS(gs, k_ransac_score, n, n, poses, cand_labels, tmatches, TC1C2, h->aabb, h->sym, h->s_max, dists);
  CB_LAUNCH_CHECK();
  return COSYB200_OK;
}

int cosyb200_symmetric_distance(cosyb200_handle* h, int64_t n, const float* T1, const float* T2,
                                const int32_t* label_ids, float* dists, int32_t* best_sym, void* stream) {
  CB_CHECK_ARG(h != nullptr && n >= 0, "symmetric_distance: bad arguments");
  if (!h->sym) { set_error("symmetric_distance: meshes not set"); return COSYB200_ESTATE; }
  if (n == 0) return COSYB200_OK;
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  int gs = pick_gs(h->s_max);
  LaunchScope ls(h, CAT_RANSAC, st);
  DISPATCH_GS(gs, k_symmetric_distance, n, n, T1, T2, label_ids, h->aabb, h->sym, h->s_max, dists, best_sym);
  CB_LAUNCH_CHECK();
  return COSYB200_OK;
}

int cosyb200_compose_inv(cosyb200_handle* h, int64_t n, const float* A, const int32_t* ia, const float* B,
                         const int32_t* ib, float* out, void* stream) {
  CB_CHECK_ARG(h != nullptr && n >= 0 && A && B && out, "compose_inv: bad arguments");
  if (n == 0) return COSYB200_OK;
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope ls(h, CAT_RANSAC, st);
  k_compose_inv<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, A, ia, B, ib, out);
  CB_LAUNCH_CHECK();
  return COSYB200_OK;
}

int cosyb200_ba_linearize(cosyb200_handle* h, int n_cand, int n_obj, int n_view, int n_pts,
                          const float* cand_TCO, const int32_t* cand_obj, const int32_t* cand_view,
                          const int32_t* cand_label, const float* TWO_9d, const float* TCW_9d,
                          const float* K, const float* points, float residuals_threshold,
                          float* align_dists, float* aligned, float* errors, float* Jc, float* JtJ,
                          float* Jte, float* loss, void* stream) {
  CB_CHECK_ARG(h != nullptr && n_cand >= 1 && n_obj >= 1 && n_view >= 1 && n_pts >= 1, "ba_linearize: bad sizes");
  CB_CHECK_ARG(cand_TCO && cand_obj && cand_view && cand_label && TWO_9d && TCW_9d && K && points,
               "ba_linearize: null input");
  CB_CHECK_ARG(align_dists && aligned && errors && Jc && loss, "ba_linearize: null output");
  if (!h->sym) { set_error("ba_linearize: meshes not set"); return COSYB200_ESTATE; }
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  {
    LaunchScope ls(h, CAT_RANSAC, st);
    k_ba_align<<<(n_cand + 63) / 64, 64, 0, st>>>(n_cand, n_pts, cand_TCO, cand_obj, cand_view, cand_label, TWO_9d,
                                                  TCW_9d, K, points, h->sym, h->n_sym, h->s_max, align_dists, aligned);
  }
  CB_LAUNCH_CHECK();
  {
    LaunchScope ls(h, CAT_RANSAC, st);
    k_ba_residuals<<<(n_cand * n_pts + 63) / 64, 64, 0, st>>>(n_cand, n_pts, aligned, cand_obj, cand_view, cand_label,
                                                              TWO_9d, TCW_9d, K, points, errors, Jc);
  }
  CB_LAUNCH_CHECK();
  {
    LaunchScope ls(h, CAT_RANSAC, st);
    k_ba_loss<<<1, 256, 0, st>>>(n_cand * n_pts * 2, errors, residuals_threshold, loss);
  }
  CB_LAUNCH_CHECK();
  if (JtJ && Jte) {
    const int n_params = 9 * (n_obj + n_view);
    dim3 block(32, 8), grid((n_params + 1 + 31) / 32, (n_params + 7) / 8);
    LaunchScope ls(h, CAT_RANSAC, st);
    k_ba_normal<<<grid, block, 0, st>>>(n_cand, n_pts, n_obj, n_view, cand_obj, cand_view, Jc, errors, JtJ, Jte);
    CB_LAUNCH_CHECK();
  }
  return COSYB200_OK;
}

// ADD / ADD-S errors of n (prediction, ground truth) pairs (kernels_eval.cuh)
int cosyb200_pose_errors(cosyb200_handle* h, int n, int n_points, const float* T_pred, const float* T_gt,
                         const float* points, const int32_t* symmetric, float* dists, float* norm_avg,
                         float* xyz_avg, float* tco_xyz, float* tco_norm, void* stream) {
  CB_CHECK_ARG(h != nullptr && n >= 1 && n_points >= 1 && n_points <= 16000, "pose_errors: bad sizes (n_points <= 16000)");
  CB_CHECK_ARG(T_pred && T_gt && points && norm_avg && xyz_avg && tco_xyz && tco_norm, "pose_errors: null pointer");
  DeviceGuard guard(h->device);
  LaunchScope ls(h, CAT_RANSAC, (cudaStream_t)stream);
  k_pose_errors<<<n, EVAL_THREADS, (size_t)n_points * 12, (cudaStream_t)stream>>>(
      n, n_points, T_pred, T_gt, points, symmetric, dists, norm_avg, xyz_avg, tco_xyz, tco_norm);
  CB_LAUNCH_CHECK();
  return COSYB200_OK;
}

// find_ransac_inliers on the device (kernels_ransac.cuh, namespace vote): everything stays on the GPU until the two
// counters and the (short) ordered lists are read back by the caller.
int cosyb200_ransac_inliers_dev(cosyb200_handle* h, int64_t n_seeds, int64_t n_pairs, const int32_t* pair_start_dev,
                                int64_t n_mtc, const int32_t* mtc_hyp_dev, const int32_t* mtc_c1_dev,
                                const int32_t* mtc_c2_dev, const float* dists_dev, float thr, int n_min_inliers,
                                int32_t* out_c1_dev, int32_t* out_c2_dev, int32_t* best_dev, int64_t* counts_dev,
                                void* stream) {
  CB_CHECK_ARG(h != nullptr && n_seeds >= 1 && n_pairs >= 1 && n_mtc >= 1 && n_mtc < (1ll << 31), "ransac_inliers_dev: bad sizes");
  CB_CHECK_ARG(pair_start_dev && mtc_hyp_dev && mtc_c1_dev && mtc_c2_dev && dists_dev && out_c1_dev && out_c2_dev &&
               best_dev && counts_dev, "ransac_inliers_dev: null pointer");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  // workspace: flags u8 [n_mtc] | acc_rows i32 [n_mtc] | row_start, n_inl i32 [n_seeds] | dsum f32 [n_seeds] |
  //            best_h, cnt i32 [n_pairs]
  const size_t need = (size_t)n_mtc * 8 + (size_t)n_seeds * 12 + (size_t)n_pairs * 8 + 64;
  if (h->vote_ws_bytes < need) {
    if (h->vote_ws) { CB_CUDA(cudaStreamSynchronize(st)); cudaFree(h->vote_ws); h->vote_ws = nullptr; }
    CB_CUDA(cudaMalloc(&h->vote_ws, need));
    h->vote_ws_bytes = need;
  }
  int32_t* acc_rows = (int32_t*)h->vote_ws;
  int32_t* row_start = acc_rows + n_mtc;
  int32_t* n_inl = row_start + n_seeds;
  float* dsum = (float*)(n_inl + n_seeds);
  int32_t* best_h = (int32_t*)(dsum + n_seeds);
  int32_t* cnt = best_h + n_pairs;
  uint8_t* flags = (uint8_t*)(cnt + n_pairs);
  {
    LaunchScope ls(h, CAT_RANSAC, st);
    vote::k_vote_greedy<<<(unsigned)((n_seeds + vote::WARPS - 1) / vote::WARPS), vote::WARPS * 32, 0, st>>>(
        (int)n_seeds, (int)n_mtc, mtc_hyp_dev, mtc_c1_dev, mtc_c2_dev, dists_dev, thr, flags, acc_rows, row_start, n_inl, dsum);
  }
  CB_LAUNCH_CHECK();
  {
    LaunchScope ls(h, CAT_RANSAC, st);
    vote::k_vote_best<<<(unsigned)((n_pairs + vote::WARPS - 1) / vote::WARPS), vote::WARPS * 32, 0, st>>>(
        (int)n_pairs, pair_start_dev, n_inl, dsum, n_min_inliers, best_h, cnt);
  }
  CB_LAUNCH_CHECK();
  {
    LaunchScope ls(h, CAT_RANSAC, st);
    vote::k_vote_emit<<<1, 1024, 0, st>>>((int)n_pairs, best_h, cnt, row_start, acc_rows, mtc_c1_dev, mtc_c2_dev,
                                         out_c1_dev, out_c2_dev, best_dev, counts_dev);
  }
  CB_LAUNCH_CHECK();
  return COSYB200_OK;
}

// float64 evaluation of the same linearisation (see kernels_ba.cuh): JtJ64 [n_params^2], Jte64 [n_params], loss64 [1]
int cosyb200_ba_linearize_f64(cosyb200_handle* h, int n_cand, int n_obj, int n_view, int n_pts,
                              const float* cand_TCO, const int32_t* cand_obj, const int32_t* cand_view,
                              const int32_t* cand_label, const float* TWO_9d, const float* TCW_9d,
                              const float* K, const float* points, float residuals_threshold,
                              float* align_dists, float* aligned, double* JtJ64, double* Jte64, double* loss64,
                              void* stream) {
  CB_CHECK_ARG(h != nullptr && n_cand >= 1 && n_obj >= 1 && n_view >= 1 && n_pts >= 1, "ba_linearize_f64: bad sizes");
  CB_CHECK_ARG(cand_TCO && cand_obj && cand_view && cand_label && TWO_9d && TCW_9d && K && points,
               "ba_linearize_f64: null input");
  CB_CHECK_ARG(align_dists && aligned && loss64, "ba_linearize_f64: null output");
  if (!h->sym) { set_error("ba_linearize_f64: meshes not set"); return COSYB200_ESTATE; }
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_res = (size_t)n_cand * n_pts * 2;
  const size_t need = n_res * 19;
  if (h->ba_ws_elems < need) {
    if (h->ba_ws) { CB_CUDA(cudaStreamSynchronize(st)); cudaFree(h->ba_ws); h->ba_ws = nullptr; }
    CB_CUDA(cudaMalloc((void**)&h->ba_ws, need * 8));
    h->ba_ws_elems = need;
  }
  double* err64 = h->ba_ws;
  double* Jc64 = h->ba_ws + n_res;
  {
    LaunchScope ls(h, CAT_RANSAC, st);
    k_ba_align<<<(n_cand + 63) / 64, 64, 0, st>>>(n_cand, n_pts, cand_TCO, cand_obj, cand_view, cand_label, TWO_9d,
                                                  TCW_9d, K, points, h->sym, h->n_sym, h->s_max, align_dists, aligned);
  }
  CB_LAUNCH_CHECK();
  {
    LaunchScope ls(h, CAT_RANSAC, st);
    k_ba_residuals_d<<<(n_cand * n_pts + 63) / 64, 64, 0, st>>>(n_cand, n_pts, aligned, cand_obj, cand_view, cand_label,
                                                                TWO_9d, TCW_9d, K, points, err64, Jc64);
  }
  CB_LAUNCH_CHECK();
  {
    LaunchScope ls(h, CAT_RANSAC, st);
    k_ba_loss_d<<<1, 256, 0, st>>>((int)n_res, err64, (double)residuals_threshold, loss64);
  }
  CB_LAUNCH_CHECK();
  if (JtJ64 && Jte64) {
    const int n_params = 9 * (n_obj + n_view);
    dim3 block(32, 8), grid((n_params + 1 + 31) / 32, (n_params + 7) / 8);
    LaunchScope ls(h, CAT_RANSAC, st);
    k_ba_normal_d<<<grid, block, 0, st>>>(n_cand, n_pts, n_obj, n_view, cand_obj, cand_view, Jc64, err64, JtJ64, Jte64);
    CB_LAUNCH_CHECK();
  }
  return COSYB200_OK;
}

// step [n] = (JtJ64 + lambda I)^-1 Jte64 on the device (float64 Cholesky, one CTA); n_bad_pivots_dev may be NULL
int cosyb200_lm_solve(cosyb200_handle* h, int n, const double* JtJ64, const double* Jte64, double lambda,
                      float* step, int32_t* n_bad_pivots_dev, void* stream) {
  CB_CHECK_ARG(h != nullptr && n >= 1 && n <= LM_MAX_N && JtJ64 && Jte64 && step, "lm_solve: bad arguments (n <= %d)", LM_MAX_N);
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t need = (size_t)n * n;
  if (h->lm_ws_elems < need) {
    if (h->lm_ws) { CB_CUDA(cudaStreamSynchronize(st)); cudaFree(h->lm_ws); h->lm_ws = nullptr; }
    CB_CUDA(cudaMalloc((void**)&h->lm_ws, need * 8));
    h->lm_ws_elems = need;
  }
  LaunchScope ls(h, CAT_RANSAC, st);
  if (n <= LM_SMEM_MAX_N)
    k_lm_solve<true><<<1, LM_THREADS, (size_t)n * (n + 1) / 2 * 8, st>>>(n, JtJ64, Jte64, lambda, h->lm_ws, step, n_bad_pivots_dev);
  else
    k_lm_solve<false><<<1, LM_THREADS, 0, st>>>(n, JtJ64, Jte64, lambda, h->lm_ws, step, n_bad_pivots_dev);
  CB_LAUNCH_CHECK();
  return COSYB200_OK;
}

}  // extern "C"
