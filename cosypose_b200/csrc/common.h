// Shared declarations of the cosyb200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/cosyb200.h"

namespace cosyb {

void set_error(const char* fmt, ...);

#define CB_CHECK_ARG(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      cosyb::set_error(__VA_ARGS__);     \
      return COSYB200_EINVAL;            \
    }                                    \
  } while (0)

#define CB_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (expr);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      cosyb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
      return COSYB200_ECUDA;                                                            \
    }                                                                                   \
  } while (0)

#define CB_LAUNCH_CHECK() CB_CUDA(cudaGetLastError())

constexpr int RENDER_H = COSYB200_RENDER_H;
constexpr int RENDER_W = COSYB200_RENDER_W;
constexpr int N_SAMPLE = COSYB200_N_SAMPLE_POINTS;
constexpr int IN_CH = 6;
constexpr int N_FEATURES = 1536;
constexpr int POSE_DIM = 9;
constexpr float BN_EPS = 1e-3f;

// One MBConv block of the trunk (see effnet_table.h).
struct BlockSpec {
  int k, s, e, cin, cexp, cse, cout, pad_lo, pad_hi, skip;
  int hin, win, hout, wout;  // at the 240x320 render size
};

// Packed, BN-folded weights of one block, all on the device.
struct BlockWeights {
  // 1x1 convolutions are stored twice: [N][K] (K contiguous; tensor-core operand) and
  // [K][N] (N contiguous; CUDA-core kernel operand).  BN scale is folded into the rows.
  float* expand_nk = nullptr;   // [cexp][cin]
  float* expand_kn = nullptr;   // [cin][cexp]
  float* expand_bias = nullptr; // [cexp]
  float* dw_w = nullptr;        // [k*k][cexp] tap-major, BN scale folded
  float* dw_bias = nullptr;     // [cexp]
  float* se_r_w = nullptr;      // [cse][cexp]
  float* se_r_b = nullptr;      // [cse]
  float* se_e_w = nullptr;      // [cse][cexp] (transposed at load)
  float* se_e_b = nullptr;      // [cexp]
  float* proj_nk = nullptr;     // [cout][cexp]
  float* proj_kn = nullptr;     // [cexp][cout]
  float* proj_bias = nullptr;   // [cout]
  float* expand_tc = nullptr;   // tensor-core image of expand_nk (kernels_tc.cuh pack_weights)
  float* proj_tc = nullptr;     // tensor-core image of proj_nk
  void* expand_p2 = nullptr;    // fp16 hi/lo image of expand_nk (kernels_pw2.cuh pack_weights)
  void* proj_p2 = nullptr;      // fp16 hi/lo image of proj_nk
  void* expand_x = nullptr;     // fp16 hi/lo chunk images of expand_nk + bias row for the fused kernel (kernels_xdw.cuh)
  float expand_x_inv = 1.f;
  float expand_p2_inv = 1.f;    // 1 / (power-of-two weight scale) of the images
  float proj_p2_inv = 1.f;
};

// Stem weights travel as a kernel PARAMETER (constant bank): every FFMA of the stem takes its weight straight from
// c[0x0][imm], no shared-memory read (8.8 KB of the 32 KB parameter space).
struct StemWeights {
  float w[54][40];   // (ky, kx, ci) major, BN scale folded
  float bias[40];
};

struct PoseModel {
  bool loaded = false;
  float* stem_w = nullptr;     // [54][40]: (ky,kx,ci) major, BN scale folded
  float* stem_bias = nullptr;  // [40]
  StemWeights stem_host;       // the same, host copy passed by value at every stem launch
  std::vector<BlockWeights> blocks;
  float* head_nk = nullptr;    // [1536][384]
  float* head_kn = nullptr;    // [384][1536]
  float* head_bias = nullptr;  // [1536]
  float* head_tc = nullptr;    // tensor-core image of head_nk
  void* head_p2 = nullptr;     // fp16 hi/lo image of head_nk
  float head_p2_inv = 1.f;
  float* fc_w = nullptr;       // [9][1536]
  float* fc_b = nullptr;       // [9]
  std::vector<void*> allocs;
};

}  // namespace cosyb

struct cosyb200_handle {
  int device = 0;
  int max_batch = 0;
  int n_sms = 148;
  std::vector<cosyb::BlockSpec> blocks;
  cosyb::PoseModel models[2];
  // mesh tables
  int n_labels = 0, s_max = 0;
  float* pts_sampled = nullptr;  // [L][2000][3]
  float* sym = nullptr;          // [L][s_max][4][4]
  int32_t* n_sym = nullptr;      // [L]
  float* aabb = nullptr;         // [L][8][3]
  // workspaces (sized for max_batch)
  float* act[2] = {nullptr, nullptr};  // block-boundary activations, NHWC
  float* buf_e = nullptr;              // expanded activations / head output
  float* buf_d = nullptr;              // depthwise output
  float* pool_partial = nullptr;       // [B][tiles][cexp]
  float* gate = nullptr;               // [B][cexp_max]
  float* crops = nullptr;              // [B][3][240][320]
  float* pose9 = nullptr;              // [B][9] scratch
  float* pw_ws = nullptr; int* pw_flags = nullptr;   // k_pw2: partial sums + flags of tiles split over two CTAs
  // device rasteriser (cosyb200_set_render_meshes / cosyb200_render; kernels_raster.cuh)
  float* r_verts = nullptr;            // [n_vertices][3] object frame
  float* r_colors = nullptr;           // [n_vertices][3] in [0, 1]
  int32_t* r_faces = nullptr;          // [n_faces][3] vertex ids
  int32_t* r_face_off = nullptr;       // [r_labels + 1] first face of every label
  int r_labels = 0, r_max_faces = 0;
  unsigned long long* r_zbuf = nullptr;   // [max_batch][240][320] depth / triangle keys
  int* r_big_cnt = nullptr;               // [max_batch] triangles too large for a warp, per hypothesis
  int* r_big_list = nullptr;              // [max_batch][r_max_faces] their face ids
  unsigned char* r_frames = nullptr;      // [max_batch][240][320][3] views of the current iteration (refine_n without views)
  size_t act_elems = 0, e_elems = 0, d_elems = 0, partial_elems = 0;
  // launch accounting / optional per-category device timing (cosyb200_profile_*)
  static constexpr int N_CAT = 11;
  int64_t launches[N_CAT] = {0};
  double cat_ms[N_CAT] = {0};
  static constexpr int N_BLK = 32;     // per-MBConv-block split of cat_ms (index 31: outside the blocks)
  double blk_ms[N_CAT][N_BLK] = {{0}};
  int cur_block = N_BLK - 1;
  std::vector<int> ev_blk;
  bool profiling = false;
  int tc_groups = 0;   // 3xTF32 kernel: 0 = pick the producer-group variant per layer, 1 / 2 = force it
  // CUDA graphs of cosyb200_refine_n, keyed by every argument that is baked into the captured launches
  struct RefineGraph {
    std::vector<uint64_t> key;
    cudaGraphExec_t exec = nullptr;
    int64_t launches[N_CAT] = {0};
    uint64_t last_use = 0;
  };
  std::vector<RefineGraph> graphs;
  static constexpr int GRAPH_MAX_ITER = 8;
  float* io_buf = nullptr;             // staging of the small per-hypothesis inputs / outputs of the graph path
  uint64_t graph_clock = 0;
  int graph_miss_streak = 0;
  uint64_t model_epoch = 0;            // bumped whenever weights or mesh tables are (re)loaded: invalidates the graphs
  cudaStream_t cap_stream = nullptr;   // capture happens here (the caller's stream may be the legacy stream)
  double* ba_ws = nullptr; size_t ba_ws_elems = 0;   // float64 residuals + compact Jacobian of ba_linearize_f64
  void* vote_ws = nullptr; size_t vote_ws_bytes = 0;   // scratch of ransac_inliers_dev
  double* lm_ws = nullptr; size_t lm_ws_elems = 0;   // Cholesky workspace of lm_solve
  void* nccl_comm = nullptr;   // ncclComm_t created by cosyb200_nccl_comm_init
  int nccl_world = 1, nccl_rank = 0;
  int use_graph = 1;      // 1: refine_n replays a captured graph when its arguments repeat
  int trace_block = -1;   // debugging: the fused kernel of this block stamps its phases into the debug trace buffer
  int xdw = 1;         // 1: blocks with a kernels_xdw.cuh plan run expand + depthwise + pooling fused
  int dw_impl = 1;     // depthwise of the small-spatial blocks: 0 = rolling window + k_se_gate, 1 = k_dw_tile + k_se_fc2
  // debugging aid (cosyb200_debug_dump): copies of block `dump_block`'s internal tensors
  int dump_block = -1;
  float* dump_e = nullptr; float* dump_d = nullptr; float* dump_gate = nullptr;
  int pw2_nt = 0;      // tuning aid: forced n-tile column count of k_pw2 (0 = cost model)
  int gemm_impl = 2;   // 1x1 convolutions: 0 = CUDA-core fp32 kernel, 1 = tcgen05 3xTF32 kernel, 2 = tcgen05 3xFP16 kernel (kernels_pw2.cuh)
  std::vector<cudaEvent_t> ev_pool;   // pairs: [2*i] start, [2*i+1] stop
  std::vector<int> ev_cat;            // category of each recorded pair
  cudaStream_t ev_stream = nullptr;
};

namespace cosyb {
enum Cat { CAT_GEOMETRY = 0, CAT_CROP, CAT_STEM, CAT_EXPAND, CAT_DW, CAT_SE, CAT_PROJECT, CAT_HEAD,
           CAT_POOL_FC, CAT_RANSAC, CAT_RENDER };
int prof_resolve(cosyb200_handle* h);
// Scope object: counts the launch and, when profiling, brackets it with events on `st`.
struct LaunchScope {
  cosyb200_handle* h;
  cudaStream_t st;
  bool rec = false;
  LaunchScope(cosyb200_handle* h_, int cat, cudaStream_t st_) : h(h_), st(st_) {
    h->launches[cat] += 1;
    if (!h->profiling) return;
    if (h->ev_cat.size() * 2 + 2 > h->ev_pool.size()) {
      if (h->ev_pool.size() >= 16384) prof_resolve(h);
      else for (int i = 0; i < 2; ++i) { cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e); }
    }
    h->ev_stream = st;
    cudaEventRecord(h->ev_pool[h->ev_cat.size() * 2], st);
    h->ev_cat.push_back(cat);
    h->ev_blk.push_back(h->cur_block);
    rec = true;
  }
  ~LaunchScope() {
    if (rec) cudaEventRecord(h->ev_pool[(h->ev_cat.size() - 1) * 2 + 1], st);
  }
};
}  // namespace cosyb
