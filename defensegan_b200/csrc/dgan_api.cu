// C-ABI of the B200-native Defense-GAN projection loop (see include/defensegan_b200.h).
// Host side: generator plan (pixel-graph tables), weight re-layout, workspace carving and the
// on-device L-step driver.  Everything is enqueued on the caller's stream; nothing here
// synchronises the host.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <new>

#include "common.cuh"
#include "kernels_simt.cuh"
#include "kernels_tc.cuh"
#include "kernels_tc2.cuh"
#include "kernels_loop.cuh"

namespace dgan {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

// ---------------------------------------------------------------------------------------
// geometry tables
// ---------------------------------------------------------------------------------------
PairTable deconv_fwd_pairs(int h_in, int w_in, int h_used, int w_used, int in_raster) {
  if (in_raster <= 0) in_raster = w_in;
  PairTable t;
  t.off.push_back(0);
  for (int i = 0; i < h_used; ++i)
    for (int j = 0; j < w_used; ++j) {
      for (int ka = 0; ka < 5; ++ka) {
        const int oo = i + 1 - ka;
        if (oo < 0 || (oo & 1) || (oo >> 1) >= h_in) continue;
        for (int kb = 0; kb < 5; ++kb) {
          const int pp = j + 1 - kb;
          if (pp < 0 || (pp & 1) || (pp >> 1) >= w_in) continue;
          t.pairs.push_back(make_int2((oo >> 1) * in_raster + (pp >> 1), ka * 5 + kb));
        }
      }
      t.off.push_back((int)t.pairs.size());
    }
  return t;
}

PairTable deconv_bwd_pairs(int h_in, int w_in, int h_used, int w_used, int in_raster) {
  if (in_raster <= 0) in_raster = w_in;
  PairTable t;
  t.off.push_back(0);
  for (int o = 0; o < in_raster; ++o)
    for (int p = 0; p < in_raster; ++p) {
      // raster pixels outside the consumed h_in x w_in window receive no gradient (empty list -> zeros)
      for (int ka = 0; ka < 5 && o < h_in && p < w_in; ++ka) {
        const int i = 2 * o + ka - 1;
        if (i < 0 || i >= h_used) continue;
        for (int kb = 0; kb < 5; ++kb) {
          const int j = 2 * p + kb - 1;
          if (j < 0 || j >= w_used) continue;
          t.pairs.push_back(make_int2(i * w_used + j, ka * 5 + kb));
        }
      }
      t.off.push_back((int)t.pairs.size());
    }
  return t;
}

PairTable linear_fwd_pairs(int n_pix) {
  PairTable t;
  t.off.push_back(0);
  for (int q = 0; q < n_pix; ++q) {
    t.pairs.push_back(make_int2(0, q));
    t.off.push_back((int)t.pairs.size());
  }
  return t;
}

PairTable linear_bwd_pairs(int n_pix) {
  PairTable t;
  t.off.push_back(0);
  for (int q = 0; q < n_pix; ++q) t.pairs.push_back(make_int2(q, q));
  t.off.push_back((int)t.pairs.size());
  return t;
}

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
struct DevTable {
  int* off = nullptr;
  int2* pairs = nullptr;
  int n_out = 0;
  int n_pairs = 0;
};

struct GemmLayer {
  // forward: [P_in][N][C_in] -> [P_out][N][C_out]
  int P_in, C_in, P_out, C_out;
  int h_in, w_in, h_used, w_used;  // spatial geometry (Linear: 1x1 -> 4x4)
  bool relu;                       // ReLU after bias (false: CelebA Generator.5)
  DevTable fwd, bwd;
  PairTable fwd_host, bwd_host;
  // fp32 weight tiles.  forward tile t: rows = C_in (K), cols = C_out; backward: rows = C_out, cols = C_in
  const float* wf = nullptr; int wf_tile_stride = 0, wf_ld = 0;
  const float* wb = nullptr; int wb_tile_stride = 0, wb_ld = 0;
  const float* bias = nullptr;
  int bias_pstride = 0;            // Linear: bias is per flat feature f = pixel*C_out + c
  const float* bn_offset = nullptr;   // use_bn: Generator.BN{1,2,3}.offset / .scale (else null)
  const float* bn_scale = nullptr;
  int bn_per_pixel = 0;               // BN1 normalises each flat feature (axes [0]); BN2/3 each channel (axes [0,1,2])
  // fp16 K-major tiles for the tensor-core path (kernels_tc.cuh): [tile][N rows][K cols]
  TcWeights tc_f, tc_b;
  TcWeights2 tc2_f, tc2_b;
};

struct FinalLayer {
  int h_in, w_in, C_in, C_out, act;
  const float* w = nullptr;  // [25][C_out][C_in] == the TF filter layout
  const float* bias = nullptr;
  int n_bands = 0;
  size_t fwd_smem = 0, bwd_smem = 0;
};

}  // namespace dgan

using namespace dgan;

// which workspace tensor a segment of the fp16 loop kernel reads / writes
enum SegTensor : int { T_ZH = 0, T_ACT, T_DACT, T_DBLK, T_GPART };
struct SegBind {
  int in_kind = T_ZH, in_idx = 0, out_kind = T_ACT, out_idx = 0;
  const dgan::TcWeights* w1 = nullptr;
  const dgan::TcWeights2* w2 = nullptr;
  const float* bias = nullptr;
  int bias_pstride = 0;
  int mb_out_layer = -1, mb_in_layer = -1;   // hidden layer whose 1-bit ReLU masks the epilogue writes / reads
  int epi = 0, out_bytes = 2;
};

// one L-step plan (kernels_loop.cuh) uploaded for a given number of 256-row pairs
struct DevPlan {
  int n_mpairs = 0, n_pairs = 0;
  dgan::LoopPlan host;
  dgan::TcItem2* items[dgan::LOOP_MAX_SEG] = {nullptr};
  dgan::TcRec* tmpl_p[2] = {nullptr, nullptr};
  dgan::TcRec* tmpl_m = nullptr;
  uint32_t *win_rec_off = nullptr, *succ_off = nullptr, *succ = nullptr, *need = nullptr;
  unsigned long long* q_init = nullptr;
};

struct dgan_ctx {
  dgan_desc desc;
  int H = 0, W = 0, C = 0, hwc = 0;
  std::vector<GemmLayer> layers;
  FinalLayer fin;
  std::vector<void*> allocs;
  int64_t macs_per_row = 0;
  int64_t last_launches = 0;
  int64_t launches = 0;
  TcState tc;
  TcFinal tc_fin;
  TcWeights2 tc2_fin_f, tc2_fin_b;
  // fp16 path: the persistent projection-loop kernel
  std::vector<LoopSegSpec> segs;          // segments of one L-step in execution order
  std::vector<SegBind> binds;
  int n_fwd_seg = 0;                      // segments 0 .. n_fwd_seg-1 are the generator forward (+ loss)
  int n_pairs = 0;                        // co-resident CTA pairs (clusters of 2) the kernel is launched with
  std::vector<std::unique_ptr<DevPlan>> plans;
  LoopParams lp;                          // launch parameters of the most recent workspace (tensor maps are encoded once)
  const void* lp_ws = nullptr; int lp_rows = -1; const DevPlan* lp_plan = nullptr;
  uint32_t* last_status = nullptr;        // device status word of the most recent loop launch (lives in its workspace)
  // profiling (dgan_profile_*): time every launch of the production path; the fp16 loop kernel also records in-kernel
  // segment spans, per-CTA stall counters and a per-item trace.  Never on in a throughput pass.
  int profile = 0;
  unsigned long long* prof_dev = nullptr; size_t prof_cap = 0; int prof_L = 0; const DevPlan* prof_plan = nullptr;
  int loop_passes = 0;                    // generator passes (forward or backward) of the last loop launch: its FLOPs
  unsigned long long* dbg_dev = nullptr; int dbg_ctas = 0;   // per-CTA stall counters of the last profiled loop launch
  unsigned long long* trace_dev = nullptr; size_t trace_items = 0; const DevPlan* trace_plan = nullptr;   // per-item timestamps of one L-step
  int trace_step = -1;
  int n_rows_cur = 0;
  struct ProfRec { int kind; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  std::vector<std::string> kind_names;
  std::vector<double> kind_macs_per_row;
};

namespace dgan {

struct ProfScope {
  dgan_ctx* c; cudaStream_t s; bool on; dgan_ctx::ProfRec r;
  ProfScope(dgan_ctx* c_, int kind, cudaStream_t s_) : c(c_), s(s_), on(c_->profile != 0) {
    if (!on) return;
    r.kind = kind;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(r.b, s);
    c->prof.push_back(r);
  }
};

static int dev_alloc(dgan_ctx* c, void** p, size_t bytes) {
  DGAN_CUDA_CHECK(cudaMalloc(p, bytes));
  c->allocs.push_back(*p);
  return 0;
}

static int upload_table(dgan_ctx* c, const PairTable& t, DevTable* d, cudaStream_t s) {
  d->n_out = (int)t.off.size() - 1;
  d->n_pairs = (int)t.pairs.size();
  int rc;
  if ((rc = dev_alloc(c, (void**)&d->off, t.off.size() * sizeof(int)))) return rc;
  if ((rc = dev_alloc(c, (void**)&d->pairs, t.pairs.size() * sizeof(int2)))) return rc;
  // pageable-source async copies are staged by the runtime before returning
  DGAN_CUDA_CHECK(cudaMemcpyAsync(d->off, t.off.data(), t.off.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  DGAN_CUDA_CHECK(cudaMemcpyAsync(d->pairs, t.pairs.data(), t.pairs.size() * sizeof(int2), cudaMemcpyHostToDevice, s));
  return 0;
}

// out[t][c][r] = in[t][r][c]   (per-tile transpose; rows x cols -> cols x rows)
__global__ void transpose_tiles_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols,
                                       size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t per = (size_t)rows * cols;
  const size_t t = i / per, rem = i % per;
  const int r = (int)(rem / cols), cc = (int)(rem % cols);
  out[t * per + (size_t)cc * rows + r] = in[i];
}

__global__ void scale_copy_kernel(const float* __restrict__ in, int n_parts, size_t part_stride,
                                  float* __restrict__ out, float s, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = in[i];
  for (int p = 1; p < n_parts; ++p) g += in[i + (size_t)p * part_stride];
  out[i] = g * s;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------
struct Workspace {
  int n_rows = 0, n_pad = 0;
  float *z = nullptr, *v = nullptr, *g = nullptr;
  std::vector<float*> act, dact;     // fp32 path: per hidden layer output [P][n_pad][C]
  std::vector<float*> pre;           // use_bn: pre-normalisation outputs (null otherwise)
  std::vector<float*> bn_part;       // use_bn: [4][kBnSplits][G] partial sums (mean, var, S1, S2)
  std::vector<__half*> act_h, dact_h;  // fp16 path
  __half* z_h = nullptr;
  std::vector<unsigned long long*> maskbits;   // fp16 path: 1-bit ReLU masks per hidden layer output
  unsigned* mom_counter = nullptr;     // fp16 path: [n_pad / 128] tickets of the split-K Linear backward's momentum tail
  // fp16 path, the loop kernel's scheduling state: [status(4) | q_ctl(4) | depcnt(n_counters)] and the queue
  uint32_t *status = nullptr, *q_ctl = nullptr, *depcnt = nullptr;
  unsigned long long* queue = nullptr;
  size_t n_counters = 0, q_cap = 0;
  __half* dblk = nullptr;              // fp16 path: [n_blocks][n_pad][64] scaled dL/dpre of the last layer
  int n_loss_parts = 0, n_g_parts = 1;
  size_t loss_stride_n = 1, loss_stride_b = 1;   // loss_part index = n * stride_n + part * stride_b
  float *y = nullptr, *dpre = nullptr, *loss_part = nullptr, *loss = nullptr;
  size_t bytes = 0;
};

static Workspace carve(const dgan_ctx* c, int n_rows, void* base) {
  Workspace w;
  w.n_rows = n_rows;
  w.n_pad = (int)align_up((size_t)std::max(n_rows, 1), c->desc.precision == DGAN_PREC_FP16 ? 2 * kRowTile : kRowTile);
  size_t off = 0;
  char* b = (char*)base;
  auto take = [&](size_t bytes) -> void* {
    void* p = b ? (void*)(b + off) : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  const size_t np = (size_t)w.n_pad;
  const int latent = c->desc.latent_dim;
  w.z = (float*)take(np * latent * 4);
  w.v = (float*)take(np * latent * 4);
  const bool tc = c->desc.precision == DGAN_PREC_FP16;
  w.n_g_parts = tc ? TC_LINEAR_SPLIT : 1;
  w.g = (float*)take(np * latent * 4 * w.n_g_parts);
  if (tc) w.z_h = (__half*)take(np * latent * 2);
  if (tc) w.mom_counter = (unsigned*)take(np / kRowTile * sizeof(unsigned));
  if (tc) {
    // one counter per (segment, window, row pair); a window holds at least one output pixel, which bounds the plan's
    // needs whatever tiling it picks (the queue capacity follows loop_plan's rule on that bound)
    size_t per_mp = 0;
    for (const LoopSegSpec& sp : c->segs) per_mp += sp.tab->off.size() - 1;
    w.n_counters = per_mp * (np / (2 * kRowTile));
    w.status = (uint32_t*)take((8 + w.n_counters) * sizeof(uint32_t));
    if (w.status != nullptr) { w.q_ctl = w.status + 4; w.depcnt = w.status + 8; }
    w.q_cap = 1024;
    while (w.q_cap < 4 * (w.n_counters + (size_t)c->n_pairs) + 64) w.q_cap <<= 1;
    w.queue = (unsigned long long*)take(w.q_cap * sizeof(unsigned long long));
  }
  if (tc) w.dblk = (__half*)take((size_t)c->tc_fin.n_blocks * np * 64 * 2);
  w.n_loss_parts = tc ? c->tc_fin.n_blocks : c->fin.n_bands;
  w.loss_stride_n = tc ? 1 : (size_t)w.n_loss_parts;          // fp16 path: [block][n_pad] (coalesced epilogue stores)
  w.loss_stride_b = tc ? (size_t)np : 1;
  for (const GemmLayer& l : c->layers) {
    const size_t elems = (size_t)l.P_out * np * l.C_out;
    if (tc) {
      w.act_h.push_back((__half*)take(elems * 2));
      w.dact_h.push_back((__half*)take(elems * 2));
      w.maskbits.push_back((unsigned long long*)take(elems / 8));
    } else {
      w.act.push_back((float*)take(elems * 4));
      w.dact.push_back((float*)take(elems * 4));
      if (l.bn_scale != nullptr) {
        const size_t G = l.bn_per_pixel ? (size_t)l.P_out * l.C_out : (size_t)l.C_out;
        w.pre.push_back((float*)take(elems * 4));
        w.bn_part.push_back((float*)take((size_t)4 * kBnSplits * G * 4));
      } else {
        w.pre.push_back(nullptr);
        w.bn_part.push_back(nullptr);
      }
    }
  }
  w.y = (float*)take(np * c->hwc * 4);
  w.dpre = (float*)take(np * c->hwc * 4);
  w.loss_part = (float*)take(np * w.n_loss_parts * 4);
  w.loss = (float*)take(np * 4);
  w.bytes = off;
  return w;
}

// ---------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------
#define DGAN_LAUNCH_CHECK(c)                                                     \
  do {                                                                           \
    (c)->launches++;                                                             \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) {                                                     \
      set_error(std::string("kernel launch: ") + cudaGetErrorString(_e));        \
      return DGAN_ERR_CUDA;                                                      \
    }                                                                            \
  } while (0)

static int launch_bsgemm_f32(dgan_ctx* c, int epi, const float* in, int C_in, int n_pad, const float* wt,
                             int tile_stride, int ldw, const DevTable& tab, float* out, int C_out,
                             const float* bias, int bias_pstride, const float* mask_src, cudaStream_t s) {
  dim3 grid(n_pad / kRowTile, tab.n_out, C_out / 64), block(256);
  switch (epi) {
    case EPI_BIAS_RELU:
      bsgemm_f32_kernel<EPI_BIAS_RELU><<<grid, block, 0, s>>>(in, C_in, n_pad, wt, tile_stride, ldw, tab.off,
                                                              tab.pairs, out, C_out, bias, bias_pstride, mask_src);
      break;
    case EPI_BIAS:
      bsgemm_f32_kernel<EPI_BIAS><<<grid, block, 0, s>>>(in, C_in, n_pad, wt, tile_stride, ldw, tab.off, tab.pairs,
                                                         out, C_out, bias, bias_pstride, mask_src);
      break;
    case EPI_MASK:
      bsgemm_f32_kernel<EPI_MASK><<<grid, block, 0, s>>>(in, C_in, n_pad, wt, tile_stride, ldw, tab.off, tab.pairs,
                                                         out, C_out, bias, bias_pstride, mask_src);
      break;
    default:
      bsgemm_f32_kernel<EPI_NONE><<<grid, block, 0, s>>>(in, C_in, n_pad, wt, tile_stride, ldw, tab.off, tab.pairs,
                                                         out, C_out, bias, bias_pstride, mask_src);
      break;
  }
  DGAN_LAUNCH_CHECK(c);
  return 0;
}

template <typename TIN>
static int launch_final_fwd(dgan_ctx* c, const TIN* hin, const Workspace& w, const float* x, int R, int B,
                            bool want_grad, cudaStream_t s) {
  const FinalLayer& f = c->fin;
  dim3 grid(w.n_rows, f.n_bands), block(128);
  const size_t smem = f.fwd_smem;
  float* dpre = want_grad ? w.dpre : nullptr;
  float* lp = x ? w.loss_part : nullptr;
#define FF(CO, ACT)                                                                                         \
  final_fwd_loss_kernel<TIN, CO, ACT><<<grid, block, smem, s>>>(hin, w.n_pad, f.h_in, f.w_in, f.C_in, f.w, \
                                                                f.bias, x, R, B, w.y, dpre ? dpre : w.dpre, lp)
  if (f.C_out == 1 && f.act == ACT_SIGMOID) FF(1, ACT_SIGMOID);
  else if (f.C_out == 3 && f.act == ACT_TANH) FF(3, ACT_TANH);
  else { set_error("unsupported final layer"); return DGAN_ERR_UNSUPPORTED; }
#undef FF
  DGAN_LAUNCH_CHECK(c);
  return 0;
}

template <typename TOUT>
static int launch_final_bwd(dgan_ctx* c, const Workspace& w, const TOUT* mask_src, float gscale, TOUT* din,
                            cudaStream_t s) {
  const FinalLayer& f = c->fin;
  const size_t work = (size_t)f.h_in * f.w_in * w.n_pad * (f.C_in / 4);
  dim3 grid((unsigned)((work + 255) / 256)), block(256);
  if (f.C_out == 1)
    final_bwd_kernel<TOUT, 1><<<grid, block, f.bwd_smem, s>>>(w.dpre, w.n_pad, f.h_in, f.w_in, f.C_in, f.w, mask_src,
                                                              gscale, din);
  else
    final_bwd_kernel<TOUT, 3><<<grid, block, f.bwd_smem, s>>>(w.dpre, w.n_pad, f.h_in, f.w_in, f.C_in, f.w, mask_src,
                                                              gscale, din);
  DGAN_LAUNCH_CHECK(c);
  return 0;
}

// ---- one generator forward (+ loss and dL/dpre when x != null), fp32 path ------------------
static int run_forward(dgan_ctx* c, const Workspace& w, const float* x, int R, int B, bool want_grad,
                       cudaStream_t s) {
  int rc;
  const int nl = (int)c->layers.size();
  const float* in = w.z;
  for (int l = 0; l < nl; ++l) {
    const GemmLayer& L = c->layers[l];
    ProfScope ps(c, 2 * l, s);
    if (L.bn_scale != nullptr) {
      // pre = GEMM + bias;  act = relu(BN_batchstat(pre))
      if ((rc = launch_bsgemm_f32(c, EPI_BIAS, in, L.C_in, w.n_pad, L.wf, L.wf_tile_stride, L.wf_ld, L.fwd, w.pre[l], L.C_out,
                                  L.bias, L.bias_pstride, nullptr, s)))
        return rc;
      const int G = L.bn_per_pixel ? L.P_out * L.C_out : L.C_out;
      float* part = w.bn_part[l];
      float *mean_p = part, *var_p = part + (size_t)kBnSplits * G;
      dim3 rgrid(G / 32, kBnSplits);
      bn_reduce_kernel<0><<<rgrid, 256, 0, s>>>(w.pre[l], nullptr, nullptr, nullptr, nullptr, L.P_out, w.n_rows, w.n_pad, L.C_out,
                                                L.bn_per_pixel, mean_p, nullptr);
      DGAN_LAUNCH_CHECK(c);
      bn_reduce_kernel<1><<<rgrid, 256, 0, s>>>(w.pre[l], nullptr, nullptr, mean_p, nullptr, L.P_out, w.n_rows, w.n_pad, L.C_out,
                                                L.bn_per_pixel, var_p, nullptr);
      DGAN_LAUNCH_CHECK(c);
      const size_t total = (size_t)L.P_out * w.n_pad * L.C_out;
      bn_apply_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w.pre[l], mean_p, var_p, L.bn_scale, L.bn_offset, L.P_out,
                                                                           w.n_rows, w.n_pad, L.C_out, L.bn_per_pixel, w.act[l]);
      DGAN_LAUNCH_CHECK(c);
    } else if ((rc = launch_bsgemm_f32(c, L.relu ? EPI_BIAS_RELU : EPI_BIAS, in, L.C_in, w.n_pad, L.wf, L.wf_tile_stride,
                                       L.wf_ld, L.fwd, w.act[l], L.C_out, L.bias, L.bias_pstride, nullptr, s))) {
      return rc;
    }
    in = w.act[l];
  }
  ProfScope ps(c, 2 * nl, s);
  return launch_final_fwd<float>(c, in, w, x, R, B, want_grad, s);
}

// ---- backward-to-z, fp32 path: w.g = J^T dpre (unscaled by 2/HWC) -----
static int run_backward(dgan_ctx* c, const Workspace& w, cudaStream_t s) {
  int rc;
  const int nl = (int)c->layers.size();
  // d(act) -> d(pre) through ReLU + batch-statistics BN of layer l (in place in w.dact[l])
  auto bn_backward = [&](int l) -> int {
    const GemmLayer& L = c->layers[l];
    const int G = L.bn_per_pixel ? L.P_out * L.C_out : L.C_out;
    float* part = w.bn_part[l];
    float *mean_p = part, *var_p = part + (size_t)kBnSplits * G, *s1_p = part + (size_t)2 * kBnSplits * G,
          *s2_p = part + (size_t)3 * kBnSplits * G;
    dim3 rgrid(G / 32, kBnSplits);
    bn_reduce_kernel<2><<<rgrid, 256, 0, s>>>(w.pre[l], w.act[l], w.dact[l], mean_p, var_p, L.P_out, w.n_rows, w.n_pad, L.C_out,
                                              L.bn_per_pixel, s1_p, s2_p);
    DGAN_LAUNCH_CHECK(c);
    const size_t total = (size_t)L.P_out * w.n_pad * L.C_out;
    bn_apply_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w.pre[l], w.act[l], mean_p, var_p, s1_p, s2_p, L.bn_scale,
                                                                         L.P_out, w.n_rows, w.n_pad, L.C_out, L.bn_per_pixel,
                                                                         w.dact[l]);
    DGAN_LAUNCH_CHECK(c);
    return 0;
  };
  const GemmLayer& last = c->layers[nl - 1];
  {
    ProfScope ps(c, 2 * nl + 1, s);
    const bool bn = last.bn_scale != nullptr;
    if ((rc = launch_final_bwd<float>(c, w, (last.relu && !bn) ? w.act[nl - 1] : nullptr, 1.f, w.dact[nl - 1], s))) return rc;
    if (bn && (rc = bn_backward(nl - 1))) return rc;
  }
  for (int l = nl - 1; l >= 1; --l) {
    const GemmLayer& L = c->layers[l];
    const bool bn = c->layers[l - 1].bn_scale != nullptr;
    const bool mask = c->layers[l - 1].relu && !bn;
    ProfScope ps(c, 2 * l + 1, s);
    if ((rc = launch_bsgemm_f32(c, mask ? EPI_MASK : EPI_NONE, w.dact[l], L.C_out, w.n_pad, L.wb, L.wb_tile_stride,
                                L.wb_ld, L.bwd, w.dact[l - 1], L.C_in, nullptr, 0, mask ? w.act[l - 1] : nullptr, s)))
      return rc;
    if (bn && (rc = bn_backward(l - 1))) return rc;
  }
  const GemmLayer& L0 = c->layers[0];
  ProfScope ps(c, 1, s);
  return launch_bsgemm_f32(c, EPI_NONE, w.dact[0], L0.C_out, w.n_pad, L0.wb, L0.wb_tile_stride, L0.wb_ld, L0.bwd, w.g,
                           L0.C_in, nullptr, 0, nullptr, s);
}

static int run_init_z(dgan_ctx* c, const Workspace& w, const float* z0, uint64_t seed, cudaStream_t s, size_t row_offset = 0) {
  const int latent = c->desc.latent_dim;
  const size_t total4 = (size_t)w.n_pad * latent / 4;
  // the last layer's block tensor is K-padded to 64 columns; the epilogue only ever writes the 16*C_out valid ones
  if (w.dblk != nullptr) DGAN_CUDA_CHECK(cudaMemsetAsync(w.dblk, 0, (size_t)c->tc_fin.n_blocks * w.n_pad * 64 * sizeof(__half), s));
  if (w.mom_counter != nullptr) DGAN_CUDA_CHECK(cudaMemsetAsync(w.mom_counter, 0, (size_t)w.n_pad / kRowTile * sizeof(unsigned), s));
  // fp32 path: the last layer's forward writes dL/dpre for the real rows only while its backward walks all n_pad rows;
  // the tile-padding rows are never observed, but they must not be read uninitialised
  if (w.dblk == nullptr && w.n_pad > w.n_rows)
    DGAN_CUDA_CHECK(cudaMemsetAsync(w.dpre + (size_t)w.n_rows * c->hwc, 0, (size_t)(w.n_pad - w.n_rows) * c->hwc * sizeof(float), s));
  init_z_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, s>>>(w.z, w.v, w.z_h, z0, w.n_rows, w.n_pad, latent, seed,
                                                                 sqrtf(1.0f / (float)latent), row_offset * latent);
  DGAN_LAUNCH_CHECK(c);
  return 0;
}

static int check_ws(const dgan_ctx* c, int n_rows, void* ws, size_t ws_bytes, Workspace* out) {
  if (ws == nullptr) { set_error("workspace is NULL"); return DGAN_ERR_WORKSPACE; }
  if (((uintptr_t)ws & 1023) != 0) { set_error("workspace must be 1024-byte aligned"); return DGAN_ERR_WORKSPACE; }
  *out = carve(c, n_rows, ws);
  if (out->bytes > ws_bytes) {
    set_error("workspace too small: need " + std::to_string(out->bytes) + " bytes, got " + std::to_string(ws_bytes));
    return DGAN_ERR_WORKSPACE;
  }
  return 0;
}

// element counts of the weight tensors in creation order (include/defensegan_b200.h, dgan_num_weights)
static std::vector<size_t> weight_counts(const dgan_desc* d) {
  const size_t nd = (size_t)d->net_dim, latent = (size_t)d->latent_dim, feat = 16 * 4 * nd;
  std::vector<size_t> n = {latent * feat, feat};
  if (d->use_bn) { n.push_back(feat); n.push_back(feat); }
  std::vector<std::pair<size_t, size_t>> dc = {{4 * nd, 2 * nd}, {2 * nd, nd}};
  if (d->arch == DGAN_ARCH_CELEBA) { dc.push_back({nd, nd}); dc.push_back({nd, 3}); } else dc.push_back({nd, 1});
  for (size_t i = 0; i < dc.size(); ++i) {
    n.push_back(25 * dc[i].first * dc[i].second); n.push_back(dc[i].second);
    if (d->use_bn && i < 2) { n.push_back(dc[i].second); n.push_back(dc[i].second); }
  }
  return n;
}

static float grad_multiplier(const dgan_ctx* c) {
  float m = 2.0f / (float)c->hwc;  // d/dy mean_{HWC}(y-x)^2
  if (c->desc.precision == DGAN_PREC_FP16) m /= c->tc.grad_scale;
  return m;
}

// ---------------------------------------------------------------------------------------
// fp16 path: the persistent projection-loop kernel (kernels_loop.cuh)
// ---------------------------------------------------------------------------------------
static int loop_kind_of(int N, int epi, int out_bytes) {
  if (epi == EPI_FINAL_SIGMOID1) return LK_FINAL16;
  if (epi == EPI_FINAL_TANH3) return LK_FINAL48;
  if (out_bytes == 4) return N == 64 ? LK_NONE64F : (N == 128 ? LK_NONE128F : (N == 256 ? LK_NONE256F : -1));
  if (epi == EPI_BIAS_RELU) return N == 64 ? LK_BR64 : (N == 128 ? LK_BR128 : (N == 256 ? LK_BR256 : -1));
  if (epi == EPI_MASK) return N == 64 ? LK_MASK64 : (N == 128 ? LK_MASK128 : (N == 256 ? LK_MASK256 : -1));
  if (epi == EPI_BIAS) return N == 64 ? LK_B64 : -1;
  if (epi == EPI_NONE) return N == 64 ? LK_NONE64H : -1;
  return -1;
}

// The segments of one L-step in execution order: generator forward, last layer + loss, then backward-to-z.
static int build_segments(dgan_ctx* c) {
  const int nl = (int)c->layers.size();
  c->segs.clear(); c->binds.clear();
  auto add = [&](const std::string& name, const TcWeights& w1, const TcWeights2& w2, int epi, int out_bytes, int in_seg,
                 double macs, SegBind b) -> int {
    LoopSegSpec sp;
    sp.name = name; sp.N = w1.N; sp.K = w1.K; sp.kind = loop_kind_of(w1.N, epi, out_bytes);
    if (sp.kind < 0) { set_error(name + ": no epilogue variant for this layer shape (net_dim 64, latent 64/128/256 only)"); return DGAN_ERR_UNSUPPORTED; }
    sp.tab = &w2.tab; sp.h_grid = w2.h_grid; sp.w_grid = w2.w_grid; sp.max_acc = w2.max_acc; sp.in_seg = in_seg;
    sp.macs_per_row = macs;
    sp.fwd = (int)c->segs.size() <= (int)c->layers.size();      // Linear + hidden layers + last layer forward
    b.w1 = &w1; b.w2 = &w2; b.epi = epi; b.out_bytes = out_bytes;
    c->segs.push_back(sp); c->binds.push_back(b);
    return 0;
  };
  static const char* lname[] = {"Linear", "Generator.2", "Generator.3", "Generator.5"};
  const std::string fname = c->desc.arch == DGAN_ARCH_CELEBA ? "Generator.6" : "Generator.5";
  int rc;
  double conv_macs = 0;
  for (int l = 0; l < nl; ++l) {
    const GemmLayer& L = c->layers[l];
    const double macs = (double)L.fwd_host.pairs.size() * L.C_in * L.C_out;
    conv_macs += macs;
    SegBind b;
    b.in_kind = l == 0 ? T_ZH : T_ACT; b.in_idx = l - 1; b.out_kind = T_ACT; b.out_idx = l;
    b.bias = L.bias; b.bias_pstride = L.bias_pstride; b.mb_out_layer = L.relu ? l : -1;
    if ((rc = add(std::string(lname[l]) + ".fwd", L.tc_f, L.tc2_f, L.relu ? EPI_BIAS_RELU : EPI_BIAS, 2, l - 1, macs, b))) return rc;
  }
  const double fin_macs = (double)c->macs_per_row - conv_macs;
  {
    SegBind b;
    b.in_kind = T_ACT; b.in_idx = nl - 1; b.out_kind = T_DBLK; b.bias = c->fin.bias;
    if ((rc = add(fname + "+loss.fwd", c->tc_fin.f, c->tc2_fin_f, c->tc_fin.C_out == 1 ? EPI_FINAL_SIGMOID1 : EPI_FINAL_TANH3, 2, nl - 1, fin_macs, b))) return rc;
  }
  c->n_fwd_seg = nl + 1;
  {
    const bool relu = c->layers[nl - 1].relu;
    SegBind b;
    b.in_kind = T_DBLK; b.out_kind = T_DACT; b.out_idx = nl - 1; b.mb_in_layer = relu ? nl - 1 : -1;
    if ((rc = add(fname + ".bwd", c->tc_fin.b, c->tc2_fin_b, relu ? EPI_MASK : EPI_NONE, 2, nl, fin_macs, b))) return rc;
  }
  for (int l = nl - 1; l >= 1; --l) {
    const GemmLayer& L = c->layers[l];
    const bool relu = c->layers[l - 1].relu;
    SegBind b;
    b.in_kind = T_DACT; b.in_idx = l; b.out_kind = T_DACT; b.out_idx = l - 1; b.mb_in_layer = relu ? l - 1 : -1;
    if ((rc = add(std::string(lname[l]) + ".bwd", L.tc_b, L.tc2_b, relu ? EPI_MASK : EPI_NONE, 2, (int)c->segs.size() - 1,
                  (double)L.fwd_host.pairs.size() * L.C_in * L.C_out, b)))
      return rc;
  }
  {
    const GemmLayer& L0 = c->layers[0];
    SegBind b;
    b.in_kind = T_DACT; b.in_idx = 0; b.out_kind = T_GPART;
    if ((rc = add("Linear.bwd+momentum", L0.tc_b, L0.tc2_b, EPI_NONE, 4, (int)c->segs.size() - 1,
                  (double)L0.fwd_host.pairs.size() * L0.C_in * L0.C_out, b)))
      return rc;
  }
  if ((int)c->segs.size() > LOOP_MAX_SEG) { set_error("too many segments"); return DGAN_ERR_UNSUPPORTED; }
  return 0;
}

template <typename T>
static int upload_vec(dgan_ctx* c, const std::vector<T>& v, T** dev) {
  const size_t bytes = std::max<size_t>(1, v.size()) * sizeof(T);
  DGAN_CUDA_CHECK(cudaMalloc((void**)dev, bytes));
  c->allocs.push_back(*dev);
  if (!v.empty()) DGAN_CUDA_CHECK(cudaMemcpy(*dev, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// Plan (or fetch) the L-step schedule for `n_rows` latent rows.  Planning allocates and copies synchronously: it happens
// in dgan_workspace_bytes - which a caller needs before its first dgan_reconstruct of a batch size anyway - so that
// dgan_reconstruct itself never allocates or synchronises (it only falls back to planning here if the caller sized the
// workspace some other way).
static int get_plan(dgan_ctx* c, int n_rows, const DevPlan** out) {
  const int n_pad = (int)align_up((size_t)std::max(n_rows, 1), 2 * kRowTile), n_mpairs = n_pad / (2 * kRowTile);
  for (auto& p : c->plans)
    if (p->n_mpairs == n_mpairs) { *out = p.get(); return 0; }
  std::unique_ptr<DevPlan> dp(new DevPlan());
  dp->n_mpairs = n_mpairs; dp->n_pairs = c->n_pairs;
  int rc;
  if ((rc = loop_plan(c->segs, n_mpairs, c->n_pairs, &dp->host))) return rc;
  const LoopPlan& pl = dp->host;
  for (int v = 0; v < pl.n_seg; ++v)
    if ((rc = upload_vec(c, pl.hdrs[(size_t)v], &dp->items[v]))) return rc;
  for (int r = 0; r < 2; ++r)
    if ((rc = upload_vec(c, pl.tmpl_p[r], &dp->tmpl_p[r]))) return rc;
  if ((rc = upload_vec(c, pl.tmpl_m, &dp->tmpl_m))) return rc;
  if ((rc = upload_vec(c, pl.win_rec_off, &dp->win_rec_off))) return rc;
  if ((rc = upload_vec(c, pl.succ_off, &dp->succ_off))) return rc;
  if ((rc = upload_vec(c, pl.succ, &dp->succ))) return rc;
  if ((rc = upload_vec(c, pl.need, &dp->need))) return rc;
  if ((rc = upload_vec(c, pl.q_init, &dp->q_init))) return rc;
  c->plans.push_back(std::move(dp));
  *out = c->plans.back().get();
  return 0;
}

// Launch parameters for this workspace: per segment the TMA descriptors of its input / weight / output tensors and its
// epilogue bindings.  Encoded once per (workspace, row count); later calls only patch the per-call scalars.
static int build_params(dgan_ctx* c, const Workspace& w, const void* ws_base, const DevPlan* dp) {
  if (c->lp_ws == ws_base && c->lp_rows == w.n_rows && c->lp_plan == dp) return 0;
  LoopParams& P = c->lp;
  P = LoopParams{};
  const LoopPlan& pl = dp->host;
  if (pl.n_item_slots > w.n_counters || pl.q_cap > w.q_cap) { set_error("internal: scheduling state smaller than the plan needs"); return DGAN_ERR_WORKSPACE; }
  int rc;
  auto tensor = [&](int kind, int idx, const void** base, int* chan, int* pix) {
    const int latent = c->desc.latent_dim;
    switch (kind) {
      case T_ZH: *base = w.z_h; *chan = latent; *pix = 1; break;
      case T_ACT: *base = w.act_h[(size_t)idx]; *chan = c->layers[(size_t)idx].C_out; *pix = c->layers[(size_t)idx].P_out; break;
      case T_DACT: *base = w.dact_h[(size_t)idx]; *chan = c->layers[(size_t)idx].C_out; *pix = c->layers[(size_t)idx].P_out; break;
      case T_DBLK: *base = w.dblk; *chan = 64; *pix = c->tc_fin.n_blocks; break;
      default: *base = w.g; *chan = latent; *pix = TC_LINEAR_SPLIT; break;
    }
  };
  for (int v = 0; v < pl.n_seg; ++v) {
    const int s = v;
    const SegBind& b = c->binds[(size_t)s];
    const LoopSegSpec& sp = c->segs[(size_t)s];
    LoopSeg& g = P.seg[v];
    const void* base; int chan, pix;
    tensor(b.in_kind, b.in_idx, &base, &chan, &pix);
    if (chan != sp.K) { set_error("internal: segment input channels"); return DGAN_ERR_INVALID_ARG; }
    if ((rc = tc_make_map(c->tc, &g.tm_a, base, (uint64_t)chan, (uint64_t)w.n_pad, (uint64_t)pix, 128))) return rc;
    g.tm_b = b.w2->tm_b;
    tensor(b.out_kind, b.out_idx, &base, &chan, &pix);
    g.out = const_cast<void*>(base);
    g.tm_out = g.tm_a;                                  // placeholder for the epilogues that do not store by TMA
    if (tc2_tma_epilogue(sp.N, b.epi, b.out_bytes) &&
        (rc = tc_make_map(c->tc, &g.tm_out, base, (uint64_t)chan, (uint64_t)w.n_pad, (uint64_t)pix, 128)))
      return rc;
    g.bias = b.bias;
    g.mb_out = b.mb_out_layer >= 0 ? w.maskbits[(size_t)b.mb_out_layer] : nullptr;
    g.mb_in = b.mb_in_layer >= 0 ? w.maskbits[(size_t)b.mb_in_layer] : nullptr;
    g.items = dp->items[v];
    g.n_tile = (uint32_t)sp.N; g.kind = (uint32_t)sp.kind; g.bias_pstride = (uint32_t)b.bias_pstride;
    g.acc_stride = (uint32_t)tc2_acc_stride(sp.N);
    g.idesc = make_idesc_f16(256, sp.N);
    g.half_b = (uint32_t)(sp.N / 2) * 128u;
    g.win_base = pl.win_base[(size_t)v]; g.item_base = pl.item_base[(size_t)v]; g.n_windows = pl.n_windows[(size_t)v];
  }
  P.tmpl_p[0] = dp->tmpl_p[0]; P.tmpl_p[1] = dp->tmpl_p[1]; P.tmpl_m = dp->tmpl_m;
  P.win_rec_off = dp->win_rec_off; P.succ_off = dp->succ_off; P.succ = dp->succ; P.need = dp->need;
  P.queue = w.queue; P.q_ctl = w.q_ctl; P.depcnt = w.depcnt;
  P.q_cap = pl.q_cap; P.q_shift = 0;
  while ((1u << P.q_shift) < pl.q_cap) ++P.q_shift;
  P.q_init = (uint32_t)pl.q_init.size(); P.n_pairs = (uint32_t)dp->n_pairs;
  P.status = w.status; P.prof = nullptr; P.dbg = nullptr; P.trace = nullptr; P.trace_step = -1;
  P.n_seg = pl.n_seg; P.n_fwd = pl.n_fwd; P.n_pad = w.n_pad; P.n_mpairs = dp->n_mpairs;
  P.y = w.y; P.loss_part = w.loss_part; P.n_rows = w.n_rows; P.nbx = c->tc_fin.nbx; P.w_out = c->tc_fin.w_out;
  P.gscale = c->tc.grad_scale;
  P.mz = w.z; P.mv = w.v; P.mz_h = w.z_h; P.m_nparts = w.n_g_parts; P.m_count = (size_t)w.n_pad * c->desc.latent_dim;
  c->lp_ws = ws_base; c->lp_rows = w.n_rows; c->lp_plan = dp;
  return 0;
}

enum LoopMode : int { LOOP_RECONSTRUCT = 0, LOOP_FORWARD = 1, LOOP_LOSS_GRAD = 2 };

// Enqueue the loop kernel: rec_iters L-steps of (forward, loss, backward-to-z, momentum); the final L-step is forward
// only (the loop returns the pre-update forward of iteration L-1, models/gan.py:419-421) except for
// dgan_loss_grad, which wants the gradient of its one evaluation.  Three small memory operations put the scheduling
// state in its initial condition first: counters zero, queue slots unwritten, the first segment's items of L-step 0 ready.
static int launch_loop(dgan_ctx* c, const Workspace& w, const void* ws_base, const float* x, int R, int B, int rec_iters,
                       float lr, float mu, int decay_lr, int mode, cudaStream_t s) {
  const DevPlan* dp = nullptr;
  int rc;
  if (rec_iters < 1 || rec_iters > 0xFFFF) { set_error("rec_iters must be in [1, 65535] on the fp16 path"); return DGAN_ERR_INVALID_ARG; }
  if ((rc = get_plan(c, w.n_rows, &dp))) return rc;
  if ((rc = build_params(c, w, ws_base, dp))) return rc;
  const LoopPlan& pl = dp->host;
  LoopParams P = c->lp;
  P.x = x; P.R = R; P.B = B;
  P.m_gmul = grad_multiplier(c); P.m_lr = lr; P.m_mu = mu;
  P.m_counter = (mode == LOOP_RECONSTRUCT) ? w.mom_counter : nullptr;
  P.decay_step = decay_lr ? (int)std::ceil(rec_iters * 0.8) : 0;
  P.last_step = rec_iters - 1;
  P.full_last = (mode == LOOP_LOSS_GRAD) ? 1 : 0;
  const unsigned long long total = loop_total_parts(c->segs, pl, rec_iters, mode == LOOP_LOSS_GRAD);
  if (total + (unsigned long long)dp->n_pairs >= 0xFFFFFFFFull) { set_error("too many work items for one launch (batch x rec_rr x rec_iters)"); return DGAN_ERR_UNSUPPORTED; }
  P.n_parts_total = (uint32_t)total;
  DGAN_CUDA_CHECK(cudaMemsetAsync(w.status, 0, (8 + w.n_counters) * sizeof(uint32_t), s));
  DGAN_CUDA_CHECK(cudaMemsetAsync(w.queue, 0xFF, (size_t)pl.q_cap * sizeof(unsigned long long), s));
  DGAN_CUDA_CHECK(cudaMemcpyAsync(w.queue, dp->q_init, pl.q_init.size() * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
  if (w.mom_counter != nullptr) DGAN_CUDA_CHECK(cudaMemsetAsync(w.mom_counter, 0, (size_t)w.n_pad / kRowTile * sizeof(unsigned), s));
  c->last_status = w.status;
  c->loop_passes = (mode == LOOP_LOSS_GRAD) ? 2 * rec_iters : 2 * rec_iters - 1;
  const dim3 grid((unsigned)(2 * dp->n_pairs)), block(LOOP_THREADS);
  if (c->profile == 1) {
    // in-kernel spans: [L-step][segment] {min start, max end} of %globaltimer; per-CTA stall counters; item trace of one L-step
    const size_t need = (size_t)rec_iters * P.n_seg * 2;
    if (need > c->prof_cap) {
      if (c->prof_dev) cudaFree(c->prof_dev);
      c->prof_dev = nullptr; c->prof_cap = 0;
      DGAN_CUDA_CHECK(cudaMalloc((void**)&c->prof_dev, need * sizeof(unsigned long long)));
      c->prof_cap = need;
    }
    std::vector<unsigned long long> init(need);
    for (size_t i = 0; i < need; i += 2) { init[i] = ~0ull; init[i + 1] = 0ull; }
    DGAN_CUDA_CHECK(cudaMemcpyAsync(c->prof_dev, init.data(), need * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
    DGAN_CUDA_CHECK(cudaStreamSynchronize(s));     // `init` dies with this scope (profiling mode only)
    P.prof = c->prof_dev; c->prof_L = rec_iters; c->prof_plan = dp;
    const int n_ctas = 2 * dp->n_pairs;
    if (c->dbg_dev == nullptr || c->dbg_ctas < n_ctas) {
      if (c->dbg_dev) cudaFree(c->dbg_dev);
      c->dbg_dev = nullptr; c->dbg_ctas = 0;
      DGAN_CUDA_CHECK(cudaMalloc((void**)&c->dbg_dev, (size_t)n_ctas * DBG_COUNT * sizeof(unsigned long long)));
      c->dbg_ctas = n_ctas;
    }
    DGAN_CUDA_CHECK(cudaMemsetAsync(c->dbg_dev, 0, (size_t)n_ctas * DBG_COUNT * sizeof(unsigned long long), s));
    P.dbg = c->dbg_dev;
    const size_t n_items = pl.n_item_slots;
    if (c->trace_dev == nullptr || c->trace_items < n_items) {
      if (c->trace_dev) cudaFree(c->trace_dev);
      c->trace_dev = nullptr; c->trace_items = 0;
      DGAN_CUDA_CHECK(cudaMalloc((void**)&c->trace_dev, n_items * 4 * sizeof(unsigned long long)));
      c->trace_items = n_items;
    }
    DGAN_CUDA_CHECK(cudaMemsetAsync(c->trace_dev, 0, n_items * 4 * sizeof(unsigned long long), s));
    P.trace = c->trace_dev; c->trace_plan = dp;
    P.trace_step = std::max(0, rec_iters - 3);         // an L-step in the steady state
    c->trace_step = P.trace_step;
  }
  cudaError_t e;
  {
    ProfScope ps(c, (int)c->kind_names.size() - 1, s);
    if (c->desc.arch == DGAN_ARCH_CELEBA) projection_loop_kernel<DGAN_ARCH_CELEBA><<<grid, block, LOOP_SMEM_BYTES, s>>>(P);
    else projection_loop_kernel<DGAN_ARCH_MNIST><<<grid, block, LOOP_SMEM_BYTES, s>>>(P);
    c->launches++;
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) { set_error(std::string("projection_loop launch: ") + cudaGetErrorString(e)); return DGAN_ERR_CUDA; }
  return 0;
}

}  // namespace dgan

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int dgan_abi_version(void) { return DGAN_ABI_VERSION; }
const char* dgan_last_error(void) { return g_last_error.c_str(); }

int dgan_num_weights(const dgan_desc* d) {
  if (d == nullptr) return DGAN_ERR_INVALID_ARG;
  const int n_deconv = d->arch == DGAN_ARCH_CELEBA ? 4 : 3;
  return 2 + 2 * n_deconv + (d->use_bn ? 6 : 0);
}

static int create_impl(dgan_ctx* c, const dgan_desc* d, const float* const* weights_in, cudaStream_t s) {
  c->desc = *d;
  const bool celeba = d->arch == DGAN_ARCH_CELEBA;
  const int nd = d->net_dim, latent = d->latent_dim;
  c->H = celeba ? 64 : 28; c->W = c->H; c->C = celeba ? 3 : 1;
  c->hwc = c->H * c->W * c->C;
  int rc = 0;
  auto fail = [](int code) { return code; };   // the caller destroys the half-built handle
  // The handle owns copies of every weight tensor: the caller may free or reuse `weights_dev` as soon as the
  // copies enqueued here have run (i.e. after synchronising `stream`).
  std::vector<const float*> wown;
  {
    const std::vector<size_t> counts = weight_counts(d);
    size_t total = 0;
    for (size_t n : counts) total += align_up(n * sizeof(float), 256);
    char* base = nullptr;
    if ((rc = dev_alloc(c, (void**)&base, total))) return fail(rc);
    size_t off = 0;
    for (size_t i = 0; i < counts.size(); ++i) {
      cudaError_t e = cudaMemcpyAsync(base + off, weights_in[i], counts[i] * sizeof(float), cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) { set_error(std::string("weight copy: ") + cudaGetErrorString(e)); return fail(DGAN_ERR_CUDA); }
      wown.push_back((const float*)(base + off));
      off += align_up(counts[i] * sizeof(float), 256);
    }
  }
  const float* const* weights = wown.data();

  // ---- Linear (Generator.Input): [1][N][latent] -> [16][N][4*nd]
  {
    GemmLayer L{};
    L.P_in = 1; L.C_in = latent; L.P_out = 16; L.C_out = 4 * nd; L.h_in = 1; L.w_in = 1; L.h_used = 4; L.w_used = 4;
    L.relu = true;
    L.fwd_host = linear_fwd_pairs(16); L.bwd_host = linear_bwd_pairs(16);
    const float* W = weights[0];             // (latent, 16*4nd), column f = pixel*4nd + c
    L.wf = W; L.wf_tile_stride = L.C_out; L.wf_ld = 16 * L.C_out;
    float* Wt = nullptr;                     // [16*4nd][latent]: backward tile q rows = c, cols = latent
    if ((rc = dev_alloc(c, (void**)&Wt, (size_t)latent * 16 * L.C_out * 4))) return fail(rc);
    const size_t total = (size_t)latent * 16 * L.C_out;
    transpose_tiles_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(W, Wt, latent, 16 * L.C_out, total);
    L.wb = Wt; L.wb_tile_stride = L.C_out * latent; L.wb_ld = latent;
    L.bias = weights[1]; L.bias_pstride = L.C_out;   // bias index f = pixel*C_out + c
    if (d->use_bn) { L.bn_offset = weights[2]; L.bn_scale = weights[3]; L.bn_per_pixel = 1; }   // Generator.BN1, axes [0]
    c->layers.push_back(L);
  }
  // ---- hidden deconvs
  struct DSpec { int c_in, c_out, h_in, h_used; bool relu; int in_raster; };
  std::vector<DSpec> specs;
  if (celeba) specs = {{4 * nd, 2 * nd, 4, 8, true, 4}, {2 * nd, nd, 8, 16, true, 8}, {nd, nd, 16, 32, false, 16}};
  else if (d->use_bn)   // BN2's batch statistics cover all 8x8 outputs of Generator.2; the 7x7 crop comes after BN+ReLU
    specs = {{4 * nd, 2 * nd, 4, 8, true, 4}, {2 * nd, nd, 7, 14, true, 8}};
  else specs = {{4 * nd, 2 * nd, 4, 7, true, 4}, {2 * nd, nd, 7, 14, true, 7}};
  int wi = d->use_bn ? 4 : 2;
  int di = 0;
  for (const DSpec& sp : specs) {
    GemmLayer L{};
    L.P_in = sp.in_raster * sp.in_raster; L.C_in = sp.c_in; L.P_out = sp.h_used * sp.h_used; L.C_out = sp.c_out;
    L.h_in = L.w_in = sp.in_raster; L.h_used = L.w_used = sp.h_used; L.relu = sp.relu;
    L.fwd_host = deconv_fwd_pairs(sp.h_in, sp.h_in, sp.h_used, sp.h_used, sp.in_raster);
    L.bwd_host = deconv_bwd_pairs(sp.h_in, sp.h_in, sp.h_used, sp.h_used, sp.in_raster);
    const float* F = weights[wi];            // (5,5,C_out,C_in)
    float* Ff = nullptr;                     // [25][C_in][C_out]
    const size_t total = (size_t)kTaps * sp.c_out * sp.c_in;
    if ((rc = dev_alloc(c, (void**)&Ff, total * 4))) return fail(rc);
    transpose_tiles_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(F, Ff, sp.c_out, sp.c_in, total);
    L.wf = Ff; L.wf_tile_stride = sp.c_in * sp.c_out; L.wf_ld = sp.c_out;
    L.wb = F;  L.wb_tile_stride = sp.c_in * sp.c_out; L.wb_ld = sp.c_in;
    L.bias = weights[wi + 1];
    wi += 2;
    if (d->use_bn && di < 2) {               // Generator.BN2 / BN3 follow Generator.2 / Generator.3 (axes [0,1,2])
      L.bn_offset = weights[wi]; L.bn_scale = weights[wi + 1]; L.bn_per_pixel = 0;
      wi += 2;
    }
    ++di;
    c->layers.push_back(L);
  }
  // ---- final layer
  {
    FinalLayer& f = c->fin;
    f.h_in = f.w_in = celeba ? 32 : 14; f.C_in = nd; f.C_out = c->C; f.act = celeba ? ACT_TANH : ACT_SIGMOID;
    f.w = weights[wi]; f.bias = weights[wi + 1];
    f.n_bands = (2 * f.h_in + kBandRows - 1) / kBandRows;
    f.fwd_smem = ((size_t)kTaps * f.C_out * f.C_in + (size_t)(kBandRows / 2 + 2) * f.w_in * (f.C_in + 4)) * 4;
    f.bwd_smem = (size_t)kTaps * f.C_out * f.C_in * 4;
  }
  // exact in-bounds MACs per latent row
  c->macs_per_row = 0;
  for (const GemmLayer& L : c->layers) c->macs_per_row += (int64_t)L.fwd_host.pairs.size() * L.C_in * L.C_out;
  {
    PairTable ft = deconv_fwd_pairs(c->fin.h_in, c->fin.w_in, 2 * c->fin.h_in, 2 * c->fin.w_in);
    c->macs_per_row += (int64_t)ft.pairs.size() * c->fin.C_in * c->fin.C_out;
  }
  for (GemmLayer& L : c->layers) {
    if ((rc = upload_table(c, L.fwd_host, &L.fwd, s))) return fail(rc);
    if ((rc = upload_table(c, L.bwd_host, &L.bwd, s))) return fail(rc);
  }
  // opt in to > 48 KB dynamic shared memory where needed
#define OPTIN(K, BYTES) DGAN_CUDA_CHECK(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES)))
  OPTIN((final_fwd_loss_kernel<float, 1, ACT_SIGMOID>), 100 * 1024);
  OPTIN((final_fwd_loss_kernel<float, 3, ACT_TANH>), 100 * 1024);
  OPTIN((final_fwd_loss_kernel<__half, 1, ACT_SIGMOID>), 100 * 1024);
  OPTIN((final_fwd_loss_kernel<__half, 3, ACT_TANH>), 100 * 1024);
#undef OPTIN
  if (d->precision == DGAN_PREC_FP16) {
    std::vector<TcLayerSpec> tspecs;
    for (GemmLayer& L : c->layers) {
      TcLayerSpec t{};
      t.P_in = L.P_in; t.C_in = L.C_in; t.P_out = L.P_out; t.C_out = L.C_out;
      t.h_in = L.h_in; t.w_in = L.w_in; t.h_used = L.h_used; t.w_used = L.w_used;
      t.fwd = &L.fwd_host; t.bwd = &L.bwd_host;
      t.w_fwd_kmajor_src = (&L == &c->layers[0]) ? nullptr : L.wb;  // F[t][co][ci]: rows co (N), cols ci (K)
      t.w_bwd_kmajor_src = (&L == &c->layers[0]) ? nullptr : L.wf;  // Ff[t][ci][co]: rows ci (N), cols co (K)
      t.linear_W = (&L == &c->layers[0]) ? weights[0] : nullptr;
      t.linear_Wt = (&L == &c->layers[0]) ? L.wb : nullptr;
      t.out_f = &L.tc_f; t.out_b = &L.tc_b; t.bias_pstride = L.bias_pstride;
      tspecs.push_back(t);
    }
    if ((rc = tc_build(c->tc, tspecs, latent, &c->allocs, s))) return fail(rc);
    if ((rc = tc_build_final(c->tc, &c->tc_fin, c->fin.w, c->fin.h_in, c->fin.w_in, c->fin.C_in, c->fin.C_out,
                             c->fin.act, &c->allocs, s)))
      return fail(rc);
    for (size_t l = 0; l < c->layers.size(); ++l) {
      GemmLayer& L = c->layers[l];
      if ((rc = tc2_build_direction(c->tc, L.tc_f, &L.tc2_f, L.fwd_host, L.h_used, L.w_used, 0))) return fail(rc);
      if (l == 0) {
        const PairTable split = linear_split_pairs(L.P_out);
        if ((rc = tc2_build_direction(c->tc, L.tc_b, &L.tc2_b, split, 1, TC_LINEAR_SPLIT, 1))) return fail(rc);
      } else if ((rc = tc2_build_direction(c->tc, L.tc_b, &L.tc2_b, L.bwd_host, L.h_in, L.w_in, 0))) {
        return fail(rc);
      }
    }
    const PairTable ft = final_block_fwd_pairs(c->fin.h_in, c->fin.w_in), bt = final_block_bwd_pairs(c->fin.h_in, c->fin.w_in);
    if ((rc = tc2_build_direction(c->tc, c->tc_fin.f, &c->tc2_fin_f, ft, c->fin.h_in / 2, c->fin.w_in / 2, 0))) return fail(rc);
    if ((rc = tc2_build_direction(c->tc, c->tc_fin.b, &c->tc2_fin_b, bt, c->fin.h_in, c->fin.w_in, 0))) return fail(rc);
    if ((rc = build_segments(c))) return fail(rc);
    // the loop kernel: opt in to its shared memory and size the grid to the clusters that can be co-resident (its CTA
    // pairs wait on each other's flags, so every one of them must be on an SM)
    DGAN_CUDA_CHECK(cudaFuncSetAttribute(projection_loop_kernel<DGAN_ARCH_MNIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, LOOP_SMEM_BYTES));
    DGAN_CUDA_CHECK(cudaFuncSetAttribute(projection_loop_kernel<DGAN_ARCH_CELEBA>, cudaFuncAttributeMaxDynamicSharedMemorySize, LOOP_SMEM_BYTES));
    {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(2 * (c->tc.num_sms / 2))); cfg.blockDim = dim3(LOOP_THREADS); cfg.dynamicSmemBytes = LOOP_SMEM_BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n_clusters = 0;
      cudaError_t qe = celeba ? cudaOccupancyMaxActiveClusters(&n_clusters, projection_loop_kernel<DGAN_ARCH_CELEBA>, &cfg)
                              : cudaOccupancyMaxActiveClusters(&n_clusters, projection_loop_kernel<DGAN_ARCH_MNIST>, &cfg);
      if (qe != cudaSuccess) { (void)cudaGetLastError(); n_clusters = c->tc.num_sms / 2; }
      c->n_pairs = std::max(1, std::min(n_clusters, c->tc.num_sms / 2));
    }
    for (const LoopSegSpec& sp : c->segs) { c->kind_names.push_back(sp.name); c->kind_macs_per_row.push_back(sp.macs_per_row); }
    c->kind_names.push_back("projection_loop (all segments, all L-steps, one launch)");
    c->kind_macs_per_row.push_back((double)c->macs_per_row);
  } else {
    static const char* lname[] = {"Linear", "Generator.2", "Generator.3", "Generator.5"};
    for (size_t l = 0; l < c->layers.size(); ++l) {
      const double macs = (double)c->layers[l].fwd_host.pairs.size() * c->layers[l].C_in * c->layers[l].C_out;
      c->kind_names.push_back(std::string(lname[l]) + ".fwd"); c->kind_macs_per_row.push_back(macs);
      c->kind_names.push_back(std::string(lname[l]) + ".bwd"); c->kind_macs_per_row.push_back(macs);
    }
    double fmacs = 0;
    for (const GemmLayer& L : c->layers) fmacs += (double)L.fwd_host.pairs.size() * L.C_in * L.C_out;
    fmacs = (double)c->macs_per_row - fmacs;
    const std::string fn = celeba ? "Generator.6" : "Generator.5";
    c->kind_names.push_back(fn + "+loss.fwd"); c->kind_macs_per_row.push_back(fmacs);
    c->kind_names.push_back(fn + ".bwd"); c->kind_macs_per_row.push_back(fmacs);
    c->kind_names.push_back("momentum"); c->kind_macs_per_row.push_back(0.0);
  }
  DGAN_CUDA_CHECK(cudaGetLastError());
  return DGAN_OK;
}

int dgan_create(dgan_handle* out, const dgan_desc* d, const float* const* weights, int n_weights, void* stream) {
  if (out == nullptr || d == nullptr || weights == nullptr) { set_error("NULL argument"); return DGAN_ERR_INVALID_ARG; }
  *out = nullptr;
  if (d->abi_version != DGAN_ABI_VERSION) { set_error("ABI version mismatch"); return DGAN_ERR_INVALID_ARG; }
  if (d->arch != DGAN_ARCH_MNIST && d->arch != DGAN_ARCH_CELEBA) { set_error("unknown arch"); return DGAN_ERR_INVALID_ARG; }
  if (d->precision != DGAN_PREC_FP32 && d->precision != DGAN_PREC_FP16) { set_error("unknown precision"); return DGAN_ERR_INVALID_ARG; }
  if (d->use_bn && d->precision != DGAN_PREC_FP32) {
    set_error("use_bn=True (batch-statistics BatchNorm, tflib/ops/batchnorm.py:80-93) is built for precision fp32 only");
    return DGAN_ERR_UNSUPPORTED;
  }
  if (d->net_dim <= 0 || d->net_dim % 64 != 0) { set_error("net_dim must be a positive multiple of 64"); return DGAN_ERR_UNSUPPORTED; }
  if (d->latent_dim <= 0 || d->latent_dim % 64 != 0) { set_error("latent_dim must be a positive multiple of 64"); return DGAN_ERR_UNSUPPORTED; }
  if (n_weights != dgan_num_weights(d)) { set_error("wrong number of weight tensors"); return DGAN_ERR_INVALID_ARG; }
  for (int i = 0; i < n_weights; ++i)
    if (weights[i] == nullptr) { set_error("NULL weight pointer"); return DGAN_ERR_INVALID_ARG; }
  int dev_major = 0, dev = 0;
  DGAN_CUDA_CHECK(cudaGetDevice(&dev));
  DGAN_CUDA_CHECK(cudaDeviceGetAttribute(&dev_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (dev_major != 10) { set_error("defensegan_b200 requires an sm_100 (B200) device"); return DGAN_ERR_UNSUPPORTED; }

  dgan_ctx* c = new (std::nothrow) dgan_ctx();
  if (c == nullptr) { set_error("out of host memory"); return DGAN_ERR_INVALID_ARG; }
  const int rc = create_impl(c, d, weights, (cudaStream_t)stream);
  if (rc != DGAN_OK) { dgan_destroy(c); return rc; }   // every failure path frees device memory, streams and events
  *out = c;
  return DGAN_OK;
}

int dgan_destroy(dgan_handle h) {
  if (h == nullptr) return DGAN_OK;
  for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  if (h->prof_dev) cudaFree(h->prof_dev);
  if (h->dbg_dev) cudaFree(h->dbg_dev);
  if (h->trace_dev) cudaFree(h->trace_dev);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return DGAN_OK;
}

size_t dgan_workspace_bytes(dgan_handle h, int batch, int rec_rr) {
  if (h == nullptr || batch <= 0 || rec_rr <= 0) return 0;
  if (h->desc.precision == DGAN_PREC_FP16) {
    // plan (and upload) the loop kernel's schedule for this row count now, so that dgan_reconstruct never has to
    const DevPlan* dp = nullptr;
    if (get_plan(h, batch * rec_rr, &dp) != 0) return 0;
  }
  return carve(h, batch * rec_rr, nullptr).bytes;
}

int64_t dgan_last_launch_count(dgan_handle h) { return h ? h->last_launches : 0; }
int64_t dgan_macs_per_row(dgan_handle h) { return h ? h->macs_per_row : 0; }

int dgan_last_status(dgan_handle h, int* status_out) {
  if (h == nullptr || status_out == nullptr) return DGAN_ERR_INVALID_ARG;
  *status_out = 0;
  if (h->last_status == nullptr) return DGAN_OK;
  uint32_t v = 0;
  DGAN_CUDA_CHECK(cudaMemcpy(&v, h->last_status, sizeof(v), cudaMemcpyDeviceToHost));   // synchronises with the device
  *status_out = (int)v;
  if (v != 0) set_error("projection_loop: a dependency wait timed out (status " + std::to_string(v) + "); results of that call are invalid");
  return DGAN_OK;
}

int dgan_forward(dgan_handle h, const float* z_dev, int n_rows, float* y_dev, void* ws, size_t ws_bytes, void* stream) {
  if (h == nullptr || z_dev == nullptr || y_dev == nullptr || n_rows <= 0) { set_error("invalid argument"); return DGAN_ERR_INVALID_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  Workspace w;
  int rc;
  const int64_t launches0 = h->launches;
  if ((rc = check_ws(h, n_rows, ws, ws_bytes, &w))) return rc;
  if ((rc = run_init_z(h, w, z_dev, 0, s))) return rc;
  if (h->desc.precision == DGAN_PREC_FP16) rc = launch_loop(h, w, ws, nullptr, 1, 1, 1, 0.f, 0.f, 0, LOOP_FORWARD, s);
  else rc = run_forward(h, w, nullptr, 1, 1, false, s);
  if (rc) return rc;
  DGAN_CUDA_CHECK(cudaMemcpyAsync(y_dev, w.y, (size_t)n_rows * h->hwc * 4, cudaMemcpyDeviceToDevice, s));
  h->last_launches = h->launches - launches0;
  return DGAN_OK;
}

int dgan_loss_grad(dgan_handle h, const float* x_dev, int batch, int rec_rr, const float* z_dev, float* y_dev,
                   float* loss_dev, float* grad_dev, void* ws, size_t ws_bytes, void* stream) {
  if (h == nullptr || x_dev == nullptr || z_dev == nullptr || batch <= 0 || rec_rr <= 0) { set_error("invalid argument"); return DGAN_ERR_INVALID_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  const int n_rows = batch * rec_rr;
  Workspace w;
  int rc;
  if ((rc = check_ws(h, n_rows, ws, ws_bytes, &w))) return rc;
  if ((rc = run_init_z(h, w, z_dev, 0, s))) return rc;
  if (h->desc.precision == DGAN_PREC_FP16) {
    if ((rc = launch_loop(h, w, ws, x_dev, rec_rr, batch, 1, 0.f, 0.f, 0, LOOP_LOSS_GRAD, s))) return rc;
  } else {
    if ((rc = run_forward(h, w, x_dev, rec_rr, batch, true, s))) return rc;
    if ((rc = run_backward(h, w, s))) return rc;
  }
  loss_finish_kernel<<<(n_rows + 255) / 256, 256, 0, s>>>(w.loss_part, w.n_loss_parts, w.loss_stride_n, w.loss_stride_b, 1.0f / (float)h->hwc, n_rows, w.loss);
  DGAN_LAUNCH_CHECK(h);
  if (y_dev) DGAN_CUDA_CHECK(cudaMemcpyAsync(y_dev, w.y, (size_t)n_rows * h->hwc * 4, cudaMemcpyDeviceToDevice, s));
  if (loss_dev) DGAN_CUDA_CHECK(cudaMemcpyAsync(loss_dev, w.loss, (size_t)n_rows * 4, cudaMemcpyDeviceToDevice, s));
  if (grad_dev) {
    const size_t n = (size_t)n_rows * h->desc.latent_dim;
    scale_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w.g, w.n_g_parts, (size_t)w.n_pad * h->desc.latent_dim,
                                                                  grad_dev, grad_multiplier(h), n);
    DGAN_LAUNCH_CHECK(h);
  }
  return DGAN_OK;
}

int dgan_sample_z0(dgan_handle h, uint64_t seed, uint64_t z_row_offset, int n_rows, float* z_dev, void* stream) {
  if (h == nullptr || z_dev == nullptr || n_rows <= 0) { set_error("invalid argument"); return DGAN_ERR_INVALID_ARG; }
  const int latent = h->desc.latent_dim;
  const size_t total4 = (size_t)n_rows * latent / 4;
  init_z_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(z_dev, nullptr, nullptr, nullptr, n_rows, n_rows, latent, seed,
                                                                                      sqrtf(1.0f / (float)latent), (size_t)z_row_offset * latent);
  DGAN_LAUNCH_CHECK(h);
  return DGAN_OK;
}

int dgan_reconstruct(dgan_handle h, const dgan_rec_params* prm, const float* x_dev, const float* z0_dev, float* rec_dev,
                     float* loss_dev, int32_t* idx_dev, void* ws, size_t ws_bytes, void* stream) {
  if (h == nullptr || prm == nullptr || x_dev == nullptr || rec_dev == nullptr) { set_error("NULL argument"); return DGAN_ERR_INVALID_ARG; }
  const int batch = prm->batch, rec_rr = prm->rec_rr, rec_iters = prm->rec_iters;
  if (batch <= 0 || rec_rr <= 0 || rec_iters <= 0) { set_error("batch, rec_rr and rec_iters must be positive"); return DGAN_ERR_INVALID_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  const int n_rows = batch * rec_rr;
  Workspace w;
  int rc;
  if ((rc = check_ws(h, n_rows, ws, ws_bytes, &w))) return rc;
  const int64_t launches0 = h->launches;
  h->n_rows_cur = n_rows;
  if ((rc = run_init_z(h, w, z0_dev, prm->seed, s, (size_t)prm->z_row_offset))) return rc;
  if (h->desc.precision == DGAN_PREC_FP16) {
    // the whole L-step loop is one persistent kernel (kernels_loop.cuh)
    if ((rc = launch_loop(h, w, ws, x_dev, rec_rr, batch, rec_iters, prm->rec_lr, prm->momentum, prm->decay_lr, LOOP_RECONSTRUCT, s))) return rc;
  } else {
    const int decay_iter = (int)std::ceil(rec_iters * 0.8);
    const int latent = h->desc.latent_dim;
    for (int t = 0; t < rec_iters; ++t) {
      const bool last = (t == rec_iters - 1);
      float lr = prm->rec_lr;
      if (prm->decay_lr && t >= decay_iter) lr = prm->rec_lr * 0.1f;
      // The loop returns the pre-update forward of iteration L-1 (models/gan.py:419-421):
      // the L-th update is never observed, so its backward pass is not run.
      if ((rc = run_forward(h, w, x_dev, rec_rr, batch, !last, s))) return rc;
      if (last) continue;
      if ((rc = run_backward(h, w, s))) return rc;
      const size_t zcount = (size_t)w.n_pad * latent;
      ProfScope ps(h, 2 * (int)h->layers.size() + 2, s);
      DGAN_CUDA_CHECK(launch_pdl(momentum_kernel, dim3((unsigned)((zcount + 255) / 256)), dim3(256), 0, s, w.z, w.v,
                                 (const float*)w.g, w.n_g_parts, grad_multiplier(h), lr, prm->momentum, zcount, w.z_h));
      DGAN_LAUNCH_CHECK(h);
    }
  }
  loss_finish_kernel<<<(n_rows + 255) / 256, 256, 0, s>>>(w.loss_part, w.n_loss_parts, w.loss_stride_n, w.loss_stride_b, 1.0f / (float)h->hwc, n_rows, w.loss);
  DGAN_LAUNCH_CHECK(h);
  select_kernel<<<batch, 256, 0, s>>>(w.loss, w.y, rec_rr, h->hwc, rec_dev, loss_dev, idx_dev);
  DGAN_LAUNCH_CHECK(h);
  h->last_launches = h->launches - launches0;
  return DGAN_OK;
}

int dgan_profile_enable(dgan_handle h, int enable) {
  if (h == nullptr) return DGAN_ERR_INVALID_ARG;
  for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  h->prof.clear();
  h->profile = enable != 0 ? 1 : 0;
  h->prof_L = 0;
  return DGAN_OK;
}

int dgan_profile_num_kinds(dgan_handle h) { return h ? (int)h->kind_names.size() : 0; }

const char* dgan_profile_kind_name(dgan_handle h, int kind) {
  if (h == nullptr || kind < 0 || kind >= (int)h->kind_names.size()) return "";
  return h->kind_names[kind].c_str();
}

int dgan_profile_read(dgan_handle h, int max_kinds, double* ms_out, int64_t* launches_out, double* flops_per_launch_out) {
  if (h == nullptr || ms_out == nullptr || launches_out == nullptr || flops_per_launch_out == nullptr) return DGAN_ERR_INVALID_ARG;
  const int nk = std::min(max_kinds, (int)h->kind_names.size());
  const bool tc = h->desc.precision == DGAN_PREC_FP16;
  for (int k = 0; k < nk; ++k) {
    ms_out[k] = 0.0; launches_out[k] = 0;
    flops_per_launch_out[k] = 2.0 * h->kind_macs_per_row[k] * (double)h->n_rows_cur;
  }
  for (auto& r : h->prof) {
    DGAN_CUDA_CHECK(cudaEventSynchronize(r.b));
    float ms = 0.f;
    DGAN_CUDA_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
    if (r.kind >= 0 && r.kind < nk) { ms_out[r.kind] += ms; launches_out[r.kind]++; }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  h->prof.clear();
  if (tc && h->profile == 1 && h->prof_L > 0 && h->prof_dev != nullptr && h->prof_plan != nullptr) {
    // the fused launch: its FLOPs are those of `loop_passes` generator passes; its segments are reported as in-kernel
    // spans (first item start .. last item end over all CTA pairs, per L-step; items of neighbouring segments - and of
    // different row pairs, which drift apart - overlap, so the spans add up to more than the launch)
    const LoopPlan& pl = h->prof_plan->host;
    const int n_seg = pl.n_seg, L = h->prof_L;
    if (nk == n_seg + 1) flops_per_launch_out[n_seg] *= 0.5 * (double)h->loop_passes;
    std::vector<unsigned long long> st((size_t)L * n_seg * 2);
    DGAN_CUDA_CHECK(cudaMemcpy(st.data(), h->prof_dev, st.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int t = 0; t < L; ++t)
      for (int sg = 0; sg < n_seg; ++sg) {
        const unsigned long long a = st[((size_t)t * n_seg + sg) * 2], b = st[((size_t)t * n_seg + sg) * 2 + 1];
        if (sg >= nk || a == ~0ull || b <= a) continue;
        ms_out[sg] += (double)(b - a) * 1e-6;
        launches_out[sg]++;
      }
    h->prof_L = 0;
  }
  return DGAN_OK;
}

// Developer aid (not in the public header): per-CTA stall counters (clock64 ticks, LoopDbg order, 16 per CTA) of the most
// recent loop launch made under dgan_profile_enable(h, 1).  Returns the number of CTAs copied.
int dgan_debug_loop_stalls(dgan_handle h, unsigned long long* out, int max_ctas) {
  if (h == nullptr || out == nullptr || h->dbg_dev == nullptr) return 0;
  const int n = std::min(max_ctas, h->dbg_ctas);
  cudaDeviceSynchronize();
  cudaMemcpy(out, h->dbg_dev, (size_t)n * dgan::DBG_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  return n;
}

// Developer aid (not in the public header): the traced L-step of the most recent profiled loop launch.  Per item (segment
// after segment, row pair major, window minor): out[8i..] = {CTA pair, segment, window, row pair, time the item was taken
// from the queue, time its operands' first MMA step could start (accumulator buffer free), epilogue begin, epilogue end}
// (ns of %globaltimer; 0 = not recorded).  deps_out (if not NULL) receives per item up to `max_deps` indices of the items
// it waits for (-1 padded; -2 = the previous L-step's z update).  Returns the number of items.
int dgan_debug_loop_trace(dgan_handle h, unsigned long long* out, long long* deps_out, int max_deps, int max_items) {
  if (h == nullptr || out == nullptr || h->trace_dev == nullptr || h->trace_plan == nullptr) return 0;
  const LoopPlan& pl = h->trace_plan->host;
  const size_t n_all = pl.n_item_slots;
  const int n = (int)std::min<size_t>((size_t)max_items, n_all);
  std::vector<unsigned long long> raw(n_all * 4);
  cudaDeviceSynchronize();
  cudaMemcpy(raw.data(), h->trace_dev, raw.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  // inverse of the successor lists: the windows an item waits for
  std::vector<std::vector<uint32_t>> preds(pl.n_win);
  for (int sg = 0; sg + 1 < pl.n_seg; ++sg)
    for (uint32_t w = 0; w < pl.n_windows[(size_t)sg]; ++w)
      for (uint32_t si = pl.succ_off[pl.win_base[(size_t)sg] + w]; si < pl.succ_off[pl.win_base[(size_t)sg] + w + 1]; ++si)
        preds[pl.win_base[pl.succ[si] >> 16] + (pl.succ[si] & 0xFFFFu)].push_back(w);
  int sg = 0;
  for (int e = 0; e < n; ++e) {
    while (sg + 1 < pl.n_seg && (uint32_t)e >= pl.item_base[(size_t)sg + 1]) ++sg;
    const uint32_t local = (uint32_t)e - pl.item_base[(size_t)sg], mp = local / pl.n_windows[(size_t)sg], win = local % pl.n_windows[(size_t)sg];
    out[(size_t)e * 8 + 0] = raw[(size_t)e * 4 + 0] >> 48; out[(size_t)e * 8 + 1] = (unsigned long long)sg; out[(size_t)e * 8 + 2] = win; out[(size_t)e * 8 + 3] = mp;
    out[(size_t)e * 8 + 4] = raw[(size_t)e * 4 + 0] & 0xFFFFFFFFFFFFull;
    for (int k = 1; k < 4; ++k) out[(size_t)e * 8 + 4 + k] = raw[(size_t)e * 4 + k];
    if (deps_out != nullptr) {
      int k = 0;
      if (sg == 0) deps_out[(size_t)e * max_deps + k++] = -2;
      else
        for (uint32_t u : preds[pl.win_base[(size_t)sg] + win]) {
          if (k >= max_deps) break;
          deps_out[(size_t)e * max_deps + k++] = (long long)(pl.item_base[(size_t)sg - 1] + mp * pl.n_windows[(size_t)sg - 1] + u);
        }
      for (; k < max_deps; ++k) deps_out[(size_t)e * max_deps + k] = -1;
    }
  }
  return n;
}

// Host-only developer/test aid (not in the public header): plan one L-step of the fp16 path for `n_rows` latent rows on
// `n_pairs` CTA pairs exactly as dgan_reconstruct would, and validate the plan with loop_check_plan.  Needs no GPU.
// `mutate` != 0 damages the plan in one specific way first: the check must then fail (self-test of the validator).
// Returns 0, or an error code with the failing check in dgan_last_error().
int dgan_debug_check_plans(const dgan_desc* d, int n_rows, int n_pairs, int mutate) {
  using namespace dgan;
  if (d == nullptr || n_rows <= 0 || n_pairs <= 0) { set_error("invalid argument"); return DGAN_ERR_INVALID_ARG; }
  const bool celeba = d->arch == DGAN_ARCH_CELEBA;
  const int nd = d->net_dim, latent = d->latent_dim;
  const int n_pad = ((n_rows + 2 * kRowTile - 1) / (2 * kRowTile)) * 2 * kRowTile, n_mpairs = n_pad / (2 * kRowTile);
  struct Dir { std::string name; int N, K; PairTable tab; int h, w, force_acc, epi, out_bytes; bool fwd; };
  std::vector<Dir> fwd, bwd;
  fwd.push_back({"Linear.fwd", 4 * nd, latent, linear_fwd_pairs(16), 4, 4, 0, EPI_BIAS_RELU, 2, true});
  struct DSpec { int c_in, c_out, h_in, h_used, in_raster; bool relu; };
  std::vector<DSpec> specs;
  if (celeba) specs = {{4 * nd, 2 * nd, 4, 8, 4, true}, {2 * nd, nd, 8, 16, 8, true}, {nd, nd, 16, 32, 16, false}};
  else specs = {{4 * nd, 2 * nd, 4, 7, 4, true}, {2 * nd, nd, 7, 14, 7, true}};
  int li = 2;
  bool prev_relu = true;     // the Linear's output goes through a ReLU
  for (const DSpec& sp : specs) {
    const std::string nm = "Generator." + std::to_string(li == 4 ? 5 : li);
    fwd.push_back({nm + ".fwd", sp.c_out, sp.c_in, deconv_fwd_pairs(sp.h_in, sp.h_in, sp.h_used, sp.h_used, sp.in_raster), sp.h_used, sp.h_used, 0,
                   sp.relu ? EPI_BIAS_RELU : EPI_BIAS, 2, true});
    bwd.push_back({nm + ".bwd", sp.c_in, sp.c_out, deconv_bwd_pairs(sp.h_in, sp.h_in, sp.h_used, sp.h_used, sp.in_raster), sp.in_raster, sp.in_raster, 0,
                   prev_relu ? EPI_MASK : EPI_NONE, 2, false});
    prev_relu = sp.relu;
    ++li;
  }
  const int fh = celeba ? 32 : 14, c_img = celeba ? 3 : 1;
  fwd.push_back({"last.fwd", 16 * c_img, nd, final_block_fwd_pairs(fh, fh), fh / 2, fh / 2, 0, celeba ? EPI_FINAL_TANH3 : EPI_FINAL_SIGMOID1, 2, true});
  std::vector<Dir> dirs = fwd;
  dirs.push_back({"last.bwd", nd, 64, final_block_bwd_pairs(fh, fh), fh, fh, 0, prev_relu ? EPI_MASK : EPI_NONE, 2, false});
  for (size_t i = bwd.size(); i-- > 0;) dirs.push_back(bwd[i]);
  dirs.push_back({"Linear.bwd", latent, 4 * nd, linear_split_pairs(16), 1, TC_LINEAR_SPLIT, 1, EPI_NONE, 4, false});
  std::vector<LoopSegSpec> segs;
  for (size_t i = 0; i < dirs.size(); ++i) {
    const Dir& dr = dirs[i];
    LoopSegSpec sp;
    sp.name = dr.name; sp.N = dr.N; sp.K = dr.K; sp.kind = loop_kind_of(dr.N, dr.epi, dr.out_bytes);
    if (sp.kind < 0) { set_error(dr.name + ": unsupported layer shape"); return DGAN_ERR_UNSUPPORTED; }
    sp.tab = &dirs[i].tab; sp.h_grid = dr.h; sp.w_grid = dr.w;
    sp.max_acc = tc2_maxb(dr.N);
    if (dr.force_acc > 0) sp.max_acc = std::min(sp.max_acc, dr.force_acc);
    sp.in_seg = (int)i - 1; sp.fwd = dr.fwd;
    segs.push_back(sp);
  }
  LoopPlan plan;
  int rc = loop_plan(segs, n_mpairs, n_pairs, &plan);
  if (rc) return rc;
  if (mutate != 0) {
    // damage a step in the middle of a multi-step window of Generator.3 fwd (segment 2) - or a table entry next to it
    const int sg = 2;
    uint32_t win = 0;
    for (uint32_t w = 0; w < plan.n_windows[sg]; ++w)
      if (plan.win_rec_off[plan.win_base[sg] + w + 1] - plan.win_rec_off[plan.win_base[sg] + w] >= 3) { win = w; break; }
    const uint32_t wi = plan.win_base[sg] + win;
    const size_t r0 = plan.win_rec_off[wi], r1 = plan.win_rec_off[wi + 1];
    if (r1 - r0 < 3) { set_error("plan too small to mutate"); return DGAN_ERR_INVALID_ARG; }
    const size_t k = r0 + 1;
    TcRec& m = plan.tmpl_m[k];
    TcRec* pp[2] = {&plan.tmpl_p[0][k], &plan.tmpl_p[1][k]};
    switch (mutate) {
      case 1: m.w[2] ^= 1u << 10; break;                                  // first-MMA flag of an op
      case 2: m.w[2] ^= 1u << 7; break;                                   // accumulator of an op
      case 3: pp[0]->w[4] ^= 0x01; break;                                 // weight tile staged by rank 0 only
      case 4: pp[0]->w[2] ^= 0x01; pp[1]->w[2] ^= 0x01; break;            // input pixel of an A tile
      case 5: for (int r = 0; r < 2; ++r) pp[r]->w[0] = (pp[r]->w[0] & ~(0xFu << 8)) | ((((pp[r]->w[0] >> 8) & 0xF) ^ 1u) << 8); break;   // k-chunk
      case 6: plan.win_rec_off[wi + 1] -= 1; break;                       // a window loses its last step to its neighbour
      case 7: m.w[0] ^= 1u << 18; break;                                  // MMA warp and producer disagree on the step's size (ring placement)
      case 8: for (int r = 0; r < 2; ++r) pp[r]->w[0] |= 0xBFu; break;    // a run-time field is not blank
      case 9: std::swap(plan.tmpl_m[k], plan.tmpl_m[k + 1]);              // two steps out of order
              for (int r = 0; r < 2; ++r) std::swap(plan.tmpl_p[r][k], plan.tmpl_p[r][k + 1]);
              break;
      case 10: plan.need[wi] += 1; break;                                 // an item waits for one completion too many: never ready
      case 11: plan.succ[plan.succ_off[wi]] ^= 1u; break;                 // an item wakes the wrong window
      case 12: plan.succ_off[wi + 1] -= 1; plan.succ_off[wi] += 0; for (uint32_t j = wi + 1; j < plan.n_win; ++j) { if (j > wi + 1) plan.succ_off[j] -= 1; } plan.succ_off[plan.n_win] -= 1;
               plan.succ.erase(plan.succ.begin() + plan.succ_off[wi + 1]); break;      // a successor is missing
      case 13: plan.q_init.pop_back(); break;                             // a first-segment item is never started
      case 14: plan.q_cap = 64; break;                                    // queue too small for what can be ready at once
      default: break;
    }
  }
  std::string err;
  if ((rc = loop_check_plan(segs, plan, &err))) { set_error(err); return rc; }
  // summary of the plan (read it with dgan_last_error() after a successful call)
  std::string sum = "segments:";
  for (int v = 0; v < plan.n_seg; ++v)
    sum += " [" + segs[(size_t)v].name + " window " +
           std::to_string(plan.shape[4 * v]) + "x" + std::to_string(plan.shape[4 * v + 1]) + " stride " + std::to_string(plan.shape[4 * v + 2]) + "x" +
           std::to_string(plan.shape[4 * v + 3]) + ", " + std::to_string(plan.hdrs[(size_t)v].size()) + " windows, cost " +
           std::to_string((int)(plan.cost_total[(size_t)v] / 1024)) + " KB, largest item " + std::to_string((int)(plan.cost_max[(size_t)v] / 1024)) + " KB]";
  sum += " x " + std::to_string(plan.n_mpairs) + " row pairs; per L-step: " + std::to_string(plan.n_steps) + " steps, " + std::to_string(plan.n_mma) + " MMAs, " +
         std::to_string(2.0 * plan.n_bytes / 1e6) + " MB staged, " + std::to_string((size_t)(plan.win_fwd + plan.win_bwd) * plan.n_mpairs) + " items, " +
         std::to_string(plan.succ.size()) + " graph edges per row pair, queue capacity " + std::to_string(plan.q_cap);
  set_error(sum);
  return 0;
}

}  // extern "C"
