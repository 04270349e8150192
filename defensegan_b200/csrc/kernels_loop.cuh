// The projection loop as ONE persistent kernel (DGAN_PREC_FP16).
//
// Round 1 ran every layer-direction of an L-step as its own persistent tcgen05 kernel: 9 launches per step, 1798 per
// call, each paying ~2 us of launch gap, ~1.5 us until its first operands landed and 2-6 us of un-overlapped last
// epilogue - a quarter of the step.  Here the whole call - L x (generator forward, loss, backward-to-z, momentum) -
// is one launch: 74 CTA pairs (one per TPC, all co-resident) each walk a host-planned stream of
// (segment, window, row pair) items for one L-step and replay it L times.  A "segment" is one layer-direction
// (Linear fwd, Generator.2 fwd, ..., Linear bwd); its items are exactly those of the per-layer kernels (same windows,
// same step records, same canonical accumulation order, hence the same bits).  What used to be a grid-wide kernel
// boundary is now a per-item dependency: the epilogue of an item publishes "window w of row pair mp is written" in
// a global flag word (release), and the TMA producer of a consuming item waits (acquire) for exactly the windows
// whose pixels it stages.  Rows are independent (no BatchNorm on this path), so dependencies never cross row pairs
// and the tail of one layer overlaps the head of the next; the momentum update is applied by the CTA that completes
// a row tile's last Linear-backward partial sum and releases that row pair's next L-step.
//
// Deadlock freedom: every CTA pair processes its items in an order consistent with (L-step, segment), an item only
// waits for items of the previous segment (or the previous L-step's update), no role that signals ever waits on a
// flag, and the grid is sized to the number of co-resident clusters.  The host proves it for each plan by simulating
// the streams (loop_check_plan), and every flag wait has a time-out that raises a status word instead of hanging.
#pragma once
#include "kernels_tc2.cuh"

namespace dgan {

constexpr int LOOP_MAX_SEG = 20;      // virtual segments: (row-pair group) x (layer-direction)
constexpr int LOOP_N_SEC = 4;         // stream sections per CTA pair (see LoopPlan)
// Warp roles, by warpgroup so that registers can be re-balanced with setmaxnreg: warpgroup 0 = TMA producer (warp 0),
// MMA issuer (warp 1) and two idle warps, trimmed to LOOP_REGS_CTRL registers; warpgroups 1-2 = the 8 epilogue warps,
// raised to LOOP_REGS_EPI (the epilogue holds two 32-column TMEM loads in flight while it converts a 64-column unit:
// at the launch-time 168 registers it spilled about a kilobyte per thread).
constexpr int LOOP_THREADS = 128 + 32 * TC2_EPI_WARPS;
constexpr int LOOP_EPI_WARP0 = 4;
constexpr int LOOP_REGS_CTRL = 64, LOOP_REGS_EPI = 216;      // 128*64 + 256*216 <= 384*168 (the launch-time pool)
constexpr int LOOP_EPI_TILES = 2;                                              // one output staging tile per epilogue half
constexpr int LOOP_RING_BYTES = tc2_ring_bytes(64, EPI_BIAS_RELU, 2);           // operand ring next to those tiles
constexpr int LOOP_SMEM_BYTES = LOOP_RING_BYTES + LOOP_EPI_TILES * TC2_TILE_BYTES + TC2_STAGING_BYTES + 1024 + 256;
constexpr uint32_t LOOP_ARRIVALS = 2 * TC2_EPI_WARPS;      // flag increments per item and L-step: 8 epilogue warps x 2 CTAs
constexpr uint32_t LOOP_DEP_PREV = 0x80000000u;             // dependency on the PREVIOUS L-step's value of the flag (z update)

// epilogue variants (N_TILE, epilogue, output type) that occur in the two generators
enum LoopKind : int {
  LK_BR256 = 0, LK_BR128, LK_BR64,        // bias + ReLU (+ 1-bit mask out), fp16 tile via TMA store
  LK_B64,                                  // bias only (CelebA Generator.5)
  LK_MASK64, LK_MASK128, LK_MASK256,       // ReLU-gradient mask in
  LK_NONE64H,                              // plain fp16 (backward into CelebA Generator.5's linear output)
  LK_NONE64F, LK_NONE128F, LK_NONE256F,    // Linear backward: fp32 split-K partial sums (+ momentum tail)
  LK_FINAL16, LK_FINAL48,                  // last layer + sigmoid/tanh + MSE + dL/dpre
  LK_COUNT
};

struct __align__(64) LoopSeg {
  CUtensorMap tm_a, tm_b, tm_out;
  void* out;
  const float* bias;
  unsigned long long* mb_out;
  const unsigned long long* mb_in;
  const TcItem2* items;
  uint32_t n_tile, kind, bias_pstride, acc_stride;
  uint32_t idesc, half_b, phys, group;      // phys: layer-direction index (profiling); group: which row-pair group's L-step counter applies
};

// One entry of the launch's program: run section `sec` of every CTA pair's stream with the given L-step index per row-pair
// group.  flags bit 0: the operand ring must be drained first (the section's ring plan assumes another predecessor).
struct LoopProg { int32_t sec, t0, t1, flags; };

// Entry `pi` of the program of a launch over `n_groups` row-pair groups (see LoopPlan for the sections):
//   two groups:  sec0 [A.fwd(0)],  then for j = 0 .. L-2:  sec1 [A.bwd(j) | B.fwd(j)],  sec2 [A.fwd(j+1) | B.bwd(j)],  then sec3 [B.fwd(L-1)]
//   one group:   sec2 [fwd(0)],    then for j = 0 .. L-2:  sec1 [bwd(j)],  sec2 [fwd(j+1)]      (+ sec1 [bwd(L-1)] for dgan_loss_grad)
__host__ __device__ inline LoopProg loop_prog_entry(int n_groups, int n_prog, int pi) {
  LoopProg e;
  if (n_groups == 2) {
    if (pi == 0) { e.sec = 0; e.t0 = 0; e.t1 = -1; e.flags = 0; return e; }
    if (pi == n_prog - 1) { e.sec = 3; e.t0 = -1; e.t1 = (n_prog - 2) / 2; e.flags = 1; return e; }
    const int j = (pi - 1) >> 1;
    if (pi & 1) { e.sec = 1; e.t0 = j; e.t1 = j; e.flags = (j == 0) ? 1 : 0; }
    else { e.sec = 2; e.t0 = j + 1; e.t1 = j; e.flags = 0; }
    return e;
  }
  if (pi == 0) { e.sec = 2; e.t0 = e.t1 = 0; e.flags = 0; return e; }
  const int j = (pi - 1) >> 1;
  e.sec = (pi & 1) ? 1 : 2;
  e.t0 = e.t1 = (pi & 1) ? j : j + 1;
  e.flags = 0;
  return e;
}
inline int loop_prog_length(int n_groups, int rec_iters, bool full_last) {
  return n_groups == 2 ? 2 * rec_iters : 2 * rec_iters - 1 + (full_last ? 1 : 0);
}

struct LoopParams {
  LoopSeg seg[LOOP_MAX_SEG];       // virtual segments
  const TcRec* stream_p[2];        // producer records per cluster rank: per CTA pair its LOOP_N_SEC sections, back to back
  const TcRec* stream_m;           // MMA records, same indexing
  const uint32_t* stream_off;      // [n_pairs][LOOP_N_SEC + 1] record offsets
  const uint4* eitems;             // items in stream order, all pairs: x = vseg << 16 | window, y = row pair, z = flag index
  const uint32_t* eitem_off;       // [n_pairs][LOOP_N_SEC + 1] item offsets
  const uint32_t* dep_off;         // [items + 1] -> deps
  const uint32_t* deps;            // flag indices (| LOOP_DEP_PREV)
  uint32_t* flags;                 // zeroed per call
  uint32_t* status;                // [0] != 0: a flag wait timed out (results invalid)
  unsigned long long* prof;        // optional [t][n_vseg][2] globaltimer min-start / max-end
  unsigned long long* dbg;         // optional [CTA][16] stall counters of the roles (clock64 ticks), see LoopDbg
  unsigned long long* trace;       // optional [items of the traced program entry][4] globaltimer: dependency wait begin / end, epilogue begin / end
  int trace_entry;                 // the program entries trace_entry and trace_entry + 1 are traced
  int n_prog, n_groups, n_vseg;    // program length (loop_prog_entry), row-pair groups, virtual segments
  int last_step;                   // index of the call's final L-step (rec_iters - 1): its forward writes G(z) and the loss
  int n_pad, n_mpairs;
  // last layer / loss (models/gan.py:411-414)
  const float* x; float* y; float* loss_part;
  int R, B, n_rows, nbx, w_out;
  float gscale;
  // momentum update (models/gan.py:389-391), applied in the tail of the Linear backward
  float* mz; float* mv; __half* mz_h;
  float m_gmul, m_lr, m_mu;
  unsigned* m_counter;             // [n_pad / 128] tickets; NULL = leave the partial sums (dgan_loss_grad)
  int m_nparts, decay_step;        // decay_step > 0: lr x0.1 from that L-step on (opt-in)
  size_t m_count;
  uint32_t zflag_base;
};

namespace ptx {
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
}  // namespace ptx

// Spin until *flag >= target (acquire).  A wait that lasts seconds means a broken plan or a faulted peer: raise the
// status word and fall through (every later wait then falls through as well) instead of hanging the GPU.
__device__ __forceinline__ void loop_wait_flag(const uint32_t* flag, uint32_t target, uint32_t* status) {
  if (ptx::ld_acquire_gpu(flag) >= target) return;
  const unsigned long long t0 = ptx::globaltimer();
  uint32_t spins = 0;
  while (ptx::ld_acquire_gpu(flag) < target) {
    if ((++spins & 255u) == 0u) {
      if (*reinterpret_cast<volatile uint32_t*>(status) != 0u) return;
      if (ptx::globaltimer() - t0 > 4000000000ull) { atomicExch(status, 1u); return; }
    }
  }
}

// "this thread's global writes of the item are done": order them before the flag increment, for readers in both proxies
__device__ __forceinline__ void loop_publish_fence() {
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  ptx::fence_proxy_async_all();
}

// per-CTA stall counters written when LoopParams::dbg != NULL (developer aid: tools/loop_stalls.py)
enum LoopDbg : int { DBG_P_FLAG = 0, DBG_P_RING, DBG_P_TOTAL, DBG_M_FULL, DBG_M_ACC, DBG_M_TOTAL, DBG_E_ACC, DBG_E_TILE, DBG_E_TOTAL,
                     DBG_S_TILE, DBG_S_DONE, DBG_S_TOTAL, DBG_P_SLOW, DBG_COUNT = 16 };

struct LoopCtx {                     // per-thread view of the CTA's pipeline state handed to the epilogue variants
  uint32_t tmem_base, bar_acc_full, bar_acc_empty, epi_base, bar_base;
  int warp, lane, rank;
  uint32_t item_count, tile_count;
  long long t_acc, t_tile;          // stall ticks (debug)
};

// 32 accumulator columns of one row -> (bias | ReLU + mask bits out | mask bits in) -> 16 packed fp16 pairs
template <int EPI>
__device__ __forceinline__ void loop_convert_half(const uint32_t (&r)[32], const float* __restrict__ bias32, uint32_t mask_in,
                                                  uint32_t* __restrict__ pk16, uint32_t& mask_out) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (EPI == EPI_BIAS_RELU || EPI == EPI_BIAS) {
    const float4* bp = reinterpret_cast<const float4*>(bias32);
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 b = __ldg(bp + j4);
      v[j4 * 4 + 0] += b.x; v[j4 * 4 + 1] += b.y; v[j4 * 4 + 2] += b.z; v[j4 * 4 + 3] += b.w;
    }
    if (EPI == EPI_BIAS_RELU) {
      uint32_t bits = 0u;
#pragma unroll
      for (int j = 0; j < 32; ++j) { v[j] = fmaxf(v[j], 0.f); bits |= (uint32_t)(v[j] > 0.f) << j; }
      mask_out = bits;
    }
  }
  if (EPI == EPI_MASK) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (!((mask_in >> j) & 1u)) v[j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) pk16[j] = pack_half2(v[2 * j], v[2 * j + 1]);
}

// ------------------------------------------------------------------------------------------
// One item's epilogue.  Same arithmetic as the per-layer kernels of round 1 (the code below is that epilogue,
// parameterised at run time by the segment); ends by releasing the accumulator buffer and publishing the item.
// ------------------------------------------------------------------------------------------
template <int N_TILE, int EPI, typename TOUT>
__device__ __forceinline__ void loop_epilogue_item(const LoopParams& P, const LoopSeg& sg, LoopCtx& cx, const TcFinalArgs& fa,
                                                   int win, int mp, uint32_t* flag) {
  constexpr bool TMA_EPI = tc2_tma_epilogue(N_TILE, EPI, (int)sizeof(TOUT));
  constexpr int ACC_STRIDE = tc2_acc_stride(N_TILE);
  constexpr bool FINAL = (EPI == EPI_FINAL_SIGMOID1 || EPI == EPI_FINAL_TANH3);
  const int warp = cx.warp, lane = cx.lane, rank = cx.rank;
  const int lq = warp & 3;                          // TMEM lanes this warp may access
  const int half = (warp - LOOP_EPI_WARP0) >> 2;    // 0 | 1: which of the two warps of this quarter
  const int row = lq * 32 + lane;
  const int n_pad = P.n_pad;
  const TcItem2* ip = sg.items + win;
  const int n_acc = (int)ip->n_acc;
  const size_t n = (size_t)(2 * mp + rank) * kRowTile + row;
  const uint32_t buf = cx.item_count & 1;
  const uint32_t tbuf = cx.tmem_base + ((uint32_t)(lq * 32) << 16) + buf * TC2_BUF_COLS;
  TOUT* __restrict__ out = reinterpret_cast<TOUT*>(sg.out);
  const float* __restrict__ bias = sg.bias;
  const int bias_pstride = (int)sg.bias_pstride;

  float4 xq_next[FINAL ? (EPI == EPI_FINAL_SIGMOID1 ? 4 : 12) : 1];
  if (FINAL && half < n_acc)     // first block's target pixels: in flight while the MMAs finish
    tc_final_targets<(EPI == EPI_FINAL_SIGMOID1 ? 1 : 3)>(reinterpret_cast<float4(&)[EPI == EPI_FINAL_SIGMOID1 ? 4 : 12]>(xq_next), fa, ip->q[half], (int)n);
  {
    const long long tw0 = P.dbg ? clock64() : 0;
    ptx::mbar_wait(cx.bar_acc_full + 8 * buf, (cx.item_count >> 1) & 1);
    if (P.dbg) cx.t_acc += clock64() - tw0;
  }
  ptx::tc_fence_after();
  if (FINAL) {
    constexpr int CO = (EPI == EPI_FINAL_SIGMOID1) ? 1 : 3;
    float4 xq[4 * CO];
    for (int a = half; a < n_acc; a += 2) {
#pragma unroll
      for (int j = 0; j < 4 * CO; ++j) xq[j] = xq_next[j];
      if (a + 2 < n_acc) tc_final_targets<CO>(reinterpret_cast<float4(&)[4 * CO]>(xq_next), fa, ip->q[a + 2], (int)n);
      const uint32_t taddr = tbuf + (uint32_t)(a * ACC_STRIDE);
      if (EPI == EPI_FINAL_SIGMOID1)
        tc_final_epilogue<1, ACT_SIGMOID>(taddr, fa, bias, ip->q[a], (int)n, n_pad, reinterpret_cast<__half*>(out),
                                          reinterpret_cast<const float4(&)[4]>(xq));
      else
        tc_final_epilogue<3, ACT_TANH>(taddr, fa, bias, ip->q[a], (int)n, n_pad, reinterpret_cast<__half*>(out),
                                       reinterpret_cast<const float4(&)[12]>(xq));
    }
  } else if (TMA_EPI) {
    // ---- 64-column units through shared memory: TMEM -> regs -> (bias|ReLU|mask) -> fp16 ->
    //      128B-swizzled smem tile -> one TMA store per 128x64 tile.
    constexpr int G = N_TILE / 64;                    // 64-column groups per accumulator
    const int n_units = n_acc * G;
    // The staging tile of this epilogue half is handed to the half's STORE WARP (warp 2 + half): it issues the TMA store,
    // frees the tile when the store has read it, and publishes the item when the item's stores have completed - so no
    // thread that does arithmetic ever waits for global-memory latency.
    const uint32_t tile_full = cx.bar_base + 168 + 8 * (uint32_t)half, tile_free = cx.bar_base + 184 + 8 * (uint32_t)half;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t s_out = cx.epi_base + (uint32_t)half * TC2_TILE_BYTES;
    uint32_t r0[32], r1[32];
    unsigned long long mbits = ~0ull, mbits_next = ~0ull;
    if (half < n_units) {
      const int a = half / G, g = half % G;
      if (EPI == EPI_MASK) mbits_next = __ldcg(sg.mb_in + ((size_t)ip->q[a] * n_pad + n) * G + g);
      ptx::tmem_ld32(tbuf + (uint32_t)(a * ACC_STRIDE + g * 64), r0);
      ptx::tmem_ld32(tbuf + (uint32_t)(a * ACC_STRIDE + g * 64 + 32), r1);
    }
    for (int u = half; u < n_units; u += 2) {
      const int a = u / G, g = u % G, q = ip->q[a];
      mbits = mbits_next;
      ptx::tmem_ld_wait();
      uint32_t pk[32];
      uint32_t mlo = 0u, mhi = 0u;
      // two 32-column halves one after the other: only 32 fp32 values are live at a time
      loop_convert_half<EPI>(r0, bias + (size_t)q * bias_pstride + g * 64, (uint32_t)mbits, &pk[0], mlo);
      loop_convert_half<EPI>(r1, bias + (size_t)q * bias_pstride + g * 64 + 32, (uint32_t)(mbits >> 32), &pk[16], mhi);
      if (EPI == EPI_BIAS_RELU && sg.mb_out != nullptr)
        sg.mb_out[((size_t)q * n_pad + n) * G + g] = ((unsigned long long)mhi << 32) | mlo;
      if (u + 2 < n_units) {                           // next unit's accumulator columns: in flight during the store phase
        const int a2 = (u + 2) / G, g2 = (u + 2) % G;
        if (EPI == EPI_MASK) mbits_next = __ldcg(sg.mb_in + ((size_t)ip->q[a2] * n_pad + n) * G + g2);
        ptx::tmem_ld32(tbuf + (uint32_t)(a2 * ACC_STRIDE + g2 * 64), r0);
        ptx::tmem_ld32(tbuf + (uint32_t)(a2 * ACC_STRIDE + g2 * 64 + 32), r1);
      }
      const long long tw0 = P.dbg ? clock64() : 0;
      ptx::mbar_wait(tile_free, (cx.tile_count & 1) ^ 1);          // the store that last read s_out is done (first use: free)
      if (P.dbg) cx.t_tile += clock64() - tw0;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        ptx::st_shared_v4(s_out + (uint32_t)row * 128u + (((uint32_t)c ^ swz) << 4), pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(tile_full);                                  // 128 arrivals: tile (and this unit's mask words) complete
      ++cx.tile_count;
    }
  } else {
    constexpr int CH = N_TILE >= 32 ? N_TILE / 32 : 1;     // 32-column chunks per accumulator
    const int n_units = n_acc * CH;
    uint32_t rA[32], rB[32];
    int u = half;
    if (u < n_units) ptx::tmem_ld32(tbuf + (uint32_t)((u / CH) * ACC_STRIDE + (u % CH) * 32), rA);
    for (; u < n_units; u += 4) {
      ptx::tmem_ld_wait();
      if (u + 2 < n_units) ptx::tmem_ld32(tbuf + (uint32_t)(((u + 2) / CH) * ACC_STRIDE + ((u + 2) % CH) * 32), rB);
      tc_store_chunk<N_TILE, EPI, TOUT>(rA, ip->q[u / CH], (u % CH) * 32, n, n_pad, out, bias, bias_pstride);
      if (u + 2 < n_units) {
        ptx::tmem_ld_wait();
        if (u + 4 < n_units) ptx::tmem_ld32(tbuf + (uint32_t)(((u + 4) / CH) * ACC_STRIDE + ((u + 4) % CH) * 32), rA);
        tc_store_chunk<N_TILE, EPI, TOUT>(rB, ip->q[(u + 2) / CH], ((u + 2) % CH) * 32, n, n_pad, out, bias, bias_pstride);
      }
    }
  }
  // ---- hand the accumulator buffer back to the MMA warp
  ptx::tc_fence_before();
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive_remote(cx.bar_acc_empty + 8 * buf, 0);

  // ---- publish the item
  if (TMA_EPI) {
    // published by the half's store warp once the item's tile stores have completed
  } else if (EPI == EPI_NONE && sizeof(TOUT) == 4 && P.m_counter != nullptr) {
    // ---- momentum in the tail of the split-K Linear backward (tf.train.MomentumOptimizer, models/gan.py:389-391).
    //      Every epilogue thread has stored its share of this item's partial sums; the CTA that completes the last
    //      partial of its 128-row tile applies v <- mu v + g, z <- z - lr v (parts summed in the fixed order 0, 1, 2, ...)
    //      and releases the row pair's next L-step.
    const uint32_t flag_addr = cx.bar_base + 200;
    const unsigned rt = 2u * (unsigned)mp + (unsigned)rank;
    __threadfence();
    ptx::named_bar_sync(3, 32 * TC2_EPI_WARPS);
    if (warp == LOOP_EPI_WARP0 && lane == 0) {
      const unsigned ticket = atomicAdd(P.m_counter + rt, 1u);
      ptx::st_shared_u32(flag_addr, ticket == (unsigned)P.m_nparts - 1u ? 1u : 0u);
    }
    ptx::named_bar_sync(3, 32 * TC2_EPI_WARPS);
    if (ptx::ld_shared_u32(flag_addr) != 0u) {
      __threadfence();
      const float* __restrict__ gp = reinterpret_cast<const float*>(out);
      const size_t base = (size_t)rt * kRowTile * N_TILE;
      const int tid = (warp - LOOP_EPI_WARP0) * 32 + lane;
      // latency-bound (every operand is an L2 read): all partial sums, v and z of UNR positions are requested before any
      // is used - one round trip per iteration instead of one per partial sum
      constexpr int STRIDE = 4 * 32 * TC2_EPI_WARPS, UNR = 4, MAXP = TC_LINEAR_SPLIT;
      for (int e0 = tid * 4; e0 < kRowTile * N_TILE; e0 += UNR * STRIDE) {
        float4 gs[MAXP][UNR], vv[UNR], zz[UNR];
#pragma unroll
        for (int pp = 0; pp < MAXP; ++pp)
#pragma unroll
          for (int k = 0; k < UNR; ++k)
            gs[pp][k] = (pp < P.m_nparts) ? __ldcg(reinterpret_cast<const float4*>(gp + base + (size_t)(e0 + k * STRIDE) + (size_t)pp * P.m_count))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
          const size_t i = base + (size_t)(e0 + k * STRIDE);
          vv[k] = __ldcg(reinterpret_cast<const float4*>(P.mv + i));
          zz[k] = __ldcg(reinterpret_cast<const float4*>(P.mz + i));
        }
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
          const size_t i = base + (size_t)(e0 + k * STRIDE);
          float4 g4 = gs[0][k];
#pragma unroll
          for (int pp = 1; pp < MAXP; ++pp)          // fixed order: parts 0, 1, 2, ... (absent parts add +0)
            if (pp < P.m_nparts) { g4.x += gs[pp][k].x; g4.y += gs[pp][k].y; g4.z += gs[pp][k].z; g4.w += gs[pp][k].w; }
          float4 v4 = vv[k], z4 = zz[k];
          v4.x = fmaf(fa.m_mu, v4.x, fa.m_gmul * g4.x); v4.y = fmaf(fa.m_mu, v4.y, fa.m_gmul * g4.y);
          v4.z = fmaf(fa.m_mu, v4.z, fa.m_gmul * g4.z); v4.w = fmaf(fa.m_mu, v4.w, fa.m_gmul * g4.w);
          z4.x -= fa.m_lr * v4.x; z4.y -= fa.m_lr * v4.y; z4.z -= fa.m_lr * v4.z; z4.w -= fa.m_lr * v4.w;
          *reinterpret_cast<float4*>(P.mv + i) = v4;
          *reinterpret_cast<float4*>(P.mz + i) = z4;
          *reinterpret_cast<uint2*>(P.mz_h + i) = make_uint2(pack_half2(z4.x, z4.y), pack_half2(z4.z, z4.w));
        }
      }
      loop_publish_fence();
      ptx::named_bar_sync(3, 32 * TC2_EPI_WARPS);
      if (warp == LOOP_EPI_WARP0 && lane == 0) {
        P.m_counter[rt] = 0u;                            // ready for the next L-step's tickets
        __threadfence();
        ptx::red_release_gpu_add(P.flags + P.zflag_base + mp, LOOP_ARRIVALS / 2);   // this 128-row tile's half of z[mp]
      }
    }
    // (nothing waits on the partial sums themselves except through the ticket)
  } else {
    loop_publish_fence();
    __syncwarp();
    if (lane == 0) ptx::red_release_gpu_add(flag, 1u);
  }
}

template <int ARCH>
__device__ __forceinline__ void loop_epilogue_dispatch(const LoopParams& P, const LoopSeg& sg, LoopCtx& cx, const TcFinalArgs& fa,
                                                       int win, int mp, uint32_t* flag) {
  switch (sg.kind) {
    case LK_BR256: loop_epilogue_item<256, EPI_BIAS_RELU, __half>(P, sg, cx, fa, win, mp, flag); break;
    case LK_BR128: loop_epilogue_item<128, EPI_BIAS_RELU, __half>(P, sg, cx, fa, win, mp, flag); break;
    case LK_BR64: loop_epilogue_item<64, EPI_BIAS_RELU, __half>(P, sg, cx, fa, win, mp, flag); break;
    case LK_MASK64: loop_epilogue_item<64, EPI_MASK, __half>(P, sg, cx, fa, win, mp, flag); break;
    case LK_MASK128: loop_epilogue_item<128, EPI_MASK, __half>(P, sg, cx, fa, win, mp, flag); break;
    case LK_MASK256: loop_epilogue_item<256, EPI_MASK, __half>(P, sg, cx, fa, win, mp, flag); break;
    case LK_NONE64F: loop_epilogue_item<64, EPI_NONE, float>(P, sg, cx, fa, win, mp, flag); break;
    case LK_NONE128F: loop_epilogue_item<128, EPI_NONE, float>(P, sg, cx, fa, win, mp, flag); break;
    case LK_NONE256F: loop_epilogue_item<256, EPI_NONE, float>(P, sg, cx, fa, win, mp, flag); break;
    case LK_B64: if (ARCH == DGAN_ARCH_CELEBA) loop_epilogue_item<64, EPI_BIAS, __half>(P, sg, cx, fa, win, mp, flag); break;
    case LK_NONE64H: if (ARCH == DGAN_ARCH_CELEBA) loop_epilogue_item<64, EPI_NONE, __half>(P, sg, cx, fa, win, mp, flag); break;
    case LK_FINAL48: if (ARCH == DGAN_ARCH_CELEBA) loop_epilogue_item<48, EPI_FINAL_TANH3, __half>(P, sg, cx, fa, win, mp, flag); break;
    case LK_FINAL16: if (ARCH == DGAN_ARCH_MNIST) loop_epilogue_item<16, EPI_FINAL_SIGMOID1, __half>(P, sg, cx, fa, win, mp, flag); break;
    default: break;
  }
}

template <int ARCH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LOOP_THREADS, 1)
projection_loop_kernel(const __grid_constant__ LoopParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + LOOP_RING_BYTES;                       // two output staging tiles
  const uint32_t stg_base = epi_base + LOOP_EPI_TILES * TC2_TILE_BYTES;        // [producer ring][MMA ring] of TcRec
  const uint32_t bar_base = stg_base + TC2_STAGING_BYTES;
  // full[s] @ +8s (s<8), empty[s] @ +64+8s, acc_full[2] @ +128, acc_empty[2] @ +144, tmem slot @ +160,
  // tile_full[2] @ +168, tile_free[2] @ +184 (epilogue half <-> store warp), momentum-tail flag @ +200
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 64, bar_acc_full = bar_base + 128, bar_acc_empty = bar_base + 144;
  const uint32_t tmem_slot = bar_base + 160;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < TC2_NSLOT; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);    // leader's producer arrive.expect_tx (bytes of both CTAs)
      ptx::mbar_init(bar_empty + 8 * s, 1);   // one multicast commit per CTA
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_acc_full + 8 * b, 1);
      ptx::mbar_init(bar_acc_empty + 8 * b, 2 * TC2_EPI_WARPS);   // epilogue warps of both CTAs (used on the leader only)
      ptx::mbar_init(bar_base + 168 + 8 * b, 128);                // tile_full[half]: the 4 warps of an epilogue half
      ptx::mbar_init(bar_base + 184 + 8 * b, 1);                  // tile_free[half]: the half's store warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(tmem_slot, 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();                     // barriers of BOTH CTAs initialised before any remote signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const uint32_t* __restrict__ soff = P.stream_off + (size_t)pair * (LOOP_N_SEC + 1);
  const uint32_t* __restrict__ eoff = P.eitem_off + (size_t)pair * (LOOP_N_SEC + 1);

  if (warp < LOOP_EPI_WARP0) {
   ptx::setmaxnreg_dec<LOOP_REGS_CTRL>();
   if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const TcRec* __restrict__ stream = rank ? P.stream_p[1] : P.stream_p[0];
    const uint32_t ring = stg_base;
    uint32_t it = 0;                                      // steps issued so far, over all replays (barrier slot / phase)
    long long t_flag = 0, t_ring = 0, n_slow = 0;
    const long long t_p0 = P.dbg ? clock64() : 0;
    for (int pi = 0; pi < P.n_prog; ++pi) {
      const LoopProg pe = loop_prog_entry(P.n_groups, P.n_prog, pi);
      const uint32_t rbeg = __ldg(soff + pe.sec), rend = __ldg(soff + pe.sec + 1);
      if (rbeg >= rend) continue;
      if (pe.flags & 1) {
        // the section's ring plan assumes an empty ring: wait until every step issued so far has been consumed
        for (uint32_t j = it > TC2_NSLOT ? it - TC2_NSLOT : 0; j < it; ++j) ptx::mbar_wait(bar_empty + 8 * (j & (TC2_NSLOT - 1)), (j >> 3) & 1);
      }
      // dependencies of the section's first item; later items are described one item ahead by the records themselves
      // (w[6], w[7] of an item's first step = dependency range of the NEXT item)
      uint32_t first_d0 = 0, first_cnt = 0;
      {
        const uint32_t i0 = __ldg(eoff + pe.sec), i1 = __ldg(eoff + pe.sec + 1);
        if (i0 < i1) { first_d0 = __ldg(P.dep_off + i0); first_cnt = __ldg(P.dep_off + i0 + 1) - first_d0; }
      }
      uint4 mine = make_uint4(0, 0, 0, 0);
      if (2 * rbeg + lane < 2 * rend) mine = __ldg(reinterpret_cast<const uint4*>(stream + rbeg) + lane);
      // the item about to start: dependency range, this lane's entry (first 32) and its flag value if already fetched
      uint32_t cur_d0 = first_d0, cur_cnt = first_cnt, cur_e = 0, cur_f = 0;
      bool cur_f_valid = false;
      if (lane < cur_cnt) cur_e = __ldg(P.deps + cur_d0 + lane);
      uint32_t nxt_d0 = 0, nxt_cnt = 0, nxt_e = 0;
      uint32_t trace_item = __ldg(eoff + pe.sec);               // index of the item about to start (trace only)
      for (uint32_t base = rbeg; base < rend; base += TC2_REC_BATCH) {
        ptx::st_shared_v4(ring + lane * 16u, mine.x, mine.y, mine.z, mine.w);
        __syncwarp();
        if (2 * (base + TC2_REC_BATCH) + lane < 2 * rend) mine = __ldg(reinterpret_cast<const uint4*>(stream + base + TC2_REC_BATCH) + lane);
        const uint32_t cnt = min((uint32_t)TC2_REC_BATCH, rend - base);
        for (uint32_t i = 0; i < cnt; ++i, ++it) {
          const uint4 r0 = ptx::ld_shared_v4(ring + i * 32u);
          const uint4 r1 = ptx::ld_shared_v4(ring + i * 32u + 16u);
          const uint32_t slot = it & (TC2_NSLOT - 1);
          const int kc = (r0.x >> 8) & 0xF, nA = (r0.x >> 12) & 0x7, nB = (r0.x >> 15) & 0xF;
          const uint32_t dep = (r0.x >> 19) & 0xF;
          const int seg = (int)((r0.y >> 16) & 0x1Fu);
          const LoopSeg& sg = P.seg[seg];
          if ((r0.y >> 21) & 1u) {
            // first step of an item: everything it stages must have been published.  Fast path: the flag values were
            // fetched while the previous item's last step was issued and already satisfy the target.
            const long long tw0 = P.dbg ? clock64() : 0;
            const int t = sg.group ? pe.t1 : pe.t0;                       // this item's L-step
            const bool tracing = P.trace != nullptr && (pi == P.trace_entry || pi == P.trace_entry + 1) && leader && lane == 0;
            if (tracing) P.trace[(size_t)trace_item * 4 + 0] = ptx::globaltimer();
            if (cur_cnt > 0) {
              const uint32_t target = LOOP_ARRIVALS * (uint32_t)((cur_e & LOOP_DEP_PREV) ? t : t + 1);
              const bool ok = (lane >= cur_cnt) || (cur_f_valid && cur_f >= target);
              if (!__all_sync(0xffffffffu, ok) || cur_cnt > 32u) {
                ++n_slow;
                for (uint32_t d = cur_d0 + lane; d < cur_d0 + cur_cnt; d += 32) {
                  const uint32_t e = __ldg(P.deps + d);
                  loop_wait_flag(P.flags + (e & ~LOOP_DEP_PREV), LOOP_ARRIVALS * (uint32_t)((e & LOOP_DEP_PREV) ? t : t + 1), P.status);
                }
                __syncwarp();
              }
              ptx::fence_proxy_async_all();             // acquired generic-proxy view -> the TMA (async proxy) reads below
            }
            if (P.dbg) t_flag += clock64() - tw0;
            if (tracing) P.trace[(size_t)trace_item * 4 + 1] = ptx::globaltimer();
            ++trace_item;
            // the NEXT item's dependency range rides in this record: fetch this lane's entry now, its flag at the item's last step
            nxt_d0 = r1.z; nxt_cnt = r1.w; nxt_e = 0;
            if (lane < nxt_cnt) nxt_e = __ldg(P.deps + nxt_d0 + lane);
          }
          const int row0 = (2 * (int)(r0.y & 0xFFFFu) + (int)rank) * kRowTile;
          const long long tr0 = P.dbg ? clock64() : 0;
          if (it >= dep) ptx::mbar_wait(bar_empty + 8 * ((it - dep) & (TC2_NSLOT - 1)), ((it - dep) >> 3) & 1);   // step it-dep consumed
          if (dep != TC2_NSLOT && it >= TC2_NSLOT) ptx::mbar_wait(bar_empty + 8 * slot, ((it - TC2_NSLOT) >> 3) & 1);
          if (P.dbg) t_ring += clock64() - tr0;
          const uint32_t full = bar_full + 8 * slot;
          const uint32_t sa = smem_base + ((r0.x & 0xFFu) << 10);
          const uint32_t half_b = sg.half_b;
          const int n_half = (int)(sg.n_tile >> 1);
          if (ptx::elect_one()) {
            if (leader) ptx::mbar_expect_tx(full, 2u * ((uint32_t)nA * TC_A_BYTES + (uint32_t)nB * half_b));
#pragma unroll
            for (int a = 0; a < TC2_MAX_A; ++a) {
              if (a >= nA) break;
              const int p = (int)((((a < 2) ? r0.z : r0.w) >> (16 * (a & 1))) & 0xFFFFu);
              ptx::tma_load_3d_2sm(sa + a * TC_A_BYTES, &sg.tm_a, full, kc * 64, row0, p);
            }
            const uint32_t sb = sa + nA * TC_A_BYTES;
#pragma unroll
            for (int b = 0; b < TC2_MAX_BSLOTS; ++b) {
              if (b >= nB) break;
              const uint32_t e = ((b < 4) ? r1.x : r1.y) >> (8 * (b & 3));
              ptx::tma_load_3d_2sm(sb + b * half_b, &sg.tm_b, full, kc * 64, (int)((e >> 5) & 1u) * n_half, (int)(e & 0x1Fu));
            }
          }
          __syncwarp();
          if ((r0.y >> 22) & 1u) {
            // last step of the item issued: look at the next item's flags now, so that the answer is (usually) there by
            // the time its first step comes up
            cur_d0 = nxt_d0; cur_cnt = nxt_cnt; cur_e = nxt_e; cur_f = 0; cur_f_valid = true;
            if (lane < cur_cnt) cur_f = ptx::ld_acquire_gpu(P.flags + (cur_e & ~LOOP_DEP_PREV));
          }
        }
        __syncwarp();
      }
    }
    if (P.dbg && lane == 0) {
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_P_FLAG] = (unsigned long long)t_flag;
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_P_RING] = (unsigned long long)t_ring;
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_P_TOTAL] = (unsigned long long)(clock64() - t_p0);
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_P_SLOW] = (unsigned long long)n_slow;
    }
    // drain: nobody leaves while MMAs may still read this CTA's shared memory
    for (uint32_t j = it > TC2_NSLOT ? it - TC2_NSLOT : 0; j < it; ++j) ptx::mbar_wait(bar_empty + 8 * (j & (TC2_NSLOT - 1)), (j >> 3) & 1);
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const TcRec* __restrict__ stream = P.stream_m;
      const uint32_t ring = stg_base + TC2_REC_BATCH * (uint32_t)sizeof(TcRec);
      const uint64_t desc0 = make_smem_desc_sw128(smem_base);
      const uint32_t desc_lo0 = (uint32_t)desc0, desc_hi = (uint32_t)(desc0 >> 32);
      uint32_t it = 0, item_count = 0, buf = 0;
      uint32_t idesc = 0, acc_stride = 0, half_b16 = 0, n_merge = 0;
      long long t_full = 0, t_acc = 0;
      const long long t_m0 = P.dbg ? clock64() : 0;
      for (int pi = 0; pi < P.n_prog; ++pi) {
        const int sec = loop_prog_entry(P.n_groups, P.n_prog, pi).sec;
        const uint32_t rbeg = __ldg(soff + sec), rend = __ldg(soff + sec + 1);
        if (rbeg >= rend) continue;
        uint4 mine = make_uint4(0, 0, 0, 0);
        if (2 * rbeg + lane < 2 * rend) mine = __ldg(reinterpret_cast<const uint4*>(stream + rbeg) + lane);
        for (uint32_t base = rbeg; base < rend; base += TC2_REC_BATCH) {
          ptx::st_shared_v4(ring + lane * 16u, mine.x, mine.y, mine.z, mine.w);
          __syncwarp();
          if (2 * (base + TC2_REC_BATCH) + lane < 2 * rend) mine = __ldg(reinterpret_cast<const uint4*>(stream + base + TC2_REC_BATCH) + lane);
          const uint32_t cnt = min((uint32_t)TC2_REC_BATCH, rend - base);
          for (uint32_t i = 0; i < cnt; ++i, ++it) {
            const uint4 r0 = ptx::ld_shared_v4(ring + i * 32u);
            const uint4 r1 = ptx::ld_shared_v4(ring + i * 32u + 16u);
            const uint32_t slot = it & (TC2_NSLOT - 1), phase = (it >> 3) & 1;
            const int nA = (r0.x >> 8) & 0x7, n_ops = (r0.x >> 11) & 0x1F;
            const uint32_t flags = (r0.x >> 16) & 0x3u;
            if (flags & 1u) {                                   // first step of an item: its accumulator buffer must be drained
              const LoopSeg& sg = P.seg[r0.y & 0x1Fu];
              idesc = sg.idesc; acc_stride = sg.acc_stride; half_b16 = sg.half_b >> 4; n_merge = (sg.n_tile >> 3) << 17;
              buf = item_count & 1;
              const long long ta0 = P.dbg ? clock64() : 0;
              ptx::mbar_wait(bar_acc_empty + 8 * buf, ((item_count >> 1) & 1) ^ 1);
              if (P.dbg) t_acc += clock64() - ta0;
            }
            const long long tf0 = P.dbg ? clock64() : 0;
            ptx::mbar_wait(bar_full + 8 * slot, phase);
            if (P.dbg) t_full += clock64() - tf0;
            ptx::tc_fence_after();
            // descriptors differ only in the 14-bit start-address field: one 32-bit add each (smem < 256 KB, no carry)
            const uint32_t a_lo0 = desc_lo0 + ((r0.x & 0xFFu) << 6);
            const uint32_t b_lo0 = a_lo0 + (uint32_t)nA * (uint32_t)(TC_A_BYTES >> 4);
            if (ptx::elect_one()) {
              const uint32_t d0 = tmem_base + buf * TC2_BUF_COLS;
              const uint32_t opw[6] = {r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int oi = 0; oi < TC2_MAX_OPS; ++oi) {
                if (oi >= n_ops) break;
                const uint32_t e = opw[oi >> 1] >> (16 * (oi & 1));
                const uint32_t first_mma = (e >> 10) & 1u;
                const uint32_t a_lo = a_lo0 + (e & 3u) * (uint32_t)(TC_A_BYTES >> 4);
                const uint32_t b_lo = b_lo0 + ((e >> 2) & 7u) * half_b16;
                const uint32_t d = d0 + ((e >> 7) & 7u) * acc_stride;
                const uint32_t idg = idesc + ((e >> 5) & 3u) * n_merge;   // N = slots * N_TILE
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  ptx::umma_f16_2sm(d, ((uint64_t)desc_hi << 32) | (a_lo + 2u * k), ((uint64_t)desc_hi << 32) | (b_lo + 2u * k), idg,
                                    (k > 0 || !first_mma) ? 1u : 0u);
              }
              ptx::umma_commit_2sm(bar_empty + 8 * slot);           // this step is consumed (both CTAs)
              if (flags & 2u) ptx::umma_commit_2sm(bar_acc_full + 8 * buf);   // last step: accumulators complete in both CTAs
            }
            __syncwarp();
            if (flags & 2u) ++item_count;
          }
          __syncwarp();
        }
      }
      // drain: observe the release of the last (up to two) accumulator buffers by the epilogue warps of both CTAs
      for (uint32_t j = item_count > 2 ? item_count - 2 : 0; j < item_count; ++j) ptx::mbar_wait(bar_acc_empty + 8 * (j & 1), (j >> 1) & 1);
      if (P.dbg && lane == 0) {
        P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_M_FULL] = (unsigned long long)t_full;
        P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_M_ACC] = (unsigned long long)t_acc;
        P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_M_TOTAL] = (unsigned long long)(clock64() - t_m0);
      }
    }
   } else if (lane == 0) {
    // ===================== store warps (warp 2 + h serves epilogue half h; one lane) =====================
    // Takes the epilogue half's staged 128x64 fp16 tiles, stores them by TMA, frees the staging tile as soon as the
    // store has READ it, and publishes the item once its stores have COMPLETED (only the issuing thread can wait for that).
    const int h = warp - 2;
    const uint32_t tile_full = bar_base + 168 + 8 * (uint32_t)h, tile_free = bar_base + 184 + 8 * (uint32_t)h;
    const uint32_t s_out = epi_base + (uint32_t)h * TC2_TILE_BYTES;
    uint32_t tcount = 0;
    long long t_tile = 0, t_done = 0;
    const long long t_s0 = P.dbg ? clock64() : 0;
    for (int pi = 0; pi < P.n_prog; ++pi) {
      const int sec = loop_prog_entry(P.n_groups, P.n_prog, pi).sec;
      const uint32_t e_beg = __ldg(eoff + sec), e_end = __ldg(eoff + sec + 1);
      for (uint32_t k = e_beg; k < e_end; ++k) {
        const uint4 cur = __ldg(P.eitems + k);
        const int seg = (int)(cur.x >> 16), win = (int)(cur.x & 0xFFFFu), mp = (int)cur.y;
        const LoopSeg& sg = P.seg[seg];
        const uint32_t kind = sg.kind;
        if (!(kind <= LK_NONE64H)) continue;                       // fp32 / last-layer epilogues store (and publish) themselves
        const TcItem2* ip = sg.items + win;
        const int G = (int)(sg.n_tile >> 6), n_units = (int)__ldg(&ip->n_acc) * G;
        const int row0 = (2 * mp + (int)rank) * kRowTile;
        for (int u = h; u < n_units; u += 2) {
          const int q = (int)__ldg(&ip->q[u / G]);
          const long long tw0 = P.dbg ? clock64() : 0;
          ptx::mbar_wait(tile_full, tcount & 1);
          if (P.dbg) t_tile += clock64() - tw0;
          ptx::tma_store_3d(&sg.tm_out, s_out, (u % G) * 64, row0, q);
          ptx::bulk_commit();
          ptx::bulk_wait_read0();
          ptx::mbar_arrive(tile_free);
          ++tcount;
        }
        const long long tw1 = P.dbg ? clock64() : 0;
        ptx::bulk_wait_all0();                                      // the item's tiles are in global memory
        ptx::fence_proxy_async_all();
        ptx::red_release_gpu_add(P.flags + cur.z, 4u);   // for this half's four warps
        if (P.dbg) t_done += clock64() - tw1;
      }
    }
    if (P.dbg) {
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_S_TILE + 0] = (unsigned long long)t_tile;    // (warp 3 overwrites warp 2: same order of magnitude)
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_S_DONE] = (unsigned long long)t_done;
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_S_TOTAL] = (unsigned long long)(clock64() - t_s0);
    }
   }
  } else {
    ptx::setmaxnreg_inc<LOOP_REGS_EPI>();
    // ===================== epilogue (warps 4..11, both CTAs) =====================
    LoopCtx cx;
    cx.tmem_base = tmem_base; cx.bar_acc_full = bar_acc_full; cx.bar_acc_empty = bar_acc_empty; cx.epi_base = epi_base; cx.bar_base = bar_base;
    cx.warp = warp; cx.lane = lane; cx.rank = (int)rank; cx.item_count = 0; cx.tile_count = 0; cx.t_acc = 0; cx.t_tile = 0;
    const long long t_e0 = P.dbg ? clock64() : 0;
    TcFinalArgs fa{};
    fa.x = P.x; fa.y = P.y; fa.loss_part = P.loss_part; fa.R = P.R; fa.B = P.B; fa.n_rows = P.n_rows; fa.nbx = P.nbx; fa.w_out = P.w_out;
    fa.gscale = P.gscale; fa.m_gmul = P.m_gmul; fa.m_mu = P.m_mu;
    for (int pi = 0; pi < P.n_prog; ++pi) {
      const LoopProg pe = loop_prog_entry(P.n_groups, P.n_prog, pi);
      const uint32_t e_beg = __ldg(eoff + pe.sec), e_end = __ldg(eoff + pe.sec + 1);
      uint4 nxt = make_uint4(0, 0, 0, 0);
      if (e_beg < e_end) nxt = __ldg(P.eitems + e_beg);
      for (uint32_t k = e_beg; k < e_end; ++k, ++cx.item_count) {
        const uint4 cur = nxt;
        if (k + 1 < e_end) nxt = __ldg(P.eitems + k + 1);             // one item ahead
        const int seg = (int)(cur.x >> 16), win = (int)(cur.x & 0xFFFFu), mp = (int)cur.y;
        const LoopSeg& sg = P.seg[seg];
        const int t = sg.group ? pe.t1 : pe.t0;                       // this item's L-step
        fa.write_y = (t == P.last_step) ? 1 : 0;      // G(z) and the loss are consumed after the final forward only
        fa.m_lr = (P.decay_step > 0 && t >= P.decay_step) ? P.m_lr * 0.1f : P.m_lr;
        uint32_t* flag = P.flags + cur.z;
        unsigned long long ts = 0;
        if (P.prof != nullptr && warp == LOOP_EPI_WARP0 && lane == 0 && leader) ts = ptx::globaltimer();
        loop_epilogue_dispatch<ARCH>(P, sg, cx, fa, win, mp, flag);
        if (P.prof != nullptr && warp == LOOP_EPI_WARP0 && lane == 0 && leader) {
          unsigned long long* pr = P.prof + ((size_t)t * P.n_vseg + seg) * 2;
          const unsigned long long te = ptx::globaltimer();
          atomicMin(pr, ts);
          atomicMax(pr + 1, te);
          if (P.trace != nullptr && (pi == P.trace_entry || pi == P.trace_entry + 1)) { P.trace[(size_t)k * 4 + 2] = ts; P.trace[(size_t)k * 4 + 3] = te; }
        }
      }
    }
    if (P.dbg && warp == LOOP_EPI_WARP0 && lane == 0) {
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_E_ACC] = (unsigned long long)cx.t_acc;
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_E_TILE] = (unsigned long long)cx.t_tile;
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_E_TOTAL] = (unsigned long long)(clock64() - t_e0);
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();     // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}


// ------------------------------------------------------------------------------------------
// host side: the plan
// ------------------------------------------------------------------------------------------
struct LoopSegSpec {              // one layer-direction ("physical segment") as the planner sees it
  std::string name;
  int N = 0, K = 0;               // MMA N (output channels per pixel) and K (input channels per pixel)
  int kind = 0;                   // LoopKind
  const PairTable* tab = nullptr; // (input pixel, weight tile) contributions of every output pixel
  int h_grid = 1, w_grid = 1;     // raster of the output pixels (window shapes)
  int max_acc = 1;                // accumulators per window (TMEM columns / layer-specific cap)
  int in_seg = -1;                // segment whose output this one reads; -1: z, published by the momentum tail
  bool fwd = true;                // part of the generator forward (+ loss) or of the backward-to-z
  double macs_per_row = 0.0;      // exact in-bounds MACs per latent row (profiling only)
};

// The plan.  Row pairs are split into `n_groups` groups whose L-steps run HALF A STEP OUT OF PHASE: while group A is in
// its backward half, group B is in its forward half, and their segments alternate in every CTA pair's stream
// (A.bwd_0, B.fwd_0, A.bwd_1, B.fwd_1, ...).  Consecutive segments of one group are then separated by a segment of the
// other group, so the latency of "epilogue -> store completes -> flag -> dependent item's producer" (and of the momentum
// update at the L-step boundary) is covered by independent work instead of stalling the in-order streams - software
// pipelining across row-pair groups.  A (group, layer-direction) pair is a VIRTUAL SEGMENT with its own window tiling,
// items and flags.  Every CTA pair's records are four sections:
//   section 0 = A.fwd alone (prologue: A's first forward)      section 1 = A.bwd interleaved with B.fwd
//   section 3 = B.fwd alone (epilogue: B's last forward)       section 2 = A.fwd interleaved with B.bwd
// and a launch runs  sec0, (sec1, sec2) x (L - 1), sec3.   With one group (small batches, dgan_forward / dgan_loss_grad)
// sections 1 / 2 are that group's backward / forward and a launch runs  sec2, (sec1, sec2) x (L - 1).
struct LoopPlan {
  int n_phys = 0, n_groups = 1, n_vseg = 0, n_pairs = 0, n_mpairs = 0;
  std::vector<int> vseg_phys, vseg_group;         // vseg = group * n_phys + phys
  std::vector<std::vector<int>> group_mps;        // row pairs of each group
  std::vector<std::vector<TcItem2>> hdrs;         // per vseg: window headers
  std::vector<int> shape;                         // per vseg: wh, ww, sy, sx
  std::vector<uint32_t> flag_base, n_windows;     // per vseg: flags [row pair of the group][window]
  uint32_t zflag_base = 0, n_flags = 0;
  std::vector<int> sec_vsegs[LOOP_N_SEC];         // virtual segments of each section in stream order
  std::vector<TcRec> stream_p[2], stream_m;       // CTA pair after CTA pair, each: sections 0..3
  std::vector<uint32_t> stream_off;               // [n_pairs][LOOP_N_SEC + 1]
  std::vector<uint4> eitems;                      // x = vseg << 16 | window, y = row pair, z = flag index
  std::vector<uint32_t> eitem_off;                // [n_pairs][LOOP_N_SEC + 1]
  std::vector<uint32_t> dep_off, deps;
  long long n_steps = 0, n_mma = 0, n_bytes = 0;  // of sections 1 + 2 (= one L-step of every row pair)
};

#ifndef DGAN_COST_EPI_KB
#define DGAN_COST_EPI_KB 24.0
#endif
#ifndef DGAN_COST_FIXED_KB
#define DGAN_COST_FIXED_KB 48.0
#endif
#ifndef DGAN_LOOP_GROUPS
#define DGAN_LOOP_GROUPS 2        // row-pair groups of a projection (2 = half an L-step out of phase; 1 = all in phase, for A/B runs)
#endif
constexpr int LOOP_STEP_MAX_BYTES = 48 * 1024;    // measured optimum of the operand-ring kernels (round 1): 2 A tiles + weights

static double loop_item_cost(const Tc2HostItem& it, int N) {
  return it.stage_bytes + DGAN_COST_EPI_KB * 1024.0 * it.hdr.n_acc * std::max(1, N / 64) + DGAN_COST_FIXED_KB * 1024.0;
}

// Longest-processing-time assignment of the (window, row pair) items of one virtual segment to the CTA pairs, on top of
// the load `load` they already carry (the other virtual segment of the same slot: there is no barrier between them, so
// they are balanced jointly).  Items come out per CTA pair as (window, index into `mps`).
static void loop_lpt(const std::vector<Tc2HostItem>& items, int N, int n_mps, std::vector<double>* load,
                     std::vector<std::vector<std::pair<int, int>>>* lists) {
  const size_t n_pairs = load->size();
  lists->assign(n_pairs, {});
  std::vector<size_t> order(items.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t l, size_t r) { return items[l].stage_bytes > items[r].stage_bytes; });
  for (size_t oi = 0; oi < order.size(); ++oi)
    for (int m = 0; m < n_mps; ++m) {                     // cost-descending; the row pair is the fast index
      size_t best = 0;
      for (size_t pr = 1; pr < n_pairs; ++pr)
        if ((*load)[pr] < (*load)[best]) best = pr;
      (*load)[best] += loop_item_cost(items[order[oi]], N);
      (*lists)[best].push_back({(int)order[oi], m});
    }
}

// Window tiling of one virtual segment: every candidate shape (wh x ww accumulators, strides 1 or 2 - stride 2 gathers
// outputs of equal parity of a stride-2 transposed conv, which share weight tiles) is scored by the makespan of an LPT
// assignment of its items (cost = operand bytes staged + a per-accumulator epilogue charge + a fixed per-item charge).
static int loop_choose_tiling(const LoopSegSpec& sp, int n_mps, int n_pairs, std::vector<Tc2HostItem>* items_out, int shape_out[4]) {
  const int N = sp.N, K = sp.K;
  const int max_g = (N >= 64) ? std::min(4, 256 / N) : 1;
  const int step_max = std::min((LOOP_RING_BYTES / 2) & ~1023, LOOP_STEP_MAX_BYTES);
  double best_cost = 1e300;
  std::vector<Tc2HostItem> best_items;
  std::vector<std::vector<int>> wins;
  for (int wh = 1; wh <= 2; ++wh)
    for (int ww = 1; ww <= 8; ++ww)
      for (int sy = 1; sy <= (wh > 1 ? 2 : 1); ++sy)
        for (int sx = 1; sx <= (ww > 1 ? 2 : 1); ++sx) {
          if (wh * ww > sp.max_acc || wh > sp.h_grid || ww > std::max(sp.w_grid, 1)) continue;
          tc2_enumerate_windows(sp.h_grid, std::max(sp.w_grid, 1), wh, ww, sy, sx, &wins);
          if (wins.size() > 0xFFFFu) continue;
          std::vector<Tc2HostItem> items(wins.size());
          for (size_t i = 0; i < wins.size(); ++i) tc2_build_item(*sp.tab, wins[i], N, K, max_g, TC2_MAX_A, step_max, &items[i]);
          std::vector<double> load((size_t)n_pairs, 0.0);
          std::vector<std::vector<std::pair<int, int>>> lists;
          loop_lpt(items, N, n_mps, &load, &lists);
          const double makespan = *std::max_element(load.begin(), load.end());
          if (makespan < best_cost) {
            best_cost = makespan;
            shape_out[0] = wh; shape_out[1] = ww; shape_out[2] = sy; shape_out[3] = sx;
            best_items.swap(items);
          }
        }
  if (best_items.empty()) { set_error("no window tiling for segment " + sp.name); return DGAN_ERR_UNSUPPORTED; }
  items_out->swap(best_items);
  return 0;
}

// Plan for `n_mpairs` row pairs on `n_pairs` CTA pairs (see LoopPlan).
static int loop_plan(const std::vector<LoopSegSpec>& specs, int n_mpairs, int n_pairs, int n_groups, LoopPlan* plan) {
  const int n_phys = (int)specs.size();
  if (n_groups < 1 || n_groups > 2 || n_phys < 1 || n_phys * n_groups > LOOP_MAX_SEG) { set_error("segment count out of range"); return DGAN_ERR_UNSUPPORTED; }
  if (n_mpairs < n_groups || n_pairs < 1) { set_error("nothing to plan"); return DGAN_ERR_INVALID_ARG; }
  std::vector<int> fwd_list, bwd_list;
  for (int s = 0; s < n_phys; ++s) (specs[(size_t)s].fwd ? fwd_list : bwd_list).push_back(s);
  if (n_groups == 2 && fwd_list.size() != bwd_list.size()) { set_error("forward and backward halves differ in length"); return DGAN_ERR_UNSUPPORTED; }
  LoopPlan& pl = *plan;
  pl = LoopPlan{};
  pl.n_phys = n_phys; pl.n_groups = n_groups; pl.n_vseg = n_phys * n_groups; pl.n_pairs = n_pairs; pl.n_mpairs = n_mpairs;
  pl.group_mps.assign((size_t)n_groups, {});
  for (int mp = 0; mp < n_mpairs; ++mp) pl.group_mps[(size_t)(n_groups == 2 && mp >= (n_mpairs + 1) / 2 ? 1 : 0)].push_back(mp);
  const int nv = pl.n_vseg;
  pl.vseg_phys.resize((size_t)nv); pl.vseg_group.resize((size_t)nv);
  pl.hdrs.assign((size_t)nv, {}); pl.shape.assign((size_t)4 * nv, 1);
  pl.flag_base.assign((size_t)nv, 0); pl.n_windows.assign((size_t)nv, 0);
  std::vector<std::vector<Tc2HostItem>> items((size_t)nv);
  std::vector<std::vector<int>> pix2win((size_t)nv);
  int rc;
  uint32_t n_flags = 0;
  for (int v = 0; v < nv; ++v) {
    const int g = v / n_phys, s = v % n_phys;
    pl.vseg_phys[(size_t)v] = s; pl.vseg_group[(size_t)v] = g;
    const LoopSegSpec& sp = specs[(size_t)s];
    const int n_mps = (int)pl.group_mps[(size_t)g].size();
    if ((rc = loop_choose_tiling(sp, n_mps, n_pairs, &items[(size_t)v], &pl.shape[(size_t)4 * v]))) return rc;
    const size_t nw = items[(size_t)v].size();
    pl.hdrs[(size_t)v].resize(nw);
    pix2win[(size_t)v].assign(sp.tab->off.size() - 1, -1);
    for (size_t w = 0; w < nw; ++w) {
      pl.hdrs[(size_t)v][w] = items[(size_t)v][w].hdr;
      for (uint32_t a = 0; a < items[(size_t)v][w].hdr.n_acc; ++a) pix2win[(size_t)v][items[(size_t)v][w].hdr.q[a]] = (int)w;
    }
    pl.flag_base[(size_t)v] = n_flags;
    pl.n_windows[(size_t)v] = (uint32_t)nw;
    n_flags += (uint32_t)(nw * (size_t)n_mps);
  }
  pl.zflag_base = n_flags;
  n_flags += (uint32_t)n_mpairs;
  pl.n_flags = n_flags;
  // ---- sections
  auto vs = [&](int g, int s) { return g * n_phys + s; };
  for (int k = 0; k < LOOP_N_SEC; ++k) pl.sec_vsegs[k].clear();
  if (n_groups == 2) {
    for (int s : fwd_list) pl.sec_vsegs[0].push_back(vs(0, s));
    for (size_t k = 0; k < bwd_list.size(); ++k) { pl.sec_vsegs[1].push_back(vs(0, bwd_list[k])); pl.sec_vsegs[1].push_back(vs(1, fwd_list[k])); }
    for (size_t k = 0; k < fwd_list.size(); ++k) { pl.sec_vsegs[2].push_back(vs(0, fwd_list[k])); pl.sec_vsegs[2].push_back(vs(1, bwd_list[k])); }
    for (int s : fwd_list) pl.sec_vsegs[3].push_back(vs(1, s));
  } else {
    for (int s : bwd_list) pl.sec_vsegs[1].push_back(vs(0, s));
    for (int s : fwd_list) pl.sec_vsegs[2].push_back(vs(0, s));
  }
  // ---- assignment: per section, slot by slot (a slot = the two virtual segments that run side by side, balanced jointly)
  std::vector<std::vector<std::vector<std::pair<int, int>>>> lists[LOOP_N_SEC];     // [section][position in section][pair] -> (window, mp index)
  for (int k = 0; k < LOOP_N_SEC; ++k) {
    lists[k].resize(pl.sec_vsegs[k].size());
    std::vector<double> load((size_t)n_pairs, 0.0);
    for (size_t i = 0; i < pl.sec_vsegs[k].size(); ++i) {
      const int v = pl.sec_vsegs[k][i];
      const bool joint = (n_groups == 2 && (k == 1 || k == 2) && (i & 1));     // second member of a slot: on top of the first's load
      if (!joint) std::fill(load.begin(), load.end(), 0.0);
      loop_lpt(items[(size_t)v], specs[(size_t)pl.vseg_phys[(size_t)v]].N, (int)pl.group_mps[(size_t)pl.vseg_group[(size_t)v]].size(), &load, &lists[k][i]);
    }
  }
  // ---- emission
  const int ring_kb = LOOP_RING_BYTES / 1024;
  pl.stream_off.assign((size_t)n_pairs * (LOOP_N_SEC + 1), 0);
  pl.eitem_off.assign((size_t)n_pairs * (LOOP_N_SEC + 1), 0);
  pl.dep_off.assign(1, 0);
  for (int pr = 0; pr < n_pairs; ++pr) {
    std::vector<int> kb_of[LOOP_N_SEC];
    size_t sec_beg[LOOP_N_SEC + 1];
    for (int k = 0; k < LOOP_N_SEC; ++k) {
      sec_beg[k] = pl.stream_m.size();
      pl.stream_off[(size_t)pr * (LOOP_N_SEC + 1) + k] = (uint32_t)pl.stream_m.size();
      pl.eitem_off[(size_t)pr * (LOOP_N_SEC + 1) + k] = (uint32_t)pl.eitems.size();
      size_t prev_first_rec = (size_t)-1;
      for (size_t i = 0; i < pl.sec_vsegs[k].size(); ++i) {
        const int v = pl.sec_vsegs[k][i], g = pl.vseg_group[(size_t)v], s = pl.vseg_phys[(size_t)v];
        const LoopSegSpec& sp = specs[(size_t)s];
        std::vector<std::pair<int, int>> mine = lists[k][i][(size_t)pr];
        // row-pair major: a CTA pair meets the row pairs in the same order in every segment (measured: cost-descending
        // order instead is 12 % slower); inside a row pair keep the cost-descending order
        std::stable_sort(mine.begin(), mine.end(), [](const std::pair<int, int>& l, const std::pair<int, int>& r) { return l.second < r.second; });
        for (const auto& wm : mine) {
          const int win = wm.first, mi = wm.second, mp = pl.group_mps[(size_t)g][(size_t)mi];
          const Tc2HostItem& itm = items[(size_t)v][(size_t)win];
          const uint32_t flag = pl.flag_base[(size_t)v] + (uint32_t)mi * pl.n_windows[(size_t)v] + (uint32_t)win;
          pl.eitems.push_back(make_uint4(((uint32_t)v << 16) | (uint32_t)win, (uint32_t)mp, flag, 0u));
          // dependencies: the windows of the producing virtual segment (same group) that cover the staged input pixels
          std::vector<uint32_t> dl;
          if (sp.in_seg < 0) {
            dl.push_back((pl.zflag_base + (uint32_t)mp) | LOOP_DEP_PREV);
          } else {
            const int vin = vs(g, sp.in_seg);
            const std::vector<int>& p2w = pix2win[(size_t)vin];
            for (const Tc2HostStep& hs : itm.steps)
              for (int a = 0; a < hs.nA; ++a) {
                const int p = hs.a_pix[a];
                if (p < 0 || (size_t)p >= p2w.size() || p2w[(size_t)p] < 0) { set_error(sp.name + ": input pixel without a producer"); return DGAN_ERR_UNSUPPORTED; }
                const uint32_t f = pl.flag_base[(size_t)vin] + (uint32_t)mi * pl.n_windows[(size_t)vin] + (uint32_t)p2w[(size_t)p];
                if (std::find(dl.begin(), dl.end(), f) == dl.end()) dl.push_back(f);
              }
          }
          const uint32_t this_d0 = (uint32_t)pl.deps.size();
          pl.deps.insert(pl.deps.end(), dl.begin(), dl.end());
          pl.dep_off.push_back((uint32_t)pl.deps.size());
          // the first-step record of the PREVIOUS item of this section announces this item's dependency range, so that
          // the producer can fetch the flags one item ahead
          if (prev_first_rec != (size_t)-1)
            for (int r = 0; r < 2; ++r) { pl.stream_p[r][prev_first_rec].w[6] = this_d0; pl.stream_p[r][prev_first_rec].w[7] = (uint32_t)dl.size(); }
          prev_first_rec = pl.stream_m.size();
          for (size_t j = 0; j < itm.steps.size(); ++j) {
            const Tc2HostStep& hs = itm.steps[j];
            const int kb = (hs.bytes + 1023) / 1024;
            if (kb > ring_kb / 2) { set_error("tensor-core step larger than half the operand ring"); return DGAN_ERR_UNSUPPORTED; }
            kb_of[k].push_back(kb);
            const uint32_t flags = (j == 0 ? 1u : 0u) | (j + 1 == itm.steps.size() ? 2u : 0u);
            TcRec rm{};
            rm.w[0] = ((uint32_t)hs.nA << 8) | ((uint32_t)hs.n_ops << 11) | (flags << 16);   // ring offset filled below
            rm.w[1] = (uint32_t)v;
            for (int o = 0; o < hs.n_ops; ++o) rm.w[2 + o / 2] |= (uint32_t)hs.ops[o] << (16 * (o & 1));
            pl.stream_m.push_back(rm);
            for (int r = 0; r < 2; ++r) {
              TcRec rp{};
              rp.w[0] = ((uint32_t)hs.kc << 8) | ((uint32_t)hs.nA << 12) | ((uint32_t)hs.nB << 15);   // offset + dep filled below
              rp.w[1] = (uint32_t)mp | ((uint32_t)v << 16) | ((j == 0 ? 1u : 0u) << 21) | ((j + 1 == itm.steps.size() ? 1u : 0u) << 22);
              for (int a = 0; a < hs.nA; ++a) rp.w[2 + a / 2] |= (uint32_t)(hs.a_pix[a] & 0xFFFF) << (16 * (a & 1));
              for (int b = 0; b < hs.nB; ++b) rp.w[4 + b / 4] |= (uint32_t)hs.b_ent[r][b] << (8 * (b & 3));
              pl.stream_p[r].push_back(rp);
            }
            if (k == 1 || k == 2) { pl.n_mma += hs.n_ops; pl.n_steps += 1; pl.n_bytes += hs.bytes; }
          }
        }
      }
    }
    sec_beg[LOOP_N_SEC] = pl.stream_m.size();
    pl.stream_off[(size_t)pr * (LOOP_N_SEC + 1) + LOOP_N_SEC] = (uint32_t)pl.stream_m.size();
    pl.eitem_off[(size_t)pr * (LOOP_N_SEC + 1) + LOOP_N_SEC] = (uint32_t)pl.eitems.size();
    // ---- circular operand ring of this CTA pair.  Sequential allocation, wrap when a step does not fit.  Sections 1 and 2
    //      alternate for the whole launch: they are planned as ONE cyclic sequence (a step's dependency distance may reach
    //      back across the section boundary, also from the start of section 1 to the end of section 2).  Sections 0 and 3
    //      start on an empty ring (kernel start / explicit drain) and only look back inside themselves.
    auto plan_ring = [&](const std::vector<int>& kbs, size_t rec0, bool cyclic) {
      const int n = (int)kbs.size();
      std::vector<int> beg((size_t)n), end((size_t)n);
      int cursor = 0;
      for (int k2 = 0; k2 < n; ++k2) {
        if (cursor + kbs[(size_t)k2] > ring_kb) cursor = 0;
        beg[(size_t)k2] = cursor; end[(size_t)k2] = cursor + kbs[(size_t)k2];
        cursor = end[(size_t)k2];
      }
      // dep = distance (in steps) to the latest earlier step whose region overlaps this step's: the producer may overwrite
      // the region once that step is consumed (8 = barrier-slot reuse only)
      for (int k2 = 0; k2 < n; ++k2) {
        int dep = TC2_NSLOT;
        for (int d = 1; d < TC2_NSLOT; ++d) {
          int c = k2 - d;
          if (c < 0) { if (!cyclic) break; c = (c % n + n) % n; }
          if (beg[(size_t)c] < end[(size_t)k2] && beg[(size_t)k2] < end[(size_t)c]) { dep = d; break; }
        }
        const size_t ri = rec0 + (size_t)k2;
        pl.stream_m[ri].w[0] |= (uint32_t)beg[(size_t)k2];
        for (int r = 0; r < 2; ++r) pl.stream_p[r][ri].w[0] |= (uint32_t)beg[(size_t)k2] | ((uint32_t)dep << 19);
      }
    };
    plan_ring(kb_of[0], sec_beg[0], false);
    {
      std::vector<int> both = kb_of[1];
      both.insert(both.end(), kb_of[2].begin(), kb_of[2].end());
      plan_ring(both, sec_beg[1], true);            // sections 1 and 2 are adjacent in the record arrays
    }
    plan_ring(kb_of[3], sec_beg[3], false);
  }
  return 0;
}

// The section sequence of a launch (LoopProg entries): `full_last` = the final L-step also runs its backward half without
// the momentum update (dgan_loss_grad); otherwise the final L-step is forward only (the loop returns the pre-update
// forward of iteration L-1, models/gan.py:419-421).
static std::vector<LoopProg> loop_program(const LoopPlan& pl, int rec_iters, bool full_last) {
  const int n = loop_prog_length(pl.n_groups, rec_iters, full_last);
  std::vector<LoopProg> pg((size_t)n);
  for (int pi = 0; pi < n; ++pi) pg[(size_t)pi] = loop_prog_entry(pl.n_groups, n, pi);
  return pg;
}

// Independent validation of a plan (host only; dgan_debug_check_plans and the CPU tests):
//  * per (section, virtual segment), everything tc2_check_plan re-derives from the records (every (output pixel, input
//    pixel, tap, k-chunk) contribution exactly once into the right accumulator, first-MMA flags, canonical accumulation
//    order, every item of the group assigned exactly once);
//  * the operand ring: a step's region overlaps none of the `dep - 1` steps before it (which may still be unread when its
//    loads start) - cyclically over sections 1 + 2, linearly inside sections 0 and 3;
//  * producer, MMA and epilogue streams agree on the item sequence of every CTA pair; every record announces the next
//    item's dependency range; the flag index of an item is the one its consumers wait for;
//  * every input pixel an item stages is covered by a dependency on the window (producing virtual segment of the same
//    group, same row pair) that writes it, and the first segment waits for the previous L-step's z update;
//  * no deadlock: executing a 3-step launch program with every CTA pair in order and an item startable only when its
//    dependencies have reached the value it waits for, every item of the program completes.
static int loop_check_plan(const std::vector<LoopSegSpec>& specs, const LoopPlan& pl, std::string* err) {
  auto fail = [&](const std::string& m) { *err = m; return DGAN_ERR_INVALID_ARG; };
  const int n_phys = pl.n_phys, nv = pl.n_vseg, n_pairs = pl.n_pairs;
  if ((int)specs.size() != n_phys || nv != n_phys * pl.n_groups) return fail("segment count");
  if (pl.stream_off.size() != (size_t)n_pairs * (LOOP_N_SEC + 1) || pl.eitem_off.size() != pl.stream_off.size()) return fail("offset table size");
  if (pl.stream_p[0].size() != pl.stream_m.size() || pl.stream_p[1].size() != pl.stream_m.size()) return fail("stream sizes differ");
  if (pl.dep_off.size() != pl.eitems.size() + 1) return fail("dep_off size");
  const int ring_kb = LOOP_RING_BYTES / 1024;
  std::vector<std::vector<int>> pix2win((size_t)nv);
  for (int v = 0; v < nv; ++v) {
    pix2win[(size_t)v].assign(specs[(size_t)pl.vseg_phys[(size_t)v]].tab->off.size() - 1, -1);
    for (size_t w = 0; w < pl.hdrs[(size_t)v].size(); ++w)
      for (uint32_t a = 0; a < pl.hdrs[(size_t)v][w].n_acc; ++a) pix2win[(size_t)v][pl.hdrs[(size_t)v][w].q[a]] = (int)w;
  }
  std::vector<int> mp_index((size_t)pl.n_mpairs, -1), mp_group((size_t)pl.n_mpairs, -1);
  for (int g = 0; g < pl.n_groups; ++g)
    for (size_t i = 0; i < pl.group_mps[(size_t)g].size(); ++i) { mp_index[(size_t)pl.group_mps[(size_t)g][i]] = (int)i; mp_group[(size_t)pl.group_mps[(size_t)g][i]] = g; }
  for (int mp = 0; mp < pl.n_mpairs; ++mp)
    if (mp_index[(size_t)mp] < 0) return fail("a row pair belongs to no group");
  // ---- per (section, virtual segment): slice and reuse the single-layer validator
  for (int k = 0; k < LOOP_N_SEC; ++k)
    for (int v : pl.sec_vsegs[k]) {
      const LoopSegSpec& sp = specs[(size_t)pl.vseg_phys[(size_t)v]];
      const int g = pl.vseg_group[(size_t)v], n_mps = (int)pl.group_mps[(size_t)g].size();
      Tc2Plan sub;
      sub.n_pairs = n_pairs;
      sub.hdrs = pl.hdrs[(size_t)v];
      sub.stream_off.assign((size_t)n_pairs + 1, 0);
      std::vector<std::vector<int>> per_pair((size_t)n_pairs);
      for (int pr = 0; pr < n_pairs; ++pr) {
        const uint32_t r0 = pl.stream_off[(size_t)pr * (LOOP_N_SEC + 1) + k], r1 = pl.stream_off[(size_t)pr * (LOOP_N_SEC + 1) + k + 1];
        if (r0 > r1 || r1 > pl.stream_m.size()) return fail("stream offsets not monotone");
        sub.stream_off[(size_t)pr] = (uint32_t)sub.stream_m.size();
        for (uint32_t ri = r0; ri < r1; ++ri) {
          TcRec m = pl.stream_m[ri], p0 = pl.stream_p[0][ri], p1 = pl.stream_p[1][ri];
          if ((int)((p0.w[1] >> 16) & 0x1F) != (int)(m.w[1] & 0x1F) || p0.w[1] != p1.w[1]) return fail(sp.name + ": records of one step name different segments");
          if ((int)(m.w[1] & 0x1F) != v) continue;
          if (((p0.w[1] >> 21) & 1u) != ((m.w[0] >> 16) & 1u) || ((p0.w[1] >> 22) & 1u) != ((m.w[0] >> 17) & 1u)) return fail(sp.name + ": first/last-step marks of producer and MMA records disagree");
          const int mp = (int)(p0.w[1] & 0xFFFFu);
          if (mp >= pl.n_mpairs || mp_group[(size_t)mp] != g) return fail(sp.name + ": a record's row pair is not in its group");
          p0.w[1] = (uint32_t)mp_index[(size_t)mp]; p1.w[1] = p0.w[1]; m.w[1] = 0;
          p0.w[6] = p0.w[7] = p1.w[6] = p1.w[7] = 0;
          sub.stream_m.push_back(m); sub.stream_p[0].push_back(p0); sub.stream_p[1].push_back(p1);
        }
        const uint32_t e0 = pl.eitem_off[(size_t)pr * (LOOP_N_SEC + 1) + k], e1 = pl.eitem_off[(size_t)pr * (LOOP_N_SEC + 1) + k + 1];
        if (e0 > e1 || e1 > pl.eitems.size()) return fail("item offsets not monotone");
        for (uint32_t e = e0; e < e1; ++e) {
          const uint4 it = pl.eitems[e];
          if ((int)(it.x >> 16) != v) continue;
          const uint32_t win = it.x & 0xFFFFu;
          if (win > 0x7FFFu || it.y >= (uint32_t)pl.n_mpairs || mp_group[it.y] != g) return fail(sp.name + ": bad item");
          if (it.z != pl.flag_base[(size_t)v] + (uint32_t)mp_index[it.y] * pl.n_windows[(size_t)v] + win) return fail(sp.name + ": an item publishes the wrong flag");
          per_pair[(size_t)pr].push_back((int)((win << 16) | (uint32_t)mp_index[it.y]));
        }
      }
      sub.stream_off[(size_t)n_pairs] = (uint32_t)sub.stream_m.size();
      size_t n_slots = 0;
      for (auto& l : per_pair) n_slots = std::max(n_slots, l.size());
      sub.n_slots = (int)n_slots;
      sub.eitems.assign(n_slots * (size_t)n_pairs, -1);
      for (int pr = 0; pr < n_pairs; ++pr)
        for (size_t i = 0; i < per_pair[(size_t)pr].size(); ++i) sub.eitems[i * (size_t)n_pairs + pr] = per_pair[(size_t)pr][i];
      std::string e2;
      if (tc2_check_plan(sp.N, sp.K, *sp.tab, n_mps, LOOP_RING_BYTES, sub, &e2, /*check_ring=*/false)) return fail(sp.name + " (section " + std::to_string(k) + "): " + e2);
    }
  // ---- ring, stream/item agreement, dependency coverage: per CTA pair and section
  for (int pr = 0; pr < n_pairs; ++pr) {
    const uint32_t* so = &pl.stream_off[(size_t)pr * (LOOP_N_SEC + 1)];
    const uint32_t* eo = &pl.eitem_off[(size_t)pr * (LOOP_N_SEC + 1)];
    auto region = [&](uint32_t ri, int* beg, int* end, int* dep) {
      const TcRec& p0 = pl.stream_p[0][ri];
      const int v = (int)((p0.w[1] >> 16) & 0x1F);
      const int nA = (int)((p0.w[0] >> 12) & 7), nB = (int)((p0.w[0] >> 15) & 0xF);
      *beg = (int)(p0.w[0] & 0xFF);
      *end = *beg + (nA * TC_A_BYTES + nB * (specs[(size_t)pl.vseg_phys[(size_t)v]].N / 2) * 128 + 1023) / 1024;
      *dep = (int)((p0.w[0] >> 19) & 0xF);
    };
    auto check_ring = [&](uint32_t r0, uint32_t r1, bool cyclic) -> int {
      const int n = (int)(r1 - r0);
      for (int k2 = 0; k2 < n; ++k2) {
        int b, e, d;
        region(r0 + k2, &b, &e, &d);
        if (e > ring_kb) return fail("step region outside the ring");
        if (d < 1 || d > TC2_NSLOT) return fail("dep out of range");
        if ((pl.stream_m[r0 + k2].w[0] & 0xFF) != (uint32_t)b) return fail("producer and MMA records place a step differently");
        for (int dd = 1; dd < d; ++dd) {
          int c = k2 - dd;
          if (c < 0) { if (!cyclic) break; c = (c % n + n) % n; }
          int cb, ce, cd;
          region(r0 + c, &cb, &ce, &cd);
          if (cb < e && b < ce) return fail("ring hazard: a region may be overwritten while it can still be read");
        }
      }
      return 0;
    };
    int rc2;
    if ((rc2 = check_ring(so[0], so[1], false)) || (rc2 = check_ring(so[1], so[3], true)) || (rc2 = check_ring(so[3], so[4], false))) return rc2;
    for (int k = 0; k < LOOP_N_SEC; ++k) {
      uint32_t e = eo[k];
      std::vector<uint32_t> need;
      const int n = (int)(so[k + 1] - so[k]);
      for (int k2 = 0; k2 <= n; ++k2) {
        const bool first = (k2 < n) && ((pl.stream_p[0][so[k] + k2].w[1] >> 21) & 1u);
        if ((first || k2 == n) && k2 > 0) {
          const uint32_t d0 = pl.dep_off[e], d1 = pl.dep_off[e + 1];
          for (uint32_t f : need)
            if (std::find(pl.deps.begin() + d0, pl.deps.begin() + d1, f) == pl.deps.begin() + d1) return fail("an item stages a pixel it has no dependency on");
          for (uint32_t d = d0; d < d1; ++d)
            if ((pl.deps[d] & ~LOOP_DEP_PREV) >= pl.n_flags) return fail("dependency index out of range");
          need.clear();
          ++e;
        }
        if (k2 == n) break;
        const TcRec& p0 = pl.stream_p[0][so[k] + k2];
        const int v = (int)((p0.w[1] >> 16) & 0x1F), mp = (int)(p0.w[1] & 0xFFFF), nA = (int)((p0.w[0] >> 12) & 7);
        if (v >= nv) return fail("record names an unknown segment");
        const int g = pl.vseg_group[(size_t)v];
        if (first) {
          if (e >= eo[k + 1] || (int)(pl.eitems[e].x >> 16) != v || (int)pl.eitems[e].y != mp) return fail("producer stream and item list disagree");
          const uint32_t want_d0 = (e + 1 < eo[k + 1]) ? pl.dep_off[e + 1] : 0u, want_cnt = (e + 1 < eo[k + 1]) ? pl.dep_off[e + 2] - pl.dep_off[e + 1] : 0u;
          for (int r = 0; r < 2; ++r) {
            const TcRec& pq = pl.stream_p[r][so[k] + k2];
            if (pq.w[7] != want_cnt || (want_cnt != 0 && pq.w[6] != want_d0)) return fail("a record announces the wrong dependency range for the next item");
          }
        }
        const int in_seg = specs[(size_t)pl.vseg_phys[(size_t)v]].in_seg;
        if (in_seg < 0) {
          const uint32_t f = (pl.zflag_base + (uint32_t)mp) | LOOP_DEP_PREV;
          if (std::find(need.begin(), need.end(), f) == need.end()) need.push_back(f);
        } else {
          const int vin = g * n_phys + in_seg;
          for (int a = 0; a < nA; ++a) {
            const int p = (int)((p0.w[2 + a / 2] >> (16 * (a & 1))) & 0xFFFF);
            if ((size_t)p >= pix2win[(size_t)vin].size() || pix2win[(size_t)vin][(size_t)p] < 0) return fail("staged pixel has no producing window");
            const uint32_t f = pl.flag_base[(size_t)vin] + (uint32_t)mp_index[(size_t)mp] * pl.n_windows[(size_t)vin] + (uint32_t)pix2win[(size_t)vin][(size_t)p];
            if (std::find(need.begin(), need.end(), f) == need.end()) need.push_back(f);
          }
        }
      }
      if (e != eo[k + 1]) return fail("item count of a section's stream and its item list differ");
    }
  }
  // ---- deadlock check over a 3-step launch program.  flag value = completions so far; an item of L-step t waits for
  //      "dependency completed t + 1 times" (or, for the z update, t times: the momentum tail of L-step t - 1).
  {
    const int L = 3;
    const std::vector<LoopProg> pg = loop_program(pl, L, false);
    std::vector<int> done(pl.n_flags, 0), lin_parts((size_t)pl.n_mpairs * L, 0);
    const int last_phys = n_phys - 1;          // the Linear backward: its items complete the z update of their row pair
    int parts_per_mp = 0;
    for (int v = 0; v < nv; ++v)
      if (pl.vseg_phys[(size_t)v] == last_phys && pl.vseg_group[(size_t)v] == 0) parts_per_mp = (int)pl.n_windows[(size_t)v];
    struct Cur { size_t pi; uint32_t e; };
    std::vector<Cur> cur((size_t)n_pairs, Cur{0, 0});
    for (int pr = 0; pr < n_pairs; ++pr) cur[(size_t)pr].e = pl.eitem_off[(size_t)pr * (LOOP_N_SEC + 1) + pg[0].sec];
    size_t remaining = 0;
    for (const LoopProg& pe : pg)
      for (int pr = 0; pr < n_pairs; ++pr) remaining += pl.eitem_off[(size_t)pr * (LOOP_N_SEC + 1) + pe.sec + 1] - pl.eitem_off[(size_t)pr * (LOOP_N_SEC + 1) + pe.sec];
    bool progress = true;
    while (remaining > 0 && progress) {
      progress = false;
      for (int pr = 0; pr < n_pairs; ++pr) {
        Cur& c = cur[(size_t)pr];
        while (c.pi < pg.size()) {
          const LoopProg& pe = pg[c.pi];
          const uint32_t lim = pl.eitem_off[(size_t)pr * (LOOP_N_SEC + 1) + pe.sec + 1];
          if (c.e >= lim) {
            ++c.pi;
            if (c.pi < pg.size()) c.e = pl.eitem_off[(size_t)pr * (LOOP_N_SEC + 1) + pg[c.pi].sec];
            continue;
          }
          const uint4 it = pl.eitems[c.e];
          const int v = (int)(it.x >> 16), t = pl.vseg_group[(size_t)v] ? pe.t1 : pe.t0;
          if (t < 0) return fail("the program runs a group's item in a section where that group has no L-step");
          bool ready = true;
          for (uint32_t d = pl.dep_off[c.e]; d < pl.dep_off[c.e + 1] && ready; ++d) {
            const uint32_t f = pl.deps[d] & ~LOOP_DEP_PREV;
            if (pl.deps[d] & LOOP_DEP_PREV) ready = done[f] >= t;
            else ready = done[f] >= t + 1;
          }
          if (!ready) break;
          if (done[it.z] != t) return fail("an item runs out of L-step order");
          done[it.z] = t + 1;
          if (pl.vseg_phys[(size_t)v] == last_phys && t < L - 1) {
            if (++lin_parts[(size_t)it.y * L + t] == parts_per_mp) done[pl.zflag_base + it.y] = t + 1;
          }
          ++c.e; --remaining; progress = true;
        }
      }
    }
    if (remaining > 0) return fail("deadlock: some items can never start (dependency cycle across in-order streams)");
  }
  return 0;
}

}  // namespace dgan
