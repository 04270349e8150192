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

constexpr int LOOP_MAX_SEG = 10;
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
  uint32_t idesc, half_b, flag_base, n_windows;
};

struct LoopParams {
  LoopSeg seg[LOOP_MAX_SEG];
  const TcRec* stream_p[2];        // producer records per cluster rank: per CTA pair, the steps of ONE L-step
  const TcRec* stream_m;           // MMA records, same indexing
  const uint32_t* stream_off;      // [n_pairs][n_seg + 1] record offsets (segment boundaries inside a pair's stream)
  const uint2* eitems;             // items in stream order, all pairs: x = seg << 16 | window, y = row pair
  const uint32_t* eitem_off;       // [n_pairs][n_seg + 1] item offsets
  const uint32_t* dep_off;         // [items + 1] -> deps
  const uint32_t* deps;            // flag indices (| LOOP_DEP_PREV)
  uint32_t* flags;                 // zeroed per call
  uint32_t* status;                // [0] != 0: a flag wait timed out (results invalid)
  unsigned long long* prof;        // optional [t][n_seg][2] globaltimer min-start / max-end
  unsigned long long* dbg;         // optional [CTA][16] stall counters of the roles (clock64 ticks), see LoopDbg
  unsigned long long* trace;       // optional [items of one L-step][4] globaltimer: dependency wait begin / end, epilogue begin / end
  int trace_step;                  // the L-step that is traced
  int n_seg, n_seg_last;           // segments per L-step; segments of the LAST L-step (forward only: SURVEY F4)
  int seg_begin, seg_end;          // segment sub-range of this launch (whole step: 0, n_seg)
  int t_begin, t_end;              // L-steps of this launch
  int last_step;                   // index of the call's final L-step (rec_iters - 1)
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
      constexpr int STRIDE = 4 * 32 * TC2_EPI_WARPS, UNR = 4;
      for (int e0 = tid * 4; e0 < kRowTile * N_TILE; e0 += UNR * STRIDE) {
        float4 gs[UNR], vv[UNR], zz[UNR];
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
          const size_t i = base + (size_t)(e0 + k * STRIDE);
          gs[k] = __ldcg(reinterpret_cast<const float4*>(gp + i));
          vv[k] = __ldcg(reinterpret_cast<const float4*>(P.mv + i));
          zz[k] = __ldcg(reinterpret_cast<const float4*>(P.mz + i));
        }
        for (int pp = 1; pp < P.m_nparts; ++pp) {          // fixed order: parts 0, 1, 2, ...
          float4 tt[UNR];
#pragma unroll
          for (int k = 0; k < UNR; ++k) tt[k] = __ldcg(reinterpret_cast<const float4*>(gp + base + (size_t)(e0 + k * STRIDE) + (size_t)pp * P.m_count));
#pragma unroll
          for (int k = 0; k < UNR; ++k) { gs[k].x += tt[k].x; gs[k].y += tt[k].y; gs[k].z += tt[k].z; gs[k].w += tt[k].w; }
        }
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
          const size_t i = base + (size_t)(e0 + k * STRIDE);
          float4 v4 = vv[k], z4 = zz[k];
          v4.x = fmaf(fa.m_mu, v4.x, fa.m_gmul * gs[k].x); v4.y = fmaf(fa.m_mu, v4.y, fa.m_gmul * gs[k].y);
          v4.z = fmaf(fa.m_mu, v4.z, fa.m_gmul * gs[k].z); v4.w = fmaf(fa.m_mu, v4.w, fa.m_gmul * gs[k].w);
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

  const uint32_t* __restrict__ soff = P.stream_off + (size_t)pair * (P.n_seg + 1);
  const uint32_t* __restrict__ eoff = P.eitem_off + (size_t)pair * (P.n_seg + 1);

  if (warp < LOOP_EPI_WARP0) {
   ptx::setmaxnreg_dec<LOOP_REGS_CTRL>();
   if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const TcRec* __restrict__ stream = rank ? P.stream_p[1] : P.stream_p[0];
    const uint32_t ring = stg_base;
    uint32_t it = 0;                                      // steps issued so far, over all replays (barrier slot / phase)
    const uint32_t rbeg = __ldg(soff + P.seg_begin);
    uint4 first = make_uint4(0, 0, 0, 0);                 // the replays all start with the same records
    {
      const uint32_t rend_max = __ldg(soff + P.seg_end);
      if (2 * rbeg + lane < 2 * rend_max) first = __ldg(reinterpret_cast<const uint4*>(stream + rbeg) + lane);
    }
    // dependencies of the replay's first item (the same every replay); later items are described one item ahead by the
    // records themselves (w[6], w[7] of an item's first step = dependency range of the NEXT item)
    uint32_t first_d0 = 0, first_cnt = 0;
    {
      const uint32_t i0 = __ldg(eoff + P.seg_begin), i1 = __ldg(eoff + P.seg_end);
      if (i0 < i1) { first_d0 = __ldg(P.dep_off + i0); first_cnt = __ldg(P.dep_off + i0 + 1) - first_d0; }
    }
    long long t_flag = 0, t_ring = 0, n_slow = 0;
    const long long t_p0 = P.dbg ? clock64() : 0;
    for (int t = P.t_begin; t < P.t_end; ++t) {
      const int s_end = min(P.seg_end, (t == P.last_step) ? P.n_seg_last : P.n_seg);
      if (s_end <= P.seg_begin) continue;
      const uint32_t rend = __ldg(soff + s_end);
      // the item about to start: dependency range, this lane's entry (first 32) and its flag value if already fetched
      uint32_t cur_d0 = first_d0, cur_cnt = first_cnt, cur_e = 0, cur_f = 0;
      bool cur_f_valid = false;
      if (lane < cur_cnt) cur_e = __ldg(P.deps + cur_d0 + lane);
      uint32_t nxt_d0 = 0, nxt_cnt = 0, nxt_e = 0;
      uint32_t trace_item = __ldg(eoff + P.seg_begin);          // index of the item about to start (trace only)
      uint4 mine = first;
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
          const int seg = (int)((r0.y >> 16) & 0xFu);
          const LoopSeg& sg = P.seg[seg];
          if ((r0.y >> 20) & 1u) {
            // first step of an item: everything it stages must have been published.  Fast path: the flag values were
            // fetched while the previous item's last step was issued and already satisfy the target.
            const long long tw0 = P.dbg ? clock64() : 0;
            const bool tracing = P.trace != nullptr && t == P.trace_step && leader && lane == 0;
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
          if ((r0.y >> 21) & 1u) {
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
      const uint32_t rbeg = __ldg(soff + P.seg_begin);
      uint4 first = make_uint4(0, 0, 0, 0);
      {
        const uint32_t rend_max = __ldg(soff + P.seg_end);
        if (2 * rbeg + lane < 2 * rend_max) first = __ldg(reinterpret_cast<const uint4*>(stream + rbeg) + lane);
      }
      for (int t = P.t_begin; t < P.t_end; ++t) {
        const int s_end = min(P.seg_end, (t == P.last_step) ? P.n_seg_last : P.n_seg);
        if (s_end <= P.seg_begin) continue;
        const uint32_t rend = __ldg(soff + s_end);
        uint4 mine = first;
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
              const LoopSeg& sg = P.seg[r0.y & 0xFu];
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
    for (int t = P.t_begin; t < P.t_end; ++t) {
      const int s_end = min(P.seg_end, (t == P.last_step) ? P.n_seg_last : P.n_seg);
      if (s_end <= P.seg_begin) continue;
      const uint32_t e_beg = __ldg(eoff + P.seg_begin), e_end = __ldg(eoff + s_end);
      for (uint32_t k = e_beg; k < e_end; ++k) {
        const uint2 cur = __ldg(P.eitems + k);
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
        ptx::red_release_gpu_add(P.flags + sg.flag_base + (size_t)mp * sg.n_windows + win, 4u);   // for this half's four warps
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
    for (int t = P.t_begin; t < P.t_end; ++t) {
      const int s_end = min(P.seg_end, (t == P.last_step) ? P.n_seg_last : P.n_seg);
      if (s_end <= P.seg_begin) continue;
      fa.write_y = (t == P.last_step) ? 1 : 0;      // G(z) and the loss are consumed after the final forward only
      fa.m_lr = (P.decay_step > 0 && t >= P.decay_step) ? P.m_lr * 0.1f : P.m_lr;
      const uint32_t e_beg = __ldg(eoff + P.seg_begin), e_end = __ldg(eoff + s_end);
      uint2 nxt = make_uint2(0, 0);
      if (e_beg < e_end) nxt = __ldg(P.eitems + e_beg);
      for (uint32_t k = e_beg; k < e_end; ++k, ++cx.item_count) {
        const uint2 cur = nxt;
        if (k + 1 < e_end) nxt = __ldg(P.eitems + k + 1);             // one item ahead
        const int seg = (int)(cur.x >> 16), win = (int)(cur.x & 0xFFFFu), mp = (int)cur.y;
        const LoopSeg& sg = P.seg[seg];
        uint32_t* flag = P.flags + sg.flag_base + (size_t)mp * sg.n_windows + win;
        unsigned long long ts = 0;
        if (P.prof != nullptr && warp == LOOP_EPI_WARP0 && lane == 0 && leader) ts = ptx::globaltimer();
        loop_epilogue_dispatch<ARCH>(P, sg, cx, fa, win, mp, flag);
        if (P.prof != nullptr && warp == LOOP_EPI_WARP0 && lane == 0 && leader) {
          unsigned long long* pr = P.prof + ((size_t)t * P.n_seg + seg) * 2;
          const unsigned long long te = ptx::globaltimer();
          atomicMin(pr, ts);
          atomicMax(pr + 1, te);
          if (P.trace != nullptr && t == P.trace_step) { P.trace[(size_t)k * 4 + 2] = ts; P.trace[(size_t)k * 4 + 3] = te; }
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
// host side: the plan of one L-step
// ------------------------------------------------------------------------------------------
struct LoopSegSpec {              // one layer-direction as the planner sees it
  std::string name;
  int N = 0, K = 0;               // MMA N (output channels per pixel) and K (input channels per pixel)
  int kind = 0;                   // LoopKind
  const PairTable* tab = nullptr; // (input pixel, weight tile) contributions of every output pixel
  int h_grid = 1, w_grid = 1;     // raster of the output pixels (window shapes)
  int max_acc = 1;                // accumulators per window (TMEM columns / layer-specific cap)
  int in_seg = -1;                // segment whose output this one reads; -1: z, published by the momentum tail
  double macs_per_row = 0.0;      // exact in-bounds MACs per latent row (profiling only)
};

struct LoopPlan {
  int n_seg = 0, n_pairs = 0, n_mpairs = 0;
  std::vector<std::vector<TcItem2>> hdrs;         // per segment: window headers
  std::vector<int> shape;                         // per segment: wh, ww, sy, sx
  std::vector<TcRec> stream_p[2], stream_m;       // one L-step, CTA pair after CTA pair
  std::vector<uint32_t> stream_off;               // [n_pairs][n_seg + 1]
  std::vector<uint2> eitems;                      // x = seg << 16 | window, y = row pair
  std::vector<uint32_t> eitem_off;                // [n_pairs][n_seg + 1]
  std::vector<uint32_t> dep_off, deps;
  std::vector<uint32_t> flag_base, n_windows;     // per segment
  uint32_t zflag_base = 0, n_flags = 0;
  long long n_steps = 0, n_mma = 0, n_bytes = 0;
};

#ifndef DGAN_COST_EPI_KB
#define DGAN_COST_EPI_KB 24.0
#endif
#ifndef DGAN_COST_FIXED_KB
#define DGAN_COST_FIXED_KB 48.0
#endif
#ifndef DGAN_LOOP_ORDER
#define DGAN_LOOP_ORDER 0        // experiment switch: 0 = row-pair-major item order inside a segment, 1 = LPT (cost-descending) order
#endif
#ifndef DGAN_LOOP_CARRY
#define DGAN_LOOP_CARRY 0        // experiment switch: 1 = a CTA pair's surplus load in one segment is deducted in the next
#endif
constexpr int LOOP_STEP_MAX_BYTES = 48 * 1024;    // measured optimum of the operand-ring kernels (round 1): 2 A tiles + weights

// Window tiling of one segment and the assignment of its (window, row pair) items to CTA pairs: every candidate shape
// (wh x ww accumulators, strides 1 or 2 - stride 2 gathers outputs of equal parity of a stride-2 transposed conv, which
// share weight tiles) is scored by a longest-processing-time assignment with cost = operand bytes staged + a
// per-accumulator epilogue charge + a fixed per-item charge; the smallest makespan wins.  `carry` (in/out) is each
// CTA pair's load imbalance inherited from the previous segment: there is no barrier between segments, so a pair that
// drew the short straw in one segment takes less of the next.
static int loop_assign(const LoopSegSpec& sp, int n_mpairs, int n_pairs, std::vector<double>* carry,
                       std::vector<Tc2HostItem>* items_out, std::vector<std::vector<int>>* lists_out, int shape_out[4]) {
  const int N = sp.N, K = sp.K;
  const int max_g = (N >= 64) ? std::min(4, 256 / N) : 1;
  const int step_max = std::min((LOOP_RING_BYTES / 2) & ~1023, LOOP_STEP_MAX_BYTES);
  double best_cost = 1e300;
  std::vector<Tc2HostItem> best_items;
  std::vector<std::vector<int>> best_lists;
  std::vector<double> best_load;
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
          std::stable_sort(items.begin(), items.end(), [](const Tc2HostItem& l, const Tc2HostItem& r) { return l.stage_bytes > r.stage_bytes; });
          std::vector<double> icost(items.size());
          for (size_t i = 0; i < items.size(); ++i)
            icost[i] = items[i].stage_bytes + DGAN_COST_EPI_KB * 1024.0 * items[i].hdr.n_acc * std::max(1, N / 64) + DGAN_COST_FIXED_KB * 1024.0;
          const long long total = (long long)items.size() * n_mpairs;
          std::vector<double> load = *carry;
          std::vector<std::vector<int>> lists((size_t)n_pairs);
          for (long long idx = 0; idx < total; ++idx) {        // items[] is sorted by cost, mp is the fast index: cost-descending
            size_t best = 0;
            for (size_t pr = 1; pr < (size_t)n_pairs; ++pr)
              if (load[pr] < load[best]) best = pr;
            load[best] += icost[(size_t)(idx / n_mpairs)];
            lists[best].push_back((int)idx);
          }
          const double makespan = *std::max_element(load.begin(), load.end());
          if (makespan < best_cost) {
            best_cost = makespan;
            shape_out[0] = wh; shape_out[1] = ww; shape_out[2] = sy; shape_out[3] = sx;
            best_items.swap(items); best_lists.swap(lists); best_load.swap(load);
          }
        }
  if (best_items.empty()) { set_error("no window tiling for segment " + sp.name); return DGAN_ERR_UNSUPPORTED; }
  const double lo = *std::min_element(best_load.begin(), best_load.end());
  for (size_t pr = 0; pr < best_load.size(); ++pr) (*carry)[pr] = best_load[pr] - lo;
  items_out->swap(best_items);
  lists_out->swap(best_lists);
  return 0;
}

// Plan one L-step for `n_mpairs` row pairs on `n_pairs` CTA pairs: per segment the window tiling + assignment, then
// per CTA pair the concatenated step stream (segment-major, row-pair-major inside a segment), the circular operand
// ring simulated CYCLICALLY (the stream is replayed L times, so a step's dependency distance may reach back into
// the previous replay), the item list of the epilogue warps and every item's dependency list.
static int loop_plan(const std::vector<LoopSegSpec>& specs, int n_mpairs, int n_pairs, bool carry_load, LoopPlan* plan) {
  const int n_seg = (int)specs.size();
  if (n_seg < 1 || n_seg > LOOP_MAX_SEG) { set_error("segment count out of range"); return DGAN_ERR_UNSUPPORTED; }
  if (n_mpairs < 1 || n_pairs < 1) { set_error("nothing to plan"); return DGAN_ERR_INVALID_ARG; }
  plan->n_seg = n_seg; plan->n_pairs = n_pairs; plan->n_mpairs = n_mpairs;
  plan->hdrs.assign((size_t)n_seg, {});
  plan->shape.assign((size_t)4 * n_seg, 1);
  plan->flag_base.assign((size_t)n_seg, 0);
  plan->n_windows.assign((size_t)n_seg, 0);
  std::vector<std::vector<Tc2HostItem>> items((size_t)n_seg);
  std::vector<std::vector<std::vector<int>>> lists((size_t)n_seg);
  std::vector<std::vector<int>> pix2win((size_t)n_seg);       // output pixel of a segment -> window that writes it
  std::vector<double> carry((size_t)n_pairs, 0.0);
  int rc;
  uint32_t n_flags = 0;
  for (int s = 0; s < n_seg; ++s) {
    const LoopSegSpec& sp = specs[(size_t)s];
    if (!carry_load) std::fill(carry.begin(), carry.end(), 0.0);
    if ((rc = loop_assign(sp, n_mpairs, n_pairs, &carry, &items[(size_t)s], &lists[(size_t)s], &plan->shape[(size_t)4 * s]))) return rc;
    const size_t nw = items[(size_t)s].size();
    plan->hdrs[(size_t)s].resize(nw);
    pix2win[(size_t)s].assign(sp.tab->off.size() - 1, -1);
    for (size_t w = 0; w < nw; ++w) {
      plan->hdrs[(size_t)s][w] = items[(size_t)s][w].hdr;
      for (uint32_t a = 0; a < items[(size_t)s][w].hdr.n_acc; ++a) pix2win[(size_t)s][items[(size_t)s][w].hdr.q[a]] = (int)w;
    }
    plan->flag_base[(size_t)s] = n_flags;
    plan->n_windows[(size_t)s] = (uint32_t)nw;
    n_flags += (uint32_t)(nw * (size_t)n_mpairs);
  }
  plan->zflag_base = n_flags;
  n_flags += (uint32_t)n_mpairs;
  plan->n_flags = n_flags;

  const int ring_kb = LOOP_RING_BYTES / 1024;
  plan->stream_p[0].clear(); plan->stream_p[1].clear(); plan->stream_m.clear();
  plan->stream_off.assign((size_t)n_pairs * (n_seg + 1), 0);
  plan->eitem_off.assign((size_t)n_pairs * (n_seg + 1), 0);
  plan->eitems.clear(); plan->dep_off.assign(1, 0); plan->deps.clear();
  plan->n_steps = plan->n_mma = plan->n_bytes = 0;
  for (int pr = 0; pr < n_pairs; ++pr) {
    const size_t stream_beg = plan->stream_m.size();
    size_t prev_first_rec = (size_t)-1;
    std::vector<int> kb_of;                                    // KB of every step of this pair's stream
    for (int s = 0; s < n_seg; ++s) {
      const LoopSegSpec& sp = specs[(size_t)s];
      plan->stream_off[(size_t)pr * (n_seg + 1) + s] = (uint32_t)plan->stream_m.size();
      plan->eitem_off[(size_t)pr * (n_seg + 1) + s] = (uint32_t)plan->eitems.size();
      std::vector<int> mine = lists[(size_t)s][(size_t)pr];
      // row-pair major: a pair meets the row pairs in the same order in every segment, so what it waits for was
      // produced a whole segment-phase ago; inside a row pair keep the cost-descending order
#if DGAN_LOOP_ORDER == 0
      std::stable_sort(mine.begin(), mine.end(), [&](int l, int r) { return (l % n_mpairs) < (r % n_mpairs); });
#endif
      const int half_b = (sp.N / 2) * 128;
      for (int idx : mine) {
        const int win = idx / n_mpairs, mp = idx % n_mpairs;
        const Tc2HostItem& itm = items[(size_t)s][(size_t)win];
        plan->eitems.push_back(make_uint2(((uint32_t)s << 16) | (uint32_t)win, (uint32_t)mp));
        // dependencies: the windows of the producing segment that cover the input pixels this item stages
        std::vector<uint32_t> dl;
        if (sp.in_seg < 0) {
          dl.push_back((plan->zflag_base + (uint32_t)mp) | LOOP_DEP_PREV);
        } else {
          const std::vector<int>& p2w = pix2win[(size_t)sp.in_seg];
          for (const Tc2HostStep& hs : itm.steps)
            for (int a = 0; a < hs.nA; ++a) {
              const int p = hs.a_pix[a];
              if (p < 0 || (size_t)p >= p2w.size() || p2w[(size_t)p] < 0) { set_error(sp.name + ": input pixel without a producer"); return DGAN_ERR_UNSUPPORTED; }
              const uint32_t f = plan->flag_base[(size_t)sp.in_seg] + (uint32_t)mp * plan->n_windows[(size_t)sp.in_seg] + (uint32_t)p2w[(size_t)p];
              if (std::find(dl.begin(), dl.end(), f) == dl.end()) dl.push_back(f);
            }
        }
        const uint32_t this_d0 = (uint32_t)plan->deps.size();
        plan->deps.insert(plan->deps.end(), dl.begin(), dl.end());
        plan->dep_off.push_back((uint32_t)plan->deps.size());
        // the first-step record of the PREVIOUS item of this CTA pair announces this item's dependency range, so that the
        // producer can fetch the flags one item ahead
        if (prev_first_rec != (size_t)-1)
          for (int r = 0; r < 2; ++r) { plan->stream_p[r][prev_first_rec].w[6] = this_d0; plan->stream_p[r][prev_first_rec].w[7] = (uint32_t)dl.size(); }
        prev_first_rec = plan->stream_m.size();
        for (size_t j = 0; j < itm.steps.size(); ++j) {
          const Tc2HostStep& hs = itm.steps[j];
          const int kb = (hs.bytes + 1023) / 1024;
          if (kb > ring_kb / 2) { set_error("tensor-core step larger than half the operand ring"); return DGAN_ERR_UNSUPPORTED; }
          kb_of.push_back(kb);
          const uint32_t flags = (j == 0 ? 1u : 0u) | (j + 1 == itm.steps.size() ? 2u : 0u);
          TcRec rm{};
          rm.w[0] = ((uint32_t)hs.nA << 8) | ((uint32_t)hs.n_ops << 11) | (flags << 16);   // ring offset filled below
          rm.w[1] = (uint32_t)s;
          for (int o = 0; o < hs.n_ops; ++o) rm.w[2 + o / 2] |= (uint32_t)hs.ops[o] << (16 * (o & 1));
          plan->stream_m.push_back(rm);
          for (int r = 0; r < 2; ++r) {
            TcRec rp{};
            rp.w[0] = ((uint32_t)hs.kc << 8) | ((uint32_t)hs.nA << 12) | ((uint32_t)hs.nB << 15);   // offset + dep filled below
            rp.w[1] = (uint32_t)mp | ((uint32_t)s << 16) | ((j == 0 ? 1u : 0u) << 20) | ((j + 1 == itm.steps.size() ? 1u : 0u) << 21);
            for (int a = 0; a < hs.nA; ++a) rp.w[2 + a / 2] |= (uint32_t)(hs.a_pix[a] & 0xFFFF) << (16 * (a & 1));
            for (int b = 0; b < hs.nB; ++b) rp.w[4 + b / 4] |= (uint32_t)hs.b_ent[r][b] << (8 * (b & 3));
            plan->stream_p[r].push_back(rp);
          }
          plan->n_mma += hs.n_ops; plan->n_steps += 1; plan->n_bytes += hs.bytes;
          (void)half_b;
        }
      }
    }
    plan->stream_off[(size_t)pr * (n_seg + 1) + n_seg] = (uint32_t)plan->stream_m.size();
    plan->eitem_off[(size_t)pr * (n_seg + 1) + n_seg] = (uint32_t)plan->eitems.size();
    // ---- circular operand ring of this CTA pair.  Sequential allocation, wrap when a step does not fit; every replay
    //      starts at offset 0 so that the records are the same for every L-step.
    const int n = (int)kb_of.size();
    std::vector<int> beg((size_t)n), end((size_t)n);
    int cursor = 0;
    for (int k = 0; k < n; ++k) {
      if (cursor + kb_of[(size_t)k] > ring_kb) cursor = 0;
      beg[(size_t)k] = cursor; end[(size_t)k] = cursor + kb_of[(size_t)k];
      cursor = end[(size_t)k];
    }
    // dep = distance (in steps, across the replay boundary if need be) to the latest earlier step whose region
    // overlaps this step's: the producer may overwrite the region once that step is consumed (8 = barrier-slot reuse only)
    for (int k = 0; k < n; ++k) {
      int dep = TC2_NSLOT;
      for (int d = 1; d < TC2_NSLOT; ++d) {
        const int c = ((k - d) % n + n) % n;
        if (beg[(size_t)c] < end[(size_t)k] && beg[(size_t)k] < end[(size_t)c]) { dep = d; break; }
      }
      const size_t ri = stream_beg + (size_t)k;
      plan->stream_m[ri].w[0] |= (uint32_t)beg[(size_t)k];
      for (int r = 0; r < 2; ++r) plan->stream_p[r][ri].w[0] |= (uint32_t)beg[(size_t)k] | ((uint32_t)dep << 19);
    }
  }
  return 0;
}

// Independent validation of a plan (host only; dgan_debug_check_plans and the CPU tests):
//  * per segment, everything tc2_check_plan re-derives from the records (every (output pixel, input pixel, tap, k-chunk)
//    contribution exactly once into the right accumulator, first-MMA flags, canonical accumulation order, every item
//    assigned exactly once);
//  * the operand ring, cyclically: a step's region overlaps none of the `dep - 1` steps before it (which may still be
//    unread when its loads start), also across the replay boundary;
//  * producer, MMA and epilogue streams agree on the item sequence (segment, row pair) of every CTA pair;
//  * every input pixel an item stages is covered by a dependency on the window (of the producing segment, same row pair)
//    that writes it, and segment 0 waits for the previous L-step's z update;
//  * no deadlock: with every CTA pair executing its items in order and an item startable only when its dependencies are
//    complete, all items of an L-step complete.
static int loop_check_plan(const std::vector<LoopSegSpec>& specs, const LoopPlan& pl, std::string* err) {
  auto fail = [&](const std::string& m) { *err = m; return DGAN_ERR_INVALID_ARG; };
  const int n_seg = pl.n_seg, n_pairs = pl.n_pairs, n_mpairs = pl.n_mpairs;
  if ((int)specs.size() != n_seg) return fail("segment count");
  if (pl.stream_off.size() != (size_t)n_pairs * (n_seg + 1) || pl.eitem_off.size() != pl.stream_off.size()) return fail("offset table size");
  if (pl.stream_p[0].size() != pl.stream_m.size() || pl.stream_p[1].size() != pl.stream_m.size()) return fail("stream sizes differ");
  if (pl.dep_off.size() != pl.eitems.size() + 1) return fail("dep_off size");
  // ---- per segment: slice the streams and reuse the single-layer validator
  for (int s = 0; s < n_seg; ++s) {
    const LoopSegSpec& sp = specs[(size_t)s];
    Tc2Plan sub;
    sub.n_pairs = n_pairs;
    sub.hdrs = pl.hdrs[(size_t)s];
    sub.stream_off.assign((size_t)n_pairs + 1, 0);
    size_t n_slots = 0;
    for (int pr = 0; pr < n_pairs; ++pr) {
      const uint32_t e0 = pl.eitem_off[(size_t)pr * (n_seg + 1) + s], e1 = pl.eitem_off[(size_t)pr * (n_seg + 1) + s + 1];
      if (e0 > e1 || e1 > pl.eitems.size()) return fail("item offsets not monotone");
      n_slots = std::max(n_slots, (size_t)(e1 - e0));
    }
    sub.n_slots = (int)n_slots;
    sub.eitems.assign(n_slots * (size_t)n_pairs, -1);
    for (int pr = 0; pr < n_pairs; ++pr) {
      const uint32_t r0 = pl.stream_off[(size_t)pr * (n_seg + 1) + s], r1 = pl.stream_off[(size_t)pr * (n_seg + 1) + s + 1];
      if (r0 > r1 || r1 > pl.stream_m.size()) return fail("stream offsets not monotone");
      sub.stream_off[(size_t)pr] = (uint32_t)sub.stream_m.size();
      for (uint32_t ri = r0; ri < r1; ++ri) {
        TcRec m = pl.stream_m[ri], p0 = pl.stream_p[0][ri], p1 = pl.stream_p[1][ri];
        if ((int)(m.w[1] & 0xF) != s || (int)((p0.w[1] >> 16) & 0xF) != s || (int)((p1.w[1] >> 16) & 0xF) != s) return fail(sp.name + ": record carries another segment's id");
        if (((p0.w[1] >> 20) & 1u) != ((m.w[0] >> 16) & 1u) || p0.w[1] != p1.w[1]) return fail(sp.name + ": first-step marks of producer and MMA records disagree");
        p0.w[1] &= 0xFFFFu; p1.w[1] &= 0xFFFFu; m.w[1] = 0;
        sub.stream_m.push_back(m); sub.stream_p[0].push_back(p0); sub.stream_p[1].push_back(p1);
      }
      const uint32_t e0 = pl.eitem_off[(size_t)pr * (n_seg + 1) + s], e1 = pl.eitem_off[(size_t)pr * (n_seg + 1) + s + 1];
      for (uint32_t e = e0; e < e1; ++e) {
        const uint2 it = pl.eitems[e];
        if ((int)(it.x >> 16) != s) return fail(sp.name + ": item list out of segment order");
        if ((it.x & 0xFFFFu) > 0x7FFFu || it.y > 0xFFFFu) return fail(sp.name + ": item index too large");
        sub.eitems[(size_t)(e - e0) * n_pairs + pr] = (int)(((it.x & 0xFFFFu) << 16) | it.y);
      }
    }
    sub.stream_off[(size_t)n_pairs] = (uint32_t)sub.stream_m.size();
    std::string e2;
    if (tc2_check_plan(sp.N, sp.K, *sp.tab, n_mpairs, LOOP_RING_BYTES, sub, &e2)) return fail(sp.name + ": " + e2);
  }
  // ---- cyclic ring check + dependency coverage, per CTA pair
  const int ring_kb = LOOP_RING_BYTES / 1024;
  std::vector<std::vector<int>> pix2win((size_t)n_seg);
  for (int s = 0; s < n_seg; ++s) {
    pix2win[(size_t)s].assign(specs[(size_t)s].tab->off.size() - 1, -1);
    for (size_t w = 0; w < pl.hdrs[(size_t)s].size(); ++w)
      for (uint32_t a = 0; a < pl.hdrs[(size_t)s][w].n_acc; ++a) pix2win[(size_t)s][pl.hdrs[(size_t)s][w].q[a]] = (int)w;
  }
  for (int pr = 0; pr < n_pairs; ++pr) {
    const uint32_t r0 = pl.stream_off[(size_t)pr * (n_seg + 1)], r1 = pl.stream_off[(size_t)pr * (n_seg + 1) + n_seg];
    const int n = (int)(r1 - r0);
    std::vector<int> beg((size_t)n), end((size_t)n), dep((size_t)n);
    for (int k = 0; k < n; ++k) {
      const TcRec& p0 = pl.stream_p[0][r0 + k];
      const int s = (int)((p0.w[1] >> 16) & 0xF);
      const int nA = (int)((p0.w[0] >> 12) & 7), nB = (int)((p0.w[0] >> 15) & 0xF);
      beg[(size_t)k] = (int)(p0.w[0] & 0xFF);
      end[(size_t)k] = beg[(size_t)k] + (nA * TC_A_BYTES + nB * (specs[(size_t)s].N / 2) * 128 + 1023) / 1024;
      dep[(size_t)k] = (int)((p0.w[0] >> 19) & 0xF);
      if (end[(size_t)k] > ring_kb) return fail("step region outside the ring");
      if (dep[(size_t)k] < 1 || dep[(size_t)k] > TC2_NSLOT) return fail("dep out of range");
    }
    for (int k = 0; k < n; ++k)
      for (int d = 1; d < dep[(size_t)k]; ++d) {
        const int c = ((k - d) % n + n) % n;
        if (beg[(size_t)c] < end[(size_t)k] && beg[(size_t)k] < end[(size_t)c])
          return fail("ring hazard: a region may be overwritten while it can still be read (cyclic check)");
      }
    // dependency coverage: walk the items of the pair in stream order
    uint32_t e = pl.eitem_off[(size_t)pr * (n_seg + 1)];
    std::vector<uint32_t> need;
    for (int k = 0; k <= n; ++k) {
      const bool first = (k < n) && ((pl.stream_p[0][r0 + k].w[1] >> 20) & 1u);
      if ((first || k == n) && k > 0) {
        // close the previous item: compare with its dependency list
        const uint32_t d0 = pl.dep_off[e], d1 = pl.dep_off[e + 1];
        for (uint32_t f : need)
          if (std::find(pl.deps.begin() + d0, pl.deps.begin() + d1, f) == pl.deps.begin() + d1) return fail("an item stages a pixel it has no dependency on");
        for (uint32_t d = d0; d < d1; ++d)
          if ((pl.deps[d] & ~LOOP_DEP_PREV) >= pl.n_flags) return fail("dependency index out of range");
        need.clear();
        ++e;
      }
      if (k == n) break;
      const TcRec& p0 = pl.stream_p[0][r0 + k];
      const int s = (int)((p0.w[1] >> 16) & 0xF), mp = (int)(p0.w[1] & 0xFFFF), nA = (int)((p0.w[0] >> 12) & 7);
      if (first) {
        if (e >= pl.eitems.size() || (int)(pl.eitems[e].x >> 16) != s || (int)pl.eitems[e].y != mp) return fail("producer stream and item list disagree");
        // the record announces the NEXT item's dependency range (none after the pair's last item)
        const uint32_t lim = pl.eitem_off[(size_t)pr * (n_seg + 1) + n_seg];
        const uint32_t want_d0 = (e + 1 < lim) ? pl.dep_off[e + 1] : 0u, want_cnt = (e + 1 < lim) ? pl.dep_off[e + 2] - pl.dep_off[e + 1] : 0u;
        for (int r = 0; r < 2; ++r) {
          const TcRec& pq = pl.stream_p[r][r0 + k];
          if (pq.w[7] != want_cnt || (want_cnt != 0 && pq.w[6] != want_d0)) return fail("a record announces the wrong dependency range for the next item");
        }
      }
      if ((((p0.w[1] >> 21) & 1u) != 0u) != (((pl.stream_m[r0 + k].w[0] >> 17) & 1u) != 0u)) return fail("last-step marks of producer and MMA records disagree");
      const int in_seg = specs[(size_t)s].in_seg;
      if (in_seg < 0) {
        const uint32_t f = (pl.zflag_base + (uint32_t)mp) | LOOP_DEP_PREV;
        if (std::find(need.begin(), need.end(), f) == need.end()) need.push_back(f);
      } else {
        if (in_seg >= s) return fail("a segment reads a later segment's output");
        for (int a = 0; a < nA; ++a) {
          const int p = (int)((p0.w[2 + a / 2] >> (16 * (a & 1))) & 0xFFFF);
          if ((size_t)p >= pix2win[(size_t)in_seg].size() || pix2win[(size_t)in_seg][(size_t)p] < 0) return fail("staged pixel has no producing window");
          const uint32_t f = pl.flag_base[(size_t)in_seg] + (uint32_t)mp * pl.n_windows[(size_t)in_seg] + (uint32_t)pix2win[(size_t)in_seg][(size_t)p];
          if (std::find(need.begin(), need.end(), f) == need.end()) need.push_back(f);
        }
      }
    }
    if (e != pl.eitem_off[(size_t)pr * (n_seg + 1) + n_seg]) return fail("item count of a pair's stream and its item list differ");
  }
  // ---- deadlock check: in-order execution per CTA pair, an item starts only when its dependencies are complete
  {
    std::vector<char> done(pl.n_flags, 0);
    std::vector<uint32_t> head((size_t)n_pairs);
    for (int pr = 0; pr < n_pairs; ++pr) head[(size_t)pr] = pl.eitem_off[(size_t)pr * (n_seg + 1)];
    size_t remaining = pl.eitems.size();
    bool progress = true;
    while (remaining > 0 && progress) {
      progress = false;
      for (int pr = 0; pr < n_pairs; ++pr) {
        const uint32_t lim = pl.eitem_off[(size_t)pr * (n_seg + 1) + n_seg];
        while (head[(size_t)pr] < lim) {
          const uint32_t e = head[(size_t)pr];
          bool ready = true;
          for (uint32_t d = pl.dep_off[e]; d < pl.dep_off[e + 1] && ready; ++d)
            if (!(pl.deps[d] & LOOP_DEP_PREV) && !done[pl.deps[d]]) ready = false;
          if (!ready) break;
          const uint2 it = pl.eitems[e];
          const int s = (int)(it.x >> 16);
          const uint32_t f = pl.flag_base[(size_t)s] + it.y * pl.n_windows[(size_t)s] + (it.x & 0xFFFFu);
          if (f >= pl.zflag_base || done[f]) return fail("item flag out of range or completed twice");
          done[f] = 1;
          ++head[(size_t)pr]; --remaining; progress = true;
        }
      }
    }
    if (remaining > 0) return fail("deadlock: some items can never start (dependency cycle across in-order streams)");
    for (uint32_t f = 0; f < pl.zflag_base; ++f)
      if (!done[f]) return fail("an item is missing from the L-step");
  }
  return 0;
}

}  // namespace dgan
