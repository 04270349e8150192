// The projection loop as ONE persistent kernel (DGAN_PREC_FP16) with dataflow scheduling.
//
// Round 1 ran every layer-direction of an L-step as its own persistent tcgen05 kernel: 9 launches per step, 1798 per
// call, each paying ~2 us of launch gap, ~1.5 us until its first operands landed and 2-6 us of un-overlapped last
// epilogue - a quarter of the step.  Here the whole call - L x (generator forward, loss, backward-to-z, momentum) -
// is one launch of 74 co-resident CTA pairs (one per TPC).
//
// Work is cut into ITEMS: (segment = layer-direction, window of 1-8 output pixels, row pair of 256 latent rows, L-step).
// The steps of an item (what to stage, which MMAs to issue) are the same host-planned records as in round 1, so every
// accumulator still sums its contributions in one canonical order and the bits do not depend on the schedule.  What is
// new is WHO runs an item and WHEN: items form a dataflow graph (an item needs the windows of the previous segment that
// cover the pixels it stages - same row pair, rows are independent - and the first segment needs the row pair's momentum
// update of the previous L-step).  A global ready queue holds the items whose inputs are complete; a CTA pair that
// finishes issuing an item pops the next one.  When the stores of an item have completed, the warp that observes the last
// of them bumps the arrival counter of each successor and pushes those that became ready.  Static per-CTA streams with
// flags (the first version of this kernel, profiles/r2_loop_trace_1group.md) lost 30 % of the time to head-of-line
// blocking: an in-order stream cannot cover the publish latency at segment and L-step boundaries, and at 10 row pairs
// there is no static order that gives every dependency an independent separator.  With the queue, row pairs drift apart by
// themselves and a CTA pair only idles when nothing at all is ready.
//
// Taking an item costs a chain of global round trips (queue index, queue slot, record offsets - about 2.5 us).  It is
// paid by a FETCHER warp (the peer CTA's otherwise idle warp 1), which runs up to LOOP_AHEAD items ahead of the pair's TMA
// producer and hands items - with their record range - to every role of both CTAs through a shared-memory mailbox; the
// producer and the MMA issuer peek one mailbox entry ahead to have the next item's first records in registers when the
// current item ends.
//
// Deadlock freedom: only ready items are ever taken, an item's completion never waits on anything but its own stores, and
// the kernel ends through sentinels pushed when the last item has completed; the grid is sized to the co-resident
// clusters (a CTA pair that is never scheduled would only leave items to the others).  Queue pops have a time-out that
// raises a status word instead of hanging.
#pragma once
#include "kernels_tc2.cuh"

namespace dgan {

constexpr int LOOP_MAX_SEG = 10;
// Warp roles, by warpgroup so that registers can be re-balanced with setmaxnreg: warpgroup 0 = TMA producer (warp 0),
// MMA issuer (warp 1) and the two store warps, trimmed to LOOP_REGS_CTRL registers; warpgroups 1-2 = the 8 epilogue
// warps, raised to LOOP_REGS_EPI (the epilogue holds two 32-column TMEM loads in flight while it converts a 64-column
// unit: at the launch-time 168 registers it spilled about a kilobyte per thread).
constexpr int LOOP_THREADS = 128 + 32 * TC2_EPI_WARPS;
constexpr int LOOP_EPI_WARP0 = 4;
constexpr int LOOP_REGS_CTRL = 104, LOOP_REGS_EPI = 200;     // 128*104 + 256*200 = 384*168 (the launch-time pool)
constexpr int LOOP_EPI_TILES = 2;                                              // one output staging tile per epilogue half
constexpr int LOOP_RING_BYTES = tc2_ring_bytes(64, EPI_BIAS_RELU, 2);           // operand ring next to those tiles
constexpr int LOOP_BAR_BYTES = 512;                                            // mbarriers + mailbox
constexpr int LOOP_SMEM_BYTES = LOOP_RING_BYTES + LOOP_EPI_TILES * TC2_TILE_BYTES + TC2_STAGING_BYTES + 1024 + LOOP_BAR_BYTES;
static_assert(LOOP_SMEM_BYTES <= TC2_SMEM_MAX, "shared memory budget");
constexpr int LOOP_MAIL = 8;                                // mailbox slots: items the fetcher may run ahead of the pair's slowest role
constexpr int LOOP_AHEAD = 2;                               // items the fetcher may hold beyond the one the producer is issuing - while
                                                            // the queue has a backlog; when other pairs are waiting for work it only
                                                            // takes the next item once the producer has issued the current one
constexpr uint32_t LOOP_SENTINEL = 0x000F0000u;             // queue entry (segment 15) that ends a CTA pair
constexpr uint32_t LOOP_LAP_SHIFT = 20, LOOP_LAP_MASK = 0x7FFu;   // entries carry the lap of their queue index (bits 20..30 of lo)
constexpr unsigned long long LOOP_EMPTY = ~0ull;            // queue slot not written yet (lap field 0xFFF matches no lap)

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
  const TcItem2* items;            // window headers
  uint32_t n_tile, kind, bias_pstride, acc_stride;
  uint32_t idesc, half_b;
  uint32_t win_base;               // index of window 0 in the per-window tables (record offsets, successors, needs)
  uint32_t item_base, n_windows;   // item (window w, row pair mp) = item_base + mp * n_windows + w in the counter arrays
};

struct LoopParams {
  LoopSeg seg[LOOP_MAX_SEG];
  const TcRec* tmpl_p[2];          // producer step records per cluster rank: per (segment, window), windows back to back
  const TcRec* tmpl_m;             // MMA step records, same indexing
  const uint32_t* win_rec_off;     // [windows + 1] record offsets
  const uint32_t* succ_off;        // [windows + 1] -> succ
  const uint32_t* succ;            // successors of a window: (segment << 16 | window) in the consuming segment, same row pair
  const uint32_t* need;            // [windows] completions of producing items a window waits for, per L-step
  unsigned long long* queue;       // ready queue: lo = lap << 20 | segment << 16 | window, hi = row pair | L-step << 16
  uint32_t* q_ctl;                 // [0] popped, [1] pushed, [2] completed items
  uint32_t* depcnt;                // per item: completed parts of the items it waits for
  uint32_t* status;                // [0] != 0: a queue pop timed out (results invalid)
  unsigned long long* prof;        // optional [L-step][n_seg][2] globaltimer min-start / max-end of the epilogues
  unsigned long long* dbg;         // optional [CTA][16] stall counters of the roles (clock64 ticks), see LoopDbg
  unsigned long long* trace;       // optional [item slot][4]: pop time | CTA pair << 48, accumulator granted, epilogue begin, epilogue end (L-step trace_step)
  uint32_t q_cap, q_shift, n_pairs; // queue capacity = 1 << q_shift
  uint32_t q_init;                 // entries the host placed in the queue (the first segment's items of L-step 0): pushes start behind them
  int trace_step;
  uint32_t n_parts_total;          // item parts of the whole launch: the completion of the last one pushes the sentinels
  int n_seg, n_fwd;                // segments; the first n_fwd are the generator forward (+ loss)
  int last_step;                   // index of the call's final L-step (rec_iters - 1): its forward writes G(z) and the loss
  int full_last;                   // 1: the final L-step also runs its backward half (dgan_loss_grad); 0: forward only
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
};

namespace ptx {
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_acq_rel_gpu(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_shared_cluster_u32(uint32_t local_addr, uint32_t cta, uint32_t v) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "st.shared::cluster.b32 [ra], %2;\n\t}" ::"r"(local_addr), "r"(cta), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t local_bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(local_bar), "r"(cta) : "memory");
}
__device__ __forceinline__ bool mbar_test_cluster(uint32_t bar, uint32_t parity) {   // non-blocking, acquire at cluster scope
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint32_t ld_shared_volatile_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope (peer's writes)
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAITC_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAITC_DONE;\n\t"
      "bra WAITC_LOOP;\n\t"
      "WAITC_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
}  // namespace ptx

// "this thread's global writes of the item are done": order them before the completion counters, for readers in both proxies
__device__ __forceinline__ void loop_publish_fence() {
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  ptx::fence_proxy_async_all();
}

// per-CTA stall counters written when LoopParams::dbg != NULL (developer aid: tools/loop_stalls.py)
enum LoopDbg : int { DBG_P_POP = 0, DBG_P_RING, DBG_P_TOTAL, DBG_M_FULL, DBG_M_ACC, DBG_M_TOTAL, DBG_E_ACC, DBG_E_TILE, DBG_E_TOTAL,
                     DBG_S_TILE, DBG_S_DONE, DBG_S_TOTAL, DBG_P_ITEMS, DBG_M_MAIL, DBG_F_QUEUE, DBG_F_CREDIT, DBG_COUNT = 16 };

// shared-memory control block of a CTA (offsets from bar_base)
constexpr uint32_t LB_FULL = 0, LB_EMPTY = 64, LB_ACC_FULL = 128, LB_ACC_EMPTY = 144, LB_TMEM = 160, LB_TILE_FULL = 168,
                   LB_TILE_FREE = 184, LB_MOM = 200, LB_CREDIT = 204, LB_MAIL_FULL = 208, LB_MAIL_EMPTY = 272, LB_MAIL_DATA = 384;
static_assert(LB_MAIL_DATA + 16 * LOOP_MAIL <= LOOP_BAR_BYTES, "control block");

struct LoopItem { uint32_t seg, win, mp, t, rbeg, rend; };    // rbeg..rend: the window's step records
__device__ __forceinline__ LoopItem loop_unpack(uint32_t lo, uint32_t hi) {
  LoopItem it;
  it.seg = (lo >> 16) & 0xFu; it.win = lo & 0xFFFFu; it.mp = hi & 0xFFFFu; it.t = hi >> 16;
  it.rbeg = it.rend = 0;
  return it;
}

// Append a ready item to the queue.
__device__ __forceinline__ void loop_push(const LoopParams& P, uint32_t lo, uint32_t hi) {
  const uint32_t idx = atomicAdd(P.q_ctl + 1, 1u) + P.q_init;
  const uint32_t lap = (idx >> P.q_shift) & LOOP_LAP_MASK;
  ptx::st_release_gpu_u64(P.queue + (idx & (P.q_cap - 1u)), ((unsigned long long)hi << 32) | lo | (lap << LOOP_LAP_SHIFT));
}

// Successors of a window, fetched ahead of the item's completion (one per lane; windows with more than 32 successors
// loop in loop_complete): everything loop_complete needs from global memory that does not depend on the stores.
struct LoopSucc { uint32_t s0, s1, e, need; };
__device__ __forceinline__ LoopSucc loop_succ_prefetch(const LoopParams& P, const LoopSeg& sg, uint32_t win, int lane) {
  LoopSucc q;
  const uint32_t wi = sg.win_base + win;
  q.s0 = __ldg(P.succ_off + wi); q.s1 = __ldg(P.succ_off + wi + 1);
  q.e = 0; q.need = 0;
  if (q.s0 + (uint32_t)lane < q.s1) {
    q.e = __ldg(P.succ + q.s0 + lane);
    q.need = __ldg(P.need + P.seg[q.e >> 16].win_base + (q.e & 0xFFFFu));
  }
  return q;
}

// One PART of the item (seg, win, mp) of L-step t has completed (this warp's share of its stores is in global memory): an
// item has 4 parts (the store warps of both CTAs) when its epilogue stores by TMA, else 16 (the epilogue warps).  Every part
// bumps the successors' counters itself - `need` is in parts - so the last part to finish wakes them without a second
// round trip through an arrival counter.  Called by a converged warp.
__device__ __forceinline__ void loop_complete(const LoopParams& P, uint32_t seg, uint32_t mp, uint32_t t, int lane, const LoopSucc q) {
  // the backward half of the final L-step is never run: the loop returns the pre-update forward (models/gan.py:419-421)
  const bool stop = ((int)seg == P.n_fwd - 1) && ((int)t == P.last_step) && !P.full_last;
  if (!stop)
    for (uint32_t s = q.s0 + (uint32_t)lane; s < q.s1; s += 32) {
      uint32_t e = q.e, need = q.need;
      if (s >= q.s0 + 32u) { e = __ldg(P.succ + s); need = __ldg(P.need + P.seg[e >> 16].win_base + (e & 0xFFFFu)); }
      const LoopSeg& sn = P.seg[e >> 16];
      const uint32_t c = ptx::atom_add_acq_rel_gpu(P.depcnt + sn.item_base + mp * sn.n_windows + (e & 0xFFFFu), 1u) + 1u;
      if (c == need * (t + 1u)) loop_push(P, e, mp | (t << 16));
    }
  __syncwarp();
  if (lane == 0) {
    const uint32_t d = atomicAdd(P.q_ctl + 2, 1u) + 1u;
    if (d == P.n_parts_total)
      for (uint32_t k = 0; k < P.n_pairs; ++k) loop_push(P, LOOP_SENTINEL, 0u);
  }
}

struct LoopCtx {                     // per-thread view of the CTA's pipeline state handed to the epilogue variants
  uint32_t tmem_base, bar_base, epi_base;
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
// parameterised at run time by the segment); ends by releasing the accumulator buffer and reporting completion.
// ------------------------------------------------------------------------------------------
template <int N_TILE, int EPI, typename TOUT>
__device__ __forceinline__ void loop_epilogue_item(const LoopParams& P, const LoopSeg& sg, LoopCtx& cx, const TcFinalArgs& fa,
                                                   const LoopItem it) {
  const int win = (int)it.win, mp = (int)it.mp;
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
    ptx::mbar_wait(cx.bar_base + LB_ACC_FULL + 8 * buf, (cx.item_count >> 1) & 1);
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
    constexpr int G = N_TILE >= 64 ? N_TILE / 64 : 1;   // 64-column groups per accumulator
    const int n_units = n_acc * G;
    // The staging tile of this epilogue half is handed to the half's STORE WARP (warp 2 + half): it issues the TMA store,
    // frees the tile when the store has read it, and reports the item's completion when its stores have completed - so
    // no thread that does arithmetic ever waits for global-memory latency.
    const uint32_t tile_full = cx.bar_base + LB_TILE_FULL + 8 * (uint32_t)half, tile_free = cx.bar_base + LB_TILE_FREE + 8 * (uint32_t)half;
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
  if (lane == 0) ptx::mbar_arrive_remote(cx.bar_base + LB_ACC_EMPTY + 8 * buf, 0);

  // ---- completion
  if (TMA_EPI) {
    // reported by the half's store warp once the item's tile stores have completed
  } else {
    if (EPI == EPI_NONE && sizeof(TOUT) == 4 && P.m_counter != nullptr) {
      // ---- momentum in the tail of the split-K Linear backward (tf.train.MomentumOptimizer, models/gan.py:389-391).
      //      Every epilogue thread has stored its share of this item's partial sums; the CTA that completes the last
      //      partial of its 128-row tile applies v <- mu v + g, z <- z - lr v (parts summed in the fixed order 0, 1, 2, ...)
      //      and releases the row pair's next L-step.
      const uint32_t flag_addr = cx.bar_base + LB_MOM;
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
        constexpr int STRIDE = 4 * 32 * TC2_EPI_WARPS, UNR = 2, MAXP = TC_LINEAR_SPLIT;
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
        if (warp == LOOP_EPI_WARP0) {
          if (lane == 0) P.m_counter[rt] = 0u;                 // ready for the next L-step's tickets
          // z of this 128-row tile is updated: the row pair's next Linear forward waits for both tiles
          if ((int)it.t < P.last_step) {
            const LoopSeg& s0g = P.seg[0];
            for (uint32_t w2 = (uint32_t)lane; w2 < s0g.n_windows; w2 += 32) {
              const uint32_t c = ptx::atom_add_acq_rel_gpu(P.depcnt + s0g.item_base + (uint32_t)mp * s0g.n_windows + w2, 1u) + 1u;
              if (c == __ldg(P.need + s0g.win_base + w2) * (it.t + 1u)) loop_push(P, w2, (uint32_t)mp | ((it.t + 1u) << 16));
            }
          }
        }
      }
    }
    // every epilogue warp of both CTAs reports its share of the item; the last one to do so wakes the successors
    const LoopSucc sq = loop_succ_prefetch(P, sg, it.win, lane);
    loop_publish_fence();
    __syncwarp();
    loop_complete(P, it.seg, it.mp, it.t, lane, sq);
  }
}

template <int ARCH>
__device__ __forceinline__ void loop_epilogue_dispatch(const LoopParams& P, const LoopSeg& sg, LoopCtx& cx, const TcFinalArgs& fa,
                                                       const LoopItem it) {
  switch (sg.kind) {
    case LK_BR256: loop_epilogue_item<256, EPI_BIAS_RELU, __half>(P, sg, cx, fa, it); break;
    case LK_BR128: loop_epilogue_item<128, EPI_BIAS_RELU, __half>(P, sg, cx, fa, it); break;
    case LK_BR64: loop_epilogue_item<64, EPI_BIAS_RELU, __half>(P, sg, cx, fa, it); break;
    case LK_MASK64: loop_epilogue_item<64, EPI_MASK, __half>(P, sg, cx, fa, it); break;
    case LK_MASK128: loop_epilogue_item<128, EPI_MASK, __half>(P, sg, cx, fa, it); break;
    case LK_MASK256: loop_epilogue_item<256, EPI_MASK, __half>(P, sg, cx, fa, it); break;
    case LK_NONE64F: loop_epilogue_item<64, EPI_NONE, float>(P, sg, cx, fa, it); break;
    case LK_NONE128F: loop_epilogue_item<128, EPI_NONE, float>(P, sg, cx, fa, it); break;
    case LK_NONE256F: loop_epilogue_item<256, EPI_NONE, float>(P, sg, cx, fa, it); break;
    case LK_B64: if (ARCH == DGAN_ARCH_CELEBA) loop_epilogue_item<64, EPI_BIAS, __half>(P, sg, cx, fa, it); break;
    case LK_NONE64H: if (ARCH == DGAN_ARCH_CELEBA) loop_epilogue_item<64, EPI_NONE, __half>(P, sg, cx, fa, it); break;
    case LK_FINAL48: if (ARCH == DGAN_ARCH_CELEBA) loop_epilogue_item<48, EPI_FINAL_TANH3, __half>(P, sg, cx, fa, it); break;
    case LK_FINAL16: if (ARCH == DGAN_ARCH_MNIST) loop_epilogue_item<16, EPI_FINAL_SIGMOID1, __half>(P, sg, cx, fa, it); break;
    default: break;
  }
}

// The circular operand ring, allocated at run time: steps take consecutive regions (wrapping when one does not fit); a
// region may be overwritten once the latest earlier step that overlaps it has been consumed.  Producer warps of both CTAs
// and the MMA warp run this same function on the same step sequence, so they agree on every offset.  Lanes 0..7 hold the
// regions of the last 8 steps (barrier slots).  Returns the step's offset in KB; `dep` = distance to the step to wait for
// (8 = only the barrier slot's previous user).
struct LoopRing {
  uint32_t cursor = 0, beg_l = 0, end_l = 0;     // beg_l / end_l: this lane's slot (lanes 0..7)
};
__device__ __forceinline__ uint32_t loop_ring_alloc(LoopRing& r, uint32_t it, uint32_t kb, int lane, uint32_t* dep) {
  constexpr uint32_t RING_KB = LOOP_RING_BYTES / 1024;
  if (r.cursor + kb > RING_KB) r.cursor = 0;
  const uint32_t beg = r.cursor, end = beg + kb;
  r.cursor = end;
  const uint32_t slot = it & (TC2_NSLOT - 1);
  // slot j holds step it - ((slot - j) & 7) (j != slot) - valid if that step exists
  const uint32_t dist = (slot - (uint32_t)lane) & (TC2_NSLOT - 1);
  const bool overlap = lane < TC2_NSLOT && dist != 0 && dist <= it && r.beg_l < end && beg < r.end_l;
  const uint32_t mask = __ballot_sync(0xffffffffu, overlap);
  uint32_t d = TC2_NSLOT;
#pragma unroll
  for (uint32_t k = TC2_NSLOT - 1; k >= 1; --k)
    if ((mask >> ((slot - k) & (TC2_NSLOT - 1))) & 1u) d = k;
  *dep = d;
  if ((uint32_t)lane == slot) { r.beg_l = beg; r.end_l = end; }
  return beg;
}

template <int ARCH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LOOP_THREADS, 1)
projection_loop_kernel(const __grid_constant__ LoopParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + LOOP_RING_BYTES;                       // two output staging tiles
  const uint32_t stg_base = epi_base + LOOP_EPI_TILES * TC2_TILE_BYTES;        // [producer ring][MMA ring] of TcRec
  const uint32_t bar_base = stg_base + TC2_STAGING_BYTES;                      // control block, LB_* offsets
  const uint32_t bar_full = bar_base + LB_FULL, bar_empty = bar_base + LB_EMPTY;
  const uint32_t bar_acc_full = bar_base + LB_ACC_FULL, bar_acc_empty = bar_base + LB_ACC_EMPTY;
  const uint32_t tmem_slot = bar_base + LB_TMEM;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < TC2_NSLOT; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);    // leader's producer arrive.expect_tx (bytes of both CTAs)
      ptx::mbar_init(bar_empty + 8 * s, 1);   // one multicast commit per CTA
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_acc_full + 8 * b, 1);
      ptx::mbar_init(bar_acc_empty + 8 * b, 2 * TC2_EPI_WARPS);   // epilogue warps of both CTAs (used on the leader only)
      ptx::mbar_init(bar_base + LB_TILE_FULL + 8 * b, 128);       // the 4 warps of an epilogue half
      ptx::mbar_init(bar_base + LB_TILE_FREE + 8 * b, 1);         // the half's store warp
    }
    for (int m = 0; m < LOOP_MAIL; ++m) {
      ptx::mbar_init(bar_base + LB_MAIL_FULL + 8 * m, 1);         // the fetcher, once per item (in both CTAs)
      // readers of a mailbox slot, both CTAs (used on the peer CTA, where the fetcher lives): producer + 2 store warps +
      // 8 epilogue warps of each CTA, and the leader's MMA warp
      ptx::mbar_init(bar_base + LB_MAIL_EMPTY + 8 * m, 2 * (3 + TC2_EPI_WARPS) + 1);
    }
    ptx::st_shared_u32(bar_base + LB_CREDIT, 0u);
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

  // Every role of both CTAs learns the pair's k-th item from mailbox slot k % LOOP_MAIL (written by the fetcher, which lives
  // in the peer CTA: the leader's roles acquire at cluster scope).
  auto mail_take = [&](uint32_t k) -> LoopItem {                   // the slot is known to be full
    const uint32_t m = k & (LOOP_MAIL - 1);
    const uint4 d = ptx::ld_shared_v4(bar_base + LB_MAIL_DATA + 16 * m);
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive_remote(bar_base + LB_MAIL_EMPTY + 8 * m, 1);      // this warp is done with the slot
    LoopItem it = loop_unpack(d.x, d.y);
    it.rbeg = d.z; it.rend = d.w;
    if ((d.x & 0xF0000u) == LOOP_SENTINEL) it.seg = 0xFFFFu;
    return it;
  };
  auto mail_read = [&](uint32_t k) -> LoopItem {
    const uint32_t m = k & (LOOP_MAIL - 1), par = (k / LOOP_MAIL) & 1;
    if (leader) ptx::mbar_wait_cluster(bar_base + LB_MAIL_FULL + 8 * m, par);
    else ptx::mbar_wait(bar_base + LB_MAIL_FULL + 8 * m, par);
    return mail_take(k);
  };
  auto mail_ready = [&](uint32_t k) -> bool {                      // non-blocking
    return ptx::mbar_test_cluster(bar_base + LB_MAIL_FULL + 8 * (k & (LOOP_MAIL - 1)), (k / LOOP_MAIL) & 1);
  };

  if (warp < LOOP_EPI_WARP0) {
   ptx::setmaxnreg_dec<LOOP_REGS_CTRL>();
   if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const TcRec* __restrict__ stream = rank ? P.tmpl_p[1] : P.tmpl_p[0];
    const uint32_t ring = stg_base;
    uint32_t it = 0;                                      // steps issued so far (barrier slot / phase)
    LoopRing rs;
    long long t_pop = 0, t_ring = 0, n_items = 0;
    const long long t_p0 = P.dbg ? clock64() : 0;
    LoopItem cur, nxt;
    uint4 mine = make_uint4(0, 0, 0, 0), mine_nxt = make_uint4(0, 0, 0, 0);
    bool have_nxt = false;
    {
      const long long tw0 = P.dbg ? clock64() : 0;
      cur = mail_read(0);
      if (P.dbg) t_pop += clock64() - tw0;
      if (cur.seg != 0xFFFFu && 2 * cur.rbeg + lane < 2 * cur.rend) mine = __ldg(reinterpret_cast<const uint4*>(stream + cur.rbeg) + lane);
    }
    for (uint32_t k = 0; cur.seg != 0xFFFFu; ++k) {
      if (leader && lane == 0) ptx::st_shared_cluster_u32(bar_base + LB_CREDIT, 1, 2 * k + 1);  // item k started (fetcher's credit)
      ptx::fence_proxy_async_all();                   // acquired generic-proxy view -> the TMA (async proxy) reads below
      if (P.trace != nullptr && (int)cur.t == P.trace_step && lane == 0 && leader)
        P.trace[(size_t)(P.seg[cur.seg].item_base + cur.mp * P.seg[cur.seg].n_windows + cur.win) * 4] =
            (ptx::globaltimer() & 0xFFFFFFFFFFFFull) | ((unsigned long long)(blockIdx.x >> 1) << 48);
      ++n_items;
      const LoopSeg& sg = P.seg[cur.seg];
      const uint32_t rbeg = cur.rbeg, rend = cur.rend;
      const uint32_t half_b = sg.half_b;
      const int n_half = (int)(sg.n_tile >> 1);
      const int row0 = (2 * (int)cur.mp + (int)rank) * kRowTile;
      for (uint32_t base = rbeg; base < rend; base += TC2_REC_BATCH) {
        ptx::st_shared_v4(ring + lane * 16u, mine.x, mine.y, mine.z, mine.w);
        __syncwarp();
        if (2 * (base + TC2_REC_BATCH) + lane < 2 * rend) mine = __ldg(reinterpret_cast<const uint4*>(stream + base + TC2_REC_BATCH) + lane);
        // the next item, if the fetcher already has it: its first records travel while this item's loads are issued
        if (!have_nxt && mail_ready(k + 1)) {
          nxt = mail_take(k + 1);
          have_nxt = true;
          if (nxt.seg != 0xFFFFu && 2 * nxt.rbeg + lane < 2 * nxt.rend) mine_nxt = __ldg(reinterpret_cast<const uint4*>(stream + nxt.rbeg) + lane);
        }
        const uint32_t cnt = min((uint32_t)TC2_REC_BATCH, rend - base);
        for (uint32_t i = 0; i < cnt; ++i, ++it) {
          const uint4 r0 = ptx::ld_shared_v4(ring + i * 32u);
          const uint2 r1 = ptx::ld_shared_v2(ring + i * 32u + 16u);
          const uint32_t slot = it & (TC2_NSLOT - 1);
          const int kc = (r0.x >> 8) & 0xF, nA = (r0.x >> 12) & 0x7, nB = (r0.x >> 15) & 0xF;
          uint32_t dep;
          const uint32_t off_kb = loop_ring_alloc(rs, it, ((uint32_t)nA * TC_A_BYTES + (uint32_t)nB * half_b + 1023u) >> 10, lane, &dep);
          const long long tr0 = P.dbg ? clock64() : 0;
          if (it >= dep) ptx::mbar_wait(bar_empty + 8 * ((it - dep) & (TC2_NSLOT - 1)), ((it - dep) >> 3) & 1);   // step it-dep consumed
          if (dep != TC2_NSLOT && it >= TC2_NSLOT) ptx::mbar_wait(bar_empty + 8 * slot, ((it - TC2_NSLOT) >> 3) & 1);
          if (P.dbg) t_ring += clock64() - tr0;
          const uint32_t full = bar_full + 8 * slot;
          const uint32_t sa = smem_base + (off_kb << 10);
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
        }
        __syncwarp();
      }
      if (leader && lane == 0) ptx::st_shared_cluster_u32(bar_base + LB_CREDIT, 1, 2 * k + 2);  // item k issued
      if (!have_nxt) {
        const long long tw0 = P.dbg ? clock64() : 0;
        nxt = mail_read(k + 1);
        if (P.dbg) t_pop += clock64() - tw0;
        if (nxt.seg != 0xFFFFu && 2 * nxt.rbeg + lane < 2 * nxt.rend) mine_nxt = __ldg(reinterpret_cast<const uint4*>(stream + nxt.rbeg) + lane);
      }
      cur = nxt; mine = mine_nxt; have_nxt = false;
    }
    // drain: nobody leaves while MMAs may still read this CTA's shared memory
    for (uint32_t j = it > TC2_NSLOT ? it - TC2_NSLOT : 0; j < it; ++j) ptx::mbar_wait(bar_empty + 8 * (j & (TC2_NSLOT - 1)), (j >> 3) & 1);
    if (P.dbg && lane == 0) {
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_P_POP] = (unsigned long long)t_pop;
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_P_RING] = (unsigned long long)t_ring;
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_P_TOTAL] = (unsigned long long)(clock64() - t_p0);
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_P_ITEMS] = (unsigned long long)n_items;
    }
   } else if (warp == 1) {
    if (!leader) {
      // ===================== fetcher (peer CTA's warp 1): ready queue -> mailbox of both CTAs =====================
      long long t_wait = 0, t_credit = 0;
      for (uint32_t k = 0;; ++k) {
        const uint32_t m = k & (LOOP_MAIL - 1);
        // not further than LOOP_AHEAD items beyond the one the producer is issuing (an item held here is an item no other
        // CTA pair can take), and the mailbox slot must have been read by everybody
        {
          const long long tw0 = P.dbg ? clock64() : 0;
          // credit = 2j+1: the producer has started item j; 2j+2: it has issued all of item j's loads
          for (;;) {
            const uint32_t credit = ptx::ld_shared_volatile_u32(bar_base + LB_CREDIT);
            if (2 * k <= credit) break;                                   // item k-1 fully issued (or k = 0): always allowed
            if (2 * k <= credit + 1 + 2 * (uint32_t)LOOP_AHEAD) {         // within the look-ahead window: only with a backlog
              const uint32_t claimed = *reinterpret_cast<volatile uint32_t*>(P.q_ctl), pushed = *reinterpret_cast<volatile uint32_t*>(P.q_ctl + 1) + P.q_init;
              if ((int)(pushed - claimed) > (int)P.n_pairs) break;
            }
            __nanosleep(40);
          }
          if (k >= LOOP_MAIL) ptx::mbar_wait_cluster(bar_base + LB_MAIL_EMPTY + 8 * m, ((k / LOOP_MAIL) - 1) & 1);
          if (P.dbg) t_credit += clock64() - tw0;
        }
        uint32_t lo = 0, hi = 0, rbeg = 0, rend = 0;
        const long long tw1 = P.dbg ? clock64() : 0;
        if (lane == 0) {
          const uint32_t idx = atomicAdd(P.q_ctl, 1u);
          const unsigned long long* slot_p = P.queue + (idx & (P.q_cap - 1u));
          const uint32_t lap = (idx >> P.q_shift) & LOOP_LAP_MASK;
          unsigned long long e = ptx::ld_acquire_gpu_u64(slot_p);
          if ((((uint32_t)e >> LOOP_LAP_SHIFT) & 0xFFFu) != lap) {
            const unsigned long long t0 = ptx::globaltimer();
            uint32_t spins = 0;
            while (e = ptx::ld_acquire_gpu_u64(slot_p), (((uint32_t)e >> LOOP_LAP_SHIFT) & 0xFFFu) != lap) {
              if ((++spins & 255u) == 0u) {
                // nothing became ready for seconds: a broken plan or a faulted peer - give up instead of hanging the GPU
                if (*reinterpret_cast<volatile uint32_t*>(P.status) != 0u || ptx::globaltimer() - t0 > 4000000000ull) {
                  atomicExch(P.status, 1u);
                  e = LOOP_SENTINEL;
                  break;
                }
              }
            }
          }
          e &= ~((unsigned long long)0xFFFu << LOOP_LAP_SHIFT);
          lo = (uint32_t)e; hi = (uint32_t)(e >> 32);
          if ((lo & 0xF0000u) != LOOP_SENTINEL) {
            const uint32_t wi = P.seg[lo >> 16].win_base + (lo & 0xFFFFu);
            rbeg = __ldg(P.win_rec_off + wi); rend = __ldg(P.win_rec_off + wi + 1);
          }
          const uint32_t da = bar_base + LB_MAIL_DATA + 16 * m;
          ptx::st_shared_u32(da, lo); ptx::st_shared_u32(da + 4, hi); ptx::st_shared_u32(da + 8, rbeg); ptx::st_shared_u32(da + 12, rend);
          ptx::st_shared_cluster_u32(da, 0, lo); ptx::st_shared_cluster_u32(da + 4, 0, hi);
          ptx::st_shared_cluster_u32(da + 8, 0, rbeg); ptx::st_shared_cluster_u32(da + 12, 0, rend);
          ptx::mbar_arrive_remote_release(bar_base + LB_MAIL_FULL + 8 * m, 0);
          ptx::mbar_arrive_remote_release(bar_base + LB_MAIL_FULL + 8 * m, 1);
        }
        lo = __shfl_sync(0xffffffffu, lo, 0);
        if (P.dbg) t_wait += clock64() - tw1;
        if ((lo & 0xF0000u) == LOOP_SENTINEL) break;
      }
      if (P.dbg && lane == 0) {
        P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_F_QUEUE] = (unsigned long long)t_wait;
        P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_F_CREDIT] = (unsigned long long)t_credit;
      }
    } else {
      // ===================== MMA issuer (leader CTA only) =====================
      const TcRec* __restrict__ stream = P.tmpl_m;
      const uint32_t ring = stg_base + TC2_REC_BATCH * (uint32_t)sizeof(TcRec);
      const uint64_t desc0 = make_smem_desc_sw128(smem_base);
      const uint32_t desc_lo0 = (uint32_t)desc0, desc_hi = (uint32_t)(desc0 >> 32);
      uint32_t it = 0, item_count = 0;
      LoopRing rs;
      long long t_full = 0, t_acc = 0, t_mail = 0;
      const long long t_m0 = P.dbg ? clock64() : 0;
      LoopItem cur, nxt;
      uint4 mine = make_uint4(0, 0, 0, 0), mine_nxt = make_uint4(0, 0, 0, 0);
      bool have_nxt = false;
      {
        const long long tm0 = P.dbg ? clock64() : 0;
        cur = mail_read(0);
        if (P.dbg) t_mail += clock64() - tm0;
        if (cur.seg != 0xFFFFu && 2 * cur.rbeg + lane < 2 * cur.rend) mine = __ldg(reinterpret_cast<const uint4*>(stream + cur.rbeg) + lane);
      }
      for (uint32_t k = 0; cur.seg != 0xFFFFu; ++k) {
        const LoopSeg& sg = P.seg[cur.seg];
        const uint32_t idesc = sg.idesc, acc_stride = sg.acc_stride, half_b = sg.half_b, half_b16 = half_b >> 4, n_merge = (sg.n_tile >> 3) << 17;
        const uint32_t rbeg = cur.rbeg, rend = cur.rend;
        const uint32_t buf = item_count & 1;
        {                                                       // the item's accumulator buffer must be drained
          const long long ta0 = P.dbg ? clock64() : 0;
          ptx::mbar_wait(bar_acc_empty + 8 * buf, ((item_count >> 1) & 1) ^ 1);
          if (P.dbg) t_acc += clock64() - ta0;
          if (P.trace != nullptr && (int)cur.t == P.trace_step && lane == 0)
            P.trace[(size_t)(sg.item_base + cur.mp * sg.n_windows + cur.win) * 4 + 1] = ptx::globaltimer();
        }
        for (uint32_t base = rbeg; base < rend; base += TC2_REC_BATCH) {
          ptx::st_shared_v4(ring + lane * 16u, mine.x, mine.y, mine.z, mine.w);
          __syncwarp();
          if (2 * (base + TC2_REC_BATCH) + lane < 2 * rend) mine = __ldg(reinterpret_cast<const uint4*>(stream + base + TC2_REC_BATCH) + lane);
          if (!have_nxt && mail_ready(k + 1)) {
            nxt = mail_take(k + 1);
            have_nxt = true;
            if (nxt.seg != 0xFFFFu && 2 * nxt.rbeg + lane < 2 * nxt.rend) mine_nxt = __ldg(reinterpret_cast<const uint4*>(stream + nxt.rbeg) + lane);
          }
          const uint32_t cnt = min((uint32_t)TC2_REC_BATCH, rend - base);
          for (uint32_t i = 0; i < cnt; ++i, ++it) {
            const uint4 r0 = ptx::ld_shared_v4(ring + i * 32u);
            const uint4 r1 = ptx::ld_shared_v4(ring + i * 32u + 16u);
            const uint32_t slot = it & (TC2_NSLOT - 1), phase = (it >> 3) & 1;
            const int nA = (r0.x >> 8) & 0x7, n_ops = (r0.x >> 11) & 0x1F, nB = (r0.x >> 18) & 0xF;
            const bool last = (base + i + 1 == rend);
            uint32_t dep;
            const uint32_t off_kb = loop_ring_alloc(rs, it, ((uint32_t)nA * TC_A_BYTES + (uint32_t)nB * half_b + 1023u) >> 10, lane, &dep);
            const long long tf0 = P.dbg ? clock64() : 0;
            ptx::mbar_wait(bar_full + 8 * slot, phase);
            if (P.dbg) t_full += clock64() - tf0;
            ptx::tc_fence_after();
            // descriptors differ only in the 14-bit start-address field: one 32-bit add each (smem < 256 KB, no carry)
            const uint32_t a_lo0 = desc_lo0 + (off_kb << 6);
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
                for (int kk = 0; kk < 4; ++kk)
                  ptx::umma_f16_2sm(d, ((uint64_t)desc_hi << 32) | (a_lo + 2u * kk), ((uint64_t)desc_hi << 32) | (b_lo + 2u * kk), idg,
                                    (kk > 0 || !first_mma) ? 1u : 0u);
              }
              ptx::umma_commit_2sm(bar_empty + 8 * slot);           // this step is consumed (both CTAs)
              if (last) ptx::umma_commit_2sm(bar_acc_full + 8 * buf);   // last step: accumulators complete in both CTAs
            }
            __syncwarp();
          }
          __syncwarp();
        }
        ++item_count;
        if (!have_nxt) {
          const long long tm0 = P.dbg ? clock64() : 0;
          nxt = mail_read(k + 1);
          if (P.dbg) t_mail += clock64() - tm0;
          if (nxt.seg != 0xFFFFu && 2 * nxt.rbeg + lane < 2 * nxt.rend) mine_nxt = __ldg(reinterpret_cast<const uint4*>(stream + nxt.rbeg) + lane);
        }
        cur = nxt; mine = mine_nxt; have_nxt = false;
      }
      // drain: observe the release of the last (up to two) accumulator buffers by the epilogue warps of both CTAs
      for (uint32_t j = item_count > 2 ? item_count - 2 : 0; j < item_count; ++j) ptx::mbar_wait(bar_acc_empty + 8 * (j & 1), (j >> 1) & 1);
      if (P.dbg && lane == 0) {
        P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_M_FULL] = (unsigned long long)t_full;
        P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_M_ACC] = (unsigned long long)t_acc;
        P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_M_MAIL] = (unsigned long long)t_mail;
        P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_M_TOTAL] = (unsigned long long)(clock64() - t_m0);
      }
    }
   } else {
    // ===================== store warps (warp 2 + h serves epilogue half h) =====================
    // Takes the epilogue half's staged 128x64 fp16 tiles, stores them by TMA (lane 0), frees the staging tile as soon as
    // the store has READ it, and reports the item's completion once its stores have COMPLETED (only the issuing thread
    // can wait for that); the warp that reports the last of an item's four parts wakes its successors.
    const int h = warp - 2;
    const uint32_t tile_full = bar_base + LB_TILE_FULL + 8 * (uint32_t)h, tile_free = bar_base + LB_TILE_FREE + 8 * (uint32_t)h;
    const uint32_t s_out = epi_base + (uint32_t)h * TC2_TILE_BYTES;
    uint32_t tcount = 0;
    long long t_tile = 0, t_done = 0;
    const long long t_s0 = P.dbg ? clock64() : 0;
    for (uint32_t k = 0;; ++k) {
      const LoopItem cur = mail_read(k);
      if (cur.seg == 0xFFFFu) break;
      const LoopSeg& sg = P.seg[cur.seg];
      if (!(sg.kind <= LK_NONE64H)) continue;                    // fp32 / last-layer epilogues store (and report) themselves
      const TcItem2* ip = sg.items + cur.win;
      const int G = (int)(sg.n_tile >> 6), n_units = (int)__ldg(&ip->n_acc) * G;
      const int row0 = (2 * (int)cur.mp + (int)rank) * kRowTile;
      const LoopSucc sq = loop_succ_prefetch(P, sg, cur.win, lane);     // in flight while the tiles are stored
      if (lane == 0) {
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
        ptx::bulk_wait_all0();                                    // the item's tiles are in global memory
        loop_publish_fence();
        if (P.dbg) t_done += clock64() - tw1;
      }
      __syncwarp();
      loop_complete(P, cur.seg, cur.mp, cur.t, lane, sq);         // one of the item's 4 parts (2 halves x 2 CTAs)
    }
    if (P.dbg && lane == 0) {
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_S_TILE] = (unsigned long long)t_tile;    // (warp 3 overwrites warp 2: same order of magnitude)
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_S_DONE] = (unsigned long long)t_done;
      P.dbg[(size_t)blockIdx.x * DBG_COUNT + DBG_S_TOTAL] = (unsigned long long)(clock64() - t_s0);
    }
   }
  } else {
    ptx::setmaxnreg_inc<LOOP_REGS_EPI>();
    // ===================== epilogue (warps 4..11, both CTAs) =====================
    LoopCtx cx;
    cx.tmem_base = tmem_base; cx.bar_base = bar_base; cx.epi_base = epi_base;
    cx.warp = warp; cx.lane = lane; cx.rank = (int)rank; cx.item_count = 0; cx.tile_count = 0; cx.t_acc = 0; cx.t_tile = 0;
    const long long t_e0 = P.dbg ? clock64() : 0;
    TcFinalArgs fa{};
    fa.x = P.x; fa.y = P.y; fa.loss_part = P.loss_part; fa.R = P.R; fa.B = P.B; fa.n_rows = P.n_rows; fa.nbx = P.nbx; fa.w_out = P.w_out;
    fa.gscale = P.gscale; fa.m_gmul = P.m_gmul; fa.m_mu = P.m_mu;
    for (uint32_t k = 0;; ++k, ++cx.item_count) {
      const LoopItem cur = mail_read(k);
      if (cur.seg == 0xFFFFu) break;
      const LoopSeg& sg = P.seg[cur.seg];
      const int t = (int)cur.t;
      fa.write_y = (t == P.last_step) ? 1 : 0;      // G(z) and the loss are consumed after the final forward only
      fa.m_lr = (P.decay_step > 0 && t >= P.decay_step) ? P.m_lr * 0.1f : P.m_lr;
      unsigned long long ts = 0;
      if (P.prof != nullptr && warp == LOOP_EPI_WARP0 && lane == 0 && leader) ts = ptx::globaltimer();
      loop_epilogue_dispatch<ARCH>(P, sg, cx, fa, cur);
      if (P.prof != nullptr && warp == LOOP_EPI_WARP0 && lane == 0 && leader) {
        const unsigned long long te = ptx::globaltimer();
        unsigned long long* pr = P.prof + ((size_t)t * P.n_seg + cur.seg) * 2;
        atomicMin(pr, ts);
        atomicMax(pr + 1, te);
        if (P.trace != nullptr && t == P.trace_step) {
          unsigned long long* tr = P.trace + (size_t)(sg.item_base + cur.mp * sg.n_windows + cur.win) * 4;
          tr[2] = ts; tr[3] = te;
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
struct LoopSegSpec {              // one layer-direction as the planner sees it
  std::string name;
  int N = 0, K = 0;               // MMA N (output channels per pixel) and K (input channels per pixel)
  int kind = 0;                   // LoopKind
  const PairTable* tab = nullptr; // (input pixel, weight tile) contributions of every output pixel
  int h_grid = 1, w_grid = 1;     // raster of the output pixels (window shapes)
  int max_acc = 1;                // accumulators per window (TMEM columns / layer-specific cap)
  int in_seg = -1;                // segment whose output this one reads; -1: z, released by the momentum tail
  bool fwd = true;                // part of the generator forward (+ loss) or of the backward-to-z
  double macs_per_row = 0.0;      // exact in-bounds MACs per latent row (profiling only)
};

// The plan of one L-step: per segment its window tiling and per window its step records (templates: the row pair and the
// ring placement are filled in at run time), plus the dataflow graph between windows of consecutive segments.
struct LoopPlan {
  int n_seg = 0, n_fwd = 0, n_pairs = 0, n_mpairs = 0;
  std::vector<std::vector<TcItem2>> hdrs;         // per segment: window headers
  std::vector<int> shape;                         // per segment: wh, ww, sy, sx
  std::vector<double> cost_total, cost_max;       // per segment: planner's cost of one row pair's items / of its largest item (bytes)
  std::vector<uint32_t> win_base, item_base, n_windows;     // per segment
  uint32_t n_win = 0, n_item_slots = 0;           // windows of all segments; counters = windows x row pairs
  std::vector<TcRec> tmpl_p[2], tmpl_m;           // windows back to back in (segment, window) order
  std::vector<uint32_t> win_rec_off;              // [n_win + 1]
  std::vector<uint32_t> succ_off, succ, need;     // [n_win + 1], successor entries (segment << 16 | window), [n_win]
  std::vector<unsigned long long> q_init;         // the queue's initial content: the first segment's items of L-step 0
  uint32_t win_fwd = 0, win_bwd = 0;              // windows of the forward / backward half
  uint32_t q_cap = 0;
  long long n_steps = 0, n_mma = 0, n_bytes = 0;  // of one L-step of every row pair
};

#ifndef DGAN_COST_EPI_KB
#define DGAN_COST_EPI_KB 24.0
#endif
#ifndef DGAN_COST_FIXED_KB
#define DGAN_COST_FIXED_KB 48.0
#endif
constexpr int LOOP_STEP_MAX_BYTES = 48 * 1024;    // measured optimum of the operand-ring kernels (round 1): 2 A tiles + weights
constexpr uint32_t LOOP_TAIL_NEED = 2;            // the first segment waits for the momentum tails of the row pair's two 128-row tiles
constexpr uint32_t LOOP_PARTS_TMA = 4, LOOP_PARTS_WARPS = 2 * TC2_EPI_WARPS;     // completion parts of an item (see loop_complete)
static inline uint32_t loop_parts_of(int kind) { return kind <= LK_NONE64H ? LOOP_PARTS_TMA : LOOP_PARTS_WARPS; }

static double loop_item_cost(const Tc2HostItem& it, int N) {
  return it.stage_bytes + DGAN_COST_EPI_KB * 1024.0 * it.hdr.n_acc * std::max(1, N / 64) + DGAN_COST_FIXED_KB * 1024.0;
}

// Window tiling of one segment.  Every candidate shape (wh x ww accumulators, strides 1 or 2 - stride 2 gathers outputs
// of equal parity of a stride-2 transposed conv, which share weight tiles) is scored by a proxy of what it adds to the
// L-step:  (sum of item costs) x row pairs / CTA pairs   - its share of the machine's time when everything is busy -
//   plus  alpha x (largest item cost)                    - what it adds to a row pair's critical path (a row pair's
// segments run one after the other, and with few row pairs in flight the chain, not the machine, sets the pace).
// Item cost = operand bytes staged + a per-accumulator epilogue charge + a fixed per-item charge.
#ifndef DGAN_TILE_ALPHA
#define DGAN_TILE_ALPHA 1.0
#endif
static double loop_tile_alpha() {
  const char* e = std::getenv("DGAN_TILE_ALPHA");        // developer knob (planner experiments)
  return e ? std::atof(e) : (double)DGAN_TILE_ALPHA;
}
static int loop_choose_tiling(const LoopSegSpec& sp, int n_mps, int n_pairs, std::vector<Tc2HostItem>* items_out, int shape_out[4],
                              double* total_out = nullptr, double* max_out = nullptr) {
  const int N = sp.N, K = sp.K;
  const int max_g = (N >= 64) ? std::min(4, 256 / N) : 1;
  const int step_max = std::min((LOOP_RING_BYTES / 2) & ~1023, LOOP_STEP_MAX_BYTES);
  const double alpha = loop_tile_alpha();
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
          double total = 0.0, largest = 0.0;
          for (size_t i = 0; i < wins.size(); ++i) {
            tc2_build_item(*sp.tab, wins[i], N, K, max_g, TC2_MAX_A, step_max, &items[i]);
            const double c = loop_item_cost(items[i], N);
            total += c; largest = std::max(largest, c);
          }
          const double score = total * (double)n_mps / (double)n_pairs + alpha * largest;
          if (score < best_cost) {
            best_cost = score;
            shape_out[0] = wh; shape_out[1] = ww; shape_out[2] = sy; shape_out[3] = sx;
            best_items.swap(items);
            if (total_out) *total_out = total;
            if (max_out) *max_out = largest;
          }
        }
  if (best_items.empty()) { set_error("no window tiling for segment " + sp.name); return DGAN_ERR_UNSUPPORTED; }
  items_out->swap(best_items);
  return 0;
}

// Plan for `n_mpairs` row pairs on `n_pairs` CTA pairs.
static int loop_plan(const std::vector<LoopSegSpec>& specs, int n_mpairs, int n_pairs, LoopPlan* plan) {
  const int n_seg = (int)specs.size();
  if (n_seg < 1 || n_seg > LOOP_MAX_SEG) { set_error("segment count out of range"); return DGAN_ERR_UNSUPPORTED; }
  if (n_mpairs < 1 || n_mpairs > 0xFFFF || n_pairs < 1) { set_error("nothing to plan"); return DGAN_ERR_INVALID_ARG; }
  LoopPlan& pl = *plan;
  pl = LoopPlan{};
  pl.n_seg = n_seg; pl.n_pairs = n_pairs; pl.n_mpairs = n_mpairs;
  for (int s = 0; s < n_seg; ++s) {
    if (specs[(size_t)s].fwd) { if (pl.n_fwd != s) { set_error("forward segments must come first"); return DGAN_ERR_UNSUPPORTED; } pl.n_fwd = s + 1; }
    if (specs[(size_t)s].in_seg != s - 1) { set_error("segments must form a chain"); return DGAN_ERR_UNSUPPORTED; }
  }
  pl.hdrs.assign((size_t)n_seg, {}); pl.shape.assign((size_t)4 * n_seg, 1);
  pl.cost_total.assign((size_t)n_seg, 0.0); pl.cost_max.assign((size_t)n_seg, 0.0);
  pl.win_base.assign((size_t)n_seg, 0); pl.item_base.assign((size_t)n_seg, 0); pl.n_windows.assign((size_t)n_seg, 0);
  std::vector<std::vector<Tc2HostItem>> items((size_t)n_seg);
  std::vector<std::vector<int>> pix2win((size_t)n_seg);
  const int ring_kb = LOOP_RING_BYTES / 1024;
  int rc;
  for (int s = 0; s < n_seg; ++s) {
    const LoopSegSpec& sp = specs[(size_t)s];
    if ((rc = loop_choose_tiling(sp, n_mpairs, n_pairs, &items[(size_t)s], &pl.shape[(size_t)4 * s], &pl.cost_total[(size_t)s], &pl.cost_max[(size_t)s]))) return rc;
    const size_t nw = items[(size_t)s].size();
    pl.hdrs[(size_t)s].resize(nw);
    pix2win[(size_t)s].assign(sp.tab->off.size() - 1, -1);
    for (size_t w = 0; w < nw; ++w) {
      pl.hdrs[(size_t)s][w] = items[(size_t)s][w].hdr;
      for (uint32_t a = 0; a < items[(size_t)s][w].hdr.n_acc; ++a) pix2win[(size_t)s][items[(size_t)s][w].hdr.q[a]] = (int)w;
    }
    pl.win_base[(size_t)s] = pl.n_win;
    pl.item_base[(size_t)s] = pl.n_item_slots;
    pl.n_windows[(size_t)s] = (uint32_t)nw;
    pl.n_win += (uint32_t)nw;
    pl.n_item_slots += (uint32_t)(nw * (size_t)n_mpairs);
    (sp.fwd ? pl.win_fwd : pl.win_bwd) += (uint32_t)nw;
  }
  // ---- step records (templates) and the dependency sets: the windows of the previous segment that write the pixels an
  //      item stages
  std::vector<std::vector<uint32_t>> deps(pl.n_win);
  pl.win_rec_off.assign(1, 0);
  for (int s = 0; s < n_seg; ++s) {
    const LoopSegSpec& sp = specs[(size_t)s];
    for (size_t w = 0; w < items[(size_t)s].size(); ++w) {
      const Tc2HostItem& itm = items[(size_t)s][w];
      std::vector<uint32_t>& dl = deps[pl.win_base[(size_t)s] + w];
      if (itm.steps.empty()) { set_error(sp.name + ": a window without steps"); return DGAN_ERR_UNSUPPORTED; }
      for (size_t j = 0; j < itm.steps.size(); ++j) {
        const Tc2HostStep& hs = itm.steps[j];
        const int kb = (hs.bytes + 1023) / 1024;
        if (kb > ring_kb / 2) { set_error("tensor-core step larger than half the operand ring"); return DGAN_ERR_UNSUPPORTED; }
        if (sp.in_seg >= 0)
          for (int a = 0; a < hs.nA; ++a) {
            const std::vector<int>& p2w = pix2win[(size_t)sp.in_seg];
            const int p = hs.a_pix[a];
            if (p < 0 || (size_t)p >= p2w.size() || p2w[(size_t)p] < 0) { set_error(sp.name + ": input pixel without a producer"); return DGAN_ERR_UNSUPPORTED; }
            if (std::find(dl.begin(), dl.end(), (uint32_t)p2w[(size_t)p]) == dl.end()) dl.push_back((uint32_t)p2w[(size_t)p]);
          }
        const uint32_t flags = (j == 0 ? 1u : 0u) | (j + 1 == itm.steps.size() ? 2u : 0u);
        TcRec rm{};
        rm.w[0] = ((uint32_t)hs.nA << 8) | ((uint32_t)hs.n_ops << 11) | (flags << 16) | ((uint32_t)hs.nB << 18);
        for (int o = 0; o < hs.n_ops; ++o) rm.w[2 + o / 2] |= (uint32_t)hs.ops[o] << (16 * (o & 1));
        pl.tmpl_m.push_back(rm);
        for (int r = 0; r < 2; ++r) {
          TcRec rp{};
          rp.w[0] = ((uint32_t)hs.kc << 8) | ((uint32_t)hs.nA << 12) | ((uint32_t)hs.nB << 15);
          for (int a = 0; a < hs.nA; ++a) rp.w[2 + a / 2] |= (uint32_t)(hs.a_pix[a] & 0xFFFF) << (16 * (a & 1));
          for (int b = 0; b < hs.nB; ++b) rp.w[4 + b / 4] |= (uint32_t)hs.b_ent[r][b] << (8 * (b & 3));
          pl.tmpl_p[r].push_back(rp);
        }
        pl.n_mma += (long long)hs.n_ops * n_mpairs; pl.n_steps += n_mpairs; pl.n_bytes += (long long)hs.bytes * n_mpairs;
      }
      pl.win_rec_off.push_back((uint32_t)pl.tmpl_m.size());
    }
  }
  // ---- the graph: need = size of the dependency set; successors = its inverse
  pl.need.assign(pl.n_win, 0);
  std::vector<std::vector<uint32_t>> succ(pl.n_win);
  for (int s = 0; s < n_seg; ++s)
    for (uint32_t w = 0; w < pl.n_windows[(size_t)s]; ++w) {
      const uint32_t wi = pl.win_base[(size_t)s] + w;
      if (specs[(size_t)s].in_seg < 0) { pl.need[wi] = LOOP_TAIL_NEED; continue; }
      pl.need[wi] = (uint32_t)deps[wi].size() * loop_parts_of(specs[(size_t)specs[(size_t)s].in_seg].kind);
      for (uint32_t u : deps[wi]) succ[pl.win_base[(size_t)specs[(size_t)s].in_seg] + u].push_back(((uint32_t)s << 16) | w);
    }
  pl.succ_off.assign(1, 0);
  for (uint32_t wi = 0; wi < pl.n_win; ++wi) {
    pl.succ.insert(pl.succ.end(), succ[wi].begin(), succ[wi].end());
    pl.succ_off.push_back((uint32_t)pl.succ.size());
  }
  if (pl.succ.empty()) pl.succ.push_back(0);
  // ---- the queue.  At any time a row pair has ready items of ONE L-step only (its next L-step is released by its last
  //      Linear-backward tail), so ready-but-untaken entries never exceed windows-per-step x row pairs (+ the sentinels).
  //      Entries carry the lap of their index, so a slot needs no reset; 4x head-room keeps a slot from being rewritten
  //      before its previous entry has been read.
  {
    const unsigned long long bound = 4ull * ((unsigned long long)(pl.win_fwd + pl.win_bwd) * (unsigned long long)n_mpairs + (unsigned long long)n_pairs) + 64ull;
    uint32_t cap = 1024;
    while (cap < bound && cap < (1u << 30)) cap <<= 1;
    if (cap < bound) { set_error("ready queue too large"); return DGAN_ERR_UNSUPPORTED; }
    pl.q_cap = cap;
  }
  for (uint32_t w = 0; w < pl.n_windows[0]; ++w)            // window-major: the first pops spread over the row pairs
    for (int mp = 0; mp < n_mpairs; ++mp) pl.q_init.push_back(((unsigned long long)(uint32_t)mp << 32) | w);
  return 0;
}

// Items of a launch over `rec_iters` L-steps (the final L-step is forward only unless `full_last`), and their parts.
static inline unsigned long long loop_total_items(const LoopPlan& pl, int rec_iters, bool full_last) {
  return (unsigned long long)pl.n_mpairs * ((unsigned long long)rec_iters * pl.win_fwd + (unsigned long long)(rec_iters - (full_last ? 0 : 1)) * pl.win_bwd);
}
static inline unsigned long long loop_total_parts(const std::vector<LoopSegSpec>& specs, const LoopPlan& pl, int rec_iters, bool full_last) {
  unsigned long long n = 0;
  for (int s = 0; s < pl.n_seg; ++s)
    n += (unsigned long long)pl.n_windows[(size_t)s] * loop_parts_of(specs[(size_t)s].kind) * (unsigned long long)(specs[(size_t)s].fwd ? rec_iters : rec_iters - (full_last ? 0 : 1));
  return n * (unsigned long long)pl.n_mpairs;
}

// Host twin of loop_ring_alloc (the ring placement the kernel computes at run time).
struct LoopRingHost {
  int cursor = 0;
  std::vector<std::pair<int, int>> hist;          // regions of all steps so far
  int alloc(int kb, int* dep) {
    const int ring_kb = LOOP_RING_BYTES / 1024;
    if (cursor + kb > ring_kb) cursor = 0;
    const int beg = cursor, end = beg + kb;
    cursor = end;
    int d = TC2_NSLOT;
    for (int k = TC2_NSLOT - 1; k >= 1; --k) {
      const int c = (int)hist.size() - k;
      if (c >= 0 && hist[(size_t)c].first < end && beg < hist[(size_t)c].second) d = k;
    }
    *dep = d;
    hist.push_back({beg, end});
    return beg;
  }
};

// Independent validation of a plan (host only; dgan_debug_check_plans and the CPU tests):
//  * per segment, everything tc2_check_plan re-derives from the records of its windows (every (output pixel, input pixel,
//    tap, k-chunk) contribution exactly once into the right accumulator, first-MMA flags, canonical accumulation order,
//    every window present exactly once);
//  * the run-time ring placement, replayed over a pseudo-random item sequence: regions stay inside the ring and a region
//    is only reused after the wait the kernel performs covers every step that overlaps it;
//  * the graph: every pixel a window stages is written by a window in its dependency set, `need` equals that set's size,
//    the successor lists are exactly the inverse of the dependency sets;
//  * liveness: executing the dataflow rules of the kernel (loop_complete, the momentum tail) for 3 L-steps in a
//    pseudo-random order runs every item exactly once per L-step, in L-step order, and pushes exactly the expected number
//    of queue entries; the queue never holds more than a quarter of its capacity.
static int loop_check_plan(const std::vector<LoopSegSpec>& specs, const LoopPlan& pl, std::string* err) {
  auto fail = [&](const std::string& m) { *err = m; return DGAN_ERR_INVALID_ARG; };
  const int n_seg = pl.n_seg;
  if ((int)specs.size() != n_seg || n_seg < 1 || n_seg > LOOP_MAX_SEG) return fail("segment count");
  if (pl.tmpl_p[0].size() != pl.tmpl_m.size() || pl.tmpl_p[1].size() != pl.tmpl_m.size()) return fail("record arrays differ in size");
  if (pl.win_rec_off.size() != (size_t)pl.n_win + 1 || pl.succ_off.size() != (size_t)pl.n_win + 1 || pl.need.size() != pl.n_win) return fail("per-window table size");
  if (pl.win_rec_off.back() != pl.tmpl_m.size()) return fail("record offsets do not cover the records");
  uint32_t wsum = 0, isum = 0, wf = 0, wb = 0;
  for (int s = 0; s < n_seg; ++s) {
    if (pl.win_base[(size_t)s] != wsum || pl.item_base[(size_t)s] != isum || pl.n_windows[(size_t)s] != pl.hdrs[(size_t)s].size()) return fail("segment bases");
    if (pl.n_windows[(size_t)s] == 0 || pl.n_windows[(size_t)s] > 0xFFFFu) return fail("window count of a segment");
    wsum += pl.n_windows[(size_t)s]; isum += pl.n_windows[(size_t)s] * (uint32_t)pl.n_mpairs;
    (specs[(size_t)s].fwd ? wf : wb) += pl.n_windows[(size_t)s];
    if (specs[(size_t)s].fwd != (s < pl.n_fwd)) return fail("forward segments must come first");
    if (specs[(size_t)s].in_seg != s - 1) return fail("segments must form a chain");
  }
  if (wsum != pl.n_win || isum != pl.n_item_slots || wf != pl.win_fwd || wb != pl.win_bwd) return fail("window totals");
  std::vector<std::vector<int>> pix2win((size_t)n_seg);
  for (int s = 0; s < n_seg; ++s) {
    pix2win[(size_t)s].assign(specs[(size_t)s].tab->off.size() - 1, -1);
    for (size_t w = 0; w < pl.hdrs[(size_t)s].size(); ++w)
      for (uint32_t a = 0; a < pl.hdrs[(size_t)s][w].n_acc; ++a) {
        const uint32_t q = pl.hdrs[(size_t)s][w].q[a];
        if ((size_t)q >= pix2win[(size_t)s].size()) return fail(specs[(size_t)s].name + ": window pixel out of range");
        if (pix2win[(size_t)s][q] != -1) return fail(specs[(size_t)s].name + ": an output pixel belongs to two windows");
        pix2win[(size_t)s][q] = (int)w;
      }
    for (int v : pix2win[(size_t)s])
      if (v < 0) return fail(specs[(size_t)s].name + ": an output pixel belongs to no window");
  }
  // ---- records: one pseudo CTA pair that runs every window of the segment once for row pair 0
  for (int s = 0; s < n_seg; ++s) {
    const LoopSegSpec& sp = specs[(size_t)s];
    Tc2Plan sub;
    sub.n_pairs = 1;
    sub.hdrs = pl.hdrs[(size_t)s];
    sub.stream_off = {0u, 0u};
    sub.n_slots = (int)pl.n_windows[(size_t)s];
    for (uint32_t w = 0; w < pl.n_windows[(size_t)s]; ++w) {
      const uint32_t r0 = pl.win_rec_off[pl.win_base[(size_t)s] + w], r1 = pl.win_rec_off[pl.win_base[(size_t)s] + w + 1];
      if (r0 >= r1 || r1 > pl.tmpl_m.size()) return fail(sp.name + ": record offsets not increasing");
      for (uint32_t ri = r0; ri < r1; ++ri) {
        TcRec m = pl.tmpl_m[ri], p0 = pl.tmpl_p[0][ri], p1 = pl.tmpl_p[1][ri];
        if ((p0.w[0] & 0xFFu) != 0 || (m.w[0] & 0xFFu) != 0 || ((p0.w[0] >> 19) & 0xFu) != 0 || p0.w[1] != 0 || p1.w[1] != 0 || m.w[1] != 0)
          return fail(sp.name + ": template record carries run-time fields");
        if (((m.w[0] >> 18) & 0xFu) != ((p0.w[0] >> 15) & 0xFu)) return fail(sp.name + ": MMA record's weight-slot count differs from the producer's");
        if ((((m.w[0] >> 16) & 1u) != 0) != (ri == r0) || (((m.w[0] >> 17) & 1u) != 0) != (ri + 1 == r1)) return fail(sp.name + ": first/last-step marks do not match the window's record range");
        m.w[0] &= ~(0xFu << 18);                            // tc2_check_plan's format: no slot count in the MMA record,
        p0.w[0] |= (uint32_t)TC2_NSLOT << 19; p1.w[0] |= (uint32_t)TC2_NSLOT << 19;     // a dep distance in the producer's
        sub.stream_m.push_back(m); sub.stream_p[0].push_back(p0); sub.stream_p[1].push_back(p1);
      }
      sub.eitems.push_back((int)(w << 16));
    }
    sub.stream_off[1] = (uint32_t)sub.stream_m.size();
    std::string e2;
    if (tc2_check_plan(sp.N, sp.K, *sp.tab, 1, LOOP_RING_BYTES, sub, &e2, /*check_ring=*/false)) return fail(sp.name + ": " + e2);
  }
  // ---- graph
  std::vector<std::vector<uint32_t>> deps(pl.n_win);
  for (int s = 0; s < n_seg; ++s)
    for (uint32_t w = 0; w < pl.n_windows[(size_t)s]; ++w) {
      const uint32_t wi = pl.win_base[(size_t)s] + w;
      if (specs[(size_t)s].in_seg < 0) {
        if (pl.need[wi] != LOOP_TAIL_NEED) return fail("first segment must wait for the row pair's two momentum tails");
        continue;
      }
      std::vector<uint32_t>& dl = deps[wi];
      for (uint32_t ri = pl.win_rec_off[wi]; ri < pl.win_rec_off[wi + 1]; ++ri) {
        const TcRec& p0 = pl.tmpl_p[0][ri];
        const int nA = (int)((p0.w[0] >> 12) & 7);
        for (int a = 0; a < nA; ++a) {
          const int p = (int)((p0.w[2 + a / 2] >> (16 * (a & 1))) & 0xFFFF);
          const std::vector<int>& p2w = pix2win[(size_t)specs[(size_t)s].in_seg];
          if ((size_t)p >= p2w.size()) return fail("staged pixel out of range");
          if (std::find(dl.begin(), dl.end(), (uint32_t)p2w[(size_t)p]) == dl.end()) dl.push_back((uint32_t)p2w[(size_t)p]);
        }
      }
      if (pl.need[wi] != dl.size() * loop_parts_of(specs[(size_t)specs[(size_t)s].in_seg].kind)) return fail(specs[(size_t)s].name + ": `need` differs from the number of windows that write the staged pixels");
    }
  {
    std::vector<std::vector<uint32_t>> inv(pl.n_win);
    for (int s = 0; s < n_seg; ++s)
      for (uint32_t w = 0; w < pl.n_windows[(size_t)s]; ++w)
        for (uint32_t u : deps[pl.win_base[(size_t)s] + w]) inv[pl.win_base[(size_t)specs[(size_t)s].in_seg] + u].push_back(((uint32_t)s << 16) | w);
    for (uint32_t wi = 0; wi < pl.n_win; ++wi) {
      if (pl.succ_off[wi] > pl.succ_off[wi + 1] || pl.succ_off[wi + 1] > pl.succ.size()) return fail("successor offsets");
      std::vector<uint32_t> got(pl.succ.begin() + pl.succ_off[wi], pl.succ.begin() + pl.succ_off[wi + 1]);
      std::sort(got.begin(), got.end()); std::sort(inv[wi].begin(), inv[wi].end());
      if (got != inv[wi]) return fail("a successor list is not the inverse of the dependency sets");
    }
  }
  // ---- queue capacity and initial content
  if (pl.q_cap == 0 || (pl.q_cap & (pl.q_cap - 1)) != 0) return fail("queue capacity must be a power of two");
  if (pl.q_init.size() != (size_t)pl.n_windows[0] * (size_t)pl.n_mpairs) return fail("initial queue content");
  {
    std::vector<int> seen((size_t)pl.n_windows[0] * (size_t)pl.n_mpairs, 0);
    for (unsigned long long e : pl.q_init) {
      const uint32_t lo = (uint32_t)e, hi = (uint32_t)(e >> 32);
      if ((lo >> 16) != 0 || (lo & 0xFFFFu) >= pl.n_windows[0] || (hi >> 16) != 0 || (hi & 0xFFFFu) >= (uint32_t)pl.n_mpairs) return fail("initial queue entry is not a first-segment item of L-step 0");
      if (seen[(size_t)(hi & 0xFFFFu) * pl.n_windows[0] + (lo & 0xFFFFu)]++) return fail("initial queue entry twice");
    }
  }
  // ---- liveness and queue bound: replay the kernel's dataflow rules for 3 L-steps, taking ready items in a pseudo-random
  //      order; the ring placement is replayed along the way for one pseudo CTA pair that runs everything
  {
    const int L = 3;
    std::vector<uint32_t> depcnt(pl.n_item_slots, 0), runs(pl.n_item_slots, 0), tails((size_t)pl.n_mpairs, 0);
    std::vector<unsigned long long> ready(pl.q_init.begin(), pl.q_init.end());
    unsigned long long pushed = ready.size(), executed = 0;
    size_t max_ready = ready.size();
    uint32_t rng = 12345u;
    LoopRingHost ring;
    const int ring_kb = LOOP_RING_BYTES / 1024;
    while (!ready.empty()) {
      rng = rng * 1664525u + 1013904223u;
      const size_t pick = (size_t)(rng >> 8) % ready.size();
      const unsigned long long e = ready[pick];
      ready[pick] = ready.back(); ready.pop_back();
      const uint32_t seg = (uint32_t)e >> 16, win = (uint32_t)e & 0xFFFFu, mp = (uint32_t)(e >> 32) & 0xFFFFu, t = (uint32_t)(e >> 48);
      if ((int)seg >= n_seg || win >= pl.n_windows[seg] || (int)mp >= pl.n_mpairs || (int)t >= L) return fail("replay: bad queue entry");
      const uint32_t slot = pl.item_base[seg] + mp * pl.n_windows[seg] + win;
      if (runs[slot] != t) return fail("replay: an item runs twice or out of L-step order");
      runs[slot] = t + 1; ++executed;
      for (uint32_t ri = pl.win_rec_off[pl.win_base[seg] + win]; ri < pl.win_rec_off[pl.win_base[seg] + win + 1]; ++ri) {
        const TcRec& p0 = pl.tmpl_p[0][ri];
        const int kb = ((int)((p0.w[0] >> 12) & 7) * TC_A_BYTES + (int)((p0.w[0] >> 15) & 0xF) * (specs[seg].N / 2) * 128 + 1023) / 1024;
        int dep;
        const int beg = ring.alloc(kb, &dep);
        if (beg < 0 || beg + kb > ring_kb) return fail("replay: step region outside the ring");
        const int k = (int)ring.hist.size() - 1;
        for (int c = k - 1; c >= 0 && c > k - dep && c > k - TC2_NSLOT; --c)
          if (ring.hist[(size_t)c].first < beg + kb && beg < ring.hist[(size_t)c].second) return fail("replay: ring hazard");
      }
      const bool stop = ((int)seg == pl.n_fwd - 1) && ((int)t == L - 1);
      if (!stop)
        for (uint32_t part = 0; part < loop_parts_of(specs[seg].kind); ++part)
          for (uint32_t si = pl.succ_off[pl.win_base[seg] + win]; si < pl.succ_off[pl.win_base[seg] + win + 1]; ++si) {
            const uint32_t s2 = pl.succ[si] >> 16, w2 = pl.succ[si] & 0xFFFFu;
            const uint32_t c = ++depcnt[pl.item_base[s2] + mp * pl.n_windows[s2] + w2];
            if (c == pl.need[pl.win_base[s2] + w2] * (t + 1)) { ready.push_back(((unsigned long long)(mp | (t << 16)) << 32) | pl.succ[si]); ++pushed; }
          }
      if ((int)seg == n_seg - 1 && (int)t < L - 1 && ++tails[mp] == pl.n_windows[seg] * (t + 1)) {
        // all split-K parts of the row pair are in: both 128-row tiles run their momentum tail
        for (uint32_t tile = 0; tile < LOOP_TAIL_NEED; ++tile)
          for (uint32_t w2 = 0; w2 < pl.n_windows[0]; ++w2) {
            const uint32_t c = ++depcnt[pl.item_base[0] + mp * pl.n_windows[0] + w2];
            if (c == pl.need[pl.win_base[0] + w2] * (t + 1)) { ready.push_back(((unsigned long long)(mp | ((t + 1) << 16)) << 32) | w2); ++pushed; }
          }
      }
      max_ready = std::max(max_ready, ready.size());
    }
    if (executed != loop_total_items(pl, L, false)) return fail("replay: not every item ran (an item never became ready)");
    if (pushed != executed) return fail("replay: queue entries and executed items differ");
    if (4 * (max_ready + (size_t)pl.n_pairs) > pl.q_cap) return fail("replay: the ready queue may hold more than a quarter of its capacity");
  }
  return 0;
}

}  // namespace dgan
