// CTA-pair (cta_group::2) pixel-graph GEMM: record formats, PTX wrappers and the host-side step builder.
//
// ncu on the first tensor-core kernel (profiles/r1a_*) showed it bound by L2->SM delivery (~7-8 TB/s with the tensor
// pipe ~29 % busy): every CTA streamed its own copy of every weight tile.  Here two CTAs on the two SMs of a TPC form
// one MMA of M = 256 latent rows:
//   * each CTA stages its own 128-row activation tile A and only HALF of each weight tile (N/2 rows) - the tensor
//     cores read the other half from the peer's shared memory, so weight traffic per SM is halved;
//   * one CTA per SM owns all 512 TMEM columns: two buffers of 256 (the MMAs of item i+1 overlap the epilogue of item
//     i), each holding the 1-8 accumulators of a window of output pixels;
//   * operands live in a circular shared-memory ring of variable-size steps planned on the host: per CTA pair one
//     contiguous stream of step records.
// Roles per CTA: TMA producer warp / MMA issuer warp / 8 epilogue warps; only the even (leader) CTA issues
// tcgen05.mma.cta_group::2, commits are multicast to both CTAs, TMA completions of both CTAs land on the leader's
// "full" barrier (peer-bit mask), and the epilogue warps of both CTAs release the accumulators on the leader's
// "acc_empty" barrier.  The kernel itself (kernels_loop.cuh) runs the whole projection loop in one launch.
#pragma once
#include "kernels_tc.cuh"

namespace dgan {

constexpr int TC2_SMEM_MAX = 232448;         // 227 KB opt-in limit per CTA
constexpr int TC2_TILE_BYTES = 128 * 128;    // one 128-row x 64-channel fp16 tile (TMA box, 128B swizzle)
constexpr int TC2_BUF_COLS = 256;            // TMEM columns per accumulator buffer (2 buffers: MMA i+1 overlaps epilogue i)
constexpr int TC2_EPI_WARPS = 8;             // two epilogue warps per TMEM lane quarter
constexpr int TC2_THREADS = 64 + 32 * TC2_EPI_WARPS;

struct __align__(16) TcItem2 {
  uint16_t q[16];
  uint32_t n_acc, step_beg, n_steps, pad;
};

// TMEM columns between the accumulators of one window.  N <= 32 (MNIST last layer: 16 outputs per block) packs
// 8 accumulators into a 256-column buffer, so a window can span a whole row of blocks.
__host__ __device__ constexpr int tc2_acc_stride(int n_tile) { return n_tile <= 32 ? 32 : (n_tile < 64 ? 64 : n_tile); }

// Does this instantiation stage its output (and ReLU-mask) tiles through shared memory + TMA?
__host__ __device__ constexpr bool tc2_tma_epilogue(int n_tile, int epi, int out_bytes) {
  return out_bytes == 2 && n_tile >= 64 && epi != EPI_FINAL_SIGMOID1 && epi != EPI_FINAL_TANH3;
}

// One step of a CTA pair's work stream (32 bytes).  The host concatenates, per CTA pair, the steps of all the items
// assigned to it (LPT order), so producer and MMA warps read one contiguous array: 16 records per coalesced
// warp load, staged in shared memory, the next batch always in flight - no table-load stalls on the issue path.
//
// A step stages up to 4 input-pixel (A) tiles and up to 8 weight half-tiles (B slots) for one k-chunk into a
// variable-size region of a circular shared-memory ring (offset chosen by the host, which simulates the ring),
// then issues up to 12 MMAs that combine them.  Several A tiles per step let one weight tile serve several input
// pixels (stride-2 transposed conv: outputs of equal parity use the same tap with neighbouring inputs), which
// is what the L2->SM byte count - the limiter of these kernels - cares about.
//
// producer record (per cluster rank):
//   w[0]: ring offset / 1 KB [0,8) | k-chunk [8,12) | A tiles [12,15) | B slots [15,19) | dep [19,23)
//         dep = D: the region overlaps that of step k-D (or D = 8, barrier-slot reuse): wait until step k-D is consumed
//   w[1]: row pair mp [0,16) | segment (layer-direction) [16,20) | first step of an item [20,21): wait for its dependencies
//   w[2..3]: 4 x u16 input pixel of A tile i
//   w[4..5]: 8 x u8 per B slot: weight tile [0,5) | half (row offset N/2) [5,6)
// MMA record:
//   w[0]: ring offset / 1 KB [0,8) | A tiles [8,11) | ops [11,16) | flags [16,18): 1 = first step of an item, 2 = last
//   w[1]: segment [0,4)
//   w[2..7]: 12 x u16 per MMA: A tile [0,2) | first B slot [2,5) | slots - 1 [5,7) | accumulator [7,10) | first MMA into it [10,11)
struct __align__(16) TcRec { uint32_t w[8]; };
constexpr int TC2_MAX_A = 4, TC2_MAX_BSLOTS = 8, TC2_MAX_OPS = 12, TC2_NSLOT = 8;
// Merged-N groups: when one input pixel feeds g accumulators that sit side by side in TMEM (acc, acc+1, ...) through
// weight tiles nobody else in the step uses, the g tiles are staged back to back and ONE MMA of N = g * N_TILE
// updates all of them.  With cta_group::2 the merged B operand [W_0 | W_1 | ...] is split in halves across the pair:
// CTA r stages half-tiles x = r*g + j (j < g) of the sequence W_0.lo, W_0.hi, W_1.lo, ... - hence per-rank producer
// streams.  g = 1 reduces to "each CTA stages its half of the tile".
constexpr int TC2_REC_BATCH = 16;
constexpr int TC2_STAGING_BYTES = 2 * TC2_REC_BATCH * (int)sizeof(TcRec);   // producer + MMA warp rings

// Output staging tiles per CTA: one per epilogue half (double-buffering them was measured: no gain).
__host__ __device__ constexpr int tc2_epi_tiles(int n_tile, int epi, int out_bytes) {
  return tc2_tma_epilogue(n_tile, epi, out_bytes) ? 2 : 0;
}
__host__ __device__ constexpr int tc2_ring_bytes(int n_tile, int epi, int out_bytes) {
  const int epi_b = tc2_epi_tiles(n_tile, epi, out_bytes) * TC2_TILE_BYTES;
  const int raw = ((TC2_SMEM_MAX - 1024 - 256 - TC2_STAGING_BYTES - epi_b) / 1024) * 1024;
  return raw > 255 * 1024 ? 255 * 1024 : raw;        // ring offsets are 8-bit KB
}

namespace ptx {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load executed by either CTA of the pair into its OWN shared memory; the transaction bytes
// are credited to the barrier of the even CTA (peer bit cleared).
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {   // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_load_3d_local(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_shared_v2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(local_bar), "r"(cta) : "memory");
}
}  // namespace ptx

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct TcWeights2 {
  CUtensorMap tm_b;            // box {64, N/2, 1}: one CTA's half of a weight tile
  PairTable tab;               // host copy of the layer's (input pixel, tap) contributions
  int h_grid = 0, w_grid = 0, max_acc = 1;
};

static int tc2_maxb(int N) { return TC2_BUF_COLS / tc2_acc_stride(N); }

// host-side description of one step (same for every row pair; ring offset and dep are filled per CTA-pair stream)
struct Tc2HostStep {
  int kc = 0, nA = 0, nB = 0, n_ops = 0;
  int a_pix[TC2_MAX_A] = {0, 0, 0, 0};
  uint8_t b_ent[2][TC2_MAX_BSLOTS] = {{0}, {0}};
  uint16_t ops[TC2_MAX_OPS] = {0};
  int n_tile_mmas = 0;         // un-merged count (statistics)
  int bytes = 0;               // operand bytes staged per CTA
};
struct Tc2HostItem {
  TcItem2 hdr{};
  std::vector<Tc2HostStep> steps;
  double stage_bytes = 0.0;
};

// Steps of one window (accumulator a <-> output pixel qs[a]).  Input pixels are taken in ascending order and packed
// greedily into steps of <= max_a A tiles (max_a = 1: one input pixel per step); a weight tile needed by several
// pixels of a step is staged once.  (Re-using the weight tiles of the PREVIOUS step as well was measured in round 1:
// 5-15 % fewer bytes, but 2 % slower - less ring capacity in flight, fewer merged-N MMAs - and is gone.)  Within a pixel, runs of consecutive accumulators whose tiles nobody else in
// the step uses (and whose first-MMA flags agree) become one merged-N MMA.
static void tc2_build_item(const PairTable& tab, const std::vector<int>& qs, int N, int K, int max_g, int max_a,
                           int step_max_bytes, Tc2HostItem* out) {
  const int kch = K / 64, half_b = (N / 2) * 128;
  out->hdr = TcItem2{};
  out->hdr.n_acc = (uint32_t)qs.size();
  for (size_t a = 0; a < qs.size(); ++a) out->hdr.q[a] = (uint16_t)qs[a];
  out->steps.clear();
  std::vector<std::pair<int, std::vector<std::pair<int, int>>>> by_p;   // pixel -> (tile, acc), sorted by acc
  for (size_t a = 0; a < qs.size(); ++a)
    for (int e = tab.off[qs[a]]; e < tab.off[qs[a] + 1]; ++e) {
      const int p = tab.pairs[e].x, t = tab.pairs[e].y;
      size_t g = 0;
      for (; g < by_p.size(); ++g)
        if (by_p[g].first == p) break;
      if (g == by_p.size()) by_p.push_back({p, {}});
      by_p[g].second.push_back({t, (int)a});
    }
  std::sort(by_p.begin(), by_p.end(), [](const auto& l, const auto& r) { return l.first < r.first; });
  for (auto& g : by_p)
    std::stable_sort(g.second.begin(), g.second.end(), [](const auto& l, const auto& r) { return l.second < r.second; });
  // a pixel with more entries than one step can hold is split (Linear layers: 16 tiles per input "pixel")
  std::vector<std::pair<int, std::vector<std::pair<int, int>>>> px;
  const int ent_cap = std::min({TC2_MAX_BSLOTS, TC2_MAX_OPS, std::max(1, (step_max_bytes - TC_A_BYTES) / half_b)});
  for (auto& g : by_p)
    for (size_t b0 = 0; b0 < g.second.size(); b0 += (size_t)ent_cap)
      px.push_back({g.first, std::vector<std::pair<int, int>>(g.second.begin() + b0,
                                                               g.second.begin() + std::min(g.second.size(), b0 + (size_t)ent_cap))});
  // ---- phase 1: greedy groups
  struct Group { size_t i0, i1; std::vector<int> staged; };
  std::vector<Group> groups;
  {
    size_t i0 = 0;
    while (i0 < px.size()) {
      size_t i1 = i0;
      std::vector<int> staged;      // tiles this group loads itself
      int n_ent = 0;
      auto have = [&](int t) { return std::find(staged.begin(), staged.end(), t) != staged.end(); };
      while (i1 < px.size() && (int)(i1 - i0) < max_a) {
        int fresh = 0;
        std::vector<int> fresh_tiles;
        for (auto& ta : px[i1].second)
          if (!have(ta.first) && std::find(fresh_tiles.begin(), fresh_tiles.end(), ta.first) == fresh_tiles.end()) {
            fresh_tiles.push_back(ta.first); ++fresh;
          }
        const int nA = (int)(i1 - i0) + 1, nB = (int)staged.size() + fresh;
        const bool dup_pixel = (i1 > i0 && px[i1].first == px[i1 - 1].first);   // split halves of one pixel stay apart
        if (i1 > i0 && (dup_pixel || nB > TC2_MAX_BSLOTS || n_ent + (int)px[i1].second.size() > TC2_MAX_OPS ||
                        nA * TC_A_BYTES + nB * half_b > step_max_bytes))
          break;
        for (int t : fresh_tiles) staged.push_back(t);
        n_ent += (int)px[i1].second.size();
        ++i1;
      }
      groups.push_back({i0, i1, staged});
      i0 = i1;
    }
  }
  // ---- phase 2: ops + B slots.  A tile is "single use" (mergeable into an N = g*N_TILE MMA) only if no other pixel of
  //      its group needs it in the plain half-per-CTA layout.
  uint32_t seen = 0;
  for (size_t gi = 0; gi < groups.size(); ++gi) {
    const Group& G = groups[gi];
    std::vector<int> use(32, 0);
    for (size_t i = G.i0; i < G.i1; ++i)
      for (auto& ta : px[i].second) ++use[ta.first];
    Tc2HostStep st;
    st.nA = (int)(G.i1 - G.i0);
    int slot_of[32];
    for (int t = 0; t < 32; ++t) slot_of[t] = -1;
    for (size_t i = G.i0; i < G.i1; ++i) {
      st.a_pix[i - G.i0] = px[i].first;
      const auto& ent = px[i].second;
      for (size_t e = 0; e < ent.size();) {
        const int acc0 = ent[e].second, t0 = ent[e].first;
        const bool f0 = !(seen & (1u << acc0));
        size_t g = 1;
        int slot;
        if (use[t0] == 1) {
          while ((int)g < max_g && e + g < ent.size() && ent[e + g].second == acc0 + (int)g && use[ent[e + g].first] == 1 &&
                 std::find(G.staged.begin(), G.staged.end(), ent[e + g].first) != G.staged.end() &&
                 (!(seen & (1u << ent[e + g].second))) == f0)
            ++g;
          slot = st.nB;
          for (int r = 0; r < 2; ++r)
            for (size_t jj = 0; jj < g; ++jj) {
              const size_t x = (size_t)r * g + jj;              // half-tile index in W_0.lo, W_0.hi, W_1.lo, ...
              st.b_ent[r][slot + jj] = (uint8_t)((ent[e + x / 2].first & 0x1F) | ((x & 1) << 5));
            }
          st.nB += (int)g;
        } else if (slot_of[t0] >= 0) {
          slot = slot_of[t0];
        } else {
          slot = slot_of[t0] = st.nB;
          for (int r = 0; r < 2; ++r) st.b_ent[r][slot] = (uint8_t)((t0 & 0x1F) | (r << 5));
          st.nB += 1;
        }
        st.ops[st.n_ops++] = (uint16_t)((i - G.i0) | (slot << 2) | ((g - 1) << 5) | (acc0 << 7) | ((f0 ? 1 : 0) << 10));
        st.n_tile_mmas += (int)g;
        for (size_t jj = 0; jj < g; ++jj) seen |= 1u << ent[e + jj].second;
        e += g;
      }
    }
    st.bytes = st.nA * TC_A_BYTES + st.nB * half_b;
    out->steps.push_back(st);
  }
  // k-chunk outermost: every accumulator then sums its (k-chunk, input pixel) contributions in one canonical order
  // - ascending k-chunk, ascending pixel - whatever the window shape and step grouping, so results do not depend
  // on the batch size (the schedule does) and a sharded batch reproduces the unsharded one bit for bit.
  const size_t n_groups = out->steps.size();
  for (int kc = 1; kc < kch; ++kc)
    for (size_t gi = 0; gi < n_groups; ++gi) {
      Tc2HostStep sk = out->steps[gi];
      sk.kc = kc;
      for (int o = 0; o < sk.n_ops; ++o) sk.ops[o] &= (uint16_t)~(1u << 10);
      out->steps.push_back(sk);
    }
  out->stage_bytes = 0.0;
  for (auto& stp : out->steps) out->stage_bytes += stp.bytes;
}

// Windows of wh x ww accumulators with strides (sy, sx) over the output grid.  Stride 2 gathers outputs of equal
// parity of a stride-2 transposed conv: they use the same taps with neighbouring inputs, so weight tiles are shared.
static void tc2_enumerate_windows(int h_grid, int w_grid, int wh, int ww, int sy, int sx, std::vector<std::vector<int>>* wins) {
  wins->clear();
  for (int by = 0; by < h_grid; by += wh * sy)
    for (int bx = 0; bx < w_grid; bx += ww * sx)
      for (int ry = 0; ry < sy; ++ry)
        for (int rx = 0; rx < sx; ++rx) {
          std::vector<int> qs;
          for (int i = 0; i < wh; ++i)
            for (int j = 0; j < ww; ++j) {
              const int y = by + ry + i * sy, x = bx + rx + j * sx;
              if (y < h_grid && x < w_grid) qs.push_back(y * w_grid + x);
            }
          if (!qs.empty()) wins->push_back(qs);
        }
}

static int tc2_build_direction(TcState& st, const TcWeights& w1, TcWeights2* w2, const PairTable& tab, int h_grid,
                               int w_grid, int force_max_acc) {
  const int N = w1.N, K = w1.K;
  int max_acc = tc2_maxb(N);
  if (force_max_acc > 0) max_acc = std::min(max_acc, force_max_acc);
  w2->tab = tab; w2->h_grid = h_grid; w2->w_grid = w_grid; w2->max_acc = max_acc;
  return tc_make_map(st, &w2->tm_b, w1.w, (uint64_t)K, (uint64_t)N, (uint64_t)w1.n_tiles, (uint32_t)(N / 2));
}

// One segment's slice of a plan in the form the validator below reads (kernels_loop.cuh builds these views).
struct Tc2Plan {
  int n_slots = 0, n_pairs = 0;
  std::vector<TcItem2> hdrs;
  std::vector<TcRec> stream_p[2], stream_m;
  std::vector<uint32_t> stream_off;
  std::vector<int> eitems;       // [n_slots][n_pairs] (window << 16 | row pair), -1 = none
};

// Independent validation of a plan against the pair table it was built from (host only; used by
// dgan_debug_check_plans and the CPU tests).  Re-derives from the uploaded records alone:
//  * every (output pixel, input pixel, tap, k-chunk) contribution of every item happens exactly once, into the right
//    accumulator, with the weight half-tiles each CTA stages forming exactly the operand the MMA reads;
//  * the first MMA into an accumulator - and only that one - overwrites it;
//  * every accumulator sums in the canonical order (k-chunk major, input pixel ascending): results then do not
//    depend on the schedule (batch-size / sharding invariance);
//  * ring safety: when a step's loads may start (step k - dep consumed), no earlier step that can still be read
//    overlaps its region, regions stay inside the ring, dep <= number of barrier slots;
//  * every (window, row pair) item is assigned to exactly one CTA pair.
static int tc2_check_plan(int N, int K, const PairTable& tab, int n_mpairs, int ring_bytes, const Tc2Plan& pl, std::string* err,
                          bool check_ring = true) {
  auto fail = [&](const std::string& m) { *err = m; return DGAN_ERR_INVALID_ARG; };
  const int kch = K / 64, half_b = (N / 2) * 128, acc_stride = tc2_acc_stride(N), max_acc = TC2_BUF_COLS / acc_stride;
  const size_t n_pairs = (size_t)pl.n_pairs;
  if (pl.stream_off.size() != n_pairs + 1) return fail("stream_off size");
  if (pl.stream_p[0].size() != pl.stream_m.size() || pl.stream_p[1].size() != pl.stream_m.size()) return fail("stream sizes differ");
  if (pl.eitems.size() != (size_t)pl.n_slots * n_pairs) return fail("eitems size");
  std::vector<int> assigned(pl.hdrs.size() * (size_t)n_mpairs, 0);
  for (const TcItem2& h : pl.hdrs) {
    if (h.n_acc < 1 || (int)h.n_acc > max_acc) return fail("window with too many accumulators");
    for (uint32_t a = 0; a < h.n_acc; ++a)
      if ((size_t)h.q[a] + 1 >= tab.off.size()) return fail("window pixel out of range");
  }
  struct Step { int beg, end, nB, kc; uint8_t b0[8], b1[8]; };
  for (size_t pr = 0; pr < n_pairs; ++pr) {
    const uint32_t r_beg = pl.stream_off[pr], r_end = pl.stream_off[pr + 1];
    if (r_beg > r_end || r_end > pl.stream_m.size()) return fail("stream_off not monotone");
    std::vector<Step> steps;
    int item_k = -1, win = -1, mp = -1;
    bool in_item = false;
    uint32_t seen = 0;
    std::vector<std::pair<int, int>> last_kp;                     // per accumulator: last (kc, p)
    std::vector<std::vector<std::pair<int, int>>> contrib;        // per accumulator: (p * 32 + tile, kc)
    for (uint32_t ri = r_beg; ri < r_end; ++ri) {
      const TcRec &m = pl.stream_m[ri], &p0 = pl.stream_p[0][ri], &p1 = pl.stream_p[1][ri];
      const int k = (int)steps.size();
      Step st{};
      st.beg = (int)(p0.w[0] & 0xFF);
      const int nA = (int)((p0.w[0] >> 12) & 7), nB = (int)((p0.w[0] >> 15) & 0xF), dep = (int)((p0.w[0] >> 19) & 0xF);
      st.kc = (int)((p0.w[0] >> 8) & 0xF); st.nB = nB;
      if (p1.w[0] != p0.w[0] || p1.w[1] != p0.w[1] || p1.w[2] != p0.w[2] || p1.w[3] != p0.w[3]) return fail("producer records of the two ranks disagree");
      if ((int)(m.w[0] & 0xFF) != st.beg || (int)((m.w[0] >> 8) & 7) != nA) return fail("MMA record disagrees with the producer record");
      if (nA < 1 || nA > TC2_MAX_A || nB > TC2_MAX_BSLOTS || st.kc >= kch) return fail("step field out of range");
      st.end = st.beg + (nA * TC_A_BYTES + nB * half_b + 1023) / 1024;
      if (st.end * 1024 > ring_bytes) return fail("step region outside the ring");
      if (dep < 1 || dep > TC2_NSLOT) return fail("dep out of range");
      for (int b = 0; b < 8; ++b) { st.b0[b] = (uint8_t)(p0.w[4 + b / 4] >> (8 * (b & 3))); st.b1[b] = (uint8_t)(p1.w[4 + b / 4] >> (8 * (b & 3))); }
      const uint32_t flags = (m.w[0] >> 16) & 3u;
      const int n_ops = (int)((m.w[0] >> 11) & 0x1F);
      if (n_ops > TC2_MAX_OPS) return fail("too many ops in a step");
      if (flags & 1u) {
        if (in_item) return fail("item starts inside an item");
        in_item = true; ++item_k;
        if (item_k >= pl.n_slots) return fail("more items than slots");
        const int e = pl.eitems[(size_t)item_k * n_pairs + pr];
        if (e < 0) return fail("stream has an item the epilogue list lacks");
        win = e >> 16; mp = e & 0xFFFF;
        if ((size_t)win >= pl.hdrs.size() || mp >= n_mpairs) return fail("item index out of range");
        if (assigned[(size_t)win * n_mpairs + mp]++) return fail("item assigned twice");
        seen = 0;
        last_kp.assign(pl.hdrs[win].n_acc, {-1, -1});
        contrib.assign(pl.hdrs[win].n_acc, {});
      }
      if (!in_item) return fail("step outside an item");
      if ((int)(p0.w[1] & 0xFFFF) != mp) return fail("row pair of a step differs from its item");
      const TcItem2& hdr = pl.hdrs[win];
      for (int oi = 0; oi < n_ops; ++oi) {
        const uint32_t e = (m.w[2 + oi / 2] >> (16 * (oi & 1))) & 0xFFFFu;
        const int a_idx = e & 3, slot = (e >> 2) & 7, g = ((e >> 5) & 3) + 1, acc0 = (e >> 7) & 7;
        const bool first = (e >> 10) & 1, prev = (e >> 11) & 1;
        if (a_idx >= nA) return fail("op reads an A tile the step does not stage");
        if (acc0 + g > (int)hdr.n_acc) return fail("op writes past the window's accumulators");
        if (g > 1 && (acc_stride != N || g * N > 256)) return fail("merged MMA too wide");
        const int p = (int)((p0.w[2 + a_idx / 2] >> (16 * (a_idx & 1))) & 0xFFFF);
        int tiles[4];
        if (prev) return fail("op refers to a previous step's weight tiles (not supported)");
        {
          if (slot + g > nB) return fail("op reads a B slot the step does not stage");
          for (int i = 0; i < g; ++i) {
            const int x0 = 2 * i, x1 = 2 * i + 1;
            const uint8_t* lo = (x0 / g) ? st.b1 : st.b0; const uint8_t* hi = (x1 / g) ? st.b1 : st.b0;
            const uint8_t el = lo[slot + x0 % g], eh = hi[slot + x1 % g];
            if ((el & 0x20) != 0 || (eh & 0x20) == 0 || (el & 0x1F) != (eh & 0x1F)) return fail("staged weight halves do not form the MMA operand");
            tiles[i] = el & 0x1F;
          }
        }
        for (int i = 0; i < g; ++i) {
          const int acc = acc0 + i;
          const bool unseen = !(seen & (1u << acc));
          if (first != unseen) return fail(first ? "overwrite of a live accumulator" : "accumulate into an uninitialised accumulator");
          const std::pair<int, int> kp{st.kc, p};
          if (!(last_kp[acc] < kp)) return fail("accumulation order is not canonical (k-chunk major, pixel ascending)");
          last_kp[acc] = kp;
          contrib[acc].push_back({p * 32 + tiles[i], st.kc});
        }
        for (int i = 0; i < g; ++i) seen |= 1u << (acc0 + i);
      }
      // ring safety inside the segment (the cyclic check over a CTA pair's whole stream is loop_check_plan's)
      for (int c = k - 1; check_ring && c >= 0 && c >= k - 4 * TC2_NSLOT; --c) {
        const Step& o = steps[(size_t)c];
        if (!(o.beg < st.end && st.beg < o.end)) continue;
        if (c > k - dep) return fail("ring hazard: a region may be overwritten while it can still be read");
      }
      steps.push_back(st);
      if (flags & 2u) {
        for (uint32_t a = 0; a < hdr.n_acc; ++a) {
          std::vector<std::pair<int, int>> want;
          for (int kc = 0; kc < kch; ++kc)
            for (int e2 = tab.off[hdr.q[a]]; e2 < tab.off[hdr.q[a] + 1]; ++e2) want.push_back({tab.pairs[e2].x * 32 + tab.pairs[e2].y, kc});
          std::vector<std::pair<int, int>> got = contrib[a];
          std::sort(want.begin(), want.end()); std::sort(got.begin(), got.end());
          if (want != got) return fail("an item's MMAs do not cover exactly its pair list");
        }
        in_item = false;
      }
    }
    if (in_item) return fail("stream ends inside an item");
    for (int kk = item_k + 1; kk < pl.n_slots; ++kk)
      if (pl.eitems[(size_t)kk * n_pairs + pr] != -1) return fail("epilogue list has an item the stream lacks");
  }
  for (int v : assigned)
    if (v != 1) return fail("an item is not assigned to any CTA pair");
  return 0;
}

}  // namespace dgan
