// Tensor-core path, version 2: CTA-pair (cta_group::2) pixel-graph GEMM.
//
// ncu on version 1 (profiles/r1a_*) showed the tcgen05 kernels bound by L2->SM delivery
// (~7-8 TB/s with the tensor pipe ~29 % busy): every CTA streamed its own copy of every weight
// tile.  Here two CTAs on the two SMs of a TPC form one MMA of M = 256 latent rows:
//   * each CTA stages its own 128-row activation tile A and only HALF of each weight tile
//     (N/2 rows) - the tensor cores read the other half from the peer's shared memory, so
//     weight traffic per SM is halved;
//   * one CTA per SM owns all 512 TMEM columns: two buffers of 256 (the MMAs of item i+1 overlap the
//     epilogue of item i), each holding the 1-8 accumulators of a window of output pixels;
//   * operands live in a circular shared-memory ring of variable-size steps planned on the host
//     (tc2_get_schedule): per CTA pair one contiguous stream of step records, LPT-assigned.
// Roles per CTA: TMA producer warp / MMA issuer warp / 8 epilogue warps; only the even (leader) CTA
// issues tcgen05.mma.cta_group::2, commits are multicast to both CTAs, TMA completions of both CTAs
// land on the leader's "full" barrier (peer-bit mask), and the epilogue warps of both CTAs release
// the accumulators on the leader's "acc_empty" barrier.  The epilogue warps are independent of each other: each
// converts the 32 rows it can read from TMEM and stores them with its own TMA store; what a unit needs from global
// memory (output pixel, bias row, mask word) is fetched ahead of time (profiles/r3_epilogue_stalls.md).
#pragma once
#include "kernels_tc.cuh"

namespace dgan {

constexpr int TC2_SMEM_MAX = 232448;         // 227 KB opt-in limit per CTA
constexpr int TC2_TILE_BYTES = 128 * 128;    // one 128-row x 64-channel fp16 tile (TMA box, 128B swizzle)
constexpr int TC2_BUF_COLS = 256;            // TMEM columns per accumulator buffer (2 buffers: MMA i+1 overlaps epilogue i)
constexpr int TC2_EPI_WARPS = 8;             // two epilogue warps per TMEM lane quarter
constexpr int TC2_THREADS = 64 + 32 * TC2_EPI_WARPS;
constexpr int TC2_STORE_ROWS = 32;           // rows per output TMA store: each epilogue warp stores the 32 rows it holds

// Where each CTA pair's work starts, passed in the kernel's parameter space (constant bank): the first records and the
// first item's header are then ONE global round trip away from the kernel's entry, and that round trip overlaps the
// barrier set-up and the TMEM allocation (the prologue of the last CTAs to start sits on the critical path between two
// kernels of the chain).
constexpr int TC2_MAX_PAIRS = 80;
struct Tc2Heads {
  uint32_t off[TC2_MAX_PAIRS + 1];    // record offsets of the pairs' step streams
  int first[TC2_MAX_PAIRS];           // first item (window << 16 | row pair) of each pair, -1 = none
};

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
//   w[1]: row pair mp [0,16)
//   w[2..3]: 4 x u16 input pixel of A tile i
//   w[4..5]: 8 x u8 per B slot: weight tile [0,5) | half (row offset N/2) [5,6)
// MMA record:
//   w[0]: ring offset / 1 KB [0,8) | A tiles [8,11) | ops [11,16) | flags [16,18): 1 = first step of an item, 2 = last
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

// Output staging tiles per CTA: one per epilogue half (two per half - the next tile written while the store of the
// previous one still reads shared memory - was measured in round 1: no gain).
__host__ __device__ constexpr int tc2_epi_tiles(int n_tile, int epi, int out_bytes) {
  return tc2_tma_epilogue(n_tile, epi, out_bytes) ? 2 : 0;
}
// The item's bias row staged in shared memory by the TMA-store epilogues: N_TILE floats; a per-pixel bias (the Linear:
// N_TILE = 256, one accumulator per item) changes from item to item and is double-buffered by item parity.
// (Sized per instantiation: the operand ring of the N = 64 / 128 layers must stay at 192 KB = 4 steps of 48 KB.)
__host__ __device__ constexpr int tc2_bias_bytes(int n_tile, int epi, int out_bytes) {
  return (tc2_tma_epilogue(n_tile, epi, out_bytes) && (epi == EPI_BIAS_RELU || epi == EPI_BIAS)) ? (n_tile == 256 ? 2048 : n_tile * 4) : 0;
}
__host__ __device__ constexpr int tc2_ring_bytes(int n_tile, int epi, int out_bytes) {
  const int epi_b = tc2_epi_tiles(n_tile, epi, out_bytes) * TC2_TILE_BYTES;
  const int raw = ((TC2_SMEM_MAX - 1024 - 256 - tc2_bias_bytes(n_tile, epi, out_bytes) - TC2_STAGING_BYTES - epi_b) / 1024) * 1024;
  return raw > 255 * 1024 ? 255 * 1024 : raw;
}

template <int N_TILE, int EPI = EPI_NONE, int OUT_BYTES = 2>
struct Tc2Cfg {
  static constexpr int HALF_B = (N_TILE / 2) * 128;                      // bytes of this CTA's half weight tile
  static constexpr int ACC_STRIDE = tc2_acc_stride(N_TILE);
  static constexpr int MAXB = TC2_BUF_COLS / ACC_STRIDE;                  // = accumulators per window (8 / 4 / 2 / 1)
  static constexpr bool TMA_EPI = tc2_tma_epilogue(N_TILE, EPI, OUT_BYTES);
  static constexpr int EPI_TILES = tc2_epi_tiles(N_TILE, EPI, OUT_BYTES);   // output staging tiles (0 or 2)
  static constexpr int EPI_BYTES = EPI_TILES * TC2_TILE_BYTES;
  static constexpr int RING_BYTES = tc2_ring_bytes(N_TILE, EPI, OUT_BYTES);          // operand ring (offsets are 8-bit KB)
  static constexpr int SMEM_BYTES = RING_BYTES + EPI_BYTES + TC2_STAGING_BYTES + 1024 + 256 + tc2_bias_bytes(N_TILE, EPI, OUT_BYTES);
};

namespace ptx {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Execution barrier of the pair without the GPU-scope fence of a releasing arrive (end of the kernel: nothing another CTA
// reads depends on it - shared-memory and TMEM lifetimes are ordered by the mbarriers and tcgen05 fences).
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
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

// k-th item (window << 16 | row pair) of CTA pair `pair` from the host-computed table [slot][pair] (-1 = no more work).
// The host assigns items largest-first to the least-loaded pair (LPT) with the cost model of tc2_get_schedule.
__device__ __forceinline__ int tc2_item_at(const int* __restrict__ order, int k, int pair, int n_pairs, int n_slots) {
  return k < n_slots ? __ldg(order + (size_t)k * n_pairs + pair) : -1;
}

#ifdef DGAN_PROBE
// Developer build only (-DDGAN_PROBE): per kernel instantiation and CTA, summed over launches: cycles from the PDL wait to
// the end of the CTA's work, launches, cycles from kernel entry to the PDL wait, cycles the MMA warp waited for operands.
// [4] cycles of the set-up (kernel entry to the PDL trigger), [5] / [6] %globaltimer (ns) at kernel entry / at the end of
// the CTA's work in the LAST launch, [7] %globaltimer when the CTA's first operands had landed (leaders, last launch).
__device__ unsigned long long g_tc2_probe[48][160][8];
__device__ __forceinline__ unsigned long long probe_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__host__ __device__ constexpr int tc2_probe_key(int n_tile, int epi, int out_bytes) {
  return (n_tile == 256 ? 0 : n_tile == 128 ? 1 : n_tile == 64 ? 2 : n_tile == 48 ? 3 : 4) * 8 + (out_bytes == 4 ? 6 : (epi < 4 ? epi : epi - 4));
}
#endif

template <int N_TILE, int EPI, typename TOUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
tc_bsgemm2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                  const __grid_constant__ CUtensorMap tm_out,
                  const TcItem2* __restrict__ items, const TcRec* __restrict__ stream_p0, const TcRec* __restrict__ stream_p1,
                  const TcRec* __restrict__ stream_m, const __grid_constant__ Tc2Heads heads,
                  const int* __restrict__ eitems, int n_slots,
                  TOUT* __restrict__ out, int n_pad, const float* __restrict__ bias, int bias_pstride, const TcFinalArgs fa) {
  using Cfg = Tc2Cfg<N_TILE, EPI, (int)sizeof(TOUT)>;
  constexpr bool TMA_EPI = Cfg::TMA_EPI;
  constexpr int HALF_B = Cfg::HALF_B, ACC_STRIDE = Cfg::ACC_STRIDE;
  extern __shared__ uint8_t smem_raw[];
#ifdef DGAN_PROBE
  const long long probe_t_start = clock64();
  const unsigned long long probe_g_start = probe_gtime();
#endif
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + Cfg::RING_BYTES;            // output staging tiles: EPI_TILES / 2 per epilogue half
  const uint32_t stg_base = epi_base + Cfg::EPI_BYTES;              // [producer ring][MMA ring] of TcRec
  const uint32_t bar_base = stg_base + TC2_STAGING_BYTES;
  // full[s] @ +8s (s<8), empty[s] @ +64+8s, acc_full[2] @ +128, acc_empty[2] @ +144, tmem slot @ +160, momentum-tail flag @ +200
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 64, bar_acc_full = bar_base + 128, bar_acc_empty = bar_base + 144;
  const uint32_t tmem_slot = bar_base + 160;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  // The schedule tables are constants: each role's first entries are requested before anything else, so that their
  // latency overlaps the set-up below (and, for CTAs that start early, the previous kernel's tail).
  uint32_t rbeg = 0, rend = 0;
  uint4 mine = make_uint4(0, 0, 0, 0);
  int item_first = -1;
  constexpr bool HAS_BIAS = TMA_EPI && (EPI == EPI_BIAS_RELU || EPI == EPI_BIAS);
  const int et = (warp - 2) * 32 + lane;              // 0..255 over the epilogue threads
  uint32_t q_next = 0, nacc_next = 0;                 // epilogue: output pixel of accumulator `lane` (lanes 0..15), accumulator count
  float bias_next = 0.f;                              // epilogue: element `et` of the next item's bias row
  const TcRec* __restrict__ stream = warp == 1 ? stream_m : (rank ? stream_p1 : stream_p0);
  if (warp <= 1) {
    rbeg = heads.off[pair]; rend = heads.off[pair + 1];
    if (2 * rbeg + lane < 2 * rend) mine = __ldg(reinterpret_cast<const uint4*>(stream + rbeg) + lane);   // lane = 16-byte half
  } else {
    item_first = heads.first[pair];
    if (item_first >= 0) {           // header of the first item (see the epilogue): tables and bias are constants too
      const TcItem2* ip0 = items + (item_first >> 16);
      if (lane < 16) q_next = (uint32_t)__ldg(&ip0->q[lane]);
      nacc_next = __ldg(&ip0->n_acc);
      if (HAS_BIAS && bias_pstride == 0 && et < N_TILE) bias_next = __ldg(bias + et);
    }
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_b);
    for (int s = 0; s < TC2_NSLOT; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);    // leader's producer arrive.expect_tx (bytes of both CTAs)
      ptx::mbar_init(bar_empty + 8 * s, 1);   // one multicast commit per CTA
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_acc_full + 8 * b, 1);
      ptx::mbar_init(bar_acc_empty + 8 * b, 2 * TC2_EPI_WARPS);   // epilogue warps of both CTAs (used on the leader only)
    }
    if (TMA_EPI) ptx::prefetch_tmap(&tm_out);
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
  if (HAS_BIAS && warp >= 2 && bias_pstride != 0 && item_first >= 0 && et < N_TILE)   // per-pixel bias: row of the first item's pixel
    bias_next = __ldg(bias + (size_t)__shfl_sync(0xffffffffu, q_next, 0) * bias_pstride + et);
  // everything above overlapped the previous kernel's tail (PDL); from here on we read what it wrote
#ifdef DGAN_PROBE
  const long long probe_t_entry = clock64();
#endif
  pdl_launch_dependents();
  pdl_wait();
#ifdef DGAN_PROBE
  const long long probe_t_go = clock64();
  long long probe_wait_full = 0;
#endif

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    // The whole warp walks the step list convergently; every table value is loaded from a
    // warp-uniform address so the TMA operands live in uniform registers (no per-instruction
    // R2UR/ELECT loop), and one elected lane issues.
    uint32_t it = 0;
    const uint32_t ring = stg_base;
    for (uint32_t base = rbeg; base < rend; base += TC2_REC_BATCH) {
      ptx::st_shared_v4(ring + lane * 16u, mine.x, mine.y, mine.z, mine.w);
      __syncwarp();
      if (2 * (base + TC2_REC_BATCH) + lane < 2 * rend) mine = __ldg(reinterpret_cast<const uint4*>(stream + base + TC2_REC_BATCH) + lane);
      const uint32_t cnt = min((uint32_t)TC2_REC_BATCH, rend - base);
      for (uint32_t i = 0; i < cnt; ++i, ++it) {
        const uint4 r0 = ptx::ld_shared_v4(ring + i * 32u);
        const uint2 r1 = ptx::ld_shared_v2(ring + i * 32u + 16u);
        const uint32_t slot = it & (TC2_NSLOT - 1);
        const int kc = (r0.x >> 8) & 0xF, nA = (r0.x >> 12) & 0x7, nB = (r0.x >> 15) & 0xF;
        const uint32_t dep = (r0.x >> 19) & 0xF;
        const int row0 = (2 * (int)(r0.y & 0xFFFFu) + (int)rank) * kRowTile;
        if (it >= dep) ptx::mbar_wait(bar_empty + 8 * ((it - dep) & (TC2_NSLOT - 1)), ((it - dep) >> 3) & 1);   // step it-dep consumed
        // implied by the wait above (steps are consumed in order); observing every phase of this slot exactly once
        // before it is re-armed keeps the barrier protocol checkable (compute-sanitizer synccheck)
        if (dep != TC2_NSLOT && it >= TC2_NSLOT) ptx::mbar_wait(bar_empty + 8 * slot, ((it - TC2_NSLOT) >> 3) & 1);
        const uint32_t full = bar_full + 8 * slot;
        const uint32_t sa = smem_base + ((r0.x & 0xFFu) << 10);
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_expect_tx(full, 2u * (uint32_t)(nA * TC_A_BYTES + nB * HALF_B));
#pragma unroll
          for (int a = 0; a < TC2_MAX_A; ++a) {
            if (a >= nA) break;
            const int p = (int)((((a < 2) ? r0.z : r0.w) >> (16 * (a & 1))) & 0xFFFFu);
            ptx::tma_load_3d_2sm(sa + a * TC_A_BYTES, &tm_a, full, kc * 64, row0, p);
          }
          const uint32_t sb = sa + nA * TC_A_BYTES;
#pragma unroll
          for (int b = 0; b < TC2_MAX_BSLOTS; ++b) {
            if (b >= nB) break;
            const uint32_t e = ((b < 4) ? r1.x : r1.y) >> (8 * (b & 3));
            ptx::tma_load_3d_2sm(sb + b * HALF_B, &tm_b, full, kc * 64, (int)((e >> 5) & 1u) * (N_TILE / 2), (int)(e & 0x1Fu));
          }
        }
        __syncwarp();
      }
      __syncwarp();
    }
    // drain: the last steps' "consumed" signals are otherwise never observed (nobody leaves while MMAs still read smem)
    for (uint32_t j = it > TC2_NSLOT ? it - TC2_NSLOT : 0; j < it; ++j) ptx::mbar_wait(bar_empty + 8 * (j & (TC2_NSLOT - 1)), (j >> 3) & 1);
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_f16(256, N_TILE);
      uint32_t it = 0, item_count = 0;
      const uint32_t ring = stg_base + TC2_REC_BATCH * (uint32_t)sizeof(TcRec);
      const uint64_t desc0 = make_smem_desc_sw128(smem_base);
      const uint32_t desc_lo0 = (uint32_t)desc0, desc_hi = (uint32_t)(desc0 >> 32);
      uint32_t buf = 0;
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
            buf = item_count & 1;
            ptx::mbar_wait(bar_acc_empty + 8 * buf, ((item_count >> 1) & 1) ^ 1);
          }
#ifdef DGAN_PROBE
          const long long probe_w0 = clock64();
#endif
          ptx::mbar_wait(bar_full + 8 * slot, phase);
#ifdef DGAN_PROBE
          probe_wait_full += clock64() - probe_w0;
          if (it == 0 && lane == 0) g_tc2_probe[tc2_probe_key(N_TILE, EPI, (int)sizeof(TOUT))][blockIdx.x][7] = probe_gtime();
#endif
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
              const uint32_t first = (e >> 10) & 1u;
              const uint32_t a_lo = a_lo0 + (e & 3u) * (uint32_t)(TC_A_BYTES >> 4);
              const uint32_t b_lo = b_lo0 + ((e >> 2) & 7u) * (uint32_t)(HALF_B >> 4);
              const uint32_t d = d0 + ((e >> 7) & 7u) * ACC_STRIDE;
              const uint32_t idg = idesc + ((e >> 5) & 3u) * ((uint32_t)(N_TILE >> 3) << 17);   // N = slots * N_TILE
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_f16_2sm(d, ((uint64_t)desc_hi << 32) | (a_lo + 2u * k), ((uint64_t)desc_hi << 32) | (b_lo + 2u * k), idg,
                                  (k > 0 || !first) ? 1u : 0u);
            }
            ptx::umma_commit_2sm(bar_empty + 8 * slot);           // this step is consumed (both CTAs)
            if (flags & 2u) ptx::umma_commit_2sm(bar_acc_full + 8 * buf);   // last step: accumulators complete in both CTAs
          }
          __syncwarp();
          if (flags & 2u) ++item_count;
        }
        __syncwarp();
      }
      // drain: observe the release of the last (up to two) accumulator buffers by the epilogue warps of both CTAs
      for (uint32_t j = item_count > 2 ? item_count - 2 : 0; j < item_count; ++j) ptx::mbar_wait(bar_acc_empty + 8 * (j & 1), (j >> 1) & 1);
    }
  } else {
    // ===================== epilogue (warps 2..9, both CTAs) =====================
    // Two warps per TMEM lane quarter split the (accumulator, 32-column chunk) units of an item;
    // the TMEM load of a warp's next unit is in flight while it converts and stores the current one.
    // Nothing on the per-unit path depends on a global load issued in the same unit (ncu source counters of the
    // round-2 build: a third of the epilogue's time went to the chain item header -> output pixel -> bias / mask word):
    // the window's output pixels live in lanes 0..15 (one shuffle per use), fetched one item ahead; the item's bias row
    // is staged in shared memory before the accumulators are awaited; mask words are fetched two units ahead.
    const int lq = warp & 3;                          // TMEM lanes this warp may access
    const int half = (warp - 2) >> 2;                 // 0 | 1: which of the two warps of this quarter
    const int row = lq * 32 + lane;
    const uint32_t bias_s = bar_base + 256;           // [2][256] floats
    uint32_t item_count = 0;
    for (int kk = 0, item_e = item_first, item_next; item_e >= 0; ++kk, ++item_count, item_e = item_next) {
      item_next = tc2_item_at(eitems, kk + 1, pair, n_pairs, n_slots);      // (window << 16 | row pair), one item ahead
      const int mp = item_e & 0xFFFF;
      const uint32_t q_mine = q_next;
      const int n_acc = (int)nacc_next;
      auto q_of = [&](int a) { return (int)__shfl_sync(0xffffffffu, q_mine, a); };
      if (HAS_BIAS && (item_count == 0 || bias_pstride != 0)) {
        // a per-pixel bias (Linear) comes with one accumulator per item (N_TILE = 256): the row of q[0] serves the item
        if (et < N_TILE) ptx::st_shared_u32(bias_s + (item_count & 1u) * 1024u + (uint32_t)et * 4u, __float_as_uint(bias_next));
        ptx::named_bar_sync(4, 32 * TC2_EPI_WARPS);   // also: every warp is done with the row of item_count - 2
      }
      if (item_next >= 0) {                            // next item's header, in flight during this item
        const TcItem2* ipn = items + (item_next >> 16);
        if (lane < 16) q_next = (uint32_t)__ldg(&ipn->q[lane]);
        nacc_next = __ldg(&ipn->n_acc);
        if (HAS_BIAS && bias_pstride != 0 && et < N_TILE)
          bias_next = __ldg(bias + (size_t)__shfl_sync(0xffffffffu, q_next, 0) * bias_pstride + et);
      }
      const size_t n = (size_t)(2 * mp + (int)rank) * kRowTile + row;
      const uint32_t buf = item_count & 1;
      const uint32_t tbuf = tmem_base + ((uint32_t)(lq * 32) << 16) + buf * TC2_BUF_COLS;
      constexpr bool FINAL = (EPI == EPI_FINAL_SIGMOID1 || EPI == EPI_FINAL_TANH3);
      float4 xq_next[FINAL ? (EPI == EPI_FINAL_SIGMOID1 ? 4 : 12) : 1];
      if (FINAL && half < n_acc)     // first block's target pixels: in flight while the MMAs finish
        tc_final_targets<(EPI == EPI_FINAL_SIGMOID1 ? 1 : 3)>(reinterpret_cast<float4(&)[EPI == EPI_FINAL_SIGMOID1 ? 4 : 12]>(xq_next), fa, q_of(half < n_acc ? half : 0), (int)n);
      // EPI_MASK: mask words are fetched two units ahead; the first two are in flight while the MMAs finish
      unsigned long long mbits = ~0ull, mbits_next = ~0ull, mbits_next2 = ~0ull;
      if (TMA_EPI && EPI == EPI_MASK) {
        constexpr int G0 = N_TILE >= 64 ? N_TILE / 64 : 1;
        const int nu = n_acc * G0, u1 = half + 2;
        const int q0 = q_of(half < nu ? half / G0 : 0), q1 = q_of(u1 < nu ? u1 / G0 : 0);
        if (half < nu) mbits_next = __ldg(fa.mb_in + ((size_t)q0 * n_pad + n) * G0 + half % G0);
        if (u1 < nu) mbits_next2 = __ldg(fa.mb_in + ((size_t)q1 * n_pad + n) * G0 + u1 % G0);
      }
      ptx::mbar_wait(bar_acc_full + 8 * buf, (item_count >> 1) & 1);
      ptx::tc_fence_after();
      if (EPI == EPI_FINAL_SIGMOID1 || EPI == EPI_FINAL_TANH3) {
        constexpr int CO = (EPI == EPI_FINAL_SIGMOID1) ? 1 : 3;
        float4 xq[4 * CO];
        for (int a = half; a < n_acc; a += 2) {
#pragma unroll
          for (int j = 0; j < 4 * CO; ++j) xq[j] = xq_next[j];
          const int qa = q_of(a), qa2 = q_of(a + 2 < n_acc ? a + 2 : a);
          if (a + 2 < n_acc) tc_final_targets<CO>(reinterpret_cast<float4(&)[4 * CO]>(xq_next), fa, qa2, (int)n);   // next block's targets in flight
          const uint32_t taddr = tbuf + (uint32_t)(a * ACC_STRIDE);
          if (EPI == EPI_FINAL_SIGMOID1)
            tc_final_epilogue<1, ACT_SIGMOID>(taddr, fa, bias, qa, (int)n, n_pad, reinterpret_cast<__half*>(out),
                                              reinterpret_cast<const float4(&)[4]>(xq));
          else
            tc_final_epilogue<3, ACT_TANH>(taddr, fa, bias, qa, (int)n, n_pad, reinterpret_cast<__half*>(out),
                                           reinterpret_cast<const float4(&)[12]>(xq));
        }
      } else if (TMA_EPI) {
        // ---- 64-column units through shared memory: TMEM -> regs -> (bias|ReLU|mask) -> fp16 -> this warp's 32-row
        //      slice of a 128B-swizzled tile -> one TMA store of 32 rows x 64 channels per warp.  The warps of a half
        //      share nothing: no block-level barrier on the unit path, each warp waits for its own previous store.
        constexpr int G = N_TILE >= 64 ? N_TILE / 64 : 1;   // 64-column groups per accumulator
        const int n_units = n_acc * G;
        const int row0 = (2 * mp + (int)rank) * kRowTile + lq * 32;
        const uint32_t swz = (uint32_t)(lane & 7);
        const uint32_t s_out = epi_base + (uint32_t)half * TC2_TILE_BYTES + (uint32_t)lq * 4096u;
        const uint32_t bias_row = bias_s + ((bias_pstride != 0) ? (item_count & 1u) * 1024u : 0u);
        uint32_t r0[32], r1[32];
        if (half < n_units) {
          const int a = half / G, g = half % G;
          ptx::tmem_ld32(tbuf + (uint32_t)(a * ACC_STRIDE + g * 64), r0);
          ptx::tmem_ld32(tbuf + (uint32_t)(a * ACC_STRIDE + g * 64 + 32), r1);
        }
        for (int u = half; u < n_units; u += 2) {
          const int a = u / G, g = u % G, q = q_of(a);
          mbits = mbits_next; mbits_next = mbits_next2;
          ptx::tmem_ld_wait();
          uint32_t pk[32];
          unsigned long long bits = 0ull;
          {
            float v[64];
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
            if (HAS_BIAS) {
#pragma unroll
              for (int j4 = 0; j4 < 16; ++j4) {
                const uint4 b = ptx::ld_shared_v4(bias_row + (uint32_t)(g * 256 + j4 * 16));
                v[j4 * 4 + 0] += __uint_as_float(b.x); v[j4 * 4 + 1] += __uint_as_float(b.y);
                v[j4 * 4 + 2] += __uint_as_float(b.z); v[j4 * 4 + 3] += __uint_as_float(b.w);
              }
              if (EPI == EPI_BIAS_RELU) {
#pragma unroll
                for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f);
              }
            }
            if (EPI == EPI_BIAS_RELU && fa.mb_out != nullptr) {
#pragma unroll
              for (int j = 0; j < 64; ++j) bits |= (unsigned long long)(v[j] > 0.f) << j;
            }
            if (EPI == EPI_MASK) {
#pragma unroll
              for (int j = 0; j < 64; ++j)
                if (!((mbits >> j) & 1ull)) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) pk[j] = pack_half2(v[2 * j], v[2 * j + 1]);
          }
          if (u + 2 < n_units) {                           // next unit's accumulator columns: in flight during the store phase
            const int a2 = (u + 2) / G, g2 = (u + 2) % G;
            ptx::tmem_ld32(tbuf + (uint32_t)(a2 * ACC_STRIDE + g2 * 64), r0);
            ptx::tmem_ld32(tbuf + (uint32_t)(a2 * ACC_STRIDE + g2 * 64 + 32), r1);
          }
          if (lane == 0) ptx::bulk_wait_read0();           // this warp's previous store has read its slice
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 8; ++c)
            ptx::st_shared_v4(s_out + (uint32_t)lane * 128u + (((uint32_t)c ^ swz) << 4), pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_3d(&tm_out, s_out, g * 64, row0, q);
            ptx::bulk_commit();
          }
          // global traffic of this unit after the proxy fence (which waits for the thread's outstanding accesses)
          if (EPI == EPI_BIAS_RELU && fa.mb_out != nullptr) fa.mb_out[((size_t)q * n_pad + n) * G + g] = bits;
          if (EPI == EPI_MASK && u + 4 < n_units) {
            const int a4 = (u + 4) / G, g4 = (u + 4) % G;
            mbits_next2 = __ldg(fa.mb_in + ((size_t)q_of(a4) * n_pad + n) * G + g4);
          }
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
          tc_store_chunk<N_TILE, EPI, TOUT>(rA, q_of(u / CH), (u % CH) * 32, n, n_pad, out, bias, bias_pstride);
          if (u + 2 < n_units) {
            ptx::tmem_ld_wait();
            if (u + 4 < n_units) ptx::tmem_ld32(tbuf + (uint32_t)(((u + 4) / CH) * ACC_STRIDE + ((u + 4) % CH) * 32), rA);
            tc_store_chunk<N_TILE, EPI, TOUT>(rB, q_of((u + 2) / CH), ((u + 2) % CH) * 32, n, n_pad, out, bias, bias_pstride);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_remote(bar_acc_empty + 8 * buf, 0);
      if (EPI == EPI_NONE && sizeof(TOUT) == 4 && fa.m_counter != nullptr) {
        // ---- momentum in the tail of the split-K Linear backward.  Every epilogue thread has stored its share of this
        //      item's partial sums; the CTA that completes the last partial of its 128-row tile applies the update
        //      (same arithmetic and summation order as momentum_kernel: parts 0, 1, 2, ...).
        const uint32_t flag_addr = bar_base + 200;
        const unsigned rt = 2u * (unsigned)mp + rank;
        __threadfence();
        ptx::named_bar_sync(3, 32 * TC2_EPI_WARPS);
        if (warp == 2 && lane == 0) {
          const unsigned ticket = atomicAdd(fa.m_counter + rt, 1u);
          ptx::st_shared_u32(flag_addr, ticket == (unsigned)TC_LINEAR_SPLIT - 1u ? 1u : 0u);
        }
        ptx::named_bar_sync(3, 32 * TC2_EPI_WARPS);
        if (ptx::ld_shared_u32(flag_addr) != 0u) {
          __threadfence();
          const float* __restrict__ gp = reinterpret_cast<const float*>(out);
          const size_t base = (size_t)rt * kRowTile * N_TILE;
          const int tid = (warp - 2) * 32 + lane;
          // 4 float4 positions per thread in flight at a time (the loop is latency-bound: 6 L2 reads per position)
          constexpr int STRIDE = 4 * 32 * TC2_EPI_WARPS, UNR = 4;
          for (int e0 = tid * 4; e0 < kRowTile * N_TILE; e0 += UNR * STRIDE) {
            float4 gs[UNR], vv[UNR], zz[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
              const size_t i = base + (size_t)(e0 + u * STRIDE);
              gs[u] = __ldcg(reinterpret_cast<const float4*>(gp + i));
              vv[u] = *reinterpret_cast<const float4*>(fa.mv + i);
              zz[u] = *reinterpret_cast<const float4*>(fa.mz + i);
            }
            float4 t[TC_LINEAR_SPLIT - 1][UNR];      // all partial sums of the 4 positions in flight together
#pragma unroll
            for (int pp = 1; pp < TC_LINEAR_SPLIT; ++pp)
#pragma unroll
              for (int u = 0; u < UNR; ++u)
                t[pp - 1][u] = __ldcg(reinterpret_cast<const float4*>(gp + base + (size_t)(e0 + u * STRIDE) + (size_t)pp * fa.m_count));
#pragma unroll
            for (int pp = 1; pp < TC_LINEAR_SPLIT; ++pp)      // fixed order: parts 0, 1, 2, 3 (as momentum_kernel)
#pragma unroll
              for (int u = 0; u < UNR; ++u) { gs[u].x += t[pp - 1][u].x; gs[u].y += t[pp - 1][u].y; gs[u].z += t[pp - 1][u].z; gs[u].w += t[pp - 1][u].w; }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
              const size_t i = base + (size_t)(e0 + u * STRIDE);
              float4 v4 = vv[u], z4 = zz[u];
              v4.x = fmaf(fa.m_mu, v4.x, fa.m_gmul * gs[u].x); v4.y = fmaf(fa.m_mu, v4.y, fa.m_gmul * gs[u].y);
              v4.z = fmaf(fa.m_mu, v4.z, fa.m_gmul * gs[u].z); v4.w = fmaf(fa.m_mu, v4.w, fa.m_gmul * gs[u].w);
              z4.x -= fa.m_lr * v4.x; z4.y -= fa.m_lr * v4.y; z4.z -= fa.m_lr * v4.z; z4.w -= fa.m_lr * v4.w;
              *reinterpret_cast<float4*>(fa.mv + i) = v4;
              *reinterpret_cast<float4*>(fa.mz + i) = z4;
              if (fa.mz_h != nullptr)
                *reinterpret_cast<uint2*>(fa.mz_h + i) = make_uint2(pack_half2(z4.x, z4.y), pack_half2(z4.z, z4.w));
            }
          }
          if (tid == 0) fa.m_counter[rt] = 0u;             // ready for the next launch
        }
      }
    }
    if (TMA_EPI && lane == 0) ptx::bulk_wait_read0();   // shared memory no longer read; the writes complete with the grid
  }

#ifdef DGAN_PROBE
  {
    constexpr int key = tc2_probe_key(N_TILE, EPI, (int)sizeof(TOUT));
    if (warp == 1 && lane == 0) atomicAdd(&g_tc2_probe[key][blockIdx.x][3], (unsigned long long)probe_wait_full);
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicAdd(&g_tc2_probe[key][blockIdx.x][0], (unsigned long long)(clock64() - probe_t_go));
      atomicAdd(&g_tc2_probe[key][blockIdx.x][1], 1ull);
      atomicAdd(&g_tc2_probe[key][blockIdx.x][2], (unsigned long long)(probe_t_go - probe_t_entry));
      atomicAdd(&g_tc2_probe[key][blockIdx.x][4], (unsigned long long)(probe_t_entry - probe_t_start));
      g_tc2_probe[key][blockIdx.x][5] = probe_g_start;
      g_tc2_probe[key][blockIdx.x][6] = probe_gtime();
    }
  }
#endif
  ptx::tc_fence_before();
  ptx::cluster_sync_relaxed();   // the leader's MMAs read the peer's shared memory and TMEM is freed for the pair: nobody leaves early
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct Tc2Schedule {           // one window tiling of a layer-direction + its item -> CTA-pair assignment, uploaded
  TcItem2* items = nullptr;
  TcRec* stream_p[2] = {nullptr, nullptr};   // producer records per cluster rank; per CTA pair: its items' steps, concatenated
  TcRec* stream_m = nullptr;       // MMA records, same indexing
  Tc2Heads heads{};                // per pair: record offsets into the streams, first item (kernel parameter)
  int* eitems = nullptr;           // [n_slots][n_pairs] (window << 16 | row pair) for the epilogue warps, or -1
  int n_slots = 0, n_pairs = 0;
  int n_windows = 0;
  int wh = 0, ww = 0, sy = 1, sx = 1;
};
struct TcWeights2 {
  CUtensorMap tm_b;            // box {64, N/2, 1}
  PairTable tab;               // host copy: schedules are built lazily per batch size
  int h_grid = 0, w_grid = 0, max_acc = 1;
  mutable std::vector<std::pair<int, Tc2Schedule>> by_mpairs;   // chosen schedule per n_mpairs (lazy cache)
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
// pixels of a step is staged once.  Within a pixel, runs of consecutive accumulators whose tiles nobody else in
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
  // ---- phase 1: greedy groups.  (Re-using the weight tiles of the PREVIOUS step as well was measured in round 1:
  //      5-15 % fewer bytes, but 2 % slower - less ring capacity in flight, fewer merged-N MMAs - and is gone.)
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
                               int w_grid, int force_max_acc, std::vector<void*>* allocs, cudaStream_t s) {
  (void)allocs; (void)s;
  const int N = w1.N, K = w1.K;
  int max_acc = tc2_maxb(N);
  if (force_max_acc > 0) max_acc = std::min(max_acc, force_max_acc);
  w2->tab = tab; w2->h_grid = h_grid; w2->w_grid = w_grid; w2->max_acc = max_acc;
  return tc_make_map(st, &w2->tm_b, w1.w, (uint64_t)K, (uint64_t)N, (uint64_t)w1.n_tiles, (uint32_t)(N / 2));
}

#ifndef DGAN_STEP_MAX_KB
#define DGAN_STEP_MAX_KB 48
#endif
#ifndef DGAN_STEP_MAX_KB_N64
#define DGAN_STEP_MAX_KB_N64 64
#endif
#ifndef DGAN_COST_EPI_KB
#define DGAN_COST_EPI_KB 24.0
#endif
#ifndef DGAN_COST_FIXED_KB
#define DGAN_COST_FIXED_KB 48.0
#endif
#ifndef TC2_REFINE_BUDGET
#define TC2_REFINE_BUDGET (1LL << 26)    // candidate evaluations of the assignment refinement per window shape (tc2_plan)
#endif

// Pick (and build on first use) the window tiling for `n_mpairs` row pairs on `n_pairs` CTA pairs: every candidate
// shape (wh x ww accumulators, strides 1 or 2) is scored by an LPT assignment of its items (window, row pair) to the
// CTA pairs with cost = operand bytes staged + a per-accumulator epilogue charge + a fixed per-item charge;
// the smallest makespan wins.  Then each pair's items are concatenated into its step streams, the circular operand
// ring is simulated to give every step its offset and its dependency distance, and everything is uploaded.
struct Tc2Plan {               // host result of the planner (what tc2_get_schedule uploads)
  int shape[4] = {1, 1, 1, 1};   // wh, ww, sy, sx
  int n_slots = 0, n_pairs = 0;
  std::vector<TcItem2> hdrs;
  std::vector<TcRec> stream_p[2], stream_m;
  std::vector<uint32_t> stream_off;
  std::vector<int> eitems;
  long long n_mma = 0, n_single = 0, n_steps = 0, n_bytes = 0;
  double load_max = 0.0, load_mean = 0.0;   // cost-model load of the busiest CTA pair / the mean over pairs (balance of the LPT assignment)
};

static int tc2_plan(int N, int K, const PairTable& tab, int h_grid, int w_grid, int max_acc, int n_mpairs, int n_pairs,
                    int ring_bytes, Tc2Plan* plan) {
  const int max_g = (N >= 64) ? std::min(4, 256 / N) : 1;      // merged-N MMAs (see TC2_MAX_A above)
  const int max_a = TC2_MAX_A;
  // Step size: a step is consumed only once all of it has landed, so big steps cost pipeline depth (4 x 48 KB fit the
  // ring); measured on C2: 32 KB (= one A tile per step) 5359, 40 KB 5466, 48-56 KB 5660, 64 KB 5553, 96 KB 5385 images/s.
  // (second sweep, final round-2 epilogue: 40 KB 5679, 48 KB 6047 / 6004, 56 KB 6064, 64 KB 6041 - flat from 48 KB on, except that
  //  the N = 64, K = 128 layer (Generator.3 forward: 4 KB weight half-tiles, 3 activation tiles + their taps per 64 KB step)
  //  gains 2 - 3 us per launch with 64 KB steps while the N = 128 layers lose as much: configs[1] 5912 -> 5965 images/s on one
  //  box; CelebA's layer of that shape is indifferent: 1556 vs 1546)
  const int step_kb = (N == 64 && K == 128) ? DGAN_STEP_MAX_KB_N64 : DGAN_STEP_MAX_KB;
  const int step_max = std::min((ring_bytes / 2) & ~1023, step_kb * 1024);
  double best_cost = 1e300;
  int best_shape[4] = {1, 1, 1, 1};
  std::vector<Tc2HostItem> best_items;
  std::vector<std::vector<int>> best_lists;
  std::vector<std::vector<int>> wins;
  for (int wh = 1; wh <= 2; ++wh)
    for (int ww = 1; ww <= 8; ++ww)
      for (int sy = 1; sy <= (wh > 1 ? 2 : 1); ++sy)
        for (int sx = 1; sx <= (ww > 1 ? 2 : 1); ++sx) {
          if (wh * ww > max_acc || wh > h_grid || ww > std::max(w_grid, 1)) continue;
          tc2_enumerate_windows(h_grid, std::max(w_grid, 1), wh, ww, sy, sx, &wins);
          std::vector<Tc2HostItem> items(wins.size());
          for (size_t i = 0; i < wins.size(); ++i) tc2_build_item(tab, wins[i], N, K, max_g, max_a, step_max, &items[i]);
          std::stable_sort(items.begin(), items.end(), [](const Tc2HostItem& l, const Tc2HostItem& r) { return l.stage_bytes > r.stage_bytes; });
          std::vector<double> icost(items.size());
          for (size_t i = 0; i < items.size(); ++i)
            icost[i] = items[i].stage_bytes + DGAN_COST_EPI_KB * 1024.0 * items[i].hdr.n_acc * std::max(1, N / 64) + DGAN_COST_FIXED_KB * 1024.0;
          // LPT: items (window, mp) largest-first, each to the currently least-loaded CTA pair
          const long long total = (long long)items.size() * n_mpairs;
          std::vector<double> load((size_t)n_pairs, 0.0);
          std::vector<std::vector<int>> lists((size_t)n_pairs);
          for (long long idx = 0; idx < total; ++idx) {        // items[] is sorted by cost, mp is the fast index: cost-descending
            size_t best = 0;
            for (size_t pr = 1; pr < (size_t)n_pairs; ++pr)
              if (load[pr] < load[best]) best = pr;
            load[best] += icost[(size_t)(idx / n_mpairs)];
            lists[best].push_back((int)idx);
          }
          // Refinement: while the busiest pair can hand an item to - or swap one with - another pair so that both end
          // up below its load, do the best such move (LPT alone leaves e.g. 35 on a mean of 30.4 for Generator.2 bwd's
          // 160 items of cost 4..25).
          auto cost_of = [&](int idx) { return icost[(size_t)(idx / n_mpairs)]; };
          // One pass looks at |P| x (1 + |Q|) candidates for every other pair Q: quadratic in the items per pair.  With many
          // items per pair (large batches: 160 row pairs x 1024 windows) LPT alone is already within one small item of
          // the mean and the search would take minutes, so it runs on a budget of candidate evaluations that the
          // benchmarked sizes (<= 20 row pairs) never reach.
          long long work = 0;
          for (int iter = 0; iter < 4096 && work < TC2_REFINE_BUDGET; ++iter) {
            const size_t P = (size_t)(std::max_element(load.begin(), load.end()) - load.begin());
            double best_peak = load[P];
            size_t bq = P; int bi = -1, bj = -1;
            for (size_t Q = 0; Q < (size_t)n_pairs; ++Q) {
              if (Q == P) continue;
              work += (long long)lists[P].size() * (long long)(1 + lists[Q].size());
              for (size_t i = 0; i < lists[P].size(); ++i) {
                const double ci = cost_of(lists[P][i]);
                double peak = std::max(load[P] - ci, load[Q] + ci);           // move i: P -> Q
                if (peak < best_peak - 1e-9) { best_peak = peak; bq = Q; bi = (int)i; bj = -1; }
                for (size_t j = 0; j < lists[Q].size(); ++j) {                 // swap i <-> j
                  const double cj = cost_of(lists[Q][j]);
                  if (cj >= ci) continue;
                  peak = std::max(load[P] - ci + cj, load[Q] + ci - cj);
                  if (peak < best_peak - 1e-9) { best_peak = peak; bq = Q; bi = (int)i; bj = (int)j; }
                }
              }
            }
            if (bi < 0) break;
            const int it_i = lists[P][(size_t)bi];
            const double ci = cost_of(it_i);
            if (bj < 0) {
              lists[P].erase(lists[P].begin() + bi);
              lists[bq].push_back(it_i);
              load[P] -= ci; load[bq] += ci;
            } else {
              const int it_j = lists[bq][(size_t)bj];
              const double cj = cost_of(it_j);
              lists[P][(size_t)bi] = it_j; lists[bq][(size_t)bj] = it_i;
              load[P] += cj - ci; load[bq] += ci - cj;
            }
          }
          for (auto& l : lists)      // biggest first: a pair's last item is its smallest (shortest un-overlapped epilogue)
            std::stable_sort(l.begin(), l.end(), [&](int a, int b) { return cost_of(a) > cost_of(b); });
          const double makespan = *std::max_element(load.begin(), load.end());
          if (makespan < best_cost) {
            best_cost = makespan;
            plan->load_max = makespan;
            plan->load_mean = std::accumulate(load.begin(), load.end(), 0.0) / (double)n_pairs;
            best_shape[0] = wh; best_shape[1] = ww; best_shape[2] = sy; best_shape[3] = sx;
            best_items.swap(items); best_lists.swap(lists);
          }
        }
  plan->shape[0] = best_shape[0]; plan->shape[1] = best_shape[1]; plan->shape[2] = best_shape[2]; plan->shape[3] = best_shape[3];
  plan->n_pairs = n_pairs;
  size_t n_slots = 0;
  for (auto& l : best_lists) n_slots = std::max(n_slots, l.size());
  plan->n_slots = (int)n_slots;
  std::vector<int>& eitems = plan->eitems;
  eitems.assign(n_slots * (size_t)n_pairs, -1);
  std::vector<uint32_t>& stream_off = plan->stream_off;
  stream_off.assign((size_t)n_pairs + 1, 0);
  std::vector<TcRec>* stream_p = plan->stream_p;
  std::vector<TcRec>& stream_m = plan->stream_m;
  stream_p[0].clear(); stream_p[1].clear(); stream_m.clear();
  long long n_mma = 0, n_single = 0, n_steps = 0, n_bytes = 0;
  for (size_t pr = 0; pr < best_lists.size(); ++pr) {
    stream_off[pr] = (uint32_t)stream_m.size();
    // circular operand ring of this CTA pair: sequential allocation, wrap when the step does not fit
    std::vector<std::pair<int, int>> region;     // [begin, end) in KB of every step of this stream
    int cursor = 0;
    for (size_t k = 0; k < best_lists[pr].size(); ++k) {
      const int win = best_lists[pr][k] / n_mpairs, mp = best_lists[pr][k] % n_mpairs;
      if (win > 0x7FFF || mp > 0xFFFF) { set_error("tensor-core schedule limits exceeded"); return DGAN_ERR_UNSUPPORTED; }
      eitems[k * (size_t)n_pairs + pr] = (win << 16) | mp;
      const Tc2HostItem& itm = best_items[(size_t)win];
      for (size_t j = 0; j < itm.steps.size(); ++j) {
        const Tc2HostStep& hs = itm.steps[j];
        const int kb = (hs.bytes + 1023) / 1024;
        if (kb * 1024 > ring_bytes) { set_error("tensor-core step larger than the operand ring"); return DGAN_ERR_UNSUPPORTED; }
        if (cursor + kb > ring_bytes / 1024) cursor = 0;
        const int beg = cursor, end = cursor + kb;
        cursor = end;
        // The producer may overwrite a region once the step that used it is consumed: dep = distance to the latest
        // earlier step whose region overlaps this one (8 = barrier-slot reuse only).
        int dep = TC2_NSLOT;
        const int kidx = (int)region.size();
        for (int d = 1; d <= TC2_NSLOT && d <= kidx; ++d) {
          const auto& rg = region[(size_t)(kidx - d)];
          if (rg.first < end && beg < rg.second) { dep = d; break; }
        }
        region.push_back({beg, end});
        const uint32_t flags = (j == 0 ? 1u : 0u) | (j + 1 == itm.steps.size() ? 2u : 0u);
        TcRec rm{};
        rm.w[0] = (uint32_t)beg | ((uint32_t)hs.nA << 8) | ((uint32_t)hs.n_ops << 11) | (flags << 16);
        for (int o = 0; o < hs.n_ops; ++o) rm.w[2 + o / 2] |= (uint32_t)hs.ops[o] << (16 * (o & 1));
        stream_m.push_back(rm);
        for (int r = 0; r < 2; ++r) {
          TcRec rp{};
          rp.w[0] = (uint32_t)beg | ((uint32_t)hs.kc << 8) | ((uint32_t)hs.nA << 12) | ((uint32_t)hs.nB << 15) | ((uint32_t)dep << 19);
          rp.w[1] = (uint32_t)mp;
          for (int a = 0; a < hs.nA; ++a) rp.w[2 + a / 2] |= (uint32_t)(hs.a_pix[a] & 0xFFFF) << (16 * (a & 1));
          for (int b = 0; b < hs.nB; ++b) rp.w[4 + b / 4] |= (uint32_t)hs.b_ent[r][b] << (8 * (b & 3));
          stream_p[r].push_back(rp);
        }
        n_mma += hs.n_ops; n_single += hs.n_tile_mmas; n_steps += 1; n_bytes += hs.bytes;
      }
    }
  }
  stream_off[(size_t)n_pairs] = (uint32_t)stream_m.size();
  plan->hdrs.resize(best_items.size());
  for (size_t i = 0; i < best_items.size(); ++i) plan->hdrs[i] = best_items[i].hdr;
  plan->n_mma = n_mma; plan->n_single = n_single; plan->n_steps = n_steps; plan->n_bytes = n_bytes;
  return 0;
}

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
static int tc2_check_plan(int N, int K, const PairTable& tab, int n_mpairs, int ring_bytes, const Tc2Plan& pl, std::string* err) {
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
      // ring safety
      for (int c = k - 1; c >= 0 && c >= k - 4 * TC2_NSLOT; --c) {
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

// Pick (and build on first use) the schedule of one layer-direction for `n_mpairs` row pairs and upload it.
static int tc2_get_schedule(TcState& st, const TcWeights& w1, const TcWeights2& w2, int n_mpairs, int n_pairs, int ring_bytes,
                            std::vector<void*>* allocs, cudaStream_t s, const Tc2Schedule** out) {
  (void)st;
  for (auto& kv : w2.by_mpairs)
    if (kv.first == n_mpairs) { *out = &kv.second; return 0; }
  Tc2Plan plan;
  int rc;
  if ((rc = tc2_plan(w1.N, w1.K, w2.tab, w2.h_grid, w2.w_grid, w2.max_acc, n_mpairs, n_pairs, ring_bytes, &plan))) return rc;
  Tc2Schedule sc;
  sc.wh = plan.shape[0]; sc.ww = plan.shape[1]; sc.sy = plan.shape[2]; sc.sx = plan.shape[3];
  sc.n_windows = (int)plan.hdrs.size(); sc.n_pairs = n_pairs; sc.n_slots = plan.n_slots;
  if ((rc = tc_upload(allocs, plan.hdrs.data(), plan.hdrs.size() * sizeof(TcItem2), (void**)&sc.items, s))) return rc;
  for (int r = 0; r < 2; ++r)
    if ((rc = tc_upload(allocs, plan.stream_p[r].data(), plan.stream_p[r].size() * sizeof(TcRec), (void**)&sc.stream_p[r], s))) return rc;
  if ((rc = tc_upload(allocs, plan.stream_m.data(), plan.stream_m.size() * sizeof(TcRec), (void**)&sc.stream_m, s))) return rc;
  if (n_pairs > TC2_MAX_PAIRS) { set_error("more CTA pairs than the kernel's parameter block holds"); return DGAN_ERR_UNSUPPORTED; }
  for (int pr = 0; pr <= n_pairs; ++pr) sc.heads.off[pr] = plan.stream_off[(size_t)pr];
  for (int pr = 0; pr < n_pairs; ++pr) sc.heads.first[pr] = plan.n_slots > 0 ? plan.eitems[(size_t)pr] : -1;
  if ((rc = tc_upload(allocs, plan.eitems.data(), plan.eitems.size() * sizeof(int), (void**)&sc.eitems, s))) return rc;
  w2.by_mpairs.push_back({n_mpairs, sc});
  *out = &w2.by_mpairs.back().second;
  return 0;
}

template <int NT, int EP, typename TOUT>
static cudaError_t tc2_optin() {
  return cudaFuncSetAttribute(tc_bsgemm2_kernel<NT, EP, TOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              Tc2Cfg<NT, EP, (int)sizeof(TOUT)>::SMEM_BYTES);
}

static int tc2_optin_all() {
#define TC2_OPTIN(NT, EP, T) DGAN_CUDA_CHECK((tc2_optin<NT, EP, T>()))
  TC2_OPTIN(64, EPI_BIAS_RELU, __half); TC2_OPTIN(128, EPI_BIAS_RELU, __half); TC2_OPTIN(256, EPI_BIAS_RELU, __half);
  TC2_OPTIN(64, EPI_BIAS, __half); TC2_OPTIN(128, EPI_BIAS, __half); TC2_OPTIN(256, EPI_BIAS, __half);
  TC2_OPTIN(64, EPI_MASK, __half); TC2_OPTIN(128, EPI_MASK, __half); TC2_OPTIN(256, EPI_MASK, __half);
  TC2_OPTIN(64, EPI_NONE, __half); TC2_OPTIN(128, EPI_NONE, __half); TC2_OPTIN(256, EPI_NONE, __half);
  TC2_OPTIN(64, EPI_NONE, float); TC2_OPTIN(128, EPI_NONE, float); TC2_OPTIN(256, EPI_NONE, float);
  TC2_OPTIN(64, EPI_BIAS, float); TC2_OPTIN(128, EPI_BIAS, float); TC2_OPTIN(256, EPI_BIAS, float);   // use_bn: fp32 pre-activations
  TC2_OPTIN(16, EPI_FINAL_SIGMOID1, __half); TC2_OPTIN(48, EPI_FINAL_TANH3, __half);
#undef TC2_OPTIN
  return 0;
}

template <typename TOUT>
static int tc2_launch_impl(TcState& st, int64_t* launches, const TcWeights& w, const TcWeights2& w2m, const __half* in,
                           TOUT* out, int n_pad, int epi, const float* bias, cudaStream_t s, const TcFinalArgs* final_args = nullptr, const CUtensorMap* pre_a = nullptr,
                           const CUtensorMap* pre_out = nullptr) {
  TcFinalArgs fa{};
  if (final_args) fa = *final_args;
  CUtensorMap tm_a;
  int rc;
  if (pre_a != nullptr) tm_a = *pre_a;          // encoded once per workspace by the caller
  else if ((rc = tc_make_map(st, &tm_a, in, (uint64_t)w.K, (uint64_t)n_pad, (uint64_t)w.P_in, 128))) return rc;
  CUtensorMap tm_out = tm_a;                     // placeholder when unused
  if (tc2_tma_epilogue(w.N, epi, (int)sizeof(TOUT))) {
    if (pre_out != nullptr) tm_out = *pre_out;
    else if ((rc = tc_make_map(st, &tm_out, out, (uint64_t)w.N, (uint64_t)n_pad, (uint64_t)w.P_out, TC2_STORE_ROWS))) return rc;
  }
  if (n_pad % (2 * kRowTile) != 0) { set_error("pair kernel needs n_pad % 256 == 0"); return DGAN_ERR_INVALID_ARG; }
  if (w.bias_pstride != 0 && w.N != 256) { set_error("a per-pixel bias needs one accumulator per item (N = 256)"); return DGAN_ERR_UNSUPPORTED; }
  const int n_mpairs = n_pad / (2 * kRowTile);
  const Tc2Schedule* schp = nullptr;
  const int pairs_avail = st.num_sms / 2;
  const int ring_bytes = tc2_ring_bytes(w.N, epi, (int)sizeof(TOUT));
  if ((rc = tc2_get_schedule(st, w, w2m, n_mpairs, pairs_avail, ring_bytes, st.allocs, s, &schp))) return rc;
  const Tc2Schedule& w2s = *schp;
  const int grid = 2 * w2s.n_pairs;       // pairs without work find -1 in slot 0 and fall through
  cudaError_t le = cudaSuccess;
#define TC2_GO(NT, EP)                                                                                                 \
  le = launch_pdl(tc_bsgemm2_kernel<NT, EP, TOUT>, dim3(grid), dim3(TC2_THREADS), Tc2Cfg<NT, EP, (int)sizeof(TOUT)>::SMEM_BYTES, s, \
                  tm_a, w2m.tm_b, tm_out, w2s.items, w2s.stream_p[0], w2s.stream_p[1], w2s.stream_m, w2s.heads, w2s.eitems, w2s.n_slots, out, n_pad, bias, w.bias_pstride, fa)
#define TC2_GO_H(NT, EP)                                                                                               \
  le = launch_pdl(tc_bsgemm2_kernel<NT, EP, __half>, dim3(grid), dim3(TC2_THREADS), Tc2Cfg<NT, EP, 2>::SMEM_BYTES, s,   \
                  tm_a, w2m.tm_b, tm_out, w2s.items, w2s.stream_p[0], w2s.stream_p[1], w2s.stream_m, w2s.heads, w2s.eitems, w2s.n_slots, reinterpret_cast<__half*>(out), n_pad, bias, 0, fa)
#define TC2_BY_N(EP)                    \
  do {                                  \
    if (w.N == 64) TC2_GO(64, EP);      \
    else if (w.N == 128) TC2_GO(128, EP); \
    else TC2_GO(256, EP);               \
  } while (0)
  if (sizeof(TOUT) == 4) { if (epi == EPI_BIAS) TC2_BY_N(EPI_BIAS); else TC2_BY_N(EPI_NONE); }
  else if (epi == EPI_FINAL_SIGMOID1) { TC2_GO_H(16, EPI_FINAL_SIGMOID1); }
  else if (epi == EPI_FINAL_TANH3) { TC2_GO_H(48, EPI_FINAL_TANH3); }
  else if (epi == EPI_BIAS_RELU) { TC2_BY_N(EPI_BIAS_RELU); }
  else if (epi == EPI_BIAS) { TC2_BY_N(EPI_BIAS); }
  else if (epi == EPI_MASK) { TC2_BY_N(EPI_MASK); }
  else { TC2_BY_N(EPI_NONE); }
#undef TC2_BY_N
#undef TC2_GO
#undef TC2_GO_H
  (*launches)++;
  cudaError_t e = (le != cudaSuccess) ? le : cudaGetLastError();
  if (e != cudaSuccess) { set_error(std::string("tc_bsgemm2 launch: ") + cudaGetErrorString(e)); return DGAN_ERR_CUDA; }
  return 0;
}

}  // namespace dgan
