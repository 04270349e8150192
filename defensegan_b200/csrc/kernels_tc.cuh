// Tensor-core path (DGAN_PREC_FP16): shared pieces of the pixel-graph GEMM on tcgen05.
//
//   out[q][n][:] = epi( sum_{(p,t) in pairs(q)} in[p][n][:] x W_t )        fp16 operands, fp32 accumulate
//
// Activations are pixel-major [P][N rows][C] fp16, so one input pixel x 128 latent rows x 64 channels is a dense
// 16 KB box = one TMA box = one 128B-swizzled K-major UMMA operand; weights are stored per tap as [N][K] fp16
// tiles.  The conv geometry is data, not code: only in-bounds (input pixel, tap) pairs are listed, so exactly the
// algorithmic MACs are issued - no zero-insertion, no padding rows, no im2col buffer.
// This header holds the PTX wrappers, the UMMA descriptors, the epilogue arithmetic (bias / ReLU / mask / last layer +
// sigmoid|tanh + MSE + dL/dpre) and the weight re-layout; the CTA-pair kernel that uses them and its
// host-side planning are in kernels_tc2.cuh.
#pragma once
#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through the runtime (no -lcuda)

#include "common.cuh"

namespace dgan {

constexpr int TC_A_BYTES = 128 * 128;        // one activation tile: 128 rows x 64 fp16
constexpr int TC_LINEAR_SPLIT = 4;           // partial sums of the Linear backward (dz)

// One direction (forward or backward) of one layer on the tensor-core path.
struct TcWeights {
  __half* w = nullptr;          // [tiles][N][K] fp16, K contiguous
  int N = 0, K = 0, P_in = 0, P_out = 0, n_tiles = 0;
  int bias_pstride = 0;
};

struct TcState {
  float grad_scale = 64.f;      // fp16 gradient scaling (undone in the z update)
  void* encode_fn = nullptr;    // cuTensorMapEncodeTiled
  std::vector<void*>* allocs = nullptr;   // the handle's allocation list (lazily built schedules are freed with it)
  int num_sms = 148;
};

struct TcLayerSpec {
  int P_in, C_in, P_out, C_out, h_in, w_in, h_used, w_used;
  const PairTable *fwd, *bwd;
  const float *w_fwd_kmajor_src, *w_bwd_kmajor_src, *linear_W, *linear_Wt;
  TcWeights *out_f, *out_b;
  int bias_pstride;
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 x fp16 -> fp32, M=128, K=16
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
}  // namespace ptx

// K-major, 128B-swizzled operand tile: rows 128 B apart, 8-row groups 1024 B apart (SBO),
// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D = F32 (bits 4-5 = 1), A = B = F16 (0), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// fp32 pair -> packed fp16, round-to-nearest, saturating to +-65504: a backward activation that exceeds the fp16 range
// (trained filters, |y - x| up to 2 on CelebA) must not become inf and turn z into NaN for the rest of the loop.
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// Extra epilogue kinds of the tensor-core path: the generator's last layer (C_out <= 3) is run as
// a pixel-graph GEMM whose "output pixels" are 4x4 blocks of image pixels (N = 16*C_out columns =
// the block's pre-activations), so that the output non-linearity, the squared error, its
// derivative and the per-row loss partial are all lane-local in the epilogue
// (models/dataset_models.py:68-69,161-163; models/gan.py:411-414).
enum TcEpilogue : int { EPI_FINAL_SIGMOID1 = 8, EPI_FINAL_TANH3 = 9 };

struct TcFinalArgs {
  const float* x;        // [B][H*W*C] target images (NULL: forward only)
  float* y;              // [n_pad][H*W*C] G(z)
  float* loss_part;      // [n_blocks][n_pad] sum over the block of (y-x)^2 (written when write_y)
  int R, B, n_rows;      // restarts per image, images, valid latent rows
  int nbx, w_out;        // blocks per image row, image width
  int write_y;           // store G(z) (only the last iteration / forward-only calls consume it)
  float gscale;          // fp16 gradient scaling applied to dL/dpre
  // 1-bit ReLU masks of the CTA-pair path: bit j of word [(q*n_pad + n)*(N/64) + g] <=> out[q][n][g*64+j] > 0.
  // Written by the forward EPI_BIAS_RELU epilogue, read by the backward EPI_MASK epilogue
  // (tf.nn.relu's gradient passes where the forward output was > 0).
  unsigned long long* mb_out;
  const unsigned long long* mb_in;
  // split-K Linear backward with the momentum update (tf.train.MomentumOptimizer, models/gan.py:389-391) in its tail:
  // the CTA whose partial sums complete a 128-row tile (ticket from m_counter[tile]) applies
  // v <- mu v + gmul * sum(parts), z <- z - lr v for that tile
  float* mz; float* mv; __half* mz_h;   // z, velocity [n_pad][latent] fp32, fp16 copy of z
  float m_gmul, m_lr, m_mu;
  unsigned* m_counter;   // [n_pad / 128] arrival tickets, self-resetting; NULL = plain partial sums
  int m_nparts;          // partial sums per row tile
  size_t m_count;        // elements per partial-sum array (n_pad * latent)
};

// Target pixels (4x4 block of image n / R) of one row: loaded one accumulator ahead of their use.
template <int C_OUT>
__device__ __forceinline__ void tc_final_targets(float4 (&xq)[4 * C_OUT], const TcFinalArgs& fa, int blk, int n) {
  if (fa.x == nullptr) return;
  const int by = blk / fa.nbx, bx = blk % fa.nbx;
  const int hwc = fa.w_out * fa.w_out * C_OUT;
  const int img = min(n / fa.R, fa.B - 1);
#pragma unroll
  for (int li = 0; li < 4; ++li) {
    const size_t off = (size_t)((4 * by + li) * fa.w_out + 4 * bx) * C_OUT;
#pragma unroll
    for (int e4 = 0; e4 < C_OUT; ++e4) xq[li * C_OUT + e4] = __ldg(reinterpret_cast<const float4*>(fa.x + (size_t)img * hwc + off) + e4);
  }
}

template <int C_OUT, int ACT>
__device__ __forceinline__ void tc_final_epilogue(uint32_t taddr, const TcFinalArgs& fa, const float* __restrict__ bias,
                                                  int blk, int n, int n_pad, __half* __restrict__ dblk,
                                                  const float4 (&xq)[4 * C_OUT]) {
  constexpr int NV = 16 * C_OUT;
  const int by = blk / fa.nbx, bx = blk % fa.nbx;
  const int hwc = fa.w_out * fa.w_out * C_OUT;
  float v[NV];
#pragma unroll
  for (int c = 0; c < NV; c += 16) {
    uint32_t r[16];
    ptx::tmem_ld16(taddr + c, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[c + j] = __uint_as_float(r[j]);
  }
  float bsv[C_OUT];
#pragma unroll
  for (int co = 0; co < C_OUT; ++co) bsv[co] = __ldg(bias + co);
  float lsum = 0.f;
  uint32_t packed[NV / 2];   // the 16*C_OUT valid fp16 of this row of the block tensor
#pragma unroll
  for (int li = 0; li < 4; ++li) {
    const size_t off = (size_t)((4 * by + li) * fa.w_out + 4 * bx) * C_OUT;
    float yv[4 * C_OUT], dv[4 * C_OUT];
#pragma unroll
    for (int e = 0; e < 4 * C_OUT; ++e) {
      const float pre = v[li * 4 * C_OUT + e] + bsv[e % C_OUT];
      float yy, dact;
      if (ACT == ACT_SIGMOID) { yy = __fdividef(1.f, 1.f + __expf(-pre)); dact = yy * (1.f - yy); }
      else { const float t = __expf(-2.f * fabsf(pre)); yy = copysignf(__fdividef(1.f - t, 1.f + t), pre); dact = 1.f - yy * yy; }
      yv[e] = yy;
      float d = 0.f;
      if (fa.x != nullptr) {
        const float4 xr = xq[li * C_OUT + e / 4];
        const float xe = (e % 4 == 0) ? xr.x : (e % 4 == 1) ? xr.y : (e % 4 == 2) ? xr.z : xr.w;
        d = yy - xe;
        lsum = fmaf(d, d, lsum);
      }
      dv[e] = d * dact * fa.gscale;
    }
    if (fa.write_y) {
      float4* yp = reinterpret_cast<float4*>(fa.y + (size_t)n * hwc + off);
#pragma unroll
      for (int e4 = 0; e4 < C_OUT; ++e4) yp[e4] = make_float4(yv[e4 * 4], yv[e4 * 4 + 1], yv[e4 * 4 + 2], yv[e4 * 4 + 3]);
    }
#pragma unroll
    for (int e2 = 0; e2 < 2 * C_OUT; ++e2) packed[li * 2 * C_OUT + e2] = pack_half2(dv[2 * e2], dv[2 * e2 + 1]);
  }
  if (fa.x != nullptr) {
    uint4* dp = reinterpret_cast<uint4*>(dblk + ((size_t)blk * n_pad + n) * 64);
#pragma unroll
    for (int j4 = 0; j4 < NV / 8; ++j4)   // the K-padding columns [NV, 64) stay zero (cleared once per call)
      dp[j4] = make_uint4(packed[j4 * 4], packed[j4 * 4 + 1], packed[j4 * 4 + 2], packed[j4 * 4 + 3]);
    if (fa.write_y) fa.loss_part[(size_t)blk * n_pad + n] = lsum;   // the loss is consumed after the last forward only
  }
}

// One 32-column chunk of an accumulator (already in registers) -> epilogue -> out[q][n][c0..c0+32)
template <int N_TILE, int EPI, typename TOUT>
__device__ __forceinline__ void tc_store_chunk(const uint32_t (&r)[32], int q, int c0, size_t n, int n_pad,
                                               TOUT* __restrict__ out, const float* __restrict__ bias, int bias_pstride) {
  const size_t orow = ((size_t)q * n_pad + n) * N_TILE;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (EPI == EPI_BIAS_RELU || EPI == EPI_BIAS) {
    const float4* bp = reinterpret_cast<const float4*>(bias + (size_t)q * bias_pstride + c0);
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 b = __ldg(bp + j4);
      v[j4 * 4 + 0] += b.x; v[j4 * 4 + 1] += b.y; v[j4 * 4 + 2] += b.z; v[j4 * 4 + 3] += b.w;
    }
    if (EPI == EPI_BIAS_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
  }
  if (sizeof(TOUT) == 2) {
    uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(out) + orow + c0);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4)
      op[j4] = make_uint4(pack_half2(v[j4 * 8 + 0], v[j4 * 8 + 1]), pack_half2(v[j4 * 8 + 2], v[j4 * 8 + 3]),
                          pack_half2(v[j4 * 8 + 4], v[j4 * 8 + 5]), pack_half2(v[j4 * 8 + 6], v[j4 * 8 + 7]));
  } else {
    float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + orow + c0);
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) op[j4] = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// fp32 [tiles][rows][cols] -> fp16, same layout
__global__ void tc_convert_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}
// Linear backward tiles: out[q][k][c] = W[k][q*C + c]   (W is (latent, 16*C))
__global__ void tc_linear_bwd_tiles_kernel(const float* __restrict__ W, __half* __restrict__ out, int latent, int C,
                                           int n_pix) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)n_pix * latent * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int k = (int)((i / C) % latent);
  const int q = (int)(i / ((size_t)C * latent));
  out[i] = __float2half_rn(W[(size_t)k * n_pix * C + (size_t)q * C + c]);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D fp16 tensor [d2][d1][d0] (d0 contiguous), box {64, box1, 1}, 128B swizzle.
static int tc_make_map(const TcState& st, CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                       uint32_t box1) {
  if (st.encode_fn == nullptr) { set_error("cuTensorMapEncodeTiled unavailable"); return DGAN_ERR_CUDA; }
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
  const cuuint32_t box[3] = {64, box1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ((PFN_encodeTiled)st.encode_fn)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims,
                                               strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return DGAN_ERR_CUDA;
  }
  return 0;
}

static int tc_upload(std::vector<void*>* allocs, const void* host, size_t bytes, void** dev, cudaStream_t s) {
  DGAN_CUDA_CHECK(cudaMalloc(dev, bytes));
  allocs->push_back(*dev);
  DGAN_CUDA_CHECK(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, s));
  // the host vectors die when the builder returns: make sure the copy has been staged
  DGAN_CUDA_CHECK(cudaStreamSynchronize(s));
  return 0;
}

static int tc_set_direction(TcWeights* w, int N, int K, int n_tiles, int P_in, int P_out) {
  if (N != 16 && N != 48 && N != 64 && N != 128 && N != 256) { set_error("tensor-core path needs 64/128/256 output channels per pixel"); return DGAN_ERR_UNSUPPORTED; }
  if (K % 64 != 0) { set_error("tensor-core path needs input channels in multiples of 64"); return DGAN_ERR_UNSUPPORTED; }
  if (n_tiles > 32 || P_in > 65535 || P_out > 65535) { set_error("tensor-core schedule limits exceeded"); return DGAN_ERR_UNSUPPORTED; }
  w->N = N; w->K = K; w->P_in = P_in; w->P_out = P_out; w->n_tiles = n_tiles;
  return 0;
}

// ---- the generator's last layer as block GEMMs ---------------------------------------------
// Image pixels are grouped into 4x4 blocks; block (by,bx) receives from the 4x4 input pixels
// o = 2by-1+ry, p = 2bx-1+rx (ry,rx in 0..3): out row 4by+li = 2o+ka-1  =>  ka = li - 2ry + 3.
// Weight tile (ry,rx), forward:  rows (li*4+lj)*C_out+co, cols ci   = F[ka][kb][co][ci] or 0
//                      backward: rows ci, cols (li*4+lj)*C_out+co (zero padded to 64)
__global__ void tc_final_tiles_kernel(const float* __restrict__ F /*[25][C_out][C_in]*/, int C_out, int C_in,
                                      __half* __restrict__ wf /*[16][16*C_out][C_in]*/,
                                      __half* __restrict__ wb /*[16][C_in][64]*/) {
  const int tile = blockIdx.x, ry = tile >> 2, rx = tile & 3;
  const int nrow = 16 * C_out;
  for (int e = threadIdx.x; e < nrow * C_in; e += blockDim.x) {
    const int r = e / C_in, ci = e % C_in;
    const int co = r % C_out, l = r / C_out, li = l >> 2, lj = l & 3;
    const int ka = li - 2 * ry + 3, kb = lj - 2 * rx + 3;
    float v = 0.f;
    if (ka >= 0 && ka < 5 && kb >= 0 && kb < 5) v = F[((size_t)(ka * 5 + kb) * C_out + co) * C_in + ci];
    wf[((size_t)tile * nrow + r) * C_in + ci] = __float2half_rn(v);
  }
  for (int e = threadIdx.x; e < C_in * 64; e += blockDim.x) {
    const int ci = e / 64, k = e % 64;
    float v = 0.f;
    if (k < nrow) {
      const int co = k % C_out, l = k / C_out, li = l >> 2, lj = l & 3;
      const int ka = li - 2 * ry + 3, kb = lj - 2 * rx + 3;
      if (ka >= 0 && ka < 5 && kb >= 0 && kb < 5) v = F[((size_t)(ka * 5 + kb) * C_out + co) * C_in + ci];
    }
    wb[((size_t)tile * C_in + ci) * 64 + k] = __float2half_rn(v);
  }
}

// Pair table of the tensor-core path: an output pixel without contributions (use_bn on MNIST: Generator.3's backward
// over the 8x8 raster of which 7x7 is used - the cropped outputs get no gradient) receives (pixel 0, zero tile).
static PairTable tc_with_zero_tile(const PairTable& t, int zero_tile) {
  PairTable r;
  r.off.push_back(0);
  for (size_t q = 0; q + 1 < t.off.size(); ++q) {
    for (int e = t.off[q]; e < t.off[q + 1]; ++e) r.pairs.push_back(t.pairs[(size_t)e]);
    if (t.off[q] == t.off[q + 1]) r.pairs.push_back(make_int2(0, zero_tile));
    r.off.push_back((int)r.pairs.size());
  }
  return r;
}

// dz = sum over the Linear's 16 output pixels, as TC_LINEAR_SPLIT partial sums
static PairTable linear_split_pairs(int n_pix) {
  PairTable split;
  split.off.push_back(0);
  const int per = n_pix / TC_LINEAR_SPLIT;
  for (int part = 0; part < TC_LINEAR_SPLIT; ++part) {
    for (int q = part * per; q < (part + 1) * per; ++q) split.pairs.push_back(make_int2(q, q));
    split.off.push_back((int)split.pairs.size());
  }
  return split;
}

static PairTable final_block_fwd_pairs(int h_in, int w_in) {
  PairTable t;
  t.off.push_back(0);
  const int nby = h_in / 2, nbx = w_in / 2;   // (2*h_in)/4 blocks per side
  for (int by = 0; by < nby; ++by)
    for (int bx = 0; bx < nbx; ++bx) {
      for (int ry = 0; ry < 4; ++ry)
        for (int rx = 0; rx < 4; ++rx) {
          const int o = 2 * by - 1 + ry, p = 2 * bx - 1 + rx;
          if (o < 0 || o >= h_in || p < 0 || p >= w_in) continue;
          t.pairs.push_back(make_int2(o * w_in + p, ry * 4 + rx));
        }
      t.off.push_back((int)t.pairs.size());
    }
  return t;
}

static PairTable final_block_bwd_pairs(int h_in, int w_in) {
  PairTable t;
  t.off.push_back(0);
  const int nby = h_in / 2, nbx = w_in / 2;
  for (int o = 0; o < h_in; ++o)
    for (int p = 0; p < w_in; ++p) {
      for (int by = 0; by < nby; ++by) {
        const int ry = o - 2 * by + 1;
        if (ry < 0 || ry > 3) continue;
        for (int bx = 0; bx < nbx; ++bx) {
          const int rx = p - 2 * bx + 1;
          if (rx < 0 || rx > 3) continue;
          t.pairs.push_back(make_int2(by * nbx + bx, ry * 4 + rx));
        }
      }
      t.off.push_back((int)t.pairs.size());
    }
  return t;
}

struct TcFinal {
  TcWeights f, b;
  int n_blocks = 0, nbx = 0, w_out = 0, C_out = 0, act = 0;
};

static int tc_build_final(TcState& st, TcFinal* tf, const float* F, int h_in, int w_in, int C_in, int C_out, int act,
                          std::vector<void*>* allocs, cudaStream_t s) {
  if (C_in != 64) { set_error("tensor-core final layer needs net_dim == 64"); return DGAN_ERR_UNSUPPORTED; }
  if ((h_in % 2) || (w_in % 2) || h_in != w_in) { set_error("final layer geometry unsupported"); return DGAN_ERR_UNSUPPORTED; }
  tf->C_out = C_out; tf->act = act; tf->nbx = w_in / 2; tf->n_blocks = (h_in / 2) * (w_in / 2); tf->w_out = 2 * w_in;
  const int nrow = 16 * C_out;
  DGAN_CUDA_CHECK(cudaMalloc((void**)&tf->f.w, (size_t)16 * nrow * C_in * 2)); allocs->push_back(tf->f.w);
  DGAN_CUDA_CHECK(cudaMalloc((void**)&tf->b.w, (size_t)16 * C_in * 64 * 2)); allocs->push_back(tf->b.w);
  tc_final_tiles_kernel<<<16, 256, 0, s>>>(F, C_out, C_in, tf->f.w, tf->b.w);
  DGAN_CUDA_CHECK(cudaGetLastError());
  int rc;
  (void)st;
  if ((rc = tc_set_direction(&tf->f, nrow, C_in, 16, h_in * w_in, tf->n_blocks))) return rc;
  if ((rc = tc_set_direction(&tf->b, C_in, 64, 16, tf->n_blocks, h_in * w_in))) return rc;
  return 0;
}

static int tc_build(TcState& st, std::vector<TcLayerSpec>& specs, int latent, std::vector<void*>* allocs, cudaStream_t s) {
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  DGAN_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not found in the driver"); return DGAN_ERR_CUDA; }
  st.encode_fn = fn;
  int dev = 0;
  DGAN_CUDA_CHECK(cudaGetDevice(&dev));
  DGAN_CUDA_CHECK(cudaDeviceGetAttribute(&st.num_sms, cudaDevAttrMultiProcessorCount, dev));
  int rc;
  for (TcLayerSpec& sp : specs) {
    const bool linear = sp.linear_W != nullptr;
    const int n_tiles = linear ? sp.P_out : kTaps;
    const size_t elems = (size_t)n_tiles * sp.C_out * sp.C_in;
    // one extra, all-zero tile per direction: tc_with_zero_tile() gives output pixels that receive nothing a single
    // (pixel 0, zero tile) contribution, so their accumulators hold exact zeros and every item has a step
    const size_t tile_elems = (size_t)sp.C_out * sp.C_in;
    __half *wf = nullptr, *wb = nullptr;
    DGAN_CUDA_CHECK(cudaMalloc((void**)&wf, (elems + tile_elems) * 2)); allocs->push_back(wf);
    DGAN_CUDA_CHECK(cudaMalloc((void**)&wb, (elems + tile_elems) * 2)); allocs->push_back(wb);
    DGAN_CUDA_CHECK(cudaMemsetAsync(wf + elems, 0, tile_elems * 2, s));
    DGAN_CUDA_CHECK(cudaMemsetAsync(wb + elems, 0, tile_elems * 2, s));
    const unsigned blocks = (unsigned)((elems + 255) / 256);
    if (linear) {
      // forward tile q: rows c (N = C_out), cols k (K = latent) = Wt[q*C + c][k]
      tc_convert_kernel<<<blocks, 256, 0, s>>>(sp.linear_Wt, wf, elems);
      // backward tile q: rows k (N = latent), cols c (K = C_out) = W[k][q*C + c]
      tc_linear_bwd_tiles_kernel<<<blocks, 256, 0, s>>>(sp.linear_W, wb, latent, sp.C_out, sp.P_out);
    } else {
      tc_convert_kernel<<<blocks, 256, 0, s>>>(sp.w_fwd_kmajor_src, wf, elems);   // F[t][co][ci]
      tc_convert_kernel<<<blocks, 256, 0, s>>>(sp.w_bwd_kmajor_src, wb, elems);   // Ff[t][ci][co]
    }
    DGAN_CUDA_CHECK(cudaGetLastError());
    sp.out_f->w = wf; sp.out_b->w = wb;
    sp.out_f->bias_pstride = sp.bias_pstride;
    // forward: N = C_out, K = C_in;  backward: N = C_in, K = C_out (the Linear's dz is summed over its 16 pixels as
    // TC_LINEAR_SPLIT partial outputs: more items, and the z update adds the partials in a fixed order)
    if ((rc = tc_set_direction(sp.out_f, sp.C_out, sp.C_in, n_tiles + 1, sp.P_in, sp.P_out))) return rc;
    if ((rc = tc_set_direction(sp.out_b, sp.C_in, sp.C_out, n_tiles + 1, sp.P_out, linear ? TC_LINEAR_SPLIT : sp.P_in))) return rc;
  }
  (void)latent;
  return 0;
}

}  // namespace dgan
