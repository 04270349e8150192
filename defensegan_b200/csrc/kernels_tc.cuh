// placeholder - replaced by the tcgen05 implementation
#pragma once
#include "common.cuh"
namespace dgan {
struct TcWeights { int dummy = 0; };
struct TcState { float grad_scale = 1.f; };
struct TcLayerSpec {
  int P_in, C_in, P_out, C_out, h_in, w_in, h_used, w_used;
  const PairTable *fwd, *bwd;
  const float *w_fwd_kmajor_src, *w_bwd_kmajor_src, *linear_W, *linear_Wt;
  TcWeights *out_f, *out_b;
};
inline int tc_build(TcState&, std::vector<TcLayerSpec>&, int, std::vector<void*>*, cudaStream_t) { set_error("tensor-core path not built"); return DGAN_ERR_UNSUPPORTED; }
inline int tc_launch(TcState&, int64_t*, const TcWeights&, const __half*, __half*, int, int, const float*, const __half*, float, cudaStream_t) { return DGAN_ERR_UNSUPPORTED; }
inline int tc_launch_f32out(TcState&, int64_t*, const TcWeights&, const __half*, float*, int, cudaStream_t) { return DGAN_ERR_UNSUPPORTED; }
}
