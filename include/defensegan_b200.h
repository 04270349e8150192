/*
 * defensegan_b200.h - C ABI of the B200-native Defense-GAN projection loop.
 *
 * The reference (kabkabm/defensegan) has no FFI/plugin interface: the boundary is the Python
 * method DefenseGANBase.reconstruct (models/gan.py:333-449) plus the eval driver
 * utils/gan_defense.py:32-179.  This header is the C-ABI that sits directly underneath that
 * Python surface; each entry point cites the reference lines it replaces.
 * Plain pointers and sizes only - no torch types.  All device pointers are owned by the
 * caller; the handle owns its own copies of the weights (plain and re-laid-out), so the tensors passed
 * to dgan_create may be freed once `stream` has run the copies.  Nothing is freed across the boundary.  No function synchronises the host: work is enqueued on `stream`.
 *
 * Every function returns 0 on success, a negative dgan_status otherwise; the message is
 * available (thread-local) from dgan_last_error().  No C++ exception crosses the boundary.
 */
#ifndef DEFENSEGAN_B200_H_
#define DEFENSEGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGAN_ABI_VERSION 2

typedef struct dgan_ctx* dgan_handle;

enum dgan_status {
  DGAN_OK = 0,
  DGAN_ERR_INVALID_ARG = -1,
  DGAN_ERR_CUDA = -2,
  DGAN_ERR_UNSUPPORTED = -3,
  DGAN_ERR_WORKSPACE = -4
};

/* generator architectures (models/dataset_models.py:36-71 mnist_generator - also used by
 * F-MNIST, models/gan.py:688-698 - and :127-165 celeba_generator) */
enum dgan_arch { DGAN_ARCH_MNIST = 0, DGAN_ARCH_CELEBA = 1 };

/* arithmetic of the contractions.  State (z, momentum, loss, accumulators) is always fp32. */
enum dgan_precision {
  DGAN_PREC_FP32 = 0, /* fp32 operands, CUDA-core FMA: the reference's arithmetic type */
  DGAN_PREC_FP16 = 1  /* fp16 operands, fp32 accumulate, tcgen05 tensor cores */
};

typedef struct dgan_desc {
  int32_t abi_version; /* DGAN_ABI_VERSION */
  int32_t arch;        /* dgan_arch */
  int32_t latent_dim;  /* LATENT_DIM (experiments/cfgs/gans/default.yml:4) */
  int32_t net_dim;     /* NET_DIM    (default.yml:7) */
  int32_t use_bn;      /* USE_BN     (default.yml:3); batch-statistics BN, tflib/ops/batchnorm.py:80-93 (both precisions; fp16 path: fp32 pre-activations and statistics, fp16 activations) */
  int32_t precision;   /* dgan_precision */
} dgan_desc;

/* Number of weight tensors dgan_create expects for a descriptor, in the reference's variable
 * creation order (tflib/__init__.py:7-33 names):
 *   Generator.Input.W [latent,4096*] , Generator.Input.b,
 *   [BN1.offset, BN1.scale,]
 *   Generator.2.Filters [5,5,Cout,Cin], Generator.2.Biases, [BN2.offset, BN2.scale,]
 *   Generator.3.Filters, Generator.3.Biases, [BN3.offset, BN3.scale,]
 *   Generator.5.Filters, Generator.5.Biases, (celeba: Generator.6.Filters, Generator.6.Biases)
 * All fp32, device memory, TF layouts (Linear W is (in,out), tflib/ops/linear.py:129-133;
 * Deconv filters are (kh,kw,Cout,Cin), tflib/ops/deconv2d.py:67,104-110). */
int dgan_num_weights(const dgan_desc* desc);

/* Replaces DefenseGANBase.load_generator + the tflib.param registry (models/gan.py:80-87,
 * tflib/__init__.py:7-33): copies the device-resident weights into handle-owned memory and builds the
 * re-laid-out forms (per-tap tiles, transposes, fp16 copies) and the launch schedules on `stream`. */
int dgan_create(dgan_handle* out, const dgan_desc* desc, const float* const* weights_dev,
                int n_weights, void* stream);
int dgan_destroy(dgan_handle h);

/* Bytes of caller-owned scratch needed by dgan_reconstruct / dgan_forward / dgan_loss_grad
 * for `batch` images x `rec_rr` restarts.  Also plans and uploads the launch schedules for that many latent rows
 * (cached in the handle; this is where the one-time allocation and synchronisation of a batch size happen). */
size_t dgan_workspace_bytes(dgan_handle h, int batch, int rec_rr);

/* Hyper-parameters of one projection call: the attributes DefenseGANBase.reconstruct reads from the model object at
 * call time (models/gan.py:333-349: rec_rr, rec_iters, rec_lr) plus the optimiser constant of gan.py:389-391. */
typedef struct dgan_rec_params {
  int32_t batch;          /* images in x_dev */
  int32_t rec_rr;         /* R: random restarts per image (REC_RR, mnist.yml:7) */
  int32_t rec_iters;      /* L: gradient steps (REC_ITERS, mnist.yml:5) */
  float rec_lr;           /* constant: the reference's decay is dead code */
  float momentum;         /* 0.7 in the reference (gan.py:389-391) */
  int32_t decay_lr;       /* 0 = reference behaviour; 1 = the evidently intended x0.1 from step ceil(0.8 L) */
  uint64_t seed;          /* Philox key of the z0 stream when z0_dev == NULL */
  uint64_t z_row_offset;  /* index of this call's first latent row in that stream: a caller that shards one batch over
                             several GPUs passes (first image of the shard) * rec_rr so that the draw equals the
                             single-GPU draw row for row; 0 otherwise */
} dgan_rec_params;

/* Replaces one sess.run of the op built by DefenseGANBase.reconstruct (models/gan.py:333-449)
 * preceded by tf.local_variables_initializer() (utils/gan_defense.py:119):
 *   x_dev     [batch, H, W, C] fp32 NHWC, already input-transformed
 *   z0_dev    [batch*rec_rr, latent] fp32 (the reference's z_init_val, gan.py:395-397) or NULL:
 *             z0 ~ N(0, 1/latent) from the Philox stream (seed, z_row_offset) (gan.py:370-377)
 *   rec_dev   [batch, H, W, C] fp32: G(z_{L-1}) of the arg-min restart (gan.py:438-449)
 *   loss_dev  [batch] fp32 min per-image MSE, nullable;  idx_dev [batch] int32 chosen restart, nullable
 * The call enqueues the whole L-step loop on `stream` (8 kernels per L-step with DGAN_PREC_FP16) and never
 * synchronises the host; it does not allocate either once dgan_workspace_bytes has been called for this batch x rec_rr. */
int dgan_reconstruct(dgan_handle h, const dgan_rec_params* params, const float* x_dev,
                     const float* z0_dev, float* rec_dev, float* loss_dev, int32_t* idx_dev,
                     void* workspace, size_t workspace_bytes, void* stream);

/* The z_hat initialiser alone (models/gan.py:370-377): z_dev [n_rows, latent] ~ N(0, 1/latent), rows
 * [z_row_offset, z_row_offset + n_rows) of the Philox stream keyed by `seed` - exactly what dgan_reconstruct
 * draws when z0_dev == NULL. */
int dgan_sample_z0(dgan_handle h, uint64_t seed, uint64_t z_row_offset, int n_rows, float* z_dev,
                   void* stream);

/* generator_fn(z) (models/gan.py:657-665,726-735): y_dev [n_rows, H*W*C] fp32. */
int dgan_forward(dgan_handle h, const float* z_dev, int n_rows, float* y_dev, void* workspace,
                 size_t workspace_bytes, void* stream);

/* One evaluation of the loop body (models/gan.py:409-417) without the update, for known-answer
 * tests: y [batch*rec_rr, HWC], loss [batch*rec_rr], grad = d(sum loss)/dz [batch*rec_rr, latent]. */
int dgan_loss_grad(dgan_handle h, const float* x_dev, int batch, int rec_rr, const float* z_dev,
                   float* y_dev, float* loss_dev, float* grad_dev, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Kernels run by the most recent dgan_reconstruct on this handle (1 + 8 L - 4 + 2 with DGAN_PREC_FP16 on the MNIST stack). */
int64_t dgan_last_launch_count(dgan_handle h);

/* Stream operations the HOST issued for it.  The L-step loop only touches the workspace, so it is captured into a CUDA
 * graph the first time a (workspace, batch, rec_rr, rec_iters, rec_lr, momentum, decay_lr) combination is seen and replayed
 * with one cudaGraphLaunch afterwards: z0 initialiser (+ its memsets), image copy, graph, loss sum, arg-min select. */
int64_t dgan_last_enqueue_count(dgan_handle h);

/* Algorithmic multiply-accumulates of one generator forward per latent row (exact in-bounds
 * taps); backward-to-z has the same count. */
int64_t dgan_macs_per_row(dgan_handle h);

/* Per-kernel device timing for roofline reports (no reference counterpart).  While enabled every
 * kernel launch of dgan_reconstruct is bracketed by CUDA events on the launching stream; never
 * enable it in a timed throughput pass.  dgan_profile_read synchronises on the recorded events
 * and returns, per kernel kind (layer x direction), the summed milliseconds, the launch count and
 * the algorithmic FLOPs of one launch (2 x exact in-bounds MACs x latent rows of the last call). */
int dgan_profile_enable(dgan_handle h, int enable);
int dgan_profile_num_kinds(dgan_handle h);
const char* dgan_profile_kind_name(dgan_handle h, int kind);
int dgan_profile_read(dgan_handle h, int max_kinds, double* ms_out, int64_t* launches_out,
                      double* flops_per_launch_out);

const char* dgan_last_error(void);
int dgan_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DEFENSEGAN_B200_H_ */
