/* nsvgd.h — C ABI of the B200-native SVGD particle-interaction engine (libnsvgd.so).
 *
 * Drop-in boundary for the SVGD hot path of ratschlab/projects-2020-Neural-SVGD
 * (reference paths below are relative to that repository).  The reference is pure
 * Python/JAX; what a `jax.ffi` custom call (or the ctypes/torch host mirror in
 * projects-2020-neural-svgd_b200/nsvgd/) binds for this path is exactly the entry
 * points declared here: plain pointers and sizes, no torch / XLA types.
 *
 * Conventions
 *   - All matrices are row-major fp32 DEVICE pointers with an explicit leading
 *     dimension (ld*, in elements).  X (n x d) particles, S (n x d) scores
 *     grad log p(x_j), Q (m x d) query points (Q = X in an SVGD step).
 *   - `stream` is a cudaStream_t passed as void*.  Every call is ENQUEUE-ONLY:
 *     no host synchronisation, no allocation; scratch comes from the caller
 *     (`ws`, size from nsvgd_workspace_bytes).  Safe to capture in a CUDA graph.
 *   - The bandwidth h lives in device memory (h_dev, 1 float) so the median
 *     heuristic can feed phi* without a host round trip.
 *   - Return value: 0 = NSVGD_OK, otherwise an nsvgd_status; the message is
 *     available from nsvgd_last_error() (thread-local).  Nothing throws/aborts.
 *   - NaN in -> NaN out (h = NaN if any particle is NaN, like jnp.median).
 *   - There is NO CPU fallback: without a CUDA device every compute entry point
 *     returns NSVGD_ERR_CUDA.
 */
#ifndef NSVGD_H_
#define NSVGD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  NSVGD_OK = 0,
  NSVGD_ERR_INVALID = 1,     /* bad shape / null pointer / bad enum (reference: ValueError, kernels.py:13-28, stein.py:29-31) */
  NSVGD_ERR_CUDA = 2,        /* CUDA runtime / driver error, or no device */
  NSVGD_ERR_UNSUPPORTED = 3, /* e.g. forced tensor-core path with d > 128 (reference: NotImplementedError) */
  NSVGD_ERR_WORKSPACE = 4    /* ws too small */
} nsvgd_status;

/* kernels.median_heuristic semantics (kernels.py:160-172); see SURVEY.md App. B */
typedef enum {
  NSVGD_MEDIAN_LITERAL = 0,  /* what kernels.py:169 executes: remove_diagonal on the PARTICLES + clamped gather */
  NSVGD_MEDIAN_OFFDIAG = 1,  /* docstring intent: median of the n(n-1) off-diagonal squared distances */
  NSVGD_MEDIAN_FULL = 2      /* full n x n matrix including the zero diagonal (original SVGD code) */
} nsvgd_median_mode;

typedef enum {
  NSVGD_PATH_AUTO = 0,
  NSVGD_PATH_SIMT = 1,       /* small-d direct-difference FP32/MUFU kernel (d <= 32) */
  NSVGD_PATH_TC = 2,         /* fused tcgen05/TMEM kernel, bf16 hi/lo split, 3 MMA passes (d <= 128) */
  NSVGD_PATH_GENERIC = 3,    /* any d: K row-chunks materialised in ws + SIMT GEMM */
  NSVGD_PATH_LARGE = 4       /* d > 128: cuBLAS bf16 tensor-op GEMMs (2-piece split, 3 passes) around a fused pair kernel */
} nsvgd_path;

typedef enum {
  NSVGD_OP_MEDIAN = 0,
  NSVGD_OP_PHI = 1,
  NSVGD_OP_KSD = 2,
  NSVGD_OP_MMD = 3,
  NSVGD_OP_STEP = 4
} nsvgd_op;

typedef enum {
  NSVGD_KSD_RECT = 0,        /* stein.ksd_squared   (stein.py:273-295): mean over n x m */
  NSVGD_KSD_U = 1,           /* stein.ksd_squared_u (stein.py:299-323): sum_{i!=j} / (n(n-1)) */
  NSVGD_KSD_V = 2            /* stein.ksd_squared_v (stein.py:327-349): sum / n^2 */
} nsvgd_ksd_variant;

/* ---- introspection ------------------------------------------------------- */
int nsvgd_version(void);
const char* nsvgd_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
uint64_t nsvgd_launch_count(void);

/* Optional device-side timing (bench.py's roofline leg): when enabled, the library brackets its
 * dominant kernels with CUDA events on the launching stream.  tag 0 = phi* main kernel,
 * 1 = median histogram passes, 2 = whole phi* op, 3 = whole median op.  read() synchronises on the
 * recorded events, returns the summed milliseconds and launch count, and clears the tag. */
int nsvgd_profile_enable(int on);
int nsvgd_profile_read(int tag, double* total_ms, int* count);

/* Scratch size for one call of `op` (upper bound over paths unless `path` forces one). */
int nsvgd_workspace_bytes(int op, int64_t n, int64_t m, int64_t d, int path, size_t* out_bytes);

/* ---- A2: median-heuristic bandwidth  (kernels.py:160-172, utils.py:178-197, 653-658)
 * h = sqrt(0.5 * median(D) / log(n+1)) in fp32, exact order statistics of the
 * defined-order fp32 pair distances (bit-exact vs oracle/svgd_oracle.py). */
int nsvgd_median_h_f32(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx,
                       int mode, float* h_dev, void* ws, size_t ws_bytes);

/* The same with a warm start for SVGD loops (kernels.py:160-172 called every step, models.py:657): hint_dev = 32 bytes of
 * caller-owned, zero-initialised device memory kept between calls.  On the tensor-core filtered path (d <= 128, n >= 8192) the
 * bracket then comes from the previous call's order statistics (margin = 4 x the last relative change, 2^-13 .. 2^-7) and the
 * sampling + bracket-select phases are skipped; exact counts verify the bracket, a miss runs the exact fallback and drops the
 * hint.  Bit-identical result with or without it.  nsvgd_p2p_median_fast_run_hinted: the row-sharded form. */
int nsvgd_median_h_hinted_f32(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, int mode, float* h_dev,
                              void* hint_dev, void* ws, size_t ws_bytes);

/* Building blocks of the same select for row-sharded multi-GPU runs: every rank
 * histograms the tile rows  tile_row_start, +tile_row_stride, ...  of the upper
 * triangle, the host all-reduces `hist` (NCCL sum over uint64[2*2048]) between
 * passes, and advance/finish then take identical decisions on every rank. */
#define NSVGD_SELECT_BINS 2048
#define NSVGD_SELECT_PASSES 3
#define NSVGD_SELECT_STATE_BYTES 256
int nsvgd_median_prepare(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, int mode,
                         float* Z /* n x dz out (literal mode) or NULL */, int64_t* dz_out, float* w_last_out,
                         void* state_dev);
int nsvgd_median_hist_pass(void* stream, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last,
                           int64_t tile_row_start, int64_t tile_row_stride,
                           const void* state_dev, int pass, uint64_t* hist_dev /* 2*NSVGD_SELECT_BINS, zeroed by the call */);
int nsvgd_median_advance(void* stream, void* state_dev, const uint64_t* hist_dev, int pass);
int nsvgd_median_finish(void* stream, const void* state_dev, int64_t n, float* h_dev);

/* Tensor-core filtered variant of the same exact select (n >= 8192, 32 < dz <= 128),
 * split into stages so that a row-sharded run can all-reduce between them:
 *   sample  : centred bf16 pieces + this rank's slice (part of nparts) of the 2^22 sampled pair distances;
 *             returns where the sample buffer lives inside ws  [host: all-gather the slices in place]
 *   stage1  : bracket from the sampled distances (identical on every rank), tcgen05 filter over the 128-row blocks
 *             block_row_start, +block_row_stride, ..., exact re-evaluation of the uncertain pairs;
 *             reduce_dev (int64[4]) = {pairs below the bracket, candidates, overflow flag, listed pairs}
 *   [host: all-reduce(sum) reduce_dev]
 *   resolve : ranks inside the candidates, or raise the device-side fallback flag
 *   3 x { cand_hist -> [all-reduce(sum) hist] -> cand_advance },  commit
 *   3 x { fallback_hist -> [all-reduce] -> fallback_advance }   (kernels return at once unless flagged)
 *   nsvgd_median_finish.
 * `ws` is a scratch region of nsvgd_median_fast_workspace_bytes() that must stay untouched between stages. */
int nsvgd_median_fast_supported(int64_t n, int64_t dz);
/* d > 128 (csrc/median_large.cu): nsvgd_median_h_f32 runs the tensor-core filtered exact select on ONE GPU (symmetric CTA-pair
 * Gram with a classify epilogue, per-tile lists, tile-wise defined-order evaluation; 9 ms at n = 16384, d = 4096 vs 191 ms for
 * the 3-pass SIMT select).  Row-sharded callers run it redundantly on the gathered array: the result is an exact order
 * statistic, hence identical on every rank. */
int nsvgd_median_large_supported(int64_t n, int64_t dz);
int nsvgd_median_fast_workspace_bytes(int64_t n, int64_t dz, size_t* out_bytes);
int nsvgd_median_fast_sample(void* stream, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last,
                             int64_t part, int64_t nparts, void* ws, size_t ws_bytes, size_t* samples_offset_bytes,
                             int64_t* samples_count);
int nsvgd_median_fast_stage1(void* stream, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last, int mode,
                             int64_t block_row_start, int64_t block_row_stride, void* state_dev, uint64_t* hist_dev,
                             void* ws, size_t ws_bytes, int64_t* reduce_dev);
int nsvgd_median_fast_resolve(void* stream, void* state_dev, void* ws, size_t ws_bytes, int64_t n, int64_t dz,
                              const int64_t* reduce_dev);
int nsvgd_median_fast_cand_hist(void* stream, void* ws, size_t ws_bytes, int64_t n, int64_t dz, int pass, uint64_t* hist_dev);
int nsvgd_median_fast_cand_advance(void* stream, void* ws, size_t ws_bytes, int64_t n, int64_t dz, const uint64_t* hist_dev, int pass);
int nsvgd_median_fast_commit(void* stream, void* state_dev, void* ws, size_t ws_bytes, int64_t n, int64_t dz);
int nsvgd_median_fallback_hist(void* stream, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last,
                               int64_t tile_row_start, int64_t tile_row_stride, const void* state_dev, void* ws,
                               size_t ws_bytes, int pass, uint64_t* hist_dev);
int nsvgd_median_fallback_advance(void* stream, void* state_dev, void* ws, size_t ws_bytes, int64_t n, int64_t dz,
                                  const uint64_t* hist_dev, int pass);

/* ---- row-sharded runs — small all-reduces over NVLink PEER MEMORY, fused into their consumer kernels.
 * Between the heavy kernels a row-sharded step exchanges only tiny vectors (radix-select histograms, candidate counts,
 * the KSD/l2 sums).  Instead of one NCCL all-reduce (launch + ~20 us) each, the consumer kernel publishes its vector in a
 * peer-mapped block, flags every peer (st.release.sys), waits for theirs (ld.acquire.sys) and sums the peers' vectors in
 * rank order — integer sums and fixed-order double sums are bit-identical on every rank (csrc/p2p.cuh).
 * Plumbing: each rank allocates ONE block (nsvgd_p2p_alloc: cudaMalloc + zero + cudaIpcGetMemHandle, 64-byte handle), the
 * host layer exchanges the handles (any transport; nsvgd/parallel.py uses one torch.distributed all-gather), opens the
 * peers' blocks (nsvgd_p2p_open) and passes `blocks[world]` (own block at index `rank`) to the calls below.  All ranks
 * must issue the same sequence of p2p calls.  Enqueue-only, CUDA-graph capturable (epochs live in device memory); a peer
 * that does not arrive within ~2 s raises the block's timeout flag instead of hanging (nsvgd_p2p_timeout_flag). */
size_t nsvgd_p2p_block_bytes(void);
int nsvgd_p2p_alloc(void** dev_ptr, unsigned char* handle_out /* 64 bytes */);
int nsvgd_p2p_open(const unsigned char* handle /* 64 bytes */, void** peer_ptr);
int nsvgd_p2p_close(void* peer_ptr);
int nsvgd_p2p_free(void* dev_ptr);
int nsvgd_p2p_timeout_flag(const void* own_block, int* flag_out);
/* stand-alone form: in-place sum of buf[0..nwords) (uint64 adds, or doubles in rank order), nwords <= 8192 */
int nsvgd_p2p_allreduce(void* stream, int rank, int world, void* const* blocks, void* buf, int64_t nwords, int is_f64);
/* = all-reduce(hist) + nsvgd_median_advance */
int nsvgd_p2p_median_advance(void* stream, int rank, int world, void* const* blocks, void* state_dev, uint64_t* hist_dev, int pass);
/* The whole tensor-core filtered select for one rank of a row-sharded group in ONE call: sample slice `rank` of `world`,
 * fused 3-pass bracket select (histograms summed across ranks inside the kernel), filter over the 128-row blocks
 * rank, rank + world, ..., exact evaluation, pack + resolve, fused candidate select, conditional SIMT fallback.  Leaves the
 * resolved order statistics in state_dev (then nsvgd_median_finish).  world = 1: the single-GPU path. */
int nsvgd_p2p_median_fast_run(void* stream, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last, int mode,
                              int rank, int world, void* const* blocks, void* state_dev, uint64_t* hist_dev, void* ws,
                              size_t ws_bytes);
int nsvgd_p2p_median_fast_run_hinted(void* stream, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last, int mode,
                                     int rank, int world, void* const* blocks, void* state_dev, uint64_t* hist_dev,
                                     void* hint_dev, void* ws, size_t ws_bytes);
/* = all-reduce(reduce_dev[4]) + nsvgd_median_fast_resolve */
int nsvgd_p2p_median_fast_resolve(void* stream, int rank, int world, void* const* blocks, void* state_dev, void* ws,
                                  size_t ws_bytes, int64_t n, int64_t dz, int64_t* reduce_dev);
/* hist_dev = uint64[4*2048]: [candidate histogram | fallback histogram] of radix pass `pass`;
 * = all-reduce(hist) + nsvgd_median_fast_cand_advance + (iff the fallback flag is set) nsvgd_median_fallback_advance */
int nsvgd_p2p_median_fast_advance2(void* stream, int rank, int world, void* const* blocks, void* state_dev, void* ws,
                                   size_t ws_bytes, int64_t n, int64_t dz, uint64_t* hist_dev, int pass);
/* = all-reduce(stats_dev[4], doubles) + nsvgd_apply_update_f32 */
int nsvgd_p2p_apply_update_f32(void* stream, int rank, int world, void* const* blocks, float* Xrows, const float* phi,
                               int64_t m, int64_t d, int64_t ldx, int64_t ldphi, double* stats_dev, float lr, int scaled,
                               float lambda_reg, float* grads, double* step_aux_dev);

/* ---- A1/A3/A4: phi*  (stein.py:71-96, 180-222; kernels.py:52-65)
 * phi[i,:] = (1/n) sum_j k_ij s_j + k_ij (q_i - x_j)/h^2,  k_ij = exp(-|q_i-x_j|^2/(2h^2)).
 * Optional outputs (NULL to skip):
 *   aux   (m x 2 x d): [drift, repulsion] split of stein.phistar (stein.py:206-222)
 *   stats (double[4], device): {sum_i (s_i.phi_i + div phi(q_i)), sum_i |phi_i|^2, m, 0}
 *          — needs Sq (scores at the queries, = S when Q = X); feeds
 *          stein.stein_discrepancy(f=phi*) (stein.py:99-106) and utils.l2_norm_squared (utils.py:483-487).
 */
int nsvgd_phi_f32(void* stream,
                  const float* Q, const float* Sq, int64_t m, int64_t ldq,
                  const float* X, const float* S, int64_t n, int64_t ldx,
                  int64_t d, const float* h_dev,
                  float* phi, int64_t ldphi, float* aux, double* stats_dev,
                  int path, void* ws, size_t ws_bytes);

/* ---- A5/A7/A8: one SVGD step, sgd particle optimizer  (models.py:143-167, 654-689)
 * grads = -phi (scaled == 0) or -alpha*phi with alpha = ksd/(2*lambda_reg*l2);
 * X <- X - lr*grads.  h_dev must already hold the bandwidth (fixed, or from
 * nsvgd_median_h_f32).  Rows [row_begin,row_end) of X are the queries this call
 * updates (whole range on one GPU; the local block when row-sharded — then
 * pass stats_allreduced = 0 and use the two-phase entry points below).
 * step_aux_dev (double[8]): {ksd, l2, alpha, sum grads^2, sum updates^2, 0,0,0}. */
int nsvgd_step_sgd_f32(void* stream, float* X, const float* S, int64_t n, int64_t d, int64_t ldx,
                       const float* h_dev, float lr, int scaled, float lambda_reg,
                       float* grads /* n x d out, may be NULL */, double* step_aux_dev,
                       int path, void* ws, size_t ws_bytes);

/* two-phase form for row-sharded runs: (1) nsvgd_phi_f32 on the local rows with
 * stats, (2) host all-reduces stats (double[4]), (3) apply: */
int nsvgd_apply_update_f32(void* stream, float* Xrows, const float* phi, int64_t m, int64_t d, int64_t ldx, int64_t ldphi,
                           const double* stats_dev /* global sums, or NULL when !scaled */,
                           float lr, int scaled, float lambda_reg,
                           float* grads, double* step_aux_dev);

/* ---- A6: KSD^2  (stein.py:273-349), closed form
 * u_p(x,y) = k [ s_x.s_y + (s_x-s_y).(x-y)/h^2 + d/h^2 - r^2/h^4 ].  out_dev: 1 double. */
int nsvgd_ksd_f32(void* stream, const float* X, const float* Sx, int64_t n, int64_t ldx,
                  const float* Y, const float* Sy, int64_t m, int64_t ldy,
                  int64_t d, const float* h_dev, int variant, double* out_dev, void* ws, size_t ws_bytes);

/* The U and V statistics through the fused phi* kernels (tensor cores for d > 32): stein_discrepancy(phi*) (stein.py:99-106)
 * is the V-statistic (stein.py:327-349) and U = (n^2 V - sum_i(|s_i|^2 + d/h^2)) / (n(n-1)) (stein.py:299-323).  One phi* pass
 * instead of n^2/2 fp32 SIMT pair evaluations: the route `stein.ksd_squared_u / _v` take for n >= 4096. */
int nsvgd_ksd_tc_workspace_bytes(int64_t n, int64_t d, size_t* out_bytes);
int nsvgd_ksd_tc_f32(void* stream, const float* X, const float* S, int64_t n, int64_t d, int64_t ldx, const float* h_dev,
                     int variant, double* out_dev, void* ws, size_t ws_bytes);

/* ---- A9: MMD^2  (metrics.py:88-127; funnel != 0 applies kernels.defunnelize, kernels.py:112-123)
 * out_dev: 1 double = mean offdiag Kxx + mean offdiag Kyy - 2 mean Kxy. */
int nsvgd_mmd_f32(void* stream, const float* X, int64_t n, int64_t ldx,
                  const float* Y, int64_t m, int64_t ldy, int64_t d,
                  float h, int funnel, double* out_dev, void* ws, size_t ws_bytes);

/* sum of the RBF kernel matrix between two point sets (building block of MMD; symmetric != 0
 * requires A == B and returns sum_{i != j} = 2 sum_{i<j}, the diagonal excluded).  out_dev: 1 double. */
int nsvgd_ksum_f32(void* stream, const float* A, int64_t na, int64_t lda,
                   const float* B, int64_t nb, int64_t ldb, int64_t d,
                   float h, int symmetric, double* out_dev, void* ws, size_t ws_bytes);

/* adam particle optimizer used by models.Particles(optimizer="adam") (utils.py:711-715):
 * standard optax.adam(lr) on `grads`; m1/m2 are caller-owned state, t the 1-based step. */
int nsvgd_adam_update_f32(void* stream, float* X, const float* grads, float* m1, float* m2,
                          int64_t count, float lr, float b1, float b2, float eps, int64_t t);

/* ---- K complete SVGD steps in ONE launch for latency-bound sizes: one persistent CTA runs, per step,
 * log -> [tracers every `cadence`-th step] -> device scores -> exact median bandwidth -> phi* + KSD/l2 -> scaled sgd update,
 * i.e. models.Particles.step + log (models.py:143-197) for KernelGradient with optax.sgd, tracer cadence on the device.
 * Limits: nsvgd_loop_small_supported(n, d) (n <= 1024, d <= 32, n * d <= 8192); target 0 = funnel, 1 = gaussian with
 * tparams = [mean (d) | precision (d*d, or its diagonal if tparams_diag)].  h_fixed > 0 skips the median heuristic.
 * Logs (each may be NULL): particles [K][n][d], mean / std [K][d], bandwidth [K], logp [K][n], grad norms [K],
 * stats [K][3] = {ksd, l2, alpha}.  Tracer (Y != NULL): tracer_kind 1 = RBF MMD, 2 = RBF + funnel-kernel MMD
 * (metrics.py:112-127) into mmd_out [.][2], mmd_step [.], *mmd_count (zeroed by the caller). */
int nsvgd_loop_small_supported(int64_t n, int64_t d);
int nsvgd_svgd_loop_small_f32(void* stream, float* X, int64_t n, int64_t d, int target, const float* tparams,
                              int tparams_diag, float log_norm, int median_mode, float h_fixed, float lr, int scaled,
                              float lambda_reg, int64_t K, int64_t step0, float* log_particles, float* log_mean,
                              float* log_std, float* log_bandwidth, float* log_logp, float* log_gnorm,
                              float* log_gnorm_post, double* log_stats, const float* Y, int64_t m, int cadence,
                              int tracer_kind, float tracer_h, float* mmd_out, int64_t* mmd_step, int* mmd_count);

/* ---- vector (ARD) and full-matrix Gaussian kernels (kernels.py:52-65 with a (d,) bandwidth, kernels.py:83-88;
 * nets.RBFKernel nets.py:24-49).  phi* and MMD run on the unit-bandwidth engine after a linear map of the
 * particles (x' = x / h or x' = x Sigma^T; scores s' = s h or s Sigma^-1; phi = phi' / h or phi' Sigma): */
int nsvgd_scale_cols_f32(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, const float* v, int invert,
                         float* out, int64_t ldo);
int nsvgd_matmul_f32(void* stream, const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, float* C,
                     int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate);
/* the KSD is not invariant under that map; closed form with per-coordinate weights inv_h2[a] = 1/h_a^2 (device, d floats) */
int nsvgd_ksd_ard_f32(void* stream, const float* X, const float* Sx, int64_t n, int64_t ldx, const float* Y,
                      const float* Sy, int64_t m, int64_t ldy, int64_t d, const float* inv_h2, float inv_h2_sum,
                      int variant, double* out_dev, void* ws, size_t ws_bytes);

/* linear-time KSD estimator (stein.py:353-380): mean_i u_p(x_i, x_{i + n/2}) over the two halves of the sample (an odd
 * count drops the last one).  out_dev: 2 doubles {mean, std(ddof = 1) / (n/2)}; ws: >= (n/2) floats. */
int nsvgd_ksd_linear_f32(void* stream, const float* X, const float* S, int64_t n, int64_t d, int64_t ld,
                         const float* h_dev, double* out_dev, void* ws, size_t ws_bytes);

/* ---- differentiable operators: backward rules -------------------------------------------------------
 * The reference differentiates through its kernels and phi* with jax.grad / jacfwd (stein.py:99-106, 180-203; the
 * kernel-learning notebooks differentiate the KSD w.r.t. the bandwidth).  These entry points are the closed-form
 * backward rules a jax.custom_vjp (or torch.autograd.Function, nsvgd/autodiff.py) binds. */
/* KSD (variant as nsvgd_ksd_f32) and d KSD / d h in one pass; out_dev, dh_dev: 1 double each */
int nsvgd_ksd_dh_f32(void* stream, const float* X, const float* Sx, int64_t n, int64_t ldx, const float* Y,
                     const float* Sy, int64_t m, int64_t ldy, int64_t d, const float* h_dev, int variant,
                     double* out_dev, double* dh_dev, void* ws, size_t ws_bytes);
/* cotangents of phi = phistar(Q; X, S, h) for an upstream cotangent G (m x d, leading dimension ldq like Q):
 * dQ (m x d), dX (n x d), dS (n x d) (each may be NULL), dh_dev (1 double, may be NULL).  Q == X (the SVGD case): the
 * caller adds dQ and dX.  Scratch: nsvgd_phi_vjp_workspace_bytes. */
int nsvgd_phi_vjp_workspace_bytes(int64_t m, int64_t n, int64_t d, size_t* out_bytes);
int nsvgd_phi_vjp_f32(void* stream, const float* Q, const float* G, int64_t m, int64_t ldq, const float* X, const float* S,
                      int64_t n, int64_t ldx, int64_t d, const float* h_dev, float* dQ, int64_t lddq, float* dX, int64_t lddx,
                      float* dS, int64_t ldds, double* dh_dev, void* ws, size_t ws_bytes);

/* ---- grad log p (and optionally log p) of the reference's canonical targets, on the device ----------
 * Replace vmap(grad(logp)) of stein.py:86-89 / models.py:662-675 for targets whose score has a closed form, so that a
 * whole SVGD step is enqueued without the framework's autodiff.  S: n x d scores; logp: n values or NULL. */
/* Neal's funnel (distributions.py:275-319): last coordinate y ~ N(0, 9), x_k | y ~ N(0, e^y) */
int nsvgd_score_funnel_f32(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, float* S, int64_t lds,
                           float* logp);
/* N(mean, cov) (distributions.py:89-154, 627-639): prec = cov^-1 row-major d x d, or its diagonal (prec_is_diag);
 * log_norm = -(d/2) log 2 pi - (1/2) log det cov */
int nsvgd_score_gaussian_f32(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, const float* mean,
                             const float* prec, int prec_is_diag, float log_norm, float* S, int64_t lds, float* logp);
/* Bayesian logistic regression posterior (experiments/bayesian_logistic_regression.py:73-114, 169-190):
 * params P[i] = [w (F), log alpha], d = F + 1; minibatch Xb (B x F, row-major), yb (B) in {0,1};
 * log p = prior_logp(w, log alpha) + num_datapoints * SUM_batch log sigmoid(y' x.w)   (the reference's literal scaling) */
int nsvgd_score_logreg_f32(void* stream, const float* P, int64_t n, int64_t d, int64_t ldp, const float* Xb, const float* yb,
                           int64_t B, int64_t F, float num_datapoints, float a0, float b0, float* S, int64_t lds,
                           float* logp);

#ifdef __cplusplus
}
#endif
#endif /* NSVGD_H_ */
