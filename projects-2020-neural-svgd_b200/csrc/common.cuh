// common.cuh — shared helpers for libnsvgd (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/nsvgd.h"

namespace nsvgd {

// ---- error plumbing: status codes + thread-local message, never throw -------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launch_count;
inline void count_launch(int k = 1) { g_launch_count.fetch_add((uint64_t)k, std::memory_order_relaxed); }

#define NSVGD_REQUIRE(cond, code, ...)        \
  do {                                        \
    if (!(cond)) {                            \
      ::nsvgd::set_error(__VA_ARGS__);        \
      return (code);                          \
    }                                         \
  } while (0)

#define NSVGD_CUDA(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::nsvgd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                         __FILE__, __LINE__);                                         \
      return NSVGD_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

#define NSVGD_LAUNCH_CHECK()                                                          \
  do {                                                                                \
    ::nsvgd::count_launch();                                                          \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      ::nsvgd::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),  \
                         __FILE__, __LINE__);                                         \
      return NSVGD_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

#define NSVGD_PROPAGATE(expr)     \
  do {                            \
    int _s = (expr);              \
    if (_s != NSVGD_OK) return _s; \
  } while (0)

// ---- workspace carving -------------------------------------------------------------------
struct WsCarver {
  char* base;
  size_t off;
  size_t cap;
  bool dry;  // dry run: only measure
  WsCarver(void* p, size_t bytes, bool dry_run = false) : base((char*)p), off(0), cap(bytes), dry(dry_run) {}
  template <typename T>
  T* take(size_t count, size_t align = 256) {
    off = (off + align - 1) / align * align;
    T* r = dry ? nullptr : (T*)(base + off);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return dry || off <= cap; }
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

int device_sm_count();
// cudaFuncAttributeMaxDynamicSharedMemorySize, once per (kernel, device): the attribute is per device, so a second GPU
// in the same process (JAX, multi-device torch) needs its own call
int ensure_dyn_smem(const void* kernel, int bytes);
void prof_begin(int tag, cudaStream_t st);
void prof_end(int tag, cudaStream_t st);

// ---- device helpers ------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of doubles; result valid in thread 0.  sh must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int nw = (blockDim.x * blockDim.y + 31) >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = lane < nw ? sh[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

// log2(e) / 2 : k = exp(-r2/(2 h^2)) = exp2(-r2 * kHalfLog2e / h^2)
#define NSVGD_HALF_LOG2E 0.72134752044448170368f
#endif

// ---- internal kernels shared between translation units (declared here, defined in the .cu) --
// rowacc layout produced by every phi path and consumed by the epilogue (update.cu):
//   DIRECT form: KS[m x d] = sum_j k s_j ; R[m x d] = sum_j k (q_i - x_j) ;
//                scal[m x 4] = {rowsum, a = sum_j k s_j.(x_j - q_i), b = sum_j k r2, 0}
//   GRAM form  : O[split][m_pad x ldo] = [K S | K Xc] partial over j-splits;
//                scal[split][m_pad x 4] = {rowsum, kc = sum_j k c_j, b, 0}; needs mu (column mean)
//   GRAMW form : O[split][m_pad x ldo] = K W partial, W = S - Xc/h^2 (one block: the second GEMM is half as wide);
//                scal[...] = {rowsum, kc, b, kg = sum_j k g_ij = qc_i.(K Xc)_i};
//                phi_i = (1/n)[(K W)_i + rowsum_i qc_i/h^2],  qc_i.(K S)_i = qc_i.(K W)_i + kg_i/h^2
//   GRAMW2 form: O[split][m_pad x ldo] = K [W | 1 | c | 0..] partial (ldo = dpad + 16): rowsum and kc are columns dpad, dpad + 1;
//                scal[parts][m_pad x 2] = {b, kg}
struct EpilogueArgs {
  int form;               // 0 = DIRECT, 1 = GRAM, 2 = GRAMW, 3 = GRAMW2 (rowsum / K c in columns dpad, dpad + 1 of O; scal = {b, kg})
  const uint32_t* err_flag;   // tensor-core paths: non-zero when a kernel aborted on its watchdog -> phi and stats are NaN
  const float* KS;        // DIRECT: m x d (ld = d)
  const float* R;         // DIRECT
  const float* O;         // GRAM: nsplit x m_pad x ldo
  const float* scal;      // DIRECT: m x 4 ; GRAM: nsplit x m_pad x 4
  int nsplit;
  int scal_parts;         // GRAM: number of scal slices to sum (0 -> nsplit)
  int64_t m_pad;
  int64_t ldo;
  int64_t dpad;           // GRAM: column offset of the K Xc block inside a row of O
  const float* mu;        // GRAM: column mean used for centring (d)
  const float* Q; int64_t ldq;
  const float* Sq;        // may be null when stats are not requested
  int64_t m, n, d;
  const float* h_dev;
  float* phi; int64_t ldphi;
  float* aux;             // m x 2 x d or null
  double* stats_dev;      // double[4] or null
  double* partials;       // ws: >= 2 * epilogue_blocks(m) doubles
};
int epilogue_blocks(int64_t m);
int launch_epilogue(cudaStream_t st, const EpilogueArgs& a);

// phi paths: each fills rowacc buffers carved from ws and then calls launch_epilogue.
struct PhiArgs {
  const float* Q; const float* Sq; int64_t m, ldq;
  const float* X; const float* S; int64_t n, ldx;
  int64_t d; const float* h_dev;
  float* phi; int64_t ldphi; float* aux; double* stats_dev;
};
int phi_simt(cudaStream_t st, const PhiArgs& a, void* ws, size_t ws_bytes, bool dry, size_t* need);
int phi_generic(cudaStream_t st, const PhiArgs& a, void* ws, size_t ws_bytes, bool dry, size_t* need);
int phi_tc(cudaStream_t st, const PhiArgs& a, void* ws, size_t ws_bytes, bool dry, size_t* need);
int phi_large(cudaStream_t st, const PhiArgs& a, void* ws, size_t ws_bytes, bool dry, size_t* need);

// column mean (phi_generic.cu): deterministic two-stage; part needs colmean_partials(n,d) doubles
int launch_colmean(cudaStream_t st, const float* X, int64_t n, int64_t d, int64_t ldx, double* part, float* mu);
int64_t colmean_partials(int64_t n, int64_t d);
size_t median_workspace(int64_t n, int64_t d);

// pair engine entry points (pair_engine.cu)
int launch_kmat(cudaStream_t st, const float* Q, int64_t m, int64_t ldq, const float* X, int64_t n, int64_t ldx,
                int64_t d, const float* h_dev, float* K, int64_t ldk);
int launch_ksum(cudaStream_t st, const float* A, int64_t na, int64_t lda, const float* B, int64_t nb, int64_t ldb,
                int64_t d, const float* h_dev, float h_host, int symmetric, double* out_dev, double scale,
                int accumulate, void* ws, size_t ws_bytes, bool dry, size_t* need);

}  // namespace nsvgd
