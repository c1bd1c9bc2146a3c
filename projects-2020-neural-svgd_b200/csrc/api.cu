// api.cu — extern "C" entry points of libnsvgd.so: error plumbing, path dispatch, the composite
// SVGD step.  Reference boundary: models.KernelGradient.gradient (models.py:662-675) and
// models.Particles._step (models.py:143-167).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace nsvgd {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int device_sm_count() {
  static std::atomic<int> cached[64];          // per device ordinal
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (int c = cached[dev].load(std::memory_order_relaxed)) return c;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
  cached[dev].store(sms, std::memory_order_relaxed);
  return sms;
}

int ensure_dyn_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;   // (kernel, device) -> largest size configured so far
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  std::lock_guard<std::mutex> lk(mu);
  int& cur = done[{kernel, dev}];
  if (cur >= bytes) return NSVGD_OK;
  NSVGD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  cur = bytes;
  return NSVGD_OK;
}

// ---- optional event timing of the dominant kernels (off by default; not graph-capturable when on)
struct ProfSlot { cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfSlot> g_prof[16];
static std::mutex g_prof_mu;

void prof_begin(int tag, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfSlot s{};
  if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
  cudaEventRecord(s.a, st);
  g_prof[tag].push_back(s);
}
void prof_end(int tag, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof[tag].empty()) cudaEventRecord(g_prof[tag].back().b, st);
}

static int have_device() {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    (void)cudaGetLastError();
    set_error("no CUDA device available (%s); libnsvgd has no CPU fallback", cudaGetErrorString(e));
    return NSVGD_ERR_CUDA;
  }
  return NSVGD_OK;
}

// path choice (SURVEY 7.1 / north_star): SIMT FP32+MUFU for small d, tcgen05 where the
// contraction is a real GEMM, generic otherwise.  NSVGD_FORCE_PATH={simt,tc,generic} overrides AUTO.
static int choose_path(int path, int64_t n, int64_t m, int64_t d) {
  if (path == NSVGD_PATH_AUTO) {
    const char* env = getenv("NSVGD_FORCE_PATH");
    if (env && *env) {
      if (!strcmp(env, "simt")) path = NSVGD_PATH_SIMT;
      else if (!strcmp(env, "tc")) path = NSVGD_PATH_TC;
      else if (!strcmp(env, "generic")) path = NSVGD_PATH_GENERIC;
      else if (!strcmp(env, "large")) path = NSVGD_PATH_LARGE;
    }
  }
  if (path != NSVGD_PATH_AUTO) return path;
  if (d <= 32) return NSVGD_PATH_SIMT;
  if (d <= 128 && n >= 1024 && m >= 256) return NSVGD_PATH_TC;
  if (d > 128 && n >= 1024 && m >= 256) return NSVGD_PATH_LARGE;
  return NSVGD_PATH_GENERIC;
}

static int phi_dispatch(cudaStream_t st, const PhiArgs& a, int path, void* ws, size_t ws_bytes, bool dry, size_t* need) {
  int p = choose_path(path, a.n, a.m, a.d);
  switch (p) {
    case NSVGD_PATH_SIMT: return phi_simt(st, a, ws, ws_bytes, dry, need);
    case NSVGD_PATH_TC: return phi_tc(st, a, ws, ws_bytes, dry, need);
    case NSVGD_PATH_GENERIC: return phi_generic(st, a, ws, ws_bytes, dry, need);
    case NSVGD_PATH_LARGE: return phi_large(st, a, ws, ws_bytes, dry, need);
    default: set_error("phi: bad path %d", path); return NSVGD_ERR_INVALID;
  }
}

}  // namespace nsvgd

using namespace nsvgd;

extern "C" int nsvgd_version(void) { return 100; }

extern "C" int nsvgd_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  for (auto& v : g_prof) {
    for (auto& s : v) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    v.clear();
  }
  return NSVGD_OK;
}

extern "C" int nsvgd_profile_read(int tag, double* total_ms, int* count) {
  NSVGD_REQUIRE(tag >= 0 && tag < 16 && total_ms && count, NSVGD_ERR_INVALID, "profile_read: bad arguments");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double tot = 0.0; int c = 0;
  for (auto& s : g_prof[tag]) {
    float ms = 0.f;
    if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { tot += ms; ++c; }
    cudaEventDestroy(s.a); cudaEventDestroy(s.b);
  }
  (void)cudaGetLastError();
  g_prof[tag].clear();
  *total_ms = tot; *count = c;
  return NSVGD_OK;
}
extern "C" const char* nsvgd_last_error(void) { return g_err; }
extern "C" uint64_t nsvgd_launch_count(void) { return g_launch_count.load(); }

static int check_phi_shapes(const char* who, const void* Q, int64_t m, int64_t ldq, const void* X, const void* S,
                            int64_t n, int64_t ldx, int64_t d) {
  NSVGD_REQUIRE(Q && X && S, NSVGD_ERR_INVALID, "%s: null pointer", who);
  NSVGD_REQUIRE(n >= 1 && m >= 1 && d >= 1, NSVGD_ERR_INVALID, "%s: need n, m, d >= 1 (n=%lld m=%lld d=%lld)", who,
                (long long)n, (long long)m, (long long)d);
  NSVGD_REQUIRE(ldq >= d && ldx >= d, NSVGD_ERR_INVALID, "%s: leading dimension smaller than d", who);
  return NSVGD_OK;
}

extern "C" int nsvgd_phi_f32(void* stream, const float* Q, const float* Sq, int64_t m, int64_t ldq, const float* X,
                             const float* S, int64_t n, int64_t ldx, int64_t d, const float* h_dev, float* phi,
                             int64_t ldphi, float* aux, double* stats_dev, int path, void* ws, size_t ws_bytes) {
  NSVGD_PROPAGATE(check_phi_shapes("phi", Q, m, ldq, X, S, n, ldx, d));
  NSVGD_REQUIRE(h_dev && phi && ldphi >= d, NSVGD_ERR_INVALID, "phi: null h_dev/phi or ldphi < d");
  NSVGD_REQUIRE(!stats_dev || Sq, NSVGD_ERR_INVALID, "phi: stats need the scores at the queries (Sq)");
  NSVGD_REQUIRE(path >= 0 && path <= 4, NSVGD_ERR_INVALID, "phi: bad path %d", path);
  NSVGD_REQUIRE(ws, NSVGD_ERR_WORKSPACE, "phi: null workspace");
  NSVGD_PROPAGATE(have_device());
  PhiArgs a{Q, Sq, m, ldq, X, S, n, ldx, d, h_dev, phi, ldphi, aux, stats_dev};
  prof_begin(2, (cudaStream_t)stream);
  const int rc = phi_dispatch((cudaStream_t)stream, a, path, ws, ws_bytes, false, nullptr);
  prof_end(2, (cudaStream_t)stream);
  return rc;
}

extern "C" int nsvgd_step_sgd_f32(void* stream, float* X, const float* S, int64_t n, int64_t d, int64_t ldx,
                                  const float* h_dev, float lr, int scaled, float lambda_reg, float* grads,
                                  double* step_aux_dev, int path, void* ws, size_t ws_bytes) {
  NSVGD_PROPAGATE(check_phi_shapes("step", X, n, ldx, X, S, n, ldx, d));
  NSVGD_REQUIRE(h_dev && step_aux_dev && ws, NSVGD_ERR_INVALID, "step: null pointer");
  NSVGD_REQUIRE(path >= 0 && path <= 4, NSVGD_ERR_INVALID, "step: bad path %d", path);
  NSVGD_PROPAGATE(have_device());
  WsCarver w(ws, ws_bytes);
  float* phi = w.take<float>((size_t)(n * d));
  double* stats = w.take<double>(4);
  w.off = (w.off + 255) / 256 * 256;
  NSVGD_REQUIRE(w.off <= ws_bytes, NSVGD_ERR_WORKSPACE, "step: workspace too small");
  PhiArgs a{X, S, n, ldx, X, S, n, ldx, d, h_dev, phi, d, nullptr, stats};
  prof_begin(2, (cudaStream_t)stream);
  const int rc = phi_dispatch((cudaStream_t)stream, a, path, (char*)ws + w.off, ws_bytes - w.off, false, nullptr);
  prof_end(2, (cudaStream_t)stream);
  NSVGD_PROPAGATE(rc);
  return nsvgd_apply_update_f32(stream, X, phi, n, d, ldx, d, stats, lr, scaled, lambda_reg, grads, step_aux_dev);
}

// ---- KSD U / V statistics through the fused phi* kernels ------------------------------------------------
// stein_discrepancy(xs, logp, phi*) (stein.py:99-106) IS the V-statistic ksd_squared_v (stein.py:327-349), and the diagonal of
// u_p is |s_i|^2 + d/h^2, so  U = (n^2 V - sum_i (|s_i|^2 + d/h^2)) / (n (n - 1))  (stein.py:299-323).  The statistics come out
// of the phi* kernels for free, i.e. on the tensor cores for d > 32: cfg4-sized KSD tracking costs one phi* pass (3.3 ms) instead
// of the fp32 SIMT pair tiles (n^2/2 pairs x ~3 d flop).
namespace nsvgd {
constexpr int KSD_SUMSQ_CTAS = 256;
// per-CTA partial sums of squares (grid-stride in a fixed order; summed in index order below: deterministic)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ A, int64_t n, int64_t d, int64_t lda, double* partials) {
  __shared__ double sh[32];
  double acc = 0.0;
  const int64_t total = n * d;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const float v = A[(t / d) * lda + (t % d)];
    acc += (double)v * (double)v;
  }
  const double tot = block_sum(acc, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}
__global__ void ksd_from_stats_kernel(const double* stats, const double* sumsq_partials, int64_t n, int64_t d, const float* h_dev,
                                      int variant, double* out) {
  const double nn = (double)n;
  const double V = stats[0] / nn;                             // mean_i [ s_i . phi_i + div phi_i ]
  if (variant == NSVGD_KSD_V) { *out = V; return; }
  double sumsq = 0.0;
  for (int i = 0; i < KSD_SUMSQ_CTAS; ++i) sumsq += sumsq_partials[i];
  const double h = (double)*h_dev;
  *out = (nn * nn * V - (sumsq + nn * (double)d / (h * h))) / (nn * (nn - 1.0));
}
}  // namespace nsvgd

extern "C" int nsvgd_ksd_tc_workspace_bytes(int64_t n, int64_t d, size_t* out_bytes);

extern "C" int nsvgd_ksd_tc_f32(void* stream, const float* X, const float* S, int64_t n, int64_t d, int64_t ldx, const float* h_dev,
                                int variant, double* out_dev, void* ws, size_t ws_bytes) {
  NSVGD_REQUIRE(X && S && h_dev && out_dev && ws, NSVGD_ERR_INVALID, "ksd_tc: null pointer");
  NSVGD_REQUIRE(variant == NSVGD_KSD_U || variant == NSVGD_KSD_V, NSVGD_ERR_INVALID, "ksd_tc: the U and V statistics only");
  NSVGD_REQUIRE(n >= 2 && d >= 1 && ldx >= d, NSVGD_ERR_INVALID, "ksd_tc: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  WsCarver w(ws, ws_bytes);
  float* phi = w.take<float>((size_t)(n * d));
  double* stats = w.take<double>(4);
  double* sumsq = w.take<double>(KSD_SUMSQ_CTAS);
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "ksd_tc: workspace too small (%zu < %zu)", ws_bytes, w.off);
  w.off = (w.off + 255) / 256 * 256;
  PhiArgs a{X, S, n, ldx, X, S, n, ldx, d, h_dev, phi, d, nullptr, stats};
  NSVGD_PROPAGATE(phi_dispatch(st, a, NSVGD_PATH_AUTO, (char*)ws + w.off, ws_bytes - w.off, false, nullptr));
  if (variant == NSVGD_KSD_U) {
    sumsq_kernel<<<KSD_SUMSQ_CTAS, 256, 0, st>>>(S, n, d, ldx, sumsq);
    NSVGD_LAUNCH_CHECK();
  }
  ksd_from_stats_kernel<<<1, 1, 0, st>>>(stats, sumsq, n, d, h_dev, variant, out_dev);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

extern "C" int nsvgd_workspace_bytes(int op, int64_t n, int64_t m, int64_t d, int path, size_t* out_bytes) {
  NSVGD_REQUIRE(out_bytes, NSVGD_ERR_INVALID, "workspace_bytes: null out pointer");
  NSVGD_REQUIRE(n >= 1 && d >= 1, NSVGD_ERR_INVALID, "workspace_bytes: bad shape");
  if (m < 1) m = n;
  size_t need = 0;
  auto phi_need = [&](int64_t mm) -> size_t {
    PhiArgs a{};
    a.m = mm; a.n = n; a.d = d; a.ldq = d; a.ldx = d; a.ldphi = d;
    size_t best = 0;
    const int p0 = path == NSVGD_PATH_AUTO ? NSVGD_PATH_SIMT : path;
    const int p1 = path == NSVGD_PATH_AUTO ? NSVGD_PATH_LARGE : path;
    for (int p = p0; p <= p1; ++p) {
      if (p == NSVGD_PATH_SIMT && d > 32) continue;
      if (p == NSVGD_PATH_TC && d > 128) continue;
      if (p == NSVGD_PATH_LARGE && d <= 128 && path == NSVGD_PATH_AUTO) continue;
      size_t s = 0;
      if (phi_dispatch(nullptr, a, p, nullptr, 0, true, &s) == NSVGD_OK) best = std::max(best, s);
    }
    return best;
  };
  const int64_t tiles_n = ceil_div(n, 64), tiles_m = ceil_div(m, 64);
  switch (op) {
    case NSVGD_OP_MEDIAN: need = median_workspace(n, d); break;
    case NSVGD_OP_PHI: need = phi_need(m); break;
    case NSVGD_OP_STEP: need = phi_need(n) + (size_t)(n * d) * sizeof(float) + 1024; break;
    case NSVGD_OP_KSD: need = 2 * (size_t)(tiles_n * std::max(tiles_n, tiles_m)) * sizeof(double) + 256; break;   // KSD + d/dh partials
    case NSVGD_OP_MMD: {
      const int64_t t = std::max(tiles_n, tiles_m);
      need = (size_t)(t * t) * sizeof(double) + (size_t)((n + m) * d) * sizeof(float) + 1024;
      break;
    }
    default: set_error("workspace_bytes: bad op %d", op); return NSVGD_ERR_INVALID;
  }
  *out_bytes = need + 512;
  return NSVGD_OK;
}

extern "C" int nsvgd_ksd_tc_workspace_bytes(int64_t n, int64_t d, size_t* out_bytes) {
  NSVGD_REQUIRE(out_bytes, NSVGD_ERR_INVALID, "ksd_tc_workspace_bytes: null pointer");
  size_t phi = 0;
  NSVGD_PROPAGATE(nsvgd_workspace_bytes(NSVGD_OP_PHI, n, n, d, NSVGD_PATH_AUTO, &phi));
  *out_bytes = phi + (size_t)(n * d) * sizeof(float) + 4096;
  return NSVGD_OK;
}
