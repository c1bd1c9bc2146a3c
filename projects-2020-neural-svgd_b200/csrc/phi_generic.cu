// phi_generic.cu — phi* for any d: K row-chunks materialised in workspace (pair engine, direct
// differences) followed by a SIMT fp32 GEMM  O = K [S | Xc | 1 | c | xn]  in centred coordinates.
// [stein.py:180-222, kernels.py:52-65].  Serves d in (32,128) u (128,inf) and is the on-GPU
// cross-check of the tcgen05 path.  Centring (Xc = X - column mean) removes the cancellation in
// rowsum(K) q_i - K X for clouds far from the origin.
#include <algorithm>

#include "common.cuh"

namespace nsvgd {

// ---- column mean, deterministic two-stage (double accumulation) ------------------------------
// stage 1: one CTA per 64 rows; lane -> column (coalesced 128-byte rows), warp -> every 8th row; the 8 warp sums
// are combined in a fixed order.  stage 2: one CTA per 32 columns, 8 warps stride over the row blocks.
// (The first version used 128 CTAs with one thread per column over 512 rows: 160 us for n=65536, d=128.)
constexpr int MEAN_ROWS = 64;

__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ X, int64_t n, int64_t d,
                                                             int64_t ldx, double* __restrict__ part) {
  __shared__ double sh[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * MEAN_ROWS;
  const int64_t r1 = min(r0 + MEAN_ROWS, n);
  for (int64_t c0 = 0; c0 < d; c0 += 32) {
    const int64_t c = c0 + lane;
    double s = 0.0;
    if (c < d)
      for (int64_t r = r0 + w; r < r1; r += 8) s += (double)X[r * ldx + c];
    sh[w][lane] = s;
    __syncthreads();
    if (w == 0 && c < d) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += sh[i][lane];
      part[(int64_t)blockIdx.x * d + c] = t;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) colmean_final_kernel(const double* __restrict__ part, int64_t blocks, int64_t n,
                                                             int64_t d, float* __restrict__ mu) {
  // 32 warps per 32 columns: each lane's chain of (dependent) L2 round trips is blocks / 32 long
  __shared__ double sh[32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + lane;
  double s = 0.0;
  if (c < d) {
    double s0 = 0.0, s1 = 0.0;
    int64_t b = w;
    for (; b + 32 < blocks; b += 64) { s0 += part[b * d + c]; s1 += part[(b + 32) * d + c]; }
    for (; b < blocks; b += 32) s0 += part[b * d + c];
    s = s0 + s1;
  }
  sh[w][lane] = s;
  __syncthreads();
  if (w == 0 && c < d) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += sh[i][lane];
    mu[c] = (float)(t / (double)n);
  }
}

int launch_colmean(cudaStream_t st, const float* X, int64_t n, int64_t d, int64_t ldx, double* part, float* mu) {
  const int64_t blocks = ceil_div(n, MEAN_ROWS);
  colsum_partial_kernel<<<(unsigned)blocks, 256, 0, st>>>(X, n, d, ldx, part);
  NSVGD_LAUNCH_CHECK();
  colmean_final_kernel<<<(unsigned)ceil_div(d, 32), 1024, 0, st>>>(part, blocks, n, d, mu);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}
int64_t colmean_partials(int64_t n, int64_t d) { return ceil_div(n, MEAN_ROWS) * d; }

// ---- V = [S | Xc | 1 | c | xn | 0...]  (one warp per row) -------------------------------------
__global__ void __launch_bounds__(256) prep_v_kernel(const float* __restrict__ X, const float* __restrict__ S,
                                                     int64_t n, int64_t d, int64_t ldx, const float* __restrict__ mu,
                                                     float* __restrict__ V, int64_t nv) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  float cs = 0.f, xn = 0.f;
  float* v = V + row * nv;
  for (int64_t c = lane; c < d; c += 32) {
    const float s = S[row * ldx + c];
    const float xc = X[row * ldx + c] - mu[c];
    v[c] = s;
    v[d + c] = xc;
    cs = fmaf(s, xc, cs);
    xn = fmaf(xc, xc, xn);
  }
  cs = warp_sum(cs); xn = warp_sum(xn);
  if (lane == 0) {
    v[2 * d] = 1.f; v[2 * d + 1] = cs; v[2 * d + 2] = xn;
    for (int64_t c = 2 * d + 3; c < nv; ++c) v[c] = 0.f;
  }
}

// ---- SIMT GEMM  C[M x N] = A[M x K] B[K x N]  (row-major, 64x64x16 tiles, 4x4 per thread) -----
__global__ void __launch_bounds__(256) sgemm_nn_kernel(const float* __restrict__ A, int64_t lda,
                                                       const float* __restrict__ B, int64_t ldb,
                                                       float* __restrict__ C, int64_t ldc, int64_t M, int64_t N,
                                                       int64_t K) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][68];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * 64, j0 = (int64_t)blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ar = threadIdx.x >> 4, ac = threadIdx.x & 15;    // A: 16 rows x 16 k per pass
  const int br = threadIdx.x >> 6, bc = threadIdx.x & 63;    // B: 4 k x 64 cols per pass
  for (int64_t k0 = 0; k0 < K; k0 += 16) {
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int r = ar + 16 * p;
      const int64_t gi = i0 + r, gk = k0 + ac;
      As[ac][r] = (gi < M && gk < K) ? A[gi * lda + gk] : 0.f;
      const int kk = br + 4 * p;
      const int64_t gk2 = k0 + kk, gj = j0 + bc;
      Bs[kk][bc] = (gk2 < K && gj < N) ? B[gk2 * ldb + gj] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gi = i0 + ty * 4 + i;
    if (gi >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gj = j0 + tx * 4 + j;
      if (gj < N) C[gi * ldc + gj] = acc[i][j];
    }
  }
}

// ---- scal[row] = {rowsum, kc, b}: b = |qc|^2 rowsum + (K xn) - 2 qc.(K Xc) ----------------------
__global__ void __launch_bounds__(256) generic_scal_kernel(const float* __restrict__ O, int64_t nv, int64_t m, int64_t d,
                                                           const float* __restrict__ Q, int64_t ldq,
                                                           const float* __restrict__ mu, float* __restrict__ scal) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= m) return;
  const float* o = O + row * nv;
  float qq = 0.f, qkx = 0.f;
  for (int64_t c = lane; c < d; c += 32) {
    const float qc = Q[row * ldq + c] - mu[c];
    qq = fmaf(qc, qc, qq);
    qkx = fmaf(qc, o[d + c], qkx);
  }
  qq = warp_sum(qq); qkx = warp_sum(qkx);
  if (lane == 0) {
    const float rowsum = o[2 * d], kc = o[2 * d + 1], kxn = o[2 * d + 2];
    *reinterpret_cast<float4*>(&scal[row * 4]) = make_float4(rowsum, kc, qq * rowsum + kxn - 2.f * qkx, 0.f);
  }
}

int phi_generic(cudaStream_t st, const PhiArgs& a, void* ws, size_t ws_bytes, bool dry, size_t* need) {
  const int64_t nv = round_up(2 * a.d + 3, 4);
  // K chunk: aim for <= 256 MiB, at least one 64-row tile
  int64_t rows_chunk = std::max<int64_t>(64, (int64_t)((256ll << 20) / (4 * a.n)) / 64 * 64);
  rows_chunk = std::min<int64_t>(rows_chunk, round_up(a.m, 64));
  WsCarver w(ws, ws_bytes, dry);
  float* mu = w.take<float>((size_t)a.d);
  double* mpart = w.take<double>((size_t)colmean_partials(a.n, a.d));
  float* V = w.take<float>((size_t)(a.n * nv));
  float* Kc = w.take<float>((size_t)(rows_chunk * a.n));
  float* O = w.take<float>((size_t)(a.m * nv));
  float* scal = w.take<float>((size_t)(a.m * 4));
  double* partials = w.take<double>((size_t)(2 * epilogue_blocks(a.m)));
  if (need) *need = w.off;
  if (dry) return NSVGD_OK;
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "phi_generic: workspace too small (%zu < %zu)", ws_bytes, w.off);

  NSVGD_PROPAGATE(launch_colmean(st, a.X, a.n, a.d, a.ldx, mpart, mu));
  prep_v_kernel<<<(unsigned)ceil_div(a.n, 8), 256, 0, st>>>(a.X, a.S, a.n, a.d, a.ldx, mu, V, nv);
  NSVGD_LAUNCH_CHECK();
  for (int64_t r0 = 0; r0 < a.m; r0 += rows_chunk) {
    const int64_t rows = std::min<int64_t>(rows_chunk, a.m - r0);
    NSVGD_PROPAGATE(launch_kmat(st, a.Q + r0 * a.ldq, rows, a.ldq, a.X, a.n, a.ldx, a.d, a.h_dev, Kc, a.n));
    dim3 g((unsigned)ceil_div(nv, 64), (unsigned)ceil_div(rows, 64));
    sgemm_nn_kernel<<<g, 256, 0, st>>>(Kc, a.n, V, nv, O + r0 * nv, nv, rows, nv, a.n);
    NSVGD_LAUNCH_CHECK();
  }
  generic_scal_kernel<<<(unsigned)ceil_div(a.m, 8), 256, 0, st>>>(O, nv, a.m, a.d, a.Q, a.ldq, mu, scal);
  NSVGD_LAUNCH_CHECK();
  EpilogueArgs e{};
  e.form = 1; e.O = O; e.scal = scal; e.nsplit = 1; e.m_pad = a.m; e.ldo = nv; e.dpad = a.d; e.mu = mu;
  e.Q = a.Q; e.ldq = a.ldq; e.Sq = a.Sq; e.m = a.m; e.n = a.n; e.d = a.d; e.h_dev = a.h_dev;
  e.phi = a.phi; e.ldphi = a.ldphi; e.aux = a.aux; e.stats_dev = a.stats_dev; e.partials = partials;
  return launch_epilogue(st, e);
}

}  // namespace nsvgd
