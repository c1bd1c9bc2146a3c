// vjp.cu — backward rule of phi* (SURVEY 8f-N1): the cotangents of
//     phi_i = (1/n) sum_j k_ij v_ij,   v_ij = s_j + (q_i - x_j)/h^2,   k_ij = exp(-|q_i - x_j|^2 / (2 h^2))
// with respect to the queries Q, the particles X, the scores S and the bandwidth h, for an upstream cotangent G (m x d).
// The reference obtains them by differentiating through stein.phistar_i / get_phistar with jax.grad / jacfwd
// (stein.py:180-203; used by stein_discrepancy, stein.py:99-106, and the kernel-learning notebooks).  Closed forms, with
// c_ij = g_i . v_ij:
//     dS_j = (1/n)      sum_i k_ij g_i                                             = (K^T G)_j / n
//     dQ_i = (1/(n h^2)) [ sum_j k_ij g_i - sum_j k_ij c_ij (q_i - x_j) ]          = [rowsumK_i g_i - rowsumC_i q_i + (C X)_i] / (n h^2)
//     dX_j = (1/(n h^2)) [ sum_i k_ij c_ij (q_i - x_j) - sum_i k_ij g_i ]          = [(C^T Q)_j - colsumC_j x_j - (K^T G)_j] / (n h^2)
//     dh   = (1/n) sum_ij k_ij [ r2_ij c_ij - 2 g_i.(q_i - x_j) ] / h^3
// The pair engine (pair_engine.cu, MODE_VJP: direct differences, fp32) materialises K and C = K o c for a chunk of query
// rows; the contractions are plain fp32 SIMT GEMMs.  This is the differentiable-API path (n in the hundreds to a few
// thousand in the reference's notebooks), not the SVGD step.
#include <algorithm>

#include "common.cuh"

namespace nsvgd {

int launch_pair_vjp(cudaStream_t st, const float* Q, const float* G, int64_t m, int64_t ldq, const float* X, const float* S,
                    int64_t n, int64_t ldx, int64_t d, const float* h_dev, float* K, float* C, int64_t ldk, double* partials_dh);

// C[M x N] (+)= op(A) B, op(A) = A (M x K, TA = false) or A^T (A stored K x M, TA = true); B is K x N; row-major.
// 64 x 64 x 16 tiles, 256 threads, 4 x 4 outputs per thread.
template <bool TA>
__global__ void __launch_bounds__(256) vjp_gemm_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                       int64_t ldb, float* __restrict__ C, int64_t ldc, int64_t M, int64_t N,
                                                       int64_t K, int accumulate) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][68];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * 64, j0 = (int64_t)blockIdx.x * 64;
  float acc[4][4] = {};
  for (int64_t k0 = 0; k0 < K; k0 += 16) {
    __syncthreads();
    for (int t = threadIdx.x; t < 16 * 64; t += 256) {
      int kk, r;
      if (TA) { kk = t >> 6; r = t & 63; } else { kk = t & 15; r = t >> 4; }       // coalesced along the stored rows
      const int64_t gi = i0 + r, gk = k0 + kk;
      float v = 0.f;
      if (gi < M && gk < K) v = TA ? A[gk * lda + gi] : A[gi * lda + gk];
      As[kk][r] = v;
      const int kb = t >> 6, c = t & 63;
      const int64_t gj = j0 + c, gkb = k0 + kb;
      Bs[kb][c] = (gj < N && gkb < K) ? B[gkb * ldb + gj] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[ii][jj] = fmaf(ar[ii], br[jj], acc[ii][jj]);
    }
  }
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    const int64_t gi = i0 + ty * 4 + ii;
    if (gi >= M) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int64_t gj = j0 + tx * 4 + jj;
      if (gj < N) C[gi * ldc + gj] = (accumulate ? C[gi * ldc + gj] : 0.f) + acc[ii][jj];
    }
  }
}

template <bool TA>
static int launch_gemm(cudaStream_t st, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                       int64_t N, int64_t K, int accumulate) {
  dim3 g((unsigned)ceil_div(N, 64), (unsigned)ceil_div(M, 64));
  vjp_gemm_kernel<TA><<<g, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, accumulate);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

// [A | 1]: append a column of ones (row sums / column sums then ride along in the GEMMs)
__global__ void vjp_augment_kernel(const float* __restrict__ A, int64_t rows, int64_t d, int64_t lda, float* __restrict__ out) {
  const int64_t total = rows * (d + 1);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / (d + 1), c = t % (d + 1);
    out[t] = c < d ? A[r * lda + c] : 1.f;
  }
}

// dQ rows of one chunk: [rowsumK g - rowsumC q + C X] / (n h^2);  KXa = K [X | 1] (only its last column, rowsumK, is used
// here), CXa = C [X | 1]
__global__ void vjp_dq_kernel(const float* __restrict__ CXa, const float* __restrict__ KXa, const float* __restrict__ Q,
                              const float* __restrict__ G, int64_t rows, int64_t d, int64_t ldq, const float* __restrict__ h_dev,
                              int64_t n, float* __restrict__ dQ, int64_t lddq) {
  const float h = *h_dev;
  const float sc = 1.f / ((float)n * h * h);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < rows * d; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / d, c = t % d;
    const float rsC = CXa[r * (d + 1) + d], rsK = KXa[r * (d + 1) + d];
    dQ[r * lddq + c] = sc * (rsK * G[r * ldq + c] - rsC * Q[r * ldq + c] + CXa[r * (d + 1) + c]);
  }
}

// dX, dS from the accumulated transposed products: KtG = K^T [G | 1], CtQ = C^T [Q | 1]
__global__ void vjp_dx_ds_kernel(const float* __restrict__ KtG, const float* __restrict__ CtQ, const float* __restrict__ X,
                                 int64_t n, int64_t d, int64_t ldx, const float* __restrict__ h_dev, float* __restrict__ dX,
                                 int64_t lddx, float* __restrict__ dS, int64_t ldds) {
  const float h = *h_dev;
  const float inv_n = 1.f / (float)n;
  const float sc = inv_n / (h * h);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * d; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = t / d, c = t % d;
    const float ktg = KtG[j * (d + 1) + c];
    if (dS) dS[j * ldds + c] = inv_n * ktg;
    if (dX) dX[j * lddx + c] = sc * (CtQ[j * (d + 1) + c] - CtQ[j * (d + 1) + d] * X[j * ldx + c] - ktg);
  }
}

__global__ void __launch_bounds__(1024) vjp_dh_kernel(const double* __restrict__ p, int64_t count, int64_t n, double* out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t t = threadIdx.x; t < count; t += 1024) s += p[t];
  const double tot = block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = tot / (double)n;
}

constexpr int64_t VJP_CHUNK_ELEMS = (int64_t)1 << 24;     // K and C chunks of <= 16M floats (64 MB) each

static int64_t vjp_rows_chunk(int64_t m, int64_t n) {
  int64_t r = std::max<int64_t>(64, (VJP_CHUNK_ELEMS / std::max<int64_t>(n, 1)) / 64 * 64);
  return std::min(r, round_up(m, 64));
}

size_t phi_vjp_workspace(int64_t m, int64_t n, int64_t d) {
  const int64_t rc = vjp_rows_chunk(m, n);
  const int64_t tiles = (ceil_div(m, rc) + ceil_div(m, 64)) * ceil_div(n, 64);
  WsCarver w(nullptr, 0, true);      // the same carving as nsvgd_phi_vjp_f32
  w.take<float>((size_t)(rc * n)); w.take<float>((size_t)(rc * n));
  w.take<float>((size_t)(n * (d + 1))); w.take<float>((size_t)(m * (d + 1))); w.take<float>((size_t)(m * (d + 1)));
  w.take<float>((size_t)(rc * (d + 1))); w.take<float>((size_t)(rc * (d + 1)));
  w.take<float>((size_t)(n * (d + 1))); w.take<float>((size_t)(n * (d + 1)));
  w.take<double>((size_t)tiles);
  return w.off + 1024;
}

}  // namespace nsvgd

using namespace nsvgd;

// out = X * v or X / v per column: the pre / post transforms of a vector (ARD) bandwidth around the unit-bandwidth engine
__global__ void scale_cols_kernel(const float* __restrict__ X, int64_t n, int64_t d, int64_t ldx, const float* __restrict__ v,
                                  int invert, float* __restrict__ out, int64_t ldo) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * d; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / d, c = t % d;
    out[r * ldo + c] = invert ? X[r * ldx + c] / v[c] : X[r * ldx + c] * v[c];
  }
}

extern "C" int nsvgd_scale_cols_f32(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, const float* v, int invert,
                                    float* out, int64_t ldo) {
  NSVGD_REQUIRE(X && v && out && n >= 1 && d >= 1 && ldx >= d && ldo >= d, NSVGD_ERR_INVALID, "scale_cols: bad arguments");
  scale_cols_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n * d, 256), 4096), 256, 0, (cudaStream_t)stream>>>(X, n, d, ldx, v, invert,
                                                                                                             out, ldo);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

// C[M x N] (+)= op(A) B in fp32 (SIMT): the linear pre / post maps of a full-matrix Gaussian kernel (kernels.py:83-88)
extern "C" int nsvgd_matmul_f32(void* stream, const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, float* Cm,
                                int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate) {
  NSVGD_REQUIRE(A && B && Cm && M >= 1 && N >= 1 && K >= 1 && ldb >= N && ldc >= N, NSVGD_ERR_INVALID, "matmul: bad arguments");
  NSVGD_REQUIRE(trans_a ? lda >= M : lda >= K, NSVGD_ERR_INVALID, "matmul: bad lda");
  return trans_a ? launch_gemm<true>((cudaStream_t)stream, A, lda, B, ldb, Cm, ldc, M, N, K, accumulate)
                 : launch_gemm<false>((cudaStream_t)stream, A, lda, B, ldb, Cm, ldc, M, N, K, accumulate);
}

extern "C" int nsvgd_phi_vjp_workspace_bytes(int64_t m, int64_t n, int64_t d, size_t* out_bytes) {
  NSVGD_REQUIRE(out_bytes && m >= 1 && n >= 1 && d >= 1, NSVGD_ERR_INVALID, "phi_vjp_workspace_bytes: bad arguments");
  *out_bytes = phi_vjp_workspace(m, n, d);
  return NSVGD_OK;
}

extern "C" int nsvgd_phi_vjp_f32(void* stream, const float* Q, const float* G, int64_t m, int64_t ldq, const float* X,
                                 const float* S, int64_t n, int64_t ldx, int64_t d, const float* h_dev, float* dQ, int64_t lddq,
                                 float* dX, int64_t lddx, float* dS, int64_t ldds, double* dh_dev, void* ws, size_t ws_bytes) {
  cudaStream_t st = (cudaStream_t)stream;
  NSVGD_REQUIRE(Q && G && X && S && h_dev && ws, NSVGD_ERR_INVALID, "phi_vjp: null pointer");
  NSVGD_REQUIRE(m >= 1 && n >= 1 && d >= 1 && ldq >= d && ldx >= d, NSVGD_ERR_INVALID, "phi_vjp: bad shape");
  NSVGD_REQUIRE((!dQ || lddq >= d) && (!dX || lddx >= d) && (!dS || ldds >= d), NSVGD_ERR_INVALID, "phi_vjp: bad output leading dimension");
  const int64_t rc = vjp_rows_chunk(m, n);
  const int64_t tiles = (ceil_div(m, rc) + ceil_div(m, 64)) * ceil_div(n, 64);
  WsCarver w(ws, ws_bytes);
  float* K = w.take<float>((size_t)(rc * n));
  float* C = w.take<float>((size_t)(rc * n));
  float* Xa = w.take<float>((size_t)(n * (d + 1)));
  float* Qa = w.take<float>((size_t)(m * (d + 1)));
  float* Ga = w.take<float>((size_t)(m * (d + 1)));
  float* CXa = w.take<float>((size_t)(rc * (d + 1)));
  float* KXa = w.take<float>((size_t)(rc * (d + 1)));
  float* KtG = w.take<float>((size_t)(n * (d + 1)));
  float* CtQ = w.take<float>((size_t)(n * (d + 1)));
  double* pdh = w.take<double>((size_t)tiles);
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "phi_vjp: workspace too small (%zu < %zu)", ws_bytes, w.off);
  auto aug = [&](const float* A, int64_t rows, int64_t lda, float* out) -> int {
    vjp_augment_kernel<<<(unsigned)std::min<int64_t>(ceil_div(rows * (d + 1), 256), 4096), 256, 0, st>>>(A, rows, d, lda, out);
    NSVGD_LAUNCH_CHECK();
    return NSVGD_OK;
  };
  NSVGD_PROPAGATE(aug(X, n, ldx, Xa));
  NSVGD_PROPAGATE(aug(Q, m, ldq, Qa));
  NSVGD_PROPAGATE(aug(G, m, ldq, Ga));
  int64_t tile_off = 0;
  for (int64_t r0 = 0; r0 < m; r0 += rc) {
    const int64_t rows = std::min(rc, m - r0);
    NSVGD_PROPAGATE(launch_pair_vjp(st, Q + r0 * ldq, G + r0 * ldq, rows, ldq, X, S, n, ldx, d, h_dev, K, C, n, pdh + tile_off));
    tile_off += ceil_div(rows, 64) * ceil_div(n, 64);
    if (dQ) {
      NSVGD_PROPAGATE(launch_gemm<false>(st, C, n, Xa, d + 1, CXa, d + 1, rows, d + 1, n, 0));
      NSVGD_PROPAGATE(launch_gemm<false>(st, K, n, Xa, d + 1, KXa, d + 1, rows, d + 1, n, 0));
      vjp_dq_kernel<<<(unsigned)std::min<int64_t>(ceil_div(rows * d, 256), 4096), 256, 0, st>>>(CXa, KXa, Q + r0 * ldq, G + r0 * ldq,
                                                                                              rows, d, ldq, h_dev, n, dQ + r0 * lddq, lddq);
      NSVGD_LAUNCH_CHECK();
    }
    // transposed products accumulate over the row chunks: K^T [G | 1], C^T [Q | 1]   (A stored rows x n -> TA)
    NSVGD_PROPAGATE(launch_gemm<true>(st, K, n, Ga + r0 * (d + 1), d + 1, KtG, d + 1, n, d + 1, rows, r0 > 0));
    NSVGD_PROPAGATE(launch_gemm<true>(st, C, n, Qa + r0 * (d + 1), d + 1, CtQ, d + 1, n, d + 1, rows, r0 > 0));
  }
  if (dX || dS) {
    vjp_dx_ds_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n * d, 256), 4096), 256, 0, st>>>(KtG, CtQ, X, n, d, ldx, h_dev, dX, lddx,
                                                                                            dS, ldds);
    NSVGD_LAUNCH_CHECK();
  }
  if (dh_dev) {
    vjp_dh_kernel<<<1, 1024, 0, st>>>(pdh, tile_off, n, dh_dev);
    NSVGD_LAUNCH_CHECK();
  }
  return NSVGD_OK;
}
