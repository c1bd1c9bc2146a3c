// phi_large.cu — phi* for d > 128 (BNN weight space: d = 4096, n = 16384) on the tensor cores.
// [stein.py:180-222 + kernels.py:52-65; closed forms in SURVEY.md 8a; SURVEY 7.2-3: "d > 128 cannot keep the output
//  resident in TMEM ... materialise K and run a plain GEMM"]
//
// With d in the thousands both contractions are plain GEMMs, so they go to the library (cuBLAS bf16 tensor-op GEMMs with
// fp32 accumulation; the prompt's rule: cuBLAS for plain library GEMMs, hand-written kernels for the fused ops) and the
// hand-written part is what sits between them:
//   prep   : centre, split every operand into two bf16 pieces (x = b0 + b1, 2^-18 relative), |xc|^2, s.xc,
//            W = S - Xc/h^2 (COMBINED form, phi_tc.cu)
//   Gram   : G[mc x nk] = Qc Xc^T          3 GEMM passes  b0 b0' + b1 b0' + b0 b1'   (K = d)
//   pair   : k = exp2(2c g - c|q|^2 - c|x|^2) -> two bf16 pieces of k, row statistics rowsum / sum k c_j / sum k r2 /
//            sum k g (one CTA per row, fp32x2 math, block reduction)                    [pair_large_kernel]
//   P.W    : O_s[mc x d] = P W               3 GEMM passes                              (K = nk <= 4096)
//   epilogue (update.cu, GRAMW form) adds the column-chunk partials O_s in fp32 RN.
// Column chunks of <= 4096 particles keep every tensor-core accumulation run at <= 768 MMAs, the same budget as the fused
// kernel (TMEM accumulator truncation).  Scratch: G 64 MiB + P pieces 64 MiB per 4096 x 4096 chunk.
#include <cublas_v2.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace nsvgd {

using namespace tc;

constexpr int64_t LG_MC = 4096, LG_NK = 4096;
constexpr size_t LG_CUBLAS_WS = 32ull << 20;

// one warp per row: centred bf16 pieces, norms, c_j, and (when S is given) the pieces of W = S - Xc/h^2
__global__ void __launch_bounds__(256) large_prep_kernel(const float* __restrict__ X, const float* __restrict__ S, int64_t n,
                                                         int64_t rows_pad, int64_t d, int64_t ldx, int64_t ldk,
                                                         const float* __restrict__ mu, const float* __restrict__ h_dev,
                                                         __nv_bfloat16* __restrict__ Xb0, __nv_bfloat16* __restrict__ Xb1,
                                                         __nv_bfloat16* __restrict__ Wb0, __nv_bfloat16* __restrict__ Wb1,
                                                         float* __restrict__ xn, float* __restrict__ cv) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows_pad) return;
  const bool rv = row < n;
  float inv_h2 = 0.f;
  if (S) { const float h = *h_dev; inv_h2 = 1.f / (h * h); }
  float acc_n = 0.f, acc_c = 0.f;
  for (int64_t c0 = 2 * lane; c0 < ldk; c0 += 64) {
    float v2[2] = {0.f, 0.f}, w2[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int64_t c = c0 + e;
      if (rv && c < d) {
        const float v = X[row * ldx + c] - mu[c];
        acc_n = fmaf(v, v, acc_n);
        v2[e] = v;
        if (S) {
          const float sv = S[row * ldx + c];
          acc_c = fmaf(sv, v, acc_c);
          w2[e] = fmaf(-v, inv_h2, sv);
        }
      }
    }
    const uint32_t p0 = pack_bf16x2(v2[0], v2[1]);
    const uint32_t p1 = pack_bf16x2(v2[0] - bf16lo_to_float(p0), v2[1] - bf16hi_to_float(p0));
    *reinterpret_cast<uint32_t*>(&Xb0[row * ldk + c0]) = p0;
    *reinterpret_cast<uint32_t*>(&Xb1[row * ldk + c0]) = p1;
    if (S) {
      const uint32_t q0 = pack_bf16x2(w2[0], w2[1]);
      const uint32_t q1 = pack_bf16x2(w2[0] - bf16lo_to_float(q0), w2[1] - bf16hi_to_float(q0));
      *reinterpret_cast<uint32_t*>(&Wb0[row * ldk + c0]) = q0;
      *reinterpret_cast<uint32_t*>(&Wb1[row * ldk + c0]) = q1;
    }
  }
  acc_n = warp_sum(acc_n); acc_c = warp_sum(acc_c);
  if (lane == 0) {
    xn[row] = rv ? acc_n : 1e30f;     // padded particle: k underflows to 0
    if (cv) cv[row] = rv ? acc_c : 0.f;
  }
}

// one CTA per query row of the chunk: G row -> two bf16 pieces of k + the four row statistics of this column chunk
__global__ void __launch_bounds__(256) pair_large_kernel(const float* __restrict__ G, int64_t nk, const float* __restrict__ xn,
                                                         const float* __restrict__ cv, const float* __restrict__ qn,
                                                         const float* __restrict__ h_dev, __nv_bfloat16* __restrict__ Pb0,
                                                         __nv_bfloat16* __restrict__ Pb1, float* __restrict__ scal) {
  __shared__ float sh[4][8];
  const int64_t row = blockIdx.x;
  const float h = *h_dev;
  const float cs = NSVGD_HALF_LOG2E / (h * h);
  const float ai = -cs * qn[row];
  const float* g = G + row * nk;
  float rowsum = 0.f, kc = 0.f, ke = 0.f, kg = 0.f;
  for (int64_t j = 2 * threadIdx.x; j < nk; j += 512) {          // nk is a multiple of 8: pairs never straddle the end
    const float2 gg = *reinterpret_cast<const float2*>(g + j);
    const float2 xs = *reinterpret_cast<const float2*>(xn + j);
    const float2 cc = *reinterpret_cast<const float2*>(cv + j);
    const float ex0 = fmaf(xs.x, -cs, fmaf(gg.x, 2.f * cs, ai));
    const float ex1 = fmaf(xs.y, -cs, fmaf(gg.y, 2.f * cs, ai));
    const float k0 = ex2_approx(ex0), k1 = ex2_approx(ex1);
    const uint32_t a0 = pack_bf16x2(k0, k1);
    const float h0 = bf16lo_to_float(a0), h1 = bf16hi_to_float(a0);
    const uint32_t a1 = pack_bf16x2(k0 - h0, k1 - h1);
    *reinterpret_cast<uint32_t*>(&Pb0[row * nk + j]) = a0;
    *reinterpret_cast<uint32_t*>(&Pb1[row * nk + j]) = a1;
    const float q0 = h0 + bf16lo_to_float(a1), q1 = h1 + bf16hi_to_float(a1);   // the quantised k the GEMM consumes
    rowsum += q0 + q1;
    kc = fmaf(q0, cc.x, fmaf(q1, cc.y, kc));
    ke = fmaf(q0, ex0, fmaf(q1, ex1, ke));
    kg = fmaf(q0, gg.x, fmaf(q1, gg.y, kg));
  }
  rowsum = warp_sum(rowsum); kc = warp_sum(kc); ke = warp_sum(ke); kg = warp_sum(kg);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[0][w] = rowsum; sh[1][w] = kc; sh[2][w] = ke; sh[3][w] = kg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f, c = 0.f, e = 0.f, gsum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { r += sh[0][i]; c += sh[1][i]; e += sh[2][i]; gsum += sh[3][i]; }
    *reinterpret_cast<float4*>(scal + row * 4) = make_float4(r, c, -e / cs, gsum);   // b = sum k r2 = -sum k e / cs
  }
}

static cublasHandle_t large_handle() {
  static cublasHandle_t handles[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!handles[dev]) {
    if (cublasCreate(&handles[dev]) != CUBLAS_STATUS_SUCCESS) { handles[dev] = nullptr; return nullptr; }
  }
  return handles[dev];
}

#define NSVGD_CUBLAS(expr)                                                                   \
  do {                                                                                       \
    cublasStatus_t s_ = (expr);                                                              \
    NSVGD_REQUIRE(s_ == CUBLAS_STATUS_SUCCESS, NSVGD_ERR_CUDA, "cuBLAS status %d at %s:%d", (int)s_, __FILE__, __LINE__); \
  } while (0)

// C[rows_c x cols_c] (row-major, ldc) (+)= A[rows_c x k] B  with the row-major operands described in the file header
static int gemm3(cublasHandle_t hd, cublasOperation_t opa, int m, int n, int k, const __nv_bfloat16* A0, const __nv_bfloat16* A1,
                 int lda, const __nv_bfloat16* B0, const __nv_bfloat16* B1, int ldb, float* Cm, int ldc) {
  const float one = 1.f, zero = 0.f;
  const __nv_bfloat16* As[3] = {A0, A1, A0};
  const __nv_bfloat16* Bs[3] = {B0, B0, B1};
  for (int pass = 0; pass < 3; ++pass)
    NSVGD_CUBLAS(cublasGemmEx(hd, opa, CUBLAS_OP_N, m, n, k, &one, As[pass], CUDA_R_16BF, lda, Bs[pass], CUDA_R_16BF, ldb,
                              pass == 0 ? &zero : &one, Cm, CUDA_R_32F, ldc, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT_TENSOR_OP));
  return NSVGD_OK;
}

int phi_large(cudaStream_t st, const PhiArgs& a, void* ws, size_t ws_bytes, bool dry, size_t* need) {
  const int64_t ldk = round_up(a.d, 8);
  const int64_t n_pad = round_up(a.n, 8);
  const int64_t ldo = round_up(a.d, 4);
  const int nsplit = (int)ceil_div(n_pad, LG_NK);
  const int64_t mc = std::min<int64_t>(LG_MC, a.m), nkc = std::min<int64_t>(LG_NK, n_pad);
  const bool q_is_x = (a.Q == a.X && a.m == a.n && a.ldq == a.ldx);
  WsCarver w(ws, ws_bytes, dry);
  float* mu = w.take<float>((size_t)a.d);
  double* mpart = w.take<double>((size_t)colmean_partials(a.n, a.d));
  __nv_bfloat16* Xb0 = w.take<__nv_bfloat16>((size_t)(n_pad * ldk));
  __nv_bfloat16* Xb1 = w.take<__nv_bfloat16>((size_t)(n_pad * ldk));
  __nv_bfloat16* Wb0 = w.take<__nv_bfloat16>((size_t)(n_pad * ldk));
  __nv_bfloat16* Wb1 = w.take<__nv_bfloat16>((size_t)(n_pad * ldk));
  float* xn = w.take<float>((size_t)n_pad);
  float* cv = w.take<float>((size_t)n_pad);
  __nv_bfloat16 *Qb0 = Xb0, *Qb1 = Xb1;
  float* qn = xn;
  if (!q_is_x || dry) {
    Qb0 = w.take<__nv_bfloat16>((size_t)(a.m * ldk));
    Qb1 = w.take<__nv_bfloat16>((size_t)(a.m * ldk));
    qn = w.take<float>((size_t)a.m);
  }
  float* G = w.take<float>((size_t)(mc * nkc));
  __nv_bfloat16* Pb0 = w.take<__nv_bfloat16>((size_t)(mc * nkc));
  __nv_bfloat16* Pb1 = w.take<__nv_bfloat16>((size_t)(mc * nkc));
  float* O = w.take<float>((size_t)nsplit * a.m * ldo);
  float* scal = w.take<float>((size_t)nsplit * a.m * 4);
  double* partials = w.take<double>((size_t)(2 * epilogue_blocks(a.m)));
  void* cws = w.take<char>(LG_CUBLAS_WS);
  if (need) *need = w.off;
  if (dry) return NSVGD_OK;
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "phi_large: workspace too small (%zu < %zu)", ws_bytes, w.off);
  NSVGD_REQUIRE(!a.aux, NSVGD_ERR_UNSUPPORTED, "phi_large: the [drift, repulsion] aux output is served by the generic path");
  NSVGD_REQUIRE(n_pad < (1ll << 31) && a.m < (1ll << 31) && ldk < (1ll << 31), NSVGD_ERR_INVALID, "phi_large: shape too large");
  cublasHandle_t hd = large_handle();
  NSVGD_REQUIRE(hd, NSVGD_ERR_CUDA, "phi_large: cublasCreate failed");
  NSVGD_CUBLAS(cublasSetStream(hd, st));
  NSVGD_CUBLAS(cublasSetWorkspace(hd, cws, LG_CUBLAS_WS));

  NSVGD_PROPAGATE(launch_colmean(st, a.X, a.n, a.d, a.ldx, mpart, mu));
  large_prep_kernel<<<(unsigned)ceil_div(n_pad * 32, 256), 256, 0, st>>>(a.X, a.S, a.n, n_pad, a.d, a.ldx, ldk, mu, a.h_dev, Xb0,
                                                                        Xb1, Wb0, Wb1, xn, cv);
  NSVGD_LAUNCH_CHECK();
  if (!q_is_x) {
    large_prep_kernel<<<(unsigned)ceil_div(a.m * 32, 256), 256, 0, st>>>(a.Q, nullptr, a.m, a.m, a.d, a.ldq, ldk, mu, a.h_dev, Qb0,
                                                                        Qb1, nullptr, nullptr, qn, nullptr);
    NSVGD_LAUNCH_CHECK();
  }
  prof_begin(0, st);
  for (int64_t i0 = 0; i0 < a.m; i0 += LG_MC) {
    const int64_t mi = std::min<int64_t>(LG_MC, a.m - i0);
    for (int s = 0; s < nsplit; ++s) {
      const int64_t j0 = (int64_t)s * LG_NK;
      const int64_t nk = std::min<int64_t>(LG_NK, n_pad - j0);
      // G^T (nk x mi, col-major, ld nk) = Xc_chunk (row-major nk x d) . Qc_chunk^T
      NSVGD_PROPAGATE(gemm3(hd, CUBLAS_OP_T, (int)nk, (int)mi, (int)ldk, Xb0 + j0 * ldk, Xb1 + j0 * ldk, (int)ldk, Qb0 + i0 * ldk,
                            Qb1 + i0 * ldk, (int)ldk, G, (int)nk));
      pair_large_kernel<<<(unsigned)mi, 256, 0, st>>>(G, nk, xn + j0, cv + j0, qn + i0, a.h_dev, Pb0, Pb1,
                                                      scal + ((int64_t)s * a.m + i0) * 4);
      NSVGD_LAUNCH_CHECK();
      // O_s^T (d x mi, col-major, ld ldo) = W_chunk^T (row-major nk x d viewed col-major d x nk) . P^T (nk x mi)
      NSVGD_PROPAGATE(gemm3(hd, CUBLAS_OP_N, (int)a.d, (int)mi, (int)nk, Wb0 + j0 * ldk, Wb1 + j0 * ldk, (int)ldk, Pb0, Pb1, (int)nk,
                            O + ((int64_t)s * a.m + i0) * ldo, (int)ldo));
    }
  }
  prof_end(0, st);

  EpilogueArgs e{};
  e.form = 2; e.O = O; e.scal = scal; e.nsplit = nsplit; e.scal_parts = nsplit; e.m_pad = a.m; e.ldo = ldo; e.dpad = 0; e.mu = mu;
  e.Q = a.Q; e.ldq = a.ldq; e.Sq = a.Sq; e.m = a.m; e.n = a.n; e.d = a.d; e.h_dev = a.h_dev;
  e.phi = a.phi; e.ldphi = a.ldphi; e.aux = nullptr; e.stats_dev = a.stats_dev; e.partials = partials;
  return launch_epilogue(st, e);
}

}  // namespace nsvgd
