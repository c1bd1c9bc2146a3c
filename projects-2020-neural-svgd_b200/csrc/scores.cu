// scores.cu — grad log p (and log p) of the reference's canonical targets, evaluated on the device so that a whole SVGD
// step (scores -> bandwidth -> phi* -> update) is enqueued without the framework's autodiff in the loop.
//   funnel   : Neal's funnel, y ~ N(0, 9), x_k | y ~ N(0, e^y)                       distributions.py:275-319
//   gaussian : N(mean, cov), diagonal or full covariance (precision given)          distributions.py:89-154, 627-639
//   logreg   : Bayesian logistic regression posterior of the covertype experiment, params = [w (F), log alpha]:
//              log p = prior_logp(w, log alpha) + N * sum_batch log sigmoid(y' x.w)
//              experiments/bayesian_logistic_regression.py:73-92 (prior), 103-114 (likelihood: a SUM over the batch),
//              169-190 (get_minibatch_logp: the sum multiplied by num_datapoints -- reproduced literally)
// Closed forms of what the reference obtains through vmap(grad(logp)); additive constants of log p are kept.
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace nsvgd {

constexpr float kLog2Pi = 1.8378770664093453f;

// ---------------------------------------------------------------------------- funnel: one warp per particle
__global__ void __launch_bounds__(256) score_funnel_kernel(const float* __restrict__ X, int64_t n, int64_t d, int64_t ldx,
                                                           float* __restrict__ S, int64_t lds, float* __restrict__ logp) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const float* x = X + row * ldx;
  const float y = x[d - 1];
  const float ey = expf(-y);
  float ss = 0.f;
  for (int64_t c = lane; c < d - 1; c += 32) {
    const float v = x[c];
    ss = fmaf(v, v, ss);
    S[row * lds + c] = -v * ey;                                   // d/dx_k = -x_k e^{-y}
  }
  ss = warp_sum(ss);
  if (lane == 0) {
    const float dm1 = (float)(d - 1);
    S[row * lds + d - 1] = -y / 9.f - 0.5f * dm1 + 0.5f * ey * ss;   // d/dy
    if (logp)      // norm.logpdf(y; 0, 3) + mvn.logpdf(x; 0, e^y I)
      logp[row] = -y * y / 18.f - logf(3.f) - 0.5f * kLog2Pi - 0.5f * ey * ss - 0.5f * dm1 * y - 0.5f * dm1 * kLog2Pi;
  }
}

// ---------------------------------------------------------------------------- gaussian: one warp per particle
// s = -(x - mu) P  (P symmetric precision, row-major d x d, or its diagonal); log p up to the normalisation constant
// log_norm supplied by the host (-(d/2) log 2 pi - (1/2) log det cov).
__global__ void __launch_bounds__(256) score_gaussian_kernel(const float* __restrict__ X, int64_t n, int64_t d, int64_t ldx,
                                                             const float* __restrict__ mean, const float* __restrict__ prec,
                                                             int diag, float log_norm, float* __restrict__ S, int64_t lds,
                                                             float* __restrict__ logp) {
  extern __shared__ float xc_s[];                 // [warps per CTA][d]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  float* xc = xc_s + (size_t)w * d;
  for (int64_t c = lane; c < d; c += 32) xc[c] = X[row * ldx + c] - mean[c];
  __syncwarp();
  float q = 0.f;
  for (int64_t c = lane; c < d; c += 32) {
    float acc = 0.f;
    if (diag) acc = xc[c] * prec[c];
    else
      for (int64_t k = 0; k < d; ++k) acc = fmaf(xc[k], prec[k * d + c], acc);     // coalesced over c
    S[row * lds + c] = -acc;
    q = fmaf(xc[c], acc, q);
  }
  q = warp_sum(q);
  if (lane == 0 && logp) logp[row] = -0.5f * q + log_norm;
}

// ---------------------------------------------------------------------------- logistic regression posterior
// One warp per particle; the minibatch (B x F) lives in shared memory.  Phase 1: lane i-strided over the data points,
// z_i = x_i . w, g_i = N y'_i sigmoid(-y'_i z_i); phase 2: lane f-strided over the features, s_f = sum_i g_i x_if - w_f e^{la}.
__global__ void __launch_bounds__(256) score_logreg_kernel(const float* __restrict__ P, int64_t n, int64_t ldp,
                                                           const float* __restrict__ Xb, const float* __restrict__ yb, int B,
                                                           int F, float num_datapoints, float a0, float b0,
                                                           float* __restrict__ S, int64_t lds, float* __restrict__ logp) {
  extern __shared__ float sm[];
  const int FS = F | 1;                                  // odd row stride: conflict-free for both access patterns
  float* xb = sm;                                        // [B][FS]
  float* ys = xb + (size_t)B * FS;                       // [B] in {-1, +1}
  float* wbuf = ys + B;                                  // [warps][F]
  float* gbuf = wbuf + (size_t)(blockDim.x >> 5) * F;    // [warps][B]
  for (int t = threadIdx.x; t < B * F; t += blockDim.x) xb[(t / F) * FS + (t % F)] = Xb[t];
  for (int t = threadIdx.x; t < B; t += blockDim.x) ys[t] = (yb[t] - 0.5f) * 2.f;           // {0,1} -> {-1,+1}  (:110)
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* wv = wbuf + (size_t)w * F;
  float* gv = gbuf + (size_t)w * B;
  for (int64_t row = (int64_t)blockIdx.x * nw + w; row < n; row += (int64_t)gridDim.x * nw) {
    const float* p = P + row * ldp;
    float w2 = 0.f;
    for (int f = lane; f < F; f += 32) { const float v = p[f]; wv[f] = v; w2 = fmaf(v, v, w2); }
    w2 = warp_sum(w2);
    const float la = p[F];
    const float alpha = expf(la);
    __syncwarp();
    float ll = 0.f;
    for (int i = lane; i < B; i += 32) {
      float z = 0.f;
      for (int f = 0; f < F; ++f) z = fmaf(xb[i * FS + f], wv[f], z);
      const float t = ys[i] * z;
      // log sigmoid(t) = -softplus(-t), evaluated stably; sigmoid(-t) = 1 - sigmoid(t)
      const float e = expf(-fabsf(t));
      ll += (t >= 0.f ? 0.f : t) - log1pf(e);
      const float sig_neg = t >= 0.f ? e / (1.f + e) : 1.f / (1.f + e);
      gv[i] = num_datapoints * ys[i] * sig_neg;
    }
    __syncwarp();
    for (int f = lane; f < F; f += 32) {
      float acc = 0.f;
      for (int i = 0; i < B; ++i) acc = fmaf(gv[i], xb[i * FS + f], acc);
      S[row * lds + f] = acc - wv[f] * alpha;                    // likelihood + d/dw of sum_f norm.logpdf(w_f; 1/sqrt(alpha))
    }
    ll = warp_sum(ll);
    if (lane == 0) {
      // d/d(log alpha) of  gamma.logpdf(alpha; a0, scale 1/b0) + sum_f norm.logpdf(w_f; scale alpha^-1/2)   (no Jacobian
      // term: the reference differentiates the density of alpha w.r.t. log alpha as it stands, :73-92)
      S[row * lds + F] = (a0 - 1.f) - b0 * alpha + 0.5f * (float)F - 0.5f * alpha * w2;
      if (logp) {
        const float lg = (a0 - 1.f) * la - b0 * alpha + a0 * logf(b0) - lgammaf(a0);
        const float lw = -0.5f * alpha * w2 + 0.5f * (float)F * la - 0.5f * (float)F * kLog2Pi;
        logp[row] = lg + lw + num_datapoints * ll;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------- linear-time KSD (stein.py:353-380)
// u_i = u_p(x_i, y_i) over the two halves of the sample, one warp per pair; then mean and the reference's
// std(ddof = 1) / (n/2) by one deterministic block.
__global__ void __launch_bounds__(256) ksd_linear_pairs_kernel(const float* __restrict__ X, const float* __restrict__ Sx,
                                                               const float* __restrict__ Y, const float* __restrict__ Sy,
                                                               int64_t half, int64_t d, int64_t ld, const float* __restrict__ h_dev,
                                                               float* __restrict__ u) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= half) return;
  float r2 = 0.f, ss = 0.f, sd = 0.f;
  for (int64_t c = lane; c < d; c += 32) {
    const float diff = X[i * ld + c] - Y[i * ld + c];
    const float sx = Sx[i * ld + c], sy = Sy[i * ld + c];
    r2 = fmaf(diff, diff, r2); ss = fmaf(sx, sy, ss); sd = fmaf(sx - sy, diff, sd);
  }
  r2 = warp_sum(r2); ss = warp_sum(ss); sd = warp_sum(sd);
  if (lane == 0) {
    const float h = *h_dev;
    const float inv_h2 = 1.f / (h * h);
    u[i] = ex2_approx(-NSVGD_HALF_LOG2E * inv_h2 * r2) * (ss + sd * inv_h2 + (float)d * inv_h2 - r2 * inv_h2 * inv_h2);
  }
}

__global__ void __launch_bounds__(1024) ksd_linear_reduce_kernel(const float* __restrict__ u, int64_t half, double* out) {
  __shared__ double sh[32];
  __shared__ double mean_s;
  double s = 0.0;
  for (int64_t t = threadIdx.x; t < half; t += 1024) s += (double)u[t];
  const double tot = block_sum(s, sh);
  if (threadIdx.x == 0) mean_s = tot / (double)half;
  __syncthreads();
  const double mean = mean_s;
  double q = 0.0;
  for (int64_t t = threadIdx.x; t < half; t += 1024) { const double dv = (double)u[t] - mean; q += dv * dv; }
  const double qq = block_sum(q, sh);
  if (threadIdx.x == 0) {
    out[0] = mean;
    out[1] = half > 1 ? sqrt(qq / (double)(half - 1)) / (double)half : 0.0;      // np.std(outs, ddof=1) / xs.shape[0]
  }
}

}  // namespace nsvgd

using namespace nsvgd;

extern "C" int nsvgd_ksd_linear_f32(void* stream, const float* X, const float* S, int64_t n, int64_t d, int64_t ld,
                                    const float* h_dev, double* out_dev, void* ws, size_t ws_bytes) {
  NSVGD_REQUIRE(X && S && h_dev && out_dev && ws, NSVGD_ERR_INVALID, "ksd_linear: null pointer");
  NSVGD_REQUIRE(n >= 2 && d >= 1 && ld >= d, NSVGD_ERR_INVALID, "ksd_linear: need n >= 2");
  const int64_t half = n / 2;                      // an odd count drops the last sample (stein.py:362-364)
  NSVGD_REQUIRE(ws_bytes >= (size_t)half * sizeof(float), NSVGD_ERR_WORKSPACE, "ksd_linear: workspace too small");
  float* u = (float*)ws;
  ksd_linear_pairs_kernel<<<(unsigned)ceil_div(half, 8), 256, 0, (cudaStream_t)stream>>>(X, S, X + half * ld, S + half * ld, half, d,
                                                                                        ld, h_dev, u);
  NSVGD_LAUNCH_CHECK();
  ksd_linear_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(u, half, out_dev);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

extern "C" int nsvgd_score_funnel_f32(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, float* S, int64_t lds,
                                      float* logp) {
  NSVGD_REQUIRE(X && S, NSVGD_ERR_INVALID, "score_funnel: null pointer");
  NSVGD_REQUIRE(n >= 1 && d >= 2 && ldx >= d && lds >= d, NSVGD_ERR_INVALID, "score_funnel: need n >= 1, d >= 2 (n=%lld d=%lld)",
                (long long)n, (long long)d);
  score_funnel_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>(X, n, d, ldx, S, lds, logp);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

extern "C" int nsvgd_score_gaussian_f32(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, const float* mean,
                                        const float* prec, int prec_is_diag, float log_norm, float* S, int64_t lds,
                                        float* logp) {
  NSVGD_REQUIRE(X && S && mean && prec, NSVGD_ERR_INVALID, "score_gaussian: null pointer");
  NSVGD_REQUIRE(n >= 1 && d >= 1 && ldx >= d && lds >= d, NSVGD_ERR_INVALID, "score_gaussian: bad shape");
  const size_t smem = (size_t)8 * d * sizeof(float);
  NSVGD_REQUIRE(smem <= 200 * 1024, NSVGD_ERR_UNSUPPORTED, "score_gaussian: d=%lld too large for the shared-memory row buffer", (long long)d);
  if (smem > 48 * 1024) NSVGD_PROPAGATE(ensure_dyn_smem((const void*)score_gaussian_kernel, (int)smem));
  score_gaussian_kernel<<<(unsigned)ceil_div(n, 8), 256, smem, (cudaStream_t)stream>>>(X, n, d, ldx, mean, prec, prec_is_diag,
                                                                                     log_norm, S, lds, logp);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

extern "C" int nsvgd_score_logreg_f32(void* stream, const float* P, int64_t n, int64_t d, int64_t ldp, const float* Xb,
                                      const float* yb, int64_t B, int64_t F, float num_datapoints, float a0, float b0, float* S,
                                      int64_t lds, float* logp) {
  NSVGD_REQUIRE(P && Xb && yb && S, NSVGD_ERR_INVALID, "score_logreg: null pointer");
  NSVGD_REQUIRE(n >= 1 && B >= 1 && F >= 1 && d == F + 1 && ldp >= d && lds >= d, NSVGD_ERR_INVALID,
                "score_logreg: params are [w (F), log alpha]: need d == F + 1 (d=%lld F=%lld)", (long long)d, (long long)F);
  const int FS = (int)F | 1;
  const size_t smem = ((size_t)B * FS + (size_t)B + (size_t)8 * F + (size_t)8 * B) * sizeof(float);
  NSVGD_REQUIRE(smem <= 200 * 1024, NSVGD_ERR_UNSUPPORTED, "score_logreg: minibatch %lld x %lld does not fit in shared memory",
                (long long)B, (long long)F);
  if (smem > 48 * 1024) NSVGD_PROPAGATE(ensure_dyn_smem((const void*)score_logreg_kernel, (int)smem));
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n, 8), (int64_t)device_sm_count() * 4);
  score_logreg_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(P, n, ldp, Xb, yb, (int)B, (int)F, num_datapoints, a0, b0, S,
                                                                  lds, logp);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}
