// small_loop.cu — K complete SVGD steps in ONE launch for latency-bound problem sizes.
//
// At n = 100, d = 2 (BASELINE configs[0]) a step is ~10^4 pairs: as separate kernels it costs ~45 launches of a few
// microseconds each (205 us per step even as a replayed CUDA graph).  Here ONE persistent CTA keeps particles, scores and
// phi* in shared memory and runs, per step,
//     log (particles, mean, std)  ->  [every `cadence`-th step: MMD tracers]  ->  scores of an analytic target
//     ->  exact median bandwidth (3-pass radix select over the defined-order pair distances, recomputed per pass)
//     ->  phi* + KSD / l2 statistics  ->  alpha  ->  x += lr alpha phi
// i.e. models.Particles.step + log for KernelGradient(scaled) with optax.sgd (models.py:143-197, 654-689), with the tracer
// cadence of models.py:180 on the device.  The bandwidth is bit-identical to nsvgd_median_h_f32 (same defined-order
// arithmetic, same order statistics, same finish formula); phi* uses direct differences in fp32.
// Limits: n <= 1024, d <= 32, n * d <= 8192, targets with device-side scores (funnel, Gaussian).
#include <math.h>

#include <algorithm>

#include "select.cuh"

namespace nsvgd {

constexpr int SL_THREADS = 1024;

struct SmallLoopArgs {
  float* X; int n, d;                       // particles, updated in place (contiguous)
  int target;                               // 0 funnel, 1 gaussian (tparams = [mean (d) | prec (d*d) or diag (d)]), tdiag
  const float* tparams; int tdiag; float log_norm;
  int median_mode;                          // 0 literal, 1 offdiag, 2 full; ignored when h_fixed > 0
  float h_fixed;
  long long r_lo, r_hi; float log_np1;      // order statistics of the median (host: median_ranks_host)
  float lr; int scaled; float lambda_reg;
  int K; long long step0;
  // per-step logs (any may be null): models.py:187-197 + the aux dict of KernelGradient.gradient
  float* log_particles; float* log_mean; float* log_std; float* log_bandwidth; float* log_logp; float* log_gnorm; float* log_gnorm_post;
  double* log_stats;                        // [K][3] = ksd, l2, alpha (or null)
  // tracers (metrics.py:112-127): Y (m x d) target samples, every `cadence`-th global step
  const float* Y; int m; int cadence; int tracer_kind;   // 1: rbf mmd only, 2: rbf + funnel mmd
  float tracer_h;
  float* mmd_out; long long* mmd_step; int* mmd_count;   // [.][2], [.], [1]
};

template <int DMAX>
__global__ void __launch_bounds__(SL_THREADS, 1) svgd_loop_small_kernel(const SmallLoopArgs a) {
  extern __shared__ __align__(16) float sl_smem[];
  const int n = a.n, d = a.d;
  const int dz = a.median_mode == 0 ? (n > d ? d : n - 1) : d;
  const float w_last = (a.median_mode == 0 && n > d) ? (float)(n - d) : 1.f;
  float* Xs = sl_smem;                       // [n][d]
  float* Ss = Xs + n * d;                    // [n][d]
  float* Ps = Ss + n * d;                    // [n][d] phi
  float* Zs = Ps + n * d;                    // [n][dz] literal transform (or unused)
  float* lp = Zs + n * d;                    // [n] log p
  float* prm = lp + n;                       // target parameters (<= d + d*d)
  unsigned int* hist = reinterpret_cast<unsigned int*>(prm + d + d * d);   // [2][NSVGD_SELECT_BINS]
  __shared__ SelectState sel;
  __shared__ double red[32];
  __shared__ double bc[4];
  __shared__ float h_s;
  __shared__ double kyy[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARP = SL_THREADS / 32;

  for (int t = tid; t < n * d; t += SL_THREADS) Xs[t] = a.X[t];
  if (a.target == 1) {
    const int np = d + (a.tdiag ? d : d * d);
    for (int t = tid; t < np; t += SL_THREADS) prm[t] = a.tparams[t];
  }
  __syncthreads();

  auto block_sum_d = [&](double v) -> double {          // deterministic, result broadcast to all threads
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
      double r = lane < NWARP ? red[lane] : 0.0;
      r = warp_sum(r);
      if (lane == 0) bc[0] = r;
    }
    __syncthreads();
    return bc[0];
  };
  // sum of k over pairs of two point sets held anywhere (A: na x d, B: nb x d), optional defunnelize of both sides;
  // symmetric: i < j only, doubled
  auto ksum = [&](const float* A, int na, const float* B, int nb, bool sym, bool funnel, float h) -> double {
    const float nc = -NSVGD_HALF_LOG2E / (h * h);
    double acc = 0.0;
    const long long total = (long long)na * nb;
    for (long long t = tid; t < total; t += SL_THREADS) {
      const int i = (int)(t / nb), j = (int)(t % nb);
      if (sym && i >= j) continue;
      const float* x = A + i * d; const float* y = B + j * d;
      float r2 = 0.f;
      if (funnel) {                       // kernels.defunnelize: [x e^{-y/2}, y/3]          kernels.py:112-116
        const float ex = expf(-0.5f * x[d - 1]), ey = expf(-0.5f * y[d - 1]);
        for (int c = 0; c < d - 1; ++c) { const float df = x[c] * ex - y[c] * ey; r2 = fmaf(df, df, r2); }
        const float dl = x[d - 1] / 3.f - y[d - 1] / 3.f;
        r2 = fmaf(dl, dl, r2);
      } else {
        for (int c = 0; c < d; ++c) { const float df = x[c] - y[c]; r2 = fmaf(df, df, r2); }
      }
      acc += (double)ex2_approx(r2 * nc);
    }
    const double s = block_sum_d(acc);
    return sym ? 2.0 * s : s;
  };

  if (a.Y && a.tracer_kind > 0) {          // mean offdiag Kyy is constant over the run
    const double s0 = ksum(a.Y, a.m, a.Y, a.m, true, false, a.tracer_h);
    const double s1 = a.tracer_kind > 1 ? ksum(a.Y, a.m, a.Y, a.m, true, true, a.tracer_h) : 0.0;
    if (tid == 0) { kyy[0] = s0 / ((double)a.m * (a.m - 1)); kyy[1] = s1 / ((double)a.m * (a.m - 1)); }
    __syncthreads();
  }

  for (int k = 0; k < a.K; ++k) {
    const long long gstep = a.step0 + k;
    // ---- log: particles, mean, std (population std, as torch.std(unbiased=False))                  models.py:187-197
    if (a.log_particles)
      for (int t = tid; t < n * d; t += SL_THREADS) a.log_particles[(size_t)k * n * d + t] = Xs[t];
    if (a.log_mean || a.log_std) {
      for (int c = warp; c < d; c += NWARP) {
        float s = 0.f;
        for (int i = lane; i < n; i += 32) s += Xs[i * d + c];
        s = warp_sum(s);
        const float mean = s / (float)n;
        float q = 0.f;
        for (int i = lane; i < n; i += 32) { const float df = Xs[i * d + c] - mean; q = fmaf(df, df, q); }
        q = warp_sum(q);
        if (lane == 0) {
          if (a.log_mean) a.log_mean[(size_t)k * d + c] = mean;
          if (a.log_std) a.log_std[(size_t)k * d + c] = sqrtf(q / (float)n);
        }
      }
    }
    // ---- tracers every `cadence`-th global step                                                      models.py:180
    if (a.Y && a.tracer_kind > 0 && a.cadence > 0 && gstep % a.cadence == 0) {
      const double nn = (double)n * (n - 1), nm = (double)n * a.m;
      const double xx0 = ksum(Xs, n, Xs, n, true, false, a.tracer_h);
      const double xy0 = ksum(Xs, n, a.Y, a.m, false, false, a.tracer_h);
      double xx1 = 0.0, xy1 = 0.0;
      if (a.tracer_kind > 1) {
        xx1 = ksum(Xs, n, Xs, n, true, true, a.tracer_h);
        xy1 = ksum(Xs, n, a.Y, a.m, false, true, a.tracer_h);
      }
      if (tid == 0) {
        const int slot = *a.mmd_count;
        a.mmd_out[2 * slot + 0] = (float)(xx0 / nn + kyy[0] - 2.0 * xy0 / nm);          // metrics.py:96-104
        a.mmd_out[2 * slot + 1] = (float)(xx1 / nn + kyy[1] - 2.0 * xy1 / nm);
        a.mmd_step[slot] = gstep;
        *a.mmd_count = slot + 1;
      }
      __syncthreads();
    }
    // ---- scores and log p of the target
    if (a.target == 0) {                 // funnel                                           distributions.py:275-319
      for (int i = tid; i < n; i += SL_THREADS) {
        const float y = Xs[i * d + d - 1];
        const float ey = expf(-y);
        float ss = 0.f;
        for (int c = 0; c < d - 1; ++c) { const float v = Xs[i * d + c]; ss = fmaf(v, v, ss); Ss[i * d + c] = -v * ey; }
        const float dm1 = (float)(d - 1);
        Ss[i * d + d - 1] = -y / 9.f - 0.5f * dm1 + 0.5f * ey * ss;
        lp[i] = -y * y / 18.f - logf(3.f) - 0.5f * 1.8378770664093453f - 0.5f * ey * ss - 0.5f * dm1 * y - 0.5f * dm1 * 1.8378770664093453f;
      }
    } else {                             // gaussian: s = -(x - mean) P                      distributions.py:89-154
      const float* mean = prm; const float* P = prm + d;
      for (int t = tid; t < n * d; t += SL_THREADS) {
        const int i = t / d, c = t % d;
        float acc = 0.f;
        if (a.tdiag) acc = (Xs[i * d + c] - mean[c]) * P[c];
        else for (int q = 0; q < d; ++q) acc = fmaf(Xs[i * d + q] - mean[q], P[q * d + c], acc);
        Ss[t] = -acc;
      }
      __syncthreads();
      for (int i = tid; i < n; i += SL_THREADS) {
        float q = 0.f;
        for (int c = 0; c < d; ++c) q = fmaf(Xs[i * d + c] - mean[c], -Ss[i * d + c], q);
        lp[i] = -0.5f * q + a.log_norm;
      }
    }
    __syncthreads();
    if (a.log_logp) for (int i = tid; i < n; i += SL_THREADS) a.log_logp[(size_t)k * n + i] = lp[i];

    // ---- bandwidth: exact median heuristic                                   kernels.py:160-172, utils.py:178-197, 653-658
    float h;
    if (a.h_fixed > 0.f) {
      h = a.h_fixed;
    } else {
      const float* Pz = Xs;
      if (a.median_mode == 0) {          // literal z-transform (pair_engine.cu::literal_z_kernel)
        for (int t = tid; t < n * dz; t += SL_THREADS) {
          const int r = t / dz, c = t % dz;
          int src;
          if (n > d) src = (c == d - 1) ? d - 1 : ((r < d - 1 && c >= r) ? c + 1 : c);
          else src = c >= r ? c + 1 : c;
          Zs[t] = Xs[r * d + src];
        }
        Pz = Zs;
      }
      if (tid == 0) {
        sel.prefix[0] = sel.prefix[1] = 0u; sel.mask = 0u; sel.nan_flag = 0u;
        sel.is_zero[0] = a.r_lo < 0; sel.is_zero[1] = a.r_hi < 0;
        sel.rank[0] = a.r_lo < 0 ? 0ull : (unsigned long long)a.r_lo;
        sel.rank[1] = a.r_hi < 0 ? 0ull : (unsigned long long)a.r_hi;
        sel.value[0] = sel.value[1] = 0.f; sel.log_np1 = a.log_np1; sel.done = 0;
      }
      __syncthreads();
      for (int pass = 0; pass < NSVGD_SELECT_PASSES; ++pass) {
        for (int t = tid; t < 2 * NSVGD_SELECT_BINS; t += SL_THREADS) hist[t] = 0u;
        __syncthreads();
        const unsigned int mask = sel.mask, p0 = sel.prefix[0], p1 = sel.prefix[1];
        const int shift = pass_shift(pass);
        const unsigned int binmask = (unsigned int)pass_bins(pass) - 1u;
        bool any_nan = false;
        for (int t = tid; t < n * n; t += SL_THREADS) {
          const int i = t / n, j = t % n;
          if (i >= j) continue;
          const float v = exact_pair_sqdist(Pz + i * dz, Pz + j * dz, dz, w_last);
          if (v != v) { any_nan = true; continue; }
          const unsigned int bits = __float_as_uint(v);
          const unsigned int bin = (bits >> shift) & binmask;
          if ((bits & mask) == p0) atomicAdd(&hist[bin], 1u);
          if ((bits & mask) == p1) atomicAdd(&hist[NSVGD_SELECT_BINS + bin], 1u);
        }
        if (any_nan) sel.nan_flag = 1u;
        __syncthreads();
        select_advance_smem32(&sel, hist, pass);
      }
      if (tid == 0) {
        float hh;
        if (sel.nan_flag) hh = __uint_as_float(0x7fc00000u);
        else {
          const float med = __fmul_rn(__fadd_rn(sel.value[0], sel.value[1]), 0.5f);
          hh = __fsqrt_rn(__fdiv_rn(__fmul_rn(0.5f, med), sel.log_np1));
        }
        h_s = hh;
      }
      __syncthreads();
      h = h_s;
    }
    if (a.log_bandwidth && tid == 0) a.log_bandwidth[k] = h;

    // ---- phi* and the KSD / l2 statistics: one warp per query row, lanes over the particles        stein.py:180-203, 99-106
    const float inv_h2 = 1.f / (h * h);
    const float nc = -NSVGD_HALF_LOG2E * inv_h2;
    const float inv_n = 1.f / (float)n;
    double t_acc = 0.0, l_acc = 0.0;
    for (int i = warp; i < n; i += NWARP) {
      float qi[DMAX], ks[DMAX], rr[DMAX];
#pragma unroll
      for (int c = 0; c < DMAX; ++c) { qi[c] = c < d ? Xs[i * d + c] : 0.f; ks[c] = 0.f; rr[c] = 0.f; }
      float rowsum = 0.f, av = 0.f, bv = 0.f;
      for (int j = lane; j < n; j += 32) {
        float r2 = 0.f, sd = 0.f;
        float df[DMAX];
#pragma unroll
        for (int c = 0; c < DMAX; ++c) {
          df[c] = c < d ? qi[c] - Xs[j * d + c] : 0.f;
          r2 = fmaf(df[c], df[c], r2);
        }
        const float kk = ex2_approx(r2 * nc);
#pragma unroll
        for (int c = 0; c < DMAX; ++c) {
          if (c < d) {
            const float sj = Ss[j * d + c];
            ks[c] = fmaf(kk, sj, ks[c]);
            rr[c] = fmaf(kk, df[c], rr[c]);
            sd = fmaf(sj, -df[c], sd);                 // s_j . (x_j - q_i)
          }
        }
        rowsum += kk; av = fmaf(kk, sd, av); bv = fmaf(kk, r2, bv);
      }
      rowsum = warp_sum(rowsum); av = warp_sum(av); bv = warp_sum(bv);
      float sphi = 0.f, l2 = 0.f;
#pragma unroll
      for (int c = 0; c < DMAX; ++c) {
        if (c < d) {
          const float ksr = warp_sum(ks[c]), rrr = warp_sum(rr[c]);
          const float p = ksr * inv_n + rrr * inv_h2 * inv_n;
          if (lane == 0) Ps[i * d + c] = p;
          sphi = fmaf(Ss[i * d + c], p, sphi);
          l2 = fmaf(p, p, l2);
        }
      }
      const float div = (av * inv_h2 + (float)d * inv_h2 * rowsum - bv * inv_h2 * inv_h2) * inv_n;
      if (lane == 0) { t_acc += (double)sphi + (double)div; l_acc += (double)l2; }
    }
    const double t_sum = block_sum_d(t_acc);
    const double l_sum = block_sum_d(l_acc);
    // ---- step scaling and update (update.cu::step_aux_body / apply_update_kernel)                    models.py:158-163, 677-689
    const double ksd = t_sum / (double)n, l2m = l_sum / (double)n;
    const double alpha_d = a.scaled ? ksd / (2.0 * (double)a.lambda_reg * l2m) : 1.0;
    const float alpha = (float)alpha_d;
    if (tid == 0) {
      const double g2 = alpha_d * alpha_d * l_sum;
      if (a.log_gnorm) a.log_gnorm[k] = (float)sqrt(g2);
      if (a.log_gnorm_post) a.log_gnorm_post[k] = (float)sqrt((double)a.lr * (double)a.lr * g2);
      if (a.log_stats) { a.log_stats[3 * k] = ksd; a.log_stats[3 * k + 1] = l2m; a.log_stats[3 * k + 2] = alpha_d; }
    }
    for (int t = tid; t < n * d; t += SL_THREADS) {
      const float g = -alpha * Ps[t];
      Xs[t] = Xs[t] - a.lr * g;
    }
    __syncthreads();
  }
  for (int t = tid; t < n * d; t += SL_THREADS) a.X[t] = Xs[t];
}

static size_t small_loop_smem(int n, int d) {
  return ((size_t)4 * n * d + n + d + (size_t)d * d) * sizeof(float) + (size_t)2 * NSVGD_SELECT_BINS * sizeof(unsigned int) + 64;
}

}  // namespace nsvgd

using namespace nsvgd;

extern "C" int nsvgd_loop_small_supported(int64_t n, int64_t d) {
  return (n >= 2 && n <= 1024 && d >= 1 && d <= 32 && n * d <= 8192 && small_loop_smem((int)n, (int)d) <= 200 * 1024) ? 1 : 0;
}

extern "C" int nsvgd_svgd_loop_small_f32(void* stream, float* X, int64_t n, int64_t d, int target, const float* tparams,
                                         int tparams_diag, float log_norm, int median_mode, float h_fixed, float lr, int scaled,
                                         float lambda_reg, int64_t K, int64_t step0, float* log_particles, float* log_mean,
                                         float* log_std, float* log_bandwidth, float* log_logp, float* log_gnorm,
                                         float* log_gnorm_post, double* log_stats, const float* Y, int64_t m, int cadence,
                                         int tracer_kind, float tracer_h, float* mmd_out, int64_t* mmd_step, int* mmd_count) {
  NSVGD_REQUIRE(X && K >= 1, NSVGD_ERR_INVALID, "loop_small: null particles or K < 1");
  NSVGD_REQUIRE(nsvgd_loop_small_supported(n, d), NSVGD_ERR_UNSUPPORTED,
                "loop_small: needs 2 <= n <= 1024, d <= 32, n * d <= 8192 (n=%lld d=%lld)", (long long)n, (long long)d);
  NSVGD_REQUIRE(target == 0 || (target == 1 && tparams), NSVGD_ERR_INVALID, "loop_small: target 0 (funnel) or 1 (gaussian + parameters)");
  NSVGD_REQUIRE(target != 0 || d >= 2, NSVGD_ERR_INVALID, "loop_small: the funnel needs d >= 2");
  NSVGD_REQUIRE(h_fixed > 0.f || (median_mode >= 0 && median_mode <= 2), NSVGD_ERR_INVALID, "loop_small: bad median mode %d", median_mode);
  NSVGD_REQUIRE(!Y || tracer_kind == 0 || (m >= 2 && mmd_out && mmd_step && mmd_count && tracer_h > 0.f), NSVGD_ERR_INVALID,
                "loop_small: the tracer needs m >= 2 target samples and its output buffers");
  SmallLoopArgs a{};
  a.X = X; a.n = (int)n; a.d = (int)d; a.target = target; a.tparams = tparams; a.tdiag = tparams_diag; a.log_norm = log_norm;
  a.median_mode = median_mode; a.h_fixed = h_fixed;
  long long r_lo = 0, r_hi = 0;
  if (!(h_fixed > 0.f)) median_ranks_host(n, median_mode, &r_lo, &r_hi);
  a.r_lo = r_lo; a.r_hi = r_hi; a.log_np1 = (float)log((double)(n + 1));
  a.lr = lr; a.scaled = scaled; a.lambda_reg = lambda_reg; a.K = (int)K; a.step0 = step0;
  a.log_particles = log_particles; a.log_mean = log_mean; a.log_std = log_std; a.log_bandwidth = log_bandwidth;
  a.log_logp = log_logp; a.log_gnorm = log_gnorm; a.log_gnorm_post = log_gnorm_post; a.log_stats = log_stats;
  a.Y = Y; a.m = (int)m; a.cadence = cadence; a.tracer_kind = Y ? tracer_kind : 0; a.tracer_h = tracer_h;
  a.mmd_out = mmd_out; a.mmd_step = (long long*)mmd_step; a.mmd_count = mmd_count;
  const size_t smem = small_loop_smem((int)n, (int)d);
  cudaStream_t st = (cudaStream_t)stream;
  if (d <= 2) {
    NSVGD_PROPAGATE(ensure_dyn_smem((const void*)svgd_loop_small_kernel<2>, (int)smem));
    svgd_loop_small_kernel<2><<<1, SL_THREADS, smem, st>>>(a);
  } else if (d <= 8) {
    NSVGD_PROPAGATE(ensure_dyn_smem((const void*)svgd_loop_small_kernel<8>, (int)smem));
    svgd_loop_small_kernel<8><<<1, SL_THREADS, smem, st>>>(a);
  } else {
    NSVGD_PROPAGATE(ensure_dyn_smem((const void*)svgd_loop_small_kernel<32>, (int)smem));
    svgd_loop_small_kernel<32><<<1, SL_THREADS, smem, st>>>(a);
  }
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}
