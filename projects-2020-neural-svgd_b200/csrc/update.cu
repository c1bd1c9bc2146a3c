// update.cu — phi* epilogue (shared by all phi paths), KSD/l2 statistics, particle updates.
//   phi_i   = (1/n) [ (K S)_i + R_i / h^2 ]                                   stein.py:180-203
//   div_i   = (1/n) [ a_i/h^2 + (d/h^2) rowsum_i - b_i/h^4 ]                   stein.py:39-43 (closed form)
//   stats   = { sum_i (s_i.phi_i + div_i), sum_i |phi_i|^2, m }               stein.py:99-106, utils.py:483-487
//   update  : grads = -alpha phi, alpha = ksd/(2 lambda l2) | 1 ; x <- x - lr grads   models.py:158-160, 677-689
#include <algorithm>

#include "common.cuh"
#include "p2p.cuh"

namespace nsvgd {

constexpr int EPI_ROWS = 8;  // one warp per row, 8 rows per CTA

int epilogue_blocks(int64_t m) { return (int)ceil_div(m, EPI_ROWS); }

__global__ void __launch_bounds__(256) epilogue_kernel(const EpilogueArgs a) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * EPI_ROWS + w;
  __shared__ double sh_t[EPI_ROWS], sh_l[EPI_ROWS];
  const float h = *a.h_dev;
  const float inv_h2 = 1.f / (h * h);
  const float inv_n = 1.f / (float)a.n;
  // a tensor-core kernel that aborted on its mbarrier watchdog leaves partial sums behind: never turn them into a
  // normal-looking phi / step -- NaN out, like any other invalid input
  const bool poisoned = a.err_flag && *a.err_flag != 0u;
  double t_row = 0.0, l_row = 0.0;
  if (row < a.m) {
    float rowsum = 0.f, aval = 0.f, bval = 0.f, kg = 0.f;
    if (a.form == 0) {
      const float4 sc = *reinterpret_cast<const float4*>(&a.scal[row * 4]);
      rowsum = sc.x; aval = sc.y; bval = sc.z;
    } else if (a.form == 3) {
      for (int s = 0; s < a.nsplit; ++s) {
        const float* o = a.O + ((int64_t)s * a.m_pad + row) * a.ldo + a.dpad;
        rowsum += o[0]; aval += o[1];                              // (K 1)_i and (K c)_i: columns of the second GEMM
      }
      for (int s = 0; s < a.scal_parts; ++s) {
        const float2 sc = *reinterpret_cast<const float2*>(&a.scal[((int64_t)s * a.m_pad + row) * 2]);
        bval += sc.x; kg += sc.y;
      }
    } else {
      const int parts = a.scal_parts > 0 ? a.scal_parts : a.nsplit;
      for (int s = 0; s < parts; ++s) {
        const float4 sc = *reinterpret_cast<const float4*>(&a.scal[((int64_t)s * a.m_pad + row) * 4]);
        rowsum += sc.x; aval += sc.y; bval += sc.z; kg += sc.w;   // aval = kc so far
      }
    }
    float sphi = 0.f, l2 = 0.f, qks = 0.f;
    for (int64_t c = lane; c < a.d; c += 32) {
      float ks, rr;
      if (a.form == 0) {
        ks = a.KS[row * a.d + c];
        rr = a.R[row * a.d + c];
      } else if (a.form >= 2) {
        float kw = 0.f;
        for (int s = 0; s < a.nsplit; ++s) kw += a.O[((int64_t)s * a.m_pad + row) * a.ldo + c];
        const float qc = a.Q[row * a.ldq + c] - a.mu[c];
        qks = fmaf(qc, kw, qks);
        const float p = fmaf(rowsum * qc, inv_h2, kw) * inv_n;
        a.phi[row * a.ldphi + c] = p;
        if (a.Sq) sphi = fmaf(a.Sq[row * a.ldq + c], p, sphi);
        l2 = fmaf(p, p, l2);
        continue;
      } else {
        ks = 0.f;
        float kx = 0.f;
        for (int s = 0; s < a.nsplit; ++s) {
          const float* o = a.O + ((int64_t)s * a.m_pad + row) * a.ldo;
          ks += o[c];
          kx += o[a.dpad + c];
        }
        const float qc = a.Q[row * a.ldq + c] - a.mu[c];
        rr = rowsum * qc - kx;
        qks = fmaf(qc, ks, qks);
      }
      const float drift = ks * inv_n;
      const float rep = rr * inv_h2 * inv_n;
      const float p = drift + rep;
      a.phi[row * a.ldphi + c] = p;
      if (a.aux) {
        a.aux[(row * 2 + 0) * a.d + c] = drift;
        a.aux[(row * 2 + 1) * a.d + c] = rep;
      }
      if (a.Sq) sphi = fmaf(a.Sq[row * a.ldq + c], p, sphi);
      l2 = fmaf(p, p, l2);
    }
    sphi = warp_sum(sphi); l2 = warp_sum(l2);
    if (a.form == 1) { qks = warp_sum(qks); aval = aval - qks; }   // a_i = (K c)_i - qc_i.(K S)_i
    if (a.form >= 2) { qks = warp_sum(qks); aval = aval - fmaf(kg, inv_h2, qks); }   // qc.(K S) = qc.(K W) + kg/h^2
    const float div = (aval * inv_h2 + (float)a.d * inv_h2 * rowsum - bval * inv_h2 * inv_h2) * inv_n;
    t_row = (double)sphi + (double)div;
    l_row = (double)l2;
    if (poisoned) {
      for (int64_t c = lane; c < a.d; c += 32) a.phi[row * a.ldphi + c] = nanf("");
      t_row = l_row = (double)nanf("");
    }
  }
  if (a.stats_dev) {
    if (lane == 0) { sh_t[w] = t_row; sh_l[w] = l_row; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0, l = 0.0;
#pragma unroll
      for (int i = 0; i < EPI_ROWS; ++i) { t += sh_t[i]; l += sh_l[i]; }
      a.partials[2 * (int64_t)blockIdx.x + 0] = t;
      a.partials[2 * (int64_t)blockIdx.x + 1] = l;
    }
  }
}


// GRAM form, vectorised: one warp per row, each lane owns 4 consecutive columns (float4), all split
// partials of a column group in flight at once.  Requires d, ldo, dpad, ldq, ldphi multiples of 4.
__global__ void __launch_bounds__(256) epilogue_gram4_kernel(const EpilogueArgs a) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * EPI_ROWS + w;
  __shared__ double sh_t[EPI_ROWS], sh_l[EPI_ROWS];
  const float h = *a.h_dev;
  const float inv_h2 = 1.f / (h * h);
  const float inv_n = 1.f / (float)a.n;
  const bool poisoned = a.err_flag && *a.err_flag != 0u;
  double t_row = 0.0, l_row = 0.0;
  if (row < a.m) {
    float rowsum = 0.f, aval = 0.f, bval = 0.f, kg = 0.f;
    const int parts = a.scal_parts > 0 ? a.scal_parts : a.nsplit;
    if (a.form == 3) {
      for (int s = lane; s < a.nsplit; s += 32) {
        const float2 rc = *reinterpret_cast<const float2*>(a.O + ((int64_t)s * a.m_pad + row) * a.ldo + a.dpad);
        rowsum += rc.x; aval += rc.y;                              // (K 1)_i and (K c)_i: columns of the second GEMM
      }
      for (int s = lane; s < parts; s += 32) {
        const float2 sc = *reinterpret_cast<const float2*>(&a.scal[((int64_t)s * a.m_pad + row) * 2]);
        bval += sc.x; kg += sc.y;
      }
    } else {
      for (int s = lane; s < parts; s += 32) {
        const float4 sc = *reinterpret_cast<const float4*>(&a.scal[((int64_t)s * a.m_pad + row) * 4]);
        rowsum += sc.x; aval += sc.y; bval += sc.z; kg += sc.w;
      }
    }
    rowsum = warp_sum(rowsum); aval = warp_sum(aval); bval = warp_sum(bval); kg = warp_sum(kg);
    const bool gramw = a.form >= 2;
    float sphi = 0.f, l2 = 0.f, qks = 0.f;
    for (int64_t c = 4 * lane; c < a.d; c += 128) {
      float4 ks = make_float4(0.f, 0.f, 0.f, 0.f), kx = ks;
      for (int s = 0; s < a.nsplit; ++s) {
        const float* o = a.O + ((int64_t)s * a.m_pad + row) * a.ldo;
        const float4 v0 = *reinterpret_cast<const float4*>(o + c);
        ks.x += v0.x; ks.y += v0.y; ks.z += v0.z; ks.w += v0.w;
        if (!gramw) {
          const float4 v1 = *reinterpret_cast<const float4*>(o + a.dpad + c);
          kx.x += v1.x; kx.y += v1.y; kx.z += v1.z; kx.w += v1.w;
        }
      }
      const float4 q4 = *reinterpret_cast<const float4*>(a.Q + row * a.ldq + c);
      const float4 m4 = *reinterpret_cast<const float4*>(a.mu + c);
      const float qc[4] = {q4.x - m4.x, q4.y - m4.y, q4.z - m4.z, q4.w - m4.w};
      const float ksv[4] = {ks.x, ks.y, ks.z, ks.w}, kxv[4] = {kx.x, kx.y, kx.z, kx.w};
      float pv[4], dr[4], rp[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float rr = rowsum * qc[e] - kxv[e];
        qks = fmaf(qc[e], ksv[e], qks);              // GRAMW: ksv = (K W), qks = qc.(K W)
        dr[e] = ksv[e] * inv_n;
        rp[e] = rr * inv_h2 * inv_n;
        pv[e] = gramw ? fmaf(rowsum * qc[e], inv_h2, ksv[e]) * inv_n : dr[e] + rp[e];
        l2 = fmaf(pv[e], pv[e], l2);
      }
      *reinterpret_cast<float4*>(a.phi + row * a.ldphi + c) = make_float4(pv[0], pv[1], pv[2], pv[3]);
      if (a.aux) {
        *reinterpret_cast<float4*>(a.aux + (row * 2 + 0) * a.d + c) = make_float4(dr[0], dr[1], dr[2], dr[3]);
        *reinterpret_cast<float4*>(a.aux + (row * 2 + 1) * a.d + c) = make_float4(rp[0], rp[1], rp[2], rp[3]);
      }
      if (a.Sq) {
        const float4 s4 = *reinterpret_cast<const float4*>(a.Sq + row * a.ldq + c);
        sphi = fmaf(s4.x, pv[0], sphi); sphi = fmaf(s4.y, pv[1], sphi);
        sphi = fmaf(s4.z, pv[2], sphi); sphi = fmaf(s4.w, pv[3], sphi);
      }
    }
    sphi = warp_sum(sphi); l2 = warp_sum(l2); qks = warp_sum(qks);
    if (gramw) qks = fmaf(kg, inv_h2, qks);                      // qc.(K S) = qc.(K W) + (sum_j k g_ij)/h^2
    aval = aval - qks;                                           // a_i = (K c)_i - qc_i.(K S)_i
    const float div = (aval * inv_h2 + (float)a.d * inv_h2 * rowsum - bval * inv_h2 * inv_h2) * inv_n;
    t_row = (double)sphi + (double)div;
    l_row = (double)l2;
    if (poisoned) {
      for (int64_t c = lane; c < a.d; c += 32) a.phi[row * a.ldphi + c] = nanf("");
      t_row = l_row = (double)nanf("");
    }
  }
  if (a.stats_dev) {
    if (lane == 0) { sh_t[w] = t_row; sh_l[w] = l_row; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0, l = 0.0;
#pragma unroll
      for (int i = 0; i < EPI_ROWS; ++i) { t += sh_t[i]; l += sh_l[i]; }
      a.partials[2 * (int64_t)blockIdx.x + 0] = t;
      a.partials[2 * (int64_t)blockIdx.x + 1] = l;
    }
  }
}

__global__ void __launch_bounds__(1024) stats_final_kernel(const double* __restrict__ partials, int64_t blocks,
                                                           double m, double* stats) {
  __shared__ double sh[32];
  double t = 0.0, l = 0.0;
  for (int64_t b = threadIdx.x; b < blocks; b += 1024) { t += partials[2 * b]; l += partials[2 * b + 1]; }
  const double tt = block_sum(t, sh);
  const double ll = block_sum(l, sh);
  if (threadIdx.x == 0) { stats[0] = tt; stats[1] = ll; stats[2] = m; stats[3] = 0.0; }
}

int launch_epilogue(cudaStream_t st, const EpilogueArgs& a) {
  const int blocks = epilogue_blocks(a.m);
  NSVGD_REQUIRE(!a.stats_dev || a.Sq, NSVGD_ERR_INVALID, "phi: stats need the scores at the queries (Sq)");
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  const bool vec4 = a.form >= 1 && a.d % 4 == 0 && a.ldo % 4 == 0 && a.dpad % 4 == 0 && a.ldq % 4 == 0 && a.ldphi % 4 == 0 &&
                    al16(a.O) && al16(a.Q) && al16(a.mu) && al16(a.phi) && (!a.Sq || al16(a.Sq)) && (!a.aux || al16(a.aux));
  if (vec4) epilogue_gram4_kernel<<<blocks, 256, 0, st>>>(a);
  else epilogue_kernel<<<blocks, 256, 0, st>>>(a);
  NSVGD_LAUNCH_CHECK();
  if (a.stats_dev) {
    stats_final_kernel<<<1, 1024, 0, st>>>(a.partials, blocks, (double)a.m, a.stats_dev);
    NSVGD_LAUNCH_CHECK();
  }
  return NSVGD_OK;
}

// ---- update ----------------------------------------------------------------------------------
__device__ __forceinline__ void step_aux_body(const double* stats, int scaled, float lambda_reg, float lr, double* aux);

__global__ void step_aux_kernel(const double* __restrict__ stats, int scaled, float lambda_reg, float lr,
                                double* __restrict__ aux) {
  if (threadIdx.x == 0 && blockIdx.x == 0) step_aux_body(stats, scaled, lambda_reg, lr, aux);
}

// row-sharded: the all-reduce of the KSD / l2 sums (4 doubles, fixed rank order -> bit-identical alpha on every
// rank) over peer memory fused into the kernel that turns them into the step scaling
__global__ void __launch_bounds__(64) step_aux_p2p_kernel(const P2PComm c, double* stats, int scaled, float lambda_reg, float lr,
                                                          double* aux) {
  p2p_allreduce_cta(c, reinterpret_cast<unsigned long long*>(stats), 4, true);   // late peer: stats poisoned with NaN -> alpha, X NaN
  if (threadIdx.x == 0) step_aux_body(stats, scaled, lambda_reg, lr, aux);
}

__device__ __forceinline__ void step_aux_body(const double* stats, int scaled, float lambda_reg, float lr, double* aux) {
  {
    double ksd = 0.0, l2 = 0.0, alpha = 1.0, sum_l = 0.0;
    if (stats) {
      const double m = stats[2];
      ksd = stats[0] / m; l2 = stats[1] / m; sum_l = stats[1];
      if (scaled) alpha = ksd / (2.0 * (double)lambda_reg * l2);     // models.py:688
    }
    aux[0] = ksd; aux[1] = l2; aux[2] = alpha;
    aux[3] = alpha * alpha * sum_l;                    // sum grads^2   (global_grad_norm^2, models.py:162)
    aux[4] = (double)lr * (double)lr * aux[3];         // sum updates^2 (models.py:163)
    aux[5] = aux[6] = aux[7] = 0.0;
  }
}

__global__ void __launch_bounds__(256) apply_update_kernel(float* __restrict__ X, const float* __restrict__ phi,
                                                           int64_t m, int64_t d, int64_t ldx, int64_t ldphi,
                                                           const double* __restrict__ aux, float lr,
                                                           float* __restrict__ grads) {
  const float alpha = (float)aux[2];
  const int64_t total = m * d;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / d, c = t % d;
    const float g = -alpha * phi[i * ldphi + c];     // grads = -alpha phi      (models.py:660, 689)
    if (grads) grads[t] = g;
    X[i * ldx + c] = X[i * ldx + c] - lr * g;        // optax.sgd + apply_updates (models.py:159-160)
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ X, const float* __restrict__ g,
                                                   float* __restrict__ m1, float* __restrict__ m2, int64_t count,
                                                   float lr, float b1, float b2, float eps, float c1, float c2) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
    const float gg = g[t];
    const float m = b1 * m1[t] + (1.f - b1) * gg;
    const float v = b2 * m2[t] + (1.f - b2) * gg * gg;
    m1[t] = m; m2[t] = v;
    const float mh = m / c1, vh = v / c2;
    X[t] = X[t] - lr * mh / (sqrtf(vh) + eps);
  }
}

}  // namespace nsvgd

using namespace nsvgd;

extern "C" int nsvgd_apply_update_f32(void* stream, float* Xrows, const float* phi, int64_t m, int64_t d, int64_t ldx,
                                      int64_t ldphi, const double* stats_dev, float lr, int scaled, float lambda_reg,
                                      float* grads, double* step_aux_dev) {
  cudaStream_t st = (cudaStream_t)stream;
  NSVGD_REQUIRE(Xrows && phi && step_aux_dev, NSVGD_ERR_INVALID, "apply_update: null pointer");
  NSVGD_REQUIRE(m >= 1 && d >= 1 && ldx >= d && ldphi >= d, NSVGD_ERR_INVALID, "apply_update: bad shape");
  NSVGD_REQUIRE(!scaled || stats_dev, NSVGD_ERR_INVALID, "apply_update: scaled update needs stats");
  step_aux_kernel<<<1, 32, 0, st>>>(stats_dev, scaled, lambda_reg, lr, step_aux_dev);
  NSVGD_LAUNCH_CHECK();
  const int64_t total = m * d;
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)device_sm_count() * 8);
  apply_update_kernel<<<blocks, 256, 0, st>>>(Xrows, phi, m, d, ldx, ldphi, step_aux_dev, lr, grads);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

extern "C" int nsvgd_p2p_apply_update_f32(void* stream, int rank, int world, void* const* blocks, float* Xrows,
                                          const float* phi, int64_t m, int64_t d, int64_t ldx, int64_t ldphi,
                                          double* stats_dev, float lr, int scaled, float lambda_reg, float* grads,
                                          double* step_aux_dev) {
  cudaStream_t st = (cudaStream_t)stream;
  P2PComm c;
  NSVGD_REQUIRE(p2p_make_comm(&c, rank, world, blocks), NSVGD_ERR_INVALID, "p2p_apply_update: bad communicator");
  NSVGD_REQUIRE(Xrows && phi && step_aux_dev && stats_dev, NSVGD_ERR_INVALID, "p2p_apply_update: null pointer");
  NSVGD_REQUIRE(m >= 1 && d >= 1 && ldx >= d && ldphi >= d, NSVGD_ERR_INVALID, "p2p_apply_update: bad shape");
  step_aux_p2p_kernel<<<1, 64, 0, st>>>(c, stats_dev, scaled, lambda_reg, lr, step_aux_dev);
  NSVGD_LAUNCH_CHECK();
  const int64_t total = m * d;
  const unsigned blocks_ = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)device_sm_count() * 8);
  apply_update_kernel<<<blocks_, 256, 0, st>>>(Xrows, phi, m, d, ldx, ldphi, step_aux_dev, lr, grads);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

extern "C" int nsvgd_adam_update_f32(void* stream, float* X, const float* grads, float* m1, float* m2, int64_t count,
                                     float lr, float b1, float b2, float eps, int64_t t) {
  cudaStream_t st = (cudaStream_t)stream;
  NSVGD_REQUIRE(X && grads && m1 && m2, NSVGD_ERR_INVALID, "adam: null pointer");
  NSVGD_REQUIRE(count >= 1 && t >= 1, NSVGD_ERR_INVALID, "adam: bad count/step");
  const float c1 = 1.f - powf(b1, (float)t), c2 = 1.f - powf(b2, (float)t);
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(count, 256), (int64_t)device_sm_count() * 8);
  adam_kernel<<<blocks, 256, 0, st>>>(X, grads, m1, m2, count, lr, b1, b2, eps, c1, c2);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}
