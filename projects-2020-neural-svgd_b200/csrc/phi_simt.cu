// phi_simt.cu — small-d (d <= 32) phi* kernel: direct differences, FP32 + MUFU.EX2,
// register accumulators, warp-shuffle reductions.  [stein.py:180-203, kernels.py:52-65]
//
// One warp owns R query rows; its 32 lanes split the inducing particles j (coalesced loads of
// x_j, s_j, each reused for the R rows).  Per pair and row: d SUB + d FFMA (r2) + 1 MUFU.EX2 +
// 2d FFMA (k s_j, k (q_i - x_j)) + 3 FFMA for the KSD extras.  The 2d+3 partial sums per row
// are reduced across lanes with xor-shuffles; lane 0 writes the DIRECT-form rowacc buffers.
#include <algorithm>

#include "common.cuh"

namespace nsvgd {

template <int D, int R>
__global__ void __launch_bounds__(128) phi_smalld_kernel(const float* __restrict__ Q, int64_t m, int64_t ldq,
                                                          const float* __restrict__ X, const float* __restrict__ S,
                                                          int64_t n, int64_t ldx, int d,
                                                          const float* __restrict__ h_dev,
                                                          float* __restrict__ KS, float* __restrict__ Rr,
                                                          float* __restrict__ scal) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t row0 = warp * R;
  if (row0 >= m) return;
  const float h = *h_dev;
  const float nc = -NSVGD_HALF_LOG2E / (h * h);

  float q[R][D], accS[R][D], accR[R][D], rs[R], aa[R], bb[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t row = min(row0 + r, m - 1);   // clamp: duplicate work, masked at the store
#pragma unroll
    for (int c = 0; c < D; ++c) {
      q[r][c] = c < d ? Q[row * ldq + c] : 0.f;
      accS[r][c] = 0.f; accR[r][c] = 0.f;
    }
    rs[r] = 0.f; aa[r] = 0.f; bb[r] = 0.f;
  }

  for (int64_t j = lane; j < n; j += 32) {
    float x[D], s[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
      x[c] = c < d ? X[j * ldx + c] : 0.f;
      s[c] = c < d ? S[j * ldx + c] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float diff[D];
      float r2 = 0.f, sd = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        diff[c] = q[r][c] - x[c];
        r2 = fmaf(diff[c], diff[c], r2);
        sd = fmaf(s[c], diff[c], sd);
      }
      const float k = ex2_approx(r2 * nc);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        accS[r][c] = fmaf(k, s[c], accS[r][c]);
        accR[r][c] = fmaf(k, diff[c], accR[r][c]);
      }
      rs[r] += k;
      aa[r] = fmaf(-k, sd, aa[r]);      // sum_j k s_j.(x_j - q_i)
      bb[r] = fmaf(k, r2, bb[r]);       // sum_j k r2
    }
  }

#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t row = row0 + r;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const float vs = warp_sum(accS[r][c]);
      const float vr = warp_sum(accR[r][c]);
      if (lane == 0 && row < m && c < d) {
        KS[row * d + c] = vs;
        Rr[row * d + c] = vr;
      }
    }
    const float v0 = warp_sum(rs[r]), v1 = warp_sum(aa[r]), v2 = warp_sum(bb[r]);
    if (lane == 0 && row < m) {
      *reinterpret_cast<float4*>(&scal[row * 4]) = make_float4(v0, v1, v2, 0.f);
    }
  }
}

template <int D, int R>
static int launch_smalld(cudaStream_t st, const PhiArgs& a, float* KS, float* Rr, float* scal) {
  const int64_t warps = ceil_div(a.m, R);
  const int64_t blocks = ceil_div(warps, 4);
  prof_begin(0, st);
  phi_smalld_kernel<D, R><<<(unsigned)blocks, 128, 0, st>>>(a.Q, a.m, a.ldq, a.X, a.S, a.n, a.ldx, (int)a.d, a.h_dev,
                                                            KS, Rr, scal);
  prof_end(0, st);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

int phi_simt(cudaStream_t st, const PhiArgs& a, void* ws, size_t ws_bytes, bool dry, size_t* need) {
  WsCarver w(ws, ws_bytes, dry);
  float* KS = w.take<float>((size_t)(a.m * a.d));
  float* Rr = w.take<float>((size_t)(a.m * a.d));
  float* scal = w.take<float>((size_t)(a.m * 4));
  double* partials = w.take<double>((size_t)(2 * epilogue_blocks(a.m)));
  if (need) *need = w.off;
  if (dry) return NSVGD_OK;
  NSVGD_REQUIRE(a.d <= 32, NSVGD_ERR_UNSUPPORTED, "phi_simt: small-d path supports d <= 32 (d=%lld)", (long long)a.d);
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "phi_simt: workspace too small (%zu < %zu)", ws_bytes, w.off);
  // rows per warp: more reuse of (x_j, s_j) when there are enough rows to fill the machine
  const bool big = a.m * 32 >= (int64_t)device_sm_count() * 2048 * 4;
  int s;
  const int d = (int)a.d;
  if (d <= 2)       s = big ? launch_smalld<2, 4>(st, a, KS, Rr, scal) : launch_smalld<2, 1>(st, a, KS, Rr, scal);
  else if (d <= 4)  s = big ? launch_smalld<4, 4>(st, a, KS, Rr, scal) : launch_smalld<4, 1>(st, a, KS, Rr, scal);
  else if (d <= 8)  s = big ? launch_smalld<8, 2>(st, a, KS, Rr, scal) : launch_smalld<8, 1>(st, a, KS, Rr, scal);
  else if (d <= 16) s = launch_smalld<16, 1>(st, a, KS, Rr, scal);
  else              s = launch_smalld<32, 1>(st, a, KS, Rr, scal);
  NSVGD_PROPAGATE(s);
  EpilogueArgs e{};
  e.form = 0; e.KS = KS; e.R = Rr; e.scal = scal; e.nsplit = 1; e.m_pad = a.m;
  e.Q = a.Q; e.ldq = a.ldq; e.Sq = a.Sq; e.m = a.m; e.n = a.n; e.d = a.d; e.h_dev = a.h_dev;
  e.phi = a.phi; e.ldphi = a.ldphi; e.aux = a.aux; e.stats_dev = a.stats_dev; e.partials = partials;
  return launch_epilogue(st, e);
}

}  // namespace nsvgd
