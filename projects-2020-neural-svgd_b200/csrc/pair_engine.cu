// pair_engine.cu — SIMT tiled pairwise engine (any d), direct differences.
//
// One 64x64 tile of pairs per CTA (256 threads, 4x4 pairs per thread), the coordinate
// axis streamed through shared memory in chunks of 16.  Four epilogues share the tile loop:
//   HIST : radix-select histogram of the fp32 bit patterns of the squared distances, with
//          the DEFINED summation order of oracle/svgd_oracle.py::pair_sqdist_f32
//          (sequential over coordinates, separate mul and add, weighted last coordinate)
//          -> bit-exact median bandwidth   [kernels.py:160-172, utils.py:178-197, 653-658]
//   KSUM : sum of RBF kernel values (MMD building block)        [metrics.py:96-104]
//   KMAT : materialise a row-chunk of K for the generic phi path [stein.py:180-203]
//   KSD  : sum of u_p(x_i, y_j)                                  [stein.py:273-349]
#include <algorithm>
#include <math.h>
#include <stdlib.h>

#include "p2p.cuh"
#include "select.cuh"

namespace nsvgd {

constexpr int TILE = 64;
constexpr int DK = 16;
constexpr int LDS_ = TILE + 4;  // padded row of the transposed smem tiles (floats)

enum PairMode { MODE_HIST = 0, MODE_KSUM = 1, MODE_KMAT = 2, MODE_KSD = 3, MODE_VJP = 4 };


struct PairArgs {
  const float* A; int64_t na, lda;       // "row" points (tile index blockIdx.y)
  const float* B; int64_t nb, ldb;       // "column" points (tile index blockIdx.x)
  const float* SA; const float* SB;      // scores (KSD only)
  int64_t d;
  float w_last;                          // HIST: weight of the last coordinate
  int symmetric;                         // only tiles J >= I, pairs i < j
  int64_t tile_row_start, tile_row_stride;
  const float* h_dev; float h_host;      // bandwidth (device pointer wins)
  // HIST
  const SelectState* state; int pass; unsigned long long* hist; const uint32_t* run_flag;
  float* dstore;                         // HIST, pass 0: also keep every distance (one 64 x 64 block per upper-triangular tile, +inf
                                         // where i >= j) so that the later passes histogram the buffer instead of recomputing
  // KSUM / KSD
  double* partials; float diag_weight;
  float wsum;                            // KSD, ARD: sum_a 1/h_a^2 (host)
  const float* wcoord;                   // KSD: per-coordinate weights 1/h_a^2 of an ARD bandwidth (null: scalar bandwidth)
  double* partials_dh;                   // KSD / VJP: per-CTA partial sums of the derivative w.r.t. the bandwidth (or null)
  // KMAT / VJP
  float* K; int64_t ldk;
  float* C;                              // VJP: C_ij = k_ij (g_i.s_j + g_i.(q_i - x_j) / h^2), same layout as K
};


template <int MODE>
__global__ void __launch_bounds__(256) pair_tile_kernel(const PairArgs a) {
  constexpr bool EXACT = (MODE == MODE_HIST);
  if (MODE == MODE_HIST && a.run_flag && *a.run_flag == 0u) return;   // conditional fallback pass
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  __shared__ __align__(16) float As[DK][LDS_];
  __shared__ __align__(16) float Bs[DK][LDS_];
  constexpr bool SCORES = (MODE == MODE_KSD || MODE == MODE_VJP);   // a second operand per side: scores (KSD) / cotangent and scores (VJP)
  __shared__ __align__(16) float SAs[SCORES ? DK : 1][LDS_];
  __shared__ __align__(16) float SBs[SCORES ? DK : 1][LDS_];
  __shared__ float Ws[MODE == MODE_KSD ? DK : 1];      // ARD weights of the current coordinate chunk
  __shared__ unsigned int hist_s[MODE == MODE_HIST ? 2 * NSVGD_SELECT_BINS : 1];
  __shared__ double red_s[32];

  if (MODE == MODE_HIST) {
    for (int t = threadIdx.x; t < 2 * NSVGD_SELECT_BINS; t += 256) hist_s[t] = 0u;
  }
  __syncthreads();
  bool any_nan = false;
  // HIST: persistent grid-stride walk over the upper-triangular tiles (rows strided across ranks),
  // one shared-memory histogram per CTA flushed once.  Other modes: one tile per CTA.
  const int64_t tiles_j = (a.nb + TILE - 1) / TILE;
  const int64_t tiles_i = (a.na + TILE - 1) / TILE;
  const int64_t n_rows = a.tile_row_start < tiles_i ? (tiles_i - a.tile_row_start + a.tile_row_stride - 1) / a.tile_row_stride : 0;
  for (int64_t rr = blockIdx.y; rr < (MODE == MODE_HIST ? n_rows : (int64_t)blockIdx.y + 1); rr += gridDim.y) {
  const int64_t I = a.tile_row_start + a.tile_row_stride * rr;
  for (int64_t J = (MODE == MODE_HIST ? I + blockIdx.x : (int64_t)blockIdx.x);
       J < (MODE == MODE_HIST ? tiles_j : (int64_t)blockIdx.x + 1); J += gridDim.x) {
  if (MODE != MODE_HIST && a.symmetric && J < I) return;
  const int64_t i0 = I * TILE, j0 = J * TILE;

  float acc[4][4];
  float acc_ss[SCORES ? 4 : 1][SCORES ? 4 : 1];
  float acc_sd[SCORES ? 4 : 1][SCORES ? 4 : 1];
  float acc_q4[MODE == MODE_KSD ? 4 : 1][MODE == MODE_KSD ? 4 : 1];     // ARD: sum_a diff_a^2 / h_a^4
  const bool ard = MODE == MODE_KSD && a.wcoord != nullptr;
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      acc[ii][jj] = 0.f;
      if (SCORES) { acc_ss[ii][jj] = 0.f; acc_sd[ii][jj] = 0.f; }
      if (MODE == MODE_KSD) acc_q4[ii][jj] = 0.f;
    }

  // EXACT: the last coordinate is handled separately (weighted), main loop covers [0, d-1)
  const int64_t dmain = EXACT ? a.d - 1 : a.d;
  const int lr = threadIdx.x >> 4;   // 0..15: row within a 16-row load pass
  const int lc = threadIdx.x & 15;   // coordinate within the chunk
  for (int64_t c0 = 0; c0 < dmain; c0 += DK) {
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int r = lr + 16 * p;
      const int64_t c = c0 + lc;
      const bool cv = c < dmain;
      const int64_t gi = i0 + r, gj = j0 + r;
      As[lc][r] = (cv && gi < a.na) ? a.A[gi * a.lda + c] : 0.f;
      Bs[lc][r] = (cv && gj < a.nb) ? a.B[gj * a.ldb + c] : 0.f;
      if (SCORES) {
        SAs[lc][r] = (cv && gi < a.na) ? a.SA[gi * a.lda + c] : 0.f;
        SBs[lc][r] = (cv && gj < a.nb) ? a.SB[gj * a.ldb + c] : 0.f;
      }
    }
    if (MODE == MODE_KSD && threadIdx.x < DK) Ws[threadIdx.x] = (ard && c0 + threadIdx.x < dmain) ? a.wcoord[c0 + threadIdx.x] : 1.f;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < DK; ++c) {
      const float4 av = *reinterpret_cast<const float4*>(&As[c][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[c][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
      float sar[4], sbr[4];
      if (SCORES) {
        const float4 sa = *reinterpret_cast<const float4*>(&SAs[c][ty * 4]);
        const float4 sb = *reinterpret_cast<const float4*>(&SBs[c][tx * 4]);
        sar[0] = sa.x; sar[1] = sa.y; sar[2] = sa.z; sar[3] = sa.w;
        sbr[0] = sb.x; sbr[1] = sb.y; sbr[2] = sb.z; sbr[3] = sb.w;
      }
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float diff = ar[ii] - br[jj];
          if (EXACT) {
            acc[ii][jj] = __fadd_rn(acc[ii][jj], __fmul_rn(diff, diff));
          } else if (MODE == MODE_KSD) {
            // ARD: r'^2 = sum w diff^2, (s_x-s_y).(x-y)' = sum w (s_x-s_y) diff, sum w^2 diff^2; w = 1 for a scalar bandwidth
            const float wd = Ws[c] * diff;
            acc[ii][jj] = fmaf(wd, diff, acc[ii][jj]);
            acc_q4[ii][jj] = fmaf(wd, wd, acc_q4[ii][jj]);
            diff = wd;
          } else {
            acc[ii][jj] = fmaf(diff, diff, acc[ii][jj]);
          }
          if (SCORES) {
            acc_ss[ii][jj] = fmaf(sar[ii], sbr[jj], acc_ss[ii][jj]);                                    // s_x.s_y   | g_i.s_j
            acc_sd[ii][jj] = fmaf(MODE == MODE_VJP ? sar[ii] : sar[ii] - sbr[jj], diff, acc_sd[ii][jj]);  // (s_x-s_y).(x-y) | g_i.(q_i-x_j)
          }
        }
    }
  }

  if (EXACT) {
    // weighted last coordinate: t = fl(diff*diff); t = fl(w*t); acc = fl(acc + t)
    float al[4], bl[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t gi = i0 + ty * 4 + q, gj = j0 + tx * 4 + q;
      al[q] = gi < a.na ? a.A[gi * a.lda + (a.d - 1)] : 0.f;
      bl[q] = gj < a.nb ? a.B[gj * a.ldb + (a.d - 1)] : 0.f;
    }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float diff = al[ii] - bl[jj];
        float t = __fmul_rn(diff, diff);
        t = __fmul_rn(a.w_last, t);
        acc[ii][jj] = __fadd_rn(acc[ii][jj], t);
      }
  }

  // ------------------------------------------------------------------ epilogues
  if (MODE == MODE_HIST) {
    const SelectState stt = *a.state;
    const int shift = pass_shift(a.pass);
    const unsigned int binmask = (unsigned int)pass_bins(a.pass) - 1u;
    if (a.dstore) {      // tile t(I, J) of the row-major upper triangle (tile_row_start = 0, stride 1)
      float* blk = a.dstore + (size_t)(I * tiles_j - I * (I - 1) / 2 + (J - I)) * (TILE * TILE);
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        float o[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int64_t gi = i0 + ty * 4 + ii, gj = j0 + tx * 4 + jj;
          o[jj] = (gi < a.na && gj < a.nb && gi < gj) ? acc[ii][jj] : __int_as_float(0x7f800000);
        }
        *reinterpret_cast<float4*>(blk + (ty * 4 + ii) * TILE + tx * 4) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int64_t gi = i0 + ty * 4 + ii, gj = j0 + tx * 4 + jj;
        if (gi < a.na && gj < a.nb && gi < gj) {
          const float v = acc[ii][jj];
          if (v != v) { any_nan = true; continue; }
          const unsigned int bits = __float_as_uint(v);
          const unsigned int bin = (bits >> shift) & binmask;
          if ((bits & stt.mask) == stt.prefix[0]) atomicAdd(&hist_s[bin], 1u);
          if ((bits & stt.mask) == stt.prefix[1]) atomicAdd(&hist_s[NSVGD_SELECT_BINS + bin], 1u);
        }
      }
    continue;
  }

  const float h = ard ? 1.f : (a.h_dev ? *a.h_dev : a.h_host);       // ARD: the weights carry the bandwidths
  const float inv_h2 = 1.f / (h * h);
  const float nc = -NSVGD_HALF_LOG2E * inv_h2;

  if (MODE == MODE_KMAT) {
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int64_t gi = i0 + ty * 4 + ii;
      if (gi >= a.na) continue;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int64_t gj = j0 + tx * 4 + jj;
        if (gj < a.nb) a.K[gi * a.ldk + gj] = ex2_approx(acc[ii][jj] * nc);
      }
    }
    return;
  }

  if (MODE == MODE_VJP) {
    // K_ij, C_ij = k_ij c_ij with c_ij = g_i.v_ij, v_ij = s_j + (q_i - x_j)/h^2; d/dh of sum_ij g_i.(k_ij v_ij):
    // k r2/h^3 c - 2 k g_i.(q_i - x_j)/h^3
    double ldh = 0.0;
    const float inv_h3 = inv_h2 / h;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int64_t gi = i0 + ty * 4 + ii;
      if (gi >= a.na) continue;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int64_t gj = j0 + tx * 4 + jj;
        if (gj >= a.nb) continue;
        const float r2 = acc[ii][jj];
        const float k = ex2_approx(r2 * nc);
        const float c = fmaf(acc_sd[ii][jj], inv_h2, acc_ss[ii][jj]);
        a.K[gi * a.ldk + gj] = k;
        a.C[gi * a.ldk + gj] = k * c;
        ldh += (double)(k * inv_h3 * (r2 * c - 2.f * acc_sd[ii][jj]));
      }
    }
    if (a.partials_dh) {
      const double tot = block_sum(ldh, red_s);
      if (threadIdx.x == 0) a.partials_dh[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
    return;
  }

  // KSUM / KSD: weighted block sum -> partials[block]
  double local = 0.0, local_dh = 0.0;
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int64_t gi = i0 + ty * 4 + ii, gj = j0 + tx * 4 + jj;
      if (gi >= a.na || gj >= a.nb) continue;
      float wgt = 1.f;
      if (a.symmetric) wgt = gi < gj ? 2.f : (gi == gj ? a.diag_weight : 0.f);
      if (wgt == 0.f) continue;
      const float k = ex2_approx(acc[ii][jj] * nc);
      float v = k;
      if (MODE == MODE_KSD) {
        const float r2 = acc[ii][jj];
        // trace term: d/h^2 - r2/h^4, or sum_a 1/h_a^2 - sum_a diff_a^2/h_a^4 under an ARD bandwidth
        const float u = ard ? acc_ss[ii][jj] + acc_sd[ii][jj] + a.wsum - acc_q4[ii][jj]
                            : acc_ss[ii][jj] + acc_sd[ii][jj] * inv_h2 + (float)a.d * inv_h2 - r2 * inv_h2 * inv_h2;
        v = k * u;
        if (a.partials_dh) {
          // d/dh [k u]: dk/dh = k r2/h^3;  du/dh = -2 (s_x-s_y).(x-y)/h^3 - 2 d/h^3 + 4 r2/h^5
          const float inv_h3 = inv_h2 / h;
          const float du = inv_h3 * (-2.f * acc_sd[ii][jj] - 2.f * (float)a.d + 4.f * r2 * inv_h2);
          local_dh += (double)(wgt * k * (r2 * inv_h3 * u + du));
        }
      }
      local += (double)(wgt * v);
    }
  const double tot = block_sum(local, red_s);
  if (threadIdx.x == 0) a.partials[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
  if (MODE == MODE_KSD && a.partials_dh) {
    const double tdh = block_sum(local_dh, red_s);
    if (threadIdx.x == 0) a.partials_dh[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = tdh;
  }
  }  // J
  }  // rr
  if (MODE == MODE_HIST) {
    if (any_nan) atomicOr(const_cast<uint32_t*>(&a.state->nan_flag), 1u);
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * NSVGD_SELECT_BINS; t += 256) {
      const unsigned int v = hist_s[t];
      if (v) atomicAdd(&a.hist[t], (unsigned long long)v);
    }
  }
}

// deterministic single-block reduction of the per-CTA partials: out = [out +] scale * sum
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double* __restrict__ p, int64_t count,
                                                               double scale, int accumulate, double* out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t t = threadIdx.x; t < count; t += 1024) s += p[t];
  const double tot = block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.0) + scale * tot;
}

__global__ void zero_partials_kernel(double* p, int64_t count) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x)
    p[t] = 0.0;
}

static dim3 tile_grid(int64_t na, int64_t nb, int64_t start, int64_t stride) {
  const int64_t ti = ceil_div(na, TILE), tj = ceil_div(nb, TILE);
  const int64_t rows = start < ti ? ceil_div(ti - start, stride) : 0;
  return dim3((unsigned)tj, (unsigned)rows, 1);
}

int launch_kmat(cudaStream_t st, const float* Q, int64_t m, int64_t ldq, const float* X, int64_t n, int64_t ldx,
                int64_t d, const float* h_dev, float* K, int64_t ldk) {
  PairArgs a{};
  a.A = Q; a.na = m; a.lda = ldq; a.B = X; a.nb = n; a.ldb = ldx; a.d = d;
  a.tile_row_start = 0; a.tile_row_stride = 1; a.h_dev = h_dev; a.K = K; a.ldk = ldk;
  dim3 g = tile_grid(m, n, 0, 1);
  NSVGD_REQUIRE(g.y <= 65535, NSVGD_ERR_INVALID, "kmat: too many row tiles (%u)", g.y);
  pair_tile_kernel<MODE_KMAT><<<g, 256, 0, st>>>(a);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

int launch_ksum(cudaStream_t st, const float* A, int64_t na, int64_t lda, const float* B, int64_t nb, int64_t ldb,
                int64_t d, const float* h_dev, float h_host, int symmetric, double* out_dev, double scale,
                int accumulate, void* ws, size_t ws_bytes, bool dry, size_t* need) {
  dim3 g = tile_grid(na, nb, 0, 1);
  const int64_t nblocks = (int64_t)g.x * g.y;
  WsCarver w(ws, ws_bytes, dry);
  double* partials = w.take<double>((size_t)nblocks);
  if (need) *need = w.off;
  if (dry) return NSVGD_OK;
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "ksum: workspace too small (%zu < %zu)", ws_bytes, w.off);
  NSVGD_REQUIRE(g.y <= 65535, NSVGD_ERR_INVALID, "ksum: too many row tiles (%u)", g.y);
  PairArgs a{};
  a.A = A; a.na = na; a.lda = lda; a.B = B; a.nb = nb; a.ldb = ldb; a.d = d;
  a.symmetric = symmetric; a.tile_row_start = 0; a.tile_row_stride = 1;
  a.h_dev = h_dev; a.h_host = h_host; a.partials = partials; a.diag_weight = 0.f;
  if (symmetric) {  // skipped (J < I) tiles never write their partial
    zero_partials_kernel<<<(unsigned)std::min<int64_t>(ceil_div(nblocks, 256), 1024), 256, 0, st>>>(partials, nblocks);
    NSVGD_LAUNCH_CHECK();
  }
  pair_tile_kernel<MODE_KSUM><<<g, 256, 0, st>>>(a);
  NSVGD_LAUNCH_CHECK();
  reduce_partials_kernel<<<1, 1024, 0, st>>>(partials, nblocks, scale, accumulate, out_dev);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

// ------------------------------------------------------------------ median select
__global__ void select_init_kernel(SelectState* s, long long r_lo, long long r_hi, float log_np1) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    s->prefix[0] = s->prefix[1] = 0u;
    s->mask = 0u;
    s->nan_flag = 0u;
    s->is_zero[0] = r_lo < 0; s->is_zero[1] = r_hi < 0;
    s->rank[0] = r_lo < 0 ? 0ull : (unsigned long long)r_lo;
    s->rank[1] = r_hi < 0 ? 0ull : (unsigned long long)r_hi;
    s->value[0] = s->value[1] = 0.f;
    s->log_np1 = log_np1;
    s->done = 0;
  }
}

// literal mode: Z[a] = [x_a without column a (a < d-1) | x_a[:d-1], x_a[d-1]]   (n > d)
//               Z[a] = x_a[:n] without column a                                 (n <= d)
// ldz >= dz: columns [dz, ldz) are written as zeros (the large-d select reads float4 groups)
__global__ void literal_z_kernel(const float* __restrict__ X, int64_t n, int64_t d, int64_t ldx,
                                 float* __restrict__ Z, int64_t dz, int64_t ldz) {
  const int64_t total = n * ldz;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = t / ldz, c = t % ldz;
    if (c >= dz) { Z[t] = 0.f; continue; }
    int64_t src;
    if (n > d) {
      if (c == d - 1) src = d - 1;
      else src = (a < d - 1 && c >= a) ? c + 1 : c;
    } else {
      src = c >= a ? c + 1 : c;
    }
    Z[t] = X[a * ldx + src];
  }
}

__global__ void __launch_bounds__(1024) select_advance_kernel(SelectState* s, const unsigned long long* hist, int pass,
                                                              const uint32_t* run_flag) {
  if (run_flag && *run_flag == 0u) return;
  select_advance_body(s, hist, pass);
}

// row-sharded SIMT select: all-reduce of the histogram over peer memory fused into the decision kernel
__global__ void __launch_bounds__(1024) select_advance_p2p_kernel(const P2PComm c, SelectState* s, unsigned long long* hist,
                                                                  int pass, const uint32_t* run_flag) {
  if (run_flag && *run_flag == 0u) return;       // the flag is identical on every rank (it derives from all-reduced counts)
  if (!p2p_allreduce_cta<1>(c, hist, 2 * NSVGD_SELECT_BINS, false) && threadIdx.x == 0) s->nan_flag = 1u;   // late peer -> h = NaN
  select_advance_body(s, hist, pass);
}

__global__ void select_finish_kernel(const SelectState* s, float* h_dev) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float h;
    if (s->nan_flag) {
      h = __uint_as_float(0x7fc00000u);
    } else {
      const float med = __fmul_rn(__fadd_rn(s->value[0], s->value[1]), 0.5f);   // np.median midpoint rule
      h = __fsqrt_rn(__fdiv_rn(__fmul_rn(0.5f, med), s->log_np1));             // kernels.py:171
    }
    h_dev[0] = h;
  }
}

// histogram of an in-memory value buffer (sampled distances / candidate distances) for the same select
__global__ void __launch_bounds__(256) buffer_hist_kernel(const float* __restrict__ vals, const uint32_t* counts_dev,
                                                          int count_stride, int nseg, uint32_t seg_cap, uint32_t count_host,
                                                          const SelectState* state, int pass, unsigned long long* hist,
                                                          const uint32_t* run_flag) {
  if (run_flag && *run_flag == 0u) return;
  __shared__ unsigned int hist_s[2 * NSVGD_SELECT_BINS];
  for (int t = threadIdx.x; t < 2 * NSVGD_SELECT_BINS; t += 256) hist_s[t] = 0u;
  __syncthreads();
  const SelectState stt = *state;
  const int shift = pass_shift(pass);
  const unsigned int binmask = (unsigned int)pass_bins(pass) - 1u;
  bool any_nan = false;
  for (int sgm = 0; sgm < nseg; ++sgm) {
    const uint32_t count = counts_dev ? min(counts_dev[sgm * count_stride], seg_cap) : count_host;
    const float* v0 = vals + (size_t)sgm * seg_cap;
    for (uint32_t t = blockIdx.x * 256u + threadIdx.x; t < count; t += gridDim.x * 256u) {
      const float v = v0[t];
      if (v != v) { any_nan = true; continue; }
      const unsigned int bits = __float_as_uint(v);
      const unsigned int bin = (bits >> shift) & binmask;
      if ((bits & stt.mask) == stt.prefix[0]) atomicAdd(&hist_s[bin], 1u);
      if ((bits & stt.mask) == stt.prefix[1]) atomicAdd(&hist_s[NSVGD_SELECT_BINS + bin], 1u);
    }
  }
  if (any_nan) atomicOr(const_cast<uint32_t*>(&state->nan_flag), 1u);
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * NSVGD_SELECT_BINS; t += 256) {
    const unsigned int v = hist_s[t];
    if (v) atomicAdd(&hist[t], (unsigned long long)v);
  }
}

int launch_select_init(cudaStream_t st, SelectState* s, long long r_lo, long long r_hi, float log_np1) {
  select_init_kernel<<<1, 32, 0, st>>>(s, r_lo, r_hi, log_np1);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

int launch_buffer_hist(cudaStream_t st, const float* vals, const uint32_t* counts_dev, int count_stride, int nseg,
                       uint32_t seg_cap, uint32_t count_host, const SelectState* state, int pass, uint64_t* hist,
                       const uint32_t* run_flag) {
  NSVGD_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint64_t) * 2 * NSVGD_SELECT_BINS, st));
  buffer_hist_kernel<<<device_sm_count() * 4, 256, 0, st>>>(vals, counts_dev, count_stride, nseg, seg_cap, count_host, state,
                                                            pass, (unsigned long long*)hist, run_flag);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

int launch_select_advance(cudaStream_t st, SelectState* s, const uint64_t* hist, int pass, const uint32_t* run_flag) {
  select_advance_kernel<<<1, 64, 0, st>>>(s, (const unsigned long long*)hist, pass, run_flag);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

int launch_select_advance_p2p(cudaStream_t st, const P2PComm& c, SelectState* s, uint64_t* hist, int pass, const uint32_t* run_flag) {
  select_advance_p2p_kernel<<<1, 1024, 0, st>>>(c, s, (unsigned long long*)hist, pass, run_flag);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

int launch_pair_hist(cudaStream_t st, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last,
                     int64_t tile_row_start, int64_t tile_row_stride, const SelectState* state, int pass,
                     uint64_t* hist, const uint32_t* run_flag, float* dstore) {
  NSVGD_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint64_t) * 2 * NSVGD_SELECT_BINS, st));
  PairArgs a{};
  a.dstore = dstore;
  a.A = Z; a.na = n; a.lda = ldz; a.B = Z; a.nb = n; a.ldb = ldz; a.d = dz; a.w_last = w_last;
  a.symmetric = 1; a.tile_row_start = tile_row_start; a.tile_row_stride = tile_row_stride;
  a.state = state; a.pass = pass; a.hist = (unsigned long long*)hist; a.run_flag = run_flag;
  const int64_t tiles = ceil_div(n, TILE);
  const int64_t rows = tile_row_start < tiles ? ceil_div(tiles - tile_row_start, tile_row_stride) : 0;
  if (rows == 0) return NSVGD_OK;
  // persistent 2-D grid-stride walk: ~8 CTAs per SM, never more CTAs than tiles
  const int64_t target = (int64_t)device_sm_count() * 8;
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(tiles, 32));
  const unsigned gy = (unsigned)std::max<int64_t>(1, std::min<int64_t>(rows, ceil_div(target, gx)));
  pair_tile_kernel<MODE_HIST><<<dim3(gx, gy), 256, 0, st>>>(a);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

}  // namespace nsvgd

namespace nsvgd {
void median_ranks_host(int64_t n, int mode, long long* r_lo, long long* r_hi) {
  // SURVEY.md Appendix B; mirrors oracle/svgd_oracle.py::median_ranks
  const long long P = (long long)n * (n - 1) / 2;
  if (mode == NSVGD_MEDIAN_OFFDIAG) {
    *r_lo = (P - 1) / 2; *r_hi = P / 2;
    return;
  }
  const long long N = (long long)n * n;
  auto f = [&](long long r) -> long long { return r < n ? -1 : (r - n) / 2; };
  if (N % 2 == 0) { *r_lo = f(N / 2 - 1); *r_hi = f(N / 2); }
  else { *r_lo = *r_hi = f((N - 1) / 2); }
}
}  // namespace nsvgd

using namespace nsvgd;

// leading dimension of the literal-mode Z the single-GPU driver allocates: padded to a float4 multiple for the large-d select
static int64_t literal_ldz(int64_t dz) { return dz > 128 ? round_up(dz, 4) : dz; }

static int median_prepare_ld(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, int mode, float* Z, int64_t ldz_pad,
                             int64_t* dz_out, float* w_last_out, void* state_dev);

extern "C" int nsvgd_median_prepare(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, int mode,
                                    float* Z, int64_t* dz_out, float* w_last_out, void* state_dev) {
  return median_prepare_ld(stream, X, n, d, ldx, mode, Z, 0, dz_out, w_last_out, state_dev);
}

// ldz_pad = 0: Z is dense (ld = dz, the public entry point); else its leading dimension
static int median_prepare_ld(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, int mode, float* Z, int64_t ldz_pad,
                             int64_t* dz_out, float* w_last_out, void* state_dev) {
  cudaStream_t st = (cudaStream_t)stream;
  NSVGD_REQUIRE(X && state_dev && dz_out && w_last_out, NSVGD_ERR_INVALID, "median_prepare: null pointer");
  NSVGD_REQUIRE(n >= 2 && d >= 1 && ldx >= d, NSVGD_ERR_INVALID, "median_prepare: need n >= 2, d >= 1, ldx >= d (n=%lld d=%lld)", (long long)n, (long long)d);
  NSVGD_REQUIRE(mode >= 0 && mode <= 2, NSVGD_ERR_INVALID, "median_prepare: bad mode %d", mode);
  long long r_lo, r_hi;
  median_ranks_host(n, mode, &r_lo, &r_hi);
  const float log_np1 = (float)log((double)(n + 1));
  NSVGD_PROPAGATE(launch_select_init(st, (SelectState*)state_dev, r_lo, r_hi, log_np1));
  if (mode == NSVGD_MEDIAN_LITERAL) {
    NSVGD_REQUIRE(Z, NSVGD_ERR_INVALID, "median_prepare: literal mode needs a Z buffer");
    const int64_t dz = n > d ? d : n - 1;
    *dz_out = dz;
    *w_last_out = n > d ? (float)(n - d) : 1.f;
    const int64_t ldz = ldz_pad ? ldz_pad : dz;
    const int64_t total = n * ldz;
    literal_z_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 4096), 256, 0, st>>>(X, n, d, ldx, Z, dz, ldz);
    NSVGD_LAUNCH_CHECK();
  } else {
    *dz_out = d;
    *w_last_out = 1.f;
  }
  return NSVGD_OK;
}

extern "C" int nsvgd_median_hist_pass(void* stream, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last,
                                      int64_t tile_row_start, int64_t tile_row_stride, const void* state_dev, int pass,
                                      uint64_t* hist_dev) {
  NSVGD_REQUIRE(Z && state_dev && hist_dev, NSVGD_ERR_INVALID, "median_hist_pass: null pointer");
  NSVGD_REQUIRE(pass >= 0 && pass < NSVGD_SELECT_PASSES, NSVGD_ERR_INVALID, "median_hist_pass: bad pass %d", pass);
  NSVGD_REQUIRE(tile_row_stride >= 1 && tile_row_start >= 0, NSVGD_ERR_INVALID, "median_hist_pass: bad tile row partition");
  return launch_pair_hist((cudaStream_t)stream, Z, n, dz, ldz, w_last, tile_row_start, tile_row_stride,
                          (const SelectState*)state_dev, pass, hist_dev, nullptr, nullptr);
}

extern "C" int nsvgd_median_advance(void* stream, void* state_dev, const uint64_t* hist_dev, int pass) {
  NSVGD_REQUIRE(state_dev && hist_dev, NSVGD_ERR_INVALID, "median_advance: null pointer");
  NSVGD_REQUIRE(pass >= 0 && pass < NSVGD_SELECT_PASSES, NSVGD_ERR_INVALID, "median_advance: bad pass %d", pass);
  return launch_select_advance((cudaStream_t)stream, (SelectState*)state_dev, hist_dev, pass, nullptr);
}

extern "C" int nsvgd_p2p_median_advance(void* stream, int rank, int world, void* const* blocks, void* state_dev,
                                        uint64_t* hist_dev, int pass) {
  P2PComm c;
  NSVGD_REQUIRE(p2p_make_comm(&c, rank, world, blocks), NSVGD_ERR_INVALID, "p2p_median_advance: bad communicator");
  NSVGD_REQUIRE(state_dev && hist_dev, NSVGD_ERR_INVALID, "p2p_median_advance: null pointer");
  NSVGD_REQUIRE(pass >= 0 && pass < NSVGD_SELECT_PASSES, NSVGD_ERR_INVALID, "p2p_median_advance: bad pass %d", pass);
  return launch_select_advance_p2p((cudaStream_t)stream, c, (SelectState*)state_dev, hist_dev, pass, nullptr);
}

extern "C" int nsvgd_median_finish(void* stream, const void* state_dev, int64_t n, float* h_dev) {
  (void)n;
  cudaStream_t st = (cudaStream_t)stream;
  NSVGD_REQUIRE(state_dev && h_dev, NSVGD_ERR_INVALID, "median_finish: null pointer");
  select_finish_kernel<<<1, 32, 0, st>>>((const SelectState*)state_dev, h_dev);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

// Shapes the tensor-core filter does not take (dz > 128, or n < 8192) with enough coordinates that a distance costs more
// than a 4-byte store and two 4-byte loads: ONE pass of the exact SIMT pair engine computes every distance, histograms the
// first radix digit and keeps the values; the other two radix passes histogram the buffer (HBM-bound, < 1 % of a compute
// pass at d = 4096).  The values are the same fp32 numbers, so the result is bit-identical to the 3-pass select.
static int64_t median_store_values(int64_t n, int64_t dz) {
  if (median_fast_supported(n, dz) || median_large_supported(n, dz) || dz <= 32 || n < 256) return 0;
  const int64_t t = ceil_div(n, TILE);
  const int64_t vals = t * (t + 1) / 2 * (TILE * TILE);
  const char* env = getenv("NSVGD_MEDIAN_STORE");
  if (env && env[0] == '0') return 0;
  return vals <= (1ll << 31) ? vals : 0;
}

// workspace: [state | hist | Z (literal) | fast-path scratch (d <= 128: median_tc.cu, d > 128: median_large.cu) | stored distances]
static void median_ws_layout(int64_t n, int64_t d, int mode, WsCarver& w, void** state, uint64_t** hist, float** Z,
                             void** fast, size_t* fast_bytes, float** dstore = nullptr) {
  *state = w.take<char>(NSVGD_SELECT_STATE_BYTES);
  *hist = w.take<uint64_t>(2 * NSVGD_SELECT_BINS);
  *Z = nullptr;
  const int64_t dz = mode == NSVGD_MEDIAN_LITERAL ? (n > d ? d : n - 1) : d;
  if (mode == NSVGD_MEDIAN_LITERAL) *Z = w.take<float>((size_t)(n * literal_ldz(dz)));
  *fast = nullptr; *fast_bytes = 0;
  if (median_fast_supported(n, dz)) {
    *fast_bytes = median_fast_workspace(n, dz);
    *fast = w.take<char>(*fast_bytes, 1024);
  } else if (median_large_supported(n, dz)) {
    *fast_bytes = median_large_workspace(n, dz);
    *fast = w.take<char>(*fast_bytes, 1024);
  }
  const int64_t sv = median_store_values(n, dz);
  float* ds = sv ? w.take<float>((size_t)sv, 1024) : nullptr;
  if (dstore) *dstore = ds;
}

namespace nsvgd {
size_t median_workspace(int64_t n, int64_t d) {
  size_t best = 0;
  for (int mode = 0; mode < 3; ++mode) {
    WsCarver w(nullptr, 0, true);
    void* s; uint64_t* h; float* Z; void* f; size_t fb;
    median_ws_layout(n, d, mode, w, &s, &h, &Z, &f, &fb);
    best = std::max(best, w.off);
  }
  return best + 256;
}
}  // namespace nsvgd

extern "C" int nsvgd_median_h_hinted_f32(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, int mode, float* h_dev,
                                         void* hint_dev, void* ws, size_t ws_bytes);
extern "C" int nsvgd_median_h_f32(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, int mode,
                                  float* h_dev, void* ws, size_t ws_bytes) {
  return nsvgd_median_h_hinted_f32(stream, X, n, d, ldx, mode, h_dev, nullptr, ws, ws_bytes);
}

// hint_dev (may be null): 32 bytes of caller-owned, zero-initialised device memory that persists between calls on particle
// clouds that change slowly (an SVGD loop): the bracket of the tensor-core filtered select (d <= 128, n >= 8192) is then
// warm-started from the previous call's order statistics and the sampling + bracket-select phases are skipped.  The result is
// bit-identical with or without it (the select verifies the bracket with exact counts); shapes on other paths ignore it.
extern "C" int nsvgd_median_h_hinted_f32(void* stream, const float* X, int64_t n, int64_t d, int64_t ldx, int mode, float* h_dev,
                                         void* hint_dev, void* ws, size_t ws_bytes) {
  NSVGD_REQUIRE(X && h_dev && ws, NSVGD_ERR_INVALID, "median_h: null pointer");
  NSVGD_REQUIRE(n >= 2 && d >= 1 && ldx >= d, NSVGD_ERR_INVALID, "median_h: need n >= 2, d >= 1, ldx >= d (n=%lld d=%lld)", (long long)n, (long long)d);
  NSVGD_REQUIRE(mode >= 0 && mode <= 2, NSVGD_ERR_INVALID, "median_h: bad mode %d", mode);
  cudaStream_t st = (cudaStream_t)stream;
  WsCarver w(ws, ws_bytes);
  void* state; uint64_t* hist; float* Z; void* fast; size_t fast_bytes; float* dstore;
  median_ws_layout(n, d, mode, w, &state, &hist, &Z, &fast, &fast_bytes, &dstore);
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "median_h: workspace too small (%zu < %zu)", ws_bytes, w.off);
  int64_t dz; float w_last;
  prof_begin(3, st);
  prof_begin(12, st);
  const int64_t dz_lit = n > d ? d : n - 1;
  NSVGD_PROPAGATE(median_prepare_ld(stream, X, n, d, ldx, mode, Z, mode == NSVGD_MEDIAN_LITERAL ? literal_ldz(dz_lit) : 0, &dz, &w_last, state));
  prof_end(12, st);
  const float* P = mode == NSVGD_MEDIAN_LITERAL ? Z : X;
  const int64_t ldp = mode == NSVGD_MEDIAN_LITERAL ? literal_ldz(dz) : ldx;
  const char* env = getenv("NSVGD_MEDIAN_FAST");
  const bool want_fast = fast && !(env && env[0] == '0');
  if (want_fast) {
    // tensor-core filtered exact select; falls back on the device to the 3-pass SIMT select if its
    // bracket misses (probability ~1e-9) or a candidate list overflows (massive ties)
    long long r_lo, r_hi;
    median_ranks_host(n, mode, &r_lo, &r_hi);
    // float4 rows; columns [dz, ceil4(dz)) must be zero: our own Z is padded that way, a caller's array only if dz % 4 == 0
    const bool aligned = ldp % 4 == 0 && (((uintptr_t)P) & 15) == 0 && (dz % 4 == 0 || mode == NSVGD_MEDIAN_LITERAL);
    if (median_fast_supported(n, dz))
      NSVGD_PROPAGATE(median_fast(st, P, n, dz, ldp, w_last, r_lo, r_hi, 0, 1, (SelectState*)state, hist, fast, fast_bytes,
                                  (MedianHint*)hint_dev));
    else if (aligned)
      NSVGD_PROPAGATE(median_large(st, P, n, dz, ldp, w_last, r_lo, r_hi, (SelectState*)state, hist, fast, fast_bytes,
                                   (MedianHint*)hint_dev));
    else
      for (int pass = 0; pass < NSVGD_SELECT_PASSES; ++pass) {      // unaligned rows: the recomputing 3-pass select
        NSVGD_PROPAGATE(launch_pair_hist(st, P, n, dz, ldp, w_last, 0, 1, (const SelectState*)state, pass, hist, nullptr));
        NSVGD_PROPAGATE(launch_select_advance(st, (SelectState*)state, hist, pass, nullptr));
      }
  } else {
    const int64_t sv = dstore ? median_store_values(n, dz) : 0;
    for (int pass = 0; pass < NSVGD_SELECT_PASSES; ++pass) {
      if (sv && pass > 0)
        NSVGD_PROPAGATE(launch_buffer_hist(st, dstore, nullptr, 0, 1, 0, (uint32_t)sv, (const SelectState*)state, pass, hist, nullptr));
      else
        NSVGD_PROPAGATE(launch_pair_hist(st, P, n, dz, ldp, w_last, 0, 1, (const SelectState*)state, pass, hist, nullptr,
                                         sv ? dstore : nullptr));
      NSVGD_PROPAGATE(launch_select_advance(st, (SelectState*)state, hist, pass, nullptr));
    }
  }
  const int rc = nsvgd_median_finish(stream, state, n, h_dev);
  prof_end(3, st);
  return rc;
}

extern "C" int nsvgd_ksum_f32(void* stream, const float* A, int64_t na, int64_t lda, const float* B, int64_t nb,
                              int64_t ldb, int64_t d, float h, int symmetric, double* out_dev, void* ws, size_t ws_bytes) {
  NSVGD_REQUIRE(A && B && out_dev, NSVGD_ERR_INVALID, "ksum: null pointer");
  NSVGD_REQUIRE(na >= 1 && nb >= 1 && d >= 1 && lda >= d && ldb >= d, NSVGD_ERR_INVALID, "ksum: bad shape");
  NSVGD_REQUIRE(!symmetric || (A == B && na == nb), NSVGD_ERR_INVALID, "ksum: symmetric needs A == B");
  return launch_ksum((cudaStream_t)stream, A, na, lda, B, nb, ldb, d, nullptr, h, symmetric, out_dev,
                     symmetric ? 1.0 : 1.0, 0, ws, ws_bytes, false, nullptr);
}

static int ksd_impl(cudaStream_t st, const float* X, const float* Sx, int64_t n, int64_t ldx, const float* Y, const float* Sy,
                    int64_t m, int64_t ldy, int64_t d, const float* h_dev, int variant, double* out_dev, double* dh_dev, void* ws,
                    size_t ws_bytes, const float* wcoord = nullptr, float wsum = 0.f) {
  NSVGD_REQUIRE(X && Sx && (h_dev || wcoord) && out_dev, NSVGD_ERR_INVALID, "ksd: null pointer");
  NSVGD_REQUIRE(variant >= 0 && variant <= 2, NSVGD_ERR_INVALID, "ksd: bad variant %d", variant);
  NSVGD_REQUIRE(n >= 1 && d >= 1 && ldx >= d, NSVGD_ERR_INVALID, "ksd: bad shape");
  const bool rect = variant == NSVGD_KSD_RECT;
  if (rect) {
    NSVGD_REQUIRE(Y && Sy && m >= 1 && ldy >= d, NSVGD_ERR_INVALID, "ksd: rectangular variant needs Y, Sy");
  } else {
    NSVGD_REQUIRE(n >= 2 || variant == NSVGD_KSD_V, NSVGD_ERR_INVALID, "ksd: U-statistic needs n >= 2");
    Y = X; Sy = Sx; m = n; ldy = ldx;
  }
  dim3 g = tile_grid(n, m, 0, 1);
  const int64_t nblocks = (int64_t)g.x * g.y;
  WsCarver w(ws, ws_bytes);
  double* partials = w.take<double>((size_t)nblocks * (dh_dev ? 2 : 1));
  double* partials_dh = dh_dev ? partials + nblocks : nullptr;
  NSVGD_REQUIRE(ws && w.ok(), NSVGD_ERR_WORKSPACE, "ksd: workspace too small (%zu < %zu)", ws_bytes, w.off);
  NSVGD_REQUIRE(g.y <= 65535, NSVGD_ERR_INVALID, "ksd: too many row tiles");
  PairArgs a{};
  a.A = X; a.na = n; a.lda = ldx; a.B = Y; a.nb = m; a.ldb = ldy; a.SA = Sx; a.SB = Sy; a.d = d;
  a.symmetric = rect ? 0 : 1; a.tile_row_start = 0; a.tile_row_stride = 1; a.h_dev = h_dev;
  a.partials = partials; a.partials_dh = partials_dh; a.diag_weight = variant == NSVGD_KSD_V ? 1.f : 0.f;
  a.wcoord = wcoord; a.wsum = wsum;
  if (!rect) {
    const int64_t nz = nblocks * (dh_dev ? 2 : 1);
    zero_partials_kernel<<<(unsigned)std::min<int64_t>(ceil_div(nz, 256), 1024), 256, 0, st>>>(partials, nz);
    NSVGD_LAUNCH_CHECK();
  }
  pair_tile_kernel<MODE_KSD><<<g, 256, 0, st>>>(a);
  NSVGD_LAUNCH_CHECK();
  double scale;
  if (variant == NSVGD_KSD_RECT) scale = 1.0 / ((double)n * (double)m);
  else if (variant == NSVGD_KSD_U) scale = 1.0 / ((double)n * (double)(n - 1));
  else scale = 1.0 / ((double)n * (double)n);
  reduce_partials_kernel<<<1, 1024, 0, st>>>(partials, nblocks, scale, 0, out_dev);
  NSVGD_LAUNCH_CHECK();
  if (dh_dev) {
    reduce_partials_kernel<<<1, 1024, 0, st>>>(partials_dh, nblocks, scale, 0, dh_dev);
    NSVGD_LAUNCH_CHECK();
  }
  return NSVGD_OK;
}

// ------------------------------------------------------------------ KSD (stein.py:273-349)
extern "C" int nsvgd_ksd_f32(void* stream, const float* X, const float* Sx, int64_t n, int64_t ldx, const float* Y,
                             const float* Sy, int64_t m, int64_t ldy, int64_t d, const float* h_dev, int variant,
                             double* out_dev, void* ws, size_t ws_bytes) {
  return ksd_impl((cudaStream_t)stream, X, Sx, n, ldx, Y, Sy, m, ldy, d, h_dev, variant, out_dev, nullptr, ws, ws_bytes);
}

// KSD and its derivative w.r.t. the bandwidth in one pass (the backward rule of the differentiable KSD)
extern "C" int nsvgd_ksd_dh_f32(void* stream, const float* X, const float* Sx, int64_t n, int64_t ldx, const float* Y,
                                const float* Sy, int64_t m, int64_t ldy, int64_t d, const float* h_dev, int variant,
                                double* out_dev, double* dh_dev, void* ws, size_t ws_bytes) {
  NSVGD_REQUIRE(dh_dev, NSVGD_ERR_INVALID, "ksd_dh: null dh pointer");
  return ksd_impl((cudaStream_t)stream, X, Sx, n, ldx, Y, Sy, m, ldy, d, h_dev, variant, out_dev, dh_dev, ws, ws_bytes);
}

// KSD under a per-coordinate (ARD) bandwidth h_a (kernels.py:52-65 with a vector bandwidth):
// u_p = k [ s_x.s_y + sum_a (s_x-s_y)_a (x-y)_a / h_a^2 + sum_a 1/h_a^2 - sum_a (x-y)_a^2 / h_a^4 ],  k = exp(-sum_a (x-y)_a^2 / (2 h_a^2)).
// inv_h2: d weights 1/h_a^2 on the device; inv_h2_sum their sum (host).
extern "C" int nsvgd_ksd_ard_f32(void* stream, const float* X, const float* Sx, int64_t n, int64_t ldx, const float* Y,
                                 const float* Sy, int64_t m, int64_t ldy, int64_t d, const float* inv_h2, float inv_h2_sum,
                                 int variant, double* out_dev, void* ws, size_t ws_bytes) {
  NSVGD_REQUIRE(inv_h2, NSVGD_ERR_INVALID, "ksd_ard: null weights");
  return ksd_impl((cudaStream_t)stream, X, Sx, n, ldx, Y, Sy, m, ldy, d, nullptr, variant, out_dev, nullptr, ws, ws_bytes, inv_h2,
                  inv_h2_sum);
}

// K, C and the bandwidth-derivative partials of the phi* backward rule (vjp.cu)
namespace nsvgd {
int launch_pair_vjp(cudaStream_t st, const float* Q, const float* G, int64_t m, int64_t ldq, const float* X, const float* S,
                    int64_t n, int64_t ldx, int64_t d, const float* h_dev, float* K, float* C, int64_t ldk, double* partials_dh) {
  PairArgs a{};
  a.A = Q; a.SA = G; a.na = m; a.lda = ldq; a.B = X; a.SB = S; a.nb = n; a.ldb = ldx; a.d = d;
  a.tile_row_start = 0; a.tile_row_stride = 1; a.h_dev = h_dev; a.K = K; a.C = C; a.ldk = ldk; a.partials_dh = partials_dh;
  dim3 g = tile_grid(m, n, 0, 1);
  NSVGD_REQUIRE(g.y <= 65535, NSVGD_ERR_INVALID, "phi_vjp: too many row tiles (%u)", g.y);
  pair_tile_kernel<MODE_VJP><<<g, 256, 0, st>>>(a);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}
}  // namespace nsvgd

// ------------------------------------------------------------------ MMD (metrics.py:88-127)
namespace nsvgd {
__global__ void defunnelize_kernel(const float* __restrict__ Z, int64_t n, int64_t d, int64_t ldz, float* __restrict__ out) {
  // kernels.py:112-116: [x * exp(-y/2), y/3]
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * d; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / d, c = t % d;
    const float y = Z[i * ldz + d - 1];
    out[t] = c == d - 1 ? y / 3.f : Z[i * ldz + c] * expf(-0.5f * y);
  }
}
}  // namespace nsvgd

extern "C" int nsvgd_mmd_f32(void* stream, const float* X, int64_t n, int64_t ldx, const float* Y, int64_t m,
                             int64_t ldy, int64_t d, float h, int funnel, double* out_dev, void* ws, size_t ws_bytes) {
  cudaStream_t st = (cudaStream_t)stream;
  NSVGD_REQUIRE(X && Y && out_dev && ws, NSVGD_ERR_INVALID, "mmd: null pointer");
  NSVGD_REQUIRE(n >= 2 && m >= 2 && d >= 1 && ldx >= d && ldy >= d, NSVGD_ERR_INVALID, "mmd: need n, m >= 2");
  NSVGD_REQUIRE(h > 0.f, NSVGD_ERR_INVALID, "mmd: bandwidth must be positive");
  WsCarver w(ws, ws_bytes);
  const float* A = X; const float* B = Y; int64_t lda = ldx, ldb = ldy;
  if (funnel) {
    float* Xf = w.take<float>((size_t)(n * d));
    float* Yf = w.take<float>((size_t)(m * d));
    NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "mmd: workspace too small");
    defunnelize_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n * d, 256), 2048), 256, 0, st>>>(X, n, d, ldx, Xf);
    NSVGD_LAUNCH_CHECK();
    defunnelize_kernel<<<(unsigned)std::min<int64_t>(ceil_div(m * d, 256), 2048), 256, 0, st>>>(Y, m, d, ldy, Yf);
    NSVGD_LAUNCH_CHECK();
    A = Xf; B = Yf; lda = d; ldb = d;
  }
  w.off = (w.off + 255) / 256 * 256;
  char* sub = (char*)ws + w.off;
  NSVGD_REQUIRE(w.off <= ws_bytes, NSVGD_ERR_WORKSPACE, "mmd: workspace too small");
  const size_t sub_bytes = ws_bytes - w.off;
  // mean offdiag Kxx = 2 sum_{i<j} k / (n(n-1))   (k(x,x) = 1 is excluded analytically)
  NSVGD_PROPAGATE(launch_ksum(st, A, n, lda, A, n, lda, d, nullptr, h, 1, out_dev, 1.0 / ((double)n * (n - 1)), 0, sub, sub_bytes, false, nullptr));
  NSVGD_PROPAGATE(launch_ksum(st, B, m, ldb, B, m, ldb, d, nullptr, h, 1, out_dev, 1.0 / ((double)m * (m - 1)), 1, sub, sub_bytes, false, nullptr));
  NSVGD_PROPAGATE(launch_ksum(st, A, n, lda, B, m, ldb, d, nullptr, h, 0, out_dev, -2.0 / ((double)n * m), 1, sub, sub_bytes, false, nullptr));
  return NSVGD_OK;
}
