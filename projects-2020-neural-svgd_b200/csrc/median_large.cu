// median_large.cu — the exact median-heuristic order statistics for d > 128 with a tensor-core filter (single GPU)
// [kernels.py:160-172, utils.py:178-197, 653-658; BASELINE configs[4]: n = 16384, d = 4096].
// Same contract as median_tc.cu (d <= 128): the tensor cores only DISCARD pairs that provably lie outside a bracket; every value
// that can influence the result is recomputed in the defined-order fp32 arithmetic, so the bandwidth is bit-identical to the
// SIMT select and to the oracle.  What differs at large d:
//   sample  : 2^18 row-stratified pairs (row i against pseudo-random partners; a warp moves the rows through shared memory one
//             128-byte column slice at a time and lane p adds pair p's terms in the defined order)
//   filter  : the persistent CTA-pair GEMM kernel of phi_large.cu over the symmetric tiles with a CLASSIFY epilogue.  Radius
//             eps = kg N2 + kd (d~ + kg N2), N2 = |z_i|^2 + |z_j|^2:  kg = 2^-15 + 1.5 T 2^-22 covers the bf16 2-piece products
//             (3 2^-18), the norms, the centring and T = 3 ceil(d/16) accumulating MMAs at <= 4 ulp of the largest partial sum
//             each (Cauchy-Schwarz: every partial sum <= |z_i||z_j| <= N2 / 2), kd = (d + 4) 2^-24 the sequential rounding of
//             the defined-order value itself (all terms non-negative: relative to the distance, not to the norms)
//   exact   : the uncertain pairs (~3 % at cfg5) are listed PER 128 x 256 TILE as 16-bit (row, column); one CTA per 256 x 256
//             tile streams the tile's 512 rows through shared memory in column slices and every thread adds up to 4 listed
//             pairs in the defined order -- 8 MB of rows per ~1900 pairs instead of 32 KB per pair
//   resolve : as in median_tc.cu (pack, candidate select, conditional SIMT fallback when the bracket missed or a list overflowed)
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "select.cuh"
#include "tc_common.cuh"

namespace nsvgd {

using namespace tc;

constexpr uint32_t ML_SAMPLES = 1u << 18;
constexpr uint32_t ML_TCAP = 4096;             // listed pairs per 128 x 256 tile (12.5 %)
constexpr int ML_TR = 128, ML_TC = 256;        // the CTA tile of the filter
constexpr int ML_XT = 512;                     // threads of the exact-tile kernel
constexpr int ML_XP = 4;                       // pairs per thread and round
constexpr int ML_XSTRIDE = 36;                 // floats per row slice in shared memory (32 + 4: 16-byte aligned rows)
constexpr int ML_XROWS_BYTES = 512 * ML_XSTRIDE * (int)sizeof(float);
constexpr int ML_XSMEM = ML_XROWS_BYTES + 2 * (int)ML_TCAP * (int)sizeof(unsigned short);      // row slices + the tile's row-sorted pair list

bool median_large_supported(int64_t n, int64_t dz) {
  const char* env = getenv("NSVGD_MEDIAN_LARGE");
  if (env && env[0] == '0') return false;
  return dz > 128 && n >= 2048 && n <= (1ll << 17);     // per-tile lists: 8 KB per 128 x 256 tile
  // (rows must be 16-byte aligned with a leading dimension that is a multiple of 4 and ZERO columns in [dz, ceil4(dz)): true for
  //  dz % 4 == 0, and for the literal-mode Z the driver pads itself -- e.g. the reference's BNN dimension 4594)
}

struct MlWs {
  FastMedState* fs; SelectState* bracket; SelectState* cand_state;
  float* mu; double* mpart; __nv_bfloat16* Zb0; __nv_bfloat16* Zb1; float* zn; float* Ds;
  unsigned short* tlist; uint32_t* tcount; float* cand;
  char* zero_block; size_t zero_bytes;
  uint32_t* err_flag; uint32_t* list_counts; uint32_t* cand_counts; FusedSelScratch* sel;
  uint32_t cand_cap; int64_t n_pad, ldk; int64_t tiles;
};

static size_t ml_layout(int64_t n, int64_t dz, WsCarver& w, MlWs* o) {
  MlWs f{};
  f.n_pad = round_up(n, ML_TC);
  f.ldk = round_up(dz, 8);
  f.tiles = (f.n_pad / ML_TR) * (f.n_pad / ML_TC);
  const double P = 0.5 * (double)n * (double)(n - 1);
  f.cand_cap = (uint32_t)(std::min<double>(P / 24.0 + (double)(1 << 22), 4.0e9) / MF_NSEG);
  f.fs = w.take<FastMedState>(1);
  f.bracket = (SelectState*)w.take<char>(NSVGD_SELECT_STATE_BYTES);
  f.cand_state = (SelectState*)w.take<char>(NSVGD_SELECT_STATE_BYTES);
  f.mu = w.take<float>((size_t)dz);
  f.mpart = w.take<double>((size_t)colmean_partials(n, dz));
  f.Zb0 = w.take<__nv_bfloat16>((size_t)(f.n_pad * f.ldk), 1024);
  f.Zb1 = w.take<__nv_bfloat16>((size_t)(f.n_pad * f.ldk), 1024);
  f.zn = w.take<float>((size_t)f.n_pad);
  f.Ds = w.take<float>(ML_SAMPLES);
  f.tlist = w.take<unsigned short>((size_t)f.tiles * ML_TCAP);
  f.tcount = w.take<uint32_t>((size_t)f.tiles);
  f.cand = w.take<float>((size_t)f.cand_cap * MF_NSEG);
  w.off = (w.off + 1023) / 1024 * 1024;
  const size_t z0 = w.off;
  f.err_flag = w.take<uint32_t>(64);        // [0] error / overflow flag; bytes 64.. = the pack summary
  f.list_counts = w.take<uint32_t>(MF_NSEG * MF_CSTRIDE);      // stays zero: the per-tile lists report overflow through err_flag
  f.cand_counts = w.take<uint32_t>(MF_NSEG * MF_CSTRIDE);
  f.sel = w.take<FusedSelScratch>(2);
  f.zero_block = w.dry ? nullptr : w.base + z0;
  f.zero_bytes = w.off - z0;
  if (o) *o = f;
  return w.off;
}

size_t median_large_workspace(int64_t n, int64_t dz) {
  WsCarver w(nullptr, 0, true);
  return ml_layout(n, dz, w, nullptr) + 1024;
}

// one warp per row: z = (p - mu) with the last coordinate scaled by sqrt(w) AFTER centring, two bf16 pieces, |z|^2
// (rows in [n, rows_pad): zero pieces, |z|^2 = 1e30 -- never a valid pair)
__global__ void __launch_bounds__(256) ml_prep_kernel(const float* __restrict__ P, int64_t n, int64_t rows_pad, int64_t d, int64_t ldp,
                                                      int64_t ldk, const float* __restrict__ mu, float last_scale,
                                                      __nv_bfloat16* __restrict__ Zb0, __nv_bfloat16* __restrict__ Zb1,
                                                      float* __restrict__ zn) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows_pad) return;
  const bool rv = row < n;
  float acc = 0.f;
  for (int64_t c0 = 2 * lane; c0 < ldk; c0 += 64) {
    float v2[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int64_t c = c0 + e;
      if (rv && c < d) {
        float v = P[row * ldp + c] - mu[c];
        if (c == d - 1) v *= last_scale;
        acc = fmaf(v, v, acc);
        v2[e] = v;
      }
    }
    const uint32_t p0 = pack_bf16x2(v2[0], v2[1]);
    const uint32_t p1 = pack_bf16x2(v2[0] - bf16lo_to_float(p0), v2[1] - bf16hi_to_float(p0));
    *reinterpret_cast<uint32_t*>(&Zb0[row * ldk + c0]) = p0;
    *reinterpret_cast<uint32_t*>(&Zb1[row * ldk + c0]) = p1;
  }
  acc = warp_sum(acc);
  if (lane == 0) zn[row] = rv ? acc : 1e30f;
}

__device__ __forceinline__ uint64_t ml_mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// Defined-order distances of 32 pairs by one warp (the slice-cooperative scheme of exact_list_coop_kernel, median_tc.cu): 8 lanes
// load one 128-byte row slice, 4 rows per instruction; the loading lane holds both rows of its pair for its 4 coordinates and
// parks the squared differences (the last coordinate already weighted) in the warp's tile; lane p then adds pair p's terms in
// order.  Rows 16-byte aligned, ld % 4 == 0, columns [d, ceil4(d)) zero: the weighted coordinate d - 1 sits at position
// (d - 1) & 3 of its float4 and the zero padding behind it adds exact +0.
__device__ __forceinline__ float ml_warp_exact32(const float* __restrict__ P, int64_t ldp, int d, float w_last, uint32_t i, uint32_t j,
                                                 float* tA, int lane) {
  const int r = lane >> 3, q = lane & 7;
  uint32_t ia[8], jb[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    ia[u] = __shfl_sync(0xffffffffu, i, 4 * u + r);
    jb[u] = __shfl_sync(0xffffffffu, j, 4 * u + r);
  }
  const int nslices = (d + 31) >> 5;
  const int last_group = (d - 1) >> 2, last_pos = (d - 1) & 3;
  const int d4 = (d + 3) & ~3;
  float acc = 0.f;
  for (int sl = 0; sl < nslices; ++sl) {
    const int col = 32 * sl + 4 * q;
    const bool inb = col < d;
    const bool wgt = (col >> 2) == last_group;
    float4 va[8], vb[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      va[u] = inb ? __ldg(reinterpret_cast<const float4*>(P + (int64_t)ia[u] * ldp + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
      vb[u] = inb ? __ldg(reinterpret_cast<const float4*>(P + (int64_t)jb[u] * ldp + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float d0 = va[u].x - vb[u].x, d1 = va[u].y - vb[u].y, d2 = va[u].z - vb[u].z, d3 = va[u].w - vb[u].w;
      float t0 = __fmul_rn(d0, d0), t1 = __fmul_rn(d1, d1), t2 = __fmul_rn(d2, d2), t3 = __fmul_rn(d3, d3);
      if (wgt) {
        if (last_pos == 0) t0 = __fmul_rn(w_last, t0);
        else if (last_pos == 1) t1 = __fmul_rn(w_last, t1);
        else if (last_pos == 2) t2 = __fmul_rn(w_last, t2);
        else t3 = __fmul_rn(w_last, t3);
      }
      *reinterpret_cast<float4*>(tA + (4 * u + r) * ML_XSTRIDE + 4 * q) = make_float4(t0, t1, t2, t3);
    }
    __syncwarp();
    const int kmax = min(8, (d4 - 32 * sl) >> 2);
    for (int k = 0; k < kmax; ++k) {
      const float4 t4 = *reinterpret_cast<const float4*>(tA + lane * ML_XSTRIDE + 4 * k);
      acc = __fadd_rn(acc, t4.x);
      acc = __fadd_rn(acc, t4.y);
      acc = __fadd_rn(acc, t4.z);
      acc = __fadd_rn(acc, t4.w);
    }
  }
  return acc;
}

// sample s = 32 g + p: row i_g = floor(g n / G) (G = samples / 32 groups: the rows are visited evenly) against a pseudo-random
// partner j != i_g; deterministic in (n, s)
__global__ void __launch_bounds__(128) ml_sample_kernel(const float* __restrict__ P, int64_t n, int d, int64_t ldp, float w_last,
                                                        float* __restrict__ Ds, const uint32_t* skip_flag) {
  if (skip_flag && *skip_flag) return;          // the bracket comes from the caller's hint this time (MedianHint, select.cuh)
  __shared__ __align__(16) float tiles[4][32 * ML_XSTRIDE];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t groups = ML_SAMPLES / 32u;
  for (uint32_t g = blockIdx.x * 4u + w; g < groups; g += gridDim.x * 4u) {
    const uint32_t i = (uint32_t)(((uint64_t)g * (uint64_t)n) / groups);
    uint32_t j = (uint32_t)(ml_mix(0xC0FFEEull ^ ((uint64_t)g * 32u + lane)) % (uint64_t)(n - 1));
    if (j >= i) ++j;
    const float dv = ml_warp_exact32(P, ldp, d, w_last, i, j, tiles[w], lane);
    Ds[g * 32u + lane] = dv;
  }
}

// one CTA per 256 x 256 tile on or above the diagonal: rows [256 I, +256) and [256 J, +256) stream through shared memory in
// 32-coordinate slices; thread t adds the listed pairs t, t + 512, ... (4 per round) in the defined order
__global__ void __launch_bounds__(ML_XT, 2) ml_exact_tiles_kernel(const float* __restrict__ P, int64_t n, int d, int64_t ldp, float w_last,
                                                                  FastMedState* fs, const unsigned short* __restrict__ tlist,
                                                                  const uint32_t* __restrict__ tcount, int tiles_n,
                                                                  float* __restrict__ cand, uint32_t* cand_counts, uint32_t cand_cap,
                                                                  SelectState* main_state) {
  extern __shared__ __align__(16) float xs[];        // [512 rows][ML_XSTRIDE] | sorted pair list
  __shared__ unsigned int sh_below;
  __shared__ uint32_t row_hist[257];
  const int I = blockIdx.y, J = blockIdx.x;
  if (J < I) return;
  const uint32_t c0 = tcount[(size_t)(2 * I) * tiles_n + J], c1 = tcount[(size_t)(2 * I + 1) * tiles_n + J];
  const uint32_t total = c0 + c1;
  if (total == 0u) return;
  if (threadIdx.x == 0) sh_below = 0u;
  __syncthreads();
  const unsigned short* l0 = tlist + ((size_t)(2 * I) * tiles_n + J) * ML_TCAP;
  const unsigned short* l1 = tlist + ((size_t)(2 * I + 1) * tiles_n + J) * ML_TCAP;
  const float Lv = fs->L, Uv = fs->U;
  const int lane = threadIdx.x & 31;
  const int seg = (int)((blockIdx.y * gridDim.x + blockIdx.x) % MF_NSEG);
  float* seg_cand = cand + (size_t)seg * cand_cap;
  const int nslices = (d + 31) >> 5;
  const int last_group = (d - 1) >> 2, last_pos = (d - 1) & 3;  // the weighted coordinate; the zero padding behind it adds exact +0
  const int d4 = (d + 3) & ~3;
  const int lrow = threadIdx.x >> 3, lq = threadIdx.x & 7;      // loads: 8 threads per row slice, 64 rows per pass
  // The filter appends the pairs of a tile in arbitrary order.  Counting sort by row (key = row << 8 | column, row in [0, 256)):
  // consecutive lanes then hold pairs of the SAME row, whose slice reads broadcast -- the kernel is bound by shared-memory
  // wavefronts (ncu: 80 % of peak, 41 % bank conflicts with random row pairs).  Measured at cfg5: 5.94 -> 5.33 ms.  Two
  // bank-aware orders of the column side (round-robin over the classes c mod 8 inside a row; 8 class queues interleaved slot by
  // slot) were measured too and did not help (5.7 / 5.9 ms): kept simple.
  unsigned short* sorted = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(xs) + ML_XROWS_BYTES);
  for (int t = threadIdx.x; t < 257; t += ML_XT) row_hist[t] = 0u;
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < total; t += ML_XT) {
    const unsigned short e = t < c0 ? l0[t] : l1[t - c0];
    atomicAdd(&row_hist[(t < c0 ? 0 : 128) + (e >> 8) + 1], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 32) {                 // exclusive prefix over the 256 rows: 8 per lane + a warp scan
    uint32_t loc[8], sum = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) { loc[u] = row_hist[1 + threadIdx.x * 8 + u]; sum += loc[u]; }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    uint32_t run = inc - sum;
#pragma unroll
    for (int u = 0; u < 8; ++u) { row_hist[1 + threadIdx.x * 8 + u] = run; run += loc[u]; }
  }
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < total; t += ML_XT) {
    const unsigned short e = t < c0 ? l0[t] : l1[t - c0];
    const uint32_t row = (t < c0 ? 0u : 128u) + (uint32_t)(e >> 8);
    const uint32_t pos = atomicAdd(&row_hist[1 + row], 1u);
    sorted[pos] = (unsigned short)((row << 8) | (e & 255u));
  }
  __syncthreads();
  const uint32_t nwork = total;
  uint32_t below = 0;
  bool any_nan = false;
  for (uint32_t base = 0; base < nwork; base += ML_XT * ML_XP) {
    int ra[ML_XP], rb[ML_XP];
    float acc[ML_XP];
    bool live[ML_XP];
#pragma unroll
    for (int k = 0; k < ML_XP; ++k) {
      const uint32_t t = base + threadIdx.x + ML_XT * k;
      live[k] = t < nwork;
      const unsigned short e = sorted[min(t, nwork - 1u)];                    // idle slots shadow the last pair
      ra[k] = e >> 8;
      rb[k] = 256 + (e & 255);
      acc[k] = 0.f;
    }
    for (int sl = 0; sl < nslices; ++sl) {
      const int col = 32 * sl + 4 * lq;
      __syncthreads();                                          // the previous slice has been consumed
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int lr = lrow + 64 * u;
        const int64_t grow = lr < 256 ? (int64_t)I * 256 + lr : (int64_t)J * 256 + (lr - 256);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grow < n && col < d) v = __ldg(reinterpret_cast<const float4*>(P + grow * ldp + col));
        *reinterpret_cast<float4*>(xs + lr * ML_XSTRIDE + 4 * lq) = v;
      }
      __syncthreads();
      const int kmax = min(8, (d4 - 32 * sl) >> 2);
#pragma unroll
      for (int k = 0; k < ML_XP; ++k) {
        const float* a = xs + ra[k] * ML_XSTRIDE;
        const float* b = xs + rb[k] * ML_XSTRIDE;
        float s = acc[k];
        for (int g = 0; g < kmax; ++g) {
          const float4 av = *reinterpret_cast<const float4*>(a + 4 * g), bv = *reinterpret_cast<const float4*>(b + 4 * g);
          const float d0 = av.x - bv.x, d1 = av.y - bv.y, d2 = av.z - bv.z, d3 = av.w - bv.w;
          float t0 = __fmul_rn(d0, d0), t1 = __fmul_rn(d1, d1), t2 = __fmul_rn(d2, d2), t3 = __fmul_rn(d3, d3);
          if (8 * sl + g == last_group) {
            if (last_pos == 0) t0 = __fmul_rn(w_last, t0);
            else if (last_pos == 1) t1 = __fmul_rn(w_last, t1);
            else if (last_pos == 2) t2 = __fmul_rn(w_last, t2);
            else t3 = __fmul_rn(w_last, t3);
          }
          s = __fadd_rn(s, t0);
          s = __fadd_rn(s, t1);
          s = __fadd_rn(s, t2);
          s = __fadd_rn(s, t3);
        }
        acc[k] = s;
      }
    }
#pragma unroll
    for (int k = 0; k < ML_XP; ++k) {
      const uint32_t t = base + threadIdx.x + ML_XT * k;
      const float dv = acc[k];
      bool is_cand = false;
      if (live[k]) {
        if (dv != dv) any_nan = true;
        else if (dv < Lv) ++below;
        else if (dv <= Uv) is_cand = true;
      }
      const uint32_t ball = __ballot_sync(0xffffffffu, is_cand);
      if (ball) {
        uint32_t pbase = 0;
        if (lane == 0) pbase = atomicAdd(&cand_counts[seg * MF_CSTRIDE], (uint32_t)__popc(ball));
        pbase = __shfl_sync(0xffffffffu, pbase, 0);
        if (is_cand) {
          const uint32_t pos = pbase + __popc(ball & ((1u << lane) - 1u));
          if (pos < cand_cap) seg_cand[pos] = dv;
        }
      }
    }
  }
  if (any_nan) atomicOr(&main_state->nan_flag, 1u);
  if (below) atomicAdd(&sh_below, below);
  __syncthreads();
  if (threadIdx.x == 0 && sh_below) atomicAdd(&fs->below, (unsigned long long)sh_below);
}

int median_large(cudaStream_t st, const float* P, int64_t n, int64_t dz, int64_t ldp, float w_last, long long r_lo, long long r_hi,
                 SelectState* main_state, uint64_t* hist, void* ws, size_t ws_bytes, MedianHint* hint) {
  NSVGD_REQUIRE(median_large_supported(n, dz), NSVGD_ERR_UNSUPPORTED, "median_large: unsupported shape");
  NSVGD_REQUIRE(ldp % 4 == 0 && (((uintptr_t)P) & 15) == 0, NSVGD_ERR_UNSUPPORTED, "median_large: rows must be 16-byte aligned");
  NSVGD_REQUIRE(r_lo >= 0 && r_hi >= r_lo, NSVGD_ERR_UNSUPPORTED, "median_large: needs genuine pair ranks");
  WsCarver w(ws, ws_bytes);
  MlWs f;
  ml_layout(n, dz, w, &f);
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "median_large: workspace too small (%zu < %zu)", ws_bytes, w.off);
  const int sms = device_sm_count();
  long long* red = (long long*)(f.err_flag) + 8;

  // --- prep + sample
  prof_begin(8, st);
  NSVGD_PROPAGATE(launch_colmean(st, P, n, dz, ldp, f.mpart, f.mu));
  NSVGD_CUDA(cudaMemsetAsync(f.zero_block, 0, f.zero_bytes, st));
  uint32_t* skip_flag = f.err_flag + 2;          // inside the zero block: 0 unless a valid hint arms the warm-started bracket
  const uint32_t hint_key = median_hint_key(n, dz, r_lo, w_last);
  if (hint) NSVGD_PROPAGATE(launch_hint_apply(st, hint, f.fs, skip_flag, hint_key));
  ml_prep_kernel<<<(unsigned)ceil_div(f.n_pad * 32, 256), 256, 0, st>>>(P, n, f.n_pad, dz, ldp, f.ldk, f.mu, sqrtf(w_last), f.Zb0, f.Zb1, f.zn);
  NSVGD_LAUNCH_CHECK();
  prof_end(8, st);
  prof_begin(4, st);
  ml_sample_kernel<<<(unsigned)std::min<int64_t>(sms * 8, ML_SAMPLES / 128), 128, 0, st>>>(P, n, (int)dz, ldp, w_last, f.Ds, skip_flag);
  NSVGD_LAUNCH_CHECK();
  prof_end(4, st);

  // --- bracket [L, U]: 5 sigma of the binomial rank error + slack (a miss only costs the device-side fallback)
  const double Ppairs = 0.5 * (double)n * (double)(n - 1);
  const double ns = (double)ML_SAMPLES;
  const double p_lo = ((double)r_lo + 0.5) / Ppairs, p_hi = ((double)r_hi + 0.5) / Ppairs;
  const double sig_lo = sqrt(ns * p_lo * (1.0 - p_lo)), sig_hi = sqrt(ns * p_hi * (1.0 - p_hi));
  long long a = (long long)floor(p_lo * ns - 5.0 * sig_lo - 16.0);
  long long b = (long long)ceil(p_hi * ns + 5.0 * sig_hi + 16.0);
  const int l_is_zero = a < 0, u_is_inf = b >= (long long)ML_SAMPLES;
  if (a < 0) a = 0;
  if (b >= (long long)ML_SAMPLES) b = (long long)ML_SAMPLES - 1;
  prof_begin(9, st);
  NSVGD_PROPAGATE(launch_select_init(st, f.bracket, a, b, 1.f));
  NSVGD_PROPAGATE(launch_fused_select(st, f.Ds, nullptr, 0, 1, 0, 0u, ML_SAMPLES, f.bracket, &f.sel[0], nullptr, 1, f.fs, nullptr,
                                      l_is_zero, u_is_inf, skip_flag));
  prof_end(9, st);

  // --- filter: symmetric Gram + classify
  const int T = 3 * (int)ceil_div(dz, 16);
  // NSVGD_MEDIAN_LARGE_KAPPA_SCALE (tests only): shrink both radii to measure how much margin the bound has on a given cloud
  float kscale = 1.f;
  if (const char* ks = getenv("NSVGD_MEDIAN_LARGE_KAPPA_SCALE")) kscale = (float)atof(ks);
  const float kappa_g = kscale * (3.0517578125e-5f + 1.5f * (float)T * 2.384185791015625e-7f);     // 2^-15 + 1.5 T 2^-22
  const float kappa_d = kscale * ((float)(dz + 4) * 5.9604644775390625e-8f * 1.0001f);              // (d + 4) 2^-24
  prof_begin(6, st);
  NSVGD_PROPAGATE(lg_launch_classify(st, f.Zb0, f.Zb1, n, f.n_pad, f.ldk, f.zn, f.fs, f.tlist, f.tcount, ML_TCAP, kappa_g, kappa_d, f.err_flag));
  prof_end(6, st);

  // --- exact evaluation of the listed pairs, tile by tile
  prof_begin(5, st);
  NSVGD_PROPAGATE(ensure_dyn_smem((const void*)ml_exact_tiles_kernel, ML_XSMEM));
  const int tiles_n = (int)(f.n_pad / ML_TC);
  ml_exact_tiles_kernel<<<dim3((unsigned)tiles_n, (unsigned)tiles_n), ML_XT, ML_XSMEM, st>>>(P, n, (int)dz, ldp, w_last, f.fs, f.tlist, f.tcount,
                                                                                          tiles_n, f.cand, f.cand_counts, f.cand_cap, main_state);
  NSVGD_LAUNCH_CHECK();
  prof_end(5, st);

  // --- resolve among the exact candidates; conditional SIMT fallback
  prof_begin(10, st);
  NSVGD_PROPAGATE(launch_fast_pack_resolve(st, f.fs, f.list_counts, 0xFFFFFFFFu, f.cand_counts, f.cand_cap, f.err_flag, main_state, f.cand_state, red));
  NSVGD_PROPAGATE(launch_fused_select(st, f.cand, f.cand_counts, MF_CSTRIDE, MF_NSEG, f.cand_cap, 0u, 0u, f.cand_state, &f.sel[1], nullptr, 2,
                                      f.fs, main_state, 0, 0));
  prof_end(10, st);
  prof_begin(11, st);
  for (int pass = 0; pass < NSVGD_SELECT_PASSES; ++pass) {
    NSVGD_PROPAGATE(launch_pair_hist(st, P, n, dz, ldp, w_last, 0, 1, main_state, pass, hist, &f.fs->fallback));
    NSVGD_PROPAGATE(launch_select_advance(st, main_state, hist, pass, &f.fs->fallback));
  }
  prof_end(11, st);
  if (hint) NSVGD_PROPAGATE(launch_hint_update(st, hint, main_state, f.fs, skip_flag, hint_key));
  return NSVGD_OK;
}

}  // namespace nsvgd

// the row-sharded host logic asks whether every rank should simply run the filtered select on the gathered array
extern "C" int nsvgd_median_large_supported(int64_t n, int64_t dz) { return nsvgd::median_large_supported(n, dz) ? 1 : 0; }
