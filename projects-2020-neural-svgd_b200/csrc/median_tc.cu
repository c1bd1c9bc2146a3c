// median_tc.cu — exact median-heuristic order statistics with a tensor-core filter
// [kernels.py:160-172, utils.py:178-197, 653-658].  Bit-exact with the 3-pass SIMT select and the
// oracle: the tensor cores only DISCARD pairs that provably lie outside a bracket; every value that
// can influence the result is recomputed with the defined-order fp32 arithmetic.
//
//   1. sample  : 2^22 pseudo-random pairs, exact distances, radix-select two sample quantiles
//                L <= (target order statistics) <= U   (6 sigma + slack of the binomial rank error)
//   2. filter  : ONE pass over all pairs on the tensor cores: Gram tile (bf16 2-piece split, 3 MMA
//                passes, TMEM accumulator) -> dt = |z_i|^2 + |z_j|^2 - 2 g.  With the rigorous error
//                radius eps_ij = kappa (|z_i|^2 + |z_j|^2):  dt + eps < L -> counted "below";
//                dt - eps > U -> dropped; otherwise the pair (i,j) is appended to a list.
//   3. exact   : listed pairs are recomputed in the defined order: < L counted, in [L,U] kept as
//                candidates, > U dropped.
//   4. resolve : rank' = rank - below must fall inside the candidates; radix-select among them.
//                If the bracket missed or a list overflowed (massive ties), a device-side flag makes
//                the conditional SIMT 3-pass kernels run instead — still enqueue-only, still exact.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "p2p.cuh"
#include "select.cuh"
#include "tc_common.cuh"

namespace nsvgd {

using namespace tc;

constexpr int MF_BM = 128, MF_BN = 64, MF_CHUNK = 64, MF_XSTAGES = 3;
constexpr int MF_CLS_WARPS = 8;                          // classification warps: 2 per TMEM lane quarter
constexpr int MF_THREADS = 64 + 32 * MF_CLS_WARPS;
constexpr int MF_NVEC = 4;                               // ring of per-tile |z_j|^2 vectors
constexpr int MF_STAGE = 128;                            // per-warp staging buffer (uncertain pairs) in shared memory
constexpr uint32_t MF_SAMPLES = 1u << 22;
// MF_NSEG (independent append segments, one atomic counter per 128-byte line) and MF_CSTRIDE (counter stride in uint32): select.cuh
// error radius of the filter relative to |z_i|^2 + |z_j|^2: bf16 2-piece products 2^-15, norms 2^-17,
// defined-order rounding of the exact value 2^-16, centring 2^-22  =>  < 2^-14.2; kappa = 2^-13.
constexpr float MF_KAPPA = 1.220703125e-4f;

bool median_fast_supported(int64_t n, int64_t dz) { return n >= 8192 && n <= (1ll << 22) && dz > 32 && dz <= 128; }

static uint32_t list_capacity(int64_t n) {
  const double P = 0.5 * (double)n * (double)(n - 1);
  return (uint32_t)std::min<double>(P / 48.0 + (double)(1 << 22), 4.0e9);
}
static uint32_t cand_capacity(int64_t n) {
  const double P = 0.5 * (double)n * (double)(n - 1);
  return (uint32_t)std::min<double>(P / 96.0 + (double)(1 << 22), 4.0e9);
}

// FusedSelScratch (scratch of one fused 3-pass select): select.cuh

struct FastWs {
  FusedSelScratch* sel;  // [2]: bracket select, candidate select (inside the zero block)
  char* zero_block; size_t zero_bytes;
  FastMedState* fs; SelectState* bracket; SelectState* cand_state;
  float* mu; double* mpart; __nv_bfloat16* Zb0; __nv_bfloat16* Zb1; float* zn; float* zlo; float* zhi; float* Ds;
  unsigned long long* list; float* cand; uint32_t* err_flag;
  uint32_t* list_counts; uint32_t* cand_counts;     // MF_NSEG counters each, MF_CSTRIDE apart
  uint32_t list_cap, cand_cap; int64_t n_pad; int dp;   // capacities PER SEGMENT
};

static size_t fast_layout(int64_t n, int64_t dz, WsCarver& w, FastWs* o) {
  const int dp = dz <= 64 ? 64 : 128;
  const int64_t n_pad = round_up(n, 128);
  FastWs f{};
  f.dp = dp; f.n_pad = n_pad; f.list_cap = list_capacity(n) / MF_NSEG; f.cand_cap = cand_capacity(n) / MF_NSEG;
  f.fs = w.take<FastMedState>(1);
  f.bracket = (SelectState*)w.take<char>(NSVGD_SELECT_STATE_BYTES);
  f.cand_state = (SelectState*)w.take<char>(NSVGD_SELECT_STATE_BYTES);
  f.mu = w.take<float>((size_t)dz);
  f.mpart = w.take<double>((size_t)colmean_partials(n, dz));
  f.Zb0 = w.take<__nv_bfloat16>((size_t)(n_pad * dp), 1024);
  f.Zb1 = w.take<__nv_bfloat16>((size_t)(n_pad * dp), 1024);
  f.zn = w.take<float>((size_t)n_pad);
  f.zlo = w.take<float>((size_t)n_pad);
  f.zhi = w.take<float>((size_t)n_pad);
  f.Ds = w.take<float>(MF_SAMPLES);
  f.list = w.take<unsigned long long>((size_t)f.list_cap * MF_NSEG);
  f.cand = w.take<float>((size_t)f.cand_cap * MF_NSEG);
  // one contiguous block that a single memset clears per call: error flag, list counters, select scratch
  w.off = (w.off + 1023) / 1024 * 1024;
  const size_t z0 = w.off;
  f.err_flag = w.take<uint32_t>(64);   // [0] error flag; bytes 64.. = the single-GPU reduce block
  f.list_counts = w.take<uint32_t>(MF_NSEG * MF_CSTRIDE);
  f.cand_counts = w.take<uint32_t>(MF_NSEG * MF_CSTRIDE);
  f.sel = w.take<FusedSelScratch>(2);
  f.zero_block = w.dry ? nullptr : w.base + z0;
  f.zero_bytes = w.off - z0;
  if (o) *o = f;
  return w.off;
}

size_t median_fast_workspace(int64_t n, int64_t dz) {
  WsCarver w(nullptr, 0, true);
  return fast_layout(n, dz, w, nullptr) + 1024;
}

// ---------------------------------------------------------------------------- exact pair distance, vectorised loads
__device__ __forceinline__ float exact_pair_sqdist_v4(const float* __restrict__ a, const float* __restrict__ b, int d,
                                                      float w_last) {
  // Loads are issued in batches of 8 float4 per row BEFORE the (inherently sequential) defined-order adds: the kernel
  // is latency-bound (ncu: issue 12 %, L2 26 %), and one round trip per 4 coordinates cost 32 serial L2 latencies per
  // pair; batching makes it 4.  The arithmetic and its order are unchanged (bit-exact contract).
  // Only the last float4 group (or the scalar tail) can hold the weighted coordinate d-1: the batched groups need no
  // per-element test.
  float acc = 0.f;
  const int dm = d - 1;            // coordinates [0, dm) unweighted, dm weighted
  const int groups = d >> 2;
  const int plain = (d & 3) ? groups : groups - 1;     // float4 groups made of unweighted coordinates only
  float last_diff = 0.f;
  constexpr int B = 8;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  int g0 = 0;
  for (; g0 + B <= plain; g0 += B) {
    float4 av[B], bv[B];
#pragma unroll
    for (int u = 0; u < B; ++u) { av[u] = __ldg(a4 + g0 + u); bv[u] = __ldg(b4 + g0 + u); }
#pragma unroll
    for (int u = 0; u < B; ++u) {
      const float d0 = av[u].x - bv[u].x, d1 = av[u].y - bv[u].y, d2 = av[u].z - bv[u].z, d3 = av[u].w - bv[u].w;
      acc = __fadd_rn(acc, __fmul_rn(d0, d0));
      acc = __fadd_rn(acc, __fmul_rn(d1, d1));
      acc = __fadd_rn(acc, __fmul_rn(d2, d2));
      acc = __fadd_rn(acc, __fmul_rn(d3, d3));
    }
  }
  for (int g = g0; g < groups; ++g) {
    const float4 av = __ldg(a4 + g), bv = __ldg(b4 + g);
    const float da[4] = {av.x - bv.x, av.y - bv.y, av.z - bv.z, av.w - bv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * g + e;
      if (c < dm) acc = __fadd_rn(acc, __fmul_rn(da[e], da[e]));
      else last_diff = da[e];
    }
  }
  for (int c = groups << 2; c < d; ++c) {
    const float diff = a[c] - b[c];
    if (c < dm) acc = __fadd_rn(acc, __fmul_rn(diff, diff));
    else last_diff = diff;
  }
  float t = __fmul_rn(last_diff, last_diff);
  t = __fmul_rn(w_last, t);
  return __fadd_rn(acc, t);
}

__device__ __forceinline__ float exact_pair(const float* P, int64_t ldp, int64_t i, int64_t j, int d, float w_last,
                                            bool vec) {
  return vec ? exact_pair_sqdist_v4(P + i * ldp, P + j * ldp, d, w_last)
             : exact_pair_sqdist(P + i * ldp, P + j * ldp, d, w_last);
}

// one element of the filter: below = (t1 < lo) -> ++cnt;  uncertain = !(below || t2 > hi) -> set `bit` in unc.
// Four SASS instructions: FSETP, FSETP.OR, @p IADD, @!q LOP.
__device__ __forceinline__ void classify(uint32_t& cnt, uint32_t& unc, float t1, float lo, float t2, float hi, uint32_t bit) {
  asm("{\n\t.reg .pred p, q;\n\t"
      "setp.lt.f32 p, %2, %3;\n\t"
      "setp.gt.or.f32 q, %4, %5, p;\n\t"
      "@p add.u32 %0, %0, 1;\n\t"
      "@!q or.b32 %1, %1, %6;\n\t}"
      : "+r"(cnt), "+r"(unc)
      : "f"(t1), "f"(lo), "f"(t2), "f"(hi), "r"(bit));
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}


// ---- warp-cooperative exact distances: 32 pairs per warp, rows read with fully coalesced 128-byte
// requests (lane l loads coordinates l, l+32, l+64, l+96), squared differences parked in shared
// memory, then lane p adds pair p's terms in the DEFINED sequential order.  Bit-identical to
// exact_pair_sqdist; halves the L2 sector traffic of the one-thread-per-pair gather.
constexpr int WX_STRIDE = 129;                       // floats per pair row in shared memory (dz <= 128)
constexpr int WX_WARPS = 4;
__device__ __forceinline__ float warp_exact32(const float* __restrict__ P, int64_t ldp, int d, float w_last, int64_t i,
                                              int64_t j, float* __restrict__ sq, int lane) {
  for (int pp = 0; pp < 32; ++pp) {
    const int64_t ip = __shfl_sync(0xffffffffu, i, pp), jp = __shfl_sync(0xffffffffu, j, pp);
    const float* a = P + ip * ldp;
    const float* b = P + jp * ldp;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = lane + 32 * e;
      if (c < d) {
        const float diff = __ldg(a + c) - __ldg(b + c);
        float t = __fmul_rn(diff, diff);
        if (c == d - 1) t = __fmul_rn(w_last, t);
        sq[pp * WX_STRIDE + c] = t;
      }
    }
  }
  __syncwarp();
  float acc = 0.f;
  const float* mine = sq + lane * WX_STRIDE;
  for (int c = 0; c < d; ++c) acc = __fadd_rn(acc, mine[c]);
  __syncwarp();
  return acc;
}

__global__ void __launch_bounds__(256) sample_pairs_kernel(const float* __restrict__ P, int64_t n, int d, int64_t ldp,
                                                           float w_last, int vec, uint32_t ns, float* __restrict__ Ds) {
  for (uint32_t s = blockIdx.x * 256u + threadIdx.x; s < ns; s += gridDim.x * 256u) {
    const uint64_t z = splitmix64(s);
    const uint64_t i = ((z >> 32) * (uint64_t)n) >> 32;
    uint64_t j = ((z & 0xFFFFFFFFull) * (uint64_t)(n - 1)) >> 32;
    if (j >= i) ++j;
    Ds[s] = exact_pair(P, ldp, (int64_t)min(i, j), (int64_t)max(i, j), d, w_last, vec != 0);
  }
}

__global__ void __launch_bounds__(32 * WX_WARPS) sample_pairs_warp_kernel(const float* __restrict__ P, int64_t n, int d,
                                                                         int64_t ldp, float w_last, uint32_t s_begin,
                                                                         uint32_t ns, float* __restrict__ Ds) {
  // samples [s_begin, ns): sample s is a pure function of s, so ranks can each take a slice
  extern __shared__ float wx_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* sq = wx_smem + w * 32 * WX_STRIDE;
  for (uint32_t s0 = s_begin + (blockIdx.x * WX_WARPS + w) * 32u; s0 < ns; s0 += gridDim.x * WX_WARPS * 32u) {
    const uint32_t sidx = min(s0 + lane, ns - 1);
    const uint64_t z = splitmix64(sidx);
    const uint64_t i = ((z >> 32) * (uint64_t)n) >> 32;
    uint64_t j = ((z & 0xFFFFFFFFull) * (uint64_t)(n - 1)) >> 32;
    if (j >= i) ++j;
    const float dv = warp_exact32(P, ldp, d, w_last, (int64_t)min(i, j), (int64_t)max(i, j), sq, lane);
    if (s0 + lane < ns) Ds[s0 + lane] = dv;
  }
}

// Stratified sample for the bracket: warp-block b (32 consecutive samples) takes row i = floor(b n / #blocks) -- every row
// the same number of times (+-1) -- against 32 pseudo-random partners j != i.  For a proportion the variance of a
// proportionally stratified sample never exceeds the i.i.d. binomial variance the bracket width is computed from, every
// unordered pair is still equally likely, and row i is read once per 32 pairs: half the L2 traffic of i.i.d. pairs
// (0.85 -> 0.45 ms at n=65536, d=128).  The values only become the thresholds L, U, so plain FMA arithmetic is enough here;
// everything that can reach the result is recomputed in the defined order by exact_list_kernel.
__global__ void __launch_bounds__(32 * WX_WARPS) sample_rows_kernel(const float* __restrict__ P, int64_t n, int d, int64_t ldp,
                                                                   float w_last, uint32_t s_begin, uint32_t ns,
                                                                   float* __restrict__ Ds, uint32_t NB, const uint32_t* skip_flag) {
  // NB = warp-blocks of the whole sample set (MF_SAMPLES / 32, or the size of a REPLICATED reduced set / 32)
  if (skip_flag && *skip_flag) return;      // the bracket comes from the caller's hint this time
  __shared__ float part[WX_WARPS][32 * 33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* mine = part[w];
  for (uint32_t s0 = s_begin + (blockIdx.x * WX_WARPS + w) * 32u; s0 < ns; s0 += gridDim.x * WX_WARPS * 32u) {
    const uint64_t i = ((uint64_t)(s0 >> 5) * (uint64_t)n) / NB;
    const float* a = P + i * ldp;
    float xi[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) xi[e] = (lane + 32 * e < d) ? __ldg(a + lane + 32 * e) : 0.f;
    const uint64_t z = splitmix64(s0 + lane);
    uint64_t j = ((z >> 32) * (uint64_t)(n - 1)) >> 32;
    if (j >= i) ++j;
#pragma unroll 4
    for (int pp = 0; pp < 32; ++pp) {
      const int64_t jp = __shfl_sync(0xffffffffu, (int64_t)j, pp);
      const float* b = P + jp * ldp;
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = lane + 32 * e;
        if (c < d) {
          const float diff = xi[e] - __ldg(b + c);
          const float t = diff * diff;
          acc += (c == d - 1) ? w_last * t : t;
        }
      }
      mine[pp * 33 + lane] = acc;
    }
    __syncwarp();
    float sum = 0.f;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) sum += mine[lane * 33 + l];
    Ds[s0 + lane] = sum;
    __syncwarp();
  }
}

// per-column classification thresholds: below <=> fma(g,-2,A1) < zlo[j], above <=> fma(g,-2,A2) > zhi[j]
__global__ void filter_vectors_kernel(const float* __restrict__ zn, int64_t n_pad, float* __restrict__ zlo,
                                      float* __restrict__ zhi) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_pad; t += (int64_t)gridDim.x * blockDim.x) {
    const float z = zn[t];
    zlo[t] = -z * (1.f + MF_KAPPA);
    zhi[t] = -z * (1.f - MF_KAPPA);
  }
}

// ---------------------------------------------------------------------------- the tensor-core filter
struct MfParams {
  const uint32_t* Qb0; const uint32_t* Qb1;   // centred bf16 pieces [n_pad x DP] viewed as packed pairs (query block -> TMEM)
  const float* zn; const float* zlo; const float* zhi;
  int64_t n;
  int n_tiles;                 // n_pad / 64
  int64_t rb_start, rb_stride; // 128-row blocks handled by this rank
  FastMedState* fs;
  unsigned long long* list;
  uint32_t* list_counts;
  uint32_t list_cap;           // per segment
  uint32_t* err_flag;
};

template <int DP>
struct MfSmem {
  static constexpr int KH = DP / 64;
  static constexpr int QSLAB = MF_BM * 128;
  static constexpr int XSLAB = MF_BN * 128;
  // The 128-row query block lives in TENSOR MEMORY (A operand of tcgen05.mma TS, as in phi_tc.cu): shared memory holds
  // only the 3-stage ring of column tiles, 107 KB per CTA, so TWO CTAs share an SM (256 TMEM columns each) and one CTA's
  // TMEM read + classification overlaps the other's MMAs.
  static constexpr int X_STAGE_BYTES = 2 * KH * XSLAB;
  static constexpr int OFF_X = 0;
  static constexpr int OFF_VEC = OFF_X + MF_XSTAGES * X_STAGE_BYTES;
  static constexpr int OFF_STAGE = OFF_VEC + MF_NVEC * 2 * MF_BN * 4;
  static constexpr int OFF_BAR = OFF_STAGE + MF_CLS_WARPS * MF_STAGE * 8;
  static constexpr int NBARS = 1 + 2 * MF_XSTAGES + 2 * MF_NVEC + 4;
  static constexpr int DYN_BYTES = OFF_BAR + NBARS * 8 + 16 + 1024;
  static_assert(2 * (DYN_BYTES + 1024) <= 232448, "two CTAs per SM must fit");
};

constexpr uint32_t MF_TMEM_COLS = 256, MF_TM_Q = 128;   // G double buffer [0,128) | Q piece b0 [128,192) | Q piece b1 [192,256)

template <int DP>
__global__ void __launch_bounds__(MF_THREADS, 2)
median_filter_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                        const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                        const MfParams p) {
  using L = MfSmem<DP>;
  constexpr int KH = L::KH;
  const int64_t rb = p.rb_start + p.rb_stride * (int64_t)blockIdx.y;
  const int t_begin = (int)(2 * rb) + (int)blockIdx.x * MF_CHUNK;
  if (t_begin >= p.n_tiles) return;
  const int t_end = min(t_begin + MF_CHUNK, p.n_tiles);
  const int T = t_end - t_begin;
  const int m0 = (int)(rb * MF_BM);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-B aligned, stays in the shared address space
  uint8_t* sX = smem + L::OFF_X;
  uint64_t* bars = (uint64_t*)(smem + L::OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* x_full = q_full + 1;
  uint64_t* x_empty = x_full + MF_XSTAGES;
  uint64_t* n_full = x_empty + MF_XSTAGES;
  uint64_t* n_empty = n_full + MF_NVEC;
  uint64_t* g_full = n_empty + MF_NVEC;
  uint64_t* g_empty = g_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(g_empty + 2);
  float* sVec = (float*)(smem + L::OFF_VEC);
  unsigned long long* sStage = (unsigned long long*)(smem + L::OFF_STAGE);
  volatile uint32_t* abort_s = tmem_slot + 1;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  Watch watch{abort_s, p.err_flag};

  if (threadIdx.x == 0) {
    *abort_s = (*(volatile uint32_t*)p.err_flag) ? 1u : 0u;
    mbar_init(q_full, MF_CLS_WARPS);             // query pieces stored to TMEM by the classification warps
    for (int i = 0; i < MF_XSTAGES; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < MF_NVEC; ++i) { mbar_init(&n_full[i], 1); mbar_init(&n_empty[i], MF_CLS_WARPS); }
    for (int i = 0; i < 2; ++i) { mbar_init(&g_full[i], 1); mbar_init(&g_empty[i], MF_CLS_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, MF_TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    {
      const bool leader = elect_one();
      for (int t = 0; t < T; ++t) {
        const int st = t % MF_XSTAGES, u = t / MF_XSTAGES;
        mbar_wait(&x_empty[st], (u & 1) ^ 1, watch, 0x100 + st);
        uint8_t* dst = sX + st * L::X_STAGE_BYTES;
        const int j0 = (t_begin + t) * MF_BN;
        if (leader) {
          mbar_arrive_expect_tx(&x_full[st], L::X_STAGE_BYTES);
          for (int kh = 0; kh < KH; ++kh) {
            tma_load_2d(dst + (0 * KH + kh) * L::XSLAB, &tmB0, &x_full[st], kh * 64, j0);
            tma_load_2d(dst + (1 * KH + kh) * L::XSLAB, &tmB1, &x_full[st], kh * 64, j0);
          }
        }
        const int ns = t % MF_NVEC, un = t / MF_NVEC;
        mbar_wait(&n_empty[ns], (un & 1) ^ 1, watch, 0x180 + ns);
        if (leader) {
          mbar_arrive_expect_tx(&n_full[ns], 2 * MF_BN * 4);
          bulk_load_1d(sVec + ns * 2 * MF_BN, p.zlo + j0, MF_BN * 4, &n_full[ns]);
          bulk_load_1d(sVec + ns * 2 * MF_BN + MF_BN, p.zhi + j0, MF_BN * 4, &n_full[ns]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(MF_BM, MF_BN);
      const uint32_t sX_a = smem_u32(sX);
      mbar_wait(q_full, 0, watch, 0x300);
      tc_fence_after();
      for (int t = 0; t < T; ++t) {
        const int b = t & 1, ub = t >> 1;
        const int st = t % MF_XSTAGES, us = t / MF_XSTAGES;
        mbar_wait(&g_empty[b], (ub & 1) ^ 1, watch, 0x310 + b);
        mbar_wait(&x_full[st], us & 1, watch, 0x320 + st);
        tc_fence_after();
        const uint32_t d_tm = tmem + b * MF_BN;
        const uint32_t xs = sX_a + st * L::X_STAGE_BYTES;
        if (leader) {
          uint32_t acc = 0;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const int pa = pass == 2 ? 1 : 0;
            const int pb = pass == 1 ? 1 : 0;
#pragma unroll
            for (int kh = 0; kh < KH; ++kh) {
              const uint64_t bd = umma_desc_sw128(xs + (pb * KH + kh) * L::XSLAB);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                umma_ts(d_tm, tmem + MF_TM_Q + pa * 64 + (kh * 4 + ks) * 8, umma_desc_advance(bd, ks * 32), idesc, acc);
                acc = 1;
              }
            }
          }
          umma_commit(&x_empty[st]);
          umma_commit(&g_full[b]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---- classification warps: thread <-> (row i, 32-column half of the tile)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int64_t gi = (int64_t)m0 + row_in_tile;
    {
      // the CTA's 128 query rows (two bf16 pieces) -> TMEM, once: A operand of every Gram MMA
      constexpr int QW = DP / 4;                       // packed words per thread and piece (2 column halves per row)
      const int64_t qoff = gi * (DP / 2) + half * QW;
#pragma unroll
      for (int piece = 0; piece < 2; ++piece) {
        const uint4* src = reinterpret_cast<const uint4*>((piece == 0 ? p.Qb0 : p.Qb1) + qoff);
#pragma unroll
        for (int c0 = 0; c0 < QW; c0 += 16) {
          uint32_t qv[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 v = __ldg(src + c0 / 4 + q);
            qv[q * 4] = v.x; qv[q * 4 + 1] = v.y; qv[q * 4 + 2] = v.z; qv[q * 4 + 3] = v.w;
          }
          tmem_st16(tmem + lane_addr + MF_TM_Q + piece * 64 + half * QW + c0, qv);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full);
    }
    const float qn = p.zn[gi];
    const float Lv = p.fs->L, Uv = p.fs->U;
    // below  <=>  -2g + qn(1+k) - L < -zn_j(1+k)       above  <=>  -2g + qn(1-k) - U > -zn_j(1-k)
    const float A1 = fmaf(qn, 1.f + MF_KAPPA, -Lv), A2 = fmaf(qn, 1.f - MF_KAPPA, -Uv);
    const float2 A1_2 = make_float2(A1, A1), A2_2 = make_float2(A2, A2), m2_2 = make_float2(-2.f, -2.f);
    uint32_t cnt = 0;
    const int seg = (int)((blockIdx.y * gridDim.x + blockIdx.x) % MF_NSEG);
    unsigned long long* seg_list = p.list + (size_t)seg * p.list_cap;
    unsigned long long* stage = sStage + (warp - 2) * MF_STAGE;
    uint32_t fill = 0;                                   // warp-uniform
    auto flush = [&]() {
      if (fill == 0u) return;
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&p.list_counts[seg * MF_CSTRIDE], fill);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (uint32_t e = lane; e < fill; e += 32)
        if (base + e < p.list_cap) seg_list[base + e] = stage[e];
      __syncwarp();
      fill = 0u;
    };
    for (int t = 0; t < T; ++t) {
      const int b = t & 1, ub = t >> 1;
      const int ns = t % MF_NVEC, un = t / MF_NVEC;
      const int tile = t_begin + t;
      const int64_t j0 = (int64_t)tile * MF_BN + half * 32;
      mbar_wait(&g_full[b], ub & 1, watch, 0x400 + b);
      __syncwarp();
      tc_fence_after();
      uint32_t g[32];
      tmem_ld32(tmem + lane_addr + b * MF_BN + half * 32, g);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&g_empty[b]);
      mbar_wait(&n_full[ns], un & 1, watch, 0x420 + ns);
      const float4* zlo4 = reinterpret_cast<const float4*>(sVec + ns * 2 * MF_BN + half * 32);
      const float4* zhi4 = reinterpret_cast<const float4*>(sVec + ns * 2 * MF_BN + MF_BN + half * 32);
      // boundary tiles (diagonal or padded) need the i < j < n mask
      const bool boundary = (tile <= 2 * (int)rb + 1) || ((int64_t)(tile + 1) * MF_BN > p.n) || ((int64_t)m0 + MF_BM > p.n);
      uint32_t unc = 0u;
      if (!boundary) {
        // interior tile (the common case): per element 2 FSETP (the second OR-combined), 1 predicated IADD, 1 predicated LOP
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 lo4 = zlo4[q4];
          const float4 hi4 = zhi4[q4];
          const float2 g01 = make_float2(__uint_as_float(g[q4 * 4]), __uint_as_float(g[q4 * 4 + 1]));
          const float2 g23 = make_float2(__uint_as_float(g[q4 * 4 + 2]), __uint_as_float(g[q4 * 4 + 3]));
          const float2 ta = __ffma2_rn(g01, m2_2, A1_2), tb = __ffma2_rn(g23, m2_2, A1_2);
          const float2 ua = __ffma2_rn(g01, m2_2, A2_2), ub2 = __ffma2_rn(g23, m2_2, A2_2);
          classify(cnt, unc, ta.x, lo4.x, ua.x, hi4.x, 1u << (q4 * 4));
          classify(cnt, unc, ta.y, lo4.y, ua.y, hi4.y, 1u << (q4 * 4 + 1));
          classify(cnt, unc, tb.x, lo4.z, ub2.x, hi4.z, 1u << (q4 * 4 + 2));
          classify(cnt, unc, tb.y, lo4.w, ub2.y, hi4.w, 1u << (q4 * 4 + 3));
        }
      } else {
        // diagonal / padded tile: mask to i < j < n  (fully unrolled: g[] must stay in registers)
#pragma unroll
        for (int col = 0; col < 32; ++col) {
          const float gv = __uint_as_float(g[col]);
          const float lo = sVec[ns * 2 * MF_BN + half * 32 + col], hi = sVec[ns * 2 * MF_BN + MF_BN + half * 32 + col];
          const int64_t gj = j0 + col;
          const bool valid = (gi < gj) && (gj < p.n) && (gi < p.n);
          const bool below = fmaf(gv, -2.f, A1) < lo;
          const bool above = fmaf(gv, -2.f, A2) > hi;
          cnt += (below && valid) ? 1u : 0u;
          if (valid && !(below || above)) unc |= 1u << col;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&n_empty[ns]);
      // uncertain pairs -> per-warp staging buffer in shared memory (flushed with ONE global atomic)
      const uint32_t c = __popc(unc);
      if (__any_sync(0xffffffffu, c != 0u)) {
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (fill + total > (uint32_t)MF_STAGE) flush();
        if (total > (uint32_t)MF_STAGE) {
          // pathological tile (ties): straight to global memory with its own reservation
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(&p.list_counts[seg * MF_CSTRIDE], total);
          base = __shfl_sync(0xffffffffu, base, 0);
          uint32_t pos = base + incl - c;
          while (unc) {
            const int cc = __ffs(unc) - 1;
            unc &= unc - 1;
            if (pos < p.list_cap) seg_list[pos] = ((unsigned long long)gi << 32) | (unsigned long long)(j0 + cc);
            ++pos;
          }
        } else {
          uint32_t pos = fill + incl - c;
          while (unc) {
            const int cc = __ffs(unc) - 1;
            unc &= unc - 1;
            stage[pos++] = ((unsigned long long)gi << 32) | (unsigned long long)(j0 + cc);
          }
          fill += total;
          __syncwarp();
        }
      }
    }
    flush();
    uint32_t tot = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0 && tot) atomicAdd(&p.fs->below, (unsigned long long)tot);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, MF_TMEM_COLS); }
}

// ---------------------------------------------------------------------------- exact evaluation of the uncertain list
__global__ void __launch_bounds__(32 * WX_WARPS) exact_list_warp_kernel(const float* __restrict__ P, int d, int64_t ldp,
                                                                       float w_last, FastMedState* fs,
                                                                       const unsigned long long* __restrict__ list,
                                                                       const uint32_t* __restrict__ list_counts,
                                                                       uint32_t list_cap, float* __restrict__ cand,
                                                                       uint32_t* cand_counts, uint32_t cand_cap,
                                                                       SelectState* main_state) {
  extern __shared__ float wx_smem[];
  __shared__ unsigned int sh_below;
  if (threadIdx.x == 0) sh_below = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* sq = wx_smem + w * 32 * WX_STRIDE;
  const float Lv = fs->L, Uv = fs->U;
  uint32_t below = 0;
  bool any_nan = false;
  for (int seg = blockIdx.y; seg < MF_NSEG; seg += gridDim.y) {
    const uint32_t count = min(list_counts[seg * MF_CSTRIDE], list_cap);
    const unsigned long long* seg_list = list + (size_t)seg * list_cap;
    float* seg_cand = cand + (size_t)seg * cand_cap;
    for (uint32_t t0 = (blockIdx.x * WX_WARPS + w) * 32u; t0 < count; t0 += gridDim.x * WX_WARPS * 32u) {
      const uint32_t t = t0 + lane;
      const unsigned long long e = seg_list[min(t, count - 1)];
      const float dv = warp_exact32(P, ldp, d, w_last, (int64_t)(e >> 32), (int64_t)(e & 0xFFFFFFFFull), sq, lane);
      bool is_cand = false;
      if (t < count) {
        if (dv != dv) any_nan = true;
        else if (dv < Lv) ++below;
        else if (dv <= Uv) is_cand = true;
      }
      const uint32_t ball = __ballot_sync(0xffffffffu, is_cand);
      if (ball) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&cand_counts[seg * MF_CSTRIDE], (uint32_t)__popc(ball));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (is_cand) {
          const uint32_t pos = base + __popc(ball & ((1u << lane) - 1u));
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

__global__ void __launch_bounds__(256, 2) exact_list_kernel(const float* __restrict__ P, int d, int64_t ldp, float w_last,
                                                         int vec, FastMedState* fs, const unsigned long long* __restrict__ list,
                                                         const uint32_t* __restrict__ list_counts, uint32_t list_cap,
                                                         float* __restrict__ cand, uint32_t* cand_counts, uint32_t cand_cap,
                                                         SelectState* main_state) {
  __shared__ unsigned int sh_below;
  if (threadIdx.x == 0) sh_below = 0u;
  __syncthreads();
  const float Lv = fs->L, Uv = fs->U;
  const int lane = threadIdx.x & 31;
  uint32_t below = 0;
  bool any_nan = false;
  for (int seg = blockIdx.y; seg < MF_NSEG; seg += gridDim.y) {
    const uint32_t count = min(list_counts[seg * MF_CSTRIDE], list_cap);
    const unsigned long long* seg_list = list + (size_t)seg * list_cap;
    float* seg_cand = cand + (size_t)seg * cand_cap;
    // warp-uniform trip count so that the ballot below sees the whole warp
    for (uint32_t t0 = (blockIdx.x * 256u + threadIdx.x) & ~31u; t0 < count; t0 += gridDim.x * 256u) {
      const uint32_t t = t0 + lane;
      bool is_cand = false;
      float dv = 0.f;
      if (t < count) {
        const unsigned long long e = seg_list[t];
        const int64_t i = (int64_t)(e >> 32), j = (int64_t)(e & 0xFFFFFFFFull);
        dv = exact_pair(P, ldp, i, j, d, w_last, vec != 0);
        if (dv != dv) any_nan = true;
        else if (dv < Lv) ++below;
        else if (dv <= Uv) is_cand = true;
      }
      const uint32_t ball = __ballot_sync(0xffffffffu, is_cand);
      if (ball) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&cand_counts[seg * MF_CSTRIDE], (uint32_t)__popc(ball));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (is_cand) {
          const uint32_t pos = base + __popc(ball & ((1u << lane) - 1u));
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

// Slice-cooperative exact evaluation of the listed pairs.  exact_list_kernel (one thread per pair, each lane walking its
// own two rows) is bound by the L1 tag stage: every LDG.128 of a warp touches 32 different lines and each line is looked up
// 8 times (ncu: l1tex throughput 85 %, L2 26 %, issue 12 %; profiles/r01_ncu_exact_final.txt).  Here a warp takes 32 pairs
// and moves their 64 rows through shared memory one 128-byte column slice at a time: 8 lanes read one line of one row
// (one lookup per line, 16 loads in flight per lane); the loading lane holds both rows of its pair for its 4 coordinates,
// so it parks the squared differences (last coordinate weighted) in ONE padded 32 x 32 tile, and lane p then adds pair p's
// 32 terms of the slice in the DEFINED sequential order (same arithmetic, bit-identical results).
// Requires d % 4 == 0, ldp % 4 == 0, 16-byte aligned rows, d <= 128.
constexpr int XL_STRIDE = 36;                         // floats per tile row: 32 + 4 pad -> conflict-free LDS.128 / STS.128
constexpr int XL_TILE = 32 * XL_STRIDE;
constexpr int XL_SMEM = 8 * XL_TILE * (int)sizeof(float);
__global__ void __launch_bounds__(256, 2) exact_list_coop_kernel(const float* __restrict__ P, int d, int64_t ldp, float w_last,
                                                                 FastMedState* fs, const unsigned long long* __restrict__ list,
                                                                 const uint32_t* __restrict__ list_counts, uint32_t list_cap,
                                                                 float* __restrict__ cand, uint32_t* cand_counts, uint32_t cand_cap,
                                                                 SelectState* main_state) {
  extern __shared__ __align__(16) float xl_smem[];
  __shared__ unsigned int sh_below;
  if (threadIdx.x == 0) sh_below = 0u;
  __syncthreads();
  const float Lv = fs->L, Uv = fs->U;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = lane >> 3, q = lane & 7;              // cooperative loads: 8 lanes per row, 4 rows per instruction
  float* tA = xl_smem + w * XL_TILE;
  const int nslices = (d + 31) >> 5;
  const int last_group = (d >> 2) - 1;                // the float4 group that holds the weighted coordinate d-1
  uint32_t below = 0;
  bool any_nan = false;
  for (int seg = blockIdx.y; seg < MF_NSEG; seg += gridDim.y) {
    const uint32_t count = min(list_counts[seg * MF_CSTRIDE], list_cap);
    const unsigned long long* seg_list = list + (size_t)seg * list_cap;
    float* seg_cand = cand + (size_t)seg * cand_cap;
    for (uint32_t t0 = (blockIdx.x * 256u + threadIdx.x) & ~31u; t0 < count; t0 += gridDim.x * 256u) {
      const uint32_t t = t0 + lane;
      const unsigned long long e = seg_list[min(t, count - 1u)];      // idle lanes shadow the last pair (valid rows)
      const uint32_t i = (uint32_t)(e >> 32), j = (uint32_t)(e & 0xFFFFFFFFull);
      uint32_t ia[8], jb[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        ia[u] = __shfl_sync(0xffffffffu, i, 4 * u + r);
        jb[u] = __shfl_sync(0xffffffffu, j, 4 * u + r);
      }
      float acc = 0.f;
      for (int sl = 0; sl < nslices; ++sl) {
        const int col = 32 * sl + 4 * q;
        const bool inb = col < d;
        const bool wgt = (col >> 2) == last_group;     // this lane's float4 ends with the weighted coordinate d-1
        float4 va[8], vb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          va[u] = inb ? __ldg(reinterpret_cast<const float4*>(P + (int64_t)ia[u] * ldp + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
          vb[u] = inb ? __ldg(reinterpret_cast<const float4*>(P + (int64_t)jb[u] * ldp + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();                                  // the previous slice has been consumed
        // the loading lane holds both rows of its pair for these 4 coordinates: park the squared differences (the last
        // coordinate already weighted: w * (diff * diff), as the defined order prescribes) -- one tile instead of two
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float d0 = va[u].x - vb[u].x, d1 = va[u].y - vb[u].y, d2 = va[u].z - vb[u].z, d3 = va[u].w - vb[u].w;
          float t3 = __fmul_rn(d3, d3);
          if (wgt) t3 = __fmul_rn(w_last, t3);
          *reinterpret_cast<float4*>(tA + (4 * u + r) * XL_STRIDE + 4 * q) =
              make_float4(__fmul_rn(d0, d0), __fmul_rn(d1, d1), __fmul_rn(d2, d2), t3);
        }
        __syncwarp();
        const int kmax = min(8, (d - 32 * sl) >> 2);
        for (int k = 0; k < kmax; ++k) {
          const float4 t4 = *reinterpret_cast<const float4*>(tA + lane * XL_STRIDE + 4 * k);
          acc = __fadd_rn(acc, t4.x);
          acc = __fadd_rn(acc, t4.y);
          acc = __fadd_rn(acc, t4.z);
          acc = __fadd_rn(acc, t4.w);                  // for the last group this is acc + w (diff diff): the final add
        }
      }
      const float dv = acc;
      bool is_cand = false;
      if (t < count) {
        if (dv != dv) any_nan = true;
        else if (dv < Lv) ++below;
        else if (dv <= Uv) is_cand = true;
      }
      const uint32_t ball = __ballot_sync(0xffffffffu, is_cand);
      if (ball) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&cand_counts[seg * MF_CSTRIDE], (uint32_t)__popc(ball));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (is_cand) {
          const uint32_t pos = base + __popc(ball & ((1u << lane) - 1u));
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

// local summary of stage 1 (per rank): red = {below, candidates (clamped), overflow|error, listed pairs}
__global__ void fast_pack_kernel(FastMedState* fs, const uint32_t* list_counts, uint32_t list_cap, uint32_t* cand_counts,
                                 uint32_t cand_cap, const uint32_t* err_flag, long long* red) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long bad = *err_flag != 0u ? 1 : 0;
  unsigned long long ncand = 0ull, nlist = 0ull;
  for (int sgm = 0; sgm < MF_NSEG; ++sgm) {
    const uint32_t lc = list_counts[sgm * MF_CSTRIDE], cc = cand_counts[sgm * MF_CSTRIDE];
    if (lc > list_cap || cc > cand_cap) bad = 1;                   // a list overflowed (massive ties)
    if (cc > cand_cap) cand_counts[sgm * MF_CSTRIDE] = cand_cap;
    ncand += min(cc, cand_cap); nlist += lc;
  }
  fs->list_count = (uint32_t)min(nlist, 0xFFFFFFFFull);
  fs->cand_count = (uint32_t)min(ncand, 0xFFFFFFFFull);
  red[0] = (long long)fs->below; red[1] = (long long)ncand; red[2] = bad; red[3] = (long long)nlist;
}

// red holds the GLOBAL sums (all-reduced across ranks when row-sharded)
__device__ __forceinline__ void fast_resolve_body(FastMedState* fs, const SelectState* main_state, SelectState* cs,
                                                  const long long* red);

__global__ void fast_resolve_kernel(FastMedState* fs, const SelectState* main_state, SelectState* cs, const long long* red) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  fast_resolve_body(fs, main_state, cs, red);
}

// row-sharded: the all-reduce of {below, candidates, bad, listed} over peer memory fused into the resolve step
__global__ void __launch_bounds__(64) fast_resolve_p2p_kernel(const P2PComm c, FastMedState* fs, const SelectState* main_state,
                                                              SelectState* cs, long long* red) {
  const bool ok = p2p_allreduce_cta(c, reinterpret_cast<unsigned long long*>(red), 4, false);
  if (threadIdx.x == 0) {
    if (!ok) const_cast<SelectState*>(main_state)->nan_flag = 1u;       // late peer -> h = NaN on this rank
    fast_resolve_body(fs, main_state, cs, red);
  }
}

// row-sharded: ONE kernel per radix pass = all-reduce of [candidate histogram | fallback histogram] over peer memory
// + the candidate-select decision + (only when the fallback flag is set) the SIMT-select decision
__global__ void __launch_bounds__(1024) fast_advance2_p2p_kernel(const P2PComm c, const FastMedState* fs, SelectState* cand_state,
                                                                 SelectState* main_state, unsigned long long* hist, int pass) {
  if (!p2p_allreduce_cta<1>(c, hist, 4 * NSVGD_SELECT_BINS, false) && threadIdx.x == 0) main_state->nan_flag = 1u;
  select_advance_body(cand_state, hist, pass);
  if (fs->fallback) select_advance_body(main_state, hist + 2 * NSVGD_SELECT_BINS, pass);
}

__device__ __forceinline__ void fast_resolve_body(FastMedState* fs, const SelectState* main_state, SelectState* cs,
                                                  const long long* red) {
  const unsigned long long below = (unsigned long long)red[0], ncand = (unsigned long long)red[1];
  uint32_t fb = red[2] != 0 ? 1u : 0u;
  cs->prefix[0] = cs->prefix[1] = 0u; cs->mask = 0u; cs->nan_flag = 0u;
  cs->value[0] = cs->value[1] = 0.f; cs->log_np1 = main_state->log_np1; cs->done = 0;
  for (int t = 0; t < 2; ++t) {
    cs->is_zero[t] = main_state->is_zero[t];
    cs->rank[t] = 0ull;
    if (main_state->is_zero[t]) continue;
    const unsigned long long r = main_state->rank[t];
    if (r < below || r - below >= ncand) fb = 1u;          // the sampled bracket missed the target
    else cs->rank[t] = r - below;
  }
  fs->fallback = fb;
}

// pack + (peer-memory all-reduce when row-sharded) + resolve in one launch
__device__ __forceinline__ void fast_pack_body(FastMedState* fs, const uint32_t* list_counts, uint32_t list_cap, uint32_t* cand_counts,
                                               uint32_t cand_cap, const uint32_t* err_flag, long long* red) {
  long long bad = *err_flag != 0u ? 1 : 0;
  unsigned long long ncand = 0ull, nlist = 0ull;
  for (int sgm = 0; sgm < MF_NSEG; ++sgm) {
    const uint32_t lc = list_counts[sgm * MF_CSTRIDE], cc = cand_counts[sgm * MF_CSTRIDE];
    if (lc > list_cap || cc > cand_cap) bad = 1;                   // a list overflowed (massive ties)
    if (cc > cand_cap) cand_counts[sgm * MF_CSTRIDE] = cand_cap;
    ncand += min(cc, cand_cap); nlist += lc;
  }
  fs->list_count = (uint32_t)min(nlist, 0xFFFFFFFFull);
  fs->cand_count = (uint32_t)min(ncand, 0xFFFFFFFFull);
  red[0] = (long long)fs->below; red[1] = (long long)ncand; red[2] = bad; red[3] = (long long)nlist;
}

__global__ void __launch_bounds__(64) fast_pack_resolve_kernel(const P2PComm c, int world, FastMedState* fs, const uint32_t* list_counts,
                                                               uint32_t list_cap, uint32_t* cand_counts, uint32_t cand_cap,
                                                               const uint32_t* err_flag, SelectState* main_state, SelectState* cs,
                                                               long long* red) {
  if (threadIdx.x == 0) fast_pack_body(fs, list_counts, list_cap, cand_counts, cand_cap, err_flag, red);
  __syncthreads();
  bool ok = true;
  if (world > 1) ok = p2p_allreduce_cta(c, reinterpret_cast<unsigned long long*>(red), 4, false);
  if (threadIdx.x == 0) {
    if (!ok) main_state->nan_flag = 1u;
    fast_resolve_body(fs, main_state, cs, red);
  }
}

// single-GPU pack + resolve for the large-d filter (median_large.cu)
int launch_fast_pack_resolve(cudaStream_t st, FastMedState* fs, const uint32_t* list_counts, uint32_t list_cap, uint32_t* cand_counts,
                             uint32_t cand_cap, const uint32_t* err_flag, SelectState* main_state, SelectState* cand_state,
                             long long* red);

__global__ void fast_commit_kernel(const FastMedState* fs, const SelectState* cs, SelectState* main_state) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (fs->fallback) return;     // the conditional SIMT passes will resolve main_state instead
  for (int t = 0; t < 2; ++t) main_state->value[t] = main_state->is_zero[t] ? 0.f : cs->value[t];
  main_state->done = 1;
}

int launch_fast_pack_resolve(cudaStream_t st, FastMedState* fs, const uint32_t* list_counts, uint32_t list_cap, uint32_t* cand_counts,
                             uint32_t cand_cap, const uint32_t* err_flag, SelectState* main_state, SelectState* cand_state,
                             long long* red) {
  P2PComm c{};
  c.rank = 0; c.world = 1;
  fast_pack_resolve_kernel<<<1, 64, 0, st>>>(c, 1, fs, list_counts, list_cap, cand_counts, cand_cap, err_flag, main_state, cand_state, red);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

// ---------------------------------------------------------------------------- fused 3-pass radix select over a value buffer
// ONE launch resolves both target order statistics of a buffer (the sampled distances -> bracket, the exact candidates ->
// result) instead of 3 x (memset + histogram + advance): per pass every CTA histograms its share in shared memory
// (with a warp-uniform fast path: the distances of one cloud share their leading radix digits, plain shared-memory atomics
// serialised on one or two bins -- 35 us per pass over 4M values), adds it to the pass's global histogram, and the LAST CTA to finish
// (atomic ticket) takes the select decision -- after summing the histogram over the ranks through peer memory when
// row-sharded -- and releases a flag the other CTAs of the (co-resident, <= #SM) grid wait for.
// final_mode: 1 -> publish the bracket [L, U] (bracket_set), 2 -> commit the resolved values to the main select state.
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256, 4) fused_select_kernel(const float* __restrict__ vals, const uint32_t* __restrict__ counts_dev,
                                                           int count_stride, int nseg, uint32_t seg_cap, uint32_t v_begin,
                                                           uint32_t v_end, SelectState* state, FusedSelScratch* sc, const P2PComm c,
                                                           int world, int final_mode, FastMedState* fs, SelectState* main_state,
                                                           int l_is_zero, int u_is_inf, const uint32_t* skip_flag) {
  if (skip_flag && *skip_flag) return;      // (identical on every CTA and every rank: nobody enters the hand-shakes)
  constexpr uint32_t NARROW_BINS = 8192;    // <= the 3 x 4096 counters of sc->hist, <= 2 x the words of a peer-memory slot
  __shared__ unsigned int hist_s[NARROW_BINS];
  __shared__ uint32_t s_last;
  const int lane = threadIdx.x & 31;
  if (final_mode == 2) {
    // NARROW BRACKET (a warm-started bracket is ~2^-12 wide): all candidates lie in [L, U] and the fp32 patterns of L and U are
    // fewer than 8192 apart, so ONE histogram over key = bits(v) - bits(L) resolves both order statistics exactly -- one pass
    // and one exchange instead of three.  L, U are identical on every rank, hence so is this decision.
    const uint32_t bl = __float_as_uint(fs->L), bu = __float_as_uint(fs->U);
    if (!fs->u_is_inf && !(bl >> 31) && bu >= bl && bu - bl < NARROW_BINS) {
      unsigned int* gh = &sc->hist[0][0];
      for (int t = threadIdx.x; t < (int)NARROW_BINS; t += 256) hist_s[t] = 0u;
      __syncthreads();
      const uint32_t G = min((uint32_t)nseg, gridDim.x);
      const uint32_t peers_n = (gridDim.x - (blockIdx.x % G) + G - 1u) / G;
      const uint32_t my_idx = blockIdx.x / G;
      for (int sgm = (int)(blockIdx.x % G); sgm < nseg; sgm += (int)G) {
        const uint32_t begin = counts_dev ? 0u : v_begin;
        const uint32_t end = counts_dev ? min(counts_dev[sgm * count_stride], seg_cap) : v_end;
        const float* v0 = vals + (size_t)sgm * seg_cap;
        for (uint32_t t = begin + my_idx * 256u + threadIdx.x; t < end; t += peers_n * 256u) {
          const uint32_t key = __float_as_uint(v0[t]) - bl;      // candidates are finite and inside [L, U]
          if (key < NARROW_BINS) atomicAdd(&hist_s[key], 1u);
        }
      }
      __syncthreads();
      for (int t = threadIdx.x; t < (int)NARROW_BINS; t += 256) {
        const unsigned int v = hist_s[t];
        if (v) atomicAdd(&gh[t], v);
      }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) s_last = (atomicAdd(&sc->done[0], 1u) == gridDim.x - 1u) ? 1u : 0u;
      __syncthreads();
      if (!s_last) return;                                       // single pass: nobody waits for the decision
      __threadfence();
      if (world > 1 && !p2p_allreduce_cta(c, reinterpret_cast<unsigned long long*>(gh), NARROW_BINS / 2, false) && threadIdx.x == 0)
        main_state->nan_flag = 1u;
      __syncthreads();
      for (int t = threadIdx.x; t < (int)NARROW_BINS; t += 256) hist_s[t] = gh[t];
      __syncthreads();
      // thread i owns bins [32 i, 32 i + 32): exclusive prefix over the 256 threads, then the owner walks its bins
      __shared__ unsigned long long warp_tot[8];
      unsigned long long mine = 0ull;
      for (int b = 0; b < 32; ++b) mine += hist_s[threadIdx.x * 32 + b];
      unsigned long long inc = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
      if (lane == 31) warp_tot[threadIdx.x >> 5] = inc;
      __syncthreads();
      unsigned long long excl = inc - mine;
      for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) excl += warp_tot[w];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (state->is_zero[t]) continue;
        const unsigned long long r = state->rank[t];
        if (r >= excl && r < excl + mine) {
          unsigned long long cum = excl;
          for (int b = 0; b < 32; ++b) {
            const unsigned long long cnt = hist_s[threadIdx.x * 32 + b];
            if (r < cum + cnt) { state->value[t] = __uint_as_float(bl + (uint32_t)(threadIdx.x * 32 + b)); break; }
            cum += cnt;
          }
        }
      }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0 && !fs->fallback) {                   // candidates resolved: commit (else the conditional SIMT passes do)
        for (int t = 0; t < 2; ++t) main_state->value[t] = main_state->is_zero[t] ? 0.f : state->value[t];
        main_state->done = 1;
      }
      return;
    }
  }
  for (int pass = 0; pass < NSVGD_SELECT_PASSES; ++pass) {
    for (int t = threadIdx.x; t < 2 * NSVGD_SELECT_BINS; t += 256) hist_s[t] = 0u;
    __syncthreads();
    const uint32_t mask = *(volatile uint32_t*)&state->mask;
    const uint32_t p0 = *(volatile uint32_t*)&state->prefix[0], p1 = *(volatile uint32_t*)&state->prefix[1];
    const int shift = pass_shift(pass);
    const unsigned int binmask = (unsigned int)pass_bins(pass) - 1u;
    // segmented buffers: the CTAs are dealt out over the segments (CTA b works on segments b % G, b % G + G, ... with its
    // G-group peers, G = min(nseg, grid)) so that every CTA has work whatever the segment sizes
    const uint32_t G = min((uint32_t)nseg, gridDim.x);
    const uint32_t peers_n = (gridDim.x - (blockIdx.x % G) + G - 1u) / G;      // CTAs that share my segments
    const uint32_t my_idx = blockIdx.x / G;
    for (int sgm = (int)(blockIdx.x % G); sgm < nseg; sgm += (int)G) {
      const uint32_t begin = counts_dev ? 0u : v_begin;
      const uint32_t end = counts_dev ? min(counts_dev[sgm * count_stride], seg_cap) : v_end;
      const float* v0 = vals + (size_t)sgm * seg_cap;
      // a warp takes 128 consecutive values per trip: four independent coalesced loads in flight per lane, then four
      // warp-aggregated histogram updates (warp-uniform trip count: the match needs the whole warp)
      for (uint32_t base = begin + (my_idx * 8u + (threadIdx.x >> 5)) * 128u; base < end; base += peers_n * 8u * 128u) {
        unsigned int bits4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t t = base + lane + 32u * u;
          bits4[u] = t < end ? __float_as_uint(v0[t]) : 0x7fc00000u;      // out of range -> NaN pattern -> ignored
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const unsigned int bits = bits4[u];
          const bool ok = (bits & 0x7fffffffu) <= 0x7f800000u;             // NaNs are reported by the producers of the buffer
          const unsigned int bin = (bits >> shift) & binmask;
          const bool m0 = ok && (bits & mask) == p0;
          const bool m1 = ok && (bits & mask) == p1;
          const unsigned int key = (m0 || m1) ? (bin | (m0 ? 0x1000u : 0u) | (m1 ? 0x2000u : 0u)) : 0x80000000u;
          // uniform fast path: the whole warp in one bin (the leading radix digits of one cloud's distances; plain
          // shared-memory atomics serialised on one or two bins here) -> ONE atomic; otherwise the lanes are spread out
          const unsigned int key0 = __shfl_sync(0xffffffffu, key, 0);
          if (__all_sync(0xffffffffu, key == key0)) {
            if (lane == 0 && (m0 || m1)) {
              if (m0) atomicAdd(&hist_s[bin], 32u);
              if (m1) atomicAdd(&hist_s[NSVGD_SELECT_BINS + bin], 32u);
            }
          } else {
            if (m0) atomicAdd(&hist_s[bin], 1u);
            if (m1) atomicAdd(&hist_s[NSVGD_SELECT_BINS + bin], 1u);
          }
        }
      }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * NSVGD_SELECT_BINS; t += 256) {
      const unsigned int v = hist_s[t];
      if (v) atomicAdd(&sc->hist[pass][t], v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&sc->done[pass], 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (world > 1 && !p2p_allreduce_cta(c, reinterpret_cast<unsigned long long*>(sc->hist[pass]), NSVGD_SELECT_BINS, false) && threadIdx.x == 0)
        (main_state ? main_state : state)->nan_flag = 1u;            // late peer -> h = NaN instead of a partial sum
      // stage the summed histogram in shared memory (independent coalesced loads; the counts of a value buffer fit 32 bits)
      // before the scan: 64 dependent L2 round trips per lane cost ~20 us per pass otherwise
      __syncthreads();
      for (int t = threadIdx.x; t < 2 * NSVGD_SELECT_BINS; t += 256) hist_s[t] = sc->hist[pass][t];
      __syncthreads();
      select_advance_smem32(state, hist_s, pass);
      if (pass == NSVGD_SELECT_PASSES - 1 && threadIdx.x == 0) {
        if (final_mode == 1) {             // bracket [L, U] for the tensor-core filter
          fs->below = 0ull; fs->list_count = 0u; fs->cand_count = 0u; fs->overflow = 0u; fs->fallback = 0u;
          fs->L = l_is_zero ? 0.f : state->value[0];
          fs->U = u_is_inf ? __uint_as_float(0x7f800000u) : state->value[1];
          fs->u_is_inf = (uint32_t)u_is_inf; fs->pad = 0u;
        } else if (final_mode == 2 && !fs->fallback) {   // candidates resolved: commit (else the conditional SIMT passes do)
          for (int t = 0; t < 2; ++t) main_state->value[t] = main_state->is_zero[t] ? 0.f : state->value[t];
          main_state->done = 1;
        }
      }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) st_release_gpu_u32(&sc->flag[pass], 1u);
    } else if (pass + 1 < NSVGD_SELECT_PASSES) {
      if (threadIdx.x == 0) {
        // back off between polls: ~600 CTAs re-reading one L2 line back to back slow down everything that shares the
        // slice -- the last CTA's own loads and, when row-sharded, the peers' reads of this rank's slot (measured:
        // 150 us per exchange at 8 GPUs without the back-off)
        const long long t0 = clock64();
        unsigned int ns = 100u;
        while (ld_acquire_gpu_u32(&sc->flag[pass]) == 0u) {
          __nanosleep(ns);
          if (ns < 1600u) ns <<= 1;
          if (clock64() - t0 > 8000000000ll) { (main_state ? main_state : state)->nan_flag = 1u; break; }   // never hang
        }
      }
      __syncthreads();
    }
  }
}

int launch_fused_select(cudaStream_t st, const float* vals, const uint32_t* counts_dev, int count_stride, int nseg,
                               uint32_t seg_cap, uint32_t v_begin, uint32_t v_end, SelectState* state, FusedSelScratch* sc,
                               const P2PComm* comm, int final_mode, FastMedState* fs, SelectState* main_state, int l_is_zero,
                               int u_is_inf, const uint32_t* skip_flag) {
  P2PComm c{};
  if (comm) c = *comm; else { c.rank = 0; c.world = 1; }
  // the grid MUST be co-resident (CTAs wait for each other): ask the occupancy calculator, at most four CTAs per SM
  static std::atomic<int> per_sm_cached{0};
  int per_sm = per_sm_cached.load();
  if (!per_sm) {
    NSVGD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_select_kernel, 256, 0));
    NSVGD_REQUIRE(per_sm >= 1, NSVGD_ERR_CUDA, "fused_select_kernel does not fit on an SM");
    per_sm = std::min(per_sm, 4);
    per_sm_cached.store(per_sm);
  }
  fused_select_kernel<<<device_sm_count() * per_sm, 256, 0, st>>>(vals, counts_dev, count_stride, nseg, seg_cap, v_begin, v_end, state, sc,
                                                         c, c.world, final_mode, fs, main_state, l_is_zero, u_is_inf, skip_flag);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

template <int DP>
static int launch_filter(cudaStream_t st, const CUtensorMap* maps, const MfParams& prm, int64_t row_blocks) {
  using L = MfSmem<DP>;
  NSVGD_PROPAGATE(ensure_dyn_smem((const void*)median_filter_tc_kernel<DP>, L::DYN_BYTES));
  const int64_t my_blocks = prm.rb_start < row_blocks ? ceil_div(row_blocks - prm.rb_start, prm.rb_stride) : 0;
  if (my_blocks == 0) return NSVGD_OK;
  const unsigned chunks = (unsigned)ceil_div(prm.n_tiles, MF_CHUNK);
  NSVGD_REQUIRE(my_blocks <= 65535, NSVGD_ERR_INVALID, "median filter: too many row blocks");
  prof_begin(6, st);
  median_filter_tc_kernel<DP><<<dim3(chunks, (unsigned)my_blocks), MF_THREADS, L::DYN_BYTES, st>>>(maps[0], maps[1], maps[2],
                                                                                                  maps[3], prm);
  prof_end(6, st);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

// stage 0: centred bf16 pieces, norms, thresholds, and this rank's slice of the sampled pair distances
// ---------------------------------------------------------------------------- warm-started bracket (MedianHint, select.cuh)
// hint valid -> the bracket is [lo (1 - delta), hi (1 + delta)] around the previous call's order statistics, the sample and
// the bracket select return at once (skip flag), and everything downstream is unchanged: the resolve step checks with exact
// counts that the target ranks fall among the candidates and triggers the exact SIMT select otherwise.
__global__ void hint_apply_kernel(const MedianHint* hint, FastMedState* fs, uint32_t* skip_flag, uint32_t key) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const bool use = hint && hint->valid && hint->key == key && hint->lo == hint->lo && hint->hi == hint->hi && hint->hi >= hint->lo && hint->lo >= 0.f;
  *skip_flag = use ? 1u : 0u;
  if (!use) return;
  fs->below = 0ull; fs->list_count = 0u; fs->cand_count = 0u; fs->overflow = 0u; fs->fallback = 0u;
  fs->L = hint->lo * (1.f - hint->delta);
  fs->U = hint->hi * (1.f + hint->delta);
  fs->u_is_inf = 0u; fs->pad = 0u;
}
// after the select: remember the order statistics.  margin = 4 x the relative change between the last two calls (>= 2^-13); the
// hint is only armed while that margin stays <= 2^-10 (about the width of the sampled bracket: a fast-moving cloud gains
// nothing from a wide warm-started bracket and would overflow the candidate lists), and a miss (the hinted bracket did not
// contain the target, so the exact SIMT fallback ran) disarms it for 16 << penalty calls, penalty growing with every further
// miss (<= 6) and shrinking with every hit: the amortised cost of misses stays bounded on any dynamics.
// `seen` packs: bit 0 = a previous value exists, bits 1-15 = remaining cool-down calls, bits 16-23 = penalty.
__global__ void hint_update_kernel(MedianHint* hint, const SelectState* main_state, const FastMedState* fs, const uint32_t* skip_flag,
                                   uint32_t key) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (hint->key != key) { hint->valid = 0u; hint->seen = 0u; hint->key = key; }      // another mode / shape wrote the old values
  uint32_t seen = hint->seen & 1u, cool = (hint->seen >> 1) & 0x7FFFu, pen = (hint->seen >> 16) & 0xFFu;
  const bool hinted = *skip_flag != 0u;
  const bool missed = hinted && fs->fallback;
  const float lo = main_state->value[0], hi = main_state->value[1];
  const bool usable = !main_state->nan_flag && !main_state->is_zero[0] && !main_state->is_zero[1] && lo == lo && hi == hi && hi > 0.f;
  if (hinted) { if (missed) ++hint->misses; else { ++hint->hits; if (pen) --pen; } }
  uint32_t valid = 0u;
  if (!usable) {
    seen = 0u;
  } else {
    float drift = 1.f;
    if (seen) {
      const float mid_old = 0.5f * (hint->lo + hint->hi), mid_new = 0.5f * (lo + hi);
      drift = fabsf(mid_new - mid_old) / mid_new;
    }
    hint->lo = lo; hint->hi = hi;                                  // exact either way (the fallback is exact too)
    if (missed) { pen = min(pen + 1u, 6u); cool = 16u << pen; }
    else if (cool) --cool;
    const float delta = fmaxf(4.f * drift, 1.220703125e-4f);       // >= 2^-13
    hint->delta = delta;
    valid = (seen && !cool && !missed && delta <= 9.765625e-4f) ? 1u : 0u;      // armed only while slow: margin <= 2^-10
    seen = 1u;
  }
  hint->valid = valid;
  hint->seen = seen | (cool << 1) | (pen << 16);
}

uint32_t median_hint_key(int64_t n, int64_t dz, long long r_lo, float w_last) {
  uint32_t key = 0x9E3779B9u, wb;
  memcpy(&wb, &w_last, 4);
  const uint64_t parts[4] = {(uint64_t)n, (uint64_t)dz, (uint64_t)r_lo, (uint64_t)wb};
  for (uint64_t v : parts) key ^= (uint32_t)(v ^ (v >> 32)) + 0x9E3779B9u + (key << 6) + (key >> 2);
  return key ? key : 1u;
}
int launch_hint_apply(cudaStream_t st, const MedianHint* hint, FastMedState* fs, uint32_t* skip_flag, uint32_t key) {
  hint_apply_kernel<<<1, 32, 0, st>>>(hint, fs, skip_flag, key);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}
int launch_hint_update(cudaStream_t st, MedianHint* hint, const SelectState* main_state, const FastMedState* fs,
                       const uint32_t* skip_flag, uint32_t key) {
  hint_update_kernel<<<1, 32, 0, st>>>(hint, main_state, fs, skip_flag, key);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

// replicated = true: every rank computes the SAME reduced sample set of MF_SAMPLES / nparts pairs (rows spread over all n),
// so the bracket select needs no exchange at all (median_fast_run)
int median_fast_sample(cudaStream_t st, const float* P, int64_t n, int64_t dz, int64_t ldp, float w_last, int64_t part,
                       int64_t nparts, void* ws, size_t ws_bytes, size_t* samples_off, int64_t* samples_count, bool replicated,
                       const MedianHint* hint = nullptr, uint32_t hint_key = 0u) {
  NSVGD_REQUIRE(median_fast_supported(n, dz), NSVGD_ERR_UNSUPPORTED, "median_fast: unsupported shape");
  NSVGD_REQUIRE(nparts >= 1 && part >= 0 && part < nparts && MF_SAMPLES % nparts == 0 && (MF_SAMPLES / nparts) % 32 == 0,
                NSVGD_ERR_INVALID, "median_fast_sample: the sample count %u must split evenly over %lld parts", MF_SAMPLES,
                (long long)nparts);
  WsCarver w(ws, ws_bytes);
  FastWs f;
  fast_layout(n, dz, w, &f);
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "median_fast: workspace too small (%zu < %zu)", ws_bytes, w.off);
  const int sms = device_sm_count();
  prof_begin(8, st);
  NSVGD_PROPAGATE(launch_colmean(st, P, n, dz, ldp, f.mpart, f.mu));
  NSVGD_CUDA(cudaMemsetAsync(f.zero_block, 0, f.zero_bytes, st));   // error flag, list / candidate counters, select scratch
  uint32_t* skip_flag = f.err_flag + 2;                             // (inside the zero block: 0 unless a valid hint sets it)
  if (hint) {
    hint_apply_kernel<<<1, 32, 0, st>>>(hint, f.fs, skip_flag, hint_key);
    NSVGD_LAUNCH_CHECK();
  }
  // weighted last coordinate scaled by sqrt(w) AFTER centring
  NSVGD_PROPAGATE(launch_tc_prep(st, P, nullptr, n, dz, ldp, f.n_pad, f.dp, f.mu, sqrtf(w_last), f.Zb0, f.Zb1, nullptr, nullptr,
                                 nullptr, 0, f.zn, nullptr, nullptr, nullptr));
  filter_vectors_kernel<<<(unsigned)std::min<int64_t>(ceil_div(f.n_pad, 256), 1024), 256, 0, st>>>(f.zn, f.n_pad, f.zlo, f.zhi);
  NSVGD_LAUNCH_CHECK();
  prof_end(8, st);
  constexpr int WX_SMEM = WX_WARPS * 32 * WX_STRIDE * (int)sizeof(float);
  NSVGD_PROPAGATE(ensure_dyn_smem((const void*)sample_pairs_warp_kernel, WX_SMEM));
  NSVGD_PROPAGATE(ensure_dyn_smem((const void*)exact_list_warp_kernel, WX_SMEM));
  const uint32_t per = MF_SAMPLES / (uint32_t)nparts;
  const uint32_t s_begin = replicated ? 0u : per * (uint32_t)part, s_end = s_begin + per;
  const uint32_t nb_total = (replicated ? per : MF_SAMPLES) / 32u;
  prof_begin(4, st);
  const char* iid_env = getenv("NSVGD_MEDIAN_IID_SAMPLE");
  if (iid_env && iid_env[0] == '1') {     // A/B: the first version, i.i.d. pairs with defined-order arithmetic
    const unsigned blocks = (unsigned)std::min<int64_t>(sms * 3, ceil_div(per, 32 * WX_WARPS));
    sample_pairs_warp_kernel<<<blocks, 32 * WX_WARPS, WX_SMEM, st>>>(P, n, (int)dz, ldp, w_last, s_begin, s_end, f.Ds);
  } else {
    const unsigned blocks = (unsigned)std::min<int64_t>(sms * 8, ceil_div(per, 32 * WX_WARPS));
    sample_rows_kernel<<<blocks, 32 * WX_WARPS, 0, st>>>(P, n, (int)dz, ldp, w_last, s_begin, s_end, f.Ds, nb_total, skip_flag);
  }
  prof_end(4, st);
  NSVGD_LAUNCH_CHECK();
  if (samples_off) *samples_off = (size_t)((char*)f.Ds - (char*)ws);
  if (samples_count) *samples_count = (int64_t)MF_SAMPLES;
  return NSVGD_OK;
}

// stage 1 (all MF_SAMPLES sampled distances present in the workspace): bracket, filter, exact list, summary
// comm != null: row-sharded over peer memory -- the bracket select runs over THIS rank's sample slice [s_begin, s_end) and
// sums the histograms across ranks inside the kernel; red (the local summary) is then written by the resolve kernel.
int median_fast_stage1(cudaStream_t st, const float* P, int64_t n, int64_t dz, int64_t ldp, float w_last, long long r_lo,
                       long long r_hi, int64_t block_row_start, int64_t block_row_stride, SelectState* main_state,
                       uint64_t* hist, void* ws, size_t ws_bytes, long long* red, const P2PComm* comm, uint32_t s_begin,
                       uint32_t s_end, uint32_t ns_total = MF_SAMPLES, bool bracket_local = false) {
  // ns_total: size of the sample set the bracket ranks refer to; bracket_local: the set [s_begin, s_end) = [0, ns_total) is
  // replicated on every rank, so the bracket select runs without an exchange
  NSVGD_REQUIRE(median_fast_supported(n, dz), NSVGD_ERR_UNSUPPORTED, "median_fast: unsupported shape");
  NSVGD_REQUIRE(r_lo >= 0 && r_hi >= r_lo, NSVGD_ERR_UNSUPPORTED, "median_fast: needs genuine pair ranks");
  WsCarver w(ws, ws_bytes);
  FastWs f;
  fast_layout(n, dz, w, &f);
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "median_fast: workspace too small (%zu < %zu)", ws_bytes, w.off);
  const int vec = (ldp % 4 == 0) && (((uintptr_t)P & 15) == 0);
  const int sms = device_sm_count();
  constexpr int WX_SMEM = WX_WARPS * 32 * WX_STRIDE * (int)sizeof(float);

  // --- 1. bracket [L, U] from the sample quantiles: 6 sigma of the binomial rank error + slack
  //        (deterministic pseudo-random pairs: every rank computes the identical bracket)
  const double Ppairs = 0.5 * (double)n * (double)(n - 1);
  const double ns = (double)ns_total;
  const double p_lo = ((double)r_lo + 0.5) / Ppairs, p_hi = ((double)r_hi + 0.5) / Ppairs;
  const double sig_lo = sqrt(ns * p_lo * (1.0 - p_lo)), sig_hi = sqrt(ns * p_hi * (1.0 - p_hi));
  // 5 sigma + slack on either side: a miss (probability ~6e-7 per step) only costs the device-side fallback to the SIMT
  // select (still exact); the bracket is 17 % narrower than at 6 sigma and so is the list of pairs to re-evaluate
  long long a = (long long)floor(p_lo * ns - 5.0 * sig_lo - 16.0);
  long long b = (long long)ceil(p_hi * ns + 5.0 * sig_hi + 16.0);
  const int l_is_zero = a < 0, u_is_inf = b >= (long long)ns_total;
  if (a < 0) a = 0;
  if (b >= (long long)ns_total) b = (long long)ns_total - 1;
  (void)hist;
  prof_begin(9, st);
  NSVGD_PROPAGATE(launch_select_init(st, f.bracket, a, b, 1.f));
  NSVGD_PROPAGATE(launch_fused_select(st, f.Ds, nullptr, 0, 1, 0, s_begin, s_end, f.bracket, &f.sel[0], bracket_local ? nullptr : comm, 1,
                                      f.fs, nullptr, l_is_zero, u_is_inf, f.err_flag + 2));
  prof_end(9, st);

  // --- 2. tensor-core filter over the upper triangle (128-row blocks strided across ranks)
  CUtensorMap maps[4];
  NSVGD_PROPAGATE(make_tmap_bf16_2d(&maps[0], f.Zb0, f.dp, f.n_pad, (uint64_t)f.dp * 2, 64, MF_BM));
  NSVGD_PROPAGATE(make_tmap_bf16_2d(&maps[1], f.Zb1, f.dp, f.n_pad, (uint64_t)f.dp * 2, 64, MF_BM));
  NSVGD_PROPAGATE(make_tmap_bf16_2d(&maps[2], f.Zb0, f.dp, f.n_pad, (uint64_t)f.dp * 2, 64, MF_BN));
  NSVGD_PROPAGATE(make_tmap_bf16_2d(&maps[3], f.Zb1, f.dp, f.n_pad, (uint64_t)f.dp * 2, 64, MF_BN));
  MfParams prm{};
  prm.Qb0 = (const uint32_t*)f.Zb0; prm.Qb1 = (const uint32_t*)f.Zb1;
  prm.zn = f.zn; prm.zlo = f.zlo; prm.zhi = f.zhi; prm.n = n; prm.n_tiles = (int)(f.n_pad / MF_BN); prm.rb_start = block_row_start; prm.rb_stride = block_row_stride;
  prm.fs = f.fs; prm.list = f.list; prm.list_counts = f.list_counts; prm.list_cap = f.list_cap; prm.err_flag = f.err_flag;
  const int64_t row_blocks = f.n_pad / MF_BM;
  if (f.dp == 64) NSVGD_PROPAGATE(launch_filter<64>(st, maps, prm, row_blocks));
  else NSVGD_PROPAGATE(launch_filter<128>(st, maps, prm, row_blocks));

  // --- 3. exact evaluation of the uncertain pairs
  prof_begin(5, st);
  // listed pairs come out of the filter grouped by row quarter and tile, so the one-thread-per-pair
  // gather re-uses rows through L1 and beats the warp-cooperative variant here (1.85 vs 2.5 ms at cfg4)
  const char* xl_env = getenv("NSVGD_MEDIAN_EXACT");        // "thread": the one-thread-per-pair kernel (A/B measurements)
  if (vec && dz % 4 == 0 && !(xl_env && xl_env[0] == 't')) {
    NSVGD_PROPAGATE(ensure_dyn_smem((const void*)exact_list_coop_kernel, XL_SMEM));
    exact_list_coop_kernel<<<dim3((unsigned)std::max(1, sms * 8 / MF_NSEG), MF_NSEG), 256, XL_SMEM, st>>>(
        P, (int)dz, ldp, w_last, f.fs, f.list, f.list_counts, f.list_cap, f.cand, f.cand_counts, f.cand_cap, main_state);
  } else if (vec) {
    exact_list_kernel<<<dim3((unsigned)std::max(1, sms * 8 / MF_NSEG), MF_NSEG), 256, 0, st>>>(
        P, (int)dz, ldp, w_last, vec, f.fs, f.list, f.list_counts, f.list_cap, f.cand, f.cand_counts, f.cand_cap, main_state);
  } else {
    exact_list_warp_kernel<<<dim3((unsigned)std::max(1, sms * 3 / MF_NSEG + 1), MF_NSEG), 32 * WX_WARPS, WX_SMEM, st>>>(
        P, (int)dz, ldp, w_last, f.fs, f.list, f.list_counts, f.list_cap, f.cand, f.cand_counts, f.cand_cap, main_state);
  }
  prof_end(5, st);
  NSVGD_LAUNCH_CHECK();
  if (red) {       // staged use with an external all-reduce of the summary in between
    fast_pack_kernel<<<1, 32, 0, st>>>(f.fs, f.list_counts, f.list_cap, f.cand_counts, f.cand_cap, f.err_flag, red);
    NSVGD_LAUNCH_CHECK();
  }
  return NSVGD_OK;
}

static int fast_ws(int64_t n, int64_t dz, void* ws, size_t ws_bytes, FastWs* f) {
  NSVGD_REQUIRE(median_fast_supported(n, dz), NSVGD_ERR_UNSUPPORTED, "median_fast: unsupported shape");
  WsCarver w(ws, ws_bytes);
  fast_layout(n, dz, w, f);
  NSVGD_REQUIRE(ws && w.ok(), NSVGD_ERR_WORKSPACE, "median_fast: workspace too small (%zu < %zu)", ws_bytes, w.off);
  return NSVGD_OK;
}

int median_fast_resolve(cudaStream_t st, SelectState* main_state, void* ws, size_t ws_bytes, int64_t n, int64_t dz,
                        const long long* red) {
  FastWs f;
  NSVGD_PROPAGATE(fast_ws(n, dz, ws, ws_bytes, &f));
  fast_resolve_kernel<<<1, 32, 0, st>>>(f.fs, main_state, f.cand_state, red);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

int median_fast_cand_hist(cudaStream_t st, void* ws, size_t ws_bytes, int64_t n, int64_t dz, int pass, uint64_t* hist) {
  FastWs f;
  NSVGD_PROPAGATE(fast_ws(n, dz, ws, ws_bytes, &f));
  return launch_buffer_hist(st, f.cand, f.cand_counts, MF_CSTRIDE, MF_NSEG, f.cand_cap, 0, f.cand_state, pass, hist, nullptr);
}

int median_fast_cand_advance(cudaStream_t st, void* ws, size_t ws_bytes, int64_t n, int64_t dz, const uint64_t* hist, int pass) {
  FastWs f;
  NSVGD_PROPAGATE(fast_ws(n, dz, ws, ws_bytes, &f));
  return launch_select_advance(st, f.cand_state, hist, pass, nullptr);
}

int median_fast_commit(cudaStream_t st, SelectState* main_state, void* ws, size_t ws_bytes, int64_t n, int64_t dz) {
  FastWs f;
  NSVGD_PROPAGATE(fast_ws(n, dz, ws, ws_bytes, &f));
  fast_commit_kernel<<<1, 32, 0, st>>>(f.fs, f.cand_state, main_state);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

const uint32_t* median_fast_fallback_flag(void* ws, size_t ws_bytes, int64_t n, int64_t dz) {
  FastWs f;
  if (fast_ws(n, dz, ws, ws_bytes, &f) != NSVGD_OK) return nullptr;
  return &f.fs->fallback;
}

// the whole filtered select for one rank: comm == null -> single GPU; else row-sharded over peer memory (128-row blocks
// rank, rank + world, ... of the filter; sample slice rank of world; every exchange fused into its consumer kernel).
// 13 launches + 1 memset (+ 9 early-exit launches of the conditional SIMT fallback) instead of ~45.
int median_fast_run(cudaStream_t st, const float* P, int64_t n, int64_t dz, int64_t ldp, float w_last, long long r_lo,
                    long long r_hi, const P2PComm* comm, SelectState* main_state, uint64_t* hist, void* ws, size_t ws_bytes,
                    MedianHint* hint = nullptr) {
  FastWs f;
  NSVGD_PROPAGATE(fast_ws(n, dz, ws, ws_bytes, &f));
  long long* red = (long long*)(f.err_flag) + 8;      // err_flag block is 256 bytes, 16 used
  const int rank = comm ? comm->rank : 0, world = comm ? comm->world : 1;
  const bool slice = comm && MF_SAMPLES % (uint32_t)world == 0 && (MF_SAMPLES / (uint32_t)world) % 32 == 0;
  const uint32_t per = slice ? MF_SAMPLES / (uint32_t)world : MF_SAMPLES;
  // NSVGD_MEDIAN_REPLICATED_SAMPLE=1 (row-sharded, off by default): every rank computes the SAME reduced sample set of
  // 2^22 / world pairs and selects the bracket locally -- no exchange in the bracket phase (three grid-wide passes with a
  // peer-memory all-reduce each).  The bracket is sqrt(world) x wider, so the exact list grows.  Measured at 2 ranks: bracket
  // 0.153 -> 0.064 ms, exact list 0.225 -> 0.309 ms, step 3.415 -> 3.439 ms; extrapolated to 8 ranks: -0.20 + 0.15 ms.
  // Bit-identical h either way (tests/test_sharded_gpu.py).
  bool replicated = false;
  if (const char* e = getenv("NSVGD_MEDIAN_REPLICATED_SAMPLE")) replicated = slice && e[0] == '1';
  // which problem a hint belongs to: shape, target rank (differs between the median modes), last-coordinate weight
  const uint32_t hint_key = median_hint_key(n, dz, r_lo, w_last);
  const uint32_t s_begin = (slice && !replicated) ? per * (uint32_t)rank : 0u, s_end = s_begin + per;
  NSVGD_PROPAGATE(median_fast_sample(st, P, n, dz, ldp, w_last, slice ? rank : 0, slice ? world : 1, ws, ws_bytes, nullptr, nullptr,
                                     replicated, hint, hint_key));
  // (an unsliceable world size: every rank samples everything and contributes 1/world-th... not needed: world | 2^22 in practice)
  NSVGD_REQUIRE(!comm || slice, NSVGD_ERR_UNSUPPORTED, "median_fast_run: world size %d does not divide the sample count", world);
  NSVGD_PROPAGATE(median_fast_stage1(st, P, n, dz, ldp, w_last, r_lo, r_hi, rank, world, main_state, hist, ws, ws_bytes, nullptr,
                                     comm, s_begin, s_end, replicated ? per : MF_SAMPLES, replicated));
  P2PComm c{};
  if (comm) c = *comm; else { c.rank = 0; c.world = 1; }
  prof_begin(10, st);
  fast_pack_resolve_kernel<<<1, 64, 0, st>>>(c, world, f.fs, f.list_counts, f.list_cap, f.cand_counts, f.cand_cap, f.err_flag,
                                             main_state, f.cand_state, red);
  NSVGD_LAUNCH_CHECK();
  NSVGD_PROPAGATE(launch_fused_select(st, f.cand, f.cand_counts, MF_CSTRIDE, MF_NSEG, f.cand_cap, 0u, 0u, f.cand_state, &f.sel[1], comm,
                                      2, f.fs, main_state, 0, 0));
  prof_end(10, st);
  prof_begin(11, st);
  // conditional fallback: the SIMT 3-pass select; kernels return at once unless fs->fallback is set (identical on all ranks)
  for (int pass = 0; pass < NSVGD_SELECT_PASSES; ++pass) {
    NSVGD_PROPAGATE(launch_pair_hist(st, P, n, dz, ldp, w_last, rank, world, main_state, pass, hist, &f.fs->fallback));
    if (comm) NSVGD_PROPAGATE(launch_select_advance_p2p(st, *comm, main_state, hist, pass, &f.fs->fallback));
    else NSVGD_PROPAGATE(launch_select_advance(st, main_state, hist, pass, &f.fs->fallback));
  }
  prof_end(11, st);
  if (hint) {
    hint_update_kernel<<<1, 32, 0, st>>>(hint, main_state, f.fs, f.err_flag + 2, hint_key);
    NSVGD_LAUNCH_CHECK();
  }
  return NSVGD_OK;
}

int median_fast(cudaStream_t st, const float* P, int64_t n, int64_t dz, int64_t ldp, float w_last, long long r_lo,
                long long r_hi, int64_t tile_row_start, int64_t tile_row_stride, SelectState* main_state, uint64_t* hist,
                void* ws, size_t ws_bytes, MedianHint* hint) {
  (void)tile_row_start; (void)tile_row_stride;
  return median_fast_run(st, P, n, dz, ldp, w_last, r_lo, r_hi, nullptr, main_state, hist, ws, ws_bytes, hint);
}

}  // namespace nsvgd

using namespace nsvgd;

// ---- C ABI: building blocks of the tensor-core filtered select for row-sharded runs (include/nsvgd.h)
extern "C" int nsvgd_median_fast_supported(int64_t n, int64_t dz) { return median_fast_supported(n, dz) ? 1 : 0; }

extern "C" int nsvgd_median_fast_workspace_bytes(int64_t n, int64_t dz, size_t* out_bytes) {
  NSVGD_REQUIRE(out_bytes, NSVGD_ERR_INVALID, "median_fast_workspace_bytes: null pointer");
  NSVGD_REQUIRE(median_fast_supported(n, dz), NSVGD_ERR_UNSUPPORTED, "median_fast: unsupported shape (n=%lld dz=%lld)", (long long)n, (long long)dz);
  *out_bytes = median_fast_workspace(n, dz);
  return NSVGD_OK;
}

extern "C" int nsvgd_median_fast_sample(void* stream, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last,
                                        int64_t part, int64_t nparts, void* ws, size_t ws_bytes, size_t* samples_offset_bytes,
                                        int64_t* samples_count) {
  NSVGD_REQUIRE(Z && ws, NSVGD_ERR_INVALID, "median_fast_sample: null pointer");
  return median_fast_sample((cudaStream_t)stream, Z, n, dz, ldz, w_last, part, nparts, ws, ws_bytes, samples_offset_bytes,
                            samples_count, false);
}

extern "C" int nsvgd_median_fast_stage1(void* stream, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last, int mode,
                                        int64_t block_row_start, int64_t block_row_stride, void* state_dev, uint64_t* hist_dev,
                                        void* ws, size_t ws_bytes, int64_t* reduce_dev) {
  NSVGD_REQUIRE(Z && state_dev && hist_dev && ws && reduce_dev, NSVGD_ERR_INVALID, "median_fast_stage1: null pointer");
  NSVGD_REQUIRE(mode >= 0 && mode <= 2 && block_row_start >= 0 && block_row_stride >= 1, NSVGD_ERR_INVALID, "median_fast_stage1: bad arguments");
  long long r_lo, r_hi;
  median_ranks_host(n, mode, &r_lo, &r_hi);
  return median_fast_stage1((cudaStream_t)stream, Z, n, dz, ldz, w_last, r_lo, r_hi, block_row_start, block_row_stride,
                            (SelectState*)state_dev, hist_dev, ws, ws_bytes, (long long*)reduce_dev, nullptr, 0u, MF_SAMPLES);
}
extern "C" int nsvgd_p2p_median_fast_run_hinted(void* stream, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last, int mode,
                                                int rank, int world, void* const* blocks, void* state_dev, uint64_t* hist_dev,
                                                void* hint_dev, void* ws, size_t ws_bytes);
extern "C" int nsvgd_p2p_median_fast_run(void* stream, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last, int mode,
                                         int rank, int world, void* const* blocks, void* state_dev, uint64_t* hist_dev, void* ws,
                                         size_t ws_bytes) {
  return nsvgd_p2p_median_fast_run_hinted(stream, Z, n, dz, ldz, w_last, mode, rank, world, blocks, state_dev, hist_dev, nullptr, ws,
                                          ws_bytes);
}
extern "C" int nsvgd_p2p_median_fast_run_hinted(void* stream, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last, int mode,
                                                int rank, int world, void* const* blocks, void* state_dev, uint64_t* hist_dev,
                                                void* hint_dev, void* ws, size_t ws_bytes) {
  NSVGD_REQUIRE(Z && state_dev && hist_dev && ws, NSVGD_ERR_INVALID, "p2p_median_fast_run: null pointer");
  NSVGD_REQUIRE(mode >= 0 && mode <= 2, NSVGD_ERR_INVALID, "p2p_median_fast_run: bad mode %d", mode);
  P2PComm c;
  NSVGD_REQUIRE(p2p_make_comm(&c, rank, world, blocks), NSVGD_ERR_INVALID, "p2p_median_fast_run: bad communicator");
  long long r_lo, r_hi;
  median_ranks_host(n, mode, &r_lo, &r_hi);
  return median_fast_run((cudaStream_t)stream, Z, n, dz, ldz, w_last, r_lo, r_hi, world > 1 ? &c : nullptr, (SelectState*)state_dev,
                         hist_dev, ws, ws_bytes, (MedianHint*)hint_dev);
}

extern "C" int nsvgd_median_fast_resolve(void* stream, void* state_dev, void* ws, size_t ws_bytes, int64_t n, int64_t dz,
                                         const int64_t* reduce_dev) {
  NSVGD_REQUIRE(state_dev && reduce_dev, NSVGD_ERR_INVALID, "median_fast_resolve: null pointer");
  return median_fast_resolve((cudaStream_t)stream, (SelectState*)state_dev, ws, ws_bytes, n, dz, (const long long*)reduce_dev);
}
extern "C" int nsvgd_p2p_median_fast_resolve(void* stream, int rank, int world, void* const* blocks, void* state_dev, void* ws,
                                             size_t ws_bytes, int64_t n, int64_t dz, int64_t* reduce_dev) {
  P2PComm c;
  NSVGD_REQUIRE(p2p_make_comm(&c, rank, world, blocks), NSVGD_ERR_INVALID, "p2p_median_fast_resolve: bad communicator");
  NSVGD_REQUIRE(state_dev && reduce_dev, NSVGD_ERR_INVALID, "p2p_median_fast_resolve: null pointer");
  FastWs f;
  NSVGD_PROPAGATE(fast_ws(n, dz, ws, ws_bytes, &f));
  fast_resolve_p2p_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(c, f.fs, (const SelectState*)state_dev, f.cand_state,
                                                             (long long*)reduce_dev);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

/* hist_dev: uint64[4 * NSVGD_SELECT_BINS] = [candidate histogram (2 targets) | fallback histogram (2 targets)] */
extern "C" int nsvgd_p2p_median_fast_advance2(void* stream, int rank, int world, void* const* blocks, void* state_dev, void* ws,
                                              size_t ws_bytes, int64_t n, int64_t dz, uint64_t* hist_dev, int pass) {
  P2PComm c;
  NSVGD_REQUIRE(p2p_make_comm(&c, rank, world, blocks), NSVGD_ERR_INVALID, "p2p_median_fast_advance2: bad communicator");
  NSVGD_REQUIRE(state_dev && hist_dev, NSVGD_ERR_INVALID, "p2p_median_fast_advance2: null pointer");
  NSVGD_REQUIRE(pass >= 0 && pass < NSVGD_SELECT_PASSES, NSVGD_ERR_INVALID, "p2p_median_fast_advance2: bad pass %d", pass);
  FastWs f;
  NSVGD_PROPAGATE(fast_ws(n, dz, ws, ws_bytes, &f));
  fast_advance2_p2p_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(c, f.fs, f.cand_state, (SelectState*)state_dev,
                                                                (unsigned long long*)hist_dev, pass);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

extern "C" int nsvgd_median_fast_cand_hist(void* stream, void* ws, size_t ws_bytes, int64_t n, int64_t dz, int pass, uint64_t* hist_dev) {
  NSVGD_REQUIRE(hist_dev && pass >= 0 && pass < NSVGD_SELECT_PASSES, NSVGD_ERR_INVALID, "median_fast_cand_hist: bad arguments");
  return median_fast_cand_hist((cudaStream_t)stream, ws, ws_bytes, n, dz, pass, hist_dev);
}
extern "C" int nsvgd_median_fast_cand_advance(void* stream, void* ws, size_t ws_bytes, int64_t n, int64_t dz, const uint64_t* hist_dev, int pass) {
  NSVGD_REQUIRE(hist_dev && pass >= 0 && pass < NSVGD_SELECT_PASSES, NSVGD_ERR_INVALID, "median_fast_cand_advance: bad arguments");
  return median_fast_cand_advance((cudaStream_t)stream, ws, ws_bytes, n, dz, hist_dev, pass);
}
extern "C" int nsvgd_median_fast_commit(void* stream, void* state_dev, void* ws, size_t ws_bytes, int64_t n, int64_t dz) {
  NSVGD_REQUIRE(state_dev, NSVGD_ERR_INVALID, "median_fast_commit: null pointer");
  return median_fast_commit((cudaStream_t)stream, (SelectState*)state_dev, ws, ws_bytes, n, dz);
}
extern "C" int nsvgd_median_fallback_hist(void* stream, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last,
                                          int64_t tile_row_start, int64_t tile_row_stride, const void* state_dev, void* ws,
                                          size_t ws_bytes, int pass, uint64_t* hist_dev) {
  NSVGD_REQUIRE(Z && state_dev && hist_dev && pass >= 0 && pass < NSVGD_SELECT_PASSES, NSVGD_ERR_INVALID, "median_fallback_hist: bad arguments");
  const uint32_t* flag = median_fast_fallback_flag(ws, ws_bytes, n, dz);
  NSVGD_REQUIRE(flag, NSVGD_ERR_WORKSPACE, "median_fallback_hist: bad workspace");
  return launch_pair_hist((cudaStream_t)stream, Z, n, dz, ldz, w_last, tile_row_start, tile_row_stride,
                          (const SelectState*)state_dev, pass, hist_dev, flag);
}
extern "C" int nsvgd_median_fallback_advance(void* stream, void* state_dev, void* ws, size_t ws_bytes, int64_t n, int64_t dz,
                                             const uint64_t* hist_dev, int pass) {
  NSVGD_REQUIRE(state_dev && hist_dev && pass >= 0 && pass < NSVGD_SELECT_PASSES, NSVGD_ERR_INVALID, "median_fallback_advance: bad arguments");
  const uint32_t* flag = median_fast_fallback_flag(ws, ws_bytes, n, dz);
  NSVGD_REQUIRE(flag, NSVGD_ERR_WORKSPACE, "median_fallback_advance: bad workspace");
  return launch_select_advance((cudaStream_t)stream, (SelectState*)state_dev, hist_dev, pass, flag);
}
