// phi_tc.cu — fused flash-style phi* on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), d <= 128.
// [stein.py:180-222 + kernels.py:52-65; closed forms in SURVEY.md 8a]
//
// Per CTA: 128 query rows x a range of 64-wide tiles of inducing particles j.  For every tile
//   (1) Gram   G = Qc Xc^T                 tcgen05.mma SS, M=128 N=64, operands TMA-staged (SWIZZLE_128B)
//   (2) k_ij = exp2(2c g - c|q_i|^2 - c|x_j|^2), c = log2(e)/(2h^2), in registers: one thread per
//       row (tcgen05.ld 32x32b), MUFU.EX2; rowsum, sum k c_j, sum k r2 ride along in registers
//   (3) O += P [S | Xc]                     tcgen05.mma TS (P from TMEM), M=128 N=2*DP, accumulators in TMEM
// The n x n kernel matrix is never materialised.  fp32 accuracy on bf16 tensor cores comes from a
// 2-piece split of every operand (x = b0 + b1, |x - b0 - b1| <= 2^-18 |x|) and 3 MMA passes
// (b0*b0' + b0*b1' + b1*b0'), the same error-compensation scheme as 3xTF32 at half the shared-memory
// and TMEM footprint and twice the MMA rate (DESIGN.md "why bf16 pieces").
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 2-5 = pair-function / epilogue warps (TMEM lane quarter = warp_idx % 4).
// TMEM map (512 cols): O [0,256) | Q piece b0 [256,320) | Q piece b1 [320,384) | G [384,448) | P [448,512)
// The query block lives in TENSOR MEMORY as the A operand of the Gram MMAs (tcgen05.mma TS form) instead
// of shared memory: the SM's 128 B/clk shared-memory port was the binding resource (Q re-read by every
// Gram MMA = 96 KB of the 336 KB moved per tile; profiles/r01_phi_tc_notes.txt).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <cuda_fp16.h>

#include "select.cuh"
#include "tc_common.cuh"

namespace nsvgd {

using namespace tc;

// ---------------------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !p) {
    (void)cudaGetLastError();
    return nullptr;
  }
  fn = (EncodeTiledFn)p;
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  NSVGD_REQUIRE(fn, NSVGD_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NSVGD_REQUIRE(r == CUDA_SUCCESS, NSVGD_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return NSVGD_OK;
}

// ---------------------------------------------------------------------------- prep (K0)
// Per 64-row block: centre, split into bf16 pieces, row norms, c_j = s_j . xc_j, and the
// transposed [S | Xc] pieces (K-major B operand of the second GEMM) via a shared-memory transpose.
//   Xb0/Xb1 : [rows_pad x DP] bf16        Vt0/Vt1 : [2*DP x rows_pad] bf16 (rows 0..DP-1 = S^T, DP.. = Xc^T)
//   Xh      : [rows_pad x DP] fp16, the single-piece operand of the one-pass Gram (phi_tcw.cu); values below the fp16
//             normal range are stored as 0 so that no subnormal ever reaches the tensor core
//   xn      : |xc_j|^2 (1e30 for padded rows, which zeroes k)      cv : s_j . xc_j
//   max_bits: [0] = max_j |xc_j|^2, [1] = max_j |xc_j - fp16(xc_j)|^2 (float bits, atomicMax): the inputs of the
//             rigorous error bound that decides between one and three Gram passes on the device
// COMBINED form (h_combine != null): one transposed block W^T = (S - Xc/h^2)^T in rows [0, dp) of Vt
// (phi_i = (1/n)[(K W)_i + rowsum_i qc_i / h^2]: the second GEMM has N = dp instead of 2 dp columns); with
// vt_extra = 16 the rows [dp, dp + 16) hold {1, c_j, 0, ...}: rowsum_i = (K 1)_i and (K c)_i come out of the same MMAs.
__global__ void __launch_bounds__(256, 2) tc_prep_kernel(const float* __restrict__ X, const float* __restrict__ S,
                                                      int64_t n, int64_t d, int64_t ldx, int64_t rows_pad, int dp,
                                                      const float* __restrict__ mu, float last_scale,
                                                      __nv_bfloat16* __restrict__ Xb0, __nv_bfloat16* __restrict__ Xb1,
                                                      __half* __restrict__ Xh,
                                                      __nv_bfloat16* __restrict__ Vt0, __nv_bfloat16* __restrict__ Vt1,
                                                      int vt_extra, float* __restrict__ xn, float* __restrict__ cv,
                                                      const float* __restrict__ h_combine, uint32_t* __restrict__ max_bits) {
  __shared__ float T[64][129];   // one half (S or Xc) of the block, fp32
  __shared__ float cvs[64], ones[64];
  __shared__ unsigned int smax[2];          // per-CTA maxima: 65536 x 2 same-address global atomics cost ~90 us otherwise
  if (threadIdx.x < 2) smax[threadIdx.x] = 0u;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * 64;
  const bool combine = h_combine != nullptr;
  float inv_h2 = 0.f;
  if (combine) { const float h = *h_combine; inv_h2 = 1.f / (h * h); }
  for (int half = 0; half < 2; ++half) {      // 0: Xc (also writes Xb, xn, cv), 1: S
    if (half == 1 && (!Vt0 || combine)) break;
    __syncthreads();
    // Each warp owns 8 consecutive rows.  ALL of their loads are issued before any arithmetic (one DRAM round trip per
    // warp instead of eight: the row-at-a-time form was latency-bound, ~110 us for 155 MB of traffic at n = 65536).
    const int npair = dp >> 6;                         // column pairs per lane: 1 (dp = 64) or 2 (dp = 128)
    const bool vec2 = ((ldx & 1) == 0) && ((((uintptr_t)X) & 7) == 0) && (!S || (((uintptr_t)S) & 7) == 0);
#pragma unroll 1
    for (int batch = 0; batch < 2; ++batch) {          // two batches of 4 rows: 32 loads in flight per lane, <= 128 registers
    float xv[4][2][2], sv[4][2][2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t row = r0 + w * 8 + batch * 4 + q;
      const bool rv = row < n;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c0 = 2 * lane + 64 * u;
        xv[q][u][0] = xv[q][u][1] = sv[q][u][0] = sv[q][u][1] = 0.f;
        if (u < npair && rv) {
          const float* src = (half == 0 ? X : S) + row * ldx + c0;
          if (vec2 && c0 + 1 < d) {
            const float2 t = *reinterpret_cast<const float2*>(src);
            xv[q][u][0] = t.x; xv[q][u][1] = t.y;
            if (half == 0 && S) { const float2 t2 = *reinterpret_cast<const float2*>(S + row * ldx + c0); sv[q][u][0] = t2.x; sv[q][u][1] = t2.y; }
          } else {
#pragma unroll
            for (int e = 0; e < 2; ++e)
              if (c0 + e < d) {
                xv[q][u][e] = src[e];
                if (half == 0 && S) sv[q][u][e] = S[row * ldx + c0 + e];
              }
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int rr = w * 8 + batch * 4 + q;
      const int64_t row = r0 + rr;
      const bool rv = row < n;
      float acc_n = 0.f, acc_c = 0.f, acc_r = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u >= npair) break;
        const int c0 = 2 * lane + 64 * u;
        float v2[2] = {0.f, 0.f}, t2[2] = {0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = c0 + e;
          float v = 0.f, tv = 0.f;
          if (rv && c < d) {
            if (half == 0) {
              v = xv[q][u][e] - mu[c];
              if (c == d - 1) v *= last_scale;
              acc_n = fmaf(v, v, acc_n);
              if (S) {
                const float s_ = sv[q][u][e];
                acc_c = fmaf(s_, v, acc_c);
                if (combine) tv = fmaf(-v, inv_h2, s_);
              }
            } else {
              v = xv[q][u][e];
            }
          }
          v2[e] = v; t2[e] = combine ? tv : v;
        }
        T[rr][c0] = t2[0]; T[rr][c0 + 1] = t2[1];
        if (half == 0) {
          const uint32_t p0 = pack_bf16x2(v2[0], v2[1]);
          const uint32_t p1 = pack_bf16x2(v2[0] - bf16lo_to_float(p0), v2[1] - bf16hi_to_float(p0));
          *reinterpret_cast<uint32_t*>(&Xb0[row * dp + c0]) = p0;
          *reinterpret_cast<uint32_t*>(&Xb1[row * dp + c0]) = p1;
          if (Xh) {
            __half2 hv = __floats2half2_rn(v2[0], v2[1]);
            float f0 = __low2float(hv), f1 = __high2float(hv);
            if (fabsf(f0) < 6.103515625e-05f) f0 = 0.f;     // below 2^-14: flush (never feed fp16 subnormals to the MMA)
            if (fabsf(f1) < 6.103515625e-05f) f1 = 0.f;
            hv = __floats2half2_rn(f0, f1);                   // exact: f0, f1 are fp16 values
            *reinterpret_cast<__half2*>(&Xh[row * dp + c0]) = hv;
            const float e0 = v2[0] - f0, e1 = v2[1] - f1;   // inf / NaN propagate -> bound fails -> three bf16 passes
            acc_r = fmaf(e0, e0, acc_r); acc_r = fmaf(e1, e1, acc_r);
          }
        }
      }
      if (half == 0) {
        acc_n = warp_sum(acc_n); acc_c = warp_sum(acc_c); acc_r = warp_sum(acc_r);
        if (lane == 0) {
          xn[row] = rv ? acc_n : 1e30f;   // padded row: k underflows to 0, and 0 * finite stays 0
          if (cv) cv[row] = rv ? acc_c : 0.f;
          cvs[rr] = rv ? acc_c : 0.f; ones[rr] = rv ? 1.f : 0.f;
          if (max_bits && rv) {
            atomicMax(&smax[0], __float_as_uint(acc_n));       // >= 0: uint order = float order; NaN is largest
            if (Xh) atomicMax(&smax[1], __float_as_uint(acc_r));
          }
        }
      }
    }
    }
    __syncthreads();
    if (half == 0 && max_bits && threadIdx.x < 2 && smax[threadIdx.x]) atomicMax(&max_bits[threadIdx.x], smax[threadIdx.x]);
    if (Vt0) {
      // transposed write: warp handles one coordinate c at a time, lane -> rows (2 lane, 2 lane + 1)
      const int64_t vrow0 = (half == 0 && !combine) ? dp : 0;   // Xc^T goes to rows [dp, 2dp), S^T (or W^T) to [0, dp)
      const int cmax = dp + ((half == 0 && combine) ? vt_extra : 0);
      for (int c = w; c < cmax; c += 8) {
        float a, b;
        if (c < dp) { a = T[2 * lane][c]; b = T[2 * lane + 1][c]; }
        else if (c == dp) { a = ones[2 * lane]; b = ones[2 * lane + 1]; }
        else if (c == dp + 1) { a = cvs[2 * lane]; b = cvs[2 * lane + 1]; }
        else { a = 0.f; b = 0.f; }
        const uint32_t p0 = pack_bf16x2(a, b);
        const uint32_t p1 = pack_bf16x2(a - bf16lo_to_float(p0), b - bf16hi_to_float(p0));
        const int64_t off = (vrow0 + c) * rows_pad + r0 + 2 * lane;
        *reinterpret_cast<uint32_t*>(&Vt0[off]) = p0;
        *reinterpret_cast<uint32_t*>(&Vt1[off]) = p1;
      }
    }
  }
}

int launch_tc_prep(cudaStream_t st, const float* X, const float* S, int64_t n, int64_t d, int64_t ldx, int64_t rows_pad,
                   int dp, const float* mu, float last_scale, void* Xb0, void* Xb1, void* Xh, void* Vt0, void* Vt1,
                   int vt_extra, float* xn, float* cv, const float* h_combine, uint32_t* max_bits) {
  tc_prep_kernel<<<(unsigned)(rows_pad / 64), 256, 0, st>>>(X, S, n, d, ldx, rows_pad, dp, mu, last_scale,
                                                            (__nv_bfloat16*)Xb0, (__nv_bfloat16*)Xb1, (__half*)Xh,
                                                            (__nv_bfloat16*)Vt0, (__nv_bfloat16*)Vt1, vt_extra, xn, cv,
                                                            h_combine, max_bits);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

// ---------------------------------------------------------------------------- main kernel
struct TcParams {
  const uint32_t* Qb0; const uint32_t* Qb1;   // query pieces, [m_pad x DP] bf16 viewed as packed pairs
  const float* xn; const float* cv; const float* qn;
  const float* h_dev;
  float* O; float* scal;          // [nsplit][m_pad][2*DP], [nsplit][m_pad][4]
  int64_t m_pad;
  int tiles_total, tiles_per_split;
  uint32_t* err_flag;
  const uint32_t* xn_max_bits;    // max over particles and queries of |xc|^2 (float bits), written by tc_prep_kernel
  float gram_tau;                 // ONE Gram pass when max|xc|^2 / h^2 <= gram_tau (0: always three)
  int dbg;                        // NSVGD_TC_DEBUG timing experiments: 1 = no MMAs issued, 2 = pair math skipped (results invalid)
};

constexpr int TC_BM = 128, TC_BN = 64;
constexpr int TC_DEFAULT_PAIR_WARPS = 8;
constexpr int TC_MAX_RUN_TILES = 64;   // [S | Xc] form: 12 accumulating MMAs per tile -> <= 768 sequential TMEM accumulations per run
constexpr int TC_MAX_RUN_TILES_COMBINED = 96;   // COMBINED form: 8 per tile into O -> the same 768
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_O = 0, TM_Q = 256, TM_G = 384, TM_P = 448;   // O 256 | Q 2x64 | G 64 | P 64 (b0 32 + b1 32)
// ping-pong (PP = 2, COMBINED form only: O needs 128 columns): a second G/P pair in the freed columns
//   O [0,128) | G1 [128,192) | P1 [192,256) | Q [256,384) | G0 [384,448) | P0 [448,512)
__device__ __forceinline__ constexpr uint32_t tm_g(int b) { return b == 0 ? TM_G : 128u; }
__device__ __forceinline__ constexpr uint32_t tm_p(int b) { return b == 0 ? TM_P : 192u; }

template <int DP, int VC, int PP = 1>
struct TcSmem {
  static constexpr int KH = DP / 64;                    // 64-column (128-byte) K slabs of the Gram B operand
  static constexpr int XSLAB = TC_BN * 128;             // 8 KB
  static constexpr int VSLOT = VC * 128;                // [VC rows x 64 j] bf16 = 32 KB (VC = 2 DP = 256)
  // TMA prefetch depth.  The COMBINED form spends ~1600 cycles per tile on the tensor pipe, about one L2 round trip
  // of a 16 KB TMA box: with 2 X stages / 2 tiles of V the pipe starved (70 % active, pair warps parked on g_full,
  // profiles/r01_ncu_phi_tc_combined.txt); 3 tiles of X and 3 tiles of V in flight hide it.
  static constexpr int XSTAGES = VC <= 128 ? 3 : 2;
  static constexpr int VSLOTS = VC <= 128 ? 6 : 4;
  static constexpr int NVEC = 4;                        // ring of per-tile column vectors {xn[64], cv[64]}
  static constexpr int VEC_BYTES = 2 * TC_BN * 4;
  static constexpr int X_STAGE_BYTES = 2 * KH * XSLAB;
  static constexpr int OFF_X = 0;
  static constexpr int OFF_V = OFF_X + XSTAGES * X_STAGE_BYTES;
  static constexpr int OFF_VEC = OFF_V + VSLOTS * VSLOT;
  static constexpr int OFF_BAR = OFF_VEC + NVEC * VEC_BYTES;
  static constexpr int NBARS = 1 + 2 * XSTAGES + 2 * VSLOTS + 2 * NVEC + 4 * PP + 1;
  static constexpr int TOTAL = OFF_BAR + NBARS * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;         // slack for manual 1024-B alignment
  static_assert(DYN_BYTES <= 232448, "shared memory budget exceeded");
};

// VC = columns of the second GEMM: 2*DP -> O = P [S | Xc] (drift and repulsion separately: aux output),
//                                  DP   -> O = P W, W = S - Xc/h^2 (COMBINED form: 1/3 fewer MMA flops)
// PP = 2: PING-PONG.  The pair-function warps form two groups that take alternate tiles, each with its own G and P
// buffer in TMEM, so one group's TMEM read + exp + TMEM write (~1400-2200 cycles per tile, profiles/r01_phi_pipeline.md)
// overlaps the other group's and the tensor pipe (~1570 cycles per tile) never waits for a P tile.
template <int DP, int NW, int VC, int PP>
__global__ void __launch_bounds__(64 + 32 * NW, 1)
phi_tc_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
              const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmV1,
              const TcParams p) {
  using L = TcSmem<DP, VC, PP>;
  constexpr int KH = L::KH;
  constexpr bool COMBINED = (VC == DP);
  static_assert(PP == 1 || (COMBINED && NW % 8 == 0), "ping-pong needs the COMBINED TMEM map and two equal groups");
  constexpr int GW = NW / PP;                          // pair-function warps per group
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-B aligned, stays in the shared address space
  uint8_t* sX = smem + L::OFF_X;
  uint8_t* sV = smem + L::OFF_V;
  float* sVec = (float*)(smem + L::OFF_VEC);
  uint64_t* bars = (uint64_t*)(smem + L::OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* x_full = q_full + 1;
  uint64_t* x_empty = x_full + L::XSTAGES;
  uint64_t* v_full = x_empty + L::XSTAGES;
  uint64_t* v_empty = v_full + L::VSLOTS;
  uint64_t* n_full = v_empty + L::VSLOTS;
  uint64_t* n_empty = n_full + L::NVEC;
  uint64_t* g_full = n_empty + L::NVEC;          // [PP] each
  uint64_t* g_empty = g_full + PP;
  uint64_t* p_full = g_empty + PP;
  uint64_t* p_empty = p_full + PP;
  uint64_t* o_full = p_empty + PP;
  uint32_t* tmem_slot = (uint32_t*)(o_full + 1);
  volatile uint32_t* abort_s = tmem_slot + 1;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM;
  const int split = blockIdx.y;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(t_begin + p.tiles_per_split, p.tiles_total);
  const int T = t_end - t_begin;
  Watch watch{abort_s, p.err_flag};

  if (threadIdx.x == 0) {
    *abort_s = (*(volatile uint32_t*)p.err_flag) ? 1u : 0u;
    mbar_init(q_full, NW);                       // query pieces stored to TMEM by the pair-function warps
    for (int i = 0; i < L::XSTAGES; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < L::VSLOTS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < L::NVEC; ++i) { mbar_init(&n_full[i], 1); mbar_init(&n_empty[i], GW); }
    // one arrival per pair-function warp of the group that owns the buffer (lane 0 after __syncwarp)
    for (int b = 0; b < PP; ++b) {
      mbar_init(&g_full[b], 1); mbar_init(&g_empty[b], GW);
      mbar_init(&p_full[b], GW); mbar_init(&p_empty[b], 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX0); tma_prefetch_desc(&tmX1); tma_prefetch_desc(&tmV0); tma_prefetch_desc(&tmV1);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (T > 0) {
      const bool leader = elect_one();
      int vcount = 0;
      auto load_x = [&](int t) {
        const int st = t % L::XSTAGES, u = t / L::XSTAGES;
        const int j0 = (t_begin + t) * TC_BN;
        mbar_wait(&x_empty[st], (u & 1) ^ 1, watch, 0x100 + st);
        uint8_t* dst = sX + st * L::X_STAGE_BYTES;
        if (leader) {
          if (p.dbg & 4) { if (t < L::XSTAGES) mbar_arrive_expect_tx(&x_full[st], L::X_STAGE_BYTES); else mbar_arrive(&x_full[st]); }
          else mbar_arrive_expect_tx(&x_full[st], L::X_STAGE_BYTES);
          if (!(p.dbg & 4) || t < L::XSTAGES)
          for (int kh = 0; kh < KH; ++kh) {
            tma_load_2d(dst + (0 * KH + kh) * L::XSLAB, &tmX0, &x_full[st], kh * 64, j0);
            tma_load_2d(dst + (1 * KH + kh) * L::XSLAB, &tmX1, &x_full[st], kh * 64, j0);
          }
        }
        // per-tile column vectors |xc_j|^2 and s_j.xc_j for the pair-function warps
        const int ns = t % L::NVEC, un = t / L::NVEC;
        mbar_wait(&n_empty[ns], (un & 1) ^ 1, watch, 0x180 + ns);
        if (leader) {
          mbar_arrive_expect_tx(&n_full[ns], L::VEC_BYTES);
          bulk_load_1d(sVec + ns * 2 * TC_BN, p.xn + j0, TC_BN * 4, &n_full[ns]);
          bulk_load_1d(sVec + ns * 2 * TC_BN + TC_BN, p.cv + j0, TC_BN * 4, &n_full[ns]);
        }
        __syncwarp();
      };
      auto load_v = [&](int t) {
        const int j0 = (t_begin + t) * TC_BN;
        for (int piece = 0; piece < 2; ++piece) {
          const int sl = vcount % L::VSLOTS, u = vcount / L::VSLOTS;
          ++vcount;
          mbar_wait(&v_empty[sl], (u & 1) ^ 1, watch, 0x200 + sl);
          if (leader) {
            if ((p.dbg & 4) && vcount > L::VSLOTS) mbar_arrive(&v_full[sl]);      // timing experiment: no L2 -> smem traffic
            else {
              mbar_arrive_expect_tx(&v_full[sl], L::VSLOT);
              tma_load_2d(sV + sl * L::VSLOT, piece == 0 ? &tmV0 : &tmV1, &v_full[sl], j0, 0);
            }
          }
          __syncwarp();
        }
      };
      // same order as the MMA warp consumes: X(0) .. X(PP-1), then X(t+PP), V(t)
      for (int t = 0; t < PP && t < T; ++t) load_x(t);
      for (int t = 0; t < T; ++t) {
        if (t + PP < T) load_x(t + PP);
        load_v(t);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (T > 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc_g = umma_idesc_bf16(TC_BM, TC_BN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(TC_BM, VC);
      const uint32_t sX_a = smem_u32(sX), sV_a = smem_u32(sV);
      int vcount = 0;
      // ADAPTIVE GRAM PRECISION.  delta k / k = delta g / h^2 and a single bf16 piece per operand gives
      // |delta g| <= 2^-8 |q||x|: for wide kernels (max|xc|^2 / h^2 <= gram_tau = 0.1) one pass b0 b0' keeps phi* within
      // ~2e-6 of float64 on generic clouds (measured 1.3e-6 at cfg4's ratio 0.07, 3.9e-5 at ratio 2.2:
      // profiles/r01_passcount_study.txt) and within 2^-8 * 0.1 = 4e-4 even if every coordinate's bf16 residual lined up,
      // so the two cross passes are skipped -- a third of the kernel's MMA flops.  Everything else keeps all three.
      int npass = 3;
      if (p.gram_tau > 0.f) {
        const float hh = *p.h_dev;
        const float ratio = __uint_as_float(*p.xn_max_bits) / (hh * hh);
        if (ratio <= p.gram_tau) npass = 1;          // false for NaN
      }
      mbar_wait(q_full, 0, watch, 0x300);
      tc_fence_after();
      auto issue_gram = [&](int t) {
        const int st = t % L::XSTAGES, us = t / L::XSTAGES;
        const int b = t % PP, ub = t / PP;
        mbar_wait(&g_empty[b], (ub & 1) ^ 1, watch, 0x310);
        mbar_wait(&x_full[st], us & 1, watch, 0x320 + st);
        tc_fence_after();
        const uint32_t d_tm = tmem + (b == 0 ? tm_g(0) : tm_g(1));
        const uint32_t xs = sX_a + st * L::X_STAGE_BYTES;
        if (leader) {
          uint32_t acc = 0;
          if (!(p.dbg & 1))
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            if (pass >= npass) break;
            const int pa = pass == 2 ? 1 : 0;     // Q piece: b0, b0, b1   (A operand, in TMEM)
            const int pb = pass == 1 ? 1 : 0;     // X piece: b0, b1, b0   (B operand, shared memory)
#pragma unroll
            for (int kh = 0; kh < KH; ++kh) {
              const uint64_t bd = umma_desc_sw128(xs + (pb * KH + kh) * L::XSLAB);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                umma_ts(d_tm, tmem + TM_Q + pa * 64 + (kh * 4 + ks) * 8, umma_desc_advance(bd, ks * 32), idesc_g, acc);
                acc = 1;
              }
            }
          }
          umma_commit(&x_empty[st]);
          umma_commit(&g_full[b]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int u) {
        const int b = u % PP, ub = u / PP;
        mbar_wait(&p_full[b], ub & 1, watch, 0x330);
        tc_fence_after();
        const uint32_t p0 = tmem + (b == 0 ? tm_p(0) : tm_p(1)), p1 = p0 + 32;
        {
          const int sl = vcount % L::VSLOTS, uv = vcount / L::VSLOTS; ++vcount;
          mbar_wait(&v_full[sl], uv & 1, watch, 0x340 + sl);
          tc_fence_after();
          const uint64_t bd = umma_desc_sw128(sV_a + sl * L::VSLOT);
          if (leader) {
            if (!(p.dbg & 1)) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_ts(tmem + TM_O, p0 + ks * 8, umma_desc_advance(bd, ks * 32), idesc_pv, (u > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_ts(tmem + TM_O, p1 + ks * 8, umma_desc_advance(bd, ks * 32), idesc_pv, 1u);
            }
            umma_commit(&v_empty[sl]);
          }
          __syncwarp();
        }
        {
          const int sl = vcount % L::VSLOTS, uv = vcount / L::VSLOTS; ++vcount;
          mbar_wait(&v_full[sl], uv & 1, watch, 0x350 + sl);
          tc_fence_after();
          const uint64_t bd = umma_desc_sw128(sV_a + sl * L::VSLOT);
          if (leader) {
            if (!(p.dbg & 1))
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ts(tmem + TM_O, p0 + ks * 8, umma_desc_advance(bd, ks * 32), idesc_pv, 1u);
            umma_commit(&v_empty[sl]);
            umma_commit(&p_empty[b]);
          }
          __syncwarp();
        }
      };
      // tensor-pipe order: G0 | G1 PV0 | G2 PV1 | ...  (single G and P buffers: G(t+1) is issued once the
      // pair warps have loaded G(t), and runs while they compute P(t); P.V(t-1) covers the load itself)
      // ping-pong (PP = 2): G0 G1 | G2 PV0 | G3 PV1 | ...  (G(t+2) is issued once group t%2 has loaded G(t))
      for (int t = 0; t < PP && t < T; ++t) issue_gram(t);
      for (int t = 0; t < T; ++t) {
        if (t + PP < T) issue_gram(t + PP);
        issue_pv(t);
      }
      if (leader) umma_commit(o_full);
      __syncwarp();
    }
  } else {
    // =============================== pair-function warps ==========================================
    // thread <-> (row, column part): quarter = warp % 4 selects the TMEM lanes (rows), part = which CPT
    // of the tile's 64 columns.  Several warps per SM sub-partition hide each other's latencies.
    constexpr int NPART = NW / 4;                      // warps per TMEM lane quarter (epilogue / Q-load column parts)
    constexpr int GPART = GW / 4;                      // column parts of a tile within one group
    constexpr int CPT = TC_BN / GPART;                 // columns per thread and tile (32 or 16)
    constexpr int OCOLS = VC / NPART;                  // accumulator columns per thread in the epilogue
    static_assert(OCOLS % 32 == 0, "epilogue reads 32 accumulator columns per tcgen05.ld");
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;                  // 0 .. NPART-1
    const int group = part / GPART;                    // which tiles: t % PP == group
    const int gpart = part % GPART;                    // column part inside the group's tile
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float h = *p.h_dev;
    const float cs = NSVGD_HALF_LOG2E / (h * h);          // k = exp2(-cs r2)
    const float two_cs = 2.f * cs, neg_cs = -cs;
    const float ai = neg_cs * p.qn[m0 + row_in_tile];
    const float2 two_cs2 = make_float2(two_cs, two_cs), neg_cs2 = make_float2(neg_cs, neg_cs), ai2 = make_float2(ai, ai);
    const float2 neg_one2 = make_float2(-1.f, -1.f);
    float2 rowsum2 = make_float2(0.f, 0.f), kc2 = rowsum2, ke2 = rowsum2, kg2 = rowsum2;
    if (T > 0) {
      // the CTA's 128 query rows (two bf16 pieces) -> TMEM, once: A operand of every Gram MMA
      constexpr int QW = DP / 2 / NPART;               // packed words per thread and piece
      const int64_t qoff = ((int64_t)m0 + row_in_tile) * (DP / 2) + part * QW;
#pragma unroll
      for (int piece = 0; piece < 2; ++piece) {
        const uint4* src = reinterpret_cast<const uint4*>((piece == 0 ? p.Qb0 : p.Qb1) + qoff);
        constexpr int QSTEP = QW >= 16 ? 16 : 8;
#pragma unroll
        for (int c0 = 0; c0 < QW; c0 += QSTEP) {
          uint32_t qv[QSTEP];
#pragma unroll
          for (int q = 0; q < QSTEP / 4; ++q) {
            const uint4 v = __ldg(src + c0 / 4 + q);
            qv[q * 4] = v.x; qv[q * 4 + 1] = v.y; qv[q * 4 + 2] = v.z; qv[q * 4 + 3] = v.w;
          }
          if (QSTEP == 16) tmem_st16(tmem + lane_addr + TM_Q + piece * 64 + part * QW + c0, qv);
          else tmem_st8(tmem + lane_addr + TM_Q + piece * 64 + part * QW + c0, qv);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full);
    }
    const uint32_t tm_g_mine = tmem + lane_addr + (group == 0 ? tm_g(0) : tm_g(1)) + gpart * CPT;
    const uint32_t tm_p_mine = tmem + lane_addr + (group == 0 ? tm_p(0) : tm_p(1));
    for (int t = group; t < T; t += PP) {
      const int ns = t % L::NVEC, un = t / L::NVEC;
      const int ub = t / PP;
      mbar_wait(&g_full[group], ub & 1, watch, 0x400);
      __syncwarp();
      tc_fence_after();
      uint32_t g[CPT];
      if (CPT == 32) tmem_ld32(tm_g_mine, g);
      else tmem_ld16(tm_g_mine, g);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&g_empty[group]);
      mbar_wait(&n_full[ns], un & 1, watch, 0x420 + ns);
      const float4* xn4 = reinterpret_cast<const float4*>(sVec + ns * 2 * TC_BN + gpart * CPT);
      const float4* cv4 = reinterpret_cast<const float4*>(sVec + ns * 2 * TC_BN + TC_BN + gpart * CPT);
      uint32_t pk0[CPT / 2], pk1[CPT / 2];
      if (p.dbg & 2) {
#pragma unroll
        for (int q = 0; q < CPT / 2; ++q) { pk0[q] = g[2 * q]; pk1[q] = g[2 * q + 1]; }
      } else
      // packed fp32x2 math (FFMA2/FADD2): two columns per instruction
#pragma unroll
      for (int q4 = 0; q4 < CPT / 4; ++q4) {
        const float4 xv = xn4[q4];
        const float4 cvv = cv4[q4];
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
          const float2 gg = make_float2(__uint_as_float(g[q4 * 4 + e]), __uint_as_float(g[q4 * 4 + e + 1]));
          const float2 xs = e == 0 ? make_float2(xv.x, xv.y) : make_float2(xv.z, xv.w);
          const float2 cc = e == 0 ? make_float2(cvv.x, cvv.y) : make_float2(cvv.z, cvv.w);
          float2 ex = __ffma2_rn(gg, two_cs2, ai2);
          ex = __ffma2_rn(xs, neg_cs2, ex);
          const float2 k = make_float2(ex2_approx(ex.x), ex2_approx(ex.y));
          // k = p0 + p1 (+ <= 2^-16 k): the MMA consumes the pieces, so the row statistics use
          // kq = p0 + p1 as well -> rowsum*q_i - K X cancels the self term exactly
          const uint32_t a0 = pack_bf16x2(k.x, k.y);
          const float2 hh = make_float2(bf16lo_to_float(a0), bf16hi_to_float(a0));
          const float2 r = __ffma2_rn(hh, neg_one2, k);
          const uint32_t a1 = pack_bf16x2(r.x, r.y);
          pk0[q4 * 2 + (e >> 1)] = a0;
          pk1[q4 * 2 + (e >> 1)] = a1;
          const float2 kq = __fadd2_rn(hh, make_float2(bf16lo_to_float(a1), bf16hi_to_float(a1)));
          rowsum2 = __fadd2_rn(rowsum2, kq);
          kc2 = __ffma2_rn(kq, cc, kc2);
          ke2 = __ffma2_rn(kq, ex, ke2);
          if (COMBINED) kg2 = __ffma2_rn(kq, gg, kg2);      // sum_j k g_ij = qc_i . (K Xc)_i, recovers qc_i . (K S)_i
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&n_empty[ns]);
      // single P buffer: P(t) is stored as soon as P.V(t-1) has consumed P(t-1); G(t+1) keeps the tensor pipe busy meanwhile
      mbar_wait(&p_empty[group], (ub & 1) ^ 1, watch, 0x410);
      __syncwarp();
      tc_fence_after();
      if (CPT == 32) {
        tmem_st16(tm_p_mine + gpart * 16, pk0);
        tmem_st16(tm_p_mine + 32 + gpart * 16, pk1);
      } else {
        tmem_st8(tm_p_mine + gpart * 8, pk0);
        tmem_st8(tm_p_mine + 32 + gpart * 8, pk1);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[group]);
    }
    // ---- epilogue: O (TMEM) -> workspace; the NPART warps of a quarter take OCOLS columns each
    const int64_t row = (int64_t)m0 + row_in_tile;
    float* orow = p.O + ((int64_t)split * p.m_pad + row) * VC + part * OCOLS;
    float* srow = p.scal + (((int64_t)split * NPART + part) * p.m_pad + row) * 4;
    if (T > 0) {
      mbar_wait(o_full, 0, watch, 0x500);
      __syncwarp();
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < OCOLS; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + TM_O + part * OCOLS + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(orow + c0 + q * 4) =
              make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]), __uint_as_float(v[q * 4 + 2]),
                          __uint_as_float(v[q * 4 + 3]));
      }
      // b = sum k r2 = -sum k e / cs  (this thread's part of the columns)
      *reinterpret_cast<float4*>(srow) = make_float4(rowsum2.x + rowsum2.y, kc2.x + kc2.y, -(ke2.x + ke2.y) / cs, kg2.x + kg2.y);
      tc_fence_before();
    } else {
      // empty j-range (tail split): contribute zeros
      for (int c = 0; c < OCOLS; c += 4) *reinterpret_cast<float4*>(orow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(srow) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------- COMBINED-form kernel (phi_tcw)
// The kernel the SVGD step runs (no aux split requested): O = P W' with W' = [S - Xc/h^2 | 1 | c | 0 ...] (N = DP + 16).
// Versus phi_tc_kernel above:
//  (1) rowsum_i = (K 1)_i and (K c)_i are two extra COLUMNS OF THE SECOND GEMM instead of per-pair register work: the pair
//      function no longer rebuilds the quantised k (4 unpacks + 2 FADD2 per column pair) nor accumulates k c_j -- 11 instead of
//      23 issue slots per column pair; the ncu capture of the previous kernel (profiles/r02_ncu_phi_tc_before.txt) showed the
//      tensor pipe 57 % active, waiting for the pair warps (~800 issue slots per tile and scheduler vs 1024 MMA cycles).
//      rowsum comes from the same P pieces as K W, so rowsum_i qc_i - (K Xc)_i still cancels the self term exactly.
//  (2) P is double-buffered in TMEM: P(t+1) is stored while P(t) V(t) runs; G(t+2) fills the tensor pipe behind it.
//  (3) ONE-PASS GRAM IN FP16, decided on the device from a RIGOROUS bound: with x^ = fp16(xc) (11-bit significands, exact
//      products, fp32 accumulation), |g~ - g| <= |q| |x - x^| + |q - q^| |x^| <= 2 N R + R^2, N = max |xc|, R = max |xc - x^|
//      (both reduced by tc_prep_kernel over particles AND queries), so |delta ln k| = |delta g| / h^2 <= (2 N R + R^2) / h^2.
//      One pass is used iff that bound <= gram_eps (default 5e-5, NSVGD_TC_GRAM_EPS; 0 = always three bf16-split passes);
//      the guaranteed relative kernel error of the one-pass branch is therefore <= 5e-5 < the 1e-4 contract.
// TMEM map (DP = 128): O [0,144) | Q [160,288) | G [288,352) | P0 [352,416) | P1 [416,480)
struct TcwParams {
  const uint32_t* Qb0; const uint32_t* Qb1; const uint32_t* Qh;   // query pieces [m_pad x DP]: bf16 b0, b1 and fp16 single piece
  const float* xn; const float* qn;
  const float* h_dev;
  float* O; float* scal;          // O [slab][m_pad][DP + 16], scal [slab * NPART + part][m_pad][2] = {sum k r2, sum k g}
  int64_t m_pad;
  int tiles_total, tiles_per_split;
  int run_tiles, runs_per_split;  // a CTA flushes its TMEM accumulator every run_tiles tiles into slab split * runs_per_split + run
  uint32_t* err_flag;
  const uint32_t* max_bits;       // [0] max |xc|^2, [1] max |xc - fp16(xc)|^2 (float bits), written by tc_prep_kernel
  float gram_eps;
  int dbg;
};

__device__ __forceinline__ bool tcw_onepass(const TcwParams& p) {
  if (!(p.gram_eps > 0.f)) return false;
  const float h = *p.h_dev;
  const float nmax = sqrtf(__uint_as_float(p.max_bits[0])), r2 = __uint_as_float(p.max_bits[1]);
  const float bound = (2.f * nmax * sqrtf(r2) + r2) / (h * h);
  return bound <= p.gram_eps;          // false for NaN / inf (overflowed fp16, NaN particles, h = 0)
}

template <int DP>
struct TcwSmem {
  static constexpr int KH = DP / 64;
  static constexpr int VC = DP + 16;
  static constexpr int XSLAB = TC_BN * 128;             // 8 KB: 64 particles x 64 columns x 2 bytes
  static constexpr int VSLOT = VC * 128;                // [VC rows x 64 j] bf16
  static constexpr int XSTAGES = 3, VSLOTS = 6, NVEC = 4;
  static constexpr int X_STAGE_BYTES = 2 * KH * XSLAB;  // both bf16 pieces (the fp16 single piece uses the first half)
  static constexpr int OFF_X = 0;
  static constexpr int OFF_V = OFF_X + XSTAGES * X_STAGE_BYTES;
  static constexpr int OFF_VEC = OFF_V + VSLOTS * VSLOT;
  static constexpr int OFF_BAR = OFF_VEC + NVEC * TC_BN * 4;
  static constexpr int NBARS = 1 + 2 * XSTAGES + 2 * VSLOTS + 2 * NVEC + 3 + 4 + 2;
  static constexpr int TOTAL = OFF_BAR + NBARS * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static_assert(DYN_BYTES <= 232448, "shared memory budget exceeded");
  // TMEM columns
  static constexpr uint32_t TM_O = 0, TM_Q = VC + 16, TM_G = TM_Q + DP, TM_P = TM_G + 64;   // P0 = TM_P, P1 = TM_P + 64
  static_assert(TM_P + 128 <= 512, "TMEM budget exceeded");
};

// NG = 2: two groups of 8 pair-function warps take alternate tiles (group g owns P buffer g; the G buffer is shared and
// re-issued as soon as the consuming group has loaded it), and the Gram runs TWO tiles ahead: a tile's TMEM read + exp +
// split + TMEM write (~1400 cycles per warp, serial) then has two tensor-pipe periods instead of one.
template <int DP, int NG>
__global__ void __launch_bounds__(96 + 32 * 8 * NG, 1)
phi_tcw_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
               const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmV0,
               const __grid_constant__ CUtensorMap tmV1, const TcwParams p) {
  using L = TcwSmem<DP>;
  constexpr int KH = L::KH, VC = L::VC, NW = 8 * NG, GW = 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sX = smem + L::OFF_X;
  uint8_t* sV = smem + L::OFF_V;
  float* sVec = (float*)(smem + L::OFF_VEC);
  uint64_t* bars = (uint64_t*)(smem + L::OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* x_full = q_full + 1;
  uint64_t* x_empty = x_full + L::XSTAGES;
  uint64_t* v_full = x_empty + L::XSTAGES;
  uint64_t* v_empty = v_full + L::VSLOTS;
  uint64_t* n_full = v_empty + L::VSLOTS;
  uint64_t* n_empty = n_full + L::NVEC;
  uint64_t* g_full = n_empty + L::NVEC;    // [2]: one per consuming group (a group only ever waits for its own tiles,
  uint64_t* g_empty = g_full + 2;          //      so a parity wait can never alias a completion meant for the other group)
  uint64_t* p_full = g_empty + 1;          // [2]
  uint64_t* p_empty = p_full + 2;          // [2]
  uint64_t* o_full = p_empty + 2;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_slot = (uint32_t*)(o_empty + 1);
  volatile uint32_t* abort_s = tmem_slot + 1;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM;
  const int split = blockIdx.y;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(t_begin + p.tiles_per_split, p.tiles_total);
  const int T = t_end - t_begin;
  Watch watch{abort_s, p.err_flag};

  if (threadIdx.x == 0) {
    *abort_s = (*(volatile uint32_t*)p.err_flag) ? 1u : 0u;
    mbar_init(q_full, NW);
    for (int i = 0; i < L::XSTAGES; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < L::VSLOTS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < L::NVEC; ++i) { mbar_init(&n_full[i], 1); mbar_init(&n_empty[i], GW); }
    mbar_init(&g_full[0], 1); mbar_init(&g_full[1], 1); mbar_init(g_empty, GW);       // one arrival per warp of the group that consumes the tile
    for (int b = 0; b < 2; ++b) { mbar_init(&p_full[b], GW); mbar_init(&p_empty[b], 1); }
    mbar_init(o_full, 1); mbar_init(o_empty, NW);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX0); tma_prefetch_desc(&tmX1); tma_prefetch_desc(&tmXh); tma_prefetch_desc(&tmV0); tma_prefetch_desc(&tmV1);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool onepass = tcw_onepass(p);      // the same three device words in every role -> the same decision

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // ring positions are running counters (index + phase bit): no divisions on the single-warp critical paths
    if (T > 0) {
      const bool leader = elect_one();
      int xi = 0, ni = 0, vi = 0;
      uint32_t xph = 0, nph = 0, vph = 0;
      auto load_x = [&](int t) {
        const int j0 = (t_begin + t) * TC_BN;
        mbar_wait(&x_empty[xi], xph ^ 1, watch, 0x100 + xi);
        uint8_t* dst = sX + xi * L::X_STAGE_BYTES;
        if (leader) {
          if (onepass) {
            mbar_arrive_expect_tx(&x_full[xi], KH * L::XSLAB);
            for (int kh = 0; kh < KH; ++kh) tma_load_2d(dst + kh * L::XSLAB, &tmXh, &x_full[xi], kh * 64, j0);
          } else {
            mbar_arrive_expect_tx(&x_full[xi], L::X_STAGE_BYTES);
            for (int kh = 0; kh < KH; ++kh) {
              tma_load_2d(dst + (0 * KH + kh) * L::XSLAB, &tmX0, &x_full[xi], kh * 64, j0);
              tma_load_2d(dst + (1 * KH + kh) * L::XSLAB, &tmX1, &x_full[xi], kh * 64, j0);
            }
          }
        }
        if (++xi == L::XSTAGES) { xi = 0; xph ^= 1; }
        mbar_wait(&n_empty[ni], nph ^ 1, watch, 0x180 + ni);
        if (leader) {
          mbar_arrive_expect_tx(&n_full[ni], TC_BN * 4);
          bulk_load_1d(sVec + ni * TC_BN, p.xn + j0, TC_BN * 4, &n_full[ni]);
        }
        if (++ni == L::NVEC) { ni = 0; nph ^= 1; }
        __syncwarp();
      };
      auto load_v = [&](int t) {
        const int j0 = (t_begin + t) * TC_BN;
#pragma unroll
        for (int piece = 0; piece < 2; ++piece) {
          mbar_wait(&v_empty[vi], vph ^ 1, watch, 0x200 + vi);
          if (leader) {
            if ((p.dbg & 4) && (piece == 1 || (p.dbg & 8)) && t >= L::VSLOTS) mbar_arrive(&v_full[vi]);   // timing experiment: less L2 -> smem traffic
            else {
              mbar_arrive_expect_tx(&v_full[vi], L::VSLOT);
              tma_load_2d(sV + vi * L::VSLOT, piece == 0 ? &tmV0 : &tmV1, &v_full[vi], j0, 0);
            }
          }
          if (++vi == L::VSLOTS) { vi = 0; vph ^= 1; }
          __syncwarp();
        }
      };
      for (int t = 0; t < NG && t < T; ++t) load_x(t);
      for (int t = 0; t < T; ++t) {
        if (t + NG < T) load_x(t + NG);
        load_v(t);
      }
    }
  } else if (warp == 1) {
    // =============================== Gram MMA issuer ===============================
    // The MMAs are issued by TWO warps (this one: G = Qc Xc^T; warp 2 + NW: O += P V).  With one issuing warp the ncu
    // source page showed that warp busy 87 % of the time with its own ~175 instructions per tile (uniform-datapath
    // descriptor arithmetic, mbarrier polls, UTCHMMA / UTCBAR through the MIO queue: ~9.6 cycles each) while the tensor
    // pipe sat at 67 %: the issue thread, not the pipe or the pair warps, paced the kernel.  Every dependency between the
    // two streams already goes through an mbarrier, so they need no ordering between each other.
    if (T > 0) {
      const bool leader = elect_one();
      const uint32_t idesc_g = onepass ? umma_idesc_f16(TC_BM, TC_BN) : umma_idesc_bf16(TC_BM, TC_BN);
      const uint32_t sX_a = smem_u32(sX);
      const int npass = onepass ? 1 : 3;
      int xi = 0;
      uint32_t xph = 0;
      mbar_wait(q_full, 0, watch, 0x300);
      tc_fence_after();
      for (int t = 0; t < T; ++t) {
        mbar_wait(g_empty, (t & 1) ^ 1, watch, 0x310);       // the consuming group has loaded G(t-1)
        mbar_wait(&x_full[xi], xph, watch, 0x320 + xi);
        tc_fence_after();
        const uint32_t xs = sX_a + xi * L::X_STAGE_BYTES;
        if (leader) {
          uint32_t acc = 0;
          if (!(p.dbg & 1))
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            if (pass >= npass) break;
            const int pa = pass == 2 ? 1 : 0;     // Q piece: b0, b0, b1   (A operand, in TMEM)
            const int pb = pass == 1 ? 1 : 0;     // X piece: b0, b1, b0   (B operand, shared memory)
#pragma unroll
            for (int kh = 0; kh < KH; ++kh) {
              const uint64_t bd = umma_desc_sw128(xs + (pb * KH + kh) * L::XSLAB);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                umma_ts(tmem + L::TM_G, tmem + L::TM_Q + pa * (DP / 2) + (kh * 4 + ks) * 8, umma_desc_advance(bd, ks * 32), idesc_g, acc);
                acc = 1;
              }
            }
          }
          umma_commit(&x_empty[xi]);
          umma_commit(&g_full[NG == 2 ? (t & 1) : 0]);
        }
        __syncwarp();
        if (++xi == L::XSTAGES) { xi = 0; xph ^= 1; }
      }
    }
  } else if (warp == 2 + NW) {
    // =============================== P V MMA issuer ===============================
    if (T > 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc_pv = umma_idesc_bf16(TC_BM, VC);
      const uint32_t sV_a = smem_u32(sV);
      int vi = 0;
      uint32_t vph = 0;
      // ACCUMULATION RUNS.  The TMEM accumulator truncates (bias ~2^-25 per accumulating MMA, profiles/r01_accuracy_vs_nsplit.txt),
      // so O is flushed to a partial slab every run_tiles <= 96 tiles and restarted; the slabs are added in fp32 RN by the
      // epilogue kernel.  The flush happens INSIDE a long-lived CTA (a few j-splits, chosen for whole waves) instead of ending
      // the CTA: the fixed cost of a CTA (launch, TMEM alloc, query block -> TMEM, pipeline fill and drain: ~6 us, measured by
      // varying the split count) is paid 7 instead of 45 times per SM at n = 65536.
      int run = 0, run_start = 0, run_end = min(p.run_tiles, T);
      for (int u = 0; u < T; ++u) {
        const int b = u & 1;
        const bool first = u == run_start;
        if (first && run > 0) mbar_wait(o_empty, (run - 1) & 1, watch, 0x360);     // the pair warps have drained the previous run
        mbar_wait(&p_full[b], (u >> 1) & 1, watch, 0x330 + b);
        const uint32_t p0 = tmem + L::TM_P + b * 64, p1 = p0 + 32;
        {
          mbar_wait(&v_full[vi], vph, watch, 0x340 + vi);
          tc_fence_after();
          const uint64_t bd = umma_desc_sw128(sV_a + vi * L::VSLOT);
          if (leader) {
            if (!(p.dbg & 1)) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_ts(tmem + L::TM_O, p0 + ks * 8, umma_desc_advance(bd, ks * 32), idesc_pv, (!first || ks > 0) ? 1u : 0u);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_ts(tmem + L::TM_O, p1 + ks * 8, umma_desc_advance(bd, ks * 32), idesc_pv, 1u);
            }
            umma_commit(&v_empty[vi]);
          }
          __syncwarp();
          if (++vi == L::VSLOTS) { vi = 0; vph ^= 1; }
        }
        {
          mbar_wait(&v_full[vi], vph, watch, 0x350 + vi);
          tc_fence_after();
          const uint64_t bd = umma_desc_sw128(sV_a + vi * L::VSLOT);
          if (leader) {
            if (!(p.dbg & 1))
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ts(tmem + L::TM_O, p0 + ks * 8, umma_desc_advance(bd, ks * 32), idesc_pv, 1u);
            umma_commit(&v_empty[vi]);
            umma_commit(&p_empty[b]);
            if (u + 1 == run_end) umma_commit(o_full);
          }
          __syncwarp();
          if (++vi == L::VSLOTS) { vi = 0; vph ^= 1; }
        }
        if (u + 1 == run_end) { run_start = run_end; run_end = min(run_end + p.run_tiles, T); ++run; }
      }
    }
  } else {
    // =============================== pair-function warps ==========================================
    constexpr int CPT = TC_BN / 2;                     // 32 columns per thread and tile
    constexpr int NPART = 2 * NG;                      // warps per TMEM lane quarter
    constexpr int OCOLS = VC / NPART;                  // accumulator columns per thread in the epilogue (72 / 40 / 36 / 20)
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;                  // 0 .. NPART-1
    const int group = part >> 1;                       // which tiles: t % NG == group
    const int gpart = part & 1;                        // column half inside the tile
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float h = *p.h_dev;
    const float cs = NSVGD_HALF_LOG2E / (h * h);
    const float two_cs = 2.f * cs, neg_cs = -cs;
    const float ai = neg_cs * p.qn[m0 + row_in_tile];
    const float2 two_cs2 = make_float2(two_cs, two_cs), neg_cs2 = make_float2(neg_cs, neg_cs), ai2 = make_float2(ai, ai);
    const float2 neg_one2 = make_float2(-1.f, -1.f);
    float2 ke2 = make_float2(0.f, 0.f), kg2 = ke2;
    if (T > 0) {
      // the CTA's 128 query rows -> TMEM, once: A operand of every Gram MMA (fp16 single piece, or the two bf16 pieces)
      constexpr int QW = DP / 2 / NPART;               // packed words per thread and piece (32 / 16 / 8)
      constexpr int QSTEP = QW >= 16 ? 16 : 8;
      const int64_t qoff = ((int64_t)m0 + row_in_tile) * (DP / 2) + part * QW;
      const int npieces = onepass ? 1 : 2;
      for (int piece = 0; piece < npieces; ++piece) {
        const uint4* src = reinterpret_cast<const uint4*>((onepass ? p.Qh : (piece == 0 ? p.Qb0 : p.Qb1)) + qoff);
#pragma unroll
        for (int c0 = 0; c0 < QW; c0 += QSTEP) {
          uint32_t qv[QSTEP];
#pragma unroll
          for (int q = 0; q < QSTEP / 4; ++q) {
            const uint4 v = __ldg(src + c0 / 4 + q);
            qv[q * 4] = v.x; qv[q * 4 + 1] = v.y; qv[q * 4 + 2] = v.z; qv[q * 4 + 3] = v.w;
          }
          if (QSTEP == 16) tmem_st16(tmem + lane_addr + L::TM_Q + piece * (DP / 2) + part * QW + c0, qv);
          else tmem_st8(tmem + lane_addr + L::TM_Q + piece * (DP / 2) + part * QW + c0, qv);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full);
    }
    const uint32_t tm_g_mine = tmem + lane_addr + L::TM_G + gpart * CPT;
    const int64_t row = (int64_t)m0 + row_in_tile;
    // ---- flush of one accumulation run: O (TMEM) -> partial slab; the NPART warps of a quarter take OCOLS columns each
    auto store4 = [&](float* dst, const uint32_t* v) {
      *reinterpret_cast<float4*>(dst) = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
    };
    auto flush = [&](int run) {
      const int64_t slab = (int64_t)split * p.runs_per_split + run;
      float* orow = p.O + (slab * p.m_pad + row) * VC + part * OCOLS;
      float* srow = p.scal + ((slab * NPART + part) * p.m_pad + row) * 2;
      mbar_wait(o_full, run & 1, watch, 0x500);
      __syncwarp();
      tc_fence_after();
      // OCOLS = 72 / 40 / 36 / 20 columns of this thread's row: chunks of 32, then 16, 8, 4
      constexpr int N32 = OCOLS / 32, R32 = OCOLS % 32;
      static_assert(OCOLS % 4 == 0, "epilogue chunks");
      const uint32_t tm_o = tmem + lane_addr + L::TM_O + part * OCOLS;
#pragma unroll 1
      for (int c0 = 0; c0 < N32 * 32; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tm_o + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 8; ++q) store4(orow + c0 + q * 4, v + q * 4);
      }
      int c0 = N32 * 32;
      if constexpr ((R32 & 16) != 0) {
        uint32_t v[16];
        tmem_ld16(tm_o + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) store4(orow + c0 + q * 4, v + q * 4);
        c0 += 16;
      }
      if constexpr ((R32 & 8) != 0) {
        uint32_t v[8];
        tmem_ld8(tm_o + c0, v);
        tmem_wait_ld();
        store4(orow + c0, v); store4(orow + c0 + 4, v + 4);
        c0 += 8;
      }
      if constexpr ((R32 & 4) != 0) {
        uint32_t v[4];
        tmem_ld4(tm_o + c0, v);
        tmem_wait_ld();
        store4(orow + c0, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      // b = sum k r2 = -sum k log2(k) / cs  (this thread's columns of this run)
      *reinterpret_cast<float2*>(srow) = make_float2(-(ke2.x + ke2.y) / cs, kg2.x + kg2.y);
      ke2 = make_float2(0.f, 0.f); kg2 = ke2;
    };
    int run = 0, run_end = min(p.run_tiles, T);
    for (int t = group; t < T; t += NG) {
      while (t >= run_end) { flush(run); ++run; run_end = min(run_end + p.run_tiles, T); }   // this group is done with the run
      const int ns = t % L::NVEC, un = t / L::NVEC;
      const int b = t & 1, ub = t >> 1;
      mbar_wait(&g_full[group], (t / NG) & 1, watch, 0x400);
      __syncwarp();
      tc_fence_after();
      uint32_t g[CPT];
      tmem_ld32(tm_g_mine, g);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_empty);
      mbar_wait(&n_full[ns], un & 1, watch, 0x420 + ns);
      const float4* xn4 = reinterpret_cast<const float4*>(sVec + ns * TC_BN + gpart * CPT);
      uint32_t pk0[CPT / 2], pk1[CPT / 2];
#pragma unroll
      for (int q4 = 0; q4 < CPT / 4; ++q4) {
        const float4 xv = xn4[q4];
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
          const float2 gg = make_float2(__uint_as_float(g[q4 * 4 + e]), __uint_as_float(g[q4 * 4 + e + 1]));
          const float2 xs = e == 0 ? make_float2(xv.x, xv.y) : make_float2(xv.z, xv.w);
          float2 ex = __ffma2_rn(gg, two_cs2, ai2);
          ex = __ffma2_rn(xs, neg_cs2, ex);
          const float2 k = make_float2(ex2_approx(ex.x), ex2_approx(ex.y));
          const uint32_t a0 = pack_bf16x2(k.x, k.y);
          const float2 hh = make_float2(bf16lo_to_float(a0), bf16hi_to_float(a0));
          const float2 r = __ffma2_rn(hh, neg_one2, k);
          pk0[q4 * 2 + (e >> 1)] = a0;
          pk1[q4 * 2 + (e >> 1)] = pack_bf16x2(r.x, r.y);
          ke2 = __ffma2_rn(k, ex, ke2);                // sum k log2(k): b = sum k r2 = -(.) / cs
          kg2 = __ffma2_rn(k, gg, kg2);                // sum_j k g_ij = qc_i . (K Xc)_i, recovers qc_i . (K S)_i
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&n_empty[ns]);
      mbar_wait(&p_empty[b], (ub & 1) ^ 1, watch, 0x410 + b);
      __syncwarp();
      tc_fence_after();
      const uint32_t tm_p_mine = tmem + lane_addr + L::TM_P + b * 64;
      tmem_st16(tm_p_mine + gpart * 16, pk0);
      tmem_st16(tm_p_mine + 32 + gpart * 16, pk1);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[b]);
    }
    if (T > 0) {
      const int nruns = (T + p.run_tiles - 1) / p.run_tiles;
      while (run < nruns) { flush(run); ++run; }
    }
    // runs this CTA does not have (tail split): contribute zeros
    for (; run < p.runs_per_split; ++run) {
      const int64_t slab = (int64_t)split * p.runs_per_split + run;
      float* orow = p.O + (slab * p.m_pad + row) * VC + part * OCOLS;
      for (int c = 0; c < OCOLS; c += 4) *reinterpret_cast<float4*>(orow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float2*>(p.scal + ((slab * NPART + part) * p.m_pad + row) * 2) = make_float2(0.f, 0.f);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------- host driver
struct TcPlan {
  int dp; int64_t n_pad, m_pad; int nsplit, tiles_total, tiles_per_split; bool q_is_x;
  int run_tiles, runs_per_split;      // phi_tcw: in-kernel accumulation runs (slabs = nsplit * runs_per_split)
};

// j-splits of the COMBINED-form kernel.  A CTA costs ~6 tile-times of fixed overhead (measured: ~6 us of launch, TMEM
// alloc, query block -> TMEM, pipeline fill / drain) and one CTA fits per SM, so the step time is
// waves x (tiles per CTA + 6): pick the split count that minimises it (whole waves, long CTAs).
static void plan_tcw(TcPlan& pl) {
  const int sms = device_sm_count();
  const int64_t row_blocks = pl.m_pad / TC_BM;
  int best = 1; double best_cost = 1e300;
  for (int s = 1; s <= 32 && s <= pl.tiles_total; ++s) {
    const int64_t tps = ceil_div(pl.tiles_total, s);
    if (tps < 16 && s > 1) break;
    const int64_t waves = ceil_div(row_blocks * s, sms);
    const double cost = (double)waves * ((double)tps + 6.0);
    if (cost < best_cost * 0.995) { best_cost = cost; best = s; }
  }
  if (const char* env = getenv("NSVGD_TC_NSPLIT")) {   // tuning experiments
    const int v = atoi(env);
    if (v >= 1 && v <= 64) best = std::min(v, pl.tiles_total);
  }
  pl.nsplit = best;
  pl.tiles_per_split = (int)ceil_div(pl.tiles_total, pl.nsplit);
  pl.runs_per_split = (int)ceil_div(pl.tiles_per_split, TC_MAX_RUN_TILES_COMBINED);
  pl.run_tiles = (int)ceil_div(pl.tiles_per_split, pl.runs_per_split);
}

static TcPlan make_plan(const PhiArgs& a, bool combined) {
  TcPlan pl{};
  pl.dp = a.d <= 64 ? 64 : 128;
  pl.n_pad = round_up(a.n, 128);
  pl.m_pad = round_up(a.m, 128);
  pl.q_is_x = (a.Q == a.X && a.m == a.n && a.ldq == a.ldx);
  pl.tiles_total = (int)(pl.n_pad / TC_BN);
  // Split the j range.  (1) Accuracy: the TMEM accumulator truncates (measured bias ~2^-25 per
  // accumulating MMA, profiles/r01_accuracy_vs_nsplit.txt), so one CTA accumulates at most
  // TC_MAX_RUN_TILES tiles (12 MMAs each) and the partial sums are added in fp32 RN by the epilogue.
  // (2) Occupancy: row blocks x splits CTAs should cover the SMs in whole waves (1 CTA/SM).
  const int sms = device_sm_count();
  const int64_t row_blocks = pl.m_pad / TC_BM;
  const int min_split = (int)ceil_div(pl.tiles_total, combined ? TC_MAX_RUN_TILES_COMBINED : TC_MAX_RUN_TILES);
  int best = min_split; double best_eff = 0.0;
  for (int s = min_split; s <= min_split + 8; ++s) {
    if (s > pl.tiles_total) break;
    if (pl.tiles_total / s < 8 && s > min_split) break;
    const int64_t ctas = row_blocks * s;
    const double eff = (double)ctas / (double)(ceil_div(ctas, sms) * sms);
    if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
  }
  pl.nsplit = best;
  if (const char* env = getenv("NSVGD_TC_NSPLIT")) {   // accuracy / tuning experiments
    const int v = atoi(env);
    if (v >= 1 && v <= 64) pl.nsplit = std::min(v, pl.tiles_total);
  }
  pl.tiles_per_split = (int)ceil_div(pl.tiles_total, pl.nsplit);
  return pl;
}

template <int DP, int NW, int VC, int PP>
static int launch_tc_variant(cudaStream_t st, const TcPlan& pl, const CUtensorMap* maps, const TcParams& prm) {
  using L = TcSmem<DP, VC, PP>;
  NSVGD_PROPAGATE(ensure_dyn_smem((const void*)phi_tc_kernel<DP, NW, VC, PP>, L::DYN_BYTES));
  dim3 grid((unsigned)(pl.m_pad / TC_BM), (unsigned)pl.nsplit);
  prof_begin(0, st);
  phi_tc_kernel<DP, NW, VC, PP><<<grid, 64 + 32 * NW, L::DYN_BYTES, st>>>(maps[2], maps[3], maps[4], maps[5], prm);
  prof_end(0, st);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

// kernel variant: NSVGD_TC_PAIR_WARPS in {8,16} (tuning knob; 8 measured best, profiles/r01_phi_variants.txt)
static int tc_pair_warps() {
  const char* e = getenv("NSVGD_TC_PAIR_WARPS");
  return (e && atoi(e) == 16) ? 16 : ((e && atoi(e) == 8) ? 8 : TC_DEFAULT_PAIR_WARPS);
}

// NSVGD_TC_PINGPONG=1: ping-pong, 2 x 8 pair warps.  Default off: measured 5.04 vs 4.99 ms at n=65536, d=128
// (profiles/r01_phi_pipeline.md) -- the kernel is not latency-bound, so the extra warps buy nothing.
static bool tc_pingpong() {
  const char* e = getenv("NSVGD_TC_PINGPONG");
  return e && atoi(e) == 1;
}

template <int DP>
static int launch_tc(cudaStream_t st, const TcPlan& pl, const CUtensorMap* maps, const TcParams& prm, bool combined) {
  if (combined) {
    if constexpr (DP == 128) {      // ping-pong needs 128 accumulator columns per lane quarter split over 4 warps
      if (tc_pingpong()) return launch_tc_variant<DP, 16, DP, 2>(st, pl, maps, prm);
    }
    return launch_tc_variant<DP, 8, DP, 1>(st, pl, maps, prm);
  }
  if (tc_pair_warps() == 16) return launch_tc_variant<DP, 16, 2 * DP, 1>(st, pl, maps, prm);
  return launch_tc_variant<DP, 8, 2 * DP, 1>(st, pl, maps, prm);
}

// COMBINED form (second GEMM against W = S - Xc/h^2, N = dp): whenever the caller does not ask for the
// separate [drift, repulsion] aux output.  NSVGD_TC_COMBINE=0 forces the [S | Xc] form (A/B measurements).
static bool tc_combined(const PhiArgs& a) {
  const char* e = getenv("NSVGD_TC_COMBINE");
  if (e && atoi(e) == 0) return false;
  return a.aux == nullptr;
}

template <int DP, int NG>
static int launch_tcw(cudaStream_t st, const TcPlan& pl, const CUtensorMap* maps, const TcwParams& prm) {
  using L = TcwSmem<DP>;
  NSVGD_PROPAGATE(ensure_dyn_smem((const void*)phi_tcw_kernel<DP, NG>, L::DYN_BYTES));
  dim3 grid((unsigned)(pl.m_pad / TC_BM), (unsigned)pl.nsplit);
  prof_begin(0, st);
  phi_tcw_kernel<DP, NG><<<grid, 96 + 32 * 8 * NG, L::DYN_BYTES, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], prm);
  prof_end(0, st);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

// NSVGD_TC_GROUPS=1: a single group of 8 pair-function warps (A/B measurements); default 2
static int tcw_groups() {
  const char* e = getenv("NSVGD_TC_GROUPS");
  return (e && atoi(e) == 1) ? 1 : 2;
}

// NSVGD_TC_KERNEL=v1: the round-1 kernel also for the COMBINED form (A/B measurements)
static bool tc_v1() {
  const char* e = getenv("NSVGD_TC_KERNEL");
  return e && !strcmp(e, "v1");
}

int phi_tc(cudaStream_t st, const PhiArgs& a, void* ws, size_t ws_bytes, bool dry, size_t* need) {
  if (!dry) NSVGD_REQUIRE(a.d <= 128, NSVGD_ERR_UNSUPPORTED, "phi_tc: tensor-core path supports d <= 128 (d=%lld)", (long long)a.d);
  const bool combined = !dry && tc_combined(a);      // dry run sizes for the larger of the two forms
  const bool use_w = combined && !tc_v1();
  TcPlan pl = make_plan(a, combined);
  TcPlan plw = pl;
  plan_tcw(plw);
  const size_t o_floats = std::max((size_t)pl.nsplit * 2 * pl.dp, (size_t)plw.nsplit * plw.runs_per_split * (pl.dp + 16)) * (size_t)pl.m_pad;
  const size_t scal_floats = std::max((size_t)pl.nsplit * 16, (size_t)plw.nsplit * plw.runs_per_split * 8) * (size_t)pl.m_pad;
  if (use_w) pl = plw;
  const int dp = pl.dp;
  const int vc = use_w ? dp + 16 : (combined ? dp : 2 * dp);
  WsCarver w(ws, ws_bytes, dry);
  float* mu = w.take<float>((size_t)a.d);
  double* mpart = w.take<double>((size_t)colmean_partials(a.n, a.d));
  __nv_bfloat16* Xb0 = w.take<__nv_bfloat16>((size_t)(pl.n_pad * dp), 1024);
  __nv_bfloat16* Xb1 = w.take<__nv_bfloat16>((size_t)(pl.n_pad * dp), 1024);
  __half* Xh = w.take<__half>((size_t)(pl.n_pad * dp), 1024);
  __nv_bfloat16* Vt0 = w.take<__nv_bfloat16>((size_t)(2 * dp * pl.n_pad), 1024);
  __nv_bfloat16* Vt1 = w.take<__nv_bfloat16>((size_t)(2 * dp * pl.n_pad), 1024);
  float* xn = w.take<float>((size_t)pl.n_pad);
  float* cv = w.take<float>((size_t)pl.n_pad);
  __nv_bfloat16 *Qb0 = Xb0, *Qb1 = Xb1;
  __half* Qh = Xh;
  float* qn = xn;
  if (!pl.q_is_x || dry) {   // dry run sizes for the rectangular worst case
    Qb0 = w.take<__nv_bfloat16>((size_t)(pl.m_pad * dp), 1024);
    Qb1 = w.take<__nv_bfloat16>((size_t)(pl.m_pad * dp), 1024);
    Qh = w.take<__half>((size_t)(pl.m_pad * dp), 1024);
    qn = w.take<float>((size_t)pl.m_pad);
  }
  float* O = w.take<float>(o_floats, 1024);
  float* scal = w.take<float>(scal_floats);
  uint32_t* err_flag = w.take<uint32_t>(4);
  double* partials = w.take<double>((size_t)(2 * epilogue_blocks(a.m)));
  if (need) *need = w.off;
  if (dry) return NSVGD_OK;
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "phi_tc: workspace too small (%zu < %zu)", ws_bytes, w.off);

  prof_begin(13, st);
  NSVGD_PROPAGATE(launch_colmean(st, a.X, a.n, a.d, a.ldx, mpart, mu));
  NSVGD_CUDA(cudaMemsetAsync(err_flag, 0, 16, st));
  uint32_t* max_bits = err_flag + 1;      // [1] max |xc|^2, [2] max fp16 residual^2: zeroed with the error flag just above
  NSVGD_PROPAGATE(launch_tc_prep(st, a.X, a.S, a.n, a.d, a.ldx, pl.n_pad, dp, mu, 1.f, Xb0, Xb1, use_w ? Xh : nullptr, Vt0, Vt1,
                                 use_w ? 16 : 0, xn, cv, combined ? a.h_dev : nullptr, max_bits));
  if (!pl.q_is_x)
    NSVGD_PROPAGATE(launch_tc_prep(st, a.Q, nullptr, a.m, a.d, a.ldq, pl.m_pad, dp, mu, 1.f, Qb0, Qb1, use_w ? Qh : nullptr,
                                   nullptr, nullptr, 0, qn, nullptr, nullptr, max_bits));
  prof_end(13, st);
  EpilogueArgs e{};
  if (use_w) {
    CUtensorMap maps[5];
    NSVGD_PROPAGATE(make_tmap_bf16_2d(&maps[0], Xb0, dp, pl.n_pad, (uint64_t)dp * 2, 64, TC_BN));
    NSVGD_PROPAGATE(make_tmap_bf16_2d(&maps[1], Xb1, dp, pl.n_pad, (uint64_t)dp * 2, 64, TC_BN));
    NSVGD_PROPAGATE(make_tmap_bf16_2d(&maps[2], Xh, dp, pl.n_pad, (uint64_t)dp * 2, 64, TC_BN));
    NSVGD_PROPAGATE(make_tmap_bf16_2d(&maps[3], Vt0, pl.n_pad, vc, (uint64_t)pl.n_pad * 2, TC_BN, vc));
    NSVGD_PROPAGATE(make_tmap_bf16_2d(&maps[4], Vt1, pl.n_pad, vc, (uint64_t)pl.n_pad * 2, TC_BN, vc));
    TcwParams prm{};
    prm.Qb0 = (const uint32_t*)Qb0; prm.Qb1 = (const uint32_t*)Qb1; prm.Qh = (const uint32_t*)Qh; prm.xn = xn; prm.qn = qn;
    prm.h_dev = a.h_dev; prm.O = O; prm.scal = scal; prm.m_pad = pl.m_pad;
    prm.tiles_total = pl.tiles_total; prm.tiles_per_split = pl.tiles_per_split; prm.err_flag = err_flag;
    prm.run_tiles = pl.run_tiles; prm.runs_per_split = pl.runs_per_split;
    prm.max_bits = max_bits;
    prm.gram_eps = 5e-5f;           // guaranteed bound on |delta k / k| of the one-pass branch (see phi_tcw_kernel)
    if (const char* env = getenv("NSVGD_TC_GRAM_EPS")) prm.gram_eps = (float)atof(env);   // 0 -> always three Gram passes
    if (const char* env = getenv("NSVGD_TC_DEBUG")) prm.dbg = atoi(env);
    const int ng = tcw_groups();
    if (dp == 64) NSVGD_PROPAGATE(ng == 2 ? (launch_tcw<64, 2>(st, pl, maps, prm)) : (launch_tcw<64, 1>(st, pl, maps, prm)));
    else NSVGD_PROPAGATE(ng == 2 ? (launch_tcw<128, 2>(st, pl, maps, prm)) : (launch_tcw<128, 1>(st, pl, maps, prm)));
    e.form = 3; e.nsplit = pl.nsplit * pl.runs_per_split; e.scal_parts = 2 * ng * e.nsplit;
  } else {
    CUtensorMap maps[6];
    NSVGD_PROPAGATE(make_tmap_bf16_2d(&maps[2], Xb0, dp, pl.n_pad, (uint64_t)dp * 2, 64, TC_BN));
    NSVGD_PROPAGATE(make_tmap_bf16_2d(&maps[3], Xb1, dp, pl.n_pad, (uint64_t)dp * 2, 64, TC_BN));
    NSVGD_PROPAGATE(make_tmap_bf16_2d(&maps[4], Vt0, pl.n_pad, vc, (uint64_t)pl.n_pad * 2, TC_BN, vc));
    NSVGD_PROPAGATE(make_tmap_bf16_2d(&maps[5], Vt1, pl.n_pad, vc, (uint64_t)pl.n_pad * 2, TC_BN, vc));
    TcParams prm{};
    prm.Qb0 = (const uint32_t*)Qb0; prm.Qb1 = (const uint32_t*)Qb1; prm.xn = xn; prm.cv = cv; prm.qn = qn; prm.h_dev = a.h_dev; prm.O = O; prm.scal = scal; prm.m_pad = pl.m_pad;
    prm.tiles_total = pl.tiles_total; prm.tiles_per_split = pl.tiles_per_split; prm.err_flag = err_flag;
    if (const char* env = getenv("NSVGD_TC_DEBUG")) prm.dbg = atoi(env);
    prm.xn_max_bits = max_bits;
    prm.gram_tau = 0.f;             // the bf16 single-pass Gram of round 1 has no bound inside the 1e-4 contract: opt-in only
    if (const char* env = getenv("NSVGD_TC_GRAM_TAU")) prm.gram_tau = (float)atof(env);
    if (dp == 64) NSVGD_PROPAGATE(launch_tc<64>(st, pl, maps, prm, combined));
    else NSVGD_PROPAGATE(launch_tc<128>(st, pl, maps, prm, combined));
    e.form = combined ? 2 : 1;
    e.scal_parts = ((combined ? ((tc_pingpong() && dp == 128) ? 16 : 8) : tc_pair_warps()) / 4) * pl.nsplit;
    e.nsplit = pl.nsplit;
  }
  e.O = O; e.scal = scal; e.err_flag = err_flag;
  e.m_pad = pl.m_pad; e.ldo = vc; e.dpad = dp; e.mu = mu;
  e.Q = a.Q; e.ldq = a.ldq; e.Sq = a.Sq; e.m = a.m; e.n = a.n; e.d = a.d; e.h_dev = a.h_dev;
  e.phi = a.phi; e.ldphi = a.ldphi; e.aux = a.aux; e.stats_dev = a.stats_dev; e.partials = partials;
  prof_begin(14, st);
  const int rc = launch_epilogue(st, e);
  prof_end(14, st);
  return rc;
}

}  // namespace nsvgd
