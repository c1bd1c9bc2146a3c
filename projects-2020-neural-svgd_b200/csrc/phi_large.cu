// phi_large.cu — phi* for d > 128 (BNN weight space: d = 4096, n = 16384) on hand-written tcgen05 GEMM kernels.
// [stein.py:180-222 + kernels.py:52-65; closed forms in SURVEY.md 8a; SURVEY 7.2-3: "d > 128 cannot keep the output
//  resident in TMEM ... materialise K and run a plain GEMM"]
//
// With d in the thousands the 128 x d output block of the fused kernel (phi_tc.cu) no longer fits TMEM, so phi* becomes two
// dense contractions with the pair function fused into the epilogue of the first.  Both run on ONE persistent, warp-
// specialised kernel (`lg_gemm`): TMA (SWIZZLE_128B) -> shared-memory ring -> tcgen05.mma (bf16 pieces, fp32 accumulate in
// TMEM, two accumulator buffers so the epilogue of tile t overlaps the MMAs of tile t + 1) -> tcgen05.ld epilogue:
//   prep   : centre, split every operand into two bf16 pieces (x = b0 + b1, 2^-18 relative), |xc|^2, s.xc,
//            W^T = (S - Xc/h^2)^T (COMBINED form, phi_tc.cu; transposed so that the particle index is the K dimension)
//   launch 1 (EPI_PAIR)  : G tile = Qc Xc^T (K = d; b0 b0' + b1 b0' + b0 b1' accumulated in one run) and, straight from
//            TMEM, k = exp2(2c g - c|q|^2 - c|x|^2) -> the two bf16 pieces of k (P, row-major: the A operand of launch 2)
//            + the four row statistics rowsum / sum k c_j / sum k r2 / sum k g per 256-column tile.  G never reaches HBM.
//   launch 2 (EPI_STORE) : O_s[m x d] = P W over the particles of chunk s (K = 4096: 768 accumulating MMAs, the TMEM
//            truncation budget of DESIGN.md 4.1), one fp32 slab per chunk
//   epilogue (update.cu, GRAMW form) adds the slabs and the per-tile statistics in fp32 RN.
// Piece reuse: one pipeline stage holds A0, A1, B0, B1 of a 64-wide k-block and feeds all three passes (4 tile loads per
// 3 MMA blocks instead of 6).  CTA pairs (cta_group::2, M = 256 across the two SMs of a TPC): each CTA stages only half
// of the B tile, which takes the L2 -> shared-memory traffic from 96 to 64 KB per k-block and CTA -- below the L2
// throughput cap at full tensor rate.  NSVGD_LARGE_CTAS=1 selects the single-CTA (M = 128) variant.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "select.cuh"
#include "tc_common.cuh"

namespace nsvgd {

using namespace tc;

constexpr int LG_BM = 128, LG_BN = 256, LG_BK = 64;
constexpr int64_t LG_NK = 4096;          // particles per accumulation run of the second GEMM
constexpr int LG_THREADS = 192;          // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr uint32_t LG_TMEM_COLS = 512;   // two 128 x 256 fp32 accumulators
enum { LG_EPI_PAIR = 0, LG_EPI_STORE = 1, LG_EPI_CLASSIFY = 2 };

template <int CTAS>
struct LgSmem {
  static constexpr int A_TILE = LG_BM * LG_BK * 2;                 // 16 KB: 128 rows x 64 bf16
  static constexpr int B_TILE = (LG_BN / CTAS) * LG_BK * 2;        // 32 KB, or 16 KB per CTA of a pair
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;            // A0 | A1 | B0 | B1
  static constexpr int STAGES = CTAS == 1 ? 2 : 3;
  static constexpr int OFF_BAR = STAGES * STAGE;
  static constexpr int NBAR = 2 * STAGES + 4;
  static constexpr int OFF_COL = OFF_BAR + 128;                    // float4 [4 warps][256 columns]: column statistics of mirrored tiles
  static constexpr int BYTES = OFF_COL + 4 * LG_BN * 16 + 1024;    // + slack for the 1024-byte alignment of the ring
  static_assert(NBAR * 8 + 16 <= 128, "barrier block");
};

struct LgParams {
  int epi;
  int tiles_m, tiles_n, nsplit, group_m, group_n;   // tiles_m in units of 128 * CTAS rows
  int total;                               // work units (tiles x splits) of the launch
  int kb_total, kb_per_split;              // 64-wide k-blocks: all of them / per accumulation run
  int m_rows;                              // valid rows of this launch
  int64_t row_off;                         // first row of this launch within the whole problem
  // EPI_PAIR
  const float* qn; const float* xn; const float* cv; const float* h_dev;
  __nv_bfloat16* P0; __nv_bfloat16* P1; int64_t ldp;
  float* scal; int64_t scal_stride;        // scal[(slot * scal_stride + row_off + r) * 4]
  int sym;                                 // Q == X: only tiles on or above the diagonal are computed; an off-diagonal tile also
                                           // writes its transpose (P^T and the column statistics).  slot = 2 * tile_n for the row
                                           // statistics, the 128-row block index of the tile's rows for the mirrored ones; else tile_n
  int ctas;
  int64_t diag_off;                        // >= 0: the query rows ARE the particles diag_off, diag_off + 1, ... (Q == X, or the row block of a
                                           // row-sharded rank): entry (i, diag_off + i) is a particle against itself -> g = |xc_i|^2 exactly
  // EPI_STORE
  float* O; int64_t ldo; int64_t o_split_stride; int n_cols;
  // EPI_CLASSIFY (median filter for d > 128, median_large.cu): symmetric tiles, pairs i < j < n_valid; xn = |z|^2
  FastMedState* fs;                        // bracket [L, U] (read), below (atomic), overflow -> err_flag
  unsigned short* tlist; uint32_t* tcount; uint32_t tcap;   // per 128 x 256 CTA tile: the uncertain pairs as (row << 8 | column)
  float kappa_g, kappa_d; int n_valid;
  uint32_t* err_flag;
};

__device__ __forceinline__ void lg_decode_sym(const LgParams& p, int u, int& mt, int& nt);
__device__ __forceinline__ void lg_decode(const LgParams& p, int u, int& mt, int& nt, int& s) {
  if (p.sym) { s = 0; lg_decode_sym(p, u, mt, nt); return; }
  const int per_split = p.tiles_m * p.tiles_n;
  s = u / per_split;
  int r = u - s * per_split;
  // super-tiles of group_m row tiles x group_n column tiles (~ one wave of CTAs), column-major inside; the super-tiles walk down
  // the rows of one column group before moving to the next: consecutive waves share the column group's B operand in L2 and the
  // tiles in flight share both operands
  const int per_ngroup = p.tiles_m * p.group_n;
  const int gn = min(r / per_ngroup, (p.tiles_n - 1) / p.group_n);
  r -= gn * per_ngroup;
  const int wn = min(p.group_n, p.tiles_n - gn * p.group_n);
  const int band = p.group_m * wn;
  const int b = r / band;
  r -= b * band;
  const int gm = min(p.group_m, p.tiles_m - b * p.group_m);
  nt = gn * p.group_n + r / gm;
  mt = b * p.group_m + (r - (r / gm) * gm);
}

// symmetric launch: row tile mt (128 * ctas rows) takes the column tiles nt >= (mt * ctas) / 2; bands of group_m row tiles,
// column-major inside a band (same enumeration on the host: lg_sym_tiles)
__host__ __device__ __forceinline__ int lg_sym_nmin(int mt, int ctas) { return (mt * ctas) >> 1; }
__device__ __forceinline__ void lg_decode_sym(const LgParams& p, int u, int& mt, int& nt) {
  int r0 = 0, gm = 0;
  for (int b = 0;; ++b) {
    r0 = b * p.group_m;
    gm = min(p.group_m, p.tiles_m - r0);
    int cnt = 0;
    for (int i = 0; i < gm; ++i) cnt += p.tiles_n - lg_sym_nmin(r0 + i, p.ctas);
    if (u < cnt || r0 + gm >= p.tiles_m) break;
    u -= cnt;
  }
  nt = lg_sym_nmin(r0, p.ctas);
  for (;;) {
    const int rmax = p.ctas == 2 ? nt : 2 * nt + 1;          // last row tile that takes column tile nt
    const int rows = min(gm, rmax - r0 + 1);
    if (rows >= gm) break;
    if (u < rows) { mt = r0 + u; return; }
    u -= rows;
    ++nt;
  }
  nt += u / gm;
  mt = r0 + u % gm;
}
static int64_t lg_sym_tiles(int tiles_m, int tiles_n, int ctas) {
  int64_t t = 0;
  for (int mt = 0; mt < tiles_m; ++mt) t += tiles_n - lg_sym_nmin(mt, ctas);
  return t;
}

// t[e] of lane l -> returns sum over the 32 lanes of t[l'] ... for column l: 31 shuffles instead of 32 x 5
__device__ __forceinline__ float warp_transpose_sum32(float (&t)[32], int lane) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int half = 16 >> s;
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = up ? t[i + half] : t[i];
      const float send = up ? t[i] : t[i + half];
      t[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return t[0];
}

template <int CTAS>
__device__ __forceinline__ void lg_body(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmB0,
                                        const CUtensorMap& tmB1, const LgParams& p) {
  using L = LgSmem<CTAS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + L::OFF_BAR);
  uint64_t* full = bars;                   // [STAGES] TMA -> MMA        (pair: the leader's are used by both CTAs)
  uint64_t* empty = full + L::STAGES;      // [STAGES] MMA -> TMA        (pair: multicast commit)
  uint64_t* tfull = empty + L::STAGES;     // [2] accumulator ready      (pair: multicast commit)
  uint64_t* tempty = tfull + 2;            // [2] accumulator drained    (pair: both CTAs' epilogue warps arrive on the leader's)
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  volatile uint32_t* abort_s = tmem_slot + 1;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const int unit = blockIdx.x / CTAS, nunits = gridDim.x / CTAS;
  const int total = p.total;
  Watch watch{abort_s, p.err_flag};

  if (threadIdx.x == 0) {
    *abort_s = (*(volatile uint32_t*)p.err_flag) ? 1u : 0u;
    for (int i = 0; i < L::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4 * CTAS); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB0); tma_prefetch_desc(&tmB1);
  }
  if (warp == 1) {
    if (CTAS == 2) { tmem_alloc_pair(tmem_slot, LG_TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, LG_TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer (both CTAs of a pair) ===============================
    const bool leader = elect_one();
    int si = 0;
    uint32_t ph = 0;
    for (int u = unit; u < total; u += nunits) {
      if (*abort_s) break;
      int mt, nt, s;
      lg_decode(p, u, mt, nt, s);
      const int row0 = (mt * CTAS + (int)rank) * LG_BM;
      const int col0 = nt * LG_BN + (int)rank * (LG_BN / CTAS);
      const int kb0 = s * p.kb_per_split;
      const int nkb = min(p.kb_per_split, p.kb_total - kb0);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[si], ph ^ 1, watch, 0x500 + si);
        if (leader) {
          uint8_t* dst = smem + si * L::STAGE;
          const int kc = (kb0 + kb) * LG_BK;
          if (CTAS == 2) {
            const uint32_t fb = mapa_u32(smem_u32(&full[si]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full[si], 2 * L::STAGE);
            tma_load_2d_pair(dst, &tmA0, fb, kc, row0);
            tma_load_2d_pair(dst + 2 * L::A_TILE, &tmB0, fb, kc, col0);
            tma_load_2d_pair(dst + L::A_TILE, &tmA1, fb, kc, row0);
            tma_load_2d_pair(dst + 2 * L::A_TILE + L::B_TILE, &tmB1, fb, kc, col0);
          } else {
            mbar_arrive_expect_tx(&full[si], L::STAGE);
            tma_load_2d(dst, &tmA0, &full[si], kc, row0);
            tma_load_2d(dst + 2 * L::A_TILE, &tmB0, &full[si], kc, col0);
            tma_load_2d(dst + L::A_TILE, &tmA1, &full[si], kc, row0);
            tma_load_2d(dst + 2 * L::A_TILE + L::B_TILE, &tmB1, &full[si], kc, col0);
          }
        }
        __syncwarp();
        if (++si == L::STAGES) { si = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (the leader CTA only) ===============================
    if (rank == 0) {
      const bool leader = elect_one();
      const uint32_t idesc = umma_idesc_bf16(LG_BM * CTAS, LG_BN);
      const uint32_t sbase = smem_u32(smem);
      int si = 0, it = 0;
      uint32_t ph = 0;
      for (int u = unit; u < total; u += nunits, ++it) {
        if (*abort_s) break;
        int mt, nt, s;
        lg_decode(p, u, mt, nt, s);
        const int kb0 = s * p.kb_per_split;
        const int nkb = min(p.kb_per_split, p.kb_total - kb0);
        const int as = it & 1;
        mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1, watch, 0x520 + as);     // both CTAs have drained this accumulator
        tc_fence_after();
        const uint32_t dcol = tmem + (uint32_t)as * LG_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[si], ph, watch, 0x540 + si);
          tc_fence_after();
          if (leader) {
            const uint32_t st = sbase + si * L::STAGE;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              const uint64_t ad = umma_desc_sw128(st + (pass == 1 ? L::A_TILE : 0));                       // A piece: b0, b1, b0
              const uint64_t bd = umma_desc_sw128(st + 2 * L::A_TILE + (pass == 2 ? L::B_TILE : 0));       // B piece: b0, b0, b1
#pragma unroll
              for (int ks = 0; ks < LG_BK / 16; ++ks) {
                const uint32_t acc = (kb | pass | ks) ? 1u : 0u;
                if (CTAS == 2) umma_ss_pair(dcol, umma_desc_advance(ad, ks * 32), umma_desc_advance(bd, ks * 32), idesc, acc);
                else umma_ss(dcol, umma_desc_advance(ad, ks * 32), umma_desc_advance(bd, ks * 32), idesc, acc);
              }
            }
            if (CTAS == 2) umma_commit_pair(&empty[si]); else umma_commit(&empty[si]);
          }
          __syncwarp();
          if (++si == L::STAGES) { si = 0; ph ^= 1; }
        }
        if (leader) { if (CTAS == 2) umma_commit_pair(&tfull[as]); else umma_commit(&tfull[as]); }
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue: one thread per accumulator row ===============================
    const int q = warp & 3;                                    // the TMEM lane quarter this warp may read
    const int r = q * 32 + lane;
    const uint32_t tempty_remote = CTAS == 2 ? mapa_u32(smem_u32(&tempty[0]), 0) : 0u;
    float4* colbuf = reinterpret_cast<float4*>(smem + L::OFF_COL);
    float cs = 0.f;
    if (p.epi == LG_EPI_PAIR) { const float h = *p.h_dev; cs = NSVGD_HALF_LOG2E / (h * h); }
    int it = 0;
    for (int u = unit; u < total; u += nunits, ++it) {
      if (*abort_s) break;
      int mt, nt, s;
      lg_decode(p, u, mt, nt, s);
      const int as = it & 1;
      const int grow = (mt * CTAS + (int)rank) * LG_BM + r;    // row of this launch
      const bool valid = grow < p.m_rows;
      const int col0 = nt * LG_BN;
      mbar_wait(&tfull[as], (it >> 1) & 1, watch, 0x560 + as);
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)as * LG_BN;
      if (p.epi == LG_EPI_PAIR) {
        // sym: every row of the launch is a (possibly padded) particle, so qn is defined (1e30 for the padding: k = 0)
        const float ai = (valid || p.sym) ? -cs * p.qn[p.row_off + grow] : 0.f;
        const float cs2 = 2.f * cs;
        const int rb128 = mt * CTAS + (int)rank;                 // 128-row block of this CTA's rows
        const bool mirror = p.sym && nt > (rb128 >> 1);          // off-diagonal tile: also emit its transpose
        const float ci = mirror ? p.cv[grow] : 0.f;
        float rowsum = 0.f, kc = 0.f, ke = 0.f, kg = 0.f;
        __nv_bfloat16* p0 = p.P0 + (int64_t)grow * p.ldp + col0;
        __nv_bfloat16* p1 = p.P1 + (int64_t)grow * p.ldp + col0;
#pragma unroll 1
        for (int c = 0; c < LG_BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_wait_ld();
          if (c == LG_BN / 32 - 1) {           // the accumulator is in registers: hand the buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CTAS == 2) mbar_arrive_cluster(tempty_remote + as * 8); else mbar_arrive(&tempty[as]); }
          }
          // The accumulator of a diagonal entry grows monotonically to |xc_i|^2 over 3 d / 16 MMAs and picks up the TMEM
          // truncation bias (2^-25 per MMA): at d = 4096 that is 2e-3 in g_ii, i.e. k_ii = 1 - 2e-4 where the
          // reference has exp(0) = 1.  The entry is known in closed form, so it is set, not computed.
          if (p.diag_off >= 0) {
            const int64_t jd = p.diag_off + p.row_off + grow - col0 - c * 32;      // column of this row's own particle in the chunk
            if (__any_sync(0xffffffffu, jd >= 0 && jd < 32)) {
              const uint32_t gii = __float_as_uint((valid || p.sym) ? p.qn[p.row_off + grow] : 0.f);
              const int jdi = (int)jd;
#pragma unroll
              for (int e = 0; e < 32; ++e)      // predicated selects on named registers (a plain `if (e == jdi) v[e] = ..` becomes a
                                                // dynamically indexed local-memory array)
                asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\tselp.b32 %0, %3, %0, p;\n\t}" : "+r"(v[e]) : "r"(jdi), "r"(e), "r"(gii));
            }
          }
          uint32_t w0[16], w1[16];
          float qv[32], qe[32];
          const float4* xn4 = reinterpret_cast<const float4*>(p.xn + col0 + c * 32);
          const float4* cv4 = reinterpret_cast<const float4*>(p.cv + col0 + c * 32);
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4) {
            const float4 xs = __ldg(xn4 + e4), cc = __ldg(cv4 + e4);
            const float xa[4] = {xs.x, xs.y, xs.z, xs.w}, ca[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const int e = e4 * 4 + e2 * 2;
              const float g0 = __uint_as_float(v[e]), g1 = __uint_as_float(v[e + 1]);
              const float ex0 = fmaf(xa[e2 * 2], -cs, fmaf(g0, cs2, ai));
              const float ex1 = fmaf(xa[e2 * 2 + 1], -cs, fmaf(g1, cs2, ai));
              const float k0 = ex2_approx(ex0), k1 = ex2_approx(ex1);
              const uint32_t a0 = pack_bf16x2(k0, k1);
              const float h0 = bf16lo_to_float(a0), h1 = bf16hi_to_float(a0);
              const uint32_t a1 = pack_bf16x2(k0 - h0, k1 - h1);
              w0[e >> 1] = a0; w1[e >> 1] = a1;
              const float q0 = h0 + bf16lo_to_float(a1), q1 = h1 + bf16hi_to_float(a1);   // the quantised k launch 2 consumes
              rowsum += q0 + q1;
              kc = fmaf(q0, ca[e2 * 2], fmaf(q1, ca[e2 * 2 + 1], kc));
              ke = fmaf(q0, ex0, fmaf(q1, ex1, ke));
              kg = fmaf(q0, g0, fmaf(q1, g1, kg));
              qv[e] = q0; qv[e + 1] = q1;
              qe[e] = q0 * ex0; qe[e + 1] = q1 * ex1;
            }
          }
          if (valid) {
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              *reinterpret_cast<uint4*>(p0 + c * 32 + x * 8) = make_uint4(w0[4 * x], w0[4 * x + 1], w0[4 * x + 2], w0[4 * x + 3]);
              *reinterpret_cast<uint4*>(p1 + c * 32 + x * 8) = make_uint4(w1[4 * x], w1[4 * x + 1], w1[4 * x + 2], w1[4 * x + 3]);
            }
          }
          if (mirror) {
            // k_ji = k_ij: the transposed tile of P (a warp writes 32 consecutive rows' entries of one column: 64 bytes)
            // and the statistics of the rows j of the column block, summed over this CTA's 128 rows i
            unsigned short* t0 = reinterpret_cast<unsigned short*>(p.P0) + (int64_t)(col0 + c * 32) * p.ldp + grow;
            unsigned short* t1 = reinterpret_cast<unsigned short*>(p.P1) + (int64_t)(col0 + c * 32) * p.ldp + grow;
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              t0[(int64_t)e * p.ldp] = (unsigned short)(w0[e >> 1] & 0xFFFFu);
              t0[(int64_t)(e + 1) * p.ldp] = (unsigned short)(w0[e >> 1] >> 16);
              t1[(int64_t)e * p.ldp] = (unsigned short)(w1[e >> 1] & 0xFFFFu);
              t1[(int64_t)(e + 1) * p.ldp] = (unsigned short)(w1[e >> 1] >> 16);
            }
            float t[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) t[e] = qv[e] * __uint_as_float(v[e]);
            const float s_kg = warp_transpose_sum32(t, lane);
            const float s_ke = warp_transpose_sum32(qe, lane);
#pragma unroll
            for (int e = 0; e < 32; ++e) t[e] = qv[e] * ci;
            const float s_kc = warp_transpose_sum32(t, lane);
            const float s_rs = warp_transpose_sum32(qv, lane);
            colbuf[q * LG_BN + c * 32 + lane] = make_float4(s_rs, s_kc, s_ke, s_kg);
          }
        }
        if (valid)      // b = sum k r2 = -sum k e / cs
          *reinterpret_cast<float4*>(p.scal + ((int64_t)(p.sym ? 2 * nt : nt) * p.scal_stride + p.row_off + grow) * 4) =
              make_float4(rowsum, kc, -ke / cs, kg);
        if (mirror) {
          asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
          for (int hcol = 0; hcol < 2; ++hcol) {
            const int col = hcol * 128 + r;
            float4 acc = colbuf[col];
#pragma unroll
            for (int wq = 1; wq < 4; ++wq) {
              const float4 o = colbuf[wq * LG_BN + col];
              acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            }
            if (col0 + col < p.m_rows)
              *reinterpret_cast<float4*>(p.scal + ((int64_t)rb128 * p.scal_stride + col0 + col) * 4) =
                  make_float4(acc.x, acc.y, -acc.z / cs, acc.w);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      } else if (p.epi == LG_EPI_CLASSIFY) {
        // d~ = |z_i|^2 + |z_j|^2 - 2 g~ against the bracket with the rigorous radius eps = kg N2 + kd (d~ + kg N2), N2 = |z_i|^2 + |z_j|^2:
        // certainly below L -> counted, certainly above U -> dropped, else listed for the defined-order evaluation
        uint32_t* ucount = reinterpret_cast<uint32_t*>(smem + L::OFF_COL);
        const int rb128 = mt * CTAS + (int)rank;
        const size_t tile_id = (size_t)rb128 * p.tiles_n + nt;
        unsigned short* tl = p.tlist + tile_id * p.tcap;
        if (r == 0) *ucount = 0u;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const float Lv = p.fs->L, Uv = p.fs->U;
        const float zi = p.xn[grow];
        const float kt = p.kappa_g * (1.f + p.kappa_d);
        uint32_t below = 0;
#pragma unroll 1
        for (int c = 0; c < LG_BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_wait_ld();
          if (c == LG_BN / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CTAS == 2) mbar_arrive_cluster(tempty_remote + as * 8); else mbar_arrive(&tempty[as]); }
          }
          const float4* xn4 = reinterpret_cast<const float4*>(p.xn + col0 + c * 32);
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4) {
            const float4 xs = __ldg(xn4 + e4);
            const float xa[4] = {xs.x, xs.y, xs.z, xs.w};
#pragma unroll
            for (int e1 = 0; e1 < 4; ++e1) {
              const int e = e4 * 4 + e1;
              const int gj = col0 + c * 32 + e;
              const float n2 = zi + xa[e1];
              const float dt = fmaf(-2.f, __uint_as_float(v[e]), n2);
              const float eps = fmaf(p.kappa_d, dt, kt * n2);
              const bool ok = grow < gj && gj < p.n_valid;
              const bool is_below = dt + eps < Lv;
              const bool is_above = dt - eps > Uv;
              if (ok && is_below) ++below;
              if (ok && !is_below && !is_above) {
                const uint32_t pos = atomicAdd(ucount, 1u);
                if (pos < p.tcap) tl[pos] = (unsigned short)((r << 8) | (c * 32 + e));
              }
            }
          }
        }
        below = __reduce_add_sync(0xffffffffu, below);
        if (lane == 0 && below) atomicAdd(&p.fs->below, (unsigned long long)below);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (r == 0) {
          const uint32_t cnt = *ucount;
          p.tcount[tile_id] = min(cnt, p.tcap);
          if (cnt > p.tcap) atomicExch(p.err_flag, 0x40000000u);      // list overflow (massive ties): the exact SIMT select takes over
        }
      } else {
        float* o = p.O + (int64_t)s * p.o_split_stride + (int64_t)grow * p.ldo + col0;
#pragma unroll 1
        for (int c = 0; c < LG_BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_wait_ld();
          if (c == LG_BN / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CTAS == 2) mbar_arrive_cluster(tempty_remote + as * 8); else mbar_arrive(&tempty[as]); }
          }
          if (valid) {
#pragma unroll
            for (int x = 0; x < 8; ++x)
              if (col0 + c * 32 + x * 4 < p.n_cols)
                *reinterpret_cast<uint4*>(o + c * 32 + x * 4) = make_uint4(v[4 * x], v[4 * x + 1], v[4 * x + 2], v[4 * x + 3]);
          }
        }
      }
    }
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    if (CTAS == 2) tmem_dealloc_pair(tmem, LG_TMEM_COLS); else tmem_dealloc(tmem, LG_TMEM_COLS);
  }
}

__global__ void __launch_bounds__(LG_THREADS, 1)
lg_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ LgParams p) {
  lg_body<1>(tmA0, tmA1, tmB0, tmB1, p);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LG_THREADS, 1)
lg_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ LgParams p) {
  lg_body<2>(tmA0, tmA1, tmB0, tmB1, p);
}

static bool lg_sym_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NSVGD_LARGE_SYM"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v != 0;
}
static int lg_ctas() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NSVGD_LARGE_CTAS"); v = (e && atoi(e) == 1) ? 1 : 2; }
  return v;
}

// A pieces: rows x K (row stride lda elements); B pieces: cols x K
static int lg_launch(cudaStream_t st, int ctas, const __nv_bfloat16* A0, const __nv_bfloat16* A1, int64_t a_rows, int64_t lda,
                     const __nv_bfloat16* B0, const __nv_bfloat16* B1, int64_t b_rows, int64_t ldb, int64_t kdim, LgParams prm) {
  CUtensorMap mA0, mA1, mB0, mB1;
  NSVGD_PROPAGATE(make_tmap_bf16_2d(&mA0, A0, (uint64_t)kdim, (uint64_t)a_rows, (uint64_t)lda * 2, LG_BK, LG_BM));
  NSVGD_PROPAGATE(make_tmap_bf16_2d(&mA1, A1, (uint64_t)kdim, (uint64_t)a_rows, (uint64_t)lda * 2, LG_BK, LG_BM));
  NSVGD_PROPAGATE(make_tmap_bf16_2d(&mB0, B0, (uint64_t)kdim, (uint64_t)b_rows, (uint64_t)ldb * 2, LG_BK, (uint32_t)(LG_BN / ctas)));
  NSVGD_PROPAGATE(make_tmap_bf16_2d(&mB1, B1, (uint64_t)kdim, (uint64_t)b_rows, (uint64_t)ldb * 2, LG_BK, (uint32_t)(LG_BN / ctas)));
  prm.tiles_m = (int)ceil_div(prm.sym ? a_rows : prm.m_rows, LG_BM * ctas);
  prm.ctas = ctas;
  // bands of 4 (pair) / 8 row tiles across ALL column tiles.  A 2-D super-tile raster (9 x 8 tiles ~ one wave, column groups
  // outermost) was measured for the P.W launch at cfg5: DRAM reads 7.6 -> 9.9 GB, time unchanged (4.31 vs 4.34 ms) -- not kept.
  prm.group_m = ctas == 2 ? 4 : 8;
  prm.group_n = prm.tiles_n;
  const int64_t total = prm.sym ? lg_sym_tiles(prm.tiles_m, prm.tiles_n, ctas) : (int64_t)prm.tiles_m * prm.tiles_n * prm.nsplit;
  NSVGD_REQUIRE(total < (1ll << 31), NSVGD_ERR_INVALID, "phi_large: too many tiles");
  prm.total = (int)total;
  const int units = (int)std::min<int64_t>(total, std::max(1, device_sm_count() / ctas));
  if (ctas == 2) {
    NSVGD_PROPAGATE(ensure_dyn_smem((const void*)lg_gemm_pair_kernel, LgSmem<2>::BYTES));
    lg_gemm_pair_kernel<<<units * 2, LG_THREADS, LgSmem<2>::BYTES, st>>>(mA0, mA1, mB0, mB1, prm);
  } else {
    NSVGD_PROPAGATE(ensure_dyn_smem((const void*)lg_gemm_kernel, LgSmem<1>::BYTES));
    lg_gemm_kernel<<<units, LG_THREADS, LgSmem<1>::BYTES, st>>>(mA0, mA1, mB0, mB1, prm);
  }
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

// median filter for d > 128 (median_large.cu): one symmetric Gram launch with the classify epilogue
int lg_launch_classify(cudaStream_t st, const void* Zb0, const void* Zb1, int64_t n, int64_t n_pad, int64_t ldk, const float* zn,
                       FastMedState* fs, unsigned short* tlist, uint32_t* tcount, uint32_t tcap, float kappa_g, float kappa_d,
                       uint32_t* err_flag) {
  LgParams g{};
  g.epi = LG_EPI_CLASSIFY; g.tiles_n = (int)(n_pad / LG_BN); g.nsplit = 1; g.sym = 1; g.diag_off = -1;
  g.kb_total = g.kb_per_split = (int)ceil_div(ldk, LG_BK);
  g.m_rows = (int)n_pad; g.row_off = 0;
  g.xn = zn; g.fs = fs; g.tlist = tlist; g.tcount = tcount; g.tcap = tcap; g.kappa_g = kappa_g; g.kappa_d = kappa_d; g.n_valid = (int)n;
  g.err_flag = err_flag;
  return lg_launch(st, lg_ctas(), (const __nv_bfloat16*)Zb0, (const __nv_bfloat16*)Zb1, n_pad, ldk, (const __nv_bfloat16*)Zb0,
                   (const __nv_bfloat16*)Zb1, n_pad, ldk, ldk, g);
}
int lg_tile_rows() { return LG_BM; }

// one warp per row: centred bf16 pieces, |xc|^2 and c_j = s_j . xc_j; rows in [n, rows_pad) are zero particles with
// |xc|^2 = 1e30, which makes their k underflow to exactly 0
__global__ void __launch_bounds__(256) large_prep_kernel(const float* __restrict__ X, const float* __restrict__ S, int64_t n,
                                                         int64_t rows_pad, int64_t d, int64_t ldx, int64_t ldk,
                                                         const float* __restrict__ mu, __nv_bfloat16* __restrict__ Xb0,
                                                         __nv_bfloat16* __restrict__ Xb1, float* __restrict__ xn,
                                                         float* __restrict__ cv) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows_pad) return;
  const bool rv = row < n;
  float acc_n = 0.f, acc_c = 0.f;
  for (int64_t c0 = 2 * lane; c0 < ldk; c0 += 64) {
    float v2[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int64_t c = c0 + e;
      if (rv && c < d) {
        const float v = X[row * ldx + c] - mu[c];
        acc_n = fmaf(v, v, acc_n);
        v2[e] = v;
        if (S) acc_c = fmaf(S[row * ldx + c], v, acc_c);
      }
    }
    const uint32_t p0 = pack_bf16x2(v2[0], v2[1]);
    const uint32_t p1 = pack_bf16x2(v2[0] - bf16lo_to_float(p0), v2[1] - bf16hi_to_float(p0));
    *reinterpret_cast<uint32_t*>(&Xb0[row * ldk + c0]) = p0;
    *reinterpret_cast<uint32_t*>(&Xb1[row * ldk + c0]) = p1;
  }
  acc_n = warp_sum(acc_n); acc_c = warp_sum(acc_c);
  if (lane == 0) {
    xn[row] = rv ? acc_n : 1e30f;
    if (cv) cv[row] = rv ? acc_c : 0.f;
  }
}

// W^T pieces, K-major for launch 2: Wt[c][j] = S[j][c] - (X[j][c] - mu[c]) / h^2 ; 64 particles x 32 coordinates per CTA
__global__ void __launch_bounds__(256) large_wt_kernel(const float* __restrict__ X, const float* __restrict__ S, int64_t n,
                                                       int64_t d, int64_t ldx, const float* __restrict__ mu,
                                                       const float* __restrict__ h_dev, __nv_bfloat16* __restrict__ Wt0,
                                                       __nv_bfloat16* __restrict__ Wt1, int64_t ldw) {
  __shared__ float t[32][65];
  const float h = *h_dev;
  const float inv_h2 = 1.f / (h * h);
  const int64_t j0 = (int64_t)blockIdx.x * 64, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = c0 + tx;
  const float m = c < d ? mu[c] : 0.f;
#pragma unroll
  for (int jj = ty; jj < 64; jj += 8) {
    const int64_t j = j0 + jj;
    float w = 0.f;
    if (j < n && c < d) w = fmaf(-(X[j * ldx + c] - m), inv_h2, S[j * ldx + c]);
    t[tx][jj] = w;
  }
  __syncthreads();
  const int cc = threadIdx.x >> 3, jj0 = (threadIdx.x & 7) * 8;
  if (c0 + cc < d) {
    uint32_t w0[4], w1[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = t[cc][jj0 + 2 * e], b = t[cc][jj0 + 2 * e + 1];
      w0[e] = pack_bf16x2(a, b);
      w1[e] = pack_bf16x2(a - bf16lo_to_float(w0[e]), b - bf16hi_to_float(w0[e]));
    }
    *reinterpret_cast<uint4*>(&Wt0[(c0 + cc) * ldw + j0 + jj0]) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
    *reinterpret_cast<uint4*>(&Wt1[(c0 + cc) * ldw + j0 + jj0]) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
  }
}

// [S | Xc]^T pieces for the aux split (stein.phistar's [drift, repulsion], stein.py:206-222): rows [0, d) = S^T, rows
// [dpad, dpad + d) = Xc^T; 64 particles x 32 coordinates per CTA, blockIdx.z selects the block
__global__ void __launch_bounds__(256) large_vt_kernel(const float* __restrict__ X, const float* __restrict__ S, int64_t n,
                                                       int64_t d, int64_t ldx, const float* __restrict__ mu, int64_t dpad,
                                                       __nv_bfloat16* __restrict__ Vt0, __nv_bfloat16* __restrict__ Vt1, int64_t ldw) {
  __shared__ float t[32][65];
  const int64_t j0 = (int64_t)blockIdx.x * 64, c0 = (int64_t)blockIdx.y * 32;
  const bool xblock = blockIdx.z == 1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = c0 + tx;
  const float m = (xblock && c < d) ? mu[c] : 0.f;
#pragma unroll
  for (int jj = ty; jj < 64; jj += 8) {
    const int64_t j = j0 + jj;
    float w = 0.f;
    if (j < n && c < d) w = xblock ? X[j * ldx + c] - m : S[j * ldx + c];
    t[tx][jj] = w;
  }
  __syncthreads();
  const int cc = threadIdx.x >> 3, jj0 = (threadIdx.x & 7) * 8;
  if (c0 + cc < d) {
    uint32_t w0[4], w1[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = t[cc][jj0 + 2 * e], b = t[cc][jj0 + 2 * e + 1];
      w0[e] = pack_bf16x2(a, b);
      w1[e] = pack_bf16x2(a - bf16lo_to_float(w0[e]), b - bf16hi_to_float(w0[e]));
    }
    const int64_t row = (xblock ? dpad : 0) + c0 + cc;
    *reinterpret_cast<uint4*>(&Vt0[row * ldw + j0 + jj0]) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
    *reinterpret_cast<uint4*>(&Vt1[row * ldw + j0 + jj0]) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
  }
}

int phi_large(cudaStream_t st, const PhiArgs& a, void* ws, size_t ws_bytes, bool dry, size_t* need) {
  const int ctas = lg_ctas();
  const int64_t ldk = round_up(a.d, 8);
  const int64_t n_pad = round_up(a.n, LG_BN);
  // aux split requested: the second GEMM runs against [S | Xc] (2 d columns, GRAM form of the epilogue) instead of W
  const bool want_aux = a.aux != nullptr;
  const int64_t dpad = round_up(a.d, 4);
  const int64_t v_rows = (want_aux || dry) ? dpad + a.d : a.d;       // rows of the K-major B operand of the second GEMM
  const int64_t ldo = (want_aux || dry) ? 2 * dpad : dpad;
  const int nsplit = (int)ceil_div(n_pad, LG_NK);
  const int tiles_nx = (int)(n_pad / LG_BN);                 // 256-particle column tiles of the Gram
  // the queries are a row range of the particle array (the SVGD step: all of it; a row-sharded rank: its block of the
  // gathered array): their pieces and norms are the particles', and the pairs (i, diag_off + i) are self pairs
  int64_t diag_off = -1;
  if (a.ldq == a.ldx && a.Q >= a.X && (a.Q - a.X) % a.ldx == 0 && (a.Q - a.X) / a.ldx + a.m <= a.n) diag_off = (a.Q - a.X) / a.ldx;
  const bool q_in_x = diag_off >= 0;
  const bool q_is_x = q_in_x && a.m == a.n;
  // query rows per pair of launches: the two pieces of P (rows x n_pad bf16 each) stay within 4 GiB
  const int64_t mc = std::min<int64_t>(round_up(a.m, 256), std::max<int64_t>(256, ((4ll << 30) / (n_pad * 4)) / 256 * 256));
  WsCarver w(ws, ws_bytes, dry);
  float* mu = w.take<float>((size_t)a.d);
  double* mpart = w.take<double>((size_t)colmean_partials(a.n, a.d));
  __nv_bfloat16* Xb0 = w.take<__nv_bfloat16>((size_t)(n_pad * ldk));
  __nv_bfloat16* Xb1 = w.take<__nv_bfloat16>((size_t)(n_pad * ldk));
  __nv_bfloat16* Wt0 = w.take<__nv_bfloat16>((size_t)(v_rows * n_pad));
  __nv_bfloat16* Wt1 = w.take<__nv_bfloat16>((size_t)(v_rows * n_pad));
  float* xn = w.take<float>((size_t)n_pad);
  float* cv = w.take<float>((size_t)n_pad);
  __nv_bfloat16 *Qb0 = Xb0, *Qb1 = Xb1;
  float* qn = xn;
  if (!q_in_x || dry) {
    Qb0 = w.take<__nv_bfloat16>((size_t)(a.m * ldk));
    Qb1 = w.take<__nv_bfloat16>((size_t)(a.m * ldk));
    qn = w.take<float>((size_t)a.m);
  }
  __nv_bfloat16* P0 = w.take<__nv_bfloat16>((size_t)(mc * n_pad));
  __nv_bfloat16* P1 = w.take<__nv_bfloat16>((size_t)(mc * n_pad));
  float* O = w.take<float>((size_t)nsplit * a.m * ldo);
  // Q == X in one row chunk: the Gram / pair launch computes the tiles on or above the diagonal only and mirrors them
  const bool sym = q_is_x && mc >= n_pad && lg_sym_enabled();
  const int scal_parts = sym || dry ? 2 * tiles_nx : tiles_nx;
  float* scal = w.take<float>((size_t)scal_parts * a.m * 4);
  double* partials = w.take<double>((size_t)(2 * epilogue_blocks(a.m)));
  uint32_t* err_flag = w.take<uint32_t>(64);
  if (need) *need = w.off;
  if (dry) return NSVGD_OK;
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "phi_large: workspace too small (%zu < %zu)", ws_bytes, w.off);
  NSVGD_REQUIRE(n_pad < (1ll << 31) && a.m < (1ll << 31) && ldk < (1ll << 31), NSVGD_ERR_INVALID, "phi_large: shape too large");

  NSVGD_CUDA(cudaMemsetAsync(err_flag, 0, 64 * sizeof(uint32_t), st));
  NSVGD_PROPAGATE(launch_colmean(st, a.X, a.n, a.d, a.ldx, mpart, mu));
  large_prep_kernel<<<(unsigned)ceil_div(n_pad * 32, 256), 256, 0, st>>>(a.X, a.S, a.n, n_pad, a.d, a.ldx, ldk, mu, Xb0, Xb1, xn, cv);
  NSVGD_LAUNCH_CHECK();
  if (want_aux) {
    if (dpad != a.d) {      // the (at most 3) rows between the two blocks
      NSVGD_CUDA(cudaMemsetAsync(Wt0 + a.d * n_pad, 0, (size_t)((dpad - a.d) * n_pad) * sizeof(__nv_bfloat16), st));
      NSVGD_CUDA(cudaMemsetAsync(Wt1 + a.d * n_pad, 0, (size_t)((dpad - a.d) * n_pad) * sizeof(__nv_bfloat16), st));
    }
    large_vt_kernel<<<dim3((unsigned)(n_pad / 64), (unsigned)ceil_div(a.d, 32), 2), 256, 0, st>>>(a.X, a.S, a.n, a.d, a.ldx, mu, dpad,
                                                                                               Wt0, Wt1, n_pad);
  } else {
    large_wt_kernel<<<dim3((unsigned)(n_pad / 64), (unsigned)ceil_div(a.d, 32)), 256, 0, st>>>(a.X, a.S, a.n, a.d, a.ldx, mu, a.h_dev,
                                                                                            Wt0, Wt1, n_pad);
  }
  NSVGD_LAUNCH_CHECK();
  if (q_in_x) { Qb0 = Xb0 + diag_off * ldk; Qb1 = Xb1 + diag_off * ldk; qn = xn + diag_off; }
  else {
    large_prep_kernel<<<(unsigned)ceil_div(a.m * 32, 256), 256, 0, st>>>(a.Q, nullptr, a.m, a.m, a.d, a.ldq, ldk, mu, Qb0, Qb1, qn,
                                                                        nullptr);
    NSVGD_LAUNCH_CHECK();
  }
  prof_begin(0, st);
  for (int64_t i0 = 0; i0 < a.m; i0 += mc) {
    const int64_t mi = std::min<int64_t>(mc, a.m - i0);
    LgParams g{};
    g.epi = LG_EPI_PAIR; g.tiles_n = tiles_nx; g.nsplit = 1; g.sym = sym ? 1 : 0; g.diag_off = diag_off;
    g.kb_total = g.kb_per_split = (int)ceil_div(ldk, LG_BK);
    g.m_rows = (int)mi; g.row_off = i0;
    g.qn = qn; g.xn = xn; g.cv = cv; g.h_dev = a.h_dev; g.P0 = P0; g.P1 = P1; g.ldp = n_pad; g.scal = scal; g.scal_stride = a.m;
    g.err_flag = err_flag;
    if (sym) NSVGD_CUDA(cudaMemsetAsync(scal, 0, (size_t)scal_parts * a.m * 4 * sizeof(float), st));   // slots no tile writes
    NSVGD_PROPAGATE(lg_launch(st, ctas, Qb0 + i0 * ldk, Qb1 + i0 * ldk, sym ? n_pad : mi, ldk, Xb0, Xb1, n_pad, ldk, ldk, g));
    LgParams o{};
    o.epi = LG_EPI_STORE; o.tiles_n = (int)ceil_div(v_rows, LG_BN); o.nsplit = nsplit;
    o.kb_total = (int)(n_pad / LG_BK); o.kb_per_split = (int)(LG_NK / LG_BK);
    o.m_rows = (int)mi; o.row_off = i0;
    o.O = O + i0 * ldo; o.ldo = ldo; o.o_split_stride = a.m * ldo; o.n_cols = (int)ldo;
    o.err_flag = err_flag;
    NSVGD_PROPAGATE(lg_launch(st, ctas, P0, P1, mi, n_pad, Wt0, Wt1, v_rows, n_pad, n_pad, o));
  }
  prof_end(0, st);

  EpilogueArgs e{};
  e.form = want_aux ? 1 : 2; e.err_flag = err_flag; e.O = O; e.scal = scal; e.nsplit = nsplit; e.scal_parts = scal_parts; e.m_pad = a.m;
  e.ldo = ldo; e.dpad = want_aux ? dpad : 0; e.mu = mu;
  e.Q = a.Q; e.ldq = a.ldq; e.Sq = a.Sq; e.m = a.m; e.n = a.n; e.d = a.d; e.h_dev = a.h_dev;
  e.phi = a.phi; e.ldphi = a.ldphi; e.aux = a.aux; e.stats_dev = a.stats_dev; e.partials = partials;
  return launch_epilogue(st, e);
}

}  // namespace nsvgd
