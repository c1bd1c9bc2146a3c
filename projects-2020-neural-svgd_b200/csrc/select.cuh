// select.cuh — device state of the exact radix select (shared by pair_engine.cu and median_tc.cu)
#pragma once
#include "common.cuh"
#include "p2p.cuh"

namespace nsvgd {

struct SelectState {
  uint32_t prefix[2];          // bits decided so far for the two target ranks
  uint32_t mask;               // which bits are decided
  uint32_t nan_flag;           // any NaN distance -> h = NaN (jnp.median semantics)
  unsigned long long rank[2];  // remaining rank inside the current prefix
  int32_t is_zero[2];          // target is one of the n diagonal zeros (full / literal modes)
  float value[2];              // resolved order statistics
  float log_np1;
  int32_t done;
};
static_assert(sizeof(SelectState) <= NSVGD_SELECT_STATE_BYTES, "state too large");

// state of the tensor-core-filtered fast path (median_tc.cu)
struct FastMedState {
  unsigned long long below;    // pairs with exact distance < L
  uint32_t list_count;         // uncertain pairs emitted by the filter
  uint32_t cand_count;         // exact distances inside [L, U]
  uint32_t overflow;           // a list overflowed
  uint32_t fallback;           // 1 -> the bracket missed or a list overflowed: run the SIMT 3-pass select
  float L, U;
  uint32_t u_is_inf, pad;
};

// Warm start of the bracket [L, U] of the filtered select (caller-owned, persistent, 32 bytes, zero-initialised): the two order
// statistics of the previous call and the relative margin to put around them.  Only ever a HINT: the select verifies with exact
// counts that the target ranks fall among the candidates inside the bracket and otherwise runs its exact fallback, so the
// result is bit-identical with or without it.
struct MedianHint {
  float lo, hi;        // order statistics (squared distances) of the previous call
  float delta;         // relative margin for the next bracket
  uint32_t valid;      // 1: lo / hi / delta usable
  uint32_t seen;       // bit 0: a previous value exists (drift = change between two consecutive calls); bits 1-15: cool-down calls
                       // left after a miss; bits 16-23: penalty (the cool-down doubles with every further miss)
  uint32_t hits, misses;
  uint32_t key;        // (n, dz, target rank, last-coordinate weight) of the call that wrote lo / hi: a different problem -> no hint
};
static_assert(sizeof(MedianHint) == 32, "hint layout");

#ifdef __CUDACC__
__device__ __forceinline__ int pass_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
__device__ __forceinline__ int pass_bins(int pass) { return pass == 2 ? 1024 : 2048; }

// One radix-select decision from a histogram.  select_advance_one: warp 0 of the CTA resolves target `w` from hh[0..bins)
// (global or shared memory) and parks the result; select_advance_commit publishes both.  select_advance_body = the two
// targets from one global histogram [2][NSVGD_SELECT_BINS] (warp t resolves target t).  Shared by the plain advance kernel,
// the fused select and the kernels that first all-reduce the histogram over peer memory (p2p.cuh).
// Call from ALL threads of a CTA with blockDim.x >= 64.
template <typename HT>
__device__ __forceinline__ void select_scan_warp(const SelectState* s, const HT* hh, int pass, int w, int lane,
                                                 uint32_t* new_prefix, unsigned long long* new_rank) {
  const int nb = pass == 2 ? 1024 : 2048;
  const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
  const unsigned long long target = s->rank[w];
  const int per = nb / 32;
  unsigned long long lsum = 0;
  for (int b = 0; b < per; ++b) lsum += hh[lane * per + b];
  unsigned long long incl = lsum;   // inclusive scan over lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned long long excl = incl - lsum;
  const bool mine = target >= excl && target < incl;
  const unsigned int ball = __ballot_sync(0xffffffffu, mine);
  if (ball == 0u) {  // rank beyond the population (cannot happen for valid ranks)
    if (lane == 0) { new_prefix[w] = s->prefix[w]; new_rank[w] = target; }
  } else if (lane == (__ffs(ball) - 1)) {
    unsigned long long cum = excl;
    int b = 0;
    for (; b < per; ++b) {
      const unsigned long long c = hh[lane * per + b];
      if (target < cum + c) break;
      cum += c;
    }
    new_prefix[w] = s->prefix[w] | ((uint32_t)(lane * per + b) << shift);
    new_rank[w] = target - cum;
  }
}

struct SelectScratch { uint32_t new_prefix[2]; unsigned long long new_rank[2]; };
__device__ __forceinline__ SelectScratch* select_scratch() {
  __shared__ SelectScratch sc;
  return &sc;
}

// both targets from a shared-memory histogram of 32-bit counts [2][NSVGD_SELECT_BINS] (warp t resolves target t)
__device__ __forceinline__ void select_advance_smem32(SelectState* s, const uint32_t* hist32, int pass);

__device__ __forceinline__ void select_advance_commit(SelectState* s, int pass) {
  SelectScratch* sc = select_scratch();
  const int nb = pass == 2 ? 1024 : 2048;
  const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
  __syncthreads();
  if (threadIdx.x == 0) {
    s->prefix[0] = sc->new_prefix[0]; s->prefix[1] = sc->new_prefix[1];
    s->rank[0] = sc->new_rank[0]; s->rank[1] = sc->new_rank[1];
    s->mask |= ((uint32_t)(nb - 1)) << shift;
    if (pass == NSVGD_SELECT_PASSES - 1) {
      s->value[0] = s->is_zero[0] ? 0.f : __uint_as_float(s->prefix[0]);
      s->value[1] = s->is_zero[1] ? 0.f : __uint_as_float(s->prefix[1]);
      s->done = 1;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void select_advance_smem32(SelectState* s, const uint32_t* hist32, int pass) {
  SelectScratch* sc = select_scratch();
  const int w = threadIdx.x >> 5;
  if (w < 2) select_scan_warp(s, hist32 + (size_t)w * NSVGD_SELECT_BINS, pass, w, threadIdx.x & 31, sc->new_prefix, sc->new_rank);
  select_advance_commit(s, pass);
}

__device__ __forceinline__ void select_advance_body(SelectState* s, const unsigned long long* hist, int pass) {
  SelectScratch* sc = select_scratch();
  const int w = threadIdx.x >> 5;
  if (w < 2) select_scan_warp(s, hist + (size_t)w * NSVGD_SELECT_BINS, pass, w, threadIdx.x & 31, sc->new_prefix, sc->new_rank);
  select_advance_commit(s, pass);
}

// DEFINED-ORDER fp32 squared distance of one pair (oracle/svgd_oracle.py::pair_sqdist_f32):
// sequential over coordinates, separate mul and add, last coordinate weighted by w.
__device__ __forceinline__ float exact_pair_sqdist(const float* __restrict__ a, const float* __restrict__ b,
                                                    int64_t d, float w_last) {
  float acc = 0.f;
  for (int64_t c = 0; c + 1 < d; ++c) {
    const float diff = a[c] - b[c];
    acc = __fadd_rn(acc, __fmul_rn(diff, diff));
  }
  const float diff = a[d - 1] - b[d - 1];
  float t = __fmul_rn(diff, diff);
  t = __fmul_rn(w_last, t);
  return __fadd_rn(acc, t);
}
#endif

// pair_engine.cu
int launch_select_init(cudaStream_t st, SelectState* s, long long r_lo, long long r_hi, float log_np1);
int launch_buffer_hist(cudaStream_t st, const float* vals, const uint32_t* counts_dev, int count_stride, int nseg,
                       uint32_t seg_cap, uint32_t count_host, const SelectState* state, int pass, uint64_t* hist,
                       const uint32_t* run_flag);
int launch_select_advance(cudaStream_t st, SelectState* s, const uint64_t* hist, int pass, const uint32_t* run_flag);
int launch_select_advance_p2p(cudaStream_t st, const P2PComm& c, SelectState* s, uint64_t* hist, int pass, const uint32_t* run_flag);
int launch_pair_hist(cudaStream_t st, const float* Z, int64_t n, int64_t dz, int64_t ldz, float w_last,
                     int64_t tile_row_start, int64_t tile_row_stride, const SelectState* state, int pass,
                     uint64_t* hist, const uint32_t* run_flag, float* dstore = nullptr);
void median_ranks_host(int64_t n, int mode, long long* r_lo, long long* r_hi);

// median_tc.cu: tensor-core filtered exact median.  Returns NSVGD_ERR_UNSUPPORTED when the shape is
// outside its range (caller then uses the SIMT 3-pass select).
bool median_fast_supported(int64_t n, int64_t dz);
size_t median_fast_workspace(int64_t n, int64_t dz);
int median_fast(cudaStream_t st, const float* P, int64_t n, int64_t dz, int64_t ldp, float w_last, long long r_lo,
                long long r_hi, int64_t tile_row_start, int64_t tile_row_stride, SelectState* main_state, uint64_t* hist,
                void* ws, size_t ws_bytes, MedianHint* hint = nullptr);

// shared by the two filtered selects (median_tc.cu, median_large.cu)
constexpr int MF_NSEG = 64;          // independent append segments (one atomic counter per 128-byte line)
constexpr int MF_CSTRIDE = 32;       // counter stride in uint32
// scratch of one fused 3-pass select (fused_select_kernel): one global histogram per pass + the grid-wide hand-shake
struct FusedSelScratch {
  uint32_t hist[NSVGD_SELECT_PASSES][2 * NSVGD_SELECT_BINS];   // 32-bit counts (a value buffer holds < 2^32 values over all ranks):
                                                               // two counts per 64-bit word of the peer-memory exchange
  uint32_t done[4];      // CTAs that have flushed their histogram of pass p
  uint32_t flag[4];      // set by the last CTA of pass p once the select state has advanced
};
int launch_fused_select(cudaStream_t st, const float* vals, const uint32_t* counts_dev, int count_stride, int nseg,
                        uint32_t seg_cap, uint32_t v_begin, uint32_t v_end, SelectState* state, FusedSelScratch* sc,
                        const P2PComm* comm, int final_mode, FastMedState* fs, SelectState* main_state, int l_is_zero,
                        int u_is_inf, const uint32_t* skip_flag = nullptr);
int launch_fast_pack_resolve(cudaStream_t st, FastMedState* fs, const uint32_t* list_counts, uint32_t list_cap, uint32_t* cand_counts,
                             uint32_t cand_cap, const uint32_t* err_flag, SelectState* main_state, SelectState* cand_state,
                             long long* red);

// warm-started bracket (median_tc.cu): which problem a hint belongs to, arm (sets *skip_flag and the bracket in fs), update
uint32_t median_hint_key(int64_t n, int64_t dz, long long r_lo, float w_last);
int launch_hint_apply(cudaStream_t st, const MedianHint* hint, FastMedState* fs, uint32_t* skip_flag, uint32_t key);
int launch_hint_update(cudaStream_t st, MedianHint* hint, const SelectState* main_state, const FastMedState* fs,
                       const uint32_t* skip_flag, uint32_t key);

// median_large.cu: the tensor-core filtered exact median for d > 128 (single GPU); phi_large.cu: its Gram + classify launch
bool median_large_supported(int64_t n, int64_t dz);
size_t median_large_workspace(int64_t n, int64_t dz);
int median_large(cudaStream_t st, const float* P, int64_t n, int64_t dz, int64_t ldp, float w_last, long long r_lo, long long r_hi,
                 SelectState* main_state, uint64_t* hist, void* ws, size_t ws_bytes, MedianHint* hint = nullptr);
int lg_launch_classify(cudaStream_t st, const void* Zb0, const void* Zb1, int64_t n, int64_t n_pad, int64_t ldk, const float* zn,
                       FastMedState* fs, unsigned short* tlist, uint32_t* tcount, uint32_t tcap, float kappa_g, float kappa_d,
                       uint32_t* err_flag);

// phi_tc.cu: centre + bf16 2-piece split (+ optional transposed [S|Xc] pieces); last_scale multiplies the
// centred last coordinate (literal median mode)
int launch_tc_prep(cudaStream_t st, const float* X, const float* S, int64_t n, int64_t d, int64_t ldx, int64_t rows_pad,
                   int dp, const float* mu, float last_scale, void* Xb0, void* Xb1, void* Xh, void* Vt0, void* Vt1,
                   int vt_extra, float* xn, float* cv, const float* h_combine, uint32_t* max_bits);

}  // namespace nsvgd
