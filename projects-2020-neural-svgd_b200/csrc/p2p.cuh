// p2p.cuh — small all-reduce over NVLink peer memory, callable from INSIDE a consumer kernel.
//
// Row-sharded runs exchange only tiny integer/double vectors between the heavy kernels (radix-select histograms,
// candidate counts, the KSD/l2 sums).  An NCCL all-reduce costs a launch plus ~15-25 us of latency for
// each of them; here the CONSUMER kernel (select advance, bracket resolve, step scaling) performs the reduction itself:
// every rank owns a block in peer-mapped memory (cudaIpc), publishes its vector there, raises a flag in every peer's
// block (st.release.sys), waits for the peers' flags (ld.acquire.sys) and sums the peers' vectors over NVLink in rank
// order 0..W-1, so all ranks obtain bit-identical sums (integers trivially; doubles because the order is fixed).
//
// Block layout (one per rank, cudaMalloc'ed by nsvgd_p2p_alloc, zero-initialised):
//   [0, 64)      uint32 flags[16]   flags[r] = last epoch published by rank r
//   [64, 68)     uint32 epoch       (local) number of completed all-reduces
//   [68, 72)     uint32 timeout     (local) set when a peer did not arrive within ~2 s (results invalid, no hang)
//   [1024, ...)  two data slots of NSVGD_P2P_SLOT_BYTES, used alternately (epoch parity)
// Reuse is safe without a second barrier: a rank overwrites slot p at epoch e+2 only after it passed the barrier of
// epoch e+1, which every peer signals after it finished reading epoch e.  Epochs live in device memory, so the
// kernels can be captured in a CUDA graph and replayed.
#pragma once
#include <stdint.h>

#define NSVGD_P2P_MAX_WORLD 16
#define NSVGD_P2P_SLOT_WORDS 8192
#define NSVGD_P2P_SLOT_BYTES (NSVGD_P2P_SLOT_WORDS * 8)
#define NSVGD_P2P_BLOCK_BYTES (1024 + 2 * NSVGD_P2P_SLOT_BYTES)

namespace nsvgd {

struct P2PComm {
  int rank, world;
  unsigned char* blk[NSVGD_P2P_MAX_WORLD];   // blk[r]: rank r's block as mapped into THIS process
};

#ifdef __CUDACC__
__device__ __forceinline__ void p2p_st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t p2p_ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long p2p_ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// In-place sum of buf[0..nwords) over all ranks; ALL threads of ONE CTA must call it (blockDim.x >= world).
// f64 = false: unsigned 64-bit integer adds; true: double adds in rank order.
// Returns false (on every thread) when a peer did not arrive within the watchdog time: the block's timeout flag is set
// and buf is POISONED (NaN doubles / all-ones integers) so that the consumer cannot turn a partial sum into a
// normal-looking bandwidth or step; the callers additionally raise their own NaN flags.
template <int BATCH = 4>
__device__ __forceinline__ bool p2p_allreduce_cta(const P2PComm& c, unsigned long long* buf, int nwords, bool f64) {
  unsigned char* me = c.blk[c.rank];
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(me);
  volatile uint32_t* epoch_p = reinterpret_cast<volatile uint32_t*>(me + 64);
  const uint32_t epoch = *epoch_p + 1u;
  const size_t slot_off = 1024 + (size_t)(epoch & 1u) * NSVGD_P2P_SLOT_BYTES;
  unsigned long long* my_slot = reinterpret_cast<unsigned long long*>(me + slot_off);
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) my_slot[i] = buf[i];
  __threadfence_system();
  __syncthreads();
  int late = 0;
  if ((int)threadIdx.x < c.world) {
    p2p_st_release_sys(reinterpret_cast<uint32_t*>(c.blk[threadIdx.x]) + c.rank, epoch);      // publish to peer
    const long long t0 = clock64();
    while ((int32_t)(p2p_ld_acquire_sys(my_flags + threadIdx.x) - epoch) < 0) {               // wait for peer
      if (clock64() - t0 > 4000000000ll) { *reinterpret_cast<volatile uint32_t*>(me + 68) = 1u; late = 1; break; }
    }
  }
  const bool timed_out = __syncthreads_or(late) != 0;
  if (timed_out) {
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) buf[i] = f64 ? 0x7ff8000000000000ull : ~0ull;
    __syncthreads();
    if (threadIdx.x == 0) *epoch_p = epoch;
    return false;
  }
  // All peer loads of up to BATCH words are issued BEFORE the first add: the asm statements keep their order, so
  // `acc += load` per peer would wait for every NVLink round trip (~2 us) in turn -- 128 of them per thread for a
  // 32 KB histogram on 8 ranks.
  for (int i0 = threadIdx.x; i0 < nwords; i0 += BATCH * blockDim.x) {
    unsigned long long v[BATCH][NSVGD_P2P_MAX_WORLD];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int i = i0 + u * blockDim.x;
#pragma unroll
      for (int r = 0; r < NSVGD_P2P_MAX_WORLD; ++r)
        v[u][r] = (i < nwords && r < c.world)
                      ? p2p_ld_relaxed_sys(reinterpret_cast<const unsigned long long*>(c.blk[r] + slot_off) + i) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i >= nwords) continue;
      if (f64) {
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < NSVGD_P2P_MAX_WORLD; ++r)
          if (r < c.world) acc += __longlong_as_double((long long)v[u][r]);       // fixed rank order
        buf[i] = (unsigned long long)__double_as_longlong(acc);
      } else {
        unsigned long long acc = 0ull;
#pragma unroll
        for (int r = 0; r < NSVGD_P2P_MAX_WORLD; ++r) acc += v[u][r];
        buf[i] = acc;
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *epoch_p = epoch;
  return true;
}
#endif

// host: fills comm from (rank, world, blocks[world]); returns false on bad arguments
inline bool p2p_make_comm(P2PComm* c, int rank, int world, void* const* blocks) {
  if (!blocks || world < 1 || world > NSVGD_P2P_MAX_WORLD || rank < 0 || rank >= world) return false;
  c->rank = rank; c->world = world;
  for (int r = 0; r < NSVGD_P2P_MAX_WORLD; ++r) c->blk[r] = r < world ? (unsigned char*)blocks[r] : nullptr;
  for (int r = 0; r < world; ++r)
    if (!c->blk[r]) return false;
  return true;
}

}  // namespace nsvgd
