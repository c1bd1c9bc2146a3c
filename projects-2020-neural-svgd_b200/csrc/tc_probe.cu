// tc_probe.cu — tcgen05 bring-up probes (TEST-ONLY: built into nsvgd/libnsvgd_probe.so, not into the product library).
// Two single-CTA kernels that pin the descriptor / layout conventions of tc_common.cuh against plain matrix products:
// TMA (SWIZZLE_128B) -> shared memory -> tcgen05.mma SS -> TMEM -> tcgen05.ld, and tcgen05.st -> tcgen05.mma TS.
#include "select.cuh"
#include "tc_common.cuh"

namespace nsvgd {

using namespace tc;

// ---------------------------------------------------------------------------- bring-up probes
// probe_gram: G[128 x 64] = bf16(A[128 x 128]) bf16(B[64 x 128])^T through TMA -> smem -> tcgen05.mma SS -> TMEM -> ld
__global__ void __launch_bounds__(128) tc_probe_gram_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB, float* G,
                                                            uint32_t* err_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-B aligned, stays in the shared address space
  uint8_t* sA = smem;                 // 2 slabs x 16 KB
  uint8_t* sB = smem + 32768;         // 2 slabs x 8 KB
  uint64_t* bars = (uint64_t*)(smem + 49152);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2);
  volatile uint32_t* abort_s = tmem_slot + 1;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  Watch watch{abort_s, err_flag};
  if (threadIdx.x == 0) {
    *abort_s = 0u;
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(&bars[0], 49152);
    for (int kh = 0; kh < 2; ++kh) {
      tma_load_2d(sA + kh * 16384, &tmA, &bars[0], kh * 64, 0);
      tma_load_2d(sB + kh * 8192, &tmB, &bars[0], kh * 64, 0);
    }
    mbar_wait(&bars[0], 0, watch, 1);
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    uint32_t acc = 0;
    for (int kh = 0; kh < 2; ++kh) {
      const uint64_t ad = umma_desc_sw128(smem_u32(sA + kh * 16384));
      const uint64_t bd = umma_desc_sw128(smem_u32(sB + kh * 8192));
      for (int ks = 0; ks < 4; ++ks) {
        umma_ss(tmem, umma_desc_advance(ad, ks * 32), umma_desc_advance(bd, ks * 32), idesc, acc);
        acc = 1;
      }
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0, watch, 2);
  __syncwarp();
  tc_fence_after();
  uint32_t v[64];
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  tmem_ld32(tmem + lane_addr, v);
  tmem_ld32(tmem + lane_addr + 32, v + 32);
  tmem_wait_ld();
  const int row = warp * 32 + lane;
  for (int c = 0; c < 64; ++c) G[row * 64 + c] = __uint_as_float(v[c]);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

// probe_pv: O[128 x 256] = bf16(P[128 x 64]) bf16(V[64 x 256]) with P stored to TMEM by the threads
// (tcgen05.st, packed bf16 pairs) and consumed as the A operand of tcgen05.mma TS, N = 256.
__global__ void __launch_bounds__(128) tc_probe_pv_kernel(const __grid_constant__ CUtensorMap tmV, const float* P,
                                                          float* O, uint32_t* err_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-B aligned, stays in the shared address space
  uint8_t* sV = smem;                 // 256 rows x 128 B = 32 KB
  uint64_t* bars = (uint64_t*)(smem + 32768);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2);
  volatile uint32_t* abort_s = tmem_slot + 1;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  Watch watch{abort_s, err_flag};
  if (threadIdx.x == 0) {
    *abort_s = 0u;
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  const int row = warp * 32 + lane;
  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(&bars[0], 32768);
    tma_load_2d(sV, &tmV, &bars[0], 0, 0);
  }
  uint32_t pk[32];
  for (int c = 0; c < 32; ++c) pk[c] = pack_bf16x2(P[row * 64 + 2 * c], P[row * 64 + 2 * c + 1]);
  tmem_st32(tmem + lane_addr + 384, pk);
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && lane == 0) {
    mbar_wait(&bars[0], 0, watch, 1);
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
    const uint64_t bd = umma_desc_sw128(smem_u32(sV));
    for (int ks = 0; ks < 4; ++ks)
      umma_ts(tmem, tmem + 384 + ks * 8, umma_desc_advance(bd, ks * 32), idesc, ks > 0 ? 1u : 0u);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0, watch, 2);
  __syncwarp();
  tc_fence_after();
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + lane_addr + c0, v);
    tmem_wait_ld();
    for (int c = 0; c < 32; ++c) O[row * 256 + c0 + c] = __uint_as_float(v[c]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int cols,
                                   int transpose) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < rows * cols; t += gridDim.x * blockDim.x) {
    const int r = t / cols, c = t % cols;
    if (transpose) dst[c * rows + r] = __float2bfloat16_rn(src[t]);
    else dst[t] = __float2bfloat16_rn(src[t]);
  }
}

}  // namespace nsvgd

using namespace nsvgd;

extern "C" int nsvgd_tc_probe_gram(void* stream, const float* A, const float* B, float* G, void* ws, size_t ws_bytes) {
  cudaStream_t st = (cudaStream_t)stream;
  NSVGD_REQUIRE(A && B && G && ws, NSVGD_ERR_INVALID, "tc_probe_gram: null pointer");
  WsCarver w(ws, ws_bytes);
  __nv_bfloat16* Ab = w.take<__nv_bfloat16>(128 * 128, 1024);
  __nv_bfloat16* Bb = w.take<__nv_bfloat16>(64 * 128, 1024);
  uint32_t* err = w.take<uint32_t>(4);
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "tc_probe_gram: workspace too small");
  NSVGD_CUDA(cudaMemsetAsync(err, 0, 16, st));
  f32_to_bf16_kernel<<<64, 256, 0, st>>>(A, Ab, 128, 128, 0);
  NSVGD_LAUNCH_CHECK();
  f32_to_bf16_kernel<<<32, 256, 0, st>>>(B, Bb, 64, 128, 0);
  NSVGD_LAUNCH_CHECK();
  CUtensorMap ma, mb;
  NSVGD_PROPAGATE(make_tmap_bf16_2d(&ma, Ab, 128, 128, 256, 64, 128));
  NSVGD_PROPAGATE(make_tmap_bf16_2d(&mb, Bb, 128, 64, 256, 64, 64));
  const int smem = 49152 + 64 + 1024;
  NSVGD_CUDA(cudaFuncSetAttribute(tc_probe_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc_probe_gram_kernel<<<1, 128, smem, st>>>(ma, mb, G, err);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}

extern "C" int nsvgd_tc_probe_pv(void* stream, const float* P, const float* V, float* O, void* ws, size_t ws_bytes) {
  cudaStream_t st = (cudaStream_t)stream;
  NSVGD_REQUIRE(P && V && O && ws, NSVGD_ERR_INVALID, "tc_probe_pv: null pointer");
  WsCarver w(ws, ws_bytes);
  __nv_bfloat16* Vt = w.take<__nv_bfloat16>(256 * 64, 1024);
  uint32_t* err = w.take<uint32_t>(4);
  NSVGD_REQUIRE(w.ok(), NSVGD_ERR_WORKSPACE, "tc_probe_pv: workspace too small");
  NSVGD_CUDA(cudaMemsetAsync(err, 0, 16, st));
  f32_to_bf16_kernel<<<64, 256, 0, st>>>(V, Vt, 64, 256, 1);   // Vt[256 x 64]
  NSVGD_LAUNCH_CHECK();
  CUtensorMap mv;
  NSVGD_PROPAGATE(make_tmap_bf16_2d(&mv, Vt, 64, 256, 128, 64, 256));
  const int smem = 32768 + 64 + 1024;
  NSVGD_CUDA(cudaFuncSetAttribute(tc_probe_pv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc_probe_pv_kernel<<<1, 128, smem, st>>>(mv, P, O, err);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}
