// p2p.cu — peer-memory plumbing (cudaIpc) + the stand-alone form of the fused all-reduce (p2p.cuh).
#include <string.h>

#include "common.cuh"
#include "p2p.cuh"

namespace nsvgd {

__global__ void __launch_bounds__(1024) p2p_allreduce_kernel(const P2PComm c, unsigned long long* buf, int nwords, int f64) {
  p2p_allreduce_cta<1>(c, buf, nwords, f64 != 0);
}

}  // namespace nsvgd

using namespace nsvgd;

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "nsvgd.h promises 64-byte handles");

extern "C" size_t nsvgd_p2p_block_bytes(void) { return NSVGD_P2P_BLOCK_BYTES; }

extern "C" int nsvgd_p2p_alloc(void** dev_ptr, unsigned char* handle_out) {
  NSVGD_REQUIRE(dev_ptr && handle_out, NSVGD_ERR_INVALID, "p2p_alloc: null pointer");
  void* p = nullptr;
  NSVGD_CUDA(cudaMalloc(&p, NSVGD_P2P_BLOCK_BYTES));
  NSVGD_CUDA(cudaMemset(p, 0, NSVGD_P2P_BLOCK_BYTES));
  NSVGD_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  NSVGD_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle_out, &h, 64);
  *dev_ptr = p;
  return NSVGD_OK;
}

extern "C" int nsvgd_p2p_open(const unsigned char* handle, void** peer_ptr) {
  NSVGD_REQUIRE(handle && peer_ptr, NSVGD_ERR_INVALID, "p2p_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* p = nullptr;
  NSVGD_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *peer_ptr = p;
  return NSVGD_OK;
}

extern "C" int nsvgd_p2p_close(void* peer_ptr) {
  if (peer_ptr) NSVGD_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return NSVGD_OK;
}

extern "C" int nsvgd_p2p_free(void* dev_ptr) {
  if (dev_ptr) NSVGD_CUDA(cudaFree(dev_ptr));
  return NSVGD_OK;
}

extern "C" int nsvgd_p2p_timeout_flag(const void* own_block, int* flag_out) {
  NSVGD_REQUIRE(own_block && flag_out, NSVGD_ERR_INVALID, "p2p_timeout_flag: null pointer");
  uint32_t v = 0;
  NSVGD_CUDA(cudaMemcpy(&v, (const unsigned char*)own_block + 68, 4, cudaMemcpyDeviceToHost));
  *flag_out = (int)v;
  return NSVGD_OK;
}

extern "C" int nsvgd_p2p_allreduce(void* stream, int rank, int world, void* const* blocks, void* buf, int64_t nwords,
                                   int is_f64) {
  P2PComm c;
  NSVGD_REQUIRE(p2p_make_comm(&c, rank, world, blocks), NSVGD_ERR_INVALID, "p2p_allreduce: bad communicator");
  NSVGD_REQUIRE(buf && nwords >= 1 && nwords <= NSVGD_P2P_SLOT_WORDS, NSVGD_ERR_INVALID,
                "p2p_allreduce: 1 <= nwords <= %d", NSVGD_P2P_SLOT_WORDS);
  p2p_allreduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(c, (unsigned long long*)buf, (int)nwords, is_f64);
  NSVGD_LAUNCH_CHECK();
  return NSVGD_OK;
}
