// xla_ffi_shim.cc — thin XLA FFI (jax.ffi) custom-call layer over the C ABI of include/nsvgd.h.
//
// The reference keeps its JAX-facing signatures (kernels.median_heuristic, stein.get_phistar, ...)
// and dispatches into libnsvgd.so through these handlers (INTEGRATION.md shows the Python side).
// The XLA FFI headers (xla/ffi/api/ffi.h, shipped inside jaxlib) are NOT present in this image, so
// THIS FILE HAS NEVER BEEN THROUGH A COMPILER (rounds 1 and 2 alike); it is compile-guarded: `make ffi XLA_INCLUDE=$(python -c "import jax.ffi; print(jax.ffi.include_dir())")`
// builds nsvgd/libnsvgd_ffi.so where jaxlib exists.  Nothing here computes: every handler forwards
// device pointers, the XLA stream and an XLA scratch allocation to the C ABI.
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime.h>

#include "../../include/nsvgd.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error status_to_error(int st) {
  if (st == NSVGD_OK) return ffi::Error::Success();
  const auto code = st == NSVGD_ERR_INVALID ? ffi::ErrorCode::kInvalidArgument
                    : st == NSVGD_ERR_UNSUPPORTED ? ffi::ErrorCode::kUnimplemented
                                                  : ffi::ErrorCode::kInternal;
  return ffi::Error(code, nsvgd_last_error());
}

// h = median_heuristic(x)                                                        kernels.py:160-172
static ffi::Error MedianImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F32> x, int32_t mode,
                             ffi::ResultBuffer<ffi::F32> h) {
  const auto dims = x.dimensions();
  if (dims.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "median_heuristic: x must be (n, d)");
  const int64_t n = dims[0], d = dims[1];
  size_t bytes = 0;
  if (int st = nsvgd_workspace_bytes(NSVGD_OP_MEDIAN, n, n, d, NSVGD_PATH_AUTO, &bytes)) return status_to_error(st);
  auto ws = scratch.Allocate(bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "median_heuristic: scratch allocation failed");
  return status_to_error(nsvgd_median_h_f32(stream, x.typed_data(), n, d, d, mode, h->typed_data(), *ws, bytes));
}

// phi[i] = phi*(q_i) for all rows of q (vmap of stein.get_phistar folded into ONE call)      stein.py:180-203
// Attributes: want_aux = 1 -> the [drift, repulsion] split of stein.phistar is written (slower [S | Xc] form; d > 128 then runs
// the generic path), 0 -> `aux` is a 1-element dummy result and the COMBINED form runs; want_stats = 1 -> `sq` holds the scores
// AT THE QUERIES (an explicit argument: equal row counts do not make q the same points as x) and `stats` is written.
static ffi::Error PhiImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F32> q,
                          ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> s, ffi::Buffer<ffi::F32> sq, ffi::Buffer<ffi::F32> h,
                          int32_t path, int32_t want_aux, int32_t want_stats, ffi::ResultBuffer<ffi::F32> phi,
                          ffi::ResultBuffer<ffi::F32> aux, ffi::ResultBuffer<ffi::F64> stats) {
  const auto qd = q.dimensions(), xd = x.dimensions();
  if (qd.size() != 2 || xd.size() != 2 || qd[1] != xd[1])
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "phistar: q (m,d) and x (n,d) must share d");
  const int64_t m = qd[0], n = xd[0], d = xd[1];
  if (want_stats && (sq.dimensions().size() != 2 || sq.dimensions()[0] != m || sq.dimensions()[1] != d))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "phistar: want_stats needs sq of the shape of q");
  size_t bytes = 0;
  if (int st = nsvgd_workspace_bytes(NSVGD_OP_PHI, n, m, d, path, &bytes)) return status_to_error(st);
  auto ws = scratch.Allocate(bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "phistar: scratch allocation failed");
  return status_to_error(nsvgd_phi_f32(stream, q.typed_data(), want_stats ? sq.typed_data() : nullptr, m, d, x.typed_data(),
                                       s.typed_data(), n, d, d, h.typed_data(), phi->typed_data(), d,
                                       want_aux ? aux->typed_data() : nullptr, want_stats ? stats->typed_data() : nullptr, path,
                                       *ws, bytes));
}

// one fused sgd step (models.py:143-167 with KernelGradient, models.py:654-689); x is donated (input_output_aliases)
static ffi::Error StepImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F32> x,
                           ffi::Buffer<ffi::F32> s, ffi::Buffer<ffi::F32> h, float lr, int32_t scaled, float lambda_reg,
                           int32_t path, ffi::ResultBuffer<ffi::F32> x_out, ffi::ResultBuffer<ffi::F32> grads,
                           ffi::ResultBuffer<ffi::F64> step_aux) {
  const auto xd = x.dimensions();
  const int64_t n = xd[0], d = xd[1];
  size_t bytes = 0;
  if (int st = nsvgd_workspace_bytes(NSVGD_OP_STEP, n, n, d, path, &bytes)) return status_to_error(st);
  auto ws = scratch.Allocate(bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "step: scratch allocation failed");
  if (x_out->typed_data() != x.typed_data() &&
      cudaMemcpyAsync(x_out->typed_data(), x.typed_data(), sizeof(float) * n * d, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "step: copying the particles into the output buffer failed");
  return status_to_error(nsvgd_step_sgd_f32(stream, x_out->typed_data(), s.typed_data(), n, d, d, h.typed_data(), lr, scaled,
                                            lambda_reg, grads->typed_data(), step_aux->typed_data(), path, *ws, bytes));
}

static ffi::Error KsdImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> sx,
                          ffi::Buffer<ffi::F32> y, ffi::Buffer<ffi::F32> sy, ffi::Buffer<ffi::F32> h, int32_t variant,
                          ffi::ResultBuffer<ffi::F64> out) {
  const int64_t n = x.dimensions()[0], d = x.dimensions()[1], m = y.dimensions()[0];
  size_t bytes = 0;
  if (int st = nsvgd_workspace_bytes(NSVGD_OP_KSD, n, m, d, 0, &bytes)) return status_to_error(st);
  auto ws = scratch.Allocate(bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "ksd: scratch allocation failed");
  return status_to_error(nsvgd_ksd_f32(stream, x.typed_data(), sx.typed_data(), n, d, y.typed_data(), sy.typed_data(), m, d, d,
                                       h.typed_data(), variant, out->typed_data(), *ws, bytes));
}

static ffi::Error MmdImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> y,
                          float h, int32_t funnel, ffi::ResultBuffer<ffi::F64> out) {
  const int64_t n = x.dimensions()[0], d = x.dimensions()[1], m = y.dimensions()[0];
  size_t bytes = 0;
  if (int st = nsvgd_workspace_bytes(NSVGD_OP_MMD, n, m, d, 0, &bytes)) return status_to_error(st);
  auto ws = scratch.Allocate(bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "mmd: scratch allocation failed");
  return status_to_error(nsvgd_mmd_f32(stream, x.typed_data(), n, d, y.typed_data(), m, d, d, h, funnel, out->typed_data(), *ws, bytes));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(nsvgd_ffi_median, MedianImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Attr<int32_t>("mode").Ret<ffi::Buffer<ffi::F32>>(),
                              {xla::ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(nsvgd_ffi_phi, PhiImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Attr<int32_t>("path")
                                  .Attr<int32_t>("want_aux").Attr<int32_t>("want_stats")
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F64>>(),
                              {xla::ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(nsvgd_ffi_step, StepImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("lr").Attr<int32_t>("scaled").Attr<float>("lambda_reg").Attr<int32_t>("path")
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F64>>(),
                              {xla::ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(nsvgd_ffi_ksd, KsdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Attr<int32_t>("variant")
                                  .Ret<ffi::Buffer<ffi::F64>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(nsvgd_ffi_mmd, MmdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Attr<float>("h").Attr<int32_t>("funnel")
                                  .Ret<ffi::Buffer<ffi::F64>>());
#else
// XLA FFI headers not available: build nothing (the tested boundary is the C ABI via ctypes).
extern "C" int nsvgd_ffi_unavailable(void) { return 1; }
#endif
