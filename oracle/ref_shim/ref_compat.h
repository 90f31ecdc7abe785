// Build-time compatibility shim for compiling the UNMODIFIED reference SoftRas
// extension (third_party/softras/soft_renderer/cuda/soft_rasterize_cuda{.cpp,_kernel.cu})
// against torch 2.11.  Force-included with `-include`; no reference source is
// copied or edited.  It only repairs two host-side API drifts:
//   1. AT_DISPATCH_FLOATING_TYPES(tensor.type(), ...) -- torch 2.x wants a
//      ScalarType, the reference passes DeprecatedTypeProperties
//      (soft_rasterize_cuda_kernel.cu:701,716,780).
//   2. <torch/torch.h> no longer pulls in pybind11 (soft_rasterize_cuda.cpp:1,135).
// Device code is untouched.
#pragma once
#include <torch/extension.h>
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>

namespace lasr_b200_ref_compat {
inline at::ScalarType to_scalar_type(at::ScalarType t) { return t; }
inline at::ScalarType to_scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace lasr_b200_ref_compat

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...)                                   \
  AT_DISPATCH_SWITCH(::lasr_b200_ref_compat::to_scalar_type(TYPE), NAME,              \
                     AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
