// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Force-included (-include) when oracle/build_ref.py compiles the reference's own
// C++ sources *where they lie* under /root/reference/external/maskrcnn_benchmark/csrc.
// The 2019-era sources call AT_DISPATCH_FLOATING_TYPES(tensor.type(), ...)
// (cpu/ROIAlign_cpu.cpp:266, cpu/nms_cpu.cpp:95).  torch 2.11's dispatch macro
// resolves `::detail::scalar_type(the_type)` and no longer ships an overload for
// at::DeprecatedTypeProperties, so we supply that one overload here instead of
// patching (or copying) the reference sources.
#pragma once
#include <torch/extension.h>
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) {
  return t.scalarType();
}
}  // namespace detail
