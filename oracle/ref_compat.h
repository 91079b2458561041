// Force-included when compiling the reference's CUDA op against torch >= 2.x (oracle/build_ref.py).
// AT_DISPATCH_FLOATING_TYPES(value.type(), ...) in the reference
// (/root/reference/models/ops/src/cuda/ms_deform_attn_cuda.cu:64,134) needs this overload, which
// torch removed; supplying it here keeps the reference sources byte-for-byte untouched.
#pragma once
#include <ATen/ATen.h>
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
