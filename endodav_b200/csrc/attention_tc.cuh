// tcgen05 flash attention for the spatial (per-frame) self-attention.  (placeholder: until the
// tensor-core kernel lands this forwards to the CUDA-core kernel on the same 16-bit operands)
#pragma once
#include "attention_simt.cuh"
#include "launch.h"
#include "tc_common.cuh"

namespace tc {
template <typename T>
void launch_attention_tc(edv::Launch& L, int dtype, const void* qkv, void* out, int F, int S, int heads) {
  dim3 grid((S + 127) / 128, heads, F);
  spatial_attention_simt_kernel<T><<<grid, 128, 0, L.stream>>>((const T*)qkv, (T*)out, S, heads);
  L.check("spatial_attention(simt placeholder)");
}
}  // namespace tc
