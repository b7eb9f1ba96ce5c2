// Warp-per-row device helpers over 256-channel rows.  Lane l holds channels [4l,4l+4) and
// [128+4l, 128+4l+4) so every global/shared access is a conflict-free 16-byte vector.
#pragma once
#include "common.cuh"

namespace seeme {

struct Row8 {
  float v[8];
};

__device__ __forceinline__ Row8 row_load(const float* __restrict__ p, int lane) {
  Row8 r;
  float4 a = reinterpret_cast<const float4*>(p)[lane];
  float4 b = reinterpret_cast<const float4*>(p)[lane + 32];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void row_store(float* __restrict__ p, int lane, const Row8& r) {
  reinterpret_cast<float4*>(p)[lane] = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  reinterpret_cast<float4*>(p)[lane + 32] = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
// bf16 split store of a row: hi = bf16(x), lo = bf16(x - hi) (lo may be null)
__device__ __forceinline__ void row_store_split(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int lane,
                                                const Row8& r) {
  __align__(8) __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = __float2bfloat16_rn(r.v[i]);
    l[i] = __float2bfloat16_rn(r.v[i] - __bfloat162float(h[i]));
  }
  reinterpret_cast<uint2*>(hi)[lane] = reinterpret_cast<const uint2*>(h)[0];
  reinterpret_cast<uint2*>(hi)[lane + 32] = reinterpret_cast<const uint2*>(h)[1];
  if (lo) {
    reinterpret_cast<uint2*>(lo)[lane] = reinterpret_cast<const uint2*>(l)[0];
    reinterpret_cast<uint2*>(lo)[lane + 32] = reinterpret_cast<const uint2*>(l)[1];
  }
}
__device__ __forceinline__ float row_dot(const Row8& a, const Row8& b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s = fmaf(a.v[i], b.v[i], s);
  return warp_sum(s);
}
// LayerNorm(256) with affine, biased variance, eps 1e-5 (F.layer_norm semantics)
__device__ __forceinline__ Row8 row_layernorm(const Row8& x, const float* __restrict__ g, const float* __restrict__ b, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x.v[i];
  const float mean = warp_sum(s) * (1.0f / 256.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float d = x.v[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / 256.0f) + 1e-5f);
  Row8 gg = row_load(g, lane), bb = row_load(b, lane), y;
#pragma unroll
  for (int i = 0; i < 8; ++i) y.v[i] = (x.v[i] - mean) * rstd * gg.v[i] + bb.v[i];
  return y;
}

}  // namespace seeme
