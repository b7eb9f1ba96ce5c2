// Latent denoiser (MldDenoiser.forward, mld/models/architectures/mld_denoiser.py:151-244) and the
// DDIM sampling loop with classifier-free guidance (MLD._diffusion_reverse,
// mld/models/modeltype/mld.py:432-511; diffusers DDIMScheduler.step restated in SURVEY App. B).
//
// One denoiser call = 5 MotionDiffuse-style blocks (mdiff_transformer.py:286-304) in the 2-1-2 skip
// topology of cross_attention.py:67-83, on ONE latent token per row.  Algebra used:
//  H1  only token 0 of the 4-token self-attention is consumed (mdiff_transformer.py:295-297):
//      q / out-proj / LN / FFN run for token 0 only; K,V exist for {x, cond tokens, time token};
//  H2  the "linear attention" einsum pair (:231-237) is sum_n (softmax_d(q) . softmax_n(k)_n) v_n;
//  H3  K/V of the condition tokens do not change across steps -> computed once per run;
//  H4  the time embedding, the time token's K/V and every FiLM (scale, shift) depend on t only ->
//      tables with one row per timestep, built once per timestep list.
// The sampler is a chain of these kernels captured in a CUDA graph; the CFG combine and the DDIM
// update are fused into the kernel that applies the final LayerNorm.
#include "common.cuh"
#include "rowops.cuh"
#include "umma.cuh"
#include "denoiser.cuh"
#include <stdlib.h>

namespace seeme {

// x[r] = lat[r % B] + pe[0]      (torch.cat([latents]*2), mld.py:469-473; query_pos, mld_denoiser.py:210)
__global__ void den_prep_kernel(const float* __restrict__ lat, const float* __restrict__ pe0, float* __restrict__ x,
                                __nv_bfloat16* __restrict__ xh, __nv_bfloat16* __restrict__ xl, int R, int B) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  Row8 v = row_load(lat + (size_t)(row % B) * 256, lane);
  const Row8 p = row_load(pe0, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) v.v[i] += p.v[i];
  row_store(x + (size_t)row * 256, lane, v);
  row_store_split(xh + (size_t)row * 256, xl + (size_t)row * 256, lane, v);
}

// token-0 self-attention over keys {x_r, cond_0..cond_{Nc-1}, time}; qkv [R,768] (q pre-scaled),
// kvc [Nc*R,512] = (k|v) of the cond tokens (row n*R + r), tkv [512] = (k|v) of the time token.
// NC = number of cond tokens as a template parameter: with a run-time count the per-token arrays lived in local memory
template <int NC>
__global__ void den_sa_attn_kernel(const float* __restrict__ qkv, const float* __restrict__ kvc,
                                   const float* __restrict__ tkv, int Nc_unused, int R, __nv_bfloat16* __restrict__ oh,
                                   __nv_bfloat16* __restrict__ ol) {
  constexpr int Nc = NC;
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* base = qkv + (size_t)row * 768;
  const Row8 q = row_load(base, lane);
  float sc[NC + 2];
  sc[0] = row_dot(q, row_load(base + 256, lane));
#pragma unroll
  for (int n = 0; n < Nc; ++n) sc[1 + n] = row_dot(q, row_load(kvc + ((size_t)n * R + row) * 512, lane));
  sc[1 + Nc] = row_dot(q, row_load(tkv, lane));
  float m = sc[0];
#pragma unroll
  for (int j = 1; j < Nc + 2; ++j) m = fmaxf(m, sc[j]);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < Nc + 2; ++j) { sc[j] = expf(sc[j] - m); sum += sc[j]; }
  const float inv = 1.0f / sum;
  Row8 acc;
  {
    const Row8 v = row_load(base + 512, lane);
    const float p = sc[0] * inv;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = p * v.v[i];
  }
#pragma unroll
  for (int n = 0; n < Nc; ++n) {
    const Row8 v = row_load(kvc + ((size_t)n * R + row) * 512 + 256, lane);
    const float p = sc[1 + n] * inv;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(p, v.v[i], acc.v[i]);
  }
  {
    const Row8 v = row_load(tkv + 256, lane);
    const float p = sc[1 + Nc] * inv;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(p, v.v[i], acc.v[i]);
  }
  row_store_split(oh + (size_t)row * 256, ol + (size_t)row * 256, lane, acc);
}

// FiLM tail of a StylizationBlock (mdiff_transformer.py:152-163) before its out Linear:
//   h = SiLU(LN(y) * (1 + scale) + shift),  film = [scale(256) | shift(256)] for this timestep
__device__ __forceinline__ Row8 film_silu(const Row8& y, const float* __restrict__ film, const float* __restrict__ g,
                                          const float* __restrict__ b, int lane) {
  Row8 h = row_layernorm(y, g, b, lane);
  const Row8 sc = row_load(film, lane), sh = row_load(film + 256, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) h.v[i] = silu(h.v[i] * (1.0f + sc.v[i]) + sh.v[i]);
  return h;
}

// LinearTemporalCrossAttention core (mdiff_transformer.py:219-237) for one query token, H = 1:
//   qs = softmax_d(q); ks_n = softmax over the Nc tokens (per channel); y = sum_n (qs . ks_n) v_n
// followed by the FiLM tail.  q [R,256]; kv2 [Nc*R,512] = (key|value) rows n*R + r.
template <int NC>
__global__ void den_ca_kernel(const float* __restrict__ q, const float* __restrict__ kv2, int Nc_unused, int R,
                              const float* __restrict__ film, const float* __restrict__ g, const float* __restrict__ b,
                              __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  Row8 qs = row_load(q + (size_t)row * 256, lane);
  float m = qs.v[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, qs.v[i]);
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { qs.v[i] = expf(qs.v[i] - m); sum += qs.v[i]; }
  const float inv = 1.0f / warp_sum(sum);
#pragma unroll
  for (int i = 0; i < 8; ++i) qs.v[i] *= inv;
  constexpr int Nc = NC;
  Row8 ks[NC];
#pragma unroll
  for (int n = 0; n < Nc; ++n) ks[n] = row_load(kv2 + ((size_t)n * R + row) * 512, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float km = ks[0].v[i];
#pragma unroll
    for (int n = 1; n < Nc; ++n) km = fmaxf(km, ks[n].v[i]);
    float s = 0.f;
#pragma unroll
    for (int n = 0; n < Nc; ++n) { ks[n].v[i] = expf(ks[n].v[i] - km); s += ks[n].v[i]; }
    const float is = 1.0f / s;
#pragma unroll
    for (int n = 0; n < Nc; ++n) ks[n].v[i] *= is;
  }
  Row8 y;
#pragma unroll
  for (int i = 0; i < 8; ++i) y.v[i] = 0.f;
#pragma unroll
  for (int n = 0; n < Nc; ++n) {
    const float w = row_dot(qs, ks[n]);
    const Row8 v = row_load(kv2 + ((size_t)n * R + row) * 512 + 256, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) y.v[i] = fmaf(w, v.v[i], y.v[i]);
  }
  row_store_split(oh + (size_t)row * 256, ol + (size_t)row * 256, lane, film_silu(y, film, g, b, lane));
}

__global__ void den_film_kernel(const float* __restrict__ y, const float* __restrict__ film, const float* __restrict__ g,
                                const float* __restrict__ b, __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol,
                                int R) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  row_store_split(oh + (size_t)row * 256, ol + (size_t)row * 256, lane,
                  film_silu(row_load(y + (size_t)row * 256, lane), film, g, b, lane));
}

// y = LN(x) -> fp32 and bf16 (hi, lo);  optionally a second LayerNorm chained on the result:
// y2 = LN2(y) -> bf16 only (norm2 of the self-attention block followed by the cross-attention's input norm)
__global__ void den_ln_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                              float* __restrict__ y, __nv_bfloat16* __restrict__ yh, __nv_bfloat16* __restrict__ yl,
                              const float* __restrict__ g2, const float* __restrict__ b2, __nv_bfloat16* __restrict__ y2h,
                              __nv_bfloat16* __restrict__ y2l, int R) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const Row8 v = row_layernorm(row_load(x + (size_t)row * 256, lane), g, b, lane);
  if (y) row_store(y + (size_t)row * 256, lane, v);
  if (yh) row_store_split(yh + (size_t)row * 256, yl + (size_t)row * 256, lane, v);
  if (g2) row_store_split(y2h + (size_t)row * 256, y2l + (size_t)row * 256, lane, row_layernorm(v, g2, b2, lane));
}

// eps = LN_final(x); [CFG] eps = eps_u + s (eps_c - eps_u) with u = rows [0,B), c = rows [B,2B)
// (mld.py:488-492); DDIM eta=0 update with unfused fp32 ops in diffusers' order:
//   x0 = (x - c0 eps) / c1 ; x_prev = c2 x0 + c3 eps
// coef = device [4] for this step, gscale = device scalar.
__global__ void den_final_ddim_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                      float* __restrict__ lat, int B, int cfg, const float* __restrict__ coef,
                                      const float* __restrict__ gscale) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  Row8 e = row_layernorm(row_load(x + (size_t)row * 256, lane), g, b, lane);
  if (cfg) {
    const Row8 ec = row_layernorm(row_load(x + (size_t)(B + row) * 256, lane), g, b, lane);
    const float s = *gscale;
#pragma unroll
    for (int i = 0; i < 8; ++i) e.v[i] = __fadd_rn(e.v[i], __fmul_rn(s, __fsub_rn(ec.v[i], e.v[i])));
  }
  const float c0 = coef[0], c1 = coef[1], c2 = coef[2], c3 = coef[3];
  Row8 l = row_load(lat + (size_t)row * 256, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float x0 = __fdiv_rn(__fsub_rn(l.v[i], __fmul_rn(c0, e.v[i])), c1);
    l.v[i] = __fadd_rn(__fmul_rn(c2, x0), __fmul_rn(c3, e.v[i]));
  }
  row_store(lat + (size_t)row * 256, lane, l);
}

}  // namespace seeme

using namespace seeme;

namespace {
constexpr int MAX_STEPS = DEN_MAX_STEPS;

size_t den_tensor_elems(int i) {
  if (i < DN_BLK) {
    static const size_t e[11] = {256 * 256, 256, 256 * 256, 256, 500 * 256, 256, 256, 256 * 512, 256, 256 * 512, 256};
    return e[i];
  }
  static const size_t e[38] = {768 * 256, 768, 256 * 256, 256, 1024 * 256, 1024, 256 * 1024, 256, 256, 256, 256, 256,
                               256, 256, 256, 256, 256 * 256, 256, 256 * 256, 256, 256 * 256, 256, 512 * 256, 512,
                               256, 256, 256 * 256, 256,
                               128 * 256, 128, 256 * 128, 256, 512 * 256, 512, 256, 256, 256 * 256, 256};
  return e[(i - DN_BLK) % DN_BLK_STRIDE];
}
}  // namespace


extern "C" int seeme_denoiser_create(seeme_denoiser_t* out, const float* const* w, int n_w, int max_rows) {
  SEEME_REQUIRE(out && w, SEEME_EINVAL, "seeme_denoiser_create: null argument");
  SEEME_REQUIRE(n_w == SEEME_DENOISER_NUM_TENSORS, SEEME_EINVAL, "seeme_denoiser_create: expected %d tensors, got %d",
                SEEME_DENOISER_NUM_TENSORS, n_w);
  SEEME_REQUIRE(max_rows > 0, SEEME_EINVAL, "seeme_denoiser_create: max_rows must be positive");
  seeme_denoiser* h = new seeme_denoiser();
  SEEME_CUDA(cudaGetDevice(&h->device));
  h->max_rows = max_rows;
  const char* ng = seeme_exp_env("SEEME_NO_GRAPH");
  h->use_graph = !(ng && ng[0] == '1');
  size_t wbytes = 0;
  for (int i = 0; i < n_w; ++i) wbytes += pad256(den_tensor_elems(i) * 4);
  const size_t R = max_rows, NC = SEEME_MAX_COND_TOKENS;
  size_t tbytes = 3 * pad256((size_t)MAX_STEPS * 256 * 4) + 15 * pad256((size_t)MAX_STEPS * 512 * 4);
  size_t ws = 2 * pad256(NC * R * 256 * 4) + 10 * pad256(NC * R * 512 * 4) + 12 * pad256(R * 256 * 4) + pad256(R * 768 * 4) +
              22 * pad256(R * 256 * 2) + 2 * pad256(R * 1024 * 2) + 2 * pad256(R * 128 * 2) + pad256(MAX_STEPS * 4 * 4) + 256;
  // bf16 (hi, lo) copies of the per-step weights: ~5.6 M parameters x 2 x 2 bytes
  const size_t pbytes = 2 * 2 * (size_t)(5 * (768 * 256 + 256 * 256 + 2 * 1024 * 256 + 3 * 256 * 256 + 2 * 128 * 256) + 2 * 256 * 512) + 64 * 1024;
  int rc = h->arena.init(wbytes + tbytes + ws + pbytes + 65536);
  if (rc) { delete h; return rc; }
  for (int i = 0; i < n_w; ++i) {
    size_t n = den_tensor_elems(i);
    h->w[i] = h->arena.take<float>(n);
    if (!h->w[i] || !w[i]) { set_error("seeme_denoiser_create: tensor %d null or arena exhausted", i); h->arena.release(); delete h; return SEEME_EINVAL; }
    cudaError_t e = cudaMemcpy(h->w[i], w[i], n * 4, cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) { set_error("seeme_denoiser_create: copy of tensor %d failed: %s", i, cudaGetErrorString(e)); h->arena.release(); delete h; return SEEME_ECUDA; }
  }
  for (int l = 0; l < 5; ++l) {   // fold 1/sqrt(256) into the q rows of in_proj
    scale_kernel_launch(blkw(h, l, SA_IN_W), 256 * 256, 0.0625f);
    scale_kernel_launch(blkw(h, l, SA_IN_B), 256, 0.0625f);
  }
  SEEME_CUDA(cudaDeviceSynchronize());
  h->sinus = h->arena.take<float>((size_t)MAX_STEPS * 256);
  h->t1 = h->arena.take<float>((size_t)MAX_STEPS * 256);
  h->temb = h->arena.take<float>((size_t)MAX_STEPS * 256);
  for (int l = 0; l < 5; ++l) {
    h->tkv[l] = h->arena.take<float>((size_t)MAX_STEPS * 512);
    h->film_ca[l] = h->arena.take<float>((size_t)MAX_STEPS * 512);
    h->film_ff[l] = h->arena.take<float>((size_t)MAX_STEPS * 512);
  }
  h->cond = h->arena.take<float>(NC * R * 256);
  h->tn = h->arena.take<float>(NC * R * 256);
  for (int l = 0; l < 5; ++l) { h->kvc[l] = h->arena.take<float>(NC * R * 512); h->kv2[l] = h->arena.take<float>(NC * R * 512); }
  h->lat = h->arena.take<float>(R * 256);
  h->t0 = h->arena.take<float>(R * 256);
  h->caq = h->arena.take<float>(R * 256);
  h->y = h->arena.take<float>(R * 256);
  h->qkv = h->arena.take<float>(R * 768);
  auto mk = [&](ActBuf& a, int ld, bool f32, bool b16) {
    a.ld = ld;
    a.f = f32 ? h->arena.take<float>(R * ld) : nullptr;
    a.h = b16 ? h->arena.take<__nv_bfloat16>(R * ld) : nullptr;
    a.l = b16 ? h->arena.take<__nv_bfloat16>(R * ld) : nullptr;
  };
  mk(h->x, 256, true, true);
  for (int l = 0; l < 5; ++l) mk(h->L[l], 256, true, true);
  mk(h->att, 256, false, true);
  mk(h->x1, 256, true, true);
  mk(h->x2, 256, true, false);
  mk(h->ln, 256, false, true);
  mk(h->hb, 256, false, true);
  mk(h->ff, 1024, false, true);
  mk(h->g1, 128, false, true);
  h->d_coef = h->arena.take<float>((size_t)MAX_STEPS * 4);
  h->d_gscale = h->arena.take<float>(1);
  if (!h->d_gscale) { set_error("seeme_denoiser_create: arena exhausted (workspace)"); h->arena.release(); delete h; return SEEME_ENOMEM; }
  // pack the per-step weights for the tcgen05 linears (q rows already carry the 1/16 attention scale)
  const char* pe = getenv("SEEME_DENOISER_PRECISION");
  h->npass = (pe && atoi(pe) == 1) ? 1 : 3;
  rc = SEEME_OK;
  for (int l = 0; l < 5 && !rc; ++l) {
    rc = pack_linear(h->arena, h->Wqkv[l], blkw(h, l, SA_IN_W), 256, 768, 256, blkw(h, l, SA_IN_B));
    if (!rc) rc = pack_linear(h->arena, h->Wout[l], blkw(h, l, SA_OUT_W), 256, 256, 256, blkw(h, l, SA_OUT_B));
    if (!rc) rc = pack_linear(h->arena, h->Wl1[l], blkw(h, l, SA_L1_W), 256, 1024, 256, blkw(h, l, SA_L1_B));
    if (!rc) rc = pack_linear(h->arena, h->Wl2[l], blkw(h, l, SA_L2_W), 1024, 256, 1024, blkw(h, l, SA_L2_B));
    if (!rc) rc = pack_linear(h->arena, h->Wcaq[l], blkw(h, l, CA_Q_W), 256, 256, 256, blkw(h, l, CA_Q_B));
    if (!rc) rc = pack_linear(h->arena, h->Wcaout[l], blkw(h, l, CA_OUT_W), 256, 256, 256, blkw(h, l, CA_OUT_B));
    if (!rc) rc = pack_linear(h->arena, h->Wf1[l], blkw(h, l, FF_L1_W), 256, 128, 256, blkw(h, l, FF_L1_B));
    if (!rc) rc = pack_linear(h->arena, h->Wf2[l], blkw(h, l, FF_L2_W), 128, 256, 128, blkw(h, l, FF_L2_B));
    if (!rc) rc = pack_linear(h->arena, h->Wfout[l], blkw(h, l, FF_OUT_W), 256, 256, 256, blkw(h, l, FF_OUT_B));
  }
  for (int i = 0; i < 2 && !rc; ++i)
    rc = pack_linear(h->arena, h->Wskip[i], h->w[DN_LB0_W + 2 * i], 512, 256, 512, h->w[DN_LB0_B + 2 * i]);
  if (!rc && cudaDeviceSynchronize() != cudaSuccess) { set_error("seeme_denoiser_create: weight packing failed"); rc = SEEME_ECUDA; }
  if (rc) { h->arena.release(); delete h; return rc; }
  SEEME_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
  {
    // SEEME_SAMPLER=graph keeps the round-1 back-end (a CUDA graph of ~3 750 kernels) for A/B measurements
    const char* sm = getenv("SEEME_SAMPLER");
    if (!(sm && strcmp(sm, "graph") == 0)) {
      rc = den_persist_create(h);
      if (rc) { den_persist_destroy(h); cudaStreamDestroy(h->cap_stream); h->arena.release(); delete h; return rc; }
    } else {
      h->backend = SEEME_SAMPLER_GRAPH;
    }
  }
  if (!(seeme_exp_env("SEEME_NO_CARVEOUT") && seeme_exp_env("SEEME_NO_CARVEOUT")[0] == '1')) {
    // the row-wise kernels use no shared memory; asking for the maximum carve-out anyway keeps the SMs in the
    // configuration of the 193 KB GEMM kernels they alternate with (no L1/shared re-partitioning between launches)
    cudaFuncSetAttribute(den_prep_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(den_sa_attn_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(den_sa_attn_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(den_sa_attn_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(den_sa_attn_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(den_ca_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(den_ca_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(den_ca_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(den_ca_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(den_film_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(den_ln_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(den_final_ddim_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  *out = h;
  return SEEME_OK;
}

// Build the per-timestep tables for `ts` (H4).  sinus_host: optional HOST [n,256] sinusoid
// ([cos | sin], embeddings.py:263-285) computed by the caller with the reference's own tensor ops;
// when NULL it is computed here in double precision and rounded to fp32.
static int den_build_tables(seeme_denoiser* h, const int* ts, int n, const float* sinus_host, cudaStream_t s) {
  SEEME_REQUIRE(n > 0 && n <= MAX_STEPS, SEEME_EINVAL, "denoiser: number of timesteps %d out of range (1..%d)", n, MAX_STEPS);
  std::vector<float> tmp;
  if (!sinus_host) {
    tmp.resize((size_t)n * 256);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < 128; ++i) {
        const float expo = (float)(-9.210340371976184) * (float)i / 128.0f;
        const float freq = (float)exp((double)expo);
        const float arg = (float)ts[j] * freq;
        tmp[(size_t)j * 256 + i] = (float)cos((double)arg);         // flip_sin_to_cos: cos first
        tmp[(size_t)j * 256 + 128 + i] = (float)sin((double)arg);
      }
    sinus_host = tmp.data();
  }
  SEEME_CUDA(cudaMemcpyAsync(h->sinus, sinus_host, (size_t)n * 256 * 4, cudaMemcpyHostToDevice, s));
  SEEME_CUDA(cudaStreamSynchronize(s));   // host staging buffer may be a temporary
  GemmP g = gemm_params(h->sinus, 256, h->w[DN_TE1_W], 256, h->w[DN_TE1_B], h->t1, 256, n, 256, 256);
  g.act = ACT_SILU;
  SEEME_TRY(gemm_f32(g, s));
  SEEME_TRY(gemm_f32(gemm_params(h->t1, 256, h->w[DN_TE2_W], 256, h->w[DN_TE2_B], h->temb, 256, n, 256, 256), s));
  for (int l = 0; l < 5; ++l) {
    // time token K,V: in_proj rows [256,768)
    SEEME_TRY(gemm_f32(gemm_params(h->temb, 256, blkw(h, l, SA_IN_W) + 256 * 256, 256, blkw(h, l, SA_IN_B) + 256, h->tkv[l],
                                   512, n, 512, 256), s));
    GemmP f1 = gemm_params(h->temb, 256, blkw(h, l, CA_EMB_W), 256, blkw(h, l, CA_EMB_B), h->film_ca[l], 512, n, 512, 256);
    f1.pre_act = ACT_SILU;
    SEEME_TRY(gemm_f32(f1, s));
    GemmP f2 = gemm_params(h->temb, 256, blkw(h, l, FF_EMB_W), 256, blkw(h, l, FF_EMB_B), h->film_ff[l], 512, n, 512, 256);
    f2.pre_act = ACT_SILU;
    SEEME_TRY(gemm_f32(f2, s));
  }
  if (h->persist) SEEME_TRY(den_persist_build_tables(h, n, s));
  h->table_ts.assign(ts, ts + n);
  return SEEME_OK;
}

// cond-token projections for all layers (H3); cond [Nc*R,256] in h->cond
static int den_cond_precompute(seeme_denoiser* h, int Nc, int R, cudaStream_t s) {
  const int rows = Nc * R;
  for (int l = 0; l < 5; ++l) {
    SEEME_TRY(gemm_f32(gemm_params(h->cond, 256, blkw(h, l, SA_IN_W) + 256 * 256, 256, blkw(h, l, SA_IN_B) + 256, h->kvc[l], 512,
                                   rows, 512, 256), s));
    SEEME_TRY(layernorm256(h->cond, nullptr, 0, blkw(h, l, CA_TN_W), blkw(h, l, CA_TN_B), h->tn, rows, s));
    SEEME_TRY(gemm_f32(gemm_params(h->tn, 256, blkw(h, l, CA_K_W), 256, blkw(h, l, CA_K_B), h->kv2[l], 512, rows, 256, 256), s));
    SEEME_TRY(gemm_f32(gemm_params(h->tn, 256, blkw(h, l, CA_V_W), 256, blkw(h, l, CA_V_B), h->kv2[l] + 256, 512, rows, 256, 256), s));
  }
  return SEEME_OK;
}

// one block (mdiff_transformer.py:286-304) for table row `ti`: 9 tcgen05 linears + 5 row-wise kernels
static int den_block(seeme_denoiser* h, int l, int ti, const ActBuf& xin, const ActBuf& xout, int Nc, int R, cudaStream_t s) {
  const int nb = (R + 7) / 8, np = h->npass;
  ActBuf qkv; qkv.f = h->qkv; qkv.ld = 768;
  ActBuf t0; t0.f = h->t0; t0.ld = 256;
  ActBuf caq; caq.f = h->caq; caq.ld = 256;
  ActBuf y; y.f = h->y; y.ld = 256;
  // self-attention over {x, cond tokens, time token}, token 0 only (H1)
  SEEME_TRY(run_linear(h->Wqkv[l], xin, nullptr, R, ACT_NONE, nullptr, 0, qkv, np, s));
  {
    auto k = Nc == 1 ? den_sa_attn_kernel<1> : Nc == 2 ? den_sa_attn_kernel<2> : Nc == 3 ? den_sa_attn_kernel<3> : den_sa_attn_kernel<4>;
    SEEME_CUDA(launch_pdl(k, dim3(nb), dim3(256), 0, s, h->qkv, h->kvc[l], h->tkv[l] + (size_t)ti * 512, Nc, R, h->att.h, h->att.l));
  }
  SEEME_LAUNCH_CHECK();
  SEEME_TRY(run_linear(h->Wout[l], h->att, nullptr, R, ACT_NONE, xin.f, 256, t0, np, s));
  SEEME_CUDA(launch_pdl(den_ln_kernel, dim3(nb), dim3(256), 0, s, h->t0, blkw(h, l, SA_N1_W), blkw(h, l, SA_N1_B), h->x1.f, h->x1.h, h->x1.l, nullptr, nullptr,
                                   nullptr, nullptr, R));
  SEEME_LAUNCH_CHECK();
  SEEME_TRY(run_linear(h->Wl1[l], h->x1, nullptr, R, ACT_RELU, nullptr, 0, h->ff, np, s));
  SEEME_TRY(run_linear(h->Wl2[l], h->ff, nullptr, R, ACT_NONE, h->x1.f, 256, t0, np, s));
  // x2 = norm2(.) and the cross-attention's input norm chained in one kernel
  SEEME_CUDA(launch_pdl(den_ln_kernel, dim3(nb), dim3(256), 0, s, h->t0, blkw(h, l, SA_N2_W), blkw(h, l, SA_N2_B), h->x2.f, nullptr, nullptr, blkw(h, l, CA_N_W),
                                   blkw(h, l, CA_N_B), h->ln.h, h->ln.l, R));
  SEEME_LAUNCH_CHECK();
  // linear cross-attention to the cond tokens (H2) + FiLM
  SEEME_TRY(run_linear(h->Wcaq[l], h->ln, nullptr, R, ACT_NONE, nullptr, 0, caq, np, s));
  {
    auto k = Nc == 1 ? den_ca_kernel<1> : Nc == 2 ? den_ca_kernel<2> : Nc == 3 ? den_ca_kernel<3> : den_ca_kernel<4>;
    SEEME_CUDA(launch_pdl(k, dim3(nb), dim3(256), 0, s, h->caq, h->kv2[l], Nc, R, h->film_ca[l] + (size_t)ti * 512, blkw(h, l, CA_PN_W),
                          blkw(h, l, CA_PN_B), h->hb.h, h->hb.l));
  }
  SEEME_LAUNCH_CHECK();
  SEEME_TRY(run_linear(h->Wcaout[l], h->hb, nullptr, R, ACT_NONE, h->x2.f, 256, h->x1, np, s));    // x3 -> x1 (fp32 + bf16)
  // FFN + FiLM
  SEEME_TRY(run_linear(h->Wf1[l], h->x1, nullptr, R, ACT_GELU, nullptr, 0, h->g1, np, s));
  SEEME_TRY(run_linear(h->Wf2[l], h->g1, nullptr, R, ACT_NONE, nullptr, 0, y, np, s));
  SEEME_CUDA(launch_pdl(den_film_kernel, dim3(nb), dim3(256), 0, s, h->y, h->film_ff[l] + (size_t)ti * 512, blkw(h, l, FF_PN_W), blkw(h, l, FF_PN_B), h->hb.h,
                                     h->hb.l, R));
  SEEME_LAUNCH_CHECK();
  SEEME_TRY(run_linear(h->Wfout[l], h->hb, nullptr, R, ACT_NONE, h->x1.f, 256, xout, np, s));
  return SEEME_OK;
}

// the skip stack (cross_attention.py:67-83) on h->x for table row ti; result (pre final norm) in h->L[4]
static int den_stack(seeme_denoiser* h, int ti, int Nc, int R, cudaStream_t s) {
  const ActBuf* x = &h->x;
  for (int l = 0; l < 5; ++l) {
    if (l >= 3) {   // x = Linear(cat[x, skip]); skip = L[1] for l == 3, L[0] for l == 4
      SEEME_TRY(run_linear(h->Wskip[l - 3], *x, &h->L[l == 3 ? 1 : 0], R, ACT_NONE, nullptr, 0, h->x, h->npass, s));
      x = &h->x;
    }
    SEEME_TRY(den_block(h, l, ti, *x, h->L[l], Nc, R, s));
    x = &h->L[l];
  }
  return SEEME_OK;
}

static int check_rows(seeme_denoiser* h, int Nc, int R, const char* who) {
  SEEME_REQUIRE(Nc >= 1 && Nc <= SEEME_MAX_COND_TOKENS, SEEME_EINVAL, "%s: Nc=%d unsupported (1..%d)", who, Nc, SEEME_MAX_COND_TOKENS);
  SEEME_REQUIRE(R > 0, SEEME_EINVAL, "%s: empty batch", who);
  SEEME_REQUIRE(R <= h->max_rows, SEEME_ECAP, "%s: %d rows exceed capacity %d", who, R, h->max_rows);
  return SEEME_OK;
}

extern "C" int seeme_denoiser_forward(seeme_denoiser_t h, const float* sample, int timestep, const float* cond, int Nc,
                                      int R, float* out, void* stream) {
  SEEME_REQUIRE(h && sample && cond && out, SEEME_EINVAL, "seeme_denoiser_forward: null argument");
  SEEME_TRY(check_rows(h, Nc, R, "seeme_denoiser_forward"));
  cudaStream_t s = (cudaStream_t)stream;
  if (!(h->table_ts.size() == 1 && h->table_ts[0] == timestep)) SEEME_TRY(den_build_tables(h, &timestep, 1, nullptr, s));
  SEEME_CUDA(cudaMemcpyAsync(h->cond, cond, (size_t)Nc * R * 256 * 4, cudaMemcpyDeviceToDevice, s));
  if (h->persist && h->backend != SEEME_SAMPLER_GRAPH) return den_persist_run(h, 1, sample, Nc, R, R, 0, 1, out, s);
  SEEME_TRY(den_cond_precompute(h, Nc, R, s));
  SEEME_CUDA(launch_pdl(den_prep_kernel, dim3((R + 7) / 8), dim3(256), 0, s, sample, h->w[DN_PE], h->x.f, h->x.h, h->x.l, R, R));
  SEEME_LAUNCH_CHECK();
  SEEME_TRY(den_stack(h, 0, Nc, R, s));
  SEEME_TRY(layernorm256(h->L[4].f, nullptr, 0, h->w[DN_NORM_W], h->w[DN_NORM_B], out, R, s));
  return SEEME_OK;
}

extern "C" int seeme_denoiser_set_time_table(seeme_denoiser_t h, const int32_t* timesteps, int n, const float* sinusoid,
                                             void* stream) {
  SEEME_REQUIRE(h && timesteps, SEEME_EINVAL, "seeme_denoiser_set_time_table: null argument");
  return den_build_tables(h, timesteps, n, sinusoid, (cudaStream_t)stream);
}

static int sampler_enqueue(seeme_denoiser* h, int Nc, int B, int R, int cfg, int n_steps, cudaStream_t s) {
  SEEME_TRY(den_cond_precompute(h, Nc, R, s));
  for (int i = 0; i < n_steps; ++i) {
    SEEME_CUDA(launch_pdl(den_prep_kernel, dim3((R + 7) / 8), dim3(256), 0, s, h->lat, h->w[DN_PE], h->x.f, h->x.h, h->x.l, R, B));
    SEEME_LAUNCH_CHECK();
    SEEME_TRY(den_stack(h, i, Nc, R, s));
    SEEME_CUDA(launch_pdl(den_final_ddim_kernel, dim3((B + 7) / 8), dim3(256), 0, s, h->L[4].f, h->w[DN_NORM_W], h->w[DN_NORM_B], h->lat, B, cfg,
                                                      h->d_coef + 4 * i, h->d_gscale));
    SEEME_LAUNCH_CHECK();
  }
  return SEEME_OK;
}

extern "C" int seeme_sampler_run(seeme_denoiser_t h, const float* x_T, const float* cond, int Nc, int B,
                                 float guidance_scale, int n_steps, const int32_t* timesteps, const float* coef, float* z,
                                 void* stream) {
  SEEME_REQUIRE(h && x_T && cond && timesteps && coef && z, SEEME_EINVAL, "seeme_sampler_run: null argument");
  const int cfg = guidance_scale > 1.0f ? 1 : 0;     // do_classifier_free_guidance, mld.py:331
  const int R = cfg ? 2 * B : B;
  SEEME_TRY(check_rows(h, Nc, R, "seeme_sampler_run"));
  SEEME_REQUIRE(n_steps > 0 && n_steps <= MAX_STEPS, SEEME_EINVAL, "seeme_sampler_run: n_steps=%d out of range", n_steps);
  cudaStream_t s = (cudaStream_t)stream;
  bool same = (int)h->table_ts.size() == n_steps;
  for (int i = 0; same && i < n_steps; ++i) same = h->table_ts[i] == timesteps[i];
  if (!same) SEEME_TRY(den_build_tables(h, timesteps, n_steps, nullptr, s));
  // scheduler coefficients / guidance scale live in device memory so the captured graph is generic;
  // they are re-uploaded (with a sync: the sources are caller-owned host temporaries) only on change
  if (h->coef_host.size() != (size_t)n_steps * 4 || memcmp(h->coef_host.data(), coef, (size_t)n_steps * 16) != 0 ||
      h->gscale_host != guidance_scale) {
    SEEME_CUDA(cudaMemcpyAsync(h->d_coef, coef, (size_t)n_steps * 4 * 4, cudaMemcpyHostToDevice, s));
    SEEME_CUDA(cudaMemcpyAsync(h->d_gscale, &guidance_scale, 4, cudaMemcpyHostToDevice, s));
    SEEME_CUDA(cudaStreamSynchronize(s));
    h->coef_host.assign(coef, coef + (size_t)n_steps * 4);
    h->gscale_host = guidance_scale;
  }
  SEEME_CUDA(cudaMemcpyAsync(h->cond, cond, (size_t)Nc * R * 256 * 4, cudaMemcpyDeviceToDevice, s));
  if (h->persist && h->backend != SEEME_SAMPLER_GRAPH) return den_persist_run(h, 0, x_T, Nc, B, R, cfg, n_steps, z, s);
  SEEME_CUDA(cudaMemcpyAsync(h->lat, x_T, (size_t)B * 256 * 4, cudaMemcpyDeviceToDevice, s));
  if (h->use_graph) {
    if (!(h->gexec && h->g_Nc == Nc && h->g_B == B && h->g_cfg == cfg && h->g_steps == n_steps)) {
      if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
      cudaGraph_t graph = nullptr;
      const unsigned long long before = g_launch_count;
      SEEME_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
      int rc = sampler_enqueue(h, Nc, B, R, cfg, n_steps, h->cap_stream);
      cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
      h->g_kernels = (int)(g_launch_count - before);
      g_launch_count = before;   // captured, not launched
      if (rc != SEEME_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (e != cudaSuccess) { set_error("seeme_sampler_run: graph capture failed: %s", cudaGetErrorString(e)); return SEEME_ECUDA; }
      e = cudaGraphInstantiate(&h->gexec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) { h->gexec = nullptr; set_error("seeme_sampler_run: graph instantiate failed: %s", cudaGetErrorString(e)); return SEEME_ECUDA; }
      h->g_Nc = Nc; h->g_B = B; h->g_cfg = cfg; h->g_steps = n_steps;
    }
    {
      ProfScope prof(PROF_SAMPLER_GRAPH, s);
      SEEME_CUDA(cudaGraphLaunch(h->gexec, s));
    }
    count_launch(h->g_kernels);
  } else {
    SEEME_TRY(sampler_enqueue(h, Nc, B, R, cfg, n_steps, s));
  }
  SEEME_CUDA(cudaMemcpyAsync(z, h->lat, (size_t)B * 256 * 4, cudaMemcpyDeviceToDevice, s));
  return SEEME_OK;
}

extern "C" int seeme_denoiser_set_backend(seeme_denoiser_t h, int backend) {
  SEEME_REQUIRE(h, SEEME_EINVAL, "seeme_denoiser_set_backend: null handle");
  SEEME_REQUIRE(backend == SEEME_SAMPLER_PERSISTENT || backend == SEEME_SAMPLER_GRAPH || backend == SEEME_SAMPLER_TILE, SEEME_EINVAL,
                "seeme_denoiser_set_backend: unknown back-end %d", backend);
  SEEME_REQUIRE(backend == SEEME_SAMPLER_GRAPH || h->persist, SEEME_EINVAL,
                "seeme_denoiser_set_backend: the persistent sampler was disabled at create (SEEME_SAMPLER=graph)");
  h->backend = backend;
  return SEEME_OK;
}

extern "C" int seeme_denoiser_destroy(seeme_denoiser_t h) {
  if (!h) return SEEME_OK;
  if (h->gexec) cudaGraphExecDestroy(h->gexec);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  den_persist_destroy(h);
  h->arena.release();
  delete h;
  return SEEME_OK;
}
