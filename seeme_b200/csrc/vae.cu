// Motion VAE: MldVae.encode / MldVae.decode (mld/models/architectures/mld_vae.py:128-256) over the
// skip-connected DETR stacks of mld/models/operator/cross_attention.py (post-norm, 1 head, d=256,
// ff=128, GELU(erf), learned PE added once).
//
// Internal layout is sample-major [B, S, 256] (the reference is [S, B, 256]; per-token math is
// layout independent).  Algebra used: the decoder's cross-attention has ONE key
// (memory = z[1,B,256]) so its softmax is identically 1 and its output is
// out_proj(v_proj(z_b)) for every frame of sample b -- a per-sample vector computed up-front for
// all five layers; the q/k projections of that attention never influence the result.
#include "common.cuh"
#include "rowops.cuh"
#include "umma.cuh"
#include <stdlib.h>

namespace seeme {

#ifdef SEEME_EXPERIMENTAL   // retired fp32 CUDA-core attention kernels (SEEME_VAE_ATTN=0 / 1 for A/B measurements)
// Single-head attention over <= 64 tokens per sample.  One CTA per sample, K and V of the sample
// staged in shared memory (2 x S x 1 KB), one warp per query row, fp32 softmax.
// qkv [B*S, 768] = (q * 1/16 | k | v);  key j is valid iff j < n_prefix + lengths[b].
__global__ void __launch_bounds__(256) mha1_kernel(const float* __restrict__ qkv, const int* __restrict__ lengths,
                                                   int n_prefix, int S, __nv_bfloat16* __restrict__ oh,
                                                   __nv_bfloat16* __restrict__ ol) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;
  float* Vs = sm + (size_t)S * 256;
  const int b = blockIdx.x;
  const float* base = qkv + (size_t)b * S * 768;
  for (int i = threadIdx.x; i < S * 64; i += blockDim.x) {
    int row = i >> 6, c4 = i & 63;
    reinterpret_cast<float4*>(Ks)[row * 64 + c4] = reinterpret_cast<const float4*>(base + (size_t)row * 768 + 256)[c4];
    reinterpret_cast<float4*>(Vs)[row * 64 + c4] = reinterpret_cast<const float4*>(base + (size_t)row * 768 + 512)[c4];
  }
  __syncthreads();
  int nvalid = n_prefix + lengths[b];
  nvalid = nvalid < S ? nvalid : S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int qi = warp; qi < S; qi += (blockDim.x >> 5)) {
    const Row8 q = row_load(base + (size_t)qi * 768, lane);
    // lane j holds the score of key j and key j+32
    float s0 = -INFINITY, s1 = -INFINITY;
    for (int j = 0; j < nvalid; ++j) {
      const Row8 k = row_load(Ks + (size_t)j * 256, lane);
      const float d = row_dot(q, k);
      if (j < 32) { if (lane == j) s0 = d; } else { if (lane == j - 32) s1 = d; }
    }
    const float m = warp_max(fmaxf(s0, s1));
    const float e0 = (s0 == -INFINITY) ? 0.f : expf(s0 - m);
    const float e1 = (s1 == -INFINITY) ? 0.f : expf(s1 - m);
    const float inv = 1.0f / warp_sum(e0 + e1);
    Row8 acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
    for (int j = 0; j < nvalid; ++j) {
      const float p = __shfl_sync(0xffffffffu, j < 32 ? e0 : e1, j & 31) * inv;
      const Row8 v = row_load(Vs + (size_t)j * 256, lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(p, v.v[i], acc.v[i]);
    }
    row_store_split(oh + ((size_t)b * S + qi) * 256, ol + ((size_t)b * S + qi) * 256, lane, acc);
  }
}


// Register-tiled version of the same attention (S <= 64, d = 256): one CTA per sample, Q/K/V staged in shared memory
// with a 260-float row pitch, each thread owns a 4x4 tile of the 64x64 score matrix (rows ty + 16 i, keys tx + 16 j:
// consecutive rows per quarter-warp -> conflict-free 16-byte reads), row softmax by one warp per row, then a 4 x 16
// tile of P.V per thread.  ~10x fewer instructions than one warp-reduction per (query, key) pair.
constexpr int MH_P = 260;      // Q/K/V row pitch (floats)
constexpr int MH_PP = 65;      // P row pitch
constexpr int MH_SMEM = (3 * 64 * MH_P + 64 * MH_PP) * 4;
__global__ void __launch_bounds__(256) mha1_tiled_kernel(const float* __restrict__ qkv, const int* __restrict__ lengths,
                                                         int n_prefix, int S, __nv_bfloat16* __restrict__ oh,
                                                         __nv_bfloat16* __restrict__ ol) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;
  float* Ks = Qs + 64 * MH_P;
  float* Vs = Ks + 64 * MH_P;
  float* Ps = Vs + 64 * MH_P;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* base = qkv + (size_t)b * S * 768;
  for (int i = tid; i < 64 * 64 * 3; i += 256) {           // 64 rows x 192 float4 (q | k | v)
    const int row = i / 192, c4 = i % 192;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < S) v = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * 768) + c4);
    float* dst = (c4 < 64 ? Qs : c4 < 128 ? Ks : Vs) + row * MH_P + (c4 & 63) * 4;
    *reinterpret_cast<float4*>(dst) = v;
  }
  __syncthreads();
  int nvalid = n_prefix + lengths[b];
  nvalid = nvalid < S ? nvalid : S;
  const int tx = tid & 15, ty = tid >> 4;
  {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int c = 0; c < 256; c += 4) {
      float4 q[4], k[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) q[i] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * MH_P + c);
#pragma unroll
      for (int j = 0; j < 4; ++j) k[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * MH_P + c);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j] = fmaf(q[i].w, k[j].w, fmaf(q[i].z, k[j].z, fmaf(q[i].y, k[j].y, fmaf(q[i].x, k[j].x, acc[i][j]))));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Ps[(ty + 16 * i) * MH_PP + tx + 16 * j] = (tx + 16 * j) < nvalid ? acc[i][j] : -INFINITY;
  }
  __syncthreads();
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < 64; r += 8) {
      const float s0 = Ps[r * MH_PP + lane], s1 = Ps[r * MH_PP + lane + 32];
      const float m = warp_max(fmaxf(s0, s1));
      const float e0 = (s0 == -INFINITY) ? 0.f : expf(s0 - m);
      const float e1 = (s1 == -INFINITY) ? 0.f : expf(s1 - m);
      const float inv = 1.0f / warp_sum(e0 + e1);
      Ps[r * MH_PP + lane] = e0 * inv;
      Ps[r * MH_PP + lane + 32] = e1 * inv;
    }
  }
  __syncthreads();
  {
    float4 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kk = 0; kk < nvalid; ++kk) {
      float p[4];
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = Ps[(ty + 16 * i) * MH_PP + kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(Vs + kk * MH_P + tx * 4 + 64 * j);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j].x = fmaf(p[i], v[j].x, acc[i][j].x); acc[i][j].y = fmaf(p[i], v[j].y, acc[i][j].y);
          acc[i][j].z = fmaf(p[i], v[j].z, acc[i][j].z); acc[i][j].w = fmaf(p[i], v[j].w, acc[i][j].w);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = ty + 16 * i;
      if (row < S) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float f[4] = {acc[i][j].x, acc[i][j].y, acc[i][j].z, acc[i][j].w};
          __align__(8) __nv_bfloat16 h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            h[e] = __float2bfloat16_rn(f[e]);
            l[e] = __float2bfloat16_rn(f[e] - __bfloat162float(h[e]));
          }
          const size_t o = ((size_t)b * S + row) * 256 + tx * 4 + 64 * j;
          *reinterpret_cast<uint2*>(oh + o) = *reinterpret_cast<const uint2*>(h);
          *reinterpret_cast<uint2*>(ol + o) = *reinterpret_cast<const uint2*>(l);
        }
      }
    }
  }
}

#endif  // SEEME_EXPERIMENTAL

// Single-head attention over <= 64 tokens per sample on the tensor cores: one CTA per PAIR of samples, both contractions on tcgen05.
//     S = Q K^T : [128 x 256] x [256 x 128]   rows / keys = (sample 0 | sample 1) x 64 tokens; only the two diagonal 64 x 64
//                                             blocks are used (the off-diagonal half is the price of M = 128)
//     softmax   : thread = query row, 64 scores from tensor memory, key-padding mask, fp32
//     O = P V   : [128 x 128] x [128 x 256]   P block-diagonal (zeros for the other sample's keys)
// Operands are split fp16 (hi + lo, three products hi.hi + lo.hi + hi.lo, fp32 accumulation: ~2^-22 relative), written by
// the threads straight into the 128-byte-swizzled K-major tile images the MMA reads: Q and K per 64-column chunk into a
// double buffer (staging of chunk c+1 overlaps the MMAs of chunk c), V TRANSPOSED ([dim][key]: 8 consecutive keys of one
// dim are one 16-byte chunk) into the same 128 KB once S is complete, P next to it.
constexpr int MU_CH = 128 * 128;                       // [128 rows x 64 fp16] operand chunk
constexpr int MU_R0 = 8 * MU_CH;                       // Q/K double buffer, later V^T hi | lo (2 x [2 K-blocks x 256 x 128 B])
constexpr int MU_R1 = 4 * MU_CH;                       // P: hi (2 K-blocks) | lo (2 K-blocks)
constexpr int MU_SMEM = MU_R0 + MU_R1 + 1024;
constexpr int MU_THREADS = 256;


__device__ __forceinline__ uint32_t mu_sw128(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ void mu_split8(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(MU_THREADS, 1) mha1_umma_kernel(const float* __restrict__ qkv, const int* __restrict__ lengths,
                                                                    int n_prefix, int S, int B, __nv_bfloat16* __restrict__ oh,
                                                                    __nv_bfloat16* __restrict__ ol) {
  extern __shared__ __align__(1024) uint8_t mu_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mu_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* r0 = smem;
  uint8_t* r1 = smem + MU_R0;
  __shared__ __align__(8) uint64_t qk_done[2], o_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float sm_red[4][128];         // softmax: row max / row sum of the two key halves
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * 2;

  if (tid == 0) {
    mbar_init(&qk_done[0], 1);
    mbar_init(&qk_done[1], 1);
    mbar_init(&o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t TS = tmem_base, TO = tmem_base + 128u;

  // ---- S = Q K^T over four 64-column chunks ------------------------------------------------------------------------
  // item = (matrix, row, 8-float chunk): 2 x 128 x 8 per K-chunk, 8 per thread; the loads of chunk c + 1 are in flight while
  // chunk c is converted
  auto load_qk = [&](int c, float4 (&x)[16]) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = it * MU_THREADS + tid;
      const int mat = idx >> 10, row = (idx >> 3) & 127, ch = idx & 7;
      const int sl = row >> 6, sq = row & 63;
      if (sq < S && b0 + sl < B) {
        const float4* src = reinterpret_cast<const float4*>(qkv + ((size_t)(b0 + sl) * S + sq) * 768 + mat * 256 + c * 64 + ch * 8);
        x[2 * it] = __ldg(src);
        x[2 * it + 1] = __ldg(src + 1);
      } else {
        x[2 * it] = x[2 * it + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  float4 cur[16];
  load_qk(0, cur);
  for (int c = 0; c < 4; ++c) {
    uint8_t* buf = r0 + (c & 1) * (4 * MU_CH);                  // [Qh | Ql | Kh | Kl]
    float4 nxt[16];
    if (c < 3) load_qk(c + 1, nxt);
    if (c >= 2) mbar_wait(&qk_done[c & 1], 0);                   // the MMAs of chunk c - 2 have read this buffer
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = it * MU_THREADS + tid;
      const int mat = idx >> 10, row = (idx >> 3) & 127, ch = idx & 7;
      const float x[8] = {cur[2 * it].x, cur[2 * it].y, cur[2 * it].z, cur[2 * it].w,
                          cur[2 * it + 1].x, cur[2 * it + 1].y, cur[2 * it + 1].z, cur[2 * it + 1].w};
      uint4 hi, lo;
      mu_split8(x, hi, lo);
      uint8_t* t = buf + mat * (2 * MU_CH) + mu_sw128(row, ch);
      *reinterpret_cast<uint4*>(t) = hi;
      *reinterpret_cast<uint4*>(t + MU_CH) = lo;
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (umma_elect_one()) {
        constexpr uint32_t idesc = umma_idesc_f16(128);
        const uint64_t qh = umma_desc_k128(smem_u32(buf)), ql = umma_desc_add(qh, MU_CH >> 4);
        const uint64_t kh = umma_desc_add(qh, (2 * MU_CH) >> 4), kl = umma_desc_add(qh, (3 * MU_CH) >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_bf16(TS, umma_desc_add(qh, ks * 2), umma_desc_add(kh, ks * 2), idesc, (c | ks) != 0);
          umma_bf16(TS, umma_desc_add(ql, ks * 2), umma_desc_add(kh, ks * 2), idesc, 1);
          umma_bf16(TS, umma_desc_add(qh, ks * 2), umma_desc_add(kl, ks * 2), idesc, 1);
        }
        umma_commit(&qk_done[c & 1]);
      }
      __syncwarp();
    }
    if (c < 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
    }
  }
  // V^T: task = (group of 8 keys, half of the dims); a lane owns dims lane + 32 i of the half: consecutive rows of the operand
  // tile -> conflict-free 16-byte stores, 128-byte coalesced loads.  4 tasks per warp; the 32 loads of a task are issued together,
  // those of the first task before the wait for the S MMAs (they are in flight under the softmax)
  const int t_begin = warp * 4, t_end = t_begin + 4;
  auto load_task = [&](int t, float (&x)[4][8]) {
    const int kg = t >> 1, half = t & 1;
    const int sl = kg >> 3, s0 = (kg & 7) * 8;
    const bool sample_ok = b0 + sl < B;
    const float* vb = qkv + ((size_t)(b0 + sl) * S + s0) * 768 + 512 + half * 128 + lane;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) x[i][jj] = (sample_ok && s0 + jj < S) ? __ldg(vb + (size_t)jj * 768 + i * 32) : 0.f;
  };
  auto store_task = [&](int t, float (&x)[4][8]) {
    const int kg = t >> 1, half = t & 1, sl = kg >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 hi, lo;
      mu_split8(x[i], hi, lo);
      const int d = half * 128 + i * 32 + lane;
      uint8_t* dst = r0 + sl * (2 * MU_CH) + mu_sw128(d, kg & 7);              // K-block sl: [256 dims x 64 keys]
      *reinterpret_cast<uint4*>(dst) = hi;
      *reinterpret_cast<uint4*>(dst + 4 * MU_CH) = lo;
    }
  };
  float vx[4][8];
  load_task(t_begin, vx);
  mbar_wait(&qk_done[0], 1);
  mbar_wait(&qk_done[1], 1);          // every MMA of S has completed: the score columns are final, the Q/K buffers are free
  tc_fence_after();

  // ---- softmax: two threads per query row (warps w and w + 4 share TMEM lane quarter w: 32 keys each) --------------------
  {
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane, sl = row >> 6;
    int nvalid = 0;
    if (b0 + sl < B) {
      nvalid = n_prefix + __ldg(lengths + b0 + sl);
      nvalid = nvalid < S ? nvalid : S;
    }
    nvalid -= half * 32;                                       // valid keys among this thread's 32
    uint32_t raw[32];
    tmem_ld32(TS + ((uint32_t)(q * 32) << 16) + (uint32_t)(sl * 64 + half * 32), raw);
    tmem_ld_wait();
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) m = (j < nvalid) ? fmaxf(m, __uint_as_float(raw[j])) : m;
    sm_red[half][row] = m;
    __syncthreads();
    m = fmaxf(sm_red[0][row], sm_red[1][row]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float e = (j < nvalid) ? exp2f((__uint_as_float(raw[j]) - m) * 1.4426950408889634f) : 0.f;
      raw[j] = __float_as_uint(e);
      sum += e;
    }
    sm_red[2 + half][row] = sum;
    __syncthreads();
    const float inv = 1.0f / (sm_red[2][row] + sm_red[3][row]);
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(raw[c4 * 8 + i]) * inv;
      uint4 hi, lo;
      mu_split8(x, hi, lo);
      const uint32_t off = mu_sw128(row, half * 4 + c4);
      *reinterpret_cast<uint4*>(r1 + sl * MU_CH + off) = hi;                       // own sample's keys: K-block sl
      *reinterpret_cast<uint4*>(r1 + 2 * MU_CH + sl * MU_CH + off) = lo;
      *reinterpret_cast<uint4*>(r1 + (1 - sl) * MU_CH + off) = z4;                 // the other sample's keys
      *reinterpret_cast<uint4*>(r1 + 2 * MU_CH + (1 - sl) * MU_CH + off) = z4;
    }
  }
  for (int t = t_begin; t < t_end; ++t) {
    float nx[4][8];
    if (t + 1 < t_end) load_task(t + 1, nx);
    store_task(t, vx);
    if (t + 1 < t_end) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) vx[i][jj] = nx[i][jj];
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();

  // ---- O = P V ------------------------------------------------------------------------------------------------------
  if (warp == 0) {
    tc_fence_after();
    if (umma_elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(256);
      const uint64_t ph = umma_desc_k128(smem_u32(r1)), pl = umma_desc_add(ph, (2 * MU_CH) >> 4);
      const uint64_t vh = umma_desc_k128(smem_u32(r0)), vl = umma_desc_add(vh, (4 * MU_CH) >> 4);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t ao = (uint32_t)(kb * (MU_CH >> 4) + ks * 2), bo = (uint32_t)(kb * ((2 * MU_CH) >> 4) + ks * 2);
          umma_bf16(TO, umma_desc_add(ph, ao), umma_desc_add(vh, bo), idesc, (kb | ks) != 0);
          umma_bf16(TO, umma_desc_add(pl, ao), umma_desc_add(vh, bo), idesc, 1);
          umma_bf16(TO, umma_desc_add(ph, ao), umma_desc_add(vl, bo), idesc, 1);
        }
      umma_commit(&o_full);
    }
    __syncwarp();
  }
  mbar_wait(&o_full, 0);
  tc_fence_after();
  {
    // thread = (row, column half): 128 accumulator columns -> bf16 (hi, lo) into a [128 x 512 B] staging tile each (the operand
    // buffers are free now), 16-byte chunks XOR-swizzled by the row so that both these row-per-lane stores and the
    // row-per-warp reads below are conflict-free; then every warp copies 16 rows with 512-byte coalesced global stores
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t ta = TO + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 128);
    uint8_t* th = r0 + row * 512;
    uint32_t rawb[2][32];
    tmem_ld32(ta, rawb[0]);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      tmem_ld_wait();
      if (g < 3) tmem_ld32(ta + (g + 1) * 32, rawb[(g + 1) & 1]);
      const uint32_t* raw = rawb[g & 1];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float f0 = __uint_as_float(raw[jj * 8 + 2 * i]), f1 = __uint_as_float(raw[jj * 8 + 2 * i + 1]);
          const __nv_bfloat162 hh = __floats2bfloat162_rn(f0, f1);
          const float2 hf = __bfloat1622float2(hh);
          const __nv_bfloat162 ll = __floats2bfloat162_rn(f0 - hf.x, f1 - hf.y);
          h[i] = *reinterpret_cast<const uint32_t*>(&hh);
          l[i] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        const int ch = (half * 16 + g * 4 + jj) ^ (row & 7);
        *reinterpret_cast<uint4*>(th + ch * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(th + 128 * 512 + ch * 16) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
    tc_fence_before();
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int r = warp * 16 + i, sl = r >> 6, sq = r & 63;
      if (sq < S && b0 + sl < B) {
        const size_t o = ((size_t)(b0 + sl) * S + sq) * 256 + lane * 8;
        const uint8_t* src = r0 + r * 512 + ((lane ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(oh + o) = *reinterpret_cast<const uint4*>(src);
        *reinterpret_cast<uint4*>(ol + o) = *reinterpret_cast<const uint4*>(src + 128 * 512);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// xseq[b, s] = (s < 2 ? global_motion_token[s] : emb[b, s-2]) + pe[s]     (mld_vae.py:147-164)
__global__ void vae_enc_assemble_kernel(const float* __restrict__ emb, const float* __restrict__ token,
                                        const float* __restrict__ pe, float* __restrict__ x,
                                        __nv_bfloat16* __restrict__ xh, __nv_bfloat16* __restrict__ xl, int B, int T) {
  const int S = T + 2;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * S) return;
  const int b = row / S, s = row % S;
  Row8 v = (s < 2) ? row_load(token + (size_t)s * 256, lane) : row_load(emb + ((size_t)b * T + (s - 2)) * 256, lane);
  const Row8 p = row_load(pe + (size_t)s * 256, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) v.v[i] += p.v[i];
  row_store(x + (size_t)row * 256, lane, v);
  row_store_split(xh + (size_t)row * 256, xl + (size_t)row * 256, lane, v);
}

// queries[b, t] = 0 + pe[t]                                                (mld_vae.py:198,230)
__global__ void vae_dec_queries_kernel(const float* __restrict__ pe, float* __restrict__ x, __nv_bfloat16* __restrict__ xh,
                                       __nv_bfloat16* __restrict__ xl, int B, int T) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * T) return;
  const Row8 v = row_load(pe + (size_t)(row % T) * 256, lane);
  row_store(x + (size_t)row * 256, lane, v);
  row_store_split(xh + (size_t)row * 256, xl + (size_t)row * 256, lane, v);
}

// final encoder LayerNorm on tokens 0/1 of each sample, then mu/logvar -> std = exp(logvar)^0.5,
// z = mu + std * eps                                                        (mld_vae.py:181-192)
__global__ void vae_sample_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta,
                                  const float* __restrict__ eps, float* __restrict__ z, float* __restrict__ mu_out,
                                  float* __restrict__ std_out, int B, int S) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const Row8 mu = row_layernorm(row_load(x + ((size_t)b * S + 0) * 256, lane), g, bta, lane);
  const Row8 lv = row_layernorm(row_load(x + ((size_t)b * S + 1) * 256, lane), g, bta, lane);
  const Row8 e = row_load(eps + (size_t)b * 256, lane);
  Row8 sd, zz;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sd.v[i] = sqrtf(expf(lv.v[i]));
    zz.v[i] = mu.v[i] + sd.v[i] * e.v[i];
  }
  row_store(z + (size_t)b * 256, lane, zz);
  if (mu_out) row_store(mu_out + (size_t)b * 256, lane, mu);
  if (std_out) row_store(std_out + (size_t)b * 256, lane, sd);
}

// y = LN(x (+ r[row / r_group_rows])) -> fp32 and split bf16
__global__ void vae_ln_kernel(const float* __restrict__ x, const float* __restrict__ r, int r_group_rows,
                              const float* __restrict__ g, const float* __restrict__ b, float* __restrict__ y,
                              __nv_bfloat16* __restrict__ yh, __nv_bfloat16* __restrict__ yl, int rows) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  Row8 v = row_load(x + (size_t)row * 256, lane);
  if (r) {
    const Row8 a = row_load(r + (size_t)(row / r_group_rows) * 256, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) v.v[i] += a.v[i];
  }
  v = row_layernorm(v, g, b, lane);
  if (y) row_store(y + (size_t)row * 256, lane, v);
  if (yh) row_store_split(yh + (size_t)row * 256, yl + (size_t)row * 256, lane, v);
}

}  // namespace seeme

using namespace seeme;

namespace {
enum { V_TOKEN = 0, V_PE_ENC = 1, V_PE_DEC = 2, V_SKEL_W = 3, V_SKEL_B = 4, V_FINAL_W = 5, V_FINAL_B = 6,
       V_ENC = 7, V_ENC_BLK = 13, V_DEC = 73, V_DEC_BLK = 79 };
// per-stack header: norm.w, norm.b, lb0.w, lb0.b, lb1.w, lb1.b
// encoder block (12): in_w in_b out_w out_b l1w l1b l2w l2b n1w n1b n2w n2b
// decoder block (18): sa(in_w in_b out_w out_b) ca(in_w in_b out_w out_b) l1w l1b l2w l2b n1 n2 n3 (w,b)
}  // namespace

struct seeme_vae {
  int device = 0, nfeats = 0, max_batch = 0, max_frames = 0;
  Arena arena;
  float* w[SEEME_VAE_NUM_TENSORS];
  float *emb, *qkv, *t0, *ca[5], *vtmp;
  int npass = 3;
  int attn = 2;                 // 2 = tcgen05 (mha1_umma_kernel, default), 1 = register-tiled fp32, 0 = one warp per query row
  // per stack (0 encoder, 1 decoder) and block: packed (hi, lo) weights of the tcgen05 linears
  PackedLinear Wqkv[2][5], Wout[2][5], Wl1[2][5], Wl2[2][5], Wskip[2][2];
  ActBuf x0, x, L[5], att, x1, x2, ff;
};

static size_t vae_tensor_elems(int i, int nfeats) {
  if (i == V_TOKEN) return 2 * 256;
  if (i == V_PE_ENC || i == V_PE_DEC) return 500 * 256;
  if (i == V_SKEL_W) return (size_t)256 * nfeats;
  if (i == V_SKEL_B) return 256;
  if (i == V_FINAL_W) return (size_t)nfeats * 256;
  if (i == V_FINAL_B) return nfeats;
  auto header = [](int k) -> size_t { return k < 2 ? 256 : (k % 2 == 0 ? 256 * 512 : 256); };
  if (i < V_ENC_BLK) return header(i - V_ENC);
  if (i < V_DEC) {
    static const size_t e[12] = {768 * 256, 768, 256 * 256, 256, 128 * 256, 128, 256 * 128, 256, 256, 256, 256, 256};
    return e[(i - V_ENC_BLK) % 12];
  }
  if (i < V_DEC_BLK) return header(i - V_DEC);
  static const size_t d[18] = {768 * 256, 768, 256 * 256, 256, 768 * 256, 768, 256 * 256, 256,
                               128 * 256, 128, 256 * 128, 256, 256, 256, 256, 256, 256, 256};
  return d[(i - V_DEC_BLK) % 18];
}

extern "C" int seeme_vae_create(seeme_vae_t* out, const float* const* w, int n_w, int nfeats, int max_batch,
                                int max_frames) {
  SEEME_REQUIRE(out && w, SEEME_EINVAL, "seeme_vae_create: null argument");
  SEEME_REQUIRE(n_w == SEEME_VAE_NUM_TENSORS, SEEME_EINVAL, "seeme_vae_create: expected %d tensors, got %d",
                SEEME_VAE_NUM_TENSORS, n_w);
  SEEME_REQUIRE(nfeats > 0 && nfeats <= 512 && max_batch > 0 && max_frames > 0 && max_frames + 2 <= 64, SEEME_EINVAL,
                "seeme_vae_create: unsupported sizes (nfeats=%d, max_batch=%d, max_frames=%d; frames+2 must be <= 64)",
                nfeats, max_batch, max_frames);
  seeme_vae* h = new seeme_vae();
  SEEME_CUDA(cudaGetDevice(&h->device));
  h->nfeats = nfeats; h->max_batch = max_batch; h->max_frames = max_frames;
  size_t wbytes = 0;
  for (int i = 0; i < n_w; ++i) wbytes += pad256(vae_tensor_elems(i, nfeats) * 4);
  const size_t rows = (size_t)max_batch * (max_frames + 2);
  // fp32: emb, t0, x0, x, L[5], x1, x2 (11 x 256) + qkv (768); bf16 pairs: x0, x, L[5], att, x1, x2 (10 x 256) + ff (128)
  size_t ws = 11 * pad256(rows * 256 * 4) + pad256(rows * 768 * 4) + 20 * pad256(rows * 256 * 2) + 2 * pad256(rows * 128 * 2) +
              6 * pad256((size_t)max_batch * 256 * 4);
  const size_t pbytes = 2 * 2 * 2 * (size_t)(5 * (768 * 256 + 256 * 256 + 2 * 128 * 256) + 2 * 256 * 512) + 64 * 1024;
  int rc = h->arena.init(wbytes + ws + pbytes + 65536);
  if (rc) { delete h; return rc; }
  for (int i = 0; i < n_w; ++i) {
    size_t n = vae_tensor_elems(i, nfeats);
    h->w[i] = h->arena.take<float>(n);
    if (!h->w[i] || !w[i]) { set_error("seeme_vae_create: tensor %d null or arena exhausted", i); h->arena.release(); delete h; return SEEME_EINVAL; }
    cudaError_t e = cudaMemcpy(h->w[i], w[i], n * 4, cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) { set_error("seeme_vae_create: copy of tensor %d failed: %s", i, cudaGetErrorString(e)); h->arena.release(); delete h; return SEEME_ECUDA; }
  }
  // fold the 1/sqrt(256) attention scale into the q rows of every self-attention in_proj
  auto scale_q = [&](int wi, int bi) {
    scale_kernel_launch(h->w[wi], 256 * 256, 0.0625f);
    scale_kernel_launch(h->w[bi], 256, 0.0625f);
  };
  for (int l = 0; l < 5; ++l) {
    scale_q(V_ENC_BLK + 12 * l, V_ENC_BLK + 12 * l + 1);
    scale_q(V_DEC_BLK + 18 * l, V_DEC_BLK + 18 * l + 1);
  }
  SEEME_CUDA(cudaDeviceSynchronize());
  h->emb = h->arena.take<float>(rows * 256);
  h->t0 = h->arena.take<float>(rows * 256);
  h->qkv = h->arena.take<float>(rows * 768);
  auto mk = [&](ActBuf& a, int ld, bool f32) {
    a.ld = ld;
    a.f = f32 ? h->arena.take<float>(rows * ld) : nullptr;
    a.h = h->arena.take<__nv_bfloat16>(rows * ld);
    a.l = h->arena.take<__nv_bfloat16>(rows * ld);
  };
  mk(h->x0, 256, true);
  mk(h->x, 256, true);
  for (int l = 0; l < 5; ++l) mk(h->L[l], 256, true);
  mk(h->att, 256, false);
  mk(h->x1, 256, true);
  mk(h->x2, 256, true);
  mk(h->ff, 128, false);
  for (int l = 0; l < 5; ++l) h->ca[l] = h->arena.take<float>((size_t)max_batch * 256);
  h->vtmp = h->arena.take<float>((size_t)max_batch * 256);
  if (!h->vtmp) { set_error("seeme_vae_create: arena exhausted (workspace)"); h->arena.release(); delete h; return SEEME_ENOMEM; }
  const char* pe = getenv("SEEME_VAE_PRECISION");
  h->npass = (pe && atoi(pe) == 1) ? 1 : 3;
  rc = SEEME_OK;
  for (int st = 0; st < 2 && !rc; ++st) {
    const int hdr = st ? V_DEC : V_ENC, blk = st ? V_DEC_BLK : V_ENC_BLK, stride = st ? 18 : 12, fo = st ? 8 : 4;
    for (int l = 0; l < 5 && !rc; ++l) {
      float* const* wb = h->w + blk + stride * l;
      rc = pack_linear(h->arena, h->Wqkv[st][l], wb[0], 256, 768, 256, wb[1]);
      if (!rc) rc = pack_linear(h->arena, h->Wout[st][l], wb[2], 256, 256, 256, wb[3]);
      if (!rc) rc = pack_linear(h->arena, h->Wl1[st][l], wb[fo + 0], 256, 128, 256, wb[fo + 1]);
      if (!rc) rc = pack_linear(h->arena, h->Wl2[st][l], wb[fo + 2], 128, 256, 128, wb[fo + 3]);
    }
    for (int i = 0; i < 2 && !rc; ++i) rc = pack_linear(h->arena, h->Wskip[st][i], h->w[hdr + 2 + 2 * i], 512, 256, 512, h->w[hdr + 3 + 2 * i]);
  }
  if (!rc && cudaDeviceSynchronize() != cudaSuccess) { set_error("seeme_vae_create: weight packing failed"); rc = SEEME_ECUDA; }
  if (rc) { h->arena.release(); delete h; return rc; }
#ifdef SEEME_EXPERIMENTAL
  SEEME_CUDA(cudaFuncSetAttribute(mha1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * 256 * 4));
  SEEME_CUDA(cudaFuncSetAttribute(mha1_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MH_SMEM));
#endif
  SEEME_CUDA(cudaFuncSetAttribute(mha1_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MU_SMEM));
  {
    const char* e = seeme_exp_env("SEEME_VAE_ATTN");      // A/B measurements: 0 / 1 select the fp32 CUDA-core kernels
    h->attn = e ? atoi(e) : 2;
  }
  *out = h;
  return SEEME_OK;
}

// self-attention + (optional single-key cross-attention vector) + FFN, post-norm; 4 tcgen05 linears,
// the fused attention kernel and 2-3 LayerNorm kernels.  st = 0 encoder stack, 1 decoder stack.
static int vae_layer(seeme_vae* h, int st, int l, const ActBuf& xin, const ActBuf& xout, const int* lengths, int n_prefix, int B,
                     int S, cudaStream_t s) {
  const int rows = B * S, nb = (rows + 7) / 8, np = h->npass;
  const bool dec = st == 1;
  float* const* wb = h->w + (dec ? V_DEC_BLK + 18 * l : V_ENC_BLK + 12 * l);
  float* const* f = wb + (dec ? 8 : 4);     // l1w l1b l2w l2b n1w n1b n2w n2b (n3w n3b)
  ActBuf qkv; qkv.f = h->qkv; qkv.ld = 768;
  ActBuf t0; t0.f = h->t0; t0.ld = 256;
  SEEME_TRY(run_linear(h->Wqkv[st][l], xin, nullptr, rows, ACT_NONE, nullptr, 0, qkv, np, s));
  {
    ProfScope prof(PROF_VAE_ATTN, s);
#ifdef SEEME_EXPERIMENTAL
    if (h->attn == 1) mha1_tiled_kernel<<<B, 256, MH_SMEM, s>>>(h->qkv, lengths, n_prefix, S, h->att.h, h->att.l);
    else if (h->attn == 0) mha1_kernel<<<B, 256, (size_t)2 * S * 256 * 4, s>>>(h->qkv, lengths, n_prefix, S, h->att.h, h->att.l);
    else
#endif
      mha1_umma_kernel<<<(B + 1) / 2, MU_THREADS, MU_SMEM, s>>>(h->qkv, lengths, n_prefix, S, B, h->att.h, h->att.l);
  }
  SEEME_LAUNCH_CHECK();
  SEEME_TRY(run_linear(h->Wout[st][l], h->att, nullptr, rows, ACT_NONE, xin.f, 256, t0, np, s));
  vae_ln_kernel<<<nb, 256, 0, s>>>(h->t0, nullptr, 0, f[4], f[5], h->x1.f, h->x1.h, h->x1.l, rows);
  SEEME_LAUNCH_CHECK();
  const ActBuf* xa = &h->x1;
  if (dec) {   // tgt = norm2(tgt + ca[b]): the one-key cross-attention adds a per-sample vector (H6)
    vae_ln_kernel<<<nb, 256, 0, s>>>(h->x1.f, h->ca[l], S, f[6], f[7], h->x2.f, h->x2.h, h->x2.l, rows);
    SEEME_LAUNCH_CHECK();
    xa = &h->x2;
  }
  SEEME_TRY(run_linear(h->Wl1[st][l], *xa, nullptr, rows, ACT_GELU, nullptr, 0, h->ff, np, s));
  SEEME_TRY(run_linear(h->Wl2[st][l], h->ff, nullptr, rows, ACT_NONE, xa->f, 256, t0, np, s));
  vae_ln_kernel<<<nb, 256, 0, s>>>(h->t0, nullptr, 0, dec ? f[8] : f[6], dec ? f[9] : f[7], xout.f, xout.h, xout.l, rows);
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

// the 2-1-2 skip topology of SkipTransformerEncoder/Decoder (cross_attention.py:42-65, 108-147); input h->x0
static int vae_stack(seeme_vae* h, int st, const int* lengths, int n_prefix, int B, int S, cudaStream_t s) {
  const ActBuf* x = &h->x0;
  for (int l = 0; l < 5; ++l) {
    if (l >= 3) {   // x = Linear(cat[x, skip]);  skip = L[1] for l == 3, L[0] for l == 4
      SEEME_TRY(run_linear(h->Wskip[st][l - 3], *x, &h->L[l == 3 ? 1 : 0], B * S, ACT_NONE, nullptr, 0, h->x, h->npass, s));
      x = &h->x;
    }
    SEEME_TRY(vae_layer(h, st, l, *x, h->L[l], lengths, n_prefix, B, S, s));
    x = &h->L[l];
  }
  return SEEME_OK;
}

extern "C" int seeme_vae_encode(seeme_vae_t h, const float* features, const int32_t* lengths, const float* eps, int B,
                                int T, float* z, float* mu, float* std, void* stream) {
  SEEME_REQUIRE(h && features && lengths && eps && z, SEEME_EINVAL, "seeme_vae_encode: null argument");
  SEEME_REQUIRE(B > 0 && T > 0, SEEME_EINVAL, "seeme_vae_encode: empty input (B=%d, T=%d)", B, T);
  SEEME_REQUIRE(B <= h->max_batch && T <= h->max_frames, SEEME_ECAP, "seeme_vae_encode: B=%d T=%d exceeds capacity (%d, %d)",
                B, T, h->max_batch, h->max_frames);
  cudaStream_t s = (cudaStream_t)stream;
  const int S = T + 2, rows = B * S;
  // skel_embedding: K = nfeats (75) is not a multiple of the tensor-core K granule -> fp32 CUDA-core GEMM
  SEEME_TRY(gemm_f32(gemm_params(features, h->nfeats, h->w[V_SKEL_W], h->nfeats, h->w[V_SKEL_B], h->emb, 256, B * T, 256,
                                 h->nfeats), s));
  vae_enc_assemble_kernel<<<(rows + 7) / 8, 256, 0, s>>>(h->emb, h->w[V_TOKEN], h->w[V_PE_ENC], h->x0.f, h->x0.h, h->x0.l, B, T);
  SEEME_LAUNCH_CHECK();
  SEEME_TRY(vae_stack(h, 0, lengths, 2, B, S, s));
  vae_sample_kernel<<<(B + 7) / 8, 256, 0, s>>>(h->L[4].f, h->w[V_ENC], h->w[V_ENC + 1], eps, z, mu, std, B, S);
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

extern "C" int seeme_vae_decode(seeme_vae_t h, const float* z, const int32_t* lengths, int B, int T, float* feats,
                                void* stream) {
  SEEME_REQUIRE(h && z && lengths && feats, SEEME_EINVAL, "seeme_vae_decode: null argument");
  SEEME_REQUIRE(B > 0 && T > 0, SEEME_EINVAL, "seeme_vae_decode: empty input (B=%d, T=%d)", B, T);
  SEEME_REQUIRE(B <= h->max_batch && T <= h->max_frames, SEEME_ECAP, "seeme_vae_decode: B=%d T=%d exceeds capacity (%d, %d)",
                B, T, h->max_batch, h->max_frames);
  cudaStream_t s = (cudaStream_t)stream;
  const int rows = B * T;
  // single-key cross-attention vectors for all five layers: ca_l[b] = out_proj(v_proj(z_b))   (B rows: fp32 GEMMs)
  for (int l = 0; l < 5; ++l) {
    float* const* wb = h->w + V_DEC_BLK + 18 * l;
    SEEME_TRY(gemm_f32(gemm_params(z, 256, wb[4] + 512 * 256, 256, wb[5] + 512, h->vtmp, 256, B, 256, 256), s));
    SEEME_TRY(gemm_f32(gemm_params(h->vtmp, 256, wb[6], 256, wb[7], h->ca[l], 256, B, 256, 256), s));
  }
  vae_dec_queries_kernel<<<(rows + 7) / 8, 256, 0, s>>>(h->w[V_PE_DEC], h->x0.f, h->x0.h, h->x0.l, B, T);
  SEEME_LAUNCH_CHECK();
  SEEME_TRY(vae_stack(h, 1, lengths, 0, B, T, s));
  SEEME_TRY(layernorm256(h->L[4].f, nullptr, 0, h->w[V_DEC], h->w[V_DEC + 1], h->t0, rows, s));
  // final_layer: N = nfeats (75) -> fp32 CUDA-core GEMM writing the [B,T,nfeats] output directly
  SEEME_TRY(gemm_f32(gemm_params(h->t0, 256, h->w[V_FINAL_W], 256, h->w[V_FINAL_B], feats, h->nfeats, rows, h->nfeats, 256), s));
  return SEEME_OK;
}

extern "C" int seeme_vae_destroy(seeme_vae_t h) {
  if (!h) return SEEME_OK;
  h->arena.release();
  delete h;
  return SEEME_OK;
}
