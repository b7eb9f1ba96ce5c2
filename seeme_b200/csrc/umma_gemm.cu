// Generic tcgen05 "linear" for sm_100a:  D[M,N] = epilogue([A1|A2][M,K] . W[N,K]^T)
//
//   warp 0      TMA producer: 128x64 bf16 activation boxes and BNx64 weight boxes (SWIZZLE_128B) into a
//               STAGES-deep shared-memory ring, completion on "full" mbarriers
//   warp 1      allocates TMEM, then one elected lane issues tcgen05.mma (M=128, N=BN, K=16 per
//               instruction; fp32 accumulator in TMEM); tcgen05.commit releases ring slots ("empty"
//               mbarriers) and finally signals the epilogue
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns per warp and step, thread = output row),
//               bias / per-sample bias / rank-3 xyz fold / activation / residual, then fp32 and/or bf16
//               (hi, lo) stores and the fused per-sample column max (scene-encoder pooling)
//
// Precision: NPASS = 1 multiplies bf16(A) x bf16(W); NPASS = 3 adds the two first-order correction
// terms lo(A).hi(W) + hi(A).lo(W) of the split x = hi + lo (|lo| <= 2^-9 |x|), which restores ~16 mantissa
// bits at 3x the MMA count -- used where the reference's fp32 results must be tracked closely.
#include "umma.cuh"
#include <cudaTypedefs.h>

namespace seeme {

struct UmmaEpi {
  int M, N, nkb1, nkb;           // nkb1 k-blocks come from A1, the remaining from A2
  const float* bias; int bias_group_rows;
  const float* pfold; const float* xyz;
  int act;
  const float* R; int ldr;
  float* Y; int ldy;
  __nv_bfloat16 *Yh, *Yl, *Zh, *Zl; int ldb;
  unsigned* colmax; int colmax_group_rows;
};

template <int BN, int NPASS>
struct UmmaCfg {
  static constexpr int A_BYTES = 128 * 64 * 2;
  static constexpr int W_BYTES = BN * 64 * 2;
  static constexpr int STAGE_BYTES = (A_BYTES + W_BYTES) * (NPASS == 3 ? 2 : 1);
};

template <int BN, int NPASS, int STAGES>
__global__ void __launch_bounds__(192) umma_linear_kernel(const __grid_constant__ CUtensorMap tmA1h,
                                                          const __grid_constant__ CUtensorMap tmA1l,
                                                          const __grid_constant__ CUtensorMap tmA2h,
                                                          const __grid_constant__ CUtensorMap tmA2l,
                                                          const __grid_constant__ CUtensorMap tmWh,
                                                          const __grid_constant__ CUtensorMap tmWl, const UmmaEpi e) {
  using Cfg = UmmaCfg<BN, NPASS>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned colmax_s[2][BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1h);
    tma_prefetch_desc(&tmWh);
    if (NPASS == 3) { tma_prefetch_desc(&tmA1l); tma_prefetch_desc(&tmWl); }
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, BN < 32 ? 32 : BN);
  if (e.colmax) for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) (&colmax_s[0][0])[i] = 0u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < e.nkb; ++kb) {
        const int st = kb % STAGES;
        mbar_wait(&empty_bar[st], ((kb / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[st], Cfg::STAGE_BYTES);
        uint8_t* sa = smem + (size_t)st * Cfg::STAGE_BYTES;
        uint8_t* sw = sa + Cfg::A_BYTES * (NPASS == 3 ? 2 : 1);
        const bool first = kb < e.nkb1;
        const int kc = (first ? kb : kb - e.nkb1) * 64;
        tma_load_2d(sa, first ? &tmA1h : &tmA2h, &full_bar[st], kc, m0);
        if (NPASS == 3) tma_load_2d(sa + Cfg::A_BYTES, first ? &tmA1l : &tmA2l, &full_bar[st], kc, m0);
        tma_load_2d(sw, &tmWh, &full_bar[st], kb * 64, n0);
        if (NPASS == 3) tma_load_2d(sw + Cfg::W_BYTES, &tmWl, &full_bar[st], kb * 64, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BN);
      for (int kb = 0; kb < e.nkb; ++kb) {
        const int st = kb % STAGES;
        mbar_wait(&full_bar[st], (kb / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)st * Cfg::STAGE_BYTES);
        const uint32_t sw = sa + Cfg::A_BYTES * (NPASS == 3 ? 2 : 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {     // 4 x (K = 16) inside the 128-byte swizzle atom: +32 bytes each
          const uint64_t da = umma_desc_k128(sa + k * 32), dw = umma_desc_k128(sw + k * 32);
          umma_bf16(tmem_base, da, dw, idesc, (kb | k) != 0);
          if (NPASS == 3) {
            const uint64_t dal = umma_desc_k128(sa + Cfg::A_BYTES + k * 32), dwl = umma_desc_k128(sw + Cfg::W_BYTES + k * 32);
            umma_bf16(tmem_base, dal, dw, idesc, 1);
            umma_bf16(tmem_base, da, dwl, idesc, 1);
          }
        }
        umma_commit(&empty_bar[st]);      // slot reusable once these MMAs have read it
      }
      umma_commit(&tmem_full_bar);        // accumulator complete
    }
  } else {
    // ---- epilogue: warp w may only touch TMEM lanes [32 (w % 4), 32 (w % 4) + 32) ----------------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool valid = m < e.M;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const float* brow = nullptr;
    if (e.bias) brow = e.bias + (e.bias_group_rows ? (size_t)((valid ? m : e.M - 1) / e.bias_group_rows) * e.N : 0);
    float px = 0.f, py = 0.f, pz = 0.f;
    if (e.pfold && valid) { px = e.xyz[(size_t)m * 3]; py = e.xyz[(size_t)m * 3 + 1]; pz = e.xyz[(size_t)m * 3 + 2]; }
    int g0 = 0, gmine = 0;
    if (e.colmax) { g0 = m0 / e.colmax_group_rows; gmine = (valid ? m : m0) / e.colmax_group_rows - g0; }
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t raw[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, raw);
      tmem_ld_wait();
      const int n = n0 + c0;
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(raw[i]);
      if (brow) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(brow + n + i));
          f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
        }
      }
      if (e.pfold) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float4 p = __ldg(reinterpret_cast<const float4*>(e.pfold) + n + i);
          f[i] += p.x * px + p.y * py + p.z * pz;
        }
      }
      if (e.act != ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = apply_act(f[i], e.act);
      }
      if (e.R && valid) {
        const float4* rp = reinterpret_cast<const float4*>(e.R + (size_t)m * e.ldr + n);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 r = __ldg(rp + i);
          f[4 * i] += r.x; f[4 * i + 1] += r.y; f[4 * i + 2] += r.z; f[4 * i + 3] += r.w;
        }
      }
      if (valid) {
        if (e.Y) {
          float4* yp = reinterpret_cast<float4*>(e.Y + (size_t)m * e.ldy + n);
#pragma unroll
          for (int i = 0; i < 8; ++i) yp[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        }
        if (e.Yh) {
          __align__(16) __nv_bfloat16 hb[32], lb[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) { hb[i] = __float2bfloat16_rn(f[i]); lb[i] = __float2bfloat16_rn(f[i] - __bfloat162float(hb[i])); }
          uint4* hp = reinterpret_cast<uint4*>(e.Yh + (size_t)m * e.ldb + n);
#pragma unroll
          for (int i = 0; i < 4; ++i) hp[i] = reinterpret_cast<const uint4*>(hb)[i];
          if (e.Yl) {
            uint4* lp = reinterpret_cast<uint4*>(e.Yl + (size_t)m * e.ldb + n);
#pragma unroll
            for (int i = 0; i < 4; ++i) lp[i] = reinterpret_cast<const uint4*>(lb)[i];
          }
        }
        if (e.Zh) {
          __align__(16) __nv_bfloat16 hb[32], lb[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float r = fmaxf(f[i], 0.f);
            hb[i] = __float2bfloat16_rn(r);
            lb[i] = __float2bfloat16_rn(r - __bfloat162float(hb[i]));
          }
          uint4* hp = reinterpret_cast<uint4*>(e.Zh + (size_t)m * e.ldb + n);
#pragma unroll
          for (int i = 0; i < 4; ++i) hp[i] = reinterpret_cast<const uint4*>(hb)[i];
          if (e.Zl) {
            uint4* lp = reinterpret_cast<uint4*>(e.Zl + (size_t)m * e.ldb + n);
#pragma unroll
            for (int i = 0; i < 4; ++i) lp[i] = reinterpret_cast<const uint4*>(lb)[i];
          }
        }
      }
      if (e.colmax) {
        // per-sample max over the rows of this tile; a tile spans at most two samples (group >= 128 rows)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const unsigned o = f2ord(f[i]);
          const unsigned a = __reduce_max_sync(0xffffffffu, (valid && gmine == 0) ? o : 0u);
          const unsigned b = __reduce_max_sync(0xffffffffu, (valid && gmine == 1) ? o : 0u);
          if (lane == 0) {
            if (a) atomicMax(&colmax_s[0][c0 + i], a);
            if (b) atomicMax(&colmax_s[1][c0 + i], b);
          }
        }
      }
    }
    if (e.colmax) {
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps
      const int G = (e.M + e.colmax_group_rows - 1) / e.colmax_group_rows;
      for (int i = threadIdx.x - 64; i < 2 * BN; i += 128) {
        const int which = i / BN, c = i % BN;
        const unsigned v = colmax_s[which][c];
        if (v && g0 + which < G) atomicMax(e.colmax + (size_t)(g0 + which) * e.N + n0 + c, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
  }
}

// ---- host side ------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static int get_encoder() {
  if (g_encode) return SEEME_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SEEME_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  SEEME_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, SEEME_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return SEEME_OK;
}

// bf16 [rows, cols] with row pitch ld (elements), box = box_rows x 64, SWIZZLE_128B, OOB -> zeros
static int make_map(CUtensorMap* map, const __nv_bfloat16* ptr, int rows, int cols, int ld, int box_rows) {
  SEEME_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld % 8) == 0, SEEME_EINVAL,
                "umma_linear: operand base must be 16-byte aligned and its pitch a multiple of 8 elements (ld=%d)", ld);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(ptr), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SEEME_REQUIRE(r == CUDA_SUCCESS, SEEME_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%d cols=%d ld=%d)", (int)r, rows, cols, ld);
  return SEEME_OK;
}

template <int BN, int NPASS, int STAGES>
static int launch(const UmmaLinear& g, const CUtensorMap* maps, const UmmaEpi& e, cudaStream_t s) {
  constexpr int smem = UmmaCfg<BN, NPASS>::STAGE_BYTES * STAGES + 1024;
  static bool configured = false;
  if (!configured) {
    SEEME_CUDA(cudaFuncSetAttribute(umma_linear_kernel<BN, NPASS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(g.N / BN, (g.M + 127) / 128);
  ProfScope prof(g.prof_id - 1, s);
  umma_linear_kernel<BN, NPASS, STAGES><<<grid, 192, smem, s>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], e);
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

int umma_linear(const UmmaLinear& g, int npass, cudaStream_t s) {
  SEEME_TRY(get_encoder());
  const int K = g.K1 + g.K2;
  SEEME_REQUIRE(g.M > 0 && g.N > 0 && g.K1 > 0 && g.K1 % 64 == 0 && g.K2 % 64 == 0, SEEME_EINVAL,
                "umma_linear: unsupported shape M=%d N=%d K1=%d K2=%d (K multiples of 64)", g.M, g.N, g.K1, g.K2);
  SEEME_REQUIRE(npass == 1 || npass == 3, SEEME_EINVAL, "umma_linear: npass must be 1 or 3");
  SEEME_REQUIRE(npass == 1 || (g.A1.lo && g.W.lo && (g.K2 == 0 || g.A2.lo)), SEEME_EINVAL, "umma_linear: split-bf16 needs lo operands");
  SEEME_REQUIRE(!g.colmax || g.colmax_group_rows >= 128, SEEME_EINVAL, "umma_linear: colmax groups must have >= 128 rows");
  // few rows (the latency-bound sampler / per-sample GEMMs): narrow tiles spread one GEMM over more SMs;
  // many rows: the widest tile that divides N
  const int BN = (g.M <= 2048 && !g.colmax) ? 64 : ((g.N % 256 == 0) ? 256 : 128);
  SEEME_REQUIRE(g.N % BN == 0, SEEME_EINVAL, "umma_linear: N=%d must be a multiple of %d", g.N, BN);
  CUtensorMap maps[6];
  memset(maps, 0, sizeof(maps));
  SEEME_TRY(make_map(&maps[0], g.A1.hi, g.M, g.K1, g.A1.ld, 128));
  if (npass == 3) SEEME_TRY(make_map(&maps[1], g.A1.lo, g.M, g.K1, g.A1.ld, 128)); else maps[1] = maps[0];
  if (g.K2) {
    SEEME_TRY(make_map(&maps[2], g.A2.hi, g.M, g.K2, g.A2.ld, 128));
    if (npass == 3) SEEME_TRY(make_map(&maps[3], g.A2.lo, g.M, g.K2, g.A2.ld, 128)); else maps[3] = maps[2];
  } else { maps[2] = maps[0]; maps[3] = maps[1]; }
  SEEME_TRY(make_map(&maps[4], g.W.hi, g.N, K, g.W.ld, BN));
  if (npass == 3) SEEME_TRY(make_map(&maps[5], g.W.lo, g.N, K, g.W.ld, BN)); else maps[5] = maps[4];
  UmmaEpi e;
  e.M = g.M; e.N = g.N; e.nkb1 = g.K1 / 64; e.nkb = K / 64;
  e.bias = g.bias; e.bias_group_rows = g.bias_group_rows; e.pfold = g.pfold; e.xyz = g.xyz; e.act = g.act;
  e.R = g.R; e.ldr = g.ldr; e.Y = g.Y; e.ldy = g.ldy; e.Yh = g.Yh; e.Yl = g.Yl; e.Zh = g.Zh; e.Zl = g.Zl; e.ldb = g.ldb;
  e.colmax = g.colmax; e.colmax_group_rows = g.colmax_group_rows;
  if (BN == 256) return npass == 1 ? launch<256, 1, 2>(g, maps, e, s) : launch<256, 3, 2>(g, maps, e, s);
  if (BN == 64) return npass == 1 ? launch<64, 1, 4>(g, maps, e, s) : launch<64, 3, 4>(g, maps, e, s);
  return npass == 1 ? launch<128, 1, 4>(g, maps, e, s) : launch<128, 3, 2>(g, maps, e, s);
}

__global__ void to_bf16_split_kernel(const float* __restrict__ x, int ldx, int rows, int cols, __nv_bfloat16* __restrict__ hi,
                                     __nv_bfloat16* __restrict__ lo, int ld_out, int relu) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  float v = x[(size_t)r * ldx + c];
  if (relu) v = fmaxf(v, 0.f);
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const size_t o = (size_t)r * ld_out + c;      // columns [cols, ld_out) of the destination are left untouched
  hi[o] = h;
  if (lo) lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
}

int to_bf16_split(const float* x, int ldx, int rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld_out, int relu,
                  cudaStream_t s) {
  const size_t n = (size_t)rows * cols;
  if (!n) return SEEME_OK;
  to_bf16_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, ldx, rows, cols, hi, lo, ld_out, relu);
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

int pack_linear(Arena& arena, PackedLinear& out, const float* W, int ldw, int N, int K, const float* bias) {
  out.hi = arena.take<__nv_bfloat16>((size_t)N * K);
  out.lo = arena.take<__nv_bfloat16>((size_t)N * K);
  SEEME_REQUIRE(out.hi && out.lo, SEEME_ENOMEM, "pack_linear: arena exhausted (N=%d K=%d)", N, K);
  out.N = N; out.K = K; out.bias = bias;
  return to_bf16_split(W, ldw, N, K, out.hi, out.lo, K, 0, 0);
}

int run_linear(const PackedLinear& W, const ActBuf& A, const ActBuf* A2, int M, int act, const float* R, int ldr,
               const ActBuf& out, int npass, cudaStream_t s, int bias_group_rows, const float* bias_override) {
  UmmaLinear g;
  g.M = M; g.N = W.N;
  g.A1 = {A.h, A.l, A.ld};
  if (A2) {
    g.A2 = {A2->h, A2->l, A2->ld};
    g.K1 = W.K / 2; g.K2 = W.K / 2;
  } else {
    g.K1 = W.K;
  }
  g.W = {W.hi, W.lo, W.K};
  g.bias = bias_override ? bias_override : W.bias;
  g.bias_group_rows = bias_group_rows;
  g.act = act;
  g.R = R; g.ldr = ldr;
  g.Y = out.f; g.ldy = out.ld;
  g.Yh = out.h; g.Yl = out.l; g.ldb = out.ld;
  return umma_linear(g, npass, s);
}

}  // namespace seeme

// ---- test entry point: Y = act(A W^T + bias) (+R) with fp32 inputs converted on the fly -----------------
extern "C" int seeme_test_umma_linear(const float* A, const float* W, const float* bias, const float* R, float* Y, int M, int N,
                                      int K, int act, int npass, unsigned* colmax, int colmax_group_rows, void* stream) {
  using namespace seeme;
  SEEME_REQUIRE(A && W && Y, SEEME_EINVAL, "seeme_test_umma_linear: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  __nv_bfloat16 *ah, *al, *wh, *wl;
  SEEME_CUDA(cudaMalloc(&ah, (size_t)M * K * 2));
  SEEME_CUDA(cudaMalloc(&al, (size_t)M * K * 2));
  SEEME_CUDA(cudaMalloc(&wh, (size_t)N * K * 2));
  SEEME_CUDA(cudaMalloc(&wl, (size_t)N * K * 2));
  int rc = to_bf16_split(A, K, M, K, ah, al, K, 0, s);
  if (!rc) rc = to_bf16_split(W, K, N, K, wh, wl, K, 0, s);
  if (!rc) {
    UmmaLinear g;
    g.A1 = {ah, al, K}; g.W = {wh, wl, K};
    g.M = M; g.N = N; g.K1 = K; g.K2 = 0;
    if (K >= 128 && K % 128 == 0) {    // exercise the split-K-source path: second half of K from "A2"
      g.K1 = K / 2; g.K2 = K / 2;
      g.A2 = {ah + K / 2, al + K / 2, K};
    }
    g.bias = bias; g.act = act; g.R = R; g.ldr = N; g.Y = Y; g.ldy = N;
    g.colmax = colmax; g.colmax_group_rows = colmax_group_rows;
    rc = umma_linear(g, npass, s);
  }
  cudaStreamSynchronize(s);
  cudaFree(ah); cudaFree(al); cudaFree(wh); cudaFree(wl);
  return rc;
}
