// Generic tcgen05 "linear" for sm_100a:  D[M,N] = epilogue([A1|A2][M,K] . W[N,K]^T)
//
//   warp 0      TMA producer: 128x64 bf16 activation boxes and BNx64 weight boxes (SWIZZLE_128B) into a
//               STAGES-deep shared-memory ring, completion on "full" mbarriers
//   warp 1      allocates TMEM, then one elected lane issues tcgen05.mma (M=128, N=BN, K=16 per
//               instruction; fp32 accumulator in TMEM); tcgen05.commit releases ring slots ("empty"
//               mbarriers) and finally signals the epilogue
//   warps 2..5  epilogue, 64 output columns at a time: tcgen05.ld (thread = output row), bias /
//               per-sample bias / rank-3 xyz fold / activation / fp32 residual (TMA-prefetched into
//               shared memory while the main loop runs), then the results are written as 128-byte-
//               swizzled tiles into a double-buffered shared-memory staging area (conflict-free 16-byte
//               stores) and leave the SM as TMA tensor stores (fp32 and/or bf16 hi/lo, plus bf16 of
//               relu(.)); the fused per-sample column max (scene-encoder pooling) is a warp redux on
//               order-preserving integers followed by one atomicMax per (tile, column).
//   Thread-per-row global loads/stores would touch 32 different 128-byte lines per instruction; the
//   ncu capture of that first version (profiles/r1_umma_v1_*.md) showed the epilogue's LSU wavefronts
//   bounding the kernel at 8-14 % tensor-pipe activity, hence the TMA staging.
//
// Precision: NPASS = 1 multiplies bf16(A) x bf16(W); NPASS = 3 adds the two first-order correction
// terms lo(A).hi(W) + hi(A).lo(W) of the split x = hi + lo (|lo| <= 2^-9 |x|), which restores ~16 mantissa
// bits at 3x the MMA count -- used where the reference's fp32 results must be tracked closely.
#include "umma.cuh"
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace seeme {

struct UmmaMaps {
  CUtensorMap a1h, a1l, a2h, a2l, wh, wl;   // operands
  CUtensorMap yh, yl, zh, zl, yf, rf;       // bf16 outputs, fp32 output, fp32 residual
};

struct UmmaEpi {
  int M, N, nkb1, nkb;           // nkb1 k-blocks come from A1, the remaining from A2
  const float* bias; int bias_group_rows;
  const float* pfold; const float* xyz;
  int act;
  int has_r, has_yf, has_yh, has_yl, has_zh, has_zl;
  int act_post;                 // activation applied after the residual add
  unsigned* colmax; int colmax_group_rows;
  int fp16;
};

constexpr int TILE_BYTES = 128 * 128;            // one 128-row x 128-byte swizzled staging tile
constexpr int STORE_BUF_BYTES = 4 * TILE_BYTES;  // slots: Yh, Yl, Zh | Yf(cols 0-31), Zl | Yf(cols 32-63)
constexpr int RES_BUF_BYTES = 2 * TILE_BYTES;    // residual: two 32-column fp32 tiles per 64-column group

template <int BN, int NPASS>
struct UmmaCfg {
  static constexpr int A_BYTES = 128 * 64 * 2;
  static constexpr int W_BYTES = BN * 64 * 2;
  static constexpr int STAGE_BYTES = (A_BYTES + W_BYTES) * (NPASS == 3 ? 2 : 1);
  static constexpr int NG = BN / 64;             // 64-column store groups
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// (a, b) -> packed bf16x2 hi = rn(a), rn(b) and lo = rn(a - hi_a), rn(b - hi_b): two cvt.rn.bf16x2.f32
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);          // .x = a (low 16 bits), .y = b
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// fp16 flavour: hi = packed half2 of (a, b); there is no lo part
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// byte offset of logical 16-byte chunk j of row r inside a 128-byte-swizzled [128 x 128 B] tile
__device__ __forceinline__ uint32_t sw128(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

// STAGES: depth of the operand ring; NBUF: store-staging (and residual) buffers; MINB: CTAs per SM the
// register allocation must allow (small-footprint configurations co-reside so that one CTA's epilogue
// overlaps another's main loop)
// activation of 32 accumulator values; the (uniform) selector is tested once, not per element (a per-element switch
// compiles to 64 indirect branches per thread and tile)
__device__ __forceinline__ void epi_act32(float (&f)[32], int act) {
  if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
  } else if (act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
  } else if (act == ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = silu(f[i]);
  }
}

template <int BN, int NPASS, int STAGES, int NBUF, int MINB>
__global__ void __launch_bounds__(192, MINB) umma_linear_kernel(const __grid_constant__ UmmaMaps tm, const UmmaEpi e) {
  using Cfg = UmmaCfg<BN, NPASS>;
  constexpr int RING_BYTES = Cfg::STAGE_BYTES * STAGES;
  constexpr int STORE_BYTES = NBUF * STORE_BUF_BYTES;
  constexpr int MAIN_BYTES = RING_BYTES > STORE_BYTES ? RING_BYTES : STORE_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.  Layout: [ring, reused as store staging after the last MMA | residual staging]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* res_stage = smem + MAIN_BYTES;
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar, res_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned colmax_s[2][BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a1h);
    tma_prefetch_desc(&tm.wh);
    if (NPASS == 3) { tma_prefetch_desc(&tm.a1l); tma_prefetch_desc(&tm.wl); }
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tmem_full_bar, 1);
    mbar_init(&res_bar[0], 1);
    mbar_init(&res_bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, BN < 32 ? 32 : BN);
  if (e.colmax) for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) (&colmax_s[0][0])[i] = 0u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_prologue();     // everything above overlaps the previous kernel of the chain; its outputs are read only below

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
        tma_load_2d(sa, first ? &tm.a1h : &tm.a2h, &full_bar[st], kc, m0);
        if (NPASS == 3) tma_load_2d(sa + Cfg::A_BYTES, first ? &tm.a1l : &tm.a2l, &full_bar[st], kc, m0);
        tma_load_2d(sw, &tm.wh, &full_bar[st], kb * 64, n0);
        if (NPASS == 3) tma_load_2d(sw + Cfg::W_BYTES, &tm.wl, &full_bar[st], kb * 64, n0);
      }
    }
  } else if (warp == 1) {
    // the whole warp runs the (uniform) loop; one elected lane issues the MMAs and commits -- this lets ptxas keep the
    // descriptors in uniform registers and issue the tcgen05.mma instructions back to back
    const uint32_t idesc = e.fp16 ? umma_idesc_f16(BN) : umma_idesc_bf16(BN);
    const uint64_t d0 = umma_desc_k128(smem_u32(smem));
    for (int kb = 0; kb < e.nkb; ++kb) {
      const int st = kb % STAGES;
      mbar_wait(&full_bar[st], (kb / STAGES) & 1);
      tc_fence_after();
      if (umma_elect_one()) {
        const uint64_t da0 = umma_desc_add(d0, (uint32_t)(st * (Cfg::STAGE_BYTES >> 4)));
        const uint64_t dw0 = umma_desc_add(da0, (uint32_t)((Cfg::A_BYTES * (NPASS == 3 ? 2 : 1)) >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) {     // 4 x (K = 16) inside the 128-byte swizzle atom: +32 bytes each
          const uint64_t da = umma_desc_add(da0, k * 2), dw = umma_desc_add(dw0, k * 2);
          umma_bf16(tmem_base, da, dw, idesc, (kb | k) != 0);
          if (NPASS == 3) {
            const uint64_t dal = umma_desc_add(da, Cfg::A_BYTES >> 4), dwl = umma_desc_add(dw, Cfg::W_BYTES >> 4);
            umma_bf16(tmem_base, dal, dw, idesc, 1);
            umma_bf16(tmem_base, da, dwl, idesc, 1);
          }
        }
        umma_commit(&empty_bar[st]);      // slot reusable once these MMAs have read it
        if (kb == e.nkb - 1) umma_commit(&tmem_full_bar);   // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue: warp w may only touch TMEM lanes [32 (w % 4), 32 (w % 4) + 32) ----------------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool valid = m < e.M;
    const bool elected = (warp == 2 && lane == 0);
    // residual tiles of the first store groups travel while the main loop runs
    if (e.has_r && elected) {
      for (int g = 0; g < NBUF && g < Cfg::NG; ++g) {
        mbar_arrive_expect_tx(&res_bar[g], RES_BUF_BYTES);
        tma_load_2d(res_stage + g * RES_BUF_BYTES, &tm.rf, &res_bar[g], n0 + g * 64, m0);
        tma_load_2d(res_stage + g * RES_BUF_BYTES + TILE_BYTES, &tm.rf, &res_bar[g], n0 + g * 64 + 32, m0);
      }
    }
    const float* brow = nullptr;
    if (e.bias) brow = e.bias + (e.bias_group_rows ? (size_t)((valid ? m : e.M - 1) / e.bias_group_rows) * e.N : 0);
    float px = 0.f, py = 0.f, pz = 0.f;
    if (e.pfold && valid) { px = e.xyz[(size_t)m * 3]; py = e.xyz[(size_t)m * 3 + 1]; pz = e.xyz[(size_t)m * 3 + 2]; }
    int g0 = 0, gmine = 0;
    if (e.colmax) { g0 = m0 / e.colmax_group_rows; gmine = (valid ? m : m0) / e.colmax_group_rows - g0; }
    mbar_wait(&tmem_full_bar, 0);       // all MMAs done: the accumulator is complete and the ring is free
    tc_fence_after();
#pragma unroll 1
    for (int g = 0; g < Cfg::NG; ++g) {
      const int buf = g % NBUF;
      uint8_t* sbuf = smem + buf * STORE_BUF_BYTES;
      if (g >= NBUF) {                  // the stores of group g-NBUF must have drained this staging buffer
        if (elected) tma_store_wait_read<NBUF - 1>();
        epi_bar_sync();
      }
      if (e.has_r) mbar_wait(&res_bar[buf], (g / NBUF) & 1);
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int c0 = g * 64 + half * 32;
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
        if (e.act != ACT_NONE && !e.act_post) epi_act32(f, e.act);
        if (e.has_r) {
          const uint8_t* rt = res_stage + buf * RES_BUF_BYTES + half * TILE_BYTES;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r = *reinterpret_cast<const float4*>(rt + sw128(row, j));
            f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
          }
        }
        if (e.act != ACT_NONE && e.act_post) epi_act32(f, e.act);
        if (e.has_yf) {                 // fp32 tile of 32 columns: slot 2 (half 0) / slot 3 (half 1)
          uint8_t* t = sbuf + (2 + half) * TILE_BYTES;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(t + sw128(row, j)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
        if (e.has_yh) {
          uint32_t hb[16], lb[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (e.fp16) { hb[i] = pack_f16x2(f[2 * i], f[2 * i + 1]); lb[i] = 0u; }
            else split_bf16x2(f[2 * i], f[2 * i + 1], hb[i], lb[i]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(sbuf + sw128(row, half * 4 + j)) = make_uint4(hb[4 * j], hb[4 * j + 1], hb[4 * j + 2], hb[4 * j + 3]);
          if (e.has_yl) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(sbuf + TILE_BYTES + sw128(row, half * 4 + j)) =
                  make_uint4(lb[4 * j], lb[4 * j + 1], lb[4 * j + 2], lb[4 * j + 3]);
          }
        }
        if (e.has_zh) {
          uint32_t hb[16], lb[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (e.fp16) { hb[i] = pack_f16x2(fmaxf(f[2 * i], 0.f), fmaxf(f[2 * i + 1], 0.f)); lb[i] = 0u; }
            else split_bf16x2(fmaxf(f[2 * i], 0.f), fmaxf(f[2 * i + 1], 0.f), hb[i], lb[i]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(sbuf + 2 * TILE_BYTES + sw128(row, half * 4 + j)) =
                make_uint4(hb[4 * j], hb[4 * j + 1], hb[4 * j + 2], hb[4 * j + 3]);
          if (e.has_zl) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(sbuf + 3 * TILE_BYTES + sw128(row, half * 4 + j)) =
                  make_uint4(lb[4 * j], lb[4 * j + 1], lb[4 * j + 2], lb[4 * j + 3]);
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
      fence_proxy_async();              // staging writes (generic proxy) -> visible to the TMA (async proxy)
      epi_bar_sync();
      if (elected) {
        const int c = n0 + g * 64;
        if (e.has_yh) tma_store_2d(&tm.yh, sbuf, c, m0);
        if (e.has_yl) tma_store_2d(&tm.yl, sbuf + TILE_BYTES, c, m0);
        if (e.has_zh) tma_store_2d(&tm.zh, sbuf + 2 * TILE_BYTES, c, m0);
        if (e.has_zl) tma_store_2d(&tm.zl, sbuf + 3 * TILE_BYTES, c, m0);
        if (e.has_yf) {
          tma_store_2d(&tm.yf, sbuf + 2 * TILE_BYTES, c, m0);
          tma_store_2d(&tm.yf, sbuf + 3 * TILE_BYTES, c + 32, m0);
        }
        tma_store_commit();
        if (e.has_r && g + NBUF < Cfg::NG) {   // everyone is past the reads of this residual buffer
          mbar_arrive_expect_tx(&res_bar[buf], RES_BUF_BYTES);
          tma_load_2d(res_stage + buf * RES_BUF_BYTES, &tm.rf, &res_bar[buf], n0 + (g + NBUF) * 64, m0);
          tma_load_2d(res_stage + buf * RES_BUF_BYTES + TILE_BYTES, &tm.rf, &res_bar[buf], n0 + (g + NBUF) * 64 + 32, m0);
        }
      }
    }
    if (elected) tma_store_wait_read<0>();   // shared memory must outlive the bulk stores reading it
    if (e.colmax) {
      epi_bar_sync();
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


// ---- persistent variant for many-row GEMMs (VAE stacks, image backbone) ------------------------------------------------
// One CTA per SM loops over [128 x 128] output tiles (n fastest: the CTAs that run at the same time share their A row tiles
// in L2): a 2-stage operand ring, TWO 128-column accumulators in tensor memory so that the MMAs of tile i+1 run under the
// epilogue of tile i, and an epilogue that keeps the TMA-staged stores / residual loads of the one-tile kernel above.
// The one-tile-per-CTA kernel pays barrier init + TMEM allocation + a cold ring per tile and serialises load, MMA and
// epilogue inside a CTA (2-3 co-resident CTAs hide part of it); here that cost is paid once per SM.
// Split-bf16 (x3), bias per column, activation before the residual, fp32 and / or bf16 (hi, lo) outputs.
constexpr int PL_STAGES = 2;
constexpr int PL_STAGE_BYTES = UmmaCfg<128, 3>::STAGE_BYTES;                 // 64 KB
constexpr int PL_SMEM = PL_STAGES * PL_STAGE_BYTES + STORE_BUF_BYTES + RES_BUF_BYTES + 1024;

__global__ void __launch_bounds__(192, 1) umma_plinear_kernel(const __grid_constant__ UmmaMaps tm, const UmmaEpi e, int n_tiles_n, int n_tiles) {
  using Cfg = UmmaCfg<128, 3>;
  extern __shared__ __align__(1024) uint8_t pl_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pl_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sbuf = smem + PL_STAGES * PL_STAGE_BYTES;
  uint8_t* res_stage = sbuf + STORE_BUF_BYTES;
  __shared__ __align__(8) uint64_t full_bar[PL_STAGES], empty_bar[PL_STAGES], acc_full[2], acc_empty[2], res_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a1h); tma_prefetch_desc(&tm.a1l); tma_prefetch_desc(&tm.wh); tma_prefetch_desc(&tm.wl);
    for (int i = 0; i < PL_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    mbar_init(&res_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_prologue();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (t / n_tiles_n) * 128, n0 = (t % n_tiles_n) * 128;
        for (int kb = 0; kb < e.nkb; ++kb, ++it) {
          const uint32_t st = it % PL_STAGES;
          mbar_wait(&empty_bar[st], ((it / PL_STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&full_bar[st], PL_STAGE_BYTES);
          uint8_t* sa = smem + (size_t)st * PL_STAGE_BYTES;
          uint8_t* sw = sa + 2 * Cfg::A_BYTES;
          const bool first = kb < e.nkb1;
          const int kc = (first ? kb : kb - e.nkb1) * 64;
          tma_load_2d(sa, first ? &tm.a1h : &tm.a2h, &full_bar[st], kc, m0);
          tma_load_2d(sa + Cfg::A_BYTES, first ? &tm.a1l : &tm.a2l, &full_bar[st], kc, m0);
          tma_load_2d(sw, &tm.wh, &full_bar[st], kb * 64, n0);
          tma_load_2d(sw + Cfg::W_BYTES, &tm.wl, &full_bar[st], kb * 64, n0);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128);
    const uint64_t d0 = umma_desc_k128(smem_u32(smem));
    uint32_t it = 0, ti = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
      const uint32_t buf = ti & 1u;
      mbar_wait(&acc_empty[buf], ((ti >> 1) & 1u) ^ 1u);       // the epilogue has drained this accumulator (tile ti - 2)
      tc_fence_after();
      const uint32_t d_t = tmem_base + buf * 128u;
      for (int kb = 0; kb < e.nkb; ++kb, ++it) {
        const uint32_t st = it % PL_STAGES;
        mbar_wait(&full_bar[st], (it / PL_STAGES) & 1u);
        tc_fence_after();
        if (umma_elect_one()) {
          const uint64_t da0 = umma_desc_add(d0, st * (PL_STAGE_BYTES >> 4));
          const uint64_t dw0 = umma_desc_add(da0, (uint32_t)((2 * Cfg::A_BYTES) >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_add(da0, k * 2), dw = umma_desc_add(dw0, k * 2);
            umma_bf16(d_t, da, dw, idesc, (kb | k) != 0);
            umma_bf16(d_t, umma_desc_add(da, Cfg::A_BYTES >> 4), dw, idesc, 1);
            umma_bf16(d_t, da, umma_desc_add(dw, Cfg::W_BYTES >> 4), idesc, 1);
          }
          umma_commit(&empty_bar[st]);
          if (kb == e.nkb - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool elected = (warp == 2 && lane == 0);
    uint32_t ti = 0, ng = 0;        // ng: residual groups consumed so far (res_bar phase)
    auto load_res = [&](int m0, int c) {      // fp32 residual of 64 columns [c, c + 64): two 32-column tiles
      mbar_arrive_expect_tx(&res_bar, RES_BUF_BYTES);
      tma_load_2d(res_stage, &tm.rf, &res_bar, c, m0);
      tma_load_2d(res_stage + TILE_BYTES, &tm.rf, &res_bar, c + 32, m0);
    };
    if (e.has_r && elected && (int)blockIdx.x < n_tiles) load_res((blockIdx.x / n_tiles_n) * 128, (blockIdx.x % n_tiles_n) * 128);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
      const uint32_t buf = ti & 1u;
      const int m0 = (t / n_tiles_n) * 128, n0 = (t % n_tiles_n) * 128;
      const int tn = t + gridDim.x;
      mbar_wait(&acc_full[buf], (ti >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        // the stores of the previous group must have read the staging buffer before it is rewritten
        if (elected) tma_store_wait_read<0>();
        epi_bar_sync();
        if (e.has_r) { mbar_wait(&res_bar, ng & 1u); ++ng; }
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int c0 = g * 64 + half * 32;
          uint32_t raw[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128u + (uint32_t)c0, raw);
          tmem_ld_wait();
          if (g == 1 && half == 1) {          // the whole accumulator is in registers / staged: the MMAs of tile ti + 2 may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
          }
          const int n = n0 + c0;
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(raw[i]);
          if (e.bias) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n + i));
              f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
            }
          }
          if (e.act != ACT_NONE) epi_act32(f, e.act);
          if (e.has_r) {
            const uint8_t* rt = res_stage + half * TILE_BYTES;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 r = *reinterpret_cast<const float4*>(rt + sw128(row, j));
              f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
            }
          }
          if (e.has_yf) {
            uint8_t* tt = sbuf + (2 + half) * TILE_BYTES;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(tt + sw128(row, j)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
          if (e.has_yh) {
            uint32_t hb[16], lb[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) split_bf16x2(f[2 * i], f[2 * i + 1], hb[i], lb[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(sbuf + sw128(row, half * 4 + j)) = make_uint4(hb[4 * j], hb[4 * j + 1], hb[4 * j + 2], hb[4 * j + 3]);
            if (e.has_yl) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(sbuf + TILE_BYTES + sw128(row, half * 4 + j)) =
                    make_uint4(lb[4 * j], lb[4 * j + 1], lb[4 * j + 2], lb[4 * j + 3]);
            }
          }
        }
        fence_proxy_async();
        epi_bar_sync();                 // staging complete; every thread is past its reads of the residual tiles
        if (elected) {
          const int c = n0 + g * 64;
          if (e.has_yh) tma_store_2d(&tm.yh, sbuf, c, m0);
          if (e.has_yl) tma_store_2d(&tm.yl, sbuf + TILE_BYTES, c, m0);
          if (e.has_yf) {
            tma_store_2d(&tm.yf, sbuf + 2 * TILE_BYTES, c, m0);
            tma_store_2d(&tm.yf, sbuf + 3 * TILE_BYTES, c + 32, m0);
          }
          tma_store_commit();
          if (e.has_r) {                // next residual group: the second half of this tile, or the first of this CTA's next tile
            if (g == 0) load_res(m0, n0 + 64);
            else if (tn < n_tiles) load_res((tn / n_tiles_n) * 128, (tn % n_tiles_n) * 128);
          }
        }
      }
    }
    if (elected) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
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

// [rows, cols] matrix with row pitch ld (elements); box = box_rows x 128 bytes, SWIZZLE_128B, OOB reads -> zeros
static int make_map(CUtensorMap* map, const void* ptr, bool f32, int rows, int cols, int ld, int box_rows) {
  const int esz = f32 ? 4 : 2;
  SEEME_REQUIRE(ptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ((size_t)ld * esz) % 16 == 0, SEEME_EINVAL,
                "umma_linear: tensor base must be 16-byte aligned and its pitch a multiple of 16 bytes (ld=%d)", ld);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim,
                        gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SEEME_REQUIRE(r == CUDA_SUCCESS, SEEME_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%d cols=%d ld=%d)", (int)r, rows, cols, ld);
  return SEEME_OK;
}

int umma_tensor_map_bf16(CUtensorMap* map, const void* ptr, int rows, int cols, int ld_elems, int box_rows) {
  SEEME_TRY(get_encoder());
  return make_map(map, ptr, false, rows, cols, ld_elems, box_rows);
}

template <int BN, int NPASS, int STAGES, int NBUF, int MINB>
static int launch(const UmmaLinear& g, const UmmaMaps& maps, const UmmaEpi& e, cudaStream_t s) {
  using Cfg = UmmaCfg<BN, NPASS>;
  constexpr int ring = Cfg::STAGE_BYTES * STAGES, store = NBUF * STORE_BUF_BYTES;
  // BN = 256 tiles never stage a residual (umma_linear picks BN <= 128 when R is given)
  constexpr int res_max = BN == 256 ? 0 : NBUF * RES_BUF_BYTES;
  constexpr int smem_max = (ring > store ? ring : store) + res_max + 1024;
  static_assert(smem_max <= 227 * 1024, "shared memory budget exceeded");
  static bool configured = false;
  if (!configured) {
    SEEME_CUDA(cudaFuncSetAttribute(umma_linear_kernel<BN, NPASS, STAGES, NBUF, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    smem_max));
    configured = true;
  }
  SEEME_REQUIRE(!(e.has_r && BN == 256), SEEME_EINVAL, "umma_linear: residual staging is not available for 256-wide tiles");
  const int smem = (ring > store ? ring : store) + (e.has_r ? res_max : 0) + 1024;
  dim3 grid(g.N / BN, (g.M + 127) / 128);
  ProfScope prof(g.prof_id - 1, s);
  SEEME_CUDA(launch_pdl(umma_linear_kernel<BN, NPASS, STAGES, NBUF, MINB>, grid, dim3(192), (size_t)smem, s, maps, e));
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

int umma_linear(const UmmaLinear& g, int npass, cudaStream_t s) {
  SEEME_TRY(get_encoder());
  const int K = g.K1 + g.K2;
  SEEME_REQUIRE(g.M > 0 && g.N > 0 && g.K1 > 0 && g.K1 % 64 == 0 && g.K2 % 64 == 0, SEEME_EINVAL,
                "umma_linear: unsupported shape M=%d N=%d K1=%d K2=%d (K multiples of 64)", g.M, g.N, g.K1, g.K2);
  SEEME_REQUIRE(npass == 1 || npass == 3, SEEME_EINVAL, "umma_linear: npass must be 1 or 3");
  SEEME_REQUIRE(!g.fp16 || (npass == 1 && !g.Yl && !g.Zl), SEEME_EINVAL, "umma_linear: fp16 operands are single-pass (no lo parts)");
  SEEME_REQUIRE(npass == 1 || (g.A1.lo && g.W.lo && (g.K2 == 0 || g.A2.lo)), SEEME_EINVAL, "umma_linear: split-bf16 needs lo operands");
  SEEME_REQUIRE(!g.colmax || g.colmax_group_rows >= 128, SEEME_EINVAL, "umma_linear: colmax groups must have >= 128 rows");
  SEEME_REQUIRE(!(g.Y && g.Zh), SEEME_EINVAL, "umma_linear: the fp32 output and the relu bf16 output share staging slots");
  SEEME_REQUIRE((!g.Yl || g.Yh) && (!g.Zl || g.Zh), SEEME_EINVAL, "umma_linear: lo outputs need their hi companion");
  // few rows (the latency-bound sampler / per-sample GEMMs): narrow tiles spread one GEMM over more SMs;
  // many rows: the widest tile that divides N (128 when a residual tile has to be staged as well)
  // SEEME_UMMA_SMALL_BN=128: experiment knob (fewer, fatter CTAs for the few-row GEMMs of the sampler chain)
  static const int small_bn = seeme_exp_env("SEEME_UMMA_SMALL_BN") ? atoi(seeme_exp_env("SEEME_UMMA_SMALL_BN")) : 64;
  int BN = (g.M <= 2048 && !g.colmax) ? (small_bn == 128 ? 128 : 64) : 128;
  if (g.N % BN != 0) BN = 64;
  SEEME_REQUIRE(g.N % BN == 0, SEEME_EINVAL, "umma_linear: N=%d must be a multiple of 64", g.N);
  UmmaMaps maps;
  memset(&maps, 0, sizeof(maps));
  SEEME_TRY(make_map(&maps.a1h, g.A1.hi, false, g.M, g.K1, g.A1.ld, 128));
  if (npass == 3) SEEME_TRY(make_map(&maps.a1l, g.A1.lo, false, g.M, g.K1, g.A1.ld, 128)); else maps.a1l = maps.a1h;
  if (g.K2) {
    SEEME_TRY(make_map(&maps.a2h, g.A2.hi, false, g.M, g.K2, g.A2.ld, 128));
    if (npass == 3) SEEME_TRY(make_map(&maps.a2l, g.A2.lo, false, g.M, g.K2, g.A2.ld, 128)); else maps.a2l = maps.a2h;
  } else { maps.a2h = maps.a1h; maps.a2l = maps.a1l; }
  SEEME_TRY(make_map(&maps.wh, g.W.hi, false, g.N, K, g.W.ld, BN));
  if (npass == 3) SEEME_TRY(make_map(&maps.wl, g.W.lo, false, g.N, K, g.W.ld, BN)); else maps.wl = maps.wh;
  maps.yh = maps.yl = maps.zh = maps.zl = maps.yf = maps.rf = maps.a1h;   // placeholders for unused outputs
  if (g.Yh) SEEME_TRY(make_map(&maps.yh, g.Yh, false, g.M, g.N, g.ldb, 128));
  if (g.Yl) SEEME_TRY(make_map(&maps.yl, g.Yl, false, g.M, g.N, g.ldb, 128));
  if (g.Zh) SEEME_TRY(make_map(&maps.zh, g.Zh, false, g.M, g.N, g.ldb, 128));
  if (g.Zl) SEEME_TRY(make_map(&maps.zl, g.Zl, false, g.M, g.N, g.ldb, 128));
  if (g.Y) SEEME_TRY(make_map(&maps.yf, g.Y, true, g.M, g.N, g.ldy, 128));
  if (g.R) SEEME_TRY(make_map(&maps.rf, g.R, true, g.M, g.N, g.ldr, 128));
  UmmaEpi e;
  e.M = g.M; e.N = g.N; e.nkb1 = g.K1 / 64; e.nkb = K / 64;
  e.bias = g.bias; e.bias_group_rows = g.bias_group_rows; e.pfold = g.pfold; e.xyz = g.xyz; e.act = g.act;
  e.has_r = g.R != nullptr; e.has_yf = g.Y != nullptr; e.has_yh = g.Yh != nullptr; e.has_yl = g.Yl != nullptr;
  e.has_zh = g.Zh != nullptr; e.has_zl = g.Zl != nullptr;
  e.act_post = g.act_after_residual;
  e.colmax = g.colmax; e.colmax_group_rows = g.colmax_group_rows;
  e.fp16 = g.fp16;
  // many rows, split precision, plain epilogue: the persistent kernel (one CTA per SM, MMAs of the next tile under this tile's epilogue)
  if (BN == 128 && npass == 3 && g.M >= 2048 && !g.fp16 && !g.pfold && !g.colmax && !g.Zh && !g.bias_group_rows && !g.act_after_residual &&
      !g.dense_ctas && (!g.Yl || g.Yh)) {
    static bool configured = false;
    if (!configured) {
      SEEME_CUDA(cudaFuncSetAttribute(umma_plinear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PL_SMEM));
      configured = true;
    }
    const int tiles_n = g.N / 128, tiles = tiles_n * ((g.M + 127) / 128);
    const int grid = tiles < NUM_SMS ? tiles : NUM_SMS;
    ProfScope prof(g.prof_id - 1, s);
    SEEME_CUDA(launch_pdl(umma_plinear_kernel, dim3(grid), dim3(192), (size_t)PL_SMEM, s, maps, e, tiles_n, tiles));
    SEEME_LAUNCH_CHECK();
    return SEEME_OK;
  }
  // 64-wide tiles: the whole K = 256 of a latency-bound GEMM is in flight at once (4 stages), one CTA per SM.
  // 128-wide tiles (many rows): a shallow ring (64 KB in split mode) and one staging buffer keep the footprint
  // small enough for 2-3 CTAs per SM, which is what overlaps TMA, MMA and epilogue phases across tiles.
  if (BN == 64) {
    static const bool small = !(seeme_exp_env("SEEME_UMMA64_STAGES4") && seeme_exp_env("SEEME_UMMA64_STAGES4")[0] == '1');
    // SEEME_UMMA64_STAGES=1: experiment knob (one-stage ring, 65 KB: three CTAs per SM when several chains share the GPU)
    static const bool one = seeme_exp_env("SEEME_UMMA64_STAGES") && seeme_exp_env("SEEME_UMMA64_STAGES")[0] == '1';
    if (npass == 1) return launch<64, 1, 4, 1, 1>(g, maps, e, s);
    if (one || g.dense_ctas) return launch<64, 3, 1, 1, 3>(g, maps, e, s);
    return small ? launch<64, 3, 2, 1, 2>(g, maps, e, s) : launch<64, 3, 4, 1, 1>(g, maps, e, s);
  }
  // SEEME_UMMA_DENSE_ALL=1: experiment knob -- the three-CTA build for every many-row split-bf16 GEMM (VAE stacks)
  static const bool dense_all = seeme_exp_env("SEEME_UMMA_DENSE_ALL") && seeme_exp_env("SEEME_UMMA_DENSE_ALL")[0] == '1';
  if ((g.dense_ctas || (dense_all && !g.colmax)) && npass == 3) return launch<128, 3, 1, 1, 3>(g, maps, e, s);
  return npass == 1 ? launch<128, 1, 2, 1, 2>(g, maps, e, s) : launch<128, 3, 1, 1, 2>(g, maps, e, s);
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

__global__ void bf16_join_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, float* __restrict__ out,
                                 size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}

}  // namespace seeme

// ---- test entry point: Y = act(A W^T + bias) (+R) with fp32 inputs converted on the fly -----------------
extern "C" int seeme_test_umma_linear(const float* A, const float* W, const float* bias, const float* R, float* Y, int M, int N,
                                      int K, int act, int npass, unsigned* colmax, int colmax_group_rows, float* Ysplit,
                                      float* Zsplit, void* stream) {
  using namespace seeme;
  SEEME_REQUIRE(A && W && (Y || Ysplit), SEEME_EINVAL, "seeme_test_umma_linear: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  __nv_bfloat16 *ah, *al, *wh, *wl, *ob = nullptr;
  SEEME_CUDA(cudaMalloc(&ah, (size_t)M * K * 2));
  SEEME_CUDA(cudaMalloc(&al, (size_t)M * K * 2));
  SEEME_CUDA(cudaMalloc(&wh, (size_t)N * K * 2));
  SEEME_CUDA(cudaMalloc(&wl, (size_t)N * K * 2));
  SEEME_CUDA(cudaMalloc(&ob, (size_t)M * N * 2 * 4));
  const size_t mn = (size_t)M * N;
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
    g.bias = bias; g.act = act; g.R = R; g.ldr = N;
    if (Y && !Zsplit) { g.Y = Y; g.ldy = N; }
    if (Ysplit) { g.Yh = ob; g.Yl = ob + mn; g.ldb = N; }
    if (Zsplit) { g.Zh = ob + 2 * mn; g.Zl = ob + 3 * mn; g.ldb = N; }
    g.colmax = colmax; g.colmax_group_rows = colmax_group_rows;
    rc = umma_linear(g, npass, s);
    if (!rc && Ysplit) bf16_join_kernel<<<(unsigned)((mn + 255) / 256), 256, 0, s>>>(ob, ob + mn, Ysplit, mn);
    if (!rc && Zsplit) bf16_join_kernel<<<(unsigned)((mn + 255) / 256), 256, 0, s>>>(ob + 2 * mn, ob + 3 * mn, Zsplit, mn);
  }
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(ah); cudaFree(al); cudaFree(wh); cudaFree(wl); cudaFree(ob);
  if (!rc && e != cudaSuccess) { set_error("seeme_test_umma_linear: %s", cudaGetErrorString(e)); rc = SEEME_ECUDA; }
  return rc;
}
