// SMPL blend shapes + linear-blend skinning with the blend contraction on the tensor cores.
//
//   v_posed[f, v, c] = v_template[v, c] + sum_k coef[f, k] * basis[k, c, v]       (K = 10 shape + 207 pose = 217)
//   verts[f, v, :]   = (sum_{n<4} w[v,n] A[f, j[v,n]]) . [v_posed[f, v, :]; 1]
//
// smplx evaluates the first line as two fp32 GEMMs and materialises v_shaped / v_posed / T [F,6890,4,4] in HBM.
// Here one CTA owns a group of 64 frames (their coef rows are the resident B operand, their joint transforms sit in
// shared memory) and streams 128-vertex basis tiles from L2: per tile and coordinate plane a [128 v x 256 k] x
// [256 k x 64 f] tcgen05 GEMM in split-bf16 (hi.hi + lo.hi + hi.lo, ~16 mantissa bits, fp32 accumulation) leaves
// x, y, z of a vertex in the SAME TMEM lane (three 64-column accumulators), so the skinning epilogue is thread = vertex
// with no shuffles, and the only HBM traffic is the 82 680 B per frame of output vertices.
//
//   warp 0      TMA producer: basis chunks [128 v x 64 k] (hi, lo) through a 5-slot ring
//   warp 1      MMA issuer (elected lane); accumulators double-buffered in TMEM (2 x 3 x 64 columns)
//   warps 2..17 epilogue: lane quarter x frame quarter; tcgen05.ld of x/y/z for 16 frames, sparse (<= 4) skinning, stores
//               (16 warps: the per-frame LDS -> FFMA chains are latency-bound, so occupancy is what hides them)
#include "umma.cuh"
#include "smpl_tc.cuh"
#include <stdlib.h>
#include <cuda_fp16.h>

namespace seeme {

constexpr int ST_V = 6890, ST_VP = 6912, ST_J = 24, ST_NF = 64, ST_KP = 256, ST_NST = 5;
constexpr int ST_CHUNK = 128 * 128;                  // [128 v x 64 k] bf16
constexpr int ST_BCHUNK = 64 * 128;                  // [64 f x 64 k] bf16
constexpr int ST_B_BYTES = 4 * 2 * ST_BCHUNK;        // 4 k-chunks x (hi, lo)
constexpr int ST_A_BYTES = ST_NF * ST_J * 12 * 4;    // joint transforms of the frame group
constexpr int ST_SMEM = ST_B_BYTES + ST_A_BYTES + ST_NST * ST_CHUNK + 1024;
constexpr int ST_THREADS = 576;                     // 2 control warps + 16 epilogue warps

struct StMaps { CUtensorMap bh, bl, ch, cl; };
struct StArgs {
  const float* A;            // [F,24,12]
  const float* vt;           // [3][VP]
  const float* w4;           // [VP][4]
  const unsigned char* i4;   // [VP][4]
  float* verts;              // [F,6890,3]
  int F, tiles_per_cta;
};

__device__ __forceinline__ void st_tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

#ifdef SEEME_EXPERIMENTAL   // retired variant (v1: sparse skinning from shared memory in the epilogue)
// CL: thread-block cluster size along the frame-group axis.  The CTAs of a cluster work on different frame groups but the
// SAME vertex tiles, so the basis chunk stream is identical: chunk n is fetched from L2 once, by CTA n % CL, and multicast
// into the ring slot of every CTA of the cluster (the kernel is bound by this L2 -> SM stream, not by the MMAs).
template <int CL>
__global__ void __launch_bounds__(ST_THREADS, 1) smpl_skin_tc_kernel(const __grid_constant__ StMaps tm, const StArgs a) {
  extern __shared__ __align__(1024) uint8_t st_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(st_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bop = smem;                                   // coef B operand: [kc][hi|lo][64 f x 64 k]
  uint8_t* ring = smem + ST_B_BYTES;
  float* As = reinterpret_cast<float*>(smem + ST_B_BYTES + ST_NST * ST_CHUNK);
  __shared__ __align__(8) uint64_t r_full[ST_NST], r_empty[ST_NST], b_full, acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f0 = blockIdx.x * ST_NF;
  const int t0 = blockIdx.y * a.tiles_per_cta, nt = a.tiles_per_cta;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm.bh); tma_prefetch_desc(&tm.bl); tma_prefetch_desc(&tm.ch); tma_prefetch_desc(&tm.cl);
    for (int i = 0; i < ST_NST; ++i) { mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], CL); }
    mbar_init(&b_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_free[i], 16); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  // joint transforms of this frame group (zero beyond F)
  {
    const size_t base = (size_t)f0 * ST_J * 12;
    const int nvalid = (a.F - f0 < ST_NF ? a.F - f0 : ST_NF) * ST_J * 12;
    for (int i = threadIdx.x; i < ST_NF * ST_J * 12; i += ST_THREADS) As[i] = i < nvalid ? __ldg(a.A + base + i) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  uint32_t crank = 0;
  if (CL > 1) {
    cluster_sync_all();          // every CTA's barriers exist before any remote arrive / multicast write
    crank = cluster_ctarank();
  }
  constexpr uint16_t cmask = (uint16_t)((1u << CL) - 1u);

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&b_full, ST_B_BYTES);
      for (int kc = 0; kc < 4; ++kc) {
        tma_load_2d(bop + (kc * 2 + 0) * ST_BCHUNK, &tm.ch, &b_full, kc * 64, f0);
        tma_load_2d(bop + (kc * 2 + 1) * ST_BCHUNK, &tm.cl, &b_full, kc * 64, f0);
      }
      uint32_t st = 0, ph = 1, n = 0;
      for (int i = 0; i < nt; ++i) {
        const int tile = t0 + i;
        for (int c = 0; c < 3; ++c)
          for (int kc = 0; kc < 4; ++kc)
            for (int hl = 0; hl < 2; ++hl, ++n) {
              mbar_wait(&r_empty[st], ph);        // the slot has been consumed by every CTA of the cluster
              mbar_arrive_expect_tx(&r_full[st], ST_CHUNK);
              if (CL == 1) tma_load_2d(ring + st * ST_CHUNK, hl ? &tm.bl : &tm.bh, &r_full[st], kc * 64, c * ST_VP + tile * 128);
              else if (n % CL == crank)
                tma_load_2d_mc(ring + st * ST_CHUNK, hl ? &tm.bl : &tm.bh, &r_full[st], kc * 64, c * ST_VP + tile * 128, cmask);
              if (++st == ST_NST) { st = 0; ph ^= 1u; }
            }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(ST_NF);
    const uint64_t rdesc0 = umma_desc_k128(smem_u32(ring));
    const uint64_t bdesc0 = umma_desc_k128(smem_u32(bop));
    uint32_t st = 0, ph = 0;
    mbar_wait(&b_full, 0);
    for (int i = 0; i < nt; ++i) {
      const int p = i & 1;
      mbar_wait(&acc_free[p], ((uint32_t)(i >> 1) & 1u) ^ 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        const uint32_t d = tmem_base + (uint32_t)(p * 192 + c * 64);
#pragma unroll 1
        for (int kc = 0; kc < 4; ++kc) {
          const uint64_t bh = umma_desc_add(bdesc0, (uint32_t)((kc * 2 + 0) * (ST_BCHUNK >> 4)));
          const uint64_t bl = umma_desc_add(bdesc0, (uint32_t)((kc * 2 + 1) * (ST_BCHUNK >> 4)));
          // slot with the hi part of the basis chunk: hi.hi and hi.lo
          mbar_wait(&r_full[st], ph);
          tc_fence_after();
          if (umma_elect_one()) {
            const uint64_t ad = umma_desc_add(rdesc0, st * (ST_CHUNK >> 4));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_bf16(d, umma_desc_add(ad, ks * 2), umma_desc_add(bh, ks * 2), idesc, (kc | ks) != 0);
              umma_bf16(d, umma_desc_add(ad, ks * 2), umma_desc_add(bl, ks * 2), idesc, 1);
            }
            if (CL == 1) umma_commit(&r_empty[st]); else umma_commit_mc(&r_empty[st], cmask);
          }
          __syncwarp();
          if (++st == ST_NST) { st = 0; ph ^= 1u; }
          // slot with the lo part: lo.hi
          mbar_wait(&r_full[st], ph);
          tc_fence_after();
          if (umma_elect_one()) {
            const uint64_t ad = umma_desc_add(rdesc0, st * (ST_CHUNK >> 4));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_bf16(d, umma_desc_add(ad, ks * 2), umma_desc_add(bh, ks * 2), idesc, 1);
            if (CL == 1) umma_commit(&r_empty[st]); else umma_commit_mc(&r_empty[st], cmask);
            if (c == 2 && kc == 3) umma_commit(&acc_full[p]);
          }
          __syncwarp();
          if (++st == ST_NST) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    const int q = warp & 3;                  // TMEM lane quarter
    const int fq = (warp - 2) >> 2;          // frame quarter: frames [16 fq, 16 fq + 16) of the group
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    for (int i = 0; i < nt; ++i) {
      const int p = i & 1;
      const int v = (t0 + i) * 128 + q * 32 + lane;
      float w[4];
      int jo[4];
      {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(a.w4) + v);
        const uchar4 iv = __ldg(reinterpret_cast<const uchar4*>(a.i4) + v);
        w[0] = wv.x; w[1] = wv.y; w[2] = wv.z; w[3] = wv.w;
        jo[0] = iv.x * 12; jo[1] = iv.y * 12; jo[2] = iv.z * 12; jo[3] = iv.w * 12;
      }
      const float vx = __ldg(a.vt + v), vy = __ldg(a.vt + ST_VP + v), vz = __ldg(a.vt + 2 * ST_VP + v);
      mbar_wait(&acc_full[p], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      uint32_t X[16], Y[16], Z[16];
      const uint32_t tb = tmem_base + lane_off + (uint32_t)(p * 192 + fq * 16);
      st_tmem_ld16(tb, X);
      st_tmem_ld16(tb + 64, Y);
      st_tmem_ld16(tb + 128, Z);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_free[p]);     // the accumulators are in registers: the MMAs of tile i+2 may proceed
      if (v < ST_V) {
        float* out = a.verts + ((size_t)(f0 + fq * 16) * ST_V + v) * 3;
        const int nf = a.F - (f0 + fq * 16);
#pragma unroll
        for (int f = 0; f < 16; ++f) {
          if (f < nf) {
            const float* Af = As + (fq * 16 + f) * (ST_J * 12);
            float T[12];
#pragma unroll
            for (int e = 0; e < 12; ++e) T[e] = 0.f;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              const float4* ap = reinterpret_cast<const float4*>(Af + jo[n]);
              const float4 a0 = ap[0], a1 = ap[1], a2 = ap[2];
              T[0] = fmaf(w[n], a0.x, T[0]); T[1] = fmaf(w[n], a0.y, T[1]); T[2] = fmaf(w[n], a0.z, T[2]); T[3] = fmaf(w[n], a0.w, T[3]);
              T[4] = fmaf(w[n], a1.x, T[4]); T[5] = fmaf(w[n], a1.y, T[5]); T[6] = fmaf(w[n], a1.z, T[6]); T[7] = fmaf(w[n], a1.w, T[7]);
              T[8] = fmaf(w[n], a2.x, T[8]); T[9] = fmaf(w[n], a2.y, T[9]); T[10] = fmaf(w[n], a2.z, T[10]); T[11] = fmaf(w[n], a2.w, T[11]);
            }
            const float x = vx + __uint_as_float(X[f]), y = vy + __uint_as_float(Y[f]), z = vz + __uint_as_float(Z[f]);
            float* o = out + (size_t)f * ST_V * 3;
            o[0] = T[0] * x + T[1] * y + T[2] * z + T[3];
            o[1] = T[4] * x + T[5] * y + T[6] * z + T[7];
            o[2] = T[8] * x + T[9] * y + T[10] * z + T[11];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

#endif  // SEEME_EXPERIMENTAL

// ---- no-swizzle ("interleaved") K-major operand tiles with 64-byte rows (K = 32 fp16) -------------------------------------
// canonical layout (cute: INTERLEAVE, ((8,m),(T,2)):((1T,SBO),(1,LBO)) in 16-byte units): core matrix = 8 rows x 16 bytes,
// stored contiguously (128 B); the 4 core matrices of an 8-row group along K follow each other (LBO = 128 B), 8-row groups
// are SBO = 512 B apart.  Element (r, k) of a [rows x 32] fp16 tile sits at byte (r/8)*512 + (k/8)*128 + (r%8)*16 + (k%8)*2.
__host__ __device__ constexpr uint32_t st_il_off(int r, int k) { return (uint32_t)((r >> 3) * 512 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2); }
__device__ __forceinline__ uint64_t st_desc_il(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46);
}

// micro-test of the descriptor above (tools/test_interleave.py): D[128, 48] = A[128, 32] . B[48, 32]^T, fp16 in, fp32 out
__global__ void __launch_bounds__(128, 1) st_interleave_test_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D) {
  __shared__ __align__(1024) uint8_t sa[128 * 64];
  __shared__ __align__(1024) uint8_t sb[48 * 64];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * 32; i += 128) *reinterpret_cast<__half*>(sa + st_il_off(i / 32, i % 32)) = A[i];
  for (int i = threadIdx.x; i < 48 * 32; i += 128) *reinterpret_cast<__half*>(sb + st_il_off(i / 32, i % 32)) = B[i];
  if (threadIdx.x == 0) { mbar_init(&done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  if (threadIdx.x == 0) {
    const uint64_t ad = st_desc_il(smem_u32(sa)), bd = st_desc_il(smem_u32(sb));
    umma_bf16(tb, ad, bd, umma_idesc_f16(48), 0);
    umma_bf16(tb, umma_desc_add(ad, 256 >> 4), umma_desc_add(bd, 256 >> 4), umma_idesc_f16(48), 1);     // K = 16..31: two core matrices on
    umma_commit(&done);
  }
  mbar_wait(&done, 0);
  tc_fence_after();
  uint32_t v[48];
  const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16);
  st_tmem_ld16(ta, v); st_tmem_ld16(ta + 16, v + 16); st_tmem_ld16(ta + 32, v + 32);
  tmem_ld_wait();
  for (int j = 0; j < 48; ++j) D[(warp * 32 + lane) * 48 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 64); }
}

extern "C" int seeme_test_umma_interleave(const void* A, const void* B, float* D, void* stream) {
  st_interleave_test_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(A), reinterpret_cast<const __half*>(B), D);
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Version 2: the per-(vertex, frame) blend of the joint transforms  T[v,f] = sum_j W[v,j] A[f,j]  runs on the tensor cores
// too.  ncu on the kernel above: 12 x LDS.128 per (vertex, frame) = 192 B of shared-memory reads, i.e. the 128 B/clk
// shared-memory bandwidth bounds it (tensor pipe 24 %, DRAM 15 %).  Here, per 128-vertex tile and group of 4 frames,
//     T[128 v, 48 (f, e)] = Wtile[128 v, 32 j] . Aop[48 (f, e), 32 j]^T          (K = 24 joints padded to 32)
// in split fp16 (hi.hi + lo.hi + hi.lo, fp32 accumulation: ~2^-22), and the epilogue reads the 12 blended entries of a
// (vertex, frame) from tensor memory (its own lane) next to x, y, z: no shared-memory reads at all, and dense skinning
// weights cost the same as sparse ones.
//   warp 0      producer: per tile the 16 KB W tile (pre-packed no-swizzle image, double-buffered), then 15 basis chunks
//               (fp16 hi for all K, lo for K-chunk 0 only; 5-slot ring)
//   warp 1      blend-shape MMA issuer (single accumulator set: x / y / z = 3 x 64 columns; the epilogue copies it to registers first)
//   warps 2-17  epilogue, lane quarter q x frame quarter fq (16 frames).  Per tile and 4-frame sub-block the 4 warps of a
//               frame quarter convert the sub-block's joint transforms to fp16 (hi, lo) into their own 6 KB operand buffer
//               (next sub-block prefetched from L2 into registers), wait for T, apply it, store.
//   warp 18     transform-blend MMA issuer: round-robin over the 4 frame quarters, 6 MMAs (N = 48) per sub-block into the
//               quarter's own 48 TMEM columns
__global__ void smpl_coef_split_f16_kernel(const float* __restrict__ coef, int ld, int F, int n, __half* __restrict__ ch, __half* __restrict__ cl);

constexpr int S2_NST = 5;
constexpr int S2_WT = 128 * 64 * 2;                    // W tile: hi 8 KB | lo 8 KB
constexpr int S2_AOP = 48 * 64 * 2;                    // Aop buffer of one frame quarter: hi 3 KB | lo 3 KB
constexpr int S2_SMEM = ST_B_BYTES + S2_NST * ST_CHUNK + 2 * S2_WT + 8 * S2_AOP + 1024;     // Aop: 2 buffers per frame quarter
constexpr int S2_THREADS = 608;
constexpr uint32_t S2_TCOL = 192;                      // first T column (after x | y | z)

struct S2Args {
  const float* A;            // [F,24,12]
  const float* vt;           // [3][VP]
  const uint8_t* wblob;      // [54] x 16 KB W tiles
  const uint8_t* aopblob;    // [ceil(F/4)] x 6 KB transform operands of 4 frames (hi 3 KB | lo 3 KB), written by smpl_aop_pack_kernel
  float* verts;              // [F,6890,3]
  int F, tiles_per_cta;
};

__device__ __forceinline__ void st_tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void st_tmem_ld4(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(S2_THREADS, 1) smpl_skin_tc2_kernel(const __grid_constant__ StMaps tm, const S2Args a) {
  extern __shared__ __align__(1024) uint8_t st_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(st_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bop = smem;                                   // coef B operand: [kc][hi|lo][64 f x 64 k]
  uint8_t* ring = smem + ST_B_BYTES;
  uint8_t* wt = ring + S2_NST * ST_CHUNK;
  uint8_t* aop = wt + 2 * S2_WT;
  __shared__ __align__(8) uint64_t r_full[S2_NST], r_empty[S2_NST], b_full, acc_full, acc_free, wt_full[2], wt_free[2], aop_full[2], t_free, t_full;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f0 = blockIdx.x * ST_NF;
  const int t0 = blockIdx.y * a.tiles_per_cta, nt = a.tiles_per_cta;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm.bh); tma_prefetch_desc(&tm.bl); tma_prefetch_desc(&tm.ch); tma_prefetch_desc(&tm.cl);
    for (int i = 0; i < S2_NST; ++i) { mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], 1); }
    mbar_init(&b_full, 1);
    mbar_init(&acc_full, 1);
    mbar_init(&acc_free, 16);
    for (int i = 0; i < 2; ++i) { mbar_init(&wt_full[i], 1); mbar_init(&wt_free[i], 1); }
    for (int i = 0; i < 2; ++i) mbar_init(&aop_full[i], 4);      // one arrive.expect_tx per frame quarter
    mbar_init(&t_free, 16);
    mbar_init(&t_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&b_full, ST_B_BYTES);
      for (int kc = 0; kc < 4; ++kc) {
        tma_load_2d(bop + (kc * 2 + 0) * ST_BCHUNK, &tm.ch, &b_full, kc * 64, f0);
        tma_load_2d(bop + (kc * 2 + 1) * ST_BCHUNK, &tm.cl, &b_full, kc * 64, f0);
      }
      uint32_t st = 0, ph = 1;
      for (int i = 0; i < nt; ++i) {
        const int tile = t0 + i;
        mbar_wait(&wt_free[i & 1], ((uint32_t)(i >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&wt_full[i & 1], S2_WT);
        st_bulk_load(wt + (i & 1) * S2_WT, a.wblob + (size_t)tile * S2_WT, S2_WT, &wt_full[i & 1]);
        // fp16 basis: the lo halves are streamed for the first K-chunk only (the 10 shape directions + 54 pose directions);
        // a pose-direction term carries a few millimetres at most, so its fp16 rounding (2^-12) stays below 1e-6 m
        for (int c = 0; c < 3; ++c)
          for (int kc = 0; kc < 4; ++kc)
            for (int hl = 0; hl < (kc == 0 ? 2 : 1); ++hl) {
              mbar_wait(&r_empty[st], ph);
              mbar_arrive_expect_tx(&r_full[st], ST_CHUNK);
              tma_load_2d(ring + st * ST_CHUNK, hl ? &tm.bl : &tm.bh, &r_full[st], kc * 64, c * ST_VP + tile * 128);
              if (++st == S2_NST) { st = 0; ph ^= 1u; }
            }
      }
    }
  } else if (warp == 1) {
    // ---- blend-shape GEMM: x / y / z planes, [128 v x 256 k] x [256 k x 64 f] each, split fp16 (basis lo: K-chunk 0 only) ----
    constexpr uint32_t idesc = umma_idesc_f16(ST_NF);
    const uint64_t rdesc0 = umma_desc_k128(smem_u32(ring));
    const uint64_t bdesc0 = umma_desc_k128(smem_u32(bop));
    uint32_t st = 0, ph = 0;
    mbar_wait(&b_full, 0);
    for (int i = 0; i < nt; ++i) {
      mbar_wait(&acc_free, ((uint32_t)i & 1u) ^ 1u);       // the epilogue has copied tile i-1's x / y / z to registers
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        const uint32_t d = tmem_base + (uint32_t)(c * 64);
#pragma unroll 1
        for (int kc = 0; kc < 4; ++kc) {
          const uint64_t bh = umma_desc_add(bdesc0, (uint32_t)((kc * 2 + 0) * (ST_BCHUNK >> 4)));
          const uint64_t bl = umma_desc_add(bdesc0, (uint32_t)((kc * 2 + 1) * (ST_BCHUNK >> 4)));
          mbar_wait(&r_full[st], ph);
          tc_fence_after();
          if (umma_elect_one()) {
            const uint64_t ad = umma_desc_add(rdesc0, st * (ST_CHUNK >> 4));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_bf16(d, umma_desc_add(ad, ks * 2), umma_desc_add(bh, ks * 2), idesc, (kc | ks) != 0);
              umma_bf16(d, umma_desc_add(ad, ks * 2), umma_desc_add(bl, ks * 2), idesc, 1);
            }
            umma_commit(&r_empty[st]);
            if (c == 2 && kc == 3) umma_commit(&acc_full);
          }
          __syncwarp();
          if (++st == S2_NST) { st = 0; ph ^= 1u; }
          if (kc == 0) {
            mbar_wait(&r_full[st], ph);
            tc_fence_after();
            if (umma_elect_one()) {
              const uint64_t ad = umma_desc_add(rdesc0, st * (ST_CHUNK >> 4));
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma_bf16(d, umma_desc_add(ad, ks * 2), umma_desc_add(bh, ks * 2), idesc, 1);
              umma_commit(&r_empty[st]);
            }
            __syncwarp();
            if (++st == S2_NST) { st = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 18) {
    // ---- transform-blend GEMM issuer --------------------------------------------------------------------------------
    // ONE N = 192 group per 4-frame sub-block round: the operand images of the four frame quarters sit next to each other in a
    // buffer ([q0 | q1 | q2 | q3] hi, then lo), so 6 instructions serve all 16 frames (a tcgen05.mma with a small N costs
    // ~70 cycles whatever N is: 24 instead of 96 instructions per tile)
    constexpr uint32_t idesc = umma_idesc_f16(192);
    const uint64_t wdesc0 = st_desc_il(smem_u32(wt));
    const uint64_t adesc0 = st_desc_il(smem_u32(aop));
    for (int i = 0; i < nt; ++i) {
      mbar_wait(&wt_full[i & 1], (uint32_t)(i >> 1) & 1u);
      const uint64_t wh = umma_desc_add(wdesc0, (uint32_t)((i & 1) * (S2_WT >> 4))), wl = umma_desc_add(wh, 8192 >> 4);
#pragma unroll 1
      for (int r = 0; r < 4; ++r) {
        // round n = i * 4 + r: the four operand images have landed (buffer n & 1) and every frame quarter has read T of round n - 1
        const uint32_t n = (uint32_t)(i * 4 + r);
        mbar_wait(&aop_full[n & 1u], (n >> 1) & 1u);
        mbar_wait(&t_free, (n & 1u) ^ 1u);
        tc_fence_after();
        if (umma_elect_one()) {
          const uint64_t ah = umma_desc_add(adesc0, (uint32_t)((n & 1u) * ((4 * S2_AOP) >> 4))), al = umma_desc_add(ah, (4 * 3072) >> 4);
          const uint32_t d = tmem_base + S2_TCOL;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            umma_bf16(d, umma_desc_add(wh, ks * 16), umma_desc_add(ah, ks * 16), idesc, ks != 0);
            umma_bf16(d, umma_desc_add(wl, ks * 16), umma_desc_add(ah, ks * 16), idesc, 1);
            umma_bf16(d, umma_desc_add(wh, ks * 16), umma_desc_add(al, ks * 16), idesc, 1);
          }
          umma_commit(&t_full);
          if (r == 3) umma_commit(&wt_free[i & 1]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---- epilogue -------------------------------------------------------------------------------------------------------
    const int q = warp & 3;                  // TMEM lane quarter
    const int fq = (warp - 2) >> 2;          // frame quarter: frames [16 fq, 16 fq + 16) of the group
    const int tg = q * 32 + lane;            // thread within the frame quarter's 4 warps
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    // the transform operand of sub-block n (4 frames, 6 KB, pre-packed by smpl_aop_pack_kernel) arrives with one bulk copy into
    // buffer n & 1 of this frame quarter; thread 0 of the quarter issues it as soon as the MMAs of sub-block n - 2 are complete
    const bool loader = tg == 0;
    auto load_aop = [&](uint32_t n) {        // n = i * 4 + r over the CTA's tiles (the 4 sub-blocks repeat every tile)
      const int sb = (f0 + fq * 16) / 4 + (int)(n & 3u);
      uint8_t* buf = aop + (size_t)(n & 1u) * (4 * S2_AOP);        // [4 quarters x 3 KB hi | 4 quarters x 3 KB lo]
      mbar_arrive_expect_tx(&aop_full[n & 1u], S2_AOP);
      st_bulk_load(buf + fq * 3072, a.aopblob + (size_t)sb * S2_AOP, 3072, &aop_full[n & 1u]);
      st_bulk_load(buf + 4 * 3072 + fq * 3072, a.aopblob + (size_t)sb * S2_AOP + 3072, 3072, &aop_full[n & 1u]);
    };
    const uint32_t n_total = (uint32_t)nt * 4u;
    if (loader) { load_aop(0); if (n_total > 1) load_aop(1); }
    float vx, vy, vz;
    {
      const int v0 = t0 * 128 + q * 32 + lane;
      vx = __ldg(a.vt + v0); vy = __ldg(a.vt + ST_VP + v0); vz = __ldg(a.vt + 2 * ST_VP + v0);
    }
    for (int i = 0; i < nt; ++i) {
      const int v = (t0 + i) * 128 + q * 32 + lane;
      const float cx = vx, cy = vy, cz = vz;
      if (i + 1 < nt) { vx = __ldg(a.vt + v + 128); vy = __ldg(a.vt + ST_VP + v + 128); vz = __ldg(a.vt + 2 * ST_VP + v + 128); }
      uint32_t X[16], Y[16], Z[16];
      mbar_wait(&acc_full, (uint32_t)i & 1u);
      tc_fence_after();
      {
        const uint32_t tb = tmem_base + lane_off + (uint32_t)(fq * 16);
        st_tmem_ld16(tb, X);
        st_tmem_ld16(tb + 64, Y);
        st_tmem_ld16(tb + 128, Z);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_free);        // x / y / z are in registers: the blend GEMM of tile i+1 may proceed
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t n = (uint32_t)(i * 4 + r);
        mbar_wait(&t_full, n & 1u);
        tc_fence_after();
        // the MMAs of sub-block n are complete: its operand buffer is free for sub-block n + 2
        if (loader && n + 2 < n_total) load_aop(n + 2);
        const uint32_t tt = tmem_base + lane_off + S2_TCOL + (uint32_t)fq * 48u;
        const int fl = fq * 16 + r * 4;                    // first frame of the sub-block within the group
        float* out = a.verts + ((size_t)(f0 + fl) * ST_V + v) * 3;
        uint32_t T[4][12];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          st_tmem_ld8(tt + f * 12, T[f]);
          st_tmem_ld4(tt + f * 12 + 8, T[f] + 8);
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_free);              // T is in registers: the MMAs of round n + 1 may overwrite the columns
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          if (v < ST_V && f0 + fl + f < a.F) {
            const float x = cx + __uint_as_float(X[r * 4 + f]), y = cy + __uint_as_float(Y[r * 4 + f]), z = cz + __uint_as_float(Z[r * 4 + f]);
            float* o = out + (size_t)f * ST_V * 3;
            o[0] = __uint_as_float(T[f][0]) * x + __uint_as_float(T[f][1]) * y + __uint_as_float(T[f][2]) * z + __uint_as_float(T[f][3]);
            o[1] = __uint_as_float(T[f][4]) * x + __uint_as_float(T[f][5]) * y + __uint_as_float(T[f][6]) * z + __uint_as_float(T[f][7]);
            o[2] = __uint_as_float(T[f][8]) * x + __uint_as_float(T[f][9]) * y + __uint_as_float(T[f][10]) * z + __uint_as_float(T[f][11]);
          }
        }
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

// W [VP][24] fp32 (dense, rows >= 6890 zero) -> per 128-vertex tile the fp16 (hi | lo) no-swizzle operand image [128 x 32]
__global__ void smpl_wtile_pack_kernel(const float* __restrict__ w24, uint8_t* __restrict__ blob) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ST_VP * 32) return;
  const int v = i >> 5, k = i & 31, tile = v >> 7, r = v & 127;
  const float x = k < ST_J ? w24[(size_t)v * ST_J + k] : 0.f;
  const __half hi = __float2half_rn(x);
  uint8_t* t = blob + (size_t)tile * S2_WT;
  *reinterpret_cast<__half*>(t + st_il_off(r, k)) = hi;
  *reinterpret_cast<__half*>(t + 8192 + st_il_off(r, k)) = __float2half_rn(x - __half2float(hi));
}

// A [F,24,12] fp32 -> per 4 frames the fp16 (hi 3 KB | lo 3 KB) no-swizzle B-operand image [48 (f, e) x 32 j]; joints 24..31
// and frames >= F are zero
__global__ void smpl_aop_pack_kernel(const float* __restrict__ A, int F, uint8_t* __restrict__ blob, int n_sub) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)n_sub * 48 * 32) return;
  const int k = (int)(i & 31), row = (int)((i >> 5) % 48), sb = (int)(i / (48 * 32));
  const int f = sb * 4 + row / 12, e = row % 12;
  const float x = (k < ST_J && f < F) ? A[((size_t)f * ST_J + k) * 12 + e] : 0.f;
  const __half hi = __float2half_rn(x);
  uint8_t* t = blob + (size_t)sb * S2_AOP;
  *reinterpret_cast<__half*>(t + st_il_off(row, k)) = hi;
  *reinterpret_cast<__half*>(t + 3072 + st_il_off(row, k)) = __float2half_rn(x - __half2float(hi));
}

size_t smpl_tc_aop_bytes(size_t frames) { return (frames + ST_NF) / 4 * S2_AOP; }

size_t smpl_tc_wblob_bytes() { return (size_t)(ST_VP / 128) * S2_WT; }

int smpl_tc_pack_wtiles(const float* w24, void* wblob) {
  smpl_wtile_pack_kernel<<<(ST_VP * 32 + 255) / 256, 256>>>(w24, reinterpret_cast<uint8_t*>(wblob));
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

int smpl_skin_tc2(const void* bh, const void* bl, const float* coef, int ld_coef, int n_coef, void* ch, void* cl, const float* A,
                  const float* vt, const void* wblob, void* aopblob, int F, float* verts, int prof_id, cudaStream_t s) {
  const int groups = (F + ST_NF - 1) / ST_NF;
  {
    const int n_sub = groups * (ST_NF / 4);
    smpl_aop_pack_kernel<<<(unsigned)(((size_t)n_sub * 48 * 32 + 255) / 256), 256, 0, s>>>(A, F, reinterpret_cast<uint8_t*>(aopblob), n_sub);
    SEEME_LAUNCH_CHECK();
  }
  // coef [F, n_coef] fp32 -> fp16 (hi, lo) [F, 256]; the tensor maps only move 2-byte elements, so the bf16-typed maps serve
  smpl_coef_split_f16_kernel<<<(unsigned)(((size_t)F * n_coef + 255) / 256), 256, 0, s>>>(coef, ld_coef, F, n_coef, reinterpret_cast<__half*>(ch),
                                                                                       reinterpret_cast<__half*>(cl));
  SEEME_LAUNCH_CHECK();
  StMaps maps;
  memset(&maps, 0, sizeof(maps));
  SEEME_TRY(umma_tensor_map_bf16(&maps.bh, bh, 3 * ST_VP, ST_KP, ST_KP, 128));
  SEEME_TRY(umma_tensor_map_bf16(&maps.bl, bl, 3 * ST_VP, ST_KP, ST_KP, 128));
  SEEME_TRY(umma_tensor_map_bf16(&maps.ch, ch, F, ST_KP, ST_KP, ST_NF));
  SEEME_TRY(umma_tensor_map_bf16(&maps.cl, cl, F, ST_KP, ST_KP, ST_NF));
  static const int divs[8] = {1, 2, 3, 6, 9, 18, 27, 54};
  int vsplit = 54;
  for (int i = 0; i < 8; ++i)
    if (groups * divs[i] >= 4 * NUM_SMS) { vsplit = divs[i]; break; }
  S2Args a;
  a.A = A; a.vt = vt; a.wblob = reinterpret_cast<const uint8_t*>(wblob); a.aopblob = reinterpret_cast<const uint8_t*>(aopblob);
  a.verts = verts; a.F = F;
  a.tiles_per_cta = 54 / vsplit;
  static bool configured = false;
  if (!configured) {
    SEEME_CUDA(cudaFuncSetAttribute(smpl_skin_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM));
    configured = true;
  }
  ProfScope prof(prof_id - 1, s);
  smpl_skin_tc2_kernel<<<dim3(groups, vsplit), S2_THREADS, S2_SMEM, s>>>(maps, a);
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

// basis [SK][3][VP] fp32 -> bf16 (hi, lo) [3*VP][256], row = c*VP + v, zero padded in k
__global__ void smpl_basis_pack_kernel(const float* __restrict__ basis, int SK, __nv_bfloat16* __restrict__ bh,
                                       __nv_bfloat16* __restrict__ bl) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)3 * ST_VP * ST_KP) return;
  const int k = (int)(i % ST_KP);
  const size_t row = i / ST_KP;
  const int c = (int)(row / ST_VP), v = (int)(row % ST_VP);
  const float x = k < SK ? basis[((size_t)k * 3 + c) * ST_VP + v] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  bh[i] = h;
  bl[i] = __float2bfloat16_rn(x - __bfloat162float(h));
}

// fp16 flavour for smpl_skin_tc2_kernel (same [3*VP][256] layout, IEEE half bits in the 2-byte buffers)
__global__ void smpl_basis_pack_f16_kernel(const float* __restrict__ basis, int SK, __half* __restrict__ bh, __half* __restrict__ bl) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)3 * ST_VP * ST_KP) return;
  const int k = (int)(i % ST_KP);
  const size_t row = i / ST_KP;
  const int c = (int)(row / ST_VP), v = (int)(row % ST_VP);
  const float x = k < SK ? basis[((size_t)k * 3 + c) * ST_VP + v] : 0.f;
  const __half h = __float2half_rn(x);
  bh[i] = h;
  bl[i] = __float2half_rn(x - __half2float(h));
}
__global__ void smpl_coef_split_f16_kernel(const float* __restrict__ coef, int ld, int F, int n, __half* __restrict__ ch, __half* __restrict__ cl) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)F * n) return;
  const int f = (int)(i / n), k = (int)(i % n);
  const float x = coef[(size_t)f * ld + k];
  const __half h = __float2half_rn(x);
  ch[(size_t)f * ST_KP + k] = h;
  cl[(size_t)f * ST_KP + k] = __float2half_rn(x - __half2float(h));
}

int smpl_tc_pack_basis_f16(const float* basis, int SK, void* bh, void* bl) {
  const size_t n = (size_t)3 * ST_VP * ST_KP;
  smpl_basis_pack_f16_kernel<<<(unsigned)((n + 255) / 256), 256>>>(basis, SK, reinterpret_cast<__half*>(bh), reinterpret_cast<__half*>(bl));
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

size_t smpl_tc_basis_elems() { return (size_t)3 * ST_VP * ST_KP; }

int smpl_tc_pack_basis(const float* basis, int SK, void* bh, void* bl) {
  const size_t n = smpl_tc_basis_elems();
  smpl_basis_pack_kernel<<<(unsigned)((n + 255) / 256), 256>>>(basis, SK, reinterpret_cast<__nv_bfloat16*>(bh),
                                                                reinterpret_cast<__nv_bfloat16*>(bl));
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

int smpl_skin_tc(const void* bh, const void* bl, const float* coef, int ld_coef, int n_coef, void* ch, void* cl, const float* A,
                 const float* vt, const float* w4, const unsigned char* i4, int F, float* verts, int prof_id, cudaStream_t s) {
#ifndef SEEME_EXPERIMENTAL
  (void)bh; (void)bl; (void)coef; (void)ld_coef; (void)n_coef; (void)ch; (void)cl; (void)A; (void)vt; (void)w4; (void)i4; (void)F; (void)verts;
  (void)prof_id; (void)s;
  SEEME_REQUIRE(false, SEEME_EINVAL, "smpl_skin_tc_kernel (v1) is compiled in experimental builds only (-DSEEME_EXPERIMENTAL)");
#else
  // coef [F, n_coef] fp32 -> bf16 (hi, lo) [F, 256]; columns [n_coef, 256) were zeroed at create
  SEEME_TRY(to_bf16_split(coef, ld_coef, F, n_coef, reinterpret_cast<__nv_bfloat16*>(ch), reinterpret_cast<__nv_bfloat16*>(cl), ST_KP, 0, s));
  StMaps maps;
  memset(&maps, 0, sizeof(maps));
  SEEME_TRY(umma_tensor_map_bf16(&maps.bh, bh, 3 * ST_VP, ST_KP, ST_KP, 128));
  SEEME_TRY(umma_tensor_map_bf16(&maps.bl, bl, 3 * ST_VP, ST_KP, ST_KP, 128));
  SEEME_TRY(umma_tensor_map_bf16(&maps.ch, ch, F, ST_KP, ST_KP, ST_NF));
  SEEME_TRY(umma_tensor_map_bf16(&maps.cl, cl, F, ST_KP, ST_KP, ST_NF));
  static int csz = -1;
  if (csz < 0) {
    const char* e = seeme_exp_env("SEEME_SMPL_CLUSTER");
    csz = e ? atoi(e) : 1;   // measured on B200: multicast halves the L2 reads but not the time (the SM-inbound side bounds it)
    if (csz != 1 && csz != 2 && csz != 4) csz = 1;
  }
  // frame groups padded to a multiple of the cluster size: a ghost group (f0 >= F) only consumes the chunk stream
  const int groups = ((F + ST_NF - 1) / ST_NF + csz - 1) / csz * csz;
  // split the 54 vertex tiles so that the grid has a few waves of CTAs; the split must divide 54
  static const int divs[8] = {1, 2, 3, 6, 9, 18, 27, 54};
  int vsplit = 54;
  for (int i = 0; i < 8; ++i)
    if (groups * divs[i] >= 4 * NUM_SMS) { vsplit = divs[i]; break; }
  StArgs a;
  a.A = A; a.vt = vt; a.w4 = w4; a.i4 = i4; a.verts = verts; a.F = F;
  a.tiles_per_cta = 54 / vsplit;
  static bool configured = false;
  if (!configured) {
    SEEME_CUDA(cudaFuncSetAttribute(smpl_skin_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
    SEEME_CUDA(cudaFuncSetAttribute(smpl_skin_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
    SEEME_CUDA(cudaFuncSetAttribute(smpl_skin_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
    configured = true;
  }
  ProfScope prof(prof_id - 1, s);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(groups, vsplit); cfg.blockDim = dim3(ST_THREADS); cfg.dynamicSmemBytes = ST_SMEM; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = csz > 1 ? 1 : 0;
  if (csz == 1) SEEME_CUDA(cudaLaunchKernelEx(&cfg, smpl_skin_tc_kernel<1>, maps, a));
  else if (csz == 2) SEEME_CUDA(cudaLaunchKernelEx(&cfg, smpl_skin_tc_kernel<2>, maps, a));
  else SEEME_CUDA(cudaLaunchKernelEx(&cfg, smpl_skin_tc_kernel<4>, maps, a));
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
#endif
}

}  // namespace seeme
