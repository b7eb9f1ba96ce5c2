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
    const char* e = getenv("SEEME_SMPL_CLUSTER");
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
}

}  // namespace seeme
