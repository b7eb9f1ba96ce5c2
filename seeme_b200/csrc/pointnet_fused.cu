// Fused ResnetBlockFC for the scene encoder (EgoHMR/models/respointnet.py:88-97 applied to [net | pooled],
// :36-48) -- one persistent tcgen05 kernel per residual block i >= 1, fp16 operands, fp32 accumulation.
//
// Per 128-point tile (points of ONE sample; the pooled half of the reference's concat is a per-sample bias,
// see pointnet.cu) the whole block runs on chip:
//
//     X   = net tile [128, 256] fp16                       TMA load (3-D map [256, N, B], OOB rows = 0)
//     S   : OUT  = X . Ws^T                                 tcgen05.mma, A and B from shared memory
//     relu: X   <- relu(X) in place                         epilogue warps, K-chunk by K-chunk behind S
//     G1  : H    = relu(X) . W0^T                           two N-halves, each committed separately
//     epiH: H16 = fp16(relu(H + c0[b]))                     TMEM -> registers -> TMEM (in place, packed 2/column)
//     G2  : OUT += H16 . W1^T                               tcgen05.mma with the A operand read from TMEM
//     epiOUT: out = OUT + cs[b]; per-sample column max (pooling); fp16 tile staged in the (now free) X buffer
//             and written back with TMA stores
//
// The activations never leave the SM between the three GEMMs; HBM sees 512 B in + 512 B out per point and block.
// The weights of a block (3 x 256x256 fp16 = 384 KB) are streamed from L2 as 24 chunks of [128 n, 64 k] per tile
// through a 6-deep ring.  TMEM (512 columns) holds two 256-column regions that alternate between the OUT
// accumulator of tile j and the H accumulator of tile j+1, so tile j's output epilogue overlaps tile j+1's
// shortcut GEMM.
//
//   warp 0      weight-ring TMA producer
//   warp 1      TMEM allocator + the single MMA-issuing thread
//   warps 2..5  H group (TMEM lane quarter = warp % 4): in-place relu of X behind the shortcut GEMM, H epilogue
//   warps 6..13 O group (lane quarter x column half): OUT epilogue (bias, pooled column max via a transpose-reduce
//               butterfly, fp16 staging), TMA stores and the X-tile TMA loads (one elected thread)
// The two groups run concurrently, so tile j's OUT epilogue overlaps tile j+1's shortcut GEMM, relu pass and G1.
#include "umma.cuh"
#include "pointnet_fused.cuh"
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

namespace seeme {

// Optional event trace (build with -DPF_TRACE): CTA 0 records clock64() at protocol events of its first tiles into
// pf_trace[tile][slot]; read back with seeme_pf_trace_read (tools/pf_trace.py).
#ifdef PF_TRACE
constexpr int PF_TRACE_TILES = 24, PF_TRACE_SLOTS = 64;
__device__ long long pf_trace[PF_TRACE_TILES * PF_TRACE_SLOTS];
#define PF_TR(tile, slot)                                                                                 \
  do {                                                                                                    \
    if (blockIdx.x == 0 && (tile) < PF_TRACE_TILES) pf_trace[(tile) * PF_TRACE_SLOTS + (slot)] = clock64(); \
  } while (0)
#else
#define PF_TR(tile, slot) do { } while (0)
#endif

constexpr int PF_NST = 6;                           // weight ring depth
constexpr int PF_CHUNK = 128 * 128;                 // [128 rows x 64 fp16], SWIZZLE_128B
constexpr int PF_XBUF = 4 * PF_CHUNK;               // one activation tile [128 x 256] fp16
constexpr int PF_THREADS = 448;            // 2 control warps + 4 H-group warps + 8 O-group warps
constexpr int PF_WCHUNKS = 24;                      // weight chunks per tile: S 8, G1 8, G2 8
constexpr int PF_SMEM = 2 * PF_XBUF + PF_NST * PF_CHUNK + 1024;     // pair kernel (manual 1 KB alignment slack)
constexpr int PF_SMEM_BLK = 2 * PF_XBUF + PF_NST * PF_CHUNK;       // pointnet_block_kernel (1 KB of static bias copy instead)

struct PfMaps { CUtensorMap xin, xout, w; };
struct PfArgs {
  int n_points, tiles_per_sample, n_tiles;
  const float* bias_h;   // [samples, 256]  c0: added to H before its relu
  const float* bias_o;   // [samples, 256]  cs: added to OUT
  unsigned* colmax;      // [samples, 256]  order-preserving uint, zeroed by the caller
  int store_out;
  const uint8_t* wblob;  // 24 pre-swizzled 16 KB weight chunks (byte image of the shared-memory tiles)
  int debug_skip;        // diagnostics only (SEEME_PF_DEBUG_SKIP, wrong results): 1 = no relu pass, 2 = no pooled max, 4 = no H epilogue math
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// plain (non-tensor) bulk copy of `bytes` contiguous bytes global -> shared, completion on an mbarrier.  The weight chunks
// are stored in global memory as the byte image of their SWIZZLE_128B shared-memory tile, so one 16 KB burst replaces a
// 128-row tiled TMA box (128 separate 128-byte requests).
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void pf_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pf_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void pf_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T : the A tile sits in TMEM as 128 lanes x (K/2) columns, two fp16 per column
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float warp_max_f32(float v) {
  float m;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  return m;
}
__device__ __forceinline__ uint32_t pf_pack(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// one lane of a converged warp (the same lane every time)
__device__ __forceinline__ bool pf_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// advance the 14-bit start-address field of a shared-memory matrix descriptor by off16 (units of 16 bytes)
__device__ __forceinline__ uint64_t pf_desc_add(uint64_t d, uint32_t off16) {
  return (d & 0xffffffff00000000ull) | (uint64_t)((uint32_t)d + off16);
}
__device__ __forceinline__ uint32_t pf_sw128(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

// per-warp column max of a 32x32 block held one row per lane: returns the max of column `lane`
// (transpose-reduce butterfly: 31 shuffles; v is destroyed)
__device__ __forceinline__ float pf_colmax32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float keep = up ? v[i + off] : v[i];
      const float send = up ? v[i] : v[i + off];
      v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
    }
  }
  return v[0];
}

// same for 32 columns held as 16 packed fp16 pairs (columns 2i, 2i+1 in v[i]): 16 shuffles + 16 HMNMX2 instead of 31 + 31.
// max and fp16 rounding commute (rounding is monotonic), so this is fp16(max over the fp32 values).  Returns the max of
// column `lane` as fp32; v is destroyed.
__device__ __forceinline__ float pf_colmax32_h2(uint32_t (&v)[16], int lane) {
#pragma unroll
  for (int off = 16; off >= 2; off >>= 1) {
    const bool up = (lane & off) != 0;
    const int n = off >> 1;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const uint32_t keep = up ? v[i + n] : v[i];
      const uint32_t send = up ? v[i] : v[i + n];
      const uint32_t got = __shfl_xor_sync(0xffffffffu, send, off);
      const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&keep), *reinterpret_cast<const __half2*>(&got));
      v[i] = *reinterpret_cast<const uint32_t*>(&m);
    }
  }
  const uint32_t got = __shfl_xor_sync(0xffffffffu, v[0], 1);
  const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&v[0]), *reinterpret_cast<const __half2*>(&got));
  return (lane & 1) ? __high2float(m) : __low2float(m);
}

// Lean barrier primitives for the statically scheduled kernels below: 32-bit shared addresses computed once, a wait loop
// the compiler may not unroll (the generic mbar_wait is unrolled x4 with its printf path at every call site, ~150 SASS
// instructions per wait), the timeout path out of line.
__device__ __noinline__ void pf_wait_timeout(uint32_t addr) {
  printf("seeme_b200: mbarrier wait timed out (block %d thread %d barrier 0x%x)\n", blockIdx.x, threadIdx.x, addr);
  __trap();
}
__device__ __forceinline__ void pf_wait(uint32_t addr, uint32_t parity) {
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  pf_wait_timeout(addr);
}
__device__ __forceinline__ void pf_arrive(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void pf_arrive_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pf_commit(uint32_t addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void pf_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}


__device__ __forceinline__ uint32_t pf_relu_pack(uint32_t a, uint32_t b) {
  uint32_t d;      // {lo = fp16(max(a, 0)), hi = fp16(max(b, 0))}: conversion, relu and packing in ONE instruction
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(b)), "f"(__uint_as_float(a)));
  return d;
}

// H_TMEM: G2 reads its A operand (H) from tensor memory; false = H is written over relu(X) in shared memory
template <bool H_TMEM>
__global__ void __launch_bounds__(PF_THREADS, 1) pointnet_block_kernel(const __grid_constant__ PfMaps tm, const PfArgs a) {
  // no manual alignment slack here: the 1 KB goes to the per-sample bias copy below (the static + dynamic total is exactly
  // the 227 KB limit); the declared alignment of the dynamic array is checked instead
  extern __shared__ __align__(1024) uint8_t pf_smem_raw[];
  uint8_t* smem = pf_smem_raw;
  if ((smem_u32(pf_smem_raw) & 1023u) != 0u) __trap();
  uint8_t* xbuf = smem;
  uint8_t* wring = smem + 2 * PF_XBUF;
  // Static schedule: 12 steps per tile, step s consumes the 32 KB weight pair s % 3 (chunks 2s, 2s+1 of the blob):
  //   s = 0..3   S(kc)        step_full[s]: the pair's bytes
  //   s = 4..7   G1(kc)       N = 256 (A is read once; N = 128 halves re-read it and ran at the shared-memory bandwidth);
  //                           step_full[s] also collects the 4 relu-warp arrivals for X chunk kc
  //   s = 8..11  G2(kc)       step_full[s] also collects the 4 H-epilogue warps' arrivals
  // one phase per tile and barrier (parity = tile & 1); w_empty[s % 3] is committed once per step (4 ring rounds per tile, so
  // its parities are compile-time constants too)
  __shared__ __align__(8) uint64_t step_full[12], w_empty[3], x_full[2], s_done[2][4], h_full[2][2], out_full[2], out_drained[2], staged[2];
  __shared__ uint32_t tmem_slot;
  // c0[b] / cs[b] of the current sample: the L1 next to 225 KB of shared memory is tiny, and both epilogues waited ~800 cycles
  // per 32 columns on __ldg of them.  (The pooled column maxima are kept in registers across the tiles of a sample instead of
  // a shared-memory array, which is what makes room for the second vector.)
  __shared__ __align__(16) float bias_h_s[256];
  __shared__ __align__(16) float bias_o_s[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = a.n_tiles / (int)gridDim.x, rem = a.n_tiles % (int)gridDim.x;
  const int t_begin = (int)blockIdx.x * per + ((int)blockIdx.x < rem ? (int)blockIdx.x : rem);
  const int nt = per + ((int)blockIdx.x < rem ? 1 : 0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm.xin);
    tma_prefetch_desc(&tm.w);
    tma_prefetch_desc(&tm.xout);
    for (int i = 0; i < 12; ++i) mbar_init(&step_full[i], i < 4 ? 1 : 5);
    for (int i = 0; i < 3; ++i) mbar_init(&w_empty[i], 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&x_full[b], 1);
      mbar_init(&out_full[b], 1);
      mbar_init(&out_drained[b], 8);
      mbar_init(&staged[b], 8);
      for (int k = 0; k < 4; ++k) mbar_init(&s_done[b][k], 1);
      mbar_init(&h_full[b][0], 1);
      mbar_init(&h_full[b][1], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  const uint32_t sf0 = smem_u32(step_full), we0 = smem_u32(w_empty);
  if (warp == 0) {
    // ---- weight-ring producer: one 32 KB bulk copy per step ---------------------------------------------------
    if (lane == 0) {
      const uint32_t wr = smem_u32(wring);
      for (int j = 0; j < nt; ++j) {
#pragma unroll
        for (int s = 0; s < 12; ++s) {
          pf_wait(we0 + (uint32_t)(s % 3) * 8u, (uint32_t)((s / 3) & 1) ^ 1u);
          pf_arrive_tx(sf0 + (uint32_t)s * 8u, 2 * PF_CHUNK);
          if (s >= 4 && s < 8) {      // G1(kc): the blob keeps W0 n-half major (chunks 8 + nh * 4 + kc); both halves of K-chunk kc form the pair
            pf_bulk_load(wr + (uint32_t)(s % 3) * 2 * PF_CHUNK, a.wblob + (size_t)(8 + s - 4) * PF_CHUNK, PF_CHUNK, sf0 + (uint32_t)s * 8u);
            pf_bulk_load(wr + (uint32_t)(s % 3) * 2 * PF_CHUNK + PF_CHUNK, a.wblob + (size_t)(12 + s - 4) * PF_CHUNK, PF_CHUNK, sf0 + (uint32_t)s * 8u);
          } else {
            // G2 runs its K-chunks in the order 0, 2, 1, 3: the H epilogue converts both column halves in parallel, so
            // chunks 0 and 2 are ready together (then 1 and 3)
            const int pair = s < 8 ? s : 8 + (((s - 8) & 1) << 1 | ((s - 8) >> 1));
            pf_bulk_load(wr + (uint32_t)(s % 3) * 2 * PF_CHUNK, a.wblob + (size_t)pair * 2 * PF_CHUNK, 2 * PF_CHUNK, sf0 + (uint32_t)s * 8u);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: the whole warp runs the (uniform) control flow, one elected lane issues tcgen05.mma / commit ----
    constexpr uint32_t idesc256 = umma_idesc_f16(256);
    const uint64_t wdesc0 = umma_desc_k128(smem_u32(wring));
    for (int j = 0; j < nt; ++j) {
      const int b = j & 1;
      const uint32_t pj = (uint32_t)j & 1u, p2 = (uint32_t)(j >> 1) & 1u;
      // TMEM: OUT(j) lives in region j & 1, H(j) in the other one, i.e. where OUT(j-1) was: tile j's shortcut GEMM
      // overlaps tile j-1's output epilogue
      const uint32_t Ra = tmem_base + (uint32_t)b * 256u, Rb = tmem_base + (uint32_t)(b ^ 1) * 256u;
      const uint64_t xdesc = umma_desc_k128(smem_u32(xbuf + b * PF_XBUF));
      if (lane == 0) PF_TR(j, 0);
      mbar_wait(&x_full[b], p2);
      if (lane == 0) PF_TR(j, 1);
      // S: OUT = X . Ws^T  -- per K-chunk one N = 256 MMA group over the two chunks (n-halves) of the pair
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        constexpr int S0 = 0;
        pf_wait(sf0 + (uint32_t)(S0 + kc) * 8u, pj);
        tc_fence_after();
        if (pf_elect_one()) {
          PF_TR(j, 2 + kc);
          const uint64_t wd = pf_desc_add(wdesc0, (uint32_t)((S0 + kc) % 3) * (2 * PF_CHUNK >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(Ra, pf_desc_add(xdesc, kc * (PF_CHUNK >> 4) + ks * 2), pf_desc_add(wd, ks * 2), idesc256, (kc | ks) != 0);
          pf_commit(we0 + (uint32_t)((S0 + kc) % 3) * 8u);
          umma_commit(&s_done[b][kc]);
        }
        __syncwarp();
      }
      // the H region of this tile was the OUT region of the previous one: its epilogue must have drained it
      if (j > 0) mbar_wait(&out_drained[b ^ 1], (uint32_t)((j - 1) >> 1) & 1u);
      if (lane == 0) PF_TR(j, 6);
      // G1: H = relu(X) . W0^T  (N = 256 per K-chunk)
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        const int s = 4 + kc;
        pf_wait(sf0 + (uint32_t)s * 8u, pj);
        tc_fence_after();
        if (pf_elect_one()) {
          PF_TR(j, 3 + s);
          const uint64_t wd = pf_desc_add(wdesc0, (uint32_t)(s % 3) * (2 * PF_CHUNK >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(Rb, pf_desc_add(xdesc, kc * (PF_CHUNK >> 4) + ks * 2), pf_desc_add(wd, ks * 2), idesc256, (kc | ks) != 0);
          pf_commit(we0 + (uint32_t)(s % 3) * 8u);
          if (kc == 3) umma_commit(&h_full[b][0]);
        }
        __syncwarp();
      }
      // G2: OUT += H16 . W1^T  (N = 256 per K-chunk)
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        const int s = 8 + i4;
        const int kc = ((i4 & 1) << 1) | (i4 >> 1);      // 0, 2, 1, 3
        pf_wait(sf0 + (uint32_t)s * 8u, pj);
        tc_fence_after();
        if (pf_elect_one()) {
          PF_TR(j, 3 + s);
          const uint64_t wd = pf_desc_add(wdesc0, (uint32_t)(s % 3) * (2 * PF_CHUNK >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (H_TMEM)
              umma_f16_ts(Ra, Rb + (uint32_t)((kc >> 1) * 128 + (kc & 1) * 32 + ks * 8), pf_desc_add(wd, ks * 2), idesc256, 1);
            else
              umma_bf16(Ra, pf_desc_add(xdesc, kc * (PF_CHUNK >> 4) + ks * 2), pf_desc_add(wd, ks * 2), idesc256, 1);
          }
          pf_commit(we0 + (uint32_t)(s % 3) * 8u);
          if (i4 == 3) umma_commit(&out_full[b]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ---- H group (4 warps): in-place relu of X, chunk by chunk behind the shortcut GEMM ------------------------
    const int th = (int)threadIdx.x - 64;           // 0..127
    // Thread 0 of this group also writes the staged output tiles back and prefetches the X tiles: when the output group's
    // elected thread did it, its wait for the TMA store to read the staging buffer delayed that warp's H epilogue of the
    // NEXT tile, i.e. the first G2 step (event trace: ~700 cycles per tile)
    auto store_and_prefetch = [&](int jp) {          // tile jp has been staged in X buffer jp & 1
      const int bp = jp & 1;
      const int t = t_begin + jp;
      const int sample = t / a.tiles_per_sample, n0 = (t % a.tiles_per_sample) * 128;
      mbar_wait(&staged[bp], (uint32_t)(jp >> 1) & 1u);
      if (a.store_out) {
        for (int kc = 0; kc < 4; ++kc) tma_store_3d(&tm.xout, xbuf + bp * PF_XBUF + kc * PF_CHUNK, kc * 64, n0, sample);
        pf_store_commit();
        pf_store_wait_read();
      }
      if (jp + 2 < nt) {
        const int t2 = t_begin + jp + 2;
        const int s2 = t2 / a.tiles_per_sample, m0 = (t2 % a.tiles_per_sample) * 128;
        mbar_arrive_expect_tx(&x_full[bp], PF_XBUF);
        for (int kc = 0; kc < 4; ++kc) tma_load_3d(xbuf + bp * PF_XBUF + kc * PF_CHUNK, &tm.xin, &x_full[bp], kc * 64, m0, s2);
      }
    };
    for (int j = 0; j < nt; ++j) {
      const int b = j & 1;
      const uint32_t p2 = (uint32_t)(j >> 1) & 1u;
      uint8_t* xb = xbuf + b * PF_XBUF;
      for (int kc = 0; kc < 4; ++kc) {
        mbar_wait(&s_done[b][kc], p2);
        if (th == 0) PF_TR(j, 16 + 2 * kc);
        uint4* p = reinterpret_cast<uint4*>(xb + kc * PF_CHUNK) + th;
        const __half2 z = __float2half2_rn(0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#ifdef SEEME_EXPERIMENTAL
          if (a.debug_skip & 1) break;
#endif
          uint4 v = p[i * 128];
          __half2* h = reinterpret_cast<__half2*>(&v);
          h[0] = __hmax2(h[0], z); h[1] = __hmax2(h[1], z); h[2] = __hmax2(h[2], z); h[3] = __hmax2(h[3], z);
          p[i * 128] = v;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) pf_arrive(sf0 + (uint32_t)(4 + kc) * 8u);
        if (th == 0) PF_TR(j, 17 + 2 * kc);
      }
      if (th == 0 && j > 0) store_and_prefetch(j - 1);
    }
    if (th == 0 && nt > 0) store_and_prefetch(nt - 1);
  } else {
    // ---- O group (8 warps: lane quarter x column half): OUT epilogue, pooled column max, TMA stores, X-tile loads ---------
    const int te = (int)threadIdx.x - 192;         // 0..255
    const int q = warp & 3;
    const int hsel = (warp - 6) >> 2;
    const int row = q * 32 + lane;
    const bool elected = te == 0;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    auto load_x = [&](int j) {
      const int t = t_begin + j, b = j & 1;
      const int sample = t / a.tiles_per_sample, n0 = (t % a.tiles_per_sample) * 128;
      mbar_arrive_expect_tx(&x_full[b], PF_XBUF);
      for (int kc = 0; kc < 4; ++kc) tma_load_3d(xbuf + b * PF_XBUF + kc * PF_CHUNK, &tm.xin, &x_full[b], kc * 64, n0, sample);
    };
    // pooled column maxima of this warp's 32 rows, columns hsel * 128 + g * 32 + lane, accumulated over the tiles of a sample
    float cmax[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    auto flush_colmax = [&](int sample) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (cmax[g] > -INFINITY) atomicMax(a.colmax + (size_t)sample * 256 + hsel * 128 + g * 32 + lane, f2ord(cmax[g]));
        cmax[g] = -INFINITY;
      }
    };
    if (elected) {
      if (nt > 0) load_x(0);
      if (nt > 1) load_x(1);
    }
    int cur_sample = -1;
    for (int j = 0; j < nt; ++j) {
      const int b = j & 1;
      const uint32_t p2 = (uint32_t)(j >> 1) & 1u;
      const int t = t_begin + j;
      const int sample = t / a.tiles_per_sample, n0 = (t % a.tiles_per_sample) * 128;
      uint8_t* xb = xbuf + b * PF_XBUF;
      if (sample != cur_sample) {
        if (cur_sample >= 0) flush_colmax(cur_sample);
        pf_epi_sync();                                       // nobody reads the previous sample's biases any more
        bias_h_s[te] = __ldg(a.bias_h + (size_t)sample * 256 + te);
        bias_o_s[te] = __ldg(a.bias_o + (size_t)sample * 256 + te);
        pf_epi_sync();
        cur_sample = sample;
      }
      // H epilogue of THIS tile (both column halves in parallel on the 8 warps of this group; the 4 relu warps stay on the
      // relu pass): H16 = fp16(relu(H + c0[b])), in place in TMEM (or over relu(X) in shared memory)
      {
        const float* bh = bias_h_s + hsel * 128;
        const uint32_t thh = tmem_base + (uint32_t)(b ^ 1) * 256u + lane_off + (uint32_t)hsel * 128u;
        mbar_wait(&h_full[b][0], p2);
        if (elected) PF_TR(j, 24);
        tc_fence_after();
        uint32_t raw[2][32];
        tmem_ld32(thh, raw[0]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float4 bv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) bv[i] = *(reinterpret_cast<const float4*>(bh + g * 32) + i);
          tmem_ld_wait();
          if (g < 3) tmem_ld32(thh + (g + 1) * 32, raw[(g + 1) & 1]);
          const uint32_t* r = raw[g & 1];
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            pk[2 * i] = pf_relu_pack(__float_as_uint(__uint_as_float(r[4 * i]) + bv[i].x), __float_as_uint(__uint_as_float(r[4 * i + 1]) + bv[i].y));
            pk[2 * i + 1] = pf_relu_pack(__float_as_uint(__uint_as_float(r[4 * i + 2]) + bv[i].z), __float_as_uint(__uint_as_float(r[4 * i + 3]) + bv[i].w));
          }
          if (H_TMEM) {
            tmem_st16(thh + g * 16, pk);         // in place: these 16 columns were read in this or an earlier group
          } else {
            uint8_t* ct = xb + (hsel * 2 + (g >> 1)) * PF_CHUNK;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              *reinterpret_cast<uint4*>(ct + pf_sw128(row, (g & 1) * 4 + jj)) = make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
          }
          if (g & 1) {
            if (H_TMEM) { tmem_st_wait(); tc_fence_before(); }
            else fence_proxy_async();
            __syncwarp();
            if (lane == 0) pf_arrive(sf0 + (uint32_t)(8 + (((g >> 1) << 1) | hsel)) * 8u);      // K-chunk hsel * 2 + (g >> 1) is step 8 + 2 (g >> 1) + hsel
          }
        }
        if (elected) PF_TR(j, 25);
      }
      const uint32_t to = tmem_base + (uint32_t)b * 256u + (uint32_t)hsel * 128u + lane_off;
      const bool valid = n0 + row < a.n_points;
      mbar_wait(&out_full[b], p2);
      if (elected) PF_TR(j, 28);
      tc_fence_after();
      // drain first: the accumulator region is the next tile's H region, so G1(j+1) waits for these four loads.  Each
      // 32-column group is biased and packed to fp16 (16 words) straight away; staging and the pooled max (on the packed
      // values: max and fp16 rounding commute) come after the region has been released
      uint32_t pk[4][16];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t raw[32];
        tmem_ld32(to + g * 32, raw);
        float4 bv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bv[i] = *(reinterpret_cast<const float4*>(bias_o_s + hsel * 128 + g * 32) + i);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          pk[g][2 * i] = pf_pack(__uint_as_float(raw[4 * i]) + bv[i].x, __uint_as_float(raw[4 * i + 1]) + bv[i].y);
          pk[g][2 * i + 1] = pf_pack(__uint_as_float(raw[4 * i + 2]) + bv[i].z, __uint_as_float(raw[4 * i + 3]) + bv[i].w);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&out_drained[b]);
      if (elected) PF_TR(j, 29);
      if (a.store_out) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint8_t* ct = xb + (hsel * 2 + (g >> 1)) * PF_CHUNK;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            *reinterpret_cast<uint4*>(ct + pf_sw128(row, (g & 1) * 4 + jj)) = make_uint4(pk[g][4 * jj], pk[g][4 * jj + 1], pk[g][4 * jj + 2], pk[g][4 * jj + 3]);
        }
      }
#ifdef SEEME_EXPERIMENTAL
      if (!(a.debug_skip & 2))
#endif
      {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (!valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[g][i] = 0xfc00fc00u;      // -inf, -inf
          }
          cmax[g] = fmaxf(cmax[g], pf_colmax32_h2(pk[g], lane));
        }
      }
      // the staged tile is complete for this warp: the relu group's thread 0 stores it and prefetches X tile j + 2
      if (a.store_out) fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&staged[b]);
      if (elected) PF_TR(j, 30);
    }
    if (cur_sample >= 0) flush_colmax(cur_sample);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

#ifdef SEEME_EXPERIMENTAL   // retired variant (CUDA-core generator of block 0): experimental builds only
// ---------------------------------------------------------------------------------------------------------------------
// Block 0 (+ fc_pos): the block's input relu(fc_pos(p)) [128, 512] is GENERATED on chip from the xyz coordinates (K = 3,
// CUDA cores) chunk by chunk into a 4-slot shared-memory ring -- it never exists in HBM -- and its shortcut, an affine
// map of p (rank 3, see pointnet.cu), is evaluated in the output epilogue.
//     gen : A[:, kc] = fp16(relu(Wp[kc] p + bp[kc]))          H group, one K-chunk ahead of the MMAs
//     G1  : H   = A . W0^T  (K = 512)                          8 N=256 MMA groups
//     epiH: H16 = fp16(relu(H + b0)) in place in TMEM          H group
//     G2  : OUT = H16 . W1^T                                   A operand from TMEM
//     epiOUT: out = OUT + cst0 + Pf p; pooled column max; fp16 tile -> staging -> TMA store      O group
// TMEM: H in columns [0,256), OUT in [256,512); G1 of tile j+1 overlaps the OUT epilogue of tile j.
constexpr int P0_ASLOTS = 4;
constexpr int P0_SMEM = P0_ASLOTS * PF_CHUNK + PF_XBUF + PF_NST * PF_CHUNK + 1024;

struct P0Args {
  int n_points, tiles_per_sample, n_tiles;
  const float* xyz;      // [samples, n_points, 3]
  const float4* wpb;     // [512] (Wp[c][0..2], bp[c])
  const float* b0;       // [256]
  const float* cst0;     // [256]  Ws bp + b1
  const float4* pfold;   // [256]  (Ws Wp)[c][0..2], 0
  unsigned* colmax;
  const uint8_t* wblob;  // 24 pre-swizzled 16 KB weight chunks
};

__global__ void __launch_bounds__(PF_THREADS, 1) pointnet_block0_kernel(const __grid_constant__ PfMaps tm, const P0Args a) {
  extern __shared__ __align__(1024) uint8_t pf_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pf_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* aring = smem;                                  // 4 x 16 KB generated A chunks
  uint8_t* stage = smem + P0_ASLOTS * PF_CHUNK;           // 64 KB output staging
  uint8_t* wring = stage + PF_XBUF;
  __shared__ __align__(8) uint64_t w_full[PF_NST], w_empty[PF_NST], a_ready[P0_ASLOTS], a_free[P0_ASLOTS], h_full, h_ready[4], h_free,
      out_full, out_drained;
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned colmax_s[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = a.n_tiles / (int)gridDim.x, rem = a.n_tiles % (int)gridDim.x;
  const int t_begin = (int)blockIdx.x * per + ((int)blockIdx.x < rem ? (int)blockIdx.x : rem);
  const int nt = per + ((int)blockIdx.x < rem ? 1 : 0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm.w);
    tma_prefetch_desc(&tm.xout);
    for (int i = 0; i < PF_NST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < P0_ASLOTS; ++i) { mbar_init(&a_ready[i], 4); mbar_init(&a_free[i], 1); }
    for (int i = 0; i < 4; ++i) mbar_init(&h_ready[i], 4);
    mbar_init(&h_full, 1);
    mbar_init(&h_free, 1);
    mbar_init(&out_full, 1);
    mbar_init(&out_drained, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  if (threadIdx.x < 256) colmax_s[threadIdx.x] = 0u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t RH = tmem_base, RO = tmem_base + 256u;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t st = 0, ph = 1;
      for (int j = 0; j < nt; ++j) {
        for (int i = 0; i < PF_WCHUNKS; ++i) {
          mbar_wait(&w_empty[st], ph);
          mbar_arrive_expect_tx(&w_full[st], PF_CHUNK);
          bulk_load(wring + st * PF_CHUNK, a.wblob + (size_t)i * PF_CHUNK, PF_CHUNK, &w_full[st]);
          if (++st == PF_NST) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc256 = umma_idesc_f16(256);
    const uint64_t wdesc0 = umma_desc_k128(smem_u32(wring));
    const uint64_t adesc0 = umma_desc_k128(smem_u32(aring));
    uint32_t st = 0, wph = 0;
    for (int j = 0; j < nt; ++j) {
      const uint32_t pj = (uint32_t)j & 1u;
      // G1 writes the H region: G2 of the previous tile (which read H16 from it) is ordered before by the MMA pipe
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        const int as = kc & 3;
        mbar_wait(&a_ready[as], (uint32_t)(kc >> 2) & 1u);          // 8 chunks per tile = exactly 2 ring rounds per tile
        mbar_wait(&w_full[st], wph);
        mbar_wait(&w_full[st + 1], wph);
        tc_fence_after();
        if (pf_elect_one()) {
          const uint64_t wd = pf_desc_add(wdesc0, st * (PF_CHUNK >> 4));
          const uint64_t ad = pf_desc_add(adesc0, as * (PF_CHUNK >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(RH, pf_desc_add(ad, ks * 2), pf_desc_add(wd, ks * 2), idesc256, (kc | ks) != 0);
          umma_commit(&w_empty[st]);
          umma_commit(&w_empty[st + 1]);
          umma_commit(&a_free[as]);
          if (kc == 7) umma_commit(&h_full);
        }
        __syncwarp();
        st += 2;
        if (st == PF_NST) { st = 0; wph ^= 1u; }
      }
      // G2 writes the OUT region: the previous tile's output epilogue must have drained it
      if (j > 0) mbar_wait(&out_drained, (uint32_t)(j - 1) & 1u);
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        mbar_wait(&h_ready[kc], pj);
        mbar_wait(&w_full[st], wph);
        mbar_wait(&w_full[st + 1], wph);
        tc_fence_after();
        if (pf_elect_one()) {
          const uint64_t wd = pf_desc_add(wdesc0, st * (PF_CHUNK >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16_ts(RO, RH + (uint32_t)((kc >> 1) * 128 + (kc & 1) * 32 + ks * 8), pf_desc_add(wd, ks * 2), idesc256, (kc | ks) != 0);
          umma_commit(&w_empty[st]);
          umma_commit(&w_empty[st + 1]);
          if (kc == 3) { umma_commit(&out_full); umma_commit(&h_free); }
        }
        __syncwarp();
        st += 2;
        if (st == PF_NST) { st = 0; wph ^= 1u; }
      }
    }
  } else if (warp < 6) {
    // ---- H group: thread = point.  Generates the A chunks, then the H epilogue ---------------------------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    for (int j = 0; j < nt; ++j) {
      const uint32_t pj = (uint32_t)j & 1u;
      const int t = t_begin + j;
      const int sample = t / a.tiles_per_sample, n0 = (t % a.tiles_per_sample) * 128;
      float px = 0.f, py = 0.f, pz = 0.f;
      if (n0 + row < a.n_points) {
        const float* pp = a.xyz + ((size_t)sample * a.n_points + n0 + row) * 3;
        px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2);
      }
#pragma unroll 1
      for (int kc = 0; kc < 8; ++kc) {
        const int as = kc & 3;
        // slot reuse: chunk kc of this tile overwrites the chunk consumed 4 chunks earlier (2 ring rounds per tile)
        mbar_wait(&a_free[as], ((uint32_t)(kc >> 2) & 1u) ^ 1u);
        uint8_t* ct = aring + as * PF_CHUNK;
        const float4* wp = a.wpb + kc * 64;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 w0 = __ldg(wp + jj * 8 + 2 * i), w1 = __ldg(wp + jj * 8 + 2 * i + 1);
            const float v0 = fmaxf(fmaf(w0.z, pz, fmaf(w0.y, py, fmaf(w0.x, px, w0.w))), 0.f);
            const float v1 = fmaxf(fmaf(w1.z, pz, fmaf(w1.y, py, fmaf(w1.x, px, w1.w))), 0.f);
            pk[i] = pf_pack(v0, v1);
          }
          *reinterpret_cast<uint4*>(ct + pf_sw128(row, jj)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[as]);
      }
      // H epilogue (in place in TMEM)
      mbar_wait(&h_full, pj);
      tc_fence_after();
#pragma unroll 1
      for (int hsel = 0; hsel < 2; ++hsel) {
        const uint32_t thh = RH + lane_off + (uint32_t)hsel * 128u;
        uint32_t raw[2][32];
        tmem_ld32(thh, raw[0]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float4 bv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) bv[i] = __ldg(reinterpret_cast<const float4*>(a.b0 + hsel * 128 + g * 32) + i);
          tmem_ld_wait();
          if (g < 3) tmem_ld32(thh + (g + 1) * 32, raw[(g + 1) & 1]);
          const uint32_t* r = raw[g & 1];
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float f0 = fmaxf(__uint_as_float(r[4 * i]) + bv[i].x, 0.f), f1 = fmaxf(__uint_as_float(r[4 * i + 1]) + bv[i].y, 0.f);
            const float f2 = fmaxf(__uint_as_float(r[4 * i + 2]) + bv[i].z, 0.f), f3 = fmaxf(__uint_as_float(r[4 * i + 3]) + bv[i].w, 0.f);
            pk[2 * i] = pf_pack(f0, f1);
            pk[2 * i + 1] = pf_pack(f2, f3);
          }
          tmem_st16(thh + g * 16, pk);
          if (g & 1) {
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h_ready[hsel * 2 + (g >> 1)]);
          }
        }
      }
      // the next tile's G1 may only overwrite the H region after this tile's G2 has consumed H16: that is MMA-pipe
      // order; but this group's NEXT H epilogue reads the region only after h_full, so nothing to wait for here.
      (void)h_free;
    }
  } else {
    // ---- O group ---------------------------------------------------------------------------------------------------
    const int te = (int)threadIdx.x - 192;
    const int q = warp & 3;
    const int hsel = (warp - 6) >> 2;
    const int row = q * 32 + lane;
    const bool elected = te == 0;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    auto flush_colmax = [&](int sample) {
      pf_epi_sync();
      const unsigned v = colmax_s[te];
      if (v) atomicMax(a.colmax + (size_t)sample * 256 + te, v);
      colmax_s[te] = 0u;
      pf_epi_sync();
    };
    int cur_sample = -1;
    for (int j = 0; j < nt; ++j) {
      const uint32_t pj = (uint32_t)j & 1u;
      const int t = t_begin + j;
      const int sample = t / a.tiles_per_sample, n0 = (t % a.tiles_per_sample) * 128;
      if (sample != cur_sample) {
        if (cur_sample >= 0) flush_colmax(cur_sample);
        cur_sample = sample;
      }
      const bool valid = n0 + row < a.n_points;
      float px = 0.f, py = 0.f, pz = 0.f;
      if (valid) {
        const float* pp = a.xyz + ((size_t)sample * a.n_points + n0 + row) * 3;
        px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2);
      }
      const uint32_t to = RO + lane_off + (uint32_t)hsel * 128u;
      mbar_wait(&out_full, pj);
      tc_fence_after();
      uint32_t raw[2][32];
      tmem_ld32(to, raw[0]);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c0 = hsel * 128 + g * 32;
        tmem_ld_wait();
        if (g < 3) {
          tmem_ld32(to + (g + 1) * 32, raw[(g + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&out_drained);
        }
        const uint32_t* r = raw[g & 1];
        float f[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(a.cst0 + c0) + i);
          f[4 * i] = __uint_as_float(r[4 * i]) + bv.x; f[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bv.y;
          f[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bv.z; f[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bv.w;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float4 pfv = __ldg(a.pfold + c0 + i);
          f[i] = fmaf(pfv.z, pz, fmaf(pfv.y, py, fmaf(pfv.x, px, f[i])));
        }
        {
          uint8_t* ct = stage + (hsel * 2 + (g >> 1)) * PF_CHUNK;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            *reinterpret_cast<uint4*>(ct + pf_sw128(row, (g & 1) * 4 + jj)) =
                make_uint4(pf_pack(f[8 * jj], f[8 * jj + 1]), pf_pack(f[8 * jj + 2], f[8 * jj + 3]), pf_pack(f[8 * jj + 4], f[8 * jj + 5]),
                           pf_pack(f[8 * jj + 6], f[8 * jj + 7]));
        }
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = -INFINITY;
        }
        const float mine = pf_colmax32(f, lane);
        atomicMax(&colmax_s[c0 + lane], f2ord(mine));
      }
      fence_proxy_async();
      pf_epi_sync();
      if (elected) {
        for (int kc = 0; kc < 4; ++kc) tma_store_3d(&tm.xout, stage + kc * PF_CHUNK, kc * 64, n0, sample);
        pf_store_commit();
        pf_store_wait_read();
      }
      pf_epi_sync();        // the staging buffer is free again for the next tile's writes
    }
    if (cur_sample >= 0) flush_colmax(cur_sample);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

#endif  // SEEME_EXPERIMENTAL
// ---------------------------------------------------------------------------------------------------------------------
// Block 0, tensor-core generator (default).  ncu on the kernel above showed the 4 generating warps (3 FMA + max + pack per
// value, 2 400 instructions per thread and tile behind table loads) taking ~4x the tile's MMA time.  Here fc_pos itself, the
// two biases and the rank-3 shortcut fold all become K = 16 tcgen05 MMAs on a 16-column "coordinate operand"
//     A16[r] = [xh yh zh | xl yl zl | xh yh zh | 1 1 | 0 0 0 0 0]        (x = xh + xl, two fp16 terms; fp32 accumulation)
// against constant B rows [wh(3) | wh(3) | wl(3) | bh bl | 0..]: the products are exact in fp32, so the result carries the
// fp32 affine map to ~2^-22 (the dropped xl.wl term), i.e. the same value the CUDA-core generator rounds to fp16.
//     F(kc) : TMEM[128 x 64] = A16 . Wp16[kc]^T           one MMA per 64-channel chunk, double-buffered in the OUT region
//     gen   : A[:, kc] = fp16(relu(F(kc)))                 H group: tcgen05.ld -> cvt.pack -> hmax2 -> swizzled st.shared
//     G1    : H   = A16 . B0_16^T + A . W0^T  (K = 512)    bias by MMA
//     epiH  : H16 = fp16(relu(H)) in place in TMEM
//     G2    : OUT = A16 . PF16^T + H16 . W1^T              shortcut fold + its constant by MMA
//     epiOUT: pooled column max; fp16 tile -> 32 KB staging (two rounds) -> TMA store
// The F staging buffers are OUT-region columns [0,64) and [128,192): the output epilogue drains those first and releases
// them (fr_free) long before the rest (out_drained), so the next tile's generator starts ~2 TMEM loads after G2 completes.
// Shared memory (14 x 16 KB): A ring 3 | staging 2 | constant tiles 2 | A16 1 | weight ring 6.
constexpr int P0T_ASLOTS = 3;
// Generator groups: 1 = warps 2-5 convert every chunk.  2 adds warps 14-17 for the odd chunks (independent chains), but 18
// warps leave 96 registers per thread (5 warps per scheduler) and measured slower (1.07 vs 0.90 ms per 128 clouds).
constexpr int P0T_GEN_GROUPS = 1;
constexpr int P0T_THREADS = PF_THREADS + (P0T_GEN_GROUPS - 1) * 128;
constexpr int P0T_SMEM = (P0T_ASLOTS + 2 + 2 + 1 + PF_NST) * PF_CHUNK + 1024;

struct P0TArgs {
  int n_points, tiles_per_sample, n_tiles;
  const float* xyz;       // [samples, n_points, 3]
  unsigned* colmax;
  const uint8_t* wblob;   // 24 pre-swizzled 16 KB weight chunks (G1 16, G2 8)
  const uint8_t* ctblob;  // 2 pre-swizzled 16 KB constant tiles: Wp16 | (B0_16, PF16)
};


// Schedule (per tile, 12 static steps; step s uses weight-ring pair s % 3 = 32 KB = two adjacent chunks):
//   steps 0..7  G1(kc): ONE barrier full[kc] collects the 4 generator warps' arrivals for A chunk kc (A-ring slot
//               kc % 3) AND the weight pair's TMA bytes; ONE commit empty[s % 3] releases both
//   steps 8..11 G2(kc): g2_full[kc] collects the 4 H-epilogue warps' arrivals and the weight pair's bytes
// 12 steps = 4 ring rounds per tile, so every slot index and phase parity below is a compile-time constant.
// F staging: chunks 0 and 1 of tile j+1 are produced DURING tile j's G2 into the H-region columns the in-place fp16
// conversion has freed ([64,128) after h_ready 0-1, [192,256) after h_ready 2-3), so G1(j+1) starts right behind G2(j);
// chunks 2..7 use OUT-region columns [0,64) / [128,192) once the output epilogue has released them (fr_free).
enum P0TBar { B_FULL = 0, B_EMPTY = 8, B_G2FULL = 11, B_FFULL = 15, B_A16 = 23, B_CT = 25, B_HFULL = 26, B_OUTFULL = 27, B_FRFREE = 28,
              B_DRAINED = 29, B_GEN1 = 30, B_COUNT = 31 };

__global__ void __launch_bounds__(P0T_THREADS, 1) pointnet_block0_tc_kernel(const __grid_constant__ PfMaps tm, const P0TArgs a) {
  extern __shared__ __align__(1024) uint8_t pf_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pf_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* aring = smem;                                  // 3 x 16 KB generated A chunks
  uint8_t* stage = aring + P0T_ASLOTS * PF_CHUNK;         // 2 x 16 KB output staging
  uint8_t* ct = stage + 2 * PF_CHUNK;                     // constant tiles
  uint8_t* a16 = ct + 2 * PF_CHUNK;                       // coordinate operand, K-slice (tile & 1)
  uint8_t* wring = a16 + PF_CHUNK;                        // 3 x 32 KB weight pairs
  __shared__ __align__(8) uint64_t bars[B_COUNT];
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned colmax_s[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = a.n_tiles / (int)gridDim.x, rem = a.n_tiles % (int)gridDim.x;
  const int t_begin = (int)blockIdx.x * per + ((int)blockIdx.x < rem ? (int)blockIdx.x : rem);
  const int nt = per + ((int)blockIdx.x < rem ? 1 : 0);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [bar0](int i) { return bar0 + (uint32_t)i * 8u; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm.xout);
    for (int i = 0; i < 8; ++i) mbar_init(&bars[B_FULL + i], 5);
    for (int i = 0; i < 3; ++i) mbar_init(&bars[B_EMPTY + i], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bars[B_G2FULL + i], 5);
    for (int i = 0; i < 8; ++i) mbar_init(&bars[B_FFULL + i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&bars[B_A16 + i], 4);
    mbar_init(&bars[B_CT], 1);
    mbar_init(&bars[B_HFULL], 1);
    mbar_init(&bars[B_OUTFULL], 1);
    mbar_init(&bars[B_FRFREE], 8);
    mbar_init(&bars[B_DRAINED], 8);
    mbar_init(&bars[B_GEN1], 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  if (threadIdx.x < 256) colmax_s[threadIdx.x] = 0u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t RH = tmem_base, RO = tmem_base + 256u;

  if (warp == 0) {
    // ---- weight producer: one 32 KB bulk copy per step ---------------------------------------------------------
    if (lane == 0) {
      const uint32_t wr = smem_u32(wring);
      pf_arrive_tx(BAR(B_CT), 2 * PF_CHUNK);
      pf_bulk_load(smem_u32(ct), a.ctblob, 2 * PF_CHUNK, BAR(B_CT));
      for (int j = 0; j < nt; ++j) {
#pragma unroll
        for (int s = 0; s < 12; ++s) {
          const uint32_t fb = s < 8 ? BAR(B_FULL + s) : BAR(B_G2FULL + s - 8);
          pf_wait(BAR(B_EMPTY + s % 3), (uint32_t)((s / 3) & 1) ^ 1u);
          PF_TR(j, 48 + s);
          pf_arrive_tx(fb, 2 * PF_CHUNK);
          pf_bulk_load(wr + (uint32_t)(s % 3) * 2 * PF_CHUNK, a.wblob + (size_t)s * 2 * PF_CHUNK, 2 * PF_CHUNK, fb);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer --------------------------------------------------------------------------------------------------
    constexpr uint32_t idesc256 = umma_idesc_f16(256), idesc128 = umma_idesc_f16(128), idesc64 = umma_idesc_f16(64);
    const uint64_t wdesc0 = umma_desc_k128(smem_u32(wring));
    const uint64_t adesc0 = umma_desc_k128(smem_u32(aring));
    const uint64_t ct1desc = umma_desc_k128(smem_u32(ct)), ct2desc = umma_desc_k128(smem_u32(ct + PF_CHUNK));
    const uint64_t a16desc0 = umma_desc_k128(smem_u32(a16));
    // F(kc): 64 channels kc*64.. = K-slice (kc >> 1), rows (kc & 1) * 64.. of the Wp16 tile
    auto issue_f = [&](uint64_t ad, int kc, bool commit = true) {
      const uint32_t dst = kc == 0 ? RH + 64u : kc == 1 ? RH + 192u : RO + (uint32_t)(kc & 1) * 128u;
      umma_bf16(dst, ad, pf_desc_add(ct1desc, (uint32_t)(kc & 1) * (64 * 128 >> 4) + (uint32_t)(kc >> 1) * 2), idesc64, 0);
      if (commit) pf_commit(BAR(B_FFULL + kc));
    };
    pf_wait(BAR(B_CT), 0);
    pf_wait(BAR(B_A16 + 0), 0);
    tc_fence_after();
    if (pf_elect_one()) {
      issue_f(a16desc0, 0);
      issue_f(a16desc0, 1);
    }
    __syncwarp();
    for (int j = 0; j < nt; ++j) {
      const uint32_t pj = (uint32_t)j & 1u;
      const uint64_t ad16 = pf_desc_add(a16desc0, pj * 2);
      const uint64_t ad16n = pf_desc_add(a16desc0, (pj ^ 1u) * 2);
      const bool more = j + 1 < nt;
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        if (lane == 0) PF_TR(j, 1 + 2 * kc);
        pf_wait(BAR(B_FULL + kc), pj);
        if (kc == 0) pf_wait(BAR(B_GEN1), pj);         // chunk 1 generated too (its weights may still be in flight): both H-region F buffers have been read
        tc_fence_after();
        if (pf_elect_one()) {
          PF_TR(j, 2 + 2 * kc);
          const uint64_t wd = pf_desc_add(wdesc0, (uint32_t)(kc % 3) * (2 * PF_CHUNK >> 4));
          const uint64_t ad = pf_desc_add(adesc0, (uint32_t)(kc % 3) * (PF_CHUNK >> 4));
          if (kc == 0) {
            // H = bias (accumulate = 0): G2 of the previous tile, which read H16 from this region, is ordered before by the MMA pipe
            umma_bf16(RH, ad16, ct2desc, idesc128, 0);
            umma_bf16(RH + 128u, ad16, pf_desc_add(ct2desc, 2), idesc128, 0);
          }
          if (kc >= 2 && kc + 2 < 8) issue_f(ad16, kc + 2);      // its F buffer (kc & 1) was read for chunk kc
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(RH, pf_desc_add(ad, ks * 2), pf_desc_add(wd, ks * 2), idesc256, 1);
          pf_commit(BAR(B_EMPTY + kc % 3));
          if (kc == 7) pf_commit(BAR(B_HFULL));
        }
        __syncwarp();
        if (kc == 0) {
          // chunks 2 and 3 go to the OUT-region buffers: the previous tile's epilogue must have read those columns
          if (j > 0) pf_wait(BAR(B_FRFREE), (uint32_t)(j - 1) & 1u);
          tc_fence_after();
          if (pf_elect_one()) {
            issue_f(ad16, 2, false);     // one commit (f_full[3]) covers both; the generator waits on it for chunk 2 as well
            issue_f(ad16, 3);
          }
          __syncwarp();
        }
      }
      // G2 writes the whole OUT region (incl. the F buffers, all read by now): the previous tile's epilogue must have drained it
      if (j > 0) pf_wait(BAR(B_DRAINED), (uint32_t)(j - 1) & 1u);
      if (more) pf_wait(BAR(B_A16 + (pj ^ 1u)), (uint32_t)((j + 1) >> 1) & 1u);
      tc_fence_after();
      if (pf_elect_one()) {
        PF_TR(j, 17);
        umma_bf16(RO, ad16, pf_desc_add(ct2desc, 4), idesc128, 0);
        umma_bf16(RO + 128u, ad16, pf_desc_add(ct2desc, 6), idesc128, 0);
      }
      __syncwarp();
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        constexpr int S0 = 8;
        if (lane == 0) PF_TR(j, 18 + 2 * kc);
        pf_wait(BAR(B_G2FULL + kc), pj);
        tc_fence_after();
        if (pf_elect_one()) {
          PF_TR(j, 19 + 2 * kc);
          const uint64_t wd = pf_desc_add(wdesc0, (uint32_t)((S0 + kc) % 3) * (2 * PF_CHUNK >> 4));
          // next tile's chunk 0 / 1 into the H-region columns freed by the fp16 conversion of H columns [0,128) / [128,256)
          if (kc == 1 && more) issue_f(ad16n, 0);
          if (kc == 3 && more) issue_f(ad16n, 1);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16_ts(RO, RH + (uint32_t)((kc >> 1) * 128 + (kc & 1) * 32 + ks * 8), pf_desc_add(wd, ks * 2), idesc256, 1);
          pf_commit(BAR(B_EMPTY + (S0 + kc) % 3));
          if (kc == 3) pf_commit(BAR(B_OUTFULL));
        }
        __syncwarp();
      }
    }
  } else if (warp < 6 || warp >= 14) {
    // ---- generator groups: thread = point.  Group 0 (warps 2-5) converts the even chunks and writes the coordinate
    // operand, group 1 (warps 14-17) the odd chunks; each group owns one F staging buffer, so the two chains are independent
    const int gsel = warp >= 14 ? 1 : 0;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t ar = smem_u32(aring), a16a = smem_u32(a16);
    float px = 0.f, py = 0.f, pz = 0.f;
    bool pvalid = false;
    auto load_xyz = [&](int j) {
      px = py = pz = 0.f;
      pvalid = false;
      if (j < nt) {
        const int t = t_begin + j;
        const int sample = t / a.tiles_per_sample, n0 = (t % a.tiles_per_sample) * 128;
        if (n0 + row < a.n_points) {
          const float* pp = a.xyz + ((size_t)sample * a.n_points + n0 + row) * 3;
          px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2);
          pvalid = true;
        }
      }
    };
    auto st128 = [](uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
    };
    auto write_a16 = [&](int buf) {
      const __half xh = __float2half_rn(px), yh = __float2half_rn(py), zh = __float2half_rn(pz);
      const __half xl = __float2half_rn(px - __half2float(xh)), yl = __float2half_rn(py - __half2float(yh)),
                   zl = __float2half_rn(pz - __half2float(zh));
      const __half one = __float2half_rn(pvalid ? 1.f : 0.f), zero = __float2half_rn(0.f);
      auto pk = [](__half lo, __half hi) { return (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16); };
      st128(a16a + pf_sw128(row, 2 * buf), pk(xh, yh), pk(zh, xl), pk(yl, zl), pk(xh, yh));
      st128(a16a + pf_sw128(row, 2 * buf + 1), pk(zh, one), pk(one, zero), 0u, 0u);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) pf_arrive(BAR(B_A16 + buf));
    };
    if (gsel == 0) {
      load_xyz(0);
      write_a16(0);
      load_xyz(1);
    }
    for (int j = 0; j < nt; ++j) {
      const uint32_t pj = (uint32_t)j & 1u;
#pragma unroll
      for (int kc2 = 0; kc2 < 8; kc2 += P0T_GEN_GROUPS) {
        const int kc = kc2 + gsel;
        pf_wait(BAR(B_FFULL + (kc == 2 ? 3 : kc)), pj);
        if (threadIdx.x == 64) PF_TR(j, 26 + 2 * kc);
        tc_fence_after();
        uint32_t r[64];
        const uint32_t tf = (kc == 0 ? RH + 64u : kc == 1 ? RH + 192u : RO + (uint32_t)(kc & 1) * 128u) + lane_off;
        tmem_ld32(tf, r);
        tmem_ld32(tf + 32, r + 32);
        // A-ring slot kc % 3: released by the commit of the step that used ring index kc % 3 before (3 steps earlier)
        pf_wait(BAR(B_EMPTY + kc % 3), (uint32_t)((kc / 3) & 1) ^ 1u);
        const uint32_t ctile = ar + (uint32_t)(kc % 3) * PF_CHUNK;
        if (threadIdx.x == 64 && kc == 4) PF_TR(j, 60);
        tmem_ld_wait();
        if (threadIdx.x == 64 && kc == 4) PF_TR(j, 61);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          st128(ctile + pf_sw128(row, jj), pf_relu_pack(r[8 * jj], r[8 * jj + 1]), pf_relu_pack(r[8 * jj + 2], r[8 * jj + 3]),
                pf_relu_pack(r[8 * jj + 4], r[8 * jj + 5]), pf_relu_pack(r[8 * jj + 6], r[8 * jj + 7]));
        if (threadIdx.x == 64 && kc == 4) PF_TR(j, 62);
        tc_fence_before();
        fence_proxy_async();
        if (threadIdx.x == 64 && kc == 4) PF_TR(j, 63);
        __syncwarp();
        if (lane == 0) {
          pf_arrive(BAR(B_FULL + kc));
          if (kc == 1) pf_arrive(BAR(B_GEN1));
        }
        if (threadIdx.x == 64) PF_TR(j, 27 + 2 * kc);
      }
      // coordinate operand of the next tile: its K-slice was last read by tile j-1's MMAs, all complete before f_full(j, *)
      if (gsel == 0) {
        if (j + 1 < nt) write_a16((j + 1) & 1);
        load_xyz(j + 2);
      }
    }
  } else {
    // ---- O group ---------------------------------------------------------------------------------------------------
    const int te = (int)threadIdx.x - 192;
    const int q = warp & 3;
    const int hsel = (warp - 6) >> 2;
    const int row = q * 32 + lane;
    const bool elected = te == 0;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    auto flush_colmax = [&](int sample) {
      pf_epi_sync();
      const unsigned v = colmax_s[te];
      if (v) atomicMax(a.colmax + (size_t)sample * 256 + te, v);
      colmax_s[te] = 0u;
      pf_epi_sync();
    };
    const uint32_t st0 = smem_u32(stage) + (uint32_t)hsel * PF_CHUNK + (uint32_t)row * 128u;
    const uint32_t swz = (uint32_t)(row & 7);
    auto st128 = [](uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
    };
    auto pack2 = [](uint32_t lo, uint32_t hi) {
      const __half2 h = __floats2half2_rn(__uint_as_float(lo), __uint_as_float(hi));
      return *reinterpret_cast<const uint32_t*>(&h);
    };
    // pooled column max of 32 columns (hsel * 128 + g * 32 ..) held one row per lane; r is destroyed
    auto colmax = [&](uint32_t (&r)[32], int g, bool valid) {
      float(&f)[32] = reinterpret_cast<float(&)[32]>(r);
      if (!valid) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = -INFINITY;
      }
      const float mine = pf_colmax32(f, lane);
      atomicMax(&colmax_s[hsel * 128 + g * 32 + lane], f2ord(mine));
    };
    // fp16 of the 32 columns straight into the staging chunk (16-byte chunks (g & 1) * 4 .. of this thread's row)
    auto stage_from_raw = [&](const uint32_t (&r)[32], int g) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        st128(st0 + ((((uint32_t)((g & 1) * 4 + jj)) ^ swz) << 4), pack2(r[8 * jj], r[8 * jj + 1]), pack2(r[8 * jj + 2], r[8 * jj + 3]),
              pack2(r[8 * jj + 4], r[8 * jj + 5]), pack2(r[8 * jj + 6], r[8 * jj + 7]));
    };
    int cur_sample = -1;
    for (int j = 0; j < nt; ++j) {
      const uint32_t pj = (uint32_t)j & 1u;
      const int t = t_begin + j;
      const int sample = t / a.tiles_per_sample, n0 = (t % a.tiles_per_sample) * 128;
      if (sample != cur_sample) {
        if (cur_sample >= 0) flush_colmax(cur_sample);
        cur_sample = sample;
      }
      // H epilogue (in place in TMEM): relu + fp16 of H columns hsel * 128 .., the bias is already in the accumulator.
      // Done by this group (8 warps, idle between two output epilogues) so that the generator warps keep generating.
      pf_wait(BAR(B_HFULL), pj);
      if (elected) PF_TR(j, 42);
      tc_fence_after();
      {
        const uint32_t thh = RH + lane_off + (uint32_t)hsel * 128u;
        uint32_t raw[2][32];
        tmem_ld32(thh, raw[0]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          tmem_ld_wait();
          if (g < 3) tmem_ld32(thh + (g + 1) * 32, raw[(g + 1) & 1]);
          const uint32_t* r = raw[g & 1];
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pf_relu_pack(r[2 * i], r[2 * i + 1]);
          tmem_st16(thh + g * 16, pk);
          if (g & 1) {
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) pf_arrive(BAR(B_G2FULL + hsel * 2 + (g >> 1)));
          }
        }
      }
      if (elected) PF_TR(j, 43);
      const bool valid = n0 + row < a.n_points;
      const uint32_t to = RO + lane_off + (uint32_t)hsel * 128u;
      pf_wait(BAR(B_OUTFULL), pj);
      if (elected) PF_TR(j, 44);
      tc_fence_after();
      uint32_t raw[2][32];
      tmem_ld32(to, raw[0]);
      tmem_ld32(to + 32, raw[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) pf_arrive(BAR(B_FRFREE));       // columns [0,64) and [128,192): the next tile's F staging buffers
      if (elected) { PF_TR(j, 45); pf_store_wait_read(); }     // the previous tile's second store round has read the staging chunks
      pf_epi_sync();
      stage_from_raw(raw[0], 0);
      colmax(raw[0], 0, valid);
      tmem_ld32(to + 64, raw[0]);
      stage_from_raw(raw[1], 1);
      colmax(raw[1], 1, valid);
      tmem_ld_wait();
      tmem_ld32(to + 96, raw[1]);
      fence_proxy_async();
      pf_epi_sync();
      if (elected) {
        tma_store_3d(&tm.xout, stage, 0, n0, sample);
        tma_store_3d(&tm.xout, stage + PF_CHUNK, 128, n0, sample);
        pf_store_commit();
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) pf_arrive(BAR(B_DRAINED));
      if (elected) PF_TR(j, 46);
      // second round: pack (16 words per 32 columns) and pool first -- that hides the first round's store reading the staging chunks
      uint32_t pk[2][16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[0][i] = pack2(raw[0][2 * i], raw[0][2 * i + 1]);
      colmax(raw[0], 2, valid);
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[1][i] = pack2(raw[1][2 * i], raw[1][2 * i + 1]);
      colmax(raw[1], 3, valid);
      if (elected) pf_store_wait_read();
      pf_epi_sync();        // the staging chunks are free for the second round
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          st128(st0 + ((((uint32_t)(h2 * 4 + jj)) ^ swz) << 4), pk[h2][4 * jj], pk[h2][4 * jj + 1], pk[h2][4 * jj + 2], pk[h2][4 * jj + 3]);
      fence_proxy_async();
      pf_epi_sync();
      if (elected) {
        tma_store_3d(&tm.xout, stage, 64, n0, sample);
        tma_store_3d(&tm.xout, stage + PF_CHUNK, 192, n0, sample);
        pf_store_commit();
        PF_TR(j, 47);
      }
    }
    if (elected) pf_store_wait_read();
    if (cur_sample >= 0) flush_colmax(cur_sample);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

#ifdef SEEME_EXPERIMENTAL   // retired variant (slower than the single-CTA kernel): experimental builds only
// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair version of pointnet_block_kernel: tcgen05.mma.cta_group::2, M = 256 (two 128-point tiles, one per SM), N = 256.
// Each CTA holds only ITS half of every weight K-chunk ([128 n x 64 k]: CTA 0 the output columns 0-127, CTA 1 128-255),
// so the per-SM weight traffic from L2 -- what bounds the single-CTA kernel -- halves and the 6-slot ring holds 6 K-steps.
// Protocol: the leader CTA's warp 1 issues every MMA; tcgen05.commit multicasts completion to the barriers of both CTAs;
// the peer CTA's warp 1 relays "my X tile / my weight slot has landed" to the leader with remote mbarrier arrives; the
// epilogue warps of the peer arrive remotely on the leader's r_done / h_ready / out_drained barriers.
__device__ __forceinline__ uint32_t pf_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void pf_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// arrive on the leader CTA's copy of a barrier
__device__ __forceinline__ void pf_arrive_leader(uint64_t* bar, uint32_t rank) {
  if (rank == 0) asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
  else pf_arrive_remote(pf_mapa(smem_u32(bar), 0));
}
// wait with cluster-scope acquire (the barrier receives arrivals from the peer CTA)
__device__ __forceinline__ void pf_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  pf_wait_timeout(addr);
}
__device__ __forceinline__ void umma2_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of this thread's cta_group::2 MMAs -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__host__ __device__ constexpr uint32_t umma2_idesc_f16(int n) {     // M = 256 over the CTA pair
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

constexpr int PF2_WSTEPS = 12;      // weight K-steps per tile: S 4, G1 4, G2 4 (each CTA loads its 128-row half)

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PF_THREADS, 1)
    pointnet_block_pair_kernel(const __grid_constant__ PfMaps tm, const PfArgs a) {
  extern __shared__ __align__(1024) uint8_t pf_smem_raw[];
  uint8_t* smem = pf_smem_raw;                       // (the 1 KB of alignment slack holds the bias copy, as in pointnet_block_kernel)
  if ((smem_u32(pf_smem_raw) & 1023u) != 0u) __trap();
  uint8_t* xbuf = smem;
  uint8_t* wring = smem + 2 * PF_XBUF;
  // local barriers (every CTA): w_full, x_full (TMA), w_empty, s_done, h_full, out_full (multicast commits)
  // leader-only use: wp_full, xp_full (relayed by the peer), r_done, h_ready, out_drained (both CTAs' epilogue warps)
  __shared__ __align__(8) uint64_t w_full[PF_NST], w_empty[PF_NST], wp_full[PF_NST], x_full[2], xp_full[2], s_done[2][4], r_done[2][4],
      h_full[2], h_ready[2][4], out_full[2], out_drained[2];
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned colmax_s[256];
  __shared__ __align__(16) float bias_o_s[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)blockIdx.x >> 1, npairs = (int)gridDim.x >> 1;
  // tile pairs [pb, pb + np) of this CTA pair; this CTA takes tile 2 * (pb + j) + rank (a ghost tile beyond the end has
  // no valid rows: zero-filled loads, clipped stores, no pooling contribution)
  const int tile_pairs = (a.n_tiles + 1) / 2;
  const int per = tile_pairs / npairs, rem = tile_pairs % npairs;
  const int pb = pair * per + (pair < rem ? pair : rem);
  const int np = per + (pair < rem ? 1 : 0);
  auto tile_of = [&](int j, int& sample, int& n0) {
    const int t = 2 * (pb + j) + (int)rank;
    if (t < a.n_tiles) { sample = t / a.tiles_per_sample; n0 = (t % a.tiles_per_sample) * 128; }
    else { sample = (a.n_tiles - 1) / a.tiles_per_sample; n0 = a.tiles_per_sample * 128; }      // ghost: rows >= n_points
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm.xin);
    tma_prefetch_desc(&tm.w);
    tma_prefetch_desc(&tm.xout);
    for (int i = 0; i < PF_NST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); mbar_init(&wp_full[i], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&x_full[b], 1);
      mbar_init(&xp_full[b], 1);
      mbar_init(&out_full[b], 1);
      mbar_init(&out_drained[b], 16);
      mbar_init(&h_full[b], 1);
      for (int k = 0; k < 4; ++k) { mbar_init(&s_done[b][k], 1); mbar_init(&r_done[b][k], 8); mbar_init(&h_ready[b][k], 8); }
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(&tmem_slot, 512);
  if (threadIdx.x < 256) colmax_s[threadIdx.x] = 0u;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // both CTAs' barriers and TMEM exist before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ---- weight producer: this CTA's half (128 output columns) of each K-step ---------------------------------------
    if (lane == 0) {
      uint32_t st = 0, ph = 1;
      for (int j = 0; j < np; ++j) {
        for (int i = 0; i < PF2_WSTEPS; ++i) {
          const int phase = i >> 2, kc = i & 3;
          // blob chunk order: S (kc, nh), G1 (nh, kc), G2 (kc, nh)
          const int kk = phase == 2 ? (((kc & 1) << 1) | (kc >> 1)) : kc;      // G2 runs its K-chunks in the order 0, 2, 1, 3
          const int chunk = phase == 1 ? 8 + (int)rank * 4 + kc : phase * 8 + kk * 2 + (int)rank;
          mbar_wait(&w_empty[st], ph);
          mbar_arrive_expect_tx(&w_full[st], PF_CHUNK);
          bulk_load(wring + st * PF_CHUNK, a.wblob + (size_t)chunk * PF_CHUNK, PF_CHUNK, &w_full[st]);
          if (++st == PF_NST) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && rank == 1) {
    // ---- peer relay: tell the leader's MMA thread when this CTA's X tile / weight slot has landed ---------------------
    if (lane == 0) {
      uint32_t st = 0, wph = 0;
      for (int j = 0; j < np; ++j) {
        const int b = j & 1;
        mbar_wait(&x_full[b], (uint32_t)(j >> 1) & 1u);
        pf_arrive_remote(pf_mapa(smem_u32(&xp_full[b]), 0));
        for (int i = 0; i < PF2_WSTEPS; ++i) {
          mbar_wait(&w_full[st], wph);
          pf_arrive_remote(pf_mapa(smem_u32(&wp_full[st]), 0));
          if (++st == PF_NST) { st = 0; wph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA): M = 256 over the pair ---------------------------------------------------------------
    constexpr uint32_t idesc = umma2_idesc_f16(256);
    const uint64_t wdesc0 = umma_desc_k128(smem_u32(wring));
    uint32_t st = 0, wph = 0;
    auto wait_w = [&]() {
      mbar_wait(&w_full[st], wph);
      pf_wait_cluster(&wp_full[st], wph);
    };
    auto next_w = [&]() {
      if (++st == PF_NST) { st = 0; wph ^= 1u; }
    };
    for (int j = 0; j < np; ++j) {
      const int b = j & 1;
      const uint32_t p2 = (uint32_t)(j >> 1) & 1u;
      const uint32_t Ra = tmem_base + (uint32_t)b * 256u, Rb = tmem_base + (uint32_t)(b ^ 1) * 256u;
      const uint64_t xdesc = umma_desc_k128(smem_u32(xbuf + b * PF_XBUF));
      mbar_wait(&x_full[b], p2);
      pf_wait_cluster(&xp_full[b], p2);
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {          // S: OUT = X . Ws^T
        wait_w();
        tc_fence_after();
        if (pf_elect_one()) {
          const uint64_t wd = pf_desc_add(wdesc0, st * (PF_CHUNK >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma2_f16_ss(Ra, pf_desc_add(xdesc, kc * (PF_CHUNK >> 4) + ks * 2), pf_desc_add(wd, ks * 2), idesc, (kc | ks) != 0);
          umma2_commit(&w_empty[st]);
          umma2_commit(&s_done[b][kc]);
        }
        __syncwarp();
        next_w();
      }
      if (j > 0) pf_wait_cluster(&out_drained[b ^ 1], (uint32_t)((j - 1) >> 1) & 1u);
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {          // G1: H = relu(X) . W0^T
        pf_wait_cluster(&r_done[b][kc], p2);
        wait_w();
        tc_fence_after();
        if (pf_elect_one()) {
          const uint64_t wd = pf_desc_add(wdesc0, st * (PF_CHUNK >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma2_f16_ss(Rb, pf_desc_add(xdesc, kc * (PF_CHUNK >> 4) + ks * 2), pf_desc_add(wd, ks * 2), idesc, (kc | ks) != 0);
          umma2_commit(&w_empty[st]);
          if (kc == 3) umma2_commit(&h_full[b]);
        }
        __syncwarp();
        next_w();
      }
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {          // G2: OUT += H16 . W1^T, A operand from TMEM; K-chunk order 0, 2, 1, 3
        const int kc = ((i4 & 1) << 1) | (i4 >> 1);
        pf_wait_cluster(&h_ready[b][kc], p2);
        wait_w();
        tc_fence_after();
        if (pf_elect_one()) {
          const uint64_t wd = pf_desc_add(wdesc0, st * (PF_CHUNK >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma2_f16_ts(Ra, Rb + (uint32_t)((kc >> 1) * 128 + (kc & 1) * 32 + ks * 8), pf_desc_add(wd, ks * 2), idesc, 1);
          umma2_commit(&w_empty[st]);
          if (i4 == 3) umma2_commit(&out_full[b]);
        }
        __syncwarp();
        next_w();
      }
    }
  } else if (warp < 6) {
    // ---- H group ------------------------------------------------------------------------------------------------------
    const int th = (int)threadIdx.x - 64;
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    for (int j = 0; j < np; ++j) {
      const int b = j & 1;
      const uint32_t p2 = (uint32_t)(j >> 1) & 1u;
      int sample, n0;
      tile_of(j, sample, n0);
      uint8_t* xb = xbuf + b * PF_XBUF;
      for (int kc = 0; kc < 4; ++kc) {
        mbar_wait(&s_done[b][kc], p2);
        uint4* p = reinterpret_cast<uint4*>(xb + kc * PF_CHUNK) + th;
        const __half2 z = __float2half2_rn(0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 v = p[i * 128];
          __half2* h = reinterpret_cast<__half2*>(&v);
          h[0] = __hmax2(h[0], z); h[1] = __hmax2(h[1], z); h[2] = __hmax2(h[2], z); h[3] = __hmax2(h[3], z);
          p[i * 128] = v;
        }
        asm volatile("fence.proxy.async;" ::: "memory");     // generic writes -> async proxy (the leader's MMA reads this tile)
        __syncwarp();
        if (lane == 0) pf_arrive_leader(&r_done[b][kc], rank);
      }
    }
  } else {
    // ---- O group ------------------------------------------------------------------------------------------------------
    const int te = (int)threadIdx.x - 192;
    const int q = warp & 3;
    const int hsel = (warp - 6) >> 2;
    const int row = q * 32 + lane;
    const bool elected = te == 0;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    auto load_x = [&](int j) {
      int sample, n0;
      tile_of(j, sample, n0);
      const int b = j & 1;
      mbar_arrive_expect_tx(&x_full[b], PF_XBUF);
      for (int kc = 0; kc < 4; ++kc) tma_load_3d(xbuf + b * PF_XBUF + kc * PF_CHUNK, &tm.xin, &x_full[b], kc * 64, n0, sample);
    };
    auto flush_colmax = [&](int sample) {
      pf_epi_sync();
      const unsigned v = colmax_s[te];
      if (v) atomicMax(a.colmax + (size_t)sample * 256 + te, v);
      colmax_s[te] = 0u;
      pf_epi_sync();
    };
    if (elected) {
      if (np > 0) load_x(0);
      if (np > 1) load_x(1);
    }
    int cur_sample = -1;
    for (int j = 0; j < np; ++j) {
      const int b = j & 1;
      const uint32_t p2 = (uint32_t)(j >> 1) & 1u;
      int sample, n0;
      tile_of(j, sample, n0);
      uint8_t* xb = xbuf + b * PF_XBUF;
      if (sample != cur_sample) {
        if (cur_sample >= 0) flush_colmax(cur_sample);       // (ends with a group barrier: nobody reads the old bias any more)
        bias_o_s[te] = __ldg(a.bias_o + (size_t)sample * 256 + te);
        pf_epi_sync();
        cur_sample = sample;
      }
      // H epilogue of this CTA's tile (8 warps, both column halves in parallel); arrivals go to the leader's h_ready
      {
        const float* bh = a.bias_h + (size_t)sample * 256 + hsel * 128;
        const uint32_t thh = tmem_base + (uint32_t)(b ^ 1) * 256u + lane_off + (uint32_t)hsel * 128u;
        float4 bv0[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bv0[i] = __ldg(reinterpret_cast<const float4*>(bh) + i);
        mbar_wait(&h_full[b], p2);
        tc_fence_after();
        uint32_t raw[2][32];
        tmem_ld32(thh, raw[0]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float4 bv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) bv[i] = g == 0 ? bv0[i] : __ldg(reinterpret_cast<const float4*>(bh + g * 32) + i);
          tmem_ld_wait();
          if (g < 3) tmem_ld32(thh + (g + 1) * 32, raw[(g + 1) & 1]);
          const uint32_t* r = raw[g & 1];
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            pk[2 * i] = pf_relu_pack(__float_as_uint(__uint_as_float(r[4 * i]) + bv[i].x), __float_as_uint(__uint_as_float(r[4 * i + 1]) + bv[i].y));
            pk[2 * i + 1] = pf_relu_pack(__float_as_uint(__uint_as_float(r[4 * i + 2]) + bv[i].z), __float_as_uint(__uint_as_float(r[4 * i + 3]) + bv[i].w));
          }
          tmem_st16(thh + g * 16, pk);
          if (g & 1) {
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) pf_arrive_leader(&h_ready[b][hsel * 2 + (g >> 1)], rank);
          }
        }
      }
      const uint32_t to = tmem_base + (uint32_t)b * 256u + (uint32_t)hsel * 128u + lane_off;
      const bool valid = n0 + row < a.n_points;
      mbar_wait(&out_full[b], p2);
      tc_fence_after();
      // drain first (the region is the next tile's H accumulator), then stage and pool on the packed values
      uint32_t pk[4][16];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t raw[32];
        tmem_ld32(to + g * 32, raw);
        float4 bv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bv[i] = *(reinterpret_cast<const float4*>(bias_o_s + hsel * 128 + g * 32) + i);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          pk[g][2 * i] = pf_pack(__uint_as_float(raw[4 * i]) + bv[i].x, __uint_as_float(raw[4 * i + 1]) + bv[i].y);
          pk[g][2 * i + 1] = pf_pack(__uint_as_float(raw[4 * i + 2]) + bv[i].z, __uint_as_float(raw[4 * i + 3]) + bv[i].w);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) pf_arrive_leader(&out_drained[b], rank);
      if (a.store_out) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint8_t* ct = xb + (hsel * 2 + (g >> 1)) * PF_CHUNK;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            *reinterpret_cast<uint4*>(ct + pf_sw128(row, (g & 1) * 4 + jj)) = make_uint4(pk[g][4 * jj], pk[g][4 * jj + 1], pk[g][4 * jj + 2], pk[g][4 * jj + 3]);
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[g][i] = 0xfc00fc00u;      // -inf, -inf
        }
        const float mine = pf_colmax32_h2(pk[g], lane);
        atomicMax(&colmax_s[hsel * 128 + g * 32 + lane], f2ord(mine));
      }
      if (a.store_out) fence_proxy_async();
      pf_epi_sync();
      if (elected) {
        if (a.store_out) {
          for (int kc = 0; kc < 4; ++kc) tma_store_3d(&tm.xout, xb + kc * PF_CHUNK, kc * 64, n0, sample);
          pf_store_commit();
          pf_store_wait_read();
        }
        if (j + 2 < np) load_x(j + 2);
      }
    }
    if (cur_sample >= 0) flush_colmax(cur_sample);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer may still be reading this CTA's shared / tensor memory through the pair MMAs
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

#endif  // SEEME_EXPERIMENTAL
// ---- host side ----------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_pf_encode = nullptr;
static int pf_encoder() {
  if (g_pf_encode) return SEEME_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SEEME_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  SEEME_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, SEEME_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  g_pf_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return SEEME_OK;
}

// activations [samples, n_points, 256] fp16 as a 3-D map (channels, points, samples), box [64, 128, 1]
static int pf_act_map(CUtensorMap* map, const void* ptr, int samples, int n_points) {
  cuuint64_t gdim[3] = {256, (cuuint64_t)n_points, (cuuint64_t)samples};
  cuuint64_t gstr[2] = {512, (cuuint64_t)n_points * 512};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_pf_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SEEME_REQUIRE(r == CUDA_SUCCESS, SEEME_ECUDA, "cuTensorMapEncodeTiled (activations) failed with CUresult %d", (int)r);
  return SEEME_OK;
}

static int pf_w_map(CUtensorMap* map, const void* ptr) {
  cuuint64_t gdim[2] = {64, (cuuint64_t)PF_WCHUNKS * 128};
  cuuint64_t gstr[1] = {128};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_pf_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SEEME_REQUIRE(r == CUDA_SUCCESS, SEEME_ECUDA, "cuTensorMapEncodeTiled (weights) failed with CUresult %d", (int)r);
  return SEEME_OK;
}

// element index of (row r, column cc) inside a [128 x 64] fp16 chunk stored as its SWIZZLE_128B shared-memory image
__device__ __forceinline__ int pf_swz_elem(int r, int cc) { return r * 64 + ((((cc >> 3) ^ (r & 7)) << 3) | (cc & 7)); }

// chunk blob layout (24 x [128, 64] fp16, each the byte image of its swizzled shared-memory tile):
// S (kc, nh) from Ws[:, :256]; G1 (nh, kc) from W0[:, :256]; G2 (kc, nh) from W1
__global__ void pf_pack_weights_kernel(const float* __restrict__ ws, const float* __restrict__ w0, const float* __restrict__ w1,
                                       __half* __restrict__ blob) {
  const int chunk = blockIdx.x;
  const int phase = chunk / 8, idx = chunk % 8;
  int kc, nh;
  if (phase == 1) { nh = idx / 4; kc = idx % 4; } else { kc = idx / 2; nh = idx % 2; }
  const float* src = phase == 0 ? ws : phase == 1 ? w0 : w1;
  const int ld = phase == 2 ? 256 : 512;
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int r = i / 64, cc = i % 64;
    blob[(size_t)chunk * 128 * 64 + pf_swz_elem(r, cc)] = __float2half_rn(src[(size_t)(nh * 128 + r) * ld + kc * 64 + cc]);
  }
}

// block 0 blob: G1 (kc 0..7, nh) from W0 [256,512]; G2 (kc 0..3, nh) from W1 [256,256]
__global__ void pf_pack_weights0_kernel(const float* __restrict__ w0, const float* __restrict__ w1, __half* __restrict__ blob) {
  const int chunk = blockIdx.x;
  const bool g1 = chunk < 16;
  const int idx = g1 ? chunk : chunk - 16;
  const int kc = idx / 2, nh = idx % 2;
  const float* src = g1 ? w0 : w1;
  const int ld = g1 ? 512 : 256;
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int r = i / 64, cc = i % 64;
    blob[(size_t)chunk * 128 * 64 + pf_swz_elem(r, cc)] = __float2half_rn(src[(size_t)(nh * 128 + r) * ld + kc * 64 + cc]);
  }
}

__global__ void pf_pack_wpb_kernel(const float* __restrict__ wp, const float* __restrict__ bp, float4* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < 512) out[c] = make_float4(wp[c * 3], wp[c * 3 + 1], wp[c * 3 + 2], bp[c]);
}

int pf_pack_block0(const float* w0, const float* w1, const float* wp, const float* bp, void* blob, void* wpb) {
  pf_pack_weights0_kernel<<<PF_WCHUNKS, 256>>>(w0, w1, reinterpret_cast<__half*>(blob));
  SEEME_LAUNCH_CHECK();
  pf_pack_wpb_kernel<<<2, 256>>>(wp, bp, reinterpret_cast<float4*>(wpb));
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

// constant tiles of the tensor-core block 0: two [128 rows x 64 k] fp16 SWIZZLE_128B images, K-slice s = columns 16 s..
//   tile 0, slice s   : Wp16 rows of channels 128 s + n      [wh(3) | wh(3) | wl(3) | bh bl | 0 x 5]
//   tile 1, slice 0/1 : B0_16 of channels n / 128 + n        [0 x 9 | b0h b0l | 0 x 5]
//   tile 1, slice 2/3 : PF16  of channels n / 128 + n        [ph(3) | ph(3) | pl(3) | ch cl | 0 x 5]   (pfold, cst0)
__global__ void pf_pack_ct_kernel(const float* __restrict__ wp, const float* __restrict__ bp, const float* __restrict__ b0,
                                  const float* __restrict__ pfold, const float* __restrict__ cst0, __half* __restrict__ blob) {
  const int tile = blockIdx.x;
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int n = i / 64, cc = i % 64, sl = cc >> 4, k = cc & 15;
    float w3[3] = {0.f, 0.f, 0.f}, bias = 0.f;
    if (tile == 0) {
      const int c = sl * 128 + n;
      w3[0] = wp[c * 3]; w3[1] = wp[c * 3 + 1]; w3[2] = wp[c * 3 + 2]; bias = bp[c];
    } else if (sl < 2) {
      bias = b0[sl * 128 + n];
    } else {
      const int c = (sl - 2) * 128 + n;
      w3[0] = pfold[c * 4]; w3[1] = pfold[c * 4 + 1]; w3[2] = pfold[c * 4 + 2]; bias = cst0[c];
    }
    float v = 0.f;
    if (k < 9) {
      const float w = w3[k % 3];
      const __half hi = __float2half_rn(w);
      v = k < 6 ? __half2float(hi) : w - __half2float(hi);
    } else if (k < 11) {
      const __half hi = __float2half_rn(bias);
      v = k == 9 ? __half2float(hi) : bias - __half2float(hi);
    }
    blob[(size_t)tile * 128 * 64 + pf_swz_elem(n, cc)] = __float2half_rn(v);
  }
}

int pf_pack_block0_ct(const float* wp, const float* bp, const float* b0, const float* pfold, const float* cst0, void* ctblob) {
  pf_pack_ct_kernel<<<2, 256>>>(wp, bp, b0, pfold, cst0, reinterpret_cast<__half*>(ctblob));
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

// Persistent grid size of the scene-encoder kernels.  Their CTAs take a whole SM each (225 KB of shared memory), so a full
// grid excludes every other kernel while it runs; SEEME_PF_GRID < 148 leaves SMs to the other pipeline slots' kernels (the
// latency-bound sampler chain, VAE, SMPL), see DESIGN.md 4.4.
static int pf_grid_limit() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SEEME_PF_GRID");
    v = e ? atoi(e) : NUM_SMS;
    if (v < 1 || v > NUM_SMS) v = NUM_SMS;
  }
  return v;
}

// SEEME_PF_BLOCK0=cuda selects the CUDA-core generator (pointnet_block0_kernel) for A/B measurements
static bool pf_block0_use_tc() {
#ifdef SEEME_EXPERIMENTAL
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SEEME_PF_BLOCK0");
    v = (e && strcmp(e, "cuda") == 0) ? 0 : 1;
  }
  return v == 1;
#else
  return true;
#endif
}

int pf_block0_forward(const float* xyz, void* x_out, const void* w_blob, const void* wpb, const void* ctblob, const float* b0,
                      const float* cst0, const float* pfold, unsigned* colmax, int samples, int n_points, int prof_id, cudaStream_t s) {
  SEEME_TRY(pf_encoder());
  SEEME_REQUIRE(n_points >= 128, SEEME_EINVAL, "pf_block0_forward: needs >= 128 points per sample (got %d)", n_points);
  PfMaps maps;
  memset(&maps, 0, sizeof(maps));
  SEEME_TRY(pf_act_map(&maps.xout, x_out, samples, n_points));
  maps.xin = maps.xout;
  SEEME_TRY(pf_w_map(&maps.w, w_blob));
  const int tiles_per_sample = (n_points + 127) / 128, n_tiles = tiles_per_sample * samples;
  const int grid = n_tiles < pf_grid_limit() ? n_tiles : pf_grid_limit();
  static bool configured = false;
  if (!configured) {
#ifdef SEEME_EXPERIMENTAL
    SEEME_CUDA(cudaFuncSetAttribute(pointnet_block0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P0_SMEM));
#endif
    SEEME_CUDA(cudaFuncSetAttribute(pointnet_block0_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P0T_SMEM));
    configured = true;
  }
  ProfScope prof(prof_id - 1, s);
  if (pf_block0_use_tc()) {
    P0TArgs a;
    a.n_points = n_points; a.tiles_per_sample = tiles_per_sample; a.n_tiles = n_tiles;
    a.xyz = xyz;
    a.colmax = colmax;
    a.wblob = reinterpret_cast<const uint8_t*>(w_blob);
    a.ctblob = reinterpret_cast<const uint8_t*>(ctblob);
    pointnet_block0_tc_kernel<<<grid, P0T_THREADS, P0T_SMEM, s>>>(maps, a);
  } else {
#ifdef SEEME_EXPERIMENTAL
    P0Args a;
    a.n_points = n_points; a.tiles_per_sample = tiles_per_sample; a.n_tiles = n_tiles;
    a.xyz = xyz;
    a.wpb = reinterpret_cast<const float4*>(wpb);
    a.b0 = b0; a.cst0 = cst0;
    a.pfold = reinterpret_cast<const float4*>(pfold);
    a.colmax = colmax;
    a.wblob = reinterpret_cast<const uint8_t*>(w_blob);
    pointnet_block0_kernel<<<grid, PF_THREADS, P0_SMEM, s>>>(maps, a);
#else
    (void)wpb; (void)b0; (void)cst0; (void)pfold;
    SEEME_REQUIRE(false, SEEME_EINVAL, "the CUDA-core block-0 generator is compiled in experimental builds only (-DSEEME_EXPERIMENTAL)");
#endif
  }
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

int pf_pack_block(const float* ws, const float* w0, const float* w1, void* blob) {
  pf_pack_weights_kernel<<<PF_WCHUNKS, 256>>>(ws, w0, w1, reinterpret_cast<__half*>(blob));
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

#ifdef PF_TRACE
extern "C" int seeme_pf_trace_read(long long* host, int n) {
  const int m = n < PF_TRACE_TILES * PF_TRACE_SLOTS ? n : PF_TRACE_TILES * PF_TRACE_SLOTS;
  return cudaMemcpyFromSymbol(host, pf_trace, (size_t)m * sizeof(long long)) == cudaSuccess ? m : -1;
}
#endif

size_t pf_blob_bytes() { return (size_t)PF_WCHUNKS * 128 * 64 * 2; }

int pf_block_forward(const void* x_in, void* x_out, const void* w_blob, const float* bias_h, const float* bias_o, unsigned* colmax,
                     int samples, int n_points, int h_in_tmem, int prof_id, cudaStream_t s) {
  SEEME_TRY(pf_encoder());
  SEEME_REQUIRE(n_points >= 128, SEEME_EINVAL, "pf_block_forward: needs >= 128 points per sample (got %d)", n_points);
  PfMaps maps;
  memset(&maps, 0, sizeof(maps));
  SEEME_TRY(pf_act_map(&maps.xin, x_in, samples, n_points));
  if (x_out) SEEME_TRY(pf_act_map(&maps.xout, x_out, samples, n_points)); else maps.xout = maps.xin;
  SEEME_TRY(pf_w_map(&maps.w, w_blob));
  PfArgs a;
  a.n_points = n_points;
  a.tiles_per_sample = (n_points + 127) / 128;
  a.n_tiles = a.tiles_per_sample * samples;
  a.bias_h = bias_h; a.bias_o = bias_o; a.colmax = colmax;
  a.store_out = x_out != nullptr;
  a.wblob = reinterpret_cast<const uint8_t*>(w_blob);
#ifdef SEEME_EXPERIMENTAL
  static const int dbg = getenv("SEEME_PF_DEBUG_SKIP") ? atoi(getenv("SEEME_PF_DEBUG_SKIP")) : 0;
  static bool warned = false;
  if (dbg && !warned) {
    fprintf(stderr, "seeme_b200: SEEME_PF_DEBUG_SKIP=%d -- timing diagnostics only, the scene-encoder results are WRONG\n", dbg);
    warned = true;
  }
  a.debug_skip = dbg;
#else
  a.debug_skip = 0;
#endif
  static bool configured = false;
  if (!configured) {
    SEEME_CUDA(cudaFuncSetAttribute(pointnet_block_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM_BLK));
    SEEME_CUDA(cudaFuncSetAttribute(pointnet_block_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM_BLK));
    configured = true;
  }
#ifdef SEEME_EXPERIMENTAL
  static bool configured2 = false;
  if (!configured2) {
    SEEME_CUDA(cudaFuncSetAttribute(pointnet_block_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM_BLK));
    configured2 = true;
  }
#endif
  ProfScope prof(prof_id - 1, s);
  if (h_in_tmem == 2) {            // CTA pairs (cta_group::2): 74 clusters of 2
#ifdef SEEME_EXPERIMENTAL
    const int pairs = (a.n_tiles + 1) / 2 < NUM_SMS / 2 ? (a.n_tiles + 1) / 2 : NUM_SMS / 2;
    pointnet_block_pair_kernel<<<2 * pairs, PF_THREADS, PF_SMEM_BLK, s>>>(maps, a);
    SEEME_LAUNCH_CHECK();
    return SEEME_OK;
#else
    SEEME_REQUIRE(false, SEEME_EINVAL, "the CTA-pair scene-encoder kernel (precision 18) is compiled in experimental builds only (-DSEEME_EXPERIMENTAL)");
#endif
  }
  const int grid = a.n_tiles < pf_grid_limit() ? a.n_tiles : pf_grid_limit();
  if (h_in_tmem) pointnet_block_kernel<true><<<grid, PF_THREADS, PF_SMEM_BLK, s>>>(maps, a);
  else pointnet_block_kernel<false><<<grid, PF_THREADS, PF_SMEM_BLK, s>>>(maps, a);
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

}  // namespace seeme
