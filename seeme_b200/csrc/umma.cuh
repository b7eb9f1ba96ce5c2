// tcgen05 / TMEM / TMA / mbarrier primitives for sm_100a (inline PTX) and the generic
// "linear" GEMM built on them (umma_gemm.cu).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace seeme {

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU box.  The loop must NOT be unrolled and the
// timeout path stays out of line: the unrolled form (x4 with an inlined printf at every call site) was ~150 SASS
// instructions per wait and the instruction-cache misses showed up as no_inst stalls of the MMA-issuing warp.
static __device__ __noinline__ void mbar_wait_timeout(uint32_t addr) {
  printf("seeme_b200: mbarrier wait timed out (block %d,%d thread %d barrier 0x%x)\n", blockIdx.x, blockIdx.y, threadIdx.x, addr);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
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
  mbar_wait_timeout(addr);
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled TMA load global -> shared, completion on an mbarrier (c0 = innermost coordinate)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// multicast variant: the box lands at the same CTA-relative offset in every CTA of cta_mask and completes tx bytes on
// the mbarrier at the same CTA-relative offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(cta_mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (lane_base + i)
// same, arriving on the mbarrier at this CTA-relative offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile (rows of 64 bf16 = 128 B, 8-row groups 1024 B apart):
// start address >> 4 | LBO = 1 (ignored for swizzled K-major) | SBO = 1024 >> 4 | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// advance the 14-bit start-address field of a descriptor by off16 (units of 16 bytes; shared memory is < 256 KB: no carry)
__device__ __forceinline__ uint64_t umma_desc_add(uint64_t d, uint32_t off16) {
  return (d & 0xffffffff00000000ull) | (uint64_t)((uint32_t)d + off16);
}
// one lane of a converged warp (the same lane every time)
__device__ __forceinline__ bool umma_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// same with fp16 operands (a_format = b_format = 0): 10 mantissa bits instead of 7 at the same MMA rate
__host__ __device__ constexpr uint32_t umma_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ---- host-side description of one tcgen05 linear -----------------------------------------------
// D[M,N] = epilogue( [A1 | A2][M, K1+K2] . W[N, K1+K2]^T ), operands bf16 (optionally split hi+lo for
// near-fp32 accuracy: hi.hi + lo.hi + hi.lo), fp32 accumulation in TMEM.
struct UmmaOperand {
  const __nv_bfloat16* hi = nullptr;
  const __nv_bfloat16* lo = nullptr;   // null in single-pass mode
  int ld = 0;                          // row pitch in elements
};
struct UmmaLinear {
  UmmaOperand A1, A2, W;         // A2 optional (K2 = 0); W is [N, K1+K2]
  int M = 0, N = 0, K1 = 0, K2 = 0;
  const float* bias = nullptr;   // [N] or [G,N]
  int bias_group_rows = 0;
  const float* pfold = nullptr;  // [N,4] rank-3 fold: + P[n][0..2] . xyz[m]  (scene encoder block 0)
  const float* xyz = nullptr;    // [M,3]
  int act = ACT_NONE;            // applied to acc + bias (+ pfold)
  const float* R = nullptr;      // fp32 residual added after the activation (before it when act_after_residual)
  int ldr = 0;
  int act_after_residual = 0;    // out = act(acc + bias + R): the bottleneck tail of the image backbone
  int dense_ctas = 0;            // many-row, short-K, bandwidth-bound GEMMs (image backbone): 112-register build with a
                                 // one-stage ring so that three CTAs share an SM (the default builds fit two)
  float* Y = nullptr;            // fp32 output (nullable)
  int ldy = 0;
  __nv_bfloat16 *Yh = nullptr, *Yl = nullptr;   // bf16 (hi, lo) of the output (nullable)
  __nv_bfloat16 *Zh = nullptr, *Zl = nullptr;   // bf16 (hi, lo) of relu(output) (nullable)
  int ldb = 0;
  unsigned* colmax = nullptr;    // [G,N] order-preserving uint max over the rows of each group (nullable)
  int colmax_group_rows = 0;
  int prof_id = 0;
  int fp16 = 0;                  // operands and 16-bit outputs are IEEE fp16 (stored in the bf16-typed buffers); npass 1 only
};
// precision: 1 = single bf16 pass, 3 = split-bf16 (hi.hi + lo.hi + hi.lo)
int umma_linear(const UmmaLinear& g, int npass, cudaStream_t s);

// 2-D bf16 tensor map over a [rows, cols] matrix (pitch ld_elems), box = box_rows x 64 elements, SWIZZLE_128B, OOB = 0
int umma_tensor_map_bf16(CUtensorMap* map, const void* ptr, int rows, int cols, int ld_elems, int box_rows);

// ---- convenience layer used by the transformer stacks ----------------------------------------------
struct PackedLinear {           // nn.Linear weight [N,K] packed once as bf16 (hi, lo); bias stays fp32
  __nv_bfloat16 *hi = nullptr, *lo = nullptr;
  const float* bias = nullptr;
  int N = 0, K = 0;
};
struct ActBuf {                 // an activation tensor [rows, ld]: fp32 copy and/or bf16 (hi, lo) copy
  float* f = nullptr;
  __nv_bfloat16 *h = nullptr, *l = nullptr;
  int ld = 0;
};
int pack_linear(Arena& arena, PackedLinear& out, const float* W, int ldw, int N, int K, const float* bias);
// out = act(A W^T + bias) (+ R); A2 (nullable) supplies the second half of K (torch.cat([A, A2], -1));
// the fp32 and bf16 members of `out` that are non-null are written
int run_linear(const PackedLinear& W, const ActBuf& A, const ActBuf* A2, int M, int act, const float* R, int ldr,
               const ActBuf& out, int npass, cudaStream_t s, int bias_group_rows = 0, const float* bias_override = nullptr);

// fp32 -> bf16 hi (+ lo = bf16(x - hi)) conversion of a [rows, cols] matrix (pitch ldx) into columns [0, cols)
// of a destination with pitch ld_out (the other destination columns are not written), optional relu on load
int to_bf16_split(const float* x, int ldx, int rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld_out, int relu,
                  cudaStream_t s);

}  // namespace seeme
