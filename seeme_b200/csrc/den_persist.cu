// Persistent samplers: the whole 50-step DDIM loop of MLD._diffusion_reverse (mld/models/modeltype/mld.py:432-511)
// over the skip-connected denoiser (mld_denoiser.py:151-244, cross_attention.py:67-83, mdiff_transformer.py:286-304)
// as ONE kernel launch.  Two kernels: den_persist_kernel (an 8-CTA cluster per 128-row tile, described here; the default)
// and den_mono_kernel (one CTA per tile, further down).  Measurements and what bounds each: DESIGN.md 4.2.
//
// Decomposition.  A cluster of CL = 8 CTAs owns a tile of 128 denoiser rows (under CFG: 64 latents x {uncond, cond}) for
// the whole run; clusters never talk to one another.  Inside a cluster every GEMM is split over its OUTPUT features: CTA c
// computes columns [32c, 32c+32) of every 256-wide activation (its "slice") on tcgen05 (M = 128 rows, N = 16..64 per
// instruction, fp32 accumulators in tensor memory, split-bf16 x3 operands).  TMEM, mbarriers and the
// per-row state (residual stream slice, latent slice) live across all ~3 350 GEMM units of a run.
//
//   warp 0      producer: streams this CTA's weight "tape" (pre-swizzled bf16 hi|lo blobs in consumption order, one
//               cp.async.bulk per K-block, ring of 4 x 16 KB) and the A operand K-blocks (one 32 KB bulk copy each from the
//               cluster's exchange images in L2, ring of 4 x (16 KB hi + 16 KB lo) = one full K = 256 activation)
//   warp 1      one elected lane issues tcgen05.mma; tcgen05.commit releases ring slots and signals the epilogue
//   warps 2-5   epilogue, thread = row: tcgen05.ld, bias / attention / LayerNorm / FiLM / CFG + DDIM on its 32-column
//               slice, then publishes the slice (bf16 hi, lo) into the exchange matrix for the next GEMM's A operand
//
// Exchanges (all mbarrier based, double buffered; no barrier.cluster in the loop):
//   * activation all-gather through L2: st.global of the slice -> fence -> remote mbarrier arrive on every CTA of the
//     cluster -> the producers bulk-load the full [128 x 256] activation.  FFN 256 -> 1024 -> 256 is K-split instead:
//     each CTA keeps its 128 hidden units private (its epilogue writes relu(h) straight into the A ring) and the 8
//     partial [128 x 256] results are reduce-scattered through a fp32 scratch (one exchange instead of a 512 KB all-gather).
//   * row statistics through distributed shared memory: every reduction over the 256 features of a row (the 4-key
//     attention logits, softmax_d of the linear cross-attention, every LayerNorm as a (mean, M2) Chan combine) sends
//     <= 4 floats per row to each peer with st.async (...mbarrier::complete_tx): data and signal in one instruction.
//
// Algebra: H1 token-0-only self-attention, H2 linear attention as dot products, H3 cond-token K/V once
// per run (on tensor cores, see den_persist_run), H4 per-timestep tables; plus the out-projection of the self-attention
// folded into the value projection (W_ov = W_o W_v, computed in fp64 at create): sa = sum_j p_j (W_ov t_j + W_o b_v) + b_o.
#include "denoiser.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace seeme {

constexpr int DP_CL = 8;               // CTAs per cluster
constexpr int DP_NS = 256 / DP_CL;     // slice width of a 256-wide activation
constexpr int DP_HS = 1024 / DP_CL;    // hidden units of the 1024-wide FFN per CTA
constexpr int DP_GS = 128 / DP_CL;     // hidden units of the 128-wide FFN per CTA
constexpr int DP_THREADS = 192;
constexpr int DP_NA = 4, DP_NW = 4;
constexpr int DP_A_SLOT = 32768, DP_W_SLOT = 16384;
constexpr int DP_STAT_FLOATS = 2 * DP_CL * 4 * 128;
constexpr int DP_SMEM = DP_NA * DP_A_SLOT + DP_NW * DP_W_SLOT + DP_STAT_FLOATS * 4 + 1024;
static_assert(DP_NS == 32, "the epilogues are written for 32-column slices (CL = 8)");
// columns of the per-cluster exchange matrix [128 rows x XC_COLS] (bf16 hi and lo copies)
enum { XC_P0 = 0, XC_P1 = 256, XC_L0 = 512, XC_L1 = 768, XC_XA = 1024, XC_LN = 1280, XC_HB0 = 1536, XC_HB1 = 1792,
       XC_X3 = 2048, XC_G = 2304, XC_FF = 2432, XC_COLS = 3456 };
constexpr int XC_KB = XC_COLS / 64;     // K-blocks (64 columns) per tile
constexpr int TM_XRES = 256, TM_LAT = 256 + DP_NS;   // tensor-memory columns of the per-row state

struct DpUnit {             // one GEMM unit: acc[:, acc_col : acc_col + n] = A[128, 64 nkb] . W_unit^T
  int a_col, a_col2;        // exchange-matrix column of K-block 0 / of K-block nkb1 (torch.cat([x, skip]))
  int a_rstride;            // + rank * a_rstride (CTA-private FFN hidden units)
  int nkb1, nkb, n, acc_col;
  int wait;                 // 0: A visible already, 1: cluster exchange, 2: CTA-local hand-over
  int reuse_a, release_a;   // sub-units of one stage share the A ring contents
  int last;                 // the stage's epilogue runs after this unit
  int cat;                  // 1: hi.hi and hi.lo in ONE instruction (B = [W_hi; W_lo], N = 2n: the lo tile follows the hi tile in the
                            //    blob); the hi.lo products land in columns [acc_col + n, acc_col + 2n) and the epilogue adds them
  unsigned w_off;           // byte offset of the unit's first blob in the CTA's tape (blob = n x 256 bytes)
};
constexpr int DP_MAX_UNITS = 80;

struct DpLayerP {
  // bqkv [768]: b_q / 16 | b_k | b_o + W_o b_v (the softmax weights sum to 1); the FiLM LayerNorm affines live in the film tables
  const float *bqkv, *n1g, *n1b, *b1, *b2, *n2g, *n2b, *cng, *cnb, *bcaq, *bcaout, *bf1, *bf2, *bfout;
  // [steps][512]: (k | ov) of the time token; FiLM folded into the LayerNorm affine: (g (1 + scale) | b (1 + scale) + shift)
  const float *kt, *film_ca, *film_ff;
  const float* ctab;                     // [tile][4][Nc][256][128 rows]: k, ov (self-attention), softmax_n(key), value (cross-attention)
};
struct DpParams {
  DpLayerP L[5];
  const float *skip_b[2], *fng, *fnb, *pe0;
  const uint8_t* tape;
  unsigned long long tape_cta_bytes;
  const DpUnit* units;
  int n_units;
  uint8_t* ximg;        // exchange matrix as shared-memory tile images: [tile][XC_KB K-blocks][hi 16 KB | lo 16 KB]
  float* part;          // [tiles][CL][256][128]
  const float* x_in;
  float* out;
  const float *coef, *gscale;
  int B, R, cfg, n_steps, mode, rows_pad;
  unsigned long long* trace;   // diagnostics (SEEME_DP_TRACE): 3 roles x 4096 (clock64 << 8 | event) of cluster 0 / rank 0 at trace_step
  int trace_step;
};
#define DP_TR(role, code)                                                                                       \
  do {                                                                                                          \
    if (p.trace && blockIdx.x == 0 && step == p.trace_step && tr_n < 4096)                                      \
      p.trace[(role) * 4096 + tr_n++] = ((unsigned long long)clock64() << 8) | (unsigned long long)(code);      \
  } while (0)

// ---- PTX helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dp_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void dp_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// relaxed flavour: the caller has already ordered its writes with ONE release (every mbarrier.arrive.release.cluster costs a
// MEMBAR.ALL.GPU -- eight of them per signal were ~5 000 cycles in the first trace)
__device__ __forceinline__ void dp_arrive_remote_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 16 bytes into a peer's shared memory, completing 16 tx bytes on the peer's mbarrier: data and signal travel together,
// no fence on either side
__device__ __forceinline__ void dp_st_async4(uint32_t cluster_addr, uint32_t cluster_bar, float a, float b, float c, float d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%2, %3, %4, %5}, [%1];" ::"r"(cluster_addr),
               "r"(cluster_bar), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d))
               : "memory");
}
static __device__ __noinline__ void dp_wait_timeout(uint32_t addr, int what) {
  printf("seeme_b200: den_persist wait timed out (block %d thread %d barrier 0x%x kind %d)\n", blockIdx.x, threadIdx.x, addr, what);
  __trap();
}
// wait with cluster-scope acquire (arrivals come from peer CTAs)
__device__ __forceinline__ void dp_wait_cluster(uint64_t* bar, uint32_t parity, int what) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 22); ++it) {
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
  dp_wait_timeout(addr, what);
}
__device__ __forceinline__ void dp_wait(uint64_t* bar, uint32_t parity, int what) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 22); ++it) {
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
  dp_wait_timeout(addr, what);
}
__device__ __forceinline__ void dp_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dp_fence_proxy_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait_dp() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void dp_ld32(uint32_t taddr, float (&f)[32]) {
  uint32_t raw[32];
  tmem_ld32(taddr, raw);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(raw[i]);
}
// accumulator of a `cat` unit: main columns + the hi.lo cross-term columns
__device__ __forceinline__ void dp_ld32x2(uint32_t ta, uint32_t tb, float (&f)[32]) {
  uint32_t r0[32];
  tmem_ld32(ta, r0);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r0[i]);
  tmem_ld32(tb, r0);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] += __uint_as_float(r0[i]);
}
__device__ __forceinline__ void dp_st32(uint32_t taddr, const float (&f)[32]) {
  uint32_t raw[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(f[i]);
  tmem_st32(taddr, raw);
  tmem_st_wait_dp();
}
__device__ __forceinline__ void dp_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// publish N (16 or 32) consecutive columns [col, col + N) of this thread's row into the tile's exchange images: the 16-byte
// chunks go to their SWIZZLE_128B positions, so a K-block (hi image | lo image, 32 KB) is ONE contiguous bulk copy away
// from being an A operand
template <int N>
__device__ __forceinline__ void dp_publish(uint8_t* ximg_tile, int col, int row, const float* f) {
  uint32_t hb[N / 2], lb[N / 2];
#pragma unroll
  for (int i = 0; i < N / 2; ++i) dp_split2(f[2 * i], f[2 * i + 1], hb[i], lb[i]);
  uint8_t* line = ximg_tile + (size_t)(col >> 6) * 32768 + row * 128;
  const int j0 = (col & 63) >> 3, sw = row & 7;
#pragma unroll
  for (int j = 0; j < N / 8; ++j) {
    const int c = ((j0 + j) ^ sw) << 4;
    *reinterpret_cast<uint4*>(line + c) = make_uint4(hb[4 * j], hb[4 * j + 1], hb[4 * j + 2], hb[4 * j + 3]);
    *reinterpret_cast<uint4*>(line + 16384 + c) = make_uint4(lb[4 * j], lb[4 * j + 1], lb[4 * j + 2], lb[4 * j + 3]);
  }
}
// 32 consecutive columns [c0, c0 + 32) of this thread's row into the shared-memory A operand (bf16 hi | lo K-block tiles)
__device__ __forceinline__ void dm_write_a(uint8_t* a_buf, int c0, int row, const float (&f)[32]) {
  uint32_t hb[16], lb[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dp_split2(f[2 * i], f[2 * i + 1], hb[i], lb[i]);
  uint8_t* line = a_buf + (c0 >> 6) * 32768 + row * 128;
  const int j0 = (c0 & 63) >> 3, sw = row & 7;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = ((j0 + j) ^ sw) << 4;
    *reinterpret_cast<uint4*>(line + c) = make_uint4(hb[4 * j], hb[4 * j + 1], hb[4 * j + 2], hb[4 * j + 3]);
    *reinterpret_cast<uint4*>(line + 16384 + c) = make_uint4(lb[4 * j], lb[4 * j + 1], lb[4 * j + 2], lb[4 * j + 3]);
  }
}

// 32 consecutive floats of a vector shared by all rows (bias, LayerNorm affine, FiLM, time-token tables)
__device__ __forceinline__ void dp_ldvec(const float* __restrict__ p, float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + j);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}

// 32-element reductions with four independent accumulators: the epilogue runs ONE warp per scheduler, so a serial chain of 32
// dependent FMAs (4 cycles each) is pure latency
__device__ __forceinline__ float dp_sum32(const float (&a)[32]) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) { s0 += a[i]; s1 += a[i + 1]; s2 += a[i + 2]; s3 += a[i + 3]; }
  return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ float dp_dot32(const float (&a)[32], const float (&b)[32]) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    s0 = fmaf(a[i], b[i], s0); s1 = fmaf(a[i + 1], b[i + 1], s1); s2 = fmaf(a[i + 2], b[i + 2], s2); s3 = fmaf(a[i + 3], b[i + 3], s3);
  }
  return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ float dp_sqdev32(const float (&a)[32], float mean) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float d0 = a[i] - mean, d1 = a[i + 1] - mean, d2 = a[i + 2] - mean, d3 = a[i + 3] - mean;
    s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2, d2, s2); s3 = fmaf(d3, d3, s3);
  }
  return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ float dp_max32(const float (&a)[32]) {
  float m0 = a[0], m1 = a[1], m2 = a[2], m3 = a[3];
#pragma unroll
  for (int i = 4; i < 32; i += 4) { m0 = fmaxf(m0, a[i]); m1 = fmaxf(m1, a[i + 1]); m2 = fmaxf(m2, a[i + 2]); m3 = fmaxf(m3, a[i + 3]); }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// per-row statistics exchange over distributed shared memory
struct DpStat {
  float* stats;      // [2][CL][4][128]
  uint64_t* bar;     // [2][4]
  uint32_t cnt;
  int q, lane, row, rank;
  const float* cur;
  unsigned long long* tr;     // diagnostics: (clock64 << 8 | 7) before / 8 after each exchange
  template <int K>
  __device__ __forceinline__ void send(const float (&v)[K]) {
    static_assert(K <= 4, "at most 4 values per round");
    if (tr) *tr++ = ((unsigned long long)clock64() << 8) | 7ull;
    const uint32_t par = cnt & 1u;
    // stats[par][src][row][4]; this thread's 16 bytes land in every CTA of the cluster (its own included) and complete
    // 16 tx bytes on that CTA's barrier of this warp's lane quarter: 32 rows x CL sources x 16 B per phase
    const uint32_t laddr = smem_u32(stats + ((par * DP_CL + rank) * 128 + row) * 4);
    const uint32_t lbar = smem_u32(&bar[par * 4 + q]);
    if (lane == 0) mbar_arrive_expect_tx(&bar[par * 4 + q], 32 * DP_CL * 16);
    const float v0 = v[0], v1 = K > 1 ? v[K > 1 ? 1 : 0] : 0.f, v2 = K > 2 ? v[K > 2 ? 2 : 0] : 0.f, v3 = K > 3 ? v[K > 3 ? 3 : 0] : 0.f;
#pragma unroll
    for (int pr = 0; pr < DP_CL; ++pr) dp_st_async4(dp_mapa(laddr, pr), dp_mapa(lbar, pr), v0, v1, v2, v3);
    dp_wait(&bar[par * 4 + q], (cnt >> 1) & 1u, 10);
    cur = stats + ((par * DP_CL) * 128 + row) * 4;
    ++cnt;
    if (tr) *tr++ = ((unsigned long long)clock64() << 8) | 8ull;
  }
  __device__ __forceinline__ float get(int s, int k) const { return cur[s * 512 + k]; }
  // v[k] <- sum over the cluster
  template <int K>
  __device__ __forceinline__ void sum(float (&v)[K]) {
    constexpr int K0 = K > 4 ? 4 : K;
    {
      float a[K0];
#pragma unroll
      for (int k = 0; k < K0; ++k) a[k] = v[k];
      send<K0>(a);
#pragma unroll
      for (int k = 0; k < K0; ++k) {
        float t = 0.f;
#pragma unroll
        for (int s = 0; s < DP_CL; ++s) t += get(s, k);
        v[k] = t;
      }
    }
    if constexpr (K > 4) {
      float a[K - 4];
#pragma unroll
      for (int k = 0; k < K - 4; ++k) a[k] = v[4 + k];
      send<K - 4>(a);
#pragma unroll
      for (int k = 0; k < K - 4; ++k) {
        float t = 0.f;
#pragma unroll
        for (int s = 0; s < DP_CL; ++s) t += get(s, k);
        v[4 + k] = t;
      }
    }
  }
  // LayerNorm(256) of a row whose slice is t[32]: Chan combine of the per-slice (mean, M2); biased variance, eps 1e-5
  __device__ __forceinline__ void layernorm(float (&t)[32], const float* __restrict__ g, const float* __restrict__ b) {
    const float mc = dp_sum32(t) * (1.0f / 32.0f);
    const float m2 = dp_sqdev32(t, mc);
    const float a[2] = {mc, m2};
    float gg[32], bb[32];          // requested before the exchange so that their L2 round trip overlaps it
    dp_ldvec(g, gg);
    dp_ldvec(b, bb);
    send<2>(a);
    float mean = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < DP_CL; ++s2) mean += get(s2, 0);
    mean *= (1.0f / DP_CL);
    float M2 = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < DP_CL; ++s2) { const float d = get(s2, 0) - mean; M2 += get(s2, 1) + 32.0f * d * d; }
    const float rstd = rsqrtf(M2 * (1.0f / 256.0f) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 32; ++i) t[i] = (t[i] - mean) * rstd * gg[i] + bb[i];
  }
};

template <int NC>
__global__ void __cluster_dims__(DP_CL, 1, 1) __launch_bounds__(DP_THREADS, 1) den_persist_kernel(const __grid_constant__ DpParams p) {
  extern __shared__ __align__(1024) uint8_t dp_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dp_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* w_ring = smem + DP_NA * DP_A_SLOT;
  float* stats = reinterpret_cast<float*>(w_ring + DP_NW * DP_W_SLOT);
  __shared__ __align__(8) uint64_t full_a[DP_NA], empty_a[DP_NA], full_w[DP_NW], empty_w[DP_NW], acc_full, xbar[2], lbar[2], sbar[8], rbar[8];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int tile = blockIdx.x / DP_CL;

  if (threadIdx.x == 0) {
    for (int i = 0; i < DP_NA; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < DP_NW; ++i) { mbar_init(&full_w[i], 1); mbar_init(&empty_w[i], 1); }
    mbar_init(&acc_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&xbar[i], 4 * DP_CL); mbar_init(&lbar[i], 4); }
    for (int i = 0; i < 8; ++i) { mbar_init(&sbar[i], 1); mbar_init(&rbar[i], DP_CL); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();      // every CTA's barriers exist before the first remote arrive
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ---- producer -----------------------------------------------------------------------------------
    if (lane == 0) {
      uint32_t ia = 0, iw = 0, nx = 0, nl = 0;
      int tr_n = 0;
      const uint8_t* tape = p.tape + (size_t)rank * p.tape_cta_bytes;
      const uint8_t* ximg_tile = p.ximg + (size_t)tile * XC_KB * 32768;
      for (int step = 0; step < p.n_steps; ++step) {
        for (int u = 0; u < p.n_units; ++u) {
          const DpUnit un = p.units[u];
          const uint32_t wbytes = (uint32_t)un.n * 256u;
          const uint8_t* wsrc = tape + un.w_off;
          int kw = 0;
          DP_TR(0, 1);
          // the weights do not depend on the exchange: up to NW K-blocks travel while the previous epilogue runs
          for (; kw < un.nkb && kw < DP_NW; ++kw) {
            const uint32_t sw = iw % DP_NW;
            dp_wait(&empty_w[sw], ((iw / DP_NW) & 1u) ^ 1u, 1);
            mbar_arrive_expect_tx(&full_w[sw], wbytes);
            dp_bulk_load(w_ring + sw * DP_W_SLOT, wsrc + (size_t)kw * wbytes, wbytes, &full_w[sw]);
            ++iw;
          }
          if (un.wait == 1) { dp_wait(&xbar[nx & 1u], (nx >> 1) & 1u, 2); ++nx; dp_fence_proxy_all(); }
          else if (un.wait == 2) { dp_wait(&lbar[nl & 1u], (nl >> 1) & 1u, 3); ++nl; dp_fence_proxy_all(); }
          const int acol = un.a_col + rank * un.a_rstride;
          DP_TR(0, 2);
          for (int kb = 0; kb < un.nkb; ++kb) {
            if (!un.reuse_a) {
              const uint32_t sa = ia % DP_NA;
              if (un.wait == 2) {
                // CTA-private operand (relu(h) of this CTA's FFN hidden units): the epilogue threads have written it straight
                // into the ring slots (always slots 0, 1: the K-block count of a step is a multiple of the ring size up to
                // here), so there is nothing to load -- only the slot's phase to complete
                mbar_arrive(&full_a[sa]);
              } else {
                dp_wait(&empty_a[sa], ((ia / DP_NA) & 1u) ^ 1u, 4);
                mbar_arrive_expect_tx(&full_a[sa], DP_A_SLOT);
                const int col = kb < un.nkb1 ? acol + kb * 64 : un.a_col2 + (kb - un.nkb1) * 64;
                dp_bulk_load(a_ring + sa * DP_A_SLOT, ximg_tile + (size_t)(col >> 6) * 32768, DP_A_SLOT, &full_a[sa]);
              }
              ++ia;
            }
            if (kb >= kw) {
              const uint32_t sw = iw % DP_NW;
              dp_wait(&empty_w[sw], ((iw / DP_NW) & 1u) ^ 1u, 5);
              mbar_arrive_expect_tx(&full_w[sw], wbytes);
              dp_bulk_load(w_ring + sw * DP_W_SLOT, wsrc + (size_t)kb * wbytes, wbytes, &full_w[sw]);
              ++iw;
            }
          }
          DP_TR(0, 3);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer -----------------------------------------------------------------------------------
    uint32_t ia = 0, iw = 0, a0 = 0;
    int tr_n = 0;
    const uint64_t da_base = umma_desc_k128(smem_u32(a_ring));
    const uint64_t dw_base = umma_desc_k128(smem_u32(w_ring));
    for (int step = 0; step < p.n_steps; ++step) {
      for (int u = 0; u < p.n_units; ++u) {
        const DpUnit un = p.units[u];
        if (!un.reuse_a) a0 = ia;
        const uint32_t idesc = umma_idesc_bf16(un.n), idesc2 = umma_idesc_bf16(2 * un.n);
        const uint32_t wl16 = (uint32_t)(un.n * 128) >> 4;      // lo tile follows the hi tile
        for (int kb = 0; kb < un.nkb; ++kb) {
          const uint32_t ai = a0 + (uint32_t)kb;
          const uint32_t sa = ai % DP_NA;
          if (!un.reuse_a) dp_wait(&full_a[sa], (ai / DP_NA) & 1u, 6);
          const uint32_t sw = iw % DP_NW;
          dp_wait(&full_w[sw], (iw / DP_NW) & 1u, 7);
          tc_fence_after();
          if (lane == 0 && kb == 0) DP_TR(1, 4);
          if (umma_elect_one()) {
            const uint64_t da0 = umma_desc_add(da_base, sa * (DP_A_SLOT >> 4));
            const uint64_t dw0 = umma_desc_add(dw_base, sw * (DP_W_SLOT >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = umma_desc_add(da0, k * 2), dw = umma_desc_add(dw0, k * 2);
              // a tcgen05.mma with N <= 64 costs ~73 cycles whatever N is (A-operand read): two instructions instead of three
              if (un.cat) {
                umma_bf16(tmem_base + (uint32_t)un.acc_col, da, dw, idesc2, (kb | k) != 0);
                umma_bf16(tmem_base + (uint32_t)un.acc_col, umma_desc_add(da, 16384 >> 4), dw, idesc, 1);
              } else {
                umma_bf16(tmem_base + (uint32_t)un.acc_col, da, dw, idesc, (kb | k) != 0);
                umma_bf16(tmem_base + (uint32_t)un.acc_col, umma_desc_add(da, 16384 >> 4), dw, idesc, 1);
                umma_bf16(tmem_base + (uint32_t)un.acc_col, da, umma_desc_add(dw, wl16), idesc, 1);
              }
            }
            umma_commit(&empty_w[sw]);
            if (un.release_a) umma_commit(&empty_a[sa]);
            if (un.last && kb == un.nkb - 1) umma_commit(&acc_full);
          }
          __syncwarp();
          ++iw;
        }
        if (lane == 0) DP_TR(1, 5);
        if (!un.reuse_a) ia += (uint32_t)un.nkb;
      }
    }
  } else {
    // ---- epilogue: thread = row ------------------------------------------------------------------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const int S = rank * DP_NS;
    const int prow = tile * 128 + row;
    int latrow, grow;
    bool valid;
    if (p.cfg) {
      latrow = tile * 64 + (row & 63);
      valid = latrow < p.B;
      grow = (row >> 6) * p.B + latrow;
    } else {
      latrow = prow;
      grow = prow;
      valid = prow < p.R;
    }
    uint8_t* const xt = p.ximg + (size_t)tile * XC_KB * 32768;
    // per-row cond-token tables: ONE base pointer per layer, every (table, token, column) at a compile-time offset
    const size_t ct_off = ((size_t)tile * 4 * NC * 256 + S) * 128 + row;
#define DP_CT(w, n, i) (((w) * NC + (n)) * 256 + (i)) * 128
    uint32_t nstage = 0, nx = 0, nl = 0, nr = 0;
    int tr_n = 0, step = 0;
    const bool tr_me = warp == 2 && lane == 0;
    DpStat ex;
    ex.stats = stats; ex.bar = sbar; ex.cnt = 0; ex.q = q; ex.lane = lane; ex.row = row; ex.rank = rank; ex.cur = stats; ex.tr = nullptr;

    auto acc_wait = [&]() {
      dp_wait(&acc_full, nstage & 1u, 8);
      ++nstage;
      tc_fence_after();
      if (tr_me) DP_TR(2, 6);
    };
    auto signal_x = [&]() {      // the slice just written is part of the next A operand of every CTA of the cluster
      if (tr_me) DP_TR(2, 9);
      dp_fence_proxy_all();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // ONE release (MEMBAR.ALL.GPU: this warp's slice stores are performed at the L2) covers all eight arrives
        const uint32_t b = smem_u32(&xbar[nx & 1u]);
        dp_arrive_remote(dp_mapa(b, 0));
#pragma unroll
        for (int pr = 1; pr < DP_CL; ++pr) dp_arrive_remote_relaxed(dp_mapa(b, pr));
      }
      ++nx;
      if (tr_me) DP_TR(2, 10);
    };
    auto signal_l = [&]() {      // CTA-private hand-over (FFN hidden units)
      dp_fence_proxy_all();        // st.shared (generic proxy) -> tcgen05.mma operand reads (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&lbar[nl & 1u]);
      ++nl;
    };

    {   // initial state: x = latents (or the given sample) + learned PE row 0 (mld_denoiser.py:210)
      float x[32], pe[32];
      dp_ldvec(p.pe0 + S, pe);
      if (valid) {
        const float* src = p.x_in + (size_t)(p.mode == 0 ? latrow : grow) * 256 + S;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(src) + j);
          x[4 * j] = t.x; x[4 * j + 1] = t.y; x[4 * j + 2] = t.z; x[4 * j + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = 0.f;
      }
      dp_st32(tl + TM_LAT, x);
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] += pe[i];
      dp_st32(tl + TM_XRES, x);
      dp_publish<32>(xt, XC_P0 + S, row, x);
      signal_x();
    }

    for (step = 0; step < p.n_steps; ++step) {
      const size_t toff = (size_t)step * 512;
      ex.tr = (p.trace && blockIdx.x == 0 && tr_me && step == p.trace_step) ? p.trace + 3 * 4096 : nullptr;
#pragma unroll 1
      for (int l = 0; l < 5; ++l) {
        const DpLayerP& L = p.L[l];
        const float* __restrict__ ct = L.ctab + ct_off;
        const int xout_col = l == 0 ? XC_L0 : l == 1 ? XC_L1 : XC_P1;
        if (l >= 3) {
          // x = Linear(cat[x, skip]) (cross_attention.py:77-80)
          float x[32], b[32];
          dp_ldvec(p.skip_b[l - 3] + S, b);
          acc_wait();
          dp_ld32x2(tl, tl + 32, x);
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] += b[i];
          dp_st32(tl + TM_XRES, x);
          dp_publish<32>(xt, XC_P0 + S, row, x);
          signal_x();
        }
        // ---- self-attention over {x, cond tokens, time token}, query = token 0 (mdiff_transformer.py:291-297) ----
        {
          float pr[NC + 2];
          {
            // everything that does not depend on the accumulator is requested while the GEMM runs
            float bq[32], bk[32], ktk[32], kc[NC < 2 ? NC : 2][32];
            dp_ldvec(L.bqkv + S, bq);
            dp_ldvec(L.bqkv + 256 + S, bk);
            dp_ldvec(L.kt + toff + S, ktk);
#pragma unroll
            for (int n = 0; n < NC && n < 2; ++n) {
#pragma unroll
              for (int i = 0; i < 32; ++i) kc[n][i] = __ldg(ct + DP_CT(0, n, i));
            }
            acc_wait();
            float qv[32];
            dp_ld32x2(tl, tl + 64, qv);
#pragma unroll
            for (int i = 0; i < 32; ++i) qv[i] += bq[i];
            pr[NC + 1] = dp_dot32(qv, ktk);
            {
              float kv[32];
              dp_ld32x2(tl + 32, tl + 96, kv);
#pragma unroll
              for (int i = 0; i < 32; ++i) kv[i] += bk[i];
              pr[0] = dp_dot32(qv, kv);
            }
#pragma unroll
            for (int n = 0; n < NC; ++n) {
              if (n < 2) {
                pr[1 + n] = dp_dot32(qv, kc[n < 2 ? n : 0]);
              } else {
                float t[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) t[i] = __ldg(ct + DP_CT(0, n, i));
                pr[1 + n] = dp_dot32(qv, t);
              }
            }
          }
          asm volatile("" ::: "memory");   // keep the loads below from being hoisted into the first half (register pressure)
          // operands of the second half, requested before the exchange (their L2 round trip overlaps it): merged bias
          // b_o + W_o b_v (the softmax weights sum to 1), ov of the cond tokens and of the time token
          float bs[32], ovc[NC][32], ovt[32];
          dp_ldvec(L.bqkv + 512 + S, bs);
#pragma unroll
          for (int n = 0; n < NC; ++n) {
#pragma unroll
            for (int i = 0; i < 32; ++i) ovc[n][i] = __ldg(ct + DP_CT(1, n, i));
          }
          dp_ldvec(L.kt + toff + 256 + S, ovt);
          ex.sum<NC + 2>(pr);
          {
            float m = pr[0];
#pragma unroll
            for (int j = 1; j < NC + 2; ++j) m = fmaxf(m, pr[j]);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < NC + 2; ++j) { pr[j] = __expf(pr[j] - m); sum += pr[j]; }
            const float inv = 1.0f / sum;
#pragma unroll
            for (int j = 0; j < NC + 2; ++j) pr[j] *= inv;
          }
          float t0[32];
          dp_ld32x2(tl + 128, tl + 160, t0);
#pragma unroll
          for (int i = 0; i < 32; ++i) t0[i] = bs[i] + pr[0] * t0[i] + pr[NC + 1] * ovt[i];
#pragma unroll
          for (int n = 0; n < NC; ++n) {
#pragma unroll
            for (int i = 0; i < 32; ++i) t0[i] = fmaf(pr[1 + n], ovc[n][i], t0[i]);
          }
          asm volatile("" ::: "memory");
          {
            float xr[32];
            dp_ld32(tl + TM_XRES, xr);
#pragma unroll
            for (int i = 0; i < 32; ++i) t0[i] += xr[i];
          }
          ex.layernorm(t0, L.n1g + S, L.n1b + S);      // x1 = norm1(x + sa)
          dp_st32(tl + TM_XRES, t0);
          dp_publish<32>(xt, XC_XA + S, row, t0);
          signal_x();
        }
        // ---- FFN 256 -> 1024 (ReLU) -> 256, K-split over the cluster ----
        {
          acc_wait();
#pragma unroll 1
          for (int ch = 0; ch < DP_HS / 32; ++ch) {
            float f[32], b[32];
            dp_ldvec(L.b1 + rank * DP_HS + ch * 32, b);
            dp_ld32x2(tl + (ch >> 1) * 128 + (ch & 1) * 32, tl + (ch >> 1) * 128 + (ch & 1) * 32 + 64, f);
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i] + b[i], 0.f);
            dm_write_a(a_ring, ch * 32, row, f);      // K-blocks 0, 1 of the K-split second GEMM: ring slots 0, 1 (see the producer)
          }
          signal_l();
        }
        {
          acc_wait();
          float* mine = p.part + ((size_t)(tile * DP_CL + rank) * 256) * 128 + row;
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float f[32];
            dp_ld32(tl + ch * 32, f);
#pragma unroll
            for (int i = 0; i < 32; ++i) __stcg(mine + (size_t)(ch * 32 + i) * 128, f[i]);
          }
          tc_fence_before();
          __syncwarp();
          const uint32_t par = nr & 1u;
          if (lane == 0) {
            const uint32_t b = smem_u32(&rbar[par * 4 + q]);
            dp_arrive_remote(dp_mapa(b, 0));
#pragma unroll
            for (int pr2 = 1; pr2 < DP_CL; ++pr2) dp_arrive_remote_relaxed(dp_mapa(b, pr2));
          }
          dp_wait(&rbar[par * 4 + q], (nr >> 1) & 1u, 9);
          ++nr;
          float t1[32], v[32];
          dp_ld32(tl + TM_XRES, t1);
          dp_ldvec(L.b2 + S, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) t1[i] += v[i];
#pragma unroll
          for (int s = 0; s < DP_CL; ++s) {
            const float* src = p.part + ((size_t)(tile * DP_CL + s) * 256 + S) * 128 + row;
#pragma unroll
            for (int i = 0; i < 32; ++i) t1[i] += __ldcg(src + (size_t)i * 128);
          }
          ex.layernorm(t1, L.n2g + S, L.n2b + S);      // x2 = norm2(x1 + ffn)
          dp_st32(tl + TM_XRES, t1);
          ex.layernorm(t1, L.cng + S, L.cnb + S);      // input norm of the cross-attention
          dp_publish<32>(xt, XC_LN + S, row, t1);
          signal_x();
        }
        // ---- linear cross-attention to the cond tokens + FiLM (mdiff_transformer.py:219-239, 152-163) ----
        {
          float y[32];
          {
            float qv[32], v[32];
            dp_ldvec(L.bcaq + S, v);
            acc_wait();
            dp_ld32x2(tl, tl + 32, qv);
#pragma unroll
            for (int i = 0; i < 32; ++i) qv[i] += v[i];
            const float m = dp_max32(qv);
            float st[2 + NC];
            st[0] = m;
#pragma unroll
            for (int i = 0; i < 32; ++i) qv[i] = __expf(qv[i] - m);
            st[1] = dp_sum32(qv);
#pragma unroll
            for (int n = 0; n < NC; ++n) {
              float t[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) t[i] = __ldg(ct + DP_CT(2, n, i));
              st[2 + n] = dp_dot32(qv, t);
            }
            // the value slices are requested before the exchange
            float v2[NC][32];
#pragma unroll
            for (int n = 0; n < NC; ++n) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v2[n][i] = __ldg(ct + DP_CT(3, n, i));
            }
            // combine the per-slice softmax pieces: M = max m_s, s = sum s_s e^(m_s - M), w_n = sum dn_s e^(m_s - M) / s
            float wn[NC];
            {
              constexpr int K0 = (2 + NC) > 4 ? 4 : (2 + NC);
              float a[K0];
#pragma unroll
              for (int k = 0; k < K0; ++k) a[k] = st[k];
              ex.send<K0>(a);
              float M = ex.get(0, 0);
#pragma unroll
              for (int s = 1; s < DP_CL; ++s) M = fmaxf(M, ex.get(s, 0));
              float sc[DP_CL];
              float tot = 0.f;
#pragma unroll
              for (int s = 0; s < DP_CL; ++s) { sc[s] = __expf(ex.get(s, 0) - M); tot = fmaf(ex.get(s, 1), sc[s], tot); }
#pragma unroll
              for (int n = 0; n < NC && n < 2; ++n) {
                float t = 0.f;
#pragma unroll
                for (int s = 0; s < DP_CL; ++s) t = fmaf(ex.get(s, 2 + n), sc[s], t);
                wn[n] = t;
              }
              if constexpr (NC > 2) {
                float a2[NC - 2];
#pragma unroll
                for (int k = 0; k < NC - 2; ++k) a2[k] = st[4 + k];
                ex.send<NC - 2>(a2);
#pragma unroll
                for (int n = 2; n < NC; ++n) {
                  float t = 0.f;
#pragma unroll
                  for (int s = 0; s < DP_CL; ++s) t = fmaf(ex.get(s, n - 2), sc[s], t);
                  wn[n] = t;
                }
              }
              const float inv = 1.0f / tot;
#pragma unroll
              for (int n = 0; n < NC; ++n) wn[n] *= inv;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = 0.f;
#pragma unroll
            for (int n = 0; n < NC; ++n) {
#pragma unroll
              for (int i = 0; i < 32; ++i) y[i] = fmaf(wn[n], v2[n][i], y[i]);
            }
          }
          ex.layernorm(y, L.film_ca + toff + S, L.film_ca + toff + 256 + S);    // LN affine and FiLM folded (table)
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] = __fdividef(y[i], 1.0f + __expf(-y[i]));            // SiLU
          dp_publish<32>(xt, XC_HB0 + S, row, y);
          signal_x();
        }
        {   // x3 = x2 + out(h)
          float x[32], xr[32], b[32];
          dp_ldvec(L.bcaout + S, b);
          dp_ld32(tl + TM_XRES, xr);
          acc_wait();
          dp_ld32x2(tl, tl + 32, x);
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] += b[i] + xr[i];
          dp_st32(tl + TM_XRES, x);
          dp_publish<32>(xt, XC_X3 + S, row, x);
          signal_x();
        }
        // ---- FFN 256 -> 128 (GELU) -> 256 + FiLM (mdiff_transformer.py:241-254) ----
        {
          float f[32], b[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) b[i] = __ldg(L.bf1 + rank * DP_GS + i);
          acc_wait();
          dp_ld32(tl, f);                 // columns [0, 16): this CTA's hidden units, [16, 32): their hi.lo cross terms
          float g[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) g[i] = gelu_erf(f[i] + f[16 + i] + b[i]);
          dp_publish<16>(xt, XC_G + rank * DP_GS, row, g);
          signal_x();
        }
        {
          float y[32], b[32];
          dp_ldvec(L.bf2 + S, b);
          acc_wait();
          dp_ld32x2(tl, tl + 32, y);
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] += b[i];
          ex.layernorm(y, L.film_ff + toff + S, L.film_ff + toff + 256 + S);
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] = __fdividef(y[i], 1.0f + __expf(-y[i]));
          dp_publish<32>(xt, XC_HB1 + S, row, y);
          signal_x();
        }
        {   // block output = x3 + out(h)
          float x[32], xr[32], b[32];
          dp_ldvec(L.bfout + S, b);
          dp_ld32(tl + TM_XRES, xr);
          acc_wait();
          dp_ld32x2(tl, tl + 32, x);
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] += b[i] + xr[i];
          if (l < 4) {
            dp_st32(tl + TM_XRES, x);
            dp_publish<32>(xt, xout_col + S, row, x);
            signal_x();
          } else {
            // final LayerNorm (cross_attention.py:82), then CFG combine + DDIM update (mld.py:488-497)
            ex.layernorm(x, p.fng + S, p.fnb + S);
            if (p.mode == 1) {
              if (valid) {
                float4* dst = reinterpret_cast<float4*>(p.out + (size_t)grow * 256 + S);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
              }
            } else {
              if (p.cfg) {
                // rows r (uncond) and r + 64 (cond) hold the two guidance branches of one latent; both halves apply the
                // same update.  The statistics buffers are quiescent here: no peer can start the next exchange before
                // this CTA has published the next step's input.
                float* sw = stats;
#pragma unroll
                for (int i = 0; i < 32; ++i) sw[i * 128 + row] = x[i];
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const float gs = __ldg(p.gscale);
                const bool is_u = row < 64;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float o = sw[i * 128 + (row ^ 64)];
                  const float eu = is_u ? x[i] : o, ec = is_u ? o : x[i];
                  x[i] = __fadd_rn(eu, __fmul_rn(gs, __fsub_rn(ec, eu)));
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
              }
              const float c0 = __ldg(p.coef + step * 4), c1 = __ldg(p.coef + step * 4 + 1), c2 = __ldg(p.coef + step * 4 + 2),
                          c3 = __ldg(p.coef + step * 4 + 3);
              float lt[32];
              dp_ld32(tl + TM_LAT, lt);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x0 = __fdiv_rn(__fsub_rn(lt[i], __fmul_rn(c0, x[i])), c1);
                lt[i] = __fadd_rn(__fmul_rn(c2, x0), __fmul_rn(c3, x[i]));
              }
              if (step == p.n_steps - 1) {
                if (valid && (!p.cfg || row < 64)) {
                  float4* dst = reinterpret_cast<float4*>(p.out + (size_t)latrow * 256 + S);
#pragma unroll
                  for (int j = 0; j < 8; ++j) dst[j] = make_float4(lt[4 * j], lt[4 * j + 1], lt[4 * j + 2], lt[4 * j + 3]);
                }
              } else {
                dp_st32(tl + TM_LAT, lt);
                float pe[32];
                dp_ldvec(p.pe0 + S, pe);
#pragma unroll
                for (int i = 0; i < 32; ++i) lt[i] += pe[i];
                dp_st32(tl + TM_XRES, lt);
                dp_publish<32>(xt, XC_P0 + S, row, lt);
                signal_x();
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // no CTA leaves while a peer may still arrive on its barriers / write its statistics buffers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =====================================================================================================================
// Monolithic per-tile sampler ("tile" back-end): ONE CTA owns a tile of 128 denoiser rows for the whole run and computes
// every GEMM in full -- no cluster, no exchange.  Same algebra and tables as the cluster kernel above, plus:
//   * the self-attention logit of the latent token against itself, (W_q x + b_q).(W_k x + b_k), is evaluated as
//     x.(M x + m) + c with M = W_q^T W_k (fp64 product at create): the q and k rows never have to exist at the same time
//     (tensor memory holds ONE 256-column accumulator next to the fp32 residual stream);
//   * the A operand [128 x 256] (bf16 hi | lo, SWIZZLE_128B K-blocks) lives in shared memory and is rewritten in place by
//     the epilogue threads; only the weights stream (a tape of 16 KB tiles, 5-slot ring);
//   * the FFN 256 -> 1024 -> 256 runs in 8 hidden chunks of 128: relu(h) goes back to tensor memory as packed bf16 (hi, lo)
//     and is the A operand of the second GEMM straight from there (tcgen05.mma with A in TMEM); while it runs, the fp32
//     residual is parked in an L2-resident scratch row.
// SM time per run is ~1/8 of the cluster kernel's (4 CTAs instead of 32 for 512 rows) at about the same latency, which is
// what the batch pipeline wants next to the scene encoder; with >= 148 tiles it is the saturating configuration.
constexpr int DM_THREADS = 192;
constexpr int DM_NW = 6;                     // 96 KB in flight: the weight stream is what bounds the MMA phases (trace: 26 B/clk with 5 slots)
constexpr int DM_W_SLOT = 16384;             // one [128 n x 64 k] bf16 tile (hi or lo)
constexpr int DM_A_BYTES = 4 * 32768;        // 4 K-blocks x (hi 16 KB | lo 16 KB)
constexpr int DM_SMEM = DM_A_BYTES + DM_NW * DM_W_SLOT + 1024;    // the CFG pair exchange goes through the L2 scratch row
constexpr int DM_XR = 256;                   // tensor-memory columns: [0,256) accumulator, [256,512) fp32 residual stream;
constexpr int DM_ACC1 = 256, DM_HHI = 384, DM_HLO = 448;   // during the FFN: chunk accumulator | relu(h) hi | lo

struct DmUnit {       // one [128 rows x 128 outputs] GEMM unit; its weights are the next 2 nkb tiles of the tape
  int nkb, acc_col, a_base, ts, first, commit, accum, swap;
};
constexpr int DM_MAX_UNITS = 400;

struct DmParams {
  DpLayerP L[5];
  const float* mu[5];          // [257]: m = W_q^T b_k + W_k^T b_q, then c = b_q . b_k
  const float *skip_b[2], *fng, *fnb, *pe0;
  const uint8_t* tape;
  const DmUnit* units;
  int n_units;
  float *lat, *xs1;            // [tiles][256][128] fp32 scratch: latents, residual parked during the FFN
  uint8_t* skipimg;            // [tiles][2][4 K-blocks][32 KB]: outputs of blocks 0 / 1 as A-operand images
  const float* x_in;
  float* out;
  const float *coef, *gscale;
  int B, R, cfg, n_steps, mode;
  unsigned long long* trace;   // diagnostics (SEEME_DP_TRACE): role 1 = MMA issuer (4 go received, 5 group committed), 2 = epilogue row 64
  int trace_step;              // (6 accumulator ready, 10 go signalled), CTA 0 at trace_step
};

__device__ __forceinline__ void dm_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void dm_tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// A 256-float vector shared by all rows (bias, LayerNorm affine, FiLM, time-token table), distributed over the lanes of a
// warp: lane l holds elements [8 l, 8 l + 8).  Loaded with two 16-byte loads per lane BEFORE the accumulator wait (the L2
// round trip overlaps the GEMM) and read back with shuffles -- with 225 KB of shared memory the L1 is ~3 KB, so a per-chunk
// __ldg of such a vector is an L2 round trip on the epilogue's critical path (55 % of a step in the first trace).
struct RVec { float r[8]; };
__device__ __forceinline__ RVec rv_load(const float* __restrict__ p, int lane) {
  RVec v;
  const float4 a = __ldg(reinterpret_cast<const float4*>(p) + 2 * lane), b = __ldg(reinterpret_cast<const float4*>(p) + 2 * lane + 1);
  v.r[0] = a.x; v.r[1] = a.y; v.r[2] = a.z; v.r[3] = a.w; v.r[4] = b.x; v.r[5] = b.y; v.r[6] = b.z; v.r[7] = b.w;
  return v;
}
// element ch * 32 + i (i compile-time)
#define RV_GET(v, ch, i) __shfl_sync(0xffffffffu, (v).r[(i) & 7], (ch) * 4 + ((i) >> 3))
// 128-float vector: lane l holds [4 l, 4 l + 4); element c4 * 32 + i
struct RVec4 { float r[4]; };
__device__ __forceinline__ RVec4 rv4_load(const float* __restrict__ p, int lane) {
  RVec4 v;
  const float4 a = __ldg(reinterpret_cast<const float4*>(p) + lane);
  v.r[0] = a.x; v.r[1] = a.y; v.r[2] = a.z; v.r[3] = a.w;
  return v;
}
#define RV4_GET(v, c4, i) __shfl_sync(0xffffffffu, (v).r[(i) & 3], (c4) * 8 + ((i) >> 2))

template <int NC>
__global__ void __launch_bounds__(DM_THREADS, 1) den_mono_kernel(const __grid_constant__ DmParams p) {
  extern __shared__ __align__(1024) uint8_t dm_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dm_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_buf = smem;
  uint8_t* w_ring = smem + DM_A_BYTES;
  __shared__ __align__(8) uint64_t full_w[DM_NW], empty_w[DM_NW], acc_full, go, a_free, a_full2;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < DM_NW; ++i) { mbar_init(&full_w[i], 1); mbar_init(&empty_w[i], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&go, 4);
    mbar_init(&a_free, 1);
    mbar_init(&a_full2, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ---- producer: the weight tape, in order; the saved block output for the second half of a skip fusion ----
    if (lane == 0) {
      uint32_t iw = 0, nsw = 0;
      for (int step = 0; step < p.n_steps; ++step) {
        const uint8_t* src = p.tape;
        for (int u = 0; u < p.n_units; ++u) {
          const DmUnit un = p.units[u];
          if (un.swap) {
            dp_wait(&a_free, nsw & 1u, 21);
            ++nsw;
            mbar_arrive_expect_tx(&a_full2, DM_A_BYTES);
            const uint8_t* img = p.skipimg + ((size_t)(tile * 2 + (un.swap - 1)) * 4) * 32768;
            dp_fence_proxy_all();
            for (int kb = 0; kb < 4; ++kb) dp_bulk_load(a_buf + kb * 32768, img + (size_t)kb * 32768, 32768, &a_full2);
          }
          for (int t = 0; t < 2 * un.nkb; ++t) {
            const uint32_t sw = iw % DM_NW;
            dp_wait(&empty_w[sw], ((iw / DM_NW) & 1u) ^ 1u, 22);
            mbar_arrive_expect_tx(&full_w[sw], DM_W_SLOT);
            dp_bulk_load(w_ring + sw * DM_W_SLOT, src, DM_W_SLOT, &full_w[sw]);
            src += DM_W_SLOT;
            ++iw;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    uint32_t iw = 0, ngo = 0, nsw = 0;
    int tr_n = 0;
    const uint64_t da_base = umma_desc_k128(smem_u32(a_buf));
    const uint64_t dw_base = umma_desc_k128(smem_u32(w_ring));
    const uint32_t idesc = umma_idesc_bf16(128);
    for (int step = 0; step < p.n_steps; ++step) {
      for (int u = 0; u < p.n_units; ++u) {
        const DmUnit un = p.units[u];
        if (un.first) {
          dp_wait(&go, ngo & 1u, 23);
          ++ngo;
          tc_fence_after();
          if (lane == 0) DP_TR(1, 4);
        }
        if (un.swap) {
          if (umma_elect_one()) umma_commit(&a_free);      // the MMAs of the first K half have read the A operand
          __syncwarp();
          dp_wait(&a_full2, nsw & 1u, 24);
          ++nsw;
        }
        const uint32_t d_t = tmem_base + (uint32_t)un.acc_col;
        for (int kb = 0; kb < un.nkb; ++kb) {
          const uint32_t s0 = iw % DM_NW, s1 = (iw + 1) % DM_NW;
          dp_wait(&full_w[s0], (iw / DM_NW) & 1u, 25);
          dp_wait(&full_w[s1], ((iw + 1) / DM_NW) & 1u, 26);
          tc_fence_after();
          if (umma_elect_one()) {
            const uint64_t dwh0 = umma_desc_add(dw_base, s0 * (DM_W_SLOT >> 4));
            const uint64_t dwl0 = umma_desc_add(dw_base, s1 * (DM_W_SLOT >> 4));
            if (!un.ts) {
              const uint64_t da0 = umma_desc_add(da_base, (uint32_t)(un.a_base + kb) * (32768 >> 4));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_desc_add(da0, k * 2), dwh = umma_desc_add(dwh0, k * 2), dwl = umma_desc_add(dwl0, k * 2);
                umma_bf16(d_t, da, dwh, idesc, ((kb | k) != 0) || un.accum);
                umma_bf16(d_t, umma_desc_add(da, 16384 >> 4), dwh, idesc, 1);
                umma_bf16(d_t, da, dwl, idesc, 1);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t ta = tmem_base + (uint32_t)(un.a_base + kb * 32 + k * 8);
                const uint64_t dwh = umma_desc_add(dwh0, k * 2), dwl = umma_desc_add(dwl0, k * 2);
                dm_umma_ts(d_t, ta, dwh, idesc, ((kb | k) != 0) || un.accum);
                dm_umma_ts(d_t, ta + (DM_HLO - DM_HHI), dwh, idesc, 1);
                dm_umma_ts(d_t, ta, dwl, idesc, 1);
              }
            }
            umma_commit(&empty_w[s0]);
            umma_commit(&empty_w[s1]);
            if (un.commit && kb == un.nkb - 1) umma_commit(&acc_full);
          }
          __syncwarp();
          iw += 2;
        }
        if (un.commit && lane == 0) DP_TR(1, 5);
      }
    }
  } else {
    // ---- epilogue: thread = row, 256 columns in chunks of 32 ----
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int prow = tile * 128 + row;
    int latrow, grow;
    bool valid;
    if (p.cfg) {
      latrow = tile * 64 + (row & 63);
      valid = latrow < p.B;
      grow = (row >> 6) * p.B + latrow;
    } else {
      latrow = prow;
      grow = prow;
      valid = prow < p.R;
    }
    float* const lat_g = p.lat + (size_t)tile * 256 * 128 + row;
    float* const xs1_g = p.xs1 + (size_t)tile * 256 * 128 + row;
    uint8_t* const skip_g = p.skipimg + (size_t)tile * 2 * 4 * 32768;
    const size_t ct_off = (size_t)tile * 4 * NC * 256 * 128 + row;
    uint32_t nacc = 0;
    int tr_n = 0, step = 0;
    const bool tr_me = warp == 2 && lane == 0;
    auto acc_wait = [&]() {
      dp_wait(&acc_full, nacc & 1u, 27);
      ++nacc;
      tc_fence_after();
      if (tr_me) DP_TR(2, 6);
    };
    auto signal_go = [&]() {
      dp_fence_proxy_all();        // shared-memory A operand written with st.shared -> read by the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&go);
      if (tr_me) DP_TR(2, 10);
    };
    // LayerNorm over the 256 columns stored at TMEM columns [base, base + 256): returns (mean, rstd); biased variance, eps 1e-5
    auto ln_stats = [&](uint32_t base, float sum, float& mean, float& rstd) {
      mean = sum * (1.0f / 256.0f);
      float m2 = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < 8; ++ch) {
        float v[32];
        dp_ld32(tl + base + ch * 32, v);
        m2 += dp_sqdev32(v, mean);
      }
      rstd = rsqrtf(m2 * (1.0f / 256.0f) + 1e-5f);
    };

    // per-row cond-token table chunk: 32 columns of token n, table w (0 k, 1 ov, 2 softmax_n(key), 3 value)
    auto ct_load = [&](const float* __restrict__ ct, int w, int n, int ch, float (&v)[32]) {
      const float* src = ct + (size_t)DP_CT(w, n, 0) + (size_t)ch * 32 * 128;
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __ldg(src + i * 128);
    };
    {   // initial state: x = latents (or the given sample) + learned PE row 0
      const RVec pe = rv_load(p.pe0, lane);
#pragma unroll 1
      for (int ch = 0; ch < 8; ++ch) {
        float x[32];
        if (valid) {
          const float* src = p.x_in + (size_t)(p.mode == 0 ? latrow : grow) * 256 + ch * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(src) + j);
            x[4 * j] = t.x; x[4 * j + 1] = t.y; x[4 * j + 2] = t.z; x[4 * j + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) { lat_g[(size_t)(ch * 32 + i) * 128] = x[i]; x[i] += RV_GET(pe, ch, i); }
        dp_st32(tl + DM_XR + ch * 32, x);
        dm_write_a(a_buf, ch * 32, row, x);
      }
      signal_go();
    }

    for (step = 0; step < p.n_steps; ++step) {
      const size_t toff = (size_t)step * 512;
#pragma unroll 1
      for (int l = 0; l < 5; ++l) {
        const DpLayerP& L = p.L[l];
        const float* __restrict__ ct = L.ctab + ct_off;
        if (l >= 3) {      // x = Linear(cat[x, skip])
          const RVec b = rv_load(p.skip_b[l - 3], lane);
          acc_wait();
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float x[32];
            dp_ld32(tl + ch * 32, x);
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] += RV_GET(b, ch, i);
            dp_st32(tl + DM_XR + ch * 32, x);
            dm_write_a(a_buf, ch * 32, row, x);
          }
          signal_go();
        }
        // ---- self-attention of the latent token over {x, cond tokens, time token} ----
        float pr[NC + 2];
        {   // u = M x: logit against itself = x . (u + m) + c
          const RVec m = rv_load(p.mu[l], lane);
          const float c0 = __ldg(p.mu[l] + 256);
          acc_wait();
          float d = 0.f;
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float u[32], xr[32];
            dp_ld32(tl + ch * 32, u);
            dp_ld32(tl + DM_XR + ch * 32, xr);
#pragma unroll
            for (int i = 0; i < 32; ++i) u[i] += RV_GET(m, ch, i);
            d += dp_dot32(xr, u);
          }
          pr[0] = d + c0;
          signal_go();
        }
        {   // q: logits against the cond tokens and the time token
          const RVec bq = rv_load(L.bqkv, lane), ktk = rv_load(L.kt + toff, lane);
          float kc[NC][32];
#pragma unroll
          for (int n = 0; n < NC; ++n) ct_load(ct, 0, n, 0, kc[n]);
          acc_wait();
          float dc[NC], dt = 0.f;
#pragma unroll
          for (int n = 0; n < NC; ++n) dc[n] = 0.f;
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float qv[32];
            dp_ld32(tl + ch * 32, qv);
            {
              float kt32[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) { qv[i] += RV_GET(bq, ch, i); kt32[i] = RV_GET(ktk, ch, i); }
              dt += dp_dot32(qv, kt32);
            }
#pragma unroll
            for (int n = 0; n < NC; ++n) {
              dc[n] += dp_dot32(qv, kc[n]);
              if (ch < 7) ct_load(ct, 0, n, ch + 1, kc[n]);      // the next chunk's table values travel during this chunk's tail
            }
          }
          signal_go();
#pragma unroll
          for (int n = 0; n < NC; ++n) pr[1 + n] = dc[n];
          pr[NC + 1] = dt;
          float m = pr[0];
#pragma unroll
          for (int j = 1; j < NC + 2; ++j) m = fmaxf(m, pr[j]);
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < NC + 2; ++j) { pr[j] = __expf(pr[j] - m); sum += pr[j]; }
          const float inv = 1.0f / sum;
#pragma unroll
          for (int j = 0; j < NC + 2; ++j) pr[j] *= inv;
        }
        {   // ov: t0 = x + b + sum_j p_j ov_j; x1 = norm1(t0)
          const RVec bs = rv_load(L.bqkv + 512, lane), ovt = rv_load(L.kt + toff + 256, lane);
          const RVec g1 = rv_load(L.n1g, lane), b1n = rv_load(L.n1b, lane);
          float oc[NC][32];
#pragma unroll
          for (int n = 0; n < NC; ++n) ct_load(ct, 1, n, 0, oc[n]);
          acc_wait();
          float sum = 0.f;
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float t0[32];
            dp_ld32(tl + ch * 32, t0);
#pragma unroll
            for (int i = 0; i < 32; ++i) t0[i] = RV_GET(bs, ch, i) + pr[0] * t0[i] + pr[NC + 1] * RV_GET(ovt, ch, i);
#pragma unroll
            for (int n = 0; n < NC; ++n) {
#pragma unroll
              for (int i = 0; i < 32; ++i) t0[i] = fmaf(pr[1 + n], oc[n][i], t0[i]);
              if (ch < 7) ct_load(ct, 1, n, ch + 1, oc[n]);
            }
            {
              float xr[32];
              dp_ld32(tl + DM_XR + ch * 32, xr);
#pragma unroll
              for (int i = 0; i < 32; ++i) t0[i] += xr[i];
            }
            sum += dp_sum32(t0);
            dp_st32(tl + ch * 32, t0);
          }
          float mean, rstd;
          ln_stats(0, sum, mean, rstd);
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float t0[32];
            dp_ld32(tl + ch * 32, t0);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              t0[i] = (t0[i] - mean) * rstd * RV_GET(g1, ch, i) + RV_GET(b1n, ch, i);
              xs1_g[(size_t)(ch * 32 + i) * 128] = t0[i];       // the fp32 residual waits in L2 while the FFN uses its columns
            }
            dm_write_a(a_buf, ch * 32, row, t0);
          }
          signal_go();
        }
        // ---- FFN 256 -> 1024 (ReLU) -> 256 in 8 hidden chunks of 128; relu(h) returns to tensor memory as bf16 (hi, lo) ----
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          const RVec4 b = rv4_load(L.b1 + j * 128, lane);
          acc_wait();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            float f[32];
            dp_ld32(tl + DM_ACC1 + c4 * 32, f);
            uint32_t hb[16], lb[16];
#pragma unroll
            for (int i = 0; i < 16; ++i)
              dp_split2(fmaxf(f[2 * i] + RV4_GET(b, c4, 2 * i), 0.f), fmaxf(f[2 * i + 1] + RV4_GET(b, c4, 2 * i + 1), 0.f), hb[i], lb[i]);
            dm_tmem_st16(tl + DM_HHI + c4 * 16, hb);
            dm_tmem_st16(tl + DM_HLO + c4 * 16, lb);
          }
          tmem_st_wait_dp();
          signal_go();
        }
        {   // x2 = norm2(x1 + ffn); then the input norm of the cross-attention
          const RVec b2 = rv_load(L.b2, lane), g2 = rv_load(L.n2g, lane), b2n = rv_load(L.n2b, lane);
          const RVec gc = rv_load(L.cng, lane), bc = rv_load(L.cnb, lane);
          float xs[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) xs[i] = xs1_g[(size_t)i * 128];
          acc_wait();
          float sum = 0.f;
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float t1[32];
            dp_ld32(tl + ch * 32, t1);
#pragma unroll
            for (int i = 0; i < 32; ++i) t1[i] += RV_GET(b2, ch, i) + xs[i];
            sum += dp_sum32(t1);
            if (ch < 7) {
#pragma unroll
              for (int i = 0; i < 32; ++i) xs[i] = xs1_g[(size_t)((ch + 1) * 32 + i) * 128];
            }
            dp_st32(tl + ch * 32, t1);
          }
          float mean, rstd;
          ln_stats(0, sum, mean, rstd);
          float sum2 = 0.f;
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float t1[32];
            dp_ld32(tl + ch * 32, t1);
#pragma unroll
            for (int i = 0; i < 32; ++i) t1[i] = (t1[i] - mean) * rstd * RV_GET(g2, ch, i) + RV_GET(b2n, ch, i);
            sum2 += dp_sum32(t1);
            dp_st32(tl + DM_XR + ch * 32, t1);      // the FFN is done with these columns: the residual stream is back
          }
          ln_stats(DM_XR, sum2, mean, rstd);
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float t1[32];
            dp_ld32(tl + DM_XR + ch * 32, t1);
#pragma unroll
            for (int i = 0; i < 32; ++i) t1[i] = (t1[i] - mean) * rstd * RV_GET(gc, ch, i) + RV_GET(bc, ch, i);
            dm_write_a(a_buf, ch * 32, row, t1);
          }
          signal_go();
        }
        {   // linear cross-attention to the cond tokens + FiLM
          const RVec bq2 = rv_load(L.bcaq, lane), fg = rv_load(L.film_ca + toff, lane), fb = rv_load(L.film_ca + toff + 256, lane);
          float tb[NC][32];
#pragma unroll
          for (int n = 0; n < NC; ++n) ct_load(ct, 2, n, 0, tb[n]);
          acc_wait();
          float m = -INFINITY;
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float qv[32];
            dp_ld32(tl + ch * 32, qv);
#pragma unroll
            for (int i = 0; i < 32; ++i) qv[i] += RV_GET(bq2, ch, i);
            m = fmaxf(m, dp_max32(qv));
            dp_st32(tl + ch * 32, qv);
          }
          float ssum = 0.f, dn[NC];
#pragma unroll
          for (int n = 0; n < NC; ++n) dn[n] = 0.f;
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float qv[32];
            dp_ld32(tl + ch * 32, qv);
#pragma unroll
            for (int i = 0; i < 32; ++i) qv[i] = __expf(qv[i] - m);
            ssum += dp_sum32(qv);
#pragma unroll
            for (int n = 0; n < NC; ++n) {
              dn[n] += dp_dot32(qv, tb[n]);
              if (ch < 7) ct_load(ct, 2, n, ch + 1, tb[n]);
              else ct_load(ct, 3, n, 0, tb[n]);                  // then the value table, chunk 0
            }
          }
          const float inv = 1.0f / ssum;
#pragma unroll
          for (int n = 0; n < NC; ++n) dn[n] *= inv;
          float sum = 0.f;
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float y[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = 0.f;
#pragma unroll
            for (int n = 0; n < NC; ++n) {
#pragma unroll
              for (int i = 0; i < 32; ++i) y[i] = fmaf(dn[n], tb[n][i], y[i]);
              if (ch < 7) ct_load(ct, 3, n, ch + 1, tb[n]);
            }
            sum += dp_sum32(y);
            dp_st32(tl + ch * 32, y);
          }
          float mean, rstd;
          ln_stats(0, sum, mean, rstd);
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float y[32];
            dp_ld32(tl + ch * 32, y);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float t = (y[i] - mean) * rstd * RV_GET(fg, ch, i) + RV_GET(fb, ch, i);
              y[i] = __fdividef(t, 1.0f + __expf(-t));
            }
            dm_write_a(a_buf, ch * 32, row, y);
          }
          signal_go();
        }
        {   // x3 = x2 + out(h)
          const RVec b = rv_load(L.bcaout, lane);
          acc_wait();
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float x[32], xr[32];
            dp_ld32(tl + ch * 32, x);
            dp_ld32(tl + DM_XR + ch * 32, xr);
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] += RV_GET(b, ch, i) + xr[i];
            dp_st32(tl + DM_XR + ch * 32, x);
            dm_write_a(a_buf, ch * 32, row, x);
          }
          signal_go();
        }
        {   // FFN 256 -> 128 (GELU)
          const RVec4 b = rv4_load(L.bf1, lane);
          acc_wait();
#pragma unroll 1
          for (int c4 = 0; c4 < 4; ++c4) {
            float f[32];
            dp_ld32(tl + c4 * 32, f);
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i] + RV4_GET(b, c4, i));
            dm_write_a(a_buf, c4 * 32, row, f);
          }
          signal_go();
        }
        {   // -> 256, FiLM
          const RVec b = rv_load(L.bf2, lane), fg = rv_load(L.film_ff + toff, lane), fb = rv_load(L.film_ff + toff + 256, lane);
          acc_wait();
          float sum = 0.f;
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float y[32];
            dp_ld32(tl + ch * 32, y);
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] += RV_GET(b, ch, i);
            sum += dp_sum32(y);
            dp_st32(tl + ch * 32, y);
          }
          float mean, rstd;
          ln_stats(0, sum, mean, rstd);
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float y[32];
            dp_ld32(tl + ch * 32, y);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float t = (y[i] - mean) * rstd * RV_GET(fg, ch, i) + RV_GET(fb, ch, i);
              y[i] = __fdividef(t, 1.0f + __expf(-t));
            }
            dm_write_a(a_buf, ch * 32, row, y);
          }
          signal_go();
        }
        {   // block output = x3 + out(h)
          const RVec b = rv_load(L.bfout, lane);
          acc_wait();
          float sum = 0.f;
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            float x[32], xr[32];
            dp_ld32(tl + ch * 32, x);
            dp_ld32(tl + DM_XR + ch * 32, xr);
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] += RV_GET(b, ch, i) + xr[i];
            sum += dp_sum32(x);
            if (l < 4) {
              dp_st32(tl + DM_XR + ch * 32, x);
              dm_write_a(a_buf, ch * 32, row, x);
              if (l < 2) dp_publish<32>(skip_g + (size_t)l * 4 * 32768, ch * 32, row, x);    // saved for the skip fusion of block 4 - l
            } else {
              dp_st32(tl + ch * 32, x);
            }
          }
          if (l < 4) {
            if (l < 2) __threadfence();
            signal_go();
          } else {
            // final LayerNorm, then CFG combine + DDIM update
            const RVec fg = rv_load(p.fng, lane), fb = rv_load(p.fnb, lane), pe = rv_load(p.pe0, lane);
            float mean, rstd;
            ln_stats(0, sum, mean, rstd);
            const float c0 = __ldg(p.coef + step * 4), c1 = __ldg(p.coef + step * 4 + 1), c2 = __ldg(p.coef + step * 4 + 2),
                        c3 = __ldg(p.coef + step * 4 + 3);
            const float gs = __ldg(p.gscale);
            const bool last = step == p.n_steps - 1;
            if (p.mode == 0 && p.cfg) {
              // rows r (uncond) and r + 64 (cond) are the two guidance branches of one latent: every row publishes its eps in the
              // L2 scratch row (free between the FFN blocks), one CTA-wide epilogue barrier, then each row reads its partner's
#pragma unroll 1
              for (int ch = 0; ch < 8; ++ch) {
                float e[32];
                dp_ld32(tl + ch * 32, e);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  e[i] = (e[i] - mean) * rstd * RV_GET(fg, ch, i) + RV_GET(fb, ch, i);
                  __stcg(xs1_g + (size_t)(ch * 32 + i) * 128, e[i]);
                }
                dp_st32(tl + ch * 32, e);
              }
              __threadfence_block();
              asm volatile("bar.sync 1, 128;" ::: "memory");
            }
#pragma unroll 1
            for (int ch = 0; ch < 8; ++ch) {
              float e[32];
              dp_ld32(tl + ch * 32, e);
              if (p.mode == 1 || !p.cfg) {
#pragma unroll
                for (int i = 0; i < 32; ++i) e[i] = (e[i] - mean) * rstd * RV_GET(fg, ch, i) + RV_GET(fb, ch, i);
              }
              if (p.mode == 1) {
                if (valid) {
                  float4* dst = reinterpret_cast<float4*>(p.out + (size_t)grow * 256 + ch * 32);
#pragma unroll
                  for (int j = 0; j < 8; ++j) dst[j] = make_float4(e[4 * j], e[4 * j + 1], e[4 * j + 2], e[4 * j + 3]);
                }
                continue;
              }
              if (p.cfg) {
                const float* partner = p.xs1 + (size_t)tile * 256 * 128 + (row ^ 64);
                const bool is_u = row < 64;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float o = __ldcg(partner + (size_t)(ch * 32 + i) * 128);
                  const float eu = is_u ? e[i] : o, ec = is_u ? o : e[i];
                  e[i] = __fadd_rn(eu, __fmul_rn(gs, __fsub_rn(ec, eu)));
                }
              }
              float lt[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x0 = __fdiv_rn(__fsub_rn(lat_g[(size_t)(ch * 32 + i) * 128], __fmul_rn(c0, e[i])), c1);
                lt[i] = __fadd_rn(__fmul_rn(c2, x0), __fmul_rn(c3, e[i]));
              }
              if (last) {
                if (valid && (!p.cfg || row < 64)) {
                  float4* dst = reinterpret_cast<float4*>(p.out + (size_t)latrow * 256 + ch * 32);
#pragma unroll
                  for (int j = 0; j < 8; ++j) dst[j] = make_float4(lt[4 * j], lt[4 * j + 1], lt[4 * j + 2], lt[4 * j + 3]);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) { lat_g[(size_t)(ch * 32 + i) * 128] = lt[i]; lt[i] += RV_GET(pe, ch, i); }
                dp_st32(tl + DM_XR + ch * 32, lt);
                dm_write_a(a_buf, ch * 32, row, lt);
              }
            }
            if (p.mode == 0 && !last) {
              if (p.cfg) asm volatile("bar.sync 1, 128;" ::: "memory");   // every partner read is done before the scratch row is reused
              signal_go();
            }
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

// transposes the cond-token projections into the per-row tables the epilogue reads column by column:
//   tab[tile][which][n][col][row of the tile], which = 0: k, 1: ov (self-attention; kov rows n R + grow: k | ov), 2: softmax over the
//   tokens of the cross-attention key, 3: its value (kv2: key | value).  Rows beyond the batch are zero.
__global__ void dp_ctab_kernel(const float* __restrict__ kov, const float* __restrict__ kv2, float* __restrict__ tab, int Nc, int B,
                               int R, int cfg, int rows_pad) {
  const int prow = blockIdx.x * 32 + threadIdx.x;
  const int col = blockIdx.y * 8 + threadIdx.y;
  if (prow >= rows_pad) return;
  const int tile = prow >> 7, r = prow & 127;
  int grow;
  bool valid;
  if (cfg) {
    const int latrow = tile * 64 + (r & 63);
    valid = latrow < B;
    grow = (r >> 6) * B + latrow;
  } else {
    grow = prow;
    valid = prow < R;
  }
  float k2[SEEME_MAX_COND_TOKENS];
  float m = -INFINITY;
  for (int n = 0; n < Nc; ++n) {
    k2[n] = valid ? kv2[((size_t)n * R + grow) * 512 + col] : 0.f;
    m = fmaxf(m, k2[n]);
  }
  float s = 0.f;
  for (int n = 0; n < Nc; ++n) { k2[n] = expf(k2[n] - m); s += k2[n]; }
  const float inv = 1.0f / s;
  float* t = tab + (size_t)tile * 4 * Nc * 256 * 128 + r;
  for (int n = 0; n < Nc; ++n) {
    const size_t src = ((size_t)n * R + grow) * 512 + col;
    t[((size_t)(0 * Nc + n) * 256 + col) * 128] = valid ? kov[src] : 0.f;
    t[((size_t)(1 * Nc + n) * 256 + col) * 128] = valid ? kov[src + 256] : 0.f;
    t[((size_t)(2 * Nc + n) * 256 + col) * 128] = valid ? k2[n] * inv : 0.f;
    t[((size_t)(3 * Nc + n) * 256 + col) * 128] = valid ? kv2[src + 256] : 0.f;
  }
}

// ---- host side ------------------------------------------------------------------------------------
struct DenPersist {
  Arena arena;
  int max_tiles = 0, rows_pad_max = 0;
  uint8_t* tape = nullptr;
  size_t tape_cta_bytes = 0;
  DpUnit* d_units = nullptr;
  int n_units = 0;
  uint8_t* ximg = nullptr;
  float* part = nullptr;
  float* ctab[5] = {};
  float* bqkv[5] = {};           // [768]: bq / 16 | bk | W_o b_v
  float* wkov_f32[5] = {};       // [512,256]: W_k | W_o W_v  (time-token table GEMM)
  float* bkov[5] = {};           // [512]
  float* bk2v2[5] = {};          // [512]
  float* tkov[5] = {};           // [MAX_STEPS,512]
  float* film2_ca[5] = {};       // [MAX_STEPS,512]: g (1 + scale) | b (1 + scale) + shift
  float* film2_ff[5] = {};
  PackedLinear Wkov[5], Wk2v2[5];
  ActBuf condb, tnb;             // bf16 (hi, lo) of the cond tokens / of their text_norm
  unsigned long long* trace = nullptr;
  // monolithic per-tile kernel
  uint8_t* tape_m = nullptr;
  DmUnit* d_units_m = nullptr;
  int n_units_m = 0;
  float* mu[5] = {};
  float *lat_m = nullptr, *xs1_m = nullptr;
  uint8_t* skipimg = nullptr;
};

static inline uint16_t f2bf(float x) {    // round to nearest even, like __float2bfloat16_rn (finite inputs)
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf2f(uint16_t h) {
  const uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

namespace {
struct HostW {      // host copy of one fp32 matrix [rows, ld]
  std::vector<float> v;
  int ld = 0;
};
// appends the blobs of one unit: rows `rowsel` (n of them) of W, K-blocks [k0, k0 + 64 nkb)
void append_unit(std::vector<uint8_t>& tape, const HostW& W, const std::vector<int>& rowsel, int k0, int nkb) {
  const int n = (int)rowsel.size();
  for (int kb = 0; kb < nkb; ++kb) {
    const size_t base = tape.size();
    tape.resize(base + (size_t)n * 256);
    uint8_t* hi = tape.data() + base;
    uint8_t* lo = hi + (size_t)n * 128;
    for (int r = 0; r < n; ++r)
      for (int k = 0; k < 64; ++k) {
        const float x = W.v[(size_t)rowsel[r] * W.ld + k0 + kb * 64 + k];
        const uint16_t h = f2bf(x);
        const uint16_t l = f2bf(x - bf2f(h));
        const size_t off = (size_t)r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1));   // SWIZZLE_128B image
        memcpy(hi + off, &h, 2);
        memcpy(lo + off, &l, 2);
      }
  }
}
std::vector<int> iota_rows(int first, int n) {
  std::vector<int> r(n);
  for (int i = 0; i < n; ++i) r[i] = first + i;
  return r;
}
}  // namespace

int den_persist_create(seeme_denoiser* h) {
  DenPersist* P = new DenPersist();
  h->persist = P;
  P->max_tiles = (h->max_rows + 127) / 128;
  P->rows_pad_max = P->max_tiles * 128;
  // host copies of the GEMM weights (the q rows already carry the 1/16 attention scale)
  auto fetch = [&](const float* d, size_t n, std::vector<float>& out) -> int {
    out.resize(n);
    SEEME_CUDA(cudaMemcpy(out.data(), d, n * 4, cudaMemcpyDeviceToHost));
    return SEEME_OK;
  };
  std::vector<HostW> Wqkv(5), Wl1(5), Wl2(5), Wcaq(5), Wcaout(5), Wf1(5), Wf2(5), Wfout(5);
  HostW Wskip[2];
  std::vector<std::vector<float>> bqkv(5), bkov(5);
  for (int l = 0; l < 5; ++l) {
    std::vector<float> in_w, in_b, out_w;
    SEEME_TRY(fetch(blkw(h, l, SA_IN_W), 768 * 256, in_w));
    SEEME_TRY(fetch(blkw(h, l, SA_IN_B), 768, in_b));
    SEEME_TRY(fetch(blkw(h, l, SA_OUT_W), 256 * 256, out_w));
    Wqkv[l].v.assign(768 * 256, 0.f);
    Wqkv[l].ld = 256;
    memcpy(Wqkv[l].v.data(), in_w.data(), 512 * 256 * 4);          // q, k rows
    // W_ov = W_o W_v and b_ov = W_o b_v in double precision
    std::vector<double> acc(256);
    std::vector<float> bov(256);
    for (int n = 0; n < 256; ++n) {
      for (int k = 0; k < 256; ++k) acc[k] = 0.0;
      double bacc = 0.0;
      for (int j = 0; j < 256; ++j) {
        const double wo = out_w[(size_t)n * 256 + j];
        const float* wv = &in_w[(size_t)(512 + j) * 256];
        for (int k = 0; k < 256; ++k) acc[k] += wo * (double)wv[k];
        bacc += wo * (double)in_b[512 + j];
      }
      for (int k = 0; k < 256; ++k) Wqkv[l].v[(size_t)(512 + n) * 256 + k] = (float)acc[k];
      bov[n] = (float)bacc;
    }
    {
      std::vector<float> bo;
      SEEME_TRY(fetch(blkw(h, l, SA_OUT_B), 256, bo));
      // the softmax weights sum to 1: b_o + W_o b_v is added once; the (k | ov) tables of the cond / time tokens carry b_k only
      for (int n = 0; n < 256; ++n) in_b[512 + n] = bo[n] + bov[n];
    }
    bqkv[l] = in_b;
    bkov[l].assign(in_b.begin() + 256, in_b.end());
    for (int n = 0; n < 256; ++n) bkov[l][256 + n] = 0.f;
    SEEME_TRY(fetch(blkw(h, l, SA_L1_W), 1024 * 256, Wl1[l].v)); Wl1[l].ld = 256;
    SEEME_TRY(fetch(blkw(h, l, SA_L2_W), 256 * 1024, Wl2[l].v)); Wl2[l].ld = 1024;
    SEEME_TRY(fetch(blkw(h, l, CA_Q_W), 256 * 256, Wcaq[l].v)); Wcaq[l].ld = 256;
    SEEME_TRY(fetch(blkw(h, l, CA_OUT_W), 256 * 256, Wcaout[l].v)); Wcaout[l].ld = 256;
    SEEME_TRY(fetch(blkw(h, l, FF_L1_W), 128 * 256, Wf1[l].v)); Wf1[l].ld = 256;
    SEEME_TRY(fetch(blkw(h, l, FF_L2_W), 256 * 128, Wf2[l].v)); Wf2[l].ld = 128;
    SEEME_TRY(fetch(blkw(h, l, FF_OUT_W), 256 * 256, Wfout[l].v)); Wfout[l].ld = 256;
  }
  for (int i = 0; i < 2; ++i) { SEEME_TRY(fetch(h->w[DN_LB0_W + 2 * i], 256 * 512, Wskip[i].v)); Wskip[i].ld = 512; }

  // unit list of one step and the tapes of the 8 CTAs (identical unit order, different rows)
  std::vector<DpUnit> units;
  std::vector<std::vector<uint8_t>> tapes(DP_CL);
  for (int c = 0; c < DP_CL; ++c) {
    std::vector<uint8_t>& T = tapes[c];
    T.reserve(2700000);
    const int S = c * DP_NS;
    auto add = [&](const HostW& W, const std::vector<int>& rows, int k0, int nkb, DpUnit u) {
      u.n = (int)rows.size();
      u.nkb = nkb;
      u.w_off = (unsigned)T.size();
      if (c == 0) units.push_back(u);
      append_unit(T, W, rows, k0, nkb);
    };
    auto U = [&](int a_col, int wait) {
      DpUnit u;
      memset(&u, 0, sizeof(u));
      u.a_col = a_col; u.wait = wait; u.nkb1 = 1 << 20; u.release_a = 1; u.last = 1; u.cat = 1;
      return u;
    };
    for (int l = 0; l < 5; ++l) {
      const int xin = l == 0 ? XC_P0 : l == 1 ? XC_L0 : l == 2 ? XC_L1 : XC_P0;
      if (l >= 3) {     // skip fusion: K = 512 = [current x (P1) | saved block output]
        DpUnit u = U(XC_P1, 1);
        u.nkb1 = 4;
        u.a_col2 = l == 3 ? XC_L1 : XC_L0;
        add(Wskip[l - 3], iota_rows(S, DP_NS), 0, 8, u);
      }
      {   // q | k slices (N = 64), then ov slice (N = 32) on the same A
        std::vector<int> r = iota_rows(S, DP_NS), r2 = iota_rows(256 + S, DP_NS);
        r.insert(r.end(), r2.begin(), r2.end());
        DpUnit u = U(xin, 1);
        u.release_a = 0; u.last = 0;
        add(Wqkv[l], r, 0, 4, u);
        DpUnit v = U(xin, 0);
        v.reuse_a = 1; v.acc_col = 128;
        add(Wqkv[l], iota_rows(512 + S, DP_NS), 0, 4, v);
      }
      for (int j = 0; j < DP_HS / 64; ++j) {   // hidden units [c HS + 64 j, +64)
        DpUnit u = U(XC_XA, j == 0 ? 1 : 0);
        u.reuse_a = j > 0; u.acc_col = 128 * j;
        u.release_a = j == DP_HS / 64 - 1; u.last = u.release_a;
        add(Wl1[l], iota_rows(c * DP_HS + 64 * j, 64), 0, 4, u);
      }
      for (int j = 0; j < 4; ++j) {            // partial result columns [64 j, +64) over this CTA's K slice
        DpUnit u = U(XC_FF, j == 0 ? 2 : 0);
        u.a_rstride = DP_HS;
        u.cat = 0;                               // 256 result columns: no room for the cross-term columns
        u.reuse_a = j > 0; u.acc_col = 64 * j;
        u.release_a = j == 3; u.last = j == 3;
        add(Wl2[l], iota_rows(64 * j, 64), c * DP_HS, DP_HS / 64, u);
      }
      add(Wcaq[l], iota_rows(S, DP_NS), 0, 4, U(XC_LN, 1));
      add(Wcaout[l], iota_rows(S, DP_NS), 0, 4, U(XC_HB0, 1));
      add(Wf1[l], iota_rows(c * DP_GS, DP_GS), 0, 4, U(XC_X3, 1));
      add(Wf2[l], iota_rows(S, DP_NS), 0, 2, U(XC_G, 1));
      add(Wfout[l], iota_rows(S, DP_NS), 0, 4, U(XC_HB1, 1));
    }
  }
  P->n_units = (int)units.size();
  SEEME_REQUIRE(P->n_units <= DP_MAX_UNITS, SEEME_EINVAL, "den_persist: %d units", P->n_units);
  P->tape_cta_bytes = tapes[0].size();
  for (int c = 1; c < DP_CL; ++c)
    SEEME_REQUIRE(tapes[c].size() == P->tape_cta_bytes, SEEME_EINVAL, "den_persist: tape size mismatch");

  // ---- monolithic kernel: M = W_q^T W_k, m = W_q^T b_k + W_k^T b_q, c = b_q . b_k (fp64), its unit list and tape ----
  std::vector<HostW> Wu(5);
  std::vector<std::vector<float>> mu(5);
  for (int l = 0; l < 5; ++l) {
    const std::vector<float>& W = Wqkv[l].v;          // rows: q (scaled) | k | ov
    Wu[l].v.assign(256 * 256, 0.f);
    Wu[l].ld = 256;
    std::vector<double> acc(256 * 256, 0.0);
    for (int n = 0; n < 256; ++n) {
      const float* wq = &W[(size_t)n * 256];
      const float* wk = &W[(size_t)(256 + n) * 256];
      for (int i = 0; i < 256; ++i) {
        const double a = wq[i];
        double* row = &acc[(size_t)i * 256];
        for (int j = 0; j < 256; ++j) row[j] += a * (double)wk[j];
      }
    }
    for (size_t i = 0; i < acc.size(); ++i) Wu[l].v[i] = (float)acc[i];
    mu[l].assign(257, 0.f);
    double c = 0.0;
    for (int i = 0; i < 256; ++i) {
      double m = 0.0;
      for (int n = 0; n < 256; ++n) m += (double)W[(size_t)n * 256 + i] * (double)bqkv[l][256 + n] + (double)W[(size_t)(256 + n) * 256 + i] * (double)bqkv[l][n];
      mu[l][i] = (float)m;
    }
    for (int n = 0; n < 256; ++n) c += (double)bqkv[l][n] * (double)bqkv[l][256 + n];
    mu[l][256] = (float)c;
  }
  std::vector<DmUnit> munits;
  std::vector<uint8_t> mtape;
  mtape.reserve(21u << 20);
  {
    auto MU = [&](int nkb, int acc_col, int a_base, int ts, int first, int commit, int accum, int swap) {
      DmUnit u;
      u.nkb = nkb; u.acc_col = acc_col; u.a_base = a_base; u.ts = ts; u.first = first; u.commit = commit; u.accum = accum; u.swap = swap;
      munits.push_back(u);
    };
    // a [256 outputs x K] linear on the shared-memory A operand: two 128-output units
    auto linear = [&](const HostW& W, int row0, int k0, int nkb, int accum, int swap_first, bool first, bool commit) {
      for (int half = 0; half < 2; ++half) {
        MU(nkb, 128 * half, 0, 0, first && half == 0, commit && half == 1, accum, half == 0 ? swap_first : 0);
        append_unit(mtape, W, iota_rows(row0 + 128 * half, 128), k0, nkb);
      }
    };
    HostW Wq[5], Wov[5];
    for (int l = 0; l < 5; ++l) {
      Wq[l].ld = Wov[l].ld = 256;
      Wq[l].v.assign(Wqkv[l].v.begin(), Wqkv[l].v.begin() + 256 * 256);
      Wov[l].v.assign(Wqkv[l].v.begin() + 512 * 256, Wqkv[l].v.end());
    }
    for (int l = 0; l < 5; ++l) {
      if (l >= 3) {     // K = 512: current x, then (after the A operand has been refilled) the saved output of block 4 - l
        linear(Wskip[l - 3], 0, 0, 4, 0, 0, true, false);
        linear(Wskip[l - 3], 0, 256, 4, 1, l == 3 ? 2 : 1, false, true);      // swap: 1 = block 0's output, 2 = block 1's
      }
      linear(Wu[l], 0, 0, 4, 0, 0, true, true);
      linear(Wq[l], 0, 0, 4, 0, 0, true, true);
      linear(Wov[l], 0, 0, 4, 0, 0, true, true);
      // FFN: [G1_0] [G2_0 G1_1] ... [G2_6 G1_7] [G2_7]
      auto g1 = [&](int j, bool first, bool commit) {
        MU(4, DM_ACC1, 0, 0, first, commit, 0, 0);
        append_unit(mtape, Wl1[l], iota_rows(128 * j, 128), 0, 4);
      };
      g1(0, true, true);
      for (int j = 0; j < 8; ++j) {
        for (int half = 0; half < 2; ++half) {
          MU(2, 128 * half, DM_HHI, 1, half == 0, j == 7 && half == 1, j > 0, 0);
          append_unit(mtape, Wl2[l], iota_rows(128 * half, 128), 128 * j, 2);
        }
        if (j < 7) g1(j + 1, false, true);
      }
      linear(Wcaq[l], 0, 0, 4, 0, 0, true, true);
      linear(Wcaout[l], 0, 0, 4, 0, 0, true, true);
      MU(4, 0, 0, 0, 1, 1, 0, 0);
      append_unit(mtape, Wf1[l], iota_rows(0, 128), 0, 4);
      linear(Wf2[l], 0, 0, 2, 0, 0, true, true);
      linear(Wfout[l], 0, 0, 4, 0, 0, true, true);
    }
  }
  P->n_units_m = (int)munits.size();
  SEEME_REQUIRE(P->n_units_m <= DM_MAX_UNITS, SEEME_EINVAL, "den_persist: %d monolithic units", P->n_units_m);

  const size_t Rp = (size_t)P->rows_pad_max, NCM = SEEME_MAX_COND_TOKENS;
  size_t bytes = pad256(P->tape_cta_bytes * DP_CL) + pad256(sizeof(DpUnit) * DP_MAX_UNITS) + pad256((size_t)P->max_tiles * XC_KB * 32768) +
                 pad256((size_t)P->max_tiles * DP_CL * 256 * 128 * 4) + 5 * pad256(4 * NCM * 256 * Rp * 4) +
                 5 * (pad256(768 * 4) + pad256(512 * 256 * 4) + 2 * pad256(512 * 4) + 3 * pad256((size_t)DEN_MAX_STEPS * 512 * 4)) +
                 5 * 2 * 2 * pad256(512 * 256 * 2) + 5 * pad256(512 * 256 * 4) + 4 * pad256(NCM * (size_t)h->max_rows * 256 * 2) + 65536 +
                 pad256(mtape.size()) + pad256(sizeof(DmUnit) * DM_MAX_UNITS) + 5 * pad256(257 * 4) +
                 2 * pad256((size_t)P->max_tiles * 256 * 128 * 4) + pad256((size_t)P->max_tiles * 2 * 4 * 32768);
  SEEME_TRY(P->arena.init(bytes));
  P->tape = P->arena.take<uint8_t>(P->tape_cta_bytes * DP_CL);
  P->d_units = P->arena.take<DpUnit>(DP_MAX_UNITS);
  P->ximg = P->arena.take<uint8_t>((size_t)P->max_tiles * XC_KB * 32768);
  P->part = P->arena.take<float>((size_t)P->max_tiles * DP_CL * 256 * 128);
  SEEME_REQUIRE(P->part, SEEME_ENOMEM, "den_persist: arena exhausted");
  P->tape_m = P->arena.take<uint8_t>(mtape.size());
  P->d_units_m = P->arena.take<DmUnit>(DM_MAX_UNITS);
  P->lat_m = P->arena.take<float>((size_t)P->max_tiles * 256 * 128);
  P->xs1_m = P->arena.take<float>((size_t)P->max_tiles * 256 * 128);
  P->skipimg = P->arena.take<uint8_t>((size_t)P->max_tiles * 2 * 4 * 32768);
  SEEME_REQUIRE(P->skipimg, SEEME_ENOMEM, "den_persist: arena exhausted (monolithic kernel)");
  SEEME_CUDA(cudaMemcpy(P->tape_m, mtape.data(), mtape.size(), cudaMemcpyHostToDevice));
  SEEME_CUDA(cudaMemcpy(P->d_units_m, munits.data(), sizeof(DmUnit) * munits.size(), cudaMemcpyHostToDevice));
  for (int l = 0; l < 5; ++l) {
    P->mu[l] = P->arena.take<float>(257);
    SEEME_REQUIRE(P->mu[l], SEEME_ENOMEM, "den_persist: arena exhausted");
    SEEME_CUDA(cudaMemcpy(P->mu[l], mu[l].data(), 257 * 4, cudaMemcpyHostToDevice));
  }
  for (int c = 0; c < DP_CL; ++c)
    SEEME_CUDA(cudaMemcpy(P->tape + (size_t)c * P->tape_cta_bytes, tapes[c].data(), P->tape_cta_bytes, cudaMemcpyHostToDevice));
  SEEME_CUDA(cudaMemcpy(P->d_units, units.data(), sizeof(DpUnit) * units.size(), cudaMemcpyHostToDevice));
  SEEME_CUDA(cudaMemset(P->ximg, 0, (size_t)P->max_tiles * XC_KB * 32768));
  for (int l = 0; l < 5; ++l) {
    P->ctab[l] = P->arena.take<float>(4 * NCM * 256 * Rp);
    P->bqkv[l] = P->arena.take<float>(768);
    P->wkov_f32[l] = P->arena.take<float>(512 * 256);
    P->bkov[l] = P->arena.take<float>(512);
    P->bk2v2[l] = P->arena.take<float>(512);
    P->tkov[l] = P->arena.take<float>((size_t)DEN_MAX_STEPS * 512);
    P->film2_ca[l] = P->arena.take<float>((size_t)DEN_MAX_STEPS * 512);
    P->film2_ff[l] = P->arena.take<float>((size_t)DEN_MAX_STEPS * 512);
    SEEME_REQUIRE(P->film2_ff[l], SEEME_ENOMEM, "den_persist: arena exhausted");
    SEEME_CUDA(cudaMemcpy(P->bqkv[l], bqkv[l].data(), 768 * 4, cudaMemcpyHostToDevice));
    SEEME_CUDA(cudaMemcpy(P->wkov_f32[l], Wqkv[l].v.data() + 256 * 256, 512 * 256 * 4, cudaMemcpyHostToDevice));
    SEEME_CUDA(cudaMemcpy(P->bkov[l], bkov[l].data(), 512 * 4, cudaMemcpyHostToDevice));
    SEEME_CUDA(cudaMemcpy(P->bk2v2[l], blkw(h, l, CA_K_B), 256 * 4, cudaMemcpyDeviceToDevice));
    SEEME_CUDA(cudaMemcpy(P->bk2v2[l] + 256, blkw(h, l, CA_V_B), 256 * 4, cudaMemcpyDeviceToDevice));
    float* k2v2 = P->arena.take<float>(512 * 256);
    SEEME_REQUIRE(k2v2, SEEME_ENOMEM, "den_persist: arena exhausted");
    SEEME_CUDA(cudaMemcpy(k2v2, blkw(h, l, CA_K_W), 256 * 256 * 4, cudaMemcpyDeviceToDevice));
    SEEME_CUDA(cudaMemcpy(k2v2 + 256 * 256, blkw(h, l, CA_V_W), 256 * 256 * 4, cudaMemcpyDeviceToDevice));
    SEEME_TRY(pack_linear(P->arena, P->Wkov[l], P->wkov_f32[l], 256, 512, 256, P->bkov[l]));
    SEEME_TRY(pack_linear(P->arena, P->Wk2v2[l], k2v2, 256, 512, 256, P->bk2v2[l]));
  }
  auto mk = [&](ActBuf& a) {
    a.ld = 256; a.f = nullptr;
    a.h = P->arena.take<__nv_bfloat16>(NCM * (size_t)h->max_rows * 256);
    a.l = P->arena.take<__nv_bfloat16>(NCM * (size_t)h->max_rows * 256);
  };
  mk(P->condb);
  mk(P->tnb);
  SEEME_REQUIRE(P->tnb.l, SEEME_ENOMEM, "den_persist: arena exhausted (cond buffers)");
  SEEME_CUDA(cudaFuncSetAttribute(den_persist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM));
  SEEME_CUDA(cudaFuncSetAttribute(den_persist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM));
  SEEME_CUDA(cudaFuncSetAttribute(den_persist_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM));
  SEEME_CUDA(cudaFuncSetAttribute(den_persist_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM));
  SEEME_CUDA(cudaFuncSetAttribute(den_mono_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM));
  SEEME_CUDA(cudaFuncSetAttribute(den_mono_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM));
  SEEME_CUDA(cudaFuncSetAttribute(den_mono_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM));
  SEEME_CUDA(cudaFuncSetAttribute(den_mono_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM));
  SEEME_CUDA(cudaDeviceSynchronize());
  return SEEME_OK;
}

void den_persist_destroy(seeme_denoiser* h) {
  if (!h->persist) return;
  h->persist->arena.release();
  delete h->persist;
  h->persist = nullptr;
}

// FiLM (mdiff_transformer.py:152-163) folded into the affine of the LayerNorm it follows:
//   LN(y) (1 + scale) + shift = n (g (1 + scale)) + (b (1 + scale) + shift),  n = the normalised row
__global__ void dp_film_fold_kernel(const float* __restrict__ film, const float* __restrict__ g, const float* __restrict__ b,
                                    float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 256) return;
  const int t = i >> 8, c = i & 255;
  const float sc = film[(size_t)t * 512 + c], sh = film[(size_t)t * 512 + 256 + c];
  out[(size_t)t * 512 + c] = g[c] * (1.0f + sc);
  out[(size_t)t * 512 + 256 + c] = b[c] * (1.0f + sc) + sh;
}

int den_persist_build_tables(seeme_denoiser* h, int n, cudaStream_t s) {
  DenPersist* P = h->persist;
  for (int l = 0; l < 5; ++l) {
    // (k | ov) of the time token: rows of [W_k | W_o W_v], bias (b_k | 0)
    SEEME_TRY(gemm_f32(gemm_params(h->temb, 256, P->wkov_f32[l], 256, P->bkov[l], P->tkov[l], 512, n, 512, 256), s));
    dp_film_fold_kernel<<<(n * 256 + 255) / 256, 256, 0, s>>>(h->film_ca[l], blkw(h, l, CA_PN_W), blkw(h, l, CA_PN_B), P->film2_ca[l], n);
    SEEME_LAUNCH_CHECK();
    dp_film_fold_kernel<<<(n * 256 + 255) / 256, 256, 0, s>>>(h->film_ff[l], blkw(h, l, FF_PN_W), blkw(h, l, FF_PN_B), P->film2_ff[l], n);
    SEEME_LAUNCH_CHECK();
  }
  return SEEME_OK;
}

static int dp_dump_trace(unsigned long long* dtrace, const char* path, cudaStream_t s) {
  std::vector<unsigned long long> host(4 * 4096);
  SEEME_CUDA(cudaStreamSynchronize(s));
  SEEME_CUDA(cudaMemcpy(host.data(), dtrace, host.size() * 8, cudaMemcpyDeviceToHost));
  FILE* f = fopen(path, "w");
  if (f) {
    for (int role = 0; role < 4; ++role)
      for (int i = 0; i < 4096 && host[role * 4096 + i]; ++i)
        fprintf(f, "%d %llu %d\n", role, host[role * 4096 + i] >> 8, (int)(host[role * 4096 + i] & 255));
    fclose(f);
  }
  return SEEME_OK;
}

int den_persist_run(seeme_denoiser* h, int mode, const float* x_in, int Nc, int B, int R, int cfg, int n_steps, float* out,
                    cudaStream_t s) {
  DenPersist* P = h->persist;
  const int tiles = cfg ? (B + 63) / 64 : (R + 127) / 128;
  SEEME_REQUIRE(tiles <= P->max_tiles, SEEME_ECAP, "den_persist: %d row tiles exceed capacity %d", tiles, P->max_tiles);
  const int rows_pad = tiles * 128;
  const int rows = Nc * R;
  // cond-token projections, once per run, on tensor cores (H3): (k | ov) for the self-attention, (key | value) of
  // text_norm(cond) for the cross-attention
  SEEME_TRY(to_bf16_split(h->cond, 256, rows, 256, P->condb.h, P->condb.l, 256, 0, s));
  for (int l = 0; l < 5; ++l) {
    ActBuf o1; o1.f = h->kvc[l]; o1.ld = 512;
    SEEME_TRY(run_linear(P->Wkov[l], P->condb, nullptr, rows, ACT_NONE, nullptr, 0, o1, 3, s));
    SEEME_TRY(layernorm256(h->cond, nullptr, 0, blkw(h, l, CA_TN_W), blkw(h, l, CA_TN_B), h->tn, rows, s));
    SEEME_TRY(to_bf16_split(h->tn, 256, rows, 256, P->tnb.h, P->tnb.l, 256, 0, s));
    ActBuf o2; o2.f = h->kv2[l]; o2.ld = 512;
    SEEME_TRY(run_linear(P->Wk2v2[l], P->tnb, nullptr, rows, ACT_NONE, nullptr, 0, o2, 3, s));
    dp_ctab_kernel<<<dim3(rows_pad / 32, 32), dim3(32, 8), 0, s>>>(h->kvc[l], h->kv2[l], P->ctab[l], Nc, B, R, cfg, rows_pad);
    SEEME_LAUNCH_CHECK();
  }
  if (h->backend == SEEME_SAMPLER_TILE) {
    DmParams m;
    memset(&m, 0, sizeof(m));
    for (int l = 0; l < 5; ++l) {
      DpLayerP& L = m.L[l];
      L.bqkv = P->bqkv[l]; L.n1g = blkw(h, l, SA_N1_W); L.n1b = blkw(h, l, SA_N1_B);
      L.b1 = blkw(h, l, SA_L1_B); L.b2 = blkw(h, l, SA_L2_B); L.n2g = blkw(h, l, SA_N2_W); L.n2b = blkw(h, l, SA_N2_B);
      L.cng = blkw(h, l, CA_N_W); L.cnb = blkw(h, l, CA_N_B); L.bcaq = blkw(h, l, CA_Q_B); L.bcaout = blkw(h, l, CA_OUT_B);
      L.bf1 = blkw(h, l, FF_L1_B); L.bf2 = blkw(h, l, FF_L2_B); L.bfout = blkw(h, l, FF_OUT_B);
      L.kt = P->tkov[l]; L.film_ca = P->film2_ca[l]; L.film_ff = P->film2_ff[l]; L.ctab = P->ctab[l];
      m.mu[l] = P->mu[l];
    }
    m.skip_b[0] = h->w[DN_LB0_B]; m.skip_b[1] = h->w[DN_LB1_B];
    m.fng = h->w[DN_NORM_W]; m.fnb = h->w[DN_NORM_B]; m.pe0 = h->w[DN_PE];
    m.tape = P->tape_m; m.units = P->d_units_m; m.n_units = P->n_units_m;
    m.lat = P->lat_m; m.xs1 = P->xs1_m; m.skipimg = P->skipimg;
    m.x_in = x_in; m.out = out; m.coef = h->d_coef; m.gscale = h->d_gscale;
    m.B = B; m.R = R; m.cfg = cfg; m.n_steps = n_steps; m.mode = mode;
    const char* tpath = getenv("SEEME_DP_TRACE");
    if (tpath && !P->trace) SEEME_CUDA(cudaMalloc(&P->trace, 4 * 4096 * 8));
    if (tpath) {
      SEEME_CUDA(cudaMemsetAsync(P->trace, 0, 4 * 4096 * 8, s));
      m.trace = P->trace;
      m.trace_step = n_steps > 1 ? 1 : 0;
    }
    {
      ProfScope prof(PROF_SAMPLER_GRAPH, s);
      const dim3 grid(tiles), block(DM_THREADS);
      switch (Nc) {
        case 1: den_mono_kernel<1><<<grid, block, DM_SMEM, s>>>(m); break;
        case 2: den_mono_kernel<2><<<grid, block, DM_SMEM, s>>>(m); break;
        case 3: den_mono_kernel<3><<<grid, block, DM_SMEM, s>>>(m); break;
        default: den_mono_kernel<4><<<grid, block, DM_SMEM, s>>>(m); break;
      }
    }
    SEEME_LAUNCH_CHECK();
    if (tpath) SEEME_TRY(dp_dump_trace(P->trace, tpath, s));
    return SEEME_OK;
  }
  DpParams p;
  memset(&p, 0, sizeof(p));
  for (int l = 0; l < 5; ++l) {
    DpLayerP& L = p.L[l];
    L.bqkv = P->bqkv[l]; L.n1g = blkw(h, l, SA_N1_W); L.n1b = blkw(h, l, SA_N1_B);
    L.b1 = blkw(h, l, SA_L1_B); L.b2 = blkw(h, l, SA_L2_B); L.n2g = blkw(h, l, SA_N2_W); L.n2b = blkw(h, l, SA_N2_B);
    L.cng = blkw(h, l, CA_N_W); L.cnb = blkw(h, l, CA_N_B); L.bcaq = blkw(h, l, CA_Q_B);
    L.bcaout = blkw(h, l, CA_OUT_B);
    L.bf1 = blkw(h, l, FF_L1_B); L.bf2 = blkw(h, l, FF_L2_B);
    L.bfout = blkw(h, l, FF_OUT_B);
    L.kt = P->tkov[l]; L.film_ca = P->film2_ca[l]; L.film_ff = P->film2_ff[l];
    L.ctab = P->ctab[l];
  }
  p.skip_b[0] = h->w[DN_LB0_B]; p.skip_b[1] = h->w[DN_LB1_B];
  p.fng = h->w[DN_NORM_W]; p.fnb = h->w[DN_NORM_B]; p.pe0 = h->w[DN_PE];
  p.tape = P->tape; p.tape_cta_bytes = P->tape_cta_bytes;
  p.units = P->d_units; p.n_units = P->n_units;
  p.ximg = P->ximg; p.part = P->part;
  p.x_in = x_in; p.out = out; p.coef = h->d_coef; p.gscale = h->d_gscale;
  p.B = B; p.R = R; p.cfg = cfg; p.n_steps = n_steps; p.mode = mode; p.rows_pad = rows_pad;
  const char* trace_path = getenv("SEEME_DP_TRACE");     // diagnostics: event trace of cluster 0 / rank 0 at step 1
  if (trace_path && !P->trace) SEEME_CUDA(cudaMalloc(&P->trace, 4 * 4096 * 8));
  if (trace_path) {
    SEEME_CUDA(cudaMemsetAsync(P->trace, 0, 4 * 4096 * 8, s));
    p.trace = P->trace;
    p.trace_step = n_steps > 1 ? 1 : 0;
  }
  {
    ProfScope prof(PROF_SAMPLER_GRAPH, s);
    const dim3 grid(tiles * DP_CL), block(DP_THREADS);
    switch (Nc) {
      case 1: den_persist_kernel<1><<<grid, block, DP_SMEM, s>>>(p); break;
      case 2: den_persist_kernel<2><<<grid, block, DP_SMEM, s>>>(p); break;
      case 3: den_persist_kernel<3><<<grid, block, DP_SMEM, s>>>(p); break;
      default: den_persist_kernel<4><<<grid, block, DP_SMEM, s>>>(p); break;
    }
  }
  SEEME_LAUNCH_CHECK();
  if (trace_path) SEEME_TRY(dp_dump_trace(P->trace, trace_path, s));
  return SEEME_OK;
}

}  // namespace seeme
