// Error state, the generic fp32 linear kernel and row-wise LayerNorm shared by all stages.
#include "common.cuh"
#include <stdlib.h>

namespace seeme {

static thread_local char g_err[1024] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool g_prof_on = false;
bool g_pdl_on = seeme_exp_env("SEEME_PDL") && seeme_exp_env("SEEME_PDL")[0] == '1';   // measured: no gain inside the sampler graph
namespace {
struct ProfPair { cudaEvent_t a, b; };
std::vector<ProfPair> g_prof_pairs[PROF_COUNT];
cudaEvent_t g_prof_open[PROF_COUNT];
double g_prof_ms[PROF_COUNT];
long long g_prof_n[PROF_COUNT];
}  // namespace
void prof_begin(int id, cudaStream_t s) {
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, s);
  g_prof_open[id] = e;
}
void prof_end(int id, cudaStream_t s) {
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, s);
  g_prof_pairs[id].push_back({g_prof_open[id], e});
}

// ------------------------------------------------------------------------------------------------
// gemm_f32: register-tiled CUDA-core SGEMM for the "TN" case both operands K-contiguous
// (activations [M,K] row-major, nn.Linear weight [N,K] row-major).  Used where a contraction must
// stay bit-level fp32 (small-K prologues, reference-precision mode) -- the tcgen05 path in
// umma_gemm.cu takes the large dense contractions.
// ------------------------------------------------------------------------------------------------
// GK: K depth of one shared-memory stage.  The few-row GEMMs (per-sample biases, cond-token projections, 256-row latent
// maps) are bound by the latency of the K loop (one L2 round trip per stage), so they take small tiles and 64-deep stages:
// same k order, hence bit-identical sums, in a quarter of the iterations and four times the CTAs.
template <int BM, int BN, int TM, int TN, int GK = 16>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_f32_kernel(const GemmP p) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int HM = TM / 2, HN = TN / 2;   // each thread owns two HM-row and two HN-col strips
  __shared__ __align__(16) float As[2][GK][BM + 4];
  __shared__ __align__(16) float Bs[2][GK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const bool vecA = (p.K % 4 == 0) && (p.ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.X) & 15) == 0);
  const bool vecB = (p.K % 4 == 0) && (p.ldw % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.W) & 15) == 0);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  constexpr int LA = (BM * GK / 4 + NT - 1) / NT;   // float4 loads per thread for the A tile
  constexpr int LB = (BN * GK / 4 + NT - 1) / NT;
  float4 ra[LA], rb[LB];

  auto load_tile = [&](int k0) {
#pragma unroll
    for (int l = 0; l < LA; ++l) {
      int idx = tid + l * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < BM * GK / 4) {
        int row = idx / (GK / 4), kq = (idx % (GK / 4)) * 4;
        int gm = m0 + row, gk = k0 + kq;
        if (gm < p.M) {
          const float* src = p.X + (size_t)gm * p.ldx + gk;
          if (vecA && gk + 3 < p.K) {
            v = __ldg(reinterpret_cast<const float4*>(src));
          } else {
            if (gk + 0 < p.K) v.x = __ldg(src + 0);
            if (gk + 1 < p.K) v.y = __ldg(src + 1);
            if (gk + 2 < p.K) v.z = __ldg(src + 2);
            if (gk + 3 < p.K) v.w = __ldg(src + 3);
          }
          if (p.pre_act) { v.x = apply_act(v.x, p.pre_act); v.y = apply_act(v.y, p.pre_act); v.z = apply_act(v.z, p.pre_act); v.w = apply_act(v.w, p.pre_act); }
        }
      }
      ra[l] = v;
    }
#pragma unroll
    for (int l = 0; l < LB; ++l) {
      int idx = tid + l * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < BN * GK / 4) {
        int row = idx / (GK / 4), kq = (idx % (GK / 4)) * 4;
        int gn = n0 + row, gk = k0 + kq;
        if (gn < p.N) {
          const float* src = p.W + (size_t)gn * p.ldw + gk;
          if (vecB && gk + 3 < p.K) {
            v = __ldg(reinterpret_cast<const float4*>(src));
          } else {
            if (gk + 0 < p.K) v.x = __ldg(src + 0);
            if (gk + 1 < p.K) v.y = __ldg(src + 1);
            if (gk + 2 < p.K) v.z = __ldg(src + 2);
            if (gk + 3 < p.K) v.w = __ldg(src + 3);
          }
        }
      }
      rb[l] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int l = 0; l < LA; ++l) {
      int idx = tid + l * NT;
      if (idx < BM * GK / 4) {
        int row = idx / (GK / 4), kq = (idx % (GK / 4)) * 4;
        As[buf][kq + 0][row] = ra[l].x; As[buf][kq + 1][row] = ra[l].y;
        As[buf][kq + 2][row] = ra[l].z; As[buf][kq + 3][row] = ra[l].w;
      }
    }
#pragma unroll
    for (int l = 0; l < LB; ++l) {
      int idx = tid + l * NT;
      if (idx < BN * GK / 4) {
        int row = idx / (GK / 4), kq = (idx % (GK / 4)) * 4;
        Bs[buf][kq + 0][row] = rb[l].x; Bs[buf][kq + 1][row] = rb[l].y;
        Bs[buf][kq + 2][row] = rb[l].z; Bs[buf][kq + 3][row] = rb[l].w;
      }
    }
  };

  const int nk = (p.K + GK - 1) / GK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < HM; ++i) {
        a[i] = As[buf][k][ty * HM + i];
        a[HM + i] = As[buf][k][BM / 2 + ty * HM + i];
      }
#pragma unroll
      for (int j = 0; j < HN; ++j) {
        b[j] = Bs[buf][k][tx * HN + j];
        b[HN + j] = Bs[buf][k][BN / 2 + tx * HN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i < HM ? ty * HM + i : BM / 2 + ty * HM + (i - HM));
    if (m >= p.M) continue;
    const float* brow = p.bias ? p.bias + (p.bias_group_rows ? (size_t)(m / p.bias_group_rows) * p.N : 0) : nullptr;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (j < HN ? tx * HN + j : BN / 2 + tx * HN + (j - HN));
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (brow) v += brow[n];
      v = apply_act(v, p.act);
      if (p.R) v += p.R[(size_t)m * p.ldr + n];
      float* y = p.Y + (size_t)m * p.ldy + n;
      if (p.accumulate) v += *y;
      *y = v;
    }
  }
}

int gemm_f32(const GemmP& p, cudaStream_t s) {
  if (p.M <= 0 || p.N <= 0) return SEEME_OK;
  ProfScope prof(p.prof_id - 1, s);
  if (p.M >= 2048 && p.N >= 128) {
    dim3 grid((p.N + 127) / 128, (p.M + 127) / 128);
    gemm_f32_kernel<128, 128, 8, 8><<<grid, 256, 0, s>>>(p);
  } else if ((long long)p.M * p.N <= 1024ll * 512 && p.K >= 128) {
    dim3 grid((p.N + 31) / 32, (p.M + 31) / 32);
    gemm_f32_kernel<32, 32, 2, 2, 64><<<grid, 256, 0, s>>>(p);
  } else {
    dim3 grid((p.N + 63) / 64, (p.M + 63) / 64);
    gemm_f32_kernel<64, 64, 4, 4><<<grid, 256, 0, s>>>(p);
  }
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over 256 channels, one warp per row, two-pass in registers (matches
// F.layer_norm: biased variance, eps 1e-5 inside the sqrt).
// ------------------------------------------------------------------------------------------------
__global__ void layernorm256_kernel(const float* __restrict__ x, const float* __restrict__ r, int r_group_rows,
                                    const float* __restrict__ g, const float* __restrict__ b,
                                    float* __restrict__ y, int rows) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xp = reinterpret_cast<const float4*>(x + (size_t)row * 256);
  float4 v0 = xp[lane], v1 = xp[lane + 32];
  if (r) {
    const float4* rp = reinterpret_cast<const float4*>(r + (size_t)(r_group_rows ? row / r_group_rows : row) * 256);
    float4 a = rp[lane], c = rp[lane + 32];
    v0.x += a.x; v0.y += a.y; v0.z += a.z; v0.w += a.w;
    v1.x += c.x; v1.y += c.y; v1.z += c.z; v1.w += c.w;
  }
  float s = v0.x + v0.y + v0.z + v0.w + v1.x + v1.y + v1.z + v1.w;
  const float mean = warp_sum(s) * (1.0f / 256.0f);
  float d, q = 0.f;
  d = v0.x - mean; q += d * d; d = v0.y - mean; q += d * d; d = v0.z - mean; q += d * d; d = v0.w - mean; q += d * d;
  d = v1.x - mean; q += d * d; d = v1.y - mean; q += d * d; d = v1.z - mean; q += d * d; d = v1.w - mean; q += d * d;
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / 256.0f) + 1e-5f);
  const float4* gp = reinterpret_cast<const float4*>(g);
  const float4* bp = reinterpret_cast<const float4*>(b);
  float4 g0 = gp[lane], g1 = gp[lane + 32], b0 = bp[lane], b1 = bp[lane + 32];
  float4 o0, o1;
  o0.x = (v0.x - mean) * rstd * g0.x + b0.x; o0.y = (v0.y - mean) * rstd * g0.y + b0.y;
  o0.z = (v0.z - mean) * rstd * g0.z + b0.z; o0.w = (v0.w - mean) * rstd * g0.w + b0.w;
  o1.x = (v1.x - mean) * rstd * g1.x + b1.x; o1.y = (v1.y - mean) * rstd * g1.y + b1.y;
  o1.z = (v1.z - mean) * rstd * g1.z + b1.z; o1.w = (v1.w - mean) * rstd * g1.w + b1.w;
  float4* yp = reinterpret_cast<float4*>(y + (size_t)row * 256);
  yp[lane] = o0;
  yp[lane + 32] = o1;
}

int layernorm256(const float* x, const float* r, int r_group_rows, const float* g, const float* b, float* y,
                 int rows, cudaStream_t s) {
  if (rows <= 0) return SEEME_OK;
  layernorm256_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, r, r_group_rows, g, b, y, rows);
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

__global__ void scale_kernel(float* p, size_t n, float s) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] *= s;
}
void scale_kernel_launch(float* p, size_t n, float s) {
  scale_kernel<<<(unsigned)((n + 255) / 256), 256>>>(p, n, s);
}

__global__ void ddim_step_kernel(const float* __restrict__ eps, const float* __restrict__ x, float* __restrict__ out,
                                 size_t n, float c0, float c1, float c2, float c3) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float e = eps[i];   // unfused fp32 ops in diffusers' order
  const float x0 = __fdiv_rn(__fsub_rn(x[i], __fmul_rn(c0, e)), c1);
  out[i] = __fadd_rn(__fmul_rn(c2, x0), __fmul_rn(c3, e));
}

}  // namespace seeme

extern "C" {
int seeme_abi_version(void) { return SEEME_ABI_VERSION; }
const char* seeme_last_error(void) { return seeme::g_err; }
unsigned long long seeme_launch_count(void) { return seeme::g_launch_count; }

int seeme_prof_enable(int on) {
  seeme::g_prof_on = on != 0;
  return SEEME_OK;
}
int seeme_prof_read(int id, double* total_ms, long long* count) {
  using namespace seeme;
  SEEME_REQUIRE(id >= 0 && id < PROF_COUNT && total_ms && count, SEEME_EINVAL, "seeme_prof_read: bad argument");
  for (auto& pr : g_prof_pairs[id]) {
    SEEME_CUDA(cudaEventSynchronize(pr.b));
    float ms = 0.f;
    SEEME_CUDA(cudaEventElapsedTime(&ms, pr.a, pr.b));
    g_prof_ms[id] += ms;
    g_prof_n[id] += 1;
    cudaEventDestroy(pr.a);
    cudaEventDestroy(pr.b);
  }
  g_prof_pairs[id].clear();
  *total_ms = g_prof_ms[id];
  *count = g_prof_n[id];
  g_prof_ms[id] = 0.0;
  g_prof_n[id] = 0;
  return SEEME_OK;
}

int seeme_ddim_step(const float* eps, const float* sample, float* prev, size_t n, float c0, float c1, float c2,
                    float c3, void* stream) {
  if (n == 0) return SEEME_OK;
  SEEME_REQUIRE(eps && sample && prev, SEEME_EINVAL, "seeme_ddim_step: null pointer");
  seeme::ddim_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(eps, sample, prev, n, c0, c1, c2, c3);
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}
}
