// Shared helpers for the seeme_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include "../../include/seeme_b200.h"

namespace seeme {

// ---- error plumbing (thread-local message behind seeme_last_error) --------------------------
void set_error(const char* fmt, ...);
extern unsigned long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += (unsigned long long)n; }

#define SEEME_CUDA(expr)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      seeme::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SEEME_ECUDA;                                                                         \
    }                                                                                             \
  } while (0)

#define SEEME_REQUIRE(cond, code, ...)   \
  do {                                   \
    if (!(cond)) {                       \
      seeme::set_error(__VA_ARGS__);     \
      return (code);                     \
    }                                    \
  } while (0)

#define SEEME_TRY(expr)          \
  do {                           \
    int _r = (expr);             \
    if (_r != SEEME_OK) return _r; \
  } while (0)

#define SEEME_LAUNCH_CHECK()                                                                 \
  do {                                                                                       \
    cudaError_t _e = cudaPeekAtLastError();                                                  \
    if (_e != cudaSuccess) {                                                                 \
      seeme::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,                 \
                       cudaGetErrorString(_e));                                              \
      return SEEME_ECUDA;                                                                    \
    }                                                                                        \
    seeme::count_launch();                                                                   \
  } while (0)

// ---- optional per-kernel-class device timing (bench.py's live roofline measurement) ----------
// When enabled with seeme_prof_enable(1), each instrumented launch is bracketed by a cudaEvent pair on
// its own stream; seeme_prof_read() synchronises and returns the summed duration and launch count.
enum ProfId { PROF_POINTNET_GEMM = 0, PROF_SMPL_SKIN = 1, PROF_SMPL_POSE = 2, PROF_SAMPLER_GRAPH = 3, PROF_VAE_ATTN = 4,
              PROF_UMMA_GEMM = 5, PROF_POINTNET_FUSED = 6, PROF_COUNT = 8 };
extern bool g_prof_on;
void prof_begin(int id, cudaStream_t s);
void prof_end(int id, cudaStream_t s);
struct ProfScope {
  int id; cudaStream_t s; bool on;
  ProfScope(int id_, cudaStream_t s_) : id(id_), s(s_), on(g_prof_on && id_ >= 0) { if (on) prof_begin(id, s); }
  ~ProfScope() { if (on) prof_end(id, s); }
};

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------
// Kernels of a dependent chain call pdl_prologue() before touching any data produced by earlier kernels: it lets
// the NEXT kernel of the stream/graph start its own prologue (launch latency, barrier init, TMEM allocation, tensor-map
// prefetch) while this one is still running, then blocks until the PREVIOUS kernel has completed and flushed.  Both
// instructions are no-ops for a kernel launched without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
extern bool g_pdl_on;   // SEEME_PDL=1 enables the attribute (default: plain stream order)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- a tiny bump allocator over one cudaMalloc'd slab (everything allocated at create) -------
struct Arena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  int init(size_t bytes) {
    cap = bytes;
    used = 0;
    cudaError_t e = cudaMalloc((void**)&base, bytes);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
      base = nullptr;
      return SEEME_ENOMEM;
    }
    return SEEME_OK;
  }
  template <typename T>
  T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    if (used + bytes > cap) return nullptr;
    T* p = reinterpret_cast<T*>(base + used);
    used += bytes;
    return p;
  }
  void release() {
    if (base) cudaFree(base);
    base = nullptr;
  }
};
inline size_t pad256(size_t bytes) { return (bytes + 255) & ~size_t(255); }

// Experiment knobs (tile shapes, ring depths, retired kernel variants, PDL ...) are read in experimental builds only
// (python -m seeme_b200.build with SEEME_EXPERIMENTAL=1 in the environment adds -DSEEME_EXPERIMENTAL); the product library
// ignores them.  The knobs a user may set are listed in INTEGRATION.md.
inline const char* seeme_exp_env(const char* name) {
#ifdef SEEME_EXPERIMENTAL
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

constexpr int D_MODEL = 256;
constexpr int NUM_SMS = 148;

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

// order-preserving float <-> uint mapping (for atomicMax / redux on floats)
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_SILU = 3 };
// NOTE: inside an unrolled per-element loop this switch compiles to one indirect branch (BRX) per element (an if-chain
// is turned back into the same jump table) -- callers on a hot path test the common selectors once, outside the loop
__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_GELU: return gelu_erf(v);
    case ACT_SILU: return silu(v);
    default: return v;
  }
}

// ---- generic fp32 linear: Y = act(pre(X) W^T + bias) (+R) ------------------------------------
struct GemmP {
  const float* X; int ldx;      // [M,K]
  const float* W; int ldw;      // [N,K] (nn.Linear weight layout)
  const float* bias;            // [N] or [G,N] (per row-group); nullable
  int bias_group_rows;          // 0: one bias; else bias row = m / bias_group_rows
  const float* R; int ldr;      // residual added after the activation; nullable
  float* Y; int ldy;
  int M, N, K;
  int pre_act;                  // Act applied to X on load (ACT_NONE/ACT_RELU/ACT_SILU)
  int act;                      // Act applied to (acc + bias)
  int accumulate;               // Y += instead of Y =
  int prof_id;                  // ProfId + 1 to time this launch, 0 = not timed
};
inline GemmP gemm_params(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y,
                         int ldy, int M, int N, int K) {
  GemmP p;
  memset(&p, 0, sizeof(p));
  p.X = X; p.ldx = ldx; p.W = W; p.ldw = ldw; p.bias = bias; p.Y = Y; p.ldy = ldy;
  p.M = M; p.N = N; p.K = K;
  return p;
}
int gemm_f32(const GemmP& p, cudaStream_t s);

// ---- row-wise ops on [rows,256] (one warp per row) -------------------------------------------
// y = LN(x (+ r)) * g + b      (r nullable; r_group_rows > 0: r row = row / r_group_rows, i.e. a
// per-sample vector broadcast over frames)
int layernorm256(const float* x, const float* r, int r_group_rows, const float* g, const float* b,
                 float* y, int rows, cudaStream_t s);

// p[i] *= s on the default stream (create-time weight folding)
void scale_kernel_launch(float* p, size_t n, float s);

}  // namespace seeme
