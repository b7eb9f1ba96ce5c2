// Denoiser handle shared by the two sampler back-ends: the CUDA-graph kernel chain (denoiser.cu) and the persistent
// cluster kernel (den_persist.cu).
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace seeme {
// tensor order of seeme_denoiser_create (include/seeme_b200.h; seeme_b200/ops.py:denoiser_keys)
enum { DN_TE1_W = 0, DN_TE1_B, DN_TE2_W, DN_TE2_B, DN_PE, DN_NORM_W, DN_NORM_B, DN_LB0_W, DN_LB0_B, DN_LB1_W, DN_LB1_B,
       DN_BLK = 11, DN_BLK_STRIDE = 38 };
enum { SA_IN_W = 0, SA_IN_B, SA_OUT_W, SA_OUT_B, SA_L1_W, SA_L1_B, SA_L2_W, SA_L2_B, SA_N1_W, SA_N1_B, SA_N2_W, SA_N2_B,
       CA_N_W = 12, CA_N_B, CA_TN_W, CA_TN_B, CA_Q_W, CA_Q_B, CA_K_W, CA_K_B, CA_V_W, CA_V_B, CA_EMB_W, CA_EMB_B,
       CA_PN_W, CA_PN_B, CA_OUT_W, CA_OUT_B,
       FF_L1_W = 28, FF_L1_B, FF_L2_W, FF_L2_B, FF_EMB_W, FF_EMB_B, FF_PN_W, FF_PN_B, FF_OUT_W, FF_OUT_B };
constexpr int DEN_MAX_STEPS = 1024;
struct DenPersist;   // state of the persistent cluster sampler (den_persist.cu)
}  // namespace seeme

struct seeme_denoiser {
  int device = 0, max_rows = 0;
  seeme::Arena arena;
  float* w[SEEME_DENOISER_NUM_TENSORS];
  // per-timestep tables (H4)
  float *sinus, *t1, *temb, *tkv[5], *film_ca[5], *film_ff[5];
  std::vector<int> table_ts;      // timesteps the tables currently hold
  // per-run cond projections (H3) and activations
  float *cond, *tn, *kvc[5], *kv2[5];
  float *lat, *qkv, *t0, *caq, *y;
  // tcgen05 path: packed (hi, lo) weights and activations carried as fp32 (residuals, row-wise ops) and/or
  // split bf16 (GEMM A operands)
  int npass = 3;
  seeme::PackedLinear Wqkv[5], Wout[5], Wl1[5], Wl2[5], Wcaq[5], Wcaout[5], Wf1[5], Wf2[5], Wfout[5], Wskip[2];
  seeme::ActBuf x, L[5], att, x1, x2, ln, hb, ff, g1;
  float *d_coef, *d_gscale;
  std::vector<float> coef_host;   // what d_coef / d_gscale currently hold
  float gscale_host = -1.f;
  // graph cache
  cudaStream_t cap_stream = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int g_Nc = -1, g_B = -1, g_cfg = -1, g_steps = -1, g_kernels = 0;
  bool use_graph = true;
  // persistent cluster sampler (den_persist.cu); null when disabled (SEEME_SAMPLER=graph)
  seeme::DenPersist* persist = nullptr;
  int backend = SEEME_SAMPLER_PERSISTENT;
};

static inline float* blkw(seeme_denoiser* h, int l, int k) { return h->w[seeme::DN_BLK + seeme::DN_BLK_STRIDE * l + k]; }

namespace seeme {
// den_persist.cu
int den_persist_create(seeme_denoiser* h);
void den_persist_destroy(seeme_denoiser* h);
// builds the per-timestep tables of the persistent sampler from h->temb (n rows)
int den_persist_build_tables(seeme_denoiser* h, int n, cudaStream_t s);
// mode 0: 50-step sampler (x_in = x_T [B,256], out = z [B,256]); mode 1: one denoiser call (x_in = sample [R,256], out = eps [R,256])
int den_persist_run(seeme_denoiser* h, int mode, const float* x_in, int Nc, int B, int R, int cfg, int n_steps, float* out,
                    cudaStream_t s);
}  // namespace seeme
