// SMPL forward (smplx 0.1.28 SMPL.forward / lbs.py, restated in SURVEY App. C): Rodrigues,
// shape+pose blendshapes, 24-joint kinematic chain, linear-blend skinning over 6890 vertices,
// plus aa_to_quat (mld/utils/geometry2.py:33-54) and the float64 renorm prologue
// (mld/data/EgoBody.py:151-157).
//
// Two kernels per call:
//  smpl_pose_kernel  one warp per frame, lane j = joint j: Rodrigues, rest joints from the folded
//                    regressor J = Jt + Jd.beta (J_regressor.(v_template +
//                    shapedirs.beta) is linear in beta, so a [24,3] + [24,3,10] table replaces the
//                    [24,6890] contraction), level-synchronous parent chain via warp shuffles,
//                    skinning transforms A_j (3x4), posed joints, quaternion, and the blend
//                    coefficient row c = [beta(10) | vec(R_1..23 - I)(207) | 0-pad] for kernel 2.
//  smpl_skin_kernel  thread = vertex, CTA = 128 vertices x FT frames:
//                    v_posed = v_template + Basis^T c   (Basis = [shapedirs | posedirs], K = 217,
//                    stored as [K][3][Vpad] planes so a warp reads 128 contiguous bytes per k),
//                    then T_v = sum_{<=4} w A_j (SMPL's skinning weights have <=4 non-zeros per vertex;
//                    a dense-24 layout is used when a model violates that) and the 3x4 apply.
#include "common.cuh"
#include "smpl_tc.cuh"
#include <stdlib.h>

namespace seeme {

constexpr int SV = 6890, SJ = 24, SVP = 6912 /* 54 * 128 */, SK = 217, SKP = 224;
constexpr int FT = 16;   // frames per CTA in the skinning kernel

struct SmplTopo {
  int parent[SJ];
  int depth[SJ];
  int max_depth;
};

__device__ __forceinline__ void rodrigues(float rx, float ry, float rz, float* R) {
  // smplx batch_rodrigues: angle = ||r + 1e-8||, dir = r / angle, R = I + sin K + (1 - cos) K K
  const float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
  const float angle = sqrtf(ax * ax + ay * ay + az * az);
  const float x = rx / angle, y = ry / angle, z = rz / angle;
  float s, c;
  sincosf(angle, &s, &c);
  const float t = 1.0f - c;
  // K = [[0,-z,y],[z,0,-x],[-y,x,0]];  K K = [[-(y2+z2), xy, xz],[xy, -(x2+z2), yz],[xz, yz, -(x2+y2)]]
  R[0] = 1.0f + t * (-(y * y + z * z)); R[1] = -s * z + t * (x * y);          R[2] = s * y + t * (x * z);
  R[3] = s * z + t * (x * y);           R[4] = 1.0f + t * (-(x * x + z * z)); R[5] = -s * x + t * (y * z);
  R[6] = -s * y + t * (x * z);          R[7] = s * x + t * (y * z);           R[8] = 1.0f + t * (-(x * x + y * y));
}

// Inputs either as separate tensors (feats == nullptr) or as the normalised feature row
// feats[F,Dn] with float64 stats (fused renorm): go = m[0:3], body = m[3:3+n_body] (zero padded
// to 69), transl = m[Dn-3:Dn].
__global__ void __launch_bounds__(256) smpl_pose_kernel(
    const float* __restrict__ betas, const float* __restrict__ body_pose, const float* __restrict__ global_orient,
    const float* __restrict__ transl, const float* __restrict__ feats, int Dn, const double* __restrict__ mean,
    const double* __restrict__ stdv, int n_body, double* __restrict__ m_out, const float* __restrict__ Jt,
    const float* __restrict__ Jd, const SmplTopo topo, int F, float* __restrict__ A_out, float* __restrict__ coef_out,
    float* __restrict__ joints_out, float* __restrict__ quat_out) {
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= F) return;
  const int j = lane < SJ ? lane : SJ - 1;   // lanes 24..31 shadow joint 23 (no stores)
  const bool active = lane < SJ;

  float r[3], tr[3] = {0.f, 0.f, 0.f};
  if (feats) {
    const float* row = feats + (size_t)f * Dn;
    // renorm in float64 then .float(), like feats*std+mean with float64 stats followed by .float()
    for (int i = lane; i < Dn; i += 32) {
      const double m = (double)row[i] * stdv[i] + mean[i];
      if (m_out) m_out[(size_t)f * Dn + i] = m;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int idx = 3 * j + c;
      r[c] = (idx < 3 + n_body) ? (float)((double)row[idx] * stdv[idx] + mean[idx]) : 0.f;
      const int ti = Dn - 3 + c;
      tr[c] = (float)((double)row[ti] * stdv[ti] + mean[ti]);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      r[c] = (j == 0) ? global_orient[(size_t)f * 3 + c] : body_pose[(size_t)f * 69 + 3 * (j - 1) + c];
      if (transl) tr[c] = transl[(size_t)f * 3 + c];
    }
  }
  float R[9];
  rodrigues(r[0], r[1], r[2], R);

  // rest-pose joint of this lane: J = Jt + Jd . beta
  float bt[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) bt[k] = betas[(size_t)f * 10 + k];
  float Jj[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float a = Jt[j * 3 + c];
#pragma unroll
    for (int k = 0; k < 10; ++k) a = fmaf(Jd[(j * 3 + c) * 10 + k], bt[k], a);
    Jj[c] = a;
  }
  const int par = topo.parent[j] < 0 ? 0 : topo.parent[j];
  float rel[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float pj = __shfl_sync(0xffffffffu, Jj[c], par);
    rel[c] = (topo.parent[j] < 0) ? Jj[c] : Jj[c] - pj;
  }
  // G = [R | rel] for the root; children: G = G_parent * [R | rel], one tree level at a time
  float G[12];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    G[a * 4 + 0] = R[a * 3 + 0]; G[a * 4 + 1] = R[a * 3 + 1]; G[a * 4 + 2] = R[a * 3 + 2]; G[a * 4 + 3] = rel[a];
  }
  const int my_depth = topo.depth[j];
  for (int d = 1; d <= topo.max_depth; ++d) {
    float P[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) P[e] = __shfl_sync(0xffffffffu, G[e], par);
    if (my_depth == d) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int b = 0; b < 3; ++b)
          G[a * 4 + b] = P[a * 4 + 0] * R[0 * 3 + b] + P[a * 4 + 1] * R[1 * 3 + b] + P[a * 4 + 2] * R[2 * 3 + b];
        G[a * 4 + 3] = P[a * 4 + 0] * rel[0] + P[a * 4 + 1] * rel[1] + P[a * 4 + 2] * rel[2] + P[a * 4 + 3];
      }
    }
  }
  if (active) {
    float* Ao = A_out + ((size_t)f * SJ + j) * 12;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      Ao[a * 4 + 0] = G[a * 4 + 0]; Ao[a * 4 + 1] = G[a * 4 + 1]; Ao[a * 4 + 2] = G[a * 4 + 2];
      // A = G - pad(G [J;0]): translation column minus G_R J; the global translation is folded in
      Ao[a * 4 + 3] = G[a * 4 + 3] - (G[a * 4 + 0] * Jj[0] + G[a * 4 + 1] * Jj[1] + G[a * 4 + 2] * Jj[2]) + tr[a];
      joints_out[((size_t)f * SJ + j) * 3 + a] = G[a * 4 + 3] + tr[a];
    }
    float* co = coef_out + (size_t)f * SKP;
    if (j == 0) {
#pragma unroll
      for (int k = 0; k < 10; ++k) co[k] = bt[k];
#pragma unroll
      for (int k = SK; k < SKP; ++k) co[k] = 0.f;
      if (quat_out) {
        // aa_to_quat: norm(theta + 1e-8), normalized = theta / angle, half-angle cos/sin
        const float ax = r[0] + 1e-8f, ay = r[1] + 1e-8f, az = r[2] + 1e-8f;
        const float ang = sqrtf(ax * ax + ay * ay + az * az);
        float s, c;
        sincosf(ang * 0.5f, &s, &c);
        float* qo = quat_out + (size_t)f * 4;
        qo[0] = c; qo[1] = s * (r[0] / ang); qo[2] = s * (r[1] / ang); qo[3] = s * (r[2] / ang);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 9; ++e) co[10 + (j - 1) * 9 + e] = R[e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
    }
  }
}

// v_posed + skinning.  grid (SVP/128, ceil(F/FT)), 128 threads.
template <int NNZ>
__global__ void __launch_bounds__(128) smpl_skin_kernel(const float* __restrict__ basis /*[SK][3][SVP]*/,
                                                        const float* __restrict__ vt /*[3][SVP]*/,
                                                        const float* __restrict__ lbs_w /*[SVP][NNZ]*/,
                                                        const unsigned char* __restrict__ lbs_i /*[SVP][NNZ]*/,
                                                        const float* __restrict__ A /*[F][24][12]*/,
                                                        const float* __restrict__ coef /*[F][SKP]*/, int F,
                                                        float* __restrict__ verts /*[F][SV][3]*/) {
  __shared__ __align__(16) float cs[SK][FT];        // coef transposed: [k][frame]
  __shared__ __align__(16) float As[FT][SJ][12];
  const int f0 = blockIdx.y * FT;
  const int nf = min(FT, F - f0);
  const int tid = threadIdx.x;
  for (int i = tid; i < FT * SK; i += 128) {
    const int ff = i / SK, k = i % SK;
    cs[k][ff] = ff < nf ? coef[(size_t)(f0 + ff) * SKP + k] : 0.f;
  }
  for (int i = tid; i < FT * SJ * 12; i += 128) {
    const int ff = i / (SJ * 12);
    (&As[0][0][0])[i] = ff < nf ? A[(size_t)f0 * SJ * 12 + i] : 0.f;
  }
  __syncthreads();
  const int v = blockIdx.x * 128 + tid;
  float acc[FT][3];
  {
    const float x = vt[v], y = vt[SVP + v], z = vt[2 * SVP + v];
#pragma unroll
    for (int ff = 0; ff < FT; ++ff) { acc[ff][0] = x; acc[ff][1] = y; acc[ff][2] = z; }
  }
  const float* bp = basis + v;
#pragma unroll 2
  for (int k = 0; k < SK; ++k) {
    const float b0 = __ldg(bp), b1 = __ldg(bp + SVP), b2 = __ldg(bp + 2 * SVP);
    bp += 3 * SVP;
    const float4* cr = reinterpret_cast<const float4*>(&cs[k][0]);
#pragma unroll
    for (int q = 0; q < FT / 4; ++q) {
      const float4 c = cr[q];
      acc[q * 4 + 0][0] = fmaf(c.x, b0, acc[q * 4 + 0][0]); acc[q * 4 + 0][1] = fmaf(c.x, b1, acc[q * 4 + 0][1]); acc[q * 4 + 0][2] = fmaf(c.x, b2, acc[q * 4 + 0][2]);
      acc[q * 4 + 1][0] = fmaf(c.y, b0, acc[q * 4 + 1][0]); acc[q * 4 + 1][1] = fmaf(c.y, b1, acc[q * 4 + 1][1]); acc[q * 4 + 1][2] = fmaf(c.y, b2, acc[q * 4 + 1][2]);
      acc[q * 4 + 2][0] = fmaf(c.z, b0, acc[q * 4 + 2][0]); acc[q * 4 + 2][1] = fmaf(c.z, b1, acc[q * 4 + 2][1]); acc[q * 4 + 2][2] = fmaf(c.z, b2, acc[q * 4 + 2][2]);
      acc[q * 4 + 3][0] = fmaf(c.w, b0, acc[q * 4 + 3][0]); acc[q * 4 + 3][1] = fmaf(c.w, b1, acc[q * 4 + 3][1]); acc[q * 4 + 3][2] = fmaf(c.w, b2, acc[q * 4 + 3][2]);
    }
  }
  if (v >= SV) return;
  float w[NNZ];
  int ji[NNZ];
#pragma unroll
  for (int n = 0; n < NNZ; ++n) { w[n] = lbs_w[(size_t)v * NNZ + n]; ji[n] = lbs_i[(size_t)v * NNZ + n]; }
#pragma unroll
  for (int ff = 0; ff < FT; ++ff) {
    if (ff >= nf) break;
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
#pragma unroll
    for (int n = 0; n < NNZ; ++n) {
      const float4* a = reinterpret_cast<const float4*>(&As[ff][ji[n]][0]);
      const float4 a0 = a[0], a1 = a[1], a2 = a[2];
      T[0] = fmaf(w[n], a0.x, T[0]); T[1] = fmaf(w[n], a0.y, T[1]); T[2] = fmaf(w[n], a0.z, T[2]); T[3] = fmaf(w[n], a0.w, T[3]);
      T[4] = fmaf(w[n], a1.x, T[4]); T[5] = fmaf(w[n], a1.y, T[5]); T[6] = fmaf(w[n], a1.z, T[6]); T[7] = fmaf(w[n], a1.w, T[7]);
      T[8] = fmaf(w[n], a2.x, T[8]); T[9] = fmaf(w[n], a2.y, T[9]); T[10] = fmaf(w[n], a2.z, T[10]); T[11] = fmaf(w[n], a2.w, T[11]);
    }
    const float x = acc[ff][0], y = acc[ff][1], z = acc[ff][2];
    float* o = verts + ((size_t)(f0 + ff) * SV + v) * 3;
    o[0] = T[0] * x + T[1] * y + T[2] * z + T[3];
    o[1] = T[4] * x + T[5] * y + T[6] * z + T[7];
    o[2] = T[8] * x + T[9] * y + T[10] * z + T[11];
  }
}

// ---- create-time re-layout kernels -------------------------------------------------------------
// Jfold[j][c] (c<3: J_regressor . v_template ; c = 3 + cc*10 + k: J_regressor . shapedirs[:,cc,k])
__global__ void smpl_jfold_kernel(const float* __restrict__ Jreg, const float* __restrict__ vt,
                                  const float* __restrict__ shapedirs, float* __restrict__ Jt, float* __restrict__ Jd) {
  const int j = blockIdx.x, col = blockIdx.y;   // col in [0,33)
  double s = 0.0;
  for (int v = threadIdx.x; v < SV; v += blockDim.x) {
    const float wv = Jreg[(size_t)j * SV + v];
    const float x = col < 3 ? vt[v * 3 + col] : shapedirs[(size_t)v * 30 + (col - 3)];
    s += (double)wv * (double)x;
  }
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (col < 3) Jt[j * 3 + col] = (float)red[0];
    else Jd[j * 30 + (col - 3)] = (float)red[0];
  }
}

__global__ void smpl_basis_kernel(const float* __restrict__ vt, const float* __restrict__ shapedirs,
                                  const float* __restrict__ posedirs, float* __restrict__ basis, float* __restrict__ vtp) {
  // basis[k][c][v] ; k<10 shapedirs[v][c][k] ; k>=10 posedirs[k-10][v*3+c] ; zero for v >= SV
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < (size_t)3 * SVP) {
    const int c = (int)(i / SVP), v = (int)(i % SVP);
    vtp[i] = v < SV ? vt[v * 3 + c] : 0.f;
  }
  if (i >= (size_t)SK * 3 * SVP) return;
  const int k = (int)(i / (3 * SVP)), c = (int)((i / SVP) % 3), v = (int)(i % SVP);
  float x = 0.f;
  if (v < SV) x = k < 10 ? shapedirs[(size_t)v * 30 + c * 10 + k] : posedirs[(size_t)(k - 10) * (SV * 3) + v * 3 + c];
  basis[i] = x;
}

// sparse (<=4) and dense-24 skinning layouts; *max_nnz gets the largest per-vertex non-zero count
__global__ void smpl_lbs_pack_kernel(const float* __restrict__ W, float* __restrict__ w4, unsigned char* __restrict__ i4,
                                     float* __restrict__ w24, unsigned char* __restrict__ i24, int* __restrict__ max_nnz) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= SVP) return;
  int n = 0;
  for (int j = 0; j < SJ; ++j) {
    const float x = v < SV ? W[(size_t)v * SJ + j] : 0.f;
    w24[(size_t)v * SJ + j] = x;
    i24[(size_t)v * SJ + j] = (unsigned char)j;
    if (x != 0.f) {
      if (n < 4) { w4[v * 4 + n] = x; i4[v * 4 + n] = (unsigned char)j; }
      ++n;
    }
  }
  for (int k = n; k < 4; ++k) { w4[v * 4 + k] = 0.f; i4[v * 4 + k] = 0; }
  atomicMax(max_nnz, n);
}

}  // namespace seeme

using namespace seeme;

struct seeme_smpl {
  int device = 0, max_frames = 0, max_nnz = 0;
  Arena arena;
  SmplTopo topo;
  float *Jt, *Jd, *basis, *vtp, *w4, *w24;
  unsigned char *i4, *i24;
  float *A, *coef;
  // tensor-core blend path (smpl_tc.cu): packed bf16 (hi, lo) basis and per-call coefficient scratch
  void *bh = nullptr, *bl = nullptr, *ch = nullptr, *cl = nullptr, *wblob = nullptr, *aopblob = nullptr;
  bool use_tc = true;
  int tc_version = 2;      // SEEME_SMPL_TC=1: sparse skinning in the epilogue from shared memory (<= 4 weights per vertex)
};

extern "C" int seeme_smpl_create(seeme_smpl_t* out, const float* v_template, const float* shapedirs,
                                 const float* posedirs, const float* J_regressor, const float* lbs_weights,
                                 const int32_t* parents, int max_frames) {
  SEEME_REQUIRE(out && v_template && shapedirs && posedirs && J_regressor && lbs_weights && parents, SEEME_EINVAL,
                "seeme_smpl_create: null argument");
  SEEME_REQUIRE(max_frames > 0, SEEME_EINVAL, "seeme_smpl_create: max_frames must be positive");
  seeme_smpl* h = new seeme_smpl();
  SEEME_CUDA(cudaGetDevice(&h->device));
  h->max_frames = max_frames;
  h->topo.max_depth = 0;
  for (int j = 0; j < SJ; ++j) {
    const int p = parents[j];
    if (!((j == 0 && p < 0) || (j > 0 && p >= 0 && p < j))) {
      set_error("seeme_smpl_create: parents[%d]=%d is not a topologically ordered tree", j, p);
      delete h;
      return SEEME_EINVAL;
    }
    h->topo.parent[j] = p;
    h->topo.depth[j] = j == 0 ? 0 : h->topo.depth[p] + 1;
    if (h->topo.depth[j] > h->topo.max_depth) h->topo.max_depth = h->topo.depth[j];
  }
  const size_t fpad = ((size_t)max_frames + FT - 1) / FT * FT;
  size_t bytes = pad256(SJ * 3 * 4) + pad256(SJ * 30 * 4) + pad256((size_t)SK * 3 * SVP * 4) + pad256((size_t)3 * SVP * 4) +
                 pad256((size_t)SVP * 4 * 4) + pad256((size_t)SVP * SJ * 4) + pad256((size_t)SVP * 4) + pad256((size_t)SVP * SJ) +
                 pad256(fpad * SJ * 12 * 4) + pad256(fpad * SKP * 4) + 4096 +
                 2 * pad256(smpl_tc_basis_elems() * 2) + 2 * pad256((fpad + 64) * 256 * 2) + pad256(smpl_tc_wblob_bytes()) + pad256(smpl_tc_aop_bytes(fpad));
  int rc = h->arena.init(bytes);
  if (rc) { delete h; return rc; }
  h->Jt = h->arena.take<float>(SJ * 3);
  h->Jd = h->arena.take<float>(SJ * 30);
  h->basis = h->arena.take<float>((size_t)SK * 3 * SVP);
  h->vtp = h->arena.take<float>((size_t)3 * SVP);
  h->w4 = h->arena.take<float>((size_t)SVP * 4);
  h->w24 = h->arena.take<float>((size_t)SVP * SJ);
  h->i4 = h->arena.take<unsigned char>((size_t)SVP * 4);
  h->i24 = h->arena.take<unsigned char>((size_t)SVP * SJ);
  h->A = h->arena.take<float>(fpad * SJ * 12);
  h->coef = h->arena.take<float>(fpad * SKP);
  h->bh = h->arena.take<__nv_bfloat16>(smpl_tc_basis_elems());
  h->bl = h->arena.take<__nv_bfloat16>(smpl_tc_basis_elems());
  h->ch = h->arena.take<__nv_bfloat16>((fpad + 64) * 256);
  h->cl = h->arena.take<__nv_bfloat16>((fpad + 64) * 256);
  h->wblob = h->arena.take<char>(smpl_tc_wblob_bytes());
  h->aopblob = h->arena.take<char>(smpl_tc_aop_bytes(fpad));
  {
    const char* e = getenv("SEEME_SMPL_FP32");
    h->use_tc = !(e && e[0] == '1');
#ifdef SEEME_EXPERIMENTAL
    const char* v = getenv("SEEME_SMPL_TC");
    h->tc_version = (v && v[0] == '1') ? 1 : 2;
#else
    h->tc_version = 2;
#endif
  }
  int* d_nnz = h->arena.take<int>(1);
  if (!d_nnz) { set_error("seeme_smpl_create: arena exhausted"); h->arena.release(); delete h; return SEEME_ENOMEM; }
  smpl_jfold_kernel<<<dim3(SJ, 33), 256>>>(J_regressor, v_template, shapedirs, h->Jt, h->Jd);
  smpl_basis_kernel<<<(unsigned)(((size_t)SK * 3 * SVP + 255) / 256), 256>>>(v_template, shapedirs, posedirs, h->basis, h->vtp);
  cudaMemset(d_nnz, 0, sizeof(int));
  if (h->cl) {
    cudaMemset(h->ch, 0, (fpad + 64) * 256 * 2);
    cudaMemset(h->cl, 0, (fpad + 64) * 256 * 2);
    if ((h->tc_version == 2 ? smpl_tc_pack_basis_f16(h->basis, SK, h->bh, h->bl) : smpl_tc_pack_basis(h->basis, SK, h->bh, h->bl)) != SEEME_OK) { h->arena.release(); delete h; return SEEME_ECUDA; }
  }
  smpl_lbs_pack_kernel<<<(SVP + 127) / 128, 128>>>(lbs_weights, h->w4, h->i4, h->w24, h->i24, d_nnz);
  if (h->wblob && smpl_tc_pack_wtiles(h->w24, h->wblob) != SEEME_OK) { h->arena.release(); delete h; return SEEME_ECUDA; }
  cudaError_t e = cudaMemcpy(&h->max_nnz, d_nnz, sizeof(int), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { set_error("seeme_smpl_create: packing failed: %s", cudaGetErrorString(e)); h->arena.release(); delete h; return SEEME_ECUDA; }
  *out = h;
  return SEEME_OK;
}

static int smpl_run(seeme_smpl* h, const float* betas, const float* body_pose, const float* global_orient,
                    const float* transl, const float* feats, int Dn, const double* mean, const double* stdv, int n_body,
                    double* m_out, int F, float* vertices, float* joints, float* quat, cudaStream_t s) {
  SEEME_REQUIRE(F > 0, SEEME_EINVAL, "seeme_smpl: empty batch");
  SEEME_REQUIRE(F <= h->max_frames, SEEME_ECAP, "seeme_smpl: %d frames exceed capacity %d", F, h->max_frames);
  {
    ProfScope prof(PROF_SMPL_POSE, s);
    smpl_pose_kernel<<<(F + 7) / 8, 256, 0, s>>>(betas, body_pose, global_orient, transl, feats, Dn, mean, stdv, n_body, m_out,
                                                 h->Jt, h->Jd, h->topo, F, h->A, h->coef, joints, quat);
  }
  SEEME_LAUNCH_CHECK();
  if (vertices && h->use_tc && h->tc_version == 2 && h->wblob && h->aopblob) {
    // blend shapes AND the skinning-transform blend on the tensor cores (smpl_tc.cu, version 2)
    SEEME_TRY(smpl_skin_tc2(h->bh, h->bl, h->coef, SKP, SKP, h->ch, h->cl, h->A, h->vtp, h->wblob, h->aopblob, F, vertices, PROF_SMPL_SKIN + 1, s));
  } else if (vertices && h->use_tc && h->max_nnz <= 4) {
    // blend contraction on the tensor cores, skinning in its epilogue (smpl_tc.cu)
    SEEME_TRY(smpl_skin_tc(h->bh, h->bl, h->coef, SKP, SKP, h->ch, h->cl, h->A, h->vtp, h->w4, h->i4, F, vertices,
                           PROF_SMPL_SKIN + 1, s));
  } else if (vertices) {
    ProfScope prof(PROF_SMPL_SKIN, s);
    dim3 grid(SVP / 128, (F + FT - 1) / FT);
    if (h->max_nnz <= 4)
      smpl_skin_kernel<4><<<grid, 128, 0, s>>>(h->basis, h->vtp, h->w4, h->i4, h->A, h->coef, F, vertices);
    else
      smpl_skin_kernel<SJ><<<grid, 128, 0, s>>>(h->basis, h->vtp, h->w24, h->i24, h->A, h->coef, F, vertices);
    SEEME_LAUNCH_CHECK();
  }
  return SEEME_OK;
}

extern "C" int seeme_smpl_forward(seeme_smpl_t h, const float* betas, const float* body_pose, const float* global_orient,
                                  const float* transl, int F, float* vertices, float* joints, float* quat, void* stream) {
  SEEME_REQUIRE(h && betas && body_pose && global_orient && joints, SEEME_EINVAL, "seeme_smpl_forward: null argument");
  return smpl_run(h, betas, body_pose, global_orient, transl, nullptr, 0, nullptr, nullptr, 69, nullptr, F, vertices, joints,
                  quat, (cudaStream_t)stream);
}

extern "C" int seeme_smpl_forward_feats(seeme_smpl_t h, const float* feats, int Dn, const double* mean, const double* std,
                                        int n_body, const float* betas, int F, double* m_out, float* vertices, float* joints,
                                        float* quat, void* stream) {
  SEEME_REQUIRE(h && feats && mean && std && betas && joints, SEEME_EINVAL, "seeme_smpl_forward_feats: null argument");
  SEEME_REQUIRE(n_body > 0 && n_body <= 69 && n_body % 3 == 0 && Dn >= 3 + n_body + 3, SEEME_EINVAL,
                "seeme_smpl_forward_feats: bad layout (Dn=%d, n_body=%d)", Dn, n_body);
  return smpl_run(h, betas, nullptr, nullptr, nullptr, feats, Dn, mean, std, n_body, m_out, F, vertices, joints, quat,
                  (cudaStream_t)stream);
}

extern "C" int seeme_smpl_destroy(seeme_smpl_t h) {
  if (!h) return SEEME_OK;
  h->arena.release();
  delete h;
  return SEEME_OK;
}
