// Scene encoder: ResnetPointnet (EgoHMR/models/respointnet.py:33-59,88-97) + output_scene
// (mld/models/modeltype/mld.py:257-261).
//
// Dataflow (per chunk of samples, all per-point tensors [rows = C*N, channels]):
//   x0  = fc_pos_0(p)                                     [rows,512]
//   blk0: h = fc_0(relu(x0)); net = shortcut(x0) + b1 + fc_1(relu(h))
//   blk i>=1: the reference concatenates the per-sample max-pool to every point
//     ([net, pooled], respointnet.py:37-46).  The pooled half of each contraction is constant over the
//     points of a sample, so it is folded into per-sample bias vectors:
//       c0[b] = W0[:,256:] relu(pool_b) + b0 ;  cs[b] = Ws[:,256:] pool_b + b1
//       h = W0[:,:256] relu(net) + c0[b] ;  net' = Ws[:,:256] net + cs[b] + fc_1(relu(h))
//   pooling: per-(sample,channel) max via order-preserving atomicMax.
#include "common.cuh"

namespace seeme {

__global__ void colmax_kernel(const float* __restrict__ x, unsigned* __restrict__ out, int N, int rows_per_cta) {
  // x [B,N,256]; grid (splits, B); 256 threads: thread = channel
  const int b = blockIdx.y, c = threadIdx.x;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(N, r0 + rows_per_cta);
  const float* p = x + ((size_t)b * N + r0) * 256 + c;
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    m0 = fmaxf(m0, p[0]); m1 = fmaxf(m1, p[256]); m2 = fmaxf(m2, p[512]); m3 = fmaxf(m3, p[768]);
    p += 1024;
  }
  for (; r < r1; ++r) { m0 = fmaxf(m0, p[0]); p += 256; }
  atomicMax(out + (size_t)b * 256 + c, f2ord(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3))));
}

__global__ void ord_decode_kernel(const unsigned* __restrict__ in, float* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = ord2f(in[i]);
}

}  // namespace seeme

using namespace seeme;

struct seeme_pointnet {
  int device = 0, max_batch = 0, max_points = 0, chunk = 0;
  Arena arena;
  // packed copies of the weights (contiguous fp32)
  float *fc_pos_w, *fc_pos_b;
  float *w0[4], *b0[4], *w1[4], *b1[4], *ws[4];
  float *wc, *bc, *wo, *bo;
  // workspace
  float *x0, *h, *net[2], *pool, *c0, *cs, *feat;
  unsigned* pool_ord;
};

static int copy_w(Arena& a, float*& dst, const float* src, size_t n) {
  dst = a.take<float>(n);
  SEEME_REQUIRE(dst != nullptr, SEEME_ENOMEM, "pointnet: arena exhausted");
  SEEME_CUDA(cudaMemcpy(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice));
  return SEEME_OK;
}

extern "C" int seeme_pointnet_create(seeme_pointnet_t* out, const float* const* w, int n_w, int max_batch,
                                     int max_points) {
  SEEME_REQUIRE(out && w, SEEME_EINVAL, "seeme_pointnet_create: null argument");
  SEEME_REQUIRE(n_w == SEEME_POINTNET_NUM_TENSORS, SEEME_EINVAL, "seeme_pointnet_create: expected %d tensors, got %d",
                SEEME_POINTNET_NUM_TENSORS, n_w);
  SEEME_REQUIRE(max_batch > 0 && max_points > 0, SEEME_EINVAL, "seeme_pointnet_create: bad capacity");
  for (int i = 0; i < n_w; ++i) SEEME_REQUIRE(w[i] != nullptr, SEEME_EINVAL, "seeme_pointnet_create: tensor %d is null", i);
  seeme_pointnet* h = new seeme_pointnet();
  SEEME_CUDA(cudaGetDevice(&h->device));
  h->max_batch = max_batch;
  h->max_points = max_points;
  // samples per pass: bound the per-point workspace to ~3.3 GB
  h->chunk = max_batch < 32 ? max_batch : 32;
  const size_t rows = (size_t)h->chunk * max_points;
  size_t wbytes = pad256(512 * 3 * 4) + pad256(512 * 4) + 4 * (pad256(256 * 512 * 4) * 2 + pad256(256 * 256 * 4) + 2 * pad256(256 * 4)) +
                  pad256(512 * 256 * 4) + pad256(512 * 4) + pad256(256 * 512 * 4) + pad256(256 * 4);
  size_t ws = pad256(rows * 512 * 4) + 3 * pad256(rows * 256 * 4) + 4 * pad256((size_t)max_batch * 256 * 4) +
              pad256((size_t)max_batch * 512 * 4);
  int rc = h->arena.init(wbytes + ws + 4096);
  if (rc != SEEME_OK) { delete h; return rc; }
  int k = 0;
  rc = copy_w(h->arena, h->fc_pos_w, w[k++], 512 * 3);
  if (!rc) rc = copy_w(h->arena, h->fc_pos_b, w[k++], 512);
  for (int i = 0; i < 4 && !rc; ++i) {
    rc = copy_w(h->arena, h->w0[i], w[k++], 256 * 512);
    if (!rc) rc = copy_w(h->arena, h->b0[i], w[k++], 256);
    if (!rc) rc = copy_w(h->arena, h->w1[i], w[k++], 256 * 256);
    if (!rc) rc = copy_w(h->arena, h->b1[i], w[k++], 256);
    if (!rc) rc = copy_w(h->arena, h->ws[i], w[k++], 256 * 512);
  }
  if (!rc) rc = copy_w(h->arena, h->wc, w[k++], 512 * 256);
  if (!rc) rc = copy_w(h->arena, h->bc, w[k++], 512);
  if (!rc) rc = copy_w(h->arena, h->wo, w[k++], 256 * 512);
  if (!rc) rc = copy_w(h->arena, h->bo, w[k++], 256);
  if (rc) { h->arena.release(); delete h; return rc; }
  h->x0 = h->arena.take<float>(rows * 512);
  h->h = h->arena.take<float>(rows * 256);
  h->net[0] = h->arena.take<float>(rows * 256);
  h->net[1] = h->arena.take<float>(rows * 256);
  h->pool = h->arena.take<float>((size_t)max_batch * 256);
  h->c0 = h->arena.take<float>((size_t)max_batch * 256);
  h->cs = h->arena.take<float>((size_t)max_batch * 256);
  h->pool_ord = h->arena.take<unsigned>((size_t)max_batch * 256);
  h->feat = h->arena.take<float>((size_t)max_batch * 512);
  if (!h->feat) { set_error("pointnet: arena exhausted (workspace)"); h->arena.release(); delete h; return SEEME_ENOMEM; }
  *out = h;
  return SEEME_OK;
}

static int pool_stage(seeme_pointnet* h, const float* net, int C, int N, cudaStream_t s) {
  SEEME_CUDA(cudaMemsetAsync(h->pool_ord, 0, (size_t)C * 256 * sizeof(unsigned), s));
  const int rows_per_cta = 256;
  dim3 grid((N + rows_per_cta - 1) / rows_per_cta, C);
  colmax_kernel<<<grid, 256, 0, s>>>(net, h->pool_ord, N, rows_per_cta);
  SEEME_LAUNCH_CHECK();
  ord_decode_kernel<<<(C * 256 + 255) / 256, 256, 0, s>>>(h->pool_ord, h->pool, C * 256);
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

extern "C" int seeme_pointnet_forward(seeme_pointnet_t h, const float* pcd, int B, int N, float* feat512,
                                      float* emb256, void* stream) {
  SEEME_REQUIRE(h && pcd, SEEME_EINVAL, "seeme_pointnet_forward: null argument");
  SEEME_REQUIRE(B > 0 && N > 0, SEEME_EINVAL, "seeme_pointnet_forward: empty input (B=%d, N=%d)", B, N);
  SEEME_REQUIRE(B <= h->max_batch && N <= h->max_points, SEEME_ECAP,
                "seeme_pointnet_forward: B=%d N=%d exceeds capacity (%d, %d)", B, N, h->max_batch, h->max_points);
  cudaStream_t s = (cudaStream_t)stream;
  for (int b0 = 0; b0 < B; b0 += h->chunk) {
    const int C = (B - b0 < h->chunk) ? (B - b0) : h->chunk;
    const int rows = C * N;
    const float* p = pcd + (size_t)b0 * N * 3;
    // fc_pos_0
    SEEME_TRY(gemm_f32(gemm_params(p, 3, h->fc_pos_w, 3, h->fc_pos_b, h->x0, 512, rows, 512, 3), s));
    float* net = h->net[0];
    float* nxt = h->net[1];
    {  // block_0 (K = 512, no pooled half)
      GemmP g = gemm_params(h->x0, 512, h->w0[0], 512, h->b0[0], h->h, 256, rows, 256, 512);
      g.pre_act = ACT_RELU; g.prof_id = PROF_POINTNET_GEMM + 1;
      SEEME_TRY(gemm_f32(g, s));
      GemmP gs0 = gemm_params(h->x0, 512, h->ws[0], 512, h->b1[0], net, 256, rows, 256, 512);
      gs0.prof_id = PROF_POINTNET_GEMM + 1;
      SEEME_TRY(gemm_f32(gs0, s));
      GemmP g1 = gemm_params(h->h, 256, h->w1[0], 256, nullptr, net, 256, rows, 256, 256);
      g1.pre_act = ACT_RELU; g1.accumulate = 1; g1.prof_id = PROF_POINTNET_GEMM + 1;
      SEEME_TRY(gemm_f32(g1, s));
    }
    for (int i = 1; i < 4; ++i) {
      SEEME_TRY(pool_stage(h, net, C, N, s));
      GemmP gc0 = gemm_params(h->pool, 256, h->w0[i] + 256, 512, h->b0[i], h->c0, 256, C, 256, 256);
      gc0.pre_act = ACT_RELU;
      SEEME_TRY(gemm_f32(gc0, s));
      SEEME_TRY(gemm_f32(gemm_params(h->pool, 256, h->ws[i] + 256, 512, h->b1[i], h->cs, 256, C, 256, 256), s));
      GemmP g0 = gemm_params(net, 256, h->w0[i], 512, h->c0, h->h, 256, rows, 256, 256);
      g0.pre_act = ACT_RELU; g0.bias_group_rows = N; g0.prof_id = PROF_POINTNET_GEMM + 1;
      SEEME_TRY(gemm_f32(g0, s));
      GemmP gs = gemm_params(net, 256, h->ws[i], 512, h->cs, nxt, 256, rows, 256, 256);
      gs.bias_group_rows = N; gs.prof_id = PROF_POINTNET_GEMM + 1;
      SEEME_TRY(gemm_f32(gs, s));
      GemmP g1 = gemm_params(h->h, 256, h->w1[i], 256, nullptr, nxt, 256, rows, 256, 256);
      g1.pre_act = ACT_RELU; g1.accumulate = 1; g1.prof_id = PROF_POINTNET_GEMM + 1;
      SEEME_TRY(gemm_f32(g1, s));
      float* t = net; net = nxt; nxt = t;
    }
    SEEME_TRY(pool_stage(h, net, C, N, s));
    // fc_c(relu(pool)) -> [C,512]; output_scene: Linear(relu(.)) -> [C,256]
    GemmP gc = gemm_params(h->pool, 256, h->wc, 256, h->bc, h->feat + (size_t)b0 * 512, 512, C, 512, 256);
    gc.pre_act = ACT_RELU;
    SEEME_TRY(gemm_f32(gc, s));
    if (feat512)
      SEEME_CUDA(cudaMemcpyAsync(feat512 + (size_t)b0 * 512, h->feat + (size_t)b0 * 512, (size_t)C * 512 * 4,
                                 cudaMemcpyDeviceToDevice, s));
    if (emb256) {
      GemmP go = gemm_params(h->feat + (size_t)b0 * 512, 512, h->wo, 512, h->bo, emb256 + (size_t)b0 * 256, 256, C, 256, 512);
      go.pre_act = ACT_RELU;
      SEEME_TRY(gemm_f32(go, s));
    }
  }
  return SEEME_OK;
}

extern "C" int seeme_pointnet_destroy(seeme_pointnet_t h) {
  if (!h) return SEEME_OK;
  h->arena.release();
  delete h;
  return SEEME_OK;
}
