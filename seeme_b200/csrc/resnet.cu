// Image backbone of the scene/image-conditioned variants: ProHMRScene.encode_image =
// ResNet-50 (EgoHMR/models/prohmr/prohmr_scene.py:99-100, EgoHMR/models/resnet.py:60-180; Bottleneck [3,4,6,3],
// stride on the 3x3 convolution, 1x1 stride-s downsample shortcut, eval-mode BatchNorm, mean over the 7x7 map).
//
// Layout: activations are NHWC so that every convolution is ONE row-major GEMM  out[M = B.Ho.Wo, Cout] =
// A[M, K] . W[Cout, K]^T  on the tcgen05 linear of umma_gemm.cu (split-bf16 x3 operands, fp32 accumulation):
//   1x1 stride 1    A is the activation itself (no copy);
//   1x1 stride 2    A = the even-pixel rows gathered by rn_gather_kernel (taps = 1);
//   3x3 / 7x7       A = the patch matrix written by rn_gather_kernel / rn_stem_patch_kernel, K = (ky, kx, c).
// BatchNorm is folded into the packed weights at create time (W' = W . g / sqrt(var + eps), b' = beta - mean . g /
// sqrt(var + eps)), so conv + BN + ReLU (+ residual) is the GEMM epilogue: bias, [residual], relu, and the result is
// written as the fp32 copy (next residual) and/or the bf16 (hi, lo) copy (next A operand) by the same kernel.
#include "common.cuh"
#include "umma.cuh"
#include <stdlib.h>
#include <vector>

namespace seeme {

// 16-bit copies of an fp32 value: bf16 (hi, lo) pair, or one IEEE fp16 stored in the bf16-typed `hi` buffer (f16 != 0)
__device__ __forceinline__ void rn_store16(float v, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t i, int f16) {
  if (f16) {
    reinterpret_cast<__half*>(hi)[i] = __float2half_rn(v);
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

__global__ void rn_to_f16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) reinterpret_cast<__half*>(out)[i] = __float2half_rn(x[i]);
}

// W'[n, (ky,kx,c)] = W[n,c,ky,kx] * s[n] (zero-padded to Kp columns), b'[n] = beta[n] - mean[n] * s[n]
__global__ void rn_fold_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float* __restrict__ wf,
                               float* __restrict__ bf, int Cout, int Cin, int kh, int kw, int Kp) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)Cout * Kp) return;
  const int n = (int)(i / Kp), k = (int)(i % Kp);
  const float s = gamma[n] / sqrtf(var[n] + 1e-5f);
  float v = 0.f;
  if (k < kh * kw * Cin) {
    const int tap = k / Cin, c = k % Cin;
    v = w[((size_t)n * Cin + c) * (kh * kw) + tap] * s;
  }
  wf[i] = v;
  if (k == 0) bf[n] = beta[n] - mean[n] * s;
}

// stem patches: x [B,3,224,224] fp32 (NCHW, the reference's input) -> A [B*112*112, 192] bf16 (hi, lo),
// k = (ky*7 + kx)*3 + c for the 7x7 stride-2 pad-3 convolution, columns 147..191 zero
__global__ void rn_stem_patch_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                     int B, int f16) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;     // one thread per 8 consecutive k (16-byte stores)
  if (i >= (size_t)B * 12544 * 24) return;
  const int k0 = (int)(i % 24) * 8;
  const size_t pix = i / 24;
  const int ox = (int)(pix % 112), oy = (int)((pix / 112) % 112), b = (int)(pix / 12544);
  __align__(16) unsigned short vh[8], vl[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = k0 + j;
    float v = 0.f;
    if (k < 147) {
      const int c = k % 3, kx = (k / 3) % 7, ky = k / 21;
      const int iy = oy * 2 - 3 + ky, ix = ox * 2 - 3 + kx;
      if (iy >= 0 && iy < 224 && ix >= 0 && ix < 224) v = __ldg(x + (((size_t)b * 3 + c) * 224 + iy) * 224 + ix);
    }
    if (f16) {
      vh[j] = __half_as_ushort(__float2half_rn(v));
      vl[j] = 0;
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      vh[j] = __bfloat16_as_ushort(h);
      vl[j] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
    }
  }
  *reinterpret_cast<uint4*>(hi + pix * 192 + k0) = *reinterpret_cast<const uint4*>(vh);
  if (!f16) *reinterpret_cast<uint4*>(lo + pix * 192 + k0) = *reinterpret_cast<const uint4*>(vl);
}

// 3x3 stride-2 pad-1 max pooling of the stem output y [B,112,112,64] fp32 -> [B,56,56,64] fp32 + bf16 (hi, lo)
__global__ void rn_maxpool_kernel(const float* __restrict__ y, float* __restrict__ o, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int B, int f16) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;     // one thread per 4 channels
  if (i >= (size_t)B * 3136 * 16) return;
  const int c4 = (int)(i % 16);
  const size_t pix = i / 16;
  const int ox = (int)(pix % 56), oy = (int)((pix / 56) % 56), b = (int)(pix / 3136);
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
    if (iy < 0 || iy >= 112) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - 1 + kx;
      if (ix < 0 || ix >= 112) continue;
      const float4 v = *reinterpret_cast<const float4*>(y + (((size_t)b * 112 + iy) * 112 + ix) * 64 + c4 * 4);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  *reinterpret_cast<float4*>(o + pix * 64 + c4 * 4) = m;
  const float f[4] = {m.x, m.y, m.z, m.w};
  for (int j = 0; j < 4; ++j) rn_store16(f[j], hi, lo, pix * 64 + c4 * 4 + j, f16);
}

// patch / subsample gather on the bf16 (hi, lo) NHWC activation: out[(b,oy,ox), tap*C + c] = in[b, oy*s - pad + ky,
// ox*s - pad + kx, c] (0 outside); one thread moves 8 channels (16 B) of both halves
__global__ void rn_gather_kernel(const __nv_bfloat16* __restrict__ ih, const __nv_bfloat16* __restrict__ il,
                                 __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol, int B, int H, int W, int C,
                                 int ksz, int stride, int pad, int Ho, int Wo) {
  const int c8n = C / 8;
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * Ho * Wo * ksz * ksz * c8n;
  if (i >= total) return;
  const int c8 = (int)(i % c8n);
  size_t r = i / c8n;
  const int tap = (int)(r % (ksz * ksz));
  r /= (ksz * ksz);
  const int ox = (int)(r % Wo), oy = (int)((r / Wo) % Ho), b = (int)(r / ((size_t)Wo * Ho));
  const int iy = oy * stride - pad + tap / ksz, ix = ox * stride - pad + tap % ksz;
  uint4 vh = make_uint4(0, 0, 0, 0), vl = vh;
  if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
    const size_t src = (((size_t)b * H + iy) * W + ix) * C + c8 * 8;
    vh = *reinterpret_cast<const uint4*>(ih + src);
    if (il) vl = *reinterpret_cast<const uint4*>(il + src);
  }
  const size_t dst = r * ((size_t)ksz * ksz * C) + (size_t)tap * C + c8 * 8;
  *reinterpret_cast<uint4*>(oh + dst) = vh;
  if (il) *reinterpret_cast<uint4*>(ol + dst) = vl;
}

// x4.mean(dim=(2,3)) (resnet.py:180): [B,49,2048] fp32 -> [B,2048]
__global__ void rn_avgpool_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int HW, int C) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)B * C) return;
  const int c = (int)(i % C), b = (int)(i / C);
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += x[((size_t)b * HW + p) * C + c];
  out[i] = s / (float)HW;
}

}  // namespace seeme

using namespace seeme;

namespace {
constexpr int RN_LAYERS[4] = {3, 4, 6, 3};
struct RnConv {
  PackedLinear w;         // folded weights [Cout, Kp] (hi, lo) + folded bias
  int Cin = 0, Cout = 0, ksz = 1, Kp = 0;
};
struct RnBlock {
  RnConv c1, c2, c3, ds;
  bool has_ds = false;
  int stride = 1, planes = 0, cin = 0;
};
}  // namespace

struct seeme_resnet50 {
  int device = 0, max_batch = 0, chunk = 0;
  int f16 = 0;            // 1: IEEE fp16 operands, one MMA pass (SEEME_RESNET_PRECISION=16); 0: split-bf16 x3
  Arena arena;
  RnConv stem;
  std::vector<RnBlock> blocks;
  // workspace of one chunk of images
  __nv_bfloat16 *patch_h = nullptr, *patch_l = nullptr;     // patch matrices (largest: stem, 12544 x 192 per image)
  float* stem_out = nullptr;                                // [chunk,112,112,64]
  ActBuf x[2];                                              // block input / output (fp32 + hi/lo), <= 3136 x 256 per image
  ActBuf t1, t2;                                            // bottleneck intermediates (hi/lo)
  float* res = nullptr;                                     // downsample shortcut (fp32)
  // output_images tail (mld.py:251-255): relu -> Linear(2048, 256)
  PackedLinear tail;
  float* feat = nullptr;                                    // [max_batch, 2048]
  __nv_bfloat16 *feat_h = nullptr, *feat_l = nullptr;       // relu(feat) as bf16 (hi, lo)
};

static void rn_destroy(seeme_resnet50* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  h->arena.release();
  delete h;
}

// fold conv + BN into a packed GEMM weight; `t` points at {conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var}
static int rn_pack(seeme_resnet50* h, RnConv& c, const float* const* t, int Cin, int Cout, int ksz, float* scratch) {
  c.Cin = Cin; c.Cout = Cout; c.ksz = ksz;
  const int K = ksz * ksz * Cin;
  c.Kp = (K + 63) / 64 * 64;
  float* bias = h->arena.take<float>(Cout);
  SEEME_REQUIRE(bias, SEEME_ENOMEM, "seeme_resnet50_create: arena exhausted");
  const size_t n = (size_t)Cout * c.Kp;
  rn_fold_kernel<<<(unsigned)((n + 255) / 256), 256>>>(t[0], t[1], t[2], t[3], t[4], scratch, bias, Cout, Cin, ksz, ksz, c.Kp);
  SEEME_LAUNCH_CHECK();
  SEEME_TRY(pack_linear(h->arena, c.w, scratch, c.Kp, Cout, c.Kp, bias));
  if (h->f16) {
    rn_to_f16_kernel<<<(unsigned)((n + 255) / 256), 256>>>(scratch, c.w.hi, n);
    SEEME_LAUNCH_CHECK();
  }
  return SEEME_OK;
}

extern "C" int seeme_resnet50_create(seeme_resnet50_t* out, const float* const* w, int n_w, int max_batch) {
  SEEME_REQUIRE(out && w, SEEME_EINVAL, "seeme_resnet50_create: null argument");
  SEEME_REQUIRE(n_w == SEEME_RESNET50_NUM_TENSORS, SEEME_EINVAL, "seeme_resnet50_create: expected %d tensors, got %d",
                SEEME_RESNET50_NUM_TENSORS, n_w);
  SEEME_REQUIRE(max_batch > 0, SEEME_EINVAL, "seeme_resnet50_create: max_batch must be positive");
  for (int i = 0; i < n_w; ++i) SEEME_REQUIRE(w[i] != nullptr, SEEME_EINVAL, "seeme_resnet50_create: tensor %d is null", i);
  seeme_resnet50* h = new seeme_resnet50();
  SEEME_CUDA(cudaGetDevice(&h->device));
  h->max_batch = max_batch;
  const char* pe = getenv("SEEME_RESNET_PRECISION");
  const int prec = pe ? atoi(pe) : 3;
  if (prec != 3 && prec != 16) { set_error("SEEME_RESNET_PRECISION must be 3 (split-bf16) or 16 (fp16), got %d", prec); delete h; return SEEME_EINVAL; }
  h->f16 = prec == 16;
  const char* ce = seeme_exp_env("SEEME_RESNET_CHUNK");
  const int cap = ce ? atoi(ce) : 128;
  h->chunk = max_batch < cap ? max_batch : (cap > 0 ? cap : 128);
  const size_t C = (size_t)h->chunk;
  // packed weights: 23.5 M folded parameters as bf16 (hi, lo) + biases + the fp32 fold scratch (largest conv: 512 x 4608)
  const size_t wbytes = (size_t)26 * 1000 * 1000 * 4 + 4096 + pad256((size_t)2048 * 1024 * 4 > (size_t)512 * 4608 * 4 ? (size_t)2048 * 1024 * 4
                                                                                                             : (size_t)512 * 4608 * 4) +
                        (size_t)60 * 4 * 256 * 16;
  const size_t rows1 = C * 3136;
  const size_t tailb = 2 * pad256((size_t)256 * 2048 * 2) + pad256((size_t)max_batch * 2048 * 4) + 2 * pad256((size_t)max_batch * 2048 * 2);
  const size_t ws = tailb + 2 * pad256(C * 12544 * 192 * 2) + pad256(C * 12544 * 64 * 4) + 2 * (pad256(rows1 * 256 * 4) + 2 * pad256(rows1 * 256 * 2)) +
                    2 * 2 * pad256(rows1 * 128 * 2) + pad256(rows1 * 256 * 4);
  int rc = h->arena.init(wbytes + ws + 65536);
  if (rc != SEEME_OK) { delete h; return rc; }
  float* scratch = h->arena.take<float>((size_t)2048 * 1024 > (size_t)512 * 4608 ? (size_t)2048 * 1024 : (size_t)512 * 4608);
  int k = 0;
  rc = rn_pack(h, h->stem, w + k, 3, 64, 7, scratch);
  k += 5;
  int cin = 64;
  for (int L = 0; L < 4 && !rc; ++L) {
    const int planes = 64 << L;
    for (int b = 0; b < RN_LAYERS[L] && !rc; ++b) {
      RnBlock blk;
      blk.planes = planes; blk.cin = cin;
      blk.stride = (b == 0 && L > 0) ? 2 : 1;
      blk.has_ds = b == 0;
      rc = rn_pack(h, blk.c1, w + k, cin, planes, 1, scratch); k += 5;
      if (!rc) { rc = rn_pack(h, blk.c2, w + k, planes, planes, 3, scratch); k += 5; }
      if (!rc) { rc = rn_pack(h, blk.c3, w + k, planes, planes * 4, 1, scratch); k += 5; }
      if (!rc && blk.has_ds) { rc = rn_pack(h, blk.ds, w + k, cin, planes * 4, 1, scratch); k += 5; }
      cin = planes * 4;
      h->blocks.push_back(blk);
    }
  }
  if (!rc) rc = pack_linear(h->arena, h->tail, w[k], 2048, 256, 2048, w[k + 1]);
  if (!rc) {   // the handle keeps its own copy of the tail bias (the caller's tensors may be freed after create)
    float* tb = h->arena.take<float>(256);
    if (!tb) { set_error("seeme_resnet50_create: arena exhausted"); rc = SEEME_ENOMEM; }
    else if (cudaMemcpy(tb, w[k + 1], 256 * 4, cudaMemcpyDeviceToDevice) != cudaSuccess) { set_error("seeme_resnet50_create: bias copy failed"); rc = SEEME_ECUDA; }
    else h->tail.bias = tb;
  }
  if (!rc) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { set_error("seeme_resnet50_create: weight packing failed: %s", cudaGetErrorString(e)); rc = SEEME_ECUDA; }
  }
  if (rc) { rn_destroy(h); return rc; }
  h->patch_h = h->arena.take<__nv_bfloat16>(C * 12544 * 192);
  h->patch_l = h->arena.take<__nv_bfloat16>(C * 12544 * 192);
  h->stem_out = h->arena.take<float>(C * 12544 * 64);
  for (int i = 0; i < 2; ++i) {
    h->x[i].f = h->arena.take<float>(rows1 * 256);
    h->x[i].h = h->arena.take<__nv_bfloat16>(rows1 * 256);
    h->x[i].l = h->arena.take<__nv_bfloat16>(rows1 * 256);
  }
  h->t1.h = h->arena.take<__nv_bfloat16>(rows1 * 128);
  h->t1.l = h->arena.take<__nv_bfloat16>(rows1 * 128);
  h->t2.h = h->arena.take<__nv_bfloat16>(rows1 * 128);
  h->t2.l = h->arena.take<__nv_bfloat16>(rows1 * 128);
  h->res = h->arena.take<float>(rows1 * 256);
  h->feat = h->arena.take<float>((size_t)max_batch * 2048);
  h->feat_h = h->arena.take<__nv_bfloat16>((size_t)max_batch * 2048);
  h->feat_l = h->arena.take<__nv_bfloat16>((size_t)max_batch * 2048);
  if (!h->res || !h->feat_l) { set_error("seeme_resnet50_create: arena exhausted (workspace)"); rn_destroy(h); return SEEME_ENOMEM; }
  *out = h;
  return SEEME_OK;
}

extern "C" int seeme_resnet50_destroy(seeme_resnet50_t h) {
  rn_destroy(h);
  return SEEME_OK;
}

// out = [relu](A . W'^T + b' [+ R]) with the requested copies
static int rn_gemm(const seeme_resnet50* h, const RnConv& c, const __nv_bfloat16* ah, const __nv_bfloat16* al, int lda, int M, bool relu,
                   const float* R, float* yf, __nv_bfloat16* yh, __nv_bfloat16* yl, cudaStream_t s) {
  UmmaLinear g;
  const bool f16 = h->f16 != 0;
  g.fp16 = f16;
  if (f16) { al = nullptr; yl = nullptr; }
  g.A1 = {const_cast<__nv_bfloat16*>(ah), const_cast<__nv_bfloat16*>(al), lda};
  g.W = {c.w.hi, f16 ? nullptr : c.w.lo, c.Kp};
  g.M = M; g.N = c.Cout; g.K1 = c.Kp; g.K2 = 0;
  g.bias = c.w.bias;
  g.act = relu ? ACT_RELU : ACT_NONE;
  g.R = R; g.ldr = c.Cout; g.act_after_residual = R != nullptr;
  static const bool dense = !(seeme_exp_env("SEEME_RESNET_DENSE") && seeme_exp_env("SEEME_RESNET_DENSE")[0] == '0');
  g.dense_ctas = dense && M > 2048;
  g.Y = yf; g.ldy = c.Cout;
  g.Yh = yh; g.Yl = yl; g.ldb = c.Cout;
  return umma_linear(g, f16 ? 1 : 3, s);
}

static int rn_gather(const seeme_resnet50* h, const ActBuf& in, __nv_bfloat16* oh, __nv_bfloat16* ol, int B, int H, int W, int C, int ksz,
                     int stride, int pad, int Ho, int Wo, cudaStream_t s) {
  const size_t total = (size_t)B * Ho * Wo * ksz * ksz * (C / 8);
  rn_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in.h, h->f16 ? nullptr : in.l, oh, ol, B, H, W, C, ksz, stride, pad, Ho,
                                                                      Wo);
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

static int rn_chunk(seeme_resnet50* h, const float* img, int B, float* out, cudaStream_t s) {
  {  // stem: conv 7x7/2 + BN + ReLU, max-pool 3x3/2
    const size_t n = (size_t)B * 12544 * 24;
    rn_stem_patch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(img, h->patch_h, h->patch_l, B, h->f16);
    SEEME_LAUNCH_CHECK();
    SEEME_TRY(rn_gemm(h, h->stem, h->patch_h, h->patch_l, 192, B * 12544, true, nullptr, h->stem_out, nullptr, nullptr, s));
    const size_t m = (size_t)B * 3136 * 16;
    rn_maxpool_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(h->stem_out, h->x[0].f, h->x[0].h, h->x[0].l, B, h->f16);
    SEEME_LAUNCH_CHECK();
  }
  int cur = 0, H = 56;
  for (const RnBlock& b : h->blocks) {
    const ActBuf& xi = h->x[cur];
    const ActBuf& xo = h->x[cur ^ 1];
    const int Ho = H / b.stride, Min = B * H * H, Mout = B * Ho * Ho, P = b.planes;
    // conv1 1x1 + BN + ReLU
    SEEME_TRY(rn_gemm(h, b.c1, xi.h, xi.l, b.cin, Min, true, nullptr, nullptr, h->t1.h, h->t1.l, s));
    // conv2 3x3 (stride) + BN + ReLU on the patch matrix
    SEEME_TRY(rn_gather(h, h->t1, h->patch_h, h->patch_l, B, H, H, P, 3, b.stride, 1, Ho, Ho, s));
    SEEME_TRY(rn_gemm(h, b.c2, h->patch_h, h->patch_l, 9 * P, Mout, true, nullptr, nullptr, h->t2.h, h->t2.l, s));
    // shortcut
    const float* R = xi.f;
    if (b.has_ds) {
      if (b.stride == 1) {
        SEEME_TRY(rn_gemm(h, b.ds, xi.h, xi.l, b.cin, Min, false, nullptr, h->res, nullptr, nullptr, s));
      } else {
        SEEME_TRY(rn_gather(h, xi, h->patch_h, h->patch_l, B, H, H, b.cin, 1, b.stride, 0, Ho, Ho, s));
        SEEME_TRY(rn_gemm(h, b.ds, h->patch_h, h->patch_l, b.cin, Mout, false, nullptr, h->res, nullptr, nullptr, s));
      }
      R = h->res;
    }
    // conv3 1x1 + BN, + shortcut, ReLU: fp32 copy (next shortcut) and bf16 copies (next A operand)
    SEEME_TRY(rn_gemm(h, b.c3, h->t2.h, h->t2.l, P, Mout, true, R, xo.f, xo.h, xo.l, s));
    cur ^= 1;
    H = Ho;
  }
  const size_t n = (size_t)B * 2048;
  rn_avgpool_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->x[cur].f, out, B, 49, 2048);
  SEEME_LAUNCH_CHECK();
  return SEEME_OK;
}

extern "C" int seeme_resnet50_forward(seeme_resnet50_t h, const float* images, int B, float* feat2048, float* emb256, void* stream) {
  SEEME_REQUIRE(h && images && (feat2048 || emb256), SEEME_EINVAL, "seeme_resnet50_forward: null argument");
  SEEME_REQUIRE(B > 0, SEEME_EINVAL, "seeme_resnet50_forward: empty batch");
  SEEME_REQUIRE(B <= h->max_batch, SEEME_ECAP, "seeme_resnet50_forward: batch %d exceeds capacity %d", B, h->max_batch);
  cudaStream_t s = (cudaStream_t)stream;
  for (int b0 = 0; b0 < B; b0 += h->chunk) {
    const int nb = B - b0 < h->chunk ? B - b0 : h->chunk;
    SEEME_TRY(rn_chunk(h, images + (size_t)b0 * 3 * 224 * 224, nb, h->feat + (size_t)b0 * 2048, s));
  }
  if (feat2048) SEEME_CUDA(cudaMemcpyAsync(feat2048, h->feat, (size_t)B * 2048 * 4, cudaMemcpyDeviceToDevice, s));
  if (emb256) {   // output_images: Sequential(ReLU, Linear(2048, 256))
    SEEME_TRY(to_bf16_split(h->feat, 2048, B, 2048, h->feat_h, h->feat_l, 2048, 1, s));
    ActBuf a; a.h = h->feat_h; a.l = h->feat_l; a.ld = 2048;
    ActBuf o; o.f = emb256; o.ld = 256;
    SEEME_TRY(run_linear(h->tail, a, nullptr, B, ACT_NONE, nullptr, 0, o, 3, s));
  }
  return SEEME_OK;
}
