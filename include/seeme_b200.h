/*
 * seeme_b200 -- C ABI of the B200-native (sm_100a) SEE-ME inference hot path.
 *
 * The reference (L-Scofano/SEEME) is pure Python and has no FFI plugin interface; its extension
 * points are the YAML `target:` classes (mld/config.py:25-32) and the attributes of the `MLD`
 * LightningModule (mld/models/modeltype/mld.py:149-172,185,257-287).  This header is what a C-ABI
 * replacement of those operators exports; the Python mirror of the reference classes in
 * `seeme_b200/` binds it through ctypes (see INTEGRATION.md for the stub a reference maintainer
 * would add).  Each entry point cites the reference interface it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch's allocator) unless marked HOST;
 *    tensors are dense, row-major, fp32 unless stated otherwise;
 *  - every call enqueues work on the `stream` argument (a cudaStream_t passed as void*) and returns
 *    without synchronising; results are valid in stream order;
 *  - a handle is bound to the device that was current at `*_create`, is not thread-safe, owns all of
 *    its workspace (allocated at create; nothing is allocated afterwards, so the calls can be
 *    captured in a CUDA graph) and packs/re-lays-out the weights once at create -- the caller's
 *    `state_dict` tensors stay untouched and interchangeable with reference checkpoints;
 *  - return value 0 = OK, negative = SEEME_E* ; `seeme_last_error()` gives the message of the last
 *    failure on the calling thread.  There is no CPU fallback anywhere behind this ABI.
 */
#ifndef SEEME_B200_H
#define SEEME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEEME_ABI_VERSION 1

enum {
  SEEME_OK = 0,
  SEEME_EINVAL = -1,   /* bad argument / unsupported shape */
  SEEME_ECUDA = -2,    /* CUDA runtime error (message carries cudaGetErrorString) */
  SEEME_ENOMEM = -3,   /* workspace allocation failed at create */
  SEEME_ECAP = -4      /* call exceeds the capacity given at create (max_batch, max_frames, ...) */
};

int seeme_abi_version(void);
const char* seeme_last_error(void);
/* number of kernels launched by this library on the calling process since load (for bench.py's
 * `gpu_launches`); graph replays count the kernels inside the graph. */
unsigned long long seeme_launch_count(void);

/* Per-kernel-class device timing for bench.py's live roofline measurement: with profiling enabled,
 * every launch of an instrumented kernel class is bracketed by a cudaEvent pair on its stream.
 * ids: 0 scene-encoder GEMMs, 1 SMPL blend+skinning kernel, 2 SMPL pose kernel, 3 sampler graph,
 * 4 VAE attention, 5 tcgen05 GEMM.  seeme_prof_read synchronises, returns and clears the sums. */
int seeme_prof_enable(int on);
int seeme_prof_read(int id, double* total_ms, long long* count);

/* Self-test hook for the tcgen05/TMA linear used by every dense contraction on the path:
 * Y[M,N] = act(A[M,K] W[N,K]^T + bias[N]) (+ R[M,N]); fp32 device inputs are converted to (split) bf16
 * internally.  K multiple of 64, N multiple of 128; npass 1 = bf16, 3 = split-bf16 (hi.hi+lo.hi+hi.lo).
 * colmax (nullable): [ceil(M/group), N] order-preserving-uint column max per group of rows (zeroed by
 * the caller).  Ysplit / Zsplit (nullable, fp32 [M,N]): hi+lo of the kernel's bf16 outputs of the result and
 * of relu(result) (when Zsplit is given the fp32 output Y is not produced: they share staging).
 * Synchronises the stream. */
int seeme_test_umma_linear(const float* A, const float* W, const float* bias, const float* R, float* Y,
                           int M, int N, int K, int act, int npass, unsigned* colmax,
                           int colmax_group_rows, float* Ysplit, float* Zsplit, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Scene encoder.  Replaces `ProHMRScene.encode_scene` -> `ResnetPointnet.forward`
 * (EgoHMR/models/prohmr/prohmr_scene.py:102-104, EgoHMR/models/respointnet.py:33-59) and
 * `MLD.output_scene` = Sequential(ReLU, Linear(512,256)) (mld/models/modeltype/mld.py:257-261).
 * Weight order (26 tensors): fc_pos_0.{weight,bias}; for i in 0..3: block_i.fc_0.{weight,bias},
 * block_i.fc_1.{weight,bias}, block_i.shortcut.weight; fc_c.{weight,bias}; output_scene.1.{weight,bias}.
 * ---------------------------------------------------------------------------------------------- */
typedef struct seeme_pointnet* seeme_pointnet_t;
#define SEEME_POINTNET_NUM_TENSORS 26
int seeme_pointnet_create(seeme_pointnet_t* out, const float* const* weights /*HOST array of device ptrs*/,
                          int n_weights, int max_batch, int max_points);
/* Same with an explicit GEMM operand format for the per-point contractions (fp32 accumulation everywhere):
 *   3  split-bf16 (hi.hi + lo.hi + hi.lo, ~16 mantissa bits; the default of seeme_pointnet_create unless the
 *      environment variable SEEME_POINTNET_PRECISION overrides it), layer-by-layer tcgen05 GEMMs;
 *   1  plain bf16, layer-by-layer;   0  fp32 CUDA-core GEMMs, layer-by-layer;
 *   16 fp16 operands, residual blocks 1..3 fused into one persistent tcgen05 kernel each (activations stay in
 *      shared/tensor memory between the block's three GEMMs); 17 = 16 with the hidden activation staged through
 *      shared memory instead of tensor memory. */
int seeme_pointnet_create_ex(seeme_pointnet_t* out, const float* const* weights, int n_weights, int max_batch,
                             int max_points, int precision);
/* pcd [B,N,3] -> feat512 [B,512] (= encode_scene, may be NULL) and emb256 [B,256]
 * (= output_scene(encode_scene), may be NULL). */
int seeme_pointnet_forward(seeme_pointnet_t h, const float* pcd, int B, int N, float* feat512,
                           float* emb256, void* stream);
int seeme_pointnet_destroy(seeme_pointnet_t h);

/* ------------------------------------------------------------------------------------------------
 * Image backbone.  Replaces `ProHMRScene.encode_image` -> `ResNet.forward`
 * (EgoHMR/models/prohmr/prohmr_scene.py:99-100, EgoHMR/models/resnet.py:60-180: ResNet-50, Bottleneck
 * [3,4,6,3], eval-mode BatchNorm, mean over the final 7x7 map) and `MLD.output_images` =
 * Sequential(ReLU, Linear(2048,256)) (mld/models/modeltype/mld.py:251-255).
 * Weight order (267 tensors; 265 `backbone.*` state_dict names, then output_images.1.{weight,bias}): conv1.weight, bn1.{weight,bias,running_mean,
 * running_var}; then for every block layer{1..4}.{0..}: conv1.weight, bn1.{4}, conv2.weight, bn2.{4},
 * conv3.weight, bn3.{4} and, for block 0 of each layer, downsample.0.weight, downsample.1.{4}.
 * ---------------------------------------------------------------------------------------------- */
typedef struct seeme_resnet50* seeme_resnet50_t;
#define SEEME_RESNET50_NUM_TENSORS 267
int seeme_resnet50_create(seeme_resnet50_t* out, const float* const* weights /*HOST array of device ptrs*/,
                          int n_weights, int max_batch);
/* images [B,3,224,224] fp32 (NCHW, as the reference feeds them) -> feat2048 [B,2048] (= encode_image, may be NULL)
 * and emb256 [B,256] (= output_images(encode_image), may be NULL) */
int seeme_resnet50_forward(seeme_resnet50_t h, const float* images, int B, float* feat2048, float* emb256,
                           void* stream);
int seeme_resnet50_destroy(seeme_resnet50_t h);

/* ------------------------------------------------------------------------------------------------
 * Motion VAE.  Replaces `MldVae.encode` / `MldVae.decode`
 * (mld/models/architectures/mld_vae.py:128-193, 195-256; stacks in
 * mld/models/operator/cross_attention.py:18-147,258-403).
 * Weight order (169 tensors): global_motion_token, query_pos_encoder.pe, query_pos_decoder.pe,
 * skel_embedding.{w,b}, final_layer.{w,b}; encoder.norm.{w,b}, encoder.linear_blocks.{0,1}.{w,b};
 * for blk in [input_blocks.0, input_blocks.1, middle_block, output_blocks.0, output_blocks.1]:
 *   encoder.blk.{self_attn.in_proj_weight, self_attn.in_proj_bias, self_attn.out_proj.{w,b},
 *                linear1.{w,b}, linear2.{w,b}, norm1.{w,b}, norm2.{w,b}};
 * decoder.norm.{w,b}, decoder.linear_blocks.{0,1}.{w,b}; for blk (same order):
 *   decoder.blk.{self_attn.(4), multihead_attn.(4), linear1.{w,b}, linear2.{w,b}, norm1, norm2, norm3 .{w,b}}.
 * ---------------------------------------------------------------------------------------------- */
typedef struct seeme_vae* seeme_vae_t;
#define SEEME_VAE_NUM_TENSORS 169
int seeme_vae_create(seeme_vae_t* out, const float* const* weights, int n_weights, int nfeats,
                     int max_batch, int max_frames);
/* features [B,T,nfeats], lengths int32 [B] (device), eps [B,256] (the N(0,1) draw of
 * Normal.rsample, mld_vae.py:190-192) -> z = mu + std*eps [B,256], mu [B,256], std [B,256]
 * (mu/std may be NULL). */
int seeme_vae_encode(seeme_vae_t h, const float* features, const int32_t* lengths, const float* eps,
                     int B, int T, float* z, float* mu, float* std, void* stream);
/* z [B,256], lengths int32 [B] (device), T = max(lengths) -> feats [B,T,nfeats]; padded frames are
 * NOT zeroed, like the reference (mld_vae.py:253 is commented out). */
int seeme_vae_decode(seeme_vae_t h, const float* z, const int32_t* lengths, int B, int T,
                     float* feats, void* stream);
int seeme_vae_destroy(seeme_vae_t h);

/* ------------------------------------------------------------------------------------------------
 * Denoiser and sampler.  `seeme_denoiser_forward` replaces `MldDenoiser.forward`
 * (mld/models/architectures/mld_denoiser.py:151-244; blocks mdiff_transformer.py:137-163,206-304;
 * skip stack cross_attention.py:67-83; timestep embedding tools/embeddings.py:245-322).
 * `seeme_sampler_run` replaces `MLD._diffusion_reverse` (mld/models/modeltype/mld.py:432-511):
 * the 50-step loop with the classifier-free-guidance combine (:488-492) and
 * `DDIMScheduler.step` (diffusers; configs/modules/scheduler.yaml:1-14) fused in.
 * Weight order (201 tensors): time_embedding.linear_1.{w,b}, time_embedding.linear_2.{w,b},
 * query_pos.pe, encoder.norm.{w,b}, encoder.linear_blocks.{0,1}.{w,b}; for blk (order as above):
 *   sa_block.{self_attn.in_proj_weight, in_proj_bias, out_proj.{w,b}, linear1.{w,b}, linear2.{w,b},
 *             norm1.{w,b}, norm2.{w,b}},
 *   ca_block.{norm.{w,b}, text_norm.{w,b}, query.{w,b}, key.{w,b}, value.{w,b},
 *             proj_out.emb_layers.1.{w,b}, proj_out.norm.{w,b}, proj_out.out_layers.2.{w,b}},
 *   ffn.{linear1.{w,b}, linear2.{w,b}, proj_out.emb_layers.1.{w,b}, proj_out.norm.{w,b},
 *        proj_out.out_layers.2.{w,b}}.
 * ---------------------------------------------------------------------------------------------- */
typedef struct seeme_denoiser* seeme_denoiser_t;
#define SEEME_DENOISER_NUM_TENSORS 201
#define SEEME_MAX_COND_TOKENS 4
int seeme_denoiser_create(seeme_denoiser_t* out, const float* const* weights, int n_weights,
                          int max_rows);
/* Build the per-timestep tables (time embedding, time-token K/V, FiLM scale/shift of all blocks) for
 * `timesteps` (HOST int32 [n]).  `sinusoid` (HOST fp32 [n,256], may be NULL) is
 * `get_timestep_embedding(t, 256, flip_sin_to_cos=True, downscale_freq_shift=0)`
 * (tools/embeddings.py:245-285) computed by the caller; NULL = computed internally.  Called
 * implicitly by the two entry points below when their timesteps differ from the cached ones. */
int seeme_denoiser_set_time_table(seeme_denoiser_t h, const int32_t* timesteps, int n,
                                  const float* sinusoid, void* stream);
/* sample [R,256] (R = B' rows; latent_dim[0] == 1), timestep (integer, as `t` in mld.py:467),
 * cond [Nc,R,256] (= encoder_hidden_states as the denoiser receives it) -> out [R,256]. */
int seeme_denoiser_forward(seeme_denoiser_t h, const float* sample, int timestep, const float* cond,
                           int Nc, int R, float* out, void* stream);
/* x_T [B,256]; cond [Nc,R,256] with R = 2B when guidance_scale > 1 (rows [0,B) are the branch the
 * reference calls "uncond", rows [B,2B) the "text" branch, mld.py:488-492) else R = B.
 * timesteps: HOST int32 [n_steps] (e.g. 981,961,...,1); coef: HOST fp32 [n_steps,4] =
 * {sqrt(1-abar_t), sqrt(abar_t), sqrt(abar_prev), sqrt(1-abar_prev)} so that
 *   x0 = (x - c0*eps)/c1 ; x_prev = c2*x0 + c3*eps      (DDIM, eta = 0).
 * -> z [B,256].  The whole chain is captured in a CUDA graph owned by the handle. */
int seeme_sampler_run(seeme_denoiser_t h, const float* x_T, const float* cond, int Nc, int B,
                      float guidance_scale, int n_steps, const int32_t* timesteps, const float* coef,
                      float* z, void* stream);
/* Sampler / denoiser back-end of this handle.  SEEME_SAMPLER_PERSISTENT (default): ONE launch of the persistent
 * cluster kernel per run (8-CTA clusters per 128-row tile; lowest latency, 8 SMs per tile for the whole run).
 * SEEME_SAMPLER_GRAPH: the CUDA graph of small kernels (less SM time per run: the choice when many chains share the
 * GPU with the scene encoder, as in the batch pipeline).  Same results to fp32 rounding. */
#define SEEME_SAMPLER_PERSISTENT 0
#define SEEME_SAMPLER_GRAPH 1
#define SEEME_SAMPLER_TILE 2   /* persistent kernel, ONE CTA per 128-row tile (no cluster): 1/8 of the SM time */
int seeme_denoiser_set_backend(seeme_denoiser_t h, int backend);
int seeme_denoiser_destroy(seeme_denoiser_t h);
/* standalone `DDIMScheduler.step` (eta = 0, epsilon prediction) for the scheduler duck type. */
int seeme_ddim_step(const float* eps, const float* sample, float* prev, size_t n, float c0, float c1,
                    float c2, float c3, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SMPL body model.  Replaces `smplx.SMPL.forward` (smplx==0.1.28, requirements.txt:170; call sites
 * mld/models/modeltype/mld.py:1476-1482,1522-1528,1557-1563,1677-1683,1723-1729), `aa_to_quat`
 * (mld/utils/geometry2.py:33-54) and, in the *_feats variant, `datamodule.renorm`
 * (mld/data/EgoBody.py:151-157) plus the slicing at mld.py:1458-1520 / 1655-1722.
 * ---------------------------------------------------------------------------------------------- */
typedef struct seeme_smpl* seeme_smpl_t;
int seeme_smpl_create(seeme_smpl_t* out, const float* v_template /*[6890,3]*/,
                      const float* shapedirs /*[6890,3,10]*/, const float* posedirs /*[207,20670]*/,
                      const float* J_regressor /*[24,6890]*/, const float* lbs_weights /*[6890,24]*/,
                      const int32_t* parents /*HOST [24]*/, int max_frames);
/* betas [F,10], body_pose [F,69], global_orient [F,3], transl [F,3] or NULL ->
 * vertices [F,6890,3] (NULL = skip skinning), joints [F,24,3] (the 24 kinematic joints; ego_eval
 * only reads joints[:, :24], mld.py:1485-1487), quat [F,4] = aa_to_quat(global_orient) or NULL. */
int seeme_smpl_forward(seeme_smpl_t h, const float* betas, const float* body_pose,
                       const float* global_orient, const float* transl, int F, float* vertices,
                       float* joints, float* quat, void* stream);
/* Fused prologue: feats [F,Dn] normalised network output / ground truth (Dn = 75; for gimo the
 * caller passes the 69-d concat of dims 0:66 and -3:), mean/std float64 [Dn] (device);
 * n_body = 69 (egobody) or 63 (gimo: 6 zero hand DoF appended, mld.py:1659-1665).
 * m_out float64 [F,Dn] = renorm(feats) (may be NULL). */
int seeme_smpl_forward_feats(seeme_smpl_t h, const float* feats, int Dn, const double* mean,
                             const double* std, int n_body, const float* betas, int F,
                             double* m_out, float* vertices, float* joints, float* quat, void* stream);
int seeme_smpl_destroy(seeme_smpl_t h);

#ifdef __cplusplus
}
#endif
#endif /* SEEME_B200_H */
