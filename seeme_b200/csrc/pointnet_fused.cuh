// Fused residual-block kernel of the scene encoder (pointnet_fused.cu).
#pragma once
#include "common.cuh"

namespace seeme {

// bytes of one block's packed fp16 weight-chunk blob
size_t pf_blob_bytes();
// pack Ws[:, :256] ([256,512] fp32), W0[:, :256] ([256,512]) and W1 ([256,256]) of one block (default stream)
int pf_pack_block(const float* ws, const float* w0, const float* w1, void* blob);
// x_in / x_out: [samples, n_points, 256] fp16 (x_out nullable: only the pooled column max is produced);
// bias_h / bias_o: [samples, 256] fp32; colmax: [samples, 256] order-preserving uint (zeroed by the caller)
int pf_block_forward(const void* x_in, void* x_out, const void* w_blob, const float* bias_h, const float* bias_o, unsigned* colmax,
                     int samples, int n_points, int h_in_tmem, int prof_id, cudaStream_t s);

// block 0 (+ fc_pos): weight-chunk blob from W0 [256,512] / W1 [256,256] and the packed (Wp, bp) table [512] float4
int pf_pack_block0(const float* w0, const float* w1, const float* wp, const float* bp, void* blob, void* wpb);
// constant tiles (32 KB) of the tensor-core block 0: fc_pos, the biases and the rank-3 shortcut fold as K = 16 MMA operands.
// pfold [256] float4 and cst0 [256] must already be computed (device pointers)
int pf_pack_block0_ct(const float* wp, const float* bp, const float* b0, const float* pfold, const float* cst0, void* ctblob);
// xyz [samples, n_points, 3] fp32 -> x_out [samples, n_points, 256] fp16 and the pooled column max;
// b0 [256], cst0 [256] = Ws bp + b1, pfold [256] float4 = rank-3 fold of the shortcut (CUDA-core generator only)
int pf_block0_forward(const float* xyz, void* x_out, const void* w_blob, const void* wpb, const void* ctblob, const float* b0,
                      const float* cst0, const float* pfold, unsigned* colmax, int samples, int n_points, int prof_id, cudaStream_t s);

}  // namespace seeme
