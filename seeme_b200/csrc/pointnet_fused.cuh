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

}  // namespace seeme
