// Tensor-core SMPL blend + skinning kernel (smpl_tc.cu).
#pragma once
#include "common.cuh"

namespace seeme {

size_t smpl_tc_basis_elems();     // elements of each of the packed bf16 basis buffers (hi, lo)
// basis [SK][3][6912] fp32 -> bf16 (hi, lo) [3*6912][256]  (default stream, create time)
int smpl_tc_pack_basis(const float* basis, int SK, void* bh, void* bl);
// coef [F, n_coef] fp32 (pitch ld_coef); ch / cl: bf16 [F_pad, 256] scratch whose columns >= n_coef are zero;
// A [F,24,12]; vt [3][6912]; w4/i4 [6912][4]; verts [F,6890,3]
int smpl_skin_tc(const void* bh, const void* bl, const float* coef, int ld_coef, int n_coef, void* ch, void* cl, const float* A,
                 const float* vt, const float* w4, const unsigned char* i4, int F, float* verts, int prof_id, cudaStream_t s);

// version 2 (transform blend on the tensor cores as well; any number of skinning weights per vertex):
// w24 [6912][24] dense fp32 -> wblob (smpl_tc_wblob_bytes()) at create; same call otherwise
int smpl_tc_pack_basis_f16(const float* basis, int SK, void* bh, void* bl);   // version 2 streams an fp16 (hi, lo) basis
size_t smpl_tc_wblob_bytes();
size_t smpl_tc_aop_bytes(size_t frames);      // per-call scratch for the fp16 transform operands
int smpl_tc_pack_wtiles(const float* w24, void* wblob);
int smpl_skin_tc2(const void* bh, const void* bl, const float* coef, int ld_coef, int n_coef, void* ch, void* cl, const float* A,
                  const float* vt, const void* wblob, void* aopblob, int F, float* verts, int prof_id, cudaStream_t s);

}  // namespace seeme
