// dense_hp.cuh -- entry points of the fp32-storage ("high precision", dtype code 2) variants of the dense-path
// streaming kernels.  dense_misc.cu / attention.cu route to these when their `bf16` selector is 2
// (half_precision_weights=false, models/guidance/dreammat_guidance.py:56,92-94).
#pragma once
#include <stdint.h>

int hp_groupnorm(const float* x, int n_img, int HW, int C, int ld, int G, const float* gamma, const float* beta, float eps,
                 int silu, float* y, int ldy, float* stats, void* stream);
int hp_groupnorm_bwd(const float* x, const float* dz, int n_img, int HW, int C, int G, const float* gamma, const float* beta,
                     float eps, int silu, const float* stats, const float* dx_add, float* dx, void* stream);
int hp_layernorm(const float* x, int64_t M, int C, const float* gamma, const float* beta, float eps, float* y, void* stream);
int hp_geglu(const float* h, int64_t M, int D, float* out, void* stream);
int hp_axpby2d(const float* s1, int64_t ld1, float a, const float* s2, int64_t ld2, float b, int64_t rows, int cols,
               float* dst, int64_t ldd, void* stream);
int hp_transpose(const float* x, int batch, int R, int C, int64_t ldx, int64_t bsx, float* y, int64_t ldy, int64_t bsy,
                 void* stream);
int hp_pad_convert(const float* x, int64_t rows, int cin, int cpad, float scale, float shift, float* y, void* stream);
int hp_attention(const float* q, int64_t ldq, int64_t q_bs, const float* k, const float* v, int64_t ldkv, int64_t kv_bs,
                 float* out, int64_t ldo, int64_t out_bs, int batch, int heads, int Nq, int Nk, float scale, void* stream);
