// Device-side descriptors shared by the two conv-GEMM paths (fp32 CUDA-core, bf16 tcgen05).
#pragma once
#include "common.cuh"

namespace avdf {

struct EpiParams {
  const float* bias; const unsigned char* row_mask; const float* ln_w; const float* ln_b; int act;
  const float* pe; const float* residual; const float* gamma;
  float* out_f32; void* out_h; int out_h_f16;
  int n_out;
  const float* dot_w; int dot_n; float* dot_out;      // row dot products (avdf_conv_gemm_args.dot_*)
};

struct SegInfo {
  int n_seg, batch;
  int t_out[AVDF_MAX_LEVELS];
  int a_row[AVDF_MAX_LEVELS];
  int o_row[AVDF_MAX_LEVELS];
  int w_row[AVDF_MAX_LEVELS];
  long long a_rows, o_rows;     // rows per video in A / out
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == AVDF_ACT_RELU) return fmaxf(v, 0.f);
  if (act == AVDF_ACT_GELU) return gelu_erf(v);
  return v;
}

int conv_gemm_f32(const avdf_conv_gemm_args* a, cudaStream_t st);
int conv_gemm_tc(const avdf_conv_gemm_args* a, cudaStream_t st);

inline void fill_epi(const avdf_conv_gemm_args* a, EpiParams& e) {
  e.bias = a->bias; e.row_mask = a->row_mask; e.ln_w = a->ln_w; e.ln_b = a->ln_b; e.act = a->act;
  e.pe = a->pe; e.residual = a->residual; e.gamma = a->gamma;
  e.out_f32 = a->out_f32; e.out_h = a->out_h; e.out_h_f16 = a->out_h_dtype == AVDF_DTYPE_F16;
  e.n_out = a->n_out;
  e.dot_w = a->dot_w; e.dot_n = a->dot_n; e.dot_out = a->dot_out;
}
inline void fill_taps(const avdf_conv_gemm_args* a, int* tab) {
  for (int j = 0; j < AVDF_MAX_TAPS; ++j)
    tab[j] = j >= a->taps ? 0 : (a->tap_rows ? a->tap_rows[j] : (a->tap_mode ? j : j - (a->taps >> 1)));
}
inline void fill_seg(const avdf_conv_gemm_args* a, SegInfo& s) {
  s.n_seg = a->n_seg; s.batch = a->batch;
  for (int i = 0; i < a->n_seg; ++i) { s.t_out[i] = a->seg_t_out[i]; s.a_row[i] = a->seg_a_row[i]; s.o_row[i] = a->seg_o_row[i]; s.w_row[i] = a->seg_w_row[i]; }
  s.a_rows = a->a_rows_per_video; s.o_rows = a->o_rows_per_video;
}

}  // namespace avdf
