// fp32 parity path of avdf_conv_gemm: CUDA-core implicit-GEMM conv (raw accumulators to a workspace)
// followed by a warp-per-row epilogue. This is the "fp32 mode" of BASELINE.json (1e-4 bar) and the
// on-device cross-check of the tcgen05 path; it is not the throughput path.
// Restates MaskedConv1D (libs/modeling/blocks.py:41-63) + LayerNorm (blocks.py:97-112).
#include "gemm_common.cuh"

namespace avdf {

constexpr int SBM = 64, SBN = 64, SBK = 16;

struct SimtParams {
  SegInfo seg;
  int seg_m_start[AVDF_MAX_LEVELS + 1];   // prefix of batch * t_out over levels
  const float* a; const float* w; float* raw;
  int n_out, c_in, taps, stride, m_total, k_total;
  int tap_tab[AVDF_MAX_TAPS];
};

__device__ __forceinline__ void simt_decode_row(const SimtParams& p, int m, int& seg, int& b, int& t) {
  seg = 0;
  while (seg + 1 < p.seg.n_seg && m >= p.seg_m_start[seg + 1]) ++seg;
  const int local = m - p.seg_m_start[seg];
  b = local / p.seg.t_out[seg];
  t = local - b * p.seg.t_out[seg];
}

__global__ void __launch_bounds__(256) conv_gemm_f32_kernel(const SimtParams p) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;
  // loader mapping: one float4 of A and one of W per thread per k-chunk
  const int lrow = tid >> 2, lk = (tid & 3) << 2;
  int a_seg = 0, a_b = 0, a_t = 0;
  const bool a_ok = (m0 + lrow) < p.m_total;
  if (a_ok) simt_decode_row(p, m0 + lrow, a_seg, a_b, a_t);
  const int t_in_len = a_ok ? p.seg.t_out[a_seg] * p.stride : 0;
  const float* a_base = p.a + ((size_t)a_b * p.seg.a_rows + (a_ok ? p.seg.a_row[a_seg] : 0)) * p.c_in;
  const bool w_ok = (n0 + lrow) < p.n_out;
  const float* w_base = p.w + (size_t)(n0 + lrow) * p.k_total;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int kk = 0; kk < p.k_total; kk += SBK) {
    const int tap = kk / p.c_in, c = kk - tap * p.c_in;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a_ok) {
      const int ti = p.stride * a_t + p.tap_tab[tap];
      if (ti >= 0 && ti < t_in_len) av = *reinterpret_cast<const float4*>(a_base + (size_t)ti * p.c_in + c + lk);
    }
    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w_ok) wv = *reinterpret_cast<const float4*>(w_base + kk + lk);
    As[lk + 0][lrow] = av.x; As[lk + 1][lrow] = av.y; As[lk + 2][lrow] = av.z; As[lk + 3][lrow] = av.w;
    Bs[lk + 0][lrow] = wv.x; Bs[lk + 1][lrow] = wv.y; Bs[lk + 2][lrow] = wv.z; Bs[lk + 3][lrow] = wv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {a4.x, a4.y, a4.z, a4.w}, br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.m_total) continue;
    int seg, b, t;
    simt_decode_row(p, m, seg, b, t);
    const size_t orow = (size_t)b * p.seg.o_rows + p.seg.o_row[seg] + t;
    const int n = n0 + tx * 4;
    if (n + 3 < p.n_out)
      *reinterpret_cast<float4*>(p.raw + orow * p.n_out + n) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    else
      for (int j = 0; j < 4; ++j) if (n + j < p.n_out) p.raw[orow * p.n_out + n + j] = acc[i][j];
  }
}

// warp per output row; two-pass LayerNorm in fp32
__global__ void __launch_bounds__(256) row_epilogue_kernel(const SimtParams p, const EpiParams e) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= p.m_total) return;
  int seg, b, t;
  simt_decode_row(p, m, seg, b, t);
  const size_t orow = (size_t)b * p.seg.o_rows + p.seg.o_row[seg] + t;
  const int N = e.n_out;
  const float* raw = p.raw + orow * N;
  const float mk = e.row_mask ? (e.row_mask[orow] ? 1.f : 0.f) : 1.f;
  float mean = 0.f, rstd = 1.f;
  if (e.ln_w) {
    float s = 0.f;
    for (int n = lane; n < N; n += 32) s += (raw[n] + (e.bias ? e.bias[n] : 0.f)) * mk;
    mean = warp_sum(s) / (float)N;
    float q = 0.f;
    for (int n = lane; n < N; n += 32) { float d = (raw[n] + (e.bias ? e.bias[n] : 0.f)) * mk - mean; q += d * d; }
    rstd = 1.f / sqrtf(warp_sum(q) / (float)N + 1e-5f);
  }
  for (int n = lane; n < N; n += 32) {
    float v = (raw[n] + (e.bias ? e.bias[n] : 0.f)) * mk;
    if (e.ln_w) v = (v - mean) * rstd * e.ln_w[n] + e.ln_b[n];
    v = apply_act(v, e.act);
    if (e.pe) v += e.pe[(size_t)t * N + n] * mk;
    if (e.residual) v = e.residual[orow * N + n] * mk + (e.gamma ? e.gamma[n] : 1.f) * v;
    if (e.out_f32) e.out_f32[orow * N + n] = v;
    if (e.out_h) {
      if (e.out_h_f16) reinterpret_cast<__half*>(e.out_h)[orow * N + n] = __float2half_rn(v);
      else reinterpret_cast<__nv_bfloat16*>(e.out_h)[orow * N + n] = __float2bfloat16_rn(v);
    }
  }
}

static int conv_gemm_f32_one(const avdf_conv_gemm_args* a, cudaStream_t st);

// Segments with their own weight block (seg_w_row != 0) run as one launch per segment with the weight / vector
// pointers advanced on the host; the shared-weight case (pyramid levels) is a single launch.
int conv_gemm_f32(const avdf_conv_gemm_args* a, cudaStream_t st) {
  bool shared = true;
  for (int i = 0; i < a->n_seg; ++i) shared = shared && a->seg_w_row[i] == 0;
  if (shared) return conv_gemm_f32_one(a, st);
  for (int i = 0; i < a->n_seg; ++i) {
    avdf_conv_gemm_args b = *a;
    const int wo = a->seg_w_row[i];
    b.n_seg = 1;
    b.seg_t_out[0] = a->seg_t_out[i]; b.seg_a_row[0] = a->seg_a_row[i]; b.seg_o_row[0] = a->seg_o_row[i]; b.seg_w_row[0] = 0;
    b.w = reinterpret_cast<const float*>(a->w) + (size_t)wo * a->taps * a->c_in;
    if (a->bias) b.bias = a->bias + wo;
    if (a->ln_w) { b.ln_w = a->ln_w + wo; b.ln_b = a->ln_b + wo; }
    if (a->gamma) b.gamma = a->gamma + wo;
    int rc = conv_gemm_f32_one(&b, st);
    if (rc) return rc;
  }
  return AVDF_OK;
}

static int conv_gemm_f32_one(const avdf_conv_gemm_args* a, cudaStream_t st) {
  SimtParams p{};
  fill_seg(a, p.seg);
  int m = 0;
  for (int i = 0; i < a->n_seg; ++i) { p.seg_m_start[i] = m; m += a->batch * a->seg_t_out[i]; }
  p.seg_m_start[a->n_seg] = m;
  p.m_total = m;
  p.a = reinterpret_cast<const float*>(a->a); p.w = reinterpret_cast<const float*>(a->w);
  p.raw = reinterpret_cast<float*>(a->workspace);
  p.n_out = a->n_out; p.c_in = a->c_in; p.taps = a->taps; p.stride = a->stride; p.k_total = a->taps * a->c_in;
  fill_taps(a, p.tap_tab);
  AVDF_CHECK_ARG(!a->ln_after_residual, "ln_after_residual is a feature of the 16-bit path");
  AVDF_CHECK_ARG(!a->dot_out, "dot_out is a feature of the 16-bit path");
  AVDF_CHECK_ARG(a->c_in % SBK == 0, "c_in must be a multiple of 16");
  AVDF_CHECK_ARG(a->workspace && a->workspace_bytes >= (size_t)a->batch * a->o_rows_per_video * a->n_out * sizeof(float),
                 "fp32 path needs a workspace of batch*o_rows_per_video*n_out floats");
  if (m == 0) return AVDF_OK;
  EpiParams e{};
  fill_epi(a, e);
  dim3 grid(ceil_div(m, SBM), ceil_div(a->n_out, SBN));
  conv_gemm_f32_kernel<<<grid, 256, 0, st>>>(p);
  int rc = check_launch("conv_gemm_f32_kernel");
  if (rc) return rc;
  row_epilogue_kernel<<<ceil_div(m, 8), 256, 0, st>>>(p, e);
  return check_launch("row_epilogue_kernel");
}

}  // namespace avdf

using namespace avdf;

extern "C" size_t avdf_conv_gemm_workspace_bytes(const avdf_conv_gemm_args* a) {
  if (!a || a->dtype != AVDF_DTYPE_F32) return 0;
  return (size_t)a->batch * a->o_rows_per_video * a->n_out * sizeof(float);
}

extern "C" int avdf_conv_gemm(const avdf_conv_gemm_args* a, void* stream) {
  AVDF_CHECK_ARG(a != nullptr, "args is null");
  AVDF_CHECK_ARG(a->batch >= 0 && a->n_out > 0 && a->c_in > 0, "bad sizes");
  AVDF_CHECK_ARG(a->tap_mode == 0 || a->tap_mode == 1, "tap_mode must be 0 or 1");
  if (a->tap_rows) AVDF_CHECK_ARG(a->taps >= 1 && a->taps <= AVDF_MAX_TAPS && a->stride == 1 && !a->tap_mode, "tap_rows: 1 <= taps <= 9, stride 1, tap_mode 0");
  else AVDF_CHECK_ARG(a->tap_mode ? (a->taps == 2 && a->stride == 1) : (a->taps == 1 || a->taps == 3), "taps must be 1 or 3 (centred) or 2 (forward, stride 1)");
  AVDF_CHECK_ARG(!a->ln_after_residual || (a->ln_w && a->residual && a->out_f32 && a->out_h && a->act == AVDF_ACT_NONE && !a->pe),
                 "ln_after_residual needs ln_w, residual, out_f32 and out_h, no activation and no pe");
  AVDF_CHECK_ARG(a->stride == 1 || a->stride == 2, "stride must be 1 or 2");
  AVDF_CHECK_ARG(a->n_seg >= 1 && a->n_seg <= AVDF_MAX_LEVELS, "n_seg out of range");
  AVDF_CHECK_ARG(a->a && a->w, "null operand");
  AVDF_CHECK_ARG(a->out_f32 || a->out_h || a->dot_out, "no output");
  AVDF_CHECK_ARG(!a->out_h || a->out_h_dtype == AVDF_DTYPE_BF16 || a->out_h_dtype == AVDF_DTYPE_F16, "out_h_dtype must be BF16 or F16");
  AVDF_CHECK_ARG((a->ln_w == nullptr) == (a->ln_b == nullptr), "ln_w / ln_b must come together");
  for (int i = 0; i < a->n_seg; ++i) AVDF_CHECK_ARG(a->seg_t_out[i] > 0, "seg_t_out must be positive");
  for (int i = 0; i < a->n_seg; ++i)
    AVDF_CHECK_ARG(a->seg_w_row[i] >= 0 && a->seg_w_row[i] + a->n_out <= (a->n_w_rows > 0 ? a->n_w_rows : a->n_out), "seg_w_row out of range");
  if (a->batch == 0) return AVDF_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->dtype == AVDF_DTYPE_F32) return conv_gemm_f32(a, st);
  if (a->dtype == AVDF_DTYPE_BF16 || a->dtype == AVDF_DTYPE_F16) return conv_gemm_tc(a, st);
  set_error("avdf_conv_gemm: unknown dtype %d", a->dtype);
  return AVDF_ERR_INVALID;
}
