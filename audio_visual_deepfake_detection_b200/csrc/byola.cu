// the BYOL-A feature extractor (audio_feature/content_audio) in front of the localization path.
//   avdf_logmel            wav -> normalised log-mel spectrogram (extract_audio_feature_one.py:34-42, 66; torchaudio
//                          MelSpectrogram defaults: centre/reflect, periodic Hann, power 2, HTK mel triangles)
//   avdf_byola_conv1_pool  features.0-3 of AudioNTT2020Task6 (byol_a/models.py:54-57): conv3x3 1->64 + BN + ReLU + maxpool 2x2
//   avdf_byola_pool        the maxpool 2x2 behind features.4-6 / 8-10 (models.py:59-67); the two 64->64 convolutions
//                          themselves and the fc layers (models.py:70-76) run on the tensor cores as avdf_conv_gemm
//                          with a tap table (taps = 9 row offsets in the "grid layout" below)
// Grid layout of a level with M mel rows: all clips of a batch packed along time, one all-zero time step before every
// clip and behind the last one, every time step padded to M + 2 rows (a zero row below mel 0 and above mel M-1), 64
// channels per row: row = step * (M + 2) + mel + 1. A 3x3 convolution over (time, mel) is then a 1-D convolution over rows
// with the nine row offsets {-1,0,1} * (M + 2) + {-1,0,1}: zero padding at the clip and mel borders comes from the stored
// zero rows, and the outputs AT the padding rows are forced back to zero by the GEMM's row mask.
#include "common.cuh"

namespace avdf {

constexpr int FFT_N = 1024, FFT_HOP = 160, FFT_BINS = FFT_N / 2 + 1, MELS = 64, BCH = 64;

struct LogMelParams {
  const float* wav; const long long* clip_sample_off; const int* clip_frame_off; const int* clip_pair_off; int n_clips, total_frames;
  const float* window; const float2* twiddle; const int* mel_lo; const int* mel_cnt; const float* mel_w; int mel_stride;
  float mean, inv_std, eps; float* lms;
};

__device__ __forceinline__ int find_clip(const int* off, int n, int idx) {      // off[c] <= idx < off[c + 1]
  int lo = 0, hi = n;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(off + mid) <= idx) lo = mid; else hi = mid; }
  return lo;
}

// One CTA = two consecutive frames of one clip: frame A rides in the real part, frame B in the imaginary
// part of ONE 1024-point complex FFT (radix-4 Stockham in shared memory, twiddles from a host-computed table), the two
// spectra are separated with the conjugate symmetry of real signals; power -> 64 mel triangles -> log -> normalise.
__global__ void __launch_bounds__(256) logmel_kernel(const LogMelParams p) {
  __shared__ float2 buf[2][FFT_N];
  __shared__ float pw[2][FFT_BINS + 3];
  const int tid = threadIdx.x;
  // frames are paired INSIDE a clip (its frames 2j and 2j + 1), so a clip's result does not depend on its batch neighbours
  const int c = find_clip(p.clip_pair_off, p.n_clips, blockIdx.x);
  const int j2 = 2 * (blockIdx.x - __ldg(p.clip_pair_off + c));
  const int fo = __ldg(p.clip_frame_off + c), nfr = __ldg(p.clip_frame_off + c + 1) - fo;
  const long long s0 = __ldg(p.clip_sample_off + c);
  const int ns = (int)(__ldg(p.clip_sample_off + c + 1) - s0);
  const float* src[2]; int nsamp[2], fr[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const bool on = j2 + h < nfr;
    src[h] = on ? p.wav + s0 : nullptr; nsamp[h] = ns; fr[h] = j2 + h;
  }
  const int f0 = fo + j2;                                      // row of frame A in lms
  const int f_end = fo + nfr;
  // first radix-4 pass straight from global memory (all its twiddles are 1): thread i owns samples i, i + 256, i + 512, i + 768
  float2 x4[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = tid + j * (FFT_N / 4);
    const float w = __ldg(p.window + n);
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      v[h] = 0.f;
      if (src[h]) {
        int i = fr[h] * FFT_HOP - FFT_N / 2 + n;               // centre = True, reflect padding (no edge repeat)
        if (i < 0) i = -i;
        if (i >= nsamp[h]) i = 2 * (nsamp[h] - 1) - i;
        v[h] = __ldg(src[h] + i) * w;
      }
    }
    x4[j] = make_float2(v[0], v[1]);
  }
  {
    const float2 t0 = make_float2(x4[0].x + x4[2].x, x4[0].y + x4[2].y), t1 = make_float2(x4[0].x - x4[2].x, x4[0].y - x4[2].y);
    const float2 t2 = make_float2(x4[1].x + x4[3].x, x4[1].y + x4[3].y);
    const float2 t3 = make_float2(x4[1].y - x4[3].y, x4[3].x - x4[1].x);           // -i (x1 - x3)
    const int o = tid << 2;
    buf[0][o] = make_float2(t0.x + t2.x, t0.y + t2.y);
    buf[0][o + 1] = make_float2(t1.x + t3.x, t1.y + t3.y);
    buf[0][o + 2] = make_float2(t0.x - t2.x, t0.y - t2.y);
    buf[0][o + 3] = make_float2(t1.x - t3.x, t1.y - t3.y);
  }
  __syncthreads();
  int cur = 0;
#pragma unroll 1
  for (int s = 4, sh = 6; s < FFT_N; s <<= 2, sh -= 2) {     // radix 4: thread = one butterfly; w = table[k * (256 / s)]
    const int i = tid;
    const int k = i & (s - 1);
    const int tb = k << sh;
    const float2 w1 = __ldg(p.twiddle + tb), w2 = __ldg(p.twiddle + 2 * tb), w3 = __ldg(p.twiddle + 3 * tb);
    const float2 x0 = buf[cur][i], a1 = buf[cur][i + FFT_N / 4], a2 = buf[cur][i + FFT_N / 2], a3 = buf[cur][i + 3 * FFT_N / 4];
    const float2 x1 = make_float2(w1.x * a1.x - w1.y * a1.y, w1.x * a1.y + w1.y * a1.x);
    const float2 x2 = make_float2(w2.x * a2.x - w2.y * a2.y, w2.x * a2.y + w2.y * a2.x);
    const float2 x3 = make_float2(w3.x * a3.x - w3.y * a3.y, w3.x * a3.y + w3.y * a3.x);
    const float2 t0 = make_float2(x0.x + x2.x, x0.y + x2.y), t1 = make_float2(x0.x - x2.x, x0.y - x2.y);
    const float2 t2 = make_float2(x1.x + x3.x, x1.y + x3.y);
    const float2 t3 = make_float2(x1.y - x3.y, x3.x - x1.x);                       // -i (x1 - x3)
    const int o = ((i - k) << 2) + k;
    buf[cur ^ 1][o] = make_float2(t0.x + t2.x, t0.y + t2.y);
    buf[cur ^ 1][o + s] = make_float2(t1.x + t3.x, t1.y + t3.y);
    buf[cur ^ 1][o + 2 * s] = make_float2(t0.x - t2.x, t0.y - t2.y);
    buf[cur ^ 1][o + 3 * s] = make_float2(t1.x - t3.x, t1.y - t3.y);
    cur ^= 1;
    __syncthreads();
  }
  for (int k = tid; k < FFT_BINS; k += 256) {                 // X_A = (Z[k] + conj Z[N-k]) / 2, X_B = (Z[k] - conj Z[N-k]) / 2i
    const float2 z = buf[cur][k], y = buf[cur][(FFT_N - k) & (FFT_N - 1)];
    const float ar = 0.5f * (z.x + y.x), ai = 0.5f * (z.y - y.y);
    const float br = 0.5f * (z.y + y.y), bi = 0.5f * (y.x - z.x);
    pw[0][k] = ar * ar + ai * ai;
    pw[1][k] = br * br + bi * bi;
  }
  __syncthreads();
  if (tid < 2 * MELS) {                                        // thread = (frame, mel filter): its triangle covers <= 39 consecutive bins
    const int h = tid >> 6, m = tid & (MELS - 1);
    if (f0 + h < f_end) {
      const int lo = __ldg(p.mel_lo + m), cnt = __ldg(p.mel_cnt + m);
      const float* w = p.mel_w + (size_t)m * p.mel_stride;
      float mel = 0.f;
      for (int k = 0; k < cnt; ++k) mel = fmaf(pw[h][lo + k], __ldg(w + k), mel);
      p.lms[(size_t)(f0 + h) * MELS + m] = (logf(mel + p.eps) - p.mean) * p.inv_std;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
struct Conv1Params {
  const float* lms; const int* clip_frame_off; const float* w; const float* b;
  const int* clip_of_step; const int* t_of_step; int n_steps; void* out; uint8_t* mask_out;
};

// thread = (output position of the level-1 grid layout, 8 of the 64 channels): the 4x4 log-mel patch under the 2x2 pooling
// window, four 3x3 convolutions per channel with the BatchNorm folded into w / b, max, ReLU
template <typename T>
__global__ void __launch_bounds__(256) byola_conv1_pool_kernel(const Conv1Params p) {
  __shared__ float sw[BCH * 9];
  __shared__ float sb[BCH];
  for (int i = threadIdx.x; i < BCH * 9; i += 256) sw[i] = __ldg(p.w + i);
  if (threadIdx.x < BCH) sb[threadIdx.x] = __ldg(p.b + threadIdx.x);
  __syncthreads();
  constexpr int MP = MELS / 2 + 2;
  const long long pos = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
  const int cg = threadIdx.x & 7;
  if (pos >= (long long)p.n_steps * MP) return;
  const int s = (int)(pos / MP), mp = (int)(pos - (long long)s * MP);
  const int clip = __ldg(p.clip_of_step + s);
  const bool live = clip >= 0 && mp >= 1 && mp <= MELS / 2;
  T* orow = reinterpret_cast<T*>(p.out) + pos * BCH + cg * 8;
  if (p.mask_out && cg == 0) p.mask_out[pos] = live ? 1 : 0;
  float r[8];
  if (!live) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = 0.f;
    Row8<T>::store(orow, r);
    return;
  }
  const int t = __ldg(p.t_of_step + s), m = mp - 1;
  const int fo = __ldg(p.clip_frame_off + clip), nf = __ldg(p.clip_frame_off + clip + 1) - fo;
  float x[4][4];                                               // [frame 2t-1 .. 2t+2][mel 2m-1 .. 2m+2]
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int f = 2 * t - 1 + a;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int mm = 2 * m - 1 + c;
      x[a][c] = (f >= 0 && f < nf && mm >= 0 && mm < MELS) ? __ldg(p.lms + (size_t)(fo + f) * MELS + mm) : 0.f;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = cg * 8 + j;
    const float* w = sw + ch * 9;                              // [dm][df] as Conv2d(mel, time) stores it
    float best = -INFINITY;
#pragma unroll
    for (int df = 0; df < 2; ++df)
#pragma unroll
      for (int dm = 0; dm < 2; ++dm) {
        float acc = sb[ch];
#pragma unroll
        for (int km = 0; km < 3; ++km)
#pragma unroll
          for (int kf = 0; kf < 3; ++kf) acc = fmaf(w[km * 3 + kf], x[df + kf][dm + km], acc);
        best = fmaxf(best, acc);
      }
    r[j] = fmaxf(best, 0.f);
  }
  Row8<T>::store(orow, r);
}

struct PoolParams {
  const void* in; int mel_in; const int* clip_step_in; const int* clip_of_step; const int* t_of_step; int n_steps_out, pad_out;
  void* out; uint8_t* mask_out;
};

// maxpool 2x2 from one grid layout into the next (pad_out = 1) or into the dense [step, mel * 64 + channel] rows the fc
// layers read (pad_out = 0): thread = (output position, 8 channels). The inputs are post-ReLU, the order of ReLU and max
// does not matter.
template <typename T>
__global__ void __launch_bounds__(256) byola_pool_kernel(const PoolParams p) {
  const int mo = p.mel_in / 2, mpo = mo + 2 * p.pad_out, mpi = p.mel_in + 2;
  const long long pos = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
  const int cg = threadIdx.x & 7;
  if (pos >= (long long)p.n_steps_out * mpo) return;
  const int s = (int)(pos / mpo), mp = (int)(pos - (long long)s * mpo);
  const int clip = __ldg(p.clip_of_step + s);
  const bool live = clip >= 0 && mp >= p.pad_out && mp < mo + p.pad_out;
  T* orow = reinterpret_cast<T*>(p.out) + pos * BCH + cg * 8;
  if (p.mask_out && cg == 0) p.mask_out[pos] = live ? 1 : 0;
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = 0.f;
  if (live) {
    const int t = __ldg(p.t_of_step + s), m = mp - p.pad_out;
    const long long row0 = ((long long)__ldg(p.clip_step_in + clip) + 2 * t) * mpi + 2 * m + 1;
    const T* in = reinterpret_cast<const T*>(p.in) + cg * 8;
    float v[8];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        Row8<T>::load(in + (row0 + (long long)a * mpi + c) * BCH, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = fmaxf(r[j], v[j]);
      }
  }
  Row8<T>::store(orow, r);
}

}  // namespace avdf

using namespace avdf;

extern "C" int avdf_logmel(const float* wav, const int64_t* clip_sample_off, const int32_t* clip_frame_off,
                           const int32_t* clip_pair_off, int32_t n_clips, int32_t total_frames, int32_t total_pairs, const float* window, const float* twiddle, const int32_t* mel_lo, const int32_t* mel_cnt, const float* mel_w,
                           int32_t mel_stride, float mean, float std, float* lms, void* stream) {
  AVDF_CHECK_ARG(wav && clip_sample_off && clip_frame_off && clip_pair_off && window && twiddle && mel_lo && mel_cnt && mel_w && lms, "null pointer");
  AVDF_CHECK_ARG(n_clips >= 1 && total_frames >= 0 && total_pairs >= 0 && total_pairs <= total_frames && std > 0.f && mel_stride >= 1, "bad sizes");
  AVDF_CHECK_ARG((reinterpret_cast<uintptr_t>(twiddle) & 7) == 0, "twiddle table must be 8-byte aligned");
  if (total_frames == 0) return AVDF_OK;
  LogMelParams p{wav, reinterpret_cast<const long long*>(clip_sample_off), clip_frame_off, clip_pair_off, n_clips, total_frames, window,
                 reinterpret_cast<const float2*>(twiddle), mel_lo, mel_cnt, mel_w, mel_stride, mean, 1.0f / std, 1.1920928955078125e-07f, lms};
  logmel_kernel<<<total_pairs, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("logmel_kernel");
}

extern "C" int avdf_byola_conv1_pool(const float* lms, const int32_t* clip_frame_off, const float* w, const float* b,
                                     const int32_t* clip_of_step, const int32_t* t_of_step, int32_t n_steps, void* out,
                                     int32_t out_dtype, uint8_t* mask_out, void* stream) {
  AVDF_CHECK_ARG(lms && clip_frame_off && w && b && clip_of_step && t_of_step && out, "null pointer");
  AVDF_CHECK_ARG(n_steps >= 0, "bad n_steps");
  AVDF_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 31) == 0, "out must be 32-byte aligned");
  if (n_steps == 0) return AVDF_OK;
  Conv1Params p{lms, clip_frame_off, w, b, clip_of_step, t_of_step, n_steps, out, mask_out};
  const long long positions = (long long)n_steps * (MELS / 2 + 2);
  const int grid = (int)((positions + 31) / 32);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (out_dtype == AVDF_DTYPE_F32) byola_conv1_pool_kernel<float><<<grid, 256, 0, st>>>(p);
  else if (out_dtype == AVDF_DTYPE_BF16) byola_conv1_pool_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else if (out_dtype == AVDF_DTYPE_F16) byola_conv1_pool_kernel<__half><<<grid, 256, 0, st>>>(p);
  else AVDF_CHECK_ARG(false, "out_dtype must be F32, BF16 or F16");
  return check_launch("byola_conv1_pool_kernel");
}

extern "C" int avdf_byola_pool(const void* in, int32_t dtype, int32_t mel_in, const int32_t* clip_step_in,
                               const int32_t* clip_of_step, const int32_t* t_of_step, int32_t n_steps_out, int32_t pad_out,
                               void* out, uint8_t* mask_out, void* stream) {
  AVDF_CHECK_ARG(in && clip_step_in && clip_of_step && t_of_step && out, "null pointer");
  AVDF_CHECK_ARG(mel_in >= 2 && mel_in % 2 == 0 && n_steps_out >= 0 && (pad_out == 0 || pad_out == 1), "bad sizes");
  AVDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(in)) & 31) == 0, "in / out must be 32-byte aligned");
  if (n_steps_out == 0) return AVDF_OK;
  PoolParams p{in, mel_in, clip_step_in, clip_of_step, t_of_step, n_steps_out, pad_out, out, mask_out};
  const long long positions = (long long)n_steps_out * (mel_in / 2 + 2 * pad_out);
  const int grid = (int)((positions + 31) / 32);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == AVDF_DTYPE_F32) byola_pool_kernel<float><<<grid, 256, 0, st>>>(p);
  else if (dtype == AVDF_DTYPE_BF16) byola_pool_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else if (dtype == AVDF_DTYPE_F16) byola_pool_kernel<__half><<<grid, 256, 0, st>>>(p);
  else AVDF_CHECK_ARG(false, "dtype must be F32, BF16 or F16");
  return check_launch("byola_pool_kernel");
}
