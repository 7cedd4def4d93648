// Banded (window 7) softmax attention of LocalMaskedMHCA / LocalMaskedMMHCA (libs/modeling/blocks.py:977-1224, 535-781) on
// the tensor cores for 16-bit q / k / v: 4 heads x 64 channels, one warp per 16 consecutive query rows.
//
// The CUDA-core kernel (blocks.cu: attention_banded_kernel, warp per query row) executes ~460 instructions per row, most of
// them 16-bit -> fp32 conversions, the 2 x 8 FMAs per (row, key) and the shuffle reductions of the per-head dot products.
// Here a warp treats its 16 query rows as a dense 16 x 32 problem per head - keys [i0 - 8, i0 + 24) cover the band
// |key - row| <= 3 of every row - and masks the band afterwards:
//     S = Q K^T    mma.sync m16n8k16 (fp32 accumulate): 4 key tiles x 4 channel steps per head
//     softmax      in the accumulator layout (a row lives in 4 lanes: two shuffles per reduction), base-2, the reference's
//                  additive -1e4 on masked keys and -inf outside the sequence / band
//     O = P V      P re-used as the A operand straight from the accumulators, split into a 16-bit head and a 16-bit
//                  remainder (two MMAs) so that the probabilities keep fp32-level accuracy; V fragments by ldmatrix.trans
// 4.6x more multiply-adds than the band needs, ~7x fewer instructions. A CTA owns 64 (or 32) query rows and ONE PAIR of heads
// (blockIdx.y): the 128 channels of its K / V rows (+16 halo rows) are staged once by cp.async with a 272-byte row pitch
// (conflict-free ldmatrix; 43 KB per CTA, so 4-5 CTAs are resident per SM); rows outside the sequence are zero-filled; the
// Q fragments are fetched into registers while the K / V rows are in flight.
// Measured (batch 32, CUDA-graph replay, scripts/attention_bench.py): T = 768 / 384 / 192 / 96 / 48: 14.9 / 8.4 / 5.9 / 4.9 /
// 4.4 us against 19.3 / 11.9 / 8.3 / 7.0 / 6.9 us of the CUDA-core kernel.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"

namespace avdf {

constexpr int kC = 256;               // channels (4 heads x 64)
constexpr int AM_HEADS = 2;               // heads per CTA (blockIdx.y selects the pair): 128 channels of K / V staged per row
constexpr int AM_PITCH = AM_HEADS * 64 * 2 + 16;   // bytes per staged row (272: consecutive rows 4 banks apart)

template <bool BF16>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// p -> 16-bit head and 16-bit remainder, two values per register
template <bool BF16> __device__ __forceinline__ void split2(float p0, float p1, uint32_t& hi, uint32_t& lo) {
  if (BF16) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(p0), h1 = __float2bfloat16_rn(p1);
    hi = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
    lo = pack_bf16x2(p0 - __bfloat162float(h0), p1 - __bfloat162float(h1));
  } else {
    const __half h0 = __float2half_rn(p0), h1 = __float2half_rn(p1);
    hi = pack_f16x2(__half2float(h0), __half2float(h1));
    lo = pack_f16x2(p0 - __half2float(h0), p1 - __half2float(h1));
  }
}
__device__ __forceinline__ void store2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store2(__half* p, float a, float b) { *reinterpret_cast<uint32_t*>(p) = pack_f16x2(a, b); }
__device__ __forceinline__ void store2(__nv_bfloat16* p, float a, float b) { *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(a, b); }

template <typename InT, typename OutT, int AM_WARPS>
__global__ void __launch_bounds__(AM_WARPS * 32) attention_banded_mma_kernel(const InT* __restrict__ q, const InT* __restrict__ k,
                                                                            const InT* __restrict__ v, const unsigned char* __restrict__ kv_mask,
                                                                            OutT* __restrict__ out, int B, int T, int RPV) {
  pdl_trigger();                 // the output-projection GEMM may start its prologue now (it waits for this grid)
  pdl_wait();                    // launched as a programmatic dependent of the q/k/v GEMM: its CTAs are placed while that grid drains
  constexpr bool BF16 = std::is_same<InT, __nv_bfloat16>::value;
  constexpr int HALF = 3;
  constexpr int AM_ROWS = 16 * AM_WARPS, AM_KEYS = AM_ROWS + 16;
  extern __shared__ __align__(16) unsigned char am_smem[];
  unsigned char* ks = am_smem;
  unsigned char* vs = am_smem + AM_KEYS * AM_PITCH;
  __shared__ unsigned char s_mask[AM_KEYS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_per_video = (T + AM_ROWS - 1) / AM_ROWS;
  const int b = blockIdx.x / tiles_per_video;
  const int t0 = (blockIdx.x - b * tiles_per_video) * AM_ROWS;
  const int lo = t0 - 8;                                       // staged row r holds key lo + r
  const size_t qb = (size_t)b * RPV, base = (size_t)b * T;
  const int cb = blockIdx.y * AM_HEADS * 64;                   // first channel of this CTA's heads
  constexpr int CHUNKS = AM_HEADS * 64 * 2 / 16;               // 16-byte chunks per staged row
  for (int i = threadIdx.x; i < AM_KEYS * CHUNKS; i += AM_WARPS * 32) {
    const int r = i / CHUNKS, c = i - r * CHUNKS;
    const int j = lo + r;
    unsigned char* dk = ks + r * AM_PITCH + c * 16;
    unsigned char* dv = vs + r * AM_PITCH + c * 16;
    if (j >= 0 && j < T) {
      const size_t g = ((qb + j) * kC + cb) * sizeof(InT) + (size_t)c * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dk)), "l"(reinterpret_cast<const unsigned char*>(k) + g) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dv)), "l"(reinterpret_cast<const unsigned char*>(v) + g) : "memory");
    } else {                                                   // 0 x garbage must not reach the accumulators
      *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
    }
  }
  for (int i = threadIdx.x; i < AM_KEYS; i += AM_WARPS * 32) {
    const int j = lo + i;
    s_mask[i] = (j >= 0 && j < T) ? (kv_mask ? kv_mask[base + j] : 1) : 0;
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  const int i0 = t0 + 16 * warp;                               // this warp's first query row
  const int g = lane >> 2, t = lane & 3;
  const int r0 = i0 + g, r1 = r0 + 8;
  const bool ok0 = r0 < T, ok1 = r1 < T;
  // the A fragments of Q for the CTA's heads (rows r0, r1; 32 registers), fetched while the K / V rows are in flight - loaded
  // head by head their ~1 us of latency was exposed four times per warp
  uint32_t qa[AM_HEADS][4][4];
  {
    const uint32_t* q0 = reinterpret_cast<const uint32_t*>(q + (qb + (ok0 ? r0 : 0)) * kC);
    const uint32_t* q1 = reinterpret_cast<const uint32_t*>(q + (qb + (ok1 ? r1 : 0)) * kC);
#pragma unroll
    for (int h = 0; h < AM_HEADS; ++h)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int w32 = (cb + h * 64 + kk * 16) / 2 + t;       // 32-bit word of channels (2t, 2t + 1) of this step
        qa[h][kk][0] = ok0 ? __ldg(q0 + w32) : 0u; qa[h][kk][1] = ok1 ? __ldg(q1 + w32) : 0u;
        qa[h][kk][2] = ok0 ? __ldg(q0 + w32 + 4) : 0u; qa[h][kk][3] = ok1 ? __ldg(q1 + w32 + 4) : 0u;
      }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (i0 >= T) return;
  const int kl0 = 16 * warp;                                   // staged row of key i0 - 8
  const int mi = lane >> 3, mr = lane & 7;                     // ldmatrix: this lane addresses row mr of matrix mi
  const float scale = 0.125f * 1.4426950408889634f;            // 1/sqrt(64) * log2(e)
  const float masked = -1e4f * 1.4426950408889634f;            // blocks.py:1194-1195, base-2 domain
  const uint32_t ks_u = (uint32_t)__cvta_generic_to_shared(ks), vs_u = (uint32_t)__cvta_generic_to_shared(vs);
#pragma unroll
  for (int h = 0; h < AM_HEADS; ++h) {                         // h: head within the CTA's pair (staged columns h * 64 ..)
    float s[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {                           // 16 channels of the head per step
      const uint32_t (&a)[4] = qa[h][kk];
#pragma unroll
      for (int np = 0; np < 2; ++np) {                         // key tiles 2 np, 2 np + 1
        uint32_t bf[4];
        ldmatrix_x4(bf, ks_u + (uint32_t)((kl0 + (2 * np + (mi >> 1)) * 8 + mr) * AM_PITCH + (h * 64 + kk * 16 + (mi & 1) * 8) * 2));
        mma16816<BF16>(s[2 * np], a, bf[0], bf[1]);
        mma16816<BF16>(s[2 * np + 1], a, bf[2], bf[3]);
      }
    }
    // ---- band / sequence / key mask, base-2 softmax of rows r0 (s[.][0..1]) and r1 (s[.][2..3])
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kl = kl0 + j * 8 + 2 * t + e;                // staged row of this key
        const int key = lo + kl;
        const bool inseq = key >= 0 && key < T;
        const float add = s_mask[kl] ? 0.f : masked;
        const int d0 = key - r0, d1 = key - r1;
        const float x0 = (inseq && d0 >= -HALF && d0 <= HALF) ? fmaf(s[j][e], scale, add) : -INFINITY;
        const float x1 = (inseq && d1 >= -HALF && d1 <= HALF) ? fmaf(s[j][2 + e], scale, add) : -INFINITY;
        s[j][e] = x0; s[j][2 + e] = x1;
        m0 = fmaxf(m0, x0); m1 = fmaxf(m1, x1);
      }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    if (m0 == -INFINITY) m0 = 0.f;                              // (a row beyond the sequence: every p becomes 0)
    if (m1 == -INFINITY) m1 = 0.f;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float p0, p1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(s[j][e] - m0));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(s[j][2 + e] - m1));
        s[j][e] = p0; s[j][2 + e] = p1;
        l0 += p0; l1 += p1;
      }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // ---- O = P V: the accumulator tiles 2 kk, 2 kk + 1 are the A fragment of key step kk
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t ah[4], al[4];
      split2<BF16>(s[2 * kk][0], s[2 * kk][1], ah[0], al[0]);
      split2<BF16>(s[2 * kk][2], s[2 * kk][3], ah[1], al[1]);
      split2<BF16>(s[2 * kk + 1][0], s[2 * kk + 1][1], ah[2], al[2]);
      split2<BF16>(s[2 * kk + 1][2], s[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {                         // channel tiles 2 np, 2 np + 1 of the head
        uint32_t bf[4];
        ldmatrix_x4_trans(bf, vs_u + (uint32_t)((kl0 + kk * 16 + (mi & 1) * 8 + mr) * AM_PITCH + (h * 64 + (2 * np + (mi >> 1)) * 8) * 2));
        mma16816<BF16>(o[2 * np], ah, bf[0], bf[1]);
        mma16816<BF16>(o[2 * np], al, bf[0], bf[1]);
        mma16816<BF16>(o[2 * np + 1], ah, bf[2], bf[3]);
        mma16816<BF16>(o[2 * np + 1], al, bf[2], bf[3]);
      }
    }
    const float inv0 = (s_mask[kl0 + 8 + g] && l0 > 0.f) ? 1.f / l0 : 0.f;       // blocks.py:1208-1209
    const float inv1 = (s_mask[kl0 + 16 + g] && l1 > 0.f) ? 1.f / l1 : 0.f;
    OutT* o0 = out + (base + r0) * kC + cb + h * 64 + 2 * t;
    OutT* o1 = out + (base + r1) * kC + cb + h * 64 + 2 * t;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      if (ok0) store2(o0 + n * 8, o[n][0] * inv0, o[n][1] * inv0);
      if (ok1) store2(o1 + n * 8, o[n][2] * inv1, o[n][3] * inv1);
    }
  }
}

// 16-bit q / k / v, window 7, 4 heads x 64 channels; returns AVDF_ERR_UNSUPPORTED for anything else
int attention_banded_mma(const void* q, const void* k, const void* v, const unsigned char* kv_mask, void* out, int in_dtype,
                         int out_dtype, int batch, int t, int rpv, cudaStream_t st) {
  if (in_dtype != AVDF_DTYPE_F16 && in_dtype != AVDF_DTYPE_BF16) return AVDF_ERR_UNSUPPORTED;
  static const int warps_env = getenv("AVDF_ATT_WARPS") ? atoi(getenv("AVDF_ATT_WARPS")) : 0;
  // 64 query rows per CTA while that still gives every SM two CTAs, else 32 (measured, batch 32: T = 768 14.9 vs 15.4 us,
  // T = 192 6.6 vs 5.9 us)
  const int sms = device_sm_count();
  const int warps = warps_env == 4 || warps_env == 2 || warps_env == 1 ? warps_env
                                                                         : (batch * ((t + 63) / 64) * (4 / AM_HEADS) >= 2 * sms ? 4 : 2);
  const int rows = 16 * warps;
  const dim3 grid(batch * ((t + rows - 1) / rows), 4 / AM_HEADS);
  const int smem = 2 * (rows + 16) * AM_PITCH;
#define AVDF_AM(InT) do { if (warps == 4) AVDF_AM_W(InT, 4); else if (warps == 1) AVDF_AM_W(InT, 1); else AVDF_AM_W(InT, 2); } while (0)
#define AVDF_AM_W(InT, W)                                                                                                  \
  AVDF_DISPATCH_DTYPE(out_dtype, OutT, {                                                                                   \
    AVDF_SMEM_ATTR_ONCE((attention_banded_mma_kernel<InT, OutT, W>), 2 * (64 + 16) * AM_PITCH);                            \
    cudaError_t le_ = launch_pdl(attention_banded_mma_kernel<InT, OutT, W>, grid, W * 32, smem, st, (const InT*)q, (const InT*)k,   \
                                 (const InT*)v, kv_mask, (OutT*)out, batch, t, rpv);                                       \
    if (le_ != cudaSuccess) { set_error("attention_banded_mma_kernel: launch failed: %s", cudaGetErrorString(le_)); return AVDF_ERR_CUDA; } \
  })
  if (in_dtype == AVDF_DTYPE_F16) AVDF_AM(__half); else AVDF_AM(__nv_bfloat16);
#undef AVDF_AM
#undef AVDF_AM_W
  return check_launch("attention_banded_mma_kernel");
}

}  // namespace avdf
