// HBM-bound kernels of the ConvTransformer blocks, the FPN and the head / video-level tails.
// All activations are token-major [batch, T, C] with C contiguous; one warp owns one row of
// C = 256 channels (lane l holds channels 8l .. 8l+7: a 1 KB fp32 row is one fully coalesced
// 2 x 512 B request, a bf16 row one 512 B request), row statistics are warp shuffles.
//
// Reference code restated (paths relative to the reference root):
//   ln_dwconv_ln     LayerNorm (libs/modeling/blocks.py:97-112) -> depthwise MaskedConv1D k3
//                    (blocks.py:41-63; q/k/v convs of LocalMaskedMHCA :1159-1165, LocalMaskedMMHCA
//                    :726-741, MaskedMHCA :291-297) -> LayerNorm; the nearest up/down-sampling between
//                    scales (libs/modeling/backbones.py:487,490) is an index map on the source rows,
//                    and the stride-2 skip MaxPool1d(3,2,1) (blocks.py:1277-1281) is a by-product.
//   attention        banded softmax attention == the sliding-chunk code of blocks.py:977-1224
//                    (verified to 1.8e-7) and the global MaskedMHCA blocks.py:299-309
//   ln_rows          blocks.py:97-112 (ln2 before the MLP, blocks.py:1311)
//   instnorm_lrelu   nn.InstanceNorm1d + LeakyReLU(0.2) of DownBlock, blocks.py:1508-1515
//   fpn_fuse         FPN1D.forward top-down path + fpn_convs + fpn_norms, libs/modeling/necks.py:75-93
//   head_final       cls_head / offset_head + Scale + ReLU, libs/modeling/av_fd_no_recon.py:82-89,152-159
//   vcls_exp12       DeepInterpolator.classifier, blocks.py:1608-1626
//   vcls_exp13       SegmentandCls.segment, blocks.py:1682-1700
#include <math.h>
#include <type_traits>
#include "common.cuh"

namespace avdf {

constexpr float kLnEps = 1e-5f;
constexpr int kC = 256;               // channels handled by the warp-per-row kernels

// mean / 1/sqrt(var+eps) of one 256-channel row spread over a warp (8 values per lane), two-pass
__device__ __forceinline__ void row_stats(const float (&v)[8], float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += v[k];
  mean = warp_sum(s) * (1.f / kC);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { const float d = v[k] - mean; q = fmaf(d, d, q); }
  rstd = 1.f / sqrtf(warp_sum(q) * (1.f / kC) + kLnEps);
}

__device__ __forceinline__ void lds8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ------------------------------------------------------------------------------------------------
// LN -> depthwise conv k3 (stride 1|2) * mask -> LN for up to 3 streams sharing one source
// ------------------------------------------------------------------------------------------------
struct LdlParams {
  const float* src; const unsigned char* mask_out;
  const float* ln_in_w[3]; const float* ln_in_b[3]; const float* dw_w[3];
  const float* ln_out_w[3]; const float* ln_out_b[3];
  void* out[3]; float* skip_out;
  int B, t_src, t_virt, shift, t_out, n_streams, out_rows, rows;
};
constexpr int LDL_MAX_ROWS = 8;        // output rows per warp tile: 8, 4 or 2 (host picks the largest that still fills the SMs)
constexpr int LDL_WARPS = 8;
constexpr int LDL_DEPTH = 4;           // source rows in flight per warp (cp.async ring): 24 warps x 3 KB per SM keeps HBM busy

// A lane owns channels [4 lane, 4 lane + 4) and [128 + 4 lane, 128 + 4 lane + 4): every 16-byte shared-memory access of a
// warp then covers 512 contiguous bytes (conflict-free LDS.128; 8 contiguous channels per lane put lanes i and i + 4
// on the same banks and doubled the wavefronts of what ncu showed to be the limiting pipe).
__device__ __forceinline__ void lds8p(const float* row, int lane, f32x2 (&v)[4]) {       // 8 floats as 4 packed pairs
  const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(row + 4 * lane);
  const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(row + 128 + 4 * lane);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ float hsum2(f32x2 v) { float a, b; upk2(v, a, b); return a + b; }
__device__ __forceinline__ void store4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
__device__ __forceinline__ void store4(__half* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_f16x2(a, b), pack_f16x2(c, d));
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}
template <typename OutT> __device__ __forceinline__ void store8p(OutT* row, int lane, const f32x2 (&v)[4]) {
  float f[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) upk2(v[k], f[2 * k], f[2 * k + 1]);
  store4(row + 4 * lane, f[0], f[1], f[2], f[3]);
  store4(row + 128 + 4 * lane, f[4], f[5], f[6], f[7]);
}

// With u = xhat * w_in + b_in (xhat = the normalised source row, shared by all streams) the depthwise conv is
//   acc[c] = sum_j dw[c][j] * u_j[c] = sum_j (dw[c][j] w_in[c]) * xhat_j[c] + b_in[c] * sum_j dw[c][j]
// so the per-stream constants A_j = dw_j * w_in, Bj = dw_j * b_in, Bsum = sum_j Bj are folded once per CTA into shared
// memory and an interior output row costs 3 FMA per channel and stream (edge rows, where a tap falls outside
// [0, t_virt), use the per-tap Bj instead of Bsum).
// The kernel is bound by issue slots (ncu: ~600 instructions per output row before this layout), so all the
// element-wise math runs on packed fp32 pairs (fma.rn.f32x2: two channels per instruction), each LayerNorm carries
// its (sum, sum of squares) as ONE packed pair through the 5 butterfly steps, and 1/sqrt is the MUFU rsqrt (2 ulp).
// The NS streams of an output row are computed together so that their reductions overlap.
template <typename OutT, int STRIDE, int NS>
__global__ void __launch_bounds__(LDL_WARPS * 32, 3) ln_dwconv_ln_kernel(const LdlParams p) {
  // per stream: A0, A1, A2, Bsum, B0, B1, B2, ln_out_w, ln_out_b  (9 x 256 floats)
  pdl_trigger();                 // the projection GEMM that follows may start its prologue now (it waits for this grid)
  extern __shared__ __align__(16) float ldl_smem[];
  float (*sp)[9][kC] = reinterpret_cast<float (*)[9][kC]>(ldl_smem);                                   // [NS][9][256]
  float (*ring)[LDL_DEPTH][kC] = reinterpret_cast<float (*)[LDL_DEPTH][kC]>(ldl_smem + NS * 9 * kC);   // [warps][depth][256]
  for (int i = threadIdx.x; i < NS * kC; i += blockDim.x) {
    const int s = i / kC, c = i - s * kC;
    const float w = p.ln_in_w[s][c], bb = p.ln_in_b[s][c];
    const float d0 = p.dw_w[s][3 * c], d1 = p.dw_w[s][3 * c + 1], d2 = p.dw_w[s][3 * c + 2];
    sp[s][0][c] = d0 * w; sp[s][1][c] = d1 * w; sp[s][2][c] = d2 * w;
    sp[s][4][c] = d0 * bb; sp[s][5][c] = d1 * bb; sp[s][6][c] = d2 * bb;
    sp[s][3][c] = d0 * bb + d1 * bb + d2 * bb;
    sp[s][7][c] = p.ln_out_w[s][c]; sp[s][8][c] = p.ln_out_b[s][c];
  }
  pdl_wait();                    // launched as a programmatic dependent: everything above (weights only) overlapped the previous kernel's tail
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int R = p.rows;
  const int tiles_per_video = (p.t_out + R - 1) / R;
  const long long n_tiles = (long long)p.B * tiles_per_video;
  const f32x2 zero2 = pk2(0.f);
  for (long long tile = (long long)blockIdx.x * LDL_WARPS + warp; tile < n_tiles; tile += (long long)gridDim.x * LDL_WARPS) {
    const int b = (int)(tile / tiles_per_video);
    const int t0 = (int)(tile - (long long)b * tiles_per_video) * R;
    const int t1 = min(t0 + R, p.t_out);
    const float* src_b = p.src + (size_t)b * p.t_src * kC;
    f32x2 xh[3][4];                 // normalised rows of the 3-position window
    f32x2 rw[STRIDE == 2 ? 3 : 1][4];   // raw rows (only the stride-2 MaxPool skip needs them)
    bool ok[3] = {false, false, false};
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) xh[j][k] = zero2;
    // the tile's mask bytes, one per lane, fetched once (a per-row byte load sat on the critical path of every row)
    unsigned tile_mask = 0xffffffffu;
    if (p.mask_out) {
      const bool mbit = (t0 + lane < t1) ? (p.mask_out[(size_t)b * p.t_out + t0 + lane] != 0) : true;
      tile_mask = __ballot_sync(0xffffffffu, mbit);
    }
    // software pipeline: the source rows of positions pos + 1 .. pos + LDL_DEPTH - 1 are in flight (cp.async into a
    // per-warp smem ring; every lane copies and later reads its own 32 bytes, so no cross-lane synchronisation)
    const int pos_first = STRIDE * t0 - 1, pos_last = STRIDE * (t1 - 1) + 1;
    const unsigned ring_base = (unsigned)__cvta_generic_to_shared(&ring[warp][0][4 * lane]);
    auto issue = [&](int pos) {
      if (pos >= 0 && pos < p.t_virt && pos <= pos_last) {
        const int r = p.shift >= 0 ? (pos >> p.shift) : (pos << (-p.shift));
        const float* g = src_b + (size_t)r * kC + 4 * lane;
        const unsigned d = ring_base + (unsigned)(((pos - pos_first) % LDL_DEPTH) * kC * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 512), "l"(g + 128) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int d = 0; d < LDL_DEPTH; ++d) issue(pos_first + d);
    for (int pos = pos_first; pos <= pos_last; ++pos) {
      asm volatile("cp.async.wait_group %0;" ::"n"(LDL_DEPTH - 1) : "memory");
      // slide the window, normalise the new row
#pragma unroll
      for (int k = 0; k < 4; ++k) { xh[0][k] = xh[1][k]; xh[1][k] = xh[2][k]; }
      ok[0] = ok[1]; ok[1] = ok[2];
      ok[2] = pos >= 0 && pos < p.t_virt;
      f32x2 nxt[4];
      if (ok[2]) lds8p(ring[warp][(pos - pos_first) % LDL_DEPTH], lane, nxt);
      else {
#pragma unroll
        for (int k = 0; k < 4; ++k) nxt[k] = zero2;
      }
      if (STRIDE == 2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { rw[0][k] = rw[STRIDE == 2 ? 1 : 0][k]; rw[STRIDE == 2 ? 1 : 0][k] = rw[STRIDE == 2 ? 2 : 0][k]; rw[STRIDE == 2 ? 2 : 0][k] = nxt[k]; }
      }
      if (ok[2]) {                                // input LayerNorm: shifted single pass - sum and sum of squares of
        float n0, n1;                             // (x - pivot) travel as one packed pair through the 5 shuffle steps
        upk2(nxt[0], n0, n1);
        const float pivot = __shfl_sync(0xffffffffu, n0, 0);
        const f32x2 npiv = pk2(-pivot);
        f32x2 s2 = zero2, q2 = zero2;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const f32x2 d = add2(nxt[k], npiv); s2 = add2(s2, d); q2 = fma2(d, d, q2); }
        f32x2 st = pk2(hsum2(s2), hsum2(q2));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) st = add2(st, __shfl_xor_sync(0xffffffffu, st, o));
        float su, sq;
        upk2(st, su, sq);
        const float dm = su * (1.f / kC);
        const float rstd = rsqrtf(fmaxf(sq * (1.f / kC) - dm * dm, 0.f) + kLnEps);
        const f32x2 nmean = pk2(-(pivot + dm)), rs2 = pk2(rstd);
#pragma unroll
        for (int k = 0; k < 4; ++k) xh[2][k] = mul2(add2(nxt[k], nmean), rs2);
      }
      issue(pos + LDL_DEPTH);                    // refill the slot just consumed
      const int rel = pos - (STRIDE * t0 + 1);
      if (rel < 0 || (rel % STRIDE) != 0) continue;
      const int t = t0 + rel / STRIDE;          // output row whose taps are window[0..2]
      const size_t orow = (size_t)b * p.t_out + t;              // dense row: mask / skip
      const size_t srow = (size_t)b * p.out_rows + t;           // row in the (possibly interleaved) output buffers
      const bool keep = (tile_mask >> (t - t0)) & 1u;
      const bool interior = ok[0] && ok[2];                      // ok[1] always holds for an output row
      f32x2 acc[NS][4];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (!keep) {
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[s][k] = zero2;
        } else if (interior) {
          f32x2 a0[4], a1[4], a2[4];
          lds8p(sp[s][3], lane, acc[s]); lds8p(sp[s][0], lane, a0); lds8p(sp[s][1], lane, a1); lds8p(sp[s][2], lane, a2);
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[s][k] = fma2(a2[k], xh[2][k], fma2(a1[k], xh[1][k], fma2(a0[k], xh[0][k], acc[s][k])));
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[s][k] = zero2;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (ok[j]) {
              f32x2 a[4], bj[4];
              lds8p(sp[s][j], lane, a); lds8p(sp[s][4 + j], lane, bj);
#pragma unroll
              for (int k = 0; k < 4; ++k) acc[s][k] = add2(acc[s][k], fma2(a[k], xh[j][k], bj[k]));
            }
          }
        }
      }
      f32x2 st[NS];                               // (sum, sum of squares) per stream
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const f32x2 s2 = add2(add2(acc[s][0], acc[s][1]), add2(acc[s][2], acc[s][3]));
        f32x2 q2 = mul2(acc[s][0], acc[s][0]);
#pragma unroll
        for (int k = 1; k < 4; ++k) q2 = fma2(acc[s][k], acc[s][k], q2);
        st[s] = pk2(hsum2(s2), hsum2(q2));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int s = 0; s < NS; ++s) st[s] = add2(st[s], __shfl_xor_sync(0xffffffffu, st[s], o));
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        float su, sq;
        upk2(st[s], su, sq);
        const float m2 = su * (1.f / kC);
        const float r2 = rsqrtf(fmaxf(sq * (1.f / kC) - m2 * m2, 0.f) + kLnEps);
        const f32x2 rr = pk2(r2), mm = pk2(-m2 * r2);
        f32x2 w[4], bb[4];
        lds8p(sp[s][7], lane, w); lds8p(sp[s][8], lane, bb);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[s][k] = fma2(fma2(acc[s][k], rr, mm), w[k], bb[k]);
        store8p(reinterpret_cast<OutT*>(p.out[s]) + srow * kC, lane, acc[s]);
      }
      if (STRIDE == 2 && p.skip_out) {           // MaxPool1d(3, 2, 1) of the raw rows, -inf padding
        float mx[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float m0, m1, l0, l1, h0, h1;
          upk2(rw[STRIDE == 2 ? 1 : 0][k], m0, m1);
          upk2(rw[0][k], l0, l1);
          upk2(rw[STRIDE == 2 ? 2 : 0][k], h0, h1);
          if (ok[0]) { m0 = fmaxf(m0, l0); m1 = fmaxf(m1, l1); }
          if (ok[2]) { m0 = fmaxf(m0, h0); m1 = fmaxf(m1, h1); }
          mx[2 * k] = m0; mx[2 * k + 1] = m1;
        }
        store4(p.skip_out + orow * kC + 4 * lane, mx[0], mx[1], mx[2], mx[3]);
        store4(p.skip_out + orow * kC + 128 + 4 * lane, mx[4], mx[5], mx[6], mx[7]);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// The same operator, second layout (default). The kernel above runs one warp per output row with all three streams'
// folded constants in shared memory: 36 constant LDS.128 per row put it on the LSU data pipe (ncu: 61 % of peak, twice
// the issue-slot utilisation), and the input LayerNorm of a row is the head of every row's dependency chain. Here a CTA
// owns LDL2_ROWS consecutive output rows of one video and works in two phases:
//   A  all warps: input LayerNorm statistics of the tile's source rows (+ halo), ONCE per row; the normalised rows go to
//      shared memory (pre-affine: the per-stream affine is folded into the conv constants as before)
//   B  warp = (stream, slice of the tile's rows): the stream's folded constants live in REGISTERS for the whole slice
//      (A0, A1, A2, Bsum, ln_out w / b for the lane's 8 channels: 48 registers), so a row costs 2 LDS.128 (the new
//      window row) + 12 FMA2 + the LayerNorm reduction + 4 FMA2 + the 16-byte stores. Edge rows (a tap outside the
//      sequence) fetch the per-tap bias terms from a small shared table.
// Same arithmetic and operation order per element as the first layout (1 / sqrt is the bare MUFU here: last-ulp differences).
// ------------------------------------------------------------------------------------------------
constexpr int LDL2_ROWS = 32;          // output rows per CTA
constexpr int LDL2_WARPS = 6;          // 1, 2 or 3 streams -> 6, 3 or 2 warps per stream
template <int STRIDE> __host__ __device__ constexpr int ldl2_src_rows() { return STRIDE * (LDL2_ROWS - 1) + 3; }

__device__ __forceinline__ float rsqrt_fast(float x) {      // MUFU.RSQ, 2 ulp; the argument is >= 1e-5 (no denormal path)
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// (Measured and dropped: a persistent form with the next tile's raw rows prefetched by cp.async into a shared-memory double
// buffer - 26.1 against 25.5 us per level-0 launch: an iteration takes as long as a whole one-tile CTA, the global loads were
// not what the warps wait for. The kernel runs at 0.33 instructions per cycle and scheduler with 3 warps per scheduler: every
// warp issues once in ~8 cycles - dependent FMA2 / shuffle / LDS chains - and the register file (150 per thread) caps both the
// warps per SM and the rows a warp can interleave: 2, 3 CTAs per SM x 2, 4 rows per iteration all land on 25-27 us.)
// Both phases are chains of dependent instructions per row (load -> sums -> 5 butterfly steps -> rsqrt -> scale), ~16
// cycles from issue to issue (ncu: 0.28 warp instructions per cycle and scheduler with 4.5 resident warps): every warp
// therefore works on LDL2_PA source rows (phase A) and LDL2_RPI<STRIDE> output rows (phase B) at once, their chains
// interleaved by the unrolled loops.
#ifndef AVDF_LDL2_PA
#define AVDF_LDL2_PA 6
#endif
constexpr int LDL2_PA = AVDF_LDL2_PA;      // (6: every warp's source rows of a stride-1 tile are one round of loads - one exposed latency instead of two)
#ifndef AVDF_LDL2_RPI1
#define AVDF_LDL2_RPI1 4
#endif
#ifndef AVDF_LDL2_MINB
#define AVDF_LDL2_MINB 2
#endif
template <int STRIDE> __host__ __device__ constexpr int ldl2_rpi() { return STRIDE == 1 ? AVDF_LDL2_RPI1 : 2; }

template <typename OutT, int STRIDE, int NS, bool SKIP>
__global__ void __launch_bounds__(LDL2_WARPS * 32, AVDF_LDL2_MINB) ln_dwconv_ln2_kernel(const LdlParams p) {
  pdl_trigger();                 // the projection GEMM that follows may start its prologue now (it waits for this grid)
  constexpr int SRC = ldl2_src_rows<STRIDE>();
  constexpr int RPI = ldl2_rpi<STRIDE>();
  constexpr int WIN = STRIDE * (RPI - 1) + 3;         // window rows of one iteration
  extern __shared__ __align__(16) float ldl2_smem[];
  float (*xh)[kC] = reinterpret_cast<float (*)[kC]>(ldl2_smem);                         // [SRC][256] normalised source rows
  float (*eb)[3][kC] = reinterpret_cast<float (*)[3][kC]>(ldl2_smem + SRC * kC);        // [NS][tap][256] dw_j * b_in
  float (*raw)[kC] = reinterpret_cast<float (*)[kC]>(ldl2_smem + (SRC + NS * 3) * kC);  // [SRC][256] raw rows (SKIP only)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_per_video = (p.t_out + LDL2_ROWS - 1) / LDL2_ROWS;
  const int b = blockIdx.x / tiles_per_video;
  const int t0 = (blockIdx.x - b * tiles_per_video) * LDL2_ROWS;
  const int t1 = min(t0 + LDL2_ROWS, p.t_out);
  const float* src_b = p.src + (size_t)b * p.t_src * kC;
  const int pos_first = STRIDE * t0 - 1, pos_last = STRIDE * (t1 - 1) + 1;
  const f32x2 zero2 = pk2(0.f);
  // the tile's mask bytes, one per lane
  unsigned tile_mask = 0xffffffffu;
  if (p.mask_out) {
    const bool mbit = (t0 + lane < t1) ? (p.mask_out[(size_t)b * p.t_out + t0 + lane] != 0) : true;
    tile_mask = __ballot_sync(0xffffffffu, mbit);
  }
  // Interior tiles (a full tile, no tap outside the sequence, no masked row - all but the first and last tile of a video
  // when the masks are all-true, i.e. always under force_upsampling) take the phase-B loop with every edge test compiled out:
  // the tests, the zero fills and the divergence bookkeeping they drag in were ~25 % of the executed instructions
  // (ncu source page: ISETP 8 %, CS2R 7 %, BRA / BSSY / BSYNC 7.6 %). Same arithmetic, same order: bit-identical results.
  // They also never read the per-tap bias table, so only edge tiles build it (768 x NS entries from global memory: ~1 us of
  // exposed latency at the head of every CTA).
  const bool interior = (t1 - t0) == LDL2_ROWS && pos_first >= 0 && pos_last < p.t_virt && tile_mask == 0xffffffffu;
  if (!interior) {
    for (int i = threadIdx.x; i < NS * kC; i += LDL2_WARPS * 32) {
      const int s = i / kC, c = i - s * kC;
      const float bb = p.ln_in_b[s][c];
      eb[s][0][c] = p.dw_w[s][3 * c] * bb; eb[s][1][c] = p.dw_w[s][3 * c + 1] * bb; eb[s][2][c] = p.dw_w[s][3 * c + 2] * bb;
    }
  }
  // the constants of this warp's phase-B stream, in registers for the whole slice: fetched HERE so that their latency overlaps
  // phase A's loads instead of following them
  constexpr int WPS = LDL2_WARPS / NS;          // warps per stream
  const int s = warp / WPS, part = warp - s * WPS;
  f32x2 A0[4], A1[4], A2[4], Bs[4], Wo[4], Bo[4];
  {
    const float* wi = p.ln_in_w[s]; const float* bi = p.ln_in_b[s]; const float* dw = p.dw_w[s];
    const float* wo = p.ln_out_w[s]; const float* bo = p.ln_out_b[s];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = (k < 2 ? 0 : 128) + 4 * lane + 2 * (k & 1);
      float a0[2], a1[2], a2[2], bs[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float w = __ldg(wi + c + e), bb = __ldg(bi + c + e);
        const float d0 = __ldg(dw + 3 * (c + e)), d1 = __ldg(dw + 3 * (c + e) + 1), d2 = __ldg(dw + 3 * (c + e) + 2);
        a0[e] = d0 * w; a1[e] = d1 * w; a2[e] = d2 * w; bs[e] = d0 * bb + d1 * bb + d2 * bb;
      }
      A0[k] = pk2(a0[0], a0[1]); A1[k] = pk2(a1[0], a1[1]); A2[k] = pk2(a2[0], a2[1]); Bs[k] = pk2(bs[0], bs[1]);
      Wo[k] = pk2(__ldg(wo + c), __ldg(wo + c + 1)); Bo[k] = pk2(__ldg(bo + c), __ldg(bo + c + 1));
    }
  }
  // launched as a programmatic dependent of the kernel that wrote `src`: the mask bytes, the bias table and the stream constants
  // above are weights / host-written tables and were fetched while that kernel was still draining
  pdl_wait();
  // ---- phase A: normalise every source position of the tile once
  for (int base = pos_first + warp; base <= pos_last; base += LDL2_WARPS * LDL2_PA) {
    f32x2 nxt[LDL2_PA][4];
    bool ok[LDL2_PA];
#pragma unroll
    for (int j = 0; j < LDL2_PA; ++j) {
      const int pos = base + j * LDL2_WARPS;
      ok[j] = pos <= pos_last && pos >= 0 && pos < p.t_virt;
      const int r = ok[j] ? (p.shift >= 0 ? (pos >> p.shift) : (pos << (-p.shift))) : 0;
      const float* g = src_b + (size_t)r * kC;
      const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2*>(g + 4 * lane));
      const ulonglong2 c = __ldg(reinterpret_cast<const ulonglong2*>(g + 128 + 4 * lane));
      nxt[j][0] = a.x; nxt[j][1] = a.y; nxt[j][2] = c.x; nxt[j][3] = c.y;
    }
    float pivot[LDL2_PA];
    f32x2 st[LDL2_PA];
#pragma unroll
    for (int j = 0; j < LDL2_PA; ++j) {
      float n0, n1;
      upk2(nxt[j][0], n0, n1);
      pivot[j] = __shfl_sync(0xffffffffu, n0, 0);
      const f32x2 npiv = pk2(-pivot[j]);
      f32x2 s2 = zero2, q2 = zero2;
#pragma unroll
      for (int k = 0; k < 4; ++k) { const f32x2 d = add2(nxt[j][k], npiv); s2 = add2(s2, d); q2 = fma2(d, d, q2); }
      st[j] = pk2(hsum2(s2), hsum2(q2));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int j = 0; j < LDL2_PA; ++j) st[j] = add2(st[j], __shfl_xor_sync(0xffffffffu, st[j], o));
    }
#pragma unroll
    for (int j = 0; j < LDL2_PA; ++j) {
      if (!ok[j]) continue;
      const int pos = base + j * LDL2_WARPS;
      float su, sq;
      upk2(st[j], su, sq);
      const float dm = su * (1.f / kC);
      const float rstd = rsqrt_fast(fmaxf(sq * (1.f / kC) - dm * dm, 0.f) + kLnEps);
      const f32x2 nmean = pk2(-(pivot[j] + dm)), rs2 = pk2(rstd);
      f32x2 o4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) o4[k] = mul2(add2(nxt[j][k], nmean), rs2);
      float* d = xh[pos - pos_first];
      *reinterpret_cast<ulonglong2*>(d + 4 * lane) = make_ulonglong2(o4[0], o4[1]);
      *reinterpret_cast<ulonglong2*>(d + 128 + 4 * lane) = make_ulonglong2(o4[2], o4[3]);
      if (SKIP) {
        float* rr = raw[pos - pos_first];
        *reinterpret_cast<ulonglong2*>(rr + 4 * lane) = make_ulonglong2(nxt[j][0], nxt[j][1]);
        *reinterpret_cast<ulonglong2*>(rr + 128 + 4 * lane) = make_ulonglong2(nxt[j][2], nxt[j][3]);
      }
    }
  }
  // ---- phase B: this warp's stream and row slice (its constants were fetched above)
  __syncthreads();
  const int n_rows = t1 - t0;
  const int per = ((n_rows + WPS - 1) / WPS + RPI - 1) / RPI * RPI;      // rows per warp, a multiple of RPI
  const int ra = t0 + part * per, rb = min(ra + per, t1);
  OutT* out_s = reinterpret_cast<OutT*>(p.out[s]);
  auto phase_b = [&](auto fast_tag) {
  constexpr bool FAST = decltype(fast_tag)::value;
  for (int tb = ra; tb < rb; tb += RPI) {
    const int w0 = STRIDE * (tb - t0);           // window row j of this iteration = xh[w0 + j] = position STRIDE tb - 1 + j
    const int pos0 = STRIDE * tb - 1;
    f32x2 xw[WIN][4];
#pragma unroll
    for (int j = 0; j < WIN; ++j) {
      const int pos = pos0 + j;
      if (FAST || (pos >= 0 && pos < p.t_virt && pos <= pos_last)) lds8p(xh[w0 + j], lane, xw[j]);
      else {
#pragma unroll
        for (int k = 0; k < 4; ++k) xw[j][k] = zero2;
      }
    }
    f32x2 acc[RPI][4], st[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      const int t = tb + r;
      const int j0 = STRIDE * r;                 // taps: window rows j0, j0 + 1, j0 + 2
      const bool ok0 = FAST || pos0 + j0 >= 0, ok2 = FAST || pos0 + j0 + 2 < p.t_virt;       // the centre tap is always inside
      const bool keep = FAST || (t < rb && ((tile_mask >> (t - t0)) & 1u));
      if (!keep) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[r][k] = zero2;
      } else if (ok0 && ok2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[r][k] = fma2(A2[k], xw[j0 + 2][k], fma2(A1[k], xw[j0 + 1][k], fma2(A0[k], xw[j0][k], Bs[k])));
      } else {                                   // edge row: per-tap bias terms, in the first layout's order
        f32x2 b0[4], b1[4], b2[4];
        lds8p(eb[s][0], lane, b0); lds8p(eb[s][1], lane, b1); lds8p(eb[s][2], lane, b2);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          f32x2 a = zero2;
          if (ok0) a = add2(a, fma2(A0[k], xw[j0][k], b0[k]));
          a = add2(a, fma2(A1[k], xw[j0 + 1][k], b1[k]));
          if (ok2) a = add2(a, fma2(A2[k], xw[j0 + 2][k], b2[k]));
          acc[r][k] = a;
        }
      }
      const f32x2 s2 = add2(add2(acc[r][0], acc[r][1]), add2(acc[r][2], acc[r][3]));
      f32x2 q2 = mul2(acc[r][0], acc[r][0]);
#pragma unroll
      for (int k = 1; k < 4; ++k) q2 = fma2(acc[r][k], acc[r][k], q2);
      st[r] = pk2(hsum2(s2), hsum2(q2));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < RPI; ++r) st[r] = add2(st[r], __shfl_xor_sync(0xffffffffu, st[r], o));
    }
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      const int t = tb + r;
      if (!FAST && t >= rb) continue;
      float su, sq;
      upk2(st[r], su, sq);
      const float m2 = su * (1.f / kC);
      const float r2 = rsqrt_fast(fmaxf(sq * (1.f / kC) - m2 * m2, 0.f) + kLnEps);
      const f32x2 rr = pk2(r2), mm = pk2(-m2 * r2);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[r][k] = fma2(fma2(acc[r][k], rr, mm), Wo[k], Bo[k]);
      store8p(out_s + ((size_t)b * p.out_rows + t) * kC, lane, acc[r]);
    }
    if (SKIP && s == 0) {                        // MaxPool1d(3, 2, 1) of the raw rows, -inf padding (STRIDE == 2)
#pragma unroll
      for (int r = 0; r < RPI; ++r) {
        const int t = tb + r;
        if (!FAST && t >= rb) continue;
        const int j0 = STRIDE * r;
        const bool ok0 = FAST || pos0 + j0 >= 0, ok2 = FAST || pos0 + j0 + 2 < p.t_virt;
        f32x2 r0[4], r1[4], r2v[4];
        lds8p(raw[w0 + j0 + 1], lane, r1);
        if (ok0) lds8p(raw[w0 + j0], lane, r0);
        if (ok2) lds8p(raw[w0 + j0 + 2], lane, r2v);
        float mx[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float m0, m1, l0, l1, h0, h1;
          upk2(r1[k], m0, m1);
          if (ok0) { upk2(r0[k], l0, l1); m0 = fmaxf(m0, l0); m1 = fmaxf(m1, l1); }
          if (ok2) { upk2(r2v[k], h0, h1); m0 = fmaxf(m0, h0); m1 = fmaxf(m1, h1); }
          mx[2 * k] = m0; mx[2 * k + 1] = m1;
        }
        float* so = p.skip_out + ((size_t)b * p.t_out + t) * kC;
        store4(so + 4 * lane, mx[0], mx[1], mx[2], mx[3]);
        store4(so + 128 + 4 * lane, mx[4], mx[5], mx[6], mx[7]);
      }
    }
  }
  };
  if (interior) phase_b(std::true_type{}); else phase_b(std::false_type{});
}

// ------------------------------------------------------------------------------------------------
// attention: warp per query row, 8 lanes per head (head dim 64), online softmax in fp32
// ------------------------------------------------------------------------------------------------
constexpr int ATT_ROWS = 4, ATT_WARPS = 8;

template <typename InT, typename OutT>
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_kernel(const InT* __restrict__ q, const InT* __restrict__ k,
                                                                  const InT* __restrict__ v, const unsigned char* __restrict__ kv_mask,
                                                                  OutT* __restrict__ out, int B, int T, int RPV, int window) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = lane * 8;
  const int tiles_per_video = (T + ATT_ROWS - 1) / ATT_ROWS;
  const long long n_tiles = (long long)B * tiles_per_video;
  const int half = window > 1 ? window / 2 : 0;
  const float scale = 0.125f;          // 1/sqrt(64)
  for (long long tile = (long long)blockIdx.x * ATT_WARPS + warp; tile < n_tiles; tile += (long long)gridDim.x * ATT_WARPS) {
    const int b = (int)(tile / tiles_per_video);
    const int t0 = (int)(tile - (long long)b * tiles_per_video) * ATT_ROWS;
    const size_t base = (size_t)b * T;          // mask / output rows
    const size_t qb = (size_t)b * RPV;          // q / k / v rows
    for (int i = t0; i < min(t0 + ATT_ROWS, T); ++i) {
      float qv[8];
      Row8<InT>::load(q + (qb + i) * kC + c0, qv);
#pragma unroll
      for (int d = 0; d < 8; ++d) qv[d] *= scale;
      const int lo = window > 1 ? max(0, i - half) : 0;
      const int hi = window > 1 ? min(T - 1, i + half) : T - 1;
      float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] = 0.f;
      for (int j = lo; j <= hi; ++j) {
        const bool mk = kv_mask ? (kv_mask[base + j] != 0) : true;
        if (window <= 1 && !mk) continue;                 // global: masked keys are -inf
        float kv[8], vv[8];
        Row8<InT>::load(k + (qb + j) * kC + c0, kv);
        Row8<InT>::load(v + (qb + j) * kC + c0, vv);
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) s = fmaf(qv[d], kv[d], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (!mk) s += -1e4f;                                // banded: additive -1e4 (blocks.py:1194-1195)
        const float mn = fmaxf(m, s);
        const float corr = expf(m - mn);                    // m = -inf on the first key -> 0
        const float pj = expf(s - mn);
        l = fmaf(l, corr, pj);
#pragma unroll
        for (int d = 0; d < 8; ++d) acc[d] = fmaf(acc[d], corr, pj * vv[d]);
        m = mn;
      }
      const bool qok = (window > 1 && kv_mask) ? (kv_mask[base + i] != 0) : true;   // blocks.py:1208-1209
      const float inv = (qok && l > 0.f) ? 1.f / l : 0.f;
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] *= inv;
      Row8<OutT>::store(out + (base + i) * kC + c0, acc);
    }
  }
}


// Banded attention with the K / V rows of a query tile staged in shared memory: a CTA owns ATB_ROWS consecutive
// query rows of one video and cp.async-loads rows [t0 - HALF, t0 + ATB_ROWS + HALF) of K and V once (each K/V row is
// otherwise re-read from L1/L2 by 2*HALF+1 different queries). One warp per query row, 8 lanes per head; the
// 2*HALF+1 scores of a row are computed first (independent dot products), then a two-pass softmax in base 2
// (log2(e) folded into the query scale, ex2.approx), then the weighted sum of V.
constexpr int ATB_MAX_ROWS = 64, ATB_WARPS = 8;   // query rows per CTA are a launch parameter (<= ATB_MAX_ROWS)

template <typename InT, typename OutT, int HALF>
__global__ void __launch_bounds__(ATB_WARPS * 32) attention_banded_kernel(const InT* __restrict__ q, const InT* __restrict__ k,
                                                                         const InT* __restrict__ v, const unsigned char* __restrict__ kv_mask,
                                                                         OutT* __restrict__ out, int B, int T, int RPV, int ATB_ROWS) {
  extern __shared__ __align__(16) unsigned char att_smem[];
  constexpr int ROWB = kC * (int)sizeof(InT);                  // bytes per K or V row
  constexpr int W = 2 * HALF + 1;
  const int NROW = ATB_ROWS + 2 * HALF;
  InT* ks = reinterpret_cast<InT*>(att_smem);
  InT* vs = reinterpret_cast<InT*>(att_smem + (size_t)NROW * ROWB);
  __shared__ unsigned char s_mask[ATB_MAX_ROWS + 2 * HALF];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_per_video = (T + ATB_ROWS - 1) / ATB_ROWS;
  const int b = blockIdx.x / tiles_per_video;
  const int t0 = (blockIdx.x - b * tiles_per_video) * ATB_ROWS;
  const int t1 = min(t0 + ATB_ROWS, T);
  const int lo = t0 - HALF;                                    // smem row r holds key row lo + r (if inside [0, T))
  const size_t qb = (size_t)b * RPV, base = (size_t)b * T;
  constexpr int CH = ROWB / 16;                                // 16-byte chunks per row
  for (int i = threadIdx.x; i < NROW * CH; i += blockDim.x) {
    const int r = i / CH, c = i - r * CH;
    const int j = lo + r;
    if (j < 0 || j >= T || j >= t1 + HALF) continue;
    const size_t g = ((qb + j) * kC) * sizeof(InT) + (size_t)c * 16;
    const unsigned sk = (unsigned)__cvta_generic_to_shared(reinterpret_cast<unsigned char*>(ks) + (size_t)r * ROWB + c * 16);
    const unsigned sv = (unsigned)__cvta_generic_to_shared(reinterpret_cast<unsigned char*>(vs) + (size_t)r * ROWB + c * 16);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sk), "l"(reinterpret_cast<const unsigned char*>(k) + g) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sv), "l"(reinterpret_cast<const unsigned char*>(v) + g) : "memory");
  }
  for (int i = threadIdx.x; i < NROW; i += blockDim.x) {
    const int j = lo + i;
    s_mask[i] = (j >= 0 && j < T) ? (kv_mask ? kv_mask[base + j] : 1) : 0;
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int c0 = lane * 8;
  const float scale = 0.125f * 1.4426950408889634f;           // 1/sqrt(64) * log2(e)
  for (int i = t0 + warp; i < t1; i += ATB_WARPS) {
    float qv[8];
    Row8<InT>::load(q + (qb + i) * kC + c0, qv);
#pragma unroll
    for (int d = 0; d < 8; ++d) qv[d] *= scale;
    const int r0 = i - HALF - lo;                               // smem row of key i - HALF (= i - t0 >= 0)
    float sc[W];
    float m = -INFINITY;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int j = i - HALF + w;
      float s = 0.f;
      if (j >= 0 && j < T) {                                    // warp-uniform
        float kv[8];
        Row8<InT>::load(ks + (size_t)(r0 + w) * kC + c0, kv);
#pragma unroll
        for (int d = 0; d < 8; ++d) s = fmaf(qv[d], kv[d], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (j < 0 || j >= T) s = -INFINITY;                       // outside the sequence
      else if (!s_mask[r0 + w]) s += -1e4f * 1.4426950408889634f;   // blocks.py:1194-1195 (additive, base-2 domain)
      sc[w] = s;
      m = fmaxf(m, s);
    }
    float l = 0.f, acc[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] = 0.f;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int j = i - HALF + w;
      if (j >= 0 && j < T) {
        float pj;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pj) : "f"(sc[w] - m));
        l += pj;
        float vv[8];
        Row8<InT>::load(vs + (size_t)(r0 + w) * kC + c0, vv);
#pragma unroll
        for (int d = 0; d < 8; ++d) acc[d] = fmaf(pj, vv[d], acc[d]);
      }
    }
    const float inv = (s_mask[i - lo] && l > 0.f) ? 1.f / l : 0.f;   // blocks.py:1208-1209
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] *= inv;
    Row8<OutT>::store(out + (base + i) * kC + c0, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over channels of fp32 rows (C = 256 * NCH)
// ------------------------------------------------------------------------------------------------
// A warp normalises LNR_ROWS rows per iteration (their loads are in flight together: one 1 KB row per warp left the
// kernel latency-bound at ~3 TB/s) and the host sizes the grid so that every CTA runs the same number of iterations.
constexpr int LNR_ROWS = 2;
template <typename OutT, int NCH>
__global__ void __launch_bounds__(256) ln_rows_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bvec, OutT* __restrict__ out, long long rows) {
  pdl_trigger();                 // the fused MLP that follows may start its prologue now (it waits for this grid)
  const int lane = threadIdx.x & 31;
  const int C = kC * NCH;
  for (long long r0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * LNR_ROWS; r0 < rows; r0 += (long long)gridDim.x * 8 * LNR_ROWS) {
    float v[LNR_ROWS][NCH][8];
    float s[LNR_ROWS];
#pragma unroll
    for (int i = 0; i < LNR_ROWS; ++i) {
      const long long r = r0 + i < rows ? r0 + i : rows - 1;      // a tail warp re-reads the last row and skips the store
#pragma unroll
      for (int h = 0; h < NCH; ++h) Row8<float>::load(x + (size_t)r * C + h * kC + lane * 8, v[i][h]);
    }
#pragma unroll
    for (int i = 0; i < LNR_ROWS; ++i) {
      s[i] = 0.f;
#pragma unroll
      for (int h = 0; h < NCH; ++h)
#pragma unroll
        for (int k = 0; k < 8; ++k) s[i] += v[i][h][k];
    }
    float mean[LNR_ROWS], qq[LNR_ROWS];
#pragma unroll
    for (int i = 0; i < LNR_ROWS; ++i) mean[i] = warp_sum(s[i]) / (float)C;
#pragma unroll
    for (int i = 0; i < LNR_ROWS; ++i) {
      qq[i] = 0.f;
#pragma unroll
      for (int h = 0; h < NCH; ++h)
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float d = v[i][h][k] - mean[i]; qq[i] = fmaf(d, d, qq[i]); }
    }
    float rstd[LNR_ROWS];
#pragma unroll
    for (int i = 0; i < LNR_ROWS; ++i) rstd[i] = 1.f / sqrtf(warp_sum(qq[i]) / (float)C + kLnEps);
#pragma unroll
    for (int h = 0; h < NCH; ++h) {
      float ww[8], bb[8];
      Row8<float>::load(w + h * kC + lane * 8, ww);
      Row8<float>::load(bvec + h * kC + lane * 8, bb);
#pragma unroll
      for (int i = 0; i < LNR_ROWS; ++i) {
        if (r0 + i >= rows) break;
#pragma unroll
        for (int k = 0; k < 8; ++k) v[i][h][k] = fmaf((v[i][h][k] - mean[i]) * rstd[i], ww[k], bb[k]);
        Row8<OutT>::store(out + (size_t)(r0 + i) * C + h * kC + lane * 8, v[i][h]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// InstanceNorm1d over T + LeakyReLU. CTA = (video, 32-channel slab); threadIdx.x = channel (128 B
// coalesced rows), threadIdx.y splits T. Two-pass statistics (second and third pass hit L2).
// ------------------------------------------------------------------------------------------------
constexpr int IN_TY = 16;
template <typename OutT>
__global__ void __launch_bounds__(32 * IN_TY) instnorm_lrelu_kernel(const float* __restrict__ x, OutT* __restrict__ out,
                                                                   int T, int C, float slope) {
  __shared__ float red[IN_TY][33];
  const int b = blockIdx.y, c = blockIdx.x * 32 + threadIdx.x;
  const float* xb = x + (size_t)b * T * C + c;
  OutT* ob = out + (size_t)b * T * C + c;
  float s = 0.f;
  for (int t = threadIdx.y; t < T; t += IN_TY) s += xb[(size_t)t * C];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < IN_TY; ++i) tot += red[i][threadIdx.x];
  const float mean = tot / (float)T;
  __syncthreads();
  float q = 0.f;
  for (int t = threadIdx.y; t < T; t += IN_TY) { const float d = xb[(size_t)t * C] - mean; q = fmaf(d, d, q); }
  red[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  tot = 0.f;
#pragma unroll
  for (int i = 0; i < IN_TY; ++i) tot += red[i][threadIdx.x];
  const float rstd = 1.f / sqrtf(tot / (float)T + kLnEps);
  for (int t = threadIdx.y; t < T; t += IN_TY) {
    float v = (xb[(size_t)t * C] - mean) * rstd;
    v = v >= 0.f ? v : v * slope;
    store1(ob + (size_t)t * C, v);
  }
}

// Same, with the thread's T / IN_TY values held in registers: x is read once instead of three times (the launches of
// this path have T <= 768). Same summation order as the kernel above.
template <typename OutT, int NR>
__global__ void __launch_bounds__(32 * IN_TY) instnorm_lrelu_reg_kernel(const float* __restrict__ x, OutT* __restrict__ out,
                                                                       int T, int C, float slope) {
  __shared__ float red[IN_TY][33];
  const int b = blockIdx.y, c = blockIdx.x * 32 + threadIdx.x;
  const float* xb = x + (size_t)b * T * C + c;
  OutT* ob = out + (size_t)b * T * C + c;
  float v[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int t = threadIdx.y + i * IN_TY;
    v[i] = t < T ? xb[(size_t)t * C] : 0.f;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NR; ++i) s += v[i];            // rows beyond T hold 0
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < IN_TY; ++i) tot += red[i][threadIdx.x];
  const float mean = tot / (float)T;
  __syncthreads();
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    if (threadIdx.y + i * IN_TY < T) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  }
  red[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  tot = 0.f;
#pragma unroll
  for (int i = 0; i < IN_TY; ++i) tot += red[i][threadIdx.x];
  const float rstd = 1.f / sqrtf(tot / (float)T + kLnEps);
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int t = threadIdx.y + i * IN_TY;
    if (t < T) {
      float y = (v[i] - mean) * rstd;
      y = y >= 0.f ? y : y * slope;
      store1(ob + (size_t)t * C, y);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// FPN top-down fuse + depthwise conv + LN, warp per pyramid row
// ------------------------------------------------------------------------------------------------
struct FpnParams {
  const float* lat; const unsigned char* mask; const float* dw_w; const float* ln_w; const float* ln_b;
  void* out; int B, n_levels, P;
  int lvl_off[AVDF_MAX_LEVELS], lvl_len[AVDF_MAX_LEVELS];
};

// Latency-bound as first written (18 dependent 1 KB row loads per output row behind run-time loop bounds: 125 us for
// 74 MB): the level loop is unrolled over AVDF_MAX_LEVELS with predicates so that the loads of a tap (one row per
// pyramid level above the output row) are in flight together, and a warp keeps the depthwise / LayerNorm weights of
// its current level in registers (rows of a video are sorted by level).
template <typename OutT>
__global__ void __launch_bounds__(256, 2) fpn_fuse_kernel(const FpnParams p) {
  const int lane = threadIdx.x & 31;
  const int c0 = lane * 8;
  const long long rows = (long long)p.B * p.P;
  int cur_l = -1;
  float dwr[3][8], lw[8], lb[8];
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * 8) {
    const int b = (int)(r / p.P);
    const int pr = (int)(r - (long long)b * p.P);
    int l = 0;
    while (l + 1 < p.n_levels && pr >= p.lvl_off[l + 1]) ++l;
    const int t = pr - p.lvl_off[l], T = p.lvl_len[l];
    if (l != cur_l) {
      const float* dw = p.dw_w + (size_t)l * kC * 3;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        dwr[0][k] = __ldg(dw + (c0 + k) * 3); dwr[1][k] = __ldg(dw + (c0 + k) * 3 + 1); dwr[2][k] = __ldg(dw + (c0 + k) * 3 + 2);
      }
      if (p.ln_w) {
        Row8<float>::load(p.ln_w + (size_t)l * kC + c0, lw);
        Row8<float>::load(p.ln_b + (size_t)l * kC + c0, lb);
      }
      cur_l = l;
    }
    const float* lat_b = p.lat + (size_t)b * p.P * kC;
    const float mk = p.mask ? (p.mask[r] ? 1.f : 0.f) : 1.f;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int tt = t + j - 1;
      const bool tin = tt >= 0 && tt < T;
      // L_l[tt] = lat_l[tt] + (lat_{l+1}[tt>>1] + (... )) summed from the top level down (necks.py:76-80)
      float a[AVDF_MAX_LEVELS][8];
#pragma unroll
      for (int jl = 0; jl < AVDF_MAX_LEVELS; ++jl) {
        const int dl = jl - l;
        const int ts = tt >> (dl > 0 ? dl : 0);
        const bool ok = tin && jl < p.n_levels && dl >= 0 && ts < p.lvl_len[jl];
        if (ok) Row8<float>::load(lat_b + (size_t)(p.lvl_off[jl] + ts) * kC + c0, a[jl]);
        else {
#pragma unroll
          for (int k = 0; k < 8; ++k) a[jl][k] = 0.f;
        }
      }
      float sum[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) sum[k] = 0.f;
#pragma unroll
      for (int jl = AVDF_MAX_LEVELS - 1; jl >= 0; --jl)
#pragma unroll
        for (int k = 0; k < 8; ++k) sum[k] += a[jl][k];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(dwr[j][k], sum[k], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= mk;
    if (p.ln_w) {
      float m, rs;
      row_stats(acc, m, rs);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf((acc[k] - m) * rs, lw[k], lb[k]);
    }
    Row8<OutT>::store(reinterpret_cast<OutT*>(p.out) + (size_t)r * kC + c0, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// last conv (k3) of both heads, warp per pyramid row: 1 + 2 dot products of length 3 * 256
// ------------------------------------------------------------------------------------------------
struct HeadParams {
  const void* cls_feat; const void* reg_feat; const unsigned char* mask;
  const float* cls_w; const float* cls_b; const float* reg_w; const float* reg_b;
  float* logits; float* offsets; int B, n_levels, P;
  int lvl_off[AVDF_MAX_LEVELS], lvl_len[AVDF_MAX_LEVELS]; float lvl_scale[AVDF_MAX_LEVELS];
};

// Column-blocked variant: the kernel above re-reads every lateral row up to 18 times through L2 (~600 MB for a 50 MB
// tensor: L2-bandwidth bound, ~100 us). Here a CTA owns FPB level-0 positions of one video and the rows of all coarser
// levels above them (FPB >> l rows at level l, plus one halo row on each side for the k3 taps): phase 1 builds the
// top-down sums L_l[t] = L_{l+1}[t >> 1] + lat_l[t] once per row in shared memory (coarsest level first, same
// summation order as necks.py:76-80), phase 2 runs the depthwise conv + mask + LayerNorm from shared memory.
// Every lateral row is read once per CTA (plus halos).
constexpr int FPB = 32;                       // level-0 positions per CTA
template <typename OutT>
__global__ void __launch_bounds__(256) fpn_fuse_cols_kernel(const FpnParams p) {
  extern __shared__ __align__(16) float fpn_smem[];          // [sum_l (FPB >> l) + 2][256]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = lane * 8;
  const int blocks_per_video = p.lvl_len[0] / FPB;
  const int b = blockIdx.x / blocks_per_video;
  const int blk = blockIdx.x - b * blocks_per_video;
  const float* lat_b = p.lat + (size_t)b * p.P * kC;
  int row0[AVDF_MAX_LEVELS];                                 // first smem row of level l (its row i holds position s_l - 1 + i)
  {
    int acc_rows = 0;
#pragma unroll
    for (int l = 0; l < AVDF_MAX_LEVELS; ++l) { row0[l] = acc_rows; acc_rows += (FPB >> l) + 2; }
  }
  // phase 0: every lateral row this CTA needs, all levels at once (cp.async, one wait): fetched level by level behind
  // the top-down dependency the six global-load round trips were serialised (~80 us for 75 MB)
  for (int l = 0; l < p.n_levels; ++l) {
    const int n = (FPB >> l) + 2, s = ((blk * FPB) >> l) - 1, T = p.lvl_len[l];
    for (int i = warp; i < n; i += 8) {
      const int pos = s + i;
      float* d = fpn_smem + (size_t)(row0[l] + i) * kC + c0;
      if (pos >= 0 && pos < T) {
        const float* g = lat_b + (size_t)(p.lvl_off[l] + pos) * kC + c0;
        const unsigned sd = (unsigned)__cvta_generic_to_shared(d);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sd), "l"(g) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sd + 16), "l"(g + 4) : "memory");
      } else {                                                 // outside the level: the conv's zero padding
        *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(d + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // phase 1: top-down sums in shared memory, coarsest level first (L_l[t] = L_{l+1}[t >> 1] + lat_l[t])
  for (int l = p.n_levels - 2; l >= 0; --l) {
    const int n = (FPB >> l) + 2, s = ((blk * FPB) >> l) - 1, T = p.lvl_len[l];
    for (int i = warp; i < n; i += 8) {
      const int pos = s + i;
      if (pos >= 0 && pos < T && (pos >> 1) < p.lvl_len[l + 1]) {
        const int pi = (pos >> 1) - (((blk * FPB) >> (l + 1)) - 1);
        float* d = fpn_smem + (size_t)(row0[l] + i) * kC + c0;
        float v[8], u[8];
        lds8(d, v);
        lds8(fpn_smem + (size_t)(row0[l + 1] + pi) * kC + c0, u);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = u[k] + v[k];
        *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    __syncthreads();
  }
  // phase 2: depthwise k3 + mask + LayerNorm. The output rows are cut into items of <= 8 rows of ONE level and an item
  // belongs to one warp, which loads that level's weights once, as six 16-byte loads of its 24 contiguous floats (every
  // warp walking every level re-loaded them six times with 24 scalar loads 96 bytes apart: ~4600 L1 sector requests per
  // warp, what bounded the kernel at ~80 us).
  int item = 0;
  for (int l = 0; l < p.n_levels; ++l) {
    const int n = FPB >> l, s = (blk * FPB) >> l;
    for (int first = 0; first < n; first += 8, ++item) {
      if ((item & 7) != warp) continue;
      const int cnt = min(8, n - first);
      float dwr[3][8], lw[8], lb[8];
      {
        const float4* dw4 = reinterpret_cast<const float4*>(p.dw_w + (size_t)l * kC * 3 + (size_t)c0 * 3);   // [8 channels][3 taps]
        float wv[24];
#pragma unroll
        for (int q = 0; q < 6; ++q) { const float4 t4 = __ldg(dw4 + q); wv[4 * q] = t4.x; wv[4 * q + 1] = t4.y; wv[4 * q + 2] = t4.z; wv[4 * q + 3] = t4.w; }
#pragma unroll
        for (int k = 0; k < 8; ++k) { dwr[0][k] = wv[3 * k]; dwr[1][k] = wv[3 * k + 1]; dwr[2][k] = wv[3 * k + 2]; }
      }
      if (p.ln_w) {
        Row8<float>::load(p.ln_w + (size_t)l * kC + c0, lw);
        Row8<float>::load(p.ln_b + (size_t)l * kC + c0, lb);
      }
      for (int i = first; i < first + cnt; ++i) {
        const size_t r = (size_t)b * p.P + p.lvl_off[l] + s + i;
        const float mk = p.mask ? (p.mask[r] ? 1.f : 0.f) : 1.f;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float u[8];
          lds8(fpn_smem + (size_t)(row0[l] + i + j) * kC + c0, u);   // smem row i + j holds position s + i + j - 1
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaf(dwr[j][k], u[k], acc[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] *= mk;
        if (p.ln_w) {
          float m, rs;
          row_stats(acc, m, rs);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaf((acc[k] - m) * rs, lw[k], lb[k]);
        }
        Row8<OutT>::store(reinterpret_cast<OutT*>(p.out) + r * kC + c0, acc);
      }
    }
  }
}

// The 3 x 3 x 256 weights live in registers (72 per lane, loaded once per warp) and the six activation rows of an
// output row are loaded together: as first written every row re-read 9 KB of weights through L1 (58 us for 50 MB).
template <typename InT>
__global__ void __launch_bounds__(256) head_final_kernel(const HeadParams p) {
  const int lane = threadIdx.x & 31;
  const int c0 = lane * 8;
  const long long rows = (long long)p.B * p.P;
  const InT* cf = reinterpret_cast<const InT*>(p.cls_feat);
  const InT* rf = reinterpret_cast<const InT*>(p.reg_feat);
  float wc[3][8], wr0[3][8], wr1[3][8];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    Row8<float>::load(p.cls_w + j * kC + c0, wc[j]);
    Row8<float>::load(p.reg_w + j * kC + c0, wr0[j]);
    Row8<float>::load(p.reg_w + 3 * kC + j * kC + c0, wr1[j]);
  }
  const float cb = p.cls_b[0], rb0 = p.reg_b[0], rb1 = p.reg_b[1];
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * 8) {
    const int b = (int)(r / p.P);
    const int pr = (int)(r - (long long)b * p.P);
    int l = 0;
    while (l + 1 < p.n_levels && pr >= p.lvl_off[l + 1]) ++l;
    const int t = pr - p.lvl_off[l], T = p.lvl_len[l];
    float xc[3][8], xr[3][8];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int tt = t + j - 1;
      if (tt >= 0 && tt < T) {
        const size_t row = (size_t)(r + (j - 1));
        Row8<InT>::load(cf + row * kC + c0, xc[j]);
        Row8<InT>::load(rf + row * kC + c0, xr[j]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) { xc[j][k] = 0.f; xr[j][k] = 0.f; }
      }
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
      for (int k = 0; k < 8; ++k) a0 = fmaf(wc[j][k], xc[j][k], a0);
#pragma unroll
      for (int k = 0; k < 8; ++k) a1 = fmaf(wr0[j][k], xr[j][k], a1);
#pragma unroll
      for (int k = 0; k < 8; ++k) a2 = fmaf(wr1[j][k], xr[j][k], a2);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) {
      const float mk = p.mask ? (p.mask[r] ? 1.f : 0.f) : 1.f;
      p.logits[r] = (a0 + cb) * mk;
      const float sc = p.lvl_scale[l];
      p.offsets[2 * r] = fmaxf(((a1 + rb0) * mk) * sc, 0.f);
      p.offsets[2 * r + 1] = fmaxf(((a2 + rb1) * mk) * sc, 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// exp12 video-level tail, one CTA (256 threads) per video, thread = output channel:
//   g = LeakyReLU_.2(IN_T(W0 z)) [T, C]; pooled = [max_t g | mean_t g]; h = ReLU(LN(W1 pooled)); out = w2.h + b2
// The 1x1 conv is 256 rank-1 updates: thread c reads W0T[k][c] (coalesced) and z[k][0..T) (smem broadcast) and keeps
// its channel's T values in registers, so the InstanceNorm over T and the pooling need no communication at all.
// Weights arrive TRANSPOSED ([in, out]) for exactly that access pattern.
// ------------------------------------------------------------------------------------------------
constexpr int VC12_MAXT = 32;
template <typename InT>
__global__ void __launch_bounds__(256) vcls_exp12_kernel(const InT* __restrict__ z, const float* __restrict__ w0t,
                                                         const float* __restrict__ w1t, const float* __restrict__ ln_w,
                                                         const float* __restrict__ ln_b, const float* __restrict__ w2,
                                                         const float* __restrict__ b2, float* __restrict__ out, int T) {
  __shared__ __align__(16) float zs[kC][VC12_MAXT];      // z transposed: [k][t]
  __shared__ float pooled[2 * kC];
  __shared__ float red[2][8];
  const int b = blockIdx.x, c = threadIdx.x, lane = c & 31, warp = c >> 5;
  for (int i = c; i < T * kC; i += 256) { const int t = i / kC, k = i - t * kC; zs[k][t] = load1(z + (size_t)b * T * kC + i); }
  for (int i = c; i < kC * (VC12_MAXT - T); i += 256) { const int k = i / (VC12_MAXT - T), t = T + i - k * (VC12_MAXT - T); zs[k][t] = 0.f; }
  __syncthreads();
  float g[VC12_MAXT];
#pragma unroll
  for (int t = 0; t < VC12_MAXT; ++t) g[t] = 0.f;
#pragma unroll 8
  for (int k = 0; k < kC; ++k) {            // unrolled: 8 independent weight loads in flight per thread
    const float w = __ldg(w0t + (size_t)k * kC + c);
    const float4* zr = reinterpret_cast<const float4*>(&zs[k][0]);
#pragma unroll
    for (int j = 0; j < VC12_MAXT / 4; ++j) {
      const float4 zz = zr[j];
      g[4 * j] = fmaf(w, zz.x, g[4 * j]); g[4 * j + 1] = fmaf(w, zz.y, g[4 * j + 1]);
      g[4 * j + 2] = fmaf(w, zz.z, g[4 * j + 2]); g[4 * j + 3] = fmaf(w, zz.w, g[4 * j + 3]);
    }
  }
  float s1 = 0.f;
#pragma unroll
  for (int t = 0; t < VC12_MAXT; ++t) if (t < T) s1 += g[t];
  const float mean = s1 / (float)T;
  float qv = 0.f;
#pragma unroll
  for (int t = 0; t < VC12_MAXT; ++t) if (t < T) { const float d = g[t] - mean; qv = fmaf(d, d, qv); }
  const float rstd = 1.f / sqrtf(qv / (float)T + kLnEps);
  float mx = -INFINITY, sum_g = 0.f;
#pragma unroll
  for (int t = 0; t < VC12_MAXT; ++t) if (t < T) {
    float v = (g[t] - mean) * rstd;
    v = v >= 0.f ? v : 0.2f * v;
    mx = fmaxf(mx, v); sum_g += v;
  }
  pooled[c] = mx; pooled[kC + c] = sum_g / (float)T;
  __syncthreads();
  float hc = 0.f;                                        // h[c] = sum_j W1[c][j] pooled[j]
#pragma unroll 16
  for (int j = 0; j < 2 * kC; ++j) hc = fmaf(__ldg(w1t + (size_t)j * kC + c), pooled[j], hc);
  // LayerNorm over the 256 channels (two-pass), ReLU, dot with w2
  float s = warp_sum(hc);
  if (lane == 0) red[0][warp] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[0][i];
  const float m = tot / (float)kC;
  const float d = hc - m;
  float q2 = warp_sum(d * d);
  if (lane == 0) red[1][warp] = q2;
  __syncthreads();
  tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[1][i];
  const float rs = 1.f / sqrtf(tot / (float)kC + kLnEps);
  const float y = fmaxf(fmaf(d * rs, ln_w[c], ln_b[c]), 0.f);
  float a = warp_sum(w2[c] * y);
  __syncthreads();
  if (lane == 0) red[0][warp] = a;
  __syncthreads();
  if (c == 0) {
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += red[0][i];
    out[b] = r + b2[0];
  }
}

// ------------------------------------------------------------------------------------------------
// exp13 video-level tail, one CTA per video, C (<= 128) channels, thread = (channel c, time lane):
//   g = LeakyReLU_.2(IN_T(W0 z)) [T, C]; s_t = seg_w . g_t + seg_b; out = cls_w . [max_t s, mean_t s] + cls_b
// Each thread keeps row c of W0 in registers; z rows are read once per pass (L1 broadcast across the C threads of a
// time lane). Pass 1 accumulates sum / sum of squares of g per channel, pass 2 recomputes g, normalises and reduces
// over channels with one shuffle tree per time step.
// ------------------------------------------------------------------------------------------------
template <typename InT, int C>
__global__ void __launch_bounds__(256) vcls_exp13_kernel(const InT* __restrict__ z, const float* __restrict__ w0,
                                                         const float* __restrict__ seg_w, const float* __restrict__ seg_b,
                                                         const float* __restrict__ cls_w, const float* __restrict__ cls_b,
                                                         float* __restrict__ out, int T) {
  constexpr int TL = 256 / C;                            // time lanes
  __shared__ float red[2][256];
  __shared__ float stat[2][C];
  __shared__ float zrow[TL][C];
  __shared__ float spart[TL][C / 32];
  __shared__ float sfin[2][TL];
  const int b = blockIdx.x, tid = threadIdx.x, c = tid % C, tl = tid / C, lane = tid & 31;
  const InT* zb = z + (size_t)b * T * C;
  float w[C];
#pragma unroll
  for (int k = 0; k < C; ++k) w[k] = __ldg(w0 + (size_t)c * C + k);
  const int steps = (T + TL - 1) / TL;
  float s1 = 0.f, s2 = 0.f;
  for (int i = 0; i < steps; ++i) {
    const int t = i * TL + tl;
    __syncthreads();
    if (t < T) zrow[tl][c] = load1(zb + (size_t)t * C + c);
    __syncthreads();
    if (t < T) {
      float g = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) g = fmaf(w[k], zrow[tl][k], g);
      s1 += g; s2 = fmaf(g, g, s2);
    }
  }
  red[0][tid] = s1; red[1][tid] = s2;
  __syncthreads();
  if (tid < C) {
    float a = 0.f, q = 0.f;
    for (int j = 0; j < TL; ++j) { a += red[0][j * C + tid]; q += red[1][j * C + tid]; }
    const float m = a / (float)T;
    stat[0][tid] = m;
    stat[1][tid] = 1.f / sqrtf(fmaxf(q / (float)T - m * m, 0.f) + kLnEps);
  }
  __syncthreads();
  const float mean = stat[0][c], rstd = stat[1][c], sw = seg_w[c];
  float smax = -INFINITY, ssum = 0.f;
  for (int i = 0; i < steps; ++i) {
    const int t = i * TL + tl;
    __syncthreads();
    if (t < T) zrow[tl][c] = load1(zb + (size_t)t * C + c);
    __syncthreads();
    float v = 0.f;
    if (t < T) {
      float g = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) g = fmaf(w[k], zrow[tl][k], g);
      v = (g - mean) * rstd;
      v = (v >= 0.f ? v : 0.2f * v) * sw;
    }
    v = warp_sum(v);                                     // C is a multiple of 32: a warp never straddles time lanes
    if (lane == 0) spart[tl][(tid % C) / 32] = v;
    __syncthreads();
    if (c == 0 && t < T) {
      float st = seg_b[0];
#pragma unroll
      for (int j = 0; j < C / 32; ++j) st += spart[tl][j];
      smax = fmaxf(smax, st); ssum += st;
    }
  }
  if (c == 0) { sfin[0][tl] = smax; sfin[1][tl] = ssum; }
  __syncthreads();
  if (tid == 0) {
    float mx = -INFINITY, sm_ = 0.f;
    for (int j = 0; j < TL; ++j) { mx = fmaxf(mx, sfin[0][j]); sm_ += sfin[1][j]; }
    out[b] = cls_w[0] * mx + cls_w[1] * (sm_ / (float)T) + cls_b[0];
  }
}

int attention_banded_mma(const void* q, const void* k, const void* v, const unsigned char* kv_mask, void* out, int in_dtype,
                         int out_dtype, int batch, int t, int rpv, cudaStream_t st);       // attention_mma.cu

static int grid_for(long long work_items, int per_cta, int sms) {
  long long want = (work_items + per_cta - 1) / per_cta;
  long long cap = (long long)sms * 8;
  if (want > cap) {                      // grid-stride: the same number of iterations for every CTA (no ragged last wave)
    const long long iters = (want + cap - 1) / cap;
    want = (want + iters - 1) / iters;
  }
  return (int)(want < 1 ? 1 : want);
}
static int sm_count() { return device_sm_count(); }

}  // namespace avdf

using namespace avdf;

extern "C" int avdf_ln_dwconv_ln(const avdf_ln_dwconv_ln_args* a, void* stream) {
  AVDF_CHECK_ARG(a != nullptr, "args is null");
  AVDF_CHECK_ARG(a->channels == kC, "channels must be 256");
  AVDF_CHECK_ARG(a->stride == 1 || a->stride == 2, "stride must be 1 or 2");
  AVDF_CHECK_ARG(a->n_streams >= 1 && a->n_streams <= 3, "n_streams out of range");
  AVDF_CHECK_ARG(a->batch >= 0 && a->t_src > 0 && a->t_virt > 0 && a->t_virt % a->stride == 0, "bad sizes");
  AVDF_CHECK_ARG(a->shift > -16 && a->shift < 16, "shift out of range");
  AVDF_CHECK_ARG(a->shift >= 0 ? (((a->t_virt - 1) >> a->shift) < a->t_src) : (((a->t_virt - 1) << -a->shift) < a->t_src),
                 "virtual length maps outside the source");
  AVDF_CHECK_DTYPE(a->out_dtype, "out_dtype");
  AVDF_CHECK_ARG(a->src != nullptr, "src is null");
  AVDF_CHECK_ARG(a->skip_out == nullptr || (a->stride == 2 && a->shift == 0), "skip_out needs stride 2 and no resampling");
  LdlParams p{};
  p.src = a->src; p.mask_out = a->mask_out; p.skip_out = a->skip_out;
  for (int s = 0; s < a->n_streams; ++s) {
    AVDF_CHECK_ARG(a->ln_in_w[s] && a->ln_in_b[s] && a->dw_w[s] && a->ln_out_w[s] && a->ln_out_b[s] && a->out[s], "null stream parameter");
    p.ln_in_w[s] = a->ln_in_w[s]; p.ln_in_b[s] = a->ln_in_b[s]; p.dw_w[s] = a->dw_w[s];
    p.ln_out_w[s] = a->ln_out_w[s]; p.ln_out_b[s] = a->ln_out_b[s]; p.out[s] = a->out[s];
  }
  p.B = a->batch; p.t_src = a->t_src; p.t_virt = a->t_virt; p.shift = a->shift; p.t_out = a->t_virt / a->stride;
  p.out_rows = a->out_rows_per_video > 0 ? a->out_rows_per_video : p.t_out;
  AVDF_CHECK_ARG(p.out_rows >= p.t_out, "out_rows_per_video smaller than the output length");
  p.n_streams = a->n_streams;
  if (a->batch == 0) return AVDF_OK;
  AVDF_CHECK_ARG(a->tile_rows == 0 || a->tile_rows == 2 || a->tile_rows == 4 || a->tile_rows == 8, "tile_rows must be 0 (auto), 2, 4 or 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const int ldl_v1 = getenv("AVDF_LDL_V1") ? atoi(getenv("AVDF_LDL_V1")) : 0;
  if (a->tile_rows == 0 && !ldl_v1 && a->n_streams >= 2) {
    // second layout (see ln_dwconv_ln2_kernel): one CTA per 32 output rows of a video. Measured (batch 32, T = 768, us):
    // three streams 30.9 -> 27.7, stride 2 39.1 -> 33.2, ONE stream 16.5 -> 18.6 (nothing to share between streams: the
    // first layout's per-warp cp.async pipeline wins), so single-stream launches, an explicit tile_rows (tests, scripts)
    // and AVDF_LDL_V1=1 keep the first layout
    const int grid2 = a->batch * ((p.t_out + LDL2_ROWS - 1) / LDL2_ROWS);
#define AVDF_LDL2(S, NS, SK)                                                                                         \
  AVDF_DISPATCH_DTYPE(a->out_dtype, OutT, {                                                                          \
    const size_t smem = (size_t)(ldl2_src_rows<S>() * (SK ? 2 : 1) + NS * 3) * kC * sizeof(float);                   \
    AVDF_SMEM_ATTR_ONCE((ln_dwconv_ln2_kernel<OutT, S, NS, SK>), smem);                                              \
    { cudaError_t le_ = launch_pdl(ln_dwconv_ln2_kernel<OutT, S, NS, SK>, grid2, LDL2_WARPS * 32, smem, st, p); if (le_ != cudaSuccess) { set_error("ln_dwconv_ln2_kernel: launch failed: %s", cudaGetErrorString(le_)); return AVDF_ERR_CUDA; } } \
  })
    if (a->stride == 1) { if (a->n_streams == 1) AVDF_LDL2(1, 1, false); else if (a->n_streams == 2) AVDF_LDL2(1, 2, false); else AVDF_LDL2(1, 3, false); }
    else if (a->skip_out) { if (a->n_streams == 1) AVDF_LDL2(2, 1, true); else if (a->n_streams == 2) AVDF_LDL2(2, 2, true); else AVDF_LDL2(2, 3, true); }
    else { if (a->n_streams == 1) AVDF_LDL2(2, 1, false); else if (a->n_streams == 2) AVDF_LDL2(2, 2, false); else AVDF_LDL2(2, 3, false); }
#undef AVDF_LDL2
    return check_launch("ln_dwconv_ln2_kernel");
  }
  // rows per warp tile: the largest of 8 / 4 / 2 that still gives (almost) every resident warp a tile - shorter tiles
  // re-normalise 2 halo rows per tile but shorten the serial per-warp chain of the small pyramid levels
  const long long warp_slots = (long long)sm_count() * 3 * LDL_WARPS;
  int rows = a->tile_rows;
  if (rows == 0) {
    rows = LDL_MAX_ROWS;
    while (rows > 2 && (long long)a->batch * ((p.t_out + rows - 1) / rows) * 4 < warp_slots * 3) rows /= 2;
  }
  p.rows = rows;
  const long long tiles = (long long)a->batch * ((p.t_out + rows - 1) / rows);
  const int grid = grid_for(tiles, LDL_WARPS, sm_count());
#define AVDF_LDL(S, NS)                                                                                              \
  AVDF_DISPATCH_DTYPE(a->out_dtype, OutT, {                                                                          \
    const size_t smem = (size_t)(NS * 9 + LDL_WARPS * LDL_DEPTH) * kC * sizeof(float);                               \
    AVDF_SMEM_ATTR_ONCE((ln_dwconv_ln_kernel<OutT, S, NS>), smem);                                                   \
    { cudaError_t le_ = launch_pdl(ln_dwconv_ln_kernel<OutT, S, NS>, grid, LDL_WARPS * 32, smem, st, p); if (le_ != cudaSuccess) { set_error("ln_dwconv_ln_kernel: launch failed: %s", cudaGetErrorString(le_)); return AVDF_ERR_CUDA; } } \
  })
  if (a->stride == 1) { if (a->n_streams == 1) AVDF_LDL(1, 1); else if (a->n_streams == 2) AVDF_LDL(1, 2); else AVDF_LDL(1, 3); }
  else { if (a->n_streams == 1) AVDF_LDL(2, 1); else if (a->n_streams == 2) AVDF_LDL(2, 2); else AVDF_LDL(2, 3); }
#undef AVDF_LDL
  return check_launch("ln_dwconv_ln_kernel");
}

extern "C" int avdf_attention(const void* q, const void* k, const void* v, const uint8_t* kv_mask, void* out,
                              int32_t in_dtype, int32_t out_dtype, int32_t batch, int32_t t, int32_t qkv_rows_per_video,
                              int32_t channels, int32_t n_head, int32_t window, void* stream) {
  AVDF_CHECK_ARG(q && k && v && out, "null pointer");
  AVDF_CHECK_ARG(channels == kC && n_head == 4, "attention supports 4 heads x 64 channels");
  AVDF_CHECK_ARG(batch >= 0 && t > 0, "bad sizes");
  AVDF_CHECK_ARG(window <= 1 || (window & 1), "window must be odd (or <= 1 for global attention)");
  AVDF_CHECK_DTYPE(in_dtype, "in_dtype");
  AVDF_CHECK_DTYPE(out_dtype, "out_dtype");
  const int rpv = qkv_rows_per_video > 0 ? qkv_rows_per_video : t;
  AVDF_CHECK_ARG(rpv >= t, "qkv_rows_per_video smaller than t");
  if (batch == 0) return AVDF_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (window == 7 && in_dtype != AVDF_DTYPE_F32) {   // 16-bit operands: tensor-core kernel (attention_mma.cu)
    static const bool use_mma = !(getenv("AVDF_ATT_MMA") && atoi(getenv("AVDF_ATT_MMA")) == 0);
    if (use_mma) return attention_banded_mma(q, k, v, kv_mask, out, in_dtype, out_dtype, batch, t, rpv, st);
  }
  if (window == 7) {                            // the window every shipped config uses: smem-tiled specialisation
    constexpr int HALF = 3;
    // query rows per CTA: 24 divides every level of the 768-row pyramid, and at 30 KB of K/V per CTA (16-bit) the
    // 32 x 32 tiles of a batch-32 level-0 launch are resident in ONE wave (7 CTAs per SM)
    static const int rows_env = getenv("AVDF_ATB_ROWS") ? atoi(getenv("AVDF_ATB_ROWS")) : 0;
    const int rows = rows_env > 0 && rows_env <= ATB_MAX_ROWS ? rows_env : 24;
    const int grid = batch * ((t + rows - 1) / rows);
    AVDF_DISPATCH_DTYPE(in_dtype, InT, AVDF_DISPATCH_DTYPE(out_dtype, OutT, {
      const size_t smem = 2 * (size_t)(rows + 2 * HALF) * kC * sizeof(InT);
      AVDF_SMEM_ATTR_ONCE((attention_banded_kernel<InT, OutT, HALF>), 2 * (size_t)(ATB_MAX_ROWS + 2 * HALF) * kC * sizeof(InT));
      attention_banded_kernel<InT, OutT, HALF><<<grid, ATB_WARPS * 32, smem, st>>>((const InT*)q, (const InT*)k, (const InT*)v, kv_mask,
                                                                                  (OutT*)out, batch, t, rpv, rows);
    }));
    return check_launch("attention_banded_kernel");
  }
  const long long tiles = (long long)batch * ((t + ATT_ROWS - 1) / ATT_ROWS);
  const int grid = grid_for(tiles, ATT_WARPS, sm_count());
  AVDF_DISPATCH_DTYPE(in_dtype, InT, AVDF_DISPATCH_DTYPE(out_dtype, OutT, (attention_kernel<InT, OutT><<<grid, ATT_WARPS * 32, 0, st>>>(
      (const InT*)q, (const InT*)k, (const InT*)v, kv_mask, (OutT*)out, batch, t, rpv, window))));
  return check_launch("attention_kernel");
}

extern "C" int avdf_ln_rows(const float* x, const float* w, const float* b, void* out, int32_t out_dtype, int64_t rows,
                            int32_t channels, void* stream) {
  AVDF_CHECK_ARG(x && w && b && out, "null pointer");
  AVDF_CHECK_ARG(channels == 256 || channels == 512 || channels == 1024, "channels must be 256, 512 or 1024");
  AVDF_CHECK_ARG(rows >= 0, "rows < 0");
  AVDF_CHECK_DTYPE(out_dtype, "out_dtype");
  if (rows == 0) return AVDF_OK;
  const int grid = grid_for(rows, 8 * LNR_ROWS, sm_count());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (channels == 256) AVDF_DISPATCH_DTYPE(out_dtype, OutT, (ln_rows_kernel<OutT, 1><<<grid, 256, 0, st>>>(x, w, b, (OutT*)out, rows)));
  else if (channels == 512) AVDF_DISPATCH_DTYPE(out_dtype, OutT, (ln_rows_kernel<OutT, 2><<<grid, 256, 0, st>>>(x, w, b, (OutT*)out, rows)));
  else AVDF_DISPATCH_DTYPE(out_dtype, OutT, (ln_rows_kernel<OutT, 4><<<grid, 256, 0, st>>>(x, w, b, (OutT*)out, rows)));
  return check_launch("ln_rows_kernel");
}

extern "C" int avdf_instnorm_lrelu(const float* x, void* out, int32_t out_dtype, int32_t batch, int32_t t, int32_t channels,
                                   float slope, void* stream) {
  AVDF_CHECK_ARG(x && out, "null pointer");
  AVDF_CHECK_ARG(batch >= 0 && t > 0 && channels > 0 && channels % 32 == 0, "channels must be a multiple of 32");
  AVDF_CHECK_DTYPE(out_dtype, "out_dtype");
  AVDF_CHECK_ARG(batch <= 65535, "batch too large for one launch");
  if (batch == 0) return AVDF_OK;
  dim3 grid(channels / 32, batch), block(32, IN_TY);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nr = (t + IN_TY - 1) / IN_TY;
#define AVDF_IN_REG(NR) AVDF_DISPATCH_DTYPE(out_dtype, OutT, (instnorm_lrelu_reg_kernel<OutT, NR><<<grid, block, 0, st>>>(x, (OutT*)out, t, channels, slope)))
  if (nr <= 3) AVDF_IN_REG(3);
  else if (nr <= 6) AVDF_IN_REG(6);
  else if (nr <= 12) AVDF_IN_REG(12);
  else if (nr <= 24) AVDF_IN_REG(24);
  else if (nr <= 48) AVDF_IN_REG(48);
  else AVDF_DISPATCH_DTYPE(out_dtype, OutT, (instnorm_lrelu_kernel<OutT><<<grid, block, 0, st>>>(x, (OutT*)out, t, channels, slope)));
#undef AVDF_IN_REG
  return check_launch("instnorm_lrelu_kernel");
}

static int fill_levels(int n_levels, const int32_t* level_len, int* off, int* len) {
  int P = 0;
  for (int l = 0; l < n_levels; ++l) { off[l] = P; len[l] = level_len[l]; P += level_len[l]; }
  return P;
}

extern "C" int avdf_fpn_fuse(const float* lat, const uint8_t* mask, const float* dw_w, const float* ln_w, const float* ln_b,
                             void* out, int32_t out_dtype, int32_t batch, int32_t channels, int32_t n_levels,
                             const int32_t* level_len, void* stream) {
  AVDF_CHECK_ARG(lat && dw_w && out && level_len, "null pointer");
  AVDF_CHECK_ARG((ln_w == nullptr) == (ln_b == nullptr), "ln_w / ln_b must come together");
  AVDF_CHECK_ARG(channels == kC, "channels must be 256");
  AVDF_CHECK_ARG(n_levels >= 1 && n_levels <= AVDF_MAX_LEVELS, "n_levels out of range");
  AVDF_CHECK_DTYPE(out_dtype, "out_dtype");
  for (int l = 0; l + 1 < n_levels; ++l) AVDF_CHECK_ARG(level_len[l] == 2 * level_len[l + 1], "levels must halve");
  FpnParams p{};
  p.lat = lat; p.mask = mask; p.dw_w = dw_w; p.ln_w = ln_w; p.ln_b = ln_b; p.out = out; p.B = batch; p.n_levels = n_levels;
  p.P = fill_levels(n_levels, level_len, p.lvl_off, p.lvl_len);
  if (batch == 0) return AVDF_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (level_len[0] % FPB == 0 && (FPB >> (n_levels - 1)) >= 1) {    // column-blocked kernel (every shipped config)
    int smem_rows = 0;
    for (int l = 0; l < AVDF_MAX_LEVELS; ++l) smem_rows += (FPB >> l) + 2;
    const size_t smem = (size_t)smem_rows * kC * sizeof(float);
    const int grid = batch * (level_len[0] / FPB);
    AVDF_DISPATCH_DTYPE(out_dtype, OutT, {
      AVDF_SMEM_ATTR_ONCE(fpn_fuse_cols_kernel<OutT>, smem);
      fpn_fuse_cols_kernel<OutT><<<grid, 256, smem, st>>>(p);
    });
    return check_launch("fpn_fuse_cols_kernel");
  }
  const int grid = grid_for((long long)batch * p.P, 8, sm_count());
  AVDF_DISPATCH_DTYPE(out_dtype, OutT, (fpn_fuse_kernel<OutT><<<grid, 256, 0, st>>>(p)));
  return check_launch("fpn_fuse_kernel");
}

extern "C" int avdf_head_final(const void* cls_feat, const void* reg_feat, int32_t dtype, const uint8_t* mask,
                               const float* cls_w, const float* cls_b, const float* reg_w, const float* reg_b,
                               const float* level_scale, float* logits, float* offsets, int32_t batch, int32_t channels,
                               int32_t n_levels, const int32_t* level_len, void* stream) {
  AVDF_CHECK_ARG(cls_feat && reg_feat && cls_w && cls_b && reg_w && reg_b && level_scale && logits && offsets && level_len, "null pointer");
  AVDF_CHECK_ARG(channels == kC, "channels must be 256");
  AVDF_CHECK_ARG(n_levels >= 1 && n_levels <= AVDF_MAX_LEVELS, "n_levels out of range");
  HeadParams p{};
  p.cls_feat = cls_feat; p.reg_feat = reg_feat; p.mask = mask; p.cls_w = cls_w; p.cls_b = cls_b; p.reg_w = reg_w; p.reg_b = reg_b;
  p.logits = logits; p.offsets = offsets; p.B = batch; p.n_levels = n_levels;
  p.P = fill_levels(n_levels, level_len, p.lvl_off, p.lvl_len);
  for (int l = 0; l < n_levels; ++l) p.lvl_scale[l] = level_scale[l];
  if (batch == 0) return AVDF_OK;
  const int grid = grid_for((long long)batch * p.P, 8, sm_count());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  AVDF_CHECK_DTYPE(dtype, "dtype");
  AVDF_DISPATCH_DTYPE(dtype, InT, (head_final_kernel<InT><<<grid, 256, 0, st>>>(p)));
  return check_launch("head_final_kernel");
}

// Heads from per-tap partial sums (avdf_conv_gemm dot_out of the last tower layers): one thread per pyramid row adds the
// three taps of its neighbours inside the level, bias, mask, Scale and ReLU (av_fd_no_recon.py:82-89, 152-159).
namespace avdf {
__global__ void __launch_bounds__(256) head_combine_kernel(const HeadParams p, const float* __restrict__ cd, const float* __restrict__ rd) {
  const long long rows = (long long)p.B * p.P;
  const float cb = p.cls_b[0], rb0 = p.reg_b[0], rb1 = p.reg_b[1];
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const int pr = (int)(r % p.P);
    int l = 0;
    while (l + 1 < p.n_levels && pr >= p.lvl_off[l + 1]) ++l;
    const int t = pr - p.lvl_off[l], T = p.lvl_len[l];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {                  // tap j multiplies the row at t + j - 1 (zero outside the level)
      const int tt = t + j - 1;
      if (tt >= 0 && tt < T) {
        const long long rr = r + (j - 1);
        a0 += cd[rr * 3 + j];
        a1 += rd[rr * 6 + j];
        a2 += rd[rr * 6 + 3 + j];
      }
    }
    const float mk = p.mask ? (p.mask[r] ? 1.f : 0.f) : 1.f;
    p.logits[r] = (a0 + cb) * mk;
    const float sc = p.lvl_scale[l];
    p.offsets[2 * r] = fmaxf(((a1 + rb0) * mk) * sc, 0.f);
    p.offsets[2 * r + 1] = fmaxf(((a2 + rb1) * mk) * sc, 0.f);
  }
}
}  // namespace avdf

extern "C" int avdf_head_combine(const float* cls_dots, const float* reg_dots, const uint8_t* mask, const float* cls_b,
                                 const float* reg_b, const float* level_scale, float* logits, float* offsets, int32_t batch,
                                 int32_t n_levels, const int32_t* level_len, void* stream) {
  AVDF_CHECK_ARG(cls_dots && reg_dots && cls_b && reg_b && level_scale && logits && offsets && level_len, "null pointer");
  AVDF_CHECK_ARG(n_levels >= 1 && n_levels <= AVDF_MAX_LEVELS, "n_levels out of range");
  HeadParams p{};
  p.mask = mask; p.cls_b = cls_b; p.reg_b = reg_b; p.logits = logits; p.offsets = offsets; p.B = batch; p.n_levels = n_levels;
  p.P = fill_levels(n_levels, level_len, p.lvl_off, p.lvl_len);
  for (int l = 0; l < n_levels; ++l) p.lvl_scale[l] = level_scale[l];
  if (batch == 0) return AVDF_OK;
  const long long rows = (long long)batch * p.P;
  const int grid = (int)((rows + 255) / 256);
  head_combine_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, cls_dots, reg_dots);
  return check_launch("head_combine_kernel");
}

extern "C" int avdf_vcls_exp12(const void* z, int32_t dtype, const float* conv0_wt, const float* lin1_wt, const float* ln_w,
                               const float* ln_b, const float* lin2_w, const float* lin2_b, float* out, int32_t batch,
                               int32_t t, int32_t channels, void* stream) {
  AVDF_CHECK_ARG(z && conv0_wt && lin1_wt && ln_w && ln_b && lin2_w && lin2_b && out, "null pointer");
  AVDF_CHECK_ARG(channels == kC, "channels must be 256");
  AVDF_CHECK_ARG(t > 0 && t <= VC12_MAXT, "t out of range (1..32)");
  AVDF_CHECK_DTYPE(dtype, "dtype");
  if (batch == 0) return AVDF_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  AVDF_DISPATCH_DTYPE(dtype, InT, (vcls_exp12_kernel<InT><<<batch, 256, 0, st>>>((const InT*)z, conv0_wt, lin1_wt, ln_w, ln_b, lin2_w, lin2_b, out, t)));
  return check_launch("vcls_exp12_kernel");
}

extern "C" int avdf_vcls_exp13(const void* z, int32_t dtype, const float* conv0_w, const float* seg_w, const float* seg_b,
                               const float* cls_w, const float* cls_b, float* out, int32_t batch, int32_t t,
                               int32_t channels, void* stream) {
  AVDF_CHECK_ARG(z && conv0_w && seg_w && seg_b && cls_w && cls_b && out, "null pointer");
  AVDF_CHECK_ARG(channels == 32 || channels == 64, "channels must be 32 or 64");
  AVDF_CHECK_ARG(t > 0, "t must be positive");
  AVDF_CHECK_DTYPE(dtype, "dtype");
  if (batch == 0) return AVDF_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (channels == 64) AVDF_DISPATCH_DTYPE(dtype, InT, (vcls_exp13_kernel<InT, 64><<<batch, 256, 0, st>>>((const InT*)z, conv0_w, seg_w, seg_b, cls_w, cls_b, out, t)));
  else AVDF_DISPATCH_DTYPE(dtype, InT, (vcls_exp13_kernel<InT, 32><<<batch, 256, 0, st>>>((const InT*)z, conv0_w, seg_w, seg_b, cls_w, cls_b, out, t)));
  return check_launch("vcls_exp13_kernel");
}
