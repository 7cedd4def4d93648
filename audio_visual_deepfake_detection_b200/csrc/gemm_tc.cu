// 16-bit tensor-core path of avdf_conv_gemm for sm_100a: TMA-fed, warp-specialised, persistent implicit-GEMM 1-D
// convolution on tcgen05 with the accumulator in TMEM and the whole MaskedConv1D -> (+bias) -> mask -> LayerNorm ->
// activation -> (+PE) -> gamma * x + residual epilogue fused.
//
// Tile: 128 output tokens x BN output channels, K stepped in 64-channel blocks per tap. 256 threads:
//   warp 0      TMA producer: A = 4-D box (64 ch, parity, TT steps, BB videos) of the token-major activation -- conv taps
//               are shifted boxes, zero padding is TMA out-of-bounds fill, stride 2 is the parity dimension;
//               W = 2-D box (64, BN) of the [N, taps*C] weights.
//   warp 1      MMA issuer: one lane issues tcgen05.mma (M=128, N=BN, K=16, bf16|fp16 -> fp32 in TMEM); tcgen05.commit
//               releases smem stages / publishes the accumulator.
//   warp 2      TMEM allocator.   warp 3 idle.
//   warps 4-7   epilogue, one TMEM lane quarter each (thread = output row): tcgen05.ld -> math in registers ->
//               swizzled smem tile -> TMA store; residual tiles arrive by TMA load. Global memory is touched by TMA only.
// Two configurations of the same kernel:
//   narrow (BN <= 128): 2 stages x (16 KB A + 16 KB W), 2 x 128 TMEM columns, ~103 KB smem, <= 128 registers
//                       -> TWO CTAs per SM. The K=256 GEMMs that dominate the launch count are latency-bound per tile
//                       (TMA -> MMA -> commit -> tcgen05.ld -> store is a chain of ~2 us); two independent tile
//                       pipelines per SM overlap those chains.
//   wide   (BN = 256):  3 stages x (16 KB A + 32 KB W), 2 x 256 TMEM columns, one CTA per SM (LayerNorm over the row
//                       needs all 256 channels in one CTA; the big-K embedding GEMMs are MMA-bound).
//   weight-stationary (BN = 256, 1x1, K <= 256): the CTA's 256 x K weight block (128 KB) is loaded ONCE and stays in
//                       shared memory; a CTA is pinned to one (segment, n-tile) group and walks its 128-row tiles, so
//                       only A (16 KB per K block, 3 stages) streams. The narrow configuration re-reads a 64 KB weight
//                       tile for every 128 x 128 output tile: 128 KB of L2 -> SM traffic per tile at ~80 GB/s per SM is
//                       1.6 us, the bound of the K = 256 launches (q/k/v, attention projection, FPN laterals).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include "gemm_common.cuh"
#include "tc_ptx.cuh"

namespace avdf {
namespace tc {

constexpr int BM = 128, BK = 64, MAX_BN = 256, MAX_STAGES = 4;
constexpr int A_STAGE = BM * BK * 2;          // 16384
constexpr int EPI_VEC_BYTES = 4 * MAX_BN * 4; // bias / ln_w / ln_b / gamma of the current n-tile
constexpr int STAGE_TILE_BYTES = 4 * 2 * 4096;   // per epilogue warp: a 4 KB result tile + a 4 KB residual / second result tile
constexpr int THREADS = 256;                  // 4 control warps + 4 epilogue warps
__host__ __device__ constexpr int n_stages_of(int bn) { return bn <= 128 ? 2 : 3; }
__host__ __device__ constexpr int b_stage_of(int bn) { return (bn <= 128 ? 128 : 256) * BK * 2; }
__host__ __device__ constexpr int smem_bytes_of(int bn) {
  return n_stages_of(bn) * (A_STAGE + b_stage_of(bn)) + 1024 /*barriers*/ + EPI_VEC_BYTES + STAGE_TILE_BYTES + 1024 /*align slack*/;
}
// weight-stationary: 2 x 16 KB of A in flight, 4 x 32 KB of resident W, EIGHT epilogue warps (a weight-stationary CTA is alone
// on its SM; with four, the epilogue - ~4.2 us per 128 x 256 tile - was the limit, scripts/gemm_ws_bench.py), each with its
// two 4 KB staging tiles; bias + gamma only (no fused LayerNorm). 231,936 of the 232,448 bytes a CTA can have: the
// dynamic shared-memory window must start 1024-aligned (it does: the first KB of an SM's shared memory is reserved).
// The 16-bit-output variants (the q/k/v projection) need only ONE 4 KB staging tile per warp, which pays for FOUR A stages:
// with two, the globaltimer stamps (scripts/gemm_epilogue_timeline.py) showed the epilogue idle 2.5 of every 4.0 us - the
// TMA -> MMA -> commit -> refill chain of a stage is ~1.2 us and a tile needs four of them.
constexpr int WS_W_BLOCKS = 4, WS_EPI_WARPS = 8;
__host__ __device__ constexpr int ws_a_stages(bool out16_only) { return out16_only ? 4 : 2; }
__host__ __device__ constexpr int ws_stage_bytes(bool out16_only) { return out16_only ? 4096 : 8192; }   // per epilogue warp
constexpr int WS_BAR_BYTES = 512, WS_VEC_BYTES = 2 * MAX_BN * 4;
constexpr int WS_SMEM_BYTES = ws_a_stages(false) * A_STAGE + WS_W_BLOCKS * MAX_BN * BK * 2 + WS_BAR_BYTES + WS_VEC_BYTES + WS_EPI_WARPS * ws_stage_bytes(false);
static_assert(WS_SMEM_BYTES == ws_a_stages(true) * A_STAGE + WS_W_BLOCKS * MAX_BN * BK * 2 + WS_BAR_BYTES + WS_VEC_BYTES + WS_EPI_WARPS * ws_stage_bytes(true), "both layouts use the same total");
constexpr int WS_THREADS = 128 + WS_EPI_WARPS * 32;

struct Params {
  CUtensorMap a_map[AVDF_MAX_LEVELS];
  CUtensorMap w_map;
  CUtensorMap w_half_map;                    // mc: box of bn / 2 weight rows (each CTA of a pair fetches one half)
  CUtensorMap o32_map[AVDF_MAX_LEVELS];      // fp32 output, 3-D (n, t, video) per segment, box = one epilogue warp's 32 rows x 32 columns
  CUtensorMap o16_map[AVDF_MAX_LEVELS];      // 16-bit output copy
  CUtensorMap o16w_map[AVDF_MAX_LEVELS];     // 16-bit output, 64-column box (32 rows x 128 B, swizzle 128B) for the wide epilogue pass
  CUtensorMap res_map[AVDF_MAX_LEVELS];      // residual (same geometry as the fp32 output)
  SegInfo seg;
  int seg_tile_start[AVDF_MAX_LEVELS + 1];   // prefix of m-tiles per level
  int seg_tt[AVDF_MAX_LEVELS];               // time steps per tile (power of two <= 128)
  int n_out, c_in, taps, stride, bn, n_tiles_n, n_tiles_m, total_tiles;
  int tap_tab[AVDF_MAX_TAPS];        // row offset of every tap (avdf_conv_gemm_args.tap_mode / tap_rows)
  int ws, ws_groups, ws_per;                 // weight-stationary: (segment, n-tile) groups, CTAs per group
  int mc;                                    // clusters of two CTAs on adjacent row tiles share every weight tile by TMA multicast
  int pair;                                  // wide configuration as CTA pairs: cta_group::2 MMAs (M = 256 over two SMs, each CTA holds half of W)
  unsigned idesc;
  EpiParams epi;
  unsigned long long* dbg;                   // optional per-CTA phase timestamps (globaltimer ns), 8 per CTA
};

// ---------------------------------------------------------------- debug timeline
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#ifdef AVDF_GEMM_TIMELINE
#define AVDF_TS(slot) do { if (p.dbg && lane == 0) p.dbg[blockIdx.x * 16 + (slot)] = gtime(); } while (0)
#else
#define AVDF_TS(slot) do { } while (0)
#endif

struct TileCoord { int seg, b0, t0, tt, n0; };
__device__ __forceinline__ TileCoord decode_tile(const Params& p, int tile) {
  TileCoord c;
  // m fastest: consecutive tiles of a persistent CTA (tile, tile + grid, ...) mostly share the weight tile and the
  // per-channel epilogue vectors
  const int nt = tile / p.n_tiles_m;
  const int mt = tile - nt * p.n_tiles_m;
  c.n0 = nt * p.bn;
  int s = 0;
  while (s + 1 < p.seg.n_seg && mt >= p.seg_tile_start[s + 1]) ++s;
  c.seg = s;
  c.tt = p.seg_tt[s];
  const int local = mt - p.seg_tile_start[s];
  const int tblocks = p.seg.t_out[s] / c.tt;
  const int bb = local / tblocks;
  c.t0 = (local - bb * tblocks) * c.tt;
  c.b0 = bb * (BM / c.tt);
  return c;
}

// MODE >= 0 fixes the epilogue variant at compile time (bit 0 LayerNorm, bits 1-2 activation, bit 3 residual,
// bit 4 positional encoding) so the epilogue carries no dead branches; MODE < 0 reads the flags at run time.
constexpr int W8_SMEM_BYTES_FWD = 223232;      // = W8_SMEM_BYTES (defined below, next to the kernel)
constexpr int mode_of(bool ln, int act, bool res, bool pe) { return (ln ? 1 : 0) | (act << 1) | (res ? 8 : 0) | (pe ? 16 : 0); }
// bit 5: the LayerNorm follows the residual step (avdf_conv_gemm_args.ln_after_residual): attention projection + LN2 in
// one launch. Only instantiated for the wide eight-epilogue-warp configuration with both outputs.
constexpr int MODE_POSTLN = mode_of(true, AVDF_ACT_NONE, true, false) | 32;
// bit 6: row dot products (avdf_conv_gemm_args.dot_*): the last layer of the head towers emits the per-tap partial sums of
// the heads' final convolution instead of its 256-channel fp32 output. Wide eight-warp configuration only.
constexpr int MODE_HEADDOT = mode_of(true, AVDF_ACT_RELU, false, false) | 64;
constexpr int MAX_DOTS = 6;
constexpr int W8D_SMEM_BYTES = W8_SMEM_BYTES_FWD + MAX_DOTS * MAX_BN * 4;

// OUTK >= 0 fixes which outputs exist: bit 0 fp32, bit 1 16-bit copy, bit 2 the 16-bit copy is fp16 (else bf16).
// CFG (compile-time, so that the streaming variants carry none of the other configurations' state):
//   0  streaming, four epilogue warps (narrow tiles: two CTAs per SM)
//   1  weight-stationary, eight epilogue warps
//   2  wide (BN = 256: fused LayerNorm, K >= 1024) with EIGHT epilogue warps: the two warps of a TMEM lane quarter split the
//      columns; LayerNorm statistics are exchanged through shared memory (one named barrier per row block). With four
//      warps the LayerNorm epilogue (two TMEM passes over 256 columns) took longer than the K = 768 main loop.
constexpr int W8_SMEM_BYTES = smem_bytes_of(MAX_BN) + 4 * 8192 /* four more staging pairs */ + 4096 /* LayerNorm partial sums */;
static_assert(W8_SMEM_BYTES == W8_SMEM_BYTES_FWD, "keep the forward declaration in step");
template <int MODE, int OUTK, int CFG = 0>
__global__ void __launch_bounds__(CFG ? WS_THREADS : THREADS, CFG ? 1 : 2) conv_gemm_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024 B alignment for the 128B swizzle atoms (an offset into the array keeps the shared address space visible
  // to the compiler: LDS/STS instead of generic loads)
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  constexpr bool ws = CFG == 1, w8 = CFG == 2 || CFG == 3;      // CFG 3: the wide configuration as CTA pairs (cta_group::2)
  constexpr bool pair = CFG == 3;
  constexpr bool OUT16_ONLY = MODE >= 0 && OUTK >= 0 && (OUTK & 3) == 2 && (MODE & 24) == 0;   // the "wide" epilogue pass
  const int n_stages = ws ? ws_a_stages(OUT16_ONLY) : (pair ? MAX_STAGES : n_stages_of(p.bn));
  const int b_stage = pair ? (p.bn >> 1) * BK * 2 : b_stage_of(p.bn);
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem + n_stages * A_STAGE;       // ws: WS_W_BLOCKS resident K blocks of the weights
  unsigned char* after = smem_b + (ws ? WS_W_BLOCKS : n_stages) * b_stage;
  // ws: the staging tiles come first (they need the 1024-byte alignment `after` has), then barriers and vectors
  uint64_t* bars = reinterpret_cast<uint64_t*>(ws ? after + WS_EPI_WARPS * ws_stage_bytes(OUT16_ONLY) : after);
  // bars: full[3], empty[3], tmem_full[2], tmem_empty[2], tmem ptr, residual[4]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  float* epi_smem = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + (ws ? WS_BAR_BYTES : 1024));
  unsigned char* stage_smem = ws ? after : after + 1024 + EPI_VEC_BYTES;
  constexpr int EPI_WARPS = CFG ? WS_EPI_WARPS : 4;
  unsigned char* part_smem = stage_smem + EPI_WARPS * 8192;       // w8: LayerNorm partial sums [tile parity][team][row] (float2)
  if (ws && (smem_u32(smem_dyn) & 1023u) != 0) {          // no slack for re-alignment in this configuration
    if (threadIdx.x == 0) printf("avdf gemm_tc: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + 2 + s); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * MAX_STAGES + 21);       // ws: the resident weight block has landed
  // i-th tile of this CTA (-1: none). Persistent launches stride the m-fastest tile list by the grid; weight-stationary
  // launches pin a CTA to one (segment, n-tile) group - one weight block - and stride that group's m-tiles.
  auto tile_at = [&](int i) -> int {
    if (!ws) { const int t_ = blockIdx.x + i * gridDim.x; return t_ < p.total_tiles ? t_ : -1; }
    const int g = blockIdx.x % p.ws_groups, j = blockIdx.x / p.ws_groups + i * p.ws_per;
    const int sg = g % p.seg.n_seg, nt = g / p.seg.n_seg;
    return j < p.seg_tile_start[sg + 1] - p.seg_tile_start[sg] ? nt * p.n_tiles_m + p.seg_tile_start[sg] + j : -1;
  };
  const int acc_cols = p.bn <= 128 ? 128 : 256;      // two accumulators: 256 or 512 TMEM columns per CTA
  const uint32_t tmem_cols = 2u * acc_cols;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) AVDF_TS(0);
  // CTA pairs (wide configuration only): the two CTAs of a cluster own adjacent row tiles of one n-tile; the leader (rank 0)
  // issues ONE tcgen05.mma.cta_group::2 per K step for both (M = 256), each CTA loads its own A rows and HALF of the weight
  // tile. Barriers the MMA warp waits on live in the leader, barriers it signals are reached in both CTAs by multicast commits
  // (the scheme of mlp_fused.cu). What it buys: a CTA writes and reads 32 instead of 48 KB of shared memory per K step - the
  // K >= 512 launches were bound by shared-memory bandwidth (TMA writes + operand reads), not by L2 (the multicast variant
  // that only halved the L2 reads did not move them).
  // A kernel that contains cta_group::2 instructions can only be launched as clusters (a plain launch fails with 'cluster
  // misconfiguration'), hence a separate instantiation (CFG 3). A stage then holds 16 + 16 KB instead of 16 + 32 KB: FOUR stages
  // instead of three in less shared memory - the big-K launches are bound by the latency of the TMA -> MMA -> commit -> refill
  // chain (~1.2 us) times the bytes in flight, and a pair needs half the bytes per K step.

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.seg.n_seg; ++s)
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.a_map[s]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.w_map) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), p.mc ? 2 : 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), pair ? 2 * EPI_WARPS : EPI_WARPS); }
    mbar_init(wfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (pair) {                                  // one warp of EACH CTA of the pair
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  // mc: the peer's barriers must exist before a multicast load or commit of ours can land on them
  const bool clustered = !ws && (p.mc != 0 || pair);
  const uint32_t crank = clustered ? cluster_ctarank() : 0u;
  if (clustered) cluster_sync_all();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp == 0) AVDF_TS(1);
  // programmatic dependent launch: everything above overlapped the previous kernel's tail; its outputs (A, residual) are
  // touched only from here on. Our own dependents may start their prologue right away.
  pdl_trigger();
  pdl_wait();

  const int kb_per_tap = p.c_in / BK;
  const int k_iters = p.taps * kb_per_tap;
  const uint32_t stage_bytes = (uint32_t)A_STAGE + (uint32_t)p.bn * BK * 2;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    // (the whole warp runs the loop and one elected lane issues: see elect_one() in tc_ptx.cuh)
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    for (int it = 0, tile; (tile = tile_at(it)) >= 0; ++it) {
      const TileCoord tc_ = decode_tile(p, tile);
      if (ws && it == 0 && leader) {             // the group's weight block: k_iters boxes of (64, bn), once
        mbar_arrive_expect_tx(wfull_bar, (uint32_t)(k_iters * p.bn * BK * 2));
        for (int kb = 0; kb < k_iters; ++kb)
          tma_load_2d(smem_u32(smem_b + kb * b_stage), &p.w_map, wfull_bar, kb * BK, tc_.n0 + p.seg.w_row[tc_.seg]);
      }
      for (int tap = 0; tap < p.taps; ++tap) {
        const int d = p.tap_tab[tap];
        int par = 0, dt = d;
        if (p.stride == 2) { par = d & 1; dt = (d - par) / 2; }
        for (int kb = 0; kb < kb_per_tap; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (pair) {
            // both CTAs' bytes (own A rows + own half of the weight tile) are credited to the LEADER's full barrier
            if (leader) {
              if (crank == 0) mbar_arrive_expect_tx(full_bar(stage), 2u * ((uint32_t)A_STAGE + (uint32_t)(p.bn >> 1) * 128u));
              const uint32_t fb = mapa_rank(full_bar(stage), 0u);
              tma_load_4d_pair(smem_u32(smem_a + stage * A_STAGE), &p.a_map[tc_.seg], fb, kb * BK, par, tc_.t0 + dt, tc_.b0);
              tma_load_2d_pair(smem_u32(smem_b + stage * b_stage), &p.w_half_map, fb, tap * p.c_in + kb * BK,
                               tc_.n0 + p.seg.w_row[tc_.seg] + (int)crank * (p.bn >> 1));
            }
          } else if (leader) {
            mbar_arrive_expect_tx(full_bar(stage), ws ? (uint32_t)A_STAGE : stage_bytes);
            tma_load_4d(smem_u32(smem_a + stage * A_STAGE), &p.a_map[tc_.seg], full_bar(stage), kb * BK, par, tc_.t0 + dt, tc_.b0);
            if (!ws && p.mc)             // our half of the weight tile, into both CTAs of the pair
              tma_load_2d_mc(smem_u32(smem_b + stage * b_stage) + crank * (uint32_t)(p.bn >> 1) * 128u, &p.w_half_map, full_bar(stage),
                             tap * p.c_in + kb * BK, tc_.n0 + p.seg.w_row[tc_.seg] + (int)crank * (p.bn >> 1), (unsigned short)3);
            else if (!ws) tma_load_2d(smem_u32(smem_b + stage * b_stage), &p.w_map, full_bar(stage), tap * p.c_in + kb * BK, tc_.n0 + p.seg.w_row[tc_.seg]);
          }
          __syncwarp();
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    // All 32 lanes run the loop (and wait on the barriers); ONE elected lane issues tcgen05.mma / commit from
    // warp-uniform code. Issued under `if (lane == 0)` every instruction was wrapped in a per-thread election loop
    // (~53 ns per issue measured, scripts/umma_pace.py) - longer than the 35 ns a 128x128x16 instruction takes.
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; (!pair || crank == 0) && tile_at(it) >= 0; ++it) {      // pair: the leader CTA issues for both
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if (ws && it == 0) mbar_wait(wfull_bar, 0);
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * acc_cols);
      for (int ki = 0; ki < k_iters; ++ki) {
        mbar_wait(full_bar(stage), phase);
        tcgen05_fence_after();
        if (leader) {
          const uint64_t da = make_sw128_desc(smem_u32(smem_a + stage * A_STAGE));
          const uint64_t db = make_sw128_desc(smem_u32(smem_b + (ws ? ki : stage) * b_stage));
          if (pair) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_pair(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (ki > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair(empty_bar(stage));
            if (ki == k_iters - 1) umma_commit_pair(tfull_bar(acc));
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)          // +32 B per K=16 step inside the swizzle atom
              umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (ki > 0 || k > 0) ? 1u : 0u);
            if (!ws && p.mc) umma_commit_mc(empty_bar(stage), (unsigned short)3); else umma_commit(empty_bar(stage));
            if (ki == k_iters - 1) umma_commit(tfull_bar(acc));
          }
        }
        __syncwarp();
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue: 4 warps, warp q owns TMEM lanes / tile rows 32q .. 32q+31
    // The math runs in the layout tcgen05.ld delivers (thread = output row, 32 consecutive columns per load):
    // per-row quantities (mask, LayerNorm mean / rstd) are plain registers, per-column vectors are broadcast
    // LDS.128. Results go to 128B/64B-swizzled smem tiles and leave with cp.async.bulk.tensor stores (3-D box =
    // this warp's 32 rows x 32 or 64 columns; rows beyond the batch are clipped by the tensor map); the residual
    // block of the next chunk is prefetched by TMA into a swizzled tile (mbarrier).
    const EpiParams& e = p.epi;
    // the accumulator is free again: tell the MMA warp (pair: the leader CTA's, through shared::cluster)
    auto release_acc = [&](int acc_) {
      if (pair) mbar_arrive_cluster(mapa_rank(tempty_bar(acc_), 0u)); else mbar_arrive(tempty_bar(acc_));
    };
    const bool ep_leader = elect_one();        // the lane that issues this warp's TMA loads / stores, commits and waits
    const int wi = warp - 4;                     // epilogue warp 0 .. EPI_WARPS - 1
    const int q = wi & 3;                        // TMEM lane quarter (= warp % 4) / 32-row block of the tile
    const int team = wi >> 2;                    // ws: two warps share a row block and split its columns
    const int N = p.n_out;
    const int et = threadIdx.x - 128;            // 0..127 among the epilogue threads
    float* s_bias = epi_smem; float* s_lnw = epi_smem + MAX_BN; float* s_lnb = epi_smem + 2 * MAX_BN; float* s_gam = epi_smem + (ws ? 1 : 3) * MAX_BN;
    if (ws) { s_lnw = s_bias; s_lnb = s_bias; }   // (never read: the weight-stationary configuration has no fused LayerNorm)
    constexpr int STG = ws ? ws_stage_bytes(OUT16_ONLY) : 8192;
    constexpr bool ONE_TILE = STG == 4096;                  // ws, 16-bit output: a single staging tile per warp
    unsigned char* t32 = stage_smem + wi * STG;             // result tile (fp32: swizzle 128B)
    unsigned char* trs = ONE_TILE ? t32 : t32 + 4096;       // residual tile / second result tile
    const uint32_t res_bar = bar_base + 8u * (2 * MAX_STAGES + 5 + wi);
    const uint32_t res_bar1 = bar_base + 8u * (2 * MAX_STAGES + 13 + wi);   // second residual tile (in-place fp32 path)
    const int chunks = (p.bn >> 5) / (EPI_WARPS / 4);     // 32-column chunks this warp handles: [ch0, ch1)
    const int ch0 = team * chunks, ch1 = ch0 + chunks;
    constexpr bool POSTLN = MODE >= 0 && (MODE & 32) != 0;
    constexpr bool HEADDOT = MODE >= 0 && (MODE & 64) != 0;
    static_assert(!HEADDOT || CFG == 2 || CFG == 3, "row dot products run in the wide eight-warp configuration");
    float* s_dot = reinterpret_cast<float*>(part_smem + 4096);        // HEADDOT: dot_w [dot_n][256]
    static_assert(!POSTLN || CFG == 2 || CFG == 3, "the post-residual LayerNorm runs in the wide eight-warp configuration");
    const bool has_ln = MODE < 0 ? (e.ln_w != nullptr) : ((MODE & 1) != 0);
    const bool has_res = MODE < 0 ? (e.residual != nullptr) : ((MODE & 8) != 0);
    const bool has_pe = MODE < 0 ? (e.pe != nullptr) : ((MODE & 16) != 0);
    const int act = MODE < 0 ? e.act : ((MODE >> 1) & 3);
    const bool has32 = OUTK < 0 ? (e.out_f32 != nullptr) : ((OUTK & 1) != 0);
    const bool has16 = OUTK < 0 ? (e.out_h != nullptr) : ((OUTK & 2) != 0);
    const bool o16_f16 = OUTK < 0 ? (e.out_h_f16 != 0) : ((OUTK & 4) != 0);
    // 16-bit result tile (32 columns: swizzle 64B): shares the fp32 tile when there is no fp32 output, else the residual tile
    unsigned char* t16 = has32 ? trs : t32;
    const int sw7 = lane & 7;                    // 128B swizzle: 16-byte chunk j of row `lane` lives in slot j ^ (lane & 7)
    const int sw3 = (lane >> 1) & 3;             // 64B swizzle: chunk j of row `lane` lives in slot j ^ ((lane >> 1) & 3)
    if (ep_leader) { mbar_init(res_bar, 1); mbar_init(res_bar1, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t res_phase = 0, res_phase1 = 0, store_seq = 0;
    int loaded_n0 = -1;
    // row mask of a tile's row owned by this thread (a global load: fetched one tile ahead so that its latency is
    // off the critical path of the tile's epilogue)
    auto row_mask_of = [&](int tile_) -> float {
      if (tile_ < 0 || !e.row_mask) return 1.f;
      const TileCoord c = decode_tile(p, tile_);
      const int r_ = q * 32 + lane;
      const int b_ = c.b0 + r_ / c.tt, t_ = c.t0 + (r_ & (c.tt - 1));
      if (b_ >= p.seg.batch) return 1.f;
      return e.row_mask[(size_t)b_ * p.seg.o_rows + p.seg.o_row[c.seg] + t_] ? 1.f : 0.f;
    };
    float mk_next = row_mask_of(tile_at(0));
    for (int it = 0, tile; (tile = tile_at(it)) >= 0; ++it) {
      const TileCoord tc_ = decode_tile(p, tile);
      const int vec0 = tc_.n0 + p.seg.w_row[tc_.seg];   // first entry of this tile's per-channel vectors
      if (vec0 != loaded_n0) {                   // per-channel epilogue vectors of this n-tile -> smem
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        for (int i = et; i < p.bn; i += EPI_WARPS * 32) {
          s_bias[i] = e.bias ? __ldg(e.bias + vec0 + i) : 0.f;
          if (!ws) {
            s_lnw[i] = e.ln_w ? __ldg(e.ln_w + vec0 + i) : 1.f;
            s_lnb[i] = e.ln_b ? __ldg(e.ln_b + vec0 + i) : 0.f;
          }
          s_gam[i] = e.gamma ? __ldg(e.gamma + vec0 + i) : 1.f;
        }
        if (HEADDOT)
          for (int i = et; i < e.dot_n * MAX_BN; i += EPI_WARPS * 32) s_dot[i] = __ldg(e.dot_w + i);
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        loaded_n0 = vec0;
      }
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      // my row (thread = row) and this warp's box origin inside the segment
      const int r = q * 32 + lane;
      const int b = tc_.b0 + r / tc_.tt;
      const int t = tc_.t0 + (r & (tc_.tt - 1));
      const bool valid = b < p.seg.batch;
      const int wb = tc_.b0 + (q * 32) / tc_.tt, wt = tc_.t0 + ((q * 32) & (tc_.tt - 1));   // box origin (video, time)
      const float mk = mk_next;
      mk_next = row_mask_of(tile_at(it + 1));     // in flight during this tile's epilogue
      auto fetch_residual = [&](int ch) {         // 32 rows x 128 B of the residual -> trs (TMA, swizzle 128B)
        if (ep_leader) {
          mbar_arrive_expect_tx(res_bar, 4096);
          tma_load_3d(smem_u32(trs), &p.res_map[tc_.seg], res_bar, tc_.n0 + ch * 32, wt, wb);
        }
      };
      // ---- residual + fp32 output only (attention projection, MLP down-projection): the residual block arrives by TMA in
      //      the very tile the result leaves from (updated in place by the thread that owns the row), so the two 4 KB
      //      tiles of this warp double-buffer the chunks: the residual of chunk c + 1 is in flight while chunk c is computed
      //      (with one residual tile its ~1 us load latency was exposed on every 32-column chunk: the top stall in ncu)
      constexpr bool RES32_OK = (MODE >= 0 && OUTK == 1 && (MODE & 8) != 0 && (MODE & 16) == 0) || POSTLN;
      auto fetch_residual_into = [&](int ch) {    // chunk ch -> tile ch & 1 (lane 0; the tile's previous store has been read)
        const uint32_t bar = (ch & 1) ? res_bar1 : res_bar;
        mbar_arrive_expect_tx(bar, 4096);
        tma_load_3d(smem_u32(t32 + ((ch & 1) << 12)), &p.res_map[tc_.seg], bar, tc_.n0 + ch * 32, wt, wb);
      };
      if (RES32_OK) {
        if (ep_leader) {
          tma_store_wait_read();                  // the previous tile's stores have read both tiles
          fetch_residual_into(ch0);
          if (chunks > 1) fetch_residual_into(ch0 + 1);
        }
      } else if (has_res) {
        fetch_residual(ch0);
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_cols);
      float mean = 0.f, rstd = 1.f;
      if constexpr (POSTLN) {
        // ---- attention projection + LN2: pass 1 builds the residual stream y = residual * mask + gamma * ((acc + bias) *
        //      mask) chunk by chunk in the tile its residual block arrived in (as above), sends it out (fp32), writes it
        //      BACK over the accumulator (tcgen05.st) and sums y, y^2; pass 2 re-reads y from TMEM, normalises and emits
        //      the 16-bit operand of the MLP through the wide 64-column boxes
        float s = 0.f, ss = 0.f;
        uint32_t vr[32];
        tmem_ld32_issue(taddr + ch0 * 32, vr);
        for (int ch = ch0; ch < ch1; ++ch) {
          tmem_ld_wait();
          float x[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(vr[i]);
          if (ch + 1 < ch1) tmem_ld32_issue(taddr + (ch + 1) * 32, vr);
          const int cl = ch * 32;
          unsigned char* tb = t32 + ((ch & 1) << 12);
          if (ch > ch0 && ch + 1 < ch1 && ep_leader) {
            tma_store_wait_read();
            fetch_residual_into(ch + 1);
          }
          if (ch & 1) { mbar_wait(res_bar1, res_phase1); res_phase1 ^= 1; }
          else { mbar_wait(res_bar, res_phase); res_phase ^= 1; }
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + cl);
          const float4* g4 = reinterpret_cast<const float4*>(s_gam + cl);
          uint32_t yb[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4* slot = reinterpret_cast<float4*>(tb + lane * 128 + ((j ^ sw7) << 4));
            const float4 rv = *slot;
            const float4 bb = b4[j], gg = g4[j];
            const float4 y = make_float4(fmaf(gg.x, (x[4 * j] + bb.x) * mk, rv.x * mk), fmaf(gg.y, (x[4 * j + 1] + bb.y) * mk, rv.y * mk),
                                         fmaf(gg.z, (x[4 * j + 2] + bb.z) * mk, rv.z * mk), fmaf(gg.w, (x[4 * j + 3] + bb.w) * mk, rv.w * mk));
            *slot = y;
            s += (y.x + y.y) + (y.z + y.w);
            ss = fmaf(y.x, y.x, ss); ss = fmaf(y.y, y.y, ss); ss = fmaf(y.z, y.z, ss); ss = fmaf(y.w, y.w, ss);
            yb[4 * j] = __float_as_uint(y.x); yb[4 * j + 1] = __float_as_uint(y.y); yb[4 * j + 2] = __float_as_uint(y.z); yb[4 * j + 3] = __float_as_uint(y.w);
          }
          tmem_st32(taddr + ch * 32, yb);
          fence_async_smem();
          __syncwarp();
          if (ep_leader) {
            tma_store_3d(&p.o32_map[tc_.seg], smem_u32(tb), tc_.n0 + cl, wt, wb);
            tma_store_commit();
          }
        }
        {                                         // the other half of the row's columns belongs to the partner warp
          float2* part = reinterpret_cast<float2*>(part_smem) + (it & 1) * 256;
          part[team * 128 + q * 32 + lane] = make_float2(s, ss);
          asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
          const float2 pa = part[q * 32 + lane], pb = part[128 + q * 32 + lane];
          s = pa.x + pb.x; ss = pa.y + pb.y;
        }
        mean = s / (float)p.bn;
        rstd = rsqrtf(fmaxf(ss / (float)p.bn - mean * mean, 0.f) + 1e-5f);
        tmem_st_wait();
        const f32x2 nmean2 = pk2(-mean), rstd2 = pk2(rstd);
        int tsel = 0;                             // pass 1 left its older store in tile 0, the younger one in tile 1
        for (int ch = ch0; ch < ch1; ch += 2) {
          uint32_t va[32], vb[32];
          tmem_ld32_issue(taddr + ch * 32, va);
          tmem_ld32_issue(taddr + (ch + 1) * 32, vb);
          tmem_ld_wait();
          if (ch + 2 >= ch1) {                     // all TMEM reads of this warp done: release the accumulator
            tcgen05_fence_before();
            __syncwarp();
            if (ep_leader) release_acc(acc);
          }
          unsigned char* tw = t32 + (tsel << 12);
          tsel ^= 1;
          if (ep_leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int cl = (ch + half) * 32;
            const float4* w4 = reinterpret_cast<const float4*>(s_lnw + cl);
            const float4* l4 = reinterpret_cast<const float4*>(s_lnb + cl);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t* vv = half == 0 ? va : vb;
              f32x2 y[4];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const float4 ww = w4[2 * j + u], ll = l4[2 * j + u];
                y[2 * u] = fma2(mul2(add2(pk2(__uint_as_float(vv[8 * j + 4 * u]), __uint_as_float(vv[8 * j + 4 * u + 1])), nmean2), rstd2), pk2(ww.x, ww.y), pk2(ll.x, ll.y));
                y[2 * u + 1] = fma2(mul2(add2(pk2(__uint_as_float(vv[8 * j + 4 * u + 2]), __uint_as_float(vv[8 * j + 4 * u + 3])), nmean2), rstd2), pk2(ww.z, ww.w), pk2(ll.z, ll.w));
              }
              uint4 uo;
              float f0, f1;
              if (o16_f16) {
                upk2(y[0], f0, f1); uo.x = pack_f16x2(f0, f1); upk2(y[1], f0, f1); uo.y = pack_f16x2(f0, f1);
                upk2(y[2], f0, f1); uo.z = pack_f16x2(f0, f1); upk2(y[3], f0, f1); uo.w = pack_f16x2(f0, f1);
              } else {
                upk2(y[0], f0, f1); uo.x = pack_bf16x2(f0, f1); upk2(y[1], f0, f1); uo.y = pack_bf16x2(f0, f1);
                upk2(y[2], f0, f1); uo.z = pack_bf16x2(f0, f1); upk2(y[3], f0, f1); uo.w = pack_bf16x2(f0, f1);
              }
              *reinterpret_cast<uint4*>(tw + lane * 128 + (((half * 4 + j) ^ sw7) << 4)) = uo;
            }
          }
          fence_async_smem();
          __syncwarp();
          if (ep_leader) {
            tma_store_3d(&p.o16w_map[tc_.seg], smem_u32(tw), tc_.n0 + ch * 32, wt, wb);
            tma_store_commit();
          }
        }
        continue;                                   // next tile
      }
      if (has_ln) {                               // row statistics over all bn columns (this thread owns the whole row)
        float s = 0.f, ss = 0.f;
        for (int ch = ch0; ch < ch1; ++ch) {
          float v[32];
          tmem_ld32(taddr + ch * 32, v);
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + ch * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[i];
            const float x0 = (v[4 * i] + bb.x) * mk, x1 = (v[4 * i + 1] + bb.y) * mk, x2 = (v[4 * i + 2] + bb.z) * mk, x3 = (v[4 * i + 3] + bb.w) * mk;
            s += (x0 + x1) + (x2 + x3);
            ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss); ss = fmaf(x2, x2, ss); ss = fmaf(x3, x3, ss);
          }
        }
        if (w8) {                                 // the other half of the row's columns belongs to the partner warp
          float2* part = reinterpret_cast<float2*>(part_smem) + (it & 1) * 256;
          part[team * 128 + q * 32 + lane] = make_float2(s, ss);
          asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
          const float2 pa = part[q * 32 + lane], pb = part[128 + q * 32 + lane];
          s = pa.x + pb.x; ss = pa.y + pb.y;
        }
        mean = s / (float)p.bn;
        const float var = fmaxf(ss / (float)p.bn - mean * mean, 0.f);
        rstd = rsqrtf(var + 1e-5f);
      }
      // ---- wide pass (16-bit output only, no residual / PE): 64 columns per step halve the per-step fixed latencies
      //      (TMEM wait, proxy fence, warp sync, TMA issue, store-read wait); the two 4 KB tiles alternate
      constexpr bool WIDE_OK = MODE >= 0 && OUTK >= 0 && (OUTK & 3) == 2 && (MODE & 24) == 0;
      if (WIDE_OK && (chunks & 1) == 0) {
        for (int ch = ch0; ch < ch1; ch += 2) {
          uint32_t va[32], vb[32];
          if (wi == 0 && it == 0) AVDF_TS(2 + 6 * ((ch - ch0) >> 1 & 1));       // accumulator ready, step starts
          tmem_ld32_issue(taddr + ch * 32, va);
          tmem_ld32_issue(taddr + (ch + 1) * 32, vb);
          tmem_ld_wait();
          if (wi == 0 && it == 0) AVDF_TS(3 + 6 * ((ch - ch0) >> 1 & 1));       // TMEM read done
          if (ch + 2 >= ch1) {                     // all TMEM reads of this warp done: release the accumulator
            tcgen05_fence_before();
            __syncwarp();
            if (ep_leader) release_acc(acc);
          }
          unsigned char* tw = (store_seq++ & 1) ? trs : t32;
          const f32x2 mk2 = pk2(mk), nmean2 = pk2(-mean), rstd2 = pk2(rstd);
          if (ep_leader) {
            if (ONE_TILE) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          }
          __syncwarp();
          if (wi == 0 && it == 0) AVDF_TS(4 + 6 * ((ch - ch0) >> 1 & 1));       // staging tile free
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int cl = (ch + half) * 32;
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + cl);
            const float4* w4 = reinterpret_cast<const float4*>(s_lnw + cl);
            const float4* l4 = reinterpret_cast<const float4*>(s_lnb + cl);
#pragma unroll
            for (int j = 0; j < 4; ++j) {          // 8 columns -> one 16-byte chunk of the 128 B row; packed fp32 pairs
              const uint32_t* vv = half == 0 ? va : vb;
              f32x2 y[4];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const float4 bb = b4[2 * j + u];
                f32x2 xa = mul2(add2(pk2(__uint_as_float(vv[8 * j + 4 * u]), __uint_as_float(vv[8 * j + 4 * u + 1])), pk2(bb.x, bb.y)), mk2);
                f32x2 xb = mul2(add2(pk2(__uint_as_float(vv[8 * j + 4 * u + 2]), __uint_as_float(vv[8 * j + 4 * u + 3])), pk2(bb.z, bb.w)), mk2);
                if (has_ln) {
                  const float4 ww = w4[2 * j + u], ll = l4[2 * j + u];
                  xa = fma2(mul2(add2(xa, nmean2), rstd2), pk2(ww.x, ww.y), pk2(ll.x, ll.y));
                  xb = fma2(mul2(add2(xb, nmean2), rstd2), pk2(ww.z, ww.w), pk2(ll.z, ll.w));
                }
                y[2 * u] = act_tc2(xa, act); y[2 * u + 1] = act_tc2(xb, act);
              }
              uint4 uo;
              float f0, f1;
              if (o16_f16) {
                upk2(y[0], f0, f1); uo.x = pack_f16x2(f0, f1); upk2(y[1], f0, f1); uo.y = pack_f16x2(f0, f1);
                upk2(y[2], f0, f1); uo.z = pack_f16x2(f0, f1); upk2(y[3], f0, f1); uo.w = pack_f16x2(f0, f1);
              } else {
                upk2(y[0], f0, f1); uo.x = pack_bf16x2(f0, f1); upk2(y[1], f0, f1); uo.y = pack_bf16x2(f0, f1);
                upk2(y[2], f0, f1); uo.z = pack_bf16x2(f0, f1); upk2(y[3], f0, f1); uo.w = pack_bf16x2(f0, f1);
              }
              *reinterpret_cast<uint4*>(tw + lane * 128 + (((half * 4 + j) ^ sw7) << 4)) = uo;
            }
          }
          if (wi == 0 && it == 0) AVDF_TS(5 + 6 * ((ch - ch0) >> 1 & 1));       // math + STS issued
          fence_async_smem();
          __syncwarp();
          if (wi == 0 && it == 0) AVDF_TS(6 + 6 * ((ch - ch0) >> 1 & 1));       // proxy fence done
          if (ep_leader) {
            tma_store_3d(&p.o16w_map[tc_.seg], smem_u32(tw), tc_.n0 + ch * 32, wt, wb);
            tma_store_commit();
          }
          if (wi == 0 && it == 0) AVDF_TS(7 + 6 * ((ch - ch0) >> 1 & 1));       // TMA store issued
        }
        if (wi == 0 && it == 1) AVDF_TS(14);        // second tile's epilogue starts
        continue;                                   // next tile
      }
      float dots[MAX_DOTS];
#pragma unroll
      for (int j = 0; j < MAX_DOTS; ++j) dots[j] = 0.f;
      uint32_t vr[32];                            // accumulator block of the current chunk (raw bits)
      tmem_ld32_issue(taddr + ch0 * 32, vr);
      for (int ch = ch0; ch < ch1; ++ch) {
        tmem_ld_wait();
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(vr[i]);
        if (ch + 1 < ch1) {
          tmem_ld32_issue(taddr + (ch + 1) * 32, vr);   // next block's TMEM read overlaps this block's math
        } else {                                  // all TMEM reads of this warp are issued: once they complete the
          tcgen05_fence_before();                 // MMA warp may overwrite the accumulator
          __syncwarp();
          if (ep_leader) release_acc(acc);
        }
        const int cl = ch * 32;
        {                                         // (acc + bias) * mask -> LayerNorm -> activation
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + cl);
          const float4* w4 = reinterpret_cast<const float4*>(s_lnw + cl);
          const float4* l4 = reinterpret_cast<const float4*>(s_lnb + cl);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = b4[j];
            x[4 * j] = (x[4 * j] + bb.x) * mk; x[4 * j + 1] = (x[4 * j + 1] + bb.y) * mk;
            x[4 * j + 2] = (x[4 * j + 2] + bb.z) * mk; x[4 * j + 3] = (x[4 * j + 3] + bb.w) * mk;
            if (has_ln) {
              const float4 ww = w4[j], ll = l4[j];
              x[4 * j] = fmaf((x[4 * j] - mean) * rstd, ww.x, ll.x); x[4 * j + 1] = fmaf((x[4 * j + 1] - mean) * rstd, ww.y, ll.y);
              x[4 * j + 2] = fmaf((x[4 * j + 2] - mean) * rstd, ww.z, ll.z); x[4 * j + 3] = fmaf((x[4 * j + 3] - mean) * rstd, ww.w, ll.w);
            }
            x[4 * j] = act_tc(x[4 * j], act); x[4 * j + 1] = act_tc(x[4 * j + 1], act);
            x[4 * j + 2] = act_tc(x[4 * j + 2], act); x[4 * j + 3] = act_tc(x[4 * j + 3], act);
          }
        }
        if (HEADDOT) {                            // this row's partial dot products over the chunk's 32 channels
#pragma unroll
          for (int j = 0; j < MAX_DOTS; ++j) {
            if (j < e.dot_n) {
              const float4* d4 = reinterpret_cast<const float4*>(s_dot + j * MAX_BN + cl);
              float a_ = dots[j];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 ww = d4[i];
                a_ = fmaf(ww.x, x[4 * i], a_); a_ = fmaf(ww.y, x[4 * i + 1], a_); a_ = fmaf(ww.z, x[4 * i + 2], a_); a_ = fmaf(ww.w, x[4 * i + 3], a_);
              }
              dots[j] = a_;
            }
          }
        }
        if (has_pe && valid) {                    // + PE[t, n] * mask (embedding only: one launch per pass)
          const float4* pe4 = reinterpret_cast<const float4*>(e.pe + (size_t)t * N + tc_.n0 + cl);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 pv = __ldg(pe4 + j);
            x[4 * j] = fmaf(pv.x, mk, x[4 * j]); x[4 * j + 1] = fmaf(pv.y, mk, x[4 * j + 1]);
            x[4 * j + 2] = fmaf(pv.z, mk, x[4 * j + 2]); x[4 * j + 3] = fmaf(pv.w, mk, x[4 * j + 3]);
          }
        }
        if (RES32_OK) {
          unsigned char* tb = t32 + ((ch & 1) << 12);
          if (ch > ch0 && ch + 1 < ch1 && ep_leader) {   // tile (ch + 1) & 1 held chunk ch - 1: its store was committed one
            tma_store_wait_read();                         // iteration ago; refill it with the residual of chunk ch + 1
            fetch_residual_into(ch + 1);
          }
          if (ch & 1) { mbar_wait(res_bar1, res_phase1); res_phase1 ^= 1; }
          else { mbar_wait(res_bar, res_phase); res_phase ^= 1; }
          const float4* g4 = reinterpret_cast<const float4*>(s_gam + cl);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4* slot = reinterpret_cast<float4*>(tb + lane * 128 + ((j ^ sw7) << 4));
            const float4 rv = *slot;
            const float4 gg = g4[j];
            *slot = make_float4(fmaf(gg.x, x[4 * j], rv.x * mk), fmaf(gg.y, x[4 * j + 1], rv.y * mk),
                                fmaf(gg.z, x[4 * j + 2], rv.z * mk), fmaf(gg.w, x[4 * j + 3], rv.w * mk));
          }
          fence_async_smem();
          __syncwarp();
          if (ep_leader) {
            tma_store_3d(&p.o32_map[tc_.seg], smem_u32(tb), tc_.n0 + cl, wt, wb);
            tma_store_commit();
          }
          continue;                                 // next chunk
        }
        if (has_res) {                            // residual * mask + gamma * x
          mbar_wait(res_bar, res_phase);
          res_phase ^= 1;
          const float4* g4 = reinterpret_cast<const float4*>(s_gam + cl);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 rv = *reinterpret_cast<const float4*>(trs + lane * 128 + ((j ^ sw7) << 4));
            const float4 gg = g4[j];
            x[4 * j] = fmaf(gg.x, x[4 * j], rv.x * mk); x[4 * j + 1] = fmaf(gg.y, x[4 * j + 1], rv.y * mk);
            x[4 * j + 2] = fmaf(gg.z, x[4 * j + 2], rv.z * mk); x[4 * j + 3] = fmaf(gg.w, x[4 * j + 3], rv.w * mk);
          }
          __syncwarp();                           // every lane has read the residual tile
        }
        // the previous chunk's TMA stores must have finished READING the tiles before they are overwritten
        unsigned char* t16c = t16;
        if (!has32) {                               // 16-bit only: alternate halves of the 4 KB tile, one store may stay in flight
          t16c = t32 + ((store_seq++ & 1) << 11);
          if (ep_leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        } else {
          if (ep_leader) tma_store_wait_read();
        }
        __syncwarp();
        if (has32) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(t32 + lane * 128 + ((j ^ sw7) << 4)) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
        }
        if (has16) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            if (o16_f16) {
              u.x = pack_f16x2(x[8 * j], x[8 * j + 1]); u.y = pack_f16x2(x[8 * j + 2], x[8 * j + 3]);
              u.z = pack_f16x2(x[8 * j + 4], x[8 * j + 5]); u.w = pack_f16x2(x[8 * j + 6], x[8 * j + 7]);
            } else {
              u.x = pack_bf16x2(x[8 * j], x[8 * j + 1]); u.y = pack_bf16x2(x[8 * j + 2], x[8 * j + 3]);
              u.z = pack_bf16x2(x[8 * j + 4], x[8 * j + 5]); u.w = pack_bf16x2(x[8 * j + 6], x[8 * j + 7]);
            }
            *reinterpret_cast<uint4*>(t16c + lane * 64 + ((j ^ sw3) << 4)) = u;
          }
        }
        fence_async_smem();                       // generic-proxy smem writes -> visible to the async proxy (TMA)
        __syncwarp();
        if (ep_leader) {
          if (has32) tma_store_3d(&p.o32_map[tc_.seg], smem_u32(t32), tc_.n0 + cl, wt, wb);
          if (has16) tma_store_3d(&p.o16_map[tc_.seg], smem_u32(t16c), tc_.n0 + cl, wt, wb);
          tma_store_commit();
        }
        if (has_res && ch + 1 < ch1) {         // next chunk's residual; if the 16-bit tile aliases the residual
          if (has16 && has32) {                   // tile, its store must have read it first
            if (ep_leader) tma_store_wait_read();
            __syncwarp();
          }
          fetch_residual(ch + 1);
        }
      }
      if (has_res && has16 && has32) {            // before the next tile's first residual prefetch
        if (ep_leader) tma_store_wait_read();
        __syncwarp();
      }
      if (HEADDOT) {
        // the two warps of a row quarter hold the two column halves: team 0 stores its partial sums, team 1 adds to them
        // (two addends: the order does not matter, the result is deterministic)
        float* drow = e.dot_out + ((size_t)b * p.seg.o_rows + p.seg.o_row[tc_.seg] + t) * e.dot_n;
        if (team == 0 && valid)
          for (int j = 0; j < e.dot_n; ++j) drow[j] = dots[j];
        __threadfence_block();
        asm volatile("bar.sync %0, 64;" ::"r"(6 + q) : "memory");
        if (team == 1 && valid)
          for (int j = 0; j < e.dot_n; ++j) atomicAdd(drow + j, dots[j]);
      }
    }
    if (ep_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before exit
  }
  tcgen05_fence_before();
  __syncthreads();
  if (clustered) cluster_sync_all();        // no CTA leaves while its peer can still write into its shared memory / barriers / TMEM
  if (warp == 2) {
    tcgen05_fence_after();
    if (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

}  // namespace tc

static unsigned long long* g_dbg = nullptr;
static bool g_w8 = !(getenv("AVDF_GEMM_W8") && atoi(getenv("AVDF_GEMM_W8")) == 0);   // wide tiles: eight epilogue warps (0: four)
static int g_ws_mode = -1;                    // weight-stationary configuration: -1 auto (default), 0 never, 1 wherever it is legal
// weight multicast between CTA pairs: smallest K (taps * c_in) it is used for; 0 switches it off (AVDF_GEMM_MC)
// (measured: no gain - these launches are bound by shared-memory bandwidth, not by L2 - so it is off by default)
static int g_mc_min_k = getenv("AVDF_GEMM_MC") ? atoi(getenv("AVDF_GEMM_MC")) : 0;
// CTA pairs (cta_group::2) for wide-configuration launches with taps * c_in >= this; 0 switches them off (AVDF_GEMM_PAIR)
static int g_pair_min_k = getenv("AVDF_GEMM_PAIR") ? atoi(getenv("AVDF_GEMM_PAIR")) : 512;

int conv_gemm_tc(const avdf_conv_gemm_args* a, cudaStream_t st) {
  using namespace tc;
  AVDF_CHECK_ARG(a->c_in % BK == 0, "bf16 path: c_in must be a multiple of 64");
  AVDF_CHECK_ARG(a->n_out % 32 == 0, "bf16 path: n_out must be a multiple of 32");
  // N tile: 256 only where LayerNorm needs the whole row in one CTA; everything else uses the narrow configuration
  // (BN <= 128, two CTAs per SM)
  int bn = a->ln_w ? (a->n_out >= MAX_BN ? MAX_BN : a->n_out) : (a->n_out % 128 == 0 ? 128 : (a->n_out > MAX_BN ? MAX_BN : a->n_out));
  // long-K launches (the MLP down-projection, K = 1024) are bound by the L2 -> SM operand traffic ((BM + bn) K bytes per
  // tile): one 256-wide tile per row block reads the activations once instead of twice (measured 29.2 -> 26.7 us)
  if (!a->ln_w && a->taps * a->c_in >= 1024 && a->n_out % 256 == 0) bn = 256;
  // weight-stationary: 1x1, K <= 256, 256-wide n-tiles, every segment with the same number of m-tiles (a CTA is
  // pinned to one (segment, n-tile) group). Worth it when a CTA gets to reuse its weight block over several tiles.
  const int sms = device_sm_count();
  bool ws = g_ws_mode != 0 && a->taps == 1 && a->stride == 1 && a->c_in <= WS_W_BLOCKS * BK && a->n_out % MAX_BN == 0 && a->n_seg >= 1 && !a->ln_w;
  int ws_groups = 0, ws_per = 0;
  if (ws) {
    int per_seg = -1;
    for (int s = 0; s < a->n_seg; ++s) {
      const int T = a->seg_t_out[s];
      int tt = T & (-T);
      if (tt > BM) tt = BM;
      const int n = T > 0 ? (T / tt) * ceil_div(a->batch, BM / tt) : 0;
      if (per_seg < 0) per_seg = n; else if (n != per_seg) ws = false;
    }
    ws_groups = a->n_seg * (a->n_out / MAX_BN);
    ws_per = sms / ws_groups < per_seg ? sms / ws_groups : per_seg;
    // auto: only where it was measured to win (scripts/gemm_ws_bench.py: the stacked q/k/v projection at level 0, 14.8 vs
    // 16.7 us) - 16-bit output without activation / residual (the instantiated fast variant), three or more weight groups
    // and >= 3.5 tiles per CTA; everywhere else the two co-resident CTAs of the narrow configuration are as fast or faster
    const bool qkv_like = a->out_h && !a->out_f32 && a->out_h_dtype == AVDF_DTYPE_F16 && a->act == AVDF_ACT_NONE && !a->residual && !a->pe;
    if (ws_per < 1 || (g_ws_mode < 0 && !(qkv_like && ws_groups >= 3 && per_seg * 2 >= ws_per * 7))) ws = false;
  }
  if (ws) bn = MAX_BN;
  AVDF_CHECK_ARG(a->n_out % bn == 0, "bf16 path: n_out must be <= 256 or a multiple of 256");
  AVDF_CHECK_ARG(bn % 16 == 0 && bn >= 32, "bf16 path: unsupported n_out");
  AVDF_CHECK_ARG(!a->ln_w || a->n_out == bn, "bf16 path: fused LayerNorm needs n_out <= 256");
  AVDF_CHECK_ARG(a->stride == 1 || a->taps == 3 || a->taps == 1, "bad stride/taps");
  AVDF_CHECK_ARG(!a->tap_rows || a->stride == 1, "tap_rows needs stride 1");
  AVDF_CHECK_ARG((reinterpret_cast<uintptr_t>(a->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0, "operands must be 16-byte aligned");
  EncodeFn encode = get_encode();
  if (!encode) { set_error("avdf_conv_gemm: cuTensorMapEncodeTiled not available from the driver"); return AVDF_ERR_CUDA; }

  const bool f16 = a->dtype == AVDF_DTYPE_F16;
  const CUtensorMapDataType tm_dtype = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  static thread_local Params p;     // large (tensor maps); filled per call, copied into the launch by value
  memset(&p, 0, sizeof(p));
  fill_seg(a, p.seg);
  fill_epi(a, p.epi);
  p.dbg = g_dbg;
  p.n_out = a->n_out; p.c_in = a->c_in; p.taps = a->taps; p.stride = a->stride; p.bn = bn;
  fill_taps(a, p.tap_tab);
  p.n_tiles_n = a->n_out / bn;
  int tiles = 0;
  for (int s = 0; s < a->n_seg; ++s) {
    const int T = a->seg_t_out[s];
    int tt = T & (-T);
    if (tt > BM) tt = BM;
    p.seg_tt[s] = tt;
    const int bb = BM / tt;
    p.seg_tile_start[s] = tiles;
    tiles += (T / tt) * ceil_div(a->batch, bb);
    // A view of this level: dims (c, parity, t, b)
    const cuuint64_t t_in = (cuuint64_t)T * a->stride;
    cuuint64_t dims[4] = {(cuuint64_t)a->c_in, (cuuint64_t)a->stride, (cuuint64_t)T, (cuuint64_t)a->batch};
    cuuint64_t strides[3] = {(cuuint64_t)a->c_in * 2, (cuuint64_t)a->c_in * 2 * a->stride, (cuuint64_t)a->a_rows_per_video * a->c_in * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, 1u, (cuuint32_t)tt, (cuuint32_t)bb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    (void)t_in;
    void* base = const_cast<unsigned char*>(reinterpret_cast<const unsigned char*>(a->a)) + (size_t)a->seg_a_row[s] * a->c_in * 2;
    CUresult r = encode(&p.a_map[s], tm_dtype, 4, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("avdf_conv_gemm: cuTensorMapEncodeTiled(A, level %d) failed with %d", s, (int)r); return AVDF_ERR_CUDA; }
    // epilogue maps of this segment: 3-D (n, t, video); box = 32 columns x the 32 rows one epilogue warp owns
    {
      const int tw = tt < 32 ? tt : 32, bw = 32 / tw;
      struct { CUtensorMap* map; const void* base; CUtensorMapDataType dt; int es; CUtensorMapSwizzle sw; } outs[3] = {
          {&p.o32_map[s], a->out_f32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, CU_TENSOR_MAP_SWIZZLE_128B},
          {&p.o16_map[s], a->out_h, a->out_h_dtype == AVDF_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_64B},
          {&p.res_map[s], a->residual, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, CU_TENSOR_MAP_SWIZZLE_128B}};
      if (a->out_h && (!a->out_f32 || a->ln_after_residual) && bn % 64 == 0) {     // wide 16-bit box
        cuuint64_t odims[3] = {(cuuint64_t)a->n_out, (cuuint64_t)T, (cuuint64_t)a->batch};
        cuuint64_t ostr[2] = {(cuuint64_t)a->n_out * 2, (cuuint64_t)a->o_rows_per_video * a->n_out * 2};
        cuuint32_t obox[3] = {64u, (cuuint32_t)tw, (cuuint32_t)bw};
        cuuint32_t oes[3] = {1, 1, 1};
        void* obase = const_cast<unsigned char*>(reinterpret_cast<const unsigned char*>(a->out_h)) + (size_t)a->seg_o_row[s] * a->n_out * 2;
        CUresult ro = encode(&p.o16w_map[s], a->out_h_dtype == AVDF_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                             obase, odims, ostr, obox, oes, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (ro != CUDA_SUCCESS) { set_error("avdf_conv_gemm: cuTensorMapEncodeTiled(wide output, level %d) failed with %d", s, (int)ro); return AVDF_ERR_CUDA; }
      }
      for (int o = 0; o < 3; ++o) {
        if (!outs[o].base) continue;
        AVDF_CHECK_ARG((reinterpret_cast<uintptr_t>(outs[o].base) & 15) == 0, "outputs / residual must be 16-byte aligned");
        cuuint64_t odims[3] = {(cuuint64_t)a->n_out, (cuuint64_t)T, (cuuint64_t)a->batch};
        cuuint64_t ostr[2] = {(cuuint64_t)a->n_out * outs[o].es, (cuuint64_t)a->o_rows_per_video * a->n_out * outs[o].es};
        cuuint32_t obox[3] = {32u, (cuuint32_t)tw, (cuuint32_t)bw};
        cuuint32_t oes[3] = {1, 1, 1};
        void* obase = const_cast<unsigned char*>(reinterpret_cast<const unsigned char*>(outs[o].base)) + (size_t)a->seg_o_row[s] * a->n_out * outs[o].es;
        CUresult ro = encode(outs[o].map, outs[o].dt, 3, obase, odims, ostr, obox, oes, CU_TENSOR_MAP_INTERLEAVE_NONE, outs[o].sw,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (ro != CUDA_SUCCESS) { set_error("avdf_conv_gemm: cuTensorMapEncodeTiled(output %d, level %d) failed with %d", o, s, (int)ro); return AVDF_ERR_CUDA; }
      }
    }
  }
  p.seg_tile_start[a->n_seg] = tiles;
  p.n_tiles_m = tiles;
  p.total_tiles = tiles * p.n_tiles_n;
  if (ws) { p.ws = 1; p.ws_groups = ws_groups; p.ws_per = ws_per; }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a->taps * a->c_in, (cuuint64_t)(a->n_w_rows > 0 ? a->n_w_rows : a->n_out)};
    cuuint64_t strides[1] = {(cuuint64_t)a->taps * a->c_in * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.w_map, tm_dtype, 2, const_cast<void*>(a->w), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("avdf_conv_gemm: cuTensorMapEncodeTiled(W) failed with %d", (int)r); return AVDF_ERR_CUDA; }
  }
  // instruction descriptor: D=f32 (bits 4-5 = 1), A/B format at bits 7-9 / 10-12 (0 = f16, 1 = bf16), K-major both,
  // N>>3 at 17, M>>4 at 24
  const unsigned fmt = f16 ? 0u : 1u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(bn >> 3) << 17) | ((unsigned)(BM >> 4) << 24);
  if (p.total_tiles == 0) return AVDF_OK;

  static DeviceOnce attrs_once;          // per device; marked only after every attribute call succeeded
  const int cur_dev = current_device();
  if (!attrs_once.done(cur_dev)) {
    // (mode, output kind) pairs the inference path uses get their own instantiation; anything else runs the generic one
#define AVDF_TC_VARIANTS(X)                                                                                   \
    X(mode_of(false, AVDF_ACT_NONE, false, false), 1) X(mode_of(false, AVDF_ACT_NONE, false, false), 6)           \
    X(mode_of(false, AVDF_ACT_NONE, false, false), 2) X(mode_of(false, AVDF_ACT_NONE, true, false), 1)            \
    X(mode_of(false, AVDF_ACT_NONE, true, false), 7) X(mode_of(false, AVDF_ACT_NONE, true, false), 3)             \
    X(mode_of(false, AVDF_ACT_GELU, false, false), 6) X(mode_of(false, AVDF_ACT_GELU, false, false), 2)           \
    X(mode_of(true, AVDF_ACT_RELU, false, false), 6) X(mode_of(true, AVDF_ACT_RELU, false, false), 2)             \
    X(mode_of(true, AVDF_ACT_RELU, false, false), 1) X(mode_of(true, AVDF_ACT_RELU, false, true), 1)
    // weight-stationary instantiations: q/k/v (16-bit out), attention projection (residual, fp32 out), FPN laterals (fp32 out)
#define AVDF_TC_WS_VARIANTS(X)                                                                                \
    X(mode_of(false, AVDF_ACT_NONE, false, false), 6) X(mode_of(false, AVDF_ACT_NONE, true, false), 1)            \
    X(mode_of(false, AVDF_ACT_NONE, false, false), 1)
#define AVDF_SET_SMEM(M, O) AVDF_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<M, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_of(MAX_BN)));
    // wide tiles with eight epilogue warps: LayerNorm + ReLU (embedding, head towers), K >= 1024 (DownBlock convs, MLP down-projection)
#define AVDF_TC_W8_VARIANTS(X)                                                                                \
    X(mode_of(true, AVDF_ACT_RELU, false, false), 6) X(mode_of(true, AVDF_ACT_RELU, false, false), 2)             \
    X(mode_of(true, AVDF_ACT_RELU, false, false), 1) X(mode_of(true, AVDF_ACT_RELU, false, true), 1)              \
    X(mode_of(false, AVDF_ACT_NONE, false, false), 1) X(mode_of(false, AVDF_ACT_NONE, true, false), 1)            \
    X(mode_of(false, AVDF_ACT_NONE, true, false), 7) X(mode_of(false, AVDF_ACT_NONE, true, false), 3)           \
    X(MODE_POSTLN, 7) X(MODE_POSTLN, 3) X(MODE_HEADDOT, 0)
#define AVDF_SET_SMEM_WS(M, O) AVDF_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<M, O, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES));
#define AVDF_SET_SMEM_W8(M, O) AVDF_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<M, O, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ((M) >= 0 && ((M) & 64)) ? W8D_SMEM_BYTES : W8_SMEM_BYTES)); \
    AVDF_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<M, O, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ((M) >= 0 && ((M) & 64)) ? W8D_SMEM_BYTES : W8_SMEM_BYTES));
    AVDF_SET_SMEM(-1, -1)
    AVDF_TC_VARIANTS(AVDF_SET_SMEM)
    AVDF_SET_SMEM_WS(-1, -1)
    AVDF_TC_WS_VARIANTS(AVDF_SET_SMEM_WS)
    AVDF_SET_SMEM_W8(-1, -1)
    AVDF_TC_W8_VARIANTS(AVDF_SET_SMEM_W8)
#undef AVDF_SET_SMEM
#undef AVDF_SET_SMEM_WS
#undef AVDF_SET_SMEM_W8
    attrs_once.mark(cur_dev);
  }
  AVDF_CHECK_ARG((long long)a->batch * a->o_rows_per_video * a->n_out < (1ll << 31), "output larger than 2^31 elements");
  const int ctas_per_sm = bn <= 128 ? 2 : 1;      // narrow tiles: two co-resident CTAs per SM
  int grid = ws ? ws_groups * ws_per : (p.total_tiles < sms * ctas_per_sm ? p.total_tiles : sms * ctas_per_sm);
  // weight multicast: the launches whose L2 -> SM operand stream is dominated by the weight tile every CTA re-reads (the
  // K >= 512 convolutions: embedding, video branch, head towers). CTAs 2c, 2c + 1 form a cluster and walk adjacent row tiles
  // of the SAME n-tile in lock step (m is the fastest tile index: an even number of m-tiles and an even grid keep a pair
  // inside one n-tile and give both CTAs the same number of tiles).
  const bool pair = !ws && bn == MAX_BN && g_w8 && g_pair_min_k > 0 && a->taps * a->c_in >= g_pair_min_k && p.n_tiles_m % 2 == 0 && grid >= 2;
  const bool mc = pair || (!ws && g_mc_min_k > 0 && a->taps * a->c_in >= g_mc_min_k && p.n_tiles_m % 2 == 0 && grid >= 2 && bn % 16 == 0);
  if (mc) {
    grid &= ~1;
    p.mc = pair ? 0 : 1;
    p.pair = pair ? 1 : 0;
    if (pair) p.idesc = (p.idesc & ~(0x1fu << 24)) | ((unsigned)((2 * BM) >> 4) << 24);     // M = 256 over the two CTAs
    cuuint64_t dims[2] = {(cuuint64_t)a->taps * a->c_in, (cuuint64_t)(a->n_w_rows > 0 ? a->n_w_rows : a->n_out)};
    cuuint64_t strides[1] = {(cuuint64_t)a->taps * a->c_in * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(bn / 2)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.w_half_map, tm_dtype, 2, const_cast<void*>(a->w), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("avdf_conv_gemm: cuTensorMapEncodeTiled(W half) failed with %d", (int)r); return AVDF_ERR_CUDA; }
  }
  int smem_bytes_ = 0;
  auto go = [&](auto kernel, int threads) {
    return mc ? launch_pdl_cluster(kernel, grid, threads, smem_bytes_, st, 2, p) : launch_pdl(kernel, grid, threads, smem_bytes_, st, p);
  };
  const bool w8 = !ws && bn == MAX_BN && g_w8;
  const int smem_bytes = ws ? WS_SMEM_BYTES : (w8 ? (a->dot_out ? W8D_SMEM_BYTES : W8_SMEM_BYTES) : smem_bytes_of(bn));
  smem_bytes_ = smem_bytes;
  const int mode = mode_of(a->ln_w != nullptr, a->act, a->residual != nullptr, a->pe != nullptr) | (a->ln_after_residual ? 32 : 0) | (a->dot_out ? 64 : 0);
  if (a->dot_out) {
    AVDF_CHECK_ARG(w8 && a->n_out == MAX_BN && a->dot_w && a->dot_n >= 1 && a->dot_n <= MAX_DOTS, "dot_out needs n_out = 256, dot_w and 1 <= dot_n <= 6");
  }
  if (a->ln_after_residual) {
    AVDF_CHECK_ARG(w8 && a->n_out == MAX_BN, "ln_after_residual needs n_out = 256 and the eight-warp wide configuration");
  }
  const int outk = (a->out_f32 ? 1 : 0) | (a->out_h ? 2 : 0) | ((a->out_h && a->out_h_dtype == AVDF_DTYPE_F16) ? 4 : 0);
  bool launched = false;
  cudaError_t lerr = cudaSuccess;
#define AVDF_LAUNCH(M, O) if (!launched && !ws && !w8 && mode == (M) && outk == (O)) { lerr = go(conv_gemm_tc_kernel<M, O>, THREADS); launched = true; }
#define AVDF_LAUNCH_WS(M, O) if (!launched && ws && mode == (M) && outk == (O)) { lerr = go(conv_gemm_tc_kernel<M, O, 1>, WS_THREADS); launched = true; }
#define AVDF_LAUNCH_W8(M, O) if (!launched && w8 && mode == (M) && outk == (O)) { lerr = pair ? go(conv_gemm_tc_kernel<M, O, 3>, WS_THREADS) : go(conv_gemm_tc_kernel<M, O, 2>, WS_THREADS); launched = true; }
  AVDF_TC_VARIANTS(AVDF_LAUNCH)
  AVDF_TC_WS_VARIANTS(AVDF_LAUNCH_WS)
  AVDF_TC_W8_VARIANTS(AVDF_LAUNCH_W8)
#undef AVDF_LAUNCH
#undef AVDF_LAUNCH_WS
#undef AVDF_LAUNCH_W8
  if (!launched && a->dot_out) { set_error("avdf_conv_gemm: no dot_out instantiation for this epilogue / output combination"); return AVDF_ERR_UNSUPPORTED; }
  if (!launched && a->ln_after_residual) { set_error("avdf_conv_gemm: no ln_after_residual instantiation for this output combination"); return AVDF_ERR_UNSUPPORTED; }
  if (!launched) {
    if (ws) lerr = go(conv_gemm_tc_kernel<-1, -1, 1>, WS_THREADS);
    else if (w8) lerr = pair ? go(conv_gemm_tc_kernel<-1, -1, 3>, WS_THREADS) : go(conv_gemm_tc_kernel<-1, -1, 2>, WS_THREADS);
    else lerr = go(conv_gemm_tc_kernel<-1, -1>, THREADS);
  }
  if (lerr != cudaSuccess) {
    set_error("conv_gemm_tc_kernel: launch failed: %s (grid %d, smem %d, bn %d, ws %d, w8 %d, multicast %d, pair %d)", cudaGetErrorString(lerr), grid,
              smem_bytes, bn, (int)ws, (int)w8, p.mc, p.pair);
    return AVDF_ERR_CUDA;
  }
  return check_launch("conv_gemm_tc_kernel");
}

}  // namespace avdf

// Debug hook (not part of the C-ABI contract in include/avdf.h's operator list): device buffer of 8 uint64 per CTA
// that receives globaltimer stamps of the kernel phases; NULL switches it off.
extern "C" __attribute__((visibility("default"))) int avdf_debug_gemm_timeline(unsigned long long* dev_buf) {
  avdf::g_dbg = dev_buf;
  return 0;
}
// Debug / test hook: wide tiles with eight (1, default) or four (0) epilogue warps. Returns the previous setting.
extern "C" __attribute__((visibility("default"))) int avdf_debug_gemm_w8(int on) {
  const int prev = avdf::g_w8 ? 1 : 0;
  avdf::g_w8 = on != 0;
  return prev;
}
// Debug / test hook: weight multicast between CTA pairs for launches with taps * c_in >= min_k (0: never). Returns the
// previous setting.
extern "C" __attribute__((visibility("default"))) int avdf_debug_gemm_mc(int min_k) {
  const int prev = avdf::g_mc_min_k;
  avdf::g_mc_min_k = min_k;
  return prev;
}
// Debug / test hook: CTA pairs (cta_group::2) for wide launches with taps * c_in >= min_k (0: never). Returns the previous setting.
extern "C" __attribute__((visibility("default"))) int avdf_debug_gemm_pair(int min_k) {
  const int prev = avdf::g_pair_min_k;
  avdf::g_pair_min_k = min_k;
  return prev;
}
// Debug / test hook: weight-stationary configuration -1 auto (default), 0 never, 1 wherever it is legal. Returns the
// previous setting.
extern "C" __attribute__((visibility("default"))) int avdf_debug_gemm_ws(int mode) {
  const int prev = avdf::g_ws_mode;
  avdf::g_ws_mode = mode;
  return prev;
}
