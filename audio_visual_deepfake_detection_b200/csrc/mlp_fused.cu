// avdf_mlp_fused: the transformer block's MLP (blocks.py:1236-1243, 1315-1316) as ONE tcgen05 kernel for sm_100a:
//     out = residual * mask + gamma * ((GELU(x W1^T + b1) W2^T + b2) * mask)         x [rows, 256] 16-bit, out fp32
// The two GEMMs of the unfused path (256 -> 1024 with a GELU epilogue, 1024 -> 256 with the residual epilogue) move the
// [rows, 1024] hidden activations through L2/HBM twice and re-read their operands per 128x128 tile; here a CTA owns
// 128 rows, keeps x resident in shared memory and walks the hidden dimension in chunks of 128:
//     G1(j): acc1[j & 1] (TMEM, 128 cols)  = x . W1[j]^T                        K = 256
//     E1(j): TMEM -> +b1 -> GELU -> 16-bit -> shared memory, in the K-major 128B-swizzled layout tcgen05.mma reads
//     G2(j): acc2 (TMEM, 256 cols)        += h_j . W2[:, j]^T                   K = 128
// so the hidden activations never leave the SM.
//
// The kernel runs as clusters of TWO CTAs on the two SMs of a TPC and issues cta_group::2 MMAs (M = 256: 128 rows per
// CTA). Each CTA keeps its own x / hidden tiles and accumulators but loads only HALF of every weight tile (64 of the 128
// W1 rows of a chunk, 128 of the 256 W2 rows): per 128 rows an SM ingests 0.5 MB of weights instead of 1 MB and the
// tensor core reads 6-8 KB instead of 8 KB of shared memory per instruction. The single-CTA version of this kernel was
// bound by exactly that: its globaltimer stamps showed the MMA warp issuing back to back at ~70 ns per 128x128x16
// instruction (35 ns when nothing else uses shared memory) - operand reads (256 KB per chunk) + TMA writes (128 KB) + the
// epilogue's stores were ~420 KB per chunk against 128 B/clk of shared-memory bandwidth (profiles/README.md, round 2).
// Barriers the MMA warp WAITS on live in the leader CTA (cluster rank 0): both CTAs' TMA loads credit their bytes there
// (cp.async.bulk.tensor .cta_group::2) and both CTAs' epilogue warps arrive there (mapa + mbarrier.arrive
// .shared::cluster); barriers the MMA warp SIGNALS are reached in both CTAs by multicast commits.
//
// 384 threads per CTA: warp 0 TMA producer (x tile + a 5-slot ring of 16 KB weight half-tiles), warp 1 MMA issuer (leader
// CTA only; G1(j+1) is issued before G2(j): the tensor pipe works while the epilogue warps run GELU on chunk j), warp 2
// TMEM allocator, warps 4-11 epilogue (two warpgroups: TMEM lane quarter = warp % 4, column half = warpgroup). Final
// epilogue: acc2 -> +b2, mask, gamma, + residual, thread = row in the TMEM-native layout: the residual block arrives by TMA
// in the swizzled 4 KB tile the result leaves from (two tiles per warp, the first residual block is fetched while the
// chunk loop still runs), global memory is touched by TMA only (rows beyond the tensor are clipped by the tensor map).
//
// PROJ variant (avdf_mlp_fused_args.att != NULL): the block's attention output projection and LN2 run in the same kernel
// ("block tail": blocks.py:1223, 1309-1316 / 779, 868-872 in one launch):
//     y   = skip * mask + gamma_a * ((att Wo^T + bo) * mask)            acc2 <- att . Wo^T (4 K slices through the weight ring)
//     x   = LN2(y)                                                        never leaves the SM: written into the x tile in the
//                                                                         K-major swizzled operand layout by the epilogue warps
//     out = (y + GELU(x W1^T + b1) W2'^T + b2') * mask                     W2' = diag(gamma_m) W2, b2' = gamma_m b2 (folded by the
//                                                                         caller): y STAYS in the accumulator and the G2
//                                                                         instructions accumulate on top of it
// The att tile arrives by TMA in the (then idle) hidden + staging tiles; the projection epilogue works like the GEMM's
// post-residual LayerNorm epilogue (gemm_tc.cu): pass 1 builds y in the tiles its residual blocks arrived in (they live in
// the x tile's memory, free until pass 2), sends it out, writes it back over the accumulator (tcgen05.st) and sums y, y^2;
// pass 2 re-reads y from TMEM, normalises and writes the 16-bit x tile. Neither y nor the LN2 output leaves the SM: per 128
// rows the kernel reads skip (128 KB) + att (64 KB) and writes out (128 KB) - the fp32 residual stream's minimum - where the
// separate launches moved y out and in twice more. (y is also stored when the caller passes a y pointer: tests.)
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/avdf.h"
#include "tc_ptx.cuh"

namespace avdf {
namespace mlpf {
using namespace tc;

constexpr int C = 256, HID = 1024, BM = 128, HC = 128, NCHUNK = HID / HC;
constexpr int UNIT = 128 * 64 * 2;            // one [128 rows x 64 K] 16-bit operand tile, 128B-swizzled
constexpr int RING = 5;                       // a chunk needs 4 slots (2 of W1, 2 of W2)
#ifndef AVDF_MLP_GELU_WARPS
#define AVDF_MLP_GELU_WARPS 8
#endif
constexpr int EPI_WARPS = 8;                  // warps of the final epilogue (two per TMEM lane quarter: 128 columns each)
constexpr int GELU_WARPS = AVDF_MLP_GELU_WARPS;   // warps of the per-chunk GELU epilogue: 8 (64 hidden columns each) or 16 (32 each)
static_assert(GELU_WARPS == 8 || GELU_WARPS == 16, "GELU warps");
constexpr int GCOLS = HC / (GELU_WARPS / 4);  // hidden columns of a chunk per GELU warp
constexpr int THREADS = 128 + GELU_WARPS * 32;
// shared memory (offsets from a 1024-aligned base)
constexpr int OFF_X = 0;                      // 4 units
constexpr int OFF_H = OFF_X + 4 * UNIT;       // 2 units (the hidden chunk); the final epilogue reuses them as 8 x 4 KB tiles
constexpr int OFF_STG = OFF_H + 2 * UNIT;     // 8 x 4 KB: every epilogue warp's first residual / result tile
constexpr int OFF_RING = OFF_STG + EPI_WARPS * 4096;   // RING units
constexpr int OFF_VEC = OFF_RING + RING * UNIT;        // b1[1024], b2[256], gamma[256]; PROJ: bo, gamma_a, ln2 w, ln2 b [256 each]
constexpr int OFF_PART = OFF_VEC + (HID + 6 * C) * 4;  // PROJ: LayerNorm partial sums, [2 column halves][128 rows] float2
constexpr int OFF_BAR = OFF_PART + 2 * BM * 8;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;       // + alignment slack

struct Params {
  CUtensorMap x_map, w1_map, w2_map;   // boxes: x (64, 128), W1 (64, 64 rows: this CTA's half of a chunk), W2 (64, 128 rows)
  CUtensorMap res_map, out_map;        // fp32 [rows, 256], box (32 columns, 32 rows), swizzle 128B
  CUtensorMap att_map, wo_map, skip_map;   // PROJ: att [rows, 256] (64, 128), Wo [256, 256] (64 K, 128 rows: this CTA's half), skip fp32 like res_map
  const float* bo; const float* gamma_a; const float* ln2_w; const float* ln2_b;
  int y_store;                         // PROJ: also store y through res_map (0: y never leaves the SM)
  void* out_h;
  const float* b1; const float* b2; const float* gamma; const unsigned char* row_mask;
  int rows, tiles;
  int oh_t, oh_pitch, oh_row0;  // 16-bit copy: row r = b * oh_t + t lands at row b * oh_pitch + oh_row0 + t (oh_t = 0: dense)
  unsigned idesc, idesc2;       // M 256 x N 128 (G1), M 256 x N 256 (G2)
  unsigned long long* dbg;      // optional (debug hook): globaltimer stamps of CTA 0's first tile, 3 x 96 slots
};
__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define MLPF_TS(role, slot) do { if (p.dbg && blockIdx.x == 0 && it == 0 && (slot) < 96 && (threadIdx.x & 31) == 0) p.dbg[(role) * 96 + (slot)] = gtime_ns(); } while (0)

// barrier slots
enum { B_XFULL = 0, B_XEMPTY, B_RFULL, B_REMPTY = B_RFULL + RING, B_A1FULL = B_REMPTY + RING, B_A1EMPTY = B_A1FULL + 2,
       B_HFULL = B_A1EMPTY + 2, B_HEMPTY, B_A2FULL, B_A2EMPTY, B_ATTFULL, B_ATTEMPTY, B_RES, B_RES2 = B_RES + 2 * EPI_WARPS,
       B_COUNT = B_RES2 + 2 * EPI_WARPS };
static_assert(B_COUNT * 8 + 8 <= 512, "barrier block");

template <bool F16, bool PROJ>
__global__ void __launch_bounds__(THREADS, 1) mlp_fused_kernel(const __grid_constant__ Params p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* sx = smem + OFF_X;
  unsigned char* sh = smem + OFF_H;
  unsigned char* sstg = smem + OFF_STG;
  unsigned char* sring = smem + OFF_RING;
  float* s_b1 = reinterpret_cast<float*>(smem + OFF_VEC);
  float* s_b2 = s_b1 + HID;
  float* s_gam = s_b2 + C;
  float* s_bo = s_gam + C; float* s_gama = s_bo + C; float* s_l2w = s_gama + C; float* s_l2b = s_l2w + C;   // PROJ
  float2* s_part = reinterpret_cast<float2*>(smem + OFF_PART);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + B_COUNT);
  const uint32_t bar_base = smem_u32(bars);
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rank of this CTA in its pair; the unit of the tile loop is a PAIR of 128-row tiles (tile 2u + rank is this CTA's)
  const uint32_t rank = cluster_ctarank();
  const int unit0 = (int)(blockIdx.x >> 1), unit_step = (int)(gridDim.x >> 1), n_units = (p.tiles + 1) / 2;
  // barrier `i` of the leader CTA as a shared::cluster address (the barriers the MMA warp waits on)
  auto lbar = [&](int i) { return mapa_rank(bar(i), 0u); };
  // hidden chunk handled in step j: the same order in every CTA, so a row's fp32 accumulation order (and with it the
  // result bits) does not depend on which tile / CTA the row falls into (batch invariance). Rotating the order per CTA
  // to spread the weight reads over L2 slices was measured: no gain.
  auto chunk_of = [&](int j) { return j; };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.x_map) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.w1_map) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.w2_map) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.res_map) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.out_map) : "memory");
    if (PROJ) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.att_map) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.wo_map) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.skip_map) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar(B_XFULL), PROJ ? 2 * EPI_WARPS : 1); mbar_init(bar(B_XEMPTY), 1);   // PROJ: the x tile is written by the epilogue warps
    mbar_init(bar(B_ATTFULL), 1); mbar_init(bar(B_ATTEMPTY), EPI_WARPS);
    for (int s = 0; s < RING; ++s) { mbar_init(bar(B_RFULL + s), 1); mbar_init(bar(B_REMPTY + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar(B_A1FULL + s), 1); mbar_init(bar(B_A1EMPTY + s), 2 * GELU_WARPS); }
    mbar_init(bar(B_HFULL), 2 * GELU_WARPS); mbar_init(bar(B_HEMPTY), 1);
    mbar_init(bar(B_A2FULL), 1); mbar_init(bar(B_A2EMPTY), 2 * EPI_WARPS);
    for (int s = 0; s < 4 * EPI_WARPS; ++s) mbar_init(bar(B_RES + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {       // one warp of EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp >= 4) {                               // per-channel vectors, once per CTA
    for (int i = threadIdx.x - 128; i < HID; i += GELU_WARPS * 32) s_b1[i] = p.b1 ? __ldg(p.b1 + i) : 0.f;
    for (int i = threadIdx.x - 128; i < C; i += GELU_WARPS * 32) {
      s_b2[i] = p.b2 ? __ldg(p.b2 + i) : 0.f;
      s_gam[i] = p.gamma ? __ldg(p.gamma + i) : 1.f;
      if (PROJ) {
        s_bo[i] = p.bo ? __ldg(p.bo + i) : 0.f;
        s_gama[i] = p.gamma_a ? __ldg(p.gamma_a + i) : 1.f;
        s_l2w[i] = __ldg(p.ln2_w + i); s_l2b[i] = __ldg(p.ln2_b + i);
      }
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();                            // the peer's barriers are initialised before anything signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tm_acc2 = tmem_base + 256u;
  // programmatic dependent launch: the prologue above (barriers, TMEM, bias vectors) overlapped the previous kernel's tail
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (both CTAs)
    // (whole warp in the loop, one elected lane issues from warp-uniform code: elect_one() in tc_ptx.cuh)
    // A ring slot (16 KB) holds this CTA's half of one MMA group's B operand: two K slices of 64 W1 rows (G1) or one K
    // slice of 128 W2 rows (G2). The bytes of both CTAs are credited to the LEADER's full barrier (its producer posts the
    // expected 32 KB); slots are released in both CTAs by multicast commits.
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0; int it = 0;
    int nu = 0;
    auto begin_slot = [&]() -> uint32_t {
      mbar_wait(bar(B_REMPTY + stage), phase ^ 1);
      MLPF_TS(0, 1 + nu); ++nu;
      if (leader && rank == 0) mbar_arrive_expect_tx(bar(B_RFULL + stage), 2 * UNIT);
      return lbar(B_RFULL + stage);
    };
    auto end_slot = [&]() { __syncwarp(); if (++stage == RING) { stage = 0; phase ^= 1; } };
    auto load_w1 = [&](int chunk, int u) {       // K slices 2u, 2u + 1 of this CTA's 64 rows of W1 chunk `chunk`
      const uint32_t fb = begin_slot();
      if (leader) {
        tma_load_2d_pair(smem_u32(sring + stage * UNIT), &p.w1_map, fb, (2 * u) * 64, chunk * HC + (int)rank * 64);
        tma_load_2d_pair(smem_u32(sring + stage * UNIT + UNIT / 2), &p.w1_map, fb, (2 * u + 1) * 64, chunk * HC + (int)rank * 64);
      }
      end_slot();
    };
    auto load_w2 = [&](int chunk, int sl) {      // K slice sl of this CTA's 128 rows of W2[:, chunk]
      const uint32_t fb = begin_slot();
      if (leader) tma_load_2d_pair(smem_u32(sring + stage * UNIT), &p.w2_map, fb, chunk * HC + sl * 64, (int)rank * 128);
      end_slot();
    };
    for (int unit = unit0; unit < n_units; unit += unit_step, ++it) {
      const int tile = 2 * unit + (int)rank;
      nu = 0;
      if (!PROJ) {
        mbar_wait(bar(B_XEMPTY), (uint32_t)(it & 1) ^ 1);
        MLPF_TS(0, 0);
        if (leader) {
          if (rank == 0) mbar_arrive_expect_tx(bar(B_XFULL), 8 * UNIT);
          const uint32_t xb = lbar(B_XFULL);
          for (int s = 0; s < 4; ++s) tma_load_2d_pair(smem_u32(sx + s * UNIT), &p.x_map, xb, s * 64, tile * BM);
        }
        __syncwarp();
      } else {
        // the att tile lands in the hidden + staging tiles (4 consecutive units), free once the previous tile's final
        // epilogue has drained them; then the four K slices of this CTA's half of Wo go through the ring
        mbar_wait(bar(B_ATTEMPTY), (uint32_t)(it & 1) ^ 1);
        MLPF_TS(0, 0);
        if (leader) {
          if (rank == 0) mbar_arrive_expect_tx(bar(B_ATTFULL), 8 * UNIT);
          const uint32_t ab = lbar(B_ATTFULL);
          for (int s = 0; s < 4; ++s) tma_load_2d_pair(smem_u32(sh + s * UNIT), &p.att_map, ab, s * 64, tile * BM);
        }
        __syncwarp();
        for (int sl = 0; sl < 4; ++sl) {
          const uint32_t fb = begin_slot();
          if (leader) tma_load_2d_pair(smem_u32(sring + stage * UNIT), &p.wo_map, fb, sl * 64, (int)rank * 128);
          end_slot();
        }
      }
      for (int u = 0; u < 2; ++u) load_w1(chunk_of(0), u);
      for (int j = 0; j < NCHUNK; ++j) {
        if (j + 1 < NCHUNK)
          for (int u = 0; u < 2; ++u) load_w1(chunk_of(j + 1), u);
        for (int sl = 0; sl < 2; ++sl) load_w2(chunk_of(j), sl);
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ---------------------------------------------------------------- MMA issuer (leader CTA, for both CTAs)
    // G1(j): 2 ring slots x 2 K slices x 4 instructions of 256 x 128 x 16; G2(j): 2 ring slots x 4 instructions of
    // 256 x 256 x 16 (each CTA holds half of the B rows at the same shared-memory offsets)
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0; int it = 0;
    uint32_t n_a1[2] = {0, 0}, n_h = 0, n_a2e = 0;
    auto g1 = [&](int j) {
      const int b = j & 1;
      mbar_wait(bar(B_A1EMPTY + b), (n_a1[b] & 1) ^ 1);
      tcgen05_fence_after();
      for (int u = 0; u < 2; ++u) {
        mbar_wait(bar(B_RFULL + stage), phase);
        tcgen05_fence_after();
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t da = make_sw128_desc(smem_u32(sx + (2 * u + ks) * UNIT));
            const uint64_t db = make_sw128_desc(smem_u32(sring + stage * UNIT + ks * (UNIT / 2)));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_pair(tmem_base + (uint32_t)(b * HC), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (u > 0 || ks > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(bar(B_REMPTY + stage));
          if (u == 1) umma_commit_pair(bar(B_A1FULL + b));
        }
        __syncwarp();
        if (++stage == RING) { stage = 0; phase ^= 1; }
      }
      ++n_a1[b];
    };
    for (int unit = unit0; unit < n_units; unit += unit_step, ++it) {
      if (PROJ) {
        // acc2 <- att . Wo^T: 4 ring slots x 4 instructions of 256 x 256 x 16
        mbar_wait(bar(B_A2EMPTY), (n_a2e & 1) ^ 1); ++n_a2e;       // the previous tile's final epilogue has drained acc2
        mbar_wait(bar(B_ATTFULL), (uint32_t)(it & 1));
        tcgen05_fence_after();
        for (int sl = 0; sl < 4; ++sl) {
          mbar_wait(bar(B_RFULL + stage), phase);
          tcgen05_fence_after();
          if (leader) {
            const uint64_t da = make_sw128_desc(smem_u32(sh + sl * UNIT));
            const uint64_t db = make_sw128_desc(smem_u32(sring + stage * UNIT));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_pair(tm_acc2, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc2, (sl > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair(bar(B_REMPTY + stage));
            if (sl == 3) umma_commit_pair(bar(B_A2FULL));
          }
          __syncwarp();
          if (++stage == RING) { stage = 0; phase ^= 1; }
        }
      }
      mbar_wait(bar(B_XFULL), (uint32_t)(it & 1));
      tcgen05_fence_after();
      MLPF_TS(1, 0);
      g1(0);
      MLPF_TS(1, 1);
      for (int j = 0; j < NCHUNK; ++j) {
        if (j + 1 < NCHUNK) g1(j + 1);
        MLPF_TS(1, 2 + 4 * j);                  // G1(j+1) issued
        if (j == NCHUNK - 2 && leader) umma_commit_pair(bar(B_XEMPTY));   // every G1 of this tile is issued: x may be replaced
        if (j == 0) { mbar_wait(bar(B_A2EMPTY), (n_a2e & 1) ^ 1); ++n_a2e; tcgen05_fence_after(); }   // (PROJ: LN2 pass 2 has read y)
        mbar_wait(bar(B_HFULL), n_h & 1);
        tcgen05_fence_after();
        MLPF_TS(1, 3 + 4 * j);                  // hidden chunk j ready
        for (int sl = 0; sl < 2; ++sl) {
          mbar_wait(bar(B_RFULL + stage), phase);
          tcgen05_fence_after();
          if (leader) {
            const uint64_t da = make_sw128_desc(smem_u32(sh + sl * UNIT));
            const uint64_t db = make_sw128_desc(smem_u32(sring + stage * UNIT));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_pair(tm_acc2, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc2, (PROJ || j > 0 || sl > 0 || k > 0) ? 1u : 0u);   // PROJ: on top of y
            umma_commit_pair(bar(B_REMPTY + stage));
          }
          __syncwarp();
          if (++stage == RING) { stage = 0; phase ^= 1; }
        }
        if (leader) {
          umma_commit_pair(bar(B_HEMPTY));
          if (j == NCHUNK - 1) umma_commit_pair(bar(B_A2FULL));
        }
        __syncwarp();
        ++n_h;
        MLPF_TS(1, 4 + 4 * j);                  // G2(j) issued
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue warps (both CTAs)
    const int wi = warp - 4;                    // 0..GELU_WARPS-1
    const int q = warp & 3;                     // TMEM lane quarter this warp may read
    const int cg = wi >> 2;                     // GELU: this warp's group of GCOLS hidden columns of a chunk
    const int g = cg & 1;                       // final epilogue (warps 0..7): column half
    const bool fin = wi < EPI_WARPS;            // takes part in the final epilogue
    const int r = q * 32 + lane;                // tile row owned by this thread
    const int sw7 = lane & 7;                   // (r & 7) == (lane & 7)
    // final epilogue: this warp's two 32 x 32 fp32 tiles (128B-swizzled rows). tA is the warp's own all the time; tB is a
    // slice of the hidden tile, free from the moment acc2 is complete until the next tile's first hidden chunk
    unsigned char* tA = sstg + (wi & 7) * 4096;
    unsigned char* tB = sh + (wi & 7) * 4096;
    const uint32_t res_bar[2] = {bar(B_RES + 2 * (wi & 7)), bar(B_RES + 2 * (wi & 7) + 1)};
    uint32_t res_phase[2] = {0, 0};
    const uint32_t a1empty_bar[2] = {lbar(B_A1EMPTY), lbar(B_A1EMPTY + 1)};
    const uint32_t hfull_bar = lbar(B_HFULL), a2empty_bar = lbar(B_A2EMPTY);
    uint32_t n_a1[2] = {0, 0}, n_h = 0, n_a2f = 0;
    const uint32_t xfull_bar = lbar(B_XFULL);
    int it = 0;
    for (int unit = unit0; unit < n_units; unit += unit_step, ++it) {
      const int tile = 2 * unit + (int)rank;
      const int row = tile * BM + r;
      float mk = 1.f;
      if (p.row_mask) mk = (row < p.rows && __ldg(p.row_mask + row)) ? 1.f : 0.f;
      const int wrow = tile * BM + q * 32;      // first row of this warp's 32-row boxes
      const int col0 = g * 128;                 // first of this warp's 128 output columns
      auto fetch_residual = [&](int ch) {       // lane 0: residual block of chunk ch -> tile ch & 1
        mbar_arrive_expect_tx(res_bar[ch & 1], 4096);
        tma_load_2d(smem_u32((ch & 1) ? tB : tA), &p.res_map, res_bar[ch & 1], col0 + ch * 32, wrow);
      };
      if (!PROJ && fin && lane == 0) fetch_residual(0);  // tA was drained at the end of the previous tile
      if (PROJ) {
        // ---- attention projection epilogue + LN2 (warps = final-epilogue warps: all 8)
        // the skip blocks of the four chunks: two tiles inside the (still unused) x tile, requested right away, and - once the
        // projection has consumed the att tile - this warp's staging tile and its slice of the hidden tile
        unsigned char* tX[4] = {sx + wi * 8192, sx + wi * 8192 + 4096, tA, tB};
        const uint32_t sk_bar[4] = {res_bar[0], res_bar[1], bar(B_RES2 + 2 * wi), bar(B_RES2 + 2 * wi + 1)};
        auto fetch_skip = [&](int ch) {
          mbar_arrive_expect_tx(sk_bar[ch], 4096);
          tma_load_2d(smem_u32(tX[ch]), &p.skip_map, sk_bar[ch], col0 + ch * 32, wrow);
        };
        mbar_wait(bar(B_XEMPTY), (uint32_t)(it & 1) ^ 1);       // the previous tile's G1s have read the x tile
        if (lane == 0) { fetch_skip(0); fetch_skip(1); }
        if (wi == 0 && lane == 0) MLPF_TS(2, 42);              // skip tiles requested
        mbar_wait(bar(B_A2FULL), n_a2f & 1); ++n_a2f;
        tcgen05_fence_after();
        if (wi == 0 && lane == 0) MLPF_TS(2, 43);              // projection accumulator complete
        if (lane == 0) { fetch_skip(2); fetch_skip(3); }        // the att tile is consumed
        const uint32_t taddr2 = tm_acc2 + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
        float s = 0.f, ss = 0.f;
        uint32_t vr[32];
        tmem_ld32_issue(taddr2, vr);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          tmem_ld_wait();
          float x[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(vr[i]);
          if (ch + 1 < 4) tmem_ld32_issue(taddr2 + (ch + 1) * 32, vr);
          unsigned char* tb = tX[ch];
          if (ch < 2) { mbar_wait(sk_bar[ch], res_phase[ch]); res_phase[ch] ^= 1; }
          else mbar_wait(sk_bar[ch], (uint32_t)(it & 1));
          const float4* b4 = reinterpret_cast<const float4*>(s_bo + col0 + ch * 32);
          const float4* g4 = reinterpret_cast<const float4*>(s_gama + col0 + ch * 32);
          uint32_t yb[32];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {        // y = gamma_a * ((acc + bo) * mask) + skip * mask, in place (the GEMM epilogue's formula)
            float4* slot = reinterpret_cast<float4*>(tb + lane * 128 + ((jj ^ sw7) << 4));
            const float4 rv = *slot;
            const float4 bb = b4[jj], gg = g4[jj];
            const float4 y = make_float4(fmaf(gg.x, (x[4 * jj] + bb.x) * mk, rv.x * mk), fmaf(gg.y, (x[4 * jj + 1] + bb.y) * mk, rv.y * mk),
                                         fmaf(gg.z, (x[4 * jj + 2] + bb.z) * mk, rv.z * mk), fmaf(gg.w, (x[4 * jj + 3] + bb.w) * mk, rv.w * mk));
            *slot = y;
            s += (y.x + y.y) + (y.z + y.w);
            ss = fmaf(y.x, y.x, ss); ss = fmaf(y.y, y.y, ss); ss = fmaf(y.z, y.z, ss); ss = fmaf(y.w, y.w, ss);
            yb[4 * jj] = __float_as_uint(y.x); yb[4 * jj + 1] = __float_as_uint(y.y); yb[4 * jj + 2] = __float_as_uint(y.z); yb[4 * jj + 3] = __float_as_uint(y.w);
          }
          tmem_st32(taddr2 + ch * 32, yb);        // y replaces the projection in the accumulator: the LayerNorm pass and G2 build on it
          if (p.y_store) {                        // (optional copy of y for the caller)
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&p.res_map, smem_u32(tb), col0 + ch * 32, wrow);
              tma_store_commit();
            }
          } else {
            __syncwarp();                         // every lane has read the tile
          }
        }
        if (wi == 0 && lane == 0) MLPF_TS(2, 44);              // pass 1 done (y stored, written back to TMEM)
        s_part[g * BM + r] = make_float2(s, ss);
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");   // the two warps of this row quarter
        const float2 pa = s_part[r], pb = s_part[BM + r];
        const float mean = (pa.x + pb.x) * (1.f / C);
        const float rstd = rsqrtf(fmaxf((pa.y + pb.y) * (1.f / C) - mean * mean, 0.f) + 1e-5f);
        tmem_st_wait();
        if (p.y_store && lane == 0) tma_store_wait_read();      // the y stores have read this warp's tiles
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");   // every warp's tiles are free: the memory becomes the x tile
        const f32x2 nmean2 = pk2(-mean), rstd2 = pk2(rstd);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {           // pass 2: x = LN2(y) in the K-major 128B-swizzled operand layout
          uint32_t vy[32];
          tmem_ld32_issue(taddr2 + ch * 32, vy);
          tmem_ld_wait();
          if (ch == 3) {                           // y is in registers: acc2 may take G2
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(a2empty_bar);
          }
          const int c0 = col0 + ch * 32;
          unsigned char* xrow = sx + (c0 >> 6) * UNIT + r * 128;
          const int slot0 = (c0 & 63) >> 3;
          const float4* w4 = reinterpret_cast<const float4*>(s_l2w + c0);
          const float4* l4 = reinterpret_cast<const float4*>(s_l2b + c0);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            f32x2 y[4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float4 ww = w4[2 * jj + u], ll = l4[2 * jj + u];
              y[2 * u] = fma2(mul2(add2(pk2(__uint_as_float(vy[8 * jj + 4 * u]), __uint_as_float(vy[8 * jj + 4 * u + 1])), nmean2), rstd2), pk2(ww.x, ww.y), pk2(ll.x, ll.y));
              y[2 * u + 1] = fma2(mul2(add2(pk2(__uint_as_float(vy[8 * jj + 4 * u + 2]), __uint_as_float(vy[8 * jj + 4 * u + 3])), nmean2), rstd2), pk2(ww.z, ww.w), pk2(ll.z, ll.w));
            }
            uint4 uo;
            float f0, f1;
            if (F16) {
              upk2(y[0], f0, f1); uo.x = pack_f16x2(f0, f1); upk2(y[1], f0, f1); uo.y = pack_f16x2(f0, f1);
              upk2(y[2], f0, f1); uo.z = pack_f16x2(f0, f1); upk2(y[3], f0, f1); uo.w = pack_f16x2(f0, f1);
            } else {
              upk2(y[0], f0, f1); uo.x = pack_bf16x2(f0, f1); upk2(y[1], f0, f1); uo.y = pack_bf16x2(f0, f1);
              upk2(y[2], f0, f1); uo.z = pack_bf16x2(f0, f1); upk2(y[3], f0, f1); uo.w = pack_bf16x2(f0, f1);
            }
            *reinterpret_cast<uint4*>(xrow + (((slot0 + jj) ^ sw7) << 4)) = uo;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (wi == 0 && lane == 0) MLPF_TS(2, 45);              // pass 2 done: the x tile is written
        if (lane == 0) mbar_arrive_cluster(xfull_bar);
      }
      for (int j = 0; j < NCHUNK; ++j) {
        const int b = j & 1;
        mbar_wait(bar(B_A1FULL + b), n_a1[b] & 1);
        ++n_a1[b];
        tcgen05_fence_after();
        if (wi == 0 && lane == 0) MLPF_TS(2, 4 * j);          // acc1 chunk j complete
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * HC + cg * GCOLS);
        uint32_t va[GCOLS];
        tmem_ld32_issue(taddr, *reinterpret_cast<uint32_t(*)[32]>(va));
        if (GCOLS == 64) tmem_ld32_issue(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(va + GCOLS - 32));
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(a1empty_bar[b]);
        // this warp's GCOLS hidden columns inside the chunk's K-major tile: K slice (64 columns) and first 16-byte slot
        unsigned char* hrow = sh + ((cg * GCOLS) >> 6) * UNIT + r * 128;
        const int slot0 = ((cg * GCOLS) & 63) >> 3;
        const float* bias = s_b1 + chunk_of(j) * HC + cg * GCOLS;
        uint4 hv[GCOLS / 8];                      // the row's activations, 16-bit: computed before the buffer is free
        const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
        for (int jj = 0; jj < GCOLS / 8; ++jj) { // 8 columns -> one 16-byte chunk of the 128 B row
          f32x2 y[4];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 bb = b4[2 * jj + u];
            y[2 * u] = gelu_poly2(add2(pk2(__uint_as_float(va[8 * jj + 4 * u]), __uint_as_float(va[8 * jj + 4 * u + 1])), pk2(bb.x, bb.y)));
            y[2 * u + 1] = gelu_poly2(add2(pk2(__uint_as_float(va[8 * jj + 4 * u + 2]), __uint_as_float(va[8 * jj + 4 * u + 3])), pk2(bb.z, bb.w)));
          }
          uint4 uo;
          float f0, f1;
          if (F16) {
            upk2(y[0], f0, f1); uo.x = pack_f16x2(f0, f1); upk2(y[1], f0, f1); uo.y = pack_f16x2(f0, f1);
            upk2(y[2], f0, f1); uo.z = pack_f16x2(f0, f1); upk2(y[3], f0, f1); uo.w = pack_f16x2(f0, f1);
          } else {
            upk2(y[0], f0, f1); uo.x = pack_bf16x2(f0, f1); upk2(y[1], f0, f1); uo.y = pack_bf16x2(f0, f1);
            upk2(y[2], f0, f1); uo.z = pack_bf16x2(f0, f1); upk2(y[3], f0, f1); uo.w = pack_bf16x2(f0, f1);
          }
          hv[jj] = uo;
        }
        if (wi == 0 && lane == 0) MLPF_TS(2, 4 * j + 1);      // GELU math done
        mbar_wait(bar(B_HEMPTY), (n_h & 1) ^ 1);           // G2 of the previous chunk has read the hidden tile
        ++n_h;
        if (wi == 0 && lane == 0) MLPF_TS(2, 4 * j + 2);      // hidden buffer free
#pragma unroll
        for (int c = 0; c < GCOLS / 8; ++c) *reinterpret_cast<uint4*>(hrow + (((slot0 + c) ^ sw7) << 4)) = hv[c];
        fence_async_smem();                       // generic-proxy writes -> visible to tcgen05.mma (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(hfull_bar);
        if (wi == 0 && lane == 0) MLPF_TS(2, 4 * j + 3);      // hidden chunk published
      }
      // ---- final epilogue: this warp's 32 rows x 128 columns of acc2 in 4 chunks of 32 columns, thread = row.
      //      chunk c lives in tile c & 1: residual in (TMA), updated in place, result out (TMA); the residual of chunk
      //      c + 1 is in flight while chunk c is computed, the one of chunk c + 2 is fetched once chunk c's store has
      //      read the tile
      if (fin) {
      mbar_wait(bar(B_A2FULL), n_a2f & 1); ++n_a2f;
      tcgen05_fence_after();
      if (wi == 0 && lane == 0) MLPF_TS(2, 40);
      if (!PROJ && lane == 0) fetch_residual(1);  // the hidden tile (tB) is free now
      const uint32_t taddr2 = tm_acc2 + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
      // optional 16-bit copy of the result (the operand of the FPN lateral conv / the next level's block): same dtype as x;
      // rows leave transposed (lane = column) as 64-byte row pieces
      unsigned short* outh_base = p.out_h ? reinterpret_cast<unsigned short*>(p.out_h) + col0 + lane : nullptr;
      int my_dr = wrow + lane;                    // destination row of this lane's row of the block
      if (p.out_h && p.oh_t) my_dr = (my_dr / p.oh_t) * p.oh_pitch + p.oh_row0 + (my_dr % p.oh_t);
      const int nrow = min(32, p.rows - wrow);    // rows of this warp's block inside the tensor (<= 0: none)
      uint32_t vr[32];
      tmem_ld32_issue(taddr2, vr);
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        tmem_ld_wait();
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(vr[i]);
        if (ch + 1 < 4) {
          tmem_ld32_issue(taddr2 + (ch + 1) * 32, vr);
        } else {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(a2empty_bar);
        }
        unsigned char* tb = (ch & 1) ? tB : tA;
        const float4* b4 = reinterpret_cast<const float4*>(s_b2 + col0 + ch * 32);
        if (PROJ) {
          // the accumulator already holds y + the MLP (b2 / W2 pre-scaled by the caller): (acc + b2') * mask
          if (ch >= 2) {                          // this tile's previous store (chunk ch - 2) has read it
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
          }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const float4 bb = b4[jj];
            *reinterpret_cast<float4*>(tb + lane * 128 + ((jj ^ sw7) << 4)) =
                make_float4((x[4 * jj] + bb.x) * mk, (x[4 * jj + 1] + bb.y) * mk, (x[4 * jj + 2] + bb.z) * mk, (x[4 * jj + 3] + bb.w) * mk);
          }
        } else {
        if (ch >= 1 && ch + 1 < 4 && lane == 0) {   // tile (ch + 1) & 1 held chunk ch - 1: its store was issued one chunk ago
          tma_store_wait_read();
          fetch_residual(ch + 1);
        }
        mbar_wait(res_bar[ch & 1], res_phase[ch & 1]);
        res_phase[ch & 1] ^= 1;
        const float4* g4 = reinterpret_cast<const float4*>(s_gam + col0 + ch * 32);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {          // residual * mask + gamma * ((acc + b2) * mask), in place
          float4* slot = reinterpret_cast<float4*>(tb + lane * 128 + ((jj ^ sw7) << 4));
          const float4 rv = *slot;
          const float4 bb = b4[jj], gg = g4[jj];
          *slot = make_float4(fmaf(rv.x, mk, gg.x * ((x[4 * jj] + bb.x) * mk)), fmaf(rv.y, mk, gg.y * ((x[4 * jj + 1] + bb.y) * mk)),
                              fmaf(rv.z, mk, gg.z * ((x[4 * jj + 2] + bb.z) * mk)), fmaf(rv.w, mk, gg.w * ((x[4 * jj + 3] + bb.w) * mk)));
        }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&p.out_map, smem_u32(tb), col0 + ch * 32, wrow);
          tma_store_commit();
        }
        if (outh_base != nullptr) {               // (warp-uniform)
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {       // row rr of the block: 32 lanes = 32 consecutive columns
            const float o = *reinterpret_cast<const float*>(tb + rr * 128 + ((((lane >> 2) ^ (rr & 7)) << 4) | ((lane & 3) << 2)));
            const int dr = __shfl_sync(0xffffffffu, my_dr, rr);
            unsigned short hb;
            if (F16) { const __half hh = __float2half_rn(o); hb = *reinterpret_cast<const unsigned short*>(&hh); }
            else { const __nv_bfloat16 hh = __float2bfloat16_rn(o); hb = *reinterpret_cast<const unsigned short*>(&hh); }
            if (rr < nrow) outh_base[(size_t)dr * C + ch * 32] = hb;
          }
          __syncwarp();
        }
      }
      if (wi == 0 && lane == 0) MLPF_TS(2, 41);
      if (lane == 0) {
        tma_store_wait_read();                    // both tiles drained: tB returns to the hidden tile, tA takes the next residual
        if (PROJ) mbar_arrive(bar(B_ATTEMPTY));   // ... or, with tB, the next att tile
      }
      }
      // every warp's tB lives in the hidden tile, which other warps overwrite in the next tile's first chunk
      asm volatile("bar.sync 1, %0;" ::"n"(GELU_WARPS * 32) : "memory");
    }
    if (fin && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before exit
  }
  tcgen05_fence_before();
  cluster_sync_all();                             // the leader's MMAs read the peer's shared memory and write its TMEM
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace mlpf
}  // namespace avdf

using namespace avdf;

static unsigned long long* g_mlpf_dbg = nullptr;
// Debug hook (not in include/avdf.h): device buffer of 288 uint64 that receives globaltimer stamps of CTA 0's first tile
// (producer / MMA issuer / epilogue warp 4); NULL switches it off.
extern "C" __attribute__((visibility("default"))) int avdf_debug_mlp_timeline(unsigned long long* dev_buf) {
  g_mlpf_dbg = dev_buf;
  return 0;
}

extern "C" int avdf_mlp_fused(const avdf_mlp_fused_args* a, void* stream) {
  using namespace avdf::mlpf;
  AVDF_CHECK_ARG(a != nullptr, "args is null");
  AVDF_CHECK_ARG(a->channels == C && a->hidden == HID, "avdf_mlp_fused supports channels = 256, hidden = 1024");
  AVDF_CHECK_ARG(a->dtype == AVDF_DTYPE_F16 || a->dtype == AVDF_DTYPE_BF16, "dtype must be a 16-bit type");
  AVDF_CHECK_ARG(a->rows >= 0, "rows must be >= 0");
  const bool proj = a->att != nullptr;
  AVDF_CHECK_ARG(a->w1 && a->w2 && a->out, "null pointer");
  if (proj) {
    AVDF_CHECK_ARG(a->w_o && a->ln2_w && a->ln2_b && a->skip, "block tail: w_o, ln2_w, ln2_b and skip are required");
    AVDF_CHECK_ARG(a->gamma == nullptr, "block tail: fold the MLP's scale into w2 / b2 (gamma must be NULL)");
    AVDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(a->att) | reinterpret_cast<uintptr_t>(a->w_o) | reinterpret_cast<uintptr_t>(a->skip) |
                     reinterpret_cast<uintptr_t>(a->y)) & 15) == 0, "pointers must be 16-byte aligned");
    AVDF_CHECK_ARG(GELU_WARPS == EPI_WARPS, "block tail needs the eight-warp build");
  } else {
    AVDF_CHECK_ARG(a->x && a->residual, "null pointer");
  }
  AVDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(proj ? a->att : a->x) | reinterpret_cast<uintptr_t>(a->w1) | reinterpret_cast<uintptr_t>(a->w2) |
                   reinterpret_cast<uintptr_t>(proj ? nullptr : a->residual) | reinterpret_cast<uintptr_t>(a->out)) & 15) == 0, "pointers must be 16-byte aligned");
  if (a->rows == 0) return AVDF_OK;
  tc::EncodeFn encode = tc::get_encode();
  if (!encode) { set_error("avdf_mlp_fused: cuTensorMapEncodeTiled not available from the driver"); return AVDF_ERR_CUDA; }
  const bool f16 = a->dtype == AVDF_DTYPE_F16;
  const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  static thread_local Params p;
  memset(&p, 0, sizeof(p));
  auto enc2 = [&](CUtensorMap* m, CUtensorMapDataType t, int es, const void* base, long long cols, long long rows_, int bc, int br,
                  const char* what) -> int {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows_};
    cuuint64_t strides[1] = {(cuuint64_t)cols * es};
    cuuint32_t box[2] = {(cuuint32_t)bc, (cuuint32_t)br};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(m, t, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("avdf_mlp_fused: cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r); return AVDF_ERR_CUDA; }
    return AVDF_OK;
  };
  int rc;
  if (!proj && (rc = enc2(&p.x_map, dt, 2, a->x, C, a->rows, 64, 128, "x"))) return rc;
  if (proj) {
    if ((rc = enc2(&p.att_map, dt, 2, a->att, C, a->rows, 64, 128, "att"))) return rc;
    if ((rc = enc2(&p.wo_map, dt, 2, a->w_o, C, C, 64, 128, "w_o"))) return rc;
    if ((rc = enc2(&p.skip_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->skip, C, a->rows, 32, 32, "skip"))) return rc;
    p.bo = a->b_o; p.gamma_a = a->gamma_attn; p.ln2_w = a->ln2_w; p.ln2_b = a->ln2_b;
  }
  if ((rc = enc2(&p.w1_map, dt, 2, a->w1, C, HID, 64, 64, "w1"))) return rc;
  if ((rc = enc2(&p.w2_map, dt, 2, a->w2, HID, C, 64, 128, "w2"))) return rc;
  p.y_store = (proj && a->y) ? 1 : 0;
  if ((!proj || a->y) && (rc = enc2(&p.res_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, proj ? a->y : a->residual, C, a->rows, 32, 32, "residual"))) return rc;
  if ((rc = enc2(&p.out_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->out, C, a->rows, 32, 32, "out"))) return rc;
  p.out_h = a->out_h;
  p.oh_t = a->out_h_t; p.oh_pitch = a->out_h_pitch; p.oh_row0 = a->out_h_row0;
  AVDF_CHECK_ARG(a->out_h_t >= 0 && (a->out_h_t == 0 || (a->rows % a->out_h_t == 0 && a->out_h_pitch >= a->out_h_t + a->out_h_row0)),
                 "out_h_t must divide rows and fit the destination pitch");
  AVDF_CHECK_ARG(!a->out_h || (reinterpret_cast<uintptr_t>(a->out_h) & 15) == 0, "out_h must be 16-byte aligned");
  p.b1 = a->b1; p.b2 = a->b2; p.gamma = a->gamma; p.row_mask = a->row_mask;
  p.rows = a->rows; p.tiles = (a->rows + BM - 1) / BM;
  p.dbg = g_mlpf_dbg;
  const unsigned fmt = f16 ? 0u : 1u;
  // instruction descriptors: D = f32, A/B format, N >> 3 at bit 17, M >> 4 at bit 24 (M = 256 over the two CTAs of a pair)
  const unsigned m_field = (unsigned)((2 * BM) >> 4) << 24;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(128 >> 3) << 17) | m_field;
  p.idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(256 >> 3) << 17) | m_field;
  const int sms = device_sm_count();
  AVDF_SMEM_ATTR_ONCE((mlp_fused_kernel<true, false>), SMEM_BYTES);
  AVDF_SMEM_ATTR_ONCE((mlp_fused_kernel<false, false>), SMEM_BYTES);
  AVDF_SMEM_ATTR_ONCE((mlp_fused_kernel<true, true>), SMEM_BYTES);
  AVDF_SMEM_ATTR_ONCE((mlp_fused_kernel<false, true>), SMEM_BYTES);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // clusters of two CTAs (one TPC); a pair walks pairs of 128-row tiles
  const int units = (p.tiles + 1) / 2, max_pairs = sms / 2;
  const int grid = 2 * (units < max_pairs ? units : max_pairs);
  cudaError_t lerr;
  if (proj) lerr = f16 ? launch_pdl_cluster(mlp_fused_kernel<true, true>, grid, THREADS, SMEM_BYTES, st, 2, p)
                       : launch_pdl_cluster(mlp_fused_kernel<false, true>, grid, THREADS, SMEM_BYTES, st, 2, p);
  else lerr = f16 ? launch_pdl_cluster(mlp_fused_kernel<true, false>, grid, THREADS, SMEM_BYTES, st, 2, p)
                  : launch_pdl_cluster(mlp_fused_kernel<false, false>, grid, THREADS, SMEM_BYTES, st, 2, p);
  if (lerr != cudaSuccess) { set_error("mlp_fused_kernel: launch failed: %s", cudaGetErrorString(lerr)); return AVDF_ERR_CUDA; }
  return check_launch("mlp_fused_kernel");
}
