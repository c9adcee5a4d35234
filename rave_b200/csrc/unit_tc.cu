// Fused Residual(DilatedUnit) forward on tcgen05 (sm_100a):
//
//     out = x + Conv1x1( LeakyReLU( Conv3_dil( LeakyReLU(x) ) ) )            rave/blocks.py:31-45 (Residual), 83-112 (DilatedUnit)
//
// in ONE kernel.  The reference runs 2 cuDNN convs + 2 activations + pads + the add (7 kernels, 5 HBM round trips);
// the per-layer tcgen05 path (conv_tc.cu) still needed 2 launches with the intermediate operand a1 = LeakyReLU(conv3)
// going through HBM.  Here a CTA owns a tile of 128 time steps x ALL C channels:
//
//   phase 1   acc[128 x NC] (TMEM) = sum_{tap k<3} sum_kb  A_k[128 x BK] (TMA, rows shifted by k*dil - pad) * W3_k[NC x BK]^T
//   epilogue1 TMEM -> LeakyReLU -> bf16 -> shared memory, written directly in the K-major swizzled layout the tensor
//             core reads (the A operand of phase 2 never leaves the SM); optionally also to HBM (training keeps a1 for the
//             backward: write-only, no read)
//   phase 2   acc[128 x NC] = sum_kb  A2[128 x BK] (shared memory) * W1[NC x BK]^T
//   epilogue2 TMEM + skip -> out (bf16 operand of the next layer and / or the fp32 stream); the skip h = x is
//             recovered from the unit's own input operand a = LeakyReLU(x) (inverse LeakyReLU), as conv_tc.cu does
//
// C = 384 runs both phases in two N chunks of 192 (UMMA N <= 256; the 128 x 384 bf16 A2 tile is 96 KB of shared memory).
// Two 256-column TMEM buffers alternate between consecutive (phase, chunk) jobs, so the epilogue of one job overlaps
// the MMAs of the next -- except at the phase-1 -> phase-2 boundary of a tile, where phase 2 needs the complete A2.
// Algorithmic HBM bytes per unit: 2 B*L*C (operand in) + 2 B*L*C (operand out) [+ 2 B*L*C a1 when training] + 8 C^2
// (weights) -- C = 96, L = 4096, B = 32: 50 MB instead of 126 MB for the two-launch form.
//
// Warp roles (320 threads): 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..9 = epilogue (two warps per TMEM lane
// quadrant, alternate 32-column chunks).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace rave {
namespace tc {

constexpr int U_THREADS = 320;
constexpr int U_EPI_WARPS = 8;

struct UnitParams {
  int B, C, L, pitch;          // channel-last [B][pitch][C] operand tensors; L valid rows
  int dil, pad_l;
  int BL, BB, n_lt, n_bg;      // tile = BB batches x BL rows (BL * BB == 128)
  float slope_in_inv;          // 1 / slope of the LeakyReLU that produced the input operand (skip recovery)
  float slope_mid;             // LeakyReLU between the two convs
  int act_out;                 // activation applied to the written operand (RAVE_ACT_NONE / RAVE_ACT_LEAKY)
  float slope_out;
  const __nv_bfloat16 *xa;     // input operand (also the skip source)
  __nv_bfloat16 *a1_out;       // [B][pitch][C] or null: the intermediate operand, kept for the backward
  float *out_f32;              // [B][pitch][C] or null
  __nv_bfloat16 *out_act;      // [B][pitch][C] or null
};

template <int C, int BK>
struct UnitCfg {
  static constexpr int NC = C <= 256 ? C : C / 2;             // N chunk (UMMA N <= 256)
  static constexpr int NCH = C / NC;
  static constexpr int KB = C / BK;                           // K blocks per tap
  static constexpr int SWZ = BK * 2;
  static constexpr int A_BYTES = 128 * BK * 2;
  static constexpr int B_BYTES = NC * BK * 2;
  static constexpr int B_PAD = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = A_BYTES + B_PAD;
  static constexpr int A2_SLAB = 128 * BK * 2;
  static constexpr int A2_BYTES = KB * A2_SLAB;               // the whole [128 x C] bf16 tile
  static constexpr int MAX_STAGES = (222 * 1024 - A2_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = MAX_STAGES > 6 ? 6 : MAX_STAGES;
  static constexpr int A2_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = A2_OFFSET + A2_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
  static_assert(STAGES >= 2, "not enough shared memory for a 2-stage pipeline");
  static_assert(NC % 32 == 0 && NC <= 256, "chunk must be a multiple of 32 columns");
};

__device__ __forceinline__ float bfl(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bfh(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

template <int C, int BK>
__global__ void __launch_bounds__(U_THREADS, 1)
dilated_unit_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w3,
                       const __grid_constant__ CUtensorMap tmap_w1, const UnitParams p) {
  using L = UnitCfg<C, BK>;
  constexpr int STAGES = L::STAGES, NC = L::NC, NCH = L::NCH, KB = L::KB, SWZ = L::SWZ;
  constexpr uint32_t TMEM_COLS = 512;                   // two accumulator buffers at columns 0 and 256

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *a2 = smem + L::A2_OFFSET;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L::BAR_OFFSET);
  uint64_t *empty_bar = full_bar + STAGES;
  uint64_t *tfull_bar = empty_bar + STAGES;             // [2]
  uint64_t *tempty_bar = tfull_bar + 2;                 // [2]
  uint64_t *a2_ready = tempty_bar + 2;                  // epilogue -> MMA: the tile's A2 operand is complete
  uint64_t *a2_free = a2_ready + 1;                     // MMA -> epilogue: phase 2 has finished reading A2
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(a2_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.n_lt * p.n_bg;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w3);
    tma_prefetch_desc(&tmap_w1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], U_EPI_WARPS);
    }
    mbar_init(a2_ready, U_EPI_WARPS);
    mbar_init(a2_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  griddep_launch_dependents();      // dependents may begin their prologue ...
  griddep_wait();                   // ... and this kernel touches global memory only after its predecessors are done

  if (warp == 0) {
    // =========================== TMA producer ===========================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int lt = tile % p.n_lt, bg = tile / p.n_lt;
      const int l0 = lt * p.BL, b0 = bg * p.BB;
      for (int ch = 0; ch < NCH; ++ch) {                 // phase 1: activation rows + conv3 weights
        for (int k = 0; k < 3; ++k) {
          const int row = l0 + k * p.dil - p.pad_l;
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t *sa = smem + stage * L::STAGE_BYTES;
            if (elect_one()) {
              mbar_arrive_expect_tx(&full_bar[stage], L::A_BYTES + L::B_BYTES);
              tma_load_4d(sa, &tmap_a, &full_bar[stage], kb * BK, 0, row, b0);
              tma_load_2d(sa + L::A_BYTES, &tmap_w3, &full_bar[stage], kb * BK, k * C + ch * NC);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
      for (int ch = 0; ch < NCH; ++ch) {                 // phase 2: conv1 weights only (A2 is resident)
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t *sa = smem + stage * L::STAGE_BYTES;
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[stage], L::B_BYTES);
            tma_load_2d(sa + L::A_BYTES, &tmap_w1, &full_bar[stage], kb * BK, ch * NC);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc_bf16(128, NC);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t a2_base = smem_u32(a2);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    int stage = 0;
    uint32_t phase = 0;
    int job = 0, it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      for (int ch = 0; ch < NCH; ++ch, ++job) {          // ---- phase 1
        const int buf = job & 1;
        mbar_wait(&tempty_bar[buf], ((job >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_u + buf * 256;
        for (int u = 0; u < 3 * KB; ++u) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
          const uint64_t adesc = make_kmajor_desc(sa, SWZ);
          const uint64_t bdesc = make_kmajor_desc(sa + L::A_BYTES, SWZ);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
              umma_f16(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (u > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (u == 3 * KB - 1) umma_commit(&tfull_bar[buf]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      mbar_wait(a2_ready, it & 1);                       // every epilogue warp has written (and fenced) its A2 rows
      tc_fence_after();
      for (int ch = 0; ch < NCH; ++ch, ++job) {          // ---- phase 2
        const int buf = job & 1;
        mbar_wait(&tempty_bar[buf], ((job >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_u + buf * 256;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sb = smem_base + stage * L::STAGE_BYTES + L::A_BYTES;
          const uint64_t adesc = make_kmajor_desc(a2_base + kb * L::A2_SLAB, SWZ);
          const uint64_t bdesc = make_kmajor_desc(sb, SWZ);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
              umma_f16(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (kb == KB - 1) {
              umma_commit(&tfull_bar[buf]);
              if (ch == NCH - 1) umma_commit(a2_free);   // all MMAs that read A2 have retired when this fires
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =========================== epilogue (8 warps) ===========================
    const int quad = warp & 3;                           // TMEM lane quadrant this warp may access
    const int part = (warp - 2) >> 2;                    // 0 / 1: alternate 32-column chunks
    const int row = quad * 32 + lane;
    // byte offset of this row's 16-byte chunk c16 inside an A2 slab: row * SWZ + ((c16 ^ swz(row)) << 4)
    const uint32_t row_xor = (SWZ == 128) ? (uint32_t)(row & 7) : (SWZ == 64) ? (uint32_t)((row >> 1) & 3)
                                                                                : (uint32_t)((row >> 2) & 1);
    int job = 0, it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int lt = tile % p.n_lt, bg = tile / p.n_lt;
      const int b = bg * p.BB + row / p.BL;
      const int l = lt * p.BL + row % p.BL;
      const bool valid = (b < p.B) && (l < p.L);
      const size_t grow = ((size_t)b * p.pitch + l) * C;
      if (it > 0) mbar_wait(a2_free, (it - 1) & 1);      // the previous tile's phase 2 no longer reads A2
      // ---- epilogue 1: acc -> LeakyReLU -> bf16 -> A2 (swizzled K-major rows) [+ HBM copy for the backward]
      for (int ch = 0; ch < NCH; ++ch, ++job) {
        const int buf = job & 1;
        mbar_wait(&tfull_bar[buf], (job >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256;
#pragma unroll 1
        for (int c0 = part * 32; c0 < NC; c0 += 64) {
          float v[32];
          tmem_ld_32x32(taddr + c0, v);
          uint32_t pk[16];
#pragma unroll
          for (int w = 0; w < 16; ++w) {
            const float a0 = fmaxf(v[2 * w], v[2 * w] * p.slope_mid), a1 = fmaxf(v[2 * w + 1], v[2 * w + 1] * p.slope_mid);
            __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
            pk[w] = *reinterpret_cast<uint32_t *>(&h);
          }
          const int col = ch * NC + c0;                  // first channel of this 32-column chunk
#pragma unroll
          for (int q = 0; q < 4; ++q) {                  // four 16-byte pieces (8 channels each)
            const int cc = col + 8 * q;
            const uint32_t off = (uint32_t)(cc / BK) * L::A2_SLAB + (uint32_t)row * SWZ +
                                 ((((uint32_t)(cc % BK) >> 3) ^ row_xor) << 4);
            *reinterpret_cast<uint4 *>(a2 + off) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
          if (p.a1_out && valid) {
            stg256(p.a1_out + grow + col, pk);
            stg256(p.a1_out + grow + col + 16, pk + 8);
          }
        }
        tc_fence_before();
        if (ch == NCH - 1) fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&tempty_bar[buf]);
          if (ch == NCH - 1) mbar_arrive(a2_ready);
        }
      }
      // ---- epilogue 2: acc + skip -> outputs
      for (int ch = 0; ch < NCH; ++ch, ++job) {
        const int buf = job & 1;
        // skip rows of this thread's chunks: issue the loads before waiting for the accumulator
        uint32_t sk[(NC + 63) / 64][16];
#pragma unroll
        for (int i = 0; i < (NC + 63) / 64; ++i) {
          const int c0 = part * 32 + 64 * i;
          if (c0 < NC && valid) {
            ldg256(p.xa + grow + ch * NC + c0, sk[i]);
            ldg256(p.xa + grow + ch * NC + c0 + 16, sk[i] + 8);
          }
        }
        mbar_wait(&tfull_bar[buf], (job >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256;
#pragma unroll
        for (int i = 0; i < (NC + 63) / 64; ++i) {
          const int c0 = part * 32 + 64 * i;
          if (c0 < NC) {
            float v[32];
            tmem_ld_32x32(taddr + c0, v);
            if (valid) {
              const int col = ch * NC + c0;
#pragma unroll
              for (int w = 0; w < 16; ++w) {
                const float s0 = bfl(sk[i][w]), s1 = bfh(sk[i][w]);
                v[2 * w] += fminf(s0, s0 * p.slope_in_inv);          // inverse LeakyReLU of the input operand
                v[2 * w + 1] += fminf(s1, s1 * p.slope_in_inv);
              }
              if (p.out_f32) {
                uint32_t o[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(v[j]);
#pragma unroll
                for (int j = 0; j < 4; ++j) stg256(p.out_f32 + grow + col + 8 * j, o + 8 * j);
              }
              if (p.out_act) {
                uint32_t pk[16];
#pragma unroll
                for (int w = 0; w < 16; ++w) {
                  float a0 = v[2 * w], a1 = v[2 * w + 1];
                  if (p.act_out == RAVE_ACT_LEAKY) {
                    a0 = fmaxf(a0, a0 * p.slope_out);
                    a1 = fmaxf(a1, a1 * p.slope_out);
                  }
                  __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
                  pk[w] = *reinterpret_cast<uint32_t *>(&h);
                }
                stg256(p.out_act + grow + col, pk);
                stg256(p.out_act + grow + col + 16, pk + 8);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =============================================================================================
// Weight-stationary variant for the narrow units (C = 96: the six longest launches of the encoder / generator forward).
// dilated_unit_tc_kernel re-streams, for EVERY 128-row tile, the three tap-shifted copies of the activation rows
// (3 x 24 KB) and the whole weight set (W3 54 KB + W1 18 KB) from L2: 144 KB per tile, 35 B/clk/SM sustained -- the
// L2 -> SM path (~42 B/clk/SM), not the tensor pipe (18 % active) or HBM, bounded it (profiles/r2_ncu_unit96.md).
// Here the weights are loaded ONCE per CTA and stay in shared memory (72 KB), and each tile brings ONE haloed
// activation tile (128 + 2 dil rows): tap k is the same tile read (k dil) rows further down -- K-major operand, 64-byte
// swizzle, descriptor start address moved by whole rows (the swizzle is a function of the absolute shared-memory
// address, which is how TMA wrote it).  29 KB instead of 144 KB per tile.
// =============================================================================================
constexpr int UW_AROWS = 152;                       // 128 + 2 * dil rows, dil <= 12
constexpr int UW_ASLAB = UW_AROWS * 64;             // one 32-channel K block of the haloed tile (64-byte rows)
constexpr int UW_WSLAB = 96 * 64;                   // one (tap, K block) weight slab: 96 rows x 32 channels
constexpr int UW_STAGES = 3;
constexpr int UW_W3_OFF = 0, UW_W1_OFF = 9 * UW_WSLAB, UW_A_OFF = 12 * UW_WSLAB;
constexpr int UW_A2_OFF = UW_A_OFF + UW_STAGES * 3 * UW_ASLAB;
constexpr int UW_OUT_OFF = UW_A2_OFF + 3 * (128 * 64);      // staging tile of the bf16 output (TMA store source)
constexpr int UW_BAR_OFF = UW_OUT_OFF + 3 * (128 * 64);
constexpr int UW_TOTAL = UW_BAR_OFF + 256 + 1024;

__global__ void __launch_bounds__(U_THREADS, 1)
dilated_unit_ws96_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w3,
                         const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_out,
                         const __grid_constant__ CUtensorMap tmap_a1, const UnitParams p) {
  constexpr int C = 96, BK = 32, KB = 3, SWZ = 64, A2_SLAB = 128 * 64;
  constexpr uint32_t TMEM_COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *a2 = smem + UW_A2_OFF;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + UW_BAR_OFF);
  uint64_t *empty_bar = full_bar + UW_STAGES;
  uint64_t *tfull_bar = empty_bar + UW_STAGES;          // [2]
  uint64_t *tempty_bar = tfull_bar + 2;                 // [2]
  uint64_t *a2_ready = tempty_bar + 2;
  uint64_t *a2_free = a2_ready + 1;
  uint64_t *w_bar = a2_free + 1;                        // the resident weights have landed
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.n_lt * p.n_bg;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w3);
    tma_prefetch_desc(&tmap_w1);
    for (int s = 0; s < UW_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], U_EPI_WARPS);     // released by the epilogue: it reads the skip rows from the tile
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], U_EPI_WARPS);
    }
    mbar_init(a2_ready, U_EPI_WARPS);
    mbar_init(a2_free, 1);
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  griddep_launch_dependents();
  griddep_wait();
  const int arows = 128 + 2 * p.dil;                    // rows of the haloed tile actually loaded

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {                                   // weights: once
      mbar_arrive_expect_tx(w_bar, 12 * UW_WSLAB);
      for (int k = 0; k < 3; ++k)
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(smem + UW_W3_OFF + (k * KB + kb) * UW_WSLAB, &tmap_w3, w_bar, kb * BK, k * C);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(smem + UW_W1_OFF + kb * UW_WSLAB, &tmap_w1, w_bar, kb * BK, 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int lt = tile % p.n_lt, b0 = tile / p.n_lt;          // BB == 1
      const int row0 = lt * 128 - p.pad_l;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t *sa = smem + UW_A_OFF + stage * 3 * UW_ASLAB;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(KB * arows * 64));
        for (int kb = 0; kb < KB; ++kb) tma_load_4d(sa + kb * UW_ASLAB, &tmap_a, &full_bar[stage], kb * BK, 0, row0, b0);
      }
      __syncwarp();
      if (++stage == UW_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc_bf16(128, C);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t a2_base = smem_u32(a2);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    mbar_wait(w_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    int job = 0, it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      {                                                   // ---- phase 1: conv3 over the haloed tile
        const int buf = job & 1;
        mbar_wait(&tempty_bar[buf], ((job >> 1) & 1) ^ 1);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t tmem_d = tmem_u + buf * 256;
        const uint32_t sa = smem_base + UW_A_OFF + stage * 3 * UW_ASLAB;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
              const uint64_t adesc = make_kmajor_desc(sa + kb * UW_ASLAB + (uint32_t)(k * p.dil) * 64u, SWZ);
              const uint64_t bdesc = make_kmajor_desc(smem_base + UW_W3_OFF + (k * KB + kb) * UW_WSLAB, SWZ);
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk)
                umma_f16(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k > 0 || kb > 0 || kk > 0) ? 1u : 0u);
            }
          }
          umma_commit(&tfull_bar[buf]);
        }
        __syncwarp();
        if (++stage == UW_STAGES) { stage = 0; phase ^= 1; }
        ++job;
      }
      mbar_wait(a2_ready, it & 1);
      tc_fence_after();
      {                                                   // ---- phase 2: 1x1 conv on the resident A2 / W1
        const int buf = job & 1;
        mbar_wait(&tempty_bar[buf], ((job >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_u + buf * 256;
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t adesc = make_kmajor_desc(a2_base + kb * A2_SLAB, SWZ);
            const uint64_t bdesc = make_kmajor_desc(smem_base + UW_W1_OFF + kb * UW_WSLAB, SWZ);
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
              umma_f16(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&tfull_bar[buf]);
          umma_commit(a2_free);
        }
        __syncwarp();
        ++job;
      }
    }
  } else {
    // =========================== epilogue (8 warps) ===========================
    // Per-thread row accesses to global memory touch 32 different 128-byte lines per instruction (L1 wavefront bound:
    // the skip read and the output write were ~1.6k clocks each per tile).  Here the skip comes from the haloed tile
    // that is already in shared memory (its centre rows), and both bf16 outputs leave through TMA: a1 straight from
    // the A2 tile (same K-major swizzled layout as its tensor map's boxes), the unit's output from a staging tile.
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t row_xor = (uint32_t)((row >> 1) & 3);          // 64-byte swizzle of tile row `row`
    const int srow = row + p.pad_l;                               // the same position inside the haloed tile
    const uint32_t srow_off = (uint32_t)srow * 64u, srow_xor = (uint32_t)((srow >> 1) & 3);
    const bool issuer = threadIdx.x == 64;
    uint8_t *outb = smem + UW_OUT_OFF;
    int job = 0, it = 0;
    int stage = 0;
    uint32_t sphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int lt = tile % p.n_lt, b = tile / p.n_lt;
      const int l0 = lt * 128;
      const int l = l0 + row;
      const bool valid = (b < p.B) && (l < p.L);
      const size_t grow = ((size_t)b * p.pitch + l) * C;
      if (it > 0) mbar_wait(a2_free, (it - 1) & 1);
      if (p.a1_out && it > 0) {                                   // the a1 store of the previous tile has read A2
        if (issuer) {                                             // (bulk groups complete in order: a1, out, a1, out ...)
          if (p.out_act) bulk_wait_read<1>();
          else bulk_wait_read<0>();
        }
        named_bar_sync(2, 256);
      }
      {   // ---- epilogue 1: acc -> LeakyReLU -> bf16 -> A2 (swizzled K-major rows)
        const int buf = job & 1;
        mbar_wait(&tfull_bar[buf], (job >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256;
        const float2 s2 = make_float2(p.slope_mid, p.slope_mid);
#pragma unroll 1
        for (int c0 = part * 32; c0 < C; c0 += 64) {
          float v[32];
          tmem_ld_32x32(taddr + c0, v);
          uint32_t pk[16];
#pragma unroll
          for (int w = 0; w < 16; ++w) {
            const float2 t = make_float2(v[2 * w], v[2 * w + 1]);
            const float2 u = __fmul2_rn(t, s2);
            const __nv_bfloat162 h = __hmax2(__floats2bfloat162_rn(t.x, t.y), __floats2bfloat162_rn(u.x, u.y));
            pk[w] = *reinterpret_cast<const uint32_t *>(&h);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)      // chunk c0 = K block c0 / 32: 16-byte unit q of this row
            sts128(a2 + (uint32_t)(c0 / BK) * A2_SLAB + (uint32_t)row * SWZ + (((uint32_t)q ^ row_xor) << 4), pk + 4 * q);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&tempty_bar[buf]);
          mbar_arrive(a2_ready);
        }
        if (p.a1_out) {                                           // training: keep a1 for the backward (write-only)
          named_bar_sync(2, 256);
          if (issuer) {
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) tma_store_4d(&tmap_a1, a2 + kb * A2_SLAB, kb * BK, 0, l0, b);
            bulk_commit();
          }
        }
        ++job;
      }
      {   // ---- epilogue 2: acc + skip -> outputs
        const int buf = job & 1;
        mbar_wait(&tfull_bar[buf], (job >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256;
        const uint8_t *sa = smem + UW_A_OFF + stage * 3 * UW_ASLAB;      // this tile's haloed activation rows
        mbar_wait(&full_bar[stage], sphase);                     // (long complete: makes the TMA writes visible here)
        if (p.out_act) {                                          // the previous tile's output store has read `outb`
          if (issuer) {
            if (p.a1_out) bulk_wait_read<1>();                    // ... this tile's a1 store may still be running
            else bulk_wait_read<0>();
          }
          named_bar_sync(1, 256);
        }
#pragma unroll 1
        for (int c0 = part * 32; c0 < C; c0 += 64) {
          uint32_t sk[16];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            lds128(sa + (uint32_t)(c0 / BK) * UW_ASLAB + srow_off + (((uint32_t)q ^ srow_xor) << 4), sk + 4 * q);
          float v[32];
          tmem_ld_32x32(taddr + c0, v);
#pragma unroll
          for (int w = 0; w < 16; ++w) {
            const float s0 = bfl(sk[w]), s1 = bfh(sk[w]);
            v[2 * w] += fminf(s0, s0 * p.slope_in_inv);          // inverse LeakyReLU of the input operand
            v[2 * w + 1] += fminf(s1, s1 * p.slope_in_inv);
          }
          if (p.out_f32 && valid) {
            uint32_t o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(v[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) stg256(p.out_f32 + grow + c0 + 8 * j, o + 8 * j);
          }
          if (p.out_act) {
            uint32_t pk[16];
            const float2 so = make_float2(p.slope_out, p.slope_out);
#pragma unroll
            for (int w = 0; w < 16; ++w) {
              const float2 t = make_float2(v[2 * w], v[2 * w + 1]);
              __nv_bfloat162 h = __floats2bfloat162_rn(t.x, t.y);
              if (p.act_out == RAVE_ACT_LEAKY) {
                const float2 u = __fmul2_rn(t, so);
                h = __hmax2(h, __floats2bfloat162_rn(u.x, u.y));
              }
              pk[w] = *reinterpret_cast<const uint32_t *>(&h);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
              sts128(outb + (uint32_t)(c0 / BK) * A2_SLAB + (uint32_t)row * SWZ + (((uint32_t)q ^ row_xor) << 4), pk + 4 * q);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&tempty_bar[buf]);
          mbar_arrive(&empty_bar[stage]);                          // skip rows read: the producer may refill this stage
        }
        if (p.out_act) {
          fence_proxy_async();
          named_bar_sync(1, 256);
          if (issuer) {
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) tma_store_4d(&tmap_out, outb + kb * A2_SLAB, kb * BK, 0, l0, b);
            bulk_commit();
          }
        }
        ++job;
      }
      if (++stage == UW_STAGES) { stage = 0; sphase ^= 1; }
    }
    if (issuer) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFnU)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFnU unit_encode_fn() {
  static EncodeTiledFnU fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFnU)ptr;
  }
  return fn;
}

template <int C, int BK>
static int launch_unit(const CUtensorMap &ta, const CUtensorMap &t3, const CUtensorMap &t1, const UnitParams &p,
                       cudaStream_t stream) {
  using L = UnitCfg<C, BK>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(dilated_unit_tc_kernel<C, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::TOTAL);
    if (e != cudaSuccess) {
      set_error("dilated_unit_tc: cudaFuncSetAttribute(%d bytes): %s", L::TOTAL, cudaGetErrorString(e));
      return 2;
    }
    attr = true;
  }
  const int tiles = p.n_lt * p.n_bg;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = tiles < sms ? tiles : sms;
  launch_pdl(dilated_unit_tc_kernel<C, BK>, dim3(grid), dim3(U_THREADS), L::TOTAL, stream, ta, t3, t1, p);
  RAVE_CHECK_LAUNCH("dilated_unit_tc");
  return 0;
}

}  // namespace tc
}  // namespace rave

extern "C" int rave_dilated_unit_tc_supported(int C, int L) {
  return (C == 96 || C == 192 || C == 384) && L >= 8;
}

extern "C" int rave_dilated_unit_tc_fwd(const void *xa, const void *w3t, const void *w1t, void *a1_out, float *out_f32,
                                        void *out_act, int B, int C, int L, int pitch, int dil, int pad_l,
                                        float slope_in, float slope_mid, int act_out, float slope_out, void *stream) {
  using namespace rave;
  using namespace rave::tc;
  RAVE_CHECK_ARG(xa && w3t && w1t && (out_f32 || out_act), "dilated_unit_tc: null pointer");
  RAVE_CHECK_ARG(rave_dilated_unit_tc_supported(C, L), "dilated_unit_tc: unsupported width C=%d (96, 192, 384)", C);
  RAVE_CHECK_ARG(B > 0 && L > 0 && dil >= 1 && pad_l >= 0, "dilated_unit_tc: bad shape");
  if (pitch <= 0) pitch = L;
  RAVE_CHECK_ARG(pitch >= L, "dilated_unit_tc: pitch %d < L %d", pitch, L);
  RAVE_CHECK_ARG(slope_in > 0.f && slope_in <= 1.f && slope_mid >= 0.f && slope_mid <= 1.f && slope_out >= 0.f &&
                     slope_out <= 1.f, "dilated_unit_tc: LeakyReLU slopes must lie in (0, 1]");
  RAVE_CHECK_ARG(act_out == RAVE_ACT_NONE || act_out == RAVE_ACT_LEAKY, "dilated_unit_tc: output activation %d", act_out);
  RAVE_CHECK_ARG((((uintptr_t)xa | (uintptr_t)a1_out | (uintptr_t)out_f32 | (uintptr_t)out_act) & 31) == 0 &&
                     (((uintptr_t)w3t | (uintptr_t)w1t) & 15) == 0, "dilated_unit_tc: tensors must be 32-byte aligned");
  EncodeTiledFnU enc = unit_encode_fn();
  RAVE_CHECK_ARG(enc, "dilated_unit_tc: cuTensorMapEncodeTiled not available");
  const int BK = C % 64 == 0 ? 64 : 32;
  UnitParams p;
  p.B = B; p.C = C; p.L = L; p.pitch = pitch; p.dil = dil; p.pad_l = pad_l;
  int BL = 128;
  while (BL > L && BL > 8) BL >>= 1;
  p.BL = BL; p.BB = 128 / BL;
  p.n_lt = ceil_div(L, BL);
  p.n_bg = ceil_div(B, p.BB);
  p.slope_in_inv = 1.f / slope_in;
  p.slope_mid = slope_mid;
  p.act_out = act_out;
  p.slope_out = slope_out;
  p.xa = (const __nv_bfloat16 *)xa;
  p.a1_out = (__nv_bfloat16 *)a1_out;
  p.out_f32 = out_f32;
  p.out_act = (__nv_bfloat16 *)out_act;
  const CUtensorMapSwizzle swz = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const CUtensorMapL2promotion promo = BK == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
  const int NC = C <= 256 ? C : C / 2;
  CUtensorMap ta, t3, t1;
  {
    cuuint64_t dims[4] = {(cuuint64_t)C, 1, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2, (cuuint64_t)C * 2 * pitch};
    cuuint32_t box[4] = {(cuuint32_t)BK, 1, (cuuint32_t)p.BL, (cuuint32_t)p.BB};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(xa), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RAVE_CHECK_ARG(r == CUDA_SUCCESS, "dilated_unit_tc: tensor map A encode failed (%d)", (int)r);
  }
  for (int which = 0; which < 2; ++which) {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)(which == 0 ? 3 : 1) * C};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)NC};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(which == 0 ? &t3 : &t1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void *>(which == 0 ? w3t : w1t), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RAVE_CHECK_ARG(r == CUDA_SUCCESS, "dilated_unit_tc: weight tensor map encode failed (%d)", (int)r);
  }
  cudaStream_t s = (cudaStream_t)stream;
  {
    const char *e = getenv("RAVE_UNIT_WS");
    if (C == 96 && p.BB == 1 && dil <= 12 && !(e && e[0] == '0')) {
      // weight-stationary kernel: its activation map carries the haloed box (128 + 2 dil rows), its weight maps one
      // (tap, K block) slab per box
      CUtensorMap tah, t3s, t1s;
      {
        cuuint64_t dims[4] = {(cuuint64_t)C, 1, (cuuint64_t)L, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2, (cuuint64_t)C * 2 * pitch};
        cuuint32_t box[4] = {32, 1, (cuuint32_t)(128 + 2 * dil), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tah, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(xa), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RAVE_CHECK_ARG(r == CUDA_SUCCESS, "dilated_unit_tc(ws): tensor map A encode failed (%d)", (int)r);
      }
      for (int which = 0; which < 2; ++which) {
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)(which == 0 ? 3 : 1) * C};
        cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {32, 96};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(which == 0 ? &t3s : &t1s, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                         const_cast<void *>(which == 0 ? w3t : w1t), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RAVE_CHECK_ARG(r == CUDA_SUCCESS, "dilated_unit_tc(ws): weight tensor map encode failed (%d)", (int)r);
      }
      static bool attr = false;
      if (!attr) {
        cudaError_t er = cudaFuncSetAttribute(dilated_unit_ws96_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UW_TOTAL);
        if (er != cudaSuccess) {
          set_error("dilated_unit_tc(ws): cudaFuncSetAttribute(%d bytes): %s", UW_TOTAL, cudaGetErrorString(er));
          return 2;
        }
        attr = true;
      }
      const int tiles = p.n_lt * p.n_bg;
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      // bf16 outputs leave through TMA: [B][pitch][96] viewed as (c, 1, row, b) boxes of one 32-channel K block x 128 rows
      // (rows >= L and batches >= B are clipped by the TMA unit)
      CUtensorMap tout, ta1;
      memset(&tout, 0, sizeof(tout));
      memset(&ta1, 0, sizeof(ta1));
      for (int which = 0; which < 2; ++which) {
        void *base = which == 0 ? out_act : a1_out;
        if (!base) continue;
        cuuint64_t dims[4] = {(cuuint64_t)C, 1, (cuuint64_t)L, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2, (cuuint64_t)C * 2 * pitch};
        cuuint32_t box[4] = {32, 1, 128, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(which == 0 ? &tout : &ta1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RAVE_CHECK_ARG(r == CUDA_SUCCESS, "dilated_unit_tc(ws): output tensor map encode failed (%d)", (int)r);
      }
      launch_pdl(dilated_unit_ws96_kernel, dim3(tiles < sms ? tiles : sms), dim3(U_THREADS), UW_TOTAL, s, tah, t3s, t1s,
                 tout, ta1, p);
      RAVE_CHECK_LAUNCH("dilated_unit_tc(ws)");
      return 0;
    }
  }
  switch (C) {
    case 96: return launch_unit<96, 32>(ta, t3, t1, p, s);
    case 192: return launch_unit<192, 64>(ta, t3, t1, p, s);
    case 384: return launch_unit<384, 64>(ta, t3, t1, p, s);
  }
  set_error("dilated_unit_tc: no kernel for C=%d", C);
  return 1;
}
