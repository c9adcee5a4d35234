// PQMF 16-band analysis / synthesis kernels (fp32 FMA, polyphase-in-shared-memory form).
//
// Reference: CachedPQMF.forward / .inverse, rave/pqmf.py:279-294 (F.pad + F.conv1d with a
// [16,1,513] stride-16 filter, resp. a [16,16,33] filter followed by *16, channel flip and a
// channel->time interleave) and reverse_half, rave/pqmf.py:13-17.
//
// Layout: the decimate-by-16 FIR reads x at stride 16, which is a 16-way shared-memory bank
// conflict if the signal tile is stored linearly.  The tile is therefore stored de-interleaved by
// phase, xs[p][q] = x[16 q + p], so that frame n / tap (16 jq + jp) reads xs[jp][n + jq]:
// consecutive frames -> consecutive words.  Each thread owns 4 consecutive frames x 4 bands and
// slides a 36-word register window over jq, so one 16-byte shared load feeds 16 FMAs.
#include "common.cuh"
#include "tc_common.cuh"

namespace rave {

constexpr int PQ_M = 16;        // bands
constexpr int PQ_FR = 256;      // frames (analysis) / band-rate samples (synthesis) per CTA
constexpr int PQ_JQ = 33;       // taps per phase (ceil(528/16))
constexpr int PQ_XQ = PQ_FR + PQ_JQ - 1;  // 288 window columns
constexpr int PQ_XP = 292;      // padded pitch (multiple of 4 floats)
constexpr int PQ_SMEM = (PQ_JQ * 16 * 16 + 16 * PQ_XP) * sizeof(float);

// y[b][k][n] = sgn(k,n) * sum_j taps[k][j] * x[b][16 n + j - pad_l]
__global__ void __launch_bounds__(256)
pqmf_analysis_kernel(const float *__restrict__ x, const float *__restrict__ taps,
                     float *__restrict__ y, int T, int Lout, int ntaps, int pad_l, int flip_sign) {
  extern __shared__ __align__(16) float smem[];
  float *ts = smem;                      // [528][16]  ts[j*16 + k]
  float *xs = smem + PQ_JQ * 16 * 16;    // [16][PQ_XP]
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * PQ_FR;

  for (int i = tid; i < PQ_JQ * 16 * 16; i += 256) {
    int j = i >> 4, k = i & 15;
    ts[i] = (j < ntaps) ? __ldg(taps + k * ntaps + j) : 0.f;
  }
  const float *xb = x + (size_t)b * T;
  const long base = (long)16 * n0 - pad_l;
  for (int i = tid; i < 16 * PQ_XQ; i += 256) {
    long gi = base + i;
    float v = (gi >= 0 && gi < T) ? __ldg(xb + gi) : 0.f;
    xs[(i & 15) * PQ_XP + (i >> 4)] = v;
  }
  __syncthreads();

  const int np = tid & 63;   // frame quad
  const int bg = tid >> 6;   // band group (4 bands)
  float acc[4][4];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[f][k] = 0.f;

  for (int jp = 0; jp < 16; ++jp) {
    float xr[36];
    const float4 *xrow = reinterpret_cast<const float4 *>(xs + jp * PQ_XP + 4 * np);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      float4 v = xrow[i];
      xr[4 * i + 0] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int jq = 0; jq < PQ_JQ; ++jq) {
      const float4 t4 = *reinterpret_cast<const float4 *>(ts + (jq * 16 + jp) * 16 + bg * 4);
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const float xv = xr[jq + f];
        acc[f][0] = fmaf(xv, t4.x, acc[f][0]);
        acc[f][1] = fmaf(xv, t4.y, acc[f][1]);
        acc[f][2] = fmaf(xv, t4.z, acc[f][2]);
        acc[f][3] = fmaf(xv, t4.w, acc[f][3]);
      }
    }
  }

  const int n = n0 + 4 * np;
  if (n >= Lout) return;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int k = bg * 4 + kk;
    float v[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      // reverse_half: odd bands, even time steps (n is a multiple of 4, so f parity == time parity)
      const bool neg = flip_sign && (k & 1) && !(f & 1);
      v[f] = neg ? -acc[f][kk] : acc[f][kk];
    }
    float *yp = y + ((size_t)b * PQ_M + k) * Lout + n;
    if (n + 3 < Lout && (Lout & 3) == 0) {
      *reinterpret_cast<float4 *>(yp) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int f = 0; f < 4; ++f)
        if (n + f < Lout) yp[f] = v[f];
    }
  }
}

// out[b][16 t + 15 - m] = scale * sum_c sum_j w[m][c][j] * sgn(c,tau) * x[b][c][tau], tau = t + j - pad_l
__global__ void __launch_bounds__(256)
pqmf_synthesis_kernel(const float *__restrict__ x, const float *__restrict__ w,
                      float *__restrict__ out, int L, int K, int pad_l, float scale, int flip_sign) {
  extern __shared__ __align__(16) float smem[];
  float *ws = smem;                      // [16 c][33 j][16 m]
  float *xs = smem + PQ_JQ * 16 * 16;    // [16 c][PQ_XP]
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * PQ_FR;

  for (int i = tid; i < PQ_JQ * 16 * 16; i += 256) {
    int m = i & 15, cj = i >> 4;
    int c = cj / PQ_JQ, j = cj - c * PQ_JQ;
    ws[i] = (j < K) ? __ldg(w + ((size_t)m * 16 + c) * K + j) : 0.f;
  }
  for (int i = tid; i < 16 * PQ_XQ; i += 256) {
    int c = i / PQ_XQ, q = i - c * PQ_XQ;
    int tau = t0 + q - pad_l;
    float v = 0.f;
    if (tau >= 0 && tau < L) {
      v = __ldg(x + ((size_t)b * 16 + c) * L + tau);
      if (flip_sign && (c & 1) && !(tau & 1)) v = -v;
    }
    xs[c * PQ_XP + q] = v;
  }
  __syncthreads();

  const int tp = tid & 63;
  const int bg = tid >> 6;
  float acc[4][4];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[f][k] = 0.f;

  for (int c = 0; c < 16; ++c) {
    float xr[36];
    const float4 *xrow = reinterpret_cast<const float4 *>(xs + c * PQ_XP + 4 * tp);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      float4 v = xrow[i];
      xr[4 * i + 0] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < PQ_JQ; ++j) {
      const float4 w4 = *reinterpret_cast<const float4 *>(ws + (c * PQ_JQ + j) * 16 + bg * 4);
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const float xv = xr[j + f];
        acc[f][0] = fmaf(xv, w4.x, acc[f][0]);
        acc[f][1] = fmaf(xv, w4.y, acc[f][1]);
        acc[f][2] = fmaf(xv, w4.z, acc[f][2]);
        acc[f][3] = fmaf(xv, w4.w, acc[f][3]);
      }
    }
  }

  const int t = t0 + 4 * tp;
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    if (t + f >= L) break;
    // m = 4 bg + mm  ->  time slot 15 - m : the 4 slots 12-4bg .. 15-4bg in reversed band order
    float4 v = make_float4(scale * acc[f][3], scale * acc[f][2], scale * acc[f][1], scale * acc[f][0]);
    float *op = out + (size_t)b * 16 * L + (size_t)16 * (t + f) + 12 - 4 * bg;
    *reinterpret_cast<float4 *>(op) = v;
  }
}


// =============================================================================================
// Factorised ("fast") PQMF kernels.
//
// A pseudo-QMF bank is a cosine-modulated prototype (rave/pqmf.py:32-52): hk[k][n] = 2 h[n] cos((2k+1) pi/(2M) (n - c) + phi_k),
// and the modulating cosine changes sign every 2M = 32 taps.  Hence every filter table the four PQMF operators
// use (analysis taps, synthesis weights, and the re-indexed tables of their adjoints) is RANK ONE PER RESIDUE:
//
//        T[k][32 i + r] = C[k][r] * Q[r][i]            (k: band, r: tap index mod 32, i: tap index div 32)
//
// (the host verifies this numerically on the tables it is given -- rave_b200/pqmf.py::_factorise -- and keeps the
// dense kernels above for a bank that is not).  The dense 512-MAC-per-sample FIR then splits into
//   analysis form :  u_r[n] = sum_i Q[r][i] x[16 n + 32 i + r - pad]   (polyphase filter, 34 MAC / sample)
//                    y[k][n] = sum_r C[k][r] u_r[n]                       (16 x 32 modulation, 32 MAC / sample)
//   synthesis form:  v_r[tau] = sum_c C[c][r] s(c,tau) x[c][tau]          (32 x 16 demodulation, 32 MAC / sample)
//                    out[16 t + 15 - m] = scale sum_{e<2} sum_i Q[16e+m][i] v_{16e+m}[t + 2i + e - pad]   (33 MAC / sample)
// i.e. 8x fewer FMAs than the dense form (66 instead of 512/528 per sample): ~140 MFMA for a 32 x 65536 batch, which
// is what lets the operator approach its 16.8 MB HBM floor on the fp32 pipes.  Summation order differs from the
// dense conv, so results agree to fp32 rounding (~3e-7 rel-L2), not bit for bit.
//
// Staging: the signal tile (and the modulation table) is brought into shared memory by ONE bulk TMA copy
// (cp.async.bulk, mbarrier complete_tx) per tile row; out-of-range borders are zero-filled by the threads.
// =============================================================================================
constexpr int PF_NI = 17;                        // taps per residue: ceil(544 / 32)
constexpr int PF_NF = 256;                       // frames per CTA (analysis form)
constexpr int PF_XT = 16 * PF_NF + 32 * PF_NI;   // 4640 samples of signal per tile
constexpr int PF_UP = 36;                        // pitch of us[n][r] (conflict-free float4 reads at 36 n)
constexpr int PF_A_SMEM = (PF_XT + PF_NF * PF_UP + 32 * 16) * 4 + 16;

__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::smem_u32(smem_dst)),
               "l"(gmem), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

// y[b][k][n] = sgn(k,n) * sum_r C[k][r] * sum_i Q[r][i] * x[b][16 n + 32 i + r - pad_l]
//   Ct: [32 r][16 k] (modulation, transposed), Qt: [PF_NI i][32 r] (prototype polyphase components)
__global__ void __launch_bounds__(256, 3)
pqmf_analysis_fast_kernel(const float *__restrict__ x, const float *__restrict__ Ct, const float *__restrict__ Qt,
                          float *__restrict__ y, int T, int Lout, int pad_l, int flip_sign) {
  extern __shared__ __align__(16) float smem[];
  float *xs = smem;                         // [PF_XT]
  float *us = xs + PF_XT;                   // [PF_NF][PF_UP]
  float *cs = us + PF_NF * PF_UP;           // [32][16]
  uint64_t *bar = reinterpret_cast<uint64_t *>(cs + 32 * 16);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * PF_NF;
  const float *xb = x + (size_t)b * T;
  const long base = (long)16 * n0 - pad_l;

  // ---- stage 1: signal tile -> xs (bulk TMA for the in-range part, zero fill for the borders)
  long lo = base < 0 ? 0 : base, hi = base + PF_XT > T ? T : base + PF_XT;
  if (hi < lo) hi = lo;
  const bool bulk_ok = (((uintptr_t)(xb + lo)) & 15) == 0 && ((lo - base) & 3) == 0;
  const int n_bulk = bulk_ok ? (int)((hi - lo) & ~3L) : 0;          // floats moved by the bulk copy
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    tc::mbar_arrive_expect_tx(bar, (uint32_t)(n_bulk + 32 * 16) * 4u);
    if (n_bulk) bulk_g2s(xs + (lo - base), xb + lo, (uint32_t)n_bulk * 4u, bar);
    bulk_g2s(cs, Ct, 32 * 16 * 4, bar);
  }
  {
    const int i0 = (int)(lo - base), i1 = i0 + n_bulk, i2 = (int)(hi - base);
    for (int i = tid; i < i0; i += 256) xs[i] = 0.f;
    for (int i = i1 + tid; i < i2; i += 256) xs[i] = __ldg(xb + base + i);      // unaligned / ragged remainder
    for (int i = i2 + tid; i < PF_XT; i += 256) xs[i] = 0.f;
  }
  float q[PF_NI];
#pragma unroll
  for (int i = 0; i < PF_NI; ++i) q[i] = __ldg(Qt + i * 32 + lane);
  __syncthreads();
  tc::mbar_wait(bar, 0);

  // ---- stage 2: polyphase components.  lane = residue r; frames of one parity form a plain 17-tap FIR over the
  //      stride-32 sequence s[m] = xs[32 m + 16 par + r]  (frame n = 2 m' + par reads s[m' .. m' + 16])
  {
    const int par = warp & 1;
    const float *sp = xs + 16 * par + lane;
#pragma unroll 1
    for (int blk = 0; blk < 2; ++blk) {
      const int m0 = (warp >> 1) * 32 + blk * 16;
      float sw[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) sw[j] = sp[32 * (m0 + j)];
#pragma unroll
      for (int f = 0; f < 16; ++f) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < PF_NI; ++i) acc = fmaf(q[i], sw[f + i], acc);
        us[(2 * (m0 + f) + par) * PF_UP + lane] = acc;
      }
    }
  }
  __syncthreads();

  // ---- stage 3: modulation.  thread = frames {fq, fq+64, fq+128, fq+192} x bands 4 bg .. 4 bg + 3
  const int fq = tid & 63, bg = tid >> 6;
  float acc[4][4];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[f][k] = 0.f;
#pragma unroll
  for (int r = 0; r < 32; r += 4) {
    float4 u[4], c[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) u[f] = *reinterpret_cast<const float4 *>(us + (fq + 64 * f) * PF_UP + r);
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = *reinterpret_cast<const float4 *>(cs + (r + j) * 16 + 4 * bg);
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const float uv[4] = {u[f].x, u[f].y, u[f].z, u[f].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[f][0] = fmaf(uv[j], c[j].x, acc[f][0]);
        acc[f][1] = fmaf(uv[j], c[j].y, acc[f][1]);
        acc[f][2] = fmaf(uv[j], c[j].z, acc[f][2]);
        acc[f][3] = fmaf(uv[j], c[j].w, acc[f][3]);
      }
    }
  }
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const int n = n0 + fq + 64 * f;
    if (n >= Lout) continue;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = 4 * bg + kk;
      const bool neg = flip_sign && (k & 1) && !(n & 1);       // reverse_half: odd bands, even time steps
      y[((size_t)b * PQ_M + k) * Lout + n] = neg ? -acc[f][kk] : acc[f][kk];
    }
  }
}

constexpr int PS_NT = 224;                       // band-rate output positions per CTA (synthesis form)
constexpr int PS_NQ = 256;                       // input positions per tile: PS_NT + 32
constexpr int PS_VP = 290;                       // pitch of vs[r][q]  (290 = 2 mod 32: lanes (m, parity) hit 32 banks)
constexpr int PS_S_SMEM = (16 * PS_NQ + 32 * PS_VP + 16 * 32) * 4 + 16;

// out[b][16 t + 15 - m] = scale * sum_{e<2} sum_i Q[16e+m][i] * v_{16e+m}[t + 2i + e - pad_l],
//   v_r[tau] = sum_c C[c][r] * sgn(c,tau) * x[b][c][tau];   Cc: [16 c][32 r], Qt: [PF_NI i][32 r]
__global__ void __launch_bounds__(256, 3)
pqmf_synthesis_fast_kernel(const float *__restrict__ x, const float *__restrict__ Cc, const float *__restrict__ Qt,
                           float *__restrict__ out, int L, int pad_l, float scale, int flip_sign) {
  extern __shared__ __align__(16) float smem[];
  float *xs = smem;                         // [16 c][PS_NQ]
  float *vs = xs + 16 * PS_NQ;              // [32 r][PS_VP]
  float *cs = vs + 32 * PS_VP;              // [16 c][32 r]
  uint64_t *bar = reinterpret_cast<uint64_t *>(cs + 16 * 32);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * PS_NT;
  const int tau0 = t0 - pad_l;                                 // input position of tile column 0
  const float *xb = x + (size_t)b * 16 * L;

  // ---- stage 1: 16 band rows -> xs (one bulk TMA copy per row where aligned; borders zero-filled)
  int lo = tau0 < 0 ? 0 : tau0, hi = tau0 + PS_NQ > L ? L : tau0 + PS_NQ;
  if (hi < lo) hi = lo;
  const bool bulk_ok = ((((uintptr_t)xb) | ((uintptr_t)L * 4)) & 15) == 0 && (lo & 3) == 0 && ((lo - tau0) & 3) == 0;
  const int n_bulk = bulk_ok ? ((hi - lo) & ~3) : 0;
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    tc::mbar_arrive_expect_tx(bar, (uint32_t)(16 * n_bulk + 16 * 32) * 4u);
    if (n_bulk)
      for (int c = 0; c < 16; ++c) bulk_g2s(xs + c * PS_NQ + (lo - tau0), xb + (size_t)c * L + lo, (uint32_t)n_bulk * 4u, bar);
    bulk_g2s(cs, Cc, 16 * 32 * 4, bar);
  }
  {
    const int i0 = lo - tau0, i1 = i0 + n_bulk, i2 = hi - tau0;
    for (int i = tid; i < 16 * PS_NQ; i += 256) {
      const int c = i >> 8, qq = i & (PS_NQ - 1);
      if (qq < i0 || qq >= i2) xs[i] = 0.f;
      else if (qq >= i1) xs[i] = __ldg(xb + (size_t)c * L + tau0 + qq);
    }
  }
  __syncthreads();
  tc::mbar_wait(bar, 0);

  // ---- stage A: demodulation, thread = input position q (all 32 residues)
  {
    const int qq = tid;
    const float sgn_odd = (flip_sign && !((tau0 + qq) & 1)) ? -1.f : 1.f;   // reverse_half on the way in
    float xv[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float v = xs[c * PS_NQ + qq];
      xv[c] = (c & 1) ? v * sgn_odd : v;
    }
#pragma unroll 1
    for (int r0 = 0; r0 < 32; r0 += 16) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 c4 = *reinterpret_cast<const float4 *>(cs + c * 32 + r0 + j);
          acc[j] = fmaf(xv[c], c4.x, acc[j]);
          acc[j + 1] = fmaf(xv[c], c4.y, acc[j + 1]);
          acc[j + 2] = fmaf(xv[c], c4.z, acc[j + 2]);
          acc[j + 3] = fmaf(xv[c], c4.w, acc[j + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) vs[(r0 + j) * PS_VP + qq] = acc[j];
    }
  }
  __syncthreads();

  // ---- stage B: prototype polyphase filters.  warp = 28 consecutive t; lane = (m, parity of t): outputs of one
  //      parity are a 17-tap FIR over the stride-2 subsequence of v_m plus a 16-tap one over v_{16+m}
  {
    const int m = lane & 15, par = lane >> 4;
    const int tb = 28 * warp + par;                               // tile-relative t of f = 0
    float acc[14];
#pragma unroll
    for (int f = 0; f < 14; ++f) acc[f] = 0.f;
    {
      float qa[PF_NI], a[30];
#pragma unroll
      for (int i = 0; i < PF_NI; ++i) qa[i] = __ldg(Qt + i * 32 + m);
      const float *vp = vs + m * PS_VP + tb;
#pragma unroll
      for (int j = 0; j < 30; ++j) a[j] = vp[2 * j];
#pragma unroll
      for (int f = 0; f < 14; ++f)
#pragma unroll
        for (int i = 0; i < PF_NI; ++i) acc[f] = fmaf(qa[i], a[f + i], acc[f]);
    }
    {
      float qb[PF_NI - 1], a[29];
#pragma unroll
      for (int i = 0; i < PF_NI - 1; ++i) qb[i] = __ldg(Qt + i * 32 + 16 + m);
      const float *vp = vs + (16 + m) * PS_VP + tb + 1;
#pragma unroll
      for (int j = 0; j < 29; ++j) a[j] = vp[2 * j];
#pragma unroll
      for (int f = 0; f < 14; ++f)
#pragma unroll
        for (int i = 0; i < PF_NI - 1; ++i) acc[f] = fmaf(qb[i], a[f + i], acc[f]);
    }
    float *ob = out + (size_t)b * 16 * L;
#pragma unroll
    for (int f = 0; f < 14; ++f) {
      const int t = t0 + tb + 2 * f;
      if (t < L) ob[(size_t)16 * t + 15 - m] = scale * acc[f];
    }
  }
}

}  // namespace rave

extern "C" int rave_pqmf_analysis_fwd(const float *x, const float *taps, float *y, int B, int T,
                                      int Lout, int ntaps, int pad_l, int flip_sign, void *stream) {
  using namespace rave;
  RAVE_CHECK_ARG(x && taps && y, "pqmf_analysis: null pointer");
  RAVE_CHECK_ARG(ntaps > 0 && ntaps <= PQ_JQ * 16, "pqmf_analysis: ntaps %d > %d", ntaps, PQ_JQ * 16);
  RAVE_CHECK_ARG(B > 0 && T > 0 && Lout > 0 && pad_l >= 0, "pqmf_analysis: bad shape");
  RAVE_CHECK_ARG(B <= 65535, "pqmf_analysis: B %d > 65535", B);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pqmf_analysis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PQ_SMEM);
    attr = true;
  }
  dim3 grid(ceil_div(Lout, PQ_FR), B);
  pqmf_analysis_kernel<<<grid, 256, PQ_SMEM, (cudaStream_t)stream>>>(x, taps, y, T, Lout, ntaps,
                                                                    pad_l, flip_sign);
  RAVE_CHECK_LAUNCH("pqmf_analysis");
  return 0;
}

extern "C" int rave_pqmf_synthesis_fwd(const float *x, const float *w, float *out, int B, int L,
                                       int K, int pad_l, float scale, int flip_sign, void *stream) {
  using namespace rave;
  RAVE_CHECK_ARG(x && w && out, "pqmf_synthesis: null pointer");
  RAVE_CHECK_ARG(K > 0 && K <= PQ_JQ, "pqmf_synthesis: K %d > %d", K, PQ_JQ);
  RAVE_CHECK_ARG(B > 0 && L > 0 && pad_l >= 0, "pqmf_synthesis: bad shape");
  RAVE_CHECK_ARG(B <= 65535, "pqmf_synthesis: B %d > 65535", B);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pqmf_synthesis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PQ_SMEM);
    attr = true;
  }
  dim3 grid(ceil_div(L, PQ_FR), B);
  pqmf_synthesis_kernel<<<grid, 256, PQ_SMEM, (cudaStream_t)stream>>>(x, w, out, L, K, pad_l, scale,
                                                                     flip_sign);
  RAVE_CHECK_LAUNCH("pqmf_synthesis");
  return 0;
}

// Factorised forms (see the kernel comments): Ct [32][16] / Cc [16][32] modulation, Qt [17][32] polyphase prototype.
extern "C" int rave_pqmf_analysis_fast(const float *x, const float *Ct, const float *Qt, float *y, int B, int T,
                                       int Lout, int pad_l, int flip_sign, void *stream) {
  using namespace rave;
  RAVE_CHECK_ARG(x && Ct && Qt && y, "pqmf_analysis_fast: null pointer");
  RAVE_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && Lout > 0 && pad_l >= 0, "pqmf_analysis_fast: bad shape");
  RAVE_CHECK_ARG((((uintptr_t)Ct) & 15) == 0, "pqmf_analysis_fast: tables must be 16-byte aligned");
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pqmf_analysis_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_A_SMEM);
    attr = true;
  }
  dim3 grid(ceil_div(Lout, PF_NF), B);
  pqmf_analysis_fast_kernel<<<grid, 256, PF_A_SMEM, (cudaStream_t)stream>>>(x, Ct, Qt, y, T, Lout, pad_l, flip_sign);
  RAVE_CHECK_LAUNCH("pqmf_analysis_fast");
  return 0;
}

extern "C" int rave_pqmf_synthesis_fast(const float *x, const float *Cc, const float *Qt, float *out, int B, int L,
                                        int pad_l, float scale, int flip_sign, void *stream) {
  using namespace rave;
  RAVE_CHECK_ARG(x && Cc && Qt && out, "pqmf_synthesis_fast: null pointer");
  RAVE_CHECK_ARG(B > 0 && B <= 65535 && L > 0 && pad_l >= 0, "pqmf_synthesis_fast: bad shape");
  RAVE_CHECK_ARG((((uintptr_t)Cc) & 15) == 0, "pqmf_synthesis_fast: tables must be 16-byte aligned");
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pqmf_synthesis_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_S_SMEM);
    attr = true;
  }
  dim3 grid(ceil_div(L, PS_NT), B);
  pqmf_synthesis_fast_kernel<<<grid, 256, PS_S_SMEM, (cudaStream_t)stream>>>(x, Cc, Qt, out, L, pad_l, scale, flip_sign);
  RAVE_CHECK_LAUNCH("pqmf_synthesis_fast");
  return 0;
}
