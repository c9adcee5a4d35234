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
