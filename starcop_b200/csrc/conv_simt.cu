// fp32-FMA implicit-GEMM convolutions (NHWC).  This is the exact-fp32 path: it carries the parity
// claim against the reference's fp32 PyTorch network and serves every layer the tcgen05 path does
// not (stem Cin=4 stride 2, channel counts that are not multiples of 16).  The depthwise 3x3 kernels live in dwconv.cu.
#include "common.cuh"

using namespace sc;

// ------------------------------------------------------------------------------------------------
// dense conv fprop: M = output pixels, N = Cout, K = KH*KW*Cin (tap-major, channel-minor)
// 256 threads, block tile BM x BN, thread tile TM x TN, BK = 16, register-prefetch double buffer.
// ------------------------------------------------------------------------------------------------
template <typename T, int BM, int BN, int TM, int TN, bool VEC>
__global__ void __launch_bounds__(256)
conv_fprop_kernel(const T* __restrict__ x, int ldx, const float* __restrict__ wp, const float* __restrict__ bias,
                  T* __restrict__ y, int ldy, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                  int stride, int pad, int Ho, int Wo, int accumulate) {
  constexpr int BK = 16;
  constexpr int TX = BN / TN;            // threads along N
  static_assert((BM / TM) * TX == 256, "tile/thread mismatch");
  constexpr int A_PER_T = BM * BK / 256; // elements of A per thread (8 for BM=128, 16 for BM=256)
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int64_t M = (int64_t)N * Ho * Wo;
  const int K = KH * KW * Cin;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A loader: thread owns pixel row a_m and A_PER_T consecutive k starting at a_k
  constexpr int KGROUPS = BK / 8;                 // 2 groups of 8 k per pixel
  constexpr int ROWS_PER_PASS = 256 / KGROUPS;    // 128 pixel rows per pass
  constexpr int PASSES = BM / ROWS_PER_PASS;      // 1 (BM=128) or 2 (BM=256)
  static_assert(A_PER_T == 8 * PASSES, "A loader geometry");
  const int a_r = tid % ROWS_PER_PASS;       // consecutive threads -> consecutive pixels: conflict-free As stores
  const int a_kg = tid / ROWS_PER_PASS;
  int a_n[PASSES], a_h[PASSES], a_w[PASSES];
  bool a_ok[PASSES];
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps) {
    int64_t m = m0 + a_r + ps * ROWS_PER_PASS;
    a_ok[ps] = m < M;
    int64_t mm = a_ok[ps] ? m : 0;
    a_w[ps] = (int)(mm % Wo);
    int64_t t = mm / Wo;
    a_h[ps] = (int)(t % Ho);
    a_n[ps] = (int)(t / Ho);
  }
  // B loader: BK x BN floats, 4 per thread
  constexpr int B_VECS = BK * BN / 4;
  float areg[PASSES][8];
  float4 breg[(B_VECS + 255) / 256];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      int kb = k0 + a_kg * 8;
      if (VEC) {
        // Cin % 8 == 0: the 8 k of this group share one tap
        bool ok = a_ok[ps] && kb < K;
        int tap = ok ? kb / Cin : 0, ci = ok ? kb - tap * Cin : 0;
        int kh = tap / KW, kw = tap - kh * KW;
        int ih = a_h[ps] * stride - pad + kh, iw = a_w[ps] * stride - pad + kw;
        ok = ok && ih >= 0 && ih < H && iw >= 0 && iw < W;
        if (ok) {
          f8 v = load8<T>(x + (((int64_t)a_n[ps] * H + ih) * W + iw) * ldx + ci);
#pragma unroll
          for (int i = 0; i < 8; ++i) areg[ps][i] = v.v[i];
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) areg[ps][i] = 0.f;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int k = kb + i;
          float v = 0.f;
          if (a_ok[ps] && k < K) {
            int tap = k / Cin, ci = k - tap * Cin;
            int kh = tap / KW, kw = tap - kh * KW;
            int ih = a_h[ps] * stride - pad + kh, iw = a_w[ps] * stride - pad + kw;
            if (ih >= 0 && ih < H && iw >= 0 && iw < W)
              v = to_f<T>(x[(((int64_t)a_n[ps] * H + ih) * W + iw) * ldx + ci]);
          }
          areg[ps][i] = v;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < (B_VECS + 255) / 256; ++j) {
      int v = tid + j * 256;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < B_VECS) {
        int kk = v / (BN / 4), nn = (v % (BN / 4)) * 4;
        int k = k0 + kk, n = n0 + nn;
        if (k < K) {
          const float* src = wp + (int64_t)k * Cout + n;
          if (n + 3 < Cout && (Cout % 4 == 0)) {
            r = *reinterpret_cast<const float4*>(src);
          } else {
            if (n < Cout) r.x = src[0];
            if (n + 1 < Cout) r.y = src[1];
            if (n + 2 < Cout) r.z = src[2];
            if (n + 3 < Cout) r.w = src[3];
          }
        }
      }
      breg[j] = r;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps)
#pragma unroll
      for (int i = 0; i < 8; ++i) As[a_kg * 8 + i][a_r + ps * ROWS_PER_PASS] = areg[ps][i];
#pragma unroll
    for (int j = 0; j < (B_VECS + 255) / 256; ++j) {
      int v = tid + j * 256;
      if (v < B_VECS) {
        int kk = v / (BN / 4), nn = (v % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[kk][nn]) = breg[j];
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // Two-level summation: each BK-long chunk is accumulated in `part` and folded into `acc` once, so
  // the rounding error grows like sqrt(BK) + sqrt(K/BK) instead of sqrt(K) (K reaches 12384 in the
  // decoder).  This is the fp32 PARITY path: its noise on the zero-true-gradient parameters (anything
  // feeding a BatchNorm) is compared with PyTorch's blocked CPU kernels by the tests.
  float part[TM][TN];
  load_tiles(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    store_tiles();
    __syncthreads();
    if (k0 + BK < K) load_tiles(k0 + BK);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] += part[i][j];
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= Cout) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      T* o = y + m * ldy + n;
      if (accumulate) v += to_f<T>(*o);
      *o = from_f<T>(v);
    }
  }
}

template <typename T, int BM, int BN, int TM, int TN>
static int launch_fprop(const void* x, int ldx, const float* wp, const float* bias, void* y, int ldy, int N,
                        int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int accumulate,
                        cudaStream_t st) {
  int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  int64_t M = (int64_t)N * Ho * Wo;
  dim3 grid(ceil_div(M, BM), ceil_div(Cout, BN));
  bool vec = (Cin % 8 == 0) && (ldx % 8 == 0);
  if (vec)
    conv_fprop_kernel<T, BM, BN, TM, TN, true><<<grid, 256, 0, st>>>((const T*)x, ldx, wp, bias, (T*)y, ldy, N, H, W,
                                                                       Cin, Cout, KH, KW, stride, pad, Ho, Wo, accumulate);
  else
    conv_fprop_kernel<T, BM, BN, TM, TN, false><<<grid, 256, 0, st>>>((const T*)x, ldx, wp, bias, (T*)y, ldy, N, H, W,
                                                                        Cin, Cout, KH, KW, stride, pad, Ho, Wo, accumulate);
  return check_launch();
}

extern "C" int sc_conv_fprop(const void* x, int ldx, const float* w_packed, const float* bias, void* y, int ldy,
                             int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int dtype,
                             int accumulate, void* stream) {
  if (!x || !w_packed || !y || N <= 0 || Cin <= 0 || Cout <= 0 || stride < 1) return SC_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  SC_DISPATCH_DTYPE(dtype, {
    if (Cout > 32)
      return launch_fprop<T, 128, 64, 8, 4>(x, ldx, w_packed, bias, y, ldy, N, H, W, Cin, Cout, KH, KW, stride, pad, accumulate, st);
    if (Cout > 16)
      return launch_fprop<T, 128, 32, 4, 4>(x, ldx, w_packed, bias, y, ldy, N, H, W, Cin, Cout, KH, KW, stride, pad, accumulate, st);
    return launch_fprop<T, 256, 16, 4, 4>(x, ldx, w_packed, bias, y, ldy, N, H, W, Cin, Cout, KH, KW, stride, pad, accumulate, st);
  });
  return SC_OK;
}

// ------------------------------------------------------------------------------------------------
// dense conv wgrad: per tap, dW[ci][co] = sum_p x[p@tap][ci] * dy[p][co]; 64x64 tile, split over
// pixels.  Each pixel split writes its own OIHW-shaped partial gradient into the workspace and a
// second kernel sums the splits in order (deterministic: no floating-point atomics).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const T* __restrict__ x, int ldx, const T* __restrict__ dy, int lddy, float* __restrict__ dw,
                  float* __restrict__ partials, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
                  int pad, int Ho, int Wo, int ci_tiles, int psplit) {
  constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
  __shared__ float As[BK][BM + 4];   // [pixel][ci]
  __shared__ float Bs[BK][BN + 4];   // [pixel][co]
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int ci0 = (blockIdx.x % ci_tiles) * BM;
  const int tap = blockIdx.x / ci_tiles;
  const int co0 = blockIdx.y * BN;
  const int kh = tap / KW, kw = tap - kh * KW;
  const int64_t P = (int64_t)N * Ho * Wo;
  const int64_t chunk = ((P + psplit - 1) / psplit + BK - 1) / BK * BK;
  const int64_t p_begin = (int64_t)blockIdx.z * chunk;
  int64_t p_end = p_begin + chunk;
  if (p_end > P) p_end = P;

  // loader: 16 pixels x 64 channels = 1024 elements, 4 per thread: thread -> (pixel lr, 4 channels lc)
  const int lr = tid / 16, lc = (tid % 16) * 4;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int64_t p0 = p_begin; p0 < p_end; p0 += BK) {
    int64_t p = p0 + lr;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < p_end) {
      int wo = (int)(p % Wo);
      int64_t t = p / Wo;
      int ho = (int)(t % Ho);
      int n = (int)(t / Ho);
      int ih = ho * stride - pad + kh, iw = wo * stride - pad + kw;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
        const T* xp = x + (((int64_t)n * H + ih) * W + iw) * ldx;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (ci0 + lc + i < Cin) a[i] = to_f<T>(xp[ci0 + lc + i]);
      }
      const T* dp = dy + p * lddy;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (co0 + lc + i < Cout) b[i] = to_f<T>(dp[co0 + lc + i]);
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lr][lc]) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(&Bs[lr][lc]) = make_float4(b[0], b[1], b[2], b[3]);
    __syncthreads();
    // two-level summation (see conv_fprop_kernel): 16-pixel chunks folded into the running sums
    float part[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
      float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) part[i][j] = fmaf(aa[i], bb[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] += part[i][j];
  }
  const int KK = KH * KW;
  // one split: this block is the only writer of its elements; several: split z owns copy z of the workspace
  float* dst = psplit == 1 ? dw : partials + (int64_t)blockIdx.z * Cout * Cin * KK;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int ci = ci0 + ty * TM + i;
    if (ci >= Cin) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int co = co0 + tx * TN + j;
      if (co >= Cout) continue;
      const int64_t o = ((int64_t)co * Cin + ci) * KK + tap;
      if (psplit == 1) dst[o] += acc[i][j];
      else dst[o] = acc[i][j];
    }
  }
}

__global__ void conv_wgrad_sum_kernel(const float* __restrict__ partials, int psplit, int64_t n, float* __restrict__ dw) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = partials[i];
    for (int z = 1; z < psplit; ++z) s += partials[(int64_t)z * n + i];     // fixed order
    dw[i] += s;
  }
}

static int simt_wgrad_split(int N, int Ho, int Wo, int Cin, int Cout, int KH, int KW) {
  int64_t P = (int64_t)N * Ho * Wo;
  int ci_tiles = ceil_div(Cin, 64), co_tiles = ceil_div(Cout, 64);
  int base = ci_tiles * KH * KW * co_tiles;
  int psplit = (kNumSMs * 4 + base - 1) / base;
  int64_t max_split = (P + 255) / 256;
  if (psplit > max_split) psplit = (int)max_split;
  if (psplit < 1) psplit = 1;
  // every split non-empty (chunks are multiples of 16 pixels)
  const int64_t chunk = ((P + psplit - 1) / psplit + 15) / 16 * 16;
  return (int)((P + chunk - 1) / chunk);
}

extern "C" int64_t sc_conv_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad) {
  int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  int psplit = simt_wgrad_split(N, Ho, Wo, Cin, Cout, KH, KW);
  return psplit > 1 ? (int64_t)psplit * Cout * Cin * KH * KW * sizeof(float) : 0;
}

extern "C" int sc_conv_wgrad(const void* x, int ldx, const void* dy, int lddy, float* dw_oihw, float* workspace, int N,
                             int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int dtype, void* stream) {
  if (!x || !dy || !dw_oihw || N <= 0) return SC_ERR_BAD_ARG;
  int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  int ci_tiles = ceil_div(Cin, 64), co_tiles = ceil_div(Cout, 64);
  const int psplit = simt_wgrad_split(N, Ho, Wo, Cin, Cout, KH, KW);
  if (psplit > 1 && !workspace) return SC_ERR_BAD_ARG;
  dim3 grid(ci_tiles * KH * KW, co_tiles, psplit);
  cudaStream_t st = (cudaStream_t)stream;
  SC_DISPATCH_DTYPE(dtype, (conv_wgrad_kernel<T><<<grid, 256, 0, st>>>(
                               (const T*)x, ldx, (const T*)dy, lddy, dw_oihw, workspace, N, H, W, Cin, Cout, KH, KW,
                               stride, pad, Ho, Wo, ci_tiles, psplit)));
  int rc = check_launch();
  if (rc != SC_OK || psplit == 1) return rc;
  const int64_t n = (int64_t)Cout * Cin * KH * KW;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  conv_wgrad_sum_kernel<<<(int)blocks, 256, 0, st>>>(workspace, psplit, n, dw_oihw);
  return check_launch();
}
