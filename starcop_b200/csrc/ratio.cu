// Band-ratio product with outlier-robust gain matching (starcop/data/feature_extration.py:37-56):
//   lo, hi = np.percentile(band, 5), np.percentile(band, 95)          (linear interpolation)
//   sum_b  = sum of band values with lo <= v <= hi                    (per band, per tile)
//   c = sum_bg / sum_sig ;  R = (c*sig - bg) / (bg + 1e-6) ;  R = zero_value where both < 1e-6
// Kernel 1 (one CTA per tile x band): exact order statistics by a 4-pass radix select on the
// monotone integer image of the floats (four ranks at once: k, k+1 for each percentile), then the
// inlier sum.  Kernel 2: the coalesced, vectorised elementwise product.
#include "common.cuh"

using namespace sc;

namespace {

constexpr int kSelThreads = 1024;

__device__ __forceinline__ uint32_t f2key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// np.percentile's _lerp (numpy/lib/_function_base_impl.py), evaluated in double
__device__ __forceinline__ double np_lerp(double a, double b, double t) {
  double d = b - a;
  double r = a + d * t;
  if (t >= 0.5) r = b - d * (1.0 - t);
  return r;
}

struct SelectRanks {
  int64_t r[4];     // k_lo, k_lo+1, k_hi, k_hi+1 (0-based ranks in the sorted band)
  double t_lo, t_hi;
};

__global__ void __launch_bounds__(kSelThreads)
ratio_select_kernel(const float* __restrict__ bg, const float* __restrict__ sig, int64_t HW, SelectRanks sr,
                    double* __restrict__ ws) {
  const int band = blockIdx.x, tile = blockIdx.y;
  const float* src = (band == 0 ? bg : sig) + (int64_t)tile * HW;
  __shared__ unsigned int hist[4][256];
  __shared__ uint32_t prefix[4];
  __shared__ int64_t rank[4];
  __shared__ double s_red[kSelThreads / 32];
  const int tid = threadIdx.x;
  if (tid < 4) {
    prefix[tid] = 0;
    rank[tid] = sr.r[tid];
  }
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 4 * 256; i += kSelThreads) (&hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    uint32_t pf[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) pf[q] = prefix[q];
    for (int64_t i = tid; i < HW; i += kSelThreads) {
      uint32_t k = f2key(src[i]);
      uint32_t b = (k >> shift) & 255u;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if ((k & himask) == pf[q]) atomicAdd(&hist[q][b], 1u);
    }
    __syncthreads();
    if (tid < 4) {
      int64_t r = rank[tid];
      uint32_t b = 0;
      for (; b < 256; ++b) {
        unsigned int c = hist[tid][b];
        if (r < (int64_t)c) break;
        r -= c;
      }
      rank[tid] = r;
      prefix[tid] |= b << shift;
    }
    __syncthreads();
  }
  // prefix[q] is now the exact key of the rank-q order statistic
  const double a0 = key2f(prefix[0]), a1 = key2f(prefix[1]), b0 = key2f(prefix[2]), b1 = key2f(prefix[3]);
  const double lo = np_lerp(a0, a1, sr.t_lo), hi = np_lerp(b0, b1, sr.t_hi);
  double s = 0.0;
  for (int64_t i = tid; i < HW; i += kSelThreads) {
    double v = (double)src[i];
    if (v >= lo && v <= hi) s += v;
  }
  s = warp_sum(s);
  if ((tid & 31) == 0) s_red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < kSelThreads / 32; ++i) t += s_red[i];
    double* o = ws + ((int64_t)tile * 2 + band) * 3;
    o[0] = t;
    o[1] = lo;
    o[2] = hi;
  }
}

__global__ void ratio_apply_kernel(const float* __restrict__ bg, const float* __restrict__ sig, float* __restrict__ out,
                                   int64_t HW, int64_t total, const double* __restrict__ ws, float zero_value) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t tile = i / HW;
    // numpy: c is a float32 scalar (sum of float32 / sum of float32)
    float c = (float)ws[(tile * 2 + 0) * 3] / (float)ws[(tile * 2 + 1) * 3];
    float b = bg[i], s = sig[i];
    float r = (c * s - b) / (b + 1e-6f);
    if (s < 1e-6f && b < 1e-6f) r = zero_value;
    out[i] = r;
  }
}

}  // namespace

extern "C" int64_t sc_ratio_workspace_bytes(int T, int64_t HW) {
  (void)HW;
  return (int64_t)T * 2 * 3 * sizeof(double);
}

extern "C" int sc_ratio_product(const float* bg, const float* sig, float* out, int T, int64_t HW, float percentile,
                                float zero_value, void* workspace, void* stream) {
  if (!bg || !sig || !out || !workspace || T <= 0 || HW < 2 || percentile < 0.f || percentile > 50.f) return SC_ERR_BAD_ARG;
  SelectRanks sr;
  // numpy: virtual index = (n - 1) * (q / 100), floor + fraction, all in float64
  double vlo = (double)(HW - 1) * ((double)percentile / 100.0);
  double vhi = (double)(HW - 1) * ((100.0 - (double)percentile) / 100.0);
  int64_t klo = (int64_t)floor(vlo), khi = (int64_t)floor(vhi);
  sr.t_lo = vlo - (double)klo;
  sr.t_hi = vhi - (double)khi;
  sr.r[0] = klo;
  sr.r[1] = klo + 1 < HW ? klo + 1 : HW - 1;
  sr.r[2] = khi;
  sr.r[3] = khi + 1 < HW ? khi + 1 : HW - 1;
  cudaStream_t st = (cudaStream_t)stream;
  ratio_select_kernel<<<dim3(2, T), kSelThreads, 0, st>>>(bg, sig, HW, sr, (double*)workspace);
  int64_t total = (int64_t)T * HW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  ratio_apply_kernel<<<blocks, 256, 0, st>>>(bg, sig, out, HW, total, (const double*)workspace, zero_value);
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// A12: multiple linear regression of a target band on K background bands over one tile
// (feature_extration.py:58-124, sklearn LinearRegression with intercept), then the reconstruction
// recon = X.coef + intercept.  Normal equations of the CENTRED data accumulated in fp64.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kMlrMaxK = 9;
constexpr int kMlrThreads = 256;

// pass 1: per-tile sums  s_x[k], s_y, then pass 2 uses the means: G = sum (x-mx)(x-mx)^T, r = sum (x-mx)(y-my)
__global__ void __launch_bounds__(kMlrThreads)
mlr_means_kernel(const float* __restrict__ bands, const float* __restrict__ target, int K, int64_t HW, double* __restrict__ ws) {
  const int tile = blockIdx.y;
  const float* xb = bands + (int64_t)tile * K * HW;
  const float* yb = target + (int64_t)tile * HW;
  double s[kMlrMaxK + 1];
  for (int k = 0; k <= K; ++k) s[k] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
    for (int k = 0; k < K; ++k) s[k] += (double)xb[(int64_t)k * HW + i];
    s[K] += (double)yb[i];
  }
  double* out = ws + (int64_t)tile * 128;          // [0..K] sums
  for (int k = 0; k <= K; ++k) {
    double v = warp_sum(s[k]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&out[k], v);
  }
}

__global__ void __launch_bounds__(kMlrThreads)
mlr_gram_kernel(const float* __restrict__ bands, const float* __restrict__ target, int K, int64_t HW, double* __restrict__ ws) {
  const int tile = blockIdx.y;
  const float* xb = bands + (int64_t)tile * K * HW;
  const float* yb = target + (int64_t)tile * HW;
  double* base = ws + (int64_t)tile * 128;
  double mean[kMlrMaxK + 1];
  for (int k = 0; k <= K; ++k) mean[k] = base[k] / (double)HW;
  double g[(kMlrMaxK + 1) * (kMlrMaxK + 2) / 2];   // packed lower triangle over [x_0..x_{K-1}, y]
  const int NT = (K + 1) * (K + 2) / 2;
  for (int e = 0; e < NT; ++e) g[e] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
    double v[kMlrMaxK + 1];
    for (int k = 0; k < K; ++k) v[k] = (double)xb[(int64_t)k * HW + i] - mean[k];
    v[K] = (double)yb[i] - mean[K];
    int e = 0;
    for (int a = 0; a <= K; ++a)
      for (int b = 0; b <= a; ++b) g[e++] += v[a] * v[b];
  }
  double* out = base + 16;                          // packed Gram matrix
  for (int e = 0; e < NT; ++e) {
    double v = warp_sum(g[e]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&out[e], v);
  }
}

// one thread per tile: solve the K x K system (Cholesky), write coef[0..K) and intercept
__global__ void mlr_solve_kernel(int K, int64_t HW, double* __restrict__ ws, int T) {
  const int tile = blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= T) return;
  double* base = ws + (int64_t)tile * 128;
  double* G = base + 16;
  double L[kMlrMaxK][kMlrMaxK], rhs[kMlrMaxK], z[kMlrMaxK], c[kMlrMaxK];
  auto tri = [](int a, int b) { return a * (a + 1) / 2 + b; };
  for (int a = 0; a < K; ++a) {
    rhs[a] = G[tri(K, a)];
    for (int b = 0; b <= a; ++b) {
      double sdot = G[tri(a, b)];
      for (int k = 0; k < b; ++k) sdot -= L[a][k] * L[b][k];
      L[a][b] = (a == b) ? sqrt(sdot > 0.0 ? sdot : 1e-300) : sdot / L[b][b];
    }
  }
  for (int a = 0; a < K; ++a) {
    double sdot = rhs[a];
    for (int k = 0; k < a; ++k) sdot -= L[a][k] * z[k];
    z[a] = sdot / L[a][a];
  }
  for (int a = K - 1; a >= 0; --a) {
    double sdot = z[a];
    for (int k = a + 1; k < K; ++k) sdot -= L[k][a] * c[k];
    c[a] = sdot / L[a][a];
  }
  double icpt = base[K] / (double)HW;
  for (int a = 0; a < K; ++a) icpt -= c[a] * (base[a] / (double)HW);
  double* o = base + 96;                            // [coef..., intercept]
  for (int a = 0; a < K; ++a) o[a] = c[a];
  o[K] = icpt;
}

__global__ void mlr_recon_kernel(const float* __restrict__ bands, int K, int64_t HW, const double* __restrict__ ws,
                                 float* __restrict__ recon, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t tile = i / HW, px = i - tile * HW;
    const double* o = ws + tile * 128 + 96;
    const float* xb = bands + tile * K * HW;
    double acc = o[K];
    for (int k = 0; k < K; ++k) acc += o[k] * (double)xb[(int64_t)k * HW + px];
    recon[i] = (float)acc;
  }
}

__global__ void zero_override_kernel(const float* __restrict__ ref, float* __restrict__ out, int64_t n, float value) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (ref[i] == 0.0f) out[i] = value;
}

// A17: EMIT input rescale (emit_tools/emit_dataset.py:62-101): crop to x32, clip(mf/240,0,2)*1750,
// clip(rgb/20,0,2)*60, nan_to_num
__global__ void emit_rescale_kernel(const float* __restrict__ magic, const float* __restrict__ rgb, float* __restrict__ out,
                                    int H, int W, int H32, int W32) {
  int64_t total = (int64_t)4 * H32 * W32;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int w = (int)(i % W32);
    int h = (int)((i / W32) % H32);
    int c = (int)(i / ((int64_t)W32 * H32));
    float v;
    if (c == 0) {
      v = magic[(int64_t)h * W + w] / 240.f;
      v = fminf(fmaxf(v, 0.f), 2.f) * 1750.f;
      if (isnan(magic[(int64_t)h * W + w])) v = nanf("");
    } else {
      float r = rgb[((int64_t)(c - 1) * H + h) * W + w];
      v = fminf(fmaxf(r / 20.f, 0.f), 2.f) * 60.f;
      if (isnan(r)) v = nanf("");
    }
    // torch.nan_to_num: nan -> 0, +-inf -> +-float max
    if (isnan(v)) v = 0.f;
    else if (isinf(v)) v = v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    out[i] = v;
  }
}
}  // namespace

extern "C" int64_t sc_mlr_workspace_bytes(int T) { return (int64_t)T * 128 * sizeof(double); }

extern "C" int sc_mlr_reconstruct(const float* bands, const float* target, float* recon, int T, int K, int64_t HW,
                                  void* workspace, void* stream) {
  if (!bands || !target || !recon || !workspace || T <= 0 || K < 1 || K > kMlrMaxK || HW < K + 1) return SC_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(workspace, 0, (size_t)sc_mlr_workspace_bytes(T), st);
  int bx = (int)((HW + kMlrThreads * 8 - 1) / (kMlrThreads * 8));
  int cap = (kNumSMs * 4 + T - 1) / T;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  mlr_means_kernel<<<dim3(bx, T), kMlrThreads, 0, st>>>(bands, target, K, HW, (double*)workspace);
  mlr_gram_kernel<<<dim3(bx, T), kMlrThreads, 0, st>>>(bands, target, K, HW, (double*)workspace);
  mlr_solve_kernel<<<(T + 31) / 32, 32, 0, st>>>(K, HW, (double*)workspace, T);
  int64_t total = (int64_t)T * HW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  mlr_recon_kernel<<<blocks, 256, 0, st>>>(bands, K, HW, (const double*)workspace, recon, total);
  return check_launch();
}

extern "C" int sc_zero_override(const float* ref, float* out, int64_t n, float value, void* stream) {
  if (!ref || !out || n <= 0) return SC_ERR_BAD_ARG;
  int blocks = (int)((n + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  zero_override_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ref, out, n, value);
  return check_launch();
}

extern "C" int sc_emit_rescale(const float* magic, const float* rgb, float* out, int H, int W, void* stream) {
  if (!magic || !rgb || !out || H < 32 || W < 32) return SC_ERR_BAD_ARG;
  int H32 = H / 32 * 32, W32 = W / 32 * 32;
  int64_t total = (int64_t)4 * H32 * W32;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  emit_rescale_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(magic, rgb, out, H, W, H32, W32);
  return check_launch();
}
