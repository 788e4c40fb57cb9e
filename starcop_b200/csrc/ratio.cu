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
