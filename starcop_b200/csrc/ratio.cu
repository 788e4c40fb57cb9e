// Band-ratio product with outlier-robust gain matching (starcop/data/feature_extration.py:37-56):
//   lo, hi = np.percentile(band, 5), np.percentile(band, 95)          (linear interpolation)
//   sum_b  = sum of band values with lo <= v <= hi                    (per band, per tile)
//   c = sum_bg / sum_sig ;  R = (c*sig - bg) / (bg + 1e-6) ;  R = zero_value where both < 1e-6
// Kernel 1 (one CTA per tile x band): exact order statistics by a 4-pass radix select on the
// monotone integer image of the floats (four ranks at once: k, k+1 for each percentile), then the
// inlier sum.  Kernel 2: the coalesced, vectorised elementwise product.
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

using namespace sc;
namespace cg = cooperative_groups;

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


// ------------------------------------------------------------------------------------------------
// Cluster-resident variant (tiles of <= 512 x 512 pixels): ONE thread-block cluster of 8 CTAs per tile.
// Each CTA keeps its eighth of the current band in shared memory (<= 128 KB), so every band is read from
// HBM once; the radix-select histograms of the 8 CTAs are merged through distributed shared memory
// (each CTA sums all eight histograms itself: no broadcast step, one cluster barrier per pass thanks to
// ping-pong histogram buffers).  Keys are made relative to the band's minimum and only the
// ceil(bits(max-min)/8) significant digits are scanned (3 passes for typical radiances instead of 4, and
// a well-spread first histogram instead of everything landing in one exponent bucket).  The ratio is
// applied straight from the resident signal band.  HBM traffic = bg + sig read once (+ bg re-read from
// L2) + the output = the product's algorithmic 12 B / pixel.
// ------------------------------------------------------------------------------------------------
constexpr int kClu = 8;
constexpr int kCluThreads = 1024;
constexpr int kCluMaxSlice = 32768;      // floats per CTA: 128 KB

__device__ __forceinline__ double block_sum_1024(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < kCluThreads / 32; ++i) t += red[i];      // fixed order: deterministic
  return t;
}

__global__ void __cluster_dims__(kClu, 1, 1) __launch_bounds__(kCluThreads, 1)
ratio_cluster_kernel(const float* __restrict__ bg, const float* __restrict__ sig, float* __restrict__ out, int64_t HW,
                     SelectRanks sr, float zero_value, double* __restrict__ ws) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int tile = blockIdx.x / kClu;
  extern __shared__ __align__(16) float data[];              // this CTA's slice of the current band
  __shared__ unsigned int hist[2][4][256];                   // ping-pong, read remotely by the whole cluster
  __shared__ unsigned int tot[4][256];
  __shared__ uint32_t s_mm[2];                               // local min / max key, read remotely
  __shared__ double s_part;                                  // local inlier sum, read remotely
  __shared__ uint32_t s_g[2];                                // cluster-wide min / max key
  __shared__ uint32_t prefix[4];
  __shared__ long long rank[4];
  __shared__ double red[kCluThreads / 32];
  __shared__ uint32_t wred[2][kCluThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slice = (int)((((HW + kClu - 1) / kClu) + 3) & ~(int64_t)3);
  const int64_t start = (int64_t)crank * slice;
  const int n = (int)(HW - start < 0 ? 0 : (HW - start < slice ? HW - start : slice));
  double band_sum[2] = {0.0, 0.0}, band_lo[2] = {0.0, 0.0}, band_hi[2] = {0.0, 0.0};
  int pp = 0;                                                // ping-pong index, advances once per pass

  for (int band = 0; band < 2; ++band) {
    const float* src = (band == 0 ? bg : sig) + (int64_t)tile * HW + start;
    // ---- load the slice, local min / max key ---------------------------------------------------------
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    if ((HW & 3) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(data);
      for (int i = tid; i < n / 4; i += kCluThreads) {
        const float4 v = s4[i];
        d4[i] = v;
        const uint32_t k0 = f2key(v.x), k1 = f2key(v.y), k2 = f2key(v.z), k3 = f2key(v.w);
        kmin = min(min(kmin, k0), min(min(k1, k2), k3));
        kmax = max(max(kmax, k0), max(max(k1, k2), k3));
      }
    } else {
      for (int i = tid; i < n; i += kCluThreads) {
        const float v = src[i];
        data[i] = v;
        const uint32_t k = f2key(v);
        kmin = min(kmin, k);
        kmax = max(kmax, k);
      }
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) {
      wred[0][warp] = kmin;
      wred[1][warp] = kmax;
    }
    __syncthreads();
    if (warp == 0) {
      kmin = __reduce_min_sync(0xffffffffu, wred[0][lane]);
      kmax = __reduce_max_sync(0xffffffffu, wred[1][lane]);
      if (lane == 0) {
        s_mm[0] = kmin;
        s_mm[1] = kmax;
      }
    }
    cluster.sync();
    if (warp == 0) {
      uint32_t a = 0xffffffffu, b = 0u;
      if (lane < kClu) {
        const uint32_t* r = cluster.map_shared_rank(s_mm, lane);
        a = r[0];
        b = r[1];
      }
      a = __reduce_min_sync(0xffffffffu, a);
      b = __reduce_max_sync(0xffffffffu, b);
      if (lane == 0) {
        s_g[0] = a;
        s_g[1] = b;
      }
      if (lane < 4) {
        prefix[lane] = 0;
        rank[lane] = sr.r[lane];
      }
    }
    __syncthreads();
    const uint32_t gmin = s_g[0];
    const uint32_t range = s_g[1] - gmin;
    const int nb = range ? 32 - __clz(range) : 1;
    const int npass = (nb + 7) / 8, nbp = 8 * npass;
    // ---- radix select of the four ranks over the significant digits ---------------------------------------
    for (int pass = 0; pass < npass; ++pass, pp ^= 1) {
      const int shift = nbp - 8 * (pass + 1);
      unsigned int (*h)[256] = hist[pp];
      for (int i = tid; i < 4 * 256; i += kCluThreads) (&h[0][0])[i] = 0;
      // ranks whose prefix equals an earlier rank's share that rank's histogram
      uint32_t pf[4];
      int uq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        pf[q] = prefix[q];
        uq[q] = q;
        for (int q2 = q - 1; q2 >= 0; --q2)
          if (pf[q2] == pf[q]) uq[q] = q2;
      }
      __syncthreads();
      const int n_pad = (n + 31) & ~31;
      for (int i = tid; i < n_pad; i += kCluThreads) {
        const bool valid = i < n;
        const uint32_t rk = valid ? f2key(data[i]) - gmin : 0u;
        const uint32_t b = (rk >> shift) & 255u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (uq[q] != q) continue;                                   // block-uniform
          const bool hit = valid && (pass == 0 || (rk >> (shift + 8)) == (pf[q] >> (shift + 8)));
          // the relative keys spread the first digit over all 256 buckets, so plain shared atomics are
          // conflict-light; a warp whose 32 elements all fall in ONE bucket (constant / nodata regions)
          // adds once (MATCH.ALL is a single compare; MATCH.ANY was measured ~10x slower here)
          int all_same;
          __match_all_sync(0xffffffffu, hit ? b : 0xffffffffu, &all_same);
          if (all_same) {
            if (hit && lane == 0) atomicAdd(&h[q][b], 32u);
          } else if (hit) {
            atomicAdd(&h[q][b], 1u);
          }
        }
      }
      __syncthreads();
      cluster.sync();
      {
        const int q = tid >> 8, b = tid & 255;
        unsigned int t = 0;
        if (uq[q] == q) {
#pragma unroll
          for (int r = 0; r < kClu; ++r) t += cluster.map_shared_rank(&hist[pp][q][b], r)[0];
        }
        tot[q][b] = t;
      }
      __syncthreads();
      if (warp < 4) {
        // warp q finds the bucket holding rank[q]: 8 bins per lane, warp scan of the lane totals
        const int q = warp;
        const unsigned int* tq = tot[uq[q]];
        unsigned int c8[8], ls = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          c8[k] = tq[lane * 8 + k];
          ls += c8[k];
        }
        unsigned int incl = ls;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        const long long r = rank[q];
        const long long excl = (long long)incl - ls;
        const bool mine = r >= excl && r < (long long)incl;
        if (mine) {
          long long rr = r - excl;
          int k = 0;
          for (; k < 7; ++k) {
            if (rr < (long long)c8[k]) break;
            rr -= c8[k];
          }
          rank[q] = rr;
          prefix[q] = pf[q] | ((uint32_t)(lane * 8 + k) << shift);
        }
      }
      __syncthreads();
    }
    // ---- percentile bounds (np.percentile linear interpolation) and the inlier sum -----------------------------
    const double a0 = key2f(prefix[0] + gmin), a1 = key2f(prefix[1] + gmin);
    const double b0 = key2f(prefix[2] + gmin), b1 = key2f(prefix[3] + gmin);
    const double lo = np_lerp(a0, a1, sr.t_lo), hi = np_lerp(b0, b1, sr.t_hi);
    double sacc = 0.0;
    for (int i = tid; i < n; i += kCluThreads) {
      const double v = (double)data[i];
      if (v >= lo && v <= hi) sacc += v;
    }
    sacc = block_sum_1024(sacc, red);
    if (tid == 0) s_part = sacc;
    cluster.sync();
    double total = 0.0;
    for (int r = 0; r < kClu; ++r) total += *cluster.map_shared_rank(&s_part, r);    // fixed order: deterministic
    band_sum[band] = total;
    band_lo[band] = lo;
    band_hi[band] = hi;
    cluster.sync();            // everyone has read s_part / the last histograms before they are reused
  }
  if (crank == 0 && tid == 0 && ws) {
    for (int band = 0; band < 2; ++band) {
      double* o = ws + ((int64_t)tile * 2 + band) * 3;
      o[0] = band_sum[band];
      o[1] = band_lo[band];
      o[2] = band_hi[band];
    }
  }
  // ---- apply: the signal band is resident, the background band is re-read (L2) -------------------------------
  const float c = (float)band_sum[0] / (float)band_sum[1];     // numpy: float32 scalar
  const float* bgp = bg + (int64_t)tile * HW + start;
  float* op = out + (int64_t)tile * HW + start;
  for (int i = tid; i < n; i += kCluThreads) {
    const float b = bgp[i], sv = data[i];
    float r = (c * sv - b) / (b + 1e-6f);
    if (sv < 1e-6f && b < 1e-6f) r = zero_value;
    op[i] = r;
  }
}


// ------------------------------------------------------------------------------------------------
// Sixteen-CTA cluster variant (tiles of <= 512 x 512 pixels): BOTH bands of a tile are resident, a sixteenth of each
// per CTA (2 x 64 KB of sort keys), so the two bands share every pass and every cluster barrier (6 barriers per tile
// instead of 12), each band is read from HBM exactly once and nothing is re-read from L2: HBM traffic = the
// algorithmic 12 B / pixel.  Eight tiles occupy 128 SMs instead of 64.  The data are stored as the monotone integer
// image of the floats (the key is computed once, at load), the scans use 16-byte shared-memory loads, and the select
// is the same exact radix select on the significant digits of key - min as above.
// ------------------------------------------------------------------------------------------------
constexpr int kC16 = 16;
constexpr int kC16Threads = 1024;
constexpr int kC16MaxSlice = 16384;      // keys per band per CTA: 2 x 64 KB
constexpr int kC16MaxSlice12 = 21848;    // with 12 CTAs per tile: 2 x 85.3 KB

__global__ void __launch_bounds__(kC16Threads, 1)
ratio_cluster16_kernel(const float* __restrict__ bg, const float* __restrict__ sig, float* __restrict__ out, int64_t HW,
                       SelectRanks sr, float zero_value, double* __restrict__ ws) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int csize = (int)cluster.num_blocks();               // 16, or 12 when that fills the GPU in fewer waves
  const int tile = blockIdx.x / csize;
  extern __shared__ __align__(16) uint32_t keys[];           // [2][slice] sort keys of this CTA's part of both bands
  __shared__ unsigned int hist[2][2][4][256];                // [ping-pong][band][rank][digit], read remotely
  __shared__ unsigned int tot[2][4][256];
  __shared__ uint32_t s_mm[2][2];                            // [band]{min, max} key of this CTA, read remotely
  __shared__ double s_part[2];                               // local inlier sums, read remotely
  __shared__ uint32_t s_gmin[2];
  __shared__ int s_nbp;
  __shared__ uint32_t prefix[2][4];
  __shared__ long long rank[2][4];
  __shared__ double red[2][kC16Threads / 32];
  __shared__ uint32_t wred[4][kC16Threads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slice = (int)((((HW + csize - 1) / csize) + 3) & ~(int64_t)3);
  const int64_t start = (int64_t)crank * slice;
  const int n = (int)(HW - start < 0 ? 0 : (HW - start < slice ? HW - start : slice));
  const int nfull = n / 4;                                   // complete 16-byte groups; <= 3 tail elements
  uint32_t* kb[2] = {keys, keys + slice};

  // ---- load both bands once, as keys; local min / max ------------------------------------------------------------
  {
    uint32_t kmin[2] = {0xffffffffu, 0xffffffffu}, kmax[2] = {0u, 0u};
#pragma unroll
    for (int band = 0; band < 2; ++band) {
      const float* src = (band == 0 ? bg : sig) + (int64_t)tile * HW + start;
      if ((HW & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(kb[band]);
        for (int i = tid; i < n / 4; i += kC16Threads) {
          const float4 v = __ldg(s4 + i);
          const uint4 k = make_uint4(f2key(v.x), f2key(v.y), f2key(v.z), f2key(v.w));
          d4[i] = k;
          kmin[band] = min(min(kmin[band], k.x), min(min(k.y, k.z), k.w));
          kmax[band] = max(max(kmax[band], k.x), max(max(k.y, k.z), k.w));
        }
      } else {
        for (int i = tid; i < n; i += kC16Threads) {
          const uint32_t k = f2key(src[i]);
          kb[band][i] = k;
          kmin[band] = min(kmin[band], k);
          kmax[band] = max(kmax[band], k);
        }
      }
    }
#pragma unroll
    for (int band = 0; band < 2; ++band) {
      const uint32_t a = __reduce_min_sync(0xffffffffu, kmin[band]), b = __reduce_max_sync(0xffffffffu, kmax[band]);
      if (lane == 0) {
        wred[2 * band][warp] = a;
        wred[2 * band + 1][warp] = b;
      }
    }
    __syncthreads();
    if (warp < 4) {
      const uint32_t v = wred[warp][lane];
      const uint32_t r = (warp & 1) ? __reduce_max_sync(0xffffffffu, v) : __reduce_min_sync(0xffffffffu, v);
      if (lane == 0) s_mm[warp >> 1][warp & 1] = r;
    }
    cluster.sync();
    if (warp < 2) {                                            // warp = band: cluster-wide min / max
      uint32_t a = 0xffffffffu, b = 0u;
      if (lane < csize) {
        const uint32_t* r = cluster.map_shared_rank(&s_mm[warp][0], lane);
        a = r[0];
        b = r[1];
      }
      a = __reduce_min_sync(0xffffffffu, a);
      b = __reduce_max_sync(0xffffffffu, b);
      if (lane == 0) {
        s_gmin[warp] = a;
        const uint32_t range = b - a;
        wred[0][warp] = range ? 32 - __clz(range) : 1;         // significant bits of this band
      }
      if (lane < 4) {
        prefix[warp][lane] = 0;
        rank[warp][lane] = sr.r[lane];
      }
    }
    __syncthreads();
    if (tid == 0) {
      const int nb = max((int)wred[0][0], (int)wred[0][1]);
      s_nbp = 8 * ((nb + 7) / 8);
    }
    __syncthreads();
  }
  const int nbp = s_nbp, npass = nbp / 8;
  const uint32_t gmin[2] = {s_gmin[0], s_gmin[1]};
  int pp = 0;
  // ---- radix select of the four ranks of both bands over the significant digits ---------------------------------------
  for (int pass = 0; pass < npass; ++pass, pp ^= 1) {
    const int shift = nbp - 8 * (pass + 1);
    for (int i = tid; i < 2 * 4 * 256; i += kC16Threads) (&hist[pp][0][0][0])[i] = 0;
    uint32_t pf[2][4];
    int uq[2][4];
#pragma unroll
    for (int band = 0; band < 2; ++band)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        pf[band][q] = prefix[band][q];
        uq[band][q] = q;
        for (int q2 = q - 1; q2 >= 0; --q2)
          if (pf[band][q2] == pf[band][q]) uq[band][q] = q2;   // ranks with equal prefixes share a histogram
      }
    __syncthreads();
#pragma unroll
    for (int band = 0; band < 2; ++band) {
      const uint4* k4 = reinterpret_cast<const uint4*>(kb[band]);
      const uint32_t gm = gmin[band];
      unsigned int* hb = &hist[pp][band][0][0];
      // one element: a warp whose 32 elements all fall in ONE bucket (constant / nodata regions) adds once
      auto add = [&](unsigned int* h, uint32_t b, bool hit) {
        int all_same;
        __match_all_sync(0xffffffffu, hit ? b : 0xffffffffu, &all_same);
        if (all_same) {
          if (hit && lane == 0) atomicAdd(h + b, 32u);
        } else if (hit) {
          atomicAdd(h + b, 1u);
        }
      };
      if (pass == 0) {
        // every element counts, the four ranks share one histogram (all prefixes are empty)
        for (int i0 = 0; i0 < nfull; i0 += kC16Threads) {      // warp-uniform trip count
          const int i = i0 + tid;
          const bool ok = i < nfull;
          uint4 kv = make_uint4(0, 0, 0, 0);
          if (ok) kv = k4[i];
          add(hb, ((kv.x - gm) >> shift) & 255u, ok);
          add(hb, ((kv.y - gm) >> shift) & 255u, ok);
          add(hb, ((kv.z - gm) >> shift) & 255u, ok);
          add(hb, ((kv.w - gm) >> shift) & 255u, ok);
        }
        if (warp == 0) {
          const int i = 4 * nfull + lane;
          const bool ok = lane < 4 && i < n;
          add(hb, (((ok ? kb[band][i] : gm) - gm) >> shift) & 255u, ok);
        }
      } else {
        // only the elements whose upper digits equal one of the (<= 4, usually 2) distinct prefixes matter: most
        // warps skip a group after a subtract, a shift and the compares
        uint32_t up[4];
        int sl[4], nu = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (uq[band][q] == q) {
            up[nu] = pf[band][q] >> (shift + 8);
            sl[nu] = q;
            ++nu;
          }
        auto visit = [&](uint32_t key, bool ok) {
          const uint32_t rk = key - gm, u = rk >> (shift + 8);
          bool any = false;
          for (int j = 0; j < nu; ++j) any |= ok && u == up[j];
          if (!__any_sync(0xffffffffu, any)) return;
          for (int j = 0; j < nu; ++j) add(hb + sl[j] * 256, (rk >> shift) & 255u, ok && u == up[j]);
        };
        for (int i0 = 0; i0 < nfull; i0 += kC16Threads) {
          const int i = i0 + tid;
          const bool ok = i < nfull;
          uint4 kv = make_uint4(0, 0, 0, 0);
          if (ok) kv = k4[i];
          visit(kv.x, ok);
          visit(kv.y, ok);
          visit(kv.z, ok);
          visit(kv.w, ok);
        }
        if (warp == 0) {
          const int i = 4 * nfull + lane;
          const bool ok = lane < 4 && i < n;
          visit(ok ? kb[band][i] : gm, ok);
        }
      }
    }
    __syncthreads();
    cluster.sync();
    for (int o = tid; o < 2 * 4 * 256; o += kC16Threads) {    // every CTA sums the histograms of the whole cluster itself
      const int band = o >> 10, q = (o >> 8) & 3, b = o & 255;
      unsigned int t = 0;
      if (uq[band][q] == q) {
        // all remote loads in flight together (one DSMEM latency, not csize of them), summed in rank order
        unsigned int v[kC16];
#pragma unroll
        for (int r = 0; r < kC16; ++r) v[r] = r < csize ? cluster.map_shared_rank(&hist[pp][band][q][b], r)[0] : 0u;
#pragma unroll
        for (int r = 0; r < kC16; ++r) t += v[r];
      }
      tot[band][q][b] = t;
    }
    __syncthreads();
    if (warp < 8) {
      // warp (band, q) finds the bucket holding rank[band][q]: 8 bins per lane, warp scan of the lane totals
      const int band = warp >> 2, q = warp & 3;
      const unsigned int* tq = tot[band][uq[band][q]];
      unsigned int c8[8], ls = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        c8[k] = tq[lane * 8 + k];
        ls += c8[k];
      }
      unsigned int incl = ls;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const long long r = rank[band][q];
      const long long excl = (long long)incl - ls;
      if (r >= excl && r < (long long)incl) {
        long long rr = r - excl;
        int k = 0;
        for (; k < 7; ++k) {
          if (rr < (long long)c8[k]) break;
          rr -= c8[k];
        }
        rank[band][q] = rr;
        prefix[band][q] = pf[band][q] | ((uint32_t)(lane * 8 + k) << shift);
      }
    }
    __syncthreads();
  }
  // ---- percentile bounds (np.percentile linear interpolation) and the inlier sums -----------------------------------
  double lo[2], hi[2], sacc[2] = {0.0, 0.0};
#pragma unroll
  for (int band = 0; band < 2; ++band) {
    const double a0 = key2f(prefix[band][0] + gmin[band]), a1 = key2f(prefix[band][1] + gmin[band]);
    const double b0 = key2f(prefix[band][2] + gmin[band]), b1 = key2f(prefix[band][3] + gmin[band]);
    lo[band] = np_lerp(a0, a1, sr.t_lo);
    hi[band] = np_lerp(b0, b1, sr.t_hi);
    for (int i = tid; i < n; i += kC16Threads) {
      const double v = (double)key2f(kb[band][i]);
      if (v >= lo[band] && v <= hi[band]) sacc[band] += v;
    }
    sacc[band] = warp_sum(sacc[band]);
    if (lane == 0) red[band][warp] = sacc[band];
  }
  __syncthreads();
  if (tid < 2) {
    double t = 0.0;
    for (int i = 0; i < kC16Threads / 32; ++i) t += red[tid][i];          // fixed order: deterministic
    s_part[tid] = t;
  }
  cluster.sync();
  double band_sum[2] = {0.0, 0.0};
  for (int r = 0; r < csize; ++r) {                                         // fixed order: deterministic
    const double* rp = cluster.map_shared_rank(&s_part[0], r);
    band_sum[0] += rp[0];
    band_sum[1] += rp[1];
  }
  if (crank == 0 && tid == 0 && ws) {
    for (int band = 0; band < 2; ++band) {
      double* o = ws + ((int64_t)tile * 2 + band) * 3;
      o[0] = band_sum[band];
      o[1] = lo[band];
      o[2] = hi[band];
    }
  }
  // ---- apply from the resident keys ---------------------------------------------------------------------------------
  const float c = (float)band_sum[0] / (float)band_sum[1];     // numpy: float32 scalar
  float* op = out + (int64_t)tile * HW + start;
  if ((HW & 3) == 0) {
    const uint4* b4 = reinterpret_cast<const uint4*>(kb[0]);
    const uint4* s4 = reinterpret_cast<const uint4*>(kb[1]);
    float4* o4 = reinterpret_cast<float4*>(op);
    for (int i = tid; i < n / 4; i += kC16Threads) {
      const uint4 bk = b4[i], sk = s4[i];
      const float b[4] = {key2f(bk.x), key2f(bk.y), key2f(bk.z), key2f(bk.w)};
      const float sv[4] = {key2f(sk.x), key2f(sk.y), key2f(sk.z), key2f(sk.w)};
      float r[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        r[e] = (c * sv[e] - b[e]) / (b[e] + 1e-6f);
        if (sv[e] < 1e-6f && b[e] < 1e-6f) r[e] = zero_value;
      }
      o4[i] = make_float4(r[0], r[1], r[2], r[3]);
    }
  } else {
    for (int i = tid; i < n; i += kC16Threads) {
      const float b = key2f(kb[0][i]), sv = key2f(kb[1][i]);
      float r = (c * sv - b) / (b + 1e-6f);
      if (sv < 1e-6f && b < 1e-6f) r = zero_value;
      op[i] = r;
    }
  }
  cluster.sync();            // no CTA exits while its shared memory may still be read remotely
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
  // measured on B200 (scripts/ratio_bench.py): one cluster per tile wins while the tiles do not fill the GPU
  // (T = 8: 99 us vs 259 us); from ~64 tiles on, two independent CTAs per tile keep more SMs busy
  // sixteen-CTA clusters (both bands resident: every byte crosses HBM once) whenever the tile fits and the device
  // co-schedules such a cluster; otherwise the eight-CTA / single-CTA kernels below
  // measured on B200 (scripts/ratio_bench.py, 512 x 512 tiles): clusters win while the tiles do not fill the GPU with
  // independent CTAs (T = 8: 89 us); from ~100 tiles on, two independent CTAs per tile keep more SMs busy
  if (HW <= (int64_t)12 * kC16MaxSlice12 && (T <= 96 || getenv("STARCOP_RATIO_CLUSTER")) && !getenv("STARCOP_RATIO_NOCLUSTER") &&
      !getenv("STARCOP_RATIO_CLUSTER8")) {
    // cluster size: 16 CTAs (2 x 64 KB of keys each) or 12 (2 x 86 KB) -- whichever needs fewer cluster-waves x work per
    // CTA for this T (B200 co-schedules 7 sixteen-CTA clusters: 8 tiles would take two waves)
    static int active[2] = {-1, -1};                          // max co-resident clusters of 16 / 12 CTAs (per process)
    const int sizes[2] = {kC16, 12};
    auto config = [&](cudaLaunchConfig_t& c, cudaLaunchAttribute& a, int cl, unsigned grid) {
      const int slice = (int)((((HW + cl - 1) / cl) + 3) & ~(int64_t)3);
      c = cudaLaunchConfig_t{};
      c.gridDim = dim3(grid);
      c.blockDim = dim3(kC16Threads);
      c.dynamicSmemBytes = (size_t)2 * slice * sizeof(uint32_t);
      c.stream = st;
      a.id = cudaLaunchAttributeClusterDimension;
      a.val.clusterDim.x = cl;
      a.val.clusterDim.y = 1;
      a.val.clusterDim.z = 1;
      c.attrs = &a;
      c.numAttrs = 1;
    };
    if (active[0] < 0) {
      active[0] = active[1] = 0;
      if (cudaFuncSetAttribute(ratio_cluster16_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
          cudaFuncSetAttribute(ratio_cluster16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               2 * kC16MaxSlice12 * (int)sizeof(uint32_t)) == cudaSuccess) {
        for (int k = 0; k < 2; ++k) {
          cudaLaunchConfig_t qc;
          cudaLaunchAttribute qa;
          config(qc, qa, sizes[k], (unsigned)sizes[k]);
          qc.dynamicSmemBytes = (size_t)2 * (k == 0 ? kC16MaxSlice : kC16MaxSlice12) * sizeof(uint32_t);
          int ncl = 0;
          if (cudaOccupancyMaxActiveClusters(&ncl, ratio_cluster16_kernel, &qc) == cudaSuccess) active[k] = ncl;
        }
      }
      (void)cudaGetLastError();
      if (getenv("STARCOP_RATIO_DEBUG")) fprintf(stderr, "ratio clusters co-resident: %d x 16 CTAs, %d x 12 CTAs\n", active[0], active[1]);
    }
    int best = -1;
    double best_cost = 0.0;
    for (int k = 0; k < 2; ++k) {
      if (active[k] < 1 || HW > (int64_t)sizes[k] * (k == 0 ? kC16MaxSlice : kC16MaxSlice12)) continue;
      const double cost = (double)((T + active[k] - 1) / active[k]) / sizes[k];
      if (best < 0 || cost < best_cost) {
        best = k;
        best_cost = cost;
      }
    }
    if (getenv("STARCOP_RATIO_CLUSTER12") && active[1] >= 1) best = 1;
    if (best >= 0) {
      cudaLaunchConfig_t cfg;
      cudaLaunchAttribute at;
      config(cfg, at, sizes[best], (unsigned)(T * sizes[best]));
      cudaError_t e = cudaLaunchKernelEx(&cfg, ratio_cluster16_kernel, bg, sig, out, HW, sr, zero_value, (double*)workspace);
      if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }
      return check_launch();
    }
  }
  if (HW <= (int64_t)kClu * kCluMaxSlice && (T < 64 || getenv("STARCOP_RATIO_CLUSTER") || getenv("STARCOP_RATIO_CLUSTER8")) && !getenv("STARCOP_RATIO_NOCLUSTER")) {
    const int slice = (int)((((HW + kClu - 1) / kClu) + 3) & ~(int64_t)3);
    const size_t smem = (size_t)slice * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(ratio_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kCluMaxSlice * (int)sizeof(float));
      if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }
      attr_set = true;
    }
    ratio_cluster_kernel<<<T * kClu, kCluThreads, smem, st>>>(bg, sig, out, HW, sr, zero_value, (double*)workspace);
    return check_launch();
  }
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
