// Bandwidth-bound kernels of the STARCOP hot path: input normalisation, training-mode BatchNorm
// (statistics / finalize / apply / backward), gradient routing, segmentation head, fused weighted
// BCE + decisions + confusion counts, Adam.  All NHWC, 8-channel vectors, fp32 math, fp64 sums.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace sc {
thread_local cudaError_t g_last_error = cudaSuccess;
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("STARCOP_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
}
using namespace sc;

extern "C" int sc_abi_version(void) { return 1; }
// host-side check of the tile-index division constants (common.cuh, FastDiv): the same formula the kernels evaluate
extern "C" unsigned int sc_debug_fastdiv(unsigned int x, unsigned int d) {
  if (d == 0) return 0;
  const FastDiv f = make_fastdiv(d);
  return (unsigned int)(((((unsigned long long)f.m * x) >> 32) + x) >> f.s);
}
extern "C" const char* sc_last_cuda_error(void) { return cudaGetErrorString(sc::g_last_error); }

// ------------------------------------------------------------------------------------------------
// A1 normalize_x: clamp((x-off)/fac, lo, hi).float()  (normalizer_module.py:134-135)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void normalize_pack_kernel(const float* __restrict__ x, const double* __restrict__ off,
                                      const double* __restrict__ fac, const double* __restrict__ lo,
                                      const double* __restrict__ hi, int f64_path, int C, int64_t HW,
                                      int64_t total_px, T* __restrict__ out_nhwc, int ld_out,
                                      float* __restrict__ out_nchw) {
  sc::pdl_wait();
  // one thread per pixel: NCHW reads are coalesced per channel plane, NHWC writes are ld_out wide
  const bool fast = !f64_path && C <= 8 && ld_out == 8 && out_nhwc && (reinterpret_cast<uintptr_t>(out_nhwc) & (8 * sizeof(T) - 1)) == 0;
  // the fp32 copies of the parameters are made ONCE per thread: sixteen F2F.F32.F64 per pixel on the fp64 pipe
  // (plus eight 2-byte stores) held this kernel at 1.3 TB/s
  float foff[8], ffac[8], flo[8], fhi[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const bool ok = fast && c < C;
    foff[c] = ok ? (float)off[c] : 0.f;
    ffac[c] = ok ? (float)fac[c] : 1.f;
    flo[c] = ok ? (float)lo[c] : 0.f;
    fhi[c] = ok ? (float)hi[c] : 0.f;
  }
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total_px;
       p += (int64_t)gridDim.x * blockDim.x) {
    int64_t b, s;
    if (total_px < (1ll << 31)) {
      const uint32_t b32 = (uint32_t)p / (uint32_t)HW;
      b = b32;
      s = (uint32_t)p - b32 * (uint32_t)HW;
    } else {
      b = p / HW;
      s = p - b * HW;
    }
    if (fast) {
      // the network's input (4 products padded to 8 channels): one 16 / 32-byte store per pixel
      f8 o;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float r = 0.f;
        if (c < C) {
          const float v = x[(b * C + c) * HW + s];
          const float d = (v - foff[c]) / ffac[c];                 // IEEE division, like ATen
          r = d < flo[c] ? flo[c] : (d > fhi[c] ? fhi[c] : d);
          if (out_nchw) out_nchw[(b * C + c) * HW + s] = r;
        }
        o.v[c] = r;
      }
      store8<T>(out_nhwc + p * 8, o);
      continue;
    }
    for (int c = 0; c < ld_out; ++c) {
      float r = 0.f;
      if (c < C) {
        float v = x[(b * C + c) * HW + s];
        if (f64_path) {
          // bit0: offsets are float64 (x-off promotes to f64); bit1: factors are float64
          double num = (f64_path & 1) ? (double)v - off[c] : (double)(v - (float)off[c]);
          double d = num / fac[c];
          d = d < lo[c] ? lo[c] : (d > hi[c] ? hi[c] : d);   // NaN propagates like torch.clamp
          r = (float)d;
        } else {
          float d = (v - (float)off[c]) / (float)fac[c];     // IEEE division, like ATen
          float l = (float)lo[c], h = (float)hi[c];
          r = d < l ? l : (d > h ? h : d);
        }
        if (out_nchw) out_nchw[(b * C + c) * HW + s] = r;
      }
      if (out_nhwc) out_nhwc[p * ld_out + c] = from_f<T>(r);
    }
  }
}

// the network-input case (C <= 8 products padded to 8 NHWC channels, fp32 arithmetic, H*W a multiple of 4): FOUR
// consecutive pixels per thread -- one 16-byte load per channel plane, four vector stores -- so a warp moves 512-byte
// runs in both directions (the pixel-per-thread kernel above reads 4 bytes per lane and plane: 104 us for 134 MB)
template <typename T>
__global__ void __launch_bounds__(256)
normalize_pack4_kernel(const float* __restrict__ x, const double* __restrict__ off, const double* __restrict__ fac,
                       const double* __restrict__ lo, const double* __restrict__ hi, int C, int64_t HW, int64_t total_q,
                       T* __restrict__ out_nhwc, float* __restrict__ out_nchw) {
  sc::pdl_wait();
  float foff[8], ffac[8], flo[8], fhi[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const bool ok = c < C;
    foff[c] = ok ? (float)off[c] : 0.f;
    ffac[c] = ok ? (float)fac[c] : 1.f;
    flo[c] = ok ? (float)lo[c] : 0.f;
    fhi[c] = ok ? (float)hi[c] : 0.f;
  }
  const int64_t HWq = HW >> 2;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total_q; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = q / HWq, s = (q - b * HWq) << 2;       // image, first of the four pixels
    f8 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) o[j].v[c] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < C) {
        const float4 v = *reinterpret_cast<const float4*>(x + (b * C + c) * HW + s);
        const float vv[4] = {v.x, v.y, v.z, v.w};
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d = (vv[j] - foff[c]) / ffac[c];             // IEEE division, like ATen
          r[j] = d < flo[c] ? flo[c] : (d > fhi[c] ? fhi[c] : d);  // NaN propagates like torch.clamp
          o[j].v[c] = r[j];
        }
        if (out_nchw) *reinterpret_cast<float4*>(out_nchw + (b * C + c) * HW + s) = make_float4(r[0], r[1], r[2], r[3]);
      }
    }
    T* op = out_nhwc + (b * HW + s) * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) store8<T>(op + j * 8, o[j]);
  }
}

extern "C" int sc_normalize_pack(const float* x, const double* off, const double* fac, const double* lo,
                                 const double* hi, int f64_path, int B, int C, int H, int W,
                                 void* out_nhwc, int ld_out, int dtype, float* out_nchw, void* stream) {
  if (!x || B <= 0 || C <= 0 || ld_out < C) return SC_ERR_BAD_ARG;
  int64_t HW = (int64_t)H * W, total = HW * B;
  if (!f64_path && C <= 8 && ld_out == 8 && out_nhwc && HW % 4 == 0 && !(reinterpret_cast<uintptr_t>(x) & 15) &&
      !(reinterpret_cast<uintptr_t>(out_nhwc) & 31) && !(reinterpret_cast<uintptr_t>(out_nchw) & 15)) {
    const int64_t tq = total / 4;
    int qb = (int)std::min<int64_t>((tq + 255) / 256, (int64_t)kNumSMs * 16);
    SC_DISPATCH_DTYPE(dtype, (sc::launch_pdl((normalize_pack4_kernel<T>), qb, 256, 0, (cudaStream_t)stream, x, off, fac, lo, hi, C,
                                             HW, tq, (T*)out_nhwc, out_nchw)));
    return check_launch();
  }
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  SC_DISPATCH_DTYPE(dtype, (sc::launch_pdl((normalize_pack_kernel<T>), blocks, 256, 0, (cudaStream_t)stream, 
                               x, off, fac, lo, hi, f64_path, C, HW, total, (T*)out_nhwc, ld_out, out_nchw)));
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// per-channel reductions over NHWC pixels: thread = (pixel lane, 8-channel vector)
// ------------------------------------------------------------------------------------------------
struct RedGeom {
  int CV, CVB, PL, gy;
};
static RedGeom red_geom(int C) {
  RedGeom g;
  g.CV = C / 8;
  g.CVB = g.CV < 256 ? g.CV : 256;
  g.PL = 256 / g.CVB;
  g.gy = (g.CV + g.CVB - 1) / g.CVB;
  return g;
}
static int red_blocks(int64_t P, int PL, int gy) {
  int64_t want = (P + PL * 8 - 1) / (PL * 8);   // >= 8 pixels per thread
  int64_t cap = (kNumSMs * 4) / gy;
  if (cap < 1) cap = 1;
  if (cap > SC_BN_MAX_PARTIALS) cap = SC_BN_MAX_PARTIALS;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

// Block (x, y) reduces its pixel slice for channel group y and writes ONE row of partial sums:
// partials[x][0..C) = sum, partials[x][C..2C) = sum of squares.  No global atomics (thousands of
// fp64 atomics on a few hundred addresses serialise in L2), and the result is deterministic.
template <typename T>
__global__ void bn_stats_kernel(const T* __restrict__ y, int ldy, double* __restrict__ partials, int64_t P,
                                int C, int CVB, int PL) {
  sc::pdl_wait();
  extern __shared__ double sm[];   // [PL][CVB][16]
  int cvl = threadIdx.x % CVB, pl = threadIdx.x / CVB;
  int cv = blockIdx.y * CVB + cvl;
  if (pl < PL && cv * 8 >= C) {
    double* mine = sm + ((size_t)pl * CVB + cvl) * 16;
    for (int i = 0; i < 16; ++i) mine[i] = 0.0;
  }
  if (pl < PL && cv * 8 < C) {
    // fp32 partial sums over runs of 16 pixels, flushed into fp64 accumulators (keeps the fp64 pipe idle)
    double s[8], q[8];
    float fs[8], fq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i] = q[i] = 0.0;
      fs[i] = fq[i] = 0.f;
    }
    // 4 pixels per trip: all four 16-byte loads are issued before any arithmetic (the grid is capped at
    // 296 CTAs by the partial-row protocol, so per-thread memory parallelism is what fills the HBM pipe)
    int run = 0;
    const int64_t stride = (int64_t)gridDim.x * PL;
    int64_t p = (int64_t)blockIdx.x * PL + pl;
    for (; p + 3 * stride < P; p += 4 * stride) {
      f8 v4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v4[u] = load8<T>(y + (p + u * stride) * ldy + cv * 8);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          fs[i] += v4[u].v[i];
          fq[i] = fmaf(v4[u].v[i], v4[u].v[i], fq[i]);
        }
      if ((run += 4) >= 16) {
        run = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += (double)fs[i];
          q[i] += (double)fq[i];
          fs[i] = fq[i] = 0.f;
        }
      }
    }
    for (; p < P; p += stride) {
      f8 v = load8<T>(y + p * ldy + cv * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        fs[i] += v.v[i];
        fq[i] = fmaf(v.v[i], v.v[i], fq[i]);
      }
    }
    // stage this thread's 16 partials: sm[pl][cvl][16] (no shared-memory atomics: with few channels
    // hundreds of threads would contend on a handful of fp64 CAS loops)
    double* mine = sm + ((size_t)pl * CVB + cvl) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mine[i] = s[i] + (double)fs[i];
      mine[8 + i] = q[i] + (double)fq[i];
    }
  }
  __syncthreads();
  double* row = partials + (int64_t)blockIdx.x * 2 * C;
  for (int o = threadIdx.x; o < CVB * 16; o += blockDim.x) {
    const int cvo = o / 16, k = o % 16;          // k < 8: sum, k >= 8: second moment
    const int c = (blockIdx.y * CVB + cvo) * 8 + (k & 7);
    if (c >= C) continue;
    double t = 0.0;
    for (int j = 0; j < PL; ++j) t += sm[((size_t)j * CVB + cvo) * 16 + k];
    row[(k < 8 ? 0 : C) + c] = t;
  }
}

extern "C" int64_t sc_bn_partials_bytes(int C) { return (int64_t)(SC_BN_MAX_PARTIALS + 2) * 2 * C * sizeof(double); }

int bn_stats_flat(const void* y, double* partials, int* nrows_host, int64_t P, int C, int dtype, cudaStream_t st);   // below

extern "C" int sc_bn_stats(const void* y, int ldy, double* partials, int* nrows_host, int64_t P, int C, int dtype,
                           void* stream) {
  if (!y || !partials || !nrows_host || C % 8 || ldy % 8 || P <= 0) return SC_ERR_BAD_ARG;
  if (ldy == C && C <= 256 && P >= (int64_t)1 << 18 && (dtype == SC_F32 || dtype == SC_BF16) &&
      !(reinterpret_cast<uintptr_t>(y) & 15) && !getenv("STARCOP_BN_NOFLAT"))
    return bn_stats_flat(y, partials, nrows_host, P, C, dtype, (cudaStream_t)stream);
  RedGeom g = red_geom(C);
  dim3 grid(red_blocks(P, g.PL, g.gy), g.gy);
  *nrows_host = (int)grid.x;
  size_t smem = (size_t)g.PL * g.CVB * 16 * sizeof(double);   // <= 32 KB
  SC_DISPATCH_DTYPE(dtype, (sc::launch_pdl((bn_stats_kernel<T>), grid, 256, smem, (cudaStream_t)stream, 
                               (const T*)y, ldy, partials, P, C, g.CVB, g.PL)));
  return check_launch();
}

// block = 32 channels x kRedY row slices: the nrows partial rows are summed kRedY-way in parallel
// (<= 10 independent loads per thread: these tiny kernels are pure latency)
constexpr int kRedY = 32;
constexpr int kRedRows = (SC_BN_MAX_PARTIALS + kRedY - 1) / kRedY;     // partial rows per thread (10)
__global__ void bn_finalize_kernel(const double* __restrict__ partials, int nrows, int64_t P, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* running_mean, float* running_var, float momentum, float eps,
                                   int training, float* scale, float* shift, float* save_mean,
                                   float* save_invstd) {
  sc::pdl_wait();
#ifdef SC_PDL_TRIGGER_TINY
  sc::pdl_trigger();
#endif
  __shared__ double sh[2][kRedY][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 32 + tx;
  // the owner thread's parameter loads are issued BEFORE the reduction: their latency overlaps the partial-row loads
  // instead of following the shared-memory combine (this kernel is pure latency, 62 times per step)
  float pf_g = 1.f, pf_b = 0.f, pf_rm = 0.f, pf_rv = 0.f;
  if (ty == 0 && c < C) {
    if (gamma) pf_g = gamma[c];
    if (beta) pf_b = beta[c];
    if (running_mean) {
      pf_rm = running_mean[c];
      pf_rv = running_var[c];
    }
  }
  double s1 = 0.0, s2 = 0.0;
  if (training && c < C) {
    // all of this thread's rows are loaded before the first add: the kernel is one memory latency long, not
    // nrows / 32 of them (it sits between every convolution and its activation, 62 times per step)
    double a[kRedRows], b[kRedRows];
#pragma unroll
    for (int u = 0; u < kRedRows; ++u) {
      const int r = ty + u * kRedY;
      a[u] = r < nrows ? partials[(int64_t)r * 2 * C + c] : 0.0;
      b[u] = r < nrows ? partials[(int64_t)r * 2 * C + C + c] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kRedRows; ++u) {      // same order as a sequential loop over the rows
      s1 += a[u];
      s2 += b[u];
    }
  }
  sh[0][ty][tx] = s1;
  sh[1][ty][tx] = s2;
  __syncthreads();
  if (ty != 0 || c >= C) return;
  float mean, invstd;
  if (training) {
    // fixed summation order: deterministic
#pragma unroll
    for (int j = 1; j < kRedY; ++j) {
      s1 += sh[0][j][tx];
      s2 += sh[1][j][tx];
    }
    // reciprocal multiplies instead of three fp64 divisions on the critical path (the owner thread's chain IS the
    // kernel): 1/P and P/(P-1) are exact-to-rounding constants, rsqrt(double) is accurate to 1 ulp before the cast
    const double invP = 1.0 / (double)P;
    double m = s1 * invP;
    double var = s2 * invP - m * m;        // biased, used for normalisation
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)rsqrt(var + (double)eps);
    if (running_mean) {
      double unb = P > 1 ? var * ((double)P * (1.0 / (double)(P - 1))) : var;
      running_mean[c] = (1.f - momentum) * pf_rm + momentum * mean;
      running_var[c] = (1.f - momentum) * pf_rv + momentum * (float)unb;
    }
  } else {
    mean = pf_rm;
    invstd = 1.f / sqrtf(pf_rv + eps);
  }
  float g = pf_g, b = pf_b;
  float sc_ = g * invstd;
  scale[c] = sc_;
  shift[c] = b - mean * sc_;
  if (save_mean) save_mean[c] = mean;
  if (save_invstd) save_invstd[c] = invstd;
}

extern "C" int sc_bn_finalize(const double* sums, int nrows, int64_t P, int C, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, float momentum, float eps,
                              int training, float* scale, float* shift, float* save_mean,
                              float* save_invstd, void* stream) {
  if (C <= 0 || !scale || !shift || (training && (!sums || nrows > SC_BN_MAX_PARTIALS)) ||
      (!training && (!running_mean || !running_var)))
    return SC_ERR_BAD_ARG;
  sc::launch_pdl((bn_finalize_kernel), (C + 31) / 32, dim3(32, kRedY), 0, (cudaStream_t)stream, 
      sums, nrows, P, C, gamma, beta, running_mean, running_var, momentum, eps, training, scale, shift,
      save_mean, save_invstd);
  return check_launch();
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_kernel(const T* __restrict__ y, int ldy, const float* __restrict__ scale, const float* __restrict__ shift,
              int act, const T* __restrict__ res, int ldr, T* __restrict__ z, int ldz, int64_t P, int C, int H,
              int W, int up2, int CVB, int PL) {
  sc::pdl_wait();
  const int cvl = threadIdx.x % CVB, pl = threadIdx.x / CVB;
  const int cv = blockIdx.y * CVB + cvl;
  if (pl >= PL || cv * 8 >= C) return;
  f8 sc_, sh;
  const bool has_bn = scale != nullptr;
  if (has_bn) {
    sc_ = load8<float>(scale + cv * 8);
    sh = load8<float>(shift + cv * 8);
  }
  for (int64_t p = (int64_t)blockIdx.x * PL + pl; p < P; p += (int64_t)gridDim.x * PL) {
    f8 v = load8<T>(y + p * ldy + cv * 8);
    if (has_bn) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] = apply_act(fmaf(v.v[i], sc_.v[i], sh.v[i]), act);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] = apply_act(v.v[i], act);
    }
    if (res) {
      f8 r = load8<T>(res + p * ldr + cv * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] += r.v[i];
    }
    if (!up2) {
      store8<T>(z + p * ldz + cv * 8, v);
    } else {
      int w = (int)(p % W);
      int64_t t = p / W;
      int h = (int)(t % H);
      int64_t n = t / H;
      int64_t base = ((n * 2 * H + 2 * h) * (2 * (int64_t)W) + 2 * w);
      T* o = z + base * ldz + cv * 8;
      store8<T>(o, v);
      store8<T>(o + ldz, v);
      store8<T>(o + (int64_t)2 * W * ldz, v);
      store8<T>(o + ((int64_t)2 * W + 1) * ldz, v);
    }
  }
}

static int ew_blocks(int64_t total) {
  int64_t b = (total + 255) / 256;
  int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

extern "C" int sc_bn_act(const void* y, int ldy, const float* scale, const float* shift, int act,
                         const void* residual, int ldr, void* z, int ldz, int N, int H, int W, int C,
                         int upsample2, int dtype, void* stream) {
  if (!y || !z || C % 8 || ldy % 8 || ldz % 8 || (residual && ldr % 8)) return SC_ERR_BAD_ARG;
  int64_t P = (int64_t)N * H * W;
  RedGeom g = red_geom(C);
  int64_t want = (P + g.PL * 4 - 1) / (g.PL * 4);
  int64_t cap = (kNumSMs * 8) / g.gy;
  if (cap < 1) cap = 1;
  dim3 grid((unsigned)(want < cap ? (want < 1 ? 1 : want) : cap), g.gy);
  SC_DISPATCH_DTYPE(dtype, (sc::launch_pdl((bn_act_kernel<T>), grid, 256, 0, (cudaStream_t)stream, 
                               (const T*)y, ldy, scale, shift, act, (const T*)residual, ldr, (T*)z, ldz, P, C, H, W,
                               upsample2, g.CVB, g.PL)));
  return check_launch();
}

// dz gather: plain or 2x2 sum-pooled from the (N,2H,2W,ld) buffer of an upsampled consumer
template <typename T>
__device__ __forceinline__ f8 load_dz(const T* dz, int lddz, int pooled, int64_t p, int cv, int H, int W) {
  if (!pooled) return load8<T>(dz + p * lddz + cv * 8);
  int w = (int)(p % W);
  int64_t t = p / W;
  int h = (int)(t % H);
  int64_t n = t / H;
  int64_t base = ((n * 2 * H + 2 * h) * (2 * (int64_t)W) + 2 * w);
  const T* o = dz + base * lddz + cv * 8;
  f8 a = load8<T>(o), b = load8<T>(o + lddz), c = load8<T>(o + (int64_t)2 * W * lddz),
     d = load8<T>(o + ((int64_t)2 * W + 1) * lddz);
#pragma unroll
  for (int i = 0; i < 8; ++i) a.v[i] = (a.v[i] + b.v[i]) + (c.v[i] + d.v[i]);
  return a;
}

template <typename T>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(const T* __restrict__ dz, int lddz, int pooled, const T* __restrict__ y,
                                     int ldy, const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ mean, const float* __restrict__ invstd, int act,
                                     double* __restrict__ partials, int64_t P, int C, int H, int W, int CVB, int PL) {
  sc::pdl_wait();
  extern __shared__ double sm[];   // [PL][CVB][16]
  int cvl = threadIdx.x % CVB, pl = threadIdx.x / CVB;
  int cv = blockIdx.y * CVB + cvl;
  if (pl < PL && cv * 8 >= C) {
    double* mine = sm + ((size_t)pl * CVB + cvl) * 16;
    for (int i = 0; i < 16; ++i) mine[i] = 0.0;
  }
  if (pl < PL && cv * 8 < C) {
    f8 sc_ = load8<float>(scale + cv * 8), sh = load8<float>(shift + cv * 8);
    f8 mu = load8<float>(mean + cv * 8), is = load8<float>(invstd + cv * 8);
    double s[8], q[8];
    float fs[8], fq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i] = q[i] = 0.0;
      fs[i] = fq[i] = 0.f;
    }
    // 2 pixels per trip, loads first (see bn_stats_kernel: the 296-CTA cap makes per-thread memory
    // parallelism the lever; measured 2.8 TB/s with one pixel per trip)
    int run = 0;
    const int64_t stride = (int64_t)gridDim.x * PL;
    int64_t p = (int64_t)blockIdx.x * PL + pl;
    auto accum = [&](const f8& yv, const f8& g) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float gi = g.v[i] * act_mask(fmaf(yv.v[i], sc_.v[i], sh.v[i]), act);
        float xh = (yv.v[i] - mu.v[i]) * is.v[i];
        fs[i] += gi;
        fq[i] = fmaf(gi, xh, fq[i]);
      }
    };
    for (; p + stride < P; p += 2 * stride) {
      f8 y4[2], g4[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) y4[u] = load8<T>(y + (p + u * stride) * ldy + cv * 8);
#pragma unroll
      for (int u = 0; u < 2; ++u) g4[u] = load_dz<T>(dz, lddz, pooled, p + u * stride, cv, H, W);
#pragma unroll
      for (int u = 0; u < 2; ++u) accum(y4[u], g4[u]);
      if ((run += 2) >= 16) {
        run = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += (double)fs[i];
          q[i] += (double)fq[i];
          fs[i] = fq[i] = 0.f;
        }
      }
    }
    for (; p < P; p += stride) {
      f8 yv = load8<T>(y + p * ldy + cv * 8);
      f8 g = load_dz<T>(dz, lddz, pooled, p, cv, H, W);
      accum(yv, g);
    }
    // stage this thread's 16 partials: sm[pl][cvl][16] (no shared-memory atomics: with few channels
    // hundreds of threads would contend on a handful of fp64 CAS loops)
    double* mine = sm + ((size_t)pl * CVB + cvl) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mine[i] = s[i] + (double)fs[i];
      mine[8 + i] = q[i] + (double)fq[i];
    }
  }
  __syncthreads();
  double* row = partials + (int64_t)blockIdx.x * 2 * C;
  for (int o = threadIdx.x; o < CVB * 16; o += blockDim.x) {
    const int cvo = o / 16, k = o % 16;          // k < 8: sum, k >= 8: second moment
    const int c = (blockIdx.y * CVB + cvo) * 8 + (k & 7);
    if (c >= C) continue;
    double t = 0.0;
    for (int j = 0; j < PL; ++j) t += sm[((size_t)j * CVB + cvo) * 16 + k];
    row[(k < 8 ? 0 : C) + c] = t;
  }
}

// Flat variant for the large contiguous layers (C <= 256 a multiple of 8 -- 16 / 32 / 64 of the decoder, 96 / 144 of the
// encoder's expanded tensors; threads beyond PL * CV idle when C / 8 does not divide the block --, dz and y contiguous): both tensors are
// plain arrays, so a CTA streams them through shared memory with 1-D bulk copies (cp.async.bulk, a ring of
// three chunks in flight) instead of holding every in-flight byte in registers -- the register-bound kernel
// above tops out at ~2.8 TB/s on these layers with the 296-CTA cap of the partial-row protocol.
constexpr int kFlatThreads = 256;
constexpr int kFlatStages = 3;
constexpr int kFlatPixPerThread = 4;

template <typename T>
__global__ void __launch_bounds__(kFlatThreads, 2)
bn_bwd_reduce_flat_kernel(const T* __restrict__ dz, const T* __restrict__ y, const float* __restrict__ scale,
                          const float* __restrict__ shift, const float* __restrict__ mean,
                          const float* __restrict__ invstd, int act, double* __restrict__ partials, int64_t P, int C) {
  sc::pdl_wait();
  extern __shared__ __align__(128) uint8_t fsm[];
  const int CV = C / 8, PL = kFlatThreads / CV;
  const int PPC = PL * kFlatPixPerThread;                         // pixels per chunk
  const uint32_t chunk_bytes = (uint32_t)PPC * C * sizeof(T);     // per tensor
  T* bufs = reinterpret_cast<T*>(fsm);                            // [stage][2][PPC*C]
  uint64_t* full = reinterpret_cast<uint64_t*>(fsm + (size_t)kFlatStages * 2 * chunk_bytes);
  double* sm = reinterpret_cast<double*>(fsm);          // [PL][CV][16] reduction scratch, aliases the (drained) ring
  const int tid = threadIdx.x, cv = tid % CV, pl = tid / CV;
  const bool active = pl < PL;                                    // C / 8 need not divide the block size
  const int64_t nchunks = (P + PPC - 1) / PPC;
  const int64_t my_n = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int i = 0; i < kFlatStages; ++i) tc::mbar_init(&full[i], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int64_t i) {
    const int64_t c = blockIdx.x + i * gridDim.x;
    const int64_t p0 = c * PPC;
    const int64_t np = P - p0 < PPC ? P - p0 : PPC;
    const uint32_t bytes = (uint32_t)(np * C * sizeof(T));
    const int slot = (int)(i % kFlatStages);
    T* b = bufs + (size_t)slot * 2 * PPC * C;
    tc::mbar_arrive_expect_tx(&full[slot], 2 * bytes);
    tc::bulk_load_1d(b, y + p0 * C, bytes, &full[slot]);
    tc::bulk_load_1d(b + (size_t)PPC * C, dz + p0 * C, bytes, &full[slot]);
  };
  if (tid == 0)
    for (int i = 0; i < kFlatStages && i < my_n; ++i) issue(i);
  const f8 sc_ = load8<float>(scale + cv * 8), sh = load8<float>(shift + cv * 8);
  const f8 mu = load8<float>(mean + cv * 8), is = load8<float>(invstd + cv * 8);
  double s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.0;
  int slot = 0;
  uint32_t phase = 0;
  for (int64_t i = 0; i < my_n; ++i) {
    const int64_t p0 = (blockIdx.x + i * gridDim.x) * PPC;
    const T* by = bufs + (size_t)slot * 2 * PPC * C;
    const T* bg = by + (size_t)PPC * C;
    tc::mbar_wait(&full[slot], phase);
    float fs[8], fq[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) fs[k] = fq[k] = 0.f;
#pragma unroll
    for (int u = 0; u < kFlatPixPerThread; ++u) {
      const int pp = pl + u * PL;
      if (active && p0 + pp < P) {
        const f8 yv = load8<T>(by + (size_t)pp * C + cv * 8);
        const f8 g = load8<T>(bg + (size_t)pp * C + cv * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float gi = g.v[k] * act_mask(fmaf(yv.v[k], sc_.v[k], sh.v[k]), act);
          const float xh = (yv.v[k] - mu.v[k]) * is.v[k];
          fs[k] += gi;
          fq[k] = fmaf(gi, xh, fq[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s[k] += (double)fs[k];
      q[k] += (double)fq[k];
    }
    __syncthreads();                                  // every reader of the slot is done
    if (tid == 0 && i + kFlatStages < my_n) issue(i + kFlatStages);
    if (++slot == kFlatStages) {
      slot = 0;
      phase ^= 1;
    }
  }
  if (active) {
    double* mine = sm + ((size_t)pl * CV + cv) * 16;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mine[k] = s[k];
      mine[8 + k] = q[k];
    }
  }
  __syncthreads();
  double* row = partials + (int64_t)blockIdx.x * 2 * C;
  for (int o = tid; o < CV * 16; o += kFlatThreads) {
    const int cvo = o / 16, k = o % 16;
    double t = 0.0;
    for (int j = 0; j < PL; ++j) t += sm[((size_t)j * CV + cvo) * 16 + k];
    row[(k < 8 ? 0 : C) + cvo * 8 + (k & 7)] = t;
  }
}

// Forward statistics through the same bulk-copy ring (one input stream): sum and sum of squares per channel of a
// contiguous tensor.  The register kernel (bn_stats_kernel) is capped at 296 CTAs by the partial-row protocol and tops
// out at ~2.2 TB/s on the megapixel layers.
template <typename T>
__global__ void __launch_bounds__(kFlatThreads, 2)
bn_stats_flat_kernel(const T* __restrict__ y, double* __restrict__ partials, int64_t P, int C) {
  sc::pdl_wait();
  extern __shared__ __align__(128) uint8_t fsm[];
  const int CV = C / 8, PL = kFlatThreads / CV;
  const int PPC = PL * 2 * kFlatPixPerThread;                     // pixels per chunk (one tensor: twice the pixels)
  const uint32_t chunk_bytes = (uint32_t)PPC * C * sizeof(T);
  T* bufs = reinterpret_cast<T*>(fsm);                            // [stage][PPC*C]
  uint64_t* full = reinterpret_cast<uint64_t*>(fsm + (size_t)kFlatStages * chunk_bytes);
  double* sm = reinterpret_cast<double*>(fsm);                    // [PL][CV][16] reduction scratch (drained ring)
  const int tid = threadIdx.x, cv = tid % CV, pl = tid / CV;
  const bool active = pl < PL;
  const int64_t nchunks = (P + PPC - 1) / PPC;
  const int64_t my_n = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int i = 0; i < kFlatStages; ++i) tc::mbar_init(&full[i], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int64_t i) {
    const int64_t p0 = (blockIdx.x + i * gridDim.x) * PPC;
    const int64_t np = P - p0 < PPC ? P - p0 : PPC;
    const uint32_t bytes = (uint32_t)(np * C * sizeof(T));
    const int slot = (int)(i % kFlatStages);
    tc::mbar_arrive_expect_tx(&full[slot], bytes);
    tc::bulk_load_1d(bufs + (size_t)slot * PPC * C, y + p0 * C, bytes, &full[slot]);
  };
  if (tid == 0)
    for (int i = 0; i < kFlatStages && i < my_n; ++i) issue(i);
  double s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.0;
  int slot = 0;
  uint32_t phase = 0;
  for (int64_t i = 0; i < my_n; ++i) {
    const int64_t p0 = (blockIdx.x + i * gridDim.x) * PPC;
    const T* by = bufs + (size_t)slot * PPC * C;
    tc::mbar_wait(&full[slot], phase);
    float fs[8], fq[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) fs[k] = fq[k] = 0.f;
#pragma unroll
    for (int u = 0; u < 2 * kFlatPixPerThread; ++u) {
      const int pp = pl + u * PL;
      if (active && p0 + pp < P) {
        const f8 yv = load8<T>(by + (size_t)pp * C + cv * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          fs[k] += yv.v[k];
          fq[k] = fmaf(yv.v[k], yv.v[k], fq[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s[k] += (double)fs[k];
      q[k] += (double)fq[k];
    }
    __syncthreads();                                  // every reader of the slot is done
    if (tid == 0 && i + kFlatStages < my_n) issue(i + kFlatStages);
    if (++slot == kFlatStages) {
      slot = 0;
      phase ^= 1;
    }
  }
  if (active) {
    double* mine = sm + ((size_t)pl * CV + cv) * 16;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mine[k] = s[k];
      mine[8 + k] = q[k];
    }
  }
  __syncthreads();
  double* row = partials + (int64_t)blockIdx.x * 2 * C;
  for (int o = tid; o < CV * 16; o += kFlatThreads) {
    const int cvo = o / 16, k = o % 16;
    double t = 0.0;
    for (int j = 0; j < PL; ++j) t += sm[((size_t)j * CV + cvo) * 16 + k];
    row[(k < 8 ? 0 : C) + cvo * 8 + (k & 7)] = t;
  }
}

// shared by sc_bn_stats (declared above its definition)
template <typename T>
static int launch_bn_stats_flat(const void* y, double* partials, int* nrows_host, int64_t P, int C, cudaStream_t st) {
  const int PPC = (kFlatThreads / (C / 8)) * 2 * kFlatPixPerThread;
  const size_t smem = (size_t)kFlatStages * PPC * C * sizeof(T) + 64;     // >= 32 KB: also the reduction scratch
  const int64_t nchunks = (P + PPC - 1) / PPC;
  const int gx = (int)(nchunks < SC_BN_MAX_PARTIALS ? nchunks : SC_BN_MAX_PARTIALS);
  *nrows_host = gx;
  cudaError_t e = cudaFuncSetAttribute(bn_stats_flat_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    g_last_error = e;
    return SC_ERR_CUDA;
  }
  sc::launch_pdl((bn_stats_flat_kernel<T>), gx, kFlatThreads, smem, st, (const T*)y, partials, P, C);
  return check_launch();
}
int bn_stats_flat(const void* y, double* partials, int* nrows_host, int64_t P, int C, int dtype, cudaStream_t st) {
  if (dtype == SC_F32) return launch_bn_stats_flat<float>(y, partials, nrows_host, P, C, st);
  return launch_bn_stats_flat<__nv_bfloat16>(y, partials, nrows_host, P, C, st);
}

extern "C" int sc_bn_bwd_reduce(const void* dz, int lddz, int pooled, const void* y, int ldy,
                                const float* scale, const float* shift, const float* mean,
                                const float* invstd, int act, double* red, int* nrows_host, int N, int H, int W,
                                int C, int dtype, void* stream) {
  if (!dz || !y || !red || !nrows_host || C % 8 || ldy % 8 || lddz % 8) return SC_ERR_BAD_ARG;
  int64_t P = (int64_t)N * H * W;
  if (!pooled && lddz == C && ldy == C && C <= 256 && P >= (int64_t)1 << 18 &&
      !(reinterpret_cast<uintptr_t>(dz) & 15) && !(reinterpret_cast<uintptr_t>(y) & 15) && !getenv("STARCOP_BN_NOFLAT")) {
    const int esz = dtype == SC_F32 ? 4 : 2;
    const int PPC = (kFlatThreads / (C / 8)) * kFlatPixPerThread;
    const size_t smem = (size_t)kFlatStages * 2 * PPC * C * esz + 64;      // >= 32 KB: also the reduction scratch
    const int64_t nchunks = (P + PPC - 1) / PPC;
    const int gx = (int)(nchunks < SC_BN_MAX_PARTIALS ? nchunks : SC_BN_MAX_PARTIALS);
    *nrows_host = gx;
    cudaError_t e;
    if (dtype == SC_F32) {
      e = cudaFuncSetAttribute(bn_bwd_reduce_flat_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }
      sc::launch_pdl((bn_bwd_reduce_flat_kernel<float>), gx, kFlatThreads, smem, (cudaStream_t)stream, 
          (const float*)dz, (const float*)y, scale, shift, mean, invstd, act, red, P, C);
    } else if (dtype == SC_BF16) {
      e = cudaFuncSetAttribute(bn_bwd_reduce_flat_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }
      sc::launch_pdl((bn_bwd_reduce_flat_kernel<__nv_bfloat16>), gx, kFlatThreads, smem, (cudaStream_t)stream, 
          (const __nv_bfloat16*)dz, (const __nv_bfloat16*)y, scale, shift, mean, invstd, act, red, P, C);
    } else {
      return SC_ERR_BAD_ARG;
    }
    return check_launch();
  }
  RedGeom g = red_geom(C);
  dim3 grid(red_blocks(P, g.PL, g.gy), g.gy);
  *nrows_host = (int)grid.x;
  size_t smem = (size_t)g.PL * g.CVB * 16 * sizeof(double);   // <= 32 KB
  SC_DISPATCH_DTYPE(dtype, (sc::launch_pdl((bn_bwd_reduce_kernel<T>), grid, 256, smem, (cudaStream_t)stream, 
                               (const T*)dz, lddz, pooled, (const T*)y, ldy, scale, shift, mean, invstd, act,
                               red, P, C, H, W, g.CVB, g.PL)));
  return check_launch();
}

// dy = scale*(g - m1 - xhat*m2) with xhat = (y-mean)*invstd  ==  A*g + B*y + D per channel:
//   A = scale, B = -scale*m2*invstd, D = scale*(m2*mean*invstd - m1).
// The totals kernel builds the table [A | shift | B | D] once; the apply kernel keeps one channel
// vector per thread (coefficients in registers) and streams pixels.
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ dz, int lddz, int pooled, const T* __restrict__ y, int ldy,
                    const float* __restrict__ coef, int act, T* __restrict__ dy, int lddy, int64_t P, int C,
                    int H, int W, int CVB, int PL) {
  sc::pdl_wait();
  const int cvl = threadIdx.x % CVB, pl = threadIdx.x / CVB;
  const int cv = blockIdx.y * CVB + cvl;
  if (pl >= PL || cv * 8 >= C) return;
  const f8 cA = load8<float>(coef + cv * 8), cS = load8<float>(coef + C + cv * 8);
  const f8 cB = load8<float>(coef + 2 * C + cv * 8), cD = load8<float>(coef + 3 * C + cv * 8);
  for (int64_t p = (int64_t)blockIdx.x * PL + pl; p < P; p += (int64_t)gridDim.x * PL) {
    f8 yv = load8<T>(y + p * ldy + cv * 8);
    f8 g = load_dz<T>(dz, lddz, pooled, p, cv, H, W);
    f8 o;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float gi = g.v[i] * act_mask(fmaf(yv.v[i], cA.v[i], cS.v[i]), act);
      o.v[i] = fmaf(cA.v[i], gi, fmaf(cB.v[i], yv.v[i], cD.v[i]));
    }
    store8<T>(dy + p * lddy + cv * 8, o);
  }
}

// totals row = sum of the partial rows; the parameter gradients dbeta = sum g, dgamma = sum g*xhat;
// and the apply kernel's coefficient table
__global__ void bn_bwd_totals_kernel(double* __restrict__ partials, int nrows, int C, double invP,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     float* dgamma, float* dbeta, float* __restrict__ coef) {
  sc::pdl_wait();
#ifdef SC_PDL_TRIGGER_TINY
  sc::pdl_trigger();
#endif
  __shared__ double sh[2][kRedY][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 32 + tx;
  // the owner thread's coefficient inputs are fetched before the reduction (latency overlap, see bn_finalize_kernel)
  float pf_sc = 0.f, pf_sh = 0.f, pf_mu = 0.f, pf_is = 0.f, pf_dg = 0.f, pf_db = 0.f;
  if (ty == 0 && c < C) {
    pf_sc = scale[c];
    pf_sh = shift[c];
    pf_mu = mean[c];
    pf_is = invstd[c];
    if (dgamma) {
      pf_dg = dgamma[c];
      pf_db = dbeta[c];
    }
  }
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    double a[kRedRows], b[kRedRows];
#pragma unroll
    for (int u = 0; u < kRedRows; ++u) {
      const int r = ty + u * kRedY;
      a[u] = r < nrows ? partials[(int64_t)r * 2 * C + c] : 0.0;
      b[u] = r < nrows ? partials[(int64_t)r * 2 * C + C + c] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kRedRows; ++u) {
      s1 += a[u];
      s2 += b[u];
    }
  }
  sh[0][ty][tx] = s1;
  sh[1][ty][tx] = s2;
  __syncthreads();
  if (ty != 0 || c >= C) return;
#pragma unroll
  for (int j = 1; j < kRedY; ++j) {
    s1 += sh[0][j][tx];
    s2 += sh[1][j][tx];
  }
  partials[(int64_t)nrows * 2 * C + c] = s1;
  partials[(int64_t)nrows * 2 * C + C + c] = s2;
  if (dgamma) {
    dbeta[c] = pf_db + (float)s1;
    dgamma[c] = pf_dg + (float)s2;
  }
  const float m1 = (float)(s1 * invP), m2 = (float)(s2 * invP);
  const float sc_ = pf_sc, is = pf_is, mu = pf_mu;
  coef[c] = sc_;
  coef[C + c] = pf_sh;
  coef[2 * C + c] = -sc_ * m2 * is;
  coef[3 * C + c] = sc_ * (m2 * mu * is - m1);
}

extern "C" int sc_bn_bwd_apply(const void* dz, int lddz, int pooled, const void* y, int ldy,
                               const float* scale, const float* shift, const float* mean,
                               const float* invstd, const float* gamma, int act, double* partials, int nrows,
                               void* dy, int lddy, float* dgamma, float* dbeta, int N, int H, int W, int C,
                               int dtype, void* stream) {
  if (!dz || !y || !partials || !dy || nrows < 1 || nrows > SC_BN_MAX_PARTIALS || C % 8 || ldy % 8 || lddz % 8 ||
      lddy % 8)
    return SC_ERR_BAD_ARG;
  (void)gamma;
  int64_t P = (int64_t)N * H * W;
  // the coefficient table lives right after the totals row of the partials buffer (4*C floats = 2*C doubles)
  float* coef = reinterpret_cast<float*>(partials + ((int64_t)nrows + 1) * 2 * C);
  sc::launch_pdl((bn_bwd_totals_kernel), (C + 31) / 32, dim3(32, kRedY), 0, (cudaStream_t)stream, 
      partials, nrows, C, 1.0 / (double)P, scale, shift, mean, invstd, dgamma, dbeta, coef);
  RedGeom g = red_geom(C);
  int64_t want = (P + g.PL * 4 - 1) / (g.PL * 4);
  int64_t cap = (kNumSMs * 8) / g.gy;
  if (cap < 1) cap = 1;
  dim3 grid((unsigned)(want < cap ? (want < 1 ? 1 : want) : cap), g.gy);
  SC_DISPATCH_DTYPE(dtype, (sc::launch_pdl((bn_bwd_apply_kernel<T>), grid, 256, 0, (cudaStream_t)stream, 
                               (const T*)dz, lddz, pooled, (const T*)y, ldy, coef, act, (T*)dy, lddy, P, C, H, W,
                               g.CVB, g.PL)));
  return check_launch();
}

template <typename T>
__global__ void add_into_kernel(const T* __restrict__ a, int lda, int pooled, T* __restrict__ out, int ldo,
                                int accumulate, int64_t total, int CV, int H, int W) {
  sc::pdl_wait();
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(idx % CV);
    int64_t p = idx / CV;
    f8 v = load_dz<T>(a, lda, pooled, p, cv, H, W);
    if (accumulate) {
      f8 o = load8<T>(out + p * ldo + cv * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] += o.v[i];
    }
    store8<T>(out + p * ldo + cv * 8, v);
  }
}

extern "C" int sc_add_into(const void* a, int lda, int pooled, void* out, int ldo, int accumulate, int N,
                           int H, int W, int C, int dtype, void* stream) {
  if (!a || !out || C % 8 || lda % 8 || ldo % 8) return SC_ERR_BAD_ARG;
  int CV = C / 8;
  int64_t total = (int64_t)N * H * W * CV;
  SC_DISPATCH_DTYPE(dtype, (sc::launch_pdl((add_into_kernel<T>), ew_blocks(total), 256, 0, (cudaStream_t)stream, 
                               (const T*)a, lda, pooled, (T*)out, ldo, accumulate, total, CV, H, W)));
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// Stem as a space-to-depth convolution.  The network's first layer is Conv2d(C <= 4 -> 32, 3x3, stride 2): on the
// tensor cores that is nine stride-2 TMA boxes of 16-byte pixels per tile (fprop 94 us for 134 MB) and -- worse -- a
// weight gradient whose 166 us are the LAST thing the backward pass does, fully exposed before Adam.  With the input
// regrouped as (N, H/2, W/2, 16) -- channel (sy*2 + sx)*4 + c = input pixel (2Y + sy, 2X + sx), channel c -- the same
// convolution is a stride-1 3x3 convolution over 16 channels whose taps ty, tx in {0, 1} carry the original taps
//      ky -> (ty, sy):  0 -> (0, 1),  1 -> (1, 0),  2 -> (1, 1)      (same for kx), every other weight zero,
// i.e. a layer the halo kernels (one staged patch per tile, fprop and wgrad) already run at full speed.  Zero weights
// contribute exactly zero, so the result differs from the direct form only by summation order.
// ------------------------------------------------------------------------------------------------
__global__ void stem_s2d_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int C, __nv_bfloat16* __restrict__ xs, int H, int W,
                                int64_t total) {
  sc::pdl_wait();
  const int H2 = H / 2, W2 = W / 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % W2);
    const int64_t t = i / W2;
    const int Y = (int)(t % H2);
    const int64_t n = t / H2;
    __align__(16) __nv_bfloat16 o[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat16* src = x + ((n * H + 2 * Y + (q >> 1)) * W + 2 * X + (q & 1)) * ldx;
#pragma unroll
      for (int c = 0; c < 4; ++c) o[q * 4 + c] = c < C ? src[c] : __float2bfloat16_rn(0.f);
    }
    uint4* dst = reinterpret_cast<uint4*>(xs + i * 16);
    dst[0] = reinterpret_cast<const uint4*>(o)[0];
    dst[1] = reinterpret_cast<const uint4*>(o)[1];
  }
}
__device__ __forceinline__ bool stem_s2d_map(int ch, int ty, int tx, int C, int& c, int& ky, int& kx) {
  const int q = ch >> 2, sy = q >> 1, sx = q & 1;
  c = ch & 3;
  // (t, s) -> k:  (0,1) -> 0, (1,0) -> 1, (1,1) -> 2; (0,0) and t = 2 carry nothing
  ky = ty == 0 ? (sy == 1 ? 0 : -1) : (ty == 1 ? 1 + sy : -1);
  kx = tx == 0 ? (sx == 1 ? 0 : -1) : (tx == 1 ? 1 + sx : -1);
  return c < C && ky >= 0 && kx >= 0;
}
// w (Cout, C, 3, 3) fp32 -> bf16 [Cout][9 taps][16 channels] (the tensor-core packing of the equivalent weights)
__global__ void stem_s2d_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wb, int Cout, int C) {
  sc::pdl_wait();
  const int total = Cout * 9 * 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ch = i & 15, tap = (i >> 4) % 9, co = i / 144;
    int c, ky, kx;
    const bool ok = stem_s2d_map(ch, tap / 3, tap % 3, C, c, ky, kx);
    wb[i] = __float2bfloat16_rn(ok ? w[((co * C + c) * 3 + ky) * 3 + kx] : 0.f);
  }
}
// dW (Cout, C, 3, 3) += the non-zero positions of the equivalent weights' gradient g (Cout, 16, 3, 3)
__global__ void stem_s2d_unpack_grad_kernel(const float* __restrict__ g, float* __restrict__ dw, int Cout, int C) {
  sc::pdl_wait();
  const int total = Cout * 16 * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9, ch = (i / 9) & 15, co = i / 144;
    int c, ky, kx;
    if (stem_s2d_map(ch, tap / 3, tap % 3, C, c, ky, kx)) dw[((co * C + c) * 3 + ky) * 3 + kx] += g[i];
  }
}

extern "C" int sc_stem_s2d(const void* x_nhwc, int ldx, int C, void* xs, int N, int H, int W, void* stream) {
  if (!x_nhwc || !xs || C < 1 || C > 4 || ldx < C || (H & 1) || (W & 1) || N <= 0 || (reinterpret_cast<uintptr_t>(xs) & 15))
    return SC_ERR_BAD_ARG;
  const int64_t total = (int64_t)N * (H / 2) * (W / 2);
  sc::launch_pdl((stem_s2d_kernel), ew_blocks(total), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x_nhwc, ldx, C,
                 (__nv_bfloat16*)xs, H, W, total);
  return check_launch();
}
extern "C" int sc_stem_s2d_pack_weights(const float* w_oihw, void* w_bf16, int Cout, int C, void* stream) {
  if (!w_oihw || !w_bf16 || Cout < 1 || C < 1 || C > 4) return SC_ERR_BAD_ARG;
  sc::launch_pdl((stem_s2d_pack_kernel), (Cout * 144 + 255) / 256, 256, 0, (cudaStream_t)stream, w_oihw, (__nv_bfloat16*)w_bf16, Cout, C);
  return check_launch();
}
extern "C" int sc_stem_s2d_unpack_grad(const float* g_s2d, float* dw_oihw, int Cout, int C, void* stream) {
  if (!g_s2d || !dw_oihw || Cout < 1 || C < 1 || C > 4) return SC_ERR_BAD_ARG;
  sc::launch_pdl((stem_s2d_unpack_grad_kernel), (Cout * 144 + 255) / 256, 256, 0, (cudaStream_t)stream, g_s2d, dw_oihw, Cout, C);
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// segmentation head: Conv2d(C->1, 3x3, pad 1, bias).  AI 8.5 FLOP/B -> one thread per pixel.
// ------------------------------------------------------------------------------------------------
constexpr int kHeadMaxC = 64;

template <typename T>
__global__ void head_fprop_kernel(const T* __restrict__ x, int ldx, const float* __restrict__ w,
                                  const float* __restrict__ bias, float* __restrict__ logits, int N, int H,
                                  int W, int C) {
  sc::pdl_wait();
  __shared__ float ws[9 * kHeadMaxC];   // [tap][c]
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
    int tap = i / C, c = i % C;
    ws[i] = w[c * 9 + tap];             // torch (1,C,3,3)
  }
  __syncthreads();
  float b = bias ? bias[0] : 0.f;
  int64_t total = (int64_t)N * H * W;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total;
       p += (int64_t)gridDim.x * blockDim.x) {
    int wq = (int)(p % W);
    int64_t t = p / W;
    int h = (int)(t % H);
    float acc = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int ih = h + kh - 1;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int iw = wq + kw - 1;
        if (iw < 0 || iw >= W) continue;
        const T* xp = x + (p + (int64_t)(kh - 1) * W + (kw - 1)) * ldx;
        const float* wp = ws + (kh * 3 + kw) * C;
        for (int c = 0; c < C; c += 8) {
          f8 v = load8<T>(xp + c);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc = fmaf(v.v[i], wp[c + i], acc);
        }
      }
    }
    logits[p] = acc + b;
  }
}

// Tiled variant for the network's head (16 bf16 channels): a CTA owns 8 x 32 output pixels, stages the 10 x 34 halo
// patch ONCE in shared memory (cp.async, zero fill = the padding, double buffered against the next tile) as two planes
// of 16-byte channel vectors, and a LANE PAIR computes a pixel: each lane keeps the 72 weights of its 8 channels in
// registers, reads nine conflict-free 16-byte vectors, and the pair is combined by one shuffle.  The pixel-per-thread
// kernel above fetched every input byte nine times through L1 and its 144 weights from shared memory per pixel
// (160 us for 201 MB of traffic).
constexpr int kHeadStages = 4;
constexpr int kHtW = 32, kHtH = 8, kHpW = kHtW + 2, kHpH = kHtH + 2, kHpPix = kHpW * kHpH;      // 340 patch pixels
__device__ __forceinline__ void cp_async_zfill(void* dst, const void* src, int bytes, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  if (bytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(ok ? 16u : 0u) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(ok ? 4u : 0u) : "memory");
}
__device__ __forceinline__ void head_tile_coords(int tile, int tiles_w, int tiles_h, int& n, int& h0, int& w0) {
  const int tw = tile % tiles_w;
  const int t2 = tile / tiles_w;
  n = t2 / tiles_h;
  h0 = (t2 - n * tiles_h) * kHtH;
  w0 = tw * kHtW;
}

__global__ void __launch_bounds__(256, 2)
head_fprop_tile_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ bias,
                       float* __restrict__ logits, int H, int W, int tiles_w, int tiles_h, int total_tiles) {
  sc::pdl_wait();
  __shared__ uint4 sx[kHeadStages][2][kHpPix];        // [buffer][channel half][patch pixel]
  const int half = threadIdx.x & 1, slot = threadIdx.x >> 1;
  float wr[9][8];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int i = 0; i < 8; ++i) wr[tap][i] = w[(half * 8 + i) * 9 + tap];            // torch (1,16,3,3)
  const float b = bias ? bias[0] : 0.f;
  auto stage = [&](int tile, int buf) {
    int n, h0, w0;
    head_tile_coords(tile, tiles_w, tiles_h, n, h0, w0);
    for (int g = threadIdx.x; g < 2 * kHpPix; g += 256) {
      const int px = g >> 1, hf = g & 1;
      const int ph = px / kHpW, pw = px - ph * kHpW;
      const int ih = h0 - 1 + ph, iw = w0 - 1 + pw;
      const bool ok = (unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W;
      const __nv_bfloat16* src = ok ? x + (((int64_t)n * H + ih) * W + iw) * ldx + hf * 8 : x;
      cp_async_zfill(&sx[buf][hf][px], src, 16, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // kHeadStages - 1 patches in flight per CTA (one tile of lookahead left the loads latency-bound: 2 CTAs x 11 KB per
  // SM in flight is ~2.7 TB/s); an empty group is committed past the end so that wait_group counts stay uniform
  int buf = 0;
#pragma unroll
  for (int j = 0; j < kHeadStages - 1; ++j) {
    const int t = blockIdx.x + j * gridDim.x;
    if (t < total_tiles) stage(t, j);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, buf = buf + 1 == kHeadStages ? 0 : buf + 1) {
    {
      const int t = tile + (kHeadStages - 1) * gridDim.x;
      const int nb = buf == 0 ? kHeadStages - 1 : buf - 1;      // the buffer released at the end of the previous iteration
      if (t < total_tiles) stage(t, nb);
      else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group %0;" ::"n"(kHeadStages - 1) : "memory");
    __syncthreads();
    int n, h0, w0;
    head_tile_coords(tile, tiles_w, tiles_h, n, h0, w0);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int pix = it * 128 + slot;
      const int ty = pix / kHtW, tx = pix - ty * kHtW;
      float acc = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const uint4 u = sx[buf][half][(ty + kh) * kHpW + tx + kw];
          const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc = fmaf(__uint_as_float(uu[i] << 16), wr[kh * 3 + kw][2 * i], acc);
            acc = fmaf(__uint_as_float(uu[i] & 0xffff0000u), wr[kh * 3 + kw][2 * i + 1], acc);
          }
        }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      const int h = h0 + ty, wq = w0 + tx;
      if (half == 0 && h < H && wq < W) logits[((int64_t)n * H + h) * W + wq] = acc + b;
    }
    __syncthreads();                                  // this buffer is refilled at the top of the next iteration
  }
}

extern "C" int sc_head_fprop(const void* x, int ldx, const float* w, const float* bias, float* logits, int N,
                             int H, int W, int C, int dtype, void* stream) {
  if (!x || !w || !logits || C % 8 || C > kHeadMaxC || ldx % 8) return SC_ERR_BAD_ARG;
  int64_t total = (int64_t)N * H * W;
  if (dtype == SC_BF16 && C == 16 && !(reinterpret_cast<uintptr_t>(x) & 15) && !getenv("STARCOP_HEAD_NOTILE")) {
    const int tiles_w = (W + kHtW - 1) / kHtW, tiles_h = (H + kHtH - 1) / kHtH;
    const int64_t tiles = (int64_t)N * tiles_w * tiles_h;
    if (tiles <= INT32_MAX) {
      const int grid = (int)std::min<int64_t>(tiles, (int64_t)kNumSMs * 2);
      sc::launch_pdl((head_fprop_tile_kernel), grid, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, ldx, w, bias, logits, H, W,
                     tiles_w, tiles_h, (int)tiles);
      return check_launch();
    }
  }
  SC_DISPATCH_DTYPE(dtype, (sc::launch_pdl((head_fprop_kernel<T>), ew_blocks(total), 256, 0, (cudaStream_t)stream, 
                               (const T*)x, ldx, w, bias, logits, N, H, W, C)));
  return check_launch();
}

template <typename T>
__global__ void head_dgrad_kernel(const float* __restrict__ w, const float* __restrict__ dl, T* __restrict__ dx,
                                  int lddx, int N, int H, int W, int C) {
  sc::pdl_wait();
  __shared__ float ws[9 * kHeadMaxC];
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
    int tap = i / C, c = i % C;
    ws[i] = w[c * 9 + tap];
  }
  __syncthreads();
  int64_t total = (int64_t)N * H * W;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total;
       p += (int64_t)gridDim.x * blockDim.x) {
    int wq = (int)(p % W);
    int64_t t = p / W;
    int h = (int)(t % H);
    float g[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        // dx[h,w] += dl[h-kh+1, w-kw+1] * w[kh,kw]
        int oh = h - kh + 1, ow = wq - kw + 1;
        g[kh * 3 + kw] = (oh >= 0 && oh < H && ow >= 0 && ow < W) ? dl[p + (int64_t)(1 - kh) * W + (1 - kw)] : 0.f;
      }
    for (int c = 0; c < C; c += 8) {
      f8 o;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a = 0.f;
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) a = fmaf(g[tp], ws[tp * C + c + i], a);
        o.v[i] = a;
      }
      store8<T>(dx + p * lddx + c, o);
    }
  }
}

// the same tiling for the data gradient: the 10 x 34 dlogits patch is staged (4-byte cp.async, zero fill), a lane
// produces the 8 channels of its half of a pixel from nine shared scalars and its 72 register weights, and a warp
// stores 512 contiguous bytes
__global__ void __launch_bounds__(256, 2)
head_dgrad_tile_kernel(const float* __restrict__ w, const float* __restrict__ dl, __nv_bfloat16* __restrict__ dx, int lddx, int H, int W,
                       int tiles_w, int tiles_h, int total_tiles) {
  sc::pdl_wait();
  __shared__ float sd[kHeadStages][kHpPix];
  const int half = threadIdx.x & 1, slot = threadIdx.x >> 1;
  float wr[9][8];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int i = 0; i < 8; ++i) wr[tap][i] = w[(half * 8 + i) * 9 + tap];
  auto stage = [&](int tile, int buf) {
    int n, h0, w0;
    head_tile_coords(tile, tiles_w, tiles_h, n, h0, w0);
    for (int px = threadIdx.x; px < kHpPix; px += 256) {
      const int ph = px / kHpW, pw = px - ph * kHpW;
      const int ih = h0 - 1 + ph, iw = w0 - 1 + pw;
      const bool ok = (unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W;
      cp_async_zfill(&sd[buf][px], ok ? dl + ((int64_t)n * H + ih) * W + iw : dl, 4, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // kHeadStages - 1 patches in flight per CTA (one tile of lookahead left the loads latency-bound: 2 CTAs x 11 KB per
  // SM in flight is ~2.7 TB/s); an empty group is committed past the end so that wait_group counts stay uniform
  int buf = 0;
#pragma unroll
  for (int j = 0; j < kHeadStages - 1; ++j) {
    const int t = blockIdx.x + j * gridDim.x;
    if (t < total_tiles) stage(t, j);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, buf = buf + 1 == kHeadStages ? 0 : buf + 1) {
    {
      const int t = tile + (kHeadStages - 1) * gridDim.x;
      const int nb = buf == 0 ? kHeadStages - 1 : buf - 1;      // the buffer released at the end of the previous iteration
      if (t < total_tiles) stage(t, nb);
      else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group %0;" ::"n"(kHeadStages - 1) : "memory");
    __syncthreads();
    int n, h0, w0;
    head_tile_coords(tile, tiles_w, tiles_h, n, h0, w0);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int pix = it * 128 + slot;
      const int ty = pix / kHtW, tx = pix - ty * kHtW;
      f8 o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          // dx[h,w] += dl[h-kh+1, w-kw+1] * w[kh,kw]; patch pixel (r, c) is image pixel (h0 - 1 + r, w0 - 1 + c)
          const float g = sd[buf][(ty + 2 - kh) * kHpW + tx + 2 - kw];
#pragma unroll
          for (int i = 0; i < 8; ++i) o.v[i] = fmaf(g, wr[kh * 3 + kw][i], o.v[i]);
        }
      const int h = h0 + ty, wq = w0 + tx;
      if (h < H && wq < W) store8<__nv_bfloat16>(dx + (((int64_t)n * H + h) * W + wq) * lddx + half * 8, o);
    }
    __syncthreads();
  }
}

extern "C" int sc_head_bwd(const void* x, int ldx, const float* w, const float* dlogits, void* dx, int lddx,
                           int N, int H, int W, int C, int dtype, void* stream) {
  if (!x || !w || !dlogits || !dx || C % 8 || C > kHeadMaxC || ldx % 8 || lddx % 8) return SC_ERR_BAD_ARG;
  int64_t total = (int64_t)N * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SC_BF16 && C == 16 && !(reinterpret_cast<uintptr_t>(dx) & 15) && !getenv("STARCOP_HEAD_NOTILE")) {
    const int tiles_w = (W + kHtW - 1) / kHtW, tiles_h = (H + kHtH - 1) / kHtH;
    const int64_t tiles = (int64_t)N * tiles_w * tiles_h;
    if (tiles <= INT32_MAX) {
      const int grid = (int)std::min<int64_t>(tiles, (int64_t)kNumSMs * 2);
      sc::launch_pdl((head_dgrad_tile_kernel), grid, 256, 0, st, w, dlogits, (__nv_bfloat16*)dx, lddx, H, W, tiles_w, tiles_h, (int)tiles);
      return check_launch();
    }
  }
  SC_DISPATCH_DTYPE(dtype, (sc::launch_pdl((head_dgrad_kernel<T>), ew_blocks(total), 256, 0, st, w, dlogits, (T*)dx, lddx, N, H, W, C)));
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// A3/A4/A6: fused weighted BCE-with-logits + gradient + decisions + confusion counts
// ------------------------------------------------------------------------------------------------
__global__ void bce_fused_kernel(const float* __restrict__ logits, const float* __restrict__ y,
                                 const float* __restrict__ w, float pw, int64_t HW, float grad_scale,
                                 double* loss_sum, float* __restrict__ grad, long long* cm, long long* pred_count,
                                 long long* cm_sig, long long* pred_count_sig, float* __restrict__ prediction,
                                 float* __restrict__ loss_px, float* __restrict__ loss_px_w,
                                 long long* __restrict__ pred_binary, long long* __restrict__ differences) {
  sc::pdl_wait();
  int b = blockIdx.y;
  const int64_t base = (int64_t)b * HW;
  double lsum = 0.0;
  long long c_val[4] = {0, 0, 0, 0}, c_sig[4] = {0, 0, 0, 0};
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < HW; s += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = base + s;
    float x = logits[i], t = y[i], wt = w ? w[i] : 1.f;
    // ATen: (1-t)*x + (1+(pw-1)*t) * (log1p(exp(-|x|)) + max(-x,0))
    float lw = 1.f + (pw - 1.f) * t;
    float l = (1.f - t) * x + lw * (log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f));
    lsum += (double)(l * wt);
    // sigmoid as ATen computes it: 1/(1+exp(-x))
    float sg = 1.f / (1.f + expf(-x));
    if (grad) {
      float tt = pw * t;
      grad[i] = ((tt + 1.f - t) * sg - tt) * wt * grad_scale;
    }
    int ti = (int)(long long)t;              // y.long()
    int pv = x >= 0.f ? 1 : 0;               // model_module.py:124
    int ps = sg > 0.5f ? 1 : 0;              // model_module.py:204
    if (ti == 0 || ti == 1) {
      c_val[2 * ti + pv]++;
      c_sig[2 * ti + ps]++;
    }
    if (prediction) prediction[i] = sg;
    if (loss_px) loss_px[i] = l;
    if (loss_px_w) loss_px_w[i] = wt * l;
    if (pred_binary) pred_binary[i] = ps;
    if (differences) differences[i] = 2 * ps + (t == 1.f ? 1 : 0);   // model_module.py:268-269
  }
  __shared__ double s_l[8];
  __shared__ long long s_c[8];
  __shared__ unsigned int s_last;
  if (threadIdx.x < 8) s_c[threadIdx.x] = 0;
  __syncthreads();
  lsum = warp_sum(lsum);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    c_val[k] = warp_sum(c_val[k]);
    c_sig[k] = warp_sum(c_sig[k]);
  }
  if ((threadIdx.x & 31) == 0) {
    s_l[threadIdx.x >> 5] = lsum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd((unsigned long long*)&s_c[k], (unsigned long long)c_val[k]);        // integers: exact in any order
      atomicAdd((unsigned long long*)&s_c[4 + k], (unsigned long long)c_sig[k]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (cm)
      for (int k = 0; k < 4; ++k) atomicAdd((unsigned long long*)&cm[k], (unsigned long long)s_c[k]);
    if (cm_sig)
      for (int k = 0; k < 4; ++k) atomicAdd((unsigned long long*)&cm_sig[k], (unsigned long long)s_c[4 + k]);
    if (pred_count) atomicAdd((unsigned long long*)&pred_count[b], (unsigned long long)(s_c[1] + s_c[3]));
    if (pred_count_sig) atomicAdd((unsigned long long*)&pred_count_sig[b], (unsigned long long)(s_c[5] + s_c[7]));
  }
  if (!loss_sum) return;
  // The loss is a floating-point sum: every block publishes its partial, and the LAST block to finish (ticket in
  // loss_sum[1]) adds them up in block order -- the result does not depend on the blocks' arrival order.
  const unsigned int nblocks = gridDim.x * gridDim.y;
  const unsigned int bid = blockIdx.y * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += s_l[k];
    loss_sum[2 + bid] = t;
    __threadfence();
    const unsigned long long tk = atomicAdd(reinterpret_cast<unsigned long long*>(loss_sum + 1), 1ull);
    s_last = (tk == nblocks - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ double s_red[256];
  double t = 0.0;
  for (unsigned int k = threadIdx.x; k < nblocks; k += blockDim.x) t += __ldcg(loss_sum + 2 + k);
  s_red[threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < (int)blockDim.x; ++k) tot += s_red[k];
    loss_sum[0] += tot;
    *reinterpret_cast<unsigned long long*>(loss_sum + 1) = 0ull;     // ticket ready for the next launch
  }
}

static dim3 bce_grid(int B, int64_t HW) {
  int bx = (int)((HW + 1023) / 1024);
  int cap = (kNumSMs * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3(bx, B);
}

extern "C" int64_t sc_bce_loss_words(int B, int64_t HW) {
  if (B <= 0 || HW <= 0) return -1;
  const dim3 g = bce_grid(B, HW);
  return 2 + (int64_t)g.x * g.y;
}

extern "C" int sc_bce_fused(const float* logits, const float* y, const float* w, float pos_weight, int B,
                            int64_t HW, float grad_scale, double* loss_sum, float* grad, int64_t* cm,
                            int64_t* pred_count, int64_t* cm_sig, int64_t* pred_count_sig, float* prediction,
                            float* loss_px, float* loss_px_w, int64_t* pred_binary, int64_t* differences,
                            void* stream) {
  if (!logits || !y || B <= 0 || HW <= 0 || B > 65535) return SC_ERR_BAD_ARG;
  sc::launch_pdl((bce_fused_kernel), bce_grid(B, HW), 256, 0, (cudaStream_t)stream, 
      logits, y, w, pos_weight, HW, grad_scale, loss_sum, grad, (long long*)cm, (long long*)pred_count,
      (long long*)cm_sig, (long long*)pred_count_sig, prediction, loss_px, loss_px_w, (long long*)pred_binary,
      (long long*)differences);
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// A16 Adam (torch.optim.Adam defaults: no weight decay, no amsgrad), one flat arena
// ------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2_sqrt, float gs) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gs;
    float mi = m[i] + (gi - m[i]) * (1.f - b1);          // torch: exp_avg.lerp_(grad, 1-beta1)
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;          // exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
  }
}

extern "C" int sc_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                            float beta2, float eps, int step_host, float grad_scale, void* stream) {
  if (!p || !g || !m || !v || n <= 0 || step_host < 1) return SC_ERR_BAD_ARG;
  float bc1 = 1.f - powf(beta1, (float)step_host);
  float bc2 = 1.f - powf(beta2, (float)step_host);
  adam_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1,
                                                               sqrtf(bc2), grad_scale);
  return check_launch();
}

// device-resident step counter and learning rate: the whole train step can be replayed as a CUDA graph
__global__ void adam_step_inc_kernel(int* step) {
  sc::pdl_wait(); *step += 1; }
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, const float* __restrict__ lr_dev, float b1, float b2,
                                float eps, const int* __restrict__ step_dev, float gs) {
  sc::pdl_wait();
  const double t = (double)*step_dev;
  const float bc1 = (float)(1.0 - pow((double)b1, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
  const float step_size = *lr_dev / bc1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gs;
    float mi = m[i] + (gi - m[i]) * (1.f - b1);
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

extern "C" int sc_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev,
                                float beta1, float beta2, float eps, int* step_dev, float grad_scale, void* stream) {
  if (!p || !g || !m || !v || !lr_dev || !step_dev || n <= 0) return SC_ERR_BAD_ARG;
  sc::launch_pdl((adam_step_inc_kernel), 1, 1, 0, (cudaStream_t)stream, step_dev);
  sc::launch_pdl((adam_dev_kernel), ew_blocks(n), 256, 0, (cudaStream_t)stream, p, g, m, v, n, lr_dev, beta1, beta2, eps, step_dev,
                                                                   grad_scale);
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// weights: OIHW f32 -> [tap][Cin][Cout] f32 (optionally the flipped/transposed data-gradient filter)
// ------------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ o, int Cout, int Cin,
                                    int KK, int flip_t) {
  int64_t total = (int64_t)Cout * Cin * KK;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (!flip_t) {
      // o[tap][ci][co] = w[co][ci][tap]
      int co = (int)(i % Cout);
      int64_t t = i / Cout;
      int ci = (int)(t % Cin);
      int tap = (int)(t / Cin);
      o[i] = w[((int64_t)co * Cin + ci) * KK + tap];
    } else {
      // dgrad filter: input channels = Cout, output channels = Cin: o[tap][co][ci] = w[co][ci][KK-1-tap]
      int ci = (int)(i % Cin);
      int64_t t = i / Cin;
      int co = (int)(t % Cout);
      int tap = (int)(t / Cout);
      o[i] = w[((int64_t)co * Cin + ci) * KK + (KK - 1 - tap)];
    }
  }
}

extern "C" int sc_pack_weights(const float* w_oihw, float* w_packed, int Cout, int Cin, int KH, int KW,
                               int flip_transpose, void* stream) {
  if (!w_oihw || !w_packed) return SC_ERR_BAD_ARG;
  int64_t total = (int64_t)Cout * Cin * KH * KW;
  pack_weights_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(w_oihw, w_packed, Cout, Cin, KH * KW,
                                                                           flip_transpose);
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// A13 weight_mag1c and A14 threshold + binary opening with the 3x3 cross
// ------------------------------------------------------------------------------------------------
__global__ void weight_mag1c_kernel(const float* __restrict__ m, float* __restrict__ o, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = m[i] / 400.f;                 // np.clip(mag1c/400, .1, 1)
    o[i] = v < 0.1f ? 0.1f : (v > 1.f ? 1.f : v);
  }
}
extern "C" int sc_weight_mag1c(const float* mag1c, float* out, int64_t n, void* stream) {
  if (!mag1c || !out || n <= 0) return SC_ERR_BAD_ARG;
  weight_mag1c_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(mag1c, out, n);
  return check_launch();
}

// erosion: AND over the cross, out-of-image neighbours ignored (kornia "geodesic" border)
__global__ void threshold_erode_kernel(const float* __restrict__ pred, float thr, uint8_t* __restrict__ er,
                                       int B, int H, int W) {
  int64_t total = (int64_t)B * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int w = (int)(i % W);
    int h = (int)((i / W) % H);
    bool v = pred[i] > thr;
    if (h > 0) v = v && (pred[i - W] > thr);
    if (h < H - 1) v = v && (pred[i + W] > thr);
    if (w > 0) v = v && (pred[i - 1] > thr);
    if (w < W - 1) v = v && (pred[i + 1] > thr);
    er[i] = v ? 1 : 0;
  }
}
__global__ void dilate_kernel(const uint8_t* __restrict__ er, long long* __restrict__ out, int B, int H, int W) {
  int64_t total = (int64_t)B * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int w = (int)(i % W);
    int h = (int)((i / W) % H);
    bool v = er[i];
    if (h > 0) v = v || er[i - W];
    if (h < H - 1) v = v || er[i + W];
    if (w > 0) v = v || er[i - 1];
    if (w < W - 1) v = v || er[i + 1];
    out[i] = v ? 1 : 0;
  }
}
extern "C" int sc_threshold_opening(const float* pred, float threshold, int64_t* out, uint8_t* scratch, int B,
                                    int H, int W, void* stream) {
  if (!pred || !out || !scratch) return SC_ERR_BAD_ARG;
  int64_t total = (int64_t)B * H * W;
  threshold_erode_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(pred, threshold, scratch, B, H, W);
  dilate_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(scratch, (long long*)out, B, H, W);
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// A7: the precision / recall threshold sweep of run_validation (starcop/validation.py:37-42, 118-125) in ONE pass:
// for K ascending thresholds, hist[t][i] counts the pixels of class t (= y.long()) whose prediction exceeds exactly
// i of the thresholds (pred > thr_k, strict, like the reference); the confusion matrix of threshold k is then
// [[sum_{i<=k} hist[0][i], sum_{i>k} hist[0][i]], [sum_{i<=k} hist[1][i], sum_{i>k} hist[1][i]]].  Integer
// counts: exact.
// ------------------------------------------------------------------------------------------------
constexpr int kSweepMaxK = 64;
__global__ void threshold_sweep_kernel(const float* __restrict__ pred, const float* __restrict__ y,
                                       const float* __restrict__ thr, int K, int64_t n, unsigned long long* __restrict__ hist) {
  __shared__ float s_thr[kSweepMaxK];
  __shared__ unsigned int s_h[2 * (kSweepMaxK + 1)];
  for (int i = threadIdx.x; i < K; i += blockDim.x) s_thr[i] = thr[i];
  for (int i = threadIdx.x; i < 2 * (K + 1); i += blockDim.x) s_h[i] = 0u;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float p = pred[i];
    const int t = (int)(long long)y[i];
    if (t != 0 && t != 1) continue;
    int lo = 0, hi = K;                       // number of thresholds below p: first k with !(p > thr[k])
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (p > s_thr[mid]) lo = mid + 1; else hi = mid;
    }
    atomicAdd(&s_h[t * (K + 1) + lo], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * (K + 1); i += blockDim.x)
    if (s_h[i]) atomicAdd(&hist[i], (unsigned long long)s_h[i]);
}

extern "C" int sc_threshold_sweep(const float* pred, const float* y, const float* thresholds_ascending, int K, int64_t n,
                                  int64_t* hist, void* stream) {
  if (!pred || !y || !thresholds_ascending || !hist || K < 1 || K > kSweepMaxK || n <= 0) return SC_ERR_BAD_ARG;
  int64_t blocks = (n + 4095) / 4096;
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  threshold_sweep_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(pred, y, thresholds_ascending, K, n,
                                                                        (unsigned long long*)hist);
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// Training augmentation on the device (starcop/data/datamodule.py:128-134: kornia RandomRotation(p=.5, degrees=90)
// + RandomHorizontalFlip + RandomVerticalFlip, applied jointly to input / label / loss weight).  The composite of
// the three operators is ONE affine resampling per sample: out[b,c,y,x] = bilinear (or nearest) sample of in[b,c]
// at (sx, sy) = (m0*x + m1*y + m2, m3*x + m4*y + m5), pixel centres at integer coordinates (align_corners=True),
// zero outside the image -- kornia.warp_affine -> F.grid_sample(padding_mode="zeros") semantics.
// ------------------------------------------------------------------------------------------------
__global__ void affine_warp_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ mats,
                                   int C, int H, int W, int nearest, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    int64_t t = i / W;
    const int y = (int)(t % H);
    t /= H;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    const float* m = mats + b * 6;
    const float sx = fmaf(m[0], (float)x, fmaf(m[1], (float)y, m[2]));
    const float sy = fmaf(m[3], (float)x, fmaf(m[4], (float)y, m[5]));
    const float* src = in + ((int64_t)b * C + c) * H * W;
    float v = 0.f;
    if (nearest) {
      const int xi = (int)nearbyintf(sx), yi = (int)nearbyintf(sy);
      if (xi >= 0 && xi < W && yi >= 0 && yi < H) v = src[(int64_t)yi * W + xi];
    } else {
      const float fx = floorf(sx), fy = floorf(sy);
      const int x0 = (int)fx, y0 = (int)fy;
      const float ax = sx - fx, ay = sy - fy;
      auto at = [&](int yy, int xx) { return (xx >= 0 && xx < W && yy >= 0 && yy < H) ? src[(int64_t)yy * W + xx] : 0.f; };
      // ATen grid_sampler_2d bilinear: nw*(1-ax)(1-ay) + ne*ax(1-ay) + sw*(1-ax)ay + se*ax*ay
      v = at(y0, x0) * ((1.f - ax) * (1.f - ay)) + at(y0, x0 + 1) * (ax * (1.f - ay)) + at(y0 + 1, x0) * ((1.f - ax) * ay) +
          at(y0 + 1, x0 + 1) * (ax * ay);
    }
    out[i] = v;
  }
}

extern "C" int sc_affine_warp(const float* in, float* out, const float* mats_dst_to_src, int B, int C, int H, int W,
                              int nearest, void* stream) {
  if (!in || !out || !mats_dst_to_src || B <= 0 || C <= 0 || H <= 0 || W <= 0 || in == out) return SC_ERR_BAD_ARG;
  const int64_t total = (int64_t)B * C * H * W;
  affine_warp_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(in, out, mats_dst_to_src, C, H, W, nearest, total);
  return check_launch();
}
