// Depthwise 3x3 convolutions (pad 1, stride 1|2) of the MobileNetV2 encoder: fprop, dgrad, wgrad.
//
// These are HBM-bound stencils (9 MACs per element moved).  Every kernel here is a persistent CTA
// that walks spatial tiles of one channel block: the input halo tile is staged ONCE into shared
// memory with the producer's BatchNorm + activation already applied (so the x6 expanded tensor of
// an inverted-residual block is never stored normalised, and each element is normalised once, not
// once per tap), then outputs are computed from shared memory in register strips.  All global loads
// happen in the staging loops (many independent 16 B loads in flight per thread); several CTAs per
// SM overlap one CTA's staging with another's arithmetic.
//   * fprop optionally emits the BatchNorm partial-sum rows of its (stored) outputs, which removes
//     the separate statistics pass over y;
//   * stride-1 dgrad is fprop with the mirrored filter; stride-2 dgrad computes 2x2 output quads
//     from a (TH+1) x (TW+1) tile of dy;
//   * wgrad keeps the 9 x 8 tap sums of a thread's channel vector in registers across all its
//     tiles and writes ONE partial row per CTA (deterministic two-level reduction, no atomics).
#include "common.cuh"

using namespace sc;

namespace {

constexpr int kDwThreads = 256;
constexpr int kDwMaxRows = 296;

struct DwPlan {
  int TH, TW;            // output tile (for the stride-2 dgrad: tile in dy space)
  int IH, IW;            // staged input tile
  int CVB, n_cb;         // 8-channel vectors per block, channel blocks
  int tiles_h, tiles_w;
  int64_t n_tiles;       // N * tiles_h * tiles_w
  int grid_x;
};

static int pick_cvb(int CV) {
  for (int d = 8; d >= 1; --d)
    if (CV % d == 0) return d;
  return 1;
}

static DwPlan dw_plan(int N, int Ho, int Wo, int CV, int th, int tw, int ih, int iw, int ctas_per_sm, int max_rows) {
  DwPlan p;
  p.TH = th;
  p.TW = tw;
  p.IH = ih;
  p.IW = iw;
  p.CVB = pick_cvb(CV);
  p.n_cb = CV / p.CVB;
  p.tiles_h = (Ho + th - 1) / th;
  p.tiles_w = (Wo + tw - 1) / tw;
  p.n_tiles = (int64_t)N * p.tiles_h * p.tiles_w;
  int64_t cap = ((int64_t)kNumSMs * ctas_per_sm + p.n_cb - 1) / p.n_cb;
  if (cap < 1) cap = 1;
  if (cap > max_rows) cap = max_rows;
  p.grid_x = (int)(p.n_tiles < cap ? p.n_tiles : cap);
  return p;
}

// Stage act(x*scale+shift) of the halo tile [ih0, ih0+IH) x [iw0, iw0+IW) of image n, channel vectors
// [cv0, cv0+CVB), into sm[(px*CVB + cvl)*8]; out-of-image elements are ZERO (the conv's padding applies
// to the normalised tensor).
template <typename T>
__device__ __forceinline__ void dw_stage(T* __restrict__ sm, const T* __restrict__ x, int ldx, int n, int ih0, int iw0,
                                         int IH, int IW, int H, int W, int cv0, int CVB, const float* s_bn, int act,
                                         int cvl, int pl, int PLn) {
  // s_bn: shared [2][CVB*8] scale | shift of this channel block, or nullptr (identity).  Loaded here so the
  // 16 registers are not live during the arithmetic phase.
  const bool has_bn = s_bn != nullptr;
  f8 sc_, sh;
  if (has_bn) {
    sc_ = load8<float>(s_bn + cvl * 8);
    sh = load8<float>(s_bn + CVB * 8 + cvl * 8);
  }
  const int npx = IH * IW;
  const T* xn = x + (int64_t)n * H * W * ldx + (cv0 + cvl) * 8;
#pragma unroll 4
  for (int px = pl; px < npx; px += PLn) {
    const int r = px / IW, c = px - r * IW;
    const int ih = ih0 + r, iw = iw0 + c;
    f8 v;
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      v = load8<T>(xn + ((int64_t)ih * W + iw) * ldx);
      if (has_bn) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v.v[i] = apply_act(fmaf(v.v[i], sc_.v[i], sh.v[i]), act);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] = 0.f;
    }
    store8<T>(sm + ((size_t)px * CVB + cvl) * 8, v);
  }
}

__device__ __forceinline__ const float* dw_stage_bn(float* s_bn, const float* __restrict__ scale,
                                                    const float* __restrict__ shift, int cv0, int CVB) {
  if (!scale) return nullptr;
  for (int i = threadIdx.x; i < CVB * 8; i += blockDim.x) {
    s_bn[i] = scale[cv0 * 8 + i];
    s_bn[CVB * 8 + i] = shift[cv0 * 8 + i];
  }
  return s_bn;
}

// ---------------------------------------------------------------------------------------------------
// fprop (and stride-1 dgrad with mirror = 1): thread = (strip of TWP outputs, channel vector)
// ---------------------------------------------------------------------------------------------------
template <typename T, int STRIDE, int TWP>
__global__ void __launch_bounds__(kDwThreads, 3)
dw_fprop_kernel(const T* __restrict__ x, int ldx, const float* __restrict__ scale, const float* __restrict__ shift,
                int act, const float* __restrict__ w, int mirror, T* __restrict__ y, int ldy,
                double* __restrict__ stats, int H, int W, int C, int Ho, int Wo, DwPlan g) {
  extern __shared__ __align__(16) uint8_t dw_smem[];
  const int CVB = g.CVB;
  float* ws = reinterpret_cast<float*>(dw_smem);                 // [9][CVB*8] weights, [2][CVB*8] scale | shift
  T* tile = reinterpret_cast<T*>(ws + 11 * CVB * 8);             // [IH*IW][CVB][8]
  const int tid = threadIdx.x;
  const int PLn = kDwThreads / CVB;
  const int cvl = tid % CVB, pl = tid / CVB;
  const bool active = pl < PLn;
  const int cv0 = blockIdx.y * CVB;
  for (int i = tid; i < 9 * CVB * 8; i += kDwThreads) {
    const int tap = i / (CVB * 8), c = i - tap * CVB * 8;
    ws[i] = w[(int64_t)(cv0 * 8 + c) * 9 + (mirror ? 8 - tap : tap)];
  }
  const float* s_bn = dw_stage_bn(ws + 9 * CVB * 8, scale, shift, cv0, CVB);
  float ssum[8], ssq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) ssum[i] = ssq[i] = 0.f;
  constexpr int NI = (TWP - 1) * STRIDE + 3;
  const int strips_w = g.TW / TWP;
  const int items = g.TH * strips_w;
  for (int64_t t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
    const int tw = (int)(t % g.tiles_w);
    const int64_t t2 = t / g.tiles_w;
    const int th = (int)(t2 % g.tiles_h);
    const int n = (int)(t2 / g.tiles_h);
    const int oh0 = th * g.TH, ow0 = tw * g.TW;
    __syncthreads();                               // the previous tile's readers are done (and ws is staged)
    if (active)
      dw_stage<T>(tile, x, ldx, n, oh0 * STRIDE - 1, ow0 * STRIDE - 1, g.IH, g.IW, H, W, cv0, CVB, s_bn, act, cvl, pl,
                  PLn);
    __syncthreads();
    if (!active) continue;
    for (int it = pl; it < items; it += PLn) {
      const int r = it / strips_w, sw = it - r * strips_w;
      const int oh = oh0 + r, ow = ow0 + sw * TWP;
      if (oh >= Ho || ow >= Wo) continue;
      float acc[TWP][8];
#pragma unroll
      for (int j = 0; j < TWP; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const T* row = tile + ((size_t)((r * STRIDE + kh) * g.IW + sw * TWP * STRIDE) * CVB + cvl) * 8;
        f8 wv[3];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) wv[kw] = load8<float>(ws + (kh * 3 + kw) * CVB * 8 + cvl * 8);
        // one input vector live at a time: column c feeds output j through tap kw = c - j*STRIDE
#pragma unroll
        for (int c = 0; c < NI; ++c) {
          const f8 in = load8<T>(row + (size_t)c * CVB * 8);
#pragma unroll
          for (int j = 0; j < TWP; ++j) {
            const int kw = c - j * STRIDE;
            if (kw >= 0 && kw < 3) {
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(in.v[i], wv[kw].v[i], acc[j][i]);
            }
          }
        }
      }
      T* yp = y + (((int64_t)n * Ho + oh) * Wo + ow) * ldy + (cv0 + cvl) * 8;
#pragma unroll
      for (int j = 0; j < TWP; ++j) {
        if (ow + j >= Wo) break;
        f8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = acc[j][i];
        store8<T>(yp + (int64_t)j * ldy, o);
        if (stats) {
          // statistics of the STORED (storage-precision) values, like the separate pass would see them
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float s = to_f<T>(from_f<T>(o.v[i]));
            ssum[i] += s;
            ssq[i] = fmaf(s, s, ssq[i]);
          }
        }
      }
    }
  }
  if (stats) {
    // deterministic block reduction: stage [pl][cvl][16] floats, then 16*CVB threads sum over pl in fixed order
    __syncthreads();
    float* red = reinterpret_cast<float*>(tile);               // >= 256*16*4 = 16 KB guaranteed by the launcher
    if (active) {
      float* mine = red + ((size_t)pl * CVB + cvl) * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mine[i] = ssum[i];
        mine[8 + i] = ssq[i];
      }
    }
    __syncthreads();
    double* row = stats + (int64_t)blockIdx.x * 2 * C;
    for (int o = tid; o < CVB * 16; o += kDwThreads) {
      const int cvo = o / 16, k = o % 16;
      double s = 0.0;
      for (int j = 0; j < PLn; ++j) s += (double)red[((size_t)j * CVB + cvo) * 16 + k];
      row[(k < 8 ? 0 : C) + (cv0 + cvo) * 8 + (k & 7)] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// stride-2 dgrad: thread = (dy pixel (a,b), channel vector) -> the 2x2 quad dx[2a..2a+1][2b..2b+1]
//   dx[2a  ,2b  ] = g00 w11
//   dx[2a  ,2b+1] = g01 w10 + g00 w12
//   dx[2a+1,2b  ] = g10 w01 + g00 w21
//   dx[2a+1,2b+1] = g11 w00 + g10 w02 + g01 w20 + g00 w22          (gXY = dy[a+X][b+Y], zero outside)
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kDwThreads)
dw_dgrad_s2_kernel(const T* __restrict__ dy, int lddy, const float* __restrict__ w, T* __restrict__ dx, int lddx,
                   int H, int W, int C, int Ho, int Wo, DwPlan g) {
  extern __shared__ __align__(16) uint8_t dw_smem[];
  const int CVB = g.CVB;
  float* ws = reinterpret_cast<float*>(dw_smem);
  T* tile = reinterpret_cast<T*>(ws + 9 * CVB * 8);
  const int tid = threadIdx.x;
  const int PLn = kDwThreads / CVB;
  const int cvl = tid % CVB, pl = tid / CVB;
  const bool active = pl < PLn;
  const int cv0 = blockIdx.y * CVB;
  for (int i = tid; i < 9 * CVB * 8; i += kDwThreads) {
    const int tap = i / (CVB * 8), c = i - tap * CVB * 8;
    ws[i] = w[(int64_t)(cv0 * 8 + c) * 9 + tap];
  }
  const int items = g.TH * g.TW;
  for (int64_t t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
    const int tw = (int)(t % g.tiles_w);
    const int64_t t2 = t / g.tiles_w;
    const int th = (int)(t2 % g.tiles_h);
    const int n = (int)(t2 / g.tiles_h);
    const int a0 = th * g.TH, b0 = tw * g.TW;
    __syncthreads();
    if (active) dw_stage<T>(tile, dy, lddy, n, a0, b0, g.IH, g.IW, Ho, Wo, cv0, CVB, nullptr, 0, cvl, pl, PLn);
    __syncthreads();
    if (!active) continue;
    for (int it = pl; it < items; it += PLn) {
      const int r = it / g.TW, c = it - r * g.TW;
      const int a = a0 + r, b = b0 + c;
      if (2 * a >= H || 2 * b >= W) continue;
      const T* tp = tile + ((size_t)(r * g.IW + c) * CVB + cvl) * 8;
      const f8 g00 = load8<T>(tp), g01 = load8<T>(tp + (size_t)CVB * 8);
      const f8 g10 = load8<T>(tp + (size_t)g.IW * CVB * 8), g11 = load8<T>(tp + (size_t)(g.IW + 1) * CVB * 8);
      const float* wc = ws + cvl * 8;
      const int WS = CVB * 8;
      f8 o00, o01, o10, o11;
      {
        const f8 w11 = load8<float>(wc + 4 * WS), w10 = load8<float>(wc + 3 * WS), w12 = load8<float>(wc + 5 * WS);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o00.v[i] = g00.v[i] * w11.v[i];
          o01.v[i] = fmaf(g01.v[i], w10.v[i], g00.v[i] * w12.v[i]);
        }
      }
      {
        const f8 w01 = load8<float>(wc + 1 * WS), w21 = load8<float>(wc + 7 * WS);
#pragma unroll
        for (int i = 0; i < 8; ++i) o10.v[i] = fmaf(g10.v[i], w01.v[i], g00.v[i] * w21.v[i]);
      }
      {
        const f8 w00 = load8<float>(wc), w02 = load8<float>(wc + 2 * WS), w20 = load8<float>(wc + 6 * WS),
                 w22 = load8<float>(wc + 8 * WS);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o11.v[i] = fmaf(g11.v[i], w00.v[i], fmaf(g10.v[i], w02.v[i], fmaf(g01.v[i], w20.v[i], g00.v[i] * w22.v[i])));
      }
      T* op = dx + (((int64_t)n * H + 2 * a) * W + 2 * b) * lddx + (cv0 + cvl) * 8;
      const bool w1 = 2 * b + 1 < W, h1 = 2 * a + 1 < H;
      store8<T>(op, o00);
      if (w1) store8<T>(op + lddx, o01);
      if (h1) store8<T>(op + (int64_t)W * lddx, o10);
      if (h1 && w1) store8<T>(op + ((int64_t)W + 1) * lddx, o11);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// wgrad: dw[c][kh][kw] = sum_p xhat[p*s-1+k][c] * dy[p][c]; thread = (output pixel, channel vector)
// ---------------------------------------------------------------------------------------------------
template <typename T, int STRIDE>
__global__ void __launch_bounds__(kDwThreads, 2)
dw_wgrad_kernel(const T* __restrict__ x, int ldx, const float* __restrict__ scale, const float* __restrict__ shift,
                int act, const T* __restrict__ dy, int lddy, float* __restrict__ partial, int H, int W, int C, int Ho,
                int Wo, DwPlan g) {
  extern __shared__ __align__(16) uint8_t dw_smem[];
  const int CVB = g.CVB;
  float* s_bnbuf = reinterpret_cast<float*>(dw_smem);            // [2][CVB*8] scale | shift
  T* tile = reinterpret_cast<T*>(s_bnbuf + 2 * CVB * 8);         // [IH*IW][CVB][8]
  T* gt = tile + (size_t)g.IH * g.IW * CVB * 8;                  // [TH*TW][CVB][8] staged dy
  const int tid = threadIdx.x;
  const int PLn = kDwThreads / CVB;
  const int cvl = tid % CVB, pl = tid / CVB;
  const bool active = pl < PLn;
  const int cv0 = blockIdx.y * CVB;
  const float* s_bn = dw_stage_bn(s_bnbuf, scale, shift, cv0, CVB);
  float acc[9][8];
#pragma unroll
  for (int tp = 0; tp < 9; ++tp)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[tp][i] = 0.f;
  const int items = g.TH * g.TW;
  for (int64_t t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
    const int tw = (int)(t % g.tiles_w);
    const int64_t t2 = t / g.tiles_w;
    const int th = (int)(t2 % g.tiles_h);
    const int n = (int)(t2 / g.tiles_h);
    const int oh0 = th * g.TH, ow0 = tw * g.TW;
    __syncthreads();
    if (active) {
      dw_stage<T>(tile, x, ldx, n, oh0 * STRIDE - 1, ow0 * STRIDE - 1, g.IH, g.IW, H, W, cv0, CVB, s_bn, act, cvl, pl,
                  PLn);
      dw_stage<T>(gt, dy, lddy, n, oh0, ow0, g.TH, g.TW, Ho, Wo, cv0, CVB, nullptr, 0, cvl, pl, PLn);
    }
    __syncthreads();
    if (!active) continue;
    for (int it = pl; it < items; it += PLn) {
      const int r = it / g.TW, c = it - r * g.TW;
      // out-of-image outputs were staged as zero gradients: no branch needed
      const f8 gv = load8<T>(gt + ((size_t)it * CVB + cvl) * 8);
      const T* tp0 = tile + ((size_t)(r * STRIDE * g.IW + c * STRIDE) * CVB + cvl) * 8;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const f8 v = load8<T>(tp0 + (size_t)(kh * g.IW + kw) * CVB * 8);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[kh * 3 + kw][i] = fmaf(v.v[i], gv.v[i], acc[kh * 3 + kw][i]);
        }
    }
  }
  // deterministic block reduction, one tap at a time through [pl][cvl][8] floats (8 KB)
  float* red = reinterpret_cast<float*>(dw_smem);
  float* row = partial + (int64_t)blockIdx.x * C * 9;
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    __syncthreads();
    if (active) {
      float* mine = red + ((size_t)pl * CVB + cvl) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) mine[i] = acc[tp][i];
    }
    __syncthreads();
    for (int o = tid; o < CVB * 8; o += kDwThreads) {
      float s = 0.f;
      for (int j = 0; j < PLn; ++j) s += red[(size_t)j * CVB * 8 + o];
      row[(cv0 * 8 + o) * 9 + tp] = s;
    }
  }
}

__global__ void dw_wgrad_sum_kernel(const float* __restrict__ partial, int nrows, int n, float* __restrict__ dw) {
  __shared__ float sh[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (i < n)
    for (int r = ty; r < nrows; r += 32) s += partial[(int64_t)r * n + i];
  sh[ty][tx] = s;
  __syncthreads();
  if (ty != 0 || i >= n) return;
#pragma unroll
  for (int j = 1; j < 32; ++j) s += sh[j][tx];
  dw[i] += s;
}

// generic (any size / parity) stride-2 data gradient, one thread per input vector: fallback for odd H or W
template <typename T>
__global__ void __launch_bounds__(256)
dw_dgrad_generic_kernel(const T* __restrict__ dy, int lddy, const float* __restrict__ w, T* __restrict__ dx, int lddx,
                        int N, int H, int W, int C, int stride, int Ho, int Wo) {
  int CV = C / 8;
  int64_t total = (int64_t)N * H * W * CV;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(idx % CV);
    int64_t p = idx / CV;
    int iw = (int)(p % W);
    int64_t t = p / W;
    int ih = (int)(t % H);
    int n = (int)(t / H);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int kh = 0; kh < 3; ++kh) {
      int th = ih + 1 - kh;
      if (th < 0 || th % stride) continue;
      int ho = th / stride;
      if (ho >= Ho) continue;
      for (int kw = 0; kw < 3; ++kw) {
        int tw = iw + 1 - kw;
        if (tw < 0 || tw % stride) continue;
        int wo = tw / stride;
        if (wo >= Wo) continue;
        f8 gq = load8<T>(dy + (((int64_t)n * Ho + ho) * Wo + wo) * lddy + cv * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(gq.v[i], w[(int64_t)(cv * 8 + i) * 9 + kh * 3 + kw], acc[i]);
      }
    }
    f8 o;
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] = acc[i];
    store8<T>(dx + p * lddx + cv * 8, o);
  }
}

template <typename K>
static int set_smem(K kernel, size_t smem) {
  if (smem <= 48 * 1024) return SC_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    g_last_error = e;
    return SC_ERR_CUDA;
  }
  return SC_OK;
}

static int pick_tw(int Wo, int twp) {
  int tw = Wo >= 16 ? 16 : ((Wo + twp - 1) / twp) * twp;
  return tw;
}

template <typename T>
static int launch_fprop(const T* x, int ldx, const float* scale, const float* shift, int act, const float* w,
                        int mirror, T* y, int ldy, double* stats, int* stats_rows_host, int N, int H, int W, int C,
                        int stride, cudaStream_t st) {
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const int twp = stride == 1 ? 4 : 2;
  const int thm = stride == 1 ? 8 : 4;                 // stride 2 stages a (2*TH+1) x (2*TW+1) tile: keep it small
  const int tw = pick_tw(Wo, twp), th = Ho >= thm ? thm : Ho;
  DwPlan g = dw_plan(N, Ho, Wo, C / 8, th, tw, (th - 1) * stride + 3, (tw - 1) * stride + 3, 6,
                     stats ? SC_BN_MAX_PARTIALS : 1 << 30);
  size_t tile_bytes = (size_t)g.IH * g.IW * g.CVB * 8 * sizeof(T);
  if (stats && tile_bytes < (size_t)kDwThreads * 16 * sizeof(float)) tile_bytes = (size_t)kDwThreads * 16 * sizeof(float);
  const size_t smem = (size_t)11 * g.CVB * 8 * sizeof(float) + tile_bytes;
  if (smem > 200 * 1024) return SC_ERR_UNSUPPORTED;
  if (stats_rows_host) *stats_rows_host = g.grid_x;
  dim3 grid(g.grid_x, g.n_cb);
  int rc;
  if (stride == 1) {
    if ((rc = set_smem(dw_fprop_kernel<T, 1, 4>, smem)) != SC_OK) return rc;
    dw_fprop_kernel<T, 1, 4><<<grid, kDwThreads, smem, st>>>(x, ldx, scale, shift, act, w, mirror, y, ldy, stats, H, W, C,
                                                             Ho, Wo, g);
  } else {
    if ((rc = set_smem(dw_fprop_kernel<T, 2, 2>, smem)) != SC_OK) return rc;
    dw_fprop_kernel<T, 2, 2><<<grid, kDwThreads, smem, st>>>(x, ldx, scale, shift, act, w, mirror, y, ldy, stats, H, W, C,
                                                             Ho, Wo, g);
  }
  return check_launch();
}

}  // namespace

extern "C" int sc_dwconv_fprop(const void* x, int ldx, const float* scale, const float* shift, int act,
                               const float* w, void* y, int ldy, double* stats, int* stats_rows_host, int N, int H,
                               int W, int C, int stride, int dtype, void* stream) {
  if (!x || !w || !y || C % 8 || ldx % 8 || ldy % 8 || (stride != 1 && stride != 2) || N <= 0 || H <= 0 || W <= 0 ||
      (stats && !stats_rows_host))
    return SC_ERR_BAD_ARG;
  SC_DISPATCH_DTYPE(dtype, return launch_fprop<T>((const T*)x, ldx, scale, shift, act, w, 0, (T*)y, ldy, stats,
                                                  stats_rows_host, N, H, W, C, stride, (cudaStream_t)stream));
  return SC_OK;
}

extern "C" int sc_dwconv_dgrad(const void* dy, int lddy, const float* w, void* dx, int lddx, int N, int H, int W,
                               int C, int stride, int dtype, void* stream) {
  if (!dy || !w || !dx || C % 8 || lddy % 8 || lddx % 8 || (stride != 1 && stride != 2) || N <= 0) return SC_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (stride == 1) {
    // dx = depthwise correlation of dy with the mirrored filter
    SC_DISPATCH_DTYPE(dtype, return launch_fprop<T>((const T*)dy, lddy, nullptr, nullptr, SC_ACT_NONE, w, 1, (T*)dx,
                                                    lddx, nullptr, nullptr, N, H, W, C, 1, st));
  }
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  if ((H | W) & 1) {
    int64_t total = (int64_t)N * H * W * (C / 8);
    int64_t b = (total + 255) / 256, cap = (int64_t)kNumSMs * 8;
    SC_DISPATCH_DTYPE(dtype, (dw_dgrad_generic_kernel<T><<<(int)(b < cap ? b : cap), 256, 0, st>>>(
                                 (const T*)dy, lddy, w, (T*)dx, lddx, N, H, W, C, stride, Ho, Wo)));
    return check_launch();
  }
  const int tw = Wo >= 16 ? 16 : Wo, th = Ho >= 8 ? 8 : Ho;
  DwPlan g = dw_plan(N, Ho, Wo, C / 8, th, tw, th + 1, tw + 1, 6, 1 << 30);
  dim3 grid(g.grid_x, g.n_cb);
  SC_DISPATCH_DTYPE(dtype, {
    const size_t smem = (size_t)9 * g.CVB * 8 * sizeof(float) + (size_t)g.IH * g.IW * g.CVB * 8 * sizeof(T);
    int rc = set_smem(dw_dgrad_s2_kernel<T>, smem);
    if (rc != SC_OK) return rc;
    dw_dgrad_s2_kernel<T><<<grid, kDwThreads, smem, st>>>((const T*)dy, lddy, w, (T*)dx, lddx, H, W, C, Ho, Wo, g);
  });
  return check_launch();
}

extern "C" int64_t sc_dwconv_wgrad_workspace_bytes(int C) { return (int64_t)kDwMaxRows * C * 9 * sizeof(float); }

extern "C" int sc_dwconv_wgrad(const void* x, int ldx, const float* scale, const float* shift, int act,
                               const void* dy, int lddy, float* dw, float* workspace, int N, int H, int W, int C,
                               int stride, int dtype, void* stream) {
  if (!x || !dy || !dw || !workspace || C % 8 || ldx % 8 || lddy % 8 || (stride != 1 && stride != 2) || N <= 0)
    return SC_ERR_BAD_ARG;
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  cudaStream_t st = (cudaStream_t)stream;
  const int thm = stride == 1 ? 8 : 4;
  const int tw = Wo >= 16 ? 16 : Wo, th = Ho >= thm ? thm : Ho;
  DwPlan g = dw_plan(N, Ho, Wo, C / 8, th, tw, (th - 1) * stride + 3, (tw - 1) * stride + 3, 2, kDwMaxRows);
  dim3 grid(g.grid_x, g.n_cb);
  SC_DISPATCH_DTYPE(dtype, {
    size_t smem = (size_t)2 * g.CVB * 8 * sizeof(float) + ((size_t)g.IH * g.IW + (size_t)g.TH * g.TW) * g.CVB * 8 * sizeof(T);
    if (smem < (size_t)kDwThreads * 8 * sizeof(float)) smem = (size_t)kDwThreads * 8 * sizeof(float);
    if (smem > 110 * 1024) return SC_ERR_UNSUPPORTED;
    int rc;
    if (stride == 1) {
      if ((rc = set_smem(dw_wgrad_kernel<T, 1>, smem)) != SC_OK) return rc;
      dw_wgrad_kernel<T, 1><<<grid, kDwThreads, smem, st>>>((const T*)x, ldx, scale, shift, act, (const T*)dy, lddy,
                                                            workspace, H, W, C, Ho, Wo, g);
    } else {
      if ((rc = set_smem(dw_wgrad_kernel<T, 2>, smem)) != SC_OK) return rc;
      dw_wgrad_kernel<T, 2><<<grid, kDwThreads, smem, st>>>((const T*)x, ldx, scale, shift, act, (const T*)dy, lddy,
                                                            workspace, H, W, C, Ho, Wo, g);
    }
  });
  int rc = check_launch();
  if (rc != SC_OK) return rc;
  dw_wgrad_sum_kernel<<<(C * 9 + 31) / 32, dim3(32, 32), 0, st>>>(workspace, g.grid_x, C * 9, dw);
  return check_launch();
}
