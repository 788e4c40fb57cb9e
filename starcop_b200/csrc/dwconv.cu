// Depthwise 3x3 convolutions (pad 1, stride 1|2) of the MobileNetV2 encoder: fprop, dgrad, wgrad.
//
// These are HBM-bound stencils (9 MACs per element moved).  Every kernel here is a persistent CTA
// that walks spatial tiles of one channel block: the input halo tile is staged ONCE into shared
// memory by TMA (one 4-D box per tile, zero fill outside the image, a ring of 2-3 tiles in flight per CTA
// so the HBM pipe stays full while the CTA computes), normalised IN PLACE with the producer's
// BatchNorm + activation (so the x6 expanded tensor of an inverted-residual block is never stored
// normalised, and each element is normalised once, not once per tap), then outputs are computed from
// shared memory in register strips.  Tile geometry and the channel-block width are compile-time
// constants, so every shared-memory address is base + immediate.
//   * fprop optionally emits the BatchNorm partial-sum rows of its (stored) outputs, which removes
//     the separate statistics pass over y;
//   * stride-1 dgrad is fprop with the mirrored filter; stride-2 dgrad computes 2x2 output quads
//     from a (TH+1) x (TW+1) tile of dy;
//   * wgrad keeps the 9 x 8 tap sums of a thread's channel vector in registers across all its
//     tiles and writes ONE partial row per CTA (deterministic two-level reduction, no atomics).
#include "common.cuh"
#include "tc_common.cuh"

using namespace sc;
using namespace tc;

namespace {

constexpr int kDwThreads = 256;
#ifndef DW_FPROP_MINB
#define DW_FPROP_MINB 3
#endif
#ifndef DW_FPROP_CTAS
#define DW_FPROP_CTAS 3
#endif
constexpr int kDwMaxRows = 296;

// compile-time tile geometry for stride S, CVB 8-channel vectors per block, strips of TWP outputs
template <int S, int CVB, int TWP>
struct DwGeo {
  static constexpr int PLn = kDwThreads / CVB;                       // pixel lanes (threads per channel vector)
  static constexpr int TW = 16;
  static constexpr int STRIPS = TW / TWP;
  static constexpr int TH0 = S == 1 ? 8 : 4;
  static constexpr int TH = (TH0 * STRIPS >= PLn) ? TH0 : 2 * TH0;   // every lane gets >= 1 strip per tile
  static constexpr int IH = (TH - 1) * S + 3, IW = (TW - 1) * S + 3;
  static constexpr int ITEMS = TH * STRIPS;
  static constexpr int TILE_ELEMS = IH * IW * CVB * 8;
};

template <typename T> struct StatAcc { using type = float; };
template <> struct StatAcc<float> { using type = double; };

struct DwTiles {
  int tiles_h, tiles_w;
  int n_tiles;           // N * tiles_h * tiles_w
};

// The staged tile holds RAW values (TMA zero-filled outside the image).  Normalise it in place:
// v <- act(v*scale+shift) inside the image, 0 outside (the conv's padding applies to the NORMALISED
// tensor, and act(shift) != 0).  s_bn: shared [2][CVB*8] scale | shift.
template <typename T, int IH, int IW, int CVB>
__device__ __forceinline__ void dw_normalize(T* __restrict__ sm, int ih0, int iw0, int H, int W, const float* s_bn,
                                             int act, int cvl, int pl) {
  constexpr int PLn = kDwThreads / CVB;
  const f8 sc_ = load8<float>(s_bn + cvl * 8), sh = load8<float>(s_bn + CVB * 8 + cvl * 8);
#pragma unroll 2
  for (int px = pl; px < IH * IW; px += PLn) {
    const int r = px / IW, c = px - r * IW;
    const int ih = ih0 + r, iw = iw0 + c;
    T* p = sm + (px * CVB + cvl) * 8;
    f8 v;
    if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) {
      v = load8<T>(p);
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] = apply_act(fmaf(v.v[i], sc_.v[i], sh.v[i]), act);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] = 0.f;
    }
    store8<T>(p, v);
  }
}

// ring of `ns` staged tiles, one mbarrier each; the CTA's i-th tile lives in slot i % ns
struct DwRing {
  uint8_t* bufs;
  uint64_t* full;
  int ns, stage_bytes;
};
__device__ __forceinline__ DwRing dw_ring_init(uint8_t* smem_aligned, int ns, int stage_bytes) {
  DwRing r;
  r.bufs = smem_aligned;
  r.full = reinterpret_cast<uint64_t*>(smem_aligned + (size_t)ns * stage_bytes);
  r.ns = ns;
  r.stage_bytes = stage_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ns; ++i) mbar_init(&r.full[i], 1);
    fence_barrier_init();
  }
  return r;
}
__device__ __forceinline__ const float* dw_stage_bn(float* s_bn, const float* __restrict__ scale,
                                                    const float* __restrict__ shift, int cv0, int n) {
  if (!scale) return nullptr;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s_bn[i] = scale[cv0 * 8 + i];
    s_bn[n + i] = shift[cv0 * 8 + i];
  }
  return s_bn;
}

__device__ __forceinline__ void dw_tile_coords(int t, const DwTiles& g, int& n, int& th, int& tw) {
  tw = t % g.tiles_w;
  const int t2 = t / g.tiles_w;
  th = t2 % g.tiles_h;
  n = t2 / g.tiles_h;
}

// ---------------------------------------------------------------------------------------------------
// fprop (and stride-1 dgrad with mirror = 1): thread = (strip of TWP outputs, channel vector)
// shared memory: [ns x stage][ns mbarriers][9*CVB*8 weights | 2*CVB*8 scale, shift]
// ---------------------------------------------------------------------------------------------------
template <typename T, int S, int CVB>
__global__ void __launch_bounds__(kDwThreads, DW_FPROP_MINB)
dw_fprop_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ scale,
                const float* __restrict__ shift, int act, const float* __restrict__ w, int mirror, T* __restrict__ y,
                int ldy, double* __restrict__ stats, int H, int W, int C, int Ho, int Wo, DwTiles g, int ns,
                int stage_bytes) {
  sc::pdl_wait();
  constexpr int TWP = S == 1 ? 4 : 2;
  using G = DwGeo<S, CVB, TWP>;
  constexpr int NI = (TWP - 1) * S + 3;
  constexpr uint32_t kTileBytes = G::TILE_ELEMS * sizeof(T);
  extern __shared__ __align__(128) uint8_t dw_smem[];   // no integer casts: keeps the shared address space (LDS, not LD)
  DwRing ring = dw_ring_init(dw_smem, ns, stage_bytes);
  float* ws = reinterpret_cast<float*>(ring.full + ns + (ns & 1));   // 16 B aligned
  const int tid = threadIdx.x;
  const int cvl = tid % CVB, pl = tid / CVB;
  const bool active = pl < G::PLn;
  const int cv0 = blockIdx.y * CVB;
  for (int i = tid; i < 9 * CVB * 8; i += kDwThreads) {
    const int tap = i / (CVB * 8), c = i - tap * CVB * 8;
    ws[i] = w[(cv0 * 8 + c) * 9 + (mirror ? 8 - tap : tap)];
  }
  const float* s_bn = dw_stage_bn(ws + 9 * CVB * 8, scale, shift, cv0, CVB * 8);
  // per-thread statistics: fp32 in the bf16 throughput mode, fp64 in the fp32 parity mode
  using StatT = typename StatAcc<T>::type;
  StatT ssum[8], ssq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) ssum[i] = ssq[i] = (StatT)0;
  const float* wsc = ws + cvl * 8;
  const int my_tiles = (g.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int i) {     // thread 0: TMA the CTA's i-th tile into slot i % ns
    int n, th, tw;
    dw_tile_coords(blockIdx.x + i * gridDim.x, g, n, th, tw);
    const int slot = i % ns;
    mbar_arrive_expect_tx(&ring.full[slot], kTileBytes);
    tma_load_4d(ring.bufs + (size_t)slot * stage_bytes, &tmX, cv0 * 8, tw * G::TW * S - 1, th * G::TH * S - 1, n,
                &ring.full[slot]);
  };
  __syncthreads();                                 // barriers initialised, weights staged
  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < ns && i < my_tiles; ++i) issue(i);
  }
  int slot = 0;
  uint32_t phase = 0;
  for (int i = 0; i < my_tiles; ++i) {
    int n, th, tw;
    dw_tile_coords(blockIdx.x + i * gridDim.x, g, n, th, tw);
    const int oh0 = th * G::TH, ow0 = tw * G::TW;
    T* tile = reinterpret_cast<T*>(ring.bufs + (size_t)slot * stage_bytes);
    mbar_wait(&ring.full[slot], phase);
    if (s_bn) {
      if (active) dw_normalize<T, G::IH, G::IW, CVB>(tile, oh0 * S - 1, ow0 * S - 1, H, W, s_bn, act, cvl, pl);
      __syncthreads();
    }
    if (active) {
#pragma unroll 1
      for (int it = pl; it < G::ITEMS; it += G::PLn) {
        const int r = it / G::STRIPS, sw = it - r * G::STRIPS;
        const int oh = oh0 + r, ow = ow0 + sw * TWP;
        if (oh >= Ho || ow >= Wo) continue;
        float acc[TWP][8];
#pragma unroll
        for (int j = 0; j < TWP; ++j)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[j][k] = 0.f;
        const T* tp0 = tile + ((r * S * G::IW + sw * TWP * S) * CVB + cvl) * 8;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          f8 wv[3];
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) wv[kw] = load8<float>(wsc + (kh * 3 + kw) * CVB * 8);
          // one input vector live at a time: column c feeds output j through tap kw = c - j*S
#pragma unroll
          for (int c = 0; c < NI; ++c) {
            const f8 in = load8<T>(tp0 + (kh * G::IW + c) * CVB * 8);
#pragma unroll
            for (int j = 0; j < TWP; ++j) {
              const int kw = c - j * S;
              if (kw >= 0 && kw < 3) {
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[j][k] = fmaf(in.v[k], wv[kw].v[k], acc[j][k]);
              }
            }
          }
        }
        T* yp = y + ((int64_t)(n * Ho + oh) * Wo + ow) * ldy + (cv0 + cvl) * 8;
#pragma unroll
        for (int j = 0; j < TWP; ++j) {
          if (ow + j >= Wo) break;
          f8 o;
#pragma unroll
          for (int k = 0; k < 8; ++k) o.v[k] = acc[j][k];
          store8<T>(yp + (int64_t)j * ldy, o);
          if (stats) {
            // statistics of the STORED (storage-precision) values, like the separate pass would see them
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const StatT sv = (StatT)to_f<T>(from_f<T>(o.v[k]));
              ssum[k] += sv;
              ssq[k] += sv * sv;
            }
          }
        }
      }
    }
    __syncthreads();                               // every reader of this slot is done
    if (tid == 0 && i + ns < my_tiles) {
      fence_proxy_async();                         // generic-proxy accesses to the slot precede the async refill
      issue(i + ns);
    }
    if (++slot == ns) {
      slot = 0;
      phase ^= 1;
    }
  }
  if (stats) {
    // deterministic block reduction: stage [pl][cvl][16] floats, then 16*CVB threads sum over pl in fixed order
    StatT* red = reinterpret_cast<StatT*>(ring.bufs);          // >= 256*16*8 = 32 KB guaranteed by the launcher
    if (active) {
      StatT* mine = red + (pl * CVB + cvl) * 16;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        mine[k] = ssum[k];
        mine[8 + k] = ssq[k];
      }
    }
    __syncthreads();
    double* row = stats + (int64_t)blockIdx.x * 2 * C;
    for (int o = tid; o < CVB * 16; o += kDwThreads) {
      const int cvo = o / 16, k = o % 16;
      double sd = 0.0;
      for (int j = 0; j < G::PLn; ++j) sd += (double)red[(j * CVB + cvo) * 16 + k];
      row[(k < 8 ? 0 : C) + (cv0 + cvo) * 8 + (k & 7)] = sd;
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
template <int CVB>
struct DwGeoD2 {
  static constexpr int PLn = kDwThreads / CVB;
  static constexpr int TW = 16;
  static constexpr int TH = (8 * TW >= PLn * 2) ? 8 : 16;      // >= 2 quads per lane
  static constexpr int IH = TH + 1, IW = TW + 1;
  static constexpr int TILE_ELEMS = IH * IW * CVB * 8;
};

template <typename T, int CVB>
__global__ void __launch_bounds__(kDwThreads, 3)
dw_dgrad_s2_kernel(const __grid_constant__ CUtensorMap tmDY, const float* __restrict__ w, T* __restrict__ dx, int lddx,
                   int H, int W, int C, DwTiles g, int ns, int stage_bytes) {
  sc::pdl_wait();
  using G = DwGeoD2<CVB>;
  constexpr int IW = G::IW;
  constexpr uint32_t kTileBytes = G::TILE_ELEMS * sizeof(T);
  extern __shared__ __align__(128) uint8_t dw_smem[];   // no integer casts: keeps the shared address space (LDS, not LD)
  DwRing ring = dw_ring_init(dw_smem, ns, stage_bytes);
  float* ws = reinterpret_cast<float*>(ring.full + ns + (ns & 1));
  const int tid = threadIdx.x;
  const int cvl = tid % CVB, pl = tid / CVB;
  const bool active = pl < G::PLn;
  const int cv0 = blockIdx.y * CVB;
  for (int i = tid; i < 9 * CVB * 8; i += kDwThreads) {
    const int tap = i / (CVB * 8), c = i - tap * CVB * 8;
    ws[i] = w[(cv0 * 8 + c) * 9 + tap];
  }
  const float* wc = ws + cvl * 8;
  constexpr int WS = CVB * 8;
  const int my_tiles = (g.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int i) {
    int n, th, tw;
    dw_tile_coords(blockIdx.x + i * gridDim.x, g, n, th, tw);
    const int slot = i % ns;
    mbar_arrive_expect_tx(&ring.full[slot], kTileBytes);
    tma_load_4d(ring.bufs + (size_t)slot * stage_bytes, &tmDY, cv0 * 8, tw * G::TW, th * G::TH, n, &ring.full[slot]);
  };
  __syncthreads();
  if (tid == 0) {
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < ns && i < my_tiles; ++i) issue(i);
  }
  int slot = 0;
  uint32_t phase = 0;
  for (int i = 0; i < my_tiles; ++i) {
    int n, th, tw;
    dw_tile_coords(blockIdx.x + i * gridDim.x, g, n, th, tw);
    const int a0 = th * G::TH, b0 = tw * G::TW;
    const T* tile = reinterpret_cast<const T*>(ring.bufs + (size_t)slot * stage_bytes);
    mbar_wait(&ring.full[slot], phase);
    if (active) {
#pragma unroll 1
      for (int it = pl; it < G::TH * G::TW; it += G::PLn) {
        const int r = it / G::TW, c = it - r * G::TW;
        const int a = a0 + r, b = b0 + c;
        if (2 * a >= H || 2 * b >= W) continue;
        const T* tp = tile + ((r * IW + c) * CVB + cvl) * 8;
        const f8 g00 = load8<T>(tp), g01 = load8<T>(tp + CVB * 8);
        const f8 g10 = load8<T>(tp + IW * CVB * 8), g11 = load8<T>(tp + (IW + 1) * CVB * 8);
        f8 o00, o01, o10, o11;
        {
          const f8 w11 = load8<float>(wc + 4 * WS), w10 = load8<float>(wc + 3 * WS), w12 = load8<float>(wc + 5 * WS);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            o00.v[k] = g00.v[k] * w11.v[k];
            o01.v[k] = fmaf(g01.v[k], w10.v[k], g00.v[k] * w12.v[k]);
          }
        }
        {
          const f8 w01 = load8<float>(wc + 1 * WS), w21 = load8<float>(wc + 7 * WS);
#pragma unroll
          for (int k = 0; k < 8; ++k) o10.v[k] = fmaf(g10.v[k], w01.v[k], g00.v[k] * w21.v[k]);
        }
        {
          const f8 w00 = load8<float>(wc), w02 = load8<float>(wc + 2 * WS), w20 = load8<float>(wc + 6 * WS),
                   w22 = load8<float>(wc + 8 * WS);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            o11.v[k] = fmaf(g11.v[k], w00.v[k], fmaf(g10.v[k], w02.v[k], fmaf(g01.v[k], w20.v[k], g00.v[k] * w22.v[k])));
        }
        T* op = dx + ((int64_t)(n * H + 2 * a) * W + 2 * b) * lddx + (cv0 + cvl) * 8;
        const bool w1 = 2 * b + 1 < W, h1 = 2 * a + 1 < H;
        store8<T>(op, o00);
        if (w1) store8<T>(op + lddx, o01);
        if (h1) store8<T>(op + (int64_t)W * lddx, o10);
        if (h1 && w1) store8<T>(op + ((int64_t)W + 1) * lddx, o11);
      }
    }
    __syncthreads();
    if (tid == 0 && i + ns < my_tiles) {
      fence_proxy_async();
      issue(i + ns);
    }
    if (++slot == ns) {
      slot = 0;
      phase ^= 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// wgrad: dw[c][kh][kw] = sum_p xhat[p*s-1+k][c] * dy[p][c]; thread = (strip of 2 outputs, channel vector)
// one stage = [x halo tile][dy tile] (two TMA boxes on one mbarrier)
// ---------------------------------------------------------------------------------------------------
template <typename T, int S, int CVB>
__global__ void __launch_bounds__(kDwThreads, 2)
dw_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                const float* __restrict__ scale, const float* __restrict__ shift, int act, float* __restrict__ partial,
                int H, int W, int C, DwTiles g, int ns, int stage_bytes) {
  sc::pdl_wait();
  constexpr int TWP = 2;
  using G = DwGeo<S, CVB, TWP>;
  constexpr int NI = (TWP - 1) * S + 3;
  constexpr uint32_t kXBytes = G::TILE_ELEMS * sizeof(T);
  constexpr uint32_t kXBytesAl = (kXBytes + 127) & ~127u;
  constexpr uint32_t kGBytes = G::TH * G::TW * CVB * 8 * sizeof(T);
  extern __shared__ __align__(128) uint8_t dw_smem[];   // no integer casts: keeps the shared address space (LDS, not LD)
  DwRing ring = dw_ring_init(dw_smem, ns, stage_bytes);
  float* s_bnbuf = reinterpret_cast<float*>(ring.full + ns + (ns & 1));
  const int tid = threadIdx.x;
  const int cvl = tid % CVB, pl = tid / CVB;
  const bool active = pl < G::PLn;
  const int cv0 = blockIdx.y * CVB;
  const float* s_bn = dw_stage_bn(s_bnbuf, scale, shift, cv0, CVB * 8);
  float acc[9][8];
#pragma unroll
  for (int tp = 0; tp < 9; ++tp)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[tp][k] = 0.f;
  const int my_tiles = (g.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int i) {
    int n, th, tw;
    dw_tile_coords(blockIdx.x + i * gridDim.x, g, n, th, tw);
    const int slot = i % ns;
    uint8_t* buf = ring.bufs + (size_t)slot * stage_bytes;
    mbar_arrive_expect_tx(&ring.full[slot], kXBytes + kGBytes);
    tma_load_4d(buf, &tmX, cv0 * 8, tw * G::TW * S - 1, th * G::TH * S - 1, n, &ring.full[slot]);
    // out-of-image outputs arrive as zero gradients: no bounds checks in the arithmetic
    tma_load_4d(buf + kXBytesAl, &tmDY, cv0 * 8, tw * G::TW, th * G::TH, n, &ring.full[slot]);
  };
  __syncthreads();
  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < ns && i < my_tiles; ++i) issue(i);
  }
  int slot = 0;
  uint32_t phase = 0;
  for (int i = 0; i < my_tiles; ++i) {
    int n, th, tw;
    dw_tile_coords(blockIdx.x + i * gridDim.x, g, n, th, tw);
    T* tile = reinterpret_cast<T*>(ring.bufs + (size_t)slot * stage_bytes);
    const T* gt = reinterpret_cast<const T*>(ring.bufs + (size_t)slot * stage_bytes + kXBytesAl);
    mbar_wait(&ring.full[slot], phase);
    if (s_bn) {
      if (active)
        dw_normalize<T, G::IH, G::IW, CVB>(tile, th * G::TH * S - 1, tw * G::TW * S - 1, H, W, s_bn, act, cvl, pl);
      __syncthreads();
    }
    if (active) {
#pragma unroll 1
      for (int it = pl; it < G::ITEMS; it += G::PLn) {
        const int r = it / G::STRIPS, sw = it - r * G::STRIPS;
        f8 gv[TWP];
#pragma unroll
        for (int j = 0; j < TWP; ++j) gv[j] = load8<T>(gt + ((r * G::TW + sw * TWP + j) * CVB + cvl) * 8);
        const T* tp0 = tile + ((r * S * G::IW + sw * TWP * S) * CVB + cvl) * 8;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int c = 0; c < NI; ++c) {
            const f8 v = load8<T>(tp0 + (kh * G::IW + c) * CVB * 8);
#pragma unroll
            for (int j = 0; j < TWP; ++j) {
              const int kw = c - j * S;
              if (kw >= 0 && kw < 3) {
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[kh * 3 + kw][k] = fmaf(v.v[k], gv[j].v[k], acc[kh * 3 + kw][k]);
              }
            }
          }
      }
    }
    __syncthreads();
    if (tid == 0 && i + ns < my_tiles) {
      fence_proxy_async();
      issue(i + ns);
    }
    if (++slot == ns) {
      slot = 0;
      phase ^= 1;
    }
  }
  // deterministic block reduction, one tap at a time through [pl][cvl][8] floats (8 KB)
  float* red = reinterpret_cast<float*>(ring.bufs);
  float* row = partial + (int64_t)blockIdx.x * C * 9;
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    __syncthreads();
    if (active) {
      float* mine = red + (pl * CVB + cvl) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) mine[k] = acc[tp][k];
    }
    __syncthreads();
    for (int o = tid; o < CVB * 8; o += kDwThreads) {
      float sv = 0.f;
      for (int j = 0; j < G::PLn; ++j) sv += red[j * CVB * 8 + o];
      row[(cv0 * 8 + o) * 9 + tp] = sv;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// segmentation-head wgrad (Conv2d(C -> 1, 3x3), C <= 64): dw[c][tap] = sum_p x[p+tap][c] * dl[p], dbias = sum dl.
// Same machinery as the depthwise wgrad (TMA halo-tile ring, 72 tap sums per thread, one partial row per CTA);
// the "dy" of every channel is the single-channel fp32 logit gradient, read straight from global memory.
// ---------------------------------------------------------------------------------------------------
template <typename T, int CVB>
__global__ void __launch_bounds__(kDwThreads, 2)
head_wgrad_tile_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ dl, float* __restrict__ partial,
                       int H, int W, int C, DwTiles g, int ns, int stage_bytes) {
  sc::pdl_wait();
  constexpr int TWP = 2;
  using G = DwGeo<1, CVB, TWP>;
  constexpr int NI = TWP + 2;
  constexpr uint32_t kXBytes = G::TILE_ELEMS * sizeof(T);
  extern __shared__ __align__(128) uint8_t dw_smem[];
  DwRing ring = dw_ring_init(dw_smem, ns, stage_bytes);
  const int tid = threadIdx.x;
  const int cvl = tid % CVB, pl = tid / CVB;
  const bool active = pl < G::PLn;
  float acc[9][8];
#pragma unroll
  for (int tp = 0; tp < 9; ++tp)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[tp][k] = 0.f;
  float bsum = 0.f;
  const int my_tiles = (g.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int i) {
    int n, th, tw;
    dw_tile_coords(blockIdx.x + i * gridDim.x, g, n, th, tw);
    const int slot = i % ns;
    mbar_arrive_expect_tx(&ring.full[slot], kXBytes);
    tma_load_4d(ring.bufs + (size_t)slot * stage_bytes, &tmX, 0, tw * G::TW - 1, th * G::TH - 1, n, &ring.full[slot]);
  };
  __syncthreads();
  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < ns && i < my_tiles; ++i) issue(i);
  }
  int slot = 0;
  uint32_t phase = 0;
  for (int i = 0; i < my_tiles; ++i) {
    int n, th, tw;
    dw_tile_coords(blockIdx.x + i * gridDim.x, g, n, th, tw);
    const T* tile = reinterpret_cast<const T*>(ring.bufs + (size_t)slot * stage_bytes);
    mbar_wait(&ring.full[slot], phase);
    if (active) {
#pragma unroll 1
      for (int it = pl; it < G::ITEMS; it += G::PLn) {
        const int r = it / G::STRIPS, sw = it - r * G::STRIPS;
        const int oh = th * G::TH + r, ow = tw * G::TW + sw * TWP;
        float g2[TWP] = {0.f, 0.f};                   // this strip's logit gradients, zero outside the image
        if (oh < H) {
          const float* dp = dl + ((int64_t)n * H + oh) * W + ow;
#pragma unroll
          for (int j = 0; j < TWP; ++j)
            if (ow + j < W) g2[j] = dp[j];
        }
        if (cvl == 0) bsum += g2[0] + g2[1];
        const T* tp0 = tile + ((r * G::IW + sw * TWP) * CVB + cvl) * 8;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int c = 0; c < NI; ++c) {
            const f8 v = load8<T>(tp0 + (kh * G::IW + c) * CVB * 8);
#pragma unroll
            for (int j = 0; j < TWP; ++j) {
              const int kw = c - j;
              if (kw >= 0 && kw < 3) {
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[kh * 3 + kw][k] = fmaf(v.v[k], g2[j], acc[kh * 3 + kw][k]);
              }
            }
          }
      }
    }
    __syncthreads();
    if (tid == 0 && i + ns < my_tiles) {
      fence_proxy_async();
      issue(i + ns);
    }
    if (++slot == ns) {
      slot = 0;
      phase ^= 1;
    }
  }
  // deterministic block reduction, one tap at a time (+ the bias sum) through [pl][cvl][8] floats
  float* red = reinterpret_cast<float*>(ring.bufs);
  float* row = partial + (int64_t)blockIdx.x * (C * 9 + 1);
#pragma unroll
  for (int tp = 0; tp < 10; ++tp) {
    __syncthreads();
    if (active) {
      float* mine = red + (pl * CVB + cvl) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) mine[k] = tp < 9 ? acc[tp < 9 ? tp : 0][k] : (k == 0 ? bsum : 0.f);
    }
    __syncthreads();
    for (int o = tid; o < CVB * 8; o += kDwThreads) {
      float sv = 0.f;
      for (int j = 0; j < G::PLn; ++j) sv += red[j * CVB * 8 + o];
      if (tp < 9) row[o * 9 + tp] = sv;
    }
    if (tp == 9 && tid == 0) {
      float sv = 0.f;
      for (int j = 0; j < G::PLn * CVB; ++j) sv += red[j * 8];
      row[C * 9] = sv;
    }
  }
}

__global__ void head_wgrad_sum_kernel(const float* __restrict__ partial, int nrows, int n, float* __restrict__ dw,
                                      float* __restrict__ dbias) {
  sc::pdl_wait();
  __shared__ float sh[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (i <= n)
    for (int r = ty; r < nrows; r += 32) s += partial[(int64_t)r * (n + 1) + i];
  sh[ty][tx] = s;
  __syncthreads();
  if (ty != 0 || i > n) return;
#pragma unroll
  for (int j = 1; j < 32; ++j) s += sh[j][tx];
  if (i < n) dw[i] += s;
  else if (dbias) dbias[0] += s;
}

__global__ void dw_wgrad_sum_kernel(const float* __restrict__ partial, int nrows, int n, float* __restrict__ dw) {
  sc::pdl_wait();
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
  sc::pdl_wait();
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

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
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

// channel-block widths with a compiled kernel, widest first
static int pick_cvb(int CV) {
  static const int opts[] = {8, 6, 4, 3, 2, 1};
  for (int d : opts)
    if (CV % d == 0) return d;
  return 1;
}

static int dw_grid_x(const DwTiles& g, int n_cb, int ctas_per_sm, int max_rows) {
  int64_t cap = ((int64_t)kNumSMs * ctas_per_sm + n_cb - 1) / n_cb;
  if (cap < 1) cap = 1;
  if (cap > max_rows) cap = max_rows;
  return (int)(g.n_tiles < cap ? g.n_tiles : cap);
}

static bool dw_tiles(int N, int Ho, int Wo, int TH, int TW, DwTiles* g) {
  g->tiles_h = (Ho + TH - 1) / TH;
  g->tiles_w = (Wo + TW - 1) / TW;
  int64_t n = (int64_t)N * g->tiles_h * g->tiles_w;
  if (n > INT32_MAX) return false;
  g->n_tiles = (int)n;
  return true;
}

// ring depth: as many tiles in flight as fit a per-CTA budget that keeps `ctas` CTAs resident per SM
static int dw_ring_depth(size_t stage_bytes, size_t extra, int ctas) {
  const size_t budget = (size_t)(227 * 1024) / ctas - 1024 - extra - 256;
  int ns = (int)(budget / stage_bytes);
  if (ns > 4) ns = 4;
  return ns;
}

template <typename T, int S, int CVB>
static int launch_fprop_cvb(const T* x, int ldx, const float* scale, const float* shift, int act, const float* w,
                            int mirror, T* y, int ldy, double* stats, int* stats_rows_host, int N, int H, int W, int C,
                            cudaStream_t st) {
  using G = DwGeo<S, CVB, (S == 1 ? 4 : 2)>;
  const int Ho = (H - 1) / S + 1, Wo = (W - 1) / S + 1;
  DwTiles g;
  if (!dw_tiles(N, Ho, Wo, G::TH, G::TW, &g)) return SC_ERR_BAD_ARG;
  const int n_cb = (C / 8) / CVB;
  size_t stage = ((size_t)G::TILE_ELEMS * sizeof(T) + 127) & ~(size_t)127;
  const size_t extra = (size_t)11 * CVB * 8 * sizeof(float) + 64;
  int ctas = DW_FPROP_CTAS;
  int ns = dw_ring_depth(stage, extra, ctas);
  while (ns < 2 && ctas > 1) ns = dw_ring_depth(stage, extra, --ctas);
  if (ns < 1) return SC_ERR_UNSUPPORTED;
  if ((size_t)ns * stage < (size_t)kDwThreads * 16 * sizeof(double)) stage = (size_t)kDwThreads * 16 * sizeof(double);
  const int gx = dw_grid_x(g, n_cb, ctas, stats ? SC_BN_MAX_PARTIALS : 1 << 30);
  const size_t smem = 128 + (size_t)ns * stage + extra;
  if (stats_rows_host) *stats_rows_host = gx;
  CUtensorMap tmX;
  if (!encode_nhwc_plain(&tmX, x, (int)sizeof(T), C, W, H, N, ldx, CVB * 8, G::IW, G::IH)) return SC_ERR_NO_DEVICE;
  int rc = set_smem(dw_fprop_kernel<T, S, CVB>, smem);
  if (rc != SC_OK) return rc;
  sc::launch_pdl((dw_fprop_kernel<T, S, CVB>), dim3(gx, n_cb), kDwThreads, smem, st, tmX, scale, shift, act, w, mirror, y, ldy, stats, H,
                                                                       W, C, Ho, Wo, g, ns, (int)stage);
  return check_launch();
}

template <typename T, int CVB>
static int launch_dgrad_s2_cvb(const T* dy, int lddy, const float* w, T* dx, int lddx, int N, int H, int W, int C,
                               cudaStream_t st) {
  using G = DwGeoD2<CVB>;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  DwTiles g;
  if (!dw_tiles(N, Ho, Wo, G::TH, G::TW, &g)) return SC_ERR_BAD_ARG;
  const int n_cb = (C / 8) / CVB;
  const size_t stage = ((size_t)G::TILE_ELEMS * sizeof(T) + 127) & ~(size_t)127;
  const size_t extra = (size_t)9 * CVB * 8 * sizeof(float) + 64;
  int ctas = 3;
  int ns = dw_ring_depth(stage, extra, ctas);
  while (ns < 2 && ctas > 1) ns = dw_ring_depth(stage, extra, --ctas);
  if (ns < 1) return SC_ERR_UNSUPPORTED;
  const int gx = dw_grid_x(g, n_cb, ctas, 1 << 30);
  const size_t smem = 128 + (size_t)ns * stage + extra;
  CUtensorMap tmDY;
  if (!encode_nhwc_plain(&tmDY, dy, (int)sizeof(T), C, Wo, Ho, N, lddy, CVB * 8, G::IW, G::IH)) return SC_ERR_NO_DEVICE;
  int rc = set_smem(dw_dgrad_s2_kernel<T, CVB>, smem);
  if (rc != SC_OK) return rc;
  sc::launch_pdl((dw_dgrad_s2_kernel<T, CVB>), dim3(gx, n_cb), kDwThreads, smem, st, tmDY, w, dx, lddx, H, W, C, g, ns, (int)stage);
  return check_launch();
}

template <typename T, int S, int CVB>
static int launch_wgrad_cvb(const T* x, int ldx, const float* scale, const float* shift, int act, const T* dy, int lddy,
                            float* workspace, int* rows, int N, int H, int W, int C, cudaStream_t st) {
  using G = DwGeo<S, CVB, 2>;
  const int Ho = (H - 1) / S + 1, Wo = (W - 1) / S + 1;
  DwTiles g;
  if (!dw_tiles(N, Ho, Wo, G::TH, G::TW, &g)) return SC_ERR_BAD_ARG;
  const int n_cb = (C / 8) / CVB;
  const size_t xal = ((size_t)G::TILE_ELEMS * sizeof(T) + 127) & ~(size_t)127;
  const size_t stage = xal + (((size_t)G::TH * G::TW * CVB * 8 * sizeof(T) + 127) & ~(size_t)127);
  const size_t extra = (size_t)2 * CVB * 8 * sizeof(float) + 64;
  int ctas = 2;
  int ns = dw_ring_depth(stage, extra, ctas);
  while (ns < 2 && ctas > 1) ns = dw_ring_depth(stage, extra, --ctas);
  if (ns < 1) return SC_ERR_UNSUPPORTED;
  const int gx = dw_grid_x(g, n_cb, ctas, kDwMaxRows);
  const size_t smem = 128 + (size_t)ns * stage + extra;     // ns*stage >= 8 KB reduction scratch
  *rows = gx;
  CUtensorMap tmX, tmDY;
  if (!encode_nhwc_plain(&tmX, x, (int)sizeof(T), C, W, H, N, ldx, CVB * 8, G::IW, G::IH) ||
      !encode_nhwc_plain(&tmDY, dy, (int)sizeof(T), C, Wo, Ho, N, lddy, CVB * 8, G::TW, G::TH))
    return SC_ERR_NO_DEVICE;
  int rc = set_smem(dw_wgrad_kernel<T, S, CVB>, smem);
  if (rc != SC_OK) return rc;
  sc::launch_pdl((dw_wgrad_kernel<T, S, CVB>), dim3(gx, n_cb), kDwThreads, smem, st, tmX, tmDY, scale, shift, act, workspace, H, W, C, g,
                                                                       ns, (int)stage);
  return check_launch();
}

template <typename T, int CVB>
static int launch_head_wgrad_cvb(const T* x, int ldx, const float* dl, float* dw, float* dbias, float* workspace, int N,
                                 int H, int W, int C, cudaStream_t st) {
  using G = DwGeo<1, CVB, 2>;
  DwTiles g;
  if (!dw_tiles(N, H, W, G::TH, G::TW, &g)) return SC_ERR_BAD_ARG;
  const size_t stage = ((size_t)G::TILE_ELEMS * sizeof(T) + 127) & ~(size_t)127;
  int ns = dw_ring_depth(stage, 64, 2);
  if (ns < 1) return SC_ERR_UNSUPPORTED;
  const int gx = dw_grid_x(g, 1, 2, kDwMaxRows);
  size_t smem = 128 + (size_t)ns * stage + 64;
  if (smem < (size_t)kDwThreads * 8 * sizeof(float) + 256) smem = (size_t)kDwThreads * 8 * sizeof(float) + 256;
  CUtensorMap tmX;
  if (!encode_nhwc_plain(&tmX, x, (int)sizeof(T), C, W, H, N, ldx, CVB * 8, G::IW, G::IH)) return SC_ERR_NO_DEVICE;
  int rc = set_smem(head_wgrad_tile_kernel<T, CVB>, smem);
  if (rc != SC_OK) return rc;
  sc::launch_pdl((head_wgrad_tile_kernel<T, CVB>), gx, kDwThreads, smem, st, tmX, dl, workspace, H, W, C, g, ns, (int)stage);
  rc = check_launch();
  if (rc != SC_OK) return rc;
  sc::launch_pdl((head_wgrad_sum_kernel), (C * 9 + 1 + 31) / 32, dim3(32, 32), 0, st, workspace, gx, C * 9, dw, dbias);
  return check_launch();
}

#define DW_DISPATCH_CVB(cvb, ...)                            \
  switch (cvb) {                                             \
    case 8: { constexpr int CVB = 8; __VA_ARGS__; } break;   \
    case 6: { constexpr int CVB = 6; __VA_ARGS__; } break;   \
    case 4: { constexpr int CVB = 4; __VA_ARGS__; } break;   \
    case 3: { constexpr int CVB = 3; __VA_ARGS__; } break;   \
    case 2: { constexpr int CVB = 2; __VA_ARGS__; } break;   \
    default: { constexpr int CVB = 1; __VA_ARGS__; } break;  \
  }

template <typename T>
static int launch_fprop(const T* x, int ldx, const float* scale, const float* shift, int act, const float* w,
                        int mirror, T* y, int ldy, double* stats, int* stats_rows_host, int N, int H, int W, int C,
                        int stride, cudaStream_t st) {
  const int cvb = pick_cvb(C / 8);
  if (stride == 1) {
    DW_DISPATCH_CVB(cvb, return (launch_fprop_cvb<T, 1, CVB>(x, ldx, scale, shift, act, w, mirror, y, ldy, stats,
                                                              stats_rows_host, N, H, W, C, st)));
  } else {
    DW_DISPATCH_CVB(cvb, return (launch_fprop_cvb<T, 2, CVB>(x, ldx, scale, shift, act, w, mirror, y, ldy, stats,
                                                              stats_rows_host, N, H, W, C, st)));
  }
  return SC_ERR_BAD_ARG;
}

template <typename T>
static int launch_dgrad_s2(const T* dy, int lddy, const float* w, T* dx, int lddx, int N, int H, int W, int C,
                           cudaStream_t st) {
  const int cvb = pick_cvb(C / 8);
  DW_DISPATCH_CVB(cvb, return (launch_dgrad_s2_cvb<T, CVB>(dy, lddy, w, dx, lddx, N, H, W, C, st)));
  return SC_ERR_BAD_ARG;
}

template <typename T>
static int launch_wgrad(const T* x, int ldx, const float* scale, const float* shift, int act, const T* dy, int lddy,
                        float* workspace, int* rows, int N, int H, int W, int C, int stride, cudaStream_t st) {
  const int cvb = pick_cvb(C / 8);
  if (stride == 1) {
    DW_DISPATCH_CVB(cvb, return (launch_wgrad_cvb<T, 1, CVB>(x, ldx, scale, shift, act, dy, lddy, workspace, rows, N, H,
                                                              W, C, st)));
  } else {
    DW_DISPATCH_CVB(cvb, return (launch_wgrad_cvb<T, 2, CVB>(x, ldx, scale, shift, act, dy, lddy, workspace, rows, N, H,
                                                              W, C, st)));
  }
  return SC_ERR_BAD_ARG;
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
  if ((H | W) & 1) {
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    int64_t total = (int64_t)N * H * W * (C / 8);
    int64_t b = (total + 255) / 256, cap = (int64_t)kNumSMs * 8;
    SC_DISPATCH_DTYPE(dtype, (sc::launch_pdl((dw_dgrad_generic_kernel<T>), (int)(b < cap ? b : cap), 256, 0, st, 
                                 (const T*)dy, lddy, w, (T*)dx, lddx, N, H, W, C, stride, Ho, Wo)));
    return check_launch();
  }
  SC_DISPATCH_DTYPE(dtype, return launch_dgrad_s2<T>((const T*)dy, lddy, w, (T*)dx, lddx, N, H, W, C, st));
  return SC_ERR_BAD_ARG;
}

// head wgrad through the tile machinery: C in {8, 16, 24, 32, 48, 64} (all channels in one block)
extern "C" int64_t sc_head_wgrad_workspace_bytes(int C) { return (int64_t)kDwMaxRows * (C * 9 + 1) * sizeof(float); }
extern "C" int sc_head_wgrad_tiled(const void* x, int ldx, const float* dlogits, float* dw, float* dbias, float* workspace,
                                   int N, int H, int W, int C, int dtype, void* stream) {
  if (!x || !dlogits || !dw || !workspace || C % 8 || C > 64 || ldx % 8 || N <= 0) return SC_ERR_BAD_ARG;
  const int cv = C / 8;
  if (cv != 1 && cv != 2 && cv != 3 && cv != 4 && cv != 6 && cv != 8) return SC_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  SC_DISPATCH_DTYPE(dtype, DW_DISPATCH_CVB(cv, return (launch_head_wgrad_cvb<T, CVB>((const T*)x, ldx, dlogits, dw, dbias,
                                                                                    workspace, N, H, W, C, st))));
  return SC_ERR_BAD_ARG;
}

extern "C" int64_t sc_dwconv_wgrad_workspace_bytes(int C) { return (int64_t)kDwMaxRows * C * 9 * sizeof(float); }

extern "C" int sc_dwconv_wgrad(const void* x, int ldx, const float* scale, const float* shift, int act,
                               const void* dy, int lddy, float* dw, float* workspace, int N, int H, int W, int C,
                               int stride, int dtype, void* stream) {
  if (!x || !dy || !dw || !workspace || C % 8 || ldx % 8 || lddy % 8 || (stride != 1 && stride != 2) || N <= 0)
    return SC_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int rows = 0, rc = SC_ERR_BAD_ARG;
  SC_DISPATCH_DTYPE(dtype, rc = launch_wgrad<T>((const T*)x, ldx, scale, shift, act, (const T*)dy, lddy, workspace, &rows,
                                                N, H, W, C, stride, st));
  if (rc != SC_OK) return rc;
  sc::launch_pdl((dw_wgrad_sum_kernel), (C * 9 + 31) / 32, dim3(32, 32), 0, st, workspace, rows, C * 9, dw);
  return check_launch();
}
