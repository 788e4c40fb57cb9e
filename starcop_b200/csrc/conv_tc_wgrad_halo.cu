// tcgen05 weight gradient of the THIN 3x3 / stride-1 / pad-1 layers (decoder blocks 3-4: 16...80 input channels,
// 16 / 32 output channels at 256^2...512^2).
//
// The general wgrad kernel (conv_tc.cu) fetches one shifted activation box per filter tap; at these channel counts
// a box row is 32...64 bytes and TMA retires about one box ROW per ~2 cycles whatever its width, so the kernel is
// bound by the NUMBER of rows it asks for: 9 taps x (pixels) + dY, i.e. ~11 rows per output pixel.
//
// Here the activation HALO PATCH of a 16 x 8 pixel tile is fetched ONCE per 32-channel group (one TMA box,
// (8+2) x PW pixel rows), and the three horizontal taps come from ONE MMA through the M dimension:
//   * the patch is stored pixel-major, [patch pixel][GW channels], rows of GW*2 bytes with the matching TMA
//     swizzle (32B for GW = 16, 64B for GW = 32): this is the MN-major UMMA operand layout whose "MN blocks" are
//     one swizzle row wide (GW channels) and LBO apart;
//   * choosing LBO = ONE ROW makes MN block j the same pixels shifted by j: accumulator row (j, ci) of
//       D[(j, ci)][co] += sum_k X[pixel k + j][ci] * dY[pixel k][co]            (K = 16 pixels of one image row)
//     is the weight gradient of horizontal tap kw = j (lag j - 1) -- M = 128 rows hold 4 (GW = 32) or 8 (GW = 16)
//     lags, of which the first three are the filter taps;
//   * the vertical taps are three start addresses (patch row r + kh) into the same patch, three accumulators.
//   ~2.5 rows per output pixel instead of ~11, every input byte crosses L2 -> SM once.
// The accumulators (channel groups x 3 vertical taps x Cout fp32 columns) stay in TMEM for the CTA's whole
// lifetime; each CTA writes ONE partial gradient, and a second kernel sums the CTAs' partials in CTA order
// (deterministic, no floating-point atomics).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

using namespace sc;
using namespace tc;

namespace {

constexpr int kThreads = 192;
constexpr int kTileW = 16, kTileH = 8;          // 128 output pixels per stage
constexpr int kPatchH = kTileH + 2;
constexpr int kMaxSmem = 227 * 1024;

struct WgHaloParams {
  int N, H, W, Cin, Cout;
  int ng;                        // channel groups
  int tiles_w, tiles_h, n_tiles;
  int stages, stage_bytes, patch_bytes;    // patch_bytes: smem pitch of one group's patch (1 KB multiple)
  int box_bytes;                           // bytes one patch box actually delivers
  int tmem_cols;
  int debug;                               // timing experiments only: 1 = no TMA after the ring is primed, 2 = no MMAs
  int nsplit;                              // independent accumulator sets (image rows r % nsplit): hides the MMA dependency latency
  float* partials;               // [grid][ng][3 kh][3 kw][GW][Cout]
  float* dw;
};

template <int GW>
struct Geo {
  static constexpr int LAGS = 128 / GW;                    // 8 or 4 pixel shifts stacked along M
  static constexpr int PW = (kTileW + LAGS - 1 + 3) & ~3;  // patch width: columns w0-1 ... w0+PW-2
  static constexpr int ROW = GW * 2;                       // bytes per patch pixel = swizzle span
};

template <int GW>
__global__ void __launch_bounds__(kThreads, 2)
tc_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, WgHaloParams p) {
  sc::pdl_wait();
  using G = Geo<GW>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int rowB = p.Cout * 2;                             // bytes per dY pixel = its swizzle span

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = (uint32_t)(p.ng * p.box_bytes + kTileW * kTileH * rowB);
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int tw = tile % p.tiles_w;
        const int t2 = tile / p.tiles_w;
        const int th = t2 % p.tiles_h;
        const int img = t2 / p.tiles_h;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* s = smem + (size_t)stage * p.stage_bytes;
        if (p.debug == 1 && i >= p.stages) {
          mbar_arrive(&full[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
          continue;
        }
        mbar_arrive_expect_tx(&full[stage], tx);
        for (int g = 0; g < p.ng; ++g)
          tma_load_4d(s + (size_t)g * p.patch_bytes, &tmX, g * GW, tw * kTileW - 1, th * kTileH - 1, img, &full[stage]);
        tma_load_4d(s + (size_t)p.ng * p.patch_bytes, &tmDY, 0, tw * kTileW, th * kTileH, img, &full[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_bf16(128, p.Cout, 1, 1), idesc2 = make_idesc_bf16(128, 2 * p.Cout, 1, 1),
                     idesc3 = make_idesc_bf16(128, 3 * p.Cout, 1, 1);
      constexpr uint32_t A_LAYOUT = layout_for_row_bytes(G::ROW);
      const uint32_t b_layout = layout_for_row_bytes(rowB);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sbase = smem_u32(smem + (size_t)stage * p.stage_bytes);
        const uint32_t bbase = sbase + (uint32_t)(p.ng * p.patch_bytes);
        if (p.debug != 2) {
          // One MMA per PATCH row rho (image row rho - 1): A = that row's pixels (M = lags x channels), B = the dY
          // rows it meets as vertical taps -- rows rho-2, rho-1, rho of the tile stacked along N (N block stride =
          // one dY image row), so accumulator column block b collects tap kh = 2 - b.  Rows 0, 1, 8, 9 of the patch
          // meet only the dY rows that exist: narrower N.  Row 2 goes first so that the very first MMA of a CTA
          // (accumulate = 0) initialises all three column blocks.
          const uint32_t brow = kTileW * rowB;
          for (int g = 0; g < p.ng; ++g) {
            const uint32_t abase = sbase + (uint32_t)(g * p.patch_bytes);
            const uint32_t d = tmem_base + (uint32_t)(g * 3 * p.Cout);
#pragma unroll
            for (int k = 0; k < kPatchH; ++k) {
              const int rho = k < 6 ? k + 2 : (k == 6 ? 0 : (k == 7 ? 1 : k));      // 2..7, 0, 1, 8, 9
              const int b0 = rho < 2 ? 2 - rho : 0;                                  // first column block
              const int b1 = rho > 7 ? 9 - rho : 2;                                  // last column block
              const int nb = b1 - b0 + 1;
              const int r0 = rho - 2 + b0;                                           // first dY row used
              const uint64_t ad = make_smem_desc(abase + (uint32_t)(rho * G::PW) * G::ROW, G::ROW, 8 * G::ROW, A_LAYOUT);
              const uint64_t bd = make_smem_desc(bbase + (uint32_t)r0 * brow, brow, 8 * rowB, b_layout);
              umma_bf16(d + (uint32_t)(b0 * p.Cout), ad, bd, nb == 3 ? idesc3 : (nb == 2 ? idesc2 : idesc1), (i > 0 || k > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(&empty[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tfull);
    }
  } else if (my_tiles > 0) {
    // epilogue, once per CTA: accumulator row m = (lag j, channel ci); lags 0..2 are the taps kw = 0..2
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int j = m / GW, ci = m % GW;
    mbar_wait(tfull, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* mine = p.partials + (size_t)blockIdx.x * p.ng * 9 * GW * p.Cout;
    if (q * 32 < 3 * GW) {                                   // warp-uniform: this warp holds useful rows
      for (int g = 0; g < p.ng; ++g)
        for (int kh = 0; kh < 3; ++kh)
          for (int c0 = 0; c0 < p.Cout; c0 += 16) {
            float v[16];
            tmem_ld16(taddr + (uint32_t)((g * 3 + kh) * p.Cout + c0), v);
            for (int sp = 1; sp < p.nsplit; ++sp) {          // fold the accumulator sets in a fixed order
              float u[16];
              tmem_ld16(taddr + (uint32_t)(((sp * p.ng + g) * 3 + kh) * p.Cout + c0), u);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += u[i];
            }
            if (j < 3) {
              float* o = mine + ((size_t)((g * 3 + (2 - kh)) * 3 + j) * GW + ci) * p.Cout + c0;   // column block b holds tap kh = 2 - b
#pragma unroll
              for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
          }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// dW[co][ci][kh][kw] += sum over CTAs of partial[cta][g][kh][kw][ci_local][co], in a FIXED order: warp w of the block sums
// the rows r = w, w + 32, ... (coalesced 128-byte reads of 32 neighbouring elements), warp 0 then adds the 32 slice sums
// in slice order.  (One thread per element walking all ~300 rows serially was latency-bound: 63 us for 2304 elements.)
template <int GW>
__global__ void __launch_bounds__(1024)
wgrad_halo_reduce_kernel(const float* __restrict__ partials, int nrows, int ng, int Cin, int Cout, float* __restrict__ dw) {
  sc::pdl_wait();
  __shared__ float slice[32][33];
  const int per = ng * 9 * GW * Cout;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int e0 = blockIdx.x * 32; e0 < per; e0 += gridDim.x * 32) {
    const int e = e0 + lane;
    float s = 0.f;
    if (e < per)
      for (int r = w; r < nrows; r += 32) s += __ldcg(partials + (size_t)r * per + e);
    slice[w][lane] = s;
    __syncthreads();
    if (w == 0 && e < per) {
      float t = slice[0][lane];
#pragma unroll
      for (int k = 1; k < 32; ++k) t += slice[k][lane];
      const int co = e % Cout;
      int q = e / Cout;
      const int cil = q % GW;
      q /= GW;
      const int kw = q % 3;
      q /= 3;
      const int kh = q % 3;
      const int g = q / 3;
      const int ci = g * GW + cil;
      if (ci < Cin) dw[((size_t)co * Cin + ci) * 9 + kh * 3 + kw] += t;
    }
    __syncthreads();
  }
}

bool encode_sw(CUtensorMap* m, const void* ptr, int C, int W, int H, int N, int ld, int bc, int bw, int bh) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const int rb = bc * 2;
  const CUtensorMapSwizzle sw = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Plan {
  int gw, ng, grid, stages, stage_bytes, patch_bytes, tmem_cols, nsplit;
};

bool make_plan(int N, int H, int W, int Cin, int Cout, Plan* pl) {
  if (getenv("STARCOP_NO_WGRAD_HALO")) return false;
  if (Cin < 16 || Cin % 8 || (Cout != 16 && Cout != 32) || H % kTileH || W % kTileW) return false;
  pl->gw = Cin <= 16 ? 16 : 32;
  pl->ng = (Cin + pl->gw - 1) / pl->gw;
  const int cols = pl->ng * 3 * Cout;
  if (cols > 512) return false;
  pl->nsplit = 1;
  pl->tmem_cols = 32;
  while (pl->tmem_cols < cols * pl->nsplit) pl->tmem_cols *= 2;
  const int pw = pl->gw == 16 ? Geo<16>::PW : Geo<32>::PW;
  pl->patch_bytes = kPatchH * pw * pl->gw * 2;                        // multiples of 1 KB: 7680 -> pad
  pl->patch_bytes = (pl->patch_bytes + 1023) & ~1023;
  pl->stage_bytes = pl->ng * pl->patch_bytes + ((kTileW * kTileH * Cout * 2 + 1023) & ~1023);
  // Neither the TMA stream of narrow rows nor the chain of small MMAs saturates its unit: two (or more) CTAs per SM
  // overlap one CTA's MMA latency with the other's loads.  TMEM (512 columns per SM) and shared memory decide.
  // Measured (bs 16): 32->16 @512^2 154 -> 123 us, 16->16 @512^2 149 -> 115 us with two CTAs; layers with few tiles
  // per CTA lose (more partials to reduce, shorter accumulation runs): 32->32 @256^2 64 -> 73 us.
  const int64_t nt = (int64_t)N * (H / kTileH) * (W / kTileW);
  if (nt > INT32_MAX) return false;
  int ctas = (512 / pl->tmem_cols >= 2 && nt >= (int64_t)kNumSMs * 2 * 64) ? 2 : 1;
  while (ctas > 1 && ((220 * 1024) / ctas - 2048) / pl->stage_bytes < 3) --ctas;
  int stages = ((220 * 1024) / ctas - 2048) / pl->stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) return false;
  pl->stages = stages;
  pl->grid = (int)(nt < kNumSMs * ctas ? nt : kNumSMs * ctas);
  return true;
}

}  // namespace

// 1 when sc_tc_conv_wgrad routes this 3x3 / stride-1 layer to the halo kernel
extern "C" int sc_tc_wgrad_halo_supported(int N, int H, int W, int Cin, int Cout) {
  Plan pl;
  return make_plan(N, H, W, Cin, Cout, &pl) ? 1 : 0;
}
extern "C" int64_t sc_tc_wgrad_halo_workspace_bytes(int N, int H, int W, int Cin, int Cout) {
  Plan pl;
  if (!make_plan(N, H, W, Cin, Cout, &pl)) return -1;
  return (int64_t)pl.grid * pl.ng * 9 * pl.gw * Cout * sizeof(float);
}

extern "C" int sc_tc_wgrad_halo(const void* x, int ldx, const void* dy, int lddy, float* dw_oihw, float* partials, int N,
                                int H, int W, int Cin, int Cout, void* stream) {
  if (!x || !dy || !dw_oihw || !partials || ldx % 8 || lddy % 8) return SC_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15)) return SC_ERR_BAD_ARG;
  Plan pl;
  if (!make_plan(N, H, W, Cin, Cout, &pl)) return SC_ERR_UNSUPPORTED;
  WgHaloParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ng = pl.ng;
  p.tiles_w = W / kTileW; p.tiles_h = H / kTileH; p.n_tiles = N * p.tiles_w * p.tiles_h;
  p.stages = pl.stages; p.stage_bytes = pl.stage_bytes; p.patch_bytes = pl.patch_bytes; p.tmem_cols = pl.tmem_cols;
  p.nsplit = pl.nsplit;
  p.debug = getenv("STARCOP_WGH_DEBUG") ? atoi(getenv("STARCOP_WGH_DEBUG")) : 0;
  p.box_bytes = kPatchH * (pl.gw == 16 ? Geo<16>::PW : Geo<32>::PW) * pl.gw * 2;
  p.partials = partials; p.dw = dw_oihw;
  const int pw = pl.gw == 16 ? Geo<16>::PW : Geo<32>::PW;
  CUtensorMap tmX, tmDY;
  if (!encode_sw(&tmX, x, Cin, W, H, N, ldx, pl.gw, pw, kPatchH) || !encode_sw(&tmDY, dy, Cout, W, H, N, lddy, Cout, kTileW, kTileH))
    return SC_ERR_NO_DEVICE;
  const size_t smem = (size_t)pl.stages * pl.stage_bytes + 1024 + 256;
  cudaStream_t st = (cudaStream_t)stream;
  const int per = pl.ng * 9 * pl.gw * Cout;
  const int rblocks = (per + 31) / 32;
  cudaError_t e;
  if (pl.gw == 16) {
    e = cudaFuncSetAttribute(tc_wgrad_halo_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }
    sc::launch_pdl((tc_wgrad_halo_kernel<16>), pl.grid, kThreads, smem, st, tmX, tmDY, p);
    sc::launch_pdl((wgrad_halo_reduce_kernel<16>), rblocks, 1024, 0, st, partials, pl.grid, pl.ng, Cin, Cout, dw_oihw);
  } else {
    e = cudaFuncSetAttribute(tc_wgrad_halo_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }
    sc::launch_pdl((tc_wgrad_halo_kernel<32>), pl.grid, kThreads, smem, st, tmX, tmDY, p);
    sc::launch_pdl((wgrad_halo_reduce_kernel<32>), rblocks, 1024, 0, st, partials, pl.grid, pl.ng, Cin, Cout, dw_oihw);
  }
  return check_launch();
}
