// tcgen05 tensor-core convolutions for the HyperSTARCOP U-Net (bf16 storage, fp32 accumulate in TMEM).
//
// fprop / dgrad:  implicit GEMM, M = 128 output pixels (an 8 x 16 spatial patch), N = Cout tile,
//   K = taps x Cin.  The A tile of filter tap (dy,dx) is ONE TMA box of the NHWC activation tensor
//   shifted by (dy,dx); out-of-image elements are zero-filled by TMA, which IS the conv padding.
//   B = packed bf16 weights [Cout][tap][Cin] (K-major).  Warp-specialised, persistent:
//   warp 0 TMA producer, warp 1 MMA issuer (single thread) + TMEM owner, warps 2..5 epilogue;
//   smem ring of `stages` (A,B) tiles, two TMEM accumulator stages so the epilogue of tile i
//   overlaps the MMAs of tile i+1.
// wgrad:  dW[(tap,ci)][co] = sum_pixels X[pixel@tap][ci] * dY[pixel][co]; both operands are read in
//   their natural NHWC layout as MN-major UMMA operands (the reduction runs over pixels), M = 128
//   rows made of 128/KC shifted activation boxes, split over pixels across CTAs.  Every split writes
//   its 128 x block_n partial tile to a workspace; the LAST split of a tile to finish (a ticket counter)
//   sums the partials in split order and adds the result into the torch-layout gradient: no floating
//   point atomics, run-to-run bit-identical.
#include "common.cuh"
#include "tc_common.cuh"

using namespace sc;
using namespace tc;

// ------------------------------------------------------------------------------------------------
// driver entry point for tensor-map encoding (no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
static EncodeTiledFn get_encode() { return encode_tiled_fn(); }

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// NHWC bf16 activation tensor (C, W, H, N) with channel stride ld; box = (kc, bw, bh, 1)
// es > 1: the box samples every es-th pixel (strided convolution); TMA copies ceil(box/es) elements
static int encode_act(CUtensorMap* m, const void* ptr, int C, int W, int H, int N, int ld, int kc, int bw, int bh,
                      int es_ = 1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return SC_ERR_NO_DEVICE;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)(bw * es_), (cuuint32_t)(bh * es_), 1};
  cuuint32_t es[4] = {1, (cuuint32_t)es_, (cuuint32_t)es_, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SC_OK : SC_ERR_BAD_ARG;
}
// row-major bf16 matrix (cols, rows); box = (kc, brows)
static int encode_mat(CUtensorMap* m, const void* ptr, int64_t cols, int64_t rows, int kc, int brows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return SC_ERR_NO_DEVICE;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)brows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SC_OK : SC_ERR_BAD_ARG;
}

extern "C" int sc_tc_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return (major == 10 && get_encode() != nullptr) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// weights: OIHW f32 -> bf16 [Cout][tap][Cin] (optionally the flipped / transposed dgrad filter)
// ------------------------------------------------------------------------------------------------
__global__ void tc_pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ o, int Cout, int Cin,
                                       int KK, int flip_t, int cin_pad, int cout_pad) {
  // output logical shape [rows][KK][cols]: rows = Cout (or Cin when flip_t), cols = Cin (or Cout)
  int rows = flip_t ? cin_pad : cout_pad, cols = flip_t ? cout_pad : cin_pad;
  int64_t total = (int64_t)rows * KK * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % cols);
    int64_t t = i / cols;
    int tap = (int)(t % KK);
    int r = (int)(t / KK);
    float v = 0.f;
    if (!flip_t) {
      if (r < Cout && c < Cin) v = w[((int64_t)r * Cin + c) * KK + tap];
    } else {
      // dgrad: out channel r = ci, in channel c = co, tap mirrored
      if (r < Cin && c < Cout) v = w[((int64_t)c * Cin + r) * KK + (KK - 1 - tap)];
    }
    o[i] = __float2bfloat16_rn(v);
  }
}

extern "C" int sc_tc_pack_weights(const float* w_oihw, void* w_bf16, int Cout, int Cin, int KH, int KW,
                                  int flip_transpose, int cin_pad, int cout_pad, void* stream) {
  if (!w_oihw || !w_bf16 || cin_pad < Cin || cout_pad < Cout) return SC_ERR_BAD_ARG;
  int64_t total = (int64_t)cin_pad * cout_pad * KH * KW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  tc_pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, (__nv_bfloat16*)w_bf16, Cout, Cin, KH * KW,
                                                                   flip_transpose, cin_pad, cout_pad);
  return check_launch();
}

// All of a step's weight re-packs in ONE launch: `descs` is a device table of n jobs (see sc_tc_pack_desc in the
// header), job j covering output elements [offset_j, offset_{j+1}) of the concatenated index space.
__global__ void tc_pack_weights_batch_kernel(const sc_tc_pack_desc* __restrict__ descs, int n, int64_t total) {
  sc::pdl_wait();
  __shared__ int64_t s_off[129];
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_off[i] = descs[i].offset;
  if (threadIdx.x == 0) s_off[n] = total;
  __syncthreads();
  // Two consecutive outputs per thread (every job's row length is a multiple of 16, its offset even: a pair never
  // straddles a row or a job), 32-bit index arithmetic (a job is < 2^31 elements; the four 64-bit divisions per
  // element made this kernel 9x slower than its 78 MB of traffic), one 4-byte store.
  for (int64_t g = 2 * (blockIdx.x * (int64_t)blockDim.x + threadIdx.x); g < total; g += 2 * (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {                       // last job whose offset <= g
      const int mid = (lo + hi + 1) >> 1;
      if (s_off[mid] <= g) lo = mid; else hi = mid - 1;
    }
    const sc_tc_pack_desc d = descs[lo];
    const uint32_t i = (uint32_t)(g - d.offset);
    const uint32_t cols = (uint32_t)(d.flip_transpose ? d.cout_pad : d.cin_pad);
    const uint32_t t = i / cols, c = i - t * cols;
    const uint32_t r = t / (uint32_t)d.kk, tap = t - r * (uint32_t)d.kk;
    float v0 = 0.f, v1 = 0.f;
    if (!d.flip_transpose) {
      if (r < (uint32_t)d.cout) {
        const float* wp = d.w + ((int64_t)r * d.cin + c) * d.kk + tap;
        if (c < (uint32_t)d.cin) v0 = wp[0];
        if (c + 1 < (uint32_t)d.cin) v1 = wp[d.kk];
      }
    } else {
      if (r < (uint32_t)d.cin) {
        const float* wp = d.w + ((int64_t)c * d.cin + r) * d.kk + (d.kk - 1 - tap);
        if (c < (uint32_t)d.cout) v0 = wp[0];
        if (c + 1 < (uint32_t)d.cout) v1 = wp[(int64_t)d.cin * d.kk];
      }
    }
    *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(d.out) + i) = __floats2bfloat162_rn(v0, v1);
  }
}

extern "C" int sc_tc_pack_weights_batch(const sc_tc_pack_desc* descs_dev, int n, int64_t total, void* stream) {
  if (!descs_dev || n < 1 || n > 128 || total < 1) return SC_ERR_BAD_ARG;
  int64_t blocks = (total / 2 + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  sc::launch_pdl((tc_pack_weights_batch_kernel), (int)blocks, 256, 0, (cudaStream_t)stream, descs_dev, n, total);
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// fprop
// ------------------------------------------------------------------------------------------------
struct FpropParams {
  int N, H, W, Cin, Cout, KH, KW;   // H, W: OUTPUT spatial size
  int stride;
  int tiles_w, tiles_h;      // spatial patches per image
  int m_tiles, n_tiles;
  FastDiv div_n, div_w, div_h;   // tile index -> (N tile, image, patch row, patch column) without integer divisions
  int block_n;               // multiple of 16, <= 256
  int cchunks;               // ceil(Cin / KC)
  int groups;                // (tap, channel chunk) K groups per pipeline stage
  int stages;
  int ldy;
  int accumulate;            // y += result (gradient fan-in)
  int tmem_cols;             // 2 accumulator stages: 512, or 256 when two CTAs share an SM
  int b_resident;            // > 0: byte offset (from the barrier block) of the CTA-resident weights [n tile][K step]
  __nv_bfloat16* y;
  double* stats;             // optional [2][Cout] fp64 sum / sum of squares of the stored outputs
};

constexpr int kTcThreads = 192;
#ifndef SC_EPI_BATCH
#define SC_EPI_BATCH 4
#endif
constexpr int kEpiBatch = SC_EPI_BATCH;   // TMEM loads in flight behind one wait in the store-bound epilogue
constexpr int kMaxDynSmem = 227 * 1024;

template <int KC>
__global__ void __launch_bounds__(kTcThreads, 2)
tc_conv_fprop_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, FpropParams p) {
  sc::pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int ROW_BYTES = KC * 2;
  constexpr int A_BYTES = 128 * ROW_BYTES;
  const int B_BYTES = (p.block_n * ROW_BYTES + 1023) & ~1023;
  const int G = p.groups;
  const int STAGE_BYTES = G * (A_BYTES + (p.b_resident ? 0 : B_BYTES));       // [G x A][G x B], every box 1024-aligned
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * STAGE_BYTES);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint64_t* tempty = tfull + 2;
  uint64_t* bres = tempty + 2;                                  // the resident weight tile has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres + 1);
  float* s_stats = reinterpret_cast<float*>(tmem_slot + 4);     // [2][Cout] per-CTA partial statistics
  float* s_part = s_stats + 2 * p.Cout;                         // [4 warps][2][block_n] one tile's column sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
    mbar_init(bres, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int ksteps = p.KH * p.KW * p.cchunks;
  const int nstages_per_tile = (ksteps + G - 1) / G;
  const int pad = p.KH / 2;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // Layers whose whole packed weight matrix fits beside the pipeline: fetched ONCE per CTA.  TMA retires a box row
      // every few cycles whatever its width, and with 16-32 input channels the Cout weight rows of a tile cost as much
      // as its 128 activation rows (f2.exp fprop, no epilogue work at all: 49 us for 33 MB of input; 89 -> 52 us with
      // the stores once the weights were resident)
      uint8_t* sBres = reinterpret_cast<uint8_t*>(full) + p.b_resident;
      if (p.b_resident) {
        mbar_arrive_expect_tx(bres, (uint32_t)(p.n_tiles * ksteps * p.block_n * ROW_BYTES));
        for (int nt = 0; nt < p.n_tiles; ++nt)
          for (int ks = 0; ks < ksteps; ++ks)
            tma_load_2d(sBres + (size_t)(nt * ksteps + ks) * B_BYTES, &tmB, ks * KC, nt * p.block_n, bres);
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int mt = (int)fast_div((uint32_t)tile, p.div_n), nt = tile - mt * p.n_tiles;
        int t2 = (int)fast_div((uint32_t)mt, p.div_w);
        int tw = mt - t2 * p.tiles_w;
        int img = (int)fast_div((uint32_t)t2, p.div_h);
        int th = t2 - img * p.tiles_h;
        int n0 = nt * p.block_n;
        const int w0 = (tw * 16) * p.stride - pad, h0 = (th * 8) * p.stride - pad;
        const int group_tx = A_BYTES + (p.b_resident ? 0 : p.block_n * ROW_BYTES);
        int cc = 0, kw = 0, kh = 0;             // running (channel chunk, tap) of the next K group: no divisions
        for (int s0 = 0; s0 < ksteps; s0 += G) {
          const int g_here = min(G, ksteps - s0);
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sA = smem + (size_t)stage * STAGE_BYTES;
          uint8_t* sB = sA + G * A_BYTES;
          mbar_arrive_expect_tx(&full[stage], g_here * group_tx);
          for (int g = 0; g < g_here; ++g) {
            tma_load_4d(sA + g * A_BYTES, &tmA, cc * KC, w0 + kw, h0 + kh, img, &full[stage]);
            if (!p.b_resident) tma_load_2d(sB + g * B_BYTES, &tmB, (s0 + g) * KC, n0, &full[stage]);
            if (++cc == p.cchunks) {
              cc = 0;
              if (++kw == p.KW) {
                kw = 0;
                ++kh;
              }
            }
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.block_n, 0, 0);
      constexpr uint32_t LAYOUT = layout_for_row_bytes(ROW_BYTES);
      constexpr uint32_t SBO = 8 * ROW_BYTES;
      const uint64_t desc0 = make_smem_desc(smem_u32(smem), 16, SBO, LAYOUT);
      const uint64_t desc_hi = desc0 & 0xffffffff00000000ull;
      const uint32_t desc_lo0 = (uint32_t)desc0;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t bres_lo = desc_lo0 + (uint32_t)((p.stages * STAGE_BYTES + p.b_resident) >> 4);
      if (p.b_resident) {
        mbar_wait(bres, 0);
        tc_fence_after();
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (p.tmem_cols / 2);
        const int nt_mma = tile - (int)fast_div((uint32_t)tile, p.div_n) * p.n_tiles;
        const uint32_t bres_tile = bres_lo + (uint32_t)(nt_mma * ksteps) * (uint32_t)(B_BYTES >> 4);
        for (int si = 0; si < nstages_per_tile; ++si) {
          const int g_here = min(G, ksteps - si * G);
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          // descriptors: assembled once (desc0), then a stage / group / K step is an add on the low word (start
          // address in 16-byte units; shared memory < 256 KB, the 14-bit field never carries) -- the issuing thread
          // is the limit on the thin layers, every instruction between two MMAs counts
          const uint32_t a_lo = desc_lo0 + (uint32_t)stage * (uint32_t)(STAGE_BYTES >> 4);
          const uint32_t b_lo = p.b_resident ? bres_tile + (uint32_t)(si * G) * (uint32_t)(B_BYTES >> 4)
                                             : a_lo + (uint32_t)(G * A_BYTES >> 4);
          for (int g = 0; g < g_here; ++g) {
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) {
              const uint64_t ad = desc_hi | (uint64_t)(a_lo + (uint32_t)(g * (A_BYTES >> 4) + k * 2));
              const uint64_t bd = desc_hi | (uint64_t)(b_lo + (uint32_t)(g * (B_BYTES >> 4) + k * 2));
              umma_bf16(d_tmem, ad, bd, idesc, (si > 0 || g > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // epilogue: warp q reads TMEM lanes [32q, 32q+32)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;       // 0..127
    if (p.stats) {
      for (int i = et; i < 2 * p.Cout; i += 128) s_stats[i] = 0.f;
      asm volatile("bar.sync 1, 128;");
    }
    // after the recursive-halving reduction below, this lane owns column (lane >> 1) of each 16-column group
    const int my_col = lane >> 1;
    // every 16-channel group of a pixel is a 32-byte aligned sector
    const bool wide_ok = !p.accumulate && (p.ldy % 16) == 0 && (p.block_n % 16) == 0 &&
                         (reinterpret_cast<uintptr_t>(p.y) & 31) == 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      int mt = (int)fast_div((uint32_t)tile, p.div_n), nt = tile - mt * p.n_tiles;
      int t2 = (int)fast_div((uint32_t)mt, p.div_w);
      int tw = mt - t2 * p.tiles_w;
      int img = (int)fast_div((uint32_t)t2, p.div_h);
      int th = t2 - img * p.tiles_h;
      int n0 = nt * p.block_n;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      int h = th * 8 + row / 16, w = tw * 16 + row % 16;
      __nv_bfloat16* yp = p.y + (((int64_t)img * p.H + h) * p.W + w) * p.ldy + n0;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (p.tmem_cols / 2);
      int c_begin = 0;
      if (wide_ok && !p.stats) {
        // store-bound layers (the 1x1 expansions: one MMA per tile, 12-120 KB of output): up to four TMEM loads in
        // flight behind ONE tcgen05.wait::ld, then the 256-bit stores back to back -- the per-chunk load -> wait ->
        // store chain cost one TMEM round trip per 16 channels (f2.exp fprop 112 us for 235 MB)
        const int full = min(p.block_n, (p.Cout - n0) & ~15);       // columns covered by complete 16-channel groups
        for (; c_begin + 16 <= full; c_begin += 16 * kEpiBatch) {
          uint32_t r[kEpiBatch][16];
          const int ng = min(kEpiBatch, (full - c_begin) / 16);
#pragma unroll
          for (int g = 0; g < kEpiBatch; ++g)
            if (g < ng) tmem_ld16_nowait(taddr + c_begin + 16 * g, r[g]);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < kEpiBatch; ++g) {
            if (g < ng) {
              uint32_t w8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[g][2 * i]), __uint_as_float(r[g][2 * i + 1]));
                w8[i] = *reinterpret_cast<uint32_t*>(&h2);
              }
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(yp + c_begin + 16 * g), "r"(w8[0]), "r"(w8[1]),
                           "r"(w8[2]), "r"(w8[3]), "r"(w8[4]), "r"(w8[5]), "r"(w8[6]), "r"(w8[7])
                           : "memory");
            }
          }
        }
        c_begin = full;
      }
      for (int c0 = c_begin; c0 < p.block_n; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        if (wide_ok && n0 + c0 + 16 <= p.Cout) {
          // 16 channels = one 32-byte sector per thread: ONE 256-bit store (STG.256) instead of two half-sector
          // stores -- the wide 1x1 expansions are bound by these scattered stores, not by the MMAs
          uint32_t w8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w8[i] = *reinterpret_cast<uint32_t*>(&h2);
            if (p.stats) {
              const float2 f = __bfloat1622float2(h2);
              v[2 * i] = f.x;
              v[2 * i + 1] = f.y;
            }
          }
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(yp + c0), "r"(w8[0]), "r"(w8[1]), "r"(w8[2]),
                       "r"(w8[3]), "r"(w8[4]), "r"(w8[5]), "r"(w8[6]), "r"(w8[7])
                       : "memory");
        } else
        // channels are handled in 8-wide halves (Cout % 8 == 0; a ragged last N tile is masked)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float* vv = v + hf * 8;
          if (n0 + c0 + hf * 8 >= p.Cout) {
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] = 0.f;
            continue;
          }
          uint4 u;
          __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
          if (p.accumulate) {
            uint4 o = *reinterpret_cast<const uint4*>(yp + c0 + hf * 8);
            const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&o);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float2 f = __bfloat1622float2(oh[i]);
              vv[2 * i] += f.x;
              vv[2 * i + 1] += f.y;
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) hh[i] = __floats2bfloat162_rn(vv[2 * i], vv[2 * i + 1]);
          *reinterpret_cast<uint4*>(yp + c0 + hf * 8) = u;
          if (p.stats) {
            // statistics are taken of the STORED (bf16-rounded) values
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float2 f = __bfloat1622float2(hh[i]);
              vv[2 * i] = f.x;
              vv[2 * i + 1] = f.y;
            }
          }
        }
        if (p.stats) {
          // column sums over the warp's 32 pixel rows by recursive halving: 16+16 shuffles per 16 columns
          float sq[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) sq[i] = v[i] * v[i];
#pragma unroll
          for (int wdt = 8, mask = 16; wdt >= 1; wdt >>= 1, mask >>= 1) {
            const bool upper = (lane & mask) != 0;
#pragma unroll
            for (int i = 0; i < wdt; ++i) {
              float send = upper ? v[i] : v[i + wdt];
              float keep = upper ? v[i + wdt] : v[i];
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
              float send2 = upper ? sq[i] : sq[i + wdt];
              float keep2 = upper ? sq[i + wdt] : sq[i];
              sq[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, mask);
            }
          }
          float s1 = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
          float s2 = sq[0] + __shfl_xor_sync(0xffffffffu, sq[0], 1);
          if ((lane & 1) == 0) {
            s_part[q * 2 * p.block_n + c0 + my_col] = s1;
            s_part[q * 2 * p.block_n + p.block_n + c0 + my_col] = s2;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (p.stats) {
        // the four warps' column sums are combined in a fixed order by the channel's owner thread (no
        // shared-memory atomics: their arrival order would make the fp32 sums differ run to run)
        asm volatile("bar.sync 1, 128;");
        for (int c = et; c < p.block_n; c += 128) {
          const int ch = n0 + c;
          if (ch < p.Cout) {
            const float* sp = s_part + c;
            const int st = 2 * p.block_n;
            s_stats[ch] += (sp[0] + sp[st]) + (sp[2 * st] + sp[3 * st]);
            sp += p.block_n;
            s_stats[p.Cout + ch] += (sp[0] + sp[st]) + (sp[2 * st] + sp[3 * st]);
          }
        }
        asm volatile("bar.sync 1, 128;");
      }
    }
    if (p.stats) {
      // one partial row per CTA (BatchNorm partial-sum protocol, see sc_bn_stats): no global atomics
      asm volatile("bar.sync 1, 128;");
      double* rowp = p.stats + (int64_t)blockIdx.x * 2 * p.Cout;
      for (int i = et; i < 2 * p.Cout; i += 128) rowp[i] = (double)s_stats[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// channel chunk = swizzle span of one TMA box: the widest of 64 / 32 / 16 whose zero padding of the
// ragged last chunk stays small (padding costs MMA cycles only: TMA zero-fills out-of-range channels)
static int pick_kc(int C) {
  // channel chunk = TMA box row = one swizzle row (32 / 64 / 128 bytes).  TMA retires a box row every few cycles
  // whatever its width, and the thin layers are bound by exactly that row rate (per-tap wgrad of 152 -> 64: 46 rows
  // per pixel at 32-channel chunks = 187 us of row issue for a 208 us kernel), so above 32 channels the widest row
  // wins even when it pads K (the tensor pipe idles on these layers anyway).  STARCOP_KC_RULE=old: <= 12.5 % padding.
  static const bool old_rule = getenv("STARCOP_KC_RULE") && getenv("STARCOP_KC_RULE")[0] == 'o';
  if (C <= 16) return 16;
  if (!old_rule) return C <= 32 ? 32 : 64;
  int p64 = (C + 63) / 64 * 64;
  if ((p64 - C) * 8 <= C) return 64;      // <= 12.5 % padding
  return 32;
}
extern "C" int sc_tc_cin_pad(int Cin) {
  int kc = pick_kc(Cin);
  return (Cin + kc - 1) / kc * kc;
}

extern "C" int sc_tc_conv_fprop(const void* x, int ldx, const void* w_bf16, void* y, int ldy, double* stats,
                                int* stats_rows_host, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                                int stride, int accumulate, void* stream) {
  if (!x || !w_bf16 || !y || (stats && !stats_rows_host)) return SC_ERR_BAD_ARG;
  if (stride != 1 && stride != 2) return SC_ERR_UNSUPPORTED;
  const int Ho = stride == 1 ? H : (H + 2 * (KH / 2) - KH) / 2 + 1;
  const int Wo = stride == 1 ? W : (W + 2 * (KW / 2) - KW) / 2 + 1;
  if (Cin < 1 || Cout % 8 || Wo % 16 || Ho % 8 || ldx % 8 || ldy % 8 || KH != KW || (KH != 1 && KH != 3))
    return SC_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return SC_ERR_BAD_ARG;
  const int kc = pick_kc(Cin);
  FpropParams p;
  p.N = N; p.H = Ho; p.W = Wo; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.stride = stride;
  p.tiles_w = Wo / 16; p.tiles_h = Ho / 8;
  p.m_tiles = N * p.tiles_w * p.tiles_h;
  int nt = (Cout + 255) / 256;
  p.block_n = ((Cout + nt - 1) / nt + 15) / 16 * 16;
  p.n_tiles = (Cout + p.block_n - 1) / p.block_n;
  p.div_n = make_fastdiv((uint32_t)p.n_tiles);
  p.div_w = make_fastdiv((uint32_t)p.tiles_w);
  p.div_h = make_fastdiv((uint32_t)p.tiles_h);
  p.cchunks = (Cin + kc - 1) / kc;      // a ragged last chunk is zero-filled by TMA (and by the weight pack)
  p.ldy = ldy; p.y = (__nv_bfloat16*)y; p.stats = stats; p.accumulate = accumulate;
  const int ksteps = KH * KW * p.cchunks;
  const int group_bytes = 128 * kc * 2 + ((p.block_n * kc * 2 + 1023) & ~1023);
  // thin layers issue only a few MMA cycles per (tap, chunk) group: put several groups in one
  // pipeline stage so the MMA thread pays one barrier round trip per >= ~256 tensor-pipe cycles
  const int group_cycles = (kc / 16) * (p.block_n / 2);
  int groups = (256 + group_cycles - 1) / group_cycles;
  if (groups > ksteps) groups = ksteps;
  if (groups > 9) groups = 9;
  while (groups > 1 && groups * group_bytes > 64 * 1024) --groups;
  if (groups < 1) groups = 1;
  p.groups = groups;
  const int b_tile_bytes = (p.block_n * kc * 2 + 1023) & ~1023;
  // resident weights (kernel comment): all (N tile, K step) weight tiles, when they are small next to the pipeline
  static const int kBresMax = getenv("STARCOP_BRES_KB") ? atoi(getenv("STARCOP_BRES_KB")) * 1024 : 48 * 1024;
  const int b_all_bytes = p.n_tiles * ksteps * b_tile_bytes;
  const bool bres = b_all_bytes <= kBresMax && !getenv("STARCOP_NO_BRES");
  // 1024-aligned (swizzle atoms), behind the barrier / statistics block
  const int stats_bytes = stats ? (2 * Cout + 8 * p.block_n) * 4 : 0;
  p.b_resident = bres ? (256 + stats_bytes + 1023) & ~1023 : 0;
  const int tail = 1024 + (bres ? p.b_resident + b_all_bytes : 256 + stats_bytes);   // alignment slack + barriers + statistics (+ weights)
  const int stage_bytes = groups * (bres ? 128 * kc * 2 : group_bytes);               // resident weights: the stages hold activations only
  // Short-K layers (the 1x1 expansions / projections: one or two MMAs per tile) are bound by the per-tile
  // TMA -> MMA -> epilogue hand-offs, not by the tensor pipe: run TWO CTAs per SM (half the shared memory,
  // 256 TMEM columns each) so twice as many tiles are in flight.
  const int total_tiles = p.m_tiles * p.n_tiles;
  bool two_cta = Cin * KH * KW <= 192 && p.block_n <= 128 && total_tiles >= 4 * kNumSMs;
  if (two_cta && (100 * 1024 - tail) / stage_bytes < 3) two_cta = false;     // not enough pipeline depth in half an SM
  p.tmem_cols = two_cta ? 256 : 512;
  int stages = ((two_cta ? 100 : 200) * 1024 - tail) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return SC_ERR_UNSUPPORTED;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + tail;

  CUtensorMap tmA, tmB;
  int rc = encode_act(&tmA, x, Cin, W, H, N, ldx, kc, 16, 8, stride);
  if (rc != SC_OK) return rc;
  rc = encode_mat(&tmB, w_bf16, (int64_t)KH * KW * p.cchunks * kc, Cout, kc, p.block_n);
  if (rc != SC_OK) return rc;
  int total = p.m_tiles * p.n_tiles;
  const int slots = two_cta ? 2 * kNumSMs : kNumSMs;
  int grid = total < slots ? total : slots;
  if (stats) *stats_rows_host = grid;
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_FPROP(KC)                                                                                     \
  do {                                                                                                       \
    /* opt in to the full dynamic shared memory: a per-device function attribute, not a stream operation */  \
    cudaError_t e = cudaFuncSetAttribute(tc_conv_fprop_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         kMaxDynSmem);                                                       \
    if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }                                          \
    sc::launch_pdl((tc_conv_fprop_kernel<KC>), grid, kTcThreads, smem, st, tmA, tmB, p);                                   \
  } while (0)
  if (kc == 64) LAUNCH_FPROP(64);
  else if (kc == 32) LAUNCH_FPROP(32);
  else LAUNCH_FPROP(16);
#undef LAUNCH_FPROP
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------------
struct WgradParams {
  int N, H, W, Cin, Cout, KH, KW;   // H, W: OUTPUT (dY) spatial size
  int stride;
  int tiles_w, tiles_h, p_tiles;   // 4 x 16 pixel patches
  int m_blocks, n_blocks, ksplit;
  int block_n, nb_boxes, kcb;      // dY tile: nb_boxes boxes of kcb channels
  int cchunks, subs_total;         // sub-blocks (tap, channel chunk) of KC rows each
  int stages;
  float* dw;
  float* partials;                 // [m_blocks * n_blocks][ksplit][128][block_n] fp32 (ksplit > 1)
  int red_items, red_groups;       // split reduction: blocks of the gradient x groups of splits (one CTA each)
  float* red_scratch;              // [red_items][red_groups][64 * KH * KW] group sums (red_groups > 1)
  int* red_tickets;                // [red_items] arrival counters, cleared by the wgrad kernel itself
};

// dW += sum over the pixel splits of their partial tiles, in a fixed order (deterministic).  The partial tiles hold the
// gradient as [tap, ci][co] (co fastest), the OIHW gradient wants [co][ci][tap].  The gradient is cut into blocks of
// (all taps) x (8 input channels) x (8 output channels); the splits of a block (up to ~300 for the thin
// high-resolution layers) are cut into `red_groups` contiguous groups and ONE CTA sums one group of one block:
//   * inside the CTA the group's splits are summed in interleaved slices by different threads (split k of the group
//     belongs to slice k % slices; four loads in flight per thread) and the slice sums are added in slice order;
//   * with one group the CTA transposes the block through shared memory and adds it into dW with coalesced runs of
//     8 * KH * KW floats per output channel;
//   * with several groups it stores its group sum, takes a ticket, and the CTA that arrives LAST adds the group sums
//     in group order and writes dW.  Which CTA is last varies, the order of the additions never does.
// Every level has a fixed association, so the result does not depend on scheduling; the grid stays in the hundreds of
// CTAs for every layer (64 -> 64 3x3 at 60 splits: 64 blocks x 8 groups; the 4 -> 32 stem at 296 splits: 4 x 37).
constexpr int kWredT = 8;                        // block edge: 8 input x 8 output channels
constexpr int kWredSliceFloats = 4096;           // scratch for the slice sums (16 KB)
template <int KC>
__global__ void __launch_bounds__(256)
tc_wgrad_reduce_kernel(WgradParams p) {
  sc::pdl_wait();
  constexpr int SUBS = 128 / KC;
  constexpr int CGS = KC / kWredT;               // input-channel groups per chunk
  __shared__ float slice_s[kWredSliceFloats];
  __shared__ float tr_s[kWredT][kWredT * 9 + 1];
  __shared__ int s_last;
  const int KK = p.KH * p.KW;
  const int E = KK * kWredT * kWredT;            // elements of one block: 64 (1x1) or 576 (3x3)
  const int G = p.red_groups;
  const int per_group = (p.ksplit + G - 1) / G;
  const int co_tiles = (p.Cout + kWredT - 1) / kWredT;
  const size_t split_stride = (size_t)128 * p.block_n;
  const int item = blockIdx.x / G, grp = blockIdx.x - item * G;
  const int ct = item % co_tiles;
  const int cgi = item / co_tiles;
  const int chunk = cgi / CGS, cg = cgi - chunk * CGS;
  const int ci0 = chunk * KC + cg * kWredT;
  if (ci0 >= p.Cin) return;                      // zero-padded channels of a ragged last chunk (whole block idle)
  const int co0 = ct * kWredT;
  const int nb = co0 / p.block_n;                // block_n is a multiple of 16: a block never straddles N blocks
  const int c0 = co0 - nb * p.block_n;
  const int k0 = grp * per_group;
  const int k1 = min(p.ksplit, k0 + per_group);
  int slices = kWredSliceFloats / E;
  slices = slices > 16 ? 16 : slices;
  if (slices > k1 - k0) slices = max(k1 - k0, 1);
  for (int idx = threadIdx.x; idx < E * slices; idx += 256) {
    const int sl = idx / E, e = idx - sl * E;
    const int col = e % kWredT;
    const int t = e / kWredT;
    const int cil = cg * kWredT + t % kWredT, tap = t / kWredT;
    const int j = tap * p.cchunks + chunk;
    const int mb = j / SUBS, sidx = j - mb * SUBS;
    const float* base = p.partials + (((size_t)(nb * p.m_blocks + mb) * p.ksplit) * 128 + sidx * KC + cil) * p.block_n + c0 + col;
    float a = 0.f;
    int k = k0 + sl;
    for (; k + 3 * slices < k1; k += 4 * slices) {
      const float x0 = __ldcg(base + (size_t)k * split_stride), x1 = __ldcg(base + (size_t)(k + slices) * split_stride);
      const float x2 = __ldcg(base + (size_t)(k + 2 * slices) * split_stride), x3 = __ldcg(base + (size_t)(k + 3 * slices) * split_stride);
      a += x0; a += x1; a += x2; a += x3;
    }
    for (; k < k1; k += slices) a += __ldcg(base + (size_t)k * split_stride);
    slice_s[idx] = a;
  }
  __syncthreads();
  float* mine = p.red_scratch + ((size_t)item * G + grp) * E;
  for (int e = threadIdx.x; e < E; e += 256) {
    float a = slice_s[e];
    for (int sl = 1; sl < slices; ++sl) a += slice_s[sl * E + e];
    if (G == 1) {
      const int col = e % kWredT;
      const int t = e / kWredT;
      tr_s[col][(t % kWredT) * KK + t / kWredT] = a;
    } else {
      __stcg(mine + e, a);
    }
  }
  if (G > 1) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(p.red_tickets + item, 1) == G - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* all = p.red_scratch + (size_t)item * G * E;
    for (int e = threadIdx.x; e < E; e += 256) {
      float a = __ldcg(all + e);
      int g = 1;
      for (; g + 3 < G; g += 4) {
        const float x0 = __ldcg(all + (size_t)g * E + e), x1 = __ldcg(all + (size_t)(g + 1) * E + e);
        const float x2 = __ldcg(all + (size_t)(g + 2) * E + e), x3 = __ldcg(all + (size_t)(g + 3) * E + e);
        a += x0; a += x1; a += x2; a += x3;
      }
      for (; g < G; ++g) a += __ldcg(all + (size_t)g * E + e);
      const int col = e % kWredT;
      const int t = e / kWredT;
      tr_s[col][(t % kWredT) * KK + t / kWredT] = a;
    }
  }
  __syncthreads();
  const int n = min(kWredT, p.Cin - ci0) * KK;   // valid contiguous floats per output channel
  for (int e = threadIdx.x; e < kWredT * n; e += 256) {
    const int col = e / n, o = e - col * n;
    const int co = co0 + col;
    if (co < p.Cout) p.dw[((int64_t)co * p.Cin + ci0) * KK + o] += tr_s[col][o];
  }
}

template <int KC>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, WgradParams p) {
  sc::pdl_wait();
  if (blockIdx.x == 0 && p.red_groups > 1)       // arrival counters of the split reduction that follows in the stream
    for (int i = threadIdx.x; i < p.red_items; i += kTcThreads) p.red_tickets[i] = 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int PIX = 64;                        // pixels (K) per stage: 4 rows x 16 cols
  constexpr int SUBS = 128 / KC;                 // activation boxes per M block
  constexpr int A_SUB_BYTES = PIX * KC * 2;
  constexpr int A_BYTES = SUBS * A_SUB_BYTES;    // 16 KB
  const int B_BOX_BYTES = PIX * p.kcb * 2;
  const int B_BYTES = p.nb_boxes * B_BOX_BYTES;
  const int STAGE_BYTES = (A_BYTES + B_BYTES + 1023) & ~1023;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * STAGE_BYTES);
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
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item: (m block, n block, pixel split)
  int wid = blockIdx.x;
  const int mb = wid % p.m_blocks;
  wid /= p.m_blocks;
  const int nb = wid % p.n_blocks;
  const int ks = wid / p.n_blocks;
  const int n0 = nb * p.block_n;
  const int per = (p.p_tiles + p.ksplit - 1) / p.ksplit;
  const int t_begin = ks * per;
  const int t_end = min(p.p_tiles, t_begin + per);
  const int nsteps = max(0, t_end - t_begin);
  int valid_subs = p.subs_total - mb * SUBS;
  if (valid_subs > SUBS) valid_subs = SUBS;
  const int pad = p.KH / 2;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // this CTA's sub-blocks (channel offset, tap shift) are fixed: decode them once
      int sc0[SUBS], sdx[SUBS], sdy[SUBS];
#pragma unroll
      for (int s = 0; s < SUBS; ++s) {
        int j = mb * SUBS + s;
        int tap = j / p.cchunks, cc = j - tap * p.cchunks;
        sc0[s] = cc * KC;
        sdy[s] = tap / p.KW - pad;
        sdx[s] = tap % p.KW - pad;
      }
      const uint32_t tx = valid_subs * A_SUB_BYTES + B_BYTES;
      int tw = t_begin % p.tiles_w;
      int t2 = t_begin / p.tiles_w;
      int th = t2 % p.tiles_h;
      int img = t2 / p.tiles_h;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sA = smem + (size_t)stage * STAGE_BYTES;
        uint8_t* sB = sA + A_BYTES;
        mbar_arrive_expect_tx(&full[stage], tx);
        const int wq = (tw * 16) * p.stride, hq = (th * 4) * p.stride;
#pragma unroll
        for (int s = 0; s < SUBS; ++s)
          if (s < valid_subs) tma_load_4d(sA + s * A_SUB_BYTES, &tmX, sc0[s], wq + sdx[s], hq + sdy[s], img, &full[stage]);
        for (int b = 0; b < p.nb_boxes; ++b)
          tma_load_4d(sB + b * B_BOX_BYTES, &tmDY, n0 + b * p.kcb, tw * 16, th * 4, img, &full[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
        if (++tw == p.tiles_w) {
          tw = 0;
          if (++th == p.tiles_h) {
            th = 0;
            ++img;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.block_n, 1, 1);
      constexpr uint32_t A_LAYOUT = layout_for_row_bytes(KC * 2);
      const uint32_t b_layout = layout_for_row_bytes(p.kcb * 2);
      constexpr uint32_t A_SBO = 8 * KC * 2;
      const uint32_t b_sbo = 8 * p.kcb * 2;
      int stage = 0;
      uint32_t phase = 0;
      // descriptors assembled once; a stage / K step is an add on the low word (start address, 16-byte units)
      const uint64_t ad0 = make_smem_desc(smem_u32(smem), A_SUB_BYTES, A_SBO, A_LAYOUT);
      const uint64_t bd0 = make_smem_desc(smem_u32(smem) + A_BYTES, B_BOX_BYTES, b_sbo, b_layout);
      const uint64_t a_hi = ad0 & 0xffffffff00000000ull, b_hi = bd0 & 0xffffffff00000000ull;
      const uint32_t a_lo0 = (uint32_t)ad0, b_lo0 = (uint32_t)bd0;
      const uint32_t b_kstep = (uint32_t)(16 * p.kcb * 2) >> 4;
      for (int i = 0; i < nsteps; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t soff = (uint32_t)stage * (uint32_t)(STAGE_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < PIX / 16; ++k) {
          // K advances over pixels: 16 rows of (kc*2) bytes
          const uint64_t ad = a_hi | (uint64_t)(a_lo0 + soff + (uint32_t)(k * 16 * KC * 2 >> 4));
          const uint64_t bd = b_hi | (uint64_t)(b_lo0 + soff + (uint32_t)k * b_kstep);
          umma_bf16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tfull);
    }
  } else {
    // epilogue (every split holds >= 1 pixel tile: the host sizes ksplit that way)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int KK = p.KH * p.KW;
    // OIHW offset of accumulator row r (tap, ci) of this M block, or -1 for a padding row
    auto row_offset = [&](int r) -> int64_t {
      const int s = r / KC, cil = r % KC;
      if (s >= valid_subs) return -1;
      const int j = mb * SUBS + s;
      const int tap = j / p.cchunks;
      const int ci = (j - tap * p.cchunks) * KC + cil;
      return ci < p.Cin ? (int64_t)ci * KK + tap : -1;
    };
    mbar_wait(tfull, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (p.ksplit == 1) {
      const int64_t ro = row_offset(row);
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        if (ro >= 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int co = n0 + c0 + i;
            if (co < p.Cout) p.dw[(int64_t)co * p.Cin * KK + ro] += v[i];     // this CTA is the element's only writer
          }
        }
      }
    } else {
      // this split's partial tile; tc_wgrad_reduce_kernel sums the splits in order
      const int tile_id = nb * p.m_blocks + mb;
      const int tile_elems = 128 * p.block_n;
      float* mine = p.partials + ((size_t)tile_id * p.ksplit + ks) * tile_elems + (size_t)row * p.block_n;
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(mine + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// geometry shared by the launcher and the workspace query
static int wgrad_geometry(int N, int Ho, int Wo, int Cin, int Cout, int KH, int KW, WgradParams* p, int* kc_out) {
  const int kc = pick_kc(Cin);
  p->tiles_w = Wo / 16; p->tiles_h = Ho / 4;
  p->p_tiles = N * p->tiles_w * p->tiles_h;
  p->kcb = pick_kc(Cout);
  int nt = (Cout + 255) / 256;
  p->block_n = ((Cout + nt - 1) / nt + p->kcb - 1) / p->kcb * p->kcb;
  if (p->block_n > 256) return SC_ERR_UNSUPPORTED;
  p->n_blocks = (Cout + p->block_n - 1) / p->block_n;
  p->nb_boxes = p->block_n / p->kcb;
  p->cchunks = (Cin + kc - 1) / kc;
  p->subs_total = KH * KW * p->cchunks;
  const int subs = 128 / kc;
  p->m_blocks = (p->subs_total + subs - 1) / subs;
  int base = p->m_blocks * p->n_blocks;
  int ksplit = (kNumSMs * 2 + base - 1) / base;
  int max_split = (p->p_tiles + 7) / 8;                    // >= 8 pixel tiles (512 px) per CTA
  if (ksplit > max_split) ksplit = max_split;
  if (ksplit < 1) ksplit = 1;
  // every split must own at least one pixel tile (its partial tile enters the ordered sum)
  const int per = (p->p_tiles + ksplit - 1) / ksplit;
  p->ksplit = (p->p_tiles + per - 1) / per;
  // split reduction: one CTA per (8 x 8 channel block, group of splits); aim at ~4 CTAs per SM, >= 4 splits per group
  p->red_items = p->cchunks * (kc / kWredT) * ((Cout + kWredT - 1) / kWredT);
  int groups = (4 * kNumSMs + p->red_items - 1) / p->red_items;
  if (groups > (p->ksplit + 3) / 4) groups = (p->ksplit + 3) / 4;
  if (groups > 32) groups = 32;
  if (groups < 1) groups = 1;
  const int per_group = (p->ksplit + groups - 1) / groups;
  p->red_groups = (p->ksplit + per_group - 1) / per_group;       // no empty group
  *kc_out = kc;
  return SC_OK;
}
// workspace layout: [partial tiles][group sums][tickets]
static size_t wgrad_partials_bytes(const WgradParams& p) {
  return ((size_t)p.m_blocks * p.n_blocks * p.ksplit * 128 * p.block_n * sizeof(float) + 255) & ~(size_t)255;
}
static size_t wgrad_scratch_bytes(const WgradParams& p) {
  return p.red_groups > 1 ? ((size_t)p.red_items * p.red_groups * 64 * p.KH * p.KW * sizeof(float) + 255) & ~(size_t)255 : 0;
}

extern "C" int64_t sc_tc_conv_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride) {
  const int Ho = stride == 1 ? H : (H + 2 * (KH / 2) - KH) / 2 + 1;
  const int Wo = stride == 1 ? W : (W + 2 * (KW / 2) - KW) / 2 + 1;
  if (KH == 3 && KW == 3 && stride == 1 && sc_tc_wgrad_halo_supported(N, H, W, Cin, Cout))
    return sc_tc_wgrad_halo_workspace_bytes(N, H, W, Cin, Cout);
  WgradParams p;
  int kc;
  if (Cin < 1 || wgrad_geometry(N, Ho, Wo, Cin, Cout, KH, KW, &p, &kc) != SC_OK) return -1;
  p.KH = KH; p.KW = KW;
  return p.ksplit > 1 ? (int64_t)(wgrad_partials_bytes(p) + wgrad_scratch_bytes(p) + (size_t)p.red_items * sizeof(int)) : 0;
}

extern "C" int sc_tc_conv_wgrad(const void* x, int ldx, const void* dy, int lddy, float* dw_oihw, float* partials,
                                int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, void* stream) {
  if (!x || !dy || !dw_oihw) return SC_ERR_BAD_ARG;
  if (stride != 1 && stride != 2) return SC_ERR_UNSUPPORTED;
  const int Ho = stride == 1 ? H : (H + 2 * (KH / 2) - KH) / 2 + 1;
  const int Wo = stride == 1 ? W : (W + 2 * (KW / 2) - KW) / 2 + 1;
  if (Cin < 1 || Cout % 8 || Wo % 16 || Ho % 4 || ldx % 8 || lddy % 8 || KH != KW || (KH != 1 && KH != 3))
    return SC_ERR_UNSUPPORTED;
  if (KH == 3 && stride == 1 && sc_tc_wgrad_halo_supported(N, H, W, Cin, Cout))      // thin layers: halo-patch kernel
    return sc_tc_wgrad_halo(x, ldx, dy, lddy, dw_oihw, partials, N, H, W, Cin, Cout, stream);
  WgradParams p;
  int kc;
  p.N = N; p.H = Ho; p.W = Wo; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.stride = stride;
  int rc = wgrad_geometry(N, Ho, Wo, Cin, Cout, KH, KW, &p, &kc);
  if (rc != SC_OK) return rc;
  if (p.ksplit > 1 && !partials) return SC_ERR_BAD_ARG;
  p.dw = dw_oihw;
  p.partials = partials;
  p.red_scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(partials) + wgrad_partials_bytes(p));
  p.red_tickets = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(p.red_scratch) + wgrad_scratch_bytes(p));
  const int stage_bytes = (64 * 128 * 2 + p.nb_boxes * 64 * p.kcb * 2 + 1023) & ~1023;
  const int tail = 1024 + 256;
  int stages = (200 * 1024 - tail) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) return SC_ERR_UNSUPPORTED;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + tail;

  CUtensorMap tmX, tmDY;
  rc = encode_act(&tmX, x, Cin, W, H, N, ldx, kc, 16, 4, stride);
  if (rc != SC_OK) return rc;
  rc = encode_act(&tmDY, dy, Cout, Wo, Ho, N, lddy, p.kcb, 16, 4);
  if (rc != SC_OK) return rc;
  int grid = p.m_blocks * p.n_blocks * p.ksplit;
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_WGRAD(KC)                                                                                     \
  do {                                                                                                       \
    cudaError_t e = cudaFuncSetAttribute(tc_conv_wgrad_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         kMaxDynSmem);                                                       \
    if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }                                          \
    sc::launch_pdl((tc_conv_wgrad_kernel<KC>), grid, kTcThreads, smem, st, tmX, tmDY, p);                                  \
    if (p.ksplit > 1) {                                                                                      \
      sc::launch_pdl((tc_wgrad_reduce_kernel<KC>), p.red_items * p.red_groups, 256, 0, st, p);              \
    }                                                                                                        \
  } while (0)
  if (kc == 64) LAUNCH_WGRAD(64);
  else if (kc == 32) LAUNCH_WGRAD(32);
  else LAUNCH_WGRAD(16);
#undef LAUNCH_WGRAD
  return check_launch();
}
