// tcgen05 3x3 / stride-1 / pad-1 convolution for the THIN high-resolution layers of the U-Net decoder
// (decoder blocks 2-4: 16...80 channels at 128^2...512^2), fprop and dgrad (dgrad = the same kernel on dy
// with the flipped / transposed filter).
//
// The general kernel (conv_tc.cu) fetches one shifted TMA box per filter tap: every input pixel crosses
// L2 -> SM nine times, and at these channel counts that L2 traffic, not the tensor pipe or HBM, is the
// limit (b4c1 fprop: 9 x 268 MB in 359 us = 6.7 TB/s of L2 reads).  Here the input HALO PATCH of a
// 16 x 8 pixel output tile (18 x 10 pixels) is staged ONCE, and the nine taps are nine UMMA shared-memory
// descriptors into that one patch:
//   * the patch is stored as channel planes [Cin/8][18][10] of 16-byte granules (8 bf16 channels), i.e.
//     the canonical un-swizzled K-major UMMA layout: a core matrix (8 rows x 16 B) = 8 consecutive
//     pixels of a patch row, SBO = one patch row (160 B), LBO = one plane;
//   * with M = 128 = 16 rows x 8 pixels, A-row m = (r, q) of tap (dy, dx) is patch pixel (r+dy, q+dx):
//     the tap is just a start-address offset of (dy*10 + dx)*16 bytes;
//   * the patch is gathered by ONE producer warp with 16-byte cp.async (LDGSTS, zero fill outside the
//     image = the padding, and for the ragged last channel pair): consecutive lanes fetch the consecutive
//     channel granules of a pixel, so global reads are fully coalesced 32...160-byte runs, and the
//     granule -> (source offset, destination) map is a small table built once per CTA.  (A TMA box per
//     plane was measured first: TMA retires one box ROW per ~3-6 cycles whatever its width, so 16-byte
//     rows make the producer the limit at 16 channels.)  Each lane's share of a patch is published
//     by cp.async.mbarrier.arrive.noinc when its copies land; the MMA thread orders the generic-proxy
//     writes before the tensor core's async-proxy reads with fence.proxy.async after the barrier wait;
//   * tiles are processed in batches of up to four whose MMAs are interleaved tap by tap (independent
//     TMEM accumulators): one tile's K loop is a dependent chain that leaves the pipe idle at N <= 80;
//   * ALL taps' weights stay resident in shared memory for the CTA's lifetime ([K/8][Cout] granules).
// Warp roles / TMEM double buffering / BatchNorm statistics epilogue as in conv_tc.cu.
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

using namespace sc;
using namespace tc;

namespace {

constexpr int kThreads = 256;                         // warps 0-2 gather patches, warp 3 issues the MMAs, warps 4-7 = epilogue
constexpr int kProducers = 3;
constexpr int kTileH = 16, kTileW = 8;                 // output tile: M = 128 pixels
constexpr int kPatchH = kTileH + 2, kPatchW = kTileW + 2;
constexpr int kPlaneBytes = kPatchH * kPatchW * 16;    // 2880
constexpr int kPlaneStride = (kPlaneBytes + 127) & ~127;   // 2944: every TMA destination 128 B aligned
constexpr int kRowBytes = kPatchW * 16;                // SBO
constexpr int kMaxSmem = 227 * 1024;
constexpr int kMaxBatch = 4;                           // tiles whose MMAs / epilogues are interleaved

struct HaloParams {
  int N, H, W, Cin, Cout;
  int np;                      // channel planes per patch = cin_pad16 / 8
  int tiles_w, tiles_h, m_tiles;
  FastDiv div_w, div_h;        // tile index -> (image, tile row, tile column) without integer divisions
  int stages, ldy, accumulate;
  int batch, ncols;            // tiles per MMA batch, TMEM columns per accumulator
  int tmem_cols;               // allocation: power of two >= 2 * batch * ncols
  __nv_bfloat16* y;
  const __nv_bfloat16* w;      // [Cout][9][np*8]
  const __nv_bfloat16* x;
  int ldx;
  double* stats;
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on the mbarrier once all of this thread's earlier cp.async copies have completed (no pending-count
// increment: the barrier is initialised with one arrival per producer lane)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kThreads, 2)
tc_conv3x3_halo_kernel(HaloParams p) {
  sc::pdl_wait();
  extern __shared__ __align__(1024) uint8_t smem[];
  const int np = p.np;
  const int B_BYTES = 9 * np * p.Cout * 16;                         // [tap*np + plane][Cout] granules
  const int B_AL = (B_BYTES + 127) & ~127;
  const int STAGE_BYTES = np * kPlaneStride;
  uint8_t* sB = smem;
  uint8_t* sA = smem + B_AL;
  uint64_t* full = reinterpret_cast<uint64_t*>(sA + (size_t)p.stages * STAGE_BYTES);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_stats = reinterpret_cast<float*>(tmem_slot + 4);         // [2][Cout]
  float* s_part = s_stats + 2 * p.Cout;                             // [4 warps][2][Cout] one batch's column sums
  int2* tab = reinterpret_cast<int2*>(s_part + 8 * p.Cout);         // [np*180] granule map (8 B aligned: Cout % 16 == 0)
  const int ngran = np * kPatchH * kPatchW;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 32 * kProducers);      // every producer lane publishes its own share of a patch
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
    fence_barrier_init();
  }
  // granule g = pixel * np + plane (consecutive lanes = consecutive 16-byte runs of one pixel's channels):
  // .x = element offset from the patch origin, .y = destination offset | ph << 15 | pw << 20 | channel-valid << 24
  for (int g = threadIdx.x; g < ngran; g += kThreads) {
    const int px = g / np, pl = g - px * np;
    const int ph = px / kPatchW, pw = px - ph * kPatchW;
    tab[g] = make_int2((ph * p.W + pw) * p.ldx + pl * 8,
                       (pl * kPlaneStride + px * 16) | (ph << 15) | (pw << 20) | ((pl * 8 < p.Cin ? 1 : 0) << 24));
  }
  if (warp == kProducers) tmem_alloc(tmem_slot, p.tmem_cols);
  // resident weights: global [n][k8] granules -> shared [k8][n] granules (un-swizzled K-major B operand)
  {
    const int k8n = 9 * np;
    const uint4* wg = reinterpret_cast<const uint4*>(p.w);
    uint4* ws = reinterpret_cast<uint4*>(sB);
    for (int i = threadIdx.x; i < k8n * p.Cout; i += kThreads) {
      const int k8 = i / p.Cout, n = i - k8 * p.Cout;
      ws[i] = wg[(size_t)n * k8n + k8];
    }
    fence_proxy_async();             // generic-proxy writes -> visible to the tensor core's async proxy
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Register budget: 2 CTAs x 256 threads leave 128 registers per thread at launch; the gather / MMA warps need ~50,
  // the epilogue (a batch of four tiles' accumulators in flight + their statistics) ~170: the first warpgroup hands
  // its surplus to the second one.
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp < kProducers) {
    // producers: THREE warps gather a patch with cp.async (ncu, one producer warp: that warp was busy 100 % of the
    // time at ~200 cycles per LDGSTS -- the copies a single warp may have in flight cap the gather, not its issue
    // rate -- while the epilogue warps idled 42 % waiting for accumulators); each lane's share is published by the hardware
    // (cp.async.mbarrier.arrive.noinc fires when the lane's copies have landed), so up to `stages` patches
    // are in flight and the producer never blocks on its own loads.  (Publishing with wait_group ->
    // fence.proxy.async -> arrive in the producer was measured 1.3-1.45x slower: the blocking wait caps the
    // loads in flight.)
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t sA_u32 = smem_u32(sA);
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      const int t2 = (int)fast_div((uint32_t)tile, p.div_w);
      const int tw = tile - t2 * p.tiles_w;
      const int img = (int)fast_div((uint32_t)t2, p.div_h);
      const int th = t2 - img * p.tiles_h;
      const int h0 = th * kTileH - 1, w0 = tw * kTileW - 1;
      mbar_wait(&empty[stage], phase ^ 1);
      const __nv_bfloat16* base = p.x + (((int64_t)img * p.H + h0) * p.W + w0) * p.ldx;
      const uint32_t dst = sA_u32 + (uint32_t)stage * STAGE_BYTES;
      if (h0 >= 0 && w0 >= 0 && h0 + kPatchH <= p.H && w0 + kPatchW <= p.W) {
        // interior patch (all but the image border): no per-pixel bounds checks
#pragma unroll 4
        for (int g = warp * 32 + lane; g < ngran; g += 32 * kProducers) {
          const int2 t = tab[g];
          cp_async16_zfill(dst + (t.y & 0x7fff), base + t.x, (t.y >> 24) ? 16u : 0u);
        }
      } else {
#pragma unroll 2
        for (int g = warp * 32 + lane; g < ngran; g += 32 * kProducers) {
          const int2 t = tab[g];
          const int ph = (t.y >> 15) & 31, pw = (t.y >> 20) & 15;
          const bool ok = (unsigned)(h0 + ph) < (unsigned)p.H && (unsigned)(w0 + pw) < (unsigned)p.W && (t.y >> 24);
          cp_async16_zfill(dst + (t.y & 0x7fff), ok ? (const void*)(base + t.x) : (const void*)p.x, ok ? 16u : 0u);
        }
      }
      cp_async_arrive_noinc(&full[stage]);
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == kProducers) {
    if (lane == 0) {
      // Tiles are processed in BATCHES of G with the MMAs of the batch interleaved tap by tap: the K steps of
      // one tile form a dependent accumulation chain, and at N = 16...80 a single chain leaves the tensor pipe
      // idle for most of each MMA's latency (measured ~130-250 cycles per MMA); G independent chains hide it.
      const int G = p.batch, NP = p.ncols;
      const uint32_t idesc = make_idesc_bf16(128, p.Cout, 0, 0);
      const uint32_t b_lbo = p.Cout * 16;
      const uint32_t baddr = smem_u32(sB);
      const uint32_t sA_u32 = smem_u32(sA);
      const int kpairs = np / 2;                 // K = 16 per MMA = two planes
      const int my_tiles = (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t0 = 0; t0 < my_tiles; t0 += G, ++it) {
        const int gn = my_tiles - t0 < G ? my_tiles - t0 : G;
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        // The issuing thread is the bottleneck of this kernel (ncu: ~250 cycles per MMA with ~45 dependent
        // instructions each -- descriptor assembly, local-memory batch tables, per-MMA elect / uniform moves), so the
        // descriptors are assembled ONCE per batch and a tap / K step / tile is an integer add on their low words
        // (start address field, 16-byte units: shared memory is < 256 KB, the field never carries).
        uint32_t a_lo[kMaxBatch];
        int st_first = stage;
#pragma unroll
        for (int g = 0; g < kMaxBatch; ++g) {
          a_lo[g] = 0;
          if (g < gn) {
            mbar_wait(&full[stage], phase);
            a_lo[g] = (uint32_t)make_smem_desc(sA_u32 + (uint32_t)stage * STAGE_BYTES, kPlaneStride, kRowBytes, 0);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        tc_fence_after();
        // the patches were written through the generic proxy (cp.async); UMMA reads through the async proxy
        // (measured: the fence costs < 2 % here)
        fence_proxy_async();
        const uint32_t d_tmem = tmem_base + acc * (G * NP);
        const uint64_t a_hi = make_smem_desc(0, kPlaneStride, kRowBytes, 0) & 0xffffffff00000000ull;
        const uint64_t b_desc0 = make_smem_desc(baddr, b_lbo, 128, 0);
        const uint32_t b_hi32 = (uint32_t)(b_desc0 >> 32), b_lo0 = (uint32_t)b_desc0;
        const uint32_t b_step = b_lbo >> 4;
        auto issue = [&](auto gn_c) {
          constexpr int GN = decltype(gn_c)::value;
          uint32_t first = 0;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a_off = (tap / 3) * kPatchW + (tap % 3);                   // 16-byte granules
            uint32_t bl = b_lo0 + (uint32_t)(tap * np) * b_step;
            uint32_t ao = a_off;
            for (int kk = 0; kk < kpairs; ++kk, bl += 2 * b_step, ao += 2 * (kPlaneStride >> 4)) {
              const uint64_t bd = ((uint64_t)b_hi32 << 32) | bl;
#pragma unroll
              for (int g = 0; g < GN; ++g) umma_bf16(d_tmem + g * NP, a_hi | (uint64_t)(a_lo[g] + ao), bd, idesc, first);
              first = 1;
            }
          }
        };
        switch (gn) {
          case 1: issue(std::integral_constant<int, 1>{}); break;
          case 2: issue(std::integral_constant<int, 2>{}); break;
          case 3: issue(std::integral_constant<int, 3>{}); break;
          default: issue(std::integral_constant<int, 4>{}); break;
        }
        for (int g = 0, st = st_first; g < gn; ++g) {
          umma_commit(&empty[st]);
          if (++st == p.stages) st = 0;
        }
        umma_commit(&tfull[acc]);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    // epilogue: warp q reads TMEM lanes [32q, 32q+32); lane m = pixel (m / 8, m % 8) of the 16 x 8 tile
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;      // 0..127
    if (p.stats) {
      for (int i = et; i < 2 * p.Cout; i += 128) s_stats[i] = 0.f;
      asm volatile("bar.sync 1, 128;");
    }
    const int my_col = lane >> 1;
    const int G = p.batch, NP = p.ncols;
    const int my_tiles = (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const bool wide_ok = !p.accumulate && (p.ldy % 16) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0;
    int it = 0;
    for (int t0 = 0; t0 < my_tiles; t0 += G, ++it) {
      const int gn = my_tiles - t0 < G ? my_tiles - t0 : G;
      const int acc = it & 1;
      // this thread's pixel in each tile of the batch
      __nv_bfloat16* yp[kMaxBatch];
      bool valid[kMaxBatch];
#pragma unroll
      for (int g = 0; g < kMaxBatch; ++g) {
        valid[g] = false;
        yp[g] = p.y;
        if (g < gn) {
          const int tile = blockIdx.x + (t0 + g) * gridDim.x;
          const int t2 = (int)fast_div((uint32_t)tile, p.div_w);
          const int tw = tile - t2 * p.tiles_w;
          const int img = (int)fast_div((uint32_t)t2, p.div_h);
          const int th = t2 - img * p.tiles_h;
          const int h = th * kTileH + (row >> 3), w = tw * kTileW + (row & 7);
          valid[g] = h < p.H && w < p.W;
          yp[g] = p.y + (((int64_t)img * p.H + h) * p.W + w) * p.ldy;
        }
      }
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (G * NP);
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        // all of the batch's TMEM loads in flight together, one wait
        uint32_t r[kMaxBatch][16];
#pragma unroll
        for (int g = 0; g < kMaxBatch; ++g)
          if (g < gn) tmem_ld16_nowait(taddr + g * NP + c0, r[g]);
        tmem_ld_wait();
        float sv[16], sq[16];               // this thread's statistics over the batch: ONE shuffle reduction per batch
#pragma unroll
        for (int i = 0; i < 16; ++i) sv[i] = sq[i] = 0.f;
#pragma unroll
        for (int g = 0; g < kMaxBatch; ++g) {
          if (g < gn && valid[g] && wide_ok) {
            // 16 channels = one 32-byte sector: a single 256-bit store
            uint32_t w8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[g][2 * i]), __uint_as_float(r[g][2 * i + 1]));
              w8[i] = *reinterpret_cast<uint32_t*>(&h2);
              if (p.stats) {
                const float2 f = __bfloat1622float2(h2);
                sv[2 * i] += f.x;
                sv[2 * i + 1] += f.y;
                sq[2 * i] = fmaf(f.x, f.x, sq[2 * i]);
                sq[2 * i + 1] = fmaf(f.y, f.y, sq[2 * i + 1]);
              }
            }
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(yp[g] + c0), "r"(w8[0]), "r"(w8[1]),
                         "r"(w8[2]), "r"(w8[3]), "r"(w8[4]), "r"(w8[5]), "r"(w8[6]), "r"(w8[7])
                         : "memory");
          } else if (g < gn && valid[g]) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              float vv[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) vv[i] = __uint_as_float(r[g][hf * 8 + i]);
              uint4 u;
              __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
              if (p.accumulate) {
                const uint4 o = *reinterpret_cast<const uint4*>(yp[g] + c0 + hf * 8);
                const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&o);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = __bfloat1622float2(oh[i]);
                  vv[2 * i] += f.x;
                  vv[2 * i + 1] += f.y;
                }
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) hh[i] = __floats2bfloat162_rn(vv[2 * i], vv[2 * i + 1]);
              *reinterpret_cast<uint4*>(yp[g] + c0 + hf * 8) = u;
              if (p.stats) {
                // statistics are taken of the STORED (bf16-rounded) values
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = __bfloat1622float2(hh[i]);
                  sv[hf * 8 + 2 * i] += f.x;
                  sv[hf * 8 + 2 * i + 1] += f.y;
                  sq[hf * 8 + 2 * i] = fmaf(f.x, f.x, sq[hf * 8 + 2 * i]);
                  sq[hf * 8 + 2 * i + 1] = fmaf(f.y, f.y, sq[hf * 8 + 2 * i + 1]);
                }
              }
            }
          }
        }
        if (p.stats) {
          // column sums over the warp's 32 pixel rows by recursive halving: 16+16 shuffles per 16 columns
#pragma unroll
          for (int wdt = 8, mask = 16; wdt >= 1; wdt >>= 1, mask >>= 1) {
            const bool upper = (lane & mask) != 0;
#pragma unroll
            for (int i = 0; i < wdt; ++i) {
              const float send = upper ? sv[i] : sv[i + wdt];
              const float keep = upper ? sv[i + wdt] : sv[i];
              sv[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
              const float send2 = upper ? sq[i] : sq[i + wdt];
              const float keep2 = upper ? sq[i + wdt] : sq[i];
              sq[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, mask);
            }
          }
          const float s1 = sv[0] + __shfl_xor_sync(0xffffffffu, sv[0], 1);
          const float s2 = sq[0] + __shfl_xor_sync(0xffffffffu, sq[0], 1);
          if ((lane & 1) == 0) {
            s_part[q * 2 * p.Cout + c0 + my_col] = s1;
            s_part[q * 2 * p.Cout + p.Cout + c0 + my_col] = s2;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (p.stats) {
        // fixed-order combination of the four warps' column sums by the channel's owner thread (deterministic)
        asm volatile("bar.sync 1, 128;");
        for (int c = et; c < 2 * p.Cout; c += 128) {
          const float* sp = s_part + c;
          const int st = 2 * p.Cout;
          s_stats[c] += (sp[0] + sp[st]) + (sp[2 * st] + sp[3 * st]);
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
  if (warp == kProducers) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static size_t halo_smem(int np, int Cout, int stages, bool stats) {
  const size_t b = ((size_t)9 * np * Cout * 16 + 127) & ~(size_t)127;
  return b + (size_t)stages * np * kPlaneStride + (size_t)(2 * stages + 4) * 8 + 16 + 10 * Cout * 4 + (size_t)np * kPatchH * kPatchW * 8 + 64;
}

}  // namespace

extern "C" int sc_tc_halo_cin_pad(int Cin) { return (Cin + 15) / 16 * 16; }

// 1 when the resident-weight halo kernel applies: thin layers only (all 9 taps' weights + >= 3 patches in smem)
extern "C" int sc_tc_halo_supported(int Cin, int Cout) {
  if (Cin < 8 || Cin % 8 || Cout % 16 || Cout < 16 || Cout > 128) return 0;
  const int np = sc_tc_halo_cin_pad(Cin) / 8;
  return halo_smem(np, Cout, 3, true) <= (size_t)kMaxSmem - 1024 ? 1 : 0;
}

extern "C" int sc_tc_conv3x3_halo(const void* x, int ldx, const void* w_bf16, void* y, int ldy, double* stats,
                                  int* stats_rows_host, int N, int H, int W, int Cin, int Cout, int accumulate,
                                  void* stream) {
  if (!x || !w_bf16 || !y || (stats && !stats_rows_host) || N <= 0 || H <= 0 || W <= 0) return SC_ERR_BAD_ARG;
  if (!sc_tc_halo_supported(Cin, Cout) || ldx % 8 || ldy % 8) return SC_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) ||
      (reinterpret_cast<uintptr_t>(w_bf16) & 15))
    return SC_ERR_BAD_ARG;
  HaloParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.np = sc_tc_halo_cin_pad(Cin) / 8;
  p.tiles_w = (W + kTileW - 1) / kTileW;
  p.tiles_h = (H + kTileH - 1) / kTileH;
  p.div_w = make_fastdiv((uint32_t)p.tiles_w);
  p.div_h = make_fastdiv((uint32_t)p.tiles_h);
  const int64_t mt = (int64_t)N * p.tiles_w * p.tiles_h;
  if (mt > INT32_MAX) return SC_ERR_BAD_ARG;
  p.m_tiles = (int)mt;
  p.ldy = ldy; p.accumulate = accumulate;
  p.y = (__nv_bfloat16*)y; p.w = (const __nv_bfloat16*)w_bf16; p.stats = stats;
  // Two CTAs per SM (each with half the shared memory and <= 256 TMEM columns) double the warps that hide the
  // per-tile handshake latencies; layers whose resident weights are too big for that run one CTA per SM.
  p.ncols = (Cout + 31) / 32 * 32;
  int best_ctas = 0, best_stages = 0, best_batch = 0;
  for (int ctas = 2; ctas >= 1; --ctas) {
    const size_t budget = (size_t)kMaxSmem / ctas - 1024;
    const int tmem_budget = 512 / ctas;
    int G = tmem_budget / (2 * p.ncols);
    if (G > kMaxBatch) G = kMaxBatch;
    if (G < 1) continue;
    int stages = 8;
    while (stages > 3 && halo_smem(p.np, Cout, stages, true) > budget) --stages;
    if (halo_smem(p.np, Cout, stages, true) > budget) continue;
    if (G > stages / 2) G = stages / 2;      // the next batch's patches load while this one computes
    if (ctas * G > best_ctas * best_batch) {
      best_ctas = ctas;
      best_stages = stages;
      best_batch = G;
    }
  }
  if (!best_ctas) return SC_ERR_UNSUPPORTED;
  const int stages = best_stages;
  p.stages = stages;
  p.batch = best_batch;
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.batch * p.ncols) p.tmem_cols *= 2;
  const size_t smem = halo_smem(p.np, Cout, stages, stats != nullptr);
  p.x = (const __nv_bfloat16*)x;
  p.ldx = ldx;
  if ((int64_t)(kPatchH * W + kPatchW) * ldx + Cin > INT32_MAX) return SC_ERR_BAD_ARG;
  const int grid = p.m_tiles < kNumSMs * best_ctas ? p.m_tiles : kNumSMs * best_ctas;
  if (stats) *stats_rows_host = grid;
  {
    cudaError_t e = cudaFuncSetAttribute(tc_conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 1024);
    if (e != cudaSuccess) {
      g_last_error = e;
      return SC_ERR_CUDA;
    }
  }
  sc::launch_pdl((tc_conv3x3_halo_kernel), grid, kThreads, smem, (cudaStream_t)stream, p);
  return check_launch();
}
