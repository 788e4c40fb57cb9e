// Per-pixel spectral products over BIP hyperspectral cubes (the layout process_aviris.py:183-184 opens): every
// pixel's spectrum is read from HBM exactly once and a handful of values per pixel are written.  These kernels are
// bandwidth-bound by construction -- algorithmic bytes = pixels x (C_in + C_out) x 4 -- and are the
// "spectral-product kernel" whose achieved HBM GB/s bench.py reports against the measured peak.
//
// sc_srf_aggregate: sensor simulation by spectral response functions (starcop/data/aviris.py:262-338, the product at
// :324-326): out[k] = sum_c w[k][c] * x[c] over the bands the SRF of output band k touches, fill where any of those
// bands is nodata.  A CTA streams chunks of pixels (a multiple of 4, so that every chunk is a 16-byte aligned,
// 16-byte multiple run of the flat cube) through a ring of 1-D bulk copies (cp.async.bulk + mbarrier, three chunks
// in flight per CTA), so the bytes in flight live in shared memory, not in registers; thread = (pixel of the chunk,
// group of output bands) reads its pixel's spectrum from shared memory (pixel pitch = C floats: odd pitches are
// conflict-free) against weights staged once per CTA.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

using namespace sc;

namespace {

constexpr int kSrfThreads = 256;
#ifndef SRF_STAGES
#define SRF_STAGES 3
#endif
constexpr int kSrfStages = SRF_STAGES;
constexpr int kSrfMaxK = 16;

struct SrfRanges {
  int c0[kSrfMaxK], c1[kSrfMaxK];     // band range [c0, c1) with non-zero weight per output band
};

__global__ void __launch_bounds__(kSrfThreads, 4)
srf_aggregate_kernel(const float* __restrict__ cube, int mis, int64_t n_pixels, int64_t tile_pixels, int C, int ppc,
                     const float* __restrict__ w, SrfRanges rg, int K, float fill, float* __restrict__ out, int dbg) {
  // cube is 4-byte aligned; `mis` floats precede it back to the previous 16-byte boundary (they belong to the same
  // allocation: device allocations are 256-byte aligned).  Every chunk is fetched from that boundary: chunk c
  // covers floats [c*ppc*C - mis, ...) of the cube, and a pixel's spectrum sits `mis` floats into its slot.
  extern __shared__ __align__(128) uint8_t ssm[];
  const uint32_t chunk_bytes = (uint32_t)ppc * C * 4 + 16;
  const uint32_t chunk_pitch = (chunk_bytes + 127) & ~127u;
  float* bufs = reinterpret_cast<float*>(ssm);
  uint64_t* full = reinterpret_cast<uint64_t*>(ssm + (size_t)kSrfStages * chunk_pitch);
  float* sw = reinterpret_cast<float*>(full + kSrfStages + (kSrfStages & 1));        // [K][C] weights
  const int tid = threadIdx.x;
  const int64_t nchunks = (n_pixels + ppc - 1) / ppc;
  const int64_t my_n = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int i = 0; i < kSrfStages; ++i) tc::mbar_init(&full[i], 1);
    tc::fence_barrier_init();
  }
  const int Cp = (C + 3) & ~3;                         // weight rows padded to 16 bytes (float4 loads)
  for (int i = tid; i < K * Cp; i += kSrfThreads) {
    const int k = i / Cp, c = i - k * Cp;
    sw[i] = c < C ? w[k * C + c] : 0.f;
  }
  __syncthreads();
  auto issue = [&](int64_t i) {
    const int64_t p0 = (blockIdx.x + i * gridDim.x) * (int64_t)ppc;
    const int64_t np = n_pixels - p0 < ppc ? n_pixels - p0 : ppc;
    uint32_t bytes = (uint32_t)((np * C + mis) * 4);
    bytes &= ~15u;                                     // a tail of < 16 bytes (cube size not a 16-byte multiple) is
    const int slot = (int)(i % kSrfStages);            //   read straight from global memory below
    tc::mbar_arrive_expect_tx(&full[slot], bytes);
    tc::bulk_load_1d(reinterpret_cast<uint8_t*>(bufs) + (size_t)slot * chunk_pitch, cube + p0 * C - mis, bytes, &full[slot]);
  };
  if (tid == 0)
    for (int i = 0; i < kSrfStages && i < my_n; ++i) issue(i);
  const int ngroups = kSrfThreads / ppc;               // >= 1: the launcher keeps ppc <= 256
  const int px = tid % ppc, kg = tid / ppc;
  // per-thread invariants of the chunk loop (the loop body is counted in instructions: the kernel is issue-bound)
  const int px_off = px * C + mis;                     // this pixel's spectrum, floats into a slot
  const int px_end = px_off + C;
  const int full_valid = (((ppc * C + mis) * 4) & ~15) / 4;      // staged floats of a complete chunk
  int64_t p = (int64_t)blockIdx.x * ppc + px;          // this thread's pixel, advanced by one grid stride per chunk
  const int64_t pstep = (int64_t)gridDim.x * ppc;
  int64_t t = p / tile_pixels, sp = p - t * tile_pixels;         // (tile, pixel in tile), updated incrementally
  const int64_t tstep = pstep / tile_pixels, spstep = pstep - tstep * tile_pixels;
  int slot = 0;
  uint32_t phase = 0;
  for (int64_t i = 0; i < my_n; ++i, p += pstep) {
    const int64_t p0 = p - px;
    const float* b = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(bufs) + (size_t)slot * chunk_pitch);
    tc::mbar_wait(&full[slot], phase);
    if (kg < ngroups && p < n_pixels && !(dbg & 1)) {
      const float* xp = b + px_off;
      // only the last pixel of a chunk fetched from a misaligned cube, and the very last pixels of a cube whose byte
      // size is not a 16-byte multiple, miss the staged tail
      const int valid_floats = p0 + ppc <= n_pixels ? full_valid : ((((int)(n_pixels - p0) * C + mis) * 4) & ~15) / 4;
      const bool staged = px_end <= valid_floats;
      for (int k = kg; k < K; k += ngroups) {
        const float* wk = sw + k * Cp;
        float acc = 0.f;
        bool miss = false;
        if (staged) {
          // The kernel is issue-bound (ncu: 75 % issue-active at 60 % of DRAM peak), so the loop is written for
          // instruction count: weights come four at a time (the groups are aligned down to a multiple of four bands:
          // the extra leading bands are the pixel's own, with weight exactly 0), and "a band that counts is nodata"
          // is accumulated as sum |w| * [v == fill] -- one FSET + one FFMA instead of two compares, a select and an OR.
          float bad = 0.f;
          int c = rg.c0[k];
          const int c1 = rg.c1[k], ca = c & ~3, cb = c1 & ~3;
          if (ca < cb) {
#pragma unroll 2
            for (c = ca; c < cb; c += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(wk + c);
              const float v0 = xp[c], v1 = xp[c + 1], v2 = xp[c + 2], v3 = xp[c + 3];
              acc = fmaf(w4.x, v0, acc);                 // a zero weight contributes exactly 0 (finite radiances)
              acc = fmaf(w4.y, v1, acc);
              acc = fmaf(w4.z, v2, acc);
              acc = fmaf(w4.w, v3, acc);
              bad = fmaf(v0 == fill ? 1.f : 0.f, fabsf(w4.x), bad);
              bad = fmaf(v1 == fill ? 1.f : 0.f, fabsf(w4.y), bad);
              bad = fmaf(v2 == fill ? 1.f : 0.f, fabsf(w4.z), bad);
              bad = fmaf(v3 == fill ? 1.f : 0.f, fabsf(w4.w), bad);
            }
            c = cb;
          }
          for (; c < c1; ++c) {
            const float v = xp[c], wv = wk[c];
            acc = fmaf(wv, v, acc);
            bad = fmaf(v == fill ? 1.f : 0.f, fabsf(wv), bad);
          }
          miss = bad > 0.f;
        } else {
          for (int c = rg.c0[k]; c < rg.c1[k]; ++c) {
            const float v = px_off + c < valid_floats ? xp[c] : cube[p * C + c];
            const float wv = wk[c];
            if (wv != 0.f) {
              miss |= v == fill;
              acc = fmaf(wv, v, acc);
            }
          }
        }
        // planar per tile: (tile, K, tile_pixels) -- the reference's (K, H, W) for each scene
        out[(t * K + k) * tile_pixels + sp] = miss ? fill : acc;
      }
    }
    t += tstep;
    sp += spstep;
    if (sp >= tile_pixels) {
      sp -= tile_pixels;
      ++t;
    }
    __syncthreads();                                   // every reader of the slot is done
    if (tid == 0 && i + kSrfStages < my_n) issue(i + kSrfStages);
    if (++slot == kSrfStages) {
      slot = 0;
      phase ^= 1;
    }
  }
}

}  // namespace

extern "C" int sc_srf_aggregate(const float* cube_bip, int64_t n_pixels, int64_t tile_pixels, int C, const float* weights,
                                const int32_t* band_ranges_host, int K, float fill, float* out_planar, void* stream) {
  if (!cube_bip || !weights || !out_planar || n_pixels <= 0 || C < 1 || K < 1 || K > kSrfMaxK) return SC_ERR_BAD_ARG;
  if (tile_pixels <= 0 || n_pixels % tile_pixels) return SC_ERR_BAD_ARG;
  if (reinterpret_cast<uintptr_t>(cube_bip) & 3) return SC_ERR_BAD_ARG;
  const int mis = (int)((reinterpret_cast<uintptr_t>(cube_bip) & 15) / 4);
  // pixels per chunk: a multiple of 4 (16-byte multiple runs for any C), <= 256 threads
  // measured at 512 x 512 x 125 (bs 8): 16 KB chunks x 3 stages x 4 CTAs / SM = 5.1 TB/s (32 KB x 2 CTAs: 4.8; the
  // copy ring alone, arithmetic skipped: 7.1 TB/s)
  const int chunk_kb = getenv("STARCOP_SRF_KB") ? atoi(getenv("STARCOP_SRF_KB")) : 16;
  int ppc = (chunk_kb * 1024) / (C * 4) / 4 * 4;
  if (ppc > 256) ppc = 256;
  if (ppc < 4) return SC_ERR_UNSUPPORTED;              // spectra longer than 2048 bands
  cudaStream_t st = (cudaStream_t)stream;
  SrfRanges rg;
  for (int k = 0; k < K; ++k) {
    rg.c0[k] = band_ranges_host ? band_ranges_host[2 * k] : 0;
    rg.c1[k] = band_ranges_host ? band_ranges_host[2 * k + 1] : C;
    if (rg.c0[k] < 0 || rg.c1[k] > C || rg.c0[k] > rg.c1[k]) return SC_ERR_BAD_ARG;
  }
  const size_t chunk_pitch = ((size_t)ppc * C * 4 + 16 + 127) & ~(size_t)127;
  const size_t smem = kSrfStages * chunk_pitch + 64 + (size_t)K * ((C + 3) & ~3) * 4;
  cudaError_t e = cudaFuncSetAttribute(srf_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }
  const int64_t nchunks = (n_pixels + ppc - 1) / ppc;
  const int ctas = getenv("STARCOP_SRF_CTAS") ? atoi(getenv("STARCOP_SRF_CTAS")) : 4;
  const int grid = (int)(nchunks < ctas * kNumSMs ? nchunks : ctas * kNumSMs);
  srf_aggregate_kernel<<<grid, kSrfThreads, smem, st>>>(cube_bip, mis, n_pixels, tile_pixels, C, ppc, weights, rg, K, fill, out_planar,
                                                        getenv("STARCOP_SRF_DBG") ? atoi(getenv("STARCOP_SRF_DBG")) : 0);
  return check_launch();
}

// ------------------------------------------------------------------------------------------------
// configs[2] glue: matched-filter output + three cube bands -> the network's input, in ONE pass.
// For every pixel: ch0 = clip(mf, 0, 10000) (the cached-dataset clamp, sampling_dataset.py:292-293), ch1..3 = the
// cube's bands nearest 640 / 550 / 460 nm (TOA_AVIRIS_*nm products), then DataNormalizer.normalize_x
// (normalizer_module.py:134-135: IEEE division, clamp) and the NHWC pack at channel stride ld (channels >= 4 zero)
// that sc_normalize_pack would produce from the NCHW batch -- plus, optionally, the raw NCHW batch itself ("input" of
// the DataLoader contract) and weight_mag1c = clip(mf / 400, .1, 1) (feature_extration.py:32-35).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void chain_pack_kernel(const float* __restrict__ mf, const float* __restrict__ cube, int C, int b0, int b1, int b2,
                                  const double* __restrict__ off, const double* __restrict__ fac, const double* __restrict__ lo,
                                  const double* __restrict__ hi, int64_t HW, int64_t total, T* __restrict__ out_nhwc, int ld,
                                  float* __restrict__ out_nchw, float* __restrict__ weight) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = p / HW, s = p - b * HW;
    const float* px = cube + p * C;
    float v[4];
    const float m = mf[p];
    v[0] = m < 0.f ? 0.f : (m > 10000.f ? 10000.f : m);
    v[1] = px[b0];
    v[2] = px[b1];
    v[3] = px[b2];
    if (weight) {
      const float wv = v[0] / 400.f;
      weight[p] = wv < 0.1f ? 0.1f : (wv > 1.f ? 1.f : wv);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (out_nchw) out_nchw[(b * 4 + c) * HW + s] = v[c];
      const float d = (v[c] - (float)off[c]) / (float)fac[c];
      const float l = (float)lo[c], h = (float)hi[c];
      if (out_nhwc) out_nhwc[p * ld + c] = from_f<T>(d < l ? l : (d > h ? h : d));
    }
    if (out_nhwc)
      for (int c = 4; c < ld; ++c) out_nhwc[p * ld + c] = from_f<T>(0.f);
  }
}

extern "C" int sc_chain_pack(const float* mf, const float* cube_bip, int C, int band_r, int band_g, int band_b, const double* off,
                             const double* fac, const double* lo, const double* hi, int B, int64_t HW, void* out_nhwc, int ld,
                             int dtype, float* out_nchw, float* weight_loss, void* stream) {
  if (!mf || !cube_bip || !off || !fac || !lo || !hi || B <= 0 || HW <= 0 || (out_nhwc && ld < 4)) return SC_ERR_BAD_ARG;
  if (band_r < 0 || band_g < 0 || band_b < 0 || band_r >= C || band_g >= C || band_b >= C) return SC_ERR_BAD_ARG;
  const int64_t total = (int64_t)B * HW;
  int64_t blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  SC_DISPATCH_DTYPE(dtype, (chain_pack_kernel<T><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
                               mf, cube_bip, C, band_r, band_g, band_b, off, fac, lo, hi, HW, total, (T*)out_nhwc, ld, out_nchw,
                               weight_loss)));
  return check_launch();
}
