// mag1c, group-resident tensor-core kernel (fp32 radiance, alpha = 0, <= 512 pixels per group and <= 75 bands: the
// AVIRIS case, process_aviris.py:183-219 -> starcop/models/mag1c.py:176-348).  One pixel group per CTA.
//
// What bounds the matched filter is not the bytes (20 us / tile at the HBM roofline) but the S x S statistics of each
// group: C0 = cov(x) is a 73 x 512 x 73 contraction (38 k cycles of DFMA in the fp64-FMA kernel, mag1c.cu) followed by
// its inverse (73 dependent pivots).  Here
//   * the group is read from HBM ONCE (4-byte cp.async, every copy of the group in flight together) into a pixel-major
//     shared-memory landing zone (rows of 4 x odd words: 16-byte aligned, bank-conflict free);
//   * the spectra are centred on the (rounded) group mean and converted to 24-bit FIXED POINT per band,
//     q = rint((x - c_s) / 2^e_s), |q| < 2^23, which is within one bit of the fp32 inputs' own resolution;
//   * q = 65536 q2 + 256 q1 + q0 with balanced 8-bit digits; every digit is exactly representable in bf16, every
//     digit product is an integer < 2^14 and a 512-pixel sum of them stays below 2^24: the six slice-pair products
//     Q_a^T Q_b (a >= b) are therefore computed EXACTLY by tcgen05.mma kind::f16 with fp32 accumulators in TMEM
//     (M = 128 band rows, N = 80, K = 16 pixels per instruction; 6 accumulators x 80 columns, interleaved so that
//     six independent chains are in flight), whatever the accumulation order.  An all-ones operand row yields the
//     exact column sums of Q (the mean of the quantised data) from the same MMAs.  The covariance of the
//     fixed-point data is then assembled in fp64 and is exact to double rounding;
//   * P0 = C0^-1 by the symmetric sweep operator with the matrix in REGISTERS (warp = 5 columns, lane = 3 rows:
//     15 entries per thread; per pivot a warp reads 5 warp-uniform pivot-row entries and 3 coalesced pivot-column
//     vectors), one barrier per pivot, the next pivot's reciprocal (rcp.approx + two Newton steps) computed by its
//     owner off the critical path.  P0 never returns to shared memory: the iterations' mat-vecs use the registers;
//   * the reweighting iterations are the rank-two Woodbury updates of mag1c.cu in the centred fixed-point
//     coordinates; every scalar of an update is a combination of 14 dot products accumulated by the mat-vec threads
//     (no serial single-warp section);
//   * the two per-pixel passes of an iteration read the integers from shared memory and never convert them: a
//     biased 24-bit integer in the low word of a double with a zero high word IS the denormal u * 2^-1074, and fp64
//     FMAs take denormal operands at full speed, so a pass is one DFMA per element against coefficients pre-scaled
//     by 2^900 (no I2F / F2F / magic-number add).  All S x S and per-pixel algebra is fp64.
// Measured (B200, 8 tiles of 512 x 512 x 125, 73-band window): rmf 265 us / tile (fp64-FMA kernel: 404),
// 30 iterations 1011 us / tile (1103).  Phase clocks of one group (cycles): load 15 k, centre 6 k, convert + MMA 27 k,
// assemble 11 k, sweep 72 k, apply 10 k; one iteration 10.4 k (two shared-memory passes ~3.5 k each: 147 KB at the
// measured 120 B / clk / SM is 1.2 k; scalars 2.4 k).  Errors against the fp64 streaming kernel on fp64 data:
// 2e-6 of scale (rmf), 3e-4 (30 iterations), i.e. the sensitivity of the filter to an input perturbation of one
// fp32 ulp; the reference's own fp32 path is further from fp64 than that (tests/test_gpu_fullsize.py).
#include "common.cuh"
#include "tc_common.cuh"

using namespace sc;
using namespace tc;

namespace sc {
__device__ long long g_mag1c_clocks[16];
}

namespace {

constexpr int kT = 512;                 // threads = max pixels per group
constexpr int kSP = 80;                 // operand rows: S bands + the ones row + zero rows
constexpr int kChunk = 64;              // pixels per staged MMA chunk
constexpr int kSlice = (kSP / 8) * (kChunk / 8) * 128;   // one digit slice of a chunk, core-matrix tiled: 10240 B
constexpr int kBuf = 3 * kSlice;
constexpr int kStage = 2 * kBuf;        // 61440 B; the S x S fp64 matrix aliases it after the MMAs
constexpr int kVecs = 16;
constexpr int kCk = 128;                // pivot-row buffer: covers every column index a sweep thread can form
constexpr int kDots = 16;               // per-warp partial dot products of an iteration (14 used)
constexpr int kND = 14;
constexpr double kScaling = 1e5;        // mag1c.py:56
constexpr double kEpsilon = 1e-9;       // mag1c.py:57
constexpr double kBias = 8421504.0;      // 0x808080: the integers are kept biased (non-negative)

struct Lay {
  int lda, pitch;
  int off_vec, off_scal, off_land, off_bar, total, off_as;
};
__host__ __device__ inline Lay layout(int S) {
  Lay L;
  L.lda = S | 1;
  // pixel rows of 4 * (odd) words: 16-byte aligned for LDS.128, and eight consecutive rows tile all 32 banks
  L.pitch = 4 * (((S + 3) / 4) | 1);
  L.off_vec = kStage;
  L.off_scal = L.off_vec + (kVecs * kSP + 2 * kCk) * 8;
  L.off_land = L.off_scal + (64 + 16 * kDots) * 8;
  L.off_bar = L.off_land + kT * L.pitch * 4;
  L.total = L.off_bar + 64;
  L.off_as = (S * L.lda * 8 + 127) & ~127;     // a = R * mf per pixel, behind the matrix inside the staging area
  return L;
}

__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// The per-pixel passes never convert the integers: a biased 24-bit integer u in the LOW word of a double whose high
// word is zero IS the (denormal) double u * 2^-1074, exactly.  fp64 FMAs take denormal operands at full speed, so
// sum_s u_s c_s is one DFMA per element on coefficients pre-scaled by 2^900 (the products are normal numbers), with no
// I2F / F2F, no magic-number DADD and no per-element register moves; the bias is removed afterwards with the sum of
// the coefficients.  x * 2^900 * 2^-1074 = x * 2^-174.
__device__ __forceinline__ double u2dn(uint32_t u) { return __hiloint2double(0, (int)u); }
__device__ __forceinline__ double up900(double x) { return x * 8.452712498170644e270; }       // * 2^900 (exact)
__device__ __forceinline__ double up174(double x) { return x * 2.3945242826029513e52; }       // * 2^174 (exact)
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

#define MAG1C_CLK(i)                                              \
  do {                                                            \
    if (blockIdx.x == 0 && threadIdx.x == 0) g_mag1c_clocks[i] = clock64(); \
  } while (0)

__global__ void __launch_bounds__(kT, 1)
mag1c_tc_kernel(const float* __restrict__ x, int64_t pixel_stride, const int32_t* __restrict__ pix_idx,
                const int32_t* __restrict__ counts, int pmax, const double* __restrict__ tmpl,
                float* __restrict__ mf_out, float* __restrict__ al_out, int S, int num_iter, int skip_le,
                int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char sm[];
  const int g = blockIdx.x;
  const int P = counts ? counts[g] : pmax;
  if (P <= skip_le) return;                            // mag1c.py:166-168 (skip_le = 10): outputs keep the pre-fill
  const int32_t* idx = pix_idx + (int64_t)g * pmax;
  const Lay L = layout(S);
  const int LDA = L.lda, PITCH = L.pitch;
  double* A = reinterpret_cast<double*>(sm);           // S x LDA (aliases the operand staging)
  double* vec = reinterpret_cast<double*>(sm + L.off_vec);
  double* xbar = vec;                  // mean of the (quantised) group
  double* cen = xbar + kSP;            // centre c_s (an fp32 number)
  double* scl = cen + kSP;             // 2^e_s
  double* mq = scl + kSP;              // mean of the centred quantised data: xbar = cen + mq
  double* tp = mq + kSP;               // template
  double* tprev = tp + kSP;
  double* tcur = tprev + kSP;
  double* mu = tcur + kSP;
  double* wv = mu + kSP;
  double* pt = wv + kSP;
  double* pw = pt + kSP;
  double* pb = pw + kSP;
  double* cs = pb + kSP;               // cit * scl: the per-pixel filter coefficients of the integers
  double* xs = cs + kSP;               // xbar * scl (albedo factor pass)
  double* v = xs + kSP;                // scl * Q^T a
  double* dv = v + kSP;                // cen - mu
  double* ck0 = dv + kSP;              // [kCk] pivot row, double buffered
  double* ck1 = ck0 + kCk;
  double* scal = reinterpret_cast<double*>(sm + L.off_scal);
  double* dots = scal + 64;            // [16][kDots] per-warp partial dot products
  // front-end scratch aliases vectors the iterations initialise themselves (tprev .. pb are re-zeroed after the MMAs)
  float* cenf = reinterpret_cast<float*>(tprev);
  float* invs = cenf + kSP;
  float* partf = invs + kSP;                                      // [6][kSP]
  float* land = reinterpret_cast<float*>(sm + L.off_land);        // [512][PITCH] radiance, then offset-binary q
  uint32_t* landu = reinterpret_cast<uint32_t*>(land);
  int32_t* pidx = reinterpret_cast<int32_t*>(sm);                 // only until the group is loaded (staging area)
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sm + L.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);
  int* okp = reinterpret_cast<int*>(tmem_slot + 1);
  double* a_s = reinterpret_cast<double*>(sm + L.off_as);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kT / 32;
  const double N = (double)P;

  MAG1C_CLK(0);
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 32) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_barrier_init();
    *okp = 1;
  }
  for (int i = tid; i < P; i += kT) pidx[i] = idx[i];
  for (int i = tid; i < kVecs * kSP + 2 * kCk; i += kT) vec[i] = (i >= 4 * kSP && i < 4 * kSP + S) ? tmpl[i - 4 * kSP] : 0.0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- the ONE pass over HBM: one warp per pixel, lanes over bands, every copy asynchronous ---------------------
  for (int p = warp; p < P; p += NW) {
    const float* xp = x + (int64_t)pidx[p] * pixel_stride;
    float* dst = land + p * PITCH;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int s = lane + 32 * k;
      if (s < S) cp_async4(dst + s, xp + s);
    }
  }
  const int64_t my_pix = tid < P ? (int64_t)pidx[tid] : 0;
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  MAG1C_CLK(1);

  // ---- centre (rounded group mean) and per-band power-of-two scale ----------------------------------------------
  const int bs = tid % kSP, part = tid / kSP;            // 6 partial sums per band
  {
    float ref = 0.f, a0 = 0.f, a1 = 0.f;
    if (part < 6 && bs < S) {
      ref = land[bs];
      int p = part;
      for (; p + 6 < P; p += 12) {
        a0 += land[p * PITCH + bs] - ref;
        a1 += land[(p + 6) * PITCH + bs] - ref;
      }
      if (p < P) a0 += land[p * PITCH + bs] - ref;
      partf[part * kSP + bs] = a0 + a1;
    }
    __syncthreads();
    if (tid < S) {
      double t = 0.0;
      for (int q = 0; q < 6; ++q) t += (double)partf[q * kSP + tid];
      const float c = (float)((double)land[tid] + t / N);
      cenf[tid] = c;
      cen[tid] = (double)c;
    }
    __syncthreads();
    if (part < 6 && bs < S) {
      const float c = cenf[bs];
      float m = 0.f;
      for (int p = part; p < P; p += 6) m = fmaxf(m, fabsf(land[p * PITCH + bs] - c));
      partf[part * kSP + bs] = m;
    }
    __syncthreads();
    if (tid < S) {
      float m = 0.f;
      for (int q = 0; q < 6; ++q) m = fmaxf(m, partf[q * kSP + tid]);
      int e = 0;
      if (m > 0.f && frexpf(m, &e) > 0.99f) ++e;         // m <= 0.99 * 2^e: |q| <= 0x7F7F7F, three balanced BYTES
      e = e < -90 ? -90 : e;
      invs[tid] = ldexpf(1.f, 23 - e);                   // |q| <= 2^23
      scl[tid] = ldexp(1.0, e - 23);
    }
    __syncthreads();
  }
  MAG1C_CLK(2);

  // ---- exact covariance of the fixed-point data on the tensor cores ---------------------------------------------
  // operand rows S+1 .. 79 are zero in every chunk: written once
  for (int i = tid; i < 2 * 3 * (kSP - S - 1) * 8; i += kT) {
    const int o = i & 7, rs = i >> 3;
    const int s2 = S + 1 + rs % (kSP - S - 1), sl = rs / (kSP - S - 1);          // sl = buffer * 3 + slice
    *reinterpret_cast<uint4*>(sm + sl * kSlice + (s2 >> 3) * 1024 + o * 128 + (s2 & 7) * 16) = make_uint4(0, 0, 0, 0);
  }
  const int nchunks = (P + kChunk - 1) / kChunk;
  {
    const uint32_t idesc = make_idesc_bf16(128, kSP, 0, 0);
    // un-swizzled K-major operand: LBO (K direction) = 128 B, SBO (row groups) = 1024 B; slice / K-step offsets are
    // added to the 14-bit start-address field (>> 4)
    const uint64_t desc0 = make_smem_desc(smem_u32(sm), 128, 1024, 0);
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      if (c >= 2) mbar_wait(&mbar[buf], (uint32_t)(((c >> 1) - 1) & 1));   // the MMAs that read this buffer are done
      unsigned char* sb = sm + buf * kBuf;
      for (int u = tid; u < 8 * kSP; u += kT) {
        const int s = u % kSP, o = u / kSP;
        if (s > S) continue;                              // zero rows: written once at the start
        const int p0 = c * kChunk + o * 8;
        uint32_t w0[4], w1[4], w2[4];
        if (s < S) {
          const float cf = cenf[s], is = invs[s];
          uint32_t* lp = landu + p0 * PITCH + s;
          float d0[8], d1[8], d2[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            int q = 0;
            if (p0 + i < P) q = __float2int_rn((__uint_as_float(lp[i * PITCH]) - cf) * is);
            const uint32_t ub = (uint32_t)(q + 0x808080);  // balanced digits: byte k = digit k + 128
            lp[i * PITCH] = ub;                           // kept for the per-pixel passes
            d0[i] = __uint_as_float(__byte_perm(ub, 0x4B000000u, 0x7650)) - 8388736.f;   // 2^23 + byte - (2^23 + 128)
            d1[i] = __uint_as_float(__byte_perm(ub, 0x4B000000u, 0x7651)) - 8388736.f;
            d2[i] = __uint_as_float(__byte_perm(ub, 0x4B000000u, 0x7652)) - 8388736.f;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {                  // bf16 = upper half of the float (small integers: exact)
            w0[i] = __byte_perm(__float_as_uint(d0[2 * i]), __float_as_uint(d0[2 * i + 1]), 0x7632);
            w1[i] = __byte_perm(__float_as_uint(d1[2 * i]), __float_as_uint(d1[2 * i + 1]), 0x7632);
            w2[i] = __byte_perm(__float_as_uint(d2[2 * i]), __float_as_uint(d2[2 * i + 1]), 0x7632);
          }
        } else {                                          // s == S: the ones row (pixel counts / column sums)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            w0[i] = (p0 + 2 * i < P ? 0x3F80u : 0u) | (p0 + 2 * i + 1 < P ? 0x3F800000u : 0u);
            w1[i] = 0;
            w2[i] = 0;
          }
        }
        unsigned char* d = sb + (s >> 3) * 1024 + o * 128 + (s & 7) * 16;
        *reinterpret_cast<uint4*>(d) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
        *reinterpret_cast<uint4*>(d + kSlice) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        *reinterpret_cast<uint4*>(d + 2 * kSlice) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
      }
      fence_proxy_async();                                // generic-proxy stores -> the tensor core's async proxy
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint64_t dbase = desc0 + (uint64_t)((buf * kBuf) >> 4);
        // the K steps of one accumulator are a dependent chain (~130 cycles per MMA at N = 80): the six
        // slice pairs are interleaved so that six independent chains keep the tensor pipe busy
#pragma unroll
        for (int kk = 0; kk < kChunk / 16; ++kk) {
          int pair = 0;
#pragma unroll
          for (int a = 2; a >= 0; --a)
#pragma unroll
            for (int b = a; b >= 0; --b, ++pair) {        // (2,2) (2,1) (2,0) (1,1) (1,0) (0,0)
              umma_bf16(tmem_base + pair * kSP, dbase + (uint64_t)((a * kSlice + kk * 256) >> 4),
                        dbase + (uint64_t)((b * kSlice + kk * 256) >> 4), idesc, (c > 0 || kk > 0) ? 1u : 0u);
            }
        }
        umma_commit(&mbar[buf]);
      }
    }
    mbar_wait(&mbar[(nchunks - 1) & 1], (uint32_t)(((nchunks - 1) >> 1) & 1));   // MMAs retire in order
    tc_fence_after();
    __syncthreads();                                       // nobody still converts into the area A aliases
  }
  MAG1C_CLK(3);
  for (int i = tid; i < 7 * kSP; i += kT) tprev[i] = 0.0;   // tprev, tcur, mu, wv, pt, pw, pb (held the front-end scratch)
  const double invN = 1.0 / N;
  // TMEM -> W = (symmetric pairs) / 2 + (a > b pairs), scaled: C = (W + W^T) / N - m m^T
  {
    const int quad = warp & 3, cset = warp >> 2;
    const int i = quad * 32 + lane;
    if (quad * 32 < S) {                                   // warp-uniform
      for (int cb = cset; cb < kSP / 16; cb += 4) {
        if (cb * 16 > S) continue;
        // pair order (2,2) (2,1) (2,0) (1,1) (1,0) (0,0); symmetric pairs enter W halved
        const double coef[6] = {0.5 * 4294967296.0, 16777216.0, 65536.0, 0.5 * 65536.0, 256.0, 0.5};
        const double csum[6] = {0.0, 0.0, 65536.0, 0.0, 256.0, 1.0};
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + cb * 16;
        double w[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) w[jj] = 0.0;
        double sq = 0.0;
#pragma unroll
        for (int p3 = 0; p3 < 6; p3 += 3) {
          uint32_t r[3][16];
#pragma unroll
          for (int u = 0; u < 3; ++u) tmem_ld16_nowait(taddr + (p3 + u) * kSP, r[u]);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 3; ++u)
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const double dv2 = (double)__uint_as_float(r[u][jj]);
              w[jj] = fma(coef[p3 + u], dv2, w[jj]);
              if (cb * 16 + jj == S) sq = fma(csum[p3 + u], dv2, sq);   // the ones row: exact column sums of Q
            }
        }
        if (i < S) {
          const double si = scl[i];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const int j = cb * 16 + jj;
            if (j < S) A[i * LDA + j] = w[jj] * si * scl[j];
          }
          if (cb * 16 <= S && S < cb * 16 + 16) mq[i] = si * sq * invN;
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    const int NTRI = S * (S + 1) / 2;
    for (int e = tid; e < NTRI; e += kT) {
      int i2 = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
      while (i2 * (i2 + 1) / 2 > e) --i2;
      while ((i2 + 1) * (i2 + 2) / 2 <= e) ++i2;
      const int j2 = e - i2 * (i2 + 1) / 2;
      const double cv = (A[i2 * LDA + j2] + A[j2 * LDA + i2]) * invN - mq[i2] * mq[j2];
      A[i2 * LDA + j2] = cv;
      A[j2 * LDA + i2] = cv;
    }
    if (tid < S) {
      xbar[tid] = cen[tid] + mq[tid];
      xs[tid] = up900((cen[tid] + mq[tid]) * scl[tid]);
    }
    if (warp == NW - 1) {                                  // xbar . xbar and cen . xbar (albedo factor, mag1c.py:330)
      double mumu = 0.0, cenx = 0.0;
      for (int s2 = lane; s2 < S; s2 += 32) {
        const double xb = cen[s2] + mq[s2];
        mumu = fma(xb, xb, mumu);
        cenx = fma(cen[s2] - kBias * scl[s2], xb, cenx);   // the stored integers are biased
      }
      mumu = warp_sum_all(mumu);
      cenx = warp_sum_all(cenx);
      if (lane == 0) {
        scal[2] = mumu;
        scal[3] = cenx;
      }
    }
    __syncthreads();
  }
  MAG1C_CLK(4);

  // ---- P0 = C0^-1: symmetric sweep operator.  Warp w owns the 5 columns 5w .. 5w+4, lane l the rows l, l+32, l+64:
  // 15 entries per thread in REGISTERS for all S sweeps and for the iterations that follow (P0 is never re-read
  // from shared memory).  Per pivot a thread reads its warp's 5 pivot-row entries (warp-uniform addresses) and its 3
  // pivot-column entries (consecutive addresses): ~9 shared-memory wavefronts per warp instead of 26 with a
  // row-per-thread-group layout; one barrier per pivot, reciprocal by rcp.approx + two Newton steps.
  double av[3][5];
  const int jb = 5 * warp;
  {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        const int r = lane + 32 * ch, j = jb + c;
        av[ch][c] = (r < S && j < S) ? A[r * LDA + j] : 0.0;
      }
    // slot kCk-1 of the pivot-row buffer carries 1 / a_kk: the owner of the NEXT pivot's diagonal entry updates that
    // entry first and computes its reciprocal while the other threads are still updating, so the reciprocal chain
    // (MUFU + two Newton steps) is off the post-barrier critical path
    if (warp == 0 && lane == 0) {
      const double d0 = av[0][0];
      if (!(d0 > 0.0)) *okp = 0;
      ck0[kCk - 1] = fast_rcp(d0 > 0.0 ? d0 : 1.0);
    }
    for (int k = 0; k < S; ++k) {
      double* ck = (k & 1) ? ck1 : ck0;
      double* ckn = (k & 1) ? ck0 : ck1;
      const int kch = k >> 5, kl = k & 31;
      if (lane == kl) {                                     // row k = column k (symmetric): this warp's five entries
#pragma unroll
        for (int c = 0; c < 5; ++c) ck[jb + c] = kch == 0 ? av[0][c] : (kch == 1 ? av[1][c] : av[2][c]);
      }
      __syncthreads();
      const double inv = ck[kCk - 1];
      double cj[5], f[3];
#pragma unroll
      for (int c = 0; c < 5; ++c) cj[c] = ck[jb + c];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) f[ch] = ck[lane + 32 * ch] * inv;        // a_ik / d
      {  // next pivot's diagonal entry first: its owner publishes the reciprocal
        const int kn = k + 1, nch = kn >> 5, nl = kn & 31;
        if (kn < S && jb <= kn && kn < jb + 5 && lane == nl) {
          const int nc = kn - jb;
          double dn = 0.0;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch)
#pragma unroll
            for (int c = 0; c < 5; ++c)
              if (ch == nch && c == nc) dn = fma(-f[ch], cj[c], av[ch][c]);
          if (!(dn > 0.0)) {                               // not positive definite: the reference's Cholesky raises
            *okp = 0;
            dn = 1.0;
          }
          ckn[kCk - 1] = fast_rcp(dn);
        }
      }
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
#pragma unroll
        for (int c = 0; c < 5; ++c) av[ch][c] = fma(-f[ch], cj[c], av[ch][c]);   // a_ij <- a_ij - (a_ik / d) a_kj
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
        if (ch == kch && lane == kl) {                     // pivot row: a_kj <- a_kj / d
#pragma unroll
          for (int c = 0; c < 5; ++c) av[ch][c] = cj[c] * inv;
        }
      if (jb <= k && k < jb + 5) {                         // the warp that owns the pivot column (warp-uniform)
        const int kc = k - jb;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const double val = (ch == kch && lane == kl) ? -inv : f[ch];     // a_kk <- -1 / d;  a_ik <- a_ik / d
#pragma unroll
          for (int c = 0; c < 5; ++c)
            if (c == kc) av[ch][c] = val;
        }
      }
    }
    __syncthreads();                                       // A (= the staging area) is dead from here: reused below
  }
  MAG1C_CLK(5);

  // ---- iterations (mag1c.py:233-268, 312-335) in the centred fixed-point coordinates ---------------------------
  // x_p = cen + scl * q_p; xbar = cen + mq.  With a = R * mf, ma = mean(a), t = previous target, b = target:
  //   mean(modx) = mu = xbar - ma t;  cov(modx) = C0 + t w^T + w t^T + beta t t^T,  w = ma mq - scl * Q^T a / N
  //   (the cen * sum(a) terms of X^T a cancel exactly), beta = a.a / N - ma^2;
  //   cit = S_it^-1 b = pb - y1 pt - y2 pw (Woodbury; pb = P0 b, pt = P0 t, pw = P0 w);
  //   (x_p - mu) . cit = q_p . (scl * cit) + d . cit,  d = cen - mu.
  // Every scalar the update needs is a combination of eleven dot products of {t, w, b, d} with {pt, pw, pb}; they are
  // accumulated by the mat-vec threads themselves (one partial row per warp), and every warp then forms the scalars
  // redundantly: no single-warp serial section, three barriers between the two passes over the pixels.
  double sum_a = 0.0, sum_a2 = 0.0;
  double R_d = 0.0;
  float mf_f = 0.f;
  const uint32_t* myrow = landu + tid * PITCH;
  const int nwa = (S + 31) / 32;                           // warps whose threads own a matrix row in the mat-vec
  double* mvp = A;                                         // [2][16 warps][3 x 32 rows] partial products (A is dead)
  for (int it = 0; it <= num_iter; ++it) {
    if (tid < S) {
      if (it == 0) {
        mu[tid] = xbar[tid];
        tcur[tid] = tp[tid] * xbar[tid];           // target0 = template * mean(x)   (mag1c.py:312, :233)
        dv[tid] = -mq[tid];
      } else {
        const double ma = sum_a * invN;
        const double told = tcur[tid];
        tprev[tid] = told;
        pt[tid] = pb[tid];                         // P0 t: last iteration's P0 b
        const double m = xbar[tid] - ma * told;
        mu[tid] = m;
        dv[tid] = cen[tid] - m;
        wv[tid] = ma * mq[tid] - v[tid] * invN;
        tcur[tid] = tp[tid] * m;
      }
    }
    __syncthreads();
    if (it == 1) MAG1C_CLK(9);
    {  // pb = P0 b, pw = P0 w from the register-resident -P0: per-warp partial rows, then one thread per row
      double bj[5], wj[5], sb[3] = {0.0, 0.0, 0.0}, sw[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        bj[c] = tcur[jb + c];                              // warp-uniform; zero beyond S
        wj[c] = wv[jb + c];
      }
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          sb[ch] = fma(av[ch][c], bj[c], sb[ch]);
          sw[ch] = fma(av[ch][c], wj[c], sw[ch]);
        }
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        mvp[(warp * 3 + ch) * 32 + lane] = sb[ch];
        mvp[((NW + warp) * 3 + ch) * 32 + lane] = sw[ch];
      }
    }
    __syncthreads();
    if (it == 1) MAG1C_CLK(10);
    if (warp < nwa) {   // row r = tid: totals over the 16 column blocks, and the row's terms of the eleven dot products
      const int r = tid;
      double sb = 0.0, sw = 0.0;
#pragma unroll
      for (int wq = 0; wq < NW; ++wq) {
        sb -= mvp[wq * 96 + r];                            // av holds -P0
        sw -= mvp[(NW + wq) * 96 + r];
      }
      double e[kND];
#pragma unroll
      for (int q = 0; q < kND; ++q) e[q] = 0.0;
      if (r < S) {
        pb[r] = sb;
        pw[r] = sw;
        const double tr = tprev[r], wr = wv[r], br = tcur[r], dr = dv[r], ptr = pt[r];
        e[0] = tr * ptr; e[1] = tr * sw; e[2] = wr * sw; e[3] = tr * sb; e[4] = wr * sb;
        e[5] = br * sb;  e[6] = br * ptr; e[7] = br * sw; e[8] = dr * sb; e[9] = dr * ptr; e[10] = dr * sw;
        const double sr = scl[r];
        e[11] = sr * sb; e[12] = sr * ptr; e[13] = sr * sw;                  // sum of the integer coefficients
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < kND; ++q) e[q] += __shfl_xor_sync(0xffffffffu, e[q], o);
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < kND; ++q) dots[warp * kDots + q] = e[q];
      }
    }
    __syncthreads();
    double nrm, shift;
    {
      double tot = 0.0;
      if (lane < kND)
        for (int wq = 0; wq < nwa; ++wq) tot += dots[wq * kDots + lane];
      const double g11 = __shfl_sync(0xffffffffu, tot, 0), g12 = __shfl_sync(0xffffffffu, tot, 1),
                   g22 = __shfl_sync(0xffffffffu, tot, 2), r1 = __shfl_sync(0xffffffffu, tot, 3),
                   r2 = __shfl_sync(0xffffffffu, tot, 4), bpb = __shfl_sync(0xffffffffu, tot, 5),
                   bpt = __shfl_sync(0xffffffffu, tot, 6), bpw = __shfl_sync(0xffffffffu, tot, 7),
                   dpb = __shfl_sync(0xffffffffu, tot, 8), dpt = __shfl_sync(0xffffffffu, tot, 9),
                   dpw = __shfl_sync(0xffffffffu, tot, 10), spb = __shfl_sync(0xffffffffu, tot, 11),
                   spt = __shfl_sync(0xffffffffu, tot, 12), spw = __shfl_sync(0xffffffffu, tot, 13);
      double y1 = 0.0, y2 = 0.0;
      if (it > 0) {
        const double ma = sum_a * invN;
        const double beta = sum_a2 * invN - ma * ma;
        const double m11 = g11, m12 = 1.0 + g12, m22 = g22 - beta;
        const double idet = 1.0 / (m11 * m22 - m12 * m12);
        y1 = (r1 * m22 - m12 * r2) * idet;
        y2 = (m11 * r2 - m12 * r1) * idet;
      }
      nrm = bpb - y1 * bpt - y2 * bpw;                     // b . cit
      // (cen - mu) . cit, minus the bias of the stored integers times the sum of their coefficients
      shift = (dpb - y1 * dpt - y2 * dpw) - kBias * (spb - y1 * spt - y2 * spw);
      if (it > 0 && nrm < 1.0) nrm = 1.0;                  // mag1c.py:264-266 (not applied inside rmf)
      if (tid < kSP) cs[tid] = tid < S ? up900((pb[tid] - y1 * pt[tid] - y2 * pw[tid]) * scl[tid]) : 0.0;
    }
    __syncthreads();
    if (it == 1) MAG1C_CLK(11);
    // ---- matched-filter apply: one thread per pixel over the resident integers ----------------------------------
    const bool last = it == num_iter;
    double la = 0.0, la2 = 0.0;
    double a = 0.0;
    if (tid < P) {
      // the pixel's PITCH integers by 16-byte loads (rows are 16-byte aligned; bands S .. PITCH-1 meet zero
      // coefficients); coefficients by warp-uniform 16-byte broadcast loads
      double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
      double mf;
      const uint4* row4 = reinterpret_cast<const uint4*>(myrow);
      if (it == 0) {
        double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
        for (int s4 = 0; s4 < PITCH / 4; ++s4) {
          const uint4 qv = row4[s4];
          const double2 c01 = *reinterpret_cast<const double2*>(cs + 4 * s4);
          const double2 c23 = *reinterpret_cast<const double2*>(cs + 4 * s4 + 2);
          const double2 b01 = *reinterpret_cast<const double2*>(xs + 4 * s4);
          const double2 b23 = *reinterpret_cast<const double2*>(xs + 4 * s4 + 2);
          const double q0 = u2dn(qv.x), q1 = u2dn(qv.y), q2 = u2dn(qv.z), q3 = u2dn(qv.w);
          d0 = fma(q0, c01.x, d0);
          d1 = fma(q1, c01.y, d1);
          d2 = fma(q2, c23.x, d2);
          d3 = fma(q3, c23.y, d3);
          x0 = fma(q0, b01.x, x0);
          x1 = fma(q1, b01.y, x1);
          x2 = fma(q2, b23.x, x2);
          x3 = fma(q3, b23.y, x3);
        }
        const double R = (up174((x0 + x1) + (x2 + x3)) + scal[3]) / scal[2];  // x_p . xbar / xbar . xbar   (mag1c.py:330)
        mf = (up174((d0 + d1) + (d2 + d3)) + shift) / (R * nrm);             // mag1c.py:332
        const float R_f = (float)R;
        R_d = (double)R_f;
        al_out[my_pix] = R_f;
      } else {
#pragma unroll 4
        for (int s4 = 0; s4 < PITCH / 4; ++s4) {
          const uint4 qv = row4[s4];
          const double2 c01 = *reinterpret_cast<const double2*>(cs + 4 * s4);
          const double2 c23 = *reinterpret_cast<const double2*>(cs + 4 * s4 + 2);
          d0 = fma(u2dn(qv.x), c01.x, d0);
          d1 = fma(u2dn(qv.y), c01.y, d1);
          d2 = fma(u2dn(qv.z), c23.x, d2);
          d3 = fma(u2dn(qv.w), c23.y, d3);
        }
        const double dot = up174((d0 + d1) + (d2 + d3)) + shift;   // (x_p - mu) . cit
        const double reg = 1.0 / (R_d * ((double)mf_f + kEpsilon));   // mag1c.py:255
        mf = (dot - reg) / (R_d * nrm);                    // mag1c.py:267
      }
      mf = mf > 0.0 ? mf : 0.0;                            // relu
      mf_f = (float)mf;
      if (last) mf_out[my_pix] = (float)((double)mf_f * kScaling);
      a = R_d * (double)mf_f;
      la = a;
      la2 = a * a;
    }
    if (last) break;
    if (it == 1) MAG1C_CLK(12);
    a_s[tid] = up900(a);
    la = warp_sum_all(la);
    la2 = warp_sum_all(la2);
    if (lane == 0) {
      scal[8 + warp] = la;
      scal[8 + NW + warp] = la2;
    }
    __syncthreads();                         // also publishes a_s
    if (it == 1) MAG1C_CLK(13);
    sum_a = 0.0;
    sum_a2 = 0.0;
#pragma unroll
    for (int wq = 0; wq < NW; ++wq) {
      sum_a += scal[8 + wq];
      sum_a2 += scal[8 + NW + wq];
    }
    {  // v = scl * Q^T a: warp w owns the bands w, w+16, ..., interleaved (independent chains, batched shuffles);
       // each lane keeps its 16 pixels' a in registers
      double ar[kT / 32];
#pragma unroll
      for (int k = 0; k < kT / 32; ++k) ar[k] = a_s[lane + 32 * k];
      double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      const uint32_t* col = landu + lane * PITCH + warp;
#pragma unroll
      for (int k = 0; k < kT / 32; ++k)
#pragma unroll
        for (int b = 0; b < 5; ++b)
          if (warp + 16 * b < S) acc[b] = fma(u2dn(col[32 * k * PITCH + 16 * b]), ar[k], acc[b]);  // pad pixels: a = 0
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int b = 0; b < 5; ++b) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
      if (lane == 0) {
#pragma unroll
        for (int b = 0; b < 5; ++b)
          if (warp + 16 * b < S) v[warp + 16 * b] = (up174(acc[b]) - kBias * sum_a) * scl[warp + 16 * b];
      }
    }
    __syncthreads();
    if (it == 0) MAG1C_CLK(7);
    if (it == 1) MAG1C_CLK(14);
  }
  if (num_iter == 0) MAG1C_CLK(6);
  MAG1C_CLK(8);
  if (tid == 0 && !*okp && status) atomicAdd(status, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

namespace sc {

// -> SC_OK when the launch was made, SC_ERR_UNSUPPORTED when the shape does not fit this kernel (caller falls back)
int mag1c_tc_launch(const float* x, int64_t pixel_stride, const int32_t* pix_idx, const int32_t* counts, int pmax,
                    const double* tmpl, float* mf_out, float* al_out, int G, int S, int num_iter, int skip_le,
                    int* status, cudaStream_t st) {
  if (pmax > kT || S + 1 > kSP || S < 16) return SC_ERR_UNSUPPORTED;
  const Lay L = layout(S);
  if (L.total > 227 * 1024 || L.off_as + kT * 8 > kStage) return SC_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(mag1c_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  if (e != cudaSuccess) {
    g_last_error = e;
    return SC_ERR_CUDA;
  }
  mag1c_tc_kernel<<<G, kT, L.total, st>>>(x, pixel_stride, pix_idx, counts, pmax, tmpl, mf_out, al_out, S, num_iter,
                                          skip_le, status);
  return check_launch();
}

}  // namespace sc

extern "C" int sc_debug_mag1c_clocks(long long* out16) {
  return cudaMemcpyFromSymbol(out16, sc::g_mag1c_clocks, sizeof(long long) * 16) == cudaSuccess ? SC_OK : SC_ERR_CUDA;
}
