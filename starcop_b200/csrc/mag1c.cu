// mag1c: albedo-corrected reweighted-L1 matched filter (starcop/models/mag1c.py:176-348) for one
// pixel group per CTA.
//
// The reference recomputes, 1+num_iter times per group, the S x S covariance of
// modx = x - R*mf*target (a P x S x S bmm), its Cholesky factor and a solve.  Here the group's
// second-moment matrix X^T X is accumulated ONCE (fp64) and every later covariance follows from
// the rank-one structure modx = X - a t^T (a = R*mf per pixel, t = previous target):
//     sum modx modx^T = X^T X - v t^T - t v^T + (a.a) t t^T,   v = X^T a,
//     mean(modx)      = xbar - mean(a) t,
// so an iteration costs one pass over the group's pixels (the matched-filter apply fused with the
// accumulation of v, sum a, sum a^2 for the next iteration) plus an S x S Cholesky in shared
// memory.  All S x S arithmetic is fp64 (packed lower triangles), which is at least as accurate
// as the reference's fp32/fp64 bmm of centred data.  Pixels are addressed through an index list,
// which is how func_by_groups' gather / scatter (mag1c.py:161-172) is expressed on the device.
#include <stdlib.h>

#include "common.cuh"

using namespace sc;

namespace sc {
// mag1c_tc.cu: group-resident kernel with the covariance on the tensor cores
int mag1c_tc_launch(const float* x, int64_t pixel_stride, const int32_t* pix_idx, const int32_t* counts, int pmax,
                    const double* tmpl, float* mf_out, float* al_out, int G, int S, int num_iter, int skip_le,
                    int* status, cudaStream_t st);
}

namespace {

constexpr int kThreads = 128;
constexpr int kTile = 64;           // pixels staged per shared-memory tile (double buffered)
constexpr double kScaling = 1e5;    // mag1c.py:56
constexpr double kEpsilon = 1e-9;   // mag1c.py:57

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // j <= i

// asynchronous global -> shared copies (LDGSTS): the whole tile is in flight at once, no registers
template <int BYTES>
__device__ __forceinline__ void cp_async(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "n"(BYTES)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < kThreads / 32; ++i) s += red[i];
  return s;
}

// Left-looking Cholesky of the (S+1) x (S+1) matrix [[C, b], [b^T, .]] stored row-major with leading
// dimension LD (LD = 1 mod 16: column accesses by consecutive threads are bank-conflict free for
// doubles).  Thread i owns row i; row S holds b, so on exit it holds z = L^-1 b (the forward
// substitution comes for free).  Only the lower triangle is referenced.
__device__ void cholesky_aug(double* A, int S, int LD, int* ok, double* diag) {
  const int i = threadIdx.x;
  for (int j = 0; j < S; ++j) {
    double sdot = 0.0;
    if (i >= j && i <= S) {
      const double* ri = A + i * LD;
      const double* rj = A + j * LD;
      // four independent partial sums: the fp64 FMA chain is latency-, not throughput-bound
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = 0;
      for (; k + 4 <= j; k += 4) {
        s0 += ri[k] * rj[k];
        s1 += ri[k + 1] * rj[k + 1];
        s2 += ri[k + 2] * rj[k + 2];
        s3 += ri[k + 3] * rj[k + 3];
      }
      for (; k < j; ++k) s0 += ri[k] * rj[k];
      sdot = ri[j] - ((s0 + s1) + (s2 + s3));
      if (i == j) {
        if (!(sdot > 0.0)) {
          *ok = 0;
          sdot = 1.0;
        }
        *diag = sqrt(sdot);
      }
    }
    __syncthreads();
    if (i >= j && i <= S) A[i * LD + j] = (i == j) ? *diag : sdot / *diag;
    __syncthreads();
  }
}

// backward substitution L^T c = z (z = row S of A), one warp; c_i = (z_i - sum_{k>i} L_ki c_k) / L_ii
__device__ void chol_backsolve(const double* A, int S, int LD, double* c) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const double* z = A + S * LD;
    for (int i = S - 1; i >= 0; --i) {
      double s = 0.0;
      for (int k = i + 1 + lane; k < S; k += 32) s += A[k * LD + i] * c[k];
      s = warp_sum(s);
      if (lane == 0) c[i] = (z[i] - s) / A[i * LD + i];
      __syncwarp();
    }
  }
  __syncthreads();
}

__host__ __device__ inline int mag1c_ld(int S) { return ((S + 1) + 14) / 16 * 16 + 1; }   // >= S+1, = 1 mod 16

template <typename TS>
__global__ void __launch_bounds__(kThreads)
mag1c_kernel(const TS* __restrict__ x, int64_t pixel_stride, const int32_t* __restrict__ pix_idx,
             const int32_t* __restrict__ counts, int pmax, const double* __restrict__ tmpl, TS* __restrict__ mf_out,
             TS* __restrict__ al_out, int S, int num_iter, double alpha, int skip_le, int* __restrict__ status) {
  extern __shared__ double smd[];
  const int g = blockIdx.x;
  const int P = counts ? counts[g] : pmax;
  if (P <= skip_le) return;                            // mag1c.py:166-168 (skip_le = 10): too few pixels, outputs keep the pre-fill
  const int32_t* idx = pix_idx + (int64_t)g * pmax;
  const int NT = S * (S + 1) / 2;
  const int LD = mag1c_ld(S);
  double* XtX = smd;                 // packed lower triangle of sum x x^T
  double* A = XtX + NT;              // (S+1) x LD working covariance / Cholesky factor, row S = rhs
  double* xbar = A + (S + 1) * LD;
  double* mu = xbar + S;
  double* tprev = mu + S;            // target used inside modx
  double* tcur = tprev + S;          // target = template * mu
  double* v = tcur + S;              // X^T a
  double* cit = v + S;
  double* z = cit + S;
  double* tp = z + S;                // template
  double* red = tp + S;              // [8] block-reduction scratch
  double* a_s = red + 8;             // [kTile] a = R*mf of the staged pixels
  TS* tile0 = reinterpret_cast<TS*>(a_s + kTile);                 // 2 x [kTile][S] staged spectra (double buffer)
  uchar2* rc = reinterpret_cast<uchar2*>(tile0 + 2 * kTile * S);  // packed index -> (row, col)
  int32_t* pidx = reinterpret_cast<int32_t*>(rc + ((NT + 3) & ~3)); // [P] this group's pixel indices (8 B aligned)
  // the per-pixel state of the iteration (albedo factor R and the current matched-filter value, both in the storage
  // precision) lives in the OUTPUT arrays themselves (one L2-resident read + write per pixel and iteration): 2 x P
  // elements less shared memory per CTA, which is what lets two EMIT groups (2560 pixels, fp64) share an SM
  __shared__ int ok;
  const int tid = threadIdx.x;

  if (tid == 0) ok = 1;
  for (int i = tid; i < S; i += kThreads)
    for (int j = 0; j <= i; ++j) rc[tri(i, j)] = make_uchar2((unsigned char)i, (unsigned char)j);
  for (int i = tid; i < NT; i += kThreads) XtX[i] = 0.0;
  for (int i = tid; i < S; i += kThreads) {
    tp[i] = tmpl[i];
    xbar[i] = 0.0;
  }
  for (int i = tid; i < P; i += kThreads) pidx[i] = idx[i];
  __syncthreads();

  // stage pixels [p0, p0+np) of the group into buffer b with cp.async: consecutive threads copy
  // consecutive bands of a pixel row; one commit group per tile
  auto stage = [&](int b, int p0) {
    const int np = min(kTile, P - p0);
    TS* t = tile0 + b * kTile * S;
    for (int e = tid; e < np * S; e += kThreads) {
      int pp = e / S, s2 = e - pp * S;
      cp_async<sizeof(TS)>(t + pp * S + s2, x + (int64_t)pidx[p0 + pp] * pixel_stride + s2);
    }
    cp_async_commit();
  };
  const int nchunks = (P + kTile - 1) / kTile;

  // ---- pass 0: xbar and X^T X -----------------------------------------------------------------------
  stage(0, 0);
  for (int c = 0; c < nchunks; ++c) {
    const int p0 = c * kTile;
    const int np = min(kTile, P - p0);
    const TS* tile = tile0 + (c & 1) * kTile * S;
    if (c + 1 < nchunks) {
      stage((c + 1) & 1, p0 + kTile);     // prefetch the next tile while this one is consumed
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    for (int e = tid; e < NT; e += kThreads) {
      const int i = rc[e].x, j = rc[e].y;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int pp = 0;
      for (; pp + 4 <= np; pp += 4) {
        a0 += (double)tile[pp * S + i] * (double)tile[pp * S + j];
        a1 += (double)tile[(pp + 1) * S + i] * (double)tile[(pp + 1) * S + j];
        a2 += (double)tile[(pp + 2) * S + i] * (double)tile[(pp + 2) * S + j];
        a3 += (double)tile[(pp + 3) * S + i] * (double)tile[(pp + 3) * S + j];
      }
      for (; pp < np; ++pp) a0 += (double)tile[pp * S + i] * (double)tile[pp * S + j];
      XtX[e] += (a0 + a1) + (a2 + a3);
    }
    for (int s2 = tid; s2 < S; s2 += kThreads) {
      double acc = 0.0;
      for (int pp = 0; pp < np; ++pp) acc += (double)tile[pp * S + s2];
      xbar[s2] += acc;
    }
    __syncthreads();                        // everyone done with this buffer before it is refilled
  }
  const double N = (double)P;
  for (int s2 = tid; s2 < S; s2 += kThreads) {
    xbar[s2] /= N;
    mu[s2] = xbar[s2];
    tcur[s2] = tp[s2] * xbar[s2];      // target0 = template * mean(x)   (mag1c.py:312, :233)
    tprev[s2] = 0.0;
    v[s2] = 0.0;
  }
  __syncthreads();
  double mumu = 0.0;
  for (int s2 = 0; s2 < S; ++s2) mumu += xbar[s2] * xbar[s2];   // every thread: mu.mu for the albedo factor

  double sum_a = 0.0, sum_a2 = 0.0;       // block-uniform after block_sum
  for (int it = 0; it <= num_iter; ++it) {
    // ---- covariance of modx, C = (1-alpha) S + alpha diag(S) ------------------------------------
    if (it > 0) {
      const double ma = sum_a / N;
      for (int s2 = tid; s2 < S; s2 += kThreads) {
        tprev[s2] = tcur[s2];
        mu[s2] = xbar[s2] - ma * tcur[s2];
      }
      __syncthreads();
      for (int s2 = tid; s2 < S; s2 += kThreads) tcur[s2] = tp[s2] * mu[s2];
      __syncthreads();
    }
    for (int e = tid; e < NT; e += kThreads) {
      const int i = rc[e].x, j = rc[e].y;
      double sxx = XtX[e];
      if (it > 0) sxx += -v[i] * tprev[j] - tprev[i] * v[j] + sum_a2 * tprev[i] * tprev[j];
      double c = sxx / N - mu[i] * mu[j];
      if (i != j) c *= (1.0 - alpha);
      A[i * LD + j] = c;
    }
    for (int s2 = tid; s2 < S; s2 += kThreads) A[S * LD + s2] = tcur[s2];
    __syncthreads();
    cholesky_aug(A, S, LD, &ok, z);
    chol_backsolve(A, S, LD, cit);
    double nrm = 0.0, mucit = 0.0;
    for (int s2 = 0; s2 < S; ++s2) {
      nrm += tcur[s2] * cit[s2];
      mucit += mu[s2] * cit[s2];
    }
    if (it > 0 && nrm < 1.0) nrm = 1.0;               // mag1c.py:264-266 (not applied inside rmf)
    // ---- apply + next iteration's statistics, kTile staged pixels at a time ------------------------
    const bool last = it == num_iter;
    double la = 0.0, la2 = 0.0;
    double vacc = 0.0;                                // thread s2 = tid < S owns v[s2]
    stage(0, 0);
    for (int c = 0; c < nchunks; ++c) {
      const int p0 = c * kTile;
      const int np = min(kTile, P - p0);
      const TS* tile = tile0 + (c & 1) * kTile * S;
      if (c + 1 < nchunks) {
        stage((c + 1) & 1, p0 + kTile);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      for (int pp = tid; pp < np; pp += kThreads) {   // one thread per staged pixel (row stride S is odd: no conflicts)
        const int64_t pix = pidx[p0 + pp];
        const TS* xp = tile + pp * S;
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0, xmu = 0.0;
        int s2 = 0;
        for (; s2 + 4 <= S; s2 += 4) {
          d0 += (double)xp[s2] * cit[s2];
          d1 += (double)xp[s2 + 1] * cit[s2 + 1];
          d2 += (double)xp[s2 + 2] * cit[s2 + 2];
          d3 += (double)xp[s2 + 3] * cit[s2 + 3];
        }
        for (; s2 < S; ++s2) d0 += (double)xp[s2] * cit[s2];
        const double dot = (d0 + d1) + (d2 + d3);
        if (it == 0) {
          double x0 = 0.0, x1 = 0.0;
          for (s2 = 0; s2 + 2 <= S; s2 += 2) {
            x0 += (double)xp[s2] * xbar[s2];
            x1 += (double)xp[s2 + 1] * xbar[s2 + 1];
          }
          for (; s2 < S; ++s2) x0 += (double)xp[s2] * xbar[s2];
          xmu = x0 + x1;
        }
        double R, mf;
        if (it == 0) {
          R = xmu / mumu;                                // mag1c.py:330
          mf = (dot - mucit) / (R * nrm);                // mag1c.py:332
        } else {
          R = (double)al_out[pix];
          double mf_old = (double)mf_out[pix];
          double reg = 1.0 / (R * (mf_old + kEpsilon));  // mag1c.py:255
          mf = ((dot - mucit) - reg) / (R * nrm);        // mag1c.py:267
        }
        mf = mf > 0.0 ? mf : 0.0;                        // relu
        TS mfs = (TS)mf;                                 // working values kept in the storage precision
        if (it == 0) al_out[pix] = (TS)R;
        mf_out[pix] = last ? (TS)((double)mfs * kScaling) : mfs;
        double a = (double)(TS)R * (double)mfs;
        a_s[pp] = a;
        la += a;
        la2 += a * a;
      }
      __syncthreads();
      if (!last && tid < S) {
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        int pp = 0;
        for (; pp + 4 <= np; pp += 4) {
          v0 += a_s[pp] * (double)tile[pp * S + tid];
          v1 += a_s[pp + 1] * (double)tile[(pp + 1) * S + tid];
          v2 += a_s[pp + 2] * (double)tile[(pp + 2) * S + tid];
          v3 += a_s[pp + 3] * (double)tile[(pp + 3) * S + tid];
        }
        for (; pp < np; ++pp) v0 += a_s[pp] * (double)tile[pp * S + tid];
        vacc += (v0 + v1) + (v2 + v3);
      }
      __syncthreads();                      // buffer and a_s free for the next tile
    }
    if (last) break;
    __syncthreads();
    for (int s2 = tid; s2 < S; s2 += kThreads) {
      if (s2 == tid) v[s2] = vacc;
    }
    sum_a = block_sum(la, red);
    sum_a2 = block_sum(la2, red);
    __syncthreads();
  }
  if (tid == 0 && !ok && status) atomicAdd(status, 1);
}


// ------------------------------------------------------------------------------------------------
// Group-RESIDENT fast path (fp32 radiance, alpha = 0, <= 512 pixels per group: the AVIRIS case,
// process_aviris.py:183-219).  The group's spectra are read from HBM exactly ONCE into shared memory
// (band-major, 73 x 512 floats = 150 KB), and the 30 reweighting iterations never refactorise:
//   S_it = C0 + t w^T + w t^T + beta t t^T      (t = previous target, w = ma*xbar - X^T a / N,
//                                                beta = a.a/N - ma^2, ma = mean(a), a = R*mf)
// is a symmetric rank-two update of the group's plain covariance C0, so with P0 = C0^-1 (one in-place
// Gauss-Jordan inversion, fp64) the Woodbury identity gives
//   S_it^-1 b = P0 b - [P0 t, P0 w] (K^-1 + U^T P0 U)^-1 U^T P0 b,   U = [t, w], K^-1 = [[0,1],[1,-beta]],
// i.e. two 73 x 73 mat-vecs and a 2 x 2 solve per iteration (P0 t is last iteration's P0 b) instead of
// a covariance rebuild + Cholesky.  Same mathematics as the reference's loop (mag1c.py:233-268), all
// S x S algebra in fp64.  Per iteration the kernel then makes two passes over the RESIDENT spectra: the
// matched-filter apply (one thread per pixel) and v = X^T a (one warp per band).
// ------------------------------------------------------------------------------------------------
constexpr int kResThreads = 512;
constexpr int kResMaxP = 512;
constexpr int kResLdx = 513;       // odd row pitch: band-major rows start on different banks

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__host__ __device__ inline int res_sp(int S) { return (S + 7) / 8 * 8; }
__host__ __device__ inline int res_lda(int S) { return S | 1; }
__host__ __device__ inline size_t res_smem_bytes(int S) {
  const int SP = res_sp(S);
  return (size_t)(S * res_lda(S) + 1 + 14 * SP + kResMaxP + 64) * 8 + (size_t)SP * kResLdx * 4 + kResMaxP * 4 + 16;
}

__global__ void __launch_bounds__(kResThreads, 1)
mag1c_resident_kernel(const float* __restrict__ x, int64_t pixel_stride, const int32_t* __restrict__ pix_idx,
                      const int32_t* __restrict__ counts, int pmax, const double* __restrict__ tmpl,
                      float* __restrict__ mf_out, float* __restrict__ al_out, int S, int num_iter, int skip_le,
                      int* __restrict__ status) {
  extern __shared__ double smd[];
  const int g = blockIdx.x;
  const int P = counts ? counts[g] : pmax;
  if (P <= skip_le) return;                            // mag1c.py:166-168 (skip_le = 10): too few pixels, outputs keep the pre-fill
  const int32_t* idx = pix_idx + (int64_t)g * pmax;
  const int SP = res_sp(S), LDA = res_lda(S);
  double* A = smd;                       // S x LDA: C0, then P0 = C0^-1
  double* xbar = A + ((S * LDA + 1) & ~1);     // vectors 16-byte aligned (double2 loads of cit / xbar)
  double* tp = xbar + SP;                // template
  double* tprev = tp + SP;               // t: target inside modx
  double* tcur = tprev + SP;             // b: target = template * mu
  double* mu = tcur + SP;
  double* wv = mu + SP;
  double* pt = wv + SP;                  // P0 t
  double* pw = pt + SP;                  // P0 w
  double* pb = pw + SP;                  // P0 b
  double* cit = pb + SP;
  double* v = cit + SP;                  // X^T a
  double* rowk = v + SP;                 // Gauss-Jordan pivot row / column
  double* colk = rowk + SP;
  double* spare = colk + SP;
  double* a_s = spare + SP;              // [512] a = R*mf per pixel
  double* scal = a_s + kResMaxP;         // [64] scalars + block-reduction scratch
  float* Xt = reinterpret_cast<float*>(scal + 64);     // [SP][513] band-major spectra
  int32_t* pidx = reinterpret_cast<int32_t*>(Xt + (size_t)SP * kResLdx);
  __shared__ int ok;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kResThreads / 32;
  const double N = (double)P;

  if (tid == 0) ok = 1;
  for (int i = tid; i < P; i += kResThreads) pidx[i] = idx[i];
  for (int i = tid; i < S; i += kResThreads) tp[i] = tmpl[i];
  for (int i = tid; i < (SP - S) * kResLdx; i += kResThreads) Xt[(size_t)S * kResLdx + i] = 0.f;   // pad bands
  if (P < kResMaxP) {                                                                               // pad pixels
    const int npad = kResMaxP - P;
    for (int i = tid; i < S * npad; i += kResThreads) Xt[(size_t)(i / npad) * kResLdx + P + i % npad] = 0.f;
  }
  __syncthreads();

  // ---- the ONE pass over HBM: pixel-major global -> band-major shared.  One warp per pixel, lanes over bands:
  // coalesced 4*S-byte reads, transposed stores hit consecutive banks (odd row pitch); all loads independent.
  for (int p0 = warp; p0 < P; p0 += 4 * NW) {
    float r[4][3];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int p = p0 + q * NW;
      const float* xp = x + (int64_t)pidx[p < P ? p : p0] * pixel_stride;
#pragma unroll
      for (int k = 0; k < 3; ++k) r[q][k] = (p < P && lane + 32 * k < S) ? xp[lane + 32 * k] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int p = p0 + q * NW;
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (p < P && lane + 32 * k < S) Xt[(size_t)(lane + 32 * k) * kResLdx + p] = r[q][k];
    }
  }
  __syncthreads();

  // ---- xbar: one warp per band ------------------------------------------------------------------------------
  for (int s2 = warp; s2 < S; s2 += NW) {
    const float* row = Xt + (size_t)s2 * kResLdx;
    double acc = 0.0;
    for (int p = lane; p < P; p += 32) acc += (double)(row[p]);
    acc = warp_sum_all(acc);
    if (lane == 0) xbar[s2] = acc / N;
  }
  __syncthreads();

  // ---- C0 = X^T X / N - xbar xbar^T: one warp per 8 x 4 tile of the lower triangle, lanes over pixels --------
  {
    int t = 0;
    for (int rb = 0; rb < SP / 8; ++rb) {
      const int ncb = min(2 * rb + 2, SP / 4);
      for (int cb = 0; cb < ncb; ++cb, ++t) {
        if (t % NW != warp) continue;
        double acc[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] = 0.0;
        const float* ri = Xt + (size_t)(8 * rb) * kResLdx;
        const float* rj = Xt + (size_t)(4 * cb) * kResLdx;
        for (int p = lane; p < P; p += 32) {
          double xi[8], xj[4];
#pragma unroll
          for (int a = 0; a < 8; ++a) xi[a] = (double)(ri[(size_t)a * kResLdx + p]);
#pragma unroll
          for (int b = 0; b < 4; ++b) xj[b] = (double)(rj[(size_t)b * kResLdx + p]);
#pragma unroll
          for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a * 4 + b] = fma(xi[a], xj[b], acc[a * 4 + b]);
        }
        // 32 lanes x 32 partial sums -> lane e holds the total of entry e (recursive halving, 31 shuffles)
#pragma unroll
        for (int wdt = 16, mask = 16; wdt >= 1; wdt >>= 1, mask >>= 1) {
          const bool upper = (lane & mask) != 0;
#pragma unroll
          for (int e = 0; e < wdt; ++e) {
            const double send = upper ? acc[e] : acc[e + wdt];
            const double keep = upper ? acc[e + wdt] : acc[e];
            acc[e] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
          }
        }
        const int i = 8 * rb + (lane >> 2), j = 4 * cb + (lane & 3);
        if (i < S && j <= i) {
          const double c = acc[0] / N - xbar[i] * xbar[j];
          A[i * LDA + j] = c;
          A[j * LDA + i] = c;
        }
      }
    }
  }
  __syncthreads();

  // ---- P0 = C0^-1 by the symmetric sweep operator (Goodnight): sweeping every pivot of an SPD matrix leaves
  // -C0^-1.  Each thread keeps its <= 7 lower-triangle entries in REGISTERS for all S sweeps; per sweep only
  // the pivot row/column is published to shared memory (double buffered: one barrier per sweep).  A
  // non-positive pivot means C0 is not positive definite (the reference's Cholesky raises).
  {
    constexpr int kOwn = 7;                      // ceil(80*81/2 / 512)
    const int NT = S * (S + 1) / 2;
    double av[kOwn];
    int ei[kOwn], ej[kOwn];
#pragma unroll
    for (int m = 0; m < kOwn; ++m) {
      const int e = tid + kResThreads * m;
      int i = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
      while (i * (i + 1) / 2 > e) --i;
      while ((i + 1) * (i + 2) / 2 <= e) ++i;
      const int j = e - i * (i + 1) / 2;
      const bool valid = e < NT;
      ei[m] = valid ? i : -1;
      ej[m] = valid ? j : -1;
      av[m] = valid ? A[i * LDA + j] : 0.0;
    }
    for (int k = 0; k < S; ++k) {
      double* ck = (k & 1) ? colk : rowk;
#pragma unroll
      for (int m = 0; m < kOwn; ++m) {
        if (ej[m] == k) ck[ei[m]] = av[m];          // column k below (and on) the diagonal
        else if (ei[m] == k) ck[ej[m]] = av[m];     // row k left of the diagonal = column k above it
      }
      __syncthreads();
      double d = ck[k];
      if (!(d > 0.0)) {
        if (tid == 0) ok = 0;
        d = 1.0;
      }
      const double inv = 1.0 / d;
#pragma unroll
      for (int m = 0; m < kOwn; ++m) {
        const int i = ei[m], j = ej[m];
        if (i < 0) continue;
        if (i == k) av[m] = (j == k) ? -inv : ck[j] * inv;
        else if (j == k) av[m] = ck[i] * inv;
        else av[m] = fma(-ck[i] * inv, ck[j], av[m]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < kOwn; ++m) {
      if (ei[m] < 0) continue;
      A[ei[m] * LDA + ej[m]] = -av[m];
      A[ej[m] * LDA + ei[m]] = -av[m];
    }
  }
  __syncthreads();

  // ---- iterations ----------------------------------------------------------------------------------------------
  double sum_a = 0.0, sum_a2 = 0.0;       // block-uniform
  float R_f = 0.f, mf_f = 0.f;            // this thread's pixel, working values in the storage precision
  const int64_t my_pix = tid < P ? (int64_t)pidx[tid] : 0;
  for (int it = 0; it <= num_iter; ++it) {
    if (tid < S) {
      if (it == 0) {
        mu[tid] = xbar[tid];
        tcur[tid] = tp[tid] * xbar[tid];          // target0 = template * mean(x)   (mag1c.py:312, :233)
      } else {
        const double ma = sum_a / N;
        const double told = tcur[tid];
        tprev[tid] = told;
        pt[tid] = pb[tid];                          // P0 t: last iteration's P0 b
        const double m = xbar[tid] - ma * told;     // mean of modx = x - a t^T
        mu[tid] = m;
        wv[tid] = ma * xbar[tid] - v[tid] / N;
        tcur[tid] = tp[tid] * m;
      }
    }
    __syncthreads();
    // pb = P0 b (and pw = P0 w): four threads per matrix row
    {
      const int r = tid >> 2, sub = tid & 3;
      double sb = 0.0, sw = 0.0;
      if (r < S) {
        const double* Ar = A + r * LDA;
        for (int c = sub; c < S; c += 4) {
          const double aval = Ar[c];
          sb = fma(aval, tcur[c], sb);
          if (it > 0) sw = fma(aval, wv[c], sw);
        }
      }
      sb += __shfl_xor_sync(0xffffffffu, sb, 1);
      sb += __shfl_xor_sync(0xffffffffu, sb, 2);
      sw += __shfl_xor_sync(0xffffffffu, sw, 1);
      sw += __shfl_xor_sync(0xffffffffu, sw, 2);
      if (r < S && sub == 0) {
        pb[r] = sb;
        pw[r] = sw;
      }
    }
    __syncthreads();
    // warp 0: Woodbury 2 x 2 system, cit = S_it^-1 b, and the filter's scalars
    if (warp == 0) {
      double g11 = 0.0, g12 = 0.0, g22 = 0.0, r1 = 0.0, r2 = 0.0;
      if (it > 0) {
        for (int s2 = lane; s2 < S; s2 += 32) {
          g11 = fma(tprev[s2], pt[s2], g11);
          g12 = fma(tprev[s2], pw[s2], g12);
          g22 = fma(wv[s2], pw[s2], g22);
          r1 = fma(tprev[s2], pb[s2], r1);
          r2 = fma(wv[s2], pb[s2], r2);
        }
        g11 = warp_sum_all(g11); g12 = warp_sum_all(g12); g22 = warp_sum_all(g22);
        r1 = warp_sum_all(r1); r2 = warp_sum_all(r2);
      }
      double y1 = 0.0, y2 = 0.0;
      if (it > 0) {
        const double beta = sum_a2 / N - (sum_a / N) * (sum_a / N);
        const double m11 = g11, m12 = 1.0 + g12, m22 = g22 - beta;
        const double det = m11 * m22 - m12 * m12;
        y1 = (r1 * m22 - m12 * r2) / det;
        y2 = (m11 * r2 - m12 * r1) / det;
      }
      double nrm = 0.0, mucit = 0.0, mumu = 0.0;
      for (int s2 = lane; s2 < S; s2 += 32) {
        const double c = it > 0 ? pb[s2] - y1 * pt[s2] - y2 * pw[s2] : pb[s2];
        cit[s2] = c;
        nrm = fma(tcur[s2], c, nrm);
        mucit = fma(mu[s2], c, mucit);
        mumu = fma(xbar[s2], xbar[s2], mumu);
      }
      nrm = warp_sum_all(nrm); mucit = warp_sum_all(mucit); mumu = warp_sum_all(mumu);
      if (it > 0 && nrm < 1.0) nrm = 1.0;               // mag1c.py:264-266 (not applied inside rmf)
      if (lane == 0) {
        scal[0] = nrm;
        scal[1] = mucit;
        scal[2] = mumu;
      }
    }
    __syncthreads();
    // ---- matched-filter apply: one thread per pixel over the resident spectra ------------------------------------
    const bool last = it == num_iter;
    const double nrm = scal[0], mucit = scal[1], mumu = scal[2];
    double la = 0.0, la2 = 0.0;
    if (tid < P) {
      const float* xp = Xt + tid;
      double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
      int s2 = 0;
      for (; s2 + 4 <= S; s2 += 4) {
        // the filter coefficients are warp-uniform: two 16-byte broadcast loads per four bands
        const double2 c01 = *reinterpret_cast<const double2*>(cit + s2);
        const double2 c23 = *reinterpret_cast<const double2*>(cit + s2 + 2);
        d0 = fma((double)(xp[(size_t)s2 * kResLdx]), c01.x, d0);
        d1 = fma((double)(xp[(size_t)(s2 + 1) * kResLdx]), c01.y, d1);
        d2 = fma((double)(xp[(size_t)(s2 + 2) * kResLdx]), c23.x, d2);
        d3 = fma((double)(xp[(size_t)(s2 + 3) * kResLdx]), c23.y, d3);
      }
      for (; s2 < S; ++s2) d0 = fma((double)(xp[(size_t)s2 * kResLdx]), cit[s2], d0);
      const double dot = (d0 + d1) + (d2 + d3);
      double R, mf;
      if (it == 0) {
        double x0 = 0.0, x1 = 0.0;
        for (s2 = 0; s2 + 2 <= S; s2 += 2) {
          x0 = fma((double)(xp[(size_t)s2 * kResLdx]), xbar[s2], x0);
          x1 = fma((double)(xp[(size_t)(s2 + 1) * kResLdx]), xbar[s2 + 1], x1);
        }
        for (; s2 < S; ++s2) x0 = fma((double)(xp[(size_t)s2 * kResLdx]), xbar[s2], x0);
        R = (x0 + x1) / mumu;                            // mag1c.py:330
        mf = (dot - mucit) / (R * nrm);                  // mag1c.py:332
        R_f = (float)R;
        al_out[my_pix] = R_f;
      } else {
        R = (double)R_f;
        const double reg = 1.0 / (R * ((double)mf_f + kEpsilon));   // mag1c.py:255
        mf = ((dot - mucit) - reg) / (R * nrm);          // mag1c.py:267
      }
      mf = mf > 0.0 ? mf : 0.0;                          // relu
      mf_f = (float)mf;
      if (last) mf_out[my_pix] = (float)((double)mf_f * kScaling);
      const double a = (double)R_f * (double)mf_f;
      a_s[tid] = a;
      la = a;
      la2 = a * a;
    }
    if (last) break;
    // block sums of a and a^2
    la = warp_sum_all(la);
    la2 = warp_sum_all(la2);
    if (lane == 0) {
      scal[8 + warp] = la;
      scal[8 + NW + warp] = la2;
    }
    __syncthreads();                         // also publishes a_s
    sum_a = 0.0;
    sum_a2 = 0.0;
#pragma unroll
    for (int wq = 0; wq < NW; ++wq) {
      sum_a += scal[8 + wq];
      sum_a2 += scal[8 + NW + wq];
    }
    // v = X^T a: one warp per band; each lane keeps its 16 pixels' a in registers across the warp's bands
    {
      double ar[kResMaxP / 32];
#pragma unroll
      for (int k = 0; k < kResMaxP / 32; ++k) ar[k] = lane + 32 * k < P ? a_s[lane + 32 * k] : 0.0;
      for (int s2 = warp; s2 < S; s2 += NW) {
        const float* row = Xt + (size_t)s2 * kResLdx + lane;
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int k = 0; k < kResMaxP / 32; k += 2) {
          acc0 = fma((double)row[32 * k], ar[k], acc0);             // pad pixels are zero-filled and meet a = 0
          acc1 = fma((double)row[32 * (k + 1)], ar[k + 1], acc1);
        }
        const double acc = warp_sum_all(acc0 + acc1);
        if (lane == 0) v[s2] = acc;
      }
    }
    __syncthreads();
  }
  if (tid == 0 && !ok && status) atomicAdd(status, 1);
}

}  // namespace

extern "C" int64_t sc_mag1c_smem_bytes(int S, int pmax, int elem_bytes) {
  int64_t NT = (int64_t)S * (S + 1) / 2;
  return (NT + (int64_t)(S + 1) * mag1c_ld(S) + 8 * S + 8 + kTile) * 8 + 2 * (int64_t)kTile * S * elem_bytes + 2 * NT + 32 + 4 * (int64_t)(pmax + 2);
}

extern "C" int sc_mag1c_filter(const void* x, int64_t pixel_stride, const int32_t* pix_idx, const int32_t* counts,
                               int pmax, const double* tmpl, void* mf_out, void* albedo_out, int G, int S,
                               int num_iter, double alpha, int fp64, int skip_le, int* status, void* stream) {
  if (skip_le < 1) skip_le = 1;                        // a single pixel has no covariance
  if (!x || !pix_idx || !tmpl || !mf_out || !albedo_out || G <= 0 || S < 2 || S + 1 > kThreads || pmax < 1 || num_iter < 0)
    return SC_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  // group-resident fast paths: fp32 radiance, no diagonal loading, <= 512 pixels per group (the AVIRIS case)
  if (!fp64 && alpha == 0.0 && pmax <= kResMaxP && !getenv("STARCOP_MAG1C_STREAMING") && !getenv("STARCOP_MAG1C_NO_TC")) {
    const int r = mag1c_tc_launch((const float*)x, pixel_stride, pix_idx, counts, pmax, tmpl, (float*)mf_out,
                                  (float*)albedo_out, G, S, num_iter, skip_le, status, st);
    if (r != SC_ERR_UNSUPPORTED) return r;
  }
  if (!fp64 && alpha == 0.0 && pmax <= kResMaxP && res_smem_bytes(S) <= 227 * 1024 && !getenv("STARCOP_MAG1C_STREAMING")) {
    const size_t rs = res_smem_bytes(S);
    e = cudaFuncSetAttribute(mag1c_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs);
    if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }
    mag1c_resident_kernel<<<G, kResThreads, rs, st>>>((const float*)x, pixel_stride, pix_idx, counts, pmax, tmpl,
                                                      (float*)mf_out, (float*)albedo_out, S, num_iter, skip_le, status);
    return check_launch();
  }
  size_t smem = (size_t)sc_mag1c_smem_bytes(S, pmax, fp64 ? 8 : 4);
  if (smem > 220 * 1024) return SC_ERR_UNSUPPORTED;
  if (fp64) {
    e = cudaFuncSetAttribute(mag1c_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }
    mag1c_kernel<double><<<G, kThreads, smem, st>>>((const double*)x, pixel_stride, pix_idx, counts, pmax, tmpl,
                                                    (double*)mf_out, (double*)albedo_out, S, num_iter, alpha, skip_le, status);
  } else {
    e = cudaFuncSetAttribute(mag1c_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }
    mag1c_kernel<float><<<G, kThreads, smem, st>>>((const float*)x, pixel_stride, pix_idx, counts, pmax, tmpl,
                                                   (float*)mf_out, (float*)albedo_out, S, num_iter, alpha, skip_le, status);
  }
  return check_launch();
}
