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
#include "common.cuh"

using namespace sc;

namespace {

constexpr int kThreads = 128;
constexpr int kChunk = 32;          // pixels staged per shared-memory tile when accumulating X^T X
constexpr double kScaling = 1e5;    // mag1c.py:56
constexpr double kEpsilon = 1e-9;   // mag1c.py:57

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // j <= i

__device__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < kThreads / 32; ++i) s += red[i];
  return s;
}

// in-place packed Cholesky (lower), returns false through *ok when a pivot is not positive
__device__ void cholesky_packed(double* A, int S, int* ok, const uchar2* __restrict__ rc) {
  for (int j = 0; j < S; ++j) {
    __syncthreads();
    double d = A[tri(j, j)];
    if (threadIdx.x == 0 && !(d > 0.0)) *ok = 0;
    double ljj = sqrt(d > 0.0 ? d : 1.0);
    __syncthreads();
    // scale column j
    for (int i = j + threadIdx.x; i < S; i += kThreads) A[tri(i, j)] = (i == j) ? ljj : A[tri(i, j)] / ljj;
    __syncthreads();
    // rank-1 update of the trailing lower triangle: A[i][k] -= L[i][j] * L[k][j], j < k <= i
    const int n = S - j - 1;
    const int cnt = n * (n + 1) / 2;
    for (int e = threadIdx.x; e < cnt; e += kThreads) {
      // e -> (r, c) with c <= r in the n x n trailing triangle (a prefix of the packed enumeration)
      const uchar2 q = rc[e];
      int i = j + 1 + q.x, k = j + 1 + q.y;
      A[tri(i, k)] -= A[tri(i, j)] * A[tri(k, j)];
    }
  }
  __syncthreads();
}

// solve L L^T c = b (packed L), result in c; z is scratch.  One warp does the substitutions.
__device__ void chol_solve(const double* L, const double* b, double* z, double* c, int S) {
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    for (int i = 0; i < S; ++i) {           // forward: z_i = (b_i - sum_{k<i} L_ik z_k) / L_ii
      double s = 0.0;
      for (int k = lane; k < i; k += 32) s += L[tri(i, k)] * z[k];
      s = warp_sum(s);
      if (lane == 0) z[i] = (b[i] - s) / L[tri(i, i)];
      __syncwarp();
    }
    for (int i = S - 1; i >= 0; --i) {      // backward: c_i = (z_i - sum_{k>i} L_ki c_k) / L_ii
      double s = 0.0;
      for (int k = i + 1 + lane; k < S; k += 32) s += L[tri(k, i)] * c[k];
      s = warp_sum(s);
      if (lane == 0) c[i] = (z[i] - s) / L[tri(i, i)];
      __syncwarp();
    }
  }
  __syncthreads();
}

template <typename TS>
__global__ void __launch_bounds__(kThreads)
mag1c_kernel(const TS* __restrict__ x, int64_t pixel_stride, const int32_t* __restrict__ pix_idx,
             const int32_t* __restrict__ counts, int pmax, const double* __restrict__ tmpl, TS* __restrict__ mf_out,
             TS* __restrict__ al_out, int S, int num_iter, double alpha, int* __restrict__ status) {
  extern __shared__ double smd[];
  const int g = blockIdx.x;
  const int P = counts ? counts[g] : pmax;
  if (P <= 10) return;                                 // mag1c.py:166-168: too few pixels, outputs stay NODATA
  const int32_t* idx = pix_idx + (int64_t)g * pmax;
  const int NT = S * (S + 1) / 2;
  double* XtX = smd;                 // packed lower triangle of sum x x^T
  double* A = XtX + NT;              // working covariance / Cholesky factor
  double* xbar = A + NT;
  double* mu = xbar + S;
  double* tprev = mu + S;            // target used inside modx
  double* tcur = tprev + S;          // target = template * mu
  double* v = tcur + S;              // X^T a
  double* cit = v + S;
  double* z = cit + S;
  double* tp = z + S;                // template
  double* red = tp + S;              // [8] block-reduction scratch
  double* tiled = red + 8;                             // [kChunk][S] staging tile (fp64 view)
  float* tile = reinterpret_cast<float*>(tiled);       // same storage, fp32 view
  uchar2* rc = reinterpret_cast<uchar2*>(tiled + kChunk * S);   // packed index -> (row, col)
  __shared__ int ok;
  const int tid = threadIdx.x;
  constexpr bool kF64 = sizeof(TS) == 8;

  if (tid == 0) ok = 1;
  for (int i = tid; i < S; i += kThreads)
    for (int j = 0; j <= i; ++j) rc[tri(i, j)] = make_uchar2((unsigned char)i, (unsigned char)j);
  for (int i = tid; i < NT; i += kThreads) XtX[i] = 0.0;
  for (int i = tid; i < S; i += kThreads) {
    tp[i] = tmpl[i];
    xbar[i] = 0.0;
  }
  __syncthreads();

  // ---- pass 0: xbar and X^T X, pixels staged kChunk at a time -------------------------------------
  for (int p0 = 0; p0 < P; p0 += kChunk) {
    const int np = min(kChunk, P - p0);
    __syncthreads();
    for (int e = tid; e < np * S; e += kThreads) {
      int pp = e / S, s = e - pp * S;
      TS val = x[(int64_t)idx[p0 + pp] * pixel_stride + s];
      if (kF64) tiled[pp * S + s] = (double)val;
      else tile[pp * S + s] = (float)val;
    }
    __syncthreads();
    for (int e = tid; e < NT; e += kThreads) {
      const int i = rc[e].x, j = rc[e].y;
      double acc = 0.0;
      if (kF64) {
        for (int pp = 0; pp < np; ++pp) acc += tiled[pp * S + i] * tiled[pp * S + j];
      } else {
        for (int pp = 0; pp < np; ++pp) acc += (double)tile[pp * S + i] * (double)tile[pp * S + j];
      }
      XtX[e] += acc;
    }
    for (int s = tid; s < S; s += kThreads) {
      double acc = 0.0;
      if (kF64) {
        for (int pp = 0; pp < np; ++pp) acc += tiled[pp * S + s];
      } else {
        for (int pp = 0; pp < np; ++pp) acc += (double)tile[pp * S + s];
      }
      xbar[s] += acc;
    }
  }
  __syncthreads();
  const double N = (double)P;
  for (int s = tid; s < S; s += kThreads) {
    xbar[s] /= N;
    mu[s] = xbar[s];
    tcur[s] = tp[s] * xbar[s];      // target0 = template * mean(x)   (mag1c.py:312, :233)
    tprev[s] = 0.0;
    v[s] = 0.0;
  }
  __syncthreads();
  double mumu = 0.0;
  for (int s = 0; s < S; ++s) mumu += xbar[s] * xbar[s];   // every thread: mu.mu for the albedo factor

  double sum_a = 0.0, sum_a2 = 0.0;       // block-uniform after block_sum
  for (int it = 0; it <= num_iter; ++it) {
    // ---- covariance of modx, C = (1-alpha) S + alpha diag(S) ------------------------------------
    if (it > 0) {
      const double ma = sum_a / N;
      for (int s = tid; s < S; s += kThreads) {
        tprev[s] = tcur[s];
        mu[s] = xbar[s] - ma * tcur[s];
      }
      __syncthreads();
      for (int s = tid; s < S; s += kThreads) tcur[s] = tp[s] * mu[s];
      __syncthreads();
    }
    for (int e = tid; e < NT; e += kThreads) {
      const int i = rc[e].x, j = rc[e].y;
      double sxx = XtX[e];
      if (it > 0) sxx += -v[i] * tprev[j] - tprev[i] * v[j] + sum_a2 * tprev[i] * tprev[j];
      double c = sxx / N - mu[i] * mu[j];
      if (i != j) c *= (1.0 - alpha);
      A[e] = c;
    }
    cholesky_packed(A, S, &ok, rc);
    chol_solve(A, tcur, z, cit, S);
    double nrm = 0.0, mucit = 0.0;
    for (int s = 0; s < S; ++s) {
      nrm += tcur[s] * cit[s];
      mucit += mu[s] * cit[s];
    }
    if (it > 0 && nrm < 1.0) nrm = 1.0;               // mag1c.py:264-266 (not applied inside rmf)
    // ---- apply: one thread per pixel -------------------------------------------------------------
    double la = 0.0, la2 = 0.0;
    for (int p = tid; p < P; p += kThreads) {
      const int64_t pix = idx[p];
      const TS* xp = x + pix * pixel_stride;
      double dot = 0.0, xmu = 0.0;
      for (int s = 0; s < S; ++s) {
        double xv = (double)xp[s];
        dot += xv * cit[s];
        if (it == 0) xmu += xv * xbar[s];
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
      if (it == 0) al_out[pix] = (TS)R;
      // keep the working value in the storage precision, like the reference's tensors
      TS mfs = (TS)mf;
      mf_out[pix] = it == num_iter ? (TS)((double)mfs * kScaling) : mfs;
      double a = (double)(TS)R * (double)mfs;
      la += a;
      la2 += a * a;
    }
    if (it == num_iter) break;
    sum_a = block_sum(la, red);
    sum_a2 = block_sum(la2, red);
    __syncthreads();                                   // mf_out / al_out of this block visible below
    // ---- v = X^T a: one thread per band, pixels streamed (x rows are L1/L2 resident) ---------------
    for (int s = tid; s < S; s += kThreads) {
      double acc = 0.0;
      for (int p = 0; p < P; ++p) {
        const int64_t pix = idx[p];
        double a = (double)al_out[pix] * (double)mf_out[pix];
        acc += a * (double)x[pix * pixel_stride + s];
      }
      v[s] = acc;
    }
    __syncthreads();
  }
  if (tid == 0 && !ok && status) atomicAdd(status, 1);
}

}  // namespace

extern "C" int64_t sc_mag1c_smem_bytes(int S) {
  int64_t NT = (int64_t)S * (S + 1) / 2;
  return (2 * NT + 8 * S + 8) * 8 + (int64_t)kChunk * S * 8 + 2 * NT + 16;
}

extern "C" int sc_mag1c_filter(const void* x, int64_t pixel_stride, const int32_t* pix_idx, const int32_t* counts,
                               int pmax, const double* tmpl, void* mf_out, void* albedo_out, int G, int S,
                               int num_iter, double alpha, int fp64, int* status, void* stream) {
  if (!x || !pix_idx || !tmpl || !mf_out || !albedo_out || G <= 0 || S < 2 || S > 160 || pmax < 1 || num_iter < 0)
    return SC_ERR_BAD_ARG;
  size_t smem = (size_t)sc_mag1c_smem_bytes(S);
  if (smem > 220 * 1024) return SC_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (fp64) {
    e = cudaFuncSetAttribute(mag1c_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }
    mag1c_kernel<double><<<G, kThreads, smem, st>>>((const double*)x, pixel_stride, pix_idx, counts, pmax, tmpl,
                                                    (double*)mf_out, (double*)albedo_out, S, num_iter, alpha, status);
  } else {
    e = cudaFuncSetAttribute(mag1c_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { g_last_error = e; return SC_ERR_CUDA; }
    mag1c_kernel<float><<<G, kThreads, smem, st>>>((const float*)x, pixel_stride, pix_idx, counts, pmax, tmpl,
                                                   (float*)mf_out, (float*)albedo_out, S, num_iter, alpha, status);
  }
  return check_launch();
}
