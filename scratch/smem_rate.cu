#include <cstdio>
#include <cuda_runtime.h>
// shared-memory access patterns of the mag1c per-pixel pass: 512 x 73 words pixel-major + 80 double coefficients
constexpr int S = 73, PITCH = 73, P = 512;
constexpr double kM = 4503599627370496.0 + 8421504.0;
__device__ __forceinline__ double u2d(unsigned u) { return __hiloint2double(0x43300000, (int)u) - kM; }
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, long long* clk, int reps) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* cs = reinterpret_cast<double*>(smraw);
  unsigned* q = reinterpret_cast<unsigned*>(smraw + 1024);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 128; i += 512) cs[i] = (1.0 + i * 1e-3) * 1e270;
  for (int i = tid; i < P * PITCH; i += 512) q[i] = 0x808080 + (i * 2654435761u >> 12) % 1000;
  __syncthreads();
  double acc = 0.0;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (MODE == 0) {          // thread per pixel, coefficients by LDS.128 broadcast
      const unsigned* row = q + tid * PITCH;
      double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
      int s = 0;
      for (; s + 4 <= S; s += 4) {
        const double2 c01 = *reinterpret_cast<const double2*>(cs + s), c23 = *reinterpret_cast<const double2*>(cs + s + 2);
        d0 = fma(u2d(row[s]), c01.x, d0); d1 = fma(u2d(row[s + 1]), c01.y, d1);
        d2 = fma(u2d(row[s + 2]), c23.x, d2); d3 = fma(u2d(row[s + 3]), c23.y, d3);
      }
      for (; s < S; ++s) d0 = fma(u2d(row[s]), cs[s], d0);
      acc += (d0 + d1) + (d2 + d3);
    } else if (MODE == 5) {   // thread per pixel, coefficients by LDS.64 broadcast
      const unsigned* row = q + tid * PITCH;
      double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
      int s = 0;
      for (; s + 4 <= S; s += 4) {
        d0 = fma(u2d(row[s]), cs[s], d0); d1 = fma(u2d(row[s + 1]), cs[s + 1], d1);
        d2 = fma(u2d(row[s + 2]), cs[s + 2], d2); d3 = fma(u2d(row[s + 3]), cs[s + 3], d3);
      }
      for (; s < S; ++s) d0 = fma(u2d(row[s]), cs[s], d0);
      acc += (d0 + d1) + (d2 + d3);
    } else if (MODE == 6) {   // 256 threads x 2 pixels, LDS.128 coefficients
      if (tid < 256) {
        const unsigned* row = q + tid * PITCH;
        double d[4] = {0, 0, 0, 0};
        int s = 0;
        for (; s + 2 <= S; s += 2) {
          const double2 c01 = *reinterpret_cast<const double2*>(cs + s);
          d[0] = fma(u2d(row[s]), c01.x, d[0]);
          d[1] = fma(u2d(row[s + 1]), c01.y, d[1]);
          d[2] = fma(u2d(row[256 * PITCH + s]), c01.x, d[2]);
          d[3] = fma(u2d(row[256 * PITCH + s + 1]), c01.y, d[3]);
        }
        acc += (d[0] + d[1]) + (d[2] + d[3]);
      }
    } else if (MODE == 7) {   // v pass: warp per band GROUP, bands interleaved (independent chains + batched shuffles)
      double ar[16];
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) ar[kk] = cs[(lane + kk) & 63];
      double a[5] = {0, 0, 0, 0, 0};
      const unsigned* col = q + lane * PITCH + warp;
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
#pragma unroll
        for (int b = 0; b < 5; ++b)
          if (warp + 16 * b < S) a[b] = fma(u2d(col[32 * kk * PITCH + 16 * b]), ar[kk], a[b]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int b = 0; b < 5; ++b) a[b] += __shfl_xor_sync(0xffffffffu, a[b], o);
      acc += a[0] + a[1] + a[2] + a[3] + a[4];
    } else if (MODE == 8) {   // thread per pixel, denormal trick (no DADD), LDS.128 coef
      const unsigned* row = q + tid * PITCH;
      double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
      int s = 0;
      for (; s + 4 <= S; s += 4) {
        const double2 c01 = *reinterpret_cast<const double2*>(cs + s), c23 = *reinterpret_cast<const double2*>(cs + s + 2);
        d0 = fma(__hiloint2double(0, row[s]), c01.x, d0); d1 = fma(__hiloint2double(0, row[s + 1]), c01.y, d1);
        d2 = fma(__hiloint2double(0, row[s + 2]), c23.x, d2); d3 = fma(__hiloint2double(0, row[s + 3]), c23.y, d3);
      }
      acc += (d0 + d1) + (d2 + d3);
    } else if (MODE == 9) {   // v pass interleaved, denormal trick
      double ar[16];
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) ar[kk] = cs[(lane + kk) & 63];
      double a[5] = {0, 0, 0, 0, 0};
      const unsigned* col = q + lane * PITCH + warp;
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
#pragma unroll
        for (int b = 0; b < 5; ++b)
          if (warp + 16 * b < S) a[b] = fma(__hiloint2double(0, col[32 * kk * PITCH + 16 * b]), ar[kk], a[b]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int b = 0; b < 5; ++b) a[b] += __shfl_xor_sync(0xffffffffu, a[b], o);
      acc += a[0] + a[1] + a[2] + a[3] + a[4];
    } else if (MODE == 10) {  // denormal DFMA rate without memory: 8 chains
      double t = __hiloint2double(0, tid + 1000 + r);
      double d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < 73; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = fma(t, cs[j], d[j]);
      }
      acc += d[0] + d[1] + d[2] + d[3] + d[4] + d[5] + d[6] + d[7];
    } else if (MODE == 11) {  // normal DFMA rate, same shape
      double t = (double)(tid + 1000 + r);
      double d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < 73; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = fma(t, cs[j], d[j]);
      }
      acc += d[0] + d[1] + d[2] + d[3] + d[4] + d[5] + d[6] + d[7];
    } else if (MODE == 1) {   // only the data loads (thread per pixel), coefficient = constant
      const unsigned* row = q + tid * PITCH;
      double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
      int s = 0;
      for (; s + 4 <= S; s += 4) {
        d0 = fma(u2d(row[s]), 1.5, d0); d1 = fma(u2d(row[s + 1]), 1.25, d1);
        d2 = fma(u2d(row[s + 2]), 1.125, d2); d3 = fma(u2d(row[s + 3]), 1.0625, d3);
      }
      acc += (d0 + d1) + (d2 + d3);
    } else if (MODE == 2) {   // 128 threads x 4 pixels, coefficients by LDS.128 broadcast
      if (tid < 128) {
        const unsigned* row = q + tid * PITCH;
        double d[4] = {0, 0, 0, 0}, e[4] = {0, 0, 0, 0};
        int s = 0;
        for (; s + 2 <= S; s += 2) {
          const double2 c01 = *reinterpret_cast<const double2*>(cs + s);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            d[j] = fma(u2d(row[j * 128 * PITCH + s]), c01.x, d[j]);
            e[j] = fma(u2d(row[j * 128 * PITCH + s + 1]), c01.y, e[j]);
          }
        }
        acc += (d[0] + d[1]) + (d[2] + d[3]) + (e[0] + e[1]) + (e[2] + e[3]);
      }
    } else if (MODE == 3) {   // lanes over bands, warp per 32 pixels, coefficients in registers, transpose-free: shuffle reduce per pixel
      const double c0 = cs[lane], c1 = cs[lane + 32], c2 = lane + 64 < S ? cs[lane + 64] : 0.0;
      for (int p = warp * 32; p < warp * 32 + 32; ++p) {
        const unsigned* row = q + p * PITCH;
        double d = u2d(row[lane]) * c0;
        d = fma(u2d(row[lane + 32]), c1, d);
        if (lane + 64 < S) d = fma(u2d(row[lane + 64]), c2, d);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        acc += d;
      }
    } else if (MODE == 4) {   // v pass: warp per band, lanes over pixels
      double ar[16];
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) ar[kk] = cs[(lane + kk) & 63];
      for (int s = warp; s < S; s += 16) {
        const unsigned* col = q + lane * PITCH + s;
        double a0 = 0, a1 = 0;
#pragma unroll
        for (int kk = 0; kk < 16; kk += 2) {
          a0 = fma(u2d(col[32 * kk * PITCH]), ar[kk], a0);
          a1 = fma(u2d(col[32 * (kk + 1) * PITCH]), ar[kk + 1], a1);
        }
        double d = a0 + a1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        acc += d;
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  out[blockIdx.x * 512 + tid] = acc;
  if (tid == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  double* out; long long* clk;
  cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&clk, 64);
  const int smem = 1024 + P * PITCH * 4, reps = 64;
  const char* names[] = {"thread/pixel LDS.128 coef", "thread/pixel data only", "128 thr x 4 px", "lanes over bands + shfl", "v pass warp/band", "thread/pixel LDS.64 coef", "256 thr x 2 px", "v pass interleaved bands", "thread/pixel denormal", "v pass interleaved denormal", "584 DFMA/thread denormal operand", "584 DFMA/thread normal"};
  long long h;
#define RUN(M) cudaFuncSetAttribute(k<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<M><<<148, 512, smem>>>(out, clk, reps); \
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("%-28s %.0f cycles per pass\n", names[M], (double)h / reps);
  RUN(0) RUN(1) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11)
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
