#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(long long* out, long long* clk, int iters, int seed) {
  __shared__ int sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (i * seed) & 0xFFFFFF;
  __syncthreads();
  long long a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  int x0 = seed + threadIdx.x, x1 = x0 * 3, x2 = x0 * 5, x3 = x0 * 7;
  const int c0 = seed * 11 + 1, c1 = seed * 13 + 2;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {        // 8 independent IMAD.WIDE (32 x 32 + 64)
      a0 += (long long)x0 * c0; a1 += (long long)x1 * c0; a2 += (long long)x2 * c0; a3 += (long long)x3 * c0;
      a4 += (long long)x0 * c1; a5 += (long long)x1 * c1; a6 += (long long)x2 * c1; a7 += (long long)x3 * c1;
      x0 += 1; x1 += 1; x2 += 1; x3 += 1;
    } else if (MODE == 1) { // 4 x (LDS + 2 IMAD.WIDE): the integer apply pass
      const int b = (threadIdx.x * 73 + i * 4) & 4095;
      const int q0 = sm[b], q1 = sm[(b + 1) & 4095], q2 = sm[(b + 2) & 4095], q3 = sm[(b + 3) & 4095];
      a0 += (long long)q0 * c0; a1 += (long long)q0 * c1; a2 += (long long)q1 * c0; a3 += (long long)q1 * c1;
      a4 += (long long)q2 * c0; a5 += (long long)q2 * c1; a6 += (long long)q3 * c0; a7 += (long long)q3 * c1;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + x0 + x1 + x2 + x3;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[MODE] = t1 - t0;
}
int main() {
  long long *out, *clk;
  cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&clk, 64);
  const int iters = 4096;
  k<0><<<148, 512>>>(out, clk, iters, 3); k<1><<<148, 512>>>(out, clk, iters, 3);
  long long h[8]; cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
  printf("IMAD.WIDE x8            %lld cycles -> %.1f IMAD.WIDE/clk/SM\n", h[0], 8.0 * 512 * iters / h[0]);
  printf("LDS + 2 IMAD.WIDE x4    %lld cycles -> %.1f elements/clk/SM\n", h[1], 4.0 * 512 * iters / h[1]);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
