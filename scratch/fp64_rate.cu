#include <cstdio>
#include <cuda_runtime.h>
// measure DFMA / DADD / F2F.F64.F32 / LDS.32 / I2F throughput per SM per clock with 16 warps x 512 threads, 1 CTA per SM
template <int MODE>
__global__ void k(double* out, long long* clk, int iters, float fseed) {
  __shared__ float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (float)i * fseed;
  __syncthreads();
  double a0 = threadIdx.x, a1 = 1.0 + a0, a2 = 2.0 + a0, a3 = 3.0 + a0, a4 = 4. + a0, a5 = 5. + a0, a6 = 6. + a0, a7 = 7. + a0;
  const double c = 1.0000001, d = 1e-9;
  float f0 = fseed * threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {  // 8 independent DFMA
      a0 = fma(a0, c, d); a1 = fma(a1, c, d); a2 = fma(a2, c, d); a3 = fma(a3, c, d);
      a4 = fma(a4, c, d); a5 = fma(a5, c, d); a6 = fma(a6, c, d); a7 = fma(a7, c, d);
    } else if (MODE == 1) {  // 8 DADD
      a0 += d; a1 += d; a2 += d; a3 += d; a4 += d; a5 += d; a6 += d; a7 += d;
    } else if (MODE == 2) {  // 4 x (F2F + DFMA)
      a0 = fma((double)f0, c, a0); a1 = fma((double)f1, c, a1); a2 = fma((double)f2, c, a2); a3 = fma((double)f3, c, a3);
      f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f;
    } else if (MODE == 3) {  // 4 x (LDS + magic DADD + DFMA)
      int b = (threadIdx.x * 73 + i * 4) & 4095;
      a0 = fma(__hiloint2double(0x43300000, __float_as_int(sm[b])) - 4503599627370496.0, c, a0);
      a1 = fma(__hiloint2double(0x43300000, __float_as_int(sm[(b + 1) & 4095])) - 4503599627370496.0, c, a1);
      a2 = fma(__hiloint2double(0x43300000, __float_as_int(sm[(b + 2) & 4095])) - 4503599627370496.0, c, a2);
      a3 = fma(__hiloint2double(0x43300000, __float_as_int(sm[(b + 3) & 4095])) - 4503599627370496.0, c, a3);
    } else if (MODE == 4) {  // dependent DFMA chain (latency)
      a0 = fma(a0, c, d); a0 = fma(a0, c, d); a0 = fma(a0, c, d); a0 = fma(a0, c, d);
      a0 = fma(a0, c, d); a0 = fma(a0, c, d); a0 = fma(a0, c, d); a0 = fma(a0, c, d);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[MODE] = t1 - t0;
}
int main() {
  double* out; long long* clk;
  cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&clk, 64);
  const int iters = 4096;
  k<0><<<148, 512>>>(out, clk, iters, 1.f); k<1><<<148, 512>>>(out, clk, iters, 1.f); k<2><<<148, 512>>>(out, clk, iters, 1.f);
  k<3><<<148, 512>>>(out, clk, iters, 1.f); k<4><<<148, 512>>>(out, clk, iters, 1.f);
  long long h[8]; cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
  const char* names[] = {"DFMA x8", "DADD x8", "F2F+DFMA x4", "LDS+DADD+DFMA x4", "dependent DFMA x8"};
  const double ops[] = {8, 8, 4, 4, 8};
  for (int m = 0; m < 5; ++m)
    printf("%-20s %lld cycles  -> %.1f groups/clk/SM, %.2f cycles per warp-group-iteration\n", names[m], h[m],
           ops[m] * 512.0 * iters / h[m], (double)h[m] / iters);
  for (int w = 1; w <= 16; w *= 2) {
    k<0><<<148, 32 * w>>>(out, clk, iters, 1.f); cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
    printf("DFMA x8, %2d warps: %.1f /clk/SM\n", w, 8.0 * 32 * w * iters / h[0]);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
