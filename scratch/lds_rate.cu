#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(unsigned* out, long long* clk, int reps) {
  extern __shared__ __align__(16) unsigned q[];
  const int tid = threadIdx.x;
  for (int i = tid; i < 512 * 76; i += 512) q[i] = i * 2654435761u;
  __syncthreads();
  unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (MODE == 0) {  // LDS.32, lane stride 73 words, 72 loads per thread
      const unsigned* row = q + tid * 73;
#pragma unroll 8
      for (int s = 0; s < 72; s += 4) { a0 += row[s]; a1 += row[s + 1]; a2 += row[s + 2]; a3 += row[s + 3]; }
    } else if (MODE == 1) {  // LDS.32, consecutive lanes consecutive words
      const unsigned* row = q + tid;
#pragma unroll 8
      for (int s = 0; s < 72; s += 4) { a0 += row[s * 512]; a1 += row[(s + 1) * 512]; a2 += row[(s + 2) * 512]; a3 += row[(s + 3) * 512]; }
    } else if (MODE == 2) {  // LDS.128, lane stride 76 words, 18 loads per thread
      const uint4* row = reinterpret_cast<const uint4*>(q + tid * 76);
#pragma unroll 6
      for (int s = 0; s < 18; ++s) { uint4 v = row[s]; a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w; }
    } else if (MODE == 3) {  // LDS.64 lane stride 74 words
      const uint2* row = reinterpret_cast<const uint2*>(q + tid * 74);
#pragma unroll 6
      for (int s = 0; s < 36; s += 2) { uint2 v = row[s]; uint2 w = row[s + 1]; a0 += v.x; a1 += v.y; a2 += w.x; a3 += w.y; }
    } else if (MODE == 4) {  // LDS.128 consecutive lanes consecutive 16 B
      const uint4* row = reinterpret_cast<const uint4*>(q) + tid;
#pragma unroll 6
      for (int s = 0; s < 18; ++s) { uint4 v = row[s * 512]; a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w; }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  out[blockIdx.x * 512 + tid] = a0 + a1 + a2 + a3;
  if (tid == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  unsigned* out; long long* clk;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 64);
  const int smem = 512 * 76 * 4, reps = 64;
  const char* names[] = {"LDS.32 stride 73", "LDS.32 consecutive", "LDS.128 stride 76", "LDS.64 stride 74", "LDS.128 consecutive"};
  long long h;
#define RUN(M) cudaFuncSetAttribute(k<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<M><<<148, 512, smem>>>(out, clk, reps); \
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("%-22s %.0f cycles per 147 KB pass -> %.1f B/clk\n", names[M], (double)h / reps, 512.0 * 72 * 4 * reps / h);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
