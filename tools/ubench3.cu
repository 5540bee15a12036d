// ubench3.cu -- shared-memory load-only / store-only / RMW bandwidth per SM (development tool)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MODE 0: LDS.64 only, 1: STS.64 only, 2: LDS.64 + STS.64 (no fma dependency), 3: LDS.128 only, 4: STS.128 only
template <int MODE>
__global__ void k_bw(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 24576; i += blockDim.x) sm[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int W = (MODE >= 3) ? 16 : 8;
  const uint32_t base = smem_u32(sm) + (warp % 16) * 12288 / 2 + lane * W;  // 6 KB window per warp
  unsigned long long acc = 0, x = lane, y = warp;
  for (int it = 0; it < iters; ++it) {
    const uint32_t a = base + (it & 3) * 1024;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t ad = a + (MODE >= 3 ? (u & 1) * 512 : u * 256);
      if (MODE == 0 || MODE == 2) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(x) : "r"(ad));
      if (MODE == 1 || MODE == 2) asm volatile("st.shared.b64 [%0], %1;" ::"r"(ad), "l"(y) : "memory");
      if (MODE == 3) asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(ad));
      if (MODE == 4) asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(ad), "l"(x), "l"(y) : "memory");
    }
    acc ^= x;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (double)acc;
}
template <class F>
static float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <int MODE>
static void run(double* out, int sms, double clk, const char* name, int bytes) {
  const int smem = 24576 * 8, iters = 8000;
  cudaFuncSetAttribute(k_bw<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int threads : {256, 512, 1024}) {
    float ms = time_ms([&] { k_bw<MODE><<<sms, threads, smem>>>(out, iters); });
    const double instr = (threads / 32) * 4.0 * iters;
    printf("%-10s warps=%2d: %.2f clk per warp instr, %.0f B/clk/SM\n", name, threads / 32, ms * 1e-3 * clk / instr,
           instr * bytes / (ms * 1e-3 * clk));
  }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = khz * 1e3; const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
  run<0>(out, sms, clk, "lds64", 256);
  run<1>(out, sms, clk, "sts64", 256);
  run<2>(out, sms, clk, "lds+sts64", 512);
  run<3>(out, sms, clk, "lds128", 512);
  run<4>(out, sms, clk, "sts128", 512);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
