// ubench.cu -- calibration micro-benchmarks for the Gram pass design (development tool).
// Measures, per SM and clock: fp64 FMA issue rate from registers, DMMA m8n8k4 rate, the
// shared-memory read-modify-write rate (LDS.64 + DFMA + STS.64, conflict-free and data-dependent
// addresses), the 128-bit variant and fp64 shared-memory atomics.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void k_dfma(double* out, int iters) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double* out, int iters) {
  double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// RMW: every lane owns distinct addresses; idx stream comes from shared memory (data dependent)
template <int MODE>
__global__ void k_rmw(double* out, int iters) {
  extern __shared__ double sm[];
  const int nacc = 16384;  // 128 KB of accumulators
  uint32_t* idx = reinterpret_cast<uint32_t*>(sm + nacc);  // 1024 offsets
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) sm[i] = 0.0;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (uint32_t)((i * 37) % 256);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // each warp owns a private region of 512 doubles: lane l, step s -> region + 32 * off + lane
  double* reg = sm + warp * 512;
  const double v = 1.0 + lane;
  for (int it = 0; it < iters; ++it) {
    const uint32_t o = idx[(it * 8) & 1023];  // broadcast load
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int off = ((o + u * 5) & 15);
      if (MODE == 0) {
        double* p = reg + 32 * off + lane;
        *p = fma(v, 1.5, *p);
      } else if (MODE == 1) {
        double2* p = reinterpret_cast<double2*>(reg + 32 * (off & 7) * 2) + lane;
        double2 x = *p;
        x.x = fma(v, 1.5, x.x);
        x.y = fma(v, 2.5, x.y);
        *p = x;
      } else {
        atomicAdd(reg + 32 * off + lane, v);
      }
    }
  }
  __syncthreads();
  double s = 0;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) s += sm[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = khz * 1e3;
  printf("%s: %d SMs, %.0f MHz nominal\n", p.name, sms, clk / 1e6);
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * sms * 4 * 1024));
  const int smem = 16384 * 8 + 4096;
  CK(cudaFuncSetAttribute(k_rmw<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(k_rmw<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(k_rmw<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int threads : {256, 512, 1024}) {
    const int iters = 20000;
    float ms = time_ms([&] { k_dfma<<<sms * 2, threads>>>(out, iters); });
    printf("dfma   threads=%4d x2 CTAs: %.2f FMA/clk/SM\n", threads, 2.0 * threads * 8.0 * iters / (ms * 1e-3 * clk));
    ms = time_ms([&] { k_dmma<<<sms * 2, threads>>>(out, iters / 4); });
    printf("dmma   threads=%4d x2 CTAs: %.2f FMA/clk/SM (256 per m8n8k4)\n", threads,
           2.0 * (threads / 32) * 4.0 * 256.0 * (iters / 4) / (ms * 1e-3 * clk));
    ms = time_ms([&] { k_rmw<0><<<sms, threads, smem>>>(out, iters / 4); });
    printf("rmw64  threads=%4d: %.2f FMA/clk/SM\n", threads, 1.0 * threads * 8.0 * (iters / 4) / (ms * 1e-3 * clk));
    ms = time_ms([&] { k_rmw<1><<<sms, threads, smem>>>(out, iters / 4); });
    printf("rmw128 threads=%4d: %.2f FMA/clk/SM\n", threads, 2.0 * threads * 8.0 * (iters / 4) / (ms * 1e-3 * clk));
    ms = time_ms([&] { k_rmw<2><<<sms, threads, smem>>>(out, iters / 16); });
    printf("atom64 threads=%4d: %.2f add/clk/SM\n", threads, 1.0 * threads * 8.0 * (iters / 16) / (ms * 1e-3 * clk));
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
