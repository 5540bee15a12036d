// ubench2.cu -- shared-memory pipe calibration for the Gram pass (development tool):
//  (1) wavefronts per LDS.64 / LDS.128 for operand-style access patterns (few distinct addresses)
//  (2) read-modify-write throughput vs warps per SM and independent chains per warp (inline PTX,
//      so the compiler cannot serialise the chains)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench2 ubench2.cu && ./ubench2
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// PAT: 0 broadcast, 1: 8 distinct (lane>>2) contiguous, 2: 32 distinct contiguous,
//      3: (a = (lane>>2)&3, half = lane>>4): 4 distinct per half-warp, halves 1 KB apart (+64 B)
template <int WIDTH, int PAT>
__global__ void k_lds(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int e;
  if (PAT == 0) e = 0;
  else if (PAT == 1) e = lane >> 2;
  else if (PAT == 2) e = lane;
  else e = ((lane >> 2) & 3) + (lane >> 4) * (1024 / WIDTH + 64 / WIDTH);
  uint32_t a = smem_u32(sm) + warp * 1536 + e * WIDTH;
  unsigned long long s0 = 0, s1 = 0;
  for (int it = 0; it < iters; ++it) {
    const uint32_t ai = a + (it & 1) * 16384;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (WIDTH == 8) {
        unsigned long long x;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(x) : "r"(ai + u * 2048));
        s0 ^= x;
      } else {
        unsigned long long x, y;
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(ai + u * 2048));
        s0 ^= x;
        s1 ^= y;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (double)(s0 ^ s1);
}

// RMW with ILP independent chains per warp, conflict-free (lane-contiguous) addresses that
// change every iteration
template <int ILP>
__global__ void k_rmw(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 24576; i += blockDim.x) sm[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = smem_u32(sm) + (warp * 768 + lane) * 8;  // 24 rows of 32 doubles per warp
  const double v = 1.0 + lane;
  uint32_t r = warp;
  for (int it = 0; it < iters; ++it) {
    double x[ILP];
    uint32_t ad[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      r = r * 1664525u + 1013904223u;
      ad[u] = base + (((r >> 24) % 6u) * 4 + u % 4) * 256;  // distinct rows for the chains of one step
    }
#pragma unroll
    for (int u = 0; u < ILP; ++u) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x[u]) : "r"(ad[u]));
#pragma unroll
    for (int u = 0; u < ILP; ++u) x[u] = fma(v, 1.5, x[u]);
#pragma unroll
    for (int u = 0; u < ILP; ++u) asm volatile("st.shared.f64 [%0], %1;" ::"r"(ad[u]), "d"(x[u]) : "memory");
  }
  __syncthreads();
  double s = 0;
  for (int i = threadIdx.x; i < 24576; i += blockDim.x) s += sm[i];
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

template <int WIDTH, int PAT>
static void run_lds(double* out, int sms, double clk) {
  const int smem = 8192 * 8, threads = 512, iters = 4000;
  cudaFuncSetAttribute(k_lds<WIDTH, PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float ms = time_ms([&] { k_lds<WIDTH, PAT><<<sms, threads, smem>>>(out, iters); });
  printf("lds%-3d pattern %d: %.2f clk per warp instruction per SM\n", WIDTH * 8, PAT,
         ms * 1e-3 * clk / ((threads / 32) * 8.0 * iters));
}

template <int ILP>
static void run_rmw(double* out, int sms, double clk) {
  const int smem = 24576 * 8, iters = 4000;
  cudaFuncSetAttribute(k_rmw<ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int threads : {128, 256, 512, 1024}) {
    float ms = time_ms([&] { k_rmw<ILP><<<sms, threads, smem>>>(out, iters); });
    printf("rmw ilp=%d warps=%2d: %.2f FMA/clk/SM\n", ILP, threads / 32, 1.0 * threads * ILP * iters / (ms * 1e-3 * clk));
  }
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = khz * 1e3;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 1024);
  run_lds<8, 0>(out, sms, clk);
  run_lds<8, 1>(out, sms, clk);
  run_lds<8, 2>(out, sms, clk);
  run_lds<8, 3>(out, sms, clk);
  run_lds<16, 0>(out, sms, clk);
  run_lds<16, 1>(out, sms, clk);
  run_lds<16, 2>(out, sms, clk);
  run_lds<16, 3>(out, sms, clk);
  run_rmw<1>(out, sms, clk);
  run_rmw<2>(out, sms, clk);
  run_rmw<4>(out, sms, clk);
  run_rmw<8>(out, sms, clk);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
