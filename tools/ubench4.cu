// ubench4.cu -- shared-memory wavefronts per instruction for the access patterns of the Gram pass.
// Run under ncu and divide l1tex__data_pipe_lsu_wavefronts_mem_shared_op_{ld,st}.sum by
// smsp__inst_executed_op_shared_{ld,st}.sum per kernel (development tool).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// W = access width in bytes (4, 8, 16); PAT selects the per-lane offset
template <int W, int PAT, bool STORE>
__global__ void k_pat(double* out, int iters, int halfgap) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 6144; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int half = lane >> 4, a = (lane >> 2) & 3, b = lane & 3;
  int off;
  switch (PAT) {
    case 0: off = a * W + half * halfgap; break;             // row operand: 4 distinct per half-warp
    case 1: off = b * W + half * halfgap; break;             // column operand
    case 2: off = half * halfgap; break;                     // one address per half-warp
    case 3: off = (a * 28 + b) * 8 + half * halfgap; break;  // accumulator 4 x 4, pitch 28
    case 4: off = lane * halfgap; break;                     // lane-strided (record rows)
    default: off = 0; break;                                 // full broadcast
  }
  const uint32_t base = smem_u32(sm) + off;
  unsigned long long acc = 0;
  for (int it = 0; it < iters; ++it) {
    const uint32_t ad = base + (it & 7) * 2048;
    unsigned long long x = it, y = acc;
    if (!STORE) {
      if (W == 4) { uint32_t t; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(t) : "r"(ad)); x = t; }
      if (W == 8) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(x) : "r"(ad));
      if (W == 16) asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(ad));
      acc += x ^ y;
    } else {
      if (W == 8) asm volatile("st.shared.b64 [%0], %1;" ::"r"(ad), "l"(x) : "memory");
      if (W == 16) asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(ad), "l"(x), "l"(y) : "memory");
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (double)acc;
}

int main() {
  double* out;
  cudaMalloc(&out, 8 * 1024);
  const int smem = 6144 * 8 + 16384 + 32768;
#define RUN(W, PAT, ST, GAP)                                                                      \
  cudaFuncSetAttribute(k_pat<W, PAT, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);    \
  k_pat<W, PAT, ST><<<1, 128, smem>>>(out, 1000, GAP);                                           \
  printf("W=%d PAT=%d store=%d gap=%d: %s\n", W, PAT, ST, GAP, cudaGetErrorString(cudaDeviceSynchronize()));
  RUN(16, 0, false, 64)     // 1
  RUN(16, 0, false, 192)    // 2
  RUN(16, 1, false, 64)     // 3
  RUN(16, 5, false, 0)      // 4
  RUN(8, 0, false, 32)      // 5
  RUN(8, 0, false, 96)      // 6
  RUN(8, 1, false, 32)      // 7
  RUN(8, 1, false, 128)     // 8: both halves same banks
  RUN(4, 2, false, 4)       // 9
  RUN(4, 0, false, 16)      // 10
  RUN(8, 2, false, 8)       // 11
  RUN(8, 3, false, 5600)    // 12: accumulator load
  RUN(8, 3, true, 5600)     // 13: accumulator store
  RUN(16, 4, true, 704)     // 14: record store, stride 704
  RUN(16, 4, true, 720)     // 15: record store, stride 720
  RUN(8, 4, true, 8)        // 16: contiguous 64-bit store
  return 0;
}
