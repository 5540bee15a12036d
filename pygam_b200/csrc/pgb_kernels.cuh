// pgb_kernels.cuh -- small device helpers shared by the translation units of libpgb.
#pragma once
#include "pgb_device.cuh"

static const int kSMs = 148;                 // B200
static const size_t kSmemMax = 227 * 1024;   // opt-in shared memory per CTA on sm_100
static const int kPassThreads = 480;  // + one producer warp = 16 warps: 128 registers per thread
static const int kPassThreadsWide = 576;  // + one producer warp = 19 warps: 104 registers per thread (lean instances)

// deterministic block reduction of S per-thread sums -> dst[0..S)
template <int S>
__device__ __forceinline__ void block_reduce_store(double (&acc)[S], double* red /* S*32 smem */,
                                                   double* dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    double v = acc[s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[s * 32 + warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < S) {
    double v = 0.0;
    for (int w = 0; w < nw; ++w) v += red[threadIdx.x * 32 + w];
    dst[threadIdx.x] = v;
  }
}

// sum `count` partial vectors of length S in a fixed order: out[s] (+)= sum_c part[c*S+s].  One warp per
// scalar (launch with 32 * S threads): lane l sums the partials l, l + 32, ..., then a shuffle tree.
static __global__ void scalar_reduce_kernel(const double* __restrict__ part, int count, int S,
                                            double* __restrict__ out, int accumulate) {
  const int s = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (s >= S) return;
  double v = 0.0;
  for (int c = lane; c < count; c += 32) v += part[(int64_t)c * S + s];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) out[s] = accumulate ? out[s] + v : v;
}

struct EmitLp {
  const double* beta;
  double lp;
  __device__ __forceinline__ void operator()(int c, double v) { lp = fma(v, beta[c], lp); }
};

__device__ __forceinline__ void load_terms_smem(DTerm* s_terms, const DTerm* g_terms, int n_terms) {
  const int words = n_terms * (int)(sizeof(DTerm) / 4);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(g_terms);
  uint32_t* dst = reinterpret_cast<uint32_t*>(s_terms);
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
}

// ----------------------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: the O(k)-per-row
// passes stage tiles of the feature columns in shared memory while the previous tile is being
// evaluated, so that HBM latency never sits in the dependent chain of the basis evaluation
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(mbar), "r"(parity)
                 : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
