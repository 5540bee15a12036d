// pgb_upload.cu -- the rows of a fit from the caller's (pageable) host array into the feature-major device shard.
//
// Reference code replaced (pyGAM 0.12.0): the ingestion of GAM.fit / predict (pygam.py:861-917 -> check_X / check_y,
// utils.py:96-263), which keeps X as an (n, F) row-major host array.  The engine wants column j of X contiguous
// (DESIGN.md section 3), and a fit of 1e7 rows spends more time in a plain pageable cudaMemcpy of X (one staging
// thread inside the driver) than in its PIRLS iterations.  Here T worker threads each take every T-th block of rows:
// a worker transposes its block into a pinned staging buffer ([F][rows of the block]: F sequential write streams),
// sends it with ONE strided copy (cudaMemcpy2DAsync: F runs of `rows` doubles into the columns of the shard) on its own
// stream and goes on with its next block while the copy runs (two buffers per worker).  No device-side transpose, no
// second device copy of X.  The staging buffers are pinned once per process and reused.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include "pgb_internal.h"

namespace {

const int kMaxWorkers = 16;
const size_t kStageBytes = 4u << 20;  // per buffer; two per worker
const size_t kMaxPitch = 2147483647;   // cudaDeviceProp::memPitch on every current device

struct Worker {
  int device = -1;
  double* buf[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  cudaStream_t stream = nullptr;
};
Worker g_workers[kMaxWorkers];
std::mutex g_upload_mutex;  // one upload at a time per process (the staging buffers are shared)

// the staging buffers of the first kSlabWorkers workers are one pinned allocation made by the first call, whatever
// its size: a small first fit (few blocks, few workers) must not leave the pinning of the other workers' buffers
// to the first large one
const int kSlabWorkers = 8;
double* g_slab = nullptr;

cudaError_t slab_prepare() {
  if (g_slab) return cudaSuccess;
  const size_t bytes = (size_t)kSlabWorkers * 2 * kStageBytes;
  if (cudaHostAlloc((void**)&g_slab, bytes, cudaHostAllocPortable) != cudaSuccess) {
    // no pinned memory to be had (a locked-memory limit): pageable staging buffers still work -- the copies are then
    // staged once more inside the driver and complete before cudaMemcpy2DAsync returns -- only slower
    cudaGetLastError();
    g_slab = static_cast<double*>(aligned_alloc(4096, bytes));
    if (!g_slab) return cudaErrorMemoryAllocation;
  }
  for (int k = 0; k < kSlabWorkers; ++k)
    for (int q = 0; q < 2; ++q)
      if (!g_workers[k].buf[q]) g_workers[k].buf[q] = g_slab + (size_t)(2 * k + q) * (kStageBytes / sizeof(double));
  return cudaSuccess;
}

cudaError_t worker_prepare(Worker& w, int device) {
  if (w.device == device) return cudaSuccess;
  cudaError_t e;
  if (w.device >= 0) {  // another device than last time: the stream and events belong to the old one
    cudaSetDevice(w.device);
    cudaStreamDestroy(w.stream);
    for (int q = 0; q < 2; ++q) cudaEventDestroy(w.done[q]);
    w.device = -1;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return e;
  }
  for (int q = 0; q < 2; ++q) {
    if (!w.buf[q] && cudaHostAlloc((void**)&w.buf[q], kStageBytes, cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();  // pageable staging, as in slab_prepare
      w.buf[q] = static_cast<double*>(aligned_alloc(4096, kStageBytes));
      if (!w.buf[q]) return cudaErrorMemoryAllocation;
    }
    if ((e = cudaEventCreateWithFlags(&w.done[q], cudaEventDisableTiming)) != cudaSuccess) return e;
  }
  if ((e = cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
  w.device = device;
  return cudaSuccess;
}

// rows [r0, r0 + nr) of the row-major (n, F) array -> dst[j * nr + i]
void transpose_block(const double* __restrict__ src, int64_t F, int64_t nr, double* __restrict__ dst) {
  if (F == 1) {
    memcpy(dst, src, (size_t)nr * sizeof(double));
    return;
  }
  const int64_t RB = 256;  // rows per inner block: F write streams of 2 KB stay in L1/L2 while the rows are read once
  for (int64_t i0 = 0; i0 < nr; i0 += RB) {
    const int64_t i1 = std::min(nr, i0 + RB);
    for (int64_t j = 0; j < F; ++j) {
      const double* s = src + i0 * F + j;
      double* d = dst + j * nr + i0;
      for (int64_t i = 0; i < i1 - i0; ++i) d[i] = s[i * F];
    }
  }
}

}  // namespace

extern "C" int pgb_upload_rows(double* dst, int64_t ld, const double* src, int64_t n, int32_t F, int32_t n_threads) {
  if (!dst || !src || n < 0 || F < 1 || ld < n || n_threads < 0) {
    pgb_set_error("pgb_upload_rows: bad arguments");
    return PGB_EINVAL;
  }
  if (n == 0) return PGB_OK;
  cudaPointerAttributes attr;
  PGB_CUDA(cudaPointerGetAttributes(&attr, dst));
  if (attr.type != cudaMemoryTypeDevice) {
    pgb_set_error("pgb_upload_rows: dst is not device memory");
    return PGB_EINVAL;
  }
  const int device = attr.device;
  // rows per block: a staging buffer holds F x rows doubles; runs of at least 16 rows (or all of them)
  int64_t rows = (int64_t)(kStageBytes / sizeof(double)) / F;
  if (rows < 1) {
    pgb_set_error("pgb_upload_rows: %d features do not fit a staging buffer", (int)F);
    return PGB_EUNSUPPORTED;
  }
  rows = std::min(rows, n);
  const int64_t n_blocks = (n + rows - 1) / rows;
  int T = n_threads > 0 ? n_threads : (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
  T = (int)std::min<int64_t>(std::min(T, kMaxWorkers), n_blocks);
  std::lock_guard<std::mutex> lock(g_upload_mutex);
  {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess && n_blocks > 1) e = slab_prepare();  // (a one-block upload pins worker 0's two buffers only)
    for (int k = 0; k < kSlabWorkers && n_blocks > 1 && e == cudaSuccess; ++k) e = worker_prepare(g_workers[k], device);
    cudaSetDevice(cur);
    if (e != cudaSuccess) return pgb_cuda_fail(e, "pgb_upload_rows (staging buffers)");
  }
  cudaError_t errs[kMaxWorkers];
  auto work = [&](int k) {
    cudaError_t e = cudaSetDevice(device);
    Worker& w = g_workers[k];
    if (e == cudaSuccess) e = worker_prepare(w, device);
    int q = 0;
    bool used[2] = {false, false};
    for (int64_t b = k; b < n_blocks && e == cudaSuccess; b += T, q ^= 1) {
      const int64_t r0 = b * rows, nr = std::min(rows, n - r0);
      if (used[q] && (e = cudaEventSynchronize(w.done[q])) != cudaSuccess) break;
      transpose_block(src + r0 * F, F, nr, w.buf[q]);
      if ((size_t)ld * sizeof(double) <= kMaxPitch) {
        e = cudaMemcpy2DAsync(dst + r0, (size_t)ld * sizeof(double), w.buf[q], (size_t)nr * sizeof(double),
                              (size_t)nr * sizeof(double), (size_t)F, cudaMemcpyHostToDevice, w.stream);
      } else {  // columns further apart than the largest pitch of a strided copy (> 2.6e8 rows): one copy per column
        for (int64_t j = 0; j < F && e == cudaSuccess; ++j)
          e = cudaMemcpyAsync(dst + j * ld + r0, w.buf[q] + j * nr, (size_t)nr * sizeof(double), cudaMemcpyHostToDevice,
                              w.stream);
      }
      if (e == cudaSuccess) e = cudaEventRecord(w.done[q], w.stream);
      used[q] = true;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(w.stream);
    errs[k] = e;
  };
  if (T == 1) {
    int cur = 0;
    cudaGetDevice(&cur);
    work(0);
    cudaSetDevice(cur);
  } else {
    std::vector<std::thread> th;
    th.reserve(T);
    for (int k = 0; k < T; ++k) th.emplace_back(work, k);
    for (auto& t : th) t.join();
  }
  for (int k = 0; k < T; ++k)
    if (errs[k] != cudaSuccess) return pgb_cuda_fail(errs[k], "pgb_upload_rows");
  return PGB_OK;
}
