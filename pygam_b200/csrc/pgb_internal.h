// pgb_internal.h -- host-side model handle shared by the translation units of libpgb.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pgb.h"
#include "pgb_device.cuh"

// A role is a contiguous range of coefficient rows [rlo, rhi) of the Gram matrix whose upper
// strip G[rlo:rhi, rlo:m] (+ one extra column holding the right-hand side r) one CTA keeps in
// shared memory.  Roles partition [0, m); a model whose whole upper triangle fits has one.
struct DRole {
  int32_t rlo, rhi;
  int32_t first_term;  // first term with a coefficient >= rlo
  int32_t slot0;       // first row slot of that term
  int32_t kq;          // slots kept per data row: (k - slot0) + 1 virtual slot for r
  int32_t row_slots;   // leading slots that can hold a column < rhi
  int32_t ldp;         // padded strip row length (doubles), odd
  int32_t pad;
  int64_t strip_off;   // offset of the strip inside one partial (doubles)
  int64_t strip_size;  // (rhi - rlo) * ldp
};

struct pgb_model {
  std::vector<pgb_term> terms;
  std::vector<DTerm> dterms;
  std::vector<DRole> roles;
  DTerm* d_terms = nullptr;
  DRole* d_roles = nullptr;
  Family fam{};
  int32_t n_terms = 0, n_features = 0, m = 0, k = 0, device = 0;
  int32_t tile_rows = 0, gram_threads = 256, gram_ctas_per_role = 0, occupancy = 1;
  size_t gram_smem = 0;
  int64_t partial_stride = 0;  // doubles per chunk partial (sum of strip sizes)
  int32_t pass_ctas = 0;       // CTAs of the O(k) passes
  // staging of the host-buffer entry point (allocated on first use, reused afterwards)
  struct HostStage {
    int64_t cap_rows = 0, ws_bytes = 0;
    bool has_w = false;
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    double* dX[2] = {nullptr, nullptr};
    double* dy[2] = {nullptr, nullptr};
    float* dw[2] = {nullptr, nullptr};
    void* ws[2] = {nullptr, nullptr};
    double* d_out = nullptr;
    double* d_beta = nullptr;
  } stage;
  struct Timing {
    bool enabled = false, pending = false;
    cudaEvent_t ev[2] = {nullptr, nullptr};
  } timing;
};

void pgb_count_launch(int n);

void pgb_set_error(const char* fmt, ...);
int pgb_cuda_fail(cudaError_t e, const char* what);
#define PGB_CUDA(call)                                           \
  do {                                                           \
    cudaError_t e_ = (call);                                     \
    if (e_ != cudaSuccess) return pgb_cuda_fail(e_, #call);      \
  } while (0)
