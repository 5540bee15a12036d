// pgb_internal.h -- host-side model handle shared by the translation units of libpgb.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pgb.h"
#include "pgb_device.cuh"

// ---- plan of the Gram pass (DESIGN.md "Gram pass") ---------------------------------------
// The non-zeros of one model-matrix row occupy k "slots" (fixed per model: every term emits
// nnz slots per row) plus one virtual slot for the right-hand side.  Consecutive terms are
// grouped into column CHUNKS of at most 16 slots.  A UNIT is (row term j, column chunk g >=
// chunk(j)): the block G[coefficients of term j, columns of chunk g], K_j x width_g doubles, kept
// in shared memory by ONE warp for the whole kernel -- no two warps ever update the same
// accumulator, so the fp64 read-modify-write needs no atomics.  Units are packed into ROLES
// (what one CTA keeps in shared memory); every role walks the rows assigned to its CTAs.
struct DChunk {
  int32_t first_term, n_terms;
  int32_t n_slots;  // slots of the chunk incl. the virtual rhs slot if it lives here
  int32_t width;    // physical columns of the chunk (layout dependent, >= sum of n_coefs)
  int32_t has_rhs;  // the virtual rhs column is the last physical column of this chunk
  int32_t layout;   // 0 natural, 1 four-way interleave (bank-conflict free for cubic splines)
  int32_t pad0, pad1;
};

struct DUnit {
  int32_t row_term;  // term whose coefficients are the rows of this block
  int32_t chunk;
  int32_t rslot;     // record slot of the row term's first non-zero (role-local)
  int32_t nnz_r;     // non-zeros of the row term per data row
  int32_t cslot;     // record slot of the chunk's first slot (role-local)
  int32_t clen;      // slots of the chunk
  int32_t cmin;      // first chunk slot that belongs to a term >= row_term (earlier: lower triangle)
  int32_t width;     // physical width of the block
  int32_t k_rows;    // n_coefs of the row term
  int32_t pad;
  int64_t off;       // offset of the block inside the role's accumulator (doubles)
};

struct DRoleTerm {
  int32_t term;   // term index, or -1 for the virtual rhs slot
  int32_t slot0;  // role-local record slot of its first non-zero
};

struct DRole {
  int32_t unit0, n_units;      // range in the unit table
  int32_t term0, n_rterms;     // range in the role-term table (terms whose basis the role needs)
  int32_t n_slots;             // record slots per data row
  int32_t needs_all_terms;     // role evaluates every term (inline IRLS possible)
  int32_t cta0, n_ctas;        // CTAs of the 1-d grid that work for this role
  int32_t rhs_pos, rhs_slot;   // physical column / record slot of the virtual rhs entry (-1: none)
  int64_t acc_size;            // doubles of the role's accumulator
  int64_t part_off;            // offset of the role's first partial inside the workspace (doubles)
  int64_t map_off;             // offset of the role's element map
};

struct pgb_model {
  std::vector<pgb_term> terms;
  std::vector<DTerm> dterms;
  std::vector<DChunk> chunks;
  std::vector<DUnit> units;
  std::vector<DRoleTerm> rterms;
  std::vector<DRole> roles;
  std::vector<int32_t> elem_map;  // per accumulator element: row * (m + 1) + col of G|r, or -1
  DTerm* d_terms = nullptr;
  DChunk* d_chunks = nullptr;
  DUnit* d_units = nullptr;
  DRoleTerm* d_rterms = nullptr;
  DRole* d_roles = nullptr;
  int32_t* d_elem_map = nullptr;
  Family fam{};
  int32_t n_terms = 0, n_features = 0, m = 0, k = 0, device = 0;
  int32_t tile_rows = 0, gram_threads = 256, gram_ctas = 0;
  size_t gram_smem = 0;
  int64_t partial_doubles = 0;  // doubles of all partials of all roles
  int64_t map_size = 0;
  int32_t pass_ctas = 0;        // CTAs of the O(k) passes
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
