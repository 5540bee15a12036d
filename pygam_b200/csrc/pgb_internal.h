// pgb_internal.h -- host-side model handle shared by the translation units of libpgb.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pgb.h"
#include "pgb_device.cuh"

// ---- plan of the Gram pass (DESIGN.md "Gram pass") ---------------------------------------
// The non-zeros of one model-matrix row occupy "slots" (fixed per model: every term emits nnz
// slots per row) plus one virtual slot for the right-hand side.  Slots are grouped four by four
// into QUADS; the coefficients are partitioned into GROUPS (one per multi-slot term; single-slot
// terms -- intercept, l(), f() -- and the rhs share "misc" groups of one quad).  A BLOCK is the
// accumulator G[group gi, group gj], gi <= gj, rows x pitch doubles with pitch = 4 (mod 8), so
// that the 4 x 4 outer product of two quads (4 consecutive rows x 4 consecutive columns) lands
// in 16 distinct 8-byte banks.  An ITEM is (row quad, column quad) of a block: 16 products per
// data row, computed by the 16 lanes (a, b) of a half-warp.  A STEP is two items, one per
// half-warp; a warp owns a list of steps for the whole kernel, and every accumulator element
// is updated by exactly one warp -- the fp64 read-modify-write needs no atomics.  Blocks are
// packed into ROLES (what one CTA keeps in shared memory); every role walks the rows assigned
// to its CTAs.
struct DGroup {
  int32_t width;    // coefficients of the group (+1 for the rhs column when it lives here)
  int32_t pitch;    // row pitch (doubles) of a block whose columns are this group
  int32_t n_quads;
  int32_t quad0;    // first global quad
};

struct DBlock {
  int32_t gi, gj;   // row group <= column group
  int32_t role;
  int32_t n_items;
  int32_t tr, pad;  // tr = 1: stored transposed, [coefficients of gj][pitch of gi]
  int64_t off;      // offset inside the role's accumulator (doubles)
};

// one (shared quad, varying quad) pair of a half-step: 16 products per data row
struct DVar {
  uint32_t slot;    // role-local record slot of the varying quad's first entry
  uint32_t off8;    // block offset inside the role's accumulator, bytes
  uint32_t pitch8;  // block row pitch, bytes
  uint32_t mask;    // bit 4 * (varying slot) + (shared slot): the lane accumulates; bit 16: same group
                    // on both sides, element = (min index, max index)
};

// What one half-warp does per data row in one step: the operand of the SHARED quad is loaded once
// (lane & 3 selects its slot) and multiplied with up to PGB_MAX_K VARYING quads ((lane >> 2) & 3
// selects the slot), each product going to its own block at
// [index of the varying slot][index of the shared slot] (blocks whose shared quad is the row side
// are stored transposed).
#define PGB_MAX_K 4
struct DHalf {
  uint32_t shared_slot, pad;
  DVar v[PGB_MAX_K];
};

struct DStep {
  DHalf h[2];
};

struct DWarp {
  int32_t step0, n_steps;
};

struct DRoleTerm {
  int32_t term;      // term index, or -1 for the virtual rhs slot
  int32_t slot0;     // role-local record slot of its first non-zero
  int32_t idx_base;  // index of the term's first coefficient inside its group
  int32_t pad;
};

// a contiguous range of global record slots a role needs: one TMA bulk copy per tile
struct DRun {
  int32_t gslot0, lslot0, n_slots, pad;
};

struct DRole {
  int32_t warp0, n_warps;      // range in the warp table
  int32_t n_slots;             // record slots per data row (4 per quad)
  int32_t needs_all_terms;     // role evaluates every term (inline IRLS possible)
  int32_t cta0, n_ctas;        // CTAs of the 1-d grid that work for this role
  int32_t rhs_slot, max_steps; // record slot of the virtual rhs entry (-1: none); longest warp
  int32_t step0, n_steps;      // range in the step table (copied to shared memory by every CTA)
  int32_t run0, n_runs;        // range in the run table
  int64_t acc_size;            // doubles of the role's accumulator
  int64_t part_off;            // offset of the role's first partial inside the workspace (doubles)
  int64_t map_off;             // offset of the role's element map
};

// ---- plan of the cell-owned Gram pass (pgb_cells.cu) -----------------------------------
// A PAIR is a block G[spline term i, spline term j], i <= j; its CELLS are (interval of x_i, interval
// of x_j) -- for a diagonal pair (interval of x_i, row index mod ncb).  A ROLE = up to 512 consecutive
// cells of one pair: what the threads of one CTA own while they serve the role.
struct DCellRole {
  DMarg a, b;                       // marginals of the row term and of the column term (b = a on the diagonal)
  int32_t diag, ncell, ncb, key0;   // cells key0 .. key0 + ncell - 1 of the pair; cell = ci * ncb + cj
  int32_t pair, c_lo, c_hi;         // CTAs c_lo .. c_hi can serve this role
  uint32_t hmul;                    // diagonal role: row r of a tile goes to split (r * hmul) >> 16
  int32_t sa, sb, pad[2];           // spline index (column of the cell records) of the two terms
};
struct DCellPair {
  int32_t ti, tj;                   // terms; tj = -1: diagonal pair of ti
  int32_t off_i, off_j;             // first coefficients
  int32_t nint_i, ncb, ncells, pad;
  int64_t sum_off;                  // sums[sum_off + product * ncells + cell]
};

// ---- generic cell pairs (pgb_cells.cu): every block of [G | r] with a te(), f() or l() term on one side ----
// A term is a product of FACTORS, each with one column of the cell records: a cubic marginal (s(), a marginal
// of te(): knot interval byte + centred local coordinate), a level (f(): coefficient index byte, 0xff = the row
// has no entry in this term) or a value (l(): the feature itself).  The block of two terms is a sum over the
// rows of omega * prod(factor power vectors), so inside one CELL (fixed interval / level of every factor) it
// is a set of MOMENT sums: digit d of the product index = exponent of the d-th cubic factor ("axis").  A
// thread owns (cell, slice): the slice fixes the exponents of the OUTER cubic factors, the two INNER ones span
// the thread's 4 x 4 register sums.  Rows of a tile are ordered by cell once per (key signature, tile) --
// the ORDER -- and shared by every pair with the same key (te diagonal block, te x intercept, te x rhs).
enum { PGB_FK_CUBIC = 0, PGB_FK_LEVEL = 1, PGB_FK_VALUE = 2 };
struct DGenOrder {
  int32_t nkey, nh, ncells, spitch;   // key dimensions, hash split of the rows, cells = prod(key_n) * nh
  uint32_t hmul;
  int32_t key_col[4], key_n[4], key_mul[4], key_kind[4];
  int32_t pad;
  int64_t starts_off;                 // starts of tile t at starts_g[starts_off * n_tiles + t * spitch]
};
struct DGenPair {
  int32_t ta, tb;                     // terms, ta <= tb; tb == n_terms: the right-hand side column
  int32_t off_a, off_b;               // first coefficients
  int32_t nf;                         // factor occurrences: those of ta, then those of tb
  int32_t f_kind[4], f_side[4], f_col[4];
  int32_t f_n[4];                     // cubic: knot intervals; level: coefficients
  int32_t f_nspl[4], f_stride[4];     // coefficient index of the factor = (local index / stride) % nspl
  int32_t f_key[4], f_ax[4];          // key dimension / moment axis of the factor (-1: none)
  int32_t n_ax, n_acc, nsl, nprod;    // axes; sums of the inner factors (1, 4, 16); outer exponent combinations (1, 4, 16); nprod = n_acc * nsl
  int32_t ne, nslt;                   // exponents of the first outer factor a thread covers (1 or 4: n_acc * ne sums per thread); thread slices = nsl / ne
  int32_t in_a, in_b, out0, out1;     // factor occurrences of the inner / outer cubic factors (-1: none)
  int32_t nval, val_col[2];           // value factors multiplied into the weight
  int32_t wz;                         // weight column: omega z (rhs) instead of omega
  int32_t order, nh, ncells, spitch;  // copy of the order's geometry
  int32_t key_mul[4];
  int64_t sum_off, starts_off;
};
struct DGenRole {
  int32_t pair, key0, ncl;            // cells key0 .. key0 + ncl - 1 of the pair; thread = (cell - key0) * nslt + slice
  int32_t c_lo, c_hi, pad;
};

struct pgb_model {
  struct Cells {
    bool enabled = false;
    bool mixed = false;                // the model has te() / f() / l() terms: generic pairs next to the spline pairs
    std::vector<DCellRole> roles;
    std::vector<DCellPair> pairs;
    std::vector<int32_t> coef_term, pair_of;
    std::vector<DGenOrder> gorders;
    std::vector<DGenPair> gpairs;
    std::vector<DGenRole> groles;
    std::vector<int32_t> term_class, gpair_of;  // 0 intercept, 1 plain cubic s(), 2 generic; (ta, tb) -> generic pair
    std::vector<int32_t> grole_prefix, gcta_role_lo;  // weighted units of the generic kernel (WUnitIter)
    DCellRole* d_roles = nullptr;
    DCellPair* d_pairs = nullptr;
    int32_t* d_coef_term = nullptr;
    int32_t* d_pair_of = nullptr;
    DGenOrder* d_gorders = nullptr;
    DGenPair* d_gpairs = nullptr;
    DGenRole* d_groles = nullptr;
    int32_t* d_term_class = nullptr;
    int32_t* d_gpair_of = nullptr;
    int32_t* d_grole_prefix = nullptr;
    int32_t* d_gcta_role_lo = nullptr;
    int32_t tile_rows = 0, grid = 0, slots_per_cta = 0, icpt_coef = -1, n_spl = 0, occ = 1, threads = 512;
    int32_t ring_groups = 1;               // independent consumer groups per CTA of the ring kernel
    int32_t ncol = 0;                  // record columns (one per factor)
    int32_t gslots_per_cta = 0, ggrid = 0;  // partial slots per CTA and grid of the generic kernel
    int32_t gtile_rows = 0;                 // rows per tile of the generic orders (16-bit row indices: < 65536)
    size_t smem = 0;
    int64_t sum_doubles = 0, slot_doubles = 0;
    int64_t gsum_doubles = 0, gstarts_u16 = 0;  // generic sums; starts of all orders per tile
    int64_t gslot_doubles = 0;                   // partial slot of a generic role: kGenAcc sums x kGenSumThreads threads
  } cells;
  std::vector<pgb_term> terms;
  std::vector<DTerm> dterms;
  std::vector<DGroup> groups;
  std::vector<DBlock> blocks;
  std::vector<DStep> steps;
  std::vector<DWarp> warps;
  std::vector<DRoleTerm> rterms;  // every term and the rhs with its global record slot
  std::vector<DRun> runs;
  std::vector<int32_t> pad_slots;  // global record slots no term fills
  std::vector<DRole> roles;
  std::vector<int32_t> elem_map;  // per accumulator element: row * (m + 1) + col of G|r, or -1
  DTerm* d_terms = nullptr;
  DStep* d_steps = nullptr;
  DWarp* d_warps = nullptr;
  DRoleTerm* d_rterms = nullptr;
  DRun* d_runs = nullptr;
  int32_t* d_pad_slots = nullptr;
  DRole* d_roles = nullptr;
  int32_t* d_elem_map = nullptr;
  Family fam{};
  int32_t n_terms = 0, n_features = 0, m = 0, k = 0, device = 0;
  int32_t tile_rows = 0, gram_threads = 256, gram_ctas = 0;
  size_t gram_smem = 0, rec_smem = 0;
  int32_t n_slots_all = 0, rec_ctas = 0;
  int64_t partial_doubles = 0;  // doubles of all partials of all roles
  int64_t map_size = 0;
  int32_t pass_ctas = 0;        // CTAs of the O(k) passes
  bool general_ok = true;       // the general Gram plan exists (it needs every term-pair block of G in one SM's shared memory)
  // set only through the test-only entry point pgb_debug_model_create (tests compare the two Gram paths and
  // exercise tile / band boundaries at small n); zero in every model the product creates
  struct Debug {
    int32_t force_general = 0, cells_tile_rows = 0, cells_band_tiles = 0, cells_threads = 0, cells_kernel = 0;
  } dbg;
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
    int n_chunks = 0;
    std::vector<cudaEvent_t> chunk_ev;  // [0, 1]: whole pass; then per row chunk: before records, between, after accumulation
  } timing;
};

// pgb_core.cu: the IRLS step of a Gram pass as an instance of the O(k) pass kernel (stats_pass_kernel)
int pgb_pass_irls(const pgb_model* M, const double* X, int64_t n, int64_t ld, const double* y, const float* w,
                  const double* beta, double* om, double* omz, double* scal_part, int* grid_out, cudaStream_t st);
void pgb_count_launch(int n);
void pgb_set_error(const char* fmt, ...);
int pgb_cuda_fail(cudaError_t e, const char* what);
#define PGB_CUDA(call)                                           \
  do {                                                           \
    cudaError_t e_ = (call);                                     \
    if (e_ != cudaSuccess) return pgb_cuda_fail(e_, #call);      \
  } while (0)
