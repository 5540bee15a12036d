// ubench5.cu -- prototype of the cell-binned Gram accumulation (development tool).
//
// Idea: for a pair of cubic spline terms (i, j) every data row lands in exactly one CELL
// (interval of x_i, interval of x_j) and adds a 4 x 4 outer product to the 16 elements of G that
// belong to the cell.  Instead of scattering into shared memory (16 B of read-modify-write per FMA,
// 7.7 FMA/clk/SM), a thread OWNS a cell: the tile's rows are counting-sorted by cell in shared
// memory and every thread walks the rows of its own cell, accumulating in REGISTERS for the whole
// kernel.  Bound: the fp64 pipe (62.7 FMA/clk/SM) divided by the basis re-evaluation overhead and
// the lane imbalance of the per-tile cell counts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ubench5 ubench5.cu && ./ubench5
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../pygam_b200/csrc/pgb_kernels.cuh"

struct Marg {
  double lo, inv_span, nint_d;
  int nint, feature;
};
struct Role {
  Marg a, b;
  int diag, ncell, ncb, pad;
};

#define NACC 18
#define HSPLIT 22

__device__ __forceinline__ int cell_of(const Marg& m, double x, double& s) {
  const double u = (x - m.lo) * m.inv_span;
  s = u * m.nint_d;
  return min(max((int)s, 0), m.nint - 1);
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    cell_kernel(const double* __restrict__ X, int64_t ld, int64_t n, const double* __restrict__ om,
                const double* __restrict__ omz, const Role* __restrict__ roles, int n_roles, int TR,
                const int* __restrict__ cta_slot0, double* __restrict__ part, int mode) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ Role R;
  __shared__ int s_wtot[32];
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* xa = reinterpret_cast<double*>(smraw);
  double* xb = xa + TR;
  double* w = xb + TR;
  uint16_t* key = reinterpret_cast<uint16_t*>(w + TR);
  uint16_t* perm = key + TR;
  int* cursor = reinterpret_cast<int*>(perm + TR);
  const uint32_t bar = smem_addr(&s_bar);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  const int64_t ntiles = (n + TR - 1) / TR;
  const int64_t U = ntiles * n_roles;
  const int64_t u0 = U * blockIdx.x / gridDim.x, u1 = U * (blockIdx.x + 1) / gridDim.x;
  int cur_role = -1;
  int slot = cta_slot0[blockIdx.x];
  uint32_t phase = 0;
  double acc[NACC];
#pragma unroll
  for (int e = 0; e < NACC; ++e) acc[e] = 0.0;
  int NC = 0;
  for (int64_t u = u0; u < u1; ++u) {
    const int role = (int)(u / ntiles);
    const int64_t tile = u - (int64_t)role * ntiles;
    if (role != cur_role) {
      if (cur_role >= 0) {
        if (tid < NC) {
#pragma unroll
          for (int e = 0; e < NACC; ++e) part[((int64_t)slot * NACC + e) * 512 + tid] = acc[e];
        }
        ++slot;
#pragma unroll
        for (int e = 0; e < NACC; ++e) acc[e] = 0.0;
      }
      __syncthreads();
      if (tid == 0) R = roles[role];
      __syncthreads();
      cur_role = role;
      NC = R.ncell;
    }
    const int diag = R.diag;
    const int64_t row0 = tile * TR;
    const int cnt = (int)min((int64_t)TR, n - row0);
    const double* ca = X + (int64_t)R.a.feature * ld + row0;
    const double* cb = diag ? omz + row0 : X + (int64_t)R.b.feature * ld + row0;
    if (cnt == TR) {
      if (tid == 0) {
        mbar_expect_tx(bar, 3u * TR * 8u);
        bulk_g2s(smem_addr(xa), ca, TR * 8u, bar);
        bulk_g2s(smem_addr(xb), cb, TR * 8u, bar);
        bulk_g2s(smem_addr(w), om + row0, TR * 8u, bar);
      }
    } else {
      for (int r = tid; r < cnt; r += THREADS) {
        xa[r] = ca[r];
        xb[r] = cb[r];
        w[r] = om[row0 + r];
      }
    }
    for (int k = tid; k < NC; k += THREADS) cursor[k] = 0;
    __syncthreads();
    if (cnt == TR) {
      mbar_wait(bar, phase);
      phase ^= 1;
    }
    const int ncb = R.ncb;
    for (int r = tid; r < cnt; r += THREADS) {
      double s;
      const int ci = cell_of(R.a, xa[r], s);
      int k;
      if (diag) k = ci * HSPLIT + (r % HSPLIT);
      else k = ci * ncb + cell_of(R.b, xb[r], s);
      key[r] = (uint16_t)k;
      atomicAdd(&cursor[k], 1);
    }
    __syncthreads();
    // exclusive scan of the cell counts: thread k owns cell k
    const int c = tid < NC ? cursor[tid] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int t = lane < THREADS / 32 ? s_wtot[lane] : 0;
      int ti = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += v;
      }
      s_wtot[lane] = ti - t;
    }
    __syncthreads();
    const int start = incl - c + s_wtot[warp];
    if (tid < NC) cursor[tid] = start;
    __syncthreads();
    for (int r = tid; r < cnt; r += THREADS) {
      const int pos = atomicAdd(&cursor[key[r]], 1);
      perm[pos] = (uint16_t)r;
    }
    __syncthreads();
    if (mode != 1 && tid < NC) {
      if (!diag) {
        const double cid = (double)(tid / ncb), cjd = (double)(tid % ncb);
        const double loa = R.a.lo, isa = R.a.inv_span, na = R.a.nint_d;
        const double lob = R.b.lo, isb = R.b.inv_span, nb = R.b.nint_d;
        for (int s = 0; s < c; ++s) {
          const int r = perm[start + s];
          const double ta = ((xa[r] - loa) * isa) * na - cid;
          const double tb = ((xb[r] - lob) * isb) * nb - cjd;
          const double ww = w[r];
          double a0, a1, a2, a3, b0, b1, b2, b3;
          cubic_bspline(ta, a0, a1, a2, a3);
          cubic_bspline(tb, b0, b1, b2, b3);
          a0 *= ww; a1 *= ww; a2 *= ww; a3 *= ww;
          acc[0] = fma(a0, b0, acc[0]); acc[1] = fma(a0, b1, acc[1]); acc[2] = fma(a0, b2, acc[2]); acc[3] = fma(a0, b3, acc[3]);
          acc[4] = fma(a1, b0, acc[4]); acc[5] = fma(a1, b1, acc[5]); acc[6] = fma(a1, b2, acc[6]); acc[7] = fma(a1, b3, acc[7]);
          acc[8] = fma(a2, b0, acc[8]); acc[9] = fma(a2, b1, acc[9]); acc[10] = fma(a2, b2, acc[10]); acc[11] = fma(a2, b3, acc[11]);
          acc[12] = fma(a3, b0, acc[12]); acc[13] = fma(a3, b1, acc[13]); acc[14] = fma(a3, b2, acc[14]); acc[15] = fma(a3, b3, acc[15]);
        }
      } else {
        const double cid = (double)(tid / HSPLIT);
        const double loa = R.a.lo, isa = R.a.inv_span, na = R.a.nint_d;
        for (int s = 0; s < c; ++s) {
          const int r = perm[start + s];
          const double ta = ((xa[r] - loa) * isa) * na - cid;
          const double ww = w[r], wz = xb[r];
          double a0, a1, a2, a3;
          cubic_bspline(ta, a0, a1, a2, a3);
          const double v0 = a0 * ww, v1 = a1 * ww, v2 = a2 * ww, v3 = a3 * ww;
          acc[0] = fma(v0, a0, acc[0]); acc[1] = fma(v0, a1, acc[1]); acc[2] = fma(v0, a2, acc[2]); acc[3] = fma(v0, a3, acc[3]);
          acc[4] = fma(v1, a1, acc[4]); acc[5] = fma(v1, a2, acc[5]); acc[6] = fma(v1, a3, acc[6]);
          acc[7] = fma(v2, a2, acc[7]); acc[8] = fma(v2, a3, acc[8]); acc[9] = fma(v3, a3, acc[9]);
          acc[10] += v0; acc[11] += v1; acc[12] += v2; acc[13] += v3;
          acc[14] = fma(a0, wz, acc[14]); acc[15] = fma(a1, wz, acc[15]); acc[16] = fma(a2, wz, acc[16]); acc[17] = fma(a3, wz, acc[17]);
        }
      }
    }
    __syncthreads();
  }
  if (cur_role >= 0 && tid < NC) {
#pragma unroll
    for (int e = 0; e < NACC; ++e) part[((int64_t)slot * NACC + e) * 512 + tid] = acc[e];
  }
}

static void cubic_h(double t, double* b) {
  const double omt = 1.0 - t, t2 = t * t, t3 = t2 * t;
  b[0] = (omt * omt) * (omt * (1.0 / 6.0));
  b[3] = t3 * (1.0 / 6.0);
  b[1] = fma(t3, 0.5, fma(t2, -1.0, 2.0 / 3.0));
  b[2] = fma(t3, -0.5, fma(t2, 0.5, fma(t, 0.5, 1.0 / 6.0)));
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 4000000;
  const int F = 10, NINT = 22;
  const int check = argc > 2 ? atoi(argv[2]) : 0;
  std::vector<double> hX((size_t)F * n), hom(n), homz(n);
  uint64_t st = 88172645463325252ull;
  auto rnd = [&]() {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    return (double)(st >> 11) * (1.0 / 9007199254740992.0);
  };
  for (auto& v : hX) v = rnd();
  for (int64_t i = 0; i < n; ++i) { hom[i] = 0.1 + 0.15 * rnd(); homz[i] = hom[i] * (rnd() - 0.5); }
  std::vector<Role> roles;
  auto marg = [&](int f) { Marg m; m.lo = 0.0; m.inv_span = 1.0; m.nint = NINT; m.nint_d = NINT; m.feature = f; return m; };
  for (int i = 0; i < F; ++i) {
    Role r{}; r.a = marg(i); r.b = marg(i); r.diag = 1; r.ncell = NINT * HSPLIT; r.ncb = HSPLIT; roles.push_back(r);
    for (int j = i + 1; j < F; ++j) { Role c{}; c.a = marg(i); c.b = marg(j); c.diag = 0; c.ncell = NINT * NINT; c.ncb = NINT; roles.push_back(c); }
  }
  const int nr = (int)roles.size();
  double *dX, *dom, *domz, *dpart;
  Role* droles;
  int* dslot0;
  cudaMalloc(&dX, hX.size() * 8); cudaMalloc(&dom, n * 8); cudaMalloc(&domz, n * 8);
  cudaMalloc(&droles, nr * sizeof(Role));
  cudaMemcpy(dX, hX.data(), hX.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dom, hom.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(domz, homz.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(droles, roles.data(), nr * sizeof(Role), cudaMemcpyHostToDevice);

  auto run = [&](auto kern, int threads, int minb, int TR, int mode, const char* name) {
    const int grid = 148 * minb;
    const size_t smem = (size_t)TR * (24 + 4) + 512 * 4 + 64;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t ntiles = (n + TR - 1) / TR, U = ntiles * nr;
    std::vector<int> slot0(grid + 1, 0);
    for (int c = 0; c < grid; ++c) {
      const int64_t u0 = U * c / grid, u1 = U * (c + 1) / grid;
      const int nro = u1 > u0 ? (int)((u1 - 1) / ntiles - u0 / ntiles + 1) : 0;
      slot0[c + 1] = slot0[c] + nro;
    }
    const int nslots = slot0[grid];
    cudaMalloc(&dslot0, (grid + 1) * 4);
    cudaMemcpy(dslot0, slot0.data(), (grid + 1) * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&dpart, (size_t)nslots * NACC * 512 * 8);
    cudaMemset(dpart, 0, (size_t)nslots * NACC * 512 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) kern<<<grid, threads, smem>>>(dX, n, n, dom, domz, droles, nr, TR, dslot0, dpart, mode);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; ++i) kern<<<grid, threads, smem>>>(dX, n, n, dom, domz, droles, nr, TR, dslot0, dpart, mode);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double fma_row = 45 * 16 + 10 * 18;
    printf("%-28s TR=%5d smem=%6zu grid=%d mode=%d: %8.3f ms  %7.1f Mrows/s  %5.2f FMA/clk/SM  (%s)\n", name, TR, smem, grid,
           mode, ms, n / ms / 1e3, n * fma_row / (ms * 1e-3) / (148 * 1.965e9), cudaGetErrorString(err));
    if (check && mode == 0) {
      std::vector<double> hp((size_t)nslots * NACC * 512);
      cudaMemcpy(hp.data(), dpart, hp.size() * 8, cudaMemcpyDeviceToHost);
      // sum the slots of every role
      std::vector<double> got((size_t)nr * NACC * 512, 0.0);
      for (int c = 0; c < grid; ++c) {
        const int64_t u0 = U * c / grid, u1 = U * (c + 1) / grid;
        if (u1 <= u0) continue;
        const int r0 = (int)(u0 / ntiles), r1 = (int)((u1 - 1) / ntiles);
        for (int r = r0; r <= r1; ++r)
          for (int e = 0; e < NACC * 512; ++e) got[(size_t)r * NACC * 512 + e] += hp[(size_t)(slot0[c] + r - r0) * NACC * 512 + e];
      }
      double worst = 0.0;
      for (int r = 0; r < nr; ++r) {
        std::vector<double> ref((size_t)NACC * 512, 0.0), mag((size_t)NACC * 512, 0.0);
        const Role& R = roles[r];
        for (int64_t i = 0; i < n; ++i) {
          const double xa = hX[(size_t)R.a.feature * n + i];
          const double sa = xa * NINT;
          const int ci = std::min(std::max((int)sa, 0), NINT - 1);
          double a[4], b[4];
          cubic_h(sa - ci, a);
          if (!R.diag) {
            const double sb = hX[(size_t)R.b.feature * n + i] * NINT;
            const int cj = std::min(std::max((int)sb, 0), NINT - 1);
            cubic_h(sb - cj, b);
            const int k = ci * NINT + cj;
            for (int p = 0; p < 4; ++p)
              for (int q = 0; q < 4; ++q) ref[(size_t)(4 * p + q) * 512 + k] += hom[i] * a[p] * b[q];
          } else {
            // diagonal roles: compare sums over the hash split
            int e = 0;
            for (int p = 0; p < 4; ++p)
              for (int q = p; q < 4; ++q) ref[(size_t)(e++) * 512 + ci] += hom[i] * a[p] * a[q];
            for (int p = 0; p < 4; ++p) ref[(size_t)(10 + p) * 512 + ci] += hom[i] * a[p];
            for (int p = 0; p < 4; ++p) ref[(size_t)(14 + p) * 512 + ci] += homz[i] * a[p];
          }
        }
        for (int e = 0; e < NACC; ++e)
          for (int k = 0; k < (R.diag ? NINT : NINT * NINT); ++k) {
            double g = 0.0;
            if (R.diag) for (int h = 0; h < HSPLIT; ++h) g += got[((size_t)r * NACC + e) * 512 + k * HSPLIT + h];
            else g = got[((size_t)r * NACC + e) * 512 + k];
            const double rf = ref[(size_t)e * 512 + k];
            const double d = fabs(g - rf) / (fabs(rf) + 1e-300);
            if (rf != 0.0 && d > worst) worst = d;
            if (rf == 0.0 && g != 0.0) worst = 1.0;
          }
      }
      printf("   check vs host: worst relative error %.3e\n", worst);
    }
    cudaFree(dpart);
    cudaFree(dslot0);
  };
  for (int mode = 0; mode < 2; ++mode) {
    run(cell_kernel<512, 2>, 512, 2, 3584, mode, "512 thr x 2 CTA/SM");
    run(cell_kernel<512, 2>, 512, 2, 2048, mode, "512 thr x 2 CTA/SM");
    run(cell_kernel<512, 1>, 512, 1, 7680, mode, "512 thr x 1 CTA/SM");
    run(cell_kernel<512, 2>, 512, 2, 3072, mode, "512 thr x 2 CTA/SM");
  }
  return 0;
}
