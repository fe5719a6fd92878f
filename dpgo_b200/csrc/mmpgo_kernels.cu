// Hand-written sm_100a FP64 kernels of the MM-PGO / AMM-PGO iteration.
// See mmpgo_kernels.cuh for the data model and DESIGN.md for the rooflines.
#include <cstdlib>
#include "mmpgo_kernels.cuh"
#include "so3_project.cuh"

namespace mmpgo {

template <int D> struct Dim {
  static constexpr int R = D + 1;
  static constexpr int PB = (D + 1) * D;
  static constexpr int BB = (D + 1) * (D + 1);
  static constexpr int SYM = (D + 1) * (D + 2) / 2;
  static constexpr int TNV = 1 + D + D * D;
};

__device__ __forceinline__ int symidx(int r, int c) {
  return r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r;
}

// Deterministic block reduction of K scalars per thread: fixed shuffle tree
// inside each warp, then warp 0 adds the warp results in warp order.
template <int K, int NT>
__device__ __forceinline__ void block_reduce_store(double (&v)[K], double *dst) {
  constexpr int NW = (NT + 31) / 32;
  __shared__ double red[NW][K];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) red[warp][k] = x;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double x = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) x += red[w][threadIdx.x];
    dst[threadIdx.x] = x;
  }
}

// o = row r (run-time) of the NR x D register array M, as a chain of selects: indexing a register array with a
// run-time row would put it in local memory
template <int D, int NR>
__device__ __forceinline__ void sel_row(const double *M, int r, double *o) {
#pragma unroll
  for (int k = 0; k < D; ++k) o[k] = M[k];
#pragma unroll
  for (int rr = 1; rr < NR; ++rr) {
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = (r == rr) ? M[rr * D + k] : o[k];
  }
}

// P = V - sym(V Y^T) Y for one pose, row r of the d x d block  (SOdProduct.h:64-103)
template <int D>
__device__ __forceinline__ void proj_row(const double *V, const double *Y, int r, double *out) {
  // S[r][c] = 0.5 (V_r . Y_c + Y_r . V_c)
  double S[D], Yr[D];
  sel_row<D, D>(Y, r, Yr);
#pragma unroll
  for (int c = 0; c < D; ++c) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      a += V[r * D + k] * Y[c * D + k];
      b += Yr[k] * V[c * D + k];
    }
    S[c] = 0.5 * (a + b);
  }
#pragma unroll
  for (int k = 0; k < D; ++k) {
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) acc += S[c] * Y[c * D + k];
    out[k] = V[r * D + k] - acc;
  }
}

// L2 prefetch of a contiguous byte range (cp.async.bulk.prefetch.L2: one instruction, no registers, no smem).
// The streaming kernels here are bound by DRAM latency x the loads a resident CTA keeps in flight, not by
// bandwidth: a CTA asks for the data of the CTA that will run one wave later, which then finds it in L2.
__device__ __forceinline__ void l2_prefetch(const void *p, size_t bytes) {
  const size_t a0 = reinterpret_cast<size_t>(p) & ~size_t(15);
  const size_t a1 = (reinterpret_cast<size_t>(p) + bytes + 15) & ~size_t(15);
  for (size_t a = a0; a < a1; a += 16384) {
    const unsigned n = (unsigned)min((size_t)16384, a1 - a);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(n) : "memory");
  }
}

// The tile that a CTA `PF_DIST` blocks later will work on: first own pose and pose count; false when there is none or
// its node is masked out.  Measured on the 1 M-pose grid (k_gpass): any distance from 8 to 200 tiles gives the same
// 13-16 % (the data only has to be on its way before the CTA starts), 592 (a full wave) and more lose it again.
constexpr int PF_DIST = 64;
__device__ __forceinline__ bool tile_ahead(const Tiles &tl, int tile, int &ps, int &pc) {
  const int tf = tile + PF_DIST;
  if (tf >= tl.n_tiles) return false;
  if (tl.active && !tl.active[tl.node[tf]]) return false;
  ps = tl.start[tf]; pc = tl.cnt[tf];
  return true;
}
// rows [ps, ps + pc) of a per-pose array with `stride` doubles per pose (null: nothing)
__device__ __forceinline__ void l2_prefetch_rows(const double *base, int stride, int ps, int pc) {
  if (base) l2_prefetch(base + (size_t)ps * stride, (size_t)pc * stride * sizeof(double));
}

// =============================================================================
// K2 epilogues (shared by the two thread mappings below).  Called by every thread of the CTA;
// `valid` selects the threads that own output row `row` of local pose `pl` (global pose p) with
// the complete row acc = (G x)_row and the pose block xp of the input vector.
// =============================================================================
template <int D, int MODE, int NT>
__device__ __forceinline__ void gpass_epilogue(const bool valid, const int pl, const int row, const int p, const int tile,
                                               const double (&acc)[D], const double (&xp)[(D + 1) * D],
                                               const GPassArgs &a) {
  constexpr int PB = Dim<D>::PB;
  double xr[D];                        // row `row` of the input pose block
  sel_row<D, D + 1>(xp, row, xr);
  double sc[4] = {0.0, 0.0, 0.0, 0.0};
  if (MODE == G_EVAL) {
    if (valid) {
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double gv = a.g ? a.g[(size_t)p * PB + row * D + c] : 0.0;
        sc[0] += xr[c] * (gv + 0.5 * acc[c]);
      }
    }
    block_reduce_store<1, NT>(reinterpret_cast<double(&)[1]>(sc), a.partials + (size_t)tile * NS);
    return;
  }
  if (MODE == G_RHS_T) {
    if (valid && row == 0) {
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double gv = a.g ? a.g[(size_t)p * PB + c] : 0.0;
        a.out[(size_t)(a.out_perm ? a.out_perm[p] : p) * D + c] = gv + acc[c];
      }
    }
    return;
  }
  // modes that need the whole pose block of the result: stage in shared memory
  __shared__ double sV[TILE][PB];
  if (MODE == G_GRAD) {
    double v[D];
    if (valid) {
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double gv = a.g[(size_t)p * PB + row * D + c];
        v[c] = gv + acc[c];
        a.out[(size_t)p * PB + row * D + c] = v[c];
        sV[pl][row * D + c] = v[c];
        sc[0] += xr[c] * (gv + 0.5 * acc[c]);
        sc[2] += xr[c] * acc[c];
        sc[3] += xr[c] * gv;
      }
    }
    __syncthreads();
    if (valid) {
      if (row == 0) {
#pragma unroll
        for (int c = 0; c < D; ++c) sc[1] += v[c] * v[c];
      } else {
        double pr[D];
        proj_row<D>(&sV[pl][D], xp + D, row - 1, pr);
#pragma unroll
        for (int c = 0; c < D; ++c) sc[1] += pr[c] * pr[c];
      }
    }
    block_reduce_store<4, NT>(sc, a.partials + (size_t)tile * NS);
    return;
  }
  if (MODE == G_REDGRAD) {
    // all rows of G x are formed, so the surrogate value s1 = sum x.(g + 1/2 G x) comes for free
    if (valid) {
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double gv = a.g[(size_t)p * PB + row * D + c];
        sc[1] += xr[c] * (gv + 0.5 * acc[c]);
        if (row >= 1) {
          const double v = gv + acc[c];
          a.out[(size_t)p * PB + row * D + c] = v;
          sV[pl][row * D + c] = v;
        }
      }
    }
    __syncthreads();
    if (valid && row >= 1) {
      double pr[D];
      proj_row<D>(&sV[pl][D], xp + D, row - 1, pr);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        a.out2[(size_t)p * PB + row * D + c] = pr[c];
        sc[0] += pr[c] * pr[c];
      }
    }
    block_reduce_store<2, NT>(reinterpret_cast<double(&)[2]>(sc), a.partials + (size_t)tile * NS);
    return;
  }
  if (MODE == G_HV) {
    // xp holds the direction p = [tdot; Rdot]; Y and nab come from xref / nab
    double Y[D * D], NB[D * D];
    if (valid) {
      const double *yp = a.xref + (size_t)p * PB + D;
      const double *np = a.nab + (size_t)p * PB + D;
#pragma unroll
      for (int k = 0; k < D * D; ++k) { Y[k] = yp[k]; NB[k] = np[k]; }
    }
    if (valid && row >= 1) {
      // E_r = acc - (sym(nab Y^T) Rdot)_r     (SymBlockDiagProduct, SOdProduct.h:64-89)
      const int r = row - 1;
      double S[D], Yr[D], NBr[D];
      sel_row<D, D>(Y, r, Yr);
      sel_row<D, D>(NB, r, NBr);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double u = 0.0, w = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          u += NBr[k] * Y[c * D + k];
          w += Yr[k] * NB[c * D + k];
        }
        S[c] = 0.5 * (u + w);
      }
#pragma unroll
      for (int k = 0; k < D; ++k) {
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) t += S[c] * xp[(1 + c) * D + k];
        sV[pl][row * D + k] = acc[k] - t;
      }
    }
    __syncthreads();
    if (valid) {
      if (row >= 1) {
        double pr[D];
        proj_row<D>(&sV[pl][D], Y, row - 1, pr);
#pragma unroll
        for (int c = 0; c < D; ++c) {
          a.out[(size_t)p * PB + row * D + c] = pr[c];
          const double pv = xr[c];
          sc[0] += pv * pr[c];
          sc[1] += pr[c] * pr[c];
          sc[2] += pv * pv;
        }
      } else {
#pragma unroll
        for (int c = 0; c < D; ++c) a.out[(size_t)p * PB + c] = 0.0;
      }
    }
    block_reduce_store<3, NT>(reinterpret_cast<double(&)[3]>(sc), a.partials + (size_t)tile * NS);
    return;
  }
}

// =============================================================================
// K2: block-CSR pass.  (d+1) threads per pose, thread `row` owns one output row.
// =============================================================================
template <int D, int MODE>
__global__ void __launch_bounds__(TILE *(D + 1), (MODE == G_HV ? 3 : 4))
k_gpass(Tiles tl, GPassArgs a) {
  constexpr int R = Dim<D>::R, PB = Dim<D>::PB, BB = Dim<D>::BB, SYM = Dim<D>::SYM;
  constexpr int NT = TILE * R;
  const int tile = blockIdx.x;
  const int node = tl.node[tile];
  if (tl.active && !tl.active[node]) return;
  if (threadIdx.x < 4) {
    int ps, pc;
    if (tile_ahead(tl, tile, ps, pc)) {
      if (threadIdx.x < 2) {
        const int ef0 = a.rowptr[ps], ef1 = a.rowptr[ps + pc];
        if (threadIdx.x == 0) l2_prefetch(a.blk + (size_t)ef0 * BB, (size_t)(ef1 - ef0) * BB * sizeof(double));
        else l2_prefetch(a.col + ef0, (size_t)(ef1 - ef0) * sizeof(int));
      } else if (threadIdx.x == 2) {
        l2_prefetch_rows(a.x, PB, ps, pc);
        l2_prefetch_rows(a.diag, SYM, ps, pc);
      } else {
        l2_prefetch_rows(a.g, PB, ps, pc);
        if (MODE == G_HV) { l2_prefetch_rows(a.xref, PB, ps, pc); l2_prefetch_rows(a.nab, PB, ps, pc); }
      }
    }
  }
  const int pl = threadIdx.x / R, row = threadIdx.x % R;
  const int cnt = tl.cnt[tile];
  const int p = tl.start[tile] + pl;
  const bool valid = pl < cnt;
  // which input rows participate: G_RHS_T uses rotation rows only (G01)
  constexpr int K0 = (MODE == G_RHS_T) ? 1 : 0;
  const bool row_on = (MODE == G_RHS_T) ? (row == 0)
                    : (MODE == G_HV) ? (row >= 1) : true;

  double acc[D];
#pragma unroll
  for (int c = 0; c < D; ++c) acc[c] = 0.0;
  double xp[PB];
  if (valid) {
    const double *xpp = a.x + (size_t)p * PB;
#pragma unroll
    for (int k = 0; k < PB; ++k) xp[k] = xpp[k];
  }
  if (valid && row_on) {
    const int e0 = a.rowptr[p], e1 = a.rowptr[p + 1];
    // two entries per trip: the loads of both (column -> neighbour block is a dependent chain) are
    // in flight together
    // the column index of the next entry is fetched one trip ahead, so that the dependent
    // column -> neighbour-block gather costs one round trip per trip instead of two
    int qn = e0 < e1 ? __ldg(a.col + e0) : 0;
#pragma unroll 2
    for (int e = e0; e < e1; ++e) {
      const int q = qn;
      qn = e + 1 < e1 ? __ldg(a.col + e + 1) : 0;
      const double *b = a.blk + (size_t)e * BB + row * R;
      // 128-bit loads: the neighbour's pose block (16-byte aligned for d = 2 and 3) and, for
      // d = 3, this thread's row of the 4 x 4 block
      const double2 *xq2 = reinterpret_cast<const double2 *>(a.x + (size_t)q * PB);
      double br[R], xq[PB];
      if (R % 2 == 0) {
        const double2 *b2 = reinterpret_cast<const double2 *>(b);
#pragma unroll
        for (int k = 0; k < R / 2; ++k) { const double2 v = __ldg(b2 + k); br[2 * k] = v.x; br[2 * k + 1] = v.y; }
      } else {
#pragma unroll
        for (int k = 0; k < R; ++k) br[k] = __ldg(b + k);
      }
#pragma unroll
      for (int k = 0; k < PB / 2; ++k) { const double2 v = __ldg(xq2 + k); xq[2 * k] = v.x; xq[2 * k + 1] = v.y; }
#pragma unroll
      for (int k = K0; k < R; ++k) {
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = fma(br[k], xq[k * D + c], acc[c]);
      }
    }
    const double *dg = a.diag + (size_t)p * SYM;
#pragma unroll
    for (int k = K0; k < R; ++k) {
      const double dv = dg[symidx(row, k)];
#pragma unroll
      for (int c = 0; c < D; ++c) acc[c] = fma(dv, xp[k * D + c], acc[c]);
    }
  }

  gpass_epilogue<D, MODE, NT>(valid, pl, row, p, tile, acc, xp, a);
}

// G_RHS_T as its own kernel: rhs_t = g_t + G01 Y needs only row 0 of every block and the rotation
// rows of the neighbours.  One thread per pose (in k_gpass three of the four row threads idle in this
// mode), row 0 of the blocks from the compact array blk0; same accumulation order as k_gpass.
template <int D>
__global__ void __launch_bounds__(TILE) k_g01(Tiles tl, GPassArgs a) {
  constexpr int R = Dim<D>::R, PB = Dim<D>::PB, SYM = Dim<D>::SYM, DD = D * D;
  const int tile = blockIdx.x;
  const int node = tl.node[tile];
  if (tl.active && !tl.active[node]) return;
  if ((int)threadIdx.x >= tl.cnt[tile]) return;
  const int p = tl.start[tile] + threadIdx.x;
  double acc[D];
#pragma unroll
  for (int c = 0; c < D; ++c) acc[c] = 0.0;
  const int e0 = a.rowptr[p], e1 = a.rowptr[p + 1];
  int qn = e0 < e1 ? __ldg(a.col + e0) : 0;
#pragma unroll 2
  for (int e = e0; e < e1; ++e) {
    const int q = qn;
    qn = e + 1 < e1 ? __ldg(a.col + e + 1) : 0;
    const double *b = a.blk0 + (size_t)e * R;
    const double *xq = a.x + (size_t)q * PB + D;          // rotation rows of the neighbour
    double br[R], xr[DD];
#pragma unroll
    for (int k = 1; k < R; ++k) br[k] = __ldg(b + k);
#pragma unroll
    for (int k = 0; k < DD; ++k) xr[k] = __ldg(xq + k);
#pragma unroll
    for (int k = 1; k < R; ++k) {
#pragma unroll
      for (int c = 0; c < D; ++c) acc[c] = fma(br[k], xr[(k - 1) * D + c], acc[c]);
    }
  }
  const double *dg = a.diag + (size_t)p * SYM;
  const double *xp = a.x + (size_t)p * PB + D;
#pragma unroll
  for (int k = 1; k < R; ++k) {
    const double dv = dg[symidx(0, k)];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = fma(dv, xp[(k - 1) * D + c], acc[c]);
  }
  const size_t orow = a.out_perm ? (size_t)a.out_perm[p] : (size_t)p;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    const double gv = a.g ? a.g[(size_t)p * PB + c] : 0.0;
    a.out[orow * D + c] = gv + acc[c];
  }
}

// k_g01 for SE(3), 2 lanes per pose: lane h reads columns
// 2h, 2h+1 of the compact row (one 16-byte load) and the neighbour's rows 2h, 2h+1 (h = 0: rotation row 1 only).
__global__ void __launch_bounds__(TILE * 2) k_g01_3(Tiles tl, GPassArgs a) {
  constexpr int D = 3, R = 4, PB = 12, SYM = 10;
  const int tile = blockIdx.x;
  const int node = tl.node[tile];
  if (tl.active && !tl.active[node]) return;
  if (threadIdx.x < 4) {
    int ps, pc;
    if (tile_ahead(tl, tile, ps, pc)) {
      if (threadIdx.x < 2) {
        const int ef0 = a.rowptr[ps], ef1 = a.rowptr[ps + pc];
        if (threadIdx.x == 0) l2_prefetch(a.blk0 + (size_t)ef0 * R, (size_t)(ef1 - ef0) * R * sizeof(double));
        else l2_prefetch(a.col + ef0, (size_t)(ef1 - ef0) * sizeof(int));
      } else if (threadIdx.x == 2) {
        l2_prefetch_rows(a.x, PB, ps, pc);
      } else {
        l2_prefetch_rows(a.diag, SYM, ps, pc);
        l2_prefetch_rows(a.g, PB, ps, pc);
      }
    }
  }
  const int pl = threadIdx.x >> 1, h = threadIdx.x & 1;
  const bool valid = pl < tl.cnt[tile];
  const int p = tl.start[tile] + pl;
  double acc[D] = {0.0, 0.0, 0.0};
  if (valid) {
    const int e0 = a.rowptr[p], e1 = a.rowptr[p + 1];
    int qn = e0 < e1 ? __ldg(a.col + e0) : 0;
#pragma unroll 2
    for (int e = e0; e < e1; ++e) {
      const int q = qn;
      qn = e + 1 < e1 ? __ldg(a.col + e + 1) : 0;
      const double2 b2 = __ldg(reinterpret_cast<const double2 *>(a.blk0 + (size_t)e * R) + h);
      const double2 *xq2 = reinterpret_cast<const double2 *>(a.x + (size_t)q * PB) + 3 * h;
      const double2 v1 = __ldg(xq2 + 1), v2 = __ldg(xq2 + 2);
      if (h) {                               // column 2 (rotation row 2); column 0 is the translation row: not part of G01
        const double2 v0 = __ldg(xq2);
        acc[0] = fma(b2.x, v0.x, acc[0]); acc[1] = fma(b2.x, v0.y, acc[1]); acc[2] = fma(b2.x, v1.x, acc[2]);
      }
      acc[0] = fma(b2.y, v1.y, acc[0]); acc[1] = fma(b2.y, v2.x, acc[1]); acc[2] = fma(b2.y, v2.y, acc[2]);
    }
    const double *dg = a.diag + (size_t)p * SYM;
    const double *xh = a.x + (size_t)p * PB + 6 * h;
    const double d1 = dg[symidx(0, 2 * h + 1)];
    if (h) {
      const double d0 = dg[symidx(0, 2)];
#pragma unroll
      for (int c = 0; c < D; ++c) acc[c] = fma(d0, xh[c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = fma(d1, xh[3 + c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < D; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
  if (valid && h == 0) {
    const size_t orow = a.out_perm ? (size_t)a.out_perm[p] : (size_t)p;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const double gv = a.g ? a.g[(size_t)p * PB + c] : 0.0;
      a.out[orow * D + c] = gv + acc[c];
    }
  }
}

template <int D> void launch_gpass(int mode, const Tiles &tl, const GPassArgs &a, cudaStream_t s) {
  const dim3 grid(tl.n_tiles), block(TILE * (D + 1));
  switch (mode) {
    case G_EVAL: k_gpass<D, G_EVAL><<<grid, block, 0, s>>>(tl, a); break;
    case G_GRAD: k_gpass<D, G_GRAD><<<grid, block, 0, s>>>(tl, a); break;
    case G_RHS_T:
      if (D == 3) k_g01_3<<<grid, TILE * 2, 0, s>>>(tl, a);
      else k_g01<D><<<grid, TILE, 0, s>>>(tl, a);
      break;
    case G_REDGRAD: k_gpass<D, G_REDGRAD><<<grid, block, 0, s>>>(tl, a); break;
    case G_HV: k_gpass<D, G_HV><<<grid, block, 0, s>>>(tl, a); break;
  }
}
template void launch_gpass<2>(int, const Tiles &, const GPassArgs &, cudaStream_t);
template void launch_gpass<3>(int, const Tiles &, const GPassArgs &, cudaStream_t);

// =============================================================================
// K1: inter-node edge pass.  One thread per own pose; its half-edges are
// contiguous 128-byte records.
// =============================================================================
__device__ __forceinline__ double irls_weight(int loss, double e, double delta, double &rho) {
  // DPGOProblem.cpp:647-675; rho is the edge's contribution to fobjE
  if (loss == 0) { rho = 0.5 * e; return 1.0; }
  if (loss == 1) {  // Huber
    const double sq = sqrt(delta);
    const double resc = sqrt(fmax(e, delta));
    rho = 0.5 * fmin(2.0 * sq * resc - delta, e);
    return sq / resc;
  }
  if (loss == 2) {  // Geman-McClure
    const double s = e + delta;
    rho = 0.5 * delta * (e / s);
    return (delta * delta) / (s * s);
  }
  const double w = exp(-e / delta);  // Welsch
  rho = 0.5 * (delta - delta * w);
  return w;
}

template <int D, int MODE>
__global__ void __launch_bounds__(TILE) k_inter(Tiles tl, InterArgs a) {
  constexpr int PB = Dim<D>::PB, SYM = Dim<D>::SYM;
  const int tile = blockIdx.x;
  const int node = tl.node[tile];
  if (tl.active && !tl.active[node]) return;
  if (threadIdx.x < 3) {
    int ps, pc;
    if (tile_ahead(tl, tile, ps, pc)) {
      if (threadIdx.x == 0) {
        const int ef0 = a.rowptr[ps], ef1 = a.rowptr[ps + pc];
        l2_prefetch(a.rec + ef0, (size_t)(ef1 - ef0) * sizeof(InterRec));
      } else if (threadIdx.x == 1) {
        l2_prefetch_rows(a.xa, PB, ps, pc);
      } else {
        l2_prefetch_rows(a.xb, PB, ps, pc);
      }
    }
  }
  const int pl = threadIdx.x;
  const int p = tl.start[tile] + pl;
  const bool valid = pl < tl.cnt[tile];
  const double gam = (a.gamma && a.xb) ? a.gamma[node] : 0.0;
  double sc[7] = {0, 0, 0, 0, 0, 0, 0};
  if (valid) {
    // own pose: z (possibly extrapolated), previous z0, v (= z or z - z0)
    double z[PB], z0[PB], acc[PB];
    const double *pa = a.xa + (size_t)p * PB;
#pragma unroll
    for (int k = 0; k < PB; ++k) { z[k] = pa[k]; z0[k] = 0.0; acc[k] = 0.0; }
    if (a.xb) {
      const double *pb = a.xb + (size_t)p * PB;
#pragma unroll
      for (int k = 0; k < PB; ++k) z0[k] = pb[k];
      if (a.gamma) {
#pragma unroll
        for (int k = 0; k < PB; ++k) z[k] = z[k] + gam * (z[k] - z0[k]);
      }
    }
    if (a.yex) {
#pragma unroll
      for (int k = 0; k < PB; ++k) a.yex[(size_t)p * PB + k] = z[k];
    }
    const int e0 = a.rowptr[p], e1 = a.rowptr[p + 1];
    for (int he = e0; he < e1; ++he) {
      const InterRec *r = a.rec + he;
      const int q = r->other;
      const int own_i = r->own_is_i;
      const double tau = r->tau, kap = r->kappa;
      double t[D], Rm[D * D];
#pragma unroll
      for (int k = 0; k < D; ++k) t[k] = r->t[k];
#pragma unroll
      for (int k = 0; k < D * D; ++k) Rm[k] = r->R[k];
      double zq[PB], zq0[PB];
      const double *qa = a.xa + (size_t)q * PB;
#pragma unroll
      for (int k = 0; k < PB; ++k) { zq[k] = qa[k]; zq0[k] = 0.0; }
      if (a.xb) {
        const double *qb = a.xb + (size_t)q * PB;
#pragma unroll
        for (int k = 0; k < PB; ++k) zq0[k] = qb[k];
        if (a.gamma) {
#pragma unroll
          for (int k = 0; k < PB; ++k) zq[k] = zq[k] + gam * (zq[k] - zq0[k]);
        }
      }
      // role-resolved views: (xi, xj) of the edge i -> j
      const double *xi = own_i ? z : zq, *xj = own_i ? zq : z;
      const double *xi0 = own_i ? z0 : zq0, *xj0 = own_i ? zq0 : z0;
      if (MODE == I_TRIVIAL) {
        // off-diagonal block product (M-form, DPGO_utils.cpp:1964-2028 for S)
        //   own = i: [-tau t_j ; -tau t t_j - kappa R Y_j]
        //   own = j: [-tau (t_i + t^T Y_i) ; -kappa R^T Y_i]
        double c[PB];
        if (own_i) {
#pragma unroll
          for (int k = 0; k < D; ++k) c[k] = -tau * zq[k];
#pragma unroll
          for (int rr = 0; rr < D; ++rr)
#pragma unroll
            for (int k = 0; k < D; ++k) {
              double s = 0.0;
#pragma unroll
              for (int cc = 0; cc < D; ++cc) s += Rm[rr * D + cc] * zq[(1 + cc) * D + k];
              c[(1 + rr) * D + k] = -tau * t[rr] * zq[k] - kap * s;
            }
        } else {
#pragma unroll
          for (int k = 0; k < D; ++k) {
            double s = zq[k];
#pragma unroll
            for (int cc = 0; cc < D; ++cc) s += t[cc] * zq[(1 + cc) * D + k];
            c[k] = -tau * s;
          }
#pragma unroll
          for (int rr = 0; rr < D; ++rr)
#pragma unroll
            for (int k = 0; k < D; ++k) {
              double s = 0.0;
#pragma unroll
              for (int cc = 0; cc < D; ++cc) s += Rm[cc * D + rr] * zq[(1 + cc) * D + k];
              c[(1 + rr) * D + k] = -kap * s;
            }
        }
#pragma unroll
        for (int k = 0; k < PB; ++k) acc[k] += c[k];
        // q(v) = sum_e [1/2 v_o.(M_on v_n) - 1/4 v_n^T M_nn v_n] - 1/4 v_o^T Dinter v_o - xi/2 |v_o|^2
        // with v = z (first update: f0 = -q(z)) or v = z - z0 (Q-term), see DESIGN.md
        double vo[PB], vn[PB];
#pragma unroll
        for (int k = 0; k < PB; ++k) {
          vo[k] = a.use_diff ? z[k] - z0[k] : z[k];
          vn[k] = a.use_diff ? zq[k] - zq0[k] : zq[k];
        }
        double cross = 0.0, nn = 0.0;
        if (a.use_diff) {
          // recompute the block product on v
          if (own_i) {
#pragma unroll
            for (int k = 0; k < D; ++k) cross += vo[k] * (-tau * vn[k]);
#pragma unroll
            for (int rr = 0; rr < D; ++rr)
#pragma unroll
              for (int k = 0; k < D; ++k) {
                double s = 0.0;
#pragma unroll
                for (int cc = 0; cc < D; ++cc) s += Rm[rr * D + cc] * vn[(1 + cc) * D + k];
                cross += vo[(1 + rr) * D + k] * (-tau * t[rr] * vn[k] - kap * s);
              }
          } else {
#pragma unroll
            for (int k = 0; k < D; ++k) {
              double s = vn[k];
#pragma unroll
              for (int cc = 0; cc < D; ++cc) s += t[cc] * vn[(1 + cc) * D + k];
              cross += vo[k] * (-tau * s);
            }
#pragma unroll
            for (int rr = 0; rr < D; ++rr)
#pragma unroll
              for (int k = 0; k < D; ++k) {
                double s = 0.0;
#pragma unroll
                for (int cc = 0; cc < D; ++cc) s += Rm[cc * D + rr] * vn[(1 + cc) * D + k];
                cross += vo[(1 + rr) * D + k] * (-kap * s);
              }
          }
        } else {
#pragma unroll
          for (int k = 0; k < PB; ++k) cross += vo[k] * c[k];
        }
        // v_n^T M_nn v_n : neighbour is j (own = i): tau|t|^2 + kappa|Y|^2;
        //                  neighbour is i (own = j): tau|t + t_e^T Y|^2 + kappa|Y|^2
        double ny = 0.0;
#pragma unroll
        for (int k = D; k < PB; ++k) ny += vn[k] * vn[k];
        double nt = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double s = vn[k];
          if (!own_i) {
#pragma unroll
            for (int cc = 0; cc < D; ++cc) s += t[cc] * vn[(1 + cc) * D + k];
          }
          nt += s * s;
        }
        nn = tau * nt + kap * ny;
        sc[0] += 0.5 * cross - 0.25 * nn;
      } else {  // I_ROBUST: evaluate_E  (B-form residuals, DPGO_utils.cpp:2176-2207)
        double rt[D], rR[D * D];
        double e = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double s = xi[k] - xj[k];
#pragma unroll
          for (int cc = 0; cc < D; ++cc) s += t[cc] * xi[(1 + cc) * D + k];
          rt[k] = s;
          e += tau * s * s;
        }
#pragma unroll
        for (int rr = 0; rr < D; ++rr)
#pragma unroll
          for (int k = 0; k < D; ++k) {
            double s = -xj[(1 + rr) * D + k];
#pragma unroll
            for (int cc = 0; cc < D; ++cc) s += Rm[cc * D + rr] * xi[(1 + cc) * D + k];
            rR[rr * D + k] = s;
            e += kap * s * s;
          }
        double rho;
        const double w = irls_weight(a.loss, e, a.loss_reg, rho);
        sc[0] += rho;
        if (a.split_g && w > a.resc[he]) sc[6] += 1.0;          // a rescale is due (DPGOProblem.cpp:305-306)
        if (a.w_out) a.w_out[he] = w;
        if (a.e_out) a.e_out[he] = e;
        // weighted gradient rows of the own endpoint
        const double wt = w * tau, wk = w * kap;
        if (own_i) {
#pragma unroll
          for (int k = 0; k < D; ++k) acc[k] += wt * rt[k];
#pragma unroll
          for (int rr = 0; rr < D; ++rr)
#pragma unroll
            for (int k = 0; k < D; ++k) {
              double s = 0.0;
#pragma unroll
              for (int cc = 0; cc < D; ++cc) s += Rm[rr * D + cc] * rR[cc * D + k];
              acc[(1 + rr) * D + k] += wt * t[rr] * rt[k] + wk * s;
            }
        } else {
#pragma unroll
          for (int k = 0; k < D; ++k) acc[k] -= wt * rt[k];
#pragma unroll
          for (int k = 0; k < D * D; ++k) acc[D + k] -= wk * rR[k];
        }
        if (a.use_diff) {
          // history terms of evaluate_g_and_f (DPGOProblem.cpp:389-397):
          //   s1 += w0_e <B_e y, B_e z0>     (tr(Y^T DfobjE0), each edge once per node)
          //   s2 += y_i^T Mii y_i + y_j^T Mjj y_j    (1/2 of tr(Y^T Q Y) without xi)
          // here xi/xj are the CURRENT (unextrapolated) points, y = z - z0.
          double yi[PB], yj[PB];
#pragma unroll
          for (int k = 0; k < PB; ++k) { yi[k] = xi[k] - xi0[k]; yj[k] = xj[k] - xj0[k]; }
          double dot = 0.0, qd = 0.0, nyi = 0.0, nyj = 0.0, ntj = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            double sy = yi[k] - yj[k], s0 = xi0[k] - xj0[k], si = yi[k];
#pragma unroll
            for (int cc = 0; cc < D; ++cc) {
              sy += t[cc] * yi[(1 + cc) * D + k];
              s0 += t[cc] * xi0[(1 + cc) * D + k];
              si += t[cc] * yi[(1 + cc) * D + k];
            }
            dot += tau * sy * s0;
            qd += tau * si * si;
            ntj += yj[k] * yj[k];
          }
#pragma unroll
          for (int rr = 0; rr < D; ++rr)
#pragma unroll
            for (int k = 0; k < D; ++k) {
              double sy = -yj[(1 + rr) * D + k], s0 = -xj0[(1 + rr) * D + k];
#pragma unroll
              for (int cc = 0; cc < D; ++cc) {
                sy += Rm[cc * D + rr] * yi[(1 + cc) * D + k];
                s0 += Rm[cc * D + rr] * xi0[(1 + cc) * D + k];
              }
              dot += kap * sy * s0;
            }
#pragma unroll
          for (int k = D; k < PB; ++k) { nyi += yi[k] * yi[k]; nyj += yj[k] * yj[k]; }
          sc[1] += a.w_prev[he] * dot;
          const double se = a.resc ? a.resc[he] : 1.0;         // Rescale::Dynamic: Q carries s_e on both endpoint blocks
          sc[2] += se * (qd + kap * nyi + tau * ntj + kap * nyj);
        }
      }
    }
    // pose-local part: D-type diagonal products and the g row
    const double *dg = a.dinter + (size_t)p * SYM;
    double dz[PB];
#pragma unroll
    for (int rr = 0; rr < D + 1; ++rr)
#pragma unroll
      for (int k = 0; k < D; ++k) {
        double s = 0.0;
#pragma unroll
        for (int cc = 0; cc < D + 1; ++cc) s += dg[symidx(rr, cc)] * z[cc * D + k];
        dz[rr * D + k] = s;
      }
    if (MODE == I_TRIVIAL) {
      // g = S Z = C - (Dinter + xi) x
      double vo[PB];
#pragma unroll
      for (int k = 0; k < PB; ++k) vo[k] = a.use_diff ? z[k] - z0[k] : z[k];
      double vdv = 0.0, vv = 0.0;
#pragma unroll
      for (int rr = 0; rr < D + 1; ++rr)
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double s = 0.0;
#pragma unroll
          for (int cc = 0; cc < D + 1; ++cc) s += dg[symidx(rr, cc)] * vo[cc * D + k];
          vdv += vo[rr * D + k] * s;
          vv += vo[rr * D + k] * vo[rr * D + k];
        }
      sc[0] += -0.25 * vdv - 0.5 * a.xi * vv;
#pragma unroll
      for (int k = 0; k < PB; ++k) {
        const double gk = acc[k] - (dz[k] + a.xi * z[k]);
        a.g[(size_t)p * PB + k] = gk;
        sc[1] += z[k] * acc[k];      // x.C
        sc[2] += z[k] * dz[k];       // x^T Dinter x
        sc[3] += z[k] * z[k];        // |x|^2
      }
    } else {
      // g = DfobjE_own - D x,  D = 2 Dinter + xi   (DPGOProblem.cpp:233-239)
      double ydy = 0.0, yy = 0.0;
      if (a.use_diff) {
#pragma unroll
        for (int k = 0; k < PB; ++k) { const double y = z[k] - z0[k]; yy += y * y; }
      }
#pragma unroll
      for (int k = 0; k < PB; ++k) {
        // Rescale::Dynamic: D is rebuilt from the weights found here before it is applied (k_gfix)
        const double dx = a.split_g ? 0.0 : 2.0 * dz[k] + a.xi * z[k];
        a.g[(size_t)p * PB + k] = acc[k] - dx;
        sc[3] += z[k] * acc[k];      // x.DfobjE
        sc[4] += z[k] * dx;          // x^T D x
      }
      sc[5] += yy;                   // |y_own|^2 for the 2 xi term of Q
      (void)ydy;
    }
  }
  block_reduce_store<7, TILE>(sc, a.partials + (size_t)tile * NS);
}

template <int D> void launch_inter(int mode, const Tiles &tl, const InterArgs &a, cudaStream_t s) {
  if (mode == I_TRIVIAL) k_inter<D, I_TRIVIAL><<<tl.n_tiles, TILE, 0, s>>>(tl, a);
  else k_inter<D, I_ROBUST><<<tl.n_tiles, TILE, 0, s>>>(tl, a);
}
template void launch_inter<2>(int, const Tiles &, const InterArgs &, cudaStream_t);
template void launch_inter<3>(int, const Tiles &, const InterArgs &, cudaStream_t);

// =============================================================================
// Rescale::Dynamic.  update_quadratic_mat (DPGOProblem.cpp:751-840) for the nodes whose rescale vector is replaced:
// s_e = clamp(1.25 omega_e, 0.01, 1) (:309-311), then everything that depends on it, per own pose: the inter-node
// diagonal sum  Dinter = sum_e s_e M_own,e,  the diagonal block of G,  T, N, V' of the proximal step (0.5 xi on the
// auxiliary matrices in this builder, DPGO_utils.cpp:3621, 3635) and the diagonal of G00 (the only part of G00 that
// changes: the PCG translation solve just sees a new diagonal where the reference refactorises).
// =============================================================================
template <int D>
__global__ void __launch_bounds__(TILE) k_rescale(Tiles tl, RescaleArgs a) {
  constexpr int R = Dim<D>::R, SYM = Dim<D>::SYM, TNV = Dim<D>::TNV;
  const int tile = blockIdx.x;
  const int node = tl.node[tile];
  if (tl.active && !tl.active[node]) return;
  if ((int)threadIdx.x >= tl.cnt[tile]) return;
  const int p = tl.start[tile] + threadIdx.x;
  double dx[SYM];
#pragma unroll
  for (int k = 0; k < SYM; ++k) dx[k] = 0.0;
  for (int he = a.rowptr[p]; he < a.rowptr[p + 1]; ++he) {
    const InterRec *r = a.rec + he;
    const double s = fmax(fmin(1.25 * a.w[he], 1.0), 0.01);
    a.resc[he] = s;
    const double tau = r->tau, kap = r->kappa;
    // own = i: [[tau, tau t^T], [tau t, kappa I + tau t t^T]];  own = j: [[tau, 0], [0, kappa I]]
    dx[symidx(0, 0)] += s * tau;
#pragma unroll
    for (int k = 0; k < D; ++k) dx[symidx(1 + k, 1 + k)] += s * kap;
    if (r->own_is_i) {
#pragma unroll
      for (int k = 0; k < D; ++k) {
        dx[symidx(1 + k, 0)] += s * tau * r->t[k];
#pragma unroll
        for (int c = 0; c <= k; ++c) dx[symidx(1 + k, 1 + c)] += s * tau * r->t[k] * r->t[c];
      }
    }
  }
  double Gm[R * R], Hm[R * R];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < R; ++c) {
      const double di = a.dintra[(size_t)p * SYM + symidx(r, c)], dv = dx[symidx(r, c)];
      Gm[r * R + c] = di + 2.0 * dv + (r == c ? a.xi : 0.0);
      Hm[r * R + c] = 2.0 * di + 2.0 * dv + (r == c ? 0.5 * a.xi : 0.0);
    }
#pragma unroll
  for (int k = 0; k < SYM; ++k) a.dinter[(size_t)p * SYM + k] = dx[k];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) a.gdiag[(size_t)p * SYM + symidx(r, c)] = Gm[r * R + c];
  double *c_ = a.tnv + (size_t)p * TNV;
  const double T = 1.0 / Hm[0];
  c_[0] = T;
#pragma unroll
  for (int k = 0; k < D; ++k) c_[1 + k] = T * Hm[1 + k];
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) c_[1 + D + r * D + c] = Hm[(1 + r) * R + 1 + c] - Hm[(1 + r) * R] * (T * Hm[1 + c]);
  a.d00[p] = Gm[0];
  a.ts_rec[a.pose_rec[p]] = Gm[0];
}
template <int D> void launch_rescale(const Tiles &tl, const RescaleArgs &a, cudaStream_t s) {
  k_rescale<D><<<tl.n_tiles, TILE, 0, s>>>(tl, a);
}
template void launch_rescale<2>(const Tiles &, const RescaleArgs &, cudaStream_t);
template void launch_rescale<3>(const Tiles &, const RescaleArgs &, cudaStream_t);

// g = DfobjE_own - D x with D = 2 Dinter + xi (DPGOProblem.cpp:323-327 / :487-488, after the rescale)
template <int D>
__global__ void __launch_bounds__(TILE) k_gfix(Tiles tl, GFixArgs a) {
  constexpr int PB = Dim<D>::PB, SYM = Dim<D>::SYM;
  const int tile = blockIdx.x;
  const int node = tl.node[tile];
  if (tl.active && !tl.active[node]) return;
  const int p = tl.start[tile] + threadIdx.x;
  double sc[5] = {0, 0, 0, 0, 0};
  if ((int)threadIdx.x < tl.cnt[tile]) {
    double z[PB];
#pragma unroll
    for (int k = 0; k < PB; ++k) z[k] = a.x[(size_t)p * PB + k];
    const double *dg = a.dinter + (size_t)p * SYM;
#pragma unroll
    for (int rr = 0; rr < D + 1; ++rr)
#pragma unroll
      for (int k = 0; k < D; ++k) {
        double s = 0.0;
#pragma unroll
        for (int cc = 0; cc < D + 1; ++cc) s += dg[symidx(rr, cc)] * z[cc * D + k];
        const double dxv = 2.0 * s + a.xi * z[rr * D + k];
        a.g[(size_t)p * PB + rr * D + k] -= dxv;
        sc[4] += z[rr * D + k] * dxv;
      }
  }
  block_reduce_store<5, TILE>(sc, a.partials + (size_t)tile * NS);
}
template <int D> void launch_gfix(const Tiles &tl, const GFixArgs &a, cudaStream_t s) {
  k_gfix<D><<<tl.n_tiles, TILE, 0, s>>>(tl, a);
}
template void launch_gfix<2>(const Tiles &, const GFixArgs &, cudaStream_t);
template void launch_gfix<3>(const Tiles &, const GFixArgs &, cudaStream_t);

// Tile <-> registers through shared memory: the CTA reads / writes the tile's pose blocks as one
// contiguous run of doubles (coalesced), every thread then picks up / deposits its own pose.
// Rows are padded to PB + 1 doubles (2-way bank conflicts at most for 64-bit accesses).
template <int PB>
__device__ __forceinline__ void tile_load_n(const double *g, int p0, int cnt, double *sm, double (&v)[PB]) {
  const int n = cnt * PB;
  const double *src = g + (size_t)p0 * PB;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += TILE) sm[(i / PB) * (PB + 1) + i % PB] = src[i];
  __syncthreads();
  if (threadIdx.x < cnt) {
#pragma unroll
    for (int k = 0; k < PB; ++k) v[k] = sm[threadIdx.x * (PB + 1) + k];
  }
}
template <int PB>
__device__ __forceinline__ void tile_store_n(double *g, int p0, int cnt, double *sm, const double (&v)[PB]) {
  const int n = cnt * PB;
  double *dst = g + (size_t)p0 * PB;
  __syncthreads();
  if (threadIdx.x < cnt) {
#pragma unroll
    for (int k = 0; k < PB; ++k) sm[threadIdx.x * (PB + 1) + k] = v[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += TILE) dst[i] = sm[(i / PB) * (PB + 1) + i % PB];
}

template <int D>
__device__ __forceinline__ void tile_load(const double *g, int p0, int cnt, double *sm, double (&v)[(D + 1) * D]) {
  tile_load_n<(D + 1) * D>(g, p0, cnt, sm, v);
}
template <int D>
__device__ __forceinline__ void tile_store(double *g, int p0, int cnt, double *sm, const double (&v)[(D + 1) * D]) {
  tile_store_n<(D + 1) * D>(g, p0, cnt, sm, v);
}

// =============================================================================
// K3: fused Nesterov extrapolation + proximal step + SO(d) polar projection.
// One thread per pose.                (DPGOHash.cpp:255-262, DPGOProblem.cpp:600-632)
// =============================================================================
template <int D>
__global__ void __launch_bounds__(TILE) k_prox(Tiles tl, ProxArgs a) {
  constexpr int PB = Dim<D>::PB, TNV = Dim<D>::TNV;
  const int tile = blockIdx.x;
  const int node = tl.node[tile];
  if (tl.active && !tl.active[node]) return;
  if (threadIdx.x < 8) {
    int ps, pc;
    if (tile_ahead(tl, tile, ps, pc)) {
      const int j = threadIdx.x;
      const double *q = j == 0 ? a.xa : j == 1 ? a.xb : j == 2 ? a.dfa : j == 3 ? a.dfb : j == 4 ? a.ga : j == 5 ? a.gb
                      : j == 6 ? a.xref : a.tnv;
      l2_prefetch_rows(q, j == 7 ? TNV : PB, ps, pc);
    }
  }
  const int p = tl.start[tile] + threadIdx.x;
  const bool valid = threadIdx.x < tl.cnt[tile];
  double sc[1] = {0.0};
  if (valid) {
    const double gam = (a.xb && a.gamma) ? a.gamma[node] : 0.0;
    double y0[PB], df[PB];
#pragma unroll
    for (int k = 0; k < PB; ++k) {
      y0[k] = a.xa[(size_t)p * PB + k];
      df[k] = a.dfa[(size_t)p * PB + k];
    }
    if (a.xb) {
#pragma unroll
      for (int k = 0; k < PB; ++k) y0[k] = y0[k] + gam * (y0[k] - a.xb[(size_t)p * PB + k]);
    }
    if (a.dfb) {
#pragma unroll
      for (int k = 0; k < PB; ++k) df[k] = df[k] + gam * (df[k] - a.dfb[(size_t)p * PB + k]);
    }
    if (a.gex) {
#pragma unroll
      for (int k = 0; k < PB; ++k) {
        double gv = a.ga[(size_t)p * PB + k];
        if (a.gb) gv = gv + gam * (gv - a.gb[(size_t)p * PB + k]);
        a.gex[(size_t)p * PB + k] = gv;
      }
    }
    const double *c = a.tnv + (size_t)p * TNV;
    const double T = c[0];
    const double *N = c + 1, *V = c + 1 + D;
    // M = V' Y0 - Df_Y + N^T Df_t
    double M[D * D], Yn[D * D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int k = 0; k < D; ++k) {
        double s = 0.0;
#pragma unroll
        for (int cc = 0; cc < D; ++cc) s += V[r * D + cc] * y0[(1 + cc) * D + k];
        M[r * D + k] = s - df[(1 + r) * D + k] + N[r] * df[k];
      }
    project_to_SOd<D>(M, Yn);
    // t = t0 - N (Y - Y0) - T Df_t
    double out[PB];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < D; ++r) s += N[r] * (Yn[r * D + k] - y0[(1 + r) * D + k]);
      out[k] = y0[k] - s - T * df[k];
    }
#pragma unroll
    for (int k = 0; k < D * D; ++k) out[D + k] = Yn[k];
#pragma unroll
    for (int k = 0; k < PB; ++k) {
      a.xout[(size_t)p * PB + k] = out[k];
      if (a.xref) {
        const double dlt = out[k] - a.xref[(size_t)p * PB + k];
        sc[0] += dlt * dlt;
      }
    }
  }
  block_reduce_store<1, TILE>(sc, a.partials + (size_t)tile * NS);
}
template <int D> void launch_prox(const Tiles &tl, const ProxArgs &a, cudaStream_t s) {
  k_prox<D><<<tl.n_tiles, TILE, 0, s>>>(tl, a);
}
template void launch_prox<2>(const Tiles &, const ProxArgs &, cudaStream_t);
template void launch_prox<3>(const Tiles &, const ProxArgs &, cudaStream_t);

// =============================================================================
// Per-pose vector operations of the truncated-CG / trust-region bookkeeping.
// =============================================================================
template <int D>
__device__ __forceinline__ void apply_precon(const VecArgs &a, int p, const double *r, const double *Y,
                                             double *v) {
  // DPGOProblem::precondition (DPGOProblem.cpp:579-598): M^{-1} r, then tangent projection
  double u[D * D];
  if (a.precon == 0) {
#pragma unroll
    for (int k = 0; k < D * D; ++k) v[k] = r[k];
    return;
  }
  if (a.precon == 3) {
    // RegularizedCholesky: reg_Chol_precon_.solve(Ydot) was done by the sparse sweeps (DPGOProblem.cpp:592-594)
    const double *z = a.pre + (size_t)p * (D + 1) * D + D;
#pragma unroll
    for (int k = 0; k < D * D; ++k) u[k] = z[k];
  } else if (a.precon == 1) {
#pragma unroll
    for (int rr = 0; rr < D; ++rr)
#pragma unroll
      for (int k = 0; k < D; ++k) u[rr * D + k] = a.pinv[(size_t)p * D * D + rr * D + rr] * r[rr * D + k];
  } else {
    const double *B = a.pinv + (size_t)p * D * D;
#pragma unroll
    for (int rr = 0; rr < D; ++rr)
#pragma unroll
      for (int k = 0; k < D; ++k) {
        double s = 0.0;
#pragma unroll
        for (int cc = 0; cc < D; ++cc) s += B[rr * D + cc] * r[cc * D + k];
        u[rr * D + k] = s;
      }
  }
#pragma unroll
  for (int rr = 0; rr < D; ++rr) proj_row<D>(u, Y, rr, v + rr * D);
}

template <int D, int OP>
__global__ void __launch_bounds__(TILE) k_vec(Tiles tl, VecArgs a) {
  constexpr int PB = Dim<D>::PB, DD = D * D;
  const int tile = blockIdx.x;
  const int node = tl.node[tile];
  if (tl.active && !tl.active[node]) return;
  if (threadIdx.x < 4) {
    int ps, pc;
    if (tile_ahead(tl, tile, ps, pc)) {
      const int j = threadIdx.x;
      l2_prefetch_rows(j == 0 ? a.a : j == 1 ? a.b : j == 2 ? a.c : a.y, PB, ps, pc);
    }
  }
  const int p0 = tl.start[tile], cnt = tl.cnt[tile];
  const int p = p0 + threadIdx.x;
  const bool valid = threadIdx.x < cnt;
  const double *cf = a.coef ? a.coef + (size_t)node * MAXC : nullptr;
  __shared__ double sm[TILE * (PB + 1)];
  double sc[3] = {0, 0, 0};
  // element-wise operations: one thread per double of the tile, fully coalesced
  if (OP == V_CG_DIR || OP == V_CG_FINAL || OP == V_DOTS || OP == V_COPY_ROT || OP == V_COPY || OP == V_DIFFNORM ||
      OP == V_COPY_T) {
    const size_t base = (size_t)p0 * PB;
    const int n = cnt * PB;
    for (int i = threadIdx.x; i < n; i += TILE) {
      const bool rot = i % PB >= D;                 // rotation rows of the pose block
      const size_t o = base + i;
      if (OP == V_CG_DIR) {            // a = v; o1 = p; coef[1] = beta
        if (rot) a.o1[o] = -a.a[o] + cf[1] * a.o1[o];
      } else if (OP == V_CG_FINAL) {   // o1 = s, a = p, b = Hp, o5 = Hs; coef[2] = sigma (sign folded in by the host)
        // s gets all rows: p.t holds tdot = -G00^{-1} G01 p_Y of the last Hessian-vector product, so
        // s.t = sum alpha_k tdot_k is the first-order change of the translations along the step
        a.o1[o] = a.o1[o] + cf[2] * a.a[o];
        if (rot) a.o5[o] = a.o5[o] + cf[2] * a.b[o];
      } else if (OP == V_DOTS) {
        if (rot) { const double x = a.a[o], y = a.b[o]; sc[0] += x * y; sc[1] += x * x; sc[2] += y * y; }
      } else if (OP == V_COPY_ROT) {
        if (rot) a.o1[o] = a.a[o];
      } else if (OP == V_COPY) {
        a.o1[o] = a.a[o];
      } else if (OP == V_COPY_T) {
        if (!rot) a.o1[o] = a.a[o];
      } else {                         // V_DIFFNORM
        const double dlt = a.a[o] - a.b[o];
        sc[0] += dlt * dlt;
      }
    }
    if (OP == V_DOTS || OP == V_DIFFNORM) block_reduce_store<3, TILE>(sc, a.partials + (size_t)tile * NS);
    return;
  }
  // per-pose operations (preconditioner / tangent projection / polar projection need the whole
  // d x d block): tiles go through shared memory so that global accesses stay coalesced
  double Yb[PB], A1[PB], A2[PB], A3[PB], A4[PB];
  if (OP == V_CG_INIT) {
    // a = grad; o1 = s, o2 = r, o3 = v, o4 = p, o5 = Hs
    tile_load<D>(a.a, p0, cnt, sm, A1);
    tile_load<D>(a.y, p0, cnt, sm, Yb);
    if (valid) {
      double v[DD];
      apply_precon<D>(a, p, A1 + D, Yb + D, v);
#pragma unroll
      for (int k = 0; k < D; ++k) { A1[k] = 0.0; A2[k] = 0.0; A3[k] = 0.0; A4[k] = 0.0; }
#pragma unroll
      for (int k = 0; k < DD; ++k) { A2[D + k] = v[k]; A3[D + k] = -v[k]; A4[D + k] = 0.0; sc[0] += A1[D + k] * v[k]; }
    }
    tile_store<D>(a.o2, p0, cnt, sm, A1);     // r = grad (t rows zero)
    tile_store<D>(a.o3, p0, cnt, sm, A2);     // v
    tile_store<D>(a.o4, p0, cnt, sm, A3);     // p = -v
    tile_store<D>(a.o1, p0, cnt, sm, A4);     // s = 0
    tile_store<D>(a.o5, p0, cnt, sm, A4);     // Hs = 0
  } else if (OP == V_CG_STEP) {
    // a = p, b = Hp; o1 = s, o2 = r, o3 = v, o5 = Hs; coef[0] = alpha
    const double al = cf[0];
    tile_load<D>(a.a, p0, cnt, sm, A1);       // p
    tile_load<D>(a.b, p0, cnt, sm, A2);       // Hp
    tile_load<D>(a.o1, p0, cnt, sm, A3);      // s
    if (valid) {
#pragma unroll
      for (int k = 0; k < PB; ++k) A3[k] = A3[k] + al * A1[k];   // all rows, see V_CG_FINAL
    }
    tile_store<D>(a.o1, p0, cnt, sm, A3);
    tile_load<D>(a.o5, p0, cnt, sm, A3);      // Hs
    if (valid) {
#pragma unroll
      for (int k = D; k < PB; ++k) A3[k] = A3[k] + al * A2[k];
    }
    tile_store<D>(a.o5, p0, cnt, sm, A3);
    tile_load<D>(a.o2, p0, cnt, sm, A4);      // r
    tile_load<D>(a.y, p0, cnt, sm, Yb);
    if (valid) {
      double v[DD];
#pragma unroll
      for (int k = D; k < PB; ++k) A4[k] = A4[k] + al * A2[k];
      if (a.precon != 3) {
        apply_precon<D>(a, p, A4 + D, Yb + D, v);
#pragma unroll
        for (int k = 0; k < D; ++k) A1[k] = 0.0;
#pragma unroll
        for (int k = 0; k < DD; ++k) { A1[D + k] = v[k]; sc[0] += A4[D + k] * v[k]; }
      }
    }
    tile_store<D>(a.o2, p0, cnt, sm, A4);
    if (a.precon != 3) tile_store<D>(a.o3, p0, cnt, sm, A1);   // precon 3: v follows in V_CG_PRE, after the sweeps on the new r
  } else if (OP == V_RETRACT) {
    // a = x, b = s; o1 = xprop: rotation rows = proj(x.Y + s.Y), translation rows copied from x
    tile_load<D>(a.a, p0, cnt, sm, A1);
    tile_load<D>(a.b, p0, cnt, sm, A2);
    if (valid) {
      double M[DD], Yn[DD];
#pragma unroll
      for (int k = 0; k < DD; ++k) M[k] = A1[D + k] + A2[D + k];
      project_to_SOd<D>(M, Yn);
#pragma unroll
      for (int k = 0; k < DD; ++k) A1[D + k] = Yn[k];
#pragma unroll
      for (int k = 0; k < D; ++k) A1[k] = A1[k] + A2[k];   // t + s.t: first-order guess for recover_translations
    }
    tile_store<D>(a.o1, p0, cnt, sm, A1);
  } else if (OP == V_CG_PRE) {
    // a = r, o3 = v: v = Proj(Y, pre), s0 = r.v
    tile_load<D>(a.a, p0, cnt, sm, A1);
    tile_load<D>(a.y, p0, cnt, sm, Yb);
    if (valid) {
      double v[DD];
      apply_precon<D>(a, p, A1 + D, Yb + D, v);
#pragma unroll
      for (int k = 0; k < D; ++k) A2[k] = 0.0;
#pragma unroll
      for (int k = 0; k < DD; ++k) { A2[D + k] = v[k]; sc[0] += A1[D + k] * v[k]; }
    }
    tile_store<D>(a.o3, p0, cnt, sm, A2);
  } else if (OP == V_PRECOND) {
    tile_load<D>(a.a, p0, cnt, sm, A1);
    tile_load<D>(a.y, p0, cnt, sm, Yb);
    if (valid) {
      double v[DD];
      apply_precon<D>(a, p, A1 + D, Yb + D, v);
#pragma unroll
      for (int k = 0; k < DD; ++k) { A2[D + k] = v[k]; sc[0] += v[k] * v[k]; }
#pragma unroll
      for (int k = 0; k < D; ++k) A2[k] = 0.0;
    }
    if (a.o1) tile_store<D>(a.o1, p0, cnt, sm, A2);
  }
  if (OP == V_CG_INIT || OP == V_CG_STEP || OP == V_PRECOND || OP == V_CG_PRE)
    block_reduce_store<3, TILE>(sc, a.partials + (size_t)tile * NS);
}

// rhs[perm[p d + r]][c] = src[p][1 + r][c]: the rotation rows of a pose-block vector in the elimination order of the
// G11 factor (RegularizedCholesky preconditioner)
template <int D>
__global__ void __launch_bounds__(TILE) k_gather_rot(Tiles tl, const double *src, const int *perm, double *rhs) {
  constexpr int PB = Dim<D>::PB;
  const int tile = blockIdx.x;
  if (tl.active && !tl.active[tl.node[tile]]) return;
  const int p0 = tl.start[tile], n = tl.cnt[tile] * D * D;
  for (int i = threadIdx.x; i < n; i += TILE) {
    const int pl = i / (D * D), k = i % (D * D);
    const int p = p0 + pl;
    rhs[(size_t)__ldg(perm + p * D + k / D) * D + k % D] = src[(size_t)p * PB + D + k];
  }
}
template <int D> void launch_gather_rot(const Tiles &tl, const double *src, const int *perm, double *rhs, cudaStream_t s) {
  k_gather_rot<D><<<tl.n_tiles, TILE, 0, s>>>(tl, src, perm, rhs);
}
template void launch_gather_rot<2>(const Tiles &, const double *, const int *, double *, cudaStream_t);
template void launch_gather_rot<3>(const Tiles &, const double *, const int *, double *, cudaStream_t);

template <int D> void launch_vec(int op, const Tiles &tl, const VecArgs &a, cudaStream_t s) {
#define MMPGO_VEC_CASE(OP) case OP: k_vec<D, OP><<<tl.n_tiles, TILE, 0, s>>>(tl, a); break;
  switch (op) {
    MMPGO_VEC_CASE(V_CG_INIT) MMPGO_VEC_CASE(V_CG_STEP) MMPGO_VEC_CASE(V_CG_DIR)
    MMPGO_VEC_CASE(V_CG_FINAL) MMPGO_VEC_CASE(V_RETRACT) MMPGO_VEC_CASE(V_DOTS)
    MMPGO_VEC_CASE(V_COPY_ROT) MMPGO_VEC_CASE(V_COPY) MMPGO_VEC_CASE(V_PRECOND)
    MMPGO_VEC_CASE(V_DIFFNORM) MMPGO_VEC_CASE(V_COPY_T) MMPGO_VEC_CASE(V_CG_PRE)
  }
#undef MMPGO_VEC_CASE
}
template void launch_vec<2>(int, const Tiles &, const VecArgs &, cudaStream_t);
template void launch_vec<3>(int, const Tiles &, const VecArgs &, cudaStream_t);

// =============================================================================
// per-node reduction of tile partials (fixed order => deterministic)
// =============================================================================
__global__ void __launch_bounds__(128) k_reduce(const int *tb, const int *te, const double *partials,
                                                double *node_scal) {
  const int node = blockIdx.x;
  const int b = tb[node], e = te[node];
  const int k = threadIdx.x & 7, lane8 = threadIdx.x >> 3;  // 16 groups of NS=8 slots
  __shared__ double sm[16][NS];
  double s = 0.0;
  for (int t = b + lane8; t < e; t += 16) s += partials[(size_t)t * NS + k];
  sm[lane8][k] = s;
  __syncthreads();
  if (threadIdx.x < NS) {
    double x = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) x += sm[i][threadIdx.x];
    node_scal[(size_t)node * NS + threadIdx.x] = x;
  }
}
void launch_reduce(int num_nodes, const int *tb, const int *te, const double *partials, double *node_scal,
                   cudaStream_t s) {
  k_reduce<<<num_nodes, 128, 0, s>>>(tb, te, partials, node_scal);
}

// =============================================================================
// edge-parallel global objective  (DPGOStar::evaluate_f, DPGOStar.cpp:713-761)
// =============================================================================
// Edge data are struct-of-arrays: field f of edge e at val[f * n + e], f = tau, kappa, t[0..2],
// R[0..8] (row-major); idx[0..2][n] = i, j, inter flag.  Every load is coalesced.
template <int D>
__global__ void __launch_bounds__(256) k_edge_objective(int64_t n, const int *idx, const double *val, const double *x,
                                                        int loss, double loss_reg, double *block_partials) {
  constexpr int PB = Dim<D>::PB;
  double sc[1] = {0.0};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int pi = __ldg(idx + e), pj = __ldg(idx + n + e), inter = __ldg(idx + 2 * n + e);
    const double tau = __ldg(val + e), kap = __ldg(val + n + e);
    double t[D], R[D * D], xi[PB], xj[PB];
#pragma unroll
    for (int k = 0; k < D; ++k) t[k] = __ldg(val + (2 + k) * n + e);
#pragma unroll
    for (int k = 0; k < D * D; ++k) R[k] = __ldg(val + (5 + k) * n + e);
#pragma unroll
    for (int k = 0; k < PB; ++k) { xi[k] = x[(size_t)pi * PB + k]; xj[k] = x[(size_t)pj * PB + k]; }
    double et = 0.0, er = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      double s = xi[k] - xj[k];
#pragma unroll
      for (int cc = 0; cc < D; ++cc) s += t[cc] * xi[(1 + cc) * D + k];
      et += s * s;
    }
    if (loss == 0) {
      // M-form: kappa (|Y_i|^2 + |Y_j|^2 - 2 <R^T Y_i, Y_j>)   (DPGO_utils.cpp:500-560)
      double ni = 0.0, nj = 0.0, cr = 0.0;
#pragma unroll
      for (int rr = 0; rr < D; ++rr)
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double s = 0.0;
#pragma unroll
          for (int cc = 0; cc < D; ++cc) s += R[cc * D + rr] * xi[(1 + cc) * D + k];
          cr += s * xj[(1 + rr) * D + k];
          ni += xi[(1 + rr) * D + k] * xi[(1 + rr) * D + k];
          nj += xj[(1 + rr) * D + k] * xj[(1 + rr) * D + k];
        }
      er = ni + nj - 2.0 * cr;
      sc[0] += 0.5 * (tau * et + kap * er);
    } else {
#pragma unroll
      for (int rr = 0; rr < D; ++rr)
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double s = -xj[(1 + rr) * D + k];
#pragma unroll
          for (int cc = 0; cc < D; ++cc) s += R[cc * D + rr] * xi[(1 + cc) * D + k];
          er += s * s;
        }
      const double e2 = tau * et + kap * er;
      if (inter) {
        double rho;
        irls_weight(loss, e2, loss_reg, rho);
        sc[0] += rho;
      } else {
        sc[0] += 0.5 * e2;
      }
    }
  }
  block_reduce_store<1, 256>(sc, block_partials + blockIdx.x);
}
__global__ void k_sum_blocks(int n, const double *bp, double *out) {
  __shared__ double sm[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += bp[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sm[0];
}
template <int D>
void launch_edge_objective(int64_t n_edges, const int *idx, const double *val, const double *x, int loss, double loss_reg,
                           double *block_partials, int *n_blocks_out, cudaStream_t s) {
  int nb = (int)((n_edges + 255) / 256);
  if (nb > 148 * 8) nb = 148 * 8;
  if (nb < 1) nb = 1;
  *n_blocks_out = nb;
  k_edge_objective<D><<<nb, 256, 0, s>>>(n_edges, idx, val, x, loss, loss_reg, block_partials);
}
template void launch_edge_objective<2>(int64_t, const int *, const double *, const double *, int, double, double *, int *,
                                       cudaStream_t);
template void launch_edge_objective<3>(int64_t, const int *, const double *, const double *, int, double, double *, int *,
                                       cudaStream_t);
// out[0] = sum_i v[i * stride], fixed order (one warp)
__global__ void k_sum_strided(int n, const double *v, int stride, double *out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) s += v[(size_t)i * stride];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) out[0] = s;
}
void launch_sum_strided(int n, const double *v, int stride, double *out, cudaStream_t s) {
  k_sum_strided<<<1, 32, 0, s>>>(n, v, stride, out);
}
void launch_sum_blocks(int n_blocks, const double *bp, double *out, cudaStream_t s) {
  k_sum_blocks<<<1, 256, 0, s>>>(n_blocks, bp, out);
}

// =============================================================================
// K2b: translation solve G00 t = rhs
// =============================================================================
// dense path: one warp per output row, t_p = sum_q Ginv[p][q] rhs[q]
template <int D>
__global__ void __launch_bounds__(256) k_dense_solve(const int *node_off, const long long *dense_off,
                                                     const int *node_active, const double *ginv,
                                                     const double *rhs, double *xout) {
  constexpr int PB = Dim<D>::PB;
  const int node = blockIdx.y;
  if (node_active && !node_active[node]) return;
  const int n0 = node_off[node + 1] - node_off[node];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n0) return;
  if (dense_off[node] < 0) return;
  const double *row = ginv + (size_t)dense_off[node] + (size_t)warp * n0;
  const double *b = rhs + (size_t)node_off[node] * D;
  double acc[D];
#pragma unroll
  for (int c = 0; c < D; ++c) acc[c] = 0.0;
  for (int q = lane; q < n0; q += 32) {
    const double gv = row[q];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = fma(gv, b[(size_t)q * D + c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < D; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], o);
  }
  if (lane == 0) {
    double *o = xout + (size_t)(node_off[node] + warp) * PB;
#pragma unroll
    for (int c = 0; c < D; ++c) o[c] = -acc[c];
  }
}

template <int D>
void launch_dense_solve(int num_nodes, const int *node_off, const long long *dense_off, const int *node_active,
                        const double *ginv, const double *rhs, double *xout, int max_n0, cudaStream_t s) {
  const dim3 grid((max_n0 * 32 + 255) / 256, num_nodes);
  k_dense_solve<D><<<grid, 256, 0, s>>>(node_off, dense_off, node_active, ginv, rhs, xout);
}
template void launch_dense_solve<2>(int, const int *, const long long *, const int *, const double *, const double *,
                                    double *, int, cudaStream_t);
template void launch_dense_solve<3>(int, const int *, const long long *, const int *, const double *, const double *,
                                    double *, int, cudaStream_t);

// =============================================================================
// halo pack / unpack
// =============================================================================
template <int D>
__global__ void k_gather_poses(int64_t n, const int *idx, const double *src, double *dst) {
  constexpr int PB = Dim<D>::PB;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * PB) return;
  const int64_t p = i / PB;
  const int k = (int)(i % PB);
  dst[i] = src[(size_t)idx[p] * PB + k];
}
template <int D>
__global__ void k_scatter_poses(int64_t n, const int *idx, const double *src, double *dst) {
  constexpr int PB = Dim<D>::PB;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * PB) return;
  const int64_t p = i / PB;
  const int k = (int)(i % PB);
  dst[(size_t)idx[p] * PB + k] = src[i];
}
template <int D> void launch_gather_poses(int64_t n, const int *idx, const double *src, double *dst, cudaStream_t s) {
  if (n <= 0) return;
  const int64_t tot = n * Dim<D>::PB;
  k_gather_poses<D><<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(n, idx, src, dst);
}
template <int D> void launch_scatter_poses(int64_t n, const int *idx, const double *src, double *dst, cudaStream_t s) {
  if (n <= 0) return;
  const int64_t tot = n * Dim<D>::PB;
  k_scatter_poses<D><<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(n, idx, src, dst);
}
template <int D>
__global__ void k_copy_poses(int64_t n, const int *src_idx, const int *dst_idx, const double *src, double *dst) {
  constexpr int PB = Dim<D>::PB;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * PB) return;
  const int64_t p = i / PB;
  const int k = (int)(i % PB);
  dst[(size_t)dst_idx[p] * PB + k] = src[(size_t)src_idx[p] * PB + k];
}
template <int D>
void launch_copy_poses(int64_t n, const int *src_idx, const int *dst_idx, const double *src, double *dst, cudaStream_t s) {
  if (n <= 0) return;
  const int64_t tot = n * Dim<D>::PB;
  k_copy_poses<D><<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(n, src_idx, dst_idx, src, dst);
}
template void launch_copy_poses<2>(int64_t, const int *, const int *, const double *, double *, cudaStream_t);
template void launch_copy_poses<3>(int64_t, const int *, const int *, const double *, double *, cudaStream_t);
template void launch_gather_poses<2>(int64_t, const int *, const double *, double *, cudaStream_t);
template void launch_gather_poses<3>(int64_t, const int *, const double *, double *, cudaStream_t);
template void launch_scatter_poses<2>(int64_t, const int *, const double *, double *, cudaStream_t);
template void launch_scatter_poses<3>(int64_t, const int *, const double *, double *, cudaStream_t);

// =============================================================================
// layout conversion between the reference's global iterate (column-major ((d+1)N) x d,
// rows [t; R blocks], C++/examples/dist_pgo.cpp:502-511) and the device pose blocks
// =============================================================================
template <int D>
__global__ void k_pack_poses(int64_t n, const int64_t *gid, const double *X, int64_t ld, int64_t N, double *d0,
                             double *d1, double *d2, double *d3, double *d4) {
  constexpr int PB = Dim<D>::PB;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * PB) return;
  const int64_t p = i / PB;
  const int k = (int)(i % PB), r = k / D, c = k % D;
  const int64_t g = gid[p];
  const double v = r == 0 ? X[g + c * ld] : X[N + D * g + (r - 1) + c * ld];
  d0[i] = v;
  if (d1) d1[i] = v;
  if (d2) d2[i] = v;
  if (d3) d3[i] = v;
  if (d4) d4[i] = v;
}
template <int D>
__global__ void k_unpack_poses(int64_t n, const int64_t *gid, const double *src, double *X, int64_t ld, int64_t N) {
  constexpr int PB = Dim<D>::PB;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * PB) return;
  const int64_t p = i / PB;
  const int k = (int)(i % PB), r = k / D, c = k % D;
  const int64_t g = gid[p];
  const double v = src[i];
  if (r == 0) X[g + c * ld] = v;
  else X[N + D * g + (r - 1) + c * ld] = v;
}
template <int D>
void launch_pack_poses(int64_t n, const int64_t *gid, const double *X, int64_t ld, int64_t N, double *d0, double *d1,
                       double *d2, double *d3, double *d4, cudaStream_t s) {
  if (n <= 0) return;
  const int64_t tot = n * Dim<D>::PB;
  k_pack_poses<D><<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(n, gid, X, ld, N, d0, d1, d2, d3, d4);
}
template <int D>
void launch_unpack_poses(int64_t n, const int64_t *gid, const double *src, double *X, int64_t ld, int64_t N,
                         cudaStream_t s) {
  if (n <= 0) return;
  const int64_t tot = n * Dim<D>::PB;
  k_unpack_poses<D><<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(n, gid, src, X, ld, N);
}
template void launch_pack_poses<2>(int64_t, const int64_t *, const double *, int64_t, int64_t, double *, double *, double *, double *, double *, cudaStream_t);
template void launch_pack_poses<3>(int64_t, const int64_t *, const double *, int64_t, int64_t, double *, double *, double *, double *, double *, cudaStream_t);
template void launch_unpack_poses<2>(int64_t, const int64_t *, const double *, double *, int64_t, int64_t, cudaStream_t);
template void launch_unpack_poses<3>(int64_t, const int64_t *, const double *, double *, int64_t, int64_t, cudaStream_t);

// batched polar projection of n row-major d x d blocks (project_to_SO3n / project_to_SO2n,
// C++/DPGO/include/DPGO/DPGO_utils.h:515-565), one block per thread
template <int D> __global__ void k_project_blocks(int64_t n, const double *A, double *U) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double M[D * D], R[D * D];
#pragma unroll
  for (int k = 0; k < D * D; ++k) M[k] = A[i * D * D + k];
  project_to_SOd<D>(M, R);
#pragma unroll
  for (int k = 0; k < D * D; ++k) U[i * D * D + k] = R[k];
}
template <int D> void launch_project_blocks(int64_t n, const double *A, double *U, cudaStream_t s) {
  if (n <= 0) return;
  k_project_blocks<D><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(n, A, U);
}
template void launch_project_blocks<2>(int64_t, const double *, double *, cudaStream_t);
template void launch_project_blocks<3>(int64_t, const double *, double *, cudaStream_t);

}  // namespace mmpgo
