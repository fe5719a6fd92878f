// K2b: the translation solve  G00 u = rhs  for every local robot node, ONE persistent launch.
//
// Replaces the CHOLMOD solves of DPGOProblem::recover_translations
// (C++/DPGO/include/DPGO/DPGOProblem.h:275-294), DPGOProblem::retract
// (C++/DPGO/src/DPGOProblem.cpp:127-143) and the inner solve of the reduced Hessian-vector
// product (DPGOProblem.cpp:552-577, `L_.solve`).  G00 (n0 x n0: tau-weighted intra-node
// Laplacian + 2 tau per inter-node edge + xi) is block diagonal over nodes, so every node is an
// independent SPD system with d right-hand sides; it is solved to a relative residual of
// `translation_solve_tol` (1e-12) by Jacobi-preconditioned CG in the single-reduction form
//   w = A z;  Ap = w + beta Ap;  p = z + beta p;  alpha = rz / p.Ap        (phase A)
//   x += alpha p;  r -= alpha Ap;  z = r / diag;  beta = rz' / rz           (phase B)
// (two synchronisation points per iteration, algebraically the standard PCG recurrence).
//
// Execution model: warps are persistent and own a static round-robin set of 32-pose "warp
// tiles".  Nothing synchronises across the grid: the tiles of one node rendezvous on that
// node's arrival counter; the last arriver sums the per-tile partials in a fixed order
// (deterministic, independent of scheduling and of which other nodes share the GPU), computes
// the node's alpha / beta / convergence and publishes the node's next phase number (epoch).
// Converged nodes drop out individually.  Requires all CTAs co-resident (cooperative launch).
//
// Layout: solver vectors are [warp tile][d][32] (every load/store is one contiguous 256-byte
// line per column); the matrix is sliced ELLPACK with one slice per warp tile, entries
// [k][32] of {slot of the neighbour, -tau}.
#include "mmpgo_kernels.cuh"

namespace mmpgo {

namespace {

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int atom_add_acq_rel(int *p, int v) {
  int o;
  asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
  return o;
}
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  return x;   // valid in lane 0
}

// Fixed-order sum of K partial columns over the CTA tiles [c0, c1) of one node (one warp).
template <int K>
__device__ __forceinline__ void node_sum(const double *partials, int c0, int c1, int lane, double (&out)[K]) {
  double s[K];
#pragma unroll
  for (int k = 0; k < K; ++k) s[k] = 0.0;
  for (int c = c0 + lane; c < c1; c += 32) {
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] += __ldcg(partials + (size_t)c * 4 + k);
  }
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = __shfl_sync(0xffffffffu, warp_sum(s[k]), 0);
}

// grid-wide barrier on a monotone counter (zeroed before the launch); all CTAs are co-resident
__device__ __forceinline__ void grid_barrier(int *counter, int generation) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atom_add_acq_rel(counter, 1);
    const int target = generation * (int)gridDim.x;
    while (ld_acquire(counter) < target) { }
    __threadfence();
  }
  __syncthreads();
}

}  // namespace

constexpr int MAXCT = 32;   // CTA tiles staged per batch (metadata + partials in shared memory)

template <int D>
__global__ void __launch_bounds__(256, 4) k_tsolve(TSolveArgs a) {
  constexpr int PB = (D + 1) * D;
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  int *done_count = a.cnt + a.n_nodes;     // nodes that have finished
  int *barrier = a.cnt + a.n_nodes + 1;
  int *finish = a.cnt + a.n_nodes + 2;     // round in which every CTA leaves (0 = not known yet)
  const int last_round = 2 * a.max_iters + 3;
  // per-batch metadata of this CTA's tiles: static part + the node scalars of the round
  __shared__ int m_node[MAXCT], m_start[MAXCT], m_cnt[MAXCT], m_state[MAXCT], m_sell[MAXCT][9];
  __shared__ double m_coef[MAXCT];
  __shared__ double red[MAXCT][8][3];

  for (int round = 0; round <= last_round; ++round) {
    for (int base = blockIdx.x; base < a.n_ct; base += gridDim.x * MAXCT) {
      const int nb = min(MAXCT, (a.n_ct - base + (int)gridDim.x - 1) / (int)gridDim.x);
      __syncthreads();
      // ---- stage metadata: m_state 0 = run, 1 = skip, 2 = publish result
      if (threadIdx.x < nb) {
        const int k = threadIdx.x, ct = base + k * gridDim.x;
        const int node = __ldg(a.ct_node + ct);
        m_node[k] = node; m_start[k] = __ldg(a.ct_start + ct); m_cnt[k] = __ldg(a.ct_cnt + ct);
        int st = 0;
        double coef = 0.0;
        if (a.active && !__ldg(a.active + node)) st = 1;
        else if (round > 0) {
          const double *nst = a.nstate + (size_t)node * 8;
          const int done_at = (int)__ldcg(nst + 6);       // finished-in-round + 1, 0 = running
          if (done_at) st = done_at == round ? 2 : 1;
          else coef = __ldcg(nst + ((round & 1) ? 3 : 2));  // beta for phase A, alpha for phase B
        }
        m_state[k] = st; m_coef[k] = coef;
      }
      for (int q = threadIdx.x; q < nb * 9; q += blockDim.x) {
        const int k = q / 9, r = q % 9;
        m_sell[k][r] = __ldg(a.sell_ptr + 8 * (base + k * gridDim.x) + r);
      }
      __syncthreads();
      // ---- sweep
      for (int k = 0; k < nb; ++k) {
        const int st = m_state[k];
        if (st == 1) continue;
        const int ct = base + k * gridDim.x;
        const bool valid = 32 * wi + lane < m_cnt[k];
        const int p = m_start[k] + 32 * wi + lane;                   // own pose index
        const size_t vb = (size_t)(8 * ct + wi) * (32 * D) + lane;   // slot of (p, column 0)
        if (st == 2) {
          // node finished in the previous round: publish u into the pose array, t = -u
          if (valid) {
#pragma unroll
            for (int c = 0; c < D; ++c) a.xio[(size_t)p * PB + c] = -__ldcg(a.x + vb + 32 * c);
          }
          continue;
        }
        double part[3] = {0.0, 0.0, 0.0};
        if (round == 0) {
          // r = b - A x0 ; z = r / diag ; p = Ap = 0 ; partials rz, bb, rr
          if (valid) {
            double x0[D], acc[D], b[D];
            const double dg = __ldg(a.d00 + p);
#pragma unroll
            for (int c = 0; c < D; ++c) { x0[c] = 0.0; acc[c] = 0.0; b[c] = a.rhs[(size_t)p * D + c]; }
            if (a.warm) {
#pragma unroll
              for (int c = 0; c < D; ++c) { x0[c] = -a.xio[(size_t)p * PB + c]; acc[c] = dg * x0[c]; }
              const int e0 = __ldg(a.rowptr + p), e1 = __ldg(a.rowptr + p + 1);
              for (int e = e0; e < e1; ++e) {
                const double av = __ldg(a.a00 + e);
                const double *xq = a.xio + (size_t)__ldg(a.col + e) * PB;
#pragma unroll
                for (int c = 0; c < D; ++c) acc[c] = fma(av, -xq[c], acc[c]);
              }
            }
#pragma unroll
            for (int c = 0; c < D; ++c) {
              const double rv = b[c] - acc[c], zv = rv / dg;
              a.x[vb + 32 * c] = x0[c];
              a.z[vb + 32 * c] = zv;
              a.p[vb + 32 * c] = 0.0;
              a.ap[vb + 32 * c] = 0.0;
              part[0] += rv * zv; part[1] += b[c] * b[c]; part[2] += rv * rv;
            }
          }
        } else if (a.mode == 1) {
        } else if (round & 1) {
          // phase A: w = A z ; Ap = w + beta Ap ; p = z + beta p ; partial p.Ap
          const double beta = m_coef[k];
          const int s0 = m_sell[k][wi], s1 = m_sell[k][wi + 1];
          double zo[D], acc[D], apo[D], po[D];
#pragma unroll
          for (int c = 0; c < D; ++c) { zo[c] = 0.0; acc[c] = 0.0; apo[c] = 0.0; po[c] = 0.0; }
          if (valid) {
            const double dg = __ldg(a.d00 + p);
#pragma unroll
            for (int c = 0; c < D; ++c) {
              zo[c] = __ldcg(a.z + vb + 32 * c);
              apo[c] = __ldcg(a.ap + vb + 32 * c);
              po[c] = __ldcg(a.p + vb + 32 * c);
              acc[c] = dg * zo[c];
            }
          }
          // the slice loop is warp-uniform (padded entries have value 0); entries are fetched
          // four at a time so that the dependent gathers overlap
          for (int s = s0; s < s1; s += 4) {
            int slot[4];
            double av[4], zq[4][D];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const bool in = s + q < s1;
              slot[q] = in ? __ldg(a.sell_col + (size_t)(s + q) * 32 + lane) : (int)vb;
              av[q] = in ? __ldg(a.sell_val + (size_t)(s + q) * 32 + lane) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int c = 0; c < D; ++c) zq[q][c] = __ldcg(a.z + (size_t)slot[q] + 32 * c);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int c = 0; c < D; ++c) acc[c] = fma(av[q], zq[q][c], acc[c]);
          }
          if (valid) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
              const double apv = acc[c] + beta * apo[c];
              const double pv = zo[c] + beta * po[c];
              a.ap[vb + 32 * c] = apv;
              a.p[vb + 32 * c] = pv;
              part[0] += pv * apv;
            }
          }
        } else {
          // phase B: x += alpha p ; z -= alpha Ap / diag (r = diag z) ; partials rz, rr
          const double alpha = m_coef[k];
          if (valid) {
            const double dg = __ldg(a.d00 + p);
#pragma unroll
            for (int c = 0; c < D; ++c) {
              const double pv = __ldcg(a.p + vb + 32 * c), apv = __ldcg(a.ap + vb + 32 * c);
              const double xv = __ldcg(a.x + vb + 32 * c) + alpha * pv;
              const double rv = dg * __ldcg(a.z + vb + 32 * c) - alpha * apv;
              const double zv = rv / dg;
              a.x[vb + 32 * c] = xv;
              a.z[vb + 32 * c] = zv;
              part[0] += rv * zv; part[2] += rv * rv;
            }
          }
        }
        const double p0 = warp_sum(part[0]), p1 = warp_sum(part[1]), p2 = warp_sum(part[2]);
        if (lane == 0) { red[k][wi][0] = p0; red[k][wi][1] = p1; red[k][wi][2] = p2; }
      }
      __syncthreads();
      if (wi != 0) continue;
      // ---- warp 0: CTA partials in a fixed order, then rendezvous on the nodes
      const bool mine = lane < nb && m_state[lane] == 0;
      int old = -1, node = 0;
      if (mine) {
        node = m_node[lane];
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { s0 += red[lane][q][0]; s1 += red[lane][q][1]; s2 += red[lane][q][2]; }
        double *pw = a.partials + (size_t)(base + lane * gridDim.x) * 4;
        pw[0] = s0; pw[1] = s1; pw[2] = s2;
        __threadfence();
        if (!(a.mode == 2 && round > 0 && round < 40)) old = atom_add_acq_rel(a.cnt + node, 1);
      }
      unsigned last = __ballot_sync(0xffffffffu, mine && old == __ldg(a.node_cte + node) - __ldg(a.node_ctb + node) - 1);
      while (last) {
        // a tile of this CTA was the last of its node in this round: reduce the node, set its scalars
        const int src = __ffs(last) - 1;
        last &= last - 1;
        const int nd = __shfl_sync(0xffffffffu, node, src);
        double *nst = a.nstate + (size_t)nd * 8;
        const int cb = __ldg(a.node_ctb + nd), ce = __ldg(a.node_cte + nd);
        __threadfence();
        bool done = false;
        if (round == 0) {
          double s[3];
          node_sum<3>(a.partials, cb, ce, lane, s);
          if (lane == 0) { nst[0] = s[0]; nst[1] = s[1]; nst[5] = s[2]; nst[2] = 0.0; nst[3] = 0.0; nst[4] = 0.0; }
          done = !(s[0] > 0.0) || !(s[2] > a.tol2 * s[1]);
        } else if (round & 1) {
          double s[1];
          node_sum<1>(a.partials, cb, ce, lane, s);
          if (s[0] > 0.0) { if (lane == 0) nst[2] = __ldcg(nst + 0) / s[0]; }
          else done = true;
        } else {
          double s[3];
          node_sum<3>(a.partials, cb, ce, lane, s);
          const double rz = __ldcg(nst + 0), bb = __ldcg(nst + 1), it = __ldcg(nst + 4) + 1.0;
          __syncwarp();
          if (lane == 0) { nst[3] = s[0] / rz; nst[0] = s[0]; nst[5] = s[2]; nst[4] = it; }
          done = !(s[2] > a.tol2 * bb) || !(s[0] > 0.0) || it >= (double)a.max_iters;
        }
        if (lane == 0) {
          a.cnt[nd] = 0;
          if (round == 0) nst[6] = 0.0;
          if (done) {
            nst[6] = (double)(round + 1);
            const unsigned long long it = (unsigned long long)(round / 2);
            if (a.stats) {
              atomicAdd(a.stats, it);
              atomicAdd(a.stats + 1, it * (unsigned long long)(__ldg(a.node_off + nd + 1) - __ldg(a.node_off + nd)));
            }
            __threadfence();
            if (atomicAdd(done_count, 1) == a.n_active - 1) st_release(finish, round + 1);
          }
        }
      }
    }
    // leave one round after the last node finished (its tiles have published u by then); the
    // flag is set before the barrier of the finishing round, so every CTA sees it in time
    if (round > 0 && ld_acquire(finish) == round) break;
    grid_barrier(barrier, round + 1);
  }
}

template <int D> int launch_tsolve(const TSolveArgs &a, int grid, cudaStream_t s) {
  TSolveArgs args = a;
  void *params[] = {&args};
  return (int)cudaLaunchCooperativeKernel((const void *)k_tsolve<D>, dim3(grid), dim3(256), params, 0, s);
}
template int launch_tsolve<2>(const TSolveArgs &, int, cudaStream_t);
template int launch_tsolve<3>(const TSolveArgs &, int, cudaStream_t);

template <int D> int tsolve_max_grid(int device) {
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tsolve<D>, 256, 0) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  return per_sm * sms;
}
template int tsolve_max_grid<2>(int);
template int tsolve_max_grid<3>(int);

}  // namespace mmpgo
