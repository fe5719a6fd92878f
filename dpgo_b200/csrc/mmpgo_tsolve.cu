// K2b: the translation solve  G00 u = rhs  for every local robot node, ONE persistent launch.
//
// Replaces the CHOLMOD solves of DPGOProblem::recover_translations
// (C++/DPGO/include/DPGO/DPGOProblem.h:275-294), DPGOProblem::retract
// (C++/DPGO/src/DPGOProblem.cpp:127-143) and the inner solve of the reduced Hessian-vector
// product (DPGOProblem.cpp:552-577, `L_.solve`).  G00 (n0 x n0: tau-weighted intra-node
// Laplacian + 2 tau per inter-node edge + xi) is block diagonal over nodes, so every node is an
// independent SPD system with d right-hand sides; it is solved to a relative residual of
// `translation_solve_tol` (1e-12) by Jacobi-preconditioned CG in the single-reduction form
//   w = A z;  Ap = w + beta Ap;  p = z + beta p;  alpha = rz / p.Ap              (phase A)
//   x += alpha p;  z -= alpha Ap / diag  (r = diag z);  beta = rz' / rz           (phase B)
// (two rendezvous per iteration, algebraically the standard PCG recurrence).
//
// Execution model.  CTAs are persistent (cooperative launch: all co-resident) and own a static
// round-robin set of "CTA tiles" (<= 256 consecutive poses of one node).  Nothing synchronises
// across the grid.  Each node carries an epoch = the next phase its tiles may execute; a CTA
// runs an event loop: poll the epochs of its tiles' nodes, execute every tile that is ready (its
// own phase counter <= node epoch), then arrive on the nodes' counters.  The last tile of a node
// to arrive sums the per-tile partials in a fixed order (deterministic, independent of
// scheduling and of which other nodes share the GPU), computes alpha / beta / convergence and
// publishes the node's next epoch.  Nodes drift apart freely, so the rendezvous latency of one
// node is hidden behind the tiles of the others; converged nodes retire individually.
//
// Data movement.  Solver vectors are laid out [cta tile][warp][d][32]; the matrix is sliced
// ELLPACK, one slice per 32-pose warp slice, {slot of the neighbour, -tau}[k][32].  A tile's
// vectors, diagonal and ELLPACK rows are contiguous, and arrive in shared memory by bulk
// asynchronous copies (cp.async.bulk -> UBLKCP, completion on an mbarrier), two stages deep, so
// HBM latency is covered by the copy engine instead of by resident warps.  Neighbours inside the
// tile are gathered from shared memory, the others from L2.
#include "mmpgo_kernels.cuh"

namespace mmpgo {

namespace {

constexpr int DONE_BIT = 0x40000000;

__device__ __forceinline__ int ld_relaxed(const int *p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
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
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
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

// ---- bulk asynchronous copies global -> shared, completion on an mbarrier (TMA 1-D path) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

}  // namespace

template <int D> struct TSCfg {
  static constexpr int VEC = CTILE * D;          // doubles of one solver vector per CTA tile
  static constexpr int SELL_CAP = 80;            // 32-entry ELLPACK rows of one CTA tile that fit a stage
  static constexpr int NSTAGE = 2;
  // stage layout (bytes): v0 v1 v2 | diag | X, X = {sell values, sell columns} (phase A) or v3 (phase B)
  static constexpr int OFF_DIAG = 3 * VEC * 8;
  static constexpr int OFF_X = OFF_DIAG + CTILE * 8;
  static constexpr int OFF_COL = OFF_X + SELL_CAP * 32 * 8;
  static constexpr int STAGE_BYTES = OFF_COL + SELL_CAP * 32 * 4;
  static_assert(SELL_CAP * 32 * 12 >= VEC * 8, "phase-B vector must fit the overlay region");
  static constexpr int DYN_BYTES = NSTAGE * STAGE_BYTES;
};

template <int D>
__global__ void __launch_bounds__(256, 2) k_tsolve(TSolveArgs a) {
  typedef TSCfg<D> C;
  constexpr int PB = (D + 1) * D;
  constexpr int NST = C::NSTAGE;
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  int *epoch = a.cnt + a.n_nodes;            // [nodes] next phase the node's tiles may run (| DONE_BIT)
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ uint64_t full[NST];
  // this CTA's tiles: static description, own phase counter, per-pass scratch
  __shared__ int m_node[TS_MAXCT], m_start[TS_MAXCT], m_cnt[TS_MAXCT], m_sell[TS_MAXCT][9];
  __shared__ int t_round[TS_MAXCT];          // next phase of the tile; -1 = retired
  __shared__ int r_k[TS_MAXCT], r_ep[TS_MAXCT], n_ready_s, n_live_s;
  __shared__ double r_coef[TS_MAXCT];
  __shared__ double red[8][3];
  uint32_t ph0 = 0, ph1 = 0;                 // mbarrier parities of the two stages

  const int nb = (a.n_ct - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // <= TS_MAXCT (host-checked)
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NST; ++q) mbar_init(&full[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    n_live_s = 0;
  }
  __syncthreads();
  if (threadIdx.x < nb) {
    const int k = threadIdx.x, ct = blockIdx.x + k * gridDim.x;
    const int node = __ldg(a.ct_node + ct);
    m_node[k] = node; m_start[k] = __ldg(a.ct_start + ct); m_cnt[k] = __ldg(a.ct_cnt + ct);
    const bool on = !(a.active && !__ldg(a.active + node));
    t_round[k] = on ? 0 : -1;
    if (on) atomicAdd(&n_live_s, 1);
  }
  for (int q = threadIdx.x; q < nb * 9; q += blockDim.x) {
    const int k = q / 9, r = q % 9;
    m_sell[k][r] = __ldg(a.sell_ptr + 8 * (blockIdx.x + k * gridDim.x) + r);
  }
  __syncthreads();
  int n_live = n_live_s;

  while (n_live > 0) {
    // ---- poll: which of my tiles may run their next phase?
    if (wi == 0) {
      if (lane == 0) n_ready_s = 0;
      __syncwarp();
      for (int k0 = 0; k0 < nb; k0 += 32) {
        const int k = k0 + lane;
        bool ready = false;
        int ep = 0;
        if (k < nb && t_round[k] >= 0) {
          ep = t_round[k] == 0 ? 0 : ld_relaxed(epoch + m_node[k]);
          ready = (ep & DONE_BIT) || ep >= t_round[k];
        }
        const unsigned m = __ballot_sync(0xffffffffu, ready);
        if (ready) {
          const int pos = n_ready_s + __popc(m & ((1u << lane) - 1));
          r_k[pos] = k; r_ep[pos] = ep;
        }
        __syncwarp();
        if (lane == 0) n_ready_s += __popc(m);
        __syncwarp();
      }
    }
    __syncthreads();
    const int n_ready = n_ready_s;
    if (n_ready == 0) { __nanosleep(200); __syncthreads(); continue; }
    fence_acq_rel();                                     // what the publishers stored is visible from here on
    if (threadIdx.x < n_ready) {
      // scalars of the phase: beta for phase A (odd), alpha for phase B (even)
      const int j = threadIdx.x, k = r_k[j], r = t_round[k];
      double coef = 0.0;
      if (!(r_ep[j] & DONE_BIT) && r > 0) coef = __ldcg(a.nstate + (size_t)m_node[k] * 8 + ((r & 1) ? 3 : 2));
      r_coef[j] = coef;
    }
    if (threadIdx.x == 0) asm volatile("fence.proxy.async;" ::: "memory");   // bulk copies read others' stores
    __syncthreads();

    // ---- the tile's vectors (and its ELLPACK rows) arrive by bulk copy, two stages deep;
    // thread 0 is the producer.  kind: 0 init, 1 phase A, 2 phase B, 3 publish the result
    auto kind_of = [&](int j) {
      if (r_ep[j] & DONE_BIT) return 3;
      const int r = t_round[r_k[j]];
      return r == 0 ? 0 : ((r & 1) ? 1 : 2);
    };
    auto issue = [&](int j) {
      const int kd = kind_of(j);
      if (kd == 0 || kd == 3) return;                    // direct loads, once per solve
      const int k = r_k[j], ct = blockIdx.x + k * gridDim.x, stg = j % NST;
      unsigned char *sb = dyn + (size_t)stg * C::STAGE_BYTES;
      const size_t v0 = (size_t)ct * C::VEC;
      const int r0 = m_sell[k][0], rows = m_sell[k][8] - r0;
      const bool phA = kd == 1, sell_staged = phA && rows <= C::SELL_CAP;
      uint32_t bytes = 3 * C::VEC * 8 + CTILE * 8;
      if (phA) { if (sell_staged) bytes += rows * 32 * 12; }
      else bytes += C::VEC * 8;
      mbar_expect_tx(&full[stg], bytes);
      const double *s0 = phA ? a.z : a.p, *s1 = a.ap, *s2 = phA ? a.p : a.x;
      bulk_g2s(sb, s0 + v0, C::VEC * 8, &full[stg]);
      bulk_g2s(sb + C::VEC * 8, s1 + v0, C::VEC * 8, &full[stg]);
      bulk_g2s(sb + 2 * C::VEC * 8, s2 + v0, C::VEC * 8, &full[stg]);
      bulk_g2s(sb + C::OFF_DIAG, a.diag_s + (size_t)ct * CTILE, CTILE * 8, &full[stg]);
      if (phA) {
        if (sell_staged && rows > 0) {
          bulk_g2s(sb + C::OFF_X, a.sell_val + (size_t)r0 * 32, rows * 32 * 8, &full[stg]);
          bulk_g2s(sb + C::OFF_COL, a.sell_col + (size_t)r0 * 32, rows * 32 * 4, &full[stg]);
        }
      } else {
        bulk_g2s(sb + C::OFF_X, a.z + v0, C::VEC * 8, &full[stg]);
      }
    };
    if (threadIdx.x == 0)
      for (int j = 0; j < min(NST, n_ready); ++j) issue(j);

    for (int j = 0; j < n_ready; ++j) {
      const int k = r_k[j], ct = blockIdx.x + k * gridDim.x, stg = j % NST, kd = kind_of(j);
      const bool valid = 32 * wi + lane < m_cnt[k];
      const int p = m_start[k] + 32 * wi + lane;                     // own pose index
      const size_t vb = (size_t)(8 * ct + wi) * (32 * D) + lane;     // slot of (pose, column 0)
      double part[3] = {0.0, 0.0, 0.0};
      if (kd == 3) {
        // node finished: publish u into the pose array, t = -u
        if (valid) {
#pragma unroll
          for (int c = 0; c < D; ++c) a.xio[(size_t)p * PB + c] = -__ldcg(a.x + vb + 32 * c);
        }
      } else if (kd == 0) {
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
      } else {
        const unsigned char *sb = dyn + (size_t)stg * C::STAGE_BYTES;
        const double *sv0 = reinterpret_cast<const double *>(sb) + wi * (32 * D) + lane;
        const double *sv1 = sv0 + C::VEC, *sv2 = sv0 + 2 * C::VEC;
        if (stg == 0) { mbar_wait(&full[0], ph0); ph0 ^= 1; }
        else { mbar_wait(&full[1], ph1); ph1 ^= 1; }
        const double dg = reinterpret_cast<const double *>(sb + C::OFF_DIAG)[32 * wi + lane];
        if (kd == 1) {
          // phase A: w = A z ; Ap = w + beta Ap ; p = z + beta p ; partial p.Ap
          const double beta = r_coef[j];
          const int r0 = m_sell[k][0];
          const int s0 = m_sell[k][wi] - r0, s1 = m_sell[k][wi + 1] - r0;
          const bool sell_staged = m_sell[k][8] - r0 <= C::SELL_CAP;
          const double *sval = reinterpret_cast<const double *>(sb + C::OFF_X);
          const int *scol = reinterpret_cast<const int *>(sb + C::OFF_COL);
          const double *zt = reinterpret_cast<const double *>(sb);   // this tile's z, slots relative to tile_lo
          const int tile_lo = ct * C::VEC;
          double zo[D], acc[D];
#pragma unroll
          for (int c = 0; c < D; ++c) { zo[c] = sv0[32 * c]; acc[c] = dg * zo[c]; }
          // the slice loop is warp-uniform (padded entries have value 0); entries are taken four
          // at a time so that the gathers of out-of-tile neighbours overlap
          for (int s = s0; s < s1; s += 4) {
            int slot[4];
            double av[4], zq[4][D];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const bool in = s + q < s1;
              if (sell_staged) {
                slot[q] = in ? scol[(s + q) * 32 + lane] : (int)vb;
                av[q] = in ? sval[(s + q) * 32 + lane] : 0.0;
              } else {
                slot[q] = in ? __ldg(a.sell_col + (size_t)(r0 + s + q) * 32 + lane) : (int)vb;
                av[q] = in ? __ldg(a.sell_val + (size_t)(r0 + s + q) * 32 + lane) : 0.0;
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const unsigned rel = (unsigned)(slot[q] - tile_lo);
              if (rel < (unsigned)C::VEC) {
#pragma unroll
                for (int c = 0; c < D; ++c) zq[q][c] = zt[rel + 32 * c];
              } else {
#pragma unroll
                for (int c = 0; c < D; ++c) zq[q][c] = __ldcg(a.z + (size_t)slot[q] + 32 * c);
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int c = 0; c < D; ++c) acc[c] = fma(av[q], zq[q][c], acc[c]);
          }
          if (valid) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
              const double apv = acc[c] + beta * sv1[32 * c];
              const double pv = zo[c] + beta * sv2[32 * c];
              a.ap[vb + 32 * c] = apv;
              a.p[vb + 32 * c] = pv;
              part[0] += pv * apv;
            }
          }
        } else {
          // phase B: x += alpha p ; z -= alpha Ap / diag (r = diag z) ; partials rz, rr
          const double alpha = r_coef[j];
          const double *sv3 = reinterpret_cast<const double *>(sb + C::OFF_X) + wi * (32 * D) + lane;
          if (valid) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
              const double xv = sv2[32 * c] + alpha * sv0[32 * c];
              const double rv = dg * sv3[32 * c] - alpha * sv1[32 * c];
              const double zv = rv / dg;
              a.x[vb + 32 * c] = xv;
              a.z[vb + 32 * c] = zv;
              part[0] += rv * zv; part[2] += rv * rv;
            }
          }
        }
      }
      if (kd != 3) {
        const double p0 = warp_sum(part[0]), p1 = warp_sum(part[1]), p2 = warp_sum(part[2]);
        if (lane == 0) { red[wi][0] = p0; red[wi][1] = p1; red[wi][2] = p2; }
      }
      __syncthreads();                                   // stage buffer free again, red[] complete
      if (threadIdx.x == 0 && j + NST < n_ready) issue(j + NST);
      if (kd != 3 && threadIdx.x < 3) {
        // the tile's partial, summed over the warps in a fixed order
        double sacc = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) sacc += red[q][threadIdx.x];
        a.partials[(size_t)ct * 4 + threadIdx.x] = sacc;
      }
      __syncthreads();                                   // red[] free again
    }

    // ---- warp 0: arrive on the nodes of the tiles just executed
    if (wi == 0) {
      for (int j0 = 0; j0 < n_ready; j0 += 32) {
        const int j = j0 + lane;
        const bool mine = j < n_ready;
        int old = -1, node = 0, rnd = 0;
        bool fin = false;
        if (mine) {
          const int k = r_k[j];
          node = m_node[k]; rnd = t_round[k];
          fin = (r_ep[j] & DONE_BIT) != 0;
          if (fin) { t_round[k] = -1; atomicSub(&n_live_s, 1); }
          else {
            t_round[k] = rnd + 1;
            old = atom_add_acq_rel(a.cnt + node, 1);     // release: this CTA's stores of the phase (ordered by bar.sync)
          }
        }
        unsigned last = __ballot_sync(0xffffffffu, mine && !fin &&
                                      old == __ldg(a.node_cte + node) - __ldg(a.node_ctb + node) - 1);
        while (last) {
          // a tile of this CTA was the last of its node in this phase: reduce the node, set its scalars
          const int src = __ffs(last) - 1;
          last &= last - 1;
          const int nd = __shfl_sync(0xffffffffu, node, src);
          const int round = __shfl_sync(0xffffffffu, rnd, src);
          double *nst = a.nstate + (size_t)nd * 8;
          const int cb = __ldg(a.node_ctb + nd), ce = __ldg(a.node_cte + nd);
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
            if (done && a.stats) {
              const unsigned long long it = (unsigned long long)(round / 2);
              atomicAdd(a.stats, it);
              atomicAdd(a.stats + 1, it * (unsigned long long)(__ldg(a.node_off + nd + 1) - __ldg(a.node_off + nd)));
            }
            st_release(epoch + nd, (round + 1) | (done ? DONE_BIT : 0));
          }
        }
      }
    }
    __syncthreads();
    n_live = n_live_s;
  }
}

template <int D> int launch_tsolve(const TSolveArgs &a, int grid, cudaStream_t s) {
  TSolveArgs args = a;
  void *params[] = {&args};
  return (int)cudaLaunchCooperativeKernel((const void *)k_tsolve<D>, dim3(grid), dim3(256), params,
                                          TSCfg<D>::DYN_BYTES, s);
}
template int launch_tsolve<2>(const TSolveArgs &, int, cudaStream_t);
template int launch_tsolve<3>(const TSolveArgs &, int, cudaStream_t);

template <int D> int tsolve_max_grid(int device) {
  int per_sm = 0, sms = 0;
  if (cudaFuncSetAttribute(k_tsolve<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, TSCfg<D>::DYN_BYTES) != cudaSuccess)
    return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tsolve<D>, 256, TSCfg<D>::DYN_BYTES) != cudaSuccess)
    return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  return per_sm * sms;
}
template int tsolve_max_grid<2>(int);
template int tsolve_max_grid<3>(int);

}  // namespace mmpgo
