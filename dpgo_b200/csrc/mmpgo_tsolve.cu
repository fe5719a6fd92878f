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
// node is hidden behind the tiles of the others; converged nodes retire individually.  Inside a
// CTA the work is warp-specialised: consumer groups (one pose per thread) do the arithmetic, a
// boundary warp owns all global synchronisation, two dispatch warps feed the copy ring.
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
__device__ __forceinline__ void red_add_release(int *p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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
  static constexpr int RL = 3 * VEC + CTILE;     // doubles of one tile record {x, p, Ap, diag}
  static constexpr int WPT = TS_WPT;             // consumer warps per tile (one pose per thread)
  static constexpr int NGRP = TS_NGRP;           // consumer groups (tiles in arithmetic at the same time)
  static constexpr int HV = 2 * TS_HALO + 1;     // tiles of z staged for phase A (own tile in the middle)
  static constexpr int SELL_CAP = 40;            // 32-entry ELLPACK rows of one CTA tile that fit a stage
  static constexpr int NSTAGE = 5;
  static constexpr int THREADS = NGRP * CTILE + 96;   // consumer groups + boundary warp + two dispatch warps
  // stage layout (bytes):  Z (HV VEC) | R (record or its {p, Ap, diag} tail) | S (packed ELLPACK rows)
  //   phase A: Z = z of the tiles ct-HALO .. ct+HALO, R = {p, Ap, diag}, S = the tile's rows   (3 bulk copies)
  //   phase B: Z = z of the tile, R = {x, p, Ap, diag}                                        (2 bulk copies)
  static constexpr int OFF_R = HV * VEC * 8;
  static constexpr int OFF_S = OFF_R + RL * 8;
  static constexpr int STAGE_BYTES = OFF_S + SELL_CAP * 384;
  static constexpr int DYN_BYTES = NSTAGE * STAGE_BYTES;
};

namespace {
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta_shared(const int *p) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_shared(int *p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void consumer_barrier(int group) {
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(CTILE) : "memory");
}
}  // namespace

// tile flags (shared memory): scheduler 0 -> 1 (dispatched), consumers 1 -> 2 (phase executed),
// scheduler 2 -> 0 (arrived on the node)
template <int D>
__global__ void __launch_bounds__(TS_NGRP * CTILE + 96, 1) k_tsolve(TSolveArgs a) {
  typedef TSCfg<D> C;
  constexpr int PB = (D + 1) * D;
  constexpr int NST = C::NSTAGE;
  constexpr int WPT = C::WPT;
  const int lane = threadIdx.x & 31, wg = threadIdx.x >> 5;   // the last warp schedules, the others consume
  const int wi = wg % WPT, grp = wg / WPT;                    // consumer: warp inside its group, group
  int *epoch = a.cnt + a.n_nodes;            // [nodes] next phase the node's tiles may run (| DONE_BIT)
  extern __shared__ __align__(128) unsigned char dyn[];
  // full barriers: one per use modulo NGRP * NST, so that a barrier is always waited on by the same
  // consumer group (with one barrier per stage the groups would alternate on it, and a group
  // running ahead could mistake another group's unfinished phase for its own)
  __shared__ uint64_t full[C::NGRP * NST], empty[NST];
  __shared__ int m_node[TS_MAXCT], m_start[TS_MAXCT], m_cnt[TS_MAXCT], m_sell[TS_MAXCT][TS_WPT + 1];
  __shared__ int t_round[TS_MAXCT];          // 0: tile of an active node, -1: masked
  __shared__ int d_k[NST], d_kind[NST], d_seg[NST];   // work descriptor of a stage
  __shared__ int sg_done[TS_MAXCT];           // per segment: tiles of the current phase handed back
  __shared__ double d_coef[NST];
  __shared__ double red[TS_NGRP][2][TS_WPT][3];
  __shared__ int n_live_s;

  // CTA tiles are dealt to the CTAs in chunks of `chunk` consecutive tiles, round-robin: a node's
  // rendezvous involves only the CTAs that hold one of its chunks, yet a node that needs many more
  // iterations than the others still keeps ~tiles/chunk CTAs busy in the tail
  // (the dealing is a host-built table, a.cta_ptr / a.cta_tiles: nodes known to need many more
  // iterations than the rest are dealt in smaller chunks, i.e. spread over more CTAs)
  __shared__ int m_tile[TS_MAXCT];
  const int tb0 = __ldg(a.cta_ptr + blockIdx.x);
  const int nb = __ldg(a.cta_ptr + blockIdx.x + 1) - tb0;                            // <= TS_MAXCT (host-checked)
  for (int k = threadIdx.x; k < nb; k += blockDim.x) m_tile[k] = __ldg(a.cta_tiles + tb0 + k);
  __syncthreads();
  auto tile_of = [&](int k) { return m_tile[k]; };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NST; ++q) mbar_init(&empty[q], 1);
#pragma unroll
    for (int q = 0; q < C::NGRP * NST; ++q) mbar_init(&full[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    n_live_s = 0;
  }
  __syncthreads();
  if (threadIdx.x < nb) {
    const int k = threadIdx.x, ct = tile_of(k);
    const int node = __ldg(a.ct_node + ct);
    m_node[k] = node; m_start[k] = __ldg(a.ct_start + ct); m_cnt[k] = __ldg(a.ct_cnt + ct);
    const bool on = !(a.active && !__ldg(a.active + node));
    t_round[k] = on ? 0 : -1;
    if (on) atomicAdd(&n_live_s, 1);
  }
  for (int q = threadIdx.x; q < nb * (WPT + 1); q += blockDim.x) {
    const int k = q / (WPT + 1), r = q % (WPT + 1);
    m_sell[k][r] = __ldg(a.sell_ptr + WPT * tile_of(k) + r);
  }
  __syncthreads();

  // A segment = this CTA's tiles of one node (contiguous).  Two service warps:
  //   boundary warp : per segment and phase one arrival on the node, the node reduction when it is
  //                   the last arriver, and the poll for the node's next epoch -> opens the phase
  //   dispatch warps: two; move tiles of open phases into the copy ring (shared-memory state only)
  __shared__ int sg_k0[TS_MAXCT], sg_n[TS_MAXCT], sg_node[TS_MAXCT], sg_target[TS_MAXCT];
  __shared__ int sg_kind[TS_MAXCT];            // kind of the open phase: 0 init, 1 phase A, 2 phase B, 3 publish
  __shared__ double sg_coef[TS_MAXCT];         // beta (phase A) / alpha (phase B) of the open phase
  __shared__ int sg_open[TS_MAXCT];            // last phase opened for dispatch   (boundary -> dispatch, release)
  __shared__ int sg_claim[TS_MAXCT];           // tiles claimed so far by the dispatch warps (all phases)
  __shared__ int n_seg_s, all_done_s, use_ctr_s;
  if (wg >= C::NGRP * WPT) {
    if (wg == C::NGRP * WPT && lane == 0) {
      int ns = 0;
      for (int k = 0; k < nb; ++k) {
        if (t_round[k] < 0) continue;                    // masked node
        if (ns > 0 && sg_node[ns - 1] == m_node[k] && sg_k0[ns - 1] + sg_n[ns - 1] == k) { sg_n[ns - 1]++; continue; }
        sg_k0[ns] = k; sg_n[ns] = 1; sg_node[ns] = m_node[k];
        // CTAs that share the node = arrivals the node expects per phase
        const int nd = m_node[k];
        sg_target[ns] = __ldg(a.node_parts + nd);
        sg_kind[ns] = 0; sg_coef[ns] = 0.0; sg_open[ns] = 0; sg_done[ns] = 0; sg_claim[ns] = 0;   // phase 0 (init) is open
        ++ns;
      }
      n_seg_s = ns;
      all_done_s = 0;
      use_ctr_s = 0;
    }
    asm volatile("bar.sync 15, 96;" ::: "memory");     // the three service warps
    const int n_seg = n_seg_s;

    if (wg == C::NGRP * WPT) {
      // =========================== boundary warp ===========================
      // sg_state (private): 1 phase running, 0 waiting for the node's epoch, -1 retired
      __shared__ int sg_state[TS_MAXCT], sg_round[TS_MAXCT];
      for (int q = lane; q < n_seg; q += 32) { sg_state[q] = 1; sg_round[q] = 0; }
      __syncwarp();
      int n_live = n_seg;
      int ep_polled = 0;
      while (n_live > 0) {
        bool progress = false;
        for (int s0 = 0; s0 < n_seg; s0 += 32) {
          const int sidx = s0 + lane;
          const bool have = sidx < n_seg;
          const int state = have ? sg_state[sidx] : -1;
          int old = -1, node = 0, rnd = 0;
          bool arrived = false, retired = false;
          if (state == 1 && ld_acquire_cta_shared(&sg_done[sidx]) == sg_n[sidx]) {
            // every tile of the segment has executed the phase
            node = sg_node[sidx]; rnd = sg_round[sidx];
            sg_done[sidx] = 0;
            if (sg_kind[sidx] == 3) { sg_state[sidx] = -1; retired = true; }
            else {
              old = atom_add_acq_rel(a.cnt + node, 1);   // release: the consumers' stores of the phase
              sg_round[sidx] = rnd + 1; sg_state[sidx] = 0;
              arrived = true;
            }
          }
          n_live -= __popc(__ballot_sync(0xffffffffu, retired));
          if (__ballot_sync(0xffffffffu, arrived || retired)) progress = true;
          unsigned last = __ballot_sync(0xffffffffu, arrived && old == sg_target[sidx] - 1);
          while (last) {
            // this CTA was the last of the node's CTAs in this phase: reduce the node, set its scalars
            const int src = __ffs(last) - 1;
            last &= last - 1;
            const int nd = __shfl_sync(0xffffffffu, node, src);
            const int round = __shfl_sync(0xffffffffu, rnd, src);
            double *nst = a.nstate + (size_t)nd * 8;
            const int cb = __ldg(a.node_ctb + nd), ce = __ldg(a.node_cte + nd);
            bool done = false, handoff = false, unconv = false;
            if (round == 0) {
              double sm[3];
              node_sum<3>(a.partials, cb, ce, lane, sm);
              if (lane == 0) { nst[0] = sm[0]; nst[1] = sm[1]; nst[5] = sm[2]; nst[2] = 0.0; nst[3] = 0.0; nst[4] = 0.0; }
              done = !(sm[0] > 0.0) || !(sm[2] > a.tol2 * sm[1]);
            } else if (round & 1) {
              double sm[1];
              node_sum<1>(a.partials, cb, ce, lane, sm);
              if (sm[0] > 0.0) { if (lane == 0) nst[2] = __ldcg(nst + 0) / sm[0]; }
              else done = true;
            } else {
              double sm[3];
              node_sum<3>(a.partials, cb, ce, lane, sm);
              const double rz = __ldcg(nst + 0), bb = __ldcg(nst + 1), it = __ldcg(nst + 4) + 1.0;
              __syncwarp();
              if (lane == 0) { nst[3] = sm[0] / rz; nst[0] = sm[0]; nst[5] = sm[2]; nst[4] = it; }
              done = !(sm[2] > a.tol2 * bb) || !(sm[0] > 0.0) || it >= (double)a.max_iters;
              unconv = (sm[2] > a.tol2 * bb) && (sm[0] > 0.0) && it >= (double)a.max_iters;
              // hand-off: once only a few nodes are still iterating, a phase is bound by this kernel's
              // rendezvous chain (~17 us) and not by HBM; those nodes leave here with their CG state
              // in place and k_tsolve_lite (resume mode, ~4 us per rendezvous, all SMs) finishes them
              if (!done && a.handoff_live > 0 && a.n_active - ld_relaxed(a.cnt + 3 * a.n_nodes) <= a.handoff_live)
                handoff = true;
            }
            handoff = __shfl_sync(0xffffffffu, (int)handoff, 0) != 0;
            if (lane == 0) {
              a.cnt[nd] = 0;
              if (handoff) { a.cnt[2 * a.n_nodes + nd] = 1; done = true; }
              else if (done) atomicAdd(a.cnt + 3 * a.n_nodes, 1);
              if (done && !handoff && a.stats) {
                const unsigned long long it = (unsigned long long)(round / 2);
                atomicAdd(a.stats, it);
                atomicAdd(a.stats + 1, it * (unsigned long long)(__ldg(a.node_off + nd + 1) - __ldg(a.node_off + nd)));
                if (unconv) atomicAdd(a.stats + 2, 1ull);      // stopped on max_iters above the tolerance
                atomicMax(a.stats + 3, it);
              }
              st_release(epoch + nd, (round + 1) | (done ? DONE_BIT : 0));
            }
          }
          __syncwarp();
          // segments waiting for their node: has the next phase been published?  (the epoch was
          // polled at the end of the previous pass)
          bool fresh = false;
          const int ep = s0 == 0 ? ep_polled : ld_relaxed(epoch + sg_node[min(sidx, n_seg - 1)]);
          if (have && sg_state[sidx] == 0 && !arrived) fresh = (ep & DONE_BIT) || ep >= sg_round[sidx];
          if (__ballot_sync(0xffffffffu, fresh)) {
            fence_acq_rel();                              // what the publishers stored is visible from here on
            if (fresh) {
              const int r = sg_round[sidx];
              const int kind = (ep & DONE_BIT) ? 3 : ((r & 1) ? 1 : 2);
              if (kind != 3) sg_coef[sidx] = __ldcg(a.nstate + (size_t)sg_node[sidx] * 8 + (kind == 1 ? 3 : 2));
              sg_kind[sidx] = kind;
              sg_state[sidx] = 1;
              st_release_cta_shared(&sg_open[sidx], r);   // the dispatch warp may hand out the phase
            }
            progress = true;
          }
          __syncwarp();
        }
        ep_polled = (lane < n_seg && sg_state[lane] == 0) ? ld_relaxed(epoch + sg_node[lane]) : 0;
        if (!progress) __nanosleep(20);
      }
      if (lane == 0) st_release_cta_shared(&all_done_s, 1);
      return;
    }

    // =========================== dispatch warps (two) ===========================
    // Each claims tiles of open phases (compare-and-swap on the segment's claim counter: claim c is
    // tile c % n of phase c / n), takes the next use number of the ring and issues the tile's bulk
    // copies; two warps, because issuing the copies of one tile costs ~0.4 us.
    const bool lead = wg == C::NGRP * WPT + 1;          // the lead dispatcher also stops the consumers
    int cur = -1;                                        // segment tiles are currently claimed from
    int fenced = -1;                                     // (segment, phase) this warp has proxy-fenced for
    while (true) {
      // ---- claim a tile
      int claim = -1;
      if (cur >= 0 && lane == 0) {
        const int n = sg_n[cur];
        int c = ld_acquire_cta_shared(&sg_claim[cur]);
        while (ld_acquire_cta_shared(&sg_open[cur]) >= c / n) {
          const int old = atomicCAS(&sg_claim[cur], c, c + 1);
          if (old == c) { claim = c; break; }
          c = old;
        }
      }
      claim = __shfl_sync(0xffffffffu, claim, 0);
      if (claim < 0) {
        // pick the open segment with the oldest unclaimed phase
        int best = -1, best_round = 0x7fffffff;
        for (int s0 = 0; s0 < n_seg; s0 += 32) {
          const int sidx = s0 + lane;
          int key = 0x7fffffff;
          if (sidx < n_seg) {
            const int ph = ld_acquire_cta_shared(&sg_claim[sidx]) / sg_n[sidx];
            if (ld_acquire_cta_shared(&sg_open[sidx]) >= ph) key = ph;
          }
          int idx = sidx;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const int k2 = __shfl_xor_sync(0xffffffffu, key, o), i2 = __shfl_xor_sync(0xffffffffu, idx, o);
            if (k2 < key || (k2 == key && i2 < idx)) { key = k2; idx = i2; }
          }
          if (key < best_round) { best_round = key; best = idx; }
        }
        cur = best;
        if (best < 0) {
          if (ld_acquire_cta_shared(&all_done_s)) break;
          __nanosleep(20);
        }
        continue;
      }
      // ---- hand the tile to the consumers
      if (lane == 0) {
        const int n = sg_n[cur];
        const int kk = sg_k0[cur] + claim % n, kd = sg_kind[cur];
        const double coef = sg_coef[cur];
        const int use = atomicAdd(&use_ctr_s, 1);
        const int stage = use % NST;
        uint64_t *fb = &full[use % (C::NGRP * NST)];
        if (use >= NST) mbar_wait(&empty[stage], (unsigned)(use / NST - 1) & 1u);
        d_k[stage] = kk; d_kind[stage] = kd; d_coef[stage] = coef; d_seg[stage] = cur;
        if (kd == 1 || kd == 2) {
          if (fenced != cur * 65536 + claim / n) {
            asm volatile("fence.proxy.async;" ::: "memory");   // bulk copies read what other threads stored
            fenced = cur * 65536 + claim / n;
          }
          const int ct = tile_of(kk);
          unsigned char *sb = dyn + (size_t)stage * C::STAGE_BYTES;
          const double *rc = a.rec + (size_t)ct * C::RL;
          const size_t v0 = (size_t)ct * C::VEC;
          if (kd == 1) {
            const int r0 = m_sell[kk][0], rows = m_sell[kk][WPT] - r0;
            const bool sell_staged = rows <= C::SELL_CAP && rows > 0;
            mbar_expect_tx(fb, (C::HV * C::VEC + 2 * C::VEC + CTILE) * 8 + (sell_staged ? rows * 384 : 0));
            bulk_g2s(sb, a.z + v0 - TS_HALO * C::VEC, C::HV * C::VEC * 8, fb);   // z is padded by the halo
            bulk_g2s(sb + C::OFF_R, rc + C::VEC, (2 * C::VEC + CTILE) * 8, fb);
            if (sell_staged) bulk_g2s(sb + C::OFF_S, a.sell_pack + (size_t)r0 * 384, rows * 384, fb);
          } else {
            mbar_expect_tx(fb, (C::VEC + C::RL) * 8);
            bulk_g2s(sb, a.z + v0, C::VEC * 8, fb);
            bulk_g2s(sb + C::OFF_R, rc, C::RL * 8, fb);
          }
        } else {
          mbar_arrive(fb);                              // no staged data: init / publish use direct loads
        }
      }
      __syncwarp();
    }
    // ---- every segment retired and every tile handed back: the lead tells the consumers to stop
    if (lead && lane == 0) {
      for (int q = 0; q < C::NGRP; ++q) {                // one stop descriptor per consumer group
        const int use = atomicAdd(&use_ctr_s, 1);
        const int stage = use % NST;
        if (use >= NST) mbar_wait(&empty[stage], (unsigned)(use / NST - 1) & 1u);
        d_kind[stage] = -1;
        mbar_arrive(&full[use % (C::NGRP * NST)]);
      }
    }
    return;
  }

  // =============================== consumer warps ===================================
  // the two groups take the ring's uses alternately, so that one group's waits (far gathers,
  // barriers) overlap the other group's arithmetic
  int tick = 0;
  for (int use = grp;; use += C::NGRP) {
    const int stage = use % NST;
    mbar_wait(&full[use % (C::NGRP * NST)], (unsigned)(use / (C::NGRP * NST)) & 1u);
    const int kd = d_kind[stage];
    if (kd < 0) break;
    const int k = d_k[stage], ct = tile_of(k);
    const double coef = d_coef[stage];
    const bool valid = 32 * wi + lane < m_cnt[k];
    const int p = m_start[k] + 32 * wi + lane;                     // own pose index
    const size_t vb = (size_t)(WPT * ct + wi) * (32 * D) + lane;     // slot of (pose, column 0)
    const unsigned char *sb = dyn + (size_t)stage * C::STAGE_BYTES;
    double part[3] = {0.0, 0.0, 0.0};
    double *rc = a.rec + (size_t)ct * C::RL;             // this tile's record {x, p, Ap, diag}
    const int lo = wi * (32 * D) + lane;                 // this pose inside a tile-sized vector
    if (kd == 3) {
      // node finished: publish u into the pose array, t = -u
      if (valid) {
#pragma unroll
        for (int c = 0; c < D; ++c) a.xio[(size_t)p * PB + c] = -__ldcg(rc + lo + 32 * c);
      }
    } else if (kd == 0) {
      // r = b - A x0 ; z = r / diag ; p = Ap = 0 ; partials rz, bb, rr   (direct loads, once per solve)
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
          rc[lo + 32 * c] = x0[c];
          a.z[vb + 32 * c] = zv;
          rc[C::VEC + lo + 32 * c] = 0.0;
          rc[2 * C::VEC + lo + 32 * c] = 0.0;
          part[0] += rv * zv; part[1] += b[c] * b[c]; part[2] += rv * rv;
        }
      }
    } else if (kd == 1) {
      // phase A: w = A z ; Ap = w + beta Ap ; p = z + beta p ; partial p.Ap
      const double beta = coef;
      const double *zh = reinterpret_cast<const double *>(sb);              // z of the tiles ct-HALO .. ct+HALO
      const double *sp = reinterpret_cast<const double *>(sb + C::OFF_R) + lo;
      const double *sap = sp + C::VEC;
      const double dg = reinterpret_cast<const double *>(sb + C::OFF_R)[2 * C::VEC + 32 * wi + lane];
      const int r0 = m_sell[k][0], rows = m_sell[k][WPT] - r0;
      const int s0 = m_sell[k][wi] - r0, s1 = m_sell[k][wi + 1] - r0;
      const bool sell_staged = rows <= C::SELL_CAP;
      const unsigned char *pack = sell_staged ? sb + C::OFF_S : a.sell_pack + (size_t)r0 * 384;
      const double *sval = reinterpret_cast<const double *>(pack);
      const int *scol = reinterpret_cast<const int *>(pack + (size_t)rows * 256);
      const int halo_lo = (ct - TS_HALO) * C::VEC;
      double zo[D], acc[D];
#pragma unroll
      for (int c = 0; c < D; ++c) { zo[c] = zh[TS_HALO * C::VEC + lo + 32 * c]; acc[c] = dg * zo[c]; }
      // the slice loop is warp-uniform (padded entries have value 0); entries are taken eight at
      // a time so that the gathers of far neighbours (outside the staged tiles) overlap
      constexpr int EB = 8;
      for (int s = s0; s < s1; s += EB) {
        int slot[EB];
        double av[EB], zq[EB][D];
#pragma unroll
        for (int q = 0; q < EB; ++q) {
          const bool in = s + q < s1;
          slot[q] = in ? scol[(s + q) * 32 + lane] : (int)vb;
          av[q] = in ? sval[(s + q) * 32 + lane] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < EB; ++q) {
          const unsigned rel = (unsigned)(slot[q] - halo_lo);
          if (rel >= (unsigned)(C::HV * C::VEC)) {
#pragma unroll
            for (int c = 0; c < D; ++c) zq[q][c] = __ldcg(a.z + (size_t)slot[q] + 32 * c);
          }
        }
#pragma unroll
        for (int q = 0; q < EB; ++q) {
          const unsigned rel = (unsigned)(slot[q] - halo_lo);
          if (rel < (unsigned)(C::HV * C::VEC)) {
#pragma unroll
            for (int c = 0; c < D; ++c) zq[q][c] = zh[rel + 32 * c];
          }
        }
#pragma unroll
        for (int q = 0; q < EB; ++q)
#pragma unroll
          for (int c = 0; c < D; ++c) acc[c] = fma(av[q], zq[q][c], acc[c]);
      }
      if (valid) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const double apv = acc[c] + beta * sap[32 * c];
          const double pv = zo[c] + beta * sp[32 * c];
          rc[2 * C::VEC + lo + 32 * c] = apv;
          rc[C::VEC + lo + 32 * c] = pv;
          part[0] += pv * apv;
        }
      }
    } else {
      // phase B: x += alpha p ; z -= alpha Ap / diag (r = diag z) ; partials rz, rr
      const double alpha = coef;
      const double *sz = reinterpret_cast<const double *>(sb) + lo;
      const double *sx = reinterpret_cast<const double *>(sb + C::OFF_R) + lo;
      const double *sp = sx + C::VEC, *sap = sx + 2 * C::VEC;
      const double dg = reinterpret_cast<const double *>(sb + C::OFF_R)[3 * C::VEC + 32 * wi + lane];
      if (valid) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const double xv = sx[32 * c] + alpha * sp[32 * c];
          const double rv = dg * sz[32 * c] - alpha * sap[32 * c];
          const double zv = rv / dg;
          rc[lo + 32 * c] = xv;
          a.z[vb + 32 * c] = zv;
          part[0] += rv * zv; part[2] += rv * rv;
        }
      }
    }
    const int rb = tick & 1;
    if (kd != 3) {
      const double p0 = warp_sum(part[0]), p1 = warp_sum(part[1]), p2 = warp_sum(part[2]);
      if (lane == 0) { red[grp][rb][wi][0] = p0; red[grp][rb][wi][1] = p1; red[grp][rb][wi][2] = p2; }
    }
    consumer_barrier(grp);                               // the group is done with the stage; red[grp][rb] complete
    if (wi == 0) {
      if (kd != 3 && lane < 3) {
        // the tile's partial, summed over the warps in a fixed order
        double sacc = 0.0;
#pragma unroll
        for (int q = 0; q < WPT; ++q) sacc += red[grp][rb][q][lane];
        a.partials[(size_t)ct * 4 + lane] = sacc;
      }
      __syncwarp();
      if (lane == 0) {
        const int seg = d_seg[stage];
        mbar_arrive(&empty[stage]);                      // stage buffer free again
        // hand the tile back: the scheduler arrives on the node once the whole segment is back
        asm volatile("fence.acq_rel.cta;" ::: "memory");
        atomicAdd(&sg_done[seg], 1);
      }
    }
    ++tick;
  }
}


// =============================================================================
// K2b, small-shard variant ("lite").  Same algorithm, data layout, per-pose arithmetic and
// fixed-order reductions as k_tsolve (the two kernels give bit-identical solutions), but built
// for the regime where a GPU holds so few poses per SM that an iteration is bound by the
// rendezvous, not by HBM (<= ~250 k poses per GPU: 4 and 8 GPUs on the 1 M-pose graph):
//   * every CTA owns a contiguous run of `chunk` CTA tiles; a group of four warps works on one
//     tile with plain loads, no copy ring, no service warps; the tile records, the CTA's own z
//     tiles and its ELLPACK rows stay in shared memory for the whole solve when they fit (host
//     plan in TSolveArgs::lite_*), the rest of the working set is L2 resident;
//   * one rendezvous per phase and node: arrive on the node's monotonic counter, spin until all
//     CTAs that hold tiles of the node have arrived, then EVERY such CTA sums the node's
//     per-tile partials itself (same fixed order => same scalars everywhere).  The chain
//     "last arriver reduces -> publishes epoch -> pollers read coefficients -> dispatch" of the
//     ring kernel (~10 us per phase) shrinks to atomic + poll + one read of the partials.
// Partials are double buffered by round parity: a CTA that is already in round r+1 must not
// overwrite what a slower CTA still sums for round r.
// =============================================================================
constexpr int TSL_MAXSEG = TSL_MAXT;   // node segments (runs of tiles of one active node) per CTA
constexpr int TSL_NG = 7;              // four-warp groups per CTA, one tile each at a time (896 threads, 72 registers)

template <int D, int NG>
__global__ void __launch_bounds__(NG * CTILE, 1) k_tsolve_lite(TSolveArgs a) {
  typedef TSCfg<D> C;
  constexpr int PB = (D + 1) * D;
  constexpr int WPT = C::WPT;
  const int lane = threadIdx.x & 31, wg = threadIdx.x >> 5;
  const int wi = wg % WPT, grp = wg / WPT;
  // Tile dealing.  Fresh solve: CTA b owns the contiguous CTA tiles [b chunk, (b+1) chunk).  Resume
  // (a.resume, after a hand-off from k_tsolve): only the nodes flagged in a.cnt[2 n_nodes + node]
  // continue; their tiles are numbered consecutively and dealt evenly over the whole grid, and the
  // CG state {x, p, Ap, z; rz, |b|^2, beta, iterations} is picked up where k_tsolve left it.
  const bool resume = a.resume != 0;
  __shared__ int m_tile[TSL_MAXT], m_start[TSL_MAXT], m_cnt[TSL_MAXT], m_seg[TSL_MAXT], m_sell[TSL_MAXT][TS_WPT + 1];
  __shared__ int sg_node[TSL_MAXSEG], sg_target[TSL_MAXSEG], sg_state[TSL_MAXSEG], sg_owner[TSL_MAXSEG];
  __shared__ double sg_coef[TSL_MAXSEG], sg_rz[TSL_MAXSEG], sg_bb[TSL_MAXSEG], sg_it[TSL_MAXSEG];
  __shared__ int n_seg_s, nb_s;
  __shared__ double red[NG][2][TS_WPT][3];
  const int k0 = blockIdx.x * a.chunk;                      // first tile (fresh solve only)
  if (threadIdx.x == 0) {
    int ns = 0, nbl = 0;
    if (!resume) {
      const int tpc = a.chunk;
      nbl = max(0, min(tpc, a.n_ct - k0));
      // segments: runs of this CTA's tiles that belong to one active node
      int prev = -1;
      for (int k = 0; k < nbl; ++k) {
        m_tile[k] = k0 + k;
        const int nd = __ldg(a.ct_node + k0 + k);
        const bool on = !(a.active && !__ldg(a.active + nd));
        if (!on) { m_seg[k] = -1; continue; }
        if (nd != prev) {
          const int cb = __ldg(a.node_ctb + nd), ce = __ldg(a.node_cte + nd);
          sg_node[ns] = nd;
          sg_target[ns] = (ce - 1) / tpc - cb / tpc + 1;      // CTAs that hold tiles of the node
          sg_owner[ns] = (cb >= k0 && cb < k0 + nbl) ? 1 : 0;  // this CTA reports the node's statistics
          sg_state[ns] = 0; sg_coef[ns] = 0.0; sg_rz[ns] = 0.0; sg_bb[ns] = 0.0; sg_it[ns] = 0.0;
          prev = nd; ++ns;
        }
        m_seg[k] = ns - 1;
      }
    } else {
      const int *flag = a.cnt + 2 * a.n_nodes;
      int T = 0;
      for (int nd = 0; nd < a.n_nodes; ++nd)
        if (__ldcg(flag + nd)) T += __ldg(a.node_cte + nd) - __ldg(a.node_ctb + nd);
      const int tpc = max(1, (T + (int)gridDim.x - 1) / (int)gridDim.x);
      if (tpc > TSL_MAXT) __trap();                          // host-checked bound
      const int c0 = min((int)blockIdx.x * tpc, T), c1 = min(c0 + tpc, T);
      nbl = c1 - c0;
      int off = 0;
      for (int nd = 0; nd < a.n_nodes && off < c1; ++nd) {
        if (!__ldcg(flag + nd)) continue;
        const int cb = __ldg(a.node_ctb + nd), n = __ldg(a.node_cte + nd) - cb;
        const int lo = max(c0, off), hi = min(c1, off + n);
        if (lo < hi) {
          const double *nst = a.nstate + (size_t)nd * 8;
          sg_node[ns] = nd;
          sg_target[ns] = (off + n - 1) / tpc - off / tpc + 1;
          sg_owner[ns] = (off >= c0 && off < c1) ? 1 : 0;
          sg_state[ns] = 0;
          sg_coef[ns] = __ldcg(nst + 3);                     // beta of the phase A that comes next
          sg_rz[ns] = __ldcg(nst + 0); sg_bb[ns] = __ldcg(nst + 1); sg_it[ns] = __ldcg(nst + 4);
          for (int c = lo; c < hi; ++c) { m_tile[c - c0] = cb + (c - off); m_seg[c - c0] = ns; }
          ++ns;
        }
        off += n;
      }
    }
    n_seg_s = ns; nb_s = nbl;
  }
  __syncthreads();
  const int nb = nb_s;
  for (int k = threadIdx.x; k < nb; k += blockDim.x) {
    const int ct = m_tile[k];
    m_start[k] = __ldg(a.ct_start + ct); m_cnt[k] = __ldg(a.ct_cnt + ct);
  }
  for (int q = threadIdx.x; q < nb * (WPT + 1); q += blockDim.x) {
    const int k = q / (WPT + 1), r = q % (WPT + 1);
    m_sell[k][r] = __ldg(a.sell_ptr + WPT * m_tile[k] + r);
  }
  // the ELLPACK rows of this CTA's tiles are one contiguous run of sell_pack: staged in shared
  // memory once per launch when they fit (a.lite_stage_bytes > 0), so that phase A is one
  // shared-memory read + one L2 gather deep instead of two dependent L2 round trips
  extern __shared__ __align__(128) unsigned char dyn[];
  const int row_first = nb > 0 ? __ldg(a.sell_ptr + WPT * k0) : 0;
  const bool staged = a.lite_stage_bytes > 0;
  if (staged && nb > 0) {
    const int row_end = __ldg(a.sell_ptr + WPT * (k0 + nb));
    const int n16 = (row_end - row_first) * 24;                       // 384-byte rows as 16-byte words
    const int4 *src = reinterpret_cast<const int4 *>(a.sell_pack + (size_t)row_first * 384);
    int4 *dst = reinterpret_cast<int4 *>(dyn);
    for (int q = threadIdx.x; q < n16; q += blockDim.x) dst[q] = __ldg(src + q);
  }
  // tile-private solver state {x, p, Ap, diag} and a copy of the CTA's own z tiles live in shared
  // memory when they fit (a.lite_vec_off / a.lite_z_off >= 0): per iteration only the z values
  // and the gathers of neighbours owned by other CTAs touch L2
  const bool vres = a.lite_vec_off >= 0, zres = a.lite_z_off >= 0;
  double *vec_s = reinterpret_cast<double *>(dyn + (vres ? a.lite_vec_off : 0));
  double *z_s = reinterpret_cast<double *>(dyn + (zres ? a.lite_z_off : 0));
  const int zbase = WPT * k0 * (32 * D);                   // slot of the CTA's first pose
  const unsigned zlim = zres ? (unsigned)(nb * C::VEC) : 0u;
  if (vres) {
    for (int q = threadIdx.x; q < nb * CTILE; q += blockDim.x) {
      const int k = q / CTILE, r = q % CTILE;
      vec_s[(size_t)k * C::RL + 3 * C::VEC + r] = __ldg(a.rec + (size_t)(k0 + k) * C::RL + 3 * C::VEC + r);
    }
  }
  __syncthreads();
  const int n_seg = n_seg_s;
  if (n_seg == 0) return;
  const int round0 = resume ? 1 : 0;                        // a resumed solve starts with a phase A
  for (int round = round0;; ++round) {
    double *pbuf = a.partials + (size_t)(round & 1) * a.n_ct * 4;
    // ---------------- the phase, one tile per group at a time ----------------
    for (int k = grp; k < nb; k += NG) {
      const int seg = m_seg[k];
      if (seg < 0) continue;
      const int st = sg_state[seg];
      if (st == 2) continue;
      const int kd = st == 1 ? 3 : (round == 0 ? 0 : ((round & 1) ? 1 : 2));
      const double coef = sg_coef[seg];
      const int ct = m_tile[k];
      const bool valid = 32 * wi + lane < m_cnt[k];
      const int p = m_start[k] + 32 * wi + lane;
      const size_t vb = (size_t)(WPT * ct + wi) * (32 * D) + lane;
      double part[3] = {0.0, 0.0, 0.0};
      double *rc = vres ? vec_s + (size_t)k * C::RL : a.rec + (size_t)ct * C::RL;
      const int lo = wi * (32 * D) + lane;
      double *zown = z_s + (size_t)k * C::VEC + lo;        // this pose in the CTA's z copy (zres)
      if (kd == 3) {
        if (valid) {
#pragma unroll
          for (int c = 0; c < D; ++c) a.xio[(size_t)p * PB + c] = -rc[lo + 32 * c];
        }
      } else if (kd == 0) {
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
            rc[lo + 32 * c] = x0[c];
            a.z[vb + 32 * c] = zv;
            if (zres) zown[32 * c] = zv;
            rc[C::VEC + lo + 32 * c] = 0.0;
            rc[2 * C::VEC + lo + 32 * c] = 0.0;
            part[0] += rv * zv; part[1] += b[c] * b[c]; part[2] += rv * rv;
          }
        }
      } else if (kd == 1) {
        const double beta = coef;
        const double *sp = rc + C::VEC + lo;
        const double *sap = sp + C::VEC;
        const int r0 = m_sell[k][0], rows = m_sell[k][WPT] - r0;
        const int s0 = m_sell[k][wi] - r0, s1 = m_sell[k][wi + 1] - r0;
        const unsigned char *pack = staged ? dyn + (size_t)(r0 - row_first) * 384 : a.sell_pack + (size_t)r0 * 384;
        const double *sval = reinterpret_cast<const double *>(pack);
        const int *scol = reinterpret_cast<const int *>(pack + (size_t)rows * 256);
        double zo[D], acc[D];
#pragma unroll
        for (int c = 0; c < D; ++c) zo[c] = zres ? zown[32 * c] : __ldcg(a.z + vb + 32 * c);
        const double dg = rc[3 * C::VEC + 32 * wi + lane];
        // gathers four entries at a time (twelve loads in flight per thread; the register budget
        // of a 1024-thread CTA); the accumulation order is that of k_tsolve
        constexpr int EB = 4;
        double zq[EB][D];
        auto gather = [&](int s) {
#pragma unroll
          for (int q = 0; q < EB; ++q) {
            const int slot = s + q < s1 ? scol[(s + q) * 32 + lane] : (int)vb;
            const unsigned rel = (unsigned)(slot - zbase);
            if (rel < zlim) {                               // neighbour owned by this CTA
#pragma unroll
              for (int c = 0; c < D; ++c) zq[q][c] = z_s[rel + 32 * c];
            } else {
#pragma unroll
              for (int c = 0; c < D; ++c) zq[q][c] = __ldcg(a.z + (size_t)slot + 32 * c);
            }
          }
        };
        if (s0 < s1) gather(s0);
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = dg * zo[c];
        for (int s = s0; s < s1; s += EB) {
#pragma unroll
          for (int q = 0; q < EB; ++q) {
            const double av = s + q < s1 ? sval[(s + q) * 32 + lane] : 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] = fma(av, zq[q][c], acc[c]);
          }
          if (s + EB < s1) gather(s + EB);
        }
        if (valid) {
#pragma unroll
          for (int c = 0; c < D; ++c) {
            const double apv = acc[c] + beta * sap[32 * c];
            const double pv = zo[c] + beta * sp[32 * c];
            rc[2 * C::VEC + lo + 32 * c] = apv;
            rc[C::VEC + lo + 32 * c] = pv;
            part[0] += pv * apv;
          }
        }
      } else {
        const double alpha = coef;
        const double *sx = rc + lo;
        const double *sp = sx + C::VEC, *sap = sx + 2 * C::VEC;
        const double dg = rc[3 * C::VEC + 32 * wi + lane];
        if (valid) {
          double xo[D], po[D], apo[D], zo[D];              // all loads first: one round trip
#pragma unroll
          for (int c = 0; c < D; ++c) {
            xo[c] = sx[32 * c]; po[c] = sp[32 * c]; apo[c] = sap[32 * c];
            zo[c] = zres ? zown[32 * c] : __ldcg(a.z + vb + 32 * c);
          }
#pragma unroll
          for (int c = 0; c < D; ++c) {
            const double xv = xo[c] + alpha * po[c];
            const double rv = dg * zo[c] - alpha * apo[c];
            const double zv = rv / dg;
            rc[lo + 32 * c] = xv;
            a.z[vb + 32 * c] = zv;
            if (zres) zown[32 * c] = zv;
            part[0] += rv * zv; part[2] += rv * rv;
          }
        }
      }
      if (kd != 3) {
        const int rb = (k / NG) & 1;                       // red is double buffered per group
        const double p0 = warp_sum(part[0]), p1 = warp_sum(part[1]), p2 = warp_sum(part[2]);
        if (lane == 0) { red[grp][rb][wi][0] = p0; red[grp][rb][wi][1] = p1; red[grp][rb][wi][2] = p2; }
        consumer_barrier(grp);                             // red[grp][rb] complete
        if (wi == 0 && lane < 3) {
          double sacc = 0.0;
#pragma unroll
          for (int q = 0; q < WPT; ++q) sacc += red[grp][rb][q][lane];
          pbuf[(size_t)ct * 4 + lane] = sacc;
        }
      }
    }
    __syncthreads();
    // ---------------- rendezvous + node scalars, one warp per segment ----------------
    // all arrivals first, then the waits (a warp may serve several segments)
    for (int s = wg; s < n_seg; s += NG * WPT) {
      // release: the stores of the whole CTA (ordered before this by the barrier above)
      if (sg_state[s] == 0 && lane == 0) red_add_release(a.cnt + sg_node[s], 1);
    }
    __syncwarp();
    for (int s = wg; s < n_seg; s += NG * WPT) {
      const int st = sg_state[s];
      if (st == 1) { if (lane == 0) sg_state[s] = 2; continue; }
      if (st != 0) continue;
      const int nd = sg_node[s];
      const int target = sg_target[s] * (round - round0 + 1);
      {
        const long long t0 = clock64();
        unsigned spins = 0;
        while (ld_relaxed(a.cnt + nd) < target) {
          if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000ll) __trap();   // never hang the GPU
        }
        fence_acq_rel();                                   // acquire: what the other CTAs stored before arriving
      }
      const int cb = __ldg(a.node_ctb + nd), ce = __ldg(a.node_cte + nd);
      bool done = false, unconv = false;
      if (round == 0) {
        double sm[3];
        node_sum<3>(pbuf, cb, ce, lane, sm);
        if (lane == 0) { sg_rz[s] = sm[0]; sg_bb[s] = sm[1]; sg_it[s] = 0.0; sg_coef[s] = 0.0; }
        done = !(sm[0] > 0.0) || !(sm[2] > a.tol2 * sm[1]);
      } else if (round & 1) {
        double sm[1];
        node_sum<1>(pbuf, cb, ce, lane, sm);
        if (sm[0] > 0.0) { if (lane == 0) sg_coef[s] = sg_rz[s] / sm[0]; }      // alpha for phase B
        else done = true;
      } else {
        double sm[3];
        node_sum<3>(pbuf, cb, ce, lane, sm);
        const double rz = sg_rz[s], bb = sg_bb[s], it = sg_it[s] + 1.0;
        __syncwarp();
        if (lane == 0) { sg_coef[s] = sm[0] / rz; sg_rz[s] = sm[0]; sg_it[s] = it; }   // beta for phase A
        done = !(sm[2] > a.tol2 * bb) || !(sm[0] > 0.0) || it >= (double)a.max_iters;
        unconv = (sm[2] > a.tol2 * bb) && (sm[0] > 0.0) && it >= (double)a.max_iters;
      }
      if (done && lane == 0) {
        sg_state[s] = 1;                                   // publish in the next pass, then retire
        if (sg_owner[s] && a.stats) {
          const unsigned long long it = (unsigned long long)sg_it[s];   // completed CG iterations (= round / 2)
          atomicAdd(a.stats, it);
          atomicAdd(a.stats + 1, it * (unsigned long long)(__ldg(a.node_off + nd + 1) - __ldg(a.node_off + nd)));
          if (unconv) atomicAdd(a.stats + 2, 1ull);        // stopped on max_iters above the tolerance
          atomicMax(a.stats + 3, it);
        }
      }
    }
    __syncthreads();
    int live = 0;
    for (int s = 0; s < n_seg; ++s) live += sg_state[s] != 2;
    if (live == 0) break;
  }
}

template <int D> int launch_tsolve_lite(const TSolveArgs &a, int grid, cudaStream_t s) {
  TSolveArgs args = a;
  void *params[] = {&args};
  return (int)cudaLaunchCooperativeKernel((const void *)k_tsolve_lite<D, TSL_NG>, dim3(grid), dim3(TSL_NG * CTILE), params,
                                          (size_t)a.lite_dyn_bytes, s);
}
template int launch_tsolve_lite<2>(const TSolveArgs &, int, cudaStream_t);
template int launch_tsolve_lite<3>(const TSolveArgs &, int, cudaStream_t);

template <int D> int tsolve_lite_max_grid(int device) {
  int per_sm = 0, sms = 0;
  if (cudaFuncSetAttribute(k_tsolve_lite<D, TSL_NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, TSL_STAGE_MAX) != cudaSuccess)
    return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tsolve_lite<D, TSL_NG>, TSL_NG * CTILE, TSL_STAGE_MAX) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  return per_sm * sms;
}
template int tsolve_lite_max_grid<2>(int);
template int tsolve_lite_max_grid<3>(int);

template <int D> int launch_tsolve(const TSolveArgs &a, int grid, cudaStream_t s) {
  TSolveArgs args = a;
  void *params[] = {&args};
  return (int)cudaLaunchCooperativeKernel((const void *)k_tsolve<D>, dim3(grid), dim3(TSCfg<D>::THREADS), params,
                                          TSCfg<D>::DYN_BYTES, s);
}
template int launch_tsolve<2>(const TSolveArgs &, int, cudaStream_t);
template int launch_tsolve<3>(const TSolveArgs &, int, cudaStream_t);

template <int D> int tsolve_max_grid(int device) {
  int per_sm = 0, sms = 0;
  if (cudaFuncSetAttribute(k_tsolve<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, TSCfg<D>::DYN_BYTES) != cudaSuccess)
    return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tsolve<D>, TSCfg<D>::THREADS, TSCfg<D>::DYN_BYTES) != cudaSuccess)
    return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  return per_sm * sms;
}
template int tsolve_max_grid<2>(int);
template int tsolve_max_grid<3>(int);

}  // namespace mmpgo
