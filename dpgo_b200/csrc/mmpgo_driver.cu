// Batched host control flow of the per-node drivers: every robot node of the
// handle advances in lock step through the same kernel launches, with per-node
// scalars (objective values, step lengths, restart decisions) owned by the host.
//
// Restates  DPGOHash::{initialize,update,amm_pgo,mm_pgo,iterate,communicate}
//             C++/DPGO/src/DPGOHash.cpp:20-628, include/DPGO/DPGOHash.h:28-86
//           DPGOStar::{initialize,update_n,amm_pgo_n,mm_pgo_n,pm_pgo_n,iterate}
//             C++/DPGO/src/DPGOStar.cpp:109-711
//           Optimization::Riemannian::TNT      Optimization/Riemannian/TNT.h:242-693
//           Optimization::LinearAlgebra::STPCG Optimization/LinearAlgebra/IterativeSolvers.h:166-426
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <limits>

#include "mmpgo_driver.cuh"

namespace mmpgo {

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      set_error(std::string(#x) + ": " + cudaGetErrorString(e_));                \
      return MMPGO_ERR_CUDA;                                                     \
    }                                                                            \
  } while (0)
#define RC(x)                    \
  do {                           \
    int rc_ = (x);               \
    if (rc_) return rc_;         \
  } while (0)

typedef std::vector<int> Mask;


static bool any(const Mask &m) {
  for (int v : m) if (v) return true;
  return false;
}
static bool all(const Mask &m) {
  for (int v : m) if (!v) return false;
  return true;
}
static Mask mask_and(const Mask &a, const Mask &b) {
  Mask r(a.size());
  for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] && b[i];
  return r;
}

// Tiles view with an uploaded node mask (nullptr when all nodes are active).
static int make_tiles(Handle *h, const Mask &m, Tiles *tl, int *dst = nullptr) {
  tl->n_tiles = h->n_tiles;
  tl->node = h->d_tile_node; tl->start = h->d_tile_start; tl->cnt = h->d_tile_cnt;
  if (all(m)) { tl->active = nullptr; return 0; }
  int *d = dst ? dst : h->d_active;
  CK(cudaMemcpyAsync(d, m.data(), sizeof(int) * m.size(), cudaMemcpyHostToDevice, h->stream));
  tl->active = d;
  return 0;
}

// per-node sums of the tile partials -> host
static int reduce_to_host(Handle *h, const double **out) {
  launch_reduce(h->A, h->d_node_tb, h->d_node_te, h->d_partials, h->d_node_scal, h->stream);
  h->ctr.launches++;
  CK(cudaMemcpyAsync(h->h_pinned, h->d_node_scal, sizeof(double) * h->A * NS, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  *out = h->h_pinned;
  return 0;
}
// Deferred variant: several kernels' per-node sums are reduced into separate slots and read back
// with ONE host synchronisation (the tile partials buffer is reused as soon as its reduce kernel
// has been enqueued).  slot < RED_SLOTS.
static int reduce_async(Handle *h, int slot) {
  double *d = h->d_slot + (size_t)slot * h->A * NS;
  launch_reduce(h->A, h->d_node_tb, h->d_node_te, h->d_partials, d, h->stream);
  h->ctr.launches++;
  CK(cudaMemcpyAsync(h->h_slot + (size_t)slot * h->A * NS, d, sizeof(double) * h->A * NS, cudaMemcpyDeviceToHost,
                     h->stream));
  return 0;
}
static const double *slot_host(Handle *h, int slot) { return h->h_slot + (size_t)slot * h->A * NS; }
static int host_sync(Handle *h) {
  CK(cudaGetLastError());             // a failed launch configuration of any kernel enqueued since the last check
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
static int upload_coef(Handle *h, const std::vector<double> &c) {
  CK(cudaMemcpyAsync(h->d_coef, c.data(), sizeof(double) * c.size(), cudaMemcpyHostToDevice, h->stream));
  return 0;
}

// sum of host scalars over all ranks (no-op for a single handle)
static int allreduce(Handle *h, double *v, int n) {
  if (h->world <= 1) return 0;
  h->allreduces++;
  if (h->nccl_comm) {
    // through the device: pinned staging -> ncclAllReduce on the handle's stream -> back
    double *hp = h->h_pinned + (size_t)h->A * (2 * NS + 8) + 32;
    for (int i = 0; i < n; ++i) hp[i] = v[i];
    CK(cudaMemcpyAsync(h->d_scalar + 4, hp, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    RC(nccl_allreduce(h, h->d_scalar + 4, n));
    CK(cudaMemcpyAsync(hp, h->d_scalar + 4, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n; ++i) v[i] = hp[i];
    return 0;
  }
  if (!h->allreduce_fn) { set_error("no transport: call mmpgo_nccl_init or pass callbacks to mmpgo_set_sharding"); return MMPGO_ERR_STATE; }
  if (h->allreduce_fn(h->cb_user, v, n) != 0) { set_error("allreduce callback failed"); return MMPGO_ERR_ARG; }
  return 0;
}
// boundary chunks of `send` -> peers, theirs -> `recv`, ordered on the handle's stream
static int exchange(Handle *h, const double *send, const std::vector<int64_t> &sc, double *recv, const std::vector<int64_t> &rc) {
  if (h->nccl_comm) return nccl_exchange(h, send, sc.data(), recv, rc.data());
  if (!h->exchange_fn) { set_error("no transport: call mmpgo_nccl_init or pass callbacks to mmpgo_set_sharding"); return MMPGO_ERR_STATE; }
  if (h->exchange_fn(h->cb_user, send, sc.data(), recv, rc.data()) != 0) { set_error("exchange callback failed"); return MMPGO_ERR_ARG; }
  return 0;
}

// Tile dealing of the copy-ring translation solve: every node's CTA tiles are cut into chunks of
// `chunk` consecutive tiles, chunks go round-robin to the persistent CTAs.  A node's rendezvous
// involves the CTAs that hold one of its chunks.  (Measured on the 1M-pose grid: dealing the two
// slowest-converging nodes in chunks of 1-2 tiles, to put more CTAs on the tail of the solve, does
// not pay: 3.37 -> 3.6 ms per cold solve; chunk 8 beats 4 and 16.)
static int plan_ring(Handle *h, int grid, int chunk) {
  if (h->ts_plan_grid == grid && h->ts_plan_chunk == chunk) return 0;
  std::vector<std::vector<int>> lists(grid);
  std::vector<int> parts(h->A, 0);
  int rr = 0;
  for (int a = 0; a < h->A; ++a) {
    const int c = chunk;
    int nch = 0;
    for (int t = h->h_node_ctb[a]; t < h->h_node_cte[a]; t += c, ++rr, ++nch)
      for (int u = t; u < std::min(t + c, h->h_node_cte[a]); ++u) lists[rr % grid].push_back(u);
    parts[a] = std::min(nch, grid);
  }
  std::vector<int> ptr(grid + 1, 0), tiles;
  h->ts_plan_max = 0;
  for (int b = 0; b < grid; ++b) {
    std::sort(lists[b].begin(), lists[b].end());
    tiles.insert(tiles.end(), lists[b].begin(), lists[b].end());
    ptr[b + 1] = (int)tiles.size();
    h->ts_plan_max = std::max(h->ts_plan_max, (int)lists[b].size());
  }
  // the plan buffers may still be read by a solve in flight on the stream: stream-ordered copies
  // from a staging vector that outlives them
  h->ts_plan_stage.assign(ptr.begin(), ptr.end());
  h->ts_plan_stage.insert(h->ts_plan_stage.end(), tiles.begin(), tiles.end());
  h->ts_plan_stage.insert(h->ts_plan_stage.end(), parts.begin(), parts.end());
  const int *st = h->ts_plan_stage.data();
  CK(cudaMemcpyAsync(h->d_cta_ptr, st, sizeof(int) * ptr.size(), cudaMemcpyHostToDevice, h->stream));
  if (!tiles.empty())
    CK(cudaMemcpyAsync(h->d_cta_tiles, st + ptr.size(), sizeof(int) * tiles.size(), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_node_parts, st + ptr.size() + tiles.size(), sizeof(int) * parts.size(), cudaMemcpyHostToDevice,
                     h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->ts_plan_grid = grid; h->ts_plan_chunk = chunk;
  return 0;
}

template <int D> struct Drv {
  static constexpr int PB = (D + 1) * D;

  // boundary poses of `x` (an NP-sized pose array) -> peers; their boundary poses -> our halo rows
  static int halo_exchange(Handle *h, double *x) {
    if (h->world <= 1) return 0;
    launch_gather_poses<D>(h->n_send, h->d_send_idx, x, h->d_send, h->stream);
    h->ctr.launches++;
    // no host synchronisation: the transport orders itself on the handle's stream (mmpgo.h)
    h->halo_exchanges++;
    return exchange(h, h->d_send, h->send_dbl, x + (size_t)h->NO * PB, h->recv_dbl);
  }

  // the same for two pose arrays in ONE all-to-all (per-peer chunks [a | b])
  static int halo_exchange2(Handle *h, double *xa, double *xb) {
    if (h->world <= 1) return 0;
    launch_copy_poses<D>(h->n_send, h->d_send_idx, h->d_send2_a, xa, h->d_send2, h->stream);
    launch_copy_poses<D>(h->n_send, h->d_send_idx, h->d_send2_b, xb, h->d_send2, h->stream);
    h->ctr.launches += 2;
    h->halo_exchanges++;
    RC(exchange(h, h->d_send2, h->send_dbl2, h->d_recv2, h->recv_dbl2));
    launch_copy_poses<D>(h->NH, h->d_recv2_a, h->d_halo_row, h->d_recv2, xa, h->stream);
    launch_copy_poses<D>(h->NH, h->d_recv2_b, h->d_halo_row, h->d_recv2, xb, h->stream);
    h->ctr.launches += 2;
    return 0;
  }

  static GPassArgs gargs(Handle *h, const double *diag = nullptr) {
    GPassArgs a;
    std::memset(&a, 0, sizeof(a));
    a.rowptr = h->d_rowptr; a.col = h->d_col; a.blk = h->d_blk; a.blk0 = h->d_blk0;
    a.diag = diag ? diag : h->d_gdiag; a.partials = h->d_partials;
    a.out_perm = h->use_direct ? h->d_mf_perm : nullptr;      // G_RHS_T feeds the sparse direct solve in its elimination order
    return a;
  }

  // tr(x^T (g + 1/2 G x)) per node  (evaluate_G without the offset f)
  static int eval_G(Handle *h, const double *x, const double *g, const Mask &m, std::vector<double> &val) {
    Tiles tl; RC(make_tiles(h, m, &tl));
    GPassArgs a = gargs(h);
    a.x = x; a.g = g;
    launch_gpass<D>(G_EVAL, tl, a, h->stream);
    h->ctr.launches++; h->ctr.intra_passes++;
    const double *s; RC(reduce_to_host(h, &s));
    val.assign(h->A, 0.0);
    for (int n = 0; n < h->A; ++n) if (m[n]) val[n] = s[n * NS];
    return 0;
  }

  // ---- K2b: G00 u = rhs, then xio.t = -u      (recover_translations, DPGOProblem.h:275-294)
  static int solve_t(Handle *h, double *xio, const Mask &m, bool warm) {
    const Mask md = mask_and(m, h->dense_mask), mp = mask_and(m, h->pcg_mask);
    h->ctr.solve_calls++;
    if (any(md)) {
      CK(cudaMemcpyAsync(h->d_active2, md.data(), sizeof(int) * md.size(), cudaMemcpyHostToDevice, h->stream));
      launch_dense_solve<D>(h->A, h->d_node_off, h->d_dense_off, h->d_active2, h->d_ginv, h->rhs_t, xio,
                            h->max_dense_n0, h->stream);
      h->ctr.launches++;
    }
    if (any(mp) && h->use_direct) {
      // sparse Cholesky factor of G00 (host, at upload) applied by one persistent launch of the
      // level-scheduled supernodal sweeps (mmpgo_mfsolve.cu): exact like the reference's L_.solve
      MfSolveArgs ma;
      ma.f = h->mf;
      if (all(mp)) ma.active = nullptr;
      else {
        CK(cudaMemcpyAsync(h->d_active2, mp.data(), sizeof(int) * mp.size(), cudaMemcpyHostToDevice, h->stream));
        ma.active = h->d_active2;
      }
      ma.rhs = h->rhs_t; ma.out = xio; ma.out_stride = PB; ma.sign = -1.0; ma.dry = h->mf_dry ? 1 : 0; ma.level_sync = h->mf_force_dep ? 0 : (h->mf_level_sync || h->mf_level_sync_auto) ? 1 : 0;
      CK((cudaError_t)launch_mf_solve<D>(ma, h->mf_grid, h->stream));
      h->ctr.launches++;
      h->ctr.reserved[2]++;                                   // solves served by the sparse direct kernel
      return 0;
    }
    if (any(mp)) {
      // one persistent launch: per-node Jacobi-PCG to `translation_solve_tol` (mmpgo_tsolve.cu)
      TSolveArgs ta;
      std::memset(&ta, 0, sizeof(ta));
      ta.rowptr = h->d_rowptr; ta.col = h->d_col; ta.a00 = h->d_a00; ta.d00 = h->d_d00;
      ta.sell_ptr = h->d_sell_ptr; ta.sell_pack = h->d_sell_pack;
      ta.ct_node = h->d_ct_node; ta.ct_start = h->d_ct_start; ta.ct_cnt = h->d_ct_cnt;
      ta.n_ct = h->n_ctiles; ta.chunk = h->ts_chunk; ta.node_ctb = h->d_node_ctb; ta.node_cte = h->d_node_cte;
      if (all(mp)) ta.active = nullptr;
      else {
        CK(cudaMemcpyAsync(h->d_active2, mp.data(), sizeof(int) * mp.size(), cudaMemcpyHostToDevice, h->stream));
        ta.active = h->d_active2;
      }
      ta.rhs = h->rhs_t; ta.xio = xio; ta.warm = warm ? 1 : 0;
      ta.rec = h->ts_rec; ta.z = h->ts_z;
      ta.partials = h->ts_partials; ta.nstate = h->ts_nstate;
      ta.cnt = h->d_ts_sync; ta.n_nodes = h->A; ta.n_active = 0;
      for (int v : mp) ta.n_active += v ? 1 : 0;
      ta.stats = h->d_ts_stats; ta.node_off = h->d_node_off;
      ta.tol2 = h->opt.translation_solve_tol * h->opt.translation_solve_tol;
      ta.max_iters = h->opt.translation_solve_max_iters;
      CK(cudaMemsetAsync(h->d_ts_sync, 0, sizeof(int) * (3 * h->A + 8), h->stream));
      // small shards (few CTA tiles per SM) are rendezvous bound: k_tsolve_lite; large ones are
      // HBM bound: the copy-ring kernel.  MMPGO_TS_KERNEL=ring|lite (read at set_graph) overrides (experiments).
      const int lgrid = std::max(1, std::min(h->tsl_max_grid, h->n_ctiles));
      const int tpc = (h->n_ctiles + lgrid - 1) / lgrid;
      bool lite = tpc <= std::min(h->ts_lite_max_tiles, TSL_MAXT);
      if (h->ts_force_kernel == 1) lite = false;
      if (h->ts_force_kernel == 2) lite = tpc <= TSL_MAXT;
      if (lite) {
        ta.chunk = tpc;
        if (h->tsl_plan_tpc != tpc) {
          // shared-memory plan, in order of benefit per byte: the CTA's own z tiles, the tile
          // records {x, p, Ap, diag}, the ELLPACK rows (of the fullest CTA)
          const int64_t zb = (int64_t)tpc * CTILE * D * 8, vb = (int64_t)tpc * (3 * CTILE * D + CTILE) * 8;
          int64_t eb = 0;
          for (int c0 = 0; c0 < h->n_ctiles; c0 += tpc) {
            const int c1 = std::min(c0 + tpc, h->n_ctiles);
            eb = std::max<int64_t>(eb, (int64_t)(h->h_sell_ptr[(size_t)TS_WPT * c1] - h->h_sell_ptr[(size_t)TS_WPT * c0]) * 384);
          }
          int64_t used = 0;
          h->tsl_stage_bytes = 0; h->tsl_vec_off = -1; h->tsl_z_off = -1;
          const bool want_z = used + zb <= TSL_STAGE_MAX; if (want_z) used += zb;
          const bool want_v = used + vb <= TSL_STAGE_MAX; if (want_v) used += vb;
          const bool want_e = used + eb <= TSL_STAGE_MAX; if (want_e) used += eb;
          int64_t off = 0;
          if (want_e) { h->tsl_stage_bytes = (int)eb; off += (eb + 127) / 128 * 128; }
          if (want_v) { h->tsl_vec_off = (int)off; off += vb; }
          if (want_z) { h->tsl_z_off = (int)off; off += zb; }
          h->tsl_dyn_bytes = (int)off;
          h->tsl_plan_tpc = tpc;
        }
        const bool nores = h->ts_nores;                          // experiments: everything from L2
        ta.lite_stage_bytes = nores ? 0 : h->tsl_stage_bytes;
        ta.lite_vec_off = nores ? -1 : h->tsl_vec_off;
        ta.lite_z_off = nores ? -1 : h->tsl_z_off;
        ta.lite_dyn_bytes = nores ? 0 : h->tsl_dyn_bytes;
        CK((cudaError_t)launch_tsolve_lite<D>(ta, (h->n_ctiles + tpc - 1) / tpc, h->stream));
        h->ctr.launches++;
        h->ctr.reserved[1]++;                                 // solves served by k_tsolve_lite
        CK(cudaMemcpyAsync(h->h_ts_stats, h->d_ts_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        return 0;
      }
      int grid = std::max(1, std::min(std::min(h->ts_max_grid, 1024), h->n_ctiles));
      if (h->ts_grid_override > 0) grid = std::min(grid, h->ts_grid_override);
      RC(plan_ring(h, grid, ta.chunk));
      if (h->ts_plan_max > TS_MAXCT) {
        set_error("translation solve: too many poses per GPU for the persistent kernel (shard over more GPUs)");
        return MMPGO_ERR_UNSUPPORTED;
      }
      ta.cta_ptr = h->d_cta_ptr; ta.cta_tiles = h->d_cta_tiles; ta.node_parts = h->d_node_parts;
      // the tail of the solve (a few slowly converging nodes still iterating) is rendezvous bound:
      // those nodes are handed to k_tsolve_lite with their CG state (same arithmetic, same result)
      int max_nt = 1;
      for (int n = 0; n < h->A; ++n) max_nt = std::max(max_nt, h->h_node_cte[n] - h->h_node_ctb[n]);
      int hl = std::max(2, ta.n_active / 16);
      hl = std::min(hl, (int)((int64_t)lgrid * TSL_MAXT / max_nt));
      if (h->ts_handoff >= 0) hl = std::min(h->ts_handoff, (int)((int64_t)lgrid * TSL_MAXT / max_nt));
      ta.handoff_live = std::max(hl, 0);
      CK((cudaError_t)launch_tsolve<D>(ta, grid, h->stream));
      h->ctr.launches++;
      if (ta.handoff_live > 0) {
        ta.resume = 1; ta.chunk = 1;
        ta.lite_stage_bytes = 0; ta.lite_vec_off = -1; ta.lite_z_off = -1; ta.lite_dyn_bytes = 0;
        CK((cudaError_t)launch_tsolve_lite<D>(ta, lgrid, h->stream));
        h->ctr.launches++;
      }
      CK(cudaMemcpyAsync(h->h_ts_stats, h->d_ts_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    }
    return 0;
  }

  // t = -G00^{-1} (g_t + G01 Y)   for the rotation rows currently in xio
  static int recover_t(Handle *h, double *xio, const double *g, const Mask &m, bool warm = true) {
    Tiles tl; RC(make_tiles(h, m, &tl));
    GPassArgs a = gargs(h);
    a.x = xio; a.g = g; a.out = h->rhs_t;
    launch_gpass<D>(G_RHS_T, tl, a, h->stream);
    h->ctr.launches++; h->ctr.intra_passes++;
    return solve_t(h, xio, m, warm);
  }

  // K3 wrapper
  static int proximal(Handle *h, const double *xa, const double *xb, const double *dfa, const double *dfb,
                      const double *ga, const double *gb, double *gex, double *xout, const double *xref,
                      const Mask &m, std::vector<double> *dist2) {
    Tiles tl; RC(make_tiles(h, m, &tl));
    ProxArgs a; std::memset(&a, 0, sizeof(a));
    a.xa = xa; a.xb = xb; a.dfa = dfa; a.dfb = dfb; a.ga = ga; a.gb = gb; a.gamma = h->d_gamma;
    a.tnv = h->d_tnv; a.xref = xref; a.xout = xout; a.gex = gex; a.partials = h->d_partials;
    launch_prox<D>(tl, a, h->stream);
    h->ctr.launches++; h->ctr.prox_passes++;
    if (dist2) {
      const double *s; RC(reduce_to_host(h, &s));
      dist2->assign(h->A, 0.0);
      for (int n = 0; n < h->A; ++n) (*dist2)[n] = s[n * NS];
    }
    return 0;
  }

  // RegularizedCholesky: pre_buf (rotation rows) = (G11 + reg)^{-1} src for the nodes of m -- reg_Chol_precon_.solve
  // (DPGOProblem.cpp:592-594) as a gather into elimination order and one persistent launch of the supernodal sweeps
  static int precon_solve(Handle *h, const double *src, const Mask &m) {
    Tiles tl; RC(make_tiles(h, m, &tl));
    launch_gather_rot<D>(tl, src, h->d_mf11_perm, h->rhs11, h->stream);
    MfSolveArgs ma;
    ma.f = h->mf11;
    ma.active = tl.active;                 // the mask make_tiles uploaded (stream-ordered before the launch)
    ma.rhs = h->rhs11; ma.out = h->pre_buf; ma.out_stride = D; ma.sign = 1.0; ma.dry = 0;
    ma.level_sync = h->mf11_level_sync ? 1 : 0;
    CK((cudaError_t)launch_mf_solve<D>(ma, h->mf11_grid, h->stream));
    h->ctr.launches += 2;
    h->ctr.reserved[5]++;                  // preconditioner solves
    return 0;
  }
  static int vec(Handle *h, int op, const Mask &m, const double *a_, const double *b_, double *o1, double *o2,
                 double *o3, double *o4, const double *y, double *o5 = nullptr) {
    const bool regchol = h->opt.preconditioner == MMPGO_PRECON_REGULARIZED_CHOLESKY;
    // the preconditioner is a per-node sparse solve, not a pose-local product: do it before the kernel that
    // projects its result (CG_INIT, PRECOND) or between the two halves of CG_STEP (the solve needs the new r)
    if (regchol && (op == V_CG_INIT || op == V_PRECOND)) RC(precon_solve(h, a_, m));
    Tiles tl; RC(make_tiles(h, m, &tl));
    VecArgs a; std::memset(&a, 0, sizeof(a));
    a.a = a_; a.b = b_; a.o1 = o1; a.o2 = o2; a.o3 = o3; a.o4 = o4; a.o5 = o5; a.y = y;
    a.pinv = h->d_pinv; a.precon = h->opt.preconditioner; a.coef = h->d_coef; a.partials = h->d_partials;
    a.pre = h->pre_buf;
    launch_vec<D>(op, tl, a, h->stream);
    h->ctr.launches++; h->ctr.vector_passes++;
    if (regchol && op == V_CG_STEP) {
      RC(precon_solve(h, o2, m));          // o2 = r, just updated
      RC(make_tiles(h, m, &tl));
      a.a = o2;
      launch_vec<D>(V_CG_PRE, tl, a, h->stream);
      h->ctr.launches++; h->ctr.vector_passes++;
    }
    return 0;
  }

  // Hess[x](p): tdot = -G00^{-1} G01 p_Y; Hp = Proj(Y, (G [tdot; p_Y])_Y - sym(nab Y^T) p_Y)
  // returns per-node p.Hp, Hp.Hp, p.p        (DPGOProblem.cpp:552-577)
  // (launch part: enqueues everything and the read-back of the three sums into slot 3, no synchronisation)
  static int hess_vec_launch(Handle *h, const double *x, double *p, double *Hp, const Mask &m, bool first) {
    Tiles tl; RC(make_tiles(h, m, &tl));
    GPassArgs a = gargs(h);
    a.x = p; a.g = nullptr; a.out = h->rhs_t;
    launch_gpass<D>(G_RHS_T, tl, a, h->stream);
    h->ctr.launches++; h->ctr.intra_passes++;
    if (first) {
      // the first search direction of this call is -P(grad), close to that of the previous outer
      // iteration: start the solve from the tdot found then (only the initial guess changes)
      RC(vec(h, V_COPY_T, m, h->tdot_prev, nullptr, p, nullptr, nullptr, nullptr, nullptr));
      RC(solve_t(h, p, m, true));
      RC(vec(h, V_COPY_T, m, p, nullptr, h->tdot_prev, nullptr, nullptr, nullptr, nullptr));
    } else {
      RC(solve_t(h, p, m, false));
    }
    RC(make_tiles(h, m, &tl));
    a = gargs(h);
    a.x = p; a.xref = x; a.nab = h->nab; a.out = Hp;
    launch_gpass<D>(G_HV, tl, a, h->stream);
    h->ctr.launches++; h->ctr.intra_passes++;
    return reduce_async(h, 3);
  }
  static int hess_vec(Handle *h, const double *x, double *p, double *Hp, const Mask &m, std::vector<double> &pHp,
                      std::vector<double> &HpHp, std::vector<double> &pp, bool first = false, bool launched = false) {
    if (!launched) {
      RC(hess_vec_launch(h, x, p, Hp, m, first));
      RC(host_sync(h));
    }
    const double *s = slot_host(h, 3);
    pHp.assign(h->A, 0.0); HpHp.assign(h->A, 0.0); pp.assign(h->A, 0.0);
    for (int n = 0; n < h->A; ++n) { pHp[n] = s[n * NS]; HpHp[n] = s[n * NS + 1]; pp[n] = s[n * NS + 2]; }
    return 0;
  }

  // ---- Steihaug-Toint truncated PCG, all nodes of `m` in lock step.  Result in h->cg_s.
  // `pre`: CG_INIT (sums in slot 2) and the first Hessian-vector product (slot 3) were already
  // launched for a superset of `m` and read back by the caller (speculation in tnt()).
  static int stpcg_init_launch(Handle *h, const double *x, const Mask &m) {
    RC(vec(h, V_CG_INIT, m, h->grad, nullptr, h->cg_s, h->cg_r, h->cg_v, h->cg_p, x, h->cg_Hs));
    return reduce_async(h, 2);
  }
  static int stpcg(Handle *h, const double *x, const Mask &m, const std::vector<double> &Delta,
                   std::vector<double> &hMnorm, std::vector<int> &inner, bool pre = false) {
    const int A = h->A;
    const mmpgo_options &o = h->opt;
    const double eps = 1e-8;                                      // IterativeSolvers.h:179
    std::vector<double> rv(A, 0), sk_M_pk(A, 0), sk_M_2(A, 0), pk_M_2(A, 0), target(A, 0), coef((size_t)A * MAXC, 0.0);
    Mask act = m;
    hMnorm.assign(A, 0.0); inner.assign(A, 0);
    if (!pre) {
      RC(stpcg_init_launch(h, x, m));
      RC(host_sync(h));
    }
    const double *s = slot_host(h, 2);
    for (int n = 0; n < A; ++n) if (m[n]) {
      rv[n] = s[n * NS];
      pk_M_2[n] = rv[n];
      const double r0 = std::sqrt(rv[n]);
      target[n] = r0 * std::min(o.STPCG_kappa, std::pow(r0, o.STPCG_theta));
    }
    std::vector<double> kap, HpHp, pp;
    bool first_hv = true;
    while (any(act)) {
      for (int n = 0; n < A; ++n) if (act[n]) {
        if (inner[n] >= o.max_tCG_iterations || std::sqrt(rv[n]) <= target[n]) {
          act[n] = 0; hMnorm[n] = std::sqrt(sk_M_2[n]);
        }
      }
      if (!any(act)) break;
      RC(hess_vec(h, x, h->cg_p, h->cg_Hp, act, kap, HpHp, pp, first_hv, pre && first_hv));
      first_hv = false;
      Mask fin(A, 0), cont(A, 0), kern(A, 0);
      for (int n = 0; n < A; ++n) if (act[n]) {
        if (std::sqrt(HpHp[n]) / std::sqrt(pp[n]) < eps) { kern[n] = 1; continue; }
        const double alpha = rv[n] / kap[n];
        const double skp1 = sk_M_2[n] + 2 * alpha * sk_M_pk[n] + alpha * alpha * pk_M_2[n];
        if (kap[n] <= 0 || skp1 > Delta[n] * Delta[n]) {
          fin[n] = 1;
          coef[(size_t)n * MAXC + 2] = (-sk_M_pk[n] + std::sqrt(sk_M_pk[n] * sk_M_pk[n] +
                                        pk_M_2[n] * (Delta[n] * Delta[n] - sk_M_2[n]))) / pk_M_2[n];
        } else {
          cont[n] = 1;
          coef[(size_t)n * MAXC + 0] = alpha;
          coef[(size_t)n * MAXC + 4] = skp1;
        }
      }
      if (any(kern)) {
        // search direction in the kernel of H: follow it to the boundary (IterativeSolvers.h:305-333)
        RC(vec(h, V_DOTS, kern, h->cg_p, h->cg_r, nullptr, nullptr, nullptr, nullptr, nullptr));
        RC(reduce_to_host(h, &s));
        for (int n = 0; n < A; ++n) if (kern[n]) {
          double sgn = 1.0, smp = sk_M_pk[n];
          if (s[n * NS] < 0) { sgn = -1.0; smp = -smp; }
          const double sigma = (-smp + std::sqrt(smp * smp + pk_M_2[n] * (Delta[n] * Delta[n] - sk_M_2[n]))) / pk_M_2[n];
          coef[(size_t)n * MAXC + 2] = sgn * sigma;
          fin[n] = 1;
        }
      }
      RC(upload_coef(h, coef));
      if (any(fin)) {
        RC(vec(h, V_CG_FINAL, fin, h->cg_p, h->cg_Hp, h->cg_s, nullptr, nullptr, nullptr, nullptr, h->cg_Hs));
        for (int n = 0; n < A; ++n) if (fin[n]) { act[n] = 0; hMnorm[n] = Delta[n]; }
      }
      if (any(cont)) {
        RC(vec(h, V_CG_STEP, cont, h->cg_p, h->cg_Hp, h->cg_s, h->cg_r, h->cg_v, nullptr, x, h->cg_Hs));
        RC(reduce_to_host(h, &s));
        for (int n = 0; n < A; ++n) if (cont[n]) {
          const double alpha = coef[(size_t)n * MAXC + 0];
          const double rk_vk = s[n * NS];
          const double beta = rk_vk / (alpha * kap[n]);
          sk_M_2[n] = coef[(size_t)n * MAXC + 4];
          sk_M_pk[n] = beta * (sk_M_pk[n] + alpha * pk_M_2[n]);
          pk_M_2[n] = rk_vk + beta * beta * pk_M_2[n];
          rv[n] = rk_vk;
          coef[(size_t)n * MAXC + 1] = beta;
          inner[n]++;
          h->ctr.tcg_iterations++;
        }
        RC(upload_coef(h, coef));
        RC(vec(h, V_CG_DIR, cont, h->cg_v, nullptr, h->cg_p, nullptr, nullptr, nullptr, nullptr));
      }
    }
    return 0;
  }

  // ---- truncated-Newton trust region on the nodes of `m`, in place on x.  fval: G(x) without offset.
  static int tnt(Handle *h, double *x, const double *g, const Mask &m, std::vector<double> &fx) {
    const int A = h->A;
    const mmpgo_options &o = h->opt;
    const double sqrt_eps = std::sqrt(std::numeric_limits<double>::epsilon());
    std::vector<double> gnorm(A, 0), pgnorm(A, 0), Delta(A, 1.0);   // Delta0 = 1, TNT.h:81
    std::vector<int> it(A, 0), acc(A, 0);
    Mask run = m;
    // the quadratic model pass (reduced gradient) also yields the surrogate value G(x): one pass
    // instead of evaluate_G + gradient (TNT.h:369-382)
    fx.assign(A, 0.0);
    bool speculate = false;
    auto quad_model = [&](const Mask &mm) -> int {
      Tiles tl; RC(make_tiles(h, mm, &tl));
      GPassArgs a = gargs(h);
      a.x = x; a.g = g; a.out = h->nab; a.out2 = h->grad;
      launch_gpass<D>(G_REDGRAD, tl, a, h->stream);
      h->ctr.launches++; h->ctr.intra_passes++;
      // gradient norm and preconditioned gradient norm: two kernels, one host synchronisation
      RC(reduce_async(h, 0));
      if (o.preconditioner != MMPGO_PRECON_NONE) {
        RC(vec(h, V_PRECOND, mm, h->grad, nullptr, nullptr, nullptr, nullptr, nullptr, x));
        RC(reduce_async(h, 1));
      }
      if (speculate) {
        // The tCG start (CG_INIT) and its first Hessian-vector product do not depend on these norms
        // except through the node mask (a node whose gradient is already small drops out): launch
        // them for all nodes of `mm` behind the model pass and read everything back with ONE
        // synchronisation; results of nodes that drop out are ignored.
        RC(stpcg_init_launch(h, x, mm));
        RC(hess_vec_launch(h, x, h->cg_p, h->cg_Hp, mm, true));
      }
      RC(host_sync(h));
      const double *s = slot_host(h, 0), *s1 = slot_host(h, 1);
      for (int n = 0; n < A; ++n) if (mm[n]) { gnorm[n] = std::sqrt(s[n * NS]); fx[n] = s[n * NS + 1]; }
      if (o.preconditioner != MMPGO_PRECON_NONE) {
        for (int n = 0; n < A; ++n) if (mm[n]) pgnorm[n] = std::sqrt(s1[n * NS]);
      } else {
        for (int n = 0; n < A; ++n) if (mm[n]) pgnorm[n] = gnorm[n];
      }
      return 0;
    };
    speculate = o.max_iterations > 0 && o.max_iterations_accepted > 0;
    RC(quad_model(m));
    bool pre = speculate;
    speculate = false;
    std::vector<double> hM, sHs, HsHs, ss, fprop;
    std::vector<int> inner;
    while (true) {
      for (int n = 0; n < A; ++n) if (run[n]) {
        if (it[n] >= o.max_iterations || acc[n] >= o.max_iterations_accepted) run[n] = 0;
        else if (gnorm[n] < o.grad_norm_tol) run[n] = 0;
        else if (pgnorm[n] < o.preconditioned_grad_norm_tol) run[n] = 0;
      }
      if (!any(run)) break;
      RC(stpcg(h, x, run, Delta, hM, inner, pre));
      pre = false;
      // trial point: retract, recover translations, evaluate
      RC(vec(h, V_RETRACT, run, x, h->cg_s, h->xprop, nullptr, nullptr, nullptr, nullptr));   // writes the whole pose block
      RC(recover_t(h, h->xprop, g, run));
      // G at the trial point, grad.h, |h|^2 and h.Hess h: three kernels, one host synchronisation.
      // H is linear and the step is s = sum alpha_k p_k, so H s was accumulated from the H p_k of the
      // tCG iterations (cg_Hs): no further Hessian-vector product (the reference recomputes it,
      // TNT.h:514-515; same value up to rounding)
      {
        Tiles tl; RC(make_tiles(h, run, &tl));
        GPassArgs a = gargs(h);
        a.x = h->xprop; a.g = g;
        launch_gpass<D>(G_EVAL, tl, a, h->stream);
        h->ctr.launches++; h->ctr.intra_passes++;
        RC(reduce_async(h, 0));
      }
      RC(vec(h, V_DOTS, run, h->grad, h->cg_s, nullptr, nullptr, nullptr, nullptr, nullptr));
      RC(reduce_async(h, 1));
      RC(vec(h, V_DOTS, run, h->cg_s, h->cg_Hs, nullptr, nullptr, nullptr, nullptr, nullptr));
      RC(reduce_async(h, 2));
      RC(host_sync(h));
      std::vector<double> gh(A, 0);
      fprop.assign(A, 0.0); ss.assign(A, 0.0); sHs.assign(A, 0.0);
      {
        const double *s0 = slot_host(h, 0), *s1 = slot_host(h, 1), *s2 = slot_host(h, 2);
        for (int n = 0; n < A; ++n) {
          if (run[n]) fprop[n] = s0[n * NS];
          gh[n] = s1[n * NS]; ss[n] = s1[n * NS + 2]; sHs[n] = s2[n * NS];
        }
      }
      Mask accm(A, 0), requad(A, 0);
      for (int n = 0; n < A; ++n) if (run[n]) {
        h->st[n].tcg_iterations += inner[n];
        h->st[n].tnt_iterations++;
        h->ctr.tnt_iterations++;
        const double h_norm = std::sqrt(ss[n]);
        const double dm = -gh[n] - 0.5 * sHs[n];
        const double df = fx[n] - fprop[n];
        // the reference compares f values that include the constant offset f_k; the
        // relative decrease uses |f(x)| with that offset (TNT.h:523)
        const double rel = df / (sqrt_eps + std::fabs(fx[n] + h->st[n].f));
        const double rho = df / dm;
        const bool ok = !std::isnan(rho) && rho > 0.05;               // eta1, TNT.h:84
        bool stop = false;
        if (ok) {
          acc[n]++; accm[n] = 1; fx[n] = fprop[n];
          if (rel < o.rel_func_decrease_tol || h_norm < o.stepsize_tol) stop = true;
          else if (acc[n] < o.max_iterations_accepted && it[n] + 1 < o.max_iterations) requad[n] = 1;
        }
        if (!stop) {
          if (!std::isnan(rho) && rho >= 0.9) Delta[n] = std::max(2.5 * hM[n], Delta[n]);   // eta2, alpha2
          else if (std::isnan(rho) || rho < 0.05) {
            Delta[n] = 0.25 * hM[n];                                                        // alpha1
            if (Delta[n] < 1e-6) stop = true;                                               // Delta_tolerance
          }
        }
        it[n]++;
        if (stop) run[n] = 0;
      }
      if (any(accm)) RC(vec(h, V_COPY, accm, h->xprop, nullptr, x, nullptr, nullptr, nullptr, nullptr));
      requad = mask_and(requad, run);
      if (any(requad)) RC(quad_model(requad));
    }
    return 0;
  }

  // ---- update()  -------------------------------------------------------------------
  static int update(Handle *h) {
    const int A = h->A;
    const mmpgo_options &o = h->opt;
    Mask m(A, 0);
    for (int n = 0; n < A; ++n) m[n] = !h->st[n].updated;
    if (!any(m)) return 0;
    // all nodes of a handle advance in lock step (one batched launch per operation): batch-wide
    // decisions below (first update, history terms) are taken once for the whole handle
    for (int n = 0; n < A; ++n)
      if (!m[n] || h->st[n].iters != h->st[0].iters) {
        set_error("update(): the nodes of a handle must share the iteration count and the updated state");
        return MMPGO_ERR_STATE;
      }
    const bool star = o.algorithm == MMPGO_ALG_STAR;
    const bool trivial = o.loss == MMPGO_LOSS_NONE;
    // rotate history: the "current" slot becomes "previous"
    h->icur ^= 1;
    double *gk = h->g[h->icur], *Dfk = h->Df[h->icur];
    const double *Xk = h->X[h->ik], *Xkm1 = h->X[h->ikm1];
    Tiles tl; RC(make_tiles(h, m, &tl));
    // DPGOStar::update_n always uses the *_f0 forms (DPGOStar.cpp:337-358)
    std::vector<int> first(A);
    for (int n = 0; n < A; ++n) first[n] = star || h->st[n].iters == 0;
    const bool use_diff = !first[0];
    InterArgs ia; std::memset(&ia, 0, sizeof(ia));
    ia.rowptr = h->d_xrowptr; ia.rec = h->d_xrec; ia.xa = Xk; ia.xb = use_diff ? Xkm1 : nullptr;
    ia.dinter = h->d_dinter; ia.gamma = nullptr; ia.use_diff = use_diff ? 1 : 0;
    ia.loss = o.loss; ia.loss_reg = o.loss_reg; ia.xi = o.regularizer;
    ia.g = gk; ia.partials = h->d_partials;
    if (!trivial) {
      std::swap(h->w_cur, h->w_prev);
      ia.w_prev = h->w_prev; ia.w_out = h->w_cur;
    }
    const bool dyn = h->dynamic;
    if (dyn) { ia.resc = h->resc; ia.split_g = 1; }
    launch_inter<D>(trivial ? I_TRIVIAL : I_ROBUST, tl, ia, h->stream);
    h->ctr.launches++; h->ctr.inter_passes++;
    RC(reduce_async(h, 0));
    if (dyn) {
      // Rescale::Dynamic (DPGOProblem.cpp:301-321, 465-485): the weights just found decide per node whether its
      // rescale vector is replaced; the history terms above used the old one, g / Dfobj / f below use the new one
      RC(host_sync(h));
      Mask resc(A, 0);
      const double *s0 = slot_host(h, 0);
      for (int n = 0; n < A; ++n) if (m[n]) {
        NodeState &st = h->st[n];
        if (st.rescale_count >= o.max_rescale_count || s0[n * NS + 6] > 0.0) { resc[n] = 1; st.rescale_count = 0; st.rescales++; }
        else st.rescale_count++;
      }
      if (any(resc)) {
        Tiles tr; RC(make_tiles(h, resc, &tr, h->d_active2));
        RescaleArgs ra; std::memset(&ra, 0, sizeof(ra));
        ra.rowptr = h->d_xrowptr; ra.rec = h->d_xrec; ra.w = h->w_cur; ra.resc = h->resc; ra.dintra = h->d_dintra;
        ra.dinter = h->d_dinter; ra.gdiag = h->d_gdiag; ra.tnv = h->d_tnv; ra.d00 = h->d_d00;
        ra.ts_rec = h->ts_rec; ra.pose_rec = h->d_pose_rec; ra.xi = o.regularizer;
        launch_rescale<D>(tr, ra, h->stream);
        h->ctr.launches++;
      }
      GFixArgs ga; std::memset(&ga, 0, sizeof(ga));
      ga.x = Xk; ga.dinter = h->d_dinter; ga.g = gk; ga.xi = o.regularizer; ga.partials = h->d_partials;
      launch_gfix<D>(tl, ga, h->stream);
      h->ctr.launches++;
      RC(reduce_async(h, 2));
    }
    // gradient pass: Dfobj = g + G x   (its sums are read back with those of K1: one synchronisation)
    GPassArgs a = gargs(h);
    a.x = Xk; a.g = gk; a.out = Dfk;
    launch_gpass<D>(G_GRAD, tl, a, h->stream);
    h->ctr.launches++; h->ctr.intra_passes++;
    RC(reduce_async(h, 1));
    RC(host_sync(h));
    const double *s = slot_host(h, 0);
    std::vector<double> i0(A), i1(A), i2(A), i3(A), i4(A), i5(A);
    for (int n = 0; n < A; ++n) { i0[n] = s[n*NS]; i1[n] = s[n*NS+1]; i2[n] = s[n*NS+2]; i3[n] = s[n*NS+3]; i4[n] = s[n*NS+4]; i5[n] = s[n*NS+5]; }
    if (dyn) for (int n = 0; n < A; ++n) i4[n] = slot_host(h, 2)[n * NS + 4];      // x^T D x with the new D (k_gfix)
    s = slot_host(h, 1);
    const double xi = o.regularizer;
    for (int n = 0; n < A; ++n) if (m[n]) {
      NodeState &st = h->st[n];
      const double xgGx = s[n*NS], grad2 = s[n*NS+1], xGx = s[n*NS+2];
      double fobj, f;
      if (trivial) {
        if (first[n]) {
          f = -i0[n];                    // f0 = 1/2 tr(Z^T P0 Z) = -q(z)        (DPGOProblem.cpp:279)
          fobj = xgGx + f;               // evaluate_G(Xak, g, f)               (DPGOHash.cpp:111-115)
        } else {
          fobj = st.Gk + i0[n];          // G + 1/2 tr(Y^T Q Y)                  (DPGOProblem.cpp:530-531)
          // f = fobj + 1/2 tr(Z^T P Z);  P = -(intra M) - offdiag(inter) + xi    (:532)
          const double xAx = xGx - (2.0 * i2[n] + xi * i3[n]);   // x^T (intra M) x
          f = fobj + (-0.5 * xAx - i1[n] + 0.5 * xi * i3[n]);
        }
      } else {
        const double fobjE = i0[n];
        if (first[n]) {
          // f0 = 1/2 fobjE + tr(X^T(1/2 D X - DfobjE))                          (DPGOProblem.cpp:233-244)
          f = 0.5 * fobjE + (0.5 * i4[n] - i3[n]);
          fobj = f + xgGx;
        } else {
          // fobj = G - 1/2 fobjE0 - 1/2 tr(Y^T(DfobjE0 + 1/2 Q Y)) + 1/2 fobjE      (:387-399)
          fobj = st.Gk - 0.5 * st.fobjE - 0.5 * (i1[n] + (i2[n] + xi * i5[n])) + 0.5 * fobjE;
          f = fobj - xgGx;
        }
        st.fobjE = fobjE;
      }
      st.fobj_prev = st.fobj;
      st.fobj = fobj; st.f = f;
      const int iter = st.iters;
      if (star) st.Gk = fobj;                                                       // DPGOStar.cpp:343
      if (iter == 0) { st.Fk[0] = st.Fk[1] = fobj; st.Gk = fobj; }                  // DPGOHash.cpp:143-147
      st.gradFnorm = std::sqrt(grad2);
      if (o.scheme == MMPGO_SCHEME_AMM) {
        if (iter == 0) { st.s_cur = 1.0; if (!star) { st.oscillations.clear(); st.oscillations.push_back(1); } }
        else st.s_cur = st.s_next;
        st.s_next = 0.5 + 0.5 * std::sqrt(4.0 * st.s_cur * st.s_cur + 1.0);          // :177
        st.gamma = (st.s_cur - 1.0) / st.s_next;                                    // :179
        if (!star) {
          if (fobj <= st.Fk[1]) st.soft_restart_hits[0] = st.soft_restart_hits[0] > 2 ? st.soft_restart_hits[0] - 2 : 0;
          else st.soft_restart_hits[0]++;
          if (iter > 0) {
            if (fobj <= st.fobj_prev) { st.soft_restart_hits[1] = 0; st.oscillations.push_back(1); }
            else { st.soft_restart_hits[1]++; st.oscillations.push_back(0); }
            st.num_oscillations += st.oscillations[iter] != st.oscillations[iter - 1];
          }
          if (iter > o.oscillation_cnt_period) {
            const int k = iter - o.oscillation_cnt_period;
            st.num_oscillations -= st.oscillations[k] != st.oscillations[k - 1];
          }
          st.Fk[0] = st.Fk[0] * (1 - o.eta[0]) + fobj * o.eta[0];
          st.Fk[1] = std::max(fobj, st.Fk[1] * (1 - o.eta[1]) + fobj * o.eta[1]);
        }
      } else if (!star) {
        st.Fk[0] = st.Fk[1] = fobj;
      }
      if (star) st.Fk[0] = st.Fk[1] = fobj;                                         // DPGOStar.cpp:384-385
      st.updated = true;
    }
    return 0;
  }

  static void upload_gamma(Handle *h) {
    std::vector<double> gm(h->A);
    for (int n = 0; n < h->A; ++n) gm[n] = h->st[n].iters == 0 ? 0.0 : h->st[n].gamma;
    cudaMemcpyAsync(h->d_gamma, gm.data(), sizeof(double) * h->A, cudaMemcpyHostToDevice, h->stream);
  }

  // extrapolated proximal step of amm_pgo / amm_pgo_n: fills Xakh, gex (and Dfex for robust losses)
  static int amm_proximal(Handle *h, const Mask &m, std::vector<double> *dist2) {
    const mmpgo_options &o = h->opt;
    const bool trivial = o.loss == MMPGO_LOSS_NONE;
    const double *Xk = h->X[h->ik], *Xkm1 = h->X[h->ikm1];
    const double *gk = h->g[h->icur], *gkm1 = h->g[h->icur ^ 1];
    const double *Dfk = h->Df[h->icur], *Dfkm1 = h->Df[h->icur ^ 1];
    const bool it0 = h->st[0].iters == 0;
    upload_gamma(h);
    if (it0) {
      return proximal(h, Xk, nullptr, Dfk, nullptr, gk, nullptr, h->gex, h->Xakh, Xk, m, dist2);
    }
    if (trivial) {
      // g and Df are linear in Z: extrapolate them with gamma (DPGOHash.cpp:259-262)
      return proximal(h, Xk, Xkm1, Dfk, Dfkm1, gk, gkm1, h->gex, h->Xakh, Xk, m, dist2);
    }
    // robust: evaluate_g_and_Df at the extrapolated point (DPGOHash.cpp:264, DPGOProblem.cpp:683-749)
    Tiles tl; RC(make_tiles(h, m, &tl));
    InterArgs ia; std::memset(&ia, 0, sizeof(ia));
    ia.rowptr = h->d_xrowptr; ia.rec = h->d_xrec; ia.xa = Xk; ia.xb = Xkm1; ia.dinter = h->d_dinter;
    ia.gamma = h->d_gamma; ia.use_diff = 0; ia.loss = o.loss; ia.loss_reg = o.loss_reg; ia.xi = o.regularizer;
    ia.w_out = h->w_tmp; ia.g = h->gex; ia.yex = h->Yex; ia.partials = h->d_partials;
    launch_inter<D>(I_ROBUST, tl, ia, h->stream);
    h->ctr.launches++; h->ctr.inter_passes++;
    GPassArgs a = gargs(h);
    a.x = h->Yex; a.g = h->gex; a.out = h->Dfex;
    launch_gpass<D>(G_GRAD, tl, a, h->stream);
    h->ctr.launches++; h->ctr.intra_passes++;
    return proximal(h, h->Yex, nullptr, h->Dfex, nullptr, nullptr, nullptr, nullptr, h->Xakh, Xk, m, dist2);
  }

  // ---- DPGOHash::amm_pgo for all nodes (DPGOHash.cpp:230-444)
  static int hash_amm(Handle *h) {
    const int A = h->A;
    const mmpgo_options &o = h->opt;
    const Mask allm(A, 1);
    const double *Xk = h->X[h->ik];
    double *Xak = h->X[h->iak];
    const double *gk = h->g[h->icur], *Dfk = h->Df[h->icur];
    Mask refined(A, 0);
    for (int n = 0; n < A; ++n) {
      const NodeState &st = h->st[n];
      refined[n] = (((st.gradFnorm * st.gradFnorm / st.fobj) > o.accepted_delta) ||
                    (st.num_oscillations >= o.max_oscillations)) &&
                   o.max_iterations > 0 && o.max_iterations_accepted > 0;
      h->st[n].refined = refined[n];
    }
    std::vector<double> dist2, Gkh, Gk, fx;
    // |X^{k+1/2} - X^k|^2 (K3) and G(X^{k+1/2}) (K2): two kernels, one host synchronisation
    RC(amm_proximal(h, allm, nullptr));
    RC(reduce_async(h, 0));
    {
      Tiles tl; RC(make_tiles(h, allm, &tl));
      GPassArgs a = gargs(h);
      a.x = h->Xakh; a.g = gk;
      launch_gpass<D>(G_EVAL, tl, a, h->stream);
      h->ctr.launches++; h->ctr.intra_passes++;
      RC(reduce_async(h, 1));
    }
    RC(host_sync(h));
    dist2.assign(A, 0.0); Gkh.assign(A, 0.0);
    for (int n = 0; n < A; ++n) { dist2[n] = slot_host(h, 0)[n * NS]; Gkh[n] = slot_host(h, 1)[n * NS]; }
    std::vector<double> minG(A);
    for (int n = 0; n < A; ++n) { Gkh[n] += h->st[n].f; minG[n] = h->st[n].Fk[0] - o.psi * dist2[n]; }
    // Xak.R = Xakh.R ; t = recover(g_extrapolated)
    RC(vec(h, V_COPY, allm, h->Xakh, nullptr, Xak, nullptr, nullptr, nullptr, nullptr));
    RC(recover_t(h, Xak, h->gex, allm));
    if (any(refined)) RC(tnt(h, Xak, h->gex, refined, fx));
    RC(eval_G(h, Xak, gk, allm, Gk));
    for (int n = 0; n < A; ++n) h->st[n].Gk = Gk[n] + h->st[n].f;
    // adaptive restart (DPGOHash.cpp:385-389)
    Mask redo(A, 0);
    for (int n = 0; n < A; ++n) redo[n] = Gkh[n] > minG[n];
    if (any(redo)) {
      RC(proximal(h, Xk, nullptr, Dfk, nullptr, nullptr, nullptr, nullptr, h->Xakh, nullptr, redo, nullptr));
      std::vector<double> v; RC(eval_G(h, h->Xakh, gk, redo, v));
      for (int n = 0; n < A; ++n) if (redo[n]) Gkh[n] = v[n] + h->st[n].f;
    }
    Mask hard(A, 0), restart(A, 0), use_gk(A, 0);
    for (int n = 0; n < A; ++n) {
      const NodeState &st = h->st[n];
      hard[n] = st.Gk > st.Fk[0];
      const bool soft = (st.Gk > st.Fk[1] && st.soft_restart_hits[0] >= o.max_soft_restart_hits[0]) ||
                        (st.Gk > st.fobj && st.soft_restart_hits[1] > o.max_soft_restart_hits[1]);
      restart[n] = hard[n] || soft;
    }
    if (any(restart)) {
      Mask from_h(A, 0), from_prox(A, 0);
      for (int n = 0; n < A; ++n) if (restart[n]) {
        h->st[n].restarts++;
        use_gk[n] = 1;
        if (Gkh[n] <= h->st[n].fobj) from_h[n] = 1; else from_prox[n] = 1;
      }
      if (any(from_h)) RC(vec(h, V_COPY, from_h, h->Xakh, nullptr, Xak, nullptr, nullptr, nullptr, nullptr));
      if (any(from_prox)) RC(proximal(h, Xk, nullptr, Dfk, nullptr, nullptr, nullptr, nullptr, Xak, nullptr, from_prox, nullptr));
      RC(recover_t(h, Xak, gk, restart));
      const Mask rr = mask_and(restart, refined);
      Mask rn(A, 0);
      for (int n = 0; n < A; ++n) rn[n] = restart[n] && !refined[n];
      if (any(rr)) {
        RC(tnt(h, Xak, gk, rr, fx));
        for (int n = 0; n < A; ++n) if (rr[n]) h->st[n].Gk = fx[n] + h->st[n].f;
      }
      if (any(rn)) {
        std::vector<double> v; RC(eval_G(h, Xak, gk, rn, v));
        for (int n = 0; n < A; ++n) if (rn[n]) h->st[n].Gk = v[n] + h->st[n].f;
      }
      for (int n = 0; n < A; ++n) if (restart[n]) {
        if (hard[n]) h->st[n].s_next = std::max(0.5 * h->st[n].s_next, 1.0);
        h->st[n].soft_restart_hits[0] /= 3;
        h->st[n].soft_restart_hits[1] = 0;
      }
    }
    // final safeguard (DPGOHash.cpp:434-441)
    Mask safe(A, 0);
    for (int n = 0; n < A; ++n) {
      const NodeState &st = h->st[n];
      safe[n] = (st.Fk[0] - st.Gk) < o.phi * (st.Fk[0] - Gkh[n]);
    }
    if (any(safe)) {
      RC(vec(h, V_COPY_ROT, safe, h->Xakh, nullptr, Xak, nullptr, nullptr, nullptr, nullptr));
      const Mask s1 = mask_and(safe, use_gk);
      Mask s2(A, 0);
      for (int n = 0; n < A; ++n) s2[n] = safe[n] && !use_gk[n];
      if (any(s1)) RC(recover_t(h, Xak, gk, s1));
      if (any(s2)) RC(recover_t(h, Xak, h->gex, s2));
      std::vector<double> v; RC(eval_G(h, Xak, gk, safe, v));
      for (int n = 0; n < A; ++n) if (safe[n]) h->st[n].Gk = v[n] + h->st[n].f;
    }
    return 0;
  }

  // ---- DPGOHash::mm_pgo for all nodes (DPGOHash.cpp:446-581)
  static int hash_mm(Handle *h) {
    const int A = h->A;
    const mmpgo_options &o = h->opt;
    const Mask allm(A, 1);
    const double *Xk = h->X[h->ik];
    double *Xak = h->X[h->iak];
    const double *gk = h->g[h->icur], *Dfk = h->Df[h->icur];
    Mask refined(A, 0), plain(A, 0);
    for (int n = 0; n < A; ++n) {
      const NodeState &st = h->st[n];
      refined[n] = ((st.gradFnorm * st.gradFnorm / st.fobj) > o.accepted_delta) && o.max_iterations > 0 &&
                   o.max_iterations_accepted > 0;
      plain[n] = !refined[n];
      h->st[n].refined = refined[n];
    }
    RC(proximal(h, Xk, nullptr, Dfk, nullptr, nullptr, nullptr, nullptr, h->Xakh, nullptr, allm, nullptr));
    RC(recover_t(h, h->Xakh, gk, allm));
    RC(vec(h, V_COPY, allm, h->Xakh, nullptr, Xak, nullptr, nullptr, nullptr, nullptr));
    std::vector<double> fx;
    if (any(refined)) {
      RC(tnt(h, Xak, gk, refined, fx));
      for (int n = 0; n < A; ++n) if (refined[n]) h->st[n].Gk = fx[n] + h->st[n].f;
    }
    if (any(plain)) {
      std::vector<double> v; RC(eval_G(h, Xak, gk, plain, v));
      for (int n = 0; n < A; ++n) if (plain[n]) h->st[n].Gk = v[n] + h->st[n].f;
    }
    return 0;
  }

  // global objective of a candidate iterate held in pose blocks (own + halo rows)
  static int edge_objective(Handle *h, double *x, double *f, bool exchange = true) {
    if (exchange) RC(halo_exchange(h, x));
    int nb = 0;
    launch_edge_objective<D>(h->n_edges_owned, h->d_eidx, h->d_eval, x, h->opt.loss, h->opt.loss_reg, h->d_block_partials,
                             &nb, h->stream);
    launch_sum_blocks(nb, h->d_block_partials, h->d_scalar, h->stream);
    h->ctr.launches += 2; h->ctr.inter_passes++;
    CK(cudaMemcpyAsync(h->h_pinned, h->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *f = h->h_pinned[0];
    return allreduce(h, f, 1);
  }
  static int diff2(Handle *h, const double *a, const double *b, double *out) {
    const Mask allm(h->A, 1);
    RC(vec(h, V_DIFFNORM, allm, a, b, nullptr, nullptr, nullptr, nullptr, nullptr));
    const double *s; RC(reduce_to_host(h, &s));
    double t = 0.0;
    for (int n = 0; n < h->A; ++n) t += s[n * NS];
    *out = t;
    return allreduce(h, out, 1);
  }

  // ---- DPGOStar::iterate (DPGOStar.cpp:126-213), single-handle form (all nodes local)
  static int star_iterate(Handle *h) {
    const int A = h->A;
    const mmpgo_options &o = h->opt;
    const Mask allm(A, 1);
    double *Xk = h->X[h->ik];
    double *Xak = h->X[h->iak];      // X^{k+1} candidate ("Xkp")
    const double *gk = h->g[h->icur], *Dfk = h->Df[h->icur];
    Mask refined(A, 0), plain(A, 0);
    for (int n = 0; n < A; ++n) {
      const NodeState &st = h->st[n];
      refined[n] = (st.gradFnorm * st.gradFnorm / st.fobj) > o.accepted_delta;     // DPGOStar.cpp:515-516
      plain[n] = !refined[n];
      h->st[n].refined = refined[n];
    }
    std::vector<double> fx;
    // amm_pgo_n for all nodes
    RC(amm_proximal(h, allm, nullptr));
    RC(vec(h, V_COPY, allm, h->Xakh, nullptr, Xak, nullptr, nullptr, nullptr, nullptr));
    RC(recover_t(h, Xak, h->gex, allm));
    if (any(refined) && o.max_iterations > 0 && o.max_iterations_accepted > 0) RC(tnt(h, Xak, h->gex, refined, fx));
    // F(X^{k+1/2}), |X^{k+1/2} - X^k|^2, F(X^{k+1}), |X^{k+1} - X^k|^2: four kernels, one host
    // synchronisation and ONE all-reduce of four scalars (DPGOStar.cpp:147-159 evaluates them one
    // after the other; X^{k+1} does not depend on the outcome of the first test)
    double fobjh, fobj, dh, dp;
    {
      RC(halo_exchange2(h, h->Xakh, Xak));
      int nb = 0;
      launch_edge_objective<D>(h->n_edges_owned, h->d_eidx, h->d_eval, h->Xakh, o.loss, o.loss_reg, h->d_block_partials, &nb, h->stream);
      launch_sum_blocks(nb, h->d_block_partials, h->d_scalar, h->stream);
      launch_edge_objective<D>(h->n_edges_owned, h->d_eidx, h->d_eval, Xak, o.loss, o.loss_reg, h->d_block_partials, &nb, h->stream);
      launch_sum_blocks(nb, h->d_block_partials, h->d_scalar + 1, h->stream);
      h->ctr.launches += 4; h->ctr.inter_passes += 2;
      RC(vec(h, V_DIFFNORM, allm, h->Xakh, Xk, nullptr, nullptr, nullptr, nullptr, nullptr));
      launch_reduce(A, h->d_node_tb, h->d_node_te, h->d_partials, h->d_node_scal, h->stream);
      RC(vec(h, V_DIFFNORM, allm, Xak, Xk, nullptr, nullptr, nullptr, nullptr, nullptr));
      launch_reduce(A, h->d_node_tb, h->d_node_te, h->d_partials, h->d_node_scal2, h->stream);
      h->ctr.launches += 2;
      double *hp = h->h_pinned;
      double v[4];
      if (h->world > 1 && (h->allreduce_dev_fn || h->nccl_comm)) {
        // the four scalars are reduced on the device (stream-ordered), then read back once
        launch_sum_strided(A, h->d_node_scal, NS, h->d_scalar + 2, h->stream);
        launch_sum_strided(A, h->d_node_scal2, NS, h->d_scalar + 3, h->stream);
        h->ctr.launches += 2;
        h->allreduces++;
        if (h->nccl_comm) RC(nccl_allreduce(h, h->d_scalar, 4));
        else if (h->allreduce_dev_fn(h->cb_user, h->d_scalar, 4) != 0) { set_error("device allreduce callback failed"); return MMPGO_ERR_ARG; }
        CK(cudaMemcpyAsync(hp, h->d_scalar, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        v[0] = hp[0]; v[2] = hp[1]; v[1] = hp[2]; v[3] = hp[3];
      } else {
        CK(cudaMemcpyAsync(hp, h->d_scalar, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(hp + 8, h->d_node_scal, sizeof(double) * A * NS, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(hp + 8 + (size_t)A * NS, h->d_node_scal2, sizeof(double) * A * NS, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        v[0] = hp[0]; v[1] = 0.0; v[2] = hp[1]; v[3] = 0.0;
        for (int n = 0; n < A; ++n) { v[1] += hp[8 + (size_t)n * NS]; v[3] += hp[8 + (size_t)(A + n) * NS]; }
        RC(allreduce(h, v, 4));
      }
      fobjh = v[0]; dh = v[1]; fobj = v[2]; dp = v[3];
    }
    if (fobjh > h->starF - o.psi * dh) {
      // pm_pgo_n: plain proximal from Xk (DPGOStar.cpp:685-711)
      RC(proximal(h, Xk, nullptr, Dfk, nullptr, nullptr, nullptr, nullptr, h->Xakh, nullptr, allm, nullptr));
      RC(edge_objective(h, h->Xakh, &fobjh));
    }
    if (fobj > h->starF - o.psi * dp) {
      // global restart: mm_pgo_n for all nodes + halve s (DPGOStar.cpp:159-169)
      h->star_restarts++;
      RC(vec(h, V_COPY_ROT, allm, h->Xakh, nullptr, Xak, nullptr, nullptr, nullptr, nullptr));
      RC(recover_t(h, Xak, gk, allm));
      if (any(refined) && o.max_iterations > 0 && o.max_iterations_accepted > 0) {
        RC(tnt(h, Xak, gk, refined, fx));
        for (int n = 0; n < A; ++n) if (refined[n]) h->st[n].Gk = fx[n] + h->st[n].f;
      }
      if (any(plain)) {
        std::vector<double> v; RC(eval_G(h, Xak, gk, plain, v));
        for (int n = 0; n < A; ++n) if (plain[n]) h->st[n].Gk = v[n] + h->st[n].f;
      }
      for (int n = 0; n < A; ++n) { h->st[n].s_next = std::max(0.5 * h->st[n].s_next, 1.0); h->st[n].restarts++; }
      RC(edge_objective(h, Xak, &fobj));
    }
    if (h->starF - fobj < o.phi * (h->starF - fobjh)) {
      RC(vec(h, V_COPY_ROT, allm, h->Xakh, nullptr, Xak, nullptr, nullptr, nullptr, nullptr));
      RC(recover_t(h, Xak, gk, allm));
      RC(edge_objective(h, Xak, &fobj));
    }
    h->star_fobj = fobj;
    h->starF = h->starF * (1 - o.eta[0]) + fobj * o.eta[0];
    // every path above ends with a halo exchange of the final X^{k+1} (the merged exchange, or the
    // one inside edge_objective after a restart / safeguard): communicate() need not repeat it
    h->next_halo_current = true;
    return 0;
  }

  static int iterate(Handle *h) {
    for (int n = 0; n < h->A; ++n)
      if (!h->st[n].updated) { set_error("iterate() before update()"); return MMPGO_ERR_STATE; }
    if (h->opt.algorithm == MMPGO_ALG_STAR) RC(star_iterate(h));
    else if (h->opt.scheme == MMPGO_SCHEME_AMM) RC(hash_amm(h));
    else RC(hash_mm(h));
    for (int n = 0; n < h->A; ++n) { h->st[n].iters++; h->st[n].updated = false; }
    // publish X^{k+1}: the reference copies Xak into the head of Xk at the end of iterate()
    // (DPGOHash.cpp:612-616, DPGOStar.cpp:194-208); on the device the three pose buffers rotate.
    // The neighbour copies stay those of iteration k until communicate() refreshes them, as in
    // the reference, so communicate() may be repeated or (for a graph without remote
    // neighbours) skipped.
    const int old_km1 = h->ikm1;
    h->ikm1 = h->ik; h->ik = h->iak; h->iak = old_km1;
    if (h->NH > 0 && !h->next_halo_current)
      CK(cudaMemcpyAsync(h->X[h->ik] + (size_t)h->NO * PB, h->X[h->ikm1] + (size_t)h->NO * PB,
                         sizeof(double) * (size_t)h->NH * PB, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
  }

  // DPGOHash::communicate (DPGOHash.h:28-86) / DPGOStar::communicate (DPGOStar.cpp:215-223): refresh
  // the copies of the remote neighbours.  Neighbours on the same GPU are read in place.
  static int communicate(Handle *h) {
    if (h->next_halo_current) { h->next_halo_current = false; return 0; }
    return halo_exchange(h, h->X[h->ik]);
  }
};

// ---- layout conversion between the reference's global X and device pose blocks ------
// The host matrix goes to the device as it is (one strided copy, d columns of (d+1)N doubles)
// and is re-laid out there; nothing is converted on the host.
static int stage_alloc(Handle *h) {
  if (h->d_xstage) return 0;
  const size_t n = (size_t)(h->d + 1) * h->N * h->d;
  void *q = nullptr;
  CK(cudaMalloc(&q, n * sizeof(double)));
  CK(cudaMemsetAsync(q, 0, n * sizeof(double), h->stream));
  h->allocs.push_back(q);
  h->d_xstage = static_cast<double *>(q);
  return 0;
}
static int stage_upload(Handle *h, const double *X, int64_t ldx) {
  RC(stage_alloc(h));
  // only the rows the handle reads travel: translation rows [lo, hi) and rotation rows [N + d lo, N + d hi)
  const size_t ld = (size_t)(h->d + 1) * h->N;
  const size_t lo = (size_t)h->stage_lo, n = (size_t)(h->stage_hi - h->stage_lo), d = (size_t)h->d;
  CK(cudaMemcpy2DAsync(h->d_xstage + lo, ld * sizeof(double), X + lo, (size_t)ldx * sizeof(double), n * sizeof(double),
                       d, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpy2DAsync(h->d_xstage + h->N + d * lo, ld * sizeof(double), X + h->N + d * lo, (size_t)ldx * sizeof(double),
                       d * n * sizeof(double), d, cudaMemcpyHostToDevice, h->stream));
  return 0;
}
static void pack_dev(Handle *h, double *d0, double *d1, double *d2, double *d3, double *d4) {
  const int64_t ld = (int64_t)(h->d + 1) * h->N;
  if (h->d == 2) launch_pack_poses<2>(h->NP, h->d_pose_gid, h->d_xstage, ld, h->N, d0, d1, d2, d3, d4, h->stream);
  else launch_pack_poses<3>(h->NP, h->d_pose_gid, h->d_xstage, ld, h->N, d0, d1, d2, d3, d4, h->stream);
  h->ctr.launches++;
}

int driver_initialize(Handle *h, const double *X, int64_t ldx) {
  if (!h->graph_set) { set_error("set_graph first"); return MMPGO_ERR_STATE; }
  if (ldx < (int64_t)(h->d + 1) * h->N) { set_error("ldx too small"); return MMPGO_ERR_ARG; }
  RC(stage_upload(h, X, ldx));
  h->ik = 0; h->ikm1 = 1; h->iak = 2; h->icur = 0;
  h->next_halo_current = false;
  pack_dev(h, h->X[0], h->X[1], h->X[2], h->Xakh, h->xprop);
  CK(cudaMemsetAsync(h->tdot_prev, 0, sizeof(double) * (size_t)h->NO * (h->d + 1) * h->d, h->stream));
  for (auto &s : h->st) { s = NodeState(); s.updated = false; }
  h->star_restarts = 0;
  if (h->dynamic) {
    // Rescale::Dynamic starts from the all-ones rescale vector (DPGOProblem.cpp:31, 84): rebuild the per-pose
    // constants with the kernel that will maintain them (w = 0.8 => s = clamp(1.25 w) = 1)
    std::vector<double> w((size_t)h->n_inter_he, 0.8);
    CK(cudaMemcpyAsync(h->w_tmp, w.data(), sizeof(double) * w.size(), cudaMemcpyHostToDevice, h->stream));
    Tiles tl;
    tl.n_tiles = h->n_tiles; tl.node = h->d_tile_node; tl.start = h->d_tile_start; tl.cnt = h->d_tile_cnt; tl.active = nullptr;
    RescaleArgs ra; std::memset(&ra, 0, sizeof(ra));
    ra.rowptr = h->d_xrowptr; ra.rec = h->d_xrec; ra.w = h->w_tmp; ra.resc = h->resc; ra.dintra = h->d_dintra;
    ra.dinter = h->d_dinter; ra.gdiag = h->d_gdiag; ra.tnv = h->d_tnv; ra.d00 = h->d_d00;
    ra.ts_rec = h->ts_rec; ra.pose_rec = h->d_pose_rec; ra.xi = h->opt.regularizer;
    if (h->d == 2) launch_rescale<2>(tl, ra, h->stream); else launch_rescale<3>(tl, ra, h->stream);
    h->ctr.launches++;
    CK(cudaStreamSynchronize(h->stream));          // w is a temporary
  }
  if (h->opt.algorithm == MMPGO_ALG_STAR) {
    double f = 0.0;
    int rc = h->d == 2 ? Drv<2>::edge_objective(h, h->X[h->ik], &f, false)
                       : Drv<3>::edge_objective(h, h->X[h->ik], &f, false);
    if (rc) return rc;
    h->star_fobj = f; h->starF = f;                                   // DPGOStar.cpp:120-122
  } else {
    CK(cudaStreamSynchronize(h->stream));                             // X is borrowed for the call only
  }
  h->initialized = true;
  return 0;
}

// The reference's L_.solve is exact (CHOLMOD, DPGOProblem.cpp:140, 568); a PCG solve that stops on
// translation_solve_max_iters above translation_solve_tol is a deviation and is reported, not hidden:
// the first driver call that observes the device-side count (it travels with every PCG launch and is
// visible after the call's next host synchronisation; update() always ends with one) fails.
static int check_pcg(Handle *h) {
  if (!h->h_ts_stats || h->h_ts_stats[2] <= h->ts_unconv_seen) return 0;
  const unsigned long long n = h->h_ts_stats[2] - h->ts_unconv_seen;
  h->ts_unconv_seen = h->h_ts_stats[2];
  char buf[256];
  std::snprintf(buf, sizeof(buf), "translation solve: %llu node solve(s) stopped at translation_solve_max_iters = %d "
                "above translation_solve_tol = %g (raise the limit or use the direct solver)", n,
                h->opt.translation_solve_max_iters, h->opt.translation_solve_tol);
  set_error(buf);
  return MMPGO_ERR_NOT_CONVERGED;
}
int driver_update(Handle *h) {
  if (!h->initialized) { set_error("initialize first"); return MMPGO_ERR_STATE; }
  const int rc = h->d == 2 ? Drv<2>::update(h) : Drv<3>::update(h);
  return rc ? rc : check_pcg(h);
}
int driver_iterate(Handle *h) {
  if (!h->initialized) { set_error("initialize first"); return MMPGO_ERR_STATE; }
  const int rc = h->d == 2 ? Drv<2>::iterate(h) : Drv<3>::iterate(h);
  return rc ? rc : check_pcg(h);
}
int driver_communicate(Handle *h) {
  if (!h->initialized) { set_error("initialize first"); return MMPGO_ERR_STATE; }
  return h->d == 2 ? Drv<2>::communicate(h) : Drv<3>::communicate(h);
}

int driver_get_poses(Handle *h, double *X, int64_t ldx) {
  if (!h->initialized) { set_error("initialize first"); return MMPGO_ERR_STATE; }
  RC(stage_alloc(h));
  const int d = h->d;
  const int64_t ld = (int64_t)(d + 1) * h->N;
  if (d == 2) launch_unpack_poses<2>(h->NO, h->d_pose_gid, h->X[h->ik], h->d_xstage, ld, h->N, h->stream);
  else launch_unpack_poses<3>(h->NO, h->d_pose_gid, h->X[h->ik], h->d_xstage, ld, h->N, h->stream);
  h->ctr.launches++;
  // only the rows of the local nodes travel back: own poses are the id range [g_lo, g_hi)
  const int64_t g_lo = h->own_gid.front(), g_hi = h->own_gid.back() + 1;
  CK(cudaMemcpy2DAsync(X + g_lo, (size_t)ldx * sizeof(double), h->d_xstage + g_lo, (size_t)ld * sizeof(double),
                       (size_t)(g_hi - g_lo) * sizeof(double), d, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpy2DAsync(X + h->N + d * g_lo, (size_t)ldx * sizeof(double), h->d_xstage + h->N + d * g_lo,
                       (size_t)ld * sizeof(double), (size_t)(g_hi - g_lo) * d * sizeof(double), d,
                       cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int driver_get_weights(Handle *h, int node, double *w, int64_t cap, int64_t *count) {
  if (node < h->node_begin || node >= h->node_end) { set_error("node not local"); return MMPGO_ERR_ARG; }
  const NodeInfo &ni = h->info[node - h->node_begin];
  *count = (int64_t)ni.inter_he.size();
  if (!w) return 0;
  if (cap < *count) { set_error("weight buffer too small"); return MMPGO_ERR_ARG; }
  std::vector<double> all((size_t)h->n_inter_he);
  if (h->opt.loss == MMPGO_LOSS_NONE) {
    for (int64_t k = 0; k < *count; ++k) w[k] = 1.0;
    return 0;
  }
  CK(cudaMemcpyAsync(all.data(), h->w_cur, all.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int64_t k = 0; k < *count; ++k) w[k] = all[ni.inter_he[k]];
  return 0;
}

int driver_evaluate_f(Handle *h, const double *X, int64_t ldx, double *f) {
  if (!h->graph_set) { set_error("set_graph first"); return MMPGO_ERR_STATE; }
  RC(stage_upload(h, X, ldx));
  pack_dev(h, h->xeval, nullptr, nullptr, nullptr, nullptr);
  // evaluate_f reports the LOCAL partial sum (see mmpgo.h); halo rows come from X itself
  const int world = h->world;
  h->world = 1;
  const int rc = h->d == 2 ? Drv<2>::edge_objective(h, h->xeval, f, false) : Drv<3>::edge_objective(h, h->xeval, f, false);
  h->world = world;
  return rc;
}

// DPGOStar::evaluate_grad (DPGOStar.cpp:763-829).  At Z = X the surrogate's gradient is the
// gradient of F, so the Euclidean gradient of the own poses is K1 (inter-node edges, weights
// evaluated at X) + the K2 gradient pass, exactly what update() computes for X^k; the
// reduced-gradient pass adds the tangent projection of the rotation rows.  Scratch only
// (TNT vectors, the evaluation pose array): the solver state is not touched.
template <int D> static int evaluate_grad_t(Handle *h, const double *X, int64_t ldx, double *G, int64_t ldg) {
  typedef Drv<D> Dr;
  const mmpgo_options &o = h->opt;
  const Mask allm(h->A, 1);
  RC(stage_upload(h, X, ldx));
  pack_dev(h, h->xeval, nullptr, nullptr, nullptr, nullptr);
  Tiles tl; RC(make_tiles(h, allm, &tl));
  double *g = h->cg_r, *Df = h->cg_v, *grad = h->cg_p;
  InterArgs ia; std::memset(&ia, 0, sizeof(ia));
  ia.rowptr = h->d_xrowptr; ia.rec = h->d_xrec; ia.xa = h->xeval; ia.dinter = h->d_dinter;
  ia.loss = o.loss; ia.loss_reg = o.loss_reg; ia.xi = o.regularizer;
  ia.w_out = h->w_tmp; ia.g = g; ia.partials = h->d_partials;
  launch_inter<D>(o.loss == MMPGO_LOSS_NONE ? I_TRIVIAL : I_ROBUST, tl, ia, h->stream);
  GPassArgs a = Dr::gargs(h);
  a.x = h->xeval; a.g = g; a.out = Df;
  launch_gpass<D>(G_GRAD, tl, a, h->stream);                 // all rows of the Euclidean gradient
  a.out = h->cg_Hp; a.out2 = grad;
  launch_gpass<D>(G_REDGRAD, tl, a, h->stream);              // rotation rows, projected (SOdProduct::Proj)
  VecArgs v; std::memset(&v, 0, sizeof(v));
  v.a = Df; v.o1 = grad; v.partials = h->d_partials;
  launch_vec<D>(V_COPY_T, tl, v, h->stream);                 // translation rows stay Euclidean
  h->ctr.launches += 4; h->ctr.inter_passes++; h->ctr.intra_passes += 2; h->ctr.vector_passes++;
  // pose blocks -> the caller's layout; only the rows of the local nodes travel back
  const int d = h->d;
  const int64_t ld = (int64_t)(d + 1) * h->N;
  launch_unpack_poses<D>(h->NO, h->d_pose_gid, grad, h->d_xstage, ld, h->N, h->stream);
  h->ctr.launches++;
  const int64_t g_lo = h->own_gid.front(), g_hi = h->own_gid.back() + 1;
  CK(cudaMemcpy2DAsync(G + g_lo, (size_t)ldg * sizeof(double), h->d_xstage + g_lo, (size_t)ld * sizeof(double),
                       (size_t)(g_hi - g_lo) * sizeof(double), d, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpy2DAsync(G + h->N + d * g_lo, (size_t)ldg * sizeof(double), h->d_xstage + h->N + d * g_lo,
                       (size_t)ld * sizeof(double), (size_t)(g_hi - g_lo) * d * sizeof(double), d,
                       cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int driver_evaluate_grad(Handle *h, const double *X, int64_t ldx, double *G, int64_t ldg) {
  if (!h->graph_set) { set_error("set_graph first"); return MMPGO_ERR_STATE; }
  return h->d == 2 ? evaluate_grad_t<2>(h, X, ldx, G, ldg) : evaluate_grad_t<3>(h, X, ldx, G, ldg);
}

// t = -G00^{-1} rhs for every local node through the solve path of the handle (dense inverse, sparse
// Cholesky sweeps or PCG): the core of recover_translations (DPGOProblem.h:275-294), on caller data.
template <int D> static int translation_solve_t(Handle *h, const double *rhs, double *t) {
  typedef Drv<D> Dr;
  constexpr int PB = (D + 1) * D;
  const Mask allm(h->A, 1);
  std::vector<double> prhs;
  if (h->use_direct) {       // the sparse direct solve takes its right-hand side in elimination order
    prhs.resize((size_t)h->NO * D);
    for (int p = 0; p < h->NO; ++p)
      for (int c = 0; c < D; ++c) prhs[(size_t)h->h_mf_perm[p] * D + c] = rhs[(size_t)p * D + c];
    rhs = prhs.data();
  }
  CK(cudaMemcpyAsync(h->rhs_t, rhs, sizeof(double) * (size_t)h->NO * D, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->xeval, 0, sizeof(double) * (size_t)h->NO * PB, h->stream));
  RC(Dr::solve_t(h, h->xeval, allm, false));
  CK(cudaMemcpy2DAsync(t, sizeof(double) * D, h->xeval, sizeof(double) * PB, sizeof(double) * D, (size_t)h->NO,
                       cudaMemcpyDeviceToHost, h->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int driver_translation_solve(Handle *h, const double *rhs, double *t) {
  if (!h->graph_set) { set_error("set_graph first"); return MMPGO_ERR_STATE; }
  const int rc = h->d == 2 ? translation_solve_t<2>(h, rhs, t) : translation_solve_t<3>(h, rhs, t);
  return rc ? rc : check_pcg(h);     // the solve above ended with a host synchronisation
}

// Times `reps` back-to-back launches of one hot kernel on the handle's stream with CUDA
// events (bench.py's live roofline measurement).  kind: 0 G_EVAL, 1 G_GRAD, 2 inter pass,
// 3 fused proximal, 4 edge objective, 5 G00 SpMV, 6 G_HV, 7 G_RHS_T.
template <int D> static int profile_pass(Handle *h, int kind, int reps, float *ms_avg) {
  typedef Drv<D> Dr;
  const Mask allm(h->A, 1);
  Tiles tl; RC(make_tiles(h, allm, &tl));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double *Xk = h->X[h->ik];
  auto one = [&]() -> int {
    if (kind == 0 || kind == 1 || kind == 6 || kind == 7) {
      GPassArgs a = Dr::gargs(h);
      a.x = Xk; a.g = h->g[h->icur]; a.out = kind == 7 ? h->rhs_t : h->Dfex; a.out2 = h->grad;
      a.xref = Xk; a.nab = h->Df[h->icur];
      launch_gpass<D>(kind == 0 ? G_EVAL : kind == 1 ? G_GRAD : kind == 6 ? G_HV : G_RHS_T, tl, a, h->stream);
    } else if (kind == 2) {
      InterArgs ia; std::memset(&ia, 0, sizeof(ia));
      ia.rowptr = h->d_xrowptr; ia.rec = h->d_xrec; ia.xa = Xk; ia.xb = nullptr; ia.dinter = h->d_dinter;
      ia.loss = h->opt.loss; ia.loss_reg = h->opt.loss_reg; ia.xi = h->opt.regularizer;
      ia.w_out = h->w_tmp; ia.g = h->gex; ia.partials = h->d_partials;
      launch_inter<D>(h->opt.loss == MMPGO_LOSS_NONE ? I_TRIVIAL : I_ROBUST, tl, ia, h->stream);
    } else if (kind == 3) {
      ProxArgs a; std::memset(&a, 0, sizeof(a));
      a.xa = Xk; a.xb = h->X[h->ikm1]; a.dfa = h->Df[h->icur]; a.dfb = h->Df[h->icur ^ 1];
      a.ga = h->g[h->icur]; a.gb = h->g[h->icur ^ 1]; a.gamma = h->d_gamma; a.tnv = h->d_tnv; a.xref = Xk;
      a.xout = h->xprop; a.gex = h->gex; a.partials = h->d_partials;
      launch_prox<D>(tl, a, h->stream);
    } else if (kind == 4) {
      int nb = 0;
      launch_edge_objective<D>(h->n_edges_owned, h->d_eidx, h->d_eval, Xk, h->opt.loss, h->opt.loss_reg, h->d_block_partials,
                               &nb, h->stream);
    } else if (kind >= 8 && kind <= 11) {
      h->mf_dry = kind == 9;                 // 9: the sparse direct solve's stages and barriers without its jobs
      h->mf_level_sync = kind == 9 || kind == 10;   // 10: with a grid barrier per level (mmpgo_solver_stage_times)
      h->mf_force_dep = kind == 11;          // 11: with per-supernode dependencies whatever the automatic choice is
      // one cold translation solve on scratch (rhs = whatever recover_t left), fixed iteration count
      const double tol = h->opt.translation_solve_tol; const int mi = h->opt.translation_solve_max_iters;
      if (getenv("MMPGO_TS_ITERS")) { h->opt.translation_solve_tol = 0.0; h->opt.translation_solve_max_iters = atoi(getenv("MMPGO_TS_ITERS")); }
      h->ts_grid_override = getenv("MMPGO_TS_GRID") ? atoi(getenv("MMPGO_TS_GRID")) : 0;
      int rc = Dr::solve_t(h, h->xprop, allm, false);
      h->opt.translation_solve_tol = tol; h->opt.translation_solve_max_iters = mi; h->ts_grid_override = 0;
      h->mf_dry = false; h->mf_level_sync = false; h->mf_force_dep = false;
      if (rc) return rc;
      h->ctr.launches--;
    } else {
      set_error("unknown kernel kind");
      return MMPGO_ERR_ARG;
    }
    h->ctr.launches++;
    return 0;
  };
  for (int w = 0; w < 3; ++w) RC(one());
  CK(cudaEventRecord(e0, h->stream));
  for (int r = 0; r < reps; ++r) RC(one());
  CK(cudaEventRecord(e1, h->stream));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  *ms_avg = ms / (float)reps;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}
int driver_profile_pass(Handle *h, int kind, int reps, float *ms_avg) {
  if (!h->initialized) { set_error("initialize first"); return MMPGO_ERR_STATE; }
  if (reps <= 0) { set_error("reps must be positive"); return MMPGO_ERR_ARG; }
  return h->d == 2 ? profile_pass<2>(h, kind, reps, ms_avg) : profile_pass<3>(h, kind, reps, ms_avg);
}

// device-side solve statistics -> counters (solve_iters = node-iterations of the G00 PCG,
// reserved[0] = pose-iterations, the unit of its byte accounting)
int driver_sync_counters(Handle *h) {
  if (!h->graph_set) return 0;
  if (!h->d_ts_stats) return 0;
  unsigned long long st[4] = {0, 0, 0, 0};
  CK(cudaMemcpyAsync(st, h->d_ts_stats, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->ctr.solve_iters = (int64_t)st[0];
  h->ctr.reserved[0] = (int64_t)st[1];
  h->ctr.reserved[3] = (int64_t)st[2];
  h->ctr.reserved[4] = (int64_t)st[3];
  return 0;
}
int driver_reset_solve_stats(Handle *h) {
  if (!h->graph_set || !h->d_ts_stats) return 0;
  CK(cudaMemsetAsync(h->d_ts_stats, 0, 4 * sizeof(unsigned long long), h->stream));
  CK(cudaStreamSynchronize(h->stream));             // no PCG launch's copy of the old counts is in flight any more
  std::memset(h->h_ts_stats, 0, 4 * sizeof(unsigned long long));
  h->ts_unconv_seen = 0;
  return 0;
}

int driver_current_objective(Handle *h, double *f, double *g2) {
  if (!h->initialized) { set_error("initialize first"); return MMPGO_ERR_STATE; }
  // sum_a fobj_a = F (DPGOStar.cpp:719-722 vs DPGOHash.cpp:108-118); |grad F|^2 = sum_a |gradF_a|^2
  double sf = 0.0, sg = 0.0;
  for (const auto &s : h->st) { sf += s.fobj; sg += s.gradFnorm * s.gradFnorm; }
  *f = sf; *g2 = sg;
  return 0;
}

}  // namespace mmpgo
