// extern "C" surface of libmmpgo; see include/mmpgo.h for the contract.
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "mmpgo_driver.cuh"
#include "mmpgo_mf.cuh"

namespace mmpgo {
static thread_local std::string g_err;
void set_error(const std::string &s) { g_err = s; }
}  // namespace mmpgo

using mmpgo::Handle;

struct mmpgo_handle_s {
  Handle h;
};

#define H_OR_FAIL(hh)                                   \
  if (!(hh)) {                                          \
    mmpgo::set_error("null handle");                    \
    return MMPGO_ERR_ARG;                               \
  }                                                     \
  Handle *h = &(hh)->h;                                 \
  if (cudaSetDevice(h->opt.device) != cudaSuccess) {    \
    mmpgo::set_error("cudaSetDevice failed");           \
    return MMPGO_ERR_CUDA;                              \
  }

extern "C" {

const char *mmpgo_version(void) { return "mmpgo-b200 0.1 (sm_100a, fp64)"; }
const char *mmpgo_last_error(void) { return mmpgo::g_err.c_str(); }

void mmpgo_default_options(mmpgo_options *o) {
  std::memset(o, 0, sizeof(*o));
  o->algorithm = MMPGO_ALG_HASH;
  o->scheme = MMPGO_SCHEME_AMM;
  o->loss = MMPGO_LOSS_NONE;
  o->preconditioner = MMPGO_PRECON_BLOCK_JACOBI;
  o->regularizer = 1e-11;       // dist_pgo.cpp:120
  o->loss_reg = 0.25;           // :107
  o->accepted_delta = 5e-4;
  o->eta[0] = 5e-4; o->eta[1] = 2.5e-2;
  o->psi = 1e-10; o->phi = 1e-6;
  o->max_soft_restart_hits[0] = 10; o->max_soft_restart_hits[1] = 25;
  o->oscillation_cnt_period = 15;
  o->max_oscillations = 12;
  o->grad_norm_tol = 1e-3;
  o->preconditioned_grad_norm_tol = 1e-4;
  o->rel_func_decrease_tol = 1e-6;
  o->stepsize_tol = 1e-4;
  o->max_iterations = 10;
  o->max_iterations_accepted = 1;
  o->max_tCG_iterations = 10000;
  o->STPCG_kappa = 0.05; o->STPCG_theta = 0.9;
  o->dense_solve_max_n = 2048;
  o->translation_solve_tol = 1e-12;
  o->translation_solve_max_iters = 4000;
  o->device = 0;
  o->translation_solver = MMPGO_TSOLVE_AUTO;
  o->rescale = MMPGO_RESCALE_STATIC;        // dist_pgo.cpp:105
  o->max_rescale_count = 5;                // DPGO_types.h:131
  o->reg_Cholesky_precon_max_condition_number = 1e6;   // DPGO_types.h:159
}

int mmpgo_create(const mmpgo_options *opts, mmpgo_handle *out) {
  if (!opts || !out) { mmpgo::set_error("null argument"); return MMPGO_ERR_ARG; }
  if (opts->loss < 0 || opts->loss > 3 || opts->preconditioner < 0 || opts->preconditioner > 3 ||
      opts->algorithm < 0 || opts->algorithm > 1 || opts->scheme < 0 || opts->scheme > 1 ||
      opts->translation_solver < 0 || opts->translation_solver > 4 || opts->rescale < 0 || opts->rescale > 1) {
    mmpgo::set_error("invalid enum value in options");
    return MMPGO_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    mmpgo::set_error("no CUDA device: libmmpgo has no CPU fallback");
    return MMPGO_ERR_CUDA;
  }
  if (opts->device < 0 || opts->device >= ndev) { mmpgo::set_error("bad device ordinal"); return MMPGO_ERR_ARG; }
  if (cudaSetDevice(opts->device) != cudaSuccess) { mmpgo::set_error("cudaSetDevice failed"); return MMPGO_ERR_CUDA; }
  mmpgo_handle_s *hs = new (std::nothrow) mmpgo_handle_s();
  if (!hs) { mmpgo::set_error("out of host memory"); return MMPGO_ERR_ARG; }
  hs->h.opt = *opts;
  std::memset(&hs->h.ctr, 0, sizeof(hs->h.ctr));
  if (cudaStreamCreateWithFlags(&hs->h.stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete hs;
    mmpgo::set_error("cudaStreamCreate failed");
    return MMPGO_ERR_CUDA;
  }
  *out = hs;
  return MMPGO_OK;
}

int mmpgo_destroy(mmpgo_handle hh) {
  if (!hh) return MMPGO_OK;
  cudaSetDevice(hh->h.opt.device);
  cudaStreamSynchronize(hh->h.stream);
  mmpgo::driver_free(&hh->h);
  cudaStreamDestroy(hh->h.stream);
  delete hh;
  return MMPGO_OK;
}

int mmpgo_set_graph(mmpgo_handle hh, int32_t d, int64_t num_poses, int32_t num_nodes, int32_t node_begin,
                    int32_t node_end, int64_t num_edges, const int32_t *edge_i, const int32_t *edge_j,
                    const double *R, const double *t, const double *kappa, const double *tau) {
  H_OR_FAIL(hh);
  if (!edge_i || !edge_j || !R || !t || !kappa || !tau) { mmpgo::set_error("null edge array"); return MMPGO_ERR_ARG; }
  try {
    return mmpgo::driver_set_graph(h, d, num_poses, num_nodes, node_begin, node_end, num_edges, edge_i, edge_j, R,
                                   t, kappa, tau);
  } catch (const std::exception &e) {
    mmpgo::set_error(e.what());
    return MMPGO_ERR_ARG;
  }
}

#define GUARDED(call)                              \
  try {                                            \
    return (call);                                 \
  } catch (const std::exception &e) {              \
    mmpgo::set_error(e.what());                    \
    return MMPGO_ERR_ARG;                          \
  }

int mmpgo_initialize(mmpgo_handle hh, const double *X, int64_t ldx) {
  H_OR_FAIL(hh);
  if (!X) { mmpgo::set_error("null X"); return MMPGO_ERR_ARG; }
  GUARDED(mmpgo::driver_initialize(h, X, ldx));
}
int mmpgo_update(mmpgo_handle hh) { H_OR_FAIL(hh); GUARDED(mmpgo::driver_update(h)); }
int mmpgo_iterate(mmpgo_handle hh) { H_OR_FAIL(hh); GUARDED(mmpgo::driver_iterate(h)); }
int mmpgo_communicate(mmpgo_handle hh) { H_OR_FAIL(hh); GUARDED(mmpgo::driver_communicate(h)); }

int mmpgo_get_poses(mmpgo_handle hh, double *X, int64_t ldx) {
  H_OR_FAIL(hh);
  if (!X || ldx < (int64_t)(h->d + 1) * h->N) { mmpgo::set_error("bad X / ldx"); return MMPGO_ERR_ARG; }
  GUARDED(mmpgo::driver_get_poses(h, X, ldx));
}

int mmpgo_get_node_scalars(mmpgo_handle hh, int32_t node, mmpgo_node_scalars *out) {
  H_OR_FAIL(hh);
  if (!out || !h->graph_set || node < h->node_begin || node >= h->node_end) {
    mmpgo::set_error("node not local");
    return MMPGO_ERR_ARG;
  }
  const mmpgo::NodeState &s = h->st[node - h->node_begin];
  const mmpgo::NodeInfo &ni = h->info[node - h->node_begin];
  std::memset(out, 0, sizeof(*out));
  out->fobj = s.fobj; out->f = s.f; out->Gk = s.Gk; out->gradFnorm = s.gradFnorm;
  out->Fk[0] = s.Fk[0]; out->Fk[1] = s.Fk[1];
  out->s = s.s_cur; out->s_next = s.s_next; out->gamma = s.gamma;
  out->iters = s.iters;
  out->soft_restart_hits[0] = s.soft_restart_hits[0]; out->soft_restart_hits[1] = s.soft_restart_hits[1];
  out->num_oscillations = s.num_oscillations;
  out->refined = s.refined ? 1 : 0; out->restarts = s.restarts;
  out->tcg_iterations = s.tcg_iterations; out->tnt_iterations = s.tnt_iterations;
  out->n0 = ni.n0; out->n1 = ni.n1; out->m0 = ni.m0; out->m1 = ni.m1;
  out->reserved = s.rescales;
  out->translation_solve_iters = h->h_ts_stats ? (int32_t)h->h_ts_stats[3] : 0;
  return MMPGO_OK;
}

int mmpgo_get_weights(mmpgo_handle hh, int32_t node, double *w, int64_t capacity, int64_t *count) {
  H_OR_FAIL(hh);
  if (!count || !h->graph_set) { mmpgo::set_error("bad argument"); return MMPGO_ERR_ARG; }
  GUARDED(mmpgo::driver_get_weights(h, node, w, capacity, count));
}

int mmpgo_evaluate_f(mmpgo_handle hh, const double *X, int64_t ldx, double *fobj) {
  H_OR_FAIL(hh);
  if (!X || !fobj || ldx < (int64_t)(h->d + 1) * h->N) { mmpgo::set_error("bad X / ldx"); return MMPGO_ERR_ARG; }
  GUARDED(mmpgo::driver_evaluate_f(h, X, ldx, fobj));
}

int mmpgo_evaluate_grad(mmpgo_handle hh, const double *X, int64_t ldx, double *G, int64_t ldg) {
  H_OR_FAIL(hh);
  if (!X || !G || ldx < (int64_t)(h->d + 1) * h->N || ldg < (int64_t)(h->d + 1) * h->N) {
    mmpgo::set_error("bad X / G / leading dimension");
    return MMPGO_ERR_ARG;
  }
  GUARDED(mmpgo::driver_evaluate_grad(h, X, ldx, G, ldg));
}

int mmpgo_translation_solve(mmpgo_handle hh, const double *rhs, double *t) {
  H_OR_FAIL(hh);
  if (!rhs || !t) { mmpgo::set_error("null argument"); return MMPGO_ERR_ARG; }
  GUARDED(mmpgo::driver_translation_solve(h, rhs, t));
}

int mmpgo_current_objective(mmpgo_handle hh, double *fobj, double *grad_sqnorm) {
  H_OR_FAIL(hh);
  if (!fobj || !grad_sqnorm) { mmpgo::set_error("null output"); return MMPGO_ERR_ARG; }
  GUARDED(mmpgo::driver_current_objective(h, fobj, grad_sqnorm));
}

int mmpgo_star_objective(mmpgo_handle hh, double *F, double *fobj, int32_t *restarts) {
  H_OR_FAIL(hh);
  if (F) *F = h->starF;
  if (fobj) *fobj = h->star_fobj;
  if (restarts) *restarts = h->star_restarts;
  return MMPGO_OK;
}

int mmpgo_set_sharding(mmpgo_handle hh, int32_t rank, int32_t world_size, const int32_t *rank_node_begin,
                       mmpgo_exchange_fn exchange, mmpgo_allreduce_fn allreduce, void *user) {
  H_OR_FAIL(hh);
  GUARDED(mmpgo::driver_set_sharding(h, rank, world_size, rank_node_begin, exchange, allreduce, user));
}
int mmpgo_set_device_allreduce(mmpgo_handle hh, mmpgo_allreduce_dev_fn fn) {
  H_OR_FAIL(hh);
  h->allreduce_dev_fn = fn;
  return MMPGO_OK;
}
int mmpgo_nccl_unique_id(void *id128) {
  if (!id128) { mmpgo::set_error("null argument"); return MMPGO_ERR_ARG; }
  GUARDED(mmpgo::nccl_unique_id(id128));
}
int mmpgo_nccl_init(mmpgo_handle hh, const void *id128) {
  H_OR_FAIL(hh);
  if (!id128) { mmpgo::set_error("null argument"); return MMPGO_ERR_ARG; }
  GUARDED(mmpgo::nccl_init(h, id128));
}

int mmpgo_halo_counts(mmpgo_handle hh, int64_t *send_poses, int64_t *recv_poses) {
  H_OR_FAIL(hh);
  if (!send_poses || !recv_poses || (int)h->send_poses.size() != h->world) {
    mmpgo::set_error("set_sharding first");
    return MMPGO_ERR_STATE;
  }
  for (int q = 0; q < h->world; ++q) { send_poses[q] = h->send_poses[q]; recv_poses[q] = h->recv_poses[q]; }
  return MMPGO_OK;
}

int mmpgo_plan_halo(int64_t num_poses, int32_t num_nodes, int64_t num_edges, const int32_t *edge_i,
                    const int32_t *edge_j, int32_t world_size, const int32_t *rank_node_begin, int32_t rank,
                    int64_t *send_counts, int64_t *recv_counts, int64_t *send_gids, int64_t send_capacity,
                    int64_t *recv_gids, int64_t recv_capacity) {
  GUARDED(mmpgo::plan_halo(num_poses, num_nodes, num_edges, edge_i, edge_j, world_size, rank_node_begin, rank,
                           send_counts, recv_counts, send_gids, send_capacity, recv_gids, recv_capacity));
}

int mmpgo_plan_halo_pair(int32_t world_size, const int64_t *send_poses, const int64_t *recv_poses, int64_t n_own,
                         int32_t *send_a, int32_t *send_b, int32_t *recv_a, int32_t *recv_b, int32_t *halo_row) {
  if (world_size < 1 || !send_poses || !recv_poses || !send_a || !send_b || !recv_a || !recv_b || !halo_row) {
    mmpgo::set_error("bad argument");
    return MMPGO_ERR_ARG;
  }
  mmpgo::plan_halo_pair(world_size, send_poses, recv_poses, n_own, send_a, send_b, recv_a, recv_b, halo_row);
  return MMPGO_OK;
}

int mmpgo_profile_pass(mmpgo_handle hh, int32_t kind, int32_t reps, float *ms_avg) {
  H_OR_FAIL(hh);
  if (!ms_avg) { mmpgo::set_error("null output"); return MMPGO_ERR_ARG; }
  GUARDED(mmpgo::driver_profile_pass(h, kind, reps, ms_avg));
}

int mmpgo_get_counters(mmpgo_handle hh, mmpgo_counters *out) {
  H_OR_FAIL(hh);
  if (!out) { mmpgo::set_error("null output"); return MMPGO_ERR_ARG; }
  int rc = mmpgo::driver_sync_counters(h);
  if (rc) return rc;
  *out = h->ctr;
  return MMPGO_OK;
}
int mmpgo_reset_counters(mmpgo_handle hh) {
  H_OR_FAIL(hh);
  std::memset(&h->ctr, 0, sizeof(h->ctr));
  return mmpgo::driver_reset_solve_stats(h);
}
int mmpgo_synchronize(mmpgo_handle hh) {
  H_OR_FAIL(hh);
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) { mmpgo::set_error("stream sync failed"); return MMPGO_ERR_CUDA; }
  return MMPGO_OK;
}
void *mmpgo_stream(mmpgo_handle hh) { return hh ? (void *)hh->h.stream : nullptr; }

int mmpgo_graph_sizes(mmpgo_handle hh, int64_t *sizes /* [8] */) {
  H_OR_FAIL(hh);
  if (!sizes || !h->graph_set) { mmpgo::set_error("bad argument"); return MMPGO_ERR_ARG; }
  sizes[0] = h->NO; sizes[1] = h->NH; sizes[2] = h->n_intra_entries; sizes[3] = h->n_inter_he;
  sizes[4] = h->n_edges_owned; sizes[5] = h->n_tiles; sizes[6] = h->A; sizes[7] = h->d;
  return MMPGO_OK;
}

int mmpgo_preconditioner_info(mmpgo_handle hh, int32_t node, double *lambda_max, int64_t *factor_nnz) {
  H_OR_FAIL(hh);
  if (!h->graph_set || node < h->node_begin || node >= h->node_end) { mmpgo::set_error("node not local"); return MMPGO_ERR_ARG; }
  if (!h->use_regchol) { mmpgo::set_error("the handle does not run the RegularizedCholesky preconditioner"); return MMPGO_ERR_STATE; }
  if (lambda_max) *lambda_max = h->lambda_max[node - h->node_begin];
  if (factor_nnz) *factor_nnz = h->mf11_nnz;
  return MMPGO_OK;
}

int mmpgo_solver_info(mmpgo_handle hh, int64_t *info /* [8] */) {
  H_OR_FAIL(hh);
  if (!info || !h->graph_set) { mmpgo::set_error("bad argument"); return MMPGO_ERR_ARG; }
  info[0] = !h->any_pcg ? 0 : h->use_direct ? MMPGO_TSOLVE_DIRECT
            : h->ts_force_kernel == 1 ? MMPGO_TSOLVE_PCG_RING : h->ts_force_kernel == 2 ? MMPGO_TSOLVE_PCG_LITE : MMPGO_TSOLVE_PCG;
  info[1] = h->mf_nnz; info[2] = h->mf_entries; info[3] = h->mf_height; info[4] = h->mf_supernodes;
  info[5] = h->mf_tasks; info[6] = h->mf_grid; info[7] = h->dense_poses;
  return MMPGO_OK;
}

int mmpgo_solver_stage_times(mmpgo_handle hh, double *us, int32_t *warp_jobs, int32_t *cta_jobs, int32_t capacity, int32_t *count) {
  H_OR_FAIL(hh);
  if (!count) { mmpgo::set_error("null argument"); return MMPGO_ERR_ARG; }
  const int n = h->use_direct ? (int)h->mf_stage_jobs.size() / 2 : 0;
  *count = n;
  if (!us || n == 0) return MMPGO_OK;
  if (capacity < n) { mmpgo::set_error("buffer too small"); return MMPGO_ERR_ARG; }
  std::vector<unsigned long long> t((size_t)n + 1);
  if (cudaStreamSynchronize(h->stream) != cudaSuccess ||
      cudaMemcpy(t.data(), h->mf.stage_ns, sizeof(unsigned long long) * (n + 1), cudaMemcpyDeviceToHost) != cudaSuccess) {
    mmpgo::set_error("stage time read-back failed");
    return MMPGO_ERR_CUDA;
  }
  for (int i = 0; i < n; ++i) {
    us[i] = 1e-3 * (double)(t[i + 1] - t[i]);
    if (warp_jobs) warp_jobs[i] = h->mf_stage_jobs[2 * i];
    if (cta_jobs) cta_jobs[i] = h->mf_stage_jobs[2 * i + 1];
  }
  return MMPGO_OK;
}

int mmpgo_stage_range(mmpgo_handle hh, int64_t *lo, int64_t *hi) {
  H_OR_FAIL(hh);
  if (!lo || !hi || !h->graph_set) { mmpgo::set_error("bad argument"); return MMPGO_ERR_ARG; }
  *lo = h->stage_lo; *hi = h->stage_hi;
  return MMPGO_OK;
}

int mmpgo_project_to_sodn(int32_t d, int64_t n, const double *A, double *U, int32_t device) {
  if ((d != 2 && d != 3) || n < 0 || !A || !U) { mmpgo::set_error("bad argument"); return MMPGO_ERR_ARG; }
  if (n == 0) return MMPGO_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev || cudaSetDevice(device) != cudaSuccess) {
    mmpgo::set_error("no CUDA device: libmmpgo has no CPU fallback");
    return MMPGO_ERR_CUDA;
  }
  double *dA = nullptr, *dU = nullptr;
  const size_t bytes = (size_t)n * d * d * sizeof(double);
  int rc = MMPGO_OK;
  if (cudaMalloc(&dA, bytes) != cudaSuccess || cudaMalloc(&dU, bytes) != cudaSuccess ||
      cudaMemcpy(dA, A, bytes, cudaMemcpyHostToDevice) != cudaSuccess) rc = MMPGO_ERR_CUDA;
  if (!rc) {
    if (d == 2) mmpgo::launch_project_blocks<2>(n, dA, dU, nullptr);
    else mmpgo::launch_project_blocks<3>(n, dA, dU, nullptr);
    if (cudaMemcpy(U, dU, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) rc = MMPGO_ERR_CUDA;
  }
  if (rc) mmpgo::set_error(std::string("project_to_SOdn: ") + cudaGetErrorString(cudaGetLastError()));
  cudaFree(dA); cudaFree(dU);
  return rc;
}

int mmpgo_mf_host_solve(int32_t n, const int32_t *ptr, const int32_t *col, const double *val, int32_t block,
                        int32_t leaf, int32_t nrhs, const double *rhs, double *x, int64_t *stats) {
  if (n <= 0 || !ptr || !col || !val || block < 1 || n % block || nrhs < 1 || nrhs > 8 || !rhs || !x) {
    mmpgo::set_error("bad argument");
    return MMPGO_ERR_ARG;
  }
  try {
    mmpgo::MfMatrix A;
    A.n = n; A.ptr = ptr; A.col = col; A.val = val;
    mmpgo::MfFactor F;
    if (mmpgo::mf_factor({A}, block, leaf, false, &F) != 0) {
      mmpgo::set_error("matrix is not positive definite");
      return MMPGO_ERR_ARG;
    }
    mmpgo::mf_host_solve(F, nrhs, rhs, x);
    if (stats) {
      stats[0] = F.nnz; stats[1] = F.height; stats[2] = (int64_t)F.sn.size(); stats[3] = (int64_t)F.flops;
      stats[4] = (int64_t)F.wjobs[0].size(); stats[5] = (int64_t)F.wjobs[1].size();
      stats[6] = F.max_R; stats[7] = F.urows;
    }
    return MMPGO_OK;
  } catch (const std::exception &e) {
    mmpgo::set_error(e.what());
    return MMPGO_ERR_ARG;
  }
}

}  // extern "C"
