// Host-side driver state of libmmpgo (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/mmpgo.h"
#include "mmpgo_kernels.cuh"
#include "mmpgo_mf.cuh"

namespace mmpgo {

// DPGOResult (C++/DPGO/include/DPGO/DPGO_types.h:204-322), the scalars only;
// iterates k and k-1 live on the device.
struct NodeState {
  bool updated = true;
  int iters = 0;
  int soft_restart_hits[2] = {0, 0};
  std::vector<int> oscillations;
  int num_oscillations = 0;
  double gamma = 0.0;
  double s_cur = 1.0, s_next = 1.0;
  double Fk[2] = {0.0, 0.0};
  double Gk = 0.0;
  double fobj = 0.0, fobj_prev = 0.0, f = 0.0;
  double fobjE = 0.0;
  double gradFnorm = 0.0;
  bool refined = false;
  int rescale_count = 0;        // Rescale::Dynamic (DPGO_types.h:277)
  int rescales = 0;
  int restarts = 0, tcg_iterations = 0, tnt_iterations = 0;
};

struct NodeInfo {
  int n0 = 0, n1 = 0, m0 = 0, m1 = 0;
  int64_t first_gid = 0;
  std::vector<int> inter_he;   // half-edge ids in inter_measurements() order
  bool dense = false;
};

template <typename T> struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
};

constexpr int RED_SLOTS = 4;
struct Handle {
  mmpgo_options opt;
  int d = 0;
  cudaStream_t stream = nullptr;
  bool graph_set = false, initialized = false;
  // graph / partition
  int64_t N = 0;
  int num_nodes = 0, node_begin = 0, node_end = 0, A = 0;
  std::vector<int> node_off;          // [A+1] own pose offsets
  std::vector<int64_t> own_gid;       // [NO] global id of own pose
  std::vector<int64_t> halo_gid;      // [NH] global id of halo pose
  std::vector<int> halo_owner;        // [NH] owning node
  int NO = 0, NH = 0, NP = 0;
  int64_t stage_lo = 0, stage_hi = 0;  // id range covering own + halo poses: the rows of a host iterate that travel
  std::vector<NodeInfo> info;
  std::vector<NodeState> st;
  int64_t n_intra_entries = 0, n_inter_he = 0, n_edges_owned = 0;
  // tiles
  int n_tiles = 0;
  std::vector<int> h_tile_node, h_node_tb, h_node_te;
  int *d_tile_node = nullptr, *d_tile_start = nullptr, *d_tile_cnt = nullptr;
  int *d_node_tb = nullptr, *d_node_te = nullptr, *d_node_off = nullptr;
  int *d_active = nullptr;            // [A] current mask
  int *d_active2 = nullptr;
  // static operators
  int *d_rowptr = nullptr, *d_col = nullptr;
  double *d_blk0 = nullptr;       // [nnz][d+1] row 0 of every off-diagonal block (G01 pass)
  double *d_blk = nullptr, *d_gdiag = nullptr, *d_dintra = nullptr, *d_dinter = nullptr;
  double *d_tnv = nullptr, *d_pinv = nullptr;
  double *d_a00 = nullptr, *d_d00 = nullptr;
  int *d_xrowptr = nullptr;
  InterRec *d_xrec = nullptr;
  int *d_eidx = nullptr;              // owned edges, struct-of-arrays: i, j, inter flag
  double *d_eval = nullptr;           // tau, kappa, t[3], R[9] per owned edge, field-major
  double *d_ginv = nullptr;
  long long *d_dense_off = nullptr;
  int max_dense_n0 = 0;
  bool any_dense = false, any_pcg = false;
  std::vector<int> dense_mask, pcg_mask;
  double *d_xstage = nullptr;         // global iterate in the reference layout (lazy)
  int64_t *d_pose_gid = nullptr;      // [NP] global id of own + halo poses
  // pose vectors (NP x PB)
  double *X[3] = {nullptr, nullptr, nullptr};   // rotating: Xk, Xkm1, Xak
  int ik = 0, ikm1 = 1, iak = 2;
  double *Xakh = nullptr, *Yex = nullptr, *xprop = nullptr, *xeval = nullptr;
  // own-sized vectors
  double *g[2] = {nullptr, nullptr}, *Df[2] = {nullptr, nullptr};
  int icur = 0;
  double *gex = nullptr, *Dfex = nullptr, *nab = nullptr, *grad = nullptr;
  double *cg_s = nullptr, *cg_r = nullptr, *cg_v = nullptr, *cg_p = nullptr, *cg_Hp = nullptr, *cg_Hs = nullptr, *tdot_prev = nullptr;
  // compact (NO x D)
  double *rhs_t = nullptr;
  // persistent translation solve (mmpgo_tsolve.cu)
  int *d_sell_ptr = nullptr, *d_ts_sync = nullptr;
  unsigned char *d_sell_pack = nullptr;
  int *d_ct_node = nullptr, *d_ct_start = nullptr, *d_ct_cnt = nullptr, *d_node_ctb = nullptr, *d_node_cte = nullptr;
  int n_ctiles = 0;
  double *ts_rec = nullptr, *ts_z = nullptr, *ts_z_base = nullptr;
  double *ts_partials = nullptr, *ts_nstate = nullptr;
  unsigned long long *d_ts_stats = nullptr;   // [0] node-iterations [1] pose-iterations [2] node solves stopped on max_iters [3] max iterations of a node
  unsigned long long *h_ts_stats = nullptr;   // pinned copy, refreshed after every PCG launch
  unsigned long long ts_unconv_seen = 0;       // h_ts_stats[2] already reported
  int ts_max_grid = 0, ts_grid_override = 0;
  // tile dealing of the copy-ring kernel (host-built table)
  std::vector<int> h_node_ctb, h_node_cte;
  int *d_cta_ptr = nullptr, *d_cta_tiles = nullptr, *d_node_parts = nullptr;
  int ts_plan_grid = -1, ts_plan_chunk = -1, ts_plan_max = 0;
  std::vector<int> ts_plan_stage;
  int ts_lite_max_tiles = 16;    // use k_tsolve_lite when a CTA gets at most this many CTA tiles
  // sparse direct translation solve (mmpgo_mf.cuh): factor of G00 of every node above dense_solve_max_n
  bool use_direct = false;
  MfDevice mf;
  int *d_mf_perm = nullptr;            // original own pose -> row of the permuted right-hand side (identity for dense nodes)
  std::vector<int> h_mf_perm;
  std::vector<int> mf_stage_jobs;      // per stage (forward stages, then backward): warp jobs, CTA jobs
  int mf_grid = 0;
  // RegularizedCholesky preconditioner: factor of G11 + (lambda_max / max_cond) I per node, d scalar rows per pose
  bool use_regchol = false;
  MfDevice mf11;
  int mf11_grid = 0;
  bool mf11_level_sync = true;
  int *d_mf11_perm = nullptr;            // scalar row (pose * d + r) -> position in the elimination order
  double *rhs11 = nullptr;               // [NO d][d] right-hand side in elimination order
  double *pre_buf = nullptr;             // [NO][PB] (G11 + reg)^{-1} r in pose-block layout (rotation rows)
  int64_t mf11_nnz = 0;
  std::vector<double> lambda_max;        // per local node
  bool mf_dry = false, mf_level_sync = false, mf_level_sync_auto = true, mf_force_dep = false;
  int64_t mf_nnz = 0, mf_entries = 0, mf_tasks = 0;
  int mf_height = 0, mf_supernodes = 0;
  int64_t dense_poses = 0;
  int ts_force_kernel = 0;       // PCG path: 0 automatic, 1 copy-ring kernel, 2 small-shard kernel (options.translation_solver)
  int ts_chunk = 8;              // consecutive CTA tiles dealt to one CTA at a time (ring kernel)
  bool ts_nores = false;         // small-shard kernel without shared-memory residency
  int ts_handoff = -1;           // nodes still iterating at which the ring kernel hands off (-1: max(2, nodes/16))
  int tsl_max_grid = 0;          // co-resident CTAs of k_tsolve_lite
  int64_t sell_entries = 0;
  std::vector<int> h_sell_ptr;      // ELLPACK slice offsets (host copy: staging size of k_tsolve_lite)
  int tsl_plan_tpc = -1, tsl_stage_bytes = 0, tsl_vec_off = -1, tsl_z_off = -1, tsl_dyn_bytes = 0;   // shared-memory plan for `tpc` tiles per CTA
  // per half-edge
  double *w_cur = nullptr, *w_prev = nullptr, *w_tmp = nullptr;
  double *resc = nullptr;            // Rescale::Dynamic: rescale s_e of the owning node's majoriser (ones at set_graph)
  int *d_pose_rec = nullptr;         // own pose -> index of its diagonal entry in the PCG tile records
  bool dynamic = false;
  // scalars
  double *d_node_scal2 = nullptr;
  double *d_partials = nullptr, *d_node_scal = nullptr, *d_coef = nullptr, *d_gamma = nullptr;
  double *d_block_partials = nullptr, *d_scalar = nullptr;
  double *h_pinned = nullptr;   // pinned staging (A*NS + A*MAXC + misc)
  double *d_slot = nullptr, *h_slot = nullptr;   // RED_SLOTS x [A][NS] per-node sums read back with one synchronisation
  int *h_pinned_i = nullptr;
  // AMM-PGO* global state (DPGOStar.cpp:126-213)
  double starF = 0.0, star_fobj = 0.0;
  int star_restarts = 0;
  // halo exchange
  int rank = 0, world = 1;
  std::vector<std::pair<int, int>> boundary_pairs;   // (own pose, remote node) per inter half-edge
  std::vector<int64_t> send_poses, recv_poses;       // per peer rank
  std::vector<int64_t> send_dbl, recv_dbl;
  int *d_send_idx = nullptr;
  int64_t n_send = 0;
  double *d_send = nullptr;
  // two-array exchange: per-peer chunks [array a | array b]
  std::vector<int64_t> send_dbl2, recv_dbl2;
  int *d_send2_a = nullptr, *d_send2_b = nullptr, *d_recv2_a = nullptr, *d_recv2_b = nullptr, *d_halo_row = nullptr;
  double *d_send2 = nullptr, *d_recv2 = nullptr;
  bool next_halo_current = false;   // the halo rows of X^{k+1} already hold the peers' final X^{k+1}
  mmpgo_exchange_fn exchange_fn = nullptr;
  mmpgo_allreduce_fn allreduce_fn = nullptr;
  mmpgo_allreduce_dev_fn allreduce_dev_fn = nullptr;
  void *cb_user = nullptr;
  void *nccl_comm = nullptr;          // ncclComm_t owned by the handle (mmpgo_nccl_init); replaces the callbacks
  int64_t halo_exchanges = 0, allreduces = 0;
  mmpgo_counters ctr;
  std::vector<void *> allocs;
};

int driver_set_graph(Handle *h, int d, int64_t N, int num_nodes, int nb, int ne, int64_t E, const int32_t *ei,
                     const int32_t *ej, const double *R, const double *t, const double *kappa,
                     const double *tau);
int driver_initialize(Handle *h, const double *X, int64_t ldx);
int driver_update(Handle *h);
int driver_iterate(Handle *h);
int driver_communicate(Handle *h);
int driver_get_poses(Handle *h, double *X, int64_t ldx);
int driver_get_weights(Handle *h, int node, double *w, int64_t cap, int64_t *count);
void plan_halo_pair(int world, const int64_t *send_poses, const int64_t *recv_poses, int64_t n_own, int32_t *sa,
                    int32_t *sb, int32_t *ra, int32_t *rb, int32_t *hrow);
int driver_evaluate_f(Handle *h, const double *X, int64_t ldx, double *f);
int driver_evaluate_grad(Handle *h, const double *X, int64_t ldx, double *G, int64_t ldg);
int driver_current_objective(Handle *h, double *f, double *g2);
int driver_translation_solve(Handle *h, const double *rhs, double *t);
int driver_profile_pass(Handle *h, int kind, int reps, float *ms_avg);
int driver_sync_counters(Handle *h);
int driver_reset_solve_stats(Handle *h);
int driver_set_sharding(Handle *h, int rank, int world, const int32_t *rank_node_begin, mmpgo_exchange_fn ex,
                        mmpgo_allreduce_fn ar, void *user);
void driver_free(Handle *h);
int plan_halo(int64_t N, int num_nodes, int64_t E, const int32_t *ei, const int32_t *ej, int world,
              const int32_t *rnb, int rank, int64_t *sc, int64_t *rcn, int64_t *sg, int64_t scap, int64_t *rg,
              int64_t rcap);
void set_error(const std::string &s);
int nccl_unique_id(void *id);
int nccl_init(Handle *h, const void *id);
void nccl_destroy(Handle *h);
int nccl_exchange(Handle *h, const double *send, const int64_t *send_counts, double *recv, const int64_t *recv_counts);
int nccl_allreduce(Handle *h, double *vals_dev, int n);

}  // namespace mmpgo
