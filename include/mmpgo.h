/* mmpgo.h -- C ABI of the B200 MM-PGO / AMM-PGO* / AMM-PGO# iteration.
 *
 * One handle = the robot nodes [node_begin, node_end) of one pose graph, bound
 * to one CUDA device.  All pointers are HOST pointers borrowed for the call
 * unless the name ends in _dev.  Every function returns 0 on success and a
 * negative mmpgo_status otherwise (the reference's drivers return int 0 / -1,
 * C++/DPGO/include/DPGO/DPGOHash.h:20-28); no exceptions cross the boundary.
 * A handle is not thread-safe; different handles may be driven concurrently.
 *
 * Matrix conventions follow the reference: a global iterate X is
 * ((d+1) N) x d, column-major (Eigen::MatrixXd), rows [t_0..t_{N-1};
 * R-block_0 .. R-block_{N-1}] with d rows per rotation block
 * (C++/examples/dist_pgo.cpp:502-511, C++/DPGO/src/DPGOStar.cpp:541-547).
 *
 * Each entry point cites the reference interface it replaces.
 */
#ifndef MMPGO_H_
#define MMPGO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmpgo_handle_s *mmpgo_handle;

enum mmpgo_status {
  MMPGO_OK = 0,
  MMPGO_ERR_ARG = -1,        /* inconsistent sizes (reference: LOG(ERROR); return -1) */
  MMPGO_ERR_CUDA = -2,       /* a CUDA call failed; see mmpgo_last_error */
  MMPGO_ERR_STATE = -3,      /* call order violated (e.g. iterate before update) */
  MMPGO_ERR_UNSUPPORTED = -4,
  MMPGO_ERR_NOT_CONVERGED = -5  /* a PCG translation solve stopped on translation_solve_max_iters above the tolerance */
};

/* DPGO::Loss, C++/DPGO/include/DPGO/DPGO_types.h:67 */
enum mmpgo_loss { MMPGO_LOSS_NONE = 0, MMPGO_LOSS_HUBER = 1,
                  MMPGO_LOSS_GEMAN_MCCLURE = 2, MMPGO_LOSS_WELSCH = 3 };
/* DPGO::Scheme, DPGO_types.h:70-75 */
enum mmpgo_scheme { MMPGO_SCHEME_MM = 0, MMPGO_SCHEME_AMM = 1 };
/* DPGO::Preconditioner, DPGO_types.h:35-40, plus the per-pose d x d
 * block-Jacobi preconditioner of the device path. */
enum mmpgo_preconditioner { MMPGO_PRECON_NONE = 0, MMPGO_PRECON_JACOBI = 1,
                            MMPGO_PRECON_BLOCK_JACOBI = 2,
                            /* the reference's default (DPGO_types.h:155): sparse Cholesky factor of
                             * G11 + (lambda_max / max_cond) I per node (DPGOProblem.cpp:101-124), factored on the host
                             * at mmpgo_set_graph like G00 and applied by the same supernodal sweeps.  Exact and
                             * ~10x the bytes of a block-Jacobi application per tCG iteration. */
                            MMPGO_PRECON_REGULARIZED_CHOLESKY = 3 };
/* How G00 u = rhs is solved for nodes with more than dense_solve_max_n poses (the reference
 * keeps a CHOLMOD factor of G00, DPGOProblem.cpp:93).  DIRECT: sparse Cholesky (nested dissection,
 * multifrontal, factored at mmpgo_set_graph) applied by dependency-scheduled supernodal sweeps -- exact
 * like the reference's.  PCG: Jacobi-preconditioned CG to translation_solve_tol; _RING / _LITE pin
 * one of its two kernels (tests).  AUTO: DIRECT unless the factor would exceed the memory / setup
 * budget (thick 3-D nodes), then PCG. */
enum mmpgo_translation_solver { MMPGO_TSOLVE_AUTO = 0, MMPGO_TSOLVE_PCG = 1, MMPGO_TSOLVE_PCG_RING = 2,
                                MMPGO_TSOLVE_PCG_LITE = 3, MMPGO_TSOLVE_DIRECT = 4 };
/* DPGO::Rescale.  DYNAMIC (robust losses): every update() may replace the per-measurement rescale vector of a node by
 * clamp(1.25 omega, 0.01, 1) (DPGOProblem.cpp:301-321, 465-485) and rebuild the majoriser's inter-node diagonal
 * blocks, D, Q, T, N, V' from it (update_quadratic_mat, :751-840); G00 changes on its diagonal only, so the
 * translation solve runs the PCG kernels (the reference refactorises with CHOLMOD, :315). */
enum mmpgo_rescale { MMPGO_RESCALE_STATIC = 0, MMPGO_RESCALE_DYNAMIC = 1 };
/* Which driver class the handle emulates. */
enum mmpgo_algorithm { MMPGO_ALG_HASH = 0,   /* DPGOHash: AMM-PGO# / MM-PGO */
                       MMPGO_ALG_STAR = 1 }; /* DPGOStar: AMM-PGO*          */

/* DPGO::Options (DPGO_types.h:78-201); mmpgo_default_options() fills in the
 * values `dist_pgo` uses (C++/examples/dist_pgo.cpp:103-120) -- with ONE exception: the preconditioner defaults
 * to the per-pose block-Jacobi of this path's contract, not to the reference's RegularizedCholesky
 * (DPGO_types.h:155, which dist_pgo does not override).  Set preconditioner =
 * MMPGO_PRECON_REGULARIZED_CHOLESKY for the reference's own trajectory (exact, about twice the time per iteration). */
typedef struct mmpgo_options {
  int32_t algorithm;             /* mmpgo_algorithm */
  int32_t scheme;                /* mmpgo_scheme */
  int32_t loss;                  /* mmpgo_loss */
  int32_t preconditioner;        /* mmpgo_preconditioner; default BLOCK_JACOBI (the reference's is REGULARIZED_CHOLESKY) */
  double regularizer;            /* xi, 1e-11 */
  double loss_reg;               /* delta, 0.25 */
  double accepted_delta;         /* 5e-4 */
  double eta[2];                 /* 5e-4, 2.5e-2 */
  double psi, phi;               /* 1e-10, 1e-6 */
  int32_t max_soft_restart_hits[2];   /* 10, 25 */
  int32_t oscillation_cnt_period;     /* 15 */
  int32_t max_oscillations;           /* 12 */
  double grad_norm_tol;               /* 1e-3 */
  double preconditioned_grad_norm_tol;/* 1e-4 */
  double rel_func_decrease_tol;       /* 1e-6 */
  double stepsize_tol;                /* 1e-4 */
  int32_t max_iterations;             /* TNT outer iterations, 10 */
  int32_t max_iterations_accepted;    /* 1 */
  int32_t max_tCG_iterations;         /* 10000 */
  double STPCG_kappa, STPCG_theta;    /* 0.05, 0.9 */
  /* device-path knobs (no reference counterpart) */
  int32_t dense_solve_max_n;     /* nodes with n0 <= this use a dense G00^{-1}; else PCG */
  double translation_solve_tol;  /* relative residual of the G00 PCG, 1e-12 */
  int32_t translation_solve_max_iters;
  int32_t device;                /* CUDA device ordinal */
  int32_t translation_solver;    /* mmpgo_translation_solver; nodes above dense_solve_max_n */
  int32_t rescale;               /* mmpgo_rescale: DPGO::Rescale (DPGO_types.h:42-46, Options::rescale :128); dist_pgo
                                    pins Static (dist_pgo.cpp:105), which is the default here */
  int32_t max_rescale_count;     /* 5 (DPGO_types.h:131) */
  double reg_Cholesky_precon_max_condition_number;   /* 1e6 (DPGO_types.h:159) */
  int32_t reserved[2];
} mmpgo_options;

/* DPGOResult scalars a caller of results() reads (DPGO_types.h:204-322). */
typedef struct mmpgo_node_scalars {
  double fobj, f, Gk, gradFnorm, Fk[2], s, s_next, gamma;
  int32_t iters, soft_restart_hits[2], num_oscillations;
  int32_t refined, restarts, tcg_iterations, tnt_iterations;
  int32_t n0, n1, m0, m1;
  int32_t translation_solve_iters; /* PCG solver only: most iterations one node took in a solve since the last
                                      mmpgo_reset_counters (handle-wide, as of the last host synchronisation); 0 = direct */
  int32_t reserved;              /* Rescale::Dynamic: how often the node's rescale vector was replaced */
} mmpgo_node_scalars;

/* Kernel launch / byte accounting used by bench.py (gpu_launches, roofline). */
typedef struct mmpgo_counters {
  int64_t launches;               /* kernels launched by this handle */
  int64_t intra_passes;           /* K2 block-CSR passes */
  int64_t inter_passes;           /* K1 inter-edge passes */
  int64_t prox_passes;            /* K3 fused extrapolate+proximal+projection */
  int64_t solve_calls, solve_iters;   /* K2b G00 solves / PCG iterations */
  int64_t tcg_iterations, tnt_iterations;
  int64_t vector_passes;
  int64_t reserved[7];            /* [0] pose-iterations of the G00 PCG, [1] solves served by the small-shard PCG kernel,
                                     [2] solves served by the sparse direct kernel, [5] RegularizedCholesky preconditioner solves, [3] PCG node solves that stopped on
                                     translation_solve_max_iters above the tolerance, [4] most PCG iterations one node took */
} mmpgo_counters;

const char *mmpgo_version(void);
const char *mmpgo_last_error(void);
void mmpgo_default_options(mmpgo_options *opts);

/* Replaces the DPGOHash / DPGOStar constructors (DPGOHash.h:18, DPGOStar.h:16-19). */
int mmpgo_create(const mmpgo_options *opts, mmpgo_handle *out);
int mmpgo_destroy(mmpgo_handle h);

/* Replaces DPGO::read_g2o's partition (C++/DPGO/src/DPGO_utils.cpp:140-202),
 * generate_data_info (:326-438) and the DPGOProblem constructor
 * (C++/DPGO/src/DPGOProblem.cpp:11-125).  Edges carry GLOBAL pose ids; the
 * contiguous id-range partition into num_nodes robot nodes is the reference's.
 * R is row-major d x d per edge, t has d entries per edge. */
int mmpgo_set_graph(mmpgo_handle h, int32_t d, int64_t num_poses, int32_t num_nodes,
                    int32_t node_begin, int32_t node_end, int64_t num_edges,
                    const int32_t *edge_i, const int32_t *edge_j, const double *R,
                    const double *t, const double *kappa, const double *tau);

/* DPGOHash::initialize (DPGOHash.cpp:20-43) / DPGOStar::initialize
 * (DPGOStar.cpp:109-124) for all local nodes; X is the GLOBAL iterate, ldx its
 * leading dimension (>= (d+1) N).  Neighbour copies are taken from X too
 * (what DPGO::communicate does at dist_pgo.cpp:446). */
int mmpgo_initialize(mmpgo_handle h, const double *X, int64_t ldx);
/* DPGOHash::update (DPGOHash.cpp:84-228) / DPGOStar::update (:225-231). */
int mmpgo_update(mmpgo_handle h);
/* DPGOHash::iterate (DPGOHash.cpp:583-628) / DPGOStar::iterate (:126-213)
 * for all local nodes, batched. */
int mmpgo_iterate(mmpgo_handle h);
/* DPGOHash::communicate (DPGOHash.h:28-86) / DPGOStar::communicate (:215-223):
 * publishes the new own poses.  With remote neighbours the caller exchanges
 * the packed boundary buffers between begin and end (see halo API below). */
int mmpgo_communicate(mmpgo_handle h);

/* results().Xk (DPGO_types.h:208): writes the rows of the local nodes' own
 * poses into a GLOBAL-layout X. */
int mmpgo_get_poses(mmpgo_handle h, double *X, int64_t ldx);
int mmpgo_get_node_scalars(mmpgo_handle h, int32_t node, mmpgo_node_scalars *out);
/* DPGOProblem::evaluate_E's DiagReg (DPGOProblem.cpp:647-675): IRLS weight of
 * every inter-node measurement of `node`, in the order of the node's
 * inter_measurements() list.  *count receives m1. */
int mmpgo_get_weights(mmpgo_handle h, int32_t node, double *w, int64_t capacity,
                      int64_t *count);
/* DPGOStar::evaluate_f (DPGOStar.cpp:713-761) restricted to the edges owned
 * by the local nodes (each inter-node edge is owned by the node of its i
 * endpoint); summing over handles gives F.  X is a full GLOBAL iterate. */
int mmpgo_evaluate_f(mmpgo_handle h, const double *X, int64_t ldx, double *fobj);
/* DPGOStar::evaluate_grad (DPGOStar.cpp:763-829): Riemannian gradient of the global
 * objective at the GLOBAL iterate X (robust weights evaluated at X; rotation rows projected
 * onto the tangent space of SO(d)^n, translation rows Euclidean).  G has the layout of X
 * (column-major ((d+1)N) x d, leading dimension ldg); only the rows of the poses owned by the
 * local nodes are written (all rows for a single handle).  Does not touch the solver state. */
int mmpgo_evaluate_grad(mmpgo_handle h, const double *X, int64_t ldx, double *G, int64_t ldg);
/* The linear solve inside DPGOProblem::recover_translations (DPGOProblem.h:275-294, L_.solve):
 * t = -G00^{-1} rhs for every local node, through the handle's solve path (dense inverse, sparse
 * Cholesky sweeps or PCG, see mmpgo_translation_solver).  rhs, t: [own poses][d] row-major, own poses
 * of the local nodes in ascending global id.  Needs mmpgo_set_graph only; uses scratch. */
int mmpgo_translation_solve(mmpgo_handle h, const double *rhs, double *t);
/* Objective of the CURRENT device iterate, no host<->device pose traffic
 * (what dist_pgo logs each iteration, dist_pgo.cpp:523-530). */
int mmpgo_current_objective(mmpgo_handle h, double *fobj, double *grad_sqnorm);

/* ---- multi-GPU: robot nodes sharded over ranks, one handle per rank ----------
 * Only the boundary poses of inter-node loop closures move between GPUs
 * (DPGOHash::receive wire format, DPGOHash.cpp:45-82: one (d+1) x d block per
 * pose, ascending pose id per peer), plus a scalar all-reduce of the objective /
 * restart test for AMM-PGO* (DPGOStar.cpp:147-171).  The transport is supplied
 * by the caller as two callbacks (NCCL through torch.distributed in this repo):
 *   exchange(user, send_dev, send_counts, recv_dev, recv_counts): all-to-all of
 *     pose blocks; counts are in doubles per peer rank.  The call is ordered on the
 *     handle's stream (mmpgo_stream): the send buffer is produced by work already
 *     enqueued there, and the library reads recv_dev from kernels it enqueues there
 *     afterwards -- the transport must enqueue on, or synchronise with, that stream
 *     (no host synchronisation is required);
 *   allreduce(user, vals, n): in-place sum of n host doubles over all ranks.
 * rank_node_begin has world+1 entries: rank r owns nodes [begin[r], begin[r+1]).
 * Call after mmpgo_set_graph and before mmpgo_initialize. */
typedef int (*mmpgo_exchange_fn)(void *user, const void *send_dev, const int64_t *send_counts,
                                 void *recv_dev, const int64_t *recv_counts);
typedef int (*mmpgo_allreduce_fn)(void *user, double *vals, int32_t n);
int mmpgo_set_sharding(mmpgo_handle h, int32_t rank, int32_t world_size,
                       const int32_t *rank_node_begin, mmpgo_exchange_fn exchange,
                       mmpgo_allreduce_fn allreduce, void *user);
/* Optional: in-place sum over all ranks of n DEVICE doubles, ordered on the handle's stream
 * (ncclAllReduce).  When set, the AMM-PGO* scalars are reduced on the device and read back once. */
typedef int (*mmpgo_allreduce_dev_fn)(void *user, void *vals_dev, int32_t n);
int mmpgo_set_device_allreduce(mmpgo_handle h, mmpgo_allreduce_dev_fn allreduce_dev);
/* NCCL transport driven by the library itself (what mmpgo_communicate(handle, ncclComm_t) of SURVEY.md section 8b
 * stands for): rank 0 obtains a 128-byte ncclUniqueId, the caller distributes it to all ranks (MPI, a file,
 * torch.distributed ...), every rank calls mmpgo_nccl_init after mmpgo_set_sharding (whose callbacks may then
 * be NULL).  From then on the halo exchange is a grouped ncclSend / ncclRecv with the ranks that share an edge and
 * the AMM-PGO* scalars are one ncclAllReduce, all enqueued on the handle's stream from C++.  libnccl.so.2 is
 * resolved with dlopen at the first call; the communicator is destroyed with the handle. */
int mmpgo_nccl_unique_id(void *id128);
int mmpgo_nccl_init(mmpgo_handle h, const void *id128);
/* per-peer number of boundary poses sent / received each exchange (length world_size) */
int mmpgo_halo_counts(mmpgo_handle h, int64_t *send_poses, int64_t *recv_poses);
/* Host-only (no CUDA) restatement of the exchange plan for rank `rank`: the global ids
 * of the own poses sent to each peer (`sent_`, DPGO_utils.cpp:426-433) and of the remote
 * poses received from each peer (`recv_`, :435), concatenated in rank order, ascending
 * id per peer.  gids buffers may be NULL to query the counts only. */
int mmpgo_plan_halo(int64_t num_poses, int32_t num_nodes, int64_t num_edges, const int32_t *edge_i,
                    const int32_t *edge_j, int32_t world_size, const int32_t *rank_node_begin,
                    int32_t rank, int64_t *send_counts, int64_t *recv_counts, int64_t *send_gids,
                    int64_t send_capacity, int64_t *recv_gids, int64_t recv_capacity);

/* Host-only: index maps of the TWO-array halo exchange of AMM-PGO* (the boundary poses of
 * X^{k+1/2} and X^{k+1}, whose global objectives DPGOStar::iterate evaluates, DPGOStar.cpp:147-159,
 * travel in one all-to-all).  Per peer the wire chunk is [poses of array a | poses of array b].
 * send_a/send_b[i]: slot (in poses) of the i-th boundary pose of the send list in the send buffer;
 * recv_a/recv_b[k], halo_row[k]: slot of the k-th halo pose in the receive buffer and its row in the
 * pose arrays (n_own + k).  Lengths: sum(send_poses) and sum(recv_poses). */
int mmpgo_plan_halo_pair(int32_t world_size, const int64_t *send_poses, const int64_t *recv_poses, int64_t n_own,
                         int32_t *send_a, int32_t *send_b, int32_t *recv_a, int32_t *recv_b, int32_t *halo_row);

/* AMM-PGO* master-node scalars: F (the running average, DPGOStar.cpp:210),
 * the last accepted global objective and the number of global restarts. */
int mmpgo_star_objective(mmpgo_handle h, double *F, double *fobj, int32_t *restarts);
/* sizes[8] = {own poses, halo poses, block-CSR entries, inter half-edges,
 * owned edges, tiles, local nodes, d} */
int mmpgo_graph_sizes(mmpgo_handle h, int64_t *sizes);
/* info[8] = {translation solver in use for the large nodes (mmpgo_translation_solver; 0 = every node
 * is dense), nnz(L) of the sparse factor, stored factor entries (both copies), separator-tree height,
 * supernodes, jobs per solve, persistent CTAs, poses solved by the dense inverse} */
int mmpgo_solver_info(mmpgo_handle h, int64_t *info);
/* RegularizedCholesky preconditioner of a local node (DPGOProblem.cpp:101-124): the estimate of the largest
 * eigenvalue of G11 the regulariser lambda_max / reg_Cholesky_precon_max_condition_number was built from (the
 * reference asks Spectra for it to 1e-4), and the entries of the Cholesky factor of all local nodes' G11 + reg I.
 * MMPGO_ERR_STATE when the handle runs another preconditioner. */
int mmpgo_preconditioner_info(mmpgo_handle h, int32_t node, double *lambda_max, int64_t *factor_nnz);
/* [lo, hi): the global pose ids whose rows of a host iterate are copied to the device by mmpgo_initialize /
 * mmpgo_evaluate_f / mmpgo_evaluate_grad: the own poses of the local nodes and their remote neighbours. */
int mmpgo_stage_range(mmpgo_handle h, int64_t *lo, int64_t *hi);
/* Sparse direct solve, measurement: duration in microseconds of every stage of the LAST solve (forward
 * stages by height, then backward stages by depth; device timer of CTA 0 at the grid barriers) with the
 * warp jobs and CTA jobs of the stage.  *count receives the number of stages (0 without a sparse factor);
 * us may be NULL to query it. */
int mmpgo_solver_stage_times(mmpgo_handle h, double *us, int32_t *warp_jobs, int32_t *cta_jobs, int32_t capacity,
                             int32_t *count);

/* project_to_SO3n / project_to_SO2n (C++/DPGO/include/DPGO/DPGO_utils.h:515-565; kernels
 * C++/DPGO/src/internal/project_to_SOd.cpp:27-33,121-196): the polar projection of n row-major
 * d x d host blocks onto SO(d), on `device`; the operation the fused proximal kernel applies to
 * every pose.  Needs no handle. */
int mmpgo_project_to_sodn(int32_t d, int64_t n, const double *A, double *U, int32_t device);

/* Host-only (no CUDA): the sparse direct solver behind the translation solve (the reference's
 * CHOLMOD factor L_ = chol(G00), DPGOProblem.cpp:93, and its L_.solve call sites) on one SPD
 * matrix given as CSR (full symmetric pattern incl. the diagonal; `block` consecutive rows form
 * one graph vertex): nested-dissection multifrontal Cholesky, then the host restatement of the
 * device sweeps (same blocks, same order of every sum).  rhs, x: [n][nrhs] row-major.  stats[8]
 * (may be NULL) = {nnz(L), tree height, supernodes, factor flops, forward jobs, backward jobs,
 * largest CTA-served front, update rows}.  Used by the CPU tests of the factorisation. */
int mmpgo_mf_host_solve(int32_t n, const int32_t *ptr, const int32_t *col, const double *val, int32_t block,
                        int32_t leaf, int32_t nrhs, const double *rhs, double *x, int64_t *stats);

/* Measurement hook: average device time (CUDA events on the handle's stream) of
 * `reps` back-to-back launches of one hot kernel on the current iterate.
 * kind: 0 K2 evaluate, 1 K2 gradient, 2 K1 inter-edge pass, 3 K3 fused proximal,
 * 4 edge-parallel objective, 6 K2 Hessian-vector, 7 K2 G01 pass, 8 one K2b translation solve (one persistent
 * launch); sparse direct solve only: 9 its levels and barriers without the jobs, 10 the solve with a grid barrier
 * after every level of the separator tree (what mmpgo_solver_stage_times reports), 11 the solve with per-supernode
 * dependencies instead (the library picks one of the two by the number of jobs per resident warp). */
int mmpgo_profile_pass(mmpgo_handle h, int32_t kind, int32_t reps, float *ms_avg);
int mmpgo_get_counters(mmpgo_handle h, mmpgo_counters *out);
int mmpgo_reset_counters(mmpgo_handle h);
int mmpgo_synchronize(mmpgo_handle h);
/* the CUDA stream all kernels of this handle are launched on (cudaStream_t) */
void *mmpgo_stream(mmpgo_handle h);

#ifdef __cplusplus
}
#endif
#endif /* MMPGO_H_ */
