// Internal kernel interface of libmmpgo (sm_100a, FP64).
//
// Device data model (DESIGN.md "Data layout in HBM"):
//   * a pose is one (d+1) x d row-major block: row 0 = t_i, rows 1..d = Y_i = R_i^T
//     (the reference stores the same numbers as rows i and n0+d*i.. of X,
//     C++/DPGO/src/DPGO_utils.cpp:413-424); PB = (d+1)*d doubles.
//   * own poses of all local robot nodes are contiguous [0, NO); copies of
//     remote neighbours ("halo") follow at [NO, NP).
//   * the static majoriser G (DPGO_utils.cpp:1542-1641, 1964-2028) is a
//     block-CSR over own poses: (d+1)x(d+1) off-diagonal blocks per intra-node
//     half-edge + one packed symmetric diagonal block per pose.
//   * inter-node half-edges (one per own endpoint) are a CSR of 128-byte
//     records holding the raw measurement.
//   * work is cut into tiles of <= TILE poses that never straddle a node, so
//     every per-node scalar is a fixed-order sum of per-tile partials.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmpgo {

constexpr int TILE = 64;      // poses per tile / CTA
constexpr int NS = 8;         // scalar slots per tile
constexpr int MAXC = 8;       // per-node coefficient slots

struct Tiles {
  int n_tiles;
  const int *node;    // [n_tiles] local node of the tile
  const int *start;   // [n_tiles] first own pose
  const int *cnt;     // [n_tiles] poses in tile
  const int *active;  // [num_local_nodes] 0/1 mask (device) or nullptr = all
};

// ---- K2: block-CSR connection-Laplacian pass -------------------------------
enum GMode {
  G_EVAL = 0,     // s0 = sum x.(g + 1/2 G x)                     (evaluate_G, DPGOProblem.cpp:180-203)
  G_GRAD = 1,     // out = g + G x; s0 as above; s1 = |gradF|^2; s2 = x.Gx; s3 = x.g
  G_RHS_T = 2,    // rhs_t = g_t + G01 Y  (t rows out, Y rows in)   (DPGOProblem.h:289-290)
  G_REDGRAD = 3,  // nab = g_Y + (G x)_Y; grad = Proj(Y, nab); s0 = |grad|^2 (DPGOProblem.h:370-393); s1 = evaluate_G
  G_HV = 4        // Hp = Proj(Y, (G p)_Y - sym(nab Y^T) p_Y); s0=p.Hp s1=Hp.Hp s2=p.p (DPGOProblem.cpp:552-577)
};

struct GPassArgs {
  const int *rowptr;      // [NO+1]
  const int *col;         // [nnz] own pose index
  const double *blk;      // [nnz][(d+1)^2] off-diagonal blocks, row-major
  const double *blk0;     // [nnz][d+1] row 0 of every block (G_RHS_T reads only these)
  const double *diag;     // [NO][SYM] packed lower-triangular diagonal block
  const double *x;        // input pose-block vector (gathered)
  const double *g;        // additive vector (may be null)
  const double *xref;     // G_HV: the point Y (pose blocks); others unused
  const double *nab;      // G_HV: Euclidean gradient rows
  double *out;            // vector output (pose blocks or compact t rows for G_RHS_T)
  const int *out_perm;    // G_RHS_T: row of pose p in `out` (sparse direct solve: its elimination order); null = p
  double *out2;           // G_REDGRAD: grad
  double *partials;       // [n_tiles][NS]
};

// ---- K1: inter-node edge pass ----------------------------------------------
struct InterRec {         // one own-endpoint half-edge, 128 bytes
  int32_t other;          // pose index (own or halo) of the other endpoint
  int32_t own_is_i;       // 1: own pose is the edge's i endpoint
  double tau, kappa;
  double t[3];
  double R[9];            // row-major d x d
  double pad;
};
static_assert(sizeof(InterRec) == 128, "InterRec must be 128 bytes");

enum InterMode {
  I_TRIVIAL = 0,   // g = S Z, q(v) partials                       (DPGOProblem.cpp:269-287, 516-542)
  I_ROBUST = 1     // evaluate_E: weights, DfobjE rows, loss value (DPGOProblem.cpp:634-681)
};

struct InterArgs {
  const int *rowptr;          // [NO+1] half-edges per own pose
  const InterRec *rec;
  const double *xa;           // Z_k      (NP pose blocks)
  const double *xb;           // Z_{k-1}  (may be null)
  const double *dinter;       // [NO][SYM] sum of own diagonal blocks of inter edges
  const double *gamma;        // per-node extrapolation factor (null = 0)
  int use_diff;               // I_TRIVIAL: v = z - z_prev for q(v); I_ROBUST: add history terms
  int loss;
  double loss_reg, xi;
  const double *w_prev;       // previous IRLS weights per half-edge (I_ROBUST, use_diff)
  double *w_out;              // IRLS weights per half-edge (may be null)
  double *e_out;              // squared error per half-edge (may be null)
  double *g;                  // [NO] pose blocks: g
  double *yex;                // extrapolated own poses (may be null)
  double *partials;
  // Rescale::Dynamic (null / 0 otherwise)
  const double *resc;         // per half-edge rescale s_e of the node's majoriser: weighs the Q history term
  int split_g;                // 1: write g = DfobjE_own only and count omega_e > s_e in slot 6; D x follows in k_gfix
};

// ---- Rescale::Dynamic: per-pose constants from the rescale vector (update_quadratic_mat, DPGOProblem.cpp:751-840) ----
struct RescaleArgs {
  const int *rowptr;          // inter half-edges per own pose
  const InterRec *rec;
  const double *w;            // IRLS weights of the last evaluate_E per half-edge
  double *resc;               // out: s_e = clamp(1.25 w_e, 0.01, 1)
  const double *dintra;       // [NO][SYM]
  double *dinter, *gdiag, *tnv, *d00;
  double *ts_rec;             // PCG tile records; pose_rec[p] = index of the pose's diagonal entry
  const int *pose_rec;
  double xi;
};
struct GFixArgs {
  const double *x;            // Z_k own rows
  const double *dinter;
  double *g;                  // in: DfobjE_own, out: g = DfobjE_own - D x
  double xi;
  double *partials;           // slot 4: x^T D x
};

// ---- K3: fused extrapolation + proximal + polar projection -------------------
struct ProxArgs {
  const double *xa, *xb;      // X_k, X_{k-1} own rows (xb null -> no extrapolation)
  const double *dfa, *dfb;    // Df_k, Df_{k-1}
  const double *ga, *gb;      // g_k, g_{k-1} (may be null: no g extrapolation)
  const double *gamma;        // per node
  const double *tnv;          // [NO][1 + d + d*d]  T, N, V'
  const double *xref;         // for s0 = |Xout - xref|^2
  double *xout;               // X^{k+1/2}
  double *gex;                // extrapolated g (may be null)
  double *partials;
};

// ---- generic per-pose vector ops (tCG / TNT bookkeeping) --------------------
enum VecOp {
  V_CG_INIT = 0,   // s=0; Hs=0; r=grad; v=P(r); p=-v;     s0 = r.v, s1 = p.p... (IterativeSolvers.h:204-262)
  V_CG_STEP = 1,   // s+=a p (all rows: s.t collects the first-order translation update); Hs+=a Hp; r+=a Hp; v=P(r);   s0 = r.v
  V_CG_DIR = 2,    // p = -v + b p
  V_CG_FINAL = 3,  // s += sigma p; Hs += sigma Hp  (p optionally negated first)
  V_RETRACT = 4,   // xprop.Y = proj(x.Y + s.Y), xprop.t = x.t + s.t (initial guess of the solve that follows)  (DPGOProblem.cpp:127-143)
  V_DOTS = 5,      // s0 = a.b  s1 = a.a  s2 = b.b (rotation rows)
  V_COPY_ROT = 6,  // out.Y = a.Y
  V_COPY = 7,      // out = a
  V_PRECOND = 8,   // out = P(a); s0 = out.out
  V_DIFFNORM = 10, // s0 = |a - b|^2 (all rows)
  V_COPY_T = 13,   // o1.t = a.t
  V_CG_PRE = 14    // RegularizedCholesky, second half of V_CG_STEP: v = Proj(Y, pre); s0 = r.v   (a = r, o3 = v)
};

struct VecArgs {
  const double *a, *b, *c;
  double *o1, *o2, *o3, *o4;
  double *o5;               // tCG: H s, accumulated alongside s (H s = sum alpha_k H p_k)
  const double *y;          // point for tangent projection (pose blocks)
  const double *pinv;       // [NO][d*d] block-Jacobi inverse or [NO][d] Jacobi
  int precon;               // 0 none, 1 jacobi, 2 block jacobi, 3 regularized Cholesky (M^{-1} r precomputed in `pre`)
  const double *pre;        // precon 3: (G11 + reg)^{-1} r in pose-block layout, written by the sparse sweeps before the launch
  const double *coef;       // [num_nodes][MAXC] per-node coefficients
  double *partials;
};

// ---- K2b: translation solve --------------------------------------------------
// persistent per-node Jacobi-PCG (mmpgo_tsolve.cu).  A CTA tile is <= CTILE consecutive poses of
// one node; warp wi of the CTA owns its poses [32 wi, 32 wi + 32).  Solver vectors are laid out
// [cta tile][warp][D][32]; slice index sl = 8 * cta tile + warp.
constexpr int CTILE = 128;        // poses per CTA tile of the translation solve
constexpr int TS_WPT = CTILE / 32;  // 32-pose warp slices per CTA tile
constexpr int TS_NGRP = 3;        // consumer groups of the persistent solve kernel
constexpr int TS_HALO = 2;        // neighbour tiles staged on each side for the z gathers
constexpr int TS_MAXCT = 128;     // CTA tiles one persistent CTA can own
struct TSolveArgs {
  const int *rowptr, *col;        // G00 CSR over own poses (warm-start residual only)
  const double *a00, *d00;
  const int *sell_ptr;            // [TS_WPT n_ct + 1] slice offsets (units of 32 entries)
  const unsigned char *sell_pack; // per CTA tile (rows = its ELLPACK rows): {-tau}[rows][32] doubles, then
                                  // {slot of the neighbour's column 0}[rows][32] ints; tile at byte 384 sell_ptr[TS_WPT ct]
  const int *ct_node, *ct_start, *ct_cnt;
  int n_ct;
  int chunk;                      // k_tsolve_lite: contiguous CTA tiles per CTA
  const int *cta_ptr, *cta_tiles; // k_tsolve: CTA tiles of every persistent CTA (CSR over the grid), ascending per CTA
  const int *node_parts;          // k_tsolve: CTAs that hold tiles of the node (arrivals per phase)
  const int *node_ctb, *node_cte; // CTA-tile range of every node
  const int *active;              // per-node mask or nullptr
  const double *rhs;              // [NO][D]
  double *xio;                    // pose blocks: t rows in (warm start: u0 = -t) and out (t = -u)
  int warm;
  double *rec;                    // per CTA tile one record {x, p, Ap}[TS_WPT][D][32] + diag[CTILE] (padded with 1)
  double *z;                      // [TS_WPT n_ct][D][32], padded by TS_HALO tiles at both ends; r is kept as z = r / diag
  double *partials;               // [n_ct][4]
  double *nstate;                 // [nodes][8]: rz, bb, alpha, beta, iters, rr, finished-in-round + 1
  int *cnt;                       // [3 nodes + 8] arrival counters, epochs, hand-off flags, finished count (zeroed before launch)
  int n_nodes, n_active;
  unsigned long long *stats;      // [2]: sum of node iterations, sum of iterations x poses of the node
  const int *node_off;            // [nodes+1] own pose offsets
  double tol2;
  int max_iters;
  // k_tsolve_lite, dynamic shared memory plan (byte offsets; what does not fit stays in L2):
  int lite_stage_bytes;           // ELLPACK rows staged at offset 0 (0 = read from L2)
  int lite_vec_off;               // tile records {x, p, Ap, diag} resident (-1 = in global memory)
  int lite_z_off;                 // copy of the CTA's own z tiles (-1 = none)
  int lite_dyn_bytes;             // total
  // hand-off k_tsolve -> k_tsolve_lite.  cnt layout: [n_nodes] arrivals, [n_nodes] epochs, [n_nodes]
  // hand-off flags, then the count of finished nodes (all zeroed before the first launch)
  int handoff_live;               // k_tsolve: nodes still iterating at which the rest is handed off (0 = never)
  int resume;                     // k_tsolve_lite: continue the flagged nodes instead of starting a solve
};
template <int D> int launch_tsolve(const TSolveArgs &a, int grid, cudaStream_t s);
template <int D> int tsolve_max_grid(int device);
// small-shard variant (k_tsolve_lite): `chunk` = contiguous CTA tiles per CTA, `partials` holds two
// buffers of [n_ct][4]
constexpr int TSL_STAGE_MAX = 216 * 1024;   // dynamic shared memory of k_tsolve_lite (staged ELLPACK rows)
constexpr int TSL_MAXT = 32;      // CTA tiles (hence node segments) one lite CTA can own
template <int D> int launch_tsolve_lite(const TSolveArgs &a, int grid, cudaStream_t s);
template <int D> int tsolve_lite_max_grid(int device);

// ---- edge-parallel global objective (AMM-PGO*, DPGOStar.cpp:713-761) ----------
struct EdgeRec {            // host-side staging record; the device keeps the fields as struct-of-arrays
  int32_t i, j;             // pose indices (own / halo numbering)
  double tau, kappa;
  double t[3];
  double R[9];
  int32_t inter;            // 1: inter-node edge (robust kernel applies)
  int32_t pad;
};
static_assert(sizeof(EdgeRec) == 128, "EdgeRec must be 128 bytes");

template <int D> void launch_gpass(int mode, const Tiles &tl, const GPassArgs &a, cudaStream_t s);
template <int D> void launch_inter(int mode, const Tiles &tl, const InterArgs &a, cudaStream_t s);
template <int D> void launch_prox(const Tiles &tl, const ProxArgs &a, cudaStream_t s);
template <int D> void launch_rescale(const Tiles &tl, const RescaleArgs &a, cudaStream_t s);
template <int D> void launch_gfix(const Tiles &tl, const GFixArgs &a, cudaStream_t s);
template <int D> void launch_vec(int op, const Tiles &tl, const VecArgs &a, cudaStream_t s);
// rotation rows of a pose-block vector -> right-hand side of the G11 sweeps in elimination order
template <int D> void launch_gather_rot(const Tiles &tl, const double *src, const int *perm, double *rhs, cudaStream_t s);
void launch_reduce(int num_nodes, const int *node_tile_begin, const int *node_tile_end,
                   const double *partials, double *node_scal, cudaStream_t s);
template <int D> void launch_edge_objective(int64_t n_edges, const int *idx, const double *val, const double *x,
                                            int loss, double loss_reg, double *block_partials,
                                            int *n_blocks_out, cudaStream_t s);
void launch_sum_blocks(int n_blocks, const double *block_partials, double *out, cudaStream_t s);
void launch_sum_strided(int n, const double *v, int stride, double *out, cudaStream_t s);

// translation solve, dense path (small nodes): t = -G00^{-1} rhs
template <int D> void launch_dense_solve(int num_nodes, const int *node_off, const long long *node_dense_off,
                                         const int *node_active, const double *ginv, const double *rhs,
                                         double *xout /*pose blocks, row 0 = -Ginv rhs*/, int max_n0,
                                         cudaStream_t s);
// halo pack / unpack (DPGOHash::receive wire format, DPGOHash.cpp:45-82)
template <int D> void launch_gather_poses(int64_t n, const int *idx, const double *src, double *dst, cudaStream_t s);
template <int D> void launch_scatter_poses(int64_t n, const int *idx, const double *src, double *dst, cudaStream_t s);
// dst pose dst_idx[i] = src pose src_idx[i]  (pack / unpack of the two-array halo exchange)
template <int D> void launch_copy_poses(int64_t n, const int *src_idx, const int *dst_idx, const double *src, double *dst,
                                        cudaStream_t s);

// global iterate (column-major, rows [t; R blocks]) <-> pose blocks, on the device
template <int D> void launch_pack_poses(int64_t n, const int64_t *gid, const double *X, int64_t ld, int64_t N,
                                        double *d0, double *d1, double *d2, double *d3, double *d4, cudaStream_t s);
template <int D> void launch_unpack_poses(int64_t n, const int64_t *gid, const double *src, double *X, int64_t ld,
                                          int64_t N, cudaStream_t s);

template <int D> void launch_project_blocks(int64_t n, const double *A, double *U, cudaStream_t s);

}  // namespace mmpgo
