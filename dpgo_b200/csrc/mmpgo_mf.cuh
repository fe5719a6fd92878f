// Sparse direct solve of the per-node SPD systems (K2b, direct path): nested-dissection
// multifrontal Cholesky on the host at graph upload, level-scheduled supernodal sweeps on the
// device per solve.
//
// Replaces the reference's CHOLMOD factor `L_ = chol(G00)` (C++/DPGO/src/DPGOProblem.cpp:93) and
// its `L_.solve` call sites: recover_translations (include/DPGO/DPGOProblem.h:275-294), retract
// (DPGOProblem.cpp:127-143), the inner solve of the reduced Hessian-vector product (:552-577).
// The same machinery factors G11 + lambda I for the reference's default preconditioner
// (RegularizedCholesky, DPGOProblem.cpp:101-124, applied at :592).
//
// Data model.  The matrix of one robot node is ordered by nested dissection (separators from
// breadth-first level structures); every node of the separator tree is one SUPERNODE s with k_s
// columns (its vertices, contiguous in the new numbering) and a boundary of m_s rows in its
// ancestors (R_s = k_s + m_s rows in all).  With the front's Cholesky panel [L11; L21] the device
// keeps ONE dense block per supernode
//        M_s = [ inv(L11) ; L21 inv(L11) ]        (R_s x k_s)
// so that both sweeps are plain dense products without a triangular dependency inside a supernode:
//   forward   [ y_s ; -du ] = M_s f1,  f1 = b_s + (children's update rows),  u_s = f2 - L21 y_s
//   backward  x_s = M_s^T [ y_s ; -x_boundary ]
// M_s is stored twice: row-major (forward: a lane owns one output row, or a quarter of it, and
// streams it with 128-bit loads) and column-major (backward: the same for output columns), rows /
// columns zero-padded to an even length.  Supernodes of equal height (forward) / depth (backward)
// are independent: one grid-wide barrier per level.
// Every sum has a fixed order: results do not depend on scheduling or on which nodes share a GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace mmpgo {

struct MfSn {             // one supernode, 48 bytes
  int node;               // local robot node
  int c0;                 // first column, handle-wide permuted numbering (node offset included)
  int k, R;               // columns, rows
  long long moff;         // column-major copy in M: column j at moff + j * Rp, Rp = R rounded up to even   (backward sweep)
  long long mtoff;        // row-major copy in MT: row i at mtoff + i * kp, kp = k rounded up to even        (forward sweep)
  int rowoff;             // its R rows in pull0 / pull1
  int boff;               // its m = R - k boundary rows in bidx
  int uoff;               // first row of its update vector in the u buffer
  int nchild;             // 0 for the leaves of the separator tree (no update rows to pull)
};
static_assert(sizeof(MfSn) == 48, "MfSn must be 48 bytes");

// One warp job: rows [r0, r0 + n) of supernode sn in the forward sweep, columns [r0, r0 + n) in the backward sweep.
// It may start when the counters wait0 / wait1 (indices into MfDevice::done, -1: none) have reached need0 / need1:
// forward the finished forward jobs of the two children, backward those of the parent (the root: its own forward jobs).
struct MfJob { int sn, r0, n, wait0, wait1, need0, need1, pad; };
static_assert(sizeof(MfJob) == 32, "MfJob must be 32 bytes");

constexpr int MF_THREADS = 512;    // one persistent CTA per SM: 16 warps with up to 128 registers each (measured per solve on the
                                   // 1 M-pose grid: 256 threads 1.29, 384: 1.10, 512: 0.97, 768: 1.16, 1024 (spills): 1.62 ms)
constexpr int MF_WARPS = MF_THREADS / 32;
constexpr int MF_BLK = 128;        // rows of a front's right-hand side a warp stages at a time
constexpr int MF_BUF = MF_BLK + 16; // its shared-memory buffer (the dot loops may overrun a range by < 16 zero terms)
constexpr int MF_KS = 64;          // forward: fronts with more columns are summed in MF_Q slices by MF_Q lanes per row
constexpr int MF_RS = 128;         // backward: fronts with more rows are summed in MF_Q slices by MF_Q lanes per column
constexpr int MF_Q = 4;            // slices: term j of a block of MF_BLK belongs to slice (j % MF_BLK) / (MF_BLK / MF_Q)
constexpr int MF_QW = MF_BLK / MF_Q;
constexpr int MF_SROWS = 32 / MF_Q;   // rows (columns) per sliced job

inline int mf_even(int x) { return (x + 1) & ~1; }

// Host-side factor of all local nodes of a handle (or of one matrix in the host-only tests).
struct MfFactor {
  int nrows = 0;                       // total scalar rows (sum over nodes)
  int block = 1;                       // scalar rows per graph vertex (1: G00, d: G11)
  std::vector<MfSn> sn;
  std::vector<double> M, MT;
  std::vector<int> pull0, pull1;       // per supernode row: row of a child's update vector in u, or -1
  std::vector<int> bidx;               // per boundary row: position in the permuted numbering
  std::vector<int> iperm, perm;        // permuted position -> original row (handle-wide) and its inverse
  // per sweep (0 forward, 1 backward): the warp jobs with their ranges per stage
  std::vector<MfJob> wjobs[2];
  std::vector<int> wstage[2];              // [stage[s], stage[s+1])
  // dependencies between supernodes: a forward job waits for all forward jobs of the two children, a backward job for
  // all backward jobs of the parent (the root: for its own forward jobs)
  struct Dep { int child0, child1, parent, nf, nb, pad[3]; };
  std::vector<Dep> dep;
  int urows = 0;                       // rows of the u buffer
  int max_R = 0;                       // largest front
  int64_t nnz = 0;                     // sum k(k+1)/2 + k m  (entries of L)
  double flops = 0.0;
  int height = 0;
};

// One SPD matrix per node, CSR over scalar rows local to the node (full symmetric pattern incl.
// the diagonal; duplicate entries are summed).  `block` consecutive rows form one graph vertex.
struct MfMatrix {
  int n = 0;
  const int *ptr = nullptr, *col = nullptr;
  const double *val = nullptr;
  bool skip = false;      // rows are numbered but not factored (node served by another solve path)
};

// Symbolic pass only: fills nnz / flops / height estimates (no values).  Returns 0.
// Full factorisation: returns 0, or -1 when a pivot is not positive (matrix not SPD).
int mf_factor(const std::vector<MfMatrix> &mats, int block, int leaf, bool symbolic_only, MfFactor *out);
// Host restatement of the device sweeps (same blocks, same order of every sum): x = A^{-1} rhs,
// rhs / x are [nrows][nrhs] in the ORIGINAL numbering.
void mf_host_solve(const MfFactor &F, int nrhs, const double *rhs, double *x);

// ---- device side ---------------------------------------------------------------------------
struct MfDevice {
  const MfSn *sn = nullptr;
  const double *M = nullptr, *MT = nullptr;
  const int *pull0 = nullptr, *pull1 = nullptr, *bidx = nullptr, *iperm = nullptr;
  const MfJob *wjobs[2] = {nullptr, nullptr};
  const int *wstage[2] = {nullptr, nullptr};
  int n_stage[2] = {0, 0};
  double *y = nullptr, *xp = nullptr, *u = nullptr;     // [nrows][D] permuted, [urows][D]
  const MfFactor::Dep *dep = nullptr;                   // per supernode: children, parent, job counts
  unsigned *done = nullptr;                             // [2][supernodes] finished forward / backward jobs (zeroed before launch)
  int n_sn = 0;
  unsigned *barrier = nullptr;                          // grid barrier counter (zeroed before launch)
  unsigned long long *stage_ns = nullptr;               // [n_stage[0] + n_stage[1] + 1] globaltimer of CTA 0 at every stage boundary
  int smem_bytes = 0;
  int max_ctas = 0;                                     // CTAs the fullest stage can use (grid sizing)
};
struct MfSolveArgs {
  MfDevice f;
  const int *active;        // per-node mask or nullptr
  const double *rhs;        // [rows][D] in the PERMUTED numbering (row perm[p] holds the entry of original row p)
  double *out;              // out[row * out_stride + c] = sign * x
  int out_stride;
  double sign;
  int dry;                  // measurement: walk the stages and barriers without executing the jobs
  int level_sync;           // 1: a grid barrier after every level of the separator tree (per-level timing); 0: every job
                            // waits for the supernodes it depends on only, levels and nodes overlap
};
template <int D> int launch_mf_solve(const MfSolveArgs &a, int grid, cudaStream_t s);
template <int D> int mf_solve_max_grid(int device, int smem_bytes);

}  // namespace mmpgo
