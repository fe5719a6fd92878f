// Graph upload: partition, local indexing and the static operators of every
// local robot node, regenerated per edge instead of as Eigen sparse matrices.
//
// Restates the *output semantics* of
//   DPGO::read_g2o            C++/DPGO/src/DPGO_utils.cpp:140-202 (partition)
//   generate_data_info        :326-438  (own poses by ascending id, neighbours)
//   simplify_quadratic_data_matrix :1398-2288 and simplify_regular_data_matrix
//   (Static) :2290-2967       (G, D, H -> T, N, V')
//   DPGOProblem::DPGOProblem  C++/DPGO/src/DPGOProblem.cpp:11-125 (L_ = chol(G00),
//   preconditioner)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

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

template <typename T> static int dalloc(Handle *h, T **p, size_t n) {
  void *q = nullptr;
  CK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
  CK(cudaMemsetAsync(q, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream));
  h->allocs.push_back(q);
  *p = static_cast<T *>(q);
  return 0;
}
template <typename T> static int upload(Handle *h, T **p, const std::vector<T> &v) {
  // the copy goes through the handle's stream like dalloc's memset: the stream is non-blocking,
  // so a copy on the legacy default stream would not be ordered after that memset
  int rc = dalloc(h, p, v.size());
  if (rc) return rc;
  if (!v.empty()) {
    CK(cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // v may be a temporary
  }
  return 0;
}

void driver_free(Handle *h) {
  nccl_destroy(h);
  for (void *p : h->allocs) cudaFree(p);
  h->allocs.clear();
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  h->h_pinned = nullptr;
  if (h->h_slot) cudaFreeHost(h->h_slot);
  h->h_slot = nullptr;
  if (h->h_ts_stats) cudaFreeHost(h->h_ts_stats);
  h->h_ts_stats = nullptr;
}

// the `index` lambda of read_g2o, DPGO_utils.cpp:147-158
struct Partition {
  int64_t q, inc_n, inc;
  Partition(int64_t N, int P) : q(N / P), inc_n(N - (int64_t)P * (N / P)), inc((N - (int64_t)P * (N / P)) * (N / P + 1)) {}
  inline void operator()(int64_t g, int &node, int64_t &pose) const {
    if (g < inc) { node = (int)(g / (q + 1)); pose = g % (q + 1); }
    else { const int64_t i = g - inc; node = (int)(i / q + inc_n); pose = i % q; }
  }
  inline int64_t first_gid(int node) const {
    return node < inc_n ? (int64_t)node * (q + 1) : inc + (int64_t)(node - inc_n) * q;
  }
};

static inline int symi(int r, int c) { return r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r; }

// Per-edge blocks of M_e, rows/cols [t, Y_0..Y_{d-1}]  (DPGO_utils.cpp:1542-1641)
static void edge_blocks(int d, const double *R, const double *t, double kap, double tau, double *Mii,
                        double *Mjj, double *Mij) {
  const int Rr = d + 1;
  std::memset(Mii, 0, sizeof(double) * Rr * Rr);
  std::memset(Mjj, 0, sizeof(double) * Rr * Rr);
  std::memset(Mij, 0, sizeof(double) * Rr * Rr);
  Mii[0] = tau; Mjj[0] = tau; Mij[0] = -tau;
  for (int k = 0; k < d; ++k) {
    Mii[1 + k] = tau * t[k];
    Mii[(1 + k) * Rr] = tau * t[k];
    Mjj[(1 + k) * Rr + 1 + k] = kap;
    Mij[(1 + k) * Rr] = -tau * t[k];
    for (int c = 0; c < d; ++c) {
      Mii[(1 + k) * Rr + 1 + c] = tau * t[k] * t[c] + (k == c ? kap : 0.0);
      Mij[(1 + k) * Rr + 1 + c] = -kap * R[k * d + c];
    }
  }
}

// dense SPD inverse via Cholesky (the reference factors G00 with CHOLMOD,
// DPGOProblem.cpp:93); nodes small enough keep G00^{-1} explicitly.
static bool dense_spd_inverse(int n, std::vector<double> &A) {
  std::vector<double> L((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j <= i; ++j) {
      double s = A[(size_t)i * n + j];
      const double *li = &L[(size_t)i * n], *lj = &L[(size_t)j * n];
      for (int k = 0; k < j; ++k) s -= li[k] * lj[k];
      if (i == j) {
        if (s <= 0.0) return false;
        L[(size_t)i * n + i] = std::sqrt(s);
      } else {
        L[(size_t)i * n + j] = s / L[(size_t)j * n + j];
      }
    }
  }
  std::vector<double> Lt((size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Lt[(size_t)j * n + i] = L[(size_t)i * n + j];
#pragma omp parallel for schedule(dynamic, 8)
  for (int c = 0; c < n; ++c) {
    std::vector<double> y(n, 0.0);
    for (int i = c; i < n; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      const double *li = &L[(size_t)i * n];
      for (int k = c; k < i; ++k) s -= li[k] * y[k];
      y[i] = s / li[i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      const double *lt = &Lt[(size_t)i * n];
      for (int k = i + 1; k < n; ++k) s -= lt[k] * y[k];
      y[i] = s / lt[i];
    }
    for (int i = 0; i < n; ++i) A[(size_t)i * n + c] = y[i];
  }
  // symmetrise
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) {
      const double v = 0.5 * (A[(size_t)i * n + j] + A[(size_t)j * n + i]);
      A[(size_t)i * n + j] = v;
      A[(size_t)j * n + i] = v;
    }
  return true;
}

static bool small_inverse(int d, const double *M, double *out) {
  // Gauss-Jordan on a d x d SPD block
  double a[9], b[9];
  for (int i = 0; i < d * d; ++i) { a[i] = M[i]; b[i] = 0.0; }
  for (int i = 0; i < d; ++i) b[i * d + i] = 1.0;
  for (int c = 0; c < d; ++c) {
    int piv = c;
    for (int r = c + 1; r < d; ++r) if (std::fabs(a[r * d + c]) > std::fabs(a[piv * d + c])) piv = r;
    if (a[piv * d + c] == 0.0) return false;
    if (piv != c) for (int k = 0; k < d; ++k) { std::swap(a[c * d + k], a[piv * d + k]); std::swap(b[c * d + k], b[piv * d + k]); }
    const double iv = 1.0 / a[c * d + c];
    for (int k = 0; k < d; ++k) { a[c * d + k] *= iv; b[c * d + k] *= iv; }
    for (int r = 0; r < d; ++r) if (r != c) {
      const double f = a[r * d + c];
      for (int k = 0; k < d; ++k) { a[r * d + k] -= f * a[c * d + k]; b[r * d + k] -= f * b[c * d + k]; }
    }
  }
  for (int i = 0; i < d * d; ++i) out[i] = b[i];
  return true;
}

// Uploads a host factor (mmpgo_factor.cu) and sizes the persistent launch of its sweeps.  `rows_out`: for every
// permuted position the row of the caller's output array it is scattered to (units of out_stride doubles);
// null = the factor's own numbering.
static int mf_upload(Handle *h, const MfFactor &F, int d, const std::vector<int> *rows_out, MfDevice &m, int *grid,
                     bool *level_sync_auto, int64_t *tasks, std::vector<int> *stage_jobs) {
  int rc = 0;
  const int smem = (int)sizeof(double) * (MF_WARPS + 1) * MF_BUF * d;   // one front buffer per warp (+ one spare: the last warp's loops may overrun)
  if (smem > 200 * 1024) return 1;
  MfSn *sn_d; double *M_d, *MT_d; int *p0, *p1, *bi, *ip;
  if ((rc = upload(h, &sn_d, F.sn))) return rc;
  if ((rc = upload(h, &M_d, F.M))) return rc;
  if ((rc = upload(h, &MT_d, F.MT))) return rc;
  if ((rc = upload(h, &p0, F.pull0))) return rc;
  if ((rc = upload(h, &p1, F.pull1))) return rc;
  if ((rc = upload(h, &bi, F.bidx))) return rc;
  if ((rc = upload(h, &ip, rows_out ? *rows_out : F.iperm))) return rc;
  m.sn = sn_d; m.M = M_d; m.MT = MT_d; m.pull0 = p0; m.pull1 = p1; m.bidx = bi; m.iperm = ip;
  for (int dir = 0; dir < 2; ++dir) {
    MfJob *wj; int *ws;
    if ((rc = upload(h, &wj, F.wjobs[dir]))) return rc;
    if ((rc = upload(h, &ws, F.wstage[dir]))) return rc;
    m.wjobs[dir] = wj; m.wstage[dir] = ws;
    m.n_stage[dir] = (int)F.wstage[dir].size() - 1;
    for (int st = 0; st < m.n_stage[dir]; ++st)
      m.max_ctas = std::max(m.max_ctas, (F.wstage[dir][st + 1] - F.wstage[dir][st] + MF_WARPS - 1) / MF_WARPS);
    if (tasks) *tasks += (int64_t)F.wjobs[dir].size();
  }
  if ((rc = dalloc(h, &m.y, (size_t)F.nrows * d))) return rc;
  if ((rc = dalloc(h, &m.xp, (size_t)F.nrows * d))) return rc;
  if ((rc = dalloc(h, &m.u, (size_t)std::max(F.urows, 1) * d))) return rc;
  if ((rc = dalloc(h, &m.barrier, (size_t)4))) return rc;
  {
    MfFactor::Dep *dp;
    if ((rc = upload(h, &dp, F.dep))) return rc;
    m.dep = dp; m.n_sn = (int)F.sn.size();
    if ((rc = dalloc(h, &m.done, (size_t)2 * F.sn.size() + 2))) return rc;
  }
  if ((rc = dalloc(h, &m.stage_ns, (size_t)(F.wstage[0].size() + F.wstage[1].size() + 2)))) return rc;
  if (stage_jobs) {
    stage_jobs->clear();
    for (int dir = 0; dir < 2; ++dir)
      for (size_t st = 0; st + 1 < F.wstage[dir].size(); ++st) {
        stage_jobs->push_back(F.wstage[dir][st + 1] - F.wstage[dir][st]);
        stage_jobs->push_back(0);
      }
  }
  m.smem_bytes = smem;
  const int mg = d == 2 ? mf_solve_max_grid<2>(h->opt.device, smem) : mf_solve_max_grid<3>(h->opt.device, smem);
  if (mg <= 0) { set_error("occupancy query for the sparse direct solve failed"); return MMPGO_ERR_CUDA; }
  *grid = std::max(1, std::min(mg, m.max_ctas));
  // level barriers or per-supernode dependencies?  Same arithmetic, same bits.  Since a job requests its static
  // operands (pull indices, right-hand side, boundary index list) before it waits and polls with relaxed loads, the
  // dependency schedule wins at every shard size measured (per solve, barriers vs dependencies: 64 nodes of the
  // 1 M-pose grid on one GPU 0.86 vs 0.64 ms; 32 nodes 0.63 vs 0.44; 16 nodes 0.52 vs 0.37; 8 nodes 0.47 vs 0.35);
  // the barrier schedule stays for the per-level timers (mmpgo_solver_stage_times, profile kind 10)
  *level_sync_auto = false;
  return 0;
}

int driver_set_graph(Handle *h, int d, int64_t N, int num_nodes, int nb, int ne, int64_t E, const int32_t *ei,
                     const int32_t *ej, const double *R, const double *t, const double *kappa,
                     const double *tau) {
  if (d != 2 && d != 3) { set_error("d must be 2 or 3"); return MMPGO_ERR_ARG; }
  if (N <= 0 || num_nodes <= 0 || nb < 0 || ne > num_nodes || nb >= ne || E < 0 || N < num_nodes) {
    set_error("inconsistent graph sizes");
    return MMPGO_ERR_ARG;
  }
  if (h->graph_set) { set_error("graph already set"); return MMPGO_ERR_STATE; }
  // Rescale::Dynamic (robust losses; DPGOProblem.cpp:40-86): G00 changes with the weights, on its diagonal only --
  // no factor or dense inverse is kept, every node runs the PCG translation solve
  h->dynamic = h->opt.rescale == MMPGO_RESCALE_DYNAMIC && h->opt.loss != MMPGO_LOSS_NONE;
  if (h->dynamic) {
    if (h->opt.translation_solver == MMPGO_TSOLVE_DIRECT) {
      set_error("Rescale::Dynamic changes G00 every few iterations: translation_solver = DIRECT is not available");
      return MMPGO_ERR_UNSUPPORTED;
    }
    if (h->opt.translation_solver == MMPGO_TSOLVE_AUTO) h->opt.translation_solver = MMPGO_TSOLVE_PCG;
    h->opt.dense_solve_max_n = 0;
  }
  h->d = d; h->N = N; h->num_nodes = num_nodes; h->node_begin = nb; h->node_end = ne; h->A = ne - nb;
  const int A = h->A, Rr = d + 1, PB = (d + 1) * d, BB = Rr * Rr, SYM = Rr * (Rr + 1) / 2, TNV = 1 + d + d * d;
  const Partition part(N, num_nodes);
  for (int64_t e = 0; e < E; ++e)
    if (ei[e] < 0 || ej[e] < 0 || ei[e] >= N || ej[e] >= N || ei[e] == ej[e]) {
      set_error("edge endpoint out of range");
      return MMPGO_ERR_ARG;
    }

  // ---- which poses appear in a measurement (generate_data_info only indexes those, :358-372)
  std::vector<uint8_t> present(N, 0);
  for (int64_t e = 0; e < E; ++e) { present[ei[e]] = 1; present[ej[e]] = 1; }
  // own index of every local pose: rank among the node's present poses (:400-415)
  h->node_off.assign(A + 1, 0);
  h->info.assign(A, NodeInfo());
  std::vector<int> dev_index(N, -1);
  h->own_gid.clear();
  for (int a = 0; a < A; ++a) {
    const int node = nb + a;
    const int64_t g0 = part.first_gid(node), g1 = (node + 1 < num_nodes) ? part.first_gid(node + 1) : N;
    h->info[a].first_gid = -1;
    for (int64_t g = g0; g < g1; ++g)
      if (present[g]) {
        if (h->info[a].first_gid < 0) h->info[a].first_gid = g;
        dev_index[g] = (int)h->own_gid.size();
        h->own_gid.push_back(g);
      }
    h->node_off[a + 1] = (int)h->own_gid.size();
    h->info[a].n0 = h->node_off[a + 1] - h->node_off[a];
    if (h->info[a].n0 == 0) { set_error("a node has no measurements"); return MMPGO_ERR_ARG; }
  }
  h->NO = (int)h->own_gid.size();
  const int NO = h->NO;
  std::vector<int> own_node(NO);
  for (int a = 0; a < A; ++a) for (int p = h->node_off[a]; p < h->node_off[a + 1]; ++p) own_node[p] = a;

  // ---- classify edges, collect halo poses
  std::vector<int> en_i(E), en_j(E);
  std::vector<int64_t> halo;
  for (int64_t e = 0; e < E; ++e) {
    int64_t pp;
    part(ei[e], en_i[e], pp);
    part(ej[e], en_j[e], pp);
    const bool li = en_i[e] >= nb && en_i[e] < ne, lj = en_j[e] >= nb && en_j[e] < ne;
    if (li && !lj) halo.push_back(ej[e]);
    if (lj && !li) halo.push_back(ei[e]);
  }
  std::sort(halo.begin(), halo.end());
  halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
  h->halo_gid = halo;
  h->NH = (int)halo.size();
  h->NP = NO + h->NH;
  h->halo_owner.resize(h->NH);
  for (int k = 0; k < h->NH; ++k) {
    int nd; int64_t pp;
    part(halo[k], nd, pp);
    h->halo_owner[k] = nd;
    dev_index[halo[k]] = NO + k;
  }

  h->stage_lo = h->own_gid.front(); h->stage_hi = h->own_gid.back() + 1;
  if (h->NH > 0) { h->stage_lo = std::min(h->stage_lo, halo.front()); h->stage_hi = std::max(h->stage_hi, halo.back() + 1); }

  // ---- count rows
  std::vector<int> rowcnt(NO, 0), xrowcnt(NO, 0);
  int64_t n_owned = 0;
  for (int64_t e = 0; e < E; ++e) {
    const bool li = en_i[e] >= nb && en_i[e] < ne, lj = en_j[e] >= nb && en_j[e] < ne;
    if (!li && !lj) continue;
    if (en_i[e] == en_j[e]) {
      rowcnt[dev_index[ei[e]]]++; rowcnt[dev_index[ej[e]]]++;
      h->info[en_i[e] - nb].m0++;
      n_owned++;
    } else {
      if (li) { xrowcnt[dev_index[ei[e]]]++; h->info[en_i[e] - nb].m1++; n_owned++; }
      if (lj) { xrowcnt[dev_index[ej[e]]]++; h->info[en_j[e] - nb].m1++; }
    }
  }
  std::vector<int> rowptr(NO + 1, 0), xrowptr(NO + 1, 0);
  for (int p = 0; p < NO; ++p) { rowptr[p + 1] = rowptr[p] + rowcnt[p]; xrowptr[p + 1] = xrowptr[p] + xrowcnt[p]; }
  const int64_t nnz = rowptr[NO], nxe = xrowptr[NO];
  h->n_intra_entries = nnz; h->n_inter_he = nxe; h->n_edges_owned = n_owned;
  std::vector<int> col(nnz), fill(NO, 0), xfill(NO, 0);
  std::vector<double> blk((size_t)nnz * BB), dintra((size_t)NO * SYM, 0.0), dinter((size_t)NO * SYM, 0.0);
  std::vector<InterRec> xrec(nxe);
  std::vector<EdgeRec> erec(n_owned);
  // neighbour sets per node for n1
  std::vector<std::vector<int64_t>> nbrs(A);
  int64_t eo = 0;
  double Mii[16], Mjj[16], Mij[16];
  for (int64_t e = 0; e < E; ++e) {
    const bool li = en_i[e] >= nb && en_i[e] < ne, lj = en_j[e] >= nb && en_j[e] < ne;
    if (!li && !lj) continue;
    const double *Re = R + (size_t)e * d * d, *te = t + (size_t)e * d;
    edge_blocks(d, Re, te, kappa[e], tau[e], Mii, Mjj, Mij);
    const int pi = dev_index[ei[e]], pj = dev_index[ej[e]];
    const bool intra = en_i[e] == en_j[e];
    if (intra || li) {
      EdgeRec &r = erec[eo++];
      std::memset(&r, 0, sizeof(r));
      r.i = pi; r.j = pj; r.tau = tau[e]; r.kappa = kappa[e]; r.inter = intra ? 0 : 1;
      for (int k = 0; k < d; ++k) r.t[k] = te[k];
      for (int k = 0; k < d * d; ++k) r.R[k] = Re[k];
    }
    if (intra) {
      int s = rowptr[pi] + fill[pi]++;
      col[s] = pj;
      std::memcpy(&blk[(size_t)s * BB], Mij, sizeof(double) * BB);
      s = rowptr[pj] + fill[pj]++;
      col[s] = pi;
      for (int r = 0; r < Rr; ++r) for (int c = 0; c < Rr; ++c) blk[(size_t)s * BB + r * Rr + c] = Mij[c * Rr + r];
      for (int r = 0; r < Rr; ++r) for (int c = 0; c <= r; ++c) {
        dintra[(size_t)pi * SYM + symi(r, c)] += Mii[r * Rr + c];
        dintra[(size_t)pj * SYM + symi(r, c)] += Mjj[r * Rr + c];
      }
    } else {
      for (int side = 0; side < 2; ++side) {
        const bool own_i = side == 0;
        if (own_i ? !li : !lj) continue;
        const int po = own_i ? pi : pj, pn = own_i ? pj : pi;
        const int a = (own_i ? en_i[e] : en_j[e]) - nb;
        const int s = xrowptr[po] + xfill[po]++;
        InterRec &r = xrec[s];
        std::memset(&r, 0, sizeof(r));
        r.other = pn; r.own_is_i = own_i ? 1 : 0; r.tau = tau[e]; r.kappa = kappa[e];
        for (int k = 0; k < d; ++k) r.t[k] = te[k];
        for (int k = 0; k < d * d; ++k) r.R[k] = Re[k];
        h->info[a].inter_he.push_back(s);
        nbrs[a].push_back(own_i ? ej[e] : ei[e]);
        if (pn >= NO) h->boundary_pairs.emplace_back(po, own_i ? en_j[e] : en_i[e]);
        const double *Md = own_i ? Mii : Mjj;
        for (int rr = 0; rr < Rr; ++rr) for (int c = 0; c <= rr; ++c) dinter[(size_t)po * SYM + symi(rr, c)] += Md[rr * Rr + c];
      }
    }
  }
  for (int a = 0; a < A; ++a) {
    std::sort(nbrs[a].begin(), nbrs[a].end());
    h->info[a].n1 = (int)(std::unique(nbrs[a].begin(), nbrs[a].end()) - nbrs[a].begin());
  }

  // ---- pose-local constants: G diagonal, H -> T, N, V', preconditioner, G00
  const double xi = h->opt.regularizer;
  std::vector<double> gdiag((size_t)NO * SYM), tnv((size_t)NO * TNV), pinv((size_t)NO * d * d, 0.0), d00(NO), a00(nnz);
  for (int p = 0; p < NO; ++p) {
    double Hm[16], Gm[16];
    for (int r = 0; r < Rr; ++r) for (int c = 0; c < Rr; ++c) {
      const double di = dintra[(size_t)p * SYM + symi(r, c)], dx = dinter[(size_t)p * SYM + symi(r, c)];
      Gm[r * Rr + c] = di + 2.0 * dx + (r == c ? xi : 0.0);          // :2212-2243
      // :1679-1755, 2038-2096; the Dynamic builder carries 0.5 xi on the auxiliary matrices (:3621, :3635)
      Hm[r * Rr + c] = 2.0 * di + 2.0 * dx + (r == c ? (h->dynamic ? 0.5 : 1.5) * xi : 0.0);
    }
    for (int r = 0; r < Rr; ++r) for (int c = 0; c <= r; ++c) gdiag[(size_t)p * SYM + symi(r, c)] = Gm[r * Rr + c];
    double *c_ = &tnv[(size_t)p * TNV];
    const double T = 1.0 / Hm[0];                                      // T = T.inverse(), :2280
    c_[0] = T;
    for (int k = 0; k < d; ++k) c_[1 + k] = T * Hm[1 + k];             // N = T * N, :2282
    for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c)            // V' = H_RR - H_Rt T H_tR, :2962-2964
      c_[1 + d + r * d + c] = Hm[(1 + r) * Rr + 1 + c] - Hm[(1 + r) * Rr] * (T * Hm[1 + c]);
    d00[p] = Gm[0];
    double g11[9];
    for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) g11[r * d + c] = Gm[(1 + r) * Rr + 1 + c];
    if (h->opt.preconditioner == MMPGO_PRECON_BLOCK_JACOBI) {
      if (!small_inverse(d, g11, &pinv[(size_t)p * d * d])) { set_error("singular G11 block"); return MMPGO_ERR_ARG; }
    } else {
      for (int r = 0; r < d; ++r) pinv[(size_t)p * d * d + r * d + r] = 1.0 / g11[r * d + r];  // DPGOProblem.cpp:96-98
    }
  }
  for (int64_t s = 0; s < nnz; ++s) a00[s] = blk[(size_t)s * BB];

  // ---- tiles (never straddle a node)
  std::vector<int> tnode, tstart, tcnt;
  h->h_node_tb.assign(A, 0); h->h_node_te.assign(A, 0);
  for (int a = 0; a < A; ++a) {
    h->h_node_tb[a] = (int)tnode.size();
    for (int p = h->node_off[a]; p < h->node_off[a + 1]; p += TILE) {
      tnode.push_back(a); tstart.push_back(p); tcnt.push_back(std::min(TILE, h->node_off[a + 1] - p));
    }
    h->h_node_te[a] = (int)tnode.size();
  }
  h->n_tiles = (int)tnode.size();
  h->h_tile_node = tnode;

  // ---- persistent solve: "CTA tiles" of <= 256 consecutive poses of one node, cut into eight
  // 32-pose warp slices; sliced ELLPACK of G00 with one slice per warp slice, columns stored as
  // slots of the [cta tile][warp][d][32] vector layout
  std::vector<int> ct_node, ct_start, ct_cnt, node_ctb(A, 0), node_cte(A, 0);
  for (int a = 0; a < A; ++a) {
    node_ctb[a] = (int)ct_node.size();
    for (int p = h->node_off[a]; p < h->node_off[a + 1]; p += CTILE) {
      ct_node.push_back(a); ct_start.push_back(p); ct_cnt.push_back(std::min(CTILE, h->node_off[a + 1] - p));
    }
    node_cte[a] = (int)ct_node.size();
  }
  h->n_ctiles = (int)ct_node.size();
  h->h_node_ctb = node_ctb; h->h_node_cte = node_cte;
  h->ts_plan_grid = -1;
  h->ts_force_kernel = h->opt.translation_solver == MMPGO_TSOLVE_PCG_RING ? 1
                       : h->opt.translation_solver == MMPGO_TSOLVE_PCG_LITE ? 2 : 0;
  const int n_sl = TS_WPT * h->n_ctiles;
  std::vector<int> sell_ptr((size_t)n_sl + 1, 0), slot_of(NO, 0);
  for (int c = 0; c < h->n_ctiles; ++c)
    for (int k = 0; k < ct_cnt[c]; ++k) slot_of[ct_start[c] + k] = (TS_WPT * c + k / 32) * 32 * d + (k % 32);
  for (int sl = 0; sl < n_sl; ++sl) {
    const int c = sl / TS_WPT, c0 = std::min(32, ct_cnt[c] - 32 * (sl % TS_WPT));
    int width = 0;
    for (int k = 0; k < c0; ++k) width = std::max(width, rowcnt[ct_start[c] + 32 * (sl % TS_WPT) + k]);
    sell_ptr[sl + 1] = sell_ptr[sl] + (c0 > 0 ? width : 0);
  }
  h->sell_entries = sell_ptr.back();
  h->h_sell_ptr = sell_ptr;
  // per-tile records {x, p, Ap, diag} (one bulk copy per phase) and the packed ELLPACK rows
  const size_t RL = (size_t)3 * CTILE * d + CTILE;
  std::vector<double> rec((size_t)h->n_ctiles * RL, 0.0);
  for (int c = 0; c < h->n_ctiles; ++c)
    for (int k = 0; k < CTILE; ++k) rec[(size_t)c * RL + 3 * CTILE * d + k] = k < ct_cnt[c] ? d00[ct_start[c] + k] : 1.0;
  std::vector<unsigned char> sell_pack((size_t)h->sell_entries * 384 + 16, 0);
  for (int c = 0; c < h->n_ctiles; ++c) {
    const int r0 = sell_ptr[TS_WPT * c], rows = sell_ptr[TS_WPT * (c + 1)] - r0;
    double *pv = reinterpret_cast<double *>(&sell_pack[(size_t)r0 * 384]);
    int *pc = reinterpret_cast<int *>(&sell_pack[(size_t)r0 * 384 + (size_t)rows * 256]);
    for (int w = 0; w < TS_WPT; ++w) {
      const int sl = TS_WPT * c + w, c0 = std::min(32, ct_cnt[c] - 32 * w);
      const int p0 = ct_start[c] + 32 * w;
      for (int s_ = sell_ptr[sl]; s_ < sell_ptr[sl + 1]; ++s_)
        for (int l = 0; l < 32; ++l) {
          const int k = s_ - sell_ptr[sl];
          const size_t o = (size_t)(s_ - r0) * 32 + l;
          pc[o] = slot_of[ct_start[c]];   // padding: any valid slot, value 0
          pv[o] = 0.0;
          if (c0 > 0 && l < c0 && k < rowcnt[p0 + l]) {
            const int e = rowptr[p0 + l] + k;
            pc[o] = slot_of[col[e]];
            pv[o] = a00[e];
          }
        }
    }
  }

  // ---- dense G00^{-1} for small nodes
  std::vector<long long> dense_off(A, -1);
  std::vector<double> ginv;
  h->dense_mask.assign(A, 0); h->pcg_mask.assign(A, 0);
  h->max_dense_n0 = 0;
  for (int a = 0; a < A; ++a) {
    const int n0 = h->info[a].n0;
    if (n0 <= h->opt.dense_solve_max_n) {
      std::vector<double> M((size_t)n0 * n0, 0.0);
      const int off = h->node_off[a];
      for (int p = 0; p < n0; ++p) {
        M[(size_t)p * n0 + p] += d00[off + p];
        for (int s = rowptr[off + p]; s < rowptr[off + p + 1]; ++s) M[(size_t)p * n0 + (col[s] - off)] += a00[s];
      }
      if (!dense_spd_inverse(n0, M)) { set_error("G00 is not positive definite"); return MMPGO_ERR_ARG; }
      dense_off[a] = (long long)ginv.size();
      ginv.insert(ginv.end(), M.begin(), M.end());
      h->info[a].dense = true; h->dense_mask[a] = 1; h->any_dense = true; h->dense_poses += n0;
      h->max_dense_n0 = std::max(h->max_dense_n0, n0);
    } else {
      h->pcg_mask[a] = 1; h->any_pcg = true;
    }
  }

  // ---- sparse Cholesky of G00 for the nodes above dense_solve_max_n (the reference factors every
  // node's G00 with CHOLMOD, DPGOProblem.cpp:93): host factorisation now, device sweeps per solve
  int rc = 0;
  if (h->any_pcg && (h->opt.translation_solver == MMPGO_TSOLVE_AUTO || h->opt.translation_solver == MMPGO_TSOLVE_DIRECT)) {
    // scalar CSR of every node's G00 (local column indices, diagonal included)
    std::vector<int> mptr((size_t)NO + A, 0), mcol;
    std::vector<double> mval;
    std::vector<MfMatrix> mats(A);
    mcol.reserve((size_t)nnz + NO); mval.reserve((size_t)nnz + NO);
    std::vector<size_t> ptr_off(A);
    size_t po = 0;
    for (int a = 0; a < A; ++a) {
      const int off = h->node_off[a], n0 = h->info[a].n0;
      ptr_off[a] = po;
      for (int p = 0; p < n0; ++p) {
        mptr[po + p] = (int)mcol.size();
        if (!h->pcg_mask[a]) continue;
        mcol.push_back(p); mval.push_back(d00[off + p]);
        for (int s2 = rowptr[off + p]; s2 < rowptr[off + p + 1]; ++s2) { mcol.push_back(col[s2] - off); mval.push_back(a00[s2]); }
      }
      mptr[po + n0] = (int)mcol.size();
      po += (size_t)n0 + 1;
    }
    for (int a = 0; a < A; ++a) {
      // row pointers are absolute offsets into mcol / mval: rebase the column / value pointers per node
      mats[a].n = h->info[a].n0; mats[a].ptr = &mptr[ptr_off[a]]; mats[a].col = mcol.data(); mats[a].val = mval.data();
      mats[a].skip = !h->pcg_mask[a];
    }
    // dissection stops at 32 poses: smaller leaves mean fewer factor entries (-12 % at 16) but more and shorter
    // jobs and two more levels; measured on the 1 M-pose grid 16: 1.13, 32: 0.98, 48: 0.95, 64: 0.94 ms per solve
    const int leaf = 32;
    MfFactor F;
    mf_factor(mats, 1, leaf, true, &F);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const double bytes = 16.0 * (double)F.nnz * 1.35;          // both copies; the stored blocks carry the zero triangle of inv(L11)
    const bool fits = bytes <= 0.3 * (double)free_b && F.flops <= 1.5e12;
    if (!fits && h->opt.translation_solver == MMPGO_TSOLVE_DIRECT) {
      set_error("translation_solver = DIRECT: the sparse factor of G00 exceeds the memory / setup budget");
      return MMPGO_ERR_UNSUPPORTED;
    }
    if (fits) {
      if (mf_factor(mats, 1, leaf, false, &F) != 0) { set_error("G00 is not positive definite"); return MMPGO_ERR_ARG; }
      rc = mf_upload(h, F, d, nullptr, h->mf, &h->mf_grid, &h->mf_level_sync_auto, &h->mf_tasks, &h->mf_stage_jobs);
      if (rc < 0) return rc;
      if (rc == 0) {
        if ((rc = upload(h, &h->d_mf_perm, F.perm))) return rc;
        h->h_mf_perm = F.perm;
        h->use_direct = true;
        h->mf_nnz = F.nnz; h->mf_entries = 2 * (int64_t)F.M.size(); h->mf_height = F.height;
        h->mf_supernodes = (int)F.sn.size();
      } else if (h->opt.translation_solver == MMPGO_TSOLVE_DIRECT) {
        set_error("translation_solver = DIRECT: a front of the sparse factor does not fit shared memory");
        return MMPGO_ERR_UNSUPPORTED;
      }
    }
  }

  // ---- RegularizedCholesky preconditioner (the reference's default, DPGO_types.h:155): Cholesky factor of
  // G11 + (lambda_max / max_cond) I per node (DPGOProblem.cpp:101-124), G11 = the rotation rows / columns of the
  // majoriser G.  Same host factorisation and device sweeps as G00, d scalar rows per pose.
  h->use_regchol = false;
  if (h->opt.preconditioner == MMPGO_PRECON_REGULARIZED_CHOLESKY) {
    const int64_t rows = (int64_t)NO * d;
    std::vector<int> mptr((size_t)rows + A, 0), mcol;
    std::vector<double> mval;
    mcol.reserve(((size_t)nnz + NO) * d * d); mval.reserve(((size_t)nnz + NO) * d * d);
    std::vector<MfMatrix> mats(A);
    std::vector<size_t> ptr_off(A);
    h->lambda_max.assign(A, 0.0);
    size_t po = 0;
    for (int a = 0; a < A; ++a) {
      const int off = h->node_off[a], n0 = h->info[a].n0;
      ptr_off[a] = po;
      for (int p = 0; p < n0; ++p)
        for (int r = 0; r < d; ++r) {
          mptr[po + (size_t)p * d + r] = (int)mcol.size();
          for (int c = 0; c < d; ++c) {
            mcol.push_back(p * d + c);
            mval.push_back(gdiag[(size_t)(off + p) * SYM + symi(std::max(1 + r, 1 + c), std::min(1 + r, 1 + c))]);
          }
          for (int s2 = rowptr[off + p]; s2 < rowptr[off + p + 1]; ++s2)
            for (int c = 0; c < d; ++c) {
              mcol.push_back((col[s2] - off) * d + c);
              mval.push_back(blk[(size_t)s2 * BB + (1 + r) * Rr + 1 + c]);
            }
        }
      mptr[po + (size_t)n0 * d] = (int)mcol.size();
      po += (size_t)n0 * d + 1;
    }
    // lambda_max of every node's G11 (the reference asks Spectra for a loose estimate, tolerance 1e-4): power
    // iteration on the positive semidefinite matrix, Rayleigh quotient stationary to 1e-6 relative
#pragma omp parallel for schedule(dynamic, 1)
    for (int a = 0; a < A; ++a) {
      const int n = h->info[a].n0 * d;
      const int *ptr = &mptr[ptr_off[a]];
      std::vector<double> v((size_t)n), w((size_t)n);
      for (int i = 0; i < n; ++i) v[i] = 1.0 + 0.5 * std::sin(0.7 * i + 0.3);
      double lam = 0.0;
      for (int it = 0; it < 2000; ++it) {
        double nv = 0.0;
        for (int i = 0; i < n; ++i) nv += v[i] * v[i];
        nv = 1.0 / std::sqrt(nv);
        for (int i = 0; i < n; ++i) v[i] *= nv;
        double rq = 0.0;
        for (int i = 0; i < n; ++i) {
          double t = 0.0;
          for (int e = ptr[i]; e < ptr[i + 1]; ++e) t += mval[e] * v[mcol[e]];
          w[i] = t; rq += t * v[i];
        }
        v.swap(w);
        const bool stop = it > 8 && std::fabs(rq - lam) <= 1e-6 * std::fabs(rq);
        lam = rq;
        if (stop) break;
      }
      h->lambda_max[a] = lam;
    }
    // the regulariser goes onto the (first) diagonal entry of every row
    for (int a = 0; a < A; ++a) {
      const int n = h->info[a].n0 * d;
      const int *ptr = &mptr[ptr_off[a]];
      const double reg = h->lambda_max[a] / h->opt.reg_Cholesky_precon_max_condition_number;
      for (int i = 0; i < n; ++i)
        for (int e = ptr[i]; e < ptr[i + 1]; ++e)
          if (mcol[e] == i) { mval[e] += reg; break; }
      mats[a].n = n; mats[a].ptr = ptr; mats[a].col = mcol.data(); mats[a].val = mval.data(); mats[a].skip = false;
    }
    const int leaf11 = d == 3 ? 12 : 16;               // vertices (poses) per leaf: ~36 / 32 scalar columns
    MfFactor F;
    mf_factor(mats, d, leaf11, true, &F);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (16.0 * (double)F.nnz * 1.35 > 0.5 * (double)free_b || F.flops > 4e13) {
      set_error("preconditioner = RegularizedCholesky: the sparse factor of G11 exceeds the memory / setup budget");
      return MMPGO_ERR_UNSUPPORTED;
    }
    if (mf_factor(mats, d, leaf11, false, &F) != 0) { set_error("G11 + reg is not positive definite"); return MMPGO_ERR_ARG; }
    // permuted position -> row of a pose-block array (d + 1 rows of d doubles per pose; rotation row r of pose p is
    // row p (d + 1) + 1 + r), so that the backward sweep scatters straight into pose-block layout
    std::vector<int> rows_out(F.iperm.size());
    for (size_t i = 0; i < F.iperm.size(); ++i) rows_out[i] = (F.iperm[i] / d) * (d + 1) + 1 + F.iperm[i] % d;
    int64_t tasks = 0;
    rc = mf_upload(h, F, d, &rows_out, h->mf11, &h->mf11_grid, &h->mf11_level_sync, &tasks, nullptr);
    if (rc != 0) {
      if (rc > 0) { set_error("preconditioner = RegularizedCholesky: a front of the sparse factor does not fit shared memory"); rc = MMPGO_ERR_UNSUPPORTED; }
      return rc;
    }
    if ((rc = upload(h, &h->d_mf11_perm, F.perm))) return rc;
    if ((rc = dalloc(h, &h->rhs11, (size_t)rows * d))) return rc;
    if ((rc = dalloc(h, &h->pre_buf, (size_t)NO * (d + 1) * d))) return rc;
    h->use_regchol = true;
    h->mf11_nnz = F.nnz;
  }

  // ---- upload
  {
    std::vector<int64_t> pg(h->own_gid);
    pg.insert(pg.end(), h->halo_gid.begin(), h->halo_gid.end());
    if ((rc = upload(h, &h->d_pose_gid, pg))) return rc;
  }
  if ((rc = upload(h, &h->d_rowptr, rowptr))) return rc;
  if ((rc = upload(h, &h->d_col, col))) return rc;
  if ((rc = upload(h, &h->d_blk, blk))) return rc;
  {
    // row 0 of every off-diagonal block, contiguous: the G01 pass (k_g01) streams only these
    const int Rr0 = d + 1;
    std::vector<double> blk0((size_t)nnz * Rr0);
    for (int64_t s2 = 0; s2 < nnz; ++s2)
      for (int c = 0; c < Rr0; ++c) blk0[(size_t)s2 * Rr0 + c] = blk[(size_t)s2 * BB + c];
    if ((rc = upload(h, &h->d_blk0, blk0))) return rc;
  }
  if ((rc = upload(h, &h->d_gdiag, gdiag))) return rc;
  if ((rc = upload(h, &h->d_dintra, dintra))) return rc;
  if ((rc = upload(h, &h->d_dinter, dinter))) return rc;
  if ((rc = upload(h, &h->d_tnv, tnv))) return rc;
  if ((rc = upload(h, &h->d_pinv, pinv))) return rc;
  if ((rc = upload(h, &h->d_a00, a00))) return rc;
  if ((rc = upload(h, &h->d_d00, d00))) return rc;
  if ((rc = upload(h, &h->d_xrowptr, xrowptr))) return rc;
  if ((rc = upload(h, &h->d_xrec, xrec))) return rc;
  {
    const size_t ne = erec.size();
    std::vector<int> eidx(3 * ne);
    std::vector<double> evals(14 * ne, 0.0);
    for (size_t e = 0; e < ne; ++e) {
      eidx[e] = erec[e].i; eidx[ne + e] = erec[e].j; eidx[2 * ne + e] = erec[e].inter;
      evals[e] = erec[e].tau; evals[ne + e] = erec[e].kappa;
      for (int k = 0; k < 3; ++k) evals[(2 + k) * ne + e] = erec[e].t[k];
      for (int k = 0; k < 9; ++k) evals[(5 + k) * ne + e] = erec[e].R[k];
    }
    if ((rc = upload(h, &h->d_eidx, eidx))) return rc;
    if ((rc = upload(h, &h->d_eval, evals))) return rc;
  }
  if ((rc = upload(h, &h->d_ginv, ginv))) return rc;
  if ((rc = upload(h, &h->d_sell_ptr, sell_ptr))) return rc;
  if ((rc = upload(h, &h->d_sell_pack, sell_pack))) return rc;
  if ((rc = upload(h, &h->ts_rec, rec))) return rc;
  if ((rc = upload(h, &h->d_ct_node, ct_node))) return rc;
  if ((rc = upload(h, &h->d_ct_start, ct_start))) return rc;
  if ((rc = upload(h, &h->d_ct_cnt, ct_cnt))) return rc;
  if ((rc = upload(h, &h->d_node_ctb, node_ctb))) return rc;
  if ((rc = upload(h, &h->d_node_cte, node_cte))) return rc;
  if ((rc = upload(h, &h->d_dense_off, dense_off))) return rc;
  if ((rc = upload(h, &h->d_tile_node, tnode))) return rc;
  if ((rc = upload(h, &h->d_tile_start, tstart))) return rc;
  if ((rc = upload(h, &h->d_tile_cnt, tcnt))) return rc;
  if ((rc = upload(h, &h->d_node_tb, h->h_node_tb))) return rc;
  if ((rc = upload(h, &h->d_node_te, h->h_node_te))) return rc;
  if ((rc = upload(h, &h->d_node_off, h->node_off))) return rc;
  if ((rc = dalloc(h, &h->d_active, (size_t)A))) return rc;
  if ((rc = dalloc(h, &h->d_active2, (size_t)A))) return rc;
  const size_t np = (size_t)h->NP * PB, no = (size_t)NO * PB, nc = (size_t)NO * d;
  for (int k = 0; k < 3; ++k) if ((rc = dalloc(h, &h->X[k], np))) return rc;
  double **pv[] = {&h->Xakh, &h->Yex, &h->xprop, &h->xeval};
  for (auto q : pv) if ((rc = dalloc(h, q, np))) return rc;
  double **ov[] = {&h->g[0], &h->g[1], &h->Df[0], &h->Df[1], &h->gex, &h->Dfex, &h->nab, &h->grad,
                   &h->cg_s, &h->cg_r, &h->cg_v, &h->cg_p, &h->cg_Hp, &h->cg_Hs, &h->tdot_prev};
  for (auto q : ov) if ((rc = dalloc(h, q, no))) return rc;
  double **cv[] = {&h->rhs_t};
  for (auto q : cv) if ((rc = dalloc(h, q, nc))) return rc;
  {
    const size_t nv = (size_t)h->n_ctiles * CTILE * d;
    // z is staged with a TS_HALO-tile halo on both sides: pad the array accordingly at each end
    if ((rc = dalloc(h, &h->ts_z_base, nv + (size_t)2 * TS_HALO * CTILE * d))) return rc;
    h->ts_z = h->ts_z_base + (size_t)TS_HALO * CTILE * d;
    if ((rc = dalloc(h, &h->ts_partials, (size_t)h->n_ctiles * 4 * 2))) return rc;   // two buffers (k_tsolve_lite)
    if ((rc = dalloc(h, &h->ts_nstate, (size_t)A * 8))) return rc;
    if ((rc = dalloc(h, &h->d_ts_sync, (size_t)3 * A + 8))) return rc;
    if ((rc = dalloc(h, &h->d_ts_stats, (size_t)4))) return rc;
    CK(cudaMallocHost((void **)&h->h_ts_stats, 4 * sizeof(unsigned long long)));
    std::memset(h->h_ts_stats, 0, 4 * sizeof(unsigned long long));
    h->ts_unconv_seen = 0;
    if ((rc = dalloc(h, &h->d_cta_ptr, (size_t)1024 + 1))) return rc;
    if ((rc = dalloc(h, &h->d_cta_tiles, (size_t)h->n_ctiles + 1))) return rc;
    if ((rc = dalloc(h, &h->d_node_parts, (size_t)A))) return rc;
    h->ts_max_grid = d == 2 ? tsolve_max_grid<2>(h->opt.device) : tsolve_max_grid<3>(h->opt.device);
    if (h->ts_max_grid <= 0) { set_error("occupancy query for the translation solve failed"); return MMPGO_ERR_CUDA; }
    h->tsl_max_grid = d == 2 ? tsolve_lite_max_grid<2>(h->opt.device) : tsolve_lite_max_grid<3>(h->opt.device);
    if (h->tsl_max_grid <= 0) { set_error("occupancy query for the small-shard translation solve failed"); return MMPGO_ERR_CUDA; }
  }
  double **wv[] = {&h->w_cur, &h->w_prev, &h->w_tmp};
  for (auto q : wv) if ((rc = dalloc(h, q, (size_t)nxe))) return rc;
  if (h->dynamic) {
    std::vector<double> ones((size_t)nxe, 1.0);                  // DiagReScale_.setOnes (DPGOProblem.cpp:31)
    if ((rc = upload(h, &h->resc, ones))) return rc;
    std::vector<int> pose_rec(NO, 0);
    for (int c = 0; c < h->n_ctiles; ++c)
      for (int k = 0; k < ct_cnt[c]; ++k) pose_rec[ct_start[c] + k] = (int)((size_t)c * RL + 3 * CTILE * d + k);
    if ((rc = upload(h, &h->d_pose_rec, pose_rec))) return rc;
  }
  if ((rc = dalloc(h, &h->d_partials, (size_t)h->n_tiles * NS))) return rc;
  if ((rc = dalloc(h, &h->d_node_scal, (size_t)A * NS))) return rc;
  if ((rc = dalloc(h, &h->d_node_scal2, (size_t)A * NS))) return rc;
  if ((rc = dalloc(h, &h->d_coef, (size_t)A * MAXC))) return rc;
  if ((rc = dalloc(h, &h->d_gamma, (size_t)A))) return rc;
  if ((rc = dalloc(h, &h->d_block_partials, (size_t)148 * 8 + 8))) return rc;
  if ((rc = dalloc(h, &h->d_scalar, (size_t)40))) return rc;
  CK(cudaMallocHost((void **)&h->h_pinned, sizeof(double) * ((size_t)A * (2 * NS + 8) + 64)));   // [.. + 32, +64): all-reduce staging
  if ((rc = dalloc(h, &h->d_slot, (size_t)RED_SLOTS * A * NS))) return rc;
  CK(cudaMallocHost((void **)&h->h_slot, sizeof(double) * (size_t)RED_SLOTS * A * NS));
  std::memset(h->h_slot, 0, sizeof(double) * (size_t)RED_SLOTS * A * NS);
  CK(cudaStreamSynchronize(h->stream));
  h->st.assign(A, NodeState());
  h->graph_set = true;
  return 0;
}

int plan_halo(int64_t N, int num_nodes, int64_t E, const int32_t *ei, const int32_t *ej, int world,
              const int32_t *rnb, int rank, int64_t *sc, int64_t *rcn, int64_t *sg, int64_t scap, int64_t *rg,
              int64_t rcap) {
  if (N <= 0 || num_nodes <= 0 || world < 1 || rank < 0 || rank >= world || !rnb || !ei || !ej || !sc || !rcn ||
      rnb[0] != 0 || rnb[world] != num_nodes) {
    set_error("inconsistent sharding");
    return MMPGO_ERR_ARG;
  }
  const Partition part(N, num_nodes);
  auto owner = [&](int64_t g) {
    int nd; int64_t pp;
    part(g, nd, pp);
    int r = (int)(std::upper_bound(rnb, rnb + world + 1, nd) - rnb) - 1;
    return std::min(std::max(r, 0), world - 1);
  };
  std::vector<std::vector<int64_t>> snd(world), rcv(world);
  for (int64_t e = 0; e < E; ++e) {
    const int ri = owner(ei[e]), rj = owner(ej[e]);
    if (ri == rj) continue;
    if (ri == rank) { snd[rj].push_back(ei[e]); rcv[rj].push_back(ej[e]); }
    if (rj == rank) { snd[ri].push_back(ej[e]); rcv[ri].push_back(ei[e]); }
  }
  int64_t so = 0, ro = 0;
  for (int q = 0; q < world; ++q) {
    for (auto *v : {&snd[q], &rcv[q]}) { std::sort(v->begin(), v->end()); v->erase(std::unique(v->begin(), v->end()), v->end()); }
    sc[q] = (int64_t)snd[q].size(); rcn[q] = (int64_t)rcv[q].size();
    if (sg) { if (so + sc[q] > scap) { set_error("send_gids too small"); return MMPGO_ERR_ARG; } std::copy(snd[q].begin(), snd[q].end(), sg + so); }
    if (rg) { if (ro + rcn[q] > rcap) { set_error("recv_gids too small"); return MMPGO_ERR_ARG; } std::copy(rcv[q].begin(), rcv[q].end(), rg + ro); }
    so += sc[q]; ro += rcn[q];
  }
  return 0;
}

void plan_halo_pair(int world, const int64_t *send_poses, const int64_t *recv_poses, int64_t n_own, int32_t *sa,
                    int32_t *sb, int32_t *ra, int32_t *rb, int32_t *hrow) {
  int64_t so = 0, ro = 0;
  for (int q = 0; q < world; ++q) {
    const int64_t sc = send_poses[q], rcq = recv_poses[q];
    for (int64_t i = 0; i < sc; ++i) { sa[so + i] = (int32_t)(2 * so + i); sb[so + i] = (int32_t)(2 * so + sc + i); }
    for (int64_t i = 0; i < rcq; ++i) {
      ra[ro + i] = (int32_t)(2 * ro + i); rb[ro + i] = (int32_t)(2 * ro + rcq + i);
      hrow[ro + i] = (int32_t)(n_own + ro + i);
    }
    so += sc; ro += rcq;
  }
}

int driver_set_sharding(Handle *h, int rank, int world, const int32_t *rnb, mmpgo_exchange_fn ex,
                        mmpgo_allreduce_fn ar, void *user) {
  if (!h->graph_set) { set_error("set_graph first"); return MMPGO_ERR_STATE; }
  if (world < 1 || rank < 0 || rank >= world || !rnb || rnb[rank] != h->node_begin || rnb[rank + 1] != h->node_end ||
      rnb[0] != 0 || rnb[world] != h->num_nodes) {
    set_error("inconsistent sharding");
    return MMPGO_ERR_ARG;
  }
  if (world > 1 && ((ex == nullptr) != (ar == nullptr))) { set_error("pass both collective callbacks, or none and call mmpgo_nccl_init"); return MMPGO_ERR_ARG; }
  h->rank = rank; h->world = world; h->exchange_fn = ex; h->allreduce_fn = ar; h->cb_user = user;
  auto owner = [&](int node) {
    int r = (int)(std::upper_bound(rnb, rnb + world + 1, node) - rnb) - 1;
    return std::min(std::max(r, 0), world - 1);
  };
  const int PB = (h->d + 1) * h->d;
  // receive side: the halo is sorted by global id, hence already grouped by owner rank
  h->recv_poses.assign(world, 0);
  for (int k = 0; k < h->NH; ++k) h->recv_poses[owner(h->halo_owner[k])]++;
  // send side: own poses adjacent to a node of rank q, ascending pose id
  std::vector<std::vector<int>> lists(world);
  for (const auto &bp : h->boundary_pairs) lists[owner(bp.second)].push_back(bp.first);
  h->send_poses.assign(world, 0);
  std::vector<int> idx;
  for (int q = 0; q < world; ++q) {
    auto &l = lists[q];
    std::sort(l.begin(), l.end());
    l.erase(std::unique(l.begin(), l.end()), l.end());
    if (q == rank && !l.empty()) { set_error("internal: boundary pose sent to self"); return MMPGO_ERR_ARG; }
    h->send_poses[q] = (int64_t)l.size();
    idx.insert(idx.end(), l.begin(), l.end());
  }
  h->n_send = (int64_t)idx.size();
  h->send_dbl.resize(world); h->recv_dbl.resize(world);
  for (int q = 0; q < world; ++q) { h->send_dbl[q] = h->send_poses[q] * PB; h->recv_dbl[q] = h->recv_poses[q] * PB; }
  int rc = 0;
  if ((rc = upload(h, &h->d_send_idx, idx))) return rc;
  if ((rc = dalloc(h, &h->d_send, (size_t)h->n_send * PB))) return rc;
  // two-array exchange (AMM-PGO*: X^{k+1/2} and X^{k+1} travel in ONE all-to-all): per peer the
  // chunk is [poses of array a | poses of array b]; index maps for the pack / unpack copies
  {
    std::vector<int> sa((size_t)h->n_send), sb((size_t)h->n_send), ra((size_t)h->NH), rb((size_t)h->NH), hrow((size_t)h->NH);
    plan_halo_pair(world, h->send_poses.data(), h->recv_poses.data(), h->NO, sa.data(), sb.data(), ra.data(), rb.data(),
                   hrow.data());
    h->send_dbl2.resize(world); h->recv_dbl2.resize(world);
    for (int q = 0; q < world; ++q) { h->send_dbl2[q] = 2 * h->send_dbl[q]; h->recv_dbl2[q] = 2 * h->recv_dbl[q]; }
    if ((rc = upload(h, &h->d_send2_a, sa))) return rc;
    if ((rc = upload(h, &h->d_send2_b, sb))) return rc;
    if ((rc = upload(h, &h->d_recv2_a, ra))) return rc;
    if ((rc = upload(h, &h->d_recv2_b, rb))) return rc;
    if ((rc = upload(h, &h->d_halo_row, hrow))) return rc;
    if ((rc = dalloc(h, &h->d_send2, (size_t)2 * h->n_send * PB))) return rc;
    if ((rc = dalloc(h, &h->d_recv2, (size_t)2 * h->NH * PB))) return rc;
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

}  // namespace mmpgo
