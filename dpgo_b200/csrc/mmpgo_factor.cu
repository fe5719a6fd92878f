// Host side of the sparse direct solve (see mmpgo_mf.cuh): nested-dissection ordering,
// symbolic structure, multifrontal Cholesky and the job lists of the device sweeps.
//
// Stands where the reference calls CHOLMOD through Eigen::CholmodDecomposition
// (C++/DPGO/include/DPGO/DPGO_types.h:27; L_.compute(G00) at C++/DPGO/src/DPGOProblem.cpp:93,
// reg_Chol_precon_.compute(G11 + lambda I) at :119-123).  Nothing here is taken from SuiteSparse:
// the ordering is George's automatic nested dissection on breadth-first level structures, the
// numeric phase a textbook multifrontal method with dense fronts.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "mmpgo_mf.cuh"

namespace mmpgo {

namespace {

struct VGraph {                 // vertex graph of one node: adjacency without self loops
  int nv = 0;
  std::vector<int> ptr, adj;
};

struct NdNode {
  std::vector<int> verts;
  int child[2] = {-1, -1};
};

// ---- nested dissection ------------------------------------------------------------------
class Dissection {
 public:
  Dissection(const VGraph &g, int leaf) : g_(g), leaf_(std::max(leaf, 1)), region_(g.nv, -1), seen_(g.nv, 0), lev_(g.nv, 0) {}
  std::vector<NdNode> nodes;    // postorder: children before parents, root last

  void run() {
    std::vector<int> all(g_.nv);
    std::iota(all.begin(), all.end(), 0);
    rec(all);
  }

 private:
  const VGraph &g_;
  int leaf_;
  std::vector<int> region_, seen_, lev_;
  int stamp_ = 0, next_region_ = 0;
  std::vector<int> cur_, nxt_;

  // breadth-first level structure of the region `rid` from `root`; further components are
  // appended behind the last level, so that edges only join equal or adjacent levels
  int bfs(const std::vector<int> &S, int root, int rid, std::vector<int> &order) {
    ++stamp_;
    order.clear();
    int L = 0;
    size_t cursor = 0;
    cur_.clear();
    cur_.push_back(root);
    seen_[root] = stamp_;
    lev_[root] = 0;
    for (;;) {
      while (!cur_.empty()) {
        order.insert(order.end(), cur_.begin(), cur_.end());
        nxt_.clear();
        for (int v : cur_)
          for (int e = g_.ptr[v]; e < g_.ptr[v + 1]; ++e) {
            const int w = g_.adj[e];
            if (region_[w] == rid && seen_[w] != stamp_) { seen_[w] = stamp_; lev_[w] = L + 1; nxt_.push_back(w); }
          }
        cur_.swap(nxt_);
        ++L;
      }
      while (cursor < S.size() && seen_[S[cursor]] == stamp_) ++cursor;
      if (cursor == S.size()) break;
      const int r2 = S[cursor];
      seen_[r2] = stamp_;
      lev_[r2] = L;
      cur_.push_back(r2);
    }
    return L;
  }

  int make_leaf(std::vector<int> &S) {
    NdNode nd;
    nd.verts.swap(S);
    nodes.push_back(std::move(nd));
    return (int)nodes.size() - 1;
  }

  int rec(std::vector<int> &S) {
    if ((int)S.size() <= leaf_) return make_leaf(S);
    const int rid = next_region_++;
    for (int v : S) region_[v] = rid;
    std::vector<int> order;
    bfs(S, S[0], rid, order);
    const int far = order.back();              // pseudo-peripheral start
    const int L = bfs(S, far, rid, order);
    if (L < 3) return make_leaf(S);
    std::vector<int> cnt(L, 0), cum(L, 0);
    for (int v : order) cnt[lev_[v]]++;
    for (int l = 0, c = 0; l < L; ++l) { c += cnt[l]; cum[l] = c; }
    const int tot = (int)S.size();
    int best = -1;
    for (double thr : {0.3, 0.15, 0.0}) {
      for (int l = 1; l + 1 < L; ++l) {
        const int before = cum[l - 1], after = tot - cum[l];
        if (before <= 0 || after <= 0 || std::min(before, after) < thr * tot) continue;
        if (best < 0 || cnt[l] < cnt[best] ||
            (cnt[l] == cnt[best] && std::abs(2 * cum[l] - tot) < std::abs(2 * cum[best] - tot))) best = l;
      }
      if (best >= 0) break;
    }
    if (best < 0) return make_leaf(S);
    std::vector<int> A, B, sep;
    for (int v : order) {
      const int l = lev_[v];
      if (l < best) A.push_back(v);
      else if (l > best) B.push_back(v);
      else {
        // only the vertices of the level that touch the next one separate the halves
        bool touches = false;
        for (int e = g_.ptr[v]; e < g_.ptr[v + 1] && !touches; ++e) {
          const int w = g_.adj[e];
          touches = region_[w] == rid && lev_[w] == best + 1;
        }
        (touches ? sep : A).push_back(v);
      }
    }
    if (sep.empty()) { sep.push_back(A.back()); A.pop_back(); }   // halves already disconnected
    if (A.empty() || B.empty()) {
      S.clear(); S.insert(S.end(), order.begin(), order.end());
      return make_leaf(S);
    }
    std::vector<int>().swap(order);
    std::vector<int>().swap(S);
    const int ca = rec(A);
    const int cb = rec(B);
    NdNode nd;
    nd.verts.swap(sep);
    nd.child[0] = ca; nd.child[1] = cb;
    nodes.push_back(std::move(nd));
    return (int)nodes.size() - 1;
  }
};

// ---- one node ---------------------------------------------------------------------------
struct NodeFactor {
  int n = 0;
  std::vector<int> sperm, siperm;                 // scalar permutation (local)
  std::vector<int> c0, k, m, child0, child1, height, depth;
  std::vector<std::vector<int>> bnd;              // boundary, scalar permuted indices (local)
  std::vector<size_t> moff, mtoff;                // column-major copy (k columns of Rp) / row-major copy (R rows of kp)
  std::vector<double> M, MT;
  int64_t nnz = 0;
  double flops = 0.0;
  bool ok = true;
};

void factor_node(const MfMatrix &A, int block, int leaf, bool symbolic_only, NodeFactor &nf) {
  const int n = A.n, nv = n / block;
  nf.n = n;
  if (A.skip) {
    nf.siperm.resize(n);
    std::iota(nf.siperm.begin(), nf.siperm.end(), 0);
    nf.moff.assign(1, 0); nf.mtoff.assign(1, 0);
    return;
  }
  // vertex graph
  VGraph g;
  g.nv = nv;
  g.ptr.assign(nv + 1, 0);
  for (int r = 0; r < n; r += block)
    for (int e = A.ptr[r]; e < A.ptr[r + 1]; ++e)
      if (A.col[e] / block != r / block) g.ptr[r / block + 1]++;
  for (int v = 0; v < nv; ++v) g.ptr[v + 1] += g.ptr[v];
  g.adj.resize(g.ptr[nv]);
  {
    std::vector<int> fill(g.ptr.begin(), g.ptr.end() - 1);
    for (int r = 0; r < n; r += block)
      for (int e = A.ptr[r]; e < A.ptr[r + 1]; ++e)
        if (A.col[e] / block != r / block) g.adj[fill[r / block]++] = A.col[e] / block;
    // ascending neighbour lists: the ordering depends on the graph only, not on the order of its edges
    for (int v = 0; v < nv; ++v) std::sort(g.adj.begin() + g.ptr[v], g.adj.begin() + g.ptr[v + 1]);
  }
  Dissection nd(g, leaf);
  nd.run();
  const int ns = (int)nd.nodes.size();
  // permutation: supernodes in postorder, vertices in the order the dissection left them
  std::vector<int> vperm(nv, -1);
  nf.c0.resize(ns); nf.k.resize(ns); nf.m.resize(ns); nf.child0.resize(ns); nf.child1.resize(ns);
  nf.height.assign(ns, 0); nf.depth.assign(ns, 0);
  int pos = 0;
  for (int s = 0; s < ns; ++s) {
    nf.c0[s] = pos * block;
    nf.k[s] = (int)nd.nodes[s].verts.size() * block;
    nf.child0[s] = nd.nodes[s].child[0]; nf.child1[s] = nd.nodes[s].child[1];
    for (int v : nd.nodes[s].verts) vperm[v] = pos++;
  }
  nf.sperm.resize(n); nf.siperm.resize(n);
  for (int v = 0; v < nv; ++v)
    for (int r = 0; r < block; ++r) { nf.sperm[v * block + r] = vperm[v] * block + r; nf.siperm[vperm[v] * block + r] = v * block + r; }
  // symbolic: boundary of every supernode (vertex level, then expanded)
  std::vector<std::vector<int>> vb(ns);
  nf.bnd.resize(ns);
  for (int s = 0; s < ns; ++s) {
    const int c1 = (nf.c0[s] + nf.k[s]) / block;
    std::vector<int> &b = vb[s];
    for (int v : nd.nodes[s].verts)
      for (int e = g.ptr[v]; e < g.ptr[v + 1]; ++e)
        if (vperm[g.adj[e]] >= c1) b.push_back(vperm[g.adj[e]]);
    for (int c : {nf.child0[s], nf.child1[s]})
      if (c >= 0) {
        for (int x : vb[c]) if (x >= c1) b.push_back(x);
        nf.height[s] = std::max(nf.height[s], nf.height[c] + 1);
      }
    std::sort(b.begin(), b.end());
    b.erase(std::unique(b.begin(), b.end()), b.end());
    nf.m[s] = (int)b.size() * block;
    nf.bnd[s].resize(nf.m[s]);
    for (size_t i = 0; i < b.size(); ++i)
      for (int r = 0; r < block; ++r) nf.bnd[s][i * block + r] = b[i] * block + r;
    const double k = nf.k[s], R = nf.k[s] + nf.m[s];
    nf.nnz += (int64_t)(k * (k + 1) / 2 + k * nf.m[s]);
    nf.flops += k * R * R;
  }
  for (int s = ns - 1; s >= 0; --s)
    for (int c : {nf.child0[s], nf.child1[s]}) if (c >= 0) nf.depth[c] = nf.depth[s] + 1;
  nf.moff.assign(ns + 1, 0); nf.mtoff.assign(ns + 1, 0);
  for (int s = 0; s < ns; ++s) {
    nf.moff[s + 1] = nf.moff[s] + (size_t)nf.k[s] * mf_even(nf.k[s] + nf.m[s]);
    nf.mtoff[s + 1] = nf.mtoff[s] + (size_t)(nf.k[s] + nf.m[s]) * mf_even(nf.k[s]);
  }
  if (symbolic_only) return;

  // numeric: multifrontal Cholesky, dense fronts (column-major, lower triangle)
  nf.M.assign(nf.moff[ns], 0.0);
  nf.MT.assign(nf.mtoff[ns], 0.0);
  std::vector<int> fpos(n, -1);
  std::vector<std::vector<double>> upd(ns);
  std::vector<double> F, X;
  for (int s = 0; s < ns && nf.ok; ++s) {
    const int k = nf.k[s], m = nf.m[s], R = k + m, c0 = nf.c0[s];
    F.assign((size_t)R * R, 0.0);
    for (int t = 0; t < k; ++t) fpos[c0 + t] = t;
    for (int t = 0; t < m; ++t) fpos[nf.bnd[s][t]] = k + t;
    for (int j = 0; j < k; ++j) {
      const int r = nf.siperm[c0 + j];
      for (int e = A.ptr[r]; e < A.ptr[r + 1]; ++e) {
        const int pc = nf.sperm[A.col[e]];
        if (pc < c0) continue;
        const int i = fpos[pc];
        if (i >= j) F[(size_t)i + (size_t)j * R] += A.val[e];
      }
    }
    for (int c : {nf.child0[s], nf.child1[s]})
      if (c >= 0) {
        const int mc = nf.m[c];
        const std::vector<double> &U = upd[c];
        const std::vector<int> &bc = nf.bnd[c];
        for (int jj = 0; jj < mc; ++jj) {
          const size_t pj = (size_t)fpos[bc[jj]];
          for (int ii = jj; ii < mc; ++ii) F[(size_t)fpos[bc[ii]] + pj * R] += U[(size_t)ii + (size_t)jj * mc];
        }
        std::vector<double>().swap(upd[c]);
      }
    // partial Cholesky of the leading k columns, right-looking
    for (int j = 0; j < k; ++j) {
      double *cj = &F[(size_t)j * R];
      const double dj = cj[j];
      if (!(dj > 0.0)) { nf.ok = false; break; }
      const double l = std::sqrt(dj), inv = 1.0 / l;
      cj[j] = l;
      for (int i = j + 1; i < R; ++i) cj[i] *= inv;
      for (int jj = j + 1; jj < R; ++jj) {
        const double c = cj[jj];
        if (c == 0.0) continue;
        double *cc = &F[(size_t)jj * R];
        for (int i = jj; i < R; ++i) cc[i] -= cj[i] * c;
      }
    }
    if (!nf.ok) break;
    // X = inv(L11), lower triangular, column-major k x k
    X.assign((size_t)k * k, 0.0);
    for (int c = 0; c < k; ++c) {
      X[(size_t)c + (size_t)c * k] = 1.0 / F[(size_t)c + (size_t)c * R];
      for (int i = c + 1; i < k; ++i) {
        double sum = 0.0;
        for (int t = c; t < i; ++t) sum += F[(size_t)i + (size_t)t * R] * X[(size_t)t + (size_t)c * k];
        X[(size_t)i + (size_t)c * k] = -sum / F[(size_t)i + (size_t)i * R];
      }
    }
    double *Ms = &nf.M[nf.moff[s]], *Mt = &nf.MT[nf.mtoff[s]];
    const size_t Rp = (size_t)mf_even(R), kp = (size_t)mf_even(k);
    for (int c = 0; c < k; ++c) {
      for (int i = c; i < k; ++i) Ms[(size_t)i + (size_t)c * Rp] = X[(size_t)i + (size_t)c * k];
      for (int i = 0; i < m; ++i) {
        // W = L21 inv(L11):  W[i][c] = sum_{t >= c} L21[i][t] X[t][c]
        double sum = 0.0;
        for (int t = c; t < k; ++t) sum += F[(size_t)(k + i) + (size_t)t * R] * X[(size_t)t + (size_t)c * k];
        Ms[(size_t)(k + i) + (size_t)c * Rp] = sum;
      }
    }
    for (int i = 0; i < R; ++i)
      for (int c = 0; c < k; ++c) Mt[(size_t)i * kp + c] = Ms[(size_t)i + (size_t)c * Rp];
    if (m > 0) {
      std::vector<double> &U = upd[s];
      U.resize((size_t)m * m);
      for (int jj = 0; jj < m; ++jj)
        for (int ii = jj; ii < m; ++ii) U[(size_t)ii + (size_t)jj * m] = F[(size_t)(k + ii) + (size_t)(k + jj) * R];
    }
  }
}

}  // namespace

int mf_factor(const std::vector<MfMatrix> &mats, int block, int leaf, bool symbolic_only, MfFactor *out) {
  const int A = (int)mats.size();
  std::vector<NodeFactor> nfs(A);
#pragma omp parallel for schedule(dynamic, 1)
  for (int a = 0; a < A; ++a) factor_node(mats[a], block, leaf, symbolic_only, nfs[a]);
  MfFactor &F = *out;
  F = MfFactor();
  F.block = block;
  size_t tot_m = 0, tot_mt = 0, tot_rows = 0, tot_b = 0;
  int row_off = 0;
  for (int a = 0; a < A; ++a) {
    if (!nfs[a].ok) return -1;
    F.nnz += nfs[a].nnz; F.flops += nfs[a].flops;
    tot_m += nfs[a].moff.back(); tot_mt += nfs[a].mtoff.back();
    for (size_t s = 0; s < nfs[a].k.size(); ++s) { tot_rows += nfs[a].k[s] + nfs[a].m[s]; tot_b += nfs[a].m[s]; }
    row_off += nfs[a].n;
  }
  F.nrows = row_off;
  if (symbolic_only) {
    for (int a = 0; a < A; ++a)
      for (size_t s = 0; s < nfs[a].k.size(); ++s) F.height = std::max(F.height, nfs[a].height[s]);
    return 0;
  }
  F.M.resize(tot_m); F.MT.resize(tot_mt);
  F.pull0.assign(tot_rows, -1); F.pull1.assign(tot_rows, -1);
  F.bidx.resize(tot_b);
  F.iperm.resize(F.nrows);
  std::vector<std::vector<int>> fw, bw;        // supernodes per forward / backward stage
  size_t mo = 0, mto = 0, ro = 0, bo = 0;
  int uo = 0;
  row_off = 0;
  for (int a = 0; a < A; ++a) {
    NodeFactor &nf = nfs[a];
    const int ns = (int)nf.k.size(), base = (int)F.sn.size();
    if (!nf.M.empty()) {
      std::memcpy(&F.M[mo], nf.M.data(), nf.M.size() * sizeof(double));
      std::memcpy(&F.MT[mto], nf.MT.data(), nf.MT.size() * sizeof(double));
    }
    for (int p = 0; p < nf.n; ++p) F.iperm[row_off + p] = row_off + nf.siperm[p];
    std::vector<int> uoff(ns);
    for (int s = 0; s < ns; ++s) {
      MfSn sn;
      std::memset(&sn, 0, sizeof(sn));
      const int k = nf.k[s], m = nf.m[s], R = k + m;
      sn.node = a; sn.c0 = row_off + nf.c0[s]; sn.k = k; sn.R = R;
      sn.moff = (long long)(mo + nf.moff[s]); sn.mtoff = (long long)(mto + nf.mtoff[s]);
      sn.rowoff = (int)ro; sn.boff = (int)bo; sn.uoff = uo;
      sn.nchild = (nf.child0[s] >= 0) + (nf.child1[s] >= 0);
      uoff[s] = uo;
      MfFactor::Dep dp;
      std::memset(&dp, 0, sizeof(dp));
      dp.child0 = nf.child0[s] >= 0 ? base + nf.child0[s] : -1;
      dp.child1 = nf.child1[s] >= 0 ? base + nf.child1[s] : -1;
      dp.parent = -1;
      F.dep.push_back(dp);
      for (int t = 0; t < m; ++t) F.bidx[bo + t] = row_off + nf.bnd[s][t];
      // children's update rows land on rows of this front
      int slot = 0;
      for (int c : {nf.child0[s], nf.child1[s]}) {
        if (c >= 0) {
          std::vector<int> &pl = slot == 0 ? F.pull0 : F.pull1;
          for (int t = 0; t < nf.m[c]; ++t) {
            const int p = nf.bnd[c][t];        // local permuted index: a column or a boundary row of s
            int fp;
            if (p >= nf.c0[s] && p < nf.c0[s] + k) fp = p - nf.c0[s];
            else fp = k + (int)(std::lower_bound(nf.bnd[s].begin(), nf.bnd[s].end(), p) - nf.bnd[s].begin());
            pl[ro + fp] = uoff[c] + t;
          }
        }
        ++slot;
      }
      F.sn.push_back(sn);
      if ((int)fw.size() <= nf.height[s]) fw.resize(nf.height[s] + 1);
      if ((int)bw.size() <= nf.depth[s]) bw.resize(nf.depth[s] + 1);
      fw[nf.height[s]].push_back(base + s);
      bw[nf.depth[s]].push_back(base + s);
      F.height = std::max(F.height, nf.height[s]);
      ro += R; bo += m; uo += m;
    }
    for (int s = 0; s < ns; ++s)
      for (int c : {nf.child0[s], nf.child1[s]}) if (c >= 0) F.dep[base + c].parent = base + s;
    mo += nf.M.size(); mto += nf.MT.size();
    row_off += nf.n;
    nf = NodeFactor();     // release the node's copy
  }
  F.urows = uo;
  F.perm.resize(F.nrows);
  for (int p = 0; p < F.nrows; ++p) F.perm[F.iperm[p]] = p;
  // warp jobs.  Forward: fronts with at most MF_KS columns are cut into jobs of 64 rows (whole front if it has at
  // most 128), wider ones into jobs of 2 MF_SROWS rows summed in slices.  Backward: fronts with at most MF_RS rows
  // are one job, taller ones are cut into jobs of 2 MF_SROWS columns summed in slices.
  auto mkjob = [](int sn, int r0, int n) { MfJob j; std::memset(&j, 0, sizeof(j)); j.sn = sn; j.r0 = r0; j.n = n; j.wait0 = j.wait1 = -1; return j; };
  for (int dir = 0; dir < 2; ++dir) {
    const std::vector<std::vector<int>> &stages = dir == 0 ? fw : bw;
    F.wstage[dir].assign(1, 0);
    for (const auto &st : stages) {
      for (int s : st) {
        const MfSn &sn = F.sn[s];
        F.max_R = std::max(F.max_R, sn.R);
        if (dir == 0) {
          if (sn.k > MF_KS) for (int r0 = 0; r0 < sn.R; r0 += 2 * MF_SROWS) F.wjobs[0].push_back(mkjob(s, r0, std::min(2 * MF_SROWS, sn.R - r0)));
          else if (sn.R <= 128) F.wjobs[0].push_back(mkjob(s, 0, sn.R));
          else for (int r0 = 0; r0 < sn.R; r0 += 64) F.wjobs[0].push_back(mkjob(s, r0, std::min(64, sn.R - r0)));
        } else {
          if (sn.R > MF_RS) for (int c0 = 0; c0 < sn.k; c0 += 2 * MF_SROWS) F.wjobs[1].push_back(mkjob(s, c0, std::min(2 * MF_SROWS, sn.k - c0)));
          else F.wjobs[1].push_back(mkjob(s, 0, sn.k));
        }
      }
      F.wstage[dir].push_back((int)F.wjobs[dir].size());
    }
    for (const MfJob &jb : F.wjobs[dir]) (dir == 0 ? F.dep[jb.sn].nf : F.dep[jb.sn].nb)++;
  }
  // what every job waits for: counters of MfDevice::done ([0, nsn) forward, [nsn, 2 nsn) backward) and their targets
  const int nsn = (int)F.sn.size();
  for (MfJob &jb : F.wjobs[0]) {
    const MfFactor::Dep &dp = F.dep[jb.sn];
    if (dp.child0 >= 0) { jb.wait0 = dp.child0; jb.need0 = F.dep[dp.child0].nf; }
    if (dp.child1 >= 0) { jb.wait1 = dp.child1; jb.need1 = F.dep[dp.child1].nf; }
  }
  for (MfJob &jb : F.wjobs[1]) {
    const MfFactor::Dep &dp = F.dep[jb.sn];
    if (dp.parent >= 0) { jb.wait0 = nsn + dp.parent; jb.need0 = F.dep[dp.parent].nb; }
    else { jb.wait0 = jb.sn; jb.need0 = dp.nf; }           // the root turns around when its forward jobs are done
  }
  return 0;
}

// Host restatement of the device sweeps: same blocks, same order of every sum (fused multiply-adds
// in ascending column / row order; CTA jobs in MF_Q contiguous slices added in slice order), so the
// device result can be compared bit for bit.
void mf_host_solve(const MfFactor &F, int nrhs, const double *rhs, double *x) {
  const int D = nrhs;
  std::vector<double> y((size_t)F.nrows * D, 0.0), xp((size_t)F.nrows * D, 0.0), u((size_t)std::max(F.urows, 1) * D, 0.0);
  std::vector<double> f;
  auto stage_rows = [&](const MfSn &sn, int i, double *val) {   // rhs + children's update rows of front row i
    for (int c = 0; c < D; ++c) val[c] = i < sn.k ? rhs[(size_t)F.iperm[sn.c0 + i] * D + c] : 0.0;
    if (!sn.nchild) return;
    const int p0 = F.pull0[sn.rowoff + i], p1 = F.pull1[sn.rowoff + i];
    for (int c = 0; c < D; ++c) {
      if (p0 >= 0) val[c] += u[(size_t)p0 * D + c];
      if (p1 >= 0) val[c] += u[(size_t)p1 * D + c];
    }
  };
  // sum_{j in [b, e)} Mp[j] * fv[j] per right-hand side, fused multiply-adds in ascending order: one chain, or
  // (sliced fronts) MF_Q chains -- term j belongs to slice ((j - base) % MF_BLK) / MF_QW -- added in slice order
  auto dot = [&](bool sliced, int base, int b, int e, const double *Mp, const double *fv, double *acc) {
    double part[MF_Q][8];
    for (int q = 0; q < MF_Q; ++q) for (int c = 0; c < D; ++c) part[q][c] = 0.0;
    for (int j = b; j < e; ++j) {
      const int q = sliced ? ((j - base) % MF_BLK) / MF_QW : 0;
      for (int c = 0; c < D; ++c) part[q][c] = std::fma(Mp[j], fv[(size_t)j * D + c], part[q][c]);
    }
    for (int c = 0; c < D; ++c) {
      acc[c] = part[0][c];
      if (sliced) for (int q = 1; q < MF_Q; ++q) acc[c] += part[q][c];
    }
  };
  for (size_t st = 0; st + 1 < F.wstage[0].size(); ++st)
    for (int t = F.wstage[0][st]; t < F.wstage[0][st + 1]; ++t) {
      const MfJob &jb = F.wjobs[0][t];
      const MfSn &sn = F.sn[jb.sn];
      const int k = sn.k, kp = mf_even(k);
      f.assign((size_t)kp * D, 0.0);
      for (int j = 0; j < k; ++j) stage_rows(sn, j, &f[(size_t)j * D]);
      const double *Mt = &F.MT[sn.mtoff];
      for (int i = jb.r0; i < jb.r0 + jb.n; ++i) {
        const int len = i < k ? i + 1 : k;                            // inv(L11) is lower triangular
        double acc[8];
        dot(k > MF_KS, 0, 0, len, Mt + (size_t)i * kp, f.data(), acc);
        if (i < k) {
          for (int c = 0; c < D; ++c) y[(size_t)(sn.c0 + i) * D + c] = acc[c];
        } else {
          double f2[8];
          stage_rows(sn, i, f2);
          for (int c = 0; c < D; ++c) u[(size_t)(sn.uoff + i - k) * D + c] = f2[c] - acc[c];
        }
      }
    }
  for (size_t st = 0; st + 1 < F.wstage[1].size(); ++st)
    for (int t = F.wstage[1][st]; t < F.wstage[1][st + 1]; ++t) {
      const MfJob &jb = F.wjobs[1][t];
      const MfSn &sn = F.sn[jb.sn];
      const int k = sn.k, R = sn.R, Rp = mf_even(R);
      f.assign((size_t)Rp * D, 0.0);
      for (int i = 0; i < R; ++i)
        for (int c = 0; c < D; ++c)
          f[(size_t)i * D + c] = i < k ? y[(size_t)(sn.c0 + i) * D + c] : -xp[(size_t)F.bidx[sn.boff + i - k] * D + c];
      const double *Ms = &F.M[sn.moff];
      const bool sliced = R > MF_RS;
      for (int j = jb.r0; j < jb.r0 + jb.n; ++j) {
        // rows before the job's (pass's) first column multiply zeros of M_s^T; sliced jobs start their blocks there
        const int b = sliced ? jb.r0 : ((j - jb.r0) / 64) * 64 + jb.r0;
        double acc[8];
        dot(sliced, b, b, R, Ms + (size_t)j * Rp, f.data(), acc);
        for (int c = 0; c < D; ++c) {
          xp[(size_t)(sn.c0 + j) * D + c] = acc[c];
          x[(size_t)F.iperm[sn.c0 + j] * D + c] = acc[c];
        }
      }
    }
}

}  // namespace mmpgo
