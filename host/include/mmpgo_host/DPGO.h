// Host-side mirror of the reference's per-node driver interface, in the reference's language.
//
// Keeps the names, argument meaning and int return codes of
//   DPGO::RelativePoseMeasurement / measurements_t   C++/DPGO/include/DPGO/RelativePoseMeasurement.h:11-62
//   DPGO::Options, Loss, Scheme                      C++/DPGO/include/DPGO/DPGO_types.h:35-201
//   DPGO::read_g2o_file                              C++/DPGO/src/DPGO_utils.cpp:8-138
//   DPGO::DPGOHash  (initialize/update/iterate/communicate/results)   include/DPGO/DPGOHash.h:13-107
//   DPGO::DPGOStar  (+ evaluate_f, evaluate_grad)                                    include/DPGO/DPGOStar.h:13-93
// but one object drives ALL robot nodes [node_begin, node_end) that live on one GPU: the loops
// `for alpha: dpgo_hash[alpha]->iterate()` of C++/examples/dist_pgo.cpp:497-520 become one
// batched call through the C ABI (include/mmpgo.h).  All arithmetic happens in libmmpgo.so; there
// is no CPU fallback.  Matrices are plain column-major buffers in the reference's global layout
// [t (N rows); R blocks (dN rows)] x d (dist_pgo.cpp:502-511) -- what Eigen::MatrixXd::data() is.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <map>
#include <set>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" {
#include "../../../include/mmpgo.h"
}

namespace DPGO {

typedef double Scalar;

enum class Loss { None = 0, Huber = 1, GemanMcClure = 2, Welsch = 3 };   // DPGO_types.h:67
enum class Scheme { MM = 0, AMM = 1 };                                   // DPGO_types.h:70-75
enum class Preconditioner { None = 0, Jacobi = 1, BlockJacobi = 2, RegularizedCholesky = 3 };   // DPGO_types.h:35-40 + the device path's block-Jacobi
enum class Rescale { Static = 0, Dynamic = 1 };                          // DPGO_types.h:42-46

/** One relative pose measurement i -> j (RelativePoseMeasurement.h:11-29); poses are GLOBAL ids,
 * the node of a pose follows from the contiguous id-range partition (DPGO_utils.cpp:147-158). */
struct RelativePoseMeasurement {
  int64_t i = 0, j = 0;
  std::array<double, 9> R{};   // row-major d x d
  std::array<double, 3> t{};
  double kappa = 0, tau = 0;
};
typedef std::vector<RelativePoseMeasurement> measurements_t;

/** A column-major matrix: the subset of Eigen::MatrixXd this interface needs. */
struct Matrix {
  int64_t r = 0, c = 0;
  std::vector<double> v;
  Matrix() {}
  Matrix(int64_t rows, int64_t cols) : r(rows), c(cols), v((size_t)rows * cols, 0.0) {}
  int64_t rows() const { return r; }
  int64_t cols() const { return c; }
  double *data() { return v.data(); }
  const double *data() const { return v.data(); }
  double &operator()(int64_t i, int64_t j) { return v[(size_t)j * r + i]; }
  double operator()(int64_t i, int64_t j) const { return v[(size_t)j * r + i]; }
};

/** DPGO::Options with the values dist_pgo sets (dist_pgo.cpp:103-120). */
struct Options {
  Loss loss = Loss::None;
  Scheme scheme = Scheme::AMM;
  Preconditioner preconditioner = Preconditioner::BlockJacobi;
  Scalar regularizer = 1e-11, loss_reg = 0.25;
  Scalar reg_Cholesky_precon_max_condition_number = 1e6;   // DPGO_types.h:159
  Rescale rescale = Rescale::Static;     // what dist_pgo sets (dist_pgo.cpp:105); the struct's own default is Dynamic (:128)
  int max_rescale_count = 5;             // DPGO_types.h:131
  int device = 0;
  void fill(mmpgo_options *c, int algorithm) const {
    mmpgo_default_options(c);
    c->algorithm = algorithm;
    c->loss = (int)loss; c->scheme = (int)scheme; c->preconditioner = (int)preconditioner;
    c->regularizer = regularizer; c->loss_reg = loss_reg; c->device = device;
    c->rescale = (int)rescale; c->max_rescale_count = max_rescale_count;
    c->reg_Cholesky_precon_max_condition_number = reg_Cholesky_precon_max_condition_number;
  }
};

namespace detail {
inline bool inv3(const double *M, double *out) {   // symmetric 3 x 3 inverse
  const double a = M[0], b = M[1], c = M[2], d = M[4], e = M[5], f = M[8];
  const double det = a * (d * f - e * e) - b * (b * f - c * e) + c * (b * e - c * d);
  if (det == 0) return false;
  out[0] = (d * f - e * e) / det; out[1] = (c * e - b * f) / det; out[2] = (b * e - c * d) / det;
  out[3] = out[1]; out[4] = (a * f - c * c) / det; out[5] = (b * c - a * e) / det;
  out[6] = out[2]; out[7] = out[5]; out[8] = (a * d - b * b) / det;
  return true;
}
}  // namespace detail

/** DPGO::read_g2o_file (DPGO_utils.cpp:8-138): EDGE_SE2 / EDGE_SE3:QUAT lines; tau = d / tr(I_t^-1),
 * kappa = I_33 (2-D) or 3 / (2 tr(I_R^-1)) (3-D); VERTEX lines are ignored.  Returns d. */
inline int read_g2o_file(const std::string &filename, int64_t &num_poses, measurements_t &measurements) {
  std::ifstream in(filename);
  if (!in.is_open()) throw std::runtime_error("cannot open " + filename);
  measurements.clear();
  num_poses = 0;
  int d = 0;
  std::string line, tag;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    if (!(ss >> tag)) continue;
    RelativePoseMeasurement m;
    if (tag == "EDGE_SE2") {
      double dx, dy, dth, I11, I12, I13, I22, I23, I33;
      ss >> m.i >> m.j >> dx >> dy >> dth >> I11 >> I12 >> I13 >> I22 >> I23 >> I33;
      d = 2;
      m.t = {dx, dy, 0};
      m.R = {std::cos(dth), -std::sin(dth), std::sin(dth), std::cos(dth)};
      const double det = I11 * I22 - I12 * I12;
      m.tau = 2.0 / ((I22 + I11) / det);       // 2 / tr(TranInfo^-1)
      m.kappa = I33;
    } else if (tag == "EDGE_SE3:QUAT") {
      double dx, dy, dz, qx, qy, qz, qw, I[21];
      ss >> m.i >> m.j >> dx >> dy >> dz >> qx >> qy >> qz >> qw;
      for (double &x : I) ss >> x;
      d = 3;
      m.t = {dx, dy, dz};
      // Eigen::Quaterniond(qw,qx,qy,qz).toRotationMatrix(), no normalisation (DPGO_utils.cpp:100-101)
      const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
      m.R = {1 - (ty * qy + tz * qz), ty * qx - tz * qw, tz * qx + ty * qw,
             ty * qx + tz * qw, 1 - (tx * qx + tz * qz), tz * qy - tx * qw,
             tz * qx - ty * qw, tz * qy + tx * qw, 1 - (tx * qx + ty * qy)};
      const double It[9] = {I[0], I[1], I[2], I[1], I[6], I[7], I[2], I[7], I[11]};
      const double Ir[9] = {I[15], I[16], I[17], I[16], I[18], I[19], I[17], I[19], I[20]};
      double inv[9];
      if (!detail::inv3(It, inv)) throw std::runtime_error("singular translational information matrix");
      m.tau = 3.0 / (inv[0] + inv[4] + inv[8]);
      if (!detail::inv3(Ir, inv)) throw std::runtime_error("singular rotational information matrix");
      m.kappa = 3.0 / (2.0 * (inv[0] + inv[4] + inv[8]));
    } else if (tag.rfind("VERTEX", 0) == 0) {
      continue;
    } else {
      throw std::runtime_error("unrecognized g2o record: " + tag);
    }
    if (ss.fail()) throw std::runtime_error("malformed g2o line: " + line);
    if (m.i + 1 > num_poses) num_poses = m.i + 1;
    if (m.j + 1 > num_poses) num_poses = m.j + 1;
    measurements.push_back(m);
  }
  return d;
}

/** Scalars of DPGOResult a caller of results() reads (DPGO_types.h:204-322). */
typedef mmpgo_node_scalars DPGOResult;

class DPGODriver {
 public:
  virtual ~DPGODriver() { if (h_) mmpgo_destroy(h_); }
  DPGODriver(const DPGODriver &) = delete;
  DPGODriver &operator=(const DPGODriver &) = delete;

  /** X: global iterate ((d+1)N x d); neighbour copies are taken from X as DPGO::communicate does
   * before the first update (dist_pgo.cpp:446-462).  Returns 0, or -1 on inconsistent sizes. */
  int initialize(const Matrix &X) const {
    if (X.rows() != (int64_t)(d_ + 1) * N_ || X.cols() != d_) return -1;
    return mmpgo_initialize(h_, X.data(), X.rows());
  }
  int update() const { return mmpgo_update(h_); }
  int iterate() const { return mmpgo_iterate(h_); }
  int communicate() const { return mmpgo_communicate(h_); }
  /** results().Xk of every local node, written into the global X (rows of local nodes only). */
  int X(Matrix &X) const {
    if (X.rows() != (int64_t)(d_ + 1) * N_ || X.cols() != d_) return -1;
    return mmpgo_get_poses(h_, X.data(), X.rows());
  }
  int results(int node, DPGOResult &out) const { return mmpgo_get_node_scalars(h_, node, &out); }
  /** DPGOStar::evaluate_f (DPGOStar.cpp:713-761) over the edges owned by the local nodes. */
  int evaluate_f(const Matrix &X, Scalar &fobj) const { return mmpgo_evaluate_f(h_, X.data(), X.rows(), &fobj); }
  /** DPGOStar::evaluate_grad (DPGOStar.cpp:763-829): Riemannian gradient of the global objective at X
   *  (rows of the poses owned by the local nodes). */
  int evaluate_grad(const Matrix &X, Matrix &grad) const {
    if (X.rows() != (int64_t)(d_ + 1) * N_ || X.cols() != d_) return -1;
    if (grad.rows() != X.rows() || grad.cols() != X.cols()) grad = Matrix(X.rows(), X.cols());
    return mmpgo_evaluate_grad(h_, X.data(), X.rows(), grad.data(), grad.rows());
  }
  /** F and |grad F|^2 of the current device iterate (what dist_pgo logs, dist_pgo.cpp:523-530). */
  int objective(Scalar &F, Scalar &grad_sqnorm) const { return mmpgo_current_objective(h_, &F, &grad_sqnorm); }
  int weights(int node, std::vector<double> &w) const {
    int64_t n = 0;
    int rc = mmpgo_get_weights(h_, node, nullptr, 0, &n);
    if (rc) return rc;
    w.assign((size_t)n, 0.0);
    return n ? mmpgo_get_weights(h_, node, w.data(), n, &n) : 0;
  }
  int d() const { return d_; }
  int64_t num_poses() const { return N_; }
  mmpgo_handle handle() const { return h_; }

 protected:
  DPGODriver(int algorithm, int num_nodes, int d, int64_t num_poses, const measurements_t &meas, const Options &opt,
             int node_begin, int node_end)
      : d_(d), N_(num_poses) {
    mmpgo_options c;
    opt.fill(&c, algorithm);
    if (mmpgo_create(&c, &h_)) throw std::runtime_error(mmpgo_last_error());
    const size_t E = meas.size();
    std::vector<int32_t> ei(E), ej(E);
    std::vector<double> R(E * d * d), t(E * d), kappa(E), tau(E);
    for (size_t e = 0; e < E; ++e) {
      ei[e] = (int32_t)meas[e].i; ej[e] = (int32_t)meas[e].j;
      for (int k = 0; k < d * d; ++k) R[e * d * d + k] = meas[e].R[k];
      for (int k = 0; k < d; ++k) t[e * d + k] = meas[e].t[k];
      kappa[e] = meas[e].kappa; tau[e] = meas[e].tau;
    }
    if (node_end < 0) node_end = num_nodes;
    if (mmpgo_set_graph(h_, d, num_poses, num_nodes, node_begin, node_end, (int64_t)E, ei.data(), ej.data(), R.data(),
                        t.data(), kappa.data(), tau.data())) {
      const std::string msg = mmpgo_last_error();
      mmpgo_destroy(h_);
      h_ = nullptr;
      throw std::runtime_error(msg);
    }
  }
  mmpgo_handle h_ = nullptr;
  int d_;
  int64_t N_;
};

/** AMM-PGO# (Scheme::AMM) / MM-PGO (Scheme::MM), decentralised restarts (DPGOHash.h:13-107). */
class DPGOHash : public DPGODriver {
 public:
  DPGOHash(int num_nodes, int d, int64_t num_poses, const measurements_t &meas, const Options &opt,
           int node_begin = 0, int node_end = -1)
      : DPGODriver(MMPGO_ALG_HASH, num_nodes, d, num_poses, meas, opt, node_begin, node_end) {}
};

/** AMM-PGO*, restart decided on the global objective by the master node (DPGOStar.h:13-93). */
class DPGOStar : public DPGODriver {
 public:
  DPGOStar(int num_nodes, int d, int64_t num_poses, const measurements_t &meas, const Options &opt,
           int node_begin = 0, int node_end = -1)
      : DPGODriver(MMPGO_ALG_STAR, num_nodes, d, num_poses, meas, opt, node_begin, node_end) {}
  DPGOStar(int num_nodes, const std::string &g2o, const Options &opt) : DPGOStar(load(g2o), num_nodes, opt) {}
  int star_objective(Scalar &F, Scalar &fobj, int &restarts) const {
    int32_t r = 0;
    int rc = mmpgo_star_objective(h_, &F, &fobj, &r);
    restarts = r;
    return rc;
  }

 private:
  struct Loaded { int d; int64_t N; measurements_t m; };
  static Loaded load(const std::string &f) { Loaded l; l.d = read_g2o_file(f, l.N, l.m); return l; }
  DPGOStar(const Loaded &l, int num_nodes, const Options &opt)
      : DPGODriver(MMPGO_ALG_STAR, num_nodes, l.d, l.N, l.m, opt, 0, -1) {}
};

// ---------------------------------------------------------------------------------------------
// Per-node facade: the reference's call sequence, unchanged.
//
// The reference builds ONE driver object per node (`dpgo_hash[alpha] = make_shared<DPGOHash>(alpha, ...)`,
// dist_pgo.cpp:129-135) and loops over them for every phase (dist_pgo.cpp:446-531):
//     for alpha: initialize(Xk[alpha]); update();
//     loop { for alpha: iterate();   for alpha: X <- results().Xk;   for alpha: communicate(dpgo_hash);
//            for alpha: update(); }
// PerNode<Driver> keeps those names and that sequence on top of ONE batched driver per GPU.  Every
// method is a request; the batched call on the device runs when the LAST local node has made the
// request for the round, so each `for alpha` loop costs one C-ABI call and the objects stay in lock
// step exactly like the reference's.  A node that runs ahead of the others by more than one round
// gets -1 (the reference would compute on stale neighbour copies there).
// results() is valid once every local node has completed the phase, which is when the reference reads it.
// ---------------------------------------------------------------------------------------------

/** What dist_pgo reads from DPGOHash::results() (DPGO_types.h:204-322): the node's own poses in the
 * reference's per-node layout [t (n0 rows); R blocks (d n0 rows)] x d, and the scalars. */
struct NodeResult {
  Matrix Xk;
  DPGOResult scalars;
};

class NodeBatch {
 public:
  enum Phase { INITIALIZE = 0, UPDATE = 1, ITERATE = 2, COMMUNICATE = 3, NUM_PHASES = 4 };

  NodeBatch(std::shared_ptr<DPGODriver> drv, int num_nodes, int node_begin, int node_end)
      : drv_(std::move(drv)), A_(num_nodes), nb_(node_begin), ne_(node_end < 0 ? num_nodes : node_end),
        X_((int64_t)(drv_->d() + 1) * drv_->num_poses(), drv_->d()) {
    for (auto &c : calls_) c.assign((size_t)(ne_ - nb_), 0);
  }
  /** recv() of every node (DPGO_utils.cpp:426-435): for node alpha the poses of every other node beta it shares a
   * measurement with, ascending.  Needed by PerNode::receive only. */
  void set_measurements(const measurements_t &meas) {
    recv_.assign((size_t)A_, {});
    for (const auto &m : meas) {
      const int a = node_of((int64_t)m.i), b = node_of((int64_t)m.j);
      if (a == b) continue;
      int64_t fa, na, fb, nbp;
      range(a, fa, na); range(b, fb, nbp);
      recv_[(size_t)a][b].insert((int64_t)m.j - fb);
      recv_[(size_t)b][a].insert((int64_t)m.i - fa);
    }
  }
  int node_of(int64_t pose) const {
    const int64_t N = drv_->num_poses(), q = N / A_, r = N % A_;
    return pose < r * (q + 1) ? (int)(pose / (q + 1)) : (int)(r + (pose - r * (q + 1)) / std::max<int64_t>(q, 1));
  }
  const std::map<int, std::set<int64_t>> &recv(int node) const { return recv_[(size_t)node]; }
  /** first global pose and number of poses of a node: DPGO_utils.cpp:147-158 (the first N % A nodes hold one more) */
  void range(int node, int64_t &first, int64_t &n0) const {
    const int64_t N = drv_->num_poses(), q = N / A_, r = N % A_;
    first = (int64_t)node * q + std::min<int64_t>(node, r);
    n0 = q + (node < r ? 1 : 0);
  }
  /** rows of node `node` of a per-node iterate -> the staged global iterate */
  int stage(int node, const Matrix &Xk) {
    int64_t first, n0;
    range(node, first, n0);
    const int d = drv_->d();
    if (Xk.cols() != d || Xk.rows() < (d + 1) * n0) return -1;     // neighbour rows below the own ones are ignored
    const int64_t N = drv_->num_poses();
    for (int c = 0; c < d; ++c) {
      for (int64_t i = 0; i < n0; ++i) X_(first + i, c) = Xk(i, c);
      for (int64_t i = 0; i < d * n0; ++i) X_(N + d * first + i, c) = Xk(n0 + i, c);
    }
    return 0;
  }
  int request(int node, Phase ph) {
    if (node < nb_ || node >= ne_) return -1;
    std::vector<int64_t> &c = calls_[ph];
    if (c[(size_t)(node - nb_)] > done_[ph]) return -1;            // a second request before the round ran
    ++c[(size_t)(node - nb_)];
    for (int64_t v : c) if (v <= done_[ph]) return 0;              // somebody has not asked yet: deferred
    ++done_[ph];
    fresh_ = false;
    switch (ph) {
      case INITIALIZE: return drv_->initialize(X_);
      case UPDATE: return drv_->update();
      case ITERATE: return drv_->iterate();
      default: return drv_->communicate();
    }
  }
  /** own rows + scalars of a node from the device iterate (one mmpgo_get_poses per round, shared by all nodes) */
  int results(int node, NodeResult &out) {
    if (node < nb_ || node >= ne_) return -1;
    if (!fresh_) {
      const int rc = drv_->X(X_);
      if (rc) return rc;
      fresh_ = true;
    }
    int64_t first, n0;
    range(node, first, n0);
    const int d = drv_->d();
    const int64_t N = drv_->num_poses();
    if (out.Xk.rows() != (d + 1) * n0 || out.Xk.cols() != d) out.Xk = Matrix((d + 1) * n0, d);
    for (int c = 0; c < d; ++c) {
      for (int64_t i = 0; i < n0; ++i) out.Xk(i, c) = X_(first + i, c);
      for (int64_t i = 0; i < d * n0; ++i) out.Xk(n0 + i, c) = X_(N + d * first + i, c);
    }
    return drv_->results(node, out.scalars);
  }
  const std::shared_ptr<DPGODriver> &driver() const { return drv_; }
  int node_begin() const { return nb_; }
  int node_end() const { return ne_; }

 private:
  std::shared_ptr<DPGODriver> drv_;
  int A_, nb_, ne_;
  Matrix X_;                       // staged global iterate (initialize) / last device iterate (results)
  bool fresh_ = false;
  std::vector<std::map<int, std::set<int64_t>>> recv_;
  std::array<std::vector<int64_t>, NUM_PHASES> calls_;
  std::array<int64_t, NUM_PHASES> done_{{0, 0, 0, 0}};
};

/** One object per node with the reference's per-node methods (DPGOHash.h:18-90); Driver = DPGOHash or DPGOStar. */
template <class Driver>
class PerNode {
 public:
  PerNode(int node, std::shared_ptr<NodeBatch> batch) : node_(node), batch_(std::move(batch)) {}
  /** x: the node's iterate, own poses first ([t; R] of the n0 own poses, DPGOHash.cpp:20-43).  The neighbour
   * copies that follow in the reference's Xk are the owners' rows after DPGO::communicate (dist_pgo.cpp:446)
   * and are taken from the owners here. */
  int initialize(const Matrix &x) const {
    const int rc = batch_->stage(node_, x);
    return rc ? rc : batch_->request(node_, NodeBatch::INITIALIZE);
  }
  int update() const { return batch_->request(node_, NodeBatch::UPDATE); }
  int iterate() const { return batch_->request(node_, NodeBatch::ITERATE); }
  /** DPGOHash::communicate(pgos) (DPGOHash.h:28-86): the boundary poses move on the device */
  template <typename PGO>
  int communicate(const std::vector<std::shared_ptr<PGO>> &pgos) const {
    if ((int)pgos.size() < batch_->node_end() - batch_->node_begin()) return -1;   // "No information for node"
    return batch_->request(node_, NodeBatch::COMMUNICATE);
  }
  /** DPGOHash::receive (DPGOHash.cpp:45-82): one message per neighbour node beta, [t; R blocks] of the poses in
   * recv()[beta], ascending.  The sizes are checked like the reference's asserts (-1 on a node that is not a neighbour
   * or on a wrong row count).  The VALUES are not taken from the message: the neighbour copies of a node are the
   * owners' rows on the device (the same numbers when the message is what the reference sends, results().Xk of
   * beta), refreshed by the batched communicate() this call requests -- between GPUs by the library's exchange. */
  int receive(const std::map<int, Matrix> &msg) const {
    const int d = batch_->driver()->d();
    const auto &recv = batch_->recv(node_);
    for (const auto &beta : msg) {
      const auto it = recv.find(beta.first);
      if (it == recv.end()) return -1;                                  // "Can not find information for node"
      if (beta.second.rows() != (int64_t)(d + 1) * (int64_t)it->second.size() || beta.second.cols() != d) return -1;
    }
    return batch_->request(node_, NodeBatch::COMMUNICATE);
  }
  /** the poses of other nodes this node reads (DPGOProblem::recv(), DPGO_utils.cpp:426-435): node -> local pose ids */
  const std::map<int, std::set<int64_t>> &recv() const { return batch_->recv(node_); }
  const NodeResult &results() const {
    if (batch_->results(node_, res_)) throw std::runtime_error(mmpgo_last_error());
    return res_;
  }
  int node() const { return node_; }
  const std::shared_ptr<NodeBatch> &batch() const { return batch_; }

 private:
  int node_;
  std::shared_ptr<NodeBatch> batch_;
  mutable NodeResult res_;
};

/** dpgo_hash[alpha] for alpha in [node_begin, node_end): the replacement of dist_pgo.cpp:129-135. */
template <class Driver>
std::vector<std::shared_ptr<PerNode<Driver>>> make_per_node(int num_nodes, int d, int64_t num_poses,
                                                            const measurements_t &meas, const Options &opt,
                                                            int node_begin = 0, int node_end = -1) {
  auto drv = std::make_shared<Driver>(num_nodes, d, num_poses, meas, opt, node_begin, node_end);
  if (node_end < 0) node_end = num_nodes;
  auto batch = std::make_shared<NodeBatch>(drv, num_nodes, node_begin, node_end);
  batch->set_measurements(meas);
  std::vector<std::shared_ptr<PerNode<Driver>>> out;
  for (int a = node_begin; a < node_end; ++a) out.push_back(std::make_shared<PerNode<Driver>>(a, batch));
  return out;
}

}  // namespace DPGO
