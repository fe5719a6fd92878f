// dist_pgo: the reference's experiment driver (C++/examples/dist_pgo.cpp) on top of libmmpgo.
//
// Same flags (--dataset --num_nodes --iters --dist_init --loss --accelerated --save, :23-47), same
// hard-coded solver options (:103-120, in mmpgo_default_options), same loop order
// (iterate -> communicate -> update -> log 2F and 2|grad F|, :492-531) and the same output files
// (results_chordal_<N>_<amm|mm>.txt, estimates_<loss>.txt, :538-568).  Added, not replaced:
// `--loss gm`, `--algorithm hash|star`, `--device`.  The initialisation stays host code: a
// centralised chordal relaxation (the reference's `--dist_init false` path, :416-444); the
// distributed chordal initialisation (C++/DChordal) is out of scope and falls back to it.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>

#include "mmpgo_host/DPGO.h"

using DPGO::Matrix;
using DPGO::measurements_t;

static bool parse_bool(const std::string &v) { return v == "1" || v == "true" || v == "True" || v == "on" || v == "yes"; }

// ---- centralised chordal initialisation (host): rotations from the linear relaxation with
// pose 0 fixed to the identity, projected onto SO(d) by the device kernel; then least-squares
// translations with t_0 = 0.  Both systems are SPD and solved by Jacobi-preconditioned CG.
namespace {
struct CG {
  template <class Op>
  static int solve(const Op &A, const std::vector<double> &diag, int cols, const std::vector<double> &b,
                   std::vector<double> &x, double tol, int max_it) {
    const size_t n = b.size();
    std::vector<double> r(b), z(n), p(n), Ap(n);
    x.assign(n, 0.0);
    auto dot = [&](const std::vector<double> &u, const std::vector<double> &v) {
      double s = 0;
#pragma omp parallel for reduction(+ : s)
      for (long i = 0; i < (long)n; ++i) s += u[i] * v[i];
      return s;
    };
    const double bb = dot(b, b);
    if (bb == 0) return 0;
    for (size_t i = 0; i < n; ++i) z[i] = r[i] / diag[i / cols];
    p = z;
    double rz = dot(r, z);
    for (int it = 0; it < max_it; ++it) {
      A(p, Ap);
      const double alpha = rz / dot(p, Ap);
#pragma omp parallel for
      for (long i = 0; i < (long)n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
      if (dot(r, r) <= tol * tol * bb) return it + 1;
#pragma omp parallel for
      for (long i = 0; i < (long)n; ++i) z[i] = r[i] / diag[i / cols];
      const double rz2 = dot(r, z);
      const double beta = rz2 / rz;
      rz = rz2;
#pragma omp parallel for
      for (long i = 0; i < (long)n; ++i) p[i] = z[i] + beta * p[i];
    }
    return max_it;
  }
};
}  // namespace

static int chordal_initialization(int d, int64_t N, const measurements_t &meas, int device, Matrix &X) {
  const size_t dd = (size_t)d * d;
  // ---- rotations: minimise sum kappa |R_e^T Y_i - Y_j|^2, Y_0 = I   (Y_i = R_i^T, d x d row-major)
  auto applyR = [&](const std::vector<double> &Y, std::vector<double> &out) {   // full operator on N blocks
    std::fill(out.begin(), out.end(), 0.0);
    for (const auto &m : meas) {
      const double *Yi = &Y[(size_t)m.i * dd], *Yj = &Y[(size_t)m.j * dd];
      double *oi = &out[(size_t)m.i * dd], *oj = &out[(size_t)m.j * dd];
      for (int r = 0; r < d; ++r)
        for (int c = 0; c < d; ++c) {
          double RYj = 0, RtYi = 0;
          for (int k = 0; k < d; ++k) { RYj += m.R[r * d + k] * Yj[k * d + c]; RtYi += m.R[k * d + r] * Yi[k * d + c]; }
          oi[r * d + c] += m.kappa * (Yi[r * d + c] - RYj);
          oj[r * d + c] += m.kappa * (Yj[r * d + c] - RtYi);
        }
    }
  };
  std::vector<double> deg(N, 0.0), degt(N, 0.0);
  for (const auto &m : meas) { deg[m.i] += m.kappa; deg[m.j] += m.kappa; degt[m.i] += m.tau; degt[m.j] += m.tau; }
  for (int64_t i = 0; i < N; ++i) { if (deg[i] == 0) deg[i] = 1; if (degt[i] == 0) degt[i] = 1; }
  std::vector<double> Y0((size_t)N * dd, 0.0), rhs((size_t)N * dd), tmp((size_t)N * dd), Y;
  for (int k = 0; k < d; ++k) Y0[k * d + k] = 1.0;
  applyR(Y0, rhs);
  for (auto &v : rhs) v = -v;
  for (size_t k = 0; k < dd; ++k) rhs[k] = 0.0;
  auto A11 = [&](const std::vector<double> &v, std::vector<double> &out) {       // pose 0 pinned
    std::vector<double> vv(v);
    for (size_t k = 0; k < dd; ++k) vv[k] = 0.0;
    applyR(vv, out);
    for (size_t k = 0; k < dd; ++k) out[k] = v[k] * deg[0];
  };
  const int itR = CG::solve(A11, deg, (int)dd, rhs, Y, 1e-10, 20000);
  for (int k = 0; k < d; ++k) for (int c = 0; c < d; ++c) Y[k * d + c] = k == c ? 1.0 : 0.0;
  std::vector<double> Yp((size_t)N * dd);
  if (mmpgo_project_to_sodn(d, N, Y.data(), Yp.data(), device)) {
    std::cerr << "projection failed: " << mmpgo_last_error() << std::endl;
    return -1;
  }
  // ---- translations: minimise sum tau |t_i - t_j + t_e^T Y_i|^2, t_0 = 0
  std::vector<double> b((size_t)N * d, 0.0), t;
  for (const auto &m : meas) {
    const double *Yi = &Yp[(size_t)m.i * dd];
    for (int c = 0; c < d; ++c) {
      double ce = 0;
      for (int k = 0; k < d; ++k) ce += m.t[k] * Yi[k * d + c];
      b[(size_t)m.i * d + c] -= m.tau * ce;
      b[(size_t)m.j * d + c] += m.tau * ce;
    }
  }
  for (int c = 0; c < d; ++c) b[c] = 0.0;
  auto L11 = [&](const std::vector<double> &v, std::vector<double> &out) {
    std::fill(out.begin(), out.end(), 0.0);
    for (const auto &m : meas)
      for (int c = 0; c < d; ++c) {
        const double vi = m.i == 0 ? 0.0 : v[(size_t)m.i * d + c], vj = m.j == 0 ? 0.0 : v[(size_t)m.j * d + c];
        out[(size_t)m.i * d + c] += m.tau * (vi - vj);
        out[(size_t)m.j * d + c] += m.tau * (vj - vi);
      }
    for (int c = 0; c < d; ++c) out[c] = v[c] * degt[0];
  };
  const int itT = CG::solve(L11, degt, d, b, t, 1e-10, 50000);
  for (int c = 0; c < d; ++c) t[c] = 0.0;
  std::cout << "chordal initialization: " << itR << " / " << itT << " CG iterations (rotations / translations)" << std::endl;
  X = Matrix((d + 1) * N, d);
  for (int64_t i = 0; i < N; ++i) {
    for (int c = 0; c < d; ++c) X(i, c) = t[(size_t)i * d + c];
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c) X(N + d * i + r, c) = Yp[(size_t)i * dd + r * d + c];
  }
  return 0;
}

// dist_pgo.cpp:446-531 with the per-node objects: same loops, same reads of results().Xk, the objective and the
// gradient norm evaluated on the gathered X like the reference's dpgo_star.evaluate_f / evaluate_grad.
template <class PGO>
static int per_node_loop(const std::vector<std::shared_ptr<PGO>> &dpgo, const Matrix &X0, int d, int64_t num_poses,
                         int num_nodes, int num_iters, bool by_messages) {
  auto &batch = *dpgo[0]->batch();
  std::vector<Matrix> Xk(num_nodes);
  for (int alpha = 0; alpha < num_nodes; alpha++) {
    int64_t index, n;
    batch.range(alpha, index, n);
    Xk[alpha] = Matrix((d + 1) * n, d);
    for (int c = 0; c < d; ++c) {
      for (int64_t i = 0; i < n; ++i) Xk[alpha](i, c) = X0(index + i, c);
      for (int64_t i = 0; i < d * n; ++i) Xk[alpha](n + i, c) = X0(num_poses + d * index + i, c);
    }
  }
  for (int alpha = 0; alpha < num_nodes; alpha++) {
    if (dpgo[alpha]->initialize(Xk[alpha]) || dpgo[alpha]->update()) { std::cerr << mmpgo_last_error() << std::endl; return -1; }
  }
  Matrix X((d + 1) * num_poses, d), gradF;
  auto gather = [&]() {
    for (int alpha = 0; alpha < num_nodes; alpha++) {
      int64_t i, n;
      batch.range(alpha, i, n);
      const Matrix &R = dpgo[alpha]->results().Xk;
      for (int c = 0; c < d; ++c) {
        for (int64_t k = 0; k < n; ++k) X(i + k, c) = R(k, c);
        for (int64_t k = 0; k < d * n; ++k) X(num_poses + d * i + k, c) = R(n + k, c);
      }
    }
  };
  auto norm = [](const Matrix &G) {
    double s = 0;
    for (int64_t i = 0; i < G.rows(); ++i) for (int64_t c = 0; c < G.cols(); ++c) s += G(i, c) * G(i, c);
    return std::sqrt(s);
  };
  double fobj = 0, grad = 0;
  gather();
  if (batch.driver()->evaluate_f(X, fobj) || batch.driver()->evaluate_grad(X, gradF)) return -1;
  fobj *= 2; grad = 2 * norm(gradF);
  std::cout << "===============================================" << std::endl;
  std::cout << "Distributed PGO" << std::endl;
  std::cout << "-----------------------------------------------" << std::endl;
  for (int iter = 0; iter < num_iters; iter++) {
    std::cout << iter << ": " << std::setprecision(20) << fobj << " " << grad << std::endl;
    for (int alpha = 0; alpha < num_nodes; alpha++) if (dpgo[alpha]->iterate()) { std::cerr << mmpgo_last_error() << std::endl; return -1; }
    gather();
    if (!by_messages) {
      for (int alpha = 0; alpha < num_nodes; alpha++) if (dpgo[alpha]->communicate(dpgo)) { std::cerr << mmpgo_last_error() << std::endl; return -1; }
    } else {
      // the message-passing variant of the reference (DPGOHash::receive, DPGOHash.cpp:45-82; the buffers dist_pgo
      // sizes at :456-458): node alpha gets from every neighbour beta the poses listed in recv()[beta]
      for (int alpha = 0; alpha < num_nodes; alpha++) {
        std::map<int, Matrix> msgs;
        for (const auto &info : dpgo[alpha]->recv()) {
          const Matrix &Xb = dpgo[info.first]->results().Xk;
          const int64_t nb = Xb.rows() / (d + 1), np = (int64_t)info.second.size();
          Matrix msg((d + 1) * np, d);
          int64_t t = 0;
          for (int64_t j : info.second) {
            for (int c = 0; c < d; ++c) {
              msg(t, c) = Xb(j, c);
              for (int r = 0; r < d; ++r) msg(np + t * d + r, c) = Xb(nb + j * d + r, c);
            }
            ++t;
          }
          msgs[info.first] = msg;
        }
        if (dpgo[alpha]->receive(msgs)) { std::cerr << "receive: inconsistent message" << std::endl; return -1; }
      }
    }
    for (int alpha = 0; alpha < num_nodes; alpha++) if (dpgo[alpha]->update()) { std::cerr << mmpgo_last_error() << std::endl; return -1; }
    if (batch.driver()->evaluate_f(X, fobj) || batch.driver()->evaluate_grad(X, gradF)) return -1;
    fobj *= 2; grad = 2 * norm(gradF);
  }
  std::cout << "---------------------------------------" << std::endl;
  std::cout << "final objective: " << fobj << std::endl;
  std::cout << "final gradient: " << grad << std::endl;
  return 0;
}

int main(int argc, char *argv[]) {
  if (argc < 2) {
    std::cout << "Usage: " << argv[0] << " [input .g2o file]" << std::endl;
    return 1;
  }
  std::map<std::string, std::string> opt = {{"iters", "1000"}, {"dist_init", "true"}, {"loss", "trivial"},
                                             {"accelerated", "true"}, {"save", "true"}, {"algorithm", "hash"},
                                             {"device", "0"}, {"dist_init_fallback", "false"}, {"per_node", "false"},
                                             {"preconditioner", "block_jacobi"}};
  for (int a = 1; a < argc; ++a) {
    std::string s = argv[a];
    if (s == "--help") {
      std::cout << "Program options:\n"
                   "  --help                     produce help message\n"
                   "  --dataset arg              path to pose graph dataset\n"
                   "  --num_nodes arg            number of nodes\n"
                   "  --iters arg (=1000)        number of iterations\n"
                   "  --dist_init arg (=true)    distributed (\"true\") or centralized (\"false\") initialization\n"
                   "  --loss arg (=trivial)      loss type (\"trivial\", \"huber\", \"welsch\" or \"gm\")\n"
                   "  --accelerated arg (=1)     whether accelerated or not\n"
                   "  --save arg (=1)            whether to save the optimization results or not\n"
                   "  --algorithm arg (=hash)    \"hash\" (AMM-PGO# / MM-PGO) or \"star\" (AMM-PGO*)\n"
                   "  --device arg (=0)          CUDA device ordinal\n"
                   "  --init arg                 text file with the initial iterate ((d+1)N rows of d numbers)\n"
                   "  --dist_init_fallback arg (=false)  with --dist_init true: use the centralised chordal initialisation\n"
                   "  --per_node arg (=false)    run the reference's loop with one driver object per node (DPGO::PerNode);\n"
                   "                             \"receive\": exchange through DPGOHash::receive messages instead of communicate()\n"
                   "  --preconditioner arg (=block_jacobi)  tCG preconditioner: \"block_jacobi\", \"regularized_cholesky\" (the\n"
                   "                             reference's default, DPGO_types.h:155), \"jacobi\" or \"none\"\n"
                   "  --parse_only arg           only read the dataset and print its checksums\n";
      return 0;
    }
    if (s.rfind("--", 0) != 0) { std::cerr << "unrecognised argument " << s << std::endl; return -1; }
    s = s.substr(2);
    const size_t eq = s.find('=');
    if (eq != std::string::npos) opt[s.substr(0, eq)] = s.substr(eq + 1);
    else if (a + 1 < argc) opt[s] = argv[++a];
    else { std::cerr << "missing value for --" << s << std::endl; return -1; }
  }
  if (!opt.count("dataset")) { std::cerr << "No dataset has been specfied." << std::endl; return -1; }
  if (!opt.count("num_nodes")) { std::cerr << "No number of nodes has been specfied." << std::endl; return -1; }
  const std::string filename = opt["dataset"], loss_type = opt["loss"];
  const int num_nodes = std::atoi(opt["num_nodes"].c_str()), num_iters = std::atoi(opt["iters"].c_str());
  const bool accelerated = parse_bool(opt["accelerated"]), dist_chordal = parse_bool(opt["dist_init"]);
  const bool save = parse_bool(opt["save"]), star = opt["algorithm"] == "star";

  DPGO::Options options;
  options.device = std::atoi(opt["device"].c_str());
  if (loss_type == "trivial") options.loss = DPGO::Loss::None;
  else if (loss_type == "huber") options.loss = DPGO::Loss::Huber;
  else if (loss_type == "welsch") options.loss = DPGO::Loss::Welsch;
  else if (loss_type == "gm" || loss_type == "geman-mcclure") options.loss = DPGO::Loss::GemanMcClure;
  else {
    std::cerr << " The loss type can only be \"trivial\", \"huber\", \"welsch\" or \"gm\"." << std::endl;
    return -1;
  }
  options.scheme = accelerated ? DPGO::Scheme::AMM : DPGO::Scheme::MM;
  {
    const std::string pre = opt["preconditioner"];
    if (pre == "block_jacobi") options.preconditioner = DPGO::Preconditioner::BlockJacobi;
    else if (pre == "regularized_cholesky") options.preconditioner = DPGO::Preconditioner::RegularizedCholesky;
    else if (pre == "jacobi") options.preconditioner = DPGO::Preconditioner::Jacobi;
    else if (pre == "none") options.preconditioner = DPGO::Preconditioner::None;
    else {
      std::cerr << " The preconditioner can only be \"block_jacobi\", \"regularized_cholesky\", \"jacobi\" or \"none\"." << std::endl;
      return -1;
    }
  }

  try {
    int64_t num_poses = 0;
    measurements_t measurements;
    const int d = DPGO::read_g2o_file(filename, num_poses, measurements);
    std::cout << "read " << measurements.size() << " measurements between " << num_poses << " poses (SE(" << d
              << ")) from " << filename << std::endl;
    if (opt.count("parse_only")) {
      // reader check (no device needed): sizes and checksums of what read_g2o_file produced
      double st = 0, sk = 0, sR = 0, stt = 0;
      long long si = 0, sj = 0;
      for (const auto &m : measurements) {
        st += m.tau; sk += m.kappa; si += m.i; sj += m.j;
        for (int k = 0; k < d * d; ++k) sR += m.R[k] * (k + 1);
        for (int k = 0; k < d; ++k) stt += m.t[k] * (k + 1);
      }
      std::cout << std::setprecision(17) << "parse_only " << d << " " << num_poses << " " << measurements.size() << " " << si
                << " " << sj << " " << st << " " << sk << " " << sR << " " << stt << std::endl;
      return 0;
    }
    // --dist_init true (the reference's default, dist_pgo.cpp:144-415) asks for the distributed chordal
    // initialisation (C++/DChordal), host code outside this path: refused unless the caller opts into the
    // centralised chordal initialisation (--dist_init false, or --dist_init_fallback true) or passes --init
    if (dist_chordal && !opt.count("init") && !parse_bool(opt["dist_init_fallback"])) {
      std::cerr << "--dist_init true: the distributed chordal initialisation (DChordal) is not part of this path "
                   "(MMPGO_ERR_UNSUPPORTED).  Pass --dist_init false (centralised chordal initialisation, "
                   "dist_pgo.cpp:416-444), --dist_init_fallback true, or --init <file>." << std::endl;
      return MMPGO_ERR_UNSUPPORTED;
    }
    if (dist_chordal && !opt.count("init"))
      std::cout << "note: --dist_init_fallback: using the centralised chordal initialisation" << std::endl;
    Matrix X;
    if (opt.count("init")) {
      // an initial iterate from a text file: (d+1)N rows of d numbers, the reference's layout
      X = Matrix((d + 1) * num_poses, d);
      std::ifstream in(opt["init"]);
      if (!in.is_open()) throw std::runtime_error("cannot open " + opt["init"]);
      for (int64_t i = 0; i < X.rows(); ++i)
        for (int c = 0; c < d; ++c)
          if (!(in >> X(i, c))) throw std::runtime_error("initial iterate file too short");
    } else if (chordal_initialization(d, num_poses, measurements, options.device, X)) return -1;

    if (parse_bool(opt["per_node"]) || opt["per_node"] == "receive") {
      const bool by_messages = opt["per_node"] == "receive";
      // The reference's main loop as it stands (dist_pgo.cpp:446-531), one object per node: the objects
      // forward to one batched driver (DPGO::PerNode, mmpgo_host/DPGO.h).
      if (star)
        return per_node_loop(DPGO::make_per_node<DPGO::DPGOStar>(num_nodes, d, num_poses, measurements, options), X, d,
                             num_poses, num_nodes, num_iters, by_messages);
      return per_node_loop(DPGO::make_per_node<DPGO::DPGOHash>(num_nodes, d, num_poses, measurements, options), X, d,
                           num_poses, num_nodes, num_iters, by_messages);
    }
    std::unique_ptr<DPGO::DPGODriver> dpgo;
    if (star) dpgo.reset(new DPGO::DPGOStar(num_nodes, d, num_poses, measurements, options));
    else dpgo.reset(new DPGO::DPGOHash(num_nodes, d, num_poses, measurements, options));
    if (dpgo->initialize(X) || dpgo->update()) { std::cerr << mmpgo_last_error() << std::endl; return -1; }

    double F = 0, g2 = 0;
    dpgo->objective(F, g2);
    double fobj = 2 * F, grad = 2 * std::sqrt(g2);
    std::vector<std::array<double, 4>> results;
    results.push_back({0, 0, fobj, grad});
    double time = 0;
    std::cout << "===============================================" << std::endl;
    std::cout << "Distributed PGO" << std::endl;
    std::cout << "-----------------------------------------------" << std::endl;
    for (int iter = 0; iter < num_iters; iter++) {
      std::cout << iter << ": " << std::setprecision(20) << fobj << " " << grad << std::endl;
      const auto t0 = std::chrono::steady_clock::now();
      if (dpgo->iterate() || dpgo->communicate() || dpgo->update()) { std::cerr << mmpgo_last_error() << std::endl; return -1; }
      mmpgo_synchronize(dpgo->handle());
      time += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      dpgo->objective(F, g2);
      fobj = 2 * F; grad = 2 * std::sqrt(g2);
      results.push_back({double(iter) + 1, time, fobj, grad});
    }
    std::cout << "---------------------------------------" << std::endl;
    std::cout << "final objective: " << fobj << std::endl;
    std::cout << "final gradient: " << grad << std::endl;
    std::cout << "time: " << time / num_nodes << " s/node." << std::endl;

    if (save) {
      const std::string resfile = "results_chordal_" + std::to_string(num_nodes) + "_" + (accelerated ? "amm" : "mm") + ".txt";
      std::ofstream output(resfile);
      if (!output.is_open()) return -1;
      for (const auto &res : results)
        output << int(res[0]) << " " << std::setprecision(16) << res[1] << " " << std::setprecision(16) << res[2] << " "
               << std::setprecision(16) << res[3] << std::endl;
      output.close();
      if (dpgo->X(X)) { std::cerr << mmpgo_last_error() << std::endl; return -1; }
      // express the estimate in the frame of pose 0 (dist_pgo.cpp:553-557)
      std::vector<double> t0v(d), R0(d * d);
      for (int c = 0; c < d; ++c) t0v[c] = X(0, c);
      for (int64_t i = 0; i < num_poses; ++i) for (int c = 0; c < d; ++c) X(i, c) -= t0v[c];
      for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) R0[r * d + c] = X(num_poses + c, r);   // R = block^T
      Matrix Xo(X.rows(), d);
      for (int64_t i = 0; i < X.rows(); ++i)
        for (int c = 0; c < d; ++c) {
          double s = 0;
          for (int k = 0; k < d; ++k) s += X(i, k) * R0[k * d + c];
          Xo(i, c) = s;
        }
      output.open("./estimates_" + loss_type + ".txt");
      if (!output.is_open()) return -1;
      for (int64_t i = 0; i < Xo.rows(); ++i) {
        for (int c = 0; c < d; ++c) output << (c ? " " : "") << Xo(i, c);
        output << "\n";
      }
      output.close();
    }
  } catch (const std::exception &e) {
    std::cerr << "dist_pgo: " << e.what() << std::endl;
    return -1;
  }
  return 0;
}
