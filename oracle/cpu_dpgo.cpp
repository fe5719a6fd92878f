// oracle/cpu_dpgo.cpp -- TEST INFRASTRUCTURE / CPU BASELINE ONLY (never linked into libmmpgo).
//
// The restated reference algorithm in the reference's own language, dependency-free C++17 +
// OpenMP: what BASELINE.md section 3 / SURVEY.md section 8(d) call the "restated CPU baseline".
// It executes the per-iteration path of
//   DPGOProblem   C++/DPGO/src/DPGOProblem.cpp (evaluate_E :634-681, evaluate_g_and_f0 :222-267,
//                 evaluate_g_and_f :360-424, evaluate_none_g_and_f(0) :269-287/:516-542, evaluate_G
//                 :180-203, proximal :600-632, retract :127-143, Hessian-vector :552-577, precondition
//                 :579-598), include/DPGO/DPGOProblem.h:275-294 (recover_translations)
//   DPGOHash      C++/DPGO/src/DPGOHash.cpp:84-628        DPGOStar   C++/DPGO/src/DPGOStar.cpp:126-711
//   TNT / STPCG   C++/Optimization/include/Optimization/Riemannian/TNT.h:242-693,
//                 LinearAlgebra/IterativeSolvers.h:166-426
// on the SAME scalar row-major sparse matrices the reference builds (G, S, P, P0, Q, D, B1, U, N, V,
// G01, G10, G11; DPGO_utils.cpp:1398-2967) -- they are assembled by oracle/data_matrix.py and handed
// over as CSR arrays -- with a sparse Cholesky factor of G00 in place of CHOLMOD (up-looking, the
// fill-reducing ordering is passed in) and the reference's own AVX2 SO(d) projections, compiled from
// /root/reference and linked in (oracle/Makefile).  Threading mirrors the reference: OpenMP
// `parallel for` inside the sparse products and per-pose loops (DPGO_types.h:20-23), nodes looped
// serially (C++/examples/dist_pgo.cpp:497-520); mode 1 instead runs the nodes in parallel.
//
// It is checked against the numpy oracle (tests/test_oracle_cpu_baseline.py) and timed by
// bench.py --impl reference.  Only iterates k and k-1 are kept (the reference keeps all of them).
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <vector>

extern "C" {
int ref_project_to_SO3n(const double *A, double *U, long n);
int ref_project_to_SO2n(const double *A, double *U, long n);
}

namespace {

bool g_par_ops = true;          // OpenMP inside the operators (reference style)

struct Mat {                    // column-major dense matrix (Eigen::MatrixXd)
  int r = 0, c = 0;
  std::vector<double> v;
  Mat() {}
  Mat(int rows, int cols) : r(rows), c(cols), v((size_t)rows * cols, 0.0) {}
  double *col(int j) { return v.data() + (size_t)j * r; }
  const double *col(int j) const { return v.data() + (size_t)j * r; }
  double &operator()(int i, int j) { return v[(size_t)j * r + i]; }
  double operator()(int i, int j) const { return v[(size_t)j * r + i]; }
};

struct Csr {                    // Eigen::SparseMatrix<double, RowMajor>
  int rows = 0, cols = 0;
  std::vector<int> ptr, idx;
  std::vector<double> val;
  bool empty() const { return rows == 0; }
};

// Y(row0 + i, :) = (A X)(i, :) ; X rows [xoff, xoff + A.cols)
void spmm(const Csr &A, const Mat &X, int xoff, Mat &Y, int row0 = 0) {
  const int d = X.c;
#pragma omp parallel for schedule(static) if (g_par_ops && A.rows > 2048)
  for (int i = 0; i < A.rows; ++i) {
    double acc[3] = {0, 0, 0};
    for (int p = A.ptr[i]; p < A.ptr[i + 1]; ++p) {
      const double a = A.val[p];
      const int j = xoff + A.idx[p];
      for (int c = 0; c < d; ++c) acc[c] += a * X(j, c);
    }
    for (int c = 0; c < d; ++c) Y(row0 + i, c) = acc[c];
  }
}
Mat mul(const Csr &A, const Mat &X, int xoff = 0) {
  Mat Y(A.rows, X.c);
  spmm(A, X, xoff, Y);
  return Y;
}
// A^T X
Mat mulT(const Csr &A, const Mat &X) {
  Mat Y(A.cols, X.c);
  for (int i = 0; i < A.rows; ++i)
    for (int p = A.ptr[i]; p < A.ptr[i + 1]; ++p)
      for (int c = 0; c < X.c; ++c) Y(A.idx[p], c) += A.val[p] * X(i, c);
  return Y;
}
double tr(const Mat &A, const Mat &B) {
  double s = 0.0;
  const size_t n = A.v.size();
#pragma omp parallel for reduction(+ : s) schedule(static) if (g_par_ops && n > 65536)
  for (long i = 0; i < (long)n; ++i) s += A.v[i] * B.v[i];
  return s;
}
// rows [r0, r0 + n) of A
Mat rows(const Mat &A, int r0, int n) {
  Mat B(n, A.c);
  for (int c = 0; c < A.c; ++c) std::memcpy(B.col(c), A.col(c) + r0, sizeof(double) * n);
  return B;
}
void set_rows(Mat &A, int r0, const Mat &B) {
  for (int c = 0; c < A.c; ++c) std::memcpy(A.col(c) + r0, B.col(c), sizeof(double) * B.r);
}
void axpy(Mat &Y, double a, const Mat &X) {
  const size_t n = Y.v.size();
  for (size_t i = 0; i < n; ++i) Y.v[i] += a * X.v[i];
}

// ---- sparse Cholesky (up-looking), ordering supplied ---------------------------------------
struct SparseChol {
  int n = 0;
  std::vector<int> perm, cp, ri;       // perm[new] = old; L in CSC
  std::vector<double> lx;
  bool factor(const Csr &A, const int *perm_in) {
    n = A.rows;
    perm.assign(perm_in, perm_in + n);
    std::vector<int> inv(n);
    for (int k = 0; k < n; ++k) inv[perm[k]] = k;
    // C = P A P^T, upper triangle by columns (column k holds rows i <= k)
    std::vector<int> ccp(n + 1, 0);
    for (int i = 0; i < n; ++i)
      for (int p = A.ptr[i]; p < A.ptr[i + 1]; ++p) {
        const int a = inv[i], b = inv[A.idx[p]];
        if (a <= b) ccp[b + 1]++;
      }
    for (int k = 0; k < n; ++k) ccp[k + 1] += ccp[k];
    std::vector<int> cri(ccp[n]), fill(ccp.begin(), ccp.end() - 1);
    std::vector<double> cx(ccp[n]);
    for (int i = 0; i < n; ++i)
      for (int p = A.ptr[i]; p < A.ptr[i + 1]; ++p) {
        const int a = inv[i], b = inv[A.idx[p]];
        if (a <= b) { cri[fill[b]] = a; cx[fill[b]++] = A.val[p]; }
      }
    // elimination tree
    std::vector<int> parent(n, -1), anc(n, -1);
    for (int k = 0; k < n; ++k)
      for (int p = ccp[k]; p < ccp[k + 1]; ++p)
        for (int i = cri[p]; i != -1 && i < k;) {
          const int nx = anc[i];
          anc[i] = k;
          if (nx == -1) parent[i] = k;
          i = nx;
        }
    // row patterns (ereach) -> column counts, then the numeric pass
    std::vector<int> mark(n, -1), stack(n), cnt(n, 1);
    auto ereach = [&](int k, int &top) {
      top = n;
      mark[k] = k;
      for (int p = ccp[k]; p < ccp[k + 1]; ++p) {
        int i = cri[p], len = 0;
        if (i > k) continue;
        for (; mark[i] != k; i = parent[i]) { stack[len++] = i; mark[i] = k; }
        while (len > 0) stack[--top] = stack[--len];
      }
    };
    for (int k = 0; k < n; ++k) {
      int top;
      ereach(k, top);
      for (int t = top; t < n; ++t) cnt[stack[t]]++;
    }
    cp.assign(n + 1, 0);
    for (int k = 0; k < n; ++k) cp[k + 1] = cp[k] + cnt[k];
    ri.assign(cp[n], 0);
    lx.assign(cp[n], 0.0);
    std::vector<int> nz(cp.begin(), cp.end() - 1);
    std::vector<double> x(n, 0.0);
    std::fill(mark.begin(), mark.end(), -1);
    for (int k = 0; k < n; ++k) {
      int top;
      ereach(k, top);
      x[k] = 0.0;
      for (int p = ccp[k]; p < ccp[k + 1]; ++p) if (cri[p] <= k) x[cri[p]] += cx[p];
      double dk = x[k];
      x[k] = 0.0;
      for (; top < n; ++top) {
        const int i = stack[top];
        const double lki = x[i] / lx[cp[i]];
        x[i] = 0.0;
        for (int p = cp[i] + 1; p < nz[i]; ++p) x[ri[p]] -= lx[p] * lki;
        dk -= lki * lki;
        ri[nz[i]] = k;
        lx[nz[i]++] = lki;
      }
      if (!(dk > 0.0)) return false;
      ri[nz[k]] = k;
      lx[nz[k]++] = std::sqrt(dk);
    }
    return true;
  }
  // B <- A^{-1} B  (all columns)
  void solve(Mat &B) const {
    const int d = B.c;
#pragma omp parallel for schedule(static) if (g_par_ops && d > 1)
    for (int c = 0; c < d; ++c) {
      std::vector<double> x(n);
      double *b = B.col(c);
      for (int k = 0; k < n; ++k) x[k] = b[perm[k]];
      for (int j = 0; j < n; ++j) {
        x[j] /= lx[cp[j]];
        const double xj = x[j];
        for (int p = cp[j] + 1; p < cp[j + 1]; ++p) x[ri[p]] -= lx[p] * xj;
      }
      for (int j = n - 1; j >= 0; --j) {
        double s = x[j];
        for (int p = cp[j] + 1; p < cp[j + 1]; ++p) s -= lx[p] * x[ri[p]];
        x[j] = s / lx[cp[j]];
      }
      for (int k = 0; k < n; ++k) b[perm[k]] = x[k];
    }
  }
};

// ---- SO(d)^n geometry (SOdProduct.h:39-116) --------------------------------------------------
// R: (d n) x d, pose i = rows [d i, d i + d)
Mat so_project(const Mat &M, int d) {
  const int n = M.r / d;
  const int np = std::max(n, 4);
  std::vector<double> A((size_t)np * d * d, 0.0), U((size_t)np * d * d, 0.0);
  for (int i = n; i < np; ++i) for (int k = 0; k < d; ++k) A[(size_t)i * d * d + k * d + k] = 1.0;
#pragma omp parallel for schedule(static) if (g_par_ops && n > 4096)
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c) A[(size_t)i * d * d + r * d + c] = M(d * i + r, c);
  // the reference's AVX2 kernels, four matrices per call (DPGO_utils.h:515-565); OpenMP over chunks
  const int chunks = g_par_ops ? std::max(1, std::min(omp_get_max_threads(), np / 1024)) : 1;
#pragma omp parallel for schedule(static) if (chunks > 1)
  for (int ch = 0; ch < chunks; ++ch) {
    const long b = (long)np * ch / chunks / 4 * 4, e = ch + 1 == chunks ? np : (long)np * (ch + 1) / chunks / 4 * 4;
    if (e - b >= 4) {
      if (d == 3) ref_project_to_SO3n(&A[(size_t)b * 9], &U[(size_t)b * 9], e - b);
      else ref_project_to_SO2n(&A[(size_t)b * 4], &U[(size_t)b * 4], e - b);
    }
  }
  Mat R(M.r, d);
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c) R(d * i + r, c) = U[(size_t)i * d * d + r * d + c];
  return R;
}
// V - sym(V_i Y_i^T) Y_i   (SOdProduct::Proj, :96-103, with SymBlockDiagProduct :64-89)
Mat sym_block_diag_product(const Mat &A, const Mat &B, const Mat &C, int d) {
  // for every pose: sym(B_i^T C_i)-style product as in the reference: A_i * sym(B_i^T C_i)?  The
  // reference stores poses as row blocks Y_i (d x d) of a (d n) x d matrix and computes
  //   out_i = sym(C_i B_i^T) A_i ,  sym(M) = (M + M^T) / 2
  const int n = A.r / d;
  Mat out(A.r, d);
#pragma omp parallel for schedule(static) if (g_par_ops && n > 4096)
  for (int i = 0; i < n; ++i) {
    double S[9];
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c) {
        double u = 0.0, w = 0.0;
        for (int k = 0; k < d; ++k) { u += C(d * i + r, k) * B(d * i + c, k); w += B(d * i + r, k) * C(d * i + c, k); }
        S[r * d + c] = 0.5 * (u + w);
      }
    for (int r = 0; r < d; ++r)
      for (int k = 0; k < d; ++k) {
        double t = 0.0;
        for (int c = 0; c < d; ++c) t += S[r * d + c] * A(d * i + c, k);
        out(d * i + r, k) = t;
      }
  }
  return out;
}
Mat so_proj(const Mat &Y, const Mat &V, int d) {
  Mat P = V;
  const Mat S = sym_block_diag_product(Y, Y, V, d);
  axpy(P, -1.0, S);
  return P;
}

// ---- options -----------------------------------------------------------------------------
struct Options {
  int scheme = 1, loss = 0, precon = 2;
  double xi = 1e-11, loss_reg = 0.25, accepted_delta = 5e-4, eta[2] = {5e-4, 2.5e-2}, psi = 1e-10, phi = 1e-6;
  int max_soft_restart_hits[2] = {10, 25}, oscillation_cnt_period = 15, max_oscillations = 12;
  double grad_norm_tol = 1e-3, precon_grad_norm_tol = 1e-4, rel_func_decrease_tol = 1e-6, stepsize_tol = 1e-4;
  int max_iterations = 10, max_iterations_accepted = 1, max_tCG = 10000;
  double kappa = 0.05, theta = 0.9;
};

// ---- STPCG / TNT (IterativeSolvers.h:166-426, TNT.h:242-693) -------------------------------
typedef std::function<Mat(const Mat &)> LinOp;
struct TntOut { Mat x; double f = 0; int inner = 0, iters = 0; };

Mat stpcg(const Mat &g, const LinOp &H, double Delta, int max_it, double kappa_fgr, double theta, const LinOp *P,
          double &hMnorm, int &its) {
  const double eps = 1e-8;
  Mat s(g.r, g.c), r = g, v = P ? (*P)(r) : r, p = v;
  for (double &x : p.v) x = -x;
  double sk_M_pk = 0, sk_M_2 = 0, pk_M_2 = tr(r, v);
  const double D2 = Delta * Delta;
  const double r0 = std::sqrt(tr(r, v));
  const double target = r0 * std::min(kappa_fgr, std::pow(r0, theta));
  its = 0;
  while (its < max_it) {
    if (std::sqrt(tr(r, v)) <= target) break;
    const Mat Hp = H(p);
    const double kap = tr(p, Hp);
    if (std::sqrt(tr(Hp, Hp)) / std::sqrt(tr(p, p)) < eps) {
      if (tr(p, r) < 0) { for (double &x : p.v) x = -x; sk_M_pk = -sk_M_pk; }
      const double sigma = (-sk_M_pk + std::sqrt(sk_M_pk * sk_M_pk + pk_M_2 * (D2 - sk_M_2))) / pk_M_2;
      axpy(s, sigma, p);
      hMnorm = Delta;
      return s;
    }
    const double alpha = tr(r, v) / kap;
    const double skp1 = sk_M_2 + 2 * alpha * sk_M_pk + alpha * alpha * pk_M_2;
    if (kap <= 0 || skp1 > D2) {
      const double sigma = (-sk_M_pk + std::sqrt(sk_M_pk * sk_M_pk + pk_M_2 * (D2 - sk_M_2))) / pk_M_2;
      axpy(s, sigma, p);
      hMnorm = Delta;
      return s;
    }
    axpy(s, alpha, p);
    axpy(r, alpha, Hp);
    v = P ? (*P)(r) : r;
    const double rv = tr(r, v);
    const double beta = rv / (alpha * kap);
    sk_M_2 = skp1;
    sk_M_pk = beta * (sk_M_pk + alpha * pk_M_2);
    pk_M_2 = rv + beta * beta * pk_M_2;
    for (size_t i = 0; i < p.v.size(); ++i) p.v[i] = -v.v[i] + beta * p.v[i];
    ++its;
  }
  hMnorm = std::sqrt(sk_M_2);
  return s;
}

struct TntProblem {
  std::function<double(const Mat &)> f;
  std::function<Mat(const Mat &, Mat &nab)> grad;                  // returns Riemannian gradient, fills nab
  std::function<Mat(const Mat &, const Mat &nab, const Mat &)> hess;
  std::function<Mat(const Mat &, const Mat &)> retract, precon;
  bool has_precon = false;
};

TntOut tnt(const TntProblem &pb, const Mat &x0, const Options &o) {
  const double sqrt_eps = std::sqrt(std::numeric_limits<double>::epsilon());
  TntOut out;
  Mat x = x0, nab;
  double fx = pb.f(x);
  Mat grad = pb.grad(x, nab);
  double gnorm = std::sqrt(tr(grad, grad)), pgnorm = gnorm;
  if (pb.has_precon) { const Mat pg = pb.precon(x, grad); pgnorm = std::sqrt(tr(pg, pg)); }
  double Delta = 1.0;
  int it = 0, acc = 0;
  while (it < o.max_iterations && acc < o.max_iterations_accepted) {
    if (gnorm < o.grad_norm_tol || pgnorm < o.precon_grad_norm_tol) break;
    const LinOp H = [&](const Mat &v) { return pb.hess(x, nab, v); };
    const LinOp P = [&](const Mat &v) { return pb.precon(x, v); };
    double hM;
    int inner;
    const Mat h = stpcg(grad, H, Delta, o.max_tCG, o.kappa, o.theta, pb.has_precon ? &P : nullptr, hM, inner);
    out.inner += inner;
    const double h_norm = std::sqrt(tr(h, h));
    const Mat xp = pb.retract(x, h);
    const double fp = pb.f(xp);
    const double dm = -tr(grad, h) - 0.5 * tr(h, pb.hess(x, nab, h));      // TNT.h:514-515
    const double df = fx - fp;
    const double rel = df / (sqrt_eps + std::fabs(fx));
    const double rho = df / dm;
    const bool ok = !std::isnan(rho) && rho > 0.05;
    acc += ok ? 1 : 0;
    if (ok) {
      x = xp;
      fx = fp;
      if (rel < o.rel_func_decrease_tol || h_norm < o.stepsize_tol) break;
      grad = pb.grad(x, nab);
      gnorm = std::sqrt(tr(grad, grad));
      pgnorm = gnorm;
      if (pb.has_precon) { const Mat pg = pb.precon(x, grad); pgnorm = std::sqrt(tr(pg, pg)); }
    }
    if (!std::isnan(rho) && rho >= 0.9) Delta = std::max(2.5 * hM, Delta);
    else if (std::isnan(rho) || rho < 0.05) {
      Delta = 0.25 * hM;
      if (Delta < 1e-6) break;
    }
    ++it;
  }
  out.x = x;
  out.f = fx;
  out.iters = it;
  return out;
}

// ---- DPGOProblem -------------------------------------------------------------------------
struct Problem {
  int d = 3, n0 = 0, n1 = 0, m1 = 0;
  bool quadratic = true;
  Options o;
  Csr G, Grot, G01, G10, G11, S, P, P0, Q, D, B1, U, N, V;
  std::vector<double> T, pinv;      // T diagonal; block-Jacobi / Jacobi inverse of G11's diagonal
  SparseChol L;
  std::vector<double> w;            // last IRLS weights
  int size0() const { return (d + 1) * n0; }

  Mat recover_translations(const Mat &R, const Mat &g) const {          // DPGOProblem.h:275-294
    Mat t = mul(G01, R);
    for (int c = 0; c < d; ++c) for (int i = 0; i < n0; ++i) t(i, c) += g(i, c);
    L.solve(t);
    for (double &x : t.v) x = -x;
    return t;
  }
  double evaluate_G(const Mat &Y, const Mat &g, double f) const {       // DPGOProblem.cpp:180-203
    Mat temp = mul(G, Y);
    for (size_t i = 0; i < temp.v.size(); ++i) temp.v[i] = g.v[i] + 0.5 * temp.v[i];
    return tr(Y, temp) + f;
  }
  // DPGOProblem.cpp:634-681 -> DfobjE (full Z rows), fobjE
  void evaluate_E(const Mat &Z, Mat &DfobjE, double &fobjE) {
    Mat Err = mul(B1, Z);
    const double delta = o.loss_reg;
    w.assign(m1, 1.0);
    double fE = 0.0;
    for (int e = 0; e < m1; ++e) {
      double s = 0.0;
      for (int r = 0; r <= d; ++r) for (int c = 0; c < d; ++c) s += Err((d + 1) * e + r, c) * Err((d + 1) * e + r, c);
      double wt = 1.0;
      if (o.loss == 0) fE += 0.5 * s;
      else if (o.loss == 1) { const double resc = std::sqrt(std::max(s, delta)); wt = std::sqrt(delta) / resc; fE += 0.5 * std::min(2 * std::sqrt(delta) * resc - delta, s); }
      else if (o.loss == 2) { wt = delta * delta / ((s + delta) * (s + delta)); fE += 0.5 * delta * (s / (s + delta)); }
      else { wt = std::exp(-s / delta); fE += 0.5 * (delta - delta * wt); }
      w[e] = wt;
      for (int r = 0; r <= d; ++r) for (int c = 0; c < d; ++c) Err((d + 1) * e + r, c) *= wt;
    }
    DfobjE = mulT(B1, Err);
    fobjE = fE;
  }
  Mat evaluate_g(const Mat &Z) {                                        // DPGOProblem.cpp:683-725
    if (quadratic) return mul(S, Z);
    Mat DfE; double fE;
    evaluate_E(Z, DfE, fE);
    Mat g = rows(DfE, 0, size0());
    const Mat X = rows(Z, 0, size0());
    axpy(g, -1.0, mul(D, X));
    return g;
  }
  Mat proximal(const Mat &Z, const Mat &Df) const {                     // DPGOProblem.cpp:600-632
    const int s0 = size0();
    Mat M;
    const Mat R0 = rows(Z, n0, d * n0);
    if (quadratic) M = mul(U, Z);
    else {
      M = mul(V, R0);
      const Mat Dft = rows(Df, 0, n0);
      const Mat NtDf = mulT(N, Dft);
      for (int c = 0; c < d; ++c) for (int i = 0; i < d * n0; ++i) M(i, c) += NtDf(i, c) - Df(n0 + i, c);
    }
    const Mat R = so_project(M, d);
    Mat dR = R;
    axpy(dR, -1.0, R0);
    const Mat NdR = mul(N, dR);
    Mat out(s0, d);
    for (int c = 0; c < d; ++c) {
      for (int i = 0; i < n0; ++i) out(i, c) = Z(i, c) - NdR(i, c) - T[i] * Df(i, c);
      std::memcpy(out.col(c) + n0, R.col(c), sizeof(double) * d * n0);
    }
    return out;
  }
  Mat precondition(const Mat &Y, const Mat &V_) const {                 // DPGOProblem.cpp:579-598
    Mat u(V_.r, d);
    const int n = n0;
    if (o.precon == 1) {
      for (int c = 0; c < d; ++c) for (int i = 0; i < d * n; ++i) u(i, c) = pinv[i] * V_(i, c);
    } else {
      for (int i = 0; i < n; ++i)
        for (int r = 0; r < d; ++r)
          for (int c = 0; c < d; ++c) {
            double s = 0.0;
            for (int k = 0; k < d; ++k) s += pinv[(size_t)i * d * d + r * d + k] * V_(d * i + k, c);
            u(d * i + r, c) = s;
          }
    }
    return so_proj(rows(Y, n0, d * n0), u, d);
  }
  Mat reduced_hess(const Mat &Y, const Mat &nab, const Mat &Ydot) const {   // DPGOProblem.cpp:552-577
    const Mat R = rows(Y, n0, d * n0);
    Mat tdot = mul(G01, Ydot);
    L.solve(tdot);
    for (double &x : tdot.v) x = -x;
    Mat E = mul(G10, tdot);
    axpy(E, 1.0, mul(G11, Ydot));
    axpy(E, -1.0, sym_block_diag_product(Ydot, R, nab, d));
    return so_proj(R, E, d);
  }
};

struct NodeState {
  bool updated = true;
  int iters = 0, hits[2] = {0, 0}, num_osc = 0;
  std::vector<int> osc;
  double gamma = 0, s_cur = 1, s_next = 1, Fk[2] = {0, 0}, Gk = 0, fobj = 0, fobj_prev = 0, f = 0, fobjE = 0, gradFnorm = 0;
  Mat Xk, Xak, Xakh, X_cur, X_prev, g_cur, g_prev, Df_cur, Df_prev, DfobjE;
  bool refined = false;
  int tcg = 0, restarts = 0;
};

struct Edges {
  long E = 0;
  std::vector<long> i, j;
  std::vector<double> R, t, kappa, tau;
  std::vector<unsigned char> inter;
};

struct Driver {
  int d = 3, A = 0, algorithm = 0;
  long N = 0;
  Options o;
  std::vector<Problem> pb;
  std::vector<NodeState> st;
  std::vector<long> first;
  std::vector<std::vector<long>> own_gid, nbr_gid;
  Edges ed;
  Mat Xk, Xkh, Xkp;               // AMM-PGO* global iterates
  double F = 0, fobj = 0;
  int global_restarts = 0;
  bool par_nodes = false;

  template <class Fn> void for_nodes(Fn fn) {
    if (par_nodes) {
      const bool save = g_par_ops;
      g_par_ops = false;
#pragma omp parallel for schedule(dynamic, 1)
      for (int a = 0; a < A; ++a) fn(a);
      g_par_ops = save;
    } else {
      for (int a = 0; a < A; ++a) fn(a);
    }
  }

  // ---- global objective (DPGOStar::evaluate_f, DPGOStar.cpp:713-761), edge by edge
  double evaluate_f(const Mat &X) const {
    double f = 0.0;
#pragma omp parallel for reduction(+ : f) schedule(static)
    for (long e = 0; e < ed.E; ++e) {
      const long pi = ed.i[e], pj = ed.j[e];
      const double *R = &ed.R[(size_t)e * d * d], *t = &ed.t[(size_t)e * d];
      double et = 0.0, er = 0.0;
      for (int k = 0; k < d; ++k) {
        double s = X(pi, k) - X(pj, k);
        for (int c = 0; c < d; ++c) s += t[c] * X(N + d * pi + c, k);
        et += s * s;
      }
      if (o.loss == 0) {
        double ni = 0, nj = 0, cr = 0;
        for (int r = 0; r < d; ++r)
          for (int k = 0; k < d; ++k) {
            double s = 0.0;
            for (int c = 0; c < d; ++c) s += R[c * d + r] * X(N + d * pi + c, k);
            cr += s * X(N + d * pj + r, k);
            ni += X(N + d * pi + r, k) * X(N + d * pi + r, k);
            nj += X(N + d * pj + r, k) * X(N + d * pj + r, k);
          }
        f += 0.5 * (ed.tau[e] * et + ed.kappa[e] * (ni + nj - 2.0 * cr));
      } else {
        for (int r = 0; r < d; ++r)
          for (int k = 0; k < d; ++k) {
            double s = -X(N + d * pj + r, k);
            for (int c = 0; c < d; ++c) s += R[c * d + r] * X(N + d * pi + c, k);
            er += s * s;
          }
        const double e2 = ed.tau[e] * et + ed.kappa[e] * er, delta = o.loss_reg;
        if (!ed.inter[e]) f += 0.5 * e2;
        else if (o.loss == 1) f += 0.5 * std::min(2 * std::sqrt(delta) * std::sqrt(std::max(e2, delta)) - delta, e2);
        else if (o.loss == 2) f += 0.5 * delta * (e2 / (e2 + delta));
        else f += 0.5 * (delta - delta * std::exp(-e2 / delta));
      }
    }
    return f;
  }

  // Z = [t; R; t_nbr; R_nbr] of node a from a global X (dist_pgo.cpp:436-446, DPGO_utils.h:397-453)
  void scatter(const Mat &X, int a, Mat &Z, bool own) const {
    const Problem &p = pb[a];
    const int n0 = p.n0, n1 = p.n1;
    if (Z.r == 0) Z = Mat((d + 1) * (n0 + n1), d);
    for (int c = 0; c < d; ++c) {
      if (own)
        for (int i = 0; i < n0; ++i) {
          const long g = own_gid[a][i];
          Z(i, c) = X(g, c);
          for (int r = 0; r < d; ++r) Z(n0 + d * i + r, c) = X(N + d * g + r, c);
        }
      const int base = (d + 1) * n0;
      for (int i = 0; i < n1; ++i) {
        const long g = nbr_gid[a][i];
        Z(base + i, c) = X(g, c);
        for (int r = 0; r < d; ++r) Z(base + n1 + d * i + r, c) = X(N + d * g + r, c);
      }
    }
  }
  void put(Mat &Xg, int a, const Mat &Xa) const {      // DPGOStar.cpp:541-547
    const int n0 = pb[a].n0;
    const long i0 = first[a];
    for (int c = 0; c < d; ++c) {
      std::memcpy(Xg.col(c) + i0, Xa.col(c), sizeof(double) * n0);
      std::memcpy(Xg.col(c) + N + d * i0, Xa.col(c) + n0, sizeof(double) * d * n0);
    }
  }

  TntOut run_tnt(int a, const Mat &x0, const Mat &g, double f) {
    Problem &p = pb[a];
    const int n0 = p.n0;
    TntProblem t;
    t.f = [&](const Mat &Y) { return p.evaluate_G(Y, g, f); };
    t.grad = [&](const Mat &Y, Mat &nab) {
      nab = mul(p.Grot, Y);                                           // DPGOProblem.h:370-382
      for (int c = 0; c < d; ++c) for (int i = 0; i < d * n0; ++i) nab(i, c) += g(n0 + i, c);
      return so_proj(rows(Y, n0, d * n0), nab, d);
    };
    t.hess = [&](const Mat &Y, const Mat &nab, const Mat &v) { return p.reduced_hess(Y, nab, v); };
    t.retract = [&](const Mat &Y, const Mat &h) {                     // DPGOProblem.cpp:127-143
      Mat Rn = rows(Y, n0, d * n0);
      axpy(Rn, 1.0, h);
      const Mat Rp = so_project(Rn, d);
      const Mat tp = p.recover_translations(Rp, g);
      Mat out(p.size0(), d);
      set_rows(out, 0, tp);
      set_rows(out, n0, Rp);
      return out;
    };
    t.has_precon = o.precon != 0;
    t.precon = [&](const Mat &Y, const Mat &v) { return p.precondition(Y, v); };
    TntOut r = tnt(t, x0, o);
    st[a].tcg += r.inner;
    return r;
  }

  void grad_norm(int a, const Mat &Dfobj) {
    NodeState &s = st[a];
    const Problem &p = pb[a];
    const Mat gr = so_proj(rows(s.Xak, p.n0, d * p.n0), rows(Dfobj, p.n0, d * p.n0), d);
    double g2 = tr(gr, gr);
    for (int c = 0; c < d; ++c) for (int i = 0; i < p.n0; ++i) g2 += Dfobj(i, c) * Dfobj(i, c);
    s.gradFnorm = std::sqrt(g2);
  }

  // DPGOHash::update (DPGOHash.cpp:84-228) / DPGOStar::update_n (DPGOStar.cpp:315-390)
  void update_n(int a) {
    NodeState &s = st[a];
    Problem &p = pb[a];
    if (s.updated) return;
    const int it = s.iters, s0 = p.size0();
    const bool star = algorithm == 1;
    const bool first = star || it == 0;
    const Mat &Z = s.Xk;
    Mat g, Dfobj;
    double f, fobj;
    if (p.quadratic) {
      g = mul(p.S, Z);
      if (first) {
        f = 0.5 * tr(Z, mul(p.P0, Z));                                // DPGOProblem.cpp:269-287
        fobj = p.evaluate_G(s.Xak, g, f);
      } else {
        Mat Y = Z;                                                    // :516-542
        axpy(Y, -1.0, s.X_cur);
        fobj = s.Gk + 0.5 * tr(Y, mul(p.Q, Y));
        f = fobj + 0.5 * tr(Z, mul(p.P, Z));
      }
      Dfobj = mul(p.G, s.Xak);
      axpy(Dfobj, 1.0, g);
    } else {
      const Mat X = rows(Z, 0, s0);
      if (first) {                                                    // :222-267
        Mat DfE; double fE;
        p.evaluate_E(Z, DfE, fE);
        g = rows(DfE, 0, s0);
        Mat temp = mul(p.D, X);
        axpy(g, -1.0, temp);
        for (size_t i = 0; i < temp.v.size(); ++i) temp.v[i] = 0.5 * temp.v[i] - DfE(i % s0, (int)(i / s0));
        f = 0.5 * fE + tr(X, temp);
        temp = mul(p.G, X);
        Dfobj = g;
        axpy(Dfobj, 1.0, temp);
        for (size_t i = 0; i < temp.v.size(); ++i) temp.v[i] = 0.5 * temp.v[i] + g.v[i];
        fobj = f + tr(X, temp);
        s.DfobjE = DfE; s.fobjE = fE;
      } else {                                                        // :360-424
        Mat Y = Z;
        axpy(Y, -1.0, s.X_cur);
        Mat temp = mul(p.Q, Y);
        for (size_t i = 0; i < temp.v.size(); ++i) temp.v[i] = s.DfobjE.v[i] + 0.5 * temp.v[i];
        fobj = s.Gk - 0.5 * s.fobjE - 0.5 * tr(Y, temp);
        Mat DfE; double fE;
        p.evaluate_E(Z, DfE, fE);
        fobj += 0.5 * fE;
        g = rows(DfE, 0, s0);
        axpy(g, -1.0, mul(p.D, X));
        temp = mul(p.G, X);
        Dfobj = g;
        axpy(Dfobj, 1.0, temp);
        for (size_t i = 0; i < temp.v.size(); ++i) temp.v[i] = 0.5 * temp.v[i] + g.v[i];
        f = fobj - tr(X, temp);
        s.DfobjE = DfE; s.fobjE = fE;
      }
    }
    if (star) s.Gk = fobj;
    if (it == 0) { s.Fk[0] = s.Fk[1] = fobj; s.Gk = fobj; }
    grad_norm(a, Dfobj);
    if (it > 0) { s.X_prev.v.swap(s.X_cur.v); s.X_prev.r = s.X_cur.r; s.X_prev.c = s.X_cur.c; s.g_prev = std::move(s.g_cur); s.Df_prev = std::move(s.Df_cur); }
    s.fobj_prev = s.fobj;
    s.X_cur = Z; s.g_cur = std::move(g); s.Df_cur = std::move(Dfobj); s.fobj = fobj; s.f = f;
    if (o.scheme == 1) {
      if (it == 0) { s.s_cur = 1.0; s.osc.assign(1, 1); }
      else s.s_cur = s.s_next;
      s.s_next = 0.5 + 0.5 * std::sqrt(4.0 * s.s_cur * s.s_cur + 1.0);
      s.gamma = (s.s_cur - 1.0) / s.s_next;
      if (!star) {
        if (fobj <= s.Fk[1]) s.hits[0] = s.hits[0] > 2 ? s.hits[0] - 2 : 0; else s.hits[0]++;
        if (it > 0) {
          if (fobj <= s.fobj_prev) { s.hits[1] = 0; s.osc.push_back(1); } else { s.hits[1]++; s.osc.push_back(0); }
          s.num_osc += s.osc[it] != s.osc[it - 1];
        }
        if (it > o.oscillation_cnt_period) { const int k = it - o.oscillation_cnt_period; s.num_osc -= s.osc[k] != s.osc[k - 1]; }
        s.Fk[0] = s.Fk[0] * (1 - o.eta[0]) + fobj * o.eta[0];
        s.Fk[1] = std::max(fobj, s.Fk[1] * (1 - o.eta[1]) + fobj * o.eta[1]);
      }
    } else if (!star) s.Fk[0] = s.Fk[1] = fobj;
    if (star) s.Fk[0] = s.Fk[1] = fobj;
    s.updated = true;
  }

  // extrapolated point, g and Df of amm_pgo / amm_pgo_n (DPGOHash.cpp:255-264)
  void extrapolate(int a, Mat &Y, Mat &g, Mat &Df) {
    NodeState &s = st[a];
    Problem &p = pb[a];
    if (s.iters == 0) { Y = s.Xk; g = s.g_cur; Df = s.Df_cur; return; }
    Y = s.X_cur;
    for (size_t i = 0; i < Y.v.size(); ++i) Y.v[i] += s.gamma * (s.X_cur.v[i] - s.X_prev.v[i]);
    if (p.quadratic) {
      g = s.g_cur; Df = s.Df_cur;
      for (size_t i = 0; i < g.v.size(); ++i) { g.v[i] += s.gamma * (s.g_cur.v[i] - s.g_prev.v[i]); Df.v[i] += s.gamma * (s.Df_cur.v[i] - s.Df_prev.v[i]); }
    } else {
      g = p.evaluate_g(Y);
      Df = mul(p.G, Y);          // G has (d+1) n0 columns: only the own rows of Y enter
      axpy(Df, 1.0, g);
    }
  }
  void set_rot_and_recover(int a, Mat &Xak, const Mat &Xakh, const Mat &g) {
    const Problem &p = pb[a];
    const Mat R = rows(Xakh, p.n0, d * p.n0);
    set_rows(Xak, p.n0, R);
    set_rows(Xak, 0, p.recover_translations(R, g));
  }

  // DPGOHash::amm_pgo (DPGOHash.cpp:230-444)
  void hash_amm(int a) {
    NodeState &s = st[a];
    Problem &p = pb[a];
    Mat Y, g, Df;
    extrapolate(a, Y, g, Df);
    const double f = s.f, fobj_k = s.fobj;
    const Mat &gk = s.g_cur;
    const bool refined = (((s.gradFnorm * s.gradFnorm / fobj_k) > o.accepted_delta) || (s.num_osc >= o.max_oscillations)) &&
                         o.max_iterations > 0 && o.max_iterations_accepted > 0;
    s.refined = refined;
    s.Xakh = p.proximal(Y, Df);
    double Gkh = p.evaluate_G(s.Xakh, gk, f);
    Mat diff = s.Xakh;
    axpy(diff, -1.0, s.Xak);
    const double minG = s.Fk[0] - o.psi * tr(diff, diff);
    set_rot_and_recover(a, s.Xak, s.Xakh, g);
    if (refined) s.Xak = run_tnt(a, s.Xak, g, f).x;
    s.Gk = p.evaluate_G(s.Xak, gk, f);
    if (Gkh > minG) { s.Xakh = p.proximal(s.Xk, s.Df_cur); Gkh = p.evaluate_G(s.Xakh, gk, f); }
    const bool hard = s.Gk > s.Fk[0];
    const bool soft = (s.Gk > s.Fk[1] && s.hits[0] >= o.max_soft_restart_hits[0]) ||
                      (s.Gk > fobj_k && s.hits[1] > o.max_soft_restart_hits[1]);
    const Mat *gcur = &g;
    if (hard || soft) {
      s.restarts++;
      gcur = &gk;
      if (Gkh <= fobj_k) s.Xak = s.Xakh; else s.Xak = p.proximal(s.Xk, s.Df_cur);
      set_rows(s.Xak, 0, p.recover_translations(rows(s.Xak, p.n0, d * p.n0), gk));
      if (refined) { TntOut r = run_tnt(a, s.Xak, gk, f); s.Xak = r.x; s.Gk = r.f; }
      else s.Gk = p.evaluate_G(s.Xak, gk, f);
      if (hard) s.s_next = std::max(0.5 * s.s_next, 1.0);
      s.hits[0] /= 3; s.hits[1] = 0;
    }
    if ((s.Fk[0] - s.Gk) < o.phi * (s.Fk[0] - Gkh)) {
      set_rot_and_recover(a, s.Xak, s.Xakh, *gcur);
      s.Gk = p.evaluate_G(s.Xak, gk, f);
    }
  }
  // DPGOHash::mm_pgo (DPGOHash.cpp:446-581)
  void hash_mm(int a) {
    NodeState &s = st[a];
    Problem &p = pb[a];
    const bool refined = ((s.gradFnorm * s.gradFnorm / s.fobj) > o.accepted_delta) && o.max_iterations > 0 && o.max_iterations_accepted > 0;
    s.refined = refined;
    s.Xakh = p.proximal(s.Xk, s.Df_cur);
    set_rows(s.Xakh, 0, p.recover_translations(rows(s.Xakh, p.n0, d * p.n0), s.g_cur));
    if (refined) { TntOut r = run_tnt(a, s.Xakh, s.g_cur, s.f); s.Xak = r.x; s.Gk = r.f; }
    else { s.Xak = s.Xakh; s.Gk = p.evaluate_G(s.Xak, s.g_cur, s.f); }
  }
  void finish_n(int a) {
    NodeState &s = st[a];
    s.iters++;
    for (int c = 0; c < d; ++c) std::memcpy(s.Xk.col(c), s.Xak.col(c), sizeof(double) * pb[a].size0());
    s.updated = false;
  }

  // DPGOStar::amm_pgo_n / mm_pgo_n / pm_pgo_n (DPGOStar.cpp:392-711)
  void star_amm_n(int a) {
    NodeState &s = st[a];
    Problem &p = pb[a];
    Mat Y, g, Df;
    extrapolate(a, Y, g, Df);
    s.refined = (s.gradFnorm * s.gradFnorm / s.fobj) > o.accepted_delta;
    s.Xakh = p.proximal(Y, Df);
    set_rot_and_recover(a, s.Xak, s.Xakh, g);
    if (s.refined && o.max_iterations > 0 && o.max_iterations_accepted > 0) s.Xak = run_tnt(a, s.Xak, g, s.f).x;
    put(Xkh, a, s.Xakh);
    put(Xkp, a, s.Xak);
  }
  void star_mm_n(int a) {
    NodeState &s = st[a];
    Problem &p = pb[a];
    set_rot_and_recover(a, s.Xak, s.Xakh, s.g_cur);
    if (s.refined && o.max_iterations > 0 && o.max_iterations_accepted > 0) { TntOut r = run_tnt(a, s.Xak, s.g_cur, s.f); s.Xak = r.x; s.Gk = r.f; }
    else s.Gk = p.evaluate_G(s.Xak, s.g_cur, s.f);
    put(Xkp, a, s.Xak);
  }
  double dist2(const Mat &A_, const Mat &B_) const {
    double s = 0.0;
    const size_t n = A_.v.size();
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (long i = 0; i < (long)n; ++i) { const double t = A_.v[i] - B_.v[i]; s += t * t; }
    return s;
  }
  // DPGOStar::iterate (DPGOStar.cpp:126-213)
  void star_iterate() {
    for_nodes([&](int a) { star_amm_n(a); });
    double fobjh = evaluate_f(Xkh);
    if (fobjh > F - o.psi * dist2(Xkh, Xk)) {
      for_nodes([&](int a) { st[a].Xakh = pb[a].proximal(st[a].Xk, st[a].Df_cur); put(Xkh, a, st[a].Xakh); });
      fobjh = evaluate_f(Xkh);
    }
    double fo = evaluate_f(Xkp);
    if (fo > F - o.psi * dist2(Xkp, Xk)) {
      global_restarts++;
      for_nodes([&](int a) { star_mm_n(a); st[a].s_next = std::max(0.5 * st[a].s_next, 1.0); });
      fo = evaluate_f(Xkp);
    }
    if (F - fo < o.phi * (F - fobjh)) {
      for_nodes([&](int a) { set_rot_and_recover(a, st[a].Xak, st[a].Xakh, st[a].g_cur); put(Xkp, a, st[a].Xak); });
      fo = evaluate_f(Xkp);
    }
    for_nodes([&](int a) { finish_n(a); });
    Xk.v.swap(Xkp.v);
    fobj = fo;
    F = F * (1 - o.eta[0]) + fo * o.eta[0];
  }
};

}  // namespace

// ---- C ABI (ctypes: oracle/cpu_ref.py) -----------------------------------------------------
extern "C" {

void *cpu_dpgo_create(int d, int num_nodes, long num_poses, int algorithm, int scheme, int loss, int precon,
                      const double *opts /* xi, loss_reg, accepted_delta, eta0, eta1, psi, phi, hits0, hits1, osc_period,
                                            max_osc, grad_tol, pgrad_tol, rel_tol, step_tol, max_it, max_acc, max_tcg,
                                            kappa, theta */) {
  Driver *h = new Driver();
  h->d = d; h->A = num_nodes; h->N = num_poses; h->algorithm = algorithm;
  Options &o = h->o;
  o.scheme = scheme; o.loss = loss; o.precon = precon;
  o.xi = opts[0]; o.loss_reg = opts[1]; o.accepted_delta = opts[2]; o.eta[0] = opts[3]; o.eta[1] = opts[4];
  o.psi = opts[5]; o.phi = opts[6]; o.max_soft_restart_hits[0] = (int)opts[7]; o.max_soft_restart_hits[1] = (int)opts[8];
  o.oscillation_cnt_period = (int)opts[9]; o.max_oscillations = (int)opts[10]; o.grad_norm_tol = opts[11];
  o.precon_grad_norm_tol = opts[12]; o.rel_func_decrease_tol = opts[13]; o.stepsize_tol = opts[14];
  o.max_iterations = (int)opts[15]; o.max_iterations_accepted = (int)opts[16]; o.max_tCG = (int)opts[17];
  o.kappa = opts[18]; o.theta = opts[19];
  h->pb.resize(num_nodes); h->st.resize(num_nodes); h->first.resize(num_nodes);
  h->own_gid.resize(num_nodes); h->nbr_gid.resize(num_nodes);
  for (auto &p : h->pb) { p.d = d; p.o = o; p.quadratic = loss == 0; }
  return h;
}
void cpu_dpgo_destroy(void *hh) { delete static_cast<Driver *>(hh); }

int cpu_dpgo_set_matrix(void *hh, int node, const char *name, int rows_, int cols_, const int *ptr, const int *idx,
                        const double *val) {
  Driver *h = static_cast<Driver *>(hh);
  Problem &p = h->pb[node];
  const std::string nm(name);
  Csr *m = nm == "G" ? &p.G : nm == "G01" ? &p.G01 : nm == "G10" ? &p.G10 : nm == "G11" ? &p.G11 : nm == "S" ? &p.S
         : nm == "P" ? &p.P : nm == "P0" ? &p.P0 : nm == "Q" ? &p.Q : nm == "D" ? &p.D : nm == "B1" ? &p.B1
         : nm == "U" ? &p.U : nm == "N" ? &p.N : nm == "V" ? &p.V : nullptr;
  if (!m) return -1;
  m->rows = rows_; m->cols = cols_;
  m->ptr.assign(ptr, ptr + rows_ + 1);
  m->idx.assign(idx, idx + ptr[rows_]);
  m->val.assign(val, val + ptr[rows_]);
  return 0;
}

// own_gid / nbr_gid: global pose ids of the node's own poses and neighbour copies; G00 as CSR with a
// fill-reducing ordering chol_perm (perm[new] = old); T: diagonal of T; returns -2 if G00 is not SPD
int cpu_dpgo_set_node(void *hh, int node, int n0, int n1, int m1, long first_gid, const long *own_gid, const long *nbr_gid,
                      const double *T, const int *g00_ptr, const int *g00_idx, const double *g00_val, const int *chol_perm) {
  Driver *h = static_cast<Driver *>(hh);
  Problem &p = h->pb[node];
  const int d = h->d;
  p.n0 = n0; p.n1 = n1; p.m1 = m1;
  h->first[node] = first_gid;
  h->own_gid[node].assign(own_gid, own_gid + n0);
  h->nbr_gid[node].assign(nbr_gid, nbr_gid + n1);
  p.T.assign(T, T + n0);
  Csr G00;
  G00.rows = G00.cols = n0;
  G00.ptr.assign(g00_ptr, g00_ptr + n0 + 1);
  G00.idx.assign(g00_idx, g00_idx + g00_ptr[n0]);
  G00.val.assign(g00_val, g00_val + g00_ptr[n0]);
  if (!p.L.factor(G00, chol_perm)) return -2;
  // rotation rows of G (reduced Euclidean gradient, DPGOProblem.h:370-382)
  p.Grot.rows = d * n0; p.Grot.cols = p.G.cols;
  p.Grot.ptr.assign(d * n0 + 1, 0);
  for (int i = 0; i < d * n0; ++i) p.Grot.ptr[i + 1] = p.G.ptr[n0 + i + 1] - p.G.ptr[n0];
  p.Grot.idx.assign(p.G.idx.begin() + p.G.ptr[n0], p.G.idx.end());
  p.Grot.val.assign(p.G.val.begin() + p.G.ptr[n0], p.G.val.end());
  // preconditioner: Jacobi (DPGOProblem.cpp:96-98) or the d x d block-Jacobi of the device path
  const Options &o = h->o;
  if (o.precon == 1) {
    p.pinv.assign((size_t)d * n0, 0.0);
    for (int i = 0; i < d * n0; ++i)
      for (int q = p.G11.ptr[i]; q < p.G11.ptr[i + 1]; ++q) if (p.G11.idx[q] == i) p.pinv[i] = 1.0 / p.G11.val[q];
  } else if (o.precon == 2) {
    p.pinv.assign((size_t)d * d * n0, 0.0);
    for (int i = 0; i < n0; ++i) {
      double B[9] = {0}, a[9], b[9];
      for (int r = 0; r < d; ++r)
        for (int q = p.G11.ptr[d * i + r]; q < p.G11.ptr[d * i + r + 1]; ++q) {
          const int c = p.G11.idx[q];
          if (c / d == i) B[r * d + c % d] += p.G11.val[q];
        }
      for (int k = 0; k < d * d; ++k) { a[k] = B[k]; b[k] = 0.0; }
      for (int k = 0; k < d; ++k) b[k * d + k] = 1.0;
      for (int c = 0; c < d; ++c) {
        int piv = c;
        for (int r = c + 1; r < d; ++r) if (std::fabs(a[r * d + c]) > std::fabs(a[piv * d + c])) piv = r;
        if (piv != c) for (int k = 0; k < d; ++k) { std::swap(a[c * d + k], a[piv * d + k]); std::swap(b[c * d + k], b[piv * d + k]); }
        const double iv = 1.0 / a[c * d + c];
        for (int k = 0; k < d; ++k) { a[c * d + k] *= iv; b[c * d + k] *= iv; }
        for (int r = 0; r < d; ++r) if (r != c) {
          const double f = a[r * d + c];
          for (int k = 0; k < d; ++k) { a[r * d + k] -= f * a[c * d + k]; b[r * d + k] -= f * b[c * d + k]; }
        }
      }
      for (int k = 0; k < d * d; ++k) p.pinv[(size_t)i * d * d + k] = b[k];
    }
  }
  return 0;
}

int cpu_dpgo_set_edges(void *hh, long E, const long *i, const long *j, const double *R, const double *t, const double *kappa,
                       const double *tau, const unsigned char *inter) {
  Driver *h = static_cast<Driver *>(hh);
  const int d = h->d;
  Edges &e = h->ed;
  e.E = E;
  e.i.assign(i, i + E); e.j.assign(j, j + E);
  e.R.assign(R, R + E * d * d); e.t.assign(t, t + E * d);
  e.kappa.assign(kappa, kappa + E); e.tau.assign(tau, tau + E);
  e.inter.assign(inter, inter + E);
  return 0;
}

// mode 0: OpenMP inside the operators, nodes serial (reference); mode 1: nodes in parallel
void cpu_dpgo_set_threads(void *hh, int mode, int nthreads) {
  Driver *h = static_cast<Driver *>(hh);
  if (nthreads > 0) omp_set_num_threads(nthreads);
  h->par_nodes = mode == 1;
  g_par_ops = mode == 0;
}

int cpu_dpgo_initialize(void *hh, const double *X) {
  Driver *h = static_cast<Driver *>(hh);
  const int d = h->d;
  const long rowsX = (long)(d + 1) * h->N;
  Mat Xg((int)rowsX, d);
  std::memcpy(Xg.v.data(), X, sizeof(double) * rowsX * d);
  for (int a = 0; a < h->A; ++a) {
    NodeState &s = h->st[a];
    s = NodeState();
    h->scatter(Xg, a, s.Xk, true);
    s.Xak = rows(s.Xk, 0, h->pb[a].size0());
    s.Xakh = Mat(h->pb[a].size0(), d);
    s.updated = false;
  }
  h->Xk = Xg;
  h->Xkh = Mat((int)rowsX, d);
  h->Xkp = Xg;
  h->global_restarts = 0;
  if (h->algorithm == 1) { h->fobj = h->evaluate_f(h->Xk); h->F = h->fobj; }
  return 0;
}
int cpu_dpgo_update(void *hh) {
  Driver *h = static_cast<Driver *>(hh);
  h->for_nodes([&](int a) { h->update_n(a); });
  return 0;
}
int cpu_dpgo_iterate(void *hh) {
  Driver *h = static_cast<Driver *>(hh);
  if (h->algorithm == 1) { h->star_iterate(); return 0; }
  h->for_nodes([&](int a) {
    if (h->o.scheme == 1) h->hash_amm(a); else h->hash_mm(a);
    h->finish_n(a);
    h->put(h->Xk, a, h->st[a].Xak);           // dist_pgo.cpp:502-511 gathers the global X after iterate()
  });
  return 0;
}
// DPGOHash::communicate (DPGOHash.h:28-86) / DPGOStar::communicate (DPGOStar.cpp:215-223, 276-313)
int cpu_dpgo_communicate(void *hh) {
  Driver *h = static_cast<Driver *>(hh);
  h->for_nodes([&](int a) { h->scatter(h->Xk, a, h->st[a].Xk, false); h->st[a].updated = false; });
  return 0;
}
int cpu_dpgo_get_X(void *hh, double *X) {
  Driver *h = static_cast<Driver *>(hh);
  std::memcpy(X, h->Xk.v.data(), sizeof(double) * h->Xk.v.size());
  return 0;
}
// per node: fobj, gradFnorm, refined, tcg iterations, restarts
int cpu_dpgo_node_scalars(void *hh, double *out) {
  Driver *h = static_cast<Driver *>(hh);
  for (int a = 0; a < h->A; ++a) {
    const NodeState &s = h->st[a];
    out[5 * a] = s.fobj; out[5 * a + 1] = s.gradFnorm; out[5 * a + 2] = s.refined; out[5 * a + 3] = s.tcg;
    out[5 * a + 4] = s.restarts;
  }
  return 0;
}
double cpu_dpgo_evaluate_f(void *hh, const double *X) {
  Driver *h = static_cast<Driver *>(hh);
  Mat Xg((int)((long)(h->d + 1) * h->N), h->d);
  std::memcpy(Xg.v.data(), X, sizeof(double) * Xg.v.size());
  return h->evaluate_f(Xg);
}
int cpu_dpgo_weights(void *hh, int node, double *w) {
  Driver *h = static_cast<Driver *>(hh);
  std::copy(h->pb[node].w.begin(), h->pb[node].w.end(), w);
  return (int)h->pb[node].w.size();
}
long cpu_dpgo_chol_nnz(void *hh, int node) { return (long)static_cast<Driver *>(hh)->pb[node].L.lx.size(); }
int cpu_dpgo_star_restarts(void *hh) { return static_cast<Driver *>(hh)->global_restarts; }
int cpu_dpgo_max_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
