// oracle/_ref harness (TEST INFRASTRUCTURE ONLY): drives the reference's own header-only
// Riemannian truncated-Newton trust region and Steihaug-Toint PCG
// (C++/Optimization/include/Optimization/Riemannian/TNT.h:242-693,
//  C++/Optimization/include/Optimization/LinearAlgebra/IterativeSolvers.h:166-426), included
// from /root/reference, over a plain vector type with C callbacks -- the way DPGO calls it
// (C++/DPGO/src/DPGOHash.cpp:266-381: quadratic-model form, optional preconditioner).
#include <cstddef>
#include <optional>
#include <vector>

#include "Optimization/LinearAlgebra/IterativeSolvers.h"
#include "Optimization/Riemannian/TNT.h"

struct Vec {
  std::vector<double> v;
  Vec() {}
  explicit Vec(size_t n) : v(n, 0.0) {}
};
static Vec operator+(const Vec &a, const Vec &b) { Vec r(a.v.size()); for (size_t i = 0; i < a.v.size(); ++i) r.v[i] = a.v[i] + b.v[i]; return r; }
static Vec operator-(const Vec &a, const Vec &b) { Vec r(a.v.size()); for (size_t i = 0; i < a.v.size(); ++i) r.v[i] = a.v[i] - b.v[i]; return r; }
static Vec operator-(const Vec &a) { Vec r(a.v.size()); for (size_t i = 0; i < a.v.size(); ++i) r.v[i] = -a.v[i]; return r; }
static Vec operator*(double s, const Vec &a) { Vec r(a.v.size()); for (size_t i = 0; i < a.v.size(); ++i) r.v[i] = s * a.v[i]; return r; }
static Vec operator*(const Vec &a, double s) { return s * a; }
static Vec &operator+=(Vec &a, const Vec &b) { for (size_t i = 0; i < a.v.size(); ++i) a.v[i] += b.v[i]; return a; }
static Vec &operator-=(Vec &a, const Vec &b) { for (size_t i = 0; i < a.v.size(); ++i) a.v[i] -= b.v[i]; return a; }
static Vec &operator*=(Vec &a, double s) { for (size_t i = 0; i < a.v.size(); ++i) a.v[i] *= s; return a; }

extern "C" {
// x: a point (nx doubles); v, grad, Hessian / preconditioner output: tangent vectors (nt doubles);
// the retraction writes a point
typedef double (*f_cb)(const double *x);
typedef void (*qm_cb)(const double *x, double *grad);
typedef void (*op_cb)(const double *x, const double *v, double *out);
typedef double (*metric_cb)(const double *x, const double *v1, const double *v2);

struct ref_tnt_params {
  double gradient_tolerance, preconditioned_gradient_tolerance, relative_decrease_tolerance, stepsize_tolerance;
  double Delta0, eta1, eta2, alpha1, alpha2, Delta_tolerance, kappa_fgr, theta;
  long max_iterations, max_iterations_accepted, max_TPCG_iterations;
};
struct ref_tnt_out {
  double f, gradfx_norm;
  long status, iterations, n_inner;
  long inner_iterations[64];
  double gain_ratios[64];
};

int ref_tnt(long nx, long nt, const double *x0, f_cb f, qm_cb qm, op_cb hess, metric_cb metric, op_cb retract, op_cb precon,
            const ref_tnt_params *p, double *x_out, ref_tnt_out *out) {
  using namespace Optimization;
  using namespace Optimization::Riemannian;
  Objective<Vec, double> F = [f](const Vec &x) { return f(x.v.data()); };
  QuadraticModel<Vec, Vec> QM = [qm, hess, nt](const Vec &x, Vec &grad, LinearOperator<Vec, Vec> &H) {
    grad = Vec(nt);
    qm(x.v.data(), grad.v.data());
    H = [hess, nt](const Vec &X, const Vec &V) { Vec r(nt); hess(X.v.data(), V.v.data(), r.v.data()); return r; };
  };
  RiemannianMetric<Vec, Vec, double> M = [metric](const Vec &x, const Vec &a, const Vec &b) {
    return metric(x.v.data(), a.v.data(), b.v.data());
  };
  Retraction<Vec, Vec> R = [retract, nx](const Vec &x, const Vec &h) {
    Vec r(nx); retract(x.v.data(), h.v.data(), r.v.data()); return r;
  };
  std::optional<LinearOperator<Vec, Vec>> P;
  if (precon) P = [precon, nt](const Vec &x, const Vec &v) { Vec r(nt); precon(x.v.data(), v.v.data(), r.v.data()); return r; };
  TNTParams<double> params;
  params.gradient_tolerance = p->gradient_tolerance;
  params.preconditioned_gradient_tolerance = p->preconditioned_gradient_tolerance;
  params.relative_decrease_tolerance = p->relative_decrease_tolerance;
  params.stepsize_tolerance = p->stepsize_tolerance;
  params.Delta0 = p->Delta0; params.eta1 = p->eta1; params.eta2 = p->eta2;
  params.alpha1 = p->alpha1; params.alpha2 = p->alpha2; params.Delta_tolerance = p->Delta_tolerance;
  params.kappa_fgr = p->kappa_fgr; params.theta = p->theta;
  params.max_iterations = (size_t)p->max_iterations;
  params.max_iterations_accepted = (size_t)p->max_iterations_accepted;
  params.max_TPCG_iterations = (size_t)p->max_TPCG_iterations;
  params.verbose = false;
  Vec X0(nx);
  for (long i = 0; i < nx; ++i) X0.v[i] = x0[i];
  try {
    TNTResult<Vec, double> res = TNT<Vec, Vec, double>(F, QM, M, R, X0, P, params);
    for (long i = 0; i < nx; ++i) x_out[i] = res.x.v[i];
    out->f = res.f; out->gradfx_norm = res.gradfx_norm; out->status = (long)res.status;
    out->iterations = (long)res.inner_iterations.size();
    out->n_inner = (long)res.inner_iterations.size();
    for (size_t i = 0; i < res.inner_iterations.size() && i < 64; ++i) out->inner_iterations[i] = (long)res.inner_iterations[i];
    for (size_t i = 0; i < res.gain_ratios.size() && i < 64; ++i) out->gain_ratios[i] = res.gain_ratios[i];
  } catch (const std::exception &) {
    return -1;
  }
  return 0;
}
}
