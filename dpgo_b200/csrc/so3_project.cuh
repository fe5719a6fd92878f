// Batched polar projection onto SO(2) / SO(3), one matrix per thread.
//
// Follows the operation sequence of the reference's AVX2 kernels so that the
// device path rounds like the CPU path: project_to_SO2
// (C++/DPGO/include/DPGO/internal/project_to_SO2.h:3-18) and project_to_SO3
// (C++/DPGO/src/internal/project_to_SOd.cpp:121-196 with the macros of
// C++/DPGO/include/DPGO/internal/svd3x3.h: A^T A :3-27, 8 sweeps of the
// approximate-Givens Jacobi conjugation :29-109 accumulating V as a
// quaternion, V :111-152, A V :197-234, column sort with sign fix :236-386,
// Givens QR :388-463, U V^T project_to_SO3.h:5-41).  Every multiply / add /
// fma is spelled with a round-to-nearest intrinsic so that nvcc cannot
// contract differently from the reference's explicit mul/add/fma mix, and
// 1/sqrt is the exact division the reference uses (project_to_SOd.cpp:113).
#pragma once

namespace mmpgo {

#define MUL(a, b) __dmul_rn((a), (b))
#define ADD(a, b) __dadd_rn((a), (b))
#define SUB(a, b) __dsub_rn((a), (b))
#define FMA(a, b, c) __fma_rn((a), (b), (c))
#define RSQRT(a) __ddiv_rn(1.0, __dsqrt_rn(a))

__device__ __forceinline__ void project_to_SO2(const double *A, double *U) {
  double c = ADD(A[0], A[3]);
  double s = SUB(A[2], A[1]);
  double n2 = MUL(c, c);
  n2 = FMA(s, s, n2);
  const bool ok = n2 >= 1.0e-32;
  c = ok ? c : 1.0;
  s = ok ? s : 0.0;
  n2 = ok ? n2 : 1.0;
  const double r = RSQRT(n2);
  const double u11 = MUL(c, r), u21 = MUL(s, r);
  U[0] = u11; U[1] = -u21; U[2] = u21; U[3] = u11;
}

// one SVD3X3_JACOBI_CONJUATION on (S11,S21,S31,S22,S32,S33) and quaternion (qs; QX,QY,QZ)
__device__ __forceinline__ void jacobi_conj(double &S11, double &S21, double &S31, double &S22, double &S32,
                                            double &S33, double &qs, double &QX, double &QY, double &QZ) {
  const double kTiny = 1.0e-32;
  const double kFourGammaSq = 5.828427124746190;   // sqrt(8) + 3
  const double kSinPi8 = 0.3826834323650897;       // 0.5 sqrt(2 - sqrt 2) as the reference computes it at run time (traits.cpp:16-17)
  const double kCosPi8 = 0.9238795325112867;       // 0.5 sqrt(2 + sqrt 2)
  double sh = MUL(S21, 0.5);
  double t5 = SUB(S11, S22);
  double t2 = MUL(sh, sh);
  bool m1 = t2 >= kTiny;
  sh = m1 ? sh : 0.0;
  double ch = m1 ? t5 : 1.0;
  double t1 = MUL(sh, sh);
  t2 = MUL(ch, ch);
  double t3 = ADD(t1, t2);
  double t4 = RSQRT(t3);
  sh = MUL(t4, sh);
  ch = MUL(t4, ch);
  t1 = MUL(kFourGammaSq, t1);
  m1 = t2 <= t1;
  sh = m1 ? kSinPi8 : sh;
  ch = m1 ? kCosPi8 : ch;
  t1 = MUL(sh, sh);
  t2 = MUL(ch, ch);
  const double c = SUB(t2, t1);
  double s = MUL(ch, sh);
  s = ADD(s, s);
  t3 = ADD(t1, t2);
  S33 = MUL(S33, t3);
  S31 = MUL(S31, t3);
  S32 = MUL(S32, t3);
  S33 = MUL(S33, t3);
  t1 = MUL(s, S31);
  t2 = MUL(s, S32);
  S31 = MUL(c, S31);
  S32 = MUL(c, S32);
  S31 = ADD(t2, S31);
  S32 = SUB(S32, t1);
  t2 = MUL(s, s);
  t1 = MUL(S22, t2);
  t3 = MUL(S11, t2);
  t4 = MUL(c, c);
  S11 = MUL(S11, t4);
  S22 = MUL(S22, t4);
  S11 = ADD(S11, t1);
  S22 = ADD(S22, t3);
  t4 = SUB(t4, t2);
  t2 = ADD(S21, S21);
  S21 = MUL(S21, t4);
  t4 = MUL(c, s);
  t2 = MUL(t2, t4);
  t5 = MUL(t5, t4);
  S11 = ADD(S11, t2);
  S21 = SUB(S21, t5);
  S22 = SUB(S22, t2);
  const double tX = MUL(sh, QX), tY = MUL(sh, QY), tZ = MUL(sh, QZ);
  sh = MUL(sh, qs);
  qs = MUL(ch, qs);
  QX = MUL(ch, QX);
  QY = MUL(ch, QY);
  QZ = MUL(ch, QZ);
  QZ = ADD(QZ, sh);
  qs = SUB(qs, tZ);
  QX = ADD(QX, tY);
  QY = SUB(QY, tX);
}

__device__ __forceinline__ void cswap(bool m, double &x, double &y) {
  const double a = m ? y : x, b = m ? x : y;
  x = a; y = b;
}

// SVD3X3_QR: Givens rotation of rows (r1, r2) of A with pivots ap = A[r1][pc], an = A[r2][pc]
__device__ __forceinline__ void givens_qr(double (&A)[3][3], double (&U)[3][3], int r1, int r2, int pc) {
  const double kSmall = 1.0e-16;
  const double ap = A[r1][pc], an = A[r2][pc];
  double sh = MUL(an, an);
  sh = (sh >= kSmall) ? an : 0.0;
  double ch = SUB(0.0, ap);
  ch = fmax(ch, ap);
  ch = fmax(ch, kSmall);
  const bool m5 = ap >= 0.0;
  double t1 = MUL(ch, ch);
  double t2 = FMA(sh, sh, t1);
  t1 = RSQRT(t2);
  t1 = MUL(t1, t2);
  ch = ADD(ch, t1);
  t1 = ch;
  ch = m5 ? ch : sh;
  sh = m5 ? sh : t1;
  t1 = MUL(ch, ch);
  t2 = FMA(sh, sh, t1);
  t1 = RSQRT(t2);
  ch = MUL(ch, t1);
  sh = MUL(sh, t1);
  double s = MUL(sh, sh);
  const double c = FMA(ch, ch, -s);
  s = MUL(sh, ch);
  s = ADD(s, s);
#pragma unroll
  for (int col = 0; col < 3; ++col) {
    const double x = A[r1][col], y = A[r2][col];
    const double u1 = MUL(s, x), u2 = MUL(s, y);
    A[r1][col] = ADD(MUL(c, x), u2);
    A[r2][col] = SUB(MUL(c, y), u1);
  }
#pragma unroll
  for (int row = 0; row < 3; ++row) {
    const double x = U[row][r1], y = U[row][r2];
    const double u1 = MUL(s, x), u2 = MUL(s, y);
    U[row][r1] = ADD(MUL(c, x), u2);
    U[row][r2] = SUB(MUL(c, y), u1);
  }
}

__device__ __forceinline__ void project_to_SO3(const double *Ain, double *Out) {
  double A[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A[i][j] = Ain[i * 3 + j];
  // A^T A (lower triangle)
  double S11 = FMA(A[2][0], A[2][0], FMA(A[1][0], A[1][0], MUL(A[0][0], A[0][0])));
  double S21 = FMA(A[2][1], A[2][0], FMA(A[1][1], A[1][0], MUL(A[0][1], A[0][0])));
  double S31 = FMA(A[2][2], A[2][0], FMA(A[1][2], A[1][0], MUL(A[0][2], A[0][0])));
  double S22 = FMA(A[2][1], A[2][1], FMA(A[1][1], A[1][1], MUL(A[0][1], A[0][1])));
  double S32 = FMA(A[2][2], A[2][1], FMA(A[1][2], A[1][1], MUL(A[0][2], A[0][1])));
  double S33 = FMA(A[2][2], A[2][2], FMA(A[1][2], A[1][2], MUL(A[0][2], A[0][2])));
  double qs = 1.0, qx = 0.0, qy = 0.0, qz = 0.0;
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    jacobi_conj(S11, S21, S31, S22, S32, S33, qs, qx, qy, qz);
    jacobi_conj(S22, S32, S21, S33, S31, S11, qs, qy, qz, qx);
    jacobi_conj(S33, S31, S32, S11, S21, S22, qs, qz, qx, qy);
  }
  // V from the quaternion
  double t2 = MUL(qs, qs);
  t2 = FMA(qx, qx, t2);
  t2 = FMA(qy, qy, t2);
  t2 = FMA(qz, qz, t2);
  double t1 = RSQRT(t2);
  qs = MUL(qs, t1); qx = MUL(qx, t1); qy = MUL(qy, t1); qz = MUL(qz, t1);
  t1 = MUL(qx, qx); t2 = MUL(qy, qy);
  double t3 = MUL(qz, qz);
  double V[3][3];
  double v11 = MUL(qs, qs);
  double v22 = SUB(v11, t1);
  double v33 = SUB(v22, t2);
  v33 = ADD(v33, t3);
  v22 = ADD(v22, t2);
  v22 = SUB(v22, t3);
  v11 = ADD(v11, t1);
  v11 = SUB(v11, t2);
  v11 = SUB(v11, t3);
  t1 = ADD(qx, qx); t2 = ADD(qy, qy); t3 = ADD(qz, qz);
  double v32 = MUL(qs, t1), v13 = MUL(qs, t2), v21 = MUL(qs, t3);
  t1 = MUL(qy, t1); t2 = MUL(qz, t2); t3 = MUL(qx, t3);
  const double v12 = SUB(t1, v21), v23 = SUB(t2, v32), v31 = SUB(t3, v13);
  v21 = ADD(t1, v21); v32 = ADD(t2, v32); v13 = ADD(t3, v13);
  V[0][0] = v11; V[0][1] = v12; V[0][2] = v13;
  V[1][0] = v21; V[1][1] = v22; V[1][2] = v23;
  V[2][0] = v31; V[2][1] = v32; V[2][2] = v33;
  // A <- A V
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double x1 = A[r][0], x2 = A[r][1], x3 = A[r][2];
#pragma unroll
    for (int c = 0; c < 3; ++c) A[r][c] = FMA(V[2][c], x3, FMA(V[1][c], x2, MUL(V[0][c], x1)));
  }
  // sort columns by decreasing norm, keeping det(V) = +1
  double n1 = FMA(A[2][0], A[2][0], FMA(A[1][0], A[1][0], MUL(A[0][0], A[0][0])));
  double n2 = FMA(A[2][1], A[2][1], FMA(A[1][1], A[1][1], MUL(A[0][1], A[0][1])));
  double n3 = FMA(A[2][2], A[2][2], FMA(A[1][2], A[1][2], MUL(A[0][2], A[0][2])));
  {
    const bool m = n1 < n2;
#pragma unroll
    for (int r = 0; r < 3; ++r) { cswap(m, A[r][0], A[r][1]); cswap(m, V[r][0], V[r][1]); }
    cswap(m, n1, n2);
    const double sg = ADD(1.0, m ? -2.0 : 0.0);
#pragma unroll
    for (int r = 0; r < 3; ++r) { A[r][1] = MUL(A[r][1], sg); V[r][1] = MUL(V[r][1], sg); }
  }
  {
    const bool m = n1 < n3;
#pragma unroll
    for (int r = 0; r < 3; ++r) { cswap(m, A[r][0], A[r][2]); cswap(m, V[r][0], V[r][2]); }
    cswap(m, n1, n3);
    const double sg = ADD(1.0, m ? -2.0 : 0.0);
#pragma unroll
    for (int r = 0; r < 3; ++r) { A[r][0] = MUL(A[r][0], sg); V[r][0] = MUL(V[r][0], sg); }
  }
  {
    const bool m = n2 < n3;
#pragma unroll
    for (int r = 0; r < 3; ++r) { cswap(m, A[r][1], A[r][2]); cswap(m, V[r][1], V[r][2]); }
    cswap(m, n2, n3);
    const double sg = ADD(1.0, m ? -2.0 : 0.0);
#pragma unroll
    for (int r = 0; r < 3; ++r) { A[r][2] = MUL(A[r][2], sg); V[r][2] = MUL(V[r][2], sg); }
  }
  double U[3][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}};
  givens_qr(A, U, 0, 1, 0);
  givens_qr(A, U, 0, 2, 0);
  givens_qr(A, U, 1, 2, 1);
  // U V^T
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Out[i * 3 + j] = FMA(U[i][2], V[j][2], FMA(U[i][1], V[j][1], MUL(U[i][0], V[j][0])));
}

template <int D> __device__ __forceinline__ void project_to_SOd(const double *A, double *U);
template <> __device__ __forceinline__ void project_to_SOd<2>(const double *A, double *U) { project_to_SO2(A, U); }
template <> __device__ __forceinline__ void project_to_SOd<3>(const double *A, double *U) { project_to_SO3(A, U); }

#undef MUL
#undef ADD
#undef SUB
#undef FMA
#undef RSQRT

}  // namespace mmpgo
