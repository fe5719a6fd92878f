"""SO(d)^n geometry (oracle; test infrastructure only).

Restates C++/DPGO/include/DPGO/SOdProduct.h:39-116 and the batched polar
projections: project_to_SO2 (include/DPGO/internal/project_to_SO2.h:3-18),
project_to_SO3 (src/internal/project_to_SOd.cpp:121-196 with the macros of
include/DPGO/internal/svd3x3.h: COMPUTE_ATA :3-27, JACOBI_CONJUATION :29-109,
COMPUTE_MATRIX_V :111-152, MULTIPLY_WITH_V :197-234, SORT_SINGULAR_VALUES
:236-386, QR :388-463) and PROJECT_TO_SO3_COMPUTE_U (project_to_SO3.h:5-41),
operation for operation (numpy has no fused multiply-add: fma(a,b,c) is
evaluated as a*b+c, a <=1 ulp difference per op).
Rotation blocks are stored as d x d row blocks of a (d*n) x d matrix.
"""
from __future__ import annotations

import numpy as np

_TINY = 1.0e-32
_SMALL = 1.0e-16
_SIN_PI8 = 0.5 * np.sqrt(2.0 - np.sqrt(2.0))
_COS_PI8 = 0.5 * np.sqrt(2.0 + np.sqrt(2.0))
_FOUR_GAMMA_SQ = np.sqrt(8.0) + 3.0


def _rsqrt(x):
    return 1.0 / np.sqrt(x)       # project_to_SOd.cpp:113


def project_to_SO2(A):
    """A: (n,2,2) -> (n,2,2)."""
    a11, a12, a21, a22 = A[:, 0, 0], A[:, 0, 1], A[:, 1, 0], A[:, 1, 1]
    c = a11 + a22
    s = a21 - a12
    n2 = c * c
    n2 = s * s + n2
    ok = n2 >= _TINY
    c = np.where(ok, c, 1.0)
    s = np.where(ok, s, 0.0)
    n2 = np.where(ok, n2, 1.0)
    r = _rsqrt(n2)
    u11 = c * r
    u21 = s * r
    U = np.empty_like(A)
    # project_to_SOd.h:33-40
    U[:, 0, 0] = u11; U[:, 0, 1] = -u21
    U[:, 1, 0] = u21; U[:, 1, 1] = u11
    return U


def _jacobi(S, q, a, b, c):
    """One SVD3X3_JACOBI_CONJUATION on the symmetric matrix entries S (dict)
    with index permutation (a,b,c) of (1,2,3): SS11=S[aa], SS21=S[ba],
    SS31=S[ca], SS22=S[bb], SS32=S[cb], SS33=S[cc]; quaternion q=[s,x,y,z]."""
    def key(i, j):
        return (i, j) if i >= j else (j, i)
    k11, k21, k31 = key(a, a), key(b, a), key(c, a)
    k22, k32, k33 = key(b, b), key(c, b), key(c, c)
    S11, S21, S31, S22, S32, S33 = (S[k] for k in (k11, k21, k31, k22, k32, k33))
    qs = q[0]
    QX, QY, QZ = q[a], q[b], q[c]

    sh = S21 * 0.5
    t5 = S11 - S22
    t2 = sh * sh
    m1 = t2 >= _TINY
    sh = np.where(m1, sh, 0.0)
    ch = np.where(m1, t5, 1.0)
    t1 = sh * sh
    t2 = ch * ch
    t3 = t1 + t2
    t4 = _rsqrt(t3)
    sh = t4 * sh
    ch = t4 * ch
    t1 = _FOUR_GAMMA_SQ * t1
    m1 = t2 <= t1
    sh = np.where(m1, _SIN_PI8, sh)
    ch = np.where(m1, _COS_PI8, ch)
    t1 = sh * sh
    t2 = ch * ch
    c_ = t2 - t1
    s_ = ch * sh
    s_ = s_ + s_

    t3 = t1 + t2
    S33 = S33 * t3
    S31 = S31 * t3
    S32 = S32 * t3
    S33 = S33 * t3

    t1 = s_ * S31
    t2 = s_ * S32
    S31 = c_ * S31
    S32 = c_ * S32
    S31 = t2 + S31
    S32 = S32 - t1

    t2 = s_ * s_
    t1 = S22 * t2
    t3 = S11 * t2
    t4 = c_ * c_
    S11 = S11 * t4
    S22 = S22 * t4
    S11 = S11 + t1
    S22 = S22 + t3
    t4 = t4 - t2
    t2 = S21 + S21
    S21 = S21 * t4
    t4 = c_ * s_
    t2 = t2 * t4
    t5 = t5 * t4
    S11 = S11 + t2
    S21 = S21 - t5
    S22 = S22 - t2

    # quaternion update; STMP1..3 alias (tmp for X, Y, Z of this permutation)
    tX = sh * QX
    tY = sh * QY
    tZ = sh * QZ
    sh = sh * qs
    qs = ch * qs
    QX = ch * QX
    QY = ch * QY
    QZ = ch * QZ
    QZ = QZ + sh
    qs = qs - tZ
    QX = QX + tY
    QY = QY - tX

    for k, v in zip((k11, k21, k31, k22, k32, k33), (S11, S21, S31, S22, S32, S33)):
        S[k] = v
    q[0] = qs
    q[a], q[b], q[c] = QX, QY, QZ


def _cswap(mask, x, y):
    return np.where(mask, y, x), np.where(mask, x, y)


def _qr(A, U, piv, npiv, r1, r2):
    """SVD3X3_QR: Givens rotation on rows r1,r2 (0-based) of A with pivot
    entries A[piv], A[npiv]; accumulates into columns r1,r2 of U."""
    ap, an = A[piv], A[npiv]
    sh = an * an
    sh = np.where(sh >= _SMALL, an, 0.0)
    ch = 0.0 - ap
    ch = np.maximum(ch, ap)
    ch = np.maximum(ch, _SMALL)
    m5 = ap >= 0.0
    t1 = ch * ch
    t2 = sh * sh + t1
    t1 = _rsqrt(t2)
    t1 = t1 * t2
    ch = ch + t1
    t1 = ch
    ch = np.where(m5, ch, sh)
    sh = np.where(m5, sh, t1)
    t1 = ch * ch
    t2 = sh * sh + t1
    t1 = _rsqrt(t2)
    ch = ch * t1
    sh = sh * t1
    s_ = sh * sh
    c_ = ch * ch - s_
    s_ = sh * ch
    s_ = s_ + s_
    for col in range(3):
        x, y = A[(r1, col)], A[(r2, col)]
        t1 = s_ * x
        t2 = s_ * y
        x = c_ * x
        y = c_ * y
        A[(r1, col)] = x + t2
        A[(r2, col)] = y - t1
    for row in range(3):
        x, y = U[(row, r1)], U[(row, r2)]
        t1 = s_ * x
        t2 = s_ * y
        x = c_ * x
        y = c_ * y
        U[(row, r1)] = x + t2
        U[(row, r2)] = y - t1


def project_to_SO3(Ain):
    """Ain: (n,3,3) -> (n,3,3) nearest rotation (project_to_SOd.cpp:121-196)."""
    n = Ain.shape[0]
    A = {(i, j): Ain[:, i, j].copy() for i in range(3) for j in range(3)}
    # COMPUTE_ATA (lower triangle, keys (row>=col), 0-based)
    S = {}
    for (i, j) in ((0, 0), (1, 0), (2, 0), (1, 1), (2, 1), (2, 2)):
        v = A[(0, i)] * A[(0, j)]
        v = A[(1, i)] * A[(1, j)] + v
        v = A[(2, i)] * A[(2, j)] + v
        S[(i, j)] = v
    # _jacobi uses 1-based permutation labels mapped onto 0-based keys
    S1 = {(i + 1, j + 1): v for (i, j), v in S.items()}
    q = [np.ones(n), np.zeros(n), np.zeros(n), np.zeros(n)]
    for _ in range(8):
        _jacobi(S1, q, 1, 2, 3)
        _jacobi(S1, q, 2, 3, 1)
        _jacobi(S1, q, 3, 1, 2)
    qs, qx, qy, qz = q
    # COMPUTE_MATRIX_V
    t2 = qs * qs
    t2 = qx * qx + t2
    t2 = qy * qy + t2
    t2 = qz * qz + t2
    t1 = _rsqrt(t2)
    qs, qx, qy, qz = qs * t1, qx * t1, qy * t1, qz * t1
    t1, t2, t3 = qx * qx, qy * qy, qz * qz
    v11 = qs * qs
    v22 = v11 - t1
    v33 = v22 - t2
    v33 = v33 + t3
    v22 = v22 + t2
    v22 = v22 - t3
    v11 = v11 + t1
    v11 = v11 - t2
    v11 = v11 - t3
    t1, t2, t3 = qx + qx, qy + qy, qz + qz
    v32 = qs * t1
    v13 = qs * t2
    v21 = qs * t3
    t1 = qy * t1
    t2 = qz * t2
    t3 = qx * t3
    v12 = t1 - v21
    v23 = t2 - v32
    v31 = t3 - v13
    v21 = t1 + v21
    v32 = t2 + v32
    v13 = t3 + v13
    V = {(0, 0): v11, (0, 1): v12, (0, 2): v13, (1, 0): v21, (1, 1): v22,
         (1, 2): v23, (2, 0): v31, (2, 1): v32, (2, 2): v33}
    # MULTIPLY_WITH_V: A <- A V
    for r in range(3):
        x1, x2, x3 = A[(r, 0)], A[(r, 1)], A[(r, 2)]
        for c in range(3):
            v = V[(0, c)] * x1
            v = V[(1, c)] * x2 + v
            v = V[(2, c)] * x3 + v
            A[(r, c)] = v
    # SORT_SINGULAR_VALUES
    nrm = []
    for c in range(3):
        v = A[(0, c)] * A[(0, c)]
        v = A[(1, c)] * A[(1, c)] + v
        v = A[(2, c)] * A[(2, c)] + v
        nrm.append(v)
    for (ca, cb, neg) in ((0, 1, 1), (0, 2, 0), (1, 2, 2)):
        m = nrm[ca] < nrm[cb]
        for M_ in (A, V):
            for r in range(3):
                M_[(r, ca)], M_[(r, cb)] = _cswap(m, M_[(r, ca)], M_[(r, cb)])
        nrm[ca], nrm[cb] = _cswap(m, nrm[ca], nrm[cb])
        sgn = 1.0 + np.where(m, -2.0, 0.0)
        for M_ in (A, V):
            for r in range(3):
                M_[(r, neg)] = M_[(r, neg)] * sgn
    U = {(i, j): (np.ones(n) if i == j else np.zeros(n))
         for i in range(3) for j in range(3)}
    _qr(A, U, (0, 0), (1, 0), 0, 1)
    _qr(A, U, (0, 0), (2, 0), 0, 2)
    _qr(A, U, (1, 1), (2, 1), 1, 2)
    # PROJECT_TO_SO3_COMPUTE_U: out = U V^T
    out = np.empty_like(Ain)
    for i in range(3):
        for j in range(3):
            v = U[(i, 0)] * V[(j, 0)]
            v = U[(i, 1)] * V[(j, 1)] + v
            v = U[(i, 2)] * V[(j, 2)] + v
            out[:, i, j] = v
    return out


def project_svd(A):
    """project_to_SOdn (DPGO_utils.h:484-512): U V^T with the last column of
    U negated when det(U) det(V) <= 0."""
    U, s, Vt = np.linalg.svd(A)
    det = np.linalg.det(U) * np.linalg.det(Vt)
    U = U.copy()
    U[det <= 0, :, -1] *= -1
    return U @ Vt


def project(M, d):
    """SOdProduct::project (SOdProduct.h:39-55) on a (d*n, d) matrix.
    For fewer than 4 blocks the reference falls back to Eigen::JacobiSVD
    (DPGO_utils.h:524,548)."""
    n = M.shape[0] // d
    A = M.reshape(n, d, d)
    if n < 4:
        return project_svd(A).reshape(n * d, d)
    if d == 2:
        return project_to_SO2(A).reshape(n * d, d)
    return project_to_SO3(A).reshape(n * d, d)


def sym_block_diag_product(A, B, C, d):
    """SOdProduct::SymBlockDiagProduct (SOdProduct.h:64-89):
    P_i = sym(C_i B_i^T) A_i."""
    n = A.shape[0] // d
    A3, B3, C3 = (x.reshape(n, d, d) for x in (A, B, C))
    Gm = C3 @ np.transpose(B3, (0, 2, 1))
    Sm = 0.5 * (Gm + np.transpose(Gm, (0, 2, 1)))
    return (Sm @ A3).reshape(n * d, d)


def proj(Y, V, d):
    """SOdProduct::Proj (SOdProduct.h:96-103)."""
    return V - sym_block_diag_product(Y, Y, V, d)
