// K2b, direct path: the level-scheduled supernodal sweeps of the sparse Cholesky solve, ONE
// persistent cooperative launch per solve (data model and algebra: mmpgo_mf.cuh).
//
// Replaces `L_.solve(...)` of the reference (CHOLMOD through Eigen; call sites
// C++/DPGO/include/DPGO/DPGOProblem.h:291, C++/DPGO/src/DPGOProblem.cpp:140, :568).
//
// Execution model.  A stage = all supernodes of equal height (forward) or depth (backward) of
// every active robot node; its tasks are dealt round-robin to the persistent CTAs, stages are
// separated by a grid barrier (2 H + 1 barriers per solve, H = tree height, 14 on the 15 625-pose
// slabs of the 1 M-pose grid), one 1024-thread CTA per SM.  All work is cut into WARP JOBS: a run of
// rows (forward) or columns (backward) of one supernode, whose right-hand side the warp stages in its
// slice of shared memory (in blocks of MF_BLK rows for large fronts).  A lane owns two adjacent
// output rows / columns, so every term is one 128-bit load and a warp reads 512 contiguous bytes.
// Narrow fronts: one lane per output pair, passes of 64.  Wide fronts: MF_Q lanes per pair, each
// summing a quarter of every block, so that the dependent chain of a job stays short.  The kernel
// is bound by the latency of dependent instructions and loads per job, not by bandwidth.  Sums run
// in ascending order with fused multiply-adds, slice sums are added in slice order: deterministic,
// bit-identical to mf_host_solve.
// HBM traffic per solve = both copies of the factor once (2 x 8 B x nnz(M)) + O(rows) vectors:
// the roofline model of SURVEY.md section 8(d) for the solve.
#include "mmpgo_mf.cuh"

namespace mmpgo {

namespace {

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void red_release_u32(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// lane 0 waits until `*ctr` has reached `need` (the jobs this one depends on have published their results); the
// other lanes follow through the warp barrier.  A wait of seconds means a broken dependency list: trap rather than hang.
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_count(const unsigned *ctr, unsigned need) {
  // back off while waiting: thousands of warps polling the L2 would slow down the ones that work.  The polls are
  // relaxed loads (an acquire load invalidates the SM's L1 at every poll: 6.3 M CCTL.IVALL per solve in the ncu
  // capture of the first version); one acquire load follows when the count is there (job_wait).
  // A wait of seconds means a broken dependency list: trap rather than hang (wall clock, so that a profiler's replay
  // or a time-sliced GPU does not trip it).
  unsigned ns = 32, polls = 0;
  unsigned long long t0 = 0;
  while (ld_relaxed_u32(ctr) < need) {
    __nanosleep(ns);
    if (ns < 128u) ns *= 2;          // measured caps 128 / 256 / 512 / 1024 ns: 0.574 / 0.573 / 0.577 / 0.583 ms per solve (8-node shard 0.295 / 0.296 / 0.300 / 0.313)
    if ((++polls & 4095u) == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 8000000000ull) __trap();
    }
  }
}
// L2 prefetch of a contiguous range (cp.async.bulk.prefetch.L2): no registers, no shared memory
__device__ __forceinline__ void l2_prefetch(const void *p, size_t bytes) {
  const size_t a0 = reinterpret_cast<size_t>(p) & ~size_t(15);
  const size_t a1 = (reinterpret_cast<size_t>(p) + bytes + 15) & ~size_t(15);
  for (size_t a = a0; a < a1; a += 16384) {
    const unsigned n = (unsigned)min((size_t)16384, a1 - a);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(n) : "memory");
  }
}
// the same for `count` pieces of `bytes` bytes, `ld` doubles apart (a row range of a column-major block): one line per
// prefetch instruction, dealt over the lanes of the warp
__device__ __forceinline__ void l2_prefetch_strided(const double *base, size_t ld, int count, int bytes, int lane) {
  const int lines = (bytes + 127) / 128;
  for (int i = lane; i < count * lines; i += 32) {
    const char *q = reinterpret_cast<const char *>(base + (size_t)(i / lines) * ld) + (i % lines) * 128;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
  }
}
// What a job waits for before it may read what other jobs wrote (nothing in level-barrier mode).  Lane 0 polls,
// its acquire loads order the reads of the whole warp (through the warp barrier) after the counters.
struct JobWait { const unsigned *c0, *c1; unsigned n0, n1; };
__device__ __forceinline__ void job_wait(const JobWait &w, int lane) {
  if (w.c0 == nullptr && w.c1 == nullptr) return;
  if (lane == 0) {
    if (w.c0) wait_count(w.c0, w.n0);
    if (w.c1) wait_count(w.c1, w.n1);
    // the synchronising reads: one acquire load per counter once it is there (a fence.acq_rel here showed up as 11 %
    // stall_membar: it also waits for the loads this job has in flight)
    if (w.c0) (void)ld_acquire_u32(w.c0);
    if (w.c1) (void)ld_acquire_u32(w.c1);
  }
  __syncwarp();
}

// all threads of all CTAs; `target` counts arrivals since the launch (the counter is zeroed before it)
__device__ __forceinline__ void grid_barrier(unsigned *ctr, unsigned &target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    while (ld_acquire_u32(ctr) < target) {}
    __threadfence();
  }
  __syncthreads();
}

// right-hand side of front row i: b (columns only; rhs arrives in the permuted numbering) + the children's
// update rows
template <int D>
__device__ __forceinline__ void front_rhs(const MfSolveArgs &a, const MfSn &sn, int i, double (&v)[D]) {
  const MfDevice &f = a.f;
  int p0 = -1, p1 = -1;
  if (sn.nchild) { p0 = __ldg(f.pull0 + sn.rowoff + i); p1 = __ldg(f.pull1 + sn.rowoff + i); }
  if (i < sn.k) {
    const double *b = a.rhs + (size_t)(sn.c0 + i) * D;
#pragma unroll
    for (int c = 0; c < D; ++c) v[c] = b[c];
  } else {
#pragma unroll
    for (int c = 0; c < D; ++c) v[c] = 0.0;
  }
  if (p0 >= 0) {
#pragma unroll
    for (int c = 0; c < D; ++c) v[c] += __ldcg(f.u + (size_t)p0 * D + c);
  }
  if (p1 >= 0) {
#pragma unroll
    for (int c = 0; c < D; ++c) v[c] += __ldcg(f.u + (size_t)p1 * D + c);
  }
}

// The factor is read exactly once per sweep: latency, not reuse, has to be covered.  A lane owns TWO adjacent
// output rows (forward, column-major copy) or columns (backward, row-major copy), so that one 128-bit load per
// term serves both and the 32 lanes of a warp read 512 contiguous bytes; MF_G terms are in flight per lane.  The
// front's right-hand side comes from shared memory (broadcast reads).
constexpr int MF_G = 16;    // measured per solve on the 1 M-pose grid: 4 in flight 1.31, 8: 0.97, 12: 0.87, 16: 0.84 ms (128 registers)
__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }

// a0[c] += sum_{j in [b, e)} P[j * ld].x * fv[j * D + c], a1 likewise with .y; ascending j, fused multiply-adds.
// The loop runs to the warp-uniform bound emax >= e; terms beyond e are taken as zeros (the buffer behind fv only
// ever holds finite numbers, up to 8 rows past any range).
template <int D>
__device__ __forceinline__ void dot_pair(const double *__restrict__ P, size_t ld, const double *fv, int b, int e, int emax,
                                         double (&a0)[D], double (&a1)[D]) {
  for (int j = b; j < emax; j += MF_G) {
    double2 mv[MF_G];
#pragma unroll
    for (int t = 0; t < MF_G; ++t) mv[t] = j + t < e ? ldg2(P + (size_t)(j + t) * ld) : make_double2(0.0, 0.0);
#pragma unroll
    for (int t = 0; t < MF_G; ++t) {
      const double *fp = fv + (j + t) * D;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double fc = fp[c];
        a0[c] = fma(mv[t].x, fc, a0[c]);
        a1[c] = fma(mv[t].y, fc, a1[c]);
      }
    }
  }
}
// slice sums of the MF_Q lanes of an output pair, added in slice order; valid in the lane with q == 0
template <int D> __device__ __forceinline__ void slice_sum(double (&acc)[D], int lane) {
#pragma unroll
  for (int c = 0; c < D; ++c) {
    double s = acc[c];
#pragma unroll
    for (int q = 1; q < MF_Q; ++q) s += __shfl_sync(0xffffffffu, acc[c], (lane & ~(MF_Q - 1)) + q);
    acc[c] = s;
  }
}

// rows [j0, j0 + cnt) of the front's forward right-hand side -> buf[0 .. cnt), zero beyond k.  cnt <= MF_BLK: the
// (at most four) rows of a lane are fetched together -- first every index, then every value -- instead of one
// dependent chain after the other.  Two halves: `pre` fetches what does not depend on other jobs (pull indices,
// the permuted right-hand side) and may run BEFORE the job waits for its children; `post` adds the children's
// update rows and fills the buffer.
template <int D> struct FwdStage { int p0[MF_BLK / 32], p1[MF_BLK / 32]; double v[MF_BLK / 32][D]; };
template <int D>
__device__ __forceinline__ void stage_forward_pre(const MfSolveArgs &a, const MfSn &sn, int j0, int cnt, int lane, FwdStage<D> &st) {
  const MfDevice &f = a.f;
  constexpr int U = MF_BLK / 32;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int j = j0 + lane + 32 * u;
    st.p0[u] = -1; st.p1[u] = -1;
    if (sn.nchild && lane + 32 * u < cnt && j < sn.k) { st.p0[u] = __ldg(f.pull0 + sn.rowoff + j); st.p1[u] = __ldg(f.pull1 + sn.rowoff + j); }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int j = j0 + lane + 32 * u;
    const bool on = lane + 32 * u < cnt && j < sn.k;
#pragma unroll
    for (int c = 0; c < D; ++c) st.v[u][c] = on ? a.rhs[(size_t)(sn.c0 + j) * D + c] : 0.0;
  }
}
template <int D>
__device__ __forceinline__ void stage_forward_post(const MfSolveArgs &a, int cnt, double *buf, int lane, FwdStage<D> &st) {
  const MfDevice &f = a.f;
  constexpr int U = MF_BLK / 32;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (st.p0[u] >= 0) {
#pragma unroll
      for (int c = 0; c < D; ++c) st.v[u][c] += __ldcg(f.u + (size_t)st.p0[u] * D + c);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (st.p1[u] >= 0) {
#pragma unroll
      for (int c = 0; c < D; ++c) st.v[u][c] += __ldcg(f.u + (size_t)st.p1[u] * D + c);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (lane + 32 * u < cnt) {
#pragma unroll
      for (int c = 0; c < D; ++c) buf[(lane + 32 * u) * D + c] = st.v[u][c];
    }
  }
}
template <int D>
__device__ __forceinline__ void stage_forward(const MfSolveArgs &a, const MfSn &sn, int j0, int cnt, double *buf, int lane) {
  FwdStage<D> st;
  stage_forward_pre<D>(a, sn, j0, cnt, lane, st);
  stage_forward_post<D>(a, cnt, buf, lane, st);
}
// rows [i0, i0 + cnt) of [y_s; -x_boundary] -> buf[0 .. cnt), zero beyond R; cnt <= MF_BLK, fetched like stage_forward
template <int D> struct BwdStage { const double *src[MF_BLK / 32]; };
// `pre`: where every row comes from (the boundary rows through the index list: static data, may be fetched before the
// job waits for its parent); `post`: the values (own y, the ancestors' x)
template <int D>
__device__ __forceinline__ void stage_backward_pre(const MfSolveArgs &a, const MfSn &sn, int i0, int cnt, int lane, BwdStage<D> &st) {
  const MfDevice &f = a.f;
  constexpr int U = MF_BLK / 32;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int i = i0 + lane + 32 * u;
    st.src[u] = nullptr;
    if (lane + 32 * u < cnt && i < sn.R)
      st.src[u] = i < sn.k ? f.y + (size_t)(sn.c0 + i) * D : f.xp + (size_t)__ldg(f.bidx + sn.boff + i - sn.k) * D;
  }
}
template <int D>
__device__ __forceinline__ void stage_backward_post(const MfSn &sn, int i0, int cnt, double *buf, int lane, const BwdStage<D> &st);
template <int D>
__device__ __forceinline__ void stage_backward(const MfSolveArgs &a, const MfSn &sn, int i0, int cnt, double *buf, int lane) {
  BwdStage<D> st;
  stage_backward_pre<D>(a, sn, i0, cnt, lane, st);
  stage_backward_post<D>(sn, i0, cnt, buf, lane, st);
}
template <int D>
__device__ __forceinline__ void stage_backward_post(const MfSn &sn, int i0, int cnt, double *buf, int lane, const BwdStage<D> &st) {
  constexpr int U = MF_BLK / 32;
  const double *const *src = st.src;
  double v[U][D];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const double sg = i0 + lane + 32 * u < sn.k ? 1.0 : -1.0;
#pragma unroll
    for (int c = 0; c < D; ++c) v[u][c] = src[u] ? sg * __ldcg(src[u] + c) : 0.0;
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (lane + 32 * u < cnt) {
#pragma unroll
      for (int c = 0; c < D; ++c) buf[(lane + 32 * u) * D + c] = v[u][c];
    }
  }
}

// forward: one output row of [y_s; -du] = M_s f1; f2 = the children's update rows of a boundary row (fetched
// before the dot product, see forward_job)
template <int D>
__device__ __forceinline__ void forward_store(const MfSolveArgs &a, const MfSn &sn, int i, const double (&acc)[D], const double (&f2)[D]) {
  const MfDevice &f = a.f;
  if (i < sn.k) {
#pragma unroll
    for (int c = 0; c < D; ++c) f.y[(size_t)(sn.c0 + i) * D + c] = acc[c];
  } else {
#pragma unroll
    for (int c = 0; c < D; ++c) f.u[(size_t)(sn.uoff + i - sn.k) * D + c] = f2[c] - acc[c];
  }
}
// update rows pulled by the boundary rows i, i + 1 of a front (zero for columns and for rows beyond `end`), in two
// halves: the children's row numbers (static: may be fetched before the job waits for the children) and the values
struct PullIdx { int a0, a1, b0, b1; };
__device__ __forceinline__ void pull_idx(const MfSolveArgs &a, const MfSn &sn, int i, int end, bool on, PullIdx &pi) {
  const MfDevice &f = a.f;
  pi.a0 = pi.a1 = pi.b0 = pi.b1 = -1;
  if (on && sn.nchild) {
    if (i >= sn.k && i < end) { pi.a0 = __ldg(f.pull0 + sn.rowoff + i); pi.a1 = __ldg(f.pull1 + sn.rowoff + i); }
    if (i + 1 >= sn.k && i + 1 < end) { pi.b0 = __ldg(f.pull0 + sn.rowoff + i + 1); pi.b1 = __ldg(f.pull1 + sn.rowoff + i + 1); }
  }
}
template <int D>
__device__ __forceinline__ void pull_vals(const MfSolveArgs &a, const PullIdx &pi, double (&g0)[D], double (&g1)[D]) {
  const MfDevice &f = a.f;
#pragma unroll
  for (int c = 0; c < D; ++c) { g0[c] = 0.0; g1[c] = 0.0; }
  // same order of additions as front_rhs: first child, then second
  if (pi.a0 >= 0) {
#pragma unroll
    for (int c = 0; c < D; ++c) g0[c] += __ldcg(f.u + (size_t)pi.a0 * D + c);
  }
  if (pi.a1 >= 0) {
#pragma unroll
    for (int c = 0; c < D; ++c) g0[c] += __ldcg(f.u + (size_t)pi.a1 * D + c);
  }
  if (pi.b0 >= 0) {
#pragma unroll
    for (int c = 0; c < D; ++c) g1[c] += __ldcg(f.u + (size_t)pi.b0 * D + c);
  }
  if (pi.b1 >= 0) {
#pragma unroll
    for (int c = 0; c < D; ++c) g1[c] += __ldcg(f.u + (size_t)pi.b1 * D + c);
  }
}
template <int D>
__device__ __forceinline__ void pull_pair(const MfSolveArgs &a, const MfSn &sn, int i, int end, bool on, double (&g0)[D], double (&g1)[D]) {
  PullIdx pi;
  pull_idx(a, sn, i, end, on, pi);
  pull_vals<D>(a, pi, g0, g1);
}
// backward: one entry of x_s
template <int D>
__device__ __forceinline__ void backward_store(const MfSolveArgs &a, const MfSn &sn, int j, const double (&acc)[D]) {
  const MfDevice &f = a.f;
  double *o = a.out + (size_t)__ldg(f.iperm + sn.c0 + j) * a.out_stride;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    f.xp[(size_t)(sn.c0 + j) * D + c] = acc[c];
    o[c] = a.sign * acc[c];
  }
}

// forward job: [y_s; -du] = M_s f1 for rows [r0, r0 + n)
template <int D>
__device__ __forceinline__ void forward_job(const MfSolveArgs &a, const MfSn &sn, int r0, int n, double *buf, int lane,
                                            const JobWait &w) {
  const MfDevice &f = a.f;
  const int k = sn.k, kp = (k + 1) & ~1, Rp = (sn.R + 1) & ~1;
  const double *Mc = f.M + sn.moff;                       // column-major, leading dimension Rp
  double a0[D], a1[D];
  if (k <= MF_KS) {
    // one lane per pair of rows, passes of 64 rows; the whole right-hand side fits the buffer
    double g0[D], g1[D];
    {
      // indices and b (and the children's row numbers of the first pass) are in flight while the job waits for its
      // children; what the children wrote is requested in one go right after the wait
      FwdStage<D> st;
      PullIdx pi;
      stage_forward_pre<D>(a, sn, 0, kp, lane, st);
      pull_idx(a, sn, r0 + 2 * lane, r0 + n, r0 + 2 * lane < r0 + n, pi);
      // a job that has to wait asks for its (static) block of the factor meanwhile
      if (w.c0 || w.c1) {
        if (r0 == 0 && n == sn.R) { if (lane == 0) l2_prefetch(Mc, (size_t)Rp * k * sizeof(double)); }
        else l2_prefetch_strided(Mc + r0, (size_t)Rp, r0 < k ? min(k, r0 + 64) : k, n * (int)sizeof(double), lane);
      }
      job_wait(w, lane);
      pull_vals<D>(a, pi, g0, g1);
      stage_forward_post<D>(a, kp, buf, lane, st);
    }
    __syncwarp();
    for (int p0 = r0; p0 < r0 + n; p0 += 64) {
      const int i = p0 + 2 * lane;
      const bool mine = i < r0 + n;
      // inv(L11) is lower triangular: rows below p0 + 64 never reach beyond column p0 + 63
      const int emax = p0 < k ? min(k, p0 + 64) : k;
#pragma unroll
      for (int c = 0; c < D; ++c) { a0[c] = 0.0; a1[c] = 0.0; }
      if (p0 != r0) pull_pair<D>(a, sn, i, r0 + n, mine, g0, g1);
      dot_pair<D>(Mc + i, (size_t)Rp, buf, 0, mine ? emax : 0, emax, a0, a1);
      if (mine) {
        forward_store<D>(a, sn, i, a0, g0);
        if (i + 1 < r0 + n) forward_store<D>(a, sn, i + 1, a1, g1);
      }
    }
  } else {
    // MF_Q lanes per pair of rows (n <= 2 MF_SROWS rows), the right-hand side staged in blocks of MF_BLK columns;
    // lane q of a pair sums the q-th quarter of every block
    const int ri = lane / MF_Q, q = lane % MF_Q;
    const int i = r0 + 2 * ri;
    const bool mine = 2 * ri < n;
    const int emax = r0 < k ? min(k, r0 + 2 * MF_SROWS) : k;
#pragma unroll
    for (int c = 0; c < D; ++c) { a0[c] = 0.0; a1[c] = 0.0; }
    double g0[D], g1[D];
    {
      FwdStage<D> st;
      PullIdx pi;
      stage_forward_pre<D>(a, sn, 0, min(MF_BLK, kp), lane, st);
      pull_idx(a, sn, i, r0 + n, mine && q == 0, pi);
      if (w.c0 || w.c1) l2_prefetch_strided(Mc + r0, (size_t)Rp, emax, n * (int)sizeof(double), lane);
      job_wait(w, lane);
      pull_vals<D>(a, pi, g0, g1);
      stage_forward_post<D>(a, min(MF_BLK, kp), buf, lane, st);
      __syncwarp();
    }
    for (int cb = 0; cb < emax; cb += MF_BLK) {
      if (cb > 0) {
        __syncwarp();
        stage_forward<D>(a, sn, cb, min(MF_BLK, kp - cb), buf, lane);
        __syncwarp();
      }
      const int b = cb + q * MF_QW, em = min(emax, b + MF_QW);
      dot_pair<D>(Mc + i, (size_t)Rp, buf - cb * D, b, mine ? em : b, em, a0, a1);
    }
    slice_sum<D>(a0, lane);
    slice_sum<D>(a1, lane);
    if (mine && q == 0) {
      forward_store<D>(a, sn, i, a0, g0);
      if (2 * ri + 1 < n) forward_store<D>(a, sn, i + 1, a1, g1);
    }
  }
  __syncwarp();
}

// backward job: x_s = M_s^T [y_s; -x_boundary] for columns [c0, c0 + n)
template <int D>
__device__ __forceinline__ void backward_job(const MfSolveArgs &a, const MfSn &sn, int c0, int n, double *buf, int lane,
                                             const JobWait &w) {
  const MfDevice &f = a.f;
  const int R = sn.R, Rp = (R + 1) & ~1, kp = (sn.k + 1) & ~1;
  const double *Mr = f.MT + sn.mtoff;                     // row-major, leading dimension kp
  double a0[D], a1[D];
  if (R <= MF_RS) {
    {
      BwdStage<D> st;
      stage_backward_pre<D>(a, sn, 0, Rp, lane, st);       // the boundary index list: in flight while the job waits for its parent
      if (lane == 0 && (w.c0 || w.c1) && c0 == 0 && n == sn.k) l2_prefetch(Mr, (size_t)kp * R * sizeof(double));
      job_wait(w, lane);
      stage_backward_post<D>(sn, 0, Rp, buf, lane, st);
    }
    __syncwarp();
    for (int p0 = c0; p0 < c0 + n; p0 += 64) {            // rows before p0 multiply zeros of every column of the pass
      const int j = p0 + 2 * lane;
      const bool mine = j < c0 + n;
#pragma unroll
      for (int c = 0; c < D; ++c) { a0[c] = 0.0; a1[c] = 0.0; }
      dot_pair<D>(Mr + j, (size_t)kp, buf, p0, mine ? R : p0, R, a0, a1);
      if (mine) {
        backward_store<D>(a, sn, j, a0);
        if (j + 1 < c0 + n) backward_store<D>(a, sn, j + 1, a1);
      }
    }
  } else {
    const int lj = lane / MF_Q, q = lane % MF_Q;
    const int j = c0 + 2 * lj;
    const bool mine = 2 * lj < n;
#pragma unroll
    for (int c = 0; c < D; ++c) { a0[c] = 0.0; a1[c] = 0.0; }
    for (int rb = c0; rb < R; rb += MF_BLK) {             // rows before c0 multiply zeros of the job's columns
      __syncwarp();
      if (rb == c0) {
        BwdStage<D> st;
        stage_backward_pre<D>(a, sn, rb, min(MF_BLK, Rp - rb), lane, st);
        if (w.c0 || w.c1) l2_prefetch_strided(Mr + (size_t)c0 * kp + c0, (size_t)kp, R - c0, n * (int)sizeof(double), lane);
        job_wait(w, lane);
        stage_backward_post<D>(sn, rb, min(MF_BLK, Rp - rb), buf, lane, st);
      } else {
        stage_backward<D>(a, sn, rb, min(MF_BLK, Rp - rb), buf, lane);
      }
      __syncwarp();
      const int b = rb + q * MF_QW, em = min(R, b + MF_QW);
      dot_pair<D>(Mr + j, (size_t)kp, buf - rb * D, b, mine ? em : b, em, a0, a1);
    }
    slice_sum<D>(a0, lane);
    slice_sum<D>(a1, lane);
    if (mine && q == 0) {
      backward_store<D>(a, sn, j, a0);
      if (2 * lj + 1 < n) backward_store<D>(a, sn, j + 1, a1);
    }
  }
  __syncwarp();
}

template <int D>
__global__ void __launch_bounds__(MF_THREADS, 1) k_mf_solve(MfSolveArgs a) {
  extern __shared__ double mf_sm[];
  const MfDevice &f = a.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *wbuf = mf_sm + warp * (MF_BUF * D);
  const int gw = blockIdx.x * MF_WARPS + warp, nw = gridDim.x * MF_WARPS;
  // the dot loops run to padded bounds and multiply loaded zeros with whatever the buffer holds there: make sure
  // it only ever holds finite numbers
  for (int i = threadIdx.x; i < (MF_WARPS + 1) * MF_BUF * D; i += MF_THREADS) mf_sm[i] = 0.0;
  __syncthreads();
  unsigned target = 0;
  int stamp = 0;
  auto mark = [&]() {      // stage boundaries as seen by CTA 0 (mmpgo_solver_stage_times)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      f.stage_ns[stamp] = t;
    }
    ++stamp;
  };
  mark();
#pragma unroll 1
  for (int dir = 0; dir < 2; ++dir) {
    const MfJob *wj = f.wjobs[dir];
    for (int st = 0; st < f.n_stage[dir]; ++st) {
      // every warp strides over the job list (ordered by level), the next job record in flight while the current
      // one runs.  Without level barriers (the default) the loop covers the whole sweep at once and a job waits only
      // for the supernodes it depends on: levels overlap, and so do the robot nodes.
      const int w0 = a.level_sync ? f.wstage[dir][st] : f.wstage[dir][0];
      const int w1 = a.level_sync ? f.wstage[dir][st + 1] : f.wstage[dir][f.n_stage[dir]];
      unsigned *done = f.done + (size_t)dir * f.n_sn;
      int t = w0 + gw;
      MfJob cur = {0, 0, 0, -1, -1, 0, 0, 0};
      if (t < w1) cur = wj[t];
      while (t < w1) {
        const int tn = t + nw;
        MfJob nxt = {0, 0, 0, -1, -1, 0, 0, 0};
        if (tn < w1) nxt = wj[tn];
        const MfSn sn = f.sn[cur.sn];
        if (a.dry || (a.active && !a.active[sn.node])) { t = tn; cur = nxt; continue; }
        JobWait w = {nullptr, nullptr, 0u, 0u};
        if (!a.level_sync) {
          if (cur.wait0 >= 0) { w.c0 = f.done + cur.wait0; w.n0 = (unsigned)cur.need0; }
          if (cur.wait1 >= 0) { w.c1 = f.done + cur.wait1; w.n1 = (unsigned)cur.need1; }
        }
        // the job waits for the supernodes it depends on INSIDE, after it has requested what does not depend on them
        if (dir == 0) forward_job<D>(a, sn, cur.r0, cur.n, wbuf, lane, w); else backward_job<D>(a, sn, cur.r0, cur.n, wbuf, lane, w);
        if (!a.level_sync) {
          __syncwarp();
          if (lane == 0) red_release_u32(done + cur.sn, 1u);
        }
        t = tn; cur = nxt;
      }
      if (!a.level_sync) { mark(); break; }
      if (dir == 0 || st + 1 < f.n_stage[1]) grid_barrier(f.barrier, target);
      mark();
    }
  }
}

}  // namespace

template <int D> int launch_mf_solve(const MfSolveArgs &a, int grid, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(a.f.barrier, 0, sizeof(unsigned), s);
  if (e != cudaSuccess) return (int)e;
  if (!a.level_sync && (e = cudaMemsetAsync(a.f.done, 0, sizeof(unsigned) * 2 * (size_t)a.f.n_sn, s)) != cudaSuccess) return (int)e;
  if (a.f.smem_bytes > 48 * 1024) {
    e = cudaFuncSetAttribute(k_mf_solve<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, a.f.smem_bytes);
    if (e != cudaSuccess) return (int)e;
  }
  MfSolveArgs args = a;
  void *kargs[] = {&args};
  return (int)cudaLaunchCooperativeKernel((const void *)k_mf_solve<D>, dim3(grid), dim3(MF_THREADS), kargs,
                                          (size_t)a.f.smem_bytes, s);
}
template int launch_mf_solve<2>(const MfSolveArgs &, int, cudaStream_t);
template int launch_mf_solve<3>(const MfSolveArgs &, int, cudaStream_t);

template <int D> int mf_solve_max_grid(int device, int smem_bytes) {
  if (smem_bytes > 48 * 1024 &&
      cudaFuncSetAttribute(k_mf_solve<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
    return -1;
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mf_solve<D>, MF_THREADS, (size_t)smem_bytes) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  return per_sm * sms;
}
template int mf_solve_max_grid<2>(int, int);
template int mf_solve_max_grid<3>(int, int);

}  // namespace mmpgo
