// K2b, direct path: the level-scheduled supernodal sweeps of the sparse Cholesky solve, ONE
// persistent cooperative launch per solve (data model and algebra: mmpgo_mf.cuh).
//
// Replaces `L_.solve(...)` of the reference (CHOLMOD through Eigen; call sites
// C++/DPGO/include/DPGO/DPGOProblem.h:291, C++/DPGO/src/DPGOProblem.cpp:140, :568).
//
// Execution model.  A stage = all supernodes of equal height (forward) or depth (backward) of
// every active robot node; its tasks are dealt round-robin to the persistent CTAs, stages are
// separated by a grid barrier (2 H + 1 barriers per solve, H = tree height, 14 on the 15 625-pose
// slabs of the 1 M-pose grid), one 1024-thread CTA per SM.  Supernodes with at most MF_RW rows are
// served by single warps (32 independent warp jobs per task: 32 rows forward / 32 columns backward
// each, the front's right-hand side staged in the warp's slice of shared memory); larger ones by
// whole CTAs, MF_SPAN rows / columns per job, each summed in MF_Q contiguous slices by MF_Q threads.
// Sums run in ascending order with fused multiply-adds, slice sums are added in slice order:
// deterministic, bit-identical to mf_host_solve.  The factor is read exactly once per sweep, so the
// loops batch their loads (MF_BATCH in flight per thread) instead of counting on reuse.
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

// The factor is read exactly once per sweep: latency, not reuse, has to be covered, so the loads of MF_BATCH
// consecutive terms are issued together and the first batch is issued BEFORE the front's right-hand side is
// staged (its addresses depend on the supernode record only).
constexpr int MF_BATCH = 8;
__device__ __forceinline__ void load_batch(const double *__restrict__ Mp, size_t stride, int j, int e, double (&mv)[MF_BATCH]) {
#pragma unroll
  for (int t = 0; t < MF_BATCH; ++t) mv[t] = j + t < e ? __ldg(Mp + (size_t)(j + t) * stride) : 0.0;
}
// acc[c] = sum_{j in [b, e)} Mp[j * stride] * fv[j * D + c], ascending j, fused multiply-adds; mv = the batch
// loaded at j = b
template <int D>
__device__ __forceinline__ void dot_range(const double *__restrict__ Mp, size_t stride, const double *fv, int b, int e,
                                          double (&mv)[MF_BATCH], double (&acc)[D]) {
#pragma unroll
  for (int c = 0; c < D; ++c) acc[c] = 0.0;
  for (int j = b; j < e; j += MF_BATCH) {
    if (j > b) load_batch(Mp, stride, j, e, mv);
#pragma unroll
    for (int t = 0; t < MF_BATCH; ++t) {
      if (j + t < e) {
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = fma(mv[t], fv[(j + t) * D + c], acc[c]);
      }
    }
  }
}

// ---- warp jobs (fronts of at most MF_RW rows) ------------------------------------------------
// forward: rows [r0, r0 + 32) of the supernode; `buf` holds k x D doubles
template <int D>
__device__ __forceinline__ void forward_warp(const MfSolveArgs &a, const MfSn &sn, int r0, double *buf, int lane) {
  const MfDevice &f = a.f;
  if (a.active && !a.active[sn.node]) return;
  const int k = sn.k, R = sn.R;
  const int i = r0 + lane;
  // rows above the diagonal of inv(L11) are zero: columns beyond the chunk's last row add nothing
  const int jend = i < R ? (r0 < k ? min(k, r0 + 32) : k) : 0;
  const double *Mi = f.M + sn.moff + i;
  double mv[MF_BATCH];
  load_batch(Mi, (size_t)R, 0, jend, mv);
  for (int j = lane; j < k; j += 32) {
    double v[D];
    front_rhs<D>(a, sn, j, v);
#pragma unroll
    for (int c = 0; c < D; ++c) buf[j * D + c] = v[c];
  }
  double f2[D];
  if (i >= k && i < R) front_rhs<D>(a, sn, i, f2);
  __syncwarp();
  if (i < R) {
    double acc[D];
    dot_range<D>(Mi, (size_t)R, buf, 0, jend, mv, acc);
    if (i < k) {
#pragma unroll
      for (int c = 0; c < D; ++c) f.y[(size_t)(sn.c0 + i) * D + c] = acc[c];
    } else {
#pragma unroll
      for (int c = 0; c < D; ++c) f.u[(size_t)(sn.uoff + i - k) * D + c] = f2[c] - acc[c];
    }
  }
  __syncwarp();
}

// backward: columns [r0, r0 + 32); `buf` holds R x D doubles
template <int D>
__device__ __forceinline__ void backward_warp(const MfSolveArgs &a, const MfSn &sn, int r0, double *buf, int lane) {
  const MfDevice &f = a.f;
  if (a.active && !a.active[sn.node]) return;
  const int k = sn.k, R = sn.R;
  const int j = r0 + lane;
  const int e = j < k ? R : r0;
  const double *Mj = f.MT + sn.moff + j;
  double mv[MF_BATCH];
  load_batch(Mj, (size_t)k, r0, e, mv);
  const int orow = j < k ? __ldg(f.iperm + sn.c0 + j) : 0;
  for (int i = r0 + lane; i < R; i += 32) {      // rows before r0 multiply zeros of every column of the chunk
    const double *src = i < k ? f.y + (size_t)(sn.c0 + i) * D : f.xp + (size_t)__ldg(f.bidx + sn.boff + i - k) * D;
    const double sg = i < k ? 1.0 : -1.0;
#pragma unroll
    for (int c = 0; c < D; ++c) buf[i * D + c] = sg * __ldcg(src + c);
  }
  __syncwarp();
  if (j < k) {
    double acc[D];
    dot_range<D>(Mj, (size_t)k, buf, r0, R, mv, acc);
    double *o = a.out + (size_t)orow * a.out_stride;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      f.xp[(size_t)(sn.c0 + j) * D + c] = acc[c];
      o[c] = a.sign * acc[c];
    }
  }
  __syncwarp();
}

// ---- CTA jobs (larger fronts): MF_SPAN rows / columns, each summed in MF_Q contiguous slices by MF_Q
// threads; the slice sums are added in slice order.  `buf`: front right-hand side, `part`: [MF_Q][MF_SPAN][D]
template <int D>
__device__ __forceinline__ void forward_cta(const MfSolveArgs &a, const MfJob job, double *buf, double *part) {
  const MfDevice &f = a.f;
  const MfSn sn = f.sn[job.sn];
  const bool on = !a.active || a.active[sn.node];
  const int k = sn.k, R = sn.R;
  const int li = threadIdx.x % MF_SPAN, q = threadIdx.x / MF_SPAN;
  const int i = job.r0 + li;
  const int jend = job.r0 < k ? min(k, job.r0 + MF_SPAN) : k;
  const int per = (jend + MF_Q - 1) / MF_Q;
  const int jb = q * per, je = (on && i < R) ? min(jend, (q + 1) * per) : jb;
  const double *Mi = f.M + sn.moff + i;
  double mv[MF_BATCH];
  load_batch(Mi, (size_t)R, jb, je, mv);
  if (on) {
    for (int j = threadIdx.x; j < k; j += MF_THREADS) {
      double v[D];
      front_rhs<D>(a, sn, j, v);
#pragma unroll
      for (int c = 0; c < D; ++c) buf[j * D + c] = v[c];
    }
  }
  double f2[D];
  if (on && q == 0 && i >= k && i < R) front_rhs<D>(a, sn, i, f2);
  __syncthreads();
  if (on && i < R) {
    double acc[D];
    dot_range<D>(Mi, (size_t)R, buf, jb, je, mv, acc);
#pragma unroll
    for (int c = 0; c < D; ++c) part[(q * MF_SPAN + li) * D + c] = acc[c];
  }
  __syncthreads();
  if (on && i < R && q == 0) {
    double acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
      acc[c] = part[li * D + c];
#pragma unroll
      for (int qq = 1; qq < MF_Q; ++qq) acc[c] += part[(qq * MF_SPAN + li) * D + c];
    }
    if (i < k) {
#pragma unroll
      for (int c = 0; c < D; ++c) f.y[(size_t)(sn.c0 + i) * D + c] = acc[c];
    } else {
#pragma unroll
      for (int c = 0; c < D; ++c) f.u[(size_t)(sn.uoff + i - k) * D + c] = f2[c] - acc[c];
    }
  }
}

template <int D>
__device__ __forceinline__ void backward_cta(const MfSolveArgs &a, const MfJob job, double *buf, double *part) {
  const MfDevice &f = a.f;
  const MfSn sn = f.sn[job.sn];
  const bool on = !a.active || a.active[sn.node];
  const int k = sn.k, R = sn.R;
  const int lj = threadIdx.x % MF_SPAN, q = threadIdx.x / MF_SPAN;
  const int j = job.r0 + lj;
  const int per = (R - job.r0 + MF_Q - 1) / MF_Q;
  const int ib = job.r0 + q * per, ie = (on && j < k) ? min(R, job.r0 + (q + 1) * per) : ib;
  const double *Mj = f.MT + sn.moff + j;
  double mv[MF_BATCH];
  load_batch(Mj, (size_t)k, ib, ie, mv);
  const int orow = (on && j < k && q == 0) ? __ldg(f.iperm + sn.c0 + j) : 0;
  if (on) {
    for (int i = job.r0 + threadIdx.x; i < R; i += MF_THREADS) {
      const double *src = i < k ? f.y + (size_t)(sn.c0 + i) * D : f.xp + (size_t)__ldg(f.bidx + sn.boff + i - k) * D;
      const double sg = i < k ? 1.0 : -1.0;
#pragma unroll
      for (int c = 0; c < D; ++c) buf[i * D + c] = sg * __ldcg(src + c);
    }
  }
  __syncthreads();
  if (on && j < k) {
    double acc[D];
    dot_range<D>(Mj, (size_t)k, buf, ib, ie, mv, acc);
#pragma unroll
    for (int c = 0; c < D; ++c) part[(q * MF_SPAN + lj) * D + c] = acc[c];
  }
  __syncthreads();
  if (on && j < k && q == 0) {
    double *o = a.out + (size_t)orow * a.out_stride;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      double acc = part[lj * D + c];
#pragma unroll
      for (int qq = 1; qq < MF_Q; ++qq) acc += part[(qq * MF_SPAN + lj) * D + c];
      f.xp[(size_t)(sn.c0 + j) * D + c] = acc;
      o[c] = a.sign * acc;
    }
  }
}

template <int D>
__global__ void __launch_bounds__(MF_THREADS, 1) k_mf_solve(MfSolveArgs a) {
  extern __shared__ double mf_sm[];
  const MfDevice &f = a.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *part = mf_sm + f.part_off, *wbuf = mf_sm + warp * (MF_RW * D);
  const int gw = blockIdx.x * MF_WARPS + warp, nw = gridDim.x * MF_WARPS;
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
    const MfJob *wj = f.wjobs[dir], *cj = f.cjobs[dir];
    for (int st = 0; st < f.n_stage[dir]; ++st) {
      // CTA jobs of the stage first (the long ones), then the warp jobs, every warp striding over the
      // stage's list with the NEXT job's record already in flight
      const int c1 = f.cstage[dir][st + 1];
      for (int t = f.cstage[dir][st] + blockIdx.x; t < c1; t += gridDim.x) {
        __syncthreads();                                 // the previous job's shared memory is free
        if (dir == 0) forward_cta<D>(a, cj[t], mf_sm, part); else backward_cta<D>(a, cj[t], mf_sm, part);
      }
      __syncthreads();
      const int w1 = f.wstage[dir][st + 1];
      int t = f.wstage[dir][st] + gw;
      MfJob job = {0, 0};
      MfSn sn;
      if (t < w1) { job = wj[t]; sn = f.sn[job.sn]; }
      while (t < w1) {
        const int tn = t + nw;
        MfJob jobn = {0, 0};
        MfSn snn;
        if (tn < w1) { jobn = wj[tn]; snn = f.sn[jobn.sn]; }
        if (dir == 0) forward_warp<D>(a, sn, job.r0, wbuf, lane); else backward_warp<D>(a, sn, job.r0, wbuf, lane);
        t = tn; job = jobn; sn = snn;
      }
      if (dir == 0 || st + 1 < f.n_stage[1]) grid_barrier(f.barrier, target);
      mark();
    }
  }
}

}  // namespace

template <int D> int launch_mf_solve(const MfSolveArgs &a, int grid, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(a.f.barrier, 0, sizeof(unsigned), s);
  if (e != cudaSuccess) return (int)e;
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
