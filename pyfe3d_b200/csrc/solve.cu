// Device-resident consumers of the assembled matrix:
//
//  * plan_cg        Jacobi-preconditioned conjugate gradient on the CSR values of structured plans, the whole
//                   iteration on the device: block SpMV (assembly.cu) + three fused vector kernels, deterministic
//                   reductions, convergence decided on the device (no host synchronisation inside an iteration).
//                   It is the diagonal-scaled cg of the reference scripts
//                       D = diag(Kuu)^-1/2;  cg(D Kuu D, D f)                 (tests/test_quad4r_linear_buckling_plate.py:135-146)
//                   written as preconditioned CG on the unscaled system (the same iterates: z = D^2 r), with the
//                   boundary-condition partition K[bu,:][:,bu] (tests/test_quad4_static_point_load.py:84-99) applied
//                   as a mask inside the SpMV.
//  * csr_compact_*  K[bu,:][:,bu] as an explicit CSR matrix (row / column compaction of a device CSR matrix), for
//                   callers that hand Kuu to scipy's spsolve / eigsh.
#include <algorithm>
#include <cstdint>

#include <cub/device/device_scan.cuh>

#include "common.cuh"

struct pf3_plan;

namespace pf3 {

int plan_spmv(const pf3_plan* pl, cudaStream_t st, const double* vals, const unsigned char* free_, const double* x,
              double* y, int64_t* launches);
int plan_diagonal(const pf3_plan* pl, cudaStream_t st, const double* vals, double* diag, int64_t* launches);
int64_t plan_nrows(const pf3_plan* pl);

#define PF3_CUDA(x)                        \
  do {                                     \
    cudaError_t _e = (x);                  \
    if (_e != cudaSuccess) return int(_e); \
  } while (0)

namespace {

constexpr int kCgThreads = 256;
constexpr int kCgMaxBlocks = 148 * 8;

// device scalars of one solve
struct CgScal {
  double rz, rzn, pap, rr, bb, tol2;   // r.z, new r.z, p.Ap, the norm the criterion uses (squared), |b|^2, threshold^2
  int done;                            // 0 running, 1 converged, 2 breakdown (p.Ap <= 0 or not finite), 3 maxiter
  int iters, maxiter;
  unsigned ticket[3];
  int scaled;                          // criterion in the Jacobi-scaled norm r.D^-1.r (what cg on D K D measures)
};

// block-wide sum in a fixed order (deterministic): warp butterflies, then warp 0 over the warp sums
__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.;
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;   // valid in thread 0
}

// Grid-wide deterministic reduction: every block leaves its partial sums, the block that takes the last ticket adds
// them in block order and runs `fin(totals)`.  NV values per block.
template <int NV, class Fin>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], double* partial, unsigned* ticket, Fin fin) {
  __shared__ double sh[32];
  __shared__ bool last;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double s = block_sum(v[i], sh);
    if (threadIdx.x == 0) partial[i * gridDim.x + blockIdx.x] = s;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double tot[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double a = 0.;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) a += partial[i * gridDim.x + b];
    tot[i] = block_sum(a, sh);
  }
  if (threadIdx.x == 0) {
    *ticket = 0u;
    fin(tot);
  }
}

// minv = free && d != 0 ? 1/d : 0 ; optionally d accumulated as d += c * t first
__global__ void k_cg_diag_acc(int64_t n, double* __restrict__ d, const double* __restrict__ t, double c, int first) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    d[i] = first ? c * t[i] : d[i] + c * t[i];
}
__global__ void k_axpy(int64_t n, double* __restrict__ y, const double* __restrict__ t, double c, int first) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    y[i] = first ? c * t[i] : y[i] + c * t[i];
}
__global__ void k_mul(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    out[i] = a[i] * b[i];
}

// start: minv from the diagonal, x masked (or zeroed), r = P b - (A x already in ap when use_x0), z = minv r, p = z,
// rz = r.z, bb = |P b|^2 (or b.D^-1.b in the scaled norm)
__global__ void __launch_bounds__(kCgThreads) k_cg_start(int64_t n, const unsigned char* __restrict__ free_,
                                                         const double* __restrict__ b, double* __restrict__ minv,
                                                         double* __restrict__ x, double* __restrict__ r,
                                                         double* __restrict__ p, const double* __restrict__ ax,
                                                         int use_x0, double rtol, double atol, int maxiter,
                                                         double* partial, CgScal* S) {
  double acc[3] = {0., 0., 0.};
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const bool fr = free_ == nullptr || free_[i] != 0;
    const double d = minv[i];
    const double mi = (fr && d != 0.) ? 1. / d : 0.;
    minv[i] = mi;
    const double bi = fr ? b[i] : 0.;
    double ri = bi;
    if (use_x0) {
      if (!fr) x[i] = 0.;
      ri -= ax[i];
    } else {
      x[i] = 0.;
    }
    if (!fr) ri = 0.;
    r[i] = ri;
    const double zi = mi * ri;
    p[i] = zi;
    acc[0] += ri * zi;
    acc[1] += S->scaled ? bi * mi * bi : bi * bi;
    acc[2] += S->scaled ? ri * zi : ri * ri;
  }
  grid_reduce<3>(acc, partial, &S->ticket[0], [&](const double* t) {
    S->rz = t[0];
    S->bb = t[1];
    S->rr = t[2];
    const double tol = fmax(rtol * sqrt(t[1]), atol);
    S->tol2 = tol * tol;
    S->iters = 0;
    S->maxiter = maxiter;
    S->done = (t[2] <= tol * tol || t[0] == 0.) ? 1 : 0;
  });
}

__global__ void __launch_bounds__(kCgThreads) k_cg_dot(int64_t n, const double* __restrict__ p,
                                                       const double* __restrict__ ap, double* partial, CgScal* S) {
  if (S->done) return;
  double acc[1] = {0.};
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    acc[0] += p[i] * ap[i];
  grid_reduce<1>(acc, partial, &S->ticket[0], [&](const double* t) {
    S->pap = t[0];
    if (!(t[0] > 0.) || !isfinite(t[0])) S->done = 2;
  });
}

// x += a p ; r -= a Ap ; rzn = r.minv.r ; rr = |r|^2 (or rzn in the scaled norm); decides convergence
__global__ void __launch_bounds__(kCgThreads) k_cg_update(int64_t n, const double* __restrict__ p,
                                                          const double* __restrict__ ap,
                                                          const double* __restrict__ minv, double* __restrict__ x,
                                                          double* __restrict__ r, double* partial, CgScal* S) {
  if (S->done) return;
  const double alpha = S->rz / S->pap;
  double acc[2] = {0., 0.};
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * ap[i];
    r[i] = ri;
    acc[0] += ri * minv[i] * ri;
    acc[1] += ri * ri;
  }
  grid_reduce<2>(acc, partial, &S->ticket[1], [&](const double* t) {
    S->rzn = t[0];
    S->rr = S->scaled ? t[0] : t[1];
    S->iters += 1;
    if (S->rr <= S->tol2 || t[0] == 0.) S->done = 1;
    else if (S->iters >= S->maxiter) S->done = 3;
  });
}

// p = minv r + (rzn / rz) p ; the last block then moves rzn into rz
__global__ void __launch_bounds__(kCgThreads) k_cg_dir(int64_t n, const double* __restrict__ r,
                                                       const double* __restrict__ minv, double* __restrict__ p,
                                                       CgScal* S) {
  if (S->done) return;
  const double beta = S->rzn / S->rz;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    p[i] = minv[i] * r[i] + beta * p[i];
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(&S->ticket[2], 1u) == gridDim.x - 1;
    if (last) {
      S->ticket[2] = 0u;
      S->rz = S->rzn;
    }
  }
}

// ---- row-sharded CG: the same three passes on a rank's own rows, scalars in a caller-owned array (all-reduced by the
// caller between the passes): sc[0] p.Ap, sc[1] new r.z, sc[2] r.r, sc[3] current r.z
__global__ void __launch_bounds__(kCgThreads) k_shard_dot(int64_t n, const double* __restrict__ p,
                                                          const double* __restrict__ ap, double* partial,
                                                          unsigned* ticket, double* sc) {
  double acc[1] = {0.};
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    acc[0] += p[i] * ap[i];
  grid_reduce<1>(acc, partial, ticket, [&](const double* t) { sc[0] = t[0]; });
}
__global__ void __launch_bounds__(kCgThreads) k_shard_update(int64_t n, const double* __restrict__ p,
                                                             const double* __restrict__ ap,
                                                             const double* __restrict__ minv, double* __restrict__ x,
                                                             double* __restrict__ r, double* partial, unsigned* ticket,
                                                             double* sc) {
  // p.Ap <= 0 (breakdown, or an exactly converged system iterated on inside a batch): freeze the iterate instead of
  // spreading NaN; the host sees p.Ap and the residual at the end of the batch
  const double alpha = sc[0] > 0. ? sc[3] / sc[0] : 0.;
  double acc[2] = {0., 0.};
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * ap[i];
    r[i] = ri;
    acc[0] += ri * minv[i] * ri;
    acc[1] += ri * ri;
  }
  grid_reduce<2>(acc, partial, ticket, [&](const double* t) {
    sc[1] = t[0];
    sc[2] = t[1];
  });
}
__global__ void __launch_bounds__(kCgThreads) k_shard_dir(int64_t n, const double* __restrict__ r,
                                                          const double* __restrict__ minv, double* __restrict__ p,
                                                          unsigned* ticket, double* sc) {
  const double beta = sc[3] != 0. ? sc[1] / sc[3] : 0.;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    p[i] = minv[i] * r[i] + beta * p[i];
  __syncthreads();   // every thread of the block has read sc[3] (through beta) before the ticket is taken
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {   // the last block: nobody reads the old sc[3] any more
      *ticket = 0u;
      sc[3] = sc[1];
    }
  }
}

unsigned cg_grid(int64_t n) {
  const int64_t want = (n + kCgThreads - 1) / kCgThreads;
  return unsigned(std::max<int64_t>(1, std::min<int64_t>(want, kCgMaxBlocks)));
}

}  // namespace

// work layout of the row-sharded kernels: partial sums (2 per block) | three tickets
size_t cg_shard_work_bytes() { return size_t(2) * kCgMaxBlocks * sizeof(double) + 64; }
namespace {
inline unsigned* shard_tickets(void* work) { return reinterpret_cast<unsigned*>(static_cast<double*>(work) + 2 * kCgMaxBlocks); }
}
int cg_shard_dot(cudaStream_t st, int64_t n, const double* p, const double* ap, double* sc, void* work, int64_t* launches) {
  k_shard_dot<<<cg_grid(n), kCgThreads, 0, st>>>(n, p, ap, static_cast<double*>(work), shard_tickets(work), sc);
  ++*launches;
  return int(cudaGetLastError());
}
int cg_shard_update(cudaStream_t st, int64_t n, const double* p, const double* ap, const double* minv, double* x, double* r,
                    double* sc, void* work, int64_t* launches) {
  k_shard_update<<<cg_grid(n), kCgThreads, 0, st>>>(n, p, ap, minv, x, r, static_cast<double*>(work),
                                                    shard_tickets(work) + 1, sc);
  ++*launches;
  return int(cudaGetLastError());
}
int cg_shard_dir(cudaStream_t st, int64_t n, const double* r, const double* minv, double* p, double* sc, void* work,
                 int64_t* launches) {
  k_shard_dir<<<cg_grid(n), kCgThreads, 0, st>>>(n, r, minv, p, shard_tickets(work) + 2, sc);
  ++*launches;
  return int(cudaGetLastError());
}

struct CgOp {
  const pf3_plan* plan;
  const double* vals;
  double coef;
};

size_t cg_work_bytes(int64_t n) {
  // r, p, ap, minv, tmp | partial sums (3 values per block) | scalars
  return size_t(5) * size_t(n) * sizeof(double) + size_t(3) * kCgMaxBlocks * sizeof(double) + 256;
}

// info_out: {iterations, status (0 converged, 1 maxiter reached, 2 breakdown), residual norm, |b|}
int plan_cg(cudaStream_t st, int nops, const CgOp* ops, int64_t n, const unsigned char* free_, const double* b,
            double* x, int use_x0, double rtol, double atol, int maxiter, int flags, void* work, int* iters,
            int* status, double* resid, double* bnorm, int64_t* launches) {
  if (nops < 1 || n <= 0) return PF3_E_BAD_ARG;
  for (int i = 0; i < nops; ++i)
    if (plan_nrows(ops[i].plan) != n) return PF3_E_BAD_ARG;   // single device: every plan owns every row
  double* r = static_cast<double*>(work);
  double* p = r + n;
  double* ap = p + n;
  double* minv = ap + n;
  double* tmp = minv + n;
  double* partial = tmp + n;
  CgScal* S = reinterpret_cast<CgScal*>(partial + 3 * kCgMaxBlocks);
  const unsigned grid = cg_grid(n);
  PF3_CUDA(cudaMemsetAsync(S, 0, sizeof(CgScal), st));
  if (flags & 1) {
    const int one = 1;
    PF3_CUDA(cudaMemcpyAsync(&S->scaled, &one, sizeof(int), cudaMemcpyHostToDevice, st));
  }
  auto matvec = [&](const double* xin, double* out) -> int {
    for (int i = 0; i < nops; ++i) {
      const bool direct = i == 0 && ops[i].coef == 1.;
      int rc = plan_spmv(ops[i].plan, st, ops[i].vals, free_, xin, direct ? out : tmp, launches);
      if (rc) return rc;
      if (!direct) {
        k_axpy<<<grid, kCgThreads, 0, st>>>(n, out, tmp, ops[i].coef, i == 0);
        ++*launches;
      }
    }
    return int(cudaGetLastError());
  };
  // diagonal of the operator -> minv (finished in k_cg_start)
  for (int i = 0; i < nops; ++i) {
    int rc = plan_diagonal(ops[i].plan, st, ops[i].vals, tmp, launches);
    if (rc) return rc;
    k_cg_diag_acc<<<grid, kCgThreads, 0, st>>>(n, minv, tmp, ops[i].coef, i == 0);
    ++*launches;
  }
  if (use_x0) {
    int rc = matvec(x, ap);   // the SpMV masks constrained columns and rows itself
    if (rc) return rc;
  }
  if (maxiter <= 0) maxiter = int(std::min<int64_t>(10 * n, 2000000000));
  k_cg_start<<<grid, kCgThreads, 0, st>>>(n, free_, b, minv, x, r, p, ap, use_x0, rtol, atol, maxiter, partial, S);
  ++*launches;
  PF3_CUDA(cudaGetLastError());
  const int every = (flags >> 8) > 0 ? (flags >> 8) : 16;   // iterations between two looks at the device flag
  CgScal h;
  PF3_CUDA(cudaMemcpyAsync(&h, S, sizeof(CgScal), cudaMemcpyDeviceToHost, st));
  PF3_CUDA(cudaStreamSynchronize(st));
  // one batch = `every` iterations; every kernel returns at once when the device flag is set (converged, breakdown
  // or maxiter), so a batch may be enqueued blindly.  With PF3_CG_GRAPH the batch is captured once into a CUDA graph
  // and replayed: small systems are launch-bound (4 launches of a few microseconds per iteration).
  int64_t per_batch = 0;
  auto enqueue = [&]() -> int {
    const int64_t l0 = *launches;
    for (int k = 0; k < every; ++k) {
      int rc = matvec(p, ap);
      if (rc) return rc;
      k_cg_dot<<<grid, kCgThreads, 0, st>>>(n, p, ap, partial, S);
      k_cg_update<<<grid, kCgThreads, 0, st>>>(n, p, ap, minv, x, r, partial, S);
      k_cg_dir<<<grid, kCgThreads, 0, st>>>(n, r, minv, p, S);
      *launches += 3;
    }
    per_batch = *launches - l0;
    return int(cudaGetLastError());
  };
  cudaGraphExec_t gexec = nullptr;
  if ((flags & 2) && !h.done) {
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      const int rc = enqueue();
      const cudaError_t ce = cudaStreamEndCapture(st, &g);
      if (rc == 0 && ce == cudaSuccess && g != nullptr && cudaGraphInstantiate(&gexec, g, 0) != cudaSuccess) gexec = nullptr;
      if (g) cudaGraphDestroy(g);
      *launches -= per_batch;   // captured, not yet run
    }
    cudaGetLastError();         // a refused capture (e.g. the legacy default stream) falls back to plain launches
  }
  int64_t guard = (int64_t(maxiter) + every - 1) / every + 1;
  while (!h.done && guard-- > 0) {
    if (gexec) {
      PF3_CUDA(cudaGraphLaunch(gexec, st));
      *launches += per_batch;
    } else {
      int rc = enqueue();
      if (rc) return rc;
    }
    PF3_CUDA(cudaMemcpyAsync(&h, S, sizeof(CgScal), cudaMemcpyDeviceToHost, st));
    PF3_CUDA(cudaStreamSynchronize(st));
  }
  if (gexec) cudaGraphExecDestroy(gexec);
  if (iters) *iters = h.iters;
  if (status) *status = h.done == 1 ? 0 : (h.done == 2 ? 2 : 1);   // 0 converged, 1 maxiter, 2 breakdown
  if (resid) *resid = sqrt(h.rr);
  if (bnorm) *bnorm = sqrt(h.bb);
  return PF3_OK;
}

// y = S P A P S x (S = diag(scale)): the scaled operators the reference scripts hand to eigsh
// (tests/test_quad4r_linear_buckling_plate.py:172-180).  tmp: n doubles.
int plan_spmv_scaled(const pf3_plan* pl, cudaStream_t st, const double* vals, const unsigned char* free_,
                     const double* scale, const double* x, double* y, double* tmp, int64_t n, int64_t* launches) {
  const unsigned grid = cg_grid(n);
  k_mul<<<grid, kCgThreads, 0, st>>>(n, scale, x, tmp);
  ++*launches;
  int rc = plan_spmv(pl, st, vals, free_, tmp, y, launches);
  if (rc) return rc;
  k_mul<<<grid, kCgThreads, 0, st>>>(plan_nrows(pl), scale, y, y);
  ++*launches;
  return int(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------
// K[bu, :][:, bu] as CSR (tests/test_quad4_static_point_load.py:84-99): rows with free_[row0 + i] != 0 are kept and
// renumbered in order; columns with free_[c] != 0 are kept and renumbered by their rank among the free columns
// (free_ == NULL: everything is kept).  upper: additionally only entries with col >= row (scipy.sparse.triu): KC0, KG
// and M are symmetric, so a host consumer that accepts a triangle (CHOLMOD, PARDISO, eigsh on a symmetric operator)
// needs 5/9 of the bytes moved over PCIe.
namespace {

__global__ void k_flag_i64(int64_t n, const unsigned char* __restrict__ f, int64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    out[i] = (f == nullptr || f[i] != 0) ? 1 : 0;
}
// keep entry (row, col)?  free_ == NULL: no constraints; upper: the upper triangle col >= row only (symmetric matrices:
// half the values to move)
__device__ __forceinline__ bool keep_entry(const unsigned char* free_, bool upper, int64_t grow, int64_t col) {
  return (free_ == nullptr || free_[col] != 0) && (!upper || col >= grow);
}

// one warp per source row: number of kept entries -> cnt[rank of the row] (kept rows only)
__global__ void k_compact_count(int64_t nrows, const int64_t* __restrict__ indptr, const int64_t* __restrict__ indices,
                                const unsigned char* __restrict__ free_, int upper, int64_t row0,
                                const int64_t* __restrict__ colmap, int64_t* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5; row < nrows;
       row += (int64_t(gridDim.x) * blockDim.x) >> 5) {
    if (free_ != nullptr && !free_[row0 + row]) continue;
    int64_t c = 0;
    for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32) c += keep_entry(free_, upper != 0, row0 + row, indices[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) cnt[colmap[row0 + row]] = c;
  }
}

// one warp per source row: kept entries written in order (ballot + prefix popcount keeps the column order)
__global__ void k_compact_fill(int64_t nrows, const int64_t* __restrict__ indptr, const int64_t* __restrict__ indices,
                               const double* __restrict__ vals, const unsigned char* __restrict__ free_, int upper,
                               int64_t row0, const int64_t* __restrict__ colmap, const int64_t* __restrict__ out_ptr,
                               int64_t* __restrict__ out_idx, double* __restrict__ out_val) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5; row < nrows;
       row += (int64_t(gridDim.x) * blockDim.x) >> 5) {
    if (free_ != nullptr && !free_[row0 + row]) continue;
    int64_t o = out_ptr[colmap[row0 + row]];
    const int64_t a = indptr[row], e = indptr[row + 1];
    for (int64_t k0 = a; k0 < e; k0 += 32) {
      const int64_t k = k0 + lane;
      int64_t col = -1;
      bool keep = false;
      if (k < e) {
        col = indices[k];
        keep = keep_entry(free_, upper != 0, row0 + row, col);
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int64_t dst = o + __popc(m & ((1u << lane) - 1u));
        if (out_idx != nullptr) out_idx[dst] = colmap[col];
        if (out_val != nullptr) out_val[dst] = vals[k];
      }
      o += __popc(m);
    }
  }
}

unsigned rows_grid(int64_t nrows) {
  const int64_t want = (nrows * 32 + 255) / 256;
  return unsigned(std::max<int64_t>(1, std::min<int64_t>(want, 148 * 64)));
}

}  // namespace

// colmap[ncols + 1]: exclusive scan of the free flags (colmap[ncols] = number of free DOFs); out_ptr[nfree_rows + 1].
// Returns the number of kept rows and of kept entries through nkeep / nnz (host).
int csr_compact_symbolic(cudaStream_t st, int64_t nrows, int64_t ncols, const int64_t* indptr, const int64_t* indices,
                         const unsigned char* free_, int upper, int64_t row0, int64_t* colmap, int64_t* out_ptr,
                         int64_t* nkeep, int64_t* nnz, int64_t* launches) {
  if (nrows < 0 || ncols <= 0 || row0 < 0 || row0 + nrows > ncols) return PF3_E_BAD_ARG;
  // colmap = exclusive scan of the flags, in place
  k_flag_i64<<<cg_grid(ncols), kCgThreads, 0, st>>>(ncols, free_, colmap);
  ++*launches;
  PF3_CUDA(cudaMemsetAsync(colmap + ncols, 0, sizeof(int64_t), st));
  void* tmp = nullptr;
  size_t tb = 0;
  PF3_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, colmap, colmap, ncols + 1, st));
  PF3_CUDA(cudaMallocAsync(&tmp, tb, st));
  cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tb, colmap, colmap, ncols + 1, st);
  ++*launches;
  if (e != cudaSuccess) {
    cudaFreeAsync(tmp, st);
    return int(e);
  }
  int64_t lo = 0, hi = 0;
  PF3_CUDA(cudaMemcpyAsync(&lo, colmap + row0, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PF3_CUDA(cudaMemcpyAsync(&hi, colmap + row0 + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PF3_CUDA(cudaStreamSynchronize(st));
  const int64_t keep = hi - lo;
  // row counts of the kept rows, at their rank RELATIVE to the first kept row of this block
  PF3_CUDA(cudaMemsetAsync(out_ptr, 0, size_t(keep + 1) * sizeof(int64_t), st));
  if (nrows > 0 && keep > 0) {
    // colmap is global; the local rank is colmap[row0 + row] - lo: pass a shifted output pointer
    k_compact_count<<<rows_grid(nrows), 256, 0, st>>>(nrows, indptr, indices, free_, upper, row0, colmap, out_ptr - lo);
    ++*launches;
  }
  size_t tb2 = 0;
  PF3_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, out_ptr, out_ptr, keep + 1, st));
  if (tb2 > tb) {
    cudaFreeAsync(tmp, st);
    PF3_CUDA(cudaMallocAsync(&tmp, tb2, st));
  }
  e = cub::DeviceScan::ExclusiveSum(tmp, tb2, out_ptr, out_ptr, keep + 1, st);
  ++*launches;
  cudaFreeAsync(tmp, st);
  if (e != cudaSuccess) return int(e);
  int64_t total = 0;
  PF3_CUDA(cudaMemcpyAsync(&total, out_ptr + keep, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PF3_CUDA(cudaStreamSynchronize(st));
  if (nkeep) *nkeep = keep;
  if (nnz) *nnz = total;
  return PF3_OK;
}

int csr_compact_fill(cudaStream_t st, int64_t nrows, int64_t ncols, const int64_t* indptr, const int64_t* indices,
                     const double* vals, const unsigned char* free_, int upper, int64_t row0, const int64_t* colmap,
                     const int64_t* out_ptr, int64_t* out_idx, double* out_val, int64_t* launches) {
  if (nrows <= 0) return PF3_OK;
  if (row0 < 0 || row0 + nrows > ncols) return PF3_E_BAD_ARG;
  int64_t lo = 0;
  PF3_CUDA(cudaMemcpyAsync(&lo, colmap + row0, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PF3_CUDA(cudaStreamSynchronize(st));
  k_compact_fill<<<rows_grid(nrows), 256, 0, st>>>(nrows, indptr, indices, vals, free_, upper, row0, colmap, out_ptr - lo,
                                                   out_idx, out_val);
  ++*launches;
  return int(cudaGetLastError());
}

}  // namespace pf3
