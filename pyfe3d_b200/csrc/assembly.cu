// GPU COO -> CSR sum-duplicates assembly.
//
// Replaces the `scipy.sparse.coo_matrix((v,(r,c)),shape=(N,N)).tocsc()/.tocsr()` call every
// reference script makes right after its element loop (tests/test_quad4_static_point_load.py:80,
// tests/test_beamc_natural_freq_curved.py:93-94 in /root/reference).
//
// Two plans:
//  * structured: built from connectivity alone.  COO indices are a pure function of connectivity
//    (SURVEY Appendix B), so the symbolic phase sorts NODE-PAIR keys (16 per quad instead of 576
//    entry keys), and the numeric phase is a deterministic gather: one warp per node row block reads
//    the contiguous row slab of every incident element (coalesced) and writes its CSR rows once.
//  * generic: radix sort of (row, col) keys of arbitrary COO arrays + segmented sum.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <vector>

#include "common.cuh"
#include "pattern.hpp"

namespace pf3 {

constexpr int kMaxGroups = 8;

struct GroupDev {
  const int64_t* conn;
  const uint16_t* tab;  // per slab-local entry: dof_i | node_j<<3 | colrank<<6 (node_j==7: same node)
  int64_t ne;
  int64_t coo_offset;
  int64_t pairbase, incbase;
  int nn, size, wn, diag, npairs;
};

struct PlanDev {
  GroupDev g[kMaxGroups];
  int ngroups;
  int64_t nnodes, node_begin, node_end;
  int cnt[6], rowoff[6], mc;
};

}  // namespace pf3

struct pf3_plan {
  int device = 0;
  int generic = 0;
  int64_t nrows = 0, nnz = 0;
  // structured
  pf3::PlanDev dev{};
  int8_t colrank[6][6];
  bool umask[6][6];
  int64_t nblk = 0, ninc = 0, npair = 0, nown = 0;
  int max_nb = 0;
  int degenerate = 0;
  int64_t* d_brow_ptr = nullptr;
  int64_t* d_bcol = nullptr;
  int64_t* d_inc_ptr = nullptr;
  int64_t* d_inc_src = nullptr;
  int64_t* d_inc_pair0 = nullptr;
  int32_t* d_inc_meta = nullptr;
  int32_t* d_slot = nullptr;
  mutable pf3::NodeRec* d_noderec = nullptr;   // built on first fused use
  mutable int2* d_pftab = nullptr;             // L2 prefetch table of the fused quad kernel (common.cuh: kPfChunk)
  mutable int64_t pf_nchunks = 0;
  mutable pf3::TriaRec* d_triarec = nullptr;   // ditto, Tria3R
  mutable int rmax = 0;
  mutable std::vector<pf3::NodeRec*> d_grecs;  // per-group records of the slab assembly (built on first use)
  mutable std::vector<int> grmax;
  std::vector<int> gmask;                      // MaskId of every group
  std::vector<int> gkind;                      // element kind of every group
  mutable pf3::NodeRec* d_frecs = nullptr;     // fused-kernel records of ONE group of a multi-group plan
  mutable int frecs_group = -1, frmax = 0;
  std::vector<uint16_t*> d_tabs;
  // generic
  int64_t n = 0, nnz_coo = 0;
  int64_t* d_indptr = nullptr;
  int64_t* d_indices = nullptr;
  int64_t* d_perm = nullptr;
  int64_t* d_seg = nullptr;
};

namespace pf3 {

#define PF3_CUDA(x)                      \
  do {                                   \
    cudaError_t _e = (x);                \
    if (_e != cudaSuccess) return int(_e); \
  } while (0)

namespace {

template <class T>
int dalloc(T** p, int64_t n) {
  *p = nullptr;
  if (n <= 0) n = 1;
  return int(cudaMalloc((void**)p, size_t(n) * sizeof(T)));
}

__device__ __forceinline__ int find_group_by(const PlanDev& P, int64_t id, bool pair) {
  int g = 0;
  for (int k = 1; k < P.ngroups; ++k)
    if (id >= (pair ? P.g[k].pairbase : P.g[k].incbase)) g = k;
  return g;
}

// ---- symbolic kernels ---------------------------------------------------------------------
__global__ void k_pair_keys(const PlanDev P, int gi, int64_t* keys, uint32_t* vals) {
  const GroupDev& G = P.g[gi];
  const int64_t n = G.ne * G.npairs;
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < n; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t e = t / G.npairs;
    const int q = int(t - e * G.npairs);
    const int a = G.diag ? q : q / G.nn, b = G.diag ? q : q % G.nn;
    const int64_t na = G.conn[e * G.nn + a], nb = G.conn[e * G.nn + b];
    const bool own = na >= P.node_begin && na < P.node_end && nb >= 0 && nb < P.nnodes;   // bad ids: rejected below
    keys[G.pairbase + t] = own ? na * P.nnodes + nb : P.nnodes * P.nnodes;
    vals[G.pairbase + t] = uint32_t(G.pairbase + t);
  }
}

__global__ void k_inc_keys(const PlanDev P, int gi, int64_t* keys, uint32_t* vals, int* degenerate) {
  const GroupDev& G = P.g[gi];
  const int64_t n = G.ne * G.nn;
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < n; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t e = t / G.nn;
    const int a = int(t - e * G.nn);
    const int64_t na = G.conn[t];
    if (na < 0 || na >= P.nnodes) degenerate[1] = 1;   // node id outside [0, nnodes): would alias keys na*nnodes+nb
    for (int b = 0; b < a; ++b)
      if (G.conn[e * G.nn + b] == na) *degenerate = 1;
    const bool own = na >= P.node_begin && na < P.node_end;
    keys[G.incbase + t] = own ? na : P.nnodes;
    vals[G.incbase + t] = uint32_t(G.incbase + t);
  }
}

__global__ void k_head_flags(const int64_t* keys, int64_t n, int64_t sentinel, int32_t* flags) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < n; t += int64_t(gridDim.x) * blockDim.x)
    flags[t] = (keys[t] != sentinel && (t == 0 || keys[t] != keys[t - 1])) ? 1 : 0;
}

__global__ void k_scatter_unique(const int64_t* keys, const int32_t* flags, const int32_t* incl, int64_t n,
                                 int64_t* ukeys) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < n; t += int64_t(gridDim.x) * blockDim.x)
    if (flags[t]) ukeys[incl[t] - 1] = keys[t];
}

// lower_bound of target in sorted a[0..n)
__device__ __forceinline__ int64_t lower_bound_dev(const int64_t* a, int64_t n, int64_t target) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void k_row_ptr(const int64_t* sorted, int64_t n, int64_t first, int64_t count, int64_t scale,
                          int64_t* ptr) {
  // ptr[i] = lower_bound(sorted, (first+i)*scale), i in [0, count]
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i <= count; i += int64_t(gridDim.x) * blockDim.x)
    ptr[i] = lower_bound_dev(sorted, n, (first + i) * scale);
}

__global__ void k_bcol_maxnb(const int64_t* ukeys, int64_t nblk, int64_t nnodes, int64_t* bcol,
                             const int64_t* brow_ptr, int64_t nown, int* max_nb) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < max(nblk, nown);
       t += int64_t(gridDim.x) * blockDim.x) {
    if (t < nblk) bcol[t] = ukeys[t] % nnodes;
    if (t < nown) atomicMax(max_nb, int(brow_ptr[t + 1] - brow_ptr[t]));
  }
}

__global__ void k_pair_slots(const PlanDev P, const int64_t* keys, const uint32_t* vals, const int32_t* incl,
                             int64_t n, const int64_t* brow_ptr, int32_t* slot) {
  const int64_t sentinel = P.nnodes * P.nnodes;
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < n; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t k = keys[t];
    if (k == sentinel) {
      slot[vals[t]] = -1;
      continue;
    }
    const int64_t row = k / P.nnodes - P.node_begin;
    slot[vals[t]] = int32_t(int64_t(incl[t] - 1) - brow_ptr[row]);
  }
}

__global__ void k_inc_fill(const PlanDev P, const uint32_t* vals, int64_t ninc_valid, int64_t* inc_src,
                           int64_t* inc_pair0, int32_t* inc_meta) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < ninc_valid;
       t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t id = vals[t];
    const int gi = find_group_by(P, id, false);
    const GroupDev& G = P.g[gi];
    const int64_t loc = id - G.incbase;
    const int64_t e = loc / G.nn;
    const int a = int(loc - e * G.nn);
    inc_src[t] = G.coo_offset + e * G.size + int64_t(a) * G.wn;
    inc_pair0[t] = G.pairbase + e * G.npairs + (G.diag ? a : a * G.nn);
    inc_meta[t] = gi | (a << 8);
  }
}

// ---- pattern -----------------------------------------------------------------------------
struct MaskDev {
  int8_t cols[6][6];  // cols[d][r] = r-th column dof of row d
};

__global__ void k_pattern(const PlanDev P, const MaskDev M, const int64_t* brow_ptr, const int64_t* bcol,
                          int64_t nown, int64_t* indptr, int64_t* indices) {
  // one thread per (owned node, dof row)
  const int64_t n = nown * 6;
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t <= n; t += int64_t(gridDim.x) * blockDim.x) {
    if (t == n) {
      if (indptr) indptr[n] = brow_ptr[nown] * P.mc;
      continue;
    }
    const int64_t i = t / 6;
    const int d = int(t - i * 6);
    const int64_t b0 = brow_ptr[i], nb = brow_ptr[i + 1] - b0;
    const int64_t start = b0 * P.mc + int64_t(P.rowoff[d]) * nb;
    if (indptr) indptr[t] = start;
    if (indices)
      for (int64_t s = 0; s < nb; ++s)
        for (int r = 0; r < P.cnt[d]; ++r) indices[start + s * P.cnt[d] + r] = 6 * bcol[b0 + s] + M.cols[d][r];
  }
}

// ---- numeric -----------------------------------------------------------------------------
// One warp per owned node.  acc lives in shared memory (max_nb * mc doubles per warp).
// skip_group >= 0 (add mode): csr_v += the contributions of every group EXCEPT skip_group; nodes that have none are
// not touched (the fused kernel has already written group skip_group's share of every row).
// GLOBAL: the accumulator is the node's CSR row block itself (hub nodes coupled to hundreds of nodes, e.g. spider / RBE
// style springs, whose row block does not fit shared memory): slower read-modify-write of global memory, same order.
template <bool SAFE, bool GLOBAL = false>
__global__ void __launch_bounds__(256) k_assemble(const PlanDev P, const int64_t* __restrict__ brow_ptr,
                                                  const int64_t* __restrict__ inc_ptr,
                                                  const int64_t* __restrict__ inc_src,
                                                  const int64_t* __restrict__ inc_pair0,
                                                  const int32_t* __restrict__ inc_meta,
                                                  const int32_t* __restrict__ slot, int64_t nown, int acc_stride,
                                                  const double* __restrict__ coo_v, double* __restrict__ csr_v,
                                                  int skip_group) {
  extern __shared__ double sacc[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  double* acc = sacc + size_t(warp) * acc_stride;
  for (int64_t i = int64_t(blockIdx.x) * wpc + warp; i < nown; i += int64_t(gridDim.x) * wpc) {
    const int64_t b0 = brow_ptr[i];
    const int nb = int(brow_ptr[i + 1] - b0);
    const int nout = nb * P.mc;
    if (GLOBAL) acc = csr_v + b0 * P.mc;
    const int64_t q0 = inc_ptr[i], q1 = inc_ptr[i + 1];
    if (skip_group >= 0) {
      bool any = false;
      for (int64_t q = q0 + lane; q < q1; q += 32) any = any || ((inc_meta[q] & 0xff) != skip_group);
      if (__ballot_sync(0xffffffffu, any) == 0u) continue;
    }
    if (!(GLOBAL && skip_group >= 0))
      for (int k = lane; k < nout; k += 32) acc[k] = 0.;
    __syncwarp();
    for (int64_t q = q0; q < q1; ++q) {
      const int meta = inc_meta[q];
      if ((meta & 0xff) == skip_group) continue;
      const GroupDev& G = P.g[meta & 0xff];
      const int a = meta >> 8;
      const double* src = coo_v + inc_src[q];
      const int32_t* sl = slot + inc_pair0[q];
      const uint16_t* tab = G.tab;
      for (int l = lane; l < G.wn; l += 32) {
        const int code = tab[l];
        const int di = code & 7, bj = (code >> 3) & 7, rk = code >> 6;
        const int s = G.diag ? sl[0] : sl[bj];
        (void)a;
        const int dest = P.rowoff[di] * nb + s * P.cnt[di] + rk;
        const double v = src[l];
        if (SAFE) atomicAdd(&acc[dest], v); else acc[dest] += v;
      }
      __syncwarp();
    }
    if (!GLOBAL) {
      double* out = csr_v + b0 * P.mc;
      if (skip_group >= 0) {
        for (int k = lane; k < nout; k += 32) out[k] += acc[k];
      } else {
        for (int k = lane; k < nout; k += 32) out[k] = acc[k];
      }
    }
    __syncwarp();
  }
}

// ---- generic COO plan -----------------------------------------------------------------------
__global__ void k_coo_keys(const int64_t* r, const int64_t* c, int64_t n, int64_t nnz, int64_t* keys, int64_t* vals) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < nnz; t += int64_t(gridDim.x) * blockDim.x) {
    keys[t] = r[t] * n + c[t];
    vals[t] = t;
  }
}
__global__ void k_seg_starts(const int32_t* flags, const int32_t* incl, int64_t n, int64_t* seg) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t <= n; t += int64_t(gridDim.x) * blockDim.x) {
    if (t == n) {
      seg[n ? incl[n - 1] : 0] = n;
    } else if (flags[t]) {
      seg[incl[t] - 1] = t;
    }
  }
}
__global__ void k_indices_from_keys(const int64_t* ukeys, int64_t nnz, int64_t n, int64_t* indices) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < nnz; t += int64_t(gridDim.x) * blockDim.x)
    indices[t] = ukeys[t] % n;
}
__global__ void k_segsum(const int64_t* seg, const int64_t* perm, int64_t nnz, const double* v, double* out) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < nnz; t += int64_t(gridDim.x) * blockDim.x) {
    double s = 0.;
    for (int64_t k = seg[t]; k < seg[t + 1]; ++k) s += v[perm[k]];
    out[t] = s;
  }
}

// ---- SpMV ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_spmv(int64_t nrows, const int64_t* __restrict__ indptr,
                                              const int64_t* __restrict__ indices, const double* __restrict__ vals,
                                              const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = w; row < nrows; row += nw) {
    double s = 0.;
    for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32) s += vals[k] * x[indices[k]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = s;
  }
}

// The same with two entries per lane and 16-byte loads of indices and values (16 B per nonzero is the whole traffic of
// this kernel: the int64 CSR a scipy-shaped consumer holds): entries are paired from the first EVEN position of the
// row, an odd first / last entry is taken by lane 0 / lane 1.  Needs 16-byte aligned index and value arrays.
__global__ void __launch_bounds__(256) k_spmv_v2(int64_t nrows, const int64_t* __restrict__ indptr,
                                                 const int64_t* __restrict__ indices, const double* __restrict__ vals,
                                                 const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = w; row < nrows; row += nw) {
    const int64_t a = indptr[row], b = indptr[row + 1];
    const int64_t a2 = a + (a & 1);
    double s0 = 0., s1 = 0.;
    if (lane == 0 && a2 > a && a < b) s0 = vals[a] * x[indices[a]];
    if (lane == 1 && b > a2 && ((b - a2) & 1)) s1 = vals[b - 1] * x[indices[b - 1]];
    for (int64_t k = a2 + 2 * lane; k + 1 < b; k += 64) {
      const longlong2 ij = *reinterpret_cast<const longlong2*>(indices + k);
      const double2 v = *reinterpret_cast<const double2*>(vals + k);
      s0 += v.x * x[ij.x];
      s1 += v.y * x[ij.y];
    }
    double s = s0 + s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = s;
  }
}

inline unsigned grid_for(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  return unsigned(std::max<int64_t>(1, std::min<int64_t>(g, 148 * 32)));
}

int bits_for(int64_t maxval) {
  int b = 1;
  while (b < 63 && (int64_t(1) << b) <= maxval) ++b;
  return b;
}

template <class K, class V>
int sort_pairs(K* kin, K* kout, V* vin, V* vout, int64_t n, int end_bit, cudaStream_t st) {
  size_t bytes = 0;
  PF3_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, end_bit, st));
  void* tmp = nullptr;
  PF3_CUDA(cudaMalloc(&tmp, bytes ? bytes : 1));
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, n, 0, end_bit, st);
  cudaStreamSynchronize(st);
  cudaFree(tmp);
  return int(e);
}

int inclusive_sum(const int32_t* in, int32_t* out, int64_t n, cudaStream_t st) {
  size_t bytes = 0;
  PF3_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, n, st));
  void* tmp = nullptr;
  PF3_CUDA(cudaMalloc(&tmp, bytes ? bytes : 1));
  cudaError_t e = cub::DeviceScan::InclusiveSum(tmp, bytes, in, out, n, st);
  cudaStreamSynchronize(st);
  cudaFree(tmp);
  return int(e);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
int plan_create_structured(int device, cudaStream_t st, int matrix, int64_t nnodes, int ngroups,
                           const pf3_batch* groups, const int64_t* coo_offsets, int64_t node_begin,
                           int64_t node_end, int64_t* launches, pf3_plan** out) {
  if (ngroups <= 0 || ngroups > kMaxGroups || nnodes <= 0 || node_begin < 0 || node_end > nnodes ||
      node_begin >= node_end)
    return PF3_E_BAD_ARG;
  pf3_plan* pl = new pf3_plan();
  pl->device = device;
  PlanDev& P = pl->dev;
  P.ngroups = ngroups;
  P.nnodes = nnodes;
  P.node_begin = node_begin;
  P.node_end = node_end;
  pl->nown = node_end - node_begin;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) pl->umask[i][j] = false;
  std::vector<BlockLayout> lay(ngroups);
  int64_t pairbase = 0, incbase = 0;
  for (int g = 0; g < ngroups; ++g) {
    lay[g] = make_layout(groups[g].kind, matrix, groups[g].mtype);
    const BlockLayout& L = lay[g];
    if (L.written == 0 || groups[g].ne < 0 || groups[g].conn == nullptr) {
      delete pl;
      return PF3_E_BAD_ARG;
    }
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) pl->umask[i][j] |= mask_has(L.mask, i, j);
    pl->gmask.push_back(L.mask);
    pl->gkind.push_back(groups[g].kind);
    GroupDev& G = P.g[g];
    G.conn = groups[g].conn;
    G.ne = groups[g].ne;
    G.coo_offset = coo_offsets ? coo_offsets[g] : 0;
    G.nn = L.nn;
    G.size = L.size;
    G.wn = L.written / L.nn;
    G.diag = L.diag_pairs ? 1 : 0;
    G.npairs = L.diag_pairs ? L.nn : L.nn * L.nn;
    G.pairbase = pairbase;
    G.incbase = incbase;
    pairbase += G.ne * G.npairs;
    incbase += G.ne * G.nn;
  }
  pl->npair = pairbase;
  pl->ninc = incbase;
  if (pairbase >= (int64_t(1) << 31) || incbase >= (int64_t(1) << 31)) {   // int32 scans of head flags, uint32 ids
    delete pl;
    return PF3_E_CAPACITY;
  }
  P.mc = 0;
  for (int i = 0; i < 6; ++i) {
    P.rowoff[i] = P.mc;
    int c = 0;
    for (int j = 0; j < 6; ++j) {
      pl->colrank[i][j] = -1;
      if (pl->umask[i][j]) {
        pl->colrank[i][j] = int8_t(c);
        ++c;
      }
    }
    P.cnt[i] = c;
    P.mc += c;
  }
  // per-group slab tables
  for (int g = 0; g < ngroups; ++g) {
    const BlockLayout& L = lay[g];
    std::vector<uint16_t> tab(P.g[g].wn);
    for (int l = 0; l < P.g[g].wn; ++l) {  // slab of local node 0 (identical for every node)
      const int di = L.li[l], bj = L.diag_pairs ? 7 : L.lb[l], rk = pl->colrank[di][L.lj[l]];
      tab[l] = uint16_t(di | (bj << 3) | (rk << 6));
    }
    uint16_t* d = nullptr;
    int rc = dalloc(&d, int64_t(tab.size()));
    if (rc) { pf3_plan_destroy(pl); return rc; }
    cudaMemcpyAsync(d, tab.data(), tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    pl->d_tabs.push_back(d);
    P.g[g].tab = d;
  }

  int rc = 0;
  int64_t *keys = nullptr, *keys2 = nullptr, *ukeys = nullptr;
  uint32_t *vals = nullptr, *vals2 = nullptr;
  int32_t *flags = nullptr, *incl = nullptr;
  int* d_flagsmall = nullptr;
  auto cleanup = [&]() {
    cudaFree(keys); cudaFree(keys2); cudaFree(ukeys); cudaFree(vals); cudaFree(vals2);
    cudaFree(flags); cudaFree(incl); cudaFree(d_flagsmall);
  };
#define PF3_TRY(x) do { rc = (x); if (rc) { cleanup(); pf3_plan_destroy(pl); return rc; } } while (0)
  const int64_t NP = pl->npair, NI = pl->ninc, NMAX = std::max(NP, NI);
  PF3_TRY(dalloc(&keys, NMAX));
  PF3_TRY(dalloc(&keys2, NMAX));
  PF3_TRY(dalloc(&vals, NMAX));
  PF3_TRY(dalloc(&vals2, NMAX));
  PF3_TRY(dalloc(&flags, NMAX));
  PF3_TRY(dalloc(&incl, NMAX));
  PF3_TRY(dalloc(&d_flagsmall, 3));
  PF3_TRY(int(cudaMemsetAsync(d_flagsmall, 0, 3 * sizeof(int), st)));

  // ---- node-pair blocks
  for (int g = 0; g < ngroups; ++g) {
    k_pair_keys<<<grid_for(P.g[g].ne * P.g[g].npairs), 256, 0, st>>>(P, g, keys, vals);
    ++*launches;
  }
  const int64_t sentinel = nnodes * nnodes;
  PF3_TRY(sort_pairs(keys, keys2, vals, vals2, NP, bits_for(sentinel), st));
  k_head_flags<<<grid_for(NP), 256, 0, st>>>(keys2, NP, sentinel, flags);
  PF3_TRY(inclusive_sum(flags, incl, NP, st));
  int32_t nblk32 = 0;
  if (NP > 0) PF3_TRY(int(cudaMemcpyAsync(&nblk32, incl + NP - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st)));
  cudaStreamSynchronize(st);
  pl->nblk = nblk32;
  PF3_TRY(dalloc(&ukeys, pl->nblk));
  k_scatter_unique<<<grid_for(NP), 256, 0, st>>>(keys2, flags, incl, NP, ukeys);
  PF3_TRY(dalloc(&pl->d_brow_ptr, pl->nown + 1));
  PF3_TRY(dalloc(&pl->d_bcol, pl->nblk));
  k_row_ptr<<<grid_for(pl->nown + 1), 256, 0, st>>>(ukeys, pl->nblk, node_begin, pl->nown, nnodes, pl->d_brow_ptr);
  k_bcol_maxnb<<<grid_for(std::max(pl->nblk, pl->nown)), 256, 0, st>>>(ukeys, pl->nblk, nnodes, pl->d_bcol,
                                                                         pl->d_brow_ptr, pl->nown, d_flagsmall);
  PF3_TRY(dalloc(&pl->d_slot, NP));
  k_pair_slots<<<grid_for(NP), 256, 0, st>>>(P, keys2, vals2, incl, NP, pl->d_brow_ptr, pl->d_slot);
  *launches += 5;

  // ---- node -> (element, local node) incidences
  for (int g = 0; g < ngroups; ++g) {
    k_inc_keys<<<grid_for(P.g[g].ne * P.g[g].nn), 256, 0, st>>>(P, g, keys, vals, d_flagsmall + 1);
    ++*launches;
  }
  PF3_TRY(sort_pairs(keys, keys2, vals, vals2, NI, bits_for(nnodes), st));
  PF3_TRY(dalloc(&pl->d_inc_ptr, pl->nown + 1));
  k_row_ptr<<<grid_for(pl->nown + 1), 256, 0, st>>>(keys2, NI, node_begin, pl->nown, 1, pl->d_inc_ptr);
  // incidences of owned nodes occupy sorted positions [inc_ptr[0], inc_ptr[nown]); store them rebased
  int64_t h_first = 0, h_last = 0;
  PF3_TRY(int(cudaMemcpyAsync(&h_first, pl->d_inc_ptr, sizeof(int64_t), cudaMemcpyDeviceToHost, st)));
  PF3_TRY(int(cudaMemcpyAsync(&h_last, pl->d_inc_ptr + pl->nown, sizeof(int64_t), cudaMemcpyDeviceToHost, st)));
  int h_small[3] = {0, 0, 0};
  PF3_TRY(int(cudaMemcpyAsync(h_small, d_flagsmall, 3 * sizeof(int), cudaMemcpyDeviceToHost, st)));
  cudaStreamSynchronize(st);
  pl->max_nb = h_small[0];
  pl->degenerate = h_small[1];
  if (h_small[2]) PF3_TRY(PF3_E_BAD_ARG);   // connectivity refers to a node outside [0, nnodes)
  const int64_t nvalid = h_last;  // sorted keys < nnodes come first; owned range may start after 0
  PF3_TRY(dalloc(&pl->d_inc_src, nvalid));
  PF3_TRY(dalloc(&pl->d_inc_pair0, nvalid));
  PF3_TRY(dalloc(&pl->d_inc_meta, nvalid));
  k_inc_fill<<<grid_for(nvalid), 256, 0, st>>>(P, vals2, nvalid, pl->d_inc_src, pl->d_inc_pair0, pl->d_inc_meta);
  *launches += 2;
  (void)h_first;
  PF3_TRY(int(cudaStreamSynchronize(st)));
  PF3_TRY(int(cudaGetLastError()));
  cleanup();
#undef PF3_TRY
  pl->nrows = 6 * pl->nown;
  pl->nnz = pl->nblk * P.mc;
  *out = pl;
  return PF3_OK;
}

int plan_pattern(const pf3_plan* pl, cudaStream_t st, int64_t* indptr, int64_t* indices, int64_t* launches) {
  if (pl->generic) {
    if (indptr) PF3_CUDA(cudaMemcpyAsync(indptr, pl->d_indptr, size_t(pl->nrows + 1) * 8, cudaMemcpyDeviceToDevice, st));
    if (indices) PF3_CUDA(cudaMemcpyAsync(indices, pl->d_indices, size_t(pl->nnz) * 8, cudaMemcpyDeviceToDevice, st));
    return PF3_OK;
  }
  MaskDev MD;
  for (int i = 0; i < 6; ++i) {
    int c = 0;
    for (int j = 0; j < 6; ++j) {
      MD.cols[i][j] = 0;
      if (pl->umask[i][j]) MD.cols[i][c++] = int8_t(j);
    }
  }
  k_pattern<<<grid_for(pl->nown * 6 + 1), 256, 0, st>>>(pl->dev, MD, pl->d_brow_ptr, pl->d_bcol, pl->nown, indptr, indices);
  ++*launches;
  return int(cudaGetLastError());
}

// ---- slab assembly ------------------------------------------------------------------------------
// Per group: node records (4 incidences per round) + a kernel that stages the incident elements' COO row
// slabs in shared memory with cp.async (each slab is contiguous in the COO array) and reduces them into the
// node's CSR rows in the fixed order k = 0..3.  A half-warp per node, 8 slabs in flight per warp.
namespace {
__global__ void k_group_valence(const PlanDev P, int gi, const int64_t* inc_ptr, const int32_t* inc_meta,
                                int64_t nown, int* out) {
  for (int64_t n = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; n < nown; n += int64_t(gridDim.x) * blockDim.x) {
    int c = 0;
    for (int64_t q = inc_ptr[n]; q < inc_ptr[n + 1]; ++q) c += ((inc_meta[q] & 0xff) == gi);
    atomicMax(out, c);
  }
}
__global__ void k_group_records(const PlanDev P, int gi, const int64_t* __restrict__ brow_ptr,
                                const int64_t* __restrict__ inc_ptr, const int64_t* __restrict__ inc_pair0,
                                const int32_t* __restrict__ inc_meta, const int32_t* __restrict__ slot,
                                int64_t nown, int rmax, NodeRec* __restrict__ out) {
  const GroupDev& G = P.g[gi];
  const int64_t total = nown * rmax;
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t n = t / rmax;
    const int r = int(t - n * rmax);
    NodeRec R;
    R.b0 = brow_ptr[n];
    R.nb = uint8_t(brow_ptr[n + 1] - R.b0);
    for (int s = 0; s < 16; ++s) R.gmap[s] = 0xFFFF;
    for (int k = 0; k < 4; ++k) R.inc[k] = -1;
    int seen = 0, cnt = 0;
    for (int64_t q = inc_ptr[n]; q < inc_ptr[n + 1]; ++q) {
      const int meta = inc_meta[q];
      if ((meta & 0xff) != gi) continue;
      const int k = seen - 4 * r;
      ++seen;
      if (k < 0 || k >= 4) continue;
      const int a = meta >> 8;
      const int64_t p0 = inc_pair0[q];
      const int64_t e = (p0 - G.pairbase) / G.npairs;
      R.inc[k] = int32_t(e * G.nn + a);
      ++cnt;
      if (G.diag) {
        const int sl = slot[p0];
        if (sl >= 0 && sl < 16) R.gmap[sl] = uint16_t(R.gmap[sl] & ~(0xF << (4 * k)));
      } else {
        for (int b = 0; b < G.nn; ++b) {
          const int sl = slot[p0 + b];
          if (sl >= 0 && sl < 16) R.gmap[sl] = uint16_t((R.gmap[sl] & ~(0xF << (4 * k))) | (b << (4 * k)));
        }
      }
    }
    R.v = uint8_t(cnt);
    for (int i = 0; i < 6; ++i) R.pad[i] = 0;
    out[t] = R;
  }
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

// NNS: node blocks per slab row (element nodes, or 1 for diagonal-pair-only groups); NR x CNT: masked rows/cols
template <int NNS, int NR, int CNT>
__global__ void __launch_bounds__(256) k_assemble_slabs(const NodeRec* __restrict__ recs, int rmax, int64_t nown,
                                                        const double* __restrict__ coo, int nn, int size,
                                                        double* __restrict__ csr, int accumulate, int vec16) {
  constexpr int kSlab = NR * NNS * CNT;
  constexpr int kLd = kSlab + 2 - (kSlab & 1);      // even (16-B aligned rows for the paired path), != 0 mod 16
  constexpr bool kVec = (kSlab % 2 == 0);            // slab start is 16-B aligned iff size and wn are even
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  double* st = smem + warp * (8 * kLd);
  const int h = lane >> 4, l16 = lane & 15, k = l16 >> 2, qd = l16 & 3;
  const int64_t npairs = (nown + 1) >> 1;
  for (int64_t np = int64_t(blockIdx.x) * wpc + warp; np < npairs; np += int64_t(gridDim.x) * wpc) {
    const int64_t n = 2 * np + h;
    for (int r = 0; r < rmax; ++r) {
      int inc = -1, nb = 0;
      int64_t b0 = 0;
      const NodeRec* nr = nullptr;
      if (n < nown) {
        nr = recs + n * rmax + r;
        inc = nr->inc[k];
        nb = nr->nb;
        b0 = nr->b0;
      }
      const bool any = __ballot_sync(0xffffffffu, inc >= 0) != 0u;
      const bool first = (r == 0) && !accumulate;
      if (!any && !first) continue;
      if (inc >= 0) {
        const int e = inc / nn, a = inc - e * nn;
        const double* src = coo + int64_t(e) * size + int64_t(a) * kSlab;
        double* dst = st + (lane >> 2) * kLd;
        if (kVec && vec16) {
          for (int c = qd; c < kSlab / 2; c += 4)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst + 2 * c)), "l"(src + 2 * c) : "memory");
        } else {
          for (int c = qd; c < kSlab; c += 4)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(dst + c)), "l"(src + c) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      if (nb > 0) {
        const int w = nb * CNT;
        const double* sh = st + h * 4 * kLd;
        double* out = csr + b0 * (NR * CNT);
        for (int x = l16; x < w; x += 16) {
          const int s = x / CNT, rr = x - s * CNT;
          const unsigned gm = nr->gmap[s];
          int off[4];
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            const int nib = (gm >> (4 * k2)) & 0xF;
            off[k2] = (nib != 0xF) ? (k2 * kLd + nib * CNT + rr) : -1;
          }
#pragma unroll
          for (int d = 0; d < NR; ++d) {
            double sum = 0.;
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2)
              if (off[k2] >= 0) sum += sh[d * NNS * CNT + off[k2]];
            double* o = out + d * w + x;
            if (first) *o = sum; else *o += sum;
          }
        }
      }
      __syncwarp();
    }
  }
}

template <int NR, int CNT>
int launch_slabs(int nns, const NodeRec* recs, int rmax, int64_t nown, const double* coo, int nn, int size,
                 double* csr, int accumulate, cudaStream_t st) {
#ifndef PF3_SLAB_WPC
#define PF3_SLAB_WPC 1   // one-warp CTAs, uncapped grid: the CTA scheduler keeps the nodes in flight a narrow window (as in quad_fused.cu)
#endif
#ifndef PF3_SLAB_CAP
#define PF3_SLAB_CAP 2000000000
#endif
  const int wpc = PF3_SLAB_WPC;
  const int64_t npairs = (nown + 1) / 2;
  const unsigned grid = unsigned(std::max<int64_t>(1, std::min<int64_t>((npairs + wpc - 1) / wpc, int64_t(PF3_SLAB_CAP))));
  const int vec16 = int((((uintptr_t)coo) & 15) == 0 && (size % 2) == 0);   // 16-B cp.async needs aligned slabs
#define PF3_SLAB_CASE(N)                                                                                       \
  case N: {                                                                                                    \
    constexpr int kSlab = NR * N * CNT, kLd = kSlab + 2 - (kSlab & 1);                                          \
    const size_t smem = size_t(wpc) * 8 * kLd * sizeof(double);                                                 \
    if (smem > 48 * 1024)                                                                                       \
      cudaFuncSetAttribute(k_assemble_slabs<N, NR, CNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); \
    k_assemble_slabs<N, NR, CNT><<<grid, wpc * 32, smem, st>>>(recs, rmax, nown, coo, nn, size, csr, accumulate, vec16); \
    break;                                                                                                      \
  }
  switch (nns) {
    PF3_SLAB_CASE(1)
    PF3_SLAB_CASE(2)
    PF3_SLAB_CASE(3)
    PF3_SLAB_CASE(4)
    default: return PF3_E_UNSUPPORTED;
  }
#undef PF3_SLAB_CASE
  return int(cudaGetLastError());
}

// (NR, CNT) of the plan's union mask when every row has the same count (0 rows allowed at one end)
bool uniform_shape(const PlanDev& P, int* nr, int* cnt) {
  int c = 0, rows = 0;
  for (int i = 0; i < 6; ++i) {
    if (P.cnt[i] == 0) continue;
    if (c == 0) c = P.cnt[i];
    if (P.cnt[i] != c) return false;
    ++rows;
  }
  if (!((rows == 6) || (rows == 3 && (P.cnt[0] == 0 || P.cnt[5] == 0)))) return false;
  *nr = rows;
  *cnt = c;
  return (rows == 6 && (c == 6 || c == 5 || c == 3)) || (rows == 3 && c == 3);
}
}  // namespace

int plan_assemble_slabs(const pf3_plan* pl, cudaStream_t st, const double* coo_v, double* csr_v, int64_t* launches,
                        bool* done) {
  *done = false;
  const PlanDev& P = pl->dev;
  int nr = 0, cnt = 0;
  if (pl->degenerate || pl->max_nb > 16 || !uniform_shape(P, &nr, &cnt)) return PF3_OK;
  for (int g = 0; g < P.ngroups; ++g) {   // every group must write exactly the plan's mask
    for (int i = 0; i < 6; ++i) {
      int c = 0;
      for (int j = 0; j < 6; ++j) c += mask_has(pl->gmask[g], i, j) ? 1 : 0;
      if (c != P.cnt[i]) return PF3_OK;
    }
    if (P.g[g].ne * P.g[g].nn >= (int64_t(1) << 31)) return PF3_OK;
  }
  if (pl->d_grecs.empty()) {
    pl->d_grecs.assign(P.ngroups, nullptr);
    pl->grmax.assign(P.ngroups, 1);
    for (int g = 0; g < P.ngroups; ++g) {
      int* d_mv = nullptr;
      PF3_CUDA(cudaMalloc((void**)&d_mv, sizeof(int)));
      PF3_CUDA(cudaMemsetAsync(d_mv, 0, sizeof(int), st));
      k_group_valence<<<grid_for(pl->nown), 256, 0, st>>>(P, g, pl->d_inc_ptr, pl->d_inc_meta, pl->nown, d_mv);
      int mv = 0;
      PF3_CUDA(cudaMemcpyAsync(&mv, d_mv, sizeof(int), cudaMemcpyDeviceToHost, st));
      PF3_CUDA(cudaStreamSynchronize(st));
      cudaFree(d_mv);
      pl->grmax[g] = std::max(1, (mv + 3) / 4);
      PF3_CUDA(cudaMalloc((void**)&pl->d_grecs[g], size_t(pl->nown) * pl->grmax[g] * sizeof(NodeRec)));
      k_group_records<<<grid_for(pl->nown * pl->grmax[g]), 256, 0, st>>>(P, g, pl->d_brow_ptr, pl->d_inc_ptr,
                                                                         pl->d_inc_pair0, pl->d_inc_meta, pl->d_slot,
                                                                         pl->nown, pl->grmax[g], pl->d_grecs[g]);
      *launches += 2;
      PF3_CUDA(cudaGetLastError());
    }
  }
  for (int g = 0; g < P.ngroups; ++g) {
    const GroupDev& G = P.g[g];
    const int nns = G.diag ? 1 : G.nn;
    int rc;
    const double* coo = coo_v + G.coo_offset;
    if (nr == 6 && cnt == 6) rc = launch_slabs<6, 6>(nns, pl->d_grecs[g], pl->grmax[g], pl->nown, coo, G.nn, G.size, csr_v, g > 0, st);
    else if (nr == 6 && cnt == 5) rc = launch_slabs<6, 5>(nns, pl->d_grecs[g], pl->grmax[g], pl->nown, coo, G.nn, G.size, csr_v, g > 0, st);
    else if (nr == 6 && cnt == 3) rc = launch_slabs<6, 3>(nns, pl->d_grecs[g], pl->grmax[g], pl->nown, coo, G.nn, G.size, csr_v, g > 0, st);
    else rc = launch_slabs<3, 3>(nns, pl->d_grecs[g], pl->grmax[g], pl->nown, coo, G.nn, G.size, csr_v, g > 0, st);
    ++*launches;
    if (rc) return rc;
  }
  *done = true;
  return PF3_OK;
}

// The per-node gather over the union pattern (any mix of masks).  skip_group >= 0: add every other group's share to
// csr_v (see k_assemble).
int plan_assemble_gather(const pf3_plan* pl, cudaStream_t st, const double* coo_v, double* csr_v, int skip_group,
                         int64_t* launches) {
  if (pl->generic) return PF3_E_UNSUPPORTED;
  const int acc_stride = std::max(1, pl->max_nb * pl->dev.mc);
  int wpc = 8;
  while (wpc > 1 && size_t(wpc) * acc_stride * sizeof(double) > 200 * 1024) wpc >>= 1;
  const size_t smem = size_t(wpc) * acc_stride * sizeof(double);
  if (smem > 200 * 1024) {
    // a hub node's row block exceeds shared memory: accumulate in the CSR array itself
    const unsigned gridg = unsigned(std::max<int64_t>(1, std::min<int64_t>((pl->nown + 7) / 8, 148 * 64)));
    if (pl->degenerate)
      k_assemble<true, true><<<gridg, 256, 0, st>>>(pl->dev, pl->d_brow_ptr, pl->d_inc_ptr, pl->d_inc_src, pl->d_inc_pair0,
                                                     pl->d_inc_meta, pl->d_slot, pl->nown, 0, coo_v, csr_v, skip_group);
    else
      k_assemble<false, true><<<gridg, 256, 0, st>>>(pl->dev, pl->d_brow_ptr, pl->d_inc_ptr, pl->d_inc_src, pl->d_inc_pair0,
                                                      pl->d_inc_meta, pl->d_slot, pl->nown, 0, coo_v, csr_v, skip_group);
    ++*launches;
    return int(cudaGetLastError());
  }
  const unsigned grid = unsigned(std::max<int64_t>(1, std::min<int64_t>((pl->nown + wpc - 1) / wpc, 148 * 64)));
  if (pl->degenerate) {
    if (smem > 48 * 1024) PF3_CUDA(cudaFuncSetAttribute(k_assemble<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    k_assemble<true><<<grid, wpc * 32, smem, st>>>(pl->dev, pl->d_brow_ptr, pl->d_inc_ptr, pl->d_inc_src, pl->d_inc_pair0,
                                                   pl->d_inc_meta, pl->d_slot, pl->nown, acc_stride, coo_v, csr_v, skip_group);
  } else {
    if (smem > 48 * 1024) PF3_CUDA(cudaFuncSetAttribute(k_assemble<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    k_assemble<false><<<grid, wpc * 32, smem, st>>>(pl->dev, pl->d_brow_ptr, pl->d_inc_ptr, pl->d_inc_src, pl->d_inc_pair0,
                                                    pl->d_inc_meta, pl->d_slot, pl->nown, acc_stride, coo_v, csr_v, skip_group);
  }
  ++*launches;
  return int(cudaGetLastError());
}

int plan_assemble(const pf3_plan* pl, cudaStream_t st, const double* coo_v, double* csr_v, int64_t* launches) {
  if (!pl->generic) {
    bool done = false;
    int rc = plan_assemble_slabs(pl, st, coo_v, csr_v, launches, &done);
    if (rc || done) return rc;
  }
  if (pl->generic) {
    k_segsum<<<grid_for(pl->nnz), 256, 0, st>>>(pl->d_seg, pl->d_perm, pl->nnz, coo_v, csr_v);
    ++*launches;
    return int(cudaGetLastError());
  }
  return plan_assemble_gather(pl, st, coo_v, csr_v, -1, launches);
}

int plan_create_generic(int device, cudaStream_t st, int64_t n, int64_t nnz_coo, const int64_t* r, const int64_t* c,
                        int64_t* launches, pf3_plan** out) {
  if (n <= 0 || nnz_coo < 0 || (nnz_coo > 0 && (!r || !c))) return PF3_E_BAD_ARG;
  if (nnz_coo >= (int64_t(1) << 31)) return PF3_E_CAPACITY;
  pf3_plan* pl = new pf3_plan();
  pl->device = device;
  pl->generic = 1;
  pl->n = n;
  pl->nrows = n;
  pl->nnz_coo = nnz_coo;
  int rc = 0;
  int64_t *keys = nullptr, *keys2 = nullptr, *vals = nullptr, *ukeys = nullptr;
  int32_t *flags = nullptr, *incl = nullptr;
  auto cleanup = [&]() { cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(ukeys); cudaFree(flags); cudaFree(incl); };
#define PF3_TRY(x) do { rc = (x); if (rc) { cleanup(); pf3_plan_destroy(pl); return rc; } } while (0)
  PF3_TRY(dalloc(&keys, nnz_coo));
  PF3_TRY(dalloc(&keys2, nnz_coo));
  PF3_TRY(dalloc(&vals, nnz_coo));
  PF3_TRY(dalloc(&pl->d_perm, nnz_coo));
  PF3_TRY(dalloc(&flags, nnz_coo));
  PF3_TRY(dalloc(&incl, nnz_coo));
  k_coo_keys<<<grid_for(nnz_coo), 256, 0, st>>>(r, c, n, nnz_coo, keys, vals);
  PF3_TRY(sort_pairs(keys, keys2, vals, pl->d_perm, nnz_coo, bits_for(n * n), st));
  k_head_flags<<<grid_for(nnz_coo), 256, 0, st>>>(keys2, nnz_coo, int64_t(-1), flags);
  PF3_TRY(inclusive_sum(flags, incl, nnz_coo, st));
  int32_t nu = 0;
  if (nnz_coo > 0) PF3_TRY(int(cudaMemcpyAsync(&nu, incl + nnz_coo - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st)));
  cudaStreamSynchronize(st);
  pl->nnz = nu;
  PF3_TRY(dalloc(&ukeys, pl->nnz));
  PF3_TRY(dalloc(&pl->d_seg, pl->nnz + 1));
  PF3_TRY(dalloc(&pl->d_indptr, n + 1));
  PF3_TRY(dalloc(&pl->d_indices, pl->nnz));
  k_scatter_unique<<<grid_for(nnz_coo), 256, 0, st>>>(keys2, flags, incl, nnz_coo, ukeys);
  k_seg_starts<<<grid_for(nnz_coo + 1), 256, 0, st>>>(flags, incl, nnz_coo, pl->d_seg);
  k_row_ptr<<<grid_for(n + 1), 256, 0, st>>>(ukeys, pl->nnz, 0, n, n, pl->d_indptr);
  k_indices_from_keys<<<grid_for(pl->nnz), 256, 0, st>>>(ukeys, pl->nnz, n, pl->d_indices);
  *launches += 6;
  PF3_TRY(int(cudaStreamSynchronize(st)));
  PF3_TRY(int(cudaGetLastError()));
  cleanup();
#undef PF3_TRY
  *out = pl;
  return PF3_OK;
}

// y = P A P x with P = diag(free): the boundary-condition partition K[bu,:][:,bu] of the reference scripts
// (tests/test_quad4_static_point_load.py:84-99) applied on the fly, without extracting a sub-matrix.
__global__ void __launch_bounds__(256) k_spmv_masked(int64_t nrows, const int64_t* __restrict__ indptr,
                                                     const int64_t* __restrict__ indices,
                                                     const double* __restrict__ vals,
                                                     const unsigned char* __restrict__ free_,
                                                     const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = w; row < nrows; row += nw) {
    double s = 0.;
    if (free_[row])
      for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32) {
        const int64_t c = indices[k];
        if (free_[c]) s += vals[k] * x[c];
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = s;
  }
}
__global__ void k_csr_diag(int64_t nrows, const int64_t* __restrict__ indptr, const int64_t* __restrict__ indices,
                           const double* __restrict__ vals, int64_t row0, double* __restrict__ d) {
  for (int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; r < nrows; r += int64_t(gridDim.x) * blockDim.x) {
    double v = 0.;
    for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k)
      if (indices[k] == r + row0) v += vals[k];
    d[r] = v;
  }
}

int spmv_csr_masked(cudaStream_t st, int64_t nrows, const int64_t* indptr, const int64_t* indices, const double* vals,
                    const unsigned char* free_, const double* x, double* y, int64_t* launches) {
  if (nrows <= 0) return PF3_OK;
  const int64_t blocks = (nrows * 32 + 255) / 256;
  k_spmv_masked<<<unsigned(std::min<int64_t>(blocks, 148 * 64)), 256, 0, st>>>(nrows, indptr, indices, vals, free_, x, y);
  ++*launches;
  return int(cudaGetLastError());
}
int csr_diagonal(cudaStream_t st, int64_t nrows, const int64_t* indptr, const int64_t* indices, const double* vals,
                 int64_t row0, double* d, int64_t* launches) {
  if (nrows <= 0) return PF3_OK;
  k_csr_diag<<<grid_for(nrows), 256, 0, st>>>(nrows, indptr, indices, vals, row0, d);
  ++*launches;
  return int(cudaGetLastError());
}

int spmv_csr(cudaStream_t st, int64_t nrows, const int64_t* indptr, const int64_t* indices, const double* vals,
             const double* x, double* y, int64_t* launches) {
  if (nrows <= 0) return PF3_OK;
  const int64_t blocks = (nrows * 32 + 255) / 256;
  if (((((uintptr_t)indices) | ((uintptr_t)vals)) & 15) == 0)
    k_spmv_v2<<<unsigned(std::min<int64_t>(blocks, 148 * 64)), 256, 0, st>>>(nrows, indptr, indices, vals, x, y);
  else
    k_spmv<<<unsigned(std::min<int64_t>(blocks, 148 * 64)), 256, 0, st>>>(nrows, indptr, indices, vals, x, y);
  ++*launches;
  return int(cudaGetLastError());
}

// ---- block-structured SpMV on the plan's own layout --------------------------------------------------------
// The CSR value array of a structured plan is a sequence of node row blocks [dof row d][column block s][masked
// column r] whose column blocks are listed once per NODE in bcol: y = A x needs 8 B/nonzero of values plus 8 B per
// 6x6 block of indices (8.2 B/nnz against 16 B/nnz for int64 CSR), and the 6 rows of a node share their x gathers.
// A half-warp takes one node; lanes stride over the (column block, masked column) positions of a row, rows
// inner.  free_ (nullable): P A P with P = diag(free), the boundary-condition partition of the reference scripts
// (tests/test_quad4_static_point_load.py:84-99).  Deterministic (fixed per-lane order + butterfly reduction).
struct SpmvMask {
  int8_t cols[6][6];
  int cnt[6], rowoff[6], mc;
};
// CNT > 0: every non-empty row has CNT masked columns per block (0: general); SAME: all non-empty rows share one
// column set, so the x gather (and its mask byte) is done once per position instead of once per row
template <int CNT, bool SAME>
__global__ void __launch_bounds__(128) k_plan_spmv(const SpmvMask M, const int64_t* __restrict__ brow_ptr,
                                                   const int64_t* __restrict__ bcol, int64_t nown, int64_t node_begin,
                                                   const double* __restrict__ vals,
                                                   const unsigned char* __restrict__ free_,
                                                   const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31, h = lane >> 4, l16 = lane & 15;
  const int64_t i = (int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + h;
  const bool live = i < nown;
  int64_t b0 = 0;
  int nb = 0;
  if (live) {
    b0 = brow_ptr[i];
    nb = int(brow_ptr[i + 1] - b0);
  }
  const double* v = vals + b0 * M.mc;
  double acc[6] = {0., 0., 0., 0., 0., 0.};
  if constexpr (CNT > 0) {
    const int w = nb * CNT;
    for (int xx = l16; xx < w; xx += 16) {
      const int s = xx / CNT, r = xx - s * CNT;
      const int64_t nc = 6 * bcol[b0 + s];
      if constexpr (SAME) {
        int c0 = 0;
#pragma unroll
        for (int d = 5; d >= 0; --d)
          if (M.cnt[d] != 0) c0 = M.cols[d][r];
        double xv = x[nc + c0];
        if (free_ != nullptr && !free_[nc + c0]) xv = 0.;
#pragma unroll
        for (int d = 0; d < 6; ++d)
          if (M.cnt[d] != 0) acc[d] += v[M.rowoff[d] * nb + xx] * xv;
      } else {
#pragma unroll
        for (int d = 0; d < 6; ++d) {
          if (M.cnt[d] == 0) continue;
          const int64_t col = nc + M.cols[d][r];
          double xv = x[col];
          if (free_ != nullptr && !free_[col]) xv = 0.;
          acc[d] += v[M.rowoff[d] * nb + xx] * xv;
        }
      }
    }
  } else {
#pragma unroll
    for (int d = 0; d < 6; ++d) {
      const int cnt = M.cnt[d];
      if (cnt == 0) continue;
      const int w = nb * cnt;
      for (int xx = l16; xx < w; xx += 16) {
        const int s = xx / cnt, r = xx - s * cnt;
        const int64_t col = 6 * bcol[b0 + s] + M.cols[d][r];
        double xv = x[col];
        if (free_ != nullptr && !free_[col]) xv = 0.;
        acc[d] += v[M.rowoff[d] * nb + xx] * xv;
      }
    }
  }
#pragma unroll
  for (int d = 0; d < 6; ++d)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o);
  if (live && l16 < 6) {
    double out = acc[0];
#pragma unroll
    for (int d = 1; d < 6; ++d) out = (l16 == d) ? acc[d] : out;
    const int64_t row = 6 * i + l16;
    if (free_ != nullptr && !free_[6 * node_begin + row]) out = 0.;
    y[row] = out;
  }
}

// Full 6x6 blocks (KC0 of shells and beams): 16-byte loads of values and of x, two columns per lane.
__global__ void __launch_bounds__(128) k_plan_spmv_full(const int64_t* __restrict__ brow_ptr,
                                                        const int64_t* __restrict__ bcol, int64_t nown, int64_t node_begin,
                                                        const double* __restrict__ vals,
                                                        const unsigned char* __restrict__ free_,
                                                        const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31, h = lane >> 4, l16 = lane & 15;
  const int64_t i = (int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + h;
  const bool live = i < nown;
  int64_t b0 = 0;
  int nb = 0;
  if (live) {
    b0 = brow_ptr[i];
    nb = int(brow_ptr[i + 1] - b0);
  }
  const double* v = vals + b0 * 36;
  const int w = nb * 6;
  double acc[6] = {0., 0., 0., 0., 0., 0.};
  for (int xx = 2 * l16; xx < w; xx += 32) {
    const int s = xx / 6, r = xx - s * 6;
    const int64_t col = 6 * bcol[b0 + s] + r;
    double2 xv = *reinterpret_cast<const double2*>(x + col);
    if (free_ != nullptr) {
      const uchar2 f = *reinterpret_cast<const uchar2*>(free_ + col);
      if (!f.x) xv.x = 0.;
      if (!f.y) xv.y = 0.;
    }
#pragma unroll
    for (int d = 0; d < 6; ++d) {
      const double2 a = *reinterpret_cast<const double2*>(v + d * w + xx);
      acc[d] += a.x * xv.x;
      acc[d] += a.y * xv.y;
    }
  }
#pragma unroll
  for (int d = 0; d < 6; ++d)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o);
  if (live && l16 < 6) {
    double out = acc[0];
#pragma unroll
    for (int d = 1; d < 6; ++d) out = (l16 == d) ? acc[d] : out;
    const int64_t row = 6 * i + l16;
    if (free_ != nullptr && !free_[6 * node_begin + row]) out = 0.;
    y[row] = out;
  }
}

// diag[6 i + d] = A[row, row] (0 when the pattern has no diagonal entry there): Jacobi scaling
__global__ void k_plan_diag(const SpmvMask M, const int64_t* __restrict__ brow_ptr, const int64_t* __restrict__ bcol,
                            int64_t nown, int64_t node_begin, const double* __restrict__ vals,
                            double* __restrict__ diag) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < nown * 6; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t i = t / 6;
    const int d = int(t - i * 6);
    const int64_t b0 = brow_ptr[i];
    const int nb = int(brow_ptr[i + 1] - b0);
    double out = 0.;
    int r = -1;
    for (int q = 0; q < M.cnt[d]; ++q)
      if (M.cols[d][q] == d) r = q;
    if (r >= 0)
      for (int s = 0; s < nb; ++s)
        if (bcol[b0 + s] == node_begin + i) out = vals[b0 * M.mc + int64_t(M.rowoff[d]) * nb + s * M.cnt[d] + r];
    diag[t] = out;
  }
}

static SpmvMask spmv_mask_of(const pf3_plan* pl) {
  SpmvMask M;
  for (int i = 0; i < 6; ++i) {
    int c = 0;
    for (int j = 0; j < 6; ++j) {
      M.cols[i][j] = 0;
      if (pl->umask[i][j]) M.cols[i][c++] = int8_t(j);
    }
    M.cnt[i] = pl->dev.cnt[i];
    M.rowoff[i] = pl->dev.rowoff[i];
  }
  M.mc = pl->dev.mc;
  return M;
}

int plan_spmv(const pf3_plan* pl, cudaStream_t st, const double* vals, const unsigned char* free_, const double* x,
              double* y, int64_t* launches) {
  if (pl->generic) return PF3_E_UNSUPPORTED;
  if (pl->nown <= 0) return PF3_OK;
  const SpmvMask M = spmv_mask_of(pl);
  int cnt = 0;
  bool uniform = true;
  for (int d = 0; d < 6; ++d) {
    if (M.cnt[d] == 0) continue;
    if (cnt == 0) cnt = M.cnt[d];
    if (M.cnt[d] != cnt) uniform = false;
  }
  const int wpc = 4;
  const unsigned grid = unsigned((pl->nown + 2 * wpc - 1) / (2 * wpc));
  const int64_t nb0 = pl->dev.node_begin;
  bool same = uniform;
  for (int d = 0, d0 = -1; d < 6 && same; ++d) {
    if (M.cnt[d] == 0) continue;
    if (d0 < 0) d0 = d;
    for (int r = 0; r < cnt; ++r) same = same && M.cols[d][r] == M.cols[d0][r];
  }
#define PF3_SPMV_CASE(C, S) \
  k_plan_spmv<C, S><<<grid, wpc * 32, 0, st>>>(M, pl->d_brow_ptr, pl->d_bcol, pl->nown, nb0, vals, free_, x, y)
  const bool al16 = ((((uintptr_t)vals) | ((uintptr_t)x)) & 15) == 0 && (free_ == nullptr || (((uintptr_t)free_) & 1) == 0);
  if (uniform && cnt == 6 && same && M.mc == 36 && al16)
    k_plan_spmv_full<<<grid, wpc * 32, 0, st>>>(pl->d_brow_ptr, pl->d_bcol, pl->nown, nb0, vals, free_, x, y);
  else if (uniform && cnt == 6 && same) PF3_SPMV_CASE(6, true);
  else if (uniform && cnt == 3 && same) PF3_SPMV_CASE(3, true);
  else if (uniform && cnt == 6) PF3_SPMV_CASE(6, false);
  else if (uniform && cnt == 5) PF3_SPMV_CASE(5, false);
  else if (uniform && cnt == 3) PF3_SPMV_CASE(3, false);
  else PF3_SPMV_CASE(0, false);
#undef PF3_SPMV_CASE
  ++*launches;
  return int(cudaGetLastError());
}

int plan_diagonal(const pf3_plan* pl, cudaStream_t st, const double* vals, double* diag, int64_t* launches) {
  if (pl->generic) return PF3_E_UNSUPPORTED;
  if (pl->nown <= 0) return PF3_OK;
  k_plan_diag<<<grid_for(pl->nown * 6), 256, 0, st>>>(spmv_mask_of(pl), pl->d_brow_ptr, pl->d_bcol, pl->nown,
                                                      pl->dev.node_begin, vals, diag);
  ++*launches;
  return int(cudaGetLastError());
}

// deterministic fint gather: fint[6*node + d] += sum over incident (element, local node) of fe
__global__ void k_fint_gather(const uint32_t* __restrict__ sorted_inc, const int64_t* __restrict__ inc_ptr,
                              int64_t nnodes, const double* __restrict__ fe, double* __restrict__ fint) {
  const int64_t n = nnodes * 6;
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < n; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t node = t / 6;
    const int d = int(t - node * 6);
    double s = 0.;
    for (int64_t q = inc_ptr[node]; q < inc_ptr[node + 1]; ++q) s += fe[int64_t(sorted_inc[q]) * 6 + d];
    fint[t] += s;
  }
}
__global__ void k_simple_inc_keys(const int64_t* conn, int64_t n, int64_t* keys, uint32_t* vals) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < n; t += int64_t(gridDim.x) * blockDim.x) {
    keys[t] = conn[t];
    vals[t] = uint32_t(t);
  }
}

int fint_gather(cudaStream_t st, int64_t ne, int nn, int64_t nnodes, const int64_t* conn, const double* fe,
                double* fint, int64_t* launches) {
  const int64_t NI = ne * nn;
  if (NI >= (int64_t(1) << 32)) return PF3_E_CAPACITY;
  int64_t *keys = nullptr, *keys2 = nullptr, *ptr = nullptr;
  uint32_t *vals = nullptr, *vals2 = nullptr;
  int rc = 0;
  auto cleanup = [&]() { cudaFree(keys); cudaFree(keys2); cudaFree(ptr); cudaFree(vals); cudaFree(vals2); };
#define PF3_TRY(x) do { rc = (x); if (rc) { cleanup(); return rc; } } while (0)
  PF3_TRY(dalloc(&keys, NI));
  PF3_TRY(dalloc(&keys2, NI));
  PF3_TRY(dalloc(&vals, NI));
  PF3_TRY(dalloc(&vals2, NI));
  PF3_TRY(dalloc(&ptr, nnodes + 1));
  k_simple_inc_keys<<<grid_for(NI), 256, 0, st>>>(conn, NI, keys, vals);
  PF3_TRY(sort_pairs(keys, keys2, vals, vals2, NI, bits_for(nnodes), st));
  k_row_ptr<<<grid_for(nnodes + 1), 256, 0, st>>>(keys2, NI, 0, nnodes, 1, ptr);
  // fe is [ne][nn][6]: incidence id t = e*nn + a addresses fe[t*6 .. t*6+5]
  k_fint_gather<<<grid_for(nnodes * 6), 256, 0, st>>>(vals2, ptr, nnodes, fe, fint);
  *launches += 3;
  PF3_TRY(int(cudaStreamSynchronize(st)));
  PF3_TRY(int(cudaGetLastError()));
  cleanup();
#undef PF3_TRY
  return PF3_OK;
}

int fused_max_slots();

namespace {
__global__ void k_max_valence(const int64_t* inc_ptr, int64_t nown, int* out) {
  for (int64_t n = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; n < nown; n += int64_t(gridDim.x) * blockDim.x)
    atomicMax(out, int(inc_ptr[n + 1] - inc_ptr[n]));
}
// L2 prefetch table (common.cuh: kPfChunk): first the lowest node pair that uses each element, then per chunk of pairs
// and quarter of the element range the [min, max] of the elements used first there.
template <class REC, int NINC, int DIV>
__global__ void k_pf_first_use(const REC* __restrict__ rec, int64_t total, int rmax, int* __restrict__ first_use) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int pair = int((t / rmax) >> 1);
#pragma unroll
    for (int k = 0; k < NINC; ++k) {
      const int inc = rec[t].inc[k];
      if (inc >= 0) atomicMin(first_use + inc / DIV, pair);
    }
  }
}
__global__ void k_pf_ranges(const int* __restrict__ first_use, int64_t ne, int64_t seglen, int* __restrict__ lo,
                            int* __restrict__ hi) {
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < ne; e += int64_t(gridDim.x) * blockDim.x) {
    const int f = first_use[e];
    if (f == 0x7f7f7f7f) continue;
    const int64_t slot = int64_t(f / kPfChunk) * kPfRuns + min(int64_t(kPfRuns - 1), e / seglen);
    atomicMin(lo + slot, int(e));
    atomicMax(hi + slot, int(e));
  }
}
__global__ void k_pf_table(const int* __restrict__ lo, const int* __restrict__ hi, int64_t nslots, int2* __restrict__ tab) {
  for (int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; c < nslots; c += int64_t(gridDim.x) * blockDim.x) {
    int2 t = make_int2(0, 0);
    // a chunk of kPfChunk node pairs first-uses about 2 kPfChunk elements; a run much longer than that means the
    // numbering scatters them (prefetching the hull would read far more than is used): no prefetch of this run
    if (hi[c] >= lo[c] && hi[c] - lo[c] + 1 <= 8 * kPfChunk) t = make_int2(lo[c], hi[c] - lo[c] + 1);
    tab[c] = t;
  }
}
}  // namespace

// Build the prefetch table of a plan from its node records (quads: NodeRec, 4 incidences, pair id e*16+..; triangles:
// TriaRec, 10 incidences, e*9+..).
template <class REC, int NINC, int DIV>
static int build_pf_table(const pf3_plan* pl, const REC* rec, int64_t ne, cudaStream_t st, int64_t* launches) {
  const int64_t npairs = (pl->nown + 1) / 2;
  const int64_t nchunks = (npairs + kPfChunk - 1) / kPfChunk;
  if (nchunks <= 4 * kPfAhead || npairs >= (int64_t(1) << 31) || ne <= 0) return PF3_OK;   // too small to matter
  const int64_t nslots = nchunks * kPfRuns;
  int *d_first = nullptr, *d_lo = nullptr, *d_hi = nullptr;
  PF3_CUDA(cudaMalloc((void**)&d_first, size_t(ne) * sizeof(int)));
  PF3_CUDA(cudaMalloc((void**)&d_lo, size_t(nslots) * sizeof(int)));
  PF3_CUDA(cudaMalloc((void**)&d_hi, size_t(nslots) * sizeof(int)));
  PF3_CUDA(cudaMalloc((void**)&pl->d_pftab, size_t(nslots) * sizeof(int2)));
  PF3_CUDA(cudaMemsetAsync(d_first, 0x7f, size_t(ne) * sizeof(int), st));   // 0x7f7f7f7f: "never" (> any pair)
  PF3_CUDA(cudaMemsetAsync(d_lo, 0x7f, size_t(nslots) * sizeof(int), st));
  PF3_CUDA(cudaMemsetAsync(d_hi, 0xff, size_t(nslots) * sizeof(int), st));    // -1
  k_pf_first_use<REC, NINC, DIV><<<grid_for(pl->nown * pl->rmax), 256, 0, st>>>(rec, pl->nown * pl->rmax, pl->rmax, d_first);
  k_pf_ranges<<<grid_for(ne), 256, 0, st>>>(d_first, ne, (ne + kPfRuns - 1) / kPfRuns, d_lo, d_hi);
  k_pf_table<<<grid_for(nslots), 256, 0, st>>>(d_lo, d_hi, nslots, pl->d_pftab);
  *launches += 3;
  PF3_CUDA(cudaGetLastError());
  PF3_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_first);
  cudaFree(d_lo);
  cudaFree(d_hi);
  pl->pf_nchunks = nchunks;
  return PF3_OK;
}

namespace {
__global__ void k_node_records(const int64_t* __restrict__ brow_ptr, const int64_t* __restrict__ inc_ptr,
                               const int64_t* __restrict__ inc_pair0, const int32_t* __restrict__ slot, int64_t nown,
                               int rmax, NodeRec* __restrict__ out) {
  const int64_t total = nown * rmax;
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t n = t / rmax;
    const int r = int(t - n * rmax);
    NodeRec R;
    R.b0 = brow_ptr[n];
    R.nb = uint8_t(brow_ptr[n + 1] - R.b0);
    const int64_t q0 = inc_ptr[n];
    const int v = int(inc_ptr[n + 1] - q0);
    for (int s = 0; s < 16; ++s) R.gmap[s] = 0xFFFF;
    int cnt = 0;
    for (int k = 0; k < 4; ++k) {
      const int kk = 4 * r + k;
      R.inc[k] = -1;
      if (kk < v) {
        const int64_t p0 = inc_pair0[q0 + kk];
        R.inc[k] = int32_t(p0);
        ++cnt;
        for (int b = 0; b < 4; ++b) {
          const int s = slot[p0 + b];
          if (s >= 0 && s < 16) R.gmap[s] = uint16_t((R.gmap[s] & ~(0xF << (4 * k))) | (b << (4 * k)));
        }
      }
    }
    R.v = uint8_t(cnt);
    for (int i = 0; i < 6; ++i) R.pad[i] = 0;
    out[t] = R;
  }
}
}  // namespace

namespace {
__global__ void k_tria_records(const int64_t* __restrict__ brow_ptr, const int64_t* __restrict__ inc_ptr,
                               const int64_t* __restrict__ inc_pair0, const int32_t* __restrict__ slot, int64_t nown,
                               int rmax, TriaRec* __restrict__ out) {
  const int64_t total = nown * rmax;
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t n = t / rmax;
    const int r = int(t - n * rmax);
    TriaRec R;
    R.b0 = brow_ptr[n];
    R.nb = uint8_t(brow_ptr[n + 1] - R.b0);
    const int64_t q0 = inc_ptr[n];
    const int v = int(inc_ptr[n + 1] - q0);
    for (int s = 0; s < 16; ++s) R.gmap[s] = 0xFFFFFFFFu;
    int cnt = 0;
    for (int k = 0; k < 10; ++k) {
      const int kk = 10 * r + k;
      R.inc[k] = -1;
      if (kk < v) {
        const int64_t p0 = inc_pair0[q0 + kk];
        R.inc[k] = int32_t(p0);
        ++cnt;
        for (int b = 0; b < 3; ++b) {
          const int s = slot[p0 + b];
          if (s >= 0 && s < 16) R.gmap[s] = (R.gmap[s] & ~(3u << (2 * k))) | (uint32_t(b) << (2 * k));
        }
      }
    }
    R.v = uint8_t(cnt);
    for (int i = 0; i < 14; ++i) R.pad[i] = 0;
    out[t] = R;
  }
}
}  // namespace

namespace {
// fused-kernel node records of ONE group of a multi-group plan: incidences filtered by group, pair ids local to the
// group (e*16 + a*4), slots relative to the union row block
__global__ void k_node_records_group(const PlanDev P, int gi, const int64_t* __restrict__ brow_ptr,
                                     const int64_t* __restrict__ inc_ptr, const int64_t* __restrict__ inc_pair0,
                                     const int32_t* __restrict__ inc_meta, const int32_t* __restrict__ slot,
                                     int64_t nown, int rmax, NodeRec* __restrict__ out) {
  const GroupDev& G = P.g[gi];
  const int64_t total = nown * rmax;
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t n = t / rmax;
    const int r = int(t - n * rmax);
    NodeRec R;
    R.b0 = brow_ptr[n];
    R.nb = uint8_t(brow_ptr[n + 1] - R.b0);
    for (int s = 0; s < 16; ++s) R.gmap[s] = 0xFFFF;
    for (int k = 0; k < 4; ++k) R.inc[k] = -1;
    int seen = 0, cnt = 0;
    for (int64_t q = inc_ptr[n]; q < inc_ptr[n + 1]; ++q) {
      if ((inc_meta[q] & 0xff) != gi) continue;
      const int k = seen - 4 * r;
      ++seen;
      if (k < 0 || k >= 4) continue;
      const int64_t p0 = inc_pair0[q];
      R.inc[k] = int32_t(p0 - G.pairbase);
      ++cnt;
      for (int b = 0; b < 4; ++b) {
        const int sl = slot[p0 + b];
        if (sl >= 0 && sl < 16) R.gmap[sl] = uint16_t((R.gmap[sl] & ~(0xF << (4 * k))) | (b << (4 * k)));
      }
    }
    R.v = uint8_t(cnt);
    for (int i = 0; i < 6; ++i) R.pad[i] = 0;
    out[t] = R;
  }
}
}  // namespace

// Union layout of `matrix` (with mass type mtype for every group) over the plan's groups, seen from group `group`.
int plan_union_map(const pf3_plan* pl, int group, int matrix, int mtype, UnionMap* um) {
  if (pl->generic || group < 0 || group >= pl->dev.ngroups) return PF3_E_BAD_ARG;
  bool u[6][6], own[6][6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) u[i][j] = own[i][j] = false;
  for (int g = 0; g < pl->dev.ngroups; ++g) {
    const BlockLayout L = make_layout(pl->gkind[g], matrix, mtype);
    if (L.written == 0) return PF3_E_UNSUPPORTED;
    // a group that contributes only diagonal node pairs to this matrix (lumped beam / truss mass) gives the matrix a
    // block structure that may differ from this plan's: the fused kernel, which walks this plan's blocks, cannot write it
    if (L.diag_pairs && !pl->dev.g[g].diag) return PF3_E_UNSUPPORTED;
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        u[i][j] = u[i][j] || mask_has(L.mask, i, j);
        if (g == group) own[i][j] = mask_has(L.mask, i, j);
      }
  }
  bool same = true;
  int mc = 0, orow = 0;
  for (int i = 0; i < 6; ++i) {
    um->rowoff[i] = mc;
    int c = 0, oc = 0, ownc = 0;
    for (int j = 0; j < 6; ++j) ownc += own[i][j] ? 1 : 0;
    for (int j = 0; j < 6; ++j) {
      um->col[i][j] = -1;
      if (u[i][j]) {
        um->col[i][c] = own[i][j] ? int8_t(oc) : int8_t(-1);
        ++c;
      }
      if (own[i][j]) ++oc;
      same = same && (u[i][j] == own[i][j]);
    }
    um->row[i] = ownc > 0 ? int8_t(orow++) : int8_t(-1);
    um->cnt[i] = c;
    mc += c;
    um->colbits[i] = 0;
    for (int j = 0; j < 6; ++j) um->colbits[i] |= unsigned(um->col[i][j] < 0 ? 0xF : um->col[i][j]) << (4 * j);
  }
  um->mc = mc;
  um->active = same ? 0 : 1;
  return PF3_OK;
}

// FusedArgs for group `group` (a Quad4 / Quad4R group) of a multi-group plan.
int plan_fused_args_group(const pf3_plan* pl, int group, FusedArgs* F, cudaStream_t st, int64_t* launches) {
  if (pl->generic || pl->degenerate || group < 0 || group >= pl->dev.ngroups) return PF3_E_UNSUPPORTED;
  const GroupDev& G = pl->dev.g[group];
  if (G.nn != 4 || G.diag || G.npairs != 16) return PF3_E_UNSUPPORTED;
  if (pl->max_nb > fused_max_slots()) return PF3_E_CAPACITY;
  if (G.ne * 16 >= (int64_t(1) << 31)) return PF3_E_CAPACITY;
  if (pl->d_frecs == nullptr || pl->frecs_group != group) {
    cudaFree(pl->d_frecs);
    pl->d_frecs = nullptr;
    int* d_mv = nullptr;
    PF3_CUDA(cudaMalloc((void**)&d_mv, sizeof(int)));
    PF3_CUDA(cudaMemsetAsync(d_mv, 0, sizeof(int), st));
    k_group_valence<<<grid_for(pl->nown), 256, 0, st>>>(pl->dev, group, pl->d_inc_ptr, pl->d_inc_meta, pl->nown, d_mv);
    int mv = 0;
    PF3_CUDA(cudaMemcpyAsync(&mv, d_mv, sizeof(int), cudaMemcpyDeviceToHost, st));
    PF3_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_mv);
    pl->frmax = std::max(1, (mv + 3) / 4);
    PF3_CUDA(cudaMalloc((void**)&pl->d_frecs, size_t(pl->nown) * pl->frmax * sizeof(NodeRec)));
    k_node_records_group<<<grid_for(pl->nown * pl->frmax), 256, 0, st>>>(pl->dev, group, pl->d_brow_ptr, pl->d_inc_ptr,
                                                                        pl->d_inc_pair0, pl->d_inc_meta, pl->d_slot,
                                                                        pl->nown, pl->frmax, pl->d_frecs);
    pl->frecs_group = group;
    *launches += 2;
    PF3_CUDA(cudaGetLastError());
  }
  F->noderec = pl->d_frecs;
  F->rmax = pl->frmax;
  F->brow_ptr = pl->d_brow_ptr;
  F->inc_ptr = pl->d_inc_ptr;
  F->inc_pair0 = pl->d_inc_pair0;
  F->slot = pl->d_slot;
  F->nown = pl->nown;
  return PF3_OK;
}

int tria_fused_max_slots();
int tria_fused_incidences();
// Triangle twin of plan_fused_args.
int plan_fused_args_tria(const pf3_plan* pl, FusedArgs* F, cudaStream_t st, int64_t* launches) {
  if (pl->generic || pl->dev.ngroups != 1 || pl->degenerate) return PF3_E_UNSUPPORTED;
  const GroupDev& G = pl->dev.g[0];
  if (G.nn != 3 || G.diag || G.npairs != 9 || G.pairbase != 0) return PF3_E_UNSUPPORTED;
  if (pl->max_nb > tria_fused_max_slots()) return PF3_E_CAPACITY;
  if (G.ne * 9 >= (int64_t(1) << 31)) return PF3_E_CAPACITY;
  if (pl->d_triarec == nullptr) {
    int* d_mv = nullptr;
    PF3_CUDA(cudaMalloc((void**)&d_mv, sizeof(int)));
    PF3_CUDA(cudaMemsetAsync(d_mv, 0, sizeof(int), st));
    k_max_valence<<<grid_for(pl->nown), 256, 0, st>>>(pl->d_inc_ptr, pl->nown, d_mv);
    int mv = 0;
    PF3_CUDA(cudaMemcpyAsync(&mv, d_mv, sizeof(int), cudaMemcpyDeviceToHost, st));
    PF3_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_mv);
    const int per = tria_fused_incidences();
    pl->rmax = mv <= per ? 1 : (mv + per - 1) / per;
    PF3_CUDA(cudaMalloc((void**)&pl->d_triarec, size_t(pl->nown) * pl->rmax * sizeof(TriaRec)));
    k_tria_records<<<grid_for(pl->nown * pl->rmax), 256, 0, st>>>(pl->d_brow_ptr, pl->d_inc_ptr, pl->d_inc_pair0,
                                                                  pl->d_slot, pl->nown, pl->rmax, pl->d_triarec);
    *launches += 2;
    PF3_CUDA(cudaGetLastError());
    int rc = build_pf_table<TriaRec, 10, 9>(pl, pl->d_triarec, G.ne, st, launches);
    if (rc) return rc;
  }
  F->triarec = pl->d_triarec;
  F->pftab = pl->d_pftab;
  F->pf_nchunks = pl->pf_nchunks;
  F->rmax = pl->rmax;
  F->brow_ptr = pl->d_brow_ptr;
  F->inc_ptr = pl->d_inc_ptr;
  F->inc_pair0 = pl->d_inc_pair0;
  F->slot = pl->d_slot;
  F->nown = pl->nown;
  return PF3_OK;
}

// Fill the plan part of FusedArgs; PF3_E_UNSUPPORTED when the plan cannot drive the fused kernel.
int plan_fused_args(const pf3_plan* pl, int kind, FusedArgs* F, cudaStream_t st, int64_t* launches) {
  if (pl->generic || pl->dev.ngroups != 1 || pl->degenerate) return PF3_E_UNSUPPORTED;
  const GroupDev& G = pl->dev.g[0];
  if (G.nn != 4 || G.diag || G.npairs != 16 || G.pairbase != 0) return PF3_E_UNSUPPORTED;
  if (kind != PF3_QUAD4 && kind != PF3_QUAD4R) return PF3_E_UNSUPPORTED;
  if (pl->max_nb > fused_max_slots()) return PF3_E_CAPACITY;
  if (G.ne * 16 >= (int64_t(1) << 31)) return PF3_E_CAPACITY;
  if (pl->d_noderec == nullptr) {
    int* d_mv = nullptr;
    PF3_CUDA(cudaMalloc((void**)&d_mv, sizeof(int)));
    PF3_CUDA(cudaMemsetAsync(d_mv, 0, sizeof(int), st));
    k_max_valence<<<grid_for(pl->nown), 256, 0, st>>>(pl->d_inc_ptr, pl->nown, d_mv);
    int mv = 0;
    PF3_CUDA(cudaMemcpyAsync(&mv, d_mv, sizeof(int), cudaMemcpyDeviceToHost, st));
    PF3_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_mv);
    pl->rmax = mv <= 4 ? 1 : (mv + 3) / 4;
    PF3_CUDA(cudaMalloc((void**)&pl->d_noderec, size_t(pl->nown) * pl->rmax * sizeof(NodeRec)));
    k_node_records<<<grid_for(pl->nown * pl->rmax), 256, 0, st>>>(pl->d_brow_ptr, pl->d_inc_ptr, pl->d_inc_pair0,
                                                                  pl->d_slot, pl->nown, pl->rmax, pl->d_noderec);
    *launches += 2;
    PF3_CUDA(cudaGetLastError());
    int rc = build_pf_table<NodeRec, 4, 16>(pl, pl->d_noderec, G.ne, st, launches);
    if (rc) return rc;
  }
  F->noderec = pl->d_noderec;
  F->pftab = pl->d_pftab;
  F->pf_nchunks = pl->pf_nchunks;
  F->rmax = pl->rmax;
  F->brow_ptr = pl->d_brow_ptr;
  F->inc_ptr = pl->d_inc_ptr;
  F->inc_pair0 = pl->d_inc_pair0;
  F->slot = pl->d_slot;
  F->nown = pl->nown;
  return PF3_OK;
}
// fint[6*node + d] += sum over the node's incident (element, local node) pairs of group `group` of fe, in the
// plan's fixed incidence order: no sort, no atomics (update_fint, e.g. quad4.pyx:1339-1362).
namespace {
__global__ void k_plan_fint(const PlanDev P, int gi, const int64_t* __restrict__ inc_ptr,
                            const int64_t* __restrict__ inc_pair0, const int32_t* __restrict__ inc_meta, int64_t nown,
                            const double* __restrict__ fe, double* __restrict__ fint) {
  const GroupDev& G = P.g[gi];
  const int64_t n6 = nown * 6;
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < n6; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t i = t / 6;
    const int d = int(t - i * 6);
    double s = 0.;
    bool any = false;
    for (int64_t q = inc_ptr[i]; q < inc_ptr[i + 1]; ++q) {
      const int meta = inc_meta[q];
      if ((meta & 0xff) != gi) continue;
      const int64_t e = (inc_pair0[q] - G.pairbase) / G.npairs;
      s += fe[(e * G.nn + (meta >> 8)) * 6 + d];
      any = true;
    }
    if (any) fint[6 * (P.node_begin + i) + d] += s;
  }
}

// The same gather with one thread per NODE: the incidence list is walked once instead of once per dof, the element index
// comes from a division by the compile-time NN * NN, and the 48-byte node rows of fe / fint move as three 16-byte
// pieces.  Same summation order (incidences ascending), so the sums are bit-identical to k_plan_fint's.
template <int NN>
__global__ void k_plan_fint_rows(const PlanDev P, int gi, const int64_t* __restrict__ inc_ptr,
                                 const int64_t* __restrict__ inc_pair0, const int32_t* __restrict__ inc_meta,
                                 int64_t nown, const double* __restrict__ fe, double* __restrict__ fint) {
  const int64_t pairbase = P.g[gi].pairbase;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < nown; i += int64_t(gridDim.x) * blockDim.x) {
    double2 s0 = make_double2(0., 0.), s1 = s0, s2 = s0;
    bool any = false;
    const int64_t q1 = inc_ptr[i + 1];
    for (int64_t q = inc_ptr[i]; q < q1; ++q) {
      const int meta = inc_meta[q];
      if ((meta & 0xff) != gi) continue;
      const uint64_t e = uint64_t(inc_pair0[q] - pairbase) / unsigned(NN * NN);
      const double2* r = reinterpret_cast<const double2*>(fe + (e * NN + unsigned(meta >> 8)) * 6);
      const double2 t0 = r[0], t1 = r[1], t2 = r[2];
      s0.x += t0.x; s0.y += t0.y;
      s1.x += t1.x; s1.y += t1.y;
      s2.x += t2.x; s2.y += t2.y;
      any = true;
    }
    if (any) {
      double2* o = reinterpret_cast<double2*>(fint + 6 * (P.node_begin + i));
      double2 a0 = o[0], a1 = o[1], a2 = o[2];
      a0.x += s0.x; a0.y += s0.y;
      a1.x += s1.x; a1.y += s1.y;
      a2.x += s2.x; a2.y += s2.y;
      o[0] = a0; o[1] = a1; o[2] = a2;
    }
  }
}
}  // namespace

int plan_fint_gather(const pf3_plan* pl, cudaStream_t st, int group, const double* fe, double* fint,
                     int64_t* launches) {
  if (pl->generic || group < 0 || group >= pl->dev.ngroups) return PF3_E_BAD_ARG;
  const int nn = pl->dev.g[group].nn;
#ifdef PF3_FINT_GATHER_OLD
  const bool rows = false;
#else
  const bool rows = ((reinterpret_cast<uintptr_t>(fe) | reinterpret_cast<uintptr_t>(fint)) & 15) == 0 &&
                    pl->dev.g[group].npairs == nn * nn && nn >= 2 && nn <= 4;
#endif
#define PF3_FINT_ROWS(NN) \
  k_plan_fint_rows<NN><<<grid_for(pl->nown), 128, 0, st>>>(pl->dev, group, pl->d_inc_ptr, pl->d_inc_pair0, pl->d_inc_meta, \
                                                          pl->nown, fe, fint)
  if (rows && nn == 4) PF3_FINT_ROWS(4);
  else if (rows && nn == 3) PF3_FINT_ROWS(3);
  else if (rows && nn == 2) PF3_FINT_ROWS(2);
  else
    k_plan_fint<<<grid_for(pl->nown * 6), 256, 0, st>>>(pl->dev, group, pl->d_inc_ptr, pl->d_inc_pair0, pl->d_inc_meta,
                                                       pl->nown, fe, fint);
#undef PF3_FINT_ROWS
  ++*launches;
  return int(cudaGetLastError());
}
int plan_group_kind_nn(const pf3_plan* pl, int group) {
  return (pl->generic || group < 0 || group >= pl->dev.ngroups) ? 0 : pl->dev.g[group].nn;
}
int64_t plan_group_ne_of(const pf3_plan* pl, int group) {
  return (pl->generic || group < 0 || group >= pl->dev.ngroups) ? -1 : pl->dev.g[group].ne;
}
int64_t plan_nblocks(const pf3_plan* pl) { return pl->generic ? 0 : pl->nblk; }
int64_t plan_group_ne(const pf3_plan* pl) { return pl->generic ? 0 : pl->dev.g[0].ne; }
int64_t plan_nnz(const pf3_plan* pl) { return pl->nnz; }
int64_t plan_nrows(const pf3_plan* pl) { return pl->nrows; }

}  // namespace pf3

extern "C" int pf3_plan_destroy(pf3_plan* pl) {
  if (!pl) return PF3_OK;
  cudaFree(pl->d_brow_ptr); cudaFree(pl->d_bcol); cudaFree(pl->d_inc_ptr); cudaFree(pl->d_inc_src);
  cudaFree(pl->d_inc_pair0); cudaFree(pl->d_inc_meta); cudaFree(pl->d_slot); cudaFree(pl->d_noderec); cudaFree(pl->d_pftab); cudaFree(pl->d_triarec); cudaFree(pl->d_frecs);
  for (auto* r : pl->d_grecs) cudaFree(r);
  for (auto* t : pl->d_tabs) cudaFree(t);
  cudaFree(pl->d_indptr); cudaFree(pl->d_indices); cudaFree(pl->d_perm); cudaFree(pl->d_seg);
  delete pl;
  return PF3_OK;
}
