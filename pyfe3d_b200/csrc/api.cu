// extern "C" surface of libpyfe3d_b200.so (declared in include/pyfe3d_b200.h).
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

#include "common.cuh"
#include "pattern.hpp"

struct pf3_plan;

namespace pf3 {
cudaError_t launch_quad(int kind, const EvalArgs& A, cudaStream_t st);
cudaError_t launch_tria(const EvalArgs& A, cudaStream_t st);
cudaError_t launch_quad_aero(const EvalArgs& A, const AeroOut& O, cudaStream_t st);
cudaError_t launch_laminate_props(int64_t nrows, int nplies, const double* theta, int64_t theta_stride,
                                  const double* plyt, int64_t plyt_stride, const double* lamina, int64_t lamina_stride,
                                  const double* offset, int64_t offset_stride, int calc_scf, double* out, cudaStream_t st);
cudaError_t launch_lp_props(int64_t nrows, const double* thick, int64_t thick_stride, const double* inv,
                            int64_t inv_stride, const double* lp, int64_t lp_stride, const double* rho,
                            int64_t rho_stride, int var_mask, int grad_complete, double* out, double* grad,
                            cudaStream_t st);
cudaError_t launch_quad4_BL(int64_t n, const double* xe, double xi, double eta, double* out, cudaStream_t st);
cudaError_t launch_line(int kind, const EvalArgs& A, cudaStream_t st);
int plan_create_structured(int device, cudaStream_t st, int matrix, int64_t nnodes, int ngroups,
                           const pf3_batch* groups, const int64_t* coo_offsets, int64_t node_begin,
                           int64_t node_end, int64_t* launches, pf3_plan** out);
int plan_create_generic(int device, cudaStream_t st, int64_t n, int64_t nnz_coo, const int64_t* r, const int64_t* c,
                        int64_t* launches, pf3_plan** out);
int plan_pattern(const pf3_plan* pl, cudaStream_t st, int64_t* indptr, int64_t* indices, int64_t* launches);
int plan_assemble(const pf3_plan* pl, cudaStream_t st, const double* coo_v, double* csr_v, int64_t* launches);
int spmv_csr(cudaStream_t st, int64_t nrows, const int64_t* indptr, const int64_t* indices, const double* vals,
             const double* x, double* y, int64_t* launches);
int spmv_csr_masked(cudaStream_t st, int64_t nrows, const int64_t* indptr, const int64_t* indices, const double* vals,
                    const unsigned char* free_, const double* x, double* y, int64_t* launches);
int csr_diagonal(cudaStream_t st, int64_t nrows, const int64_t* indptr, const int64_t* indices, const double* vals,
                 int64_t row0, double* d, int64_t* launches);
int plan_spmv(const pf3_plan* pl, cudaStream_t st, const double* vals, const unsigned char* free_, const double* x,
              double* y, int64_t* launches);
int plan_diagonal(const pf3_plan* pl, cudaStream_t st, const double* vals, double* diag, int64_t* launches);
int fint_gather(cudaStream_t st, int64_t ne, int nn, int64_t nnodes, const int64_t* conn, const double* fe,
                double* fint, int64_t* launches);
int64_t plan_nnz(const pf3_plan* pl);
int64_t plan_nblocks(const pf3_plan* pl);
size_t cg_shard_work_bytes();
int cg_shard_dot(cudaStream_t st, int64_t n, const double* p, const double* ap, double* sc, void* work, int64_t* launches);
int cg_shard_update(cudaStream_t st, int64_t n, const double* p, const double* ap, const double* minv, double* x, double* r,
                    double* sc, void* work, int64_t* launches);
int cg_shard_dir(cudaStream_t st, int64_t n, const double* r, const double* minv, double* p, double* sc, void* work,
                 int64_t* launches);
int plan_fint_gather(const pf3_plan* pl, cudaStream_t st, int group, const double* fe, double* fint, int64_t* launches);
int plan_group_kind_nn(const pf3_plan* pl, int group);
int64_t plan_group_ne_of(const pf3_plan* pl, int group);
int64_t plan_group_ne(const pf3_plan* pl);
int plan_fused_args(const pf3_plan* pl, int kind, FusedArgs* F, cudaStream_t st, int64_t* launches);
cudaError_t launch_quad_fused(int kind, const FusedArgs& F, double* rec, cudaStream_t st, int64_t* launches,
                              int phases);
int fused_record_stride(int kind, const EvalArgs& A);
int plan_fused_args_tria(const pf3_plan* pl, FusedArgs* F, cudaStream_t st, int64_t* launches);
int plan_fused_args_group(const pf3_plan* pl, int group, FusedArgs* F, cudaStream_t st, int64_t* launches);
int plan_union_map(const pf3_plan* pl, int group, int matrix, int mtype, UnionMap* um);
int plan_assemble_gather(const pf3_plan* pl, cudaStream_t st, const double* coo_v, double* csr_v, int skip_group,
                         int64_t* launches);
cudaError_t launch_tria_fused(const FusedArgs& F, double* rec, cudaStream_t st, int64_t* launches);
int tria_fused_record_stride(const EvalArgs& A);
int64_t plan_nrows(const pf3_plan* pl);
struct CgOp {
  const pf3_plan* plan;
  const double* vals;
  double coef;
};
size_t cg_work_bytes(int64_t n);
int plan_cg(cudaStream_t st, int nops, const CgOp* ops, int64_t n, const unsigned char* free_, const double* b,
            double* x, int use_x0, double rtol, double atol, int maxiter, int flags, void* work, int* iters,
            int* status, double* resid, double* bnorm, int64_t* launches);
int plan_spmv_scaled(const pf3_plan* pl, cudaStream_t st, const double* vals, const unsigned char* free_,
                     const double* scale, const double* x, double* y, double* tmp, int64_t n, int64_t* launches);
int csr_compact_symbolic(cudaStream_t st, int64_t nrows, int64_t ncols, const int64_t* indptr, const int64_t* indices,
                         const unsigned char* free_, int upper, int64_t row0, int64_t* colmap, int64_t* out_ptr,
                         int64_t* nkeep, int64_t* nnz, int64_t* launches);
int csr_compact_fill(cudaStream_t st, int64_t nrows, int64_t ncols, const int64_t* indptr, const int64_t* indices,
                     const double* vals, const unsigned char* free_, int upper, int64_t row0, const int64_t* colmap,
                     const int64_t* out_ptr, int64_t* out_idx, double* out_val, int64_t* launches);
}  // namespace pf3

#define PF3_HOST_CHUNKS 8   // row ranges of the pipelined host-buffer step

struct pf3_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int64_t launches = 0;
  std::map<int, int8_t*> idx_tabs;  // (kind, matrix, mtype) -> device table [4][written]
  double* scratch = nullptr;
  size_t scratch_bytes = 0;
  double* hostio = nullptr;          // device staging of pf3_eval_assemble_host: x | u | csr_kc0 | csr_kg | csr_m
  size_t hostio_bytes = 0;
  cudaStream_t copy_stream = nullptr;   // pf3_eval_assemble_host: device->host copies of finished row ranges
  cudaEvent_t chunk_done[PF3_HOST_CHUNKS] = {};
  char* stage_host = nullptr;           // pinned + device staging of the small host-pointer calls (per-element drop-in)
  char* stage_dev = nullptr;
  void* solve_work = nullptr;           // vectors + scalars of pf3_plan_cg / pf3_plan_spmv_scaled
  size_t solve_work_bytes = 0;
};
#define PF3_STAGE_BYTES (size_t(1) << 20)

namespace {

#define PF3_CUDA(x)                        \
  do {                                     \
    cudaError_t _e = (x);                  \
    if (_e != cudaSuccess) return int(_e); \
  } while (0)

int use_device(pf3_context* ctx) {
  if (!ctx) return PF3_E_BAD_ARG;
  PF3_CUDA(cudaSetDevice(ctx->device));
  // The runtime keeps ONE last-error slot per host thread, shared with every other CUDA user in the process (torch,
  // cub, ...): a non-sticky error left there by someone else's call must not surface as the status of this call, whose
  // launches are checked with cudaGetLastError().  Sticky (fatal) errors are returned again by the next call anyway.
  const cudaError_t stale = cudaGetLastError();
  if (stale != cudaSuccess && std::getenv("PF3_DEBUG"))
    std::fprintf(stderr, "pyfe3d_b200: cleared a stale CUDA error left by an earlier call: %d (%s)\n", int(stale),
                 cudaGetErrorString(stale));
  return PF3_OK;
}

int ensure_scratch(pf3_context* ctx, size_t bytes) {
  if (bytes <= ctx->scratch_bytes) return PF3_OK;
  if (ctx->scratch) cudaFree(ctx->scratch);
  ctx->scratch = nullptr;
  ctx->scratch_bytes = 0;
  PF3_CUDA(cudaMalloc((void**)&ctx->scratch, bytes));
  ctx->scratch_bytes = bytes;
  return PF3_OK;
}

// One-shot grid: CTA b fills the 2048 consecutive written entries [2048 b, 2048 b + 2048) of the index arrays, so the
// hardware hands out the output in address order (a capped grid-stride grid loses ~25 % of the write bandwidth to the
// widening front of addresses in flight) and the 64-bit division is done once per CTA, not per entry.
constexpr int kIdxPerCta = 2048;
// VEC: two consecutive entries per thread as one 16-byte store (written and size even, arrays 16-byte aligned: a pair
// never straddles two elements).
template <bool VEC>
__global__ void __launch_bounds__(256) k_fill_indices(const int8_t* __restrict__ tab, int written, int size, int nn,
                                                      int64_t ne, const int64_t* __restrict__ conn,
                                                      int64_t* __restrict__ r, int64_t* __restrict__ c) {
  const int64_t n = ne * written;
  const int64_t t0 = int64_t(blockIdx.x) * kIdxPerCta;
  const int64_t e0 = t0 / written;
  const int l0 = int(t0 - e0 * written);
  constexpr int kStep = VEC ? 2 : 1;
#pragma unroll
  for (int i = 0; i < kIdxPerCta / (256 * kStep); ++i) {
    const int off = (i * 256 + int(threadIdx.x)) * kStep;
    if (t0 + off >= n) return;
    const unsigned lo = unsigned(l0 + off);
    const unsigned de = lo / unsigned(written);
    const int l = int(lo - de * unsigned(written));
    const int64_t e = e0 + de;
    const int64_t k = e * size + l;
    const int64_t* ce = conn + e * nn;
    if constexpr (VEC) {
      if (r) {
        const longlong2 v = make_longlong2(6 * ce[tab[l]] + tab[written + l], 6 * ce[tab[l + 1]] + tab[written + l + 1]);
        *reinterpret_cast<longlong2*>(r + k) = v;
      }
      if (c) {
        const longlong2 v = make_longlong2(6 * ce[tab[2 * written + l]] + tab[3 * written + l],
                                           6 * ce[tab[2 * written + l + 1]] + tab[3 * written + l + 1]);
        *reinterpret_cast<longlong2*>(c + k) = v;
      }
    } else {
      if (r) r[k] = 6 * ce[tab[l]] + tab[written + l];
      if (c) c[k] = 6 * ce[tab[2 * written + l]] + tab[3 * written + l];
    }
  }
}

int check_batch(const pf3_batch* b, int what) {
  if (!b || b->kind < 0 || b->kind >= PF3_NKINDS || b->ne < 0) return PF3_E_BAD_ARG;
  if (b->ne == 0) return PF3_OK;
  if (!b->conn) return PF3_E_BAD_ARG;
  if (b->kind != PF3_SPRING && !b->props) return PF3_E_BAD_ARG;
  if (!b->state) {
    if (b->kind != PF3_SPRING && !b->x) return PF3_E_BAD_ARG;
    if ((b->kind == PF3_BEAMC || b->kind == PF3_BEAMLR || b->kind == PF3_SPRING) && !b->evec) return PF3_E_BAD_ARG;
    if ((what & (PF3_KG | PF3_FINT)) && !b->u) return PF3_E_BAD_ARG;
  }
  if (b->state && (b->state_flags & PF3_STATE_REFRESH_XE) && b->kind != PF3_SPRING && !b->x) return PF3_E_BAD_ARG;
  if (b->state && (b->state_flags & PF3_STATE_REFRESH_UE) && !b->u) return PF3_E_BAD_ARG;
  if (b->kind == PF3_SPRING && !b->eparam) return PF3_E_BAD_ARG;
  if ((what & PF3_KG) && (what & PF3_KG_STRESS)) return PF3_E_BAD_ARG;
  if ((what & PF3_KG_STRESS) && b->kind > PF3_TRIA3R) return PF3_E_UNSUPPORTED;
  if ((what & PF3_KG) && pf3::kind_sparse_size(b->kind, PF3_MAT_KG) == 0) return PF3_E_UNSUPPORTED;
  if ((what & PF3_M) && pf3::kind_sparse_size(b->kind, PF3_MAT_M) == 0) return PF3_E_UNSUPPORTED;
  if (what & PF3_M) {
    const int maxm = (b->kind <= PF3_TRIA3R) ? 2 : 1;
    if (b->mtype < 0 || b->mtype > maxm) return PF3_E_BAD_ARG;
  }
  return PF3_OK;
}

int launch_eval(pf3_context* ctx, const pf3::EvalArgs& A, int kind) {
  cudaError_t e;
  if (kind == PF3_QUAD4 || kind == PF3_QUAD4R)
    e = pf3::launch_quad(kind, A, ctx->stream);
  else if (kind == PF3_TRIA3R)
    e = pf3::launch_tria(A, ctx->stream);
  else
    e = pf3::launch_line(kind, A, ctx->stream);
  ++ctx->launches;
  return int(e);
}

void base_args(const pf3_batch* b, pf3::EvalArgs& A) {
  std::memset(&A, 0, sizeof(A));
  A.ne = b->ne;
  A.conn = b->conn;
  A.x = b->x;
  A.u = b->u;
  A.props = b->props;
  A.prop_id = b->prop_id;
  A.evec = b->evec;
  A.evec_stride = b->evec_stride;
  A.eparam = b->eparam;
  A.state = b->state;
  A.state_flags = b->state_flags;
  A.mtype = b->mtype;
  A.Nxx = b->stress[0];
  A.Nyy = b->stress[1];
  A.Nxy = b->stress[2];
}

}  // namespace

extern "C" {

int pf3_version(void) { return PF3_VERSION; }

const char* pf3_error_string(int code) {
  switch (code) {
    case PF3_OK: return "ok";
    case PF3_E_BAD_ARG: return "pyfe3d_b200: bad argument";
    case PF3_E_NO_DEVICE: return "pyfe3d_b200: no CUDA device (this library has no CPU fallback)";
    case PF3_E_UNSUPPORTED: return "pyfe3d_b200: matrix not defined for this element kind";
    case PF3_E_CAPACITY: return "pyfe3d_b200: size exceeds a plan/kernel capacity limit";
    default: return code > 0 ? cudaGetErrorString(cudaError_t(code)) : "pyfe3d_b200: unknown error";
  }
}

int pf3_device_count(int* n) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    cudaGetLastError();
    c = 0;
  }
  if (n) *n = c;
  return PF3_OK;
}

int pf3_create(int device, pf3_context** out) {
  if (!out) return PF3_E_BAD_ARG;
  int c = 0;
  pf3_device_count(&c);
  if (c <= 0) return PF3_E_NO_DEVICE;
  if (device < 0 || device >= c) return PF3_E_BAD_ARG;
  PF3_CUDA(cudaSetDevice(device));
  pf3_context* ctx = new pf3_context();
  ctx->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete ctx;
    return int(e);
  }
  ctx->own_stream = true;
  *out = ctx;
  return PF3_OK;
}

int pf3_destroy(pf3_context* ctx) {
  if (!ctx) return PF3_OK;
  cudaSetDevice(ctx->device);
  for (auto& kv : ctx->idx_tabs) cudaFree(kv.second);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->hostio) cudaFree(ctx->hostio);
  if (ctx->stage_host) cudaFreeHost(ctx->stage_host);
  if (ctx->stage_dev) cudaFree(ctx->stage_dev);
  if (ctx->solve_work) cudaFree(ctx->solve_work);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (cudaEvent_t ev : ctx->chunk_done)
    if (ev) cudaEventDestroy(ev);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return PF3_OK;
}

int pf3_set_stream(pf3_context* ctx, void* s) {
  if (!ctx) return PF3_E_BAD_ARG;
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = cudaStream_t(s);
  ctx->own_stream = false;
  return PF3_OK;
}

int pf3_synchronize(pf3_context* ctx) {
  int rc = use_device(ctx);
  if (rc) return rc;
  PF3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PF3_OK;
}

int pf3_malloc(pf3_context* ctx, size_t bytes, void** p) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!p) return PF3_E_BAD_ARG;
  PF3_CUDA(cudaMalloc(p, bytes ? bytes : 1));
  return PF3_OK;
}

int pf3_free(pf3_context* ctx, void* p) {
  int rc = use_device(ctx);
  if (rc) return rc;
  PF3_CUDA(cudaFree(p));
  return PF3_OK;
}

int pf3_memcpy_h2d(pf3_context* ctx, void* dst, const void* src, size_t bytes) {
  int rc = use_device(ctx);
  if (rc) return rc;
  PF3_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  PF3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PF3_OK;
}

int pf3_memcpy_d2h(pf3_context* ctx, void* dst, const void* src, size_t bytes) {
  int rc = use_device(ctx);
  if (rc) return rc;
  PF3_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PF3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PF3_OK;
}

int pf3_memset(pf3_context* ctx, void* dst, int byte, size_t bytes) {
  int rc = use_device(ctx);
  if (rc) return rc;
  PF3_CUDA(cudaMemsetAsync(dst, byte, bytes, ctx->stream));
  return PF3_OK;
}

int pf3_launch_count(pf3_context* ctx, int64_t* n) {
  if (!ctx || !n) return PF3_E_BAD_ARG;
  *n = ctx->launches;
  return PF3_OK;
}

int pf3_num_nodes(int kind) { return pf3::kind_nodes(kind); }
int pf3_sparse_size(int kind, int matrix) { return pf3::kind_sparse_size(kind, matrix); }
int pf3_written_size(int kind, int matrix, int mtype) { return pf3::make_layout(kind, matrix, mtype).written; }

int pf3_fill_indices(pf3_context* ctx, int kind, int matrix, int mtype, int64_t ne, const int64_t* conn,
                     int64_t init_k, int64_t* r, int64_t* c) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (ne < 0 || (ne > 0 && !conn)) return PF3_E_BAD_ARG;
  const int key = kind * 100 + matrix * 10 + mtype;
  pf3::BlockLayout L = pf3::make_layout(kind, matrix, mtype);
  if (L.written == 0) return PF3_E_UNSUPPORTED;
  if (ne == 0 || (!r && !c)) return PF3_OK;
  auto it = ctx->idx_tabs.find(key);
  if (it == ctx->idx_tabs.end()) {
    std::vector<int8_t> tab(size_t(4) * L.written);
    for (int l = 0; l < L.written; ++l) {
      tab[l] = L.la[l];
      tab[L.written + l] = L.li[l];
      tab[2 * L.written + l] = L.lb[l];
      tab[3 * L.written + l] = L.lj[l];
    }
    int8_t* d = nullptr;
    PF3_CUDA(cudaMalloc((void**)&d, tab.size()));
    PF3_CUDA(cudaMemcpyAsync(d, tab.data(), tab.size(), cudaMemcpyHostToDevice, ctx->stream));
    PF3_CUDA(cudaStreamSynchronize(ctx->stream));
    it = ctx->idx_tabs.emplace(key, d).first;
  }
  const int64_t n = ne * L.written;
  const int64_t ctas = (n + kIdxPerCta - 1) / kIdxPerCta;
  if (ctas > int64_t(0x7fffffff)) return PF3_E_CAPACITY;
  const unsigned grid = unsigned(std::max<int64_t>(1, ctas));
  int64_t* rk = r ? r + init_k : nullptr;
  int64_t* ck = c ? c + init_k : nullptr;
  const bool vec = L.written % 2 == 0 && L.size % 2 == 0 && ((uintptr_t(rk) | uintptr_t(ck)) & 15) == 0;
  if (vec)
    k_fill_indices<true><<<grid, 256, 0, ctx->stream>>>(it->second, L.written, L.size, L.nn, ne, conn, rk, ck);
  else
    k_fill_indices<false><<<grid, 256, 0, ctx->stream>>>(it->second, L.written, L.size, L.nn, ne, conn, rk, ck);
  ++ctx->launches;
  PF3_CUDA(cudaGetLastError());
  return PF3_OK;
}

int pf3_eval(pf3_context* ctx, const pf3_batch* b, int what, const pf3_coo* kc0, const pf3_coo* kg,
             const pf3_coo* m, double* fint) {
  int rc = use_device(ctx);
  if (rc) return rc;
  rc = check_batch(b, what);
  if (rc) return rc;
  if (b->ne == 0) return PF3_OK;
  const int nn = pf3::kind_nodes(b->kind);
  pf3::EvalArgs A;
  base_args(b, A);
  int kwhat = 0;
  if ((what & PF3_KC0) && kc0) {
    if (kc0->v) {
      A.kc0v = kc0->v;
      A.kc0_k0 = kc0->init_k;
      A.acc_kc0 = kc0->accumulate;
      kwhat |= PF3_KC0;
    }
    if (kc0->r || kc0->c) {
      rc = pf3_fill_indices(ctx, b->kind, PF3_MAT_KC0, 0, b->ne, b->conn, kc0->init_k, kc0->r, kc0->c);
      if (rc) return rc;
    }
  }
  if ((what & (PF3_KG | PF3_KG_STRESS)) && kg) {
    if (kg->v) {
      A.kgv = kg->v;
      A.kg_k0 = kg->init_k;
      A.acc_kg = kg->accumulate;
      kwhat |= what & (PF3_KG | PF3_KG_STRESS);
    }
    if (kg->r || kg->c) {
      rc = pf3_fill_indices(ctx, b->kind, PF3_MAT_KG, 0, b->ne, b->conn, kg->init_k, kg->r, kg->c);
      if (rc) return rc;
    }
  }
  if ((what & PF3_M) && m) {
    if (m->v) {
      A.mv = m->v;
      A.m_k0 = m->init_k;
      A.acc_m = m->accumulate;
      kwhat |= PF3_M;
    }
    if (m->r || m->c) {
      rc = pf3_fill_indices(ctx, b->kind, PF3_MAT_M, b->mtype, b->ne, b->conn, m->init_k, m->r, m->c);
      if (rc) return rc;
    }
  }
  if ((what & PF3_FINT) && fint) {
    rc = ensure_scratch(ctx, size_t(b->ne) * 6 * nn * sizeof(double));
    if (rc) return rc;
    A.fe = ctx->scratch;
    kwhat |= PF3_FINT;
  }
  if (kwhat == 0) return PF3_OK;
  A.what = kwhat;
  // Quad4/Quad4R matrices without fint/state/`+=`: the record + pair-lane kernels in element mode (COO slabs
  // leave as aligned bulk copies) are ~2x faster than the thread-per-element kernel.
  const bool quad = b->kind == PF3_QUAD4 || b->kind == PF3_QUAD4R;
  const bool aligned = (((uintptr_t)A.kc0v | (uintptr_t)A.kgv | (uintptr_t)A.mv) & 15) == 0 &&
                       ((A.kc0_k0 | A.kg_k0 | A.m_k0) & 1) == 0;
  if (quad && !(kwhat & PF3_FINT) && !b->state && !A.acc_kc0 && !A.acc_kg && !A.acc_m && aligned &&
      b->ne * 16 < (int64_t(1) << 31)) {
    pf3::FusedArgs F;
    std::memset(&F, 0, sizeof(F));
    F.A = A;
    F.nown = b->ne;
    F.rmax = 1;
    rc = ensure_scratch(ctx, size_t(b->ne) * pf3::fused_record_stride(b->kind, F.A) * sizeof(double));
    if (rc) return rc;
    cudaError_t e = pf3::launch_quad_fused(b->kind, F, ctx->scratch, ctx->stream, &ctx->launches, 3);
    return int(e);
  }
  // Tria3R the same way through the triangle record + node-lane kernels (KG slabs are not 16-byte multiples and go
  // out as plain stores, so only KC0 and M need the alignment)
  const bool taligned = (((uintptr_t)A.kc0v | (uintptr_t)A.mv) & 15) == 0 && ((A.kc0_k0 | A.m_k0) & 1) == 0;
  if (b->kind == PF3_TRIA3R && !(kwhat & PF3_FINT) && !b->state && !A.acc_kc0 && !A.acc_kg && !A.acc_m && taligned &&
      b->ne * 9 < (int64_t(1) << 31)) {
    pf3::FusedArgs F;
    std::memset(&F, 0, sizeof(F));
    F.A = A;
    F.nown = (b->ne + 2) / 3;
    F.rmax = 1;
    rc = ensure_scratch(ctx, size_t(b->ne) * pf3::tria_fused_record_stride(F.A) * sizeof(double));
    if (rc) return rc;
    cudaError_t e = pf3::launch_tria_fused(F, ctx->scratch, ctx->stream, &ctx->launches);
    return int(e);
  }
  rc = launch_eval(ctx, A, b->kind);
  if (rc) return rc;
  if (kwhat & PF3_FINT) {
    rc = pf3::fint_gather(ctx->stream, b->ne, nn, b->nnodes, b->conn, ctx->scratch, fint, &ctx->launches);
    if (rc) return rc;
  }
  return PF3_OK;
}

int pf3_eval_state(pf3_context* ctx, const pf3_batch* b, double* state_out) {
  int rc = use_device(ctx);
  if (rc) return rc;
  rc = check_batch(b, 0);
  if (rc) return rc;
  if (!state_out) return PF3_E_BAD_ARG;
  if (b->ne == 0) return PF3_OK;
  pf3::EvalArgs A;
  base_args(b, A);
  A.what = 0;
  A.state_out = state_out;
  return launch_eval(ctx, A, b->kind);
}

int pf3_eval_finte(pf3_context* ctx, const pf3_batch* b, double* finte_out) {
  int rc = use_device(ctx);
  if (rc) return rc;
  rc = check_batch(b, PF3_FINT);
  if (rc) return rc;
  if (!finte_out) return PF3_E_BAD_ARG;
  if (b->ne == 0) return PF3_OK;
  pf3::EvalArgs A;
  base_args(b, A);
  A.what = PF3_FINT;
  A.finte = finte_out;
  return launch_eval(ctx, A, b->kind);
}

int pf3_quad4_update_BL(pf3_context* ctx, int64_t n, const double* xe, double xi, double eta, double* out) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!xe || !out))) return PF3_E_BAD_ARG;
  cudaError_t e = pf3::launch_quad4_BL(n, xe, xi, eta, out, ctx->stream);
  ++ctx->launches;
  return int(e);
}

int pf3_plan_create(pf3_context* ctx, int matrix, int64_t nnodes, int ngroups, const pf3_batch* groups,
                    const int64_t* coo_offsets, int64_t node_begin, int64_t node_end, pf3_plan** plan) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!groups || !plan || matrix < 0 || matrix > PF3_MAT_CA) return PF3_E_BAD_ARG;
  return pf3::plan_create_structured(ctx->device, ctx->stream, matrix, nnodes, ngroups, groups, coo_offsets,
                                     node_begin, node_end, &ctx->launches, plan);
}

int pf3_plan_create_coo(pf3_context* ctx, int64_t n, int64_t nnz_coo, const int64_t* r, const int64_t* c,
                        pf3_plan** plan) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!plan) return PF3_E_BAD_ARG;
  return pf3::plan_create_generic(ctx->device, ctx->stream, n, nnz_coo, r, c, &ctx->launches, plan);
}

int pf3_plan_nnz(const pf3_plan* plan, int64_t* nnz) {
  if (!plan || !nnz) return PF3_E_BAD_ARG;
  *nnz = pf3::plan_nnz(plan);
  return PF3_OK;
}

int pf3_plan_nrows(const pf3_plan* plan, int64_t* nrows) {
  if (!plan || !nrows) return PF3_E_BAD_ARG;
  *nrows = pf3::plan_nrows(plan);
  return PF3_OK;
}

int pf3_plan_fint(pf3_context* ctx, const pf3_plan* plan, int group, const pf3_batch* b, double* fint) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!plan || !fint) return PF3_E_BAD_ARG;
  rc = check_batch(b, PF3_FINT);
  if (rc) return rc;
  const int nn = pf3::kind_nodes(b->kind);
  if (pf3::plan_group_kind_nn(plan, group) != nn || pf3::plan_group_ne_of(plan, group) != b->ne) return PF3_E_BAD_ARG;
  if (b->ne == 0) return PF3_OK;
  rc = ensure_scratch(ctx, size_t(b->ne) * 6 * nn * sizeof(double));
  if (rc) return rc;
  pf3::EvalArgs A;
  base_args(b, A);
  A.what = PF3_FINT;
  A.fe = ctx->scratch;
  rc = launch_eval(ctx, A, b->kind);
  if (rc) return rc;
  return pf3::plan_fint_gather(plan, ctx->stream, group, ctx->scratch, fint, &ctx->launches);
}

int pf3_plan_nblocks(const pf3_plan* plan, int64_t* nblk) {
  if (!plan || !nblk) return PF3_E_BAD_ARG;
  *nblk = pf3::plan_nblocks(plan);
  return PF3_OK;
}

namespace {

// Quad4 / Quad4R host-buffer step as a pipeline: K1 once, then K2 over PF3_HOST_CHUNKS ranges of node pairs; a node's
// CSR rows are complete when its CTA retires and a range's rows are contiguous in every value array, so each range
// goes back to the host on a second stream while the next one is being evaluated.
int fused_pipelined(pf3_context* ctx, int kind, pf3::FusedArgs& F, double* const dev_out[3], double* const host_out[3],
                    const int per_block[3]) {
  if (!ctx->copy_stream) PF3_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (cudaEvent_t& ev : ctx->chunk_done)
    if (!ev) PF3_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  const int64_t npairs = (F.nown + 1) / 2;
  int64_t pair_at[PF3_HOST_CHUNKS + 1], block_at[PF3_HOST_CHUNKS + 1];
  for (int c = 0; c <= PF3_HOST_CHUNKS; ++c) {
    pair_at[c] = npairs * c / PF3_HOST_CHUNKS;
    const int64_t node = std::min<int64_t>(2 * pair_at[c], F.nown);
    PF3_CUDA(cudaMemcpyAsync(&block_at[c], F.brow_ptr + node, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  }
  PF3_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaError_t e = pf3::launch_quad_fused(kind, F, ctx->scratch, ctx->stream, &ctx->launches, 1);
  if (e != cudaSuccess) return int(e);
  for (int c = 0; c < PF3_HOST_CHUNKS; ++c) {
    F.pair_first = pair_at[c];
    F.pair_count = pair_at[c + 1] - pair_at[c];
    if (F.pair_count <= 0) continue;
    e = pf3::launch_quad_fused(kind, F, ctx->scratch, ctx->stream, &ctx->launches, 2);
    if (e != cudaSuccess) return int(e);
    PF3_CUDA(cudaEventRecord(ctx->chunk_done[c], ctx->stream));
    PF3_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_done[c], 0));
    for (int k = 0; k < 3; ++k) {
      if (!host_out[k] || !dev_out[k]) continue;
      const int64_t lo = block_at[c] * per_block[k], hi = block_at[c + 1] * per_block[k];
      if (hi > lo)
        PF3_CUDA(cudaMemcpyAsync(host_out[k] + lo, dev_out[k] + lo, size_t(hi - lo) * sizeof(double),
                                 cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
  }
  F.pair_first = F.pair_count = 0;
  PF3_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  return PF3_OK;
}

// host_out (nullable): host destinations of the three CSR value arrays; *copied reports whether they were filled here.
int eval_assemble_impl(pf3_context* ctx, const pf3_batch* b, const pf3_plan* plan, int what, const pf3_coo* kc0,
                       const pf3_coo* kg, const pf3_coo* m, double* csr_kc0, double* csr_kg, double* csr_m,
                       double* const host_out[3], bool* copied) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!plan) return PF3_E_BAD_ARG;
  rc = check_batch(b, what);
  if (rc) return rc;
  if (b->state || (what & PF3_FINT)) return PF3_E_UNSUPPORTED;
  if (b->ne != pf3::plan_group_ne(plan)) return PF3_E_BAD_ARG;
  if (((what & PF3_KC0) && !csr_kc0) || ((what & (PF3_KG | PF3_KG_STRESS)) && !csr_kg) || ((what & PF3_M) && !csr_m))
    return PF3_E_BAD_ARG;
  for (const pf3_coo* c : {kc0, kg, m}) {
    if (c && c->accumulate) return PF3_E_UNSUPPORTED;
    // COO slabs leave shared memory as 16-byte-aligned bulk copies
    if (c && c->v && (((uintptr_t)c->v & 15) != 0 || (c->init_k & 1) != 0)) return PF3_E_UNSUPPORTED;
  }
  if ((((uintptr_t)csr_kc0) | ((uintptr_t)csr_kg) | ((uintptr_t)csr_m)) & 15) return PF3_E_UNSUPPORTED;
  if (b->ne == 0) return PF3_OK;
  pf3::FusedArgs F;
  std::memset(&F, 0, sizeof(F));
  const bool tria = b->kind == PF3_TRIA3R;
  rc = tria ? pf3::plan_fused_args_tria(plan, &F, ctx->stream, &ctx->launches)
            : pf3::plan_fused_args(plan, b->kind, &F, ctx->stream, &ctx->launches);
  if (rc) return rc;
  base_args(b, F.A);
  F.A.what = what & (PF3_KC0 | PF3_KG | PF3_KG_STRESS | PF3_M);
  if (kc0 && kc0->v) { F.A.kc0v = kc0->v; F.A.kc0_k0 = kc0->init_k; }
  if (kg && kg->v) { F.A.kgv = kg->v; F.A.kg_k0 = kg->init_k; }
  if (m && m->v) { F.A.mv = m->v; F.A.m_k0 = m->init_k; }
  F.csr_kc0 = csr_kc0;
  F.csr_kg = csr_kg;
  F.csr_m = csr_m;
  rc = ensure_scratch(ctx, size_t(b->ne) * (tria ? pf3::tria_fused_record_stride(F.A) : pf3::fused_record_stride(b->kind, F.A)) *
                               sizeof(double));
  if (rc) return rc;
  // the pipeline pays off when the device->host copies dominate: enough node pairs for PF3_HOST_CHUNKS full waves
  if (host_out && !tria && F.nown >= int64_t(PF3_HOST_CHUNKS) * 16384) {
    double* const dev_out[3] = {csr_kc0, csr_kg, csr_m};
    const int per_block[3] = {36, 9, b->mtype == 2 ? 18 : 30};
    rc = fused_pipelined(ctx, b->kind, F, dev_out, host_out, per_block);
    if (rc) return rc;
    *copied = true;
  } else {
    cudaError_t e = tria ? pf3::launch_tria_fused(F, ctx->scratch, ctx->stream, &ctx->launches)
                         : pf3::launch_quad_fused(b->kind, F, ctx->scratch, ctx->stream, &ctx->launches, 3);
    if (e != cudaSuccess) return int(e);
  }
  const pf3_coo* cs[3] = {(what & PF3_KC0) ? kc0 : nullptr, (what & (PF3_KG | PF3_KG_STRESS)) ? kg : nullptr,
                          (what & PF3_M) ? m : nullptr};
  for (int k = 0; k < 3; ++k)
    if (cs[k] && (cs[k]->r || cs[k]->c)) {
      rc = pf3_fill_indices(ctx, b->kind, k, k == 2 ? b->mtype : 0, b->ne, b->conn, cs[k]->init_k, cs[k]->r, cs[k]->c);
      if (rc) return rc;
    }
  return PF3_OK;
}

}  // namespace

int pf3_eval_assemble(pf3_context* ctx, const pf3_batch* b, const pf3_plan* plan, int what, const pf3_coo* kc0,
                      const pf3_coo* kg, const pf3_coo* m, double* csr_kc0, double* csr_kg, double* csr_m) {
  return eval_assemble_impl(ctx, b, plan, what, kc0, kg, m, csr_kc0, csr_kg, csr_m, nullptr, nullptr);
}

// Fused evaluate + assemble of ONE Quad4 / Quad4R group of a multi-group plan into the union layouts.
int pf3_eval_assemble_group(pf3_context* ctx, const pf3_batch* b, const pf3_plan* plan, int group, int what,
                            const pf3_coo* kc0, const pf3_coo* kg, const pf3_coo* m, double* csr_kc0, double* csr_kg,
                            double* csr_m) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!plan) return PF3_E_BAD_ARG;
  rc = check_batch(b, what);
  if (rc) return rc;
  if (b->kind != PF3_QUAD4 && b->kind != PF3_QUAD4R) return PF3_E_UNSUPPORTED;
  if (b->state || (what & PF3_FINT)) return PF3_E_UNSUPPORTED;
  if (b->ne != pf3::plan_group_ne_of(plan, group)) return PF3_E_BAD_ARG;
  if (((what & PF3_KC0) && !csr_kc0) || ((what & (PF3_KG | PF3_KG_STRESS)) && !csr_kg) || ((what & PF3_M) && !csr_m))
    return PF3_E_BAD_ARG;
  for (const pf3_coo* c : {kc0, kg, m}) {
    if (c && c->accumulate) return PF3_E_UNSUPPORTED;
    if (c && c->v && (((uintptr_t)c->v & 15) != 0 || (c->init_k & 1) != 0)) return PF3_E_UNSUPPORTED;
  }
  if ((((uintptr_t)csr_kc0) | ((uintptr_t)csr_kg) | ((uintptr_t)csr_m)) & 15) return PF3_E_UNSUPPORTED;
  if (b->ne == 0) return PF3_OK;
  pf3::FusedArgs F;
  std::memset(&F, 0, sizeof(F));
  rc = pf3::plan_fused_args_group(plan, group, &F, ctx->stream, &ctx->launches);
  if (rc) return rc;
  const int bits[3] = {PF3_KC0, PF3_KG | PF3_KG_STRESS, PF3_M};
  for (int k = 0; k < 3; ++k) {
    if (!(what & bits[k])) continue;
    rc = pf3::plan_union_map(plan, group, k, k == 2 ? b->mtype : 0, &F.um[k]);
    if (rc) return rc;
  }
  F.zero_empty = 1;       // nodes that only other groups touch still get their rows initialised
  base_args(b, F.A);
  F.A.what = what & (PF3_KC0 | PF3_KG | PF3_KG_STRESS | PF3_M);
  if (kc0 && kc0->v) { F.A.kc0v = kc0->v; F.A.kc0_k0 = kc0->init_k; }
  if (kg && kg->v) { F.A.kgv = kg->v; F.A.kg_k0 = kg->init_k; }
  if (m && m->v) { F.A.mv = m->v; F.A.m_k0 = m->init_k; }
  F.csr_kc0 = csr_kc0;
  F.csr_kg = csr_kg;
  F.csr_m = csr_m;
  rc = ensure_scratch(ctx, size_t(b->ne) * pf3::fused_record_stride(b->kind, F.A) * sizeof(double));
  if (rc) return rc;
  cudaError_t e = pf3::launch_quad_fused(b->kind, F, ctx->scratch, ctx->stream, &ctx->launches, 3);
  if (e != cudaSuccess) return int(e);
  const pf3_coo* cs[3] = {(what & PF3_KC0) ? kc0 : nullptr, (what & (PF3_KG | PF3_KG_STRESS)) ? kg : nullptr,
                          (what & PF3_M) ? m : nullptr};
  for (int k = 0; k < 3; ++k)
    if (cs[k] && (cs[k]->r || cs[k]->c)) {
      rc = pf3_fill_indices(ctx, b->kind, k, k == 2 ? b->mtype : 0, b->ne, b->conn, cs[k]->init_k, cs[k]->r, cs[k]->c);
      if (rc) return rc;
    }
  return PF3_OK;
}

int pf3_plan_assemble_add(pf3_context* ctx, const pf3_plan* plan, const double* coo_v, double* csr_v, int skip_group) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!plan || !coo_v || !csr_v || skip_group < 0) return PF3_E_BAD_ARG;
  return pf3::plan_assemble_gather(plan, ctx->stream, coo_v, csr_v, skip_group, &ctx->launches);
}

// Host-buffer step on a fixed mesh: x, u come from host memory, the assembled CSR values go back to host memory.
int pf3_eval_assemble_host(pf3_context* ctx, const pf3_batch* b, const pf3_plan* plan, int what, const pf3_coo* kc0,
                           const pf3_coo* kg, const pf3_coo* m, double* csr_kc0_host, double* csr_kg_host,
                           double* csr_m_host) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!b || !plan || b->nnodes <= 0) return PF3_E_BAD_ARG;
  const int64_t nblk = pf3::plan_nblocks(plan);
  const size_t nx = size_t(b->nnodes) * 3, nu = b->u ? size_t(b->nnodes) * 6 : 0;
  const size_t n0 = ((what & PF3_KC0) && csr_kc0_host) ? size_t(nblk) * 36 : 0;
  const size_t n1 = ((what & (PF3_KG | PF3_KG_STRESS)) && csr_kg_host) ? size_t(nblk) * 9 : 0;
  const size_t n2 = ((what & PF3_M) && csr_m_host) ? size_t(nblk) * (b->mtype == 2 ? 18 : 30) : 0;
  auto up = [](size_t n) { return (n + 1) & ~size_t(1); };   // keep every piece 16-byte aligned
  const size_t total = (up(nx) + up(nu) + up(n0) + up(n1) + up(n2)) * sizeof(double);
  if (total > ctx->hostio_bytes) {
    if (ctx->hostio) cudaFree(ctx->hostio);
    ctx->hostio = nullptr;
    ctx->hostio_bytes = 0;
    PF3_CUDA(cudaMalloc((void**)&ctx->hostio, total));
    ctx->hostio_bytes = total;
  }
  double* dx = ctx->hostio;
  double* du = dx + up(nx);
  double* d0 = du + up(nu);
  double* d1 = d0 + up(n0);
  double* d2 = d1 + up(n1);
  if (!b->x) return PF3_E_BAD_ARG;
  PF3_CUDA(cudaMemcpyAsync(dx, b->x, nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (nu) PF3_CUDA(cudaMemcpyAsync(du, b->u, nu * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  pf3_batch db = *b;
  db.x = dx;
  db.u = nu ? du : nullptr;
  double* const host_out[3] = {n0 ? csr_kc0_host : nullptr, n1 ? csr_kg_host : nullptr, n2 ? csr_m_host : nullptr};
  bool copied = false;
  rc = eval_assemble_impl(ctx, &db, plan, what, kc0, kg, m, n0 ? d0 : nullptr, n1 ? d1 : nullptr, n2 ? d2 : nullptr,
                          host_out, &copied);
  if (rc) return rc;
  if (!copied) {
    if (n0) PF3_CUDA(cudaMemcpyAsync(csr_kc0_host, d0, n0 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (n1) PF3_CUDA(cudaMemcpyAsync(csr_kg_host, d1, n1 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (n2) PF3_CUDA(cudaMemcpyAsync(csr_m_host, d2, n2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  PF3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PF3_OK;
}

int pf3_plan_pattern(pf3_context* ctx, const pf3_plan* plan, int64_t* indptr, int64_t* indices) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!plan) return PF3_E_BAD_ARG;
  return pf3::plan_pattern(plan, ctx->stream, indptr, indices, &ctx->launches);
}

int pf3_plan_assemble(pf3_context* ctx, const pf3_plan* plan, const double* coo_v, double* csr_v) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!plan || !coo_v || !csr_v) return PF3_E_BAD_ARG;
  return pf3::plan_assemble(plan, ctx->stream, coo_v, csr_v, &ctx->launches);
}

int pf3_spmv_csr(pf3_context* ctx, int64_t nrows, const int64_t* indptr, const int64_t* indices,
                 const double* vals, const double* x, double* y) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (nrows < 0 || !indptr || !x || !y) return PF3_E_BAD_ARG;
  return pf3::spmv_csr(ctx->stream, nrows, indptr, indices, vals, x, y, &ctx->launches);
}

int pf3_spmv_csr_masked(pf3_context* ctx, int64_t nrows, const int64_t* indptr, const int64_t* indices,
                        const double* vals, const unsigned char* free_dof, const double* x, double* y) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (nrows < 0 || !indptr || !x || !y || !free_dof) return PF3_E_BAD_ARG;
  return pf3::spmv_csr_masked(ctx->stream, nrows, indptr, indices, vals, free_dof, x, y, &ctx->launches);
}

int pf3_csr_diagonal(pf3_context* ctx, int64_t nrows, const int64_t* indptr, const int64_t* indices,
                     const double* vals, int64_t row0, double* diag) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (nrows < 0 || !indptr || !diag) return PF3_E_BAD_ARG;
  return pf3::csr_diagonal(ctx->stream, nrows, indptr, indices, vals, row0, diag, &ctx->launches);
}

int pf3_eval_aero(pf3_context* ctx, const pf3_batch* b, int what, const pf3_coo* ka_beta, const pf3_coo* ka_gamma,
                  const pf3_coo* ca) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!b || b->ne < 0) return PF3_E_BAD_ARG;
  if (b->kind != PF3_QUAD4 && b->kind != PF3_QUAD4R) return PF3_E_UNSUPPORTED;
  if (b->ne == 0) return PF3_OK;
  if (!b->conn || (!b->state && !b->x)) return PF3_E_BAD_ARG;
  if (b->state && (b->state_flags & PF3_STATE_REFRESH_XE) && !b->x) return PF3_E_BAD_ARG;
  pf3::EvalArgs A;
  base_args(b, A);
  A.u = nullptr;
  A.evec = nullptr;   // the material axis plays no role in the aerodynamic matrices
  pf3::AeroOut O;
  std::memset(&O, 0, sizeof(O));
  const pf3_coo* dst[3] = {ka_beta, ka_gamma, ca};
  const int bits[3] = {PF3_KA_BETA, PF3_KA_GAMMA, PF3_CA};
  bool any = false;
  for (int w = 0; w < 3; ++w) {
    if (!(what & bits[w]) || !dst[w]) continue;
    if (dst[w]->v) {
      O.v[w] = dst[w]->v;
      O.k0[w] = dst[w]->init_k;
      O.acc[w] = dst[w]->accumulate;
      any = true;
    }
    if (dst[w]->r || dst[w]->c) {
      rc = pf3_fill_indices(ctx, b->kind, PF3_MAT_KA_BETA + w, 0, b->ne, b->conn, dst[w]->init_k, dst[w]->r,
                            dst[w]->c);
      if (rc) return rc;
    }
  }
  if (!any) return PF3_OK;
  cudaError_t e = pf3::launch_quad_aero(A, O, ctx->stream);
  ++ctx->launches;
  return int(e);
}

int pf3_laminate_props(pf3_context* ctx, int64_t nrows, int nplies, const double* thetadeg, int64_t theta_stride,
                       const double* plyt, int64_t plyt_stride, const double* lamina, int64_t lamina_stride,
                       const double* offset, int64_t offset_stride, int calc_scf, double* props_out) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (nrows < 0 || nplies <= 0 || !thetadeg || !plyt || !lamina || !props_out) return PF3_E_BAD_ARG;
  if (theta_stride < 0 || plyt_stride < 0 || lamina_stride < 0 || offset_stride < 0) return PF3_E_BAD_ARG;
  cudaError_t e = pf3::launch_laminate_props(nrows, nplies, thetadeg, theta_stride, plyt, plyt_stride, lamina,
                                             lamina_stride, offset, offset_stride, calc_scf, props_out, ctx->stream);
  ++ctx->launches;
  return int(e);
}

int pf3_lamination_parameter_props(pf3_context* ctx, int64_t nrows, const double* thickness, int64_t thickness_stride,
                                   const double* invariants, int64_t invariants_stride, const double* lp,
                                   int64_t lp_stride, const double* rho, int64_t rho_stride, int var_mask,
                                   int grad_complete, double* props_out, double* grad_out) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (nrows < 0 || !thickness || !invariants || !lp || (!props_out && !grad_out)) return PF3_E_BAD_ARG;
  if (thickness_stride < 0 || invariants_stride < 0 || lp_stride < 0 || rho_stride < 0) return PF3_E_BAD_ARG;
  if (var_mask < 0 || var_mask >= (1 << PF3_LP_NVARS) || (grad_out && !var_mask)) return PF3_E_BAD_ARG;
  cudaError_t e = pf3::launch_lp_props(nrows, thickness, thickness_stride, invariants, invariants_stride, lp, lp_stride,
                                       rho, rho_stride, var_mask, grad_complete, props_out, grad_out, ctx->stream);
  ++ctx->launches;
  return int(e);
}

int pf3_plan_spmv(pf3_context* ctx, const pf3_plan* plan, const double* vals, const unsigned char* free_dof,
                  const double* x, double* y) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!plan || !vals || !x || !y) return PF3_E_BAD_ARG;
  return pf3::plan_spmv(plan, ctx->stream, vals, free_dof, x, y, &ctx->launches);
}

int pf3_plan_diagonal(pf3_context* ctx, const pf3_plan* plan, const double* vals, double* diag) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!plan || !vals || !diag) return PF3_E_BAD_ARG;
  return pf3::plan_diagonal(plan, ctx->stream, vals, diag, &ctx->launches);
}

int pf3_plan_cg(pf3_context* ctx, int nops, const pf3_plan* const* plans, const double* const* vals,
                const double* coefs, const unsigned char* free_dof, const double* b, double* x, int use_x0,
                double rtol, double atol, int maxiter, int flags, pf3_cg_info* info) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (nops < 1 || nops > 8 || !plans || !vals || !b || !x) return PF3_E_BAD_ARG;
  pf3::CgOp ops[8];
  for (int i = 0; i < nops; ++i) {
    if (!plans[i] || !vals[i]) return PF3_E_BAD_ARG;
    ops[i] = {plans[i], vals[i], coefs ? coefs[i] : 1.};
  }
  const int64_t n = pf3::plan_nrows(plans[0]);
  if (n <= 0) return PF3_E_BAD_ARG;
  const size_t need = pf3::cg_work_bytes(n);
  if (need > ctx->solve_work_bytes) {
    if (ctx->solve_work) cudaFree(ctx->solve_work);
    ctx->solve_work = nullptr;
    ctx->solve_work_bytes = 0;
    PF3_CUDA(cudaMalloc(&ctx->solve_work, need));
    ctx->solve_work_bytes = need;
  }
  int iters = 0, status = 0;
  double resid = 0., bnorm = 0.;
  rc = pf3::plan_cg(ctx->stream, nops, ops, n, free_dof, b, x, use_x0, rtol, atol, maxiter, flags, ctx->solve_work,
                    &iters, &status, &resid, &bnorm, &ctx->launches);
  if (info) {
    info->iterations = iters;
    info->status = status;
    info->residual = resid;
    info->bnorm = bnorm;
  }
  return rc;
}

size_t pf3_cg_shard_work_bytes(void) { return pf3::cg_shard_work_bytes(); }

int pf3_cg_shard_dot(pf3_context* ctx, int64_t n, const double* p, const double* ap, double* sc, void* work) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (n <= 0 || !p || !ap || !sc || !work) return PF3_E_BAD_ARG;
  return pf3::cg_shard_dot(ctx->stream, n, p, ap, sc, work, &ctx->launches);
}

int pf3_cg_shard_update(pf3_context* ctx, int64_t n, const double* p, const double* ap, const double* minv, double* x,
                        double* r, double* sc, void* work) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (n <= 0 || !p || !ap || !minv || !x || !r || !sc || !work) return PF3_E_BAD_ARG;
  return pf3::cg_shard_update(ctx->stream, n, p, ap, minv, x, r, sc, work, &ctx->launches);
}

int pf3_cg_shard_dir(pf3_context* ctx, int64_t n, const double* r, const double* minv, double* p, double* sc, void* work) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (n <= 0 || !r || !minv || !p || !sc || !work) return PF3_E_BAD_ARG;
  return pf3::cg_shard_dir(ctx->stream, n, r, minv, p, sc, work, &ctx->launches);
}

int pf3_plan_spmv_scaled(pf3_context* ctx, const pf3_plan* plan, const double* vals, const unsigned char* free_dof,
                         const double* scale, const double* x, double* y) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!plan || !vals || !scale || !x || !y) return PF3_E_BAD_ARG;
  const int64_t n = pf3::plan_nrows(plan);
  if (n <= 0) return PF3_E_BAD_ARG;
  const size_t need = pf3::cg_work_bytes(n);
  if (need > ctx->solve_work_bytes) {
    if (ctx->solve_work) cudaFree(ctx->solve_work);
    ctx->solve_work = nullptr;
    ctx->solve_work_bytes = 0;
    PF3_CUDA(cudaMalloc(&ctx->solve_work, need));
    ctx->solve_work_bytes = need;
  }
  return pf3::plan_spmv_scaled(plan, ctx->stream, vals, free_dof, scale, x, y, static_cast<double*>(ctx->solve_work),
                               n, &ctx->launches);
}

int pf3_csr_compact_symbolic(pf3_context* ctx, int64_t nrows, int64_t ncols, const int64_t* indptr,
                             const int64_t* indices, const unsigned char* free_dof, int flags, int64_t row0,
                             int64_t* colmap, int64_t* out_indptr, int64_t* nkeep, int64_t* nnz) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!indptr || !indices || !colmap || !out_indptr) return PF3_E_BAD_ARG;
  return pf3::csr_compact_symbolic(ctx->stream, nrows, ncols, indptr, indices, free_dof, flags & PF3_COMPACT_UPPER, row0,
                                   colmap, out_indptr, nkeep, nnz, &ctx->launches);
}

int pf3_csr_compact_fill(pf3_context* ctx, int64_t nrows, int64_t ncols, const int64_t* indptr,
                         const int64_t* indices, const double* vals, const unsigned char* free_dof, int flags,
                         int64_t row0, const int64_t* colmap, const int64_t* out_indptr, int64_t* out_indices,
                         double* out_vals) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!indptr || !indices || !colmap || !out_indptr || (out_vals && !vals)) return PF3_E_BAD_ARG;
  return pf3::csr_compact_fill(ctx->stream, nrows, ncols, indptr, indices, vals, free_dof, flags & PF3_COMPACT_UPPER,
                               row0, colmap, out_indptr, out_indices, out_vals, &ctx->launches);
}

}  // extern "C"

namespace {

// Small host-pointer calls (the per-element drop-in classes make one per method call): every input is packed into ONE
// pinned staging buffer and goes to the device in one copy, the kernels run on the device image, and one copy brings
// the image back; only the `out` pieces are written to the caller's arrays.  Two PCIe transfers and one stream
// synchronisation per call instead of one cudaMalloc + copy per array.
struct Stager {
  pf3_context* ctx;
  size_t off = 0;
  bool overflow = false;
  struct Seg { void* host; size_t bytes, off; };
  Seg outs[16];
  int nouts = 0;
  explicit Stager(pf3_context* c) : ctx(c) {}
  int init() {
    if (!ctx->stage_host) {
      PF3_CUDA(cudaHostAlloc((void**)&ctx->stage_host, PF3_STAGE_BYTES, cudaHostAllocDefault));
      PF3_CUDA(cudaMalloc((void**)&ctx->stage_dev, PF3_STAGE_BYTES));
    }
    return PF3_OK;
  }
  template <class T>
  T* add(const T* host, size_t count, bool in, bool out) {
    if (!host || count == 0) return nullptr;
    const size_t bytes = count * sizeof(T);
    off = (off + 15) & ~size_t(15);
    if (off + bytes > PF3_STAGE_BYTES || (out && nouts == 16)) {
      overflow = true;
      return nullptr;
    }
    if (in) std::memcpy(ctx->stage_host + off, host, bytes);
    if (out) outs[nouts++] = Seg{(void*)host, bytes, off};
    T* d = reinterpret_cast<T*>(ctx->stage_dev + off);
    off += bytes;
    return d;
  }
  int upload() {
    if (off) PF3_CUDA(cudaMemcpyAsync(ctx->stage_dev, ctx->stage_host, off, cudaMemcpyHostToDevice, ctx->stream));
    return PF3_OK;
  }
  int download() {
    if (off) PF3_CUDA(cudaMemcpyAsync(ctx->stage_host, ctx->stage_dev, off, cudaMemcpyDeviceToHost, ctx->stream));
    PF3_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < nouts; ++i) std::memcpy(outs[i].host, ctx->stage_host + outs[i].off, outs[i].bytes);
    return PF3_OK;
  }
};

size_t batch_host_bytes(const pf3_batch* hb) {
  const int nn = pf3::kind_nodes(hb->kind);
  const int pstride = (hb->kind <= PF3_TRIA3R) ? PF3_SHELLPROP_STRIDE : PF3_BEAMPROP_STRIDE;
  size_t n = size_t(hb->ne) * nn * 8 + 256;
  if (hb->x) n += size_t(hb->nnodes) * 24;
  if (hb->u) n += size_t(hb->nnodes) * 48;
  if (hb->props) n += size_t(hb->nprop) * pstride * 8;
  if (hb->prop_id) n += size_t(hb->ne) * 4;
  if (hb->evec) n += (hb->evec_stride ? size_t(hb->ne) * hb->evec_stride : 0) * 8 + 64;
  if (hb->eparam) n += size_t(hb->ne) * PF3_EPARAM_STRIDE * 8;
  if (hb->state) n += size_t(hb->ne) * PF3_STATE_STRIDE * 8;
  return n;
}

// device image of a host batch inside the stager
void stage_batch(Stager& S, const pf3_batch* hb, pf3_batch* db) {
  const int nn = pf3::kind_nodes(hb->kind);
  const int pstride = (hb->kind <= PF3_TRIA3R) ? PF3_SHELLPROP_STRIDE : PF3_BEAMPROP_STRIDE;
  *db = *hb;
  db->conn = S.add(hb->conn, size_t(hb->ne) * nn, true, false);
  db->x = S.add(hb->x, size_t(hb->nnodes) * 3, true, false);
  db->u = S.add(hb->u, size_t(hb->nnodes) * 6, true, false);
  db->props = S.add(hb->props, size_t(hb->nprop) * pstride, true, false);
  db->prop_id = S.add(hb->prop_id, size_t(hb->ne), true, false);
  if (hb->evec) {
    const int width = (hb->kind == PF3_SPRING) ? 6 : 3;
    const size_t cnt = hb->evec_stride ? size_t(hb->ne - 1) * hb->evec_stride + width : width;
    db->evec = S.add(hb->evec, cnt, true, false);
  }
  db->eparam = S.add(hb->eparam, size_t(hb->ne) * PF3_EPARAM_STRIDE, true, false);
  db->state = S.add(hb->state, size_t(hb->ne) * PF3_STATE_STRIDE, true, false);
}

// pf3_eval on a staged image; *handled = false when the call does not fit the staging buffer
int eval_host_staged(pf3_context* ctx, const pf3_batch* hb, int what, const pf3_coo* kc0, const pf3_coo* kg,
                     const pf3_coo* m, double* fint, bool* handled) {
  *handled = false;
  const pf3_coo* hs[3] = {(what & PF3_KC0) ? kc0 : nullptr, (what & (PF3_KG | PF3_KG_STRESS)) ? kg : nullptr,
                          (what & PF3_M) ? m : nullptr};
  size_t need = batch_host_bytes(hb);
  for (int k = 0; k < 3; ++k)
    if (hs[k]) need += size_t(hb->ne) * pf3::kind_sparse_size(hb->kind, k) * 24 + 64;
  if ((what & PF3_FINT) && fint) need += size_t(hb->nnodes) * 48 + 16;
  if (need > PF3_STAGE_BYTES) return PF3_OK;
  Stager S(ctx);
  int rc = S.init();
  if (rc) return rc;
  pf3_batch b;
  stage_batch(S, hb, &b);
  pf3_coo dc[3];
  for (int k = 0; k < 3; ++k) {
    std::memset(&dc[k], 0, sizeof(pf3_coo));
    const size_t n = size_t(hb->ne) * pf3::kind_sparse_size(hb->kind, k);
    if (!hs[k] || n == 0) {
      hs[k] = nullptr;
      continue;
    }
    dc[k].accumulate = hs[k]->accumulate;
    // existing values / indices are needed for `+=` and for the unwritten lumped-mass tail: every piece goes both ways
    if (hs[k]->v) dc[k].v = S.add(hs[k]->v + hs[k]->init_k, n, true, true);
    if (hs[k]->r) dc[k].r = S.add(hs[k]->r + hs[k]->init_k, n, true, true);
    if (hs[k]->c) dc[k].c = S.add(hs[k]->c + hs[k]->init_k, n, true, true);
  }
  double* dfint = ((what & PF3_FINT) && fint) ? S.add(fint, size_t(hb->nnodes) * 6, true, true) : nullptr;
  if (S.overflow) return PF3_OK;
  rc = S.upload();
  if (rc) return rc;
  rc = pf3_eval(ctx, &b, what, hs[0] ? &dc[0] : nullptr, hs[1] ? &dc[1] : nullptr, hs[2] ? &dc[2] : nullptr, dfint);
  if (rc) return rc;
  rc = S.download();
  if (rc) return rc;
  *handled = true;
  return PF3_OK;
}

}  // namespace

extern "C" {

// ---- small host-pointer calls of the per-element drop-in classes (pyfe3d_b200/elements.py) ---------------------
int pf3_eval_state_host(pf3_context* ctx, const pf3_batch* hb, double* state_out) {
  int rc = use_device(ctx);
  if (rc) return rc;
  rc = check_batch(hb, 0);
  if (rc) return rc;
  if (!state_out) return PF3_E_BAD_ARG;
  if (hb->ne == 0) return PF3_OK;
  if (batch_host_bytes(hb) + size_t(hb->ne) * PF3_STATE_STRIDE * 8 > PF3_STAGE_BYTES) return PF3_E_CAPACITY;
  Stager S(ctx);
  rc = S.init();
  if (rc) return rc;
  pf3_batch b;
  stage_batch(S, hb, &b);
  double* dout = S.add(state_out, size_t(hb->ne) * PF3_STATE_STRIDE, false, true);
  if (S.overflow) return PF3_E_CAPACITY;
  rc = S.upload();
  if (rc) return rc;
  rc = pf3_eval_state(ctx, &b, dout);
  if (rc) return rc;
  return S.download();
}

int pf3_eval_finte_host(pf3_context* ctx, const pf3_batch* hb, double* finte_out) {
  int rc = use_device(ctx);
  if (rc) return rc;
  rc = check_batch(hb, PF3_FINT);
  if (rc) return rc;
  if (!finte_out) return PF3_E_BAD_ARG;
  if (hb->ne == 0) return PF3_OK;
  const int nn = pf3::kind_nodes(hb->kind);
  if (batch_host_bytes(hb) + size_t(hb->ne) * 6 * nn * 8 > PF3_STAGE_BYTES) return PF3_E_CAPACITY;
  Stager S(ctx);
  rc = S.init();
  if (rc) return rc;
  pf3_batch b;
  stage_batch(S, hb, &b);
  double* dout = S.add(finte_out, size_t(hb->ne) * 6 * nn, false, true);
  if (S.overflow) return PF3_E_CAPACITY;
  rc = S.upload();
  if (rc) return rc;
  rc = pf3_eval_finte(ctx, &b, dout);
  if (rc) return rc;
  return S.download();
}

int pf3_eval_aero_host(pf3_context* ctx, const pf3_batch* hb, int what, const pf3_coo* ka_beta,
                       const pf3_coo* ka_gamma, const pf3_coo* ca) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!hb || hb->ne < 0) return PF3_E_BAD_ARG;
  if (hb->kind != PF3_QUAD4 && hb->kind != PF3_QUAD4R) return PF3_E_UNSUPPORTED;
  if (hb->ne == 0) return PF3_OK;
  if (batch_host_bytes(hb) + size_t(hb->ne) * 144 * 24 * 3 > PF3_STAGE_BYTES) return PF3_E_CAPACITY;
  Stager S(ctx);
  rc = S.init();
  if (rc) return rc;
  pf3_batch b;
  stage_batch(S, hb, &b);
  const pf3_coo* hs[3] = {ka_beta, ka_gamma, ca};
  const int bits[3] = {PF3_KA_BETA, PF3_KA_GAMMA, PF3_CA};
  pf3_coo dc[3];
  const size_t n = size_t(hb->ne) * 144;
  for (int k = 0; k < 3; ++k) {
    std::memset(&dc[k], 0, sizeof(pf3_coo));
    if (!(what & bits[k])) hs[k] = nullptr;
    if (!hs[k]) continue;
    dc[k].accumulate = hs[k]->accumulate;
    if (hs[k]->v) dc[k].v = S.add(hs[k]->v + hs[k]->init_k, n, true, true);
    if (hs[k]->r) dc[k].r = S.add(hs[k]->r + hs[k]->init_k, n, true, true);
    if (hs[k]->c) dc[k].c = S.add(hs[k]->c + hs[k]->init_k, n, true, true);
  }
  if (S.overflow) return PF3_E_CAPACITY;
  rc = S.upload();
  if (rc) return rc;
  rc = pf3_eval_aero(ctx, &b, what, hs[0] ? &dc[0] : nullptr, hs[1] ? &dc[1] : nullptr, hs[2] ? &dc[2] : nullptr);
  if (rc) return rc;
  return S.download();
}

int pf3_quad4_update_BL_host(pf3_context* ctx, int64_t n, const double* xe, double xi, double eta, double* out) {
  int rc = use_device(ctx);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!xe || !out))) return PF3_E_BAD_ARG;
  if (n == 0) return PF3_OK;
  if (size_t(n) * (12 + 264) * 8 + 64 > PF3_STAGE_BYTES) return PF3_E_CAPACITY;
  Stager S(ctx);
  rc = S.init();
  if (rc) return rc;
  const double* dxe = S.add(xe, size_t(n) * 12, true, false);
  double* dout = S.add(out, size_t(n) * 264, false, true);
  rc = S.upload();
  if (rc) return rc;
  rc = pf3_quad4_update_BL(ctx, n, dxe, xi, eta, dout);
  if (rc) return rc;
  return S.download();
}

// Host-pointer convenience: every pointer in host_batch / the pf3_coo structs / fint is a HOST pointer.
int pf3_eval_host(pf3_context* ctx, const pf3_batch* hb, int what, const pf3_coo* kc0, const pf3_coo* kg,
                  const pf3_coo* m, double* fint) {
  int rc = use_device(ctx);
  if (rc) return rc;
  rc = check_batch(hb, what);
  if (rc) return rc;
  if (hb->ne == 0) return PF3_OK;
  {
    bool handled = false;   // small calls: one packed copy each way (see Stager)
    rc = eval_host_staged(ctx, hb, what, kc0, kg, m, fint, &handled);
    if (rc || handled) return rc;
  }
  const int nn = pf3::kind_nodes(hb->kind);
  const int pstride = (hb->kind <= PF3_TRIA3R) ? PF3_SHELLPROP_STRIDE : PF3_BEAMPROP_STRIDE;
  std::vector<void*> owned;
  auto up = [&](const void* h, size_t bytes, const void** d) -> int {
    *d = nullptr;
    if (!h || bytes == 0) return PF3_OK;
    void* p = nullptr;
    PF3_CUDA(cudaMalloc(&p, bytes));
    owned.push_back(p);
    PF3_CUDA(cudaMemcpyAsync(p, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *d = p;
    return PF3_OK;
  };
  auto release = [&]() {
    for (void* p : owned) cudaFree(p);
  };
#define PF3_TRYH(x) do { rc = (x); if (rc) { release(); return rc; } } while (0)
  pf3_batch b = *hb;
  PF3_TRYH(up(hb->conn, size_t(hb->ne) * nn * 8, (const void**)&b.conn));
  PF3_TRYH(up(hb->x, size_t(hb->nnodes) * 3 * 8, (const void**)&b.x));
  PF3_TRYH(up(hb->u, size_t(hb->nnodes) * 6 * 8, (const void**)&b.u));
  PF3_TRYH(up(hb->props, size_t(hb->nprop) * pstride * 8, (const void**)&b.props));
  PF3_TRYH(up(hb->prop_id, size_t(hb->ne) * 4, (const void**)&b.prop_id));
  if (hb->evec) {
    const int width = (hb->kind == PF3_SPRING) ? 6 : 3;
    const size_t cnt = hb->evec_stride ? size_t(hb->ne - 1) * hb->evec_stride + width : width;
    PF3_TRYH(up(hb->evec, cnt * 8, (const void**)&b.evec));
  }
  PF3_TRYH(up(hb->eparam, size_t(hb->ne) * PF3_EPARAM_STRIDE * 8, (const void**)&b.eparam));
  PF3_TRYH(up(hb->state, size_t(hb->ne) * PF3_STATE_STRIDE * 8, (const void**)&b.state));
  struct DevCoo {
    pf3_coo d;
    const pf3_coo* h;
    int64_t n;
  };
  DevCoo dc[3];
  const pf3_coo* hs[3] = {(what & PF3_KC0) ? kc0 : nullptr, (what & (PF3_KG | PF3_KG_STRESS)) ? kg : nullptr,
                          (what & PF3_M) ? m : nullptr};
  for (int k = 0; k < 3; ++k) {
    dc[k].h = hs[k];
    dc[k].n = hb->ne * pf3::kind_sparse_size(hb->kind, k);
    std::memset(&dc[k].d, 0, sizeof(pf3_coo));
    if (!hs[k] || dc[k].n == 0) {
      dc[k].h = nullptr;
      continue;
    }
    dc[k].d.accumulate = hs[k]->accumulate;
    dc[k].d.init_k = 0;
    if (hs[k]->v) {
      void* p = nullptr;
      PF3_TRYH(int(cudaMalloc(&p, size_t(dc[k].n) * 8)));
      owned.push_back(p);
      dc[k].d.v = (double*)p;
      // existing values are needed both for `+=` and for the unwritten lumped-mass tail
      PF3_TRYH(int(cudaMemcpyAsync(p, hs[k]->v + hs[k]->init_k, size_t(dc[k].n) * 8, cudaMemcpyHostToDevice, ctx->stream)));
    }
    if (hs[k]->r) {
      void* p = nullptr;
      PF3_TRYH(int(cudaMalloc(&p, size_t(dc[k].n) * 8)));
      owned.push_back(p);
      dc[k].d.r = (int64_t*)p;
      PF3_TRYH(int(cudaMemcpyAsync(p, hs[k]->r + hs[k]->init_k, size_t(dc[k].n) * 8, cudaMemcpyHostToDevice, ctx->stream)));
    }
    if (hs[k]->c) {
      void* p = nullptr;
      PF3_TRYH(int(cudaMalloc(&p, size_t(dc[k].n) * 8)));
      owned.push_back(p);
      dc[k].d.c = (int64_t*)p;
      PF3_TRYH(int(cudaMemcpyAsync(p, hs[k]->c + hs[k]->init_k, size_t(dc[k].n) * 8, cudaMemcpyHostToDevice, ctx->stream)));
    }
  }
  double* dfint = nullptr;
  if ((what & PF3_FINT) && fint) {
    PF3_TRYH(up(fint, size_t(hb->nnodes) * 6 * 8, (const void**)&dfint));
  }
  PF3_TRYH(pf3_eval(ctx, &b, what, dc[0].h ? &dc[0].d : nullptr, dc[1].h ? &dc[1].d : nullptr,
                    dc[2].h ? &dc[2].d : nullptr, dfint));
  for (int k = 0; k < 3; ++k) {
    if (!dc[k].h) continue;
    if (dc[k].d.v) PF3_TRYH(int(cudaMemcpyAsync(dc[k].h->v + dc[k].h->init_k, dc[k].d.v, size_t(dc[k].n) * 8, cudaMemcpyDeviceToHost, ctx->stream)));
    if (dc[k].d.r) PF3_TRYH(int(cudaMemcpyAsync(dc[k].h->r + dc[k].h->init_k, dc[k].d.r, size_t(dc[k].n) * 8, cudaMemcpyDeviceToHost, ctx->stream)));
    if (dc[k].d.c) PF3_TRYH(int(cudaMemcpyAsync(dc[k].h->c + dc[k].h->init_k, dc[k].d.c, size_t(dc[k].n) * 8, cudaMemcpyDeviceToHost, ctx->stream)));
  }
  if (dfint) PF3_TRYH(int(cudaMemcpyAsync(fint, dfint, size_t(hb->nnodes) * 6 * 8, cudaMemcpyDeviceToHost, ctx->stream)));
  PF3_TRYH(int(cudaStreamSynchronize(ctx->stream)));
#undef PF3_TRYH
  release();
  return PF3_OK;
}

}  // extern "C"
