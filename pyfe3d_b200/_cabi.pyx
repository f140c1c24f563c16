# cython: language_level=3, boundscheck=False, wraparound=False
"""Thin Cython layer over the extern "C" ABI of libpyfe3d_b200.so
(include/pyfe3d_b200.h).  Nothing here computes: it packs plain pointers and
sizes into the C structs and turns status codes into exceptions.

Pointers are passed as integers (``tensor.data_ptr()`` for device memory,
``ndarray.ctypes.data`` for the ``*_host`` entry points); 0 means NULL.
"""
from libc.stdint cimport int32_t, int64_t, uintptr_t
from libc.string cimport memset

cdef extern from "pyfe3d_b200.h":
    ctypedef struct pf3_context:
        pass
    ctypedef struct pf3_plan:
        pass
    ctypedef struct pf3_batch:
        int32_t kind
        int32_t mtype
        int64_t ne
        int64_t nnodes
        const int64_t* conn
        const double* x
        const double* u
        const double* props
        const int32_t* prop_id
        int64_t nprop
        const double* evec
        int32_t evec_stride
        const double* eparam
        const double* state
        int32_t state_flags
        double stress[3]
    ctypedef struct pf3_coo:
        int64_t* r
        int64_t* c
        double* v
        int64_t init_k
        int32_t accumulate
    int pf3_version()
    const char* pf3_error_string(int code)
    int pf3_device_count(int* n)
    int pf3_create(int device, pf3_context** ctx)
    int pf3_destroy(pf3_context* ctx)
    int pf3_set_stream(pf3_context* ctx, void* s)
    int pf3_synchronize(pf3_context* ctx)
    int pf3_launch_count(pf3_context* ctx, int64_t* n)
    int pf3_num_nodes(int kind)
    int pf3_sparse_size(int kind, int matrix)
    int pf3_written_size(int kind, int matrix, int mtype)
    int pf3_eval(pf3_context*, const pf3_batch*, int what, const pf3_coo*, const pf3_coo*, const pf3_coo*, double* fint) nogil
    int pf3_eval_host(pf3_context*, const pf3_batch*, int what, const pf3_coo*, const pf3_coo*, const pf3_coo*, double* fint) nogil
    int pf3_fill_indices(pf3_context*, int kind, int matrix, int mtype, int64_t ne, const int64_t* conn,
                         int64_t init_k, int64_t* r, int64_t* c) nogil
    int pf3_eval_state(pf3_context*, const pf3_batch*, double* out) nogil
    int pf3_eval_state_host(pf3_context*, const pf3_batch*, double* out) nogil
    int pf3_eval_finte_host(pf3_context*, const pf3_batch*, double* out) nogil
    int pf3_eval_aero_host(pf3_context*, const pf3_batch*, int what, const pf3_coo*, const pf3_coo*, const pf3_coo*) nogil
    int pf3_quad4_update_BL_host(pf3_context*, int64_t n, const double* xe, double xi, double eta, double* out) nogil
    int pf3_eval_finte(pf3_context*, const pf3_batch*, double* out) nogil
    int pf3_plan_create(pf3_context*, int matrix, int64_t nnodes, int ngroups, const pf3_batch* groups,
                        const int64_t* coo_offsets, int64_t node_begin, int64_t node_end, pf3_plan** plan) nogil
    int pf3_plan_create_coo(pf3_context*, int64_t n, int64_t nnz, const int64_t* r, const int64_t* c, pf3_plan**) nogil
    int pf3_plan_destroy(pf3_plan*)
    int pf3_plan_nnz(const pf3_plan*, int64_t*)
    int pf3_plan_nrows(const pf3_plan*, int64_t*)
    int pf3_plan_pattern(pf3_context*, const pf3_plan*, int64_t* indptr, int64_t* indices) nogil
    int pf3_plan_nblocks(const pf3_plan*, int64_t*)
    int pf3_plan_fint(pf3_context*, const pf3_plan*, int group, const pf3_batch*, double* fint) nogil
    int pf3_eval_aero(pf3_context*, const pf3_batch*, int what, const pf3_coo*, const pf3_coo*, const pf3_coo*) nogil
    int pf3_laminate_props(pf3_context*, int64_t nrows, int nplies, const double* thetadeg, int64_t theta_stride,
                           const double* plyt, int64_t plyt_stride, const double* lamina, int64_t lamina_stride,
                           const double* offset, int64_t offset_stride, int calc_scf, double* props_out) nogil
    int pf3_lamination_parameter_props(pf3_context*, int64_t nrows, const double* thickness, int64_t thickness_stride,
                                       const double* invariants, int64_t invariants_stride, const double* lp,
                                       int64_t lp_stride, const double* rho, int64_t rho_stride, int var_mask,
                                       int grad_complete, double* props_out, double* grad_out) nogil
    int pf3_plan_spmv(pf3_context*, const pf3_plan*, const double* vals, const unsigned char* free_dof,
                      const double* x, double* y) nogil
    int pf3_plan_diagonal(pf3_context*, const pf3_plan*, const double* vals, double* diag) nogil
    ctypedef struct pf3_cg_info:
        int32_t iterations
        int32_t status
        double residual
        double bnorm
    int pf3_plan_cg(pf3_context*, int nops, const pf3_plan* const* plans, const double* const* vals,
                    const double* coefs, const unsigned char* free_dof, const double* b, double* x, int use_x0,
                    double rtol, double atol, int maxiter, int flags, pf3_cg_info* info) nogil
    size_t pf3_cg_shard_work_bytes() nogil
    int pf3_cg_shard_dot(pf3_context*, int64_t n, const double* p, const double* ap, double* sc, void* work) nogil
    int pf3_cg_shard_update(pf3_context*, int64_t n, const double* p, const double* ap, const double* minv, double* x,
                            double* r, double* sc, void* work) nogil
    int pf3_cg_shard_dir(pf3_context*, int64_t n, const double* r, const double* minv, double* p, double* sc,
                         void* work) nogil
    int pf3_plan_spmv_scaled(pf3_context*, const pf3_plan*, const double* vals, const unsigned char* free_dof,
                             const double* scale, const double* x, double* y) nogil
    int pf3_csr_compact_symbolic(pf3_context*, int64_t nrows, int64_t ncols, const int64_t* indptr,
                                 const int64_t* indices, const unsigned char* free_dof, int flags, int64_t row0,
                                 int64_t* colmap, int64_t* out_indptr, int64_t* nkeep, int64_t* nnz) nogil
    int pf3_csr_compact_fill(pf3_context*, int64_t nrows, int64_t ncols, const int64_t* indptr,
                             const int64_t* indices, const double* vals, const unsigned char* free_dof, int flags,
                             int64_t row0, const int64_t* colmap, const int64_t* out_indptr, int64_t* out_indices,
                             double* out_vals) nogil
    int pf3_quad4_update_BL(pf3_context*, int64_t n, const double* xe, double xi, double eta, double* out) nogil
    int pf3_eval_assemble(pf3_context*, const pf3_batch*, const pf3_plan*, int what, const pf3_coo*, const pf3_coo*,
                          const pf3_coo*, double*, double*, double*) nogil
    int pf3_eval_assemble_host(pf3_context*, const pf3_batch*, const pf3_plan*, int what, const pf3_coo*,
                               const pf3_coo*, const pf3_coo*, double*, double*, double*) nogil
    int pf3_eval_assemble_group(pf3_context*, const pf3_batch*, const pf3_plan*, int group, int what, const pf3_coo*,
                                const pf3_coo*, const pf3_coo*, double*, double*, double*) nogil
    int pf3_plan_assemble_add(pf3_context*, const pf3_plan*, const double* coo_v, double* csr_v, int skip_group) nogil
    int pf3_plan_assemble(pf3_context*, const pf3_plan*, const double* coo_v, double* csr_v) nogil
    int pf3_spmv_csr(pf3_context*, int64_t nrows, const int64_t* indptr, const int64_t* indices,
                     const double* vals, const double* x, double* y) nogil
    int pf3_spmv_csr_masked(pf3_context*, int64_t nrows, const int64_t* indptr, const int64_t* indices,
                            const double* vals, const unsigned char* free_dof, const double* x, double* y) nogil
    int pf3_csr_diagonal(pf3_context*, int64_t nrows, const int64_t* indptr, const int64_t* indices,
                         const double* vals, int64_t row0, double* diag) nogil
    int pf3_memcpy_h2d(pf3_context*, void* dst, const void* src, size_t n) nogil
    int pf3_memcpy_d2h(pf3_context*, void* dst, const void* src, size_t n) nogil

KC0, KG, KG_STRESS, M, FINT = 1, 2, 4, 8, 16
KA_BETA, KA_GAMMA, CA = 32, 64, 128
MAT_KC0, MAT_KG, MAT_M = 0, 1, 2
MAT_KA_BETA, MAT_KA_GAMMA, MAT_CA = 3, 4, 5
QUAD4, QUAD4R, TRIA3R, BEAMC, BEAMLR, TRUSS, SPRING = range(7)
SHELLPROP_STRIDE, BEAMPROP_STRIDE, EPARAM_STRIDE, STATE_STRIDE = 32, 16, 12, 50
STATE_REFRESH_XE, STATE_REFRESH_UE = 1, 2


class Pf3Error(RuntimeError):
    pass


cdef int _check(int rc) except -1:
    if rc == 0:
        return 0
    msg = pf3_error_string(rc).decode()
    if rc == -1:
        raise ValueError(msg)
    raise Pf3Error("%s (code %d)" % (msg, rc))


def cg_shard_work_bytes():
    return pf3_cg_shard_work_bytes()


def version():
    return pf3_version()


def device_count():
    cdef int n = 0
    pf3_device_count(&n)
    return n


def num_nodes(int kind):
    return pf3_num_nodes(kind)


def sparse_size(int kind, int matrix):
    return pf3_sparse_size(kind, matrix)


def written_size(int kind, int matrix, int mtype=0):
    return pf3_written_size(kind, matrix, mtype)


cdef class Batch:
    """pf3_batch; every pointer argument is an integer address (0 = NULL)."""
    cdef pf3_batch b

    def __init__(self, int kind, int64_t ne, int64_t nnodes, uintptr_t conn=0, uintptr_t x=0, uintptr_t u=0,
                 uintptr_t props=0, uintptr_t prop_id=0, int64_t nprop=0, uintptr_t evec=0, int evec_stride=0,
                 uintptr_t eparam=0, uintptr_t state=0, int mtype=0, stress=(0., 0., 0.), int state_flags=0):
        memset(&self.b, 0, sizeof(pf3_batch))
        self.b.kind = kind
        self.b.mtype = mtype
        self.b.ne = ne
        self.b.nnodes = nnodes
        self.b.conn = <const int64_t*>conn
        self.b.x = <const double*>x
        self.b.u = <const double*>u
        self.b.props = <const double*>props
        self.b.prop_id = <const int32_t*>prop_id
        self.b.nprop = nprop
        self.b.evec = <const double*>evec
        self.b.evec_stride = evec_stride
        self.b.eparam = <const double*>eparam
        self.b.state = <const double*>state
        self.b.state_flags = state_flags
        self.b.stress[0] = stress[0]
        self.b.stress[1] = stress[1]
        self.b.stress[2] = stress[2]


cdef class Coo:
    """pf3_coo destination (r, c, v addresses; 0 = not requested)."""
    cdef pf3_coo c

    def __init__(self, uintptr_t r=0, uintptr_t c=0, uintptr_t v=0, int64_t init_k=0, int accumulate=0):
        self.c.r = <int64_t*>r
        self.c.c = <int64_t*>c
        self.c.v = <double*>v
        self.c.init_k = init_k
        self.c.accumulate = accumulate


cdef class Context:
    cdef pf3_context* ctx
    cdef public int device

    def __cinit__(self, int device=0):
        self.ctx = NULL
        self.device = device
        _check(pf3_create(device, &self.ctx))

    def __dealloc__(self):
        if self.ctx != NULL:
            pf3_destroy(self.ctx)
            self.ctx = NULL

    def set_stream(self, uintptr_t stream):
        _check(pf3_set_stream(self.ctx, <void*>stream))

    def synchronize(self):
        _check(pf3_synchronize(self.ctx))

    def launch_count(self):
        cdef int64_t n = 0
        _check(pf3_launch_count(self.ctx, &n))
        return n

    def eval(self, Batch b, int what, Coo kc0=None, Coo kg=None, Coo m=None, uintptr_t fint=0, bint host=False):
        cdef const pf3_coo* p0 = &kc0.c if kc0 is not None else NULL
        cdef const pf3_coo* p1 = &kg.c if kg is not None else NULL
        cdef const pf3_coo* p2 = &m.c if m is not None else NULL
        cdef int rc
        with nogil:
            if host:
                rc = pf3_eval_host(self.ctx, &b.b, what, p0, p1, p2, <double*>fint)
            else:
                rc = pf3_eval(self.ctx, &b.b, what, p0, p1, p2, <double*>fint)
        _check(rc)

    def laminate_props(self, int64_t nrows, int nplies, uintptr_t theta, int64_t theta_stride, uintptr_t plyt,
                       int64_t plyt_stride, uintptr_t lamina, int64_t lamina_stride, uintptr_t offset,
                       int64_t offset_stride, int calc_scf, uintptr_t out):
        cdef int rc
        with nogil:
            rc = pf3_laminate_props(self.ctx, nrows, nplies, <const double*>theta, theta_stride, <const double*>plyt,
                                    plyt_stride, <const double*>lamina, lamina_stride, <const double*>offset,
                                    offset_stride, calc_scf, <double*>out)
        _check(rc)

    def lamination_parameter_props(self, int64_t nrows, uintptr_t thickness, int64_t thickness_stride,
                                   uintptr_t invariants, int64_t invariants_stride, uintptr_t lp, int64_t lp_stride,
                                   uintptr_t rho, int64_t rho_stride, int var_mask, int grad_complete, uintptr_t out,
                                   uintptr_t grad):
        cdef int rc
        with nogil:
            rc = pf3_lamination_parameter_props(self.ctx, nrows, <const double*>thickness, thickness_stride,
                                                <const double*>invariants, invariants_stride, <const double*>lp,
                                                lp_stride, <const double*>rho, rho_stride, var_mask, grad_complete,
                                                <double*>out, <double*>grad)
        _check(rc)

    def eval_aero(self, Batch b, int what, Coo ka_beta=None, Coo ka_gamma=None, Coo ca=None, bint host=False):
        cdef const pf3_coo* p0 = &ka_beta.c if ka_beta is not None else NULL
        cdef const pf3_coo* p1 = &ka_gamma.c if ka_gamma is not None else NULL
        cdef const pf3_coo* p2 = &ca.c if ca is not None else NULL
        cdef int rc
        with nogil:
            if host:
                rc = pf3_eval_aero_host(self.ctx, &b.b, what, p0, p1, p2)
            else:
                rc = pf3_eval_aero(self.ctx, &b.b, what, p0, p1, p2)
        _check(rc)

    def fill_indices(self, int kind, int matrix, int mtype, int64_t ne, uintptr_t conn, int64_t init_k,
                     uintptr_t r, uintptr_t c):
        cdef int rc
        with nogil:
            rc = pf3_fill_indices(self.ctx, kind, matrix, mtype, ne, <const int64_t*>conn, init_k,
                                  <int64_t*>r, <int64_t*>c)
        _check(rc)

    def eval_state(self, Batch b, uintptr_t out, bint host=False):
        cdef int rc
        with nogil:
            if host:
                rc = pf3_eval_state_host(self.ctx, &b.b, <double*>out)
            else:
                rc = pf3_eval_state(self.ctx, &b.b, <double*>out)
        _check(rc)

    def quad4_update_BL(self, int64_t n, uintptr_t xe, double xi, double eta, uintptr_t out, bint host=False):
        cdef int rc
        with nogil:
            if host:
                rc = pf3_quad4_update_BL_host(self.ctx, n, <const double*>xe, xi, eta, <double*>out)
            else:
                rc = pf3_quad4_update_BL(self.ctx, n, <const double*>xe, xi, eta, <double*>out)
        _check(rc)

    def eval_finte(self, Batch b, uintptr_t out, bint host=False):
        cdef int rc
        with nogil:
            if host:
                rc = pf3_eval_finte_host(self.ctx, &b.b, <double*>out)
            else:
                rc = pf3_eval_finte(self.ctx, &b.b, <double*>out)
        _check(rc)

    def spmv_csr(self, int64_t nrows, uintptr_t indptr, uintptr_t indices, uintptr_t vals, uintptr_t x, uintptr_t y):
        cdef int rc
        with nogil:
            rc = pf3_spmv_csr(self.ctx, nrows, <const int64_t*>indptr, <const int64_t*>indices,
                              <const double*>vals, <const double*>x, <double*>y)
        _check(rc)

    def spmv_csr_masked(self, int64_t nrows, uintptr_t indptr, uintptr_t indices, uintptr_t vals, uintptr_t free_dof,
                        uintptr_t x, uintptr_t y):
        cdef int rc
        with nogil:
            rc = pf3_spmv_csr_masked(self.ctx, nrows, <const int64_t*>indptr, <const int64_t*>indices,
                                     <const double*>vals, <const unsigned char*>free_dof, <const double*>x, <double*>y)
        _check(rc)

    def csr_diagonal(self, int64_t nrows, uintptr_t indptr, uintptr_t indices, uintptr_t vals, int64_t row0,
                     uintptr_t diag):
        cdef int rc
        with nogil:
            rc = pf3_csr_diagonal(self.ctx, nrows, <const int64_t*>indptr, <const int64_t*>indices,
                                  <const double*>vals, row0, <double*>diag)
        _check(rc)

    def csr_compact_symbolic(self, int64_t nrows, int64_t ncols, uintptr_t indptr, uintptr_t indices,
                             uintptr_t free_dof, int flags, int64_t row0, uintptr_t colmap, uintptr_t out_indptr):
        cdef int rc
        cdef int64_t nkeep = 0, nnz = 0
        with nogil:
            rc = pf3_csr_compact_symbolic(self.ctx, nrows, ncols, <const int64_t*>indptr, <const int64_t*>indices,
                                          <const unsigned char*>free_dof, flags, row0, <int64_t*>colmap,
                                          <int64_t*>out_indptr, &nkeep, &nnz)
        _check(rc)
        return nkeep, nnz

    def csr_compact_fill(self, int64_t nrows, int64_t ncols, uintptr_t indptr, uintptr_t indices, uintptr_t vals,
                         uintptr_t free_dof, int flags, int64_t row0, uintptr_t colmap, uintptr_t out_indptr,
                         uintptr_t out_indices, uintptr_t out_vals):
        cdef int rc
        with nogil:
            rc = pf3_csr_compact_fill(self.ctx, nrows, ncols, <const int64_t*>indptr, <const int64_t*>indices,
                                      <const double*>vals, <const unsigned char*>free_dof, flags, row0,
                                      <const int64_t*>colmap, <const int64_t*>out_indptr, <int64_t*>out_indices,
                                      <double*>out_vals)
        _check(rc)

    def plan_cg(self, list plans, list vals, list coefs, uintptr_t free_dof, uintptr_t b, uintptr_t x, bint use_x0,
                double rtol, double atol, int maxiter, int flags):
        """pf3_plan_cg: returns (iterations, status, residual norm, |b|)."""
        cdef int n = len(plans)
        if n < 1 or n > 8 or len(vals) != n or len(coefs) != n:
            raise ValueError("1..8 (plan, values, coefficient) triples")
        cdef const pf3_plan* pp[8]
        cdef const double* vv[8]
        cdef double cc[8]
        cdef int i
        for i in range(n):
            pp[i] = (<Plan>plans[i]).plan
            vv[i] = <const double*><uintptr_t>vals[i]
            cc[i] = coefs[i]
        cdef pf3_cg_info info
        cdef int rc
        with nogil:
            rc = pf3_plan_cg(self.ctx, n, pp, vv, cc, <const unsigned char*>free_dof, <const double*>b, <double*>x,
                             use_x0, rtol, atol, maxiter, flags, &info)
        _check(rc)
        return info.iterations, info.status, info.residual, info.bnorm

    def cg_shard_dot(self, int64_t n, uintptr_t p, uintptr_t ap, uintptr_t sc, uintptr_t work):
        cdef int rc
        with nogil:
            rc = pf3_cg_shard_dot(self.ctx, n, <const double*>p, <const double*>ap, <double*>sc, <void*>work)
        _check(rc)

    def cg_shard_update(self, int64_t n, uintptr_t p, uintptr_t ap, uintptr_t minv, uintptr_t x, uintptr_t r,
                        uintptr_t sc, uintptr_t work):
        cdef int rc
        with nogil:
            rc = pf3_cg_shard_update(self.ctx, n, <const double*>p, <const double*>ap, <const double*>minv, <double*>x,
                                     <double*>r, <double*>sc, <void*>work)
        _check(rc)

    def cg_shard_dir(self, int64_t n, uintptr_t r, uintptr_t minv, uintptr_t p, uintptr_t sc, uintptr_t work):
        cdef int rc
        with nogil:
            rc = pf3_cg_shard_dir(self.ctx, n, <const double*>r, <const double*>minv, <double*>p, <double*>sc,
                                  <void*>work)
        _check(rc)

    def memcpy_h2d(self, uintptr_t dst, uintptr_t src, size_t n):
        cdef int rc
        with nogil:
            rc = pf3_memcpy_h2d(self.ctx, <void*>dst, <const void*>src, n)
        _check(rc)

    def memcpy_d2h(self, uintptr_t dst, uintptr_t src, size_t n):
        cdef int rc
        with nogil:
            rc = pf3_memcpy_d2h(self.ctx, <void*>dst, <const void*>src, n)
        _check(rc)


cdef class Plan:
    cdef pf3_plan* plan
    cdef Context owner

    def __cinit__(self):
        self.plan = NULL

    def __dealloc__(self):
        if self.plan != NULL:
            pf3_plan_destroy(self.plan)
            self.plan = NULL

    @staticmethod
    def structured(Context ctx, int matrix, int64_t nnodes, list batches, list coo_offsets,
                   int64_t node_begin, int64_t node_end):
        cdef int n = len(batches)
        if n == 0 or n > 8 or len(coo_offsets) != n:
            raise ValueError("1..8 element groups with one COO offset each")
        cdef pf3_batch groups[8]
        cdef int64_t offs[8]
        cdef int i
        for i in range(n):
            groups[i] = (<Batch>batches[i]).b
            offs[i] = coo_offsets[i]
        cdef Plan p = Plan()
        p.owner = ctx
        cdef int rc
        with nogil:
            rc = pf3_plan_create(ctx.ctx, matrix, nnodes, n, groups, offs, node_begin, node_end, &p.plan)
        _check(rc)
        return p

    @staticmethod
    def from_coo(Context ctx, int64_t n, int64_t nnz, uintptr_t r, uintptr_t c):
        cdef Plan p = Plan()
        p.owner = ctx
        cdef int rc
        with nogil:
            rc = pf3_plan_create_coo(ctx.ctx, n, nnz, <const int64_t*>r, <const int64_t*>c, &p.plan)
        _check(rc)
        return p

    @property
    def nnz(self):
        cdef int64_t n = 0
        _check(pf3_plan_nnz(self.plan, &n))
        return n

    @property
    def nrows(self):
        cdef int64_t n = 0
        _check(pf3_plan_nrows(self.plan, &n))
        return n

    @property
    def nblocks(self):
        cdef int64_t n = 0
        _check(pf3_plan_nblocks(self.plan, &n))
        return n

    def eval_assemble(self, Batch b, int what, Coo kc0=None, Coo kg=None, Coo m=None, uintptr_t csr_kc0=0,
                      uintptr_t csr_kg=0, uintptr_t csr_m=0):
        cdef const pf3_coo* p0 = &kc0.c if kc0 is not None else NULL
        cdef const pf3_coo* p1 = &kg.c if kg is not None else NULL
        cdef const pf3_coo* p2 = &m.c if m is not None else NULL
        cdef int rc
        with nogil:
            rc = pf3_eval_assemble(self.owner.ctx, &b.b, self.plan, what, p0, p1, p2, <double*>csr_kc0,
                                   <double*>csr_kg, <double*>csr_m)
        _check(rc)

    def eval_assemble_host(self, Batch b, int what, Coo kc0=None, Coo kg=None, Coo m=None, uintptr_t csr_kc0=0,
                           uintptr_t csr_kg=0, uintptr_t csr_m=0):
        cdef const pf3_coo* p0 = &kc0.c if kc0 is not None else NULL
        cdef const pf3_coo* p1 = &kg.c if kg is not None else NULL
        cdef const pf3_coo* p2 = &m.c if m is not None else NULL
        cdef int rc
        with nogil:
            rc = pf3_eval_assemble_host(self.owner.ctx, &b.b, self.plan, what, p0, p1, p2, <double*>csr_kc0,
                                        <double*>csr_kg, <double*>csr_m)
        _check(rc)

    def eval_assemble_group(self, Batch b, int group, int what, Coo kc0=None, Coo kg=None, Coo m=None,
                            uintptr_t csr_kc0=0, uintptr_t csr_kg=0, uintptr_t csr_m=0):
        cdef const pf3_coo* p0 = &kc0.c if kc0 is not None else NULL
        cdef const pf3_coo* p1 = &kg.c if kg is not None else NULL
        cdef const pf3_coo* p2 = &m.c if m is not None else NULL
        cdef int rc
        with nogil:
            rc = pf3_eval_assemble_group(self.owner.ctx, &b.b, self.plan, group, what, p0, p1, p2, <double*>csr_kc0,
                                         <double*>csr_kg, <double*>csr_m)
        _check(rc)

    def assemble_add(self, uintptr_t coo_v, uintptr_t csr_v, int skip_group):
        cdef int rc
        with nogil:
            rc = pf3_plan_assemble_add(self.owner.ctx, self.plan, <const double*>coo_v, <double*>csr_v, skip_group)
        _check(rc)

    def fint(self, int group, Batch b, uintptr_t fint):
        cdef int rc
        with nogil:
            rc = pf3_plan_fint(self.owner.ctx, self.plan, group, &b.b, <double*>fint)
        _check(rc)

    def spmv(self, uintptr_t vals, uintptr_t free_dof, uintptr_t x, uintptr_t y):
        cdef int rc
        with nogil:
            rc = pf3_plan_spmv(self.owner.ctx, self.plan, <const double*>vals, <const unsigned char*>free_dof,
                               <const double*>x, <double*>y)
        _check(rc)

    def diagonal(self, uintptr_t vals, uintptr_t diag):
        cdef int rc
        with nogil:
            rc = pf3_plan_diagonal(self.owner.ctx, self.plan, <const double*>vals, <double*>diag)
        _check(rc)

    def spmv_scaled(self, uintptr_t vals, uintptr_t free_dof, uintptr_t scale, uintptr_t x, uintptr_t y):
        cdef int rc
        with nogil:
            rc = pf3_plan_spmv_scaled(self.owner.ctx, self.plan, <const double*>vals, <const unsigned char*>free_dof,
                                      <const double*>scale, <const double*>x, <double*>y)
        _check(rc)

    def pattern(self, uintptr_t indptr, uintptr_t indices):
        cdef int rc
        with nogil:
            rc = pf3_plan_pattern(self.owner.ctx, self.plan, <int64_t*>indptr, <int64_t*>indices)
        _check(rc)

    def assemble(self, uintptr_t coo_v, uintptr_t csr_v):
        cdef int rc
        with nogil:
            rc = pf3_plan_assemble(self.owner.ctx, self.plan, <const double*>coo_v, <double*>csr_v)
        _check(rc)
