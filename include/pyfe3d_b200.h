/*
 * pyfe3d_b200 — C ABI of the B200-native element-evaluation + sparse-assembly path.
 *
 * This is the drop-in boundary for ONE hot path of saullocastro/pyfe3d
 * (SURVEY.md §8): the per-element update_KC0 / update_KG / update_KG_given_stress /
 * update_M / update_fint methods of Quad4, Quad4R, Tria3R, BeamC, BeamLR, Truss and
 * Spring, and the COO -> CSR sum-duplicates step the reference leaves to scipy.
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference checkout).
 *
 * Conventions
 *  - plain C: pointers + sizes, no C++/torch types.  All array arguments are DEVICE
 *    pointers unless the function name ends in `_host`.
 *  - int return: 0 = ok, <0 = PF3_E_* (bad argument / state), >0 = cudaError_t.
 *    pf3_error_string(code) explains either.  There is no CPU fallback: every
 *    compute call needs a CUDA device and fails with PF3_E_NO_DEVICE otherwise.
 *  - x  : float64[3*nnodes], node p at x[3p..3p+2]          (quad4.pyx:526-537)
 *    u  : float64[6*nnodes], u v w rx ry rz per node        (quad4.pyx:634-639)
 *    conn: int64[ne*nn] node POSITIONS p; the reference's c_a attribute is 6*p
 *  - COO blocks: element e of a batch owns entries [init_k + e*SIZE, +SIZE) in the
 *    reference's own order (row-major over node_i, dof_i, node_j, dof_j restricted
 *    to the per-matrix mask).  Offsets are 64-bit (the reference
 *    overflows its C-int init_k_* beyond 3 728 270 Quad4 elements, quad4.pyx:453).
 */
#ifndef PYFE3D_B200_H
#define PYFE3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define PF3_VERSION 100

/* ---- status codes ------------------------------------------------------- */
#define PF3_OK 0
#define PF3_E_BAD_ARG (-1)
#define PF3_E_NO_DEVICE (-2)
#define PF3_E_UNSUPPORTED (-3)
#define PF3_E_CAPACITY (-4)

/* ---- element kinds (pyfe3d/__init__.py:11-17) --------------------------- */
#define PF3_QUAD4 0   /* pyfe3d/quad4.pyx  */
#define PF3_QUAD4R 1  /* pyfe3d/quad4r.pyx */
#define PF3_TRIA3R 2  /* pyfe3d/tria3r.pyx */
#define PF3_BEAMC 3   /* pyfe3d/beamc.pyx  */
#define PF3_BEAMLR 4  /* pyfe3d/beamlr.pyx */
#define PF3_TRUSS 5   /* pyfe3d/truss.pyx  */
#define PF3_SPRING 6  /* pyfe3d/spring.pyx */
#define PF3_NKINDS 7

/* ---- which matrices one pf3_eval call produces (bitmask) ---------------- */
#define PF3_KC0 1        /* update_KC0              e.g. quad4.pyx:1204 */
#define PF3_KG 2         /* update_KG (from u)      e.g. quad4.pyx:1365 */
#define PF3_KG_STRESS 4  /* update_KG_given_stress  e.g. quad4.pyx:2259 (shells only) */
#define PF3_M 8          /* update_M                e.g. quad4.pyx:3083 */
#define PF3_FINT 16      /* update_fint             e.g. quad4.pyx:1316 */

/* piston-theory aerodynamic matrices of Quad4 / Quad4R (pf3_eval_aero) */
#define PF3_KA_BETA 32   /* update_KA_beta   quad4.pyx:9491,  quad4r.pyx:12789 */
#define PF3_KA_GAMMA 64  /* update_KA_gamma  quad4.pyx:10312, quad4r.pyx:13605 */
#define PF3_CA 128       /* update_CA        quad4.pyx:11115, quad4r.pyx:14403 */

/* ---- matrix ids for sizes / index fill / assembly plans ----------------- */
#define PF3_MAT_KC0 0
#define PF3_MAT_KG 1
#define PF3_MAT_M 2
#define PF3_MAT_KA_BETA 3  /* KA_BETA_SPARSE_SIZE / KA_GAMMA_SPARSE_SIZE / CA_SPARSE_SIZE = 144 (quad4.pyx:150-152): */
#define PF3_MAT_KA_GAMMA 4 /* the translational 12x12 mask of KG                                                    */
#define PF3_MAT_CA 5

/* ---- property tables ---------------------------------------------------- */
/* ShellProp scalars read by the elements (shellprop.pxd:38-46): row layout
 *  0..5  A11 A12 A16 A22 A26 A66 | 6..11 B.. | 12..17 D.. | 18..20 E44 E45 E55
 *  21 scf_k13 | 22 scf_k23 | 23 h | 24 intrho | 25 intrhoz | 26 intrhoz2 | 27..31 pad */
#define PF3_SHELLPROP_STRIDE 32
/* BeamProp (beamprop.pxd:1-3): A E G Iyy Izz Iyz J Ay Az intrho intrhoy intrhoz
 *  intrhoy2 intrhoz2 intrhoyz | pad */
#define PF3_BEAMPROP_STRIDE 16
/* per-element parameters (attributes / kwargs of the reference objects):
 *  shells : 0 K6ROT (default 100, quad4.pyx:481) | 1 alpha_shear_locking (0.7,
 *           tria3r.pyx:263) | 2..6 hgfactor_u,v,w,rx,ry (1.0, quad4r.pyx:1151-1155)
 *           | 7 != 0: slots 8..11 hold the element's PREVIOUS m11 m12 m21 m22, which
 *           update_rotation_matrix leaves untouched when xmat is null or normal to the
 *           element (sticky state, quad4.pyx:588,598); 0: start from identity
 *  spring : 0..5 kxe kye kze krxe krye krze (spring.pyx:117-124) */
#define PF3_EPARAM_STRIDE 12
/* per-element explicit state used by the per-element drop-in classes, which keep
 * r11..r33 / m11..m22 / probe.xe / area|length / probe.ue on the host object exactly
 * like the reference (quad4.pyx:450-459).  Layout (doubles):
 *  0..8 R row-major | 9..12 m11 m12 m21 m22 | 13 area or length | 14..25 xe (3*nn)
 *  | 26..49 ue (6*nn) */
#define PF3_STATE_STRIDE 50
#define PF3_STATE_REFRESH_XE 1
#define PF3_STATE_REFRESH_UE 2

typedef struct pf3_context pf3_context; /* device, stream, scratch; one host thread at a time */
typedef struct pf3_plan pf3_plan;       /* symbolic assembly result (pattern + gather map)   */

/* One batch = all elements of ONE kind.  Replaces the Python loop
 *   for e in elements: e.update_rotation_matrix(...); e.update_probe_xe(x);
 *                      e.update_probe_ue(u); e.update_KC0(...); ...
 * (tests/test_quad4_static_point_load.py:53-78). */
typedef struct pf3_batch {
  int32_t kind;           /* PF3_QUAD4 ... */
  int32_t mtype;          /* update_M mtype: shells 0 consistent,1 reduced,2 lumped; lines 0,1 */
  int64_t ne;
  int64_t nnodes;
  const int64_t* conn;    /* [ne*nn] */
  const double* x;        /* [3*nnodes]; unused by PF3_SPRING and when state != NULL */
  const double* u;        /* [6*nnodes] or NULL; needed by PF3_KG and PF3_FINT */
  const double* props;    /* [nprop*stride]; NULL for PF3_SPRING */
  const int32_t* prop_id; /* [ne] or NULL (all elements use row 0) */
  int64_t nprop;
  const double* evec;     /* shells: material direction xmat (update_rotation_matrix kwargs,
                             quad4.pyx:491); beams: vxy (beamc.pyx:158); spring: xi xj xk vxyi
                             vxyj vxyk (spring.pyx:147).  NULL => shells: no material axis */
  int32_t evec_stride;    /* doubles between elements in evec; 0 = one vector for all */
  const double* eparam;   /* [ne*PF3_EPARAM_STRIDE] or NULL (defaults) */
  const double* state;    /* [ne*PF3_STATE_STRIDE] or NULL.  When given, frames are NOT
                             recomputed from x/u: the kernels use this state verbatim */
  int32_t state_flags;    /* with state: PF3_STATE_REFRESH_XE recomputes xe and area|length from
                             x with the state's R (update_probe_xe, quad4.pyx:682);
                             PF3_STATE_REFRESH_UE recomputes ue from u (update_probe_ue, :627) */
  double stress[3];       /* Nxx Nyy Nxy for PF3_KG_STRESS */
} pf3_batch;

/* Destination COO arrays of one matrix.  r/c may be NULL ("update_*v_only=1");
 * v may be NULL (indices only).  accumulate!=0 gives the reference's `+=` on values
 * (quad4.pyx:1313); 0 overwrites, which equals `+=` into zero-initialised arrays and
 * saves reading 8 B per entry. */
typedef struct pf3_coo {
  int64_t* r;
  int64_t* c;
  double* v;
  int64_t init_k;
  int32_t accumulate;
} pf3_coo;

/* ---- context / utilities ------------------------------------------------ */
int pf3_version(void);
const char* pf3_error_string(int code);
int pf3_device_count(int* n);
int pf3_create(int device, pf3_context** ctx);
int pf3_destroy(pf3_context* ctx);
int pf3_set_stream(pf3_context* ctx, void* cuda_stream); /* borrow a cudaStream_t (e.g. torch's) */
int pf3_synchronize(pf3_context* ctx);
int pf3_malloc(pf3_context* ctx, size_t bytes, void** dptr);
int pf3_free(pf3_context* ctx, void* dptr);
int pf3_memcpy_h2d(pf3_context* ctx, void* dst, const void* src, size_t bytes);
int pf3_memcpy_d2h(pf3_context* ctx, void* dst, const void* src, size_t bytes);
int pf3_memset(pf3_context* ctx, void* dst, int byte, size_t bytes);
/* number of kernels this context has launched (bench.py's gpu_launches) */
int pf3_launch_count(pf3_context* ctx, int64_t* n);

/* ---- static element data (XData classes, e.g. quad4.pyx:142-180) -------- */
int pf3_num_nodes(int kind);
/* KC0/KG/M_SPARSE_SIZE; 0 where the element has no such matrix */
int pf3_sparse_size(int kind, int matrix);
/* entries actually written for this mtype (lumped mass writes fewer) */
int pf3_written_size(int kind, int matrix, int mtype);

/* ---- element evaluation ------------------------------------------------- */
/* what: bitmask of PF3_KC0|PF3_KG|PF3_KG_STRESS|PF3_M|PF3_FINT (KG and KG_STRESS are
 * exclusive: both write `kg`).  Unused outputs may be NULL.  fint: float64[6*nnodes],
 * accumulated (`fint[c+i] += ...`, quad4.pyx:1339) deterministically (node gather). */
int pf3_eval(pf3_context* ctx, const pf3_batch* batch, int what,
             const pf3_coo* kc0, const pf3_coo* kg, const pf3_coo* m, double* fint);
/* Piston-theory aerodynamic matrices of one Quad4 / Quad4R batch: what = PF3_KA_BETA | PF3_KA_GAMMA | PF3_CA, each
 * filling its own COO arrays (144 entries per element, the KG mask) exactly like
 * quad.update_KA_beta(KA_betar, KA_betac, KA_betav) etc. (tests/test_quad4r_piston_theory.py:116-118).  Geometry
 * comes from x (or batch->state); no property table is read.  PF3_E_UNSUPPORTED for other kinds. */
int pf3_eval_aero(pf3_context* ctx, const pf3_batch* batch, int what, const pf3_coo* ka_beta,
                  const pf3_coo* ka_gamma, const pf3_coo* ca);
/* COO row/col indices only (the unrolled `KC0r[k]=..; KC0c[k]=..` blocks) */
int pf3_fill_indices(pf3_context* ctx, int kind, int matrix, int mtype, int64_t ne,
                     const int64_t* conn, int64_t init_k, int64_t* r, int64_t* c);
/* per-element state as the reference leaves it on the object/probe after
 * update_rotation_matrix + update_probe_xe + update_probe_ue: out[ne*PF3_STATE_STRIDE] */
int pf3_eval_state(pf3_context* ctx, const pf3_batch* batch, double* state_out);
/* probe.finte (local internal force, update_probe_finte e.g. quad4.pyx:1174): out[ne*6*nn] */
int pf3_eval_finte(pf3_context* ctx, const pf3_batch* batch, double* finte_out);

/* Quad4Probe.update_BL(xi, eta) (quad4.pyx:273-395): the 11 strain-interpolation rows (BLexx BLeyy BLgxy
 * BLkxx BLkyy BLkxy BLgyz_grad BLgyz_rot BLgxz_grad BLgxz_rot BLdrilling, 24 doubles each) of n probes from
 * their local coordinates xe[n*12]; out[n*264]. */
int pf3_quad4_update_BL(pf3_context* ctx, int64_t n, const double* xe, double xi, double eta, double* out);

/* ---- assembly: replaces scipy.sparse.coo_matrix((v,(r,c))).tocsr() ------- */
/* (call site tests/test_quad4_static_point_load.py:80).
 * Structured plan: built from connectivity only (indices are a pure function of it).
 * groups: ngroups batches contributing to the SAME matrix; only kind, ne, conn, mtype are
 * read.  coo_offsets[g] = init_k of group g inside the common COO value array.
 * Rows owned by this plan: DOF rows of nodes [node_begin, node_end) — the multi-GPU
 * row-ownership shard; use 0, nnodes for everything. */
int pf3_plan_create(pf3_context* ctx, int matrix, int64_t nnodes, int ngroups,
                    const pf3_batch* groups, const int64_t* coo_offsets,
                    int64_t node_begin, int64_t node_end, pf3_plan** plan);
/* Generic plan from arbitrary COO indices (any element mix / ordering): radix sort of
 * (row,col) keys + unique.  n = matrix dimension. */
int pf3_plan_create_coo(pf3_context* ctx, int64_t n, int64_t nnz_coo, const int64_t* r,
                        const int64_t* c, pf3_plan** plan);
int pf3_plan_destroy(pf3_plan* plan);
int pf3_plan_nnz(const pf3_plan* plan, int64_t* nnz);
int pf3_plan_nrows(const pf3_plan* plan, int64_t* nrows);
/* CSR pattern: indptr[nrows+1], indices[nnz] (int64, column indices global, sorted) */
int pf3_plan_pattern(pf3_context* ctx, const pf3_plan* plan, int64_t* indptr, int64_t* indices);
/* numeric phase: csr_v[nnz] = sum of duplicates of coo_v, deterministic order */
int pf3_plan_assemble(pf3_context* ctx, const pf3_plan* plan, const double* coo_v, double* csr_v);

/* update_fint for every element of `batch` (= group `group` of the structured plan) using the plan's
 * node -> incident element lists: fint[6*nnodes] += ..., fixed summation order, no sort, no atomics; only
 * the plan's owned node rows are updated (e.g. quad4.pyx:1316-1362). */
int pf3_plan_fint(pf3_context* ctx, const pf3_plan* plan, int group, const pf3_batch* batch, double* fint);
int pf3_plan_nblocks(const pf3_plan* plan, int64_t* nblk); /* 6x6 node-pair blocks of a structured plan */

/* Fused evaluate + assemble for ONE Quad4/Quad4R batch and the structured PF3_MAT_KC0 plan built from
 * it: a single kernel writes the COO value arrays (kc0/kg/m, any may be NULL or have v == NULL: that
 * COO array is then not produced) AND the assembled CSR values, without re-reading the COO arrays
 * (DRAM traffic = SURVEY §8(d)'s 15.1 kB/element lower bound).  CSR layouts are those of structured plans
 * of the same batch: csr_kc0[nblk*36], csr_kg[nblk*9], csr_m[nblk*30] (mtype 2: nblk*18).  Replaces the
 * element loop + scipy tocsr pair of tests/test_quad4_static_point_load.py:53-80 in one call.
 * PF3_E_UNSUPPORTED for other kinds / multi-group plans (use pf3_eval + pf3_plan_assemble). */
int pf3_eval_assemble(pf3_context* ctx, const pf3_batch* batch, const pf3_plan* plan, int what,
                      const pf3_coo* kc0, const pf3_coo* kg, const pf3_coo* m, double* csr_kc0,
                      double* csr_kg, double* csr_m);

/* Mixed meshes (e.g. Quad4 skin + BeamC stiffeners in ONE matrix, BASELINE config 5): fused evaluate + assemble of
 * group `group` (a Quad4 / Quad4R batch) of a MULTI-group structured plan.  The CSR outputs use the UNION layouts of
 * KC0 / KG / M over the plan's groups (= the layouts of structured plans of the same groups for those matrices, e.g.
 * 36 entries per block for KG when a BeamC group is present); positions this group does not have are written as zeros.
 * kc0/kg/m address the plan-wide COO value arrays (init_k = the group's coo_offset).  Follow with pf3_eval for the
 * other groups and pf3_plan_assemble_add per matrix. */
int pf3_eval_assemble_group(pf3_context* ctx, const pf3_batch* batch, const pf3_plan* plan, int group, int what,
                            const pf3_coo* kc0, const pf3_coo* kg, const pf3_coo* m, double* csr_kc0, double* csr_kg,
                            double* csr_m);
/* csr_v[nnz] += the contributions of every group of the structured plan EXCEPT skip_group, read from the plan-wide
 * COO value array coo_v; rows of nodes without such contributions are not touched. */
int pf3_plan_assemble_add(pf3_context* ctx, const pf3_plan* plan, const double* coo_v, double* csr_v, int skip_group);

/* The same step with HOST buffers on a fixed mesh (the optimisation / nonlinear loop of a reference script that
 * re-runs its element loop and scipy assembly every iteration with new x / u): batch->x and batch->u are HOST
 * pointers, copied to device staging owned by the context; every other batch pointer (conn, props, prop_id, evec,
 * eparam) stays device-resident as given to pf3_plan_create; csr_*_host are HOST pointers (pinned memory
 * recommended) that receive the assembled values; kc0/kg/m are optional DEVICE COO destinations.  Returns after the
 * copies have completed. */
int pf3_eval_assemble_host(pf3_context* ctx, const pf3_batch* batch, const pf3_plan* plan, int what,
                           const pf3_coo* kc0, const pf3_coo* kg, const pf3_coo* m, double* csr_kc0_host,
                           double* csr_kg_host, double* csr_m_host);

/* y = A x for a CSR matrix with int64 indptr/indices (downstream cg / eigsh operators) */
int pf3_spmv_csr(pf3_context* ctx, int64_t nrows, const int64_t* indptr, const int64_t* indices,
                 const double* vals, const double* x, double* y);

/* y = P A P x, P = diag(free_dof != 0): the boundary-condition partition KC0[bu,:][:,bu] every reference script
 * forms right after assembly (tests/test_quad4_static_point_load.py:84-99), applied on the fly (square A). */
int pf3_spmv_csr_masked(pf3_context* ctx, int64_t nrows, const int64_t* indptr, const int64_t* indices,
                        const double* vals, const unsigned char* free_dof, const double* x, double* y);
/* diag[r] = A[r, r + row0]: Jacobi scaling used by the reference's preconditioned cg calls
 * (tests/test_quad4r_linear_buckling_plate.py:135-146) */
int pf3_csr_diagonal(pf3_context* ctx, int64_t nrows, const int64_t* indptr, const int64_t* indices,
                     const double* vals, int64_t row0, double* diag);

/* y = A x (free_dof == NULL) or y = P A P x (P = diag(free_dof != 0), free_dof[6*nnodes] over GLOBAL dofs) for the
 * CSR values of a STRUCTURED plan, using the plan's node-block structure instead of per-entry indices (8.2 B per
 * nonzero instead of 16): the operator the reference scripts hand to cg / eigsh after `KC0[bu, :][:, bu]`
 * (tests/test_quad4_static_point_load.py:84-104, tests/test_quad4r_linear_buckling_plate.py:135-180).
 * x[6*nnodes] (global), y[6*nown] (the plan's own rows).  PF3_E_UNSUPPORTED for generic (COO) plans. */
int pf3_plan_spmv(pf3_context* ctx, const pf3_plan* plan, const double* vals, const unsigned char* free_dof,
                  const double* x, double* y);
/* diag[6*nown]: the diagonal of the plan's row block (0 where the pattern has no diagonal entry) */
int pf3_plan_diagonal(pf3_context* ctx, const pf3_plan* plan, const double* vals, double* diag);

/* ---- device solves on the assembled matrix ------------------------------------------ */
/* Jacobi-preconditioned conjugate gradient for (P A P) x = P b, everything on the device:
 *   A = sum_i coefs[i] * (CSR values vals[i] of STRUCTURED plan plans[i])   -- e.g. K, or K - sigma M with the two
 *       matrices kept in their own layouts; every plan must own all rows (single device; a row-sharded solve exchanges
 *       the search direction between ranks, pyfe3d_b200/solve.py);
 *   P = diag(free_dof != 0) (NULL: no constraints): the partition K[bu,:][:,bu] of
 *       tests/test_quad4_static_point_load.py:84-99, applied inside the SpMV;
 *   b, x: float64[6*nnodes] over ALL dofs; x is zero on constrained dofs on return; use_x0 != 0 starts from x.
 * One iteration = the plans' block SpMV (8.2 B per nonzero) + three fused vector kernels (dot; x/r update with both
 * reductions; direction update), reductions in a fixed order (bit-reproducible), convergence decided on the device and
 * looked at by the host every (flags >> 8) iterations (0: 16).  This is the reference scripts' diagonally scaled solve
 *   D = diag(Kuu)^-1/2;  cg(D Kuu D, D fu, atol=...);  uu = D uu_scaled   (tests/test_quad4r_linear_buckling_plate.py:135-146)
 * written as preconditioned CG on the unscaled system (identical iterates).  Stops when the residual norm
 * <= max(rtol * |b|, atol); PF3_CG_SCALED_NORM measures both in the scaled norm (r.D^2.r) that scipy's cg sees on the
 * scaled system.  maxiter <= 0: 10 * ndof. */
#define PF3_CG_SCALED_NORM 1
#define PF3_CG_GRAPH 2 /* replay each batch of iterations as ONE CUDA graph (launch-bound small systems); needs a
                          capturable stream -- on the legacy default stream the call falls back to plain launches */
typedef struct pf3_cg_info {
  int32_t iterations;
  int32_t status;   /* 0 converged, 1 maxiter reached, 2 breakdown (p.Ap <= 0: matrix not positive definite on bu) */
  double residual;  /* the norm the criterion used */
  double bnorm;
} pf3_cg_info;
int pf3_plan_cg(pf3_context* ctx, int nops, const pf3_plan* const* plans, const double* const* vals,
                const double* coefs, const unsigned char* free_dof, const double* b, double* x, int use_x0,
                double rtol, double atol, int maxiter, int flags, pf3_cg_info* info);
/* y = S P A P S x with S = diag(scale[6*nnodes]): the scaled operators KC0uu_scaled / KGuu_scaled the reference hands to
 * eigsh (tests/test_quad4r_linear_buckling_plate.py:172-180).  Plan owning all rows only. */
int pf3_plan_spmv_scaled(pf3_context* ctx, const pf3_plan* plan, const double* vals, const unsigned char* free_dof,
                         const double* scale, const double* x, double* y);

/* The vector kernels of ONE iteration of the Jacobi-CG on a rank's OWN rows, for the row-sharded multi-GPU solve
 * (pyfe3d_b200/solve.py plan_cg_solve with a process group: block SpMV on the own rows, point-to-point halo exchange of
 * the search direction, these three fused kernels with two small all_reduces between them).  n: own rows.
 *   sc (device, 4 doubles): [0] p.Ap   [1] new r.M^-1.r   [2] r.r   [3] current r.M^-1.r
 *   _dot   : sc[0] = sum p*ap over the own rows                          -> caller all-reduces sc[0]
 *   _update: a = sc[3]/sc[0]; x += a p; r -= a ap; sc[1] = sum r*minv*r; sc[2] = sum r*r   -> caller all-reduces sc[1..2]
 *   _dir   : p = minv r + (sc[1]/sc[3]) p; then sc[3] = sc[1]
 * Reductions are deterministic (fixed block order).  work: pf3_cg_shard_work_bytes() bytes on the device, zeroed once. */
size_t pf3_cg_shard_work_bytes(void);
int pf3_cg_shard_dot(pf3_context* ctx, int64_t n, const double* p, const double* ap, double* sc, void* work);
int pf3_cg_shard_update(pf3_context* ctx, int64_t n, const double* p, const double* ap, const double* minv, double* x,
                        double* r, double* sc, void* work);
int pf3_cg_shard_dir(pf3_context* ctx, int64_t n, const double* r, const double* minv, double* p, double* sc, void* work);

/* K[bu, :][:, bu] as an explicit CSR matrix (tests/test_quad4_static_point_load.py:84-99) from a device CSR matrix whose
 * rows are the global rows [row0, row0 + nrows) (row0 = 6*node_begin of a row-sharded plan, else 0) and whose column
 * indices are global in [0, ncols): rows / columns with free_dof[.] != 0 are kept and renumbered by their rank among
 * the free dofs (columns: global rank; rows: rank within this row block).
 * free_dof == NULL keeps every row and column.  flags & PF3_COMPACT_UPPER additionally keeps only entries with
 * col >= row (scipy.sparse.triu): KC0, KG and M are symmetric, so a host consumer that accepts one triangle needs 5/9
 * of the values moved over PCIe.
 *  symbolic: colmap[ncols + 1] <- exclusive scan of the free flags (scratch the fill step needs again),
 *            out_indptr[<= nrows + 1] <- row pointers of the compacted block; *nkeep rows, *nnz entries (host).
 *  fill:     out_indices[nnz] (may be NULL: values-only refresh for a fixed pattern), out_vals[nnz] (may be NULL). */
#define PF3_COMPACT_UPPER 1
int pf3_csr_compact_symbolic(pf3_context* ctx, int64_t nrows, int64_t ncols, const int64_t* indptr,
                             const int64_t* indices, const unsigned char* free_dof, int flags, int64_t row0,
                             int64_t* colmap, int64_t* out_indptr, int64_t* nkeep, int64_t* nnz);
int pf3_csr_compact_fill(pf3_context* ctx, int64_t nrows, int64_t ncols, const int64_t* indptr,
                         const int64_t* indices, const double* vals, const unsigned char* free_dof, int flags,
                         int64_t row0, const int64_t* colmap, const int64_t* out_indptr, int64_t* out_indices,
                         double* out_vals);

/* ---- property tables on the device ---------------------------------------- */
/* props_out[nrows * PF3_SHELLPROP_STRIDE]: the ShellProp scalars of nrows laminates of nplies plies each, what
 *   laminated_plate(stack, plyts=..., laminaprops=..., rhos=..., offset=..., calc_scf=...)  (pyfe3d/shellprop_utils.py:96)
 * returns in A11..D66, E44..E55, scf_k13/k23, h, intrho, intrhoz, intrhoz2 (Lamina.rebuild shellprop.pyx:278,
 * calc_constitutive_matrix :568, calc_scf :485), one laminate per thread.  Row r reads thetadeg[r*theta_stride + p]
 * (degrees), plyt[r*plyt_stride + p], lamina[r*lamina_stride + 8*p + {e1,e2,nu12,g12,g13,g23,rho,pad}] and
 * offset[r*offset_stride] (offset may be NULL = 0); a stride of 0 shares one stack / thickness set / material set /
 * offset between all rows.  calc_scf == 0 leaves the factors at 5/6. */
int pf3_laminate_props(pf3_context* ctx, int64_t nrows, int nplies, const double* thetadeg, int64_t theta_stride,
                       const double* plyt, int64_t plyt_stride, const double* lamina, int64_t lamina_stride,
                       const double* offset, int64_t offset_stride, int calc_scf, double* props_out);

/* Lamination-parameter form of the same table and its gradient rows.
 * props_out[nrows * PF3_SHELLPROP_STRIDE] (may be NULL): what
 *   shellprop_from_LaminationParameters(thickness, mat, lp)                         (pyfe3d/shellprop.pyx:767-815)
 * stores in A11..D66, E44..E55 and h; scf_k13 = scf_k23 = 5/6 as that function leaves them.  Row r reads
 * thickness[r*thickness_stride], the material invariants u1..u7 (MatLamina.rebuild, shellprop.pyx:189-195) at
 * invariants[r*invariants_stride + 0..6] and xiA1..4, xiB1..4, xiD1..4, xiE1..2 at lp[r*lp_stride + 0..13]; a stride
 * of 0 shares the value between all rows.  rho (may be NULL) is an extension for homogeneous density: intrho = rho h,
 * intrhoz = 0, intrhoz2 = rho h^3/12; with NULL the three stay 0 like the reference.
 * grad_out[nrows * popcount(var_mask) * PF3_SHELLPROP_STRIDE] (may be NULL): for every laminate, one PROPERTY row
 * per selected variable v (bit v of var_mask; 0 = h, 1..4 = xiA1..4, 5..8 = xiB1..4, 9..12 = xiD1..4, 13..14 =
 * xiE1..2, in that order) holding d(A, B, D, E, mass integrals)/dv -- the numbers GradABDE.calc_LP_grad
 * (shellprop.pyx:933-1014) puts in gradAij/gradBij/gradDij/gradEij -- with scf and h repeated, so that update_KC0 /
 * update_M evaluated with that row are dKC0/dv and dM/dv (the element matrices are linear in these scalars; not
 * valid for Quad4R's hourglass control, which divides by the inverse of ABD, quad4r.pyx:713-746).
 * grad_complete == 0 keeps the reference's range(5) loops (shellprop.pyx:968,981,994): d(A66,B66,D66)/d(xi) = 0;
 * grad_complete != 0 stores the mathematically complete -fac*u3 for xi3. */
#define PF3_LP_NVARS 15
int pf3_lamination_parameter_props(pf3_context* ctx, int64_t nrows, const double* thickness, int64_t thickness_stride,
                                   const double* invariants, int64_t invariants_stride, const double* lp,
                                   int64_t lp_stride, const double* rho, int64_t rho_stride, int var_mask,
                                   int grad_complete, double* props_out, double* grad_out);

/* ---- host-pointer convenience (numpy callers): copies in, runs, copies out -- */
/* Every pointer in host_batch / the pf3_coo structs / fint is a HOST pointer.  Calls whose arrays fit the
 * context's 1 MiB staging buffer (the per-element methods of pyfe3d_b200/elements.py: X.update_KC0(KC0r, KC0c,
 * KC0v, prop) etc., quad4.pyx:1204) cost one packed copy in each direction and one stream synchronisation. */
int pf3_eval_host(pf3_context* ctx, const pf3_batch* host_batch, int what,
                  const pf3_coo* kc0, const pf3_coo* kg, const pf3_coo* m, double* fint);
/* Host-pointer forms of pf3_eval_state / pf3_eval_finte / pf3_eval_aero / pf3_quad4_update_BL for small batches
 * (update_rotation_matrix / update_probe_xe / update_probe_ue quad4.pyx:491,682,627; update_probe_finte :1174;
 * update_KA_beta :9491; Quad4Probe.update_BL :273).  PF3_E_CAPACITY when the arrays exceed the staging buffer:
 * use the device-pointer entry points for large batches. */
int pf3_eval_state_host(pf3_context* ctx, const pf3_batch* host_batch, double* state_out);
int pf3_eval_finte_host(pf3_context* ctx, const pf3_batch* host_batch, double* finte_out);
int pf3_eval_aero_host(pf3_context* ctx, const pf3_batch* host_batch, int what, const pf3_coo* ka_beta,
                       const pf3_coo* ka_gamma, const pf3_coo* ca);
int pf3_quad4_update_BL_host(pf3_context* ctx, int64_t n, const double* xe, double xi, double eta, double* out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* PYFE3D_B200_H */
