// Shared device helpers for the element-evaluation kernels (sm_100a, FP64).
//
// Execution shape: ONE ELEMENT PER THREAD for the arithmetic, so the
// frame / Jacobian / shape-gradient set-up is never replicated across lanes, and a
// per-warp shared-memory stage so that every global store instruction writes one
// contiguous >=256 B run of the caller's COO value array (an element's block is
// contiguous: 576 / 144 / 480 doubles).
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/pyfe3d_b200.h"

namespace pf3 {

// Function attributes (cudaFuncSetAttribute) are per DEVICE: a call site's first() is true once per device, so a
// process that drives several GPUs (tests, a multi-context host) raises the shared-memory limit on each of them.
struct PerDeviceOnce {
  std::atomic<unsigned long long> seen{0ull};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long bit = 1ull << (d & 63);
    return (seen.fetch_or(bit) & bit) == 0ull;
  }
};

constexpr int kWarpsPerCta = 4;
constexpr int kThreads = 32 * kWarpsPerCta;
constexpr int kChunk = 72;     // doubles one thread stages between flushes (3 rows x 24)
constexpr int kStageLd = 73;   // odd leading dimension: conflict-free row writes
constexpr size_t kStageBytes = size_t(kWarpsPerCta) * 32 * kStageLd * sizeof(double);

struct EvalArgs {
  int64_t ne;
  const int64_t* __restrict__ conn;
  const double* __restrict__ x;
  const double* __restrict__ u;
  const double* __restrict__ props;
  const int32_t* __restrict__ prop_id;
  const double* __restrict__ evec;
  int evec_stride;
  const double* __restrict__ eparam;
  const double* __restrict__ state;
  int state_flags;
  int what;
  int mtype;
  double Nxx, Nyy, Nxy;
  double* kc0v;
  double* kgv;
  double* mv;
  double* fe;        // [ne][6*nn] element force in global axes (PF3_FINT scratch)
  double* finte;     // [ne][6*nn] local internal force (probe.finte), optional
  double* state_out; // [ne][PF3_STATE_STRIDE], optional
  int64_t kc0_k0, kg_k0, m_k0;
  int acc_kc0, acc_kg, acc_m;
};

// Property-table row of element e.  Under compute-sanitizer's memcheck the SASS this compiles to (a predicated-off LDG +
// CS2R + predicated IMAD.WIDE) is executed with garbage row offsets -- false out-of-bounds reports at the property loads
// that follow, although the hardware runs it correctly (goldens match to 1e-12, and a build with a printf next to it is
// clean under the tool).  -DPF3_MEMCHECK_BUILD (scripts/build_memcheck_lib.sh, used by scripts/gpu_sanitizer.sh) selects
// a form with an unconditional index load and selects, which keeps memcheck usable on the whole library; it costs about
// 1 % on update_fint / config 4 (one more dependent load in front of the property loads), so it is not the default.
__device__ __forceinline__ int64_t prop_index(const EvalArgs& A, int64_t e) {
#ifdef PF3_MEMCHECK_BUILD
  const bool has = A.prop_id != nullptr;
  const int32_t* pp = has ? A.prop_id + e : reinterpret_cast<const int32_t*>(A.conn);
  const int v = *pp;
  return has ? int64_t(v) : int64_t(0);
#else
  return A.prop_id ? int64_t(A.prop_id[e]) : int64_t(0);
#endif
}

// destinations of the piston-theory matrices (quad_aero_kernel): 0 KA_beta, 1 KA_gamma, 2 CA
struct AeroOut {
  double* v[3];
  int64_t k0[3];
  int acc[3];
};

// Per-node work record of the fused kernel, built once per plan (one 64-B load replaces the
// inc_ptr -> inc_pair0 -> slot pointer chase).  One record per (node, round of 4 incidences).
struct alignas(16) NodeRec {
  int64_t b0;          // first 6x6 block of the node's CSR row block
  int32_t inc[4];      // e*16 + a*4 of the incident (element, local node) pairs of this round, -1: none
  uint16_t gmap[16];   // per column-block slot: nibble k -> local node b of incidence k feeding it, 0xF: none
  uint8_t v, nb;       // incidences in this round, column blocks of the node row
  uint8_t pad[6];
};
static_assert(sizeof(NodeRec) == 64, "NodeRec must be 64 bytes");

// The same for the triangle kernel (tria_fused.cu): one record per (node, round of 10 incidences).
struct alignas(16) TriaRec {
  int64_t b0;          // first 6x6 block of the node's CSR row block
  int32_t inc[10];     // e*9 + a*3 of the incident (element, local node) pairs of this round, -1: none
  uint32_t gmap[16];   // per column-block slot: 2 bits per incidence k -> local node b feeding it, 3: none
  uint8_t v, nb;       // incidences in this round, column blocks of the node row
  uint8_t pad[14];
};
static_assert(sizeof(TriaRec) == 128, "TriaRec must be 128 bytes");

// Where the fused kernel's own (rows x masked columns) block of one matrix lands inside the UNION row layout of a
// multi-group plan (e.g. Quad4 skin + BeamC stiffeners: the quad's 3x3 KG entries inside 6x6 union blocks).
struct UnionMap {
  int8_t row[6];      // union row d -> own row index (rank among the own non-empty rows), -1: the group has no such row
  int8_t col[6][6];   // union (row d, column rank ju) -> own masked-column rank in that row, -1: structural zero here
  int cnt[6], rowoff[6], mc;   // the union layout: columns per block in row d, block-row offset, entries per block
  unsigned colbits[6];         // col[d][*] packed 4 bits per union column rank (0xF = structural zero)
  int active;         // 0: the union layout IS the own layout (fast path)
};

// fused evaluate + assemble (quad_fused.cu): element inputs + the plan's block structure
struct FusedArgs {
  EvalArgs A;
  const int64_t* __restrict__ brow_ptr;   // [nown+1] first block of each owned node row
  const int64_t* __restrict__ inc_ptr;    // [nown+1] incident (element, local node) pairs per node
  const int64_t* __restrict__ inc_pair0;  // [ninc]   e*16 + a*4
  const int32_t* __restrict__ slot;       // [ne*16]  column-block slot of node pair (e, a, b) in a's row
  const NodeRec* __restrict__ noderec;    // [nown*rmax]   (quads)
  const TriaRec* __restrict__ triarec;    // [nown*rmax]   (triangles)
  int rmax;                               // rounds of 4 incidences per node (1 when every valence <= 4)
  int64_t nown;
  double* csr_kc0;
  double* csr_kg;
  double* csr_m;
  UnionMap um[3];                         // KC0, KG, M: only read when um[i].active
  int zero_empty;                         // multi-group plans: zero the rows of nodes this group does not touch
  int64_t pair_first;                     // quads: this launch covers node pairs [pair_first, pair_first + pair_count)
  int64_t pair_count;                     //        (pair_count == 0: all pairs) -- pf3_eval_assemble_host's pipeline
  const int2* __restrict__ pftab;         // L2 prefetch table [pf_nchunks][kPfRuns] (see kPfChunk), nullptr: no prefetch
  int64_t pf_nchunks;
};

// L2 prefetch table of the fused shell kernels (quad_fused.cu, tria_fused.cu): the node pairs are cut into chunks of
// kPfChunk; entry c holds kPfRuns contiguous runs of element records that are used FIRST by a node of chunk c, one per
// quarter of the element range (x = first element, y = count; y = 0: none, or too scattered to be worth reading) --
// structured meshes first-use one run per chunk, meshes built from several structured parts (config 4: the second
// triangle of every cell is numbered ne/2 later) one run per part.  The CTA of the first node of chunk c issues bulk L2
// prefetches for the node records and those element records of chunk c + kPfAhead.
constexpr int kPfRuns = 4;
// Measured and rejected: a PRODUCER MODE without the record launch -- in block order every chunk preceded by the 40
// one-warp producer CTAs (K1 for 32 elements each) of the chunk 8 chunks later, per-chunk completion flags, consumers
// find the records in L2.  Parity-green, 10.82 ms against 10.48 ms for K1 + K2 on the same box: the producers' gathers
// of conn / x / u are DRAM reads inside the write stream again, and their latency-bound warps hold K2's CTA slots.
// Thread-per-element kernels: the connectivity of the 32 elements a warp will gather through is streamed from DRAM and
// heads a chain of dependent round trips (connectivity -> coordinates / displacements).  One lane per warp asks L2 for the
// connectivity rows of the warp PF3_CONN_AHEAD elements further on, so that the first trip is an L2 hit.
#ifndef PF3_CONN_AHEAD
#define PF3_CONN_AHEAD 65536
#endif
__device__ __forceinline__ void conn_prefetch(const int64_t* conn, int nn, int64_t e0, int64_t ne, int lane) {
#if PF3_CONN_AHEAD > 0
  const int64_t e1 = e0 + PF3_CONN_AHEAD;
  if (lane == 0 && e1 + 32 <= ne) {
    const char* q = reinterpret_cast<const char*>(conn + e1 * nn);
    const unsigned bytes = unsigned(32 * nn * 8);
    if ((reinterpret_cast<uintptr_t>(q) & 15) == 0)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q), "r"(bytes) : "memory");
  }
#endif
}
constexpr int kPfChunk = 512;
constexpr int kPfAhead = 8;   // (4, 8 and 16 chunks ahead measure the same, 32 slightly worse)
// Halving the node records to 32 bytes (per-incidence slot lists, expanded into gmap by the kernel) was measured and
// rejected: 10.88 ms against 10.75 for these 64-byte records on the same box -- with the bulk prefetch in place their
// DRAM reads are already batched, and the expansion sits on the critical path of every CTA.

#ifndef PF3_L2_HINTS
#define PF3_L2_HINTS 1   // stores of the fused kernels carry an L2 evict-first policy
#endif
// L2 evict-first policy for everything the fused kernels write: their output (60 GB per step for 4 M Quad4) streams
// through the L2 and is never read again, while the element records are re-read by up to three later CTAs and the prefetched records must survive until
// their CTAs run.  (Measured with scripts/micro/k2_stream_reads.cu: a DRAM read among the writes costs ~12x its bytes.)
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void stg_stream(double2* p, double2 v, uint64_t pol) {
#if PF3_L2_HINTS
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
#else
  *p = v;
#endif
}
__device__ __forceinline__ void stg_stream(double* p, double v, uint64_t pol) {
#if PF3_L2_HINTS
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
#else
  *p = v;
#endif
}


struct Mat3 {
  double a[3][3];
};

__device__ __forceinline__ double dot3(const double* a, const double* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double normalize3(double* v) {
  // one division and three multiplications (the reference divides each component, quad4.pyx:548-552: the components
  // differ from that in the last bit at most; a double division is a ~15-instruction sequence and every frame has three
  // of these)
  double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const double inv = 1. / n;
  v[0] *= inv;
  v[1] *= inv;
  v[2] *= inv;
  return n;
}

// out = R * L * R^T for a local 3x3 block L = [[l00,l01,0],[l10,l11,0],[0,0,l22]]
// (translation-translation and rotation-rotation shell blocks).
__device__ __forceinline__ void rot_block_diag5(const Mat3& R, double l00, double l01, double l10,
                                                double l11, double l22, double (*o)[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double t0 = R.a[i][0] * l00 + R.a[i][1] * l10;
    double t1 = R.a[i][0] * l01 + R.a[i][1] * l11;
    double t2 = R.a[i][2] * l22;
#pragma unroll
    for (int j = 0; j < 3; ++j) o[i][j] = t0 * R.a[j][0] + t1 * R.a[j][1] + t2 * R.a[j][2];
  }
}
// out = R * L * R^T for L with only l22 == 0 (translation-rotation couplings).
__device__ __forceinline__ void rot_block_8(const Mat3& R, double l00, double l01, double l02,
                                            double l10, double l11, double l12, double l20,
                                            double l21, double (*o)[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double t0 = R.a[i][0] * l00 + R.a[i][1] * l10 + R.a[i][2] * l20;
    double t1 = R.a[i][0] * l01 + R.a[i][1] * l11 + R.a[i][2] * l21;
    double t2 = R.a[i][0] * l02 + R.a[i][1] * l12;
#pragma unroll
    for (int j = 0; j < 3; ++j) o[i][j] = t0 * R.a[j][0] + t1 * R.a[j][1] + t2 * R.a[j][2];
  }
}
// generic dense 3x3 (line elements)
__device__ __forceinline__ void rot_block_full(const Mat3& R, const double (*l)[3], double (*o)[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double t[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) t[q] = R.a[i][0] * l[0][q] + R.a[i][1] * l[1][q] + R.a[i][2] * l[2][q];
#pragma unroll
    for (int j = 0; j < 3; ++j) o[i][j] = t[0] * R.a[j][0] + t[1] * R.a[j][1] + t[2] * R.a[j][2];
  }
}

// Every lane has written `len` doubles of ITS element at stage[lane*kStageLd + k].
// Write them to out[(e0+lane)*estride + off + k] with warp-contiguous stores.
template <int LEN>
__device__ __forceinline__ void flush_chunk(const double* stage, double* __restrict__ out, int64_t e0,
                                            int nvalid, int64_t estride, int off, bool accumulate,
                                            int lane) {
  __syncwarp();
  double* base = out + e0 * estride + off;
  // lane walks positions idx = lane, lane + 32, ... of the nvalid x LEN tile; (el, k) follow by carry instead of a
  // division per store
  int el = lane / LEN, k = lane - el * LEN;
  constexpr int kStepEl = 32 / LEN, kStepK = 32 - kStepEl * LEN;
  if (accumulate) {
    while (el < nvalid) {
      double* p = base + int64_t(el) * estride + k;
      *p += stage[el * kStageLd + k];
      el += kStepEl;
      k += kStepK;
      if (k >= LEN) {
        k -= LEN;
        ++el;
      }
    }
  } else {
    if constexpr (LEN % 2 == 0) {
      // even runs on 16-byte aligned bases: two doubles per store
      if (((reinterpret_cast<uintptr_t>(base) & 15) == 0) && (estride & 1) == 0) {
        constexpr int H = LEN / 2, kStepEl2 = 32 / H, kStepK2 = 32 - kStepEl2 * H;
        int el2 = lane / H, k2 = lane - el2 * H;
        while (el2 < nvalid) {
          const double* s = stage + el2 * kStageLd + 2 * k2;
          *reinterpret_cast<double2*>(base + int64_t(el2) * estride + 2 * k2) = make_double2(s[0], s[1]);
          el2 += kStepEl2;
          k2 += kStepK2;
          if (k2 >= H) {
            k2 -= H;
            ++el2;
          }
        }
        __syncwarp();
        return;
      }
    }
    while (el < nvalid) {
      base[int64_t(el) * estride + k] = stage[el * kStageLd + k];
      el += kStepEl;
      k += kStepK;
      if (k >= LEN) {
        k -= LEN;
        ++el;
      }
    }
  }
  __syncwarp();
}

}  // namespace pf3
