// Fused, node-centric Quad4 / Quad4R evaluation + CSR assembly.
//
// One pass produces BOTH outputs the hot path owes:
//   * the reference's COO value arrays KC0v / KGv / Mv (update_KC0 quad4.pyx:1204, update_KG :1365,
//     update_KG_given_stress :2259, update_M :3083; Quad4R: quad4r.pyx:1145, :4681, :5574, :6398), and
//   * the CSR values scipy's coo_matrix(...).tocsr() would give (tests/test_quad4_static_point_load.py:80),
// without ever re-reading the COO arrays: DRAM traffic is the 15.1 kB/element lower bound of SURVEY §8(d)
// plus a 0.3-0.45 kB/element record (below).
//
// Two kernels:
//  K1 quad_record_kernel  one THREAD per element: frame (update_rotation_matrix quad4.pyx:491), local
//     coordinates (update_probe_xe :682), area, inverse Jacobian determinants, material-axis rotation of
//     A/B/D (:847-899) and, for KG, the membrane force resultants per Gauss point from the local
//     displacements (update_probe_ue :627, :1965-1967).  Everything with a sqrt/division or a gather
//     lives here, once per element; the result is a 36- or 56-double record.
//  K2 quad_fused_kernel   node-centric.  The unit of work is a NODE (= 6 CSR rows).  Its row block is the
//     sum, over the incident elements, of the element's 6x24 row slab for that node, and every (element,
//     local node) slab belongs to exactly one node.  A half-warp takes one node: 4 incident elements x 4
//     node-pair blocks = 16 lanes, each lane evaluating ONE 6x6 block (a, b) of ONE element with the same
//     instruction stream (no divergence).  Blocks are staged in shared memory, streamed out as contiguous
//     COO slabs (1152 B each) and summed per CSR slot in a fixed order (deterministic, no atomics).
#include <algorithm>

#include "shell.cuh"

namespace pf3 {

namespace {

#ifndef PF3_COO_TMA
#define PF3_COO_TMA 1   // 1: COO slabs leave as TMA bulk copies; 0: 16-B stores by the incidence's 4 lanes
#endif
constexpr double kGpF = 0.5773502691896257645092;
#ifndef PF3_FUSED_WARPS
#define PF3_FUSED_WARPS 1   // warps per CTA and CTAs per SM of the fused kernel: 12 warps/SM at 168 registers;
#define PF3_FUSED_CTAS 12   // one-warp CTAs measured best (10.82 ms vs 10.98 for 2x6 and 11.15 for 4x3)
#endif
constexpr int kFusedWarps = PF3_FUSED_WARPS;

#ifndef PF3_L2_PREFETCH
#define PF3_L2_PREFETCH 1   // bulk L2 prefetch of the node / element records kPfAhead chunks of node pairs ahead
#endif
constexpr int kMaxSlots = 16;               // column blocks per node row supported by the fused path (NodeRec::gmap)
// Element record (doubles): 0..5 the element x and y axes (R columns 0 and 1, row-major 3 x 2; z = x X y is recomputed
// by the consumer) | 6..13 the eight local edge differences | 14..17 1/detJ at the 2x2 Gauss points | 18 1/detJ at the
// centre | 19 area (Quad4: NEGATIVE when the element is "thick", h / sqrt(area) >= 1, quad4.pyx:1032) | [Quad4R: the
// drilling coefficient 1e-6 K6ROT A66 and the five hourglass coefficients w0 gamma^2 E_d, quad4r.pyx:3088-3116 -- nine
// divisions that every one of the 16 lanes working on an element would otherwise repeat] | [KG from u: 12 membrane force
// resultants Nxx, Nyy, Nxy at the 4 Gauss points] | [material axes: the rotated A, B, D, 18 doubles].  Quad4: 20 / 32 /
// 38 / 50 doubles (the three-matrix north-star call reads 256-byte records that never straddle a third 128-byte line);
// Quad4R: 6 more.
constexpr int kRecBase = 20;
constexpr int kRecHg = 6;
constexpr int kRecN = 12;
constexpr int kRecABD = 18;
constexpr int kRecMax = kRecBase + kRecHg + kRecN + kRecABD;
__host__ __device__ constexpr int rec_stride(bool quad4r, bool kg_u, bool rot) {
  return kRecBase + (quad4r ? kRecHg : 0) + (kg_u ? kRecN : 0) + (rot ? kRecABD : 0);
}
// shared-memory stride of a staged record: even (16-byte aligned) and = 2 mod 4, so that the 16-byte loads of the 8
// incidences of a warp fall into disjoint banks
__host__ __device__ constexpr int rec_ld(int stride) { return stride + ((stride & 3) == 2 ? 0 : 2); }

// ------------------------------------------------------------------------------------------ K1
#ifndef PF3_K1_CTAS
#define PF3_K1_CTAS 3
#endif
// The records of the 32 consecutive elements [e0, e0 + nvalid) by one warp: staged in shared memory (odd leading
// dimension ld = stride + 1, 32 * ld doubles at `stage`) and written out as one contiguous run of nvalid x stride doubles.
template <int KIND>
__device__ __forceinline__ void record_warp(const EvalArgs& A, double* __restrict__ rec, int stride, int64_t e0, int nvalid,
                                            double* stage, int lane) {
  const int64_t e = e0 + min(lane, nvalid - 1);
  const int ld = stride + 1;
  const bool kg_u = (A.what & PF3_KG) != 0;
  double ue[24];
  ShellGeom<4> g;
  shell_geom<4, true>(A, e, g, kg_u ? ue : nullptr);
  double* r = stage + lane * ld;
  const bool rot = A.evec != nullptr;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    r[2 * i] = g.R.a[i][0];
    r[2 * i + 1] = g.R.a[i][1];
  }
  const double d[8] = {g.X[1] - g.X[0], g.X[2] - g.X[3], g.X[3] - g.X[0], g.X[2] - g.X[1],
                       g.Y[1] - g.Y[0], g.Y[2] - g.Y[3], g.Y[3] - g.Y[0], g.Y[2] - g.Y[1]};
#pragma unroll
  for (int i = 0; i < 8; ++i) r[6 + i] = d[i];
  double J11e[2], J12e[2], J21x[2], J22x[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double t = i ? kGpF : -kGpF;
    J11e[i] = 0.25 * ((1. - t) * d[0] + (1. + t) * d[1]);
    J12e[i] = 0.25 * ((1. - t) * d[4] + (1. + t) * d[5]);
    J21x[i] = 0.25 * ((1. - t) * d[2] + (1. + t) * d[3]);
    J22x[i] = 0.25 * ((1. - t) * d[6] + (1. + t) * d[7]);
  }
  double idJ[4];
#pragma unroll
  for (int gp = 0; gp < 4; ++gp) {
    idJ[gp] = 1. / (J11e[gp & 1] * J22x[gp >> 1] - J12e[gp & 1] * J21x[gp >> 1]);
    r[14 + gp] = idJ[gp];
  }
  const double J11c = 0.25 * (d[0] + d[1]), J12c = 0.25 * (d[4] + d[5]);
  const double J21c = 0.25 * (d[2] + d[3]), J22c = 0.25 * (d[6] + d[7]);
  r[18] = 1. / (J11c * J22c - J12c * J21c);
  r[19] = g.area;
  constexpr int kHg = (KIND == PF3_QUAD4R) ? kRecHg : 0;
  if (KIND == PF3_QUAD4 && A.props != nullptr) {
    const double hh = A.props[prop_index(A, e) * PF3_SHELLPROP_STRIDE + 23];
    if (hh / sqrt(g.area) >= 1.) r[19] = -g.area;   // thick: transverse shear integrated at 2x2 (quad4.pyx:1032,1127)
  }
  if (kg_u || rot || KIND == PF3_QUAD4R) {
    ShellCoef c;
    shell_coef<4>(A, e, g, c);
    if (KIND == PF3_QUAD4R) {
      double K6ROT = 100., hgf[5] = {1., 1., 1., 1., 1.};
      if (A.eparam != nullptr) {
        const double* ep = A.eparam + e * PF3_EPARAM_STRIDE;
        K6ROT = ep[0];
#pragma unroll
        for (int q = 0; q < 5; ++q) hgf[q] = ep[2 + q];
      }
      const double* cA = c.A;
      const double hh = c.h;
      const double den = -cA[0] * cA[3] * cA[5] + cA[0] * cA[4] * cA[4] + cA[1] * cA[1] * cA[5] -
                         2 * cA[1] * cA[2] * cA[4] + cA[2] * cA[2] * cA[3];
      const double a11 = (-cA[3] * cA[5] + cA[4] * cA[4]) / den, a22 = (-cA[0] * cA[5] + cA[2] * cA[2]) / den;
      const double E1eq = 1. / (hh * a11), E2eq = 1. / (hh * a22);
      const double dd = 1.0 + 1.0 / g.area;
      const double Eu = hgf[0] * 0.1 * E1eq * hh / dd, Ev = hgf[1] * 0.1 * E2eq * hh / dd;
      const double Erx = hgf[3] * 0.1 * E2eq * hh * hh * hh / dd, Ery = hgf[4] * 0.1 * E1eq * hh * hh * hh / dd;
      const double Ew = hgf[2] * 0.5 * (Erx + Ery);
      // gamma_a = +-(j11 j22 + j12 j21)/4 with j = J0^-1 (quad4r.pyx:3116); the sign (+ - + -) stays with the lanes
      const double w0 = 4. * (J11c * J22c - J12c * J21c);
      const double gam = 0.25 * (J22c * J11c + J12c * J21c) * r[18] * r[18];
      const double wg2 = w0 * gam * gam;
      r[kRecBase + 0] = 1e-6 * K6ROT * cA[5];
      r[kRecBase + 1] = wg2 * Eu;
      r[kRecBase + 2] = wg2 * Ev;
      r[kRecBase + 3] = wg2 * Ew;
      r[kRecBase + 4] = wg2 * Erx;
      r[kRecBase + 5] = wg2 * Ery;
    }
    if (rot) {
      double* ra = r + kRecBase + kHg + (kg_u ? kRecN : 0);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        ra[i] = c.A[i];
        ra[6 + i] = c.B[i];
        ra[12 + i] = c.D[i];
      }
    }
    if (kg_u) {
      // membrane force resultants per Gauss point (quad4.pyx:1965-1967)
#pragma unroll
      for (int gp = 0; gp < 4; ++gp) {
        const int ix = gp >> 1, ie = gp & 1;
        const double xg = ix ? kGpF : -kGpF, eg = ie ? kGpF : -kGpF;
        double exx = 0, eyy = 0, gxy = 0, kxx = 0, kyy = 0, kxy = 0;
#pragma unroll
        for (int cn = 0; cn < 4; ++cn) {
          const double xc = (cn == 1 || cn == 2) ? 1. : -1., ec = (cn >= 2) ? 1. : -1.;
          const double dxc = 0.25 * xc * (1. + ec * eg), dec = 0.25 * ec * (1. + xc * xg);
          const double nx = (J22x[ix] * dxc - J12e[ie] * dec) * idJ[gp];
          const double ny = (-J21x[ix] * dxc + J11e[ie] * dec) * idJ[gp];
          exx += nx * ue[6 * cn];
          eyy += ny * ue[6 * cn + 1];
          gxy += ny * ue[6 * cn] + nx * ue[6 * cn + 1];
          kxx += nx * ue[6 * cn + 4];
          kyy -= ny * ue[6 * cn + 3];
          kxy += ny * ue[6 * cn + 4] - nx * ue[6 * cn + 3];
        }
        r[20 + kHg + gp] = c.A[0] * exx + c.A[1] * eyy + c.A[2] * gxy + c.B[0] * kxx + c.B[1] * kyy + c.B[2] * kxy;
        r[24 + kHg + gp] = c.A[1] * exx + c.A[3] * eyy + c.A[4] * gxy + c.B[1] * kxx + c.B[3] * kyy + c.B[4] * kxy;
        r[28 + kHg + gp] = c.A[2] * exx + c.A[4] * eyy + c.A[5] * gxy + c.B[2] * kxx + c.B[4] * kyy + c.B[5] * kxy;
      }
    }
  }
  __syncwarp();
  double* out = rec + e0 * stride;
  const int total = nvalid * stride;
  if (stride == rec_stride(false, true, false)) {   // the north-star call: division by a constant
    constexpr int kS = rec_stride(false, true, false);
    for (int idx = lane; idx < total; idx += 32) out[idx] = stage[(idx / kS) * ld + idx % kS];
  } else {
    for (int idx = lane; idx < total; idx += 32) out[idx] = stage[(idx / stride) * ld + idx % stride];
  }
}

template <int KIND>
__global__ void __launch_bounds__(128, PF3_K1_CTAS) quad_record_kernel(const EvalArgs A, double* __restrict__ rec, int stride,
                                                                       int64_t e_begin, int64_t e_end) {
  extern __shared__ double k1_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // this launch covers elements [e_begin, e_end)
  const int64_t e0 = e_begin + (int64_t(blockIdx.x) * (blockDim.x >> 5) + warp) * 32;
  if (e0 >= e_end) return;
  conn_prefetch(A.conn, 4, e0, e_end, lane);
  record_warp<KIND>(A, rec, stride, e0, int(min(int64_t(32), e_end - e0)), k1_smem + warp * 32 * (stride + 1), lane);
}

// ------------------------------------------------------------------------------------------ K2
// Shared-memory staging uses the COO slab layout itself (row d, node b, masked column r) with a padded,
// even slab stride, so that (i) a slab leaves as ONE bulk async copy (TMA, cp.async.bulk shared->global)
// issued by the incidence's first lane and (ii) staging stores are conflict-free.
template <int NR, int CNT>
struct SlabShape {
  static constexpr int kSlab = NR * 4 * CNT;   // doubles per (element, node) COO slab
#ifdef PF3_SLAB_PAD2
  static constexpr int kLd = kSlab + 2;        // older stride 146 / 122 / 74 / 38: 2-way conflicts on the staging stores
#else
  // padded stride 152 / 124 / 76 / 44 doubles: even (16-B aligned slabs for the bulk copy) and chosen so that the
  // staging stores of the 8 incidences of a warp fall into disjoint banks (kLd mod 16 = 8 for the 16-B stores of
  // KC0 and for KG, 12 for the 8-B stores of M); the +2 stride was 2-way conflicted on every staging store
  static constexpr int kLd = kSlab + ((CNT == 6 || NR == 3) ? 8 : 4);
#endif
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

// Lanes have written their block into slab (lane>>2).  Ship the slabs to the COO array and reduce the (up to)
// 4 incidences of each half-warp's node into its CSR rows in the fixed order k = 0..3.
template <int NR, int CNT>
__device__ __forceinline__ void emit_slabs(const double* st, const NodeRec* nr, double* __restrict__ coo,
                                           int64_t slab_base, bool act, double* __restrict__ csr, int64_t csr_base,
                                           int nb, bool first_round, int lane, uint64_t pol,
                                           const UnionMap* um = nullptr) {
  constexpr int kSlab = SlabShape<NR, CNT>::kSlab, kLd = SlabShape<NR, CNT>::kLd;
  const int h = lane >> 4, l16 = lane & 15;
#if PF3_COO_TMA
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (coo != nullptr && act && (lane & 3) == 0) {
    const uint32_t src = smem_u32(st + (lane >> 2) * kLd);
#if PF3_L2_HINTS
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(coo + slab_base),
                 "r"(src), "r"(kSlab * 8), "l"(pol)
                 : "memory");
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(coo + slab_base), "r"(src),
                 "r"(kSlab * 8)
                 : "memory");
#endif
  }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#else
  __syncwarp();
  if (coo != nullptr && act) {
    // the 4 lanes of an incidence stream their own slab: 16-B loads/stores, immediate offsets, no index math
    const double2* src = reinterpret_cast<const double2*>(st + (lane >> 2) * kLd) + (lane & 3);
    double2* dst = reinterpret_cast<double2*>(coo + slab_base) + (lane & 3);
#pragma unroll
    for (int i = 0; i < (kSlab / 2 + 3) / 4; ++i)
      if ((kSlab / 2) % 4 == 0 || 4 * i + (lane & 3) < kSlab / 2) dst[4 * i] = src[4 * i];
  }
#endif
  if (nb > 0 && csr != nullptr && um != nullptr) {
    // multi-group plan: this group's entries land inside the UNION row layout (csr_base = b0 * um->mc); positions
    // the group does not have are written as zeros so that the other groups can be added afterwards
    const double* sh = st + h * 4 * kLd;
    if (um->mc == 36) {
      // full 6x6 union blocks (shell + beam meshes): positions outer, the 6 union rows inner, column maps as bit fields
      const int wu = nb * 6;
      double* out = csr + csr_base;
#pragma unroll 1
      for (int x = l16; x < wu; x += 16) {
        const int s = x / 6, c = x - s * 6;
        const unsigned gm = nr->gmap[s];
        int off[4];
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          const int nib = (gm >> (4 * k2)) & 0xF;
          off[k2] = (nib != 0xF) ? (k2 * kLd + nib * CNT) : -1;
        }
#pragma unroll
        for (int du = 0; du < 6; ++du) {
          const int down = um->row[du];
          const int j = int((um->colbits[du] >> (4 * c)) & 0xF);
          double sum = 0.;
          if (down >= 0 && j != 0xF) {
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2)
              if (off[k2] >= 0) sum += sh[off[k2] + down * 4 * CNT + j];
          }
          double* o = out + du * wu + x;
          if (first_round) *o = sum; else *o += sum;   // (a 16-byte two-column variant measured 5 % slower)
        }
      }
    } else
#pragma unroll 1
    for (int du = 0; du < 6; ++du) {
      const int cu = um->cnt[du];
      if (cu == 0) continue;
      const int wu = nb * cu, down = um->row[du];
      double* out = csr + csr_base + um->rowoff[du] * nb;
#pragma unroll 1
      for (int x = l16; x < wu; x += 16) {
        const int s = x / cu, ju = x - s * cu;
        const int j = down >= 0 ? um->col[du][ju] : -1;
        double sum = 0.;
        if (j >= 0) {
          const unsigned gm = nr->gmap[s];
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            const int nib = (gm >> (4 * k2)) & 0xF;
            if (nib != 0xF) sum += sh[k2 * kLd + down * 4 * CNT + nib * CNT + j];
          }
        }
        if (first_round) out[x] = sum; else out[x] += sum;
      }
    }
  } else if (nb > 0 && csr != nullptr) {
    const int w = nb * CNT;
    const double* sh = st + h * 4 * kLd;
    if (CNT % 2 == 0) {
      double* out = csr + csr_base;
#pragma unroll 1
      for (int y = l16; 2 * y < w; y += 16) {
        const int x = 2 * y, s = x / CNT, rr = x - s * CNT;
        const unsigned gm = nr->gmap[s];
        int off[4];
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          const int nib = (gm >> (4 * k2)) & 0xF;
          off[k2] = (nib != 0xF) ? (k2 * kLd + nib * CNT + rr) : -1;
        }
#pragma unroll
        for (int d = 0; d < NR; ++d) {
          double2 sum = make_double2(0., 0.);
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2)
            if (off[k2] >= 0) {
#ifdef PF3_ABL_NORED
              sum.x += double(off[k2]);
#else
              const double2 t = *reinterpret_cast<const double2*>(sh + d * 4 * CNT + off[k2]);
              sum.x += t.x;
              sum.y += t.y;
#endif
            }
          double2* o = reinterpret_cast<double2*>(out + d * w + x);
          if (!first_round) {
            const double2 t = *o;
            sum.x += t.x;
            sum.y += t.y;
          }
          stg_stream(o, sum, pol);
        }
      }
    } else {
      double* out = csr + csr_base;
#pragma unroll 1
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
#ifdef PF3_ABL_NORED
            if (off[k2] >= 0) sum += double(off[k2]);
#else
            if (off[k2] >= 0) sum += sh[d * 4 * CNT + off[k2]];
#endif
          double* o = out + d * w + x;
          if (!first_round) sum += *o;
          stg_stream(o, sum, pol);
        }
      }
    }
  }
  __syncwarp();
}

// The staging area is reused by the next matrix: before writing it again, wait until the bulk copies issued from it
// have READ shared memory.  (A private staging area for KG, so that KG never waits for the previous KC0 copies, was
// measured slower: the wait is back-pressure from the store path, not a latency to hide.)
__device__ __forceinline__ void stage_reuse_wait() {
#if PF3_COO_TMA
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
#endif
}

constexpr int kStageV4 = 8 * SlabShape<6, 6>::kLd;         // 1216 doubles: the largest matrix (KC0)
// CHUNK = consecutive node pairs per warp.  1: one pair per warp, no prefetch (the store-bound three-matrix call: the
// CTA scheduler alone orders the work; measured 10.4 ms against 11.0 / 11.5 ms for chunks of 2 / 4, which widen the
// window of nodes in flight and take L1 away).  kFChunkBig = 2: the next pair's records arrive by cp.async while this
// one is computed (the latency-bound one- and two-matrix calls; with the L2 prefetch in place 2 pairs beat 4, 8 and 1:
// config 3 1.68 ms against 1.77 / 1.84 / 1.69, KC0 only at 4 M Quad4 4.84 against 5.01 / - / 5.07).
#ifndef PF3_CHUNK_BIG
#define PF3_CHUNK_BIG 2
#endif
constexpr int kFChunkBig = PF3_CHUNK_BIG;
#ifndef PF3_CHUNK_VOL
#define PF3_CHUNK_VOL 1400   // calls that write less than this per node (two-matrix calls) take kFChunkBig pairs per CTA
#endif
__host__ __device__ constexpr int fring(int chunk) { return chunk > 1 ? 3 : 1; }
__host__ __device__ constexpr int fbufs(int chunk) { return chunk > 1 ? 2 : 1; }
// per warp: slab staging | ring of node-record pairs (64 B each) | element records of 8 incidences (x2 when
// prefetching), stride rec_ld(rstride) doubles (conflict-free, 16-B aligned)
__host__ __device__ constexpr int warp_smem_doubles(int rstride, int chunk) {
  return kStageV4 + fring(chunk) * 2 * 8 + fbufs(chunk) * 8 * rec_ld(rstride);
}

// Which of the warp's 8 incidence slots holds the record of this lane's element: the two nodes of a pair usually share
// two of their elements (and in element mode all four "incidences" are the same element), so a record is fetched once
// per warp, by the lowest incidence that names it, and read by all of them.
#ifndef PF3_EREC_DEDUP
#define PF3_EREC_DEDUP 1
#endif
__device__ __forceinline__ int erec_owner(int pair0, int lane) {
#if PF3_EREC_DEDUP
  const unsigned same = __match_any_sync(0xffffffffu, pair0 >= 0 ? (pair0 >> 4) : (-1 - lane));
  return (__ffs(same) - 1) >> 2;
#else
  return lane >> 2;
#endif
}

// Stage the element records of the 8 incidences of a node pair into shared memory with 16-B cp.async; the 4 lanes
// of an incidence split the record's chunks.
__device__ __forceinline__ void erec_fetch(const double* __restrict__ rec, int rstride, double* buf, int pair0,
                                           int lane) {
  const int owner = erec_owner(pair0, lane);
  if (pair0 >= 0 && owner == (lane >> 2)) {
#ifdef PF3_ABL_ERECSAME
    const char* src = reinterpret_cast<const char*>(rec + int64_t(lane >> 2) * rstride);
#else
    const char* src = reinterpret_cast<const char*>(rec + int64_t(pair0 >> 4) * rstride);
#endif
    char* dst = reinterpret_cast<char*>(buf + (lane >> 2) * rec_ld(rstride));
    for (int c = (lane & 3) * 16; c < rstride * 8; c += 64)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + c)), "l"(src + c) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// Bring the NodeRec of node pair `np`, round r into shared memory (8 x 16-B cp.async by lanes 0..7).
__device__ __forceinline__ void noderec_fetch(const FusedArgs& F, NodeRec* dst2, int64_t np, int r, int lane) {
  if (lane < 8) {
    const int hh = lane >> 2, q = lane & 3;
    const int64_t n = 2 * np + hh;
    char* dst = reinterpret_cast<char*>(dst2 + hh) + 16 * q;
    if (n < F.nown && F.noderec != nullptr) {
      const char* src = reinterpret_cast<const char*>(F.noderec + n * F.rmax + r) + 16 * q;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    } else if (n < F.nown) {
      // element mode (COO only, no plan): "node" n is element n, its 4 incidences are its own 4 row slabs
      const int p0 = int(n) * 16;
      int4 z = make_int4(-1, -1, -1, -1);
      if (q == 0) z = make_int4(0, 0, p0, p0 + 4);
      if (q == 1) z = make_int4(p0 + 8, p0 + 12, -1, -1);
      if (q == 3) z = make_int4(-1, -1, 4, 0);   // v = 4, nb = 0
      *reinterpret_cast<int4*>(dst) = z;
    } else {
      // empty record: no incidences, no blocks
      // bytes 0-15: b0, inc[0..1] | 16-47: inc[2..3], gmap[0..11] | 48-63: gmap[12..15], v, nb, pad
      int4 z = make_int4(-1, -1, -1, -1);
      if (q == 0) z = make_int4(0, 0, -1, -1);
      if (q == 3) z = make_int4(-1, -1, 0, 0);
      *reinterpret_cast<int4*>(dst) = z;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// One warp = kFChunk consecutive node PAIRS (all their rounds), one CTA = kFusedWarps warps, and as many CTAs as
// there is work: the hardware CTA scheduler hands out node pairs in order, so the nodes in flight at any time form a
// narrow window of the mesh (their COO/CSR destinations are neighbours in DRAM) and a slow warp never holds work back.
// Measured on B200 at 4 M Quad4: persistent grid-stride warps with a 3-deep prefetch ring 12.4 ms/step, this 10.4.
// Within a warp the node records of item j+2 and the element records of item j+1 are in flight (cp.async) while item j
// is evaluated (matters for the latency-bound one- and two-matrix calls, e.g. config 3).
template <int KIND, int CHUNK>
__global__ void __launch_bounds__(32 * kFusedWarps, PF3_FUSED_CTAS) quad_fused_kernel(const FusedArgs F, const double* __restrict__ rec,
                                                                         int rstride) {
  constexpr bool PRE = CHUNK > 1;   // records of the next items arrive by cp.async while this one is evaluated
  constexpr int kFRing = PRE ? 3 : 1, kFBufs = PRE ? 2 : 1;
  extern __shared__ __align__(16) double smem[];
  const EvalArgs& A = F.A;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int eld = rec_ld(rstride);
  double* st = smem + warp * warp_smem_doubles(rstride, PRE ? kFChunkBig : 1);
  NodeRec* ring = reinterpret_cast<NodeRec*>(st + kStageV4);
  double* erec = st + kStageV4 + kFRing * 2 * 8;
  const int h = lane >> 4, l16 = lane & 15, k = l16 >> 2, b = l16 & 3;
  const UnionMap* umK = F.um[0].active ? &F.um[0] : nullptr;
  const UnionMap* umKG = F.um[1].active ? &F.um[1] : nullptr;
  const UnionMap* umM = F.um[2].active ? &F.um[2] : nullptr;
  const int64_t npairs = F.pair_count ? F.pair_first + F.pair_count : (F.nown + 1) >> 1;   // end of this launch's range
  const int rmax = F.rmax;
  const double xib = (b == 1 || b == 2) ? 1. : -1., etab = (b >= 2) ? 1. : -1.;
  // static schedule: this warp owns CHUNK consecutive pairs
  const int64_t np0 = F.pair_first + (int64_t(blockIdx.x) * kFusedWarps + warp) * CHUNK;
  if (np0 >= npairs) return;
  const uint64_t pol = l2_evict_first_policy();
#if PF3_L2_PREFETCH
  // The CTA of the first pair of a chunk pulls the node records and the first-use element records of the chunk
  // kPfAhead chunks ahead into L2 with a few bulk prefetches: the DRAM reads of this kernel then arrive as long bursts
  // instead of 128- and 256-byte reads scattered between its writes (each of which turns the DRAM bus around).
  if (F.pftab != nullptr && lane == 0 && (np0 % kPfChunk) < CHUNK) {
    const int64_t c = np0 / kPfChunk + kPfAhead;
    if (c < F.pf_nchunks) {
      constexpr unsigned kPiece = 32768;
      const int64_t n0 = c * kPfChunk * 2;
      const int64_t nn = min(int64_t(2 * kPfChunk), F.nown - n0);
      const char* q = reinterpret_cast<const char*>(F.noderec + n0 * rmax);
      for (int64_t left = nn * rmax * int64_t(sizeof(NodeRec)); left > 0; left -= kPiece, q += kPiece)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q), "r"(unsigned(min(left, int64_t(kPiece)))) : "memory");
      for (int run = 0; run < kPfRuns; ++run) {
        const int2 t = F.pftab[c * kPfRuns + run];
        q = reinterpret_cast<const char*>(rec + int64_t(t.x) * rstride);
        for (int64_t left = int64_t(t.y) * rstride * 8; left > 0; left -= kPiece, q += kPiece)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q), "r"(unsigned(min(left, int64_t(kPiece)))) : "memory");
      }
    }
  }
#endif
  const int nitems = int(min(int64_t(CHUNK), npairs - np0)) * rmax;   // item j = (pair np0 + j / rmax, round j % rmax)

  auto rec_fetch_pair = [&](int slot3, int64_t pair, int r) {
    if (pair >= 0)
      noderec_fetch(F, ring + slot3 * 2, pair, r, lane);
    else
      asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // (rounds: rmax == 1 unless a node has more than 4 incident elements -- no runtime integer divisions then)
  const bool one_round = rmax == 1;
  auto rec_fetch = [&](int j) {
    rec_fetch_pair(j % kFRing, j < nitems ? np0 + (one_round ? j : j / rmax) : int64_t(-1), one_round ? 0 : j % rmax);
  };
  auto erec_prefetch = [&](int j, bool valid) {
    if (valid)
      erec_fetch(rec, rstride, erec + (j % kFBufs) * 8 * eld, (ring + (j % kFRing) * 2 + h)->inc[k], lane);
    else
      asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if constexpr (CHUNK > 1) {
    rec_fetch(0);
    rec_fetch(1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    erec_prefetch(0, 0 < nitems);
  }

  for (int j = 0;; ++j) {
    if (j >= nitems) break;
    const int r = one_round ? 0 : j % rmax;
    if constexpr (CHUNK > 1) {
      rec_fetch(j + 2);
      asm volatile("cp.async.wait_group 1;" ::: "memory");   // node records j+1 and element records j have landed
      __syncwarp();
      erec_prefetch(j + 1, j + 1 < nitems);
    } else {
      if (j > 0) stage_reuse_wait();
      rec_fetch(j);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    }
    const NodeRec* nr = ring + (j % kFRing) * 2 + h;
    const int pair0 = nr->inc[k];
    const bool act = pair0 >= 0;
    if (__ballot_sync(0xffffffffu, act) == 0u) {
      if (F.zero_empty && r == 0 && nr->nb > 0) {
        // a node only other groups touch: initialise its rows so that their contributions can be added
        const int nbz = nr->nb;
        const int64_t bz = nr->b0;
        if ((A.what & PF3_KC0) && F.csr_kc0) {
          const int mc = umK ? umK->mc : 36;
          for (int x = l16; x < nbz * mc; x += 16) F.csr_kc0[bz * mc + x] = 0.;
        }
        if ((A.what & (PF3_KG | PF3_KG_STRESS)) && F.csr_kg) {
          const int mc = umKG ? umKG->mc : 9;
          for (int x = l16; x < nbz * mc; x += 16) F.csr_kg[bz * mc + x] = 0.;
        }
        if ((A.what & PF3_M) && F.csr_m) {
          const int mc = umM ? umM->mc : (A.mtype == 2 ? 18 : 30);
          for (int x = l16; x < nbz * mc; x += 16) F.csr_m[bz * mc + x] = 0.;
        }
      }
      continue;
    }
    if constexpr (!PRE) {
#ifndef PF3_ABL_NOEREC
      erec_prefetch(j, true);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
#endif
    }
    const int64_t b0 = nr->b0;
    const int nb = nr->nb;
    const int64_t e = act ? (pair0 >> 4) : 0;
    const int a = act ? ((pair0 >> 2) & 3) : 0;
    const bool first = (r == 0);
    const double xia = (a == 1 || a == 2) ? 1. : -1., etaa = (a >= 2) ? 1. : -1.;

    // ---------------- element record (K1) and property row
    const double* re = erec + ((j % kFBufs) * 8 + erec_owner(pair0, lane)) * eld;
    const double2* re2 = reinterpret_cast<const double2*>(re);
    double rr_[kRecBase];
#pragma unroll
    for (int i = 0; i < kRecBase / 2; ++i) {
      const double2 t = re2[i];
      rr_[2 * i] = t.x;
      rr_[2 * i + 1] = t.y;
    }
    Mat3 R;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      R.a[i][0] = rr_[2 * i];
      R.a[i][1] = rr_[2 * i + 1];
    }
    R.a[0][2] = R.a[1][0] * R.a[2][1] - R.a[2][0] * R.a[1][1];   // z = x X y
    R.a[1][2] = R.a[2][0] * R.a[0][1] - R.a[0][0] * R.a[2][1];
    R.a[2][2] = R.a[0][0] * R.a[1][1] - R.a[1][0] * R.a[0][1];
    const double dX10 = rr_[6], dX23 = rr_[7], dX30 = rr_[8], dX21 = rr_[9];
    const double dY10 = rr_[10], dY23 = rr_[11], dY30 = rr_[12], dY21 = rr_[13];
    const double idJ[4] = {rr_[14], rr_[15], rr_[16], rr_[17]};
    const double idJ0 = rr_[18], area = fabs(rr_[19]);
    const bool thick = (KIND == PF3_QUAD4) && rr_[19] < 0.;   // decided once per element by K1
    constexpr int kHg = (KIND == PF3_QUAD4R) ? kRecHg : 0;
    const bool kg_u = (A.what & PF3_KG) != 0;
    const double* prow = A.props + prop_index(A, e) * PF3_SHELLPROP_STRIDE;
    const double* abd = (A.evec != nullptr) ? re + kRecBase + kHg + (kg_u ? kRecN : 0) : prow;

    // Jacobian rows: J11,J12 depend on eta only, J21,J22 on xi only (index 0: -p, 1: +p)
    double J11e[2], J12e[2], J21x[2], J22x[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double t = i ? kGpF : -kGpF;
      J11e[i] = 0.25 * ((1. - t) * dX10 + (1. + t) * dX23);
      J12e[i] = 0.25 * ((1. - t) * dY10 + (1. + t) * dY23);
      J21x[i] = 0.25 * ((1. - t) * dX30 + (1. + t) * dX21);
      J22x[i] = 0.25 * ((1. - t) * dY30 + (1. + t) * dY21);
    }

    // gradient Gram of the pair over the 2x2 Gauss points, the mixed N / N,x sums, and Ge_ab
    double gxx = 0., gxy = 0., gyx = 0., gyy = 0., pyab = 0., pxab = 0., pyba = 0., pxba = 0., hab = 0., ge = 0.;
#pragma unroll
    for (int gp = 0; gp < 4; ++gp) {
      const int ix = gp >> 1, ie = gp & 1;
      const double xg = ix ? kGpF : -kGpF, eg = ie ? kGpF : -kGpF;
      const double dxa = 0.25 * xia * (1. + etaa * eg), dea = 0.25 * etaa * (1. + xia * xg);
      const double dxb = 0.25 * xib * (1. + etab * eg), deb = 0.25 * etab * (1. + xib * xg);
      const double wxa = J22x[ix] * dxa - J12e[ie] * dea, wya = -J21x[ix] * dxa + J11e[ie] * dea;
      const double wxb = J22x[ix] * dxb - J12e[ie] * deb, wyb = -J21x[ix] * dxb + J11e[ie] * deb;
      const double na = 0.25 * (1. + xia * xg) * (1. + etaa * eg), nbv = 0.25 * (1. + xib * xg) * (1. + etab * eg);
      const double vax = wxa * idJ[gp], vay = wya * idJ[gp];
      gxx += vax * wxb;
      gxy += vax * wyb;
      gyx += vay * wxb;
      gyy += vay * wyb;
      pyab += wya * nbv;
      pxab += wxa * nbv;
      pyba += wyb * na;
      pxba += wxb * na;
      hab += (na * nbv) * (J11e[ie] * J22x[ix] - J12e[ie] * J21x[ix]);
      if (kg_u) {
        const double nxx = re[20 + kHg + gp], nyy = re[24 + kHg + gp], nxy = re[28 + kHg + gp];
        ge += wxb * (vax * nxx + vay * nxy) + wyb * (vax * nxy + vay * nyy);
      }
    }
    if (A.what & PF3_KG_STRESS) ge = A.Nxx * gxx + A.Nxy * (gxy + gyx) + A.Nyy * gyy;

    // ---------------- KG : Ge_ab * z z^T on the translations
    if (A.what & (PF3_KG | PF3_KG_STRESS)) {
      double* sl = st + (lane >> 2) * SlabShape<3, 3>::kLd + b * 3;
      stage_reuse_wait();
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) sl[i * 12 + jj] = (R.a[i][2] * R.a[jj][2]) * ge;
      emit_slabs<3, 3>(st, nr, A.kgv ? A.kgv + A.kg_k0 : nullptr, e * 144 + a * 36, act, F.csr_kg,
                       umKG ? b0 * umKG->mc : b0 * 9, nb, first, lane, pol, umKG);
    }

    // ---------------- M : H_ab * (T6 m_l T6^T)
    if (A.what & PF3_M) {
      double H;
      if (A.mtype == 0) {
        H = hab;
      } else if (A.mtype == 1) {
        H = 0.0625 * area;
      } else {
        // Gauss-Lobatto: detJ at node a on the diagonal, zero elsewhere (quad4.pyx:8873)
        const double J11n = 0.25 * ((1. - etaa) * dX10 + (1. + etaa) * dX23), J12n = 0.25 * ((1. - etaa) * dY10 + (1. + etaa) * dY23);
        const double J21n = 0.25 * ((1. - xia) * dX30 + (1. + xia) * dX21), J22n = 0.25 * ((1. - xia) * dY30 + (1. + xia) * dY21);
        H = (a == b) ? (J11n * J22n - J12n * J21n) : 0.;
      }
      const double r0 = prow[24], r1 = prow[25], r2 = prow[26];
      double* coo = A.mv ? A.mv + A.m_k0 : nullptr;
      if (A.mtype != 2) {
        double* sl = st + (lane >> 2) * SlabShape<6, 5>::kLd + b * 5;
        stage_reuse_wait();
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double tt[3], tr[3], rq[3];
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) {
            tt[jj] = r0 * R.a[i][0] * R.a[jj][0] + r0 * R.a[i][1] * R.a[jj][1] + r0 * R.a[i][2] * R.a[jj][2];
            tr[jj] = r1 * R.a[i][0] * R.a[jj][1] - r1 * R.a[i][1] * R.a[jj][0];
            rq[jj] = r2 * R.a[i][0] * R.a[jj][0] + r2 * R.a[i][1] * R.a[jj][1];
          }
          sl[i * 20 + 0] = H * tt[0];
          sl[i * 20 + 1] = H * tt[1];
          sl[i * 20 + 2] = H * tt[2];
          sl[i * 20 + 3] = H * tr[(i == 0) ? 1 : 0];
          sl[i * 20 + 4] = H * tr[(i == 2) ? 1 : 2];
          // row 3+i: rt = tr^T = -tr (columns j != i), then rr
          sl[(3 + i) * 20 + 0] = -(H * tr[(i == 0) ? 1 : 0]);
          sl[(3 + i) * 20 + 1] = -(H * tr[(i == 2) ? 1 : 2]);
          sl[(3 + i) * 20 + 2] = H * rq[0];
          sl[(3 + i) * 20 + 3] = H * rq[1];
          sl[(3 + i) * 20 + 4] = H * rq[2];
        }
        emit_slabs<6, 5>(st, nr, coo, e * 480 + a * 120, act, F.csr_m, umM ? b0 * umM->mc : b0 * 30, nb, first, lane, pol, umM);
      } else {
        double* sl = st + (lane >> 2) * SlabShape<6, 3>::kLd + b * 3;
        stage_reuse_wait();
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) {
            sl[i * 12 + jj] = H * (r0 * R.a[i][0] * R.a[jj][0] + r0 * R.a[i][1] * R.a[jj][1] + r0 * R.a[i][2] * R.a[jj][2]);
            sl[(3 + i) * 12 + jj] = H * (r2 * R.a[i][0] * R.a[jj][0] + r2 * R.a[i][1] * R.a[jj][1]);
          }
        emit_slabs<6, 3>(st, nr, coo, e * 480 + a * 72, act, F.csr_m, umM ? b0 * umM->mc : b0 * 18, nb, first, lane, pol, umM);
      }
    }

    // ---------------- KC0 : the 6x6 block (a, b)
    if (A.what & PF3_KC0) {
      double cA[6], cB[6], cD[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        cA[i] = abd[i];
        cB[i] = abd[6 + i];
        cD[i] = abd[12 + i];
      }
      const double J11c = 0.25 * (dX10 + dX23), J12c = 0.25 * (dY10 + dY23);
      const double J21c = 0.25 * (dX30 + dX21), J22c = 0.25 * (dY30 + dY21);
      const double w0 = 4. * (J11c * J22c - J12c * J21c);
      const double N0xa = (J22c * 0.25 * xia - J12c * 0.25 * etaa) * idJ0;
      const double N0ya = (-J21c * 0.25 * xia + J11c * 0.25 * etaa) * idJ0;
      const double N0xb = (J22c * 0.25 * xib - J12c * 0.25 * etab) * idJ0;
      const double N0yb = (-J21c * 0.25 * xib + J11c * 0.25 * etab) * idJ0;
      const double k13 = prow[21], k23 = prow[22];
      const double E44 = prow[18] * k23, E45 = prow[19] * 0.5 * (k13 + k23), E55 = prow[20] * k13;
      const double sgn = ((a ^ b) & 1) ? -1. : 1.;
      double kd = 1., hg0 = 0., hg1 = 0., hg2 = 0., hg3 = 0., hg4 = 0.;
      if (KIND == PF3_QUAD4R) {   // per-element constants from the record (K1), the sign of the hourglass vector here
        const double2* hq = reinterpret_cast<const double2*>(re + kRecBase);
        const double2 h01 = hq[0], h23 = hq[1], h45 = hq[2];
        kd = h01.x;
        hg0 = sgn * h01.y;
        hg1 = sgn * h23.x;
        hg2 = sgn * h23.y;
        hg3 = sgn * h45.x;
        hg4 = sgn * h45.y;
      }
      // constitutive Gram: 2x2 Gauss (Quad4) or centre point with weight 4 detJ0 (Quad4R)
      double cxx = gxx, cxy = gxy, cyx = gyx, cyy = gyy;
      if (KIND == PF3_QUAD4R) {
        const double wa = w0 * N0xa, wb = w0 * N0ya;
        cxx = wa * N0xb;
        cxy = wa * N0yb;
        cyx = wb * N0xb;
        cyy = wb * N0yb;
      }
      const double tSa = w0 * (E44 * N0ya + E45 * N0xa), sSa = w0 * (E45 * N0ya + E55 * N0xa);
      const double tSb = w0 * (E44 * N0yb + E45 * N0xb), sSb = w0 * (E45 * N0yb + E55 * N0xb);
      double2* sl2 = reinterpret_cast<double2*>(st + (lane >> 2) * SlabShape<6, 6>::kLd + b * 6);
      double o1[3][3], o2[3][3];
      {
        const double uu = f_pp(cA, cxx, cxy, cyx, cyy) + 0.25 * kd * gyy + hg0;
        const double uv = f_pq(cA, cxx, cxy, cyx, cyy) - 0.25 * kd * gyx;
        const double vu = f_qp(cA, cxx, cxy, cyx, cyy) - 0.25 * kd * gxy;
        const double vv = f_qq(cA, cxx, cxy, cyx, cyy) + 0.25 * kd * gxx + hg1;
        const double ww = (thick ? (E44 * gyy + E45 * (gxy + gyx) + E55 * gxx) : (tSa * N0yb + sSa * N0xb)) + hg2;
        rot_block_diag5(R, uu, uv, vu, vv, ww, o1);
      }
      rot_block_8(R, -f_pq(cB, cxx, cxy, cyx, cyy), f_pp(cB, cxx, cxy, cyx, cyy), 0.5 * kd * pyab,
                  -f_qq(cB, cxx, cxy, cyx, cyy), f_qp(cB, cxx, cxy, cyx, cyy), -0.5 * kd * pxab, -0.25 * tSa,
                  0.25 * sSa, o2);
      stage_reuse_wait();
#pragma unroll
      for (int i = 0; i < 3; ++i) {   // rows u v w of node a: 6 columns of node b, 24 doubles per COO row
        sl2[i * 12 + 0] = make_double2(o1[i][0], o1[i][1]);
        sl2[i * 12 + 1] = make_double2(o1[i][2], o2[i][0]);
        sl2[i * 12 + 2] = make_double2(o2[i][1], o2[i][2]);
      }
      rot_block_8(R, -f_qp(cB, cxx, cxy, cyx, cyy), -f_qq(cB, cxx, cxy, cyx, cyy), -0.25 * tSb,
                  f_pp(cB, cxx, cxy, cyx, cyy), f_pq(cB, cxx, cxy, cyx, cyy), 0.25 * sSb, 0.5 * kd * pyba,
                  -0.5 * kd * pxba, o1);
      {
        const double c44 = w0 * E44 * 0.0625, c45 = w0 * E45 * 0.0625, c55 = w0 * E55 * 0.0625;
        const double rxrx = f_qq(cD, cxx, cxy, cyx, cyy) + c44 + hg3;
        const double rxry = -f_qp(cD, cxx, cxy, cyx, cyy) - c45;
        const double ryrx = -f_pq(cD, cxx, cxy, cyx, cyy) - c45;
        const double ryry = f_pp(cD, cxx, cxy, cyx, cyy) + c55 + hg4;
        rot_block_diag5(R, rxrx, rxry, ryrx, ryry, kd * hab, o2);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        sl2[(3 + i) * 12 + 0] = make_double2(o1[i][0], o1[i][1]);
        sl2[(3 + i) * 12 + 1] = make_double2(o1[i][2], o2[i][0]);
        sl2[(3 + i) * 12 + 2] = make_double2(o2[i][1], o2[i][2]);
      }
      emit_slabs<6, 6>(st, nr, A.kc0v ? A.kc0v + A.kc0_k0 : nullptr, e * 576 + a * 144, act, F.csr_kc0,
                       umK ? b0 * umK->mc : b0 * 36, nb, first, lane, pol, umK);
    }
  }
  // the bulk copies read this CTA's shared memory: they must have done so before the CTA retires
  stage_reuse_wait();
}

}  // namespace

size_t fused_smem_bytes(int rstride, int chunk) { return size_t(kFusedWarps) * warp_smem_doubles(rstride, chunk) * sizeof(double); }
int fused_max_slots() { return kMaxSlots; }
int fused_record_stride(int kind, const EvalArgs& A) {
  return rec_stride(kind == PF3_QUAD4R, (A.what & PF3_KG) != 0, A.evec != nullptr);
}

namespace {

cudaError_t launch_k1(int kind, const EvalArgs& A, double* rec, int stride, int64_t e_begin, int64_t e_end, int warps,
                      cudaStream_t st, int64_t* launches) {
  if (e_end <= e_begin) return cudaSuccess;
  static PerDeviceOnce once1;
  if (once1.first()) {
    const int maxs = int(size_t(4) * 32 * (kRecMax + 1) * sizeof(double));
    cudaFuncSetAttribute(quad_record_kernel<PF3_QUAD4>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs);
    cudaFuncSetAttribute(quad_record_kernel<PF3_QUAD4R>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs);
  }
  const int64_t per = int64_t(warps) * 32;
  const unsigned g1 = unsigned((e_end - e_begin + per - 1) / per);
  const size_t smem1 = size_t(warps) * 32 * (stride + 1) * sizeof(double);
  if (kind == PF3_QUAD4)
    quad_record_kernel<PF3_QUAD4><<<g1, 32 * warps, smem1, st>>>(A, rec, stride, e_begin, e_end);
  else
    quad_record_kernel<PF3_QUAD4R><<<g1, 32 * warps, smem1, st>>>(A, rec, stride, e_begin, e_end);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace

// rec: device scratch of ne * fused_record_stride doubles.  phases: bit 0 = K1 (records of ALL elements), bit 1 = K2 for
// the node pairs [F.pair_first, F.pair_first + F.pair_count) (all pairs when pair_count == 0).
cudaError_t launch_quad_fused(int kind, const FusedArgs& F, double* rec, cudaStream_t st, int64_t* launches, int phases) {
  if (F.nown <= 0 || F.A.ne <= 0) return cudaSuccess;
  const int stride = fused_record_stride(kind, F.A);
  // doubles written per element (COO + CSR share): the three-matrix north-star call is store-bound, smaller calls are
  // latency-bound and take the prefetching variant
  const int w = F.A.what;
  const int vol = ((w & PF3_KC0) ? 900 : 0) + ((w & (PF3_KG | PF3_KG_STRESS)) ? 225 : 0) + ((w & PF3_M) ? 750 : 0);
  const bool mapped = F.um[0].active || F.um[1].active || F.um[2].active;
#ifdef PF3_FORCE_CHUNK4
  const int chunk = !mapped ? kFChunkBig : 1;
#else
  const int chunk = (vol < PF3_CHUNK_VOL && !mapped) ? kFChunkBig : 1;
#endif
  if (phases & 1) {
    cudaError_t e1 = launch_k1(kind, F.A, rec, stride, 0, F.A.ne, 4, st, launches);
    if (e1 != cudaSuccess) return e1;
  }
  if (!(phases & 2)) return cudaSuccess;
  const size_t smem = fused_smem_bytes(stride, chunk);
  const int64_t npairs = F.pair_count ? F.pair_count : (F.nown + 1) / 2;
  const int64_t want = (npairs + int64_t(kFusedWarps) * chunk - 1) / (int64_t(kFusedWarps) * chunk);
  if (want > int64_t(0x7fffffff)) return cudaErrorInvalidConfiguration;
  const unsigned grid = unsigned(want < 1 ? 1 : want);
  static PerDeviceOnce once;
  if (once.first()) {
    const int m1 = int(fused_smem_bytes(kRecMax, 1)), m4 = int(fused_smem_bytes(kRecMax, kFChunkBig));
    cudaFuncSetAttribute(quad_fused_kernel<PF3_QUAD4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, m1);
    cudaFuncSetAttribute(quad_fused_kernel<PF3_QUAD4R, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, m1);
    cudaFuncSetAttribute(quad_fused_kernel<PF3_QUAD4, kFChunkBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, m4);
    cudaFuncSetAttribute(quad_fused_kernel<PF3_QUAD4R, kFChunkBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, m4);
  }
  const unsigned nt = 32 * kFusedWarps;
  if (kind == PF3_QUAD4) {
    if (chunk == 1) quad_fused_kernel<PF3_QUAD4, 1><<<grid, nt, smem, st>>>(F, rec, stride);
    else quad_fused_kernel<PF3_QUAD4, kFChunkBig><<<grid, nt, smem, st>>>(F, rec, stride);
  } else {
    if (chunk == 1) quad_fused_kernel<PF3_QUAD4R, 1><<<grid, nt, smem, st>>>(F, rec, stride);
    else quad_fused_kernel<PF3_QUAD4R, kFChunkBig><<<grid, nt, smem, st>>>(F, rec, stride);
  }
  ++*launches;
  return cudaGetLastError();
}

}  // namespace pf3
