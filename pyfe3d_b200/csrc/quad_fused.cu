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
#include "shell.cuh"

namespace pf3 {

namespace {

constexpr double kGpF = 0.5773502691896257645092;
constexpr int kFusedWarps = 4;
constexpr int kLd = 36;                     // staging leading dimension: 36 mod 16 = 4 -> <=2-way LDS conflicts
constexpr int kStageDoubles = 36 * kLd;     // one 6x6 block per lane
constexpr int kMaxSlots = 16;               // column blocks per node row supported by the fused path
constexpr int kWarpSmemDoubles = kStageDoubles + (8 * kMaxSlots) / 8;
constexpr int kRecPlain = 36;               // record doubles without / with rotated A,B,D
constexpr int kRecRot = 56;

// ------------------------------------------------------------------------------------------ K1
template <int KIND>
__global__ void __launch_bounds__(128) quad_record_kernel(const EvalArgs A, double* __restrict__ rec, int stride) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= A.ne) return;
  const bool kg_u = (A.what & PF3_KG) != 0;
  double ue[24];
  ShellGeom<4> g;
  shell_geom<4>(A, e, g, kg_u ? ue : nullptr);
  double* r = rec + e * stride;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r[3 * i + j] = g.R.a[i][j];
  const double d[8] = {g.X[1] - g.X[0], g.X[2] - g.X[3], g.X[3] - g.X[0], g.X[2] - g.X[1],
                       g.Y[1] - g.Y[0], g.Y[2] - g.Y[3], g.Y[3] - g.Y[0], g.Y[2] - g.Y[1]};
#pragma unroll
  for (int i = 0; i < 8; ++i) r[9 + i] = d[i];
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
    r[17 + gp] = idJ[gp];
  }
  const double J11c = 0.25 * (d[0] + d[1]), J12c = 0.25 * (d[4] + d[5]);
  const double J21c = 0.25 * (d[2] + d[3]), J22c = 0.25 * (d[6] + d[7]);
  r[21] = 1. / (J11c * J22c - J12c * J21c);
  r[22] = g.area;
  r[23] = 0.;
  if (kg_u || stride == kRecRot) {
    ShellCoef c;
    shell_coef<4>(A, e, g, c);
    if (stride == kRecRot) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        r[36 + i] = c.A[i];
        r[42 + i] = c.B[i];
        r[48 + i] = c.D[i];
      }
      r[54] = 0.;
      r[55] = 0.;
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
        r[24 + gp] = c.A[0] * exx + c.A[1] * eyy + c.A[2] * gxy + c.B[0] * kxx + c.B[1] * kyy + c.B[2] * kxy;
        r[28 + gp] = c.A[1] * exx + c.A[3] * eyy + c.A[4] * gxy + c.B[1] * kxx + c.B[3] * kyy + c.B[4] * kxy;
        r[32 + gp] = c.A[2] * exx + c.A[4] * eyy + c.A[5] * gxy + c.B[2] * kxx + c.B[4] * kyy + c.B[5] * kxy;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ K2
template <int NR, int CNT>
struct EmitIdx {
  static constexpr int kSlab = NR * 4 * CNT;           // doubles per (element, node) COO slab
  static constexpr int kCooIt = (kSlab + 31) / 32;
  static constexpr int kCsrIt = (kMaxSlots * CNT + 15) / 16;
  int coo_src[kCooIt];   // staging offset of slab entry p = lane + 32 i (without the incidence column)
  __device__ __forceinline__ void init(int lane) {
#pragma unroll
    for (int i = 0; i < kCooIt; ++i) {
      const int p = lane + 32 * i;
      const int d = p / (4 * CNT), rem = p - d * (4 * CNT);
      const int bb = rem / CNT, rr = rem - bb * CNT;
      coo_src[i] = (d * CNT + rr) * kLd + bb;
    }
  }
};

// The lanes have staged their block at st[t*kLd + lane], t = d*CNT + r.  Stream the 8 slabs to the COO
// array and reduce the 4 incidences of each half-warp's node into its CSR rows.
template <int NR, int CNT>
__device__ __forceinline__ void emit_staged(const EmitIdx<NR, CNT>& I, const double* st, const signed char* inv,
                                            double* __restrict__ coo, int64_t slab_base, bool act,
                                            double* __restrict__ csr, int64_t csr_base, int nb, bool first_round,
                                            int lane) {
  const int h = lane >> 4, l16 = lane & 15;
  __syncwarp();
  if (coo != nullptr) {
#pragma unroll 1
    for (int idx = 0; idx < 8; ++idx) {
      const int64_t base = __shfl_sync(0xffffffffu, slab_base, idx * 4);
      const int on = __shfl_sync(0xffffffffu, act ? 1 : 0, idx * 4);
      if (on) {
        double* dst = coo + base + lane;
        const double* src = st + idx * 4;
#pragma unroll
        for (int i = 0; i < EmitIdx<NR, CNT>::kCooIt; ++i)
          if (lane + 32 * i < EmitIdx<NR, CNT>::kSlab) dst[32 * i] = src[I.coo_src[i]];
      }
    }
  }
  if (nb > 0 && csr != nullptr) {
    const int w = nb * CNT;
    const signed char* iv = inv + h * 4 * kMaxSlots;
    const double* sh = st + h * 16;
    double* out = csr + csr_base + l16;
#pragma unroll 1
    for (int x = l16; x < w; x += 16) {
      const int s = x / CNT, rr = x - s * CNT;
      int off[4];
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        const int bs = iv[k2 * kMaxSlots + s];
        off[k2] = (bs >= 0) ? (k2 * 4 + bs + rr * kLd) : -1;
      }
#pragma unroll
      for (int d = 0; d < NR; ++d) {
        double sum = 0.;
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2)
          if (off[k2] >= 0) sum += sh[d * CNT * kLd + off[k2]];
        double* o = out + d * w + (x - l16);
        if (first_round) *o = sum; else *o += sum;
      }
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void stage9(double* my, const double (*o)[3], int row0, int col0, int cnt) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) my[((row0 + i) * cnt + col0 + j) * kLd] = o[i][j];
}

struct NodeWork {   // what a lane needs to know about its node / incidence before touching any double
  int64_t b0, pair0;
  int packed;       // v | nb << 8 | (myslot + 1) << 16
  __device__ __forceinline__ int v() const { return packed & 0xff; }
  __device__ __forceinline__ int nb() const { return (packed >> 8) & 0xff; }
  __device__ __forceinline__ int myslot() const { return (packed >> 16) - 1; }
};

__device__ __forceinline__ NodeWork fetch_work(const FusedArgs& F, int64_t np, int64_t npairs, int h, int k, int b,
                                               int round) {
  NodeWork w;
  w.b0 = 0;
  w.pair0 = 0;
  w.packed = 0;
  const int64_t n = 2 * np + h;
  if (np < npairs && n < F.nown) {
    const int64_t q0 = F.inc_ptr[n];
    const int v = int(F.inc_ptr[n + 1] - q0);
    w.b0 = F.brow_ptr[n];
    const int nb = int(F.brow_ptr[n + 1] - w.b0);
    const int kk = 4 * round + k;
    int ms = -1;
    if (kk < v) {
      w.pair0 = F.inc_pair0[q0 + kk];
      ms = F.slot[w.pair0 + b];
    }
    w.packed = v | (nb << 8) | ((ms + 1) << 16);
  }
  return w;
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int KIND>
__global__ void __launch_bounds__(32 * kFusedWarps, 3) quad_fused_kernel(const FusedArgs F, const double* __restrict__ rec,
                                                                         int rstride) {
  extern __shared__ double smem[];
  const EvalArgs& A = F.A;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* st = smem + warp * kWarpSmemDoubles;
  double* my = st + lane;
  signed char* inv = reinterpret_cast<signed char*>(st + kStageDoubles);
  const int h = lane >> 4, l16 = lane & 15, k = l16 >> 2, b = l16 & 3;
  const int64_t npairs = (F.nown + 1) >> 1;
  const double xib = (b == 1 || b == 2) ? 1. : -1., etab = (b >= 2) ? 1. : -1.;
  EmitIdx<6, 6> I66;
  EmitIdx<6, 5> I65;
  EmitIdx<6, 3> I63;
  EmitIdx<3, 3> I33;
  I66.init(lane);
  I65.init(lane);
  I63.init(lane);
  I33.init(lane);
  const int64_t stride_np = int64_t(gridDim.x) * kFusedWarps;
  int64_t np = int64_t(blockIdx.x) * kFusedWarps + warp;
  // software pipeline: index chain two node pairs ahead, element record lines one pair ahead (L1 prefetch)
  NodeWork nxt = fetch_work(F, np, npairs, h, k, b, 0);
  NodeWork nxt2 = fetch_work(F, np + stride_np, npairs, h, k, b, 0);

  for (; np < npairs; np += stride_np) {
    const NodeWork cur0 = nxt;
    nxt = nxt2;
    nxt2 = fetch_work(F, np + 2 * stride_np, npairs, h, k, b, 0);
    if (nxt.myslot() >= 0) {
      const char* pr = reinterpret_cast<const char*>(rec + (nxt.pair0 >> 4) * rstride);
      for (int off = b * 128; off < rstride * 8; off += 512) prefetch_l1(pr + off);
    }
    const int vmax = max(__shfl_sync(0xffffffffu, cur0.v(), 0), __shfl_sync(0xffffffffu, cur0.v(), 16));
    const int rounds = (vmax + 3) >> 2;

    for (int r = 0; r < rounds; ++r) {
      const NodeWork cur = (r == 0) ? cur0 : fetch_work(F, np, npairs, h, k, b, r);
      const bool act = cur.myslot() >= 0;
      const int64_t e = cur.pair0 >> 4;
      const int a = int(cur.pair0 >> 2) & 3;
      const int64_t b0 = cur.b0;
      const int nb = cur.nb();
      for (int i = lane; i < 8 * kMaxSlots; i += 32) inv[i] = -1;
      __syncwarp();
      if (act) inv[(h * 4 + k) * kMaxSlots + cur.myslot()] = (signed char)b;
      const bool first = (r == 0);
      const double xia = (a == 1 || a == 2) ? 1. : -1., etaa = (a >= 2) ? 1. : -1.;

      // ---------------- element record (K1) and property row
      const double* re = rec + e * rstride;
      Mat3 R;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R.a[i][j] = re[3 * i + j];
      const double dX10 = re[9], dX23 = re[10], dX30 = re[11], dX21 = re[12];
      const double dY10 = re[13], dY23 = re[14], dY30 = re[15], dY21 = re[16];
      double idJ[4];
#pragma unroll
      for (int gp = 0; gp < 4; ++gp) idJ[gp] = re[17 + gp];
      const double idJ0 = re[21], area = re[22];
      const double* prow = A.props + int64_t(A.prop_id ? A.prop_id[e] : 0) * PF3_SHELLPROP_STRIDE;
      const double* abd = (rstride == kRecRot) ? re + 36 : prow;

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
      const bool kg_u = (A.what & PF3_KG) != 0;

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
          const double nxx = re[24 + gp], nyy = re[28 + gp], nxy = re[32 + gp];
          ge += wxb * (vax * nxx + vay * nxy) + wyb * (vax * nxy + vay * nyy);
        }
      }
      if (A.what & PF3_KG_STRESS) ge = A.Nxx * gxx + A.Nxy * (gxy + gyx) + A.Nyy * gyy;

      // ---------------- KG : Ge_ab * z z^T on the translations
      if (A.what & (PF3_KG | PF3_KG_STRESS)) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) my[(i * 3 + j) * kLd] = (R.a[i][2] * R.a[j][2]) * ge;
        emit_staged<3, 3>(I33, st, inv, A.kgv ? A.kgv + A.kg_k0 : nullptr, e * 144 + a * 36, act, F.csr_kg, b0 * 9,
                          nb, first, lane);
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
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            double tt[3], tr[3], rr[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              tt[j] = r0 * R.a[i][0] * R.a[j][0] + r0 * R.a[i][1] * R.a[j][1] + r0 * R.a[i][2] * R.a[j][2];
              tr[j] = r1 * R.a[i][0] * R.a[j][1] - r1 * R.a[i][1] * R.a[j][0];
              rr[j] = r2 * R.a[i][0] * R.a[j][0] + r2 * R.a[i][1] * R.a[j][1];
            }
            my[(i * 5 + 0) * kLd] = H * tt[0];
            my[(i * 5 + 1) * kLd] = H * tt[1];
            my[(i * 5 + 2) * kLd] = H * tt[2];
            my[(i * 5 + 3) * kLd] = H * tr[(i == 0) ? 1 : 0];
            my[(i * 5 + 4) * kLd] = H * tr[(i == 2) ? 1 : 2];
            // row 3+i: rt = tr^T = -tr (columns j != i), then rr
            my[((3 + i) * 5 + 0) * kLd] = -(H * tr[(i == 0) ? 1 : 0]);
            my[((3 + i) * 5 + 1) * kLd] = -(H * tr[(i == 2) ? 1 : 2]);
            my[((3 + i) * 5 + 2) * kLd] = H * rr[0];
            my[((3 + i) * 5 + 3) * kLd] = H * rr[1];
            my[((3 + i) * 5 + 4) * kLd] = H * rr[2];
          }
          emit_staged<6, 5>(I65, st, inv, coo, e * 480 + a * 120, act, F.csr_m, b0 * 30, nb, first, lane);
        } else {
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              my[(i * 3 + j) * kLd] = H * (r0 * R.a[i][0] * R.a[j][0] + r0 * R.a[i][1] * R.a[j][1] + r0 * R.a[i][2] * R.a[j][2]);
              my[((3 + i) * 3 + j) * kLd] = H * (r2 * R.a[i][0] * R.a[j][0] + r2 * R.a[i][1] * R.a[j][1]);
            }
          emit_staged<6, 3>(I63, st, inv, coo, e * 480 + a * 72, act, F.csr_m, b0 * 18, nb, first, lane);
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
        const double k13 = prow[21], k23 = prow[22], hh = prow[23];
        const double E44 = prow[18] * k23, E45 = prow[19] * 0.5 * (k13 + k23), E55 = prow[20] * k13;
        const double sgn = ((a ^ b) & 1) ? -1. : 1.;
        double kd = 1., hg0 = 0., hg1 = 0., hg2 = 0., hg3 = 0., hg4 = 0.;
        if (KIND == PF3_QUAD4R) {
          double K6ROT = 100., hgf[5] = {1., 1., 1., 1., 1.};
          if (A.eparam != nullptr) {
            const double* ep = A.eparam + e * PF3_EPARAM_STRIDE;
            K6ROT = ep[0];
#pragma unroll
            for (int d = 0; d < 5; ++d) hgf[d] = ep[2 + d];
          }
          kd = 1e-6 * K6ROT * cA[5];
          const double den = -cA[0] * cA[3] * cA[5] + cA[0] * cA[4] * cA[4] + cA[1] * cA[1] * cA[5] -
                             2 * cA[1] * cA[2] * cA[4] + cA[2] * cA[2] * cA[3];
          const double a11 = (-cA[3] * cA[5] + cA[4] * cA[4]) / den, a22 = (-cA[0] * cA[5] + cA[2] * cA[2]) / den;
          const double E1eq = 1. / (hh * a11), E2eq = 1. / (hh * a22);
          const double dd = 1.0 + 1.0 / area;
          const double Eu = hgf[0] * 0.1 * E1eq * hh / dd, Ev = hgf[1] * 0.1 * E2eq * hh / dd;
          const double Erx = hgf[3] * 0.1 * E2eq * hh * hh * hh / dd, Ery = hgf[4] * 0.1 * E1eq * hh * hh * hh / dd;
          const double Ew = hgf[2] * 0.5 * (Erx + Ery);
          // gamma_a = +-(j11 j22 + j12 j21)/4 with j = J0^-1 (quad4r.pyx:3116)
          const double gam = 0.25 * (J22c * J11c + J12c * J21c) * idJ0 * idJ0;
          const double wg2 = sgn * w0 * gam * gam;
          hg0 = wg2 * Eu;
          hg1 = wg2 * Ev;
          hg2 = wg2 * Ew;
          hg3 = wg2 * Erx;
          hg4 = wg2 * Ery;
        }
        const bool thick = (KIND == PF3_QUAD4) && (hh / sqrt(area) >= 1.);
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
        double o[3][3];
        {
          const double uu = f_pp(cA, cxx, cxy, cyx, cyy) + 0.25 * kd * gyy + hg0;
          const double uv = f_pq(cA, cxx, cxy, cyx, cyy) - 0.25 * kd * gyx;
          const double vu = f_qp(cA, cxx, cxy, cyx, cyy) - 0.25 * kd * gxy;
          const double vv = f_qq(cA, cxx, cxy, cyx, cyy) + 0.25 * kd * gxx + hg1;
          const double ww = (thick ? (E44 * gyy + E45 * (gxy + gyx) + E55 * gxx) : (tSa * N0yb + sSa * N0xb)) + hg2;
          rot_block_diag5(R, uu, uv, vu, vv, ww, o);
          stage9(my, o, 0, 0, 6);
        }
        rot_block_8(R, -f_pq(cB, cxx, cxy, cyx, cyy), f_pp(cB, cxx, cxy, cyx, cyy), 0.5 * kd * pyab,
                    -f_qq(cB, cxx, cxy, cyx, cyy), f_qp(cB, cxx, cxy, cyx, cyy), -0.5 * kd * pxab, -0.25 * tSa,
                    0.25 * sSa, o);
        stage9(my, o, 0, 3, 6);
        rot_block_8(R, -f_qp(cB, cxx, cxy, cyx, cyy), -f_qq(cB, cxx, cxy, cyx, cyy), -0.25 * tSb,
                    f_pp(cB, cxx, cxy, cyx, cyy), f_pq(cB, cxx, cxy, cyx, cyy), 0.25 * sSb, 0.5 * kd * pyba,
                    -0.5 * kd * pxba, o);
        stage9(my, o, 3, 0, 6);
        {
          const double c44 = w0 * E44 * 0.0625, c45 = w0 * E45 * 0.0625, c55 = w0 * E55 * 0.0625;
          const double rxrx = f_qq(cD, cxx, cxy, cyx, cyy) + c44 + hg3;
          const double rxry = -f_qp(cD, cxx, cxy, cyx, cyy) - c45;
          const double ryrx = -f_pq(cD, cxx, cxy, cyx, cyy) - c45;
          const double ryry = f_pp(cD, cxx, cxy, cyx, cyy) + c55 + hg4;
          rot_block_diag5(R, rxrx, rxry, ryrx, ryry, kd * hab, o);
          stage9(my, o, 3, 3, 6);
        }
        emit_staged<6, 6>(I66, st, inv, A.kc0v ? A.kc0v + A.kc0_k0 : nullptr, e * 576 + a * 144, act, F.csr_kc0,
                          b0 * 36, nb, first, lane);
      }
    }
  }
}

}  // namespace

size_t fused_smem_bytes() { return size_t(kFusedWarps) * kWarpSmemDoubles * sizeof(double); }
int fused_max_slots() { return kMaxSlots; }
int fused_record_stride(const EvalArgs& A) { return A.evec != nullptr ? kRecRot : kRecPlain; }

// rec: device scratch of ne * fused_record_stride doubles
cudaError_t launch_quad_fused(int kind, const FusedArgs& F, double* rec, cudaStream_t st, int64_t* launches) {
  if (F.nown <= 0 || F.A.ne <= 0) return cudaSuccess;
  const int stride = fused_record_stride(F.A);
  const unsigned g1 = unsigned((F.A.ne + 127) / 128);
  if (kind == PF3_QUAD4)
    quad_record_kernel<PF3_QUAD4><<<g1, 128, 0, st>>>(F.A, rec, stride);
  else
    quad_record_kernel<PF3_QUAD4R><<<g1, 128, 0, st>>>(F.A, rec, stride);
  ++*launches;
  cudaError_t e1 = cudaGetLastError();
  if (e1 != cudaSuccess) return e1;
  const size_t smem = fused_smem_bytes();
  const int64_t npairs = (F.nown + 1) / 2;
  const int64_t want = (npairs + kFusedWarps - 1) / kFusedWarps;
  const int64_t cap = 148 * 3 * 8;
  const unsigned grid = unsigned(want < cap ? (want < 1 ? 1 : want) : cap);
  static bool once = false;
  if (!once) {
    cudaFuncSetAttribute(quad_fused_kernel<PF3_QUAD4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(quad_fused_kernel<PF3_QUAD4R>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    once = true;
  }
  if (kind == PF3_QUAD4)
    quad_fused_kernel<PF3_QUAD4><<<grid, 32 * kFusedWarps, smem, st>>>(F, rec, stride);
  else
    quad_fused_kernel<PF3_QUAD4R><<<grid, 32 * kFusedWarps, smem, st>>>(F, rec, stride);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace pf3
