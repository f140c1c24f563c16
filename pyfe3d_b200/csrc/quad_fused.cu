// Fused, node-centric Quad4 / Quad4R evaluation + CSR assembly.
//
// One pass produces BOTH outputs the hot path owes:
//   * the reference's COO value arrays KC0v / KGv / Mv (update_KC0 quad4.pyx:1204, update_KG :1365,
//     update_KG_given_stress :2259, update_M :3083; Quad4R: quad4r.pyx:1145, :4681, :5574, :6398), and
//   * the CSR values scipy's coo_matrix(...).tocsr() would give (tests/test_quad4_static_point_load.py:80),
// without ever re-reading the COO arrays: DRAM traffic is the 15.1 kB/element lower bound of SURVEY §8(d).
//
// Work decomposition.  The unit of work is a NODE (= 6 CSR rows).  Its row block is the sum, over the
// elements incident to the node, of the element's 6x24 row slab for that node — and every (element,
// local node) slab belongs to exactly one node.  So a half-warp takes one node: 4 incident elements x 4
// node-pair blocks = 16 lanes, each lane evaluating ONE 6x6 block (a, b) of ONE element with the same
// instruction stream (no divergence).  The blocks are staged in shared memory, streamed out as contiguous
// COO slabs (1152 B each), and summed per CSR slot in a fixed order (deterministic, no atomics).
// FP64 work is ~4x redundant in the per-element set-up; the profile of the two-pass version showed the
// FP64 pipe at 11 % while DRAM sat at 50 %, so arithmetic is the resource to spend.
#include "shell.cuh"

namespace pf3 {

namespace {

constexpr double kGpF = 0.5773502691896257645092;
constexpr int kFusedWarps = 8;
constexpr int kLd = 33;                     // staging leading dimension (doubles)
constexpr int kStageDoubles = 36 * kLd;     // one 6x6 block per lane
constexpr int kMaxSlots = 16;               // column blocks per node row supported by the fused path
constexpr int kWarpSmemDoubles = kStageDoubles + (8 * kMaxSlots) / 8;

// vals[d*CNT + r]: this lane's block, rows d < NR, CNT masked columns per row.
template <int NR, int CNT>
__device__ __forceinline__ void emit_block(const double* vals, double* st, const signed char* inv, double* coo,
                                           int64_t slab_base, bool act, double* csr, int64_t csr_base, int nb,
                                           bool first_round, int lane) {
  const int h = lane >> 4, l16 = lane & 15;
#pragma unroll
  for (int t = 0; t < NR * CNT; ++t) st[t * kLd + lane] = vals[t];
  __syncwarp();
  if (coo != nullptr) {
#pragma unroll 1
    for (int idx = 0; idx < 8; ++idx) {
      const int64_t base = __shfl_sync(0xffffffffu, slab_base, idx * 4);
      const int on = __shfl_sync(0xffffffffu, act ? 1 : 0, idx * 4);
      if (on) {
#pragma unroll
        for (int p = lane; p < NR * 4 * CNT; p += 32) {
          const int d = p / (4 * CNT), rem = p - d * (4 * CNT);
          const int bb = rem / CNT, rr = rem - bb * CNT;
          coo[base + p] = st[(d * CNT + rr) * kLd + idx * 4 + bb];
        }
      }
    }
  }
  if (nb > 0 && csr != nullptr) {
    const int w = nb * CNT;
    const signed char* iv = inv + h * 4 * kMaxSlots;
    const double* sh = st + h * 16;
#pragma unroll 1
    for (int d = 0; d < NR; ++d) {
      double* out = csr + csr_base + int64_t(d) * w;
      for (int x = l16; x < w; x += 16) {
        const int s = x / CNT, rr = x - s * CNT;
        const double* row = sh + (d * CNT + rr) * kLd;
        double sum = 0.;
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          const int bs = iv[k2 * kMaxSlots + s];
          if (bs >= 0) sum += row[k2 * 4 + bs];
        }
        if (first_round) out[x] = sum; else out[x] += sum;
      }
    }
  }
  __syncwarp();
}

template <int KIND>
__global__ void __launch_bounds__(32 * kFusedWarps, 2) quad_fused_kernel(const FusedArgs F) {
  extern __shared__ double smem[];
  const EvalArgs& A = F.A;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* st = smem + warp * kWarpSmemDoubles;
  signed char* inv = reinterpret_cast<signed char*>(st + kStageDoubles);
  const int h = lane >> 4, l16 = lane & 15, k = l16 >> 2, b = l16 & 3;
  const int64_t npairs = (F.nown + 1) >> 1;
  const double xib = (b == 1 || b == 2) ? 1. : -1., etab = (b >= 2) ? 1. : -1.;

  for (int64_t np = int64_t(blockIdx.x) * kFusedWarps + warp; np < npairs; np += int64_t(gridDim.x) * kFusedWarps) {
    const int64_t n = 2 * np + h;
    const bool nodev = n < F.nown;
    const int64_t q0 = nodev ? F.inc_ptr[n] : 0;
    const int v = nodev ? int(F.inc_ptr[n + 1] - q0) : 0;
    const int64_t b0 = nodev ? F.brow_ptr[n] : 0;
    const int nb = nodev ? int(F.brow_ptr[n + 1] - b0) : 0;
    const int vmax = max(__shfl_sync(0xffffffffu, v, 0), __shfl_sync(0xffffffffu, v, 16));
    const int rounds = (vmax + 3) >> 2;

    for (int r = 0; r < rounds; ++r) {
      const int kk = 4 * r + k;
      const bool act = kk < v;
      const int64_t pair0 = act ? F.inc_pair0[q0 + kk] : 0;
      const int64_t e = pair0 >> 4;
      const int a = int(pair0 >> 2) & 3;
      const int myslot = act ? F.slot[pair0 + b] : -1;
      for (int i = lane; i < 8 * kMaxSlots; i += 32) inv[i] = -1;
      __syncwarp();
      if (act) inv[(h * 4 + k) * kMaxSlots + myslot] = (signed char)b;
      __syncwarp();
      const bool first = (r == 0);
      const double xia = (a == 1 || a == 2) ? 1. : -1., etaa = (a >= 2) ? 1. : -1.;
      const double sgn = ((a ^ b) & 1) ? -1. : 1.;

      // ---------------- element set-up (identical for the 4 lanes of an incidence)
      double ue[24];
      ShellGeom<4> g;
      shell_geom<4>(A, e, g, (A.what & PF3_KG) ? ue : nullptr);
      ShellCoef c;
      shell_coef<4>(A, e, g, c);
      const Mat3& R = g.R;
      const double dX10 = g.X[1] - g.X[0], dX23 = g.X[2] - g.X[3], dX30 = g.X[3] - g.X[0], dX21 = g.X[2] - g.X[1];
      const double dY10 = g.Y[1] - g.Y[0], dY23 = g.Y[2] - g.Y[3], dY30 = g.Y[3] - g.Y[0], dY21 = g.Y[2] - g.Y[1];
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
      double dJ[4], idJ[4], Wxa[4], Wya[4], Wxb[4], Wyb[4], Na[4], Nb[4];
#pragma unroll
      for (int gp = 0; gp < 4; ++gp) {
        const int ix = gp >> 1, ie = gp & 1;
        const double xg = ix ? kGpF : -kGpF, eg = ie ? kGpF : -kGpF;
        dJ[gp] = J11e[ie] * J22x[ix] - J12e[ie] * J21x[ix];
        idJ[gp] = 1. / dJ[gp];
        const double dxa = 0.25 * xia * (1. + etaa * eg), dea = 0.25 * etaa * (1. + xia * xg);
        const double dxb = 0.25 * xib * (1. + etab * eg), deb = 0.25 * etab * (1. + xib * xg);
        Wxa[gp] = J22x[ix] * dxa - J12e[ie] * dea;
        Wya[gp] = -J21x[ix] * dxa + J11e[ie] * dea;
        Wxb[gp] = J22x[ix] * dxb - J12e[ie] * deb;
        Wyb[gp] = -J21x[ix] * dxb + J11e[ie] * deb;
        Na[gp] = 0.25 * (1. + xia * xg) * (1. + etaa * eg);
        Nb[gp] = 0.25 * (1. + xib * xg) * (1. + etab * eg);
      }
      const double J11c = 0.25 * (dX10 + dX23), J12c = 0.25 * (dY10 + dY23);
      const double J21c = 0.25 * (dX30 + dX21), J22c = 0.25 * (dY30 + dY21);
      const double dJ0 = J11c * J22c - J12c * J21c, idJ0 = 1. / dJ0;
      const double w0 = 4. * dJ0;
      const double N0xa = (J22c * 0.25 * xia - J12c * 0.25 * etaa) * idJ0;
      const double N0ya = (-J21c * 0.25 * xia + J11c * 0.25 * etaa) * idJ0;
      const double N0xb = (J22c * 0.25 * xib - J12c * 0.25 * etab) * idJ0;
      const double N0yb = (-J21c * 0.25 * xib + J11c * 0.25 * etab) * idJ0;

      double kd = 1., hg[5] = {0., 0., 0., 0., 0.};
      if (KIND == PF3_QUAD4R) {
        double K6ROT = 100., hgf[5] = {1., 1., 1., 1., 1.};
        if (A.eparam != nullptr) {
          const double* ep = A.eparam + e * PF3_EPARAM_STRIDE;
          K6ROT = ep[0];
#pragma unroll
          for (int d = 0; d < 5; ++d) hgf[d] = ep[2 + d];
        }
        kd = 1e-6 * K6ROT * c.A[5];
        const double A11 = c.A[0], A12 = c.A[1], A16 = c.A[2], A22 = c.A[3], A26 = c.A[4], A66 = c.A[5];
        const double den = -A11 * A22 * A66 + A11 * A26 * A26 + A12 * A12 * A66 - 2 * A12 * A16 * A26 + A16 * A16 * A22;
        const double a11 = (-A22 * A66 + A26 * A26) / den, a22 = (-A11 * A66 + A16 * A16) / den;
        const double E1eq = 1. / (c.h * a11), E2eq = 1. / (c.h * a22);
        const double dd = 1.0 + 1.0 / g.area;
        const double Eu = hgf[0] * 0.1 * E1eq * c.h / dd, Ev = hgf[1] * 0.1 * E2eq * c.h / dd;
        const double Erx = hgf[3] * 0.1 * E2eq * c.h * c.h * c.h / dd, Ery = hgf[4] * 0.1 * E1eq * c.h * c.h * c.h / dd;
        const double Ew = hgf[2] * 0.5 * (Erx + Ery);
        // gamma_a = +-(j11 j22 + j12 j21)/4 with j = J0^-1 (quad4r.pyx:3116)
        const double gam = 0.25 * (J22c * J11c + J12c * J21c) * idJ0 * idJ0;
        const double wg2 = w0 * gam * gam;
        hg[0] = wg2 * Eu;
        hg[1] = wg2 * Ev;
        hg[2] = wg2 * Ew;
        hg[3] = wg2 * Erx;
        hg[4] = wg2 * Ery;
      }
      const bool thick = (KIND == PF3_QUAD4) && (c.h / sqrt(g.area) >= 1.);

      // gradient Gram of the pair over the 2x2 Gauss points, and the mixed N / N,x sums
      double gxx = 0., gxy = 0., gyx = 0., gyy = 0., pyab = 0., pxab = 0., pyba = 0., pxba = 0., hab = 0.;
#pragma unroll
      for (int gp = 0; gp < 4; ++gp) {
        const double vax = Wxa[gp] * idJ[gp], vay = Wya[gp] * idJ[gp];
        gxx += vax * Wxb[gp];
        gxy += vax * Wyb[gp];
        gyx += vay * Wxb[gp];
        gyy += vay * Wyb[gp];
        pyab += Wya[gp] * Nb[gp];
        pxab += Wxa[gp] * Nb[gp];
        pyba += Wyb[gp] * Na[gp];
        pxba += Wxb[gp] * Na[gp];
        hab += (Na[gp] * Nb[gp]) * dJ[gp];
      }

      // ---------------- KG : Ge_ab * z z^T on the translations
      if (A.what & (PF3_KG | PF3_KG_STRESS)) {
        double ge;
        if (A.what & PF3_KG_STRESS) {
          ge = A.Nxx * gxx + A.Nxy * (gxy + gyx) + A.Nyy * gyy;
        } else {
          ge = 0.;
#pragma unroll
          for (int gp = 0; gp < 4; ++gp) {
            const int ix = gp >> 1, ie = gp & 1;
            const double xg = ix ? kGpF : -kGpF, eg = ie ? kGpF : -kGpF;
            double exx = 0, eyy = 0, gxy_ = 0, kxx = 0, kyy = 0, kxy = 0;
#pragma unroll
            for (int cn = 0; cn < 4; ++cn) {
              const double xc = (cn == 1 || cn == 2) ? 1. : -1., ec = (cn >= 2) ? 1. : -1.;
              const double dxc = 0.25 * xc * (1. + ec * eg), dec = 0.25 * ec * (1. + xc * xg);
              const double nx = (J22x[ix] * dxc - J12e[ie] * dec) * idJ[gp];
              const double ny = (-J21x[ix] * dxc + J11e[ie] * dec) * idJ[gp];
              exx += nx * ue[6 * cn];
              eyy += ny * ue[6 * cn + 1];
              gxy_ += ny * ue[6 * cn] + nx * ue[6 * cn + 1];
              kxx += nx * ue[6 * cn + 4];
              kyy -= ny * ue[6 * cn + 3];
              kxy += ny * ue[6 * cn + 4] - nx * ue[6 * cn + 3];
            }
            const double Nxx = c.A[0] * exx + c.A[1] * eyy + c.A[2] * gxy_ + c.B[0] * kxx + c.B[1] * kyy + c.B[2] * kxy;
            const double Nyy = c.A[1] * exx + c.A[3] * eyy + c.A[4] * gxy_ + c.B[1] * kxx + c.B[3] * kyy + c.B[4] * kxy;
            const double Nxy = c.A[2] * exx + c.A[4] * eyy + c.A[5] * gxy_ + c.B[2] * kxx + c.B[4] * kyy + c.B[5] * kxy;
            const double nxa = Wxa[gp] * idJ[gp], nya = Wya[gp] * idJ[gp];
            ge += Wxb[gp] * (nxa * Nxx + nya * Nxy) + Wyb[gp] * (nxa * Nxy + nya * Nyy);
          }
        }
        double vals[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) vals[i * 3 + j] = (R.a[i][2] * R.a[j][2]) * ge;
        emit_block<3, 3>(vals, st, inv, A.kgv ? A.kgv + A.kg_k0 : nullptr, e * 144 + a * 36, act, F.csr_kg,
                         b0 * 9, nb, first, lane);
      }

      // ---------------- M : H_ab * (T6 m_l T6^T)
      if (A.what & PF3_M) {
        double H;
        if (A.mtype == 0) {
          H = hab;
        } else if (A.mtype == 1) {
          H = 0.0625 * g.area;
        } else {
          // Gauss-Lobatto: detJ at node a on the diagonal, zero elsewhere (quad4.pyx:8873)
          const double J11n = 0.25 * ((1. - etaa) * dX10 + (1. + etaa) * dX23), J12n = 0.25 * ((1. - etaa) * dY10 + (1. + etaa) * dY23);
          const double J21n = 0.25 * ((1. - xia) * dX30 + (1. + xia) * dX21), J22n = 0.25 * ((1. - xia) * dY30 + (1. + xia) * dY21);
          H = (a == b) ? (J11n * J22n - J12n * J21n) : 0.;
        }
        NodalInertia Mi;
        nodal_inertia(R, c.rho0, c.rho1, c.rho2, Mi);
        double* coo = A.mv ? A.mv + A.m_k0 : nullptr;
        if (A.mtype != 2) {
          double vals[30];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            vals[i * 5 + 0] = H * Mi.tt[i][0];
            vals[i * 5 + 1] = H * Mi.tt[i][1];
            vals[i * 5 + 2] = H * Mi.tt[i][2];
            vals[i * 5 + 3] = H * Mi.tr[i][(i == 0) ? 1 : 0];
            vals[i * 5 + 4] = H * Mi.tr[i][(i == 2) ? 1 : 2];
            vals[(3 + i) * 5 + 0] = H * Mi.tr[(i == 0) ? 1 : 0][i];
            vals[(3 + i) * 5 + 1] = H * Mi.tr[(i == 2) ? 1 : 2][i];
            vals[(3 + i) * 5 + 2] = H * Mi.rr[i][0];
            vals[(3 + i) * 5 + 3] = H * Mi.rr[i][1];
            vals[(3 + i) * 5 + 4] = H * Mi.rr[i][2];
          }
          emit_block<6, 5>(vals, st, inv, coo, e * 480 + a * 120, act, F.csr_m, b0 * 30, nb, first, lane);
        } else {
          double vals[18];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              vals[i * 3 + j] = H * Mi.tt[i][j];
              vals[(3 + i) * 3 + j] = H * Mi.rr[i][j];
            }
          emit_block<6, 3>(vals, st, inv, coo, e * 480 + a * 72, act, F.csr_m, b0 * 18, nb, first, lane);
        }
      }

      // ---------------- KC0 : the 6x6 block (a, b)
      if (A.what & PF3_KC0) {
        // constitutive Gram: 2x2 Gauss (Quad4) or centre point with weight 4 detJ0 (Quad4R)
        double cxx = gxx, cxy = gxy, cyx = gyx, cyy = gyy;
        if (KIND == PF3_QUAD4R) {
          const double wa = w0 * N0xa, wb = w0 * N0ya;
          cxx = wa * N0xb;
          cxy = wa * N0yb;
          cyx = wb * N0xb;
          cyy = wb * N0yb;
        }
        const double tSa = w0 * (c.E44 * N0ya + c.E45 * N0xa), sSa = w0 * (c.E45 * N0ya + c.E55 * N0xa);
        const double tSb = w0 * (c.E44 * N0yb + c.E45 * N0xb), sSb = w0 * (c.E45 * N0yb + c.E55 * N0xb);
        const double c44 = w0 * c.E44 * 0.0625, c45 = w0 * c.E45 * 0.0625, c55 = w0 * c.E55 * 0.0625;
        double vals[36];
        double o[3][3];
        {
          double uu = f_pp(c.A, cxx, cxy, cyx, cyy) + 0.25 * kd * gyy;
          double uv = f_pq(c.A, cxx, cxy, cyx, cyy) - 0.25 * kd * gyx;
          double vu = f_qp(c.A, cxx, cxy, cyx, cyy) - 0.25 * kd * gxy;
          double vv = f_qq(c.A, cxx, cxy, cyx, cyy) + 0.25 * kd * gxx;
          double ww = thick ? (c.E44 * gyy + c.E45 * (gxy + gyx) + c.E55 * gxx) : (tSa * N0yb + sSa * N0xb);
          if (KIND == PF3_QUAD4R) {
            uu += sgn * hg[0];
            vv += sgn * hg[1];
            ww += sgn * hg[2];
          }
          rot_block_diag5(R, uu, uv, vu, vv, ww, o);
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) vals[i * 6 + j] = o[i][j];
        }
        rot_block_8(R, -f_pq(c.B, cxx, cxy, cyx, cyy), f_pp(c.B, cxx, cxy, cyx, cyy), 0.5 * kd * pyab,
                    -f_qq(c.B, cxx, cxy, cyx, cyy), f_qp(c.B, cxx, cxy, cyx, cyy), -0.5 * kd * pxab, -0.25 * tSa,
                    0.25 * sSa, o);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) vals[i * 6 + 3 + j] = o[i][j];
        rot_block_8(R, -f_qp(c.B, cxx, cxy, cyx, cyy), -f_qq(c.B, cxx, cxy, cyx, cyy), -0.25 * tSb,
                    f_pp(c.B, cxx, cxy, cyx, cyy), f_pq(c.B, cxx, cxy, cyx, cyy), 0.25 * sSb, 0.5 * kd * pyba,
                    -0.5 * kd * pxba, o);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) vals[(3 + i) * 6 + j] = o[i][j];
        {
          double rxrx = f_qq(c.D, cxx, cxy, cyx, cyy) + c44;
          double rxry = -f_qp(c.D, cxx, cxy, cyx, cyy) - c45;
          double ryrx = -f_pq(c.D, cxx, cxy, cyx, cyy) - c45;
          double ryry = f_pp(c.D, cxx, cxy, cyx, cyy) + c55;
          if (KIND == PF3_QUAD4R) {
            rxrx += sgn * hg[3];
            ryry += sgn * hg[4];
          }
          rot_block_diag5(R, rxrx, rxry, ryrx, ryry, kd * hab, o);
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) vals[(3 + i) * 6 + 3 + j] = o[i][j];
        }
        emit_block<6, 6>(vals, st, inv, A.kc0v ? A.kc0v + A.kc0_k0 : nullptr, e * 576 + a * 144, act, F.csr_kc0,
                         b0 * 36, nb, first, lane);
      }
    }
  }
}

}  // namespace

size_t fused_smem_bytes() { return size_t(kFusedWarps) * kWarpSmemDoubles * sizeof(double); }
int fused_max_slots() { return kMaxSlots; }

cudaError_t launch_quad_fused(int kind, const FusedArgs& F, cudaStream_t st) {
  if (F.nown <= 0) return cudaSuccess;
  const size_t smem = fused_smem_bytes();
  const int64_t npairs = (F.nown + 1) / 2;
  const int64_t want = (npairs + kFusedWarps - 1) / kFusedWarps;
  const unsigned grid = unsigned(want < 148 * 16 ? (want < 1 ? 1 : want) : 148 * 16);
  static bool once = false;
  if (!once) {
    cudaFuncSetAttribute(quad_fused_kernel<PF3_QUAD4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(quad_fused_kernel<PF3_QUAD4R>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    once = true;
  }
  if (kind == PF3_QUAD4)
    quad_fused_kernel<PF3_QUAD4><<<grid, 32 * kFusedWarps, smem, st>>>(F);
  else
    quad_fused_kernel<PF3_QUAD4R><<<grid, 32 * kFusedWarps, smem, st>>>(F);
  return cudaGetLastError();
}

}  // namespace pf3
