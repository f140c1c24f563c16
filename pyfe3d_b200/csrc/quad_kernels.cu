// Quad4 and Quad4R: KC0 / KG / KG_given_stress / M / fint, one element per thread.
//
// Replaces (reference, /root/reference/pyfe3d):
//   Quad4 : update_rotation_matrix quad4.pyx:491, update_probe_xe :682, update_probe_ue :627,
//           _update_probe_KC0ve :755, update_KC0 :1204, update_fint :1316, update_KG :1365,
//           update_KG_given_stress :2259, update_M :3083
//   Quad4R: quad4r.pyx:286, :476, :421, update_KC0 :1145, update_fint :4620, update_KG :4681,
//           update_KG_given_stress :5574, update_M :6398
#include "shell.cuh"

namespace pf3 {

namespace {

constexpr double kGp = 0.5773502691896257645092;  // quad4.pyx:929
// shape function values at the 2x2 Gauss points, g = 2*ixi + ieta (xi outer, quad4.pyx:932-935)
constexpr double kNhi = 0.25 * (1. + kGp) * (1. + kGp);
constexpr double kNmid = 0.25 * (1. - kGp * kGp);
constexpr double kNlo = 0.25 * (1. - kGp) * (1. - kGp);
// node a has (xi_a, eta_a) = (-1,-1),(1,-1),(1,1),(-1,1)
__device__ constexpr double kNgp[4][4] = {
    // g=0: xi=-p, eta=-p
    {kNhi, kNmid, kNlo, kNmid},
    // g=1: xi=-p, eta=+p
    {kNmid, kNlo, kNmid, kNhi},
    // g=2: xi=+p, eta=-p
    {kNmid, kNhi, kNmid, kNlo},
    // g=3: xi=+p, eta=+p
    {kNlo, kNmid, kNhi, kNmid}};

struct QuadInteg {
  double Wx[4][4], Wy[4][4];  // detJ * N_a,x , detJ * N_a,y at Gauss point g
  double idJ[4], dJ[4];
  double N0x[4], N0y[4];      // gradients at the centre
  double w0;                  // 4 * detJ(0,0)   (wij = 4, quad4.pyx:1076)
  double j0prod;              // j11*j22 + j12*j21 at the centre (hourglass vector, quad4r.pyx:3116)
};

__device__ __forceinline__ void jac_at(const double* X, const double* Y, double xi, double eta, double* wx,
                                       double* wy, double& dJ) {
  const double ome = 1. - eta, ope = 1. + eta, omx = 1. - xi, opx = 1. + xi;
  const double J11 = 0.25 * (ome * (X[1] - X[0]) + ope * (X[2] - X[3]));
  const double J12 = 0.25 * (ome * (Y[1] - Y[0]) + ope * (Y[2] - Y[3]));
  const double J21 = 0.25 * (omx * (X[3] - X[0]) + opx * (X[2] - X[1]));
  const double J22 = 0.25 * (omx * (Y[3] - Y[0]) + opx * (Y[2] - Y[1]));
  dJ = J11 * J22 - J12 * J21;
  const double dxi[4] = {-0.25 * ome, 0.25 * ome, 0.25 * ope, -0.25 * ope};
  const double det[4] = {-0.25 * omx, -0.25 * opx, 0.25 * opx, 0.25 * omx};
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    wx[a] = J22 * dxi[a] - J12 * det[a];   // detJ * (j11 dxi + j12 deta)
    wy[a] = -J21 * dxi[a] + J11 * det[a];  // detJ * (j21 dxi + j22 deta)
  }
}

__device__ __forceinline__ void quad_integ(const ShellGeom<4>& g, QuadInteg& q) {
#pragma unroll
  for (int gp = 0; gp < 4; ++gp) {
    const double xi = (gp & 2) ? kGp : -kGp;
    const double eta = (gp & 1) ? kGp : -kGp;
    jac_at(g.X, g.Y, xi, eta, q.Wx[gp], q.Wy[gp], q.dJ[gp]);
    q.idJ[gp] = 1. / q.dJ[gp];
  }
  double wx[4], wy[4], dJ0;
  jac_at(g.X, g.Y, 0., 0., wx, wy, dJ0);
  const double i0 = 1. / dJ0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    q.N0x[a] = wx[a] * i0;
    q.N0y[a] = wy[a] * i0;
  }
  q.w0 = 4. * dJ0;
  // j11 = 4*N0x[2]... recover from nodes 2,3: N_3,x = (j11+j12)/4, N_2,x = (j11-j12)/4
  const double j11 = 2. * (q.N0x[2] + q.N0x[1]), j12 = 2. * (q.N0x[2] - q.N0x[1]);
  const double j21 = 2. * (q.N0y[2] + q.N0y[1]), j22 = 2. * (q.N0y[2] - q.N0y[1]);
  q.j0prod = j11 * j22 + j12 * j21;
}

struct PairGram {
  double xx, xy, yx, yy;
};
// 2x2-Gauss gradient Gram of node pair (a,b):  sum_g detJ N_a,p N_b,q
__device__ __forceinline__ PairGram gram_full(const QuadInteg& q, int a, int b) {
  PairGram G = {0., 0., 0., 0.};
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const double vax = q.Wx[g][a] * q.idJ[g], vay = q.Wy[g][a] * q.idJ[g];
    G.xx += vax * q.Wx[g][b];
    G.xy += vax * q.Wy[g][b];
    G.yx += vay * q.Wx[g][b];
    G.yy += vay * q.Wy[g][b];
  }
  return G;
}
__device__ __forceinline__ PairGram gram_centre(const QuadInteg& q, int a, int b) {
  PairGram G;
  const double wa = q.w0 * q.N0x[a], wb = q.w0 * q.N0y[a];
  G.xx = wa * q.N0x[b];
  G.xy = wa * q.N0y[b];
  G.yx = wb * q.N0x[b];
  G.yy = wb * q.N0y[b];
  return G;
}

struct QuadExtra {       // Quad4R only
  double kd;             // 1e-6 * K6ROT * A66 (drilling penalty, quad4r.pyx:3306 ff.)
  double hg[5];          // w0 * E_d * g^2 for d = u v w rx ry (quad4r.pyx:3088-3092)
};

// ONLY = 0: any subset of the outputs.  ONLY = PF3_FINT: the instantiation for update_fint / update_probe_finte / state
// calls, which contains no matrix code (248 registers and no spills against 254 + 144 bytes of spills; forcing three
// CTAs per SM costs 1.1 kB of spills, so occupancy stays at 8 warps per SM).  A four-lanes-per-element variant (lane =
// Gauss point during the integration, = node afterwards, 168 registers, no staging) was measured and rejected: every
// lane has to form the frame and the laminate coefficients itself, and update_fint of 4 M Quad4 took 1.78 ms against
// 1.37 ms with this kernel.
#ifndef PF3_FINT_CTAS
#define PF3_FINT_CTAS 2
#endif
template <int KIND, int ONLY>
__global__ void __launch_bounds__(kThreads, ONLY ? PF3_FINT_CTAS : 1) quad_eval_kernel(const EvalArgs A) {
  const int what = ONLY ? (A.what & ONLY) : A.what;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t e0 = (int64_t(blockIdx.x) * kWarpsPerCta + warp) * 32;
  if (e0 >= A.ne) return;
  const int nvalid = int(min(int64_t(32), A.ne - e0));
  const int64_t e = e0 + min(lane, nvalid - 1);
  if (A.state == nullptr) conn_prefetch(A.conn, 4, e0, A.ne, lane);
  double* stage = smem + warp * 32 * kStageLd;
  double* my = stage + lane * kStageLd;

  const bool need_u = (what & (PF3_KG | PF3_FINT)) != 0 || A.state_out != nullptr;
  double ue[24];
  ShellGeom<4> g;
  shell_geom<4, true>(A, e, g, need_u && (A.u != nullptr || A.state != nullptr) ? ue : nullptr);
  if (A.state_out != nullptr) {
    if (lane < nvalid) store_state<4>(A, e, g, (A.u != nullptr || A.state != nullptr) ? ue : nullptr);
    if (what == 0) return;
  }
  ShellCoef c;
  shell_coef<4>(A, e, g, c);
  QuadInteg q;
  quad_integ(g, q);
  const Mat3& R = g.R;

  QuadExtra X;
  if (KIND == PF3_QUAD4R) {
    double K6ROT = 100., hgf[5] = {1., 1., 1., 1., 1.};
    if (A.eparam != nullptr) {
      const double* ep = A.eparam + e * PF3_EPARAM_STRIDE;
      K6ROT = ep[0];
#pragma unroll
      for (int d = 0; d < 5; ++d) hgf[d] = ep[2 + d];
    }
    X.kd = 1e-6 * K6ROT * c.A[5];
    const double A11 = c.A[0], A12 = c.A[1], A16 = c.A[2], A22 = c.A[3], A26 = c.A[4], A66 = c.A[5];
    const double den = -A11 * A22 * A66 + A11 * A26 * A26 + A12 * A12 * A66 - 2 * A12 * A16 * A26 + A16 * A16 * A22;
    const double a11 = (-A22 * A66 + A26 * A26) / den;
    const double a22 = (-A11 * A66 + A16 * A16) / den;
    const double E1eq = 1. / (c.h * a11), E2eq = 1. / (c.h * a22);
    const double dd = 1.0 + 1.0 / g.area;
    const double Eu = hgf[0] * 0.1 * E1eq * c.h / dd;
    const double Ev = hgf[1] * 0.1 * E2eq * c.h / dd;
    const double Erx = hgf[3] * 0.1 * E2eq * c.h * c.h * c.h / dd;
    const double Ery = hgf[4] * 0.1 * E1eq * c.h * c.h * c.h / dd;
    const double Ew = hgf[2] * 0.5 * (Erx + Ery);
    const double gam = 0.25 * q.j0prod;
    const double wg2 = q.w0 * gam * gam;
    X.hg[0] = wg2 * Eu;
    X.hg[1] = wg2 * Ev;
    X.hg[2] = wg2 * Ew;
    X.hg[3] = wg2 * Erx;
    X.hg[4] = wg2 * Ery;
  } else {
    X.kd = 1.;  // Quad4 ignores K6ROT: drilling coefficient is 1.0 (quad4.pyx:1069-1073)
  }
  // Quad4 thick/thin switch (quad4.pyx:1032,1127); Quad4R is always "thin"-style reduced shear
  const bool thick = (KIND == PF3_QUAD4) && (c.h / sqrt(g.area) >= 1.);
  // centre shear vectors: t_a = w0 (E44 N0y + E45 N0x), s_a = w0 (E45 N0y + E55 N0x)
  double tS[4], sS[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    tS[a] = q.w0 * (c.E44 * q.N0y[a] + c.E45 * q.N0x[a]);
    sS[a] = q.w0 * (c.E45 * q.N0y[a] + c.E55 * q.N0x[a]);
  }
  const double c44 = q.w0 * c.E44 * 0.0625, c45 = q.w0 * c.E45 * 0.0625, c55 = q.w0 * c.E55 * 0.0625;

  // ------------------------------------------------------------------ KG
  if (what & (PF3_KG | PF3_KG_STRESS)) {
    double Ge[4][4];
    if (what & PF3_KG_STRESS) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a; b < 4; ++b) {
          PairGram G = gram_full(q, a, b);
          Ge[a][b] = A.Nxx * G.xx + A.Nxy * (G.xy + G.yx) + A.Nyy * G.yy;
          Ge[b][a] = Ge[a][b];
        }
    } else {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) Ge[a][b] = 0.;
#pragma unroll
      for (int gp = 0; gp < 4; ++gp) {
        double nx[4], ny[4];
        double exx = 0, eyy = 0, gxy = 0, kxx = 0, kyy = 0, kxy = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          nx[a] = q.Wx[gp][a] * q.idJ[gp];
          ny[a] = q.Wy[gp][a] * q.idJ[gp];
          exx += nx[a] * ue[6 * a];
          eyy += ny[a] * ue[6 * a + 1];
          gxy += ny[a] * ue[6 * a] + nx[a] * ue[6 * a + 1];
          kxx += nx[a] * ue[6 * a + 4];
          kyy -= ny[a] * ue[6 * a + 3];
          kxy += ny[a] * ue[6 * a + 4] - nx[a] * ue[6 * a + 3];
        }
        // membrane force resultants (quad4.pyx:1965-1967)
        const double Nxx = c.A[0] * exx + c.A[1] * eyy + c.A[2] * gxy + c.B[0] * kxx + c.B[1] * kyy + c.B[2] * kxy;
        const double Nyy = c.A[1] * exx + c.A[3] * eyy + c.A[4] * gxy + c.B[1] * kxx + c.B[3] * kyy + c.B[4] * kxy;
        const double Nxy = c.A[2] * exx + c.A[4] * eyy + c.A[5] * gxy + c.B[2] * kxx + c.B[4] * kyy + c.B[5] * kxy;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const double px = nx[a] * Nxx + ny[a] * Nxy, py = nx[a] * Nxy + ny[a] * Nyy;
#pragma unroll
          for (int b = a; b < 4; ++b) Ge[a][b] += q.Wx[gp][b] * px + q.Wy[gp][b] * py;
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < a; ++b) Ge[a][b] = Ge[b][a];
    }
    double zz[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) zz[i][j] = R.a[i][2] * R.a[j][2];
    double* out = A.kgv + A.kg_k0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int j = 0; j < 3; ++j) my[i * 12 + b * 3 + j] = zz[i][j] * Ge[a][b];
      flush_chunk<36>(stage, out, e0, nvalid, 144, a * 36, A.acc_kg != 0, lane);
    }
  }

  // ------------------------------------------------------------------ M
  if (what & PF3_M) {
    double H[4][4];
    if (A.mtype == 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a; b < 4; ++b) {
          double h = 0.;
#pragma unroll
          for (int gp = 0; gp < 4; ++gp) h += (kNgp[gp][a] * kNgp[gp][b]) * q.dJ[gp];
          H[a][b] = h;
          H[b][a] = h;
        }
    } else if (A.mtype == 1) {
      const double v = 0.0625 * g.area;  // quad4.pyx:3125
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) H[a][b] = v;
    } else {
      // Gauss-Lobatto points sit on the nodes (quad4.pyx:8873): H diagonal, detJ at node a
      const double xs[4] = {-1., 1., 1., -1.}, es[4] = {-1., -1., 1., 1.};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        double wx[4], wy[4], dJn;
        jac_at(g.X, g.Y, xs[a], es[a], wx, wy, dJn);
#pragma unroll
        for (int b = 0; b < 4; ++b) H[a][b] = (a == b) ? dJn : 0.;
      }
    }
    NodalInertia Mi;
    nodal_inertia(R, c.rho0, c.rho1, c.rho2, Mi);
    double* out = A.mv + A.m_k0;
    if (A.mtype != 2) {
      // mask: row i (translation) -> cols 0,1,2 and rotations except 3+i; row 3+i -> translations
      // except i, then 3,4,5 (quad4.pyx:3158 ff.)
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            double* d = my + i * 20 + b * 5;
            d[0] = H[a][b] * Mi.tt[i][0];
            d[1] = H[a][b] * Mi.tt[i][1];
            d[2] = H[a][b] * Mi.tt[i][2];
            d[3] = H[a][b] * Mi.tr[i][(i == 0) ? 1 : 0];
            d[4] = H[a][b] * Mi.tr[i][(i == 2) ? 1 : 2];
          }
        flush_chunk<60>(stage, out, e0, nvalid, 480, a * 120, A.acc_m != 0, lane);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            double* d = my + i * 20 + b * 5;
            // rt = tr^T
            d[0] = H[a][b] * Mi.tr[(i == 0) ? 1 : 0][i];
            d[1] = H[a][b] * Mi.tr[(i == 2) ? 1 : 2][i];
            d[2] = H[a][b] * Mi.rr[i][0];
            d[3] = H[a][b] * Mi.rr[i][1];
            d[4] = H[a][b] * Mi.rr[i][2];
          }
        flush_chunk<60>(stage, out, e0, nvalid, 480, a * 120 + 60, A.acc_m != 0, lane);
      }
    } else {
      // lumped: translation-translation and rotation-rotation only, 288 of 480 entries written
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              my[i * 12 + b * 3 + j] = H[a][b] * Mi.tt[i][j];
              my[36 + i * 12 + b * 3 + j] = H[a][b] * Mi.rr[i][j];
            }
        flush_chunk<72>(stage, out, e0, nvalid, 480, a * 72, A.acc_m != 0, lane);
      }
    }
  }

  // ------------------------------------------------------------------ KC0
  if (what & PF3_KC0) {
    double* out = A.kc0v + A.kc0_k0;
#pragma unroll 1
    for (int a = 0; a < 4; ++a) {
      // ---- rows u v w of node a
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const PairGram Gd = gram_full(q, a, b);                                      // drilling / Quad4 ABD
        const PairGram G = (KIND == PF3_QUAD4R) ? gram_centre(q, a, b) : Gd;           // ABD part
        const double sgn = ((a ^ b) & 1) ? -1. : 1.;
        double uu = f_pp(c.A, G.xx, G.xy, G.yx, G.yy) + 0.25 * X.kd * Gd.yy;
        double uv = f_pq(c.A, G.xx, G.xy, G.yx, G.yy) - 0.25 * X.kd * Gd.yx;
        double vu = f_qp(c.A, G.xx, G.xy, G.yx, G.yy) - 0.25 * X.kd * Gd.xy;
        double vv = f_qq(c.A, G.xx, G.xy, G.yx, G.yy) + 0.25 * X.kd * Gd.xx;
        double ww;
        if (thick) {
          ww = c.E44 * Gd.yy + c.E45 * (Gd.xy + Gd.yx) + c.E55 * Gd.xx;
        } else {
          ww = tS[a] * q.N0y[b] + sS[a] * q.N0x[b];
        }
        if (KIND == PF3_QUAD4R) {
          uu += sgn * X.hg[0];
          vv += sgn * X.hg[1];
          ww += sgn * X.hg[2];
        }
        double o[3][3];
        rot_block_diag5(R, uu, uv, vu, vv, ww, o);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) my[i * 24 + b * 6 + j] = o[i][j];
        // translation(a) x rotation(b)
        double pyab = 0., pxab = 0.;
#pragma unroll
        for (int gp = 0; gp < 4; ++gp) {
          pyab += q.Wy[gp][a] * kNgp[gp][b];
          pxab += q.Wx[gp][a] * kNgp[gp][b];
        }
        const double u_rx = -f_pq(c.B, G.xx, G.xy, G.yx, G.yy);
        const double u_ry = f_pp(c.B, G.xx, G.xy, G.yx, G.yy);
        const double v_rx = -f_qq(c.B, G.xx, G.xy, G.yx, G.yy);
        const double v_ry = f_qp(c.B, G.xx, G.xy, G.yx, G.yy);
        rot_block_8(R, u_rx, u_ry, 0.5 * X.kd * pyab, v_rx, v_ry, -0.5 * X.kd * pxab, -0.25 * tS[a],
                    0.25 * sS[a], o);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) my[i * 24 + b * 6 + 3 + j] = o[i][j];
      }
      flush_chunk<72>(stage, out, e0, nvalid, 576, a * 144, A.acc_kc0 != 0, lane);
      // ---- rows rx ry rz of node a
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const PairGram Gd = gram_full(q, a, b);
        const PairGram G = (KIND == PF3_QUAD4R) ? gram_centre(q, a, b) : Gd;
        const double sgn = ((a ^ b) & 1) ? -1. : 1.;
        double pyba = 0., pxba = 0., hab = 0.;
#pragma unroll
        for (int gp = 0; gp < 4; ++gp) {
          pyba += q.Wy[gp][b] * kNgp[gp][a];
          pxba += q.Wx[gp][b] * kNgp[gp][a];
          hab += (kNgp[gp][a] * kNgp[gp][b]) * q.dJ[gp];
        }
        const double rx_u = -f_qp(c.B, G.xx, G.xy, G.yx, G.yy);
        const double rx_v = -f_qq(c.B, G.xx, G.xy, G.yx, G.yy);
        const double ry_u = f_pp(c.B, G.xx, G.xy, G.yx, G.yy);
        const double ry_v = f_pq(c.B, G.xx, G.xy, G.yx, G.yy);
        double o[3][3];
        rot_block_8(R, rx_u, rx_v, -0.25 * tS[b], ry_u, ry_v, 0.25 * sS[b], 0.5 * X.kd * pyba,
                    -0.5 * X.kd * pxba, o);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) my[i * 24 + b * 6 + j] = o[i][j];
        double rxrx = f_qq(c.D, G.xx, G.xy, G.yx, G.yy) + c44;
        double rxry = -f_qp(c.D, G.xx, G.xy, G.yx, G.yy) - c45;
        double ryrx = -f_pq(c.D, G.xx, G.xy, G.yx, G.yy) - c45;
        double ryry = f_pp(c.D, G.xx, G.xy, G.yx, G.yy) + c55;
        if (KIND == PF3_QUAD4R) {
          rxrx += sgn * X.hg[3];
          ryry += sgn * X.hg[4];
        }
        rot_block_diag5(R, rxrx, rxry, ryrx, ryry, X.kd * hab, o);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) my[i * 24 + b * 6 + 3 + j] = o[i][j];
      }
      flush_chunk<72>(stage, out, e0, nvalid, 576, a * 144 + 72, A.acc_kc0 != 0, lane);
    }
  }

  // ------------------------------------------------------------------ fint / finte
  if (what & PF3_FINT) {
    double f[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) f[i] = 0.;
    const double C6[6][6] = {{c.A[0], c.A[1], c.A[2], c.B[0], c.B[1], c.B[2]},
                             {c.A[1], c.A[3], c.A[4], c.B[1], c.B[3], c.B[4]},
                             {c.A[2], c.A[4], c.A[5], c.B[2], c.B[4], c.B[5]},
                             {c.B[0], c.B[1], c.B[2], c.D[0], c.D[1], c.D[2]},
                             {c.B[1], c.B[3], c.B[4], c.D[1], c.D[3], c.D[4]},
                             {c.B[2], c.B[4], c.B[5], c.D[2], c.D[4], c.D[5]}};
    // constitutive part: 2x2 Gauss for Quad4, centre point (weight w0) for Quad4R
    auto abd_point = [&](const double* nx, const double* ny, const double* wx, const double* wy) {
      double eps[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        eps[0] += nx[a] * ue[6 * a];
        eps[1] += ny[a] * ue[6 * a + 1];
        eps[2] += ny[a] * ue[6 * a] + nx[a] * ue[6 * a + 1];
        eps[3] += nx[a] * ue[6 * a + 4];
        eps[4] -= ny[a] * ue[6 * a + 3];
        eps[5] += ny[a] * ue[6 * a + 4] - nx[a] * ue[6 * a + 3];
      }
      double s[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        s[i] = 0.;
#pragma unroll
        for (int j = 0; j < 6; ++j) s[i] += C6[i][j] * eps[j];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        f[6 * a + 0] += wx[a] * s[0] + wy[a] * s[2];
        f[6 * a + 1] += wy[a] * s[1] + wx[a] * s[2];
        f[6 * a + 3] -= wy[a] * s[4] + wx[a] * s[5];
        f[6 * a + 4] += wx[a] * s[3] + wy[a] * s[5];
      }
    };
#pragma unroll
    for (int gp = 0; gp < 4; ++gp) {
      double nx[4], ny[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        nx[a] = q.Wx[gp][a] * q.idJ[gp];
        ny[a] = q.Wy[gp][a] * q.idJ[gp];
      }
      if (KIND == PF3_QUAD4) abd_point(nx, ny, q.Wx[gp], q.Wy[gp]);
      double th = 0.;
#pragma unroll
      for (int a = 0; a < 4; ++a) th += 0.5 * ny[a] * ue[6 * a] - 0.5 * nx[a] * ue[6 * a + 1] + kNgp[gp][a] * ue[6 * a + 5];
      th *= X.kd;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        f[6 * a + 0] += 0.5 * q.Wy[gp][a] * th;
        f[6 * a + 1] -= 0.5 * q.Wx[gp][a] * th;
        f[6 * a + 5] += q.dJ[gp] * kNgp[gp][a] * th;
      }
      if (thick) {
        double gyz = 0., gxz = 0.;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          gyz += ny[a] * ue[6 * a + 2];
          gxz += nx[a] * ue[6 * a + 2];
        }
        const double ty = c.E44 * gyz + c.E45 * gxz, tx = c.E45 * gyz + c.E55 * gxz;
#pragma unroll
        for (int a = 0; a < 4; ++a) f[6 * a + 2] += q.Wy[gp][a] * ty + q.Wx[gp][a] * tx;
      }
    }
    if (KIND == PF3_QUAD4R) {
      double wx[4], wy[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        wx[a] = q.w0 * q.N0x[a];
        wy[a] = q.w0 * q.N0y[a];
      }
      abd_point(q.N0x, q.N0y, wx, wy);
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        const double s = ue[d] - ue[6 + d] + ue[12 + d] - ue[18 + d];
        f[d] += X.hg[d] * s;
        f[6 + d] -= X.hg[d] * s;
        f[12 + d] += X.hg[d] * s;
        f[18 + d] -= X.hg[d] * s;
      }
    }
    {
      // centre transverse shear
      double gyzg = 0., gxzg = 0., ry = 0., rx = 0.;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        gyzg += q.N0y[a] * ue[6 * a + 2];
        gxzg += q.N0x[a] * ue[6 * a + 2];
        rx += ue[6 * a + 3];
        ry += ue[6 * a + 4];
      }
      const double gyz = gyzg - 0.25 * rx, gxz = gxzg + 0.25 * ry;
      const double ty = q.w0 * (c.E44 * gyz + c.E45 * gxz), tx = q.w0 * (c.E45 * gyz + c.E55 * gxz);
      double tyw = ty, txw = tx;
      if (thick) {  // minus the gradient-gradient part already integrated at 2x2 (quad4.pyx:1152-1171)
        tyw -= q.w0 * (c.E44 * gyzg + c.E45 * gxzg);
        txw -= q.w0 * (c.E45 * gyzg + c.E55 * gxzg);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        f[6 * a + 2] += q.N0y[a] * tyw + q.N0x[a] * txw;
        f[6 * a + 3] -= 0.25 * ty;
        f[6 * a + 4] += 0.25 * tx;
      }
    }
    if (A.finte != nullptr) {
#pragma unroll
      for (int i = 0; i < 24; ++i) my[i] = f[i];
      flush_chunk<24>(stage, A.finte, e0, nvalid, 24, 0, false, lane);
    }
    if (A.fe != nullptr) {
#pragma unroll
      for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int i = 0; i < 3; ++i)
          my[3 * t + i] = R.a[i][0] * f[3 * t] + R.a[i][1] * f[3 * t + 1] + R.a[i][2] * f[3 * t + 2];
      flush_chunk<24>(stage, A.fe, e0, nvalid, 24, 0, false, lane);
    }
  }
}

// Piston-theory aerodynamic matrices (update_KA_beta quad4.pyx:9491, update_KA_gamma :10312, update_CA :11115;
// Quad4R: quad4r.pyx:12789, :13605, :14403).  One element per thread.  All three are a 4x4 scalar matrix over the
// node pairs times z z^T on the translations (z = third column of R), 2x2 Gauss with wij = 1:
//   KA_beta_ab = - sum_gp N_a detJ (N_b,x r11 + N_b,y r21),  KA_gamma_ab = sum_gp N_a N_b detJ,  CA = -KA_gamma.
// Same 144-entry block layout as KG, staged per warp and flushed as contiguous runs.
__global__ void __launch_bounds__(kThreads) quad_aero_kernel(const EvalArgs A, const AeroOut O) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t e0 = (int64_t(blockIdx.x) * kWarpsPerCta + warp) * 32;
  if (e0 >= A.ne) return;
  const int nvalid = int(min(int64_t(32), A.ne - e0));
  const int64_t e = e0 + min(lane, nvalid - 1);
  double* stage = smem + warp * 32 * kStageLd;
  double* my = stage + lane * kStageLd;
  ShellGeom<4> g;
  shell_geom<4>(A, e, g, nullptr);
  QuadInteg q;
  quad_integ(g, q);
  const Mat3& R = g.R;
  double Hb[4][4], Hg[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      double hb = 0., hg = 0.;
#pragma unroll
      for (int gp = 0; gp < 4; ++gp) {
        hb -= kNgp[gp][a] * (q.Wx[gp][b] * R.a[0][0] + q.Wy[gp][b] * R.a[1][0]);
        hg += (kNgp[gp][a] * kNgp[gp][b]) * q.dJ[gp];
      }
      Hb[a][b] = hb;
      Hg[a][b] = hg;
    }
  double zz[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) zz[i][j] = R.a[i][2] * R.a[j][2];
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    if (O.v[w] == nullptr) continue;
    double* out = O.v[w] + O.k0[w];
    const double sgn = (w == 2) ? -1. : 1.;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int j = 0; j < 3; ++j) my[i * 12 + b * 3 + j] = zz[i][j] * (w == 0 ? Hb[a][b] : sgn * Hg[a][b]);
      flush_chunk<36>(stage, out, e0, nvalid, 144, a * 36, O.acc[w] != 0, lane);
    }
  }
}

}  // namespace

cudaError_t launch_quad_aero(const EvalArgs& A, const AeroOut& O, cudaStream_t st) {
  if (A.ne <= 0) return cudaSuccess;
  const int64_t per_cta = 32 * kWarpsPerCta;
  const unsigned grid = unsigned((A.ne + per_cta - 1) / per_cta);
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(quad_aero_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kStageBytes));
  }
  quad_aero_kernel<<<grid, kThreads, kStageBytes, st>>>(A, O);
  return cudaGetLastError();
}

// Quad4Probe.update_BL (quad4.pyx:273-395): the 11 strain-interpolation rows at one natural point.
// out[11][24] in the reference's attribute order: BLexx BLeyy BLgxy BLkxx BLkyy BLkxy BLgyz_grad BLgyz_rot
// BLgxz_grad BLgxz_rot BLdrilling.  One thread per probe (n probes: xe[n][12], out[n][264]).
__global__ void quad4_BL_kernel(int64_t n, const double* __restrict__ xe, double xi, double eta, double* __restrict__ out) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double* p = xe + 12 * t;
  const double X[4] = {p[0], p[3], p[6], p[9]}, Y[4] = {p[1], p[4], p[7], p[10]};
  double wx[4], wy[4], dJ;
  jac_at(X, Y, xi, eta, wx, wy, dJ);
  const double xs[4] = {-1., 1., 1., -1.}, es[4] = {-1., -1., 1., 1.};
  double* o = out + 264 * t;
  for (int i = 0; i < 264; ++i) o[i] = 0.;
  for (int a = 0; a < 4; ++a) {
    const double Nx = wx[a] / dJ, Ny = wy[a] / dJ, N = 0.25 * (1. + xs[a] * xi) * (1. + es[a] * eta);
    const int c = 6 * a;
    o[0 * 24 + c + 0] = Nx;            // exx
    o[1 * 24 + c + 1] = Ny;            // eyy
    o[2 * 24 + c + 0] = Ny;            // gxy
    o[2 * 24 + c + 1] = Nx;
    o[3 * 24 + c + 4] = Nx;            // kxx
    o[4 * 24 + c + 3] = -Ny;           // kyy
    o[5 * 24 + c + 3] = -Nx;           // kxy
    o[5 * 24 + c + 4] = Ny;
    o[6 * 24 + c + 2] = Ny;            // gyz_grad
    o[7 * 24 + c + 3] = -N;            // gyz_rot
    o[8 * 24 + c + 2] = Nx;            // gxz_grad
    o[9 * 24 + c + 4] = N;             // gxz_rot
    o[10 * 24 + c + 0] = Ny / 2.;      // drilling
    o[10 * 24 + c + 1] = -Nx / 2.;
    o[10 * 24 + c + 5] = N;
  }
}

cudaError_t launch_quad4_BL(int64_t n, const double* xe, double xi, double eta, double* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  quad4_BL_kernel<<<unsigned((n + 127) / 128), 128, 0, st>>>(n, xe, xi, eta, out);
  return cudaGetLastError();
}

cudaError_t launch_quad(int kind, const EvalArgs& A, cudaStream_t st) {
  if (A.ne <= 0) return cudaSuccess;
  const int64_t per_cta = 32 * kWarpsPerCta;
  const unsigned grid = unsigned((A.ne + per_cta - 1) / per_cta);
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(quad_eval_kernel<PF3_QUAD4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kStageBytes));
    cudaFuncSetAttribute(quad_eval_kernel<PF3_QUAD4R, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kStageBytes));
    cudaFuncSetAttribute(quad_eval_kernel<PF3_QUAD4, PF3_FINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kStageBytes));
    cudaFuncSetAttribute(quad_eval_kernel<PF3_QUAD4R, PF3_FINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kStageBytes));
  }
  const bool fint_only = (A.what & ~PF3_FINT) == 0;   // update_fint / update_probe_finte / state: no matrix code
  if (kind == PF3_QUAD4) {
    if (fint_only) quad_eval_kernel<PF3_QUAD4, PF3_FINT><<<grid, kThreads, kStageBytes, st>>>(A);
    else quad_eval_kernel<PF3_QUAD4, 0><<<grid, kThreads, kStageBytes, st>>>(A);
  } else {
    if (fint_only) quad_eval_kernel<PF3_QUAD4R, PF3_FINT><<<grid, kThreads, kStageBytes, st>>>(A);
    else quad_eval_kernel<PF3_QUAD4R, 0><<<grid, kThreads, kStageBytes, st>>>(A);
  }
  return cudaGetLastError();
}

}  // namespace pf3
