// Shell element device math shared by the Quad4 / Quad4R / Tria3R kernels.
//
// Formulation.  The reference evaluates, per Gauss point and per
// entry, a 36-term B^T C B product (quad4.pyx:1002-1030).  Because the laminate
// matrices are constant over an element, the same matrix is
//     K_ab = sum_{p,q in {x,y}} coef_pq * G^pq_ab,   G^pq_ab = sum_gp detJ N_a,p N_b,q
// so a thread first reduces the Gauss points into the 2x2 "gradient Gram" of a node
// pair (16 FMA) and then forms every entry of the 6x6 node-pair block with 4 FMAs,
// followed by the R (.) R^T rotation of its four 3x3 sub-blocks written sparsely.
#pragma once
#include "common.cuh"

namespace pf3 {

struct ShellCoef {
  double A[6], B[6], D[6];  // 11 12 16 22 26 66, element axes
  double E44, E45, E55;     // already multiplied by the shear correction factors
  double h, rho0, rho1, rho2;
};

template <int NN>
struct ShellGeom {
  Mat3 R;                  // columns = element x,y,z axes in global coordinates
  double m11, m12, m21, m22;
  // local coordinates RELATIVE TO NODE 0 (every matrix is translation-invariant); O = R^T x_0 is the absolute local
  // position of node 0, only needed for probe.xe.  The differences are formed in GLOBAL coordinates first and then
  // rotated, so that elements far from the origin (|x| / h ~ 1e3 on the 2000 x 2000 benchmark plate) keep full
  // precision; the reference rotates absolute positions and subtracts afterwards (quad4.pyx:724-728, :939-942) and
  // loses |x| / h * eps there -- its results move by ~1e-12 under a rigid translation, ours do not.
  double X[NN], Y[NN], Z[NN];
  double O[3];
  double area;
};

// S = 3x3 symmetric (11 12 16 22 26 66).  p-type column: (N,x ; 0 ; N,y), q-type: (0 ; N,y ; N,x)
__device__ __forceinline__ double f_pp(const double* S, double gxx, double gxy, double gyx, double gyy) {
  return S[0] * gxx + S[2] * (gxy + gyx) + S[5] * gyy;
}
__device__ __forceinline__ double f_pq(const double* S, double gxx, double gxy, double gyx, double gyy) {
  return S[2] * gxx + S[1] * gxy + S[5] * gyx + S[4] * gyy;
}
__device__ __forceinline__ double f_qp(const double* S, double gxx, double gxy, double gyx, double gyy) {
  return S[2] * gxx + S[5] * gxy + S[1] * gyx + S[4] * gyy;
}
__device__ __forceinline__ double f_qq(const double* S, double gxx, double gxy, double gyx, double gyy) {
  return S[5] * gxx + S[4] * (gxy + gyx) + S[3] * gyy;
}

// Material-axis state, reference quad4.pyx:583-624 (same block in tria3r.pyx).
// m is left at identity when xmat is absent, null, or parallel to the normal.
template <int NN>
__device__ __forceinline__ void material_axes(ShellGeom<NN>& g, const double* xm, double znorm,
                                              const double* mprev) {
  g.m11 = 1.;
  g.m12 = 0.;
  g.m21 = 0.;
  g.m22 = 1.;
  if (mprev != nullptr) {  // sticky state of a per-element object (quad4.pyx:485-488, 588, 598)
    g.m11 = mprev[0];
    g.m12 = mprev[1];
    g.m21 = mprev[2];
    g.m22 = mprev[3];
  }
  if (xm == nullptr) return;
  const double tol = znorm / 1e10;
  double v[3] = {xm[0], xm[1], xm[2]};
  double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (!(n > tol)) return;
  const double inv = 1. / n;
  v[0] *= inv;
  v[1] *= inv;
  v[2] *= inv;
  const double z[3] = {g.R.a[0][2], g.R.a[1][2], g.R.a[2][2]};
  const double xh[3] = {g.R.a[0][0], g.R.a[1][0], g.R.a[2][0]};
  const double yh[3] = {g.R.a[0][1], g.R.a[1][1], g.R.a[2][1]};
  double ym[3];
  cross3(z, v, ym);
  double ny = normalize3(ym);
  if (!(ny > tol)) return;
  double xp[3];
  cross3(ym, z, xp);
  normalize3(xp);
  double c = dot3(xp, xh);
  double s = sqrt(1. - c * c);
  g.m11 = c;
  g.m22 = c;
  g.m12 = (dot3(xp, yh) > 0) ? -s : s;
  g.m21 = -g.m12;
}

// Quad4/Quad4R.update_rotation_matrix + update_probe_xe + update_area
// (quad4.pyx:491-624, 682-752); Tria3R: tria3r.pyx:294-424, 480-547.
// EARLY_U: gather the displacements together with the coordinates instead of after the frame (one exposed DRAM round
// trip less: update_fint 0.99 -> 0.92 ms at 4 M Quad4).
template <int NN, bool EARLY_U = false>
__device__ __forceinline__ void shell_geom(const EvalArgs& A, int64_t e, ShellGeom<NN>& g, double* ue) {
  const bool from_state = A.state != nullptr;
  const bool need_x = !from_state || (A.state_flags & PF3_STATE_REFRESH_XE);
  const bool need_u = ue != nullptr && (!from_state || (A.state_flags & PF3_STATE_REFRESH_UE));
  int64_t cn[NN];
  double P[NN][3], U[NN][6];
#pragma unroll
  for (int a = 0; a < NN; ++a) {
    cn[a] = A.conn[e * NN + a];
    if (need_x) {
#pragma unroll
      for (int i = 0; i < 3; ++i) P[a][i] = A.x[3 * cn[a] + i];
    }
    if (EARLY_U && need_u) {
#pragma unroll
      for (int i = 0; i < 6; ++i) U[a][i] = A.u[6 * cn[a] + i];
    }
  }
  double xh[3], yh[3], zh[3];
  if (from_state) {
    const double* s = A.state + e * PF3_STATE_STRIDE;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) g.R.a[i][j] = s[3 * i + j];
    g.m11 = s[9];
    g.m12 = s[10];
    g.m21 = s[11];
    g.m22 = s[12];
    g.area = s[13];
    g.O[0] = s[14];
    g.O[1] = s[15];
    g.O[2] = s[16];
#pragma unroll
    for (int a = 0; a < NN; ++a) {
      g.X[a] = s[14 + 3 * a] - g.O[0];
      g.Y[a] = s[15 + 3 * a] - g.O[1];
      g.Z[a] = s[16 + 3 * a] - g.O[2];
    }
    if (ue != nullptr)
#pragma unroll
      for (int i = 0; i < 6 * NN; ++i) ue[i] = s[26 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      xh[i] = g.R.a[i][0];
      yh[i] = g.R.a[i][1];
      zh[i] = g.R.a[i][2];
    }
  } else {
    double znorm;
    if (NN == 4) {
      double v13[3], v42[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        v13[i] = P[2][i] - P[0][i];
        v42[i] = P[1][i] - P[NN - 1][i];
      }
      cross3(v42, v13, zh);
      znorm = normalize3(zh);
#pragma unroll
      for (int i = 0; i < 3; ++i) xh[i] = (v13[i] + v42[i]) / 2.;
      normalize3(xh);
    } else {
      double v12[3], v13[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        v12[i] = P[1][i] - P[0][i];
        v13[i] = P[2][i] - P[0][i];
        xh[i] = v12[i];
      }
      cross3(v12, v13, zh);
      znorm = normalize3(zh);
      normalize3(xh);
    }
    cross3(zh, xh, yh);
    normalize3(yh);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      g.R.a[i][0] = xh[i];
      g.R.a[i][1] = yh[i];
      g.R.a[i][2] = zh[i];
    }
    const double* xm = nullptr;
    if (A.evec != nullptr) xm = A.evec + e * int64_t(A.evec_stride);
    const double* mprev = nullptr;
    if (A.eparam != nullptr && A.eparam[e * PF3_EPARAM_STRIDE + 7] != 0.) mprev = A.eparam + e * PF3_EPARAM_STRIDE + 8;
    material_axes<NN>(g, xm, znorm, mprev);
  }
  if (need_x) {
    g.O[0] = xh[0] * P[0][0] + xh[1] * P[0][1] + xh[2] * P[0][2];
    g.O[1] = yh[0] * P[0][0] + yh[1] * P[0][1] + yh[2] * P[0][2];
    g.O[2] = zh[0] * P[0][0] + zh[1] * P[0][1] + zh[2] * P[0][2];
    g.X[0] = g.Y[0] = g.Z[0] = 0.;
#pragma unroll
    for (int a = 1; a < NN; ++a) {
      const double q[3] = {P[a][0] - P[0][0], P[a][1] - P[0][1], P[a][2] - P[0][2]};
      g.X[a] = xh[0] * q[0] + xh[1] * q[1] + xh[2] * q[2];
      g.Y[a] = yh[0] * q[0] + yh[1] * q[1] + yh[2] * q[2];
      g.Z[a] = zh[0] * q[0] + zh[1] * q[1] + zh[2] * q[2];
    }
    if (NN == 4) {
      g.area = 0.5 * fabs((g.X[0] * g.Y[1] + g.X[1] * g.Y[2] + g.X[2] * g.Y[NN - 1] + g.X[NN - 1] * g.Y[0]) -
                          (g.X[1] * g.Y[0] + g.X[2] * g.Y[1] + g.X[NN - 1] * g.Y[2] + g.X[0] * g.Y[NN - 1]));
    } else {
      g.area = fabs((-g.X[0] + g.X[1]) * (-g.Y[0] + g.Y[2]) / 2. + (g.X[0] - g.X[2]) * (-g.Y[0] + g.Y[1]) / 2.);
    }
  }
  if (need_u) {
    // update_probe_ue (quad4.pyx:627-679): local = R^T global, per 3-DOF triplet
#pragma unroll
    for (int a = 0; a < NN; ++a)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        double u0, u1, u2;
        if (EARLY_U) {
          u0 = U[a][3 * t], u1 = U[a][3 * t + 1], u2 = U[a][3 * t + 2];
        } else {
          const double* ug = A.u + 6 * cn[a] + 3 * t;
          u0 = ug[0], u1 = ug[1], u2 = ug[2];
        }
        ue[6 * a + 3 * t + 0] = xh[0] * u0 + xh[1] * u1 + xh[2] * u2;
        ue[6 * a + 3 * t + 1] = yh[0] * u0 + yh[1] * u1 + yh[2] * u2;
        ue[6 * a + 3 * t + 2] = zh[0] * u0 + zh[1] * u1 + zh[2] * u2;
      }
  }
}

template <int NN>
__device__ __forceinline__ void store_state(const EvalArgs& A, int64_t e, const ShellGeom<NN>& g,
                                            const double* ue) {
  double* s = A.state_out + e * PF3_STATE_STRIDE;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) s[3 * i + j] = g.R.a[i][j];
  s[9] = g.m11;
  s[10] = g.m12;
  s[11] = g.m21;
  s[12] = g.m22;
  s[13] = g.area;
#pragma unroll
  for (int a = 0; a < NN; ++a) {
    s[14 + 3 * a] = g.O[0] + g.X[a];
    s[15 + 3 * a] = g.O[1] + g.Y[a];
    s[16 + 3 * a] = g.O[2] + g.Z[a];
  }
  for (int i = 0; i < 6 * NN; ++i) s[26 + i] = ue ? ue[i] : 0.;
  for (int i = 26 + 6 * NN; i < PF3_STATE_STRIDE; ++i) s[i] = 0.;
}

// T S T^T for the material-axis rotation (quad4.pyx:871-899), 6 unique outputs
__device__ __forceinline__ void rotate_sym3(const double (*T)[3], const double* S, double* O) {
  const double F[3][3] = {{S[0], S[1], S[2]}, {S[1], S[3], S[4]}, {S[2], S[4], S[5]}};
  double P[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) P[i][j] = T[i][0] * F[0][j] + T[i][1] * F[1][j] + T[i][2] * F[2][j];
  O[0] = P[0][0] * T[0][0] + P[0][1] * T[0][1] + P[0][2] * T[0][2];
  O[1] = P[0][0] * T[1][0] + P[0][1] * T[1][1] + P[0][2] * T[1][2];
  O[2] = P[0][0] * T[2][0] + P[0][1] * T[2][1] + P[0][2] * T[2][2];
  O[3] = P[1][0] * T[1][0] + P[1][1] * T[1][1] + P[1][2] * T[1][2];
  O[4] = P[1][0] * T[2][0] + P[1][1] * T[2][1] + P[1][2] * T[2][2];
  O[5] = P[2][0] * T[2][0] + P[2][1] * T[2][1] + P[2][2] * T[2][2];
}

template <int NN>
__device__ __forceinline__ void shell_coef(const EvalArgs& A, int64_t e, const ShellGeom<NN>& g,
                                           ShellCoef& c) {
  const double* p = A.props + prop_index(A, e) * PF3_SHELLPROP_STRIDE;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    c.A[i] = p[i];
    c.B[i] = p[6 + i];
    c.D[i] = p[12 + i];
  }
  if (g.m12 != 0.) {  // quad4.pyx:847: m12 is the "material axis defined" switch
    const double T[3][3] = {{g.m11 * g.m11, g.m12 * g.m12, 2. * g.m11 * g.m12},
                            {g.m21 * g.m21, g.m22 * g.m22, 2. * g.m21 * g.m22},
                            {g.m11 * g.m21, g.m12 * g.m22, g.m11 * g.m22 + g.m12 * g.m21}};
    double O[6];
    rotate_sym3(T, c.A, O);
#pragma unroll
    for (int i = 0; i < 6; ++i) c.A[i] = O[i];
    rotate_sym3(T, c.B, O);
#pragma unroll
    for (int i = 0; i < 6; ++i) c.B[i] = O[i];
    rotate_sym3(T, c.D, O);
#pragma unroll
    for (int i = 0; i < 6; ++i) c.D[i] = O[i];
  }
  const double k13 = p[21], k23 = p[22];
  c.E44 = p[18] * k23;                  // quad4.pyx:903-905
  c.E45 = p[19] * 0.5 * (k13 + k23);
  c.E55 = p[20] * k13;
  c.h = p[23];
  c.rho0 = p[24];
  c.rho1 = p[25];
  c.rho2 = p[26];
}

// 6x6 nodal inertia block in GLOBAL axes: T6 m_l T6^T with
// m_l = diag(r0,r0,r0,r2,r2,0) + r1 couplings (0,4),(4,0) and -r1 at (1,3),(3,1)
// (quad4.pyx:4644 ff.).  tt/rr symmetric, rt = tr^T.
struct NodalInertia {
  double tt[3][3], tr[3][3], rr[3][3];
};
__device__ __forceinline__ void nodal_inertia(const Mat3& R, double r0, double r1, double r2,
                                              NodalInertia& M) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      M.tt[i][j] = r0 * R.a[i][0] * R.a[j][0] + r0 * R.a[i][1] * R.a[j][1] + r0 * R.a[i][2] * R.a[j][2];
      M.tr[i][j] = r1 * R.a[i][0] * R.a[j][1] - r1 * R.a[i][1] * R.a[j][0];
      M.rr[i][j] = r2 * R.a[i][0] * R.a[j][0] + r2 * R.a[i][1] * R.a[j][1];
    }
}

}  // namespace pf3
