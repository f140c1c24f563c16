// Tria3R: KC0 / KG / KG_given_stress / M / fint, one element per thread.
//
// Replaces (reference, /root/reference/pyfe3d/tria3r.pyx): update_rotation_matrix :294,
// update_probe_ue :426, update_probe_xe :480, update_probe_finte :549, update_KC0 :950,
// update_fint :2975, update_KG :3020, update_KG_given_stress :3576, update_M :4063.
#include "shell.cuh"

namespace pf3 {

namespace {

constexpr int kTriChunkKC0 = 54;  // 3 rows x 18 columns
constexpr int kTriChunkM = 45;    // 3 rows x 15
constexpr int kTriChunkKG = 27;   // 3 rows x 9

__global__ void __launch_bounds__(kThreads) tria_eval_kernel(const EvalArgs A) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t e0 = (int64_t(blockIdx.x) * kWarpsPerCta + warp) * 32;
  if (e0 >= A.ne) return;
  const int nvalid = int(min(int64_t(32), A.ne - e0));
  const int64_t e = e0 + min(lane, nvalid - 1);
  if (A.state == nullptr) conn_prefetch(A.conn, 3, e0, A.ne, lane);
  double* stage = smem + warp * 32 * kStageLd;
  double* my = stage + lane * kStageLd;

  const bool have_u = (A.u != nullptr || A.state != nullptr);
  double ue[18];
  ShellGeom<3> g;
  shell_geom<3, true>(A, e, g, have_u ? ue : nullptr);
  if (A.state_out != nullptr) {
    if (lane < nvalid) store_state<3>(A, e, g, have_u ? ue : nullptr);
    if (A.what == 0) return;
  }
  ShellCoef c;
  shell_coef<3>(A, e, g, c);
  const Mat3& R = g.R;
  double K6ROT = 100., alpha = 0.7;
  if (A.eparam != nullptr) {
    K6ROT = A.eparam[e * PF3_EPARAM_STRIDE + 0];
    alpha = A.eparam[e * PF3_EPARAM_STRIDE + 1];
  }
  const double dJ = 2. * g.area;  // tria3r.pyx:1017
  const double i2a = 1. / (2. * g.area);
  // constant gradients (tria3r.pyx:2116-2121)
  const double Nx[3] = {(g.Y[1] - g.Y[2]) * i2a, (-g.Y[0] + g.Y[2]) * i2a, (g.Y[0] - g.Y[1]) * i2a};
  const double Ny[3] = {(-g.X[1] + g.X[2]) * i2a, (g.X[0] - g.X[2]) * i2a, (-g.X[0] + g.X[1]) * i2a};
  // shear-locking relief (tria3r.pyx:1109-1122)
  {
    const double l12 = sqrt((g.X[0] - g.X[1]) * (g.X[0] - g.X[1]) + (g.Y[0] - g.Y[1]) * (g.Y[0] - g.Y[1]));
    const double l23 = sqrt((g.X[1] - g.X[2]) * (g.X[1] - g.X[2]) + (g.Y[1] - g.Y[2]) * (g.Y[1] - g.Y[2]));
    const double l31 = sqrt((g.X[2] - g.X[0]) * (g.X[2] - g.X[0]) + (g.Y[2] - g.Y[0]) * (g.Y[2] - g.Y[0]));
    double maxl = l12;
    if (l23 > maxl) maxl = l23;
    if (l31 > maxl) maxl = l31;
    const double fac = 1. / (1. + alpha * maxl * maxl / (c.h * c.h));
    c.E44 *= fac;
    c.E45 *= fac;
    c.E55 *= fac;
  }
  const double w = dJ * 0.5;                          // one point, weight 1/2
  const double kd = 1e-6 * K6ROT * c.A[5] * w;        // drilling penalty x total weight (3 pts x dJ/6)
  const double third = 0.333333333333333333333333333;
  double tS[3], sS[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    tS[a] = w * (c.E44 * Ny[a] + c.E45 * Nx[a]);
    sS[a] = w * (c.E45 * Ny[a] + c.E55 * Nx[a]);
  }
  const double c44 = w * c.E44 * third * third, c45 = w * c.E45 * third * third, c55 = w * c.E55 * third * third;

  // ------------------------------------------------------------------ KG
  if (A.what & (PF3_KG | PF3_KG_STRESS)) {
    double Nxx = A.Nxx, Nyy = A.Nyy, Nxy = A.Nxy;
    if (!(A.what & PF3_KG_STRESS)) {
      double exx = 0, eyy = 0, gxy = 0, kxx = 0, kyy = 0, kxy = 0;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        exx += Nx[a] * ue[6 * a];
        eyy += Ny[a] * ue[6 * a + 1];
        gxy += Ny[a] * ue[6 * a] + Nx[a] * ue[6 * a + 1];
        kxx += Nx[a] * ue[6 * a + 4];
        kyy -= Ny[a] * ue[6 * a + 3];
        kxy += Ny[a] * ue[6 * a + 4] - Nx[a] * ue[6 * a + 3];
      }
      Nxx = c.A[0] * exx + c.A[1] * eyy + c.A[2] * gxy + c.B[0] * kxx + c.B[1] * kyy + c.B[2] * kxy;
      Nyy = c.A[1] * exx + c.A[3] * eyy + c.A[4] * gxy + c.B[1] * kxx + c.B[3] * kyy + c.B[4] * kxy;
      Nxy = c.A[2] * exx + c.A[4] * eyy + c.A[5] * gxy + c.B[2] * kxx + c.B[4] * kyy + c.B[5] * kxy;
    }
    double zz[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) zz[i][j] = R.a[i][2] * R.a[j][2];
    double* out = A.kgv + A.kg_k0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double px = w * (Nx[a] * Nxx + Ny[a] * Nxy), py = w * (Nx[a] * Nxy + Ny[a] * Nyy);
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const double ge = Nx[b] * px + Ny[b] * py;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) my[i * 9 + b * 3 + j] = zz[i][j] * ge;
      }
      flush_chunk<kTriChunkKG>(stage, out, e0, nvalid, 81, a * 27, A.acc_kg != 0, lane);
    }
  }

  // ------------------------------------------------------------------ M
  if (A.what & PF3_M) {
    // H_ab: mtype 0 Cowper 3-point (tria3r.pyx:4200 ff.), 1: detJ/18 (:4108), 2: vertices
    double hd, ho;
    if (A.mtype == 0) {
      hd = dJ / 12.;
      ho = dJ / 24.;
    } else if (A.mtype == 1) {
      hd = ho = dJ / 18.;
    } else {
      hd = dJ / 6.;
      ho = 0.;
    }
    NodalInertia Mi;
    nodal_inertia(R, c.rho0, c.rho1, c.rho2, Mi);
    double* out = A.mv + A.m_k0;
    if (A.mtype != 2) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            const double h = (a == b) ? hd : ho;
            double* d = my + i * 15 + b * 5;
            d[0] = h * Mi.tt[i][0];
            d[1] = h * Mi.tt[i][1];
            d[2] = h * Mi.tt[i][2];
            d[3] = h * Mi.tr[i][(i == 0) ? 1 : 0];
            d[4] = h * Mi.tr[i][(i == 2) ? 1 : 2];
          }
        flush_chunk<kTriChunkM>(stage, out, e0, nvalid, 270, a * 90, A.acc_m != 0, lane);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            const double h = (a == b) ? hd : ho;
            double* d = my + i * 15 + b * 5;
            d[0] = h * Mi.tr[(i == 0) ? 1 : 0][i];
            d[1] = h * Mi.tr[(i == 2) ? 1 : 2][i];
            d[2] = h * Mi.rr[i][0];
            d[3] = h * Mi.rr[i][1];
            d[4] = h * Mi.rr[i][2];
          }
        flush_chunk<kTriChunkM>(stage, out, e0, nvalid, 270, a * 90 + 45, A.acc_m != 0, lane);
      }
    } else {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            const double h = (a == b) ? hd : ho;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              my[i * 9 + b * 3 + j] = h * Mi.tt[i][j];
              my[27 + i * 9 + b * 3 + j] = h * Mi.rr[i][j];
            }
          }
        flush_chunk<54>(stage, out, e0, nvalid, 270, a * 54, A.acc_m != 0, lane);
      }
    }
  }

  // ------------------------------------------------------------------ KC0
  if (A.what & PF3_KC0) {
    double* out = A.kc0v + A.kc0_k0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const double gxx = w * Nx[a] * Nx[b], gxy = w * Nx[a] * Ny[b];
        const double gyx = w * Ny[a] * Nx[b], gyy = w * Ny[a] * Ny[b];
        // in-plane part of the drilling penalty survives in update_KC0 (tria3r.pyx:2261-2325)
        const double uu = f_pp(c.A, gxx, gxy, gyx, gyy) + 0.25 * kd * Ny[a] * Ny[b];
        const double uv = f_pq(c.A, gxx, gxy, gyx, gyy) - 0.25 * kd * Ny[a] * Nx[b];
        const double vu = f_qp(c.A, gxx, gxy, gyx, gyy) - 0.25 * kd * Nx[a] * Ny[b];
        const double vv = f_qq(c.A, gxx, gxy, gyx, gyy) + 0.25 * kd * Nx[a] * Nx[b];
        const double ww = tS[a] * Ny[b] + sS[a] * Nx[b];
        double o[3][3];
        rot_block_diag5(R, uu, uv, vu, vv, ww, o);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) my[i * 18 + b * 6 + j] = o[i][j];
        // QUIRK kept for parity: the (u,v)-rz drilling couplings are never written by
        // update_KC0; update_fint keeps them (see the FINT section).
        rot_block_8(R, -f_pq(c.B, gxx, gxy, gyx, gyy), f_pp(c.B, gxx, gxy, gyx, gyy), 0.,
                    -f_qq(c.B, gxx, gxy, gyx, gyy), f_qp(c.B, gxx, gxy, gyx, gyy), 0., -third * tS[a],
                    third * sS[a], o);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) my[i * 18 + b * 6 + 3 + j] = o[i][j];
      }
      flush_chunk<kTriChunkKC0>(stage, out, e0, nvalid, 324, a * 108, A.acc_kc0 != 0, lane);
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const double gxx = w * Nx[a] * Nx[b], gxy = w * Nx[a] * Ny[b];
        const double gyx = w * Ny[a] * Nx[b], gyy = w * Ny[a] * Ny[b];
        double o[3][3];
        rot_block_8(R, -f_qp(c.B, gxx, gxy, gyx, gyy), -f_qq(c.B, gxx, gxy, gyx, gyy), -third * tS[b],
                    f_pp(c.B, gxx, gxy, gyx, gyy), f_pq(c.B, gxx, gxy, gyx, gyy), third * sS[b], 0., 0., o);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) my[i * 18 + b * 6 + j] = o[i][j];
        const double rzrz = (a == b) ? kd * (1. / 6.) : 0.;  // sum_p N_a(p)^2 / 3 = 1/6; a!=b dropped
        rot_block_diag5(R, f_qq(c.D, gxx, gxy, gyx, gyy) + c44, -f_qp(c.D, gxx, gxy, gyx, gyy) - c45,
                        -f_pq(c.D, gxx, gxy, gyx, gyy) - c45, f_pp(c.D, gxx, gxy, gyx, gyy) + c55, rzrz, o);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) my[i * 18 + b * 6 + 3 + j] = o[i][j];
      }
      flush_chunk<kTriChunkKC0>(stage, out, e0, nvalid, 324, a * 108 + 54, A.acc_kc0 != 0, lane);
    }
  }

  // ------------------------------------------------------------------ fint / finte
  if (A.what & PF3_FINT) {
    double f[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) f[i] = 0.;
    double eps[6] = {0, 0, 0, 0, 0, 0};
    double gyz = 0., gxz = 0., th = 0., rzs = 0.;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      eps[0] += Nx[a] * ue[6 * a];
      eps[1] += Ny[a] * ue[6 * a + 1];
      eps[2] += Ny[a] * ue[6 * a] + Nx[a] * ue[6 * a + 1];
      eps[3] += Nx[a] * ue[6 * a + 4];
      eps[4] -= Ny[a] * ue[6 * a + 3];
      eps[5] += Ny[a] * ue[6 * a + 4] - Nx[a] * ue[6 * a + 3];
      gyz += Ny[a] * ue[6 * a + 2] - third * ue[6 * a + 3];
      gxz += Nx[a] * ue[6 * a + 2] + third * ue[6 * a + 4];
      th += 0.5 * Ny[a] * ue[6 * a] - 0.5 * Nx[a] * ue[6 * a + 1];
      rzs += ue[6 * a + 5];
    }
    const double C6[6][6] = {{c.A[0], c.A[1], c.A[2], c.B[0], c.B[1], c.B[2]},
                             {c.A[1], c.A[3], c.A[4], c.B[1], c.B[3], c.B[4]},
                             {c.A[2], c.A[4], c.A[5], c.B[2], c.B[4], c.B[5]},
                             {c.B[0], c.B[1], c.B[2], c.D[0], c.D[1], c.D[2]},
                             {c.B[1], c.B[3], c.B[4], c.D[1], c.D[3], c.D[4]},
                             {c.B[2], c.B[4], c.B[5], c.D[2], c.D[4], c.D[5]}};
    double s[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      s[i] = 0.;
#pragma unroll
      for (int j = 0; j < 6; ++j) s[i] += C6[i][j] * eps[j];
      s[i] *= w;
    }
    const double ty = w * (c.E44 * gyz + c.E45 * gxz), tx = w * (c.E45 * gyz + c.E55 * gxz);
    // full drilling penalty over the 3 Cowper points (weights dJ/6 each):
    //   sum_p Bd(p) Bd(p)^T ue, with sum_p N_b(p) = 1 and sum_p N_a(p) N_b(p) = 1/2 (a=b), 1/4 (a!=b)
    const double kd3 = kd * third;  // per-point weight x penalty
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      // strain-like drilling measure: 3 * th (in-plane part) + sum_b (sum_p N_b(p)) rz_b = 3 th + rzs
      const double thsum = 3. * th + rzs;
      f[6 * a + 0] += Nx[a] * s[0] + Ny[a] * s[2] + 0.5 * Ny[a] * kd3 * thsum;
      f[6 * a + 1] += Ny[a] * s[1] + Nx[a] * s[2] - 0.5 * Nx[a] * kd3 * thsum;
      f[6 * a + 2] += Ny[a] * ty + Nx[a] * tx;
      f[6 * a + 3] += -(Ny[a] * s[4] + Nx[a] * s[5]) - third * ty;
      f[6 * a + 4] += Nx[a] * s[3] + Ny[a] * s[5] + third * tx;
      // rz_a: sum_p N_a(p) (th + sum_b N_b(p) rz_b) = th + (1/4) sum_b rz_b + (1/4) rz_a
      f[6 * a + 5] += kd3 * (th + 0.25 * rzs + 0.25 * ue[6 * a + 5]);
    }
    if (A.finte != nullptr) {
#pragma unroll
      for (int i = 0; i < 18; ++i) my[i] = f[i];
      flush_chunk<18>(stage, A.finte, e0, nvalid, 18, 0, false, lane);
    }
    if (A.fe != nullptr) {
#pragma unroll
      for (int t = 0; t < 6; ++t)
#pragma unroll
        for (int i = 0; i < 3; ++i)
          my[3 * t + i] = R.a[i][0] * f[3 * t] + R.a[i][1] * f[3 * t + 1] + R.a[i][2] * f[3 * t + 2];
      flush_chunk<18>(stage, A.fe, e0, nvalid, 18, 0, false, lane);
    }
  }
}

}  // namespace

cudaError_t launch_tria(const EvalArgs& A, cudaStream_t st) {
  if (A.ne <= 0) return cudaSuccess;
  const int64_t per_cta = 32 * kWarpsPerCta;
  const unsigned grid = unsigned((A.ne + per_cta - 1) / per_cta);
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(tria_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kStageBytes));
  }
  tria_eval_kernel<<<grid, kThreads, kStageBytes, st>>>(A);
  return cudaGetLastError();
}

}  // namespace pf3
