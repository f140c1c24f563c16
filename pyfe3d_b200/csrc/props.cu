// Device-side laminate property table (first step): the 27 ShellProp scalars the shell elements
// read, for MANY laminates at once -- one property row per thread -- so that an optimisation loop that changes ply
// thicknesses / angles every iteration (doc/source/tut-opt-lightweight-smeared-deflection-SLSQP.ipynb, cell 3) never
// rebuilds ShellProp objects on the host or re-uploads a [nrows, 32] table.
//
// Follows, per row: laminated_plate (pyfe3d/shellprop_utils.py:96-179), read_laminaprop (:13-93: nu21 = nu12 e2/e1),
// Lamina.rebuild (pyfe3d/shellprop.pyx:278-337: plane-stress Q-bar and rotated transverse shear moduli),
// ShellProp.calc_constitutive_matrix (:568-621) and ShellProp.calc_scf (:485-548, Vlachoutsis one-factor formula).
#include "common.cuh"

namespace pf3 {

namespace {

__global__ void __launch_bounds__(128) k_laminate_props(int64_t nrows, int nplies, const double* __restrict__ theta,
                                                        int64_t theta_stride, const double* __restrict__ plyt,
                                                        int64_t plyt_stride, const double* __restrict__ lamina,
                                                        int64_t lamina_stride, const double* __restrict__ offset,
                                                        int64_t offset_stride, int calc_scf,
                                                        double* __restrict__ out) {
  const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const double* th = theta + row * theta_stride;
  const double* tk = plyt + row * plyt_stride;
  const double* lm = lamina + row * lamina_stride;
  const double o = offset ? offset[row * offset_stride] : 0.;
  double h = 0.;
  for (int p = 0; p < nplies; ++p) h += tk[p];
  double A[6] = {0, 0, 0, 0, 0, 0}, B[6] = {0, 0, 0, 0, 0, 0}, D[6] = {0, 0, 0, 0, 0, 0}, E[3] = {0, 0, 0};
  double r0 = 0., r1 = 0., r2 = 0.;
  const double zb = -h / 2. + o;
  double z = zb;
  // calc_scf accumulators (shellprop.pyx:512-546)
  double D1 = 0., R1 = 0., den1 = 0., D2 = 0., R2 = 0., den2 = 0.;
  for (int p = 0; p < nplies; ++p) {
    const double* m = lm + 8 * p;   // e1 e2 nu12 g12 g13 g23 rho pad
    const double e1 = m[0], e2 = m[1], nu12 = m[2], g12 = m[3], g13 = m[4], g23 = m[5], rho = m[6];
    const double nu21 = nu12 * e2 / e1;
    const double t = th[p] * 0.017453292519943295769236907684886;   // deg2rad
    const double c = cos(t), s = sin(t);
    const double c2 = c * c, c3 = c2 * c, c4 = c2 * c2, s2 = s * s, s3 = s2 * s, s4 = s2 * s2, sc = s * c;
    const double den = 1. - nu12 * nu21;
    const double q11 = e1 / den, q12 = nu12 * e2 / den, q22 = e2 / den, q44 = g23, q55 = g13, q66 = g12;
    double Q[6];
    Q[0] = q11 * c4 + 2 * (q12 + 2 * q66) * s2 * c2 + q22 * s4;                            // q11L
    Q[1] = (q11 + q22 - 4 * q66) * s2 * c2 + q12 * (s4 + c4);                              // q12L
    Q[2] = (q11 - q12 - 2 * q66) * s * c3 + (q12 - q22 + 2 * q66) * s3 * c;                // q16L
    Q[3] = q11 * s4 + 2 * (q12 + 2 * q66) * s2 * c2 + q22 * c4;                            // q22L
    Q[4] = (q11 - q12 - 2 * q66) * s3 * c + (q12 - q22 + 2 * q66) * s * c3;                // q26L
    Q[5] = (q11 + q22 - 2 * q12 - 2 * q66) * s2 * c2 + q66 * (s4 + c4);                    // q66L
    const double q44L = q44 * c2 + q55 * s2, q45L = (q55 - q44) * sc, q55L = q55 * c2 + q44 * s2;
    const double z0 = z, z1 = z + tk[p];
    z = z1;
    r0 += rho * (z1 - z0);
    r1 += rho * (z1 * z1 / 2. - z0 * z0 / 2.);
    r2 += rho * (z1 * z1 * z1 / 3. - z0 * z0 * z0 / 3.);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      A[i] += Q[i] * (z1 - z0);
      B[i] += 1 / 2. * Q[i] * (z1 * z1 - z0 * z0);
      D[i] += 1 / 3. * Q[i] * (z1 * z1 * z1 - z0 * z0 * z0);
    }
    E[0] += q44L * (z1 - z0);
    E[1] += q45L * (z1 - z0);
    E[2] += q55L * (z1 - z0);
    if (calc_scf) {
      const double ee1 = e1 * c + e2 * s, ee2 = e2 * c + e1 * s;
      const double n12 = nu12 * c + nu21 * s, n21 = nu21 * c + nu12 * s;
      const double a = z0, bb = z1;
      const double a2 = a * a, a3 = a2 * a, a4 = a2 * a2, a5 = a4 * a, b2 = bb * bb, b3 = b2 * bb, b4 = b2 * b2, b5 = b4 * bb;
      const double poly = 15 * o * a4 + 30 * o * a2 * zb * (2 * o - zb) - 15 * o * b4 + 30 * o * b2 * zb * (-2 * o + zb) -
                          3 * a5 + 10 * a3 * (-2 * o * o - 2 * o * zb + zb * zb) -
                          15 * a * zb * zb * (4 * o * o - 4 * o * zb + zb * zb) + 3 * b5 +
                          10 * b3 * (2 * o * o + 2 * o * zb - zb * zb) + 15 * bb * zb * zb * (4 * o * o - 4 * o * zb + zb * zb);
      const double cub = (bb - o) * (bb - o) * (bb - o) / 3. - (a - o) * (a - o) * (a - o) / 3.;
      D1 += ee1 / (1 - n12 * n21);
      R1 += D1 * cub;
      den1 += g13 * tk[p] * (h / tk[p]) * D1 * D1 * poly / (60 * g13);
      D2 += ee2 / (1 - n12 * n21);
      R2 += D2 * cub;
      den2 += g23 * tk[p] * (h / tk[p]) * D2 * D2 * poly / (60 * g23);
    }
  }
  double* w = out + row * PF3_SHELLPROP_STRIDE;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    w[i] = A[i];
    w[6 + i] = B[i];
    w[12 + i] = D[i];
  }
  w[18] = E[0];
  w[19] = E[1];
  w[20] = E[2];
  w[21] = calc_scf ? R1 * R1 / den1 : 5. / 6.;
  w[22] = calc_scf ? R2 * R2 / den2 : 5. / 6.;
  w[23] = h;
  w[24] = r0;
  w[25] = r1;
  w[26] = r2;
#pragma unroll
  for (int i = 27; i < PF3_SHELLPROP_STRIDE; ++i) w[i] = 0.;
}


// Lamination-parameter form (second step): total thickness + the seven material invariants +
// the 14 lamination parameters -> the same property row, following shellprop_from_LaminationParameters
// (pyfe3d/shellprop.pyx:767-815), and the rows of GradABDE.calc_LP_grad (:933-1014) laid out as PROPERTY rows, so
// that -- the element matrices being linear in A, B, D, E and in the mass integrals -- update_KC0 / update_M fed
// with gradient row v ARE dKC0/dv and dM/dv of the element.  One warp per laminate, one lane per slot of the row:
// every row leaves as one coalesced 256-byte store.
//   slot  0..17 : A11 A12 A16 A22 A26 A66, B.., D..  = fac_f * (c_f * K_i + sum_k G_ik xi_f,k)
//                 fac = h, h^2/4, h^3/12; c = 1, 0, 1; K = (u1, u4, 0, u1, 0, u5); G = "gradinv" of shellprop.pyx:956
//   slot 18..20 : E44 E45 E55 = h (u6 + u7 xiE1), -h u7 xiE2, h (u6 - u7 xiE1)
//   slot 21..26 : scf_k13 = scf_k23 = 5/6 (the reference never calls calc_scf on this path), h, and -- only when a
//                 density is given, an extension: the reference leaves them 0 -- rho h, 0, rho h^3 / 12
// Variables of the gradient rows: 0 = h, 1..4 = xiA1..4, 5..8 = xiB1..4, 9..12 = xiD1..4, 13..14 = xiE1..2; the rows
// selected by var_mask are stored consecutively per laminate.  Gradient rows repeat scf and h (the element's
// thin/thick switch, quad4.pyx:1032, must take the same branch) and carry d(mass integrals)/dh in row 0.
// grad_complete == 0 reproduces the reference's range(5) loops (:968, :981, :994): d(X66)/d(xi) stays 0.
__global__ void __launch_bounds__(128) k_lp_props(int64_t nrows, const double* __restrict__ thick, int64_t thick_stride,
                                                  const double* __restrict__ inv, int64_t inv_stride,
                                                  const double* __restrict__ lp, int64_t lp_stride,
                                                  const double* __restrict__ rho, int64_t rho_stride, int var_mask,
                                                  int grad_complete, double* __restrict__ out,
                                                  double* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  const double h = thick[row * thick_stride];
  const double* u = inv + row * inv_stride;
  const double* xi = lp + row * lp_stride;
  const double u1 = u[0], u2 = u[1], u3 = u[2], u4 = u[3], u5 = u[4], u6 = u[5], u7 = u[6];
  const double r = rho ? rho[row * rho_stride] : 0.;
  const int fam = lane < 18 ? lane / 6 : 3;   // 0 A, 1 B, 2 D, 3 E / tail
  const int i = lane < 18 ? lane % 6 : lane - 18;
  double fac = 0., dfac = 0., cst = 0., g[4] = {0., 0., 0., 0.}, x[4] = {0., 0., 0., 0.};
  if (fam < 3) {
    fac = fam == 0 ? h : fam == 1 ? h * h / 4. : h * h * h / 12.;
    dfac = fam == 0 ? 1. : fam == 1 ? h / 2. : h * h / 4.;
    const double c = fam == 1 ? 0. : 1.;
    cst = c * (i == 0 || i == 3 ? u1 : i == 1 ? u4 : i == 5 ? u5 : 0.);
    g[0] = i == 0 ? u2 : i == 3 ? -u2 : 0.;
    g[1] = (i == 2 || i == 4) ? u2 / 2. : 0.;
    g[2] = (i == 0 || i == 3) ? u3 : (i == 1 || i == 5) ? -u3 : 0.;
    g[3] = i == 2 ? u3 : i == 4 ? -u3 : 0.;
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = xi[4 * fam + k];
  } else if (i < 3) {
    fac = h;
    dfac = 1.;
    cst = i == 1 ? 0. : u6;
    g[0] = i == 0 ? u7 : i == 2 ? -u7 : 0.;
    g[1] = i == 1 ? -u7 : 0.;
    x[0] = xi[12];
    x[1] = xi[13];
  }
  const double bracket = cst + g[0] * x[0] + g[1] * x[1] + g[2] * x[2] + g[3] * x[3];
  const bool stiff = lane < 21;
  // slots shared by every row: shear correction factors and thickness
  const double common = (lane == 21 || lane == 22) ? 5. / 6. : lane == 23 ? h : 0.;
  if (out) {
    double v = stiff ? fac * bracket : common;
    if (lane == 24) v = r * h;
    if (lane == 26) v = r * h * h * h / 12.;
    out[row * PF3_SHELLPROP_STRIDE + lane] = v;
  }
  if (!grad || !var_mask) return;
  double* w = grad + row * int64_t(__popc(var_mask)) * PF3_SHELLPROP_STRIDE + lane;
  if (var_mask & 1) {
    double v = stiff ? dfac * bracket : common;
    if (lane == 24) v = r;
    if (lane == 26) v = r * h * h / 4.;
    *w = v;
    w += PF3_SHELLPROP_STRIDE;
  }
  const bool keep = grad_complete || fam == 3 || i < 5;
#pragma unroll
  for (int f = 0; f < 3; ++f)
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (var_mask >> (1 + 4 * f + k) & 1) {
        *w = (fam == f && keep) ? fac * g[k] : common;
        w += PF3_SHELLPROP_STRIDE;
      }
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (var_mask >> (13 + k) & 1) {
      *w = (fam == 3 && i < 3) ? fac * g[k] : common;
      w += PF3_SHELLPROP_STRIDE;
    }
}

}  // namespace

cudaError_t launch_laminate_props(int64_t nrows, int nplies, const double* theta, int64_t theta_stride,
                                  const double* plyt, int64_t plyt_stride, const double* lamina, int64_t lamina_stride,
                                  const double* offset, int64_t offset_stride, int calc_scf, double* out, cudaStream_t st) {
  if (nrows <= 0) return cudaSuccess;
  k_laminate_props<<<unsigned((nrows + 127) / 128), 128, 0, st>>>(nrows, nplies, theta, theta_stride, plyt, plyt_stride,
                                                                lamina, lamina_stride, offset, offset_stride, calc_scf,
                                                                out);
  return cudaGetLastError();
}

cudaError_t launch_lp_props(int64_t nrows, const double* thick, int64_t thick_stride, const double* inv,
                            int64_t inv_stride, const double* lp, int64_t lp_stride, const double* rho,
                            int64_t rho_stride, int var_mask, int grad_complete, double* out, double* grad,
                            cudaStream_t st) {
  if (nrows <= 0) return cudaSuccess;
  k_lp_props<<<unsigned((nrows + 3) / 4), 128, 0, st>>>(nrows, thick, thick_stride, inv, inv_stride, lp, lp_stride, rho,
                                                        rho_stride, var_mask, grad_complete, out, grad);
  return cudaGetLastError();
}

}  // namespace pf3
