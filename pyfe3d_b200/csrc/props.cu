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

}  // namespace pf3
