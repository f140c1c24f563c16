// BeamC, BeamLR, Truss, Spring: KC0 / KG / M / fint, one element per thread.
//
// Replaces (reference, /root/reference/pyfe3d):
//   beamc.pyx : update_rotation_matrix :158, update_probe_ue :234, update_probe_xe :287,
//               update_probe_finte :352, update_KC0 :489, update_fint :1354, update_KG :1392, update_M :2181
//   beamlr.pyx: :160, :236, :289, :354, update_KC0 :407, update_fint :1188, update_KG :1226, update_M :1464
//   truss.pyx : update_rotation_matrix :203, update_KC0 :386, update_fint :802, update_M :838
//   spring.pyx: update_rotation_matrix :147, update_KC0 :295, update_fint :708
//
// These elements are tiny (72..144 values each) and purely store-bound.  The closed-form local 12x12 matrices are
// written once, as lists of `sym(K, i, j, value)` with constant (i, j); they are instantiated per 6x6 NODE-PAIR BLOCK
// through a sink type (KBlock<A, B>) whose `set` keeps only the entries of its block, so that after inlining only
// that block's 36 values are computed and they live in registers (the first version kept the whole 12x12 in
// per-thread local memory: 1152 B of stack, 0.3-0.4 of the store roofline).
#include "common.cuh"

namespace pf3 {

namespace {

struct BeamP {
  double A, E, G, Iyy, Izz, Iyz, J, Ay, Az, r0, ry, rz, ry2, rz2, ryz;
};

// Sink that keeps rows [6A, 6A+6) x columns [6B, 6B+6) of the local matrix.  Every call site passes literal (i, j):
// the tests fold at compile time.
template <int BA, int BB>
struct KBlock {
  double v[6][6];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) v[i][j] = 0.;
  }
  __device__ __forceinline__ void set(int i, int j, double x) {
    if (i / 6 == BA && j / 6 == BB) v[i % 6][j % 6] = x;
  }
};
template <class KS>
__device__ __forceinline__ void sym(KS& K, int i, int j, double v) {
  K.set(i, j, v);
  K.set(j, i, v);
}

// Timoshenko beam with consistent shape functions (Luo 2008); beamc.pyx:543-624
template <class KS>
__device__ __forceinline__ void beamc_Ke(const BeamP& p, double L, KS& K) {
  K.zero();
  const double iL = 1. / L, iL2 = iL * iL, iL3 = iL2 * iL;   // reciprocals: a double division costs ~15 instructions
  (void)iL2; (void)iL3;
  const double L2 = L * L, L3 = L2 * L;
  (void)L3;
  const double ay = 12 * p.E * p.Izz / (p.G * p.A * L2), az = 12 * p.E * p.Iyy / (p.G * p.A * L2);
  const double by = 1 / (1. - ay), bz = 1 / (1. - az);
  const double Ky = by * by * (p.A * p.G * L2 * ay * ay + 12 * p.E * p.Izz);
  const double Kz = bz * bz * (p.A * p.G * L2 * az * az + 12 * p.E * p.Iyy);
  const double Kyz = p.E * p.Iyz * by * bz;
  const double Sy = p.Az * p.G * ay * by, Sz = p.Ay * p.G * az * bz;
  const double cz = p.Az * p.E * bz * (1 - az) * iL, cy = p.Ay * p.E * by * (ay - 1) * iL;
  const double EA = p.A * p.E * iL, GJ = p.G * p.J * iL;
  sym(K, 0, 0, EA); sym(K, 6, 6, EA); sym(K, 0, 6, -EA);
  sym(K, 0, 4, cz); sym(K, 6, 10, cz); sym(K, 0, 10, -cz); sym(K, 4, 6, -cz);
  sym(K, 0, 5, cy); sym(K, 6, 11, cy); sym(K, 0, 11, -cy); sym(K, 5, 6, -cy);
  sym(K, 1, 1, Ky * iL3); sym(K, 7, 7, Ky * iL3); sym(K, 1, 7, -Ky * iL3);
  sym(K, 1, 5, Ky * (0.5 * iL2)); sym(K, 1, 11, Ky * (0.5 * iL2)); sym(K, 5, 7, -Ky * (0.5 * iL2)); sym(K, 7, 11, -Ky * (0.5 * iL2));
  sym(K, 2, 2, Kz * iL3); sym(K, 8, 8, Kz * iL3); sym(K, 2, 8, -Kz * iL3);
  sym(K, 2, 4, -Kz * (0.5 * iL2)); sym(K, 2, 10, -Kz * (0.5 * iL2)); sym(K, 4, 8, Kz * (0.5 * iL2)); sym(K, 8, 10, Kz * (0.5 * iL2));
  sym(K, 1, 2, 12 * Kyz * iL3); sym(K, 7, 8, 12 * Kyz * iL3); sym(K, 1, 8, -12 * Kyz * iL3); sym(K, 2, 7, -12 * Kyz * iL3);
  sym(K, 1, 4, -6 * Kyz * iL2); sym(K, 1, 10, -6 * Kyz * iL2); sym(K, 5, 8, -6 * Kyz * iL2); sym(K, 8, 11, -6 * Kyz * iL2);
  sym(K, 2, 5, 6 * Kyz * iL2); sym(K, 2, 11, 6 * Kyz * iL2); sym(K, 4, 7, 6 * Kyz * iL2); sym(K, 7, 10, 6 * Kyz * iL2);
  sym(K, 3, 3, GJ); sym(K, 9, 9, GJ); sym(K, 3, 9, -GJ);
  sym(K, 1, 3, Sy * iL); sym(K, 7, 9, Sy * iL); sym(K, 1, 9, -Sy * iL); sym(K, 3, 7, -Sy * iL);
  sym(K, 3, 5, Sy / 2); sym(K, 3, 11, Sy / 2); sym(K, 5, 9, -Sy / 2); sym(K, 9, 11, -Sy / 2);
  sym(K, 2, 3, -Sz * iL); sym(K, 8, 9, -Sz * iL); sym(K, 2, 9, Sz * iL); sym(K, 3, 8, Sz * iL);
  sym(K, 3, 4, Sz / 2); sym(K, 3, 10, Sz / 2); sym(K, 4, 9, -Sz / 2); sym(K, 9, 10, -Sz / 2);
  const double EIy = p.E * p.Iyy, EIz = p.E * p.Izz, AGL2 = p.A * p.G * L2;
  const double k44 = bz * bz * (AGL2 * az * az / 4 + EIy * az * az - 2 * EIy * az + 4 * EIy) * iL;
  const double k55 = by * by * (AGL2 * ay * ay / 4 + EIz * ay * ay - 2 * EIz * ay + 4 * EIz) * iL;
  sym(K, 4, 4, k44); sym(K, 10, 10, k44);
  sym(K, 4, 10, bz * bz * (AGL2 * az * az / 4 - EIy * az * az + 2 * EIy * az + 2 * EIy) * iL);
  sym(K, 5, 5, k55); sym(K, 11, 11, k55);
  sym(K, 5, 11, by * by * (AGL2 * ay * ay / 4 - EIz * ay * ay + 2 * EIz * ay + 2 * EIz) * iL);
  const double k45 = Kyz * (-ay * az + ay + az - 4) * iL, k411 = Kyz * (ay * az - ay - az - 2) * iL;
  sym(K, 4, 5, k45); sym(K, 10, 11, k45); sym(K, 4, 11, k411); sym(K, 5, 10, k411);
}

// beamc.pyx:1889-2179
template <class KS>
__device__ __forceinline__ void beamc_KGe(const BeamP& p, double L, const double* ue, KS& K) {
  K.zero();
  const double iL = 1. / L, iL2 = iL * iL, iL3 = iL2 * iL;   // reciprocals: a double division costs ~15 instructions
  (void)iL2; (void)iL3;
  const double L2 = L * L;
  const double ay = 12 * p.E * p.Izz / (p.G * p.A * L2), az = 12 * p.E * p.Iyy / (p.G * p.A * L2);
  const double by = 1 / (1. - ay), bz = 1 / (1. - az);
  const double N = p.A * p.E * (-ue[0] + ue[6]) * iL;
  double v = N * by * by * (5 * ay * ay - 10 * ay + 6) * (0.2 * iL);
  sym(K, 1, 1, v); sym(K, 7, 7, v); sym(K, 1, 7, -v);
  v = N * bz * bz * (5 * az * az - 10 * az + 6) * (0.2 * iL);
  sym(K, 2, 2, v); sym(K, 8, 8, v); sym(K, 2, 8, -v);
  v = N * by * bz * (5 * ay * az - 5 * ay - 5 * az + 6) * (0.2 * iL);
  sym(K, 1, 2, v); sym(K, 7, 8, v); sym(K, 1, 8, -v); sym(K, 2, 7, -v);
  v = N * by * by * (1. / 10);
  sym(K, 1, 5, v); sym(K, 1, 11, v); sym(K, 5, 7, -v); sym(K, 7, 11, -v);
  v = -N * bz * bz * (1. / 10);
  sym(K, 2, 4, v); sym(K, 2, 10, v); sym(K, 4, 8, -v); sym(K, 8, 10, -v);
  v = -N * by * bz * (1. / 10);
  sym(K, 1, 4, v); sym(K, 1, 10, v); sym(K, 4, 7, -v); sym(K, 7, 10, -v);
  v = N * by * bz * (1. / 10);
  sym(K, 2, 5, v); sym(K, 2, 11, v); sym(K, 5, 8, -v); sym(K, 8, 11, -v);
  v = L * N * bz * bz * (5 * az * az - 10 * az + 8) * (1. / 60);
  sym(K, 4, 4, v); sym(K, 10, 10, v);
  v = L * N * by * by * (5 * ay * ay - 10 * ay + 8) * (1. / 60);
  sym(K, 5, 5, v); sym(K, 11, 11, v);
  sym(K, 4, 10, L * N * bz * bz * (-5 * az * az + 10 * az - 2) * (1. / 60));
  sym(K, 5, 11, L * N * by * by * (-5 * ay * ay + 10 * ay - 2) * (1. / 60));
  v = L * N * by * bz * (-5 * ay * az + 5 * ay + 5 * az - 8) * (1. / 60);
  sym(K, 4, 5, v); sym(K, 10, 11, v);
  v = L * N * by * bz * (5 * ay * az - 5 * ay - 5 * az + 2) * (1. / 60);
  sym(K, 4, 11, v); sym(K, 5, 10, v);
}

// beamc.pyx:2246-3152; mtype 0 consistent, 1 lumped
template <class KS>
__device__ __forceinline__ void beamc_Me(const BeamP& p, double L, int mtype, KS& K) {
  K.zero();
  const double iL = 1. / L, iL2 = iL * iL, iL3 = iL2 * iL;   // reciprocals: a double division costs ~15 instructions
  (void)iL2; (void)iL3;
  const double L2 = L * L;
  const double ay = 12 * p.E * p.Izz / (p.G * p.A * L2), az = 12 * p.E * p.Iyy / (p.G * p.A * L2);
  const double by = 1 / (1. - ay), bz = 1 / (1. - az);
  const double r0 = p.r0, ry = p.ry, rz = p.rz, ry2 = p.ry2, rz2 = p.rz2, ryz = p.ryz;
  if (mtype == 1) {
    const double d[6] = {L * r0 / 2, L * by * by * r0 * (ay - 1) * (ay - 1) / 2, L * bz * bz * r0 * (az - 1) * (az - 1) / 2,
                         L * (ry2 + rz2) / 2, L * bz * bz * rz2 * (az - 1) * (az - 1) / 2,
                         L * by * by * ry2 * (ay - 1) * (ay - 1) / 2};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      K.set(i, i, d[i]);
      K.set(i + 6, i + 6, d[i]);
    }
    return;
  }
  const double by2 = by * by, bz2 = bz * bz, ay2 = ay * ay, az2 = az * az, bb = by * bz;
  double v;
  sym(K, 0, 0, L * r0 * (1. / 3)); sym(K, 6, 6, L * r0 * (1. / 3)); sym(K, 0, 6, L * r0 * (1. / 6));
  v = by * ry / 2; sym(K, 0, 1, v); sym(K, 1, 6, v); sym(K, 0, 7, -v); sym(K, 6, 7, -v);
  v = bz * rz / 2; sym(K, 0, 2, v); sym(K, 2, 6, v); sym(K, 0, 8, -v); sym(K, 6, 8, -v);
  v = L * bz * rz * (1 - 4 * az) * (1. / 12); sym(K, 0, 4, v); sym(K, 6, 10, v);
  v = L * by * ry * (4 * ay - 1) * (1. / 12); sym(K, 0, 5, v); sym(K, 6, 11, v);
  v = -L * bz * rz * (2 * az + 1) * (1. / 12); sym(K, 0, 10, v); sym(K, 4, 6, v);
  v = L * by * ry * (2 * ay + 1) * (1. / 12); sym(K, 0, 11, v); sym(K, 5, 6, v);
  v = by2 * (70 * L2 * ay2 * r0 - 147 * L2 * ay * r0 + 78 * L2 * r0 + 252 * ry2) * (iL * (1. / 210)); sym(K, 1, 1, v); sym(K, 7, 7, v);
  sym(K, 1, 7, by2 * (35 * L2 * ay2 * r0 - 63 * L2 * ay * r0 + 27 * L2 * r0 - 252 * ry2) * (iL * (1. / 210)));
  v = bz2 * (70 * L2 * az2 * r0 - 147 * L2 * az * r0 + 78 * L2 * r0 + 252 * rz2) * (iL * (1. / 210)); sym(K, 2, 2, v); sym(K, 8, 8, v);
  sym(K, 2, 8, bz2 * (35 * L2 * az2 * r0 - 63 * L2 * az * r0 + 27 * L2 * r0 - 252 * rz2) * (iL * (1. / 210)));
  v = 6 * bb * ryz * (0.2 * iL); sym(K, 1, 2, v); sym(K, 7, 8, v); sym(K, 1, 8, -v); sym(K, 2, 7, -v);
  v = L * by * rz * (20 * ay - 21) * (1. / 60); sym(K, 1, 3, v); sym(K, 7, 9, v);
  v = L * by * rz * (10 * ay - 9) * (1. / 60); sym(K, 1, 9, v); sym(K, 3, 7, v);
  v = L * bz * ry * (21 - 20 * az) * (1. / 60); sym(K, 2, 3, v); sym(K, 8, 9, v);
  v = L * bz * ry * (9 - 10 * az) * (1. / 60); sym(K, 2, 9, v); sym(K, 3, 8, v);
  v = -bb * ryz * (5 * az + 1) * (1. / 10); sym(K, 1, 4, v); sym(K, 1, 10, v); sym(K, 4, 7, -v); sym(K, 7, 10, -v);
  v = bb * ryz * (5 * ay + 1) * (1. / 10); sym(K, 2, 5, v); sym(K, 2, 11, v); sym(K, 5, 8, -v); sym(K, 8, 11, -v);
  v = by2 * (35 * L2 * ay2 * r0 - 77 * L2 * ay * r0 + 44 * L2 * r0 + 420 * ay * ry2 + 84 * ry2) * (1. / 840); sym(K, 1, 5, v); sym(K, 7, 11, -v);
  v = by2 * (-35 * L2 * ay2 * r0 + 63 * L2 * ay * r0 - 26 * L2 * r0 + 420 * ay * ry2 + 84 * ry2) * (1. / 840); sym(K, 1, 11, v); sym(K, 5, 7, -v);
  v = bz2 * (-35 * L2 * az2 * r0 + 77 * L2 * az * r0 - 44 * L2 * r0 - 420 * az * rz2 - 84 * rz2) * (1. / 840); sym(K, 2, 4, v); sym(K, 8, 10, -v);
  v = bz2 * (35 * L2 * az2 * r0 - 63 * L2 * az * r0 + 26 * L2 * r0 - 420 * az * rz2 - 84 * rz2) * (1. / 840); sym(K, 2, 10, v); sym(K, 4, 8, -v);
  v = L * (ry2 + rz2) * (1. / 3); sym(K, 3, 3, v); sym(K, 9, 9, v); sym(K, 3, 9, L * (ry2 + rz2) * (1. / 6));
  v = L2 * bz * ry * (5 * az - 6) * (1. / 120); sym(K, 3, 4, v); sym(K, 9, 10, -v);
  v = L2 * by * rz * (5 * ay - 6) * (1. / 120); sym(K, 3, 5, v); sym(K, 9, 11, -v);
  v = L2 * bz * ry * (4 - 5 * az) * (1. / 120); sym(K, 3, 10, v); sym(K, 4, 9, -v);
  v = L2 * by * rz * (4 - 5 * ay) * (1. / 120); sym(K, 3, 11, v); sym(K, 5, 9, -v);
  v = L * bz2 * (7 * L2 * az2 * r0 - 14 * L2 * az * r0 + 8 * L2 * r0 + 280 * az2 * rz2 - 140 * az * rz2 + 112 * rz2) * (1. / 840); sym(K, 4, 4, v); sym(K, 10, 10, v);
  v = L * by2 * (7 * L2 * ay2 * r0 - 14 * L2 * ay * r0 + 8 * L2 * r0 + 280 * ay2 * ry2 - 140 * ay * ry2 + 112 * ry2) * (1. / 840); sym(K, 5, 5, v); sym(K, 11, 11, v);
  sym(K, 4, 10, L * bz2 * (-7 * L2 * az2 * r0 + 14 * L2 * az * r0 - 6 * L2 * r0 + 140 * az2 * rz2 + 140 * az * rz2 - 28 * rz2) * (1. / 840));
  sym(K, 5, 11, L * by2 * (-7 * L2 * ay2 * r0 + 14 * L2 * ay * r0 - 6 * L2 * r0 + 140 * ay2 * ry2 + 140 * ay * ry2 - 28 * ry2) * (1. / 840));
  v = L * bb * ryz * (-20 * ay * az + 5 * ay + 5 * az - 8) * (1. / 60); sym(K, 4, 5, v); sym(K, 10, 11, v);
  v = L * bb * ryz * (-10 * ay * az - 5 * ay - 5 * az + 2) * (1. / 60); sym(K, 4, 11, v); sym(K, 5, 10, v);
}

// linear Timoshenko beam, one-point reduced integration; beamlr.pyx:460-1186
template <class KS>
__device__ __forceinline__ void beamlr_Ke(const BeamP& p, double L, KS& K) {
  K.zero();
  const double iL = 1. / L, iL2 = iL * iL, iL3 = iL2 * iL;   // reciprocals: a double division costs ~15 instructions
  (void)iL2; (void)iL3;
  const double EA = p.E * p.A * iL, EAz = p.E * p.Az * iL, EAy = p.E * p.Ay * iL;
  const double GA = p.G * p.A, GAy = p.G * p.Ay, GAz = p.G * p.Az, GJ = p.G * p.J * iL;
  sym(K, 0, 0, EA); sym(K, 6, 6, EA); sym(K, 0, 6, -EA);
  sym(K, 0, 4, EAz); sym(K, 6, 10, EAz); sym(K, 0, 10, -EAz); sym(K, 4, 6, -EAz);
  sym(K, 0, 5, -EAy); sym(K, 6, 11, -EAy); sym(K, 0, 11, EAy); sym(K, 5, 6, EAy);
  sym(K, 1, 1, GA * iL); sym(K, 7, 7, GA * iL); sym(K, 2, 2, GA * iL); sym(K, 8, 8, GA * iL);
  sym(K, 1, 7, -GA * iL); sym(K, 2, 8, -GA * iL);
  sym(K, 1, 3, -GAz * iL); sym(K, 7, 9, -GAz * iL); sym(K, 1, 9, GAz * iL); sym(K, 3, 7, GAz * iL);
  sym(K, 1, 5, GA / 2); sym(K, 1, 11, GA / 2); sym(K, 5, 7, -GA / 2); sym(K, 7, 11, -GA / 2);
  sym(K, 2, 3, GAy * iL); sym(K, 8, 9, GAy * iL); sym(K, 2, 9, -GAy * iL); sym(K, 3, 8, -GAy * iL);
  sym(K, 2, 4, -GA / 2); sym(K, 2, 10, -GA / 2); sym(K, 4, 8, GA / 2); sym(K, 8, 10, GA / 2);
  sym(K, 3, 3, GJ); sym(K, 9, 9, GJ); sym(K, 3, 9, -GJ);
  sym(K, 3, 4, -GAy / 2); sym(K, 3, 10, -GAy / 2); sym(K, 4, 9, GAy / 2); sym(K, 9, 10, GAy / 2);
  sym(K, 3, 5, -GAz / 2); sym(K, 3, 11, -GAz / 2); sym(K, 5, 9, GAz / 2); sym(K, 9, 11, GAz / 2);
  sym(K, 4, 4, GA * L / 4 + p.E * p.Iyy * iL); sym(K, 10, 10, GA * L / 4 + p.E * p.Iyy * iL);
  sym(K, 4, 10, GA * L / 4 - p.E * p.Iyy * iL);
  sym(K, 5, 5, GA * L / 4 + p.E * p.Izz * iL); sym(K, 11, 11, GA * L / 4 + p.E * p.Izz * iL);
  sym(K, 5, 11, GA * L / 4 - p.E * p.Izz * iL);
  sym(K, 4, 5, -p.E * p.Iyz * iL); sym(K, 10, 11, -p.E * p.Iyz * iL);
  sym(K, 4, 11, p.E * p.Iyz * iL); sym(K, 5, 10, p.E * p.Iyz * iL);
}

// beamlr.pyx:1388-1462 (literal 0.333.. / 0.1666.. constants of the reference kept)
template <class KS>
__device__ __forceinline__ void beamlr_KGe(const BeamP& p, double L, const double* ue, KS& K) {
  K.zero();
  const double iL = 1. / L, iL2 = iL * iL, iL3 = iL2 * iL;   // reciprocals: a double division costs ~15 instructions
  (void)iL2; (void)iL3;
  const double N = p.A * p.E * (-ue[0] + ue[6]) * iL;
  const double t = 0.333333333333333 * L * N, s = 0.166666666666667 * L * N;
  sym(K, 4, 4, t); sym(K, 5, 5, t); sym(K, 10, 10, t); sym(K, 11, 11, t);
  sym(K, 4, 5, -t); sym(K, 10, 11, -t);
  sym(K, 4, 10, s); sym(K, 5, 11, s);
  sym(K, 4, 11, -s); sym(K, 5, 10, -s);
}

// beamlr.pyx:1518-2424; truss.pyx:894-1800 (same with the ry/rz inertia removed)
template <class KS>
__device__ __forceinline__ void beamlr_Me(const BeamP& p, double L, int mtype, bool truss, KS& K) {
  K.zero();
  const double iL = 1. / L, iL2 = iL * iL, iL3 = iL2 * iL;   // reciprocals: a double division costs ~15 instructions
  (void)iL2; (void)iL3;
  double mb[6][6];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) mb[i][j] = 0.;
  mb[0][0] = mb[1][1] = mb[2][2] = p.r0;
  mb[0][4] = mb[4][0] = p.rz;
  mb[0][5] = mb[5][0] = -p.ry;
  mb[1][3] = mb[3][1] = -p.rz;
  mb[2][3] = mb[3][2] = p.ry;
  mb[3][3] = p.ry2 + p.rz2;
  mb[4][4] = p.rz2;
  mb[5][5] = p.ry2;
  mb[4][5] = mb[5][4] = -p.ryz;
  if (truss) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j)
        if (i >= 4 || j >= 4) mb[i][j] = 0.;
  }
  if (mtype == 0) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 6; ++j) K.set(6 * a + i, 6 * b + j, ((a == b) ? L / 3 : L * (1. / 6)) * mb[i][j]);
  } else {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int i = 0; i < 6; ++i) K.set(6 * a + i, 6 * a + i, L / 2 * mb[i][i]);
  }
}

template <class KS>
__device__ __forceinline__ void truss_Ke(const BeamP& p, double L, KS& K) {
  K.zero();
  const double iL = 1. / L, iL2 = iL * iL, iL3 = iL2 * iL;   // reciprocals: a double division costs ~15 instructions
  (void)iL2; (void)iL3;
  const double EA = p.E * p.A * iL, GJ = p.G * p.J * iL;
  sym(K, 0, 0, EA); sym(K, 6, 6, EA); sym(K, 0, 6, -EA);
  sym(K, 3, 3, GJ); sym(K, 9, 9, GJ); sym(K, 3, 9, -GJ);
}

template <class KS>
__device__ __forceinline__ void spring_Ke(const double* k, KS& K) {
  K.zero();
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    K.set(i, i, k[i]);
    K.set(i + 6, i + 6, k[i]);
    K.set(i, i + 6, -k[i]);
    K.set(i + 6, i, -k[i]);
  }
}

// mask ids
constexpr int MK_FULL = 0, MK_D18 = 1, MK_RR = 2;
__host__ __device__ constexpr bool in_mask(int mk, int i, int j) {
  return mk == MK_FULL ? true : (mk == MK_D18 ? ((i < 3) == (j < 3)) : (i >= 3 && j >= 3));
}
__host__ __device__ constexpr int mask_rowcnt(int mk, int i) {
  int c = 0;
  for (int j = 0; j < 6; ++j) c += in_mask(mk, i, j) ? 1 : 0;
  return c;
}
__host__ __device__ constexpr int mask_colrank(int mk, int i, int j) {
  int c = 0;
  for (int q = 0; q < j; ++q) c += in_mask(mk, i, q) ? 1 : 0;
  return c;
}
__host__ __device__ constexpr int mask_rowoff(int mk, int i, int nblocks) {
  int c = 0;
  for (int q = 0; q < i; ++q) c += mask_rowcnt(mk, q) * nblocks;
  return c;
}

// Which local matrix a block is filled with (one closed form per element kind and matrix)
enum { FILL_KC0 = 0, FILL_KG = 1, FILL_M = 2 };
template <int KIND, int WHICH>
struct LineFill {
  const BeamP& p;
  double L;
  const double* ue;
  const double* kspr;
  int mtype;
  template <class KS>
  __device__ __forceinline__ void operator()(KS& K) const {
    if (WHICH == FILL_KC0) {
      if (KIND == PF3_BEAMC) beamc_Ke(p, L, K);
      if (KIND == PF3_BEAMLR) beamlr_Ke(p, L, K);
      if (KIND == PF3_TRUSS) truss_Ke(p, L, K);
      if (KIND == PF3_SPRING) spring_Ke(kspr, K);
    } else if (WHICH == FILL_KG) {
      if (KIND == PF3_BEAMC) beamc_KGe(p, L, ue, K);
      if (KIND == PF3_BEAMLR) beamlr_KGe(p, L, ue, K);
    } else {
      if (KIND == PF3_BEAMC) beamc_Me(p, L, mtype, K);
      if (KIND == PF3_BEAMLR || KIND == PF3_TRUSS) beamlr_Me(p, L, mtype, KIND == PF3_TRUSS, K);
    }
  }
};

// One 6x6 node-pair block: fill (registers only), rotate its masked 3x3 sub-blocks, drop the entries at their
// positions of node a's row slab (reference order: dof_i, node_j, dof_j restricted to the mask).
template <int MK, bool DIAG, int BA, int BB, class Fill>
__device__ __forceinline__ void stage_line_block(const Mat3& R, const Fill& fill, double* my) {
  if (DIAG && BA != BB) return;
  KBlock<BA, BB> K;
  fill(K);
  constexpr int nblocks = DIAG ? 1 : 2, bpos = DIAG ? 0 : BB;
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (MK == MK_D18 && s != t) continue;
      if (MK == MK_RR && !(s == 1 && t == 1)) continue;
      double l[3][3], o[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) l[i][j] = K.v[3 * s + i][3 * t + j];
      rot_block_full(R, l, o);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int ri = 3 * s + i, cj = 3 * t + j;
          my[mask_rowoff(MK, ri, nblocks) + bpos * mask_rowcnt(MK, ri) + mask_colrank(MK, ri, cj)] = o[i][j];
        }
    }
}

// Emit one node-row slab per flush.
template <int MK, bool DIAG, int SLAB, class Fill>
__device__ __forceinline__ void emit_line_matrix(const Mat3& R, const Fill& fill, double* stage, double* my, double* out,
                                                 int64_t e0, int nvalid, int esize, bool acc, int lane) {
  stage_line_block<MK, DIAG, 0, 0>(R, fill, my);
  stage_line_block<MK, DIAG, 0, 1>(R, fill, my);
  flush_chunk<SLAB>(stage, out, e0, nvalid, esize, 0, acc, lane);
  stage_line_block<MK, DIAG, 1, 0>(R, fill, my);
  stage_line_block<MK, DIAG, 1, 1>(R, fill, my);
  flush_chunk<SLAB>(stage, out, e0, nvalid, esize, SLAB, acc, lane);
}

// f[6a + i] = sum_b K_ab[i][:] . ue[6b : 6b+6], block by block
template <int BA, int BB, class Fill>
__device__ __forceinline__ void line_block_matvec(const Fill& fill, const double* ue, double* f) {
  KBlock<BA, BB> K;
  fill(K);
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) f[6 * BA + i] += K.v[i][j] * ue[6 * BB + j];
}

template <int KIND>
__global__ void __launch_bounds__(kThreads) line_eval_kernel(const EvalArgs A) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t e0 = (int64_t(blockIdx.x) * (blockDim.x >> 5) + warp) * 32;   // CTAs of 1 or kWarpsPerCta warps
  if (e0 >= A.ne) return;
  const int nvalid = int(min(int64_t(32), A.ne - e0));
  const int64_t e = e0 + min(lane, nvalid - 1);
  if (A.state == nullptr) conn_prefetch(A.conn, 2, e0, A.ne, lane);
  double* stage = smem + warp * 32 * kStageLd;
  double* my = stage + lane * kStageLd;

  Mat3 R;
  double xe[6] = {0, 0, 0, 0, 0, 0}, L = 0., ue[12];
  const bool have_u = (A.u != nullptr || A.state != nullptr);
  const bool from_state = A.state != nullptr;
  const bool need_x = KIND != PF3_SPRING && (!from_state || (A.state_flags & PF3_STATE_REFRESH_XE));
  const bool need_u = have_u && (!from_state || (A.state_flags & PF3_STATE_REFRESH_UE));
  const int64_t n0 = A.conn[2 * e], n1 = A.conn[2 * e + 1];
  double xh[3], yh[3], zh[3], P0[3] = {0, 0, 0}, P1[3] = {0, 0, 0};
  if (need_x)
    for (int i = 0; i < 3; ++i) {
      P0[i] = A.x[3 * n0 + i];
      P1[i] = A.x[3 * n1 + i];
    }
  double U[2][6];   // gathered with the coordinates, not after the frame: one exposed memory round trip less
  if (need_u)
    for (int i = 0; i < 6; ++i) {
      U[0][i] = A.u[6 * n0 + i];
      U[1][i] = A.u[6 * n1 + i];
    }
  if (from_state) {
    const double* s = A.state + e * PF3_STATE_STRIDE;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R.a[i][j] = s[3 * i + j];
    L = s[13];
    for (int i = 0; i < 6; ++i) xe[i] = s[14 + i];
    for (int i = 0; i < 12; ++i) ue[i] = s[26 + i];
    for (int i = 0; i < 3; ++i) {
      xh[i] = R.a[i][0];
      yh[i] = R.a[i][1];
      zh[i] = R.a[i][2];
    }
  } else {
    double v[3];
    const double* ev = A.evec ? A.evec + e * int64_t(A.evec_stride) : nullptr;
    if (KIND == PF3_SPRING) {
      // axis given directly (spring.pyx:147-206)
      for (int i = 0; i < 3; ++i) {
        xh[i] = ev[i];
        v[i] = ev[3 + i];
      }
      normalize3(xh);
    } else {
      for (int i = 0; i < 3; ++i) xh[i] = P1[i] - P0[i];
      normalize3(xh);
      if (KIND == PF3_TRUSS) {  // arbitrary off-axis vector: cyclic shift (truss.pyx:243-245)
        v[0] = xh[1];
        v[1] = xh[2];
        v[2] = xh[0];
      } else {
        for (int i = 0; i < 3; ++i) v[i] = ev[i];
      }
    }
    cross3(xh, v, zh);   // z = x X vxy   (beamc.pyx:207-214)
    normalize3(zh);
    cross3(zh, xh, yh);  // y = z X x
    normalize3(yh);
    for (int i = 0; i < 3; ++i) {
      R.a[i][0] = xh[i];
      R.a[i][1] = yh[i];
      R.a[i][2] = zh[i];
    }
  }
  if (need_x) {
    for (int a = 0; a < 2; ++a) {
      const double* P = a ? P1 : P0;
      xe[3 * a + 0] = xh[0] * P[0] + xh[1] * P[1] + xh[2] * P[2];
      xe[3 * a + 1] = yh[0] * P[0] + yh[1] * P[1] + yh[2] * P[2];
      xe[3 * a + 2] = zh[0] * P[0] + zh[1] * P[1] + zh[2] * P[2];
    }
    // update_length (beamc.pyx:336-349) with the difference formed in global coordinates BEFORE the rotation: the
    // reference subtracts the rotated absolute positions and loses |x| / L * eps (6e-12 on the 100 k-element arc of
    // config 2); see ShellGeom in shell.cuh
    const double q[3] = {P1[0] - P0[0], P1[1] - P0[1], P1[2] - P0[2]};
    const double dx = xh[0] * q[0] + xh[1] * q[1] + xh[2] * q[2];
    const double dy = yh[0] * q[0] + yh[1] * q[1] + yh[2] * q[2];
    const double dz = zh[0] * q[0] + zh[1] * q[1] + zh[2] * q[2];
    L = sqrt(dx * dx + dy * dy + dz * dz);
  }
  if (need_u) {
    for (int a = 0; a < 2; ++a)
      for (int t = 0; t < 2; ++t) {
        const double* ug = U[a] + 3 * t;
        ue[6 * a + 3 * t + 0] = xh[0] * ug[0] + xh[1] * ug[1] + xh[2] * ug[2];
        ue[6 * a + 3 * t + 1] = yh[0] * ug[0] + yh[1] * ug[1] + yh[2] * ug[2];
        ue[6 * a + 3 * t + 2] = zh[0] * ug[0] + zh[1] * ug[1] + zh[2] * ug[2];
      }
  }
  if (A.state_out != nullptr) {
    if (lane < nvalid) {
      double* s = A.state_out + e * PF3_STATE_STRIDE;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) s[3 * i + j] = R.a[i][j];
      s[9] = 1.;
      s[10] = 0.;
      s[11] = 0.;
      s[12] = 1.;
      s[13] = L;
      for (int i = 0; i < 12; ++i) s[14 + i] = (i < 6) ? xe[i] : 0.;
      for (int i = 0; i < 24; ++i) s[26 + i] = (i < 12 && have_u) ? ue[i] : 0.;
    }
    if (A.what == 0) return;
  }

  BeamP p;
  double kspr[6];
  if (KIND == PF3_SPRING) {
    for (int i = 0; i < 6; ++i) kspr[i] = A.eparam[e * PF3_EPARAM_STRIDE + i];
  } else {
    const double* q = A.props + prop_index(A, e) * PF3_BEAMPROP_STRIDE;
    p.A = q[0]; p.E = q[1]; p.G = q[2]; p.Iyy = q[3]; p.Izz = q[4]; p.Iyz = q[5]; p.J = q[6]; p.Ay = q[7];
    p.Az = q[8]; p.r0 = q[9]; p.ry = q[10]; p.rz = q[11]; p.ry2 = q[12]; p.rz2 = q[13]; p.ryz = q[14];
  }

  if (A.what & (PF3_KC0 | PF3_FINT)) {
    const LineFill<KIND, FILL_KC0> fill{p, L, ue, kspr, A.mtype};
    if (A.what & PF3_KC0) {
      double* out = A.kc0v + A.kc0_k0;
      if (KIND == PF3_BEAMC || KIND == PF3_BEAMLR)
        emit_line_matrix<MK_FULL, false, 72>(R, fill, stage, my, out, e0, nvalid, 144, A.acc_kc0 != 0, lane);
      else
        emit_line_matrix<MK_D18, false, 36>(R, fill, stage, my, out, e0, nvalid, 72, A.acc_kc0 != 0, lane);
    }
    if (A.what & PF3_FINT) {
      double f[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) f[i] = 0.;
      line_block_matvec<0, 0>(fill, ue, f);
      line_block_matvec<0, 1>(fill, ue, f);
      line_block_matvec<1, 0>(fill, ue, f);
      line_block_matvec<1, 1>(fill, ue, f);
      if (A.finte != nullptr) {
#pragma unroll
        for (int i = 0; i < 12; ++i) my[i] = f[i];
        flush_chunk<12>(stage, A.finte, e0, nvalid, 12, 0, false, lane);
      }
      if (A.fe != nullptr) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int i = 0; i < 3; ++i)
            my[3 * t + i] = R.a[i][0] * f[3 * t] + R.a[i][1] * f[3 * t + 1] + R.a[i][2] * f[3 * t + 2];
        flush_chunk<12>(stage, A.fe, e0, nvalid, 12, 0, false, lane);
      }
    }
  }
  if ((A.what & PF3_KG) && (KIND == PF3_BEAMC || KIND == PF3_BEAMLR)) {
    double* out = A.kgv + A.kg_k0;
    const LineFill<KIND, FILL_KG> fill{p, L, ue, kspr, A.mtype};
    if (KIND == PF3_BEAMC)
      emit_line_matrix<MK_FULL, false, 72>(R, fill, stage, my, out, e0, nvalid, 144, A.acc_kg != 0, lane);
    else
      emit_line_matrix<MK_RR, false, 18>(R, fill, stage, my, out, e0, nvalid, 36, A.acc_kg != 0, lane);
  }
  if ((A.what & PF3_M) && KIND != PF3_SPRING) {
    double* out = A.mv + A.m_k0;
    const LineFill<KIND, FILL_M> fill{p, L, ue, kspr, A.mtype};
    if (A.mtype == 0)
      emit_line_matrix<MK_FULL, false, 72>(R, fill, stage, my, out, e0, nvalid, 144, A.acc_m != 0, lane);
    else  // lumped: two diagonal node blocks only, 36 of 144 entries written (beamc.pyx:2970)
      emit_line_matrix<MK_D18, true, 18>(R, fill, stage, my, out, e0, nvalid, 144, A.acc_m != 0, lane);
  }
}

}  // namespace

cudaError_t launch_line(int kind, const EvalArgs& A, cudaStream_t st) {
  if (A.ne <= 0) return cudaSuccess;
  // small batches (BASELINE config 2: 100 k beams = 3 125 warps for 148 SMs) are latency-bound: one-warp CTAs spread
  // them over every SM instead of filling a fifth of the machine with 4-warp CTAs
  const int warps = A.ne < int64_t(148) * 8 * 32 * kWarpsPerCta ? 1 : kWarpsPerCta;
  const int64_t per_cta = 32 * warps;
  const unsigned grid = unsigned((A.ne + per_cta - 1) / per_cta);
  const unsigned nthreads = unsigned(32 * warps);
  const size_t smem = kStageBytes / kWarpsPerCta * warps;
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(line_eval_kernel<PF3_BEAMC>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kStageBytes));
    cudaFuncSetAttribute(line_eval_kernel<PF3_BEAMLR>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kStageBytes));
    cudaFuncSetAttribute(line_eval_kernel<PF3_TRUSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kStageBytes));
    cudaFuncSetAttribute(line_eval_kernel<PF3_SPRING>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kStageBytes));
  }
  switch (kind) {
    case PF3_BEAMC: line_eval_kernel<PF3_BEAMC><<<grid, nthreads, smem, st>>>(A); break;
    case PF3_BEAMLR: line_eval_kernel<PF3_BEAMLR><<<grid, nthreads, smem, st>>>(A); break;
    case PF3_TRUSS: line_eval_kernel<PF3_TRUSS><<<grid, nthreads, smem, st>>>(A); break;
    case PF3_SPRING: line_eval_kernel<PF3_SPRING><<<grid, nthreads, smem, st>>>(A); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace pf3
