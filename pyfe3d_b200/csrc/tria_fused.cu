// Fused, node-centric Tria3R evaluation + CSR assembly: the triangle twin of quad_fused.cu.
//
// One pass writes the reference's COO value arrays KC0v / KGv / Mv (Tria3R.update_KC0 tria3r.pyx:950, update_KG
// :3020, update_KG_given_stress :3576, update_M :4063) AND the CSR values scipy's coo_matrix(...).tocsr() would give
// (tests/test_tria3r_natural_freq_distorted.py:130-131), without re-reading the COO arrays.
//
//  K1 tria_record_kernel  one THREAD per element: frame (update_rotation_matrix tria3r.pyx:294), local coordinates
//     (update_probe_xe :480), area, the constant gradients (:2116-2121), the shear-locking factor (:1109-1122), the
//     material-axis rotation of A/B/D (:1040-1092) and, for KG, the membrane resultants (:3399 ff.) -> a 24- or
//     44-double record.
//  K2 tria_fused_kernel   one WARP per node (one-warp CTAs, as many CTAs as nodes: the CTA scheduler keeps the nodes
//     in flight a narrow window of the mesh).  A triangle node has ~6 incident elements x 3 node-pair blocks: lane
//     3k+b evaluates block (a_k, b) of incidence k (up to 10 incidences per round).  Blocks are staged in the COO
//     slab layout, every (element, node) slab leaves as one TMA bulk copy (KG slabs are 216 B, not a 16-byte
//     multiple: plain coalesced stores), and the node's CSR rows are reduced from the staged slabs in the fixed
//     order k = 0..9 (deterministic, no atomics).
#include "shell.cuh"

namespace pf3 {

namespace {

constexpr int kTRecPlain = 24;   // record doubles without / with rotated A,B,D
constexpr int kTRecRot = 44;
constexpr int kTInc = 10;        // incidences per round (TriaRec::inc)
constexpr double kThird = 0.333333333333333333333333333;

// ------------------------------------------------------------------------------------------ K1
__global__ void __launch_bounds__(128) tria_record_kernel(const EvalArgs A, double* __restrict__ rec, int stride) {
  extern __shared__ double k1_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t e0 = (int64_t(blockIdx.x) * (blockDim.x >> 5) + warp) * 32;
  if (e0 >= A.ne) return;
  const int nvalid = int(min(int64_t(32), A.ne - e0));
  const int64_t e = e0 + min(lane, nvalid - 1);
  conn_prefetch(A.conn, 3, e0, A.ne, lane);
  const int ld = stride + 1;
  double* stage = k1_smem + warp * 32 * ld;
  const bool kg_u = (A.what & PF3_KG) != 0;
  double ue[18];
  ShellGeom<3> g;
  shell_geom<3, true>(A, e, g, kg_u ? ue : nullptr);
  ShellCoef c;
  shell_coef<3>(A, e, g, c);
  double K6ROT = 100., alpha = 0.7;
  if (A.eparam != nullptr) {
    K6ROT = A.eparam[e * PF3_EPARAM_STRIDE + 0];
    alpha = A.eparam[e * PF3_EPARAM_STRIDE + 1];
  }
  const double dJ = 2. * g.area;  // tria3r.pyx:1017
  const double i2a = 1. / (2. * g.area);
  const double Nx[3] = {(g.Y[1] - g.Y[2]) * i2a, (-g.Y[0] + g.Y[2]) * i2a, (g.Y[0] - g.Y[1]) * i2a};
  const double Ny[3] = {(-g.X[1] + g.X[2]) * i2a, (g.X[0] - g.X[2]) * i2a, (-g.X[0] + g.X[1]) * i2a};
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
  const double w = dJ * 0.5;
  double* r = stage + lane * ld;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r[3 * i + j] = g.R.a[i][j];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    r[9 + a] = Nx[a];
    r[12 + a] = Ny[a];
  }
  r[15] = w;
  r[16] = 1e-6 * K6ROT * c.A[5] * w;   // drilling penalty x total weight (tria3r.pyx:2261 ff.)
  r[17] = c.E44;
  r[18] = c.E45;
  r[19] = c.E55;
  r[20] = dJ;
  r[21] = 0.;
  r[22] = 0.;
  r[23] = 0.;
  if (kg_u) {
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
    r[21] = c.A[0] * exx + c.A[1] * eyy + c.A[2] * gxy + c.B[0] * kxx + c.B[1] * kyy + c.B[2] * kxy;
    r[22] = c.A[1] * exx + c.A[3] * eyy + c.A[4] * gxy + c.B[1] * kxx + c.B[3] * kyy + c.B[4] * kxy;
    r[23] = c.A[2] * exx + c.A[4] * eyy + c.A[5] * gxy + c.B[2] * kxx + c.B[4] * kyy + c.B[5] * kxy;
  }
  if (stride == kTRecRot) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      r[24 + i] = c.A[i];
      r[30 + i] = c.B[i];
      r[36 + i] = c.D[i];
    }
    r[42] = 0.;
    r[43] = 0.;
  }
  __syncwarp();
  double* out = rec + e0 * stride;
  const int total = nvalid * stride;
  for (int idx = lane; idx < total; idx += 32) out[idx] = stage[(idx / stride) * ld + idx % stride];
}

// ------------------------------------------------------------------------------------------ K2
template <int NR, int CNT>
struct TSlab {
  static constexpr int kSlab = NR * 3 * CNT;                 // doubles per (element, node) COO slab: 108 / 90 / 54 / 27
  // even stride (16-B aligned slabs for the bulk copy); 114 = 2 mod 16 makes the 16-B staging stores of KC0
  // conflict-free over the 10 incidences of a warp
  static constexpr int kLd = (CNT == 6) ? 114 : kSlab + 2 - (kSlab & 1);
  static constexpr bool kBulk = (kSlab * 8) % 16 == 0;       // KG slabs (216 B) cannot be bulk copies
};

__device__ __forceinline__ uint32_t smem_u32t(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void tstage_wait() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

// Lanes 3k+b have written their block into slab k.  Ship the slabs to the COO array and reduce the node's (up to)
// kTInc incidences into its CSR rows in the fixed order k = 0..v-1.
template <int NR, int CNT>
__device__ __forceinline__ void tria_emit(const double* st, const TriaRec* nr, double* __restrict__ coo,
                                          int64_t slab_base, bool act, int b, double* __restrict__ csr,
                                          int64_t csr_base, int nb, int v, bool first_round, int lane, uint64_t pol) {
  constexpr int kSlab = TSlab<NR, CNT>::kSlab, kLd = TSlab<NR, CNT>::kLd;
  if constexpr (TSlab<NR, CNT>::kBulk) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (coo != nullptr && act && b == 0) {
      const uint32_t src = smem_u32t(st + (lane / 3) * kLd);
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
  } else {
    __syncwarp();
    // the whole warp streams one slab per step: 27 contiguous doubles
    for (int k2 = 0; k2 < v; ++k2) {
      const int64_t base = __shfl_sync(0xffffffffu, slab_base, 3 * k2);
      const bool on = __shfl_sync(0xffffffffu, int(act), 3 * k2) != 0;
      if (coo != nullptr && on && lane < kSlab) stg_stream(coo + base + lane, st[k2 * kLd + lane], pol);
    }
  }
  if (nb > 0 && csr != nullptr) {
    const int w = nb * CNT;
    double* out = csr + csr_base;
    if constexpr (CNT % 2 == 0) {
#pragma unroll 1
      for (int y = lane; 2 * y < w; y += 32) {
        const int x = 2 * y, s = x / CNT, rr = x - s * CNT;
        const unsigned gm = nr->gmap[s];
        double2 sum[NR];
#pragma unroll
        for (int d = 0; d < NR; ++d) sum[d] = make_double2(0., 0.);
        for (int k2 = 0; k2 < v; ++k2) {
          const int nib = (gm >> (2 * k2)) & 3;
          if (nib != 3) {
            const double* p = st + k2 * kLd + nib * CNT + rr;
#pragma unroll
            for (int d = 0; d < NR; ++d) {
              const double2 t = *reinterpret_cast<const double2*>(p + d * 3 * CNT);
              sum[d].x += t.x;
              sum[d].y += t.y;
            }
          }
        }
#pragma unroll
        for (int d = 0; d < NR; ++d) {
          double2* o = reinterpret_cast<double2*>(out + d * w + x);
          if (!first_round) {
            const double2 t = *o;
            sum[d].x += t.x;
            sum[d].y += t.y;
          }
          stg_stream(o, sum[d], pol);
        }
      }
    } else {
#pragma unroll 1
      for (int x = lane; x < w; x += 32) {
        const int s = x / CNT, rr = x - s * CNT;
        const unsigned gm = nr->gmap[s];
        double sum[NR];
#pragma unroll
        for (int d = 0; d < NR; ++d) sum[d] = 0.;
        for (int k2 = 0; k2 < v; ++k2) {
          const int nib = (gm >> (2 * k2)) & 3;
          if (nib != 3) {
            const double* p = st + k2 * kLd + nib * CNT + rr;
#pragma unroll
            for (int d = 0; d < NR; ++d) sum[d] += p[d * 3 * CNT];
          }
        }
#pragma unroll
        for (int d = 0; d < NR; ++d) {
          double* o = out + d * w + x;
          if (!first_round) sum[d] += *o;
          stg_stream(o, sum[d], pol);
        }
      }
    }
  }
  __syncwarp();
}

constexpr int kTStage = kTInc * TSlab<6, 6>::kLd;   // 1140 doubles: the largest matrix (KC0)
#ifndef PF3_TFUSED_WARPS
#define PF3_TFUSED_WARPS 1
#define PF3_TFUSED_CTAS 15   // 14.7 kB of shared memory per one-warp CTA: 15 fit in 227 kB
#endif
#ifndef PF3_TFUSED_CHUNK
#define PF3_TFUSED_CHUNK 4    // consecutive nodes per warp: the records of the next node arrive while this one is computed
#endif
constexpr int kTWarps = PF3_TFUSED_WARPS;
constexpr int kTChunk = PF3_TFUSED_CHUNK;
constexpr int kTRing = 3;
// per warp: slab staging | ring of 3 node records (128 B each) | double-buffered element records of 10 incidences
__host__ __device__ constexpr int twarp_smem_doubles(int rstride) {
  return kTStage + kTRing * 16 + 2 * kTInc * (rstride + 2);
}

// One warp = kTChunk consecutive nodes, one CTA = one warp, as many CTAs as there is work (the CTA scheduler keeps
// the nodes in flight a narrow window of the mesh).  Within a warp the node record of item j+2 and the element
// records of item j+1 are in flight (cp.async) while item j is evaluated.
__global__ void __launch_bounds__(32 * kTWarps, PF3_TFUSED_CTAS) tria_fused_kernel(const FusedArgs F,
                                                                               const double* __restrict__ rec,
                                                                               int rstride) {
  extern __shared__ __align__(16) double smem[];
  const EvalArgs& A = F.A;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int eld = rstride + 2;
  double* st = smem + warp * twarp_smem_doubles(rstride);
  TriaRec* ring = reinterpret_cast<TriaRec*>(st + kTStage);
  double* erec = st + kTStage + kTRing * 16;
  const int k = lane / 3, b = lane - 3 * k;     // lanes 30, 31: k = 10, never active
  const uint64_t pol = l2_evict_first_policy();
  const int64_t n0 = (int64_t(blockIdx.x) * kTWarps + warp) * kTChunk;
  if (n0 >= F.nown) return;
  const int rmax = F.rmax;
  const int nitems = int(min(int64_t(kTChunk), F.nown - n0)) * rmax;   // items j = (node n0 + j / rmax, round j % rmax)
  // bulk L2 prefetch of the node records and first-use element records kPfAhead chunks ahead (see quad_fused.cu)
  if (F.pftab != nullptr && F.triarec != nullptr && lane == 0 && (n0 % (2 * kPfChunk)) < kTChunk) {
    const int64_t c = n0 / (2 * kPfChunk) + kPfAhead;
    if (c < F.pf_nchunks) {
      constexpr unsigned kPiece = 32768;
      const int64_t nn0 = c * kPfChunk * 2;
      const int64_t nn = min(int64_t(2 * kPfChunk), F.nown - nn0);
      const char* q = reinterpret_cast<const char*>(F.triarec + nn0 * rmax);
      for (int64_t left = nn * rmax * int64_t(sizeof(TriaRec)); left > 0; left -= kPiece, q += kPiece)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q), "r"(unsigned(min(left, int64_t(kPiece)))) : "memory");
      for (int run = 0; run < kPfRuns; ++run) {
        const int2 t = F.pftab[c * kPfRuns + run];
        q = reinterpret_cast<const char*>(rec + int64_t(t.x) * rstride);
        for (int64_t left = int64_t(t.y) * rstride * 8; left > 0; left -= kPiece, q += kPiece)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q), "r"(unsigned(min(left, int64_t(kPiece)))) : "memory");
      }
    }
  }

  auto rec_fetch = [&](int j) {
    if (j < nitems && F.triarec != nullptr) {
      if (lane < 8) {
        const int64_t rec_index = rmax == 1 ? n0 + j : (n0 + j / rmax) * rmax + j % rmax;   // no runtime division when
        const char* src = reinterpret_cast<const char*>(F.triarec + rec_index) + 16 * lane;    // every node has one round
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32t(reinterpret_cast<char*>(ring + j % kTRing) + 16 * lane)),
                     "l"(src)
                     : "memory");
      }
    } else if (j < nitems) {
      // element mode (COO only, no plan): "node" n stands for elements 3n, 3n+1, 3n+2, its 9 incidences are their
      // own row slabs; nothing is assembled (nb = 0)
      TriaRec* t = ring + j % kTRing;
      const int64_t n = n0 + j;
      if (lane < kTInc) {
        const int64_t el = 3 * n + lane / 3;
        t->inc[lane] = (lane < 9 && el < A.ne) ? int32_t(el * 9 + (lane % 3) * 3) : -1;
      }
      if (lane == 10) {
        t->b0 = 0;
        t->v = 9;
        t->nb = 0;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto erec_fetch = [&](int j) {
    if (j < nitems && k < kTInc) {
      const int p0 = (ring + j % kTRing)->inc[k];
      if (p0 >= 0) {
        const char* src = reinterpret_cast<const char*>(rec + int64_t(p0 / 9) * rstride);
        char* dst = reinterpret_cast<char*>(erec + ((j & 1) * kTInc + k) * eld);
        for (int c = b * 16; c < rstride * 8; c += 48)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32t(dst + c)), "l"(src + c) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  rec_fetch(0);
  rec_fetch(1);
  asm volatile("cp.async.wait_group 1;" ::: "memory");
  __syncwarp();
  erec_fetch(0);

  for (int j = 0; j < nitems; ++j) {
    rec_fetch(j + 2);
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // node record j+1 and element records j have landed
    __syncwarp();
    erec_fetch(j + 1);
    const TriaRec* nr = ring + j % kTRing;
    const int pair0 = (k < kTInc) ? nr->inc[k] : -1;
    const bool act = pair0 >= 0;
    if (__ballot_sync(0xffffffffu, act) == 0u) continue;
    const int64_t b0 = nr->b0;
    const int nb = nr->nb, v = nr->v;
    const int64_t e = act ? (pair0 / 9) : 0;
    const int a = act ? ((pair0 - int(e) * 9) / 3) : 0;
    const bool first = rmax == 1 || (j % rmax == 0);
    const int kk = (k < kTInc) ? k : 0;   // idle lanes compute on slab 0's record but never stage or store

    const double* re = erec + ((j & 1) * kTInc + kk) * eld;
    Mat3 R;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int jj = 0; jj < 3; ++jj) R.a[i][jj] = re[3 * i + jj];
    const double Nxa = re[9 + a], Nya = re[12 + a], Nxb = re[9 + b], Nyb = re[12 + b];
    const double w = re[15], kd = re[16], E44 = re[17], E45 = re[18], E55 = re[19], dJ = re[20];
    const double* prow = A.props + prop_index(A, e) * PF3_SHELLPROP_STRIDE;
    const double* abd = (rstride == kTRecRot) ? re + 24 : prow;
    const bool stager = act && k < kTInc;

    // ---------------- KG : Ge_ab * z z^T on the translations (tria3r.pyx:3399 ff.)
    if (A.what & (PF3_KG | PF3_KG_STRESS)) {
      double Nxx = A.Nxx, Nyy = A.Nyy, Nxy = A.Nxy;
      if (!(A.what & PF3_KG_STRESS)) {
        Nxx = re[21];
        Nyy = re[22];
        Nxy = re[23];
      }
      const double px = w * (Nxa * Nxx + Nya * Nxy), py = w * (Nxa * Nxy + Nya * Nyy);
      const double ge = Nxb * px + Nyb * py;
      tstage_wait();
      if (stager) {
        double* sl = st + k * TSlab<3, 3>::kLd + b * 3;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) sl[i * 9 + jj] = (R.a[i][2] * R.a[jj][2]) * ge;
      }
      tria_emit<3, 3>(st, nr, A.kgv ? A.kgv + A.kg_k0 : nullptr, e * 81 + a * 27, act, b, F.csr_kg, b0 * 9, nb, v,
                      first, lane, pol);
    }

    // ---------------- M : H_ab * (T6 m_l T6^T) (tria3r.pyx:4063 ff.)
    if (A.what & PF3_M) {
      double hd, ho;
      // (multiplications by the reciprocals: a double division is a ~15-instruction sequence per lane; the last-bit
      // difference to detJ / 12 etc. is far inside the 1e-12 tolerance)
      if (A.mtype == 0) {
        hd = dJ * (1. / 12.);
        ho = dJ * (1. / 24.);
      } else if (A.mtype == 1) {
        hd = ho = dJ * (1. / 18.);
      } else {
        hd = dJ * (1. / 6.);
        ho = 0.;
      }
      const double h = (a == b) ? hd : ho;
      NodalInertia Mi;
      nodal_inertia(R, prow[24], prow[25], prow[26], Mi);
      double* coo = A.mv ? A.mv + A.m_k0 : nullptr;
      tstage_wait();
      if (A.mtype != 2) {
        if (stager) {
          double* sl = st + k * TSlab<6, 5>::kLd + b * 5;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            sl[i * 15 + 0] = h * Mi.tt[i][0];
            sl[i * 15 + 1] = h * Mi.tt[i][1];
            sl[i * 15 + 2] = h * Mi.tt[i][2];
            sl[i * 15 + 3] = h * Mi.tr[i][(i == 0) ? 1 : 0];
            sl[i * 15 + 4] = h * Mi.tr[i][(i == 2) ? 1 : 2];
            sl[(3 + i) * 15 + 0] = h * Mi.tr[(i == 0) ? 1 : 0][i];
            sl[(3 + i) * 15 + 1] = h * Mi.tr[(i == 2) ? 1 : 2][i];
            sl[(3 + i) * 15 + 2] = h * Mi.rr[i][0];
            sl[(3 + i) * 15 + 3] = h * Mi.rr[i][1];
            sl[(3 + i) * 15 + 4] = h * Mi.rr[i][2];
          }
        }
        tria_emit<6, 5>(st, nr, coo, e * 270 + a * 90, act, b, F.csr_m, b0 * 30, nb, v, first, lane, pol);
      } else {
        if (stager) {
          double* sl = st + k * TSlab<6, 3>::kLd + b * 3;
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
              sl[i * 9 + jj] = h * Mi.tt[i][jj];
              sl[(3 + i) * 9 + jj] = h * Mi.rr[i][jj];
            }
        }
        tria_emit<6, 3>(st, nr, coo, e * 270 + a * 54, act, b, F.csr_m, b0 * 18, nb, v, first, lane, pol);
      }
    }

    // ---------------- KC0 : the 6x6 block (a, b) (tria3r.pyx:2124-2325 and the rotate/write section :2325-2973)
    if (A.what & PF3_KC0) {
      double cA[6], cB[6], cD[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        cA[i] = abd[i];
        cB[i] = abd[6 + i];
        cD[i] = abd[12 + i];
      }
      const double gxx = w * Nxa * Nxb, gxy = w * Nxa * Nyb, gyx = w * Nya * Nxb, gyy = w * Nya * Nyb;
      const double tSa = w * (E44 * Nya + E45 * Nxa), sSa = w * (E45 * Nya + E55 * Nxa);
      const double tSb = w * (E44 * Nyb + E45 * Nxb), sSb = w * (E45 * Nyb + E55 * Nxb);
      const double c44 = w * E44 * kThird * kThird, c45 = w * E45 * kThird * kThird, c55 = w * E55 * kThird * kThird;
      double o1[3][3], o2[3][3];
      {
        // in-plane part of the drilling penalty survives in update_KC0; the (u,v)-rz and rz_a-rz_b couplings are
        // never written (kept for parity)
        const double uu = f_pp(cA, gxx, gxy, gyx, gyy) + 0.25 * kd * Nya * Nyb;
        const double uv = f_pq(cA, gxx, gxy, gyx, gyy) - 0.25 * kd * Nya * Nxb;
        const double vu = f_qp(cA, gxx, gxy, gyx, gyy) - 0.25 * kd * Nxa * Nyb;
        const double vv = f_qq(cA, gxx, gxy, gyx, gyy) + 0.25 * kd * Nxa * Nxb;
        const double ww = tSa * Nyb + sSa * Nxb;
        rot_block_diag5(R, uu, uv, vu, vv, ww, o1);
      }
      rot_block_8(R, -f_pq(cB, gxx, gxy, gyx, gyy), f_pp(cB, gxx, gxy, gyx, gyy), 0., -f_qq(cB, gxx, gxy, gyx, gyy),
                  f_qp(cB, gxx, gxy, gyx, gyy), 0., -kThird * tSa, kThird * sSa, o2);
      tstage_wait();
      double2* sl2 = reinterpret_cast<double2*>(st + kk * TSlab<6, 6>::kLd + b * 6);
      if (stager) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {   // rows u v w of node a: 6 columns of node b, 18 doubles per COO row
          sl2[i * 9 + 0] = make_double2(o1[i][0], o1[i][1]);
          sl2[i * 9 + 1] = make_double2(o1[i][2], o2[i][0]);
          sl2[i * 9 + 2] = make_double2(o2[i][1], o2[i][2]);
        }
      }
      rot_block_8(R, -f_qp(cB, gxx, gxy, gyx, gyy), -f_qq(cB, gxx, gxy, gyx, gyy), -kThird * tSb,
                  f_pp(cB, gxx, gxy, gyx, gyy), f_pq(cB, gxx, gxy, gyx, gyy), kThird * sSb, 0., 0., o1);
      {
        const double rzrz = (a == b) ? kd * (1. / 6.) : 0.;
        rot_block_diag5(R, f_qq(cD, gxx, gxy, gyx, gyy) + c44, -f_qp(cD, gxx, gxy, gyx, gyy) - c45,
                        -f_pq(cD, gxx, gxy, gyx, gyy) - c45, f_pp(cD, gxx, gxy, gyx, gyy) + c55, rzrz, o2);
      }
      if (stager) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          sl2[(3 + i) * 9 + 0] = make_double2(o1[i][0], o1[i][1]);
          sl2[(3 + i) * 9 + 1] = make_double2(o1[i][2], o2[i][0]);
          sl2[(3 + i) * 9 + 2] = make_double2(o2[i][1], o2[i][2]);
        }
      }
      tria_emit<6, 6>(st, nr, A.kc0v ? A.kc0v + A.kc0_k0 : nullptr, e * 324 + a * 108, act, b, F.csr_kc0, b0 * 36, nb, v,
                      first, lane, pol);
    }
  }
  tstage_wait();   // the bulk copies read this CTA's shared memory: they must have done so before it retires
}

}  // namespace

int tria_fused_record_stride(const EvalArgs& A) { return A.evec != nullptr ? kTRecRot : kTRecPlain; }
int tria_fused_max_slots() { return 16; }
int tria_fused_incidences() { return kTInc; }

// rec: device scratch of ne * tria_fused_record_stride doubles
cudaError_t launch_tria_fused(const FusedArgs& F, double* rec, cudaStream_t st, int64_t* launches) {
  if (F.nown <= 0 || F.A.ne <= 0) return cudaSuccess;
  const int stride = tria_fused_record_stride(F.A);
  const unsigned g1 = unsigned((F.A.ne + 127) / 128);
  const size_t smem1 = size_t(4) * 32 * (stride + 1) * sizeof(double);
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(tria_record_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         int(size_t(4) * 32 * (kTRecRot + 1) * sizeof(double)));
    cudaFuncSetAttribute(tria_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         int(size_t(kTWarps) * twarp_smem_doubles(kTRecRot) * sizeof(double)));
  }
  tria_record_kernel<<<g1, 128, smem1, st>>>(F.A, rec, stride);
  ++*launches;
  cudaError_t e1 = cudaGetLastError();
  if (e1 != cudaSuccess) return e1;
  const int64_t want = (F.nown + int64_t(kTWarps) * kTChunk - 1) / (int64_t(kTWarps) * kTChunk);
  if (want > int64_t(0x7fffffff)) return cudaErrorInvalidConfiguration;
  const size_t smem = size_t(kTWarps) * twarp_smem_doubles(stride) * sizeof(double);
  tria_fused_kernel<<<unsigned(want), 32 * kTWarps, smem, st>>>(F, rec, stride);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace pf3
