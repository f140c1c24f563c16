// Host-side description of the COO block layout of every (element kind, matrix, mtype):
// the order in which the reference writes its `Xr[k] / Xc[k] / Xv[k]` triplets
// (loop form at quad4.pyx:1287-1293, unrolled everywhere else).
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/pyfe3d_b200.h"

namespace pf3 {

enum MaskId { MASK_NONE = -1, MASK_FULL = 0, MASK_TT = 1, MASK_M30 = 2, MASK_D18 = 3, MASK_RR = 4 };

inline bool mask_has(int mask, int i, int j) {
  switch (mask) {
    case MASK_FULL: return true;
    case MASK_TT: return i < 3 && j < 3;
    case MASK_M30: return !((i < 3 && j == i + 3) || (i >= 3 && j == i - 3));
    case MASK_D18: return (i < 3) == (j < 3);
    case MASK_RR: return i >= 3 && j >= 3;
    default: return false;
  }
}

struct BlockLayout {
  int nn = 0;          // nodes per element
  int size = 0;        // X_SPARSE_SIZE: stride between consecutive elements in the COO arrays
  int written = 0;     // entries actually written (lumped mass writes fewer)
  int mask = MASK_NONE;
  bool diag_pairs = false;  // only (a,a) node pairs (lumped beam mass)
  std::vector<int8_t> la, li, lb, lj;  // per written entry: node_i, dof_i, node_j, dof_j
};

inline int kind_nodes(int kind) {
  switch (kind) {
    case PF3_QUAD4: case PF3_QUAD4R: return 4;
    case PF3_TRIA3R: return 3;
    case PF3_BEAMC: case PF3_BEAMLR: case PF3_TRUSS: case PF3_SPRING: return 2;
    default: return 0;
  }
}

inline int kind_sparse_size(int kind, int matrix) {
  static const int T[PF3_NKINDS][3] = {{576, 144, 480}, {576, 144, 480}, {324, 81, 270}, {144, 144, 144},
                                       {144, 36, 144},  {72, 0, 144},    {72, 0, 0}};
  if (kind < 0 || kind >= PF3_NKINDS || matrix < 0 || matrix > PF3_MAT_CA) return 0;
  // KA_BETA / KA_GAMMA / CA_SPARSE_SIZE = 144 on the two quads only (quad4.pyx:150-152, quad4r.pyx:116-118)
  if (matrix > PF3_MAT_M) return (kind == PF3_QUAD4 || kind == PF3_QUAD4R) ? 144 : 0;
  return T[kind][matrix];
}

inline BlockLayout make_layout(int kind, int matrix, int mtype) {
  BlockLayout L;
  L.nn = kind_nodes(kind);
  L.size = kind_sparse_size(kind, matrix);
  if (L.nn == 0 || L.size == 0) return L;
  const bool shell = kind <= PF3_TRIA3R;
  if (matrix == PF3_MAT_KC0) {
    L.mask = (kind == PF3_TRUSS || kind == PF3_SPRING) ? MASK_D18 : MASK_FULL;
  } else if (matrix > PF3_MAT_M) {
    L.mask = MASK_TT;   // piston-theory matrices act on the translations like KG (quad4.pyx:9551 ff.)
  } else if (matrix == PF3_MAT_KG) {
    L.mask = shell ? MASK_TT : (kind == PF3_BEAMLR ? MASK_RR : MASK_FULL);
  } else {
    if (shell) {
      L.mask = (mtype == 2) ? MASK_D18 : MASK_M30;
    } else {
      L.mask = (mtype == 0) ? MASK_FULL : MASK_D18;
      L.diag_pairs = (mtype != 0);
    }
  }
  for (int a = 0; a < L.nn; ++a)
    for (int i = 0; i < 6; ++i)
      for (int b = 0; b < L.nn; ++b) {
        if (L.diag_pairs && a != b) continue;
        for (int j = 0; j < 6; ++j)
          if (mask_has(L.mask, i, j)) {
            L.la.push_back(int8_t(a));
            L.li.push_back(int8_t(i));
            L.lb.push_back(int8_t(b));
            L.lj.push_back(int8_t(j));
          }
      }
  L.written = int(L.la.size());
  return L;
}

}  // namespace pf3
