"""COO layout of the reference's element blocks + sum-duplicates assembly.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Every index block in the seven reference element files is row-major over
``(node_i, dof_i, node_j, dof_j)`` restricted to one per-node-pair mask
(e.g. quad4.pyx:1287-1293 for the loop form).
"""
import numpy as np

DOF = 6


def _mask(pairs):
    m = np.zeros((6, 6), bool)
    for i, js in pairs.items():
        m[i, list(js)] = True
    return m


MASKS = {
    # all 6 -> all 6: Quad4/Quad4R/Tria3R/BeamC/BeamLR KC0, BeamC KG, beam consistent M
    "full": np.ones((6, 6), bool),
    # translations only: shell KG (quad4.pyx:1500, tria3r.pyx:3155)
    "tt": _mask({0: (0, 1, 2), 1: (0, 1, 2), 2: (0, 1, 2)}),
    # shell consistent / reduced mass (quad4.pyx:3158,5603): zero diagonal of R.skew.R^T dropped
    "m30": _mask({0: (0, 1, 2, 4, 5), 1: (0, 1, 2, 3, 5), 2: (0, 1, 2, 3, 4),
                  3: (1, 2, 3, 4, 5), 4: (0, 2, 3, 4, 5), 5: (0, 1, 3, 4, 5)}),
    # trans x trans + rot x rot: shell lumped mass (quad4.pyx:8006), Truss/Spring KC0
    # (truss.pyx:439, spring.pyx:345), beam lumped mass (beamc.pyx:2970)
    "d18": _mask({0: (0, 1, 2), 1: (0, 1, 2), 2: (0, 1, 2), 3: (3, 4, 5), 4: (3, 4, 5), 5: (3, 4, 5)}),
    # rotations only: BeamLR KG (beamlr.pyx:1279)
    "rr": _mask({3: (3, 4, 5), 4: (3, 4, 5), 5: (3, 4, 5)}),
}


def block_pattern(nn, mask, diag_pairs_only=False):
    """(local_row, local_col) of the written entries of one element, in the
    reference's write order."""
    mk = MASKS[mask]
    rows, cols = [], []
    for a in range(nn):
        for i in range(6):
            for b in range(nn):
                if diag_pairs_only and a != b:
                    continue
                for j in range(6):
                    if mk[i, j]:
                        rows.append(6 * a + i)
                        cols.append(6 * b + j)
    return np.array(rows), np.array(cols)


def to_global(Kl, R):
    """T Kl T^T with T = blockdiag(R, R, ...) (quad4.pyx:1299-1313)."""
    ne, nd, _ = Kl.shape
    K4 = Kl.reshape(ne, nd // 3, 3, nd // 3, 3)
    return np.einsum("emi,eaibj,enj->eambn", R, K4, R).reshape(ne, nd, nd)


def coo_blocks(Kg, conn, mask, size, diag_pairs_only=False, with_indices=True):
    """Dense global element matrices -> the reference's COO triplets with
    ``init_k = e*size``.  Entries past the written count stay (0, 0, 0.0) as in a
    zero-initialised caller array (lumped-mass tails)."""
    ne, nn = conn.shape
    lr, lc = block_pattern(nn, mask, diag_pairs_only)
    nw = lr.size
    v = np.zeros((ne, size))
    v[:, :nw] = Kg[:, lr, lc]
    if not with_indices:
        return None, None, v.ravel()
    r = np.zeros((ne, size), np.int64)
    c = np.zeros((ne, size), np.int64)
    r[:, :nw] = DOF * conn[:, lr // 6] + lr % 6
    c[:, :nw] = DOF * conn[:, lc // 6] + lc % 6
    return r.ravel(), c.ravel(), v.ravel()


def scatter_fint(fe_local, R, conn, ndof):
    """fint[c_a + m] += R . finte (quad4.pyx:1339-1362)."""
    ne, nn = conn.shape
    F = np.einsum("emi,eai->eam", R, fe_local.reshape(ne, 2 * nn, 3)).reshape(ne, nn, 6)
    out = np.zeros(ndof)
    idx = (DOF * conn)[:, :, None] + np.arange(6)
    np.add.at(out, idx.ravel(), F.ravel())
    return out


def coo_to_csr(n, r, c, v):
    """What ``scipy.sparse.coo_matrix((v,(r,c)),shape=(n,n)).tocsr()`` followed
    by ``sum_duplicates`` yields (call site tests/test_quad4_static_point_load.py:80):
    unique (row, col) pairs in row-major order, duplicates summed; explicit zeros kept."""
    key = r.astype(np.int64) * n + c.astype(np.int64)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    first = np.ones(ks.size, bool)
    first[1:] = ks[1:] != ks[:-1]
    uk = ks[first]
    seg = np.cumsum(first) - 1
    vals = np.zeros(uk.size)
    np.add.at(vals, seg, v[order])
    rows = uk // n
    indptr = np.zeros(n + 1, np.int64)
    np.add.at(indptr, rows + 1, 1)
    return np.cumsum(indptr), (uk % n).astype(np.int64), vals
