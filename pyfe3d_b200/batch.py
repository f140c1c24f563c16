"""Batched entry points: all elements of one kind in one launch, then a device
COO -> CSR assembly.  This is the new API the north star asks for; the
per-element drop-in classes (``pyfe3d_b200.elements``) sit on the same C ABI.

What it replaces in a reference script (tests/test_quad4_static_point_load.py:53-80)::

    for quad in quads:                      ->  b = ElementBatch("quad4", conn, x, prop)
        quad.update_rotation_matrix(x)          KC0 = b.update_KC0()          # device COO
        quad.update_probe_xe(x)                 plan = AssemblyPlan("KC0", nnodes, [b])
        quad.update_KC0(KC0r, KC0c, KC0v, prop) K = plan.assemble(KC0.v)      # device CSR
    KC0 = coo_matrix(...).tocsc()

torch is used only as the owner of device memory and streams.
There is no CPU fallback: constructing a batch without a CUDA device raises.
"""
import numpy as np
import torch

from . import _cabi

KINDS = {"quad4": _cabi.QUAD4, "quad4r": _cabi.QUAD4R, "tria3r": _cabi.TRIA3R, "beamc": _cabi.BEAMC,
         "beamlr": _cabi.BEAMLR, "truss": _cabi.TRUSS, "spring": _cabi.SPRING}
MATRICES = {"KC0": _cabi.MAT_KC0, "KG": _cabi.MAT_KG, "M": _cabi.MAT_M,
            "KA_beta": _cabi.MAT_KA_BETA, "KA_gamma": _cabi.MAT_KA_GAMMA, "CA": _cabi.MAT_CA}
AERO = {"KA_beta": _cabi.KA_BETA, "KA_gamma": _cabi.KA_GAMMA, "CA": _cabi.CA}
SHELL_FIELDS = ["A11", "A12", "A16", "A22", "A26", "A66", "B11", "B12", "B16", "B22", "B26", "B66",
                "D11", "D12", "D16", "D22", "D26", "D66", "E44", "E45", "E55", "scf_k13", "scf_k23", "h",
                "intrho", "intrhoz", "intrhoz2"]
BEAM_FIELDS = ["A", "E", "G", "Iyy", "Izz", "Iyz", "J", "Ay", "Az",
               "intrho", "intrhoy", "intrhoz", "intrhoy2", "intrhoz2", "intrhoyz"]

_CONTEXTS = {}


def context(device=None):
    """The per-device C-ABI context, bound to torch's current stream on that device."""
    if not torch.cuda.is_available():
        raise RuntimeError("pyfe3d_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    ctx = _CONTEXTS.get(idx)
    if ctx is None:
        ctx = _cabi.Context(idx)
        _CONTEXTS[idx] = ctx
    ctx.set_stream(torch.cuda.current_stream(idx).cuda_stream)
    return ctx


def pack_props(kind, props):
    """ShellProp/BeamProp object(s) or an already packed table -> float64[nprop, stride]."""
    if props is None:
        return None
    if isinstance(props, (np.ndarray, torch.Tensor)):
        return props
    shell = KINDS[kind] <= _cabi.TRIA3R
    fields, stride = (SHELL_FIELDS, _cabi.SHELLPROP_STRIDE) if shell else (BEAM_FIELDS, _cabi.BEAMPROP_STRIDE)
    objs = list(props) if isinstance(props, (list, tuple)) else [props]
    out = np.zeros((len(objs), stride))
    for i, o in enumerate(objs):
        for j, f in enumerate(fields):
            out[i, j] = getattr(o, f)
    return out


def _dev(a, dtype, device):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.array(a, copy=True, order="C"), dtype=dtype).to(device)


def _ptr(t):
    return 0 if t is None else t.data_ptr()


class Coo:
    """Device COO triplets of one matrix (r, c may be None when only values were asked for)."""

    def __init__(self, r, c, v, n):
        self.r, self.c, self.v, self.n = r, c, v, n

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.coo_matrix((self.v.cpu().numpy(), (self.r.cpu().numpy(), self.c.cpu().numpy())),
                             shape=(self.n, self.n))


class ElementBatch:
    """All elements of one kind.  Inputs may be numpy arrays or torch tensors (any device);
    they are moved to the target device once and stay resident."""

    def __init__(self, kind, conn, x=None, props=None, prop_id=None, u=None, xmat=None, vxy=None,
                 axes=None, k=None, K6ROT=None, alpha_shear_locking=None, hgfactors=None,
                 nnodes=None, device=None):
        if kind not in KINDS:
            raise ValueError("unknown element kind %r" % (kind,))
        self.kind = kind
        self.kid = KINDS[kind]
        self.ctx = context(device)
        self.device = torch.device("cuda", self.ctx.device)
        self.nn = _cabi.num_nodes(self.kid)
        self.conn = _dev(conn, torch.int64, self.device).reshape(-1, self.nn)
        self.ne = int(self.conn.shape[0])
        self.x = _dev(x, torch.float64, self.device)
        if self.x is not None:
            self.x = self.x.reshape(-1)
        if nnodes is None:
            if self.x is None:
                raise ValueError("nnodes is required when no coordinate array is given")
            nnodes = self.x.numel() // 3
        self.nnodes = int(nnodes)
        self.u = _dev(u, torch.float64, self.device)
        self.props = _dev(pack_props(kind, props), torch.float64, self.device)
        if self.props is not None:
            self.props = self.props.reshape(-1, _cabi.SHELLPROP_STRIDE if self.kid <= _cabi.TRIA3R
                                            else _cabi.BEAMPROP_STRIDE)
        self.prop_id = _dev(prop_id, torch.int32, self.device)
        evec = xmat if self.kid <= _cabi.TRIA3R else (axes if kind == "spring" else vxy)
        self.evec = _dev(evec, torch.float64, self.device)
        self.evec_stride = 0
        if self.evec is not None:
            w = 6 if kind == "spring" else 3
            self.evec = self.evec.reshape(-1, w)
            if self.evec.shape[0] not in (1, self.ne):
                raise ValueError("per-element vector must have 1 or ne rows")
            self.evec_stride = w if self.evec.shape[0] == self.ne and self.ne > 1 else 0
        self.eparam = None
        if kind == "spring":
            if k is None:
                raise ValueError("spring batches need k[ne,6] = (kxe,kye,kze,krxe,krye,krze)")
            ep = torch.zeros((self.ne, _cabi.EPARAM_STRIDE), dtype=torch.float64, device=self.device)
            ep[:, :6] = _dev(k, torch.float64, self.device).reshape(self.ne, 6)
            self.eparam = ep
        elif K6ROT is not None or alpha_shear_locking is not None or hgfactors is not None:
            ep = torch.zeros((self.ne, _cabi.EPARAM_STRIDE), dtype=torch.float64, device=self.device)
            ep[:, 0] = 100. if K6ROT is None else _dev(np.broadcast_to(np.asarray(K6ROT, float), (self.ne,)),
                                                        torch.float64, self.device)
            ep[:, 1] = 0.7 if alpha_shear_locking is None else _dev(
                np.broadcast_to(np.asarray(alpha_shear_locking, float), (self.ne,)), torch.float64, self.device)
            ep[:, 2:7] = 1. if hgfactors is None else _dev(
                np.broadcast_to(np.asarray(hgfactors, float), (self.ne, 5)), torch.float64, self.device)
            self.eparam = ep
        self.sizes = {m: _cabi.sparse_size(self.kid, i) for m, i in MATRICES.items()}

    # -- plumbing ---------------------------------------------------------------------------
    def cabi_batch(self, mtype=0, stress=(0., 0., 0.), u=None, state=None):
        u = self.u if u is None else u
        return _cabi.Batch(self.kid, self.ne, self.nnodes, _ptr(self.conn), _ptr(self.x), _ptr(u),
                           _ptr(self.props), _ptr(self.prop_id),
                           0 if self.props is None else self.props.shape[0], _ptr(self.evec),
                           self.evec_stride, _ptr(self.eparam), _ptr(state), mtype,
                           tuple(float(s) for s in stress))

    def _alloc(self, matrix, indices, out):
        n = self.ne * self.sizes[matrix]
        if out is not None:
            return out
        v = torch.zeros(n, dtype=torch.float64, device=self.device)
        r = c = None
        if indices:
            r = torch.zeros(n, dtype=torch.int64, device=self.device)
            c = torch.zeros(n, dtype=torch.int64, device=self.device)
        return Coo(r, c, v, 6 * self.nnodes)

    def evaluate(self, KC0=False, KG=False, KG_given_stress=None, M=False, mtype=0, fint=None, u=None,
                 indices=True, out=None, accumulate=False):
        """One fused launch for any subset of {KC0, KG | KG_given_stress, M, fint}.

        Returns a dict name -> Coo.  ``out`` may carry preallocated Coo objects to overwrite
        (``accumulate=True`` gives the reference's ``+=``).  ``indices=False`` is the
        reference's ``update_*v_only=1``."""
        if u is not None:
            u = _dev(u, torch.float64, self.device)
        ctx = context(self.device)
        out = dict(out or {})
        what = 0
        coos = {}
        for name, on in (("KC0", KC0), ("KG", KG or KG_given_stress is not None), ("M", M)):
            if not on:
                continue
            if self.sizes[name] == 0:
                raise ValueError("%s has no %s matrix" % (self.kind, name))
            coos[name] = self._alloc(name, indices, out.get(name))
        if KC0:
            what |= _cabi.KC0
        if KG_given_stress is not None:
            what |= _cabi.KG_STRESS
        elif KG:
            what |= _cabi.KG
        if M:
            what |= _cabi.M
        if fint is not None:
            what |= _cabi.FINT
            if not (isinstance(fint, torch.Tensor) and fint.is_cuda and fint.dtype == torch.float64):
                raise TypeError("fint must be a float64 CUDA tensor (it is accumulated in place)")

        def cc(name):
            k = coos.get(name)
            if k is None:
                return None
            return _cabi.Coo(_ptr(k.r) if indices else 0, _ptr(k.c) if indices else 0, _ptr(k.v), 0,
                             1 if accumulate else 0)

        b = self.cabi_batch(mtype, (0., 0., 0.) if KG_given_stress is None else KG_given_stress, u)
        ctx.eval(b, what, cc("KC0"), cc("KG"), cc("M"), _ptr(fint))
        return coos

    def evaluate_aero(self, KA_beta=False, KA_gamma=False, CA=False, indices=True, out=None):
        """Piston-theory aerodynamic matrices of every Quad4 / Quad4R element in ONE launch (update_KA_beta,
        update_KA_gamma, update_CA of the reference: quad4.pyx:9491, 10312, 11115).  Returns {name: Coo}."""
        if self.kind not in ("quad4", "quad4r"):
            raise ValueError("the piston-theory matrices exist on Quad4 and Quad4R only")
        coos = dict(out or {})
        what = 0
        for name, on in (("KA_beta", KA_beta), ("KA_gamma", KA_gamma), ("CA", CA)):
            if on:
                what |= AERO[name]
                if name not in coos:
                    coos[name] = self._alloc(name, indices, None)

        def cc(name):
            k = coos.get(name)
            if k is None or not (what & AERO[name]):
                return None
            return _cabi.Coo(_ptr(k.r) if indices else 0, _ptr(k.c) if indices else 0, _ptr(k.v), 0, 0)

        context(self.device).eval_aero(self.cabi_batch(), what, cc("KA_beta"), cc("KA_gamma"), cc("CA"))
        return coos

    def update_KA_beta(self, **kw):
        return self.evaluate_aero(KA_beta=True, **kw)["KA_beta"]

    def update_KA_gamma(self, **kw):
        return self.evaluate_aero(KA_gamma=True, **kw)["KA_gamma"]

    def update_CA(self, **kw):
        return self.evaluate_aero(CA=True, **kw)["CA"]

    # -- reference-named conveniences --------------------------------------------------------
    def update_KC0(self, update_KC0v_only=0, **kw):
        return self.evaluate(KC0=True, indices=not update_KC0v_only, **kw)["KC0"]

    def update_KG(self, u=None, update_KGv_only=0, **kw):
        return self.evaluate(KG=True, u=u, indices=not update_KGv_only, **kw)["KG"]

    def update_KG_given_stress(self, Nxx, Nyy, Nxy, update_KGv_only=0, **kw):
        return self.evaluate(KG_given_stress=(Nxx, Nyy, Nxy), indices=not update_KGv_only, **kw)["KG"]

    def update_M(self, mtype=0, **kw):
        return self.evaluate(M=True, mtype=mtype, **kw)["M"]

    def update_fint(self, fint, u=None):
        self.evaluate(fint=fint, u=u)
        return fint

    def fill_indices(self, matrix, mtype=0, coo=None):
        coo = coo or self._alloc(matrix, True, None)
        context(self.device).fill_indices(self.kid, MATRICES[matrix], mtype, self.ne, _ptr(self.conn), 0,
                                          _ptr(coo.r), _ptr(coo.c))
        return coo

    def state(self, u=None):
        """[ne, 50] per-element state (R, m, area|length, xe, ue) as the reference leaves it
        on the element/probe after update_rotation_matrix + update_probe_xe + update_probe_ue."""
        if u is not None:
            u = _dev(u, torch.float64, self.device)
        out = torch.zeros((self.ne, _cabi.STATE_STRIDE), dtype=torch.float64, device=self.device)
        context(self.device).eval_state(self.cabi_batch(u=u), _ptr(out))
        return out

    def finte(self, u=None):
        """probe.finte of every element: [ne, 6*nn] local internal forces."""
        if u is not None:
            u = _dev(u, torch.float64, self.device)
        out = torch.zeros((self.ne, 6 * self.nn), dtype=torch.float64, device=self.device)
        context(self.device).eval_finte(self.cabi_batch(u=u), _ptr(out))
        return out


class AssemblyPlan:
    """Symbolic COO -> CSR assembly of ONE matrix from connectivity only.

    ``batches``: ElementBatch objects contributing to the matrix (<= 8 kinds, e.g. Quad4 skin +
    BeamC stiffeners); their COO value blocks are expected back to back in one value array, group
    g starting at ``coo_offsets[g]`` (default: concatenated in order).  ``node_range`` restricts
    the plan to the DOF rows of nodes [begin, end): the multi-GPU row-ownership shard."""

    def __init__(self, matrix, nnodes, batches, coo_offsets=None, node_range=None, mtype=0):
        self.matrix = matrix
        self.nnodes = int(nnodes)
        self.batches = list(batches)
        self.device = self.batches[0].device
        if coo_offsets is None:
            coo_offsets, o = [], 0
            for b in self.batches:
                coo_offsets.append(o)
                o += b.ne * b.sizes[matrix]
        self.coo_offsets = [int(o) for o in coo_offsets]
        self.coo_size = max(o + b.ne * b.sizes[matrix] for o, b in zip(self.coo_offsets, self.batches))
        nb, ne = node_range if node_range is not None else (0, self.nnodes)
        self.node_begin, self.node_end = int(nb), int(ne)
        ctx = context(self.device)
        self._plan = _cabi.Plan.structured(ctx, MATRICES[matrix], self.nnodes,
                                           [b.cabi_batch(mtype) for b in self.batches],
                                           self.coo_offsets, self.node_begin, self.node_end)
        self.nnz = self._plan.nnz
        self.nrows = self._plan.nrows
        self._pattern = None

    def pattern(self):
        """(indptr[nrows+1], indices[nnz]) int64 device tensors; rows are local to the shard,
        columns global; column indices sorted within each row."""
        if self._pattern is None:
            indptr = torch.empty(self.nrows + 1, dtype=torch.int64, device=self.device)
            indices = torch.empty(self.nnz, dtype=torch.int64, device=self.device)
            context(self.device)
            self._plan.pattern(_ptr(indptr), _ptr(indices))
            self._pattern = (indptr, indices)
        return self._pattern

    def assemble(self, coo_v, out=None):
        """Numeric phase (repeatable): csr values = deterministic sum of duplicates."""
        if coo_v.numel() < self.coo_size:
            raise ValueError("COO value array shorter than the plan's layout")
        if out is None:
            out = torch.empty(self.nnz, dtype=torch.float64, device=self.device)
        context(self.device)
        self._plan.assemble(_ptr(coo_v), _ptr(out))
        return out

    def to_scipy(self, vals):
        import scipy.sparse as sp
        indptr, indices = self.pattern()
        return sp.csr_matrix((vals.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()),
                             shape=(self.nrows, 6 * self.nnodes))

    def spmv(self, vals, x, free=None, out=None):
        """y = A x (or P A P x with P = diag(free)) for the CSR values ``vals`` of this plan, through the plan's
        node-block structure (no per-entry indices are read).  ``x``: [6*nnodes] global, ``free``: optional uint8
        [6*nnodes] DOF mask, result: [nrows] (the plan's own rows)."""
        if out is None:
            out = torch.empty(self.nrows, dtype=torch.float64, device=self.device)
        context(self.device)
        self._plan.spmv(_ptr(vals), _ptr(free) if free is not None else 0, _ptr(x), _ptr(out))
        return out

    def spmv_scaled(self, vals, scale, x, free=None, out=None):
        """y = S P A P S x with S = diag(scale): the diagonally scaled operators (``D_inv_sqrt @ KC0uu @ D_inv_sqrt``)
        the reference scripts hand to eigsh, tests/test_quad4r_linear_buckling_plate.py:172-180."""
        if out is None:
            out = torch.empty(self.nrows, dtype=torch.float64, device=self.device)
        context(self.device)
        self._plan.spmv_scaled(_ptr(vals), _ptr(free) if free is not None else 0, _ptr(scale), _ptr(x), _ptr(out))
        return out

    def diagonal(self, vals, out=None):
        """Diagonal of the plan's row block (Jacobi scaling)."""
        if out is None:
            out = torch.empty(self.nrows, dtype=torch.float64, device=self.device)
        context(self.device)
        self._plan.diagonal(_ptr(vals), _ptr(out))
        return out

    def update_fint(self, fint, u=None):
        """fint += internal forces of every batch of the plan (update_fint of the reference), gathered per
        node through the plan's incidence lists: deterministic, no per-call sort, owned rows only."""
        if not (isinstance(fint, torch.Tensor) and fint.is_cuda and fint.dtype == torch.float64):
            raise TypeError("fint must be a float64 CUDA tensor (it is accumulated in place)")
        if u is not None:
            u = _dev(u, torch.float64, self.device)
        context(self.device)
        for g, b in enumerate(self.batches):
            self._plan.fint(g, b.cabi_batch(u=u), _ptr(fint))
        return fint

    # -- fused evaluate + assemble (Quad4 / Quad4R, single batch) ---------------------------------
    def csr_sizes(self, mtype=0):
        """nnz of the KC0 / KG / M CSR value arrays the fused kernel fills (same layouts as the
        structured plans of the same batch for those matrices)."""
        nblk = self._plan.nblocks
        return {"KC0": nblk * 36, "KG": nblk * 9, "M": nblk * (18 if mtype == 2 else 30)}

    def evaluate_assemble(self, KC0=False, KG=False, KG_given_stress=None, M=False, mtype=0, u=None,
                          coo=None, csr=None, write_coo=True, indices=False):
        """Element matrices -> COO value arrays AND assembled CSR values.

        For the "KC0" plan of a single Quad4/Quad4R/Tria3R batch this is ONE fused kernel that never re-reads
        the COO arrays; for every other element kind (or when a node couples to more than 16 nodes) it runs the
        two-pass path (one evaluation launch + one slab assembly per matrix) with the same outputs.

        ``coo`` / ``csr``: optional dicts of preallocated outputs (name -> Coo / tensor).  With
        ``write_coo=False`` only the CSR values are returned.  Returns (coo, csr) dicts."""
        if self.matrix != "KC0":
            raise ValueError("evaluate_assemble needs the KC0 plan")
        if len(self.batches) != 1:
            return self._evaluate_assemble_mixed(KC0, KG, KG_given_stress, M, mtype, u, coo, csr, write_coo, indices)
        b = self.batches[0]
        if u is not None:
            u = _dev(u, torch.float64, self.device)
        if b.kind not in ("quad4", "quad4r", "tria3r") or getattr(self, "_fused_unsupported", False):
            return self._evaluate_assemble_two_pass(KC0, KG, KG_given_stress, M, mtype, u, coo, csr, write_coo,
                                                    indices)
        sizes = self.csr_sizes(mtype)
        coo = dict(coo or {})
        csr = dict(csr or {})
        what = 0
        for name, on in (("KC0", KC0), ("KG", KG or KG_given_stress is not None), ("M", M)):
            if not on:
                continue
            if name not in csr:
                csr[name] = torch.empty(sizes[name], dtype=torch.float64, device=self.device)
            if write_coo and name not in coo:
                coo[name] = b._alloc(name, indices, None)
        if KC0:
            what |= _cabi.KC0
        if KG_given_stress is not None:
            what |= _cabi.KG_STRESS
        elif KG:
            what |= _cabi.KG
        if M:
            what |= _cabi.M

        def cc(name):
            k = coo.get(name) if write_coo else None
            if k is None:
                return None
            return _cabi.Coo(_ptr(k.r) if indices else 0, _ptr(k.c) if indices else 0, _ptr(k.v), 0, 0)

        context(self.device)
        try:
            self._plan.eval_assemble(b.cabi_batch(mtype, (0., 0., 0.) if KG_given_stress is None else KG_given_stress, u), what, cc("KC0"),
                                     cc("KG"), cc("M"), _ptr(csr.get("KC0")), _ptr(csr.get("KG")),
                                     _ptr(csr.get("M")))
        except _cabi.Pf3Error as exc:
            if "capacity" not in str(exc) and "not defined" not in str(exc):
                raise
            self._fused_unsupported = True      # e.g. a node coupled to more than 16 nodes
            return self._evaluate_assemble_two_pass(KC0, KG, KG_given_stress, M, mtype, u, coo, csr, write_coo,
                                                    indices)
        return coo, csr

    def evaluate_assemble_host(self, x, u, out, KC0=False, KG=False, KG_given_stress=None, M=False, mtype=0,
                               coo=None):
        """One step with HOST buffers through ``pf3_eval_assemble_host``: ``x`` / ``u`` are host arrays (numpy or CPU
        tensors, pinned for speed) holding this step's coordinates / displacements, ``out`` maps "KC0"/"KG"/"M" to
        host arrays that receive the assembled CSR values (``csr_sizes()`` gives their lengths).  Connectivity,
        properties and the plan stay on the device; ``coo`` optionally maps names to DEVICE ``Coo`` value arrays that are
        filled as well.  Quad4 / Quad4R / Tria3R single-batch KC0 plans only."""
        if self.matrix != "KC0" or len(self.batches) != 1:
            raise ValueError("evaluate_assemble_host needs the KC0 plan of a single batch")
        b = self.batches[0]

        def hptr(a, n, name):
            if a is None:
                return 0
            t = torch.as_tensor(a) if not isinstance(a, torch.Tensor) else a
            if t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous() or t.numel() < n:
                raise ValueError("%s must be a contiguous float64 HOST array with at least %d entries" % (name, n))
            return t.data_ptr()

        sizes = self.csr_sizes(mtype)
        what = 0
        if KC0:
            what |= _cabi.KC0
        if KG_given_stress is not None:
            what |= _cabi.KG_STRESS
        elif KG:
            what |= _cabi.KG
        if M:
            what |= _cabi.M
        hb = _cabi.Batch(b.kid, b.ne, b.nnodes, _ptr(b.conn), hptr(x, 3 * b.nnodes, "x"),
                         hptr(u, 6 * b.nnodes, "u") if u is not None else 0, _ptr(b.props), _ptr(b.prop_id),
                         0 if b.props is None else b.props.shape[0], _ptr(b.evec), b.evec_stride, _ptr(b.eparam), 0,
                         mtype, tuple(float(t) for t in ((0., 0., 0.) if KG_given_stress is None else KG_given_stress)))
        def cc(name):
            k = (coo or {}).get(name)
            return None if k is None else _cabi.Coo(0, 0, _ptr(k.v), 0, 0)

        context(self.device)
        self._plan.eval_assemble_host(hb, what, cc("KC0"), cc("KG"), cc("M"),
                                      hptr(out.get("KC0"), sizes["KC0"], "out['KC0']") if KC0 else 0,
                                      hptr(out.get("KG"), sizes["KG"], "out['KG']") if (KG or KG_given_stress is not None) else 0,
                                      hptr(out.get("M"), sizes["M"], "out['M']") if M else 0)
        return out

    def _evaluate_assemble_mixed(self, KC0, KG, KG_given_stress, M, mtype, u, coo, csr, write_coo, indices):
        """Several element kinds in ONE matrix (e.g. Quad4 skin + BeamC stiffeners, BASELINE config 5).  When the first
        batch is Quad4 / Quad4R its share goes through the fused kernel straight into the union CSR layouts
        (``pf3_eval_assemble_group``); the other batches are evaluated into their slices of the plan-wide COO arrays
        and added (``pf3_plan_assemble_add``).  Otherwise: evaluation per batch + assembly per matrix.
        ``coo[name].v`` are plan-wide value arrays (batch g at the sibling plan's ``coo_offsets[g]``)."""
        if KG_given_stress is not None or indices or not write_coo:
            raise ValueError("mixed plans: KG_given_stress, index arrays and write_coo=False are not supported here")
        if u is not None:
            u = _dev(u, torch.float64, self.device)
        names = [n for n, on in (("KC0", KC0), ("KG", KG), ("M", M)) if on]
        plans = {n: self._sibling(n, mtype) for n in names}
        coo = dict(coo or {})
        csr = dict(csr or {})
        for n in names:
            if n not in coo:
                coo[n] = Coo(None, None, torch.zeros(plans[n].coo_size, dtype=torch.float64, device=self.device),
                             6 * self.nnodes)
            if n not in csr:
                csr[n] = torch.empty(plans[n].nnz, dtype=torch.float64, device=self.device)
        def views(g):
            b = self.batches[g]
            return {n: Coo(None, None, coo[n].v[plans[n].coo_offsets[g]:plans[n].coo_offsets[g] + b.ne * b.sizes[n]],
                           6 * self.nnodes) for n in names}

        b0 = self.batches[0]
        fused = b0.kind in ("quad4", "quad4r") and not getattr(self, "_fused_unsupported", False)
        # the fused kernel walks THIS (KC0) plan's node blocks for every matrix it writes: a matrix whose own plan has
        # a different block structure (lumped beam / truss mass: diagonal node pairs only, pattern.hpp diag_pairs) must
        # take the two-pass path, otherwise its CSR array would be written at the KC0 plan's offsets
        fnames = [n for n in names if fused and plans[n]._plan.nblocks == self._plan.nblocks]
        if fnames:
            what = ((_cabi.KC0 if "KC0" in fnames else 0) | (_cabi.KG if "KG" in fnames else 0)
                    | (_cabi.M if "M" in fnames else 0))

            def cc(n):
                return _cabi.Coo(0, 0, _ptr(coo[n].v), plans[n].coo_offsets[0], 0) if n in fnames else None

            context(self.device)
            try:
                self._plan.eval_assemble_group(b0.cabi_batch(mtype, (0., 0., 0.), u), 0, what, cc("KC0"), cc("KG"),
                                               cc("M"), _ptr(csr.get("KC0")) if "KC0" in fnames else 0,
                                               _ptr(csr.get("KG")) if "KG" in fnames else 0,
                                               _ptr(csr.get("M")) if "M" in fnames else 0)
            except _cabi.Pf3Error as exc:
                if "capacity" not in str(exc) and "not defined" not in str(exc):
                    raise
                self._fused_unsupported = True
                fnames = []
        rest = [n for n in names if n not in fnames]
        for g in range(len(self.batches)):
            todo = rest if g == 0 else names
            if not todo:
                continue
            v = views(g)
            self.batches[g].evaluate(KC0="KC0" in todo, KG="KG" in todo, M="M" in todo, mtype=mtype, u=u,
                                     indices=False, out={n: v[n] for n in todo})
        for n in names:
            if n in fnames:
                context(self.device)
                plans[n]._plan.assemble_add(_ptr(coo[n].v), _ptr(csr[n]), 0)
            else:
                plans[n].assemble(coo[n].v, out=csr[n])
        return coo, csr

    def _sibling(self, matrix, mtype):
        """Plan of another matrix of the same batches / row shard (cached)."""
        if matrix == self.matrix and mtype == 0:
            return self
        cache = self.__dict__.setdefault("_siblings", {})
        key = (matrix, mtype if matrix == "M" else 0)
        if key not in cache:
            cache[key] = AssemblyPlan(matrix, self.nnodes, self.batches,
                                      node_range=(self.node_begin, self.node_end), mtype=key[1])
        return cache[key]

    def _evaluate_assemble_two_pass(self, KC0, KG, KG_given_stress, M, mtype, u, coo, csr, write_coo, indices):
        b = self.batches[0]
        coo = dict(coo or {})
        csr = dict(csr or {})
        res = b.evaluate(KC0=KC0, KG=KG, KG_given_stress=KG_given_stress, M=M, mtype=mtype, u=u, indices=indices,
                         out=coo)
        for name, k in res.items():
            plan = self._sibling(name, mtype)
            csr[name] = plan.assemble(k.v, out=csr.get(name))
        return (res if write_coo else {}), csr


class CooPlan:
    """Generic plan from arbitrary COO index arrays (what scipy's tocsr does)."""

    def __init__(self, n, r, c, device=None):
        ctx = context(device)
        self.device = torch.device("cuda", ctx.device)
        self.r = _dev(r, torch.int64, self.device)
        self.c = _dev(c, torch.int64, self.device)
        self.n = int(n)
        self._plan = _cabi.Plan.from_coo(ctx, self.n, self.r.numel(), _ptr(self.r), _ptr(self.c))
        self.nnz = self._plan.nnz
        self.nrows = self.n
        self._pattern = None

    pattern = AssemblyPlan.pattern

    def assemble(self, coo_v, out=None):
        coo_v = _dev(coo_v, torch.float64, self.device)
        if out is None:
            out = torch.empty(self.nnz, dtype=torch.float64, device=self.device)
        context(self.device)
        self._plan.assemble(_ptr(coo_v), _ptr(out))
        return out

    def to_scipy(self, vals):
        import scipy.sparse as sp
        indptr, indices = self.pattern()
        return sp.csr_matrix((vals.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()),
                             shape=(self.n, self.n))


def spmv(indptr, indices, vals, x, out=None):
    """y = A @ x on the device for a CSR matrix with int64 indptr/indices."""
    nrows = indptr.numel() - 1
    if out is None:
        out = torch.empty(nrows, dtype=torch.float64, device=vals.device)
    context(vals.device).spmv_csr(nrows, _ptr(indptr), _ptr(indices), _ptr(vals), _ptr(x), _ptr(out))
    return out


def laminate_table(stack, plyts, laminaprops, rhos=0., offset=0., calc_scf=True, device=None, out=None):
    """Device property table ``[nrows, 32]`` for MANY laminates at once (``pf3_laminate_props``): row r is what the
    reference's ``laminated_plate(stack[r], plyts=plyts[r], laminaprops=..., rhos=..., offset=offset[r],
    calc_scf=calc_scf)`` (pyfe3d/shellprop_utils.py:96) stores in its ShellProp, ready to be an ``ElementBatch``'s
    ``props`` (with ``prop_id = arange(ne)`` for one laminate per element).

    ``stack``: angles in degrees ``[nplies]`` (shared) or ``[nrows, nplies]``; ``plyts``: ``[nplies]`` or
    ``[nrows, nplies]``; ``laminaprops``: ``(E, nu)``, ``(e1, e2, nu12, g12, g13, g23)`` (shared by all plies), a
    ``[nplies, 6]`` array or ``[nrows, nplies, 6]``; ``rhos``: scalar, ``[nplies]`` or ``[nrows, nplies]``;
    ``offset``: scalar or ``[nrows]``.  Arrays may be numpy or CUDA tensors; the number of rows is the largest
    leading dimension given."""
    ctx = context(device)
    dev = torch.device("cuda", ctx.device)

    def t(a):
        return a.to(dev, torch.float64) if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, float)).to(dev)

    th, tk = t(stack), t(plyts)
    nplies = th.shape[-1]
    lp = t(laminaprops)
    if lp.ndim == 1:
        if lp.numel() == 2:
            e, nu = lp[0], lp[1]
            g = e / (2 * (1 + nu))
            lp = torch.stack([e, e, nu, g, g, g])
        lp = lp[:6].expand(nplies, 6)
    lp = lp[..., :6]
    rh = t(rhos)
    rh = rh.expand(lp.shape[:-1]) if rh.ndim <= lp.ndim - 1 else rh
    if rh.ndim == 2 and lp.ndim == 2:
        lp = lp.expand(rh.shape[0], nplies, 6)
    lam = torch.zeros(tuple(lp.shape[:-1]) + (8,), dtype=torch.float64, device=dev)
    lam[..., :6] = lp
    lam[..., 6] = rh
    off = t(offset)
    nrows = max([a.shape[0] for a, nd in ((th, 2), (tk, 2), (lam, 3), (off, 1)) if a.ndim == nd] + [1])
    for a, nd, name in ((th, 2, "stack"), (tk, 2, "plyts"), (lam, 3, "laminaprops/rhos"), (off, 1, "offset")):
        if a.ndim == nd and a.shape[0] != nrows:
            raise ValueError("%s has %d rows, expected %d" % (name, a.shape[0], nrows))
    if tk.shape[-1] != nplies or lam.shape[-2] != nplies:
        raise ValueError("stack, plyts and laminaprops must describe the same number of plies")
    th, tk, lam, off = th.contiguous(), tk.contiguous(), lam.contiguous(), off.contiguous()
    if out is None:
        out = torch.empty((nrows, _cabi.SHELLPROP_STRIDE), dtype=torch.float64, device=dev)
    ctx.laminate_props(nrows, nplies, _ptr(th), nplies if th.ndim == 2 else 0, _ptr(tk), nplies if tk.ndim == 2 else 0,
                       _ptr(lam), 8 * nplies if lam.ndim == 3 else 0, _ptr(off), 1 if off.ndim == 1 else 0,
                       1 if calc_scf else 0, _ptr(out))
    return out


LP_VARIABLES = ("h", "xiA1", "xiA2", "xiA3", "xiA4", "xiB1", "xiB2", "xiB3", "xiB4",
                "xiD1", "xiD2", "xiD3", "xiD4", "xiE1", "xiE2")


def lamination_parameter_table(thickness, invariants, lp, rho=None, grad=(), grad_complete=False, device=None):
    """Device property table ``[nrows, 32]`` from total thickness, material invariants and lamination parameters
    (``pf3_lamination_parameter_props``): row r is what the reference's
    ``shellprop_from_LaminationParameters(thickness[r], mat, lp[r])`` (pyfe3d/shellprop.pyx:767) stores.

    ``thickness``: scalar or ``[nrows]``; ``invariants``: ``(u1..u7)`` (``MatLamina.invariants()``) or ``[nrows, 7]``;
    ``lp``: the 14 parameters (``LaminationParameters.as_array()``) or ``[nrows, 14]``; ``rho``: optional homogeneous
    density (scalar or ``[nrows]``) filling the mass integrals, which the reference leaves 0.

    ``grad``: names from ``LP_VARIABLES``; when given, returns ``(table, grad_table)`` with ``grad_table`` of shape
    ``[nrows, len(grad), 32]`` (rows in ``LP_VARIABLES`` order): property rows holding d(A, B, D, E, mass integrals)/dv
    -- ``GradABDE.calc_LP_grad`` (shellprop.pyx:933) -- so that an ``ElementBatch`` evaluated with
    ``props=grad_table[:, j]`` returns dKC0/dv_j and dM/dv_j.  ``grad_complete=False`` keeps the reference's zero
    d(X66)/d(xi) entries."""
    ctx = context(device)
    dev = torch.device("cuda", ctx.device)

    def t(a):
        return a.to(dev, torch.float64) if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, float)).to(dev)

    hh, uu, xx = t(thickness).contiguous(), t(invariants).contiguous(), t(lp).contiguous()
    rr = None if rho is None else t(rho).contiguous()
    if uu.shape[-1] != 7 or xx.shape[-1] != 14:
        raise ValueError("invariants must have 7 columns (u1..u7) and lp 14 (xiA1..4, xiB1..4, xiD1..4, xiE1..2)")
    lead = [a.shape[0] for a, nd in ((hh, 1), (uu, 2), (xx, 2)) if a.ndim == nd]
    if rr is not None and rr.ndim == 1:
        lead.append(rr.shape[0])
    nrows = max(lead + [1])
    if any(n != nrows for n in lead):
        raise ValueError("thickness, invariants, lp and rho disagree on the number of rows: %s" % lead)
    mask = 0
    for name in grad:
        mask |= 1 << LP_VARIABLES.index(name)
    out = torch.empty((nrows, _cabi.SHELLPROP_STRIDE), dtype=torch.float64, device=dev)
    g = torch.empty((nrows, bin(mask).count("1"), _cabi.SHELLPROP_STRIDE), dtype=torch.float64, device=dev) if mask else None
    ctx.lamination_parameter_props(nrows, _ptr(hh), 1 if hh.ndim == 1 else 0, _ptr(uu), 7 if uu.ndim == 2 else 0,
                                   _ptr(xx), 14 if xx.ndim == 2 else 0, 0 if rr is None else _ptr(rr),
                                   1 if rr is not None and rr.ndim == 1 else 0, mask, 1 if grad_complete else 0,
                                   _ptr(out), 0 if g is None else _ptr(g))
    return (out, g) if mask else out
