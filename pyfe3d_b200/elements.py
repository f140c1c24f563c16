"""Per-element drop-in classes: ``Quad4(probe).update_KC0(KC0r, KC0c, KC0v, prop)`` ...

Same class names, attributes, method names, argument order and defaults as the
reference's ``cdef class``es (pyfe3d/quad4.pyx:142-488, quad4r.pyx:108-284,
tria3r.pyx:128-292, beamc.pyx:23-156, beamlr.pyx:23-158, truss.pyx:29-148,
spring.pyx:21-145), so existing element-loop scripts run unchanged after
``import pyfe3d_b200 as pyfe3d``.

Every method is a batch-of-one call into the same CUDA kernels through the host-pointer
entry points of the C ABI (``pf3_eval_host`` / ``pf3_eval_state_host`` /
``pf3_eval_finte_host`` / ``pf3_eval_aero_host``): the caller's numpy arrays are packed
into one pinned staging buffer, copied to the device once, evaluated, and copied back
once (two PCIe transfers + one stream synchronisation per method call, on the context's
own stream; the call returns when the arrays hold the result).  The host objects only
hold state (r11..r33, m11..m22, area/length, probe.xe/ue/finte) exactly like the
reference; no arithmetic is done on the host and there is no CPU fallback.  The per-call
latency (tens of microseconds) makes this path a compatibility layer — the fast path is
:class:`pyfe3d_b200.batch.ElementBatch` (``pyfe3d_b200.elements.CALLS`` counts the
per-element calls made so far; a one-time ``UserWarning`` points there after 50 000 calls).
"""
import ctypes as _ct

import numpy as np

from . import _cabi

# module-level constants every reference element module defines (e.g. pyfe3d/quad4.pyx:20-24)
DOF = 6
INT = np.int64 if _ct.sizeof(_ct.c_long) == 8 else np.int32
DOUBLE = np.float64

_CTX = [None]
CALLS = [0]          # per-element C-ABI calls made by this process (see the module docstring)
_HINT_AT = 50000


def _ctx():
    if _CTX[0] is None:
        if _cabi.device_count() <= 0:
            raise RuntimeError("pyfe3d_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        _CTX[0] = _cabi.Context(0)   # its own stream; every per-element call is synchronous on it
    CALLS[0] += 1
    if CALLS[0] == _HINT_AT:
        import warnings
        warnings.warn("pyfe3d_b200: %d per-element calls so far, each a separate device round trip; "
                      "pyfe3d_b200.batch.ElementBatch evaluates all elements of one kind in one launch" % _HINT_AT,
                      UserWarning, stacklevel=3)
    return _CTX[0]


def _p(a):
    return 0 if a is None else a.ctypes.data


def _check_array(a, dtype, name):
    """Typed-memoryview contract of the reference (`long[::1]`, `double[::1]`): wrong dtype or
    non-contiguous input raises before entry (SURVEY §8(b) 'Array contract')."""
    if not isinstance(a, np.ndarray) or a.dtype != dtype or a.ndim != 1 or not a.flags.c_contiguous:
        raise ValueError("%s must be a C-contiguous 1-D numpy array of dtype %s" % (name, np.dtype(dtype)))
    if not a.flags.writeable:
        raise ValueError("%s must be writable" % name)
    return a


_SHELL_FIELDS = ["A11", "A12", "A16", "A22", "A26", "A66", "B11", "B12", "B16", "B22", "B26", "B66",
                 "D11", "D12", "D16", "D22", "D26", "D66", "E44", "E45", "E55", "scf_k13", "scf_k23", "h",
                 "intrho", "intrhoz", "intrhoz2"]
_BEAM_FIELDS = ["A", "E", "G", "Iyy", "Izz", "Iyz", "J", "Ay", "Az",
                "intrho", "intrhoy", "intrhoz", "intrhoy2", "intrhoz2", "intrhoyz"]


class _Data:
    KIND = None

    def __init__(self):
        self.KC0_SPARSE_SIZE = _cabi.sparse_size(self.KIND, _cabi.MAT_KC0)
        self.KG_SPARSE_SIZE = _cabi.sparse_size(self.KIND, _cabi.MAT_KG)
        self.M_SPARSE_SIZE = _cabi.sparse_size(self.KIND, _cabi.MAT_M)


class _Probe:
    KIND = None

    def __init__(self):
        nn = _cabi.num_nodes(self.KIND)
        self.xe = np.zeros(3 * nn)
        self.ue = np.zeros(6 * nn)
        self.finte = np.zeros(6 * nn)


class _Element:
    KIND = None
    SHELL = False

    def __init__(self, probe):
        self.probe = probe
        self.eid = -1
        self.pid = -1
        self._nn = _cabi.num_nodes(self.KIND)
        for a in range(self._nn):
            setattr(self, "n%d" % (a + 1), -1)
            setattr(self, "c%d" % (a + 1), -1)
        self.init_k_KC0 = 0
        self.init_k_KG = 0
        self.init_k_M = 0
        self.r11 = self.r12 = self.r13 = 0.
        self.r21 = self.r22 = self.r23 = 0.
        self.r31 = self.r32 = self.r33 = 0.
        self._conn = np.arange(self._nn, dtype=np.int64)

    # ---- state <-> C ABI ------------------------------------------------------------------
    def _geo(self):
        return self.area if self.SHELL else self.length

    def _set_geo(self, v):
        if self.SHELL:
            self.area = v
        else:
            self.length = v

    def _state(self):
        s = np.zeros(_cabi.STATE_STRIDE)
        s[0:9] = [self.r11, self.r12, self.r13, self.r21, self.r22, self.r23, self.r31, self.r32, self.r33]
        if self.SHELL:
            s[9:13] = [self.m11, self.m12, self.m21, self.m22]
        else:
            s[9:13] = [1., 0., 0., 1.]
        if self.KIND != _cabi.SPRING:
            s[13] = self._geo()
            s[14:14 + 3 * self._nn] = self.probe.xe
        s[26:26 + 6 * self._nn] = self.probe.ue
        return s

    def _cs(self):
        return [int(getattr(self, "c%d" % (a + 1))) for a in range(self._nn)]

    def _local(self, arr, per_node, div):
        """Gather the element's own nodes from a global array (c_a//div + i, quad4.pyx:526)."""
        out = np.empty(per_node * self._nn)
        for a, c in enumerate(self._cs()):
            out[per_node * a:per_node * (a + 1)] = arr[c // div:c // div + per_node]
        return out

    def _props_row(self, prop):
        if self.KIND == _cabi.SPRING:
            return None
        fields, stride = (_SHELL_FIELDS, _cabi.SHELLPROP_STRIDE) if self.SHELL else (_BEAM_FIELDS, _cabi.BEAMPROP_STRIDE)
        row = np.zeros(stride)
        for j, f in enumerate(fields):
            row[j] = getattr(prop, f)
        return row

    def _eparam(self, hg=None):
        ep = np.zeros(_cabi.EPARAM_STRIDE)
        if self.KIND == _cabi.SPRING:
            ep[:6] = [self.kxe, self.kye, self.kze, self.krxe, self.krye, self.krze]
        elif self.SHELL:
            ep[0] = self.K6ROT
            ep[1] = getattr(self, "alpha_shear_locking", 0.7)
            ep[2:7] = 1. if hg is None else hg
        return ep

    def _take_state(self, s, rot=False, xe=False, ue=False):
        if rot:
            (self.r11, self.r12, self.r13, self.r21, self.r22, self.r23,
             self.r31, self.r32, self.r33) = [float(v) for v in s[0:9]]
            if self.SHELL:
                self.m11, self.m12, self.m21, self.m22 = [float(v) for v in s[9:13]]
        if xe:
            self.probe.xe[:] = s[14:14 + 3 * self._nn]
            self._set_geo(float(s[13]))
        if ue:
            self.probe.ue[:] = s[26:26 + 6 * self._nn]

    # ---- reference methods ------------------------------------------------------------------
    def update_probe_ue(self, u):
        """probe.ue = R^T u (e.g. quad4.pyx:627)."""
        _check_array(u, np.float64, "u")
        ul = self._local(u, 6, 1)
        ep = self._eparam() if self.KIND == _cabi.SPRING else None
        out = self._eval_state(self._host_batch(None, self._state(), _cabi.STATE_REFRESH_UE, None, ul, ep))
        self._take_state(out, ue=True)

    def _eval_state(self, batch):
        out = np.zeros(_cabi.STATE_STRIDE)
        _ctx().eval_state(batch, out.ctypes.data, host=True)
        return out

    def update_probe_xe(self, x):
        """probe.xe = R^T x and area/length (e.g. quad4.pyx:682)."""
        _check_array(x, np.float64, "x")
        xl = self._local(x, 3, 2)
        out = self._eval_state(self._host_batch(None, self._state(), _cabi.STATE_REFRESH_XE, xl, None, None))
        self._take_state(out, xe=True)

    def _host_batch(self, prop, state, flags, x, u, ep, evec=None, mtype=0, stress=(0., 0., 0.), conn=None):
        """pf3_batch of ONE element over HOST arrays (kept alive on self until the call returns)."""
        if prop is None and self.KIND != _cabi.SPRING:
            prop = np.zeros(_cabi.SHELLPROP_STRIDE)
        if conn is None:
            conn = self._conn                      # local node numbers 0..nn-1: x / u hold this element's nodes only
        hold = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (prop, state, x, u, ep, evec)]
        self._keep = (hold, conn)
        prop, state, x, u, ep, evec = hold
        return _cabi.Batch(self.KIND, 1, self._nn, conn.ctypes.data, _p(x), _p(u), _p(prop), 0, 1, _p(evec), 0,
                           _p(ep), _p(state), mtype, stress, flags)

    def _run(self, what, prop, coo, mtype=0, stress=(0., 0., 0.), hg=None, values_only=False, fint=None, state=None):
        """One batch-of-one evaluation from the object's current state: ONE pf3_eval_host call."""
        ctx = _ctx()
        gconn = None
        if coo is not None:
            # the element's own global node positions, so that the COO indices come out global (x / u are not read
            # when a state is given)
            gconn = np.array([c // 6 for c in self._cs()], dtype=np.int64)
        b = self._host_batch(self._props_row(prop) if prop is not None else None,
                             self._state() if state is None else state, 0, None, None, self._eparam(hg), mtype=mtype,
                             stress=stress, conn=gconn)
        outs = [None, None, None]
        if coo is not None:
            which, (r, c, v, init_k, size) = coo
            if init_k < 0 or init_k + size > v.shape[0] or (not values_only and
                                                            (init_k + size > r.shape[0] or init_k + size > c.shape[0])):
                raise ValueError("COO arrays too short for init_k + SPARSE_SIZE")
            outs[which] = _cabi.Coo(0 if values_only else r.ctypes.data, 0 if values_only else c.ctypes.data,
                                    v.ctypes.data, init_k, 1)
        fl = None
        if fint is not None:
            fl = np.zeros(6 * self._nn)
        ctx.eval(b, what, outs[0], outs[1], outs[2], 0 if fl is None else fl.ctypes.data, host=True)
        if fint is not None:
            for a, cpos in enumerate(self._cs()):
                fint[cpos:cpos + 6] += fl[6 * a:6 * a + 6]

    def _finte(self, prop, hg=None):
        b = self._host_batch(self._props_row(prop) if prop is not None else None, self._state(), 0, None, None,
                             self._eparam(hg))
        out = np.zeros(6 * self._nn)
        _ctx().eval_finte(b, out.ctypes.data, host=True)
        self.probe.finte[:] = out


def _coo_args(KCr, KCc, KCv, init_k, size, names):
    _check_array(KCr, np.int64, names[0])
    _check_array(KCc, np.int64, names[1])
    _check_array(KCv, np.float64, names[2])
    return (KCr, KCc, KCv, int(init_k), size)


# ------------------------------------------------------------------------------------ shells
class _Shell(_Element):
    SHELL = True

    def __init__(self, probe):
        super().__init__(probe)
        self.init_k_KA_beta = 0
        self.init_k_KA_gamma = 0
        self.init_k_CA = 0
        self.area = 0.
        self.K6ROT = 100.
        self.m11, self.m12, self.m21, self.m22 = 1., 0., 0., 1.

    def update_rotation_matrix(self, x, xmati=0., xmatj=0., xmatk=0.):
        _check_array(x, np.float64, "x")
        xl = self._local(x, 3, 2)
        ep = self._eparam()
        ep[7] = 1.
        ep[8:12] = [self.m11, self.m12, self.m21, self.m22]
        b = self._host_batch(None, None, 0, xl, None, ep, evec=np.array([xmati, xmatj, xmatk], float))
        self._take_state(self._eval_state(b), rot=True)

    def update_area(self):
        # area is refreshed together with xe (update_probe_xe calls update_area, quad4.pyx:730);
        # calling it alone re-derives it from the probe's current xe with R = I
        st = self._state()
        st[0:9] = [1., 0., 0., 0., 1., 0., 0., 0., 1.]
        out = self._eval_state(self._host_batch(None, st, _cabi.STATE_REFRESH_XE, self.probe.xe.copy(), None, None))
        self.area = float(out[13])

    def update_probe_finte(self, prop):
        self._finte(prop)

    def update_KC0(self, KC0r, KC0c, KC0v, prop, update_KC0v_only=0):
        size = _cabi.sparse_size(self.KIND, _cabi.MAT_KC0)
        self._run(_cabi.KC0, prop, (0, _coo_args(KC0r, KC0c, KC0v, self.init_k_KC0, size, ("KC0r", "KC0c", "KC0v"))),
                  values_only=bool(update_KC0v_only))

    def update_fint(self, fint, prop):
        _check_array(fint, np.float64, "fint")
        self._finte(prop)
        self._run(_cabi.FINT, prop, None, fint=fint)

    def _kg_values_only(self, flag):
        return bool(flag)

    def update_KG(self, KGr, KGc, KGv, prop, update_KGv_only=0):
        size = _cabi.sparse_size(self.KIND, _cabi.MAT_KG)
        self._run(_cabi.KG, prop, (1, _coo_args(KGr, KGc, KGv, self.init_k_KG, size, ("KGr", "KGc", "KGv"))),
                  values_only=self._kg_values_only(update_KGv_only))

    def update_KG_given_stress(self, Nxx, Nyy, Nxy, KGr, KGc, KGv, update_KGv_only=0):
        size = _cabi.sparse_size(self.KIND, _cabi.MAT_KG)
        self._run(_cabi.KG_STRESS, None, (1, _coo_args(KGr, KGc, KGv, self.init_k_KG, size, ("KGr", "KGc", "KGv"))),
                  stress=(float(Nxx), float(Nyy), float(Nxy)), values_only=self._kg_values_only(update_KGv_only))

    def update_M(self, Mr, Mc, Mv, prop, mtype=0):
        size = _cabi.sparse_size(self.KIND, _cabi.MAT_M)
        self._run(_cabi.M, prop, (2, _coo_args(Mr, Mc, Mv, self.init_k_M, size, ("Mr", "Mc", "Mv"))), mtype=int(mtype))


class _QuadAero:
    """update_KA_beta / update_KA_gamma / update_CA of Quad4 and Quad4R (quad4.pyx:9491, 10312, 11115;
    quad4r.pyx:12789, 13605, 14403): piston-theory aerodynamic matrices, indices always written, values
    accumulated, from the object's current r11..r33 and probe.xe."""

    def _aero(self, which, r, c, v, init_k, names):
        ctx = _ctx()
        r, c, v, init_k, size = _coo_args(r, c, v, init_k, 144, names)
        if init_k < 0 or init_k + size > min(r.shape[0], c.shape[0], v.shape[0]):
            raise ValueError("COO arrays too short for init_k + SPARSE_SIZE")
        gconn = np.array([cc // 6 for cc in self._cs()], dtype=np.int64)
        b = self._host_batch(None, self._state(), 0, None, None, None, conn=gconn)
        outs = [None, None, None]
        outs[which] = _cabi.Coo(r.ctypes.data, c.ctypes.data, v.ctypes.data, init_k, 1)
        ctx.eval_aero(b, (_cabi.KA_BETA, _cabi.KA_GAMMA, _cabi.CA)[which], outs[0], outs[1], outs[2], host=True)

    def update_KA_beta(self, KA_betar, KA_betac, KA_betav):
        self._aero(0, KA_betar, KA_betac, KA_betav, self.init_k_KA_beta, ("KA_betar", "KA_betac", "KA_betav"))

    def update_KA_gamma(self, KA_gammar, KA_gammac, KA_gammav):
        self._aero(1, KA_gammar, KA_gammac, KA_gammav, self.init_k_KA_gamma, ("KA_gammar", "KA_gammac", "KA_gammav"))

    def update_CA(self, CAr, CAc, CAv):
        self._aero(2, CAr, CAc, CAv, self.init_k_CA, ("CAr", "CAc", "CAv"))


class Quad4Data(_Data):
    KIND = _cabi.QUAD4

    def __init__(self):
        super().__init__()
        self.KA_BETA_SPARSE_SIZE = 144
        self.KA_GAMMA_SPARSE_SIZE = 144
        self.CA_SPARSE_SIZE = 144


class Quad4Probe(_Probe):
    """Quad4Probe (quad4.pyx:183-395): besides xe/ue/finte it carries the 11 strain-interpolation rows
    filled by ``update_BL(xi, eta)`` and the local stiffness ``KC0ve`` of the last element evaluated."""
    KIND = _cabi.QUAD4
    _BL = ("BLexx", "BLeyy", "BLgxy", "BLkxx", "BLkyy", "BLkxy", "BLgyz_grad", "BLgyz_rot", "BLgxz_grad",
           "BLgxz_rot", "BLdrilling")

    def __init__(self):
        super().__init__()
        self._KC0ve = np.zeros(576)
        self._KC0ve_pending = None
        for n in self._BL:
            setattr(self, n, np.zeros(24))

    @property
    def KC0ve(self):
        """Local 24x24 stiffness of the element last passed through update_KC0 / update_probe_finte
        (quad4.pyx:755, :1196).  The reference fills it as a by-product of those calls; here the inputs of that
        call are remembered and the block is evaluated (the same kernel with R = I) when it is first read."""
        if self._KC0ve_pending is not None:
            elem, state, prop_row, ep = self._KC0ve_pending
            self._KC0ve_pending = None
            st = state.copy()
            st[0:9] = [1., 0., 0., 0., 1., 0., 0., 0., 1.]
            v = np.zeros(576)
            b = elem._host_batch(prop_row, st, 0, None, None, ep)
            _ctx().eval(b, _cabi.KC0, _cabi.Coo(0, 0, v.ctypes.data, 0, 0), None, None, 0, host=True)
            self._KC0ve[:] = v
        return self._KC0ve

    def update_BL(self, xi, eta):
        out = np.zeros(264)
        xe = np.ascontiguousarray(self.xe, dtype=np.float64)
        _ctx().quad4_update_BL(1, xe.ctypes.data, float(xi), float(eta), out.ctypes.data, host=True)
        rows = out.reshape(11, 24)
        for i, n in enumerate(self._BL):
            getattr(self, n)[:] = rows[i]


class Quad4(_QuadAero, _Shell):
    KIND = _cabi.QUAD4

    def _remember_KC0ve(self, prop):
        self.probe._KC0ve_pending = (self, self._state(), self._props_row(prop), self._eparam())

    def update_KC0(self, KC0r, KC0c, KC0v, prop, update_KC0v_only=0):
        super().update_KC0(KC0r, KC0c, KC0v, prop, update_KC0v_only)
        self._remember_KC0ve(prop)

    def _finte(self, prop, hg=None):   # update_probe_finte recomputes KC0ve in the reference (quad4.pyx:1196)
        super()._finte(prop, hg)
        self._remember_KC0ve(prop)

    # Quad4.update_KG / update_KG_given_stress take no update_KGv_only (quad4.pyx:1365, 2259):
    # indices are always written
    def update_KG(self, KGr, KGc, KGv, prop):
        super().update_KG(KGr, KGc, KGv, prop, 0)

    def update_KG_given_stress(self, Nxx, Nyy, Nxy, KGr, KGc, KGv):
        super().update_KG_given_stress(Nxx, Nyy, Nxy, KGr, KGc, KGv, 0)


class Quad4RData(Quad4Data):
    KIND = _cabi.QUAD4R


class Quad4RProbe(_Probe):
    KIND = _cabi.QUAD4R


class Quad4R(_QuadAero, _Shell):
    """hgfactor_u .. hgfactor_ry are ordinary positional-or-keyword parameters with default 1., as in
    quad4r.pyx:549-556 (update_probe_finte), :1145-1156 (update_KC0), :4620-4627 (update_fint)."""
    KIND = _cabi.QUAD4R

    def update_probe_finte(self, prop, hgfactor_u=1., hgfactor_v=1., hgfactor_w=1., hgfactor_rx=1., hgfactor_ry=1.):
        self._finte(prop, np.array([hgfactor_u, hgfactor_v, hgfactor_w, hgfactor_rx, hgfactor_ry], float))

    def update_KC0(self, KC0r, KC0c, KC0v, prop, update_KC0v_only=0, hgfactor_u=1., hgfactor_v=1., hgfactor_w=1.,
                   hgfactor_rx=1., hgfactor_ry=1.):
        size = _cabi.sparse_size(self.KIND, _cabi.MAT_KC0)
        self._run(_cabi.KC0, prop, (0, _coo_args(KC0r, KC0c, KC0v, self.init_k_KC0, size, ("KC0r", "KC0c", "KC0v"))),
                  hg=np.array([hgfactor_u, hgfactor_v, hgfactor_w, hgfactor_rx, hgfactor_ry], float),
                  values_only=bool(update_KC0v_only))

    def update_fint(self, fint, prop, hgfactor_u=1., hgfactor_v=1., hgfactor_w=1., hgfactor_rx=1., hgfactor_ry=1.):
        _check_array(fint, np.float64, "fint")
        h = np.array([hgfactor_u, hgfactor_v, hgfactor_w, hgfactor_rx, hgfactor_ry], float)
        self._finte(prop, h)
        self._run(_cabi.FINT, prop, None, hg=h, fint=fint)


class Tria3RData(_Data):
    KIND = _cabi.TRIA3R


class Tria3RProbe(_Probe):
    KIND = _cabi.TRIA3R


class Tria3R(_Shell):
    KIND = _cabi.TRIA3R

    def __init__(self, probe):
        super().__init__(probe)
        self.alpha_shear_locking = 0.7
        del self.init_k_KA_beta, self.init_k_KA_gamma, self.init_k_CA


# ------------------------------------------------------------------------------------- lines
class _Line(_Element):
    def __init__(self, probe):
        super().__init__(probe)
        self.length = 0.
        self.vxyi = self.vxyj = self.vxyk = 0.

    def update_length(self):
        st = self._state()
        st[0:9] = [1., 0., 0., 0., 1., 0., 0., 0., 1.]
        out = self._eval_state(self._host_batch(self._dummy_prop(), st, _cabi.STATE_REFRESH_XE, self.probe.xe.copy(),
                                                None, None))
        self.length = float(out[13])

    def _dummy_prop(self):
        return np.zeros(_cabi.BEAMPROP_STRIDE)

    def _rot(self, x, vxy):
        xl = self._local(x, 3, 2)
        b = self._host_batch(self._dummy_prop(), None, 0, xl, None, None, evec=vxy)
        self._take_state(self._eval_state(b), rot=True)

    def update_probe_finte(self, prop):
        self._finte(prop)

    def update_KC0(self, KC0r, KC0c, KC0v, prop, update_KC0v_only=0):
        size = _cabi.sparse_size(self.KIND, _cabi.MAT_KC0)
        self._run(_cabi.KC0, prop, (0, _coo_args(KC0r, KC0c, KC0v, self.init_k_KC0, size, ("KC0r", "KC0c", "KC0v"))),
                  values_only=bool(update_KC0v_only))

    def update_fint(self, fint, prop):
        _check_array(fint, np.float64, "fint")
        self._finte(prop)
        self._run(_cabi.FINT, prop, None, fint=fint)

    def update_KG(self, KGr, KGc, KGv, prop, update_KGv_only=0):
        size = _cabi.sparse_size(self.KIND, _cabi.MAT_KG)
        if size == 0:
            raise AttributeError("this element has no geometric stiffness matrix")
        self._run(_cabi.KG, prop, (1, _coo_args(KGr, KGc, KGv, self.init_k_KG, size, ("KGr", "KGc", "KGv"))),
                  values_only=bool(update_KGv_only))

    def update_M(self, Mr, Mc, Mv, prop, mtype=0):
        size = _cabi.sparse_size(self.KIND, _cabi.MAT_M)
        self._run(_cabi.M, prop, (2, _coo_args(Mr, Mc, Mv, self.init_k_M, size, ("Mr", "Mc", "Mv"))), mtype=int(mtype))


class BeamCData(_Data):
    KIND = _cabi.BEAMC


class BeamCProbe(_Probe):
    KIND = _cabi.BEAMC


class BeamC(_Line):
    KIND = _cabi.BEAMC

    def update_rotation_matrix(self, vxyi, vxyj, vxyk, x):
        _check_array(x, np.float64, "x")
        self.vxyi, self.vxyj, self.vxyk = float(vxyi), float(vxyj), float(vxyk)
        self._rot(x, np.array([vxyi, vxyj, vxyk], float))


class BeamLRData(_Data):
    KIND = _cabi.BEAMLR


class BeamLRProbe(_Probe):
    KIND = _cabi.BEAMLR


class BeamLR(BeamC):
    KIND = _cabi.BEAMLR


class TrussData(_Data):
    KIND = _cabi.TRUSS


class TrussProbe(_Probe):
    KIND = _cabi.TRUSS


class Truss(_Line):
    KIND = _cabi.TRUSS

    def update_rotation_matrix(self, x):
        _check_array(x, np.float64, "x")
        self._rot(x, None)

    @property
    def update_KG(self):  # the reference Truss defines no update_KG (truss.pyx:16-17)
        raise AttributeError("Truss has no update_KG")


class SpringData(_Data):
    KIND = _cabi.SPRING


class SpringProbe(_Probe):
    KIND = _cabi.SPRING

    def __init__(self):
        super().__init__()
        self.xe = np.zeros(6)


class Spring(_Element):
    KIND = _cabi.SPRING

    def __init__(self, probe):
        super().__init__(probe)
        self.kxe = self.kye = self.kze = 0.
        self.krxe = self.krye = self.krze = 0.
        self.vxyi = self.vxyj = self.vxyk = 0.
        self.r11 = self.r22 = self.r33 = 1.   # default R = I (spring.pyx:138-144)

    def update_rotation_matrix(self, xi, xj, xk, vxyi, vxyj, vxyk):
        self.vxyi, self.vxyj, self.vxyk = float(vxyi), float(vxyj), float(vxyk)
        b = self._host_batch(None, None, 0, None, None, self._eparam(),
                             evec=np.array([xi, xj, xk, vxyi, vxyj, vxyk], float))
        self._take_state(self._eval_state(b), rot=True)

    def update_probe_finte(self):
        self._finte(None)

    def update_KC0(self, KC0r, KC0c, KC0v, update_KC0v_only=0):
        self._run(_cabi.KC0, None, (0, _coo_args(KC0r, KC0c, KC0v, self.init_k_KC0, 72, ("KC0r", "KC0c", "KC0v"))),
                  values_only=bool(update_KC0v_only))

    def update_fint(self, fint):
        _check_array(fint, np.float64, "fint")
        self._finte(None)
        self._run(_cabi.FINT, None, None, fint=fint)
