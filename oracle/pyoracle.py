"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end to ``oracle/liboracle.so`` (the Eigen-free CPU restatement of eggshell's step,
see ``oracle/eggshell_oracle.cc``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module; the product
package ``eggshell_b200`` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SOLVER_DENSE_MURTY, SOLVER_PGS, SOLVER_JACOBI, SOLVER_SOR = 0, 1, 2, 3
CFM_AUTO, CFM_ALWAYS, CFM_NEVER = 0, 1, 2
QUIRK_GS_BOUNDS_SHIFT, QUIRK_DENSE_IGNORES_BOUNDS = 1, 2
QUIRKS_REFERENCE = 3
IT_JACOBI, IT_GS, IT_SOR = 0, 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_bp = C.POINTER(C.c_ubyte)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cc", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_world_create.restype = C.c_void_p
        _LIB.orc_condition_number.restype = C.c_double
        _LIB.orc_batch_step.restype = C.c_double
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def _pi(a):
    return a.ctypes.data_as(_ip)


def _pb(a):
    return a.ctypes.data_as(_bp)


# ---- geometry ---------------------------------------------------------------------------------
def collide_box_ground(c, R, side):
    c, R, side = _d(c), _d(R), _d(side)
    out = np.zeros((8, 7))
    n = lib().orc_collide_box_ground(_p(c), _p(R), _p(side), _p(out))
    return out[:n]


def collide_boxes(c1, R1, h1, c2, R2, h2):
    """Returns (hit, code, depth, axis, contacts[n,7]) for half-sides h1, h2."""
    a = [_d(x) for x in (c1, R1, h1, c2, R2, h2)]
    info = np.zeros(4)
    code = C.c_int(0)
    out = np.zeros((32, 7))
    n = lib().orc_collide_boxes(*[_p(x) for x in a], _p(info), C.byref(code), _p(out), 32)
    return n > 0, code.value, info[0], info[1:4].copy(), out[:n].copy()


def boxes_separated(c1, R1, h1, c2, R2, h2):
    a = [_d(x) for x in (c1, R1, h1, c2, R2, h2)]
    return bool(lib().orc_boxes_separated(*[_p(x) for x in a]))


def line_closest_approach(pa, ua, pb, ub):
    a = [_d(x) for x in (pa, ua, pb, ub)]
    al, be = C.c_double(), C.c_double()
    lib().orc_line_closest_approach(*[_p(x) for x in a], C.byref(al), C.byref(be))
    return al.value, be.value


def intersect_line_segment_and_line(p1, p2, n, d):
    a = [_d(x) for x in (p1, p2, n)]
    p = np.zeros(2)
    r = lib().orc_intersect_line_segment_and_line(*[_p(x) for x in a], C.c_double(d), _p(p))
    return bool(r), p


def clip_polygon(poly, n, d):
    poly, n = _d(poly), _d(n)
    out = np.zeros((64, 2))
    k = lib().orc_clip_polygon(_p(poly), len(poly), _p(n), C.c_double(d), _p(out), 64)
    return out[:k].copy()


def intersect_box_rect(cB, RB, hB, cR, RR, hR):
    a = [_d(x) for x in (cB, RB, hB, cR, RR, hR)]
    out = np.zeros((64, 2))
    k = lib().orc_intersect_box_rect(*[_p(x) for x in a], _p(out), 64)
    return out[:k].copy()


# ---- utils ------------------------------------------------------------------------------------
def cross_mat(a):
    out = np.zeros((3, 3))
    lib().orc_cross_mat(_p(_d(a)), _p(out))
    return out


def w_to_q_matrix(w, dt):
    out = np.zeros((3, 3))
    lib().orc_w_to_q_matrix(_p(_d(w)), C.c_double(dt), _p(out))
    return out


def align_vectors(a, b):
    out = np.zeros((3, 3))
    lib().orc_align_vectors(_p(_d(a)), _p(_d(b)), _p(out))
    return out


def _mask(m):
    return np.ascontiguousarray(np.asarray(m) != 0, dtype=np.uint8)


def select_submatrix(A, ri, ci):
    A, ri, ci = _d(A), _mask(ri), _mask(ci)
    out = np.zeros((int(ri.sum()), int(ci.sum())))
    lib().orc_select_submatrix(_p(A), A.shape[0], _pb(ri), _pb(ci), _p(out))
    return out


def update_submatrix(A, ri, ci, m):
    A, ri, ci, m = _d(A).copy(), _mask(ri), _mask(ci), _d(m)
    lib().orc_update_submatrix(_p(A), A.shape[0], _pb(ri), _pb(ci), _p(m), m.shape[0], m.shape[1])
    return A


def select_subvector(v, ind):
    v, ind = _d(v), _mask(ind)
    out = np.zeros(len(v))
    k = lib().orc_select_subvector(_p(v), len(v), _pb(ind), _p(out))
    return out[:k].copy()


def update_subvector(v, ind, n):
    v, ind = _d(v).copy(), _mask(ind)
    if np.isscalar(n):
        lib().orc_update_subvector_scalar(_p(v), len(v), _pb(ind), C.c_double(n))
    else:
        n = _d(n)
        lib().orc_update_subvector(_p(v), len(v), _pb(ind), _p(n), len(n))
    return v


def ldlt_solve(A, b):
    A, b = _d(A), _d(b)
    x = np.zeros(len(b))
    lib().orc_ldlt_solve(_p(A), len(b), _p(b), _p(x))
    return x


def lu_inverse(A):
    A = _d(A)
    out = np.zeros_like(A)
    lib().orc_lu_inverse(_p(A), A.shape[0], _p(out))
    return out


def condition_number(A):
    A = _d(A)
    return lib().orc_condition_number(_p(A), A.shape[0], A.shape[1])


# ---- dense LCP --------------------------------------------------------------------------------
def check_murty_solution(A, b, x, w, S, err=0.0):
    A, b, x, w, S = _d(A), _d(b), _d(x), _d(w), _mask(S).copy()
    r = lib().orc_check_murty_solution(_p(A), _p(b), _p(x), _p(w), len(b), _pb(S), C.c_double(err))
    return bool(r), S


def murty(A, b, lo=None, hi=None):
    A, b = _d(A), _d(b)
    n = len(b)
    lo = _d(np.zeros(n) if lo is None else np.broadcast_to(lo, (n,)))
    hi = _d(np.full(n, np.inf) if hi is None else np.broadcast_to(hi, (n,)))
    x, w = np.zeros(n), np.zeros(n)
    it = C.c_int(0)
    S = np.zeros(n, dtype=np.uint8)
    ok = lib().orc_murty(_p(A), _p(b), n, _p(lo), _p(hi), _p(x), _p(w), C.byref(it), _pb(S))
    return bool(ok), x, w, it.value, S


def mixed_solver(A, b, Cmask, lo, hi, honour_bounds=False):
    A, b, Cm, lo, hi = _d(A), _d(b), _mask(Cmask), _d(lo), _d(hi)
    n = len(b)
    x, w = np.zeros(n), np.zeros(n)
    it = C.c_int(0)
    ok = lib().orc_mixed_solver(_p(A), _p(b), n, _pb(Cm), _p(lo), _p(hi), int(honour_bounds), _p(x), _p(w), C.byref(it))
    return bool(ok), x, w, it.value


def dense_iteration(A, b, itype, Cmask=None, lo=None, hi=None, k_max=500, tol=1e-9):
    A, b = _d(A), _d(b)
    n = len(b)
    Cm = _mask(np.ones(n) if Cmask is None else Cmask)
    lo = _d(np.zeros(n) if lo is None else lo)
    hi = _d(np.zeros(n) if hi is None else hi)
    x = np.zeros(n)
    sweeps = lib().orc_dense_iteration(_p(A), _p(b), n, itype, _pb(Cm), _p(lo), _p(hi), k_max, C.c_double(tol), _p(x))
    return x, sweeps


# ---- worlds -----------------------------------------------------------------------------------
class World:
    """One reference ``Ensemble`` (ensembles.h:25-177) on the CPU oracle."""

    def __init__(self):
        self.h = C.c_void_p(lib().orc_world_create())

    def __del__(self):
        try:
            lib().orc_world_destroy(self.h)
        except Exception:
            pass

    def set_params(self, erp=0.2, cfm=0.01, min_dist=1e-6, tol=1e-9, k_max=500, gravity=(0, 0, -9.8),
                   solver=SOLVER_DENSE_MURTY, cfm_mode=CFM_AUTO, quirks=QUIRKS_REFERENCE):
        g = _d(gravity)
        lib().orc_world_set_params(self.h, C.c_double(erp), C.c_double(cfm), C.c_double(min_dist), C.c_double(tol),
                                   int(k_max), _p(g), int(solver), int(cfm_mode), int(quirks))

    def set_bodies(self, p, R, v, w, m, I, side=None):
        p, R, v, w, m, I = [_d(x) for x in (p, R, v, w, m, I)]
        n = len(m)
        s = None if side is None else _d(side)
        lib().orc_world_set_bodies(self.h, n, _p(p), _p(R), _p(v), _p(w), _p(m), _p(I), None if s is None else _p(s))

    def set_shapes(self, shape, dims):
        shape = np.ascontiguousarray(shape, dtype=np.int32)
        dims = _d(dims)
        lib().orc_world_set_shapes(self.h, _pi(shape), _p(dims))

    def set_state(self, p, R, v, w):
        p, R, v, w = [_d(x) for x in (p, R, v, w)]
        lib().orc_world_set_state(self.h, _p(p), _p(R), _p(v), _p(w))

    def set_joints(self, i0, i1, c0, c1):
        i0 = np.ascontiguousarray(i0, dtype=np.int32)
        i1 = np.ascontiguousarray(i1, dtype=np.int32)
        c0, c1 = _d(c0), _d(c1)
        lib().orc_world_set_joints(self.h, len(i0), _pi(i0), _pi(i1), _p(c0), _p(c1))

    def set_fext(self, f):
        f = _d(f)
        lib().orc_world_set_fext(self.h, _p(f))

    def build_chain(self, links, anchor):
        lib().orc_world_build_chain(self.h, int(links), _p(_d(anchor)))

    def build_cairn(self, rocks, xb, yb, zb):
        lib().orc_world_build_cairn(self.h, int(rocks), _p(_d(xb)), _p(_d(yb)), _p(_d(zb)))

    def init(self):
        return lib().orc_world_init(self.h)

    def init_stabilize(self, max_steps=100):
        e = C.c_double(0)
        steps = lib().orc_world_init_stabilize_n(self.h, int(max_steps), C.byref(e))
        return steps, e.value

    def post_stabilize(self, max_steps=500):
        e = C.c_double(0)
        steps = lib().orc_world_post_stabilize(self.h, int(max_steps), C.byref(e))
        return steps, e.value

    def step(self, dt):
        return lib().orc_world_step(self.h, C.c_double(dt))

    def update_contacts(self, dedupe=True):
        lib().orc_world_update_contacts(self.h, int(dedupe))

    @property
    def n(self):
        return lib().orc_world_n(self.h)

    @property
    def n_joints(self):
        return lib().orc_world_n_joints(self.h)

    @property
    def n_contacts(self):
        return lib().orc_world_n_contacts(self.h)

    def bodies(self):
        n = self.n
        p, R, v, w = np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3)), np.zeros((n, 3))
        lib().orc_world_get_bodies(self.h, _p(p), _p(R), _p(v), _p(w))
        return p, R, v, w

    def static(self):
        n = self.n
        m, I, ml, ma, f = np.zeros(n), np.zeros((n, 3, 3)), np.zeros(n), np.zeros((n, 3, 3)), np.zeros((n, 6))
        lib().orc_world_get_static(self.h, _p(m), _p(I), _p(ml), _p(ma), _p(f))
        return m, I, ml, ma, f

    def joints(self):
        k = self.n_joints
        i0, i1 = np.zeros(k, dtype=np.int32), np.zeros(k, dtype=np.int32)
        c0, c1 = np.zeros((k, 3)), np.zeros((k, 3))
        lib().orc_world_get_joints(self.h, _pi(i0), _pi(i1), _p(c0), _p(c1))
        return i0, i1, c0, c1

    def contacts(self):
        k = self.n_contacts
        i0, i1, code = np.zeros(k, dtype=np.int32), np.zeros(k, dtype=np.int32), np.zeros(k, dtype=np.int32)
        pos, nrm, depth = np.zeros((k, 3)), np.zeros((k, 3)), np.zeros(k)
        lib().orc_world_get_contacts(self.h, _pi(i0), _pi(i1), _p(pos), _p(nrm), _p(depth), _pi(code))
        return dict(i0=i0, i1=i1, pos=pos, nrm=nrm, depth=depth, code=code)

    def pair_hits(self):
        cap = max(1, self.n * (self.n - 1) // 2)
        a = [np.zeros(cap, dtype=np.int32) for _ in range(4)]
        k = lib().orc_world_get_pair_hits(self.h, *[_pi(x) for x in a], cap)
        return dict(i=a[0][:k], j=a[1][:k], code=a[2][:k], count=a[3][:k])

    def ground_counts(self):
        out = np.zeros(self.n, dtype=np.int32)
        lib().orc_world_get_ground_counts(self.h, _pi(out))
        return out

    def stats(self):
        s = np.zeros(9, dtype=np.int32)
        r = C.c_double(0)
        lib().orc_world_get_stats(self.h, _pi(s), C.byref(r))
        keys = ["n_contacts_raw", "n_contacts", "n_rows", "n_pair_tests", "n_pair_hits", "sweeps", "pivots",
                "cfm_applied", "status"]
        d = {k: int(v) for k, v in zip(keys, s)}
        d["residual"] = r.value
        return d

    def solution(self):
        nr = lib().orc_world_get_solution(self.h, None, None, None)
        lam, rhs, st = np.zeros(nr), np.zeros(nr), np.zeros(nr, dtype=np.int32)
        lib().orc_world_get_solution(self.h, _p(lam), _p(rhs), _pi(st))
        return lam, rhs, st

    def rows(self):
        nc = self.n_joints + self.n_contacts
        J0, J1 = np.zeros((nc, 3, 6)), np.zeros((nc, 3, 6))
        typ = np.zeros(3 * nc, dtype=np.uint8)
        lo, hi, err = np.zeros(3 * nc), np.zeros(3 * nc), np.zeros(3 * nc)
        i0, i1 = np.zeros(nc, dtype=np.int32), np.zeros(nc, dtype=np.int32)
        lib().orc_world_get_rows(self.h, _p(J0), _p(J1), _pb(typ), _p(lo), _p(hi), _pi(i0), _pi(i1), _p(err))
        return dict(J0=J0, J1=J1, type=typ, lo=lo, hi=hi, i0=i0, i1=i1, err=err)

    def dense_A(self, cfm=0.0):
        nr = 3 * (self.n_joints + self.n_contacts)
        A = np.zeros((nr, nr))
        lib().orc_world_dense_A(self.h, C.c_double(cfm), _p(A))
        return A

    def sparse_product(self, op, x, cfm=0.0, scale=1.0):
        x = _d(x)
        out = np.zeros_like(x)
        lib().orc_world_sparse_product(self.h, int(op), _p(x), C.c_double(cfm), C.c_double(scale), _p(out))
        return out

    def sparse_solve(self, which, rhs, cfm=0.0, scale=1.0, quirk=True):
        rhs = _d(rhs)
        out = np.zeros_like(rhs)
        lib().orc_world_sparse_solve(self.h, int(which), _p(rhs), C.c_double(cfm), C.c_double(scale), int(quirk), _p(out))
        return out

    def sparse_iteration(self, itype, rhs, cfm=0.0, k_max=500, tol=1e-9, quirk=True):
        rhs = _d(rhs)
        x = np.zeros_like(rhs)
        sweeps = lib().orc_world_sparse_iteration(self.h, int(itype), _p(rhs), C.c_double(cfm), int(k_max),
                                                  C.c_double(tol), int(quirk), _p(x))
        return x, sweeps


def batch_step(worlds, dt, steps, nthreads):
    """Step every world `steps` times on `nthreads` host threads; returns (seconds, rows*sweeps, rows)."""
    arr = (C.c_void_p * len(worlds))(*[w.h for w in worlds])
    tot = np.zeros(2)
    sec = lib().orc_batch_step(arr, len(worlds), C.c_double(dt), int(steps), int(nthreads), _p(tot))
    return sec, tot[0], tot[1]


def hardware_concurrency():
    return lib().orc_hardware_concurrency()
