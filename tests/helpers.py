"""Shared helpers for the parity tests: build CPU-oracle worlds from the same scene arrays that
feed the GPU batch, and compare one step at a time from shared state (BASELINE.json north_star)."""
import numpy as np

from oracle import pyoracle as O


def oracle_world(scene, w, **params):
    W = O.World()
    W.set_bodies(scene["p"][w], scene["R"][w], scene["v"][w], scene["w"][w], scene["m"][w], scene["I"][w])
    if scene["nj"]:
        W.set_joints(scene["i0"], scene["i1"], scene["c0"][w], scene["c1"][w])
    W.set_params(**params)
    st = W.init()
    if "f_ext" in scene:
        W.set_fext(scene["f_ext"][w])
    return W, st


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(|b|_inf, floor): relative to the field's scale, absolute below `floor`."""
    a, b = np.asarray(a), np.asarray(b)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), floor))


def compare_step(batch, worlds, scene_W_indices, tol=1e-9, check_lambda=True):
    """After both sides stepped once from the same state: discrete outputs bit-exact, state within
    `tol` relative.  Returns a dict of worst errors."""
    p, R, v, w = batch.bodies()
    con = batch.contacts()
    st = batch.status()
    worst = dict(p=0.0, R=0.0, v=0.0, w=0.0, lam=0.0, geom=0.0)
    for k, wi in enumerate(scene_W_indices):
        ow = worlds[k]
        op, oR, ov, owv = ow.bodies()
        oc = ow.contacts()
        os_ = ow.stats()
        nc = os_["n_contacts"]
        assert con["count"][wi] == nc, f"world {wi}: contact count {con['count'][wi]} != oracle {nc}"
        assert st["n_contacts_raw"][wi] == os_["n_contacts_raw"], f"world {wi}: raw contact count"
        assert st["n_pair_hits"][wi] == os_["n_pair_hits"], f"world {wi}: colliding pair count"
        assert np.array_equal(con["i0"][wi, :nc], oc["i0"]) and np.array_equal(con["i1"][wi, :nc], oc["i1"]), \
            f"world {wi}: contact body indices / order differ"
        assert np.array_equal(con["code"][wi, :nc], oc["code"]), f"world {wi}: collision codes differ"
        worst["geom"] = max(worst["geom"], rel_err(con["pos"][wi, :nc], oc["pos"]), rel_err(con["nrm"][wi, :nc], oc["nrm"]),
                            rel_err(con["depth"][wi, :nc], oc["depth"]))
        worst["p"] = max(worst["p"], rel_err(p[wi], op))
        worst["R"] = max(worst["R"], rel_err(R[wi], oR))
        worst["v"] = max(worst["v"], rel_err(v[wi], ov))
        worst["w"] = max(worst["w"], rel_err(w[wi], owv))
        if check_lambda:
            lam, rhs, rs = ow.solution()
            nr = len(lam)
            worst["lam"] = max(worst["lam"], rel_err(con["lam"][wi, :nr], lam))
    for key in ("p", "R", "v", "w", "geom"):
        assert worst[key] <= tol, f"{key} mismatch {worst[key]:.3e} > {tol}"
    return worst
