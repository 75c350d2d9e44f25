"""Shared helpers for the parity tests: build CPU-oracle worlds from the same scene arrays that
feed the GPU batch, and compare one step at a time from shared state (BASELINE.json north_star)."""
import numpy as np

from oracle import pyoracle as O


def oracle_world(scene, w, **params):
    W = O.World()
    W.set_bodies(scene["p"][w], scene["R"][w], scene["v"][w], scene["w"][w], scene["m"][w], scene["I"][w])
    if "shape" in scene:
        W.set_shapes(scene["shape"][w], scene["dims"][w])
    if scene["nj"]:
        W.set_joints(scene["i0"], scene["i1"], scene["c0"][w], scene["c1"][w])
    W.set_params(**params)
    st = W.init()
    if "f_ext" in scene:
        W.set_fext(scene["f_ext"][w])
    return W, st


# Natural scale of every compared field: an error is relative to max(|reference|_inf, scale), so a
# field that happens to be near zero in one world is judged against the size it has in the scene,
# not against 1.0.  Lengths: the cube side (body.h:91); velocities: what gravity adds in the
# shortest time step used by any scene (9.8 x 1e-3, constants.h:6-8); angular velocities: the same
# over one cube side; multipliers: the weight of one body (rows are acceleration-level,
# ensembles.cc:569-570).
SCALE = dict(p=0.3, R=1.0, v=9.8e-3, w=9.8e-3 / 0.3, lam=9.8, pos=0.3, nrm=1.0, depth=0.3, resid=1e-3)


def rel_err(a, b, floor):
    """max |a-b| / max(|b|_inf, floor); `floor` is the field's natural scale (SCALE[...])."""
    a, b = np.asarray(a), np.asarray(b)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), floor))


def compare_step(batch, worlds, scene_W_indices, tol=1e-9, check_lambda=True, lam_tol=None):
    """After both sides stepped once from the same state: discrete outputs bit-exact, state and
    multipliers within `tol` relative (north_star: 1e-9 in FP64).  Returns a dict of worst errors."""
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
        worst["geom"] = max(worst["geom"], rel_err(con["pos"][wi, :nc], oc["pos"], SCALE["pos"]),
                            rel_err(con["nrm"][wi, :nc], oc["nrm"], SCALE["nrm"]),
                            rel_err(con["depth"][wi, :nc], oc["depth"], SCALE["depth"]))
        worst["p"] = max(worst["p"], rel_err(p[wi], op, SCALE["p"]))
        worst["R"] = max(worst["R"], rel_err(R[wi], oR, SCALE["R"]))
        worst["v"] = max(worst["v"], rel_err(v[wi], ov, SCALE["v"]))
        worst["w"] = max(worst["w"], rel_err(w[wi], owv, SCALE["w"]))
        if check_lambda:
            lam, rhs, rs = ow.solution()
            nr = len(lam)
            worst["lam"] = max(worst["lam"], rel_err(con["lam"][wi, :nr], lam, SCALE["lam"]))
    for key in ("p", "R", "v", "w", "geom"):
        assert worst[key] <= tol, f"{key} mismatch {worst[key]:.3e} > {tol}"
    if check_lambda:
        lt = tol if lam_tol is None else lam_tol
        assert worst["lam"] <= lt, f"lambda mismatch {worst['lam']:.3e} > {lt}"
    return worst
