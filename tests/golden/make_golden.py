"""Generates the committed golden fixtures from the CPU oracle (the reference itself cannot be built
here: Eigen 3.3.8 / Qt5 / glog are absent).  Run from the repo root:

    python tests/golden/make_golden.py

Fixtures: step outputs of the BASELINE.json scenes (same seeded arrays the GPU batch consumes).
The oracle is pinned separately against the reference's literal vectors and property tests
(tests/test_oracle_*.py); these files freeze its step outputs so that (a) the oracle cannot drift
silently and (b) the GPU path can be checked on a box where only the fixtures travel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O                      # noqa: E402
import eggshell_b200.scenes as S                      # noqa: E402
from tests.helpers import oracle_world                # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def pack_step(W):
    p, R, v, w = W.bodies()
    c = W.contacts()
    lam, rhs, rs = W.solution()
    st = W.stats()
    ph = W.pair_hits()
    return dict(p=p, R=R, v=v, w=w, c_i0=c["i0"], c_i1=c["i1"], c_pos=c["pos"], c_nrm=c["nrm"], c_depth=c["depth"],
                c_code=c["code"], lam=lam, rhs=rhs, row_state=rs, ground=W.ground_counts(),
                hit_i=ph["i"], hit_j=ph["j"], hit_code=ph["code"], hit_count=ph["count"],
                stats=np.array([st[k] for k in ("n_contacts_raw", "n_contacts", "n_rows", "n_pair_hits", "sweeps", "pivots",
                                                "cfm_applied", "status")]), residual=np.array(st["residual"]))


def scene_steps(name, scene, nworlds, nsteps, **params):
    out = {}
    for wi in range(nworlds):
        W, st = oracle_world(scene, wi, **params)
        assert st == 0
        for s in range(nsteps):
            W.step(scene["dt"])
            for k, v in pack_step(W).items():
                out[f"w{wi}_s{s}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "written", len(out), "arrays")


def chain10():
    """C1: Chain(10,(2,2,1)), dt = 0.001, dense Murty (model.cc:28,102-115)."""
    W = O.World()
    W.build_chain(10, [2, 2, 1])
    W.set_params(solver=O.SOLVER_DENSE_MURTY)
    assert W.init() == 0
    out = {}
    for s in range(1, 501):
        W.step(0.001)
        if s % 100 == 0 or s in (1, 2, 399, 400, 401):
            for k, v in pack_step(W).items():
                out[f"s{s}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "chain10_dense.npz"), **out)
    print("chain10_dense written", len(out))


if __name__ == "__main__":
    chain10()
    scene_steps("stack10_pgs", S.stack10(2, seed=1000), 2, 2, solver=O.SOLVER_PGS)
    scene_steps("pile64_pgs_k50", S.pile64(1, seed=3000), 1, 2, solver=O.SOLVER_PGS, k_max=50)
    scene_steps("pile64_pgs_k500", S.pile64(1, seed=3000), 1, 2, solver=O.SOLVER_PGS, k_max=500)
    scene_steps("legged20_pgs_k100", S.legged20(2, seed=5000), 2, 2, solver=O.SOLVER_PGS, k_max=100)
    scene_steps("chain32_pgs_k100", S.chain32(1, seed=4000, anchor_z=0.2), 1, 2, solver=O.SOLVER_PGS, k_max=100)
    scene_steps("cairn4_pgs", S.cairn(2, rocks=4, zb=(0.2, 0.6), seed=11), 2, 5, solver=O.SOLVER_PGS)
