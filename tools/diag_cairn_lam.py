"""Which world / step / row carries the largest multiplier mismatch in test_cairn_pgs_falling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eggshell_b200 as E
from tests.helpers import oracle_world

scene = E.scenes.cairn(16, rocks=4, zb=(0.2, 1.0), seed=11)
b = E.scenes.make_batch(scene, solver=E.SOLVER_PGS, taps=True)
ows = [oracle_world(scene, wi, solver=1)[0] for wi in range(16)]
for s in range(30):
    p, R, v, w = b.bodies()
    for k in range(16):
        ows[k].set_state(p[k], R[k], v[k], w[k])
    b.step(scene["dt"])
    con = b.contacts(); st = b.status()
    for k, ow in enumerate(ows):
        ow.step(scene["dt"])
        lam, rhs, rs = ow.solution()
        if len(lam) == 0:
            continue
        d = np.abs(con["lam"][k, :len(lam)] - lam)
        i = int(np.argmax(d))
        if d[i] > 1e-10:
            os_ = ow.stats()
            print(f"step {s} world {k}: max |dlam| {d[i]:.3e} at row {i} (lam {lam[i]:.6e}, max|lam| {np.abs(lam).max():.3e}), rows {len(lam)}, sweeps {os_['sweeps']}, residual {os_['residual']:.3e}, "
                  f"dev residual {st['residual'][k]:.3e}, contacts {os_['n_contacts']}")
b.close()
