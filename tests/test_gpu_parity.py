"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle, one step at a
time from shared state.  Discrete outputs bit-exact; p, R, v, w, lambda within 1e-9 relative."""
import numpy as np
import pytest

from tests.helpers import oracle_world, compare_step, rel_err

pytestmark = pytest.mark.gpu


def _stepwise(scene, nsteps, worlds_idx, batch_kw, oracle_kw, tol=1e-9, lam_tol=1e-7):
    import eggshell_b200 as E
    b = E.scenes.make_batch(scene, taps=True, **batch_kw)
    ows = [oracle_world(scene, wi, **oracle_kw)[0] for wi in worlds_idx]
    dt = scene["dt"]
    worst_all = {}
    for s in range(nsteps):
        # shared state: the oracle adopts the GPU state before every step
        p, R, v, w = b.bodies()
        for k, wi in enumerate(worlds_idx):
            ows[k].set_state(p[wi], R[wi], v[wi], w[wi])
        b.step(dt)
        for ow in ows:
            ow.step(dt)
        worst = compare_step(b, ows, worlds_idx, tol=tol)
        st = b.status()
        for k, wi in enumerate(worlds_idx):
            os_ = ows[k].stats()
            assert st["sweeps"][wi] == os_["sweeps"], f"step {s} world {wi}: sweeps {st['sweeps'][wi]} != {os_['sweeps']}"
            lam, rhs, rs = ows[k].solution()
            con_rs = b.contacts()["row_state"][wi, :len(rs)]
            assert np.array_equal(con_rs, rs), f"step {s} world {wi}: row clamp states differ"
        assert worst["lam"] <= lam_tol, f"lambda mismatch {worst['lam']:.3e}"
        for k_, v_ in worst.items():
            worst_all[k_] = max(worst_all.get(k_, 0.0), v_)
    b.close()
    return worst_all


def test_stack10_pgs_stepwise():
    import eggshell_b200 as E
    scene = E.scenes.stack10(8, seed=1000)
    worst = _stepwise(scene, 5, list(range(8)), dict(solver=E.SOLVER_PGS), dict(solver=1))
    print("stack10 worst", worst)


def test_pile64_pgs_stepwise():
    import eggshell_b200 as E
    scene = E.scenes.pile64(4, seed=3000)
    worst = _stepwise(scene, 2, list(range(4)), dict(solver=E.SOLVER_PGS, k_max=50), dict(solver=1, k_max=50))
    print("pile64 worst", worst)


def test_cairn_pgs_falling():
    import eggshell_b200 as E
    scene = E.scenes.cairn(16, rocks=4, zb=(0.2, 1.0), seed=11)
    worst = _stepwise(scene, 30, list(range(16)), dict(solver=E.SOLVER_PGS), dict(solver=1))
    print("cairn worst", worst)


def test_legged_pgs_joints_and_contacts():
    import eggshell_b200 as E
    scene = E.scenes.legged20(4, seed=5000)
    worst = _stepwise(scene, 3, list(range(4)), dict(solver=E.SOLVER_PGS, k_max=100), dict(solver=1, k_max=100))
    print("legged worst", worst)
