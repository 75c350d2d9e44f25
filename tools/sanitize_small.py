"""Small batches of every kernel family for compute-sanitizer (racecheck / synccheck / memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eggshell_b200 as E

def run(name, scene, steps, **kw):
    b = E.scenes.make_batch(scene, **kw)
    for _ in range(steps):
        b.step(scene["dt"])
    st = b.status()
    print(name, "status_or", int(np.bitwise_or.reduce(st["status"])), "contacts", float(st["n_contacts"].mean()), flush=True)
    b.close()

run("pgs pile64", E.scenes.pile64(12), 2, solver=E.SOLVER_PGS, k_max=6)
run("pgs stack10", E.scenes.stack10(70), 2, solver=E.SOLVER_PGS, k_max=8)
run("pgs legged20", E.scenes.legged20(19), 2, solver=E.SOLVER_PGS, k_max=8)
run("dense chain32", E.scenes.chain32(6), 3, solver=E.SOLVER_DENSE_MURTY)
run("dense cairn", E.scenes.cairn(8, rocks=4, zb=(0.2, 0.5), seed=21), 6, solver=E.SOLVER_DENSE_MURTY)
run("sor cairn", E.scenes.cairn(8, rocks=4, zb=(0.2, 0.5), seed=31), 4, solver=E.SOLVER_SOR, k_max=10, cfm=0.1)
sc = E.scenes.chain(6, links=5, anchor=(0.0, 0.0, 3.0))
sc["p"] += np.random.default_rng(5).uniform(-0.02, 0.02, size=sc["p"].shape)
b = E.Batch(sc["W"], sc["n"], sc["nj"], solver=E.SOLVER_PGS)
b.set_bodies(sc["p"], sc["R"], sc["v"], sc["w"], sc["m"], sc["I"]); b.set_joints(sc["i0"], sc["i1"], sc["c0"], sc["c1"]); b.init()
print("init_stabilize", b.init_stabilize(max_steps=3)[0].tolist(), "post_stabilize", b.post_stabilize(max_steps=3)[0].tolist())
b.close()
