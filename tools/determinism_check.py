"""Run the same full-size step several times from the same state and compare every output bit for bit:
`python tools/determinism_check.py [workload] [worlds] [k_max] [repeats]`.  The kernels are
deterministic, so any difference is a data race / corruption."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eggshell_b200 as E

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
k_max = int(sys.argv[3]) if len(sys.argv) > 3 else 5
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
fn = {"c2": E.scenes.stack10, "c3": E.scenes.pile64, "c5": E.scenes.legged20}[wl]
scene = fn(W)
b = E.scenes.make_batch(scene, solver=E.SOLVER_PGS, k_max=k_max, taps=False, max_contacts=1024 if wl == "c3" else 0)
b.snapshot()
ref = None
ndiff_total = 0
for r in range(reps):
    b.restore()
    b.step(scene["dt"])
    out = [x.copy() for x in b.bodies()] + [b.contacts()["lam"].copy(), b.status()["status"].copy()]
    if ref is None:
        ref = out
        print(f"run 0: status_or {int(np.bitwise_or.reduce(out[-1]))}")
        continue
    diff_worlds = np.zeros(W, dtype=bool)
    for a, c in zip(out, ref):
        a2, c2 = a.reshape(W, -1), c.reshape(W, -1)
        diff_worlds |= (a2.view(np.uint8).reshape(W, -1) != c2.view(np.uint8).reshape(W, -1)).any(axis=1)
    nd = int(diff_worlds.sum())
    ndiff_total += nd
    worst = max(float(np.nanmax(np.abs(a - c))) for a, c in zip(out[:5], ref[:5]))
    print(f"run {r}: worlds differing from run 0: {nd} {np.nonzero(diff_worlds)[0][:12].tolist()} worst abs diff {worst:.3e} status_or {int(np.bitwise_or.reduce(out[-1]))}")
    if nd:
        dw = np.nonzero(diff_worlds)[0]
        print("   lane (world % 32) histogram of differing worlds:", np.bincount(dw % 32, minlength=32).tolist())
        print("   groups (world // 32) fully differing:", int((np.bincount(dw // 32, minlength=(W + 31) // 32) == 32).sum()), "partially:", int(((np.bincount(dw // 32, minlength=(W + 31) // 32) > 0) & (np.bincount(dw // 32, minlength=(W + 31) // 32) < 32)).sum()))
        w0 = dw[0]
        lam_a, lam_c = out[4][w0].ravel(), ref[4][w0].ravel()
        idx = np.nonzero(lam_a != lam_c)[0]
        print(f"   world {w0}: lam entries differing {len(idx)} of {np.count_nonzero(lam_c)} nonzero; first {idx[:8].tolist()}; a {lam_a[idx[:4]].tolist()} vs c {lam_c[idx[:4]].tolist()}")
b.close()
print("DETERMINISTIC" if ndiff_total == 0 else "NONDETERMINISTIC")
sys.exit(0 if ndiff_total == 0 else 1)
