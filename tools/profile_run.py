"""Small fixed workload for ncu: `python tools/profile_run.py [workload] [worlds] [k_max] [steps]`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eggshell_b200 as E

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1776
k_max = int(sys.argv[3]) if len(sys.argv) > 3 else 20
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
fn = {"c2": E.scenes.stack10, "c3": E.scenes.pile64, "c4": E.scenes.chain32, "c5": E.scenes.legged20}[wl]
scene = fn(W)
b = E.scenes.make_batch(scene, solver=E.SOLVER_DENSE_MURTY if wl == "c4" else E.SOLVER_PGS, k_max=k_max, max_contacts=1024 if wl == "c3" else 0)
b.set_profiling(True)
for s in range(steps):
    b.step(scene["dt"])
    ms = b.kernel_ms()
    st = b.status()
    print(f"step {s}: kernel ms narrow/assemble/solve = {ms[0]:.3f} {ms[1]:.3f} {ms[2]:.3f}; contacts {st['n_contacts'].mean():.1f} "
          f"sweeps {st['sweeps'].mean():.1f} pivots {st['pivots'].mean():.1f} rows {st['n_rows'].mean():.1f} status_or {int(np.bitwise_or.reduce(st['status']))}", flush=True)
bad = np.nonzero(b.status()['status'])[0]
if len(bad):
    print('worlds with status != 0:', len(bad), bad[:32].tolist(), 'status values', np.unique(b.status()['status'][bad]).tolist())
    p, R, v, w = b.bodies()
    con = b.contacts()
    for wb in bad[:4]:
        nf = lambda a: np.nonzero(~np.isfinite(a.reshape(a.shape[0], -1)).all(axis=1))[0].tolist()[:12]
        print(' world', wb, 'nonfinite p', nf(p[wb]), 'R', nf(R[wb]), 'v', nf(v[wb]), 'w', nf(w[wb]), 'max|v|', float(np.nanmax(np.abs(v[wb]))), 'max|w|', float(np.nanmax(np.abs(w[wb]))),
              'lam nonfinite', int((~np.isfinite(con['lam'][wb])).sum()), 'max|lam|', float(np.nanmax(np.abs(con['lam'][wb]))), 'count', int(con['count'][wb]))
    good = np.setdiff1d(np.arange(W), bad)[:1]
    print(' a good world', good, 'max|v|', float(np.abs(v[good]).max()), 'max|lam|', float(np.abs(con['lam'][good]).max()))
b.close()
