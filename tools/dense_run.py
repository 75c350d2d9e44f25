"""Dense (Schur + Murty) path timing: `python tools/dense_run.py [worlds] [steps]` on C4 = chain32."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eggshell_b200 as E
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scene = E.scenes.chain32(W)
b = E.scenes.make_batch(scene, solver=E.SOLVER_DENSE_MURTY)
b.set_profiling(True)
for s in range(steps):
    t0 = time.time()
    b.step(scene["dt"])
    ms = b.kernel_ms()
    st = b.status()
    print(f"step {s}: kernel ms narrow/assemble/solve = {ms[0]:.3f} {ms[1]:.3f} {ms[2]:.3f}; contacts {st['n_contacts'].mean():.1f} "
          f"pivots {st['sweeps'].mean():.1f} status_or {int(np.bitwise_or.reduce(st['status']))} wall {time.time()-t0:.2f}s", flush=True)
print("device MB", b.device_bytes / 1e6)
b.close()
