"""Quick timing of the dense path: python tools/bench_dense_quick.py [worlds] [settle] [steps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eggshell_b200 as E
W = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 4
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
scene = E.scenes.chain32(W, seed=4000)
b = E.scenes.make_batch(scene, solver=E.SOLVER_DENSE_MURTY)
b.step(scene["dt"], n_steps=settle)
b.snapshot()
b.set_profiling(True)
for s in range(steps):
    b.restore(); b.step(scene["dt"])
    ms = b.kernel_ms(); st = b.status()
    print(f"step {s}: narrow/assemble/solve ms = {ms[0]:.2f} {ms[1]:.2f} {ms[2]:.2f}  -> {W / (sum(ms[:3]) * 1e-3):.0f} world-steps/s; "
          f"pivots {st['pivots'].mean():.1f} rows {st['n_rows'].mean():.1f} status_or {int(np.bitwise_or.reduce(st['status']))} "
          f"flops/world {b.dense_work().mean():.3e}", flush=True)
    c = b.debug_counters()
    if c[31]:
        names = ["rows+A", "cfm", "schur", "check", "index", "order", "gather", "factor", "solve", "w", "best", "x_e", "out", "f:A", "f:B", "f:C1", "f:C2", "s:fwd", "s:bwd"]
        tot = float(c[:19].sum())
        print("   phase cycles per world (timing build): " + ", ".join(f"{n} {c[i] / c[31]:.0f} ({100 * c[i] / tot:.1f}%)" for i, n in enumerate(names)), flush=True)
b.close()
