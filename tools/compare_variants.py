"""Cross-check of two PGS variants on the same scene: `python tools/compare_variants.py [scene] [W] [k_max] [steps] [A] [B]`.
Sweep counts and clamp states must be equal; multipliers and state within 1e-9 relative."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eggshell_b200 as E

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
k_max = int(sys.argv[3]) if len(sys.argv) > 3 else 50
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
va = sys.argv[5] if len(sys.argv) > 5 else "stream"
vb = sys.argv[6] if len(sys.argv) > 6 else "runs"
fn = {"c2": E.scenes.stack10, "c3": E.scenes.pile64, "c5": E.scenes.legged20}[wl]
scene = fn(W)


def run(variant):
    # "stream:1" = variant stream with EGG_PGS_ISO=1 (cap on the isotropic fast-path level)
    name, _, iso = variant.partition(":")
    quirks = E.QUIRKS_REFERENCE | (16 if name == "runs" else 0)   # "runs" = EGG_OPT_PGS_RUNS, "stream" = the default kernel
    os.environ.pop("EGG_PGS_ISO", None)
    if iso:
        os.environ["EGG_PGS_ISO"] = iso
    b = E.scenes.make_batch(scene, solver=E.SOLVER_PGS, k_max=k_max, taps=True, max_contacts=1024 if wl == "c3" else 0, quirks=quirks)
    out = []
    for s in range(steps):
        b.step(scene["dt"])
        st = b.status()
        con = b.contacts()
        out.append(dict(sweeps=st["sweeps"].copy(), status=st["status"].copy(), resid=st["residual"].copy(),
                        rs=con["row_state"].copy(), lam=con["lam"].copy(), count=con["count"].copy(), bodies=[x.copy() for x in b.bodies()]))
    b.close()
    return out


A, B = run(va), run(vb)
ok = True
for s in range(steps):
    a, b = A[s], B[s]
    sw = np.array_equal(a["sweeps"], b["sweeps"])
    rs = np.array_equal(a["rs"], b["rs"])
    lam = float(np.max(np.abs(a["lam"] - b["lam"])) / max(1.0, float(np.max(np.abs(b["lam"])))))
    bod = max(float(np.max(np.abs(x - y)) / max(1.0, float(np.max(np.abs(y))))) for x, y in zip(a["bodies"], b["bodies"]))
    rr = None
    if a["resid"] is not None:
        rr = float(np.max(np.abs(a["resid"] - b["resid"]) / np.maximum(1e-300, np.abs(b["resid"]))))
    print(f"step {s}: sweeps equal {sw} (mean {a['sweeps'].mean():.1f}/{b['sweeps'].mean():.1f}, min {a['sweeps'].min()}) row_state equal {rs} "
          f"lam rel {lam:.2e} bodies rel {bod:.2e} resid rel {rr} status {int(a['status'].max())}/{int(b['status'].max())} contacts {a['count'].mean():.1f}")
    ok = ok and sw and rs and lam < 1e-9 and bod < 1e-9
print("OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
