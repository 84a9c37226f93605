"""Development: the quadcopter solve of two library builds compared (outputs, iteration counts).  usage: gpu_quad_cmp.py soA soB [B]"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import obca_b200
    from obca_b200 import quadcopter, scenarios
    B = int(sys.argv[3])
    sc = scenarios.quadcopter_batch(B, 100, seed=2)
    o = obca_b200.default_opts(); o.max_iter = 3000
    out = {}
    for sd in (1, 0):
        r = quadcopter.quadcopter_solve_batch(sc["x0"], sc["xF"], sc["N"], sc["Ts"], sc["R"], sc["obs"], sc["xWS"], 1.0, sd, o)
        for k in ("xp", "up", "ts", "lp", "iters", "exitflag"): out[f"{k}{sd}"] = r[k]
    np.savez(sys.argv[2], **out)
    sys.exit(0)
soA, soB = sys.argv[1], sys.argv[2]
B = sys.argv[3] if len(sys.argv) > 3 else "64"
res = []
for i, so in enumerate((soA, soB)):
    f = f"/tmp/qcmp{i}.npz"
    subprocess.check_call([sys.executable, __file__, "--child", f, B], env=dict(os.environ, OBCA_SO=os.path.abspath(so)))
    res.append(np.load(f))
a, b = res
for sd in (1, 0):
    print(f"sd={sd}: exitflag equal {np.array_equal(a[f'exitflag{sd}'], b[f'exitflag{sd}'])} ok {int((a[f'exitflag{sd}']>=1).sum())}/{int((b[f'exitflag{sd}']>=1).sum())}; "
          f"iters equal {int((a[f'iters{sd}'] == b[f'iters{sd}']).sum())}/{len(a[f'iters{sd}'])} (mean {a[f'iters{sd}'].mean():.1f} / {b[f'iters{sd}'].mean():.1f}); "
          f"max|dx| {np.abs(a[f'xp{sd}'] - b[f'xp{sd}']).max():.2e} max|du| {np.abs(a[f'up{sd}'] - b[f'up{sd}']).max():.2e} "
          f"max|dlam| {np.abs(a[f'lp{sd}'] - b[f'lp{sd}']).max():.2e}")
    dx = np.abs(a[f'xp{sd}'] - b[f'xp{sd}']).reshape(len(a[f'iters{sd}']), -1).max(1)
    same = a[f'iters{sd}'] == b[f'iters{sd}']
    print("   per-problem max|dx| quantiles (50/90/99/100 %):", np.quantile(dx, [0.5, 0.9, 0.99, 1.0]), "; same-iteration problems: max", dx[same].max() if same.any() else None,
          "; > 1e-3:", int((dx > 1e-3).sum()), "of which same-iteration", int(((dx > 1e-3) & same).sum()))
