"""Error-vs-N of the complex64 dim-8 kernels against the float64 oracle (tests/golden/growth_dim8.npz): FP64 arithmetic (DMMA),
FP32 as 3xTF32 with and without the compensated running product.  One subprocess per variant (the switches are read once per
process).  Writes a markdown table to stdout:   python tools/gpu_errgrowth_tf32.py > profiles/error_growth_tf32_r2.md"""
import json, os, subprocess, sys, textwrap
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = textwrap.dedent(f"""
    import sys, json, os, numpy as np, time
    sys.path.insert(0, {ROOT!r})
    import parament_b200 as pb
    from workloads import make_workload
    gold = np.load(os.path.join({ROOT!r}, "tests", "golden", "growth_dim8.npz"))
    out = {{}}
    for key in gold.files:
        pts = int(key[1:])
        w = make_workload("C5", pts=pts, batch=1)
        with pb.Parament("fp32") as ctx:
            ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="none")
            U = ctx.equiprop(w.dt, *w.carr)
            math, M = ctx.stat(15), ctx.stat(2)
        G = gold[key]
        out[pts] = [float(np.linalg.norm(U.astype(np.complex128) - G) / np.linalg.norm(G)), math, M]
    # full C5 ensemble: worst pulse
    full = np.load(os.path.join({ROOT!r}, "tests", "golden", "full_C5.npz"))
    w = make_workload("C5")
    with pb.Parament("fp32") as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="none")
        U = ctx.equiprop_batch(w.dt, w.carr)
        t = time.time(); U = ctx.equiprop_batch(w.dt, w.carr); ms = (time.time() - t) * 1e3
        math = ctx.stat(15)
    G = full["U"]                                  # the first 16 pulses of the ensemble
    errs = [float(np.linalg.norm(U[k].astype(np.complex128) - G[k]) / np.linalg.norm(G[k])) for k in range(G.shape[0])]
    out["ensemble"] = [max(errs), math, ms]
    print(json.dumps(out))
""")
variants = [("FP64 (DMMA)", {"PARAMENT_C64_MATH": "f64"}), ("3xTF32, compensated product", {"PARAMENT_C64_MATH": "tf32", "PARAMENT_TF32_COMP": "1"}),
            ("3xTF32, plain fp32 product", {"PARAMENT_C64_MATH": "tf32"}), ("automatic", {})]
res = {}
for name, env in variants:
    r = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, **env), capture_output=True, text=True)
    if r.returncode != 0:
        print(name, "FAILED", r.stderr[-1500:], file=sys.stderr)
        continue
    res[name] = json.loads(r.stdout.strip().splitlines()[-1])
names = list(res)
print("Relative Frobenius error against the float64 oracle; one dim-8 complex64 pulse of N steps (C5 Hamiltonians, x = 0.2), tolerance 1e-5.")
print()
print("| N | " + " | ".join(names) + " |")
print("|---|" + "---|" * len(names))
keys = sorted(int(k) for k in res[names[0]] if k != "ensemble")
for k in keys:
    print(f"| {k} | " + " | ".join(f"{res[n][str(k)][0]:.2e}" + (" (tf32)" if res[n][str(k)][1] == 1 else "") for n in names) + " |")
print("| C5 ensemble, worst of the golden pulses | " + " | ".join(f"{res[n]['ensemble'][0]:.2e}" + (" (tf32)" if res[n]['ensemble'][1] == 1 else "") for n in names) + " |")
print("| C5 ensemble, host call ms | " + " | ".join(f"{res[n]['ensemble'][2]:.2f}" for n in names) + " |")
