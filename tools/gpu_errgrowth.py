"""Error-vs-N curves against the CPU oracle (tests/golden/growth.npz, full_*.npz) for this library and, where it fits,
for the reference's own CUDA build (oracle/_ref/libparament.so).  Writes a markdown table to stdout."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import parament_b200 as pb
from workloads import make_workload

QUAD = {"none": 0, "midpoint": 0x01000000, "simpson": 0x02000000}
growth = np.load(os.path.join(ROOT, "tests", "golden", "growth.npz"))


def relf(a, b):
    a = np.asarray(a, dtype=np.complex128); b = np.asarray(b, dtype=np.complex128)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def ours(w, carr):
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        U = ctx.equiprop(w.dt, *carr)
        return U, ctx.stats()


def reference(w, carr):
    path = os.path.join(ROOT, "oracle", "_ref", "libparament.so")
    if not os.path.exists(path):
        return None
    n, pts = w.dim, carr.shape[1]
    fp64 = w.precision == "fp64"
    if n * n * pts >= 2 ** 31 or 3 * n * n * pts * (16 if fp64 else 8) > 120e9:
        return None            # does not fit / 32-bit index overflow
    lib = ctypes.cdll.LoadLibrary(path)
    sfx = "_fp64" if fp64 else ""
    h = ctypes.c_void_p()
    if getattr(lib, "Parament_create" + sfx)(ctypes.byref(h)) != 0:
        return None
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    H0 = np.ascontiguousarray(w.H0.ravel()); H1 = np.ascontiguousarray(w.H1.ravel()); c = np.ascontiguousarray(carr.ravel())
    out = np.zeros(n * n, dtype=w.ctype)
    getattr(lib, "Parament_setHamiltonian" + sfx)(h, vp(H0), vp(H1), ctypes.c_uint(n), ctypes.c_uint(w.amps), ctypes.c_bool(w.use_magnus), ctypes.c_int(QUAD[w.quadrature]))
    ec = getattr(lib, "Parament_equiprop" + sfx)(h, vp(c), ctypes.c_double(w.dt), ctypes.c_uint(pts), ctypes.c_uint(w.amps), vp(out))
    getattr(lib, "Parament_destroy" + sfx)(h)
    return out.reshape(n, n) if ec == 0 else None


print("| config | points | effective steps | this library: rel. Frobenius error vs oracle | degree used / table | reference CUDA build: error vs oracle |")
print("|---|---|---|---|---|---|")
for name in ("C2", "C3", "C4"):
    full = make_workload(name)
    sizes = sorted(int(k.split("_")[1]) for k in growth.files if k.startswith(name + "_"))
    golds = [(p, growth[f"{name}_{p}"]) for p in sizes]
    fg = os.path.join(ROOT, "tests", "golden", f"full_{name}.npz")
    if os.path.exists(fg):
        golds.append((full.pts, np.load(fg)["U"]))
    for pts, G in golds:
        carr = np.ascontiguousarray(full.carr[:, :pts])
        U, st = ours(full, carr)
        R = reference(full, carr)
        r = f"{relf(R, G):.2e}" if R is not None else "does not fit (SURVEY 6)"
        print(f"| {name} | {pts} | {int(st['steps'])} | {relf(U, G):.2e} | {int(st['degree_used'])} / {int(st['degree_reference'])} | {r} |", flush=True)
