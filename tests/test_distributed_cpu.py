"""world_size-2 gloo test (CPU) of the multi-GPU host logic: slice bounds, rank-ordered all-gather of the partial
propagators and their ordered combination (parament_b200/distributed.py).  The per-rank GPU work is replaced by
the oracle evaluated on the rank's slice, so only the plumbing is under test here; the CUDA slices and the
device-side combine are covered by tests/test_parity_gpu.py::test_time_slices_compose."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleContext:
    """Stands in for parament_b200.Parament on a CPU box (same three methods distributed.py uses)."""

    def __init__(self, w):
        self.w = w

    def steps_of(self, pts):
        from oracle.equiprop_oracle import effective_steps
        return effective_steps(pts, self.w.quadrature, self.w.use_magnus)[0]

    def equiprop_slice(self, dt, carr, lo, hi):
        from oracle.equiprop_oracle import _prepare, _slice_product
        from oracle.equiprop_oracle import _QUAD_NAMES
        H0, H1, c = _prepare(self.w.H0, self.w.H1, carr, self.w.precision)
        return _slice_product((H0, H1, c, dt, _QUAD_NAMES[self.w.quadrature], self.w.use_magnus, lo, hi, 4096))

    def combine(self, parts):
        acc = np.eye(parts.shape[-1], dtype=np.complex128)
        for p in parts:            # later slice on the left
            acc = p @ acc
        return acc


def _worker(rank, world, port, name, pts, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from parament_b200.distributed import slice_bounds, time_sliced_equiprop
    from workloads import make_workload
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    w = make_workload(name, pts=pts)
    U = time_sliced_equiprop(OracleContext(w), w.dt, w.carr)
    b = slice_bounds(w.steps, world)
    assert b[0] == 0 and b[-1] == w.steps and all(b[i] <= b[i + 1] for i in range(world))
    if rank == 0:
        np.save(out_path, U)
    else:
        assert U is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,pts,world", [("C2", 2001, 2), ("C1", 1000, 2), ("C2", 603, 3)])
def test_time_sliced_equiprop_gloo(tmp_path, name, pts, world):
    from oracle.equiprop_oracle import equiprop_oracle, rel_frobenius
    from workloads import make_workload
    out = str(tmp_path / "u.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, name, pts, out), nprocs=world, join=True)
    w = make_workload(name, pts=pts)
    whole = equiprop_oracle(w.H0, w.H1, w.carr, w.dt, w.quadrature, w.use_magnus, w.precision)
    assert rel_frobenius(np.load(out), whole) < 1e-12


def test_slice_bounds():
    from parament_b200.distributed import slice_bounds
    assert slice_bounds(10, 3) == [0, 3, 6, 10]
    assert slice_bounds(2, 4) == [0, 0, 1, 1, 2]
    assert slice_bounds(499999, 8)[-1] == 499999
