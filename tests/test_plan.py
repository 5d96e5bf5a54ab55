"""Host-side work partitioning (parament_b200/csrc/plan.hpp) on the CPU: how pulses and time steps are dealt to warps,
CTAs, copy groups and devices.  The header is pure integer logic, compiled here with g++ (tests/cpp/plan_check.cpp)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMS = 148          # B200


@pytest.fixture(scope="module")
def plan(tmp_path_factory):
    exe = tmp_path_factory.mktemp("plan") / "plan_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "parament_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "plan_check.cpp"), "-o", str(exe)], check=True)
    env = {k: v for k, v in os.environ.items() if not k.startswith("PARAMENT_")}

    def run(*args):
        out = subprocess.run([str(exe)] + [str(a) for a in args], capture_output=True, text=True, check=True, env=env).stdout
        return json.loads(out)
    return run


def test_single_pulse_fills_exactly_one_wave(plan):
    """C2: one pulse of 499 999 steps, dim 16, product-saving series -> 2 CTAs of 4 warps per SM, one partial per CTA."""
    p = plan("k1", 16, 1, 499999, SMS, 1)
    assert p["ctas_per_sm"] == 2 and p["warp_slots"] == SMS * 2 * 4
    assert p["grid"] == SMS * 2 and p["chunks_per_pulse"] == SMS * 2 * 4
    assert p["reduce_in_cta"] == 1 and p["partials_per_pulse"] == SMS * 2
    assert p["partial_elems"] == SMS * 2 * 16 * 16
    assert p["k3_launches"] == 2 and p["mid_elems"] == ((SMS * 2 + 15) // 16) * 256      # two-level ordered reduction
    # Clenshaw variant runs 3 CTAs per SM
    assert plan("k1", 16, 1, 499999, SMS, 0)["grid"] == SMS * 3


def test_fused_final_stage_plans(plan):
    """Plans of the single-launch path (k1_common.cuh): one long pulse -> groups of 16 CTA partials for the last-arriver
    reduction; ensembles -> a pulse is cut into 1, 2 or 4 chunks so that its warps share a CTA; no groups."""
    p = plan("k1", 16, 1, 499999, SMS, 1, 1)
    assert p["reduce_in_cta"] == 1 and p["partials_per_pulse"] == SMS * 2 and p["groups_per_pulse"] == (SMS * 2 + 15) // 16
    assert plan("k1", 16, 1, 499999, SMS, 1)["groups_per_pulse"] == 0          # legacy plan: a k3_reduce launch follows
    for batch, nsteps in ((10000, 1000), (1250, 1000), (5000, 64), (3552, 100000)):
        e = plan("k1", 8, batch, nsteps, SMS, 1, 1)
        if e["reduce_in_cta"] == 0:
            assert e["chunks_per_pulse"] in (1, 2, 4) and e["groups_per_pulse"] == 0
            assert e["grid"] * 4 >= batch * e["chunks_per_pulse"]
        else:
            assert e["groups_per_pulse"] >= 1 and e["chunks_per_pulse"] % 4 == 0
    small = plan("k1", 8, 1, 9999, SMS, 1, 1)                                    # C1: one pulse, 313 CTAs -> 20 groups
    assert small["groups_per_pulse"] == (small["partials_per_pulse"] + 15) // 16 >= 2


def test_short_pulses_keep_eight_steps_per_warp(plan):
    p = plan("k1", 16, 1, 100, SMS, 1)
    assert p["chunks_per_pulse"] * 8 <= 100 + 8 * 4 and p["grid"] == p["partials_per_pulse"] >= 1
    p = plan("k1", 8, 1, 3, SMS, 1)              # fewer steps than warps in a CTA: one warp owns the pulse
    assert p["chunks_per_pulse"] == 1 and p["grid"] == 1 and p["reduce_in_cta"] == 0
    assert plan("k1", 8, 1, 0, SMS, 1)["grid"] == 1


@pytest.mark.parametrize("batch", [1776, 3552, 10000, 3340, 123457])
def test_ensembles_prefer_short_ctas_and_balanced_sms(plan, batch):
    """C5-like: dim 8, 1000 steps per pulse.  Every warp owns 1/k of a pulse; the measured model prefers large k."""
    p = plan("k1", 8, batch, 1000, SMS, 1)
    k = p["chunks_per_pulse"]
    assert p["reduce_in_cta"] == 0 and p["partials_per_pulse"] == k and 1 <= k <= 8
    assert 1000 // k >= 64
    assert p["grid"] == -(-batch * k // 4)
    assert p["partial_elems"] == batch * k * 64
    ctas = -(-batch * k // 4)
    assert -(-ctas // SMS) / (ctas / SMS) < 1.02          # no SM carries more than 2 % above the mean
    assert k >= 6


def test_short_ensemble_pulses_are_not_split_below_64_steps(plan):
    assert plan("k1", 8, 5000, 100, SMS, 1)["chunks_per_pulse"] == 1
    assert plan("k1", 8, 5000, 200, SMS, 1)["chunks_per_pulse"] <= 3


def test_ensemble_copy_groups(plan):
    unit = SMS * 6 * 4 // 8                                   # 444 pulses = 1/8 wave at dim 8
    gb = plan("egroups", 10000, unit, 8)
    assert gb == [0, 444, 1332, 3108, 6660, 10000]            # 1, 2, 4, 8 units, then the rest (less than 16 + 1 units)
    gb = plan("egroups", 10000, unit, 3)                      # at most three groups: the third takes what is left
    assert gb == [0, 444, 1332, 10000]
    gb = plan("egroups", 700, unit, 8)                        # small ensemble: equal shares
    assert gb[0] == 0 and gb[-1] == 700 and len(gb) == 9 and max(b - a for a, b in zip(gb, gb[1:])) == 88
    assert plan("egroups", 5, unit, 8) == [0, 1, 2, 3, 4, 5]
    assert plan("egroups", 10000, unit, 1) == [0, 10000]
    for batch in (1776, 1777, 4 * unit, 4 * unit + 1, 99999):
        gb = plan("egroups", batch, unit, 8)
        assert gb[0] == 0 and gb[-1] == batch and len(gb) <= 9
        assert all(b > a for a, b in zip(gb, gb[1:]))


def test_time_copy_groups(plan):
    b = plan("tgroups", 499999, 6)
    assert b[0] == 0 and b[-1] == 499999 and len(b) == 7
    sizes = [y - x for x, y in zip(b, b[1:])]
    assert sizes[0] == 499999 // 63                           # the first copy is the only exposed one: 1 / (2^G - 1) of the pulse
    assert all(abs(y - 2 * x) <= 2 for x, y in zip(sizes, sizes[1:]))      # every group twice the one before
    assert plan("tgroups", 1000, 1) == [0, 1000]
    assert plan("tgroups", 40000, 2) == [0, 13333, 40000]


@pytest.mark.parametrize("configured,batch,nsteps,npad,expect", [
    (8, 1, 499999, 16, 8),        # C2 on 8 GPUs: 62 500 steps per device >= 16 384
    (8, 1, 70000, 16, 4),
    (8, 1, 2000, 16, 1),          # too short to share
    (8, 1, 1000000, 64, 8),       # C3
    (8, 1, 100000, 256, 8),       # C4
    (8, 1, 300, 256, 4),          # at least 64 steps per device
    (8, 10000, 1000, 8, 8),       # C5: pulses shared
    (8, 3, 1000000, 8, 3),        # never more devices than pulses
    (2, 700, 1000, 8, 2),
    (1, 1, 10 ** 9, 16, 1),
])
def test_devices_taking_part_in_a_call(plan, configured, batch, nsteps, npad, expect):
    d = plan("devices", configured, batch, nsteps, npad)
    assert d["devices"] == expect
    assert d["min_steps"] == max(64, 2 ** 26 // npad ** 3)


@pytest.mark.parametrize("BM,BN,n", [(64, 32, 256), (64, 32, 128), (64, 64, 256), (32, 32, 96), (64, 32, 64), (64, 32, 512)])
def test_hermitian_gemm_tiles(plan, BM, BN, n):
    """Hermitian-output GEMM of dim > 64 (k4_gemm.cu, GemmArgs::herm): the launch enumerates exactly the tiles that touch the upper
    triangle; every element of a skipped tile is the mirror image of an element of an enumerated tile, and no element of an
    enumerated tile is written twice (a mirror write only ever lands in a skipped tile)."""
    p = plan("herm", BM, BN, n)
    tiles = [tuple(t) for t in p["tiles"]]
    skipped = {tuple(t) for t in p["skipped"]}
    assert p["count"] == len(tiles) == len(set(tiles))
    assert set(tiles) | skipped == {(i, j) for i in range(n // BM) for j in range(n // BN)} and not (set(tiles) & skipped)
    if (BM, BN, n) == (64, 32, 256):
        assert p["count"] == 20                                 # C4: 12 of the 32 tiles are mirrored
    for (i, j) in skipped:                                      # strictly below the diagonal, mirror image entirely in computed tiles
        assert BM * i > BN * j + BN - 1
        for r in (BM * i, BM * i + BM - 1):
            for c in (BN * j, BN * j + BN - 1):
                assert (c // BM, r // BN) in set(tiles)
    for (i, j) in tiles:                                        # a computed tile that holds a diagonal or upper element is never a mirror target
        assert BM * i <= BN * j + BN - 1
