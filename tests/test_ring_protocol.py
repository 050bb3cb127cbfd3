"""The mbarrier hand-off protocol of conv_row.cu / conv_tc.cu under random schedules (tests/ring_model.py): the ring depths and role
counts the kernels ship with are free of parity-wait hazards, and the two configurations that were not (an odd
raw ring under two transform groups; two MMA-issuing warps on alternate tiles) are caught by the same model."""
import os
import re

import pytest

from ring_model import explore, first_acc_violation, first_violation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "speech-denoising-diffusion-model-2_b200", "csrc")


def _src(name):
    with open(os.path.join(CSRC, name)) as f:
        return f.read()


def test_sources_still_say_what_the_model_assumes():
    row, tc = _src("conv_row.cu"), _src("conv_tc.cu")
    assert re.search(r"constexpr int kNR = 6, kNA = 6, kNS = 5;", row)
    assert "static constexpr int NR = PAIR ? (NMAIN >= 4 ? 4 : 6) : kNR;" in row
    assert "static constexpr int NA = PAIR ? (NMAIN >= 4 ? 3 : 4) : kNA;" in row
    assert 'static_assert(NR % 2 == 0' in row
    assert "a.NR = 2; a.NA = 2;" in tc and "a.NR += 2" in tc and "a.NA += 2" in tc     # conv_tc ring depths: 2, 4, 6
    assert "nmma = 1;" in tc                                                          # one MMA-issuing warp
    assert "odd ring depth" in tc


@pytest.mark.parametrize("nr,na,per_tile", [(6, 6, 1), (6, 6, 3),            # conv_row_kernel, 128-wide level (1 - 3 slabs per row)
                                            (4, 3, 4), (6, 4, 5), (6, 4, 2),  # pair tile: 128 -> 32, 32 -> 32 + res 128, 64 -> 32
                                            (2, 2, 5), (4, 2, 3), (4, 4, 3), (6, 4, 4), (6, 6, 10)])   # conv3x3_tc_kernel plans
def test_shipped_rings_are_hazard_free(nr, na, per_tile):
    assert first_violation(nr, na, slabs=12 * per_tile, slabs_per_tile=per_tile, issuers=1, trials=150) is None


@pytest.mark.parametrize("nr,na,per_tile,issuers,what", [(5, 4, 5, 1, "the first pair-tile build"), (3, 3, 2, 1, "a conv_tc plan of round 1"),
                                                         (4, 4, 3, 2, "two issuing warps, ups.14.block2 in the bf16 mode")])
def test_the_two_races_of_round_2_are_caught(nr, na, per_tile, issuers, what):
    v = first_violation(nr, na, slabs=12 * per_tile, slabs_per_tile=per_tile, issuers=issuers, trials=300)
    assert v is not None, what
    assert "wrong slab" in v or "over-arrival" in v or "deadlock" in v


@pytest.mark.parametrize("rows", [3, 4, 5, 17, 18, 113])
def test_accumulator_slots_of_the_row_kernel(rows):
    """Five TMEM slots, one per input row; even / odd output rows on two epilogue groups; three arrivals hand a slot back, the first /
    last row of a run arrives for the rows that never come (conv_row.cu: kNS = 5, acc_empty counts 12 = 3 rows x 4 warps)."""
    assert first_acc_violation(5, rows=rows, trials=120) is None


def test_accumulator_model_has_teeth():
    assert first_acc_violation(2, rows=20, trials=50) is not None     # an output row needs three slots at once


@pytest.mark.parametrize("nr,na,per_tile", [(4, 4, 3), (6, 6, 5), (2, 2, 3), (6, 4, 10)])
def test_two_issuing_warps_are_safe_with_one_full_barrier_set_each(nr, na, per_tile):
    """The design DESIGN.md section 4.2a names for bringing the second MMA-issuing warp back (not in the kernels today): the transform
    arrives on the full_a set of the warp that owns the slab's tile and every warp counts its own uses of a stage."""
    assert first_violation(nr, na, slabs=16 * per_tile, slabs_per_tile=per_tile, issuers=2, trials=150, per_warp_full=True) is None
    assert first_violation(nr, na, slabs=16 * per_tile, slabs_per_tile=per_tile, issuers=2, trials=150) is not None     # shared set: the race


@pytest.mark.parametrize("nr,na,slabs,per_tile", [(6, 6, 36, 1), (6, 6, 39, 3), (4, 3, 32, 4), (6, 4, 40, 5), (6, 4, 30, 2),
                                                  (2, 2, 30, 5), (4, 2, 30, 3), (4, 4, 30, 3), (6, 4, 32, 4), (6, 6, 40, 10)])
def test_shipped_rings_exhaustively(nr, na, slabs, per_tile):
    """EVERY interleaving of loader / out-of-order load completion / two transform groups / one issuing warp over 5 - 15 ring wraps:
    no wrong slab, no over-arrival, no deadlock (10^3 - 10^5 reachable states each)."""
    violation, states = explore(nr, na, slabs, per_tile, issuers=1)
    assert violation is None, (violation, states)


@pytest.mark.parametrize("nr,na,slabs,per_tile,issuers", [(5, 4, 15, 5, 1), (3, 3, 8, 1, 1), (5, 5, 12, 1, 1), (4, 4, 12, 3, 2), (6, 6, 20, 5, 2)])
def test_known_bad_shapes_exhaustively(nr, na, slabs, per_tile, issuers):
    violation, _ = explore(nr, na, slabs, per_tile, issuers=issuers)
    assert violation is not None
