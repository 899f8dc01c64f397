"""The two exact-arithmetic shortcuts the error loops rely on (atsc_b200/csrc/common.cuh: div_1e5_int53,
round_half_away), replayed in exact rational arithmetic on the CPU (tools/div_check.py)."""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import div_check


def test_one_step_quotient_is_the_ieee_quotient():
    assert div_check.check_division(150, random.Random(11)) == []


def test_single_add_rounding_is_round_half_away():
    assert div_check.check_rounding(60000, random.Random(12)) == []


def test_frame_uniform_rounding_addend_is_hidden_by_the_clamp():
    """k_poly1s (poly.cuh: p1_trip) rounds a tame frame's spline values with copysign(pred(0.5), vmin); for values on
    the other side of zero the clamp of round_and_limit_f64 (utils/mod.rs:66-74) must give the same result as
    f64::round's own addend (tools/half_check.py, IEEE double arithmetic in numpy)."""
    import half_check
    assert half_check.check(n=300_000, seed=7) == 0
