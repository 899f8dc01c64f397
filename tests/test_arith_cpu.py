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
