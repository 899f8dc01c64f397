"""VSRI timestamp index: the C++ restatement (atsc_b200/csrc/vsri.cpp) against
a line-by-line Python restatement of vsri/src/lib.rs on the reference's documented example and
on random timestamp streams.  Host only: no GPU needed."""
import random

import pytest

import atsc_b200


class RefVsri:
    """vsri/src/lib.rs restated (i32 arithmetic; Rust's `/` truncates toward zero)."""

    def __init__(self):
        self.min_ts = self.max_ts = 0
        self.seg = []

    @staticmethod
    def div(a, b):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q

    def cur(self):                                   # :299-304
        return self.seg[-1] if self.seg else [0, 0, 0, 0]

    def count(self):                                 # :368-371
        c = self.cur()
        return c[3] + c[1]

    def update(self, y):                             # :249-285
        if y < self.max_ts:
            return False
        self.max_ts = y
        last = list(self.cur())
        if not self.seg:
            self.min_ts = y
            self.seg.append([0, last[1] + last[3], y, 1])
            return True
        if last[0] == 0:
            self.seg[-1] = [y - last[2], last[1], last[2], 2]          # :375-388
            return True
        b = last[2] - last[0] * last[1]
        if self.div(y - b, last[0]) == last[3] + last[1]:              # :410-428
            self.seg[-1][3] += 1
            return True
        self.seg.append([0, last[1] + last[3], y, 1])                  # :393-400
        return True

    def get_sample(self, y):                         # :312-328
        for m, x0, y0, n in self.seg:
            if y0 <= y <= y0 + m * (n - 1):
                if m == 0:
                    return None                      # the reference panics (division by zero)
                return self.div(y - (y0 - m * x0), m)
        return None

    def get_time(self, x):                           # :331-353
        if x == 0:
            return self.min_ts
        if x > self.count():
            return None
        if x == self.count():
            return self.max_ts
        for m, x0, y0, n in self.seg:
            if x0 <= x < x0 + n:
                return y0 + m * x
        return None

    def next(self, y):                               # :156-172
        if y < self.min_ts:
            return 0
        if y >= self.max_ts:
            return None
        for m, x0, y0, n in reversed(self.seg):
            if y <= y0:
                return x0
        return None

    def prev(self, y):                               # :178-197
        if y < self.min_ts:
            return None
        if y >= self.max_ts:
            return self.count()
        for m, x0, y0, n in self.seg:
            if y < y0:
                return x0 - 1
        return None

    def is_empty(self, t0, t1):                      # :202-245
        if len(self.seg) == 1:
            if (self.min_ts <= t0 <= self.max_ts) or (self.min_ts <= t1 <= self.max_ts):
                return False
            if t0 < self.min_ts and t1 > self.max_ts:
                return False
            return True
        prev_end = 0
        for i, (m, x0, y0, n) in enumerate(self.seg):
            end = y0 + m * (n - 1)
            if i >= 1 and t0 > prev_end and t1 < y0:
                return True
            if (y0 <= t0 < end) or (y0 <= t1 < end):
                return False
            if t0 < y0 and t1 > end:
                return False
            prev_end = end
        return True

    def all_ts(self):                                # :356-366
        return [f * m + y0 for m, x0, y0, n in self.seg for f in range(n)]

    def text(self):                                  # :442-462
        return f"{self.min_ts}\n{self.max_ts}\n" + "".join(f"{a},{b},{c},{d}\n" for a, b, c, d in self.seg)


def test_documented_example():
    """The index quoted in the reference's module comment (lib.rs:34-38)."""
    v = atsc_b200.Vsri()
    ts = [55745 + 15 * i for i in range(166)] + [58505 + 15 * i for i in range(63)]
    for y in ts:
        assert v.update_for_point(y)
    assert v.to_text() == "55745\n59435\n15,0,55745,166\n15,166,58505,63\n"
    assert v.all_timestamps() == ts
    assert (v.min, v.max, v.sample_count, v.segment_count) == (55745, 59435, 229, 2)
    assert v.get_sample(55745 + 15 * 7) == 7 and v.get_sample(55746) is not None  # truncating division
    assert v.get_sample(58490) is None and v.is_empty(58250, 58400)
    assert not v.update_for_point(100)              # a point in the past is refused
    w = atsc_b200.Vsri.from_text(v.to_text())
    assert w.to_text() == v.to_text() and w.all_timestamps() == ts


@pytest.mark.parametrize("seed", range(12))
def test_random_streams_match_restated_reference(seed):
    rng = random.Random(seed)
    v, r = atsc_b200.Vsri(), RefVsri()
    y = rng.randrange(0, 2000)
    for _ in range(rng.randrange(1, 400)):
        kind = rng.random()
        if kind < 0.7:
            y += rng.choice([1, 5, 15, 15, 15, 60])
        elif kind < 0.9:
            y += rng.randrange(0, 500)              # gap (or a repeated timestamp)
        else:
            y -= rng.randrange(1, 50)               # a point in the past: refused by both
        assert v.update_for_point(y) == r.update(y)
        y = max(y, r.max_ts)
    assert v.to_text() == r.text()
    assert v.all_timestamps() == r.all_ts()
    assert (v.min, v.max, v.sample_count) == (r.min_ts, r.max_ts, r.count())
    for q in [r.min_ts - 3, r.min_ts, r.max_ts, r.max_ts + 1] + [rng.randrange(r.min_ts - 5, r.max_ts + 5) for _ in range(200)]:
        assert v.get_sample(q) == r.get_sample(q), q
        assert v.get_next_sample(q) == r.next(q), q
        assert v.get_previous_sample(q) == r.prev(q), q
        t1 = q + rng.randrange(0, 40)
        assert v.is_empty(q, t1) == r.is_empty(q, t1), (q, t1)
    for x in range(-1, r.count() + 3):
        if x >= 0:
            assert v.get_time(x) == r.get_time(x), x
    w = atsc_b200.Vsri.from_text(v.to_text())
    assert w.to_text() == r.text()


def test_day_elapsed_seconds_and_malformed_text():
    assert atsc_b200.day_elapsed_seconds(1730419200) == 0          # 2024-11-01T00:00:00Z
    assert atsc_b200.day_elapsed_seconds(1730419200 + 3661) == 3661
    assert atsc_b200.day_elapsed_seconds(86399) == 86399 and atsc_b200.day_elapsed_seconds(-1) == 86399
    for bad in ("x\n1\n", "1\n2\n1,2,3\n", "1\n2\n1,2,3,4,5\n", "1\n2\n1,2,a,4\n"):
        with pytest.raises(ValueError):
            atsc_b200.Vsri.from_text(bad)


def test_hand_derived_vectors():
    """Expected values worked out BY HAND from the reference source, written as literals (no restatement in the
    loop): vsri/src/lib.rs:249-285 (update_for_point), :311-328 (get_sample), :330-350 (get_time), :364-367,
    :442-462 (text form) and csv-compressor/src/metric.rs:55-65 (milliseconds -> second of the day).

    Stream 10, 20, 30, 50, 60, 61:
      10 -> first point: min = 10, segment [0, 0, 10, 1]
      20 -> rate 0 segment becomes [20 - 10, 0, 10, 2]
      30 -> b = 10 - 10*0 = 10; (30 - 10) / 10 = 2 == n + x0 = 2  -> n = 3
      50 -> (50 - 10) / 10 = 4 != 3                               -> new segment [0, 3, 50, 1]
      60 -> rate 0 segment becomes [10, 3, 50, 2]
      61 -> b = 50 - 10*3 = 20; (61 - 20) / 10 = 4 (truncated) != 5 -> new segment [0, 5, 61, 1]"""
    v = atsc_b200.Vsri()
    for y in (10, 20, 30, 50, 60, 61):
        assert v.update_for_point(y)
    assert v.to_text() == "10\n61\n10,0,10,3\n10,3,50,2\n0,5,61,1\n"
    assert (v.min, v.max, v.sample_count, v.segment_count) == (10, 61, 6, 3)
    # get_time: 0 -> min; x == count -> max; beyond -> None; inside a segment y0 + m * x -- the reference adds
    # m * x without subtracting x0 (lib.rs:343), so samples 3 and 4 read 80 and 90 although they were taken at 50, 60
    assert [v.get_time(x) for x in range(8)] == [10, 20, 30, 80, 90, 61, 61, None]
    # get_sample: (y - b) / m, truncated, inside [y0, y0 + m (n - 1)]
    assert v.get_sample(20) == 1 and v.get_sample(25) == 1 and v.get_sample(30) == 2
    assert v.get_sample(50) == 3 and v.get_sample(60) == 4 and v.get_sample(45) is None and v.get_sample(9) is None
    assert v.all_timestamps() == [10, 20, 30, 50, 60, 61]
    # metric.rs:58: day_elapsed_seconds(sample.timestamp / 1000) -- integer division of the millisecond stamp
    assert atsc_b200.day_elapsed_seconds(1730419215999 // 1000) == 15
    assert atsc_b200.day_elapsed_seconds((1730419200000 + 86399 * 1000 + 999) // 1000) == 86399
