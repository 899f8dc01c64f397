"""Host-side logic of the path on CPU: WBRO / CSV ingest, frame sharding, and a
world_size-2 gloo run of the multi-rank plumbing bench.py uses (no GPU needed)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import atsc_b200
import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


def test_wbro_golden_bytes():
    # wavbrro/src/wavbrro.rs:223-233 (archive of one sample) behind the 12-byte header (write.rs:22)
    want = b"WBRO0000WBRO" + bytes([0, 0, 0, 0, 0, 0, 240, 63, 248, 255, 255, 255, 1, 0, 0, 0, 248, 255, 255, 255, 1,
                                    0, 0, 0, 1, 0, 0, 0, 5, 0, 0, 0])
    assert atsc_b200.wbro_encode([1.0]) == want
    assert list(atsc_b200.wbro_decode(want)) == [1.0]


def test_wbro_fixture_decodes_like_the_reference_layout():
    fx = np.load(os.path.join(G, "fixtures.npz"))
    blob = open(os.path.join(G, "go_gc_heap_goal_bytes.wbro"), "rb").read()
    got = atsc_b200.wbro_decode(blob)
    assert np.array_equal(got, fx["wbro_go_gc_heap_goal_bytes"])
    # re-encoding reproduces the reference's file byte for byte
    assert atsc_b200.wbro_encode(got) == blob


@pytest.mark.parametrize("n", [0, 1, 3, 2047, 2048, 2049, 10000])
def test_wbro_roundtrip(n):
    x = np.random.default_rng(n).standard_normal(n)
    assert np.array_equal(atsc_b200.wbro_decode(atsc_b200.wbro_encode(x)), x)


def test_wbro_rejects_bad_header():
    with pytest.raises(atsc_b200.AtscError):
        atsc_b200.wbro_decode(b"RIFF0000WAVE" + b"\0" * 32)


def test_csv_reader():
    fx = np.load(os.path.join(G, "fixtures.npz"))
    vals = fx["csv_iowait"]
    text = "time,value\n" + "".join(f"{1730419200 + 20 * i},{float(v)!r}\n" for i, v in enumerate(vals))
    assert np.array_equal(atsc_b200.csv_read_values(text), vals)
    text2 = "".join(f"{float(v)!r}\n" for v in vals)
    assert np.array_equal(atsc_b200.csv_read_values(text2, has_header=False), vals)
    # csv.rs tests: custom field names, missing fields, unparsable values
    t3 = "ts,val\n1,1.5\n2,2.5\n"
    assert list(atsc_b200.csv_read_values(t3, time_field="ts", value_field="val")) == [1.5, 2.5]
    with pytest.raises(atsc_b200.AtscError):
        atsc_b200.csv_read_values(t3)
    with pytest.raises(atsc_b200.AtscError):
        atsc_b200.csv_read_values("time,value\n1,abc\n")


def test_plan_shards_properties():
    rng = np.random.default_rng(1)
    for n_parts in (1, 2, 3, 4, 8):
        for _ in range(20):
            lens = rng.choice([64, 512, 16384, 65536, 131072], size=int(rng.integers(1, 200))).astype(np.uint32)
            first = atsc_b200.plan_shards(lens, n_parts)
            assert first[0] == 0 and first[-1] == len(lens) and len(first) == n_parts + 1
            assert all(a <= b for a, b in zip(first, first[1:]))
            tot = int(lens.astype(np.int64).sum())
            for p in range(n_parts):
                share = int(lens[first[p]:first[p + 1]].astype(np.int64).sum())
                assert share <= tot / n_parts + 131072 * 2


def test_two_rank_gloo_sharding():
    """world_size = 2 over gloo on CPU: every rank takes its shard of the frame table, the shards
    partition the work exactly, and gathering per-rank results in rank order reproduces the
    single-process order (the only inter-rank traffic the path has)."""
    script = os.path.join(HERE, "gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", script],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_SHARDING_OK" in r.stdout
