"""One context sharding a call over several GPUs of one box (north_star: "shard across the GPUs by frame range,
results are gathered to the host"; reference: the sequential stitch of data.rs:73-75).  A context over 2 (and
over all) devices must give the same frame records and byte-identical payload as a single-device context, for
compression and decompression, on a mixed fleet.  Skipped with fewer than 2 devices."""
import numpy as np
import pytest

import gen

pytestmark = pytest.mark.gpu


def fleet():
    kinds = ["periodic", "gauge", "util", "saw", "steps", "constant", "noisy"]
    return [gen.make(k, 150_000 + 7_000 * i, 90 + i) for i, k in enumerate(kinds * 2)]


def frames_of(series):
    import atsc_b200
    offs, lens, o0 = [], [], 0
    for s in series:
        for c in atsc_b200.chunk_sizes(len(s)):
            offs.append(o0); lens.append(c); o0 += c
    return np.concatenate(series), offs, lens


@pytest.mark.parametrize("ndev", [2, 0])  # 0: every device of the box
def test_sharded_context_matches_single_device(ndev):
    import torch
    import atsc_b200
    have = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if have < 2:
        pytest.skip("needs at least 2 CUDA devices")
    devs = list(range(have if ndev == 0 else ndev))
    flat, offs, lens = frames_of(fleet())
    one = atsc_b200.Context([0])
    many = atsc_b200.Context(devs)
    try:
        for comp, err, speed in ((atsc_b200.AUTO, 0.05, 0), (atsc_b200.AUTO, 0.03, 6), (atsc_b200.POLYNOMIAL, 0.01, 0),
                                 (atsc_b200.RLE, 0.05, 0)):
            ro, rp = one.compress_frames(flat, offs, lens, comp, err, speed, True)
            mo, mp = many.compress_frames(flat, offs, lens, comp, err, speed, True)
            assert len(mp) == len(rp) and np.array_equal(mp, rp), f"payload differs, devices {devs}, compressor {comp}"
            for i in range(len(lens)):
                a, b = mo[i], ro[i]
                assert (a.compressor, a.payload_len, a.payload_off, a.iterations, a.near_tie) == \
                       (b.compressor, b.payload_len, b.payload_off, b.iterations, b.near_tie), f"frame {i}"
            frames = [(ro[i].compressor, int(lens[i]), int(ro[i].payload_off), int(ro[i].payload_len), int(offs[i]))
                      for i in range(len(lens))]
            assert np.array_equal(many.decompress_frames(frames, rp), one.decompress_frames(frames, rp))
        # a small payload buffer reports the size that was needed, exactly like one device
        with pytest.raises(atsc_b200.AtscError):
            many.compress_frames(flat, offs, lens, atsc_b200.AUTO, 0.05, 0, True, payload_cap=1000)
        # device-resident samples are compressed by the GPU that owns them
        t = torch.from_numpy(flat).to(f"cuda:{devs[-1]}")
        do, dp = many.compress_frames(None, offs, lens, atsc_b200.AUTO, 0.05, 0, True, samples_ptr=t.data_ptr())
        ro, rp = one.compress_frames(flat, offs, lens, atsc_b200.AUTO, 0.05, 0, True)
        assert np.array_equal(dp, rp)
        with pytest.raises(atsc_b200.AtscError):   # ... and refused by a context that does not hold that GPU
            one.compress_frames(None, offs, lens, atsc_b200.AUTO, 0.05, 0, True, samples_ptr=t.data_ptr())
    finally:
        one.close()
        many.close()
