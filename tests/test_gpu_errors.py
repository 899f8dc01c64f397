"""Error behaviour of the C ABI on a GPU: where the reference panics (`todo!()`, `unwrap()`, index
out of bounds -- SURVEY.md 8b "Errors") the library returns a status code and stays usable."""
import numpy as np
import pytest

import gen
import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import atsc_b200
    c = atsc_b200.Context()
    yield c
    c.close()


def still_works(ctx):
    x = gen.make("gauge", 5000, 3)
    b, o = ctx.compress_bounded(O.POLYNOMIAL, x, 0.05)
    want, _, _ = O.compress_bounded(O.POLYNOMIAL, x, float(np.float32(0.05)))
    assert b == want


def code_of(fn):
    import atsc_b200
    with pytest.raises(atsc_b200.AtscError) as ei:
        fn()
    return ei.value.code


def test_compress_argument_errors(ctx):
    x = gen.make("util", 4096, 1)
    # Compressor::Auto has no unbounded compress (reference: todo!(), compressor/mod.rs:72)
    assert code_of(lambda: ctx.compress_frames(x, [0], [4096], O.AUTO, 0.05, 0, bounded=False)) == 4
    # empty frame (reference: index panic, optimizer/utils.rs:41) and a frame above MAX_FRAME_SIZE
    assert code_of(lambda: ctx.compress_frames(x, [0], [0], O.AUTO, 0.05, 0, True)) == 1
    big = np.zeros(131073)
    assert code_of(lambda: ctx.compress_frames(big, [0], [131073], O.AUTO, 0.05, 0, True)) == 1
    assert code_of(lambda: ctx.compress_frames(x, [0], [4096], O.AUTO, 0.05, 7, True)) == 1      # -c 0..6
    assert code_of(lambda: ctx.compress_frames(x, [0], [4096], 9, 0.05, 0, True)) == 1           # unknown compressor
    # unbounded FFT on a length that is not 2^a 3^b (not reachable from the CLI)
    assert code_of(lambda: ctx.compress_frames(np.arange(130.0), [0], [130], O.FFT, 0.0, 0, bounded=False)) == 4
    still_works(ctx)


def test_payload_capacity_error_reports_need(ctx):
    import atsc_b200
    x = gen.make("noisy", 20000, 2)
    with pytest.raises(atsc_b200.AtscError) as ei:
        ctx.compress_frames(x, [0], [20000], O.NOOP, 0.0, 0, False, payload_cap=100)
    assert ei.value.code == 3  # ATSC_ERR_CAPACITY
    out, pay = ctx.compress_frames(x, [0], [20000], O.NOOP, 0.0, 0, False)
    assert pay.tobytes() == O.compress(O.NOOP, x)
    still_works(ctx)


@pytest.mark.parametrize("comp", [O.CONSTANT, O.NOOP, O.RLE, O.POLYNOMIAL, O.IDW, O.FFT])
def test_decompress_malformed_payloads(ctx, comp):
    """Truncated / corrupted payloads: ATSC_ERR_FORMAT (the reference's bincode decode panics)."""
    x = gen.make("gauge", 600, 4)
    good = O.compress_bounded(comp, x, float(np.float32(0.03)))[0] if comp in (O.POLYNOMIAL, O.IDW, O.FFT) else O.compress(comp, x)
    ok = ctx.decompress_frames([(comp, 600, 0, len(good), 0)], np.frombuffer(good, dtype=np.uint8))
    assert len(ok) == 600
    for cut in (1, 2, len(good) // 2, len(good) - 1):
        if cut >= len(good):
            continue
        bad = np.frombuffer(good[:cut], dtype=np.uint8)
        assert code_of(lambda: ctx.decompress_frames([(comp, 600, 0, cut, 0)], bad)) == 5, (comp, cut)
    # frame says more payload than the buffer holds
    assert code_of(lambda: ctx.decompress_frames([(comp, 600, 0, len(good) + 8, 0)], np.frombuffer(good, dtype=np.uint8))) == 1
    # Auto is not a frame compressor (reference: todo!(), compressor/mod.rs:117)
    assert code_of(lambda: ctx.decompress_frames([(O.AUTO, 600, 0, len(good), 0)], np.frombuffer(good, dtype=np.uint8))) == 4
    still_works(ctx)


def test_decompress_bad_stream(ctx):
    import atsc_b200
    x = gen.make("steps", 3000, 6)
    bro = ctx.compress_data([x], compressor=O.AUTO, error=5)[0]
    assert np.array_equal(ctx.decompress_data([bro])[0], O.decompress_stream(bro))
    with pytest.raises(atsc_b200.AtscError):
        ctx.decompress_data([b"XXXX" + bro[4:]])       # bad magic (reference: panic, header.rs:34)
    with pytest.raises(atsc_b200.AtscError):
        ctx.decompress_data([bro[:len(bro) // 2]])     # truncated stream
    still_works(ctx)
