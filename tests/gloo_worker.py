"""Worker of tests/test_host_cpu.py::test_two_rank_gloo_sharding (run under torchrun, gloo)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import atsc_b200  # noqa: E402
import gen  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    series = [gen.make(k, 300000 + 1000 * i, 11 + i) for i, k in enumerate(["gauge", "constant", "saw", "steps", "periodic"])]
    offs, lens, o = [], [], 0
    for s in series:
        for c in atsc_b200.chunk_sizes(len(s)):
            offs.append(o)
            lens.append(c)
            o += c
    flat = np.concatenate(series)
    first = atsc_b200.plan_shards(lens, world)
    mine = range(first[rank], first[rank + 1])
    # stand-in for the per-frame GPU result: a checksum per frame (frames are independent)
    local = torch.tensor([[i, float(flat[offs[i]:offs[i] + lens[i]].sum())] for i in mine], dtype=torch.float64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros((mx, 2), dtype=torch.float64)
    pad[:local.shape[0]] = local
    outs = [torch.zeros((mx, 2), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(outs, pad)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        got = torch.cat([o_[:int(s.item())] for o_, s in zip(outs, sizes)])
        assert [int(x) for x in got[:, 0]] == list(range(len(lens))), "shards must partition the frames in order"
        want = [float(flat[offs[i]:offs[i] + lens[i]].sum()) for i in range(len(lens))]
        assert np.array_equal(got[:, 1].numpy(), np.array(want))
        assert t.item() == float(world)
        print("GLOO_SHARDING_OK", first)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
