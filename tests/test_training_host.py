"""Host-side logic of the iteration (no GPU): repaired level arithmetic against the oracle's restatement,
draw order against the reference's, arena bucketing, and the N>1 gradient exchange on gloo."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def test_repaired_levels_match_oracle_and_are_valid():
    from tdm_b200.training import repaired_levels as ours
    from oracle.iteration import repaired_levels as ref
    g = torch.Generator().manual_seed(0)
    for sep in (True, False):
        ind_t = torch.randint(1, 5, (4096,), generator=g)
        hi = 440 if sep else 900
        t_fake = torch.randint(240, hi, (4096,), generator=g)
        a, b = ours(ind_t, t_fake, 900, sep), ref(ind_t, t_fake, 900, sep)
        for x, y in zip(a, b):
            assert torch.equal(x, y) and x.dtype == torch.int64
        k, t1, t2, t_in = a
        assert (t2 >= t1).all() and (t2 <= 999).all() and (t1 >= 0).all()       # D3 repaired: never NaN
        assert set(t1.tolist()) == {0, 225, 450, 675}
        assert set(t_in.tolist()) == {224, 449, 674, 899}                          # timesteps generate_new visits
        assert torch.equal(t_in, (k + 1) * 225 - 1)
        if sep:
            assert ((t2 - t1) < 200).all()


def test_draw_order_is_the_references():
    """main.py:473-478 restated twice (product host code, oracle) must consume the generator identically."""
    from tdm_b200.training import draw_fake_branch
    from oracle.iteration import fake_branch_draws
    for sep in (True, False):
        torch.manual_seed(5)
        noise = torch.randn(3, 4, 8, 8).half()
        a = draw_fake_branch(noise, 900, sep)
        torch.manual_seed(5)
        noise = torch.randn(3, 4, 8, 8).half()
        b = fake_branch_draws(noise, 900, sep)
        assert all(torch.equal(x, y) for x, y in zip(a, b))


class _FakeArena:
    def __init__(self, names_sizes):
        self.layout, self.size = {}, 0
        for n, s in names_sizes:
            self.layout[n] = (self.size, (s,))
            self.size += s


def _arena(depth=4):
    items = [("pos_embed.w", 16), ("adaln.w", 8)]
    for i in range(depth):
        items += [(f"transformer_blocks.{i}.a", 32), (f"transformer_blocks.{i}.b", 8)]
    items += [("scale_shift_table", 4), ("proj_out.w", 12)]
    return _FakeArena(items)


def test_block_regions_partition_the_arena():
    from tdm_b200.dp import block_regions
    ar = _arena(4)
    rest, regions, order = block_regions(ar, 4)
    assert order == [0, 1, 2, 3]
    spans = sorted(rest + list(regions.values()))
    assert spans[0][0] == 0 and spans[-1][1] == ar.size
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c
    assert regions[0] == (24, 64) and regions[3] == (144, 184)


class _M:  # the attributes GradSync touches on a model
    def __init__(self, depth, rank):
        self._arena, self.depth, self.train_base = _arena(depth), depth, True
        self._grad = torch.arange(self._arena.size, dtype=torch.float32) * (rank + 1)
        self._lora_arena, self._lgrad = None, None


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tdm_b200.dp import GradSync
    m = _M(5, rank)
    gs = GradSync(bucket_blocks=2)
    hook = gs.hook(m)
    for i in reversed(range(5)):
        hook(i)
    gs.finish(m)
    gs.wait()
    expect = torch.arange(m._arena.size, dtype=torch.float32) * sum(r + 1 for r in range(world))
    q.put((rank, bool(torch.equal(m._grad, expect)), gs.launched))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(n == 3 + 2 for _, _, n in res)  # buckets {4}, {2,3}, {0,1} + head + tail
