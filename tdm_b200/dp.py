"""Data-parallel gradient exchange: the one collective of the hot path.

The reference gets it from accelerate -> torch DDP (main.py:321-323, 535): 25 MB autograd-hook buckets.
Here the backward is hand-sequenced (28 block-backward C calls in reverse order), every model's
gradients live in flat fp32 arenas laid out block by block, and `on_block_done(i)` fires when block i's
slice is final -- so a bucket is simply a contiguous arena slice, all-reduced (SUM) by NCCL over
NVLink/NVSwitch on the process group's communication stream while the remaining blocks' backward kernels
keep the SMs busy.  The 1/world_size average is folded into the clip coefficient of the fused optimizer
kernel, so no extra pass touches the gradients.  Batch shards across ranks; nothing else is exchanged.

Works on any torch.distributed backend (NCCL on the GPUs; gloo for the CPU tests of the bucketing logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def block_regions(arena, depth):
    """(start, end) element ranges of an _Arena: [head][block 0]..[block depth-1][tail]."""
    names = list(arena.layout)
    starts = {}
    for n in names:
        off = arena.layout[n][0]
        if n.startswith("transformer_blocks."):
            i = int(n.split(".")[1])
            starts.setdefault(i, off)
    order = sorted(starts)
    if not order:
        return [(0, arena.size)], {}, None
    bounds = [starts[i] for i in order]
    last_block_end = arena.size
    for n in names:  # first non-block name after the last block
        off = arena.layout[n][0]
        if not n.startswith("transformer_blocks.") and off > bounds[-1]:
            last_block_end = min(last_block_end, off)
    regions = {}
    for j, i in enumerate(order):
        end = bounds[j + 1] if j + 1 < len(bounds) else last_block_end
        regions[i] = (bounds[j], end)
    head = (0, bounds[0])
    tail = (last_block_end, arena.size)
    return [r for r in (head, tail) if r[1] > r[0]], regions, order


class GradSync:
    """Bucketed, overlapped all-reduce of flat gradient arenas."""

    def __init__(self, group=None, bucket_blocks=2, lora_bucket_blocks=7):
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket_blocks = bucket_blocks
        self.lora_bucket_blocks = lora_bucket_blocks
        self._pending = []
        self._plans = {}
        self.launched = 0

    # -- planning ---------------------------------------------------------------------------------
    def _plan(self, model):
        key = id(model)
        if key in self._plans:
            return self._plans[key]
        plan = []  # (flat grad tensor, rest regions, block regions, blocks per bucket)
        if model.train_base and model._grad is not None:
            rest, regions, _ = block_regions(model._arena, model.depth)
            plan.append((model._grad, rest, regions, self.bucket_blocks))
        if model._lgrad is not None:
            rest, regions, _ = block_regions(model._lora_arena, model.depth)
            plan.append((model._lgrad, rest, regions, self.lora_bucket_blocks))
        self._plans[key] = plan
        return plan

    def _reduce(self, t):
        if self.world_size == 1:
            return
        self._pending.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.launched += 1

    # -- hooks -------------------------------------------------------------------------------------
    def hook(self, model):
        """on_block_done(i) for Transformer2DModel.backward_impl: blocks finish in reverse order; every
        `blocks per bucket` finished blocks form one contiguous slice that is reduced at once."""
        plan = self._plan(model)

        def on_block_done(i):
            for flat, _, regions, per in plan:
                if i not in regions:
                    continue
                if i % per == 0:  # lowest block of its bucket: the bucket [i, i+per) is complete
                    hi = min(i + per - 1, max(regions))
                    self._reduce(flat[regions[i][0]:regions[hi][1]])
        return on_block_done

    def finish(self, model):
        """Reduce what is not inside a block (embedders, final layer) once the whole backward has run."""
        for flat, rest, _, _ in self._plan(model):
            for a, b in rest:
                self._reduce(flat[a:b])

    def wait(self):
        for w in self._pending:
            w.wait()
        self._pending = []


def init_from_env(backend="nccl"):
    """RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment (torchrun); one process per GPU."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def broadcast_params(models, src=0, group=None):
    """DDP's one-off parameter broadcast on the flat arenas."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for m in models:
        dist.broadcast(m._flat, src=src, group=group)
        if m._lflat is not None:
            dist.broadcast(m._lflat, src=src, group=group)
        m._flat16_version = m._lflat16_version = -1  # re-cast the bf16 copies
