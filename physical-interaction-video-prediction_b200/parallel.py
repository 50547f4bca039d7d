"""Data-parallel plumbing: one process per GPU, batch rows sharded across ranks, ONE all-reduce of the
flat gradient buffer per step.  The reference has no multi-GPU path (single process, train_model.py:788,889-892);
what must stay bit-exact under sharding is its scheduled-sampling index logic (train_model.py:93-96, 649-673):
every rank draws the SAME global ``np.random.shuffle(arange(B_global))`` and keeps its rows.

Nothing here touches CUDA at import time, so the logic is testable with the gloo backend on CPU.
"""
import os

import numpy as np


def num_ground_truth(batch_size, k, iter_num):
    """train_model.py:653-655 bit-exact: int32(round(float32(B) * (k / (k + exp(iter/k))))), round-half-even, float64 inner."""
    return np.int32(np.round(np.float32(batch_size) * (k / (k + np.exp(iter_num / k)))))


def scheduled_sample_mask(batch_size, num_gt):
    """train_model.py:93-96: one legacy global-RNG shuffle of arange(B); the first num_gt entries take the ground truth."""
    idx = np.arange(int(batch_size))
    np.random.shuffle(idx)
    take = np.zeros(int(batch_size), np.int32)
    take[idx[:int(num_gt)]] = 1
    return take


def shard_rows(batch_global, rank, world_size):
    if batch_global % world_size:
        raise ValueError("global batch %d is not divisible by world size %d" % (batch_global, world_size))
    bl = batch_global // world_size
    return slice(rank * bl, (rank + 1) * bl)


def schedule_plan(batch_global, T, iter_num, k, context_frames, train=True, rank=0, world_size=1):
    """-> (feedself, take[T-1, B_local] int32 or None, n_gt).  One shuffle per step with len(gen_images) > ctx-1
    (train_model.py:663,670), consumed identically on every rank."""
    if (not train) or k == -1:
        return True, None, None
    n_gt = num_ground_truth(batch_global, k, iter_num)
    rows = shard_rows(batch_global, rank, world_size)
    take = np.zeros((T - 1, rows.stop - rows.start), np.int32)
    for t in range(T - 1):
        if t > context_frames - 1:
            take[t] = scheduled_sample_mask(batch_global, n_gt)[rows]
    return False, take, n_gt


def init_distributed(backend=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the environment)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def allreduce_sum_(flat):
    """The step's only collective: sum of the flat gradient buffer over all ranks (NCCL over NVLink on the GPU box)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat)
    return flat


class GradSync(object):
    """Gradient all-reduce overlapped with the tail of the backward pass (the reference has no multi-GPU path).

    The flat gradient buffer is cut into UNITS, each a contiguous range that becomes final at one point of ``Engine.backward``:
    ``tail`` (LayerNorm, heads, state predictor -- accumulated step by step, final when the reverse time loop ends), ``enc`` (enc0 .. enc6,
    final after their deferred weight-gradient launches), ``lstm1`` .. ``lstm7`` (final after that layer's weight-gradient GEMM) and
    ``xform`` (cdna_kerns / stp Linears).  ``ready(unit)`` is called on the compute stream right after the launches that finish the
    unit: the side stream waits for that point and sums the unit over the ranks (NCCL over NVLink) while the compute stream goes on with
    the next weight-gradient GEMM.  ``finish()`` reduces whatever is left and makes the compute stream wait for the side stream, so the
    fused 1/N + Adam kernel that follows sees the complete sum.  Everything is stream-ordered (``wait_stream`` only), hence capturable
    in the step's CUDA graph."""

    def __init__(self, engine):
        import torch
        self.e = engine
        self.side = torch.cuda.Stream(device=engine.dev)
        rng = {}
        for s in engine.specs:
            top = s.name.split("/")[0]
            if top.startswith("lstm"):
                unit = top
            elif top.startswith("enc"):
                unit = "enc"
            elif top == "model" and not s.name.startswith("model/enc7"):
                unit = "xform"
            else:
                unit = "tail"
            lo, hi = rng.get(unit, (s.offset, s.offset + s.size))
            rng[unit] = (min(lo, s.offset), max(hi, s.offset + s.size))
        # units must not interleave in the flat buffer (layout.param_specs orders them so); pad bytes between specs belong to nobody
        spans = sorted(rng.values())
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])), "gradient units interleave in the flat buffer"
        self.range = rng
        self.done = set()

    def begin(self):
        self.done = set()

    def _reduce(self, units):
        import torch
        import torch.distributed as dist
        units = [u for u in units if u in self.range and u not in self.done]
        if not units:
            return
        self.done.update(units)
        spans = sorted(self.range[u] for u in units)
        merged = [list(spans[0])]
        for a, b in spans[1:]:
            if a - merged[-1][1] <= 4:                     # adjacent up to alignment padding
                merged[-1][1] = b
            else:
                merged.append([a, b])
        cur = torch.cuda.current_stream(self.e.dev)
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            for a, b in merged:
                dist.all_reduce(self.e.flat_g[a:b])

    def ready(self, *units):
        self._reduce(units)

    def finish(self):
        import torch
        self._reduce(list(self.range))
        torch.cuda.current_stream(self.e.dev).wait_stream(self.side)
