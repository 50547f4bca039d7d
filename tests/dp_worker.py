"""Worker of tests/test_gpu_dp.py: one process per GPU (torchrun), real NCCL all-reduce.

Each rank trains its shard of ONE global batch for STEPS steps through TrainStep (CUDA graphs + the NCCL all-reduce of the flat
gradient buffer), then the ranks compare:
  * flat parameters bit-identical on every rank after STEPS steps (same all-reduced gradient, same Adam),
  * the all-reduced gradient of step 1 (x 1/N) == the gradient of the single-process step on the whole global batch,
  * per-step global loss (mean of the rank losses) == the single-process loss,
  * the scheduled-sampling selects of the ranks concatenate to the single-process select (one global permutation, sliced).
Prints ``DP_OK`` on rank 0 when all of it holds.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pivp_b200 as pk  # noqa: E402
from pivp_b200 import parallel  # noqa: E402
from oracle import model as OM  # noqa: E402

STEPS = 3


def run(compute, H, b_local, T):
    rank, local, world = parallel.init_distributed()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    Bg = b_local * world
    cfg = OM.Config("CDNA", 10, schedsamp_k=900.0, height=H, width=H)
    params = OM.init_params(cfg)
    batch = OM.concat_examples(OM.synthetic_sequences(Bg, T, cfg))
    rows = parallel.shard_rows(Bg, rank, world)

    def train(model, data, B):
        opt = pk.Adam().setup(model)
        step = pk.TrainStep(model, opt, B, T, graph=True)
        step.load_batch(*[torch.from_numpy(np.ascontiguousarray(a)) for a in data], non_blocking=False)
        np.random.seed(21)
        losses, takes, g1 = [], [], None
        for i in range(STEPS):
            losses.append(float(step(6000 + 100 * i)))
            takes.append(model.take_gt.copy())
            if i == 0:
                torch.cuda.synchronize()
                g1 = model.engine.flat_g.clone()
        torch.cuda.synchronize()
        return losses, takes, g1, model.engine.flat_p.clone()

    m = pk.Model(10, is_cdna=True, scheduled_sampling_k=900.0, prefix="dp", height=H, width=H, device=str(dev), compute=compute,
                 rank=rank, world_size=world)
    m.load_params(params)
    losses, takes, g1, p = train(m, [a[:, rows] for a in batch], b_local)
    # ---- ranks agree bit for bit
    plist = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(plist, p)
    same = all(torch.equal(plist[0], q) for q in plist)
    lt = torch.tensor(losses, dtype=torch.float64, device=dev)
    dist.all_reduce(lt)
    tk = torch.from_numpy(np.stack(takes)).to(dev)                    # (STEPS, T-1, b_local)
    tlist = [torch.empty_like(tk) for _ in range(world)]
    dist.all_gather(tlist, tk)
    ok = True
    if rank == 0:
        # ---- single-process run on the whole global batch
        full = pk.Model(10, is_cdna=True, scheduled_sampling_k=900.0, prefix="full", height=H, width=H, device=str(dev), compute=compute)
        full.load_params(params)
        fl, ft, fg1, fp = train(full, batch, Bg)
        gl = (lt / world).cpu().numpy()
        tol = 1e-5 if compute == "f32" else 5e-3
        gerr = float((g1 / world - fg1).abs().max() / fg1.abs().max())
        lerr = float(np.abs(gl - np.array(fl)).max() / abs(fl[0]))
        takes_ok = np.array_equal(torch.cat(tlist, dim=2).cpu().numpy(), np.stack(ft))
        perr = float((p - fp).abs().max())
        print("[dp %s H=%d world=%d] ranks identical: %s | grad step 1 vs full batch: %.2e | loss: %.2e | selects equal: %s | max |dp - full| params after %d steps: %.2e"
              % (compute, H, world, same, gerr, lerr, takes_ok, STEPS, perr), flush=True)
        ok = same and takes_ok and gerr < (1e-4 if compute == "f32" else 2e-2) and lerr < tol
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    return bool(flag.item())


if __name__ == "__main__":
    ok = run("f32", 32, 2, 4) and run("bf16", 64, 2, 4)
    if int(os.environ.get("RANK", "0")) == 0:
        print("DP_OK" if ok else "DP_FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)
