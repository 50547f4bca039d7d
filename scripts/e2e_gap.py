"""Where the end-to-end number loses against the resident-input number: wall-clock ms per step of the benchmark step (CDNA 64x64 b32 T=10,
bf16, CUDA graph) under increasingly complete host loops.  Usage: python scripts/e2e_gap.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__

__graft_entry__.build()
import pivp_b200 as pk

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B, T, H, W = 32, 10, 64, 64
dev = torch.device("cuda", 0)
model = pk.Model(10, is_cdna=True, scheduled_sampling_k=900.0, prefix="train", height=H, width=W, device=str(dev), compute="bf16")
opt = pk.Adam(alpha=0.001).setup(model)
host = [torch.from_numpy(a).pin_memory() for a in pk.concat_examples(pk.data.synthetic_sequences(B, T, H, W, seed=1234))]
step = pk.TrainStep(model, opt, B, T, graph=True)
step.load_batch(*host)
np.random.seed(99)
it = 6000
for _ in range(5):
    step(it); it += 1
torch.cuda.synchronize()
dev_copy = [h.to(dev) for h in host]


def timed(name, body):
    global it
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    body()
    torch.cuda.synchronize()
    print("%-58s %.3f ms/step" % (name, (time.perf_counter() - t0) / STEPS * 1e3), flush=True)


def a_step_only():
    global it
    for _ in range(STEPS):
        step(it); it += 1


def b_loss_late():
    global it
    pending = None
    for _ in range(STEPS):
        h = step(it); it += 1
        if pending is not None:
            float(pending)
        pending = h
    float(pending)


def c_d2d():
    global it
    pending = None
    for _ in range(STEPS):
        step.load_batch(*dev_copy)
        h = step(it); it += 1
        if pending is not None:
            float(pending)
        pending = h
    float(pending)


def d_prefetch(order):
    def run():
        global it
        pf = pk.BatchPrefetcher(step, (host for _ in range(STEPS)))
        pending = None
        while pf.load_next(prefetch=(order == "before")):
            h = step(it); it += 1
            if order == "after":
                pf.prefetch()
            if pending is not None:
                float(pending)
            pending = h
        float(pending)
    return run


def e_sync_copy():
    global it
    for _ in range(STEPS):
        step.load_batch(*host)
        float(step(it)); it += 1


for rep in range(2):
    timed("A step only (inputs resident, no loss read)", a_step_only)
    timed("B + loss read one step late", b_loss_late)
    timed("C + device-to-device load_batch", c_d2d)
    timed("D BatchPrefetcher, next H2D issued before the step", d_prefetch("before"))
    timed("E BatchPrefetcher, next H2D issued after the step", d_prefetch("after"))
    timed("F synchronous load_batch from pinned host + float(loss)", e_sync_copy)
