"""Device time of the phases of the b32 training step, each captured as its own CUDA graph: forward, forward + BPTT (no deferred weight
gradients are separable, so: forward + backward), full step.  python scripts/phase_times.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pivp_b200 as pk
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T, H, W = 10, 64, 64
model = pk.Model(10, is_cdna=True, scheduled_sampling_k=900.0, prefix="p", compute="bf16")
opt = pk.Adam().setup(model)
host = [torch.from_numpy(a) for a in pk.concat_examples(pk.data.synthetic_sequences(B, T, H, W))]
step = pk.TrainStep(model, opt, B, T, graph=False)
step.load_batch(*host)
np.random.seed(0)
step(6000); torch.cuda.synchronize()
e = model.engine


def timed(fn, n=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


fwd = timed(lambda: e.forward_device(step.images, step.actions, step.states, False))
def fb():
    e.forward_device(step.images, step.actions, step.states, False); e.cleargrads(); e.backward()
fwbw = timed(fb)
def full():
    fb(); step._update()
allt = timed(full)
print("forward %.3f ms | forward + backward (BPTT + deferred weight gradients) %.3f ms | + Adam and bf16 weight refresh %.3f ms" % (fwd, fwbw, allt))
print("backward alone: %.3f ms; update: %.3f ms" % (fwbw - fwd, allt - fwbw))
