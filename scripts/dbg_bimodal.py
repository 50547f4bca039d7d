"""Why does the captured step time come in two modes (~9.6 / ~9.9 ms)?  A: re-capture the same TrainStep; B: new TrainStep objects
kept alive; C: the first object of B again.  Finding (round 1): the mode is NOT a property of the tensors or of the instantiated
graph -- the same graph object measures 9.92 ms and, seconds later, 9.61 ms (C), and candidate captures timed back to back share
the mode of the moment.  It is a state of the GPU that flips on a time scale of seconds while nvidia-smi keeps reporting
1965 MHz and no throttle reason; step-time A/B comparisons closer than 3 % have to be repeated or made at kernel level."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pivp_b200 as pk
B, T, H, W = 32, 10, 64, 64
model = pk.Model(10, is_cdna=True, scheduled_sampling_k=900.0, prefix="p", compute="bf16")
opt = pk.Adam().setup(model)
host = [torch.from_numpy(a) for a in pk.concat_examples(pk.data.synthetic_sequences(B, T, H, W))]
def timeit(step):
    for i in range(3): step(6000 + i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for i in range(5): step(6010 + i)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 5)
    return best
step = pk.TrainStep(model, opt, B, T, graph=True); step.load_batch(*host); np.random.seed(0)
print("A: same TrainStep, re-captured")
for k in range(6):
    step.graph = None
    print("   capture %d: %.3f ms  images@%x" % (k, timeit(step), step.images.data_ptr()), flush=True)
print("B: new TrainStep objects, old ones kept alive")
keep = []
for k in range(6):
    s = pk.TrainStep(model, opt, B, T, graph=True); s.load_batch(*host); keep.append(s)
    print("   object %d: %.3f ms  images@%x" % (k, timeit(s), s.images.data_ptr()), flush=True)
print("C: same object again (first of B), no re-capture")
for k in range(3):
    print("   %.3f ms" % timeit(keep[0]), flush=True)
