"""In-graph marginal cost of each C-ABI entry point: capture the training step as a CUDA graph with the calls of one
entry point dropped and report how much faster the replay gets.  (Results of such a step are garbage -- this is a
timing tool.  ncu's per-kernel times are cold-cache and serialised; this is what a kernel costs inside the real step.)

    python scripts/marginal_cost.py [name ...]       default: every entry point that launches >= 9 times per step"""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pivp_b200 as pk

B, T, H, W = 32, 10, 64, 64
model = pk.Model(10, is_cdna=True, scheduled_sampling_k=900.0, prefix="p", compute="bf16")
opt = pk.Adam().setup(model)
host = [torch.from_numpy(a) for a in pk.concat_examples(pk.data.synthetic_sequences(B, T, H, W))]
L = pk.lib()
orig = L.call


saved = None


def step_ms(skip=()):
    global saved
    counts = collections.Counter()
    e = model.engine
    if saved is None:
        e._workspace(B, T)
        saved = (e.flat_p.clone(), opt.m.clone(), opt.v.clone(), opt.step.clone())
    else:                                   # a step with dropped kernels leaves NaNs in the parameters: start every run from the same state
        e.flat_p.copy_(saved[0]); opt.m.copy_(saved[1]); opt.v.copy_(saved[2]); opt.step.copy_(saved[3])
        e.params_changed()

    def call(name, *a):
        counts[name] += 1
        if name in skip:
            return 0
        return orig(name, *a)
    L.call = call
    step = pk.TrainStep(model, opt, B, T, graph=True)
    step.load_batch(*host)
    np.random.seed(0)
    for i in range(3):
        step(6000 + i)
    torch.cuda.synchronize()
    L.call = orig
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for i in range(5):
            step(6010 + i)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 5)
    return best, counts


base, counts = step_ms()
per_step = {k: v // 4 for k, v in counts.items()}           # one eager + the captured step + ... : relative counts only
print("baseline %.3f ms/step" % base)
names = sys.argv[1:] or [k for k, v in sorted(counts.items(), key=lambda kv: -kv[1]) if v >= 4]
rows = []
for n in names:
    ms, _ = step_ms((n,))
    base, _ = step_ms()                     # baseline re-measured next to every run (clock / state drift)
    rows.append((base - ms, n, counts[n]))
    print("  without %-34s %8.3f ms vs %8.3f (saves %6.3f ms, %d calls seen)" % (n, ms, base, base - ms, counts[n]), flush=True)
print("sum of marginal costs %.3f ms of %.3f" % (sum(r[0] for r in rows), base))
