"""Per-entry-point GPU time of one eager training step (CUDA events around every C-ABI call; ~2 us overhead each)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pivp_b200 as pk

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
compute = sys.argv[2] if len(sys.argv) > 2 else "bf16"
T, H, W = 10, 64, 64
model = pk.Model(10, is_cdna=True, scheduled_sampling_k=900.0, prefix="p", compute=compute)
opt = pk.Adam().setup(model)
host = [torch.from_numpy(a) for a in pk.concat_examples(pk.data.synthetic_sequences(B, T, H, W))]
step = pk.TrainStep(model, opt, B, T, graph=False)
step.load_batch(*host)
np.random.seed(0)
for i in range(2):
    step(6000 + i)
torch.cuda.synchronize()
L = pk.lib(); orig = L.call; recs = []
def timed(name, *a):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); orig(name, *a); e.record(); recs.append((name, a, s, e))
L.call = timed
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); step(6002); ev1.record(); torch.cuda.synchronize()
L.call = orig
agg = collections.defaultdict(lambda: [0, 0.0])
for name, a, s, e in recs:
    key = name
    if name == "pivp_tc_conv5x5":
        key += " mode%d N=%d Kc=%d M=%d" % (a[9], a[7], a[5], a[2] * a[3] * a[4])
    elif name == "pivp_tc_wgrad5x5":
        key += " Cx=%d N4=%d P=%d" % (a[6], a[7], a[3] * a[4] * a[5])
    elif name in ("pivp_conv2d_fwd", "pivp_conv2d_dgrad", "pivp_conv2d_wgrad"):
        key += " " + "x".join(str(int(v)) for v in a if isinstance(v, int) and 0 < v < 100000)[:48]
    agg[key][0] += 1; agg[key][1] += s.elapsed_time(e) * 1e3
tot = sum(v[1] for v in agg.values())
print("step (events, eager) %.2f ms; sum of calls %.2f ms; calls %d" % (ev0.elapsed_time(ev1), tot / 1e3, len(recs)))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    print("%-78s n=%4d total=%8.1f us avg=%7.1f share=%4.1f%%" % (k[:78], n, t, t / n, 100 * t / tot))
