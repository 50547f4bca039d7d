"""Per-layer device time of the ConvLSTM tcgen05 launches of one b32 training step (forward with fused gates, input gradient, weight gradient).

One eager step records every pivp_tc_conv5x5 / pivp_tc_wgrad5x5 call with its real arguments; the calls of each (layer, direction) are then
replayed back to back as one CUDA graph (9 launches, one per time step, each on its own buffers) and timed with CUDA events.
    python scripts/halo_layers.py [--batch 32] [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pivp_b200 as pk

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--json", default=None)
args = ap.parse_args()
B, T, H, W = args.batch, 10, 64, 64
L = pk.lib()
model = pk.Model(10, is_cdna=True, scheduled_sampling_k=900.0, prefix="p", height=H, width=W, compute="bf16")
opt = pk.Adam().setup(model)
step = pk.TrainStep(model, opt, B, T, graph=False)
host = [torch.from_numpy(a) for a in pk.concat_examples(pk.data.synthetic_sequences(B, T, H, W, seed=1234))]
step.load_batch(*host)
np.random.seed(99)
step(6000); torch.cuda.synchronize()
orig, recs = L.call, []


def rec(name, *a):
    if name in ("pivp_tc_conv5x5", "pivp_tc_conv5x5_ln", "pivp_tc_wgrad5x5"):
        recs.append((name, a))
    orig(name, *a)


L.call = rec
step(6001); torch.cuda.synchronize()
L.call = orig
tc = model.engine.tc
from pivp_b200 import layout
lv = (2, 2, 4, 4, 8, 4, 2)
flops = [2.0 * B * (H // l) * (W // l) * 4 * c * 25 * (cin + c) for cin, c, l in zip(layout.LSTM_IN, layout.LSTM_SIZES, lv)]
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}


def replay(calls, reps=7):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream().cuda_stream
        for name, a in calls:
            orig(name, *(a[:-1] + (cur,)))
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best / len(calls)


rows = []
tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
for li in range(7):
    for kind, ptr in (("fwd", tc.Wf[li].data_ptr()), ("dgrad", tc.Wd[li].data_ptr())):
        calls = [(n, a) for n, a in recs if n in ("pivp_tc_conv5x5", "pivp_tc_conv5x5_ln") and a[6] == ptr]
        per = replay(calls)
        tot[kind] += per * len(calls)
        rows.append(dict(layer=li + 1, kind=kind, launches=len(calls), us=per * 1e6, tflops=flops[li] / per / 1e12,
                         frac=flops[li] / per / 1e12 / peaks["bf16_tflops"]))
    calls = [(n, a) for n, a in recs if n == "pivp_tc_wgrad5x5" and a[1] == tc.xh_all[li].data_ptr()]
    per = replay(calls)
    tot["wgrad"] += per
    rows.append(dict(layer=li + 1, kind="wgrad", launches=1, us=per * 1e6, tflops=flops[li] * (T - 1) / per / 1e12,
                     frac=flops[li] * (T - 1) / per / 1e12 / peaks["bf16_tflops"]))
print("| layer | kind | launches/step | us/launch | TFLOP/s | frac of %.0f |" % peaks["bf16_tflops"])
print("|---|---|---:|---:|---:|---:|")
for r in rows:
    print("| lstm%d | %s | %d | %.2f | %.0f | %.3f |" % (r["layer"], r["kind"], r["launches"], r["us"], r["tflops"], r["frac"]))
print("per step: fwd %.3f ms, dgrad %.3f ms, wgrad %.3f ms" % (tot["fwd"] * 1e3, tot["dgrad"] * 1e3, tot["wgrad"] * 1e3))
allc = [(n, a) for n, a in recs if n in ("pivp_tc_conv5x5", "pivp_tc_conv5x5_ln")]
per = replay(allc)
fl = sum(flops) * 2 * (T - 1) / len(allc)
print("all %d fwd+dgrad launches in one graph: %.2f us/launch, %.0f TFLOP/s, frac %.3f" % (len(allc), per * 1e6, fl / per / 1e12, fl / per / 1e12 / peaks["bf16_tflops"]))
if args.json:
    json.dump(dict(rows=rows, totals_ms={k: v * 1e3 for k, v in tot.items()}, all_avg_us=per * 1e6, all_frac=fl / per / 1e12 / peaks["bf16_tflops"]),
              open(args.json, "w"), indent=1)
