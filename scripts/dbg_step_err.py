"""Print the worst per-tensor gradient errors of one whole-step parity case (same code path as tests/test_gpu_model.py)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import test_gpu_model as TM
from oracle import model as OM
from oracle import npgrad as G
import pivp_b200 as pk
pk.lib()
idx = int(sys.argv[1]) if len(sys.argv) > 1 else 5
mt, nm, k, H, W, B, T, oob, use_state = TM.CASES[idx]
cfg = OM.Config(mt, nm, schedsamp_k=k, height=H, width=W, use_state=use_state, stp_oob=oob)
params = TM.perturbed(cfg)
batch = OM.concat_examples(OM.synthetic_sequences(B, T, cfg))
np.random.seed(99)
ref = OM.forward(params, batch, 6000, cfg, take_gt_log=[])
G.backward(ref["loss"])
model = TM.make_model(pk, mt, nm, k, H, W, oob, use_state=use_state)
model.load_params(params)
np.random.seed(99)
loss = model([torch.from_numpy(a) for a in batch], 6000)
model.cleargrads(); model.backward(); torch.cuda.synchronize()
grads = model.grads
worst = {}
for key, v in ref["P"].items():
    r = np.zeros_like(v.data) if v.grad is None else v.grad
    worst[key] = np.abs(grads[key].astype(np.float64) - r).max() / (np.abs(r).max() + 1e-20)
print(TM.CASES[idx], "loss rel err %.2e" % (abs(float(loss) - float(ref["loss"].data)) / abs(float(ref["loss"].data))))
for k_, e in sorted(worst.items(), key=lambda x: -x[1])[:8]:
    print("  %-32s %.3e" % (k_, e))
for t in range(T - 1):
    print("t=%d gen %.2e mask_pre %.2e state %.2e" % (t, TM.rel(model.gen_images[t], ref["gen_images"][t].data),
          TM.rel(model.engine.ws["mask_pre"][t], ref["trace"][t]["mask_pre"].data), TM.rel(model.gen_states[t], ref["gen_states"][t].data)))
print("trace keys", sorted(ref["trace"][0].keys()))
for k_, e in sorted(worst.items(), key=lambda x: -x[1])[:3]:
    r = ref["P"][k_].grad; gq = grads[k_].astype(np.float64); d = np.abs(gq - r); m = np.abs(r).max()
    bad = np.argwhere(d > 1e-4 * m)
    print(k_, r.shape, "bad elements", len(bad), "of", r.size, "first", bad[:6].tolist(), "vals", [(float(gq[tuple(i)]), float(r[tuple(i)])) for i in bad[:3]])
