"""Compare backward intermediates of time step 0 (T=3: the last step the reverse sweep visits) with the oracle's Var.grad."""
import sys, os
os.environ.setdefault("PIVP_BRANCHES", "")     # stage-by-stage diagnostics: one stream, so that the stop hook leaves a consistent state
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import test_gpu_model as TM
from oracle import model as OM
from oracle import npgrad as G
import pivp_b200 as pk
pk.lib()
mt, nm, k, H, W, B, T, oob, use_state = ("CDNA", 10, 900.0, 64, 64, 2, 3, "zeros", True)
cfg = OM.Config(mt, nm, schedsamp_k=k, height=H, width=W, use_state=use_state, stp_oob=oob, dtype=np.float64)
params = TM.perturbed(cfg)
batch = OM.concat_examples(OM.synthetic_sequences(B, T, cfg))
np.random.seed(99)
ref = OM.forward(params, batch, 6000, cfg, take_gt_log=[])
G.KEEP_INTERMEDIATE = True
G.backward(ref["loss"])
model = TM.make_model(pk, mt, nm, k, H, W, oob, use_state=use_state)
model.load_params(params)
np.random.seed(99)
loss = model([torch.from_numpy(a) for a in batch], 6000)
model.engine._stop_after_step = lambda t: t == T - 2      # diagnostics hook: leave step T-2's backward temporaries in place
model.cleargrads(); model.backward(); torch.cuda.synchronize()
ws = model.engine.ws
tr = ref["trace"][T - 2]
def nhwc(t, C): return t.reshape(B, H, W, C).permute(0, 3, 1, 2)
def show(name, got, want):
    want = np.asarray(want, np.float64); got = got.detach().cpu().numpy().astype(np.float64).reshape(want.shape)
    d = np.abs(got - want)
    i = np.unravel_index(d.argmax(), d.shape)
    print("%-14s rel %.2e  max|ref| %.3e  worst at %s got %.6e want %.6e  #>1e-4rel: %d" % (name, d.max() / (np.abs(want).max() + 1e-30), np.abs(want).max(), i, got[i], want[i], int((d > 1e-4 * np.abs(want).max()).sum())))
for key in ("mask_pre", "enc7_pre", "kern_raw"):
    g = getattr(tr[key], "saved_grad", None)
    print(key, None if g is None else g.shape)
show("d_mask_pre", ws["d_mask_pre"], tr["mask_pre"].saved_grad)
show("d_enc7_pre", ws["d_enc7_pre"], tr["enc7_pre"].saved_grad)
show("d_kern_raw", ws["d_kern_raw"], tr["kern_raw"].saved_grad)
show("d_e6", nhwc(ws["d_e6"], 64), tr["enc6"].saved_grad)
show("d_hid5", ws["d_hid5"][(T - 2) & 1].reshape(B, H // 8, W // 8, 128).permute(0, 3, 1, 2), tr["hidden5"].saved_grad) if hasattr(tr["hidden5"], "saved_grad") else None
def lvl(t, lv, cs, co, C):
    h, w = H // lv, W // lv
    return t.reshape(B, h, w, cs)[..., co:co + C].permute(0, 3, 1, 2)
hs = tr["hiddens"]
show("g hidden7", lvl(ws["d_cat6"][(T - 2) & 1], 2, 64, 0, 32), hs[6].saved_grad)
show("g hidden6", lvl(ws["d_cat5"], 4, 96, 0, 64), hs[5].saved_grad)
show("g hidden5", lvl(ws["d_hid5"][(T - 2) & 1], 8, 128, 0, 128), hs[4].saved_grad)
show("g hidden4", lvl(ws["d_hid4"], 4, 64, 0, 64), hs[3].saved_grad)
show("g hidden3", lvl(ws["dxh"][3], 4, 128, 0, 64), hs[2].saved_grad)
show("g hidden2", lvl(ws["d_hid2"], 2, 32, 0, 32), hs[1].saved_grad)
show("g hidden1", lvl(ws["dxh"][1], 2, 64, 0, 32), hs[0].saved_grad)
grads = model.grads
worst = {}
for key, v in ref["P"].items():
    r = np.zeros_like(v.data) if v.grad is None else v.grad
    worst[key] = np.abs(grads[key].astype(np.float64) - r).max() / (np.abs(r).max() + 1e-20)
for k_, e in sorted(worst.items(), key=lambda x: -x[1])[:6]:
    print("  %-32s %.3e" % (k_, e))
