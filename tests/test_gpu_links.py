"""The reference's link surface, to the letter: ``Stateless*.__call__(encs, hiddens, batch_size, prev_image, num_masks,
color_channels) -> (transformed_list, enc7)`` (train_model.py:293, 368, 434), ``Model.conv_res = encs`` (:734, eight NCHW tensors), the
``_state_cost`` summary lines (:752), and checkpoint files compatible with ``serializers.save_npz / load_npz`` (:864-869, 1035-1037)."""
import os

import numpy as np
import pytest
import torch

from oracle import model as OM
from oracle import npgrad as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pk():
    import pivp_b200
    pivp_b200.lib()
    return pivp_b200


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.abs(a.astype(np.float64) - b).max() / (np.abs(b).max() + 1e-30))


def oracle_step(mt, nm, H=32, B=2, T=3, oob="zeros"):
    cfg = OM.Config(mt, nm, schedsamp_k=900.0, height=H, width=H, stp_oob=oob, dtype=np.float64)
    params = OM.init_params(cfg)
    rs = np.random.RandomState(3)
    for k in sorted(params):                                   # biases off zero so that they matter
        if k.endswith("/b"):
            params[k] = params[k] + 0.1 * rs.standard_normal(params[k].shape)
    batch = OM.concat_examples(OM.synthetic_sequences(B, T, cfg))
    np.random.seed(99)
    return cfg, params, batch, OM.forward(params, batch, 6000, cfg)


@pytest.mark.parametrize("mt,nm,oob", [("CDNA", 10, "zeros"), ("CDNA", 4, "zeros"), ("DNA", 1, "zeros"), ("STP", 10, "zeros"), ("STP", 5, "border")])
def test_stateless_links_reference_call_form(pk, mt, nm, oob):
    cfg, params, batch, ref = oracle_step(mt, nm, oob=oob)
    tr = ref["trace"][-1]
    dev = "cuda"
    t32 = lambda v: torch.from_numpy(np.ascontiguousarray(v.data, dtype=np.float32)).to(dev)
    encs, hiddens, prev = [t32(e) for e in tr["encs"]], [t32(h) for h in tr["hiddens"]], t32(tr["prev_image"])
    link = {"CDNA": pk.StatelessCDNA, "DNA": pk.StatelessDNA, "STP": pk.StatelessSTP}[mt](nm).load(params)
    B = prev.shape[0]
    if mt == "STP":
        transformed, enc7 = link(encs, hiddens, B, prev, nm, 3, oob=oob)
    else:
        transformed, enc7 = link(encs, hiddens, B, prev, nm, 3)
    torch.cuda.synchronize()
    want = tr["transformed"]
    assert len(transformed) == len(want) == {"CDNA": nm + 1, "DNA": 1, "STP": nm}[mt]
    for got, w in zip(transformed, want):
        assert tuple(got.shape) == w.data.shape and rel(got, w.data) < 2e-5
    e7 = np.maximum(tr["enc7_pre"].data, 0) if mt != "STP" else tr["enc7_pre"].data       # ref:315, 388 relu; ref:454 none
    assert rel(enc7, e7) < 2e-5
    # the composite of the list (ref:725-728, zip truncation) equals the fused form on the same pre-activations
    a_pre = t32(tr["mask_pre"])
    args = {"CDNA": lambda: (prev, link.enc7_pre, a_pre, link.kern_raw), "DNA": lambda: (prev, link.enc7_pre, a_pre),
            "STP": lambda: (prev, link.enc7_pre, a_pre, link.theta_raw, oob)}[mt]()
    fused = link(*args)
    torch.cuda.synchronize()
    assert rel(fused, tr["output"].data) < 2e-5


def test_dna_link_rejects_more_than_one_mask(pk):
    link = pk.StatelessDNA(2)
    x = torch.zeros(1, 3, 8, 8, device="cuda")
    with pytest.raises(ValueError, match="Only one mask"):
        link([None] * 6 + [torch.zeros(1, 64, 8, 8, device="cuda")], [None] * 7, 1, x, 2, 3)


def test_model_conv_res_and_summaries_follow_the_reference(pk):
    cfg, params, batch, ref = oracle_step("CDNA", 10, H=32, B=2, T=4)
    m = pk.Model(10, is_cdna=True, scheduled_sampling_k=900.0, prefix="train", height=32, width=32)
    m.load_params(params)
    np.random.seed(99)
    m([torch.from_numpy(a) for a in batch], 6000)
    torch.cuda.synchronize()
    tr = ref["trace"][-1]
    res = m.conv_res                                            # ref:734 conv_res = encs (+ enc7 appended at :715)
    assert len(res) == 8
    want = [e.data for e in tr["encs"]] + [np.maximum(tr["enc7_pre"].data, 0)]
    for got, w in zip(res, want):
        assert tuple(got.shape) == w.shape and rel(got, w) < 1e-4
    lines = m.make_summaries()
    T, ctx = 4, 2
    names = [ln.split(":")[0] for ln in lines]
    want_names = []
    for i in range(T - ctx):
        want_names += ["train_recon_cost%d" % i, "train_psnr%d" % i]
    want_names += ["train_state_cost%d" % i for i in range(T - ctx)] + ["train_psnr_all", "train_loss"]
    assert names == want_names                                  # ref:744-759 order
    sc = [float(ln.split(":")[1]) for ln in lines if "_state_cost" in ln]
    import math
    for i, (s, gs) in enumerate(zip(batch[2][ctx:], ref["gen_states"][ctx - 1:])):
        assert abs(sc[i] - 1e-4 * float(np.mean((s - gs.data) ** 2))) <= 1e-4 * sc[i] + 1e-12
    m.reset_state()
    assert m.conv_res == [] and m.summaries == []


def test_npz_checkpoints_round_trip_and_load_a_chainer_layout_file(pk, tmp_path):
    from pivp_b200 import serializers as S
    cfg = OM.Config("CDNA", 10, schedsamp_k=900.0, height=32, width=32)
    params = OM.init_params(cfg, seed=77)
    # a file as chainer.serializers.save_npz writes it: flat npz, key = parameter path, Chainer layouts, NO .npz suffix (ref:1035)
    f_ref = str(tmp_path / "training-5")
    with open(f_ref, "wb") as f:
        np.savez_compressed(f, **params)
    m = pk.Model(10, is_cdna=True, scheduled_sampling_k=900.0, prefix="t", height=32, width=32)
    S.load_npz(f_ref, m)
    got = m.params()
    assert sorted(got) == sorted(params) and all(np.array_equal(got[k], params[k]) for k in params)
    # save -> same keys / shapes / bytes
    f_out = str(tmp_path / "training-6")
    S.save_npz(f_out, m)
    assert os.path.exists(f_out) and not os.path.exists(f_out + ".npz")
    with np.load(f_out) as z:
        assert sorted(z.files) == sorted(params)
        assert all(z[k].shape == params[k].shape and np.array_equal(z[k], params[k]) for k in params)
    # every missing entry is reported at once; a wrong-shaped entry names itself
    broken = {k: v for k, v in params.items() if k not in ("enc0/W", "lstm3/conv/b")}
    f_b = str(tmp_path / "broken")
    with open(f_b, "wb") as f:
        np.savez(f, **broken)
    with pytest.raises(KeyError) as ei:
        S.load_npz(f_b, m)
    assert "enc0/W" in str(ei.value) and "lstm3/conv/b" in str(ei.value)
    other = dict(params); other["masks/W"] = np.zeros((64, 5, 1, 1), np.float32)
    with open(f_b, "wb") as f:
        np.savez(f, **other)
    with pytest.raises(ValueError, match="masks/W"):
        S.load_npz(f_b, m)


def test_training_resumes_from_model_and_optimizer_files(pk, tmp_path):
    """ref:1035-1037 save both files, ref:864-869 load them: 2 steps + save + load + 1 step == 3 uninterrupted steps."""
    from pivp_b200 import serializers as S
    cfg = OM.Config("CDNA", 10, schedsamp_k=900.0, height=32, width=32)
    params = OM.init_params(cfg)
    batch = [torch.from_numpy(a) for a in OM.concat_examples(OM.synthetic_sequences(2, 4, cfg))]

    def fresh():
        m = pk.Model(10, is_cdna=True, scheduled_sampling_k=900.0, prefix="t", height=32, width=32)
        m.load_params(params)
        return m, pk.Adam(alpha=1e-3).setup(m)
    m1, o1 = fresh()
    np.random.seed(4)
    for i in range(3):
        o1.update(m1, batch, 6000 + i)
    m2, o2 = fresh()
    np.random.seed(4)
    for i in range(2):
        o2.update(m2, batch, 6000 + i)
    o2.new_epoch()
    S.save_npz(str(tmp_path / "training-1"), m2)
    S.save_npz(str(tmp_path / "state-1"), o2)
    with np.load(str(tmp_path / "state-1")) as z:              # Chainer's Optimizer.serialize / UpdateRule.serialize key scheme
        assert int(z["t"]) == 2 and int(z["epoch"]) == 1
        assert z["lstm1/conv/W/m"].shape == (128, 64, 5, 5) and z["lstm1/conv/W/v"].shape == (128, 64, 5, 5) and int(z["lstm1/conv/W/t"]) == 2
    rng_state = np.random.get_state()
    m3, o3 = fresh()
    m3.load(str(tmp_path / "training-1"))
    S.load_npz(str(tmp_path / "state-1"), o3)
    assert o3.t == 2 and o3.epoch == 1
    assert torch.equal(o3.m, o2.m) and torch.equal(o3.v, o2.v) and torch.equal(m3.engine.flat_p, m2.engine.flat_p)
    np.random.set_state(rng_state)
    o3.update(m3, batch, 6002)
    torch.cuda.synchronize()
    # fp32 atomics are unordered, so "identical" means: within a few alpha of the uninterrupted run (Adam turns an ulp on a near-zero
    # gradient into +-alpha); a lost m / v / t would move every parameter by ~alpha * sqrt(1000)
    d = (m3.engine.flat_p - m1.engine.flat_p).abs()
    assert float(d.max()) <= 4e-3 and float(d.mean()) < 2e-5
