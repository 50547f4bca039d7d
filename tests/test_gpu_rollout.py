"""Inference rollout (predict_model.py:99-128) and the training driver (8f-4, train_model.py:772-1048) on the GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import model as OM
from oracle.fused_ops import resize_images

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pk():
    import pivp_b200
    pivp_b200.lib()
    return pivp_b200


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.abs(a.astype(np.float64) - b).max() / (np.abs(b).max() + 1e-30))


@pytest.mark.parametrize("dtype,shape,out", [(np.float32, (2, 3, 48, 80), (16, 24)), (np.uint8, (1, 3, 512, 640), (64, 64)), (np.float32, (1, 3, 64, 64), (64, 64))])
def test_resize_images_matches_the_chainer_restatement_bit_for_bit(pk, dtype, shape, out):
    rs = np.random.RandomState(0)
    x = (rs.rand(*shape) * 255).astype(dtype)
    want = resize_images(x.astype(np.float32), out) / np.float32(255.0)                 # predict_model.py:120-121
    xt = torch.from_numpy(x).cuda()
    y = torch.empty(shape[0], shape[1], out[0], out[1], device="cuda")
    pk.lib().call("pivp_resize_images", xt.data_ptr(), 1 if dtype == np.uint8 else 0, y.data_ptr(), shape[0] * shape[1], shape[2], shape[3],
                  out[0], out[1], 255.0, 1, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), want.astype(np.float32))


@pytest.mark.parametrize("compute,mt,nm", [("f32", "CDNA", 10), ("bf16", "CDNA", 10), ("bf16", "DNA", 1)])
def test_rollout_batch_of_one_matches_oracle_test_mode(pk, compute, mt, nm):
    """predict_model.py:126-128: chainer.config.train = False => feedself; batch of ONE sequence (padded to a pair on the bf16 path)."""
    H, T = 64, 6
    cfg = OM.Config(mt, nm, schedsamp_k=900.0, height=H, width=H, dtype=np.float64, train=False)
    params = OM.init_params(cfg)
    img, act, sta = OM.concat_examples(OM.synthetic_sequences(1, T, cfg))
    rs = np.random.RandomState(1)
    raw = (resize_images(img.reshape(T, 3, H, H).astype(np.float32), (96, 128)) * 255).reshape(T, 1, 3, 96, 128).astype(np.float32)
    small = (resize_images(raw.reshape(T, 3, 96, 128), (H, H)) / np.float32(255.0)).reshape(T, 1, 3, H, H)
    ref = OM.forward(params, (small.astype(np.float64), act, sta), 0, cfg)
    assert ref["n_gt"] is None                                   # feedself: no scheduled sampling in test mode
    m = pk.Model(nm, is_cdna=mt == "CDNA", is_dna=mt == "DNA", scheduled_sampling_k=900.0, prefix="predict", height=H, width=H, compute=compute)
    m.load_params(params)
    m.train = False
    r = pk.Rollout(m, 1, T, graph=True)
    for rep in range(2):                                         # second call replays the captured graph
        r.load_raw(raw, act, sta)
        gen = r()
        torch.cuda.synchronize()
        tol = 1e-4 if compute == "f32" else 2e-2
        assert len(gen) == T - 1 and tuple(gen[0].shape) == (1, 3, H, H)
        for t in range(T - 1):
            g = gen[t].double().cpu().numpy()
            w = ref["gen_images"][t].data
            assert np.linalg.norm(g - w) / np.linalg.norm(w) < tol, (rep, t)
        assert abs(float(m.loss) - float(ref["loss"].data)) <= tol * float(ref["loss"].data)
    out = pk.predict(m, raw, act, sta)
    assert out.shape == (T - 1, 1, 3, H, H) and rel(out[-1], ref["gen_images"][-1].data) < (1e-4 if compute == "f32" else 5e-2)
    assert m.train is False


def test_train_loop_runs_checkpoints_and_resumes(pk, tmp_path, caplog):
    """train_model.py:main on a small synthetic data set: epochs, statistics rows, version file, training-/state- npz, resume."""
    from pivp_b200 import train_loop, data as D
    d = str(tmp_path / "push_train")
    D.write_dataset(d, D.synthetic_sequences(10, 4, 64, 64, seed=5))
    out = str(tmp_path / "models")
    np.random.seed(1)
    gl, gp, save_dir = train_loop.train(data_dir=d, output_dir=out, num_iterations=7, batch_size=4, train_val_split=0.8, gpu=0,
                                        save_interval=1, validation_interval=1, schedsamp_k=900.0, compute="bf16")
    # 8 training sequences, batch 4 => an epoch ends every 2 iterations: 3 complete epochs in 7 iterations
    assert len(gl) == 3 and len(gl[0]) == 5 and all(np.isfinite(r).all() for r in gl) and len(gp) == 3
    base = os.path.basename(save_dir).split("-")
    assert len(base) == 4 and base[2] == "CDNA" and base[3] == "4"                      # predict_model.py:92-95 parses this
    files = sorted(os.listdir(save_dir))
    assert "version" in files and "training-0" in files and "state-0" in files and "training-2" in files
    assert "training-global_losses.npy" in files and np.load(os.path.join(save_dir, "training-global_losses.npy")).shape[1] == 5
    with np.load(os.path.join(save_dir, "state-2")) as z:
        assert int(z["t"]) == 6 and int(z["epoch"]) == 3
    # resume from the last checkpoint (ref:864-869)
    gl2, _, _ = train_loop.train(data_dir=d, output_dir=out, num_iterations=2, batch_size=4, train_val_split=0.8, gpu=0, save_interval=100,
                                 pretrained_model=os.path.join(save_dir, "training-2"), pretrained_state=os.path.join(save_dir, "state-2"),
                                 compute="bf16")
    assert len(gl2) == 1 and gl2[0][0] < gl[0][0]                                       # continues from the trained weights: lower loss than epoch 0
