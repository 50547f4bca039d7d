"""Checkpoint files compatible with the reference's ``chainer.serializers.save_npz / load_npz`` (train_model.py:864-869, 1035-1037).

Chainer 2.0.1 writes one flat ``.npz`` per object: ``DictionarySerializer`` walks the link tree and stores every parameter under
its path without the leading slash -- exactly the names ``models/npz_keys.py`` prints and ``layout.param_specs`` uses
(``enc0/W``, ``lstm3/conv/b``, ``hidden5/norm/gamma``, ``model/cdna_kerns/W`` ...), each in Chainer's own layout (OIHW convolutions,
(in,out,kh,kw) deconvolutions, gate rows j,i,f,o, C*H*W LayerNorm vectors).  ``layout.ParamSpec`` converts to and from the
kernels' private layouts with exact permutations, so a reference checkpoint loads bit for bit and a file written here loads in
the reference.

The optimizer file (``state-<epoch>``, ref:1037) follows ``Optimizer.serialize`` + ``UpdateRule.serialize`` of Chainer 2.0.1:
``t`` and ``epoch`` at the top level and, per parameter path, ``<path>/t`` (the rule's update count) and the AdamRule state
``<path>/m``, ``<path>/v`` in the parameter's Chainer layout.  Files without the per-parameter ``t`` are accepted on load.

Like ``chainer.serializers.save_npz`` the file is written to ``filename`` as given (no ``.npz`` suffix is appended; the reference
names its files ``training-<epoch>`` / ``state-<epoch>``).
"""
import numpy as np


def _is_optimizer(obj):
    return hasattr(obj, "m") and hasattr(obj, "v") and hasattr(obj, "target")


def model_state(model):
    """{Chainer path: array in Chainer layout} of every parameter (what DictionarySerializer collects)."""
    return model.params()


def optimizer_state(opt):
    e = opt.target.engine
    t = opt.t
    out = {"t": np.asarray(t, np.int64), "epoch": np.asarray(getattr(opt, "epoch", 0), np.int64)}
    m, v = e.export_chainer(opt.m), e.export_chainer(opt.v)
    for s in e.specs:
        out[s.name + "/t"] = np.asarray(t, np.int64)
        out[s.name + "/m"] = m[s.name]
        out[s.name + "/v"] = v[s.name]
    return out


def save_npz(filename, obj, compression=True):
    """``chainer.serializers.save_npz(filename, obj)`` for a ``Model`` or an ``Adam`` of this package."""
    state = optimizer_state(obj) if _is_optimizer(obj) else model_state(obj)
    with open(filename, "wb") as f:
        (np.savez_compressed if compression else np.savez)(f, **state)


def _check_keys(npz, wanted, what, strict_extra=False):
    have = set(npz.files)
    missing = sorted(k for k in wanted if k not in have)
    if missing:
        raise KeyError("%s: %d entries missing from the checkpoint: %s" % (what, len(missing), ", ".join(missing)))
    extra = sorted(have - set(wanted))
    if extra and strict_extra:
        raise KeyError("%s: unexpected entries in the checkpoint: %s" % (what, ", ".join(extra)))
    return extra


def load_npz(filename, obj, strict=False):
    """``chainer.serializers.load_npz(filename, obj)``.  Every missing entry is reported in ONE error; entries the object does not
    have (another model type's links, persistent values) are ignored unless ``strict``; shapes are validated."""
    with np.load(filename) as npz:
        if _is_optimizer(obj):
            return _load_optimizer(npz, obj, strict)
        e = obj.engine
        _check_keys(npz, [s.name for s in e.specs], "load_npz(model)", strict)
        params = {}
        bad = []
        for s in e.specs:
            a = np.asarray(npz[s.name])
            if tuple(a.shape) != s.chainer_shape:
                bad.append("%s: file %s, model %s" % (s.name, tuple(a.shape), s.chainer_shape))
            params[s.name] = a
        if bad:
            raise ValueError("load_npz(model): shape mismatch (other model type / image size / num_masks?): " + "; ".join(bad))
        obj.load_params(params)


def _load_optimizer(npz, opt, strict):
    import torch
    e = opt.target.engine
    wanted = ["t"] + [s.name + k for s in e.specs for k in ("/m", "/v")]
    _check_keys(npz, wanted, "load_npz(optimizer)", False)
    host_m, host_v = np.zeros(e.nparam, np.float32), np.zeros(e.nparam, np.float32)
    for s in e.specs:
        for host, k in ((host_m, "/m"), (host_v, "/v")):
            a = np.asarray(npz[s.name + k])
            if tuple(a.shape) != s.chainer_shape:
                raise ValueError("load_npz(optimizer): %s has shape %s, expected %s" % (s.name + k, tuple(a.shape), s.chainer_shape))
            host[s.offset:s.offset + s.size] = s.to_internal(a).reshape(-1)
    opt.m.copy_(torch.from_numpy(host_m))
    opt.v.copy_(torch.from_numpy(host_v))
    opt.step.fill_(int(npz["t"]))
    if "epoch" in npz.files:
        opt.epoch = int(npz["epoch"])
