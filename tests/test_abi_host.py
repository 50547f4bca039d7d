"""CPU tests: the C-ABI library loads and exports every symbol include/pivp.h declares; parameter layouts are exact
permutations; the scheduled-sampling plan is bit-exact with the oracle and consistent across data-parallel ranks
(world_size 2 over gloo)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import model as OM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pk():
    sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.build()
    import pivp_b200
    return pivp_b200


def test_library_exports_every_declared_symbol(pk):
    protos = pk.parse_header()
    assert len(protos) >= 30
    cdll = ctypes.CDLL(pk.LIBPATH)
    for name in protos:
        assert hasattr(cdll, name), name
    L = pk.lib()
    assert L.query("pivp_abi_version") == 1
    assert L.query("pivp_layernorm_workspace_bytes", 32, 64 * 64 * 64) == 32 * 64 * 8
    assert L.query("pivp_cdna_fused_bwd_workspace_bytes", 32, 64, 64, 10) == 4 * (32 * 11 * 4096 + 32 * 250)


def test_every_entry_point_is_documented():
    """INTEGRATION.md section D names every `pivp_*` prototype of include/pivp.h with the reference lines it replaces."""
    import re
    header = open(os.path.join(ROOT, "include", "pivp.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    names = sorted(set(re.findall(r"\b(pivp_[a-z0-9_]+)\s*\(", header)))
    assert len(names) > 50
    missing = [n for n in names if ("`%s`" % n) not in doc]
    assert not missing, missing


def test_symbols_are_sm100a_only(pk):
    out = subprocess.run(["cuobjdump", "--list-elf", pk.LIBPATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT).stdout.decode()
    archs = set(l.split(".")[-2] for l in out.splitlines() if l.strip().endswith(".cubin"))
    assert archs == {"sm_100a"}, archs


def test_error_contract_without_gpu(pk):
    """Shape / pointer violations are reported as errors (no launch is attempted, so this runs on CPU)."""
    L = pk.lib()
    with pytest.raises(pk.PivpError) as ei:
        L.call("pivp_conv2d_fwd", 0, 1, 0, 1, 8, 8, 3, 0, 0, 4, 3, 3, 1, 1, 0, 4, 0, 8, 8, 0, 0, 0)
    assert "null pointer" in str(ei.value)
    with pytest.raises(pk.PivpError):
        L.call("pivp_lstm_gates_fwd", 1, 0, 1, 1, 48, 0, 0, 0, 0, 10, 48, 1.0, 0)       # C not a multiple of 32


@pytest.mark.parametrize("mt,nm,H", [("CDNA", 10, 64), ("DNA", 1, 64), ("STP", 10, 128), ("CDNA", 3, 16)])
def test_layout_roundtrip_and_init_match_oracle(pk, mt, nm, H):
    lay = pk.layout
    specs, n = lay.param_specs(mt, nm, True, H, H)
    cfg = OM.Config(mt, nm, height=H, width=H)
    assert {s.name: s.chainer_shape for s in specs} == {k: tuple(v) for k, v in OM.param_shapes(cfg).items()}
    init, ref = lay.lecun_normal_init(specs), OM.init_params(cfg)
    rs = np.random.RandomState(0)
    spans = []
    for s in specs:
        assert np.array_equal(init[s.name], ref[s.name]), s.name
        a = rs.standard_normal(s.chainer_shape).astype(np.float32)
        assert np.array_equal(s.to_chainer(s.to_internal(a)), a), s.name
        spans.append((s.offset, s.offset + s.size))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= n
    d = {s.name: s for s in specs}
    assert d["masks/W"].offset == d["model/enc7/W"].offset + d["model/enc7/W"].size
    assert d["masks/b"].offset == d["model/enc7/b"].offset + d["model/enc7/b"].size


def test_gate_interleave_is_a_permutation(pk):
    for C in (32, 64, 128):
        p = pk.layout.gate_perm(C)
        assert sorted(p) == list(range(4 * C))
        # channel ch of gate g lands in the 128-wide block ch//32 at g*32 + ch%32
        assert p[2 * C + 37 % C] == (37 % C // 32) * 128 + 2 * 32 + (37 % C) % 32


def test_schedule_plan_bit_exact_with_oracle(pk):
    from pivp_b200 import parallel
    for it in (0, 3000, 6000, 10000):
        assert int(parallel.num_ground_truth(32, 900.0, it)) == int(OM.num_ground_truth(32, 900.0, it))
    B, T, ctx = 32, 10, 2
    np.random.seed(99)
    feed, take, n_gt = parallel.schedule_plan(B, T, 6000, 900.0, ctx)
    np.random.seed(99)
    ref = [OM.scheduled_sample_order(B, n_gt) for _ in range(T - 1 - ctx)]          # 7 draws for T=10
    assert not feed and int(n_gt) == 17
    assert np.array_equal(take[ctx:].astype(bool), np.array(ref))
    assert not take[:ctx].any()
    assert parallel.schedule_plan(B, T, 0, -1, ctx) == (True, None, None)
    assert parallel.schedule_plan(B, T, 0, 900.0, ctx, train=False) == (True, None, None)


WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from pivp_b200 import parallel
from oracle import model as OM
rank, local, world = parallel.init_distributed("gloo")
B, T, H = 4, 4, 16
cfg = OM.Config("CDNA", 3, schedsamp_k=900.0, height=H, width=H, dtype=np.float64)
params = OM.init_params(cfg)
batch = OM.concat_examples(OM.synthetic_sequences(B, T, cfg))
np.random.seed(21)
feed, take, n_gt = parallel.schedule_plan(B, T, 6000, 900.0, 2, True, rank, world)
rows = parallel.shard_rows(B, rank, world)
# per-rank compute = the oracle on this rank's rows with this rank's slice of the global masks
class Fixed(object):
    def __init__(self, rows_take): self.q = list(rows_take)
    def shuffle(self, idx):
        t = self.q.pop(0); order = np.concatenate([np.flatnonzero(t), np.flatnonzero(t == 0)]); idx[:] = order
shard = tuple(np.ascontiguousarray(a[:, rows]) for a in batch)
cfg_local = cfg
import oracle.model as M2
M2_num = M2.num_ground_truth
M2.num_ground_truth = lambda b, k, it: np.int32(take[2].sum())          # local count of ground-truth rows
out = OM.loss_and_grads(params, shard, 6000, cfg_local, rng=Fixed(take[2:]))
flat = torch.from_numpy(np.concatenate([out["grads"][k].reshape(-1) for k in sorted(out["grads"])]))
parallel.allreduce_sum_(flat)
flat /= world
gathered = [torch.zeros(T - 1, B // world, dtype=torch.int32) for _ in range(world)]
dist.all_gather(gathered, torch.from_numpy(take))
if rank == 0:
    np.save(os.environ["OUT"] + ".grads.npy", flat.numpy())
    np.save(os.environ["OUT"] + ".take.npy", torch.cat(gathered, 1).numpy())
dist.barrier()
'''


def test_data_parallel_world2_gloo(pk, tmp_path):
    """world_size 2 over gloo: ranks slice one global permutation; all-reduced shard gradients == full-batch gradients."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    out = str(tmp_path / "res")
    env = dict(os.environ, OUT=out, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29534", str(script)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0, r.stdout.decode()[-3000:]
    from pivp_b200 import parallel
    B, T, H = 4, 4, 16
    cfg = OM.Config("CDNA", 3, schedsamp_k=900.0, height=H, width=H, dtype=np.float64)
    params = OM.init_params(cfg)
    batch = OM.concat_examples(OM.synthetic_sequences(B, T, cfg))
    np.random.seed(21)
    _, take_full, n_gt = parallel.schedule_plan(B, T, 6000, 900.0, 2)
    assert np.array_equal(np.load(out + ".take.npy"), take_full)
    np.random.seed(21)
    ref = OM.loss_and_grads(params, batch, 6000, cfg)
    flat_ref = np.concatenate([ref["grads"][k].reshape(-1) for k in sorted(ref["grads"])])
    got = np.load(out + ".grads.npy")
    assert np.abs(got - flat_ref).max() <= 1e-9 * np.abs(flat_ref).max()


def test_profile_tooling_reads_the_committed_launch_list():
    """scripts/summarize_launches.py turns the committed ncu launch list into the per-kernel table of profiles/ (tooling must not rot)."""
    src = os.path.join(ROOT, "profiles", "r01_launches_step_session3.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "summarize_launches.py"), src, "3", "title", "cmd"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "conv5x5_halo_tc_kernel" in out.stdout and "| kernel | launches / step |" in out.stdout


def test_concat_examples_matches_reference_form(pk):
    """train_model.py:51-71: np.split on axis 1 + np.rollaxis(img, 3, 1) == the transpose the package does; also vs the oracle."""
    rs = np.random.RandomState(0)
    B, T, H, W = 3, 4, 8, 6
    seqs = [[rs.rand(T, H, W, 3).astype(np.float32), rs.rand(T, 5).astype(np.float32), rs.rand(T, 5).astype(np.float32)] for _ in range(B)]
    img, act, sta = pk.concat_examples(seqs)
    # the reference's literal form
    x = np.array([s[0] for s in seqs])
    ref_img = np.array([np.rollaxis(np.squeeze(a, 1), 3, 1) for a in np.split(x, T, axis=1)])
    ref_act = np.array([np.squeeze(a, 1) for a in np.split(np.array([s[1] for s in seqs]), T, axis=1)])
    assert img.shape == (T, B, 3, H, W) and np.array_equal(img, ref_img) and np.array_equal(act, ref_act)
    o = OM.concat_examples(seqs)
    assert all(np.array_equal(a, b) for a, b in zip((img, act, sta), o))
