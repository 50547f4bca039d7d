#!/usr/bin/env python
"""Benchmark of the hot path: CDNA 64x64 training step (forward + BPTT + gradient all-reduce + Adam).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one rank per GPU (torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (NumPy restatement), rank 0 only

Prints ONE JSON line (rank 0).  Metric: train frames/sec = global_batch * T / step_time (BASELINE.json).
 value        : inputs resident in HBM, whole step replayed as a CUDA graph, CUDA-event time, max over ranks
 e2e          : same step through the public API (TrainStep.load_batch from pinned host memory + step + loss D2H read)
 roofline     : dominant kernel = the tcgen05 ConvLSTM implicit GEMM; algorithmic FLOPs / CUDA-event duration, measured
                live on the launching stream during one instrumented step
 cpu_baseline : the oracle (NumPy restatement of the reference's Chainer CPU path) timed on this box's host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train frames/sec (CDNA, 64x64, b32)"
T_SEQ, H, W, MASKS, K_SCHED, ITER0 = 10, 64, 64, 10, 900.0, 6000


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def lstm_flops(B):
    """Algorithmic FLOPs of one launch of the ConvLSTM GEMM per layer (2*M*N*K, K = 25*(Cin+C), un-padded)."""
    from importlib import import_module
    lay = import_module("physical-interaction-video-prediction_b200.layout")
    lv = (2, 2, 4, 4, 8, 4, 2)
    out = []
    for cin, c, l in zip(lay.LSTM_IN, lay.LSTM_SIZES, lv):
        M = B * (H // l) * (W // l)
        out.append(2.0 * M * 4 * c * 25 * (cin + c))
    return out


def run_reference(args, rank):
    """The reference's own CPU algorithm for this path.  Chainer 2.0.1 / Python 2 cannot run in this image, so this is the
    NumPy restatement (oracle/), all host threads the BLAS will use, on a bounded sample: ONE sequence of the b32 workload."""
    if rank != 0:
        return
    import numpy as np
    from oracle import model as OM
    cfg = OM.Config("CDNA", MASKS, schedsamp_k=K_SCHED, height=H, width=W)
    params = OM.init_params(cfg)
    adam = OM.Adam()
    batch = OM.concat_examples(OM.synthetic_sequences(1, T_SEQ, cfg))
    np.random.seed(99)
    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    for i in range(warm):
        OM.train_step(params, adam, batch, ITER0 + i, cfg)
    t0 = time.perf_counter()
    for i in range(steps):
        OM.train_step(params, adam, batch, ITER0 + warm + i, cfg)
    dt = (time.perf_counter() - t0) / steps
    cores = len(os.sched_getaffinity(0))
    val = 1 * T_SEQ / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "CDNA 64x64 T=10 10 masks train step (fwd+bwd+Adam), CPU NumPy restatement of the Chainer path",
                       "sample": "1 sequence of the b32 batch per step"},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "CDNA b1 T=10 fwd+bwd+Adam, %d steps after %d warm-up" % (steps, warm)},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline():
    import numpy as np
    from oracle import model as OM
    cfg = OM.Config("CDNA", MASKS, schedsamp_k=K_SCHED, height=H, width=W)
    params = OM.init_params(cfg)
    batch = OM.concat_examples(OM.synthetic_sequences(1, T_SEQ, cfg))
    np.random.seed(99)
    OM.loss_and_grads(params, batch, ITER0, cfg)
    best = 1e9
    for i in range(3):
        t0 = time.perf_counter()
        OM.loss_and_grads(params, batch, ITER0, cfg)
        best = min(best, time.perf_counter() - t0)
    return {"value": T_SEQ / best, "unit": "frames/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
            "sample": "oracle (NumPy restatement of the Chainer CPU path; Chainer 2.0.1 not installable): CDNA b1 T=10 fwd+bwd, best of 3 after 1 warm-up"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (weak scaling)")
    ap.add_argument("--compute", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import numpy as np
    import torch
    import __graft_entry__
    __graft_entry__.build()
    import pivp_b200 as pk
    from pivp_b200 import parallel
    rank, local, world = parallel.init_distributed()
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    W_ = max(3, args.warmup)
    B, T = args.batch, T_SEQ

    model = pk.Model(MASKS, is_cdna=True, scheduled_sampling_k=K_SCHED, prefix="train", height=H, width=W, device=str(dev),
                     compute=args.compute, rank=rank, world_size=world)
    opt = pk.Adam(alpha=0.001).setup(model)
    seqs = pk.data.synthetic_sequences(B, T, H, W, seed=1234 + rank)
    host = [torch.from_numpy(a).pin_memory() for a in pk.concat_examples(seqs)]
    step = pk.TrainStep(model, opt, B, T, graph=not args.no_graph)
    step.load_batch(*host)
    np.random.seed(99)                      # every rank draws the same global permutations
    it = ITER0

    def barrier():
        if world > 1:
            torch.distributed.barrier()

    for _ in range(W_):
        step(it); it += 1
    torch.cuda.synchronize()
    # ---------------- timed region: inputs resident, device time, max over ranks
    sampler = ClockSampler(local)
    barrier(); torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = pk.lib().query("pivp_launch_count")
    ev0.record()
    for _ in range(args.steps):
        step(it); it += 1
    ev1.record()
    torch.cuda.synchronize(); barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = (step.launches_per_step or 0) * args.steps if step.use_graph else pk.lib().query("pivp_launch_count") - n0
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
    ms = float(tmax.item())
    ms_per_step = ms / args.steps
    value = B * world * T / (ms_per_step * 1e-3)
    loss_now = float(model.loss)

    # ---------------- end to end: H2D of the batch from pinned memory + step + D2H loss read, every step
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step.load_batch(*host)
        loss = float(step(it)); it += 1
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    tmax = torch.tensor([e2e_s], device=dev)
    if world > 1:
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
    e2e_val = B * world * T / (float(tmax.item()) / args.steps)
    h2d = sum(a.numel() * 4 for a in host) + (T - 1) * B * 4
    d2h = 2 * (T - 1) * 4

    # ---------------- roofline: instrument ONE eager step with CUDA events around the dominant kernel's launches
    roof, cdna_op = None, None
    if True:                                  # every rank runs the instrumented step (it contains the all-reduce); rank 0 reports
        L = pk.lib()
        orig = L.call
        recs = []

        def timed_call(name, *a):
            if name in ("pivp_tc_conv5x5", "pivp_cdna_fused_fwd", "pivp_cdna_fused_bwd"):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); orig(name, *a); e.record()
                recs.append((name, a, s, e))
            else:
                orig(name, *a)
        L.call = timed_call
        eager = pk.TrainStep(model, opt, B, T, graph=False)
        eager.images, eager.actions, eager.states = step.images, step.actions, step.states
        eager(it); it += 1
        torch.cuda.synchronize()
        L.call = orig
        pk_, src = peaks()
        fl = lstm_flops(B)
        tcp = model.engine.tc
        by_ptr = {}
        if tcp is not None:
            for li, f in enumerate(fl):
                by_ptr[tcp.Wf[li].data_ptr()] = f          # forward launch of layer li
                by_ptr[tcp.Wd[li].data_ptr()] = f          # input-gradient launch: same 2*M*N*K
        tc_t, tc_f, n_tc = 0.0, 0.0, 0
        fw_t = fw_n = bw_t = bw_n = 0
        for name, a, s, e in recs:
            dt = s.elapsed_time(e) * 1e-3
            if name == "pivp_tc_conv5x5":
                tc_f += by_ptr[a[6]]; tc_t += dt; n_tc += 1
            elif name == "pivp_cdna_fused_fwd":
                fw_t += dt; fw_n += 1
            else:
                bw_t += dt; bw_n += 1
        if n_tc:
            ach = tc_f / tc_t / 1e12
            roof = {"bound": "tensor", "kernel": "conv5x5_tc_kernel (tcgen05 ConvLSTM implicit GEMM, fwd+dgrad launches)",
                    "achieved": ach, "peak": pk_["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk_["bf16_tflops_sustained"],
                    "traffic": None, "peak_source": src + " (sustained: kernel timed inside a long step)", "launches": n_tc,
                    "avg_launch_us": tc_t / n_tc * 1e6}
        if fw_n:
            fb, bb = 328680.0 * B, 608208.0 * B
            cdna_op = {"bound": "hbm", "unit": "GB/s", "peak": pk_["hbm_gbs"],
                       "fwd": {"achieved": fb / (fw_t / fw_n) / 1e9, "avg_launch_us": fw_t / fw_n * 1e6, "bytes": fb},
                       "bwd": {"achieved": bb / (bw_t / bw_n) / 1e9, "avg_launch_us": bw_t / bw_n * 1e6, "bytes": bb,
                               "note": "3 kernels + memset per call; v1 writes and re-reads the mu*dmu planes"}}
            cdna_op["fwd"]["frac"] = cdna_op["fwd"]["achieved"] / pk_["hbm_gbs"]
            cdna_op["bwd"]["frac"] = cdna_op["bwd"]["achieved"] / pk_["hbm_gbs"]

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": W_,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.compute == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": "CDNA 64x64 RGB, batch %d per GPU, T=10, 10 masks, scheduled sampling k=900 at iter %d; "
                                       "one step = forward + BPTT + grad all-reduce + Adam" % (B, ITER0),
                           "global_batch": B * world, "seq_len": T, "parallelism": "dp%d" % world,
                           "l2": "per-step working set ~%.1f GB of activations >> 126 MB L2 (no flush needed)" % (2.2 * B / 32),
                           "cuda_graph": bool(step.use_graph), "lstm_gemm": "tcgen05 bf16 fwd(+fused gates)/dgrad/wgrad, deconv fwd on tcgen05; remaining convs SIMT fp32" if args.compute == "bf16" else "SIMT fp32"},
                "clocks": clocks, "gpu_launches": int(launches), "loss": loss_now,
                "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}}
        if roof:
            line["roofline"] = roof
        if cdna_op:
            line["cdna_op"] = cdna_op
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
