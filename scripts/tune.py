"""Sweep the tuning switches of DESIGN.md section 3.1 on the training step (one `bench.py` process per setting, CUDA-graph step time).

    python scripts/tune.py                      # every switch, one at a time, against the defaults
    python scripts/tune.py PIVP_TC_PAIR_BN=64,96,192 PIVP_LN_APPLY_GX=8,16,32

With programmatic dependent launch on (the default) the step time is stable to ~0.01 ms between processes, so a 0.03 ms difference is
real; with PIVP_PDL=0 the plain-launch graph alternates between two modes 0.3 ms apart (scripts/dbg_bimodal.py) and single runs mislead.
"""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_SWEEP = {
    "PIVP_PDL": ["0"],
    "PIVP_TC_HALO": ["0", "3", "4", "5"],
    "PIVP_TC_HALO_NP": ["1", "2"],
    "PIVP_TC_HALO_SPLITK": ["0", "2", "8"],
    "PIVP_TC_SPLIT_N": ["1"],
    "PIVP_TC_PAIR_BN": ["32", "64", "192"],
    "PIVP_TC_TAPS_RING_KB": ["100", "190"],
    "PIVP_TC_WGRAD_HALO": ["0"],
    "PIVP_CDNA_P": ["4"],
    "PIVP_CDNA_PB": ["2"],
    "PIVP_LN_BWD_CTAS": ["592", "2368"],
    "PIVP_LN_APPLY_GX": ["8", "32", "64"],
}


def step_ms(env_extra, steps=20):
    env = dict(os.environ)
    env.update(env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(steps), "--warmup", "3", "--no-cpu-baseline"],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.strip().splitlines()
    try:
        return json.loads(out[-1])["ms_per_step"]
    except (IndexError, ValueError, KeyError):
        return None                              # the setting is not valid for this configuration (the library fails loudly)


def main():
    sweep = DEFAULT_SWEEP
    if len(sys.argv) > 1:
        sweep = {}
        for a in sys.argv[1:]:
            k, v = a.split("=", 1)
            sweep[k] = v.split(",")
    base = step_ms({})
    print("defaults: %.3f ms/step" % base)
    for k, vals in sweep.items():
        for v in vals:
            ms = step_ms({k: v})
            print("  %-24s %10s" % ("%s=%s" % (k, v), "failed" if ms is None else "%.3f ms (%+.3f)" % (ms, ms - base)), flush=True)
    print("defaults again: %.3f ms/step" % step_ms({}))


if __name__ == "__main__":
    main()
