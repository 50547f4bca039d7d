# Round-2 final evidence, one gpurun call on one B200 (every ncu run follows a plain run of the same command in this call):
#   GPU test suite, bench (ours + reference arm), phase times, per-layer GEMM table, launch list of one eager step, and `ncu --set full`
#   of the kernels that changed after scripts/evidence_r2.sh was last run (the other kernels' captures from that run stay valid).
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_gpu_tests_final.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 120 python scripts/phase_times.py > gpurun_out/r2_phase_times.txt 2>&1
PIVP_BRANCHES= PIVP_WGRAD_STREAMS=1 timeout 120 python scripts/phase_times.py >> gpurun_out/r2_phase_times.txt 2>&1
timeout 200 python scripts/halo_layers.py --json gpurun_out/r2_halo_layers_final.json > gpurun_out/r2_halo_layers_final.md 2>&1
python scripts/ncu_step.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python scripts/ncu_step.py > gpurun_out/ncu_launches.log 2>&1
for k in wgrad5x5_halo:7 conv_taps_tc:12 heads_fwd:2 fwd_fused_kernel:2 conv5x5_wgrad_tc:3; do
  name=${k%%:*}; cnt=${k##*:}; safe=$(echo $name | tr -c 'a-zA-Z0-9_' '_')
  timeout 300 ncu --profile-from-start off --set full --clock-control none -k regex:$name -c $cnt -o gpurun_out/r2_full_$safe python scripts/ncu_step.py > gpurun_out/ncu_full_$safe.log 2>&1
  ncu -i gpurun_out/r2_full_$safe.ncu-rep --page raw --csv > gpurun_out/r2_full_$safe.csv 2>/dev/null
  rm -f gpurun_out/r2_full_$safe.ncu-rep
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
