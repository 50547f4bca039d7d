set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_gpu_tests_final.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python scripts/ncu_step.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python scripts/ncu_step.py > gpurun_out/ncu_launches.log 2>&1
for k in conv5x5_halo:14 wgrad5x5_halo:7 bwd_apply_kernel:6 bwd_stats_kernel:4 "lnv::apply_kernel:4" heads_bwd:2 heads_fwd:2 conv_taps_tc:12 grad_handover:5 cdna_band:6 conv5x5_wgrad_tc:3; do
  name=${k%%:*}; cnt=${k##*:}; safe=$(echo $name | tr -c 'a-zA-Z0-9_' '_')
  ncu --profile-from-start off --set full --clock-control none -k regex:$name -c $cnt -o gpurun_out/r2_full_$safe python scripts/ncu_step.py > gpurun_out/ncu_full_$safe.log 2>&1
  ncu -i gpurun_out/r2_full_$safe.ncu-rep --page raw --csv > gpurun_out/r2_full_$safe.csv 2>/dev/null
  rm -f gpurun_out/r2_full_$safe.ncu-rep
done
ls -la gpurun_out | tail -30
