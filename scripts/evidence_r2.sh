# Round-2 evidence, one gpurun call on one B200 (every ncu run follows a plain run of the same command in this call):
#   GPU test suite, bench (ours + reference arm), launch list of one eager step, `ncu --set full` of the kernels the report cites.
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_gpu_tests_final.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 120 python scripts/phase_times.py > gpurun_out/r2_phase_times.txt 2>&1
timeout 120 python scripts/halo_timeline_all.py > gpurun_out/r2_halo_timeline_final.md 2>&1
timeout 200 python scripts/halo_layers.py --json gpurun_out/r2_halo_layers_final.json > gpurun_out/r2_halo_layers_final.md 2>&1
python scripts/ncu_step.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python scripts/ncu_step.py > gpurun_out/ncu_launches.log 2>&1
for k in conv5x5_halo:126 wgrad5x5_halo:7 bwd_fused_kernel:8 bwd_apply_kernel:3 "lnv::apply_kernel:4" heads_bwd:2 heads_fwd:2 conv_taps_tc:12 grad_handover:5 cdna_band:4 conv5x5_wgrad_tc:3; do
  name=${k%%:*}; cnt=${k##*:}; safe=$(echo $name | tr -c 'a-zA-Z0-9_' '_')
  ncu --profile-from-start off --set full --clock-control none -k regex:$name -c $cnt -o gpurun_out/r2_full_$safe python scripts/ncu_step.py > gpurun_out/ncu_full_$safe.log 2>&1
  ncu -i gpurun_out/r2_full_$safe.ncu-rep --page raw --csv > gpurun_out/r2_full_$safe.csv 2>/dev/null
  rm -f gpurun_out/r2_full_$safe.ncu-rep
done
python scripts/ncu_ops.py 256 > gpurun_out/ncu_ops_plain.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none -o gpurun_out/r2_ops_b256 python scripts/ncu_ops.py 256 > gpurun_out/ncu_ops.log 2>&1
ncu -i gpurun_out/r2_ops_b256.ncu-rep --page raw --csv > gpurun_out/r2_full_ops_b256.csv 2>/dev/null; rm -f gpurun_out/r2_ops_b256.ncu-rep
ls gpurun_out | wc -l
