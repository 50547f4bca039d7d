set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_gpu_tests_final.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 120 python scripts/phase_times.py > gpurun_out/r2_phase_times.txt 2>&1
PIVP_BRANCHES= PIVP_WGRAD_STREAMS=1 timeout 120 python scripts/phase_times.py >> gpurun_out/r2_phase_times.txt 2>&1
python scripts/ncu_step.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python scripts/ncu_step.py > gpurun_out/ncu_launches.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
