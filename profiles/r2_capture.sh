set -x
(time python -m pytest tests -m gpu -x -q) > gpurun_out/gputests.log 2>&1
(time python bench.py) > gpurun_out/bench_default.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:tqc_loss_group --launch-skip 5 -c 1 -o gpurun_out/prof_r2h_tqc python profiles/prof_tqc.py > gpurun_out/ncu_tqc.log 2>&1
$NCU -k regex:sample_gather_lean --launch-skip 3 -c 1 -o gpurun_out/prof_r2h_lean python profiles/prof_gather.py lean2 > gpurun_out/ncu_lean.log 2>&1
$NCU -k regex:fused_pass --launch-skip 40 -c 1 -o gpurun_out/prof_r2h_fused python bench.py --steps 1 --warmup 1 --no-step-graph --no-cpu-baseline --no-e2e --no-extra --no-updates --no-secondary --no-small --no-parity-check > gpurun_out/ncu_fused.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/gputests.log
ls -la gpurun_out/*.ncu-rep
