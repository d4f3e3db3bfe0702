# ncu captures of round 1 (run under gpurun on one B200); outputs land in gpurun_out/
set -x
python profiles/workload.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python profiles/workload.py > gpurun_out/ncu_launch.log 2>&1
python profiles/workload.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"k_expand_recordsILi1|k_horizon|k_score" -o gpurun_out/r01_hot python profiles/workload.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
# the narrow-level kernel (one warp per parent): the four plies of one perft(6) chain
python profiles/workload.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"k_expand_coop" -s 4 -c 4 -o gpurun_out/r01_coop python profiles/workload.py > gpurun_out/ncu_coop.log 2>&1
tail -2 gpurun_out/ncu_coop.log
# launch list of the bench command itself (formatted by tools/bench_launches.py)
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
