# ncu captures of round 2 (run under gpurun on one B200); outputs land in gpurun_out/
set -x
python profiles/workload.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python profiles/workload.py > gpurun_out/ncu_launch.log 2>&1
python profiles/workload.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"k_expand_recordsILi1|k_horizon|k_score" -o gpurun_out/r02_hot python profiles/workload.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
# instructions per pipe of the same kernels (the ALU pipe is the perft tail's bound)
python profiles/workload.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_xu.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum \
    --clock-control none --kernel-name-base mangled -k regex:"k_expand_recordsILi1|k_horizon|k_score" --csv --log-file gpurun_out/r02_pipes.csv python profiles/workload.py > gpurun_out/ncu_pipes.log 2>&1
# launch list of the bench command itself (formatted by tools/bench_launches.py)
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
# one rank's share of a strong-scaling step (perft(7) of the opening dealt over eight ranks)
python profiles/tools/shard_workload.py > gpurun_out/plain_shard.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_shard_launches.csv python profiles/tools/shard_workload.py > gpurun_out/ncu_shard.log 2>&1
tail -2 gpurun_out/ncu_shard.log
