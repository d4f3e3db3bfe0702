set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "bench rc=$?"; tail -c 4000 gpurun_out/bench_r1c.json; tail -5 gpurun_out/bench_r1c.err
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_expand" -s 12 -c 5 -o gpurun_out/prof_expand_r1c python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
