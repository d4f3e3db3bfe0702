# round-2 scaling run on one 8-GPU box: the multi-GPU tests, then bench.py at N = 1, 2, 4, 8 exactly as the driver launches it
set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q --timeout=400 > gpurun_out/r02_multi_tests_8gpu.log 2>&1; tail -3 gpurun_out/r02_multi_tests_8gpu.log
timeout 400 python bench.py --gpus 1 --steps 500 --warmup 20 > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err; echo rc=$?
for n in 2 4 8; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 500 --warmup 20 > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err; echo rc=$?
  tail -c 400 gpurun_out/r02_scale_n$n.err
done
