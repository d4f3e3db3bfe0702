# multi-GPU runs of round 1 (gpurun --gpus 8): weak-scaling headline of bench.py at 8 and 4 ranks
# (N=2 and N=1 were taken on smaller leases; configs[4] = profiles/config5_perft7.py, see r01_config5_*.json)
set -x
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
tail -c 900 gpurun_out/scale_n$n.json
done
