# multi-GPU runs of round 1 (gpurun --gpus 8): weak-scaling headline of bench.py at 8, 4, 2 and 1 ranks of ONE box,
# then configs[4] (Kiwipete perft(7) over 8 GPUs, hierarchically verified)
set -x
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
tail -c 700 gpurun_out/scale_n$n.json
done
python bench.py --gpus 1 --no-extras > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 profiles/config5_perft7.py > gpurun_out/config5_n8.json 2> gpurun_out/config5_n8.err
tail -c 300 gpurun_out/config5_n8.json
