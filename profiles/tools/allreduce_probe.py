#!/usr/bin/env python
"""Where does the multi-GPU step overhead go?  (run under torchrun)"""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import leafline_b200 as ll
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = ll.Engine(local)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
root = ll.new_world()
eng.perft(root, 6)
d_out = torch.zeros(ll.PERFT_DEVICE_WORDS, dtype=torch.int64, device="cuda")
small = d_out[:64]
h_out = torch.zeros(ll.PERFT_DEVICE_WORDS, dtype=torch.int64).pin_memory()
def timed(fn, reps=200):
    for _ in range(20): fn()
    dist.barrier(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream.synchronize(); t0 = time.perf_counter()
        a.record(stream); fn(); b.record(stream); b.synchronize()
        ts.append(((time.perf_counter() - t0) * 1e3, a.elapsed_time(b)))
    return statistics.median(t[0] for t in ts), statistics.median(t[1] for t in ts)
def only_perft(): eng.perft_device(root, 6, d_out.data_ptr()); stream.synchronize()
def perft_ar(): eng.perft_device(root, 6, d_out.data_ptr()); dist.all_reduce(d_out); stream.synchronize()
def perft_ar_small(): eng.perft_device(root, 6, d_out.data_ptr()); dist.all_reduce(small); stream.synchronize()
def full(): eng.perft_device(root, 6, d_out.data_ptr()); dist.all_reduce(d_out); h_out.copy_(d_out, non_blocking=True); stream.synchronize()
def ar_only(): dist.all_reduce(d_out); stream.synchronize()
for name, fn in (("perft_device only", only_perft), ("+ all_reduce 8 KB", perft_ar), ("+ all_reduce 512 B", perft_ar_small), ("+ D2H (the bench step)", full), ("all_reduce alone", ar_only)):
    host, dev = timed(fn)
    if rank == 0: print(f"{name:26s} host {host:.4f} ms   device {dev:.4f} ms", flush=True)
dist.barrier(); dist.destroy_process_group()
