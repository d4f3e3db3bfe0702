#!/usr/bin/env python
"""per_gpu.py -- perft(6) of the opening on every GPU of the box, one after the other (median ms of 300 L2-flushed steps):
how far apart the GPUs of one box are is the floor under any weak-scaling figure that waits for the slowest."""
import statistics
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import leafline_b200 as ll

root = ll.new_world()
for dev in range(torch.cuda.device_count()):
    torch.cuda.set_device(dev)
    eng = ll.Engine(dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")
    for _ in range(30):
        assert eng.perft(root, 6)[0] == 119060324
    ts = []
    for _ in range(300):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream.synchronize()
        a.record(stream)
        eng.perft(root, 6)
        b.record(stream)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    print("gpu %d: median %.4f ms  p10 %.4f  p90 %.4f" % (dev, statistics.median(ts), sorted(ts)[30], sorted(ts)[270]), flush=True)
    eng.close()
