import statistics, sys, os
sys.path.insert(0, os.getcwd())
import torch
import leafline_b200 as ll
eng = ll.Engine(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
root = ll.new_world()
def timed(fn, reps):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), min(ts)
for _ in range(3): eng.perft(root, 6)
for mode, name in ((2, "plain launches"), (0, "graph replay"), (2, "plain launches"), (0, "graph replay")):
    eng.force_synced(mode)
    eng.perft(root, 6)
    p6 = timed(lambda: eng.perft(root, 6), 60)
    print(f"{name:16s} perft6 {p6[0]:.4f} (min {p6[1]:.4f})")
