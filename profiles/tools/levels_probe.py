import statistics, sys, os
sys.path.insert(0, os.getcwd())
import torch
import leafline_b200 as ll
eng = ll.Engine(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
root = ll.new_world()
for _ in range(3): eng.perft(root, 6)
def run(do_flush, reps=30):
    ts=[]
    for _ in range(reps):
        if do_flush: flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); eng.perft(root, 6); b.record(stream); b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)
print("with flush", run(True), "without flush", run(False))
import ctypes as C
lib, ctx = eng._lib, eng._ctx
eng.timing(True)
for do_flush in (True, False):
    eng.timing_reset()
    if do_flush: flush.fill_(1)
    torch.cuda.synchronize()
    eng.perft(root, 6)
    # per-launch list is not exposed; print totals
    for name in ("aos_to_soa","emit_legal","perft_tail"):
        print(do_flush, name, eng.timing_query(name))
