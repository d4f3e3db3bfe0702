#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Leafline leaf layer.

Workload (BASELINE.json configs[1]): perft depth 6 from the opening WorldState -- 119 060 324 leaves
of the legal-move tree, i.e. legal move generation over the whole frontier.  One "step" is one full
perft(6).  `value` is leaf positions per second with the root resident on the device; `e2e` is the
same metric through the host-buffer C-ABI call (ll_perft: host root in, host counts out).

With --gpus N > 1 (launched under torchrun) the 2-ply prefixes of the tree are dealt round-robin to
the ranks (no positions are exchanged); the only collective is a u64 all-reduce of the per-root leaf
counts over NCCL.  --impl reference times the CPU oracle (the reference algorithm restated in C; the
Rust reference cannot be built in this image) on the host cores instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PERFT6_OPENING = 119060324
PERFT5_OPENING = 4865609
METRIC = "perft_leaf_positions_per_sec"
UNIT = "leaf positions/s"
WORKLOAD = "perft(6) from the opening WorldState (BASELINE.json configs[1]); 119060324 leaves"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu_index)],
                    capture_output=True, text=True, timeout=5).stdout.strip().split("\n")[0]
                parts = [p.strip() for p in out.split(",")]
                self.samples.append(float(parts[0]))
                self.sm_max = float(parts[1])
                for name, val in zip(names, parts[2:]):
                    if val.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=10)
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.sm_max,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def reference_arm(args, rank, world):
    """The reference's own algorithm on the host cores (oracle port; see module docstring)."""
    if rank != 0:
        return
    import oracle as O

    O.build()
    cores = os.cpu_count() or 1
    w = O.new_world()
    depth, leaves = 5, PERFT5_OPENING  # bounded sample of the workload: one ply shallower
    for _ in range(args.warmup):
        O.perft_divide(w, depth, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        got = int(O.perft_divide(w, depth, threads=cores).sum())
        assert got == leaves
    dt = time.perf_counter() - t0
    value = leaves * args.steps / dt
    sample = f"perft(5) from the opening ({leaves} leaves) per step, one thread per root move over {cores} host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--depth", type=int, default=6)
    ap.add_argument("--no-extras", action="store_true", help="skip the cpu baseline and the secondary kernels")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry

    entry.build()
    import leafline_b200 as ll

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: leafline_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = ll.Engine(local_rank)
    # a real (non-default) stream: the engine launches on it, so the CUDA events below see its kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)

    root = ll.new_world()
    depth = args.depth
    expected = {6: PERFT6_OPENING, 5: PERFT5_OPENING}.get(depth)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    counts = torch.zeros(1024, dtype=torch.int64, device="cuda")

    def step():
        """one perft over this rank's share of the tree + the count all-reduce"""
        leaves, per_root = eng.perft(root, depth, shard=(rank, world))
        if world > 1:
            counts.zero_()
            counts[: len(per_root)] = torch.from_numpy(per_root.astype(np.int64)).cuda(non_blocking=True)
            dist.all_reduce(counts)  # u64 sums as two's-complement i64
            return int(counts.sum().item())
        return leaves

    for _ in range(args.warmup):
        total = step()
        assert expected is None or total == expected, total
    launches0 = eng.launch_count

    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    device_ms = 0.0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)  # flush L2 between timed iterations (not timed)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        total = step()
        b.record(stream)
        b.synchronize()
        device_ms += a.elapsed_time(b)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    launches = eng.launch_count - launches0
    assert expected is None or total == expected, total
    t = torch.tensor([device_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    device_ms = float(t.item())
    leaves_total = total
    ms_per_step = device_ms / args.steps
    value = leaves_total / (ms_per_step * 1e-3)

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD if depth == 6 else f"perft({depth}) from the opening WorldState",
                   "depth": depth, "leaves": leaves_total,
                   "sharding": f"2-ply prefixes round-robin over {world} rank(s); u64 all-reduce of per-root counts",
                   "l2": "flushed between timed steps (256 MiB write, untimed)"},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 104, "d2h_bytes_per_step": 8 + 8 * 1024,
                "note": "every step goes through the host-buffer C-ABI call ll_perft (host root in, host counts out)"},
        "wall_ms_per_step_incl_flush": 1e3 * wall / args.steps,
    }

    if rank == 0 and not args.no_extras:
        peaks, peak_src = measured_peaks()
        # per-kernel device time of the dominant kernel, measured live (CUDA events on the launching stream)
        eng.timing(True)
        eng.timing_reset()
        reps = 3
        for _ in range(reps):
            eng.perft(root, depth, shard=(rank, world))
        kernels = {}
        for name in ("count_leaves", "count_legal", "emit_legal", "scan_block_sums", "aos_to_soa", "validate", "shard_filter"):
            ms, n = eng.timing_query(name)
            if n:
                kernels[name] = {"ms_per_step": ms / reps, "launches_per_step": n / reps}
        eng.timing(False)
        eng.timing_reset()
        result["kernels"] = kernels
        if "count_leaves" in kernels and world == 1:
            frontier = PERFT5_OPENING if depth == 6 else None
            if frontier:
                k_ms = kernels["count_leaves"]["ms_per_step"]
                alg_bytes = frontier * 104.0  # 100 B position + 4 B root id per frontier node
                achieved = alg_bytes / (k_ms * 1e-3) / 1e9
                result["roofline"] = {
                    "kernel": "k_count_leaves", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                    "note": "streams the depth-5 frontier (4865609 positions x 104 B); integer-issue bound, see profiles/",
                }
        # secondary hot-path kernel: batched static scoring of 2^24 seeded playout positions (configs[2])
        n = 1 << 24
        soa = eng.playout_soa(n)
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        eng.score_soa(soa, out.data_ptr())
        torch.cuda.synchronize()
        times = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            eng.score_soa(soa, out.data_ptr())
            b.record(stream)
            b.synchronize()
            times.append(a.elapsed_time(b))
        k_ms = statistics.median(times)
        achieved = n * 104.0 / (k_ms * 1e-3) / 1e9
        result["score_batch"] = {
            "positions": n, "ms": k_ms, "positions_per_sec": n / (k_ms * 1e-3),
            "roofline": {"kernel": "k_score", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_position": 104},
            "l2": "input 1.7 GB > 126 MB L2",
        }
        eng.soa_free(soa)
        # CPU baseline: the reference algorithm (oracle port) on this box's host cores
        import oracle as O

        cores = os.cpu_count() or 1
        w = O.new_world()
        t0 = time.perf_counter()
        got = int(O.perft_divide(w, 5, threads=cores).sum())
        dt = time.perf_counter() - t0
        assert got == PERFT5_OPENING
        result["cpu_baseline"] = {
            "value": PERFT5_OPENING / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"perft(5) from the opening ({PERFT5_OPENING} leaves), one thread per root move over {cores} host threads, {dt:.2f} s",
        }

    if rank == 0:
        print(json.dumps(result))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
