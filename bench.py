#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Leafline leaf layer.

Workload (BASELINE.json configs[1]): perft depth 6 from the opening WorldState -- 119 060 324 leaves
of the legal-move tree, i.e. legal move generation over the whole frontier.  One "step" is one full
perft(6) per GPU.  `value` is leaf positions per second timed on the device (CUDA events on the
launching stream); `e2e` is the same metric by the host's clock around the C-ABI call ll_perft with
HOST buffers (host root in, host counts out, copies inside the timed region).

With --gpus N > 1 (launched under torchrun) the path shards by independent root positions (weak
scaling): the batch holds one root per GPU -- every root is the opening WorldState, so that each
rank's count stays pinned to 119 060 324 -- and the only collective is a u64 all-reduce of the
per-root leaf counts over NCCL.  The strong-scaling figure (one perft(7) tree dealt by 2-ply prefixes)
is reported beside it as `perft7_strong`.

--impl reference times the CPU oracle (the reference algorithm restated in C; the Rust reference
cannot be built in this image) on the host cores instead.
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

PERFT_OPENING = {1: 20, 2: 400, 3: 8902, 4: 197281, 5: 4865609, 6: 119060324, 7: 3195901860}
METRIC = "perft_leaf_positions_per_sec"
UNIT = "leaf positions/s"
WORKLOAD = "perft(6) from the opening WorldState (BASELINE.json configs[1]); 119060324 leaves"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def profile_constants():
    """Per-launch instruction / DRAM byte counts taken from the committed ncu captures (profiles/constants.json)."""
    path = os.path.join(ROOT, "profiles", "constants.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


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
                self.samples.append((time.perf_counter(), float(parts[0])))
                self.sm_max = float(parts[1])
                for name, val in zip(names, parts[2:]):
                    if val.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self, t_begin=None, t_end=None):
        """median SM clock of the samples taken inside [t_begin, t_end] (the timed region); a region shorter than
        a few nvidia-smi calls falls back to every sample since the sampler started (warm-up included: same load)"""
        self._halt.set()
        self.join(timeout=10)
        inside = [mhz for t, mhz in self.samples if t_begin is not None and t_begin <= t <= t_end]
        window = "timed region"
        if len(inside) < 3:
            inside, window = [mhz for _, mhz in self.samples], "warm-up + timed region"
        return {
            "sm_mhz": statistics.median(inside) if inside else None,
            "sm_max_mhz": self.sm_max,
            "reasons": sorted(self.reasons),
            "samples": len(inside),
            "window": window,
        }


def reference_arm(args, rank, world):
    """The reference's own algorithm on the host cores (oracle port; see module docstring)."""
    if rank != 0:
        return
    import oracle as O

    O.build()
    cores = os.cpu_count() or 1
    w = O.new_world()
    depth, leaves = 5, PERFT_OPENING[5]  # bounded sample of the workload: one ply shallower
    for _ in range(args.warmup):
        O.perft_divide(w, depth, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        got = int(O.perft_divide(w, depth, threads=cores).sum())
        assert got == leaves
    dt = time.perf_counter() - t0
    value = leaves * args.steps / dt
    sample = (f"perft(5) from the opening ({leaves} leaves) per step, one thread per root move "
              f"(the reference's own model, mind.rs:399-414) over {cores} host threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def timed(stream, fn, reps, torch):
    """median CUDA-event time (ms) of fn() on `stream`"""
    times = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        b.synchronize()
        times.append(a.elapsed_time(b))
    return statistics.median(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 2000, about 1.2 s; 40 for --impl reference)")
    ap.add_argument("--warmup", type=int, default=None, help="warm-up steps (default 50; 3 for --impl reference)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--depth", type=int, default=6)
    ap.add_argument("--no-extras", action="store_true", help="skip the cpu baseline and the secondary kernels")
    args = ap.parse_args()
    # a reference step is ~0.5 s of all-core CPU work, a B200 step ~0.5 ms: different defaults, explicit values are honoured
    if args.steps is None:
        args.steps = 2000 if args.impl == "b200" else 40
    if args.warmup is None:
        args.warmup = 50 if args.impl == "b200" else 3
    if args.impl == "b200":
        args.warmup = max(args.warmup, 3)

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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: leafline_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    import leafline_b200 as ll
    from leafline_b200 import sharding

    eng = ll.Engine(local_rank)
    # a real (non-default) stream: the engine launches on it, so the CUDA events below see its kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)

    root = ll.new_world()
    depth = args.depth
    expected = PERFT_OPENING.get(depth)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    counts = torch.zeros(1024, dtype=torch.int64, device="cuda")

    def reduce_counts(per_root):
        """u64 all-reduce (sum) of the per-root leaf counts over NCCL (leafline_b200/sharding.py)"""
        return sharding.reduce_leaf_counts(per_root, dist, torch, device="cuda", scratch=counts)[0]

    d_out = torch.zeros(ll.PERFT_DEVICE_WORDS, dtype=torch.int64, device="cuda")
    h_out = torch.zeros(ll.PERFT_DEVICE_WORDS, dtype=torch.int64).pin_memory()

    def step():
        """one perft per GPU (this rank's root of the batch) + the count all-reduce"""
        if world == 1:
            return eng.perft(root, depth)[0]
        # counts stay on the device: launch chain -> NCCL u64 sum -> one D2H read of the reduced table
        eng.perft_device(root, depth, d_out.data_ptr())
        dist.all_reduce(d_out)
        h_out.copy_(d_out, non_blocking=True)
        stream.synchronize()
        if int(h_out[2]) != 0:  # some rank's level outgrew its buffer: redo through the re-sizing host path
            return reduce_counts(eng.perft(root, depth)[1])
        return int(h_out[0])

    eng.perft(root, depth)  # the first call sizes the level buffers (level-by-level path); the steps run the launch chain
    sampler = ClockSampler(local_rank)
    if rank == 0:  # one nvidia-smi poller per job: the clocks reported are rank 0's GPU
        sampler.start()
    for _ in range(args.warmup):
        total = step()
        assert expected is None or total == expected * world, total
    launches0 = eng.launch_count

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_begin = time.perf_counter()
    device_ms = 0.0
    host_s = 0.0
    for _ in range(args.steps):
        flush.fill_(1)  # flush L2 between timed iterations (not timed)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream.synchronize()
        t0 = time.perf_counter()
        a.record(stream)
        total = step()
        b.record(stream)
        b.synchronize()
        host_s += time.perf_counter() - t0
        device_ms += a.elapsed_time(b)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    launches = eng.launch_count - launches0
    assert expected is None or total == expected * world, total
    t = torch.tensor([device_ms, host_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    device_ms, host_ms = float(t[0].item()), float(t[1].item())
    leaves_total = total
    ms_per_step = device_ms / args.steps
    value = leaves_total / (ms_per_step * 1e-3)
    e2e_value = leaves_total / (host_ms / args.steps * 1e-3)

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD if depth == 6 else f"perft({depth}) from the opening WorldState",
                   "depth": depth, "leaves_per_step": leaves_total, "roots_per_step": world,
                   "sharding": (f"one root position per GPU ({world} ranks), no position ever crosses NVLink; "
                                "u64 all-reduce of per-root leaf counts") if world > 1 else "single GPU",
                   "l2": "flushed between timed steps (256 MiB write, untimed)"},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 104, "d2h_bytes_per_step": 8 + 8 * 1024,
                "ms_per_step": host_ms / args.steps,
                "note": ("host clock around the C-ABI call ll_perft (host root in, host leaf counts out, both copies inside)" if world == 1 else
                         "host clock around ll_perft_device (host root in) + NCCL u64 all-reduce of the count table + its D2H read")},
    }

    # strong-scaling companion: ONE perft(7) tree dealt to the ranks by 2-ply prefixes + u64 all-reduce
    if not args.no_extras:
        def perft7():
            leaves, per_root = eng.perft(root, 7, shard=(rank, world))
            return reduce_counts(per_root) if world > 1 else leaves

        assert perft7() == PERFT_OPENING[7]
        if world > 1:
            dist.barrier()
        ms7 = timed(stream, perft7, 3, torch)
        t7 = torch.tensor([ms7], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t7, op=dist.ReduceOp.MAX)
        ms7 = float(t7.item())
        result["perft7_strong"] = {"leaves": PERFT_OPENING[7], "ms": ms7, "leaf_positions_per_sec": PERFT_OPENING[7] / (ms7 * 1e-3),
                                   "scaling": "strong", "sharding": f"2-ply prefixes round-robin over {world} rank(s)"}

    # configs[3]: depth-5 search from the opening (TT-free exact negamax below every root move), root moves
    # dealt to the ranks, per-root scores gathered with a max over NCCL
    if not args.no_extras:
        def kickoff5():
            roots, scores, scored = eng.kickoff(root, 5, shard=(rank, world))
            if world > 1:
                scores = sharding.gather_root_scores(scores, dist, torch, device="cuda")
            return scores, scored

        scores5, scored5 = kickoff5()
        if world > 1:
            dist.barrier()
        ms5 = timed(stream, kickoff5, 3, torch)
        t5 = torch.tensor([ms5, float(scored5)], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax = t5.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t5, op=dist.ReduceOp.SUM)
            ms5, scored_all = float(tmax[0].item()), int(t5[1].item())
        else:
            scored_all = scored5
        result["kickoff5"] = {"depth": 5, "root_moves": len(scores5), "best_score": float(np.max(scores5)), "leaves_scored": scored_all,
                              "ms": ms5, "leaf_positions_per_sec": scored_all / (ms5 * 1e-3), "scaling": "strong",
                              "sharding": f"root moves round-robin over {world} rank(s); max-gather of per-root f32 scores"}

        # the same search two plies deeper (3.3 G leaves): large enough for the root-subtree sharding to show
        def kickoff7():
            roots, scores, scored = eng.kickoff(root, 7, shard=(rank, world))
            if world > 1:
                scores = sharding.gather_root_scores(scores, dist, torch, device="cuda")
            return scores, scored

        scores7, scored7 = kickoff7()
        if world > 1:
            dist.barrier()
        ms7k = timed(stream, kickoff7, 3, torch)
        t7k = torch.tensor([ms7k, float(scored7)], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax = t7k.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t7k, op=dist.ReduceOp.SUM)
            ms7k, scored7_all = float(tmax[0].item()), int(t7k[1].item())
        else:
            scored7_all = scored7
        result["kickoff7"] = {"depth": 7, "root_moves": len(scores7), "best_score": float(np.max(scores7)), "leaves_scored": scored7_all,
                              "ms": ms7k, "leaf_positions_per_sec": scored7_all / (ms7k * 1e-3), "scaling": "strong",
                              "sharding": f"root moves round-robin over {world} rank(s); max-gather of per-root f32 scores"}

    if rank == 0 and world == 1 and not args.no_extras:
        peaks, peak_src = measured_peaks()
        consts = profile_constants()
        # per-kernel device time inside one step, measured live (CUDA events on the launching stream)
        eng.timing(True)
        eng.timing_reset()
        reps = 5
        for _ in range(reps):
            eng.perft(root, depth)
        kernels = {}
        for name in ("perft_tail", "count_leaves", "count_legal", "emit_legal", "scan_block_sums", "store_world", "aos_to_soa", "validate", "shard_filter"):
            ms, n = eng.timing_query(name)
            if n:
                kernels[name] = {"ms_per_step": ms / reps, "launches_per_step": n / reps}
        eng.timing(False)
        eng.timing_reset()
        result["kernels"] = kernels
        result["kernels_note"] = ("per-kernel durations come from a separate pass of 5 steps with per-launch CUDA events enabled (ll_timing_enable); "
                                  "the headline loop runs without them")
        int_peak = {"lop3_thread_inst_per_sec": eng.measure_int_peak(0), "lop3_imad_thread_inst_per_sec": eng.measure_int_peak(1)}
        result["int_peak_measured"] = int_peak
        if "perft_tail" in kernels and depth == 6:
            k_ms = kernels["perft_tail"]["ms_per_step"]
            nodes = PERFT_OPENING[4]
            alg_bytes = nodes * 104.0 + 21 * 8  # 100 B position + 4 B root id per expanded node, per-root counts out
            achieved = alg_bytes / (k_ms * 1e-3) / 1e9
            result["roofline"] = {
                "kernel": "k_expand_records<PERFT>", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": consts.get("perft6_tail_dram_bytes"), "peak_source": peak_src,
                "ms": k_ms,
                "note": "the perft tail reads only the ply-4 frontier (197281 positions x 104 B) and builds the 4.87 M ply-5 "
                        "positions and their 119 M leaves on-chip: it is integer-issue bound, see int_roofline",
            }
            inst = consts.get("perft6_tail_thread_inst")
            if inst:
                rate = inst / (k_ms * 1e-3)
                result["int_roofline"] = {
                    "kernel": "k_expand_records<PERFT>", "bound": "int-issue", "achieved": rate / 1e12,
                    "peak": int_peak["lop3_imad_thread_inst_per_sec"] / 1e12,
                    "unit": "T thread-inst/s", "frac": rate / int_peak["lop3_imad_thread_inst_per_sec"],
                    "frac_of_alu_pipe": rate / int_peak["lop3_thread_inst_per_sec"],
                    "thread_inst_per_launch": inst, "leaves_per_launch": expected,
                    "ncu": {"alu_pipe_busy_pct": consts.get("perft6_tail_alu_pipe_pct"), "issue_slots_busy_pct": consts.get("perft6_tail_issue_active_pct"),
                            "lanes_active_of_32": consts.get("perft6_tail_lanes_active_of_32")},
                    "note": "thread-level instructions per launch from the committed ncu capture (profiles/constants.json); "
                            "peak = dependency-free LOP3+IMAD stream measured in this run (issue port), frac_of_alu_pipe = vs LOP3 only",
                }
        # movegen+eval half of the metric: horizon over seeded playout positions resident in HBM (configs[2]/[3] shape)
        n_h = 1 << 22
        soa = eng.playout_soa(n_h)
        vals = torch.empty(n_h, dtype=torch.float32, device="cuda")
        scored = eng.horizon_soa(soa, 1, vals.data_ptr())
        h_ms = timed(stream, lambda: eng.horizon_soa(soa, 1, vals.data_ptr()), 5, torch)
        h_inst = consts.get("horizon_2p20_thread_inst")
        result["horizon"] = {
            "frontier_positions": n_h, "plies": 1, "leaf_positions": scored, "ms": h_ms,
            "leaf_positions_per_sec": scored / (h_ms * 1e-3),
            "int_roofline": None if not h_inst else {
                "kernel": "k_horizon", "bound": "int-issue", "achieved": h_inst * (n_h / (1 << 20)) / (h_ms * 1e-3) / 1e12,
                "peak": int_peak["lop3_imad_thread_inst_per_sec"] / 1e12, "unit": "T thread-inst/s",
                "frac": h_inst * (n_h / (1 << 20)) / (h_ms * 1e-3) / int_peak["lop3_imad_thread_inst_per_sec"],
                "ncu": {"alu_pipe_busy_pct": consts.get("horizon_2p20_alu_pipe_pct"), "issue_slots_busy_pct": consts.get("horizon_2p20_issue_active_pct"),
                        "lanes_active_of_32": consts.get("horizon_2p20_lanes_active_of_32")},
                "note": "thread instructions of the committed ncu capture (2^20 parents) scaled to this run's parents"},
            "what": "reckless movegen of every frontier node + static score of every child + negamax max; leaves never leave the SM",
        }
        eng.soa_free(soa)
        # the reference's search structure: host alpha-beta (MVV-LVA + history, full window per root move) over device
        # frontier batches -- depth 8 from the opening, beyond what a full-width tree fits in HBM
        t0 = time.perf_counter()
        fc8, host_nodes, handed_off = eng.alpha_beta_forecast(root, 8, device_plies=4)
        result["alpha_beta_depth8"] = {"seconds": time.perf_counter() - t0, "best_score": float(fc8[0]["score"]), "host_nodes_expanded": host_nodes, "values_remembered": eng.last_search_stats["values_remembered"],
                                       "frontier_nodes_handed_to_device": handed_off, "device_plies": 4,
                                       "what": "ll_alpha_beta_forecast: exact root scores (alpha-beta prunes work, not values)"}
        # batched static scoring of 2^24 seeded playout positions (configs[2])
        n = 1 << 24
        soa = eng.playout_soa(n)
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        eng.score_soa(soa, out.data_ptr())
        torch.cuda.synchronize()
        k_ms = timed(stream, lambda: eng.score_soa(soa, out.data_ptr()), 10, torch)
        achieved = n * 104.0 / (k_ms * 1e-3) / 1e9
        pinned_in = torch.empty(n * 104, dtype=torch.uint8).pin_memory()
        w_np = pinned_in.numpy().view(ll.WORLD_DTYPE)
        chunk = 1 << 20
        for first in range(0, n, chunk):
            w_np[first:first + chunk] = eng.soa_download(soa, first, chunk)
        eng.soa_free(soa)
        pinned_out = torch.empty(n, dtype=torch.float32).pin_memory()
        eng.score_into(w_np, pinned_out.numpy())
        t0 = time.perf_counter()
        eng.score_into(w_np, pinned_out.numpy())
        e2e_s = time.perf_counter() - t0
        assert torch.equal(pinned_out, out.cpu())
        result["score_batch"] = {
            "positions": n, "ms": k_ms, "positions_per_sec": n / (k_ms * 1e-3),
            "roofline": {"kernel": "k_score", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": consts.get("score_dram_bytes"), "peak_source": peak_src,
                         "algorithmic_bytes_per_position": 104},
            "e2e": {"positions_per_sec": n / e2e_s, "ms": e2e_s * 1e3, "h2d_bytes": n * 104, "d2h_bytes": n * 4,
                    "note": "ll_score_batch with pinned HOST buffers: H2D of 1.74 GB + repack + score + D2H inside the timed region"},
            "l2": "input 1.7 GB > 126 MB L2",
        }
        del pinned_in, pinned_out, w_np
        # CPU baseline: the reference algorithm (oracle port) on this box's host cores
        import oracle as O

        cores = os.cpu_count() or 1
        w = O.new_world()
        t0 = time.perf_counter()
        reps_cpu = 0
        while time.perf_counter() - t0 < 10.0:
            got = int(O.perft_divide(w, 5, threads=cores).sum())
            assert got == PERFT_OPENING[5]
            reps_cpu += 1
        dt = (time.perf_counter() - t0) / reps_cpu
        result["cpu_baseline"] = {
            "value": PERFT_OPENING[5] / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"perft(5) from the opening ({PERFT_OPENING[5]} leaves) x{reps_cpu}, one thread per root move over {cores} host threads, {dt:.2f} s each",
        }
        # beside it: the same on ONE host thread, and the CPU scorer on a 2^20-position prefix
        t0 = time.perf_counter()
        assert int(O.perft_divide(w, 5, threads=1).sum()) == PERFT_OPENING[5]
        dt1 = time.perf_counter() - t0
        prefix = O.playout_batch(ll.PLAYOUT_SEED, 0, 1 << 14)
        prefix = np.tile(prefix, 64)
        t0 = time.perf_counter()
        O.score_batch(prefix)
        dts = time.perf_counter() - t0
        result["cpu_baseline_single_thread"] = {
            "perft_leaf_positions_per_sec": PERFT_OPENING[5] / dt1, "score_positions_per_sec": len(prefix) / dts, "cores": 1, "kind": "port",
            "sample": f"perft(5) from the opening once ({dt1:.1f} s); lo_score over 2^20 positions (2^14 playout positions x64, {dts * 1e3:.0f} ms)",
        }

    if rank == 0:
        print(json.dumps(result))
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
