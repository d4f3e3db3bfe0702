#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Leafline leaf layer.

Workload (BASELINE.json configs[1]): perft depth 6 from the opening WorldState -- 119 060 324 leaves
of the legal-move tree, i.e. legal move generation over the whole frontier.  One "step" is one full
perft(6) per GPU.  `value` is leaf positions per second timed on the device (CUDA events on the
launching stream); `e2e` is the same metric by the host's clock around the C-ABI call with HOST buffers
(host roots in, host counts out, copies inside the timed region).

With --gpus N > 1 (launched under torchrun, one process per GPU) every rank joins ONE library-owned group
(ll_init_multi_rank: NCCL communicator + the device-resident reduction live behind the C ABI) and a step is
one call of ll_perft_batch_multi on a batch of N roots -- root i counted by rank i, counts reduced on the
devices (weak scaling; every root is the opening WorldState so each count stays pinned to 119 060 324).
The partition of ONE tree is reported beside it: perft(7) (`perft7_strong`, ll_perft_multi), the depth-7
search (`kickoff7`, ll_kickoff_multi) and configs[4], perft(7) of Kiwipete (`config5`), each with the same
tree timed on one GPU in the same run (`strong_efficiency` = t1 / (N * tN)).

--impl reference times the CPU oracle (the reference algorithm restated in C; the Rust reference
cannot be built in this image) on the host cores instead.
"""
import argparse
import hashlib
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
KIWIPETE = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -"
KIWIPETE_PERFT7 = 373424113829  # the reference's rules (orthodox chess: 374 190 009 323), profiles/r01_config5_kiwipete_perft7_8gpu.json
METRIC = "perft_leaf_positions_per_sec"
UNIT = "leaf positions/s"
WORKLOAD = "perft(6) from the opening WorldState (BASELINE.json configs[1]); 119060324 leaves"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def device_code_hash():
    """sha256 over the device code the kernels are built from (what an ncu instruction count is a property of)"""
    h = hashlib.sha256()
    for name in ("ll_device.cuh", "ll_kernels.cuh"):
        with open(os.path.join(ROOT, "leafline_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def profile_constants():
    """Per-launch instruction / DRAM byte counts taken from the committed ncu captures (profiles/constants.json).  They are
    properties of the device code: when the sources no longer hash to what the capture was taken from, the instruction-count
    rooflines are withheld instead of being computed from stale counts."""
    path = os.path.join(ROOT, "profiles", "constants.json")
    if not os.path.exists(path):
        return {}, "profiles/constants.json missing"
    with open(path) as f:
        consts = json.load(f)
    if consts.get("device_code_sha256") != device_code_hash():
        return {}, ("profiles/constants.json was captured from other device code (device_code_sha256 differs): "
                    "instruction-count rooflines withheld until the kernels are re-profiled")
    return consts, None


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
    """The reference's own algorithm on the host cores (oracle port; see module docstring): perft(6) of the opening,
    every step the subtrees of the first m of its 20 root moves -- m = 20 (the whole tree) when the run fits the time
    budget, fewer otherwise (a bounded sample of the same workload, said in `config`).  Threads: below every root move
    one thread per reply (the reference's own model one ply down, mind.rs:399-414), all host threads busy."""
    if rank != 0:
        return
    import oracle as O

    O.build()
    cores = os.cpu_count() or 1
    root = O.new_world()
    children = [c["tree"].copy() for c in O.lookahead(root)]

    def subtree(i):
        return int(O.perft_divide(children[i], 5, threads=cores).sum())

    t0 = time.perf_counter()
    first = subtree(0)
    per_move = time.perf_counter() - t0  # calibration, untimed: one root move's subtree
    budget_s = 200.0
    m = int(max(1, min(len(children), budget_s / ((args.steps + args.warmup) * per_move))))
    want = None
    for _ in range(args.warmup):
        got = sum(subtree(i) for i in range(m))
        assert want is None or got == want
        want = got
    t0 = time.perf_counter()
    for _ in range(args.steps):
        got = sum(subtree(i) for i in range(m))
        assert want is None or got == want
        want = got
    dt = time.perf_counter() - t0
    if m == len(children):
        assert want == PERFT_OPENING[6], want
    assert first > 0
    value = want * args.steps / dt
    sample = (f"perft(6) from the opening: the depth-5 subtrees of the first {m} of its {len(children)} root moves per step "
              f"({want} of {PERFT_OPENING[6]} leaves), one thread per reply below each root move over {cores} host threads; "
              f"oracle built -O3 -march=native -ffp-contract=off")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_sample": sample, "root_moves_per_step": m, "leaves_per_step": want},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def timed(stream, fn, reps, torch, flush=None, align=None):
    """median CUDA-event time (ms) of fn() on `stream` (L2 flushed before every repetition when `flush` is given).
    align (a collective call is being timed): a one-element all-reduce is enqueued on the timing stream right before the start
    event, so that every rank's timed region starts together on the devices -- what is measured is the collective call, not how
    far apart the ranks' host loops happened to be"""
    times = []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if align is not None:
            align()
        a.record(stream)
        fn()
        b.record(stream)
        b.synchronize()
        times.append(a.elapsed_time(b))
    return statistics.median(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 2000, about 1.2 s; 5 for --impl reference)")
    ap.add_argument("--warmup", type=int, default=None, help="warm-up steps (default 50; 1 for --impl reference)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--depth", type=int, default=6)
    ap.add_argument("--no-extras", action="store_true", help="skip the cpu baseline and the secondary kernels")
    args = ap.parse_args()
    # a reference step is ~15 s of all-core CPU work, a B200 step ~0.4 ms: different defaults, explicit values are honoured
    if args.steps is None:
        args.steps = 2000 if args.impl == "b200" else 5
    if args.warmup is None:
        args.warmup = 50 if args.impl == "b200" else 1
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

    group = None
    if world > 1:
        # the library owns the communicator: the id is made on rank 0 and carried to the other processes by torch.distributed
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(ll.MultiEngine.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        group = ll.MultiEngine.for_rank(local_rank, rank, world, bytes(uid.cpu().numpy()))
        eng = group.member(0)
    else:
        eng = ll.Engine(local_rank)
    solo = ll.Engine(local_rank) if world > 1 else eng  # the same trees on ONE GPU, timed in the same run
    # a real (non-default) stream: the engine launches on it, so the CUDA events below see its kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    if solo is not eng:
        solo.set_stream(stream.cuda_stream)

    root = ll.new_world()
    depth = args.depth
    expected = PERFT_OPENING.get(depth)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    batch = np.array([root] * world, dtype=ll.WORLD_DTYPE)
    counts = np.zeros(world, dtype=np.uint64)

    def step():
        """one perft per GPU: ll_perft on one GPU; ll_perft_batch_multi (root i -> rank i, counts reduced on the devices) on N"""
        if world == 1:
            return eng.perft(root, depth)[0]
        return int(group.perft_batch(batch, depth, out=counts).sum())

    step()  # the first call sizes the level buffers (level-by-level path); the steps run the launch chain
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
    expanded_per_tree = sum(PERFT_OPENING[d] for d in range(1, depth)) + 1 if depth <= 7 else None  # nodes of plies 0 .. depth-1

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD if depth == 6 else f"perft({depth}) from the opening WorldState",
                   "depth": depth, "leaves_per_step": leaves_total, "roots_per_step": world,
                   "sharding": (f"ll_perft_batch_multi: one root position per GPU ({world} ranks, one process each, library-owned group), no position "
                                f"ever crosses NVLink; u64 sum of the {world} leaf counts on the devices by {group.reduce_path}") if world > 1 else "single GPU (ll_perft)",
                   "l2": "flushed between timed steps (256 MiB write, untimed)"},
        "clocks": clocks, "gpu_launches": launches,
        "expanded_nodes_per_sec": None if expanded_per_tree is None else expanded_per_tree * world / (ms_per_step * 1e-3),
        "expanded_nodes_note": ("SURVEY.md 8(d) work unit: a node whose legal children are generated (plies 0 .. depth-1; the last ply's children are "
                                "bulk-counted, never built) -- leaf positions/s counts those bulk-counted leaves"),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 104 * world, "d2h_bytes_per_step": (8 + 8 * 1024) if world == 1 else 8 * world,
                "ms_per_step": host_ms / args.steps,
                "note": ("host clock around the C-ABI call ll_perft (host root in, host leaf counts out, both copies inside)" if world == 1 else
                         "host clock around the C-ABI call ll_perft_batch_multi (host roots in, device-side reduction, host counts out)")},
    }
    if world > 1:
        result["reduce_path"] = group.reduce_path

    align_word = torch.zeros(1, dtype=torch.int32, device="cuda")

    def strong(name, fn_multi, fn_solo, reps, units, unit_name):
        """ONE tree on the whole group and, in the same run, on one GPU: ms (max over ranks), rate, efficiency t1 / (N * tN)"""
        want = fn_solo()
        got = fn_multi()
        if world > 1:
            dist.barrier()
        ms_n = timed(stream, fn_multi, reps, torch, flush, align=(lambda: dist.all_reduce(align_word)) if world > 1 else None)
        ms_1 = timed(stream, fn_solo, reps, torch, flush) if world > 1 else ms_n
        tt = torch.tensor([ms_n, ms_1], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_n, ms_1 = float(tt[0].item()), float(tt[1].item())
        out = {"ms": ms_n, unit_name: units / (ms_n * 1e-3), "scaling": "strong", "single_gpu_ms_same_run": ms_1,
               "strong_efficiency": ms_1 / (world * ms_n)}
        return want, got, out

    if not args.no_extras:
        # strong scaling, perft: ONE perft(7) tree; its ply-3 nodes dealt to the ranks (ll_perft_multi)
        want, got, out = strong("perft7", (lambda: group.perft(root, 7)[0]) if world > 1 else (lambda: eng.perft(root, 7)[0]),
                                lambda: solo.perft(root, 7)[0], 5, PERFT_OPENING[7], "leaf_positions_per_sec")
        assert want == got == PERFT_OPENING[7], (want, got)
        out.update({"leaves": PERFT_OPENING[7], "sharding": f"ply-3 nodes dealt over {world} rank(s) by (hash(parent) + canonical index) mod world; u64 sum on the devices"})
        result["perft7_strong"] = out

        # configs[3]: depth-5 search of the opening (TT-free exact negamax below every root move).  4.9 M leaves are ~0.3 ms of
        # work: the library answers such a tree whole on every rank instead of dealing it (ll_multi_set_shard_min_depth)
        def kick(e_or_g, d):
            roots, scores, scored = e_or_g.kickoff(root, d)
            return scores.tobytes(), scored

        for d, reps in ((5, 5), (7, 3)):
            want, got, out = strong(f"kickoff{d}", (lambda: kick(group, d)) if world > 1 else (lambda: kick(eng, d)), lambda: kick(solo, d), reps, 1, "x")
            assert want == got, f"kickoff{d}: group scores differ from one GPU's"
            scored = want[1]
            out.pop("x")
            out.update({"depth": d, "root_moves": len(want[0]) // 4, "best_score": float(np.frombuffer(want[0], dtype=np.float32).max()), "leaves_scored": scored,
                        "leaf_positions_per_sec": scored / (out["ms"] * 1e-3),
                        "sharding": ("answered whole by every rank (below the dealing threshold)" if d < 6 or world == 1 else
                                     f"ll_kickoff_multi: ply-3 nodes dealt round-robin over {world} rank(s); per-parent f32 keys max-reduced on the devices")})
            result[f"kickoff{d}"] = out

        # configs[4]: perft(7) of Kiwipete (ascension, secret service, passing by all live) on the whole group
        kiwi = ll.reconstruct(KIWIPETE)
        want, got, out = strong("config5", (lambda: group.perft(kiwi, 7)) if world > 1 else (lambda: eng.perft(kiwi, 7)), lambda: solo.perft(kiwi, 7), 2,
                                KIWIPETE_PERFT7, "leaf_positions_per_sec")
        assert got[0] == want[0] == KIWIPETE_PERFT7 and list(got[1]) == list(want[1]), (got[0], want[0])
        checks = {"single_gpu_recount": True, "divide_table_equal": True}
        children_sum = 0
        if rank == 0:  # sum of the 48 root children's perft(6), each counted separately on one GPU
            _, kids = solo.lookahead(kiwi)
            children_sum = sum(solo.perft(c["tree"], 6)[0] for c in kids)
            assert children_sum == KIWIPETE_PERFT7
            checks["sum_of_root_children_perft6"] = children_sum
            fixture = os.path.join(ROOT, "tests", "golden", "kiwipete_prefix_perft5.json")
            if os.path.exists(fixture):  # oracle recounts of sampled 2-ply prefixes at sub-depth 5 (committed fixture)
                with open(fixture) as f:
                    fx = json.load(f)
                for s in fx["samples"]:
                    assert solo.perft(ll.reconstruct(s["runes"]), 5)[0] == s["perft5"], s
                checks["oracle_prefix_samples_at_sub_depth_5"] = len(fx["samples"])
        out.update({"position": KIWIPETE, "depth": 7, "leaves": KIWIPETE_PERFT7, "verified": checks,
                    "note": "the reference's rules (orthodox chess counts 374190009323: SURVEY.md 8(a) quirks 1-2)"})
        result["config5"] = out

    if not args.no_extras:
        # the reference's search structure: host alpha-beta (MVV-LVA + history, full window per root move, one searcher thread per
        # root move whose device requests are combined) over device frontier batches -- depth 8 from the opening, beyond what a
        # full-width tree fits in HBM; on a group the root moves are searched concurrently, one subset per GPU
        searcher = group if world > 1 else eng
        searcher.alpha_beta_forecast(root, 6, device_plies=3)
        times = []
        for _ in range(3):  # the first search grows the level buffers (reported apart); a search keeps nothing of the one before
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            fc8, host_nodes, handed_off = searcher.alpha_beta_forecast(root, 8, device_plies=4)
            times.append(time.perf_counter() - t0)
        t8 = torch.tensor([min(times[1:]), times[0]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t8, op=dist.ReduceOp.MAX)
        result["alpha_beta_depth8"] = {"seconds": float(t8[0].item()), "first_search_seconds": float(t8[1].item()), "best_score": float(fc8[0]["score"]), "principal_variation": ll.pagan_variation(fc8[0]),
                                       "host_nodes_expanded": host_nodes, "values_remembered": searcher.last_search_stats["values_remembered"],
                                       "frontier_nodes_handed_to_device": handed_off, "device_plies": 4,
                                       "what": ("ll_alpha_beta_forecast" if world == 1 else f"ll_alpha_beta_forecast_multi over {world} ranks") +
                                               ": exact root scores (alpha-beta prunes work, not values)"}

    if rank == 0 and world == 1 and not args.no_extras:
        peaks, peak_src = measured_peaks()
        consts, consts_problem = profile_constants()
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
            hbm = {"achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "peak_source": peak_src,
                   "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": alg_bytes,
                   "note": "the tail reads only the ply-4 frontier (197281 positions x 104 B): HBM is nowhere near the bound"}
            inst = consts.get("perft6_tail_thread_inst")
            roof = {"kernel": "k_expand_records<PERFT>", "bound": "int-issue", "unit": "T thread-inst/s", "ms": k_ms,
                    "traffic": consts.get("perft6_tail_dram_bytes"), "hbm": hbm, "leaves_per_launch": expected,
                    "expanded_nodes_per_launch": PERFT_OPENING[4] + PERFT_OPENING[5]}
            if inst:
                rate = inst / (k_ms * 1e-3)
                alu_warp_inst = consts.get("perft6_tail_alu_pipe_warp_inst")
                roof.update({
                    "achieved": rate / 1e12, "peak": int_peak["lop3_imad_thread_inst_per_sec"] / 1e12,
                    "frac": rate / int_peak["lop3_imad_thread_inst_per_sec"],
                    "thread_inst_per_launch": inst, "thread_inst_per_expanded_child": inst / PERFT_OPENING[5],
                    # the binding pipe: warp instructions the ALU pipe executed, each occupying it for a full 32-lane warp, against the
                    # measured LOP3-only stream (the ALU pipe's peak)
                    "alu_pipe_frac": None if not alu_warp_inst else alu_warp_inst * 32.0 / (k_ms * 1e-3) / int_peak["lop3_thread_inst_per_sec"],
                    "ncu": {"alu_pipe_busy_pct": consts.get("perft6_tail_alu_pipe_pct"), "issue_slots_busy_pct": consts.get("perft6_tail_issue_active_pct"),
                            "lanes_active_of_32": consts.get("perft6_tail_lanes_active_of_32")},
                    "note": "achieved = thread-level instructions per launch (committed ncu capture of this device code, profiles/constants.json) / the "
                            "kernel's live CUDA-event time; peak = dependency-free LOP3+IMAD stream measured in this run (the issue port)",
                })
            else:
                roof.update({"achieved": None, "peak": int_peak["lop3_imad_thread_inst_per_sec"] / 1e12, "frac": None, "withheld": consts_problem})
            result["roofline"] = roof
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
            "int_roofline": {"withheld": consts_problem} if not h_inst else {
                "kernel": "k_horizon", "bound": "int-issue", "achieved": h_inst * (n_h / (1 << 20)) / (h_ms * 1e-3) / 1e12,
                "peak": int_peak["lop3_imad_thread_inst_per_sec"] / 1e12, "unit": "T thread-inst/s",
                "frac": h_inst * (n_h / (1 << 20)) / (h_ms * 1e-3) / int_peak["lop3_imad_thread_inst_per_sec"],
                "ncu": {"alu_pipe_busy_pct": consts.get("horizon_2p20_alu_pipe_pct"), "issue_slots_busy_pct": consts.get("horizon_2p20_issue_active_pct"),
                        "lanes_active_of_32": consts.get("horizon_2p20_lanes_active_of_32")},
                "note": "thread instructions of the committed ncu capture (2^20 parents) scaled to this run's parents"},
            "what": "reckless movegen of every frontier node + static score of every child + negamax max; leaves never leave the SM",
        }
        eng.soa_free(soa)
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
        # CPU baseline: the reference algorithm (oracle port) on this box's host cores, the SAME workload: perft(6) of the
        # opening once (about 15 s), below every root move one thread per reply
        import oracle as O

        cores = os.cpu_count() or 1
        w = O.new_world()
        kids = [c["tree"].copy() for c in O.lookahead(w)]
        t0 = time.perf_counter()
        got = sum(int(O.perft_divide(k, 5, threads=cores).sum()) for k in kids)
        dt = time.perf_counter() - t0
        assert got == PERFT_OPENING[6]
        result["cpu_baseline"] = {
            "value": PERFT_OPENING[6] / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"perft(6) from the opening once ({PERFT_OPENING[6]} leaves, {dt:.1f} s), one thread per reply below each root move over {cores} host threads; "
                      "oracle built -O3 -march=native -ffp-contract=off",
        }
        # beside it: perft(5) on ONE host thread, and the CPU scorer on a 2^20-position prefix
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
    if solo is not eng:
        solo.close()
    if group is not None:
        group.close()
    else:
        eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
