"""Host alpha-beta driver over device frontier batches: seconds, host nodes, frontier nodes and bank hits per
(depth, device_plies) from the opening and from Kiwipete.  Run under gpurun; not a bench value."""
import json
import sys
import time

sys.path.insert(0, ".")
import leafline_b200 as ll

KIWIPETE = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -"
rows = []
with ll.Engine(0) as eng:
    eng.alpha_beta_forecast(ll.new_world(), 5, device_plies=3)  # warm the level buffers
    for name, w, cases in (("opening", ll.new_world(), ((7, 3), (7, 4), (8, 2), (8, 3), (8, 4), (9, 3), (9, 4))),
                           ("kiwipete", ll.reconstruct(KIWIPETE), ((6, 3), (6, 4), (7, 4)))):
        for depth, plies, threads in [(d, p, t) for d, p in cases for t in (1, 32)]:
            eng.set_search_threads(threads)
            t0 = time.perf_counter()
            fc, expanded, handed = eng.alpha_beta_forecast(w, depth, device_plies=plies)
            rows.append({"root": name, "depth": depth, "device_plies": plies, "searcher_threads": threads, "seconds": round(time.perf_counter() - t0, 4),
                         "best": ll.pagan_variation(fc[0]), "score": float(fc[0]["score"]), **eng.last_search_stats})
            print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open("gpurun_out/alphabeta_probe.json", "w"), indent=1)
