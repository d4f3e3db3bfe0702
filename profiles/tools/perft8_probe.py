import sys, os, time
sys.path.insert(0, os.getcwd())
import leafline_b200 as ll
eng = ll.Engine(0)
w = ll.new_world()
for d, want in ((7, 3195901860), (8, 84998978956)):
    t0 = time.perf_counter(); leaves, per_root = eng.perft(w, d); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); leaves2, _ = eng.perft(w, d); dt2 = time.perf_counter() - t0
    print(d, leaves, leaves == want, "first %.3f s, again %.3f s, %.0f G leaves/s" % (dt, dt2, leaves / dt2 / 1e9), list(per_root[:4]))
