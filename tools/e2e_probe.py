"""Development tool: phase times of fabgpu_bccsp_verify_batch for several host-thread counts."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("fabric-mod_b200")
from tools import workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
w = workload.Workload(n, 64, seed=workload.DEFAULT_SEED + 2)
doff = w.dig_off()
for th in (1, 4, 8, 16, 32, 64):
    os.environ["FABGPU_GATE_THREADS"] = str(th)
    ctx = pkg.binding.Context(max_batch=n)
    for _ in range(3):
        st = ctx.bccsp_verify_batch(w.keys_xy, w.key_idx, w.digest, doff, w.sigs, w.sig_off)
    assert (st == 0).all()
    t0 = time.perf_counter()
    reps = 10
    acc = np.zeros(4)
    for _ in range(reps):
        ctx.bccsp_verify_batch(w.keys_xy, w.key_idx, w.digest, doff, w.sigs, w.sig_off)
        acc += np.array(ctx.last_timing())
    ms = (time.perf_counter() - t0) / reps * 1e3
    print("gate threads %3d: %.3f ms/call (%.1f Mverify/s)  phases us: lookup %.0f gates %.0f device %.0f scatter %.0f" %
          (th, ms, n / ms / 1e3, *(acc / reps)), flush=True)
    ctx.close()
