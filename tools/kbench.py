"""Kernel-variant timing helper (development tool): python tools/kbench.py <lib.so> [batch ...]
Times fabgpu_verify_p256_device on device-resident inputs with CUDA events; checks the mask is all-valid."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("fabric-mod_b200")
from tools import workload  # noqa: E402


def main():
    lib = sys.argv[1]
    batches = [int(x) for x in sys.argv[2:]] or [65536]
    pkg.binding.LIB_PATH = os.path.abspath(lib)
    pkg.binding._LIB = None
    dev = torch.device("cuda:0")
    nmax = max(batches)
    w = workload.Workload(nmax, 64, seed=workload.DEFAULT_SEED + 2)
    ctx = pkg.binding.Context(max_batch=nmax)
    t = [torch.from_numpy(a).to(dev) for a in (w.qx(), w.qy(), w.digest, w.r, w.s)]
    mask = torch.zeros(nmax // 32, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev)
    import time
    t0 = time.perf_counter()
    slots = ctx.keys_register(w.keys_xy) & 0xFFF          # device-resident API takes raw slot indices
    print("%-32s keys_register(64 keys) %.1f ms" % (os.path.basename(lib), (time.perf_counter() - t0) * 1e3), flush=True)
    ks = torch.from_numpy(slots[w.key_idx]).to(dev)
    only = os.environ.get("KBENCH_ONLY")                      # "cached": the key-table kernel only (variant sweeps)
    modes = ("cached",) if only == "cached" else ("generic", "cached")
    for mode, n in [(m, n) for n in batches for m in modes]:
        def go():
            if mode == "generic":
                ctx.verify_p256_device(*[x.data_ptr() for x in t], n, mask.data_ptr(), 0, st.cuda_stream)
            else:
                ctx.verify_p256_device_keyed(True, ks.data_ptr(), 0, 0, t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(), n, mask.data_ptr(), 0, st.cuda_stream)
        for _ in range(3):
            go()
        torch.cuda.synchronize()
        assert bool((mask[: n // 32] == -1).all())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            go()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("%-32s %-8s n=%7d  %8.3f ms  %8.2f Mverify/s" % (os.path.basename(lib), mode, n, ms, n / ms / 1e3), flush=True)
    # pinned-slot path: H2D + kernel + D2H + sync per call (what fabgpu_verify_p256_keyed does), wall clock
    import time
    for n in ([] if only else batches):
        hb = ctx.host_buffers(0)
        for name, arr in (("qx", w.qx()), ("qy", w.qy()), ("e", w.digest), ("r", w.r), ("s", w.s)):
            hb[name][:n] = arr[:n]
        ctx.host_key_slots(0)[:n] = slots[w.key_idx][:n]
        for keyed in (True, False):
            f = (lambda: ctx.verify_p256_keyed(0, n)) if keyed else (lambda: ctx.verify_p256(0, n))
            for _ in range(3):
                f()
            t0 = time.perf_counter()
            for _ in range(10):
                f()
            ms = (time.perf_counter() - t0) / 10 * 1e3
            print("%-32s slot-%-6s n=%7d  %8.3f ms wall per call" % (os.path.basename(lib), "keyed" if keyed else "plain", n, ms), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
