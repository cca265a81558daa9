"""Small-table tier timing helper (development tool): python tools/small_bench.py [batch ...]
Times table registration (second set of keys: the first pays CUDA's lazy module load) and ecdsa_verify_small_kernel on device-resident
inputs with CUDA events, L2 flushed between launches; checks the mask is all-valid.  FABGPU_SMALL_THREADS selects the CTA width."""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("fabric-mod_b200")
from tools import workload  # noqa: E402


def main():
    batches = [int(x) for x in sys.argv[1:]] or [65536]
    if os.environ.get("SMALL_LIB"):                            # a build with another FAB_WS (window-width sweep)
        pkg.binding.LIB_PATH = os.path.abspath(os.environ["SMALL_LIB"])
        pkg.binding._LIB = None
    keys = int(os.environ.get("SMALL_KEYS", "4096"))
    dev = torch.device("cuda:0")
    nmax = max(batches)
    ctx = pkg.binding.Context(max_batch=nmax)
    w0 = workload.Workload(1024, 512, seed=5, nthreads=os.cpu_count())
    t0 = time.perf_counter(); ctx.keys_register_small(w0.keys_xy); torch.cuda.synchronize()
    print("first registration (512 keys, incl. lazy module load) %.2f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
    w = workload.Workload(nmax, keys, seed=workload.DEFAULT_SEED + 11, nthreads=os.cpu_count())
    t0 = time.perf_counter(); hs = ctx.keys_register_small(w.keys_xy); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("registration of %d keys: host %.2f ms, until built %.2f ms (%.2f us per key)" % (keys, (t1 - t0) * 1e3, (t2 - t0) * 1e3, (t2 - t0) * 1e6 / keys), flush=True)
    codes = ctx.small_raw_codes(hs)
    t = [torch.from_numpy(a).to(dev) for a in (w.digest, w.r, w.s)]
    ks = torch.from_numpy(np.ascontiguousarray(codes[w.key_idx])).to(dev)
    mask = torch.zeros(nmax // 32, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev)
    for n in batches:
        def go():
            ctx.verify_p256_device_keyed(2, ks.data_ptr(), 0, 0, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), n, mask.data_ptr(), 0, st.cuda_stream)
        for _ in range(3):
            go()
        torch.cuda.synchronize()
        assert bool((mask[: n // 32] == -1).all())
        reps = 8
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for k in range(reps):
            flush.fill_(k)
            ev[k][0].record(); go(); ev[k][1].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / reps
        print("small  threads=%s keys=%d n=%7d  %8.3f ms  %8.2f Mverify/s" % (os.environ.get("FABGPU_SMALL_THREADS", "auto"), keys, n, ms, n / ms / 1e3), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
