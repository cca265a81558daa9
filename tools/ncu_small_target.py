"""Profiling target (development tool): a few launches of ecdsa_verify_small_kernel on a 64k batch over 4 096 small tables, for
   ncu --set full -k regex:ecdsa_verify_small_kernel -s 2 -c 1 python tools/ncu_small_target.py"""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("fabric-mod_b200")
from tools import workload
n = 65536
w = workload.Workload(n, 4096, seed=workload.DEFAULT_SEED + 11, nthreads=os.cpu_count())
ctx = pkg.binding.Context(max_batch=n)
dev = torch.device("cuda:0")
t = [torch.from_numpy(a).to(dev) for a in (w.digest, w.r, w.s)]
codes = ctx.small_raw_codes(ctx.keys_register_small(w.keys_xy))
ks = torch.from_numpy(np.ascontiguousarray(codes[w.key_idx])).to(dev)
mask = torch.zeros(n // 32, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream(dev)
for _ in range(4):
    ctx.verify_p256_device_keyed(2, ks.data_ptr(), 0, 0, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), n, mask.data_ptr(), 0, st.cuda_stream)
torch.cuda.synchronize()
assert bool((mask == -1).all())
print("ncu small target done")
