"""Diagnostic (development tool): device-stage times of one 10k-tx block whose MSP holds 2 000 client identities (FABGPU_BLOCK_EVENTS=1)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FABGPU_BLOCK_EVENTS"] = "1"
pkg = importlib.import_module("fabric-mod_b200")
from tools import blockgen
out = open(os.path.join(ROOT, "gpurun_out", "block_clients_probe.txt"), "w")
def say(*a):
    print(*a, file=out, flush=True); print(*a, flush=True)
ntx = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
c = pkg.binding.Context(max_batch=4096)
for name, net, seed in (("1 client", blockgen.Network(), 17), ("2000 clients", blockgen.Network(n_orgs=4, n_clients=2000, seed=0xC11E), 19)):
    blk, info = blockgen.build_block(net, ntx, 3, {}, seed=seed)
    ids = [(i.serialized, i.mspid, i.xy, i.valid) for i in net.msp_table]
    if len(ids) > 100:
        c.keys_register(np.stack([np.frombuffer(p.xy, np.uint8) for p in net.peers]))
    c.msp_configure(ids, net.policy_n_of(3), net.principals, net.channel)
    pin = c.block_buffer(len(info["env_blob"]))
    pin[:] = np.frombuffer(info["env_blob"], np.uint8)
    fast = os.environ.get("PROBE_FAST") == "1"
    for _ in range(1 if fast else 3):
        f = c.validate_envelopes(pin, info["env_off"])
    t0 = time.perf_counter()
    reps = 1 if fast else 5
    for _ in range(reps):
        f = c.validate_envelopes(pin, info["env_off"])
    ms = (time.perf_counter() - t0) / reps * 1e3
    say(name, "ms/block %.3f" % ms, "flags ok", not f.any(), "timing", [round(x, 1) for x in c.block_timing()], c.key_table_stats())
