"""Development tool: fabgpu_validate_block timing on a synthetic 10k-tx block for several host-thread counts."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("fabric-mod_b200")
from tools import blockgen
net = blockgen.Network()
blk, info = blockgen.build_block(net, 10000, 3, {}, seed=17)
ids = [(i.serialized, i.mspid, i.xy, i.valid) for i in net.msp_table]
for th, ev in ((16, "0"), (16, "1")):
    os.environ["FABGPU_GATE_THREADS"] = str(th)
    os.environ["FABGPU_BLOCK_EVENTS"] = ev
    ctx = pkg.binding.Context(max_batch=8192)
    ctx.msp_configure(ids, net.policy_n_of(3), net.principals, net.channel)
    eb, eo = info["env_blob"], info["env_off"]
    pinned = ctx.block_buffer(len(eb)); pinned[:] = np.frombuffer(eb, np.uint8)
    for _ in range(3):
        f = ctx.validate_envelopes(pinned, eo)
    assert not f.any()
    acc = np.zeros(10); t0 = time.perf_counter()
    for _ in range(10):
        ctx.validate_envelopes(pinned, eo)
    t1 = time.perf_counter()
    for _ in range(10):
        ctx.validate_envelopes(pinned, eo); acc += np.array(ctx.block_timing())
    t0 = t0 + (time.perf_counter() - t1)
    ms = (time.perf_counter() - t0) / 10 * 1e3
    print("threads %3d: %.2f ms/block  host us: enqueue %.0f wait-plan %.0f wait-rest %.0f dup %.0f | device us: copy+walk+hash(overlapped) %.0f - %.0f tail-resolve+sha %.0f verify %.0f decide %.0f"
          % (th, ms, *(acc[:4] / 10), *(acc[5:10] / 10)), flush=True)
    ctx.close()
