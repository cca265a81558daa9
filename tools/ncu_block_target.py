"""Profiling target (development tool): two fabgpu_validate_envelopes calls on a synthetic 10k-tx block."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("fabric-mod_b200")
from tools import blockgen
net = blockgen.Network()
blk, info = blockgen.build_block(net, 10000, 3, {}, seed=17)
ctx = pkg.binding.Context(max_batch=8192)
ctx.msp_configure([(i.serialized, i.mspid, i.xy, i.valid) for i in net.msp_table], net.policy_n_of(3), net.principals, net.channel)
eb, eo = info["env_blob"], info["env_off"]
pinned = ctx.block_buffer(len(eb)); pinned[:] = np.frombuffer(eb, np.uint8)
for _ in range(2):
    f = ctx.validate_envelopes(pinned, eo)
assert not f.any()
print("done")
