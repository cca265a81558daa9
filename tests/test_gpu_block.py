"""GPU: device SHA-256 and the block-level pre-pass (fabgpu_validate_block) against hashlib and the oracle's block validator."""
import hashlib
import random

import numpy as np
import pytest

from oracle import blockval as ob
from tools import blockgen
from util import pkg
import blockutil

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = pkg().binding.Context(max_batch=8192)
    yield c
    c.close()


def test_sha256_segments_vs_hashlib(ctx):
    rnd = random.Random(9)
    buf = bytes(rnd.getrandbits(8) for _ in range(20000))
    jobs, exp = [], []
    lens = [0, 1, 3, 4, 5, 31, 32, 33, 54, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 129, 1000, 4097]
    for la in lens:
        for lb in (0, 1, 7, 64, 333):
            for lc in (0, 5):
                oa, ob_, oc = rnd.randrange(0, 9000), rnd.randrange(0, 9000), rnd.randrange(0, 9000)
                jobs.append((oa, ob_, oc, la, lb, lc))
                exp.append(hashlib.sha256(buf[oa:oa + la] + buf[ob_:ob_ + lb] + buf[oc:oc + lc]).digest())
    got = ctx.sha256_segments(buf, np.array(jobs, np.uint32))
    assert [bytes(g) for g in got] == exp


def _configure(ctx, net, n):
    ctx.msp_configure(blockutil.identities_of(net), net.policy_n_of(n), net.principals, net.channel)


def _oracle(blk, net, nodes, policies=None):
    """The oracle's flags for a device that holds exactly net.msp_table (identities outside it -> NOT_VALIDATED)."""
    ids = blockutil.identities_of(net)
    return ob.validate_block(blk, ids, net.channel, nodes, net.principals, policies=policies, known=[i[0] for i in ids])


def test_block_fault_classes_match_oracle(ctx):
    net = blockgen.Network()
    faults = blockutil.fault_map(70)
    blk, _ = blockgen.build_block(net, 70, 3, faults, seed=11)
    exp = _oracle(blk, net, net.policy_n_of(3))
    _configure(ctx, net, 3)
    got = ctx.validate_block(blk)
    assert got.tolist() == exp.tolist()
    assert len(set(exp.tolist())) >= 10
    # other thresholds / nested policy reuse the same block
    for n in (2, 4):
        _configure(ctx, net, n)
        assert ctx.validate_block(blk).tolist() == _oracle(blk, net, net.policy_n_of(n)).tolist()


def test_block_per_chaincode_policies_and_dispatcher_checks(ctx):
    """Every fault class incl. the plugin dispatcher's (dispatcher.go:102-221) with a policy PER NAMESPACE: a transaction that writes to
    a second chaincode is held to that chaincode's policy too; one that only reads it is not; an unknown namespace or identity -> 254;
    21 endorsements (both encodings of every certificate) de-duplicate to four signers."""
    net = blockgen.Network()
    n = 3 * len(blockgen.FAULTS) + 4
    faults = blockutil.fault_map(n)
    assert set(faults.values()) == set(blockgen.FAULTS)
    blk, info = blockgen.build_block(net, n, 3, faults, seed=11)
    nodes, pol = net.policies_for(3)
    exp = _oracle(blk, net, nodes, pol)
    ctx.msp_configure(blockutil.identities_of(net), nodes, net.principals, net.channel, policies=pol)
    assert ctx.validate_block(blk).tolist() == exp.tolist()
    assert ctx.validate_envelopes(info["env_blob"], info["env_off"]).tolist() == exp.tolist()
    byf = {f: int(exp[t]) for t, f in faults.items()}
    assert byf["writes_strict_namespace"] == ob.ENDORSEMENT_POLICY_FAILURE and byf["reads_strict_namespace"] == ob.VALID
    assert byf["writes_unknown_namespace"] == ob.NOT_VALIDATED and byf["unknown_endorser"] == ob.NOT_VALIDATED and byf["unknown_creator"] == ob.NOT_VALIDATED
    assert byf["same_cert_two_encodings"] == ob.ENDORSEMENT_POLICY_FAILURE and byf["many_endorsements"] == ob.VALID
    assert {ob.BAD_HEADER_EXTENSION, ob.BAD_RESPONSE_PAYLOAD, ob.BAD_RWSET, ob.ILLEGAL_WRITESET, ob.INVALID_CHAINCODE, ob.INVALID_OTHER_REASON} <= set(exp.tolist())


def test_block_2000_tx_mixed_and_pinned_buffer(ctx):
    net = blockgen.Network(n_orgs=4, n_clients=3)
    rnd = random.Random(4)
    faults = {t: rnd.choice(blockgen.FAULTS) for t in rnd.sample(range(1, 2000), 240)}
    blk, binfo = blockgen.build_block(net, 2000, 3, faults, seed=13)
    exp = _oracle(blk, net, net.policy_n_of(3))
    _configure(ctx, net, 3)
    assert ctx.validate_block(blk).tolist() == exp.tolist()
    assert ctx.validate_envelopes(binfo["env_blob"], binfo["env_off"]).tolist() == exp.tolist()      # Block.Data.Data form
    pinned = ctx.block_buffer(len(blk))
    pinned[:] = np.frombuffer(blk, np.uint8)
    assert ctx.validate_block(pinned).tolist() == exp.tolist()
    assert 1700 < int((exp == ob.VALID).sum()) < 1850


def test_config3_block_replay_10k_tx(ctx):
    # BASELINE.json configs[2]: 10 k synthetic txs x 3 endorsements, 3-of-4 policy: 40 000 verifications, every flag VALID
    net = blockgen.Network()
    blk, info = blockgen.build_block(net, 10000, 3, {}, seed=17)
    assert info["n_sigs"] == 40000
    _configure(ctx, net, 3)
    pinned = ctx.block_buffer(len(blk))
    pinned[:] = np.frombuffer(blk, np.uint8)
    got = ctx.validate_block(pinned)
    assert got.shape[0] == 10000 and (got == ob.VALID).all()
    exp = _oracle(blk, net, net.policy_n_of(3))
    assert (exp == ob.VALID).all()
    # one flipped byte inside one endorsement signature flips exactly that transaction
    b = bytearray(blk)
    marker = net.peers[1].serialized
    pos = blk.index(marker, len(blk) // 2) + len(marker) + 10
    b[pos] ^= 0x20
    got2 = ctx.validate_block(bytes(b))
    exp2 = _oracle(bytes(b), net, net.policy_n_of(3))
    assert got2.tolist() == exp2.tolist() and int((got2 != ob.VALID).sum()) == 1


def test_block_survives_key_table_eviction():
    """Identity tables recycled by unrelated registrations between fabgpu_msp_configure and a block: still exact."""
    import os
    from tools import workload
    os.environ["FABGPU_KEY_SLOTS"] = "8"
    try:
        c = pkg().binding.Context(max_batch=4096)
    finally:
        del os.environ["FABGPU_KEY_SLOTS"]
    net = blockgen.Network()
    faults = blockutil.fault_map(60)
    blk, binfo = blockgen.build_block(net, 60, 3, faults, seed=23)
    exp = _oracle(blk, net, net.policy_n_of(3))
    c.msp_configure(blockutil.identities_of(net), net.policy_n_of(3), net.principals, net.channel)
    assert c.validate_envelopes(binfo["env_blob"], binfo["env_off"]).tolist() == exp.tolist()
    other = workload.Workload(64, 8, seed=99)
    assert (c.keys_register(other.keys_xy) >= 0).all()          # evicts every identity's table
    assert c.validate_envelopes(binfo["env_blob"], binfo["env_off"]).tolist() == exp.tolist()
    c.keys_register(other.keys_xy)
    assert c.validate_block(blk).tolist() == exp.tolist()       # serialized-Block entry point after another eviction
    c.close()


def test_two_blocks_in_flight(ctx):
    """fabgpu_validate_*_async / fabgpu_validate_wait: two different blocks on the two slots at once, several rounds, both entry
    forms; each slot returns the oracle's flags for ITS block (separate device buffers, counters and pinned staging per slot)."""
    net = blockgen.Network(n_orgs=4, n_clients=3)
    ids = blockutil.identities_of(net)
    _configure(ctx, net, 3)
    blocks = []
    for k, (ntx, nfault, seed) in enumerate([(900, 120, 31), (333, 60, 32)]):
        rnd = random.Random(seed)
        faults = {t: rnd.choice(blockgen.FAULTS) for t in rnd.sample(range(1, ntx), nfault)}
        blk, info = blockgen.build_block(net, ntx, 3, faults, seed=seed)
        exp = _oracle(blk, net, net.policy_n_of(3))
        blob = np.frombuffer(info["env_blob"], np.uint8)
        pinned = ctx.block_buffer(blob.shape[0], slot=k)
        pinned[:] = blob
        blocks.append((blk, pinned, info["env_off"], exp))
    binding = pkg().binding
    for rnd_i in range(4):
        order = (0, 1) if rnd_i % 2 == 0 else (1, 0)
        caps = {}
        for sl in order:
            blk, pinned, eoff, _ = blocks[sl]
            caps[sl] = ctx.validate_envelopes_async(sl, pinned, eoff) if rnd_i < 2 else ctx.validate_block_async(sl, blk)
        with pytest.raises(binding.FabGpuError):                              # a slot holds one block at a time
            ctx.validate_envelopes_async(0, blocks[0][1], blocks[0][2])
        for sl in order:
            got = ctx.validate_wait(sl, caps[sl])
            assert got.tolist() == blocks[sl][3].tolist(), (rnd_i, sl)
    with pytest.raises(binding.FabGpuError):                                  # nothing in flight
        ctx.validate_wait(1, 16)
    # the blocking calls still work afterwards, and an empty envelope list goes through both halves
    assert ctx.validate_envelopes(blocks[1][1], blocks[1][2]).tolist() == blocks[1][3].tolist()
    assert ctx.validate_envelopes_async(1, np.zeros(1, np.uint8), np.zeros(1, np.uint32)) == 0
    assert ctx.validate_wait(1, 1).shape[0] == 0
