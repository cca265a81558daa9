"""The device-side block logic (fabric-mod_b200/csrc/blockdev.cuh: walk, dispatcher checks, gates, decisions), compiled for the host,
against the oracle's restatement of the reference's validator (oracle/blockval.py), on the CPU: every fault class of
tools/blockgen.py, per-chaincode policies, policy-evaluator semantics, de-duplication, duplicates, structural corner cases."""
import numpy as np

from oracle import blockval as ob
from tools import blockgen
from tools import fabricpb as pb
import blockutil


# ---- the device-side implementation of the same logic (blockdev.cuh), executed on the host ---------------------------------
def _envs(blk):
    envs = pb.parse(pb.parse(blk, ob.S_BLOCK)["data"] or b"", ob.S_BLOCKDATA)["data"]
    off = np.zeros(len(envs) + 1, np.uint32)
    off[1:] = np.cumsum([len(e) for e in envs])
    return b"".join(envs), off


def _known(ids):
    return [i[0] for i in ids]


def test_device_logic_fault_classes_match_oracle():
    net = blockgen.Network()
    n = 3 * len(blockgen.FAULTS) + 4
    faults = blockutil.fault_map(n)
    assert set(faults.values()) == set(blockgen.FAULTS)
    blk, info = blockgen.build_block(net, n, 3, faults, seed=11, nthreads=2)
    ids = blockutil.identities_of(net)
    for k in (2, 3, 4):
        exp = ob.validate_block(blk, ids, net.channel, net.policy_n_of(k), net.principals, known=_known(ids))
        got = blockutil.device_logic_flags(info["env_blob"], info["env_off"], ids, net.channel, net.policy_n_of(k), net.principals)
        assert got.tolist() == exp.tolist(), k
    nodes = np.array([(0, 2, 1, 2), (1, 0, 0, 0), (0, 1, 3, 2), (1, 1, 0, 0), (1, 2, 0, 0)], np.int32)
    exp = ob.validate_block(blk, ids, net.channel, nodes, net.principals, known=_known(ids))
    assert blockutil.device_logic_flags(info["env_blob"], info["env_off"], ids, net.channel, nodes, net.principals).tolist() == exp.tolist()
    exp3 = ob.validate_block(blk, ids, net.channel, net.policy_n_of(3), net.principals, known=_known(ids))
    assert exp3[[t for t, f in faults.items() if f == "dup_txid"][0]] == ob.DUPLICATE_TXID
    for t in range(n):
        if t not in faults:
            assert exp3[t] == ob.VALID, t


def test_device_logic_per_chaincode_policies():
    """One policy per namespace (plugindispatcher/dispatcher.go:166-221): writes to a second chaincode bring its policy in, reads do
    not; namespaces / identities the device table lacks -> NOT_VALIDATED; signers de-duplicate on Mspid + certificate."""
    net = blockgen.Network()
    n = 3 * len(blockgen.FAULTS) + 4
    faults = blockutil.fault_map(n)
    blk, info = blockgen.build_block(net, n, 3, faults, seed=11, nthreads=2)
    ids = blockutil.identities_of(net)
    for k in (2, 3):
        nodes, pol = net.policies_for(k)
        exp = ob.validate_block(blk, ids, net.channel, nodes, net.principals, policies=pol, known=_known(ids))
        got = blockutil.device_logic_flags(info["env_blob"], info["env_off"], ids, net.channel, nodes, net.principals, policies=pol)
        assert got.tolist() == exp.tolist(), k
        byf = {f: int(exp[t]) for t, f in faults.items()}
        assert byf["reads_strict_namespace"] == ob.VALID and byf["writes_unknown_namespace"] == ob.NOT_VALIDATED
        assert byf["unknown_endorser"] == ob.NOT_VALIDATED and byf["unknown_creator"] == ob.NOT_VALIDATED and byf["many_endorsements"] == ob.VALID
        assert byf["writes_strict_namespace"] == (ob.ENDORSEMENT_POLICY_FAILURE if k == 3 else ob.VALID)      # 3 endorsements: 4-of-4 fails, 3-of-4 holds
        assert byf["same_cert_two_encodings"] == (ob.ENDORSEMENT_POLICY_FAILURE if k == 3 else ob.VALID)
        assert {ob.BAD_HEADER_EXTENSION, ob.BAD_RESPONSE_PAYLOAD, ob.BAD_RWSET, ob.ILLEGAL_WRITESET, ob.INVALID_CHAINCODE} <= set(exp.tolist())
    # the reference itself (every identity known to the MSP, no device in the picture) gives a verdict where the device abstains:
    # an unknown creator is a BAD_CREATOR_SIGNATURE there
    ref = ob.validate_block(blk, ids, net.channel, nodes, net.principals, policies=pol)
    t = [t for t, f in faults.items() if f == "unknown_creator"][0]
    assert ref[t] == ob.BAD_CREATOR_SIGNATURE and exp[t] == ob.NOT_VALIDATED


def test_policy_thresholds_and_dedup():
    # common/policies/policy_test.go:255-288 (two identical SignedData => one identity) and cauthdsl N-out-of semantics
    net = blockgen.Network()
    ids = blockutil.identities_of(net)
    blk, info = blockgen.build_block(net, 12, 3, {3: "dup_endorser", 5: "bad_endorsement_sig", 7: "same_cert_two_encodings"}, seed=3, nthreads=2)
    for k, expect_fail in ((2, []), (3, [3, 5, 7]), (4, list(range(12)))):
        exp = ob.validate_block(blk, ids, net.channel, net.policy_n_of(k), net.principals, known=_known(ids))
        got = blockutil.device_logic_flags(info["env_blob"], info["env_off"], ids, net.channel, net.policy_n_of(k), net.principals)
        assert got.tolist() == exp.tolist()
        assert [t for t in range(12) if exp[t] == ob.ENDORSEMENT_POLICY_FAILURE] == expect_fail


def test_device_logic_structural_corner_cases():
    net = blockgen.Network()
    ids = blockutil.identities_of(net)
    blk, _ = blockgen.build_block(net, 6, 3, {}, seed=5, nthreads=2)
    envs = pb.parse(pb.parse(blk, ob.S_BLOCK)["data"], ob.S_BLOCKDATA)["data"]
    good = envs[0]
    variants = [good, b"", b"\x0a\x05abc", pb.f_uint(1, 5), good + pb.f_bytes(9, b"unknown field is skipped"), pb.f_bytes(1, b""),
                envs[1][: len(envs[1]) // 2], pb.f_bytes(2, b"sig-only"), envs[2], envs[3]]
    blk2 = pb.block(9, variants)
    exp = ob.validate_block(blk2, ids, net.channel, net.policy_n_of(3), net.principals, known=_known(ids))
    assert exp[0] == ob.VALID and exp[4] == ob.DUPLICATE_TXID       # same transaction + an unknown field: parses, same tx id
    assert exp[1] == ob.BAD_COMMON_HEADER and exp[5] == ob.BAD_COMMON_HEADER
    blob, off = _envs(blk2)
    assert blockutil.device_logic_flags(blob, off, ids, net.channel, net.policy_n_of(3), net.principals).tolist() == exp.tolist()
    assert blockutil.device_logic_flags(b"", np.zeros(1, np.uint32), ids, net.channel, net.policy_n_of(3), net.principals).tolist() == []


def test_device_der_gate_matches_host_gate_fuzz():
    import random
    from oracle import goasn1, p256
    from util import pkg
    b = pkg().binding
    rnd = random.Random(77)
    seeds = [goasn1.marshal_ecdsa_signature(rnd.getrandbits(256), rnd.getrandbits(255)) for _ in range(30)]
    seeds += [goasn1.marshal_ecdsa_signature(rnd.getrandbits(8 * k), rnd.getrandbits(8 * j)) for k in (1, 31, 33, 40) for j in (1, 31, 32, 33)]
    seeds += [goasn1.marshal_ecdsa_signature(5, p256.HALF_N), goasn1.marshal_ecdsa_signature(5, p256.HALF_N + 1), goasn1.marshal_ecdsa_signature(-5, 7),
              goasn1.marshal_ecdsa_signature(0, 7), goasn1.marshal_ecdsa_signature(7, 0), b"", b"\x30\x00"]
    n = 0
    for sd in seeds:
        for it in range(120):
            m = bytearray(sd)
            if it:
                for _ in range(rnd.choice((1, 1, 2, 3))):
                    op = rnd.random()
                    if op < 0.6 and m:
                        m[rnd.randrange(len(m))] = rnd.getrandbits(8)
                    elif op < 0.8 and m:
                        del m[rnd.randrange(len(m))]
                    else:
                        m.insert(rnd.randrange(len(m) + 1), rnd.getrandbits(8))
            sig = bytes(m)
            st, r, s = b.gate_signature(sig)
            ok, r2, s2, dst = blockutil.device_gate(sig)
            assert ok == (st == b.ST_VALID), sig.hex()
            if sig:
                assert dst == st, sig.hex()                     # same status code as the host gate, error kinds included
            if ok:
                assert (r, s) == (r2, s2)
            n += 1
    assert n > 5000


def test_identity_groups_follow_the_certificate_not_the_bytes():
    """binding.identity_groups (what fabgpu_msp_identity_groups is fed): Mspid + certificate digest, as the reference's IdentityIdentifier
    (common/policies/policy.go:380-386, msp/identities.go:55-76) -- and the same partition as the oracle's identity_id."""
    from util import pkg
    net = blockgen.Network()
    ids = blockutil.identities_of(net)
    g = pkg().binding.identity_groups(ids)
    n_peers = len(net.peers)
    first_alt = len(ids) - n_peers
    for k in range(n_peers):
        assert g[k] == g[first_alt + k]                         # CRLF re-encoding of peer k's certificate: same identity
    assert len(set(g.tolist())) == len(ids) - n_peers
    oracle_ids = [ob.identity_id(i[0], i[1]) for i in ids]
    for a in range(len(ids)):
        for b in range(len(ids)):
            assert (g[a] == g[b]) == (oracle_ids[a] == oracle_ids[b])
    # the same certificate under ANOTHER MSP id is another identity; unparsable id_bytes fall back to the bytes themselves
    other = [(pb.serialized_identity("OtherMSP", net.peers[0].cert_pem), "OtherMSP", net.peers[0].xy, True),
             (pb.serialized_identity("Org1MSP", b"not a pem"), "Org1MSP", net.peers[0].xy, True),
             (pb.serialized_identity("Org1MSP", b"not a pem"), "Org1MSP", net.peers[0].xy, True)]
    g2 = pkg().binding.identity_groups(ids + other)
    assert g2[len(ids)] not in g2[: len(ids)].tolist() and g2[len(ids) + 1] == g2[len(ids) + 2] != g2[len(ids)]


def test_device_logic_with_hundreds_of_client_identities():
    """An MSP of 300 client certificates (plus peers): the identity lookup (sample_hash + msp_find of blockdev.cuh) must keep finding every
    identity when the table is large and the certificates share header, MSP id and PEM footer -- flags equal the oracle's on a block with
    every fault class, and the lookup's hash puts (almost) every identity into its own bucket."""
    net = blockgen.Network(n_orgs=4, n_clients=300, seed=0xC11E)
    n = 3 * len(blockgen.FAULTS) + 40
    faults = blockutil.fault_map(n)
    blk, info = blockgen.build_block(net, n, 3, faults, seed=29, nthreads=2)
    ids = blockutil.identities_of(net)
    assert len(ids) > 300
    exp = ob.validate_block(blk, ids, net.channel, net.policy_n_of(3), net.principals, known=_known(ids))
    got = blockutil.device_logic_flags(info["env_blob"], info["env_off"], ids, net.channel, net.policy_n_of(3), net.principals)
    assert got.tolist() == exp.tolist()
    assert int((exp == ob.VALID).sum()) > 40
    # the same hash, restated: FNV-style over 16 head + 32 middle + 32 pre-footer + 24 tail bytes
    M = (1 << 64) - 1

    def h(p):
        x = 1469598103934665603 ^ len(p)
        idx = list(range(min(16, len(p))))
        if len(p) >= 160:
            idx += list(range(len(p) // 2 - 16, len(p) // 2 + 16)) + list(range(len(p) - 72, len(p) - 40))
        idx += list(range(len(p) - min(24, len(p)), len(p)))
        for i in idx:
            x = ((x ^ p[i]) * 1099511628211) & M
        return x or 1
    hashes = {h(i[0]) for i in ids}
    assert len(hashes) == len(ids)                      # head + tail alone gave one value per MSP id
