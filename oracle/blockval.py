"""ORACLE (test infrastructure, NOT product code).

CPU restatement of the reference's block validator on the signature path -- what decides a transaction's flag in
TRANSACTIONS_FILTER as far as signatures are concerned:

  TxValidator.Validate / ValidateTx        core/committer/txvalidator/v20/validator.go:182-267,300-455
  validation.ValidateTransaction           core/common/validation/msgvalidation.go:248-320
    validateCommonHeader / ChannelHeader / SignatureHeader   :67-147
    checkSignatureFromCreator              :26-64   (DeserializeIdentity, Validate, Verify)
    protoutil.CheckTxID                    protoutil/proputils.go:357-376
    validateEndorserTransaction            :167-245 (one action; proposal-hash binding, protoutil/txutils.go:431-448)
  KeyLevelValidator.Validate               core/common/validation/statebased/validator_keylevel.go:243-259
                                           (SignedData = prp || endorser per endorsement)
  policies.SignatureSetToValidIdentities   common/policies/policy.go:365-402 (dedup, verify, drop on failure)
  cauthdsl compile (N-out-of / SignedBy)   common/cauthdsl/cauthdsl.go:24-92
  markTXIdDuplicates                       v20/validator.go:283-297

Out of the restated scope, exactly as the reference's own unit tests mock them away (v20/validator_test.go:152-196):
ledger lookups (duplicate tx ids already committed, chaincode definitions), read/write-set checks, key-level policies,
config transactions.  The MSP is a table `serialized identity -> (mspid, public key, Validate() outcome)`, which is what
msp/cache serves in steady state; SatisfiesPrincipal is MSP-member matching (identity.mspid == principal.mspid).

Every signature goes through oracle.bccsp_sw / oracle.fast (the restated bccsp/sw verifier), one at a time in the
reference's order; nothing is shared with the product's batching logic.
"""
import hashlib

import numpy as np

from tools import fabricpb as pb
from . import bccsp_sw as o
from . import fast

# peer.TxValidationCode (fabric-protos-go peer/transaction.pb.go)
VALID, NIL_ENVELOPE, BAD_PAYLOAD, BAD_COMMON_HEADER, BAD_CREATOR_SIGNATURE = 0, 1, 2, 3, 4
INVALID_ENDORSER_TRANSACTION, INVALID_CONFIG_TRANSACTION, UNSUPPORTED_TX_PAYLOAD, BAD_PROPOSAL_TXID, DUPLICATE_TXID = 5, 6, 7, 8, 9
ENDORSEMENT_POLICY_FAILURE, UNKNOWN_TX_TYPE, TARGET_CHAIN_NOT_FOUND = 10, 13, 14
NOT_VALIDATED, INVALID_OTHER_REASON = 254, 255

S_BLOCK = {1: ("header", "bytes"), 2: ("data", "bytes"), 3: ("metadata", "bytes")}
S_BLOCKDATA = {1: ("data", "rep_bytes")}
S_ENVELOPE = {1: ("payload", "bytes"), 2: ("signature", "bytes")}
S_PAYLOAD = {1: ("header", "bytes"), 2: ("data", "bytes")}
S_HEADER = {1: ("channel_header", "bytes"), 2: ("signature_header", "bytes")}
S_CHDR = {1: ("type", "uint"), 4: ("channel_id", "bytes"), 5: ("tx_id", "bytes"), 6: ("epoch", "uint")}
S_SHDR = {1: ("creator", "bytes"), 2: ("nonce", "bytes")}
S_TX = {1: ("actions", "rep_bytes")}
S_TXACTION = {1: ("header", "bytes"), 2: ("payload", "bytes")}
S_CAP = {1: ("chaincode_proposal_payload", "bytes"), 2: ("action", "bytes")}
S_CEA = {1: ("proposal_response_payload", "bytes"), 2: ("endorsements", "rep_bytes")}
S_ENDORSEMENT = {1: ("endorser", "bytes"), 2: ("signature", "bytes")}
S_PRP = {1: ("proposal_hash", "bytes"), 2: ("extension", "bytes")}


class Msp:
    """identities: list of (serialized: bytes, mspid: str, key_xy: 64 bytes, valid: bool)."""

    def __init__(self, identities):
        self.by_bytes = {bytes(ser): (i, mspid, bytes(xy), bool(valid)) for i, (ser, mspid, xy, valid) in enumerate(identities)}

    def deserialize(self, ser):
        return self.by_bytes.get(bytes(ser))


def _verify(xy: bytes, msg: bytes, sig: bytes) -> bool:
    """identity.Verify (msp/identities.go:169-196): SHA-256 then bccsp.Verify; any failure is an error."""
    key = o.P256PublicKey(int.from_bytes(xy[:32], "big"), int.from_bytes(xy[32:], "big"))
    return o.identity_verify(key, msg, sig, "SHA2") is None


def _verify_fast(xy: bytes, msg: bytes, sig: bytes) -> bool:
    st = fast.verify_batch(np.frombuffer(xy, np.uint8).reshape(1, 64), np.zeros(1, np.int32), np.frombuffer(hashlib.sha256(msg).digest(), np.uint8),
                           np.array([0, 32], np.uint32), np.frombuffer(sig, np.uint8) if sig else np.zeros(0, np.uint8), np.array([0, len(sig)], np.uint32))
    return int(st[0]) == o.VALID


def evaluate_policy(nodes, principals, identities_mspid):
    """cauthdsl.compile + evaluator (common/cauthdsl/cauthdsl.go:24-92, policy.go:97-108).
    nodes: int32[k,4] rows (type, n, first_child, n_children); identities_mspid: MSP ids of the valid, deduplicated signers."""
    def run(idx, used):
        t, n, first, cnt = [int(x) for x in nodes[idx]]
        if t == 0:                                   # NOutOf
            verified = 0
            for c in range(first, first + cnt):
                _used = list(used)
                if run(c, _used):
                    verified += 1
                    used[:] = _used
            return verified >= n
        want = principals[n]                          # SignedBy(n)
        for i, mspid in enumerate(identities_mspid):
            if used[i]:
                continue
            if mspid != want:                         # SatisfiesPrincipal: MSP member match
                continue
            used[i] = True
            return True
        return False
    return run(0, [False] * len(identities_mspid))


def validate_tx(env_bytes, msp: Msp, channel: str, nodes, principals, verify=_verify_fast):
    """One transaction -> (code, txid).  Mirrors ValidateTx + ValidateTransaction + the VSCC signature-policy check."""
    # zero-length data unmarshals to an empty Envelope (protoutil.GetEnvelopeFromBlock): no header -> BAD_COMMON_HEADER
    try:
        env = pb.parse(env_bytes, S_ENVELOPE)
    except pb.PbError:
        return INVALID_OTHER_REASON, ""               # v20/validator.go:313-320
    # ---- validation.ValidateTransaction ----
    try:
        payload = pb.parse(env["payload"] or b"", S_PAYLOAD)
    except pb.PbError:
        return BAD_PAYLOAD, ""
    try:
        if payload["header"] is None:
            raise pb.PbError("nil header")
        hdr = pb.parse(payload["header"], S_HEADER)
        chdr = pb.parse(hdr["channel_header"] or b"", S_CHDR)
        shdr = pb.parse(hdr["signature_header"] or b"", S_SHDR)
        if (chdr["type"] or 0) not in (1, 2, 3):
            raise pb.PbError("invalid header type")
        if (chdr["epoch"] or 0) != 0:
            raise pb.PbError("invalid epoch")
        if not shdr["nonce"]:
            raise pb.PbError("invalid nonce")
        if not shdr["creator"]:
            raise pb.PbError("invalid creator")
    except pb.PbError:
        return BAD_COMMON_HEADER, ""
    # checkSignatureFromCreator (msgvalidation.go:26-64)
    if not env["signature"] or not env["payload"]:
        return BAD_CREATOR_SIGNATURE, ""              # "nil arguments"
    ident = msp.deserialize(shdr["creator"])
    if ident is None or not ident[3]:
        return BAD_CREATOR_SIGNATURE, ""              # MSP error / certificate not valid
    if not verify(ident[2], env["payload"], env["signature"]):
        return BAD_CREATOR_SIGNATURE, ""
    htype = chdr["type"] or 0
    if htype == 1:
        return None, ""                               # CONFIG: outside the restated scope (configtx validation)
    if htype != 3:
        return UNSUPPORTED_TX_PAYLOAD, ""
    txid = (chdr["tx_id"] or b"").decode("utf-8", "replace")
    if txid != hashlib.sha256(shdr["nonce"] + shdr["creator"]).hexdigest():
        return BAD_PROPOSAL_TXID, ""
    # validateEndorserTransaction (msgvalidation.go:167-245)
    try:
        if payload["data"] is None:
            raise pb.PbError("nil arguments")
        tx = pb.parse(payload["data"], S_TX)
        if len(tx["actions"]) != 1:
            raise pb.PbError("only one action per transaction is supported")
        act = pb.parse(tx["actions"][0], S_TXACTION)
        ashdr = pb.parse(act["header"] or b"", S_SHDR)
        if not ashdr["nonce"] or not ashdr["creator"]:
            raise pb.PbError("invalid signature header")
        cap = pb.parse(act["payload"] or b"", S_CAP)
        if cap["action"] is None:
            raise pb.PbError("nil action")              # the Go code would dereference nil here; treated as invalid
        cea = pb.parse(cap["action"], S_CEA)
        prp = pb.parse(cea["proposal_response_payload"] or b"", S_PRP)
        if hdr["channel_header"] is None or act["header"] is None or cap["chaincode_proposal_payload"] is None:
            raise pb.PbError("nil arguments")
        phash = hashlib.sha256(hdr["channel_header"] + act["header"] + cap["chaincode_proposal_payload"]).digest()
        if phash != (prp["proposal_hash"] or b""):
            raise pb.PbError("proposal hash does not match")
    except pb.PbError:
        return INVALID_ENDORSER_TRANSACTION, ""
    # ---- ValidateTx ----
    if (chdr["channel_id"] or b"").decode("utf-8", "replace") != channel:
        return TARGET_CHAIN_NOT_FOUND, ""
    # ---- VSCC: endorsement policy over the signature set (validator_keylevel.go:243-259, policy.go:365-402) ----
    try:
        ends = [pb.parse(e, S_ENDORSEMENT) for e in cea["endorsements"]]
    except pb.PbError:
        return INVALID_OTHER_REASON, ""
    seen, signer_msps = set(), []
    prp_bytes = cea["proposal_response_payload"] or b""
    for e in ends:
        endorser = e["endorser"] or b""
        idn = msp.deserialize(endorser)
        if idn is None:
            continue                                    # invalid identity: skipped
        if idn[0] in seen:
            continue                                    # de-duplicated before any signature work
        if not verify(idn[2], prp_bytes + endorser, e["signature"] or b""):
            continue                                    # signature invalid: identity dropped
        seen.add(idn[0])
        if idn[3]:                                      # SatisfiesPrincipal re-validates the identity (msp/mspimpl.go:583-600)
            signer_msps.append(idn[1])
        else:
            signer_msps.append(None)
    if not evaluate_policy(nodes, principals, signer_msps):
        return ENDORSEMENT_POLICY_FAILURE, ""
    return VALID, txid


def validate_block(block_bytes, identities, channel, nodes, principals, verify=_verify_fast):
    """-> uint8 flags (TRANSACTIONS_FILTER), or raises for config transactions (outside the restated scope)."""
    msp = Msp(identities)
    blk = pb.parse(block_bytes, S_BLOCK)
    data = pb.parse(blk["data"] or b"", S_BLOCKDATA)["data"]
    flags = np.full(len(data), NOT_VALIDATED, np.uint8)
    txids = [""] * len(data)
    for i, d in enumerate(data):
        code, txid = validate_tx(d, msp, channel, nodes, principals, verify)
        if code is None:
            raise NotImplementedError("config transaction at index %d" % i)
        flags[i] = code
        if code == VALID:
            txids[i] = txid
    seen = set()
    for i, t in enumerate(txids):                       # markTXIdDuplicates (v20/validator.go:283-297)
        if not t:
            continue
        if t in seen:
            flags[i] = DUPLICATE_TXID
        else:
            seen.add(t)
    return flags
