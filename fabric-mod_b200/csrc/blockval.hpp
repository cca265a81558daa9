// Block-level signature pre-pass: walks a serialized
// common.Block the way the reference's validator does, collects every signature the block needs checked
// (one creator signature + E endorsement signatures per transaction) so that they can be verified as ONE GPU batch, and
// afterwards replays the reference's per-transaction decision sequence on the results to produce the
// TRANSACTIONS_FILTER byte array.
//
//   TxValidator.Validate / ValidateTx       core/committer/txvalidator/v20/validator.go:182-267,300-455
//   validation.ValidateTransaction          core/common/validation/msgvalidation.go:248-320 (+ :26-64, :67-147, :167-245)
//   KeyLevelValidator.Validate              core/common/validation/statebased/validator_keylevel.go:243-259
//   policies.SignatureSetToValidIdentities  common/policies/policy.go:365-402
//   cauthdsl compile / evaluate             common/cauthdsl/cauthdsl.go:24-92
//   markTXIdDuplicates                      core/committer/txvalidator/v20/validator.go:283-297
//
// Verifying a superset of what the reference would verify cannot change any outcome: validity depends
// only on (key, signed bytes, signature).  Ledger-dependent checks (tx ids already committed, chaincode definitions,
// rw-sets, key-level policies) and config transactions are outside this pre-pass, as in the reference's own unit
// tests, which mock them (v20/validator_test.go:152-196).  Host code only (no CUDA in this file).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace fabgpu { namespace blockval {

// peer.TxValidationCode
enum : uint8_t {
    TX_VALID = 0, TX_NIL_ENVELOPE = 1, TX_BAD_PAYLOAD = 2, TX_BAD_COMMON_HEADER = 3, TX_BAD_CREATOR_SIGNATURE = 4,
    TX_INVALID_ENDORSER_TRANSACTION = 5, TX_INVALID_CONFIG_TRANSACTION = 6, TX_UNSUPPORTED_TX_PAYLOAD = 7, TX_BAD_PROPOSAL_TXID = 8,
    TX_DUPLICATE_TXID = 9, TX_ENDORSEMENT_POLICY_FAILURE = 10, TX_UNKNOWN_TX_TYPE = 13, TX_TARGET_CHAIN_NOT_FOUND = 14,
    TX_NOT_VALIDATED = 254, TX_INVALID_OTHER_REASON = 255
};

struct Seg { uint32_t off = 0, len = 0; };                 // a byte range of the block buffer

// The MSP as the validator sees it in steady state (msp/cache): serialized identity bytes -> index.
// Serialized identities are ~900-byte strings that recur tens of thousands of times per block, so the lookup hashes a
// 40-byte sample (length, head, tail) and confirms with one memcmp instead of hashing every byte.
struct MspTable {
    std::vector<std::string> serialized;                    // identity i's wire bytes
    std::unordered_multimap<uint64_t, int> by_sample;       // sample hash -> identity index
    std::vector<std::string> mspid;
    static uint64_t sample_hash(const uint8_t* p, size_t n) {
        uint64_t h = 1469598103934665603ull ^ n;
        const size_t head = n < 16 ? n : 16, tail = n < 24 ? n : 24;
        for (size_t i = 0; i < head; i++) h = (h ^ p[i]) * 1099511628211ull;
        for (size_t i = n - tail; i < n; i++) h = (h ^ p[i]) * 1099511628211ull;
        return h;
    }
    void add(const uint8_t* p, size_t n) {
        by_sample.emplace(sample_hash(p, n), (int)serialized.size());
        serialized.emplace_back((const char*)p, n);
    }
    int find(const uint8_t* p, size_t n) const {
        auto range = by_sample.equal_range(sample_hash(p, n));
        for (auto it = range.first; it != range.second; ++it) {
            const std::string& s = serialized[it->second];
            if (s.size() == n && memcmp(s.data(), p, n) == 0) return it->second;
        }
        return -1;
    }
    std::vector<uint8_t> keys_xy;                           // 64 bytes per identity
    std::vector<uint8_t> valid;                             // identity.Validate() outcome
};

struct PolicyNode { int32_t type, n, first_child, n_children; };   // type 0: NOutOf(n) over children; 1: SignedBy(principal n)

struct SigJob {                                             // one signature to verify
    int identity = -1;                                      // index into MspTable
    Seg msg[2];                                             // signed bytes = msg[0] || msg[1]
    Seg sig;                                                // DER signature
};

struct Endorsement { int identity = -1; int job = -1; };

struct TxPlan {
    uint8_t early = TX_NOT_VALIDATED;                       // decided by structure alone (before any signature)
    bool needs_creator = false;
    int creator_identity = -1;                              // -1: unknown to the MSP
    int creator_job = -1;                                   // -1: no signature to check (missing signature / unknown creator)
    uint32_t htype = 0;
    bool endorser_parse_ok = false;                         // validateEndorserTransaction's structural part
    bool channel_ok = false;
    bool endorsements_parse_ok = true;
    Seg txid_ascii;                                         // ChannelHeader.tx_id
    Seg txid_msg[2];                                        // nonce, creator
    Seg phash_msg[3];                                       // channel_header, action signature_header, chaincode_proposal_payload
    Seg phash_claimed;                                      // ProposalResponsePayload.proposal_hash
    int check_job = -1;                                     // index of the (txid, proposal-hash) digest pair
    std::vector<Endorsement> ends;
};

struct BlockPlan {
    std::vector<TxPlan> txs;
    std::vector<SigJob> jobs;
    int n_check = 0;                                        // transactions that need the two check digests
    bool has_config_tx = false;
};

// Pieces for a multi-threaded caller: split the block into envelope ranges, plan disjoint transaction ranges with
// thread-local job lists (job indices local to the list), then merge (rebases the indices, numbers the check jobs).
bool split_block(const uint8_t* block, size_t len, std::vector<Seg>& envs);
// Thread-local job lists: creator jobs and endorsement jobs are kept apart so that the merged list holds all creator
// signatures first (their messages -- whole payloads -- are ~3x longer than endorsement messages; keeping the two classes
// in separate warps keeps the SHA-256 kernel's lanes balanced).
struct JobPart { std::vector<SigJob> creators, endorsements; };
void plan_range(const uint8_t* block, const std::vector<Seg>& envs, size_t lo, size_t hi, const MspTable& msp, const std::string& channel,
                TxPlan* txs, JobPart& local);
void merge_plan(BlockPlan& plan, std::vector<JobPart>& parts, const std::vector<size_t>& bounds);
void decide_range(const uint8_t* block, const BlockPlan& plan, const MspTable& msp, const std::vector<PolicyNode>& policy,
                  const std::vector<std::string>& principals, const uint8_t* sig_valid, const uint8_t* txid_digests,
                  const uint8_t* phash_digests, size_t lo, size_t hi, uint8_t* flags, uint64_t* txid_hash);
// txid_hash[t]: 64-bit hash of transaction t's tx id, filled by decide_range (so the serial duplicate pass below does not
// chase ten thousand cache lines of the block).
void mark_duplicates(const uint8_t* block, const BlockPlan& plan, const uint64_t* txid_hash, uint8_t* flags);

// Parses the block and builds the plan (single-threaded form).  Returns false when the outer Block / BlockData does not parse.
bool plan_block(const uint8_t* block, size_t len, const MspTable& msp, const std::string& channel, BlockPlan& out);

// sig_valid[j] = signature job j verified; txid_digest / phash_digest: 32 bytes per check job.
// Writes one validation code per transaction.
void decide_block(const uint8_t* block, const BlockPlan& plan, const MspTable& msp, const std::vector<PolicyNode>& policy,
                  const std::vector<std::string>& principals, const uint8_t* sig_valid, const uint8_t* txid_digests,
                  const uint8_t* phash_digests, uint8_t* flags);

} }  // namespace fabgpu::blockval
