// See blockval.hpp.
#include "blockval.hpp"

#include <cstring>
#include <string_view>
#include <algorithm>
#include <unordered_map>

namespace fabgpu { namespace blockval {

namespace {

// ---- protobuf wire reader (only what proto.Unmarshal needs for these messages) -------------------------------------
struct Reader {
    const uint8_t* base;   // start of the whole block buffer (segments are offsets from here)
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;

    bool varint(uint64_t& v) {
        v = 0;
        for (int shift = 0; shift < 70; shift += 7) {
            if (p >= end) return ok = false;
            const uint8_t c = *p++;
            v |= (uint64_t)(c & 0x7f) << shift;
            if (!(c & 0x80)) return true;
        }
        return ok = false;
    }
    // next field: returns false at end of message or on error (check ok)
    bool next(uint32_t& field, uint32_t& wt, uint64_t& val, Seg& bytes) {
        if (p >= end) return false;
        uint64_t key;
        if (!varint(key)) return false;
        field = (uint32_t)(key >> 3); wt = (uint32_t)(key & 7);
        if (field == 0) return ok = false;
        switch (wt) {
            case 0: return varint(val);
            case 1: if (end - p < 8) return ok = false; p += 8; return true;
            case 5: if (end - p < 4) return ok = false; p += 4; return true;
            case 2: {
                uint64_t ln;
                if (!varint(ln)) return false;
                if ((uint64_t)(end - p) < ln) return ok = false;
                bytes.off = (uint32_t)(p - base); bytes.len = (uint32_t)ln;
                p += ln;
                return true;
            }
            default: return ok = false;
        }
    }
};

// Field selection for one message: bytes fields (last occurrence wins), one optional varint field, one optional repeated
// bytes field.  A selected field with the wrong wire type is an error, as in Go.
struct Want { uint32_t field; Seg* dst; bool* present; };

bool parse_msg(const uint8_t* base, Seg msg, const Want* wants, int nw, uint32_t uint_field = 0, uint64_t* uint_dst = nullptr,
               uint32_t uint_field2 = 0, uint64_t* uint_dst2 = nullptr, uint32_t rep_field = 0, std::vector<Seg>* rep = nullptr)
{
    Reader r{base, base + msg.off, base + msg.off + msg.len};
    uint32_t f, wt; uint64_t v; Seg b;
    while (r.next(f, wt, v, b)) {
        if (uint_field && f == uint_field) { if (wt != 0) return false; *uint_dst = v; continue; }
        if (uint_field2 && f == uint_field2) { if (wt != 0) return false; *uint_dst2 = v; continue; }
        if (rep_field && f == rep_field) { if (wt != 2) return false; rep->push_back(b); continue; }
        for (int i = 0; i < nw; i++) {
            if (wants[i].field == f) {
                if (wt != 2) return false;
                *wants[i].dst = b;
                if (wants[i].present) *wants[i].present = true;
                break;
            }
        }
    }
    return r.ok;
}

int lookup_identity(const MspTable& msp, const uint8_t* base, Seg s)
{
    return msp.find(base + s.off, s.len);
}

bool seg_equals(const uint8_t* base, Seg s, const std::string& str)
{
    return s.len == str.size() && memcmp(base + s.off, str.data(), s.len) == 0;
}

void plan_tx(const uint8_t* base, Seg env, const MspTable& msp, const std::string& channel, TxPlan& tx, JobPart& jobs)
{
    // Envelope{payload=1, signature=2}: v20/validator.go:313-320.  Zero-length data unmarshals to an empty Envelope
    // (protoutil.GetEnvelopeFromBlock), whose empty Payload has no header -> BAD_COMMON_HEADER, like the reference.
    Seg payload, signature; bool has_payload = false, has_sig = false;
    { const Want w[] = {{1, &payload, &has_payload}, {2, &signature, &has_sig}};
      if (!parse_msg(base, env, w, 2)) { tx.early = TX_INVALID_OTHER_REASON; return; } }
    // Payload{header=1, data=2}: msgvalidation.go:258-262
    Seg header, data; bool has_header = false, has_data = false;
    { const Want w[] = {{1, &header, &has_header}, {2, &data, &has_data}};
      if (!parse_msg(base, payload, w, 2)) { tx.early = TX_BAD_PAYLOAD; return; } }
    // validateCommonHeader: msgvalidation.go:120-147
    if (!has_header) { tx.early = TX_BAD_COMMON_HEADER; return; }
    Seg chdr_b, shdr_b; bool has_chdr = false, has_shdr = false;
    { const Want w[] = {{1, &chdr_b, &has_chdr}, {2, &shdr_b, &has_shdr}};
      if (!parse_msg(base, header, w, 2)) { tx.early = TX_BAD_COMMON_HEADER; return; } }
    uint64_t htype = 0, epoch = 0; Seg channel_id, txid;
    { const Want w[] = {{4, &channel_id, nullptr}, {5, &txid, nullptr}};
      if (!parse_msg(base, chdr_b, w, 2, 1, &htype, 6, &epoch)) { tx.early = TX_BAD_COMMON_HEADER; return; } }
    Seg creator, nonce;
    { const Want w[] = {{1, &creator, nullptr}, {2, &nonce, nullptr}};
      if (!parse_msg(base, shdr_b, w, 2)) { tx.early = TX_BAD_COMMON_HEADER; return; } }
    const uint32_t ht = (uint32_t)htype;                      // int32 on the wire
    if (!(ht == 1 || ht == 2 || ht == 3) || epoch != 0 || nonce.len == 0 || creator.len == 0) { tx.early = TX_BAD_COMMON_HEADER; return; }
    tx.htype = ht;
    // checkSignatureFromCreator: msgvalidation.go:26-64
    tx.needs_creator = true;
    tx.creator_identity = lookup_identity(msp, base, creator);
    if (has_sig && signature.len > 0 && has_payload && payload.len > 0 && tx.creator_identity >= 0) {
        SigJob j; j.identity = tx.creator_identity; j.msg[0] = payload; j.sig = signature;
        tx.creator_job = (int)jobs.creators.size();
        jobs.creators.push_back(j);
    }
    if (ht != 3) return;                                       // CONFIG / CONFIG_UPDATE: decided in decide_block
    tx.txid_ascii = txid;
    tx.txid_msg[0] = nonce; tx.txid_msg[1] = creator;
    tx.channel_ok = seg_equals(base, channel_id, channel);
    // validateEndorserTransaction: msgvalidation.go:167-245 (structure now, proposal-hash comparison after hashing)
    if (!has_data) return;
    std::vector<Seg> actions;
    if (!parse_msg(base, data, nullptr, 0, 0, nullptr, 0, nullptr, 1, &actions) || actions.size() != 1) return;
    Seg act_hdr, act_payload; bool has_ah = false;
    { const Want w[] = {{1, &act_hdr, &has_ah}, {2, &act_payload, nullptr}};
      if (!parse_msg(base, actions[0], w, 2)) return; }
    Seg a_creator, a_nonce;
    { const Want w[] = {{1, &a_creator, nullptr}, {2, &a_nonce, nullptr}};
      if (!parse_msg(base, act_hdr, w, 2) || a_nonce.len == 0 || a_creator.len == 0) return; }
    Seg cpp, action; bool has_cpp = false, has_action = false;
    { const Want w[] = {{1, &cpp, &has_cpp}, {2, &action, &has_action}};
      if (!parse_msg(base, act_payload, w, 2) || !has_action) return; }   // the Go code would dereference a nil Action
    Seg prp; std::vector<Seg> endorsements; bool has_prp = false;
    { const Want w[] = {{1, &prp, &has_prp}};
      if (!parse_msg(base, action, w, 1, 0, nullptr, 0, nullptr, 2, &endorsements)) return; }
    Seg phash;
    { const Want w[] = {{1, &phash, nullptr}};
      if (!parse_msg(base, prp, w, 1)) return; }
    if (!has_chdr || !has_ah || !has_cpp) return;             // GetProposalHash2 "nil arguments"
    tx.phash_msg[0] = chdr_b; tx.phash_msg[1] = act_hdr; tx.phash_msg[2] = cpp;
    tx.phash_claimed = phash;
    tx.endorser_parse_ok = true;
    // KeyLevelValidator.Validate: SignedData{prp || endorser, endorser, signature} per endorsement
    for (const Seg& e : endorsements) {
        Seg endorser, esig;
        const Want w[] = {{1, &endorser, nullptr}, {2, &esig, nullptr}};
        if (!parse_msg(base, e, w, 2)) { tx.endorsements_parse_ok = false; break; }
        Endorsement en;
        en.identity = lookup_identity(msp, base, endorser);
        if (en.identity >= 0 && esig.len > 0) {
            SigJob j; j.identity = en.identity; j.msg[0] = prp; j.msg[1] = endorser; j.sig = esig;
            en.job = (int)jobs.endorsements.size();
            jobs.endorsements.push_back(j);
        }
        tx.ends.push_back(en);
    }
}

// cauthdsl evaluator: common/cauthdsl/cauthdsl.go:24-92 (note the copy-in / copy-out of `used` around each sub-policy)
bool eval_policy(const std::vector<PolicyNode>& nodes, const std::vector<std::string>& principals, int idx,
                 const std::vector<const std::string*>& signer_msp, std::vector<char>& used)
{
    const PolicyNode& nd = nodes[idx];
    if (nd.type == 0) {
        int verified = 0;
        std::vector<char> scratch(used.size());
        for (int c = nd.first_child; c < nd.first_child + nd.n_children; c++) {
            scratch = used;
            if (eval_policy(nodes, principals, c, signer_msp, scratch)) { verified++; used = scratch; }
        }
        return verified >= nd.n;
    }
    if (nd.n < 0 || nd.n >= (int)principals.size()) return false;
    const std::string& want = principals[nd.n];
    for (size_t i = 0; i < signer_msp.size(); i++) {
        if (used[i]) continue;
        if (!signer_msp[i] || *signer_msp[i] != want) continue;   // SatisfiesPrincipal: MSP member match on a valid identity
        used[i] = 1;
        return true;
    }
    return false;
}

const char kHex[] = "0123456789abcdef";

}  // namespace

bool split_block(const uint8_t* block, size_t len, std::vector<Seg>& envs)
{
    envs.clear();
    Seg whole; whole.off = 0; whole.len = (uint32_t)len;
    Seg data; bool has_data = false;
    { const Want w[] = {{2, &data, &has_data}};
      if (!parse_msg(block, whole, w, 1)) return false; }
    if (has_data && !parse_msg(block, data, nullptr, 0, 0, nullptr, 0, nullptr, 1, &envs)) return false;
    return true;
}

void plan_range(const uint8_t* block, const std::vector<Seg>& envs, size_t lo, size_t hi, const MspTable& msp, const std::string& channel,
                TxPlan* txs, JobPart& local)
{
    for (size_t i = lo; i < hi; i++) plan_tx(block, envs[i], msp, channel, txs[i], local);
}

void merge_plan(BlockPlan& plan, std::vector<JobPart>& parts, const std::vector<size_t>& bounds)
{
    plan.jobs.clear(); plan.n_check = 0; plan.has_config_tx = false;
    size_t n_creators = 0, total = 0;
    for (auto& p : parts) { n_creators += p.creators.size(); total += p.creators.size() + p.endorsements.size(); }
    plan.jobs.resize(total);
    size_t cbase = 0, ebase = n_creators;
    for (size_t part = 0; part < parts.size(); part++) {
        std::copy(parts[part].creators.begin(), parts[part].creators.end(), plan.jobs.begin() + cbase);
        std::copy(parts[part].endorsements.begin(), parts[part].endorsements.end(), plan.jobs.begin() + ebase);
        for (size_t t = bounds[part]; t < bounds[part + 1]; t++) {
            TxPlan& tx = plan.txs[t];
            if (tx.creator_job >= 0) tx.creator_job += (int)cbase;
            for (auto& e : tx.ends) if (e.job >= 0) e.job += (int)ebase;
        }
        cbase += parts[part].creators.size(); ebase += parts[part].endorsements.size();
    }
    for (auto& tx : plan.txs) {
        if (tx.early == TX_NOT_VALIDATED && tx.htype == 3) tx.check_job = plan.n_check++;
        if (tx.early == TX_NOT_VALIDATED && tx.htype == 1) plan.has_config_tx = true;
    }
}

bool plan_block(const uint8_t* block, size_t len, const MspTable& msp, const std::string& channel, BlockPlan& out)
{
    std::vector<Seg> envs;
    if (!split_block(block, len, envs)) return false;
    out.txs.assign(envs.size(), TxPlan());
    std::vector<JobPart> parts(1);
    plan_range(block, envs, 0, envs.size(), msp, channel, out.txs.data(), parts[0]);
    merge_plan(out, parts, {0, envs.size()});
    return true;
}

void decide_range(const uint8_t* block, const BlockPlan& plan, const MspTable& msp, const std::vector<PolicyNode>& policy,
                  const std::vector<std::string>& principals, const uint8_t* sig_valid, const uint8_t* txid_digests,
                  const uint8_t* phash_digests, size_t lo, size_t hi, uint8_t* flags, uint64_t* txid_hash)
{
    std::vector<int> seen;
    std::vector<const std::string*> signer_msp;
    std::vector<char> used;
    for (size_t t = lo; t < hi; t++) {
        const TxPlan& tx = plan.txs[t];
        uint8_t code = TX_VALID;
        do {
            if (tx.early != TX_NOT_VALIDATED) { code = tx.early; break; }
            // creator: identity known, certificate valid, signature verifies (msgvalidation.go:40-59)
            const bool creator_ok = tx.creator_identity >= 0 && msp.valid[tx.creator_identity] && tx.creator_job >= 0 && sig_valid[tx.creator_job];
            if (!creator_ok) { code = TX_BAD_CREATOR_SIGNATURE; break; }
            if (tx.htype == 1) { code = TX_NOT_VALIDATED; break; }           // config transaction: left to the CPU validator
            if (tx.htype != 3) { code = TX_UNSUPPORTED_TX_PAYLOAD; break; }
            // CheckTxID: hex(SHA-256(nonce || creator)) (proputils.go:357-376)
            {
                const uint8_t* dg = txid_digests + 32 * (size_t)tx.check_job;
                bool same = tx.txid_ascii.len == 64;
                for (int k = 0; same && k < 32; k++)
                    same = block[tx.txid_ascii.off + 2 * k] == (uint8_t)kHex[dg[k] >> 4] && block[tx.txid_ascii.off + 2 * k + 1] == (uint8_t)kHex[dg[k] & 15];
                if (!same) { code = TX_BAD_PROPOSAL_TXID; break; }
            }
            if (!tx.endorser_parse_ok) { code = TX_INVALID_ENDORSER_TRANSACTION; break; }
            if (tx.phash_claimed.len != 32 || memcmp(block + tx.phash_claimed.off, phash_digests + 32 * (size_t)tx.check_job, 32) != 0) {
                code = TX_INVALID_ENDORSER_TRANSACTION; break;
            }
            if (!tx.channel_ok) { code = TX_TARGET_CHAIN_NOT_FOUND; break; }
            if (!tx.endorsements_parse_ok) { code = TX_INVALID_OTHER_REASON; break; }
            // SignatureSetToValidIdentities (policy.go:365-402): in order; unknown identity skipped; an identity that
            // already verified is not checked again; a failed signature drops that entry only
            seen.clear(); signer_msp.clear();
            for (const Endorsement& e : tx.ends) {
                if (e.identity < 0) continue;
                bool dup = false;
                for (int s : seen) if (s == e.identity) { dup = true; break; }
                if (dup) continue;
                if (e.job < 0 || !sig_valid[e.job]) continue;
                seen.push_back(e.identity);
                signer_msp.push_back(msp.valid[e.identity] ? &msp.mspid[e.identity] : nullptr);
            }
            used.assign(signer_msp.size(), 0);
            if (policy.empty() || !eval_policy(policy, principals, 0, signer_msp, used)) { code = TX_ENDORSEMENT_POLICY_FAILURE; break; }
        } while (false);
        flags[t] = code;
        if (code == TX_VALID) {
            uint64_t h = 1469598103934665603ull;
            for (uint32_t k = 0; k < tx.txid_ascii.len; k++) h = (h ^ block[tx.txid_ascii.off + k]) * 1099511628211ull;
            txid_hash[t] = h;
        }
    }
}

// markTXIdDuplicates (v20/validator.go:283-297): among VALID transactions, a later one with an already seen tx id
void mark_duplicates(const uint8_t* block, const BlockPlan& plan, const uint64_t* txid_hash, uint8_t* flags)
{
    std::unordered_multimap<uint64_t, uint32_t> seen;                      // tx-id hash -> first transaction that carried it
    seen.reserve(plan.txs.size() * 2);
    for (size_t t = 0; t < plan.txs.size(); t++) {
        if (flags[t] != TX_VALID) continue;
        const Seg id = plan.txs[t].txid_ascii;
        if (id.len == 0) continue;
        bool dup = false;
        auto range = seen.equal_range(txid_hash[t]);
        for (auto it = range.first; it != range.second && !dup; ++it) {
            const Seg other = plan.txs[it->second].txid_ascii;            // equal hashes: confirm on the bytes
            dup = other.len == id.len && memcmp(block + other.off, block + id.off, id.len) == 0;
        }
        if (dup) flags[t] = TX_DUPLICATE_TXID;
        else seen.emplace(txid_hash[t], (uint32_t)t);
    }
}

void decide_block(const uint8_t* block, const BlockPlan& plan, const MspTable& msp, const std::vector<PolicyNode>& policy,
                  const std::vector<std::string>& principals, const uint8_t* sig_valid, const uint8_t* txid_digests,
                  const uint8_t* phash_digests, uint8_t* flags)
{
    std::vector<uint64_t> txid_hash(plan.txs.size() + 1, 0);
    decide_range(block, plan, msp, policy, principals, sig_valid, txid_digests, phash_digests, 0, plan.txs.size(), flags, txid_hash.data());
    mark_duplicates(block, plan, txid_hash.data(), flags);
}

} }  // namespace fabgpu::blockval
