// TEST HELPER (not product code): exposes the product's host-side block logic (fabric-mod_b200/csrc/blockval.cpp:
// plan_block / decide_block) without any GPU, so tests can feed it signature verdicts and digests computed by the oracle
// and hashlib and compare the resulting TRANSACTIONS_FILTER with the oracle's block validator on the CPU-only box.
#include <cstring>
#include <string>
#include <vector>
#include "../../fabric-mod_b200/csrc/blockval.cpp"

using namespace fabgpu::blockval;

struct Handle {
    std::vector<uint64_t> th;
    MspTable msp; std::vector<PolicyNode> policy; std::vector<std::string> principals; std::string channel;
    BlockPlan plan; std::vector<uint8_t> block;
};

extern "C" {

void* bv_new(const uint8_t* id_blob, const uint32_t* id_off, const uint8_t* mspid_blob, const uint32_t* mspid_off, const uint8_t* keys_xy,
             const uint8_t* valid, int n_ids, const int32_t* nodes, int n_nodes, const uint8_t* pr_blob, const uint32_t* pr_off, int n_pr,
             const char* channel)
{
    Handle* h = new Handle();
    for (int i = 0; i < n_ids; i++) {
        h->msp.add(id_blob + id_off[i], id_off[i + 1] - id_off[i]);
        h->msp.mspid.emplace_back((const char*)mspid_blob + mspid_off[i], mspid_off[i + 1] - mspid_off[i]);
    }
    h->msp.keys_xy.assign(keys_xy, keys_xy + 64 * (size_t)n_ids);
    h->msp.valid.assign(valid, valid + n_ids);
    for (int i = 0; i < n_nodes; i++) h->policy.push_back({nodes[4 * i], nodes[4 * i + 1], nodes[4 * i + 2], nodes[4 * i + 3]});
    for (int i = 0; i < n_pr; i++) h->principals.emplace_back((const char*)pr_blob + pr_off[i], pr_off[i + 1] - pr_off[i]);
    h->channel = channel;
    return h;
}
void bv_free(void* p) { delete (Handle*)p; }

// returns the number of transactions (or -1); *n_jobs, *n_check are set
int bv_plan(void* p, const uint8_t* block, size_t len, int* n_jobs, int* n_check)
{
    Handle* h = (Handle*)p;
    h->block.assign(block, block + len);
    if (!plan_block(h->block.data(), len, h->msp, h->channel, h->plan)) return -1;
    *n_jobs = (int)h->plan.jobs.size(); *n_check = h->plan.n_check;
    return (int)h->plan.txs.size();
}
// job j: identity index, message segments (off0,len0,off1,len1), signature segment (off,len)
void bv_job(void* p, int j, int* identity, uint32_t* segs)
{
    const SigJob& sj = ((Handle*)p)->plan.jobs[j];
    *identity = sj.identity;
    segs[0] = sj.msg[0].off; segs[1] = sj.msg[0].len; segs[2] = sj.msg[1].off; segs[3] = sj.msg[1].len; segs[4] = sj.sig.off; segs[5] = sj.sig.len;
}
// check job c of transaction t (returns -1 if the transaction has none): txid segments (2) and proposal-hash segments (3)
int bv_check(void* p, int t, uint32_t* segs)
{
    const TxPlan& tx = ((Handle*)p)->plan.txs[t];
    if (tx.check_job < 0) return -1;
    segs[0] = tx.txid_msg[0].off; segs[1] = tx.txid_msg[0].len; segs[2] = tx.txid_msg[1].off; segs[3] = tx.txid_msg[1].len;
    for (int k = 0; k < 3; k++) { segs[4 + 2 * k] = tx.phash_msg[k].off; segs[5 + 2 * k] = tx.phash_msg[k].len; }
    return tx.check_job;
}
void bv_decide(void* p, const uint8_t* sig_valid, const uint8_t* txid_digests, const uint8_t* phash_digests, uint8_t* flags)
{
    Handle* h = (Handle*)p;
    decide_block(h->block.data(), h->plan, h->msp, h->policy, h->principals, sig_valid, txid_digests, phash_digests, flags);
}

}  // extern "C"
