// Kernels of the device-side block pre-pass (logic in blockdev.cuh): one thread per transaction.
#pragma once
#include <cuda_runtime.h>
#include "blockdev.cuh"

namespace fabgpu { namespace bdev {

// Walks every envelope, looks identities up, gates every DER signature and emits the SHA-256 / verify jobs.
// Creator jobs sit at the transaction's own index; endorsement jobs are appended after them through one atomic counter.
__global__ void __launch_bounds__(32)
block_plan_kernel(const uint8_t* __restrict__ block, const uint32_t* __restrict__ env_off, uint32_t T, MspDev msp, const uint8_t* __restrict__ channel,
                  uint32_t channel_len, TxDev* __restrict__ txs, JobArrays ja, uint32_t* __restrict__ n_end)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    Seg env; env.off = env_off[2 * t]; env.len = env_off[2 * t + 1] - env_off[2 * t];      // (begin, end) pairs
    plan_tx(block, env, t, msp, channel, channel_len, txs[t], ja, [&](uint32_t n) { return T + atomicAdd(n_end, n); });
}

// Replays the reference's decision order for every transaction on the verification bitmask and the digests.
__global__ void __launch_bounds__(128)
block_decide_kernel(const uint8_t* __restrict__ block, const TxDev* __restrict__ txs, uint32_t T, MspDev msp, PolicyDev pol,
                    const uint32_t* __restrict__ mask, const uint8_t* __restrict__ gate_ok, const uint8_t* __restrict__ digests, uint32_t J_cap,
                    uint8_t* __restrict__ flags, uint64_t* __restrict__ txid_hash, Seg* __restrict__ txid_seg)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    uint64_t h = 0;
    flags[t] = decide_tx(block, txs[t], t, msp, pol, [&](uint32_t j) { return gate_ok[j] && ((mask[j >> 5] >> (j & 31)) & 1u); }, digests, J_cap, &h);
    txid_hash[t] = h;
    txid_seg[t] = txs[t].txid_ascii;
}

} }  // namespace fabgpu::bdev
