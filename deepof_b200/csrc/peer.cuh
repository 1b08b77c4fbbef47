// Data-parallel gradient exchange over NVLink peer memory (ONE all-reduce(sum) of the flat fp32 gradient per
// step).  Replaces what DDP's bucketed NCCL all-reduce does for the reference (deepof/clustering/training.py:1567, 164:
// loss.backward() under DistributedDataParallel), for a payload of 86 KB - 4 MB where the NCCL call is pure latency
// (c10d stream hop + launch + rendezvous = 0.26 - 0.32 ms per step measured at N = 2..8).
//
// Every rank owns one SYMMETRIC buffer [gradient | signal pad] mapped into all peers (torch symmetric memory: cuMem +
// fabric / fd handles — plumbing); the backward kernels write the rank-local gradient straight into it.  Per step:
//   peer_barrier (all gradients written)  ->  peer_reduce: every rank reads ALL peers' gradients over NVLink and sums them
//   in rank order 0..W-1 (one-shot all-reduce: bit-identical sums on every rank, no second exchange phase)
//   ->  peer_barrier (all peers done reading; the next step may overwrite).
// Signals are monotonically increasing epochs written with st.release.sys into the peer's pad and polled with
// ld.acquire.sys; spins are bounded (a lost peer faults the kernel instead of hanging the GPU).
#pragma once
#include "common.cuh"

#define DOF_PEER_MAX 16

struct PeerPtrs { void* p[DOF_PEER_MAX]; };

__global__ void peer_barrier_kernel(const PeerPtrs pads, int world, int rank, int epoch) {
    const int p = threadIdx.x;
    if (p >= world) return;
    __threadfence_system();
    int* theirs = reinterpret_cast<int*>(pads.p[p]) + rank;           // my slot in peer p's pad
    const int* mine = reinterpret_cast<const int*>(pads.p[rank]) + p;  // peer p's slot in my pad
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    int v = 0;
    for (unsigned long long i = 0; i < (1ull << 28); i++) {
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if (v >= epoch) return;
    }
    __trap();
}

// out[i] = sum_r peers[r][i], r = 0..world-1 in order; n4 = number of float4
__global__ void __launch_bounds__(256) peer_reduce_kernel(const PeerPtrs peers, int world, float4* __restrict__ out, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; r++) {
            float4 v;
            // volatile: never served from this SM's L1 (the peers rewrite the buffer every step)
            asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "l"(reinterpret_cast<const float4*>(peers.p[r]) + i) : "memory");
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        out[i] = acc;
    }
}
