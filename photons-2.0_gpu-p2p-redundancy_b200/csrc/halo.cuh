// halo.cuh -- leaf-granular halo planning of the multi-rank device path (p2p_halo_* in include/p2p_b200.h).
//
// After a rank has walked its tree against the gathered topology of its peers (p2p_tree_walk_peers_packed) its task list
// references ghost leaf ids; what has to travel is exactly the particles of the referenced leaves.  The reference ships
// 27 P pruned images whole (1_Indexing/src/remotes.c:337-446, fmm.c:1067-1105).  Everything below runs on the device:
//   need : marks of the referenced ghost leaves -> per-leaf particle counts -> exclusive scan = the ghost leaf table
//   give : the marks the peers sent -> per (requester, leaf) counts -> exclusive scan = offsets into the send buffer
//   gather: one launch copies every requested leaf of every requester
// The host only learns the per-peer totals (the split sizes of the all-to-all-v).
#pragma once
#include <cuda_runtime.h>

namespace p2p {

constexpr int kHaloPeers = 16;

struct PeerMap {
    int first[kHaloPeers + 1];       // ghost-leaf index bounds of the peer slots (my own rank has no slot)
    long long leaf_off[kHaloPeers];  // offset (in int2) of the slot's leaf table {first particle, count} in the gathered topology
    int nslot;
};

__global__ void halo_mark_kernel(const int* __restrict__ ts, long long ntask, int nleaf_local, unsigned char* __restrict__ marks) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntask) return;
    const int s = ts[i];
    if (s >= nleaf_local) marks[s - nleaf_local] = 1;
}

// particles wanted per ghost leaf (0 if the list does not reference it)
__global__ void halo_need_counts_kernel(const unsigned char* __restrict__ marks, int nghost, PeerMap M, const int2* __restrict__ topo,
                                        unsigned int* __restrict__ cnt) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nghost) return;
    unsigned int n = 0;
    if (marks[g]) {
        int s = 0;
        while (s + 1 < M.nslot && g >= M.first[s + 1]) s++;
        n = (unsigned int)topo[M.leaf_off[s] + (g - M.first[s])].y;
    }
    cnt[g] = n;
}

// the ghost leaf table behind the local leaves: {first ghost particle, count}
__global__ void halo_ghost_table_kernel(const long long* __restrict__ off, const unsigned int* __restrict__ cnt, int nghost, int base,
                                        int max_count, int2* __restrict__ leaf, unsigned int* __restrict__ err) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nghost) return;
    const unsigned int n = cnt[g];
    if ((int)n > max_count) atomicOr(err, 1u);           // a source leaf larger than the force kernel's stage
    leaf[g] = make_int2(base + (int)off[g], (int)n);
}

// totals of consecutive segments of an exclusive scan (off has n + 1 entries)
__global__ void halo_segment_totals_kernel(const long long* __restrict__ off, const int* __restrict__ bound, int nseg, long long* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nseg) out[s] = off[bound[s + 1]] - off[bound[s]];
}

// particles to give per (requester, local leaf)
__global__ void halo_give_counts_kernel(const unsigned char* __restrict__ asked, long long n, int nleaf, const int2* __restrict__ leaf,
                                        unsigned int* __restrict__ cnt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cnt[i] = asked[i] ? (unsigned int)leaf[i % nleaf].y : 0u;
}

// one warp per (requester, leaf): copy the leaf's fixed-point particles to the requester's part of the send buffer
__global__ void halo_gather_kernel(const unsigned char* __restrict__ asked, long long n, int nleaf, const int2* __restrict__ leaf,
                                   const long long* __restrict__ off, const int4* __restrict__ part, int4* __restrict__ out) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n || !asked[w]) return;
    const int2 L = leaf[w % nleaf];
    const long long o = off[w];
    for (int k = lane; k < L.y; k += 32) out[o + k] = part[L.x + k];
}

}  // namespace p2p
