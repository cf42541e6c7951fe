// device_tree.cu -- host side of the device-resident tree build and dual-tree walk (see device_tree.cuh).
// Compiled with -fmad=false and WITHOUT fast-math: the fp64 tree and MAC arithmetic must round exactly like
// the reference's host build.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "device_tree.cuh"
#include "midfield.cuh"
#include "p2p_ctx.h"

#define fail p2p_fail
using p2p::dt::ull;

constexpr int kBlockLevelMinNodes = 64;     // a level runs block-centric once it has this many nodes (and none above kBlockNodeMax particles)

struct p2p_dtree {
    // final tree
    long long npart = 0;
    int maxleaf = 0, nleaf = 0, nnode = 0, cap = 0, direct_start = 0, nlevel = 0;
    bool valid = false, built_here = false;
    DevBuf<double> box;
    DevBuf<int> son, node_npart, leaf_npart, leaf_ipart, node_leaf0, node_nleaf;
    DevBuf<double> node_split;
    // build
    DevBuf<double> x[3];
    DevBuf<int> perm, seg, seg_next, slot;
    DevBuf<unsigned char> flag;
    DevBuf<unsigned int> G, tile;
    DevBuf<int> t_start, t_len, t_parent, t_np0, t_child, t_nleaf, t_nnode, t_id, t_leafbase, child_cnt;
    DevBuf<double> t_split, t_lo, t_hi;
    DevBuf<int> choff, cmeta;
    DevBuf<double> capprox, capprox_end;
    DevBuf<long long> cinc;
    // mid-field (M2L lists from the walk, multipoles, local expansions)
    bool m2l_on = false, literal_d6 = false, mid_valid = false;
    int self_rank = 0;           // this rank's index among the peers (tags the M2L tasks of the local walk)
    DevBuf<int> mt, ms, mq;
    DevBuf<double> tb;                // tight particle bounds of the local leaves (minimal-image check of the walk)
    long long nm2l = 0;
    DevBuf<double> Mall, Lall, acc_mid;
    std::vector<int> lvl_begin, lvl_count;
    float ms_mid = 0.f;
    double walk_period = 0.0;
    int spec_min = 32768;        // nodes longer than this use the speculative chunked evaluation of the split mean
    int* d_scalar = nullptr;     // [0] next-level node count, [1] max leaf occupancy
    int* h_scalar = nullptr;     // pinned
    // walk
    DevBuf<ull> frontier[2];
    ull* d_wcount = nullptr;     // [0] next frontier, [1] tasks
    ull* h_wcount = nullptr;     // pinned
    unsigned int* d_dup = nullptr;
    int* d_maxw = nullptr;
    double max_leaf_width = 0.0, max_parent_width = -1.0;   // device-built trees; the latter bounds the separations of a local walk
    DevBuf<double> v[3], vtmp[3];     // device-resident stepping: velocities (tree order) and carry scratch
    DevBuf<int> gid, gidtmp;          // ... and the particles' ids
    bool stepping = false;
    long long resident = 0;      // particles currently in x[] / perm[]
    bool perm_is_local = true;   // perm[] indexes the array given to p2p_tree_build (false: global ids after a routing)
    long long walk_tasks = 0, walk_items = 0;
    int walk_levels = 0;
    int plain_max = p2p::dt::kSeqPlainMax;
    bool block_mode = true;
    bool node_mode = true;       // deep levels through node_level_kernel
    int block_level_min = kBlockLevelMinNodes;     // middle levels through block_level_kernel from this many nodes per level on (0: never)
    float ms_build = 0.f, ms_walk = 0.f;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    // p2p_forces_local: sums over its target chunks
    float sum_walk = 0.f, sum_csr = 0.f, sum_force = 0.f;
    int last_chunks = 0;
    long long max_chunk_tasks = 1LL << 29;
    std::vector<cudaEvent_t> chunk_ev;
    // chunk pipeline of p2p_forces_local: walk + packing of chunk k + 1 on pipe_stream while the force kernel of chunk k runs
    int pipe_chunks = 0;         // chunks a single-rank step is cut into at least (<= 1: no pipeline)
    int nranks = 1;              // p2p_set_rank; the pipeline is for single-rank steps (a multi-rank step overlaps its remote phase)
    cudaStream_t pipe_stream = nullptr;
    cudaEvent_t ev_pipe_built = nullptr, ev_pack_done[2] = {nullptr, nullptr}, ev_force_done[2] = {nullptr, nullptr};
};

void p2p_dtree_release(p2p_dtree* t) {
    if (!t) return;
    t->box.release(); t->son.release(); t->node_npart.release(); t->leaf_npart.release(); t->leaf_ipart.release();
    t->node_split.release(); t->node_leaf0.release(); t->node_nleaf.release();
    for (int k = 0; k < 3; k++) t->x[k].release();
    t->perm.release(); t->seg.release(); t->seg_next.release(); t->slot.release(); t->flag.release(); t->G.release(); t->tile.release();
    t->t_start.release(); t->t_len.release(); t->t_parent.release(); t->t_np0.release(); t->t_child.release(); t->t_nleaf.release();
    t->t_nnode.release(); t->t_id.release(); t->t_leafbase.release(); t->child_cnt.release(); t->t_split.release(); t->t_lo.release();
    t->t_hi.release();
    t->tb.release();
    t->mt.release(); t->ms.release(); t->mq.release(); t->Mall.release(); t->Lall.release(); t->acc_mid.release();
    for (int k = 0; k < 3; k++) { t->v[k].release(); t->vtmp[k].release(); }
    t->gid.release(); t->gidtmp.release();
    t->choff.release(); t->cmeta.release(); t->capprox.release(); t->capprox_end.release(); t->cinc.release();
    t->frontier[0].release(); t->frontier[1].release();
    if (t->d_scalar) cudaFree(t->d_scalar);
    if (t->h_scalar) cudaFreeHost(t->h_scalar);
    if (t->d_wcount) cudaFree(t->d_wcount);
    if (t->h_wcount) cudaFreeHost(t->h_wcount);
    if (t->d_dup) cudaFree(t->d_dup);
    if (t->d_maxw) cudaFree(t->d_maxw);
    if (t->e0) cudaEventDestroy(t->e0);
    if (t->e1) cudaEventDestroy(t->e1);
    for (cudaEvent_t e : t->chunk_ev) cudaEventDestroy(e);
    if (t->pipe_stream) cudaStreamDestroy(t->pipe_stream);
    for (cudaEvent_t e : {t->ev_pipe_built, t->ev_pack_done[0], t->ev_pack_done[1], t->ev_force_done[0], t->ev_force_done[1]})
        if (e) cudaEventDestroy(e);
    delete t;
}

namespace {

int get_tree(p2p_ctx* c, p2p_dtree** out) {
    if (!c->dtree) {
        p2p_dtree* t = new p2p_dtree();
        c->dtree = t;
        CU(cudaMalloc(&t->d_scalar, 4 * sizeof(int)));
        CU(cudaMallocHost(&t->h_scalar, 4 * sizeof(int)));
        CU(cudaMalloc(&t->d_wcount, p2p::dt::kWalkCounters * sizeof(ull)));
        CU(cudaMallocHost(&t->h_wcount, p2p::dt::kWalkCounters * sizeof(ull)));
        CU(cudaMalloc(&t->d_dup, sizeof(unsigned int)));
        CU(cudaMalloc(&t->d_maxw, 2 * sizeof(int)));
        CU(cudaEventCreate(&t->e0));
        CU(cudaEventCreate(&t->e1));
        static const int shifts[28][3] = {{0, 0, 0},
            {-1, -1, -1}, {-1, -1, 0}, {-1, -1, 1}, {-1, 0, -1}, {-1, 0, 0}, {-1, 0, 1}, {-1, 1, -1}, {-1, 1, 0}, {-1, 1, 1},
            {0, -1, -1}, {0, -1, 0}, {0, -1, 1}, {0, 0, -1}, {0, 0, 1}, {0, 1, -1}, {0, 1, 0}, {0, 1, 1},
            {1, -1, -1}, {1, -1, 0}, {1, -1, 1}, {1, 0, -1}, {1, 0, 0}, {1, 0, 1}, {1, 1, -1}, {1, 1, 0}, {1, 1, 1}, {0, 0, 0}};
        CU(cudaMemcpyToSymbol(p2p::dt::c_shift, shifts, sizeof shifts));
        CU(cudaMemcpyToSymbol(p2p::mf::c_mshift, shifts, sizeof shifts));
    }
    *out = c->dtree;
    return 0;
}

inline unsigned blocks(long long n, int b) { return (unsigned)((n + b - 1) / b); }

}  // namespace

extern "C" {

int p2p_tree_set_option(p2p_ctx* c, int seq_sum_plain_max) {
    USE(c);
    p2p_dtree* t;
    int r = get_tree(c, &t);
    if (r) return r;
    // -2: default thresholds but every node through the warp kernel (no block-per-node variant);
    // -3: block-per-node but no speculative chunked evaluation; -4: speculative evaluation for every node above one chunk;
    // -5: defaults, but the deep levels through the particle-wide kernels as well (no node-centric and no block-centric levels);
    // -6: no block-centric levels; -7: block-centric levels from the root on (every node up to kBlockNodeMax particles)
    t->node_mode = seq_sum_plain_max != -5;
    t->block_level_min = (seq_sum_plain_max == -5 || seq_sum_plain_max == -6) ? 0 : (seq_sum_plain_max == -7 ? 1 : kBlockLevelMinNodes);
    t->block_mode = seq_sum_plain_max != -2;
    t->spec_min = seq_sum_plain_max == -3 ? 0 : (seq_sum_plain_max == -4 ? 2048 : 32768);
    t->plain_max = seq_sum_plain_max < 0 ? p2p::dt::kSeqPlainMax : seq_sum_plain_max;
    return 0;
}

// Tree built elsewhere (p2p_build_localtree): boxes and sons for the device walk.  Ids as in p2p_tree_view:
// node_son holds the reference's global ids (leaf first_leaf + l, node first_node + n).
int p2p_tree_upload(p2p_ctx* c, int maxleaf, int nleaf, int nnode, int first_leaf, int first_node, const double* leaf_center,
                    const double* leaf_width, const int* node_son, const double* node_center, const double* node_width) {
    USE(c);
    if (nleaf < 0 || nnode < 1 || !node_son || !node_center || !node_width || (nleaf && (!leaf_center || !leaf_width)))
        return fail(P2P_ERR_ARG, "bad tree arrays");
    if (nleaf != c->nleaf) return fail(P2P_ERR_STATE, "tree has %d leaves, %d were uploaded with p2p_upload_leaves", nleaf, c->nleaf);
    p2p_dtree* t;
    int r = get_tree(c, &t);
    if (r) return r;
    const size_t nu = (size_t)nleaf + nnode;
    std::vector<double> box(6 * nu);
    for (int l = 0; l < nleaf; l++)
        for (int k = 0; k < 3; k++) { box[6 * (size_t)l + k] = leaf_center[3 * (size_t)l + k]; box[6 * (size_t)l + 3 + k] = leaf_width[3 * (size_t)l + k]; }
    for (int n = 0; n < nnode; n++)
        for (int k = 0; k < 3; k++) {
            box[6 * ((size_t)nleaf + n) + k] = node_center[3 * (size_t)n + k];
            box[6 * ((size_t)nleaf + n) + 3 + k] = node_width[3 * (size_t)n + k];
        }
    std::vector<int> son(2 * (size_t)nnode);
    for (size_t i = 0; i < son.size(); i++) {
        const int g = node_son[i];
        if (g < 0) son[i] = -1;
        else if (g >= first_node) {
            if (g - first_node >= nnode) return fail(P2P_ERR_ARG, "son id %d outside the node range", g);
            son[i] = nleaf + (g - first_node);
        } else {
            if (g < first_leaf || g - first_leaf >= nleaf) return fail(P2P_ERR_ARG, "son id %d outside the leaf range", g);
            son[i] = g - first_leaf;
        }
    }
    CU(t->box.reserve(box.size(), c->stream));
    CU(t->son.reserve(son.size(), c->stream));
    CU(cudaMemcpyAsync(t->box.p, box.data(), box.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(t->son.p, son.data(), son.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));           // the staging vectors die here
    t->npart = c->npart; t->maxleaf = maxleaf; t->nleaf = nleaf; t->nnode = nnode; t->valid = true; t->built_here = false; t->nm2l = 0;
    return 0;
}

// build_localtree on the device.  pos: host rows of 3 doubles in the caller's order (stride in doubles).
// On success the context holds the particles in tree order (fixed point), the leaves, and the tree for p2p_tree_walk.
}  // extern "C"

namespace {
// buffers of the level-synchronous build / routing for npart particles and up to ncap temp nodes
int reserve_build(p2p_ctx* c, p2p_dtree* t, long long npart, size_t ncap, p2p::dt::BuildArrays* A) {
    cudaStream_t st = c->stream;
    for (int k = 0; k < 3; k++) CU(t->x[k].reserve((size_t)npart, st, (size_t)std::min<long long>(t->resident, npart)));
    CU(t->perm.reserve((size_t)npart, st, (size_t)std::min<long long>(t->resident, npart)));
    CU(t->seg.reserve((size_t)npart, st)); CU(t->seg_next.reserve((size_t)npart, st));
    CU(t->slot.reserve((size_t)npart, st)); CU(t->flag.reserve((size_t)npart, st)); CU(t->G.reserve((size_t)npart + 1, st));
    const int ntile = (int)((npart + p2p::dt::kTile - 1) / p2p::dt::kTile);
    CU(t->tile.reserve((size_t)ntile + 1, st));
    CU(t->t_start.reserve(ncap, st)); CU(t->t_len.reserve(ncap, st)); CU(t->t_parent.reserve(ncap, st)); CU(t->t_np0.reserve(ncap, st));
    CU(t->t_child.reserve(2 * ncap, st)); CU(t->t_nleaf.reserve(ncap, st)); CU(t->t_nnode.reserve(ncap, st)); CU(t->t_id.reserve(ncap, st));
    CU(t->t_leafbase.reserve(ncap, st)); CU(t->child_cnt.reserve(ncap, st)); CU(t->t_split.reserve(ncap, st));
    CU(t->t_lo.reserve(3 * ncap, st)); CU(t->t_hi.reserve(3 * ncap, st));
    for (int k = 0; k < 3; k++) A->x[k] = t->x[k].p;
    A->perm = t->perm.p; A->seg = t->seg.p; A->seg_next = t->seg_next.p; A->flag = t->flag.p; A->G = t->G.p; A->slot = t->slot.p;
    A->t_start = t->t_start.p; A->t_len = t->t_len.p; A->t_parent = t->t_parent.p; A->t_np0 = t->t_np0.p; A->t_split = t->t_split.p;
    A->t_child = t->t_child.p; A->t_nleaf = t->t_nleaf.p; A->t_nnode = t->t_nnode.p; A->t_id = t->t_id.p; A->t_leafbase = t->t_leafbase.p;
    A->t_lo = t->t_lo.p; A->t_hi = t->t_hi.p;
    return 0;
}

int build_core(p2p_ctx* c, p2p_dtree* t, long long npart, int maxleaf, const double bdl[3], const double bdr[3], int direct_start);

// velocities and ids of the resident particles follow the permutation perm[] (tree build or routing partition) applied to the positions
int resident_carry(p2p_ctx* c, p2p_dtree* t) {
    const long long n = t->resident;
    p2p::dt::carry_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(t->perm.p, n, t->v[0].p, t->v[1].p, t->v[2].p, t->gid.p, t->vtmp[0].p,
                                                                              t->vtmp[1].p, t->vtmp[2].p, t->gidtmp.p);
    CU(cudaGetLastError());
    for (int k = 0; k < 3; k++) { std::swap(t->v[k].p, t->vtmp[k].p); std::swap(t->v[k].cap, t->vtmp[k].cap); }
    std::swap(t->gid.p, t->gidtmp.p); std::swap(t->gid.cap, t->gidtmp.cap);
    return 0;
}
}  // namespace

extern "C" {

int p2p_tree_build(p2p_ctx* c, const double* pos, int64_t stride, int64_t npart, int maxleaf, const double bdl[3],
                   const double bdr[3], int direct_start) {
    USE(c);
    if (npart < 1 || !pos || stride < 3 || maxleaf < 1 || !bdl || !bdr || direct_start < 0 || direct_start > 2)
        return fail(P2P_ERR_ARG, "bad arguments to p2p_tree_build");
    if (maxleaf > P2P_MAX_LEAF) return fail(P2P_ERR_ARG, "maxleaf %d exceeds P2P_MAX_LEAF %d", maxleaf, P2P_MAX_LEAF);
    if (npart > 0x7fffffffLL) return fail(P2P_ERR_ARG, "more than 2^31 particles per device");
    if (!c->box_set) return fail(P2P_ERR_STATE, "p2p_set_box must precede p2p_tree_build");
    p2p_dtree* t;
    int r = get_tree(c, &t);
    if (r) return r;
    cudaStream_t st = c->stream;
    t->valid = false;
    t->resident = 0;
    // ---- upload and split into coordinate arrays
    CU(c->stage.reserve((size_t)npart * 24, st));
    if (stride == 3) CU(cudaMemcpyAsync(c->stage.p, pos, (size_t)npart * 24, cudaMemcpyHostToDevice, st));
    else CU(cudaMemcpy2DAsync(c->stage.p, 24, pos, (size_t)stride * 8, 24, (size_t)npart, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(t->e0, st));
    for (int k = 0; k < 3; k++) CU(t->x[k].reserve((size_t)npart, st));
    CU(t->perm.reserve((size_t)npart, st)); CU(t->seg.reserve((size_t)npart, st));
    p2p::dt::soa_from_aos_kernel<<<blocks(npart, 256), 256, 0, st>>>(reinterpret_cast<const double*>(c->stage.p), npart, t->x[0].p, t->x[1].p,
                                                                     t->x[2].p, t->perm.p, t->seg.p);
    CU(cudaGetLastError());
    t->resident = npart;
    t->perm_is_local = true;
    return build_core(c, t, npart, maxleaf, bdl, bdr, direct_start);
}

}  // extern "C"

namespace {
// build_localtree over the particles resident in t->x / t->perm (current order = the order the reference would start from)
// levels with at most this many nodes evaluate the split mean with a block (or speculative chunks on all SMs) per node,
// levels with more nodes with a warp per node (P2P_B200_FEW_NODES overrides, for sweeps)
int few_nodes(int num_sm) {
    static int v = -1;
    if (v < 0) { const char* e = getenv("P2P_B200_FEW_NODES"); v = e ? atoi(e) : 0; if (v < 0) v = 0; }
    return v > 0 ? v : num_sm;
}

int build_core(p2p_ctx* c, p2p_dtree* t, long long npart, int maxleaf, const double bdl[3], const double bdr[3], int direct_start) {
    cudaStream_t st = c->stream;
    int r;
    int cap = (int)(2.0 * (double)npart / (double)maxleaf);   // the reference's NLEAF = NNODE capacity (fmm.c:203-209)
    if (cap > npart) cap = (int)npart + 1;
    const size_t ncap = (size_t)std::max(cap, 1) + 2;
    p2p::dt::BuildArrays A;
    if ((r = reserve_build(c, t, npart, ncap, &A))) return r;
    const int ntile = (int)((npart + p2p::dt::kTile - 1) / p2p::dt::kTile);
    CU(cudaMemsetAsync(A.seg, 0, (size_t)npart * sizeof(int), st));
    const int root[3] = {0, (int)npart, -1};
    CU(cudaMemcpyAsync(A.t_start, &root[0], 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(A.t_len, &root[1], 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(A.t_parent, &root[2], 4, cudaMemcpyHostToDevice, st));
    // ---- levels
    std::vector<int> lvl_begin{0}, lvl_count{1};
    int total_nodes = 1, longest = (int)npart;
    bool block_levels = false;
    for (int lvl = 0;; lvl++) {
        if (lvl > 256) return fail(P2P_ERR_ARG, "kd-tree deeper than 256 levels (more than maxleaf coincident particles?)");
        const int b = lvl_begin[lvl], n = lvl_count[lvl], dir = (direct_start + lvl) % 3;
        // few, long nodes: chunks on all SMs (speculative binade) or a block each; many nodes: a warp each
        const bool few = t->block_mode && n <= few_nodes(c->num_sm);
        const int block_min = few ? 2048 : 0x7fffffff;
        const int spec_min = (few && t->spec_min > 0) ? std::max(t->spec_min, block_min) : 0x7fffffff;
        if (longest > spec_min) {
            const int maxchunks = (int)(npart / p2p::dt::kChunk) + n + 1;
            CU(t->choff.reserve((size_t)n + 2, st)); CU(t->cmeta.reserve((size_t)maxchunks, st)); CU(t->capprox.reserve((size_t)maxchunks, st));
            CU(t->capprox_end.reserve((size_t)maxchunks, st)); CU(t->cinc.reserve(2 * (size_t)maxchunks, st));
            p2p::dt::SpecArrays Sp{t->choff.p, t->capprox.p, t->capprox_end.p, t->cinc.p, t->cmeta.p};
            p2p::dt::chunk_offsets_kernel<<<1, 1024, 0, st>>>(A, Sp, b, n, spec_min);
            p2p::dt::chunk_sum_kernel<<<std::min(maxchunks, c->num_sm * 8), 256, 0, st>>>(A, Sp, b, n, dir);
            p2p::dt::chunk_prefix_kernel<<<blocks((long long)n * 32, 128), 128, 0, st>>>(Sp, n);
            p2p::dt::chunk_transducer_kernel<<<std::min(maxchunks, c->num_sm * 4), p2p::dt::kBlockWarps * 32, 0, st>>>(A, Sp, b, n, dir);
            p2p::dt::chunk_combine_kernel<<<std::min(n, c->num_sm), p2p::dt::kBlockWarps * 32, 0, st>>>(A, Sp, b, n, dir);
        }
        // deep levels (no node above kNodeMax particles): a warp per node does the whole level (node_level_kernel)
        const bool node_mode = t->node_mode && longest <= p2p::dt::kNodeMax;
        // middle levels (enough nodes to fill the chip, none above kBlockNodeMax particles): a block per node for everything
        // but the split mean (block_level_kernel).  Sticky: its levels do not maintain the particle -> node map the
        // particle-wide kernels need.
        if (!node_mode && t->block_level_min > 0 && n >= t->block_level_min && longest <= p2p::dt::kBlockNodeMax) block_levels = true;
        const bool block_lvl = block_levels && !node_mode;
        const int warps = std::min(n, c->num_sm * 64);
        if (node_mode) {
            p2p::dt::node_level_kernel<<<blocks((long long)warps * 32, p2p::dt::kNodeWarps * 32), p2p::dt::kNodeWarps * 32, 0, st>>>(
                A, b, n, dir, maxleaf, t->plain_max, t->child_cnt.p);
        } else {
            if (longest > block_min)
                p2p::dt::mean_block_kernel<<<std::min(n, c->num_sm * 2), p2p::dt::kBlockWarps * 32, 0, st>>>(A, b, n, dir, block_min, spec_min);
            p2p::dt::mean_kernel<<<blocks((long long)warps * 32, 128), 128, 0, st>>>(A, b, n, dir, t->plain_max, block_min);
            if (block_lvl) {
                p2p::dt::block_level_kernel<<<std::min(n, c->num_sm * 3), p2p::dt::kBlockLevelThreads, 0, st>>>(A, b, n, dir, maxleaf, t->child_cnt.p);
            } else {
                p2p::dt::flag_kernel<<<blocks(npart, 256), 256, 0, st>>>(A, npart, dir);
                p2p::dt::flag_tile_sums_kernel<<<ntile, 256, 0, st>>>(A.flag, npart, t->tile.p);
                p2p::dt::flag_tile_offsets_kernel<<<1, 1024, 0, st>>>(t->tile.p, ntile);
                p2p::dt::flag_scan_apply_kernel<<<ntile, 256, 0, st>>>(A.flag, npart, t->tile.p, A.G);
                p2p::dt::split_kernel<<<blocks(n, 256), 256, 0, st>>>(A, b, n, maxleaf, t->child_cnt.p);
            }
        }
        p2p::dt::child_scan_kernel<<<1, 1024, 0, st>>>(t->child_cnt.p, n, t->d_scalar);
        CU(cudaMemsetAsync(t->d_scalar + 2, 0, sizeof(int), st));
        p2p::dt::children_kernel<<<blocks(n, 256), 256, 0, st>>>(A, b, n, maxleaf, t->child_cnt.p, b + n, (int)ncap, t->d_scalar + 2);
        if (!node_mode && !block_lvl) {
            p2p::dt::slot_kernel<<<blocks(npart, 256), 256, 0, st>>>(A, npart);
            p2p::dt::swap_kernel<<<blocks(npart, 256), 256, 0, st>>>(A, npart);
        }
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(t->h_scalar, t->d_scalar, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const int nnext = t->h_scalar[0];
        longest = t->h_scalar[2];
        std::swap(A.seg, A.seg_next);
        if (nnext == 0) { t->nlevel = lvl + 1; break; }
        total_nodes += nnext;
        if (total_nodes > cap)
            return fail(P2P_ERR_ARG, "kd-tree needs more than the reference's capacity of %d nodes (2 NPART / MAXLEAF)", cap);
        lvl_begin.push_back(b + n);
        lvl_count.push_back(nnext);
    }
    // ---- ids and boxes
    for (int lvl = t->nlevel - 1; lvl >= 0; lvl--) p2p::dt::count_up_kernel<<<blocks(lvl_count[lvl], 256), 256, 0, st>>>(A, lvl_begin[lvl], lvl_count[lvl]);
    int counts[2];
    CU(cudaMemcpyAsync(&counts[0], A.t_nleaf, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&counts[1], A.t_nnode, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const int nleaf = counts[0], nnode = counts[1];
    if (nleaf > cap) return fail(P2P_ERR_ARG, "kd-tree needs %d leaves, the reference's capacity is %d (2 NPART / MAXLEAF)", nleaf, cap);
    CU(t->box.reserve(6 * ((size_t)nleaf + nnode), st));
    CU(t->son.reserve(2 * (size_t)nnode, st));
    CU(t->node_npart.reserve((size_t)nnode, st)); CU(t->node_split.reserve((size_t)nnode, st));
    CU(t->node_leaf0.reserve((size_t)nnode, st)); CU(t->node_nleaf.reserve((size_t)nnode, st));
    CU(t->leaf_npart.reserve((size_t)nleaf + 1, st)); CU(t->leaf_ipart.reserve((size_t)nleaf + 1, st));
    p2p::dt::TreeOut O;
    O.nleaf = nleaf; O.box = t->box.p; O.son = t->son.p; O.node_npart = t->node_npart.p; O.node_split = t->node_split.p;
    O.leaf_npart = t->leaf_npart.p; O.leaf_ipart = t->leaf_ipart.p; O.max_width = t->d_maxw;
    O.node_leaf0 = t->node_leaf0.p; O.node_nleaf = t->node_nleaf.p;
    CU(cudaMemsetAsync(t->d_maxw, 0, 2 * sizeof(int), st));
    for (int lvl = 0; lvl < t->nlevel; lvl++)
        p2p::dt::assign_down_kernel<<<blocks(lvl_count[lvl], 256), 256, 0, st>>>(A, O, lvl_begin[lvl], lvl_count[lvl], (direct_start + lvl) % 3,
                                                                               bdl[0], bdl[1], bdl[2], bdr[0], bdr[1], bdr[2]);
    CU(cudaGetLastError());
    // ---- hand over to the force path: fixed-point particles in tree order, leaves, zeroed accelerations
    c->npart = npart; c->nghost = 0; c->nghostleaf = 0; c->csr_valid = false; c->nleaf = nleaf; c->bounds_n = 0;
    CU(c->part.reserve((size_t)npart + 1, st));
    CU(c->acc.reserve((size_t)npart + 1, st));
    CU(c->leaf.reserve((size_t)nleaf + 1, st));
    p2p::dt::pack_fixed_kernel<<<blocks(npart, 256), 256, 0, st>>>(A.x[0], A.x[1], A.x[2], npart, c->origin[0], c->origin[1], c->origin[2],
                                                                   4294967296.0 / c->extent, (float)c->mass, c->part.p);
    CU(cudaMemsetAsync(t->d_scalar + 1, 0, sizeof(int), st));
    if (nleaf) p2p::dt::leaf_pack_kernel<<<blocks(nleaf, 256), 256, 0, st>>>(O.leaf_npart, O.leaf_ipart, nleaf, c->leaf.p, t->d_scalar + 1);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(c->acc.p, 0, (size_t)npart * sizeof(float4), st));
    CU(cudaMemsetAsync(c->d_npairs_acc, 0, sizeof(unsigned long long), st));
    c->acc_tasks = 0;
    c->max_target_leaf = maxleaf;
    if ((r = p2p_update_occupancy(c))) return r;
    CU(cudaEventRecord(t->e1, st));
    float wmaxf[2] = {0.f, 0.f};
    CU(cudaMemcpyAsync(wmaxf, t->d_maxw, sizeof wmaxf, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    t->max_leaf_width = wmaxf[0];
    t->max_parent_width = wmaxf[1];
    CU(cudaEventElapsedTime(&t->ms_build, t->e0, t->e1));
    t->npart = npart; t->maxleaf = maxleaf; t->nleaf = nleaf; t->nnode = nnode; t->cap = cap; t->direct_start = direct_start;
    t->valid = true; t->built_here = true; t->mid_valid = false; t->nm2l = 0;
    t->lvl_begin = lvl_begin; t->lvl_count = lvl_count;
    return 0;
}
}  // namespace

extern "C" {

// ---- particle routing on the device (domain_decomposition: prepare_body_inOrderOf_domain + exchange) -----------------
// host slab -> resident arrays; perm[i] = first_index + i (the particle's global id, carried through the exchange)
int p2p_route_load(p2p_ctx* c, const double* pos, int64_t stride, int64_t n, int64_t first_index, int append) {
    USE(c);
    if (n < 0 || (n && !pos) || stride < 3 || first_index < 0 || first_index + n > 0x7fffffffLL) return fail(P2P_ERR_ARG, "bad slab");
    p2p_dtree* t;
    int r = get_tree(c, &t);
    if (r) return r;
    cudaStream_t st = c->stream;
    // append != 0: behind the particles already resident (a slab uploaded in pieces, so that neither the host nor the
    // staging buffer ever holds more than one piece)
    const long long base = append ? t->resident : 0;
    if (base + n > 0x7fffffffLL) return fail(P2P_ERR_ARG, "more than 2^31 particles per device");
    t->valid = false;
    if (!append) t->resident = 0;
    if (n == 0) return 0;
    CU(c->stage.reserve((size_t)n * 24, st));
    if (stride == 3) CU(cudaMemcpyAsync(c->stage.p, pos, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    else CU(cudaMemcpy2DAsync(c->stage.p, 24, pos, (size_t)stride * 8, 24, (size_t)n, cudaMemcpyHostToDevice, st));
    for (int k = 0; k < 3; k++) CU(t->x[k].reserve((size_t)(base + n), st, (size_t)base));
    CU(t->perm.reserve((size_t)(base + n), st, (size_t)base)); CU(t->seg.reserve((size_t)(base + n), st, (size_t)base));
    p2p::dt::soa_from_aos_kernel<<<blocks(n, 256), 256, 0, st>>>(reinterpret_cast<const double*>(c->stage.p), n, t->x[0].p + base, t->x[1].p + base,
                                                                 t->x[2].p + base, t->perm.p + base, t->seg.p + base);
    p2p::dt::iota_offset_kernel<<<blocks(n, 256), 256, 0, st>>>(t->perm.p + base, n, (int)first_index);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));              // the caller may reuse `pos` (pageable memory is staged by the driver)
    t->resident = base + n; t->perm_is_local = false;
    return 0;
}

// Partition the resident particles by the rank kd-tree (split[2P-1] in heap order, as p2p_domain_setup / _relax give
// them): afterwards they are grouped by destination rank, in rank order, each group in the order the reference's
// in-place partition leaves it; sendcount[r] = size of the group for rank r.  P must be a power of two.
int p2p_route_partition(p2p_ctx* c, int nproc, const double* split, int* sendcount) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t) return fail(P2P_ERR_STATE, "p2p_route_load first");
    if (nproc < 1 || (nproc & (nproc - 1)) || !split || !sendcount) return fail(P2P_ERR_ARG, "the device routing needs a power-of-two rank count");
    cudaStream_t st = c->stream;
    const long long n = t->resident;
    int levels = 0;
    while ((1 << levels) < nproc) levels++;
    if (n == 0) { for (int r = 0; r < nproc; r++) sendcount[r] = 0; return 0; }
    p2p::dt::BuildArrays A;
    int r = reserve_build(c, t, n, (size_t)2 * nproc + 2, &A);
    if (r) return r;
    DevBuf<double> dsplit;                      // tiny; released below
    CU(dsplit.reserve((size_t)2 * nproc, st));
    CU(cudaMemcpyAsync(dsplit.p, split, (size_t)(2 * nproc - 1) * 8, cudaMemcpyHostToDevice, st));
    const int ntile = (int)((n + p2p::dt::kTile - 1) / p2p::dt::kTile);
    CU(cudaMemsetAsync(A.seg, 0, (size_t)n * sizeof(int), st));
    const int root[3] = {0, (int)n, -1};
    CU(cudaMemcpyAsync(A.t_start, &root[0], 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(A.t_len, &root[1], 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(A.t_parent, &root[2], 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(A.t_split, dsplit.p, 8, cudaMemcpyDeviceToDevice, st));
    int begin = 0;
    for (int lvl = 0; lvl < levels; lvl++) {
        const int cnt = 1 << lvl, dir = lvl % 3;
        p2p::dt::route_flag_kernel<<<blocks(n, 256), 256, 0, st>>>(A, n, dir);
        p2p::dt::flag_tile_sums_kernel<<<ntile, 256, 0, st>>>(A.flag, n, t->tile.p);
        p2p::dt::flag_tile_offsets_kernel<<<1, 1024, 0, st>>>(t->tile.p, ntile);
        p2p::dt::flag_scan_apply_kernel<<<ntile, 256, 0, st>>>(A.flag, n, t->tile.p, A.G);
        p2p::dt::route_split_kernel<<<blocks(cnt, 128), 128, 0, st>>>(A, begin, cnt, dir);
        p2p::dt::route_children_kernel<<<blocks(cnt, 128), 128, 0, st>>>(A, begin, cnt, begin + cnt, dsplit.p, (2 << lvl) - 1, nproc - 1);
        p2p::dt::slot_kernel<<<blocks(n, 256), 256, 0, st>>>(A, n);
        p2p::dt::swap_kernel<<<blocks(n, 256), 256, 0, st>>>(A, n);
        CU(cudaGetLastError());
        std::swap(A.seg, A.seg_next);
        begin += cnt;
    }
    // the last level's runs are the per-rank groups (left to right = rank order for a power of two)
    CU(cudaMemcpyAsync(sendcount, A.t_len + begin, (size_t)nproc * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    dsplit.release();
    return 0;
}

int p2p_resident_count(p2p_ctx* c, int64_t* n) {
    if (!c || !n) return fail(P2P_ERR_ARG, "null pointer");
    *n = c->dtree ? c->dtree->resident : 0;
    return 0;
}

// copies of / into the resident particle arrays (device pointers owned by the caller: x, y, z doubles, idx int32)
int p2p_route_export(p2p_ctx* c, void* d_x, void* d_y, void* d_z, void* d_idx) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t) return fail(P2P_ERR_STATE, "nothing resident");
    const size_t n = (size_t)t->resident;
    if (n == 0) return 0;
    void* dst[3] = {d_x, d_y, d_z};
    for (int k = 0; k < 3; k++) CU(cudaMemcpyAsync(dst[k], t->x[k].p, n * 8, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(d_idx, t->perm.p, n * 4, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

int p2p_route_import(p2p_ctx* c, const void* d_x, const void* d_y, const void* d_z, const void* d_idx, int64_t n) {
    USE(c);
    if (n < 0 || n > 0x7fffffffLL || (n && (!d_x || !d_y || !d_z || !d_idx))) return fail(P2P_ERR_ARG, "bad arrays");
    p2p_dtree* t;
    int r = get_tree(c, &t);
    if (r) return r;
    t->valid = false; t->resident = 0;
    const void* src[3] = {d_x, d_y, d_z};
    for (int k = 0; k < 3; k++) {
        CU(t->x[k].reserve((size_t)n + 1, c->stream));
        if (n) CU(cudaMemcpyAsync(t->x[k].p, src[k], (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    CU(t->perm.reserve((size_t)n + 1, c->stream));
    if (n) CU(cudaMemcpyAsync(t->perm.p, d_idx, (size_t)n * 4, cudaMemcpyDeviceToDevice, c->stream));
    t->resident = n; t->perm_is_local = false;
    return 0;
}

// build_localtree over the resident particles (after p2p_route_import); perm keeps the global ids
int p2p_tree_build_resident(p2p_ctx* c, int maxleaf, const double bdl[3], const double bdr[3], int direct_start) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || t->resident < 1) return fail(P2P_ERR_STATE, "no resident particles (p2p_route_import)");
    if (maxleaf < 1 || maxleaf > P2P_MAX_LEAF || !bdl || !bdr || direct_start < 0 || direct_start > 2) return fail(P2P_ERR_ARG, "bad arguments");
    if (!c->box_set) return fail(P2P_ERR_STATE, "p2p_set_box must precede the tree build");
    t->valid = false;
    CU(cudaEventRecord(t->e0, c->stream));
    return build_core(c, t, t->resident, maxleaf, bdl, bdr, direct_start);
}

// ---- device-resident stepping: positions, velocities and ids live in HBM between steps ---------------
// pos / vel: host rows of 3 doubles (vel may be NULL = at rest); particle i gets id first_id + i
int p2p_resident_load(p2p_ctx* c, const double* pos, int64_t pos_stride, const double* vel, int64_t vel_stride, int64_t n, int64_t first_id) {
    USE(c);
    if (n < 1 || first_id < 0 || first_id + n > 0x7fffffffLL || !pos || pos_stride < 3 || (vel && vel_stride < 3)) return fail(P2P_ERR_ARG, "bad arrays");
    p2p_dtree* t;
    int r = get_tree(c, &t);
    if (r) return r;
    cudaStream_t st = c->stream;
    t->valid = false; t->resident = 0; t->stepping = false;
    CU(c->stage.reserve((size_t)n * 24, st));
    for (int k = 0; k < 3; k++) { CU(t->x[k].reserve((size_t)n, st)); CU(t->v[k].reserve((size_t)n, st)); CU(t->vtmp[k].reserve((size_t)n, st)); }
    CU(t->perm.reserve((size_t)n, st)); CU(t->seg.reserve((size_t)n, st)); CU(t->gid.reserve((size_t)n, st)); CU(t->gidtmp.reserve((size_t)n, st));
    if (vel) {
        CU(cudaMemcpy2DAsync(c->stage.p, 24, vel, (size_t)vel_stride * 8, 24, (size_t)n, cudaMemcpyHostToDevice, st));
        p2p::dt::soa_from_aos_kernel<<<blocks(n, 256), 256, 0, st>>>(reinterpret_cast<const double*>(c->stage.p), n, t->v[0].p, t->v[1].p,
                                                                     t->v[2].p, t->perm.p, t->seg.p);
    } else {
        for (int k = 0; k < 3; k++) CU(cudaMemsetAsync(t->v[k].p, 0, (size_t)n * 8, st));
    }
    CU(cudaMemcpy2DAsync(c->stage.p, 24, pos, (size_t)pos_stride * 8, 24, (size_t)n, cudaMemcpyHostToDevice, st));
    p2p::dt::soa_from_aos_kernel<<<blocks(n, 256), 256, 0, st>>>(reinterpret_cast<const double*>(c->stage.p), n, t->x[0].p, t->x[1].p, t->x[2].p,
                                                                 t->perm.p, t->seg.p);
    p2p::dt::iota_offset_kernel<<<blocks(n, 256), 256, 0, st>>>(t->gid.p, n, (int)first_id);
    CU(cudaGetLastError());
    t->resident = n; t->perm_is_local = false; t->stepping = true;
    return 0;
}

// tree of the resident particles, built from their current order (the reference too rebuilds from the order the previous
// step left, fmm_construct after the drift); velocities and ids follow the build's permutation
int p2p_resident_build(p2p_ctx* c, int maxleaf, const double bdl[3], const double bdr[3], int direct_start) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->stepping || t->resident < 1) return fail(P2P_ERR_STATE, "p2p_resident_load first");
    if (maxleaf < 1 || maxleaf > P2P_MAX_LEAF || !bdl || !bdr || direct_start < 0 || direct_start > 2) return fail(P2P_ERR_ARG, "bad arguments");
    if (!c->box_set) return fail(P2P_ERR_STATE, "p2p_set_box must precede the tree build");
    cudaStream_t st = c->stream;
    const long long n = t->resident;
    t->valid = false;
    CU(cudaEventRecord(t->e0, st));
    p2p::dt::iota_kernel<<<blocks(n, 256), 256, 0, st>>>(t->perm.p, n);
    int r = build_core(c, t, n, maxleaf, bdl, bdr, direct_start);
    if (r) return r;
    return resident_carry(c, t);
}

// short-range forces of the resident particles of ONE rank: build, walk, packing, forces and -- when the M2L lists are
// enabled (p2p_midfield_enable) -- the mid-field, which the kick then includes
int p2p_resident_forces(p2p_ctx* c, int maxleaf, const double bdl[3], const double bdr[3], int direct_start, double theta, double rcut,
                        double period) {
    int r = p2p_resident_build(c, maxleaf, bdl, bdr, direct_start);
    if (r) return r;
    p2p_dtree* t = c->dtree;
    double tc[3], tw[3];
    for (int k = 0; k < 3; k++) { tc[k] = 0.5 * (bdr[k] + bdl[k]); tw[k] = bdr[k] - bdl[k]; }
    if ((r = p2p_forces_local(c, theta, rcut, period, tc, tw, 1))) return r;
    if (t->m2l_on) return p2p_midfield_compute(c, nullptr);
    return 0;
}

// vel += (P2P + mid-field) * dkh for the resident particles (forces of the last step).  With the M2L lists enabled the
// mid-field must have been computed (p2p_resident_forces does; a multi-rank step calls p2p_midfield_compute_peers_packed):
// kicking with the P2P part alone would silently drop the pairs the walk handed to the expansions.
int p2p_resident_kick(p2p_ctx* c, double dkh) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->stepping || !t->valid) return fail(P2P_ERR_STATE, "no forces computed for the resident particles");
    if (t->m2l_on && !t->mid_valid) return fail(P2P_ERR_STATE, "the M2L lists are enabled but the mid-field of this step was not computed");
    if (t->mid_valid)
        p2p::dt::kick_mid_kernel<<<blocks(t->resident, 256), 256, 0, c->stream>>>(c->acc.p, t->acc_mid.p, t->resident, dkh, t->v[0].p, t->v[1].p, t->v[2].p);
    else
        p2p::dt::kick_kernel<<<blocks(t->resident, 256), 256, 0, c->stream>>>(c->acc.p, t->resident, dkh, t->v[0].p, t->v[1].p, t->v[2].p);
    CU(cudaGetLastError());
    return 0;
}

// ---- multi-rank resident stepping: after the drift the particles migrate to the ranks that own them, WITH their
// velocities and ids (domain_decomposition every step, 1_Indexing/src/domains.c:298-377) ---------------------------------
// partition of the resident particles by the rank kd-tree; velocities and ids follow
int p2p_resident_partition(p2p_ctx* c, int nproc, const double* split, int* sendcount) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->stepping || t->resident < 1) return fail(P2P_ERR_STATE, "p2p_resident_load first");
    p2p::dt::iota_kernel<<<blocks(t->resident, 256), 256, 0, c->stream>>>(t->perm.p, t->resident);
    int r = p2p_route_partition(c, nproc, split, sendcount);
    if (r) return r;
    return resident_carry(c, t);
}
// copies of / into the resident state; DEVICE pointers of the caller: 6 double arrays (x, y, z, vx, vy, vz), ids int32
int p2p_resident_export(p2p_ctx* c, void* const d_xv[6], void* d_id) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->stepping) return fail(P2P_ERR_STATE, "p2p_resident_load first");
    const size_t n = (size_t)t->resident;
    if (n == 0) return 0;
    for (int k = 0; k < 3; k++) {
        CU(cudaMemcpyAsync(d_xv[k], t->x[k].p, n * 8, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(d_xv[3 + k], t->v[k].p, n * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    CU(cudaMemcpyAsync(d_id, t->gid.p, n * 4, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}
int p2p_resident_import(p2p_ctx* c, const void* const d_xv[6], const void* d_id, int64_t n) {
    USE(c);
    if (n < 1 || n > 0x7fffffffLL || !d_xv || !d_id) return fail(P2P_ERR_ARG, "bad arrays");
    p2p_dtree* t;
    int r = get_tree(c, &t);
    if (r) return r;
    cudaStream_t st = c->stream;
    t->valid = false; t->resident = 0;
    for (int k = 0; k < 3; k++) {
        CU(t->x[k].reserve((size_t)n, st)); CU(t->v[k].reserve((size_t)n, st)); CU(t->vtmp[k].reserve((size_t)n, st));
        CU(cudaMemcpyAsync(t->x[k].p, d_xv[k], (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(t->v[k].p, d_xv[3 + k], (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    }
    CU(t->perm.reserve((size_t)n, st)); CU(t->seg.reserve((size_t)n, st)); CU(t->gid.reserve((size_t)n, st)); CU(t->gidtmp.reserve((size_t)n, st));
    CU(cudaMemcpyAsync(t->gid.p, d_id, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    t->resident = n; t->perm_is_local = false; t->stepping = true;
    return 0;
}

// pos += vel * dd, wrapped into [0, period) (period <= 0: no wrap); the tree is stale afterwards
int p2p_resident_drift(p2p_ctx* c, double dd, double period) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->stepping || t->resident < 1) return fail(P2P_ERR_STATE, "p2p_resident_load first");
    p2p::dt::drift_kernel<<<blocks(t->resident, 256), 256, 0, c->stream>>>(t->resident, dd, period, t->v[0].p, t->v[1].p, t->v[2].p, t->x[0].p,
                                                                           t->x[1].p, t->x[2].p);
    CU(cudaGetLastError());
    t->valid = false;
    return 0;
}

// current positions / velocities (packed rows of 3 doubles) and ids, in the resident order; NULL skips
int p2p_resident_download(p2p_ctx* c, double* pos, double* vel, int64_t* id) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->stepping || t->resident < 1) return fail(P2P_ERR_STATE, "p2p_resident_load first");
    const long long n = t->resident;
    cudaStream_t st = c->stream;
    CU(c->acc64.reserve((size_t)n * 3, st));
    if (pos) {
        p2p::dt::aos_from_soa_kernel<<<blocks(n, 256), 256, 0, st>>>(t->x[0].p, t->x[1].p, t->x[2].p, n, c->acc64.p);
        CU(cudaMemcpyAsync(pos, c->acc64.p, (size_t)n * 24, cudaMemcpyDeviceToHost, st));
    }
    if (vel) {
        p2p::dt::aos_from_soa_kernel<<<blocks(n, 256), 256, 0, st>>>(t->v[0].p, t->v[1].p, t->v[2].p, n, c->acc64.p);
        CU(cudaMemcpyAsync(vel, c->acc64.p, (size_t)n * 24, cudaMemcpyDeviceToHost, st));
    }
    if (id) {
        long long* tmp = reinterpret_cast<long long*>(c->acc64.p);
        p2p::dt::widen_index_kernel<<<blocks(n, 256), 256, 0, st>>>(t->gid.p, n, tmp);
        CU(cudaMemcpyAsync(id, tmp, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    return 0;
}

// ids of the particles in tree order (global ids after a routing, positions in the caller's array otherwise)
int p2p_download_index(p2p_ctx* c, int64_t* idx) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->valid || !t->built_here || !idx) return fail(P2P_ERR_STATE, "no device-built tree");
    // widened on the device and copied straight into the caller's buffer (full PCIe rate when that is pinned)
    CU(c->acc64.reserve((size_t)t->npart * 3, c->stream));
    long long* tmp = reinterpret_cast<long long*>(c->acc64.p);
    p2p::dt::widen_index_kernel<<<blocks(t->npart, 256), 256, 0, c->stream>>>(t->perm.p, t->npart, tmp);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(idx, tmp, (size_t)t->npart * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int p2p_tree_info(p2p_ctx* c, int* nleaf, int* nnode, int* nlevel, float* ms_build, float* ms_walk, int64_t* walk_items,
                  double* max_leaf_width) {
    if (!c || !c->dtree || !c->dtree->valid) return fail(P2P_ERR_STATE, "no device tree");
    if (nleaf) *nleaf = c->dtree->nleaf;
    if (nnode) *nnode = c->dtree->nnode;
    if (nlevel) *nlevel = c->dtree->nlevel;
    if (ms_build) *ms_build = c->dtree->ms_build;
    if (ms_walk) *ms_walk = c->dtree->ms_walk;
    if (walk_items) *walk_items = c->dtree->walk_items;
    if (max_leaf_width) *max_leaf_width = c->dtree->built_here ? c->dtree->max_leaf_width : -1.0;
    return 0;
}

// Copies of the device-built tree in the layout of p2p_tree_view / the reference (node_son: global ids with
// first_leaf = npart, first_node = npart + 2 npart / maxleaf); NULL pointers are skipped.
int p2p_tree_download(p2p_ctx* c, int64_t* perm, double* pos_sorted, int* leaf_npart, int* leaf_ipart, double* leaf_center,
                      double* leaf_width, int* node_npart, int* node_son, double* node_split, double* node_center, double* node_width) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->valid || !t->built_here) return fail(P2P_ERR_STATE, "no device-built tree");
    cudaStream_t st = c->stream;
    const long long np = t->npart;
    const int nl = t->nleaf, nn = t->nnode;
    if (perm) {
        std::vector<int> p((size_t)np);
        CU(cudaMemcpyAsync(p.data(), t->perm.p, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (long long i = 0; i < np; i++) perm[i] = p[(size_t)i];
    }
    if (pos_sorted) {
        std::vector<double> x((size_t)np);
        for (int k = 0; k < 3; k++) {
            CU(cudaMemcpyAsync(x.data(), t->x[k].p, (size_t)np * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (long long i = 0; i < np; i++) pos_sorted[3 * i + k] = x[(size_t)i];
        }
    }
    if (leaf_npart && nl) CU(cudaMemcpyAsync(leaf_npart, t->leaf_npart.p, (size_t)nl * 4, cudaMemcpyDeviceToHost, st));
    if (leaf_ipart && nl) CU(cudaMemcpyAsync(leaf_ipart, t->leaf_ipart.p, (size_t)nl * 4, cudaMemcpyDeviceToHost, st));
    if (node_npart) CU(cudaMemcpyAsync(node_npart, t->node_npart.p, (size_t)nn * 4, cudaMemcpyDeviceToHost, st));
    if (node_split) CU(cudaMemcpyAsync(node_split, t->node_split.p, (size_t)nn * 8, cudaMemcpyDeviceToHost, st));
    if (leaf_center || leaf_width || node_center || node_width) {
        std::vector<double> box(6 * ((size_t)nl + nn));
        CU(cudaMemcpyAsync(box.data(), t->box.p, box.size() * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (int l = 0; l < nl; l++)
            for (int k = 0; k < 3; k++) {
                if (leaf_center) leaf_center[3 * (size_t)l + k] = box[6 * (size_t)l + k];
                if (leaf_width) leaf_width[3 * (size_t)l + k] = box[6 * (size_t)l + 3 + k];
            }
        for (int n = 0; n < nn; n++)
            for (int k = 0; k < 3; k++) {
                if (node_center) node_center[3 * (size_t)n + k] = box[6 * ((size_t)nl + n) + k];
                if (node_width) node_width[3 * (size_t)n + k] = box[6 * ((size_t)nl + n) + 3 + k];
            }
    }
    if (node_son) {
        std::vector<int> son(2 * (size_t)nn);
        CU(cudaMemcpyAsync(son.data(), t->son.p, son.size() * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const int first_leaf = (int)np, first_node = (int)np + t->cap;
        for (size_t i = 0; i < son.size(); i++) node_son[i] = son[i] < 0 ? -1 : (son[i] < nl ? first_leaf + son[i] : first_node + (son[i] - nl));
    }
    CU(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"

namespace {
// breadth-first walk from the given items; appends the tasks to the context's list
int walk_impl(p2p_ctx* c, p2p_dtree* t, p2p::dt::WalkParams P, const std::vector<ull>& init, size_t task_guess) {
    cudaStream_t st = c->stream;
    constexpr int NC = p2p::dt::kWalkCounters;
    // the M2L list ACCUMULATES over the walks of one tree (target chunks, local then remote phase); a tree build empties it
    const size_t m0 = (size_t)t->nm2l;
    t->mid_valid = false; t->walk_period = P.period;
    size_t mcap = m0 + std::max<size_t>(task_guess / 8, 1 << 16);
    size_t fcap = std::max<size_t>(std::max<size_t>(task_guess / 2, 1 << 16), init.size());
    size_t tcap = std::max<size_t>(task_guess, 1 << 16);
    CU(cudaEventRecord(t->e0, st));
    const int batch0 = 2 * (t->built_here ? t->nlevel : 24) + 8;      // the walk descends one tree level per step on either side
    ull* h = t->h_wcount;
    for (int attempt = 0;; attempt++) {
        if (attempt > 8) return fail(P2P_ERR_CUDA, "dual-tree walk: buffers still overflow after %d attempts", attempt);
        if (t->m2l_on) { CU(t->mt.reserve(mcap, st, m0)); CU(t->ms.reserve(mcap, st, m0)); CU(t->mq.reserve(mcap, st, m0)); }
        CU(t->frontier[0].reserve(fcap, st)); CU(t->frontier[1].reserve(fcap, st));
        CU(c->tt.reserve((size_t)c->ntask + tcap, st, (size_t)c->ntask));
        CU(c->ts.reserve((size_t)c->ntask + tcap, st, (size_t)c->ntask));
        const ull cap_f = std::min(t->frontier[0].cap, t->frontier[1].cap);
        const ull cap_task = std::min(c->tt.cap, c->ts.cap) - (size_t)c->ntask;
        P.mt = t->m2l_on ? t->mt.p + m0 : nullptr; P.ms = t->ms.p + m0; P.mq = t->mq.p + m0;
        P.cap_m2l = t->m2l_on ? std::min(t->mt.cap, std::min(t->ms.cap, t->mq.cap)) - m0 : 0;
        CU(cudaMemcpyAsync(t->frontier[0].p, init.data(), init.size() * 8, cudaMemcpyHostToDevice, st));
        for (int k = 0; k < NC; k++) h[k] = 0;
        h[0] = init.size();
        CU(cudaMemcpyAsync(t->d_wcount, h, NC * sizeof(ull), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));                     // h is reused for the read-back below
        int level = 0;
        bool done = false, overflow = false;
        while (!done && !overflow) {
            if (level > 4096) return fail(P2P_ERR_CUDA, "dual-tree walk did not terminate");
            const int nb = level == 0 ? batch0 : 8;
            for (int k = 0; k < nb; k++, level++)
                p2p::dt::walk_level_kernel<<<c->num_sm * 8, 256, 0, st>>>(t->frontier[level & 1].p, cap_f, t->frontier[(level + 1) & 1].p, cap_f, t->d_wcount,
                                                                          c->tt.p + c->ntask, c->ts.p + c->ntask, cap_task, level, P);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(h, t->d_wcount, NC * sizeof(ull), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            overflow = h[6] != 0;
            done = h[level % 3] == 0;
        }
        t->walk_levels = level;
        if (!overflow) break;
        // a buffer was too small somewhere: grow what overflowed (the counters kept counting past the capacities) and redo
        if (h[6] & 1) fcap = std::max<size_t>(2 * fcap, (size_t)std::max(h[0], std::max(h[1], h[2])) + (1 << 16));
        if (h[6] & 2) tcap = std::max<size_t>(tcap + tcap / 2, (size_t)h[3] + (size_t)h[3] / 8);
        if (h[6] & 4) mcap = m0 + std::max<size_t>(2 * (mcap - m0), (size_t)h[4] + (size_t)h[4] / 8);
    }
    const ull ntask = h[3], nm2l = h[4], nviol = h[5], items = h[7];
    CU(cudaEventRecord(t->e1, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaEventElapsedTime(&t->ms_walk, t->e0, t->e1));
    c->ntask += (long long)ntask;
    c->csr_valid = false;
    t->walk_tasks = (long long)ntask; t->walk_items = (long long)items;
    t->nm2l = (long long)(m0 + nm2l);
    if (nviol)
        return fail(P2P_ERR_ARG, "%llu listed leaf pair(s) span half the period or more: the periodic box is too small for minimal-image "
                    "sources (needs roughly box > 2 (r_cut + 2 leaf widths))", nviol);
    return 0;
}

// tight bounds of the local leaves' particles (for the minimal-image check of a periodic walk)
int leaf_bounds(p2p_ctx* c, p2p_dtree* t) {
    CU(t->tb.reserve(6 * (size_t)std::max(c->nleaf, 1), c->stream));
    if (c->nleaf) {
        p2p::dt::leaf_bounds_kernel<<<blocks(c->nleaf, 128), 128, 0, c->stream>>>(c->leaf.p, c->nleaf, c->part.p, c->origin[0], c->origin[1],
                                                                               c->origin[2], c->extent / 4294967296.0, t->tb.p);
        CU(cudaGetLastError());
    }
    return 0;
}

int walk_checks(p2p_ctx* c, double theta, double rcut, double period, const double* tcenter, const double* twidth) {
    p2p_dtree* t = c->dtree;
    if (!t || !t->valid) return fail(P2P_ERR_STATE, "the device walk needs p2p_tree_build or p2p_tree_upload");
    if (t->nleaf != c->nleaf) return fail(P2P_ERR_STATE, "device tree and uploaded leaves differ");
    if (!(theta > 0.0) || !(rcut > 0.0)) return fail(P2P_ERR_ARG, "bad theta / rcut");
    if (period > 0.0 && (!tcenter || !twidth)) return fail(P2P_ERR_ARG, "periodic walk needs the target box");
    if ((long long)t->nleaf + t->nnode >= (1LL << 27)) return fail(P2P_ERR_ARG, "tree too large for the walk item encoding");
    return 0;
}
}  // namespace

extern "C" {

// walk_task_p2p over the device tree, plus (period > 0) the walks against the 26 periodic images of the same tree
// pruned against the target box {tcenter, twidth} exactly as prepare_sendtree2 / walk_task_p2p_ext do.  Image sources
// are listed under their LOCAL leaf id: the fixed-point coordinates wrap to the nearest image.  Tasks are appended to
// the context's list (p2p_build_csr packs them).
int p2p_tree_walk(p2p_ctx* c, double theta, double rcut, double period, const double tcenter[3], const double twidth[3]) {
    return p2p_tree_walk_range(c, theta, rcut, period, tcenter, twidth, 0, 0);
}

// the same walk restricted to the target leaves [leaf_lo, leaf_hi) (leaf_hi <= 0: all of them)
int p2p_tree_walk_range(p2p_ctx* c, double theta, double rcut, double period, const double tcenter[3], const double twidth[3],
                        int leaf_lo, int leaf_hi) {
    USE(c);
    int r = walk_checks(c, theta, rcut, period, tcenter, twidth);
    if (r) return r;
    p2p_dtree* t = c->dtree;
    const int nleaf = t->nleaf;
    if (nleaf == 0) return 0;
    if (leaf_hi > 0 && (!t->built_here || leaf_lo < 0 || leaf_lo >= leaf_hi || leaf_hi > nleaf))
        return fail(P2P_ERR_ARG, "bad target range [%d, %d) (needs a device-built tree with %d leaves)", leaf_lo, leaf_hi, nleaf);
    if (period > 0.0 && t->npart <= t->maxleaf)
        return fail(P2P_ERR_ARG, "image walk of a tree whose root holds <= maxleaf particles is undefined in the reference");
    p2p::dt::WalkParams P;
    memset(&P, 0, sizeof P);
    P.box = t->box.p; P.son = t->son.p; P.nleaf = nleaf; P.theta = theta; P.rcut = rcut; P.period = period > 0.0 ? period : 0.0;
    for (int k = 0; k < 3; k++) { P.tc[k] = tcenter ? tcenter[k] : 0.0; P.tw[k] = twidth ? twidth[k] : 0.0; }
    const int me = t->self_rank;              // 0 unless p2p_set_rank said otherwise: the peer index the M2L tasks carry
    P.sbox = t->box.p; P.sson = t->son.p; P.me = me; P.npeer = me + 1; P.snleaf[me] = nleaf;
    if (period > 0.0) {
        if ((r = leaf_bounds(c, t))) return r;            // (also what p2p_tree_export_packed ships)
        P.tb = t->tb.p; P.stb = t->tb.p;
        // A leaf pair enters the frontier from a pair that passed the cutoff test (cell gap < r_cut) with a node that has a leaf
        // child in the leaf's place, so no listed particle separation exceeds r_cut + 2 (largest such node).  If that is below
        // half the period the per-pair check of the tight bounds -- two 48-byte gathers for every one of the 1e8 tasks of 256^3 --
        // cannot fire and is skipped; small boxes (the 32-cell demo after a few steps) keep the exact per-pair check.
        if (t->built_here && t->max_parent_width >= 0.0 && rcut + 2.0 * t->max_parent_width * (1.0 + 1e-6) < 0.5 * period) P.tb = nullptr;
    }
    if (leaf_hi > 0) { P.node_leaf0 = t->node_leaf0.p; P.node_nleaf = t->node_nleaf.p; P.t_lo = leaf_lo; P.t_hi = leaf_hi; }
    std::vector<ull> init;
    const int root = nleaf;
    init.push_back(p2p::dt::item(root, root, me, 0));
    if (period > 0.0) for (int s = 1; s < 27; s++) init.push_back(p2p::dt::item(root, root, me, s));
    if (t->literal_d6) init.push_back(p2p::dt::item(root, root, me, 27));    // the reference's zero-shift self exchange (defect D6)
    const size_t rows = leaf_hi > 0 ? (size_t)(leaf_hi - leaf_lo) : (size_t)nleaf;
    return walk_impl(c, t, P, init, rows * (period > 0.0 ? 192 : 160) * (t->literal_d6 ? 2 : 1));
}

// ---- multi-rank: one tree per rank, every rank walks its tree against all of them; ONE topology buffer per rank for ONE all-gather ---------------------------------------------------------
// (block layout: p2p_topo_layout in p2p_ctx.h)
int p2p_set_rank(p2p_ctx* c, int rank, int nranks) {
    USE(c);
    if (nranks < 1 || nranks > p2p::dt::kMaxPeers || rank < 0 || rank >= nranks) return fail(P2P_ERR_ARG, "bad rank (at most %d ranks)", p2p::dt::kMaxPeers);
    p2p_dtree* t;
    int r = get_tree(c, &t);
    if (r) return r;
    t->self_rank = rank; t->nranks = nranks;
    return 0;
}

int p2p_topology_stride(int nleaf_max, int nnode_max, int64_t* stride_bytes) {
    if (nleaf_max < 0 || nnode_max < 0 || !stride_bytes) return fail(P2P_ERR_ARG, "bad arguments");
    long long a, b, d, s;
    p2p_topo_layout(nleaf_max, nnode_max, &a, &b, &d, &s);
    *stride_bytes = s;
    return 0;
}

int p2p_tree_export_packed(p2p_ctx* c, void* d_block, int nleaf_max, int nnode_max) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->valid) return fail(P2P_ERR_STATE, "no device tree");
    if (!d_block || t->nleaf > nleaf_max || t->nnode > nnode_max) return fail(P2P_ERR_ARG, "tree larger than the block layout");
    long long off_tb, off_son, off_leaf, stride;
    p2p_topo_layout(nleaf_max, nnode_max, &off_tb, &off_son, &off_leaf, &stride);
    char* dst = reinterpret_cast<char*>(d_block);
    const size_t nu = (size_t)t->nleaf + t->nnode;
    CU(cudaMemcpyAsync(dst, t->box.p, nu * 48, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(dst + off_son, t->son.p, (size_t)t->nnode * 8, cudaMemcpyDeviceToDevice, c->stream));
    if (t->nleaf) {
        int r = leaf_bounds(c, t);
        if (r) return r;
        CU(cudaMemcpyAsync(dst + off_tb, t->tb.p, (size_t)t->nleaf * 48, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(dst + off_leaf, c->leaf.p, (size_t)t->nleaf * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    return 0;
}

// p2p_tree_walk_peers over the P gathered blocks.  include_me = 0 skips this rank's own tree (its walk, images included,
// is p2p_tree_walk, which can run -- and feed the force kernel -- before any topology has arrived).
int p2p_tree_walk_peers_packed(p2p_ctx* c, double theta, double rcut, double period, const double tcenter[3], const double twidth[3],
                               int npeer, int me, const int* peer_nleaf, const int* peer_nnode, const void* d_all, int nleaf_max,
                               int nnode_max, int include_me) {
    USE(c);
    int r = walk_checks(c, theta, rcut, period, tcenter, twidth);
    if (r) return r;
    p2p_dtree* t = c->dtree;
    if (npeer < 1 || npeer > p2p::dt::kMaxPeers || me < 0 || me >= npeer || !peer_nleaf || !peer_nnode || !d_all || !tcenter || !twidth)
        return fail(P2P_ERR_ARG, "bad peer arrays (at most %d ranks)", p2p::dt::kMaxPeers);
    if (peer_nleaf[me] != t->nleaf || peer_nnode[me] != t->nnode) return fail(P2P_ERR_ARG, "peer %d is not this rank's tree", me);
    long long off_tb, off_son, off_leaf, stride;
    p2p_topo_layout(nleaf_max, nnode_max, &off_tb, &off_son, &off_leaf, &stride);
    p2p::dt::WalkParams P;
    memset(&P, 0, sizeof P);
    P.box = t->box.p; P.son = t->son.p; P.nleaf = t->nleaf; P.theta = theta; P.rcut = rcut; P.period = period > 0.0 ? period : 0.0;
    for (int k = 0; k < 3; k++) { P.tc[k] = tcenter[k]; P.tw[k] = twidth[k]; }
    P.sbox = reinterpret_cast<const double*>(d_all); P.sson = reinterpret_cast<const int*>(d_all); P.me = me; P.npeer = npeer;
    if (period > 0.0) {
        if ((r = leaf_bounds(c, t))) return r;
        P.tb = t->tb.p; P.stb = reinterpret_cast<const double*>(d_all);
    }
    int ghost = t->nleaf;
    std::vector<ull> init;
    for (int p = 0; p < npeer; p++) {
        if (peer_nleaf[p] < 0 || peer_nnode[p] < 1 || peer_nleaf[p] > nleaf_max || peer_nnode[p] > nnode_max)
            return fail(P2P_ERR_ARG, "peer %d has no tree or exceeds the block layout", p);
        if ((long long)peer_nleaf[p] + peer_nnode[p] >= (1LL << 27)) return fail(P2P_ERR_ARG, "peer tree too large for the walk item encoding");
        P.sbox_base[p] = (long long)p * stride / 48;
        P.stb_base[p] = ((long long)p * stride + off_tb) / 48;
        P.sson_base[p] = ((long long)p * stride + off_son) / 8;
        P.snleaf[p] = peer_nleaf[p];
        P.ts_base[p] = p == me ? 0 : ghost;
        if (p != me) ghost += peer_nleaf[p];
        if (p == me && !include_me) continue;
        for (int s = 0; s < (period > 0.0 ? 27 : 1); s++) init.push_back(p2p::dt::item(t->nleaf, peer_nleaf[p], p, s));
        if (p == me && t->literal_d6) init.push_back(p2p::dt::item(t->nleaf, peer_nleaf[p], p, 27));   // zero-shift self exchange (defect D6)
    }
    c->nghostleaf = ghost - t->nleaf;          // the ghost leaf table follows (p2p_halo_plan_need) before the list is packed
    c->nghost = 0;
    c->bounds_n = std::min(c->bounds_n, c->nleaf);
    if (init.empty()) return 0;
    return walk_impl(c, t, P, init, (size_t)t->nleaf * (include_me ? 224 : 32));
}

// number of (row, source) duplicates in the packed list: non-zero means two images of one leaf reached the same
// target, i.e. the periodic box is too small for minimal-image sources (needs box > 2 (r_cut + leaf sizes))
int p2p_csr_duplicates(p2p_ctx* c, int64_t* ndup) {
    USE(c);
    if (!c->csr_valid) return fail(P2P_ERR_STATE, "no CSR built");
    p2p_dtree* t;
    int r = get_tree(c, &t);
    if (r) return r;
    CU(cudaMemsetAsync(t->d_dup, 0, sizeof(unsigned int), c->stream));
    if (c->nleaf) p2p::dt::csr_duplicate_kernel<<<blocks(c->nleaf, 256), 256, 0, c->stream>>>(c->row_ptr.p, c->col.p, c->nleaf, t->d_dup);
    CU(cudaGetLastError());
    unsigned int d = 0;
    CU(cudaMemcpyAsync(&d, t->d_dup, sizeof d, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (ndup) *ndup = d;
    return 0;
}

// accelerations back in the ORDER OF THE POSITIONS GIVEN TO p2p_tree_build (packed rows of 3 doubles)
int p2p_download_acc_original(p2p_ctx* c, double* acc) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->valid || !t->built_here) return fail(P2P_ERR_STATE, "no device-built tree");
    if (!t->perm_is_local) return fail(P2P_ERR_STATE, "particles were routed between ranks: use p2p_download_acc + p2p_download_index");
    if (!acc) return fail(P2P_ERR_ARG, "null acc");
    const long long n = c->npart;
    CU(c->acc64.reserve((size_t)n * 3, c->stream));
    if (t->mid_valid)
        p2p::mf::acc_unpermute_sum_kernel<<<blocks(n, 256), 256, 0, c->stream>>>(c->acc.p, t->acc_mid.p, t->perm.p, n, c->acc64.p);
    else
        p2p::dt::acc_unpermute_kernel<<<blocks(n, 256), 256, 0, c->stream>>>(c->acc.p, t->perm.p, n, c->acc64.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(acc, c->acc64.p, (size_t)n * 24, cudaMemcpyDeviceToHost, c->stream));
    return p2p_synchronize(c);
}

// ---- mid-field: M2L lists come from the same walk; P2M/M2M/M2L/L2L/L2P as kernels ---------------------
// on != 0: the next walks also emit the M2L task list (pairs the acceptance criterion hands to the expansions).
// literal_d6 != 0 (test knob): additionally walk the local tree against itself with the remote rules, as the reference's
// zero-shift self exchange does (SURVEY defect D6) -- every local P2P and M2L task then appears twice, as in the reference.
int p2p_midfield_enable(p2p_ctx* c, int on, int literal_d6) {
    USE(c);
    p2p_dtree* t;
    int r = get_tree(c, &t);
    if (r) return r;
    t->m2l_on = on != 0; t->literal_d6 = literal_d6 != 0; t->mid_valid = false;
    return 0;
}

}  // extern "C"

namespace {
// P2M + M2M of the local tree
int mid_upward(p2p_ctx* c, p2p_dtree* t) {
    cudaStream_t st = c->stream;
    const int nl = t->nleaf;
    const size_t nu = (size_t)nl + t->nnode;
    CU(t->Mall.reserve(nu * p2p::mf::NM, st));
    if (nl) p2p::mf::p2m_kernel<<<blocks(nl, 128), 128, 0, st>>>(c->leaf.p, nl, t->box.p, t->x[0].p, t->x[1].p, t->x[2].p, c->mass, t->Mall.p);
    for (int lvl = t->nlevel - 1; lvl >= 0; lvl--)
        p2p::mf::m2m_level_kernel<<<blocks(t->lvl_count[lvl], 128), 128, 0, st>>>(t->t_id.p, t->lvl_begin[lvl], t->lvl_count[lvl], t->son.p, nl,
                                                                                  t->box.p, t->Mall.p);
    CU(cudaGetLastError());
    return 0;
}
// M2L of the last walk's list with the given source trees, then L2L and L2P
int mid_m2l_down(p2p_ctx* c, p2p_dtree* t, const double* sbox, const double* sM, const long long* sbase, int npeer,
                 const long long* sbase_M = nullptr) {
    cudaStream_t st = c->stream;
    const int nl = t->nleaf;
    const size_t nu = (size_t)nl + t->nnode;
    CU(t->Lall.reserve(nu * p2p::mf::NM, st)); CU(t->acc_mid.reserve((size_t)t->npart * 3, st));
    CU(cudaMemsetAsync(t->Lall.p, 0, nu * p2p::mf::NM * sizeof(double), st));
    if (t->nm2l) {
        p2p::mf::M2LParams P;
        memset(&P, 0, sizeof P);
        P.mt = t->mt.p; P.ms = t->ms.p; P.mq = t->mq.p; P.ntask = t->nm2l; P.box = t->box.p; P.sbox = sbox; P.sM = sM;
        for (int p = 0; p < npeer; p++) { P.sbase[p] = sbase[p]; P.sbase_M[p] = sbase_M ? sbase_M[p] : sbase[p]; }
        P.period = t->walk_period; P.rs = c->rs; P.L = t->Lall.p;
        p2p::mf::m2l_kernel<<<blocks(t->nm2l, 128), 128, 0, st>>>(P);
    }
    for (int lvl = 0; lvl < t->nlevel; lvl++)
        p2p::mf::l2l_level_kernel<<<blocks(t->lvl_count[lvl], 128), 128, 0, st>>>(t->t_id.p, t->lvl_begin[lvl], t->lvl_count[lvl], t->son.p, nl,
                                                                                  t->box.p, t->Lall.p);
    CU(cudaMemsetAsync(t->acc_mid.p, 0, (size_t)t->npart * 24, st));
    if (nl) p2p::mf::l2p_kernel<<<blocks((long long)nl * 32, 128), 128, 0, st>>>(c->leaf.p, nl, t->box.p, t->Lall.p, t->x[0].p, t->x[1].p, t->x[2].p,
                                                                                 t->acc_mid.p);
    CU(cudaGetLastError());
    return 0;
}
}  // namespace

extern "C" {

// P2M -> M2M -> M2L (tasks of the last walk) -> L2L -> L2P on the device-built tree; the result (fp64, per particle) is
// added by p2p_download_acc_original.  Single rank (sources are the local tree and its periodic images).
int p2p_midfield_compute(p2p_ctx* c, int64_t* nm2l) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->valid || !t->built_here) return fail(P2P_ERR_STATE, "the mid-field needs a device-built tree (p2p_tree_build)");
    if (!t->m2l_on) return fail(P2P_ERR_STATE, "enable the M2L lists (p2p_midfield_enable) before the walk");
    cudaStream_t st = c->stream;
    CU(cudaEventRecord(t->e0, st));
    int r = mid_upward(c, t);
    if (r) return r;
    long long zero[p2p::dt::kMaxPeers] = {0};          // the M2L tasks carry peer index self_rank: every slot is the local tree
    if ((r = mid_m2l_down(c, t, t->box.p, t->Mall.p, zero, p2p::dt::kMaxPeers))) return r;
    CU(cudaEventRecord(t->e1, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaEventElapsedTime(&t->ms_mid, t->e0, t->e1));
    t->mid_valid = true;
    if (nm2l) *nm2l = t->nm2l;
    return 0;
}

// multi-rank: the local multipoles (P2M + M2M), copied to d_M [(nleaf + nnode)][20] (device) for an all-gather ...
int p2p_midfield_multipoles(p2p_ctx* c, void* d_M) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->valid || !t->built_here) return fail(P2P_ERR_STATE, "the mid-field needs a device-built tree");
    int r = mid_upward(c, t);
    if (r) return r;
    if (d_M) CU(cudaMemcpyAsync(d_M, t->Mall.p, ((size_t)t->nleaf + t->nnode) * p2p::mf::NM * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}
// ... and M2L -> L2L -> L2P with the boxes (packed topology blocks, p2p_tree_export_packed) and multipoles (blocks of
// (nleaf_max + nnode_max) x 20 doubles) of ALL ranks; the M2L list is the one the local and the remote walk left
int p2p_midfield_compute_peers_packed(p2p_ctx* c, int npeer, const void* d_topo_all, int nleaf_max, int nnode_max, const void* d_M_all,
                                      int64_t* nm2l) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->valid || !t->built_here) return fail(P2P_ERR_STATE, "the mid-field needs a device-built tree");
    if (!t->m2l_on) return fail(P2P_ERR_STATE, "enable the M2L lists (p2p_midfield_enable) before the walk");
    if (npeer < 1 || npeer > p2p::dt::kMaxPeers || !d_topo_all || !d_M_all) return fail(P2P_ERR_ARG, "bad peer arrays");
    long long off_tb, off_son, off_leaf, stride, sbase[p2p::dt::kMaxPeers], sbase_M[p2p::dt::kMaxPeers];
    p2p_topo_layout(nleaf_max, nnode_max, &off_tb, &off_son, &off_leaf, &stride);
    for (int p = 0; p < npeer; p++) { sbase[p] = (long long)p * stride / 48; sbase_M[p] = (long long)p * ((long long)nleaf_max + nnode_max); }
    cudaStream_t st = c->stream;
    CU(cudaEventRecord(t->e0, st));
    int r = mid_m2l_down(c, t, reinterpret_cast<const double*>(d_topo_all), reinterpret_cast<const double*>(d_M_all), sbase, npeer, sbase_M);
    if (r) return r;
    CU(cudaEventRecord(t->e1, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaEventElapsedTime(&t->ms_mid, t->e0, t->e1));
    t->mid_valid = true;
    if (nm2l) *nm2l = t->nm2l;
    return 0;
}

// multipoles / local expansions of the device tree in the reference's layout ([leaf][20], [node][20]); NULL skips
int p2p_midfield_download(p2p_ctx* c, double* leaf_M, double* node_M, double* leaf_L, double* node_L, float* ms) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->mid_valid) return fail(P2P_ERR_STATE, "no mid-field computed");
    const size_t nl = (size_t)t->nleaf, nn = (size_t)t->nnode, W = p2p::mf::NM * sizeof(double);
    if (leaf_M && nl) CU(cudaMemcpyAsync(leaf_M, t->Mall.p, nl * W, cudaMemcpyDeviceToHost, c->stream));
    if (node_M) CU(cudaMemcpyAsync(node_M, t->Mall.p + nl * p2p::mf::NM, nn * W, cudaMemcpyDeviceToHost, c->stream));
    if (leaf_L && nl) CU(cudaMemcpyAsync(leaf_L, t->Lall.p, nl * W, cudaMemcpyDeviceToHost, c->stream));
    if (node_L) CU(cudaMemcpyAsync(node_L, t->Lall.p + nl * p2p::mf::NM, nn * W, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (ms) *ms = t->ms_mid;
    return 0;
}

// Local lists and forces of the device-built tree: walk (own tree and, with period > 0, its 26 images), packing and the
// force kernel, in TARGET CHUNKS (ranges of leaves) sized so that one chunk's list stays below max_chunk_tasks
// (p2p_set_chunk_tasks): the task list of 1024^3 particles on one GPU would take 6.7e9 x 12 bytes.  The accelerations
// accumulate; p2p_accumulated_counts gives the totals, p2p_step_timings the device times summed over the chunks.
}  // extern "C"

namespace {
// The chunks of p2p_forces_local as a two-stage pipeline: walk and packing of chunk k + 1 run on a second, high-priority
// stream into the other list set while the force kernel of chunk k runs on the caller's stream with retiring warps (so that
// the small latency-bound kernels of the walk and the packing find SM slots).  Replaces the ping-pong task buffers of the
// reference (1_Indexing/src/fmm.c:365-400, 947-1024: walk into one buffer while the other is computed -- serialised in this
// fork).  Every row still lives in exactly one chunk and is summed in the same order: results are bit-identical.
int forces_local_pipelined(p2p_ctx* c, p2p_dtree* t, int nchunk, double theta, double rcut, double period, const double tcenter[3],
                           const double twidth[3], int compute, cudaStream_t A, int user_budget) {
    if (!t->pipe_stream) {
        int least = 0, greatest = 0;
        CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CU(cudaStreamCreateWithPriority(&t->pipe_stream, cudaStreamNonBlocking, greatest));
        CU(cudaEventCreateWithFlags(&t->ev_pipe_built, cudaEventDisableTiming));
        for (int k = 0; k < 2; k++) {
            CU(cudaEventCreateWithFlags(&t->ev_pack_done[k], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&t->ev_force_done[k], cudaEventDisableTiming));
        }
    }
    cudaStream_t W = t->pipe_stream;
    const int nleaf = t->nleaf;
    CU(cudaEventRecord(t->ev_pipe_built, A));            // tree, particles, zeroed accelerations
    CU(cudaStreamWaitEvent(W, t->ev_pipe_built, 0));
    int r;
    for (int ch = 0; ch < nchunk; ch++) {
        const int lo = (int)((long long)nleaf * ch / nchunk), hi = (int)((long long)nleaf * (ch + 1) / nchunk);
        const int set = ch & 1;
        if (ch > 0 && (r = p2p_swap_lists(c))) return r;
        c->stream = W;
        if (ch >= 2) CU(cudaStreamWaitEvent(W, t->ev_force_done[set], 0));      // the kernel that consumed this set two chunks ago
        if ((r = p2p_clear_tasks(c))) return r;
        t->ms_walk = 0.f;
        if (hi > lo && (r = p2p_tree_walk_range(c, theta, rcut, period, tcenter, twidth, lo, hi))) return r;
        t->sum_walk += t->ms_walk;
        CU(cudaEventRecord(t->chunk_ev[4 * ch], W));
        c->row_lo = lo; c->row_hi = hi;
        r = hi > lo ? p2p_build_csr(c) : 0;
        c->row_lo = c->row_hi = 0;
        if (r) return r;
        CU(cudaEventRecord(t->chunk_ev[4 * ch + 1], W));
        CU(cudaEventRecord(t->ev_pack_done[set], W));
        c->stream = A;
        CU(cudaStreamWaitEvent(A, t->ev_pack_done[set], 0));
        CU(cudaEventRecord(t->chunk_ev[4 * ch + 3], A));
        c->rows_per_warp = ch + 1 < nchunk ? std::max(user_budget, 4) : user_budget;    // the last kernel has nothing to make room for
        if (compute && hi > lo && (r = p2p_compute(c))) return r;
        CU(cudaEventRecord(t->chunk_ev[4 * ch + 2], A));
        CU(cudaEventRecord(t->ev_force_done[set], A));
    }
    return 0;
}
}  // namespace

extern "C" {

int p2p_forces_local(p2p_ctx* c, double theta, double rcut, double period, const double tcenter[3], const double twidth[3], int compute) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t || !t->valid || !t->built_here) return fail(P2P_ERR_STATE, "p2p_forces_local needs a device-built tree");
    const int nleaf = t->nleaf;
    const long long est = (long long)nleaf * (period > 0.0 ? 200 : 170);
    int nchunk = (int)std::min<long long>((est + t->max_chunk_tasks - 1) / t->max_chunk_tasks, std::max(nleaf, 1));
    if (nchunk < 1) nchunk = 1;
    // single-rank steps of a useful size are cut into at least pipe_chunks chunks and pipelined
    const bool pipe = t->pipe_chunks > 1 && t->nranks <= 1 && compute && nleaf >= 4096 * t->pipe_chunks;
    if (pipe) nchunk = std::max(nchunk, t->pipe_chunks);
    while ((int)t->chunk_ev.size() < 4 * nchunk) { cudaEvent_t e; CU(cudaEventCreate(&e)); t->chunk_ev.push_back(e); }
    t->sum_walk = t->sum_csr = t->sum_force = 0.f;
    t->last_chunks = nchunk;
    int r;
    if (pipe) {
        cudaStream_t A = c->stream;
        const int user_budget = c->rows_per_warp;
        r = forces_local_pipelined(c, t, nchunk, theta, rcut, period, tcenter, twidth, compute, A, user_budget);
        c->stream = A; c->rows_per_warp = user_budget; c->row_lo = c->row_hi = 0;
        return r;
    }
    for (int ch = 0; ch < nchunk; ch++) {
        const int lo = (int)((long long)nleaf * ch / nchunk), hi = (int)((long long)nleaf * (ch + 1) / nchunk);
        if ((r = p2p_clear_tasks(c))) return r;
        if (nchunk == 1) r = p2p_tree_walk_range(c, theta, rcut, period, tcenter, twidth, 0, 0);
        else if (hi > lo) r = p2p_tree_walk_range(c, theta, rcut, period, tcenter, twidth, lo, hi);
        if (r) return r;
        t->sum_walk += t->ms_walk;
        CU(cudaEventRecord(t->chunk_ev[4 * ch], c->stream));
        if (nchunk > 1) { c->row_lo = lo; c->row_hi = hi; }      // the chunk's tasks have their targets in [lo, hi): pack those rows only
        r = hi > lo || nchunk == 1 ? p2p_build_csr(c) : 0;
        c->row_lo = c->row_hi = 0;
        if (r) return r;
        CU(cudaEventRecord(t->chunk_ev[4 * ch + 1], c->stream));
        CU(cudaEventRecord(t->chunk_ev[4 * ch + 3], c->stream));
        if (compute && (hi > lo || nchunk == 1) && (r = p2p_compute(c))) return r;
        CU(cudaEventRecord(t->chunk_ev[4 * ch + 2], c->stream));
    }
    return 0;
}

// chunks a single-rank p2p_forces_local is cut into at least, pipelined (walk + packing of the next chunk beside the force
// kernel of the current one); 0 or 1 = off
int p2p_set_chunk_pipeline(p2p_ctx* c, int min_chunks) {
    USE(c);
    p2p_dtree* t;
    int r = get_tree(c, &t);
    if (r) return r;
    if (min_chunks < 0 || min_chunks > 64) return fail(P2P_ERR_ARG, "bad chunk count");
    t->pipe_chunks = min_chunks;
    return 0;
}

int p2p_set_chunk_tasks(p2p_ctx* c, int64_t max_tasks) {
    USE(c);
    p2p_dtree* t;
    int r = get_tree(c, &t);
    if (r) return r;
    if (max_tasks < 1024) return fail(P2P_ERR_ARG, "chunk size too small");
    t->max_chunk_tasks = max_tasks;
    return 0;
}

// device times of the last step: tree build, walks, packing, force kernels (summed over the chunks of p2p_forces_local);
// synchronises the stream
int p2p_step_timings(p2p_ctx* c, float* ms_build, float* ms_walk, float* ms_csr, float* ms_force, int* nchunk) {
    USE(c);
    p2p_dtree* t = c->dtree;
    if (!t) return fail(P2P_ERR_STATE, "no device tree");
    CU(cudaStreamSynchronize(c->stream));
    float csr = 0.f, force = 0.f;
    for (int ch = 0; ch < t->last_chunks && 4 * ch + 2 < (int)t->chunk_ev.size(); ch++) {
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, t->chunk_ev[4 * ch], t->chunk_ev[4 * ch + 1]) == cudaSuccess) csr += a;
        if (cudaEventElapsedTime(&b, t->chunk_ev[4 * ch + 3], t->chunk_ev[4 * ch + 2]) == cudaSuccess) force += b;
    }
    cudaGetLastError();
    if (ms_build) *ms_build = t->ms_build;
    if (ms_walk) *ms_walk = t->sum_walk;
    if (ms_csr) *ms_csr = csr;
    if (ms_force) *ms_force = force;
    if (nchunk) *nchunk = t->last_chunks;
    return 0;
}

// The whole short-range P2P step of one rank in one call: positions (caller's order) in, accelerations (same order)
// out; tree build, walk (with the 26 periodic images when period > 0), packing and forces on the device.
int p2p_step_device(p2p_ctx* c, const double* pos, int64_t stride, int64_t npart, int maxleaf, const double bdl[3],
                    const double bdr[3], int direct_start, double theta, double rcut, double period, double* acc) {
    int r;
    if ((r = p2p_tree_build(c, pos, stride, npart, maxleaf, bdl, bdr, direct_start))) return r;
    double tc[3], tw[3];
    for (int k = 0; k < 3; k++) { tc[k] = 0.5 * (bdr[k] + bdl[k]); tw[k] = bdr[k] - bdl[k]; }   // the local root cell (toptree.c:18-45)
    if ((r = p2p_forces_local(c, theta, rcut, period, tc, tw, 1))) return r;
    return p2p_download_acc_original(c, acc);
}

}  // extern "C"
