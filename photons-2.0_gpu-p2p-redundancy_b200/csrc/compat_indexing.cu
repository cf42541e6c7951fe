// compat_indexing.cu -- the reference's Indexing GPU C-ABI (include/photoNs_CUDA_indexing.h) on top
// of the native library.  One process-wide context on device 0, like the reference
// (1_Indexing/src/photoNs_CUDA.cu:20-43); device buffers persist and grow (defect D16 fixed).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <vector>

#include "../../include/p2p_b200.h"
#include "../../include/photoNs_CUDA_indexing.h"

// link-compatibility with the globals the reference header declares extern (cuh:7-16)
double* d_particle_data = nullptr;
int* d_leaf_data = nullptr;
int* d_interaction_data = nullptr;
double* d_result_data = nullptr;
int max_particle_data = 0, max_leaf_data = 0, max_interaction_data = 0, max_result_data = 0;

namespace {
p2p_ctx* g_ctx = nullptr;
bool g_init_tried = false;
double g_rs = 0.0;
bool g_rs_from_env_checked = false;
int g_max_parts = 0;
int g_ntasks = 0, g_nleaf = 0;
std::vector<long long> g_first_task;   // first task index of every target leaf, -1 if none
std::vector<int> g_leaf_npart;
int g_state = 0;                       // status of the last copy/launch

void check_env() {
    if (g_rs_from_env_checked) return;
    g_rs_from_env_checked = true;
    const char* e = getenv("P2P_B200_RS");
    if (e && g_rs == 0.0) g_rs = atof(e);
}
}  // namespace

extern "C" {

void p2pSetSplitRadius(double rs) { g_rs = rs; g_rs_from_env_checked = true; }

void initGPU(int verbosity_gpu) {
    if (g_init_tried) return;
    g_init_tried = true;
    check_env();
    if (p2p_create(&g_ctx, 0) != 0) {
        printf("No CUDA-capable device found! (%s)\n", p2p_last_error());
        g_ctx = nullptr;
    } else if (verbosity_gpu) {
        printf(">> \tp2p_b200 context on device 0 (%s kernel)\n", g_rs > 0 ? "erfc-truncated" : "plain");
    }
}

void getGPUMemoryState(int verbosity_gpu) {
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { if (verbosity_gpu) printf("cudaMemGetInfo failed\n"); return; }
    if (verbosity_gpu) printf(">> \tFree memory: %zu bytes\n", fr);
}

int allocMemGPU(int nleafs, int maxPartsInLeaf, int maxNeighbors, int nTasks, int verbosity_gpu) {
    (void)nleafs; (void)maxNeighbors; (void)nTasks; (void)verbosity_gpu;
    initGPU(verbosity_gpu);
    if (!g_ctx) return -1;
    g_max_parts = maxPartsInLeaf;   // buffers themselves are sized on demand by the native library
    return 0;
}

int copyMemGPU(double* h_pos, int* h_leaf, int* h_int, int ntasks, int verbosity_gpu) {
    (void)verbosity_gpu;
    g_state = -1;
    if (!g_ctx) { printf("Error copy: no GPU context\n"); return -1; }
    if (g_max_parts <= 0 || ntasks < 0 || (ntasks && (!h_pos || !h_leaf || !h_int))) { printf("Error copy: bad arguments\n"); return -1; }
    check_env();
    // the leaf table's valid length is not part of the ABI: take the largest id the list uses
    int nleaf = 0;
    for (int n = 0; n < 2 * ntasks; n++) {
        if (h_int[n] < 0) { printf("ERROR : task %d references leaf %d (negative)\n", n / 2, h_int[n]); g_state = -3; return -3; }
        if (h_int[n] + 1 > nleaf) nleaf = h_int[n] + 1;
    }
    g_ntasks = ntasks; g_nleaf = nleaf;
    g_leaf_npart.assign((size_t)nleaf, 0);
    std::vector<int> ipart((size_t)nleaf);
    for (int l = 0; l < nleaf; l++) {
        int np = h_leaf[2 * l];
        if (np < 0 || np > g_max_parts) { printf("ERROR : leaf %d holds %d particles (max %d)\n", l, np, g_max_parts); g_state = -3; return -3; }
        g_leaf_npart[(size_t)l] = np;
        ipart[(size_t)l] = l * g_max_parts;          // position of the leaf's chunk in h_pos
    }
    g_first_task.assign((size_t)nleaf, -1);
    for (int n = ntasks - 1; n >= 0; n--) g_first_task[(size_t)h_int[2 * n]] = n;
    // the fixed-point frame is derived from the VALID slots only (leaf l, j < npart[l]): the padding of the chunked
    // array is never read by the reference kernel and may hold anything (bounding cube, four times padded)
    {
        double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
        bool first = true;
        for (int l = 0; l < nleaf; l++)
            for (int j = 0; j < g_leaf_npart[(size_t)l]; j++) {
                const double* q = h_pos + ((size_t)l * g_max_parts + j) * 3;
                for (int k = 0; k < 3; k++) {
                    if (first || q[k] < lo[k]) lo[k] = q[k];
                    if (first || q[k] > hi[k]) hi[k] = q[k];
                }
                first = false;
            }
        double w = 0.0;
        for (int k = 0; k < 3; k++) w = hi[k] - lo[k] > w ? hi[k] - lo[k] : w;
        if (!(w > 0.0) || !std::isfinite(w)) w = 1.0;
        const double extent = 4.0 * w;
        double origin[3];
        for (int k = 0; k < 3; k++) origin[k] = 0.5 * (lo[k] + hi[k]) - 0.5 * extent;
        if (p2p_set_box(g_ctx, origin, extent)) { printf("Error copy: %s\n", p2p_last_error()); return -1; }
    }
    // r_s fixes the position scale, so it is applied before the upload; mass/eps arrive at launch
    if (p2p_set_physics(g_ctx, 1.0, 0.0, g_rs)) { printf("Error copy: %s\n", p2p_last_error()); return -1; }
    if (p2p_upload_particles(g_ctx, h_pos, 3, (int64_t)nleaf * g_max_parts) ||
        p2p_upload_leaves(g_ctx, g_leaf_npart.data(), ipart.data(), nleaf) || p2p_clear_tasks(g_ctx) ||
        p2p_append_tasks_interleaved(g_ctx, h_int, ntasks, 0) || p2p_build_csr(g_ctx)) {
        printf("Error copy for GPU data : %s\n", p2p_last_error());
        return -1;
    }
    g_state = 0;
    return 0;
}

int LaunchKernelP2PIndexing(int nTasks, int posChunk, int leafChunk, int resultChunk, double SoftenScale, double MASSPART,
                            int verbosity_gpu) {
    (void)leafChunk; (void)verbosity_gpu;
    if (!g_ctx || g_state != 0) return g_state ? g_state : -1;
    if (nTasks != g_ntasks || posChunk != g_max_parts * 3 || resultChunk != g_max_parts * 3) {
        printf("ERROR : launch arguments do not match the uploaded data\n");
        return -1;
    }
    if (p2p_set_physics(g_ctx, MASSPART, SoftenScale, g_rs) || p2p_zero_acc(g_ctx) || p2p_compute(g_ctx) || p2p_synchronize(g_ctx)) {
        printf("error kernel ComputeP2PIndexing : %s\n", p2p_last_error());
        return -1;
    }
    return 0;
}

void readResultsGPU(double* h_acc, int nTasks, int maxPartsInLeaf, int verbosity_gpu) {
    (void)verbosity_gpu;
    if (!g_ctx || g_state != 0 || nTasks != g_ntasks || maxPartsInLeaf != g_max_parts) {
        printf("Error copy for reading results (h_acc_data)\n");
        return;
    }
    const size_t chunk = (size_t)g_max_parts * 3;
    std::vector<double> acc((size_t)g_nleaf * chunk);
    if (p2p_download_acc(g_ctx, acc.data(), 3, 0)) { printf("Error copy for reading results : %s\n", p2p_last_error()); return; }
    memset(h_acc, 0, sizeof(double) * chunk * (size_t)nTasks);
    for (int l = 0; l < g_nleaf; l++) {
        const long long n0 = g_first_task[(size_t)l];
        if (n0 < 0) continue;
        memcpy(h_acc + (size_t)n0 * chunk, acc.data() + (size_t)l * chunk, sizeof(double) * 3 * (size_t)g_leaf_npart[(size_t)l]);
    }
}

}  // extern "C"
