/* photoNs_CUDA_indexing.h -- drop-in C-ABI of the reference's "Indexing" GPU path
 * (replaces 1_Indexing/inc/photoNs_CUDA.cuh; implemented by lib/libphotoNs_CUDA_indexing.so).
 *
 * Exactly the symbols 1_Indexing/inc/photoNs_CUDA.cuh:24-33 declares, with identical signatures,
 * so 1_Indexing/src/fmm.c (task_compute_p2p, :842-911) and 1_Indexing/src/remotes.c
 * (task_compute_p2p_ext, :40-112) link unmodified.  The __global__ prototype the reference keeps
 * inside its extern "C" block (:34-39) is intentionally absent: C callers no longer need
 * cuda_runtime.h.
 *
 * Layouts (1_Indexing/src/fmm.c:721-731,851-877):
 *   h_pos[leaf*posChunk + j*3 + k]   padded fp64 positions, posChunk = maxPartsInLeaf*3
 *   h_leaf[leaf*2 + {0,1}]           {npart, ipart}
 *   h_interactions[n*2 + {0,1}]      {target leaf, source leaf}, 0-based
 *   h_acc[n*resultChunk + i*3 + k]   per-TASK partial accelerations; the caller sums them per
 *                                    target particle (1_Indexing/src/fmm.c:895-908).
 * Semantics are the INTENDED ones (SURVEY defects D1, D2 fixed: task 0 is computed, leaves up to 32
 * particles).  Because the caller only ever sums the task slots of a target leaf, the library
 * computes the per-leaf REDUCED acceleration with the CSR kernel and returns it in the slot of the
 * leaf's first task, zeros in its other slots -- the caller's sum is unchanged.
 * Return codes as the reference: 0 ok, -1 CUDA failure or no device, -3 task ids outside the
 * uploaded leaves (the reference's remote calls pass such ids, SURVEY defect D3).
 */
#ifndef PHOTONS_CUDA_INDEXING_H
#define PHOTONS_CUDA_INDEXING_H
#ifdef __cplusplus
extern "C" {
#endif
void initGPU(int verbosity_gpu);                                                         /* cu:20-43   */
void getGPUMemoryState(int verbosity_gpu);                                               /* cu:174-186 */
int allocMemGPU(int nleafs, int maxPartsInLeaf, int maxNeighbors, int nTasks, int verbosity_gpu);   /* cu:46-131 */
int copyMemGPU(double* h_pos_data, int* h_leaf_data, int* h_interactions, int numtasks, int verbosity_gpu); /* cu:133-172 */
int LaunchKernelP2PIndexing(int ntasks, int posChunk, int leafChunk, int resultChunk, double SoftenScale,
                            double MASSPART, int verbosity_gpu);                         /* cu:200-246 */
void readResultsGPU(double* h_acc_data, int nTasks, int maxPartsInLeaf, int verbosity_gpu); /* cu:188-197 */

/* Extension (not in the reference): the PM split radius never reaches the device through the
 * signatures above (SURVEY defect D4).  rs > 0 selects the erfc-truncated kernel
 * (2_Redundant/src/photoNs_CUDA.cu:443-446); the default (or rs <= 0) is the plain kernel the
 * reference compiles.  The environment variable P2P_B200_RS sets the same value. */
void p2pSetSplitRadius(double rs);
#ifdef __cplusplus
}
#endif
#endif
