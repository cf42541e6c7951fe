/* photoNs_CUDA_redundant.h -- drop-in C-ABI of the reference's "Redundant" GPU path
 * (replaces 2_Redundant/inc/photoNs_CUDA.cuh; implemented by lib/libphotoNs_CUDA_redundant.so).
 *
 * Exactly the symbols 2_Redundant/inc/photoNs_CUDA.cuh:28-51 declares.  Callers:
 * task_compute_p2p (2_Redundant/src/fmm.c:790-881) and task_compute_p2p_ext
 * (2_Redundant/src/remotes.c:40-159).
 *
 * "Redundant" layout = every task carries PRIVATE copies of its particles:
 *   local  (SelfInteractions): part_idx[n*3] = {nT, nS, target leaf}; part_data[n*partDataChunk + ...] =
 *          nT targets (x,y,z) then nS sources (x,y,z) fp64; result[n*resultDataChunk + i*3 + k].
 *   remote (DualNaive): h_pos_index[r][n*5] = {posStart, target leaf, nT, nS, resultIdx};
 *          h_pos_data[r][posStart ...] = nT targets then nS sources; result at resultIdx + i*3 + k.
 * Semantics are the intended ones (SURVEY defects D8, D10, D11, D13, D14 fixed: every task computed,
 * leaves up to 32 particles, per-target results, no spurious jp==ip skip for non-self tasks).
 */
#ifndef PHOTONS_CUDA_REDUNDANT_H
#define PHOTONS_CUDA_REDUNDANT_H
#ifdef __cplusplus
extern "C" {
#endif
void initGPU(int verbosity_gpu);                                                          /* cu:23-44 */
void getGPUMemoryState(int verbosity_gpu);                                                /* cu:136-147 */
int allocMemGPU(int PROC_SIZE, int maxPartsInLeaf, int MAXTASK, int verbosity_gpu);       /* cu:46-104 */
int copyMemGPU(double** h_pos_data, int** h_pos_index, int PROC_SIZE, int PROC_RANK, int maxPartsInLeaf, int numtasks,
               int posCounter, int verbosity_gpu);                                        /* cu:106-134 */
void readResultsGPU(double** h_acc_data, int PROC_RANK, int PROC_SIZE, int maxPartsInLeaf, int MAXTASK, int partCounter,
                    int verbosity_gpu);                                                   /* cu:149-160 */
int LaunchKernelP2PDualNaive(int PROC_SIZE, int PROC_RANK, int nTasks, double SoftenScale, double MASSPART,
                             int verbosity_gpu);                                          /* cu:184-221 */
int allocAndCopySelfInteractionsGPU(double* part_data, int* part_idx, int partDataChunk, int partIndexChunk,
                                    int resultDataChunk, int nTasks);                     /* cu:313-368 */
void LaunchKernelP2PSelfInteractions(int nTasks, int partDataChunk, int partIndexChunk, int resultDataChunk,
                                     double SoftenScale, double MASSPART);                /* cu:370-384 */
void readResultsGPUSelfInteractions(double* h_acc_data, int accDataChunk, int nTasks);    /* cu:459-466 */

/* Extension, see photoNs_CUDA_indexing.h */
void p2pSetSplitRadius(double rs);
#ifdef __cplusplus
}
#endif
#endif
