/* p2p_b200.h -- native C-ABI of the B200 near-field P2P library (libp2p_b200.so).
 *
 * The library replaces the body of the reference's GPU path behind
 *   1_Indexing/inc/photoNs_CUDA.cuh:24-33 and 2_Redundant/inc/photoNs_CUDA.cuh:28-51
 * (those exact symbols are exported by the compat shims, see photoNs_CUDA_indexing.h /
 * photoNs_CUDA_redundant.h in this directory).  The native API below carries what those
 * signatures cannot: the PM split radius r_s (never reaches the device in the reference, SURVEY
 * defect D4), persistent device-resident particles, ghost (halo / periodic image) leaves, a
 * device CSR of the (target leaf, source leaf) list and per-PARTICLE reduced accelerations.
 *
 * Conventions: plain pointers and sizes, no C++ or torch types; every function returns 0 on
 * success, a negative p2p_status on failure (p2p_last_error() gives the text); nothing is printed
 * unless P2P_B200_VERBOSE is set.  Calls on one context must be serialised by the caller (the
 * reference drives its GPU from one worker thread, 1_Indexing/src/fmm.c:390); every entry point
 * sets the context's device first, so it may be called from any host thread.
 *
 * Data model
 *   particles : int4 {xi, yi, zi, mass bits} on device: 32-bit fixed-point coordinates over the box
 *               given to p2p_set_box (resolution extent / 2^32 whatever the box size; coordinate
 *               differences wrap to the minimal image, so periodic-image and halo sources may be
 *               passed displaced or not); local particles [0, npart) then ghost particles.
 *   leaves    : {first particle, count}; local leaves [0, nleaf) (targets AND sources), then ghost
 *               leaves (sources only: halo leaves of other domains and periodic images, already
 *               displaced by the sender exactly as 1_Indexing/src/remotes.c:360-366 does).
 *   tasks     : (target leaf, source leaf) int32 pairs in any order, as walk_task_p2p emits them
 *               (1_Indexing/src/fmm.c:402-534: ts[] = source, tt[] = target); packed on device into
 *               CSR rows per target leaf with ascending source ids.
 *   result    : acc[i] += sum over listed leaf pairs of  m * dx / max(r, eps)^3 * g(r / 2 r_s),
 *               g(u) = erfc(u) + 2/sqrt(pi) u exp(-u^2)  (2_Redundant/src/photoNs_CUDA.cu:432-450),
 *               g = 1 when r_s <= 0 (the plain kernel, 1_Indexing/src/photoNs_CUDA.cu:342-354).
 *               G is not applied (1_Indexing/src/photoNs.c:161 applies it in the kick).
 */
#ifndef P2P_B200_H
#define P2P_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct p2p_ctx p2p_ctx;

enum p2p_status {
    P2P_OK = 0,
    P2P_ERR_CUDA = -1,      /* a CUDA runtime call failed (same code the reference returns) */
    P2P_ERR_ARG = -2,       /* bad argument (null pointer, negative size, leaf larger than P2P_MAX_LEAF) */
    P2P_ERR_STATE = -3,     /* call order violated (e.g. compute before build_csr) */
    P2P_ERR_NODEVICE = -4   /* no CUDA device / driver: there is NO CPU fallback */
};

#define P2P_MAX_LEAF 32     /* largest TARGET leaf occupancy the kernels accept (ghost source leaves: 384) */

/* kernel variants, selectable for the ncu comparisons (default P2P_KERNEL_AUTO) */
enum p2p_kernel_variant {
    P2P_KERNEL_AUTO = 0,
    P2P_KERNEL_SCALAR = 1,  /* one FP32 op per instruction */
    P2P_KERNEL_PACKED = 2   /* fma.rn.f32x2 on target pairs (sm_100a FFMA2) */
};

const char* p2p_last_error(void);
int p2p_device_count(void);

int p2p_create(p2p_ctx** ctx, int device);
int p2p_destroy(p2p_ctx* ctx);

/* mass: particle mass used when per-particle masses are not supplied (MASSPART);
 * eps: SoftenScale; rs: splitRadius, <= 0 selects the plain (untruncated) kernel. */
int p2p_set_physics(p2p_ctx* ctx, double mass, double eps, double rs);
int p2p_set_kernel_variant(p2p_ctx* ctx, int variant);
/* Coordinate frame of the fixed-point positions: x in [origin, origin + extent) per axis, periodic
 * with period `extent` (pass the simulation box for periodic runs; any target-source separation
 * the lists imply must stay below extent / 2 per axis).  Must precede p2p_upload_particles.
 * If never called, the box is derived from the uploaded particles (bounding cube, doubled). */
int p2p_set_box(p2p_ctx* ctx, const double origin[3], double extent);
/* kernel tuning knobs for the ncu sweeps; 0 keeps the default.  targets_per_pass 0 / 32: the production kernel (one pass
 * per row, near and far slice bodies; sources per lane 1 / 2, min resident blocks per SM 3 / 4); 8 / 16: the
 * first-generation kernel in its final tuning (A/B baseline) */
int p2p_set_tuning(p2p_ctx* ctx, int targets_per_pass, int sources_per_lane, int min_blocks);
/* near / far classification of the list (truncated kernel only): a source leaf whose particles are all at
 * r >= 2 r_s u_far from all particles of the target leaf is evaluated by the cheaper far-field body.
 * u_far < 0: the default (1.25, what the far-field fit covers); 0: no far class; values below the default are refused */
int p2p_set_far_threshold(p2p_ctx* ctx, double u_far);
/* use an externally owned cudaStream_t (e.g. torch's current stream); NULL restores the own stream */
int p2p_set_stream(p2p_ctx* ctx, void* cuda_stream);
/* The context owns TWO sets of list buffers (tasks, CSR, row schedule, pair counter, timing events); every list call
 * works on the current one and this call exchanges them.  A multi-rank step walks and packs its remote (halo) list in
 * one set, on a second stream, while the force kernel still consumes the local list in the other
 * (replaces the ping-pong task buffers of 1_Indexing/src/fmm.c:365-400). */
int p2p_swap_lists(p2p_ctx* ctx);
/* force kernel blocks: 0 = persistent warps (default); k > 0 = a warp retires once it has run for k x 2^17 SM cycles
 * (about k x 70 us; at least one row), so that kernels of higher-priority streams (NCCL, the halo walk) get SM slots while
 * the force kernel runs; one last wave of blocks is persistent and finishes the schedule */
int p2p_set_force_blocks(p2p_ctx* ctx, int budget);
/* room for ghost leaves / particles behind the local ones, reserved BEFORE a force kernel is in flight (growing the
 * particle array later would have to wait for it) */
int p2p_reserve_ghosts(p2p_ctx* ctx, int nghostleaf, int64_t nghost);

/* Local particles in tree order.  pos: npart rows of 3 doubles, consecutive rows `stride_doubles`
 * apart (3 for a packed array, 12 for the reference's Body AoS, 1_Indexing/inc/typesdef.h:48-57).
 * Also zeroes the accelerations.  Local leaves: leaf_npart/leaf_ipart as in Pack
 * (1_Indexing/inc/typesdef.h:38-46); empty leaves are legal. */
int p2p_upload_particles(p2p_ctx* ctx, const double* pos, int64_t stride_doubles, int64_t npart);
int p2p_upload_leaves(p2p_ctx* ctx, const int* leaf_npart, const int* leaf_ipart, int nleaf);

/* Ghost sources.  Appends nbody displaced bodies (rows of 3 doubles, stride as above; the
 * reference's RemoteBody has stride 4) and nleaf ghost leaves {start (relative to this batch),
 * count}.  Returns the id of the first new leaf in *first_leaf_id.  p2p_clear_ghosts drops all. */
int p2p_append_ghosts(p2p_ctx* ctx, const double* pos, int64_t stride_doubles, int64_t nbody, const int* start,
                      const int* count, int nleaf, int* first_leaf_id);
/* same, bodies already on the device as float4 {x, y, z, mass} in simulation units (NCCL halo buffers) */
int p2p_append_ghosts_device(p2p_ctx* ctx, const void* d_xyzm, int64_t nbody, const int* start, const int* count,
                             int nleaf, int* first_leaf_id);
int p2p_clear_ghosts(p2p_ctx* ctx);

/* Task list.  tt = target leaf ids (local), ts = source leaf ids; source_offset is added to every
 * ts (0 for local lists, *first_leaf_id for a ghost batch whose ids are batch-relative). */
int p2p_clear_tasks(p2p_ctx* ctx);
int p2p_append_tasks(p2p_ctx* ctx, const int* tt, const int* ts, int64_t ntask, int source_offset);
/* interleaved {target, source} pairs, the reference's interactions_data layout (1_Indexing/src/fmm.c:873-877) */
int p2p_append_tasks_interleaved(p2p_ctx* ctx, const int* ts_pairs, int64_t ntask, int source_offset);
int p2p_build_csr(p2p_ctx* ctx);

/* Launch the force kernel over the whole CSR (asynchronous on the context's stream). */
int p2p_compute(p2p_ctx* ctx);
int p2p_zero_acc(p2p_ctx* ctx);
int p2p_synchronize(p2p_ctx* ctx);

/* acc: npart rows of 3 doubles `stride_doubles` apart; accumulate != 0 does acc += result
 * (the reference's update loop, 1_Indexing/src/fmm.c:895-908), else acc = result. */
int p2p_download_acc(p2p_ctx* ctx, double* acc, int64_t stride_doubles, int accumulate);

int p2p_counts(p2p_ctx* ctx, int64_t* ntask, int64_t* npairs);
/* totals over every p2p_compute since the accelerations were last zeroed (p2p_zero_acc / p2p_upload_particles):
 * what the current accelerations contain; summed on the device, so chunked pipelines never synchronise */
int p2p_accumulated_counts(p2p_ctx* ctx, int64_t* ntask, int64_t* npairs);
/* copies of the device CSR for parity tests: row_ptr[nleaf+1] (int64), col[ntask] (int32) */
int p2p_download_csr(p2p_ctx* ctx, int64_t* row_ptr, int* col);
/* class of every CSR column (1 = far, evaluated by the far-field body) and the number of far columns of every row (a
 * row's far columns come first, each class in ascending order); NULL skips */
int p2p_download_csr_class(p2p_ctx* ctx, unsigned char* is_far, int* row_far);
/* milliseconds spent by the last p2p_compute / p2p_build_csr launch sequence (CUDA events) */
int p2p_last_timings(p2p_ctx* ctx, float* ms_compute, float* ms_csr);

/* One-call convenience used by the compat shims and the e2e benchmark: host buffers in, host
 * accelerations out (H2D, CSR, kernel, D2H all inside).  Ghost arguments may be NULL / 0. */
int p2p_step_host(p2p_ctx* ctx, const double* pos, int64_t pos_stride, int64_t npart, const int* leaf_npart,
                  const int* leaf_ipart, int nleaf, const int* tt, const int* ts, int64_t ntask, double* acc,
                  int64_t acc_stride, int accumulate);

/* Pipelined variant: the tasks come in `nchunk` groups (group g = tasks [chunk_off[g], chunk_off[g+1]), e.g. the
 * target chunks of p2p_walk_plan or the 16384-task flushes of the Redundant walk, 2_Redundant/src/fmm.c:416-418).
 * While the force kernel of group g runs, the tasks of group g+1 are copied and packed on a second stream
 * (double-buffered list storage).  Source ids >= nleaf address the ghost leaves of this call.  A target leaf
 * may appear in several groups: the accelerations accumulate.  Host buffers should be pinned. */
int p2p_step_host_chunked(p2p_ctx* ctx, const double* pos, int64_t pos_stride, int64_t npart, const int* leaf_npart,
                          const int* leaf_ipart, int nleaf, const double* ghost_pos, int64_t ghost_stride, int64_t nghost,
                          const int* ghost_start, const int* ghost_count, int nghostleaf, const int* tt, const int* ts,
                          const int64_t* chunk_off, int nchunk, double* acc, int64_t acc_stride, int accumulate);

/* raw device pointers for plumbing layers that keep data on the GPU (torch / NCCL) */
void* p2p_device_particles(p2p_ctx* ctx);   /* int4[npart + nghost], fixed-point */
void* p2p_device_acc(p2p_ctx* ctx);         /* float4[npart], accelerations in .xyz */

/* ---- device-resident tree build and dual-tree walk -------------------------------
 * The list producers of the reference run on the host (1_Indexing/src/fmm.c:29-263 build_localtree, :402-534
 * walk_task_p2p, 1_Indexing/src/remotes.c:337-446 prepare_sendtree2, :141-317 walk_task_p2p_ext).  These entry
 * points produce the SAME tree (particle permutation, leaf / node ids, kd-cell boxes, split values -- bit for
 * bit, including the sequential fp64 mean and the partition's tie rules) and the SAME task multiset on the
 * device, so that a whole short-range step needs one upload of positions and one download of accelerations. */

/* build_localtree on the device.  pos: host rows of 3 doubles (stride in doubles) in the caller's order; bdl/bdr:
 * the domain box; direct_start: first split dimension.  Needs p2p_set_box and p2p_set_physics first.  On success
 * the context holds the particles in tree order, the leaves and the tree; accelerations are zeroed.
 * Fails with P2P_ERR_ARG where the reference would overrun its 2 NPART / MAXLEAF leaf / node capacity. */
int p2p_tree_build(p2p_ctx* ctx, const double* pos, int64_t stride, int64_t npart, int maxleaf, const double bdl[3],
                   const double bdr[3], int direct_start);
/* boxes and sons of a tree built on the host (p2p_build_localtree, ids as in p2p_tree_view: node_son holds global
 * ids) for p2p_tree_walk; particles and leaves must have been uploaded with p2p_upload_particles / _leaves */
int p2p_tree_upload(p2p_ctx* ctx, int maxleaf, int nleaf, int nnode, int first_leaf, int first_node,
                    const double* leaf_center, const double* leaf_width, const int* node_son, const double* node_center,
                    const double* node_width);
int p2p_tree_info(p2p_ctx* ctx, int* nleaf, int* nnode, int* nlevel, float* ms_build, float* ms_walk, int64_t* walk_items,
                  double* max_leaf_width);
/* copies of the device-built tree in the reference's layout (NULL pointers are skipped); perm[i] = index, in the
 * array given to p2p_tree_build, of the particle now at tree position i; node_son holds global ids with
 * first_leaf = npart and first_node = npart + 2 npart / maxleaf (1_Indexing/src/fmm.c:199-212) */
int p2p_tree_download(p2p_ctx* ctx, int64_t* perm, double* pos_sorted, int* leaf_npart, int* leaf_ipart, double* leaf_center,
                      double* leaf_width, int* node_npart, int* node_son, double* node_split, double* node_center,
                      double* node_width);
/* walk_task_p2p over the device tree; with period > 0 also the walks against the 26 periodic images of the same
 * tree, pruned against the target box {tcenter, twidth} as prepare_sendtree2 does and walked as walk_task_p2p_ext
 * does.  Image sources are listed under their LOCAL leaf id (the fixed-point coordinates wrap to the nearest
 * image; needs box > 2 (r_cut + leaf size), see p2p_csr_duplicates).  Tasks are APPENDED to the context's list. */
int p2p_tree_walk(p2p_ctx* ctx, double theta, double rcut, double period, const double tcenter[3], const double twidth[3]);
/* the same walk restricted to the target leaves [leaf_lo, leaf_hi) (leaf_hi <= 0: all): a step can be split into target
 * chunks whose lists are walked, packed and consumed one after the other, which bounds the list memory (1024^3 on one
 * GPU lists 6.7e9 tasks); the union over a partition of the leaves is the task multiset of the full walk, and the M2L
 * list accumulates over the chunks.  Needs a device-built tree. */
int p2p_tree_walk_range(p2p_ctx* ctx, double theta, double rcut, double period, const double tcenter[3], const double twidth[3],
                        int leaf_lo, int leaf_hi);
/* Local lists and forces of the device-built tree in one call: walk (own tree and, with period > 0, its 26 images),
 * packing and the force kernel (compute != 0), in as many target chunks as keep one chunk's list below the limit of
 * p2p_set_chunk_tasks (default 2^29 tasks).  Accelerations accumulate; p2p_accumulated_counts gives the totals of the
 * step, p2p_counts / p2p_download_csr only see the last chunk. */
int p2p_forces_local(p2p_ctx* ctx, double theta, double rcut, double period, const double tcenter[3], const double twidth[3], int compute);
int p2p_set_chunk_tasks(p2p_ctx* ctx, int64_t max_tasks);
/* Chunk pipeline: a single-rank p2p_forces_local (no p2p_set_rank with nranks > 1) of at least 4096 x min_chunks leaves is cut
 * into at least min_chunks target chunks, and walk + packing of chunk k + 1 run on a second, high-priority stream into the
 * other list set while the force kernel of chunk k runs (the reference's ping-pong task buffers, 1_Indexing/src/fmm.c:365-400,
 * 947-1024, on the device).  Results are bit-identical to the unchunked step.  Default 0 = off: measured on B200 it gains
 * nothing at 256^3 and 1.4 % at 512^3 -- each of the ~50 dependent launches of a walk waits for force-kernel warps to
 * retire.  While it runs the context alternates its two list sets itself, so p2p_swap_lists must not be
 * used to hold a list across the call. */
int p2p_set_chunk_pipeline(p2p_ctx* ctx, int min_chunks);
/* device milliseconds of the last step: tree build, walks, packing and force kernels (summed over the chunks) */
int p2p_step_timings(p2p_ctx* ctx, float* ms_build, float* ms_walk, float* ms_csr, float* ms_force, int* nchunk);
/* number of sources listed twice in a row of the packed list (0 unless the periodic box is too small) */
int p2p_csr_duplicates(p2p_ctx* ctx, int64_t* ndup);
/* accelerations in the ORDER OF THE POSITIONS GIVEN TO p2p_tree_build, packed rows of 3 doubles */
int p2p_download_acc_original(p2p_ctx* ctx, double* acc);
/* The whole short-range P2P step of one rank in one call (replaces fmm_prepare + fmm_task + fmm_ext for the P2P part,
 * 1_Indexing/src/photoNs.c:97-123): host positions in the caller's order in, host accelerations in the same order out;
 * tree build, walk (26 periodic images when period > 0), packing and forces on the device.  Fails with P2P_ERR_ARG
 * when a listed leaf pair spans period / 2 or more along an axis (minimal-image sources would be ambiguous; the walk
 * checks every pair it lists). */
int p2p_step_device(p2p_ctx* ctx, const double* pos, int64_t stride, int64_t npart, int maxleaf, const double bdl[3],
                    const double bdr[3], int direct_start, double theta, double rcut, double period, double* acc);
/* ---- multi-rank device path: one tree per rank, every rank walks its own tree against all of them -------------------
 * Replaces fmm_ext / fmm_remote (1_Indexing/src/fmm.c:1026-1145, 1_Indexing/src/remotes.c:740-809): instead of pruning,
 * shipping and re-walking 27 x P halo images, the ranks all-gather their tree TOPOLOGY (kd cells, tight leaf bounds, sons,
 * leaf sizes: no particles), each rank walks against every peer tree with the sender-side cuts of prepare_sendtree2
 * evaluated on the fly, and only the particles of the leaves the lists actually reference travel afterwards.
 * All pointers below are DEVICE pointers owned by the caller (e.g. torch tensors used with torch.distributed).
 * Every rank copies its topology into a block of fixed stride (p2p_topology_stride of the LARGEST tree of the job), one
 * all_gather_into_tensor collects the P blocks, and p2p_tree_walk_peers_packed walks against them in place.  With
 * include_me = 0 the rank's own tree is left out: its walk (p2p_tree_walk, own periodic images included) needs no
 * communication, so the force kernel can work on the local list while topology, remote walk and halo particles are
 * still on their way (p2p_swap_lists gives the remote list its own buffers).  p2p_set_rank tells the local walk which
 * peer index this rank has (carried by the M2L tasks). */
int p2p_set_rank(p2p_ctx* ctx, int rank, int nranks);
int p2p_topology_stride(int nleaf_max, int nnode_max, int64_t* stride_bytes);
int p2p_tree_export_packed(p2p_ctx* ctx, void* d_block, int nleaf_max, int nnode_max);
int p2p_tree_walk_peers_packed(p2p_ctx* ctx, double theta, double rcut, double period, const double tcenter[3], const double twidth[3],
                               int npeer, int me, const int* peer_nleaf, const int* peer_nnode, const void* d_all, int nleaf_max,
                               int nnode_max, int include_me);
/* Leaf-granular halo fetch, planned on the device (csrc/halo.cuh); the host only learns the split sizes of the exchange.
 *  plan_need : marks (one byte per ghost leaf, peers in rank order without me; DEVICE buffer of the caller, to be sent to the
 *              owners) of the ghost leaves the current list references, the ghost leaf table behind the local leaves, and
 *              need_total[p] = particles wanted from rank p
 *  plan_give : d_asked = the marks the nreq = P - 1 other ranks sent (requester-major, nleaf bytes each) ->
 *              give_total[q] = particles to send to the q-th requester
 *  gather    : the requested leaves' fixed-point particles (int4), requester-major, into d_send
 *  set_particles : what arrived (sender-major = ghost leaf order) becomes the ghost particles */
int p2p_halo_plan_need(p2p_ctx* ctx, const void* d_topo_all, int npeer, int me, const int* peer_nleaf, int nleaf_max, int nnode_max,
                       void* d_marks, int64_t* need_total);
int p2p_halo_plan_give(p2p_ctx* ctx, const void* d_asked, int nreq, int64_t* give_total);
int p2p_halo_gather(p2p_ctx* ctx, const void* d_asked, int nreq, void* d_send);
int p2p_halo_set_particles(p2p_ctx* ctx, const void* d_recv, int64_t nbody);

/* ---- particle routing on the device (domain_decomposition: prepare_body_inOrderOf_domain + exchange,
 * 1_Indexing/src/domains.c:163-377): the slab a rank holds is partitioned by the rank kd-tree with the reference's
 * in-place partition (same group order, same order inside a group), the groups travel as device buffers, and the
 * tree is built from what arrived -- particles never return to the host between routing and forces. ------------- */
/* host slab -> resident device arrays; particle i carries the global id first_index + i through the exchange.
 * append != 0 adds the piece behind the particles already resident (a slab uploaded piece by piece). */
int p2p_route_load(p2p_ctx* ctx, const double* pos, int64_t stride, int64_t n, int64_t first_index, int append);
/* split[2P-1] in heap order (p2p_domain_setup / p2p_domain_relax); P a power of two; sendcount[r] = group size for rank r */
int p2p_route_partition(p2p_ctx* ctx, int nproc, const double* split, int* sendcount);
/* copies of / into the resident arrays; DEVICE pointers owned by the caller: x, y, z doubles, idx int32 */
int p2p_route_export(p2p_ctx* ctx, void* d_x, void* d_y, void* d_z, void* d_idx);
int p2p_route_import(p2p_ctx* ctx, const void* d_x, const void* d_y, const void* d_z, const void* d_idx, int64_t n);
/* build_localtree over the resident particles in their current order (what the reference's part[] holds after the
 * exchange); afterwards as after p2p_tree_build, except that results are read with p2p_download_acc (tree order) and
 * p2p_download_index (the global id of every tree position) */
int p2p_tree_build_resident(p2p_ctx* ctx, int maxleaf, const double bdl[3], const double bdr[3], int direct_start);
int p2p_download_index(p2p_ctx* ctx, int64_t* idx);

/* ---- device-resident stepping (the integrator part): positions, velocities and ids stay in
 * HBM between steps -- no per-step upload of positions or download of accelerations.  The kick / drift arithmetic is the
 * reference's (1_Indexing/src/photoNs.c:161-208: vel += acc dkh, pos += vel dd, wrap into [0, BOXSIZE) by its while
 * loops); like the reference (fmm_construct after the drift) every step's tree is built from the particle order the
 * previous step left.  The PM term (acc_pm) and the scale-factor integrals dk, dd belong to the caller. */
int p2p_resident_load(p2p_ctx* ctx, const double* pos, int64_t pos_stride, const double* vel, int64_t vel_stride, int64_t n, int64_t first_id);
int p2p_resident_forces(p2p_ctx* ctx, int maxleaf, const double bdl[3], const double bdr[3], int direct_start, double theta, double rcut,
                        double period);
/* the build alone (velocities and ids carried along): a multi-rank step continues with the exchange instead of the local forces */
int p2p_resident_build(p2p_ctx* ctx, int maxleaf, const double bdl[3], const double bdr[3], int direct_start);
/* vel += (P2P + mid-field) dkh.  With the M2L lists enabled (p2p_midfield_enable) p2p_resident_forces also computes the
 * mid-field, and the kick fails with P2P_ERR_STATE if it is missing (a multi-rank step computes it with
 * p2p_midfield_compute_peers_packed): the pairs the walk hands to the expansions must not be dropped silently. */
int p2p_resident_kick(p2p_ctx* ctx, double dkh);
/* multi-rank resident stepping: after the drift the particles migrate to their owners WITH velocities and ids
 * (domain_decomposition every step, 1_Indexing/src/domains.c:298-377): partition by the rank kd-tree (sendcount[r] = group
 * for rank r), export / import of the state as DEVICE arrays of the caller: 6 doubles arrays x, y, z, vx, vy, vz and int32 ids */
int p2p_resident_partition(p2p_ctx* ctx, int nproc, const double* split, int* sendcount);
int p2p_resident_export(p2p_ctx* ctx, void* const d_xv[6], void* d_id);
int p2p_resident_import(p2p_ctx* ctx, const void* const d_xv[6], const void* d_id, int64_t n);
int p2p_resident_drift(p2p_ctx* ctx, double dd, double period);
/* current positions / velocities (packed rows of 3 doubles) and ids (particle i of p2p_resident_load has id first_id + i), in the
 * resident (tree) order; NULL skips */
int p2p_resident_download(p2p_ctx* ctx, double* pos, double* vel, int64_t* id);
/* particles currently resident on this device (after p2p_route_load / _import / p2p_resident_load) */
int p2p_resident_count(p2p_ctx* ctx, int64_t* n);

/* ---- mid-field on the device: P2M / M2M / M2L / L2L / L2P --------------------------
 * The other half of the short-range FMM force (1_Indexing/src/operator.c; fmm_prepare, task_compute_m2l and the tail
 * of fmm_ext in 1_Indexing/src/fmm.c:745-790,913-945,1121-1128): third-order Cartesian expansions in fp64 with the
 * erfc-split radial factors.  The M2L task list is the set of pairs the walk's acceptance criterion hands to the
 * expansions (walk_task_m2l / walk_task_m2l_ext), emitted by p2p_tree_walk in the same pass as the P2P list. */
/* on: the following walks also emit the M2L list.  literal_d6 (test knob): also walk the local tree against itself
 * with the remote rules as the reference's zero-shift self exchange does (SURVEY defect D6: every local P2P and
 * M2L task twice), to compare with the reference's literal output. */
int p2p_midfield_enable(p2p_ctx* ctx, int on, int literal_d6);
/* P2M -> M2M -> M2L -> L2L -> L2P for the device-built tree and the last walk (single rank: local tree + periodic
 * images); p2p_download_acc_original then returns P2P + mid-field.  *nm2l: number of M2L tasks. */
int p2p_midfield_compute(p2p_ctx* ctx, int64_t* nm2l);
/* multi-rank: the local multipoles (P2M + M2M) copied to d_M [(nleaf + nnode)][20] doubles (DEVICE memory: this rank's
 * block of an all-gather), then M2L -> L2L -> L2P with the boxes (packed topology blocks) and multipoles (blocks of
 * (nleaf_max + nnode_max) x 20 doubles) of ALL ranks; the M2L list is the one the local and the remote walk left */
int p2p_midfield_multipoles(p2p_ctx* ctx, void* d_M);
int p2p_midfield_compute_peers_packed(p2p_ctx* ctx, int npeer, const void* d_topo_all, int nleaf_max, int nnode_max, const void* d_M_all,
                                      int64_t* nm2l);
/* multipoles and local expansions, [leaf][20] / [node][20] in the reference's coefficient order
 * (1_Indexing/inc/operator.h:24-67); NULL skips; *ms: device time of p2p_midfield_compute */
int p2p_midfield_download(p2p_ctx* ctx, double* leaf_M, double* node_M, double* leaf_L, double* node_L, float* ms);

/* test knob: runs up to this length use the plain in-order fold for the split mean, longer ones the exact parallel
 * evaluation of the same sequential sum.  Negative values select build paths (every one gives the same tree): -1 defaults;
 * -2 no block-per-node split mean; -3 no speculative chunks; -4 speculative chunks for every node above 2048 particles;
 * -5 particle-wide kernels for all levels; -6 no block-centric middle levels; -7 block-centric levels from the root on */
int p2p_tree_set_option(p2p_ctx* ctx, int seq_sum_plain_max);

#ifdef __cplusplus
}
#endif
#endif
