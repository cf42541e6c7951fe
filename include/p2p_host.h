/* p2p_host.h -- host-side producers of the P2P path (libp2p_host.so, plain C-ABI, no CUDA).
 *
 * These are the B200 build's own implementations of the reference host functions that feed the
 * GPU path; list parity (bit-exact trees, identical task sequences) is tested against the oracle
 * and the compiled reference.  Names follow the reference:
 *
 *   p2p_build_localtree     build_localtree + bksort_inplace/build_kdtree/center_kdtree
 *                           (1_Indexing/src/fmm.c:29-263) -- subtree-parallel, same ids/boxes
 *   p2p_walk_task_p2p       walk_task_p2p from (first_node, first_node) (1_Indexing/src/fmm.c:402-534)
 *                           -- frontier-parallel, emits the SAME SEQUENCE of (target, source) pairs
 *   p2p_prepare_sendtree    prepare_sendtree2 (1_Indexing/src/remotes.c:337-446)
 *   p2p_walk_task_p2p_ext   walk_task_p2p_ext from (first_node, 0) (1_Indexing/src/remotes.c:141-317)
 *   p2p_domain_*            setup_domain_index (1_Indexing/src/initial.c:204-228), domain_initialize
 *                           (1_Indexing/src/domains.c:401-469), center_toptree (1_Indexing/src/toptree.c:150-182),
 *                           prepare_body_inOrderOf_domain (1_Indexing/src/domains.c:163-296)
 *
 * All functions return 0 / a count on success and a negative value on failure.
 */
#ifndef P2P_HOST_H
#define P2P_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct p2p_tree p2p_tree;

/* A view of the tree arrays in the reference's layout: leaves are ids [first_leaf, first_leaf+nleaf),
 * nodes [first_node, first_node+nnode); `son` holds global ids (I/src/fmm.c:199-212). */
typedef struct {
    int npart, maxleaf, nleaf, nnode, nleaf_cap, nnode_cap, first_leaf, first_node;
    const int* leaf_npart;      /* [nleaf] */
    const int* leaf_ipart;      /* [nleaf] */
    const double* leaf_center;  /* [nleaf][3] */
    const double* leaf_width;   /* [nleaf][3] */
    const int* node_npart;      /* [nnode] */
    const int* node_son;        /* [nnode][2] */
    const double* node_split;   /* [nnode] */
    const double* node_center;  /* [nnode][3] */
    const double* node_width;   /* [nnode][3] */
} p2p_tree_view;

/* Builds the local kd-tree over `npart` records.  Each record is `stride_doubles` doubles whose
 * first three are the position (3 = packed positions, 12 = the reference's Body); whole records
 * and the optional payload are permuted in place exactly as the reference permutes part[].
 * bdl/bdr: domain box; direct_start: first split dimension (direct_local_start). */
int p2p_build_localtree(p2p_tree** out, double* records, int64_t stride_doubles, int64_t* payload, int npart,
                        int maxleaf, const double bdl[3], const double bdr[3], int direct_start, int nthreads);
void p2p_tree_free(p2p_tree* t);
int p2p_tree_get(const p2p_tree* t, p2p_tree_view* view);

/* Local dual-tree walk.  On success *tt / *ts point to library-owned arrays of *ntask 0-based leaf
 * indices (target, source) in the reference's traversal order; release with p2p_host_free. */
int p2p_walk_task_p2p(const p2p_tree* t, double theta, double rcut, int nthreads, int** tt, int** ts, int64_t* ntask);

/* Walk/compute pipeline (replaces the ping-pong task buffers of fmm_task / turn2compute_p2p,
 * 1_Indexing/src/fmm.c:365-400,947-1024): the dual-tree recursion is expanded to a frontier whose items are
 * grouped by TARGET chunk (a contiguous range of target leaves = a subtree of the local tree); the chunks
 * are then walked one at a time, so that the device can pack and compute chunk c while the host walks
 * chunk c+1.  The union of the chunks is exactly the task multiset of p2p_walk_task_p2p, and every task
 * of chunk c has its target leaf in [row_begin, row_end) of that chunk. */
typedef struct p2p_walk_plan p2p_walk_plan;
int p2p_walk_plan_create(const p2p_tree* t, double theta, double rcut, int nchunks_wanted, p2p_walk_plan** plan);
int p2p_walk_plan_nchunks(const p2p_walk_plan* plan);
int p2p_walk_plan_rows(const p2p_walk_plan* plan, int chunk, int* row_begin, int* row_end);
int p2p_walk_plan_run(const p2p_walk_plan* plan, int chunk, int nthreads, int** tt, int** ts, int64_t* ntask);
/* same, written straight into caller-owned buffers (e.g. pinned host memory that a DMA engine reads while
 * the next chunk is being walked); returns -3 with *ntask = needed size if cap is too small */
int p2p_walk_plan_run_into(const p2p_walk_plan* plan, int chunk, int nthreads, int* tt, int* ts, int64_t cap, int64_t* ntask);
void p2p_walk_plan_free(p2p_walk_plan* plan);

/* Pruned image of the local tree for one target domain box and displacement (the halo a peer
 * needs).  The image is the reference's RemoteNode/RemoteBody content (1_Indexing/inc/photoNs.h:202-214). */
typedef struct {
    int nnode, nbody;
    int* npart;        /* [nnode] */
    int* son;          /* [nnode][2]: leaves: {first body, end body}; pruned nodes: {-1,-1} */
    double* center;    /* [nnode][3] (displaced) */
    double* width;     /* [nnode][3] */
    double* body;      /* [nbody][3] (displaced) */
} p2p_image;
int p2p_prepare_sendtree(const p2p_tree* t, const double* records, int64_t stride_doubles, const double tcenter[3],
                         const double twidth[3], const double displace[3], double theta, double rcut, p2p_image* img);
void p2p_image_free(p2p_image* img);

/* Walk of the local tree against a received image: tt = 0-based local leaf, ts = image node index. */
int p2p_walk_task_p2p_ext(const p2p_tree* t, const p2p_image* img, double theta, double rcut, int nthreads, int** tt,
                          int** ts, int64_t* ntask);

void p2p_host_free(void* p);

/* Domain decomposition over P ranks (rank kd-tree in heap order, 2P-1 nodes). */
int p2p_domain_of_rank(int nproc, int rank);
int p2p_domain_setup(int nproc, double box, double* split, double* center, double* width, int* direct_of_node);
/* routes records in place; sendcount[r] = number of records bound for rank r, stored contiguously in rank order */
int p2p_domain_route(int nproc, const double* split, double* records, int64_t stride_doubles, int64_t* payload,
                     int64_t npart, int* sendcount);

/* work-weighted relaxation of the splits (measure_domain_runtime + determine_split_domtree,
 * 1_Indexing/src/domains.c:20-38,86-157); work[r] = task count of rank r; split[2P-1] updated in place */
int p2p_domain_relax(int nproc, double box, double* split, const double* work);

/* boxes (center_toptree, 1_Indexing/src/toptree.c:150-182) of the 2P-1 rank-tree nodes for given splits */
int p2p_domain_boxes(int nproc, double box, const double* split, double* center, double* width, int* direct_of_node);

/* Gadget-2 (format 1) snapshots (read_GadgetHeader / read_Particle_Gadget2 / write_Particle_Gadget2,
 * 1_Indexing/src/snapshot.c:5-22,211-293,397-503): float32 positions and velocities on disk, all particle types
 * concatenated, velocities scaled by a^(3/2) on input and back on output; ids are not stored (as in the reference). */
typedef struct {
    int64_t npart[6];           /* particles of each type in this file */
    int64_t nfile;              /* their sum */
    uint32_t npart_total[6];    /* of the whole snapshot */
    double mass[6];
    double time, redshift, box, omega0, omega_lambda, hubble;
    int num_files;
} p2p_snapshot_info;
int p2p_snapshot_header(const char* path, p2p_snapshot_info* info);
/* particles [n_start, n_start + n_count) of the file's order -> rows of doubles (pos / vel may be NULL) */
int p2p_snapshot_read(const char* path, int64_t n_start, int64_t n_count, double* pos, int64_t pos_stride, double* vel, int64_t vel_stride);
/* n_count particles as type 1 with info->mass[1], box, cosmology and redshift from info (vel NULL: zeros) */
int p2p_snapshot_write(const char* path, const p2p_snapshot_info* info, int64_t n_count, const double* pos, int64_t pos_stride,
                       const double* vel, int64_t vel_stride);

int p2p_host_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
