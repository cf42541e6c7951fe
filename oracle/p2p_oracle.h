/* p2p_oracle.h -- CPU (fp64, plain C) restatement of the photoNs-2.0 P2P hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product library never does.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference,
 * I/ = 1_Indexing/, R/ = 2_Redundant/).  Pinning: the reference ships no tests or golden vectors
 * (SURVEY.md section 4), so the oracle is pinned against the reference's OWN host code run here
 * (oracle/_ref/ref_lists, built from the mounted sources by oracle/Makefile): tree arrays, particle
 * permutation, local lists, pruned halo trees and remote lists must match bit for bit
 * (tests/test_oracle_vs_ref.py), and against the fixtures it generated (tests/golden/).
 * The pair arithmetic has no CPU definition in the reference tree (p2p_kernel is declared in
 * I/inc/kernels.h:4-8 and defined nowhere); it is restated from the device code cited below.
 *
 * Index convention (same as the reference, I/src/fmm.c:199-212): particles 0..npart-1, leaves
 * first_leaf = npart .. first_leaf+nleaf-1, nodes first_node = npart+nleaf_cap .. ; arrays passed
 * here are 0-based per kind, `son` values are the reference's global ids.
 */
#ifndef P2P_ORACLE_H
#define P2P_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* I/src/fmm.c:203-212 : capacities and id bases */
void oracle_tree_caps(int npart, int maxleaf, int* nleaf_cap, int* nnode_cap);

/* I/src/fmm.c:29-77 (bksort_inplace), :79-118 (build_kdtree), :120-174 (center_kdtree),
 * :176-263 (build_localtree).  pos[npart][3] and payload[npart] are permuted in place.
 * Returns 0, or -1 if capacities are exceeded. */
int oracle_build_localtree(int npart, int maxleaf, int direct_start, const double bdl[3], const double bdr[3],
                           double* pos, int64_t* payload,
                           int* leaf_npart, int* leaf_ipart, double* leaf_center, double* leaf_width,
                           int* node_npart, int* node_son, double* node_split, double* node_center,
                           double* node_width, int* nleaf_out, int* nnode_out);

/* I/src/fmm.c:266-325 : 0 open, 1 accept, -1 abort (LONGSHORT branch compiled in) */
int oracle_acceptance(const double wi[3], const double wj[3], const double dist[3], double theta, double rcut);

/* I/src/fmm.c:402-534 (walk_task_p2p) from (first_node, first_node).  Writes up to cap tasks as
 * global ids (tt = target leaf, ts = source leaf) in the reference's traversal order; returns the
 * total count (may exceed cap). */
int64_t oracle_walk_p2p(int npart, int nleaf_cap, int nleaf, int nnode, const double* leaf_center,
                        const double* leaf_width, const int* node_son, const double* node_center,
                        const double* node_width, double theta, double rcut, int* tt, int* ts, int64_t cap);

/* I/src/remotes.c:337-446 (prepare_sendtree2): prune the local tree against a target domain box,
 * shipping leaf bodies displaced by `displace`.  Outputs a RemoteNode/RemoteBody image
 * (I/inc/photoNs.h:202-214).  Returns 0, counts in *nnode_out / *nbody_out. */
int oracle_prune_sendtree(int npart, int nleaf_cap, const double* pos, const int* leaf_npart, const int* leaf_ipart,
                          const double* leaf_center, const double* leaf_width, const int* node_npart,
                          const int* node_son, const double* node_center, const double* node_width,
                          const double tcenter[3], const double twidth[3], const double displace[3], double theta,
                          double rcut, int cap_node, int cap_body, int* r_npart, int* r_son, double* r_center,
                          double* r_width, double* r_body, int* nnode_out, int* nbody_out);

/* I/src/remotes.c:141-317 (walk_task_p2p_ext) from (first_node, 0): local tree x received pruned
 * tree.  tt = local leaf global id, ts = remote node index. */
int64_t oracle_walk_p2p_ext(int npart, int nleaf_cap, int maxleaf, const double* leaf_center, const double* leaf_width,
                            const int* node_son, const double* node_center, const double* node_width,
                            int r_nnode, const int* r_npart, const int* r_son, const double* r_center,
                            const double* r_width, double theta, double rcut, int* tt, int* ts, int64_t cap);

/* Pair arithmetic: plain I/src/photoNs_CUDA.cu:342-354; truncated R/src/photoNs_CUDA.cu:432-450 with
 * rs = splitRadius, coeff = 2/sqrt(pi) (R/src/fmm.c:796-798).  Target/source roles I/src/fmm.c:875-876,
 * reduction `+=` per target particle I/src/fmm.c:895-908.  Targets: particles
 * [t_ipart[t], +t_npart[t]) of tpos; sources [s_start[s], +s_count[s]) of spos (local leaves, or
 * received bodies R/src/remotes.c:63-64).  tt/ts are 0-based leaf indices.  rs <= 0 selects plain.
 * acc[ntarget_particles][3] is ACCUMULATED into.  OpenMP over target leaves when nthreads != 1.
 * Returns the pair count sum n_t*n_s. */
int64_t oracle_p2p_tasks(const double* tpos, const int* t_npart, const int* t_ipart, int n_tleaf, const double* spos,
                         const int* s_count, const int* s_start, const int* tt, const int* ts, int64_t ntask,
                         double mass, double eps, double rs, double* acc, int nthreads);

/* sum |dx * ir3 * g| per target particle (the "sum of |terms|" norm of SURVEY.md section 8d / H1) */
int64_t oracle_p2p_absterms(const double* tpos, const int* t_npart, const int* t_ipart, int n_tleaf,
                            const double* spos, const int* s_count, const int* s_start, const int* tt,
                            const int* ts, int64_t ntask, double mass, double eps, double rs, double* absacc,
                            int nthreads);

/* FNV-1a style fingerprint of an int32 stream as defined in SURVEY.md section 8c */
uint64_t oracle_fingerprint(const int32_t* v, int64_t n);

/* Domain decomposition, I/src/initial.c:204-228 (setup_domain_index), I/src/domains.c:401-469
 * (domain_initialize: equal-volume splits), I/src/toptree.c:150-182 (center_toptree).
 * Fills split[2P-1], center/width[2P-1][3]; this_domain etc. derive from rank (see .c). */
void oracle_domain_setup(int nproc, double box, double* split, double* center, double* width, int* direct_of_node);
int oracle_domain_of_rank(int nproc, int rank);
/* I/src/domains.c:163-296: in-place routing partition of pos/payload by the domain tree;
 * sendcount[nproc] gets the number of particles bound for each rank, stored contiguously in
 * rank-tree traversal order exactly as the reference leaves them. */
void oracle_domain_partition(int nproc, const double* split, double* pos, int64_t* payload, int npart, int* sendcount,
                             int* sendorder);
/* work-weighted relaxation of the rank-tree splits, I/src/domains.c:20-38,86-157 */
void oracle_domain_relax(int nproc, double box, double* split, const double* frac);
/* Mid-field: P2M / M2M / M2L / L2L / L2P restated from 1_Indexing/src/operator.c:13-530 and driven as
 * fmm_prepare / fmm_task / fmm_ext drive them (1_Indexing/src/fmm.c:745-790,562-705,913-945,1026-1145;
 * 1_Indexing/src/remotes.c:477-640).  Single rank; box > 0 adds the 26 periodic images; literal_d6 replays the
 * reference's zero-shift self exchange (defect D6).  Pinned against the reference's own operators in
 * tests/test_oracle_vs_ref.py.  acc[npart][3] is ACCUMULATED into. */
int oracle_midfield(int npart, int nleaf_cap, int nleaf, int nnode, int maxleaf, const double* pos, const int* leaf_npart,
                    const int* leaf_ipart, const double* leaf_center, const double* leaf_width, const int* node_npart,
                    const int* node_son, const double* node_center, const double* node_width, double theta, double rcut,
                    double rs, double mass, double box, int literal_d6, double* leaf_M, double* node_M, double* leaf_L,
                    double* acc, int64_t* nm2l_local, int64_t* nm2l_total);
int oracle_max_threads(void);
#ifdef __cplusplus
}
#endif
#endif
