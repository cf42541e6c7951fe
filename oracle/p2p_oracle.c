/* p2p_oracle.c -- CPU (fp64, plain C) restatement of the photoNs-2.0 P2P hot path.
 * TEST INFRASTRUCTURE ONLY -- see p2p_oracle.h for who may use it and how it is pinned.
 * Compiled with -ffp-contract=off so that the fp64 tree/MAC arithmetic is evaluated exactly as the
 * reference's gcc -O2 x86-64 build evaluates it (no FMA contraction).
 */
#include "p2p_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ tree build ----------------- */

void oracle_tree_caps(int npart, int maxleaf, int* nleaf_cap, int* nnode_cap) {
    /* I/src/fmm.c:203-209 */
    int nn = (int)(2.0 * ((double)npart) / ((double)maxleaf));
    int nl = (int)(2.0 * ((double)npart) / ((double)maxleaf));
    if (nn > npart) nn = npart + 1;
    if (nl > npart) nl = npart + 1;
    *nleaf_cap = nl;
    *nnode_cap = nn;
}

typedef struct {
    int npart, maxleaf, nleaf_cap, nnode_cap;
    int first_leaf, first_node, last_leaf, last_node; /* reference ids; last_node is inclusive */
    double* pos;
    int64_t* payload;
    int *leaf_npart, *leaf_ipart, *node_npart, *node_son;
    double *leaf_center, *leaf_width, *node_split, *node_center, *node_width;
    int overflow;
} Build;

static void swap_body(Build* b, int base, int i, int j) {
    if (i == j) return;
    double* p = b->pos + 3 * (size_t)base;
    for (int k = 0; k < 3; k++) { double t = p[3 * i + k]; p[3 * i + k] = p[3 * j + k]; p[3 * j + k] = t; }
    if (b->payload) { int64_t t = b->payload[base + i]; b->payload[base + i] = b->payload[base + j]; b->payload[base + j] = t; }
}

/* I/src/fmm.c:29-77 : split a run of particles at the sequential mean of coordinate D */
static void mean_split(Build* b, int D, int base, int length, int np[2], double* split) {
    const double* p = b->pos + 3 * (size_t)base;
    np[0] = np[1] = 0;
    if (length < 2) { np[1] = length; return; }
    if (length == 2) {
        np[0] = np[1] = 1;
        *split = 0.5 * (p[D] + p[3 + D]);
        if (p[D] > p[3 + D]) swap_body(b, base, 0, 1);
        return;
    }
    double mean = 0.0;
    for (int n = 0; n < length; n++) mean += p[3 * n + D];
    mean /= (double)length;
    int but = length - 1;
    for (int n = 0; n < but; n++) {
        if (p[3 * n + D] > mean) {
            while (p[3 * but + D] > mean && but > n) but--;
            swap_body(b, base, n, but);
        }
    }
    np[0] = but;
    np[1] = length - but;
    *split = mean;
}

/* I/src/fmm.c:79-118 */
static void build_rec(Build* b, int direct, int ipart, int length, int inode) {
    if (length == 0 || b->overflow) return;
    int ni = inode - b->first_node;
    if (ni >= b->nnode_cap) { b->overflow = 1; return; }
    b->node_npart[ni] = length;
    int np[2];
    double split = 0.0;
    mean_split(b, direct, ipart, length, np, &split);
    b->node_split[ni] = split;
    int ip = ipart, nd = (direct + 1) % 3;
    for (int n = 0; n < 2; n++) {
        if (np[n] <= b->maxleaf) {
            int li = b->last_leaf - b->first_leaf;
            if (li >= b->nleaf_cap) { b->overflow = 1; return; }
            b->leaf_npart[li] = np[n];
            b->leaf_ipart[li] = ip;
            b->node_son[2 * ni + n] = b->last_leaf;
            b->last_leaf++;
        } else {
            b->last_node++;
            b->node_son[2 * ni + n] = b->last_node;
            build_rec(b, nd, ip, np[n], b->last_node);
        }
        ip += np[n];
    }
}

/* I/src/fmm.c:120-174 : boxes are the kd cells (domain box cut by the splits) */
static void center_rec(Build* b, int direct, int inode, double left[3], double right[3]) {
    int ni = inode - b->first_node;
    double* nw = b->node_width + 3 * (size_t)ni;
    double* nc = b->node_center + 3 * (size_t)ni;
    for (int k = 0; k < 3; k++) { nw[k] = right[k] - left[k]; nc[k] = 0.5 * (right[k] + left[k]); }
    int nd = (direct + 1) % 3;
    double split = b->node_split[ni];
    for (int n = 0; n < 2; n++) {
        int idx = b->node_son[2 * ni + n];
        if (idx < 0) continue;
        if (idx < b->last_leaf) {
            int li = idx - b->first_leaf;
            double* lw = b->leaf_width + 3 * (size_t)li;
            double* lc = b->leaf_center + 3 * (size_t)li;
            for (int k = 0; k < 3; k++) { lw[k] = nw[k]; lc[k] = nc[k]; }
            if (n == 0) { lw[direct] = split - left[direct]; lc[direct] = 0.5 * (left[direct] + split); }
            else        { lw[direct] = right[direct] - split; lc[direct] = 0.5 * (right[direct] + split); }
        } else if (n == 0) {
            double tmp = right[direct]; right[direct] = split;
            center_rec(b, nd, idx, left, right);
            right[direct] = tmp;
        } else {
            double tmp = left[direct]; left[direct] = split;
            center_rec(b, nd, idx, left, right);
            left[direct] = tmp;
        }
    }
}

int oracle_build_localtree(int npart, int maxleaf, int direct_start, const double bdl[3], const double bdr[3],
                           double* pos, int64_t* payload, int* leaf_npart, int* leaf_ipart, double* leaf_center,
                           double* leaf_width, int* node_npart, int* node_son, double* node_split,
                           double* node_center, double* node_width, int* nleaf_out, int* nnode_out) {
    Build b;
    memset(&b, 0, sizeof b);
    b.npart = npart; b.maxleaf = maxleaf;
    oracle_tree_caps(npart, maxleaf, &b.nleaf_cap, &b.nnode_cap);
    b.first_leaf = b.last_leaf = npart;
    b.first_node = b.last_node = npart + b.nleaf_cap;
    b.pos = pos; b.payload = payload;
    b.leaf_npart = leaf_npart; b.leaf_ipart = leaf_ipart; b.leaf_center = leaf_center; b.leaf_width = leaf_width;
    b.node_npart = node_npart; b.node_son = node_son; b.node_split = node_split; b.node_center = node_center;
    b.node_width = node_width;
    /* I/src/fmm.c:219-255 : zero / -1 initialisation */
    for (int n = 0; n < b.nnode_cap; n++) {
        node_npart[n] = 0; node_son[2 * n] = node_son[2 * n + 1] = -1; node_split[n] = 0.0;
        for (int k = 0; k < 3; k++) node_center[3 * n + k] = node_width[3 * n + k] = 0.0;
    }
    for (int n = 0; n < b.nleaf_cap; n++) {
        leaf_npart[n] = leaf_ipart[n] = 0;
        for (int k = 0; k < 3; k++) leaf_center[3 * n + k] = leaf_width[3 * n + k] = 0.0;
    }
    build_rec(&b, direct_start, 0, npart, b.first_node);
    if (b.overflow) return -1;
    double l[3] = {bdl[0], bdl[1], bdl[2]}, r[3] = {bdr[0], bdr[1], bdr[2]};
    center_rec(&b, direct_start, b.first_node, l, r);
    *nleaf_out = b.last_leaf - b.first_leaf;
    *nnode_out = b.last_node - b.first_node + 1;
    return 0;
}

/* ------------------------------------------------------------------ MAC + walks ---------------- */

/* I/src/fmm.c:266-325 */
int oracle_acceptance(const double wi[3], const double wj[3], const double dist[3], double theta, double rcut) {
    double w[3], mn[3];
    for (int k = 0; k < 3; k++) w[k] = (wi[k] + wj[k]) * 0.5;
    double dd2 = dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2];
    for (int k = 0; k < 3; k++) {
        mn[k] = dist[k];
        if (mn[k] < 0.0) mn[k] = -mn[k];
        mn[k] -= w[k];
        if (mn[k] <= 0.0) mn[k] = 0.0;
    }
    if (mn[0] + mn[1] + mn[2] < 0.0001) return 0;              /* touching boxes: open */
    double dm2 = mn[0] * mn[0] + mn[1] * mn[1] + mn[2] * mn[2];
    double c2 = rcut * rcut;
    if (dm2 >= c2) return -1;                                  /* box gap beyond the cutoff: abort */
    if (dd2 > 1.0 * c2) return 0;
    double wmax = w[0];
    if (w[1] > wmax) wmax = w[1];
    if (w[2] > wmax) wmax = w[2];
    wmax *= 2;
    if (wmax * wmax < theta * theta * dd2) return 1;           /* far enough: M2L's job, not P2P */
    return 0;
}

typedef struct {
    int first_leaf, last_leaf, first_node;
    const double *lc, *lw, *nc, *nw;
    const int* son;
    double theta, rcut;
    int *tt, *ts;
    int64_t n, cap;
    /* remote tree (walk_ext) */
    int maxleaf;
    const int *r_npart, *r_son;
    const double *r_center, *r_width;
} Walk;

static inline void emit(Walk* w, int im, int jm) {
    if (w->n < w->cap) { w->tt[w->n] = im; w->ts[w->n] = jm; }
    w->n++;
}
static inline const double* box_c(const Walk* w, int id) {
    return id < w->first_node ? w->lc + 3 * (size_t)(id - w->first_leaf) : w->nc + 3 * (size_t)(id - w->first_node);
}
static inline const double* box_w(const Walk* w, int id) {
    return id < w->first_node ? w->lw + 3 * (size_t)(id - w->first_leaf) : w->nw + 3 * (size_t)(id - w->first_node);
}
static inline int son_of(const Walk* w, int id, int n) { return w->son[2 * (size_t)(id - w->first_node) + n]; }

/* I/src/fmm.c:402-534 */
static void walk_rec(Walk* w, int im, int jm) {
    if (im == -1 || jm == -1) return;
    if (im == jm) {
        if (im < w->last_leaf) emit(w, im, jm);
        if (im >= w->first_node) {
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++) walk_rec(w, son_of(w, im, a), son_of(w, jm, b));
        }
        return;
    }
    int ileaf = im < w->first_node, jleaf = jm < w->first_node;
    if (ileaf && jleaf) { emit(w, im, jm); return; }            /* leaf-leaf: always, no MAC */
    const double *ci = box_c(w, im), *cj = box_c(w, jm);
    double dist[3] = {ci[0] - cj[0], ci[1] - cj[1], ci[2] - cj[2]};
    const double *wi = box_w(w, im), *wj = box_w(w, jm);
    /* the reference passes (leaf.width, node.width) in the mixed cases (I/src/fmm.c:455,481);
     * acceptance() only uses wi+wj, so the argument order cannot change a bit of the result */
    int flag = oracle_acceptance(wi, wj, dist, w->theta, w->rcut);
    if (flag != 0) return;
    if (ileaf) { walk_rec(w, im, son_of(w, jm, 0)); walk_rec(w, im, son_of(w, jm, 1)); return; }
    if (jleaf) { walk_rec(w, son_of(w, im, 0), jm); walk_rec(w, son_of(w, im, 1), jm); return; }
    if (wi[0] + wi[1] + wi[2] > wj[0] + wj[1] + wj[2]) {
        walk_rec(w, son_of(w, im, 0), jm); walk_rec(w, son_of(w, im, 1), jm);
    } else {
        walk_rec(w, im, son_of(w, jm, 0)); walk_rec(w, im, son_of(w, jm, 1));
    }
}

int64_t oracle_walk_p2p(int npart, int nleaf_cap, int nleaf, int nnode, const double* leaf_center,
                        const double* leaf_width, const int* node_son, const double* node_center,
                        const double* node_width, double theta, double rcut, int* tt, int* ts, int64_t cap) {
    (void)nnode;
    Walk w;
    memset(&w, 0, sizeof w);
    w.first_leaf = npart; w.last_leaf = npart + nleaf; w.first_node = npart + nleaf_cap;
    w.lc = leaf_center; w.lw = leaf_width; w.nc = node_center; w.nw = node_width; w.son = node_son;
    w.theta = theta; w.rcut = rcut; w.tt = tt; w.ts = ts; w.cap = cap;
    walk_rec(&w, w.first_node, w.first_node);
    return w.n;
}

/* I/src/remotes.c:141-317 */
static void walk_ext_rec(Walk* w, int im, int jm) {
    int ileaf = im < w->first_node;
    int jleaf = w->r_npart[jm] <= w->maxleaf;
    if (ileaf && jleaf) { emit(w, im, jm); return; }
    const double* ci = box_c(w, im);
    const double* cj = w->r_center + 3 * (size_t)jm;
    double dist[3] = {ci[0] - cj[0], ci[1] - cj[1], ci[2] - cj[2]};
    const double* wi = box_w(w, im);
    const double* wj = w->r_width + 3 * (size_t)jm;
    int flag = oracle_acceptance(wi, wj, dist, w->theta, w->rcut);
    int s0 = w->r_son[2 * (size_t)jm], s1 = w->r_son[2 * (size_t)jm + 1];
    if (flag == -1) return;
    if (ileaf) {                                               /* :183-221 */
        if (flag == 1 || s0 < 0 || s1 < 0) return;
        walk_ext_rec(w, im, s0); walk_ext_rec(w, im, s1);
        return;
    }
    if (jleaf) {                                               /* :226-262 */
        if (flag == 1) return;
        walk_ext_rec(w, son_of(w, im, 0), jm); walk_ext_rec(w, son_of(w, im, 1), jm);
        return;
    }
    if (flag == 1) return;                                     /* :267-316 */
    if (wi[0] + wi[1] + wi[2] > wj[0] + wj[1] + wj[2] || s0 < 0 || s1 < 0) {
        walk_ext_rec(w, son_of(w, im, 0), jm); walk_ext_rec(w, son_of(w, im, 1), jm);
    } else {
        walk_ext_rec(w, im, s0); walk_ext_rec(w, im, s1);
    }
}

int64_t oracle_walk_p2p_ext(int npart, int nleaf_cap, int maxleaf, const double* leaf_center, const double* leaf_width,
                            const int* node_son, const double* node_center, const double* node_width, int r_nnode,
                            const int* r_npart, const int* r_son, const double* r_center, const double* r_width,
                            double theta, double rcut, int* tt, int* ts, int64_t cap) {
    if (r_nnode <= 0) return 0;
    Walk w;
    memset(&w, 0, sizeof w);
    w.first_leaf = npart; w.first_node = npart + nleaf_cap; w.last_leaf = w.first_node;
    w.lc = leaf_center; w.lw = leaf_width; w.nc = node_center; w.nw = node_width; w.son = node_son;
    w.theta = theta; w.rcut = rcut; w.tt = tt; w.ts = ts; w.cap = cap;
    w.maxleaf = maxleaf; w.r_npart = r_npart; w.r_son = r_son; w.r_center = r_center; w.r_width = r_width;
    walk_ext_rec(&w, w.first_node, 0);
    return w.n;
}

/* ------------------------------------------------------------------ halo pruning --------------- */

typedef struct {
    int first_leaf, first_node;
    const double* pos;
    const int *leaf_npart, *leaf_ipart, *node_npart, *node_son;
    const double *lc, *lw, *nc, *nw;
    const double *tc, *tw, *disp;
    double theta, rcut;
    int cap_node, cap_body, numnode, numbody, overflow;
    int *r_npart, *r_son;
    double *r_center, *r_width, *r_body;
    int* r_src;                      /* optional: local id (reference numbering) of every image node */
} Prune;

/* I/src/remotes.c:337-446 */
static void prune_rec(Prune* p, int isend, int ilocal) {
    if (p->overflow) return;
    if (isend >= p->cap_node) { p->overflow = 1; return; }
    if (p->r_src) p->r_src[isend] = ilocal;
    if (ilocal < p->first_node && ilocal >= p->first_leaf) {
        int li = ilocal - p->first_leaf;
        p->r_npart[isend] = p->leaf_npart[li];
        for (int k = 0; k < 3; k++) {
            p->r_center[3 * isend + k] = p->lc[3 * li + k] + p->disp[k];
            p->r_width[3 * isend + k] = p->lw[3 * li + k];
        }
        p->r_son[2 * isend] = p->numbody;
        if (p->numbody + p->leaf_npart[li] > p->cap_body) { p->overflow = 1; return; }
        for (int q = p->leaf_ipart[li]; q < p->leaf_ipart[li] + p->leaf_npart[li]; q++) {
            for (int k = 0; k < 3; k++) p->r_body[3 * (size_t)p->numbody + k] = p->pos[3 * (size_t)q + k] + p->disp[k];
            p->numbody++;
        }
        p->r_son[2 * isend + 1] = p->numbody;
        p->numnode++;
        return;
    }
    int ni = ilocal - p->first_node;
    double d[3];
    for (int k = 0; k < 3; k++) {
        d[k] = p->tc[k] - p->nc[3 * ni + k] - p->disp[k];
        if (d[k] < 0.0) d[k] = -d[k];
        d[k] -= (p->tw[k] + p->nw[3 * ni + k]) * 0.5;
    }
    double dr = 0.0;
    for (int k = 0; k < 3; k++) if (d[k] > 0.0) dr += d[k] * d[k];
    dr = sqrt(dr);
    p->r_npart[isend] = p->node_npart[ni];
    for (int k = 0; k < 3; k++) {
        p->r_center[3 * isend + k] = p->nc[3 * ni + k] + p->disp[k];
        p->r_width[3 * isend + k] = p->nw[3 * ni + k];
    }
    p->numnode++;
    double wmax = p->nw[3 * ni];
    if (wmax < p->nw[3 * ni + 1]) wmax = p->nw[3 * ni + 1];
    if (wmax < p->nw[3 * ni + 2]) wmax = p->nw[3 * ni + 2];
    if (dr >= p->rcut || wmax < 0.95 * p->theta * dr) {
        p->r_son[2 * isend] = p->r_son[2 * isend + 1] = -1;
        return;
    }
    for (int n = 0; n < 2; n++) {
        int idx = p->node_son[2 * ni + n];
        if (idx >= p->first_leaf) {
            p->r_son[2 * isend + n] = p->numnode;
            prune_rec(p, p->numnode, idx);
        }
    }
}

int oracle_prune_sendtree(int npart, int nleaf_cap, const double* pos, const int* leaf_npart, const int* leaf_ipart,
                          const double* leaf_center, const double* leaf_width, const int* node_npart,
                          const int* node_son, const double* node_center, const double* node_width,
                          const double tcenter[3], const double twidth[3], const double displace[3], double theta,
                          double rcut, int cap_node, int cap_body, int* r_npart, int* r_son, double* r_center,
                          double* r_width, double* r_body, int* nnode_out, int* nbody_out) {
    Prune p;
    memset(&p, 0, sizeof p);
    p.first_leaf = npart; p.first_node = npart + nleaf_cap;
    p.pos = pos; p.leaf_npart = leaf_npart; p.leaf_ipart = leaf_ipart; p.node_npart = node_npart; p.node_son = node_son;
    p.lc = leaf_center; p.lw = leaf_width; p.nc = node_center; p.nw = node_width;
    p.tc = tcenter; p.tw = twidth; p.disp = displace; p.theta = theta; p.rcut = rcut;
    p.cap_node = cap_node; p.cap_body = cap_body;
    p.r_npart = r_npart; p.r_son = r_son; p.r_center = r_center; p.r_width = r_width; p.r_body = r_body;
    prune_rec(&p, 0, p.first_node);
    *nnode_out = p.numnode; *nbody_out = p.numbody;
    return p.overflow ? -1 : 0;
}

/* ------------------------------------------------------------------ pair arithmetic ------------ */

static void csr_by_target(const int* tt, int64_t ntask, int n_tleaf, int64_t** rowp, int64_t** order) {
    int64_t* rp = (int64_t*)calloc((size_t)n_tleaf + 1, sizeof(int64_t));
    int64_t* od = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ntask ? ntask : 1));
    for (int64_t k = 0; k < ntask; k++) rp[tt[k] + 1]++;
    for (int t = 0; t < n_tleaf; t++) rp[t + 1] += rp[t];
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_tleaf ? n_tleaf : 1));
    memcpy(cur, rp, sizeof(int64_t) * (size_t)n_tleaf);
    for (int64_t k = 0; k < ntask; k++) od[cur[tt[k]]++] = k;   /* stable: list order within a row */
    free(cur);
    *rowp = rp; *order = od;
}

static int64_t p2p_impl(const double* tpos, const int* t_npart, const int* t_ipart, int n_tleaf, const double* spos,
                        const int* s_count, const int* s_start, const int* tt, const int* ts, int64_t ntask,
                        double mass, double eps, double rs, double* acc, int nthreads, int absterms) {
    int64_t *rowp, *order;
    csr_by_target(tt, ntask, n_tleaf, &rowp, &order);
    const double coeff = 2.0 / sqrt(M_PI);                        /* R/src/fmm.c:798 */
    int64_t npairs = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(+ : npairs)
    for (int t = 0; t < n_tleaf; t++) {
        int nt = t_npart[t];
        for (int64_t e = rowp[t]; e < rowp[t + 1]; e++) {
            int s = ts[order[e]];
            int ns = s_count[s];
            if (nt <= 0 || ns <= 0) continue;                     /* empty leaves: I/src/photoNs_CUDA.cu:290-297 */
            npairs += (int64_t)nt * ns;
            const double* sp = spos + 3 * (size_t)s_start[s];
            for (int i = 0; i < nt; i++) {
                const double* tp = tpos + 3 * (size_t)(t_ipart[t] + i);
                double r0 = 0, r1 = 0, r2 = 0;
                for (int j = 0; j < ns; j++) {
                    double dx0 = sp[3 * j] - tp[0], dx1 = sp[3 * j + 1] - tp[1], dx2 = sp[3 * j + 2] - tp[2];
                    double dr = sqrt(dx0 * dx0 + dx1 * dx1 + dx2 * dx2);
                    double ir3;
                    if (dr < eps) ir3 = mass / (eps * eps * eps);   /* I/src/photoNs_CUDA.cu:346-350 */
                    else ir3 = mass / (dr * dr * dr);
                    if (rs > 0.0) {                                 /* R/src/photoNs_CUDA.cu:443-446 */
                        double drs = 0.5 * dr / rs;
                        ir3 *= (erfc(drs) + coeff * drs * exp(-drs * drs));
                    }
                    if (absterms) { r0 += fabs(dx0 * ir3); r1 += fabs(dx1 * ir3); r2 += fabs(dx2 * ir3); }
                    else { r0 += dx0 * ir3; r1 += dx1 * ir3; r2 += dx2 * ir3; }
                }
                double* a = acc + 3 * (size_t)(t_ipart[t] + i);     /* I/src/fmm.c:902-904 */
                a[0] += r0; a[1] += r1; a[2] += r2;
            }
        }
    }
    free(rowp); free(order);
    return npairs;
}

int64_t oracle_p2p_tasks(const double* tpos, const int* t_npart, const int* t_ipart, int n_tleaf, const double* spos,
                         const int* s_count, const int* s_start, const int* tt, const int* ts, int64_t ntask,
                         double mass, double eps, double rs, double* acc, int nthreads) {
    return p2p_impl(tpos, t_npart, t_ipart, n_tleaf, spos, s_count, s_start, tt, ts, ntask, mass, eps, rs, acc, nthreads, 0);
}
int64_t oracle_p2p_absterms(const double* tpos, const int* t_npart, const int* t_ipart, int n_tleaf,
                            const double* spos, const int* s_count, const int* s_start, const int* tt,
                            const int* ts, int64_t ntask, double mass, double eps, double rs, double* absacc,
                            int nthreads) {
    return p2p_impl(tpos, t_npart, t_ipart, n_tleaf, spos, s_count, s_start, tt, ts, ntask, mass, eps, rs, absacc, nthreads, 1);
}

uint64_t oracle_fingerprint(const int32_t* v, int64_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (int64_t i = 0; i < n; i++) { h ^= (uint32_t)v[i]; h *= 0x100000001b3ull; }
    return h;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ domain decomposition ------- */

static int mostleft_of(int P) {
    /* I/src/initial.c:206-215 */
    int m = 1;
    while (m < 2 * P - 1) m *= 2;
    m /= 2; m -= 1;
    if (P == 1) m = 0;
    return m;
}
int oracle_domain_of_rank(int nproc, int rank) {
    /* I/src/initial.c:218-221 */
    int d = rank + mostleft_of(nproc);
    if (d > 2 * nproc - 2) d -= nproc;
    return d;
}

static double fill_time(int P, int n, const int* son, double* tnode, double* tl, double* tr) {
    /* I/src/domains.c:5-18 */
    if (n >= P - 1) return tnode[n];
    if (son[2 * n] > 0) tl[n] = fill_time(P, son[2 * n], son, tnode, tl, tr);
    if (son[2 * n + 1] > 0) tr[n] = fill_time(P, son[2 * n + 1], son, tnode, tl, tr);
    tnode[n] = tl[n] + tr[n];
    return tnode[n];
}
static void volume_part(int P, int n, const int* son, const double* tl, const double* tr, double* split,
                        const double boxl[3], const double boxr[3], int dim) {
    /* I/src/domains.c:401-430 */
    if (n >= P - 1) return;
    double bl[3] = {boxl[0], boxl[1], boxl[2]}, br[3] = {boxr[0], boxr[1], boxr[2]};
    double norm = tl[n] + tr[n];
    double frac = bl[dim] + (br[dim] - bl[dim]) * tl[n] / norm;
    double brd = br[dim];
    br[dim] = frac;
    split[n] = frac;
    volume_part(P, son[2 * n], son, tl, tr, split, bl, br, (dim + 1) % 3);
    bl[dim] = frac; br[dim] = brd;
    volume_part(P, son[2 * n + 1], son, tl, tr, split, bl, br, (dim + 1) % 3);
}
static void center_top(int P, int direct, int n, const int* son, const double* split, double left[3], double right[3],
                       double* center, double* width, int* direct_of_node) {
    /* I/src/toptree.c:150-182 */
    for (int k = 0; k < 3; k++) { width[3 * n + k] = right[k] - left[k]; center[3 * n + k] = 0.5 * (right[k] + left[k]); }
    direct_of_node[n] = direct;
    int nd = (direct + 1) % 3;
    if (n >= P - 1) return;
    double tmp = right[direct]; right[direct] = split[n];
    center_top(P, nd, son[2 * n], son, split, left, right, center, width, direct_of_node);
    right[direct] = tmp;
    tmp = left[direct]; left[direct] = split[n];
    center_top(P, nd, son[2 * n + 1], son, split, left, right, center, width, direct_of_node);
    left[direct] = tmp;
}

void oracle_domain_setup(int nproc, double box, double* split, double* center, double* width, int* direct_of_node) {
    int P = nproc, len = 2 * P - 1;
    int* son = (int*)malloc(sizeof(int) * 2 * (size_t)len);
    double* tn = (double*)calloc((size_t)len, sizeof(double));
    double* tl = (double*)calloc((size_t)len, sizeof(double));
    double* tr = (double*)calloc((size_t)len, sizeof(double));
    for (int n = 0; n < P - 1; n++) { son[2 * n] = 2 * n + 1; son[2 * n + 1] = 2 * n + 2; split[n] = 0.0; }
    for (int n = P - 1; n < len; n++) { son[2 * n] = son[2 * n + 1] = -1; tn[n] = tl[n] = tr[n] = 1.0; split[n] = 0.0; }
    fill_time(P, 0, son, tn, tl, tr);
    double bl[3] = {0, 0, 0}, br[3] = {box, box, box};
    volume_part(P, 0, son, tl, tr, split, bl, br, 0);
    double l[3] = {0, 0, 0}, r[3] = {box, box, box};
    center_top(P, 0, 0, son, split, l, r, center, width, direct_of_node);
    free(son); free(tn); free(tl); free(tr);
}

/* I/src/domains.c:42-84 : ranks below the left / right son of a rank-tree node holding `size` ranks */
static void rank_fraction(int size, int* l, int* r) {
    int left, right;
    if (size == 1) { left = 1; right = 0; }
    else if (size == 2) { left = 1; right = 1; }
    else if (size == 3) { left = 2; right = 1; }
    else {
        left = 1; right = 2;
        while (size - left >= right - size) { left *= 2; right *= 2; }
        left >>= 1;
        right = size - left;
        if (left < right) { left = right; right = size - left; }
    }
    *l = left; *r = right;
}
/* I/src/domains.c:86-144 (determine_split_node): only the last assignment to `shift` (:120) is live */
static void relax_node(int P, int D, int nproc, int n, double* split, const double* tl, const double* tr,
                       const double bl[3], const double br[3]) {
    if (n >= P - 1) return;
    int nleft, nright;
    rank_fraction(nproc, &nleft, &nright);
    const double relax = 0.3;
    const double t1 = tl[n] / nleft, t2 = tr[n] / nright;
    const double w0l = split[n] - bl[D], w0r = br[D] - split[n];
    const double shift = 0.5 * relax * (t2 - t1) / (t1 * nleft / w0l + t2 * nright / w0r);
    const double new_split = split[n] + shift;
    double l[3] = {bl[0], bl[1], bl[2]}, r[3] = {br[0], br[1], br[2]};
    r[D] = split[n];
    relax_node(P, (D + 1) % 3, nleft, 2 * n + 1, split, tl, tr, l, r);
    l[D] = split[n]; r[D] = br[D];
    relax_node(P, (D + 1) % 3, nright, 2 * n + 2, split, tl, tr, l, r);
    split[n] = new_split;
}
/* I/src/domains.c:20-38 (measure_domain_runtime) + :146-157 (determine_split_domtree): frac[r] is rank r's
 * DTIME_FRACTION = W_r P / (sum W + 1e-4) (I/src/photoNs.c:303); split[2P-1] is updated in place */
void oracle_domain_relax(int nproc, double box, double* split, const double* frac) {
    if (nproc < 1 || nproc > (1 << 20)) return;
    const int P = nproc, len = 2 * P - 1, ml = mostleft_of(P);
    int* son = (int*)malloc(sizeof(int) * 2 * (size_t)len);
    double* tn = (double*)calloc((size_t)len, sizeof(double));
    double* tl = (double*)calloc((size_t)len, sizeof(double));
    double* tr = (double*)calloc((size_t)len, sizeof(double));
    for (int n = 0; n < P - 1; n++) { son[2 * n] = 2 * n + 1; son[2 * n + 1] = 2 * n + 2; }
    for (int n = P - 1; n < len; n++) { son[2 * n] = son[2 * n + 1] = -1; tl[n] = tr[n] = 1.0; }
    for (int r = 0; r < P; r++) {
        int idom = r + ml;
        if (idom > 2 * P - 2) idom -= P;
        tn[idom] = frac[r];
    }
    fill_time(P, 0, son, tn, tl, tr);
    const double bl[3] = {0, 0, 0}, br[3] = {box, box, box};
    relax_node(P, 0, P, 0, split, tl, tr, bl, br);
    free(son); free(tn); free(tl); free(tr);
}

typedef struct { double* pos; int64_t* payload; } Bodies;
static void swap_b(Bodies* b, int base, int i, int j) {
    if (i == j) return;
    double* p = b->pos + 3 * (size_t)base;
    for (int k = 0; k < 3; k++) { double t = p[3 * i + k]; p[3 * i + k] = p[3 * j + k]; p[3 * j + k] = t; }
    if (b->payload) { int64_t t = b->payload[base + i]; b->payload[base + i] = b->payload[base + j]; b->payload[base + j] = t; }
}
/* I/src/domains.c:163-270 */
static void split_at(Bodies* b, int D, int base, int length, int np[2], double split) {
    const double* p = b->pos + 3 * (size_t)base;
    if (length == 0) { np[0] = np[1] = 0; return; }
    if (length == 1) { if (p[D] > split) { np[0] = 0; np[1] = 1; } else { np[0] = 1; np[1] = 0; } return; }
    if (length == 2) {
        if (p[D] > p[3 + D]) swap_b(b, base, 0, 1);
        if (p[D] > split) { np[0] = 0; np[1] = 2; }
        else if (p[3 + D] <= split) { np[0] = 2; np[1] = 0; }
        else np[0] = np[1] = 1;
        return;
    }
    int top = 0;
    while (top < length && p[3 * top + D] <= split) top++;
    int but = length - 1;
    while (but >= 0 && p[3 * but + D] > split) but--;
    if (top == length) { np[0] = length; np[1] = 0; return; }
    if (but == -1) { np[1] = length; np[0] = 0; return; }
    int n;
    for (n = top; n <= but; n++) {
        if (p[3 * n + D] > split) {
            swap_b(b, base, n, but);
            while (p[3 * but + D] > split) but--;
        }
    }
    np[0] = (n == but) ? but + 1 : n;
    np[1] = length - np[0];
}
static void route_rec(Bodies* b, int P, int mostleft, const double* split, int D, int base, int length, int n, int* send) {
    /* I/src/domains.c:272-296 */
    if (n >= P - 1) { send[(n - mostleft + P) % P] = length; return; }
    int np[2];
    split_at(b, D, base, length, np, split[n]);
    route_rec(b, P, mostleft, split, (D + 1) % 3, base, np[0], 2 * n + 1, send);
    route_rec(b, P, mostleft, split, (D + 1) % 3, base + np[0], np[1], 2 * n + 2, send);
}
void oracle_domain_partition(int nproc, const double* split, double* pos, int64_t* payload, int npart, int* sendcount,
                             int* sendorder) {
    (void)sendorder;
    Bodies b = {pos, payload};
    for (int r = 0; r < nproc; r++) sendcount[r] = 0;
    route_rec(&b, nproc, mostleft_of(nproc), split, 0, 0, npart, 0, sendcount);
}


/* ------------------------------------------------------------------ mid-field ----
 * P2M / M2M / M2L / L2L / L2P restated from 1_Indexing/src/operator.c (QUADRUPOLE + OCTUPOLE: 20 coefficients in the
 * order of 1_Indexing/inc/operator.h:24-67), driven as fmm_prepare / fmm_task / fmm_ext drive them
 * (1_Indexing/src/fmm.c:745-790, 562-705, 913-945, 1026-1145; 1_Indexing/src/remotes.c:477-640).
 * A coefficient with multi-index (a,b,c) carries 1/(a! b! c!); the expansions are written over that table. */
enum { NMUL = 20 };
static const int MIDX[NMUL][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{2,0,0},{1,1,0},{1,0,1},{0,2,0},{0,1,1},{0,0,2},
                                  {3,0,0},{2,1,0},{2,0,1},{1,2,0},{1,1,1},{1,0,2},{0,3,0},{0,2,1},{0,1,2},{0,0,3}};
static int mpos(int a, int b, int c) {
    for (int n = 0; n < NMUL; n++) if (MIDX[n][0] == a && MIDX[n][1] == b && MIDX[n][2] == c) return n;
    return -1;
}
static double mfact(int n) { return n <= 1 ? 1.0 : (n == 2 ? 2.0 : 6.0); }
static void mpowers(double x, double y, double z, double pw[NMUL]) {
    for (int n = 0; n < NMUL; n++) {
        double v = 1.0;
        for (int k = 0; k < MIDX[n][0]; k++) v *= x;
        for (int k = 0; k < MIDX[n][1]; k++) v *= y;
        for (int k = 0; k < MIDX[n][2]; k++) v *= z;
        pw[n] = v / (mfact(MIDX[n][0]) * mfact(MIDX[n][1]) * mfact(MIDX[n][2]));
    }
}
/* operator.c:13-93 */
static void o_p2m(const double* pos, int ipart, int npart, const double c[3], double mass, double M[NMUL]) {
    for (int n = 0; n < NMUL; n++) M[n] = 0.0;
    for (int p = ipart; p < ipart + npart; p++) {
        double pw[NMUL];
        mpowers(pos[3 * (size_t)p] - c[0], pos[3 * (size_t)p + 1] - c[1], pos[3 * (size_t)p + 2] - c[2], pw);
        for (int n = 0; n < NMUL; n++) M[n] += ((MIDX[n][0] + MIDX[n][1] + MIDX[n][2]) & 1) ? -mass * pw[n] : mass * pw[n];
    }
}
/* operator.c:96-160: tM += M shifted by d = new centre - old centre */
static void o_m2m(const double d[3], const double M[NMUL], double tM[NMUL]) {
    double pw[NMUL];
    mpowers(d[0], d[1], d[2], pw);
    for (int n = 0; n < NMUL; n++) {
        double s = 0.0;
        for (int k = 0; k < NMUL; k++)
            if (MIDX[k][0] <= MIDX[n][0] && MIDX[k][1] <= MIDX[n][1] && MIDX[k][2] <= MIDX[n][2])
                s += M[k] * pw[mpos(MIDX[n][0] - MIDX[k][0], MIDX[n][1] - MIDX[k][1], MIDX[n][2] - MIDX[k][2])];
        tM[n] += s;
    }
}
/* operator.c:255-392 with the LONGSHORT factors :296-305 (rs <= 0: plain 1/r) */
static void o_m2l(const double x[3], const double M[NMUL], double rs, double toL[NMUL]) {
    const double r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2], dr = sqrt(r2);
    const double ir = 1.0 / dr, ir2 = ir * ir, ir3 = ir2 * ir, ir4 = ir3 * ir, ir5 = ir4 * ir, ir6 = ir5 * ir, ir7 = ir6 * ir;
    double f[4];
    if (rs > 0.0) {
        const double irs = 1.0 / rs, irs2 = irs * irs, irs3 = irs2 * irs, irs5 = irs3 * irs2;
        const double u = 0.5 * dr / rs, fe = exp(-u * u) * (1.0 / sqrt(M_PI)), fc = erfc(u);
        f[0] = ir * fc;
        f[1] = -ir3 * (fc + dr * fe * irs);
        f[2] = 3.0 * ir5 * fc + (3.0 * irs * ir4 + 0.5 * ir2 * irs3) * fe;
        f[3] = -3.0 * 5.0 * ir7 * fc - (15.0 * ir6 * irs + 2.5 * ir4 * irs3 + 0.25 * ir2 * irs5) * fe;
    } else { f[0] = ir; f[1] = -ir3; f[2] = 3.0 * ir5; f[3] = -15.0 * ir7; }
    double D[NMUL];
    D[0] = f[0];
    for (int n = 1; n < NMUL; n++) {
        int ax[3] = {0, 0, 0}, m = 0;
        for (int d = 0; d < 3; d++) for (int k = 0; k < MIDX[n][d]; k++) ax[m++] = d;
        if (m == 1) D[n] = f[1] * x[ax[0]];
        else if (m == 2) D[n] = f[2] * x[ax[0]] * x[ax[1]] + (ax[0] == ax[1] ? f[1] : 0.0);
        else {
            double t = 0.0;
            if (ax[1] == ax[2]) t += x[ax[0]];
            if (ax[0] == ax[2]) t += x[ax[1]];
            if (ax[0] == ax[1]) t += x[ax[2]];
            D[n] = f[3] * x[ax[0]] * x[ax[1]] * x[ax[2]] + f[2] * t;
        }
    }
    for (int n = 0; n < NMUL; n++) {
        double a = 0.0;
        for (int k = 0; k < NMUL; k++) {
            const int o = MIDX[n][0] + MIDX[n][1] + MIDX[n][2] + MIDX[k][0] + MIDX[k][1] + MIDX[k][2];
            if (o <= 3) a += M[k] * D[mpos(MIDX[n][0] + MIDX[k][0], MIDX[n][1] + MIDX[k][1], MIDX[n][2] + MIDX[k][2])];
        }
        toL[n] += a;
    }
}
/* operator.c:395-494: toL += L shifted by d = new centre - old centre */
static void o_l2l(const double d[3], const double L[NMUL], double toL[NMUL]) {
    double pw[NMUL];
    mpowers(d[0], d[1], d[2], pw);
    for (int n = 0; n < NMUL; n++) {
        double s = 0.0;
        for (int k = 0; k < NMUL; k++)
            if (MIDX[k][0] >= MIDX[n][0] && MIDX[k][1] >= MIDX[n][1] && MIDX[k][2] >= MIDX[n][2])
                s += L[k] * pw[mpos(MIDX[k][0] - MIDX[n][0], MIDX[k][1] - MIDX[n][1], MIDX[k][2] - MIDX[n][2])];
        toL[n] += s;
    }
}

typedef struct {
    Walk w;                       /* tree arrays, theta, rcut (reference numbering) */
    double rs;
    double *lM, *nM, *lL, *nL;    /* [leaf][20], [node][20] */
    int64_t nm2l;
    /* received image */
    const int* r_src;
} Mid;
static double* mid_M(Mid* m, int id) { return id < m->w.first_node ? m->lM + NMUL * (size_t)(id - m->w.first_leaf) : m->nM + NMUL * (size_t)(id - m->w.first_node); }
static double* mid_L(Mid* m, int id) { return id < m->w.first_node ? m->lL + NMUL * (size_t)(id - m->w.first_leaf) : m->nL + NMUL * (size_t)(id - m->w.first_node); }

/* operator.c:165-194 */
static void mid_m2m_rec(Mid* m, int inode) {
    const double* cn = box_c(&m->w, inode);
    for (int n = 0; n < 2; n++) {
        const int idx = son_of(&m->w, inode, n);
        if (idx < 0) continue;
        if (idx >= m->w.first_node) mid_m2m_rec(m, idx);
        const double* cs = box_c(&m->w, idx);
        const double d[3] = {cn[0] - cs[0], cn[1] - cs[1], cn[2] - cs[2]};
        o_m2m(d, mid_M(m, idx), mid_M(m, inode));
    }
}
/* fmm.c:562-705 + task_compute_m2l :913-945 (tasks are applied in the order they are found) */
static void mid_walk_local(Mid* m, int im, int jm) {
    Walk* w = &m->w;
    if (im == -1 || jm == -1) return;
    if (im == jm) {
        if (im >= w->first_node)
            for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) mid_walk_local(m, son_of(w, im, a), son_of(w, jm, b));
        return;
    }
    const int ileaf = im < w->first_node, jleaf = jm < w->first_node;
    if (ileaf && jleaf) return;
    const double *ci = box_c(w, im), *cj = box_c(w, jm), *wi = box_w(w, im), *wj = box_w(w, jm);
    const double dist[3] = {ci[0] - cj[0], ci[1] - cj[1], ci[2] - cj[2]};
    const int flag = oracle_acceptance(wi, wj, dist, w->theta, w->rcut);
    if (flag == 1) { o_m2l(dist, mid_M(m, jm), m->rs, mid_L(m, im)); m->nm2l++; return; }
    if (flag != 0) return;
    if (ileaf) { mid_walk_local(m, im, son_of(w, jm, 0)); mid_walk_local(m, im, son_of(w, jm, 1)); return; }
    if (jleaf) { mid_walk_local(m, son_of(w, im, 0), jm); mid_walk_local(m, son_of(w, im, 1), jm); return; }
    if (wi[0] + wi[1] + wi[2] > wj[0] + wj[1] + wj[2]) { mid_walk_local(m, son_of(w, im, 0), jm); mid_walk_local(m, son_of(w, im, 1), jm); }
    else { mid_walk_local(m, im, son_of(w, jm, 0)); mid_walk_local(m, im, son_of(w, jm, 1)); }
}
/* remotes.c:477-640 + task_compute_m2l_ext: the image node's multipole is that of the local node it was copied from */
static void mid_walk_ext(Mid* m, int im, int jm) {
    Walk* w = &m->w;
    const int ileaf = im < w->first_node, jleaf = w->r_npart[jm] <= w->maxleaf;
    if (ileaf && jleaf) return;
    const double* ci = box_c(w, im);
    const double* cj = w->r_center + 3 * (size_t)jm;
    const double dist[3] = {ci[0] - cj[0], ci[1] - cj[1], ci[2] - cj[2]};
    const double *wi = box_w(w, im), *wj = w->r_width + 3 * (size_t)jm;
    const int flag = oracle_acceptance(wi, wj, dist, w->theta, w->rcut);
    const int s0 = w->r_son[2 * (size_t)jm], s1 = w->r_son[2 * (size_t)jm + 1];
    if (flag == -1) return;
    if (ileaf) {
        if (flag == 1 || s0 < 0 || s1 < 0) { o_m2l(dist, mid_M(m, m->r_src[jm]), m->rs, mid_L(m, im)); m->nm2l++; return; }
        mid_walk_ext(m, im, s0); mid_walk_ext(m, im, s1);
        return;
    }
    if (flag == 1) { o_m2l(dist, mid_M(m, m->r_src[jm]), m->rs, mid_L(m, im)); m->nm2l++; return; }
    if (jleaf) { mid_walk_ext(m, son_of(w, im, 0), jm); mid_walk_ext(m, son_of(w, im, 1), jm); return; }
    if (wi[0] + wi[1] + wi[2] > wj[0] + wj[1] + wj[2] || s0 < 0 || s1 < 0) { mid_walk_ext(m, son_of(w, im, 0), jm); mid_walk_ext(m, son_of(w, im, 1), jm); }
    else { mid_walk_ext(m, im, s0); mid_walk_ext(m, im, s1); }
}
/* operator.c:498-530 */
static void mid_l2l_rec(Mid* m, int inode) {
    if (inode < m->w.first_node) return;
    const double* cn = box_c(&m->w, inode);
    for (int n = 0; n < 2; n++) {
        const int idx = son_of(&m->w, inode, n);
        if (idx < m->w.first_leaf) return;
        const double* cs = box_c(&m->w, idx);
        const double d[3] = {cs[0] - cn[0], cs[1] - cn[1], cs[2] - cn[2]};
        o_l2l(d, mid_L(m, inode), mid_L(m, idx));
        if (idx >= m->w.first_node) mid_l2l_rec(m, idx);
    }
}

/* Single rank: local tree + (box > 0) the 26 periodic images pruned against the root box; literal_d6 != 0 also replays
 * the reference's zero-shift self exchange (SURVEY defect D6).  Outputs: leaf_M[nleaf][20], node_M[nnode][20],
 * leaf_L[nleaf][20] (after L2L), acc[npart][3] (L2P, accumulated), *nm2l_local / *nm2l_total. */
int oracle_midfield(int npart, int nleaf_cap, int nleaf, int nnode, int maxleaf, const double* pos, const int* leaf_npart,
                    const int* leaf_ipart, const double* leaf_center, const double* leaf_width, const int* node_npart,
                    const int* node_son, const double* node_center, const double* node_width, double theta, double rcut,
                    double rs, double mass, double box, int literal_d6, double* leaf_M, double* node_M, double* leaf_L,
                    double* acc, int64_t* nm2l_local, int64_t* nm2l_total) {
    Mid m;
    memset(&m, 0, sizeof m);
    Walk* w = &m.w;
    w->first_leaf = npart; w->last_leaf = npart + nleaf; w->first_node = npart + nleaf_cap;
    w->lc = leaf_center; w->lw = leaf_width; w->nc = node_center; w->nw = node_width; w->son = node_son;
    w->theta = theta; w->rcut = rcut; w->maxleaf = maxleaf;
    m.rs = rs; m.lM = leaf_M; m.nM = node_M; m.lL = leaf_L;
    m.nL = (double*)calloc((size_t)NMUL * (size_t)(nnode > 0 ? nnode : 1), sizeof(double));
    if (!m.nL) return -1;
    memset(leaf_L, 0, sizeof(double) * NMUL * (size_t)nleaf);
    memset(node_M, 0, sizeof(double) * NMUL * (size_t)nnode);
    for (int l = 0; l < nleaf; l++) o_p2m(pos, leaf_ipart[l], leaf_npart[l], leaf_center + 3 * (size_t)l, mass, leaf_M + NMUL * (size_t)l);
    mid_m2m_rec(&m, w->first_node);
    mid_walk_local(&m, w->first_node, w->first_node);
    if (nm2l_local) *nm2l_local = m.nm2l;
    if (box > 0.0) {
        const int cap_node = nleaf + nnode + 2, cap_body = npart + 1;
        int* r_npart = (int*)malloc(sizeof(int) * (size_t)cap_node);
        int* r_son = (int*)malloc(sizeof(int) * 2 * (size_t)cap_node);
        int* r_src = (int*)malloc(sizeof(int) * (size_t)cap_node);
        double* r_center = (double*)malloc(sizeof(double) * 3 * (size_t)cap_node);
        double* r_width = (double*)malloc(sizeof(double) * 3 * (size_t)cap_node);
        double* r_body = (double*)malloc(sizeof(double) * 3 * (size_t)cap_body);
        if (!r_npart || !r_son || !r_src || !r_center || !r_width || !r_body) return -1;
        for (int si = literal_d6 ? 0 : 1; si < 27; si++) {
            /* order of 1_Indexing/src/fmm.c:1064-1106: zero displacement, then mi, mj, mk in {-1,0,1} */
            int sh[3] = {0, 0, 0};
            if (si > 0) { int q = si - 1; if (q >= 13) q++; sh[0] = q / 9 - 1; sh[1] = (q / 3) % 3 - 1; sh[2] = q % 3 - 1; }
            const double disp[3] = {sh[0] * box, sh[1] * box, sh[2] * box};
            Prune p;
            memset(&p, 0, sizeof p);
            p.first_leaf = npart; p.first_node = npart + nleaf_cap;
            p.pos = pos; p.leaf_npart = leaf_npart; p.leaf_ipart = leaf_ipart; p.node_npart = node_npart; p.node_son = node_son;
            p.lc = leaf_center; p.lw = leaf_width; p.nc = node_center; p.nw = node_width;
            p.tc = node_center; p.tw = node_width;           /* the local root box (toptree.c:18-45) */
            p.disp = disp; p.theta = theta; p.rcut = rcut; p.cap_node = cap_node; p.cap_body = cap_body;
            p.r_npart = r_npart; p.r_son = r_son; p.r_center = r_center; p.r_width = r_width; p.r_body = r_body; p.r_src = r_src;
            prune_rec(&p, 0, p.first_node);
            if (p.overflow) return -1;
            w->r_npart = r_npart; w->r_son = r_son; w->r_center = r_center; w->r_width = r_width; m.r_src = r_src;
            if (p.numnode > 0) mid_walk_ext(&m, w->first_node, 0);
        }
        free(r_npart); free(r_son); free(r_src); free(r_center); free(r_width); free(r_body);
    }
    if (nm2l_total) *nm2l_total = m.nm2l;
    mid_l2l_rec(&m, w->first_node);
    /* operator.c:197-251 */
    for (int l = 0; l < nleaf; l++) {
        const double* F = leaf_L + NMUL * (size_t)l;
        for (int p = leaf_ipart[l]; p < leaf_ipart[l] + leaf_npart[l]; p++) {
            double pw[NMUL];
            mpowers(pos[3 * (size_t)p] - leaf_center[3 * (size_t)l], pos[3 * (size_t)p + 1] - leaf_center[3 * (size_t)l + 1],
                    pos[3 * (size_t)p + 2] - leaf_center[3 * (size_t)l + 2], pw);
            for (int n = 0; n < NMUL; n++) {
                if (MIDX[n][0] + MIDX[n][1] + MIDX[n][2] > 2) continue;
                acc[3 * (size_t)p] += F[mpos(MIDX[n][0] + 1, MIDX[n][1], MIDX[n][2])] * pw[n];
                acc[3 * (size_t)p + 1] += F[mpos(MIDX[n][0], MIDX[n][1] + 1, MIDX[n][2])] * pw[n];
                acc[3 * (size_t)p + 2] += F[mpos(MIDX[n][0], MIDX[n][1], MIDX[n][2] + 1)] * pw[n];
            }
        }
    }
    free(m.nL);
    return 0;
}
