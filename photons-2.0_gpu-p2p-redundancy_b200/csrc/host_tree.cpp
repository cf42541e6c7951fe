// host_tree.cpp -- host-side producers of the P2P path (include/p2p_host.h), multi-threaded.
//
// Semantics follow the reference exactly (bit-identical fp64 results; compiled with
// -ffp-contract=off): see the per-function citations.  Structure is our own: the tree is built by
// subtree-parallel tasks into a scratch pool and renumbered afterwards to the reference's ids; the
// dual-tree walks expand a frontier of node pairs in traversal order and then run the pairs on all
// host threads, concatenating the per-pair outputs in frontier order, which reproduces the
// reference's sequential emission order.
#include "../../include/p2p_host.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

struct p2p_tree {
    int npart = 0, maxleaf = 0, nleaf = 0, nnode = 0, nleaf_cap = 0, nnode_cap = 0, first_leaf = 0, first_node = 0;
    std::vector<int> leaf_npart, leaf_ipart, node_npart, node_son;
    std::vector<double> leaf_center, leaf_width, node_split, node_center, node_width;
};

namespace {

int resolve_threads(int n) {
#ifdef _OPENMP
    return n > 0 ? n : omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

// ------------------------------------------------------------------ record permutation --------
struct Records {
    double* base;
    int64_t stride;
    int64_t* payload;
    inline double key(int64_t i, int D) const { return base[i * stride + D]; }
    inline void swap(int64_t i, int64_t j) const {
        if (i == j) return;
        double* a = base + i * stride;
        double* b = base + j * stride;
        for (int64_t k = 0; k < stride; k++) { double t = a[k]; a[k] = b[k]; b[k] = t; }
        if (payload) { int64_t t = payload[i]; payload[i] = payload[j]; payload[j] = t; }
    }
};

// bksort_inplace, 1_Indexing/src/fmm.c:29-77.  Runs of 2 split at the midpoint; longer runs at the
// SEQUENTIAL fp64 mean; the left count is the final position of the retreating cursor.
void split_at_mean(const Records& R, int D, int64_t lo, int len, int cnt[2], double* split) {
    cnt[0] = cnt[1] = 0;
    if (len < 2) { cnt[1] = len; return; }
    if (len == 2) {
        cnt[0] = cnt[1] = 1;
        *split = 0.5 * (R.key(lo, D) + R.key(lo + 1, D));
        if (R.key(lo, D) > R.key(lo + 1, D)) R.swap(lo, lo + 1);
        return;
    }
    double mean = 0.0;
    for (int n = 0; n < len; n++) mean += R.key(lo + n, D);
    mean /= (double)len;
    int hi = len - 1;
    for (int n = 0; n < hi; n++) {
        if (R.key(lo + n, D) > mean) {
            while (R.key(lo + hi, D) > mean && hi > n) hi--;
            R.swap(lo + n, lo + hi);
        }
    }
    cnt[0] = hi;
    cnt[1] = len - hi;
    *split = mean;
}

// ------------------------------------------------------------------ tree build -----------------
struct Scratch {                 // one internal node before renumbering
    int npart;
    double split;
    int kid[2];                  // scratch index of the child node, or -1 when the child is a leaf
    int leaf_n[2], leaf_ip[2];   // leaf children: occupancy and first particle
};

struct Builder {
    Records R;
    int maxleaf;
    std::vector<Scratch> pool;
    std::atomic<int> used{0};
    std::atomic<int> overflow{0};

    void grow(int D, int64_t ipart, int len, int me) {
        Scratch& s = pool[me];
        s.npart = len;
        int cnt[2];
        double split = 0.0;
        split_at_mean(R, D, ipart, len, cnt, &split);
        s.split = split;
        int64_t ip = ipart;
        int kids[2] = {-1, -1};
        int64_t kip[2] = {0, 0};
        for (int n = 0; n < 2; n++) {
            s.kid[n] = -1;
            s.leaf_n[n] = cnt[n];
            s.leaf_ip[n] = (int)ip;
            if (cnt[n] > maxleaf) {
                int k = used.fetch_add(1);
                if (k >= (int)pool.size()) { overflow.store(1); return; }
                s.kid[n] = k;
                kids[n] = k;
                kip[n] = ip;
            }
            ip += cnt[n];
        }
        const int nd = (D + 1) % 3;
        for (int n = 0; n < 2; n++) {
            if (kids[n] < 0) continue;
            const int k = kids[n], l = cnt[n];
            const int64_t p = kip[n];
#pragma omp task default(shared) firstprivate(nd, p, l, k) if (l > 8192)
            grow(nd, p, l, k);
        }
    }
};

// ------------------------------------------------------------------ MAC -------------------------
// acceptance(), 1_Indexing/src/fmm.c:266-325 with LONGSHORT: 0 open, 1 accept (M2L), -1 abort.
inline int mac(const double* wi, const double* wj, const double* dist, double theta, double rcut) {
    const double w0 = (wi[0] + wj[0]) * 0.5, w1 = (wi[1] + wj[1]) * 0.5, w2 = (wi[2] + wj[2]) * 0.5;
    const double dd2 = dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2];
    double g0 = dist[0], g1 = dist[1], g2 = dist[2];
    if (g0 < 0.0) g0 = -g0;
    if (g1 < 0.0) g1 = -g1;
    if (g2 < 0.0) g2 = -g2;
    g0 -= w0; g1 -= w1; g2 -= w2;
    if (g0 <= 0.0) g0 = 0.0;
    if (g1 <= 0.0) g1 = 0.0;
    if (g2 <= 0.0) g2 = 0.0;
    if (g0 + g1 + g2 < 0.0001) return 0;
    const double dm2 = g0 * g0 + g1 * g1 + g2 * g2;
    const double c2 = rcut * rcut;
    if (dm2 >= c2) return -1;
    if (dd2 > 1.0 * c2) return 0;
    double wmax = w0;
    if (w1 > wmax) wmax = w1;
    if (w2 > wmax) wmax = w2;
    wmax *= 2;
    return (wmax * wmax < theta * theta * dd2) ? 1 : 0;
}

// ------------------------------------------------------------------ walks -----------------------
struct Pair { int im, jm; };

struct LocalWalk {
    const p2p_tree* T;
    double theta, rcut;
    inline bool is_leaf(int id) const { return id < T->first_node; }
    inline const double* cen(int id) const {
        return is_leaf(id) ? &T->leaf_center[3 * (size_t)(id - T->first_leaf)] : &T->node_center[3 * (size_t)(id - T->first_node)];
    }
    inline const double* wid(int id) const {
        return is_leaf(id) ? &T->leaf_width[3 * (size_t)(id - T->first_leaf)] : &T->node_width[3 * (size_t)(id - T->first_node)];
    }
    inline int son(int id, int n) const { return T->node_son[2 * (size_t)(id - T->first_node) + n]; }

    // One step of walk_task_p2p (1_Indexing/src/fmm.c:402-534): either the pair is emitted, or it is
    // replaced by its child calls in the reference's call order.  Returns the number of children.
    inline int step(Pair p, bool* emit, Pair kids[4]) const {
        *emit = false;
        const int im = p.im, jm = p.jm;
        if (im == -1 || jm == -1) return 0;
        if (im == jm) {
            if (is_leaf(im)) { *emit = true; return 0; }
            int k = 0;
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++) kids[k++] = Pair{son(im, a), son(jm, b)};
            return 4;
        }
        const bool il = is_leaf(im), jl = is_leaf(jm);
        if (il && jl) { *emit = true; return 0; }
        const double *ci = cen(im), *cj = cen(jm), *wi = wid(im), *wj = wid(jm);
        const double dist[3] = {ci[0] - cj[0], ci[1] - cj[1], ci[2] - cj[2]};
        if (mac(wi, wj, dist, theta, rcut) != 0) return 0;
        bool open_i;
        if (il) open_i = false;
        else if (jl) open_i = true;
        else open_i = wi[0] + wi[1] + wi[2] > wj[0] + wj[1] + wj[2];
        if (open_i) { kids[0] = Pair{son(im, 0), jm}; kids[1] = Pair{son(im, 1), jm}; }
        else { kids[0] = Pair{im, son(jm, 0)}; kids[1] = Pair{im, son(jm, 1)}; }
        return 2;
    }
    void run(Pair root, std::vector<int>& out) const {
        std::vector<Pair> stack;
        stack.push_back(root);
        Pair kids[4];
        while (!stack.empty()) {
            Pair p = stack.back();
            stack.pop_back();
            bool emit;
            int nk = step(p, &emit, kids);
            if (emit) { out.push_back(p.im - T->first_leaf); out.push_back(p.jm - T->first_leaf); }
            for (int k = nk - 1; k >= 0; k--) stack.push_back(kids[k]);  // reversed: pop order = call order
        }
    }
};

struct ExtWalk {
    const p2p_tree* T;
    const p2p_image* I;
    double theta, rcut;
    inline bool is_leaf(int id) const { return id < T->first_node; }
    inline const double* cen(int id) const {
        return is_leaf(id) ? &T->leaf_center[3 * (size_t)(id - T->first_leaf)] : &T->node_center[3 * (size_t)(id - T->first_node)];
    }
    inline const double* wid(int id) const {
        return is_leaf(id) ? &T->leaf_width[3 * (size_t)(id - T->first_leaf)] : &T->node_width[3 * (size_t)(id - T->first_node)];
    }
    inline int son(int id, int n) const { return T->node_son[2 * (size_t)(id - T->first_node) + n]; }
    // walk_task_p2p_ext, 1_Indexing/src/remotes.c:141-317; a received node is a leaf iff npart <= MAXLEAF
    inline int step(Pair p, bool* emit, Pair kids[2]) const {
        *emit = false;
        const int im = p.im, jm = p.jm;
        const bool il = is_leaf(im), jl = I->npart[jm] <= T->maxleaf;
        if (il && jl) { *emit = true; return 0; }
        const double *ci = cen(im), *cj = I->center + 3 * (size_t)jm, *wi = wid(im), *wj = I->width + 3 * (size_t)jm;
        const double dist[3] = {ci[0] - cj[0], ci[1] - cj[1], ci[2] - cj[2]};
        const int flag = mac(wi, wj, dist, theta, rcut);
        const int s0 = I->son[2 * (size_t)jm], s1 = I->son[2 * (size_t)jm + 1];
        if (flag != 0) return 0;
        bool open_i;
        if (il) { if (s0 < 0 || s1 < 0) return 0; open_i = false; }
        else if (jl) open_i = true;
        else open_i = (wi[0] + wi[1] + wi[2] > wj[0] + wj[1] + wj[2]) || s0 < 0 || s1 < 0;
        if (open_i) { kids[0] = Pair{son(im, 0), jm}; kids[1] = Pair{son(im, 1), jm}; }
        else { kids[0] = Pair{im, s0}; kids[1] = Pair{im, s1}; }
        return 2;
    }
    void run(Pair root, std::vector<int>& out) const {
        std::vector<Pair> stack;
        stack.push_back(root);
        Pair kids[2];
        while (!stack.empty()) {
            Pair p = stack.back();
            stack.pop_back();
            bool emit;
            int nk = step(p, &emit, kids);
            if (emit) { out.push_back(p.im - T->first_leaf); out.push_back(p.jm); }
            for (int k = nk - 1; k >= 0; k--) stack.push_back(kids[k]);
        }
    }
};

struct Item { Pair p; bool emit; };

// Runs a list of frontier items on all host threads and concatenates the per-item outputs in list order.
template <class W>
int run_items(const W& w, const std::vector<Item>& cur, int nthreads, bool ext, int** tt, int** ts, int64_t* ntask,
              int* tt_into = nullptr, int* ts_into = nullptr, int64_t cap = 0) {
    const int64_t nitem = (int64_t)cur.size();
    std::vector<std::vector<int>> outs((size_t)nitem);
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
    for (int64_t i = 0; i < nitem; i++) {
        if (cur[i].emit) continue;
        w.run(cur[i].p, outs[i]);
    }
    std::vector<int64_t> off((size_t)nitem + 1, 0);
    for (int64_t i = 0; i < nitem; i++) off[i + 1] = off[i] + (cur[i].emit ? 1 : (int64_t)outs[i].size() / 2);
    const int64_t total = off[nitem];
    int *t, *s;
    if (tt_into) {
        *ntask = total;
        if (total > cap) return -3;
        t = tt_into; s = ts_into;
    } else {
        t = (int*)malloc(sizeof(int) * (size_t)(total ? total : 1));
        s = (int*)malloc(sizeof(int) * (size_t)(total ? total : 1));
        if (!t || !s) { free(t); free(s); return -1; }
    }
    const int first_leaf = w.T->first_leaf;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (int64_t i = 0; i < nitem; i++) {
        int64_t o = off[i];
        if (cur[i].emit) {
            t[o] = cur[i].p.im - first_leaf;
            s[o] = ext ? cur[i].p.jm : cur[i].p.jm - first_leaf;
            continue;
        }
        const std::vector<int>& v = outs[i];
        for (size_t k = 0; k + 1 < v.size(); k += 2) { t[o] = v[k]; s[o] = v[k + 1]; o++; }
    }
    if (!tt_into) { *tt = t; *ts = s; }
    *ntask = total;
    return 0;
}

// Expands the open items of `cur` (in traversal order) until there are at least `want` items.
template <class W, int MAXK>
void expand_items(const W& w, std::vector<Item>& cur, size_t want) {
    std::vector<Item> nxt;
    for (int round = 0; round < 64; round++) {
        size_t open = 0;
        for (const Item& it : cur) open += it.emit ? 0 : 1;
        if (open == 0 || cur.size() >= want) break;
        nxt.clear();
        nxt.reserve(cur.size() * 3);
        for (const Item& it : cur) {
            if (it.emit) { nxt.push_back(it); continue; }
            bool emit;
            Pair kids[MAXK];
            int nk = w.step(it.p, &emit, kids);
            if (emit) nxt.push_back(Item{it.p, true});
            for (int k = 0; k < nk; k++) nxt.push_back(Item{kids[k], false});
        }
        cur.swap(nxt);
    }
}

// Frontier-parallel driver shared by both walks.
template <class W, int MAXK>
int drive_walk(const W& w, Pair root, int nthreads, int** tt, int** ts, int64_t* ntask) {
    std::vector<Item> cur;
    cur.push_back(Item{root, false});
    expand_items<W, MAXK>(w, cur, (size_t)nthreads * 256);
    return run_items(w, cur, nthreads, MAXK == 2, tt, ts, ntask);
}

int mostleft_of(int P) {  // 1_Indexing/src/initial.c:206-215
    int m = 1;
    while (m < 2 * P - 1) m *= 2;
    m = m / 2 - 1;
    return P == 1 ? 0 : m;
}

}  // namespace

extern "C" {

int p2p_host_max_threads(void) { return resolve_threads(0); }
void p2p_host_free(void* p) { free(p); }

int p2p_build_localtree(p2p_tree** out, double* records, int64_t stride, int64_t* payload, int npart, int maxleaf,
                        const double bdl[3], const double bdr[3], int direct_start, int nthreads) {
    if (!out || npart < 0 || maxleaf < 1 || stride < 3 || (npart && !records)) return -2;
    nthreads = resolve_threads(nthreads);
    p2p_tree* T = new p2p_tree();
    T->npart = npart; T->maxleaf = maxleaf;
    // capacities and id bases, 1_Indexing/src/fmm.c:203-212
    int cap = (int)(2.0 * ((double)npart) / ((double)maxleaf));
    if (cap > npart) cap = npart + 1;
    T->nleaf_cap = T->nnode_cap = cap;
    T->first_leaf = npart;
    T->first_node = npart + cap;
    if (npart == 0 || cap < 1) { *out = T; return 0; }

    Builder B;
    B.R = Records{records, stride, payload};
    B.maxleaf = maxleaf;
    B.pool.resize((size_t)cap);
    B.used.store(1);
#pragma omp parallel num_threads(nthreads)
    {
#pragma omp single
        B.grow(direct_start, 0, npart, 0);
    }
    if (B.overflow.load()) { delete T; return -1; }
    const int nnode = B.used.load();

    // renumber to the reference's ids (nodes in pre-order, leaves left to right) and cut the boxes
    // (center_kdtree, 1_Indexing/src/fmm.c:120-174)
    T->node_npart.assign((size_t)nnode, 0); T->node_son.assign(2 * (size_t)nnode, -1); T->node_split.assign((size_t)nnode, 0.0);
    T->node_center.assign(3 * (size_t)nnode, 0.0); T->node_width.assign(3 * (size_t)nnode, 0.0);
    T->leaf_npart.reserve((size_t)cap); T->leaf_ipart.reserve((size_t)cap);
    T->leaf_center.reserve(3 * (size_t)cap); T->leaf_width.reserve(3 * (size_t)cap);
    struct Frame { int scratch, id, direct, next; double l[3], r[3]; };
    std::vector<Frame> st;
    int next_node = 0, nleaf = 0;
    Frame f0{0, next_node++, direct_start, 0, {bdl[0], bdl[1], bdl[2]}, {bdr[0], bdr[1], bdr[2]}};
    st.push_back(f0);
    while (!st.empty()) {
        Frame& f = st.back();
        const Scratch& s = B.pool[(size_t)f.scratch];
        if (f.next == 0) {
            T->node_npart[(size_t)f.id] = s.npart;
            T->node_split[(size_t)f.id] = s.split;
            for (int k = 0; k < 3; k++) {
                T->node_width[3 * (size_t)f.id + k] = f.r[k] - f.l[k];
                T->node_center[3 * (size_t)f.id + k] = 0.5 * (f.r[k] + f.l[k]);
            }
        }
        if (f.next == 2) { st.pop_back(); continue; }
        const int n = f.next++;
        const int D = f.direct;
        if (s.kid[n] < 0) {
            if (nleaf >= cap) { delete T; return -1; }
            T->node_son[2 * (size_t)f.id + n] = T->first_leaf + nleaf;
            T->leaf_npart.push_back(s.leaf_n[n]);
            T->leaf_ipart.push_back(s.leaf_ip[n]);
            for (int k = 0; k < 3; k++) {
                double w = T->node_width[3 * (size_t)f.id + k], c = T->node_center[3 * (size_t)f.id + k];
                if (k == D) {
                    if (n == 0) { w = s.split - f.l[D]; c = 0.5 * (f.l[D] + s.split); }
                    else { w = f.r[D] - s.split; c = 0.5 * (f.r[D] + s.split); }
                }
                T->leaf_width.push_back(w);
                T->leaf_center.push_back(c);
            }
            nleaf++;
        } else {
            Frame g{s.kid[n], next_node++, (D + 1) % 3, 0, {f.l[0], f.l[1], f.l[2]}, {f.r[0], f.r[1], f.r[2]}};
            if (n == 0) g.r[D] = s.split; else g.l[D] = s.split;
            T->node_son[2 * (size_t)f.id + n] = T->first_node + g.id;
            st.push_back(g);  // invalidates f; loop re-reads st.back()
        }
    }
    // leaf layout needs interleaved [leaf][3]; the pushes above already are
    T->nleaf = nleaf;
    T->nnode = nnode;
    *out = T;
    return 0;
}

void p2p_tree_free(p2p_tree* t) { delete t; }

int p2p_tree_get(const p2p_tree* t, p2p_tree_view* v) {
    if (!t || !v) return -2;
    v->npart = t->npart; v->maxleaf = t->maxleaf; v->nleaf = t->nleaf; v->nnode = t->nnode;
    v->nleaf_cap = t->nleaf_cap; v->nnode_cap = t->nnode_cap; v->first_leaf = t->first_leaf; v->first_node = t->first_node;
    v->leaf_npart = t->leaf_npart.data(); v->leaf_ipart = t->leaf_ipart.data();
    v->leaf_center = t->leaf_center.data(); v->leaf_width = t->leaf_width.data();
    v->node_npart = t->node_npart.data(); v->node_son = t->node_son.data(); v->node_split = t->node_split.data();
    v->node_center = t->node_center.data(); v->node_width = t->node_width.data();
    return 0;
}

int p2p_walk_task_p2p(const p2p_tree* t, double theta, double rcut, int nthreads, int** tt, int** ts, int64_t* ntask) {
    if (!t || !tt || !ts || !ntask) return -2;
    *tt = *ts = nullptr; *ntask = 0;
    if (t->nnode == 0) return 0;
    LocalWalk w{t, theta, rcut};
    return drive_walk<LocalWalk, 4>(w, Pair{t->first_node, t->first_node}, resolve_threads(nthreads), tt, ts, ntask);
}

}  // extern "C"

struct p2p_walk_plan {
    const p2p_tree* T;
    double theta, rcut;
    std::vector<int> row_begin;                     // [nchunk + 1] target-leaf boundaries
    std::vector<std::vector<Item>> items;           // per chunk, traversal order
};

extern "C" {

int p2p_walk_plan_create(const p2p_tree* t, double theta, double rcut, int nchunks_wanted, p2p_walk_plan** out) {
    if (!t || !out || nchunks_wanted < 1) return -2;
    p2p_walk_plan* P = new p2p_walk_plan();
    P->T = t; P->theta = theta; P->rcut = rcut;
    *out = P;
    if (t->nnode == 0) { P->row_begin = {0, t->nleaf}; P->items.resize(1); return 0; }
    // leaf range of every node (leaves are numbered left to right, nodes in pre-order)
    std::vector<int> lo((size_t)t->nnode, 0), hi((size_t)t->nnode, 0), depth((size_t)t->nnode, 0);
    for (int n = t->nnode - 1; n >= 0; n--) {       // children have larger pre-order ids than their parent
        int l = t->nleaf, h = 0;
        for (int k = 0; k < 2; k++) {
            const int son = t->node_son[2 * (size_t)n + k];
            if (son < 0) continue;
            if (son < t->first_node) { l = std::min(l, son - t->first_leaf); h = std::max(h, son - t->first_leaf + 1); }
            else { l = std::min(l, lo[(size_t)(son - t->first_node)]); h = std::max(h, hi[(size_t)(son - t->first_node)]); }
        }
        lo[(size_t)n] = l; hi[(size_t)n] = h;
    }
    for (int n = 0; n < t->nnode; n++)
        for (int k = 0; k < 2; k++) {
            const int son = t->node_son[2 * (size_t)n + k];
            if (son >= t->first_node) depth[(size_t)(son - t->first_node)] = depth[(size_t)n] + 1;
        }
    // chunk boundaries: subtrees at the depth that yields about nchunks_wanted pieces of similar leaf count
    int d = 0;
    while ((1 << d) < nchunks_wanted) d++;
    std::vector<int> bounds;
    bounds.push_back(0);
    std::vector<int> st;                            // global ids (leaf or node), visited left to right
    st.push_back(t->first_node);
    while (!st.empty()) {
        const int id = st.back();
        st.pop_back();
        if (id < t->first_node) { bounds.push_back(id - t->first_leaf + 1); continue; }   // a leaf above the cut depth
        const int n = id - t->first_node;
        if (depth[(size_t)n] >= d) { bounds.push_back(hi[(size_t)n]); continue; }
        for (int k = 1; k >= 0; k--) {
            const int son = t->node_son[2 * (size_t)n + k];
            if (son >= 0) st.push_back(son);
        }
    }
    std::sort(bounds.begin(), bounds.end());
    bounds.erase(std::unique(bounds.begin(), bounds.end()), bounds.end());
    if (bounds.back() != t->nleaf) bounds.push_back(t->nleaf);
    P->row_begin = bounds;
    const int nchunk = (int)bounds.size() - 1;
    P->items.resize((size_t)nchunk);
    auto chunk_of_leaf = [&](int leaf) { return (int)(std::upper_bound(bounds.begin(), bounds.end(), leaf) - bounds.begin()) - 1; };
    // expand from (root, root) until every item's target subtree lies inside one chunk
    LocalWalk w{t, theta, rcut};
    std::vector<Item> cur, nxt;
    cur.push_back(Item{Pair{t->first_node, t->first_node}, false});
    for (int round = 0; round < 256; round++) {
        bool again = false;
        nxt.clear();
        for (const Item& it : cur) {
            if (it.p.im == -1 || it.p.jm == -1) continue;
            int l, h;
            if (it.p.im < t->first_node) { l = it.p.im - t->first_leaf; h = l + 1; }
            else { l = lo[(size_t)(it.p.im - t->first_node)]; h = hi[(size_t)(it.p.im - t->first_node)]; }
            if (it.emit || chunk_of_leaf(l) == chunk_of_leaf(h - 1)) { nxt.push_back(it); continue; }
            bool emit;
            Pair kids[4];
            const int nk = w.step(it.p, &emit, kids);
            if (emit) nxt.push_back(Item{it.p, true});
            for (int k = 0; k < nk; k++) nxt.push_back(Item{kids[k], false});
            again = true;
        }
        cur.swap(nxt);
        if (!again) break;
    }
    for (const Item& it : cur) {
        const int l = it.p.im < t->first_node ? it.p.im - t->first_leaf : lo[(size_t)(it.p.im - t->first_node)];
        P->items[(size_t)chunk_of_leaf(l)].push_back(it);
    }
    return 0;
}

int p2p_walk_plan_nchunks(const p2p_walk_plan* P) { return P ? (int)P->items.size() : -2; }

int p2p_walk_plan_rows(const p2p_walk_plan* P, int c, int* b, int* e) {
    if (!P || c < 0 || c >= (int)P->items.size()) return -2;
    if (b) *b = P->row_begin[(size_t)c];
    if (e) *e = P->row_begin[(size_t)c + 1];
    return 0;
}

int p2p_walk_plan_run(const p2p_walk_plan* P, int c, int nthreads, int** tt, int** ts, int64_t* ntask) {
    if (!P || !tt || !ts || !ntask || c < 0 || c >= (int)P->items.size()) return -2;
    nthreads = resolve_threads(nthreads);
    LocalWalk w{P->T, P->theta, P->rcut};
    std::vector<Item> cur = P->items[(size_t)c];
    expand_items<LocalWalk, 4>(w, cur, (size_t)nthreads * 64);      // enough items for the threads
    return run_items(w, cur, nthreads, false, tt, ts, ntask);
}

int p2p_walk_plan_run_into(const p2p_walk_plan* P, int c, int nthreads, int* tt, int* ts, int64_t cap, int64_t* ntask) {
    if (!P || !tt || !ts || !ntask || c < 0 || c >= (int)P->items.size()) return -2;
    nthreads = resolve_threads(nthreads);
    LocalWalk w{P->T, P->theta, P->rcut};
    std::vector<Item> cur = P->items[(size_t)c];
    expand_items<LocalWalk, 4>(w, cur, (size_t)nthreads * 64);
    return run_items(w, cur, nthreads, false, nullptr, nullptr, ntask, tt, ts, cap);
}

void p2p_walk_plan_free(p2p_walk_plan* P) { delete P; }

int p2p_walk_task_p2p_ext(const p2p_tree* t, const p2p_image* img, double theta, double rcut, int nthreads, int** tt,
                          int** ts, int64_t* ntask) {
    if (!t || !img || !tt || !ts || !ntask) return -2;
    *tt = *ts = nullptr; *ntask = 0;
    if (t->nnode == 0 || img->nnode == 0) return 0;
    ExtWalk w{t, img, theta, rcut};
    return drive_walk<ExtWalk, 2>(w, Pair{t->first_node, 0}, resolve_threads(nthreads), tt, ts, ntask);
}

// prepare_sendtree2, 1_Indexing/src/remotes.c:337-446
int p2p_prepare_sendtree(const p2p_tree* t, const double* records, int64_t stride, const double tc[3], const double tw[3],
                         const double disp[3], double theta, double rcut, p2p_image* img) {
    if (!t || !img || stride < 3) return -2;
    memset(img, 0, sizeof *img);
    if (t->nnode == 0) return 0;
    std::vector<int> npart, son;
    std::vector<double> center, width, body;
    struct Job { int ilocal, parent, slot; };
    // The reference numbers image nodes in call (pre-)order; an explicit stack with child 1 pushed
    // first reproduces it.
    std::vector<Job> st;
    st.push_back(Job{t->first_node, -1, 0});
    while (!st.empty()) {
        Job j = st.back();
        st.pop_back();
        const int me = (int)npart.size();
        if (j.parent >= 0) son[2 * (size_t)j.parent + j.slot] = me;
        npart.push_back(0); son.push_back(0); son.push_back(0);
        center.resize(center.size() + 3); width.resize(width.size() + 3);
        if (j.ilocal < t->first_node) {
            const int li = j.ilocal - t->first_leaf;
            npart[(size_t)me] = t->leaf_npart[(size_t)li];
            for (int k = 0; k < 3; k++) {
                center[3 * (size_t)me + k] = t->leaf_center[3 * (size_t)li + k] + disp[k];
                width[3 * (size_t)me + k] = t->leaf_width[3 * (size_t)li + k];
            }
            son[2 * (size_t)me] = (int)(body.size() / 3);
            for (int p = t->leaf_ipart[(size_t)li]; p < t->leaf_ipart[(size_t)li] + t->leaf_npart[(size_t)li]; p++)
                for (int k = 0; k < 3; k++) body.push_back(records[(int64_t)p * stride + k] + disp[k]);
            son[2 * (size_t)me + 1] = (int)(body.size() / 3);
            continue;
        }
        const int ni = j.ilocal - t->first_node;
        const double* nc = &t->node_center[3 * (size_t)ni];
        const double* nw = &t->node_width[3 * (size_t)ni];
        double dr = 0.0;
        for (int k = 0; k < 3; k++) {
            double d = tc[k] - nc[k] - disp[k];
            if (d < 0.0) d = -d;
            d -= (tw[k] + nw[k]) * 0.5;
            if (d > 0.0) dr += d * d;
        }
        dr = sqrt(dr);
        npart[(size_t)me] = t->node_npart[(size_t)ni];
        for (int k = 0; k < 3; k++) { center[3 * (size_t)me + k] = nc[k] + disp[k]; width[3 * (size_t)me + k] = nw[k]; }
        double wmax = nw[0];
        if (wmax < nw[1]) wmax = nw[1];
        if (wmax < nw[2]) wmax = nw[2];
        if (dr >= rcut || wmax < 0.95 * theta * dr) { son[2 * (size_t)me] = son[2 * (size_t)me + 1] = -1; continue; }
        for (int n = 1; n >= 0; n--) {
            const int idx = t->node_son[2 * (size_t)ni + n];
            if (idx >= t->first_leaf) st.push_back(Job{idx, me, n});
        }
    }
    img->nnode = (int)npart.size();
    img->nbody = (int)(body.size() / 3);
    img->npart = (int*)malloc(sizeof(int) * npart.size());
    img->son = (int*)malloc(sizeof(int) * son.size());
    img->center = (double*)malloc(sizeof(double) * center.size());
    img->width = (double*)malloc(sizeof(double) * width.size());
    img->body = (double*)malloc(sizeof(double) * (body.size() ? body.size() : 1));
    if (!img->npart || !img->son || !img->center || !img->width || !img->body) { p2p_image_free(img); return -1; }
    memcpy(img->npart, npart.data(), sizeof(int) * npart.size());
    memcpy(img->son, son.data(), sizeof(int) * son.size());
    memcpy(img->center, center.data(), sizeof(double) * center.size());
    memcpy(img->width, width.data(), sizeof(double) * width.size());
    if (!body.empty()) memcpy(img->body, body.data(), sizeof(double) * body.size());
    return 0;
}

void p2p_image_free(p2p_image* img) {
    if (!img) return;
    free(img->npart); free(img->son); free(img->center); free(img->width); free(img->body);
    memset(img, 0, sizeof *img);
}

int p2p_domain_of_rank(int nproc, int rank) {  // 1_Indexing/src/initial.c:218-221
    int d = rank + mostleft_of(nproc);
    if (d > 2 * nproc - 2) d -= nproc;
    return d;
}

int p2p_domain_setup(int nproc, double box, double* split, double* center, double* width, int* direct_of_node) {
    if (nproc < 1 || !split || !center || !width || !direct_of_node) return -2;
    const int P = nproc, len = 2 * P - 1;
    // domain_initialize (1_Indexing/src/domains.c:433-469): every rank weighs 1, so a node's split
    // divides its interval in the ratio of the rank counts below it (heap order: sons 2n+1, 2n+2).
    std::vector<double> weight((size_t)len, 1.0);
    for (int n = P - 2; n >= 0; n--) weight[(size_t)n] = weight[2 * (size_t)n + 1] + weight[2 * (size_t)n + 2];
    struct Frame { int n, dim; double l[3], r[3]; };
    std::vector<Frame> st;
    st.push_back(Frame{0, 0, {0, 0, 0}, {box, box, box}});
    while (!st.empty()) {
        Frame f = st.back();
        st.pop_back();
        for (int k = 0; k < 3; k++) {  // center_toptree, 1_Indexing/src/toptree.c:153-159
            width[3 * f.n + k] = f.r[k] - f.l[k];
            center[3 * f.n + k] = 0.5 * (f.r[k] + f.l[k]);
        }
        direct_of_node[f.n] = f.dim;
        split[f.n] = 0.0;
        if (f.n >= P - 1) continue;
        const double tl = weight[2 * (size_t)f.n + 1], tr = weight[2 * (size_t)f.n + 2];
        const double frac = f.l[f.dim] + (f.r[f.dim] - f.l[f.dim]) * tl / (tl + tr);  // domains.c:416-417
        split[f.n] = frac;
        Frame a = f, b = f;
        a.n = 2 * f.n + 1; a.dim = (f.dim + 1) % 3; a.r[f.dim] = frac;
        b.n = 2 * f.n + 2; b.dim = (f.dim + 1) % 3; b.l[f.dim] = frac;
        st.push_back(b);
        st.push_back(a);
    }
    return 0;
}

// center_toptree (1_Indexing/src/toptree.c:150-182) for GIVEN splits (after p2p_domain_relax): boxes of all 2P-1 nodes
int p2p_domain_boxes(int nproc, double box, const double* split, double* center, double* width, int* direct_of_node) {
    if (nproc < 1 || !split || !center || !width) return -2;
    const int P = nproc;
    struct Frame { int n, dim; double l[3], r[3]; };
    std::vector<Frame> st;
    st.push_back(Frame{0, 0, {0, 0, 0}, {box, box, box}});
    while (!st.empty()) {
        Frame f = st.back();
        st.pop_back();
        for (int k = 0; k < 3; k++) {
            width[3 * f.n + k] = f.r[k] - f.l[k];
            center[3 * f.n + k] = 0.5 * (f.r[k] + f.l[k]);
        }
        if (direct_of_node) direct_of_node[f.n] = f.dim;
        if (f.n >= P - 1) continue;
        Frame a = f, b = f;
        a.n = 2 * f.n + 1; a.dim = (f.dim + 1) % 3; a.r[f.dim] = split[f.n];
        b.n = 2 * f.n + 2; b.dim = (f.dim + 1) % 3; b.l[f.dim] = split[f.n];
        st.push_back(b);
        st.push_back(a);
    }
    return 0;
}

// Work-weighted relaxation of the splits: measure_domain_runtime + determine_split_domtree
// (1_Indexing/src/domains.c:20-38,86-157).  work[r] is rank r's task count; the fractions are
// W_r P / (sum W + 1e-4) as 1_Indexing/src/photoNs.c:303 forms them.  Every split moves by
// 0.15 (t2 - t1) / (t1 nl / wl + t2 nr / wr), t = work per rank on either side, w = current widths, with the
// bounds of the recursion taken from the OLD splits.
int p2p_domain_relax(int nproc, double box, double* split, const double* work) {
    if (nproc < 1 || !split || !work) return -2;
    const int P = nproc, len = 2 * P - 1, ml = mostleft_of(P);
    double tot = 0.0;
    for (int r = 0; r < P; r++) tot += work[r];
    std::vector<double> tn((size_t)len, 0.0), tl((size_t)len, 1.0), tr((size_t)len, 1.0);
    for (int r = 0; r < P; r++) {
        int idom = r + ml;
        if (idom > 2 * P - 2) idom -= P;
        tn[(size_t)idom] = work[r] * P / (tot + 0.0001);
    }
    for (int n = P - 2; n >= 0; n--) {              // heap order: sons have larger indices
        tl[(size_t)n] = tn[2 * (size_t)n + 1];
        tr[(size_t)n] = tn[2 * (size_t)n + 2];
        tn[(size_t)n] = tl[(size_t)n] + tr[(size_t)n];
    }
    auto ranks_below = [](int size, int* l, int* r) {   // fraction(), 1_Indexing/src/domains.c:42-84
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
    };
    struct Job { int n, D, np; double l[3], r[3]; };
    std::vector<Job> st;
    std::vector<double> fresh(split, split + len);
    st.push_back(Job{0, 0, P, {0, 0, 0}, {box, box, box}});
    while (!st.empty()) {
        const Job j = st.back();
        st.pop_back();
        if (j.n >= P - 1) continue;
        int nl, nr;
        ranks_below(j.np, &nl, &nr);
        const double t1 = tl[(size_t)j.n] / nl, t2 = tr[(size_t)j.n] / nr;
        const double w0l = split[j.n] - j.l[j.D], w0r = j.r[j.D] - split[j.n];
        fresh[(size_t)j.n] = split[j.n] + 0.5 * 0.3 * (t2 - t1) / (t1 * nl / w0l + t2 * nr / w0r);
        Job a = j, b = j;
        a.n = 2 * j.n + 1; a.D = (j.D + 1) % 3; a.np = nl; a.r[j.D] = split[j.n];
        b.n = 2 * j.n + 2; b.D = (j.D + 1) % 3; b.np = nr; b.l[j.D] = split[j.n];
        st.push_back(b);
        st.push_back(a);
    }
    for (int n = 0; n < len; n++) split[n] = fresh[(size_t)n];
    return 0;
}

// bksort_body_inplace + prepare_body_inOrderOf_domain, 1_Indexing/src/domains.c:163-296
int p2p_domain_route(int nproc, const double* split, double* records, int64_t stride, int64_t* payload, int64_t npart,
                     int* sendcount) {
    if (nproc < 1 || !split || !sendcount || stride < 3 || (npart && !records)) return -2;
    const Records R{records, stride, payload};
    const int P = nproc, ml = mostleft_of(P);
    for (int r = 0; r < P; r++) sendcount[r] = 0;
    struct Job { int node, D; int64_t lo; int len; };
    std::vector<Job> st;
    st.push_back(Job{0, 0, 0, (int)npart});
    while (!st.empty()) {
        Job j = st.back();
        st.pop_back();
        if (j.node >= P - 1) { sendcount[(j.node - ml + P) % P] = j.len; continue; }
        const double s = split[j.node];
        const int D = j.D, len = j.len;
        const int64_t lo = j.lo;
        int left;
        if (len == 0) left = 0;
        else if (len == 1) left = R.key(lo, D) > s ? 0 : 1;
        else if (len == 2) {
            if (R.key(lo, D) > R.key(lo + 1, D)) R.swap(lo, lo + 1);
            if (R.key(lo, D) > s) left = 0;
            else if (R.key(lo + 1, D) <= s) left = 2;
            else left = 1;
        } else {
            int top = 0;
            while (top < len && R.key(lo + top, D) <= s) top++;
            int but = len - 1;
            while (but >= 0 && R.key(lo + but, D) > s) but--;
            if (top == len) left = len;
            else if (but == -1) left = 0;
            else {
                int n;
                for (n = top; n <= but; n++) {
                    if (R.key(lo + n, D) > s) {
                        R.swap(lo + n, lo + but);
                        while (R.key(lo + but, D) > s) but--;
                    }
                }
                left = (n == but) ? but + 1 : n;
            }
        }
        st.push_back(Job{2 * j.node + 2, (D + 1) % 3, lo + left, len - left});
        st.push_back(Job{2 * j.node + 1, (D + 1) % 3, lo, left});
    }
    return 0;
}

}  // extern "C"
