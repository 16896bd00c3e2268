// tree_walk.cpp — a small FDPS-like tree / interaction-list generator for SYNTHETIC inputs.
//
// FDPS (github.com/FDPS/FDPS, un-vendored by the reference, absent in this container) builds the
// tree, the local essential tree (LET) and the per-i-group interaction lists, and hands them to
// the dispatch functor.  That stays FDPS's job in a real PeTar.  This harness only exists to
// produce, for benchmarks and parity tests, inputs of exactly the shape FDPS emits at the call
// site reference src/petar.hpp:894-899 (tree type TreeForForceLong<...>::
// QuadrupoleWithSymmetrySearch, src/petar.hpp:619-627; theta = 0.3, n_leaf_limit = 20,
// n_group_limit = 512, src/petar.hpp:153-159):
//
//   * epj_sorted: all j-particles (local + LET) in Morton order;
//   * spj_sorted: one superparticle per tree cell — mass, centre of mass, RAW second moments
//     about the centre of mass (the kernels form the traceless part themselves,
//     reference src/soft_force.hpp:179-194) — followed by the LET superparticles received
//     from other domains;
//   * i-groups ("walks") = tree cells of the LOCAL tree with <= n_group_limit particles;
//   * per group: id_epj[] / id_spj[] from a walk of the global tree with the opening rule
//       open(cell)  <=>  dist^2(group box, cell c.m.) <= (cell length / theta)^2
//                        or the group's search box touches the cell's particle box
//                        or the cell's search box touches the group's particle box,
//     search boxes being grown by 0.99 * r_search (SAFTY_FACTOR_FOR_SEARCH, reference
//     src/ptcl.hpp:8, src/soft_ptcl.hpp:302-306, 349-353): every j that can be a neighbour of an i
//     of the group is in the EP list, never inside a superparticle.
//   * LET for a remote domain: the same rule with the remote domain's boxes as the "group".
//
// Also here: the Plummer-model recipe of reference src/particle_distribution_generator.hpp:173-250
// (std::mt19937 is the same MT19937 stream as FDPS's PS::MTTS).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>
#include <omp.h>

#include "petar_b200_types.h"

namespace {

constexpr double kSafetySearch = 0.99;   // reference src/ptcl.hpp:8

struct Box {
    double lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; k++) { lo[k] = 1e300; hi[k] = -1e300; } }
    void add(const double* p, double r = 0.0) {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k] - r); hi[k] = std::max(hi[k], p[k] + r); }
    }
    void merge(const Box& b) {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); }
    }
    bool valid() const { return lo[0] <= hi[0]; }
    bool overlaps(const Box& b) const {
        for (int k = 0; k < 3; k++) if (lo[k] > b.hi[k] || hi[k] < b.lo[k]) return false;
        return true;
    }
    double dist2(const double* p) const {
        double d2 = 0.0;
        for (int k = 0; k < 3; k++) {
            const double d = std::max(std::max(lo[k] - p[k], p[k] - hi[k]), 0.0);
            d2 += d * d;
        }
        return d2;
    }
};

// one tree element: a j-particle (EP) or a superparticle received through the LET (SP)
struct Elem {
    double pos[3];
    double mass;
    double rs;          // r_search (EP) or 0 (SP)
    double q[6];        // xx yy zz xy xz yz about pos (SP only; zero for EP)
    int    src;         // index in the caller's EP array (>=0) or ~index in the caller's SP array (<0)
    uint64_t key;
};

struct Node {
    int first, n;       // range in the sorted element array
    int child[8];
    int level;
    double cen[3], half;        // geometric cell
    double mass, cm[3], q[6];   // moments: raw second moments about cm
    Box inner, outer;           // particle box / search box of the elements inside
    bool leaf;
};

struct Tree {
    std::vector<Elem> el;       // Morton-sorted
    std::vector<Node> nodes;
    std::vector<int> ep_index;  // sorted position -> index in epj_sorted (or -1)
    std::vector<int> sp_index;  // sorted position -> index in the LET part of spj (or -1)
    int n_ep = 0, n_sp_let = 0;
    double root_cen[3], root_half;
    int n_leaf_limit = 20;
};

inline uint64_t spread21(uint64_t x) {
    x &= 0x1fffffULL;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}

uint64_t morton(const double* p, const double* cen, double half) {
    uint64_t k[3];
    for (int d = 0; d < 3; d++) {
        double u = (p[d] - (cen[d] - half)) / (2.0 * half);
        u = std::min(std::max(u, 0.0), 1.0 - 1e-12);
        k[d] = (uint64_t)(u * 2097152.0);
    }
    return spread21(k[0]) << 2 | spread21(k[1]) << 1 | spread21(k[2]);
}

void add_moment(Node& nd, double m, const double* p, const double* q) {
    // accumulate raw sums about the ORIGIN first; shifted to the c.m. in finish_moment
    nd.mass += m;
    for (int k = 0; k < 3; k++) nd.cm[k] += m * p[k];
    nd.q[0] += q[0] + m * p[0] * p[0]; nd.q[1] += q[1] + m * p[1] * p[1]; nd.q[2] += q[2] + m * p[2] * p[2];
    nd.q[3] += q[3] + m * p[0] * p[1]; nd.q[4] += q[4] + m * p[0] * p[2]; nd.q[5] += q[5] + m * p[1] * p[2];
}

int build_node(Tree& t, int first, int n, int level, const double* cen, double half) {
    const int id = (int)t.nodes.size();
    t.nodes.emplace_back();
    {
        Node& nd = t.nodes[id];
        nd.first = first; nd.n = n; nd.level = level; nd.half = half;
        for (int k = 0; k < 3; k++) nd.cen[k] = cen[k];
        for (int c = 0; c < 8; c++) nd.child[c] = -1;
        nd.mass = 0.0; nd.cm[0] = nd.cm[1] = nd.cm[2] = 0.0;
        for (int k = 0; k < 6; k++) nd.q[k] = 0.0;
        nd.inner.reset(); nd.outer.reset();
        nd.leaf = (n <= t.n_leaf_limit) || level >= 20;
    }
    if (t.nodes[id].leaf) {
        Node nd = t.nodes[id];
        // moments about a local reference point (the cell centre) to keep cancellation small
        for (int i = first; i < first + n; i++) {
            const Elem& e = t.el[i];
            double p[3] = {e.pos[0] - cen[0], e.pos[1] - cen[1], e.pos[2] - cen[2]};
            add_moment(nd, e.mass, p, e.q);
            nd.inner.add(e.pos);
            nd.outer.add(e.pos, kSafetySearch * e.rs);
        }
        t.nodes[id] = nd;
    } else {
        const int shift = 3 * (20 - level);
        int begin = first;
        for (int c = 0; c < 8; c++) {
            int end = begin;
            while (end < first + n && (int)((t.el[end].key >> shift) & 7) == c) end++;
            if (end > begin) {
                double cc[3] = {cen[0] + ((c & 4) ? 0.5 : -0.5) * half, cen[1] + ((c & 2) ? 0.5 : -0.5) * half,
                                cen[2] + ((c & 1) ? 0.5 : -0.5) * half};
                const int ch = build_node(t, begin, end - begin, level + 1, cc, 0.5 * half);
                t.nodes[id].child[c] = ch;
                const Node& cn = t.nodes[ch];
                Node& nd = t.nodes[id];
                // child moments are (mass, cm absolute, q about cm): re-express about this cell's centre
                double p[3] = {cn.cm[0] - cen[0], cn.cm[1] - cen[1], cn.cm[2] - cen[2]};
                add_moment(nd, cn.mass, p, cn.q);
                nd.inner.merge(cn.inner);
                nd.outer.merge(cn.outer);
            }
            begin = end;
        }
    }
    // finish: c.m. and second moments about the c.m.
    Node& nd = t.nodes[id];
    double c[3] = {0, 0, 0};
    if (nd.mass > 0.0) {
        for (int k = 0; k < 3; k++) c[k] = nd.cm[k] / nd.mass;
    } else if (nd.inner.valid()) {
        // massless cell (only zero-mass artificial particles): put the expansion centre inside
        for (int k = 0; k < 3; k++) c[k] = 0.5 * (nd.inner.lo[k] + nd.inner.hi[k]) - cen[k];
    }
    nd.q[0] -= nd.mass * c[0] * c[0]; nd.q[1] -= nd.mass * c[1] * c[1]; nd.q[2] -= nd.mass * c[2] * c[2];
    nd.q[3] -= nd.mass * c[0] * c[1]; nd.q[4] -= nd.mass * c[0] * c[2]; nd.q[5] -= nd.mass * c[1] * c[2];
    for (int k = 0; k < 3; k++) nd.cm[k] = c[k] + cen[k];
    return id;
}

void build_tree(Tree& t, int n_leaf_limit, const double* root_cen, double root_half) {
    t.n_leaf_limit = n_leaf_limit;
    for (int k = 0; k < 3; k++) t.root_cen[k] = root_cen[k];
    t.root_half = root_half;
    const long long n = (long long)t.el.size();
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) t.el[i].key = morton(t.el[i].pos, root_cen, root_half);
    std::stable_sort(t.el.begin(), t.el.end(), [](const Elem& a, const Elem& b) { return a.key < b.key; });
    t.ep_index.assign(n, -1); t.sp_index.assign(n, -1);
    t.n_ep = 0; t.n_sp_let = 0;
    for (long long i = 0; i < n; i++) {
        if (t.el[i].src >= 0) t.ep_index[i] = t.n_ep++;
        else                  t.sp_index[i] = t.n_sp_let++;
    }
    t.nodes.clear();
    t.nodes.reserve((size_t)(n / std::max(1, n_leaf_limit / 4)) + 64);
    if (n > 0) build_node(t, 0, (int)n, 0, root_cen, root_half);
}

void bounding_cube(const std::vector<Elem>& el, double* cen, double& half) {
    Box b; b.reset();
    for (const Elem& e : el) b.add(e.pos);
    half = 0.0;
    for (int k = 0; k < 3; k++) { cen[k] = 0.5 * (b.lo[k] + b.hi[k]); half = std::max(half, 0.5 * (b.hi[k] - b.lo[k])); }
    half = half * 1.001 + 1e-300;
}

// ---- i-groups: cells of the local tree with <= n_group_limit elements -----------------------
struct Group { int first, n; Box inner, outer; };

void make_groups(const Tree& t, int node, int n_group_limit, std::vector<Group>& out) {
    const Node& nd = t.nodes[node];
    if (nd.n <= n_group_limit || nd.leaf) {
        out.push_back({nd.first, nd.n, nd.inner, nd.outer});
        return;
    }
    for (int c = 0; c < 8; c++) if (nd.child[c] >= 0) make_groups(t, nd.child[c], n_group_limit, out);
}

// ---- the walk -----------------------------------------------------------------------------
// target = (inner box, outer/search box); lists are appended in tree (Morton) order
void walk(const Tree& t, int node, const Box& tin, const Box& tout, double theta_inv2,
          std::vector<int>& id_ep, std::vector<int>& id_sp) {
    const Node& nd = t.nodes[node];
    if (nd.n == 0) return;
    const double len = 2.0 * nd.half;
    const bool far_enough = tin.dist2(nd.cm) > len * len * theta_inv2;
    const bool search_touch = tout.overlaps(nd.inner) || nd.outer.overlaps(tin);
    if (far_enough && !search_touch) {
        id_sp.push_back(node);                    // superparticle of this cell: spj index = node id
        return;
    }
    if (nd.leaf) {
        for (int i = nd.first; i < nd.first + nd.n; i++) {
            if (t.ep_index[i] >= 0) id_ep.push_back(t.ep_index[i]);
            else                    id_sp.push_back((int)t.nodes.size() + t.sp_index[i]);   // LET SPs follow the cells
        }
        return;
    }
    for (int c = 0; c < 8; c++) if (nd.child[c] >= 0) walk(t, nd.child[c], tin, tout, theta_inv2, id_ep, id_sp);
}

// ---- everything one call produces ---------------------------------------------------------------
struct Result {
    Tree local, global;
    bool has_let = false;
    double t_tree = 0.0, t_walk = 0.0;          // wall-clock seconds of the tree build(s) and of all group walks
    std::vector<Group> groups;
    std::vector<long long> i_off, ej_off, sj_off;
    std::vector<int> id_epj, id_spj;
    const Tree& gt() const { return has_let ? global : local; }
};

void fill_elems(std::vector<Elem>& el, int n_ep, const double* pos, const double* mass, const double* rs,
                int src0, int n_sp, const pb_SPJQuad* sp) {
    const size_t base = el.size();
    el.resize(base + (size_t)n_ep + (size_t)n_sp);
    for (int i = 0; i < n_ep; i++) {
        Elem& e = el[base + i];
        e.pos[0] = pos[3 * i]; e.pos[1] = pos[3 * i + 1]; e.pos[2] = pos[3 * i + 2];
        e.mass = mass[i]; e.rs = rs[i];
        for (int k = 0; k < 6; k++) e.q[k] = 0.0;
        e.src = src0 + i; e.key = 0;
    }
    for (int i = 0; i < n_sp; i++) {
        Elem& e = el[base + n_ep + i];
        e.pos[0] = sp[i].pos.x; e.pos[1] = sp[i].pos.y; e.pos[2] = sp[i].pos.z;
        e.mass = sp[i].mass; e.rs = 0.0;
        e.q[0] = sp[i].qxx; e.q[1] = sp[i].qyy; e.q[2] = sp[i].qzz; e.q[3] = sp[i].qxy; e.q[4] = sp[i].qxz; e.q[5] = sp[i].qyz;
        e.src = ~i; e.key = 0;
    }
}

void node_to_spj(const Node& nd, pb_SPJQuad& s) {
    s.mass = nd.mass;
    s.pos.x = nd.cm[0]; s.pos.y = nd.cm[1]; s.pos.z = nd.cm[2];
    s.qxx = nd.q[0]; s.qyy = nd.q[1]; s.qzz = nd.q[2]; s.qxy = nd.q[3]; s.qxz = nd.q[4]; s.qyz = nd.q[5];
}

} // namespace

extern "C" {

// hz_set_skip_walk(1): hz_build makes the trees and the i-groups but no interaction lists (for inputs of the
// device-side walk at sizes where the host lists would not fit: BASELINE config 5); every group's lists are empty.
static int g_skip_walk = 0;
void hz_set_skip_walk(int on) { g_skip_walk = on; }

// The lists of ONE group of a tree built by hz_build (for spot checks when the walks were skipped): call with
// id_ep = id_sp = NULL to learn the lengths first.
void hz_walk_group(void* h, int g, double theta, long long* n_ep, long long* n_sp, int* id_ep, int* id_sp);

// Build local tree over n_loc local particles, optionally a global tree over local + LET
// (n_let_ep remote particles, n_let_sp remote superparticles), make i-groups from the local tree
// and walk the global tree for each.  Returns an opaque handle.
void* hz_build(int n_loc, const double* pos, const double* mass, const double* rsearch,
               int n_let_ep, const double* let_pos, const double* let_mass, const double* let_rsearch,
               int n_let_sp, const pb_SPJQuad* let_sp,
               double theta, int n_leaf_limit, int n_group_limit)
{
    Result* R = new Result();
    const double t_begin = omp_get_wtime();
    fill_elems(R->local.el, n_loc, pos, mass, rsearch, 0, 0, nullptr);
    double cen[3], half;
    bounding_cube(R->local.el, cen, half);
    build_tree(R->local, n_leaf_limit, cen, half);
    if (n_loc > 0) make_groups(R->local, 0, n_group_limit, R->groups);

    R->has_let = (n_let_ep > 0 || n_let_sp > 0);
    if (R->has_let) {
        fill_elems(R->global.el, n_loc, pos, mass, rsearch, 0, 0, nullptr);
        fill_elems(R->global.el, n_let_ep, let_pos, let_mass, let_rsearch, n_loc, n_let_sp, let_sp);
        bounding_cube(R->global.el, cen, half);
        build_tree(R->global, n_leaf_limit, cen, half);
    }
    const Tree& G = R->gt();
    const int ng = (int)R->groups.size();
    const double theta_inv2 = theta > 0.0 ? 1.0 / (theta * theta) : 1e300;
    std::vector<std::vector<int>> le(ng), ls(ng);
    R->t_tree = omp_get_wtime() - t_begin;
    const double t_w0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 4)
    for (int g = 0; g < ng; g++) {
        if (g_skip_walk) continue;
        le[g].reserve(4096); ls[g].reserve(2048);
        if (!G.nodes.empty()) walk(G, 0, R->groups[g].inner, R->groups[g].outer, theta_inv2, le[g], ls[g]);
    }
    R->t_walk = omp_get_wtime() - t_w0;
    R->i_off.assign(ng + 1, 0); R->ej_off.assign(ng + 1, 0); R->sj_off.assign(ng + 1, 0);
    for (int g = 0; g < ng; g++) {
        R->i_off[g + 1] = R->i_off[g] + R->groups[g].n;
        R->ej_off[g + 1] = R->ej_off[g] + (long long)le[g].size();
        R->sj_off[g + 1] = R->sj_off[g] + (long long)ls[g].size();
    }
    R->id_epj.resize((size_t)R->ej_off[ng]); R->id_spj.resize((size_t)R->sj_off[ng]);
#pragma omp parallel for schedule(static)
    for (int g = 0; g < ng; g++) {
        if (!le[g].empty()) memcpy(&R->id_epj[(size_t)R->ej_off[g]], le[g].data(), sizeof(int) * le[g].size());
        if (!ls[g].empty()) memcpy(&R->id_spj[(size_t)R->sj_off[g]], ls[g].data(), sizeof(int) * ls[g].size());
    }
    return R;
}

void hz_walk_group(void* h, int g, double theta, long long* n_ep, long long* n_sp, int* id_ep, int* id_sp) {
    Result* R = (Result*)h;
    const Tree& G = R->gt();
    std::vector<int> le, ls;
    const double theta_inv2 = theta > 0.0 ? 1.0 / (theta * theta) : 1e300;
    if (!G.nodes.empty()) walk(G, 0, R->groups[g].inner, R->groups[g].outer, theta_inv2, le, ls);
    *n_ep = (long long)le.size(); *n_sp = (long long)ls.size();
    if (id_ep && !le.empty()) memcpy(id_ep, le.data(), sizeof(int) * le.size());
    if (id_sp && !ls.empty()) memcpy(id_sp, ls.data(), sizeof(int) * ls.size());
}

// out: [n_epj, n_spj, n_walk, n_epi_total, n_id_epj, n_id_spj, n_nodes, n_let_sp]
void hz_sizes(void* h, long long* out) {
    Result* R = (Result*)h;
    const Tree& G = R->gt();
    const int ng = (int)R->groups.size();
    out[0] = G.n_ep; out[1] = (long long)G.nodes.size() + G.n_sp_let; out[2] = ng;
    out[3] = R->i_off[ng]; out[4] = R->ej_off[ng]; out[5] = R->sj_off[ng];
    out[6] = (long long)G.nodes.size(); out[7] = G.n_sp_let;
}

// epj_src[n_epj]: for each entry of epj_sorted the caller's source index (0..n_loc-1 local,
//   n_loc.. = LET EP);  epi_src[n_epi_total]: local particle index of each i (walk-major);
//   spj[n_spj]; offset tables (n_walk+1); index lists.
void hz_export(void* h, int* epj_src, int* epi_src, pb_SPJQuad* spj,
               long long* i_off, long long* ej_off, long long* sj_off, int* id_epj, int* id_spj) {
    Result* R = (Result*)h;
    const Tree& G = R->gt();
    for (size_t i = 0; i < G.el.size(); i++) {
        if (G.ep_index[i] >= 0) epj_src[G.ep_index[i]] = G.el[i].src;
    }
    const int ng = (int)R->groups.size();
    long long k = 0;
    for (int g = 0; g < ng; g++)
        for (int i = R->groups[g].first; i < R->groups[g].first + R->groups[g].n; i++) epi_src[k++] = R->local.el[i].src;
    const size_t nn = G.nodes.size();
    for (size_t i = 0; i < nn; i++) node_to_spj(G.nodes[i], spj[i]);
    for (size_t i = 0; i < G.el.size(); i++) {
        if (G.sp_index[i] >= 0) {
            pb_SPJQuad& s = spj[nn + G.sp_index[i]];
            const Elem& e = G.el[i];
            s.mass = e.mass; s.pos.x = e.pos[0]; s.pos.y = e.pos[1]; s.pos.z = e.pos[2];
            s.qxx = e.q[0]; s.qyy = e.q[1]; s.qzz = e.q[2]; s.qxy = e.q[3]; s.qxz = e.q[4]; s.qyz = e.q[5];
        }
    }
    memcpy(i_off, R->i_off.data(), sizeof(long long) * (ng + 1));
    memcpy(ej_off, R->ej_off.data(), sizeof(long long) * (ng + 1));
    memcpy(sj_off, R->sj_off.data(), sizeof(long long) * (ng + 1));
    if (!R->id_epj.empty()) memcpy(id_epj, R->id_epj.data(), sizeof(int) * R->id_epj.size());
    if (!R->id_spj.empty()) memcpy(id_spj, R->id_spj.data(), sizeof(int) * R->id_spj.size());
}

// out[0] = seconds spent building the tree(s), out[1] = seconds spent walking it for all groups (OpenMP)
void hz_timing(void* h, double* out) { Result* R = (Result*)h; out[0] = R->t_tree; out[1] = R->t_walk; }

// For the LET part of spj (entries n_nodes .. n_nodes+n_let_sp-1, Morton order): the index each
// entry had in the caller's let_sp array.
void hz_export_let_sp_src(void* h, int* out) {
    Result* R = (Result*)h;
    const Tree& G = R->gt();
    for (size_t i = 0; i < G.el.size(); i++)
        if (G.sp_index[i] >= 0) out[G.sp_index[i]] = ~G.el[i].src;
}

// The tree itself, in the layout of pb_tree_cell / pb_tree_group (include/petar_b200.h), for the
// device-side list builder.  Cell c's superparticle is spj[c]; its element range [first, first+n) indexes the
// Morton-sorted elements (= epj_sorted for a single-domain build; with LET elements see hz_export_elem_map).
struct ExportCell  { double cm[3], len, in_lo[3], in_hi[3], out_lo[3], out_hi[3]; int child[8]; int first, n, leaf, pad; };
struct ExportGroup { int first, n; double in_lo[3], in_hi[3], out_lo[3], out_hi[3]; };

void hz_export_tree(void* h, void* cells_out, void* groups_out) {
    Result* R = (Result*)h;
    const Tree& G = R->gt();
    ExportCell* c = (ExportCell*)cells_out;
    for (size_t i = 0; i < G.nodes.size(); i++) {
        const Node& nd = G.nodes[i];
        for (int k = 0; k < 3; k++) {
            c[i].cm[k] = nd.cm[k];
            c[i].in_lo[k] = nd.inner.lo[k]; c[i].in_hi[k] = nd.inner.hi[k];
            c[i].out_lo[k] = nd.outer.lo[k]; c[i].out_hi[k] = nd.outer.hi[k];
        }
        c[i].len = 2.0 * nd.half;
        for (int k = 0; k < 8; k++) c[i].child[k] = nd.child[k];
        c[i].first = nd.first; c[i].n = nd.n; c[i].leaf = nd.leaf ? 1 : 0;
        int nls = 0;                                   // LET superparticles among a leaf's elements
        if (nd.leaf) for (int e = nd.first; e < nd.first + nd.n; e++) nls += (G.sp_index[e] >= 0);
        c[i].pad = nls;
    }
    ExportGroup* g = (ExportGroup*)groups_out;
    for (size_t i = 0; i < R->groups.size(); i++) {
        const Group& gr = R->groups[i];
        g[i].first = gr.first; g[i].n = gr.n;
        for (int k = 0; k < 3; k++) {
            g[i].in_lo[k] = gr.inner.lo[k]; g[i].in_hi[k] = gr.inner.hi[k];
            g[i].out_lo[k] = gr.outer.lo[k]; g[i].out_hi[k] = gr.outer.hi[k];
        }
    }
}

// Element map of the (global) tree for the device walk: out[k] for the k-th Morton-sorted element is its
// index in epj_sorted (>= 0) or ~(index in the LET part of spj) (< 0).  Single-domain builds: the identity.
void hz_export_elem_map(void* h, int* out) {
    Result* R = (Result*)h;
    const Tree& G = R->gt();
    for (size_t i = 0; i < G.el.size(); i++) out[i] = G.ep_index[i] >= 0 ? G.ep_index[i] : ~G.sp_index[i];
}

// Local boxes of this domain: out[0..5] = particle box lo/hi, out[6..11] = search box lo/hi
void hz_local_boxes(void* h, double* out) {
    Result* R = (Result*)h;
    if (R->local.nodes.empty()) { for (int k = 0; k < 12; k++) out[k] = (k % 6) < 3 ? 1e300 : -1e300; return; }
    const Node& r = R->local.nodes[0];
    for (int k = 0; k < 3; k++) { out[k] = r.inner.lo[k]; out[3 + k] = r.inner.hi[k]; out[6 + k] = r.outer.lo[k]; out[9 + k] = r.outer.hi[k]; }
}

// LET for one remote domain (its particle box and search box): which local particles go as EP,
// which local cells go as SP.  Call with null outputs to get the counts first.
void hz_make_let(void* h, const double* remote_boxes, double theta, long long* n_ep, long long* n_sp,
                 int* ep_src, pb_SPJQuad* sp) {
    Result* R = (Result*)h;
    Box tin, tout;
    for (int k = 0; k < 3; k++) { tin.lo[k] = remote_boxes[k]; tin.hi[k] = remote_boxes[3 + k]; tout.lo[k] = remote_boxes[6 + k]; tout.hi[k] = remote_boxes[9 + k]; }
    std::vector<int> le, ls;
    const double theta_inv2 = theta > 0.0 ? 1.0 / (theta * theta) : 1e300;
    if (!R->local.nodes.empty() && tin.valid()) walk(R->local, 0, tin, tout, theta_inv2, le, ls);
    *n_ep = (long long)le.size(); *n_sp = (long long)ls.size();
    if (ep_src) {
        // ep_index -> sorted position is the identity for the local tree (no LET elements in it)
        for (size_t i = 0; i < le.size(); i++) ep_src[i] = R->local.el[le[i]].src;
    }
    if (sp) for (size_t i = 0; i < ls.size(); i++) node_to_spj(R->local.nodes[ls[i]], sp[i]);
}

void hz_free(void* h) { delete (Result*)h; }

// ---- Plummer model, reference src/particle_distribution_generator.hpp:173-250 -----------------
// (rank_seed = PS::Comm::getRank(); the reference ignores its own `seed` argument)
void hz_make_plummer(double mass_glb, long long n_glb, long long n_loc, double* mass, double* pos, double* vel,
                     double eng, unsigned rank_seed) {
    const double PI = std::atan(1.0) * 4.0;
    const double r_cutoff = 22.8 / (-3.0 * PI * mass_glb * mass_glb / (64.0 * -0.25));
    std::mt19937 gen(rank_seed);
    auto res53 = [&]() { const uint32_t a = (uint32_t)gen() >> 5, b = (uint32_t)gen() >> 6; return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0); };
    auto real2 = [&]() { return (uint32_t)gen() * (1.0 / 4294967296.0); };
    for (long long i = 0; i < n_loc; i++) {
        mass[i] = mass_glb / n_glb;
        double r = 9999.9;
        while (r > r_cutoff) { const double m = res53(); r = 1.0 / std::sqrt(std::pow(m, (-2.0 / 3.0)) - 1.0); }
        double phi = 2.0 * PI * res53();
        double cth = 2.0 * (real2() - 0.5);
        double sth = std::sqrt(1.0 - cth * cth);
        pos[3 * i] = r * sth * std::cos(phi); pos[3 * i + 1] = r * sth * std::sin(phi); pos[3 * i + 2] = r * cth;
        for (;;) {
            const double v_max = 0.1;
            const double v_try = res53();
            const double v_crit = v_max * res53();
            if (v_crit < v_try * v_try * std::pow((1.0 - v_try * v_try), 3.5)) {
                const double ve = std::sqrt(2.0) * std::pow((r * r + 1.0), -0.25);
                phi = 2.0 * PI * res53();
                cth = 2.0 * (res53() - 0.5);
                sth = std::sqrt(1.0 - cth * cth);
                vel[3 * i] = ve * v_try * sth * std::cos(phi); vel[3 * i + 1] = ve * v_try * sth * std::sin(phi); vel[3 * i + 2] = ve * v_try * cth;
                break;
            }
        }
    }
    double cp[3] = {0, 0, 0}, cv[3] = {0, 0, 0}, cm = 0.0;
    for (long long i = 0; i < n_loc; i++) {
        for (int k = 0; k < 3; k++) { cp[k] += mass[i] * pos[3 * i + k]; cv[k] += mass[i] * vel[3 * i + k]; }
        cm += mass[i];
    }
    for (int k = 0; k < 3; k++) { cp[k] /= cm; cv[k] /= cm; }
    const double r_scale = -3.0 * PI * mass_glb * mass_glb / (64.0 * eng);
    const double coef = 1.0 / std::sqrt(r_scale);
    for (long long i = 0; i < n_loc; i++)
        for (int k = 0; k < 3; k++) {
            pos[3 * i + k] = (pos[3 * i + k] - cp[k]) * r_scale;
            vel[3 * i + k] = (vel[3 * i + k] - cv[k]) * coef;
        }
}

} // extern "C"
