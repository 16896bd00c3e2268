// pb_walk.cu — device-side interaction-list construction.
//
// One warp per i-group walks the tree breadth-first with the opening rule of FDPS's
// QuadrupoleWithSymmetrySearch walk (see include/petar_b200.h, "device-side interaction lists"):
// each lane classifies one cell of the current frontier, warp ballots / scans compact the three
// outcomes — superparticle accepted, leaf opened (its particle range goes to the EP list), cell
// opened (its children go to the next frontier).  Two passes over the same traversal: COUNT
// (list lengths, needed to plan the force tasks) and FILL (writes the indices straight into the
// dispatch arena the force kernel reads — the lists never exist on the host).
//
// All geometry is fp64 and this file is compiled with -fmad=false, so every opening decision is
// bit-identical to the host walk of petar_b200/harness/tree_walk.cpp; list CONTENTS are therefore
// identical sets (the order differs: level order here, depth-first there).
#include "pb_device.h"
#include "petar_b200.h"

namespace pb {

namespace {

__device__ __forceinline__ double box_dist2(const double* lo, const double* hi, const double* p) {
    double d2 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double d = fmax(fmax(lo[k] - p[k], p[k] - hi[k]), 0.0);
        d2 += d * d;
    }
    return d2;
}

__device__ __forceinline__ bool box_overlap(const double* alo, const double* ahi, const double* blo, const double* bhi) {
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (alo[k] > bhi[k] || ahi[k] < blo[k]) return false;
    return true;
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
    int s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += t;
    }
    total = __shfl_sync(0xffffffffu, s, 31);
    return s - v;
}

} // namespace

// FILL = false: counts[g] = {n_ep, n_sp}.  FILL = true: ids written at id_e + offs[g].x / id_s + offs[g].y; with `caps`
// (speculative fill: offsets reserved from the previous step's list lengths) nothing is written past a group's
// reservation, the true lengths still go to counts[g], and overflow[1] is raised if any list did not fit.
template <bool FILL>
__global__ void __launch_bounds__(128)
walk_kernel(const pb_tree_cell* __restrict__ cells, const pb_tree_group* __restrict__ groups,
            int g0, int n_groups, double theta_inv2,
            int2* __restrict__ counts, const int2* __restrict__ offs, int* __restrict__ id_e, int* __restrict__ id_s,
            int* __restrict__ scratch, int cap, int* __restrict__ overflow,
            const int* __restrict__ elem_map, int n_cells, const int2* __restrict__ caps)
{
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    int* qa = scratch + (size_t)warp_global * 2 * cap;
    int* qb = qa + cap;

    for (int g = warp_global; g < n_groups; g += n_warps) {
        const pb_tree_group grp = groups[g0 + g];
        int* cur = qa; int* nxt = qb;
        if (lane == 0) cur[0] = 0;                 // root
        __syncwarp();
        int ncur = 1, nep = 0, nsp = 0;
        int* oe = nullptr; int* os = nullptr;
        int cap_e = 0x7fffffff, cap_s = 0x7fffffff;   // speculative fill: room reserved from the previous step's lengths
        if (FILL) { const int2 o = offs[g0 + g]; oe = id_e + o.x; os = id_s + o.y; }
        if (FILL && caps) { const int2 c = caps[g0 + g]; cap_e = c.x; cap_s = c.y; }
        while (ncur > 0) {
            int nnext = 0;
            for (int base = 0; base < ncur; base += 32) {
                const int idx = base + lane;
                int cls = 0, first = 0, n = 0, nls = 0, cell = -1;
                int child[8];
                if (idx < ncur) {
                    cell = cur[idx];
                    const pb_tree_cell& c = cells[cell];
                    n = c.n;
                    if (n > 0) {
                        const double len = c.len;
                        const bool far_enough = box_dist2(grp.in_lo, grp.in_hi, c.cm) > len * len * theta_inv2;
                        const bool touch = box_overlap(grp.out_lo, grp.out_hi, c.in_lo, c.in_hi) ||
                                           box_overlap(c.out_lo, c.out_hi, grp.in_lo, grp.in_hi);
                        if (far_enough && !touch) cls = 1;
                        else if (c.leaf) { cls = 2; first = c.first; nls = elem_map ? c.n_let_sp : 0; }
                        else {
                            cls = 3;
#pragma unroll
                            for (int k = 0; k < 8; k++) child[k] = c.child[k];
                        }
                    }
                }
                // superparticles
                const unsigned m1 = __ballot_sync(0xffffffffu, cls == 1);
                if (FILL && cls == 1) {
                    const int at = nsp + __popc(m1 & ((1u << lane) - 1u));
                    if (at < cap_s) os[at] = cell;
                }
                nsp += __popc(m1);
                // opened leaves: element ranges -> EP list (and, with a local essential tree, the superparticles
                // received from other domains -> SP list; elem_map says where each sorted element is stored)
                int tot;
                const int off_e = warp_excl_scan(cls == 2 ? n - nls : 0, lane, tot);
                if (elem_map == nullptr) {
                    if (FILL && cls == 2)
                        for (int k = 0; k < n; k++) if (nep + off_e + k < cap_e) oe[nep + off_e + k] = first + k;
                } else {
                    int tot_s;
                    const int off_s = warp_excl_scan(cls == 2 ? nls : 0, lane, tot_s);
                    if (FILL && cls == 2) {
                        int ke = nep + off_e, ks = nsp + off_s;
                        for (int k = 0; k < n; k++) {
                            const int m = elem_map[first + k];
                            if (m >= 0) { if (ke < cap_e) oe[ke] = m; ke++; }
                            else        { if (ks < cap_s) os[ks] = n_cells + ~m; ks++; }
                        }
                    }
                    nsp += tot_s;
                }
                nep += tot;
                // opened cells: children -> next frontier
                int nch = 0;
                if (cls == 3) {
#pragma unroll
                    for (int k = 0; k < 8; k++) nch += (child[k] >= 0);
                }
                const int off_c = warp_excl_scan(nch, lane, tot);
                if (nnext + tot > cap) { if (lane == 0) atomicExch(overflow, 1); tot = 0; nch = 0; }
                if (cls == 3 && nch > 0) {
                    int w = nnext + off_c;
#pragma unroll
                    for (int k = 0; k < 8; k++) if (child[k] >= 0) nxt[w++] = child[k];
                }
                nnext += tot;
            }
            __syncwarp();
            int* t = cur; cur = nxt; nxt = t;
            ncur = nnext;
        }
        if (lane == 0) {
            if (counts) counts[g0 + g] = make_int2(nep, nsp);
            if (FILL && caps && (nep > cap_e || nsp > cap_s)) atomicExch(overflow + 1, 1);   // a list outgrew its reservation
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Compact walk records.  The walk is latency-bound: per visited cell the kernel above pulls a 176-byte fp64 record
// (eleven loads over two or three cache lines) of which an accepted or opened cell needs a fraction.  Here every cell
// also gets
//   A (64 B, fp32): c.m., length (< 0: empty cell), particle box and search box rounded OUTWARD,
//   B (48 B, int):  children, element range, leaf flag, LET superparticle count,
// and the opening rule is first evaluated conservatively in fp32 on A: with S the largest |coordinate| of the tree,
// every fp32 distance component is within 2^-22 S of the fp64 one, so a cell is opened (or accepted) on the fp32
// result only when the decision holds for every value within those bounds; outward-rounded boxes that do not overlap
// prove that the exact boxes do not.  Whatever stays undecided — cells within a few ulps of the opening radius, and
// far cells whose rounded boxes touch the group's — is decided on the fp64 record exactly as before, so the lists are
// the same sets in the same order.  Two frontier chunks are classified per iteration (their loads overlap), then
// compacted one after the other.
// ------------------------------------------------------------------------------------------------------------------
struct CellA { float cx, cy, cz, len; float ilo[3], ihi[3], olo[3], ohi[3]; };       // 64 B
struct CellB { int child[8]; int first, n, leaf, n_let_sp; };                          // 48 B

__global__ void __launch_bounds__(256)
compact_cells_kernel(const pb_tree_cell* __restrict__ cells, int n_cells, CellA* __restrict__ A, CellB* __restrict__ B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const pb_tree_cell& c = cells[i];
    CellA a; CellB b;
    a.cx = (float)c.cm[0]; a.cy = (float)c.cm[1]; a.cz = (float)c.cm[2];
    a.len = c.n > 0 ? (float)c.len : -1.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        a.ilo[k] = __double2float_rd(c.in_lo[k]);  a.ihi[k] = __double2float_ru(c.in_hi[k]);
        a.olo[k] = __double2float_rd(c.out_lo[k]); a.ohi[k] = __double2float_ru(c.out_hi[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) b.child[k] = c.child[k];
    b.first = c.first; b.n = c.n; b.leaf = c.leaf; b.n_let_sp = c.n_let_sp;
    A[i] = a; B[i] = b;
}

namespace {
struct GroupF {                       // a group's boxes rounded outward to fp32
    float ilo[3], ihi[3], olo[3], ohi[3];
};
// 0: undecided, 1: accept as superparticle, 2: open
__device__ __forceinline__ int classify_f32(const CellA& a, const GroupF& g, float theta_inv, float margin) {
    float d2 = 0.f;
    const float p[3] = {a.cx, a.cy, a.cz};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float d = fmaxf(fmaxf(g.ilo[k] - p[k], p[k] - g.ihi[k]), 0.f);
        d2 += d * d;
    }
    const float d = sqrtf(d2);                       // (IEEE sqrt; -fmad=false: no contraction above)
    const float L = a.len * theta_inv;
    const float rel = 1.0e-6f;                        // >> the few 2^-24 of the fp32 operations themselves
    if (d * (1.f + rel) + margin < L * (1.f - rel)) return 2;                       // certainly inside the opening radius
    if (d * (1.f - rel) - margin > L * (1.f + rel)) {                               // certainly far enough ...
        bool touch = true;
#pragma unroll
        for (int k = 0; k < 3; k++) if (g.olo[k] > a.ihi[k] || g.ohi[k] < a.ilo[k]) touch = false;
        if (!touch) {
            touch = true;
#pragma unroll
            for (int k = 0; k < 3; k++) if (a.olo[k] > g.ihi[k] || a.ohi[k] < g.ilo[k]) touch = false;
        }
        if (!touch) return 1;                                                       // ... and certainly no search-box contact
    }
    return 0;
}
} // namespace

template <bool FILL>
__global__ void __launch_bounds__(128, 4)
walk_kernel_c(const pb_tree_cell* __restrict__ cells, const CellA* __restrict__ cA, const CellB* __restrict__ cB,
              const pb_tree_group* __restrict__ groups, int g0, int n_groups, double theta_inv2, float theta_inv, float margin,
              int2* __restrict__ counts, const int2* __restrict__ offs, int* __restrict__ id_e, int* __restrict__ id_s,
              int* __restrict__ scratch, int cap, int* __restrict__ overflow,
              const int* __restrict__ elem_map, int n_cells, const int2* __restrict__ caps)
{
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    int* qa = scratch + (size_t)warp_global * 2 * cap;
    int* qb = qa + cap;

    for (int g = warp_global; g < n_groups; g += n_warps) {
        const pb_tree_group* gp = groups + (g0 + g);     // fp64 boxes stay in memory: only the rare exact re-check reads them
        GroupF gf;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            gf.ilo[k] = __double2float_rd(gp->in_lo[k]);  gf.ihi[k] = __double2float_ru(gp->in_hi[k]);
            gf.olo[k] = __double2float_rd(gp->out_lo[k]); gf.ohi[k] = __double2float_ru(gp->out_hi[k]);
        }
        int* cur = qa; int* nxt = qb;
        if (lane == 0) cur[0] = 0;                 // root
        __syncwarp();
        int ncur = 1, nep = 0, nsp = 0;
        int* oe = nullptr; int* os = nullptr;
        int cap_e = 0x7fffffff, cap_s = 0x7fffffff;
        if (FILL) { const int2 o = offs[g0 + g]; oe = id_e + o.x; os = id_s + o.y; }
        if (FILL && caps) { const int2 c = caps[g0 + g]; cap_e = c.x; cap_s = c.y; }
        while (ncur > 0) {
            int nnext = 0;
            for (int base = 0; base < ncur; base += 64) {
                // ---- classify two chunks of 32 frontier cells: all loads first ----
                int cell2[2], cls2[2];
                CellA a2[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int idx = base + 32 * h + lane;
                    cell2[h] = idx < ncur ? cur[idx] : -1;
                }
#pragma unroll
                for (int h = 0; h < 2; h++)
                    if (cell2[h] >= 0) {
                        const float4* q = reinterpret_cast<const float4*>(cA + cell2[h]);
                        const float4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2), v3 = __ldg(q + 3);
                        a2[h].cx = v0.x; a2[h].cy = v0.y; a2[h].cz = v0.z; a2[h].len = v0.w;
                        a2[h].ilo[0] = v1.x; a2[h].ilo[1] = v1.y; a2[h].ilo[2] = v1.z; a2[h].ihi[0] = v1.w;
                        a2[h].ihi[1] = v2.x; a2[h].ihi[2] = v2.y; a2[h].olo[0] = v2.z; a2[h].olo[1] = v2.w;
                        a2[h].olo[2] = v3.x; a2[h].ohi[0] = v3.y; a2[h].ohi[1] = v3.z; a2[h].ohi[2] = v3.w;
                    }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    cls2[h] = 0;                                  // 0: nothing (no cell / empty cell), 1: accept, 2: open
                    if (cell2[h] >= 0 && a2[h].len >= 0.f) {
                        int c = classify_f32(a2[h], gf, theta_inv, margin);
                        if (c == 0) {                             // undecided in fp32: the exact rule on the fp64 record
                            const pb_tree_cell& cc = cells[cell2[h]];
                            const double len = cc.len;
                            const bool far_enough = box_dist2(gp->in_lo, gp->in_hi, cc.cm) > len * len * theta_inv2;
                            const bool touch = box_overlap(gp->out_lo, gp->out_hi, cc.in_lo, cc.in_hi) ||
                                               box_overlap(cc.out_lo, cc.out_hi, gp->in_lo, gp->in_hi);
                            c = (far_enough && !touch) ? 1 : 2;
                        }
                        cls2[h] = c;
                    }
                }
                // ---- compact chunk by chunk, in frontier order ----
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (base + 32 * h >= ncur) break;
                    const int cell = cell2[h];
                    int cls = cls2[h], first = 0, n = 0, nls = 0;
                    int child[8];
                    if (cls == 2) {
                        const int4* q = reinterpret_cast<const int4*>(cB + cell);
                        const int4 t = __ldg(q + 2);              // first, n, leaf, n_let_sp
                        n = t.y;
                        if (t.z) { first = t.x; nls = elem_map ? t.w : 0; }
                        else {
                            cls = 3;
                            const int4 c0 = __ldg(q), c1 = __ldg(q + 1);
                            child[0] = c0.x; child[1] = c0.y; child[2] = c0.z; child[3] = c0.w;
                            child[4] = c1.x; child[5] = c1.y; child[6] = c1.z; child[7] = c1.w;
                        }
                    }
                    // superparticles
                    const unsigned m1 = __ballot_sync(0xffffffffu, cls == 1);
                    if (FILL && cls == 1) {
                        const int at = nsp + __popc(m1 & ((1u << lane) - 1u));
                        if (at < cap_s) os[at] = cell;
                    }
                    nsp += __popc(m1);
                    // opened leaves
                    int tot;
                    const int off_e = warp_excl_scan(cls == 2 ? n - nls : 0, lane, tot);
                    if (elem_map == nullptr) {
                        if (FILL && cls == 2)
                            for (int k = 0; k < n; k++) if (nep + off_e + k < cap_e) oe[nep + off_e + k] = first + k;
                    } else {
                        int tot_s;
                        const int off_s = warp_excl_scan(cls == 2 ? nls : 0, lane, tot_s);
                        if (FILL && cls == 2) {
                            int ke = nep + off_e, ks = nsp + off_s;
                            for (int k = 0; k < n; k++) {
                                const int m = elem_map[first + k];
                                if (m >= 0) { if (ke < cap_e) oe[ke] = m; ke++; }
                                else        { if (ks < cap_s) os[ks] = n_cells + ~m; ks++; }
                            }
                        }
                        nsp += tot_s;
                    }
                    nep += tot;
                    // opened cells: children -> next frontier
                    int nch = 0;
                    if (cls == 3) {
#pragma unroll
                        for (int k = 0; k < 8; k++) nch += (child[k] >= 0);
                    }
                    const int off_c = warp_excl_scan(nch, lane, tot);
                    if (nnext + tot > cap) { if (lane == 0) atomicExch(overflow, 1); tot = 0; nch = 0; }
                    if (cls == 3 && nch > 0) {
                        int w = nnext + off_c;
#pragma unroll
                        for (int k = 0; k < 8; k++) if (child[k] >= 0) nxt[w++] = child[k];
                    }
                    nnext += tot;
                }
            }
            __syncwarp();
            int* t = cur; cur = nxt; nxt = t;
            ncur = nnext;
        }
        if (lane == 0) {
            if (counts) counts[g0 + g] = make_int2(nep, nsp);
            if (FILL && caps && (nep > cap_e || nsp > cap_s)) atomicExch(overflow + 1, 1);
        }
        __syncwarp();
    }
}

cudaError_t launch_compact_cells(cudaStream_t s, const void* cells, int n_cells, void* A, void* B) {
    if (n_cells <= 0) return cudaSuccess;
    compact_cells_kernel<<<(n_cells + 255) / 256, 256, 0, s>>>((const pb_tree_cell*)cells, n_cells, (CellA*)A, (CellB*)B);
    return cudaGetLastError();
}

cudaError_t launch_walk_c(cudaStream_t s, bool fill, const void* cells, const void* A, const void* B, const void* groups, int g0, int n_groups,
                          double theta_inv2, double coord_max, int2* counts, const int2* offs, int* id_e, int* id_s, int* scratch, int cap,
                          int n_ctas, int* overflow, const int* elem_map, int n_cells, const int2* caps) {
    if (n_groups <= 0) return cudaSuccess;
    const float theta_inv = (float)sqrt(theta_inv2);
    const float margin = (float)(coord_max * 2.4e-7 * 4.0);          // 2 sqrt(3) components of 2^-22 S, rounded up generously
    if (fill)
        walk_kernel_c<true><<<n_ctas, 128, 0, s>>>((const pb_tree_cell*)cells, (const CellA*)A, (const CellB*)B, (const pb_tree_group*)groups, g0, n_groups,
                                                   theta_inv2, theta_inv, margin, counts, offs, id_e, id_s, scratch, cap, overflow, elem_map, n_cells, caps);
    else
        walk_kernel_c<false><<<n_ctas, 128, 0, s>>>((const pb_tree_cell*)cells, (const CellA*)A, (const CellB*)B, (const pb_tree_group*)groups, g0, n_groups,
                                                    theta_inv2, theta_inv, margin, counts, nullptr, nullptr, nullptr, scratch, cap, overflow, elem_map, n_cells, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_walk_count(cudaStream_t s, const void* cells, const void* groups, int g0, int n_groups, double theta_inv2,
                              int2* counts, int* scratch, int cap, int n_ctas, int* overflow, const int* elem_map, int n_cells) {
    if (n_groups <= 0) return cudaSuccess;
    walk_kernel<false><<<n_ctas, 128, 0, s>>>((const pb_tree_cell*)cells, (const pb_tree_group*)groups, g0, n_groups, theta_inv2,
                                              counts, nullptr, nullptr, nullptr, scratch, cap, overflow, elem_map, n_cells, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_walk_fill(cudaStream_t s, const void* cells, const void* groups, int g0, int n_groups, double theta_inv2,
                             const int2* offs, int* id_e, int* id_s, int* scratch, int cap, int n_ctas, int* overflow,
                             const int* elem_map, int n_cells, const int2* caps, int2* counts) {
    if (n_groups <= 0) return cudaSuccess;
    walk_kernel<true><<<n_ctas, 128, 0, s>>>((const pb_tree_cell*)cells, (const pb_tree_group*)groups, g0, n_groups, theta_inv2,
                                             counts, offs, id_e, id_s, scratch, cap, overflow, elem_map, n_cells, caps);
    return cudaGetLastError();
}

} // namespace pb
