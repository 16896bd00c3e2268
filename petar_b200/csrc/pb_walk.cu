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
