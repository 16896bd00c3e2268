// pb_device.h — device-side data layout shared by the kernels (pb_kernels.cu) and the host
// engine (pb_engine.cu).  Nothing here is visible through the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb {

constexpr int kWarpsPerCta  = 8;                    // 256 threads
constexpr int kThreads      = kWarpsPerCta * 32;
constexpr int kTileJ        = kThreads;             // j staged per tile: one per thread
constexpr int kTilePairs    = kTileJ / 2;           // the inner loops consume j in pairs (packed fp32x2)
#ifndef PB_EP_UNROLL
#define PB_EP_UNROLL 4
#endif
#ifndef PB_SP_UNROLL
#define PB_SP_UNROLL 2
#endif
constexpr int kPairUnroll   = PB_EP_UNROLL;         // EP pairs per unrolled inner-loop step (tuning: tools/build_variants.py)
constexpr int kSpUnroll     = PB_SP_UNROLL;         // SP pairs per unrolled inner-loop step

// One walk (= FDPS i-group with its interaction lists) of a dispatch.  64 B.
struct __align__(16) Walk {
    int   i_off;          // first i-particle in the dispatch's epi array
    int   ni;             // number of i-particles
    int   ej_off;         // first entry of this walk's EP index list in the dispatch's id array
    int   nej;
    int   sj_off;         // first entry of this walk's SP index list
    int   nsj;
    float ohx, ohy, ohz;  // walk origin, hi part  (origin = oh + ol; all zero in absolute-coordinate mode)
    float olx, oly, olz;  // walk origin, lo part
    float hx, hy, hz;     // half extent of the i-particles' bounding box about the origin (+inf: no culling)
    float rsi2max;        // max r_search^2 over the walk's i-particles
};

// One CTA's work: a group of `nib` 32-wide i-blocks of one walk against a chunk of one of the
// walk's two j lists; `jsplit` warps share each i-block and split every j tile between them
// (nib * jsplit <= kWarpsPerCta).  48 B.
struct __align__(16) Task {
    int walk;
    int i_first;          // first i of the group, relative to the walk (multiple of 32)
    int nib;
    int jsplit;
    int kind;             // 0: EP-EP (linear cutoff + neighbour count), 1: EP-SP (quadrupole)
    int j_begin;          // chunk start within the walk's list
    int j_count;
    int part_base;        // partial-sum slot of (i-block 0, lane 0) for this task
    int blk0;             // index of the group's first i-block in the IBlock table (fused reduction: its chunk counter)
    int n_chunks;         // EP + SP chunks of the group = partial sums every one of its i-blocks receives
    int pad1, pad2;
};

// One 32-wide i-block for the reduction kernel.  32 B.
struct __align__(16) IBlock {
    int part_base;        // slot of (chunk 0, lane 0)
    int n_chunks;         // EP + SP chunks that produced partials for this i-block
    int stride;           // slots between consecutive chunks
    int out_off;          // index of lane 0 in the dispatch's output array
    int n_valid;          // lanes that are real particles
    int pad0, pad1, pad2;
};

static_assert(sizeof(Walk) == 64 && sizeof(Task) == 48 && sizeof(IBlock) == 32, "table records: sizes are part of the arena layout");

// Result record, identical to PeTar's ForceSoft (reference src/soft_ptcl.hpp:4-15) so that the
// D2H buffer can be memcpy'd straight into FDPS's force arrays.  40 B.
struct ForceOut {
    double  ax, ay, az, pot;
    int64_t n_ngb;
};

struct Params {
    float eps2;
    float rcut2;
    float rinv_cut;       // coords = 2: float(1 / sqrt(double(rcut2))), the CPU replay's 1/r inside the cutoff
    int   abs_mode;       // option "coords": 0 walk-relative two-float dx; 1 absolute float-cast dx for every pair;
                          // 2 walk-relative, but pairs that pass the neighbour test use the absolute float-cast dx
    int   i_f4;           // float4 per packed i-particle: 2, or 3 with the absolute float-cast position (coords = 2)
    // neighbour-list emission (count-only launches with option "nb_lists"): every pair that passes the
    // neighbour test appends the key (i_base + i index in the sub-batch) << 32 | j store index
    int   i_base;
    unsigned int pair_cap;
    unsigned long long* pairs;
    unsigned int* pair_cursor;
    int* meta;            // persistent launches (device-made plan): [0] task count, [4] task cursor
    // fused reduction (pb::force_kernel_ws, option "fuse_reduce"): the warp that delivers the LAST chunk of an i-block adds
    // the block's partial sums in chunk order and writes the forces, straight into page-locked host memory
    const IBlock* iblocks;
    int*   done;          // chunks delivered per i-block (returns to 0)
    ForceOut* out;        // device-visible address of the host result array
    double G;
};

// Two-float position relative to the walk origin: hi + lo = (xh + xl) - (oh + ol) to ~2^-46,
// hi = fl((xh - oh) + (xl - ol)).  The SAME code runs on the host for i-particles and in the
// kernel for j-particles, so a particle meeting itself gets bit-identical (hi, lo) and dx == 0.
__host__ __device__ inline void rel_hilo(float xh, float xl, float oh, float ol, float& hi, float& lo) {
    const float s  = xh - oh;
    const float bb = s - xh;
    const float e1 = (xh - (s - bb)) - (oh + bb);      // exact rounding error of s
    const float t  = xl - ol;
    hi = s + t;
    const float b2 = hi - s;
    const float e2 = (s - (hi - b2)) + (t - b2);        // exact rounding error of hi
    lo = e1 + e2;
}

// launchers (pb_kernels.cu)
cudaError_t launch_force(cudaStream_t s, int n_tasks, int nr_steps, int min_blocks,
                         const Walk* walks, const Task* tasks,
                         const float4* epi, const int* id_epj, const int* id_spj,
                         const float4* epj, const float4* spj,
                         double4* part4, int* partn, Params p, bool emit_pairs = false, bool two_i = false);

cudaError_t launch_force_persistent(cudaStream_t s, int n_ctas, int nr_steps,
                                    const Walk* walks, const Task* tasks, const float4* epi, const int* id_epj, const int* id_spj,
                                    const float4* epj, const float4* spj, double4* part4, int* partn, Params p, bool two_i);

// warp-specialised persistent force kernel (pb_kernels_ws.cu): 8 compute warps + 2 producer warps per CTA;
// two_i: SP tasks of groups with at least two i-blocks keep two i-particles per lane
cudaError_t launch_force_ws(cudaStream_t s, int n_ctas, int nr_steps,
                            const Walk* walks, const Task* tasks, const float4* epi, const int* id_epj, const int* id_spj,
                            const float4* epj, const float4* spj, double4* part4, int* partn, Params p, bool two_i);

// device-side i-particle preparation and task planning (pb_plan.cu)
cudaError_t launch_iprep(cudaStream_t s, const void* groups, int n_groups, const int* i_first, const int2* counts, const int2* offs,
                         const float4* epj, Walk* walks, float4* epi, int i_f4, int coords, int cull);
cudaError_t launch_devplan(cudaStream_t s, const void* groups, int n_groups, const int* i_first, const int2* counts, int U, int Us,
                           int3* goff, int* meta, int cap_tasks, long long cap_part, Task* tasks, IBlock* iblocks, const int2* caps, ForceOut* out_fused);
// device-side packing of raw host arrays (pb_pack.cu, option "raw_upload")
cudaError_t launch_pack_epj(cudaStream_t s, const void* raw, size_t stride, size_t off_pos, size_t off_mass, size_t off_rs, int n, float4* out);
cudaError_t launch_pack_spj(cudaStream_t s, const void* raw, size_t stride, size_t off_pos, size_t off_mass, size_t off_quad, int has_quad, int n, float4* out);
cudaError_t launch_gather_epj(cudaStream_t s, const float4* epj, int n_epj, const int* idx, int n, float4* out, int* err);
cudaError_t launch_expand_runs(cudaStream_t s, const int2* runtab, const int2* runs, const Walk* walks, int n_walk, int* out);
void plan_sizes_host(const int* ni, const int2* counts, int n_groups, int U, int Us, long long* n_tasks, long long* n_part, long long* n_iblk);

cudaError_t launch_reduce(cudaStream_t s, int n_iblocks, const IBlock* iblocks,
                          const double4* part4, const int* partn, ForceOut* out, double G, const int* meta = nullptr);

// device-side list building (pb_walk.cu); cells / groups are pb_tree_cell / pb_tree_group arrays
cudaError_t launch_walk_count(cudaStream_t s, const void* cells, const void* groups, int g0, int n_groups, double theta_inv2,
                              int2* counts, int* scratch, int cap, int n_ctas, int* overflow,
                              const int* elem_map = nullptr, int n_cells = 0);
cudaError_t launch_walk_fill(cudaStream_t s, const void* cells, const void* groups, int g0, int n_groups, double theta_inv2,
                             const int2* offs, int* id_e, int* id_s, int* scratch, int cap, int n_ctas, int* overflow,
                             const int* elem_map = nullptr, int n_cells = 0, const int2* caps = nullptr, int2* counts = nullptr);

// compact walk records (64-B fp32 + 48-B int per cell) and the walk that classifies on them first (pb_walk.cu)
cudaError_t launch_compact_cells(cudaStream_t s, const void* cells, int n_cells, void* A, void* B);
cudaError_t launch_walk_c(cudaStream_t s, bool fill, const void* cells, const void* A, const void* B, const void* groups, int g0, int n_groups,
                          double theta_inv2, double coord_max, int2* counts, const int2* offs, int* id_e, int* id_s, int* scratch, int cap,
                          int n_ctas, int* overflow, const int* elem_map, int n_cells, const int2* caps);

// changeover correction (pb_corr.cu): the fields of one neighbour / one corrected particle, fp64
struct CorrJ { double x, y, z, mass, r_in, r_out, mass_bk, status; long long id; };       // 72 B
struct CorrI { CorrJ p; double ax, ay, az, pot_tot, pot_soft; };                            // 112 B
struct CorrOut { double ax, ay, az, pot_tot, pot_soft; };                                   // 40 B
struct CorrParams { double eps2, r_out, G, status_no_cm; int replay_fp32; };
cudaError_t launch_corr(cudaStream_t s, int n_i, const CorrI* pi, const CorrJ* pj,
                        const int* nb_off, const int* nb_idx, CorrOut* out, CorrParams prm);

} // namespace pb
