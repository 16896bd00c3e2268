/* petar_b200.h — C ABI of the B200-native soft-force engine (libpetar_b200.so).
 *
 * This is the drop-in boundary for PeTar's long-range soft-force hot path: the device side of
 * FDPS `calcForceAllAndWriteBackMultiWalkIndex(dispatch, retrieve, ...)` that the reference
 * implements in src/force_gpu_cuda.cu.  Every entry point takes plain pointers and sizes only
 * (no FDPS, PeTar, CUDA or torch types); particle arrays are described by base pointer +
 * stride + field offsets (pb_layout_*) so any build of PeTar's EPISoft / EPJSoft / SPJ /
 * ForceSoft binds without recompiling this library.
 *
 * Which reference interface each entry point replaces:
 *
 *   pb_init             device selection + lazy allocation   src/force_gpu_cuda.cu:549-571
 *   pb_set_params       functor state eps2, rcut2, G         src/force_gpu_cuda.hpp:103-118
 *   pb_upload_j         dispatch(..., send_flag=true)        src/force_gpu_cuda.cu:577-620
 *   pb_dispatch_index   dispatch(..., send_flag=false)       src/force_gpu_cuda.cu:621-699
 *   pb_dispatch_direct  CalcForceWithLinearCutoffCUDA        src/force_gpu_cuda.cu:704-827
 *   pb_retrieve         RetrieveForceCUDA                    src/force_gpu_cuda.cu:831-880
 *   pb_get_profile      gpu_profile / gpu_counter            src/force_gpu_cuda.hpp:8-92,
 *                                                            src/force_gpu_cuda.cu:610-617,672-697,836-863
 *
 * Call protocol (same as FDPS drives the reference): once per tree step
 * pb_upload_j; then per walk group pb_dispatch_index (returns without waiting for the GPU; all
 * host inputs have been consumed when it returns) followed later by pb_retrieve for that
 * dispatch.  One dispatch may be outstanding at a time (FDPS tag_max = 1).
 *
 * All functions return 0 on success and a negative pb_status on failure; pb_last_error() then
 * describes it.  Nothing here falls back to a CPU path: without a usable CUDA device every
 * call fails with PB_ERR_NO_DEVICE.
 */
#ifndef PETAR_B200_H
#define PETAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_ABI_VERSION 6

typedef enum pb_status {
    PB_OK = 0,
    PB_ERR_NO_DEVICE = -1,   /* no CUDA device / driver */
    PB_ERR_CUDA      = -2,   /* a CUDA runtime call failed (message in pb_last_error) */
    PB_ERR_ARG       = -3,   /* invalid argument */
    PB_ERR_PROTOCOL  = -4,   /* call order violated (e.g. dispatch before upload_j) */
    PB_ERR_NCCL      = -5
} pb_status;

/* ---- array layouts: byte stride of one element and byte offsets of the fields used ------ */
typedef struct pb_layout_epi {      /* i-particles: EPISoft (reference src/soft_ptcl.hpp:271-278) */
    size_t stride;
    size_t off_pos;                 /* 3 x double */
    size_t off_rsearch;             /* double */
} pb_layout_epi;

typedef struct pb_layout_epj {      /* j-particles: EPJSoft (reference src/soft_ptcl.hpp:311-326) */
    size_t stride;
    size_t off_pos;                 /* 3 x double */
    size_t off_mass;                /* double */
    size_t off_rsearch;             /* double */
} pb_layout_epj;

typedef struct pb_layout_spj {      /* superparticles: PS::SPJQuadrupoleInAndOut / SPJMonopoleInAndOut */
    size_t stride;
    size_t off_pos;                 /* 3 x double */
    size_t off_mass;                /* double */
    size_t off_quad;                /* 6 x double in the order xx,yy,zz,xy,xz,yz; ignored if !has_quad */
    int    has_quad;                /* 0: monopole only (quadrupole terms are zero) */
} pb_layout_spj;

typedef struct pb_layout_force {    /* results: ForceSoft (reference src/soft_ptcl.hpp:4-15) */
    size_t stride;
    size_t off_acc;                 /* 3 x double, ASSIGNED */
    size_t off_pot;                 /* double,     ASSIGNED */
    size_t off_nngb;                /* int64,      ASSIGNED */
} pb_layout_force;

/* ---- profile / counters: the 4 timers + 5 counters of the reference's GPUProfile/GPUCounter,
 *      same meaning, in seconds / counts accumulated since the last reset ------------------ */
typedef struct pb_profile {
    double t_copy;                  /* host-side pack / unpack (gpu_profile.copy) */
    double t_send;                  /* H2D (gpu_profile.send), device-timed */
    double t_recv;                  /* D2H (gpu_profile.recv), device-timed */
    double t_calc;                  /* kernels (gpu_profile.calc), device-timed */
    long long n_walk, n_epi, n_epj, n_spj, n_call;   /* gpu_counter.* */
    long long n_interaction_ep;     /* sum n_epi*n_epj  (PeTar Ep-Ep_sum, src/petar.hpp:943-946) */
    long long n_interaction_sp;     /* sum n_epi*n_spj  (PeTar Ep-Sp_sum) */
    long long n_kernel_launch;      /* kernels launched by this library */
    long long h2d_bytes, d2h_bytes;
    /* host sub-phases of t_copy (seconds): task planning, packing into pinned staging, result
     * scatter; and the time spent enqueueing copies/kernels (not part of t_copy) */
    double t_plan, t_pack, t_unpack, t_enqueue;
    /* device-timed idle between walk groups: from the end of one dispatch's last kernel to the first
     * kernel of the next dispatch (the tag_max = 1 protocol drains the GPU at every retrieve) */
    double t_gap;
} pb_profile;

/* ---- lifetime ----------------------------------------------------------------------------- */
/* device < 0 selects my_rank % cudaGetDeviceCount(), as the reference does (:550-553).
 * Idempotent; pb_upload_j / pb_dispatch_* call it lazily with (0,-1) if it was never called. */
int  pb_init(int my_rank, int device);
void pb_finalize(void);
int  pb_abi_version(void);
const char* pb_last_error(void);

/* ---- parameters --------------------------------------------------------------------------- */
/* eps2 = EPISoft::eps^2, rcut2 = EPISoft::r_out^2, G = ForceSoft::grav_const. */
int  pb_set_params(double eps2, double rcut2, double G);

/* Options (key, value):
 *   "coords"     how dx of an EP-EP pair is formed (fp32 in every mode):
 *                0: positions relative to each walk's origin (centre of its i-particles), formed from fp64 on the
 *                   host (i) and from a hi/lo fp32 split on the device (j) — the most accurate kernel-level result;
 *                   pair it with an all-double correction (pb_correct_changeover with replay_fp32 = 0, or PeTar built
 *                   with the `#else` branch of src/hard.hpp:1443-1453);
 *                1: absolute coordinates cast to fp32 for EVERY pair — dx = float(xj) - float(xi), the exact
 *                   arithmetic of the reference kernel (src/force_gpu_cuda.cu:58-60);
 *                2 (default, the drop-in mode): as 0, except that a pair passing the neighbour test
 *                   r^2 < max(rs_i, rs_j)^2 is evaluated from dx = float(xj) - float(xi).  Those are the pairs PeTar's
 *                   CPU changeover correction visits afterwards, and with USE_GPU it removes the kernel's clamped
 *                   term by re-computing it in float from exactly that dx (`dr_32`, src/hard.hpp:1428-1442); with
 *                   any other dx a residual G m (dx - dr_32) / r_out^3 per neighbour would stay in the force
 *                   (tests/test_gpu_replay_gap.py measures it).  Neighbour counts are the same in all modes.
 *   "streams"    number of CUDA streams one dispatch is split across (default 8, 1..8).
 *   "jchunk"     target EP j-chunk per warp-task (default 0 = automatic).
 *   "nr"         Newton-Raphson steps after MUFU.RSQ (0 default, or 1).
 *   "cull"       1 (default): skip the neighbour test for j-tile segments that cannot reach any
 *                i-particle of the walk (results identical); 0: test every pair.
 *   "occupancy"  resident CTAs per SM the force kernel is compiled for: 2 (default, 124 registers)
 *                or 3 (80 registers).
 *   "lead"       the first of a dispatch's per-stream sub-batches is 1/(1+lead) the size of the others
 *                (the GPU idles until its copy lands): 0 = equal sizes, 3 (default), up to 15.
 *   "tree_batch" groups per force launch of pb_tree_force (default 256: 33.7 ms per step at N = 1e6 against 40 ms with 1024).
 *   "tree_fill"  pb_tree_force writes the lists with 0 (default): one step-wide launch, 1: one launch per batch on
 *                the batch's stream (measured: no gain, the force kernels own the SMs).
 *   "tree_streams"  streams pb_tree_force cycles its force launches over (default 4, measured best with tree_batch 256).
 *   "tree_spec"  1 (default): from the second tree step on, pb_tree_upload reserves list space from the previous
 *                step's list lengths (+12.5 % + 64 entries) and fills the lists in ONE walk pass, while the host still
 *                packs j; a list that outgrows its reservation is detected and pb_tree_force redoes the exact
 *                two-pass fill.  0: always count, then fill.
 *   "walk_ctas"  CTAs (4 warps each) of the tree-walk launches of pb_tree_upload / pb_tree_force, 1..2368 (default
 *                2368 = every warp slot of the GPU: the walk is latency-bound; 592 costs 2.3 ms per step at N = 1e6).
 *   "min_slot_work"  a dispatch is not cut into per-stream sub-batches smaller than this many EP-equivalent
 *                interactions (n_epi * (n_epj + 2 n_spj)); default 0 = always "streams" (measured: 4e7 saves enqueue time
 *                at 8 ranks per node but costs more pipelining than it saves at 4).
 *   "ep_runs"    1: pb_dispatch_index / pb_dispatch_count_index ship every walk's EP index list as maximal runs of
 *                consecutive indices, (start, length) pairs found while the walk is packed, and a small kernel writes the
 *                indices back out on the device — FDPS's EP lists are leaf cells in Morton order (~20 indices per run at
 *                N = 1e6: 8 B per run instead of 4 B per index); forces are bit-identical.  SP lists (tree cell numbers, ~1.6
 *                per run in depth-first numbering) stay plain indices.  0 (default): EP lists are copied as they are —
 *                measured at N = 1e6 the runs cut the H2D volume of a tree step from 0.76 to 0.57 GB but make the step 3.8 ms
 *                SLOWER (47.0 vs 43.2 ms): the PCIe time saved was hidden under the kernels anyway, while the extra
 *                expansion launch per stream and dispatch (448 per step) sits on the critical path of the tag_max = 1
 *                protocol.  The device-resident tree step removes the lists altogether instead.
 *   "raw_upload" 1: pb_upload_j / pb_upload_j_range page-lock the caller's arrays once (cudaHostRegister), copy them as they
 *                are and pack them on the device (bit-identical to the host packing) — for several ranks per node, where the
 *                host cores (4 per rank on an 8-GPU box) are scarcer than PCIe bandwidth: 2.7 -> 0.3 ms of host time per
 *                tree step at 8 ranks.  The arrays must then stay unchanged until the step's forces are back.  0 (default):
 *                packed on the host into pinned staging; the arrays are consumed when the call returns.  Setting raw_upload
 *                or raw_result to 0 releases every page-lock taken so far: do that (or pb_finalize) before freeing or moving
 *                an array the library has page-locked.
 *   "ws"         1 (default): persistent force launches (the device-resident tree step) run the warp-specialised kernel —
 *                8 compute warps that only wait for tiles and run the pair loops, 2 producer warps that fetch tasks and
 *                stage j tiles four deep (pb_kernels_ws.cu); 0: every warp stages and computes (pb::force_kernel).
 *   "fuse_reduce" 1 (default): the warp that delivers the last partial sum of a 32-particle block adds the block's partial sums
 *                (fixed chunk order: bitwise equal to the separate reduction kernel) and writes the 32 force records itself —
 *                into page-locked HOST memory, so neither a reduction kernel nor a D2H copy follows the force kernel.
 *                Device-resident tree step: -0.9 ms per step at N = 1e6; functor path (force dispatches, not the neighbour
 *                search): 46.0 -> 43.2 ms per tree step (two launches / copies fewer per stream and dispatch), at +0.7 %
 *                kernel time.  0: reduction kernel + D2H copy.
 *   "raw_result" 1: pb_tree_force_resident page-locks the caller's force array once (it must be the same array from step to
 *                step, as FDPS's is, and in the plain ForceSoft layout) and the kernel writes into it directly; 0
 *                (default): into the library's pinned buffer, copied to the caller's array by the host threads.
 *   "sp2i"       1 (default): EP-SP tasks of i-groups with at least two 32-particle blocks keep TWO i-particles per lane, so
 *                every superparticle pair read from shared memory serves four interactions (+6 % on the SP loop, -3.6 %
 *                kernel time per tree step); 0: one i-particle per lane everywhere.  Sums differ from sp2i = 0 only in
 *                the order in which a block's j are spread over warps (both are within the fp32 tolerance of the oracle).
 *   "chunk_tile" 1 (default): the j chunks of the task plan are whole 256-entry tiles, so only the last chunk of a list ends
 *                in a ragged tile; 0: equal chunks in multiples of 8 entries (the round-1 plan, kept for A/B runs).
 *   "nb_lists"   1: pb_dispatch_count_index also collects the neighbour PAIRS (see pb_retrieve_neighbors);
 *                0 (default): counts only.
 * Returns PB_ERR_ARG for an unknown key or value. */
int  pb_set_option(const char* key, long long value);
/* Current value of an option (any key of pb_set_option). */
int  pb_get_option(const char* key, long long* value);

/* ---- the hot path ------------------------------------------------------------------------- */
/* send_flag == true: publish all j of this tree step.  Replaces any previous j set. */
int  pb_upload_j(const void* epj, int n_epj, const pb_layout_epj* lepj,
                 const void* spj, int n_spj, const pb_layout_spj* lspj);

/* send_flag == false: enqueue n_walk walks.  epi[iw] -> n_epi[iw] i-particles, id_epj[iw] /
 * id_spj[iw] -> indices into the arrays given to pb_upload_j.  Asynchronous w.r.t. the GPU. */
int  pb_dispatch_index(int n_walk,
                       const void* const* epi, const int* n_epi, const pb_layout_epi* lepi,
                       const int* const* id_epj, const int* n_epj,
                       const int* const* id_spj, const int* n_spj);

/* Non-index mode: per-walk j arrays instead of indices (reference :704-827). */
int  pb_dispatch_direct(int n_walk,
                        const void* const* epi, const int* n_epi, const pb_layout_epi* lepi,
                        const void* const* epj, const int* n_epj, const pb_layout_epj* lepj,
                        const void* const* spj, const int* n_spj, const pb_layout_spj* lspj);

/* Neighbour search only: for PeTar's second tree, tree_nb, whose functor
 * SearchNeighborEpEpSimd runs on the CPU every step even in GPU builds (reference
 * src/petar.hpp:767-788, src/soft_force.hpp:11-34, 239-283).  Counts j with
 * r^2 < max(rs_i, rs_j)^2 (no eps, as SearchNeighborEpEpNoSimd) over each walk's EP list.
 * Retrieve with pb_retrieve: only n_ngb is assigned, acc / pot are left untouched. */
int  pb_dispatch_count_index(int n_walk,
                             const void* const* epi, const int* n_epi, const pb_layout_epi* lepi,
                             const int* const* id_epj, const int* n_epj);

/* Wait for the outstanding dispatch and ASSIGN force[iw][i].{acc,pot,n_ngb}, i < ni[iw].
 * n_walk / ni must equal those of the dispatch being retrieved. */
int  pb_retrieve(int n_walk, const int* ni, void* const* force, const pb_layout_force* lforce);

/* ---- direct-sum field query ---------------------------------------------
 * acc and potential at n_points arbitrary positions from n_ptcl particles, eps = 0, no cutoff, G
 * included, particles with mass <= 0 skipped: what the AMUSE worker's get_gravity_at_point /
 * get_potential_at_point compute with CalcForcePPSimd per point on the CPU (reference
 * amuse-interface/interface.cc:966-1030, src/soft_force.hpp:285-344).  `ptcl` is described by
 * stride + offsets of pos (3 doubles) and mass (double), e.g. PeTar's FPSoft array.  Outputs are
 * ASSIGNED; any of ax/ay/az/pot may be NULL.  REPLACES the content of the j store and leaves it unpublished: the next
 * pb_dispatch_* / pb_tree_force* without a fresh pb_upload_j fails with PB_ERR_PROTOCOL, and device pointers obtained
 * from pb_reserve_j before the call are invalid (FDPS re-publishes j at the start of every tree step anyway). */
int  pb_field_at_points(const double* x, const double* y, const double* z, int n_points,
                        const void* ptcl, int n_ptcl, size_t stride, size_t off_pos, size_t off_mass,
                        double G, double* ax, double* ay, double* az, double* pot);

/* ---- device-side interaction lists (beyond the drop-in: needs the tree) ------
 * Instead of receiving per-walk index lists from the host (0.7 GB per tree step at N = 1e6), the
 * library is given the TREE once per step and builds id_epj / id_spj for every i-group on the GPU
 * with the same opening rule FDPS's QuadrupoleWithSymmetrySearch walk applies,
 *     open(cell) <=> dist^2(group particle box, cell c.m.) <= (cell length / theta)^2
 *                    or group search box touches cell particle box, or vice versa,
 * evaluated in fp64 without FMA contraction (decisions identical to the host walk), then runs the
 * same force kernels on the lists without them ever crossing PCIe.
 * Requirements: pb_upload_j (or pb_reserve_j / pb_upload_j_range / pb_publish_j) published the j of this
 * step with SP store slot c == multipole of cell c.  Single domain (pb_tree_upload): the EP store order
 * is the tree's sorted particle order.  With a local essential tree (pb_tree_upload_let): the tree's
 * leaves also hold the EP and SP received from other domains, and `elem_map` says where each sorted
 * tree element lives — elem_map[k] >= 0: EP store index; < 0: the LET superparticle in SP store slot
 * n_cells + ~elem_map[k] — so the store may be in any order (e.g. local particles first, LET entries
 * behind them as they arrive over NCCL). */
typedef struct pb_tree_cell {          /* 176 B */
    double cm[3];                      /* centre of mass = expansion centre of the cell's superparticle */
    double len;                        /* geometric cell length */
    double in_lo[3], in_hi[3];         /* box of the particles inside */
    double out_lo[3], out_hi[3];       /* box of their search spheres (pos +- 0.99 r_search) */
    int    child[8];                   /* cell ids, -1 = none */
    int    first, n;                   /* element range in the tree's sorted order */
    int    leaf;
    int    n_let_sp;                   /* leaves: how many of the n elements are LET superparticles (0 without LET) */
} pb_tree_cell;

typedef struct pb_tree_group {         /* 104 B: one i-group ("walk") */
    int    first, n;                   /* its particles: sorted positions [first, first + n) */
    double in_lo[3], in_hi[3], out_lo[3], out_hi[3];
} pb_tree_group;

/* Per tree step: pb_tree_upload, pb_upload_j (either order), then pb_tree_force.  pb_tree_upload
 * returns as soon as the copy and the walk's counting pass are queued, so calling it FIRST lets the
 * GPU count list lengths while the host packs the j-particles. */
int  pb_tree_upload(const pb_tree_cell* cells, int n_cells, const pb_tree_group* groups, int n_groups, double theta);
int  pb_tree_upload_let(const pb_tree_cell* cells, int n_cells, const pb_tree_group* groups, int n_groups, double theta,
                        const int* elem_map, int n_elem);
/* Optional: pinned staging for the tree.  Returns buffers for n_cells cells and n_groups groups that stay valid until
 * the next pb_tree_stage / pb_finalize; write the tree into them (e.g. while converting FDPS's cells) and pass these
 * very pointers to pb_tree_upload / pb_tree_upload_let, which then skip their own 176 B-per-cell staging copy. */
int  pb_tree_stage(int n_cells, int n_groups, pb_tree_cell** cells, pb_tree_group** groups);
/* Forces on all i-particles, given in group order (group 0's particles first, ...); ASSIGNS
 * force[k].{acc,pot,n_ngb}.  Synchronous.  Both arrays are contiguous with the given layouts. */
int  pb_tree_force(const void* epi, const pb_layout_epi* lepi, void* force, const pb_layout_force* lforce);
/* The same step, device-resident: the i-particles are NOT passed — group g's i-particles are EP store slots
 * [groups[g].first, groups[g].first + groups[g].n) (upload the local particles in i-group order, as FDPS's epi_sorted
 * is) — and i-particle preparation, task planning, forces and reduction all run on the GPU (pb_plan.cu, one persistent
 * force launch).  From the second tree step on no host round trip sits between the uploads and the result: list
 * space, task and partial-sum buffers are reserved from the previous step's sizes, the true sizes come back with the
 * forces, and a step whose reservation was too small is detected on the device and repeated exactly.  ASSIGNS
 * force[k].{acc,pot,n_ngb} in group order.  Synchronous. */
int  pb_tree_force_resident(void* force, const pb_layout_force* lforce);
/* Device timeline of the last pb_tree_force_resident step, milliseconds between consecutive CUDA events on the step's
 * stream: ms[0] tree walk (list building), [1] wait for the j store (local upload, collectives), [2] i-particle
 * preparation + task planning, [3] force kernel, [4] reduction, [5] result copy to the host. */
int  pb_tree_timeline(float* ms, int n);
/* Test hook: the lists the last pb_tree_force built.  n_ep/n_sp: per group counts (n_groups each);
 * id_ep/id_sp: concatenated lists in group order, at most cap_* entries are written. */
int  pb_tree_lists(int* n_ep, int* n_sp, int* id_ep, long long cap_ep, int* id_sp, long long cap_sp);

/* Test hook, host only (needs no device): the task plan the library would make for ONE sub-batch of n_walk
 * walks with the given list lengths, as `n_streams_active` sub-batches run concurrently.  walks_out: 6 ints
 * per walk {i_off, ni, ej_off, nej, sj_off, nsj}; tasks_out: 10 ints per task {walk, i_first, nib, jsplit, kind,
 * j_begin, j_count, part_base, blk0, n_chunks} (at most cap_tasks are written); iblocks_out: 5 ints per 32-wide i-block
 * {part_base, n_chunks, stride, out_off, n_valid} (at most cap_iblocks).  Returns the number of tasks, the
 * number of i-blocks in *n_iblocks and the number of partial-sum slots in *n_part. */
int  pb_debug_plan(int n_walk, const int* n_epi, const int* n_epj, const int* n_spj, int n_streams_active,
                   int* walks_out, int* tasks_out, int cap_tasks, int* iblocks_out, int cap_iblocks,
                   int* n_iblocks, long long* n_part);

/* ---- profiling ---------------------------------------------------------------------------- */
int  pb_get_profile(pb_profile* out, int reset);

/* ---- device-resident replay (measurement only) --------------------------------------------
 * Between pb_record_begin and pb_record_end every dispatch also keeps its packed inputs
 * resident in HBM.  pb_replay re-launches the kernels of all recorded dispatches n_iter times
 * with no host<->device traffic, timed with CUDA events on the launching stream; it returns the
 * mean milliseconds per iteration (all recorded dispatches) in *ms_total, and the part spent in
 * the force kernel alone in *ms_force (either pointer may be NULL). */
int  pb_record_begin(void);
int  pb_record_end(void);
int  pb_replay(int n_iter, float* ms_total, float* ms_force);
/* number of force-kernel / reduce-kernel launches one replay iteration performs */
int  pb_replay_launches(void);

/* ---- multi-GPU: local-essential-tree j exchanged in device format --------------------------
 * The j store can be assembled from a host part (this rank's own j, packed and uploaded by
 * the library) and device-resident parts in the library's device j format that a collective
 * (NCCL over NVLink) wrote directly into the store.
 *
 * Device j formats (little-endian fp32):
 *   EPJ: 32 B = float4{x_hi,y_hi,z_hi,mass}, float4{x_lo,y_lo,z_lo,r_search}
 *   SPJ: 64 B = float4{x_hi,y_hi,z_hi,mass}, float4{x_lo,y_lo,z_lo,q'xx},
 *               float4{q'yy,q'zz,q'xy,q'xz},  float4{q'yz,trace,0,0}
 * with x = x_hi + x_lo (x_hi = float(x), x_lo = float(x - x_hi)) and q' = 3 q - trace * I the
 * traceless form of the raw second-moment tensor q (formed in fp64 before the cast). */
#define PB_EPJ_DEV_BYTES 32
#define PB_SPJ_DEV_BYTES 64

/* Size the store for n_epj/n_spj entries and return its device pointers (valid until the next
 * pb_reserve_j / pb_upload_j / pb_finalize). */
int  pb_reserve_j(int n_epj, int n_spj, void** d_epj, void** d_spj);
/* Pack host j (as pb_upload_j) into store slots [epj_first, epj_first+n_epj) and
 * [spj_first, spj_first+n_spj) of a store sized by pb_reserve_j. */
int  pb_upload_j_range(const void* epj, int epj_first, int n_epj, const pb_layout_epj* lepj,
                       const void* spj, int spj_first, int n_spj, const pb_layout_spj* lspj);
/* Make the upload stream wait until work queued so far on `cuda_stream` (a cudaStream_t, e.g.
 * the stream a NCCL collective writing into the store was enqueued on) has completed, and mark
 * the j store as published. */
int  pb_publish_j(void* cuda_stream);
/* LET send rows gathered ON THE DEVICE: d_out32[k] = EP store row idx[k] (device j format, 32 B), k < n.  `idx` is a
 * host array of EP store slots (the local particles a peer's domain needs, as FDPS's LET selection names them);
 * `d_out32` a device buffer (e.g. the send buffer of the NCCL all-to-all).  Queued on the library's upload stream,
 * behind pb_upload_j_range's copy of those particles: 4 B per row cross PCIe instead of a host-packed 32 B row.
 * With option raw_upload the list is copied from where it lies (page-locked once; it must stay unchanged until the
 * step's forces are back) and the indices are checked by the gather kernel: an index outside the store makes the next
 * call that synchronises with the device (pb_tree_force_resident, pb_retrieve, the next pb_let_gather_epj) fail with
 * PB_ERR_ARG; otherwise they are checked here. */
int  pb_let_gather_epj(const int* idx, int n, void* d_out32);
/* LET send rows of the other kind: d_out64[k] = pack(spj[k]) (device j format, 64 B), k < n, from a HOST array of
 * superparticles (the local tree's multipoles a peer's domain needs) into a DEVICE buffer, on the upload stream.
 * Option raw_upload: the array is copied as it is and packed on the device; otherwise packed here into pinned staging. */
int  pb_let_pack_spj(const void* spj, int n, const pb_layout_spj* l, void* d_out64);
/* Make `cuda_stream` (a cudaStream_t, e.g. the stream the collective is enqueued on) wait for everything queued so
 * far on the library's upload stream (pb_upload_j_range, pb_let_gather_epj). */
int  pb_stream_wait_upload(void* cuda_stream);
/* Pack host j to the device format into caller-provided HOST buffers (for building LET send
 * buffers). */
int  pb_pack_epj_host(const void* epj, int n, const pb_layout_epj* l, void* out32);
/* Same, gathering: out32[k] = pack(epj[idx[k]]), k < n (LET send lists are index lists into the
 * local particle array). */
int  pb_pack_epj_host_indexed(const void* epj, const long long* idx, int n, const pb_layout_epj* l, void* out32);
int  pb_pack_spj_host(const void* spj, int n, const pb_layout_spj* l, void* out64);

/* Neighbour lists of the last retrieved pb_dispatch_count_index (option "nb_lists" = 1): what
 * tree_nb.getNeighborListOneParticle would return for every i-particle of that dispatch (reference call
 * sites src/hard.hpp:1663, src/search_cluster.hpp), the particle itself included.  CSR over the
 * i-particles in dispatch order (walk 0's particles first): nb_off has sum(n_epi) + 1 entries, nb_idx holds
 * indices into the epj array of pb_upload_j, ascending within each list.  Call with nb_idx = NULL to learn
 * *n_pairs first.  Valid until the next count dispatch. */
int  pb_retrieve_neighbors(long long* n_pairs, int* nb_off, int* nb_idx, long long cap);

/* ---- changeover correction of the soft force ----------------------------------
 * Replaces the particle loop of SystemHard::correctForceWithCutoffTreeNeighborOMP (reference
 * src/hard.hpp:3366-3377): for every particle i, correctForceWithCutoffTreeNeighborOneParticleImp
 * (:1655-1691) — the self-potential term, then calcAccPotShortWithLinearCutoff (:1408-1476) for every
 * neighbour whose id differs, in list order.  acc, pot_tot and pot_soft of the i-particles are updated in
 * place, bit-identical to the reference function applied in the same order (fp64 on the device).
 *
 * i-particles (FPSoft) and neighbours (EPJSoft) are described by field offsets: pos (3 doubles), mass,
 * the changeover radii (ChangeOver::r_in_/r_out_ or EPJSoft::r_in/r_out), id (int64) and
 * group_data.artificial = {mass_backup, status} (doubles); off_acc / off_pot_tot / off_pot_soft are
 * used for the i side only.  Neighbour lists are CSR: nb_idx[nb_off[i] .. nb_off[i+1]) index ptcl_j.
 *
 * replay_fp32 = 1: the linear-cutoff term is re-evaluated in float from absolute coordinates, the
 *   reference's `USE_GPU` branch (use with option "coords" = 1, whose force kernel computes that term
 *   from the same float coordinates);  replay_fp32 = 0: all in double, the reference's `#else` branch
 *   (use with the default coordinates: the force kernel's own term is accurate to fp32 rounding of the
 *   true separation, so the exact term is what cancels it best).
 * status_no_cm: the status value of a group member without c.m. particle, -PS::LARGE_FLOAT
 *   (= -FLT_MAX/16, src/ptcl.hpp:214-221). */
typedef struct pb_layout_corr {
    size_t stride;
    size_t off_pos, off_mass, off_r_in, off_r_out, off_id, off_mass_backup, off_status;
    size_t off_acc, off_pot_tot, off_pot_soft;
} pb_layout_corr;

typedef struct pb_corr_params {
    double eps2, r_out, G;      /* EPISoft::eps^2, EPISoft::r_out, ForceSoft::grav_const */
    double status_no_cm;
    int    replay_fp32;
} pb_corr_params;

int  pb_correct_changeover(int n_i, void* ptcl_i, const pb_layout_corr* li,
                           int n_j, const void* ptcl_j, const pb_layout_corr* lj,
                           const int* nb_off, const int* nb_idx, const pb_corr_params* prm);

#ifdef __cplusplus
}
#endif
#endif /* PETAR_B200_H */
