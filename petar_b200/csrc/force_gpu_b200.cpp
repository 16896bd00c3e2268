// force_gpu_b200.cpp — the object that replaces PeTar's build/force_gpu_cuda.o
// (reference Makefile.in:179, 266-267).
//
// It DEFINES the symbols PeTar's src/force_gpu_cuda.hpp DECLARES — the dispatch functor
// CalcForceWithLinearCutoffCUDAMultiWalk::operator() (index mode, the mode `configure
// --enable-cuda` builds), CalcForceWithLinearCutoffCUDA::operator() (non-index mode),
// RetrieveForceCUDA and the globals gpu_profile / gpu_counter — and forwards to the C ABI of
// libpetar_b200.so (include/petar_b200.h).  This is the only translation unit that sees FDPS /
// PeTar types; nothing but pointers, counts and field offsets crosses into the library.
//
// Build inside PeTar (FDPS + SDAR present), same macros as the reference object gets
// (Makefile.in:28-30, 175-181):
//   mpicxx -std=c++17 -O2 -fopenmp -DUSE_GPU -DGPU_PROFILE -DUSE_QUAD
//          -DPARTICLE_SIMULATOR_GPU_MULIT_WALK_INDEX -I<FDPS>/src -I<SDAR>/src -I<PeTar>/src
//          -I<petar-b200>/include -c force_gpu_b200.cpp -o build/force_gpu_cuda.o
// and add  -L<petar-b200>/petar_b200/lib -lpetar_b200  to the link line.
//
// Build stand-alone (this repository's tests): -DPB_STANDALONE_MIRRORS, see petar_b200/build.py.
#include <cstdio>
#include <cstdlib>
#include <cstddef>

#ifdef PB_STANDALONE_MIRRORS
#include "force_gpu_b200.hpp"
#define PB_POS_OFFSET(T) offsetof(T, pos)
#define PB_QUAD_OFFSET(T) offsetof(T, qxx)
#else
#include <particle_simulator.hpp>
#include "force_gpu_cuda.hpp"       // PeTar's own declaration of the boundary
#define PB_POS_OFFSET(T) offsetof(T, pos)
#define PB_QUAD_OFFSET(T) offsetof(T, quad)
#endif

#include "petar_b200.h"

#ifdef GPU_PROFILE
GPUProfile gpu_profile;              // reference src/force_gpu_cuda.cu:7-10
GPUCounter gpu_counter;
#endif

namespace {

// PeTar's style for unrecoverable errors is message + abort (e.g. reference src/petar.hpp:515-517)
void check(int rc, const char* where) {
    if (rc != PB_OK) {
        std::fprintf(stderr, "petar_b200: %s failed (%d): %s\n", where, rc, pb_last_error());
        std::abort();
    }
}

// offsetof on PeTar's non-standard-layout classes is conditionally supported; g++ accepts it
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Winvalid-offsetof"
const pb_layout_epi kLayoutEpi = {sizeof(EPISoft), PB_POS_OFFSET(EPISoft), offsetof(EPISoft, r_search)};
const pb_layout_epj kLayoutEpj = {sizeof(EPJSoft), PB_POS_OFFSET(EPJSoft), offsetof(EPJSoft, mass), offsetof(EPJSoft, r_search)};
#ifdef USE_QUAD
const pb_layout_spj kLayoutSpj = {sizeof(SPJSoft), PB_POS_OFFSET(SPJSoft), offsetof(SPJSoft, mass), PB_QUAD_OFFSET(SPJSoft), 1};
#else
const pb_layout_spj kLayoutSpj = {sizeof(SPJSoft), PB_POS_OFFSET(SPJSoft), offsetof(SPJSoft, mass), 0, 0};
#endif
const pb_layout_force kLayoutForce = {sizeof(ForceSoft), offsetof(ForceSoft, acc), offsetof(ForceSoft, pot), offsetof(ForceSoft, n_ngb)};
#pragma GCC diagnostic pop

// device = my_rank % device count, as reference src/force_gpu_cuda.cu:550-553.  pb_init is idempotent and cheap once
// the library is initialised, so it is simply called every time: after a pb_finalize() (a PeTar that re-initialises)
// the rank's own device is selected again instead of the library's lazy default (device 0).
void first_call(PS::S32 my_rank) {
    check(pb_init(my_rank, -1), "pb_init");
}

#ifdef GPU_PROFILE
// fold the library's timers/counters into PeTar's globals (reference :610-617, 672-697, 852-861)
void harvest_profile() {
    // the library's profile is cumulative; add what accrued since the last harvest (PeTar clears
    // gpu_profile / gpu_counter itself, reference src/petar.hpp:2081-2084)
    static pb_profile last = {};
    pb_profile p;
    pb_get_profile(&p, 0);
    if (p.n_call < last.n_call) last = pb_profile{};      // somebody reset the library's profile
    gpu_profile.copy.time += p.t_copy - last.t_copy;
    gpu_profile.send.time += p.t_send - last.t_send;
    gpu_profile.recv.time += p.t_recv - last.t_recv;
    gpu_profile.calc.time += p.t_calc - last.t_calc;
    gpu_counter.n_walk += p.n_walk - last.n_walk;
    gpu_counter.n_epi  += p.n_epi - last.n_epi;
    gpu_counter.n_epj  += p.n_epj - last.n_epj;
    gpu_counter.n_spj  += p.n_spj - last.n_spj;
    gpu_counter.n_call += p.n_call - last.n_call;
    last = p;
}
#endif

} // namespace

#ifdef PARTICLE_SIMULATOR_GPU_MULIT_WALK_INDEX

// replaces reference src/force_gpu_cuda.cu:535-700
PS::S32 CalcForceWithLinearCutoffCUDAMultiWalk::operator()(const PS::S32 tag,
                                                           const PS::S32 n_walk,
                                                           const EPISoft** epi,
                                                           const PS::S32* n_epi,
                                                           const PS::S32** id_epj,
                                                           const PS::S32* n_epj,
                                                           const PS::S32** id_spj,
                                                           const PS::S32* n_spj,
                                                           const EPJSoft* epj,
                                                           const PS::S32 n_epj_tot,
                                                           const SPJSoft* spj,
                                                           const PS::S32 n_spj_tot,
                                                           const bool send_flag) {
    (void)tag;                                   // FDPS passes 0, simd_test.cxx passes 1: ignored
    first_call(my_rank);
    check(pb_set_params(eps2, rcut2, G), "pb_set_params");
    if (send_flag) {
        check(pb_upload_j(epj, n_epj_tot, &kLayoutEpj, spj, n_spj_tot, &kLayoutSpj), "pb_upload_j");
    } else {
        check(pb_dispatch_index(n_walk, (const void* const*)epi, n_epi, &kLayoutEpi,
                                (const int* const*)id_epj, n_epj, (const int* const*)id_spj, n_spj),
              "pb_dispatch_index");
    }
    return 0;
}

#else

// replaces reference src/force_gpu_cuda.cu:704-827
PS::S32 CalcForceWithLinearCutoffCUDA::operator()(const PS::S32 tag,
                                                  const PS::S32 n_walk,
                                                  const EPISoft* epi[],
                                                  const PS::S32 n_epi[],
                                                  const EPJSoft* epj[],
                                                  const PS::S32 n_epj[],
                                                  const SPJSoft* spj[],
                                                  const PS::S32 n_spj[]) {
    (void)tag;
    first_call(my_rank);
    check(pb_set_params(eps2, rcut2, G), "pb_set_params");
    check(pb_dispatch_direct(n_walk, (const void* const*)epi, n_epi, &kLayoutEpi,
                             (const void* const*)epj, n_epj, &kLayoutEpj,
                             (const void* const*)spj, n_spj, &kLayoutSpj),
          "pb_dispatch_direct");
    return 0;
}

#endif

#ifdef PARTICLE_SIMULATOR_GPU_MULIT_WALK_INDEX
// EXTENSION (not declared by PeTar's force_gpu_cuda.hpp): a multiwalk dispatch
// functor for the neighbour-search tree.  PeTar would use it by replacing, in treeNeighborSearch
// (reference src/petar.hpp:774-778),
//     tree_nb.calcForceAllAndWriteBack(SearchNeighborEpEpSimd(), system_soft, dinfo);
// with
//     tree_nb.calcForceAllAndWriteBackMultiWalkIndex(SearchNeighborCUDAMultiWalk(my_rank),
//                                                    RetrieveForceCUDA, 1, system_soft, dinfo, 200);
// Same argument list as the force functor; superparticle arguments are ignored.
struct SearchNeighborCUDAMultiWalk {
    PS::S32 my_rank;
    explicit SearchNeighborCUDAMultiWalk(PS::S32 r = 0) : my_rank(r) {}
    PS::S32 operator()(const PS::S32 tag, const PS::S32 n_walk, const EPISoft** epi, const PS::S32* n_epi,
                       const PS::S32** id_epj, const PS::S32* n_epj, const PS::S32** id_spj, const PS::S32* n_spj,
                       const EPJSoft* epj, const PS::S32 n_epj_tot, const SPJSoft* spj, const PS::S32 n_spj_tot,
                       const bool send_flag) {
        (void)tag; (void)id_spj; (void)n_spj; (void)spj; (void)n_spj_tot;
        first_call(my_rank);
        if (send_flag) check(pb_upload_j(epj, n_epj_tot, &kLayoutEpj, nullptr, 0, nullptr), "pb_upload_j");
        else check(pb_dispatch_count_index(n_walk, (const void* const*)epi, n_epi, &kLayoutEpi, (const int* const*)id_epj, n_epj),
                   "pb_dispatch_count_index");
        return 0;
    }
};
#endif

// replaces reference src/force_gpu_cuda.cu:831-880
PS::S32 RetrieveForceCUDA(const PS::S32 tag, const PS::S32 n_walk, const PS::S32* ni, ForceSoft** force) {
    (void)tag;
    check(pb_retrieve(n_walk, ni, (void* const*)force, &kLayoutForce), "pb_retrieve");
#ifdef GPU_PROFILE
    harvest_profile();
#endif
    return 0;
}

#ifdef PB_STANDALONE_MIRRORS
// ---- test hooks: drive the functors exactly as FDPS / src/simd_test.cxx:139-155 do, from C ----
extern "C" {

#ifdef PARTICLE_SIMULATOR_GPU_MULIT_WALK_INDEX
int pb_shim_dispatch(int my_rank, double eps2, double rcut2, double G, int tag, int n_walk,
                     const void** epi, const int* n_epi, const int** id_epj, const int* n_epj,
                     const int** id_spj, const int* n_spj,
                     const void* epj, int n_epj_tot, const void* spj, int n_spj_tot, int send_flag) {
    CalcForceWithLinearCutoffCUDAMultiWalk f(my_rank, eps2, rcut2, G);     // a temporary per call, as PeTar does
    return f(tag, n_walk, (const EPISoft**)epi, n_epi, id_epj, n_epj, id_spj, n_spj,
             (const EPJSoft*)epj, n_epj_tot, (const SPJSoft*)spj, n_spj_tot, send_flag != 0);
}
#else
int pb_shim_dispatch_direct(int my_rank, double eps2, double rcut2, double G, int tag, int n_walk,
                            const void** epi, const int* n_epi, const void** epj, const int* n_epj,
                            const void** spj, const int* n_spj) {
    CalcForceWithLinearCutoffCUDA f(my_rank, eps2, rcut2, G);
    return f(tag, n_walk, (const EPISoft**)epi, n_epi, (const EPJSoft**)epj, n_epj, (const SPJSoft**)spj, n_spj);
}
#endif

#ifdef PARTICLE_SIMULATOR_GPU_MULIT_WALK_INDEX
int pb_shim_dispatch_count(int my_rank, int tag, int n_walk, const void** epi, const int* n_epi,
                           const int** id_epj, const int* n_epj, const void* epj, int n_epj_tot, int send_flag) {
    SearchNeighborCUDAMultiWalk f(my_rank);
    return f(tag, n_walk, (const EPISoft**)epi, n_epi, id_epj, n_epj, nullptr, nullptr,
             (const EPJSoft*)epj, n_epj_tot, nullptr, 0, send_flag != 0);
}
#endif

int pb_shim_retrieve(int tag, int n_walk, const int* ni, void** force) {
    return RetrieveForceCUDA(tag, n_walk, ni, (ForceSoft**)force);
}

#ifdef PARTICLE_SIMULATOR_GPU_MULIT_WALK_INDEX
// The loop FDPS calcForceAllAndWriteBackMultiWalkIndex(dispatch, retrieve, tag_max = 1, ...) runs around the
// two functors (call site src/petar.hpp:894-899), over walk groups whose lists already exist: one dispatch
// with send_flag = true publishing all j, then per group retrieve(previous) + dispatch(this), and a last
// retrieve.  Group g's tables are the arrays FDPS holds when it calls dispatch: epi[g][iw], n_epi[g][iw], ...
int pb_shim_walk_group_loop(int my_rank, double eps2, double rcut2, double G, int n_group, const int* n_walk,
                            const void* const* epi, const void* const* n_epi,
                            const void* const* id_epj, const void* const* n_epj,
                            const void* const* id_spj, const void* const* n_spj, const void* const* force,
                            const void* epj, int n_epj_tot, const void* spj, int n_spj_tot, int send_flag) {
    int ret = 0;
    if (send_flag) {
        CalcForceWithLinearCutoffCUDAMultiWalk f(my_rank, eps2, rcut2, G);
        ret += f(0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                 (const EPJSoft*)epj, n_epj_tot, (const SPJSoft*)spj, n_spj_tot, true);
    }
    for (int g = 0; g < n_group; g++) {
        if (g > 0) ret += RetrieveForceCUDA(0, n_walk[g - 1], (const int*)n_epi[g - 1], (ForceSoft**)force[g - 1]);
        CalcForceWithLinearCutoffCUDAMultiWalk f(my_rank, eps2, rcut2, G);     // a temporary per call, as PeTar does
        ret += f(0, n_walk[g], (const EPISoft**)epi[g], (const int*)n_epi[g], (const int**)id_epj[g], (const int*)n_epj[g],
                 (const int**)id_spj[g], (const int*)n_spj[g],
                 (const EPJSoft*)epj, n_epj_tot, (const SPJSoft*)spj, n_spj_tot, false);
    }
    if (n_group > 0) ret += RetrieveForceCUDA(0, n_walk[n_group - 1], (const int*)n_epi[n_group - 1], (ForceSoft**)force[n_group - 1]);
    return ret;
}
#endif

#ifdef GPU_PROFILE
void pb_shim_profile(double t[4], long long n[5], int clear) {
    t[0] = gpu_profile.copy.time; t[1] = gpu_profile.send.time; t[2] = gpu_profile.recv.time; t[3] = gpu_profile.calc.time;
    n[0] = gpu_counter.n_walk.n; n[1] = gpu_counter.n_epi.n; n[2] = gpu_counter.n_epj.n; n[3] = gpu_counter.n_spj.n; n[4] = gpu_counter.n_call.n;
    if (clear) { gpu_profile.clear(); gpu_counter.clear(); }
}
#endif

} // extern "C"
#endif
