/* petar_b200_types.h — POD mirrors of the PeTar / FDPS particle layouts that cross the
 * soft-force dispatch/retrieve boundary.
 *
 * These mirror, field for field, the layouts the reference hands to its GPU functor
 * (default build: no KDKDK_4TH, no SAVE_NEIGHBOR_ID_IN_FORCE_KERNEL, USE_QUAD):
 *
 *   pb_EPISoft    <-> class EPISoft   reference src/soft_ptcl.hpp:271-308   (48 B)
 *   pb_EPJSoft    <-> class EPJSoft   reference src/soft_ptcl.hpp:311-376   (120 B)
 *   pb_SPJQuad    <-> PS::SPJQuadrupoleInAndOut (FDPS; used at src/force_gpu_cuda.cu:597-607) (80 B)
 *   pb_SPJMono    <-> PS::SPJMonopoleInAndOut   (FDPS)                      (32 B)
 *   pb_ForceSoft  <-> class ForceSoft reference src/soft_ptcl.hpp:4-24      (40 B)
 *
 * The C ABI (petar_b200.h) never depends on these structs: it takes base pointer + stride +
 * field offsets (pb_layout_*), so a build of PeTar with a different struct layout (e.g.
 * KDKDK_4TH adds `acc` to EPISoft/EPJSoft) binds without recompiling the library.
 * The mirrors exist for the in-repo C++ shim, the harness, the oracle and the tests.
 */
#ifndef PETAR_B200_TYPES_H
#define PETAR_B200_TYPES_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_f64vec { double x, y, z; } pb_f64vec;

/* reference src/soft_ptcl.hpp:271-278 */
typedef struct pb_EPISoft {
    int64_t   id;
    pb_f64vec pos;
    double    r_search;
    int32_t   rank_org;
    int32_t   type;      /* 0: orbital artificial particle; 1: others */
} pb_EPISoft;

/* reference src/soft_ptcl.hpp:311-326 (GroupDataDeliver is a 16-byte union, src/ptcl.hpp:17-30) */
typedef struct pb_EPJSoft {
    int64_t   id;
    double    mass;
    pb_f64vec pos;
    pb_f64vec vel;
    double    r_in;
    double    r_out;
    double    r_search;
    double    r_scale_next;
    int64_t   group_data[2];
    int32_t   rank_org;
    int32_t   adr_org;
} pb_EPJSoft;

/* FDPS PS::SPJQuadrupoleInAndOut: {F64 mass; F64vec pos; F64mat quad} with
 * PS::MatrixSym3 member order xx, yy, zz, xy, xz, yz (accessed by name at
 * reference src/force_gpu_cuda.cu:597-607 and src/soft_force.hpp:175-183). */
typedef struct pb_SPJQuad {
    double    mass;
    pb_f64vec pos;
    double    qxx, qyy, qzz, qxy, qxz, qyz;
} pb_SPJQuad;

typedef struct pb_SPJMono {
    double    mass;
    pb_f64vec pos;
} pb_SPJMono;

/* reference src/soft_ptcl.hpp:4-15 */
typedef struct pb_ForceSoft {
    pb_f64vec acc;
    double    pot;
    int64_t   n_ngb;
} pb_ForceSoft;

/* The fields of FPSoft / EPJSoft that the changeover correction reads (pos, mass, changeover radii, id,
 * group_data.artificial = {mass_backup, status}: reference src/ptcl.hpp, src/artificial_particles.hpp:20-28)
 * and updates (acc, pot_tot, pot_soft: src/soft_ptcl.hpp:27-60).  A stand-in layout for the oracle, the
 * harness and the tests; the ABI (pb_correct_changeover) takes field offsets, not this struct. */
typedef struct pb_PtclCorr {
    int64_t   id;
    double    mass;
    pb_f64vec pos;
    double    r_in, r_out;            /* ChangeOver::r_in_, r_out_ (FPSoft) or EPJSoft::r_in, r_out */
    double    mass_backup, status;    /* ArtificialParticleInformation */
    pb_f64vec acc;
    double    pot_tot, pot_soft;
} pb_PtclCorr;

#ifdef __cplusplus
}
static_assert(sizeof(pb_PtclCorr)  == 112, "PtclCorr must be 112 bytes");
static_assert(sizeof(pb_EPISoft)   == 48,  "EPISoft mirror must be 48 bytes");
static_assert(sizeof(pb_EPJSoft)   == 120, "EPJSoft mirror must be 120 bytes");
static_assert(sizeof(pb_SPJQuad)   == 80,  "SPJQuadrupoleInAndOut mirror must be 80 bytes");
static_assert(sizeof(pb_SPJMono)   == 32,  "SPJMonopoleInAndOut mirror must be 32 bytes");
static_assert(sizeof(pb_ForceSoft) == 40,  "ForceSoft mirror must be 40 bytes");
#endif

#endif /* PETAR_B200_TYPES_H */
