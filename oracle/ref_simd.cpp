// ref_simd.cpp — TEST INFRASTRUCTURE ONLY.
//
// Builds the reference's OWN x86 SIMD kernels (class PhantomGrapeQuad, reference
// src/phantomquad_for_p3t_x86.hpp) from the source file where it lies under /root/reference
// (include path given by oracle/Makefile; nothing is copied into this repository) and drives
// them exactly as the reference's SIMD functors do:
//
//   ref_epep_simd      <-> CalcForceEpEpWithLinearCutoffSimd   reference src/soft_force.hpp:350-425
//   ref_epsp_quad_simd <-> CalcForceEpSpQuadSimd               reference src/soft_force.hpp:491-554
//   ref_nb_simd        <-> SearchNeighborEpEpSimd              reference src/soft_force.hpp:239-283
//
// FDPS / SDAR headers are absent in this container, so soft_force.hpp itself cannot be included;
// the adapters below restate it over the POD mirrors of include/petar_b200_types.h.  The hot
// loops (the AVX2 / AVX-512 intrinsics) are the reference's, verbatim.
//
// Precision flags as the default PeTar build sets them: RSQRT_NR_EPJ_X2 defined, no SPJ Newton
// step (reference src/petar.hpp:13-14).  Output of this file goes ONLY to oracle/_ref/.
#include <iostream>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <omp.h>

namespace PS { typedef int S32; typedef double F64; }
#define RSQRT_NR_EPJ_X2
#include "phantomquad_for_p3t_x86.hpp"   // from /root/reference/src (see oracle/Makefile)

#include "petar_b200_types.h"

// `static thread_local PhantomGrapeQuad pg;` as reference src/soft_force.hpp:371 (≈9.5 MB/thread)
static PhantomGrapeQuad& pg_instance() {
    static thread_local PhantomGrapeQuad* pg = nullptr;
    if (!pg) {
        void* mem = nullptr;
        if (posix_memalign(&mem, 64, sizeof(PhantomGrapeQuad)) != 0) abort();
        pg = new (mem) PhantomGrapeQuad();
    }
    return *pg;
}

extern "C" {

const char* ref_simd_isa() {
#ifdef USE__AVX512
    return "avx512";
#else
    return "avx2";
#endif
}

int ref_nimax() { return PhantomGrapeQuad::NIMAX; }
int ref_njmax() { return PhantomGrapeQuad::NJMAX; }

// reference src/soft_force.hpp:350-425
void ref_epep_simd(const pb_EPISoft* ep_i, int n_ip, const pb_EPJSoft* ep_j, int n_jp,
                   pb_ForceSoft* force, double eps, double r_out, double G)
{
    const double eps2 = eps * eps;
    std::vector<int> ep_j_list(n_jp > 0 ? n_jp : 1), ep_i_list(n_ip > 0 ? n_ip : 1);
    int n_jp_local = 0, n_ip_local = 0;
    for (int i = 0; i < n_jp; i++)
        if (ep_j[i].mass > 0) ep_j_list[n_jp_local++] = i;                    // :360-362
    PhantomGrapeQuad& pg = pg_instance();
    pg.set_eps2(eps2);                                                        // :379
    pg.set_r_crit2(r_out * r_out);                                            // :380
    for (int i = 0; i < n_ip; i++) {
        if (ep_i[i].type == 1) {                                              // :383
            ep_i_list[n_ip_local] = i;
            pg.set_xi_one(n_ip_local, ep_i[i].pos.x, ep_i[i].pos.y, ep_i[i].pos.z, ep_i[i].r_search);
            n_ip_local++;
        }
    }
    const int loop_max = (n_jp_local - 1) / PhantomGrapeQuad::NJMAX + 1;      // :391
    for (int loop = 0; loop < loop_max; loop++) {
        const int ih = PhantomGrapeQuad::NJMAX * loop;
        const int n_jp_tmp = ((n_jp_local - ih) < PhantomGrapeQuad::NJMAX) ? (n_jp_local - ih) : PhantomGrapeQuad::NJMAX;
        const int it = ih + n_jp_tmp;
        int i_tmp = 0;
        for (int i = ih; i < it; i++, i_tmp++) {
            const int ij = ep_j_list[i];
            pg.set_epj_one(i_tmp, ep_j[ij].pos.x, ep_j[ij].pos.y, ep_j[ij].pos.z, ep_j[ij].mass, ep_j[ij].r_search);
        }
        pg.run_epj_for_p3t_with_linear_cutoff(n_ip, n_jp_tmp);                // :404 (n_ip, as written)
        for (int k = 0; k < n_ip_local; k++) {
            const int i = ep_i_list[k];
            double p = 0, a[3] = {0, 0, 0}, n_ngb = 0;
            pg.accum_accp_one(k, a[0], a[1], a[2], p, n_ngb);
            force[i].acc.x += G * a[0];
            force[i].acc.y += G * a[1];
            force[i].acc.z += G * a[2];
            force[i].pot   += G * p;
            force[i].n_ngb += (int)(n_ngb * 1.00001);                         // :421
        }
    }
}

// reference src/soft_force.hpp:491-554
void ref_epsp_quad_simd(const pb_EPISoft* ep_i, int n_ip, const pb_SPJQuad* sp_j, int n_jp,
                        pb_ForceSoft* force, double eps, double G)
{
    const double eps2 = eps * eps;
    std::vector<int> ep_i_list(n_ip > 0 ? n_ip : 1);
    int n_ip_local = 0;
    PhantomGrapeQuad& pg = pg_instance();
    pg.set_eps2(eps2);                                                        // :512
    for (int i = 0; i < n_ip; i++) {
        if (ep_i[i].type == 1) {                                              // :515
            ep_i_list[n_ip_local] = i;
            pg.set_xi_one(n_ip_local, ep_i[i].pos.x, ep_i[i].pos.y, ep_i[i].pos.z, 0.0);
            n_ip_local++;
        }
    }
    const int loop_max = (n_jp - 1) / PhantomGrapeQuad::NJMAX + 1;            // :522
    for (int loop = 0; loop < loop_max; loop++) {
        const int ih = PhantomGrapeQuad::NJMAX * loop;
        const int n_jp_tmp = ((n_jp - ih) < PhantomGrapeQuad::NJMAX) ? (n_jp - ih) : PhantomGrapeQuad::NJMAX;
        const int it = ih + n_jp_tmp;
        int i_tmp = 0;
        for (int i = ih; i < it; i++, i_tmp++) {
            // :532 the reference passes `i` (not i_tmp) as the buffer address; identical while
            // n_jp <= NJMAX, which FDPS walks always satisfy.  Use i_tmp to stay in bounds.
            pg.set_spj_one(i_tmp, sp_j[i].pos.x, sp_j[i].pos.y, sp_j[i].pos.z, sp_j[i].mass,
                           sp_j[i].qxx, sp_j[i].qyy, sp_j[i].qzz, sp_j[i].qxy, sp_j[i].qyz, sp_j[i].qxz);
        }
        pg.run_spj(n_ip, n_jp_tmp);                                           // :535
        for (int k = 0; k < n_ip_local; k++) {
            const int i = ep_i_list[k];
            double p = 0, a[3] = {0, 0, 0};
            pg.accum_accp_one(k, a[0], a[1], a[2], p);
            force[i].acc.x += G * a[0];
            force[i].acc.y += G * a[1];
            force[i].acc.z += G * a[2];
            force[i].pot   += G * p;
        }
    }
}

// reference src/soft_force.hpp:239-283
void ref_nb_simd(const pb_EPISoft* ep_i, int n_ip, const pb_EPJSoft* ep_j, int n_jp, pb_ForceSoft* force)
{
    PhantomGrapeQuad& pg = pg_instance();
    for (int i = 0; i < n_ip; i++)
        pg.set_xi_one(i, ep_i[i].pos.x, ep_i[i].pos.y, ep_i[i].pos.z, ep_i[i].r_search);
    const int loop_max = (n_jp - 1) / PhantomGrapeQuad::NJMAX + 1;
    for (int loop = 0; loop < loop_max; loop++) {
        const int ih = PhantomGrapeQuad::NJMAX * loop;
        const int n_jp_tmp = ((n_jp - ih) < PhantomGrapeQuad::NJMAX) ? (n_jp - ih) : PhantomGrapeQuad::NJMAX;
        const int it = ih + n_jp_tmp;
        int i_tmp = 0;
        for (int i = ih; i < it; i++, i_tmp++)
            pg.set_epj_one(i_tmp, ep_j[i].pos.x, ep_j[i].pos.y, ep_j[i].pos.z, ep_j[i].mass, ep_j[i].r_search);
        pg.run_epj_for_neighbor_count(n_ip, n_jp_tmp);
        for (int i = 0; i < n_ip; i++) {
            double n_ngb = 0;
            pg.accum_accp_one(i, n_ngb);
            force[i].n_ngb += (int)(n_ngb * 1.00001);                         // :279
        }
    }
}

// The reference CPU path for one index-mode multiwalk batch: what FDPS
// calcForceAllAndWriteBack(fep, fsp, ...) does per i-group on the SIMD build
// (reference src/petar.hpp:923-930): gather j by index, clear force, EP-EP + EP-SP functors.
// OpenMP over walks with one PhantomGrapeQuad per thread, `n_threads` <= 0 means all cores.
// Returns wall-clock seconds spent inside the functor calls' parallel region (gather included,
// as FDPS's own list copy is part of its calc_force phase).
double ref_walks_index(int n_walk,
                       const pb_EPISoft* const* epi, const int* n_epi,
                       const int* const* id_epj, const int* n_epj,
                       const int* const* id_spj, const int* n_spj,
                       const pb_EPJSoft* epj, const pb_SPJQuad* spj,
                       pb_ForceSoft* const* force,
                       double eps, double r_out, double G, int n_threads)
{
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    const double t0 = omp_get_wtime();
#pragma omp parallel num_threads(n_threads)
    {
        std::vector<pb_EPJSoft> ej;
        std::vector<pb_SPJQuad> sj;
#pragma omp for schedule(dynamic)
        for (int iw = 0; iw < n_walk; iw++) {
            ej.resize(n_epj[iw]);
            sj.resize(n_spj[iw]);
            for (int j = 0; j < n_epj[iw]; j++) ej[j] = epj[id_epj[iw][j]];
            for (int j = 0; j < n_spj[iw]; j++) sj[j] = spj[id_spj[iw][j]];
            pb_ForceSoft* f = force[iw];
            for (int i = 0; i < n_epi[iw]; i++) { f[i].acc.x = f[i].acc.y = f[i].acc.z = 0.0; f[i].pot = 0.0; f[i].n_ngb = 0; }
            ref_epep_simd(epi[iw], n_epi[iw], ej.data(), n_epj[iw], f, eps, r_out, G);
            if (n_spj[iw] > 0) ref_epsp_quad_simd(epi[iw], n_epi[iw], sj.data(), n_spj[iw], f, eps, G);
        }
    }
    return omp_get_wtime() - t0;
}

} // extern "C"
