/* oracle_soft_force.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU fp64 restatement of the reference's soft-force functors (reference src/soft_force.hpp)
 * and of the synthetic-input recipes the reference's own test uses (src/simd_test.cxx,
 * src/particle_distribution_generator.hpp).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call into this library; the product
 * (petar_b200/) never links or loads it.
 *
 * Parity pinning: the reference holds NO golden vectors for this path (src/simd_test.cxx never
 * asserts and stores no outputs).  This restatement is pinned instead against the reference's
 * own AVX2/AVX-512 kernels compiled verbatim from /root/reference (oracle/_ref, see
 * oracle/Makefile and tests/test_oracle.py) and against analytic known answers.
 */
#ifndef ORACLE_SOFT_FORCE_H
#define ORACLE_SOFT_FORCE_H

#include "petar_b200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- the four NoSimd functors (fp64) -------------------------------------------------- */

/* reference src/soft_force.hpp:11-34  SearchNeighborEpEpNoSimd: n_ngb ASSIGNED */
void orc_search_neighbor_epep(const pb_EPISoft* ep_i, int n_ip,
                              const pb_EPJSoft* ep_j, int n_jp, pb_ForceSoft* force);

/* reference src/soft_force.hpp:38-87  CalcForceEpEpWithLinearCutoffNoSimd:
 * acc/pot ACCUMULATED (+=), n_ngb ASSIGNED.  eps, r_out, G are the statics
 * EPISoft::eps, EPISoft::r_out, ForceSoft::grav_const (src/static_variables.hpp:9-11). */
void orc_force_epep_linear_cutoff(const pb_EPISoft* ep_i, int n_ip,
                                  const pb_EPJSoft* ep_j, int n_jp, pb_ForceSoft* force,
                                  double eps, double r_out, double G);

/* reference src/soft_force.hpp:125-158  CalcForceEpSpMonoNoSimd (mass/pos taken from the quad struct) */
void orc_force_epsp_mono(const pb_EPISoft* ep_i, int n_ip,
                         const pb_SPJQuad* sp_j, int n_jp, pb_ForceSoft* force,
                         double eps, double G);

/* reference src/soft_force.hpp:160-200  CalcForceEpSpQuadNoSimd */
void orc_force_epsp_quad(const pb_EPISoft* ep_i, int n_ip,
                         const pb_SPJQuad* sp_j, int n_jp, pb_ForceSoft* force,
                         double eps, double G);

/* reference src/soft_force.hpp:202-236  CalcForcePPNoSimd (eps2 = 0, no cutoff, no count) */
void orc_force_pp(const pb_EPISoft* ep_i, int n_ip,
                  const pb_EPJSoft* ep_j, int n_jp, pb_ForceSoft* force, double G);

/* ---- what FDPS + the functors do for an index-mode multiwalk batch -------------------- */

/* For every walk iw: gather ep_j by id_epj[iw][], sp_j by id_spj[iw][] (indices into the
 * shared sorted arrays, call protocol of reference src/petar.hpp:894-899), clear force,
 * apply the EP-EP and EP-SP(quad) NoSimd functors.  Result semantics equal what
 * RetrieveForceCUDA hands back (reference src/force_gpu_cuda.cu:865-875): acc/pot/n_ngb
 * ASSIGNED, EP and SP contributions summed, G applied.  OpenMP over walks. */
void orc_walks_index(int n_walk,
                     const pb_EPISoft* const* epi, const int* n_epi,
                     const int* const* id_epj, const int* n_epj,
                     const int* const* id_spj, const int* n_spj,
                     const pb_EPJSoft* epj, const pb_SPJQuad* spj,
                     pb_ForceSoft* const* force,
                     double eps, double r_out, double G);

/* ---- synthetic inputs ----------------------------------------------------------------- */

/* MT19937 as used by FDPS PS::MTTS (Matsumoto & Nishimura mt19937ar): exposed for tests */
typedef struct orc_mt19937 { uint32_t mt[624]; int mti; } orc_mt19937;
void     orc_mt_init(orc_mt19937* s, uint32_t seed);
uint32_t orc_mt_int32(orc_mt19937* s);
double   orc_mt_res53(orc_mt19937* s);
double   orc_mt_real2(orc_mt19937* s);

/* reference src/particle_distribution_generator.hpp:173-250 makePlummerModel.
 * `rank_seed` is what PS::Comm::getRank() returns (the function's own `seed` argument is unused
 * in the reference).  mass[n_loc], pos[3*n_loc], vel[3*n_loc] are caller-allocated. */
void orc_make_plummer(double mass_glb, long long n_glb, long long n_loc,
                      double* mass, double* pos, double* vel,
                      double eng, uint32_t rank_seed);

/* reference src/ptcl.hpp:227-238 Ptcl::calcRSearch */
double orc_calc_rsearch(const double vel[3], double dt_tree, double search_factor,
                        double r_out_i, double r_search_min);

/* reference src/changeover.hpp:44-59 ChangeOver::setR(m_fac, r_in, r_out): returns r_out_ */
double orc_changeover_rout(double m_fac, double r_in, double r_out, double* r_in_scaled);

/* reference src/simd_test.cxx:44-57 setSpj, USE_QUAD, driven by glibc rand();
 * call orc_srand(1) first to reproduce the default-seeded sequence of the test. */
void orc_srand(unsigned seed);
void orc_simdtest_set_spj(double N, pb_SPJQuad* sp);

/* The complete input set of reference src/simd_test.cxx:59-134 (Nepi=1000, Nepj=2000,
 * Nspj=1000, r_out=0.01, eps=1e-4, G=1, dt_tree=1/2048, statics search_factor=r_search_min=0).
 * Arrays are caller-allocated with n_epi / n_epj / n_spj entries. */
void orc_simdtest_inputs(int n_epi, int n_epj, int n_spj,
                         pb_EPISoft* epi, pb_EPJSoft* epj, pb_SPJQuad* spj);

/* ---- changeover correction of the soft force (oracle_changeover.c) --------------------------- */
/* reference src/changeover.hpp:69-78 (setR), :318-334 (calcAcc0W), :294-309 (calcPotW) */
void orc_changeover_w(double r_in, double r_out, double dr, double* acc0w, double* potw);
/* reference src/hard.hpp:1408-1476 calcAccPotShortWithLinearCutoff(Tpi&, const EPJSoft&); replay_fp32 selects
 * the `USE_GPU` branch (float replay of the linear-cutoff term) or the all-double `#else` branch */
void orc_changeover_pair(pb_PtclCorr* pi, const pb_PtclCorr* pj, double eps2, double r_out, double G, int replay_fp32);
/* reference src/hard.hpp:1655-1691 for every particle (loop of :3366-3377); neighbours in CSR form,
 * indices into pj; status_no_cm = -PS::LARGE_FLOAT (a member without c.m. particle) */
void orc_correct_force_tree_neighbor(pb_PtclCorr* p, int n, const int* nb_off, const int* nb_idx, const pb_PtclCorr* pj,
                                     double eps2, double r_out, double G, double status_no_cm, int replay_fp32);

#ifdef __cplusplus
}
#endif
#endif
