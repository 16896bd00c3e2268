/* oracle_soft_force.c — TEST INFRASTRUCTURE ONLY (see oracle_soft_force.h).
 *
 * fp64 restatement of the reference's NoSimd soft-force functors.  Arithmetic order follows the
 * reference source line by line (cited per function) so results agree with a real PeTar NoSimd
 * build to the last bit on the same compiler flags (-O2, no -ffast-math, no FMA contraction:
 * built with -ffp-contract=off).
 */
#include "oracle_soft_force.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * reference src/soft_force.hpp:11-34
 * ---------------------------------------------------------------------------------------- */
void orc_search_neighbor_epep(const pb_EPISoft* ep_i, int n_ip,
                              const pb_EPJSoft* ep_j, int n_jp, pb_ForceSoft* force)
{
    for (int i = 0; i < n_ip; i++) {
        const pb_f64vec xi = ep_i[i].pos;
        int n_ngb_i = 0;
        for (int j = 0; j < n_jp; j++) {
            const double rx = xi.x - ep_j[j].pos.x;
            const double ry = xi.y - ep_j[j].pos.y;
            const double rz = xi.z - ep_j[j].pos.z;
            const double r2 = rx * rx + ry * ry + rz * rz;   /* PS::F64vec operator*: x*x+y*y+z*z */
            const double r_search = fmax(ep_i[i].r_search, ep_j[j].r_search);
            if (r2 < r_search * r_search) n_ngb_i++;
        }
        force[i].n_ngb = n_ngb_i;
    }
}

/* ------------------------------------------------------------------------------------------
 * reference src/soft_force.hpp:38-87
 * ---------------------------------------------------------------------------------------- */
void orc_force_epep_linear_cutoff(const pb_EPISoft* ep_i, int n_ip,
                                  const pb_EPJSoft* ep_j, int n_jp, pb_ForceSoft* force,
                                  double eps, double r_out, double G)
{
    const double eps2   = eps * eps;          /* :44 */
    const double r_out2 = r_out * r_out;      /* :45 */
    for (int i = 0; i < n_ip; i++) {
        const pb_f64vec xi = ep_i[i].pos;
        double ax = 0.0, ay = 0.0, az = 0.0, poti = 0.0;
        int n_ngb_i = 0;
        for (int j = 0; j < n_jp; j++) {
            const double rx = xi.x - ep_j[j].pos.x;                        /* :58 */
            const double ry = xi.y - ep_j[j].pos.y;
            const double rz = xi.z - ep_j[j].pos.z;
            const double r2     = rx * rx + ry * ry + rz * rz;             /* :59 */
            const double r2_eps = r2 + eps2;                               /* :60 */
            const double r_search = fmax(ep_i[i].r_search, ep_j[j].r_search); /* :61 */
            if (r2 < r_search * r_search) n_ngb_i++;                       /* :62-64 (r2 WITHOUT eps) */
            const double r2_tmp = (r2_eps > r_out2) ? r2_eps : r_out2;     /* :65 */
            const double r_inv  = 1.0 / sqrt(r2_tmp);                      /* :66 */
            const double m_r    = ep_j[j].mass * r_inv;                    /* :67 */
            const double m_r3   = m_r * r_inv * r_inv;                     /* :68 */
            ax -= m_r3 * rx;                                               /* :69 */
            ay -= m_r3 * ry;
            az -= m_r3 * rz;
            poti -= m_r;                                                   /* :70 */
        }
        force[i].acc.x += G * ax;                                          /* :73 */
        force[i].acc.y += G * ay;
        force[i].acc.z += G * az;
        force[i].pot   += G * poti;                                        /* :77 */
        force[i].n_ngb  = n_ngb_i;                                         /* :84 */
    }
}

/* ------------------------------------------------------------------------------------------
 * reference src/soft_force.hpp:125-158
 * ---------------------------------------------------------------------------------------- */
void orc_force_epsp_mono(const pb_EPISoft* ep_i, int n_ip,
                         const pb_SPJQuad* sp_j, int n_jp, pb_ForceSoft* force,
                         double eps, double G)
{
    const double eps2 = eps * eps;
    for (int i = 0; i < n_ip; i++) {
        const pb_f64vec xi = ep_i[i].pos;
        double ax = 0.0, ay = 0.0, az = 0.0, poti = 0.0;
        for (int j = 0; j < n_jp; j++) {
            const double rx = xi.x - sp_j[j].pos.x;
            const double ry = xi.y - sp_j[j].pos.y;
            const double rz = xi.z - sp_j[j].pos.z;
            double r3_inv = rx * rx + ry * ry + rz * rz + eps2;   /* :140 */
            double r_inv  = 1.0 / sqrt(r3_inv);                   /* :141 */
            r3_inv  = r_inv * r_inv;                              /* :142 */
            r_inv  *= sp_j[j].mass;                               /* :143 */
            r3_inv *= r_inv;                                      /* :144 */
            ax -= r3_inv * rx;                                    /* :145 */
            ay -= r3_inv * ry;
            az -= r3_inv * rz;
            poti -= r_inv;                                        /* :146 */
        }
        force[i].acc.x += G * ax;
        force[i].acc.y += G * ay;
        force[i].acc.z += G * az;
        force[i].pot   += G * poti;
    }
}

/* ------------------------------------------------------------------------------------------
 * reference src/soft_force.hpp:160-200
 * ---------------------------------------------------------------------------------------- */
void orc_force_epsp_quad(const pb_EPISoft* ep_i, int n_ip,
                         const pb_SPJQuad* sp_j, int n_jp, pb_ForceSoft* force,
                         double eps, double G)
{
    const double eps2 = eps * eps;
    for (int ip = 0; ip < n_ip; ip++) {
        const pb_f64vec xi = ep_i[ip].pos;
        double ax = 0.0, ay = 0.0, az = 0.0, poti = 0.0;
        for (int jp = 0; jp < n_jp; jp++) {
            const double mj = sp_j[jp].mass;                                   /* :175 */
            const double rx = xi.x - sp_j[jp].pos.x;                           /* :177 */
            const double ry = xi.y - sp_j[jp].pos.y;
            const double rz = xi.z - sp_j[jp].pos.z;
            const double r2 = rx * rx + ry * ry + rz * rz + eps2;              /* :178 */
            const double qxx = sp_j[jp].qxx, qyy = sp_j[jp].qyy, qzz = sp_j[jp].qzz;
            const double qxy = sp_j[jp].qxy, qxz = sp_j[jp].qxz, qyz = sp_j[jp].qyz;
            const double tr  = qxx + qyy + qzz;                                /* :180 getTrace */
            const double qrx = qxx * rx + qxy * ry + qxz * rz;                 /* :181 */
            const double qry = qyy * ry + qyz * rz + qxy * rx;                 /* :182 */
            const double qrz = qzz * rz + qxz * rx + qyz * ry;                 /* :183 */
            const double qrr = qrx * rx + qry * ry + qrz * rz;                 /* :184 */
            const double r_inv  = 1.0f / sqrt(r2);                             /* :185 (1.0f literal) */
            const double r2_inv = r_inv * r_inv;                               /* :186 */
            const double r3_inv = r2_inv * r_inv;                              /* :187 */
            const double r5_inv = r2_inv * r3_inv * 1.5;                       /* :188 */
            const double qrr_r5 = r5_inv * qrr;                                /* :189 */
            const double qrr_r7 = r2_inv * qrr_r5;                             /* :190 */
            const double A = mj * r3_inv - tr * r5_inv + 5 * qrr_r7;           /* :191 */
            const double B = -2.0 * r5_inv;                                    /* :192 */
            ax -= A * rx + B * qrx;                                            /* :193 */
            ay -= A * ry + B * qry;
            az -= A * rz + B * qrz;
            poti -= mj * r_inv - 0.5 * tr * r3_inv + qrr_r5;                   /* :194 */
        }
        force[ip].acc.x += G * ax;                                             /* :196 */
        force[ip].acc.y += G * ay;
        force[ip].acc.z += G * az;
        force[ip].pot   += G * poti;                                           /* :197 */
    }
}

/* ------------------------------------------------------------------------------------------
 * reference src/soft_force.hpp:202-236
 * ---------------------------------------------------------------------------------------- */
void orc_force_pp(const pb_EPISoft* ep_i, int n_ip,
                  const pb_EPJSoft* ep_j, int n_jp, pb_ForceSoft* force, double G)
{
    const double eps2 = 0;
    for (int i = 0; i < n_ip; i++) {
        const pb_f64vec xi = ep_i[i].pos;
        double ax = 0.0, ay = 0.0, az = 0.0, poti = 0.0;
        for (int j = 0; j < n_jp; j++) {
            const double rx = xi.x - ep_j[j].pos.x;
            const double ry = xi.y - ep_j[j].pos.y;
            const double rz = xi.z - ep_j[j].pos.z;
            double r3_inv = rx * rx + ry * ry + rz * rz + eps2;
            double r_inv  = 1.0 / sqrt(r3_inv);
            r3_inv  = r_inv * r_inv;
            r_inv  *= ep_j[j].mass;
            r3_inv *= r_inv;
            ax -= r3_inv * rx;
            ay -= r3_inv * ry;
            az -= r3_inv * rz;
            poti -= r_inv;
        }
        force[i].acc.x += G * ax;
        force[i].acc.y += G * ay;
        force[i].acc.z += G * az;
        force[i].pot   += G * poti;
    }
}

/* ------------------------------------------------------------------------------------------
 * Index-mode multiwalk batch: what FDPS does around the functors
 * (call protocol reference src/petar.hpp:894-899; retrieve semantics src/force_gpu_cuda.cu:865-875)
 * ---------------------------------------------------------------------------------------- */
void orc_walks_index(int n_walk,
                     const pb_EPISoft* const* epi, const int* n_epi,
                     const int* const* id_epj, const int* n_epj,
                     const int* const* id_spj, const int* n_spj,
                     const pb_EPJSoft* epj, const pb_SPJQuad* spj,
                     pb_ForceSoft* const* force,
                     double eps, double r_out, double G)
{
#pragma omp parallel
    {
        pb_EPJSoft* ej = NULL; size_t cap_e = 0;
        pb_SPJQuad* sj = NULL; size_t cap_s = 0;
#pragma omp for schedule(dynamic)
        for (int iw = 0; iw < n_walk; iw++) {
            if ((size_t)n_epj[iw] > cap_e) { cap_e = (size_t)n_epj[iw] * 2; free(ej); ej = (pb_EPJSoft*)malloc(cap_e * sizeof(pb_EPJSoft)); }
            if ((size_t)n_spj[iw] > cap_s) { cap_s = (size_t)n_spj[iw] * 2; free(sj); sj = (pb_SPJQuad*)malloc(cap_s * sizeof(pb_SPJQuad)); }
            for (int j = 0; j < n_epj[iw]; j++) ej[j] = epj[id_epj[iw][j]];
            for (int j = 0; j < n_spj[iw]; j++) sj[j] = spj[id_spj[iw][j]];
            pb_ForceSoft* f = force[iw];
            for (int i = 0; i < n_epi[iw]; i++) {           /* ForceSoft::clear, src/soft_ptcl.hpp:16-23 */
                f[i].acc.x = f[i].acc.y = f[i].acc.z = 0.0; f[i].pot = 0.0; f[i].n_ngb = 0;
            }
            orc_force_epep_linear_cutoff(epi[iw], n_epi[iw], ej, n_epj[iw], f, eps, r_out, G);
            orc_force_epsp_quad(epi[iw], n_epi[iw], sj, n_spj[iw], f, eps, G);
        }
        free(ej); free(sj);
    }
}

/* ------------------------------------------------------------------------------------------
 * MT19937 (Matsumoto & Nishimura 1998, mt19937ar reference algorithm) — FDPS PS::MTTS wraps it.
 * ---------------------------------------------------------------------------------------- */
void orc_mt_init(orc_mt19937* s, uint32_t seed)
{
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->mti = 624;
}

uint32_t orc_mt_int32(orc_mt19937* s)
{
    if (s->mti >= 624) {
        for (int k = 0; k < 624; k++) {
            uint32_t y = (s->mt[k] & 0x80000000u) | (s->mt[(k + 1) % 624] & 0x7fffffffu);
            uint32_t v = s->mt[(k + 397) % 624] ^ (y >> 1);
            if (y & 1u) v ^= 0x9908b0dfu;
            s->mt[k] = v;
        }
        s->mti = 0;
    }
    uint32_t y = s->mt[s->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

double orc_mt_res53(orc_mt19937* s)
{
    uint32_t a = orc_mt_int32(s) >> 5, b = orc_mt_int32(s) >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}

double orc_mt_real2(orc_mt19937* s)
{
    return orc_mt_int32(s) * (1.0 / 4294967296.0);
}

/* ------------------------------------------------------------------------------------------
 * reference src/particle_distribution_generator.hpp:173-250
 * ---------------------------------------------------------------------------------------- */
void orc_make_plummer(double mass_glb, long long n_glb, long long n_loc,
                      double* mass, double* pos, double* vel,
                      double eng, uint32_t rank_seed)
{
    const double PI = atan(1.0) * 4.0;                                                      /* :183 */
    const double r_cutoff = 22.8 / (-3.0 * PI * mass_glb * mass_glb / (64.0 * -0.25));      /* :184 */
    orc_mt19937 mt;
    orc_mt_init(&mt, rank_seed);                                                            /* :191 */
    for (long long i = 0; i < n_loc; i++) {
        mass[i] = mass_glb / n_glb;                                                         /* :193 */
        double r_tmp = 9999.9;
        while (r_tmp > r_cutoff) {                                                          /* :195-198 */
            double m_tmp = orc_mt_res53(&mt);
            r_tmp = 1.0 / sqrt(pow(m_tmp, (-2.0 / 3.0)) - 1.0);
        }
        double phi = 2.0 * PI * orc_mt_res53(&mt);                                          /* :199 */
        double cth = 2.0 * (orc_mt_real2(&mt) - 0.5);                                       /* :200 */
        double sth = sqrt(1.0 - cth * cth);
        pos[3 * i + 0] = r_tmp * sth * cos(phi);
        pos[3 * i + 1] = r_tmp * sth * sin(phi);
        pos[3 * i + 2] = r_tmp * cth;
        while (1) {                                                                         /* :205-219 */
            const double v_max  = 0.1;
            const double v_try  = orc_mt_res53(&mt);
            const double v_crit = v_max * orc_mt_res53(&mt);
            if (v_crit < v_try * v_try * pow((1.0 - v_try * v_try), 3.5)) {
                const double ve = sqrt(2.0) * pow((r_tmp * r_tmp + 1.0), -0.25);
                phi = 2.0 * PI * orc_mt_res53(&mt);
                cth = 2.0 * (orc_mt_res53(&mt) - 0.5);
                sth = sqrt(1.0 - cth * cth);
                vel[3 * i + 0] = ve * v_try * sth * cos(phi);
                vel[3 * i + 1] = ve * v_try * sth * sin(phi);
                vel[3 * i + 2] = ve * v_try * cth;
                break;
            }
        }
    }
    double cp[3] = {0, 0, 0}, cv[3] = {0, 0, 0}, cm = 0.0;                                  /* :222-233 */
    for (long long i = 0; i < n_loc; i++) {
        for (int k = 0; k < 3; k++) { cp[k] += mass[i] * pos[3 * i + k]; cv[k] += mass[i] * vel[3 * i + k]; }
        cm += mass[i];
    }
    for (int k = 0; k < 3; k++) { cp[k] /= cm; cv[k] /= cm; }
    for (long long i = 0; i < n_loc; i++)
        for (int k = 0; k < 3; k++) { pos[3 * i + k] -= cp[k]; vel[3 * i + k] -= cv[k]; }
    const double r_scale = -3.0 * PI * mass_glb * mass_glb / (64.0 * eng);                  /* :235 */
    const double coef = 1.0 / sqrt(r_scale);                                                /* :236 */
    for (long long i = 0; i < n_loc; i++)
        for (int k = 0; k < 3; k++) { pos[3 * i + k] *= r_scale; vel[3 * i + k] *= coef; }
}

/* reference src/ptcl.hpp:227-231 */
double orc_calc_rsearch(const double vel[3], double dt_tree, double search_factor,
                        double r_out_i, double r_search_min)
{
    const double v = sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
    return fmax(v * dt_tree * search_factor + r_out_i, r_search_min);
}

/* reference src/changeover.hpp:44-52 */
double orc_changeover_rout(double m_fac, double r_in, double r_out, double* r_in_scaled)
{
    const double m_fac3 = fmax(pow(m_fac, (1.0 / 3.0)), 1.0);
    if (r_in_scaled) *r_in_scaled = m_fac3 * r_in;
    return m_fac3 * r_out;
}

void orc_srand(unsigned seed) { srand(seed); }

/* reference src/simd_test.cxx:44-57 (USE_QUAD); mixed int/float/double arithmetic kept as written */
void orc_simdtest_set_spj(double N, pb_SPJQuad* sp)
{
    sp->mass  = 1.0 / N + 0.001 / N * rand() / (float)RAND_MAX;
    sp->pos.x = 1.0 + 10.0 * rand() / (float)RAND_MAX;
    sp->pos.y = 1.0 + 10.0 * rand() / (float)RAND_MAX;
    sp->pos.z = 1.0 + 10.0 * rand() / (float)RAND_MAX;
    sp->qxx = 10.0 * rand() / (float)RAND_MAX;
    sp->qyy = 10.0 * rand() / (float)RAND_MAX;
    sp->qzz = 10.0 * rand() / (float)RAND_MAX;
    sp->qxy = 10.0 * rand() / (float)RAND_MAX;
    sp->qyz = 10.0 * rand() / (float)RAND_MAX;
    sp->qxz = 10.0 * rand() / (float)RAND_MAX;
}

/* reference src/simd_test.cxx:59-134 */
void orc_simdtest_inputs(int n_epi, int n_epj, int n_spj,
                         pb_EPISoft* epi, pb_EPJSoft* epj, pb_SPJQuad* spj)
{
    const int N = n_epi > n_epj ? n_epi : n_epj;                        /* :64 */
    double* mass = (double*)malloc(sizeof(double) * N);
    double* pos  = (double*)malloc(sizeof(double) * 3 * N);
    double* vel  = (double*)malloc(sizeof(double) * 3 * N);
    orc_make_plummer(1.0, N, N, mass, pos, vel, -0.25, 0);              /* :76 (rank 0) */
    double r_in;
    const double r_out = orc_changeover_rout(1.0, 0.001, 0.01, &r_in);  /* :84 */
    for (int i = 0; i < N; i++) {
        /* :108 calcRSearch(1/2048) with Ptcl::search_factor = Ptcl::r_search_min = 0
         * (src/static_variables.hpp:3-4) */
        const double rs = orc_calc_rsearch(&vel[3 * i], 1.0 / 2048.0, 0.0, r_out, 0.0);
        if (i < n_epi) {                                                /* :110-111 EPISoft::copyFromFP */
            memset(&epi[i], 0, sizeof(pb_EPISoft));
            epi[i].id = i + 1;
            epi[i].pos.x = pos[3 * i]; epi[i].pos.y = pos[3 * i + 1]; epi[i].pos.z = pos[3 * i + 2];
            epi[i].r_search = rs;
            epi[i].rank_org = 0;
            epi[i].type = 1;
        }
        if (i < n_epj) {                                                /* :130-131 EPJSoft::copyFromFP */
            memset(&epj[i], 0, sizeof(pb_EPJSoft));
            epj[i].id = i + 1;
            epj[i].mass = mass[i];
            epj[i].pos.x = pos[3 * i]; epj[i].pos.y = pos[3 * i + 1]; epj[i].pos.z = pos[3 * i + 2];
            epj[i].vel.x = vel[3 * i]; epj[i].vel.y = vel[3 * i + 1]; epj[i].vel.z = vel[3 * i + 2];
            epj[i].r_in = r_in; epj[i].r_out = r_out; epj[i].r_search = rs;
            epj[i].r_scale_next = 1.0;
            epj[i].rank_org = 0; epj[i].adr_org = i;
        }
    }
    for (int i = 0; i < n_spj; i++) orc_simdtest_set_spj((double)N, &spj[i]);   /* :133 */
    free(mass); free(pos); free(vel);
}
