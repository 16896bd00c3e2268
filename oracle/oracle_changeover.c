/* oracle_changeover.c — CPU restatement (fp64, plain C) of PeTar's changeover correction of the soft
 * force.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may call it.
 *
 * Pinned against the reference: oracle/_ref/libpetar_ref_changeover.so compiles the reference's own
 * ChangeOver class (src/changeover.hpp) and its calcAccPotShortWithLinearCutoff (extracted from
 * src/hard.hpp at build time) — tests/test_oracle.py compares the two bit for bit.
 *
 * What it follows:
 *   orc_changeover_set / _acc0w / _potw      src/changeover.hpp:69-78, 318-334, 294-309 (default build,
 *                                            no INTEGRATED_CUTOFF_FUNCTION); ...WTwo :366-375
 *   orc_changeover_pair                      src/hard.hpp:1408-1476 (the EPJSoft overload; the Ptcl overload
 *                                            :1329-1401 differs only in where r_in/r_out are read from)
 *   orc_correct_force_tree_neighbor          src/hard.hpp:1655-1691 (one particle) inside the particle loop of
 *                                            correctForceWithCutoffTreeNeighborOMP, src/hard.hpp:3366-3377
 *
 * replay_fp32 = 1 is the reference's `(!P3T_64BIT && USE_SIMD) || USE_GPU` branch: the linear-cutoff term is
 * re-evaluated in float from absolute coordinates (`1.0/sqrt(float)` resolves to the double sqrt under
 * <cmath>, then rounds to float — the way oracle/_ref compiles it); replay_fp32 = 0 is the `#else` branch,
 * everything in double.
 */
#include <math.h>
#include "oracle_soft_force.h"

typedef struct { double r_in, r_out, norm, coff, pot_off; } orc_changeover;

static orc_changeover co_set(double r_in, double r_out) {
    orc_changeover c;
    c.r_in = r_in; c.r_out = r_out;
    c.norm = 1.0 / (r_out - r_in);
    c.coff = (r_out - r_in) / (r_out + r_in);
    c.pot_off = (1.0 + c.coff) / r_out;
    return c;
}

static double co_potw(const orc_changeover* c, double dr) {
    double x = (dr - c->r_in) * c->norm;
    double k = 1.0;
    if (x >= 1.0) k = c->pot_off * dr;
    else if (x > 0.0) {
        double x2 = x * x, x3 = x2 * x, x5 = x2 * x3;
        k -= c->coff * x5 * (5.0 * x3 - 20.0 * x2 + 28.0 * x - 14.0);
    }
    return k;
}

static double co_acc0w(const orc_changeover* c, double dr) {
    double x = (dr - c->r_in) * c->norm;
    x = (x < 1.0) ? x : 1.0;
    x = (x > 0.0) ? x : 0.0;
    double x_1 = x - 1, x_2 = x_1 * x_1, x_4 = x_2 * x_2;
    double x2 = x * x, x3 = x2 * x, x4 = x2 * x2;
    return x_4 * (1.0 + 4.0 * x + 10.0 * x2 + 20.0 * x3 + 35.0 * c->coff * x4);
}

void orc_changeover_w(double r_in, double r_out, double dr, double* acc0w, double* potw) {
    orc_changeover c = co_set(r_in, r_out);
    *acc0w = co_acc0w(&c, dr);
    *potw = co_potw(&c, dr);
}

void orc_changeover_pair(pb_PtclCorr* pi, const pb_PtclCorr* pj, double eps2, double r_out_g, double G, int replay_fp32) {
    const double drx = pi->pos.x - pj->pos.x, dry = pi->pos.y - pj->pos.y, drz = pi->pos.z - pj->pos.z;
    const double dr2 = drx * drx + dry * dry + drz * drz;
    const double dr2_eps = dr2 + eps2;
    const double drinv = 1.0 / sqrt(dr2_eps);
    double gmor = G * pj->mass * drinv;
    const double drinv2 = drinv * drinv;
    const double gmor3 = gmor * drinv2;
    const double dr_eps = drinv * dr2_eps;
    const orc_changeover chi = co_set(pi->r_in, pi->r_out), chj = co_set(pj->r_in, pj->r_out);
    const orc_changeover* ch = (chi.r_out > chj.r_out) ? &chi : &chj;
    const double k = 1.0 - co_acc0w(ch, dr_eps);
    double gmor_max;
    if (replay_fp32) {
        const float r_out_32 = (float)r_out_g;
        const float r_out2 = r_out_32 * r_out_32;
        const float dx = (float)pi->pos.x - (float)pj->pos.x, dy = (float)pi->pos.y - (float)pj->pos.y, dz = (float)pi->pos.z - (float)pj->pos.z;
        const float dr2_eps_32 = (dx * dx + dy * dy + dz * dz) + (float)eps2;
        const float dr2_max = (dr2_eps_32 > r_out2) ? dr2_eps_32 : r_out2;
        const float drinv_max = (float)(1.0 / sqrt((double)dr2_max));
        const float gmor_max32 = (float)(G * pj->mass * (double)drinv_max);
        const float drinv2_max = drinv_max * drinv_max;
        const float gmor3_max = gmor_max32 * drinv2_max;
        const double gk = gmor3 * k;
        pi->acc.x -= gk * drx - (double)(gmor3_max * dx);
        pi->acc.y -= gk * dry - (double)(gmor3_max * dy);
        pi->acc.z -= gk * drz - (double)(gmor3_max * dz);
        gmor_max = (double)gmor_max32;
    } else {
        const double r_out2 = r_out_g * r_out_g;
        const double dr2_max = (dr2_eps > r_out2) ? dr2_eps : r_out2;
        const double drinv_max = 1.0 / sqrt(dr2_max);
        gmor_max = G * pj->mass * drinv_max;
        const double drinv2_max = drinv_max * drinv_max;
        const double gmor3_max = gmor_max * drinv2_max;
        const double f = gmor3 * k - gmor3_max;
        pi->acc.x -= f * drx; pi->acc.y -= f * dry; pi->acc.z -= f * drz;
    }
    const double kpot = 1.0 - co_potw(ch, dr_eps);
    if (pj->status == 0.0 && pj->mass_backup == 0.0) {            /* single */
        pi->pot_soft -= gmor * kpot - gmor_max;
        pi->pot_tot -= (gmor - gmor_max);
    } else if (pj->status < 0.0) {                                /* member: mass is zero, use the backup mass */
        gmor = G * pj->mass_backup * drinv;
        pi->pot_soft -= gmor * kpot - gmor_max;
        pi->pot_tot -= (gmor - gmor_max);
    } else {                                                      /* (orbital) artificial */
        pi->pot_soft += gmor_max;
        pi->pot_tot += gmor_max;
    }
}

void orc_correct_force_tree_neighbor(pb_PtclCorr* p, int n, const int* nb_off, const int* nb_idx, const pb_PtclCorr* pj,
                                     double eps2, double r_out_g, double G, double status_no_cm, int replay_fp32) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; i++) {
        pb_PtclCorr* pi = p + i;
        const int single = (pi->status == 0.0 && pi->mass_backup == 0.0);
        const int member_no_cm = (pi->status < 0.0) && (pi->status == status_no_cm);
        if (single || member_no_cm) {
            const double pot_cor = G * pi->mass / r_out_g;
            pi->pot_tot += pot_cor;
            pi->pot_soft += pot_cor;
        }
        for (int k = nb_off[i]; k < nb_off[i + 1]; k++) {
            const pb_PtclCorr* q = pj + nb_idx[k];
            if (q->id == pi->id) continue;
            orc_changeover_pair(pi, q, eps2, r_out_g, G, replay_fp32);
        }
    }
}
