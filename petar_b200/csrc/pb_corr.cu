// pb_corr.cu — changeover correction of the soft force on the device.
//
// One thread per particle walks its neighbour list and applies, pair by pair and in list order, what
// SystemHard::calcAccPotShortWithLinearCutoff does (reference src/hard.hpp:1408-1476) inside
// correctForceWithCutoffTreeNeighborOneParticleImp (:1655-1691): the fp64 Newtonian pair force weighted with
// the changeover function (src/changeover.hpp:294-334, the larger-r_out member of the pair decides, :366-375)
// replaces the linear-cutoff term the force kernel added for that pair, and the potentials are corrected
// according to the neighbour's kind (single / group member with backup mass / artificial).
//
// Everything is IEEE fp64 (the float replay of the reference's USE_GPU branch in IEEE fp32); this file is
// compiled with -fmad=false so that each operation rounds exactly as the host compilers' code does — the
// results are bit-identical to the CPU oracle and to the reference function compiled in oracle/_ref.
#include "pb_device.h"

namespace pb {
namespace {

struct ChangeOver { double r_in, r_out, norm, coff, pot_off; };

__device__ __forceinline__ ChangeOver co_set(double r_in, double r_out) {     // ChangeOver::setR(r_in, r_out)
    ChangeOver c;
    c.r_in = r_in; c.r_out = r_out;
    c.norm = 1.0 / (r_out - r_in);
    c.coff = (r_out - r_in) / (r_out + r_in);
    c.pot_off = (1.0 + c.coff) / r_out;
    return c;
}
__device__ __forceinline__ double co_potw(const ChangeOver& c, double dr) {   // ChangeOver::calcPotW
    const double x = (dr - c.r_in) * c.norm;
    double k = 1.0;
    if (x >= 1.0) k = c.pot_off * dr;
    else if (x > 0.0) {
        const double x2 = x * x, x3 = x2 * x, x5 = x2 * x3;
        k -= c.coff * x5 * (5.0 * x3 - 20.0 * x2 + 28.0 * x - 14.0);
    }
    return k;
}
__device__ __forceinline__ double co_acc0w(const ChangeOver& c, double dr) {  // ChangeOver::calcAcc0W
    double x = (dr - c.r_in) * c.norm;
    x = (x < 1.0) ? x : 1.0;
    x = (x > 0.0) ? x : 0.0;
    const double x_1 = x - 1, x_2 = x_1 * x_1, x_4 = x_2 * x_2;
    const double x2 = x * x, x3 = x2 * x, x4 = x2 * x2;
    return x_4 * (1.0 + 4.0 * x + 10.0 * x2 + 20.0 * x3 + 35.0 * c.coff * x4);
}

__global__ void __launch_bounds__(128)
corr_kernel(int n_i, const CorrI* __restrict__ pi, const CorrJ* __restrict__ pj,
            const int* __restrict__ nb_off, const int* __restrict__ nb_idx, CorrOut* __restrict__ out, CorrParams prm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_i) return;
    const CorrI I = pi[i];
    double ax = I.ax, ay = I.ay, az = I.az, pot_tot = I.pot_tot, pot_soft = I.pot_soft;

    // self-potential of the linear cutoff: singles, and members that have no c.m. particle
    const bool single = (I.p.status == 0.0 && I.p.mass_bk == 0.0);
    const bool member_no_cm = (I.p.status < 0.0) && (I.p.status == prm.status_no_cm);
    if (single || member_no_cm) {
        const double pot_cor = prm.G * I.p.mass / prm.r_out;
        pot_tot += pot_cor;
        pot_soft += pot_cor;
    }
    const ChangeOver chi = co_set(I.p.r_in, I.p.r_out);

    const int k1 = nb_off[i + 1];
    for (int k = nb_off[i]; k < k1; ++k) {
        const CorrJ J = pj[nb_idx[k]];
        if (J.id == I.p.id) continue;
        const double drx = I.p.x - J.x, dry = I.p.y - J.y, drz = I.p.z - J.z;
        const double dr2 = drx * drx + dry * dry + drz * drz;
        const double dr2_eps = dr2 + prm.eps2;
        const double drinv = 1.0 / sqrt(dr2_eps);
        double gmor = prm.G * J.mass * drinv;
        const double drinv2 = drinv * drinv;
        const double gmor3 = gmor * drinv2;
        const double dr_eps = drinv * dr2_eps;
        const ChangeOver chj = co_set(J.r_in, J.r_out);
        const ChangeOver& ch = (chi.r_out > chj.r_out) ? chi : chj;
        const double kacc = 1.0 - co_acc0w(ch, dr_eps);
        double gmor_max;
        if (prm.replay_fp32) {
            // the reference's USE_GPU branch: the cutoff term re-evaluated in float from absolute coordinates
            const float r_out_32 = (float)prm.r_out;
            const float r_out2 = r_out_32 * r_out_32;
            const float dx = (float)I.p.x - (float)J.x, dy = (float)I.p.y - (float)J.y, dz = (float)I.p.z - (float)J.z;
            const float dr2_eps_32 = (dx * dx + dy * dy + dz * dz) + (float)prm.eps2;
            const float dr2_max = (dr2_eps_32 > r_out2) ? dr2_eps_32 : r_out2;
            const float drinv_max = (float)(1.0 / sqrt((double)dr2_max));
            const float gmor_max32 = (float)(prm.G * J.mass * (double)drinv_max);
            const float drinv2_max = drinv_max * drinv_max;
            const float gmor3_max = gmor_max32 * drinv2_max;
            const double gk = gmor3 * kacc;
            ax -= gk * drx - (double)(gmor3_max * dx);
            ay -= gk * dry - (double)(gmor3_max * dy);
            az -= gk * drz - (double)(gmor3_max * dz);
            gmor_max = (double)gmor_max32;
        } else {
            const double r_out2 = prm.r_out * prm.r_out;
            const double dr2_max = (dr2_eps > r_out2) ? dr2_eps : r_out2;
            const double drinv_max = 1.0 / sqrt(dr2_max);
            gmor_max = prm.G * J.mass * drinv_max;
            const double drinv2_max = drinv_max * drinv_max;
            const double gmor3_max = gmor_max * drinv2_max;
            const double f = gmor3 * kacc - gmor3_max;
            ax -= f * drx; ay -= f * dry; az -= f * drz;
        }
        const double kpot = 1.0 - co_potw(ch, dr_eps);
        if (J.status == 0.0 && J.mass_bk == 0.0) {            // single
            pot_soft -= gmor * kpot - gmor_max;
            pot_tot -= (gmor - gmor_max);
        } else if (J.status < 0.0) {                          // member: its mass is zero in the soft force, use the backup
            gmor = prm.G * J.mass_bk * drinv;
            pot_soft -= gmor * kpot - gmor_max;
            pot_tot -= (gmor - gmor_max);
        } else {                                              // (orbital) artificial
            pot_soft += gmor_max;
            pot_tot += gmor_max;
        }
    }
    CorrOut o;
    o.ax = ax; o.ay = ay; o.az = az; o.pot_tot = pot_tot; o.pot_soft = pot_soft;
    out[i] = o;
}

} // namespace

cudaError_t launch_corr(cudaStream_t s, int n_i, const CorrI* pi, const CorrJ* pj,
                        const int* nb_off, const int* nb_idx, CorrOut* out, CorrParams prm)
{
    if (n_i <= 0) return cudaSuccess;
    corr_kernel<<<(n_i + 127) / 128, 128, 0, s>>>(n_i, pi, pj, nb_off, nb_idx, out, prm);
    return cudaGetLastError();
}

} // namespace pb
