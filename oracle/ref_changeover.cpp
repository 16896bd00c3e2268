// ref_changeover.cpp — compiles the REFERENCE's changeover correction for the oracle to be pinned against:
//   * class ChangeOver, verbatim from $(REFERENCE)/src/changeover.hpp (included where it lies);
//   * SystemHard::calcAccPotShortWithLinearCutoff(Tpi&, const EPJSoft&), extracted from
//     $(REFERENCE)/src/hard.hpp at build time into _ref/ref_changeover_pair.inc (oracle/Makefile, never
//     committed) and compiled here against minimal stand-ins for the FDPS / PeTar types it touches.
// TEST INFRASTRUCTURE ONLY.  Built twice: with -DUSE_GPU (float replay branch) and without (-DP3T_64BIT).
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include "petar_b200_types.h"

namespace PS {
typedef double F64; typedef float F32; typedef long long S64; typedef int S32;
template <class T> struct Vector3 {
    T x, y, z;
    Vector3() : x(0), y(0), z(0) {}
    Vector3(T a, T b, T c) : x(a), y(b), z(c) {}
    template <class U> operator Vector3<U>() const { return Vector3<U>((U)x, (U)y, (U)z); }
    Vector3 operator-(const Vector3& r) const { return Vector3(x - r.x, y - r.y, z - r.z); }
    Vector3 operator+(const Vector3& r) const { return Vector3(x + r.x, y + r.y, z + r.z); }
    T operator*(const Vector3& r) const { return (x * r.x) + (y * r.y) + (z * r.z); }      // FDPS: dot product
    Vector3 operator*(const T s) const { return Vector3(x * s, y * s, z * s); }
    friend Vector3 operator*(const T s, const Vector3& v) { return Vector3(s * v.x, s * v.y, s * v.z); }
    const Vector3& operator-=(const Vector3& r) { x -= r.x; y -= r.y; z -= r.z; return *this; }
    const Vector3& operator+=(const Vector3& r) { x += r.x; y += r.y; z += r.z; return *this; }
};
typedef Vector3<F64> F64vec; typedef Vector3<F32> F32vec;
}
#include "changeover.hpp"

struct ForceSoft { static PS::F64 grav_const; };
struct EPISoft { static PS::F64 eps, r_out; };
PS::F64 ForceSoft::grav_const = 1.0, EPISoft::eps = 0.0, EPISoft::r_out = 0.0;

struct ArtificialParticleInformation {         // reference src/artificial_particles.hpp:20-86: the three predicates used
    PS::F64 mass_backup, status;
    bool isMember() const { return status < 0.0; }
    bool isSingle() const { return status == 0.0 && mass_backup == 0.0; }
    PS::F64 getMassBackup() const { return mass_backup; }
};
struct GroupData { ArtificialParticleInformation artificial; };
struct EPJSoft { PS::F64 mass; PS::F64vec pos; PS::F64 r_in, r_out; GroupData group_data; };
struct PtclI { PS::F64vec pos, acc; PS::F64 pot_tot, pot_soft; ChangeOver changeover; };

struct SystemHardRef {
#include "_ref/ref_changeover_pair.inc"
};

extern "C" {
void ref_changeover_w(double r_in, double r_out, double dr, double* acc0w, double* potw) {
    ChangeOver c; c.setR(r_in, r_out);
    *acc0w = c.calcAcc0W(dr); *potw = c.calcPotW(dr);
}
void ref_changeover_pair(pb_PtclCorr* pi, const pb_PtclCorr* pj, double eps, double r_out, double G) {
    ForceSoft::grav_const = G; EPISoft::eps = eps; EPISoft::r_out = r_out;
    PtclI I; EPJSoft J;
    I.pos = PS::F64vec(pi->pos.x, pi->pos.y, pi->pos.z); I.acc = PS::F64vec(pi->acc.x, pi->acc.y, pi->acc.z);
    I.pot_tot = pi->pot_tot; I.pot_soft = pi->pot_soft; I.changeover.setR(pi->r_in, pi->r_out);
    J.mass = pj->mass; J.pos = PS::F64vec(pj->pos.x, pj->pos.y, pj->pos.z); J.r_in = pj->r_in; J.r_out = pj->r_out;
    J.group_data.artificial.mass_backup = pj->mass_backup; J.group_data.artificial.status = pj->status;
    SystemHardRef::calcAccPotShortWithLinearCutoff(I, J);
    pi->acc.x = I.acc.x; pi->acc.y = I.acc.y; pi->acc.z = I.acc.z; pi->pot_tot = I.pot_tot; pi->pot_soft = I.pot_soft;
}
}
