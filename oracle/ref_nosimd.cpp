// ref_nosimd.cpp — compiles the REFERENCE's own fp64 functors for the oracle to be pinned against, bit for bit:
//   SearchNeighborEpEpNoSimd, CalcForceEpEpWithLinearCutoffNoSimd, CalcForceEpSpMonoNoSimd,
//   CalcForceEpSpQuadNoSimd, CalcForcePPNoSimd          ($(REFERENCE)/src/soft_force.hpp:10-236)
// The functor text is extracted from the reference source at build time into _ref/ref_nosimd_functors.inc
// (oracle/Makefile, never committed) and compiled here against minimal stand-ins for the FDPS / PeTar types it
// touches.  The stand-ins have the byte layout of the pb_* mirrors (include/petar_b200_types.h), so the arrays the
// tests hold are passed straight through.  PS::Vector3 / PS::MatrixSym3 operators follow FDPS's definitions
// (dot product (x*x)+(y*y)+(z*z), scalar * vector = component-wise product, trace xx+yy+zz).
// TEST INFRASTRUCTURE ONLY.  Built with -O2 -ffp-contract=off, as the oracle restatement is.
#include <algorithm>
#include <cassert>
#include <cmath>
#include "petar_b200_types.h"

namespace PS {
typedef double F64; typedef float F32; typedef long long S64; typedef int S32;
template <class T> struct Vector3 {
    T x, y, z;
    Vector3() : x(0), y(0), z(0) {}
    Vector3(T s) : x(s), y(s), z(s) {}
    Vector3(T a, T b, T c) : x(a), y(b), z(c) {}
    Vector3 operator-(const Vector3& r) const { return Vector3(x - r.x, y - r.y, z - r.z); }
    Vector3 operator+(const Vector3& r) const { return Vector3(x + r.x, y + r.y, z + r.z); }
    T operator*(const Vector3& r) const { return (x * r.x) + (y * r.y) + (z * r.z); }
    Vector3 operator*(const T s) const { return Vector3(x * s, y * s, z * s); }
    friend Vector3 operator*(const T s, const Vector3& v) { return (v * s); }
    const Vector3& operator-=(const Vector3& r) { x -= r.x; y -= r.y; z -= r.z; return *this; }
    const Vector3& operator+=(const Vector3& r) { x += r.x; y += r.y; z += r.z; return *this; }
    T operator[](int k) const { return (&x)[k]; }
};
template <class T> struct MatrixSym3 {
    T xx, yy, zz, xy, xz, yz;
    T getTrace() const { return (xx + yy + zz); }
};
typedef Vector3<F64> F64vec; typedef MatrixSym3<F64> F64mat;
}

struct ForceSoft { PS::F64vec acc; PS::F64 pot; PS::S64 n_ngb; static PS::F64 grav_const; };
struct EPISoft { PS::S64 id; PS::F64vec pos; PS::F64 r_search; PS::S32 rank_org, type; static PS::F64 eps, r_out; };
struct EPJSoft {
    PS::S64 id; PS::F64 mass; PS::F64vec pos, vel; PS::F64 r_in, r_out, r_search, r_scale_next;
    PS::S64 group_data[2]; PS::S32 rank_org, adr_org;
};
struct SPJQuad {                                   // PS::SPJQuadrupoleInAndOut
    PS::F64 mass; PS::F64vec pos; PS::F64mat quad;
    PS::F64vec getPos() const { return pos; }
    PS::F64 getCharge() const { return mass; }
};
PS::F64 ForceSoft::grav_const = 1.0, EPISoft::eps = 0.0, EPISoft::r_out = 0.0;

static_assert(sizeof(ForceSoft) == sizeof(pb_ForceSoft) && sizeof(EPISoft) == sizeof(pb_EPISoft) &&
              sizeof(EPJSoft) == sizeof(pb_EPJSoft) && sizeof(SPJQuad) == sizeof(pb_SPJQuad), "stand-in layouts");

#include "_ref/ref_nosimd_functors.inc"

extern "C" {
void ref_nosimd_search(const void* epi, int ni, const void* epj, int nj, void* force) {
    SearchNeighborEpEpNoSimd()((const EPISoft*)epi, ni, (const EPJSoft*)epj, nj, (ForceSoft*)force);
}
void ref_nosimd_epep(const void* epi, int ni, const void* epj, int nj, void* force, double eps, double r_out, double G) {
    EPISoft::eps = eps; EPISoft::r_out = r_out; ForceSoft::grav_const = G;
    CalcForceEpEpWithLinearCutoffNoSimd()((const EPISoft*)epi, ni, (const EPJSoft*)epj, nj, (ForceSoft*)force);
}
void ref_nosimd_epsp_mono(const void* epi, int ni, const void* spj, int nj, void* force, double eps, double G) {
    EPISoft::eps = eps; ForceSoft::grav_const = G;
    CalcForceEpSpMonoNoSimd()((const EPISoft*)epi, ni, (const SPJQuad*)spj, nj, (ForceSoft*)force);
}
void ref_nosimd_epsp_quad(const void* epi, int ni, const void* spj, int nj, void* force, double eps, double G) {
    EPISoft::eps = eps; ForceSoft::grav_const = G;
    CalcForceEpSpQuadNoSimd()((const EPISoft*)epi, ni, (const SPJQuad*)spj, nj, (ForceSoft*)force);
}
void ref_nosimd_pp(const void* epi, int ni, const void* epj, int nj, void* force, double G) {
    ForceSoft::grav_const = G;
    CalcForcePPNoSimd<EPISoft, EPJSoft>()((const EPISoft*)epi, ni, (const EPJSoft*)epj, nj, (ForceSoft*)force);
}
}
