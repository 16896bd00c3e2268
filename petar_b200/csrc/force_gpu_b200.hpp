// force_gpu_b200.hpp — stand-alone declaration of the interface that PeTar's
// src/force_gpu_cuda.hpp declares, over POD mirrors of the FDPS / PeTar types.
//
// Inside a real PeTar build this header is NOT used: force_gpu_b200.cpp includes PeTar's own
// src/force_gpu_cuda.hpp (reference src/petar.hpp:66 includes it by name, so the symbols must be
// the ones that header declares).  It exists so that the same translation unit can be compiled
// and tested in a container that has neither FDPS nor SDAR (-DPB_STANDALONE_MIRRORS).
//
// Interface mirrored (names, members, argument order and meaning):
//   GPUProfile  {copy, send, recv, calc; n_profile}            reference src/force_gpu_cuda.hpp:8-53
//   GPUCounter  {n_walk, n_epi, n_epj, n_spj, n_call; n_counter}                          :55-92
//   SPJSoft                                                                               :95-99
//   CalcForceWithLinearCutoffCUDAMultiWalk {my_rank, eps2, rcut2, G; operator()}         :103-133
//   CalcForceWithLinearCutoffCUDA          {my_rank, eps2, rcut2, G; operator()}         :137-162
//   RetrieveForceCUDA                                                                    :165-168
//   Tprofile / NumCounter (only what the functors touch)     reference src/profile.hpp:176-265
#pragma once
#include <chrono>
#include <cstdint>
#include "petar_b200_types.h"

namespace PS {
typedef int32_t S32;
typedef int64_t S64;
typedef double  F64;
inline F64 GetWtime() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
} // namespace PS

typedef pb_EPISoft   EPISoft;
typedef pb_EPJSoft   EPJSoft;
typedef pb_ForceSoft ForceSoft;
#ifdef USE_QUAD
typedef pb_SPJQuad SPJSoft;
#else
typedef pb_SPJMono SPJSoft;
#endif

#ifdef GPU_PROFILE
struct Tprofile {
    PS::F64 time, tbar;
    const char* name;
    explicit Tprofile(const char* n) : time(0.0), tbar(0.0), name(n) {}
    void start() { time -= PS::GetWtime(); }
    void end() { const PS::F64 t = PS::GetWtime(); tbar += t; time += t; }
    void reset() { time = 0.0; tbar = 0.0; }
};
struct NumCounter {
    PS::S64 n;
    const char* name;
    explicit NumCounter(const char* nm) : n(0), name(nm) {}
    NumCounter& operator+=(PS::S64 v) { n += v; return *this; }
    NumCounter& operator=(PS::S64 v) { n = v; return *this; }
};
struct GPUProfile {
    Tprofile copy, send, recv, calc;
    const PS::S32 n_profile;
    GPUProfile() : copy("copy       "), send("send       "), recv("receive    "), calc("calc_force "), n_profile(4) {}
    void clear() { copy.reset(); send.reset(); recv.reset(); calc.reset(); }
};
struct GPUCounter {
    NumCounter n_walk, n_epi, n_epj, n_spj, n_call;
    const PS::S32 n_counter;
    GPUCounter() : n_walk("n_walk "), n_epi("n_epi  "), n_epj("n_epj  "), n_spj("n_spj  "), n_call("n_call "), n_counter(5) {}
    void clear() { n_walk = 0; n_epi = 0; n_epj = 0; n_spj = 0; n_call = 0; }
};
extern GPUProfile gpu_profile;
extern GPUCounter gpu_counter;
#endif

#ifdef PARTICLE_SIMULATOR_GPU_MULIT_WALK_INDEX
struct CalcForceWithLinearCutoffCUDAMultiWalk {
    PS::S32 my_rank;
    PS::F64 eps2, rcut2, G;
    CalcForceWithLinearCutoffCUDAMultiWalk() {}
    CalcForceWithLinearCutoffCUDAMultiWalk(PS::S32 r, PS::F64 e2, PS::F64 rc2, PS::F64 g) : my_rank(r), eps2(e2), rcut2(rc2), G(g) {}
    void initialize(PS::S32 r, PS::F64 e2, PS::F64 rc2, PS::F64 g) { my_rank = r; eps2 = e2; rcut2 = rc2; G = g; }
    PS::S32 operator()(const PS::S32 tag, const PS::S32 n_walk,
                       const EPISoft** epi, const PS::S32* n_epi,
                       const PS::S32** id_epj, const PS::S32* n_epj,
                       const PS::S32** id_spj, const PS::S32* n_spj,
                       const EPJSoft* epj, const PS::S32 n_epj_tot,
                       const SPJSoft* spj, const PS::S32 n_spj_tot,
                       const bool send_flag);
};
#else
struct CalcForceWithLinearCutoffCUDA {
    PS::S32 my_rank;
    PS::F64 eps2, rcut2, G;
    CalcForceWithLinearCutoffCUDA() {}
    CalcForceWithLinearCutoffCUDA(PS::S32 r, PS::F64 e2, PS::F64 rc2, PS::F64 g) : my_rank(r), eps2(e2), rcut2(rc2), G(g) {}
    void initialize(PS::S32 r, PS::F64 e2, PS::F64 rc2, PS::F64 g) { my_rank = r; eps2 = e2; rcut2 = rc2; G = g; }
    PS::S32 operator()(const PS::S32 tag, const PS::S32 n_walk,
                       const EPISoft* epi[], const PS::S32 n_epi[],
                       const EPJSoft* epj[], const PS::S32 n_epj[],
                       const SPJSoft* spj[], const PS::S32 n_spj[]);
};
#endif

PS::S32 RetrieveForceCUDA(const PS::S32 tag, const PS::S32 n_walk, const PS::S32* ni, ForceSoft** force);
