// pb_pairs.cuh — the arithmetic both force kernels share (pb_kernels.cu: one task per CTA / persistent with a block-wide
// tile ring; pb_kernels_ws.cu: warp-specialised persistent kernel): shared-memory tile layouts, staging of one j into
// a tile (origin shift, pair interleave, near test), the EP-EP and EP-SP pair loops, the compensated running sum.
#pragma once
#include "pb_device.h"

namespace pb {

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int NR>
__device__ __forceinline__ float2 rsqrt2(float2 x) {
    float2 y = make_float2(rsqrt_approx(x.x), rsqrt_approx(x.y));
    if (NR >= 1) {
        // y <- y * (1.5 - 0.5 x y^2)
        float2 h  = __fmul2_rn(x, make_float2(-0.5f, -0.5f));
        float2 y2 = __fmul2_rn(y, y);
        float2 c  = __ffma2_rn(h, y2, make_float2(1.5f, 1.5f));
        y = __fmul2_rn(y, c);
    }
    return y;
}

__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }

// compensated running sum (Kahan-Babuska / Neumaier): sum + comp carries the exact-ish total
struct KSum {
    float s, c;
    __device__ __forceinline__ void init() { s = 0.f; c = 0.f; }
    __device__ __forceinline__ void add(float v) {
        float t = s + v;
        float bp = t - s;
        c += (s - (t - bp)) + (v - bp);   // two-sum error term
        s = t;
    }
    __device__ __forceinline__ double value() const { return (double)s + (double)c; }
};

// ------------------------------------------------------------------------------------------
// shared-memory tiles
// ------------------------------------------------------------------------------------------
struct EpTile {                       // 128 pairs, 64 B per pair
    float4 a[kTilePairs];             // {x0, x1, y0, y1}   origin-relative position, hi part
    float4 b[kTilePairs];             // {z0, z1, m0, m1}
    float2 c[kTilePairs];             // {r_search0^2, r_search1^2}
    float4 al[kTilePairs];            // {xl0, xl1, yl0, yl1} lo part (only read in "near" segments)
    float2 bl[kTilePairs];            // {zl0, zl1}
    float4 ah[kTilePairs];            // {X0, X1, Y0, Y1}   ABSOLUTE position cast to float (coords = 2, near segments)
    float2 bh[kTilePairs];            // {Z0, Z1}
};
struct SpTile {                       // 128 pairs, 96 B per pair
    float4 q0[kTilePairs];            // {x0, x1, y0, y1}
    float4 q1[kTilePairs];            // {z0, z1, m0, m1}
    float4 q2[kTilePairs];            // {q'xx0, q'xx1, q'yy0, q'yy1}   (q' = 3q - tr I, traceless)
    float4 q3[kTilePairs];            // {qzz0, qzz1, qxy0, qxy1}
    float4 q4[kTilePairs];            // {qxz0, qxz1, qyz0, qyz1}
    float2 q5[kTilePairs];            // {-eps2 tr0, -eps2 tr1}
};
constexpr int kTileBufs = 3;         // tile ring: tile k+1 is written while tile k (and, by slower warps, tile k-1) is read
union __align__(16) Smem {
    EpTile ep[kTileBufs];
    SpTile sp[kTileBufs];
    double red[kWarpsPerCta][4][32];  // cross-warp combine when jsplit > 1
};

// Tile hand-over without a block-wide stall: every thread ARRIVES on the tile's mbarrier right after it has stored its
// j of that tile and WAITS on it only when it starts to compute the tile — one whole tile of work later.  With three
// buffers a warp may run a tile ahead of the slowest one, so the warps of a CTA drift apart and one warp's staging
// overlaps the others' arithmetic (a __syncthreads per tile made all eight warps stage, then stall, together).
struct TileBars {
    alignas(8) unsigned long long full[kTileBufs];
    __device__ __forceinline__ void init(int tid) {                       // followed by a __syncthreads
        if (tid == 0) {
            for (int b = 0; b < kTileBufs; ++b) {
                const unsigned bar = (unsigned)__cvta_generic_to_shared(&full[b]);
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(kThreads) : "memory");
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __device__ __forceinline__ void arrive(int b) {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&full[b]);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
    }
    __device__ __forceinline__ void wait(int b, unsigned parity) {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&full[b]);
        asm volatile("{\n"
                     ".reg .pred P1;\n"
                     "TB_WAIT:\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                     "@P1 bra TB_DONE;\n"
                     "bra TB_WAIT;\n"
                     "TB_DONE:\n"
                     "}" :: "r"(bar), "r"(parity) : "memory");
    }
};

constexpr float kPadPos = 1.0e10f;    // padded j: far away, zero mass, never a neighbour

// ------------------------------------------------------------------------------------------
// EP-EP: clamped ("linear cutoff") Plummer force + neighbour count
//   dx = xj - xi ; r2 = eps2 + dx.dx ; n += (r2 < max(rs_i, rs_j)^2)
//   r2c = max(r2, rcut2) ; rinv = rsqrt(r2c) ; acc += m rinv^3 dx ; pot -= m rinv
// (reference src/force_gpu_cuda.cu:58-74; G is applied once at the end, as soft_force.hpp:73-77)
// ------------------------------------------------------------------------------------------
struct EpRegs { float4 a, b; };       // one gathered j: {xh,yh,zh,m}, {xl,yl,zl,rs}

// (list indices come through IdPipe below; id < 0 marks a padding slot past the end of a chunk)
__device__ __forceinline__ EpRegs ep_load_j(const float4* __restrict__ epj, int id) {
    EpRegs r;
    if (id >= 0) {
        r.a = __ldg(epj + 2 * (size_t)id);
        r.b = __ldg(epj + 2 * (size_t)id + 1);
    } else {
        r.a = make_float4(0.f, 0.f, 0.f, 0.f);
        r.b = make_float4(0.f, 0.f, 0.f, -1.f);
    }
    return r;
}
// writes j into the pair-interleaved tile; returns whether this j is "near" the walk: within
// sqrt(max(rs_j^2, w.rsi2max)) of the bounding box of the walk's i-particles (0.1 % margin, far
// above fp32 rounding).  w.rsi2max covers max rs_i^2 AND the radius inside which the fp32
// rounding of an origin-relative coordinate (ulp of the box half-size) is not negligible against
// the pair separation.  Segments without a near j run the fast loop: no neighbour test (none is
// possible), single-float dx.  Near segments run the exact loop: neighbour test and
// dx = (xj_hi - xi_hi) + (xj_lo - xi_lo), the difference of the two-float relative positions.
__device__ __forceinline__ bool ep_store(EpTile& t, int tid, int id, const EpRegs& r, const Walk& w, int abs_mode) {
    float x, y, z, xl = 0.f, yl = 0.f, zl = 0.f, m, rs2;
    bool near = false;
    if (id >= 0) {
        rel_hilo(r.a.x, r.b.x, w.ohx, w.olx, x, xl);
        rel_hilo(r.a.y, r.b.y, w.ohy, w.oly, y, yl);
        rel_hilo(r.a.z, r.b.z, w.ohz, w.olz, z, zl);
        if (abs_mode == 1) { xl = 0.f; yl = 0.f; zl = 0.f; }   // reference arithmetic: dx = float(xj) - float(xi)
        m = r.a.w;
        rs2 = r.b.w * r.b.w;
        const float ex = fmaxf(fabsf(x) - w.hx, 0.f);
        const float ey = fmaxf(fabsf(y) - w.hy, 0.f);
        const float ez = fmaxf(fabsf(z) - w.hz, 0.f);
        const float d2 = ex * ex + ey * ey + ez * ez;      // hx = +inf (culling off): d2 = 0, always near
        near = !(d2 * 0.999f >= fmaxf(rs2, w.rsi2max));
    } else {
        x = y = z = kPadPos; m = 0.f; rs2 = -1.f;
    }
    const int p = tid >> 1, s = tid & 1;
    float* a = reinterpret_cast<float*>(&t.a[p]);
    float* b = reinterpret_cast<float*>(&t.b[p]);
    float* c = reinterpret_cast<float*>(&t.c[p]);
    float* al = reinterpret_cast<float*>(&t.al[p]);
    float* bl = reinterpret_cast<float*>(&t.bl[p]);
    a[s] = x; a[2 + s] = y;
    b[s] = z; b[2 + s] = m;
    c[s] = rs2;
    al[s] = xl; al[2 + s] = yl;
    bl[s] = zl;
    if (abs_mode == 2) {                                   // float(x_j): what the CPU replay subtracts with (src/hard.hpp:1431)
        float* ah = reinterpret_cast<float*>(&t.ah[p]);
        float* bh = reinterpret_cast<float*>(&t.bh[p]);
        ah[s] = r.a.x; ah[2 + s] = r.a.y;
        bh[s] = r.a.z;
    }
    return near;
}

// NEAR = 0: fast loop (13 packed FP ops per pair).  NEAR = 1: exact dx + neighbour count.
// NEAR = 2 (coords = 2, the drop-in default): as 1, and a pair that passes the neighbour test is evaluated from
// dx = float(x_j) - float(x_i), ABSOLUTE coordinates cast to float — the very term PeTar's CPU changeover correction
// re-computes in float and subtracts afterwards (`dr_32`, reference src/hard.hpp:1428-1442), so that it cancels; every
// other pair keeps the walk-relative two-float dx.  (xih, yih, zih) = float(x_i).
template <int NR, int NEAR>
__device__ __forceinline__ void ep_pairs(const EpTile& t, int p0, int p1,
                                         float xi, float yi, float zi, float xil, float yil, float zil, float rsi2,
                                         float xih, float yih, float zih,
                                         float eps2, float rcut2, float rinv_cut,
                                         float2& ax, float2& ay, float2& az, float2& pt, float2& cf) {
    const float2 nxi = bc(-xi), nyi = bc(-yi), nzi = bc(-zi), e2 = bc(eps2);
    const float2 nxil = bc(-xil), nyil = bc(-yil), nzil = bc(-zil);
#pragma unroll kPairUnroll
    for (int p = p0; p < p1; ++p) {
        const float4 A = t.a[p];
        const float4 B = t.b[p];
        float2 dx = __fadd2_rn(make_float2(A.x, A.y), nxi);
        float2 dy = __fadd2_rn(make_float2(A.z, A.w), nyi);
        float2 dz = __fadd2_rn(make_float2(B.x, B.y), nzi);
        if (NEAR) {
            const float4 AL = t.al[p];
            const float2 BL = t.bl[p];
            dx = __fadd2_rn(dx, __fadd2_rn(make_float2(AL.x, AL.y), nxil));
            dy = __fadd2_rn(dy, __fadd2_rn(make_float2(AL.z, AL.w), nyil));
            dz = __fadd2_rn(dz, __fadd2_rn(BL, nzil));
        }
        float2 r2 = __ffma2_rn(dx, dx, e2);
        r2 = __ffma2_rn(dy, dy, r2);
        r2 = __ffma2_rn(dz, dz, r2);
        if (NEAR) {
            // neighbour flags as 0.0f/1.0f, summed packed; exact (counts per tile are tiny)
            const float2 C = t.c[p];
            const bool h0 = r2.x < fmaxf(C.x, rsi2), h1 = r2.y < fmaxf(C.y, rsi2);
            cf = __fadd2_rn(cf, make_float2(h0 ? 1.f : 0.f, h1 ? 1.f : 0.f));
            if (NEAR == 2 && __any_sync(__activemask(), h0 || h1)) {        // rare: some lane of the warp has a neighbour in this pair
                const float4 AH = t.ah[p];
                const float2 BH = t.bh[p];
                const float2 ex = __fadd2_rn(make_float2(AH.x, AH.y), bc(-xih));
                const float2 ey = __fadd2_rn(make_float2(AH.z, AH.w), bc(-yih));
                const float2 ez = __fadd2_rn(BH, bc(-zih));
                dx = make_float2(h0 ? ex.x : dx.x, h1 ? ex.y : dx.y);
                dy = make_float2(h0 ? ey.x : dy.x, h1 ? ey.y : dy.y);
                dz = make_float2(h0 ? ez.x : dz.x, h1 ? ez.y : dz.y);
                r2 = __ffma2_rn(dx, dx, e2);
                r2 = __ffma2_rn(dy, dy, r2);
                r2 = __ffma2_rn(dz, dz, r2);
                // inside the cutoff the replay's 1/r is the constant float(1 / sqrt(double(r_out_32^2))) (src/hard.hpp:1434-1436):
                // use that very value for the neighbour instead of the 2-ulp MUFU approximation
                const float2 r2c_ = make_float2(fmaxf(r2.x, rcut2), fmaxf(r2.y, rcut2));
                float2 ri_ = rsqrt2<NR>(r2c_);
                if (h0 && r2.x <= rcut2) ri_.x = rinv_cut;
                if (h1 && r2.y <= rcut2) ri_.y = rinv_cut;
                const float2 pij_  = __fmul2_rn(make_float2(B.z, B.w), ri_);
                const float2 ri2_  = __fmul2_rn(ri_, ri_);
                const float2 mri3_ = __fmul2_rn(pij_, ri2_);
                ax = __ffma2_rn(mri3_, dx, ax);
                ay = __ffma2_rn(mri3_, dy, ay);
                az = __ffma2_rn(mri3_, dz, az);
                pt = __fadd2_rn(pt, pij_);
                continue;
            }
        }
        const float2 r2c  = make_float2(fmaxf(r2.x, rcut2), fmaxf(r2.y, rcut2));
        const float2 ri   = rsqrt2<NR>(r2c);
        const float2 pij  = __fmul2_rn(make_float2(B.z, B.w), ri);
        const float2 ri2  = __fmul2_rn(ri, ri);
        const float2 mri3 = __fmul2_rn(pij, ri2);
        ax = __ffma2_rn(mri3, dx, ax);
        ay = __ffma2_rn(mri3, dy, ay);
        az = __ffma2_rn(mri3, dz, az);
        pt = __fadd2_rn(pt, pij);
    }
}

// Neighbour search only (the kernel behind PeTar's second tree, tree_nb —
// SearchNeighborEpEpNoSimd, reference src/soft_force.hpp:11-34): n += (r2 < max(rs_i, rs_j)^2),
// exact two-float dx.  Only called for near segments; far segments cannot hold a neighbour.
template <bool EMIT>
__device__ __forceinline__ void ep_count_pairs(const EpTile& t, int p0, int p1,
                                               float xi, float yi, float zi, float xil, float yil, float zil, float rsi2,
                                               float eps2, float2& cf,
                                               const int* jid, unsigned int i_global, const Params& prm) {
    const float2 nxi = bc(-xi), nyi = bc(-yi), nzi = bc(-zi), e2 = bc(eps2);
    const float2 nxil = bc(-xil), nyil = bc(-yil), nzil = bc(-zil);
#pragma unroll kPairUnroll
    for (int p = p0; p < p1; ++p) {
        const float4 A = t.a[p], AL = t.al[p];
        const float4 B = t.b[p];
        const float2 BL = t.bl[p], C = t.c[p];
        const float2 dx = __fadd2_rn(__fadd2_rn(make_float2(A.x, A.y), nxi), __fadd2_rn(make_float2(AL.x, AL.y), nxil));
        const float2 dy = __fadd2_rn(__fadd2_rn(make_float2(A.z, A.w), nyi), __fadd2_rn(make_float2(AL.z, AL.w), nyil));
        const float2 dz = __fadd2_rn(__fadd2_rn(make_float2(B.x, B.y), nzi), __fadd2_rn(BL, nzil));
        float2 r2 = __ffma2_rn(dx, dx, e2);
        r2 = __ffma2_rn(dy, dy, r2);
        r2 = __ffma2_rn(dz, dz, r2);
        const bool h0 = r2.x < fmaxf(C.x, rsi2), h1 = r2.y < fmaxf(C.y, rsi2);
        cf = __fadd2_rn(cf, make_float2(h0 ? 1.f : 0.f, h1 ? 1.f : 0.f));
        if (EMIT && (h0 || h1) && i_global != 0xffffffffu) {
            // rare (well under 1 % of the tested pairs): one slot per hit from the launch-wide cursor.  Per lane, not
            // per warp: lanes that share a particle run different trip counts, so the warp is not converged here
            const unsigned int n = (unsigned int)h0 + (unsigned int)h1;
            const unsigned int at = atomicAdd(prm.pair_cursor, n);
            const unsigned long long hi = (unsigned long long)i_global << 32;
            if (h0 && at < prm.pair_cap) prm.pairs[at] = hi | (unsigned int)jid[2 * p];
            if (h1 && at + (unsigned int)h0 < prm.pair_cap) prm.pairs[at + (unsigned int)h0] = hi | (unsigned int)jid[2 * p + 1];
        }
    }
}

// ------------------------------------------------------------------------------------------
// EP-SP: monopole + quadrupole of a superparticle.
// The reference evaluates, with Q the RAW second-moment tensor and tr its trace
// (src/force_gpu_cuda.cu:283-308, src/soft_force.hpp:175-194),
//   dx = xi - xj ; r2 = eps2 + dx.dx ; qr = Q dx ; qrr = dx.qr
//   A = m r^-3 - 1.5 tr r^-5 + 7.5 qrr r^-7 ; B = -3 r^-5
//   acc -= A dx + B qr ; pot -= m r^-1 - 0.5 tr r^-3 + 1.5 qrr r^-5 .
// With the traceless tensor Q' = 3Q - tr I (formed in fp64 on the host) and
// S = dx.Q'dx - eps2 tr  this is, term for term and for any eps2, the same as
//   acc -= (m r^-3 + 2.5 S r^-7) dx - r^-5 Q'dx ; pot -= m r^-1 + 0.5 S r^-5
// (3 qrr - tr r2 = S because r2 = dx.dx + eps2) — 33 instead of 38 packed FP operations per
// pair of interactions.  The reference's own SIMD path uses the same traceless form
// (src/phantomquad_for_p3t_x86.hpp:144-163).
// ------------------------------------------------------------------------------------------
struct SpRegs { float4 a, b, c, d; };  // {xh,yh,zh,m}, {xl,yl,zl,q'xx}, {q'yy,q'zz,q'xy,q'xz}, {q'yz,tr,-,-}

__device__ __forceinline__ SpRegs sp_load_j(const float4* __restrict__ spj, int id) {
    SpRegs r;
    if (id >= 0) {
        const float4* p = spj + 4 * (size_t)id;
        r.a = __ldg(p); r.b = __ldg(p + 1); r.c = __ldg(p + 2); r.d = __ldg(p + 3);
    } else {
        r.a = r.b = r.c = r.d = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return r;
}
__device__ __forceinline__ void sp_store(SpTile& t, int tid, int id, const SpRegs& r, const Walk& w, float eps2) {
    float v[12];
    if (id >= 0) {
        v[0] = (r.a.x - w.ohx) + (r.b.x - w.olx);
        v[1] = (r.a.y - w.ohy) + (r.b.y - w.oly);
        v[2] = (r.a.z - w.ohz) + (r.b.z - w.olz);
        v[3] = r.a.w;                                  // m
        v[4] = r.b.w; v[5] = r.c.x; v[6] = r.c.y;      // q'xx q'yy q'zz
        v[7] = r.c.z; v[8] = r.c.w; v[9] = r.d.x;      // q'xy q'xz q'yz
        v[10] = -(eps2 * r.d.y);                       // -eps2 tr
        v[11] = 0.f;
    } else {
        v[0] = v[1] = v[2] = kPadPos;
#pragma unroll
        for (int k = 3; k < 12; ++k) v[k] = 0.f;
    }
    const int p = tid >> 1, s = tid & 1;
    float* q0 = reinterpret_cast<float*>(&t.q0[p]);
    float* q1 = reinterpret_cast<float*>(&t.q1[p]);
    float* q2 = reinterpret_cast<float*>(&t.q2[p]);
    float* q3 = reinterpret_cast<float*>(&t.q3[p]);
    float* q4 = reinterpret_cast<float*>(&t.q4[p]);
    float* q5 = reinterpret_cast<float*>(&t.q5[p]);
    q0[s] = v[0]; q0[2 + s] = v[1];
    q1[s] = v[2]; q1[2 + s] = v[3];
    q2[s] = v[4]; q2[2 + s] = v[5];
    q3[s] = v[6]; q3[2 + s] = v[7];
    q4[s] = v[8]; q4[2 + s] = v[9];
    q5[s] = v[10];
}

template <int NR>
__device__ __forceinline__ void sp_pairs(const SpTile& t, int p0, int p1,
                                         float xi, float yi, float zi, float eps2,
                                         float2& ax, float2& ay, float2& az, float2& pt) {
    const float2 vxi = bc(xi), vyi = bc(yi), vzi = bc(zi), e2 = bc(eps2);
#pragma unroll kSpUnroll
    for (int p = p0; p < p1; ++p) {
        const float4 Q0 = t.q0[p], Q1 = t.q1[p], Q2 = t.q2[p], Q3 = t.q3[p], Q4 = t.q4[p];
        const float2 mtr = t.q5[p];
        const float2 xj = make_float2(Q0.x, Q0.y), yj = make_float2(Q0.z, Q0.w), zj = make_float2(Q1.x, Q1.y);
        const float2 mj = make_float2(Q1.z, Q1.w);
        const float2 qxx = make_float2(Q2.x, Q2.y), qyy = make_float2(Q2.z, Q2.w);
        const float2 qzz = make_float2(Q3.x, Q3.y), qxy = make_float2(Q3.z, Q3.w);
        const float2 qxz = make_float2(Q4.x, Q4.y), qyz = make_float2(Q4.z, Q4.w);
        const float2 dx = __fadd2_rn(vxi, make_float2(-xj.x, -xj.y));
        const float2 dy = __fadd2_rn(vyi, make_float2(-yj.x, -yj.y));
        const float2 dz = __fadd2_rn(vzi, make_float2(-zj.x, -zj.y));
        float2 r2 = __ffma2_rn(dx, dx, e2);
        r2 = __ffma2_rn(dy, dy, r2);
        r2 = __ffma2_rn(dz, dz, r2);
        const float2 rinv = rsqrt2<NR>(r2);
        // column order: three independent operations share dx, then dy, then dz (operand reuse)
        float2 qrx = __fmul2_rn(qxx, dx), qry = __fmul2_rn(qxy, dx), qrz = __fmul2_rn(qxz, dx);
        qrx = __ffma2_rn(qxy, dy, qrx); qry = __ffma2_rn(qyy, dy, qry); qrz = __ffma2_rn(qyz, dy, qrz);
        qrx = __ffma2_rn(qxz, dz, qrx); qry = __ffma2_rn(qyz, dz, qry); qrz = __ffma2_rn(qzz, dz, qrz);
        float2 S = __ffma2_rn(qrx, dx, mtr); S = __ffma2_rn(qry, dy, S); S = __ffma2_rn(qrz, dz, S);
        const float2 rinv2 = __fmul2_rn(rinv, rinv);
        const float2 rinv3 = __fmul2_rn(rinv2, rinv);
        const float2 rinv5 = __fmul2_rn(rinv3, rinv2);
        const float2 mr3   = __fmul2_rn(mj, rinv3);
        const float2 S5    = __fmul2_rn(rinv5, S);
        const float2 S7    = __fmul2_rn(S5, rinv2);
        const float2 A     = __ffma2_rn(bc(2.5f), S7, mr3);
        const float2 nA    = make_float2(-A.x, -A.y);
        // acc -= A dx - r^-5 Q'dx
        ax = __ffma2_rn(nA, dx, ax); ax = __ffma2_rn(rinv5, qrx, ax);
        ay = __ffma2_rn(nA, dy, ay); ay = __ffma2_rn(rinv5, qry, ay);
        az = __ffma2_rn(nA, dz, az); az = __ffma2_rn(rinv5, qrz, az);
        // pot accumulates +(m r^-1 + 0.5 S r^-5); negated at the end
        pt = __ffma2_rn(bc(0.5f), S5, pt);
        pt = __ffma2_rn(mj, rinv, pt);
    }
}

// Two i-particles per lane (pb::force_kernel_ws, SP tasks of groups with at least two i-blocks): the same operations as
// sp_pairs for each of them, interleaved so that every j operand register feeds two consecutive instructions (operand
// reuse) and every shared-memory load four interactions: +6 % on the loop alone (tools/loopbench.cu).
template <int NR>
__device__ __forceinline__ void sp_pairs_2i(const SpTile& t, int p0, int p1, const float (&xi)[2], const float (&yi)[2], const float (&zi)[2],
                                            float eps2, float2 (&ax)[2], float2 (&ay)[2], float2 (&az)[2], float2 (&pt)[2]) {
    const float2 e2 = bc(eps2);
#pragma unroll 1
    for (int p = p0; p < p1; ++p) {
        const float4 Q0 = t.q0[p], Q1 = t.q1[p], Q2 = t.q2[p], Q3 = t.q3[p], Q4 = t.q4[p];
        const float2 mtr = t.q5[p];
        const float2 nxj = make_float2(-Q0.x, -Q0.y), nyj = make_float2(-Q0.z, -Q0.w), nzj = make_float2(-Q1.x, -Q1.y);
        const float2 mj = make_float2(Q1.z, Q1.w);
        const float2 qxx = make_float2(Q2.x, Q2.y), qyy = make_float2(Q2.z, Q2.w);
        const float2 qzz = make_float2(Q3.x, Q3.y), qxy = make_float2(Q3.z, Q3.w);
        const float2 qxz = make_float2(Q4.x, Q4.y), qyz = make_float2(Q4.z, Q4.w);
        float2 dx[2], dy[2], dz[2], r2[2], rinv[2], qrx[2], qry[2], qrz[2], S[2];
#pragma unroll
        for (int s = 0; s < 2; s++) { dx[s] = __fadd2_rn(bc(xi[s]), nxj); }
#pragma unroll
        for (int s = 0; s < 2; s++) { dy[s] = __fadd2_rn(bc(yi[s]), nyj); }
#pragma unroll
        for (int s = 0; s < 2; s++) { dz[s] = __fadd2_rn(bc(zi[s]), nzj); }
#pragma unroll
        for (int s = 0; s < 2; s++) { r2[s] = __ffma2_rn(dx[s], dx[s], e2); r2[s] = __ffma2_rn(dy[s], dy[s], r2[s]); r2[s] = __ffma2_rn(dz[s], dz[s], r2[s]); }
#pragma unroll
        for (int s = 0; s < 2; s++) rinv[s] = rsqrt2<NR>(r2[s]);
#pragma unroll
        for (int s = 0; s < 2; s++) qrx[s] = __fmul2_rn(qxx, dx[s]);
#pragma unroll
        for (int s = 0; s < 2; s++) qry[s] = __fmul2_rn(qxy, dx[s]);
#pragma unroll
        for (int s = 0; s < 2; s++) qrz[s] = __fmul2_rn(qxz, dx[s]);
#pragma unroll
        for (int s = 0; s < 2; s++) qrx[s] = __ffma2_rn(qxy, dy[s], qrx[s]);
#pragma unroll
        for (int s = 0; s < 2; s++) qry[s] = __ffma2_rn(qyy, dy[s], qry[s]);
#pragma unroll
        for (int s = 0; s < 2; s++) qrz[s] = __ffma2_rn(qyz, dy[s], qrz[s]);
#pragma unroll
        for (int s = 0; s < 2; s++) qrx[s] = __ffma2_rn(qxz, dz[s], qrx[s]);
#pragma unroll
        for (int s = 0; s < 2; s++) qry[s] = __ffma2_rn(qyz, dz[s], qry[s]);
#pragma unroll
        for (int s = 0; s < 2; s++) qrz[s] = __ffma2_rn(qzz, dz[s], qrz[s]);
#pragma unroll
        for (int s = 0; s < 2; s++) { S[s] = __ffma2_rn(qrx[s], dx[s], mtr); S[s] = __ffma2_rn(qry[s], dy[s], S[s]); S[s] = __ffma2_rn(qrz[s], dz[s], S[s]); }
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const float2 rinv2 = __fmul2_rn(rinv[s], rinv[s]);
            const float2 rinv3 = __fmul2_rn(rinv2, rinv[s]);
            const float2 rinv5 = __fmul2_rn(rinv3, rinv2);
            const float2 mr3   = __fmul2_rn(mj, rinv3);
            const float2 S5    = __fmul2_rn(rinv5, S[s]);
            const float2 S7    = __fmul2_rn(S5, rinv2);
            const float2 A     = __ffma2_rn(bc(2.5f), S7, mr3);
            const float2 nA    = make_float2(-A.x, -A.y);
            ax[s] = __ffma2_rn(nA, dx[s], ax[s]); ay[s] = __ffma2_rn(nA, dy[s], ay[s]); az[s] = __ffma2_rn(nA, dz[s], az[s]);
            ax[s] = __ffma2_rn(rinv5, qrx[s], ax[s]); ay[s] = __ffma2_rn(rinv5, qry[s], ay[s]); az[s] = __ffma2_rn(rinv5, qrz[s], az[s]);
            pt[s] = __ffma2_rn(bc(0.5f), S5, pt[s]);
            pt[s] = __ffma2_rn(mj, rinv[s], pt[s]);
        }
    }
}

// Fused reduction.  The warp that has just written the partial sums of an i-block for one chunk counts the chunk in (one
// release-acquire atomic by lane 0).  The warp that delivered the block's LAST chunk adds all of them in chunk order (fixed
// order: bitwise reproducible whoever comes last), applies G and writes the 32 ForceSoft-shaped records — into page-locked
// host memory, so no reduction kernel and no D2H copy follow the force kernel.  The 40-byte records leave as five coalesced
// 256-byte stores.  reduce_block is a real call: its registers must not weigh on the pair loops.
static __device__ __noinline__ void reduce_block(int bi, int lane, const double4* part4, const int* partn, const IBlock* iblocks, ForceOut* out, int* done, double G) {
    const IBlock B = iblocks[bi];
    double ax = 0.0, ay = 0.0, az = 0.0, pt = 0.0;
    long long n = 0;
#pragma unroll 4
    for (int c = 0; c < B.n_chunks; ++c) {
        const int slot = B.part_base + c * B.stride + lane;
        const double2* q = reinterpret_cast<const double2*>(part4 + slot);
        const double2 v0 = __ldcg(q), v1 = __ldcg(q + 1);
        ax += v0.x; ay += v0.y; az += v1.x; pt += v1.y;
        n += __ldcg(partn + slot);
    }
    const double f[5] = {G * ax, G * ay, G * az, -(G * pt), __longlong_as_double(n)};
    double* o = reinterpret_cast<double*>(out + B.out_off);
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int k = lane + 32 * j, src = k / 5, fld = k - 5 * src;   // 8-byte word k of the block's 160: particle src, field fld
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const double t = __shfl_sync(0xffffffffu, f[q], src);
            if (q == fld) v = t;
        }
        if (src < B.n_valid) o[k] = v;
    }
    if (lane == 0) done[bi] = 0;                              // ready for the next step
}

__device__ __forceinline__ void finish_block(int bi, int nch, int lane, const double4* part4, const int* partn, const Params& prm) {
    __syncwarp();                                             // the lanes' partial-sum stores are ordered before lane 0's release
    int old = 0;
    if (lane == 0) asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(prm.done + bi) : "memory");
    const int last = __shfl_sync(0xffffffffu, (int)(old + 1 == nch), 0);      // ... and lane 0's acquire before the other lanes' loads
    if (last) reduce_block(bi, lane, part4, partn, prm.iblocks, prm.out, prm.done, prm.G);
}

} // namespace pb
