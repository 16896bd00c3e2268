// pb_kernels.cu — hand-written sm_100a kernels of the soft-force hot path.
//
// What they compute is what the reference's two kernels compute
//   force_kernel_ep_ep / dev_gravity_ep_ep   reference src/force_gpu_cuda.cu:222-273 / :50-77
//   force_kernel_ep_sp / dev_gravity_ep_sp   reference src/force_gpu_cuda.cu:465-511 / :276-323
// (and, in fp64, the oracle functors src/soft_force.hpp:38-87, 160-200); how they compute it is
// new:
//
//  * one launch for both interaction kinds; a CTA works on a Task = (i-group of a walk) x
//    (chunk of that walk's EP or SP list); the host sizes tasks so that every CTA runs about the
//    same number of inner-loop steps and there are several waves of CTAs on 148 SMs;
//  * i-particles live in registers (one per lane), relative to the walk's origin;
//  * j tiles (256 entries) are gathered by index with coalesced index reads and 16-byte loads
//    from the L2-resident j store, shifted to the walk origin from their hi/lo fp32 split, and
//    written to shared memory as PAIRS so the inner loop runs on packed fp32x2 instructions
//    (FADD2 / FMUL2 / FFMA2, Blackwell only) — two interactions per issued FP instruction;
//    the gather of tile k+1 and the index read of tile k+2 are in flight during tile k;
//  * reciprocal square roots come from MUFU.RSQ (optionally refined by one Newton step);
//  * forces/potentials are accumulated in fp32 per tile and folded across tiles with a
//    compensated (Kahan-Babuska) sum; per-task results leave the SM as exact hi+lo doubles and a
//    second tiny kernel adds the tasks' partials in fp64 in a fixed order (deterministic, no
//    atomics), applies G and writes ForceSoft-shaped records.
//  * no tensor cores: this is not a contraction.
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

// ------------------------------------------------------------------------------------------
// Index tiles through TMA: a chunk's index list is contiguous and 16-byte aligned, so each 1 KB
// tile of it is one `cp.async.bulk` (1-D TMA bulk copy, SASS UBLKCP) into a 3-deep shared-memory
// ring, completion signalled on an mbarrier — issued by one thread three tiles ahead, no
// registers held, no per-thread global loads.  (The j records themselves are an indexed gather
// and stay 16-byte loads from L2.)  PB_TMA_IDS=0 builds the plain-load variant.
// ------------------------------------------------------------------------------------------
#ifndef PB_TMA_IDS
#define PB_TMA_IDS 1
#endif

constexpr int kIdRing = 3;      // ring slots
constexpr int kIdDirect = 2;    // tiles 0,1 of a chunk are read directly

struct IdRing {
    alignas(16) int buf[kIdRing][kTileJ];
    alignas(8) unsigned long long bar[kIdRing];
};

struct IdPipe {
    const int* ids;      // nullptr: dense list (element j is store slot dbase + j)
    int j_count, n_tiles, dbase;
    IdRing* ring;

    // ring entry r holds tile r + kIdDirect (the first kIdDirect tiles are plain loads: nothing to wait for
    // in the prologue, and the barrier that publishes the mbarrier init is the one the loop needs anyway)
    __device__ __forceinline__ void issue(int r) const {               // one thread
        const int tile = r + kIdDirect;
        const unsigned bytes = (unsigned)(((min(kTileJ, j_count - tile * kTileJ) + 3) & ~3) * 4);
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&ring->bar[r % kIdRing]);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&ring->buf[r % kIdRing][0]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(ids + (size_t)tile * kTileJ), "r"(bytes), "r"(bar) : "memory");
    }
    // call by all threads; a __syncthreads() must separate it from the first get(tile >= kIdDirect)
    __device__ __forceinline__ void init(const int* ids_, int j_count_, int n_tiles_, int dbase_, IdRing* ring_, int tid) {
        ids = ids_; j_count = j_count_; n_tiles = n_tiles_; dbase = dbase_; ring = ring_;
#if PB_TMA_IDS
        if (tid == 0) {
            for (int b = 0; b < kIdRing; ++b) {
                const unsigned bar = (unsigned)__cvta_generic_to_shared(&ring->bar[b]);
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(1) : "memory");
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            if (ids) for (int r = 0; r < min(kIdRing, n_tiles - kIdDirect); ++r) issue(r);
        }
#endif
    }
    // id of list element tile*kTileJ + tid, or -1 past the end of the chunk
    __device__ __forceinline__ int get(int tile, int tid) const {
        const int j = tile * kTileJ + tid;
        if (tile >= n_tiles) return -1;
        if (!ids) return j < j_count ? dbase + j : -1;
#if PB_TMA_IDS
        if (tile < kIdDirect) return j < j_count ? __ldg(ids + j) : -1;
        const int r = tile - kIdDirect;
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&ring->bar[r % kIdRing]);
        const unsigned parity = (unsigned)((r / kIdRing) & 1);
        asm volatile("{\n"
                     ".reg .pred P1;\n"
                     "LAB_WAIT:\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                     "@P1 bra DONE;\n"
                     "bra LAB_WAIT;\n"
                     "DONE:\n"
                     "}" :: "r"(bar), "r"(parity) : "memory");
        return j < j_count ? ring->buf[r % kIdRing][tid] : -1;
#else
        return j < j_count ? __ldg(ids + j) : -1;
#endif
    }
    // once every thread has read tiles <= k+2 = ring entries <= k (the caller knows: all of them have arrived on the
    // mbarrier of tile k+1, which they do after that read), the slot of entry k is free for entry k + kIdRing
    __device__ __forceinline__ void refill(int k, int tid) const {
#if PB_TMA_IDS
        if (ids && tid == 0 && k + kIdRing + kIdDirect < n_tiles) issue(k + kIdRing);
#endif
    }
};

// ------------------------------------------------------------------------------------------
// the force kernel
// ------------------------------------------------------------------------------------------
// PERSIST: the grid is 2 CTAs per SM and every CTA pulls task numbers from an atomic cursor until the task list is
// used up — the task count lives in device memory (meta[0], cursor meta[4]) because the plan was made on the device
// (pb_plan.cu), so no host round trip sits between the tree walk and the forces.
template <int NR, int MINB, bool EMIT = false, bool PERSIST = false>
__global__ void __launch_bounds__(kThreads, MINB)
force_kernel(const Walk* __restrict__ walks, const Task* __restrict__ tasks,
             const float4* __restrict__ epi,
             const int* __restrict__ id_epj, const int* __restrict__ id_spj,
             const float4* __restrict__ epj, const float4* __restrict__ spj,
             double4* __restrict__ part4, int* __restrict__ partn, Params prm)
{
    __shared__ Smem sm;
    __shared__ int near_flag[kTileBufs][kWarpsPerCta];   // per tile buffer, per staging warp (= 16-pair segment)
    __shared__ IdRing ring;                      // index tiles, filled by TMA bulk copies
    __shared__ TileBars bars;                    // tile hand-over (see TileBars)
    __shared__ int jid[EMIT ? kTileBufs : 1][EMIT ? kTileJ : 1];   // store index of every staged j (neighbour-list emission only)
    __shared__ int s_task;

    const int tid   = threadIdx.x;
    const int warp  = tid >> 5, lane = tid & 31;
  for (;;) {
    int task_id = blockIdx.x;
    if (PERSIST) {
        if (tid == 0) s_task = atomicAdd(prm.meta + 4, 1);
        __syncthreads();
        task_id = s_task;
        if (task_id >= prm.meta[0]) break;
    }
    const Task task = tasks[task_id];
    const Walk w    = walks[task.walk];

    // warp role: i-block `ib` of the group, j-split slot `js`
    const bool busy = warp < task.nib * task.jsplit;
    const int  ib   = busy ? warp % task.nib : 0;
    const int  js   = busy ? warp / task.nib : 0;
    const int  ppw  = kTilePairs / task.jsplit;             // pairs of each tile this warp consumes

    // A block with at most 16 (8) real i-particles — the ragged end of a walk — lets 2 (4) lanes share each particle
    // and split every j segment between them instead of idling: lane = (sub, il), il = the particle, sub = the lane's
    // share of the pairs; the shares are added by shuffles after the loops.
    const int  n_blk  = busy ? min(32, w.ni - (task.i_first + ib * 32)) : 32;
    const int  ishift = (n_blk <= 8) ? 2 : (n_blk <= 16) ? 1 : 0;     // isplit = 1 << ishift lanes per particle
    const int  lanes_i = 32 >> ishift;
    const int  il = lane & (lanes_i - 1), sub = lane >> (5 - ishift);

    // i-particle (registers): relative to the walk origin already (host formed x_i - origin in fp64)
    const int  i_loc  = task.i_first + ib * 32 + il;
    const bool ivalid = busy && (i_loc < w.ni);
    float4 pi = make_float4(0.f, 0.f, 0.f, 0.f), pil = make_float4(0.f, 0.f, 0.f, 0.f), pih = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ivalid) {
        const float4* rec = epi + (size_t)prm.i_f4 * (size_t)(w.i_off + i_loc);
        pi  = __ldg(rec);                                        // {x, y, z, r_search}: relative position, hi part
        pil = __ldg(rec + 1);                                    // {xl, yl, zl, -}: lo part
        if (prm.abs_mode == 2) pih = __ldg(rec + 2);             // {float(x), float(y), float(z), -}: absolute, as the CPU replay casts it
    }
    const float rsi2 = ivalid ? pi.w * pi.w : -1.f;

    KSum kx, ky, kz, kp;
    kx.init(); ky.init(); kz.init(); kp.init();
    int cnt = 0;

    const int n_tiles = (task.j_count + kTileJ - 1) / kTileJ;

    // Tile pipeline (both kinds): index tiles arrive by TMA up to four tiles ahead, ids are picked up two tiles ahead,
    // the gathered j records one tile ahead; tile k+1 is stored (and its mbarrier arrived on) after tile k was
    // computed, and waited for only at the start of iteration k+1.
    bars.init(tid);
    if (task.kind == 0 || task.kind == 2) {
        const bool count_only = (task.kind == 2);     // neighbour search only (tree_nb): no force math
        const int* ids = (w.ej_off >= 0) ? id_epj + w.ej_off + task.j_begin : nullptr;   // ej_off < 0: dense list
        const int dbase = task.j_begin;
        IdPipe idp;
        idp.init(ids, task.j_count, n_tiles, dbase, &ring, tid);
        int id_cur = idp.get(0, tid);
        EpRegs jr  = ep_load_j(epj, id_cur);
        int id_nxt = idp.get(1, tid);
        __syncthreads();                               // mbarrier inits (tile ring, index ring) visible to every thread
        {
            const bool nr_ = ep_store(sm.ep[0], tid, id_cur, jr, w, prm.abs_mode);
            const unsigned bal = __ballot_sync(0xffffffffu, nr_);
            if (lane == 0) near_flag[0][warp] = (bal != 0u);
            if (EMIT) jid[0][tid] = id_cur;
            bars.arrive(0);
        }
        int cb = 0; unsigned cpar = 0;                 // buffer and mbarrier phase parity of tile k
        for (int k = 0; k < n_tiles; ++k) {
            const bool more = (k + 1 < n_tiles);
            const int nb = (cb + 1 == kTileBufs) ? 0 : cb + 1;
            if (more) jr = ep_load_j(epj, id_nxt);
            const int id_nn = idp.get(k + 2, tid);
            bars.wait(cb, cpar);                       // every thread has stored its j of tile k
            if (k >= 1) idp.refill(k - 1, tid);        // ... and had read the ids of tile k+1 before that
            if (busy) {
                const EpTile& T = sm.ep[cb];
                const int nv  = min(kTileJ, task.j_count - k * kTileJ);
                const int npu = (((nv + 1) >> 1) + kPairUnroll - 1) & ~(kPairUnroll - 1);
                // a ragged (last) tile is shared out evenly, in whole 16-pair segments, between the warps of an i-block
                const int ppk = (nv == kTileJ) ? ppw : max(16, (((npu + task.jsplit - 1) / task.jsplit) + 15) & ~15);
                const int p0  = js * ppk, p1 = min(p0 + ppk, npu);
                float2 ax = bc(0.f), ay = bc(0.f), az = bc(0.f), pt = bc(0.f), cf = bc(0.f);
                // 16-pair segments = the 32 j one staging warp wrote; count only where flagged
                for (int seg0 = p0; seg0 < p1; seg0 += 16) {
                    // this lane's contiguous share of the segment (all of it unless lanes share a particle)
                    const int seg = seg0 + sub * (16 >> ishift);
                    const int e = min(min(seg0 + 16, p1), seg + (16 >> ishift));
                    if (count_only) {
                        if (near_flag[cb][seg0 >> 4])
                            ep_count_pairs<EMIT>(T, seg, e, pi.x, pi.y, pi.z, pil.x, pil.y, pil.z, rsi2, prm.eps2, cf,
                                                 jid[EMIT ? cb : 0], ivalid ? (unsigned int)(prm.i_base + w.i_off + i_loc) : 0xffffffffu, prm);
                    } else if (near_flag[cb][seg0 >> 4]) {
                        if (prm.abs_mode == 2)
                            ep_pairs<NR, 2>(T, seg, e, pi.x, pi.y, pi.z, pil.x, pil.y, pil.z, rsi2, pih.x, pih.y, pih.z, prm.eps2, prm.rcut2, prm.rinv_cut, ax, ay, az, pt, cf);
                        else
                            ep_pairs<NR, 1>(T, seg, e, pi.x, pi.y, pi.z, pil.x, pil.y, pil.z, rsi2, 0.f, 0.f, 0.f, prm.eps2, prm.rcut2, 0.f, ax, ay, az, pt, cf);
                    } else
                        ep_pairs<NR, 0>(T, seg, e, pi.x, pi.y, pi.z, pil.x, pil.y, pil.z, rsi2, 0.f, 0.f, 0.f, prm.eps2, prm.rcut2, 0.f, ax, ay, az, pt, cf);
                }
                kx.add(ax.x + ax.y); ky.add(ay.x + ay.y); kz.add(az.x + az.y); kp.add(pt.x + pt.y);
                cnt += (int)(cf.x + cf.y);
            }
            if (more) {
                const bool nr_ = ep_store(sm.ep[nb], tid, id_nxt, jr, w, prm.abs_mode);
                const unsigned bal = __ballot_sync(0xffffffffu, nr_);
                if (lane == 0) near_flag[nb][warp] = (bal != 0u);
                if (EMIT) jid[nb][tid] = id_nxt;
                bars.arrive(nb);
            }
            id_nxt = id_nn;
            cb = nb; if (nb == 0) cpar ^= 1u;
        }
    } else {
        const int* ids = id_spj + w.sj_off + task.j_begin;
        IdPipe idp;
        idp.init(ids, task.j_count, n_tiles, 0, &ring, tid);
        int id_cur = idp.get(0, tid);
        SpRegs jr  = sp_load_j(spj, id_cur);
        int id_nxt = idp.get(1, tid);
        __syncthreads();                               // mbarrier inits visible
        sp_store(sm.sp[0], tid, id_cur, jr, w, prm.eps2);
        bars.arrive(0);
        int cb = 0; unsigned cpar = 0;
        for (int k = 0; k < n_tiles; ++k) {
            const bool more = (k + 1 < n_tiles);
            const int nb = (cb + 1 == kTileBufs) ? 0 : cb + 1;
            if (more) jr = sp_load_j(spj, id_nxt);
            const int id_nn = idp.get(k + 2, tid);
            bars.wait(cb, cpar);
            if (k >= 1) idp.refill(k - 1, tid);
            if (busy) {
                const int nv  = min(kTileJ, task.j_count - k * kTileJ);
                const int npu = ((nv + 1) >> 1);          // pairs with at least one real j (padding pairs are inert anyway)
                const int ppk = (nv == kTileJ) ? ppw : (npu + task.jsplit - 1) / task.jsplit;   // ragged tile: even shares
                const int p0  = js * ppk, p1 = min(p0 + ppk, npu);
                float2 ax = bc(0.f), ay = bc(0.f), az = bc(0.f), pt = bc(0.f);
                const int plen = (max(p1 - p0, 0) + (1 << ishift) - 1) >> ishift;   // this lane's contiguous share of the warp's pairs
                const int q0 = p0 + sub * plen, q1 = min(p1, q0 + plen);
                sp_pairs<NR>(sm.sp[cb], q0, q1, pi.x, pi.y, pi.z, prm.eps2, ax, ay, az, pt);
                kx.add(ax.x + ax.y); ky.add(ay.x + ay.y); kz.add(az.x + az.y); kp.add(pt.x + pt.y);
            }
            if (more) { sp_store(sm.sp[nb], tid, id_nxt, jr, w, prm.eps2); bars.arrive(nb); }
            id_nxt = id_nn;
            cb = nb; if (nb == 0) cpar ^= 1u;
        }
    }
    __syncthreads();       // every warp is done with the tile buffers (they are reused for the cross-warp combine below)

    // per-warp totals as exact doubles (hi + lo)
    double dax = kx.value(), day = ky.value(), daz = kz.value(), dpt = kp.value();
    for (int o = lanes_i; o < 32; o <<= 1) {              // lanes that shared a particle (isplit > 1): fixed-order butterfly
        dax += __shfl_xor_sync(0xffffffffu, dax, o); day += __shfl_xor_sync(0xffffffffu, day, o);
        daz += __shfl_xor_sync(0xffffffffu, daz, o); dpt += __shfl_xor_sync(0xffffffffu, dpt, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }

    if (task.jsplit > 1) {
        // combine the jsplit warps that share an i-block, in fixed order js = 0,1,...
        if (busy) {
            sm.red[warp][0][lane] = dax; sm.red[warp][1][lane] = day;
            sm.red[warp][2][lane] = daz; sm.red[warp][3][lane] = dpt;
        }
        // neighbour counts ride in a second pass to keep the scratch small
        __syncthreads();
        if (busy && js == 0) {
            for (int s = 1; s < task.jsplit; ++s) {
                const int ww = s * task.nib + ib;
                dax += sm.red[ww][0][lane]; day += sm.red[ww][1][lane];
                daz += sm.red[ww][2][lane]; dpt += sm.red[ww][3][lane];
            }
        }
        __syncthreads();
        int* redn = reinterpret_cast<int*>(&sm.red[0][0][0]);
        if (busy) redn[warp * 32 + lane] = cnt;
        __syncthreads();
        if (busy && js == 0)
            for (int s = 1; s < task.jsplit; ++s) cnt += redn[(s * task.nib + ib) * 32 + lane];
    }

    if (busy && js == 0 && sub == 0) {
        const int slot = task.part_base + ib * 32 + il;
        part4[slot] = make_double4(dax, day, daz, dpt);
        partn[slot] = cnt;
    }
    if (!PERSIST) break;
    __syncthreads();                                   // everybody is done with this task's shared state
    if (tid == 0) {                                    // the mbarriers are re-initialised by the next task
        for (int b = 0; b < kTileBufs; ++b) {
            const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars.full[b]);
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(bar) : "memory");
        }
        for (int b = 0; b < kIdRing; ++b) {
            const unsigned bar = (unsigned)__cvta_generic_to_shared(&ring.bar[b]);
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(bar) : "memory");
        }
    }
  }
}

cudaError_t launch_force(cudaStream_t s, int n_tasks, int nr_steps, int min_blocks,
                         const Walk* walks, const Task* tasks,
                         const float4* epi, const int* id_epj, const int* id_spj,
                         const float4* epj, const float4* spj,
                         double4* part4, int* partn, Params p, bool emit_pairs)
{
    if (n_tasks <= 0) return cudaSuccess;
#define PB_LAUNCH(...) force_kernel<__VA_ARGS__><<<n_tasks, kThreads, 0, s>>>(walks, tasks, epi, id_epj, id_spj, epj, spj, part4, partn, p)
    if (emit_pairs)          PB_LAUNCH(0, 2, true);        // neighbour lists: count-only tasks, no rsqrt involved
    else if (nr_steps >= 1) { if (min_blocks >= 3) PB_LAUNCH(1, 3); else PB_LAUNCH(1, 2); }
    else                    { if (min_blocks >= 3) PB_LAUNCH(0, 3); else PB_LAUNCH(0, 2); }
#undef PB_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_force_persistent(cudaStream_t s, int n_ctas, int nr_steps,
                                    const Walk* walks, const Task* tasks, const float4* epi, const int* id_epj, const int* id_spj,
                                    const float4* epj, const float4* spj, double4* part4, int* partn, Params p)
{
    if (n_ctas <= 0) return cudaSuccess;
    if (nr_steps >= 1) force_kernel<1, 2, false, true><<<n_ctas, kThreads, 0, s>>>(walks, tasks, epi, id_epj, id_spj, epj, spj, part4, partn, p);
    else               force_kernel<0, 2, false, true><<<n_ctas, kThreads, 0, s>>>(walks, tasks, epi, id_epj, id_spj, epj, spj, part4, partn, p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// reduction: one warp per 32-wide i-block adds that block's task partials in fp64, chunk order
// fixed, applies G and writes ForceSoft-shaped records.
//   acc = G * sum ; pot = -G * sum(pot terms)   (signs: see the pair loops above)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
reduce_kernel(int n_iblocks, const IBlock* __restrict__ iblocks,
              const double4* __restrict__ part4, const int* __restrict__ partn,
              ForceOut* __restrict__ out, double G, const int* __restrict__ meta)
{
    const int b    = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= n_iblocks) return;
    if (meta && meta[3] != 0) return;                 // device-made plan did not fit its buffers: nothing was computed
    const IBlock ibk = iblocks[b];
    if (lane >= ibk.n_valid) return;
    double ax = 0.0, ay = 0.0, az = 0.0, pt = 0.0;
    long long n = 0;
#pragma unroll 8
    for (int c = 0; c < ibk.n_chunks; ++c) {          // independent loads: unrolled for memory-level parallelism
        const int slot = ibk.part_base + c * ibk.stride + lane;
        const double4 v = part4[slot];
        ax += v.x; ay += v.y; az += v.z; pt += v.w;
        n += partn[slot];
    }
    ForceOut o;
    o.ax = G * ax; o.ay = G * ay; o.az = G * az; o.pot = -(G * pt); o.n_ngb = n;
    out[ibk.out_off + lane] = o;
}

cudaError_t launch_reduce(cudaStream_t s, int n_iblocks, const IBlock* iblocks,
                          const double4* part4, const int* partn, ForceOut* out, double G, const int* meta)
{
    if (n_iblocks <= 0) return cudaSuccess;
    const int wpb = 8;
    reduce_kernel<<<(n_iblocks + wpb - 1) / wpb, wpb * 32, 0, s>>>(n_iblocks, iblocks, part4, partn, out, G, meta);
    return cudaGetLastError();
}

} // namespace pb
