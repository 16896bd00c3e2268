// microbench.cu — pipe-rate probes for the instructions the force kernel is made of (sm_100a).
// Prints warp-instructions per clock per SM for long, independent dependency chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu && ./microbench
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 4096

template <int KIND>
__global__ void probe(float* out, float seed, long long* cycles) {
    float2 a[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) a[k] = make_float2(seed + k + threadIdx.x, seed - k);
    const float2 b = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-9f, -1e-9f);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < CHAINS; k++) {
            if (KIND == 0) { a[k].x = fmaf(a[k].x, b.x, c.x); }                                   // FFMA
            if (KIND == 1) { a[k] = __ffma2_rn(a[k], b, c); }                                     // FFMA2
            if (KIND == 2) { a[k] = __fadd2_rn(a[k], c); }                                        // FADD2
            if (KIND == 3) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[k].x)); }       // MUFU.RSQ
            if (KIND == 4) { a[k].x = fmaxf(a[k].x, c.x) ; a[k].y = fminf(a[k].y, b.y); }         // 2x FMNMX (alu)
            if (KIND == 5) { a[k] = __ffma2_rn(a[k], b, c); a[k].x = fmaxf(a[k].x, c.y); }         // FFMA2 + FMNMX mix
            if (KIND == 6) { a[k].x = fmaf(a[k].x, b.x, c.x); a[k].y = fmaf(a[k].y, b.y, c.y); }  // 2x FFMA
            if (KIND == 8) { a[k] = __ffma2_rn(a[k], a[k], c); }                                  // FFMA2, 2 distinct operands
            if (KIND == 9) { a[k] = __ffma2_rn(a[k], make_float2(seed, seed), c); }               // FFMA2 with scalar-broadcast operand
            if (KIND == 10) { a[k] = __fmul2_rn(a[k], b); }                                       // FMUL2
            if (KIND == 11) { a[k] = __ffma2_rn(b, c, a[k]); }                                    // FFMA2 accumulate form
            if (KIND == 12) { a[k] = __ffma2_rn(a[k], b, c); a[(k + 1) % CHAINS].x = fmaf(a[(k + 1) % CHAINS].x, b.x, c.x); } // FFMA2 + FFMA
            if (KIND == 7) { a[k] = __ffma2_rn(a[k], b, c); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[k].y)); } // FFMA2+MUFU
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s += a[k].x + a[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int KIND>
void run(const char* name, int inst_per_step, int warps_per_sm) {
    float* out; long long* cyc; long long h = 0;
    const int threads = 32 * warps_per_sm;
    cudaMalloc(&out, sizeof(float) * 148 * threads);
    cudaMalloc(&cyc, sizeof(long long));
    probe<KIND><<<148, threads>>>(out, 1.0f, cyc);
    probe<KIND><<<148, threads>>>(out, 1.0f, cyc);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double winst = (double)ITERS * CHAINS * inst_per_step * warps_per_sm;
    printf("%-28s warps/SM %2d : %.3f warp-inst/clk/SM  (%.2f clk per warp-inst per SMSP)\n", name, warps_per_sm, winst / h, h / (winst / 4));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {8, 16, 32}) {
        run<0>("FFMA", 1, w);
        run<6>("2x FFMA (indep halves)", 2, w);
        run<1>("FFMA2", 1, w);
        run<2>("FADD2", 1, w);
        run<3>("MUFU.RSQ", 1, w);
        run<4>("2x FMNMX", 2, w);
        run<5>("FFMA2 + FMNMX", 2, w);
        run<7>("FFMA2 + MUFU.RSQ", 2, w);
        run<8>("FFMA2 a*a+c", 1, w);
        run<9>("FFMA2 a*bcast(s)+c", 1, w);
        run<10>("FMUL2", 1, w);
        run<11>("FFMA2 b*c+a", 1, w);
        run<12>("FFMA2 + FFMA", 2, w);
        printf("\n");
    }
    return 0;
}
