// pb_pack.cu — device-side packing of raw host particle arrays into the device j formats (option "raw_upload").
//
// With several ranks per node the host cores are the scarce resource (4 per rank on an 8-GPU box): packing this rank's
// own j-particles on the host (fp64 AoS, 120 / 80 B per element, -> hi/lo fp32, 32 / 64 B) cost 2.7 ms of a 9.6 ms tree
// step at 8 ranks, all ranks competing for the same memory bandwidth, while every GPU's PCIe link idled.  With
// raw_upload = 1 the caller's arrays are page-locked once (cudaHostRegister), copied as they are, and packed here.
// Same arithmetic as pack_epj / pack_spj in pb_engine.cu; compiled with -fmad=false so that the results are bit-identical
// to the host packing (3 q - tr must not be contracted into an FMA).
#include "pb_device.h"

namespace pb {

namespace {
__device__ __forceinline__ double ldd(const char* p, size_t off, int k = 0) {
    return *reinterpret_cast<const double*>(p + off + 8 * (size_t)k);
}
__device__ __forceinline__ void split(double x, float& hi, float& lo) {
    hi = (float)x;
    lo = (float)(x - (double)hi);
}
} // namespace

__global__ void __launch_bounds__(256)
pack_epj_kernel(const char* __restrict__ raw, size_t stride, size_t off_pos, size_t off_mass, size_t off_rs, int n, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const char* p = raw + (size_t)i * stride;
    float4 a, b;
    split(ldd(p, off_pos, 0), a.x, b.x);
    split(ldd(p, off_pos, 1), a.y, b.y);
    split(ldd(p, off_pos, 2), a.z, b.z);
    a.w = (float)ldd(p, off_mass);
    b.w = (float)ldd(p, off_rs);
    out[2 * (size_t)i] = a;
    out[2 * (size_t)i + 1] = b;
}

__global__ void __launch_bounds__(256)
pack_spj_kernel(const char* __restrict__ raw, size_t stride, size_t off_pos, size_t off_mass, size_t off_quad, int has_quad, int n, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const char* p = raw + (size_t)i * stride;
    float4 a, b, c, d;
    split(ldd(p, off_pos, 0), a.x, b.x);
    split(ldd(p, off_pos, 1), a.y, b.y);
    split(ldd(p, off_pos, 2), a.z, b.z);
    a.w = (float)ldd(p, off_mass);
    double q[6] = {0, 0, 0, 0, 0, 0};
    if (has_quad)
        for (int k = 0; k < 6; k++) q[k] = ldd(p, off_quad, k);
    const double tr = q[0] + q[1] + q[2];
    b.w = (float)(3.0 * q[0] - tr);
    c = make_float4((float)(3.0 * q[1] - tr), (float)(3.0 * q[2] - tr), (float)(3.0 * q[3]), (float)(3.0 * q[4]));
    d = make_float4((float)(3.0 * q[5]), (float)tr, 0.f, 0.f);
    float4* o = out + 4 * (size_t)i;
    o[0] = a; o[1] = b; o[2] = c; o[3] = d;
}

cudaError_t launch_pack_epj(cudaStream_t s, const void* raw, size_t stride, size_t off_pos, size_t off_mass, size_t off_rs, int n, float4* out) {
    if (n <= 0) return cudaSuccess;
    pack_epj_kernel<<<(n + 255) / 256, 256, 0, s>>>((const char*)raw, stride, off_pos, off_mass, off_rs, n, out);
    return cudaGetLastError();
}

cudaError_t launch_pack_spj(cudaStream_t s, const void* raw, size_t stride, size_t off_pos, size_t off_mass, size_t off_quad, int has_quad, int n, float4* out) {
    if (n <= 0) return cudaSuccess;
    pack_spj_kernel<<<(n + 255) / 256, 256, 0, s>>>((const char*)raw, stride, off_pos, off_mass, off_quad, has_quad, n, out);
    return cudaGetLastError();
}

} // namespace pb
