// Exhaustive check that MUFU.RCP + one Newton step (resolve.cuh::r_rcp_rn) equals the IEEE reciprocal for every
// binary32 mantissa (exponents 2^0, 2^-60, 2^60 and both signs): nvcc -arch=sm_100a tools/check_rcp.cu && ./a.out
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ float r_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float r_rcp_rn(float x) { const float r = r_rcp(x); return __fmaf_rn(r, __fmaf_rn(-x, r, 1.0f), r); }
__global__ void check(uint32_t expBits, unsigned long long* bad) {
    uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (1u << 23)) return;
    float x = __uint_as_float(expBits | m);
    if (__float_as_uint(r_rcp_rn(x)) != __float_as_uint(__frcp_rn(x))) atomicAdd(bad, 1ull);
}
int main() {
    unsigned long long* bad; cudaMallocManaged(&bad, 8);
    const uint32_t exps[] = { 127u << 23, 67u << 23, 187u << 23, (127u << 23) | 0x80000000u, 1u << 23, 253u << 23 };
    for (uint32_t e : exps) {
        *bad = 0;
        check<<<(1 << 23) / 256, 256>>>(e, bad);
        cudaDeviceSynchronize();
        printf("exp %08x mismatches %llu\n", e, *bad);
    }
    return 0;
}
