#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c){ unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b){ unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b){ unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template<int MODE> __global__ void k(float* out, int iters, float s) {
    float a[8], b[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; b[i] = s + i; }
    if (MODE == 0) {           // 16 scalar FFMA per iter (8 chains x2)
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) { a[i] = __fmaf_rn(a[i], s, b[i]); }
#pragma unroll
            for (int i = 0; i < 8; i++) { b[i] = __fmaf_rn(b[i], s, a[i]); }
        }
    } else if (MODE == 1) {    // 8 FFMA2 per iter = same flops
        unsigned long long A[4], B[4], S = pk(s, s);
        for (int i = 0; i < 4; i++) { A[i] = pk(a[2*i], a[2*i+1]); B[i] = pk(b[2*i], b[2*i+1]); }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 4; i++) A[i] = fma2(A[i], S, B[i]);
#pragma unroll
            for (int i = 0; i < 4; i++) B[i] = fma2(B[i], S, A[i]);
        }
        for (int i = 0; i < 4; i++) { upk(A[i], a[2*i], a[2*i+1]); upk(B[i], b[2*i], b[2*i+1]); }
    } else if (MODE == 2) {    // mixed: 8 FFMA2 + 16 integer LOP/IADD per iter (issue pressure)
        unsigned long long A[4], B[4], S = pk(s, s);
        unsigned x[8];
        for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
        for (int i = 0; i < 4; i++) { A[i] = pk(a[2*i], a[2*i+1]); B[i] = pk(b[2*i], b[2*i+1]); }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 4; i++) A[i] = fma2(A[i], S, B[i]);
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = (x[i] ^ (x[(i+1)&7] >> 3)) + it;
#pragma unroll
            for (int i = 0; i < 4; i++) B[i] = fma2(B[i], S, A[i]);
        }
        for (int i = 0; i < 4; i++) { upk(A[i], a[2*i], a[2*i+1]); upk(B[i], b[2*i], b[2*i+1]); }
        for (int i = 0; i < 8; i++) a[i] += x[i];
    } else {                   // mixed scalar: 16 FFMA + same integer work
        unsigned x[8];
        for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) { a[i] = __fmaf_rn(a[i], s, b[i]); }
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = (x[i] ^ (x[(i+1)&7] >> 3)) + it;
#pragma unroll
            for (int i = 0; i < 8; i++) { b[i] = __fmaf_rn(b[i], s, a[i]); }
        }
        for (int i = 0; i < 8; i++) a[i] += x[i];
    }
    float r = 0; for (int i = 0; i < 8; i++) r += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template<int MODE> void run(const char* name, float* out) {
    int iters = 20000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148*4, 256>>>(out, 100, 1.0001f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148*4, 256>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = 148.0*4*256*iters*16; printf("%s: %.3f ms  %.1f TFLOP/s (fp32 fma)\n", name, ms, 2*fmas/ms/1e9);
}
int main(){ float* out; cudaMalloc(&out, 148*4*256*4); run<0>("scalar FFMA", out); run<1>("FFMA2", out); run<3>("scalar FFMA + int", out); run<2>("FFMA2 + int", out); return 0; }
