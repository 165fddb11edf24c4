// Throwaway microbenchmark: dependent-chain latency of the fp64 operations on the G-test's critical path.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double *out, long long *cyc, double a, double b) {
    double x = a;
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 1000; ++i) { x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); }
    long long t1 = clock64();
    double y = a;
    #pragma unroll 1
    for (int i = 0; i < 1000; ++i) { y = __ddiv_rn(a, y + b); y = __ddiv_rn(a, y + b); }
    long long t2 = clock64();
    double z = a;
    #pragma unroll 1
    for (int i = 0; i < 1000; ++i) { z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31); z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 3) & 31); }
    long long t3 = clock64();
    double w = a;
    #pragma unroll 1
    for (int i = 0; i < 1000; ++i) { w = __dadd_rn(w, b); w = __dmul_rn(w, b); w = __dadd_rn(w, b); w = __dmul_rn(w, b); }
    long long t4 = clock64();
    if (threadIdx.x == 0) { out[0] = x + y + z + w; cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
}
int main() {
    double *o; long long *c, h[4];
    cudaMalloc(&o, 8); cudaMalloc(&c, 32);
    for (int rep = 0; rep < 2; ++rep) { lat<<<1, 32>>>(o, c, 1.000001, 0.999999); cudaDeviceSynchronize(); }
    cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
    printf("DFMA dependent: %.1f clk   (add+DDIV) dependent: %.1f clk   double SHFL (2 x 32-bit) dependent: %.1f clk   DADD/DMUL dependent: %.1f clk\n",
           h[0] / 4000.0, h[1] / 2000.0, h[2] / 2000.0, h[3] / 4000.0);
    return 0;
}
