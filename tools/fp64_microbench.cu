// FP64 pipe microbenchmark for sm_100a: DFMA vs DMMA (mma.sync f64) peak rates.
// Decides whether the score kernel of the backup uses the FMA pipe or the legacy DMMA path.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], double a0, double a1, double b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b0));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = threadIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_dmma1684(double* out, int iters, double a, double b) {
    double c[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma1684(c[i], a, b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b) {
    double c[8][4]; double af[4] = {a, b, a, b}; double bf[2] = {b, a};
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma1688(c[i], af, bf);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a, double b) {
    double c[8][4]; double af[8] = {a, b, a, b, a, b, a, b}; double bf[4] = {b, a, b, a};
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma16816(c[i], af, bf);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: DMMA and DFMA issued together (do they share the pipe?)
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b) {
    double c[8][2]; double acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = threadIdx.x; acc[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) { dmma884(c[i][0], c[i][1], a, b); acc[i] = fma(acc[i], a, b); }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs=%d clock=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
    int nb = p.multiProcessorCount * 8, nt = 256, iters = 20000;
    double* out; CK(cudaMalloc(&out, sizeof(double) * nb * nt));
    double warps = (double)nb * nt / 32.0;
    float ms;
    ms = timeit([&] { k_dfma<<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
    printf("DFMA          : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * nb * nt * 16.0 * iters / ms / 1e9);
    ms = timeit([&] { k_dmma884<<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA m8n8k4   : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * warps * 8 * 256.0 * iters / ms / 1e9);
    ms = timeit([&] { k_dmma1684<<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA m16n8k4  : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * warps * 8 * 512.0 * iters / ms / 1e9);
    ms = timeit([&] { k_dmma1688<<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA m16n8k8  : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * warps * 8 * 1024.0 * iters / ms / 1e9);
    ms = timeit([&] { k_dmma16816<<<nb, nt>>>(out, iters / 2, 1.0000001, 1e-9); });
    printf("DMMA m16n8k16 : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * warps * 8 * 2048.0 * (iters / 2) / ms / 1e9);
    ms = timeit([&] { k_mixed<<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
    printf("mixed 884+DFMA: %8.3f ms  %7.2f TFLOP/s (sum)\n", ms, 2.0 * (warps * 8 * 256.0 + (double)nb * nt * 8.0) * iters / ms / 1e9);
    CK(cudaDeviceSynchronize());
    return 0;
}
