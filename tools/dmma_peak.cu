// Micro-benchmark: FP64 tensor-core (mma.sync ... f64) throughput per shape vs plain DFMA on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template<int SHAPE> __global__ void k(double* out, int iters) {
    double c[8][4]; double a[8], b[4];
    for(int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1e-3 + i; for(int j = 0; j < 4; j++) c[i][j] = 0.0; }
    for(int i = 0; i < 4; i++) b[i] = 1e-3 * i + 1e-6 * threadIdx.x;
    for(int it = 0; it < iters; it++) {
        #pragma unroll
        for(int u = 0; u < 8; u++) {
            if(SHAPE == 0) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(a[0]), "d"(b[0]));
            if(SHAPE == 1) asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            if(SHAPE == 2) asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            if(SHAPE == 3) asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};" : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0; for(int i = 0; i < 8; i++) for(int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int SHAPE> void run(const char* name, double flops_per_mma) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 4, threads = 256, iters = 4096;
    double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for(int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); k<SHAPE><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if(rep && ms < best) best = ms;
    }
    const double total = flops_per_mma * 8.0 * iters * (double)blocks * (threads / 32);
    printf("%-12s %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, best, total / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    run<0>("m8n8k4", 2.0 * 8 * 8 * 4);
    run<1>("m16n8k4", 2.0 * 16 * 8 * 4);
    run<2>("m16n8k8", 2.0 * 16 * 8 * 8);
    run<3>("m16n8k16", 2.0 * 16 * 8 * 16);
    return 0;
}
