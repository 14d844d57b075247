// Micro-benchmark of the instruction pipes the fp32-screened sampler leans on (sm_100a): warp-instructions per clock per
// SM for DFMA, FFMA, FFMA2 (packed fp32), F2F fp64->fp32, REDUX (integer warp reduction), SHFL and MUFU.EX2, each as
// 8 independent dependency chains per thread with 32 warps per SM resident.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu && ./ubench_pipes
#include <cstdio>
#include <cuda_runtime.h>

enum Op { DFMA, FFMA, FFMA2, F2F_D2F, REDUX_SUM, REDUX_MAX, SHFL, MUFU, FMNMX3 };

template<int OP> __global__ void k(float* out, int iters, float seed) {
    double d[8]; float f[8]; float2 g[8]; int q[8];
    #pragma unroll
    for(int i = 0; i < 8; i++) { d[i] = threadIdx.x * 1e-3 + i + seed; f[i] = (float)d[i]; g[i] = make_float2(f[i], f[i] + 1.f); q[i] = (int)threadIdx.x + i; }
    const double da = 1.0000001, db = 1e-9;
    const float fa = 1.0000001f, fb = 1e-9f;
    for(int it = 0; it < iters; it++) {
        #pragma unroll
        for(int u = 0; u < 8; u++) {
            if(OP == DFMA) d[u] = fma(d[u], da, db);
            if(OP == FFMA) f[u] = fmaf(f[u], fa, fb);
            if(OP == FFMA2) g[u] = __ffma2_rn(g[u], make_float2(fa, fa), make_float2(fb, fb));
            if(OP == F2F_D2F) { f[u] = __double2float_rn(d[u]); d[u] += (double)0.0 + __longlong_as_double((long long)__float_as_int(f[u])); }
            if(OP == REDUX_SUM) q[u] = __reduce_add_sync(0xffffffffu, q[u]) + (int)threadIdx.x;
            if(OP == REDUX_MAX) q[u] = (int)__reduce_max_sync(0xffffffffu, (unsigned)q[u]) ^ (int)threadIdx.x;
            if(OP == SHFL) q[u] = __shfl_xor_sync(0xffffffffu, q[u], 1) + 1;
            if(OP == MUFU) f[u] = exp2f(f[u]) * 1e-3f;
            if(OP == FMNMX3) f[u] = fmaxf(fmaxf(f[u], f[(u + 1) & 7]), fb * it);
        }
    }
    float s = 0.f;
    #pragma unroll
    for(int i = 0; i < 8; i++) s += (float)d[i] + f[i] + g[i].x + g[i].y + (float)q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int OP> void run(const char* name, double instr_per_iter_per_chain) {
    int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int blocks = sms * 4, threads = 256, iters = 2048;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for(int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); k<OP><<<blocks, threads>>>(out, iters, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if(rep && ms < best) best = ms;
    }
    const double winstr = instr_per_iter_per_chain * 8.0 * iters * (double)blocks * (threads / 32);
    const double clocks = best * 1e-3 * khz * 1e3;
    printf("%-10s %8.3f ms  %6.3f warp-instr/clk/SM  (%s)\n", name, best, winstr / clocks / sms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    run<DFMA>("DFMA", 1); run<FFMA>("FFMA", 1); run<FFMA2>("FFMA2", 1); run<F2F_D2F>("F2F.d2f", 1);
    run<REDUX_SUM>("REDUX.SUM", 1); run<REDUX_MAX>("REDUX.MAX", 1); run<SHFL>("SHFL", 1); run<MUFU>("MUFU.EX2", 1); run<FMNMX3>("FMNMX3", 1);
    return 0;
}
