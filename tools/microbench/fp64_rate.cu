// Microbenchmark: FP64 DFMA / FP32 FFMA issue rate and dependent latency on the device.
#include <cstdio>
#include <cuda_runtime.h>

template <class T, int ILP>
__global__ void fma_kernel(T* out, int iters, T a, T b) {
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = (T)threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = acc[i] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class T, int ILP>
void run(const char* name, int blocks, int threads, int iters) {
    T* out;
    cudaMalloc(&out, sizeof(T) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fma_kernel<T, ILP><<<blocks, threads>>>(out, iters, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e0);
    fma_kernel<T, ILP><<<blocks, threads>>>(out, iters, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * blocks * threads * (double)iters * ILP;
    printf("%s blocks=%d threads=%d ILP=%d: %.3f ms  %.2f TFLOP/s  (%.2f FMA/clk/SM at 1.9 GHz, 148 SM)\n", name, blocks, threads, ILP, ms,
           fl / ms / 1e9, fl / 2 / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}

int main() {
    run<double, 8>("f64", 148 * 4, 512, 20000);
    run<double, 1>("f64 dependent (1 warp/SMSP)", 148, 128, 20000);
    run<float, 8>("f32", 148 * 4, 512, 20000);
    run<float, 1>("f32 dependent (1 warp/SMSP)", 148, 128, 20000);
    return 0;
}
