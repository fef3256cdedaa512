// fp64_peak.cu -- measured fp64 FMA peak of the GPU ("builder must measure an fp64 FMA
// microbenchmark ... before quoting utilisation").  Every thread runs ILP independent dependent-FMA chains
// from registers; enough warps per SM to cover the pipe latency.  Prints TFLOP/s (2 flops per FMA) for a
// few (blocks/SM, ILP) shapes and the best one, plus the dependent-issue latency of DFMA.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/fp64_peak scripts/native/fp64_peak.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int ILP>
__global__ void __launch_bounds__(256) fma_kernel(double* out, double a, double b, int iters) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = 1.0 + 1e-3 * (threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void latency_kernel(double* out, long long* cyc, double a, double b, int iters) {
    double x = 1.0 + threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = fma(x, a, b);
    const long long t1 = clock64();
    if (threadIdx.x == 0) { *cyc = t1 - t0; out[0] = x; }
}

template <int ILP>
double run(int sms, int blocks_per_sm, int iters, double* out) {
    const int grid = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fma_kernel<ILP><<<grid, 256>>>(out, 0.999999, 1e-6, iters);
    cudaDeviceSynchronize();
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_kernel<ILP><<<grid, 256>>>(out, 0.999999, 1e-6, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * (double)grid * 256.0 * ILP * iters / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    return best;
}

int main() {
    int dev = 0, sms = 0, clk = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 148 * 16 * 256 * 2);
    cudaMalloc(&cyc, sizeof(long long));
    double best = 0.0; int bi = 0, bb = 0;
    const int iters = 1 << 16;
    const int bps[3] = {2, 4, 8};
    for (int b = 0; b < 3; ++b) {
        const double t4 = run<4>(sms, bps[b], iters, out);
        const double t8 = run<8>(sms, bps[b], iters, out);
        printf("blocks/SM %d (256 thr): ILP4 %.2f TFLOP/s  ILP8 %.2f TFLOP/s\n", bps[b], t4, t8);
        if (t4 > best) { best = t4; bi = 4; bb = bps[b]; }
        if (t8 > best) { best = t8; bi = 8; bb = bps[b]; }
    }
    latency_kernel<<<1, 32>>>(out, cyc, 0.999999, 1e-6, 4096);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.2f cycles\n", (double)h / 4096.0);
    printf("{\"fp64_tflops\": %.3f, \"sms\": %d, \"clock_khz_attr\": %d, \"ilp\": %d, \"blocks_per_sm\": %d, "
           "\"fma_per_clk_per_sm_at_attr_clock\": %.2f, \"dfma_latency_cycles\": %.2f}\n",
           best, sms, clk, bi, bb, best * 1e12 / 2.0 / sms / (clk * 1e3), (double)h / 4096.0);
    return 0;
}
