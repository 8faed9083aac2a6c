// exp_atoms.cu -- GPU experiment: throughput of 32-bit shared-memory atomic adds on sm_100 (what bounds the windows of
// the segmented image-gradient scatter, csrc/ssm_scatter.cuh): lane-operations per clock and SM for
//   consecutive addresses (a warp = 32 consecutive cells), a 2-way same-address pattern, random cells of an 80 x 32 x 3
//   window; fire-and-forget (RED-like) and with the returned value consumed.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/exp_atoms tools/exp_atoms.cu
#include <cuda_runtime.h>
#include <cstdio>
constexpr int CELLS = 80 * 32 * 3, ITERS = 4096;
template <int PATTERN, bool USE_RETURN>
__global__ void __launch_bounds__(256) k(int* out, unsigned seed) {
    __shared__ int win[CELLS];
    for (int i = threadIdx.x; i < CELLS; i += 256) win[i] = 0;
    __syncthreads();
    unsigned s = seed ^ (blockIdx.x * 256 + threadIdx.x) * 2654435761u;
    int acc = 0;
    for (int it = 0; it < ITERS; ++it) {
        int idx;
        if (PATTERN == 0) idx = (threadIdx.x + it * 256) % CELLS;                       // consecutive
        else if (PATTERN == 1) idx = ((threadIdx.x >> 1) + it * 128) % CELLS;            // pairs of lanes share a cell
        else { s = s * 1664525u + 1013904223u; idx = (s >> 8) % CELLS; }               // random
        const int old = atomicAdd(&win[idx], it | 1);
        if (USE_RETURN) acc |= old + 0x40000000;
    }
    __syncthreads();
    if (USE_RETURN || threadIdx.x == 0) out[blockIdx.x * 256 + threadIdx.x] = acc + win[threadIdx.x];
}
template <int PATTERN, bool USE_RETURN> void run(const char* name, int* d_out, int sms, double ghz) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 4;
    k<PATTERN, USE_RETURN><<<grid, 256>>>(d_out, 1u);
    cudaEventRecord(e0);
    k<PATTERN, USE_RETURN><<<grid, 256>>>(d_out, 2u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)grid * 256 * ITERS;
    printf("{\"pattern\": \"%s\", \"use_return\": %s, \"ms\": %.3f, \"G_lane_ops_per_s\": %.1f, \"lane_ops_per_clk_per_sm\": %.2f}\n", name,
           USE_RETURN ? "true" : "false", ms, ops / ms / 1e6, ops / (ms * 1e-3) / (sms * ghz * 1e9));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz / 1e6;
    int* d_out; cudaMalloc(&d_out, p.multiProcessorCount * 4 * 256 * sizeof(int));
    run<0, false>("consecutive", d_out, p.multiProcessorCount, ghz); run<0, true>("consecutive", d_out, p.multiProcessorCount, ghz);
    run<1, false>("pairs_share_a_cell", d_out, p.multiProcessorCount, ghz); run<1, true>("pairs_share_a_cell", d_out, p.multiProcessorCount, ghz);
    run<2, false>("random", d_out, p.multiProcessorCount, ghz); run<2, true>("random", d_out, p.multiProcessorCount, ghz);
    return 0;
}
