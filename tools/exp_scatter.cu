// exp_scatter.cu -- round-2 GPU experiment (not product code, not a bench number): what does the image-gradient
// scatter of the warp backward cost, and which accumulation scheme is worth building?
//
// Workload: P pairs of 1088x1920 frames, N = 7 timesteps, 3 channels, one frame per pair (the real kernels do two):
// every (pair, timestep, pixel) adds weight * grad to the 4 bilinear taps at pixel + c_n * flow, in 64-bit (or 32-bit)
// fixed point.  Flow field: control grid at 1/8 ("rough") or 1/64 ("smooth") resolution x 20 px,
// bilinearly upsampled, as bench.py's.  Variants:
//   planar64     destination planes of int64, lane = pixel, 12 RED.64 per pixel and timestep     (what ships)
//   planar32     the same with int32 accumulators                                                 (op width)
//   planarf32    fp32 atomicAdd, planar                                       (the reference's ATen/cuDNN scheme)
//   inter64      accumulators interleaved per pixel (4 x int64 = one 32-byte sector), lanes = (pixel, channel):
//                one RED instruction carries the 3 channels of a tap in ONE sector
//   smem32       CTA = 64x16 source pixels; int32 shared-memory window (tile + 8 px halo, shifted by the tile's mean
//                displacement) collects what lands inside, the rest goes to global; non-zero cells are flushed
//                with RED.64
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/exp_scatter tools/exp_scatter.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at %d\"}\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int H = 1088, W = 1920, N = 7, C = 3;
constexpr float SCALE = 16777216.0f;   // 2^24: |contribution| < 2^7 -> 31 bits

struct Tap { int off; float w[4]; bool in[4]; };

__device__ __forceinline__ Tap make_tap(int x, int y, float u, float v) {
    float ix = (float)x + u, iy = (float)y + v;
    ix = fminf(fmaxf(ix, -2.0f), (float)W + 1.0f);
    iy = fminf(fmaxf(iy, -2.0f), (float)H + 1.0f);
    const float fx = floorf(ix), fy = floorf(iy);
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
    const int x0 = (int)fx, y0 = (int)fy;
    Tap t;
    t.off = y0 * W + x0;
    t.w[0] = wx0 * wy0; t.w[1] = wx1 * wy0; t.w[2] = wx0 * wy1; t.w[3] = wx1 * wy1;
    const bool xi0 = (unsigned)x0 < (unsigned)W, xi1 = (unsigned)(x0 + 1) < (unsigned)W;
    const bool yi0 = (unsigned)y0 < (unsigned)H, yi1 = (unsigned)(y0 + 1) < (unsigned)H;
    t.in[0] = xi0 && yi0; t.in[1] = xi1 && yi0; t.in[2] = xi0 && yi1; t.in[3] = xi1 && yi1;
    return t;
}
__device__ __forceinline__ int tap_delta(int k) { return (k & 1) + (k >> 1) * W; }
__device__ __forceinline__ float coef(int n) { return 0.1f + 0.1f * (float)n; }

// ---- planar accumulators, lane = pixel ---------------------------------------------------------------------------
template <typename ACC>
__global__ void __launch_bounds__(256) k_planar(const float* __restrict__ flow, const float* __restrict__ grad, ACC* __restrict__ acc, int P) {
    const int tiles_x = W / 32, tiles_y = H / 8, tpp = tiles_x * tiles_y;
    const int b = blockIdx.x / tpp, r = blockIdx.x % tpp, ty = r / tiles_x, tx = r % tiles_x;
    const int x = tx * 32 + (threadIdx.x & 31), y = ty * 8 + (threadIdx.x >> 5);
    const long long npx = (long long)H * W;
    const int p = y * W + x;
    const float u = flow[(b * 2LL) * npx + p], v = flow[(b * 2LL + 1) * npx + p];
    ACC* a = acc + (long long)b * C * npx;
    for (int n = 0; n < N; ++n) {
        const Tap t = make_tap(x, y, coef(n) * u, coef(n) * v);
        const float* g = grad + ((long long)(b * N + n) * C) * npx + p;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float gc = __ldcs(g + c * npx);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!t.in[k]) continue;
                const float contrib = t.w[k] * gc;
                ACC* dst = a + c * npx + t.off + tap_delta(k);
                if constexpr (sizeof(ACC) == 8) {
                    const long long q = __float2ll_rn(contrib * SCALE);
                    if (q) atomicAdd((unsigned long long*)dst, (unsigned long long)q);
                } else if constexpr (sizeof(ACC) == 4 && !__is_same(ACC, float)) {
                    const int q = __float2int_rn(contrib * SCALE);
                    if (q) atomicAdd((int*)dst, q);
                } else {
                    atomicAdd((float*)dst, contrib);
                }
            }
        }
    }
}

// ---- interleaved accumulators [pixel][4] int64, lanes = (pixel, channel) -----------------------------------------
__global__ void __launch_bounds__(256) k_inter64(const float* __restrict__ flow, const float* __restrict__ grad, long long* __restrict__ acc, int P) {
    const int tiles_x = W / 32, tiles_y = H / 8, tpp = tiles_x * tiles_y;
    const int b = blockIdx.x / tpp, r = blockIdx.x % tpp, ty = r / tiles_x, tx = r % tiles_x;
    const int lane = threadIdx.x & 31;
    const int x = tx * 32 + lane, y = ty * 8 + (threadIdx.x >> 5);
    const long long npx = (long long)H * W;
    const int p = y * W + x;
    const float u = flow[(b * 2LL) * npx + p], v = flow[(b * 2LL + 1) * npx + p];
    long long* a = acc + (long long)b * 4 * npx;
    const int ch = lane & 3, sub = lane >> 2;          // in sub-iteration s this lane serves pixel 8s + sub, channel ch
    for (int n = 0; n < N; ++n) {
        const Tap t = make_tap(x, y, coef(n) * u, coef(n) * v);
        const float* g = grad + ((long long)(b * N + n) * C) * npx + p;
        const float g0 = __ldcs(g), g1 = __ldcs(g + npx), g2 = __ldcs(g + 2 * npx);
        const unsigned inmask = (t.in[0] ? 1u : 0u) | (t.in[1] ? 2u : 0u) | (t.in[2] ? 4u : 0u) | (t.in[3] ? 8u : 0u);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int src = 8 * s + sub;
            const int off = __shfl_sync(0xffffffffu, t.off, src);
            const unsigned m = __shfl_sync(0xffffffffu, inmask, src);
            const float a0 = __shfl_sync(0xffffffffu, g0, src), a1 = __shfl_sync(0xffffffffu, g1, src), a2 = __shfl_sync(0xffffffffu, g2, src);
            const float gc = ch == 0 ? a0 : (ch == 1 ? a1 : a2);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float wk = __shfl_sync(0xffffffffu, t.w[k], src);
                if (ch < 3 && ((m >> k) & 1u)) {
                    const long long q = __float2ll_rn(wk * gc * SCALE);
                    if (q) atomicAdd((unsigned long long*)(a + (long long)(off + tap_delta(k)) * 4 + ch), (unsigned long long)q);
                }
            }
        }
    }
}

// ---- shared-memory window ----------------------------------------------------------------------------------------
constexpr int STW = 64, STH = 16, HALO = 8, WW = STW + 2 * HALO, WH = STH + 2 * HALO;   // 80 x 32 cells
__global__ void __launch_bounds__(256) k_smem32(const float* __restrict__ flow, const float* __restrict__ grad, long long* __restrict__ acc, int P) {
    __shared__ int win[C][WH][WW];
    __shared__ float red[2][8];
    __shared__ int org[2];
    const int tiles_x = W / STW, tiles_y = H / STH, tpp = tiles_x * tiles_y;
    const int b = blockIdx.x / tpp, r = blockIdx.x % tpp, ty = r / tiles_x, tx = r % tiles_x;
    const long long npx = (long long)H * W;
    long long* a = acc + (long long)b * C * npx;
    // thread -> 4 pixels: column lx (0..63), rows ly, ly+4, ly+8, ly+12
    const int lx = threadIdx.x & 63, ly = threadIdx.x >> 6;
    float u[4], v[4];
    float su = 0.f, sv = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = (ty * STH + ly + 4 * j) * W + tx * STW + lx;
        u[j] = flow[(b * 2LL) * npx + p]; v[j] = flow[(b * 2LL + 1) * npx + p];
        su += u[j]; sv += v[j];
    }
    for (int o = 16; o; o >>= 1) { su += __shfl_xor_sync(0xffffffffu, su, o); sv += __shfl_xor_sync(0xffffffffu, sv, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = su; red[1][threadIdx.x >> 5] = sv; }
    __syncthreads();
    float mu = 0.f, mv = 0.f;
    for (int k = 0; k < 8; ++k) { mu += red[0][k]; mv += red[1][k]; }
    mu *= 1.0f / (STW * STH); mv *= 1.0f / (STW * STH);
    for (int n = 0; n < N; ++n) {
        for (int i = threadIdx.x; i < C * WH * WW; i += 256) (&win[0][0][0])[i] = 0;
        const int ox = tx * STW - HALO + (int)rintf(coef(n) * mu), oy = ty * STH - HALO + (int)rintf(coef(n) * mv);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = tx * STW + lx, y = ty * STH + ly + 4 * j;
            const Tap t = make_tap(x, y, coef(n) * u[j], coef(n) * v[j]);
            const float* g = grad + ((long long)(b * N + n) * C) * npx + y * W + x;
            const int x0 = t.off % W, y0 = t.off / W;     // valid whenever a tap is in
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float gc = __ldcs(g + c * npx);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!t.in[k]) continue;
                    const int q = __float2int_rn(t.w[k] * gc * SCALE);
                    if (!q) continue;
                    const int cx = x0 + (k & 1) - ox, cy = y0 + (k >> 1) - oy;
                    if ((unsigned)cx < (unsigned)WW && (unsigned)cy < (unsigned)WH) atomicAdd(&win[c][cy][cx], q);
                    else atomicAdd((unsigned long long*)(a + c * npx + t.off + tap_delta(k)), (unsigned long long)(long long)q);
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < C * WH * WW; i += 256) {
            const int q = (&win[0][0][0])[i];
            if (!q) continue;
            const int c = i / (WH * WW), rr = i % (WH * WW), cy = rr / WW, cx = rr % WW;
            const int gx = ox + cx, gy = oy + cy;
            if ((unsigned)gx < (unsigned)W && (unsigned)gy < (unsigned)H)
                atomicAdd((unsigned long long*)(a + c * npx + gy * W + gx), (unsigned long long)(long long)q);
        }
        __syncthreads();
    }
}

static float lcg_normal(unsigned long long& s) {
    auto next = [&]() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (float)((s >> 40) + 1) / 16777217.0f; };
    const float a = next(), b = next();
    return sqrtf(-2.0f * logf(a)) * cosf(6.2831853f * b);
}

static void make_flow(std::vector<float>& f, int P, int div) {
    const int gh = H / div + 2, gw = W / div + 2;
    unsigned long long s = 12345;
    std::vector<float> grid((size_t)P * 2 * gh * gw);
    for (auto& g : grid) g = 20.0f * lcg_normal(s);
    for (int bc = 0; bc < P * 2; ++bc)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const float gy = (float)y / div, gx = (float)x / div;
                const int y0 = (int)gy, x0 = (int)gx;
                const float wy = gy - y0, wx = gx - x0;
                const float* g = &grid[(size_t)bc * gh * gw];
                f[((size_t)bc * H + y) * W + x] = (1 - wy) * ((1 - wx) * g[y0 * gw + x0] + wx * g[y0 * gw + x0 + 1]) +
                                                  wy * ((1 - wx) * g[(y0 + 1) * gw + x0] + wx * g[(y0 + 1) * gw + x0 + 1]);
            }
}

template <typename F> static float time_ms(F launch, void* acc, size_t acc_bytes, int reps = 3) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int i = 0; i < reps + 1; ++i) {
        CK(cudaMemset(acc, 0, acc_bytes));
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (i > 0 && ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    const int P = argc > 1 ? atoi(argv[1]) : 4;
    const long long npx = (long long)H * W;
    float *d_flow, *d_grad; void* d_acc;
    const size_t acc_bytes = (size_t)P * 4 * npx * 8;
    CK(cudaMalloc(&d_flow, (size_t)P * 2 * npx * 4));
    CK(cudaMalloc(&d_grad, (size_t)P * N * C * npx * 4));
    CK(cudaMalloc(&d_acc, acc_bytes));
    {   // upstream gradient: N(0,1)
        std::vector<float> g((size_t)P * N * C * npx);
        unsigned long long s = 777;
        for (auto& x : g) x = lcg_normal(s);
        CK(cudaMemcpy(d_grad, g.data(), g.size() * 4, cudaMemcpyHostToDevice));
    }
    std::vector<float> f((size_t)P * 2 * npx);
    const double contribs = (double)P * N * npx * C * 4;
    for (int div : {8, 64}) {
        make_flow(f, P, div);
        CK(cudaMemcpy(d_flow, f.data(), f.size() * 4, cudaMemcpyHostToDevice));
        const char* field = div == 8 ? "rough" : "smooth";
        const int grid = P * (W / 32) * (H / 8);
        auto report = [&](const char* name, float ms) {
            printf("{\"field\": \"%s\", \"variant\": \"%s\", \"pairs\": %d, \"ms\": %.3f, \"ms_at_16_pairs_2_frames\": %.2f, \"G_contributions_per_s\": %.1f}\n",
                   field, name, P, ms, ms * 16.0 / P * 2.0, contribs / ms / 1e6);
            fflush(stdout);
        };
        report("planar64", time_ms([&] { k_planar<long long><<<grid, 256>>>(d_flow, d_grad, (long long*)d_acc, P); }, d_acc, acc_bytes));
        report("planar32", time_ms([&] { k_planar<int><<<grid, 256>>>(d_flow, d_grad, (int*)d_acc, P); }, d_acc, acc_bytes));
        report("planarf32", time_ms([&] { k_planar<float><<<grid, 256>>>(d_flow, d_grad, (float*)d_acc, P); }, d_acc, acc_bytes));
        report("inter64", time_ms([&] { k_inter64<<<grid, 256>>>(d_flow, d_grad, (long long*)d_acc, P); }, d_acc, acc_bytes));
        report("smem32", time_ms([&] { k_smem32<<<P * (W / STW) * (H / STH), 256>>>(d_flow, d_grad, (long long*)d_acc, P); }, d_acc, acc_bytes));
    }
    return 0;
}
